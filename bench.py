#!/usr/bin/env python
"""bench.py -- CIFAR-10 DDPM samples/sec (1000 steps) on N B200s; see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

A "step" is one pass of the hot path over one batch: timestep embedding -> UNet forward (bf16 tensor-core
path) -> fused DDPM ancestral update, for `--batch` (default 256) synthetic 3x32x32 images per GPU,
random-init default UNet (seed 0).  Samples/sec (1000 steps) = images / (1000 * seconds per step).
Each rank owns its own shard of the sample batch (no data-path collective): weak scaling.

Printed (rank 0, one JSON line): value (device-resident, graph replays, CUDA events, max over ranks),
e2e (through DDPM.sampling-step graph with pinned-host x_t in / x_{t-1} out every step), roofline of the
dominant kernel (tcgen05 implicit-GEMM conv), cpu_baseline (the oracle port on the host cores), clocks.
`--impl reference` times the reference's CPU path (oracle port, torch fp32 on all host cores).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "diffusion-models-made-easy_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "CIFAR-10 DDPM samples/sec (1000 steps)"
UNIT = "samples/s"
TIMESTEPS = 1000
IMG = (3, 32, 32)
FLOP_PER_IMAGE = 9_803_923_456  # SURVEY.md par. 8d: default DDPM UNet forward, 2*MAC, 32x32


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_step_rate(batch: int, reps: int, warmup: int):
    """image-steps per second of DDPM.sampling_step on the CPU (oracle port, fp32, all cores)."""
    import dmme_oracle as O
    from dmme_b200.models.ddpm import UNet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = UNet().eval().state_dict()
    tabs = O.linear_tables(TIMESTEPS)
    torch.manual_seed(1234)
    x = torch.randn(256, *IMG)[:batch].contiguous()
    z = torch.randn(batch, *IMG)
    t = torch.tensor([500])
    times = []
    with torch.no_grad():
        for k in range(warmup + reps):
            t0 = time.perf_counter()
            eps = O.unet_forward(sd, x, t)
            O.ddpm_step(x, t, eps, z, tabs)
            dt = time.perf_counter() - t0
            if k >= warmup:
                times.append(dt)
    best = min(times)
    mean = sum(times) / len(times)
    return batch / mean, batch / best, cores, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 16
    rate_mean, rate_best, cores, times = cpu_step_rate(batch, max(1, args.steps), max(1, args.warmup))
    ms = 1e3 * sum(times) / len(times)
    value = rate_mean / TIMESTEPS
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DDPM CIFAR-10 UNet (configs/ddpm/cifar10.yaml) 1000-step ancestral sampling; "
                               "CPU sample: one denoise step on 16 images, extrapolated to samples/s at 1000 steps",
                   "batch_per_step": batch, "image": list(IMG), "timesteps": TIMESTEPS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} timed DDPM.sampling_step calls (UNet fwd + update), batch {batch}, fp32, "
                                   f"{cores} torch threads; samples/s = image-steps/s / 1000"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def conv_profile(model, x, t, reps=3):
    """Per-launch CUDA-event timing of every tcgen05 conv launch of one UNet forward (eager, same stream), grouped by
    launch signature.  Returns (groups, totals): groups[sig] = [launches, flop, ms] of the best repetition."""
    from dmme_b200 import ops
    from dmme_b200.models import _engine
    records = []
    orig = ops.conv2d_launch

    def timed(desc, weight, bias, out, temb=None, addend=None, out2=None, out3=None, stats=None, **kw):
        tc = ops.conv_uses_tc(desc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(desc, weight, bias, out, temb, addend, out2, out3, stats, **kw)
        e1.record()
        ho, wo = ops.conv_out_hw(desc)
        k = desc.ksize * desc.ksize * (desc.c0 + desc.c1) + desc.rc0 + desc.rc1
        sig = (f"{desc.ksize}x{desc.ksize} s{desc.stride} {desc.c0 + desc.c1}->{desc.cout} @{desc.h_in}x{desc.w_in}"
               + (f" +res{desc.rc0 + desc.rc1}" if desc.rc0 + desc.rc1 else "")
               + (" +GroupNorm+SiLU of the input" if kw.get("gn_ab") is not None else ""))
        records.append((tc, sig, 2.0 * desc.n * ho * wo * desc.cout * k, e0, e1))

    _engine.ops.conv2d_launch = timed
    try:
        best = None
        for _ in range(reps):
            records.clear()
            model.forward_raw(x, t)
            torch.cuda.synchronize()
            groups = {}
            for tc, sig, flop, a, b in records:
                if not tc:
                    continue
                g = groups.setdefault(sig, [0, 0.0, 0.0])
                g[0] += 1
                g[1] += flop
                g[2] += a.elapsed_time(b)
            tot = (sum(g[0] for g in groups.values()), sum(g[1] for g in groups.values()), sum(g[2] for g in groups.values()))
            if best is None or tot[2] < best[1][2]:
                best = (groups, tot)
    finally:
        _engine.ops.conv2d_launch = orig
    return best


def gn_profile(model, x, t, reps=3):
    """Per-launch CUDA-event timing of every GroupNorm(+SiLU) launch of one UNet forward, grouped by launch signature;
    algorithmic bytes = 2 B read + 2 B written per element (SURVEY par. 8d).  Returns (groups, totals) of the best
    repetition: groups[sig] = [launches, bytes, ms], totals = (launches, bytes, ms)."""
    from dmme_b200 import ops
    from dmme_b200.models import _engine
    records = []
    orig = ops.groupnorm

    def timed(src0, src1, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(src0, src1, *a, **k)
        e1.record()
        elems = src0.numel() + (src1.numel() if src1 is not None else 0)
        c = src0.shape[3] + (src1.shape[3] if src1 is not None else 0)
        records.append((f"GroupNorm+SiLU {c} ch @{src0.shape[1]}x{src0.shape[2]}", 4.0 * elems, e0, e1))
        return out

    _engine.ops.groupnorm = timed
    try:
        best = None
        for _ in range(reps):
            records.clear()
            model.forward_raw(x, t)
            torch.cuda.synchronize()
            groups = {}
            for sig, byts, a, b in records:
                g = groups.setdefault(sig, [0, 0.0, 0.0])
                g[0] += 1
                g[1] += byts
                g[2] += a.elapsed_time(b)
            tot = (len(records), sum(g[1] for g in groups.values()), sum(g[2] for g in groups.values()))
            if best is None or tot[2] < best[1][2]:
                best = (groups, tot)
    finally:
        _engine.ops.groupnorm = orig
    return best


def traffic_for(sig):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json,
    written from profiles/*_ncu_full.txt); None when that launch signature was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path)).get(sig, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def run_gpu(args):
    import torch.distributed as dist
    from dmme_b200 import DDPM, ops
    from dmme_b200.models.ddpm import UNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    torch.manual_seed(0)
    model = UNet().eval()
    ddpm = DDPM(model, TIMESTEPS).to(dev)
    # this rank's shard of the sample batch: images [rank*B, (rank+1)*B) of the seeded global x_T
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, *IMG, generator=g).pin_memory()
    x = x_host.to(dev, non_blocking=True)
    counter = torch.full((1,), TIMESTEPS, dtype=torch.int64, device=dev)
    seed = 20261018 + rank

    # ---- warm-up (eager: packs weights, sizes workspace) + launch count per step ----
    ddpm._graph_step(x, counter, seed)  # first call: also packs the weights (one pack launch per conv site)
    torch.cuda.synchronize()
    ops.reset_launch_count()
    ddpm._graph_step(x, counter, seed)
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count()
    ddpm._graph_step(x, counter, seed)
    torch.cuda.synchronize()

    # ---- capture one step ----
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ddpm._graph_step(x, counter, seed)

    def reset_state():
        x.copy_(x_host, non_blocking=True)
        counter.fill_(TIMESTEPS)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- device-resident timing ----
    reset_state()
    for _ in range(args.warmup):
        graph.replay()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        graph.replay()
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    finite = bool(torch.isfinite(x).all())

    # ---- end to end: pinned-host x_t in, x_{t-1} out, every step ----
    out_host = torch.empty_like(x_host).pin_memory()
    reset_state()
    for _ in range(args.warmup):
        x.copy_(x_host, non_blocking=True)
        graph.replay()
        out_host.copy_(x, non_blocking=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        x.copy_(x_host, non_blocking=True)
        graph.replay()
        out_host.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads x_{t-1} on the host
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    clk = clocks.stop() if rank == 0 else None

    if world > 1:
        # the one collective of sharded sampling: final gather of the samples (outside the timed region)
        gathered = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
        dist.gather(x, gathered, dst=0)

    if rank == 0:
        pk = peaks()
        # roofline of the dominant kernel: tcgen05 implicit-GEMM conv, per-launch events, eager pass
        reset_state()
        groups, (n_tc, tc_flop, tc_ms) = conv_profile(model, x, counter)
        # dominant kernel = the launch signature with the largest share of the step
        dom_sig, (dom_n, dom_flop, dom_ms) = max(groups.items(), key=lambda kv: kv[1][2])
        achieved = dom_flop / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        all_tc = tc_flop / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        gn_groups, (gn_n, gn_bytes, gn_ms) = gn_profile(model, x, counter)
        gn_gbs = gn_bytes / (gn_ms * 1e-3) / 1e9 if gn_ms > 0 else 0.0
        # dominant GroupNorm launch signature (largest share of the pass) -- the HBM-bound kernel of the step
        # the GroupNorm launch signature that actually streams through HBM: the largest tensor (the others fit the 126 MB L2
        # and are latency-bound launches of a few MB; most GroupNorms of the step now run inside the halo conv kernel)
        gd_sig, (gd_n, gd_bytes, gd_ms) = max(gn_groups.items(), key=lambda kv: kv[1][1] / kv[1][0])
        gd_gbs = gd_bytes / (gd_ms * 1e-3) / 1e9 if gd_ms > 0 else 0.0
        step_flop = FLOP_PER_IMAGE * B
        cpu = None
        if not args.no_cpu:
            rate_mean, _, cores, times = cpu_step_rate(16, 5, 2)
            cpu = {"value": rate_mean / TIMESTEPS, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{len(times)} timed oracle DDPM.sampling_step calls, batch 16, fp32, {cores} torch threads; "
                             "samples/s = image-steps/s / 1000"}
        total = B * world
        line = {
            "metric": METRIC, "value": total / ms_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "DDPM CIFAR-10 1000-step ancestral sampling, 256 images per GPU, bf16 "
                                   "(BASELINE config #2; default UNet 32.4M params, seed-0 random init)",
                       "batch_per_gpu": B, "global_batch": total, "image": list(IMG), "timesteps": TIMESTEPS,
                       "parallelism": f"sample-batch shard x{world}, no data-path collective",
                       "l2": "per-step activation working set (>1 GB) exceeds the 126 MB L2; no explicit flush",
                       "step": "CUDA graph: temb -> UNet fwd -> fused DDPM update -> t -= 1"},
            "e2e": {"value": total / ms_e2e, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "roofline": {"bound": "tensor", "kernel": f"tcgen05 implicit-GEMM conv, launch {dom_sig} x{B} images "
                                                      f"({dom_n} launches per step, largest share of the step)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic_for(dom_sig), "launches_per_step": dom_n,
                         "flop_per_launch": dom_flop / dom_n, "us_per_launch": 1e3 * dom_ms / dom_n,
                         "peak_source": pk["source"] + " sustained bf16 (kernel timed inside a long step)",
                         "all_tensor_core_convs": {"launches_per_step": n_tc, "flop_per_step": tc_flop, "ms_per_step": tc_ms,
                                                   "achieved": all_tc, "frac": all_tc / peak},
                         "step_tensor_frac": step_flop / (ms_dev * 1e-3) / 1e12 / peak},
            "roofline_hbm": {"bound": "hbm", "kernel": f"gn_apply_kernel, launch {gd_sig} x{B} images ({gd_n} launches per step, "
                                                       "the largest stand-alone GroupNorm tensor; statistics come from "
                                                       "the producing conv's epilogue)",
                             "achieved": gd_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gd_gbs / pk["hbm_gbs"],
                             "traffic": traffic_for(gd_sig), "launches_per_step": gd_n,
                             "bytes_per_launch": gd_bytes / gd_n, "us_per_launch": 1e3 * gd_ms / gd_n,
                             "peak_source": pk["source"] + " STREAM-style copy",
                             "all_groupnorm_launches": {"launches_per_step": gn_n, "bytes_per_step": gn_bytes,
                                                        "ms_per_step": gn_ms, "achieved": gn_gbs,
                                                        "frac": gn_gbs / pk["hbm_gbs"],
                                                        "note": "per-launch event pairs add ~2 us to each of the small "
                                                                "8x8 / 4x4 launches"}},
            "cpu_baseline": cpu, "clocks": clk, "finite": finite,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dmme_b200", choices=["dmme_b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 20)
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the dmme_b200 arm has no CPU path (use --impl reference)")
    args.warmup = max(3, args.warmup)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: the driver launches torchrun itself; a bare `python bench.py --gpus N` re-launches under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"),
               os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == "__main__":
    main()
