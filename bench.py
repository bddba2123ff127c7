#!/usr/bin/env python
"""bench.py -- CIFAR-10 DDPM samples/sec (1000 steps) on N B200s; see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak] [--no-cpu] [--no-train]

BASELINE config #2 as written: 256 images of 3x32x32 noise (seed 1234), default DDPM UNet (seed-0 random init), 1000-step
ancestral sampling, the sample batch sharded over the N GPUs -- rank r owns images [r*256/N, (r+1)*256/N) (STRONG
scaling, the default; no data-path collective, one gather of the finished samples at the end).  A "step" is one pass of
the hot path over this rank's shard: timestep embedding -> UNet forward (bf16 tensor-core path) -> fused DDPM ancestral
update -> t -= 1, captured once as a CUDA graph and replayed.  samples/s = 256 / (1000 * seconds per step).

Rank 0 prints ONE JSON line:
  value / ms_per_step     device-resident graph replays, CUDA events, max over ranks
  e2e                     the same step with x_t copied in from pinned host memory and x_{t-1} copied back every step
  weak_scaling (N > 1)    256 images PER GPU (global batch 256 N), same measurement
  batch_sweep  (N = 1)    ms/step at per-GPU batches 256/128/64/32 = what each rank runs at N = 1/2/4/8
  roofline / roofline_hbm / roofline_sampler   dominant tcgen05 conv, largest stand-alone GroupNorm, sampler update
  parity                  rel-L2 of the timed configuration's UNet output against the CPU oracle (16 images of this shard)
  train                   BASELINE config #4: graph-captured IDDPM hybrid-loss training step, batch 128 per GPU, bf16,
                          FusedAdamEMA, gradient all-reduce (NCCL, captured in the graph) when N > 1
  cpu_baseline (N = 1)    the unmodified reference (baseline/_ref) -- or the oracle port when it is not installed -- on the
                          host cores
`--impl reference` times the reference's own CPU implementation of the step (same metric / config keys).
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
GLOBAL_BATCH = 256
FLOP_PER_IMAGE = 9_803_923_456         # SURVEY.md par. 8d: default DDPM UNet forward, 2*MAC, 32x32
FLOP_PER_IMAGE_IDDPM = 9_999_745_024   # IDDPM UNet forward; forward + backward = 3x
SAMPLER_BYTES_PER_IMAGE = 49_152       # x_t read + write, eps, z: 3072 elements x 16 B (SURVEY par. 8d)
TRAIN_BATCH = 128                      # configs/iddpm/cifar10.yaml:86
WORKLOAD = ("DDPM CIFAR-10 1000-step ancestral sampling, 256 images, bf16, sample batch sharded over the GPUs "
            "(BASELINE config #2; default UNet 32.4M params, seed-0 random init, x_T from seed 1234)")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (baseline/_ref) on the host cores; oracle port when it is not installed
# ----------------------------------------------------------------------------------------------
def cpu_step_rate(batch: int, reps: int, warmup: int):
    """image-steps per second of DDPM.sampling_step on the CPU, fp32, all cores.  Returns (rate_mean, rate_best, cores,
    times, kind): kind "reference" = the unmodified dmme.diffusion_models.DDPM.sampling_step with dmme.models.ddpm.UNet
    (imported from baseline/_ref through oracle/ref_shim.py), "port" = oracle/dmme_oracle.py."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    x = torch.randn(GLOBAL_BATCH, *IMG)[:batch].contiguous()
    t = torch.tensor([500])
    kind = "port"
    step = None
    try:
        import ref_shim
        if ref_shim.available():
            ref_shim.load()
            from dmme.diffusion_models import DDPM as RefDDPM
            from dmme.models.ddpm import UNet as RefUNet
            torch.manual_seed(0)
            ref = RefDDPM(RefUNet().eval(), TIMESTEPS).eval()
            step = lambda: ref.sampling_step(x, t)  # noqa: E731
            kind = "reference"
    except Exception as e:  # an unusable install must not take the bench down: fall back to the port and say so
        print(f"bench.py: reference import failed ({e!r}); timing the oracle port", file=sys.stderr)
        step = None
    if step is None:
        import dmme_oracle as O
        from dmme_b200.models.ddpm import UNet
        torch.manual_seed(0)
        sd = UNet().eval().state_dict()
        tabs = O.linear_tables(TIMESTEPS)
        z = torch.randn(batch, *IMG)
        step = lambda: O.ddpm_step(x, t, O.unet_forward(sd, x, t), z, tabs)  # noqa: E731
    times = []
    with torch.no_grad():
        for k in range(warmup + reps):
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            if k >= warmup:
                times.append(dt)
    best = min(times)
    mean = sum(times) / len(times)
    return batch / mean, batch / best, cores, times, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 16
    rate_mean, rate_best, cores, times, kind = cpu_step_rate(batch, max(1, args.steps), max(1, args.warmup))
    ms = 1e3 * sum(times) / len(times)
    value = rate_mean / TIMESTEPS
    what = ("the unmodified reference (dmme.diffusion_models.DDPM.sampling_step, dmme.models.ddpm.UNet from baseline/_ref)"
            if kind == "reference" else "the oracle port of the reference's algorithm")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + "; CPU arm: one denoise step on a bounded sample of 16 of the 256 images per "
                               "timed step, converted to samples/s at 1000 steps",
                   "global_batch": GLOBAL_BATCH, "batch_per_step": batch, "image": list(IMG), "timesteps": TIMESTEPS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{len(times)} timed DDPM.sampling_step calls (UNet fwd + update) of {what}, batch {batch}, "
                                   f"fp32, {cores} torch threads; samples/s = image-steps/s / 1000"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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
    """Per-launch CUDA-event timing of every stand-alone GroupNorm(+SiLU) launch of one UNet forward, grouped by launch
    signature; algorithmic bytes = 2 B read + 2 B written per element (SURVEY par. 8d).  Returns (groups, totals) of the
    best repetition: groups[sig] = [launches, bytes, ms], totals = (launches, bytes, ms); (None, (0, 0, 0)) when every
    GroupNorm of the step runs inside a conv kernel."""
    from dmme_b200 import ops
    from dmme_b200.models import _engine
    records = []
    calls = {}  # signature -> arguments of its first launch (gn_isolated re-times the dominant one alone)
    orig = ops.groupnorm

    def timed(src0, src1, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(src0, src1, *a, **k)
        e1.record()
        elems = src0.numel() + (src1.numel() if src1 is not None else 0)
        c = src0.shape[3] + (src1.shape[3] if src1 is not None else 0)
        sig = f"GroupNorm+SiLU {c} ch @{src0.shape[1]}x{src0.shape[2]}"
        records.append((sig, 4.0 * elems, e0, e1))
        calls.setdefault(sig, (src0, src1, a, k, out))
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
    if best is None or not best[0]:
        return None, (0, 0.0, 0.0), {}
    return best[0], best[1], calls


def gn_isolated(call, rounds=4):
    """One stand-alone GroupNorm launch timed alone: back to back over ROTATING copies of its tensors whose footprint exceeds
    twice the 126 MB L2 (every launch reads its input from HBM and its output cannot stay resident), one CUDA-event pair
    around the whole train of launches on the launching stream.  Inside the eager forward a per-launch event pair around a
    10 us kernel mostly measures the host's launch latency (the GPU idles between the event and the launch), which is why
    the dominant launch is re-timed this way.  Returns (us per launch, launches timed, buffer sets)."""
    from dmme_b200 import ops
    src0, src1, a, k, out = call
    a = list(a)
    per = 2 * (src0.numel() + (src1.numel() if src1 is not None else 0)) * src0.element_size()  # read + written
    sets = max(2, int(2 * 126e6 // per) + 1)
    s0 = [src0.clone() for _ in range(sets)]
    s1 = [src1.clone() if src1 is not None else None for _ in range(sets)]
    outs = [torch.empty_like(out) for _ in range(sets)]
    has_out_kw = "out" in k

    def launch(i):
        if has_out_kw:
            ops.groupnorm(s0[i], s1[i], *a, **{**k, "out": outs[i]})
        else:
            aa = list(a)
            aa[8] = outs[i]  # positional: groups, gamma, beta, silu, scale, shift, chan_mask, eps, out
            ops.groupnorm(s0[i], s1[i], *aa, **k)

    for i in range(sets):
        launch(i)
    torch.cuda.synchronize()
    # the train of launches is captured once and replayed (as the step itself is): issued from Python one by one, a 10 us
    # kernel would be timed at the host's launch rate
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rounds):
            for i in range(sets):
                launch(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (rounds * sets), rounds * sets, sets


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


class SamplingRun:
    """One rank's shard of the sampling job: state, captured step graph, device-resident and end-to-end timing."""

    def __init__(self, ddpm, x_host, dev, seed, noise_offset):
        from dmme_b200 import ops
        self.ddpm, self.dev, self.seed = ddpm, dev, seed
        self.x_host = x_host.pin_memory()
        self.out_host = torch.empty_like(self.x_host).pin_memory()
        self.x = self.x_host.to(dev, non_blocking=True)
        self.counter = torch.full((1,), TIMESTEPS, dtype=torch.int64, device=dev)
        ddpm._noise_offset = noise_offset  # the noise this shard draws = what the single-GPU run draws for these images
        ddpm._graph_step(self.x, self.counter, seed)  # first call: packs the weights, sizes the workspace
        torch.cuda.synchronize()
        ops.reset_launch_count()
        ddpm._graph_step(self.x, self.counter, seed)
        torch.cuda.synchronize()
        self.launches_per_step = ops.launch_count()
        ddpm._graph_step(self.x, self.counter, seed)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            ddpm._graph_step(self.x, self.counter, seed)

    def reset(self):
        self.x.copy_(self.x_host, non_blocking=True)
        self.counter.fill_(TIMESTEPS)

    def time_device(self, steps, warmup, barrier):
        self.reset()
        for _ in range(warmup):
            self.graph.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.graph.replay()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps

    def time_e2e(self, steps, warmup, barrier):
        self.reset()
        for _ in range(warmup):
            self.x.copy_(self.x_host, non_blocking=True)
            self.graph.replay()
            self.out_host.copy_(self.x, non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.x.copy_(self.x_host, non_blocking=True)
            self.graph.replay()
            self.out_host.copy_(self.x, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller reads x_{t-1} on the host
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps


def sampler_roofline(ddpm, x, counter, seed, pk, reps=50):
    """ddpm_step_kernel alone: algorithmic 49,152 B per image (x_t read + write, eps, z) over its average launch time,
    timed back to back with CUDA events (the 12.6 MB working set of batch 256 stays in the 126 MB L2 between launches, as
    it does inside the step, where eps was just written by the output conv)."""
    eps = torch.randn_like(x)
    counter.fill_(500)
    for _ in range(5):
        ddpm._update_(x, eps, None, counter, seed)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ddpm._update_(x, eps, None, counter, seed)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    byts = SAMPLER_BYTES_PER_IMAGE * x.shape[0]
    gbs = byts / (us * 1e-6) / 1e9
    return {"bound": "hbm", "kernel": f"ddpm_step_kernel (fused ancestral update + in-kernel Philox noise) x{x.shape[0]} images",
            "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": None,
            "bytes_per_launch": byts, "us_per_launch": us,
            "note": "latency-bound: 12.6 MB per launch at batch 256 is ~2 us at the HBM peak; working set L2-resident"}


def parity_check(model, x, dev):
    """rel-L2 of the timed configuration's UNet output (this shard's batch, t = 1000) against the CPU oracle on 16 of its
    images (first and last 8; DDPM's UNet has no cross-sample coupling)."""
    import dmme_oracle as O
    n = x.shape[0]
    pick = sorted(set(list(range(min(8, n))) + list(range(max(0, n - 8), n))))
    t = torch.tensor([TIMESTEPS])
    got = model.forward_raw(x, t.to(dev)).float().cpu()[pick]
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = O.unet_forward(sd, x.cpu()[pick], t)
    err = float((got.double() - want.double()).norm() / want.double().norm())
    return {"unet_rel_l2_vs_oracle": err, "images": len(pick), "batch": n, "tolerance": 1e-2, "ok": err < 1e-2}


def train_bench(dev, world, rank, steps, warmup, barrier, max_over_ranks, pk):
    """BASELINE config #4: IDDPM default UNet (36.2 M parameters), cosine schedule, hybrid loss, bf16 tensor-core path,
    batch 128 per GPU, dropout 0.3, FusedAdamEMA (clip + Adam + WarmupLR + EMA).  Forward + loss + backward (+ the bucketed
    NCCL gradient all-reduce when N > 1) are ONE captured CUDA graph; the optimizer's two launches follow each replay."""
    import torch.distributed as dist
    from dmme_b200 import IDDPM
    from dmme_b200.models import iddpm as iddpm_models
    from dmme_b200.optim import FusedAdamEMA
    from dmme_b200.parallel import broadcast_parameters, enable_gradient_sync
    from dmme_b200.training import GraphedTrainingStep

    torch.manual_seed(0)
    model = iddpm_models.UNet()
    dm = IDDPM(model, TIMESTEPS).to(dev).train()
    nparams = sum(p.numel() for p in model.parameters())
    g = torch.Generator().manual_seed(3 + rank)
    x0 = (torch.rand(TRAIN_BATCH, *IMG, generator=g) * 2 - 1).to(dev)

    def measure(sync: bool):
        torch.manual_seed(17 + rank)
        if sync:
            broadcast_parameters(model)
            enable_gradient_sync(model)
        else:
            model.train_engine.grad_sync = None
        opt = FusedAdamEMA(model.parameters(), lr=2e-4, warmup=5000, max_grad_norm=1.0, ema_decay=0.9999)
        step = GraphedTrainingStep(dm, opt)
        loss = None
        for _ in range(warmup):
            loss = step(x0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step(x0)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / steps), float(loss.detach())

    ms_local, loss = measure(False)
    out = {"workload": "IDDPM cosine-schedule hybrid-loss training step (BASELINE config #4): default IDDPM UNet, batch 128 per "
                       "GPU, bf16, dropout 0.3, FusedAdamEMA; forward + loss + backward captured in one CUDA graph",
           "batch_per_gpu": TRAIN_BATCH, "global_batch": TRAIN_BATCH * world, "params": nparams, "steps": steps}
    ms = ms_local
    if world > 1:
        ms, loss = measure(True)
        out["allreduce"] = {"bytes_per_step": 4 * nparams, "dtype": "f32", "buckets_mb": 32,
                            "how": "NCCL all-reduce of the flat gradient arena in 32 MB buckets, launched as the backward "
                                   "pass moves past each bucket and captured inside the step's CUDA graph",
                            "ms_per_step_without_sync": ms_local, "exposed_ms": max(0.0, ms - ms_local)}
    flop = 3.0 * FLOP_PER_IMAGE_IDDPM * TRAIN_BATCH
    tf = flop / (ms * 1e-3) / 1e12
    out.update({"ms_per_step": ms, "value": TRAIN_BATCH * world / (ms * 1e-3), "unit": "images/s", "loss": loss,
                "tflops_per_gpu": tf, "frac_of_burst_bf16": tf / pk["bf16_tflops"],
                "flop_per_image": 3 * FLOP_PER_IMAGE_IDDPM})
    return out


def run_gpu(args):
    import torch.distributed as dist
    from dmme_b200 import DDPM
    from dmme_b200.models.ddpm import UNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

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

    strong = args.scaling == "strong"
    if strong and GLOBAL_BATCH % world:
        raise SystemExit(f"bench.py: {GLOBAL_BATCH} images do not split evenly over {world} GPUs")
    B = args.batch or (GLOBAL_BATCH // world if strong else GLOBAL_BATCH)
    total = B * world

    torch.manual_seed(0)
    model = UNet().eval()
    ddpm = DDPM(model, TIMESTEPS).to(dev)
    # the seeded global x_T (SURVEY par. 8d); this rank's shard = images [rank * B, (rank + 1) * B)
    torch.manual_seed(1234)
    x_global = torch.randn(GLOBAL_BATCH, *IMG)
    seed = 20261018

    def shard(batch, r):
        if batch * world <= GLOBAL_BATCH:
            return x_global[r * batch:(r + 1) * batch].clone()
        gg = torch.Generator().manual_seed(1234 + r)  # weak scaling: more than 256 images in total
        return torch.randn(batch, *IMG, generator=gg)

    per_image = IMG[0] * IMG[1] * IMG[2]
    run = SamplingRun(ddpm, shard(B, rank), dev, seed, rank * B * per_image)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_dev = max_over_ranks(run.time_device(args.steps, args.warmup, barrier))
    finite = bool(torch.isfinite(run.x).all())
    ms_e2e = max_over_ranks(run.time_e2e(args.steps, args.warmup, barrier))
    clk = clocks.stop() if rank == 0 else None

    if world > 1:
        # the one collective of sharded sampling: final gather of the samples (outside the timed region)
        gathered = [torch.empty_like(run.x) for _ in range(world)] if rank == 0 else None
        dist.gather(run.x, gathered, dst=0)

    # ---- the other scaling mode beside it: 256 images PER GPU when N > 1 ----
    weak = None
    if world > 1 and strong and not args.batch:
        wrun = SamplingRun(ddpm, shard(GLOBAL_BATCH, rank), dev, seed, rank * GLOBAL_BATCH * per_image)
        w_ms = max_over_ranks(wrun.time_device(args.steps, args.warmup, barrier))
        weak = {"scaling": "weak", "batch_per_gpu": GLOBAL_BATCH, "global_batch": GLOBAL_BATCH * world,
                "value": GLOBAL_BATCH * world / w_ms, "unit": UNIT, "ms_per_step": w_ms,
                "note": "256 images PER GPU (not the BASELINE config): per-GPU work fixed as N grows"}
        del wrun

    # ---- what each rank runs at N = 1 / 2 / 4 / 8: ms per step at per-GPU batches 256 / 128 / 64 / 32 (N = 1 only) ----
    sweep = None
    if world == 1 and not args.batch and not args.no_sweep:
        sweep = {}
        for b in (256, 128, 64, 32):
            if b == B:
                sweep[str(b)] = ms_dev
                continue
            srun = SamplingRun(ddpm, shard(b, 0), dev, seed, 0)
            sweep[str(b)] = srun.time_device(max(10, args.steps // 2), args.warmup, barrier)
            del srun

    train = None
    if not args.no_train:
        try:
            train = train_bench(dev, world, rank, max(5, min(args.steps, 20)), max(3, min(args.warmup, 5)), barrier,
                                max_over_ranks, peaks())
        except Exception as e:  # the headline metric must still be printed; the failure is part of the record
            if world > 1:
                raise
            train = {"error": repr(e)}

    if rank == 0:
        pk = peaks()
        # roofline of the dominant kernel: tcgen05 implicit-GEMM conv, per-launch events, eager pass
        run.reset()
        ddpm._noise_offset = rank * B * per_image
        groups, (n_tc, tc_flop, tc_ms) = conv_profile(model, run.x, run.counter)
        # dominant kernel = the launch signature with the largest share of the step
        dom_sig, (dom_n, dom_flop, dom_ms) = max(groups.items(), key=lambda kv: kv[1][2])
        achieved = dom_flop / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        all_tc = tc_flop / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        # the timed region is tens of milliseconds at full boost clocks: the BURST peak is the denominator
        peak = pk["bf16_tflops"]
        step_flop = FLOP_PER_IMAGE * B
        step_tf = step_flop / (ms_dev * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": f"tcgen05 implicit-GEMM conv, launch {dom_sig} x{B} images "
                                                 f"({dom_n} launches per step, largest share of the step)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "frac_of_sustained": achieved / pk["bf16_tflops_sustained"],
                    "traffic": traffic_for(dom_sig), "launches_per_step": dom_n,
                    "flop_per_launch": dom_flop / dom_n, "us_per_launch": 1e3 * dom_ms / dom_n,
                    "peak_source": pk["source"] + " burst bf16 (timed region << 1 s at full boost clocks)",
                    "all_tensor_core_convs": {"launches_per_step": n_tc, "flop_per_step": tc_flop, "ms_per_step": tc_ms,
                                              "achieved": all_tc, "frac": all_tc / peak},
                    "step_tflops": step_tf, "step_frac_burst": step_tf / peak,
                    "step_frac_sustained": step_tf / pk["bf16_tflops_sustained"]}
        gn_groups, (gn_n, gn_bytes, gn_ms), gn_calls = gn_profile(model, run.x, run.counter)
        roofline_hbm = None
        if gn_groups:
            gn_gbs = gn_bytes / (gn_ms * 1e-3) / 1e9 if gn_ms > 0 else 0.0
            # the GroupNorm launch signature that actually streams through HBM: the largest tensor (the others fit the
            # 126 MB L2; most GroupNorms of the step run inside the conv kernels)
            gd_sig, (gd_n, gd_bytes, gd_ms) = max(gn_groups.items(), key=lambda kv: kv[1][1] / kv[1][0])
            gd_us_in_step = 1e3 * gd_ms / gd_n
            gd_us, gd_timed, gd_sets = gn_isolated(gn_calls[gd_sig])
            gd_gbs = (gd_bytes / gd_n) / (gd_us * 1e-6) / 1e9 if gd_us > 0 else 0.0
            roofline_hbm = {"bound": "hbm", "kernel": f"gn_apply_kernel, launch {gd_sig} x{B} images ({gd_n} launches per step, "
                                                      "the largest stand-alone GroupNorm tensor; statistics come from "
                                                      "the producing conv's epilogue)",
                            "achieved": gd_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gd_gbs / pk["hbm_gbs"],
                            "traffic": traffic_for(gd_sig), "launches_per_step": gd_n,
                            "bytes_per_launch": gd_bytes / gd_n, "us_per_launch": gd_us,
                            "how": f"{gd_timed} launches back to back (one CUDA-graph replay) over {gd_sets} rotating copies of the tensors (footprint > 2x "
                                   "the 126 MB L2: every launch streams from / to HBM), one CUDA-event pair on the launching stream",
                            "us_per_launch_event_pair_in_eager_step": gd_us_in_step,
                            "peak_source": pk["source"] + " STREAM-style copy",
                            "note": "the GroupNorms of the 32x32 / 16x16 levels and of the output conv (most of the step's "
                                    "GroupNorm bytes) run inside the consuming conv kernels and have no pass of their own; the "
                                    "stand-alone launches left are 8-34 MB tensors (8x8 level), far below the 126 MB L2 and bound "
                                    "by launch + first-load latency, not by HBM",
                            "all_groupnorm_launches": {"launches_per_step": gn_n, "bytes_per_step": gn_bytes,
                                                       "ms_per_step": gn_ms, "achieved": gn_gbs,
                                                       "frac": gn_gbs / pk["hbm_gbs"],
                                                       "note": "per-launch event pairs inside the eager forward: small launches "
                                                               "include the host's launch latency (upper bound of their time)"}}
        roofline_sampler = sampler_roofline(ddpm, run.x, run.counter, seed, pk)
        run.reset()
        parity = parity_check(model, run.x, dev)
        cpu = None
        if not args.no_cpu and world == 1:  # the contract: rank 0 at N = 1 only (at N > 1 the other ranks' host threads
            # spin in the closing barrier and the CPU timing would measure that contention)
            rate_mean, _, cores, times, kind = cpu_step_rate(16, 5, 2)
            cpu = {"value": rate_mean / TIMESTEPS, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{len(times)} timed DDPM.sampling_step calls of "
                             f"{'the unmodified reference (baseline/_ref)' if kind == 'reference' else 'the oracle port'}, "
                             f"batch 16, fp32, {cores} torch threads; samples/s = image-steps/s / 1000"}
        line = {
            "metric": METRIC, "value": total / ms_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": total, "image": list(IMG),
                       "timesteps": TIMESTEPS,
                       "parallelism": f"sample-batch shard x{world} (rank r owns images [r*{B}, (r+1)*{B})), no data-path "
                                      "collective, one gather at the end",
                       "l2": f"per-step activation working set (~{4.2 * B:.0f} MB at this batch) exceeds the 126 MB L2 at "
                             "batch >= 32; no explicit flush",
                       "step": "CUDA graph: temb -> UNet fwd -> fused DDPM update -> t -= 1"},
            "e2e": {"value": total / ms_e2e, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": run.x_host.numel() * 4, "d2h_bytes_per_step": run.out_host.numel() * 4},
            "gpu_launches": run.launches_per_step * args.steps,
            "launches_per_step": run.launches_per_step,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_sampler": roofline_sampler,
            "parity": parity, "weak_scaling": weak, "batch_sweep_ms_per_step": sweep, "train": train,
            "cpu_baseline": cpu, "clocks": clk, "finite": finite,
        }
        if sweep:
            line["strong_scaling_projection"] = {str(n): sweep["256"] * 1.0 / sweep[str(256 // n)] for n in (1, 2, 4, 8)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The JSON line goes to the process's ORIGINAL stdout; everything else written to fd 1 (NCCL prints its version banner
    there when NCCL_DEBUG is set) has been redirected to stderr by quiet_stdout()."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def quiet_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dmme_b200", choices=["dmme_b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE config #2): 256 images in total; weak: 256 images per GPU")
    ap.add_argument("--batch", type=int, default=0, help="override the images per GPU (measurement sweeps)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step (config #4) leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the per-GPU batch sweep at N = 1")
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
        raise SystemExit(subprocess.call(cmd, stdout=_JSON_OUT))  # the ranks inherit the original stdout
    run_gpu(args)


if __name__ == "__main__":
    main()
