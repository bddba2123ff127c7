// A CHAIN of 3x3 stride-1 convolutions (the ResBlocks of the 8x8 / 4x4 levels, models/ddpm.py:118-133) in ONE persistent
// launch, image-stationary: a CTA owns `ipc` whole images for every conv of the chain, so
//   * nothing is exchanged between CTAs: no grid barrier, no split-K partial tiles, no finishing pass;
//   * the GroupNorm(+SiLU) between two convs (norm_act_drop_conv models/ddpm.py:25-35) is finished in the epilogue -- the
//     thread that owns an output channel sees every pixel of its images, the group statistics are a shuffle away -- and
//     the normalised tensor is written straight into shared memory as the next conv's MMA operand (it never exists in
//     HBM); raw block outputs (residual / skip / attention readers) and the normalised copies other consumers need go
//     to global memory from the same epilogue;
//   * only the WEIGHTS stream (L2 -> shared memory, a three-stage TMA ring that runs ahead into the next conv during an
//     epilogue); per conv a CTA spends one MMA phase and one epilogue instead of two launches (split-K GEMM + finish) or
//     conv + GroupNorm launches.
//
// Geometry (as conv_halo.cu): the CTA's images are a stack of zero-PADDED maps, position P = i * PP + (y+1) * WP + (x+1);
// a filter tap is a constant shift of d = (r-1) * WP + (s-1) position rows (128 bytes each, SWIZZLE_128B derived from the
// absolute shared-memory address), so one [positions][64 ch] chunk serves all nine taps through shifted operand
// descriptors.  The conv is computed transposed: D^T[256 cout][positions] = W X^T as two M = 128 accumulators of N =
// n_mma <= 256 columns each (the whole TMEM); TMEM lane = output channel.
//
// MEASURED NEGATIVE RESULT (DESIGN.md "chain kernel"; profiles/r2_chain_*): parity-green, 127 -> 86 launches per step, but
// SLOWER than the per-conv launches at every batch (256: 3.77 -> 4.13 ms, 32: 1.17 -> 1.63 ms per step).  An SM that owns
// whole images must ingest ALL the weights of every conv, and with N = 100..208 positions per weight tile the SS-mode MMA
// is bound by the shared-memory port: per 16 KB tile, 16 KB of TMA writes + 16 KB of A-operand reads + 14 KB of B reads =
// 46 KB / 128 B/clk = 360 clocks against 224 clocks of math (clock64 timeline, tools/trace_chain.py: a tile every 400-540
// clocks whatever the ring depth and the number of issuing threads).  The split-K path spreads one conv over all SMs with
// N = 256 tiles and wins despite its two launches per conv.  Built only with -DDMME_EXPERIMENTAL; the entry points stay in
// the ABI and answer "unsupported" otherwise.
//
// Warp roles: 0 weight-tile TMA producer, 1 activation-chunk TMA producer (inputs that come from global memory: the
// chain's first operand, the skip half of a concat, the raw input of a fused 1x1 residual conv), 2 MMA issuer / TMEM
// owner, 4..11 epilogue (TMEM lane quarter = warp % 4, accumulator = (warp - 4) / 4).
#include <cuda.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

#ifdef DMME_EXPERIMENTAL
namespace dmme {

constexpr int kChainMaxOps = 16;
constexpr int kChainCout = 256;
constexpr int kChainMaxWStages = 12;  // weight ring: as many 16 KB stages as fit beside the activation chunks
constexpr int kChainSlots = 2;
constexpr int kChainWTile = 128 * 128;  // [128 cout][64 k] bf16
constexpr int kChainThreads = 384;
constexpr int kChainEpiWarp0 = 4;
constexpr int kChainIssuers = 8;  // weight-tile TMA issuing threads
constexpr int kChainEpiThreads = 256;
constexpr int kChainSmemMax = 227 * 1024 - 512;  // dynamic shared memory the kernel may ask for (static barriers beside it)

struct ChainNorm {
  __nv_bfloat16* out;
  const float* gamma; const float* beta; const float* scale; const float* shift;
  int ss_rows, ss_ld, cpg, silu;
  float eps;
  int active, keep;
};

struct ChainOp {
  CUtensorMap w;        // [256][K] bf16, box [128][64]
  CUtensorMap src[4];   // src0, src1: normalised inputs of the 3x3 conv; res0, res1: raw inputs of the fused 1x1 conv;
                        // box = [64 ch][WP][WP][ipc] from (x, y) = (-1, -1): the zero fill is the conv padding
  int chunks[4];        // 64-channel chunks per source
  int resident_in;      // source 0 is the operand the previous op's epilogue left in shared memory
  const float* bias; const float* temb;
  int temb_rows, temb_ld;
  const __nv_bfloat16* addend;
  __nv_bfloat16* out;
  long long* stats;
  ChainNorm norm[2];
};

struct ChainParams {
  int nops, n, ipc, groups;
  int n_mma;   // MMA N: ipc * PP rounded up to a multiple of 16
  int rows;    // position rows per chunk buffer = n_mma + 2 * slack
  int slack;   // WP + 1 rows in front of position 0 (the taps reach that far back)
  int wstages; // weight-ring stages: a 16 KB tile takes ~2000 clocks from request to landing, so the stream's rate is
               // (stages x 16 KB) / latency -- three stages gave 18 B/clk, far below the MMAs' appetite
  long long* trace;  // debugging: clock64 timestamps of CTA 0, [role][1024] (dmme_debug_set_chain_trace), or null
  ChainOp op[kChainMaxOps];
};

// roles: 0 weight producer issued tile i, 1 MMA warp saw tile i landed, 2 epilogue of op i starts, 3 ends, 4 chunk producer
// issued streamed chunk i, 5 MMA warp saw streamed chunk i
__device__ __forceinline__ void chain_trace(long long* trace, int role, int idx) {
  if (trace && blockIdx.x == 0 && idx < 1024) trace[role * 1024 + idx] = clock64();
}

template <int W>
__global__ void __launch_bounds__(kChainThreads, 1) conv_chain_kernel(const __grid_constant__ ChainParams p) {
  constexpr int WP = W + 2, PP = WP * WP, HW = W * W;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full[kChainMaxWStages], w_empty[kChainMaxWStages];
  __shared__ __align__(8) uint64_t s_full[kChainSlots], s_empty[kChainSlots];
  __shared__ __align__(8) uint64_t acc_full, epi_done;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* wring = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int chunk_bytes = p.rows * 128;
  const int kChainWStages = p.wstages;
  uint8_t* resident = wring + kChainWStages * kChainWTile;  // 4 chunks: the 256-channel operand an epilogue leaves
  uint8_t* slots = resident + 4 * chunk_bytes;              // ring for chunks loaded from global memory

  if (threadIdx.x == 0) {
    for (int s = 0; s < kChainWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < kChainSlots; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&epi_done, kChainEpiThreads);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int k = 0; k < p.nops; ++k) tma_prefetch_desc(&p.op[k].w);
  }
  // (warp 3's issuing lanes share the descriptors warp 0 prefetched)
  if (warp == 1 && lane == 0) {
    for (int k = 0; k < p.nops; ++k)
      for (int j = 0; j < 4; ++j)
        if (p.op[k].chunks[j] && !(j == 0 && p.op[k].resident_in)) tma_prefetch_desc(&p.op[k].src[j]);
  }
  if (warp == 2) tmem_alloc(&tmem_slot, 512);
  {
    // the padding ring of the resident operand stays zero for the whole kernel: epilogues only write valid pixels
    uint4* z = reinterpret_cast<uint4*>(resident);
    const int n16 = 4 * chunk_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kChainThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();  // after the TMEM allocation (see common.cuh)

  if (warp == 0 || warp == 3) {
    // =========================== weight-tile producers ===========================
    // One thread needs ~400 clocks per TMA issue (measured: tools/trace_chain.py), i.e. 16 KB / 400 clk = 40 B/clk -- half of
    // what the MMAs consume at N = 112.  kChainIssuers threads (four lanes of two warps) therefore take the tiles of an op
    // round-robin; the lanes of a warp run the same instruction stream, so their issues overlap.
    // No more issuers than ring stages: a thread whose next tile lies two ring generations ahead of the slowest consumer
    // would pass the parity wait of the stage's `empty` barrier one generation too early.
    const int issuers = kChainIssuers < kChainWStages ? kChainIssuers : kChainWStages;
    const int me = (warp == 0 ? 0 : kChainIssuers / 2) + lane;
    if (lane < kChainIssuers / 2 && me < issuers) {
      int wit0 = 0;
      pdl_wait();  // the packed weights may come from a pack kernel launched just before
      for (int g = blockIdx.x; g < p.groups; g += gridDim.x) {
        for (int k = 0; k < p.nops; ++k) {
          const ChainOp& op = p.op[k];
          const int cchunks = op.chunks[0] + op.chunks[1];
          const int conv_tiles = 18 * cchunks;
          const int ntiles = conv_tiles + 2 * (op.chunks[2] + op.chunks[3]);
          // tile order (the MMA warp's): chunk-major, then tap, then the two 128-channel accumulators
          // thread `me` owns the tiles whose GLOBAL index is me mod issuers (also across op boundaries: two consecutive
          // tiles of a thread are never more than one ring length apart, see above)
          for (int idx = (me - wit0 % issuers + issuers) % issuers; idx < ntiles; idx += issuers) {
            int kcol, mt;
            if (idx < conv_tiles) {
              const int c = idx / 18, rem = idx - c * 18;
              mt = rem & 1;
              kcol = (c + (rem >> 1) * cchunks) * 64;  // packed K order: tap-major [tap][cat(src0, src1)]
            } else {
              const int r = idx - conv_tiles;
              mt = r & 1;
              kcol = (9 * cchunks + (r >> 1)) * 64;    // then the residual channels
            }
            const int wit = wit0 + idx;
            const int s = wit % kChainWStages;
            mbar_wait(&w_empty[s], ((wit / kChainWStages) & 1) ^ 1);
            if (me == 0) chain_trace(p.trace, 0, wit);
            mbar_expect_tx(&w_full[s], kChainWTile);
            tma_load_2d(wring + s * kChainWTile, &op.w, &w_full[s], kcol, mt * 128);
          }
          wit0 += ntiles;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== activation-chunk producer ===========================
    if (lane == 0) {
      int sit = 0, opc = 0;
      pdl_wait();
      for (int g = blockIdx.x; g < p.groups; g += gridDim.x) {
        for (int k = 0; k < p.nops; ++k, ++opc) {
          const ChainOp& op = p.op[k];
          // a global input of this op may be a raw output the previous op's epilogue just stored (the residual input);
          // waiting on every op also keeps this thread's phase bookkeeping of the barrier in step
          if (opc > 0) mbar_wait(&epi_done, (opc - 1) & 1);
          for (int j = 0; j < 4; ++j) {
            if (j == 0 && op.resident_in) continue;
            for (int c = 0; c < op.chunks[j]; ++c, ++sit) {
              const int s = sit % kChainSlots;
              mbar_wait(&s_empty[s], ((sit / kChainSlots) & 1) ^ 1);
              chain_trace(p.trace, 4, sit);
              mbar_expect_tx(&s_full[s], static_cast<uint32_t>(p.ipc) * PP * 128);
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
                  " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(slots + s * chunk_bytes + p.slack * 128)),
                  "l"(reinterpret_cast<uint64_t>(&op.src[j])), "r"(smem_u32(&s_full[s])), "r"(c * 64), "r"(-1), "r"(-1),
                  "r"(g * p.ipc)
                  : "memory");
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = umma_idesc_bf16(128, p.n_mma);
    int wit = 0, sit = 0, opc = 0;
    int ws = 0;
    uint32_t wph = 0;
    for (int g = blockIdx.x; g < p.groups; g += gridDim.x) {
      for (int k = 0; k < p.nops; ++k, ++opc) {
        const ChainOp& op = p.op[k];
        // the previous epilogue has drained the accumulators and written the resident operand
        if (opc > 0) mbar_wait(&epi_done, (opc - 1) & 1);
        tc_fence_after();
        bool first = true;
        for (int j = 0; j < 4; ++j) {
          const bool streamed = !(j == 0 && op.resident_in);
          const int ntaps = j < 2 ? 9 : 1;
          for (int c = 0; c < op.chunks[j]; ++c) {
            uint32_t x0_addr;
            int s = 0;
            if (streamed) {
              s = sit % kChainSlots;
              mbar_wait(&s_full[s], (sit / kChainSlots) & 1);
              tc_fence_after();
              if (lane == 0) chain_trace(p.trace, 5, sit);
              x0_addr = smem_u32(slots + s * chunk_bytes) + static_cast<uint32_t>(p.slack) * 128u;
            } else {
              x0_addr = smem_u32(resident + c * chunk_bytes) + static_cast<uint32_t>(p.slack) * 128u;
            }
            for (int tap = 0; tap < ntaps; ++tap) {
              const int d = ntaps == 9 ? (tap / 3 - 1) * WP + (tap % 3 - 1) : 0;
              const uint64_t xdesc = umma_desc_sw128(x0_addr + static_cast<uint32_t>(d * 128));
              for (int mt = 0; mt < 2; ++mt, ++wit) {
                mbar_wait(&w_full[ws], wph);
                tc_fence_after();
                if (lane == 0) chain_trace(p.trace, 1, wit);
                if (elect_one()) {
                  const uint64_t wdesc = umma_desc_sw128(smem_u32(wring + ws * kChainWTile));
                  const uint32_t dtm = tmem_base + static_cast<uint32_t>(mt * 256);
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(dtm, wdesc + 2 * kk, xdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                  umma_commit(&w_empty[ws]);
                }
                __syncwarp();
                if (++ws == kChainWStages) { ws = 0; wph ^= 1u; }
              }
              first = false;
            }
            if (streamed) {
              if (elect_one()) umma_commit(&s_empty[s]);
              __syncwarp();
              ++sit;
            }
          }
        }
        if (elect_one()) umma_commit(&acc_full);
        __syncwarp();
      }
    }
  } else if (warp >= kChainEpiWarp0) {
    // =========================== epilogue: thread = output channel ===========================
    const int e = warp - kChainEpiWarp0;
    const int q = e & 3, mt = e >> 2;
    const int ch = mt * 128 + q * 32 + lane;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt * 256);
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    // this channel's 2 bytes inside a position row of its chunk: 16-byte unit (XORed with the row's swizzle phase) + offset
    uint8_t* res_chunk = resident + (ch >> 6) * chunk_bytes;
    const uint32_t unit = static_cast<uint32_t>(ch & 63) >> 3, within = static_cast<uint32_t>(ch & 7) * 2u;
    int opc = 0;
    pdl_wait();
    for (int g = blockIdx.x; g < p.groups; g += gridDim.x) {
      for (int k = 0; k < p.nops; ++k, ++opc) {
        const ChainOp& op = p.op[k];
        const float bias_c = op.bias ? __ldg(op.bias + ch) : 0.f;
        mbar_wait(&acc_full, opc & 1);
        tc_fence_after();
        if (threadIdx.x == kChainEpiWarp0 * 32) chain_trace(p.trace, 2, opc);
        for (int i = 0; i < p.ipc; ++i) {
          const int n = g * p.ipc + i;
          if (n >= p.n) break;
          float add = bias_c;
          if (op.temb) add += __ldg(op.temb + static_cast<long long>(op.temb_rows == 1 ? 0 : n) * op.temb_ld + ch);
          const long long img = static_cast<long long>(n) * HW * kChainCout + ch;
          uint32_t rfp[HW / 2];  // the stored (rounded) outputs of this channel, two bf16 per register
          float s1 = 0.f, s2 = 0.f;
          // every accumulator row of the image is requested before the first is used (one TMEM round trip per image, not
          // one per row: the per-op epilogue took 15k clocks with a wait per row)
          uint32_t v[HW];
#pragma unroll
          for (int y = 0; y < W; ++y) {
            const uint32_t taddr = tbase + static_cast<uint32_t>(i * PP + (y + 1) * WP + 1);
            if constexpr (W == 8) tmem_ld8(taddr, reinterpret_cast<uint32_t(&)[8]>(v[y * W]));
            else tmem_ld4(taddr, reinterpret_cast<uint32_t(&)[4]>(v[y * W]));
          }
          tmem_ld_wait();
#pragma unroll
          for (int y = 0; y < W; ++y) {
            float av[W];
            if (op.addend) {
              // may be a raw output an earlier op of this very launch stored: read at L2.  (The engine folds identity
              // residuals into the GEMM instead -- [W | I] weights, the raw input as 1x1 chunks -- so that no epilogue
              // waits for global loads.)
#pragma unroll
              for (int x = 0; x < W; ++x)
                av[x] = __bfloat162float(__ldcg(op.addend + img + (y * W + x) * kChainCout));
            } else {
#pragma unroll
              for (int x = 0; x < W; ++x) av[x] = 0.f;
            }
#pragma unroll
            for (int x = 0; x < W; x += 2) {
              const __nv_bfloat16 r0 = __float2bfloat16_rn(__uint_as_float(v[y * W + x]) + add + av[x]);
              const __nv_bfloat16 r1 = __float2bfloat16_rn(__uint_as_float(v[y * W + x + 1]) + add + av[x + 1]);
              if (op.out) {
                op.out[img + (y * W + x) * kChainCout] = r0;
                op.out[img + (y * W + x + 1) * kChainCout] = r1;
              }
              const float f0 = __bfloat162float(r0), f1 = __bfloat162float(r1);
              s1 += f0 + f1;
              s2 = fmaf(f0, f0, fmaf(f1, f1, s2));
              rfp[(y * W + x) >> 1] = static_cast<uint32_t>(__bfloat16_as_ushort(r0)) |
                                      (static_cast<uint32_t>(__bfloat16_as_ushort(r1)) << 16);
            }
          }
          if (op.stats) {
            // [n][cout/4][2] fixed-point micro-group sums, the format every conv epilogue writes (dmme_conv_desc.stats)
            float m1 = s1, m2 = s2;
            m1 += __shfl_xor_sync(0xffffffffu, m1, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
            m1 += __shfl_xor_sync(0xffffffffu, m1, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
            if ((lane & 3) == 0) {
              unsigned long long* st = reinterpret_cast<unsigned long long*>(op.stats) +
                                       (static_cast<long long>(n) * (kChainCout >> 2) + (ch >> 2)) * 2;
              atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(m1 * kFix)));
              atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(m2 * kFix)));
            }
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const ChainNorm& nm = op.norm[j];
            if (!nm.active) continue;  // uniform
            float t1 = s1, t2 = s2;
            for (int o = 1; o < nm.cpg; o <<= 1) {
              t1 += __shfl_xor_sync(0xffffffffu, t1, o);
              t2 += __shfl_xor_sync(0xffffffffu, t2, o);
            }
            const float inv_cnt = 1.0f / static_cast<float>(HW * nm.cpg);
            const float mean = t1 * inv_cnt;
            const float var = fmaxf(t2 * inv_cnt - mean * mean, 0.f);
            const float rs = rsqrtf(var + nm.eps);
            const float ga = nm.gamma ? __ldg(nm.gamma + ch) : 1.f, be = nm.beta ? __ldg(nm.beta + ch) : 0.f;
            float aa = rs * ga, bb = be - mean * rs * ga;
            if (nm.scale) {
              const long long r = static_cast<long long>(nm.ss_rows == 1 ? 0 : n) * nm.ss_ld;
              const float sc = 1.f + __ldg(nm.scale + r + ch), sh = __ldg(nm.shift + r + ch);
              aa *= sc;
              bb = bb * sc + sh;
            }
#pragma unroll
            for (int px = 0; px < HW; ++px) {
              const uint32_t pk = rfp[px >> 1];
              const float rf = __uint_as_float((px & 1) ? (pk & 0xffff0000u) : (pk << 16));
              float yv = fmaf(rf, aa, bb);
              if (nm.silu) yv = silu_f(yv);
              const __nv_bfloat16 yb = __float2bfloat16_rn(yv);
              if (nm.keep) {
                const int row = p.slack + i * PP + (px / W + 1) * WP + (px % W + 1);
                uint8_t* ra = res_chunk + row * 128;
                const uint32_t phase = (smem_u32(ra) >> 7) & 7u;
                *reinterpret_cast<__nv_bfloat16*>(ra + ((unit ^ phase) << 4) + within) = yb;
              }
              if (nm.out) nm.out[img + px * kChainCout] = yb;
            }
          }
        }
        // resident operand: generic-proxy writes -> the MMAs' async-proxy reads; raw outputs: visible to this CTA's TMA
        // loads (the next op's residual input) before the barrier releases the chunk producer
        fence_proxy_async();
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        tc_fence_before();
        if (threadIdx.x == kChainEpiWarp0 * 32) chain_trace(p.trace, 3, opc);
        mbar_arrive(&epi_done);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int chain_wstages(int rows) {
  const int room = (kChainSmemMax - 1024 - (4 + kChainSlots) * rows * 128) / kChainWTile;
  return room > kChainMaxWStages ? kChainMaxWStages : room;
}

static int chain_geometry(int n, int w, int sms, int* ipc_out, int* n_mma_out, int* rows_out, int* smem_out) {
  const int wp = w + 2, pp = wp * wp, slack = wp + 1;
  long long best = -1;
  for (int ipc = 1; ipc * pp <= 256; ++ipc) {
    const int n_mma = ((ipc * pp + 15) / 16) * 16;
    if (n_mma > 256) break;
    const int rows = n_mma + 2 * slack;
    const int stages = chain_wstages(rows);
    if (stages < 2) break;
    const int smem = 1024 + stages * kChainWTile + (4 + kChainSlots) * rows * 128;
    const int groups = (n + ipc - 1) / ipc;
    const long long waves = (groups + sms - 1) / sms;
    // per weight tile a CTA spends max(four MMAs of N columns: 2 N clocks, the tile's share of the ring's round trip: ~2000
    // clocks / stages); the epilogue grows with the images
    const long long feed = 2000 / stages;
    const long long per_tile = 2 * n_mma > feed ? 2 * n_mma : feed;
    const long long cost = waves * (per_tile + 16 * ipc);
    if (best < 0 || cost < best) {
      best = cost;
      *ipc_out = ipc; *n_mma_out = n_mma; *rows_out = rows; *smem_out = smem;
    }
  }
  return best < 0 ? -1 : 0;
}

static int g_chain_ipc = 0;
static long long* g_chain_trace = nullptr;  // A/B: force the images per CTA (0 = cost model)

template <int W>
static int launch_chain(const ChainParams& p, int grid, int smem, cudaStream_t stream) {
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemMax);
    if (e != cudaSuccess) {
      set_error("conv_chain: cudaFuncSetAttribute(%d bytes): %s", kChainSmemMax, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured = true;
  }
  cudaError_t e = launch_pdl(conv_chain_kernel<W>, dim3(grid), dim3(kChainThreads), smem, stream, p);
  return check_launch_err(e, "conv_chain_kernel");
}

}  // namespace dmme
#endif  // DMME_EXPERIMENTAL

using namespace dmme;

#ifndef DMME_EXPERIMENTAL
extern "C" void dmme_set_conv_chain_ipc(int) {}
extern "C" void dmme_debug_set_chain_trace(long long*) {}
extern "C" int dmme_conv_chain_supported(int, int, int, int) { return 0; }
extern "C" int dmme_conv_chain_fwd(const dmme_chain_op*, int, int, int, int, void*) {
  set_error("conv_chain: built only with -DDMME_EXPERIMENTAL (measured slower than the per-conv launches, DESIGN.md)");
  return DMME_E_UNSUPPORTED;
}
#else
extern "C" void dmme_set_conv_chain_ipc(int ipc) { g_chain_ipc = ipc; }
// debugging: int64[6 * 1024] device buffer receiving CTA 0's per-role timestamps (tools/trace_chain.py), null = off
extern "C" void dmme_debug_set_chain_trace(long long* buf) { g_chain_trace = buf; }

extern "C" int dmme_conv_chain_supported(int n, int h, int w, int cout) {
  return n > 0 && h == w && (w == 4 || w == 8) && cout == kChainCout ? 1 : 0;
}

extern "C" int dmme_conv_chain_fwd(const dmme_chain_op* ops, int nops, int n, int h, int w, void* stream) {
  DMME_REQUIRE(ops != nullptr && nops >= 1 && nops <= kChainMaxOps, DMME_E_BADARG, "conv_chain: 1..%d ops (got %d)", kChainMaxOps, nops);
  DMME_REQUIRE(dmme_conv_chain_supported(n, h, w, kChainCout), DMME_E_SHAPE, "conv_chain: %dx%d maps not supported (4x4 / 8x8)", h, w);
  static std::mutex mu;
  static ChainParams p;  // ~14 KB kernel parameter block: kept off the stack, built under the lock, copied by the launch
  std::lock_guard<std::mutex> lock(mu);
  memset(&p, 0, sizeof(p));
  p.nops = nops; p.n = n;
  int smem = 0;
  const int sms = device_sm_count();
  DMME_REQUIRE(chain_geometry(n, w, sms, &p.ipc, &p.n_mma, &p.rows, &smem) == 0, DMME_E_SHAPE, "conv_chain: no geometry fits");
  if (g_chain_ipc > 0) {
    const int wp = w + 2, pp = wp * wp;
    p.ipc = g_chain_ipc;
    p.n_mma = ((p.ipc * pp + 15) / 16) * 16;
    p.rows = p.n_mma + 2 * (wp + 1);
    DMME_REQUIRE(p.n_mma <= 256 && chain_wstages(p.rows) >= 2, DMME_E_SHAPE, "conv_chain: forced ipc %d does not fit", p.ipc);
    smem = 1024 + chain_wstages(p.rows) * kChainWTile + (4 + kChainSlots) * p.rows * 128;
  }
  p.slack = w + 3;
  p.wstages = chain_wstages(p.rows);
  p.trace = g_chain_trace;
  p.groups = (n + p.ipc - 1) / p.ipc;
  for (int k = 0; k < nops; ++k) {
    const dmme_chain_op& o = ops[k];
    ChainOp& d = p.op[k];
    const int c[4] = {o.c0, o.c1, o.rc0, o.rc1};
    const void* src[4] = {o.src0, o.src1, o.res0, o.res1};
    DMME_REQUIRE(o.c0 > 0 && o.weight != nullptr, DMME_E_BADARG, "conv_chain: op %d lacks an input or its weight", k);
    d.resident_in = o.src0 == nullptr ? 1 : 0;
    if (d.resident_in) {
      DMME_REQUIRE(k > 0 && ops[k - 1].keep >= 0 && ops[k - 1].keep < 2 && o.c0 == kChainCout, DMME_E_BADARG,
                   "conv_chain: op %d reads the resident operand but op %d keeps none (or c0 != %d)", k, k - 1, kChainCout);
    }
    int ktot = 0;
    for (int j = 0; j < 4; ++j) {
      DMME_REQUIRE(c[j] >= 0 && c[j] % 64 == 0, DMME_E_SHAPE, "conv_chain: op %d source %d has %d channels (multiple of 64)", k, j, c[j]);
      DMME_REQUIRE(c[j] == 0 || src[j] != nullptr || (j == 0 && d.resident_in), DMME_E_BADARG, "conv_chain: op %d source %d is null", k, j);
      d.chunks[j] = c[j] / 64;
      ktot += (j < 2 ? 9 : 1) * c[j];
      if (c[j] && src[j]) {
        uint64_t dims[4] = {(uint64_t)c[j], (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t strides[3] = {(uint64_t)c[j] * 2, (uint64_t)w * c[j] * 2, (uint64_t)h * w * c[j] * 2};
        uint32_t box[4] = {64u, (uint32_t)(w + 2), (uint32_t)(h + 2), (uint32_t)p.ipc};
        int rc = encode_map(&d.src[j], src[j], 4, dims, strides, box);
        if (rc) return rc;
      }
    }
    {
      uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)kChainCout};
      uint64_t strides[1] = {(uint64_t)ktot * 2};
      uint32_t box[2] = {64u, 128u};
      int rc = encode_map(&d.w, o.weight, 2, dims, strides, box);
      if (rc) return rc;
    }
    d.bias = o.bias; d.temb = o.temb; d.temb_rows = o.temb_rows; d.temb_ld = o.temb_ld;
    d.addend = static_cast<const __nv_bfloat16*>(o.addend);
    d.out = static_cast<__nv_bfloat16*>(o.out);
    d.stats = o.stats;
    DMME_REQUIRE(o.keep >= -1 && o.keep < 2, DMME_E_BADARG, "conv_chain: op %d keep = %d", k, o.keep);
    for (int j = 0; j < 2; ++j) {
      const dmme_out_norm& s = o.out_norm[j];
      ChainNorm& q = d.norm[j];
      q.keep = o.keep == j ? 1 : 0;
      q.active = (s.out != nullptr || q.keep) ? 1 : 0;
      if (!q.active) continue;
      DMME_REQUIRE(s.cpg >= 1 && s.cpg <= 32 && 32 % s.cpg == 0, DMME_E_SHAPE,
                   "conv_chain: op %d out_norm[%d] channels per group must divide 32 (got %d)", k, j, s.cpg);
      DMME_REQUIRE(s.scale == nullptr || s.shift != nullptr, DMME_E_BADARG, "conv_chain: scale without shift");
      DMME_REQUIRE(s.out == nullptr || s.out != o.out, DMME_E_BADARG, "conv_chain: op %d out_norm[%d].out aliases out", k, j);
      q.out = static_cast<__nv_bfloat16*>(s.out);
      q.gamma = s.gamma; q.beta = s.beta; q.scale = s.scale; q.shift = s.shift;
      q.ss_rows = s.ss_rows; q.ss_ld = s.ss_ld; q.cpg = s.cpg; q.silu = s.silu; q.eps = s.eps;
    }
  }
  const int grid = p.groups < sms ? p.groups : sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return w == 8 ? launch_chain<8>(p, grid, smem, st) : launch_chain<4>(p, grid, smem, st);
}
#endif  // DMME_EXPERIMENTAL
