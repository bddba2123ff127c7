// 3x3 stride-1 convolution on tcgen05 with the activation halo tile reused across the nine filter taps.
//
// conv_tc.cu fetches one [128 px][64 ch] A tile per tap, so every activation byte crosses L2 -> SM nine times.
// Here the batch is viewed as a stack of ZERO-PADDED image rows: padded row PR = n*(H+2) + (y+1), position inside a
// tile P = row*(W+2) + (x+1).  In that space a filter tap is a constant shift d = (r-1)*(W+2) + (s-1), so ONE halo
// tile in shared memory (RT+2 padded rows x 64 channels, one TMA box per padded row, out-of-bounds pixels / rows /
// images zero-filled = the conv padding) serves all nine taps: the MMA operand descriptor is advanced by d rows of
// 128 bytes.  Measured on B200: the MMA derives the 128-byte-swizzle phase from the absolute shared-memory address
// exactly as TMA does when writing, so row-granular shifts need no re-layout and the descriptor's base-offset
// field stays 0 (setting it to (addr >> 7) & 7 gives wrong results).
//
// (Row tiles -- 16x16 and 32x32 maps -- share the padding between neighbours: one zero column per row and one zero row per
// image, pitch W+1; see kShared in the kernel.  The description here is the private-padding layout of the whole-image tiles.)
//
// Work unit = 128 output channels x RT whole padded rows (7 rows of 33 at 32x32, 15 rows of 17 at 16x16),
// computed TRANSPOSED: the weight tile [128 cout][64] is the MMA's M-side operand, the pixel rows are the N side
// (D^T = W X^T, one M=128, N<=256 instruction per 16 channels).  A 128x128 tile makes every MMA read 8 KB of shared
// memory per 64 cycles = the SM's whole 128 B/cycle (measured 41% tensor utilisation); 128x256 needs 96 B/cycle.
// The accumulator has TMEM lane = output channel, column = position; two accumulators (2 x 256 columns) are
// double-buffered so the epilogue of unit i overlaps the MMAs of unit i+1.  Because tiles hold whole rows, the
// epilogue reads exactly the W valid pixels of a row with one tcgen05.ld and stores them with immediate offsets:
// 6 instructions per output (an earlier position-masked version needed 25-66 and bound the kernel).
//
// Persistent CTAs (one per SM), static round-robin schedule.  Warp roles: 0..7 = epilogue (two warps per TMEM lane
// quarter), 8 = halo-tile TMA producer, 9 = weight-tile TMA producer, 10 = MMA issuer / TMEM owner.
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace dmme {

struct ConvHaloParams {
  CUtensorMap a[4];  // src0, src1, res0, res1: box = one padded row [W+2 px][64 ch]
  CUtensorMap at[4]; // the same tensors, box = the rt + 2 padded rows of a whole halo tile (tiles inside one image)
  CUtensorMap b;     // weights [cout][K] bf16, box [128][64]
  CUtensorMap b_half;  // the same tensor, box [64][64]: each CTA of a multicast pair loads half of every weight tile
  int chunks0, chunks1, rchunks0, rchunks1;
  unsigned char order[32];  // chunk order of a work unit, see halo_chunk_of
  int n, h, wp;
  int rt;            // padded rows per tile
  int n_mma;         // MMA N: rt * wp rounded up to a multiple of 16
  // tail tiles (row tiles only): row tiles [0, m_big) hold rt rows, tiles [m_big, m_tiles) hold rs < rt rows.  When the
  // equal tiles leave a last wave that occupies only some of the SMs, the rows of that wave are spread over all of them as
  // short tiles instead (1207 tiles of 7 rows on 148 SMs = 9 rounds of 240 columns -> 8 rounds + one of 80 columns)
  int m_big, rs, n_mma_s;
  int total_rows;    // n * (h + 2)
  int imgs_per_tile; // > 0: tiles hold whole padded images (small resolutions), one TMA box per image
  int m_tiles, n_tiles;
  const float* bias;
  const float* temb;
  int temb_rows, temb_ld;
  const __nv_bfloat16* addend;
  __nv_bfloat16* out;
  long long* stats;
  // fused GroupNorm(+SiLU) of the conv input (norm_act_drop_conv, models/ddpm.py:25-35): y = [silu](a * x + b) applied to
  // the halo tile in shared memory between the TMA load and the MMAs; (a, b) per (image, channel of cat(src0, src1))
  const float2* gn_ab;  // [n][gn_c] or null
  int gn_c;
  int gn_silu;
  long long* trace;  // debugging: per-role clock64 timestamps of CTA 0 (dmme_debug_set_halo_trace), or null
};

// trace slots of CTA 0: [role][chunk index]; roles: 0 stage free (producer issues the loads), 1 tile landed (transform
// starts), 2 transform done, 3 MMA warp saw the tile ready, 4 MMA warp issued the chunk's last tap, 5 accumulator full
// (epilogue starts; index = unit), 6 epilogue done
__device__ __forceinline__ void halo_trace(long long* trace, int role, int idx) {
  if (trace && blockIdx.x == 0 && idx < 256) trace[role * 256 + idx] = clock64();
}

// Chunk order inside a work unit (ConvHaloParams::order, built on the host: bit 7 = nine-tap conv chunk, low bits = its
// index).  The single-tap chunks of a fused 1x1 residual keep an activation stage busy for only 4 MMAs (~0.6k clocks against
// ~4.5k for a nine-tap chunk): issued back to back they fill the whole stage ring, and the next unit's first conv tile is
// loaded -- and, with a fused GroupNorm, transformed -- with nothing left to overlap it (per-role timeline,
// tools/trace_halo.py).  They are therefore spread between the conv chunks: conv chunk i sits at position
// i + ceil(i * rc / cc).
struct HaloChunk { bool is_conv; int idx; };
__device__ __forceinline__ HaloChunk halo_chunk_of(unsigned char code) { return HaloChunk{(code & 0x80) != 0, code & 0x7f}; }

constexpr int kHaloBN = 128;            // output channels per unit (MMA M)
constexpr int kHaloCols = 256;          // TMEM columns per accumulator stage
constexpr int kHaloASlot = 40 * 1024;   // >= ((rt + 2) * (W+2) + 1) * 128 bytes
constexpr int kHaloBSlot = kHaloBN * 128;
constexpr int kHaloAStages = 4;
constexpr int kHaloBStages = 4;
constexpr int kHaloSmem = kHaloAStages * kHaloASlot + kHaloBStages * kHaloBSlot + 1024;
constexpr int kHaloEpiWarps = 8;
constexpr int kHaloXfWarps = 8;  // GroupNorm transform warps (idle when the conv has no fused norm)
constexpr int kHaloProdA = 2;     // halo-tile producer warps: one thread issues a 4 KB row box every ~300 clocks, a tile has 9-16
constexpr int kHaloThreads = (kHaloEpiWarps + 3 + kHaloXfWarps + kHaloProdA - 1) * 32;
constexpr int kWarpProdA = kHaloEpiWarps, kWarpProdB = kHaloEpiWarps + 1, kWarpMma = kHaloEpiWarps + 2;
constexpr int kWarpXf0 = kHaloEpiWarps + 3;
constexpr int kWarpProdA2 = kWarpXf0 + kHaloXfWarps;  // further halo-tile producers: the row boxes of a tile are split between the issuing threads

// MC: clusters of two CTAs walk pairs of row tiles (same output channels) in lockstep and share the weight stream: each
// CTA loads half of every [128][64] weight tile and multicasts it into both CTAs' stage (half the L2 -> SM weight traffic,
// which is 2/3 of this kernel's operand feed); a stage is refilled once both CTAs' MMAs have released it.
template <int W, int COUT, bool MC>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kHaloAStages], a_empty[kHaloAStages];
  __shared__ __align__(8) uint64_t a_ready[kHaloAStages];  // halo tile transformed (fused GroupNorm): MMA may read it
  __shared__ __align__(8) uint64_t b_full[kHaloBStages], b_empty[kHaloBStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  // Row tiles (16x16 / 32x32 maps) use the SHARED-PADDING layout: one zero column per image row (x = -1) and one zero row
  // per image (y = -1).  With a row pitch of W + 1 that column is at once the left padding of its row and the right padding
  // of the row before it, the row the top padding of its image and the bottom padding of the image before it (the row
  // after the last image is out of bounds = zero-filled like every other padding position): 17 x 17 instead of 18 x 18
  // positions per 16x16 image, 15 instead of 14 rows per 256-column tile.  Tap shifts stay d = (r-1) * pitch + (s-1).
  // 8x8 maps take row tiles of one or two WHOLE images (9 or 18 stack rows: the fused norm keeps the coefficients of two
  // images per tile).  Whole-image tiles (4x4) keep private padding on all four sides: their halo rows are never loaded.
  constexpr bool kShared = W >= 8;
  constexpr int WP = kShared ? W + 1 : W + 2;  // row pitch in positions
  constexpr int HP = kShared ? W + 1 : W + 2;  // stack rows per image (square maps)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* abuf = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* bbuf = abuf + kHaloAStages * kHaloASlot;

  const int cchunks = p.chunks0 + p.chunks1;
  const int rchunks = p.rchunks0 + p.rchunks1;
  const int nck = cchunks + rchunks;
  const int units = p.m_tiles * p.n_tiles;
  constexpr int kRowBytes = WP * 128;
  const int nr = p.rt + 2;  // halo rows: one above and one below the tile
  // schedule: CTA (pair) `sched0` takes work items sched0, sched0 + nsched, ...; item v = (row tile [pair], channel tile)
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  const int nsched = MC ? gridDim.x / 2 : gridDim.x;
  const int sched0 = MC ? blockIdx.x / 2 : blockIdx.x;
  const int items = MC ? ((p.m_tiles + 1) / 2) * p.n_tiles : units;
  auto item_mt = [&](int v) { return MC ? 2 * (v / p.n_tiles) + static_cast<int>(crank) : v / p.n_tiles; };
  auto tile_row0 = [&](int mt) { return mt < p.m_big ? mt * p.rt : p.m_big * p.rt + (mt - p.m_big) * p.rs; };
  auto tile_rt = [&](int mt) { return mt < p.m_big ? p.rt : p.rs; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kHaloAStages; ++s) {
      mbar_init(&a_full[s], kHaloProdA);
      mbar_init(&a_empty[s], 1);
      mbar_init(&a_ready[s], kHaloXfWarps * 32);
    }
    for (int s = 0; s < kHaloBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], MC ? 2 : 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kHaloEpiWarps * 32); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (kShared && warp == 1) {
    // shared padding: tap (+1, +1) of the last pixel of the tile's last row reads the position AFTER the last halo row (the
    // padding column of the row below it).  No load ever writes that position: zero it once in every stage.
    *reinterpret_cast<uint4*>(abuf + (lane >> 3) * kHaloASlot + 128 + nr * kRowBytes + (lane & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (warp == kWarpProdA && lane == 0) {
    tma_prefetch_desc(&p.a[0]);
    if (p.chunks1) tma_prefetch_desc(&p.a[1]);
    if (p.rchunks0) tma_prefetch_desc(&p.a[2]);
    if (p.rchunks1) tma_prefetch_desc(&p.a[3]);
    if (p.imgs_per_tile == 0) {
      tma_prefetch_desc(&p.at[0]);
      if (p.chunks1) tma_prefetch_desc(&p.at[1]);
      if (p.rchunks0) tma_prefetch_desc(&p.at[2]);
      if (p.rchunks1) tma_prefetch_desc(&p.at[3]);
    }
    tma_prefetch_desc(MC ? &p.b_half : &p.b);
  }
  if (warp == kWarpMma) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  if (MC) cluster_sync_all();  // the peer's multicast loads and commits target this CTA's barriers
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();  // after the TMEM allocation (see ptx_sm100.cuh)

  if (warp == kWarpProdA || warp >= kWarpProdA2) {
    // =========================== halo-tile producers ===========================
    const int prod = warp == kWarpProdA ? 0 : warp - kWarpProdA2 + 1;
    const bool second = prod != 0;
    if (lane == 0) {
      int a_it = 0;
      pdl_wait();
      for (int u = sched0; u < items; u += nsched) {
        const int mt = item_mt(u);
        const int rt_u = tile_rt(mt), nr_u = rt_u + 2;
        const bool tail = rt_u != p.rt;
        const int pr0 = tile_row0(mt) - 1;  // first halo row (padded-row index, may be -1)
        // three of four tiles lie inside one padded image: ONE box of rt + 2 rows (rows -1 / h and everything past the batch
        // are out of bounds = zero-filled) instead of a box per row -- the issuing thread needs ~300 clocks per box
        // (shared padding: the tile's last halo row may be the next image's padding row = row h of this image)
        const int ni0 = (pr0 < 0 ? 0 : pr0) / HP;
        const int pr_last = pr0 + nr - 1 - (kShared ? 1 : 0);
        const bool one_img = p.imgs_per_tile == 0 && !tail && p.rt <= p.h && (pr_last < 0 ? 0 : pr_last) / HP == ni0;
        for (int ck = 0; ck < nck; ++ck, ++a_it) {
          int which, cc;
          const HaloChunk hc = halo_chunk_of(p.order[ck]);
          if (hc.is_conv) {
            which = hc.idx < p.chunks0 ? 0 : 1;
            cc = (which ? hc.idx - p.chunks0 : hc.idx) * 64;
          } else {
            const int rk = hc.idx;
            which = rk < p.rchunks0 ? 2 : 3;
            cc = (which == 3 ? rk - p.rchunks0 : rk) * 64;
          }
          const int as = a_it % kHaloAStages;
          mbar_wait(&a_empty[as], ((a_it / kHaloAStages) & 1) ^ 1);
          if (!second) halo_trace(p.trace, 0, a_it);
          if (p.imgs_per_tile > 0) {
            if (second) { mbar_arrive(&a_full[as]); continue; }
            // whole padded images: box = [64 ch][W+2 px][h+2 rows] from (x, y) = (-1, -1); rows outside the tile are only
            // ever read for padding-position outputs, so the two halo rows are not loaded at all
            mbar_expect_tx(&a_full[as], p.rt * kRowBytes);
            uint8_t* dst = abuf + as * kHaloASlot + 128 + kRowBytes;
            for (int i = 0; i < p.imgs_per_tile; ++i)
              tma_load_5d(dst + i * (p.h + 2) * kRowBytes, &p.a[which], &a_full[as], cc, -1, 0, -1, mt * p.imgs_per_tile + i);
            continue;
          }
          if (one_img) {
            if (second) { mbar_arrive(&a_full[as]); continue; }
            mbar_expect_tx(&a_full[as], nr * kRowBytes);
            tma_load_5d(abuf + as * kHaloASlot + 128, &p.at[which], &a_full[as], cc, -1, 0, pr0 - ni0 * HP - 1, ni0);
            continue;
          }
          // tiles spanning two images: a box per padded row, split between the producers.  The single-tap chunks of a fused
          // 1x1 residual only read the tile's own rows: the two halo rows are not loaded (whatever the slot holds there only
          // reaches padding-position accumulator columns)
          // (a tail tile also loads the row after its halo: that row's padding column is the position the one-time zeroing
          // above provides for the full-height tiles)
          int i_lo = hc.is_conv ? 0 : 1, i_hi = hc.is_conv ? nr_u + (tail ? 1 : 0) : nr_u - 1;
          const int i_cnt = i_hi - i_lo;
          i_hi = i_lo + (i_cnt * (prod + 1)) / kHaloProdA;
          i_lo = i_lo + (i_cnt * prod) / kHaloProdA;
          mbar_expect_tx(&a_full[as], (i_hi - i_lo) * kRowBytes);
          // slot layout: 128 bytes of slack (tap (-1,-1) of position 0 reaches one row back), then the halo rows
          uint8_t* dst = abuf + as * kHaloASlot + 128;
          for (int i = i_lo; i < i_hi; ++i) {
            const int pr = pr0 + i;
            int ni, yy;
            if (pr < 0) { ni = -1; yy = 0; }             // before the first image: whole row out of bounds -> zeros
            else { ni = pr / HP; yy = pr - ni * HP - 1; }  // yy = -1 or h: padding row -> zeros
            tma_load_5d(dst + i * kRowBytes, &p.a[which], &a_full[as], cc, -1, 0, yy, ni);
          }
        }
      }
    }
  } else if (warp == kWarpProdB) {
    // =========================== weight-tile producer ===========================
    if (lane == 0) {
      int b_it = 0;
      pdl_wait();  // the packed weights may come from a pack kernel launched just before this one
      for (int u = sched0; u < items; u += nsched) {
        const int col0 = (u % p.n_tiles) * kHaloBN;
        for (int ck = 0; ck < nck; ++ck) {
          const HaloChunk hc = halo_chunk_of(p.order[ck]);
          const bool is_conv = hc.is_conv;
          const int ntaps = is_conv ? 9 : 1;
          const int kb0 = is_conv ? hc.idx : 9 * cchunks + hc.idx;
          const int kbs = is_conv ? cchunks : 0;
          for (int tap = 0; tap < ntaps; ++tap, ++b_it) {
            const int bs = b_it % kHaloBStages;
            mbar_wait(&b_empty[bs], ((b_it / kHaloBStages) & 1) ^ 1);
            mbar_expect_tx(&b_full[bs], kHaloBSlot);
            if (MC)
              tma_load_2d_mc(bbuf + bs * kHaloBSlot + crank * (kHaloBSlot / 2), &p.b_half, &b_full[bs], (kb0 + tap * kbs) * 64,
                             col0 + static_cast<int>(crank) * (kHaloBN / 2), 3);
            else
              tma_load_2d(bbuf + bs * kHaloBSlot, &p.b, &b_full[bs], (kb0 + tap * kbs) * 64, col0);
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // =========================== MMA issuer ===========================
    // the whole warp walks the loop (converged waits); one elected lane issues the MMAs and their commits
    const uint32_t idesc_big = umma_idesc_bf16(kHaloBN, p.n_mma);  // M = 128 output channels, N = positions of the tile
    const uint32_t idesc_tail = umma_idesc_bf16(kHaloBN, p.n_mma_s);
    int a_it = 0, b_it = 0, u_it = 0;
    for (int u = sched0; u < items; u += nsched, ++u_it) {
      const int stage = u_it & 1;
      const uint32_t idesc = item_mt(u) < p.m_big ? idesc_big : idesc_tail;
      mbar_wait(&acc_empty[stage], ((u_it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t dtm = tmem_base + stage * kHaloCols;
      for (int ck = 0; ck < nck; ++ck, ++a_it) {
        const bool is_conv = halo_chunk_of(p.order[ck]).is_conv;
        const int ntaps = is_conv ? 9 : 1;
        const int as = a_it % kHaloAStages;
        mbar_wait(p.gn_ab ? &a_ready[as] : &a_full[as], (a_it / kHaloAStages) & 1);
        tc_fence_after();
        if (lane == 0) halo_trace(p.trace, 3, a_it);
        // position 0 of the tile = first pixel slot of the tile's first row = halo row 1
        const uint32_t x0_addr = smem_u32(abuf + as * kHaloASlot) + 128u + kRowBytes;
        for (int tap = 0; tap < ntaps; ++tap, ++b_it) {
          const int bs = b_it % kHaloBStages;
          mbar_wait(&b_full[bs], (b_it / kHaloBStages) & 1);
          tc_fence_after();
          const int d = is_conv ? (tap / 3 - 1) * WP + (tap % 3 - 1) : 0;
          if (elect_one()) {
            const uint64_t wdesc = umma_desc_sw128(smem_u32(bbuf + bs * kHaloBSlot));          // [128 cout][64]: M side
            const uint64_t xdesc = umma_desc_sw128(x0_addr + static_cast<uint32_t>(d * 128));  // shifted pixel rows: N side
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(dtm, wdesc + 2 * k, xdesc + 2 * k, idesc, (ck | tap | k) != 0 ? 1u : 0u);
            if (MC) umma_commit_mc(&b_empty[bs], 3);
            else umma_commit(&b_empty[bs]);
            if (tap == ntaps - 1) {
              umma_commit(&a_empty[as]);
              if (ck == nck - 1) umma_commit(&acc_full[stage]);
            }
          }
          __syncwarp();
        }
        if (lane == 0) halo_trace(p.trace, 4, a_it);
      }
    }
  } else if (warp >= kWarpXf0 && warp < kWarpProdA2) {
    // =========================== fused GroupNorm(+SiLU) of the halo tile ===========================
    // Thread = one 16-byte unit column (eight channels of the chunk: its 16 coefficients live in registers, for the two
    // images a tile can touch) x every 16th position row.  Padding positions (x = -1, W; rows above / below an image;
    // rows past the batch) were zero-filled by TMA and must stay zero: the reference pads AFTER norm + activation.
    // Traffic: the tile is read and written once more through the shared-memory port (about +13% on this kernel) against
    // a whole stand-alone pass over the tensor through HBM / L2 and its launch.
    if (p.gn_ab != nullptr && p.imgs_per_tile == 0) {
      const int xt = threadIdx.x - kWarpXf0 * 32;  // 0..kHaloXfWarps * 32
      const int ul = xt & 7;                       // logical 16-byte unit = channels [8 ul, 8 ul + 8) of the chunk
      const int r_first = xt >> 3;                 // rows r_first, r_first + 4 * kHaloXfWarps, ...
      int a_it = 0;
      pdl_wait();
      for (int u = sched0; u < items; u += nsched) {
        const int mt = item_mt(u);
        const int pr0 = tile_row0(mt) - 1;
        const int rows = (tile_rt(mt) + 2) * WP;
        // first image the tile touches; it touches at most n_lo + 1 too (clamped: a pair's odd tile may lie past the batch)
        // (image-aligned tiles -- 8x8 maps -- start at an image's padding row: their halo row above belongs to the image
        // before and only feeds padding-row outputs, so the tile's two images are the ones counted from its first own row)
        const int pr_first = (p.rt % HP) == 0 ? pr0 + 1 : pr0;
        const int n_lo = min((pr_first < 0 ? 0 : pr_first) / HP, p.n - 1);
        for (int ck = 0; ck < nck; ++ck, ++a_it) {
          const int as = a_it % kHaloAStages;
          const HaloChunk hc = halo_chunk_of(p.order[ck]);
          const bool is_conv = hc.is_conv;
          float2 c0v[8], c1v[8];
          if (is_conv) {
            const int cbase = hc.idx * 64 + ul * 8;
            const float4* g0 = reinterpret_cast<const float4*>(p.gn_ab + static_cast<long long>(n_lo) * p.gn_c + cbase);
            const bool has1 = n_lo + 1 < p.n;
            const float4* g1 = reinterpret_cast<const float4*>(p.gn_ab + static_cast<long long>(has1 ? n_lo + 1 : n_lo) * p.gn_c + cbase);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 t0 = __ldg(g0 + j), t1 = __ldg(g1 + j);
              c0v[2 * j] = make_float2(t0.x, t0.y); c0v[2 * j + 1] = make_float2(t0.z, t0.w);
              c1v[2 * j] = make_float2(t1.x, t1.y); c1v[2 * j + 1] = make_float2(t1.z, t1.w);
            }
          }
          mbar_wait(&a_full[as], (a_it / kHaloAStages) & 1);
          if (xt == 0) halo_trace(p.trace, 1, a_it);
          if (is_conv) {
            uint8_t* tile = abuf + as * kHaloASlot + 128;
            const uint32_t tile_addr = smem_u32(tile);
            // All loads of a batch are issued before the first value is used: the MMAs keep the shared-memory port
            // nearly saturated, so one load-use round trip per row (an earlier version) made this stage latency-bound
            // and slower than the MMAs it feeds (per-role timeline: tools/trace_halo.py).
            constexpr int kXfBatch = 5;
            for (int rb = r_first; rb < rows; rb += 4 * kHaloXfWarps * kXfBatch) {
              uint4 v[kXfBatch];
              uint32_t off[kXfBatch];
              int sel[kXfBatch];  // 0: padding position (stays zero), 1 / 2: first / second image of the tile
#pragma unroll
              for (int b = 0; b < kXfBatch; ++b) {
                const int r = rb + b * 4 * kHaloXfWarps;
                const int hr = r / WP, xx = r - hr * WP;
                const int pr = pr0 + hr;
                const int n = pr / HP;
                const int yy = pr - n * HP - 1;
                const bool valid = r < rows && xx != 0 && xx <= W && pr >= 0 && pr < p.total_rows && yy >= 0 && yy < W;
                sel[b] = valid ? (n != n_lo ? 2 : 1) : 0;
                // SWIZZLE_128B: the 16-byte unit index is XORed with address bits [7, 10) of the row
                const uint32_t phase = ((tile_addr + static_cast<uint32_t>(r) * 128u) >> 7) & 7u;
                off[b] = static_cast<uint32_t>(r) * 128u + ((static_cast<uint32_t>(ul) ^ phase) << 4);
                if (valid) v[b] = *reinterpret_cast<const uint4*>(tile + off[b]);
              }
#pragma unroll
              for (int b = 0; b < kXfBatch; ++b) {
                if (sel[b] == 0) continue;
                float f[8];
                unpack_bf16x2(v[b].x, f[0], f[1]); unpack_bf16x2(v[b].y, f[2], f[3]);
                unpack_bf16x2(v[b].z, f[4], f[5]); unpack_bf16x2(v[b].w, f[6], f[7]);
                const bool second = sel[b] == 2;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 ab = second ? c1v[j] : c0v[j];
                  f[j] = fmaf(f[j], ab.x, ab.y);
                }
                if (p.gn_silu) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
                }
                uint4 o;
                o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
                o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
                *reinterpret_cast<uint4*>(tile + off[b]) = o;
              }
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the MMA's async-proxy reads
          }
          if (xt == 0) halo_trace(p.trace, 2, a_it);
          mbar_arrive(&a_ready[as]);
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    // thread = one output channel (TMEM lane); warp (q, half) owns channels [32q, 32q+32) and every second row
    const int q = warp & 3;
    const int half = warp >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float kFix = static_cast<float>(1 << DMME_STATS_FRAC_BITS);
    const bool temb_per_image = p.temb && p.temb_rows != 1;
    int u_it = 0;
    pdl_wait();
    for (int u = sched0; u < items; u += nsched, ++u_it) {
      const int mt = item_mt(u), nt = u % p.n_tiles;
      const int row0 = tile_row0(mt), rt_u = tile_rt(mt);
      const int ch = nt * kHaloBN + q * 32 + lane;
      const float bias_c = p.bias ? __ldg(p.bias + ch) : 0.f;
      const int stage = u_it & 1;
      float s1 = 0.f, s2 = 0.f, bt = bias_c;
      int cur_n = -1;

      auto flush_stats = [&]() {  // warp-uniform call: per-image sums of this lane's channel -> micro-group atomics
        if (p.stats && cur_n >= 0) {
          float a1 = s1, a2 = s2;
          a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
          a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
          if ((lane & 3) == 0) {
            unsigned long long* st = reinterpret_cast<unsigned long long*>(p.stats) +
                                     (static_cast<long long>(cur_n) * (COUT >> 2) + (ch >> 2)) * 2;
            atomicAdd(st, static_cast<unsigned long long>(__float2ll_rn(a1 * kFix)));
            atomicAdd(st + 1, static_cast<unsigned long long>(__float2ll_rn(a2 * kFix)));
          }
        }
        s1 = 0.f; s2 = 0.f;
      };

      mbar_wait(&acc_full[stage], (u_it >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 0) halo_trace(p.trace, 5, u_it);
#pragma unroll 1
      for (int rr = half; rr < rt_u; rr += 2) {
        const int pr = row0 + rr;
        if (pr >= p.total_rows) break;
        const int n = pr / HP;
        const int yy = pr - n * HP - 1;
        if (yy < 0 || yy >= p.h) continue;  // padding row: junk accumulator columns
        if (n != cur_n) {
          flush_stats();
          cur_n = n;
          bt = bias_c;
          if (p.temb) bt += __ldg(p.temb + static_cast<long long>(temb_per_image ? n : 0) * p.temb_ld + ch);
        }
        const long long o = (static_cast<long long>(n) * p.h + yy) * W * COUT + ch;
        __nv_bfloat16* op = p.out + o;
        const uint32_t taddr = tmem_base + lane_off + static_cast<uint32_t>(stage * kHaloCols + rr * WP + 1);
        // the addend row is fetched before the accumulator row is awaited, and every load is issued before the first
        // store: out and addend may alias as far as the compiler knows, and a load placed after a store waits for it
        // (measured: 550 us instead of 71 us per launch)
        float av[W];
        if (p.addend) {
          const __nv_bfloat16* __restrict__ ap = p.addend + o;
#pragma unroll
          for (int i = 0; i < W; ++i) av[i] = __bfloat162float(__ldg(ap + i * COUT));
        } else {
#pragma unroll
          for (int i = 0; i < W; ++i) av[i] = 0.f;
        }
        uint32_t v[W];
        if constexpr (W == 32) tmem_ld32(taddr, v);
        else if constexpr (W == 16) tmem_ld16(taddr, v);
        else if constexpr (W == 8) tmem_ld8(taddr, v);
        else tmem_ld4(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < W; ++i) {
          const float val = __uint_as_float(v[i]) + bt + av[i];
          const __nv_bfloat16 r = __float2bfloat16_rn(val);
          op[i * COUT] = r;
          const float rf = __bfloat162float(r);
          s1 += rf;
          s2 = fmaf(rf, rf, s2);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[stage]);
      if (threadIdx.x == 0) halo_trace(p.trace, 6, u_it);
      flush_stats();
    }
  }

  tc_fence_before();
  if (MC) cluster_sync_all();  // neither CTA may leave while the other can still write its stages / signal its barriers
  else __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int g_halo_mode = 1;  // 1: AUTO prefers this kernel, 0: AUTO never picks it (A/B measurements)
static int g_sm_count = 0;
static long long* g_halo_trace = nullptr;

bool conv_halo_supported(const dmme_conv_desc& d) {
  if (g_halo_mode == 0) return false;
  if (d.act_dtype != DMME_BF16 || d.in_layout != DMME_IN_NHWC || d.out_layout != DMME_OUT_NHWC) return false;
  if (d.upsample || d.ksize != 3 || d.stride != 1) return false;
  if (d.c0 <= 0 || d.c0 % 64 || d.c1 % 64 || d.rc0 % 64 || d.rc1 % 64) return false;
  if (d.cout != 128 && d.cout != 256) return false;
  if (d.w_in != 4 && d.w_in != 8 && d.w_in != 16 && d.w_in != 32) return false;
  if (d.h_in != d.w_in) return false;
  if (d.h_in < 4 || static_cast<long long>(d.n) * (d.h_in + 2) > (1 << 24)) return false;
  return true;
}


// AUTO's choice between the two tcgen05 kernel families (both are correct wherever both are supported), by measurement at
// batch 256 (tools/prof_conv.py on the step's signatures, gpurun_out/r54_ab.log): the halo kernel's 9-tap reuse of the
// activation tile wins at 32x32 and 16x16 (76 vs 86 us, 78 vs 88 us), also with a fused 1x1 residual of up to 128 channels
// (16x16: 82 vs 90 us) -- and it can apply the GroupNorm of its input itself; wider residuals add single-tap chunks whose
// tile loads it cannot hide (16x16 256->256 +res512: 123 vs 102 us; 128->128 +res256 is 46 us with its GroupNorm inside
// against 40 + 12 us), and the padded-position
// overhead (36% at 8x8, 56% at 4x4) cancels the gain below 16x16 (37 vs 31 us, 27 vs 23 us).
bool conv_halo_preferred(const dmme_conv_desc& d) {
  if (!conv_halo_supported(d)) return false;
  // 8x8 maps (round 2: row tiles of two whole images in the shared-padding layout, GroupNorm inside) are parity-green and
  // SLOWER than the transposed kernel + stand-alone GroupNorm (3.11-3.12 vs 3.06 ms per step at batch 256): with N = 176
  // positions per MMA the weight tiles are 40 % of the shared-memory port bytes (a nine-tap chunk takes 4.4 k clocks for
  // 3.2 k of MMA) and 256 units need two rounds where the transposed kernel's 128 units need one.  Bit 7 of the mode takes
  // them anyway at large batches (A/B); explicit DMME_CONV_HALO requests always run.
  if (d.w_in == 8 && (!(g_halo_mode & 128) || d.n < 160)) return false;
  if (d.w_in < 8) return false;
  // (round 2, shared-padding row tiles: the halo kernel is ~13% faster at 16x16 and the wide-residual convs now win with
  // their GroupNorm inside -- whole step 3.06 vs 3.11 ms at batch 256, 1.844 vs 1.866 at 128, 1.086 vs 1.099 at 32; bit 6 of the
  // mode restores the old rule for A/B runs)
  if ((g_halo_mode & 64) && d.w_in == 16 && (d.rc0 + d.rc1) > 128 && !(d.cout == 128 && d.rc0 + d.rc1 <= 256)) return false;
  return true;
}

// weight multicast across CTA pairs (clusters of two): 0 = off (default: measured 3-15% SLOWER on every signature of the
// step -- the weight feed is not what paces this kernel and the pair runs in lockstep), 1 = launches with several work
// items per CTA, 2 = every launch
static int g_halo_mc = 0;

template <int W, int COUT, bool MC>
static int launch_halo_mc(const ConvHaloParams& p, cudaStream_t stream) {
  static DeviceOnce once_;
  bool& configured = once_.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<W, COUT, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem);
    if (e != cudaSuccess) {
      set_error("conv_halo: cudaFuncSetAttribute(%d bytes): %s", kHaloSmem, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  if (MC) {
    const int items = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int pairs = items < g_sm_count / 2 ? items : g_sm_count / 2;
    cudaError_t e = launch_pdl_pair(conv_halo_kernel<W, COUT, MC>, dim3(2 * pairs), dim3(kHaloThreads), kHaloSmem, stream, p);
    return check_launch_err(e, "conv_halo_kernel (multicast pairs)");
  }
  const int units = p.m_tiles * p.n_tiles;
  const int grid = units < g_sm_count ? units : g_sm_count;
  cudaError_t e = launch_pdl(conv_halo_kernel<W, COUT, MC>, dim3(grid), dim3(kHaloThreads), kHaloSmem, stream, p);
  return check_launch_err(e, "conv_halo_kernel");
}

template <int W, int COUT>
static int launch_halo(const ConvHaloParams& p, cudaStream_t stream) {
  // pairs pay off where CTAs walk several work items each (the 16x16 and 32x32 levels at sampling batch sizes)
#ifdef DMME_EXPERIMENTAL  // weight multicast measured 3-15% slower: built only with -DDMME_EXPERIMENTAL
  if (g_halo_mc == 2 || (g_halo_mc == 1 && p.m_tiles * p.n_tiles >= 2 * g_sm_count)) return launch_halo_mc<W, COUT, true>(p, stream);
#endif
  return launch_halo_mc<W, COUT, false>(p, stream);
}

int conv_halo_forward(const dmme_conv_desc& d, cudaStream_t stream) {
  DMME_REQUIRE(conv_halo_supported(d), DMME_E_SHAPE, "conv_halo: unsupported shape/layout");
  DMME_REQUIRE(d.src0 && d.weight && d.out, DMME_E_BADARG, "conv_halo: null src0/weight/out");
  DMME_REQUIRE(d.c1 == 0 || d.src1, DMME_E_BADARG, "conv_halo: c1 > 0 but src1 is null");
  DMME_REQUIRE(d.rc0 == 0 || d.res0, DMME_E_BADARG, "conv_halo: rc0 > 0 but res0 is null");
  DMME_REQUIRE(d.rc1 == 0 || d.res1, DMME_E_BADARG, "conv_halo: rc1 > 0 but res1 is null");
  ConvHaloParams p;
  memset(&p, 0, sizeof(p));
  p.chunks0 = d.c0 / 64; p.chunks1 = d.c1 / 64; p.rchunks0 = d.rc0 / 64; p.rchunks1 = d.rc1 / 64;
  const bool shared_pad = d.w_in >= 8;  // row tiles: shared-padding layout (see the kernel)
  const int hp = shared_pad ? d.h_in + 1 : d.h_in + 2;
  p.n = d.n; p.h = d.h_in; p.wp = shared_pad ? d.w_in + 1 : d.w_in + 2;
  {
    const int cc = p.chunks0 + p.chunks1, rc = p.rchunks0 + p.rchunks1;
    DMME_REQUIRE(cc + rc <= 32, DMME_E_SHAPE, "conv_halo: more than 32 channel chunks per work unit");
    int pos = 0, r = 0;
    for (int i = 0; i < cc; ++i) {
      const int r_before = g_halo_mode & 16 ? 0 : (i * rc + cc - 1) / cc;  // bit 4 (A/B): residual chunks after the conv chunks
      while (r < r_before) p.order[pos++] = static_cast<unsigned char>(r++);
      p.order[pos++] = static_cast<unsigned char>(0x80 | i);
    }
    while (r < rc) p.order[pos++] = static_cast<unsigned char>(r++);
  }
  p.total_rows = d.n * hp;
  p.n_tiles = d.cout / kHaloBN;
  g_sm_count = device_sm_count();
  const int img_pos = (d.h_in + 2) * (d.w_in + 2);
  p.imgs_per_tile = 0;
  if (!shared_pad) {
    // small resolutions: a tile is k whole padded images (2 at 8x8, 7 at 4x4), or fewer when that fills more SMs
    long long best_cost = -1;
    for (int k = kHaloCols / img_pos; k >= 1; k = k / 2) {
      const int n_mma = ((k * img_pos + 15) / 16) * 16;
      const long long units = static_cast<long long>((d.n + k - 1) / k) * p.n_tiles;
      const long long cost = ((units + g_sm_count - 1) / g_sm_count) * (n_mma + 48);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; p.imgs_per_tile = k; p.n_mma = n_mma; }
      if (k == 1) break;
    }
    p.rt = p.rs = p.imgs_per_tile * (d.h_in + 2);
    p.m_tiles = p.m_big = (p.total_rows + p.rt - 1) / p.rt;
    p.n_mma_s = p.n_mma;
  } else {
    // rows per tile: at most what fits 256 accumulator columns (7 rows of 34, 14 of 18); fewer rows when that fills the
    // last wave better (cost ~ waves x (MMA width + per-unit overhead; 96 columns' worth by measurement: with 48 the 32x32
    // level went to 6-row tiles and lost 4%)): at batch 256 the 128-channel 16x16 convs
    // take 11 rows (419 units = 2.8 waves) instead of 14 (330 units = 2.2 waves, a third wave for a fifth of the SMs)
    long long best_cost = -1;
    const bool whole_imgs = d.w_in == 8;  // 8x8: tiles of one or two whole images (hp rows each)
    const int rt_max = whole_imgs ? 2 * hp : kHaloCols / p.wp;
    for (int rt = rt_max; rt >= 1 && 2 * rt >= rt_max; --rt) {
      if (whole_imgs && rt % hp) continue;
      const int n_mma = ((rt * p.wp + 15) / 16) * 16;
      const int m_all = (p.total_rows + rt - 1) / rt;
      const long long units = static_cast<long long>(m_all) * p.n_tiles;
      const long long cost = ((units + g_sm_count - 1) / g_sm_count) * (n_mma + 96);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost; p.rt = rt; p.n_mma = n_mma; p.m_tiles = p.m_big = m_all; p.rs = rt; p.n_mma_s = n_mma;
      }
      // tail tiles: the full rounds as above, the rows of the partial last round as one round of short tiles
      const long long full = units / g_sm_count;
      if (shared_pad && !whole_imgs && (g_halo_mode & 32) == 0 && g_halo_mc == 0 && full >= 1 && units % g_sm_count != 0 && g_sm_count % p.n_tiles == 0) {
        const int m_big = static_cast<int>(full * g_sm_count / p.n_tiles);
        const int rows_left = p.total_rows - m_big * rt;
        const int per_round = g_sm_count / p.n_tiles;
        const int rs = (rows_left + per_round - 1) / per_round;
        if (rows_left > 0 && rs < rt) {
          const int n_mma_s = ((rs * p.wp + 15) / 16) * 16;
          const long long cost_t = full * (n_mma + 96) + (n_mma_s + 96);
          if (cost_t < best_cost) {
            best_cost = cost_t; p.rt = rt; p.n_mma = n_mma; p.m_big = m_big; p.rs = rs; p.n_mma_s = n_mma_s;
            p.m_tiles = m_big + (rows_left + rs - 1) / rs;
          }
        }
      }
    }
  }
  p.bias = d.bias; p.temb = d.temb; p.temb_rows = d.temb_rows; p.temb_ld = d.temb_ld;
  p.addend = static_cast<const __nv_bfloat16*>(d.addend);
  p.out = static_cast<__nv_bfloat16*>(d.out);
  p.stats = d.stats;
  p.gn_ab = reinterpret_cast<const float2*>(d.gn_ab);
  p.gn_c = d.c0 + d.c1;
  p.gn_silu = d.gn_silu;
  p.trace = g_halo_trace;
  DMME_REQUIRE(d.gn_ab == nullptr || p.imgs_per_tile == 0, DMME_E_UNSUPPORTED,
               "conv_halo: fused GroupNorm needs row tiles (16x16 and 32x32 maps)");
  // the MMA reads n_mma + (W+3) position rows past the first tile position; keep that inside the slot
  DMME_REQUIRE((1 + (p.rt + 2) * p.wp + (shared_pad ? 1 : 0)) * 128 <= kHaloASlot && (1 + 2 * p.wp + 1 + p.n_mma) * 128 <= kHaloASlot,
               DMME_E_SHAPE, "conv_halo: halo tile does not fit its shared-memory slot");

  auto act_map = [&](CUtensorMap* m, const void* ptr, int c, int rows = 0) -> int {
    uint64_t dims[5] = {(uint64_t)c, (uint64_t)d.w_in, 1, (uint64_t)d.h_in, (uint64_t)d.n};
    uint64_t strides[4] = {(uint64_t)c * 2, (uint64_t)d.w_in * c * 2, (uint64_t)d.w_in * c * 2,
                           (uint64_t)d.h_in * d.w_in * c * 2};
    uint32_t box[5] = {64u, (uint32_t)p.wp, 1u, rows ? (uint32_t)rows : p.imgs_per_tile > 0 ? (uint32_t)(d.h_in + 2) : 1u, 1u};
    return encode_map(m, ptr, 5, dims, strides, box);
  };
  int rc;
  if ((rc = act_map(&p.a[0], d.src0, d.c0))) return rc;
  if (d.c1 && (rc = act_map(&p.a[1], d.src1, d.c1))) return rc;
  if (d.rc0 && (rc = act_map(&p.a[2], d.res0, d.rc0))) return rc;
  if (d.rc1 && (rc = act_map(&p.a[3], d.res1, d.rc1))) return rc;
  if (p.imgs_per_tile == 0) {
    // (8x8: a tile of whole images never lies inside one image's rows -1 .. h, the whole-tile box is unused there)
    const int at_rows = p.rt + 2 <= d.h_in + 2 ? p.rt + 2 : d.h_in + 2;
    if ((rc = act_map(&p.at[0], d.src0, d.c0, at_rows))) return rc;
    if (d.c1 && (rc = act_map(&p.at[1], d.src1, d.c1, at_rows))) return rc;
    if (d.rc0 && (rc = act_map(&p.at[2], d.res0, d.rc0, at_rows))) return rc;
    if (d.rc1 && (rc = act_map(&p.at[3], d.res1, d.rc1, at_rows))) return rc;
  }
  {
    const uint64_t ktot = 9ull * (d.c0 + d.c1) + d.rc0 + d.rc1;
    uint64_t dims[2] = {ktot, (uint64_t)d.cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64u, (uint32_t)kHaloBN};
    if ((rc = encode_map(&p.b, d.weight, 2, dims, strides, box))) return rc;
    uint32_t box_half[2] = {64u, (uint32_t)kHaloBN / 2};
    if ((rc = encode_map(&p.b_half, d.weight, 2, dims, strides, box_half))) return rc;
  }
  if (d.w_in == 32) return d.cout == 128 ? launch_halo<32, 128>(p, stream) : launch_halo<32, 256>(p, stream);
  if (d.w_in == 16) return d.cout == 128 ? launch_halo<16, 128>(p, stream) : launch_halo<16, 256>(p, stream);
  if (d.w_in == 8) return d.cout == 128 ? launch_halo<8, 128>(p, stream) : launch_halo<8, 256>(p, stream);
  return d.cout == 128 ? launch_halo<4, 128>(p, stream) : launch_halo<4, 256>(p, stream);
}

}  // namespace dmme

// A/B measurement switch: 0 = AUTO never uses the halo kernel, 1 = default
extern "C" void dmme_set_conv_halo_mode(int mode) { dmme::g_halo_mode = mode; }
extern "C" int dmme_get_conv_halo_mode(void) { return dmme::g_halo_mode; }
// A/B measurement switch: 0 = default, no weight multicast (independent CTAs), 1 = clusters of two share the weight stream
// when every CTA has several work items, 2 = pairs for every launch (tests)
extern "C" void dmme_set_conv_halo_multicast(int mode) { dmme::g_halo_mc = mode; }
// debugging: int64[7 * 256] device buffer receiving CTA 0's per-role timestamps (tools/trace_halo.py), null = off
extern "C" void dmme_debug_set_halo_trace(long long* buf) { dmme::g_halo_trace = buf; }
