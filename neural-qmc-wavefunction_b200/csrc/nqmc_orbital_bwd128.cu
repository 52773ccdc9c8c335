// nqmc_orbital_bwd128.cu -- tensor-core reverse pass of the orbital networks for hidden width 128 (BASELINE configs 3
// and 5).  Same math, interface and slab layout as orb_bwd_kernel<128,...> (nqmc_orbital.cu, the FP32 FFMA parity anchor).
// Reference lines replaced: optimizer/vmc.py:120-133 (vmap(grad(log|psi|)) and the weighted mean).
//
// Why three kernels instead of the single fused one that serves width 64 (nqmc_orbital_bwd_tc.cu): at H = 128 the two
// tf32 hi/lo copies of W2 the fused kernel keeps (W2 for U2 = h1 W2, W2^T for D1 = delta2 W2^T; kind::tf32 has no
// operand format that reads one copy both ways) are 256 KB, and the K = rows operands of the weight-gradient GEMM
// (h1^T, delta2^T, hi and lo) another 256 KB per 128-row tile -- against 227 KB of shared memory.  So the sweep is
// cut where the operands change, and the per-row activations travel through HBM once, unit-major and row-contiguous
// (hT / dT / eT [network][unit][row], fT [network][feature][row]; 1.7 KB per row, coalesced on both sides):
//   1. orb_fwd_pc_kernel<..., BW> (nqmc_orbital_fwd_pc.cu): layer 1 -> h1 -> U2 = h1 W2 (tcgen05, A in TMEM) ->
//      h2, dW3 += seed h2, db3 += seed, delta2 = seed W3 (1 - h2^2); stores hT, dT, fT.          smem: W2 (128 KB)
//   2. bwd128_d1_kernel (here): D1 = delta2 W2^T (tcgen05, A = delta2 tile loaded back into TMEM), delta1 = D1 (1 - h1^2);
//      stores eT.  Two accumulators: the epilogue of tile t overlaps the MMAs of tile t+1.        smem: W2^T (128 KB)
//   3. bwd128_gram_kernel (here): the weight gradients as one K = rows contraction per network, split over CTAs:
//        G2 += hT dT^T (dW2, 128 x 128) ; G1 += eT fT^T (dW1 and db1, 128 x 32) ; GB += dT fT[0:16]^T (db2, column 0).
//      The arrays are K-major GEMM operands as stored, so TMA (SWIZZLE_128B) streams 32-row stages into shared memory,
//      8 splitter warps form tf32 hi / lo tiles in place, one thread issues the 3xTF32 MMAs; every 8 stages the
//      accumulator is swapped and flushed into registers (the tensor core's fp32 accumulation truncates).
// All three write the per-CTA slab [b1 | W1 | b2 | W2 | b3 | W3] that orb_bwd_reduce_kernel sums in fp64.
#include <cstdlib>

#include "nqmc_internal.cuh"
#include "nqmc_tma.cuh"

namespace {
using namespace nqtma;

constexpr int H = 128, ROWS = 128, NQ = 4;
constexpr int NFEAT = 32;                        // rows of fT per network: [1 | F features | zero padding]

// =====================================================================================================================
// 2. D1 = delta2 W2^T, delta1 = D1 (1 - h1^2)
// =====================================================================================================================
constexpr int D1_WORKERS = ROWS * NQ;            // 16 warps, thread = (row, quarter): 32 units of its row
constexpr int D1_THREADS = D1_WORKERS + 32;      // + the issuing warp
constexpr int UPT = H / NQ;                      // 32
constexpr int B_K4 = H * 16;                     // bytes between K-groups of 4 of the no-swizzle K-major B operand (LBO)
constexpr int B_MAT = (H / 4) * B_K4;            // 64 KB per hi / lo copy
constexpr uint32_t C_D = 0, C_AH = 256, C_AL = 384;
constexpr int D1_SMEM = 2 * B_MAT + 256;

__device__ __forceinline__ uint64_t desc_k_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(D1_THREADS, 1)
    bwd128_d1_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ hT, const float* __restrict__ dT,
                     float* __restrict__ eT, const int64_t rows_pad, const int64_t nwb, const int n_chunks) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* Bhi = smem;                           // B[n = h1 unit k][kk = delta2 unit j] = W2[k][j] : [j/4][k][j%4]
  uint8_t* Blo = smem + B_MAT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * B_MAT);   // a_full, mma_done[2], d_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (ROWS - 1), quarter = (tid >> 7) & (NQ - 1);
  const bool worker = tid < D1_WORKERS;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int ns = m < s.n_up ? s.n_up : s.n_dn;
  const MlpOff off = s.mlp[m];

  for (int idx = tid; idx < H * H; idx += D1_THREADS) {
    const int k = idx / H, j = idx % H;
    const float w = theta[off.W2 + idx];
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const int o = (j >> 2) * (H * 4) + k * 4 + (j & 3);
    reinterpret_cast<float*>(Bhi)[o] = hi;
    reinterpret_cast<float*>(Blo)[o] = w - hi;
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), D1_WORKERS / 32);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    mbar_init(smem_u32(&bars[3]), D1_WORKERS / 32);
    mbar_init(smem_u32(&bars[4]), D1_WORKERS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_full = smem_u32(&bars[0]), mma_done = smem_u32(&bars[1]), d_empty = smem_u32(&bars[3]);
  const int64_t n_items = nwb * ns;
  nq_pdl_wait();                                 // hT / dT come from the kernel in front

  if (!worker) {
    // ---- MMA issuer ------------------------------------------------------------------------------------------------
    if (elect_one()) {
      constexpr uint32_t IDESC = idesc_tf32(ROWS, H);
      const uint64_t dBh = desc_k_none(smem_u32(Bhi), B_K4, 128), dBl = desc_k_none(smem_u32(Blo), B_K4, 128);
      int n = 0;
      for (int64_t item = chunk; item < n_items; item += n_chunks, ++n) {
        const int b = n & 1;
        if (n >= 2) mbar_wait(d_empty + 8 * b, ((n >> 1) - 1) & 1);   // accumulator b has been read
        mbar_wait(a_full, n & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + C_D + (uint32_t)(b * H);
#pragma unroll 4
        for (int ks = 0; ks < H / 8; ++ks) {
          const uint64_t kb = (uint64_t)(((uint32_t)ks * 2 * B_K4) >> 4);
          const uint32_t ah = tmem_base + C_AH + 8u * ks, al = tmem_base + C_AL + 8u * ks;
          umma_ts(d, ah, dBh + kb, IDESC, ks == 0 ? 0u : 1u);
          umma_ts(d, al, dBh + kb, IDESC, 1u);
          umma_ts(d, ah, dBl + kb, IDESC, 1u);
        }
        umma_commit(mma_done + 8 * b);
      }
    }
  } else {
    // ---- workers ---------------------------------------------------------------------------------------------------
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int c0 = UPT * quarter;
    const float* dsrc = dT + ((int64_t)m * H + c0) * rows_pad;
    const float* hsrc = hT + ((int64_t)m * H + c0) * rows_pad;
    float* edst = eT + ((int64_t)m * H + c0) * rows_pad;
    float dcur[UPT];
    auto load_d = [&](const int64_t item) {
      const int64_t rid = item * ROWS + row;
#pragma unroll
      for (int j = 0; j < UPT; ++j) dcur[j] = dsrc[(int64_t)j * rows_pad + rid];
    };
    auto epilogue = [&](const int64_t item, const int n) {      // tile n of this CTA (accumulator n & 1)
      const int b = n & 1;
      const int64_t rid = item * ROWS + row;
      float hv[UPT];
#pragma unroll
      for (int j = 0; j < UPT; ++j) hv[j] = hsrc[(int64_t)j * rows_pad + rid];
      mbar_wait(mma_done + 8 * b, (n >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int sub = 0; sub < UPT / 16; ++sub) {
        float v[16];
        tmem_ld16(tmem_base + lane_base + C_D + (uint32_t)(b * H + c0 + 16 * sub), v);
        if (sub == UPT / 16 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty + 8 * b);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float h = hv[16 * sub + j];
          edst[(int64_t)(16 * sub + j) * rows_pad + rid] = v[j] * fmaf(-h, h, 1.0f);
        }
      }
    };
    if (chunk < n_items) load_d(chunk);
    int n = 0;
    int64_t prev = -1;
    for (int64_t item = chunk; item < n_items; item += n_chunks, ++n) {
      if (n > 0) {                                   // the MMAs of tile n-1 have read the A columns
        mbar_wait(mma_done + 8 * ((n - 1) & 1), ((n - 1) >> 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int sub = 0; sub < UPT / 16; ++sub) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = dcur[16 * sub + j];
          hi[j] = __float_as_uint(a) & 0xFFFFE000u;
          lo[j] = __float_as_uint(a - __uint_as_float(hi[j]));
        }
        tmem_st16(tmem_base + lane_base + C_AH + (uint32_t)(c0 + 16 * sub), hi);
        tmem_st16(tmem_base + lane_base + C_AL + (uint32_t)(c0 + 16 * sub), lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      if (item + n_chunks < n_items) load_d(item + n_chunks);   // in flight while the previous tile's epilogue runs
      if (n > 0) epilogue(prev, n - 1);
      prev = item;
    }
    if (n > 0) epilogue(prev, n - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// 3. weight gradients: G2 += hT dT^T, G1 += eT fT^T, GB += dT fT[0:16]^T over this CTA's share of the network's rows
// =====================================================================================================================
constexpr int STAGES = 2, FLUSH = 8;
constexpr int T_BIG = H * BK * 4;                // 16 KB: one 128-unit x 32-row tile
constexpr int T_F = NFEAT * BK * 4;              //  4 KB
constexpr int RAW_BYTES = 3 * T_BIG + T_F;       // hT | dT | eT | fT
constexpr int STAGE_BYTES = 2 * RAW_BYTES;       // raw/hi tiles, then the lo tiles in the same order
constexpr int NSPLIT_WARPS = 8, NSPLIT = NSPLIT_WARPS * 32;
constexpr int G_THREADS = NSPLIT + 64;
constexpr int G_SMEM = STAGES * STAGE_BYTES + 1024 + 256;
constexpr uint32_t C_G2 = 0, C_G1 = 128, C_GB = 160, ACC_COLS = 256;   // per accumulator set (168 used)

__global__ void __launch_bounds__(G_THREADS, 1)
    bwd128_gram_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmD,
                       const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmF, const DevSys s,
                       const int64_t nwb, const int n_chunks, float* __restrict__ part, const int p_mlp) {
  extern __shared__ uint8_t smem_raw[];
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int ns = m < s.n_up ? s.n_up : s.n_dn;
  const int F = s.F;
  // this CTA's stages: the network's n_rows / 32 stages dealt evenly over the chunks
  const int64_t n_st = nwb * ns * (ROWS / BK);
  const int64_t per = (n_st + n_chunks - 1) / n_chunks;
  const int64_t st_begin = (int64_t)chunk * per;
  const int n_it = (int)(st_begin >= n_st ? 0 : (st_begin + per <= n_st ? per : n_st - st_begin));
  const int n_blk = (n_it + FLUSH - 1) / FLUSH;
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);   // raw[ST], split[ST], empty[ST], full[2], free[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&bars[i]), 1);
      mbar_init(smem_u32(&bars[STAGES + i]), NSPLIT_WARPS);
      mbar_init(smem_u32(&bars[2 * STAGES + i]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars[3 * STAGES + b]), 1);
      mbar_init(smem_u32(&bars[3 * STAGES + 2 + b]), NSPLIT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmF);
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 2 * ACC_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t0 = smem_u32(tiles);
  const uint32_t bar_raw = smem_u32(&bars[0]), bar_split = smem_u32(&bars[STAGES]), bar_empty = smem_u32(&bars[2 * STAGES]);
  const uint32_t bar_full = smem_u32(&bars[3 * STAGES]), bar_free = smem_u32(&bars[3 * STAGES + 2]);
  nq_pdl_wait();                                 // the activations come from the two kernels in front

  if (warp == NSPLIT_WARPS) {
    if (elect_one()) {
      for (int it = 0; it < n_it; ++it) {
        const int sg = it % STAGES, u = it / STAGES;
        if (u > 0) mbar_wait(bar_empty + 8 * sg, (u - 1) & 1);
        const uint32_t st = t0 + sg * STAGE_BYTES, bar = bar_raw + 8 * sg;
        mbar_arrive_expect_tx(bar, RAW_BYTES);
        const int k = (int)((st_begin + it) * BK);
        tma_load_3d(st, &tmH, k, 0, m, bar);
        tma_load_3d(st + T_BIG, &tmD, k, 0, m, bar);
        tma_load_3d(st + 2 * T_BIG, &tmE, k, 0, m, bar);
        tma_load_3d(st + 3 * T_BIG, &tmF, k, 0, m, bar);
      }
    }
  } else if (warp == NSPLIT_WARPS + 1) {
    if (elect_one()) {
      constexpr uint32_t ID_G2 = idesc_tf32(H, H), ID_G1 = idesc_tf32(H, NFEAT), ID_GB = idesc_tf32(H, 16);   // M = 128 needs N % 16 == 0: 16 feature columns, column 0 (ones) is used
      for (int it = 0; it < n_it; ++it) {
        const int sg = it % STAGES, u = it / STAGES;
        const int blk = it / FLUSH, b = blk & 1;
        const bool first = it % FLUSH == 0;
        if (first && blk >= 2) mbar_wait(bar_free + 8 * b, ((blk >> 1) - 1) & 1);
        mbar_wait(bar_split + 8 * sg, u & 1);
        tc_fence_after();
        const uint32_t st = t0 + sg * STAGE_BYTES, lo = st + RAW_BYTES, acc = tmem_base + (uint32_t)(b * ACC_COLS);
        const uint64_t hh = desc_k_sw128(st), dh = desc_k_sw128(st + T_BIG), eh = desc_k_sw128(st + 2 * T_BIG), fh = desc_k_sw128(st + 3 * T_BIG);
        const uint64_t hl = desc_k_sw128(lo), dl = desc_k_sw128(lo + T_BIG), el = desc_k_sw128(lo + 2 * T_BIG), fl = desc_k_sw128(lo + 3 * T_BIG);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t o = (uint64_t)(ks * 2);
          const uint32_t accum = (!first || ks > 0) ? 1u : 0u;
          umma_ss(acc + C_G2, hh + o, dh + o, ID_G2, accum);
          umma_ss(acc + C_G2, hl + o, dh + o, ID_G2, 1u);
          umma_ss(acc + C_G2, hh + o, dl + o, ID_G2, 1u);
          umma_ss(acc + C_G1, eh + o, fh + o, ID_G1, accum);
          umma_ss(acc + C_G1, el + o, fh + o, ID_G1, 1u);
          umma_ss(acc + C_G1, eh + o, fl + o, ID_G1, 1u);
          umma_ss(acc + C_GB, dh + o, fh + o, ID_GB, accum);
          umma_ss(acc + C_GB, dl + o, fh + o, ID_GB, 1u);
          umma_ss(acc + C_GB, dh + o, fl + o, ID_GB, 1u);
        }
        umma_commit(bar_empty + 8 * sg);
        if (it % FLUSH == FLUSH - 1 || it == n_it - 1) umma_commit(bar_full + 8 * b);
      }
    }
  } else {
    // splitter warps: thread (q = lane quadrant, half): lane = unit 32 q + lane; G2 columns [64 half, +64), G1 columns
    // [16 half, +16), GB columns 0..7 (half 0 only)
    const int q = warp & 3, half = warp >> 2;
    const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16);
    float g2[64], g1[16], gb[8];
#pragma unroll
    for (int j = 0; j < 64; ++j) g2[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) g1[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) gb[j] = 0.f;
    auto flush = [&](const int blk) {
      const int b = blk & 1;
      mbar_wait(bar_full + 8 * b, (blk >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = trow + (uint32_t)(b * ACC_COLS);
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        float v[16];
        tmem_ld16(acc + C_G2 + (uint32_t)(64 * half + c), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) g2[c + j] += v[j];
      }
      {
        float v[16];
        tmem_ld16(acc + C_G1 + (uint32_t)(16 * half), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) g1[j] += v[j];
      }
      if (half == 0) {
        float v[8];
        tmem_ld8(acc + C_GB, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) gb[j] += v[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 8 * b);
    };
    int flushed = 0;
    for (int it = 0; it < n_it; ++it) {
      const int sg = it % STAGES, u = it / STAGES;
      mbar_wait(bar_raw + 8 * sg, u & 1);
      uint8_t* st = tiles + sg * STAGE_BYTES;
      split_tile_inplace(st, st + RAW_BYTES, RAW_BYTES / 16, tid, NSPLIT);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_split + 8 * sg);
      if (it % FLUSH == STAGES && it / FLUSH > flushed) flush(flushed++);
    }
    for (; flushed < n_blk; ++flushed) flush(flushed);
    // slab order: b1[H] W1[F*H] b2[H] W2[H*H] b3 W3[H]   (b3 / W3 were written by the value sweep)
    float* slab = part + (int64_t)blockIdx.x * p_mlp;
    const int o_b1 = 0, o_W1 = H, o_b2 = H + F * H, o_W2 = o_b2 + H;
    const int unit = 32 * q + lane;
#pragma unroll
    for (int j = 0; j < 64; ++j) slab[o_W2 + unit * H + 64 * half + j] = g2[j];     // G2[h1 unit][delta2 unit]
#pragma unroll
    for (int j = 0; j < 16; ++j) {                                                    // G1[delta1 unit][1 | feature]
      const int f = 16 * half + j;
      if (f == 0) slab[o_b1 + unit] = g1[j];
      else if (f - 1 < F) slab[o_W1 + (f - 1) * H + unit] = g1[j];
    }
    if (half == 0) slab[o_b2 + unit] = gb[0];                                         // GB[delta2 unit][ones]
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 2 * ACC_COLS);
}

// =====================================================================================================================
// 1'. seeds from parked activations.  When the 5-channel energy pass of the same fused step has left h1 (hT), h2 and the
// features (fT) of every row (nqmc_orbital_fwd_pc.cu, BwdOut), the value sweep (kernel 1) is not repeated: this HBM-bound
// kernel reads h2, forms seed = weight * inverse-matrix entry per row, writes delta2 = seed W3 (1 - h2^2) as dT and sums
// dW3 += seed h2, db3 += seed.  The row sums go through a padded shared-memory tile: written [row][unit] by the
// thread-per-row loaders (bank = row + unit), read back by one thread per unit.
// =====================================================================================================================
constexpr int SEED_G = 8, SEED_T = SEED_G * ROWS, SEED_U = H / SEED_G, SEED_LD = H + 1;   // thread = (row, group of 16 units)
constexpr int SEED_SMEM = (ROWS * SEED_LD + ROWS + H + SEED_G * (H + 1)) * 4;

__global__ void __launch_bounds__(SEED_T, 1)
    bwd128_seed_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ h2T, float* __restrict__ dT,
                       const float* __restrict__ inv, const double* __restrict__ hdr, const float* __restrict__ e_l,
                       const int64_t W, const int64_t rows_pad, const int64_t nwb, const int n_chunks,
                       float* __restrict__ part, const int p_mlp) {
  extern __shared__ float seed_sm[];
  float* tile = seed_sm;                         // [ROWS][SEED_LD] seed * h2
  float* sseed = tile + ROWS * SEED_LD;          // [ROWS]
  float* sW3 = sseed + ROWS;                     // [H]
  float* sred = sW3 + H;                         // [SEED_G][H + 1] partial sums of the row groups
  const int tid = threadIdx.x, row = tid & (ROWS - 1), grp = tid >> 7;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1, korb = spin == 0 ? m : m - s.n_up, ns = spin == 0 ? s.n_up : s.n_dn;
  const MlpOff off = s.mlp[m];
  if (tid < H) sW3[tid] = theta[off.W3 + tid];
  const float center = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));
  __syncthreads();
  // second role in the summing phase: thread (unit = tid % 128, part = tid / 128) adds rows 16 part .. 16 part + 15
  float acc = 0.f, accb = 0.f;
  const int64_t n_items = nwb * ns;
  ItemIter cur(chunk, n_chunks, ns);
  for (int64_t item = chunk; item < n_items; item += n_chunks, cur.next()) {
    const int64_t w = (int64_t)cur.wb * ROWS + row;
    const int i = cur.e;
    const int64_t base = ((int64_t)m * H + SEED_U * grp) * rows_pad + item * ROWS + row;
    float h2[SEED_U];
#pragma unroll
    for (int uu = 0; uu < SEED_U; ++uu) h2[uu] = __ldcs(h2T + base + (int64_t)uu * rows_pad);   // all 16 loads in flight
    float seed = 0.f;
    if (w < W) {
      const float ev = e_l[w];
      const float wt = isfinite(ev) ? ev - center : 0.f;
      seed = wt * inv[(int64_t)ent_index(s, spin, i, korb) * W + w];
      if (!isfinite(seed)) seed = 0.f;
    }
    if (grp == 0) sseed[row] = seed;
#pragma unroll
    for (int uu = 0; uu < SEED_U; ++uu) {
      dT[base + (int64_t)uu * rows_pad] = seed * sW3[SEED_U * grp + uu] * fmaf(-h2[uu], h2[uu], 1.f);
      tile[row * SEED_LD + SEED_U * grp + uu] = seed * h2[uu];
    }
    __syncthreads();
    {
      const int u = tid & (H - 1), part16 = (tid >> 7) * (ROWS / SEED_G);
      float a = 0.f;
#pragma unroll
      for (int rr = 0; rr < ROWS / SEED_G; ++rr) a += tile[(part16 + rr) * SEED_LD + u];
      acc += a;
      if (tid < ROWS) accb += sseed[tid];        // thread tid keeps row tid's seeds; summed over the rows at the end
    }
    __syncthreads();
  }
  sred[grp * (H + 1) + (tid & (H - 1))] = acc;
  if (tid < ROWS) tile[tid] = accb;
  __syncthreads();
  float* slab = part + (int64_t)blockIdx.x * p_mlp;
  const int o_b3 = 2 * H + s.F * H + H * H, o_W3 = o_b3 + 1;   // slab order: b1[H] W1[F*H] b2[H] W2[H*H] b3 W3[H]
  if (tid < H) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < SEED_G; ++g) a += sred[g * (H + 1) + tid];
    slab[o_W3 + tid] = a;
  } else if (tid == H) {
    float a = 0.f;
    for (int rr = 0; rr < ROWS; ++rr) a += tile[rr];
    slab[o_b3] = a;
  }
}

}  // namespace

int nq_launch_orb_fwd_bw128(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                            float* hT, float* dT, float* fT, int64_t rows_pad, int n_chunks, cudaStream_t st);

// scratch rows per network for W walkers
int64_t nq_bwd128_rows(const nqmc_ctx* ctx, int64_t W) {
  const DevSys& s = ctx->sys;
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  return nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn) * ROWS;
}

int nq_launch_orb_bwd128(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                         int* n_chunks_out, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H != H || s.F + 1 > NFEAT || ctx->bw_h == nullptr) return NQMC_ERR_UNSUPPORTED;
  if (const char* e = getenv("NQMC_TC_BWD"))
    if (atoi(e) == 0) return NQMC_ERR_UNSUPPORTED;
  if (s.A != 1 && s.A != 2 && s.A != 4 && s.A != 6 && s.A != 8 && s.A != 10) return NQMC_ERR_UNSUPPORTED;
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t rows_pad = ctx->bw_rows;
  if (nq_bwd128_rows(ctx, W) > rows_pad) return NQMC_ERR_STATE;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  if (max_items >= (int64_t)1 << 31) return NQMC_ERR_UNSUPPORTED;   // ItemIter counts items in 32 bits
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks * s.n_mlp > ctx->n_cta_bwd) n_chunks = ctx->n_cta_bwd / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  *n_chunks_out = n_chunks;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(bwd128_d1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D1_SMEM));
    NQMC_CUDA(cudaFuncSetAttribute(bwd128_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
    NQMC_CUDA(cudaFuncSetAttribute(bwd128_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEED_SMEM));
    configured = true;
  }
  // activations parked by the energy pass of this very step (same positions, same theta, E_L-centred weights)?
  const bool reuse = wgt == nullptr && hdr != nullptr && e_l != nullptr && ctx->act_valid_for == e_l && ctx->act_W == W &&
                     ctx->bw_h2 != nullptr;
  ctx->act_valid_for = nullptr;
  if (reuse) {
    ProfScope ps(ctx, NQ_PROF_BWD, st);
    bwd128_seed_kernel<<<n_chunks * s.n_mlp, SEED_T, SEED_SMEM, st>>>(s, ctx->theta, ctx->bw_h2, ctx->bw_d, ctx->inv, hdr, e_l, W, rows_pad, nwb,
                                                                       n_chunks, ctx->part_mlp, ctx->p_mlp);
    NQMC_LAUNCH_CHECK();
  } else {
    int rc = nq_launch_orb_fwd_bw128(ctx, r, wgt, hdr, e_l, W, ctx->bw_h, ctx->bw_d, ctx->bw_f, rows_pad, n_chunks, st);
    if (rc != NQMC_OK) return rc;
  }
  {
    ProfScope ps(ctx, NQ_PROF_BWD, st);
    if (ctx->prof_on)
      bwd128_d1_kernel<<<n_chunks * s.n_mlp, D1_THREADS, D1_SMEM, st>>>(s, ctx->theta, ctx->bw_h, ctx->bw_d, ctx->bw_e, rows_pad, nwb, n_chunks);
    else
      NQMC_CUDA(nq_launch_pdl_guarded(ctx, bwd128_d1_kernel, dim3(n_chunks * s.n_mlp), dim3(D1_THREADS), (size_t)D1_SMEM, st, s, ctx->theta,
                                      ctx->bw_h, ctx->bw_d, ctx->bw_e, rows_pad, nwb, n_chunks));
  }
  NQMC_LAUNCH_CHECK();
  CUtensorMap tmH, tmD, tmE, tmF;
  const uint64_t nm = (uint64_t)s.n_mlp, rp = (uint64_t)rows_pad;
  if (make_map_k32(&tmH, ctx->bw_h, rp, H, nm, rp, rp * H, H) || make_map_k32(&tmD, ctx->bw_d, rp, H, nm, rp, rp * H, H) ||
      make_map_k32(&tmE, ctx->bw_e, rp, H, nm, rp, rp * H, H) || make_map_k32(&tmF, ctx->bw_f, rp, NFEAT, nm, rp, rp * NFEAT, NFEAT)) {
    ctx->err = "orb_bwd128: cuTensorMapEncodeTiled failed";
    return NQMC_ERR_CUDA;
  }
  {
    ProfScope ps(ctx, NQ_PROF_BWD, st);
    bwd128_gram_kernel<<<n_chunks * s.n_mlp, G_THREADS, G_SMEM, st>>>(tmH, tmD, tmE, tmF, s, nwb, n_chunks, ctx->part_mlp, ctx->p_mlp);
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
