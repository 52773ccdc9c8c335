// nqmc_orbital_val2_tc.cu -- value-only orbital forward pass (MH proposal, src/nqmc/sampler/metropolis.py:124) with
// BOTH dense layers on the tensor cores (tcgen05 / TMEM, 3xTF32), hidden width 64.  Same math and interface as
// orb_eval1_tc_kernel (nqmc_orbital_val_tc.cu, layer 1 on the FP32 pipe) and orb_eval_kernel<64,1,...> (the FFMA
// parity anchor).  The value pass is what the reference's burn-in / thinning steps consist of, so it dominates a
// real VMC iteration (10 value-only MH steps per sampled step with the reference defaults, vmc.py:247).
//
// Why: in the FP32-layer-1 kernel two thirds of a thread's instructions are the 11 x 16 FFMAs of layer 1 and their
// weight loads.  Here a row's features [x, y, z, 1, (d_A, exp(-d_A)) pairs] (K1 = 16 columns; the ones column
// carries b1) are written as tf32 hi/lo into TENSOR MEMORY by the four threads of the row (one nucleus each) and
// layer 1 becomes 6 UMMAs (M=128, N=64, K=8); a thread is left with about 330 instructions per tile instead of 830:
//     W_a  features of tile t+2            -> FA[b]   (tcgen05.st)           | issuer: D1[b] = FA[b] x [W1;b1]
//     W_b  h1 = tanh(D1) of tile t+1       -> A2[b] hi/lo (tcgen05.ld / st)  | issuer: U2[b] = A2[b] x W2 (aliases D1[b])
//     W_c  phi = W3 . tanh(U2 + b2) + b3 of tile t (tcgen05.ld, 4-quarter combine through shared memory)
// One step = W_a, W_b, W_c for three consecutive tiles, then ONE barrier; the UMMAs issued after the barrier of step s
// are consumed in step s+1, behind the other two work items, so nobody waits for the tensor pipe.
// TMEM columns: FA 2 x 32, D1/U2 2 x 64, A2 2 x 128 = 448.
#include <cstdlib>

#include "nqmc_internal.cuh"

namespace {

constexpr int H = 64;
constexpr int NQ = 4;
constexpr int TCW = 128 * NQ;         // worker threads: (row, quarter), 16 hidden units / output columns each
constexpr int TCT = TCW + 128;        // + issuer warpgroup (two issuing warps; 640 threads -> 96 registers each, plenty here)
constexpr int NWW = TCW / 32;
constexpr int NSYNC = TCW + 64;       // workers + the two issuing warps
constexpr int ROWS = 128;
constexpr int UPT = H / NQ;           // 16
constexpr int K1 = 16;                // layer-1 K: x, y, z, 1, (d_a, exp(-d_a)) pairs, padded (4 + 2 A <= 16  ->  A <= 6)
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t C_FA = 0, C_D = 64, C_A2 = 192;     // FA[b] = C_FA + 32 b (hi 16 | lo 16); D1/U2[b] = C_D + 64 b; A2[b] = C_A2 + 128 b (hi 64 | lo 64)
constexpr int W2B = (H / 4) * H * 16;  // one W2 copy (hi or lo), K-major no swizzle [k/4][n][k%4]: 16 KB
constexpr int W1B = (K1 / 4) * H * 16; // one [W1;b1] copy: 4 KB

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, l;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 l, %5, 0;\n\t"
      "@l tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint32_t bar, uint32_t leader) {
  asm volatile("{\n\t.reg .pred l;\n\tsetp.ne.b32 l, %1, 0;\n\t"
               "@l tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(leader)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void split_tf32(float a, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(a) & 0xFFFFE000u;
  lo = __float_as_uint(a - __uint_as_float(hi));
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(TCT, 1)
    orb_eval1_l1tc_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r, const int64_t W,
                          float* __restrict__ phi, const int n_chunks, const int order) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* W2hi = smem;                                  // [k/4][n][k%4]
  uint8_t* W2lo = W2hi + W2B;
  uint8_t* W1hi = W2lo + W2B;                            // [k/4][n][k%4], k = feature (row F: b1, rows > F: 0)
  uint8_t* W1lo = W1hi + W1B;
  float* b2s = reinterpret_cast<float*>(W1lo + W1B);
  float* W3s = b2s + H;
  float* outS = W3s + H;                                 // [2 parities][NQ-1][ROWS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(outS + 2 * (NQ - 1) * ROWS);   // bar_l1[2], bar_u[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (ROWS - 1), quarter = (tid >> 7) & (NQ - 1);
  const bool worker = tid < TCW;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];
  const int A = s.A;

  // ---- setup: weights to tf32 hi/lo operand layouts (reads theta only) ------------------------------------------
  for (int q = tid; q < H; q += TCT) {
    b2s[q] = theta[off.b2 + q];
    W3s[q] = theta[off.W3 + q];
  }
  {
    float wv[H * H / TCW];
#pragma unroll
    for (int i = 0; i < H * H / TCW; ++i) wv[i] = worker ? theta[off.W2 + tid + i * TCW] : 0.f;
#pragma unroll
    for (int i = 0; i < H * H / TCW; ++i) {
      if (!worker) break;
      const int idx = tid + i * TCW, k = idx / H, n = idx % H;
      uint32_t hi, lo;
      split_tf32(wv[i], hi, lo);
      const int o = (k >> 2) * (H * 4) + n * 4 + (k & 3);
      reinterpret_cast<uint32_t*>(W2hi)[o] = hi;
      reinterpret_cast<uint32_t*>(W2lo)[o] = lo;
    }
  }
  for (int idx = tid; idx < K1 * H; idx += TCT) {
    const int k = idx / H, n = idx % H;
    // feature order of the operand: x, y, z, 1, then (d_a, exp(-d_a)) pairs -- each thread's columns are contiguous
    float v = 0.f;
    if (k < 3) v = theta[off.W1 + k * H + n];
    else if (k == 3) v = theta[off.b1 + n];
    else if (k < 4 + 2 * A) v = theta[off.W1 + (((k - 4) & 1) ? 3 + A + ((k - 4) >> 1) : 3 + ((k - 4) >> 1)) * H + n];
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    const int o = (k >> 2) * (H * 4) + n * 4 + (k & 3);
    reinterpret_cast<uint32_t*>(W1hi)[o] = hi;
    reinterpret_cast<uint32_t*>(W1lo)[o] = lo;
  }
  const float b3 = theta[off.b3];
  if (tid == 0) {
    for (int b = 0; b < 4; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bar_l1[2] = {smem_u32(&bars[0]), smem_u32(&bars[1])}, bar_u[2] = {smem_u32(&bars[2]), smem_u32(&bars[3])};
  auto tile_barrier = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NSYNC) : "memory"); };

  nq_pdl_wait();                              // everything above read theta only; the positions come from the predecessor
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;
  const int n_tiles = chunk < n_items ? (int)((n_items - chunk + n_chunks - 1) / n_chunks) : 0;   // tiles of this CTA (the launcher keeps n_items below 2^31)
  // step s = -2 .. n_tiles-1:  W_a(tile s+2), W_b(tile s+1), W_c(tile s) ; barrier ; issue L1(s+2), U2(s+1)

  if (!worker) {
    // =============================== issuer warpgroup ========================================================
    const int iw = warp - NWW;                // 0: layer-1 UMMAs, 1: layer-2 UMMAs, 2-3: idle
    if (iw >= 2) return;
    const uint32_t leader = elect_one();
    const uint64_t b1h = umma_desc(smem_u32(W1hi), H * 16, 128), b1l = umma_desc(smem_u32(W1lo), H * 16, 128);
    const uint64_t b2h = umma_desc(smem_u32(W2hi), H * 16, 128), b2l = umma_desc(smem_u32(W2lo), H * 16, 128);
    for (int st = -2; st < n_tiles; ++st) {
      tile_barrier();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (iw == 0 && st + 2 < n_tiles) {                         // D1 = [features | 1] x [W1 ; b1]
        const uint32_t b = (uint32_t)((st + 2) & 1);
        const uint32_t d = tmem_base + C_D + 64u * b, fa = tmem_base + C_FA + 32u * b;
#pragma unroll
        for (int ks = 0; ks < K1 / 8; ++ks) {
          const uint64_t bo = (uint64_t)((ks * 2 * H * 16) >> 4);
          umma_ts(d, fa + 8 * ks, b1h + bo, ks == 0 ? 0u : 1u, leader);
          umma_ts(d, fa + K1 + 8 * ks, b1h + bo, 1u, leader);
          umma_ts(d, fa + 8 * ks, b1l + bo, 1u, leader);
        }
        umma_commit(bar_l1[b], leader);
      }
      if (iw == 1 && st + 1 >= 0 && st + 1 < n_tiles) {          // U2 = h1 x W2, overwrites D1 of the same tile
        const uint32_t b = (uint32_t)((st + 1) & 1);
        const uint32_t d = tmem_base + C_D + 64u * b, a2 = tmem_base + C_A2 + 128u * b;
#pragma unroll 1
        for (int ks = 0; ks < H / 8; ++ks) {
          const uint64_t bo = (uint64_t)((ks * 2 * H * 16) >> 4);
          umma_ts(d, a2 + 8 * ks, b2h + bo, ks == 0 ? 0u : 1u, leader);
          umma_ts(d, a2 + H + 8 * ks, b2h + bo, 1u, leader);
          umma_ts(d, a2 + 8 * ks, b2l + bo, 1u, leader);
        }
        umma_commit(bar_u[b], leader);
      }
    }
    return;
  }

  // =============================== worker warps ================================================================
  const int c0 = UPT * quarter;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  // zero the feature operands once: padding columns and the lo part of the ones column stay zero for the whole kernel
  {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    tmem_st16(tmem_base + lane_base + C_FA + 16u * quarter, z);   // 4 quarters x 16 columns = FA[0] and FA[1]
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("bar.sync 2, %0;" ::"n"(TCW) : "memory");        // workers only: zeros before the first features
  uint32_t ph_l1[2] = {0u, 0u}, ph_u[2] = {0u, 0u};
  float nx = 0.f, ny = 0.f, nz = 0.f;                           // positions of the tile whose features are written next step
  ItemIter it_f(chunk, n_chunks, ns), it_c = it_f;              // next tile to fetch / tile whose phi is stored next
  auto fetch = [&]() {                                          // tiles are requested in order: 0, 1, 2, ...
    int64_t w = (int64_t)it_f.wb * ROWS + row;
    if (w >= W) w = W - 1;
    const float* p = r + (int64_t)(3 * (elec0 + it_f.e)) * W + w;   // one address, then two strides of W
    nx = p[0];
    ny = p[W];
    nz = p[2 * W];
    it_f.next();
  };
  if (n_tiles > 0) fetch();
#pragma unroll 2
  for (int st = -2; st < n_tiles; ++st) {
    // ---- W_a: features of tile st+2 -> FA[(st+2)&1].  Quarter q takes nuclei q, q+4, ..; quarter 0 also x, y, z, 1.
    // The three work items of a step touch different tiles and are independent.  The four warps that share an SM
    // sub-partition (the four quarters of 32 rows) leave the step barrier together; if all of them run a, b, c they
    // send their MUFU batches (2 per tanh) to the sub-partition's one XU pipe at the same moment and then all run their
    // FP32 work while it idles.  Odd quarters therefore run b, a, c: half of the warps are in the feature stage while
    // the other half use the XU pipe (value kernel 63.7 -> 58 us).
    const int a_pos = (order >> (2 * quarter)) & 3;             // 0: a, b, c   1: b, a, c   2: b, c, a
    auto stage_a = [&]() {
      if (!(st + 2 < n_tiles)) return;
      const float x = nx, y = ny, z = nz;
      if (st + 3 < n_tiles) fetch();
      const uint32_t fa = tmem_base + lane_base + C_FA + 32u * (uint32_t)((st + 2) & 1);
      if (quarter == 0) {                                       // columns 0..3: x, y, z, 1 (the ones column carries b1)
        uint32_t h0, l0, h1, l1, h2, l2;
        split_tf32(x, h0, l0);
        split_tf32(y, h1, l1);
        split_tf32(z, h2, l2);
        tmem_st4(fa, h0, h1, h2, 0x3f800000u);
        tmem_st4(fa + K1, l0, l1, l2, 0u);
      }
      for (int a = quarter; a < A; a += NQ) {                  // columns 4 + 2a, 5 + 2a: d_a, exp(-d_a)
        const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
        const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        uint32_t h0, l0, h1, l1;
        split_tf32(dd, h0, l0);
        split_tf32(__expf(-dd), h1, l1);
        tmem_st2(fa + (uint32_t)(4 + 2 * a), h0, h1);
        tmem_st2(fa + (uint32_t)(K1 + 4 + 2 * a), l0, l1);
      }
    };
    if (a_pos == 0) stage_a();
    // ---- W_b: h1 = tanh(D1) of tile st+1 -> A2[(st+1)&1]
    if (st + 1 >= 0 && st + 1 < n_tiles) {
      const uint32_t b = (uint32_t)((st + 1) & 1);
      mbar_wait(bar_l1[b], ph_l1[b]);
      ph_l1[b] ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[16];
      tmem_ld16(tmem_base + lane_base + C_D + 64u * b + (uint32_t)c0, v);
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < UPT; j += 2) {                         // two units per packed instruction (FMUL2 / FADD2 / FFMA2)
        const float2 h = nq_tanh2(make_float2(v[j], v[j + 1]));
        const float2 hh = make_float2(__uint_as_float(__float_as_uint(h.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(h.y) & 0xFFFFE000u));
        const float2 ll = nq_add2(h, make_float2(-hh.x, -hh.y));
        hi[j] = __float_as_uint(hh.x); hi[j + 1] = __float_as_uint(hh.y);
        lo[j] = __float_as_uint(ll.x); lo[j + 1] = __float_as_uint(ll.y);
      }
      tmem_st16(tmem_base + lane_base + C_A2 + 128u * b + (uint32_t)c0, hi);
      tmem_st16(tmem_base + lane_base + C_A2 + 128u * b + (uint32_t)(H + c0), lo);
    }
    if (a_pos == 1) stage_a();
    // ---- W_c: phi of tile st
    float out = 0.f;
    const uint32_t pb = (uint32_t)(st & 1);
    if (st >= 0) {
      mbar_wait(bar_u[pb], ph_u[pb]);
      ph_u[pb] ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[16];
      tmem_ld16(tmem_base + lane_base + C_D + 64u * pb + (uint32_t)c0, v);
      float2 o2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < UPT; j += 2) {
        const float2 h = nq_tanh2(nq_add2(make_float2(v[j], v[j + 1]), *reinterpret_cast<const float2*>(b2s + c0 + j)));
        o2 = nq_fma2(*reinterpret_cast<const float2*>(W3s + c0 + j), h, o2);
      }
      out = o2.x + o2.y;
      if (quarter != 0) outS[(pb * (NQ - 1) + quarter - 1) * ROWS + row] = out;
    }
    if (a_pos >= 2) stage_a();
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");   // feature and h1 stores of this step
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tile_barrier();        // FA(st+2), A2(st+1) complete; D1/U2 of tile st read; partial outputs visible
    if (st >= 0) {
      if (quarter == 0) {
        const int64_t wr = (int64_t)it_c.wb * ROWS + row;
        if (wr < W) {
          float o = out + b3;
#pragma unroll
          for (int q = 0; q < NQ - 1; ++q) o += outS[(pb * (NQ - 1) + q) * ROWS + row];
          phi[(int64_t)ent_index(s, spin, it_c.e, korb) * W + wr] = o;
        }
      }
      it_c.next();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("bar.sync 2, %0;" ::"n"(TCW) : "memory");        // workers only: every accumulator has been read
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

}  // namespace

int nq_launch_orb_eval1_l1tc(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H != H || s.F + 1 > K1) return NQMC_ERR_UNSUPPORTED;
  static const bool enabled = []() { const char* e = getenv("NQMC_VAL_L1TC"); return e == nullptr || atoi(e) != 0; }();
  if (!enabled) return NQMC_ERR_UNSUPPORTED;
  const size_t smem = 2 * W2B + 2 * W1B + ((size_t)2 * H + 2 * (NQ - 1) * ROWS) * 4 + 64;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(orb_eval1_l1tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  if (max_items >= (int64_t)1 << 31) return NQMC_ERR_UNSUPPORTED;   // ItemIter counts items in 32 bits
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  {
    ProfScope ps(ctx, NQ_PROF_EVAL1, st);
    // stage order per quarter, 2 bits each (see the kernel); default 0x60: quarters 0, 1: a, b, c; quarter 2: b, c, a;
    // quarter 3: b, a, c (scripts/exp/sweep_eval1_order.py: 65.7 us unstaggered, 57.9 - 59.9 us for every staggered setting)
    static const int order = [] { const char* e = getenv("NQMC_EVAL1_ORDER"); return e ? (int)strtol(e, nullptr, 0) : 0x60; }();
    if (ctx->prof_on) orb_eval1_l1tc_kernel<<<n_chunks * s.n_mlp, TCT, smem, st>>>(s, ctx->theta, r, W, phi, n_chunks, order);
    else NQMC_CUDA(nq_launch_pdl_guarded(ctx, orb_eval1_l1tc_kernel, dim3(n_chunks * s.n_mlp), dim3(TCT), smem, st, s, ctx->theta, r, W, phi, n_chunks, order));
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
