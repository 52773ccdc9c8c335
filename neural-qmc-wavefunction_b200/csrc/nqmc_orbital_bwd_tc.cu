// nqmc_orbital_bwd_tc.cu -- tensor-core (tcgen05 / TMEM, 3xTF32) reverse pass of the orbital networks for
// hidden width 64.  Same math, interface and slab layout as orb_bwd_kernel<64,...> (nqmc_orbital.cu), which
// stays as the FP32 FFMA parity anchor.
//
// Per 128-row tile (128 consecutive walkers, one electron of the network's spin), with seed = weight * inv:
//     h1 = tanh(f W1 + b1)                      FP32 FFMA, thread-per-row (K = 3+2A is tiny)
//     U2 = h1 W2                                UMMA   M=128 rows, N=64, K=64     (A,B K-major)
//     h2 = tanh(U2 + b2); dW3 += seed h2; delta2 = seed W3 (1-h2^2)
//     D1 = delta2 W2^T                          UMMA   M=128 rows, N=64, K=64     (B = the same W2 buffer, MN-major)
//     G2 += [h1 | 1]^T delta2                   UMMA   M=128 (64 units + a ones row -> db2), N=64, K=128 rows
//     delta1 = D1 (1-h1^2)
//     G1 += [f | 1]^T delta1                    UMMA   M=128 (F features + a ones row -> db1), N=64, K=128 rows
// G2 / G1 stay in tensor memory for the whole kernel and are read back once.
//
// Operand staging: a buffer [k/4][row][k%4] (16-byte granules) is at the same time
//   * a K-major  no-swizzle UMMA operand with M/N = row : LBO = 2048 B (next k/4 group), SBO = 128 B (next 8 rows)
//   * an MN-major no-swizzle UMMA operand with K   = row : SBO = 2048 B (next 4 M/N),    LBO = 128 B (next 8 rows)
// so h1 and delta2/delta1 are split into tf32 hi/lo and stored ONCE, then consumed by two GEMMs each.
#include <cstdlib>

#include "nqmc_internal.cuh"

namespace {

constexpr int H = 64;
constexpr int NQ = 4;                 // threads per row
constexpr int TCW = 128 * NQ;         // worker threads
constexpr int TCT = TCW + 32;         // + MMA issuer warp
constexpr int ROWS = 128;
constexpr int UPT = H / NQ;           // hidden units / output columns per thread = 16
constexpr uint32_t TMEM_COLS = 256;   // U2 | D1 | G2 | G1, 64 columns each
constexpr int GRP = ROWS * 16;        // one k/4 group: 128 rows x 16 B = 2048
constexpr int HB = 17 * GRP;          // h1 hi (or lo): 16 groups + the ones group            = 34816
constexpr int FB = 9 * GRP;           // features hi (or lo): up to 35 features + ones         = 18432
constexpr int DB = 16 * GRP;          // delta hi (or lo)                                     = 32768
constexpr int WB = (H / 4) * H * 16;  // W2 hi (or lo): [k/4][j][k%4]                         = 16384
constexpr int OFF_HH = 0, OFF_HL = HB, OFF_FH = 2 * HB, OFF_FL = 2 * HB + FB, OFF_DH = 2 * HB + 2 * FB,
              OFF_DL = OFF_DH + DB, OFF_WH = OFF_DL + DB, OFF_WL = OFF_WH + WB, OFF_MISC = OFF_WL + WB;
constexpr int MISC_FLOATS = 35 * H + 3 * H + 3 * H + 8;   // W1, b1, b2, W3, reduction scratch
constexpr int SMEM_BYTES = OFF_MISC + MISC_FLOATS * 4 + 64;

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// D=f32, A=B=tf32, N=64, M=128; bit 15 / 16: A / B is MN-major
constexpr uint32_t IDESC_KK = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
constexpr uint32_t IDESC_KM = IDESC_KK | (1u << 16);
constexpr uint32_t IDESC_MM = IDESC_KK | (1u << 15) | (1u << 16);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                     uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, l;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 l, %5, 0;\n\t"
      "@l tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// D (+)= A B with 3xTF32: hi.hi + lo.hi + hi.lo
__device__ __forceinline__ void umma3(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc,
                                      uint32_t accumulate, uint32_t leader) {
  umma(d, ah, bh, idesc, accumulate, leader);
  umma(d, al, bh, idesc, 1u, leader);
  umma(d, ah, bl, idesc, 1u, leader);
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
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  lo = a - hi;
}
// 16 consecutive k of one row -> 4 granules of the hi and lo buffers
__device__ __forceinline__ void store16(uint8_t* hi_base, uint8_t* lo_base, const float (&v)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float4 h4, l4;
    split_tf32(v[4 * g + 0], h4.x, l4.x);
    split_tf32(v[4 * g + 1], h4.y, l4.y);
    split_tf32(v[4 * g + 2], h4.z, l4.z);
    split_tf32(v[4 * g + 3], h4.w, l4.w);
    *reinterpret_cast<float4*>(hi_base + g * GRP) = h4;
    *reinterpret_cast<float4*>(lo_base + g * GRP) = l4;
  }
}

__global__ void __launch_bounds__(TCT, 1)
    orb_bwd_tc_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r,
                      const float* __restrict__ inv, const float* __restrict__ wgt, const double* __restrict__ hdr,
                      const float* __restrict__ e_l, const int64_t W, float* __restrict__ part, const int n_chunks,
                      const int p_mlp) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* misc = reinterpret_cast<float*>(smem + OFF_MISC);
  float* W1s = misc;                 // [F][H]
  float* b1s = W1s + 35 * H;
  float* b2s = b1s + H;
  float* W3s = b2s + H;
  float* redS = W3s + H;             // [3H + 8]: dW3[H], (spare), db3
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc + MISC_FLOATS);   // 3 mbarriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (ROWS - 1), quarter = (tid >> 7) & (NQ - 1);
  const bool worker = tid < TCW;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];
  const int F = s.F, A = s.A;
  const int n_fgrp = (F + 1 + 3) / 4;            // feature granules including the ones column

  // ---- setup ----------------------------------------------------------------------------------------------
  for (int idx = tid; idx < F * H; idx += TCT) W1s[idx] = theta[off.W1 + idx];
  for (int q = tid; q < H; q += TCT) {
    b1s[q] = theta[off.b1 + q];
    b2s[q] = theta[off.b2 + q];
    W3s[q] = theta[off.W3 + q];
  }
  for (int q = tid; q < 3 * H + 8; q += TCT) redS[q] = 0.f;
  for (int idx = tid; idx < H * H; idx += TCT) {           // W2 -> [k/4][j][k%4], hi / lo
    const int k = idx / H, j = idx % H;
    float hi, lo;
    split_tf32(theta[off.W2 + idx], hi, lo);
    const int o = (k >> 2) * (H * 4) + j * 4 + (k & 3);
    reinterpret_cast<float*>(smem + OFF_WH)[o] = hi;
    reinterpret_cast<float*>(smem + OFF_WL)[o] = lo;
  }
  // zero the operand buffers once (unused feature granules / rows must hold finite values), then the ones group of h1
  for (int idx = tid; idx < (OFF_WH) / 16; idx += TCT) reinterpret_cast<float4*>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  if (tid < ROWS) *reinterpret_cast<float4*>(smem + OFF_HH + 16 * GRP + tid * 16) = make_float4(1.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    for (int b = 0; b < 3; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[b])));
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
  const uint32_t T_U2 = tmem_base, T_D1 = tmem_base + 64, T_G2 = tmem_base + 128, T_G1 = tmem_base + 192;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_u = smem_u32(&bars[0]), bar_d = smem_u32(&bars[1]), bar_g = smem_u32(&bars[2]);
  uint32_t ph_u = 0, ph_d = 0, ph_g = 0;
  float centerf = 0.f;
  if (hdr != nullptr) centerf = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));

  float gW3[UPT];
#pragma unroll
  for (int j = 0; j < UPT; ++j) gW3[j] = 0.f;
  float gb3 = 0.f;
  const int c0 = UPT * quarter;                     // first hidden unit / output column of this thread
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  uint8_t* my_h_hi = smem + OFF_HH + (c0 / 4) * GRP + row * 16;
  uint8_t* my_h_lo = smem + OFF_HL + (c0 / 4) * GRP + row * 16;
  uint8_t* my_d_hi = smem + OFF_DH + (c0 / 4) * GRP + row * 16;
  uint8_t* my_d_lo = smem + OFF_DL + (c0 / 4) * GRP + row * 16;

  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;
  bool first_tile = true;
  for (int64_t item = chunk; item < n_items; item += n_chunks) {
    const int i = (int)(item % ns);
    const int64_t w0 = (item / ns) * ROWS;
    float h1[UPT], seed = 0.f;
    if (worker) {
      // ---- positions, seed, features, layer 1 (FP32 FFMA) --------------------------------------------------
      int64_t w = w0 + row;
      const bool valid = w < W;
      if (!valid) w = W - 1;
      const int e = elec0 + i;
      const float x = r[(int64_t)(3 * e + 0) * W + w], y = r[(int64_t)(3 * e + 1) * W + w], z = r[(int64_t)(3 * e + 2) * W + w];
      float wt = 1.0f;
      if (wgt != nullptr) wt = wgt[w];
      else if (e_l != nullptr) {
        const float ev = e_l[w];
        wt = isfinite(ev) ? ev - centerf : 0.f;
      }
      seed = wt * inv[(int64_t)ent_index(s, spin, i, korb) * W + w];
      if (!valid || !isfinite(seed)) seed = 0.f;
      // the previous tile's G2 / G1 MMAs still read the h1 / feature / delta buffers
      if (!first_tile) { mbar_wait(bar_g, ph_g); ph_g ^= 1; }
#pragma unroll
      for (int j = 0; j < UPT; ++j) h1[j] = b1s[c0 + j];
      float feat[36];                                  // indexed only with compile-time-unrollable loops below
      const float pc[3] = {x, y, z};
#pragma unroll
      for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int v = 0; v < UPT / 4; ++v) {
          const float4 wv = *reinterpret_cast<const float4*>(W1s + d * H + c0 + 4 * v);
          h1[4 * v + 0] = fmaf(pc[d], wv.x, h1[4 * v + 0]);
          h1[4 * v + 1] = fmaf(pc[d], wv.y, h1[4 * v + 1]);
          h1[4 * v + 2] = fmaf(pc[d], wv.z, h1[4 * v + 2]);
          h1[4 * v + 3] = fmaf(pc[d], wv.w, h1[4 * v + 3]);
        }
      }
      // feature granules are written by the thread whose quarter matches granule % 4; coordinates first
      float* f_hi = reinterpret_cast<float*>(smem + OFF_FH) + row * 4;
      float* f_lo = reinterpret_cast<float*>(smem + OFF_FL) + row * 4;
      auto put_feature = [&](int f, float val) {         // feature f of this row -> granule f/4, slot f%4
        if (((f >> 2) & (NQ - 1)) == quarter) {
          float hi, lo;
          split_tf32(val, hi, lo);
          f_hi[(f >> 2) * (GRP / 4) + (f & 3)] = hi;
          f_lo[(f >> 2) * (GRP / 4) + (f & 3)] = lo;
        }
      };
      put_feature(0, x);
      put_feature(1, y);
      put_feature(2, z);
      for (int a = 0; a < A; ++a) {
        const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
        const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        const float ex = __expf(-dd);
        put_feature(3 + a, dd);
        put_feature(3 + A + a, ex);
#pragma unroll
        for (int v = 0; v < UPT / 4; ++v) {
          const float4 wd = *reinterpret_cast<const float4*>(W1s + (3 + a) * H + c0 + 4 * v);
          const float4 we = *reinterpret_cast<const float4*>(W1s + (3 + A + a) * H + c0 + 4 * v);
          h1[4 * v + 0] = fmaf(dd, wd.x, fmaf(ex, we.x, h1[4 * v + 0]));
          h1[4 * v + 1] = fmaf(dd, wd.y, fmaf(ex, we.y, h1[4 * v + 1]));
          h1[4 * v + 2] = fmaf(dd, wd.z, fmaf(ex, we.z, h1[4 * v + 2]));
          h1[4 * v + 3] = fmaf(dd, wd.w, fmaf(ex, we.w, h1[4 * v + 3]));
        }
      }
      put_feature(F, 1.0f);                             // ones column -> db1
      (void)feat;
#pragma unroll
      for (int j = 0; j < UPT; ++j) h1[j] = nq_tanh(h1[j]);
      store16(my_h_hi, my_h_lo, h1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- U2 = h1 W2 ------------------------------------------------------------------------------------------
    if (warp == TCW / 32) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t leader = elect_one();
      const uint64_t ah = umma_desc(sbase + OFF_HH, GRP, 128), al = umma_desc(sbase + OFF_HL, GRP, 128);
      const uint64_t bh = umma_desc(sbase + OFF_WH, H * 16, 128), bl = umma_desc(sbase + OFF_WL, H * 16, 128);
#pragma unroll
      for (int ks = 0; ks < H / 8; ++ks) {
        const uint64_t ao = (uint64_t)((ks * 2 * GRP) >> 4), bo = (uint64_t)((ks * 2 * H * 16) >> 4);
        umma3(T_U2, ah + ao, al + ao, bh + bo, bl + bo, IDESC_KK, ks == 0 ? 0u : 1u, leader);
      }
      umma_commit(bar_u, leader);
      __syncwarp();
    }
    // ---- h2, dW3, delta2 ---------------------------------------------------------------------------------------
    if (worker) {
      mbar_wait(bar_u, ph_u);
      ph_u ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[16];
      tmem_ld16(T_U2 + lane_base + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        const float h = nq_tanh(v[j] + b2s[c0 + j]);
        gW3[j] = fmaf(seed, h, gW3[j]);
        v[j] = seed * W3s[c0 + j] * fmaf(-h, h, 1.0f);
      }
      if (quarter == 0) gb3 += seed;
      store16(my_d_hi, my_d_lo, v);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- D1 = delta2 W2^T ; G2 += [h1|1]^T delta2 ------------------------------------------------------------------
    if (warp == TCW / 32) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t leader = elect_one();
      {
        const uint64_t ah = umma_desc(sbase + OFF_DH, GRP, 128), al = umma_desc(sbase + OFF_DL, GRP, 128);
        // B[n = k][kk = j] = W2[k][j]: MN-major view of the W2 buffer: SBO = 1024 B (next 4 k), LBO = 128 B (next 8 j)
        const uint64_t bh = umma_desc(sbase + OFF_WH, 128, H * 16), bl = umma_desc(sbase + OFF_WL, 128, H * 16);
#pragma unroll
        for (int ks = 0; ks < H / 8; ++ks) {
          const uint64_t ao = (uint64_t)((ks * 2 * GRP) >> 4), bo = (uint64_t)((ks * 128) >> 4);
          umma3(T_D1, ah + ao, al + ao, bh + bo, bl + bo, IDESC_KM, ks == 0 ? 0u : 1u, leader);
        }
      }
      {
        // A[m = unit][kk = row] (MN-major view of the h1 buffer), B[n = j][kk = row] (MN-major view of the delta buffer)
        const uint64_t ah = umma_desc(sbase + OFF_HH, 128, GRP), al = umma_desc(sbase + OFF_HL, 128, GRP);
        const uint64_t bh = umma_desc(sbase + OFF_DH, 128, GRP), bl = umma_desc(sbase + OFF_DL, 128, GRP);
#pragma unroll
        for (int ks = 0; ks < ROWS / 8; ++ks) {
          const uint64_t o = (uint64_t)((ks * 128) >> 4);
          umma3(T_G2, ah + o, al + o, bh + o, bl + o, IDESC_MM, (first_tile && ks == 0) ? 0u : 1u, leader);
        }
      }
      umma_commit(bar_d, leader);
      __syncwarp();
    }
    // ---- delta1 = D1 (1 - h1^2) -----------------------------------------------------------------------------------
    if (worker) {
      mbar_wait(bar_d, ph_d);
      ph_d ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[16];
      tmem_ld16(T_D1 + lane_base + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < UPT; ++j) v[j] *= fmaf(-h1[j], h1[j], 1.0f);
      store16(my_d_hi, my_d_lo, v);                     // delta2 is dead: both of its GEMMs have completed
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- G1 += [f|1]^T delta1 --------------------------------------------------------------------------------------
    if (warp == TCW / 32) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t leader = elect_one();
      const uint64_t ah = umma_desc(sbase + OFF_FH, 128, GRP), al = umma_desc(sbase + OFF_FL, 128, GRP);
      const uint64_t bh = umma_desc(sbase + OFF_DH, 128, GRP), bl = umma_desc(sbase + OFF_DL, 128, GRP);
#pragma unroll
      for (int ks = 0; ks < ROWS / 8; ++ks) {
        const uint64_t o = (uint64_t)((ks * 128) >> 4);
        umma3(T_G1, ah + o, al + o, bh + o, bl + o, IDESC_MM, (first_tile && ks == 0) ? 0u : 1u, leader);
      }
      umma_commit(bar_g, leader);
      __syncwarp();
    }
    first_tile = false;
  }

  // ---- flush: G2 -> dW2 (+db2), G1 -> dW1 (+db1), register partials -> dW3, db3 --------------------------------------
  float* slab = part + (int64_t)blockIdx.x * p_mlp;
  const int o_b1 = 0, o_W1 = H, o_b2 = H + F * H, o_W2 = o_b2 + H, o_b3 = o_W2 + H * H, o_W3 = o_b3 + 1;
  if (worker) {
    if (!first_tile) { mbar_wait(bar_g, ph_g); ph_g ^= 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // thread (row = TMEM lane, quarter): 16 columns of lane `row` of G2 and G1
    float v[16];
    if (!first_tile) {
      tmem_ld16(T_G2 + lane_base + (uint32_t)c0, v);
      if (row < H) {
#pragma unroll
        for (int j = 0; j < UPT; ++j) slab[o_W2 + row * H + c0 + j] = v[j];
      } else if (row == H) {
#pragma unroll
        for (int j = 0; j < UPT; ++j) slab[o_b2 + c0 + j] = v[j];
      }
      tmem_ld16(T_G1 + lane_base + (uint32_t)c0, v);
      if (row < F) {
#pragma unroll
        for (int j = 0; j < UPT; ++j) slab[o_W1 + row * H + c0 + j] = v[j];
      } else if (row == F) {
#pragma unroll
        for (int j = 0; j < UPT; ++j) slab[o_b1 + c0 + j] = v[j];
      }
    } else {
      // this CTA had no tile: its slab must still be defined
      if (row < H) for (int j = 0; j < UPT; ++j) slab[o_W2 + row * H + c0 + j] = 0.f;
      if (row == H) for (int j = 0; j < UPT; ++j) slab[o_b2 + c0 + j] = 0.f;
      if (row < F) for (int j = 0; j < UPT; ++j) slab[o_W1 + row * H + c0 + j] = 0.f;
      if (row == F) for (int j = 0; j < UPT; ++j) slab[o_b1 + c0 + j] = 0.f;
    }
    // dW3 / db3: reduce the per-row partials over the 128 rows
#pragma unroll
    for (int j = 0; j < UPT; ++j) {
      float t = warp_sum(gW3[j]);
      if ((tid & 31) == 0) atomicAdd(&redS[c0 + j], t);
    }
    if (quarter == 0) {
      float t = warp_sum(gb3);
      if ((tid & 31) == 0) atomicAdd(&redS[3 * H], t);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  for (int q = tid; q < H; q += TCT) slab[o_W3 + q] = redS[q];
  if (tid == 0) slab[o_b3] = redS[3 * H];
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

}  // namespace

// returns NQMC_ERR_UNSUPPORTED when this variant does not apply (the caller then uses the FFMA kernel).
// On success the per-CTA slabs are in ctx->part_mlp, *n_chunks_out CTAs per network.
int nq_launch_orb_bwd_tc(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                         int* n_chunks_out, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H != H || s.F > 34) return NQMC_ERR_UNSUPPORTED;
  if (getenv("NQMC_TC_BWD") == nullptr) return NQMC_ERR_UNSUPPORTED;   // work in progress: off unless requested
  static bool configured = false;
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(orb_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks * s.n_mlp > ctx->n_cta_bwd) n_chunks = ctx->n_cta_bwd / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  {
    ProfScope ps(ctx, NQ_PROF_BWD, st);
    orb_bwd_tc_kernel<<<n_chunks * s.n_mlp, TCT, SMEM_BYTES, st>>>(s, ctx->theta, r, ctx->inv, wgt, hdr, e_l, W, ctx->part_mlp,
                                                                   n_chunks, ctx->p_mlp);
  }
  NQMC_LAUNCH_CHECK();
  *n_chunks_out = n_chunks;
  return NQMC_OK;
}
