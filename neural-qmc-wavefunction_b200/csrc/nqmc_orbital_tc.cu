// nqmc_orbital_tc.cu -- tensor-core (tcgen05 / TMEM) variant of the 5-channel orbital forward pass
// for hidden width 64: the H x H layer runs as 3xTF32 UMMAs with fp32 accumulators in tensor memory.
//
// Same math and same interface as orb_eval_kernel<64,5,...> (nqmc_orbital.cu), which stays as the FFMA
// parity anchor.  Why 3xTF32: log|psi| / E_L tolerances need fp32-equivalent operands (SURVEY Appendix D:
// plain TF32 misses by ~100x).  a = a_hi + a_lo with a_hi = a & 0xffffe000 (exactly representable in
// TF32) and a_lo = a - a_hi (exact in fp32, truncated by the tensor core to its top 11 bits), likewise b;
//     a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo        (dropped terms ~2^-22 |a.b|)
//
// Tile = 128 evaluation rows (128 consecutive walkers, one electron of the network's spin) of ONE orbital
// network; CTA = 256 threads = 8 warps, one CTA per SM (TMEM: 5 accumulators x 64 columns).
//   thread (row = tid & 127, half = tid >> 7)
//   layer 1 (K = 3+2A <= 35, FP32 FFMA): the thread evaluates its row's features and their gradient /
//       Laplacian in registers and computes 16 hidden units x 5 channels per pass from broadcast weight
//       loads (20 FFMA per LDS.128); tanh rules; hi/lo split; one 16-byte store per (channel, 4 hidden units)
//       straight into the UMMA K-major core-matrix layout [k/4][row][k%4] (conflict-free, no swizzle).
//   layer 2 (64 x 64, tensor cores): per pass the two halves produce K-chunks {2h+p}; one thread issues
//       5 channels x 2 chunks x 2 k-steps x 3 products = 60 tcgen05.mma (M=128, N=64, K=8, kind::tf32) and
//       commits to an mbarrier; the next pass's layer-1 FFMA work overlaps the MMAs.
//   epilogue: tcgen05.ld 32x32b -> bias, tanh rules for the 5 channels (thread-local), dot with W3,
//       the two halves of a row are combined through shared memory; phi[c][ent][w] written coalesced.
#include "nqmc_internal.cuh"

namespace {

constexpr int H = 64;
constexpr int NQ = 4;                // threads per row (each owns 64/NQ output columns and 32/NQ layer-1 units per pass)
constexpr int TCW = 128 * NQ;        // worker threads: layer 1, operand staging, epilogue
constexpr int TCT = TCW + 32;       // + one warp whose lane 0 issues the UMMAs
constexpr int ROWS = 128;           // rows per tile == UMMA M
constexpr int KC = 16;              // hidden units per K-chunk
constexpr int UPT = 32 / NQ;        // layer-1 hidden units per thread per pass
constexpr int EPC = H / NQ;         // epilogue columns per thread
constexpr int CH = 5;               // value, d/dx, d/dy, d/dz, Laplacian
constexpr uint32_t TMEM_COLS = 512;
// operand buffers (bytes)
constexpr int A_K4 = ROWS * 16;                 // one k/4 group: 128 rows x 16 B             = 2048
constexpr int A_CHUNK = (KC / 4) * A_K4;        // one (channel, hi|lo) K-chunk               = 8192
constexpr int A_BUF = CH * 2 * A_CHUNK;         // one K-chunk, all channels, hi+lo           = 81920
constexpr int B_K4 = H * 16;                    // one k/4 group of W2^T: 64 n x 16 B         = 1024
constexpr int B_MAT = (H / 4) * B_K4;           // W2 hi (or lo)                              = 16384
constexpr int SMEM_BYTES = 2 * A_BUF + 2 * B_MAT + (35 * H + 3 * H) * 4 + (NQ - 1) * ROWS * CH * 4 + 64;

// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// issued from warp-uniform code by the whole issuer warp; only the elected lane (leader != 0) executes the MMA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate,
                                          uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, l;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 l, %5, 0;\n\t"
      "@l tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(leader)
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

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  lo = a - hi;
}

// layer 1 (FP32 FFMA) for hidden units j0 .. j0+15 of one row, 5 channels, followed by the tanh rules
__device__ __forceinline__ void layer1_pass(const DevSys& s, const float* __restrict__ W1s, const float* __restrict__ b1s,
                                            const float x, const float y, const float z, const int j0,
                                            float (&acc)[CH][UPT]) {
  const int A = s.A;
#pragma unroll
  for (int j = 0; j < UPT; ++j) {
    acc[0][j] = b1s[j0 + j];
#pragma unroll
    for (int c = 1; c < CH; ++c) acc[c][j] = 0.f;
  }
  {  // coordinate features: value p_d, gradient e_d, Laplacian 0
    const float pc[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
      for (int v = 0; v < UPT / 4; ++v) {
        const float4 wv = *reinterpret_cast<const float4*>(W1s + d * H + j0 + 4 * v);
        const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc[0][4 * v + t] = fmaf(pc[d], wa[t], acc[0][4 * v + t]);
          acc[1 + d][4 * v + t] += wa[t];
        }
      }
    }
  }
  for (int a = 0; a < A; ++a) {
    const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
    const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    const float ex = __expf(-dd), inv = 1.0f / dd;
    const float ux = dx * inv, uy = dy * inv, uz = dz * inv;
    const float fd[CH] = {dd, ux, uy, uz, 2.0f * inv};
    const float fe[CH] = {ex, -ex * ux, -ex * uy, -ex * uz, ex * (1.0f - 2.0f * inv)};
#pragma unroll
    for (int v = 0; v < UPT / 4; ++v) {
      const float4 wd = *reinterpret_cast<const float4*>(W1s + (3 + a) * H + j0 + 4 * v);
      const float4 we = *reinterpret_cast<const float4*>(W1s + (3 + A + a) * H + j0 + 4 * v);
      const float wda[4] = {wd.x, wd.y, wd.z, wd.w}, wea[4] = {we.x, we.y, we.z, we.w};
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c][4 * v + t] = fmaf(fd[c], wda[t], fmaf(fe[c], wea[t], acc[c][4 * v + t]));
    }
  }
#pragma unroll
  for (int j = 0; j < UPT; ++j) {   // tanh rules, thread-local across the 5 channels
    const float h = nq_tanh(acc[0][j]);
    const float tp = fmaf(-h, h, 1.0f), tpp = -2.0f * h * tp;
    const float g2 = fmaf(acc[1][j], acc[1][j], fmaf(acc[2][j], acc[2][j], acc[3][j] * acc[3][j]));
    acc[4][j] = fmaf(tp, acc[4][j], tpp * g2);
    acc[1][j] *= tp; acc[2][j] *= tp; acc[3][j] *= tp;
    acc[0][j] = h;
  }
}

// hi/lo split and store as UMMA K-major core matrices [k/4][row][k%4]; then make the writes visible to the async proxy
__device__ __forceinline__ void store_operand(uint8_t* buf, const float (&acc)[CH][UPT]) {
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int v = 0; v < UPT / 4; ++v) {
      float4 hi4, lo4;
      split_tf32(acc[c][4 * v + 0], hi4.x, lo4.x);
      split_tf32(acc[c][4 * v + 1], hi4.y, lo4.y);
      split_tf32(acc[c][4 * v + 2], hi4.z, lo4.z);
      split_tf32(acc[c][4 * v + 3], hi4.w, lo4.w);
      *reinterpret_cast<float4*>(buf + (c * 2 + 0) * A_CHUNK + v * A_K4) = hi4;
      *reinterpret_cast<float4*>(buf + (c * 2 + 1) * A_CHUNK + v * A_K4) = lo4;
    }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// one thread: 5 channels x 2 K-chunks x 2 k-steps x 3 products, then commit to `bar`
template <int p>
__device__ __forceinline__ void issue_pass( const uint32_t tmem_base, const uint32_t a_base,
                                           const uint32_t bhi_base, const uint32_t blo_base, const uint32_t bar) {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t leader = elect_one();
  // descriptors differ only in their 14-bit start-address field: precompute the bases once
  const uint64_t dA = umma_desc(a_base, A_K4, 128), dBh = umma_desc(bhi_base, B_K4, 128), dBl = umma_desc(blo_base, B_K4, 128);
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int kchunk = 2 * hh + p;                       // hidden units 32*hh + 16*p ..
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const uint32_t d_tmem = tmem_base + (uint32_t)(c * H);
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks) {
        const uint64_t dah = dA + (uint64_t)((hh * A_BUF + (c * 2 + 0) * A_CHUNK + ks * 2 * A_K4) >> 4);
        const uint64_t dal = dA + (uint64_t)((hh * A_BUF + (c * 2 + 1) * A_CHUNK + ks * 2 * A_K4) >> 4);
        const uint64_t kb = (uint64_t)(((kchunk * (KC / 4) + ks * 2) * B_K4) >> 4);
        const uint64_t dbh = dBh + kb, dbl = dBl + kb;
        const uint32_t first = (p == 0 && hh == 0 && ks == 0) ? 0u : 1u;
        umma_tf32(d_tmem, dah, dbh, first, leader);
        umma_tf32(d_tmem, dal, dbh, 1u, leader);
        umma_tf32(d_tmem, dah, dbl, 1u, leader);
      }
    }
  }
  umma_commit(bar, leader);
  __syncwarp();
}

__global__ void __launch_bounds__(TCT, 1)
    orb_eval5_tc_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r, const int64_t W,
                        float* __restrict__ phi, const int n_chunks) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* Abuf = smem_raw;                                   // [2 halves][CH][hi,lo][KC/4][ROWS][4] fp32
  uint8_t* Bhi = Abuf + 2 * A_BUF;                            // [H/4][H n][4]
  uint8_t* Blo = Bhi + B_MAT;
  float* W1s = reinterpret_cast<float*>(Blo + B_MAT);         // [F][H]
  float* b1s = W1s + 35 * H;
  float* b2s = b1s + H;
  float* W3s = b2s + H;
  float* outS = W3s + H;                                      // [NQ-1][ROWS][CH] partial outputs of quarters 1..
  uint64_t* bars = reinterpret_cast<uint64_t*>(outS + (NQ - 1) * ROWS * CH);   // 2 mbarriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (ROWS - 1), quarter = (tid >> 7) & (NQ - 1);
  // layer-1 ownership: K-chunk buffer hh = quarter / (NQ/2), units 16*p + UPT*(quarter % (NQ/2)) inside it
  const int hh_own = quarter / (NQ / 2), sub_own = quarter % (NQ / 2);
  const bool worker = tid < TCW;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];
  const int F = s.F;

  // ---- one-time setup: weights, barriers, tensor memory -----------------------------------------
  for (int idx = tid; idx < F * H; idx += TCT) W1s[idx] = theta[off.W1 + idx];
  for (int q = tid; q < H; q += TCT) {
    b1s[q] = theta[off.b1 + q];
    b2s[q] = theta[off.b2 + q];
    W3s[q] = theta[off.W3 + q];
  }
  // B operand: B[n][k] = W2[k][n], K-major core matrices: float offset (k/4)*256 + n*4 + (k%4)
  for (int idx = tid; idx < H * H; idx += TCT) {
    const int k = idx / H, n = idx % H;
    float hi, lo;
    split_tf32(theta[off.W2 + idx], hi, lo);
    const int o = (k >> 2) * (H * 4) + n * 4 + (k & 3);
    reinterpret_cast<float*>(Bhi)[o] = hi;
    reinterpret_cast<float*>(Blo)[o] = lo;
  }
  const float b3 = theta[off.b3];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // B operand written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_base = smem_u32(Abuf), bhi_base = smem_u32(Bhi), blo_base = smem_u32(Blo);
  const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
  uint32_t phase0 = 0, phase1 = 0;
  uint8_t* my_buf = Abuf + hh_own * A_BUF + sub_own * (UPT / 4) * A_K4 + row * 16;
  const int j_own = 32 * hh_own + UPT * sub_own;             // + KC * pass

  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;
  // position of this thread's row in tile `it`
  auto load_xyz = [&](int64_t it, float& x, float& y, float& z) {
    int64_t w = (it / ns) * ROWS + row;
    if (w >= W) w = W - 1;
    const int e = elec0 + (int)(it % ns);
    x = r[(int64_t)(3 * e + 0) * W + w];
    y = r[(int64_t)(3 * e + 1) * W + w];
    z = r[(int64_t)(3 * e + 2) * W + w];
  };

  // software pipeline: pass 0 of tile t+1 is computed while the tensor cores finish tile t
  float acc[CH][UPT];
  float x = 0.f, y = 0.f, z = 0.f;
  if (worker && chunk < n_items) {
    load_xyz(chunk, x, y, z);
    layer1_pass(s, W1s, b1s, x, y, z, j_own, acc);
  }
  for (int64_t item = chunk; item < n_items; item += n_chunks) {
    const int i = (int)(item % ns);
    const int64_t w0 = (item / ns) * ROWS;
    const int64_t next = item + n_chunks;
    float xn = 0.f, yn = 0.f, zn = 0.f;
    if (worker && next < n_items) load_xyz(next, xn, yn, zn);        // prefetch: consumed after the second MMA batch
    // ---- pass 0: stage operands, start the MMAs --------------------------------------------------------
    if (worker) store_operand(my_buf, acc);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TCW / 32) issue_pass<0>(tmem_base, a_base, bhi_base, blo_base, bar0);
    // ---- pass 1: layer 1 overlaps the pass-0 MMAs ---------------------------------------------------------
    if (worker) {
      layer1_pass(s, W1s, b1s, x, y, z, j_own + KC, acc);
      mbar_wait(bar0, phase0);                                       // operand buffers free again
      phase0 ^= 1;
      store_operand(my_buf, acc);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TCW / 32) issue_pass<1>(tmem_base, a_base, bhi_base, blo_base, bar1);
    // ---- pass 0 of the next tile overlaps the pass-1 MMAs ---------------------------------------------------
    float out[CH] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (worker) {
      if (next < n_items) layer1_pass(s, W1s, b1s, xn, yn, zn, j_own, acc);
      x = xn; y = yn; z = zn;
      // ---- epilogue: accumulators -> activation rules -> output layer ------------------------------------
      mbar_wait(bar1, phase1);
      phase1 ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
#pragma unroll 1
      for (int sub = 0; sub < EPC / 8; ++sub) {
        const int c0 = EPC * quarter + 8 * sub;                     // first hidden unit of this 8-column slice
        float v[CH][8];
#pragma unroll
        for (int c = 0; c < CH; ++c) tmem_ld8(tmem_base + lane_base + (uint32_t)(c * H + c0), v[c]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = nq_tanh(v[0][j] + b2s[c0 + j]);
          const float tp = fmaf(-h, h, 1.0f), tpp = -2.0f * h * tp;
          const float g2 = fmaf(v[1][j], v[1][j], fmaf(v[2][j], v[2][j], v[3][j] * v[3][j]));
          const float w3 = W3s[c0 + j], w3tp = w3 * tp;
          out[0] = fmaf(w3, h, out[0]);
          out[1] = fmaf(w3tp, v[1][j], out[1]);
          out[2] = fmaf(w3tp, v[2][j], out[2]);
          out[3] = fmaf(w3tp, v[3][j], out[3]);
          out[4] = fmaf(w3, fmaf(tp, v[4][j], tpp * g2), out[4]);
        }
      }
      if (quarter != 0) {
#pragma unroll
        for (int c = 0; c < CH; ++c) outS[((quarter - 1) * ROWS + row) * CH + c] = out[c];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                          // accumulators are free; half-1 partials visible
    if (worker && quarter == 0) {
      const int64_t wr = w0 + row;
      if (wr < W) {
        const int ent = ent_index(s, spin, i, korb);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          float o = out[c] + (c == 0 ? b3 : 0.f);
#pragma unroll
          for (int q = 0; q < NQ - 1; ++q) o += outS[(q * ROWS + row) * CH + c];
          phi[((int64_t)c * s.n_ent + ent) * W + wr] = o;
        }
      }
    }
    // outS is rewritten only after the next tile's two barriers, so no extra barrier is needed here
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

}  // namespace

// returns NQMC_ERR_UNSUPPORTED when this variant does not apply (the caller then uses the FFMA kernel)
int nq_launch_orb_eval5_tc(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H != H || s.F > 35) return NQMC_ERR_UNSUPPORTED;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(orb_eval5_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  {
    ProfScope ps(ctx, NQ_PROF_EVAL5, st);
    orb_eval5_tc_kernel<<<n_chunks * s.n_mlp, TCT, SMEM_BYTES, st>>>(s, ctx->theta, r, W, phi, n_chunks);
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
