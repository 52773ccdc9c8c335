// nqmc_orbital_lap_tc.cu -- 5-channel (value, gradient, Laplacian) orbital forward pass on tcgen05 with the
// layer-2 A operand in TENSOR MEMORY.  Second-generation variant of nqmc_orbital_tc.cu (which stages A in shared
// memory): there the UMMAs re-read 4 KB of shared memory per 32-cycle MMA and the workers issue 80 STS.128 per tile,
// so the shared-memory pipe was ~50 % busy on top of the FFMA work; here the activated layer-1 outputs go from
// registers straight into TMEM (tcgen05.st, lane = row), only W2 is read from shared memory.
//
// Tile = 128 rows; thread (row, quarter).  TMEM: 5 accumulators x 64 columns + one 16-unit K-chunk of the A
// operand (5 channels x 16 columns x {hi, lo}) = 480 columns.  Per tile, four K-chunks:
//     layer 1 (FFMA) for units 16 q .. 16 q + 15 (4 per thread) -> tanh rules -> hi/lo -> tcgen05.st
//     issuer: 5 channels x 2 k-steps x 3 products = 30 UMMAs (A from TMEM), commit
// The A chunk is single-buffered: chunk q+1's layer-1 FFMA work overlaps chunk q's MMAs and only its tcgen05.st
// waits for them.  Per-row distance / exponential features (with gradient and Laplacian) are produced once per tile
// into shared memory and shared by the four threads of a row and the four chunks.
#include "nqmc_internal.cuh"

namespace {

constexpr int H = 64;
constexpr int NQ = 4;
constexpr int TCW = 128 * NQ;
constexpr int TCT = TCW + 32;
constexpr int ROWS = 128;
constexpr int CH = 5;
constexpr int KC = 16;                 // units per K-chunk
constexpr int UPT = KC / NQ;           // layer-1 units per thread per chunk = 4
constexpr int EPC = H / NQ;            // epilogue columns per thread = 16
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t C_AH = 320, C_AL = 400;   // A chunk: hi at 320 + 16 c, lo at 400 + 16 c
constexpr int B_K4 = H * 16;
constexpr int B_MAT = (H / 4) * B_K4;

inline size_t smem_bytes(int F, int A) {
  return 2 * B_MAT + ((size_t)F * H + 3 * H + (NQ - 1) * ROWS * CH + (size_t)A * 2 * CH * ROWS) * 4 + 64;
}

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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void split_tf32(float a, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(a) & 0xFFFFE000u;
  lo = __float_as_uint(a - __uint_as_float(hi));
}

// Fs[(a * 10 + c) * ROWS + row]: c = 0..4 distance feature (value, d/dx, d/dy, d/dz, Laplacian), 5..9 exp(-d) feature
__device__ __forceinline__ void produce_features(const DevSys& s, float* __restrict__ Fs_row, const int quarter,
                                                 const float x, const float y, const float z) {
  for (int a = quarter; a < s.A; a += NQ) {
    const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
    const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    const float ex = __expf(-dd), inv = 1.0f / dd;
    const float ux = dx * inv, uy = dy * inv, uz = dz * inv;
    float* fa = Fs_row + a * (2 * CH) * ROWS;
    fa[0 * ROWS] = dd; fa[1 * ROWS] = ux; fa[2 * ROWS] = uy; fa[3 * ROWS] = uz; fa[4 * ROWS] = 2.0f * inv;
    fa[5 * ROWS] = ex; fa[6 * ROWS] = -ex * ux; fa[7 * ROWS] = -ex * uy; fa[8 * ROWS] = -ex * uz;
    fa[9 * ROWS] = ex * (1.0f - 2.0f * inv);
  }
}

// layer 1 for units j0 .. j0+3 of this row, 5 channels, then the tanh rules
__device__ __forceinline__ void layer1_chunk(const DevSys& s, const float* __restrict__ W1s, const float* __restrict__ b1s,
                                             const float* __restrict__ Fs_row, const float x, const float y,
                                             const float z, const int j0, float (&acc)[CH][UPT]) {
  const int A = s.A;
  {
    const float4 bv = *reinterpret_cast<const float4*>(b1s + j0);
    acc[0][0] = bv.x; acc[0][1] = bv.y; acc[0][2] = bv.z; acc[0][3] = bv.w;
#pragma unroll
    for (int c = 1; c < CH; ++c)
#pragma unroll
      for (int t = 0; t < UPT; ++t) acc[c][t] = 0.f;
  }
  {  // coordinate features: value p_d, gradient e_d, Laplacian 0
    const float pc[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float4 wv = *reinterpret_cast<const float4*>(W1s + d * H + j0);
      const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        acc[0][t] = fmaf(pc[d], wa[t], acc[0][t]);
        acc[1 + d][t] += wa[t];
      }
    }
  }
  for (int a = 0; a < A; ++a) {
    const float* fa = Fs_row + a * (2 * CH) * ROWS;
    float fd[CH], fe[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      fd[c] = fa[c * ROWS];
      fe[c] = fa[(CH + c) * ROWS];
    }
    const float4 wd = *reinterpret_cast<const float4*>(W1s + (3 + a) * H + j0);
    const float4 we = *reinterpret_cast<const float4*>(W1s + (3 + A + a) * H + j0);
    const float wda[4] = {wd.x, wd.y, wd.z, wd.w}, wea[4] = {we.x, we.y, we.z, we.w};
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c][t] = fmaf(fd[c], wda[t], fmaf(fe[c], wea[t], acc[c][t]));
  }
#pragma unroll
  for (int t = 0; t < UPT; ++t) {
    const float h = nq_tanh(acc[0][t]);
    const float tp = fmaf(-h, h, 1.0f), tpp = -2.0f * h * tp;
    const float g2 = fmaf(acc[1][t], acc[1][t], fmaf(acc[2][t], acc[2][t], acc[3][t] * acc[3][t]));
    acc[4][t] = fmaf(tp, acc[4][t], tpp * g2);
    acc[1][t] *= tp; acc[2][t] *= tp; acc[3][t] *= tp;
    acc[0][t] = h;
  }
}

__global__ void __launch_bounds__(TCT, 1)
    orb_lap_tc_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r, const int64_t W,
                      float* __restrict__ phi, const int n_chunks) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* Bhi = smem_raw;                                    // [H/4][H n][4]
  uint8_t* Blo = Bhi + B_MAT;
  float* W1s = reinterpret_cast<float*>(Blo + B_MAT);         // [F][H]
  float* b1s = W1s + s.F * H;
  float* b2s = b1s + H;
  float* W3s = b2s + H;
  float* outS = W3s + H;                                      // [NQ-1][ROWS][CH]
  float* Fs = outS + (NQ - 1) * ROWS * CH;                    // [A][10][ROWS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Fs + s.A * 2 * CH * ROWS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (ROWS - 1), quarter = (tid >> 7) & (NQ - 1);
  const bool worker = tid < TCW;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];
  const int F = s.F;

  for (int idx = tid; idx < F * H; idx += TCT) W1s[idx] = theta[off.W1 + idx];
  for (int q = tid; q < H; q += TCT) {
    b1s[q] = theta[off.b1 + q];
    b2s[q] = theta[off.b2 + q];
    W3s[q] = theta[off.W3 + q];
  }
  for (int idx = tid; idx < H * H; idx += TCT) {
    const int k = idx / H, n = idx % H;
    uint32_t hi, lo;
    split_tf32(theta[off.W2 + idx], hi, lo);
    const int o = (k >> 2) * (H * 4) + n * 4 + (k & 3);
    reinterpret_cast<uint32_t*>(Bhi)[o] = hi;
    reinterpret_cast<uint32_t*>(Blo)[o] = lo;
  }
  const float b3 = theta[off.b3];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
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
  const uint32_t bhi_base = smem_u32(Bhi), blo_base = smem_u32(Blo);
  const uint32_t barA = smem_u32(&bars[0]);
  uint32_t phaseA = 0;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  float* Fs_row = Fs + row;

  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;
  auto load_xyz = [&](int64_t it, float& x, float& y, float& z) {
    int64_t w = (it / ns) * ROWS + row;
    if (w >= W) w = W - 1;
    const int e = elec0 + (int)(it % ns);
    x = r[(int64_t)(3 * e + 0) * W + w];
    y = r[(int64_t)(3 * e + 1) * W + w];
    z = r[(int64_t)(3 * e + 2) * W + w];
  };

  float acc[CH][UPT];
  float x = 0.f, y = 0.f, z = 0.f;
  if (chunk < n_items) {
    if (worker) {
      load_xyz(chunk, x, y, z);
      produce_features(s, Fs_row, quarter, x, y, z);
    }
    __syncthreads();
    if (worker) layer1_chunk(s, W1s, b1s, Fs_row, x, y, z, UPT * quarter, acc);
  }
  for (int64_t item = chunk; item < n_items; item += n_chunks) {
    const int i = (int)(item % ns);
    const int64_t w0 = (item / ns) * ROWS;
    const int64_t next = item + n_chunks;
    float xn = 0.f, yn = 0.f, zn = 0.f;
    if (worker && next < n_items) load_xyz(next, xn, yn, zn);
#pragma unroll 1
    for (int kc = 0; kc < H / KC; ++kc) {
      if (worker) {
        // the A columns are still being read by the previous chunk's MMAs (for kc = 0 the epilogue already waited)
        if (kc > 0) { mbar_wait(barA, phaseA); phaseA ^= 1; }
        const uint32_t t_hi = tmem_base + lane_base + C_AH + (uint32_t)(UPT * quarter);
        const uint32_t t_lo = tmem_base + lane_base + C_AL + (uint32_t)(UPT * quarter);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) split_tf32(acc[c][t], hi[t], lo[t]);
          tmem_st4(t_hi + 16u * c, hi[0], hi[1], hi[2], hi[3]);
          tmem_st4(t_lo + 16u * c, lo[0], lo[1], lo[2], lo[3]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (warp == TCW / 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t leader = elect_one();
        const uint64_t dBh = umma_desc(bhi_base, B_K4, 128), dBl = umma_desc(blo_base, B_K4, 128);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
#pragma unroll
          for (int ks = 0; ks < KC / 8; ++ks) {
            const uint64_t kb = (uint64_t)((((uint32_t)kc * (KC / 4) + ks * 2) * B_K4) >> 4);
            const uint32_t ah = tmem_base + C_AH + 16u * c + 8u * ks, al = tmem_base + C_AL + 16u * c + 8u * ks;
            const uint32_t d = tmem_base + (uint32_t)(c * H);
            const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
            umma_ts(d, ah, dBh + kb, first, leader);
            umma_ts(d, al, dBh + kb, 1u, leader);
            umma_ts(d, ah, dBl + kb, 1u, leader);
          }
        }
        umma_commit(barA, leader);
        __syncwarp();
      }
      // ---- layer 1 of the next chunk (or of the next tile's first chunk) overlaps these MMAs ------------------------
      if (kc + 1 < H / KC) {
        if (worker) layer1_chunk(s, W1s, b1s, Fs_row, x, y, z, KC * (kc + 1) + UPT * quarter, acc);
      } else if (next < n_items) {
        if (worker) produce_features(s, Fs_row, quarter, xn, yn, zn);   // every read of this tile's features is done
        __syncthreads();
        if (worker) layer1_chunk(s, W1s, b1s, Fs_row, xn, yn, zn, UPT * quarter, acc);
      }
    }
    x = xn; y = yn; z = zn;
    // ---- epilogue: accumulators -> activation rules -> output layer -----------------------------------------------
    float out[CH] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (worker) {
      mbar_wait(barA, phaseA);
      phaseA ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int sub = 0; sub < EPC / 8; ++sub) {
        const int c0 = EPC * quarter + 8 * sub;
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
    __syncthreads();                                          // accumulators free; partial outputs visible
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
    // outS is rewritten only after the next tile's barriers
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

}  // namespace

// returns NQMC_ERR_UNSUPPORTED when this variant does not apply (the caller falls back to nqmc_orbital_tc.cu / FFMA)
int nq_launch_orb_lap_tc(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H != H) return NQMC_ERR_UNSUPPORTED;
  const size_t smem = smem_bytes(s.F, s.A);
  if (smem > 227 * 1024) return NQMC_ERR_UNSUPPORTED;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(orb_lap_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  {
    ProfScope ps(ctx, NQ_PROF_EVAL5, st);
    orb_lap_tc_kernel<<<n_chunks * s.n_mlp, TCT, smem, st>>>(s, ctx->theta, r, W, phi, n_chunks);
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
