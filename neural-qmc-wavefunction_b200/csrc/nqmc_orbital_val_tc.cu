// nqmc_orbital_val_tc.cu -- tensor-core (tcgen05 / TMEM, 3xTF32) value-only orbital forward pass for hidden
// width 64: phi_m(x_i) for the Metropolis proposal (src/nqmc/sampler/metropolis.py:124).  Same math and
// interface as orb_eval_kernel<64,1,...> (nqmc_orbital.cu), the FP32 FFMA parity anchor.
//
// Tile = 128 rows (128 consecutive walkers, one electron of the network's spin); thread (row, quarter) owns
// 16 hidden units / output columns of its row.
//     h1 = tanh(f W1 + b1)          FP32 FFMA, features in registers, weights broadcast from shared memory
//     U2 = h1 W2                    UMMA M=128, N=64, K=64; A = h1 split into tf32 hi/lo, written by its owner
//                                   straight into TENSOR MEMORY (tcgen05.st); B = W2 hi/lo in shared memory
//     phi = W3 . tanh(U2 + b2) + b3 epilogue from TMEM, the 4 quarters of a row combined through shared memory
// Two tiles are in flight: accumulator and A-operand columns are double-buffered in TMEM, so the layer-1 work of
// tile t+1 overlaps the MMAs of tile t (one __syncthreads per tile).
#include "nqmc_internal.cuh"

namespace {

constexpr int H = 64;
constexpr int NQ = 4;
constexpr int TCW = 128 * NQ;
constexpr int TCT = TCW;              // 16 warps: 128 registers per thread
constexpr int ISSUER = TCW / 32 - 1;  // the last worker warp also issues the MMAs
constexpr int ROWS = 128;
constexpr int UPT = H / NQ;
constexpr uint32_t TMEM_COLS = 512;
// per buffer b (0/1): U2 at 64 b, A_hi at 128 + 128 b, A_lo at 192 + 128 b
constexpr int WB = (H / 4) * H * 16;

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
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  lo = a - hi;
}

// NA = number of nuclei when > 0 (feature loop fully unrolled), 0 = run-time count
template <int NA>
__global__ void __launch_bounds__(TCT, 1)
    orb_eval1_tc_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r, const int64_t W,
                        float* __restrict__ phi, const int n_chunks) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* Whi = smem;                                   // [k/4][j][k%4]
  uint8_t* Wlo = smem + WB;
  float* W1s = reinterpret_cast<float*>(smem + 2 * WB);  // [F][H]
  const int A = NA > 0 ? NA : s.A, F = 3 + 2 * A;
  float* b1s = W1s + F * H;
  float* b2s = b1s + H;
  float* W3s = b2s + H;
  float* outS = W3s + H;                                 // [2 buffers][NQ-1][ROWS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(outS + 2 * (NQ - 1) * ROWS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (ROWS - 1), quarter = (tid >> 7) & (NQ - 1);
  const bool worker = tid < TCW;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];

  for (int idx = tid; idx < F * H; idx += TCT) W1s[idx] = theta[off.W1 + idx];
  for (int q = tid; q < H; q += TCT) {
    b1s[q] = theta[off.b1 + q];
    b2s[q] = theta[off.b2 + q];
    W3s[q] = theta[off.W3 + q];
  }
  {
    // all loads of a thread in flight at once: with a cold L2 each dependent round trip costs ~1 us
    constexpr int NL = H * H / TCT;
    static_assert(NL * TCT == H * H, "W2 staging assumes the block size divides H*H");
    float wv[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) wv[i] = theta[off.W2 + tid + i * TCT];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int idx = tid + i * TCT;
      const int k = idx / H, j = idx % H;
      float hi, lo;
      split_tf32(wv[i], hi, lo);
      const int o = (k >> 2) * (H * 4) + j * 4 + (k & 3);
      reinterpret_cast<float*>(Whi)[o] = hi;
      reinterpret_cast<float*>(Wlo)[o] = lo;
    }
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
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t whi = smem_u32(Whi), wlo = smem_u32(Wlo);
  const uint32_t bar[2] = {smem_u32(&bars[0]), smem_u32(&bars[1])};
  uint32_t phase[2] = {0u, 0u};
  const int c0 = UPT * quarter;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;

  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;

  // layer 1 of tile `it` for this thread's (row, 16 units) -> tf32 hi/lo into the A columns of buffer b
  float nx = 0.f, ny = 0.f, nz = 0.f;                    // positions of the tile produced next, fetched one tile ahead
  auto fetch = [&](int64_t it) {
    int64_t w = (it / ns) * ROWS + row;
    if (w >= W) w = W - 1;
    const int e = elec0 + (int)(it % ns);
    nx = r[(int64_t)(3 * e + 0) * W + w];
    ny = r[(int64_t)(3 * e + 1) * W + w];
    nz = r[(int64_t)(3 * e + 2) * W + w];
  };
  auto produce = [&](int64_t it, int b) {
    const float x = nx, y = ny, z = nz;
    if (it + n_chunks < n_items) fetch(it + n_chunks);
    float h1[UPT];
#pragma unroll
    for (int j = 0; j < UPT; ++j) h1[j] = b1s[c0 + j];
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
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
      const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      const float ex = __expf(-dd);
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
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < UPT; ++j) {
      float h, l;
      split_tf32(nq_tanh(h1[j]), h, l);
      hi[j] = __float_as_uint(h);
      lo[j] = __float_as_uint(l);
    }
    tmem_st16(tmem_base + lane_base + 128u + 128u * b + (uint32_t)c0, hi);
    tmem_st16(tmem_base + lane_base + 192u + 128u * b + (uint32_t)c0, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  };
  auto issue = [&](int b) {   // whole issuer warp, elected lane executes
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t leader = elect_one();
    const uint64_t bh = umma_desc(whi, H * 16, 128), bl = umma_desc(wlo, H * 16, 128);
    const uint32_t d = tmem_base + 64u * b;
#pragma unroll
    for (int ks = 0; ks < H / 8; ++ks) {
      const uint64_t bo = (uint64_t)((ks * 2 * H * 16) >> 4);
      const uint32_t ah = tmem_base + 128u + 128u * b + 8 * ks, al = tmem_base + 192u + 128u * b + 8 * ks;
      umma_ts(d, ah, bh + bo, ks == 0 ? 0u : 1u, leader);
      umma_ts(d, al, bh + bo, 1u, leader);
      umma_ts(d, ah, bl + bo, 1u, leader);
    }
    umma_commit(bar[b], leader);
    __syncwarp();
  };

  nq_pdl_wait();                              // everything above read theta only; the positions come from the predecessor
  // ---- software pipeline, depth 2 -------------------------------------------------------------------------------
  int b = 0;
  if (chunk < n_items) {
    if (worker) {
      fetch(chunk);
      produce(chunk, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == ISSUER) issue(0);
  }
  for (int64_t item = chunk; item < n_items; item += n_chunks, b ^= 1) {
    const int64_t next = item + n_chunks;
    if (worker && next < n_items) produce(next, b ^ 1);       // overlaps the MMAs of `item`
    float out = 0.f;
    if (worker) {
      mbar_wait(bar[b], phase[b]);
      phase[b] ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[16];
      tmem_ld16(tmem_base + lane_base + 64u * b + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < UPT; ++j) out = fmaf(W3s[c0 + j], nq_tanh(v[j] + b2s[c0 + j]), out);
      if (quarter != 0) outS[(b * (NQ - 1) + quarter - 1) * ROWS + row] = out;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                        // A columns of `next` staged; accumulator b read; partials visible
    if (warp == ISSUER && next < n_items) issue(b ^ 1);
    if (worker && quarter == 0) {
      const int64_t wr = (item / ns) * ROWS + row;
      if (wr < W) {
        float o = out + b3;
#pragma unroll
        for (int q = 0; q < NQ - 1; ++q) o += outS[(b * (NQ - 1) + q) * ROWS + row];
        phi[(int64_t)ent_index(s, spin, (int)(item % ns), korb) * W + wr] = o;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

template <int NA>
int launch_t(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, int n_chunks, size_t smem, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(orb_eval1_tc_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  {
    ProfScope ps(ctx, NQ_PROF_EVAL1, st);
    if (ctx->prof_on) orb_eval1_tc_kernel<NA><<<n_chunks * s.n_mlp, TCT, smem, st>>>(s, ctx->theta, r, W, phi, n_chunks);
    else NQMC_CUDA(nq_launch_pdl_guarded(ctx, orb_eval1_tc_kernel<NA>, dim3(n_chunks * s.n_mlp), dim3(TCT), smem, st, s, ctx->theta, r, W, phi, n_chunks));
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

}  // namespace

int nq_launch_orb_eval1_tc(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H != H) return NQMC_ERR_UNSUPPORTED;
  const size_t smem = 2 * WB + ((size_t)s.F * H + 3 * H + 2 * (NQ - 1) * ROWS) * 4 + 64;
  if (smem > 96 * 1024) return NQMC_ERR_UNSUPPORTED;
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  switch (s.A) {
    case 1: return launch_t<1>(ctx, r, W, phi, n_chunks, smem, st);
    case 2: return launch_t<2>(ctx, r, W, phi, n_chunks, smem, st);
    case 3: return launch_t<3>(ctx, r, W, phi, n_chunks, smem, st);
    case 4: return launch_t<4>(ctx, r, W, phi, n_chunks, smem, st);
    case 6: return launch_t<6>(ctx, r, W, phi, n_chunks, smem, st);
    default: return launch_t<0>(ctx, r, W, phi, n_chunks, smem, st);
  }
}
