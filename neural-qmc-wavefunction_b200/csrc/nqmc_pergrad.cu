// nqmc_pergrad.cu -- per-walker log-derivatives of the antisymmetric ansatz, parameter-major:
//     Ot[p][w] = d log|psi(r_w)| / d theta_p
// Replaces: vmap(jax.grad(lambda p: wavefunction(p, r)))(samples) -- src/nqmc/optimizer/vmc.py:120-123 -- in the one
// case where the matrix itself is wanted: the O of Stochastic Reconfiguration ([SPEC] .github/phase6-issue-body.md:152-165).
// The walker step never forms O (its weight gradients are K = walkers contractions on the tensor cores,
// nqmc_orbital_bwd*.cu); here every walker's outer products are written out, so the kernel is bound by the P x W x 4
// bytes it stores (H4 / 64-wide, 65,536 walkers: 5.2 GB).
//   d log|det Phi_s| / d Phi_s[i,k] = inv_s[k,i]    seeds network k of spin s at electron i;
//   three-layer reverse pass per (walker, electron), FP32 FFMA; the sum over the spin's electrons is taken when the
//   outer products are formed, so every element of Ot is written exactly once.
// Backflow: the orbitals are evaluated at x(r) (nqmc_backflow.cu) and the eta network adds
//   d log|det| / d eta_p = (v_i - v_j) . (r_i - r_j)  per pair, a 2-16-16-1 reverse pass summed over the pairs.
#include "nqmc_internal.cuh"

namespace {

constexpr int TW = 32;                 // walkers per CTA (= lanes: every shared-memory access is conflict-free)
constexpr int NWARP = 8;
constexpr int BH = NQMC_BF_H;

// shared memory per electron of the spin block: features, h1, delta1, delta2, seed * h2, seed
__host__ __device__ constexpr int per_entry_floats(int F, int H) { return (F + 4 * H + 1) * TW; }

template <int H>
__global__ void __launch_bounds__(TW* NWARP) orb_o_kernel(const DevSys s, const float* __restrict__ theta,
                                                         const float* __restrict__ pos, const float* __restrict__ inv,
                                                         const int64_t W, float* __restrict__ Ot, const int64_t ld) {
  extern __shared__ float sm[];
  const int F = s.F, A = s.A;
  const int m = blockIdx.y, spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up, ns = spin == 0 ? s.n_up : s.n_dn, elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int64_t w = (int64_t)blockIdx.x * TW + lane, wc = w < W ? w : W - 1;
  const int PE = per_entry_floats(F, H);
  for (int i = 0; i < ns; ++i) {
    float* sf = sm + i * PE;             // [F][TW]
    float* sh1 = sf + F * TW;            // [H][TW]
    float* sd1 = sh1 + H * TW;
    float* sd2 = sd1 + H * TW;
    float* sh2 = sd2 + H * TW;           // seed * h2
    float* ssd = sh2 + H * TW;           // [TW]
    if (q == 0) {
      const int e = elec0 + i;
      const float x = pos[(int64_t)(3 * e) * W + wc], y = pos[(int64_t)(3 * e + 1) * W + wc], z = pos[(int64_t)(3 * e + 2) * W + wc];
      sf[lane] = x; sf[TW + lane] = y; sf[2 * TW + lane] = z;
      for (int a = 0; a < A; ++a) {
        const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
        const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        sf[(3 + a) * TW + lane] = d;
        sf[(3 + A + a) * TW + lane] = __expf(-d);
      }
      float sd = inv[(int64_t)ent_index(s, spin, i, korb) * W + wc];
      if (!isfinite(sd)) sd = 0.f;
      ssd[lane] = sd;
    }
    __syncthreads();
    const float seed = ssd[lane];
    for (int u = q; u < H; u += NWARP) {
      float a = theta[off.b1 + u];
      for (int f = 0; f < F; ++f) a = fmaf(theta[off.W1 + f * H + u], sf[f * TW + lane], a);
      sh1[u * TW + lane] = nq_tanh(a);
    }
    __syncthreads();
    for (int k = q; k < H; k += NWARP) {
      float a = theta[off.b2 + k];
#pragma unroll 8
      for (int j = 0; j < H; ++j) a = fmaf(theta[off.W2 + j * H + k], sh1[j * TW + lane], a);
      const float t = nq_tanh(a);
      sh2[k * TW + lane] = seed * t;
      sd2[k * TW + lane] = seed * theta[off.W3 + k] * fmaf(-t, t, 1.f);
    }
    __syncthreads();
    for (int j = q; j < H; j += NWARP) {
      float a = 0.f;
#pragma unroll 8
      for (int k = 0; k < H; ++k) a = fmaf(theta[off.W2 + j * H + k], sd2[k * TW + lane], a);
      const float t = sh1[j * TW + lane];
      sd1[j * TW + lane] = a * fmaf(-t, t, 1.f);
    }
    __syncthreads();
  }
  if (w >= W) return;
  float* out = Ot + w;
  auto S1 = [&](const int o_rel) {       // sum over the spin's electrons of one staged value
    float a = 0.f;
    for (int i = 0; i < ns; ++i) a += sm[i * PE + o_rel + lane];
    return a;
  };
  const int o_h1 = F * TW, o_d1 = o_h1 + H * TW, o_d2 = o_d1 + H * TW, o_h2 = o_d2 + H * TW, o_sd = o_h2 + H * TW;
  for (int j = q; j < H; j += NWARP) {
    out[(int64_t)(off.b1 + j) * ld] = S1(o_d1 + j * TW);
    out[(int64_t)(off.b2 + j) * ld] = S1(o_d2 + j * TW);
    out[(int64_t)(off.W3 + j) * ld] = S1(o_h2 + j * TW);
  }
  if (q == 0) out[(int64_t)off.b3 * ld] = S1(o_sd);
  for (int idx = q; idx < F * H; idx += NWARP) {
    const int f = idx / H, j = idx % H;
    float a = 0.f;
    for (int i = 0; i < ns; ++i) a = fmaf(sm[i * PE + f * TW + lane], sm[i * PE + o_d1 + j * TW + lane], a);
    out[(int64_t)(off.W1 + idx) * ld] = a;
  }
  for (int j = q; j < H; j += NWARP) {
    float hj[NQMC_MAX_NS];
    for (int i = 0; i < ns; ++i) hj[i] = sm[i * PE + o_h1 + j * TW + lane];
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      float a = 0.f;
      for (int i = 0; i < ns; ++i) a = fmaf(hj[i], sm[i * PE + o_d2 + k * TW + lane], a);
      out[(int64_t)(off.W2 + j * H + k) * ld] = a;
    }
  }
}

// Jastrow scalars, e-e cusp scalar and the e-n cusp networks (formulas of walker_grads_kernel, nqmc_walker.cu, without the
// sum over walkers).  grid (walker blocks, 1 + A): y == 0 the scalars, y == 1 + a the network of nucleus a.
constexpr int WT = 128;
__global__ void __launch_bounds__(WT) walker_o_kernel(const DevSys s, const float* __restrict__ theta,
                                                      const float* __restrict__ r, const int64_t W,
                                                      float* __restrict__ Ot, const int64_t ld) {
  __shared__ float sd[NQMC_MAX_ELEC * WT], se[NQMC_MAX_ELEC * WT];
  const int64_t w = (int64_t)blockIdx.x * WT + threadIdx.x;
  if (w >= W) return;
  float* out = Ot + w;
  if (blockIdx.y == 0) {
    const float a_ee = theta[s.jas_a_ee], b_ee = theta[s.jas_b_ee];
    const float a_en = theta[s.jas_a_en], b_en = theta[s.jas_b_en];
    const float cbs = s.use_cusp ? theta[s.ee_cusp_b] : 0.f;
    float g_aee = 0.f, g_bee = 0.f, g_aen = 0.f, g_ben = 0.f, g_cb = 0.f;
    for (int i = 0; i < s.N; ++i) {
      const float xi = r[(int64_t)(3 * i) * W + w], yi = r[(int64_t)(3 * i + 1) * W + w], zi = r[(int64_t)(3 * i + 2) * W + w];
      for (int j = i + 1; j < s.N; ++j) {
        const float dx = xi - r[(int64_t)(3 * j) * W + w], dy = yi - r[(int64_t)(3 * j + 1) * W + w], dz = zi - r[(int64_t)(3 * j + 2) * W + w];
        const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        const float q = 1.0f / fmaf(b_ee, d, 1.0f);
        g_aee = fmaf(d, q, g_aee);
        g_bee = fmaf(-a_ee * d * d, q * q, g_bee);
        if (s.use_cusp) {
          const bool same = (i < s.n_up) == (j < s.n_up);
          const float q2 = 1.0f / fmaf(fabsf(cbs), d, 1.0f);
          g_cb = fmaf(-(same ? s.cusp_same : s.cusp_diff) * d * d, q2 * q2, g_cb);
        }
      }
      for (int a = 0; a < s.A; ++a) {
        const float dx = xi - s.R[3 * a], dy = yi - s.R[3 * a + 1], dz = zi - s.R[3 * a + 2];
        const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        const float q = 1.0f / fmaf(b_en, d, 1.0f);
        g_aen = fmaf(d, q, g_aen);
        g_ben = fmaf(-a_en * d * d, q * q, g_ben);
      }
    }
    out[(int64_t)s.jas_a_ee * ld] = g_aee;
    out[(int64_t)s.jas_a_en * ld] = g_aen;
    out[(int64_t)s.jas_b_ee * ld] = g_bee;
    out[(int64_t)s.jas_b_en * ld] = g_ben;
    if (s.use_cusp) out[(int64_t)s.ee_cusp_b * ld] = g_cb * (cbs < 0.f ? -1.f : (cbs > 0.f ? 1.f : 0.f));
    return;
  }
  const int an = blockIdx.y - 1;
  const float* W0 = theta + s.en_cusp_W0[an];
  const float* b0 = theta + s.en_cusp_b0[an];
  const float* W1 = theta + s.en_cusp_W1[an];
  float sumd = 0.f;
  for (int i = 0; i < s.N; ++i) {
    const float dx = r[(int64_t)(3 * i) * W + w] - s.R[3 * an], dy = r[(int64_t)(3 * i + 1) * W + w] - s.R[3 * an + 1],
                dz = r[(int64_t)(3 * i + 2) * W + w] - s.R[3 * an + 2];
    const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    sd[i * WT + threadIdx.x] = d;
    se[i * WT + threadIdx.x] = __expf(-d);
    sumd += d;
  }
  out[(int64_t)s.en_cusp_b1[an] * ld] = sumd;
  for (int h = 0; h < NQMC_CUSP_H; ++h) {
    const float w0 = W0[h], w1 = W0[NQMC_CUSP_H + h], w2 = W0[2 * NQMC_CUSP_H + h], bb = b0[h], wo = W1[h];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
    for (int i = 0; i < s.N; ++i) {
      const float d = sd[i * WT + threadIdx.x], e = se[i * WT + threadIdx.x];
      const float t = nq_tanh(fmaf(w0, d, fmaf(w1, d * d, fmaf(w2, e, bb))));
      const float dz_ = d * wo * fmaf(-t, t, 1.0f);
      a0 += dz_;
      a1 = fmaf(dz_, d, a1);
      a2 = fmaf(dz_ * d, d, a2);
      a3 = fmaf(dz_, e, a3);
      a4 = fmaf(d, t, a4);
    }
    out[(int64_t)(s.en_cusp_b0[an] + h) * ld] = a0;
    out[(int64_t)(s.en_cusp_W0[an] + h) * ld] = a1;
    out[(int64_t)(s.en_cusp_W0[an] + NQMC_CUSP_H + h) * ld] = a2;
    out[(int64_t)(s.en_cusp_W0[an] + 2 * NQMC_CUSP_H + h) * ld] = a3;
    out[(int64_t)(s.en_cusp_W1[an] + h) * ld] = a4;
  }
}

// eta network of the backflow transform: thread per walker, pairs in sequence (337 parameters: the accumulators are the
// output rows themselves, held in shared memory as [parameter][walker] so that the final stores are coalesced).
constexpr int BF_P = 2 * BH + BH + BH * BH + BH + BH + 1;   // 337
constexpr int BT = 32;
__global__ void __launch_bounds__(BT) bf_o_kernel(const DevSys s, const float* __restrict__ theta,
                                                  const float* __restrict__ r, const float* __restrict__ vvec,
                                                  const int64_t W, float* __restrict__ Ot, const int64_t ld) {
  __shared__ float acc[BF_P][BT];
  const MlpOff o = s.bf;
  const int t = threadIdx.x, N = s.N;
  const int64_t w = (int64_t)blockIdx.x * BT + t, wc = w < W ? w : W - 1;
  for (int p = 0; p < BF_P; ++p) acc[p][t] = 0.f;
  // rows of acc in theta-relative order of the leaves: b1[16] W1[2][16] b2[16] W2[16][16] b3 W3[16]
  constexpr int o_b1 = 0, o_W1 = BH, o_b2 = 3 * BH, o_W2 = 4 * BH, o_b3 = 4 * BH + BH * BH, o_W3 = o_b3 + 1;
  for (int i = 0; i < N; ++i)
    for (int j = i + 1; j < N; ++j) {
      const float dx = r[(int64_t)(3 * i) * W + wc] - r[(int64_t)(3 * j) * W + wc];
      const float dy = r[(int64_t)(3 * i + 1) * W + wc] - r[(int64_t)(3 * j + 1) * W + wc];
      const float dz = r[(int64_t)(3 * i + 2) * W + wc] - r[(int64_t)(3 * j + 2) * W + wc];
      const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      const float sv = (vvec[(int64_t)(3 * i) * W + wc] - vvec[(int64_t)(3 * j) * W + wc]) * dx +
                       (vvec[(int64_t)(3 * i + 1) * W + wc] - vvec[(int64_t)(3 * j + 1) * W + wc]) * dy +
                       (vvec[(int64_t)(3 * i + 2) * W + wc] - vvec[(int64_t)(3 * j + 2) * W + wc]) * dz;
      float seed = sv * __expf(-0.5f * d);
      if (!isfinite(seed)) seed = 0.f;
      const float e = __expf(-d);
      float h1[BH], d2[BH];
#pragma unroll
      for (int q = 0; q < BH; ++q) h1[q] = nq_tanh(fmaf(theta[o.W1 + q], d, fmaf(theta[o.W1 + BH + q], e, theta[o.b1 + q])));
#pragma unroll 4
      for (int k = 0; k < BH; ++k) {
        float u = theta[o.b2 + k];
#pragma unroll
        for (int q = 0; q < BH; ++q) u = fmaf(theta[o.W2 + q * BH + k], h1[q], u);
        const float h2 = nq_tanh(u);
        d2[k] = seed * theta[o.W3 + k] * fmaf(-h2, h2, 1.f);
        acc[o_W3 + k][t] = fmaf(seed, h2, acc[o_W3 + k][t]);
        acc[o_b2 + k][t] += d2[k];
      }
      acc[o_b3][t] += seed;
#pragma unroll
      for (int q = 0; q < BH; ++q) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < BH; ++k) {
          a = fmaf(theta[o.W2 + q * BH + k], d2[k], a);
          acc[o_W2 + q * BH + k][t] = fmaf(h1[q], d2[k], acc[o_W2 + q * BH + k][t]);
        }
        const float d1 = a * fmaf(-h1[q], h1[q], 1.f);
        acc[o_b1 + q][t] += d1;
        acc[o_W1 + q][t] = fmaf(d, d1, acc[o_W1 + q][t]);
        acc[o_W1 + BH + q][t] = fmaf(e, d1, acc[o_W1 + BH + q][t]);
      }
    }
  if (w >= W) return;
  for (int p = 0; p < BF_P; ++p) {
    int dst;
    if (p < BH) dst = o.b1 + p;
    else if (p < 3 * BH) dst = o.W1 + (p - BH);
    else if (p < 4 * BH) dst = o.b2 + (p - 3 * BH);
    else if (p < 4 * BH + BH * BH) dst = o.W2 + (p - 4 * BH);
    else if (p == 4 * BH + BH * BH) dst = o.b3;
    else dst = o.W3 + (p - (4 * BH + BH * BH + 1));
    Ot[(int64_t)dst * ld + w] = acc[p][t];
  }
}

template <int H>
int launch_orb_o(nqmc_ctx* ctx, const float* pos, int64_t W, float* Ot, int64_t ld, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const int ns_max = s.n_up > s.n_dn ? s.n_up : s.n_dn;
  const size_t smem = (size_t)ns_max * per_entry_floats(s.F, H) * sizeof(float);
  if (smem > 227 * 1024) {
    ctx->err = "nqmc_log_psi_grads: electrons per spin x hidden width exceed the shared-memory staging of the per-walker kernel";
    return NQMC_ERR_UNSUPPORTED;
  }
  NQMC_CUDA(cudaFuncSetAttribute(orb_o_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  orb_o_kernel<H><<<dim3((unsigned)((W + TW - 1) / TW), (unsigned)s.n_mlp), TW * NWARP, smem, st>>>(s, ctx->theta, pos, ctx->inv, W, Ot, ld);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

}  // namespace

// needs ctx->inv (and, with backflow, ctx->x_cur and ctx->vvec) for these positions: nqmc_log_psi_grads runs the value /
// energy pass first
int nq_antisym_grads_per_walker(nqmc_ctx* ctx, const float* r, int64_t W, float* Ot, int64_t ld, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const float* pos = s.use_bf ? ctx->x_cur : r;
  int rc;
  if (s.H == 32) rc = launch_orb_o<32>(ctx, pos, W, Ot, ld, st);
  else if (s.H == 64) rc = launch_orb_o<64>(ctx, pos, W, Ot, ld, st);
  else if (s.H == 128) rc = launch_orb_o<128>(ctx, pos, W, Ot, ld, st);
  else {
    ctx->err = "nqmc_log_psi_grads: hidden width must be 32, 64 or 128";
    return NQMC_ERR_UNSUPPORTED;
  }
  if (rc) return rc;
  walker_o_kernel<<<dim3((unsigned)((W + WT - 1) / WT), (unsigned)(1 + (s.use_cusp ? s.A : 0))), WT, 0, st>>>(s, ctx->theta, r, W, Ot, ld);
  NQMC_LAUNCH_CHECK();
  if (s.use_bf) {
    bf_o_kernel<<<(unsigned)((W + BT - 1) / BT), BT, 0, st>>>(s, ctx->theta, r, ctx->vvec, W, Ot, ld);
    NQMC_LAUNCH_CHECK();
  }
  return NQMC_OK;
}
