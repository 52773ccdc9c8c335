// nqmc_simple.cu -- ansatz kind 0: the reference's only *implemented* wavefunction,
//   log|psi| = MLP(flatten r) + sum_i log sum_A exp(-alpha |r_i - R_A|)
// (src/nqmc/wavefunction/simple.py:41-70 SimpleMLP, :124-164 envelope + __call__), used by BASELINE
// config 1 (H2, scripts/train_h2.py).  It is the plumbing case (32 chains), so the mapping is
// simple: one WARP per walker, lanes across hidden units, weights staged once per CTA in shared
// memory; the forward-Laplacian carries the D = 3N input tangents explicitly
// (hamiltonian/molecular.py:52-69 computes the same trace + |grad|^2 by nested autodiff).
#include "nqmc_internal.cuh"

namespace {

constexpr int SW = 4;            // warps (walkers in flight) per CTA
constexpr int MAXH = 128;
constexpr int MAXD = 48;
constexpr int HPL = MAXH / 32;   // hidden units per lane

struct SimpleSmem {
  float* W1;   // [D][H0]
  float* W2;   // [H0][H1]
  float* b1;   // [H0]
  float* b2;   // [H1]
  float* W3;   // [H1]
  float* scr;  // per warp: h1[H0], lh1[H0], G1[D][H0], x[D], d2[H1]
};

__device__ __forceinline__ SimpleSmem carve(float* base, int D, int H0, int H1, int warp) {
  SimpleSmem s;
  s.W1 = base;
  s.W2 = s.W1 + D * H0;
  s.b1 = s.W2 + H0 * H1;
  s.b2 = s.b1 + H0;
  s.W3 = s.b2 + H1;
  const int per_warp = 2 * H0 + D * H0 + MAXD + H1;
  s.scr = s.W3 + H1 + warp * per_warp;
  return s;
}

__device__ __forceinline__ void stage_weights(const DevSys& s, const float* __restrict__ theta, const SimpleSmem& m,
                                              int D, int H0, int H1) {
  const MlpOff o = s.simple;
  for (int i = threadIdx.x; i < D * H0; i += blockDim.x) m.W1[i] = theta[o.W1 + i];
  for (int i = threadIdx.x; i < H0 * H1; i += blockDim.x) m.W2[i] = theta[o.W2 + i];
  for (int i = threadIdx.x; i < H0; i += blockDim.x) m.b1[i] = theta[o.b1 + i];
  for (int i = threadIdx.x; i < H1; i += blockDim.x) { m.b2[i] = theta[o.b2 + i]; m.W3[i] = theta[o.W3 + i]; }
}

// envelope: value, gradient (per lane: one electron-axis each), Laplacian -- done by lane 0..D-1
__device__ __forceinline__ void envelope(const DevSys& s, const float* x, int i, float& val, float& gx, float& gy,
                                         float& gz, float& lap) {
  float S = 0.f, sx = 0.f, sy = 0.f, sz = 0.f, sl = 0.f;
  const float al = s.env_alpha;
  for (int a = 0; a < s.A; ++a) {
    float dx = x[3 * i] - s.R[3 * a], dy = x[3 * i + 1] - s.R[3 * a + 1], dz = x[3 * i + 2] - s.R[3 * a + 2];
    float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    float e = __expf(-al * d);
    float id = 1.0f / d;
    S += e;
    float c = -al * e * id;
    sx = fmaf(c, dx, sx); sy = fmaf(c, dy, sy); sz = fmaf(c, dz, sz);
    sl += e * (al * al - 2.0f * al * id);
  }
  const float iS = 1.0f / S;
  val = logf(S);
  gx = sx * iS; gy = sy * iS; gz = sz * iS;
  lap = sl * iS - (gx * gx + gy * gy + gz * gz);
}

// mode 0: log|psi| ; mode 1: local energy
__global__ void __launch_bounds__(SW * 32) simple_eval_kernel(const DevSys s, const float* __restrict__ theta,
                                                              const float* __restrict__ r, const int64_t W,
                                                              float* __restrict__ logabs, float* __restrict__ e_l,
                                                              float* __restrict__ parts, const int mode) {
  extern __shared__ __align__(16) float smem[];
  const int D = 3 * s.N, H0 = s.H, H1 = s.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SimpleSmem m = carve(smem, D, H0, H1, warp);
  stage_weights(s, theta, m, D, H0, H1);
  __syncthreads();
  float* h1 = m.scr;
  float* lh1 = h1 + H0;
  float* G1 = lh1 + H0;
  float* x = G1 + D * H0;
  const float b3 = theta[s.simple.b3];
  for (int64_t w = (int64_t)blockIdx.x * SW + warp; w < W; w += (int64_t)gridDim.x * SW) {
    for (int d = lane; d < D; d += 32) x[d] = r[(int64_t)d * W + w];
    __syncwarp();
    // layer 1
    for (int j = lane; j < H0; j += 32) {
      float u = m.b1[j], w2 = 0.f;
      for (int d = 0; d < D; ++d) { float wv = m.W1[d * H0 + j]; u = fmaf(wv, x[d], u); w2 = fmaf(wv, wv, w2); }
      float h = nq_tanh(u), tp = fmaf(-h, h, 1.f);
      h1[j] = h;
      if (mode == 1) {
        lh1[j] = -2.f * h * tp * w2;
        for (int d = 0; d < D; ++d) G1[d * H0 + j] = tp * m.W1[d * H0 + j];
      }
    }
    __syncwarp();
    // layer 2 + output: lane owns hidden units k = lane + 32 t (static slots)
    float val = 0.f;
    float hk[HPL], tpk[HPL], w3k[HPL], luk[HPL], g2k[HPL];
#pragma unroll
    for (int t = 0; t < HPL; ++t) {
      const int k = lane + 32 * t;
      hk[t] = tpk[t] = w3k[t] = luk[t] = g2k[t] = 0.f;
      if (k < H1) {
        float u = m.b2[k], lu = 0.f;
        for (int j = 0; j < H0; ++j) {
          float wv = m.W2[j * H1 + k];
          u = fmaf(wv, h1[j], u);
          if (mode == 1) lu = fmaf(wv, lh1[j], lu);
        }
        float h = nq_tanh(u);
        hk[t] = h; tpk[t] = fmaf(-h, h, 1.f); w3k[t] = m.W3[k]; luk[t] = lu;
        val = fmaf(w3k[t], h, val);
      }
    }
    val = warp_sum(val) + b3;
    // envelope (lanes < N)
    float env = 0.f, egx = 0.f, egy = 0.f, egz = 0.f, elap = 0.f;
    if (lane < s.N) envelope(s, x, lane, env, egx, egy, egz, elap);
    const float envsum = warp_sum(env);
    if (mode == 0) {
      if (lane == 0) logabs[w] = val + envsum;
    } else {
      float g2 = 0.f;
#pragma unroll 1
      for (int d = 0; d < D; ++d) {
        float gpart = 0.f;
#pragma unroll
        for (int t = 0; t < HPL; ++t) {
          const int k = lane + 32 * t;
          if (k < H1) {
            float gu = 0.f;
            for (int j = 0; j < H0; ++j) gu = fmaf(m.W2[j * H1 + k], G1[d * H0 + j], gu);
            g2k[t] = fmaf(gu, gu, g2k[t]);
            gpart = fmaf(w3k[t] * tpk[t], gu, gpart);
          }
        }
        const float gm = warp_sum(gpart);
        const int i = d / 3, ax = d % 3;
        const float ge = __shfl_sync(0xffffffffu, ax == 0 ? egx : (ax == 1 ? egy : egz), i);
        g2 = fmaf(gm + ge, gm + ge, g2);
      }
      float lap = 0.f;
#pragma unroll
      for (int t = 0; t < HPL; ++t) lap = fmaf(w3k[t], fmaf(tpk[t], luk[t], -2.f * hk[t] * tpk[t] * g2k[t]), lap);
      lap = warp_sum(lap) + warp_sum(elap);
      // potentials (molecular.py:87-135), lanes split the pairs
      float ven = 0.f, vee = 0.f;
      for (int p = lane; p < s.N * s.A; p += 32) {
        int i = p / s.A, a = p % s.A;
        float dx = x[3 * i] - s.R[3 * a], dy = x[3 * i + 1] - s.R[3 * a + 1], dz = x[3 * i + 2] - s.R[3 * a + 2];
        ven -= s.Z[a] / fmaxf(sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz))), 1e-10f);
      }
      for (int p = lane; p < s.N * s.N; p += 32) {
        int i = p / s.N, j = p % s.N;
        if (i < j) {
          float dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
          vee += 1.0f / sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        }
      }
      ven = warp_sum(ven);
      vee = warp_sum(vee);
      if (lane == 0) {
        const float T = -0.5f * (lap + g2);
        e_l[w] = T + ven + vee + s.v_nn;
        if (parts) { parts[w] = T; parts[W + w] = ven; parts[2 * W + w] = vee; parts[3 * W + w] = val + envsum; }
        if (logabs) logabs[w] = val + envsum;
      }
    }
    __syncwarp();
  }
}

// weighted parameter gradient, kind 0.  Block accumulator in shared memory, flushed to a slab.
__global__ void __launch_bounds__(SW * 32) simple_grad_kernel(const DevSys s, const float* __restrict__ theta,
                                                              const float* __restrict__ r, const float* __restrict__ wgt,
                                                              const double* __restrict__ hdr,
                                                              const float* __restrict__ e_l, const int64_t W,
                                                              float* __restrict__ part) {
  extern __shared__ __align__(16) float smem[];
  const int D = 3 * s.N, H0 = s.H, H1 = s.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SimpleSmem m = carve(smem, D, H0, H1, warp);
  const int per_warp = 2 * H0 + D * H0 + MAXD + H1;
  float* acc = m.W3 + H1 + SW * per_warp;      // [P] in theta-relative order of the network block
  const int P = s.P;
  stage_weights(s, theta, m, D, H0, H1);
  for (int i = threadIdx.x; i < P; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  float* h1 = m.scr;
  float* d1 = h1 + H0;
  float* x = h1 + 2 * H0 + D * H0;
  float* d2 = x + MAXD;
  const MlpOff o = s.simple;
  float center = 0.f;
  if (hdr) center = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));
  for (int64_t w = (int64_t)blockIdx.x * SW + warp; w < W; w += (int64_t)gridDim.x * SW) {
    float wt = 1.f;
    if (wgt) wt = wgt[w];
    else if (e_l) { float e = e_l[w]; wt = isfinite(e) ? e - center : 0.f; }
    for (int d = lane; d < D; d += 32) x[d] = r[(int64_t)d * W + w];
    __syncwarp();
    for (int j = lane; j < H0; j += 32) {
      float u = m.b1[j];
      for (int d = 0; d < D; ++d) u = fmaf(m.W1[d * H0 + j], x[d], u);
      h1[j] = nq_tanh(u);
    }
    __syncwarp();
    for (int k = lane; k < H1; k += 32) {
      float u = m.b2[k];
      for (int j = 0; j < H0; ++j) u = fmaf(m.W2[j * H1 + k], h1[j], u);
      float h = nq_tanh(u);
      atomicAdd(&acc[o.W3 + k], wt * h);
      float dd = wt * m.W3[k] * fmaf(-h, h, 1.f);
      d2[k] = dd;
      atomicAdd(&acc[o.b2 + k], dd);
    }
    if (lane == 0) atomicAdd(&acc[o.b3], wt);
    __syncwarp();
    for (int j = lane; j < H0; j += 32) {
      float sacc = 0.f;
      const float hj = h1[j];
      for (int k = 0; k < H1; ++k) {
        sacc = fmaf(m.W2[j * H1 + k], d2[k], sacc);
      }
      const float dj = sacc * fmaf(-hj, hj, 1.f);
      d1[j] = dj;
      atomicAdd(&acc[o.b1 + j], dj);
      for (int d = 0; d < D; ++d) atomicAdd(&acc[o.W1 + d * H0 + j], x[d] * dj);
    }
    __syncwarp();
    // dW2[j][k] += h1[j] * d2[k]: lanes over k (coalesced shared atomics)
    for (int j = 0; j < H0; ++j) {
      const float hj = h1[j];
      for (int k = lane; k < H1; k += 32) atomicAdd(&acc[o.W2 + j * H1 + k], hj * d2[k]);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) part[(int64_t)blockIdx.x * P + i] = acc[i];
}

__global__ void simple_grad_finish_kernel(const float* __restrict__ part, const int nblk, const int P,
                                          double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  double a = 0;
  for (int b = 0; b < nblk; ++b) a += (double)part[(int64_t)b * P + i];
  out[i] = a;
}

size_t simple_smem_bytes(const DevSys& s, bool with_acc) {
  const int D = 3 * s.N, H0 = s.H, H1 = s.H;
  size_t f = (size_t)D * H0 + (size_t)H0 * H1 + H0 + 2 * H1 + (size_t)SW * (2 * H0 + D * H0 + MAXD + H1);
  if (with_acc) f += s.P;
  return f * sizeof(float);
}

}  // namespace

static int simple_launch_eval(nqmc_ctx* ctx, const float* r, int64_t W, float* logabs, float* e_l, float* parts,
                              int mode, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H > MAXH || 3 * s.N > MAXD) {
    ctx->err = "SimpleWavefunction: hidden width > 128 or more than 16 electrons";
    return NQMC_ERR_UNSUPPORTED;
  }
  const size_t smem = simple_smem_bytes(s, false);
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(simple_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int64_t nblk = (W + SW - 1) / SW;
  if (nblk > 8 * ctx->sm_count) nblk = 8 * ctx->sm_count;
  simple_eval_kernel<<<(unsigned)nblk, SW * 32, smem, st>>>(s, ctx->theta, r, W, logabs, e_l, parts, mode);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_simple_logpsi(nqmc_ctx* ctx, const float* r, int64_t W, float* logabs, cudaStream_t st) {
  return simple_launch_eval(ctx, r, W, logabs, nullptr, nullptr, 0, st);
}
int nq_simple_energy(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, cudaStream_t st) {
  return simple_launch_eval(ctx, r, W, nullptr, e_l, parts, 1, st);
}
int nq_simple_grads(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr_center, const float* e_l,
                    int64_t W, double* out, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const size_t smem = simple_smem_bytes(s, true);
  if (smem > 227 * 1024) {
    ctx->err = "SimpleWavefunction gradient: parameter block does not fit in shared memory";
    return NQMC_ERR_UNSUPPORTED;
  }
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(simple_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int64_t nblk = (W + SW - 1) / SW;
  const int64_t cap = ((int64_t)ctx->n_cta_bwd * ctx->p_mlp) / s.P;
  if (nblk > cap) nblk = cap;
  if (nblk > 2 * ctx->sm_count) nblk = 2 * ctx->sm_count;
  if (nblk < 1) nblk = 1;
  simple_grad_kernel<<<(unsigned)nblk, SW * 32, smem, st>>>(s, ctx->theta, r, wgt, hdr_center, e_l, W, ctx->part_mlp);
  NQMC_LAUNCH_CHECK();
  simple_grad_finish_kernel<<<(s.P + 255) / 256, 256, 0, st>>>(ctx->part_mlp, (int)nblk, s.P, out);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
