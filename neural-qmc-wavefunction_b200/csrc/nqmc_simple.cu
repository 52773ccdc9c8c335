// nqmc_simple.cu -- ansatz kind 0: the reference's only *implemented* wavefunction,
//   log|psi| = MLP(flatten r) + sum_i log sum_A exp(-alpha |r_i - R_A|)
// (src/nqmc/wavefunction/simple.py:41-70 SimpleMLP -- any tuple of hidden widths, 'tanh' or 'silu' --
// and :124-164 envelope + __call__), used by BASELINE config 1 (H2, scripts/train_h2.py).
// It is the plumbing case (32 chains), so the mapping is simple: one WARP per walker, lanes across the
// units of a layer, all weights staged once per CTA in shared memory.  The forward Laplacian carries the
// D = 3N input tangents explicitly through every layer (value v, tangents G[d], Laplacian L):
//   u = W^T v + b ; Gu[d] = W^T G[d] ; Lu = W^T L
//   v' = s(u) ; G'[d] = s'(u) Gu[d] ; L' = s'(u) Lu + s''(u) sum_d Gu[d]^2
// (hamiltonian/molecular.py:52-69 obtains the same trace + |grad|^2 from a dense nested-autodiff Hessian).
#include "nqmc_internal.cuh"

namespace {

constexpr int SW = 4;            // warps (walkers in flight) per CTA
constexpr int MAXD = 48;         // 3 * NQMC_MAX_ELEC
constexpr int DCH = 8;           // tangents accumulated per pass over a weight column

__device__ __forceinline__ void act_rules(const int act, const float u, float& h, float& a1, float& a2) {
  if (act == NQMC_ACT_TANH) {
    h = nq_tanh(u);
    a1 = fmaf(-h, h, 1.f);
    a2 = -2.f * h * a1;
  } else {   // silu: u * sigmoid(u)
    const float s = 1.0f / (1.0f + __expf(-u));
    h = u * s;
    a1 = s * fmaf(u, 1.0f - s, 1.0f);
    a2 = s * (1.0f - s) * fmaf(u, 1.0f - 2.0f * s, 2.0f);
  }
}

__device__ __forceinline__ void stage_theta(const DevSys& s, const float* __restrict__ theta, float* sh) {
  for (int i = threadIdx.x; i < s.P; i += blockDim.x) sh[i] = theta[i];
}

// envelope of electron i: value, gradient, Laplacian of log sum_A exp(-alpha d_iA)
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

// per-warp scratch of the evaluation kernel: two buffers [D + 2][HM] (row 0 value, 1..D tangents, D+1 Laplacian)
// + x[MAXD].  HM = widest layer (>= D).
__host__ __device__ inline int simple_hm(const DevSys& s) {
  int hm = 3 * s.N;
  for (int l = 1; l <= s.simple.n_hidden; ++l) hm = s.simple.dim[l] > hm ? s.simple.dim[l] : hm;
  return hm;
}

// mode 0: log|psi| ; mode 1: local energy
__global__ void __launch_bounds__(SW * 32) simple_eval_kernel(const DevSys s, const float* __restrict__ theta,
                                                              const float* __restrict__ r, const int64_t W,
                                                              float* __restrict__ logabs, float* __restrict__ e_l,
                                                              float* __restrict__ parts, const int mode) {
  extern __shared__ __align__(16) float smem[];
  const int D = 3 * s.N, HM = simple_hm(s), NH = s.simple.n_hidden, act = s.simple.act;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* th = smem;
  const int rows = mode == 1 ? D + 2 : 1;
  float* buf0 = smem + s.P + warp * (2 * rows * HM + MAXD);
  float* buf1 = buf0 + rows * HM;
  float* x = buf1 + rows * HM;
  stage_theta(s, theta, th);
  __syncthreads();
  for (int64_t w = (int64_t)blockIdx.x * SW + warp; w < W; w += (int64_t)gridDim.x * SW) {
    for (int d = lane; d < D; d += 32) x[d] = r[(int64_t)d * W + w];
    __syncwarp();
    float* in = buf0;
    float* out = buf1;
    // input "layer": v = x, G = identity, L = 0
    for (int k = lane; k < D; k += 32) {
      in[k] = x[k];
      if (mode == 1) {
        for (int d = 0; d < D; ++d) in[(1 + d) * HM + k] = d == k ? 1.f : 0.f;
        in[(D + 1) * HM + k] = 0.f;
      }
    }
    __syncwarp();
    for (int l = 0; l < NH; ++l) {
      const int K = s.simple.dim[l], M = s.simple.dim[l + 1];
      const float* Wl = th + s.simple.offW[l];
      const float* bl = th + s.simple.offb[l];
      for (int j = lane; j < M; j += 32) {
        float u = bl[j], lu = 0.f;
        for (int k = 0; k < K; ++k) {
          const float wv = Wl[k * M + j];
          u = fmaf(wv, in[k], u);
          if (mode == 1) lu = fmaf(wv, in[(D + 1) * HM + k], lu);
        }
        float h, a1, a2;
        act_rules(act, u, h, a1, a2);
        out[j] = h;
        if (mode == 1) {
          float g2 = 0.f;
          for (int d0 = 0; d0 < D; d0 += DCH) {
            float acc[DCH];
#pragma unroll
            for (int dd = 0; dd < DCH; ++dd) acc[dd] = 0.f;
            for (int k = 0; k < K; ++k) {
              const float wv = Wl[k * M + j];
#pragma unroll
              for (int dd = 0; dd < DCH; ++dd)
                if (d0 + dd < D) acc[dd] = fmaf(wv, in[(1 + d0 + dd) * HM + k], acc[dd]);
            }
#pragma unroll
            for (int dd = 0; dd < DCH; ++dd)
              if (d0 + dd < D) {
                g2 = fmaf(acc[dd], acc[dd], g2);
                out[(1 + d0 + dd) * HM + j] = a1 * acc[dd];
              }
          }
          out[(D + 1) * HM + j] = fmaf(a1, lu, a2 * g2);
        }
      }
      __syncwarp();
      float* t = in; in = out; out = t;
    }
    // output layer Dense(1)
    const int K = s.simple.dim[NH];
    const float* Wo = th + s.simple.offW[NH];
    float val = 0.f, lap = 0.f;
    for (int k = lane; k < K; k += 32) {
      val = fmaf(Wo[k], in[k], val);
      if (mode == 1) lap = fmaf(Wo[k], in[(D + 1) * HM + k], lap);
    }
    val = warp_sum(val) + th[s.simple.offb[NH]];
    float env = 0.f, egx = 0.f, egy = 0.f, egz = 0.f, elap = 0.f;
    if (lane < s.N) envelope(s, x, lane, env, egx, egy, egz, elap);
    const float envsum = warp_sum(env);
    if (mode == 0) {
      if (lane == 0) logabs[w] = val + envsum;
    } else {
      float g2 = 0.f;
      for (int d = 0; d < D; ++d) {
        float gp = 0.f;
        for (int k = lane; k < K; k += 32) gp = fmaf(Wo[k], in[(1 + d) * HM + k], gp);
        const float gm = warp_sum(gp);
        const int i = d / 3, ax = d % 3;
        const float ge = __shfl_sync(0xffffffffu, ax == 0 ? egx : (ax == 1 ? egy : egz), i);
        g2 = fmaf(gm + ge, gm + ge, g2);
      }
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

// weighted parameter gradient, kind 0.  Block accumulator in shared memory, flushed to a per-CTA slab.
// per-warp scratch: h[l] (inputs of every layer) and a1[l] (activation derivative), total 2 * sum(dim), + delta[2][HM].
__host__ __device__ inline int simple_act_floats(const DevSys& s) {
  int n = 0;
  for (int l = 0; l <= s.simple.n_hidden; ++l) n += s.simple.dim[l];
  return n;
}

__global__ void __launch_bounds__(SW * 32) simple_grad_kernel(const DevSys s, const float* __restrict__ theta,
                                                              const float* __restrict__ r, const float* __restrict__ wgt,
                                                              const double* __restrict__ hdr,
                                                              const float* __restrict__ e_l, const int64_t W,
                                                              float* __restrict__ part, float* __restrict__ Ot,
                                                              const int64_t ld) {
  // Ot != nullptr: per-walker mode (Stochastic Reconfiguration): instead of accumulating weight * d log|psi| / d theta,
  // every walker's log-derivatives are stored parameter-major, Ot[p * ld + w] (the warps of a block own consecutive
  // walkers, so the 4-byte stores of a parameter fall into one sector per block).
  extern __shared__ __align__(16) float smem[];
  const int HM = simple_hm(s), NH = s.simple.n_hidden, act = s.simple.act, NA = simple_act_floats(s);
  auto emit = [&](const int idx, const float v, const int64_t w) {
    if (Ot != nullptr) Ot[(int64_t)idx * ld + w] = v;
    else atomicAdd(&smem[s.P + idx], v);
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = s.P;
  float* th = smem;
  float* acc = smem + P;
  float* hbuf = acc + P + warp * (2 * NA + 2 * HM);
  float* abuf = hbuf + NA;
  float* dl0 = abuf + NA;
  float* dl1 = dl0 + HM;
  stage_theta(s, theta, th);
  for (int i = threadIdx.x; i < P; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  float center = 0.f;
  if (hdr) center = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));
  for (int64_t w = (int64_t)blockIdx.x * SW + warp; w < W; w += (int64_t)gridDim.x * SW) {
    float wt = 1.f;
    if (Ot != nullptr) wt = 1.f;
    else if (wgt) wt = wgt[w];
    else if (e_l) { float e = e_l[w]; wt = isfinite(e) ? e - center : 0.f; }
    const int D = s.simple.dim[0];
    for (int d = lane; d < D; d += 32) hbuf[d] = r[(int64_t)d * W + w];
    __syncwarp();
    int off = 0;
    for (int l = 0; l < NH; ++l) {           // forward, keeping every layer's input and s'(u)
      const int K = s.simple.dim[l], M = s.simple.dim[l + 1];
      const float* Wl = th + s.simple.offW[l];
      const float* bl = th + s.simple.offb[l];
      const float* in = hbuf + off;
      for (int j = lane; j < M; j += 32) {
        float u = bl[j];
        for (int k = 0; k < K; ++k) u = fmaf(Wl[k * M + j], in[k], u);
        float h, a1, a2;
        act_rules(act, u, h, a1, a2);
        hbuf[off + K + j] = h;
        abuf[off + K + j] = a1;
      }
      off += K;
      __syncwarp();
    }
    // output layer: d/dW_out = wt * h_last, d/db_out = wt ; delta_last = wt * W_out * s'(u_last)
    {
      const int K = s.simple.dim[NH];
      const float* Wo = th + s.simple.offW[NH];
      for (int k = lane; k < K; k += 32) {
        emit(s.simple.offW[NH] + k, wt * hbuf[off + k], w);
        dl0[k] = wt * Wo[k] * (NH > 0 ? abuf[off + k] : 1.f);
      }
      if (lane == 0) emit(s.simple.offb[NH], wt, w);
      __syncwarp();
    }
    float* dcur = dl0;
    float* dnext = dl1;
    for (int l = NH - 1; l >= 0; --l) {      // reverse
      const int K = s.simple.dim[l], M = s.simple.dim[l + 1];
      off -= K;
      const float* Wl = th + s.simple.offW[l];
      const float* in = hbuf + off;
      for (int j = lane; j < M; j += 32) emit(s.simple.offb[l] + j, dcur[j], w);
      for (int k = 0; k < K; ++k) {           // dW[k][j] += in[k] * delta[j]: lanes over j (conflict-free atomics)
        const float hk = in[k];
        for (int j = lane; j < M; j += 32) emit(s.simple.offW[l] + k * M + j, hk * dcur[j], w);
      }
      if (l > 0) {
        for (int k = lane; k < K; k += 32) {
          float sacc = 0.f;
          for (int j = 0; j < M; ++j) sacc = fmaf(Wl[k * M + j], dcur[j], sacc);
          dnext[k] = sacc * abuf[off + k];
        }
      }
      __syncwarp();
      float* t = dcur; dcur = dnext; dnext = t;
    }
    __syncwarp();
  }
  __syncthreads();
  if (Ot == nullptr)
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

size_t simple_eval_smem(const DevSys& s, int mode) {
  const int rows = mode == 1 ? 3 * s.N + 2 : 1;
  return ((size_t)s.P + (size_t)SW * (2 * rows * simple_hm(s) + MAXD)) * sizeof(float);
}
size_t simple_grad_smem(const DevSys& s) {
  return (2 * (size_t)s.P + (size_t)SW * (2 * simple_act_floats(s) + 2 * simple_hm(s))) * sizeof(float);
}

}  // namespace

static int simple_launch_eval(nqmc_ctx* ctx, const float* r, int64_t W, float* logabs, float* e_l, float* parts,
                              int mode, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const size_t smem = simple_eval_smem(s, mode);
  if (smem > 227 * 1024) {
    ctx->err = "SimpleWavefunction: network + per-walker tangent scratch exceed 227 KB of shared memory";
    return NQMC_ERR_UNSUPPORTED;
  }
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
  const size_t smem = simple_grad_smem(s);
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
  simple_grad_kernel<<<(unsigned)nblk, SW * 32, smem, st>>>(s, ctx->theta, r, wgt, hdr_center, e_l, W, ctx->part_mlp, nullptr, 0);
  NQMC_LAUNCH_CHECK();
  simple_grad_finish_kernel<<<(s.P + 255) / 256, 256, 0, st>>>(ctx->part_mlp, (int)nblk, s.P, out);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

// per-walker log-derivatives, parameter-major: Ot[p * ld + w] = d log|psi(r_w)| / d theta_p   (SimpleWavefunction)
int nq_simple_grads_per_walker(nqmc_ctx* ctx, const float* r, int64_t W, float* Ot, int64_t ld, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const size_t smem = simple_grad_smem(s);
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
  if (nblk > 4 * ctx->sm_count) nblk = 4 * ctx->sm_count;
  simple_grad_kernel<<<(unsigned)nblk, SW * 32, smem, st>>>(s, ctx->theta, r, nullptr, nullptr, nullptr, W, nullptr, Ot, ld);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
