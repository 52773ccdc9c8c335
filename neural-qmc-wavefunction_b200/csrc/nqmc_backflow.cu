// nqmc_backflow.cu -- BackflowTransform (.github/phase3-issue-body.md:134-169) and the local-energy
// assembly when orbitals are evaluated at backflow coordinates
//     x_i = r_i + sum_{j != i} eta(r_ij) (r_i - r_j),  eta(r) = MLP([r, e^-r]; 16,16,1; tanh) e^{-r/2}.
// Every orbital row now depends on all 3N coordinates.  Instead of 3N+2 tangent channels the orbital
// pass carries 10 (value, gradient, Hessian w.r.t. the orbital's own 3-vector input) and this file
// contracts them with the backflow Jacobian per walker:
//     d_mu log|det|  = tr(Phi^-1 d_mu Phi),   d2_mu log|det| = tr(Phi^-1 d2_mu Phi) - tr((Phi^-1 d_mu Phi)^2)
//     d_mu Phi[i,k]  = g_ik . d_mu x_i
//     sum_mu d2_mu Phi[i,k] = sum_ab Hess_ik[ab] G_i[ab] + g_ik . Lap x_i,   G_i = sum_mu d_mu x_i d_mu x_i^T
// Round 2: the orbital pass no longer needs the six Hessian channels.  Only the contraction sum_ab Hess_ik[ab] G_i[ab]
// enters E_L, and G_i depends on the backflow Jacobian alone, so backflow_xg_kernel computes G_i BEFORE the orbital
// pass and the orbital kernels propagate the G-weighted Laplacian L_G = sum_ab G_ab d_a d_b forward instead -- it obeys
// the same layer rules as the ordinary Laplacian (linear: L_G(W^T h) = W^T L_G(h); tanh: t' L_G(u) + t'' grad u^T G grad u)
// -- i.e. 5 channels (value, gradient, L_G) instead of 10: half the flops, and the 5-channel tensor-core kernels apply.
// The 10-channel route stays for the FP32 FFMA parity anchor (NQMC_TC=0).
// One thread owns one walker; pair tables and orbital derivative tables live in per-thread local
// memory (L1-resident), which is acceptable because this stage is a few % of the orbital GEMM work.
#include "nqmc_internal.cuh"

namespace {

constexpr int WT = 128;
constexpr int BH = NQMC_BF_H;   // 16

// eta network with derivatives w.r.t. the scalar distance d.  LEVEL 0: eta ; 1: eta, eta' ; 2: eta, eta', eta''
template <int LEVEL>
__device__ __forceinline__ void eta_eval(const float* __restrict__ theta, const MlpOff& o, const float d, float& eta,
                                         float& etap, float& etapp) {
  const float e = __expf(-d);
  float h[BH], hp[BH], hpp[BH];
#pragma unroll
  for (int j = 0; j < BH; ++j) {
    const float w0 = theta[o.W1 + j], w1 = theta[o.W1 + BH + j];
    const float u = fmaf(w0, d, fmaf(w1, e, theta[o.b1 + j]));
    const float t = nq_tanh(u);
    h[j] = t;
    if (LEVEL >= 1) {
      const float tp = fmaf(-t, t, 1.f), up = w0 - w1 * e, upp = w1 * e;
      hp[j] = tp * up;
      if (LEVEL == 2) hpp[j] = fmaf(tp, upp, -2.f * t * tp * up * up);
    }
  }
  float m = theta[o.b3], mp = 0.f, mpp = 0.f;
#pragma unroll 4
  for (int k = 0; k < BH; ++k) {
    float u = theta[o.b2 + k], up = 0.f, upp = 0.f;
#pragma unroll
    for (int j = 0; j < BH; ++j) {
      const float w = theta[o.W2 + j * BH + k];
      u = fmaf(w, h[j], u);
      if (LEVEL >= 1) up = fmaf(w, hp[j], up);
      if (LEVEL == 2) upp = fmaf(w, hpp[j], upp);
    }
    const float t = nq_tanh(u), w3 = theta[o.W3 + k];
    m = fmaf(w3, t, m);
    if (LEVEL >= 1) {
      const float tp = fmaf(-t, t, 1.f);
      mp = fmaf(w3, tp * up, mp);
      if (LEVEL == 2) mpp = fmaf(w3, fmaf(tp, upp, -2.f * t * tp * up * up), mpp);
    }
  }
  const float E = __expf(-0.5f * d);
  eta = m * E;
  if (LEVEL >= 1) etap = (mp - 0.5f * m) * E;
  if (LEVEL == 2) etapp = (mpp - mp + 0.25f * m) * E;
}

// x = r + sum_j eta(r_ij)(r_i - r_j)
__global__ void __launch_bounds__(WT) backflow_x_kernel(const DevSys s, const float* __restrict__ theta,
                                                        const float* __restrict__ r, const int64_t W,
                                                        float* __restrict__ x) {
  const int64_t w = (int64_t)blockIdx.x * WT + threadIdx.x;
  if (w >= W) return;
  float px[NQMC_MAX_ELEC], py[NQMC_MAX_ELEC], pz[NQMC_MAX_ELEC], ax[NQMC_MAX_ELEC], ay[NQMC_MAX_ELEC], az[NQMC_MAX_ELEC];
  for (int i = 0; i < s.N; ++i) {
    px[i] = ax[i] = r[(int64_t)(3 * i) * W + w];
    py[i] = ay[i] = r[(int64_t)(3 * i + 1) * W + w];
    pz[i] = az[i] = r[(int64_t)(3 * i + 2) * W + w];
  }
  for (int i = 0; i < s.N; ++i)
    for (int j = i + 1; j < s.N; ++j) {
      const float dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
      const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      float eta, t1, t2;
      eta_eval<0>(theta, s.bf, d, eta, t1, t2);
      ax[i] = fmaf(eta, dx, ax[i]); ay[i] = fmaf(eta, dy, ay[i]); az[i] = fmaf(eta, dz, az[i]);
      ax[j] = fmaf(-eta, dx, ax[j]); ay[j] = fmaf(-eta, dy, ay[j]); az[j] = fmaf(-eta, dz, az[j]);
    }
  for (int i = 0; i < s.N; ++i) {
    x[(int64_t)(3 * i) * W + w] = ax[i];
    x[(int64_t)(3 * i + 1) * W + w] = ay[i];
    x[(int64_t)(3 * i + 2) * W + w] = az[i];
  }
}

// x as above, plus G_i = sum_mu d_mu x_i d_mu x_i^T (symmetric 3x3, packed xx xy xz yy yz zz) for every electron:
//   d x_i / d r_i = A_i = S_i I + T_i,  S_i = 1 + sum_j eta_ij,  T_i = sum_j kappa_ij u u^T  (kappa = eta' d)
//   d x_i / d r_j = -(eta_ij I + kappa_ij u u^T)
//   G_i = A_i^2 + sum_j (eta^2 I + (2 eta kappa + kappa^2) u u^T)
__global__ void __launch_bounds__(WT) backflow_xg_kernel(const DevSys s, const float* __restrict__ theta,
                                                         const float* __restrict__ r, const int64_t W,
                                                         float* __restrict__ x, float* __restrict__ G) {
  const int64_t w = (int64_t)blockIdx.x * WT + threadIdx.x;
  if (w >= W) return;
  const int N = s.N;
  float px[NQMC_MAX_ELEC], py[NQMC_MAX_ELEC], pz[NQMC_MAX_ELEC], ax[NQMC_MAX_ELEC], ay[NQMC_MAX_ELEC], az[NQMC_MAX_ELEC];
  float S[NQMC_MAX_ELEC], T[NQMC_MAX_ELEC][6], Gd[NQMC_MAX_ELEC][6];
  for (int i = 0; i < N; ++i) {
    px[i] = ax[i] = r[(int64_t)(3 * i) * W + w];
    py[i] = ay[i] = r[(int64_t)(3 * i + 1) * W + w];
    pz[i] = az[i] = r[(int64_t)(3 * i + 2) * W + w];
    S[i] = 1.f;
#pragma unroll
    for (int c = 0; c < 6; ++c) { T[i][c] = 0.f; Gd[i][c] = 0.f; }
  }
  for (int i = 0; i < N; ++i)
    for (int j = i + 1; j < N; ++j) {
      const float dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
      const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      float eta, etap, t2;
      eta_eval<1>(theta, s.bf, d, eta, etap, t2);
      ax[i] = fmaf(eta, dx, ax[i]); ay[i] = fmaf(eta, dy, ay[i]); az[i] = fmaf(eta, dz, az[i]);
      ax[j] = fmaf(-eta, dx, ax[j]); ay[j] = fmaf(-eta, dy, ay[j]); az[j] = fmaf(-eta, dz, az[j]);
      const float id = 1.0f / d, u[3] = {dx * id, dy * id, dz * id};
      const float kap = etap * d;
      S[i] += eta; S[j] += eta;
      const float off = fmaf(2.0f * eta, kap, kap * kap);
      int c = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) {
          const float uu = u[a] * u[b];
          T[i][c] = fmaf(kap, uu, T[i][c]); T[j][c] = fmaf(kap, uu, T[j][c]);
          const float g = fmaf(off, uu, (a == b) ? eta * eta : 0.f);
          Gd[i][c] += g; Gd[j][c] += g;
          ++c;
        }
    }
  for (int i = 0; i < N; ++i) {
    x[(int64_t)(3 * i) * W + w] = ax[i];
    x[(int64_t)(3 * i + 1) * W + w] = ay[i];
    x[(int64_t)(3 * i + 2) * W + w] = az[i];
    const float A[6] = {S[i] + T[i][0], T[i][1], T[i][2], S[i] + T[i][3], T[i][4], S[i] + T[i][5]};
    float* g = G + (int64_t)(6 * i) * W + w;
    g[0] = Gd[i][0] + A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
    g[W] = Gd[i][1] + A[0] * A[1] + A[1] * A[3] + A[2] * A[4];
    g[2 * W] = Gd[i][2] + A[0] * A[2] + A[1] * A[4] + A[2] * A[5];
    g[3 * W] = Gd[i][3] + A[1] * A[1] + A[3] * A[3] + A[4] * A[4];
    g[4 * W] = Gd[i][4] + A[1] * A[2] + A[3] * A[4] + A[4] * A[5];
    g[5 * W] = Gd[i][5] + A[2] * A[2] + A[4] * A[4] + A[5] * A[5];
  }
}

// Gauss-Jordan inverse with partial pivoting on a runtime n x n matrix in local memory (n <= 8)
__device__ void inv_slogdet_dyn(const int n, float* a, float* inv, float& sign, float& logabs) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = (i == j) ? 1.f : 0.f;
  sign = 1.f;
  logabs = 0.f;
  for (int c = 0; c < n; ++c) {
    int p = c;
    float best = fabsf(a[c * n + c]);
    for (int rw = c + 1; rw < n; ++rw) {
      const float v = fabsf(a[rw * n + c]);
      if (v > best) { best = v; p = rw; }
    }
    if (p != c) {
      for (int j = 0; j < n; ++j) {
        float t = a[c * n + j]; a[c * n + j] = a[p * n + j]; a[p * n + j] = t;
        t = inv[c * n + j]; inv[c * n + j] = inv[p * n + j]; inv[p * n + j] = t;
      }
      sign = -sign;
    }
    const float piv = a[c * n + c];
    if (piv < 0.f) sign = -sign;
    logabs += logf(fabsf(piv));
    const float ip = 1.0f / piv;
    for (int j = 0; j < n; ++j) { a[c * n + j] *= ip; inv[c * n + j] *= ip; }
    for (int rw = 0; rw < n; ++rw)
      if (rw != c) {
        const float f = a[rw * n + c];
        for (int j = 0; j < n; ++j) {
          a[rw * n + j] = fmaf(-f, a[c * n + j], a[rw * n + j]);
          inv[rw * n + j] = fmaf(-f, inv[c * n + j], inv[rw * n + j]);
        }
      }
  }
}

__device__ __forceinline__ int pair_index(int i, int j, int N) {   // i < j
  return i * N - (i * (i + 1)) / 2 + (j - i - 1);
}

// Jastrow + cusp gradient / Laplacian / potentials on physical r (dynamic-N variant of nqmc_walker.cu's)
__device__ void jastrow_cusp_dyn(const DevSys& s, const float* __restrict__ theta, const float* px, const float* py,
                                 const float* pz, float& J, float* gx, float* gy, float* gz, float& lap, float& ven,
                                 float& vee) {
  const float a_ee = theta[s.jas_a_ee], b_ee = theta[s.jas_b_ee];
  const float a_en = theta[s.jas_a_en], b_en = theta[s.jas_b_en];
  const float cb = s.use_cusp ? fabsf(theta[s.ee_cusp_b]) : 0.f;
  J = 0.f; lap = 0.f; ven = 0.f; vee = 0.f;
  for (int i = 0; i < s.N; ++i) {
    for (int j = i + 1; j < s.N; ++j) {
      const float dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
      const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      float q = 1.0f / fmaf(b_ee, d, 1.0f);
      float u = a_ee * d * q, up = a_ee * q * q, upp = -2.0f * a_ee * b_ee * q * q * q;
      if (s.use_cusp) {
        const float c = ((i < s.n_up) == (j < s.n_up)) ? s.cusp_same : s.cusp_diff;
        q = 1.0f / fmaf(cb, d, 1.0f);
        u += c * d * q; up += c * q * q; upp += -2.0f * c * cb * q * q * q;
      }
      J += u;
      const float id = 1.0f / d, c = up * id;
      gx[i] = fmaf(c, dx, gx[i]); gy[i] = fmaf(c, dy, gy[i]); gz[i] = fmaf(c, dz, gz[i]);
      gx[j] = fmaf(-c, dx, gx[j]); gy[j] = fmaf(-c, dy, gy[j]); gz[j] = fmaf(-c, dz, gz[j]);
      lap += 2.0f * (upp + 2.0f * c);
      vee += id;
    }
    for (int a = 0; a < s.A; ++a) {
      const float dx = px[i] - s.R[3 * a], dy = py[i] - s.R[3 * a + 1], dz = pz[i] - s.R[3 * a + 2];
      const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      float q = 1.0f / fmaf(b_en, d, 1.0f);
      float u = a_en * d * q, up = a_en * q * q, upp = -2.0f * a_en * b_en * q * q * q;
      if (s.use_cusp) {
        const float e = __expf(-d);
        const float* W0 = theta + s.en_cusp_W0[a];
        const float* b0 = theta + s.en_cusp_b0[a];
        const float* W1 = theta + s.en_cusp_W1[a];
        float g = theta[s.en_cusp_b1[a]], gp = 0.f, gpp = 0.f;
#pragma unroll 4
        for (int h = 0; h < NQMC_CUSP_H; ++h) {
          const float w0 = W0[h], w1 = W0[NQMC_CUSP_H + h], w2 = W0[2 * NQMC_CUSP_H + h];
          const float t = nq_tanh(fmaf(w0, d, fmaf(w1, d * d, fmaf(w2, e, b0[h]))));
          const float wo = W1[h], tp = fmaf(-t, t, 1.0f);
          const float zp = fmaf(2.0f * w1, d, w0) - w2 * e, zpp = fmaf(w2, e, 2.0f * w1);
          g = fmaf(wo, t, g);
          gp = fmaf(wo, tp * zp, gp);
          gpp = fmaf(wo, tp * zpp - 2.0f * t * tp * zp * zp, gpp);
        }
        u += -s.Z[a] * d + d * g;
        up += -s.Z[a] + fmaf(d, gp, g);
        upp += fmaf(d, gpp, 2.0f * gp);
      }
      J += u;
      const float id = 1.0f / d, c = up * id;
      gx[i] = fmaf(c, dx, gx[i]); gy[i] = fmaf(c, dy, gy[i]); gz[i] = fmaf(c, dz, gz[i]);
      lap += upp + 2.0f * c;
      ven -= s.Z[a] / fmaxf(d, 1e-10f);
    }
  }
}

constexpr int MAXP = NQMC_MAX_ELEC * (NQMC_MAX_ELEC - 1) / 2;

// E_L with backflow.  phi holds, per orbital-matrix entry evaluated at x(r), either 10 channels (value, gradient,
// Hessian; gw == 0) or 5 (value, gradient, G-weighted Laplacian; gw == 1).
__global__ void __launch_bounds__(WT) assemble_bf_kernel(const DevSys s, const float* __restrict__ theta,
                                                         const float* __restrict__ r, const float* __restrict__ phi,
                                                         float* __restrict__ inv_out, float* __restrict__ v_out,
                                                         const int64_t W, float* __restrict__ e_l,
                                                         float* __restrict__ parts, const int gw) {
  const int64_t w = (int64_t)blockIdx.x * WT + threadIdx.x;
  if (w >= W) return;
  const int N = s.N;
  float px[NQMC_MAX_ELEC], py[NQMC_MAX_ELEC], pz[NQMC_MAX_ELEC];
  for (int i = 0; i < N; ++i) {
    px[i] = r[(int64_t)(3 * i) * W + w];
    py[i] = r[(int64_t)(3 * i + 1) * W + w];
    pz[i] = r[(int64_t)(3 * i + 2) * W + w];
  }
  // ---- pair tables ------------------------------------------------------------------------
  float p_eta[MAXP], p_kap[MAXP], p_u[MAXP][3];
  float S[NQMC_MAX_ELEC], T[NQMC_MAX_ELEC][6], lapx[NQMC_MAX_ELEC][3], Gd[NQMC_MAX_ELEC][6];
  for (int i = 0; i < N; ++i) {
    S[i] = 1.f;
#pragma unroll
    for (int c = 0; c < 6; ++c) { T[i][c] = 0.f; Gd[i][c] = 0.f; }
    lapx[i][0] = lapx[i][1] = lapx[i][2] = 0.f;
  }
  for (int i = 0; i < N; ++i)
    for (int j = i + 1; j < N; ++j) {
      const int p = pair_index(i, j, N);
      const float dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
      const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      float eta, etap, etapp;
      eta_eval<2>(theta, s.bf, d, eta, etap, etapp);
      const float id = 1.0f / d, u[3] = {dx * id, dy * id, dz * id};
      const float kap = etap * d, lf = 2.0f * fmaf(d, etapp, 4.0f * etap);
      p_eta[p] = eta; p_kap[p] = kap;
      p_u[p][0] = u[0]; p_u[p][1] = u[1]; p_u[p][2] = u[2];
      S[i] += eta; S[j] += eta;
      const float off = fmaf(2.0f * eta, kap, kap * kap);      // (2 eta kappa + kappa^2) u u^T  in G of the *other* electron
      int c = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) {
          const float uu = u[a] * u[b];
          T[i][c] = fmaf(kap, uu, T[i][c]); T[j][c] = fmaf(kap, uu, T[j][c]);
          const float g = fmaf(off, uu, (a == b) ? eta * eta : 0.f);
          Gd[i][c] += g; Gd[j][c] += g;
          ++c;
        }
#pragma unroll
      for (int a = 0; a < 3; ++a) { lapx[i][a] = fmaf(lf, u[a], lapx[i][a]); lapx[j][a] = fmaf(-lf, u[a], lapx[j][a]); }
    }
  // A_i = S_i I + T_i (symmetric, packed xx xy xz yy yz zz);  G_i = A_i^2 + Gd_i
  float Am[NQMC_MAX_ELEC][6];
  for (int i = 0; i < N; ++i) {
    Am[i][0] = S[i] + T[i][0]; Am[i][1] = T[i][1]; Am[i][2] = T[i][2];
    Am[i][3] = S[i] + T[i][3]; Am[i][4] = T[i][4]; Am[i][5] = S[i] + T[i][5];
    const float* A = Am[i];
    Gd[i][0] += A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
    Gd[i][1] += A[0] * A[1] + A[1] * A[3] + A[2] * A[4];
    Gd[i][2] += A[0] * A[2] + A[1] * A[4] + A[2] * A[5];
    Gd[i][3] += A[1] * A[1] + A[3] * A[3] + A[4] * A[4];
    Gd[i][4] += A[1] * A[2] + A[3] * A[4] + A[4] * A[5];
    Gd[i][5] += A[2] * A[2] + A[4] * A[4] + A[5] * A[5];
  }
  // ---- Slater part ------------------------------------------------------------------------
  float logabs = 0.f, sign = 1.f, term1 = 0.f, term2 = 0.f;
  float vx[NQMC_MAX_ELEC], vy[NQMC_MAX_ELEC], vz[NQMC_MAX_ELEC];
  float ginv[2][NQMC_MAX_NS * NQMC_MAX_NS];                       // inverse per spin
  float gtab[2][NQMC_MAX_NS * NQMC_MAX_NS][3];                    // g_ik per spin
  const int64_t cs = (int64_t)s.n_ent * W;
  for (int spin = 0; spin < 2; ++spin) {
    const int n = spin == 0 ? s.n_up : s.n_dn;
    if (n == 0) continue;
    const int e0 = spin == 0 ? 0 : s.n_up;
    const int base = spin == 0 ? 0 : s.n_up * s.n_up;
    float a[NQMC_MAX_NS * NQMC_MAX_NS];
    for (int q = 0; q < n * n; ++q) a[q] = phi[(int64_t)(base + q) * W + w];
    float sg, la;
    inv_slogdet_dyn(n, a, ginv[spin], sg, la);
    sign *= sg;
    logabs += la;
    const float* inv = ginv[spin];                                // inv[k*n + i]
    if (inv_out)
      for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) inv_out[(int64_t)(base + i * n + k) * W + w] = inv[k * n + i];
    for (int i = 0; i < n; ++i) {
      float sx = 0.f, sy = 0.f, sz = 0.f;
      const int e = e0 + i;
      for (int k = 0; k < n; ++k) {
        const int64_t eidx = (int64_t)(base + i * n + k) * W + w;
        const float c = inv[k * n + i];
        const float g0 = phi[eidx + cs], g1 = phi[eidx + 2 * cs], g2 = phi[eidx + 3 * cs];
        gtab[spin][i * n + k][0] = g0; gtab[spin][i * n + k][1] = g1; gtab[spin][i * n + k][2] = g2;
        sx = fmaf(c, g0, sx); sy = fmaf(c, g1, sy); sz = fmaf(c, g2, sz);
        // sum_ab Hess[ab] G[ab] (off-diagonals count twice) + g . Lap x
        const float* G = Gd[e];
        float hh;
        if (gw) hh = phi[eidx + 4 * cs];                          // the orbital pass already contracted Hess with G
        else hh = phi[eidx + 4 * cs] * G[0] + phi[eidx + 7 * cs] * G[3] + phi[eidx + 9 * cs] * G[5] +
                  2.0f * (phi[eidx + 5 * cs] * G[1] + phi[eidx + 6 * cs] * G[2] + phi[eidx + 8 * cs] * G[4]);
        hh += g0 * lapx[e][0] + g1 * lapx[e][1] + g2 * lapx[e][2];
        term1 = fmaf(c, hh, term1);
      }
      vx[e] = sx; vy[e] = sy; vz[e] = sz;
    }
  }
  if (v_out)
    for (int e = 0; e < N; ++e) {
      v_out[(int64_t)(3 * e) * W + w] = vx[e];
      v_out[(int64_t)(3 * e + 1) * W + w] = vy[e];
      v_out[(int64_t)(3 * e + 2) * W + w] = vz[e];
    }
  // ---- gradient of log|det| and the tr((Phi^-1 d_mu Phi)^2) term, coordinate by coordinate -------
  float gx[NQMC_MAX_ELEC], gy[NQMC_MAX_ELEC], gz[NQMC_MAX_ELEC];
  float gdet2 = 0.f;
  for (int e = 0; e < N; ++e) {
    float gcol[3];
    for (int c = 0; c < 3; ++c) {
      // Jcol_i = d x_i / d r_(e,c)  for every electron i
      float jc[NQMC_MAX_ELEC][3];
      for (int i = 0; i < N; ++i) {
        if (i == e) {
          const float* A = Am[e];
          jc[i][0] = c == 0 ? A[0] : (c == 1 ? A[1] : A[2]);
          jc[i][1] = c == 0 ? A[1] : (c == 1 ? A[3] : A[4]);
          jc[i][2] = c == 0 ? A[2] : (c == 1 ? A[4] : A[5]);
        } else {
          const int p = i < e ? pair_index(i, e, N) : pair_index(e, i, N);
          const float ku = p_kap[p] * p_u[p][c];
          jc[i][0] = -ku * p_u[p][0] - (c == 0 ? p_eta[p] : 0.f);
          jc[i][1] = -ku * p_u[p][1] - (c == 1 ? p_eta[p] : 0.f);
          jc[i][2] = -ku * p_u[p][2] - (c == 2 ? p_eta[p] : 0.f);
        }
      }
      float gsum = 0.f;
      for (int i = 0; i < N; ++i) gsum += vx[i] * jc[i][0] + vy[i] * jc[i][1] + vz[i] * jc[i][2];
      gcol[c] = gsum;
      for (int spin = 0; spin < 2; ++spin) {
        const int n = spin == 0 ? s.n_up : s.n_dn;
        if (n == 0) continue;
        const int e0 = spin == 0 ? 0 : s.n_up;
        const float* inv = ginv[spin];
        float D[NQMC_MAX_NS * NQMC_MAX_NS], M[NQMC_MAX_NS * NQMC_MAX_NS];
        for (int i = 0; i < n; ++i)
          for (int k = 0; k < n; ++k) {
            const float* g = gtab[spin][i * n + k];
            D[i * n + k] = g[0] * jc[e0 + i][0] + g[1] * jc[e0 + i][1] + g[2] * jc[e0 + i][2];
          }
        for (int k = 0; k < n; ++k)
          for (int l = 0; l < n; ++l) {
            float m = 0.f;
            for (int i = 0; i < n; ++i) m = fmaf(inv[k * n + i], D[i * n + l], m);
            M[k * n + l] = m;
          }
        for (int k = 0; k < n; ++k)
          for (int l = 0; l < n; ++l) term2 = fmaf(M[k * n + l], M[l * n + k], term2);
      }
    }
    gx[e] = gcol[0]; gy[e] = gcol[1]; gz[e] = gcol[2];
    gdet2 += gcol[0] * gcol[0] + gcol[1] * gcol[1] + gcol[2] * gcol[2];
  }
  (void)gdet2;
  float J, lapJ, ven, vee;
  jastrow_cusp_dyn(s, theta, px, py, pz, J, gx, gy, gz, lapJ, ven, vee);
  logabs += J;
  float g2 = 0.f;
  for (int e = 0; e < N; ++e) g2 += gx[e] * gx[e] + gy[e] * gy[e] + gz[e] * gz[e];
  const float T_kin = -0.5f * ((term1 - term2) + lapJ + g2);
  const float e_loc = T_kin + ven + vee + s.v_nn;
  if (e_l) e_l[w] = e_loc;
  if (parts) { parts[w] = T_kin; parts[W + w] = ven; parts[2 * W + w] = vee; parts[3 * W + w] = logabs; }
}

// ---------------------------------------------------------------------------------------------
// weighted parameter gradient of the eta network:
//   d log|det| / d eta_p = (v_i - v_j) . (r_i - r_j)   =>   rows = (walker, pair), a 2-16-16-1 reverse pass.
// Phase 1: one thread per row computes h1, delta2, delta1 into shared memory.
// Phase 2: register-blocked: the 256 threads are 16 row groups x 16 threads; thread (g, b) owns the 4 x 4 block
// (4 (b / 4), 4 (b % 4)) of dW2 and column b of the small leaves for rows g, g + 16, ... of the chunk -- two LDS.128
// feed 16 FMAs (round 1: thread (j, k) took every row of the chunk with two scalar loads per FMA, and three warps walked
// all rows again for the small leaves: 264 us at 32,768 walkers, H6).  The 16 groups are summed in shared memory once,
// after the last chunk.
constexpr int BT = 256;          // threads per CTA
constexpr int BR = 128;          // rows per chunk (static shared memory stays under 48 KB)
constexpr int BF_P = 2 * BH + BH + BH * BH + BH + BH + 1;   // 337

__global__ void __launch_bounds__(BT) bf_grads_kernel(const DevSys s, const float* __restrict__ theta,
                                                      const float* __restrict__ r, const float* __restrict__ vvec,
                                                      const float* __restrict__ wgt, const double* __restrict__ hdr,
                                                      const float* __restrict__ e_l, const int64_t W,
                                                      float* __restrict__ part) {
  constexpr int LD = BH + 4;       // row stride of the staging arrays: 16-byte aligned rows for the LDS.128 of phase 2
  __shared__ __align__(16) float sh_h1[BR][LD], sh_d2[BR][LD], sh_d1[BR][LD], sh_sh2[BR][LD];
  __shared__ float sh_f[BR][2], sh_seed[BR];
  __shared__ float sh_red[BF_P];
  const MlpOff o = s.bf;
  const int tid = threadIdx.x;
  const int N = s.N, NP = N * (N - 1) / 2;
  const int64_t n_rows = W * NP;
  float center = 0.f;
  if (hdr) center = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));
  const int grp = tid >> 4, b16 = tid & 15, jb = 4 * (b16 >> 2), kb = 4 * (b16 & 3);
  float accW2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) accW2[a][c] = 0.f;
  float a_b1 = 0.f, a_b2 = 0.f, a_w3 = 0.f, a_w10 = 0.f, a_w11 = 0.f, a_b3 = 0.f;   // small leaves, column b16
  for (int q = tid; q < BF_P; q += BT) sh_red[q] = 0.f;
  for (int64_t row0 = (int64_t)blockIdx.x * BR; row0 < n_rows; row0 += (int64_t)gridDim.x * BR) {
    // ---- phase 1 -------------------------------------------------------------------------
    if (tid < BR) {
      const int64_t row = row0 + tid;
      float seed = 0.f, d = 1.f;
      if (row < n_rows) {
        // rows are ordered pair-major so that consecutive threads read consecutive walkers (coalesced)
        const int p = (int)(row / W);
        const int64_t w = row % W;
        int i = 0, acc = 0;
        while (acc + (N - 1 - i) <= p) { acc += N - 1 - i; ++i; }
        const int j = i + 1 + (p - acc);
        const float dx = r[(int64_t)(3 * i) * W + w] - r[(int64_t)(3 * j) * W + w];
        const float dy = r[(int64_t)(3 * i + 1) * W + w] - r[(int64_t)(3 * j + 1) * W + w];
        const float dz = r[(int64_t)(3 * i + 2) * W + w] - r[(int64_t)(3 * j + 2) * W + w];
        d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        float wt = 1.f;
        if (wgt) wt = wgt[w];
        else if (e_l) { const float ev = e_l[w]; wt = isfinite(ev) ? ev - center : 0.f; }
        const float sv = (vvec[(int64_t)(3 * i) * W + w] - vvec[(int64_t)(3 * j) * W + w]) * dx +
                         (vvec[(int64_t)(3 * i + 1) * W + w] - vvec[(int64_t)(3 * j + 1) * W + w]) * dy +
                         (vvec[(int64_t)(3 * i + 2) * W + w] - vvec[(int64_t)(3 * j + 2) * W + w]) * dz;
        seed = wt * sv * __expf(-0.5f * d);
        if (!isfinite(seed)) seed = 0.f;
      }
      const float e = __expf(-d);
      float h1[BH], h2[BH], d2[BH];
#pragma unroll
      for (int j = 0; j < BH; ++j) h1[j] = nq_tanh(fmaf(theta[o.W1 + j], d, fmaf(theta[o.W1 + BH + j], e, theta[o.b1 + j])));
#pragma unroll 4
      for (int k = 0; k < BH; ++k) {
        float u = theta[o.b2 + k];
#pragma unroll
        for (int j = 0; j < BH; ++j) u = fmaf(theta[o.W2 + j * BH + k], h1[j], u);
        h2[k] = nq_tanh(u);
        d2[k] = seed * theta[o.W3 + k] * fmaf(-h2[k], h2[k], 1.f);
      }
      sh_f[tid][0] = d; sh_f[tid][1] = e; sh_seed[tid] = seed;
#pragma unroll
      for (int j = 0; j < BH; ++j) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < BH; ++k) a = fmaf(theta[o.W2 + j * BH + k], d2[k], a);
        sh_h1[tid][j] = h1[j];
        sh_d1[tid][j] = a * fmaf(-h1[j], h1[j], 1.f);
        sh_d2[tid][j] = d2[j];
        sh_sh2[tid][j] = seed * h2[j];
      }
    }
    __syncthreads();
    // ---- phase 2 -------------------------------------------------------------------------
#pragma unroll 2
    for (int rw = grp; rw < BR; rw += BT / 16) {
      const float4 hv = *reinterpret_cast<const float4*>(&sh_h1[rw][jb]);
      const float4 dv = *reinterpret_cast<const float4*>(&sh_d2[rw][kb]);
      const float h4[4] = {hv.x, hv.y, hv.z, hv.w}, d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) accW2[a][c] = fmaf(h4[a], d4[c], accW2[a][c]);
      const float d1v = sh_d1[rw][b16];
      a_b1 += d1v;
      a_w10 = fmaf(sh_f[rw][0], d1v, a_w10);
      a_w11 = fmaf(sh_f[rw][1], d1v, a_w11);
      a_b2 += sh_d2[rw][b16];
      a_w3 += sh_sh2[rw][b16];
      if (b16 == 0) a_b3 += sh_seed[rw];
    }
    __syncthreads();
  }
  // sum the 16 row groups; slab in theta-relative order of the backflow block: b1[16] W1[2][16] b2[16] W2[16][16] b3 W3[16]
  const int o_b1 = 0, o_W1 = BH, o_b2 = 3 * BH, o_W2 = 4 * BH, o_b3 = 4 * BH + BH * BH, o_W3 = o_b3 + 1;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) atomicAdd(&sh_red[o_W2 + (jb + a) * BH + kb + c], accW2[a][c]);
  atomicAdd(&sh_red[o_b1 + b16], a_b1);
  atomicAdd(&sh_red[o_W1 + b16], a_w10);
  atomicAdd(&sh_red[o_W1 + BH + b16], a_w11);
  atomicAdd(&sh_red[o_b2 + b16], a_b2);
  atomicAdd(&sh_red[o_W3 + b16], a_w3);
  if (b16 == 0) atomicAdd(&sh_red[o_b3], a_b3);
  __syncthreads();
  float* slab = part + (int64_t)blockIdx.x * BF_P;
  for (int q = tid; q < BF_P; q += BT) slab[q] = sh_red[q];
}

__global__ void bf_grads_finish_kernel(const DevSys s, const float* __restrict__ part, const int nblk,
                                       double* __restrict__ out) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= BF_P) return;
  double a = 0;
  for (int b = lane; b < nblk; b += 32) a += (double)part[(int64_t)b * BF_P + q];
  a = warp_sum(a);
  if (lane != 0) return;
  const MlpOff o = s.bf;
  int dst;
  if (q < BH) dst = o.b1 + q;
  else if (q < 3 * BH) dst = o.W1 + (q - BH);
  else if (q < 4 * BH) dst = o.b2 + (q - 3 * BH);
  else if (q < 4 * BH + BH * BH) dst = o.W2 + (q - 4 * BH);
  else if (q == 4 * BH + BH * BH) dst = o.b3;
  else dst = o.W3 + (q - (4 * BH + BH * BH + 1));
  out[dst] = a;
}

}  // namespace

int nq_launch_backflow_xg(nqmc_ctx* ctx, const float* r, int64_t W, float* x, float* G, cudaStream_t st) {
  backflow_xg_kernel<<<(unsigned)((W + WT - 1) / WT), WT, 0, st>>>(ctx->sys, ctx->theta, r, W, x, G);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_launch_backflow_x(nqmc_ctx* ctx, const float* r, int64_t W, float* x, cudaStream_t st) {
  backflow_x_kernel<<<(unsigned)((W + WT - 1) / WT), WT, 0, st>>>(ctx->sys, ctx->theta, r, W, x);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_launch_assemble_bf(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, int gw, cudaStream_t st) {
  ProfScope ps(ctx, NQ_PROF_ASSEMBLE, st);
  assemble_bf_kernel<<<(unsigned)((W + WT - 1) / WT), WT, 0, st>>>(ctx->sys, ctx->theta, r, ctx->phi, ctx->inv, ctx->vvec, W,
                                                                    e_l, parts, gw);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_launch_bf_grads(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr_center, const float* e_l,
                       int64_t W, double* out, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const int64_t n_rows = W * (s.N * (s.N - 1) / 2);
  if (n_rows == 0) return NQMC_OK;
  int64_t nblk = (n_rows + BR - 1) / BR;
  if (nblk > 2 * ctx->sm_count) nblk = 2 * ctx->sm_count;
  const int64_t cap = ((int64_t)ctx->n_cta_bwd * ctx->p_mlp) / BF_P;
  if (nblk > cap) nblk = cap;
  bf_grads_kernel<<<(unsigned)nblk, BT, 0, st>>>(s, ctx->theta, r, ctx->vvec, wgt, hdr_center, e_l, W, ctx->part_mlp);
  NQMC_LAUNCH_CHECK();
  bf_grads_finish_kernel<<<(BF_P * 32 + 127) / 128, 128, 0, st>>>(s, ctx->part_mlp, (int)nblk, out);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
