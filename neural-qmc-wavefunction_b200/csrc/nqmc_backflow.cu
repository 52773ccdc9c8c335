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
// enters E_L, and G_i depends on the backflow Jacobian alone, so backflow_elec_kernel computes G_i BEFORE the orbital
// pass and the orbital kernels propagate the G-weighted Laplacian L_G = sum_ab G_ab d_a d_b forward instead -- it obeys
// the same layer rules as the ordinary Laplacian (linear: L_G(W^T h) = W^T L_G(h); tanh: t' L_G(u) + t'' grad u^T G grad u)
// -- i.e. 5 channels (value, gradient, L_G) instead of 10: half the flops, and the 5-channel tensor-core kernels apply.
// The 10-channel route stays for the FP32 FFMA parity anchor (NQMC_TC=0).
// Work is spread over (pair, walker), (electron, walker) and (spin, walker) threads: at 32,768 walkers per GPU one thread
// per walker leaves 7 warps per SM running long serial chains (round 2a: 0.49 ms of the 2.26 ms H6 step).
#include "nqmc_internal.cuh"

namespace {

constexpr int WT = 128;
constexpr int BH = NQMC_BF_H;   // 16

constexpr int ET = 256;          // threads per CTA of the (pair, walker) / (electron, walker) kernels

// ---------------------------------------------------------------------------------------------
// eta table.  Round 1/2a evaluated the 2-16-16-1 eta network inside the per-walker kernels, pair after pair: with 32,768
// walkers per GPU that is 7 warps per SM each running a ~10^5-instruction serial chain (assemble_bf 295 us, x 45 us,
// xg 54 us at H6).  Now one thread per (pair, walker) row evaluates eta (and eta', eta'' w.r.t. the distance when LEVEL = 2)
// once, weights broadcast from shared memory as LDS.128, and the consumers read tab[level][pair][walker].
// LEVEL 0: eta ; 2: eta, eta', eta''
template <int LEVEL>
__global__ void __launch_bounds__(ET) eta_table_kernel(const DevSys s, const float* __restrict__ theta,
                                                       const float* __restrict__ r, const int64_t W,
                                                       float* __restrict__ tab) {
  __shared__ __align__(16) float sW2T[BH][BH];            // [k][j] = W2[j][k]
  __shared__ float sW1[2 * BH], sb1[BH], sb2[BH], sW3[BH], sb3;
  const MlpOff o = s.bf;
  const int tid = threadIdx.x;
  {
    const int j = tid >> 4, k = tid & 15;                 // ET == BH * BH
    sW2T[k][j] = theta[o.W2 + j * BH + k];
    if (tid < 2 * BH) sW1[tid] = theta[o.W1 + tid];
    if (tid < BH) { sb1[tid] = theta[o.b1 + tid]; sb2[tid] = theta[o.b2 + tid]; sW3[tid] = theta[o.W3 + tid]; }
    if (tid == 0) sb3 = theta[o.b3];
  }
  __syncthreads();
  const int N = s.N, NP = N * (N - 1) / 2;
  const int64_t n_rows = W * NP, row = (int64_t)blockIdx.x * ET + tid;
  if (row >= n_rows) return;
  // rows are ordered pair-major so that consecutive threads read consecutive walkers (coalesced)
  const int p = (int)(row / W);
  const int64_t w = row % W;
  int i = 0, acc = 0;
  while (acc + (N - 1 - i) <= p) { acc += N - 1 - i; ++i; }
  const int j = i + 1 + (p - acc);
  const float dx = r[(int64_t)(3 * i) * W + w] - r[(int64_t)(3 * j) * W + w];
  const float dy = r[(int64_t)(3 * i + 1) * W + w] - r[(int64_t)(3 * j + 1) * W + w];
  const float dz = r[(int64_t)(3 * i + 2) * W + w] - r[(int64_t)(3 * j + 2) * W + w];
  const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
  const float e = __expf(-d);
  float h[BH], hp[BH], hpp[BH];
#pragma unroll
  for (int q = 0; q < BH; ++q) {
    const float w0 = sW1[q], w1 = sW1[BH + q];
    const float t = nq_tanh(fmaf(w0, d, fmaf(w1, e, sb1[q])));
    h[q] = t;
    if (LEVEL == 2) {
      const float tp = fmaf(-t, t, 1.f), up = w0 - w1 * e, upp = w1 * e;
      hp[q] = tp * up;
      hpp[q] = fmaf(tp, upp, -2.f * t * tp * up * up);
    }
  }
  float m = sb3, mp = 0.f, mpp = 0.f;
#pragma unroll 4
  for (int k = 0; k < BH; ++k) {
    float u = sb2[k], up = 0.f, upp = 0.f;
#pragma unroll
    for (int q = 0; q < BH; q += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(&sW2T[k][q]);
      const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        u = fmaf(w4[c], h[q + c], u);
        if (LEVEL == 2) { up = fmaf(w4[c], hp[q + c], up); upp = fmaf(w4[c], hpp[q + c], upp); }
      }
    }
    const float t = nq_tanh(u), w3 = sW3[k];
    m = fmaf(w3, t, m);
    if (LEVEL == 2) {
      const float tp = fmaf(-t, t, 1.f);
      mp = fmaf(w3, tp * up, mp);
      mpp = fmaf(w3, fmaf(tp, upp, -2.f * t * tp * up * up), mpp);
    }
  }
  const float E = __expf(-0.5f * d);
  tab[row] = m * E;
  if (LEVEL == 2) {
    tab[n_rows + row] = (mp - 0.5f * m) * E;
    tab[2 * n_rows + row] = (mpp - mp + 0.25f * m) * E;
  }
}

__device__ __forceinline__ int pair_index(int i, int j, int N) {   // i < j
  return i * N - (i * (i + 1)) / 2 + (j - i - 1);
}

// One thread per (electron, walker):  x_e = r_e + sum_{j != e} eta_ej (r_e - r_j), and with FULL also what the local
// energy needs from the backflow Jacobian (u = unit vector of r_e - r_j, kappa = eta' d):
//   A_e  = d x_e / d r_e = S_e I + T_e,  S_e = 1 + sum_j eta_ej,  T_e = sum_j kappa_ej u u^T   (symmetric: xx xy xz yy yz zz)
//   d x_e / d r_j = -(eta_ej I + kappa_ej u u^T)
//   G_e  = sum_mu d_mu x_e d_mu x_e^T = A_e^2 + sum_j (eta^2 I + (2 eta kappa + kappa^2) u u^T)
//   Lap x_e = sum_j 2 (d eta'' + 4 eta') u
template <bool FULL>
__global__ void __launch_bounds__(ET) backflow_elec_kernel(const DevSys s, const float* __restrict__ r,
                                                           const float* __restrict__ tab, const int64_t W,
                                                           float* __restrict__ x, float* __restrict__ G,
                                                           float* __restrict__ lapx, float* __restrict__ Am) {
  const int N = s.N;
  const int64_t row = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (row >= (int64_t)N * W) return;
  const int e = (int)(row / W);
  const int64_t w = row % W, ls = (int64_t)(N * (N - 1) / 2) * W;
  const float rx = r[(int64_t)(3 * e) * W + w], ry = r[(int64_t)(3 * e + 1) * W + w], rz = r[(int64_t)(3 * e + 2) * W + w];
  float ax = rx, ay = ry, az = rz, S = 1.f, T[6], Gd[6], lx[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 6; ++c) { T[c] = 0.f; Gd[c] = 0.f; }
  for (int j = 0; j < N; ++j) {
    if (j == e) continue;
    const int64_t pw = (int64_t)(e < j ? pair_index(e, j, N) : pair_index(j, e, N)) * W + w;
    const float dx = rx - r[(int64_t)(3 * j) * W + w], dy = ry - r[(int64_t)(3 * j + 1) * W + w],
                dz = rz - r[(int64_t)(3 * j + 2) * W + w];
    const float eta = tab[pw];
    ax = fmaf(eta, dx, ax); ay = fmaf(eta, dy, ay); az = fmaf(eta, dz, az);
    if (FULL) {
      const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      const float etap = tab[ls + pw], etapp = tab[2 * ls + pw];
      const float id = 1.0f / d, u[3] = {dx * id, dy * id, dz * id};
      const float kap = etap * d, lf = 2.0f * fmaf(d, etapp, 4.0f * etap);
      S += eta;
      const float off = fmaf(2.0f * eta, kap, kap * kap);
      int c = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) {
          const float uu = u[a] * u[b];
          T[c] = fmaf(kap, uu, T[c]);
          Gd[c] += fmaf(off, uu, (a == b) ? eta * eta : 0.f);
          ++c;
        }
#pragma unroll
      for (int a = 0; a < 3; ++a) lx[a] = fmaf(lf, u[a], lx[a]);
    }
  }
  x[(int64_t)(3 * e) * W + w] = ax;
  x[(int64_t)(3 * e + 1) * W + w] = ay;
  x[(int64_t)(3 * e + 2) * W + w] = az;
  if (FULL) {
    const float A[6] = {S + T[0], T[1], T[2], S + T[3], T[4], S + T[5]};
    float* g = G + (int64_t)(6 * e) * W + w;
    float* am = Am + (int64_t)(6 * e) * W + w;
#pragma unroll
    for (int c = 0; c < 6; ++c) am[c * W] = A[c];
    g[0] = Gd[0] + A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
    g[W] = Gd[1] + A[0] * A[1] + A[1] * A[3] + A[2] * A[4];
    g[2 * W] = Gd[2] + A[0] * A[2] + A[1] * A[4] + A[2] * A[5];
    g[3 * W] = Gd[3] + A[1] * A[1] + A[3] * A[3] + A[4] * A[4];
    g[4 * W] = Gd[4] + A[1] * A[2] + A[3] * A[4] + A[4] * A[5];
    g[5 * W] = Gd[5] + A[2] * A[2] + A[4] * A[4] + A[5] * A[5];
#pragma unroll
    for (int a = 0; a < 3; ++a) lapx[(int64_t)(3 * e + a) * W + w] = lx[a];
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

// ---------------------------------------------------------------------------------------------
// E_L with backflow, in three launches (round 2a: one thread per walker did all of it, serially).
// phi holds, per orbital-matrix entry evaluated at x(r), either 10 channels (value, gradient, Hessian; gw == 0) or
// 5 (value, gradient, G-weighted Laplacian; gw == 1).
//
// (1) bf_slater_kernel, one thread per (spin, walker): inverse and slogdet of Phi_s, the seeds of the reverse pass, the
//     Laplacian trace  term1_s = sum_ik inv[k,i] (sum_ab Hess_ik[ab] G_i[ab] + g_ik . Lap x_i),  and
//         Q[i][i'] = sum_k inv[k,i'] g_ik   (a 3-vector per ordered pair of same-spin electrons;  Q[i][i] = v_i).
// (2) bf_elec_energy_kernel, one thread per (electron e, walker): with J_i = d x_i / d r_e (symmetric 3 x 3) and
//     P[i][i'] = J_i Q[i][i'],
//         grad_e log|det| = sum_i P[i][i],     sum_{c in xyz} tr((Phi^-1 d_(e,c) Phi)^2) = sum_{s} sum_{i,i' in s} P[i][i'] . P[i'][i]
//     (D_c[i][k] = g_ik . J_i[:,c], so (Phi^-1 D_c)[k][l] = sum_i inv[k,i] D_c[i][l] and the trace of its square regroups into
//     dot products of the P vectors; the 3N x n_s^3 products of round 2a become N x n_s^2 3-vector products), plus this
//     electron's share of the Jastrow / cusp gradient, Laplacian and of the potentials on the physical r.
// (3) bf_combine_kernel, one thread per walker: the sums over spins and electrons.
__global__ void __launch_bounds__(WT) bf_slater_kernel(const DevSys s, const float* __restrict__ phi,
                                                       const float* __restrict__ G, const float* __restrict__ lapx,
                                                       const int64_t W, float* __restrict__ inv_out,
                                                       float* __restrict__ v_out, float* __restrict__ qbf,
                                                       float* __restrict__ spart, const int gw) {
  const int64_t row = (int64_t)blockIdx.x * WT + threadIdx.x;
  if (row >= 2 * W) return;
  const int spin = (int)(row / W);
  const int64_t w = row % W;
  const int n = spin == 0 ? s.n_up : s.n_dn;
  float* sp = spart + (int64_t)(3 * spin) * W + w;            // term1, log|det|, sign
  if (n == 0) { sp[0] = 0.f; sp[W] = 0.f; sp[2 * W] = 1.f; return; }
  const int e0 = spin == 0 ? 0 : s.n_up;
  const int base = spin == 0 ? 0 : s.n_up * s.n_up;
  const int64_t cs = (int64_t)s.n_ent * W;
  float a[NQMC_MAX_NS * NQMC_MAX_NS], inv[NQMC_MAX_NS * NQMC_MAX_NS];   // inv[k*n + i]
  float gt[NQMC_MAX_NS * NQMC_MAX_NS][3];
  for (int q = 0; q < n * n; ++q) a[q] = phi[(int64_t)(base + q) * W + w];
  float sg, la, term1 = 0.f;
  inv_slogdet_dyn(n, a, inv, sg, la);
  for (int i = 0; i < n; ++i) {
    const int e = e0 + i;
    float Ge[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (!gw)
#pragma unroll
      for (int c = 0; c < 6; ++c) Ge[c] = G[(int64_t)(6 * e + c) * W + w];
    const float l0 = lapx[(int64_t)(3 * e) * W + w], l1 = lapx[(int64_t)(3 * e + 1) * W + w],
                l2 = lapx[(int64_t)(3 * e + 2) * W + w];
    for (int k = 0; k < n; ++k) {
      const int64_t eidx = (int64_t)(base + i * n + k) * W + w;
      const float c = inv[k * n + i];
      inv_out[eidx] = c;
      const float g0 = phi[eidx + cs], g1 = phi[eidx + 2 * cs], g2 = phi[eidx + 3 * cs];
      gt[i * n + k][0] = g0; gt[i * n + k][1] = g1; gt[i * n + k][2] = g2;
      // sum_ab Hess[ab] G[ab] (off-diagonals count twice) + g . Lap x
      float hh;
      if (gw) hh = phi[eidx + 4 * cs];                            // the orbital pass already contracted Hess with G
      else hh = phi[eidx + 4 * cs] * Ge[0] + phi[eidx + 7 * cs] * Ge[3] + phi[eidx + 9 * cs] * Ge[5] +
                2.0f * (phi[eidx + 5 * cs] * Ge[1] + phi[eidx + 6 * cs] * Ge[2] + phi[eidx + 8 * cs] * Ge[4]);
      hh += g0 * l0 + g1 * l1 + g2 * l2;
      term1 = fmaf(c, hh, term1);
    }
    for (int i2 = 0; i2 < n; ++i2) {
      float q0 = 0.f, q1 = 0.f, q2 = 0.f;
      for (int k = 0; k < n; ++k) {
        const float c = inv[k * n + i2];
        q0 = fmaf(c, gt[i * n + k][0], q0); q1 = fmaf(c, gt[i * n + k][1], q1); q2 = fmaf(c, gt[i * n + k][2], q2);
      }
      const int64_t qidx = (int64_t)(base + i * n + i2) * W + w;
      qbf[qidx] = q0; qbf[qidx + cs] = q1; qbf[qidx + 2 * cs] = q2;
      if (i2 == i) {
        v_out[(int64_t)(3 * e) * W + w] = q0; v_out[(int64_t)(3 * e + 1) * W + w] = q1; v_out[(int64_t)(3 * e + 2) * W + w] = q2;
      }
    }
  }
  sp[0] = term1; sp[W] = la; sp[2 * W] = sg;
}

constexpr int EP = 6;   // per-electron partials: |grad_e log|psi||^2, trace term, J, Laplacian of J, V_en, V_ee

__global__ void __launch_bounds__(WT) bf_elec_energy_kernel(const DevSys s, const float* __restrict__ theta,
                                                            const float* __restrict__ r, const float* __restrict__ tab,
                                                            const float* __restrict__ qbf, const float* __restrict__ Am,
                                                            const int64_t W, float* __restrict__ epart) {
  const int N = s.N;
  const int64_t row = (int64_t)blockIdx.x * WT + threadIdx.x;
  if (row >= (int64_t)N * W) return;
  const int e = (int)(row / W);
  const int64_t w = row % W, ls = (int64_t)(N * (N - 1) / 2) * W, cs = (int64_t)s.n_ent * W;
  const float rx = r[(int64_t)(3 * e) * W + w], ry = r[(int64_t)(3 * e + 1) * W + w], rz = r[(int64_t)(3 * e + 2) * W + w];
  // ---- determinant part ---------------------------------------------------------------------
  float gx = 0.f, gy = 0.f, gz = 0.f, term2 = 0.f;
  float P[NQMC_MAX_NS * NQMC_MAX_NS][3];
  for (int spin = 0; spin < 2; ++spin) {
    const int n = spin == 0 ? s.n_up : s.n_dn;
    const int e0 = spin == 0 ? 0 : s.n_up;
    const int base = spin == 0 ? 0 : s.n_up * s.n_up;
    for (int i = 0; i < n; ++i) {
      const int ei = e0 + i;
      float Jm[6];                                               // d x_ei / d r_e, packed xx xy xz yy yz zz
      if (ei == e) {
#pragma unroll
        for (int c = 0; c < 6; ++c) Jm[c] = Am[(int64_t)(6 * e + c) * W + w];
      } else {
        const int64_t pw = (int64_t)(ei < e ? pair_index(ei, e, N) : pair_index(e, ei, N)) * W + w;
        const float dx = r[(int64_t)(3 * ei) * W + w] - rx, dy = r[(int64_t)(3 * ei + 1) * W + w] - ry,
                    dz = r[(int64_t)(3 * ei + 2) * W + w] - rz;
        const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        const float eta = tab[pw], kap = tab[ls + pw] * d, id = 1.0f / d;
        const float u0 = dx * id, u1 = dy * id, u2 = dz * id, k0 = -kap * u0, k1 = -kap * u1, k2 = -kap * u2;
        Jm[0] = fmaf(k0, u0, -eta); Jm[1] = k0 * u1; Jm[2] = k0 * u2;
        Jm[3] = fmaf(k1, u1, -eta); Jm[4] = k1 * u2; Jm[5] = fmaf(k2, u2, -eta);
      }
      for (int i2 = 0; i2 < n; ++i2) {
        const int64_t qidx = (int64_t)(base + i * n + i2) * W + w;
        const float q0 = qbf[qidx], q1 = qbf[qidx + cs], q2 = qbf[qidx + 2 * cs];
        const float p0 = Jm[0] * q0 + Jm[1] * q1 + Jm[2] * q2, p1 = Jm[1] * q0 + Jm[3] * q1 + Jm[4] * q2,
                    p2 = Jm[2] * q0 + Jm[4] * q1 + Jm[5] * q2;
        P[i * n + i2][0] = p0; P[i * n + i2][1] = p1; P[i * n + i2][2] = p2;
        if (i2 == i) { gx += p0; gy += p1; gz += p2; }
      }
    }
    for (int i = 0; i < n; ++i) {
      const float* d = P[i * n + i];
      term2 += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      for (int i2 = i + 1; i2 < n; ++i2) {
        const float *pa = P[i * n + i2], *pb = P[i2 * n + i];
        term2 = fmaf(2.0f, pa[0] * pb[0] + pa[1] * pb[1] + pa[2] * pb[2], term2);
      }
    }
  }
  // ---- Jastrow + cusps on the physical r (formulas of nqmc_walker.cu; each e-e pair is seen from both ends) -------
  const float a_ee = theta[s.jas_a_ee], b_ee = theta[s.jas_b_ee];
  const float a_en = theta[s.jas_a_en], b_en = theta[s.jas_b_en];
  const float cb = s.use_cusp ? fabsf(theta[s.ee_cusp_b]) : 0.f;
  float J = 0.f, lap = 0.f, ven = 0.f, vee = 0.f;
  for (int j = 0; j < N; ++j) {
    if (j == e) continue;
    const float dx = rx - r[(int64_t)(3 * j) * W + w], dy = ry - r[(int64_t)(3 * j + 1) * W + w],
                dz = rz - r[(int64_t)(3 * j + 2) * W + w];
    const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    float q = 1.0f / fmaf(b_ee, d, 1.0f);
    float u = a_ee * d * q, up = a_ee * q * q, upp = -2.0f * a_ee * b_ee * q * q * q;
    if (s.use_cusp) {
      const float c = ((e < s.n_up) == (j < s.n_up)) ? s.cusp_same : s.cusp_diff;
      q = 1.0f / fmaf(cb, d, 1.0f);
      u += c * d * q; up += c * q * q; upp += -2.0f * c * cb * q * q * q;
    }
    J = fmaf(0.5f, u, J);
    const float id = 1.0f / d, c = up * id;
    gx = fmaf(c, dx, gx); gy = fmaf(c, dy, gy); gz = fmaf(c, dz, gz);
    lap += upp + 2.0f * c;
    vee = fmaf(0.5f, id, vee);
  }
  for (int a = 0; a < s.A; ++a) {
    const float dx = rx - s.R[3 * a], dy = ry - s.R[3 * a + 1], dz = rz - s.R[3 * a + 2];
    const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    float q = 1.0f / fmaf(b_en, d, 1.0f);
    float u = a_en * d * q, up = a_en * q * q, upp = -2.0f * a_en * b_en * q * q * q;
    if (s.use_cusp) {
      const float ex = __expf(-d);
      const float* W0 = theta + s.en_cusp_W0[a];
      const float* b0 = theta + s.en_cusp_b0[a];
      const float* W1 = theta + s.en_cusp_W1[a];
      float g = theta[s.en_cusp_b1[a]], gp = 0.f, gpp = 0.f;
#pragma unroll 4
      for (int h = 0; h < NQMC_CUSP_H; ++h) {
        const float w0 = W0[h], w1 = W0[NQMC_CUSP_H + h], w2 = W0[2 * NQMC_CUSP_H + h];
        const float t = nq_tanh(fmaf(w0, d, fmaf(w1, d * d, fmaf(w2, ex, b0[h]))));
        const float wo = W1[h], tp = fmaf(-t, t, 1.0f);
        const float zp = fmaf(2.0f * w1, d, w0) - w2 * ex, zpp = fmaf(w2, ex, 2.0f * w1);
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
    gx = fmaf(c, dx, gx); gy = fmaf(c, dy, gy); gz = fmaf(c, dz, gz);
    lap += upp + 2.0f * c;
    ven -= s.Z[a] / fmaxf(d, 1e-10f);
  }
  float* ep = epart + (int64_t)e * W + w;
  const int64_t es = (int64_t)N * W;
  ep[0] = gx * gx + gy * gy + gz * gz;
  ep[es] = term2; ep[2 * es] = J; ep[3 * es] = lap; ep[4 * es] = ven; ep[5 * es] = vee;
}

__global__ void __launch_bounds__(ET) bf_combine_kernel(const DevSys s, const float* __restrict__ spart,
                                                        const float* __restrict__ epart, const int64_t W,
                                                        float* __restrict__ e_l, float* __restrict__ parts) {
  const int64_t w = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (w >= W) return;
  const int N = s.N;
  const int64_t es = (int64_t)N * W;
  const float term1 = spart[w] + spart[3 * W + w];
  float logabs = spart[W + w] + spart[4 * W + w];
  float g2 = 0.f, term2 = 0.f, J = 0.f, lapJ = 0.f, ven = 0.f, vee = 0.f;
  for (int e = 0; e < N; ++e) {
    const float* ep = epart + (int64_t)e * W + w;
    g2 += ep[0]; term2 += ep[es]; J += ep[2 * es]; lapJ += ep[3 * es]; ven += ep[4 * es]; vee += ep[5 * es];
  }
  logabs += J;
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

namespace {
// regions of ctx->bfw (nqmc_reserve sizes it with nq_backflow_ws_floats): table, Lap x, A, Q, per-spin and per-electron partials
struct BfWs { float *tab, *lapx, *am, *qbf, *spart, *epart; };
BfWs bf_ws(const nqmc_ctx* ctx) {
  const DevSys& s = ctx->sys;
  const size_t c = (size_t)ctx->cap_w, N = (size_t)s.N, NP = N * (N - 1) / 2;
  BfWs ws;
  ws.tab = ctx->bfw;
  ws.lapx = ws.tab + 3 * NP * c;
  ws.am = ws.lapx + 3 * N * c;
  ws.qbf = ws.am + 6 * N * c;
  ws.spart = ws.qbf + 3 * (size_t)s.n_ent * c;
  ws.epart = ws.spart + 6 * c;
  return ws;
}
unsigned blocks(int64_t rows, int per) { return (unsigned)((rows + per - 1) / per); }
}  // namespace

size_t nq_backflow_ws_floats(const DevSys& s) {
  const size_t N = (size_t)s.N, NP = N * (N - 1) / 2;
  return 3 * NP + 3 * N + 6 * N + 3 * (size_t)s.n_ent + 6 + EP * N;
}

int nq_launch_backflow_xg(nqmc_ctx* ctx, const float* r, int64_t W, float* x, float* G, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const BfWs ws = bf_ws(ctx);
  const int64_t NP = s.N * (s.N - 1) / 2;
  if (NP > 0) {
    eta_table_kernel<2><<<blocks(W * NP, ET), ET, 0, st>>>(s, ctx->theta, r, W, ws.tab);
    NQMC_LAUNCH_CHECK();
  }
  backflow_elec_kernel<true><<<blocks(W * s.N, ET), ET, 0, st>>>(s, r, ws.tab, W, x, G, ws.lapx, ws.am);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_launch_backflow_x(nqmc_ctx* ctx, const float* r, int64_t W, float* x, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const BfWs ws = bf_ws(ctx);
  const int64_t NP = s.N * (s.N - 1) / 2;
  if (NP > 0) {
    eta_table_kernel<0><<<blocks(W * NP, ET), ET, 0, st>>>(s, ctx->theta, r, W, ws.tab);
    NQMC_LAUNCH_CHECK();
  }
  backflow_elec_kernel<false><<<blocks(W * s.N, ET), ET, 0, st>>>(s, r, ws.tab, W, x, nullptr, nullptr, nullptr);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

// needs nq_launch_backflow_xg(r) (table, Lap x, A; G when gw == 0) and the orbital pass at x(r) before it
int nq_launch_assemble_bf(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, int gw, cudaStream_t st) {
  ProfScope ps(ctx, NQ_PROF_ASSEMBLE, st);
  const DevSys& s = ctx->sys;
  const BfWs ws = bf_ws(ctx);
  bf_slater_kernel<<<blocks(2 * W, WT), WT, 0, st>>>(s, ctx->phi, ctx->gbf, ws.lapx, W, ctx->inv, ctx->vvec, ws.qbf, ws.spart, gw);
  NQMC_LAUNCH_CHECK();
  bf_elec_energy_kernel<<<blocks(W * s.N, WT), WT, 0, st>>>(s, ctx->theta, r, ws.tab, ws.qbf, ws.am, W, ws.epart);
  NQMC_LAUNCH_CHECK();
  bf_combine_kernel<<<blocks(W, ET), ET, 0, st>>>(s, ws.spart, ws.epart, W, e_l, parts);
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
