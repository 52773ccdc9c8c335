// nqmc_walker.cu -- per-walker kernels: proposal, Slater determinants / inverses, Jastrow (+cusp),
// local-energy assembly, Metropolis accept/reject, energy statistics, layout transposes.
// One thread owns one walker; all per-walker state (N <= 16 electrons, n_s x n_s orbital matrices
// with n_s <= 8) lives in registers; global traffic is SoA so every load/store is coalesced.
//
// Replaces:
//   log_slater_determinant / slater_with_spin   .github/phase2-issue-body.md:38-66
//   JastrowFactor                                .github/phase2-issue-body.md:125-152
//   ElectronNuclearCusp / ElectronElectronCusp   .github/phase3-issue-body.md:50-122
//   kinetic_energy (trace of Hessian + |grad|^2) src/nqmc/hamiltonian/molecular.py:52-69
//   electron_nuclear_potential / electron_electron_potential / local_energy  molecular.py:87-175
//   MetropolisSampler.step                       src/nqmc/sampler/metropolis.py:115-147
//   jnp.mean / jnp.std of the local energies     src/nqmc/optimizer/vmc.py:113-114
#include "nqmc_internal.cuh"

namespace {

constexpr int WT = 128;           // threads per block for per-walker kernels
constexpr int NPART = 16;         // doubles per block partial

// ---------------------------------------------------------------------------------------------
// Gauss-Jordan inverse with partial pivoting, fully unrolled (static register indexing).
// a[i][k] is destroyed; inv = a^-1 (inv[k][i]); returns sign and log|det| like slogdet.
template <int NS>
__device__ __forceinline__ void inv_slogdet(const int n, float (&a)[NS][NS], float (&inv)[NS][NS], float& sign,
                                            float& logabs) {
#pragma unroll
  for (int i = 0; i < NS; ++i)
#pragma unroll
    for (int j = 0; j < NS; ++j) inv[i][j] = (i == j) ? 1.f : 0.f;
  sign = 1.f;
  logabs = 0.f;
#pragma unroll
  for (int c = 0; c < NS; ++c) {
    if (c < n) {
      int p = c;
      float best = fabsf(a[c][c]);
#pragma unroll
      for (int rw = c + 1; rw < NS; ++rw)
        if (rw < n) {
          float v = fabsf(a[rw][c]);
          if (v > best) { best = v; p = rw; }
        }
#pragma unroll
      for (int rw = c + 1; rw < NS; ++rw)
        if (rw == p) {
#pragma unroll
          for (int j = 0; j < NS; ++j) {
            float t = a[c][j]; a[c][j] = a[rw][j]; a[rw][j] = t;
            t = inv[c][j]; inv[c][j] = inv[rw][j]; inv[rw][j] = t;
          }
          sign = -sign;
        }
      const float piv = a[c][c];
      sign = piv < 0.f ? -sign : sign;
      logabs += logf(fabsf(piv));
      const float ip = 1.0f / piv;
#pragma unroll
      for (int j = 0; j < NS; ++j) { a[c][j] *= ip; inv[c][j] *= ip; }
#pragma unroll
      for (int rw = 0; rw < NS; ++rw)
        if (rw != c && rw < n) {
          const float f = a[rw][c];
#pragma unroll
          for (int j = 0; j < NS; ++j) {
            a[rw][j] = fmaf(-f, a[c][j], a[rw][j]);
            inv[rw][j] = fmaf(-f, inv[c][j], inv[rw][j]);
          }
        }
    }
  }
}

// Pade pair term u = a r / (1 + b r) and derivatives
__device__ __forceinline__ void pade(float r, float a, float b, float& u, float& up, float& upp) {
  float q = __fdividef(1.0f, fmaf(b, r, 1.0f));      // 2-ulp reciprocal: the IEEE division sequence is ~8 instructions per pair
  u = a * r * q;
  up = a * q * q;
  upp = -2.0f * a * b * q * q * q;
}

// Jastrow + cusps on physical coordinates.  Accumulates value J, gradient g[e][3], Laplacian lap,
// and the bare potentials.  LAP=false: value only.
// jpar != nullptr: also d J / d(a_ee, a_en, b_ee, b_en) of the Pade terms (u = a d q, q = 1 / (1 + b d):
// du/da = d q, du/db = -a d^2 q^2), taken in the same pair loop.
template <int NE, bool LAP>
__device__ __forceinline__ void jastrow_cusp_terms(const DevSys& s, const float* __restrict__ theta, const float (&x)[NE],
                                                   const float (&y)[NE], const float (&z)[NE], float& J,
                                                   float (&gx)[NE], float (&gy)[NE], float (&gz)[NE], float& lap,
                                                   float& ven, float& vee, float* jpar = nullptr) {
  float g_aee = 0.f, g_aen = 0.f, g_bee = 0.f, g_ben = 0.f;
  const float a_ee = theta[s.jas_a_ee], b_ee = theta[s.jas_b_ee];
  const float a_en = theta[s.jas_a_en], b_en = theta[s.jas_b_en];
  const float cb = s.use_cusp ? fabsf(theta[s.ee_cusp_b]) : 0.f;
  J = 0.f; lap = 0.f; ven = 0.f; vee = 0.f;
#pragma unroll
  for (int i = 0; i < NE; ++i)
    if (i < s.N) {
#pragma unroll
      for (int j = i + 1; j < NE; ++j)
        if (j < s.N) {
          float dx = x[i] - x[j], dy = y[i] - y[j], dz = z[i] - z[j];
          // d and 1/d from one rsqrt (2 ulp) instead of an IEEE sqrt and an IEEE division
          const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
          const float id = rsqrtf(r2);
          float d = r2 > 0.f ? r2 * id : 0.f;
          float u, up, upp;
          pade(d, a_ee, b_ee, u, up, upp);
          if (jpar != nullptr) {
            const float q = __fdividef(1.0f, fmaf(b_ee, d, 1.0f));
            g_aee = fmaf(d, q, g_aee);
            g_bee = fmaf(-a_ee * d * d, q * q, g_bee);
          }
          if (s.use_cusp) {
            const bool same = (i < s.n_up) == (j < s.n_up);
            float u2, up2, upp2;
            pade(d, same ? s.cusp_same : s.cusp_diff, cb, u2, up2, upp2);
            u += u2; up += up2; upp += upp2;
          }
          J += u;
          if (LAP) {
            float c = up * id;
            gx[i] = fmaf(c, dx, gx[i]); gy[i] = fmaf(c, dy, gy[i]); gz[i] = fmaf(c, dz, gz[i]);
            gx[j] = fmaf(-c, dx, gx[j]); gy[j] = fmaf(-c, dy, gy[j]); gz[j] = fmaf(-c, dz, gz[j]);
            lap += 2.0f * (upp + 2.0f * c);
            vee += id;
          }
        }
      for (int a = 0; a < s.A; ++a) {
        float dx = x[i] - s.R[3 * a], dy = y[i] - s.R[3 * a + 1], dz = z[i] - s.R[3 * a + 2];
        const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        const float id = rsqrtf(r2);
        float d = r2 > 0.f ? r2 * id : 0.f;
        float u, up, upp;
        pade(d, a_en, b_en, u, up, upp);
        if (jpar != nullptr) {
          const float q = __fdividef(1.0f, fmaf(b_en, d, 1.0f));
          g_aen = fmaf(d, q, g_aen);
          g_ben = fmaf(-a_en * d * d, q * q, g_ben);
        }
        if (s.use_cusp) {
          // -Z d + d g_A(d),  g_A = W1 . tanh(W0^T [d, d^2, e^-d] + b0) + b1  (phase3-issue-body.md:66-81)
          const float e = __expf(-d);
          const float* W0 = theta + s.en_cusp_W0[a];
          const float* b0 = theta + s.en_cusp_b0[a];
          const float* W1 = theta + s.en_cusp_W1[a];
          float g = theta[s.en_cusp_b1[a]], gp = 0.f, gpp = 0.f;
#pragma unroll 4
          for (int h = 0; h < NQMC_CUSP_H; ++h) {
            float w0 = W0[h], w1 = W0[NQMC_CUSP_H + h], w2 = W0[2 * NQMC_CUSP_H + h];
            float zt = fmaf(w0, d, fmaf(w1, d * d, fmaf(w2, e, b0[h])));
            float t = nq_tanh(zt);
            float wo = W1[h];
            g = fmaf(wo, t, g);
            if (LAP) {
              float tp = fmaf(-t, t, 1.0f);
              float zp = fmaf(2.0f * w1, d, w0) - w2 * e;
              float zpp = fmaf(w2, e, 2.0f * w1);
              gp = fmaf(wo, tp * zp, gp);
              gpp = fmaf(wo, tp * zpp - 2.0f * t * tp * zp * zp, gpp);
            }
          }
          u += -s.Z[a] * d + d * g;
          if (LAP) {
            float qp = fmaf(d, gp, g);
            float qpp = fmaf(d, gpp, 2.0f * gp);
            up += -s.Z[a] + qp;
            upp += qpp;
          }
        }
        J += u;
        if (LAP) {
          float c = up * id;
          gx[i] = fmaf(c, dx, gx[i]); gy[i] = fmaf(c, dy, gy[i]); gz[i] = fmaf(c, dz, gz[i]);
          lap += upp + 2.0f * c;
          ven -= s.Z[a] * fminf(id, 1e10f);            // Z / max(d, 1e-10)   (molecular.py:97-100)
        }
      }
    }
  if (jpar != nullptr) {
    jpar[0] = g_aee; jpar[1] = g_aen; jpar[2] = g_bee; jpar[3] = g_ben;
  }
}

// ---------------------------------------------------------------------------------------------
// Slater part + Jastrow for one walker.  ENERGY: also kinetic / potential terms (needs 5-channel phi).
template <int NS, bool ENERGY>
__device__ __forceinline__ void walker_eval(const DevSys& s, const float* __restrict__ theta,
                                            const float* __restrict__ r, const float* __restrict__ phi,
                                            float* __restrict__ inv_out, const int64_t W, const int64_t w,
                                            float& logabs, float& sign, float& T, float& ven, float& vee,
                                            float* jpar = nullptr) {
  constexpr int NE = 2 * NS;
  float x[NE], y[NE], z[NE], gx[NE], gy[NE], gz[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    if (i < s.N) {
      x[i] = r[(int64_t)(3 * i) * W + w];
      y[i] = r[(int64_t)(3 * i + 1) * W + w];
      z[i] = r[(int64_t)(3 * i + 2) * W + w];
    } else {
      x[i] = y[i] = z[i] = 0.f;
    }
    gx[i] = gy[i] = gz[i] = 0.f;
  }
  logabs = 0.f;
  sign = 1.f;
  float lapsum = 0.f, gdet2 = 0.f;
#pragma unroll
  for (int spin = 0; spin < 2; ++spin) {
    const int n = spin == 0 ? s.n_up : s.n_dn;
    if (n == 0) continue;
    const int e0 = spin == 0 ? 0 : s.n_up;
    const int base = spin == 0 ? 0 : s.n_up * s.n_up;
    float a[NS][NS], inv[NS][NS];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int k = 0; k < NS; ++k) a[i][k] = (i < n && k < n) ? phi[(int64_t)(base + i * n + k) * W + w] : (i == k ? 1.f : 0.f);
    float sg, la;
    inv_slogdet<NS>(n, a, inv, sg, la);
    sign *= sg;
    logabs += la;
    if (inv_out != nullptr) {
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int k = 0; k < NS; ++k)
          if (i < n && k < n) inv_out[(int64_t)(base + i * n + k) * W + w] = inv[k][i];
    }
    if (ENERGY) {
#pragma unroll
      for (int i = 0; i < NS; ++i)
        if (i < n) {
          float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
          for (int k = 0; k < NS; ++k)
            if (k < n) {
              const int64_t eidx = (int64_t)(base + i * n + k) * W + w;
              const int64_t cs = (int64_t)s.n_ent * W;
              const float c = inv[k][i];
              sx = fmaf(c, phi[eidx + cs], sx);
              sy = fmaf(c, phi[eidx + 2 * cs], sy);
              sz = fmaf(c, phi[eidx + 3 * cs], sz);
              lapsum = fmaf(c, phi[eidx + 4 * cs], lapsum);
            }
          // gx is indexed statically: scatter through an unrolled match on the electron index
#pragma unroll
          for (int e = 0; e < NE; ++e)
            if (e == e0 + i) { gx[e] = sx; gy[e] = sy; gz[e] = sz; }
          gdet2 += fmaf(sx, sx, fmaf(sy, sy, sz * sz));
        }
    }
  }
  float J, lapJ;
  jastrow_cusp_terms<NE, ENERGY>(s, theta, x, y, z, J, gx, gy, gz, lapJ, ven, vee, jpar);
  logabs += J;
  T = 0.f;
  if (ENERGY) {
    float g2 = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (e < s.N) g2 += fmaf(gx[e], gx[e], fmaf(gy[e], gy[e], gz[e] * gz[e]));
    const float lap_log = lapsum - gdet2 + lapJ;
    T = -0.5f * (lap_log + g2);                        // molecular.py:67
  }
}

template <int NS>
__global__ void __launch_bounds__(WT) assemble_kernel(const DevSys s, const float* __restrict__ theta,
                                                      const float* __restrict__ r, const float* __restrict__ phi,
                                                      float* __restrict__ inv_out, const int64_t W,
                                                      float* __restrict__ e_l, float* __restrict__ parts,
                                                      float* __restrict__ logabs_out, float* __restrict__ sign_out,
                                                      const int want_energy) {
  const int64_t w = (int64_t)blockIdx.x * WT + threadIdx.x;
  if (w >= W) return;
  float la, sg, T, ven, vee;
  if (want_energy) {
    walker_eval<NS, true>(s, theta, r, phi, inv_out, W, w, la, sg, T, ven, vee);
    const float e = T + ven + vee + s.v_nn;            // molecular.py:175
    if (e_l) e_l[w] = e;
    if (parts) {
      parts[w] = T;
      parts[W + w] = ven;
      parts[2 * W + w] = vee;
      parts[3 * W + w] = la;
    }
  } else {
    walker_eval<NS, false>(s, theta, r, phi, inv_out, W, w, la, sg, T, ven, vee);
  }
  if (logabs_out) logabs_out[w] = la;
  if (sign_out) sign_out[w] = sg;
}

// Fused variant for the walker step (no cusp / backflow): E_L assembly + per-block partial sums of the energy
// header AND of the four Jastrow-scalar log-derivatives (sum O, sum E*O), so that neither the statistics pass nor
// the separate per-walker gradient kernel has to re-read the walkers:  sum_w (E_w - Ebar) O_w = sum E O - Ebar sum O.
// part[blk][16] = n, sum E, sum E^2, n_accept, n_bad, sum O[4] (a_ee, a_en, b_ee, b_en), sum E O[4]
template <int NS>
__global__ void __launch_bounds__(WT) assemble_stats_kernel(const DevSys s, const float* __restrict__ theta,
                                                            const float* __restrict__ r, const float* __restrict__ phi,
                                                            float* __restrict__ inv_out, const int64_t W,
                                                            float* __restrict__ e_l, const float* __restrict__ acc_flag,
                                                            double* __restrict__ part) {
  const int64_t w = (int64_t)blockIdx.x * WT + threadIdx.x;
  const bool valid = w < W;
  double acc[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) acc[i] = 0.0;
  if (valid) {
    float la, sg, T, ven, vee, jp[4];
    walker_eval<NS, true>(s, theta, r, phi, inv_out, W, w, la, sg, T, ven, vee, jp);   // jp: Jastrow log-derivatives from the same pair loop
    const float e = T + ven + vee + s.v_nn;            // molecular.py:175
    e_l[w] = e;
    if (acc_flag) acc[3] = acc_flag[w];
    if (isfinite(e)) {
      acc[0] = 1.0; acc[1] = e; acc[2] = (double)e * e;
      const float g_aee = jp[0], g_aen = jp[1], g_bee = jp[2], g_ben = jp[3];
      acc[5] = g_aee; acc[6] = g_aen; acc[7] = g_bee; acc[8] = g_ben;
      acc[9] = (double)e * g_aee; acc[10] = (double)e * g_aen; acc[11] = (double)e * g_bee; acc[12] = (double)e * g_ben;
    } else {
      acc[4] = 1.0;
    }
  }
  __shared__ double sh[WT / 32][13];
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const double t = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < 13) {
    double t = 0;
    for (int q = 0; q < WT / 32; ++q) t += sh[q][threadIdx.x];
    part[(int64_t)blockIdx.x * NPART + threadIdx.x] = t;
  }
}
// hdr[0..4] and jas[0..7] from the block partials (one warp per slot, fixed order => deterministic)
__global__ void stats13_finish_kernel(const double* __restrict__ part, const int nblk, double* __restrict__ hdr,
                                      double* __restrict__ jas) {
  nq_pdl_trigger();                                    // the reverse pass behind this kernel may start staging its weights
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double a = 0;
  if (slot < 13) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;             // four independent loads per round (latency, not bandwidth, matters here)
    for (int b = lane; b < nblk; b += 128) {
      a0 += part[(int64_t)b * NPART + slot];
      if (b + 32 < nblk) a1 += part[(int64_t)(b + 32) * NPART + slot];
      if (b + 64 < nblk) a2 += part[(int64_t)(b + 64) * NPART + slot];
      if (b + 96 < nblk) a3 += part[(int64_t)(b + 96) * NPART + slot];
    }
    a = (a0 + a1) + (a2 + a3);
  }
  a = warp_sum(a);
  if (lane == 0) {
    if (slot < 5) hdr[slot] = a;
    else if (slot < 13) jas[slot - 5] = a;
    else if (slot < 16) hdr[slot - 8] = 0.0;          // hdr[5..7]
  }
}
// out[jastrow scalar] = sum E O - Ebar sum O   with Ebar from the (all-reduced) header
__global__ void jastrow_from_sums_kernel(const DevSys s, const double* __restrict__ jas, const double* __restrict__ hdr,
                                         double* __restrict__ out) {
  nq_pdl_trigger();                                    // multi-GPU: the gradient exchange behind this kernel may get resident
  const int q = threadIdx.x;
  if (q >= 4) return;
  const double ebar = hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0);
  const int o = q == 0 ? s.jas_a_ee : (q == 1 ? s.jas_a_en : (q == 2 ? s.jas_b_ee : s.jas_b_en));
  out[o] = jas[4 + q] - ebar * jas[q];
}

// ---------------------------------------------------------------------------------------------
// proposal: r' = r + sigma * n   (metropolis.py:119-121)
__global__ void propose_kernel(const DevSys s, const float* __restrict__ r, const float* __restrict__ normals,
                               const uint64_t seed, const uint64_t step, const int64_t wid0, const float sigma_arg,
                               const float* __restrict__ sigma_dev, const int64_t W, float* __restrict__ r_prop) {
  nq_pdl_trigger();                                    // the value pass behind this kernel may start staging its weights
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int nn = 3 * s.N;
  const float sigma = sigma_dev ? *sigma_dev : sigma_arg;   // device-resident step size (adaptive burn-in)
  if (normals != nullptr) {
    for (int c = 0; c < nn; ++c) r_prop[(int64_t)c * W + w] = fmaf(sigma, normals[(int64_t)c * W + w], r[(int64_t)c * W + w]);
  } else {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int q = 0; 4 * q < nn; ++q) {
      uint4 x = philox4x32_10(make_uint4((uint32_t)(wid0 + w), 1u + q, (uint32_t)step, (uint32_t)(step >> 32)), key);
      float n[4];
      box_muller(x.x, x.y, n[0], n[1]);
      box_muller(x.z, x.w, n[2], n[3]);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int c = 4 * q + t;
        if (c < nn) r_prop[(int64_t)c * W + w] = fmaf(sigma, n[t], r[(int64_t)c * W + w]);
      }
    }
  }
}

__global__ void rng_fill_kernel(const DevSys s, const uint64_t seed, const uint64_t step, const int64_t wid0,
                                const int64_t W, float* __restrict__ normals, float* __restrict__ uniforms) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int nn = 3 * s.N;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  if (uniforms) {
    uint4 x = philox4x32_10(make_uint4((uint32_t)(wid0 + w), 0u, (uint32_t)step, (uint32_t)(step >> 32)), key);
    uniforms[w] = u01(x.x);
  }
  if (normals)
    for (int q = 0; 4 * q < nn; ++q) {
      uint4 x = philox4x32_10(make_uint4((uint32_t)(wid0 + w), 1u + q, (uint32_t)step, (uint32_t)(step >> 32)), key);
      float n[4];
      box_muller(x.x, x.y, n[0], n[1]);
      box_muller(x.z, x.w, n[2], n[3]);
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (4 * q + t < nn) normals[(int64_t)(4 * q + t) * W + w] = n[t];
    }
}

// accept / reject (metropolis.py:128-140): strict log(u) < 2 log|psi'| - logp, in fp32
template <int NS>
__global__ void __launch_bounds__(WT) accept_kernel(const DevSys s, const float* __restrict__ theta, float* __restrict__ r,
                                                    float* __restrict__ logp, const float* __restrict__ r_prop,
                                                    const float* __restrict__ phi, const float* __restrict__ uniforms,
                                                    const uint64_t seed, const uint64_t step, const int64_t wid0,
                                                    const int64_t W, uint8_t* __restrict__ accept,
                                                    float* __restrict__ n_acc, float* __restrict__ acc_flag) {
  nq_pdl_trigger();                                    // the 5-channel pass behind this kernel may start staging its weights
  const int64_t w = (int64_t)blockIdx.x * WT + threadIdx.x;
  if (w >= W) return;
  float la, sg, T, ven, vee;
  walker_eval<NS, false>(s, theta, r_prop, phi, nullptr, W, w, la, sg, T, ven, vee);
  const float logp_prop = 2.0f * la;                   // base.py:76
  float u;
  if (uniforms != nullptr) u = uniforms[w];
  else {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 x = philox4x32_10(make_uint4((uint32_t)(wid0 + w), 0u, (uint32_t)step, (uint32_t)(step >> 32)), key);
    u = u01(x.x);
  }
  const float ratio = logp_prop - logp[w];
  const bool acc = logf(u) < ratio;                    // log(0) = -inf accepts; NaN ratio rejects
  if (acc) {
    for (int c = 0; c < 3 * s.N; ++c) r[(int64_t)c * W + w] = r_prop[(int64_t)c * W + w];
    logp[w] = logp_prop;
  }
  if (accept) accept[w] = acc ? 1 : 0;
  if (n_acc) n_acc[w] += acc ? 1.0f : 0.0f;
  if (acc_flag) acc_flag[w] = acc ? 1.0f : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// energy statistics: per-block partials then a single-block finish (deterministic order)
__global__ void __launch_bounds__(256) stats_partial_kernel(const float* __restrict__ e_l, const float* __restrict__ acc_flag,
                                                            const int64_t W, double* __restrict__ part) {
  double n = 0, se = 0, se2 = 0, na = 0, nb = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (int64_t)gridDim.x * blockDim.x) {
    const float e = e_l[w];
    if (isfinite(e)) { n += 1.0; se += e; se2 += (double)e * e; } else nb += 1.0;
    if (acc_flag) na += acc_flag[w];
  }
  __shared__ double sh[8][5];
  n = warp_sum(n); se = warp_sum(se); se2 = warp_sum(se2); na = warp_sum(na); nb = warp_sum(nb);
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) { sh[wp][0] = n; sh[wp][1] = se; sh[wp][2] = se2; sh[wp][3] = na; sh[wp][4] = nb; }
  __syncthreads();
  if (threadIdx.x < 5) {
    double a = 0;
    for (int i = 0; i < 8; ++i) a += sh[i][threadIdx.x];
    part[(int64_t)blockIdx.x * NPART + threadIdx.x] = a;
  }
}
__global__ void stats_finish_kernel(const double* __restrict__ part, const int nblk, double* __restrict__ hdr) {
  // one warp per header slot; lanes stride the block partials (fixed order => deterministic)
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double a = 0;
  if (slot < 5)
    for (int b = lane; b < nblk; b += 32) a += part[(int64_t)b * NPART + slot];
  a = warp_sum(a);
  if (lane == 0 && slot < NQMC_HDR_SIZE) hdr[slot] = slot < 5 ? a : 0.0;
}

// ---------------------------------------------------------------------------------------------
// weighted parameter gradients of the per-walker terms (Jastrow scalars, cusp networks)
//   out[p] = sum_w weight_w * d(J + C)/d theta_p
// grid (walker blocks, nuclei): block (bx, a) handles the e-n cusp network of nucleus a for its walkers; blocks with a == 0
// also take the Jastrow / e-e cusp scalars.  Inside a block the sums over the electrons are formed per thread first and
// the warp reductions follow per (nucleus, hidden unit) -- not per (electron, nucleus, hidden unit) as in round 1, where
// the shuffles were 90 % of the kernel (436 us at 32,768 walkers, H6; now grid.y = A blocks each doing 1/6 of 1/6).
__global__ void __launch_bounds__(WT) walker_grads_kernel(const DevSys s, const float* __restrict__ theta,
                                                          const float* __restrict__ r, const float* __restrict__ wgt,
                                                          const double* __restrict__ hdr, const float* __restrict__ e_l,
                                                          const int64_t W, double* __restrict__ part,
                                                          const int n_slots) {
  extern __shared__ float sacc[];     // [n_slots] block accumulators, then [2][NQMC_MAX_ELEC][WT] distances / exponentials
  float* sd = sacc + ((n_slots + 3) & ~3);
  float* se = sd + NQMC_MAX_ELEC * WT;
  for (int q = threadIdx.x; q < n_slots; q += WT) sacc[q] = 0.f;
  __syncthreads();
  const int an = blockIdx.y;          // this block's nucleus
  float center = 0.f;
  if (hdr) center = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));
  const float a_ee = theta[s.jas_a_ee], b_ee = theta[s.jas_b_ee];
  const float a_en = theta[s.jas_a_en], b_en = theta[s.jas_b_en];
  const float cbs = s.use_cusp ? theta[s.ee_cusp_b] : 0.f;
  for (int64_t w = (int64_t)blockIdx.x * WT + threadIdx.x; w < (W + WT - 1) / WT * WT; w += (int64_t)gridDim.x * WT) {
    float wt = 0.f;
    const bool valid = w < W;
    const int64_t wc = valid ? w : W - 1;
    if (valid) {
      if (wgt) wt = wgt[w];
      else if (e_l) { float e = e_l[w]; wt = isfinite(e) ? e - center : 0.f; }
      else wt = 1.f;
    }
    if (an == 0) {
      float g_aee = 0.f, g_bee = 0.f, g_aen = 0.f, g_ben = 0.f, g_cb = 0.f;
      for (int i = 0; i < s.N; ++i) {
        const float xi = r[(int64_t)(3 * i) * W + wc], yi = r[(int64_t)(3 * i + 1) * W + wc], zi = r[(int64_t)(3 * i + 2) * W + wc];
        for (int j = i + 1; j < s.N; ++j) {
          float dx = xi - r[(int64_t)(3 * j) * W + wc], dy = yi - r[(int64_t)(3 * j + 1) * W + wc], dz = zi - r[(int64_t)(3 * j + 2) * W + wc];
          float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
          float q = 1.0f / fmaf(b_ee, d, 1.0f);
          g_aee = fmaf(d, q, g_aee);
          g_bee = fmaf(-a_ee * d * d, q * q, g_bee);
          if (s.use_cusp) {
            const bool same = (i < s.n_up) == (j < s.n_up);
            float q2 = 1.0f / fmaf(fabsf(cbs), d, 1.0f);
            g_cb = fmaf(-(same ? s.cusp_same : s.cusp_diff) * d * d, q2 * q2, g_cb);
          }
        }
        for (int a = 0; a < s.A; ++a) {
          float dx = xi - s.R[3 * a], dy = yi - s.R[3 * a + 1], dz = zi - s.R[3 * a + 2];
          float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
          float q = 1.0f / fmaf(b_en, d, 1.0f);
          g_aen = fmaf(d, q, g_aen);
          g_ben = fmaf(-a_en * d * d, q * q, g_ben);
        }
      }
      if (s.use_cusp) g_cb *= (cbs < 0.f ? -1.f : (cbs > 0.f ? 1.f : 0.f));
      float v0 = warp_sum(wt * g_aee), v1 = warp_sum(wt * g_aen), v2 = warp_sum(wt * g_bee), v3 = warp_sum(wt * g_ben),
            v4 = warp_sum(wt * g_cb);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sacc[0], v0); atomicAdd(&sacc[1], v1); atomicAdd(&sacc[2], v2); atomicAdd(&sacc[3], v3);
        atomicAdd(&sacc[4], v4);
      }
    }
    if (s.use_cusp) {
      // slots: 5 + a*81 + {b0[16], W0[3][16], b1, W1[16]} in *theta leaf order* of the nucleus
      const float* W0 = theta + s.en_cusp_W0[an];
      const float* b0 = theta + s.en_cusp_b0[an];
      const float* W1 = theta + s.en_cusp_W1[an];
      const int base = 5 + an * 81;
      float swd = 0.f;
      for (int i = 0; i < s.N; ++i) {                 // this thread's electron-nucleus distances, kept in shared memory
        const float dx = r[(int64_t)(3 * i) * W + wc] - s.R[3 * an], dy = r[(int64_t)(3 * i + 1) * W + wc] - s.R[3 * an + 1],
                    dz = r[(int64_t)(3 * i + 2) * W + wc] - s.R[3 * an + 2];
        const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        sd[i * WT + threadIdx.x] = d;
        se[i * WT + threadIdx.x] = __expf(-d);
        swd = fmaf(wt, d, swd);
      }
      const float sb1 = warp_sum(swd);
      if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[base + 64], sb1);
      for (int h = 0; h < NQMC_CUSP_H; ++h) {
        const float w0 = W0[h], w1 = W0[NQMC_CUSP_H + h], w2 = W0[2 * NQMC_CUSP_H + h], bb = b0[h], wo = W1[h];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
        for (int i = 0; i < s.N; ++i) {
          const float d = sd[i * WT + threadIdx.x], e = se[i * WT + threadIdx.x];
          const float wd = wt * d;
          const float t = nq_tanh(fmaf(w0, d, fmaf(w1, d * d, fmaf(w2, e, bb))));
          const float dz_ = wd * wo * fmaf(-t, t, 1.0f);
          a0 += dz_;
          a1 = fmaf(dz_, d, a1);
          a2 = fmaf(dz_ * d, d, a2);
          a3 = fmaf(dz_, e, a3);
          a4 = fmaf(wd, t, a4);
        }
        const float v0 = warp_sum(a0), v1 = warp_sum(a1), v2 = warp_sum(a2), v3 = warp_sum(a3), v4 = warp_sum(a4);
        if ((threadIdx.x & 31) == 0) {
          atomicAdd(&sacc[base + h], v0);
          atomicAdd(&sacc[base + 16 + h], v1);
          atomicAdd(&sacc[base + 32 + h], v2);
          atomicAdd(&sacc[base + 48 + h], v3);
          atomicAdd(&sacc[base + 65 + h], v4);
        }
      }
    }
  }
  __syncthreads();
  const int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int q = threadIdx.x; q < n_slots; q += WT) part[blk * n_slots + q] = (double)sacc[q];
}

__global__ void walker_grads_finish_kernel(const DevSys s, const double* __restrict__ part, const int nblk,
                                           const int n_slots, double* __restrict__ out) {
  // one warp per slot; lanes stride the block partials
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= n_slots) return;
  double a = 0;
  for (int b = lane; b < nblk; b += 32) a += part[(int64_t)b * n_slots + q];
  a = warp_sum(a);
  if (lane != 0) return;
  int o = -1;
  if (q == 0) o = s.jas_a_ee;
  else if (q == 1) o = s.jas_a_en;
  else if (q == 2) o = s.jas_b_ee;
  else if (q == 3) o = s.jas_b_en;
  else if (q == 4) o = s.use_cusp ? s.ee_cusp_b : -1;
  else {
    const int a_ = (q - 5) / 81, t = (q - 5) % 81;
    if (t < 16) o = s.en_cusp_b0[a_] + t;
    else if (t < 64) o = s.en_cusp_W0[a_] + (t - 16);
    else if (t == 64) o = s.en_cusp_b1[a_];
    else o = s.en_cusp_W1[a_] + (t - 65);
  }
  if (o >= 0) out[o] = a;
}

// ---------------------------------------------------------------------------------------------
__global__ void aos_to_soa_kernel(const float* __restrict__ aos, float* __restrict__ soa, const int64_t W, const int nn) {
  // tile transpose through shared memory: coalesced on both sides
  __shared__ float tile[32][49];
  const int64_t w0 = (int64_t)blockIdx.x * 32;
  for (int idx = threadIdx.x; idx < 32 * nn; idx += blockDim.x) {
    int64_t g = w0 * nn + idx;
    if (g < W * nn) tile[idx / nn][idx % nn] = aos[g];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * nn; idx += blockDim.x) {
    int c = idx / 32, l = idx % 32;
    if (w0 + l < W) soa[(int64_t)c * W + w0 + l] = tile[l][c];
  }
}
__global__ void soa_to_aos_kernel(const float* __restrict__ soa, float* __restrict__ aos, const int64_t W, const int nn) {
  __shared__ float tile[32][49];
  const int64_t w0 = (int64_t)blockIdx.x * 32;
  for (int idx = threadIdx.x; idx < 32 * nn; idx += blockDim.x) {
    int c = idx / 32, l = idx % 32;
    if (w0 + l < W) tile[l][c] = soa[(int64_t)c * W + w0 + l];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * nn; idx += blockDim.x) {
    int64_t g = w0 * nn + idx;
    if (g < W * nn) aos[g] = tile[idx / nn][idx % nn];
  }
}

__global__ void fill_kernel(float* __restrict__ dst, const float v, const int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

// accept / reject from a precomputed log|psi(r')| (ansatz kind 0)
__global__ void accept_logabs_kernel(const DevSys s, float* __restrict__ r, float* __restrict__ logp,
                                     const float* __restrict__ r_prop, const float* __restrict__ la_prop,
                                     const float* __restrict__ uniforms, const uint64_t seed, const uint64_t step,
                                     const int64_t wid0, const int64_t W, uint8_t* __restrict__ accept,
                                     float* __restrict__ n_acc, float* __restrict__ acc_flag) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const float logp_prop = 2.0f * la_prop[w];
  float u;
  if (uniforms != nullptr) u = uniforms[w];
  else {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 x = philox4x32_10(make_uint4((uint32_t)(wid0 + w), 0u, (uint32_t)step, (uint32_t)(step >> 32)), key);
    u = u01(x.x);
  }
  const bool acc = logf(u) < logp_prop - logp[w];
  if (acc) {
    for (int c = 0; c < 3 * s.N; ++c) r[(int64_t)c * W + w] = r_prop[(int64_t)c * W + w];
    logp[w] = logp_prop;
  }
  if (accept) accept[w] = acc ? 1 : 0;
  if (n_acc) n_acc[w] += acc ? 1.0f : 0.0f;
  if (acc_flag) acc_flag[w] = acc ? 1.0f : 0.0f;
}

// FP32 FFMA pipe microbenchmark: 16 independent FMA chains per thread (the roofline denominator of
// the FFMA-bound orbital kernels; MEASURED_PEAKS.json has no fp32 figure)
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, const int iters, const float a, const float b) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 123.456f) out[0] = s;
}

// ---------------------------------------------------------------------------------------------
// optimizer epilogue: optax.chain(clip_by_global_norm(c), adam(lr)) + apply_updates (optimizer/vmc.py:180-184,256-261)
__global__ void __launch_bounds__(256) adam_norm_kernel(const double* __restrict__ gsum, const double* __restrict__ hdr,
                                                        const int P, double* __restrict__ part) {
  const double sc = 2.0 / fmax(hdr[NQMC_HDR_N], 1.0);
  double a = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const double g = (double)(float)(sc * gsum[i]);        // the gradient is a float32 pytree in the reference
    a += g * g;
  }
  __shared__ double sh[8];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    part[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) adam_apply_kernel(const double* __restrict__ gsum, const double* __restrict__ hdr,
                                                         const double* __restrict__ part, const int nblk, const int P,
                                                         float* __restrict__ theta, float* __restrict__ m,
                                                         float* __restrict__ v, const float lr, const float clip,
                                                         const float bc1, const float bc2, double* __restrict__ norm_out) {
  double n2 = 0;
  for (int b = 0; b < nblk; ++b) n2 += part[b];            // every thread: same order => same value
  const double norm = sqrt(n2);
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  const float scale = norm >= (double)clip ? (float)((double)clip / norm) : 1.0f;   // optax keeps g when ||g|| < max_norm
  const double sc = 2.0 / fmax(hdr[NQMC_HDR_N], 1.0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const float g = (float)(sc * gsum[i]) * scale;
    const float mi = 0.9f * m[i] + 0.1f * g;
    const float vi = 0.999f * v[i] + 0.001f * g * g;
    m[i] = mi;
    v[i] = vi;
    theta[i] -= lr * (mi / bc1) / (sqrtf(vi / bc2) + 1e-8f);
  }
}

template <typename F>
int dispatch_ns(nqmc_ctx* ctx, F&& f) {
  const int ns = ctx->sys.n_up > ctx->sys.n_dn ? ctx->sys.n_up : ctx->sys.n_dn;
  if (ns <= 1) return f(std::integral_constant<int, 1>{});
  if (ns == 2) return f(std::integral_constant<int, 2>{});
  if (ns == 3) return f(std::integral_constant<int, 3>{});
  if (ns == 4) return f(std::integral_constant<int, 4>{});
  if (ns == 5) return f(std::integral_constant<int, 5>{});
  if (ns <= 8) return f(std::integral_constant<int, 8>{});
  ctx->err = "more than 8 electrons per spin are not supported";
  return NQMC_ERR_UNSUPPORTED;
}

}  // namespace

int nq_launch_propose(nqmc_ctx* ctx, const float* r, const float* normals, uint64_t seed, uint64_t step, int64_t wid0,
                      float sigma, int64_t W, float* r_prop, cudaStream_t st) {
  propose_kernel<<<(unsigned)((W + 255) / 256), 256, 0, st>>>(ctx->sys, r, normals, seed, step, wid0, sigma,
                                                              ctx->sigma_dev, W, r_prop);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

// Adaptive proposal width (the reference declares target_acceptance / adapt_step at
// src/nqmc/sampler/metropolis.py:52-54 and never reads them): after a block of `steps` burn-in moves,
//   sigma <- sigma * exp(gain * (acceptance - target)),   clamped to [1e-3, 10] bohr,
// with the block's acceptance taken from the per-walker counters, which are zeroed for the next block.
__global__ void __launch_bounds__(1024) adapt_sigma_kernel(float* __restrict__ n_acc, const int64_t W, const float steps,
                                                           const float target, const float gain, float* __restrict__ sigma) {
  __shared__ double sh[32];
  double a = 0;
  for (int64_t w = threadIdx.x; w < W; w += blockDim.x) { a += (double)n_acc[w]; n_acc[w] = 0.f; }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    const float acc = (float)(t / ((double)W * steps));
    *sigma = fminf(fmaxf(*sigma * __expf(gain * (acc - target)), 1e-3f), 10.0f);
  }
}

// Blocked statistics of per-chain energy series ([SPEC] blocking_analysis,
// .github/phase5-issue-body.md:113-132, batched over chains): e_l[s * W + w], chain w cut into n_blocks blocks of
// bs = S / n_blocks consecutive samples; out = (number of finite blocks, sum of block means, sum of squares).
__global__ void __launch_bounds__(256) blocked_stats_kernel(const float* __restrict__ e_l, const int64_t S, const int64_t W,
                                                            const int n_blocks, double* __restrict__ out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t bs = S / n_blocks;
  double cnt = 0, s1 = 0, s2 = 0;
  if (w < W)
    for (int b = 0; b < n_blocks; ++b) {
      double a = 0;
      for (int64_t t = 0; t < bs; ++t) a += (double)e_l[((int64_t)b * bs + t) * W + w];
      a /= (double)bs;
      if (isfinite(a)) { cnt += 1; s1 += a; s2 += a * a; }
    }
  cnt = warp_sum(cnt); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, cnt); atomicAdd(out + 1, s1); atomicAdd(out + 2, s2); }
}

int nq_blocked_stats(nqmc_ctx* ctx, const float* e_l, int64_t S, int64_t W, int n_blocks, double* out, cudaStream_t st) {
  NQMC_CUDA(cudaMemsetAsync(out, 0, 3 * sizeof(double), st));
  blocked_stats_kernel<<<(unsigned)((W + 255) / 256), 256, 0, st>>>(e_l, S, W, n_blocks, out);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_adapt_sigma(nqmc_ctx* ctx, float* n_acc, int64_t W, int steps, float target, float gain, float* sigma_dev,
                   cudaStream_t st) {
  adapt_sigma_kernel<<<1, 1024, 0, st>>>(n_acc, W, (float)steps, target, gain, sigma_dev);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_launch_accept(nqmc_ctx* ctx, float* r, float* logp, const float* r_prop, const float* uniforms, uint64_t seed,
                     uint64_t step, int64_t wid0, int64_t W, uint8_t* accept, float* n_acc, cudaStream_t st) {
  return dispatch_ns(ctx, [&](auto ns) -> int {
    constexpr int NS = decltype(ns)::value;
    ProfScope ps(ctx, NQ_PROF_ACCEPT, st);
    accept_kernel<NS><<<(unsigned)((W + WT - 1) / WT), WT, 0, st>>>(ctx->sys, ctx->theta, r, logp, r_prop, ctx->phi,
                                                                     uniforms, seed, step, wid0, W, accept, n_acc,
                                                                     ctx->wgt);
    NQMC_LAUNCH_CHECK();
    return (int)NQMC_OK;
  });
}

int nq_launch_assemble(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, bool want_energy,
                       float* logabs, float* sign, cudaStream_t st) {
  return dispatch_ns(ctx, [&](auto ns) -> int {
    constexpr int NS = decltype(ns)::value;
    ProfScope ps(ctx, NQ_PROF_ASSEMBLE, st);
    assemble_kernel<NS><<<(unsigned)((W + WT - 1) / WT), WT, 0, st>>>(ctx->sys, ctx->theta, r, ctx->phi,
                                                                       want_energy ? ctx->inv : nullptr, W, e_l, parts,
                                                                       logabs, sign, want_energy ? 1 : 0);
    NQMC_LAUNCH_CHECK();
    return (int)NQMC_OK;
  });
}

int nq_launch_stats(nqmc_ctx* ctx, const float* e_l, const float* acc_flag, int64_t W, double* hdr, cudaStream_t st) {
  int nblk = (int)((W + 255) / 256);
  if (nblk > ctx->n_blk_walk) nblk = ctx->n_blk_walk;
  if (nblk < 1) nblk = 1;
  stats_partial_kernel<<<nblk, 256, 0, st>>>(e_l, acc_flag, W, ctx->part_walk);
  NQMC_LAUNCH_CHECK();
  stats_finish_kernel<<<1, 32 * NQMC_HDR_SIZE, 0, st>>>(ctx->part_walk, nblk, hdr);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_launch_walker_grads(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr_center, const float* e_l,
                           int64_t W, double* out, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const int n_slots = 5 + (s.use_cusp ? 81 * s.A : 0);
  const int ny = s.use_cusp ? s.A : 1;               // one block row per nucleus (e-n cusp networks)
  int nblk = (int)((W + WT - 1) / WT);
  const int cap = (int)((int64_t)ctx->n_blk_walk * NPART / n_slots) / ny;
  if (nblk > cap) nblk = cap;
  if (nblk > 4 * ctx->sm_count) nblk = 4 * ctx->sm_count;
  if (nblk < 1) nblk = 1;
  const size_t smem = (((size_t)n_slots + 3) & ~(size_t)3) * sizeof(float) + (s.use_cusp ? 2 * NQMC_MAX_ELEC * WT * sizeof(float) : 0);
  walker_grads_kernel<<<dim3(nblk, ny), WT, smem, st>>>(s, ctx->theta, r, wgt, hdr_center, e_l, W, ctx->part_walk, n_slots);
  NQMC_LAUNCH_CHECK();
  walker_grads_finish_kernel<<<(n_slots * 32 + 127) / 128, 128, 0, st>>>(s, ctx->part_walk, nblk * ny, n_slots, out);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_aos_soa(nqmc_ctx* ctx, const float* src, float* dst, int64_t W, bool to_soa, cudaStream_t st) {
  const int nn = 3 * ctx->sys.N;
  const unsigned g = (unsigned)((W + 31) / 32);
  if (to_soa) aos_to_soa_kernel<<<g, 256, 0, st>>>(src, dst, W, nn);
  else soa_to_aos_kernel<<<g, 256, 0, st>>>(src, dst, W, nn);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_rng_fill(nqmc_ctx* ctx, uint64_t seed, uint64_t step, int64_t wid0, int64_t W, float* normals, float* uniforms,
                cudaStream_t st) {
  rng_fill_kernel<<<(unsigned)((W + 255) / 256), 256, 0, st>>>(ctx->sys, seed, step, wid0, W, normals, uniforms);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_fill(nqmc_ctx* ctx, float* dst, float v, int64_t n, cudaStream_t st) {
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, v, n);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_accept_logabs(nqmc_ctx* ctx, float* r, float* logp, const float* r_prop, const float* logabs_prop,
                     const float* uniforms, uint64_t seed, uint64_t step, int64_t wid0, int64_t W, uint8_t* accept,
                     float* n_acc, cudaStream_t st) {
  accept_logabs_kernel<<<(unsigned)((W + 255) / 256), 256, 0, st>>>(ctx->sys, r, logp, r_prop, logabs_prop, uniforms, seed,
                                                                    step, wid0, W, accept, n_acc, ctx->wgt);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

// value-only Slater pass that also stores the inverse orbital matrices (seeds of the reverse pass)
int nq_launch_assemble_inv(nqmc_ctx* ctx, const float* r, int64_t W, cudaStream_t st) {
  return dispatch_ns(ctx, [&](auto ns) -> int {
    constexpr int NS = decltype(ns)::value;
    assemble_kernel<NS><<<(unsigned)((W + WT - 1) / WT), WT, 0, st>>>(ctx->sys, ctx->theta, r, ctx->phi, ctx->inv, W,
                                                                       nullptr, nullptr, nullptr, nullptr, 0);
    NQMC_LAUNCH_CHECK();
    return (int)NQMC_OK;
  });
}

// returns FP32 FMA throughput in TFLOP/s (2 flop per FMA), best of `reps`
int nq_ffma_peak(nqmc_ctx* ctx, int reps, double* tflops, cudaStream_t st) {
  cudaEvent_t e0, e1;
  NQMC_CUDA(cudaEventCreate(&e0));
  NQMC_CUDA(cudaEventCreate(&e1));
  const int iters = 4096, blocks = ctx->sm_count * 8;
  double best = 0;
  for (int rpt = 0; rpt < reps + 1; ++rpt) {
    NQMC_CUDA(cudaEventRecord(e0, st));
    ffma_peak_kernel<<<blocks, 256, 0, st>>>(ctx->logabs, iters, 0.999f, 0.001f);
    NQMC_CUDA(cudaEventRecord(e1, st));
    NQMC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NQMC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 64.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
    if (rpt > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return NQMC_OK;
}

int nq_adam_step(nqmc_ctx* ctx, const double* gsum, const double* hdr, float* m, float* v, int64_t count, float lr,
                 float clip, double* norm_out, cudaStream_t st) {
  const int P = ctx->sys.P;
  int nblk = (P + 255) / 256;
  if (nblk > 128) nblk = 128;
  adam_norm_kernel<<<nblk, 256, 0, st>>>(gsum, hdr, P, ctx->part_walk);
  NQMC_LAUNCH_CHECK();
  const float bc1 = 1.0f - powf(0.9f, (float)count), bc2 = 1.0f - powf(0.999f, (float)count);
  ctx->theta_dirty = true;                       // see nq_launch_pdl_guarded
  ctx->act_valid_for = nullptr;
  adam_apply_kernel<<<(P + 255) / 256, 256, 0, st>>>(gsum, hdr, ctx->part_walk, nblk, P, ctx->theta, m, v, lr, clip, bc1, bc2,
                                                     norm_out);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

// E_L assembly fused with the energy header and the Jastrow-scalar sums (walker step without cusp / backflow)
int nq_launch_assemble_stats(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, const float* acc_flag, double* hdr,
                             cudaStream_t st) {
  const int nblk = (int)((W + WT - 1) / WT);
  if (nblk > ctx->n_blk_walk) {
    ctx->err = "assemble_stats: too many walkers for the partial buffer";
    return NQMC_ERR_STATE;
  }
  int rc = dispatch_ns(ctx, [&](auto ns) -> int {
    constexpr int NS = decltype(ns)::value;
    ProfScope ps(ctx, NQ_PROF_ASSEMBLE, st);
    assemble_stats_kernel<NS><<<nblk, WT, 0, st>>>(ctx->sys, ctx->theta, r, ctx->phi, ctx->inv, W, e_l, acc_flag,
                                                   ctx->part_walk);
    NQMC_LAUNCH_CHECK();
    return (int)NQMC_OK;
  });
  if (rc) return rc;
  stats13_finish_kernel<<<1, 32 * 16, 0, st>>>(ctx->part_walk, nblk, hdr, ctx->jas_sums);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int nq_launch_jastrow_from_sums(nqmc_ctx* ctx, const double* hdr, double* out, cudaStream_t st) {
  jastrow_from_sums_kernel<<<1, 32, 0, st>>>(ctx->sys, ctx->jas_sums, hdr, out);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
