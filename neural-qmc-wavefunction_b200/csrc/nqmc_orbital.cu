// nqmc_orbital.cu -- the orbital networks: batched forward (value or value+gradient+Laplacian
// channels) and the reverse pass that accumulates weighted parameter gradients.
//
// Replaces, for the [SPEC] NeuralOrbital / OrbitalSet (.github/phase2-issue-body.md:78-113):
//   * vmap(log_prob_fn) forward           src/nqmc/sampler/metropolis.py:124
//   * grad + dense hessian of log|psi|     src/nqmc/hamiltonian/molecular.py:60-64   (here: forward
//     Laplacian, SURVEY Appendix C.1 -- 5 channels {value, d/dx, d/dy, d/dz, Laplacian} per row)
//   * vmap(jax.grad(... params))           src/nqmc/optimizer/vmc.py:120-133
//
// Decomposition ("MLP-major"): a CTA owns ONE orbital network m = (spin, k), stages its weights in
// shared memory once, and loops over tiles of TR evaluation rows.  A row is (walker w, electron i of
// that spin); a tile is TR consecutive walkers at fixed i, so SoA position loads are coalesced.
// The hidden layers are register-tiled FP32 GEMMs: thread (rg, cg) owns RPT rows x 4 columns x C
// channels of the tile; activations live in shared memory as A[k][channel][row] so one LDS.128
// feeds 4 rows, weights as Wp[k][4*cg+jj] so one LDS.128 feeds 4 columns.  Column ownership is
// interleaved (thread cg owns columns cg, cg+CG, cg+2CG, cg+3CG) and the k-slice stride is
// C*TR+4 floats, which makes the epilogue's STS.128 of freshly activated columns bank-conflict free.
#include <cstdlib>

#include "nqmc_internal.cuh"

namespace {

template <int H, int C, int RPT, int NT>
struct Cfg {
  static constexpr int CG = H / 4;
  static constexpr int RG = NT / CG;
  static constexpr int TR = RPT * RG;
  static constexpr int SJ = C * TR + 4;
  static_assert(NT % CG == 0 && CG <= 32 && (32 % CG) == 0, "column groups must tile a warp");
};

__host__ __device__ constexpr int perm_col(int q, int CG) { return (q >> 2) + CG * (q & 3); }

// tanh rules of SURVEY Appendix C.1 applied in place to one (row, column) entry
template <int C>
__device__ __forceinline__ void act_rules(float (&v)[C]) {
  float h = nq_tanh(v[0]);
  if constexpr (C == 5) {
    float tp = fmaf(-h, h, 1.0f);
    float tpp = -2.0f * h * tp;
    float g2 = fmaf(v[1], v[1], fmaf(v[2], v[2], v[3] * v[3]));
    v[4] = fmaf(tp, v[4], tpp * g2);
    v[1] *= tp;
    v[2] *= tp;
    v[3] *= tp;
  }
  if constexpr (C == 10) {   // channels: value, d/dx d/dy d/dz, Hessian xx xy xz yy yz zz
    float tp = fmaf(-h, h, 1.0f);
    float tpp = -2.0f * h * tp;
    v[4] = fmaf(tp, v[4], tpp * v[1] * v[1]);
    v[5] = fmaf(tp, v[5], tpp * v[1] * v[2]);
    v[6] = fmaf(tp, v[6], tpp * v[1] * v[3]);
    v[7] = fmaf(tp, v[7], tpp * v[2] * v[2]);
    v[8] = fmaf(tp, v[8], tpp * v[2] * v[3]);
    v[9] = fmaf(tp, v[9], tpp * v[3] * v[3]);
    v[1] *= tp;
    v[2] *= tp;
    v[3] *= tp;
  }
  v[0] = h;
}

// Packed FP32 FMA (Blackwell FFMA2): two rows per instruction.  ptxas folds the duplicated weight
// into a broadcast operand (FFMA2 Rd, Ra.F32x2.HI_LO, Rw.F32, Rd.F32x2.HI_LO), so the FMA pipe sees
// the same work in half the issue slots -- the orbital GEMM loops are issue-bound otherwise.
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
// acc[rp][jj] (rows 2rp, 2rp+1 ; column jj) += a[rows] * w[jj]
template <int RPT>
__device__ __forceinline__ void fma_tile(float2 (&acc)[RPT / 2][4], const float (&a)[RPT], const float4& wv) {
#pragma unroll
  for (int rp = 0; rp < RPT / 2; ++rp) {
    const float2 ap = make_float2(a[2 * rp], a[2 * rp + 1]);
    acc[rp][0] = ffma2(ap, make_float2(wv.x, wv.x), acc[rp][0]);
    acc[rp][1] = ffma2(ap, make_float2(wv.y, wv.y), acc[rp][1]);
    acc[rp][2] = ffma2(ap, make_float2(wv.z, wv.z), acc[rp][2]);
    acc[rp][3] = ffma2(ap, make_float2(wv.w, wv.w), acc[rp][3]);
  }
}

// RPT consecutive rows from shared memory (one 64/128-bit load per 2/4 rows)
template <int RPT>
__device__ __forceinline__ void load_rows(const float* p, float (&a)[RPT]) {
  if constexpr (RPT == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    a[0] = t.x; a[1] = t.y;
  } else {
#pragma unroll
    for (int v = 0; v < RPT / 4; ++v) {
      float4 t = *reinterpret_cast<const float4*>(p + 4 * v);
      a[4 * v] = t.x; a[4 * v + 1] = t.y; a[4 * v + 2] = t.z; a[4 * v + 3] = t.w;
    }
  }
}
template <int RPT>
__device__ __forceinline__ void store_rows(float* p, const float (&a)[RPT]) {
  if constexpr (RPT == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(a[0], a[1]);
  } else {
#pragma unroll
    for (int v = 0; v < RPT / 4; ++v)
      *reinterpret_cast<float4*>(p + 4 * v) = make_float4(a[4 * v], a[4 * v + 1], a[4 * v + 2], a[4 * v + 3]);
  }
}

// orbital input features of one electron (phase2-issue-body.md:83-93) with their gradient / Laplacian
// (Appendix C.1 input rules), written as F[f][c][row]
template <int C>
__device__ __forceinline__ void write_features(const DevSys& s, float x, float y, float z, float* Fs, int SJ, int TR,
                                               int row) {
  const float p[3] = {x, y, z};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float* f = Fs + d * SJ + row;
    f[0] = p[d];
    if constexpr (C >= 5) {
      f[1 * TR] = d == 0 ? 1.f : 0.f;
      f[2 * TR] = d == 1 ? 1.f : 0.f;
      f[3 * TR] = d == 2 ? 1.f : 0.f;
#pragma unroll
      for (int c = 4; c < C; ++c) f[c * TR] = 0.f;
    }
  }
  for (int a = 0; a < s.A; ++a) {
    float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
    float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    float d = sqrtf(d2);
    float e = __expf(-d);
    float* fd = Fs + (3 + a) * SJ + row;
    float* fe = Fs + (3 + s.A + a) * SJ + row;
    fd[0] = d;
    fe[0] = e;
    if constexpr (C == 5) {
      float inv = 1.0f / d;
      float ux = dx * inv, uy = dy * inv, uz = dz * inv;
      fd[1 * TR] = ux;
      fd[2 * TR] = uy;
      fd[3 * TR] = uz;
      fd[4 * TR] = 2.0f * inv;
      fe[1 * TR] = -e * ux;
      fe[2 * TR] = -e * uy;
      fe[3 * TR] = -e * uz;
      fe[4 * TR] = e * (1.0f - 2.0f * inv);
    }
    if constexpr (C == 10) {
      // Hess d = (I - u u^T)/d ; Hess e^-d = e [u u^T (1 + 1/d) - I/d]
      float inv = 1.0f / d;
      float u[3] = {dx * inv, dy * inv, dz * inv};
      fd[1 * TR] = u[0]; fd[2 * TR] = u[1]; fd[3 * TR] = u[2];
      fe[1 * TR] = -e * u[0]; fe[2 * TR] = -e * u[1]; fe[3 * TR] = -e * u[2];
      int c = 4;
#pragma unroll
      for (int a_ = 0; a_ < 3; ++a_)
#pragma unroll
        for (int b_ = a_; b_ < 3; ++b_) {
          float uu = u[a_] * u[b_], dl = a_ == b_ ? 1.f : 0.f;
          fd[c * TR] = (dl - uu) * inv;
          fe[c * TR] = e * (uu * (1.0f + inv) - dl * inv);
          ++c;
        }
    }
  }
}

// ============================================================================================
// forward: phi[c][ent][w] for every (walker, electron-of-spin, orbital)
template <int H, int C, int RPT, int NT>
__global__ void __launch_bounds__(NT, (H <= 64) ? 2 : 1)
    orb_eval_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r, const int64_t W,
                    float* __restrict__ phi, const int n_chunks) {
  using K = Cfg<H, C, RPT, NT>;
  constexpr int CG = K::CG, TR = K::TR, SJ = K::SJ;
  extern __shared__ __align__(16) float smem[];
  const int F = s.F;
  float* W1p = smem;                 // [F][H]
  float* W2p = W1p + F * H;          // [H][H]
  float* b1p = W2p + H * H;          // [H]
  float* b2p = b1p + H;              // [H]
  float* W3p = b2p + H;              // [H]
  float* As = W3p + H;               // [H][SJ]   (features alias its first F slices)

  const int tid = threadIdx.x;
  const int cg = tid % CG, rg = tid / CG;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];

  for (int idx = tid; idx < F * H; idx += NT) {
    int f = idx / H, q = idx % H;
    W1p[idx] = theta[off.W1 + f * H + perm_col(q, CG)];
  }
  for (int idx = tid; idx < H * H; idx += NT) {
    int k = idx / H, q = idx % H;
    W2p[idx] = theta[off.W2 + k * H + perm_col(q, CG)];
  }
  for (int q = tid; q < H; q += NT) {
    int j = perm_col(q, CG);
    b1p[q] = theta[off.b1 + j];
    b2p[q] = theta[off.b2 + j];
    W3p[q] = theta[off.W3 + j];
  }
  const float b3 = theta[off.b3];
  __syncthreads();

  const int64_t nwb = (W + TR - 1) / TR;
  const int64_t n_items = nwb * ns;
  for (int64_t item = chunk; item < n_items; item += n_chunks) {
    const int i = (int)(item % ns);
    const int64_t w0 = (item / ns) * TR;
    // ---- features ------------------------------------------------------------------------
    if (tid < TR) {
      int64_t w = w0 + tid;
      if (w >= W) w = W - 1;
      const int e = elec0 + i;
      float x = r[(int64_t)(3 * e + 0) * W + w], y = r[(int64_t)(3 * e + 1) * W + w], z = r[(int64_t)(3 * e + 2) * W + w];
      write_features<C>(s, x, y, z, As, SJ, TR, tid);
    }
    __syncthreads();

    float2 acc[C][RPT / 2][4];
#define ACC(c, rr, jj) ((rr) & 1 ? acc[c][(rr) >> 1][jj].y : acc[c][(rr) >> 1][jj].x)
    // ---- layer 1 -------------------------------------------------------------------------
    {
      const float4 bv = *reinterpret_cast<const float4*>(b1p + 4 * cg);
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int rp = 0; rp < RPT / 2; ++rp) {
          acc[c][rp][0] = c == 0 ? make_float2(bv.x, bv.x) : make_float2(0.f, 0.f);
          acc[c][rp][1] = c == 0 ? make_float2(bv.y, bv.y) : make_float2(0.f, 0.f);
          acc[c][rp][2] = c == 0 ? make_float2(bv.z, bv.z) : make_float2(0.f, 0.f);
          acc[c][rp][3] = c == 0 ? make_float2(bv.w, bv.w) : make_float2(0.f, 0.f);
        }
      for (int f = 0; f < F; ++f) {
        const float4 wv = *reinterpret_cast<const float4*>(W1p + f * H + 4 * cg);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float a[RPT];
          load_rows<RPT>(As + f * SJ + c * TR + RPT * rg, a);
          fma_tile<RPT>(acc[c], a, wv);
        }
      }
    }
    __syncthreads();   // features are dead: A may now overwrite them
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = cg + CG * jj;
#pragma unroll
      for (int rr = 0; rr < RPT; ++rr) {
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = ACC(c, rr, jj);
        act_rules<C>(v);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          if (rr & 1) acc[c][rr >> 1][jj].y = v[c];
          else acc[c][rr >> 1][jj].x = v[c];
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float o[RPT];
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) o[rr] = ACC(c, rr, jj);
        store_rows<RPT>(As + j * SJ + c * TR + RPT * rg, o);
      }
    }
    __syncthreads();
    // ---- layer 2 -------------------------------------------------------------------------
    {
      const float4 bv = *reinterpret_cast<const float4*>(b2p + 4 * cg);
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int rp = 0; rp < RPT / 2; ++rp) {
          acc[c][rp][0] = c == 0 ? make_float2(bv.x, bv.x) : make_float2(0.f, 0.f);
          acc[c][rp][1] = c == 0 ? make_float2(bv.y, bv.y) : make_float2(0.f, 0.f);
          acc[c][rp][2] = c == 0 ? make_float2(bv.z, bv.z) : make_float2(0.f, 0.f);
          acc[c][rp][3] = c == 0 ? make_float2(bv.w, bv.w) : make_float2(0.f, 0.f);
        }
#pragma unroll 2
      for (int k = 0; k < H; ++k) {
        const float4 wv = *reinterpret_cast<const float4*>(W2p + k * H + 4 * cg);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float a[RPT];
          load_rows<RPT>(As + k * SJ + c * TR + RPT * rg, a);
          fma_tile<RPT>(acc[c], a, wv);
        }
      }
    }
    // ---- output layer: phi = W3 . h2 + b3 ------------------------------------------------
    {
      const float4 w3 = *reinterpret_cast<const float4*>(W3p + 4 * cg);
      const float w3a[4] = {w3.x, w3.y, w3.z, w3.w};
      float out[C][RPT];
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) out[c][rr] = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
          float v[C];
#pragma unroll
          for (int c = 0; c < C; ++c) v[c] = ACC(c, rr, jj);
          act_rules<C>(v);
#pragma unroll
          for (int c = 0; c < C; ++c) out[c][rr] = fmaf(w3a[jj], v[c], out[c][rr]);
        }
#pragma unroll
      for (int o = CG / 2; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int rr = 0; rr < RPT; ++rr) out[c][rr] += __shfl_xor_sync(0xffffffffu, out[c][rr], o);
      if (cg == 0) {
        const int ent = ent_index(s, spin, i, korb);
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
          const int64_t w = w0 + RPT * rg + rr;
          if (w < W) {
#pragma unroll
            for (int c = 0; c < C; ++c)
              phi[((int64_t)c * s.n_ent + ent) * W + w] = out[c][rr] + (c == 0 ? b3 : 0.f);
          }
        }
      }
    }
    __syncthreads();   // next tile's features overwrite A
  }
#undef ACC
}

// ============================================================================================
// reverse pass: out[p] += sum_rows weight[w] * inv[w][k][i] * d phi_m(row) / d theta_p
//
// Per tile (TR rows): recompute the value-only forward (h1, h2), then
//   delta2 = seed * W3 * (1 - h2^2)            dW3 += seed * h2          db3 += seed
//   delta1 = (delta2 W2^T) * (1 - h1^2)        dW2 += h1^T delta2        db2 += delta2
//                                              dW1 += f^T  delta1        db1 += delta1
// with seed[row] = weight[w] * inv[ent][w].  Weight gradients stay in registers across all tiles of
// the CTA and are flushed once to a per-CTA partial slab (no atomics on the hot path).
template <int H, int NT>
struct BCfg {
  static constexpr int CG = H / 4;
  static constexpr int RG = NT / CG;
  static constexpr int TR = 4 * RG;
  static constexpr int SR = TR + 4;          // stride of [k][row] arrays
  static constexpr int SH = H + 4;           // stride of [row][k] arrays
  static constexpr int KPT = H / RG;         // k's per thread in the dW2 outer product
  static constexpr int FG = NT / H;          // feature groups in the dW1 product
  static constexpr int A1SZ = (H * SR > TR * SH) ? H * SR : TR * SH;   // A1 ([H][SR]) is later reused as D1r ([TR][SH])
  static_assert(H % RG == 0, "row groups must divide the hidden width");
  static_assert(KPT == 1 || KPT == 2 || KPT % 4 == 0, "k's per thread");
  static size_t smem_floats(int F) {
    return (size_t)F * H + 2 * (size_t)H * H + 3 * H + (size_t)F * SR + A1SZ + 2 * (size_t)TR * SH + (size_t)H * SR + TR +
           4 * H + 4;
  }
};

template <int H, int NT>
__global__ void __launch_bounds__(NT, (H <= 64) ? 2 : 1)
    orb_bwd_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r,
                   const float* __restrict__ inv, const float* __restrict__ wgt, const double* __restrict__ hdr,
                   const float* __restrict__ e_l, const int64_t W, float* __restrict__ part, const int n_chunks,
                   const int p_mlp) {
  using K = BCfg<H, NT>;
  constexpr int CG = K::CG, TR = K::TR, SR = K::SR, SH = K::SH, KPT = K::KPT, FG = K::FG;
  constexpr int MAXFF = (35 + FG - 1) / FG;
  extern __shared__ __align__(16) float smem[];
  const int F = s.F;
  float* W1p = smem;                 // [F][H]    columns permuted
  float* W2p = W1p + F * H;          // [H][H]    W2p[k][q]   = W2[k][perm(q)]
  float* W2T = W2p + H * H;          // [H][H]    W2T[q][p]   = W2[perm(p)][perm(q)]
  float* b1p = W2T + H * H;
  float* b2p = b1p + H;
  float* W3p = b2p + H;
  float* Fs = W3p + H;               // [F][SR]   features (value only)
  float* A1 = Fs + F * SR;           // [H][SR]   h1 as [k][row]   (layer-2 operand) ...
  float* D1r = A1;                   // [TR][SH]  ... reused for delta1 as [row][p] once layer 2 is done
  float* H1r = A1 + K::A1SZ;         // [TR][SH]  h1 as [row][p]
  float* D2r = H1r + TR * SH;        // [TR][SH]  delta2 as [row][q]
  float* D2T = D2r + TR * SH;        // [H][SR]   delta2 as [q][row]
  float* seedS = D2T + H * SR;       // [TR]
  float* redS = seedS + TR;          // [4*H + 4] bias / W3 partial reduction

  const int tid = threadIdx.x;
  const int cg = tid % CG, rg = tid / CG;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];

  for (int idx = tid; idx < F * H; idx += NT) {
    int f = idx / H, q = idx % H;
    W1p[idx] = theta[off.W1 + f * H + perm_col(q, CG)];
  }
  for (int idx = tid; idx < H * H; idx += NT) {
    int k = idx / H, q = idx % H;
    W2p[idx] = theta[off.W2 + k * H + perm_col(q, CG)];
  }
  for (int idx = tid; idx < H * H; idx += NT) {
    int q = idx / H, p = idx % H;
    W2T[idx] = theta[off.W2 + perm_col(p, CG) * H + perm_col(q, CG)];
  }
  for (int q = tid; q < H; q += NT) {
    int j = perm_col(q, CG);
    b1p[q] = theta[off.b1 + j];
    b2p[q] = theta[off.b2 + j];
    W3p[q] = theta[off.W3 + j];
  }
  for (int q = tid; q < 4 * H + 4; q += NT) redS[q] = 0.f;
  float centerf = 0.f;
  if (hdr != nullptr) centerf = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));
  __syncthreads();

  // persistent accumulators
  float gW2[KPT][4];       // dW2[p = rg*KPT + kk][q = 4*cg + jj]
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) gW2[kk][jj] = 0.f;
  float gW1[MAXFF];        // dW1[f = fgrp + FG*ff][p = tid % H]
#pragma unroll
  for (int ff = 0; ff < MAXFF; ++ff) gW1[ff] = 0.f;
  float gb1 = 0.f;         // db1[p = tid % H] (threads with fgrp == 0)
  float gb2[4] = {0.f, 0.f, 0.f, 0.f}, gW3[4] = {0.f, 0.f, 0.f, 0.f};   // partial over this thread's rows
  float gb3 = 0.f;

  const int64_t nwb = (W + TR - 1) / TR;
  const int64_t n_items = nwb * ns;
  for (int64_t item = chunk; item < n_items; item += n_chunks) {
    const int i = (int)(item % ns);
    const int64_t w0 = (item / ns) * TR;
    const int ent = ent_index(s, spin, i, korb);
    if (tid < TR) {
      int64_t w = w0 + tid;
      const bool valid = w < W;
      if (!valid) w = W - 1;
      const int e = elec0 + i;
      float x = r[(int64_t)(3 * e + 0) * W + w], y = r[(int64_t)(3 * e + 1) * W + w], z = r[(int64_t)(3 * e + 2) * W + w];
      write_features<1>(s, x, y, z, Fs, SR, TR, tid);
      float wt = 1.0f;
      if (wgt != nullptr) wt = wgt[w];
      else if (e_l != nullptr) {
        float ev = e_l[w];
        wt = isfinite(ev) ? ev - centerf : 0.f;
      }
      float sd = wt * inv[(int64_t)ent * W + w];
      seedS[tid] = (valid && isfinite(sd)) ? sd : 0.f;
    }
    __syncthreads();
    // ---- forward layer 1 -----------------------------------------------------------------
    float h1[4][4], h2[4][4];
    {
      const float4 bv = *reinterpret_cast<const float4*>(b1p + 4 * cg);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) { h1[rr][0] = bv.x; h1[rr][1] = bv.y; h1[rr][2] = bv.z; h1[rr][3] = bv.w; }
      for (int f = 0; f < F; ++f) {
        const float4 wv = *reinterpret_cast<const float4*>(W1p + f * H + 4 * cg);
        const float4 av = *reinterpret_cast<const float4*>(Fs + f * SR + 4 * rg);
        const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          h1[rr][0] = fmaf(a[rr], wv.x, h1[rr][0]);
          h1[rr][1] = fmaf(a[rr], wv.y, h1[rr][1]);
          h1[rr][2] = fmaf(a[rr], wv.z, h1[rr][2]);
          h1[rr][3] = fmaf(a[rr], wv.w, h1[rr][3]);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) h1[rr][jj] = nq_tanh(h1[rr][jj]);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        *reinterpret_cast<float4*>(A1 + (cg + CG * jj) * SR + 4 * rg) = make_float4(h1[0][jj], h1[1][jj], h1[2][jj], h1[3][jj]);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        *reinterpret_cast<float4*>(H1r + (4 * rg + rr) * SH + 4 * cg) = make_float4(h1[rr][0], h1[rr][1], h1[rr][2], h1[rr][3]);
    }
    __syncthreads();
    // ---- forward layer 2, delta2 ---------------------------------------------------------
    {
      const float4 bv = *reinterpret_cast<const float4*>(b2p + 4 * cg);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) { h2[rr][0] = bv.x; h2[rr][1] = bv.y; h2[rr][2] = bv.z; h2[rr][3] = bv.w; }
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float4 wv = *reinterpret_cast<const float4*>(W2p + k * H + 4 * cg);
        const float4 av = *reinterpret_cast<const float4*>(A1 + k * SR + 4 * rg);
        const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          h2[rr][0] = fmaf(a[rr], wv.x, h2[rr][0]);
          h2[rr][1] = fmaf(a[rr], wv.y, h2[rr][1]);
          h2[rr][2] = fmaf(a[rr], wv.z, h2[rr][2]);
          h2[rr][3] = fmaf(a[rr], wv.w, h2[rr][3]);
        }
      }
      const float4 w3 = *reinterpret_cast<const float4*>(W3p + 4 * cg);
      const float w3a[4] = {w3.x, w3.y, w3.z, w3.w};
      const float4 sv = *reinterpret_cast<const float4*>(seedS + 4 * rg);
      const float sd[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        gb3 += (cg == 0) ? sd[rr] : 0.f;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float h = nq_tanh(h2[rr][jj]);
          gW3[jj] = fmaf(sd[rr], h, gW3[jj]);
          float d2 = sd[rr] * w3a[jj] * fmaf(-h, h, 1.0f);
          gb2[jj] += d2;
          h2[rr][jj] = d2;       // h2 now holds delta2
        }
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        *reinterpret_cast<float4*>(D2r + (4 * rg + rr) * SH + 4 * cg) = make_float4(h2[rr][0], h2[rr][1], h2[rr][2], h2[rr][3]);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        *reinterpret_cast<float4*>(D2T + (4 * cg + jj) * SR + 4 * rg) = make_float4(h2[0][jj], h2[1][jj], h2[2][jj], h2[3][jj]);
    }
    __syncthreads();   // A1 is dead from here on (D1r reuses it)
    // ---- delta1 = (delta2 W2^T) * (1 - h1^2)   thread: rows 4rg.., p = 4cg.. ----------------
    {
      float d1[4][4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) d1[rr][jj] = 0.f;
#pragma unroll 4
      for (int q = 0; q < H; ++q) {
        const float4 wv = *reinterpret_cast<const float4*>(W2T + q * H + 4 * cg);
        const float4 av = *reinterpret_cast<const float4*>(D2T + q * SR + 4 * rg);
        const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          d1[rr][0] = fmaf(a[rr], wv.x, d1[rr][0]);
          d1[rr][1] = fmaf(a[rr], wv.y, d1[rr][1]);
          d1[rr][2] = fmaf(a[rr], wv.z, d1[rr][2]);
          d1[rr][3] = fmaf(a[rr], wv.w, d1[rr][3]);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) d1[rr][jj] *= fmaf(-h1[rr][jj], h1[rr][jj], 1.0f);
        *reinterpret_cast<float4*>(D1r + (4 * rg + rr) * SH + 4 * cg) = make_float4(d1[rr][0], d1[rr][1], d1[rr][2], d1[rr][3]);
      }
    }
    // ---- dW2[p][q] += sum_rows h1[row][p] * delta2[row][q]   thread: p-group rg, q-group cg ----
    {
#pragma unroll 4
      for (int row = 0; row < TR; ++row) {
        const float4 dv = *reinterpret_cast<const float4*>(D2r + row * SH + 4 * cg);
        float hv[KPT];
        if constexpr (KPT >= 4) {
#pragma unroll
          for (int v = 0; v < KPT / 4; ++v) {
            float4 t = *reinterpret_cast<const float4*>(H1r + row * SH + rg * KPT + 4 * v);
            hv[4 * v] = t.x; hv[4 * v + 1] = t.y; hv[4 * v + 2] = t.z; hv[4 * v + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < KPT; ++kk) hv[kk] = H1r[row * SH + rg * KPT + kk];
        }
#pragma unroll
        for (int kk = 0; kk < KPT; ++kk) {
          gW2[kk][0] = fmaf(hv[kk], dv.x, gW2[kk][0]);
          gW2[kk][1] = fmaf(hv[kk], dv.y, gW2[kk][1]);
          gW2[kk][2] = fmaf(hv[kk], dv.z, gW2[kk][2]);
          gW2[kk][3] = fmaf(hv[kk], dv.w, gW2[kk][3]);
        }
      }
    }
    __syncthreads();   // D1r complete
    // ---- dW1[f][p] += sum_rows f[row] * delta1[row][p] ; db1[p] += sum_rows delta1 ------------
    {
      const int p = tid % H, fg = tid / H;
      for (int row = 0; row < TR; row += 4) {
        const float dv0 = D1r[(row + 0) * SH + p], dv1 = D1r[(row + 1) * SH + p], dv2 = D1r[(row + 2) * SH + p],
                    dv3 = D1r[(row + 3) * SH + p];
        if (fg == 0) gb1 += (dv0 + dv1) + (dv2 + dv3);
#pragma unroll
        for (int ff = 0; ff < MAXFF; ++ff) {
          const int f = fg + FG * ff;
          if (f < F) {
            const float4 fv = *reinterpret_cast<const float4*>(Fs + f * SR + row);
            gW1[ff] = fmaf(fv.x, dv0, fmaf(fv.y, dv1, fmaf(fv.z, dv2, fmaf(fv.w, dv3, gW1[ff]))));
          }
        }
      }
    }
    __syncthreads();   // tile buffers free
  }

  // ---- flush: per-CTA partial slab in the network's own theta order ---------------------------
  // slab layout = theta block of the network: b1[H] W1[F][H] b2[H] W2[H][H] b3 W3[H]  (relative offsets)
  float* slab = part + (int64_t)blockIdx.x * p_mlp;
  const int o_b1 = 0, o_W1 = H, o_b2 = H + F * H, o_W2 = o_b2 + H, o_b3 = o_W2 + H * H, o_W3 = o_b3 + 1;
#pragma unroll
  for (int kk = 0; kk < KPT; ++kk)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      slab[o_W2 + perm_col(rg * KPT + kk, CG) * H + perm_col(4 * cg + jj, CG)] = gW2[kk][jj];
  {
    const int p = tid % H, fg = tid / H;
#pragma unroll
    for (int ff = 0; ff < MAXFF; ++ff) {
      const int f = fg + FG * ff;
      if (f < F) slab[o_W1 + f * H + perm_col(p, CG)] = gW1[ff];
    }
    if (fg == 0) slab[o_b1 + perm_col(p, CG)] = gb1;
  }
  // column partials (over the thread's own rows) -> shared reduction over row groups
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    atomicAdd(&redS[4 * cg + jj], gb2[jj]);
    atomicAdd(&redS[H + 4 * cg + jj], gW3[jj]);
  }
  if (cg == 0) atomicAdd(&redS[2 * H], gb3);
  __syncthreads();
  for (int q = tid; q < H; q += NT) {
    slab[o_b2 + perm_col(q, CG)] = redS[q];
    slab[o_W3 + perm_col(q, CG)] = redS[H + q];
  }
  if (tid == 0) slab[o_b3] = redS[2 * H];
}

// reduce per-CTA slabs of every network into the fp64 output (theta order)
__global__ void __launch_bounds__(256) orb_bwd_reduce_kernel(const DevSys s, const float* __restrict__ part, const int n_chunks,
                                                             const int p_mlp, double* __restrict__ out) {
  // block = 32 consecutive slab entries x 8 chunk groups: thread (g, lane) sums the slabs c = g, g+8, ... of entry
  // 32 blockIdx.x + lane (coalesced, all loads of a thread independent), then the 8 partial sums are added in a
  // fixed order (deterministic) through shared memory.
  __shared__ double sm[8][33];
  const int m = blockIdx.y;
  const MlpOff off = s.mlp[m];
  const int H = s.H, F = s.F;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;
  double a = 0.0;
  if (idx < p_mlp) {
    double a0 = 0.0, a1 = 0.0;
    int c = g;
    for (; c + 8 < n_chunks; c += 16) {
      a0 += (double)part[(int64_t)(c * s.n_mlp + m) * p_mlp + idx];
      a1 += (double)part[(int64_t)((c + 8) * s.n_mlp + m) * p_mlp + idx];
    }
    if (c < n_chunks) a0 += (double)part[(int64_t)(c * s.n_mlp + m) * p_mlp + idx];
    a = a0 + a1;
  }
  sm[g][lane] = a;
  __syncthreads();
  if (g == 0 && idx < p_mlp) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += sm[q][lane];
    // slab order: b1[H] W1[F*H] b2[H] W2[H*H] b3 W3[H]
    int o;
    if (idx < H) o = off.b1 + idx;
    else if (idx < H + F * H) o = off.W1 + (idx - H);
    else if (idx < 2 * H + F * H) o = off.b2 + (idx - H - F * H);
    else if (idx < 2 * H + F * H + H * H) o = off.W2 + (idx - 2 * H - F * H);
    else if (idx == 2 * H + F * H + H * H) o = off.b3;
    else o = off.W3 + (idx - (2 * H + F * H + H * H + 1));
    out[o] = acc;
  }
}

template <int H, int C, int RPT, int NT>
int launch_eval_t(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st) {
  using K = Cfg<H, C, RPT, NT>;
  const DevSys& s = ctx->sys;
  size_t smem = sizeof(float) * ((size_t)s.F * H + (size_t)H * H + 3 * H + (size_t)H * K::SJ);
  auto kern = orb_eval_kernel<H, C, RPT, NT>;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int occ = 1;
  NQMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
  if (occ < 1) {
    ctx->err = "orb_eval: configuration does not fit in shared memory";
    return NQMC_ERR_UNSUPPORTED;
  }
  const int64_t nwb = (W + K::TR - 1) / K::TR;
  int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  int n_chunks = (ctx->sm_count * occ) / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  {
    ProfScope ps(ctx, C >= 5 ? NQ_PROF_EVAL5 : NQ_PROF_EVAL1, st);
    kern<<<n_chunks * s.n_mlp, NT, smem, st>>>(s, ctx->theta, r, W, phi, n_chunks);
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

template <int H, int NT>
int launch_bwd_t(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                 double* out, cudaStream_t st) {
  using K = BCfg<H, NT>;
  const DevSys& s = ctx->sys;
  size_t smem = sizeof(float) * K::smem_floats(s.F);
  auto kern = orb_bwd_kernel<H, NT>;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (smem > 227 * 1024) {
    ctx->err = "orb_bwd: configuration does not fit in shared memory";
    return NQMC_ERR_UNSUPPORTED;
  }
  int occ = 1;
  NQMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
  if (occ < 1) occ = 1;
  if (occ > 2) occ = 2;
  int n_chunks = (ctx->sm_count * occ) / s.n_mlp;
  if (n_chunks * s.n_mlp > ctx->n_cta_bwd) n_chunks = ctx->n_cta_bwd / s.n_mlp;
  const int64_t nwb = (W + K::TR - 1) / K::TR;
  int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  if (n_chunks > max_items) n_chunks = (int)max_items;
  if (n_chunks < 1) n_chunks = 1;
  {
    ProfScope ps(ctx, NQ_PROF_BWD, st);
    kern<<<n_chunks * s.n_mlp, NT, smem, st>>>(s, ctx->theta, r, ctx->inv, wgt, hdr, e_l, W, ctx->part_mlp, n_chunks,
                                               ctx->p_mlp);
  }
  NQMC_LAUNCH_CHECK();
  dim3 g((ctx->p_mlp + 31) / 32, s.n_mlp);
  orb_bwd_reduce_kernel<<<g, 256, 0, st>>>(s, ctx->part_mlp, n_chunks, ctx->p_mlp, out);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

}  // namespace

int nq_launch_orb_eval(nqmc_ctx* ctx, const float* r, int64_t W, int channels, float* phi, cudaStream_t st) {
  const int H = ctx->sys.H;
  if (channels == 1) {
    if (ctx->use_tc) {
      int rc = nq_launch_orb_eval1_l1tc(ctx, r, W, phi, st);                            // both layers on tensor cores (A <= 6)
      if (rc == NQMC_ERR_UNSUPPORTED) rc = nq_launch_orb_eval1_tc(ctx, r, W, phi, st);  // layer 1 on the FP32 pipe
      if (rc == NQMC_ERR_UNSUPPORTED) rc = nq_launch_orb_fwd_pc(ctx, r, W, 1, phi, nullptr, st);  // hidden width 128
      if (rc != NQMC_ERR_UNSUPPORTED) return rc;
    }
    if (H == 32) return launch_eval_t<32, 1, 8, 256>(ctx, r, W, phi, st);
    if (H == 64) return launch_eval_t<64, 1, 8, 256>(ctx, r, W, phi, st);
    if (H == 128) return launch_eval_t<128, 1, 8, 256>(ctx, r, W, phi, st);
  } else if (channels == 5) {
    if (ctx->use_tc) {
      int rc = nq_launch_orb_lap_pc(ctx, r, W, phi, st);         // tcgen05 producer/consumer pipeline, width 64
      if (rc == NQMC_ERR_UNSUPPORTED) rc = nq_launch_orb_fwd_pc(ctx, r, W, 5, phi, nullptr, st);  // width 128: two N = 64 passes
      if (rc != NQMC_ERR_UNSUPPORTED) return rc;
    }
    if (H == 32) return launch_eval_t<32, 5, 4, 256>(ctx, r, W, phi, st);
    if (H == 64) return launch_eval_t<64, 5, 4, 256>(ctx, r, W, phi, st);
    if (H == 128) return launch_eval_t<128, 5, 4, 256>(ctx, r, W, phi, st);
  }
  else if (channels == 10) {
    if (H == 32) return launch_eval_t<32, 10, 2, 256>(ctx, r, W, phi, st);
    if (H == 64) return launch_eval_t<64, 10, 2, 256>(ctx, r, W, phi, st);
    if (H == 128) return launch_eval_t<128, 10, 2, 256>(ctx, r, W, phi, st);
  }
  ctx->err = "orb_eval: unsupported hidden width / channel count";
  return NQMC_ERR_UNSUPPORTED;
}

int nq_launch_orb_bwd(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr_center, const float* e_l,
                      int64_t W, double* out, cudaStream_t st) {
  const int H = ctx->sys.H;
  if (ctx->use_tc) {
    int nc = 0;
    int rc = nq_launch_orb_bwd_tc(ctx, r, wgt, hdr_center, e_l, W, &nc, st);                  // hidden width 64: one fused kernel
    if (rc == NQMC_ERR_UNSUPPORTED) rc = nq_launch_orb_bwd128(ctx, r, wgt, hdr_center, e_l, W, &nc, st);   // width 128
    if (rc == NQMC_OK) {
      dim3 g((ctx->p_mlp + 31) / 32, ctx->sys.n_mlp);
      orb_bwd_reduce_kernel<<<g, 256, 0, st>>>(ctx->sys, ctx->part_mlp, nc, ctx->p_mlp, out);
      NQMC_LAUNCH_CHECK();
      return NQMC_OK;
    }
    if (rc != NQMC_ERR_UNSUPPORTED) return rc;
  }
  if (H == 32) return launch_bwd_t<32, 256>(ctx, r, wgt, hdr_center, e_l, W, out, st);
  if (H == 64) return launch_bwd_t<64, 256>(ctx, r, wgt, hdr_center, e_l, W, out, st);
  if (H == 128) return launch_bwd_t<128, 256>(ctx, r, wgt, hdr_center, e_l, W, out, st);
  ctx->err = "orb_bwd: unsupported hidden width";
  return NQMC_ERR_UNSUPPORTED;
}
