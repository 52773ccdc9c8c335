// nqmc_gram.cu -- Stochastic Reconfiguration on the device.
// Reference lines replaced: [SPEC] StochasticReconfiguration.compute_update, .github/phase6-issue-body.md:130-182
//   O = log_psi_grads - mean ; E = local_energies - mean ; S = O^T O / n + reg I ; f = 2 E^T O / n ;
//   delta_theta = solve(S, f) ; return -lr delta_theta
//
// The one large dense contraction, S = O^T O (n x P per-sample log-derivatives, P up to ~2e4, n up to ~1e6), is a
// tensor-core SYRK at fp32-grade accuracy:
//   * the caller keeps O parameter-major, Ot[p][w] (the layout a thread-per-walker kernel writes coalesced), so both
//     GEMM operands are K-major (k = walker, contiguous) and TMA (cp.async.bulk.tensor, SWIZZLE_128B, zero fill at the
//     ragged edges) drops 128 x 32 / 256 x 32 fp32 tiles into shared memory in the canonical UMMA layout;
//   * 8 splitter warps turn each raw tile in place into tf32 hi (top 19 bits) + an fp32 remainder lo tile;
//   * one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 256, K = 8): hi*hi + lo*hi + hi*lo into a
//     128 x 256 fp32 accumulator in TMEM (256 columns); a 2-stage ring of mbarriers (raw-full, split-full, empty)
//     keeps TMA, splitters and the tensor pipe overlapped;
//   * only tiles that touch the upper triangle are computed; the epilogue scales, adds reg on the diagonal and writes
//     the element and its mirror image, so S is exactly symmetric.
// Per 32-walker stage a CTA moves 48 KB from L2 for 128 x 256 x 32 x 3 tf32 MACs = 1536 tensor cycles: 31 B/cycle/SM,
// under the ~42 B/cycle/SM the L2 delivers chip-wide (a 128 x 128 tile would sit exactly on that limit).
// The linear solve is Jacobi-preconditioned conjugate gradients on the device (S is SPD once reg > 0): fp64 vectors,
// fp32 S read once per iteration (HBM-bound, P^2 x 4 bytes), no host synchronisation -- a device flag turns the
// remaining iterations into no-ops once the residual target is met.
#include <cmath>

#include "nqmc_internal.cuh"
#include "nqmc_tma.cuh"

namespace {
using namespace nqtma;

constexpr int BM = 128, BN = 256, STAGES = 2;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);          // raw/hi A, raw/hi B, lo A, lo B
constexpr int NSPLIT_WARPS = 8, NSPLIT = NSPLIT_WARPS * 32;
constexpr int NTHREADS = NSPLIT + 64;                          // + TMA producer warp + MMA issuer warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;  // + alignment slack + barriers

// FLUSH: the tensor core adds into its fp32 accumulator with truncation, so a long chain of same-sign terms (the
// diagonal of S: n squares) drifts low by ~(chain length / 8) x 3 x 2^-25 relative -- 7e-4 at n = 65,536.  The chain is
// therefore cut every FLUSH stages (256 samples): the MMAs alternate between two 256-column accumulators, and while
// one fills, the splitter warps add the other into per-thread fp32 running sums (round-to-nearest adds, 128 registers).
constexpr int FLUSH = 8;

__global__ void __launch_bounds__(NTHREADS, 1)
    gram_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const int M, const int N,
                const int64_t K, const int k_per_split, float* __restrict__ C, const int64_t ldc, const float alpha,
                const float diag, const int symmetric, const int atomic_out, const double* __restrict__ inv_alpha_dev) {
  extern __shared__ uint8_t smem_raw[];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (symmetric && n0 + BN <= m0) return;                      // tile entirely below the diagonal
  const int64_t k_begin = (int64_t)blockIdx.z * k_per_split;
  const int64_t k_end = k_begin + k_per_split < K ? k_begin + k_per_split : K;
  const int n_it = (int)((k_end - k_begin + BK - 1) / BK);
  const int n_blk = (n_it + FLUSH - 1) / FLUSH;
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // barriers: raw[ST], split[ST], empty[ST], acc_full[2], acc_free[2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[STAGES + s]), NSPLIT_WARPS);
      mbar_init(smem_u32(&bars[2 * STAGES + s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars[3 * STAGES + b]), 1);
      mbar_init(smem_u32(&bars[3 * STAGES + 2 + b]), NSPLIT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t0 = smem_u32(tiles);
  const uint32_t bar_raw = smem_u32(&bars[0]), bar_split = smem_u32(&bars[STAGES]), bar_empty = smem_u32(&bars[2 * STAGES]);
  const uint32_t bar_full = smem_u32(&bars[3 * STAGES]), bar_free = smem_u32(&bars[3 * STAGES + 2]);

  if (warp == NSPLIT_WARPS) {
    // ================= TMA producer (one elected thread) =================
    if (elect_one()) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES, u = it / STAGES;
        if (u > 0) mbar_wait(bar_empty + 8 * s, (u - 1) & 1);
        const uint32_t st = t0 + s * STAGE_BYTES;
        mbar_arrive_expect_tx(bar_raw + 8 * s, A_BYTES + B_BYTES);
        const int k = (int)(k_begin + (int64_t)it * BK);
        tma_load_3d(st, &tmA, k, m0, 0, bar_raw + 8 * s);
        tma_load_3d(st + A_BYTES, &tmB, k, n0, 0, bar_raw + 8 * s);
      }
    }
  } else if (warp == NSPLIT_WARPS + 1) {
    // ================= MMA issuer (one elected thread) =================
    if (elect_one()) {
      constexpr uint32_t IDESC = idesc_tf32(BM, BN);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % STAGES, u = it / STAGES;
        const int blk = it / FLUSH, b = blk & 1;
        const bool first = it % FLUSH == 0;
        if (first && blk >= 2) mbar_wait(bar_free + 8 * b, ((blk >> 1) - 1) & 1);   // accumulator b has been flushed
        mbar_wait(bar_split + 8 * s, u & 1);
        tc_fence_after();
        const uint32_t st = t0 + s * STAGE_BYTES, acc = tmem_base + (uint32_t)(b * BN);
        const uint64_t ah = desc_k_sw128(st), bh = desc_k_sw128(st + A_BYTES);
        const uint64_t al = desc_k_sw128(st + A_BYTES + B_BYTES), bl = desc_k_sw128(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t o = (uint64_t)(ks * 2);                 // 32 bytes >> 4
          umma_ss(acc, ah + o, bh + o, IDESC, (!first || ks > 0) ? 1u : 0u);
          umma_ss(acc, al + o, bh + o, IDESC, 1u);
          umma_ss(acc, ah + o, bl + o, IDESC, 1u);
        }
        umma_commit(bar_empty + 8 * s);
        if (it % FLUSH == FLUSH - 1 || it == n_it - 1) umma_commit(bar_full + 8 * b);
      }
    }
  } else {
    // ================= splitter warps; they also flush the accumulators and write the tile =================
    const int q = warp & 3, half = warp >> 2;
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(half * (BN / 2));
    float sums[BN / 2];
#pragma unroll
    for (int j = 0; j < BN / 2; ++j) sums[j] = 0.f;
    auto flush = [&](const int blk) {
      const int b = blk & 1;
      mbar_wait(bar_full + 8 * b, (blk >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < BN / 2; c += 16) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)(b * BN + c), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) sums[c + j] += v[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 8 * b);
    };
    int flushed = 0;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES, u = it / STAGES;
      mbar_wait(bar_raw + 8 * s, u & 1);
      uint8_t* st = tiles + s * STAGE_BYTES;
      split_tile_inplace(st, st + A_BYTES + B_BYTES, (A_BYTES + B_BYTES) / 16, tid, NSPLIT);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_split + 8 * s);
      // the ring lets the splitters run at most STAGES stages ahead of the tensor pipe: once stage STAGES of block
      // blk + 1 has arrived, every MMA of block blk has completed and its flush does not stall
      if (it % FLUSH == STAGES && it / FLUSH > flushed) flush(flushed++);
    }
    for (; flushed < n_blk; ++flushed) flush(flushed);
    const int m = m0 + 32 * q + lane;
    const float scale = inv_alpha_dev != nullptr ? (float)(1.0 / fmax(inv_alpha_dev[0], 1.0)) : alpha;
    if (m < M) {
#pragma unroll
      for (int j = 0; j < BN / 2; ++j) {
        const int n = n0 + half * (BN / 2) + j;
        if (n >= N || (symmetric && n < m)) continue;
        float val = scale * sums[j];
        if (m == n && blockIdx.z == 0) val += diag;
        if (atomic_out) {
          atomicAdd(&C[(int64_t)m * ldc + n], val);            // upper triangle only; symmetrize_kernel mirrors it
        } else {
          C[(int64_t)m * ldc + n] = val;
          if (symmetric && n > m) C[(int64_t)n * ldc + m] = val;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 2 * BN);
}

// split-K runs add their partial tiles atomically into the upper triangle; this copies it to the lower one, so the
// result is exactly symmetric whatever order the atomics landed in
__global__ void symmetrize_kernel(float* __restrict__ C, const int64_t ldc, const int N) {
  __shared__ float t[32][33];
  const int bx = blockIdx.x, by = blockIdx.y;
  if (bx < by) return;
  const int x = bx * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int y = by * 32 + j;
    t[j][threadIdx.x] = (x < N && y < N) ? C[(int64_t)y * ldc + x] : 0.f;      // element (row y, col x), x >= y region
  }
  __syncthreads();
  const int xo = by * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int yo = bx * 32 + j;                                                 // writes element (row yo, col xo) = (col, row) of the source
    if (xo < N && yo < N && yo > xo) C[(int64_t)yo * ldc + xo] = t[threadIdx.x][j];
  }
}

// ---- small kernels of the SR pipeline -----------------------------------------------------------------------------------
// stats of e_l: ws[0] = n finite, ws[1] = sum
__global__ void sr_energy_mean_kernel(const float* __restrict__ e_l, const int64_t n, double* __restrict__ sums) {
  double a = 0.0, c = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float e = e_l[i];
    if (isfinite(e)) { a += (double)e; c += 1.0; }
  }
  a = warp_sum(a);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[0], c);
    atomicAdd(&sums[1], a);
  }
}
// one CTA per parameter row p: rs[p] = sum over the finite-energy samples of Ot[p][w]
__global__ void __launch_bounds__(256) sr_rowsum_kernel(const float* __restrict__ Ot, const float* __restrict__ e_l, const int64_t n,
                                                        const int64_t ld, double* __restrict__ rs) {
  __shared__ double red[8];
  const float* row = Ot + (int64_t)blockIdx.x * ld;
  double a = 0.0;
  for (int64_t w = threadIdx.x; w < n; w += 256)
    if (isfinite(e_l[w])) a += (double)row[w];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    rs[blockIdx.x] = t;
  }
}
// centre row p in place with the (global) mean rs[p] / n_finite; samples with a non-finite energy become 0 and drop out
// of S and f; f[p] = 2 / n_finite * sum_w (E_w - Ebar) Oc[p][w]   (this rank's part of it)
__global__ void __launch_bounds__(256) sr_center_kernel(float* __restrict__ Ot, const float* __restrict__ e_l, const int64_t n,
                                                        const int64_t ld, const double* __restrict__ sums,
                                                        const double* __restrict__ rs, double* __restrict__ f) {
  __shared__ double red[8];
  float* row = Ot + (int64_t)blockIdx.x * ld;
  const double nf = fmax(sums[0], 1.0), ebar = sums[1] / nf;
  const float mean = (float)(rs[blockIdx.x] / nf);
  double g = 0.0;
  for (int64_t w = threadIdx.x; w < n; w += 256) {
    const float e = e_l[w];
    float o = 0.f;
    if (isfinite(e)) {
      o = row[w] - mean;
      g += ((double)e - ebar) * (double)o;
    }
    row[w] = o;
  }
  g = warp_sum(g);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = g;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    f[blockIdx.x] = 2.0 * t / nf;
  }
}

// CG state in `ws` (doubles): [0] rz, [1] pq, [2] rz_new, [3] bb (= f.f), [4] converged flag, [5] iterations done,
// [6] rr (unpreconditioned residual norm^2); vectors x, r, p, q, dinv follow (P each).
constexpr int CG_SCAL = 8;
// Deterministic block sum (fixed shuffle / shared-memory order): the CG scalars must come out bit-identical on every rank
// of a multi-GPU run (all ranks then take identical steps), so these kernels run as ONE block of 1024 threads instead of
// accumulating with atomics.  P <= ~2e4: a handful of elements per thread.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
  if (threadIdx.x < 32) {
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}
// d[i] = S[i][i] (this rank's part of the diagonal; summed over ranks before cg_init_kernel)
__global__ void cg_diag_kernel(const float* __restrict__ S, const int64_t lds, const int P, double* __restrict__ ws) {
  double* dinv = ws + CG_SCAL + 4 * (int64_t)P;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) dinv[i] = (double)S[(int64_t)i * lds + i];
}
__global__ void __launch_bounds__(1024) cg_init_kernel(const double* __restrict__ f, const int P, double* __restrict__ ws) {
  __shared__ double red[32];
  double* x = ws + CG_SCAL; double* r = x + P; double* p = r + P; double* dinv = p + 2 * (int64_t)P;
  double rz = 0.0, bb = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const double d = dinv[i];
    const double di = d > 0.0 ? 1.0 / d : 1.0;
    dinv[i] = di;
    x[i] = 0.0;
    r[i] = f[i];
    p[i] = di * f[i];
    rz += f[i] * di * f[i];
    bb += f[i] * f[i];
  }
  rz = block_sum(rz, red);
  bb = block_sum(bb, red);
  if (threadIdx.x == 0) {
    ws[0] = rz;
    ws[3] = bb;
    ws[6] = bb;
  }
}
// q = S p (one warp per row; with several ranks this rank's part of it, summed over ranks before cg_dot_kernel)
__global__ void __launch_bounds__(256) cg_matvec_kernel(const float* __restrict__ S, const int64_t lds, const int P, double* __restrict__ ws) {
  if (ws[4] != 0.0) return;
  const double* p = ws + CG_SCAL + 2 * (int64_t)P;
  double* q = ws + CG_SCAL + 3 * (int64_t)P;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= P) return;
  const float* sr = S + (int64_t)row * lds;
  double a0 = 0.0, a1 = 0.0;
  int j = lane;
  for (; j + 32 < P; j += 64) {
    a0 = fma((double)sr[j], p[j], a0);
    a1 = fma((double)sr[j + 32], p[j + 32], a1);
  }
  if (j < P) a0 = fma((double)sr[j], p[j], a0);
  const double a = warp_sum(a0 + a1);
  if (lane == 0) q[row] = a;
}
// pq = p.q
__global__ void __launch_bounds__(1024) cg_dot_kernel(const int P, double* __restrict__ ws) {
  __shared__ double red[32];
  if (ws[4] != 0.0) return;
  const double* p = ws + CG_SCAL + 2 * (int64_t)P;
  const double* q = p + P;
  double a = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) a += p[i] * q[i];
  a = block_sum(a, red);
  if (threadIdx.x == 0) ws[1] = a;
}
// x += a p ; r -= a q ; z = dinv r ; rz_new += r.z ; rr += r.r
__global__ void __launch_bounds__(1024) cg_update_kernel(const int P, double* __restrict__ ws) {
  __shared__ double red[32];
  if (ws[4] != 0.0) return;
  double* x = ws + CG_SCAL; double* r = x + P; double* p = r + P; double* q = p + P; double* dinv = q + P;
  const double a = ws[1] != 0.0 ? ws[0] / ws[1] : 0.0;
  double rz = 0.0, rr = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    x[i] += a * p[i];
    const double ri = r[i] - a * q[i];
    r[i] = ri;
    rz += ri * dinv[i] * ri;
    rr += ri * ri;
  }
  rz = block_sum(rz, red);
  rr = block_sum(rr, red);
  if (threadIdx.x == 0) {
    ws[2] = rz;
    ws[7] = rr;
  }
}
// p = z + b p ; rotate the scalars ; convergence test on |r| / |f|
__global__ void cg_direction_kernel(const int P, const double tol2, double* __restrict__ ws, double* __restrict__ done_count) {
  if (ws[4] != 0.0) return;
  double* r = ws + CG_SCAL + P; double* p = r + P; double* dinv = p + 2 * (int64_t)P;
  const double b = ws[0] != 0.0 ? ws[2] / ws[0] : 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) p[i] = dinv[i] * r[i] + b * p[i];
  // the last CTA to finish rotates the scalars (every CTA has read them by then)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const double t = atomicAdd(done_count, 1.0);
    if (t == (double)(gridDim.x - 1)) {
      *done_count = 0.0;
      ws[0] = ws[2];
      ws[1] = 0.0;
      ws[2] = 0.0;
      ws[6] = ws[7];
      ws[7] = 0.0;
      ws[5] += 1.0;
      if (ws[6] <= tol2 * ws[3]) ws[4] = 1.0;
      __threadfence();
    }
  }
}
__global__ void sr_finish_kernel(const int P, const float lr, const double* __restrict__ ws, double* __restrict__ delta,
                                 double* __restrict__ info) {
  const double* x = ws + CG_SCAL;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) delta[i] = -(double)lr * x[i];
  if (blockIdx.x == 0 && threadIdx.x == 0 && info != nullptr) {
    info[0] = ws[5];                                        // iterations
    info[1] = ws[3] > 0.0 ? sqrt(ws[6] / ws[3]) : 0.0;      // |r| / |f|
    info[2] = ws[4];                                        // 1: tolerance met
    info[3] = sqrt(ws[3]);                                  // |f|
  }
}

__global__ void apply_update_kernel(float* __restrict__ theta, const double* __restrict__ delta, const int P) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) theta[i] = (float)((double)theta[i] + delta[i]);
}

int launch_gram(nqmc_ctx* ctx, const float* A, const float* B, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                float alpha, float diag, int symmetric, float* C, int64_t ldc, cudaStream_t st,
                const double* inv_alpha_dev = nullptr) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || lda < K || ldb < K || ldc < N) return NQMC_ERR_ARG;
  if (M > INT32_MAX || N > INT32_MAX) return NQMC_ERR_ARG;
  if (symmetric && (M != N)) return NQMC_ERR_ARG;
  CUtensorMap tmA, tmB;
  int ra = make_map_k32(&tmA, A, (uint64_t)K, (uint64_t)M, 1, (uint64_t)lda, 0, BM);
  int rb = make_map_k32(&tmB, B, (uint64_t)K, (uint64_t)N, 1, (uint64_t)ldb, 0, BN);
  if (ra != 0 || rb != 0) {
    ctx->err = "nqmc_gram: cuTensorMapEncodeTiled failed (operands must be 16-byte aligned with leading dimensions that are multiples of 4)";
    return ra == -2 || rb == -2 ? NQMC_ERR_ARG : NQMC_ERR_CUDA;
  }
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int tm = (int)((M + BM - 1) / BM), tn = (int)((N + BN - 1) / BN);
  // split K when there are too few output tiles to fill the GPU (the partial tiles are then added atomically)
  const int64_t k_its = (K + BK - 1) / BK;
  int64_t tiles = (int64_t)tm * tn;
  if (symmetric) tiles = (tiles + 1) / 2 + tn;
  int splits = 1;
  if (tiles < ctx->sm_count) {
    splits = (int)((2 * ctx->sm_count + tiles - 1) / tiles);
    if (splits > k_its / 8) splits = (int)(k_its / 8);
    if (splits < 1) splits = 1;
  }
  const int64_t per_it = (k_its + splits - 1) / splits;
  splits = (int)((k_its + per_it - 1) / per_it);
  if (splits > 1) NQMC_CUDA(cudaMemsetAsync(C, 0, (size_t)M * ldc * sizeof(float), st));
  gram_kernel<<<dim3(tn, tm, splits), NTHREADS, SMEM_BYTES, st>>>(tmA, tmB, (int)M, (int)N, K, (int)(per_it * BK), C, ldc, alpha, diag,
                                                                 symmetric, splits > 1 ? 1 : 0, inv_alpha_dev);
  NQMC_LAUNCH_CHECK();
  if (splits > 1 && symmetric) {
    symmetrize_kernel<<<dim3((unsigned)((N + 31) / 32), (unsigned)((N + 31) / 32)), dim3(32, 8), 0, st>>>(C, ldc, (int)N);
    NQMC_LAUNCH_CHECK();
  }
  return NQMC_OK;
}

}  // namespace

extern "C" {

int nqmc_gram(nqmc_ctx* ctx, const float* A, const float* B, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
              float alpha, float diag, int symmetric, float* C, int64_t ldc, nqmc_stream stream) {
  if (!ctx) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  return launch_gram(ctx, A, B, M, N, K, lda, ldb, alpha, diag, symmetric, C, ldc, reinterpret_cast<cudaStream_t>(stream));
}

int nqmc_apply_update(nqmc_ctx* ctx, const double* delta, int64_t P, nqmc_stream stream) {
  if (!ctx || !delta || P != ctx->sys.P) return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  apply_update_kernel<<<ctx->sm_count, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ctx->theta, delta, (int)P);
  ctx->theta_dirty = true;
  ctx->act_valid_for = nullptr;
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

int64_t nqmc_sr_workspace_bytes(int64_t P) { return P > 0 ? (int64_t)sizeof(double) * (CG_SCAL + 6 * P + 8) : -1; }

int nqmc_sr_update(nqmc_ctx* ctx, float* Ot, const float* e_l, int64_t P, int64_t n, int64_t ld, float reg, float lr,
                   int32_t max_iter, double tol, float* S, double* work, double* delta, double* info, nqmc_stream stream) {
  if (!ctx || !Ot || !e_l || !S || !work || !delta || P <= 0 || n <= 0 || ld < n || max_iter < 1 || !(tol > 0.0) || P > INT32_MAX)
    return NQMC_ERR_ARG;
  NQMC_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* ws = work;                              // CG_SCAL scalars + x r p q dinv
  double* f = work + CG_SCAL + 5 * P;             // P
  double* sums = f + P;                           // 2 (+ done counter at [2])
  // With a peer-memory communicator attached the samples are sharded over the ranks: the sample count, the energy sum,
  // the row sums, f, the diagonal of S and -- every CG iteration -- q = S_r p are summed over ranks in-stream
  // (nqmc_comm.cu, bitwise identical on all ranks, so every rank takes the same CG steps and ends with the same delta).
  // S itself is never reduced: each rank keeps its own P x P part.
  const bool multi = nq_comm_active(ctx);
  double* rs = ws + CG_SCAL + 3 * P;              // row sums live in q's place until CG starts
  NQMC_CUDA(cudaMemsetAsync(ws, 0, CG_SCAL * sizeof(double), st));
  NQMC_CUDA(cudaMemsetAsync(sums, 0, 8 * sizeof(double), st));
  sr_energy_mean_kernel<<<ctx->sm_count, 256, 0, st>>>(e_l, n, sums);
  NQMC_LAUNCH_CHECK();
  int rc;
  if (multi && (rc = nq_comm_allreduce(ctx, sums, 8, st))) return rc;
  sr_rowsum_kernel<<<(unsigned)P, 256, 0, st>>>(Ot, e_l, n, ld, rs);
  NQMC_LAUNCH_CHECK();
  if (multi && (rc = nq_comm_allreduce(ctx, rs, P, st))) return rc;
  sr_center_kernel<<<(unsigned)P, 256, 0, st>>>(Ot, e_l, n, ld, sums, rs, f);
  NQMC_LAUNCH_CHECK();
  if (multi && (rc = nq_comm_allreduce(ctx, f, P, st))) return rc;
  // S_r = Oc Oc^T / n_finite (+ reg I on rank 0 only, so that the sum over ranks carries it once): the epilogue reads
  // the global n_finite from the device (sums[0]), so nothing synchronises; the centred rows of non-finite samples are
  // zero and drop out of the sum.
  const int rank0 = !multi || nq_comm_rank(ctx) == 0;
  rc = launch_gram(ctx, Ot, Ot, P, P, n, ld, ld, 1.0f / (float)n, rank0 ? reg : 0.0f, 1, S, P, st, sums);
  if (rc != NQMC_OK) return rc;
  const int g = ctx->sm_count;
  cg_diag_kernel<<<g, 256, 0, st>>>(S, P, (int)P, ws);
  NQMC_LAUNCH_CHECK();
  if (multi && (rc = nq_comm_allreduce(ctx, ws + CG_SCAL + 4 * P, P, st))) return rc;
  cg_init_kernel<<<1, 1024, 0, st>>>(f, (int)P, ws);
  NQMC_LAUNCH_CHECK();
  const double tol2 = tol * tol;
  for (int it = 0; it < max_iter; ++it) {
    cg_matvec_kernel<<<(unsigned)((P + 7) / 8), 256, 0, st>>>(S, P, (int)P, ws);
    if (multi && (rc = nq_comm_allreduce(ctx, ws + CG_SCAL + 3 * P, P, st))) return rc;
    cg_dot_kernel<<<1, 1024, 0, st>>>((int)P, ws);
    cg_update_kernel<<<1, 1024, 0, st>>>((int)P, ws);
    cg_direction_kernel<<<g, 256, 0, st>>>((int)P, tol2, ws, sums + 2);
  }
  ctx->launches += 4 * (int64_t)max_iter;
  sr_finish_kernel<<<g, 256, 0, st>>>((int)P, lr, ws, delta, info);
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

}  // extern "C"
