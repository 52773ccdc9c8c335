/* nqmc_b200.h -- C ABI of the B200-native nqmc walker loop (libnqmc_b200.so).
 *
 * The reference (Sakeeb91/neural-qmc-wavefunction) has no FFI / plugin registry: its extension
 * points are Python call signatures that drive a per-walker closure through jax.vmap / grad /
 * hessian.  This header is the coarse, batched boundary those signatures bind to instead
 * (SURVEY section 8(b)); every entry point cites the reference lines it replaces.  All paths are
 * relative to /root/reference.
 *
 * Conventions
 *   - plain C, no C++ / torch / XLA types.  `nqmc_stream` is a cudaStream_t passed as void*.
 *   - every buffer is CALLER-OWNED.  Unless a name ends in `_host`, pointers are DEVICE pointers on
 *     the ctx's device.  Calls are asynchronous on `stream`, never synchronise, never allocate
 *     (all scratch is sized by nqmc_reserve).  `_host` entry points take HOST pointers, copy in,
 *     run, copy out and synchronise the stream before returning (the e2e / ctypes convenience path).
 *   - walker positions are structure-of-arrays fp32: r[(3*i + d) * W + w] for electron i, axis d,
 *     walker w ("SoA [3N][W]").  The reference's AoS (W,N,3) row-major layout
 *     (src/nqmc/sampler/metropolis.py:33) is converted at the edge with nqmc_aos_to_soa / _soa_to_aos.
 *   - theta is ONE flat fp32 vector: the reference's parameter pytree flattened leaf by leaf in
 *     sorted-key order, row-major leaves == jax.flatten_util.ravel_pytree(params)
 *     (jax.tree.leaves order, used by the reference at src/nqmc/optimizer/vmc.py:265).
 *     nqmc_param_count / nqmc_leaf_info describe it.  Gradients come back in the same order.
 *   - return value: 0 ok, <0 error (codes below); message via nqmc_last_error.  No exceptions
 *     cross the ABI.  A ctx is bound to one device, is not thread-safe; distinct ctxs are independent.
 *   - there is NO CPU fallback: nqmc_create fails with NQMC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef NQMC_B200_H
#define NQMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NQMC_ABI_VERSION 3

typedef struct nqmc_ctx nqmc_ctx;
typedef void* nqmc_stream; /* cudaStream_t */

enum { NQMC_OK = 0, NQMC_ERR_ARG = -1, NQMC_ERR_CUDA = -2, NQMC_ERR_UNSUPPORTED = -3, NQMC_ERR_STATE = -4 };

/* wavefunction kinds */
enum {
  NQMC_KIND_SIMPLE = 0,        /* SimpleWavefunction: src/nqmc/wavefunction/simple.py:73-164 */
  NQMC_KIND_ANTISYMMETRIC = 1  /* [SPEC] AntisymmetricWavefunction: .github/phase3-issue-body.md:180-219 */
};

/* Geometry constants the path reads from the reference's Molecule
 * (src/nqmc/systems/molecule.py:49-97: positions, charges, spin; :142-153 V_nn). */
typedef struct {
  int32_t n_atoms;  /* 1..16 */
  int32_t n_elec;   /* 1..16 */
  int32_t n_up;     /* first n_up electrons are spin-up (molecule.py:177-180) */
  int32_t _pad;
  const double* R;  /* host [n_atoms][3], bohr */
  const double* Z;  /* host [n_atoms] */
  double v_nn;      /* Molecule.nuclear_repulsion_energy() */
} nqmc_system;

/* Network shape (constructor kwargs of the reference's wavefunction classes). */
typedef struct {
  int32_t kind;          /* NQMC_KIND_* */
  int32_t hidden0;       /* hidden_dims[0]; kind 1: 32, 64 or 128 */
  int32_t hidden1;       /* hidden_dims[1]; kind 1: must equal hidden0; kind 0: see n_hidden */
  int32_t use_cusp;      /* kind 1: ElectronNuclearCusp + ElectronElectronCusp (phase3-issue-body.md:50-122) */
  int32_t use_backflow;  /* kind 1: BackflowTransform (phase3-issue-body.md:134-169) */
  int32_t _pad;
  double envelope_decay; /* kind 0: SimpleWavefunction.envelope_decay (simple.py:102) */
  double cusp_same;      /* e-e cusp coefficient, same spin  (spec: 0.5, phase3-issue-body.md:109) */
  double cusp_diff;      /* e-e cusp coefficient, opposite   (spec: 0.25, :110) */
  /* kind 0 only (ABI v2): SimpleMLP accepts any tuple of hidden widths and two activations
   * (src/nqmc/wavefunction/simple.py:41,56-58,63-65).  n_hidden == 0 means "use hidden0, hidden1". */
  int32_t n_hidden;      /* 0, or 1..6 entries of hidden[] in use */
  int32_t activation;    /* NQMC_ACT_TANH / NQMC_ACT_SILU */
  int32_t hidden[6];     /* hidden_dims, each 1..128 */
} nqmc_net;

enum { NQMC_ACT_TANH = 0, NQMC_ACT_SILU = 1 };

/* ---- lifecycle ------------------------------------------------------------------------------- */

/* Replaces: wavefunction / Molecule construction (scripts/train_h2.py:155-169). */
int nqmc_create(const nqmc_system* sys, const nqmc_net* net, int device, nqmc_ctx** out);
int nqmc_destroy(nqmc_ctx* ctx);
const char* nqmc_last_error(const nqmc_ctx* ctx); /* ctx may be NULL: last create() error */
int nqmc_abi_version(void);

/* Size all internal scratch for batches of up to max_walkers walkers (the only allocating call). */
int nqmc_reserve(nqmc_ctx* ctx, int64_t max_walkers);

/* ---- parameters ------------------------------------------------------------------------------ */

/* Number of scalars P in theta (== sum of leaf sizes of wf.init(...), simple.py:166-176). */
int64_t nqmc_param_count(const nqmc_ctx* ctx);
/* Leaf `index` (0-based, sorted-key order): writes its '/'-joined path (NUL-terminated, truncated to
 * cap), rank-2 shape (rows, cols; scalar leaves are 1 x 1) and offset into theta.
 * Returns the number of leaves when index < 0. */
int64_t nqmc_leaf_info(const nqmc_ctx* ctx, int index, char* path, size_t cap, int64_t* rows, int64_t* cols,
                       int64_t* offset);
/* Upload theta (P floats).  `theta` is a device pointer; `theta_host` variant takes host memory. */
int nqmc_set_params(nqmc_ctx* ctx, const float* theta, int64_t P, nqmc_stream stream);
int nqmc_set_params_host(nqmc_ctx* ctx, const float* theta_host, int64_t P, nqmc_stream stream);

/* Read theta back (device-to-device / device-to-host). */
int nqmc_get_params(nqmc_ctx* ctx, float* theta, int64_t P, nqmc_stream stream);
int nqmc_get_params_host(nqmc_ctx* ctx, float* theta_host, int64_t P, nqmc_stream stream);
/* Fused optimizer epilogue on the device: gradient = 2 gsum / hdr[N] (float32), global-norm clip,
 * Adam (b1 0.9, b2 0.999, eps 1e-8), theta updated IN PLACE in the ctx.
 * Replaces: optax.chain(clip_by_global_norm(clip), adam(lr)).update + apply_updates -- src/nqmc/optimizer/vmc.py:180-184,256-261
 * and the grad_norm metric (:264-266).  m, v: device [P] fp32 moments, caller-owned, zero before the first step;
 * count: 1-based step number; grad_norm: device double (pre-clip norm), may be NULL. */
int nqmc_adam_step(nqmc_ctx* ctx, const double* gsum, const double* hdr, float* m, float* v, int64_t count, float lr,
                   float clip, double* grad_norm, nqmc_stream stream);

/* ---- layout helpers -------------------------------------------------------------------------- */
int nqmc_aos_to_soa(nqmc_ctx* ctx, const float* aos, float* soa, int64_t W, nqmc_stream stream);
int nqmc_soa_to_aos(nqmc_ctx* ctx, const float* soa, float* aos, int64_t W, nqmc_stream stream);

/* ---- the hot path ---------------------------------------------------------------------------- */

/* log|psi| and sign for W walkers.
 * Replaces: vmap(wavefunction(params, r)) -- src/nqmc/wavefunction/base.py:38-48, simple.py:151-164,
 * [SPEC] phase3-issue-body.md:196-219.  `sign` may be NULL. */
int nqmc_log_psi(nqmc_ctx* ctx, const float* r, int64_t W, float* logabs, float* sign, nqmc_stream stream);

/* One all-electron Metropolis-Hastings move for every walker, in place.
 * Replaces: MetropolisSampler.step -- src/nqmc/sampler/metropolis.py:115-147
 *   proposals = r + sigma * normal ; accept = log(u) < 2 log|psi'| - logp  (strict, fp32).
 * normals : SoA [3N][W] N(0,1) draws, or NULL -> on-device Philox4x32-10 keyed by (seed, step,
 *           walker_id0 + w) (parity mode supplies the reference's streams)
 * uniforms: [W] U[0,1) draws, or NULL (same Philox stream)
 * logp    : [W] current log|psi|^2, updated where accepted (carries the reference's stale-logp
 *           behaviour across parameter updates, SURVEY A.5, because it is caller state)
 * accept  : [W] 0/1 per walker, may be NULL ; n_accepted: [W] fp32 += accept, may be NULL (:140) */
int nqmc_mh_step(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms, uint64_t seed,
                 uint64_t step, int64_t walker_id0, float sigma, int64_t W, uint8_t* accept, float* n_accepted,
                 nqmc_stream stream);

/* MetropolisSampler.sample on the device: n_burn moves, counters reset, n_samples * n_thin moves of which every
 * n_thin-th position (starting with the first) is stored.  Walkers never leave the GPU and discarded positions
 * are never materialised.
 * Replaces: src/nqmc/sampler/metropolis.py:174-218 (two lax.scan loops, all_positions[::n_thin]).
 * normals / uniforms: NULL (device Philox keyed by seed_burn / seed_sample, step = move index within its phase),
 *   or explicit streams SoA [n_burn + n_samples*n_thin][3N][W] / [n_burn + n_samples*n_thin][W] (parity mode).
 * samples   : SoA [3N][n_samples][W] (flat walker index s * W + w == the reference's reshape(-1, N, 3), vmc.py:105)
 * n_accepted: [W] or NULL; on return the per-walker accept counts of the sampling phase (:194-199,216)
 * sigma_dev : NULL, or a device float holding the proposal width (overrides `sigma`).  With adapt_every > 0 it is
 *   adapted during burn-in only, every adapt_every moves: sigma *= exp((acc - target_acceptance) / sqrt(k)) for
 *   the k-th block -- the behaviour the reference's dataclass advertises (target_acceptance, adapt_step,
 *   metropolis.py:52-54) but never implements.  Sampling-phase moves use the final width, so detailed balance holds. */
int nqmc_sample(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms, uint64_t seed_burn,
                uint64_t seed_sample, int64_t walker_id0, float sigma, int64_t W, int64_t n_burn, int64_t n_samples,
                int64_t n_thin, float* samples, float* n_accepted, float* sigma_dev, float target_acceptance,
                int64_t adapt_every, nqmc_stream stream);

/* Local energy E_L = T + V_en + V_ee + V_nn with the analytic forward-mode Laplacian.
 * Replaces: local_energy_batched / kinetic_energy / electron_nuclear_potential /
 * electron_electron_potential -- src/nqmc/hamiltonian/molecular.py:52-69,87-102,116-135,159-197.
 * parts (may be NULL): SoA [4][W] = T, V_en, V_ee, log|psi|. */
int nqmc_local_energy(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, nqmc_stream stream);

/* out[p] = sum_w weight[w] * d log|psi(r_w)| / d theta_p      (fp64 result, P entries)
 * Replaces: vmap(jax.grad(lambda p: wavefunction(p, r)))(samples) followed by the weighted mean --
 * src/nqmc/optimizer/vmc.py:120-133 -- without materialising the (n_total x P) matrix.
 * weight NULL -> all ones (gives sum_w O_w). */
int nqmc_grad_accum(nqmc_ctx* ctx, const float* r, const float* weight, int64_t W, double* out, nqmc_stream stream);

/* Reduction header written by nqmc_energy_stats / nqmc_walker_step (8 doubles, all additive across ranks). */
enum {
  NQMC_HDR_N = 0,        /* walkers with finite E_L */
  NQMC_HDR_SUM_E = 1,    /* sum E_L over those */
  NQMC_HDR_SUM_E2 = 2,   /* sum E_L^2 */
  NQMC_HDR_N_ACCEPT = 3, /* accepted moves in this step */
  NQMC_HDR_N_BAD = 4,    /* walkers whose E_L was not finite (excluded from the sums) */
  NQMC_HDR_SIZE = 8
};

/* hdr[0..7] from e_l[W] (N_ACCEPT left 0).  Replaces jnp.mean / jnp.std in vmc.py:113-114. */
int nqmc_energy_stats(nqmc_ctx* ctx, const float* e_l, int64_t W, double* hdr, nqmc_stream stream);

/* Blocked statistics for honest error bars: e_l holds n_samples consecutive local energies of
 * each of W independent chains, sample-major (e_l[s * W + w], the layout nqmc_sample + nqmc_local_energy produce).
 * Every chain is cut into n_blocks blocks of n_samples / n_blocks samples; out3 = {number of finite block means,
 * their sum, the sum of their squares}: mean = out3[1]/out3[0], std_error = sqrt(var / out3[0]).
 * Batched form of the [SPEC] blocking_analysis, .github/phase5-issue-body.md:113-132; n_blocks == n_samples gives the
 * reference's naive jnp.std / sqrt(n) (src/nqmc/optimizer/vmc.py:114), which ignores autocorrelation. */
int nqmc_blocked_stats(nqmc_ctx* ctx, const float* e_l, int64_t n_samples, int64_t W, int32_t n_blocks, double* out3,
                       nqmc_stream stream);

/* The fused walker step = one sample of VMCOptimizer.step's inner work (vmc.py:242-253):
 *   phase A (nqmc_walker_step_a): MH move (as nqmc_mh_step) -> E_L at the new position -> hdr.
 *   [multi-GPU: the caller all-reduces hdr (8 doubles) here -- the only exchange before phase B]
 *   phase B (nqmc_walker_step_b): gsum[p] = sum_w (E_w - hdr.SUM_E/hdr.N) * O_w[p]   (fp64, P entries)
 *   [multi-GPU: the caller all-reduces gsum];  gradient = 2 * gsum / hdr.N  (vmc.py:126-133).
 * nqmc_walker_step runs A then B back to back (single GPU). */
int nqmc_walker_step_a(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms,
                       uint64_t seed, uint64_t step, int64_t walker_id0, float sigma, int64_t W, float* e_l,
                       float* n_accepted, double* hdr, nqmc_stream stream);
int nqmc_walker_step_b(nqmc_ctx* ctx, const float* r, const float* e_l, int64_t W, const double* hdr, double* gsum,
                       nqmc_stream stream);
int nqmc_walker_step(nqmc_ctx* ctx, float* r, float* logp, const float* normals, const float* uniforms, uint64_t seed,
                     uint64_t step, int64_t walker_id0, float sigma, int64_t W, float* e_l, float* n_accepted,
                     double* hdr, double* gsum, nqmc_stream stream);

/* Host-buffer variant of nqmc_walker_step: r_aos_host is the reference's (W,N,3) AoS array; copies
 * H2D, converts, steps, converts back, copies r / logp / e_l / hdr / gsum D2H, synchronises.
 * normals_aos_host (W,N,3) / uniforms_host (W) may be NULL (device Philox). */
int nqmc_walker_step_host(nqmc_ctx* ctx, float* r_aos_host, float* logp_host, const float* normals_aos_host,
                          const float* uniforms_host, uint64_t seed, uint64_t step, int64_t walker_id0, float sigma,
                          int64_t W, float* e_l_host, double* hdr_host, double* gsum_host, nqmc_stream stream);

/* Pipelined form of the host-buffer step, for a host loop that owns the walkers (the reference's VMCOptimizer.step /
 * MetropolisSampler.sample callers, src/nqmc/optimizer/vmc.py:238-253): the walkers, log|psi|^2, E_L and the header are
 * final after phase A and travel home while the reverse pass (phase B) still runs, and the next step's walkers -- which
 * do not depend on the gradient -- can travel to the device during that reverse pass.
 *   nqmc_walker_step_host_async : nqmc_walker_step_host without the final synchronisation.
 *   nqmc_walker_step_host_wait  : block until the phase-A outputs (r, logp, e_l, hdr: NQMC_WAIT_PHASE_A) or everything
 *                                 (gsum too: NQMC_WAIT_ALL) of the last _async call is in the host buffers.
 *   nqmc_walker_step_host_upload: start the H2D copy + AoS->SoA conversion of the NEXT step's inputs on the context's
 *                                 own copy stream.  Call it after nqmc_walker_step_host_wait(ctx, NQMC_WAIT_PHASE_A);
 *                                 the next _async / _host call that names the same host pointers and W finds its
 *                                 inputs staged and skips its own copy (any other call copies as usual).
 * The host buffers must stay valid (pinned for real overlap) until the matching wait returns. */
#define NQMC_WAIT_PHASE_A 0
#define NQMC_WAIT_ALL 1
int nqmc_walker_step_host_async(nqmc_ctx* ctx, float* r_aos_host, float* logp_host, const float* normals_aos_host,
                                const float* uniforms_host, uint64_t seed, uint64_t step, int64_t walker_id0, float sigma,
                                int64_t W, float* e_l_host, double* hdr_host, double* gsum_host, nqmc_stream stream);
int nqmc_walker_step_host_wait(nqmc_ctx* ctx, int what);
int nqmc_walker_step_host_upload(nqmc_ctx* ctx, const float* r_aos_host, const float* logp_host,
                                 const float* normals_aos_host, const float* uniforms_host, int64_t W);

/* ---- Stochastic Reconfiguration ----------------------------------------------------------
 * C[M][N] = alpha * A B^T + diag * I at fp32-grade accuracy on the tensor cores (tcgen05 kind::tf32, three products
 * per fp32 product, fp32 accumulation in TMEM), operands streamed by TMA.  A is [M][lda], B is [N][ldb], both with the
 * contraction index k contiguous (k < K <= lda, ldb); C is [M][ldc].  Operands must be 16-byte aligned with leading
 * dimensions that are multiples of 4.  symmetric != 0 (B == A, M == N): only tiles touching the upper triangle are
 * computed and every element is mirrored, so C is exactly symmetric.
 * Replaces: S = jnp.einsum('ni,nj->ij', O, O) -- [SPEC] .github/phase6-issue-body.md:168 -- with O kept
 * parameter-major (A = B = O^T, k = sample). */
int nqmc_gram(nqmc_ctx* ctx, const float* A, const float* B, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
              float alpha, float diag, int symmetric, float* C, int64_t ldc, nqmc_stream stream);
/* One SR update on the device.
 * Replaces: [SPEC] StochasticReconfiguration.compute_update -- .github/phase6-issue-body.md:150-181:
 *   O = log_psi_grads - mean ; E = local_energies - mean ; S = O^T O / n + reg I ; f = 2 E^T O / n ;
 *   delta = -lr * solve(S, f).
 * Ot   : [P][ld] per-sample log-derivatives, PARAMETER-major (Ot[p][w] = d log|psi(r_w)| / d theta_p), ld >= n,
 *        ld % 4 == 0; centred IN PLACE (samples whose local energy is not finite are zeroed and leave every sum)
 * e_l  : [n] local energies ; S: [P][P] fp32 workspace, on return the regularised S ; work: nqmc_sr_workspace_bytes(P)
 * delta: [P] fp64 update direction (-lr S^-1 f) ; info (may be NULL): [4] = CG iterations, |r|/|f|, converged, |f|
 * The solve is Jacobi-preconditioned CG (S is SPD for reg > 0) run for at most max_iter iterations or until
 * |S x - f| <= tol |f|, entirely in-stream (no host synchronisation).
 * Multi-GPU: with a communicator attached (nqmc_comm_attach) Ot / e_l hold THIS rank's samples; the sample count, the
 * energy and row sums, f, diag(S) and -- every CG iteration -- S_r p are summed over ranks in-stream (P <= the
 * communicator's vector length, i.e. the ctx's parameter count); every rank returns the bit-identical delta.  All ranks
 * must make the call with the same P, max_iter and tol. */
int64_t nqmc_sr_workspace_bytes(int64_t P);
/* Per-walker log-derivatives, parameter-major: Ot[p * ld + w] = d log|psi(r_w)| / d theta_p, ld >= W.
 * Replaces: vmap(jax.grad(lambda p: wavefunction(p, r)))(samples) -- src/nqmc/optimizer/vmc.py:120-123 -- in the one
 * case where the matrix is wanted itself (the O of Stochastic Reconfiguration).  Every ansatz: SimpleWavefunction
 * (nqmc_simple.cu) and the antisymmetric one with or without cusps / backflow (nqmc_pergrad.cu; NQMC_ERR_UNSUPPORTED only
 * when electrons-per-spin x hidden width exceed its shared-memory staging, e.g. H10 at width 128 -- whose S would be
 * 155 GB anyway).  Ot: [P][ld], ld >= W; W <= the reserved walker count.  H4 / 64-wide, 65,536 samples: 5.2 GB in 5.3 ms. */
int nqmc_log_psi_grads(nqmc_ctx* ctx, const float* r, int64_t W, float* Ot, int64_t ld, nqmc_stream stream);
/* theta += delta (fp64 update direction from nqmc_sr_update) in the ctx, on the device. */
int nqmc_apply_update(nqmc_ctx* ctx, const double* delta, int64_t P, nqmc_stream stream);
int nqmc_sr_update(nqmc_ctx* ctx, float* Ot, const float* e_l, int64_t P, int64_t n, int64_t ld, float reg, float lr,
                   int32_t max_iter, double tol, float* S, double* work, double* delta, double* info, nqmc_stream stream);

/* ---- debugging ------------------------------------------------------------------------------
 * With NQMC_GUARD=1 in the environment when nqmc_reserve runs, every internal scratch buffer is allocated with a 4 KB
 * guard band on either side.  nqmc_debug_check_guards synchronises the device and returns the number of buffers whose
 * bands were overwritten (0: none; nqmc_last_error names them), or 0 when the guard mode is off. */
int nqmc_debug_check_guards(nqmc_ctx* ctx);
/* Guard mode only: device pointer of the named internal scratch buffer ("phi", "inv", "logabs", ...), else NULL.
 * For tests of the checker itself. */
void* nqmc_debug_scratch_ptr(nqmc_ctx* ctx, const char* name);
/* Host-side run of the work-item iterator of the persistent orbital kernels (csrc/nqmc_internal.cuh, ItemIter): writes
 * the (walker block, electron-of-spin) pairs of items first, first + stride, ... (n of them); no GPU needed.  Tests
 * compare it with item / ns, item % ns. */
int nqmc_debug_item_iter(int32_t first, int32_t stride, int32_t ns, int64_t n, int32_t* walker_block, int32_t* electron);

/* ---- introspection --------------------------------------------------------------------------- */
/* Kernels launched by this ctx since creation (bench.py's gpu_launches). */
int64_t nqmc_launch_count(const nqmc_ctx* ctx);
/* Select the tcgen05 / TMEM 3xTF32 variant of the hidden-layer GEMMs where one exists (default on; the
 * environment variable NQMC_TC=0 turns it off at create time).  The FP32 FFMA kernels are the parity anchor. */
int nqmc_use_tensor_cores(nqmc_ctx* ctx, int enable);
/* Per-kernel CUDA-event timing on the launching stream (used by bench.py for the roofline leg).
 * kind: 0 orbital forward (value), 1 orbital forward (value+grad+Laplacian), 2 orbital reverse pass,
 *       3 Slater/Jastrow/E_L assembly, 4 accept.  nqmc_profile(ctx, 1) resets and enables; _read
 *       synchronises the device and returns the summed duration and the number of launches recorded. */
int nqmc_profile(nqmc_ctx* ctx, int enable);
int nqmc_profile_read(nqmc_ctx* ctx, int kind, double* total_ms, int64_t* count);
/* Measured FP32 FMA throughput of this GPU in TFLOP/s (register-resident FFMA chains, best of reps):
 * the roofline denominator of the FFMA-bound kernels (MEASURED_PEAKS.json has no fp32 figure). */
int nqmc_ffma_peak(nqmc_ctx* ctx, int reps, double* tflops, nqmc_stream stream);
/* Fill normals SoA [3N][W] / uniforms [W] from the device Philox stream (for RNG parity tests). */
int nqmc_rng_fill(nqmc_ctx* ctx, uint64_t seed, uint64_t step, int64_t walker_id0, int64_t W, float* normals,
                  float* uniforms, nqmc_stream stream);


/* ---- multi-GPU: peer-memory communicator (one process per GPU on one node) -----------------------------------
 * The reference is single-device.  Sharding its walker loop needs two sums over ranks per step:
 * the statistics header before the reverse pass (global mean of E_L, src/nqmc/optimizer/vmc.py:113-126) and the
 * gradient sum after it (vmc.py:126-133).  nqmc_comm_create allocates this rank's inbox and returns an opaque
 * handle (NQMC_COMM_HANDLE_BYTES bytes) to exchange by any out-of-band means (torch.distributed all_gather in
 * nqmc_b200/parallel.py); nqmc_comm_attach maps the peers' inboxes (handles = world x NQMC_COMM_HANDLE_BYTES bytes,
 * rank order).  Once attached, nqmc_walker_step / nqmc_walker_step_host reduce hdr and gsum over all ranks in-stream
 * (one-shot push all-reduce over NVLink, summed in rank order: bitwise identical on every rank); the _a/_b halves
 * stay rank-local so a caller can use its own collective instead.  All ranks must make the same calls in the same
 * order.  world <= 8. */
#define NQMC_COMM_HANDLE_BYTES 64
int nqmc_comm_create(nqmc_ctx* ctx, int rank, int world, void* handle_out);
int nqmc_comm_attach(nqmc_ctx* ctx, const void* handles);
int nqmc_comm_active(const nqmc_ctx* ctx); /* 1 once nqmc_comm_attach succeeded */
/* In-place sum over ranks of vec[0..len), len <= max(param_count, NQMC_HDR_SIZE). */
int nqmc_comm_allreduce(nqmc_ctx* ctx, double* vec, int64_t len, nqmc_stream stream);


/* ---- XLA FFI (csrc/nqmc_xla_ffi.cu) ------------------------------------------------------------------------
 * 1 when the library was built with the jaxlib headers on the include path and exports the custom-call handler
 * symbols NqmcLogPsi / NqmcMhStep / NqmcLocalEnergy / NqmcWalkerStep (XLA_FFI_DEFINE_HANDLER_SYMBOL); 0 otherwise
 * (this image: no jaxlib).  INTEGRATION.md section 4 shows the jax.ffi registration. */
int nqmc_xla_ffi_available(void);

/* ---- device memory, pinned host memory, streams ------------------------------------------------------------
 * The reference lets JAX own every array.  A host that has no array library of its own (the north_star's
 * "no PyTorch") allocates through these; they are validated wrappers over the CUDA runtime and carry no state.
 * Hot-path entry points keep taking caller-owned raw pointers, so XLA buffers (FFI shim) or any other device
 * allocation work unchanged.  Errors: nqmc_mem_last_error() (thread-local). */
enum { NQMC_COPY_H2D = 1, NQMC_COPY_D2H = 2, NQMC_COPY_D2D = 3 };
int nqmc_device_count(void);
int nqmc_dev_alloc(int device, size_t bytes, void** out);
int nqmc_dev_free(int device, void* p);
int nqmc_host_alloc(size_t bytes, void** out); /* pinned, portable */
int nqmc_host_free(void* p);
/* Asynchronous on `stream` (pageable host memory makes the runtime stage the copy). */
int nqmc_dev_memcpy(int device, void* dst, const void* src, size_t bytes, int kind, nqmc_stream stream);
/* rows x width_bytes device-to-device with pitches: stores one kept sample SoA [3N][W] into the sample tensor
 * [3N][n_samples][W] (replaces all_positions[::n_thin], src/nqmc/sampler/metropolis.py:206-213). */
int nqmc_dev_copy_rows(int device, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                       size_t rows, nqmc_stream stream);
int nqmc_dev_memset(int device, void* dst, int value, size_t bytes, nqmc_stream stream);
int nqmc_stream_create(int device, nqmc_stream* out);
int nqmc_stream_destroy(int device, nqmc_stream stream);
/* Work enqueued on `waiter` after this call starts only after everything enqueued on `signal` before it has finished
 * (event record + stream wait; fork / join of independent contexts that run on their own streams). */
int nqmc_stream_wait(int device, nqmc_stream waiter, nqmc_stream signal);
int nqmc_stream_sync(int device, nqmc_stream stream);
const char* nqmc_mem_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NQMC_B200_H */
