// nqmc_internal.cuh -- shared device/host definitions for libnqmc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/nqmc_b200.h"

#define NQMC_MAX_ATOMS 16
#define NQMC_MAX_ELEC 16
#define NQMC_MAX_NS 8    // electrons per spin
#define NQMC_MAX_MLP 16  // n_up + n_dn orbital networks
#define NQMC_CUSP_H 16
#define NQMC_BF_H 16

// ---------------------------------------------------------------------------------------------
// Parameter map: offsets of every block inside the flat theta vector (sorted-key pytree order;
// inside one flax Dense, 'bias' sorts before 'kernel').
constexpr int NQ_COMM_MAX_WORLD = 8;
struct nq_comm;

struct MlpOff {     // one Dense(F,H)-tanh-Dense(H,H)-tanh-Dense(H,1) network
  int b1, W1, b2, W2, b3, W3;
};

#define NQMC_SIMPLE_MAX_HIDDEN 6
struct SimpleNet {  // kind 0: Dense(dim[0],dim[1])-act-...-Dense(dim[n_hidden],1); dim[0] = 3N
  int n_hidden, act;
  int dim[NQMC_SIMPLE_MAX_HIDDEN + 1];
  int offW[NQMC_SIMPLE_MAX_HIDDEN + 1], offb[NQMC_SIMPLE_MAX_HIDDEN + 1];   // entry n_hidden = output layer
};

struct DevSys {     // passed by value to kernels (<= 4 KB kernel-parameter budget)
  int A, N, n_up, n_dn;
  int kind, H, use_cusp, use_bf;
  int F;            // kind 1: 3 + 2A orbital input features ; kind 0: 3N
  int n_mlp;        // n_up + n_dn
  int n_ent;        // n_up^2 + n_dn^2 entries of the two orbital matrices
  int P;
  float v_nn, env_alpha, cusp_same, cusp_diff;
  float R[NQMC_MAX_ATOMS * 3];
  float Z[NQMC_MAX_ATOMS];
  // theta offsets
  int jas_a_ee, jas_a_en, jas_b_ee, jas_b_en;
  int ee_cusp_b;
  int en_cusp_b0[NQMC_MAX_ATOMS], en_cusp_W0[NQMC_MAX_ATOMS], en_cusp_b1[NQMC_MAX_ATOMS], en_cusp_W1[NQMC_MAX_ATOMS];
  MlpOff bf;                   // backflow eta network (F=2, H=16)
  SimpleNet simple;            // kind 0 network
  MlpOff mlp[NQMC_MAX_MLP];    // m < n_up: spin-up orbital m ; else spin-down orbital m - n_up
};

// entry index of (spin, electron-in-spin i, orbital k) in the orbital scratch
// Work items of the persistent orbital kernels are (walker block wb, electron-of-spin e) pairs numbered wb * ns + e and
// dealt round-robin: CTA `first`, then first + stride, ...  The decomposition is advanced incrementally -- one division
// per thread at kernel start instead of a 64-bit div + mod per tile (which cost ~60 instructions of a ~500-instruction
// tile in the value kernel).  The launchers check that the item count fits 31 bits.
struct ItemIter {
  int wb, e, dq, dr, ns;
  __host__ __device__ ItemIter(int first, int stride, int ns_) : ns(ns_) {
    wb = first / ns_; e = first - wb * ns_;
    dq = stride / ns_; dr = stride - dq * ns_;
  }
  __host__ __device__ void next() {
    wb += dq; e += dr;
    if (e >= ns) { e -= ns; ++wb; }
  }
};

__host__ __device__ inline int ent_index(const DevSys& s, int spin, int i, int k) {
  return spin == 0 ? i * s.n_up + k : s.n_up * s.n_up + i * s.n_dn + k;
}

struct LeafInfo {
  std::string path;
  int64_t rows, cols, offset;
};

struct nqmc_ctx {
  int device = 0;
  int sm_count = 148;
  DevSys sys{};
  std::vector<LeafInfo> leaves;
  std::string err;
  int64_t launches = 0;
  bool use_tc = true;          // tcgen05 3xTF32 path for the H x H layers where a variant exists (NQMC_TC=0 disables)
  bool want_bf = false;        // backflow leaves exist in theta (use_backflow requested)
  // device state
  float* theta = nullptr;      // [P]
  int64_t cap_w = 0;           // reserved walkers
  float* r_prop = nullptr;     // [3N][cap_w]
  float* phi = nullptr;        // [5 or 10][n_ent][cap_w]  orbital value / grad xyz / laplacian (or 6 Hessian entries)
  float* inv = nullptr;        // [n_ent][cap_w]     inverse orbital matrices (seeds of the reverse pass)
  float* logabs = nullptr;     // [cap_w]
  float* sign = nullptr;       // [cap_w]
  float* e_scratch = nullptr;  // [cap_w]
  float* wgt = nullptr;        // [cap_w]
  float* x_cur = nullptr;      // [3N][cap_w] backflow coordinates of the current positions
  float* x_prop = nullptr;     // [3N][cap_w] backflow coordinates of the proposal
  float* vvec = nullptr;       // [3N][cap_w] v_i = sum_k inv[k,i] grad phi_k(x_i)  (backflow reverse pass)
  float* gbf = nullptr;        // [6N][cap_w] G_i = sum_mu d_mu x_i d_mu x_i^T (weights of the backflow Laplacian channel)
  float* bfw = nullptr;        // [nq_backflow_ws_floats][cap_w] eta table, Lap x, A, Q and partial sums (nqmc_backflow.cu)
  float* r_soa = nullptr;      // [3N][cap_w]  (host-entry staging)
  float* nrm_soa = nullptr;    // [3N][cap_w]
  float* aos_stage = nullptr;  // [cap_w][N][3]
  float* uni_stage = nullptr;  // [cap_w]
  float* logp_stage = nullptr; // [cap_w]
  float* nacc_stage = nullptr; // [cap_w]
  double* hdr_stage = nullptr; // [8]
  double* gsum_stage = nullptr;// [P]
  double* part_walk = nullptr; // [n_blk_walk][NPART_WALK] per-block partial sums of the per-walker kernels
  int n_blk_walk = 0;
  float* part_mlp = nullptr;   // [n_cta_bwd][P_mlp_max] per-CTA partial weight gradients
  int n_cta_bwd = 0;
  int p_mlp = 0;               // scalars in one orbital network
  // hidden width 128, tensor-core reverse pass (nqmc_orbital_bwd128.cu): per-row activations, unit-major
  float* bw_h = nullptr;       // [n_mlp][128][bw_rows] h1
  float* bw_d = nullptr;       // [n_mlp][128][bw_rows] delta2
  float* bw_e = nullptr;       // [n_mlp][128][bw_rows] delta1
  float* bw_f = nullptr;       // [n_mlp][32][bw_rows]  [1 | features]
  int64_t bw_rows = 0;
  float* l1cache = nullptr;    // [sm_count][16 chunks][10][512] layer-1 outputs parked between the two output passes of the width-128 5-channel kernel
  double* hdr_tmp = nullptr;   // [8]
  double* jas_sums = nullptr;  // [8] sum O, sum E O of the Jastrow scalars (fused walker step)
  const float* jas_valid_for = nullptr;   // e_l pointer the sums belong to (nullptr: not valid)
  // width-128 reverse pass fed by the energy pass of the same fused step (nqmc_orbital_fwd_pc.cu, BwdOut): want_act asks the
  // 5-channel kernel to park h1 / h2 / features, act_stored reports that it did, act_valid_for = the e_l pointer they belong to
  bool want_act = false, act_stored = false;
  const float* act_valid_for = nullptr;
  int64_t act_W = 0;
  float* bw_h2 = nullptr;      // [n_mlp][128][bw_rows] h2 parked by the 5-channel pass
  const float* sigma_dev = nullptr;       // device-resident proposal width (set by nqmc_sample while it runs)
  // NQMC_GUARD=1: guard bands around every scratch buffer (nqmc_debug_check_guards)
  struct Guard { char* base; size_t bytes, padded; const char* name; };
  bool guard = false;
  std::vector<Guard> guards;
  bool theta_dirty = true;                // a theta writer was enqueued since the last PDL-attributed launch (see nq_launch_pdl)
  // optional per-kernel CUDA-event timing (bench.py's roofline leg); off by default
  bool prof_on = false;
  static constexpr int PROF_KINDS = 6, PROF_RING = 512;
  cudaEvent_t prof_ev[PROF_KINDS][PROF_RING][2] = {};
  int prof_n[PROF_KINDS] = {};
  // host entry point: results of phase A travel back on a side stream while phase B (the reverse pass) runs
  struct nq_comm* comm = nullptr;   // peer-memory communicator (nqmc_comm.cu); nullptr: single GPU or NCCL outside
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_phase_a = nullptr, ev_side_done = nullptr, ev_mh_done = nullptr;
  // pipelined host entry (nqmc_walker_step_host_upload / _async / _wait): the next step's walkers are uploaded into the
  // second SoA buffer on up_stream while the reverse pass of the running step still reads r_soa
  float* r_soa2 = nullptr;     // [3N][cap_w]
  cudaStream_t up_stream = nullptr;
  cudaEvent_t ev_upload = nullptr, ev_all_done = nullptr;
  bool up_pending = false, host_step_issued = false;
  const float* up_r = nullptr; const float* up_logp = nullptr; const float* up_nrm = nullptr; const float* up_uni = nullptr;
  int64_t up_W = 0;
};
enum { NQ_PROF_EVAL1 = 0, NQ_PROF_EVAL5 = 1, NQ_PROF_BWD = 2, NQ_PROF_ASSEMBLE = 3, NQ_PROF_ACCEPT = 4, NQ_PROF_MISC = 5 };

struct ProfScope {   // records start/stop events around one launch on the launching stream
  nqmc_ctx* c; int kind; cudaStream_t st; int slot = -1;
  ProfScope(nqmc_ctx* ctx, int k, cudaStream_t s) : c(ctx), kind(k), st(s) {
    if (c->prof_on && c->prof_n[k] < nqmc_ctx::PROF_RING) {
      slot = c->prof_n[k]++;
      if (!c->prof_ev[k][slot][0]) { cudaEventCreate(&c->prof_ev[k][slot][0]); cudaEventCreate(&c->prof_ev[k][slot][1]); }
      cudaEventRecord(c->prof_ev[k][slot][0], st);
    }
  }
  ~ProfScope() { if (slot >= 0) cudaEventRecord(c->prof_ev[kind][slot][1], st); }
};

// ---------------------------------------------------------------------------------------------
#define NQMC_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                                 \
      return NQMC_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

#define NQMC_LAUNCH_CHECK()                                                                          \
  do {                                                                                               \
    ctx->launches++;                                                                                 \
    cudaError_t _e = cudaGetLastError();                                                             \
    if (_e != cudaSuccess) {                                                                         \
      ctx->err = std::string("kernel launch: ") + cudaGetErrorString(_e);                            \
      return NQMC_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// device math
// 2^t with one MUFU.EX2 and nothing else.  exp2f() wraps the same instruction in a denormal-range fix-up (FSETP, two
// predicated FMULs per value: t < -126 is evaluated as (2^(t/2))^2); the flush-to-zero form returns 0 there, and every
// caller adds 1 to the result, where 0 and a subnormal give the same sum -- so tanh is bit-identical and three
// instructions per value shorter.
__device__ __forceinline__ float nq_ex2(float t) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return e;
}
__device__ __forceinline__ float nq_tanh(float x) {
  // tanh(x) = 1 - 2/(exp(2x)+1); ex2.approx + rcp.approx: abs error ~1e-7, saturates cleanly.
  float e = nq_ex2(2.885390081777927f * x);     // 2*log2(e)
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// Packed FP32 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per instruction, each component
// rounded exactly like the scalar instruction).  The SIMT stages of the tensor-core kernels are issue-bound, and their
// per-unit math comes in pairs (two hidden units / two accumulator columns per thread), so packing halves the issue
// slots of that arithmetic.  ptxas folds a duplicated scalar into a broadcast operand.
__device__ __forceinline__ float2 nq_fma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 nq_mul2(const float2 a, const float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 nq_add2(const float2 a, const float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 nq_bc2(const float a) { return make_float2(a, a); }
// nq_tanh on a pair: same formula, same roundings (2 * rcp is exact, so fma(-2, rcp, 1) == 1 - __fdividef(2, .))
__device__ __forceinline__ float2 nq_tanh2(const float2 x) {
  const float2 t = nq_mul2(x, nq_bc2(2.885390081777927f));
  const float2 e = nq_add2(make_float2(nq_ex2(t.x), nq_ex2(t.y)), nq_bc2(1.0f));
  float rx, ry;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(e.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(e.y));
  return nq_fma2(nq_bc2(-2.0f), make_float2(rx, ry), nq_bc2(1.0f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 (Salmon et al. SC'11).  Stream layout documented in oracle/philox.py.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
__device__ __forceinline__ void box_muller(uint32_t x0, uint32_t x1, float& n0, float& n1) {
  float a = ((float)(x0 >> 8) + 1.0f) * 5.9604644775390625e-8f;   // (0,1]
  float b = (float)(x1 >> 8) * 5.9604644775390625e-8f;            // [0,1)
  float rad = sqrtf(-2.0f * logf(a));
  float s, c;
  sincospif(2.0f * b, &s, &c);
  n0 = rad * c;
  n1 = rad * s;
}

// ---------------------------------------------------------------------------------------------
// kernels' host launchers (defined across the .cu files)
int nq_launch_orb_eval(nqmc_ctx* ctx, const float* r, int64_t W, int channels, float* phi, cudaStream_t st);
int nq_launch_orb_bwd(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr_center, const float* e_l,
                      int64_t W, double* out, cudaStream_t st);
int nq_launch_propose(nqmc_ctx* ctx, const float* r, const float* normals, uint64_t seed, uint64_t step,
                      int64_t wid0, float sigma, int64_t W, float* r_prop, cudaStream_t st);
int nq_launch_accept(nqmc_ctx* ctx, float* r, float* logp, const float* r_prop, const float* uniforms, uint64_t seed,
                     uint64_t step, int64_t wid0, int64_t W, uint8_t* accept, float* n_acc, cudaStream_t st);
int nq_launch_assemble(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, bool want_energy,
                       float* logabs, float* sign, cudaStream_t st);
int nq_launch_stats(nqmc_ctx* ctx, const float* e_l, const float* n_acc_flags, int64_t W, double* hdr, cudaStream_t st);
int nq_launch_walker_grads(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr_center,
                           const float* e_l, int64_t W, double* out, cudaStream_t st);
int nq_simple_logpsi(nqmc_ctx* ctx, const float* r, int64_t W, float* logabs, cudaStream_t st);
int nq_simple_energy(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, cudaStream_t st);
int nq_simple_grads(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr_center, const float* e_l,
                    int64_t W, double* out, cudaStream_t st);
int nq_simple_grads_per_walker(nqmc_ctx* ctx, const float* r, int64_t W, float* Ot, int64_t ld, cudaStream_t st);
// Programmatic dependent launch (PDL): a kernel launched through nq_launch_pdl may become resident while its
// predecessor in the stream is still draining (once every CTA of the predecessor has called nq_pdl_trigger or
// exited); it must call nq_pdl_wait() before touching anything a predecessor wrote.  Used so that the per-CTA weight
// staging of the tensor-core kernels (reads theta only) overlaps the small kernel in front of them.
// INVARIANT: such a kernel reads ctx->theta BEFORE nq_pdl_wait(), so the kernel directly in front of it in the stream
// must not write theta.  The only writers are adam_apply_kernel (nqmc_adam_step) and the set_params copies; both set
// ctx->theta_dirty, and nq_launch_pdl_guarded() launches WITHOUT the PDL attribute (full stream serialisation) while
// it is set, clearing it afterwards.
template <typename... KArgs, typename... Args>
inline cudaError_t nq_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
// PDL launch that honours the theta invariant above.
template <typename... KArgs, typename... Args>
inline cudaError_t nq_launch_pdl_guarded(nqmc_ctx* ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st, Args&&... args) {
  if (ctx->theta_dirty) {
    ctx->theta_dirty = false;
    kern<<<grid, block, smem, st>>>(KArgs(args)...);
    return cudaGetLastError();
  }
  return nq_launch_pdl(kern, grid, block, smem, st, std::forward<Args>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void nq_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void nq_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// peer-memory communicator (nqmc_comm.cu)
int nq_comm_create(nqmc_ctx* ctx, int rank, int world, void* handle_out);
int nq_comm_attach(nqmc_ctx* ctx, const void* handles);
int nq_comm_allreduce(nqmc_ctx* ctx, double* vec, int64_t len, cudaStream_t st);
bool nq_comm_active(const nqmc_ctx* ctx);
int nq_comm_rank(const nqmc_ctx* ctx);
void nq_comm_free(nqmc_ctx* ctx);
int nq_blocked_stats(nqmc_ctx* ctx, const float* e_l, int64_t S, int64_t W, int n_blocks, double* out, cudaStream_t st);
int nq_adapt_sigma(nqmc_ctx* ctx, float* n_acc, int64_t W, int steps, float target, float gain, float* sigma_dev,
                   cudaStream_t st);
int nq_aos_soa(nqmc_ctx* ctx, const float* src, float* dst, int64_t W, bool to_soa, cudaStream_t st);
int nq_rng_fill(nqmc_ctx* ctx, uint64_t seed, uint64_t step, int64_t wid0, int64_t W, float* normals, float* uniforms,
                cudaStream_t st);
int nq_fill(nqmc_ctx* ctx, float* dst, float v, int64_t n, cudaStream_t st);
int nq_accept_logabs(nqmc_ctx* ctx, float* r, float* logp, const float* r_prop, const float* logabs_prop,
                     const float* uniforms, uint64_t seed, uint64_t step, int64_t wid0, int64_t W, uint8_t* accept,
                     float* n_acc, cudaStream_t st);
int nq_launch_assemble_inv(nqmc_ctx* ctx, const float* r, int64_t W, cudaStream_t st);
int nq_ffma_peak(nqmc_ctx* ctx, int reps, double* tflops, cudaStream_t st);
size_t nq_backflow_ws_floats(const DevSys& s);
int nq_antisym_grads_per_walker(nqmc_ctx* ctx, const float* r, int64_t W, float* Ot, int64_t ld, cudaStream_t st);
int nq_launch_backflow_x(nqmc_ctx* ctx, const float* r, int64_t W, float* x, cudaStream_t st);
int nq_launch_assemble_bf(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, float* parts, int gw, cudaStream_t st);
int nq_launch_backflow_xg(nqmc_ctx* ctx, const float* r, int64_t W, float* x, float* G, cudaStream_t st);
int nq_launch_bf_grads(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr_center, const float* e_l,
                       int64_t W, double* out, cudaStream_t st);
int nq_launch_orb_bwd_tc(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                         int* n_chunks_out, cudaStream_t st);
int nq_launch_orb_bwd128(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                         int* n_chunks_out, cudaStream_t st);      // hidden width 128: three TMA / tcgen05 kernels
int64_t nq_bwd128_rows(const nqmc_ctx* ctx, int64_t W);
int nq_launch_orb_eval1_tc(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st);
int nq_launch_orb_eval1_l1tc(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st);   // both layers on tensor cores
int nq_adam_step(nqmc_ctx* ctx, const double* gsum, const double* hdr, float* m, float* v, int64_t count, float lr,
                 float clip, double* norm_out, cudaStream_t st);
int nq_launch_assemble_stats(nqmc_ctx* ctx, const float* r, int64_t W, float* e_l, const float* acc_flag, double* hdr,
                             cudaStream_t st);
int nq_launch_jastrow_from_sums(nqmc_ctx* ctx, const double* hdr, double* out, cudaStream_t st);
// tcgen05 forward pass (nqmc_orbital_fwd_pc.cu).  G == nullptr: ordinary Laplacian channel (H = 128 only; H = 64 has
// nqmc_orbital_lap_pc.cu); G != nullptr: [6N][W] weights of the G-weighted Laplacian L_G (backflow), H = 64 or 128.
int nq_launch_orb_fwd_pc(nqmc_ctx* ctx, const float* r, int64_t W, int channels, float* phi, const float* G, cudaStream_t st);
int nq_launch_orb_lap_pc(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st);
