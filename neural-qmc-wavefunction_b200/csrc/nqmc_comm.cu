// nqmc_comm.cu -- peer-memory communicator for the two additive exchanges of the walker step
// (one process per GPU, one node, NVLink / NVSwitch).
//
// The reference has no collective; the data-parallel walker loop needs exactly two sums over
// ranks per step: the 8-double statistics header before the reverse pass (global mean of E_L,
// src/nqmc/optimizer/vmc.py:113-126) and the P-double gradient sum after it (vmc.py:126-133).  Both are
// latency-bound (64 B and ~160 KB), so instead of two NCCL launches the step runs a one-shot all-reduce
// over peer memory: every rank PUSHES its vector into an inbox slot on every peer (plain stores through
// NVLink) and sums its own inbox in rank order -- so the result is bitwise identical on all ranks and
// independent of arrival order.  The push is flag-in-data ("low latency" protocol): each 8-byte word carries
// 4 bytes of payload and a 4-byte epoch tag, and an aligned 8-byte access is single-copy atomic (a 16-byte
// vector access is two such words, each validated by its own tag), so the receiver polls the payload words
// themselves; no fence, no separate flag, one NVLink traversal.
//
//   buffer on each rank (cudaMalloc, exported with cudaIpcGetMemHandle):
//     inbox[2 parity][world][2 len_max]  uint64 = (tag << 32) | half of a double
//   Parity = epoch & 1.  Slot reuse is safe with two parities: a peer can only start epoch e+2 after it
//   finished epoch e+1, which needed this rank's e+1 words, which this rank sent after it had finished
//   reading the epoch-e inbox (stream order).
#include <cstdlib>
#include <cstring>

#include "nqmc_internal.cuh"

namespace {

constexpr int MAX_WORLD = NQ_COMM_MAX_WORLD;
constexpr int MAX_CTA = 128;

struct CommDev {
  unsigned long long* inbox[MAX_WORLD];  // inboxes of all ranks (own one included)
  int rank, world;
  int64_t len_max;
};

__device__ __forceinline__ void ld_sys_v2u64(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// One element per thread; CTA c owns elements [256 c, 256 c + 256) on every rank, so no grid-wide synchronisation
// is needed.  All loads of a poll round are independent (one L2 latency per round, not one per word).
__global__ void __launch_bounds__(256) allreduce_f64_kernel(double* __restrict__ vec, const int64_t len, const CommDev c,
                                                            const unsigned long long epoch, const long long budget,
                                                            int* __restrict__ timeout_flag) {
  // The kernel behind the header exchange is the persistent reverse-pass kernel, launched with programmatic stream
  // serialisation: let it become resident and stage its weights while this kernel waits for the peers (it reads the
  // header only after griddepcontrol.wait, i.e. after this grid has completed and flushed).
  nq_pdl_trigger();
  // ... and this kernel is itself launched that way behind the statistics / Jastrow kernel that produces `vec`
  nq_pdl_wait();
  const int par = (int)(epoch & 1ull);
  const long long t_start = clock64();
  const unsigned long long tag = ((epoch + 1ull) & 0xffffffffull) << 32;
  constexpr unsigned long long HI = 0xffffffff00000000ull, LO = 0xffffffffull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    // push: two tagged words per double into slot `rank` of every inbox
    const unsigned long long bits = (unsigned long long)__double_as_longlong(vec[i]);
    const unsigned long long w0 = tag | (bits & LO), w1 = tag | (bits >> 32);
#pragma unroll
    for (int q = 0; q < MAX_WORLD; ++q)
      if (q < c.world) {
        unsigned long long* dst = c.inbox[q] + (((int64_t)par * c.world + c.rank) * c.len_max + i) * 2;
        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
      }
    // pull: poll the own inbox until both words of every rank carry this epoch's tag; sum in rank order
    const unsigned long long* in = c.inbox[c.rank] + ((int64_t)par * c.world * c.len_max + i) * 2;
    unsigned long long a[MAX_WORLD], b[MAX_WORLD];
    bool all;
    do {
      all = true;
#pragma unroll
      for (int q = 0; q < MAX_WORLD; ++q)
        if (q < c.world) ld_sys_v2u64(in + (int64_t)q * c.len_max * 2, a[q], b[q]);
#pragma unroll
      for (int q = 0; q < MAX_WORLD; ++q)
        if (q < c.world) all = all && (a[q] & HI) == tag && (b[q] & HI) == tag;
      // bounded spin: a peer that never makes this call (crashed, or issued a different number of exchanges) must not
      // hang this GPU in a kernel only a device reset clears.  On time-out raise the host-visible flag and leave NaN.
      if (!all && clock64() - t_start > budget) {
        *timeout_flag = 1;
        __threadfence_system();
        vec[i] = __longlong_as_double(0x7ff8000000000000ll);
        return;
      }
    } while (!all);
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < MAX_WORLD; ++q)
      if (q < c.world) s += __longlong_as_double((long long)((a[q] & LO) | (b[q] << 32)));
    vec[i] = s;
  }
}

}  // namespace

struct nq_comm {
  int rank = 0, world = 1;
  int64_t len_max = 0;
  void* local = nullptr;                 // inbox of this rank
  void* peer[MAX_WORLD] = {};            // mapped peers (peer[rank] == local)
  bool attached = false;
  unsigned long long epoch = 0;
  int* timeout_flag = nullptr;           // pinned, mapped host int the kernel raises when its spin budget runs out
  long long budget_cycles = 20000000000ll;   // ~10 s at 1.965 GHz (NQMC_COMM_TIMEOUT_MS overrides)
};

void nq_comm_free(nqmc_ctx* ctx) {
  nq_comm* c = ctx->comm;
  if (!c) return;
  for (int q = 0; q < c->world; ++q)
    if (q != c->rank && c->peer[q]) cudaIpcCloseMemHandle(c->peer[q]);
  if (c->local) cudaFree(c->local);
  if (c->timeout_flag) cudaFreeHost(c->timeout_flag);
  delete c;
  ctx->comm = nullptr;
}

int nq_comm_rank(const nqmc_ctx* ctx) { return ctx->comm ? ctx->comm->rank : 0; }
bool nq_comm_active(const nqmc_ctx* ctx) { return ctx->comm && ctx->comm->attached && ctx->comm->world > 1; }

int nq_comm_create(nqmc_ctx* ctx, int rank, int world, void* handle_out) {
  if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || !handle_out) return NQMC_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) <= NQMC_COMM_HANDLE_BYTES, "handle size");
  nq_comm_free(ctx);
  nq_comm* c = new nq_comm;
  ctx->comm = c;
  c->rank = rank;
  c->world = world;
  c->len_max = ((int64_t)(ctx->sys.P > NQMC_HDR_SIZE ? ctx->sys.P : NQMC_HDR_SIZE) + 15) / 16 * 16;
  const size_t bytes = sizeof(unsigned long long) * 2 * 2 * (size_t)world * (size_t)c->len_max;
  NQMC_CUDA(cudaMalloc(&c->local, bytes));
  NQMC_CUDA(cudaMemset(c->local, 0, bytes));
  NQMC_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->timeout_flag), sizeof(int), cudaHostAllocMapped));
  *c->timeout_flag = 0;
  if (const char* e = getenv("NQMC_COMM_TIMEOUT_MS")) c->budget_cycles = (long long)(atof(e) * 1.965e6);
  NQMC_CUDA(cudaDeviceSynchronize());
  c->peer[rank] = c->local;
  cudaIpcMemHandle_t h;
  std::memset(handle_out, 0, NQMC_COMM_HANDLE_BYTES);
  if (world > 1) {
    NQMC_CUDA(cudaIpcGetMemHandle(&h, c->local));
    std::memcpy(handle_out, &h, sizeof h);
  }
  return NQMC_OK;
}

int nq_comm_attach(nqmc_ctx* ctx, const void* handles) {
  nq_comm* c = ctx->comm;
  if (!c || !handles) return NQMC_ERR_ARG;
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + (size_t)q * NQMC_COMM_HANDLE_BYTES, sizeof h);
    NQMC_CUDA(cudaIpcOpenMemHandle(&c->peer[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  c->attached = true;
  return NQMC_OK;
}

int nq_comm_allreduce(nqmc_ctx* ctx, double* vec, int64_t len, cudaStream_t st) {
  nq_comm* c = ctx->comm;
  if (!c || !c->attached) {
    ctx->err = "nqmc_comm_allreduce: no communicator attached";
    return NQMC_ERR_ARG;
  }
  if (len <= 0 || len > c->len_max || !vec) return NQMC_ERR_ARG;
  if (c->world == 1) return NQMC_OK;
  if (*static_cast<volatile int*>(c->timeout_flag)) {
    ctx->err = "peer-memory all-reduce timed out in an earlier call: a rank did not make the matching call "
               "(all ranks must issue the same exchanges in the same order); destroy the contexts and re-attach";
    return NQMC_ERR_STATE;
  }
  CommDev d;
  for (int q = 0; q < MAX_WORLD; ++q) {
    d.inbox[q] = q < c->world ? reinterpret_cast<unsigned long long*>(c->peer[q]) : nullptr;
  }
  d.rank = c->rank;
  d.world = c->world;
  d.len_max = c->len_max;
  int n_cta = (int)((len + 255) / 256);
  if (n_cta > MAX_CTA) n_cta = MAX_CTA;
  int* flag_dev = nullptr;
  NQMC_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&flag_dev), c->timeout_flag, 0));
  NQMC_CUDA(nq_launch_pdl(allreduce_f64_kernel, dim3(n_cta), dim3(256), 0, st, vec, len, d, c->epoch, c->budget_cycles, flag_dev));
  c->epoch++;
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
