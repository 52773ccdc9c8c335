// nqmc_comm.cu -- peer-memory communicator for the two additive exchanges of the walker step
// (one process per GPU, one node, NVLink / NVSwitch).
//
// The reference has no collective; the data-parallel walker loop needs exactly two sums over
// ranks per step: the 8-double statistics header before the reverse pass (global mean of E_L,
// src/nqmc/optimizer/vmc.py:113-126) and the P-double gradient sum after it (vmc.py:126-133).  Both are
// latency-bound (64 B and ~160 KB), so instead of two NCCL launches the step runs a one-shot all-reduce
// over peer memory: every rank PUSHES its vector into an inbox slot on every peer (plain stores through
// NVLink) and sums its own inbox in rank order -- so the result is bitwise identical on all ranks and
// independent of arrival order.  The push is flag-in-data ("low latency" protocol): each 8-byte store carries
// 4 bytes of payload and a 4-byte epoch tag, and an aligned 8-byte store is single-copy atomic, so the
// receiver polls the payload words themselves; no fence, no separate flag, one NVLink traversal.
//
//   buffer on each rank (cudaMalloc, exported with cudaIpcGetMemHandle):
//     inbox[2 parity][world][2 len_max]  uint64 = (tag << 32) | half of a double
//   Parity = epoch & 1.  Slot reuse is safe with two parities: a peer can only start epoch e+2 after it
//   finished epoch e+1, which needed this rank's e+1 words, which this rank sent after it had finished
//   reading the epoch-e inbox (stream order).
#include <cstring>

#include "nqmc_internal.cuh"

namespace {

constexpr int MAX_WORLD = NQ_COMM_MAX_WORLD;
constexpr int MAX_CTA = 32;

struct CommDev {
  unsigned long long* inbox[MAX_WORLD];  // inboxes of all ranks (own one included)
  int rank, world;
  int64_t len_max;
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// CTA c owns the slice [len c / n, len (c+1) / n) on every rank, so no grid-wide synchronisation is needed.
__global__ void __launch_bounds__(256) allreduce_f64_kernel(double* __restrict__ vec, const int64_t len, const CommDev c,
                                                            const unsigned long long epoch) {
  const int tid = threadIdx.x, cta = blockIdx.x;
  const int par = (int)(epoch & 1ull);
  const unsigned long long tag = ((epoch + 1ull) & 0xffffffffull) << 32;
  const int64_t lo = len * cta / gridDim.x, hi = len * (cta + 1) / gridDim.x;
  // push: two tagged words per double into slot `rank` of every inbox
  for (int64_t i = lo + tid; i < hi; i += blockDim.x) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(vec[i]);
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    for (int q = 0; q < c.world; ++q) {
      unsigned long long* dst = c.inbox[q] + (((int64_t)par * c.world + c.rank) * c.len_max + i) * 2;
      st_sys_u64(dst, w0);
      st_sys_u64(dst + 1, w1);
    }
  }
  // pull: poll the own inbox until every word of the slice carries this epoch's tag; sum in rank order
  const unsigned long long* in = c.inbox[c.rank] + (int64_t)par * c.world * c.len_max * 2;
  for (int64_t i = lo + tid; i < hi; i += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < c.world; ++q) {
      const unsigned long long* src = in + ((int64_t)q * c.len_max + i) * 2;
      unsigned long long w0, w1;
      do { w0 = ld_sys_u64(src); } while ((w0 & 0xffffffff00000000ull) != tag);
      do { w1 = ld_sys_u64(src + 1); } while ((w1 & 0xffffffff00000000ull) != tag);
      s += __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    }
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
};

void nq_comm_free(nqmc_ctx* ctx) {
  nq_comm* c = ctx->comm;
  if (!c) return;
  for (int q = 0; q < c->world; ++q)
    if (q != c->rank && c->peer[q]) cudaIpcCloseMemHandle(c->peer[q]);
  if (c->local) cudaFree(c->local);
  delete c;
  ctx->comm = nullptr;
}

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
  CommDev d;
  for (int q = 0; q < MAX_WORLD; ++q) {
    d.inbox[q] = q < c->world ? reinterpret_cast<unsigned long long*>(c->peer[q]) : nullptr;
  }
  d.rank = c->rank;
  d.world = c->world;
  d.len_max = c->len_max;
  int n_cta = (int)((len + 1023) / 1024);
  if (n_cta > MAX_CTA) n_cta = MAX_CTA;
  allreduce_f64_kernel<<<n_cta, 256, 0, st>>>(vec, len, d, c->epoch);
  c->epoch++;
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}
