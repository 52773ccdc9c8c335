// Microbenchmark: cycles per tcgen05.mma kind::tf32 (cta_group::1, K = 8) for the operand formats used in this
// repository.  One CTA, one issuing thread, NREP back-to-back MMAs into the same accumulator, one commit, wait.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
struct Cfg { int M, N, a_mn, b_mn, a_tmem, n_acc, n_iss; };   // n_acc: number of independent accumulators cycled through
__global__ void k(long long* out, Cfg c, int nrep) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;
  if (tid == 0) { for (int b = 0; b < 4; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) |
                         ((uint32_t)(c.M >> 4) << 24);
  if ((tid & 31) == 0 && (tid >> 5) < c.n_iss) {
    const int w = tid >> 5;
    const uint64_t da = c.a_mn ? desc(smem_u32(sm), 16384, 512, 1) : desc(smem_u32(sm), 128 * 16, 128, 0);
    const uint64_t db = c.b_mn ? desc(smem_u32(sm) + 65536, 16384, 512, 1) : desc(smem_u32(sm) + 65536, 64 * 16, 128, 0);
    const long long t0 = clock64();
    // 16 MMAs per loop iteration, fully unrolled; accumulator = tm + N * (j % n_acc) with compile-time j
    for (int i = 0; i < nrep; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t tmd = tm + (uint32_t)(64 * w) + (uint32_t)(c.N * (c.n_acc == 1 ? 0 : (c.n_acc == 2 ? (j & 1) : (j & 3))));
        if (c.a_tmem)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmd),
                       "r"(tm + 256), "l"(db), "r"(idesc), "r"(1) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmd), "l"(da),
                       "l"(db), "r"(idesc), "r"(1) : "memory");
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar[w])), "r"(0));
    const long long t2 = clock64();
    if (w == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const Cfg cfgs[] = {{128, 64, 0, 0, 1, 1, 1}, {128, 64, 0, 0, 1, 2, 1}, {128, 64, 0, 0, 1, 4, 1}, {128, 128, 0, 0, 1, 2, 1}, {128, 256, 0, 0, 0, 2, 1}, {128, 64, 0, 0, 1, 1, 2}, {128, 64, 0, 0, 1, 1, 4}, {128, 32, 0, 0, 1, 1, 1}, {128, 128, 0, 0, 1, 1, 1},
                      {128, 192, 0, 0, 1, 1, 1}, {128, 256, 0, 0, 1, 1, 1}, {128, 64, 1, 1, 0, 1, 1}, {128, 64, 1, 1, 0, 1, 2}, {64, 16, 1, 0, 0, 1, 1},
                      {64, 16, 1, 0, 0, 1, 4}};
  const int nrep = 512;
  for (auto& c : cfgs) {
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
      k<<<1, 128, 96 * 1024>>>(d, c, nrep);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d acc=%d issuers=%d A:%s B:%s  issue %.1f cyc/MMA   complete %.1f cyc/MMA\n", c.M, c.N, c.n_acc, c.n_iss, c.a_tmem ? "TMEM   " : (c.a_mn ? "smem MN" : "smem K "),
           c.b_mn ? "smem MN" : "smem K ", (double)h[0] / nrep, (double)h[1] / nrep);
  }
  return 0;
}
