// (see also m64_probe.cu)
// Probe: how does tcgen05.mma kind::tf32 read an MN-major B operand with layout type 1 (SWIZZLE_128B_BASE32B)?
// A = identity rows (K-major, SWIZZLE_NONE, known-good format), B smem word i = float(i)  =>  D[k][n] = index of the
// word the hardware uses as B[k][n].  Prints the map for K = 8, N = 64.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int M = 128, N = 64, K = 8;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__global__ void k(float* D, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  __shared__ __align__(1024) float sb[2048];
  __shared__ __align__(1024) float sa[2 * M * 4];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2048; i += blockDim.x) sb[i] = (float)i;
  for (int i = tid; i < 2 * M * 4; i += blockDim.x) {
    const int k4 = i / (M * 4), m = (i / 4) % M, kk = k4 * 4 + (i & 3);
    sa[i] = (m == kk) ? 1.f : 0.f;                       // [k/4][m][k%4]
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    const uint64_t da = desc(smem_u32(sa), M * 16, 128, 0), db = desc(smem_u32(sb), lbo, sbo, layout);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(da),
                 "l"(db), "r"(idesc), "r"(0)
                 : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid < 128) {
    const uint32_t base = tm + ((uint32_t)((tid >> 5) * 32) << 16);
    for (int c = 0; c < N; c += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(base + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[tid * N + c + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}
int main() {
  static float D[M * N];
  float* dD;
  cudaMalloc(&dD, sizeof D);
  const uint32_t cfg[][3] = {{1024, 512, 1}, {4096, 512, 1}, {512, 4096, 1}, {1024, 512, 2}};
  for (auto& c : cfg) {
    cudaMemset(dD, 0, sizeof D);
    k<<<1, 128>>>(dD, c[0], c[1], c[2]);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D, dD, sizeof D, cudaMemcpyDeviceToHost);
    printf("LBO=%u SBO=%u layout=%u: %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
    for (int kk = 0; kk < K; ++kk) {
      printf(" k=%d:", kk);
      for (int n = 0; n < N; ++n) printf(" %d", (int)D[kk * N + n]);
      printf("\n");
    }
  }
  return 0;
}
