// Experiment: which MN-major shared-memory operand formats does tcgen05.mma kind::tf32 accept?
//   variants 0,1: SWIZZLE_NONE MN-major (result on B200: accumulator stays all-zero, no error)
//   variants 2..5: SWIZZLE_128B_BASE32B (layout type 1): atoms of 32 MN x 4 K, 128-byte rows, 32-bit words of a
//                  16-byte chunk XOR-permuted by (chunk index mod 4); with / without the XOR, LBO/SBO roles swapped
//   D[m][n] = sum_k X[k][m] * Y[k][n],  M=128, N=64, K=16 (two K=8 instructions)
// layout under test (bytes): addr(mn, k) = (mn/4)*SBO + (k/8)*LBO + (k%8)*16 + (mn%4)*4
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o mn_major_tf32 mn_major_tf32.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 16;
constexpr int LBO = 128, SBO_A = (K / 8) * 128, SBO_B = (K / 8) * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
// SWIZZLE_128B_BASE32B as probed (scripts/exp/mn_major_probe.cu): atoms of 4 K-rows x 128 bytes (32 MN elements);
// the 32-byte chunk index inside a row is XORed with the row index; written here as a function of the byte address
__device__ __forceinline__ int off_b32(int mn, int k, int lbo, int sbo, bool xr) {
  const int lin = (mn >> 5) * lbo + (k >> 2) * sbo + (k & 3) * 128 + (mn & 31) * 4;
  return xr ? (lin ^ (((lin >> 7) & 3) << 5)) : lin;
}

__global__ void k(const float* X, const float* Y, float* D, int variant) {
  const int swap_lbo_sbo = variant & 1;
  const bool b32 = variant >= 2, xr = variant < 4;
  const int SBO32 = (variant & 1) ? 528 : 512, LBO32 = (K / 4) * SBO32 + ((variant & 1) ? 496 : 0);
  __shared__ __align__(1024) uint8_t sa[(M / 4) * SBO_A + 4096];
  __shared__ __align__(1024) uint8_t sb[(N / 4) * SBO_B + 4096];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int kk = i / M, m = i % M;
    if (b32) *reinterpret_cast<float*>(sa + off_b32(m, kk, LBO32, SBO32, xr)) = X[kk * M + m];
    else *reinterpret_cast<float*>(sa + (m / 4) * SBO_A + (kk / 8) * LBO + (kk % 8) * 16 + (m % 4) * 4) = X[kk * M + m];
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int kk = i / N, n = i % N;
    if (b32) *reinterpret_cast<float*>(sb + off_b32(n, kk, LBO32, SBO32, xr)) = Y[kk * N + n];
    else *reinterpret_cast<float*>(sb + (n / 4) * SBO_B + (kk / 8) * LBO + (kk % 8) * 16 + (n % 4) * 4) = Y[kk * N + n];
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
  // D f32, A/B tf32, A and B MN-major (bits 15, 16), N, M
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t da, db;
      if (b32) {
        const uint32_t l = LBO32, sb_ = SBO32;
        da = desc(smem_u32(sa) + ks * 2 * SBO32, l, sb_, 1);
        db = desc(smem_u32(sb) + ks * 2 * SBO32, l, sb_, 1);
      } else if (!swap_lbo_sbo) {
        da = desc(smem_u32(sa) + ks * LBO, LBO, SBO_A);
        db = desc(smem_u32(sb) + ks * LBO, LBO, SBO_B);
      } else {
        da = desc(smem_u32(sa) + ks * LBO, SBO_A, LBO);
        db = desc(smem_u32(sb) + ks * LBO, SBO_B, LBO);
      }
      const uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(da),
                   "l"(db), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid < 128) {
    const uint32_t base = tm + ((uint32_t)((tid >> 5) * 32) << 16);
    for (int c = 0; c < N; c += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(base + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[tid * N + c + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

int main() {
  static float X[K * M], Y[K * N], D[M * N], R[M * N];
  srand(1);
  for (auto& v : X) v = (float)(rand() % 17 - 8) / 4.f;     // exactly representable in tf32
  for (auto& v : Y) v = (float)(rand() % 17 - 8) / 4.f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float a = 0;
      for (int kk = 0; kk < K; ++kk) a += X[kk * M + m] * Y[kk * N + n];
      R[m * N + n] = a;
    }
  float *dX, *dY, *dD;
  cudaMalloc(&dX, sizeof X); cudaMalloc(&dY, sizeof Y); cudaMalloc(&dD, sizeof D);
  cudaMemcpy(dX, X, sizeof X, cudaMemcpyHostToDevice);
  cudaMemcpy(dY, Y, sizeof Y, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 6; ++variant) {
    cudaMemset(dD, 0, sizeof D);
    k<<<1, 128>>>(dX, dY, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D, dD, sizeof D, cudaMemcpyDeviceToHost);
    double err = 0, mag = 0;
    for (int i = 0; i < M * N; ++i) { err = fmax(err, fabs((double)D[i] - R[i])); mag = fmax(mag, fabs((double)D[i])); }
    printf("variant %d (%s%s%s): %s  max|D-ref| = %g  max|D| = %g\n", variant, variant >= 2 ? "SW128_BASE32B" : "SWIZZLE_NONE",
           variant >= 2 ? (variant < 4 ? " xor" : " no-xor") : "", (variant & 1) ? (variant >= 2 ? " SBO=528" : " LBO<->SBO swapped") : "", cudaGetErrorString(e), err, mag);
  }
  return 0;
}
