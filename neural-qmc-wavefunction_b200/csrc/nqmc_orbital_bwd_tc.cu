// nqmc_orbital_bwd_tc.cu -- tensor-core (tcgen05 / TMEM, 3xTF32) reverse pass of the orbital networks for
// hidden width 64.  Same math, interface and slab layout as orb_bwd_kernel<64,...> (nqmc_orbital.cu), which
// stays as the FP32 FFMA parity anchor.
//
// Per 128-row tile (128 consecutive walkers, one electron of the network's spin), seed = weight * inv:
//     h1 = tanh(f W1 + b1)                      FP32 FFMA, thread-per-row (K = 3+2A is tiny)
//     U2 = h1 W2                                UMMA  M=128 rows, N=64, K=64    A: h1 in TENSOR MEMORY, B: W2 smem
//     h2 = tanh(U2 + b2); dW3 += seed h2; delta2 = seed W3 (1-h2^2)
//     D1 = delta2 W2^T                          UMMA  M=128 rows, N=64, K=64    A: delta2 in TMEM,      B: W2^T smem
//     G2 += h1^T delta2                         UMMA  M=128 (64 units used), N=64, K=128 rows          (smem x smem)
//     db2 += sum_rows delta2                    per-thread partial sums in registers (a GEMM until round 2: see gb2)
//     delta1 = D1 (1-h1^2)
//     G1^T += delta1^T [f | 1]                  UMMA  M=128 (64 units used), N=16, K=128 rows: dW1^T, ones column -> db1
// G2 / G1^T accumulate in tensor memory for the whole kernel and are read back once.
//
// Operand layouts.
//   * "row" operands of U2 / D1 (A[m=row][k=unit]) are written by their owning thread straight into TMEM
//     (tcgen05.st, lane = row): no shared-memory traffic at all for them.  B = W2 / W2^T: K-major, no swizzle,
//     element (n, k) at (k/4)*LBO + n*16 + (k%4)*4 bytes, SBO = 128.
//   * "K = row" operands (h1, delta2, delta1 as [unit][row]): a thread owns one ROW, i.e. one k, and 16 consecutive
//     units, so they are MN-major.  tf32 accepts MN-major only with SWIZZLE_128B_BASE32B (layout type 1; with
//     SWIZZLE_NONE the accumulator silently stays zero).  The hardware's layout was probed
//     (scripts/exp/mn_major_probe.cu): atoms of 4 rows x 128 bytes, the 32-byte chunk index inside a row XORed with
//     row % 4.  The owning thread therefore stores two 32-byte chunks per buffer straight from registers: no
//     transposition, no shuffles (the previous version spent a third of its instructions on 4x4 quad transposes).
//   * the feature operand [f|1] (N = 16) stays K-major and goes through a small [feature][row] staging array.
#include <cstdlib>

#include "nqmc_internal.cuh"

namespace {

constexpr int H = 64;
constexpr int NQ = 4;                 // threads per row
constexpr int TCW = 128 * NQ;         // worker threads
// 16 worker warps + one warpgroup holding the MMA issuer warp and three warps that only take part in the register
// split.  20 warps get 96 registers each at launch (61,440 in the CTA's pool); setmaxnreg then moves the issuer
// warpgroup to 32 and the four worker warpgroups to 112 -- the pool is per CTA, so 512 x 112 + 128 x 32 must not
// exceed what was allotted at launch (asking for more blocks forever).  Alternatives measured: a plain 17-warp CTA
// caps every thread at 96 (spills); 16 warps with the issue duty folded into a worker warp stall that warp on the
// UMMA queue and everybody else on the next barrier.  setmaxnreg needs whole warpgroups and one register budget per
// code path, hence the role split.
constexpr int TCT = TCW + 128;
constexpr int ISSUER = TCW / 32;      // warp 16
constexpr int NSYNC = TCT;            // threads meeting at the tile barriers: workers + the four issuer warps
constexpr int ROWS = 128;
constexpr int UPT = H / NQ;           // hidden units / output columns per thread = 16
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t C_U2 = 0, C_D1 = 64, C_G1 = 192, C_F1 = 224, C_AH = 256, C_AL = 320, C_G2 = 384;   // C_F1: layer-1 features, hi 16 + lo 16 columns   // G2: 128 columns (hi-part | lo-part products)
// K = rows operands of the weight-gradient GEMMs (h1, delta2, delta1): MN-major, SWIZZLE_128B_BASE32B.  Layout as
// probed on B200 (scripts/exp/mn_major_probe.cu; tf32 MN-major with SWIZZLE_NONE silently yields zeros):
//   byte(unit, row) = (unit/32) LBO + (row/4) SBO + (row%4) 128 + (((unit%32)/8) ^ (row%4)) 32 + (unit%8) 4
// i.e. a thread that owns 16 consecutive units of its row stores two 32-byte chunks straight from registers -- no
// transposition in software.  SBO must stay 512 (a padded SBO faults as misaligned), so the two row groups a
// quarter warp touches share banks (2-way conflict on these stores).
constexpr int MN_SBO = 512;                 // 4 rows x 128 bytes
constexpr int MN_LBO = (128 / 4) * MN_SBO;  // one atom column: 32 units x 128 rows = 16 KB
constexpr int MN_BUF = (H / 32) * MN_LBO;   // 64 units: 32 KB per hi / lo buffer
constexpr int FS_LD = 128 + 4;        // leading dimension of the [feature][row] staging array (float4 reads, stride 4 banks)
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int NRG = ROWS / 4;         // 32 row groups (granule = 4 consecutive rows of one unit)
constexpr int WB = (H / 4) * H * 16;  // one W2 copy (hi or lo)                                                   = 16384 B

struct SmemMap {      // byte offsets, computed from F at run time (identical on host and device)
  int hM_hi, hM_lo, dM_hi, dM_lo, fT_hi, fT_lo, wa_hi, wa_lo, wb_hi, wb_lo, w1_hi, w1_lo, fstage, pf, misc, total;
  int n_fu, rgf;      // feature units incl. ones, padded to 4 ; row-group stride of the (K-major) feature buffer
  int n_g1;           // N of the [f|1] MMAs (a multiple of 8): columns >= n_fu read into the next row group, ignored
};
constexpr int K1 = 16;                // K of the layer-1 UMMAs: x, y, z, 1, (d_a, exp(-d_a)) pairs, zero padding (A <= 4 ... 6)
constexpr int W1B = (K1 / 4) * H * 16; // one [W1; b1] operand copy (hi or lo): 4 KB
__host__ __device__ inline SmemMap make_map(int F, bool l1tc = false) {
  SmemMap m;
  m.n_fu = 4 * ((F + 1 + 3) / 4);
  m.n_g1 = 8 * ((F + 1 + 7) / 8);
  m.rgf = m.n_fu * 16;
  int o = 0;
  m.hM_hi = o; o += MN_BUF;
  m.hM_lo = o; o += MN_BUF;
  m.dM_hi = o; o += MN_BUF;
  m.dM_lo = o; o += MN_BUF;
  m.wa_hi = o; o += WB;
  m.wa_lo = o; o += WB;
  m.wb_hi = o; o += WB;
  m.wb_lo = o; o += WB;
  m.fT_hi = o; o += NRG * m.rgf;
  m.fT_lo = o; o += NRG * m.rgf;
  m.w1_hi = o; o += l1tc ? W1B : 0;
  m.w1_lo = o; o += l1tc ? W1B : 0;
  m.fstage = o; o += m.n_fu * FS_LD * 4;
  m.pf = o; o += 5 * ROWS * 4;        // next tile: x, y, z, weight, inverse-matrix entry per row (cp.async targets)
  m.misc = o; o += (F * H + 3 * H + 2 * H + 8) * 4 + 80;
  m.total = o;
  return m;
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {   // MN-major, SWIZZLE_128B_BASE32B (layout type 1)
  return umma_desc(smem_addr, MN_LBO, MN_SBO) | (1ull << 61);
}
// D=f32, A=B=tf32, both K-major, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
// G2 = h1^T delta2: A and B MN-major, K = rows.  Round 2a ran it as M = 64 (the 64 h1 units; an M = 64 accumulator row u lives
// in TMEM lane (u / 16) * 32 + u % 16, probed: scripts/exp/m64_probe.cu -- G1^T still does) with B' = [delta2_hi | delta2_lo]
// as N = 128 (the lo buffer follows the hi buffer: two more 32-unit atom columns): a_hi b_hi in columns 0-63, a_hi b_lo in
// columns 64-127, plus a second UMMA a_lo b_hi.  Round 2b: the same concatenation on the A side, A' = [h1_hi ; h1_lo] as
// M = 128 (hM_lo follows hM_hi): ONE UMMA per k-step yields a_hi b_hi | a_hi b_lo in lanes 0-63 and a_lo b_hi | a_lo b_lo in
// lanes 64-127 (accumulator row = lane for M = 128).  An M = 64 UMMA occupies the tensor pipe as long as an M = 128 one, so
// G2 drops from 16 x (N128 + N64) to 16 x N128, and delta2_hi is read from shared memory once instead of twice: reverse
// kernel 153.5 -> 148.2 us.
constexpr uint32_t IDESC_MN128M = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// A and B from shared memory
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                        uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, l;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 l, %5, 0;\n\t"
      "@l tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// A from tensor memory, B from shared memory
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, l;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 l, %5, 0;\n\t"
      "@l tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tile_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NSYNC) : "memory"); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint32_t bar, uint32_t leader) {
  asm volatile("{\n\t.reg .pred l;\n\tsetp.ne.b32 l, %1, 0;\n\t"
               "@l tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(leader)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  lo = a - hi;
}

// 16 values of this thread's row -> tf32 hi / lo in the TMEM A-operand columns (lane = row)
__device__ __forceinline__ void stage_rows_tmem(uint32_t t_hi, uint32_t t_lo, const float (&v)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float h, l;
    split_tf32(v[j], h, l);
    hi[j] = __float_as_uint(h);
    lo[j] = __float_as_uint(l);
  }
  tmem_st16(t_hi, hi);
  tmem_st16(t_lo, lo);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 16 values of this thread's row (units c0 .. c0+15) -> tf32 hi / lo in an MN-major swizzled operand buffer:
// two 32-byte chunks per buffer at byte offsets off0 / off1 (see the layout formula above), straight from registers.
// Bank conflicts: the 8 lanes of a quarter warp own 8 consecutive rows, i.e. every row % 4 twice, and the chunk index is
// XORed with row % 4 only -- two lanes per 32-byte chunk slot.  `flip` = bit 2 of the row: lanes with flip set write the
// upper 16 bytes of their chunk first and the lower ones second, so each STS.128 of a quarter warp covers all 32 banks.
__device__ __forceinline__ void stage_mn(uint8_t* hi, uint8_t* lo, const int off0, const int off1, const float (&v)[16], const bool flip) {
  const int of_first = flip ? 16 : 0, of_second = 16 - of_first;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint8_t* ph = hi + (c == 0 ? off0 : off1);
    uint8_t* pl = lo + (c == 0 ? off0 : off1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {                 // h = 0: this lane's first store, h = 1: its second
      const int a = 8 * c, b = 8 * c + 4;
      const bool upper = (h == 1) != flip;        // which half of the chunk this store carries
      const float x0 = upper ? v[b + 0] : v[a + 0], x1 = upper ? v[b + 1] : v[a + 1];
      const float x2 = upper ? v[b + 2] : v[a + 2], x3 = upper ? v[b + 3] : v[a + 3];
      float4 h4, l4;
      split_tf32(x0, h4.x, l4.x);
      split_tf32(x1, h4.y, l4.y);
      split_tf32(x2, h4.z, l4.z);
      split_tf32(x3, h4.w, l4.w);
      const int of = h == 0 ? of_first : of_second;
      *reinterpret_cast<float4*>(ph + of) = h4;
      *reinterpret_cast<float4*>(pl + of) = l4;
    }
  }
}

// NA = number of nuclei when > 0 (feature loops fully unrolled), 0 = run-time count
template <int NA>
__global__ void __launch_bounds__(TCT, 1)
    orb_bwd_tc_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r,
                      const float* __restrict__ inv, const float* __restrict__ wgt, const double* __restrict__ hdr,
                      const float* __restrict__ e_l, const int64_t W, float* __restrict__ part, const int n_chunks,
                      const int p_mlp) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int A = NA > 0 ? NA : s.A, F = 3 + 2 * A;
  // Layer 1 on the tensor cores (A <= 4: the operand copies fit next to everything else in shared memory): the row's four
  // threads write [x, y, z, 1, d_a, exp(-d_a) ...] as tf32 hi / lo straight into TMEM (one nucleus each), six UMMAs form
  // [f|1] [W1; b1] in the (still free) U2 columns, and a thread gets its 16 pre-activations with one tcgen05.ld instead of
  // 11 x 16 FFMAs behind 44 broadcast LDS.128 -- measured with clock64 stamps, the FP32 layer 1 was 2.7 k of a tile's
  // 9.3 k cycles.  The wait for G2(t-1) and the delta1 staging run while those UMMAs are in flight.
  constexpr bool L1TC = NA >= 1 && NA <= 4;
  const SmemMap mp = make_map(F, L1TC);
  float* misc = reinterpret_cast<float*>(smem + mp.misc);
  float* W1s = misc;                 // [F][H]
  float* b1s = W1s + F * H;
  float* b2s = b1s + H;
  float* W3s = b2s + H;
  float* redS = W3s + H;             // [2H + 8]: dW3[H], db3 (+7 pad), db2[H]
  uint64_t* bars = reinterpret_cast<uint64_t*>(redS + 2 * H + 8);   // 5 mbarriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* fstage = reinterpret_cast<float*>(smem + mp.fstage);   // [n_fu][ROWS]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (ROWS - 1), quarter = (tid >> 7) & (NQ - 1);
  const bool worker = tid < TCW;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];

  // ---- setup ----------------------------------------------------------------------------------------------
  for (int idx = tid; idx < mp.wa_hi / 16; idx += TCT) reinterpret_cast<float4*>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = tid; idx < (mp.fstage - mp.fT_hi) / 16; idx += TCT)
    reinterpret_cast<float4*>(smem + mp.fT_hi)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = tid; idx < mp.n_fu * FS_LD; idx += TCT) fstage[idx] = 0.f;
  for (int idx = tid; idx < F * H; idx += TCT) W1s[idx] = theta[off.W1 + idx];
  for (int q = tid; q < H; q += TCT) {
    b1s[q] = theta[off.b1 + q];
    b2s[q] = theta[off.b2 + q];
    W3s[q] = theta[off.W3 + q];
  }
  for (int q = tid; q < 2 * H + 8; q += TCT) redS[q] = 0.f;
  if constexpr (L1TC) {
    for (int idx = tid; idx < K1 * H; idx += TCT) {
      const int k = idx / H, n = idx % H;   // operand feature order: x, y, z, 1, then (d_a, exp(-d_a)) pairs
      float v = 0.f;
      if (k < 3) v = theta[off.W1 + k * H + n];
      else if (k == 3) v = theta[off.b1 + n];
      else if (k < 4 + 2 * A) v = theta[off.W1 + (((k - 4) & 1) ? 3 + A + ((k - 4) >> 1) : 3 + ((k - 4) >> 1)) * H + n];
      float hi, lo;
      split_tf32(v, hi, lo);
      const int o = (k >> 2) * (H * 4) + n * 4 + (k & 3);
      reinterpret_cast<float*>(smem + mp.w1_hi)[o] = hi;
      reinterpret_cast<float*>(smem + mp.w1_lo)[o] = lo;
    }
  }
  {
    // all loads of a thread in flight at once: with a cold L2 each dependent round trip costs ~1 us
    constexpr int NL = H * H / TCW;
    static_assert(NL * TCW == H * H, "W2 staging assumes the worker count divides H*H");
    float wv[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) wv[i] = worker ? theta[off.W2 + tid + i * TCW] : 0.f;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      if (!worker) break;
      const int idx = tid + i * TCW;
      const int k = idx / H, j = idx % H;
      float hi, lo;
      split_tf32(wv[i], hi, lo);
      // copy a: B[n=j][kk=k] (for U2 = h1 W2)      -> [k/4][j][k%4]
      const int oa = (k >> 2) * (H * 4) + j * 4 + (k & 3);
      reinterpret_cast<float*>(smem + mp.wa_hi)[oa] = hi;
      reinterpret_cast<float*>(smem + mp.wa_lo)[oa] = lo;
      // copy b: B[n=k][kk=j] (for D1 = delta2 W2^T) -> [j/4][k][j%4]
      const int ob = (j >> 2) * (H * 4) + k * 4 + (j & 3);
      reinterpret_cast<float*>(smem + mp.wb_hi)[ob] = hi;
      reinterpret_cast<float*>(smem + mp.wb_lo)[ob] = lo;
    }
  }
  __syncthreads();
  if (tid == 0) {
    for (int b = 0; b < 5; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[b])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_u = smem_u32(&bars[0]), bar_d = smem_u32(&bars[1]), bar_g2 = smem_u32(&bars[2]), bar_g = smem_u32(&bars[3]);
  const uint32_t bar_l1 = smem_u32(&bars[4]);
  uint32_t ph_u = 0, ph_d = 0, ph_g2 = 0, ph_g = 0, ph_l1 = 0;
  nq_pdl_wait();                              // everything above read theta only; header, E_L, inverses come from predecessors
  float centerf = 0.f;
  if (hdr != nullptr) centerf = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));

  float gW3[UPT];
#pragma unroll
  for (int j = 0; j < UPT; ++j) gW3[j] = 0.f;
  float gb3 = 0.f;
  // db2 = sum over rows of delta2: per-thread partial sums, reduced once at the end like dW3.  (Until round 2 this was a
  // fourth tensor-core GEMM, delta2^T [f|1], whose 48 UMMAs re-read the whole delta2 operand from shared memory for one
  // useful output column; the kernel is shared-memory-bandwidth-bound -- LSU + tensor-core wavefronts at 90 % of peak.)
  float gb2[UPT];
#pragma unroll
  for (int j = 0; j < UPT; ++j) gb2[j] = 0.f;
  const int c0 = UPT * quarter;                     // first hidden unit / output column of this thread
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  // byte offsets of this thread's two 32-byte chunks (units c0..c0+7, c0+8..c0+15 of its row) in an MN-major buffer
  const int mn_row = (quarter >> 1) * MN_LBO + (row >> 2) * MN_SBO + (row & 3) * 128, mn_cb = 2 * (quarter & 1);
  const int mn_off0 = mn_row + ((mn_cb ^ (row & 3)) << 5), mn_off1 = mn_row + (((mn_cb + 1) ^ (row & 3)) << 5);
  const bool mn_flip = (row & 4) != 0;
  uint8_t* hM_hi = smem + mp.hM_hi;
  uint8_t* hM_lo = smem + mp.hM_lo;
  uint8_t* dM_hi = smem + mp.dM_hi;
  uint8_t* dM_lo = smem + mp.dM_lo;

  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;
  // [f|1] products: D=f32, A=B=tf32, A MN-major (delta), B K-major (features), M = 64 delta units, N = n_g1
  const uint32_t idesc_g1 =
      (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(mp.n_g1 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
  // Positions and reverse seed of one tile row.  The next tile's are fetched while the current one computes: the
  // quarter-0 thread of each row issues five 4-byte cp.async into shared memory (no registers held across the tile --
  // register prefetch was spilled by ptxas, and a spilled load stalls on the spot), waits for them at the end of P3
  // and the tile barrier there publishes them to the row's four threads.
  float* pf = reinterpret_cast<float*>(smem + mp.pf);            // [5][ROWS]
  ItemIter cur(chunk, n_chunks, ns), nxt = cur;                 // this tile / the next tile to request (in order)
  auto fetch = [&]() {
    const int i = nxt.e;
    int64_t w = (int64_t)nxt.wb * ROWS + row;
    nxt.next();
    if (w >= W) w = W - 1;
    const int e = elec0 + i;
    const float* src[5] = {r + (int64_t)(3 * e + 0) * W + w, r + (int64_t)(3 * e + 1) * W + w, r + (int64_t)(3 * e + 2) * W + w,
                           wgt != nullptr ? wgt + w : (e_l != nullptr ? e_l + w : nullptr),
                           inv + (int64_t)ent_index(s, spin, i, korb) * W + w};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      if (src[q] != nullptr)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(pf + q * ROWS + row)), "l"(src[q]) : "memory");
      else
        pf[q * ROWS + row] = 1.0f;
    }
  };
  if (worker && chunk < n_items) {
    if (quarter == 0) {
      fetch();
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    asm volatile("bar.sync 2, %0;" ::"n"(TCW) : "memory");      // workers only
  }
  // Pipeline (one tile = 3 CTA barriers; delta1 of tile t stays in registers until tile t+1's first phase, so the
  // wait for G2(t) -- which still reads the buffer delta1^T goes into -- hides behind layer 1 of tile t+1):
  //   P1  layer 1(t) -> A(h1) in TMEM ; delta1(t-1) -> dM                   | issue U2(t), G1^T(t-1) = delta1^T [f|1]
  //   P2  h1(t) -> hM ; wait U2 ; h2, dW3, delta2 -> A(delta2) in TMEM        | issue D1(t)
  //   P3  wait G1^T(t-1) ; delta2(t) -> dM ; features(t) -> fT              | issue G2(t) = h1^T delta2
  //   P4  wait D1 ; delta1(t) = D1 (1 - h1^2)  (registers)
  // The ones column of [f|1] yields db1 (from G1^T); db2 is summed in registers.
  float d1[UPT];
#pragma unroll
  for (int j = 0; j < UPT; ++j) d1[j] = 0.f;
  // dst += (delta in dM)^T [f|1]   (M = delta unit, N = feature; 48 UMMAs of N = n_g1: 8 cycles each)
  auto issue_df = [&](const uint32_t dst, const uint32_t leader, const bool first) {
    const uint64_t ah = umma_desc_mn(sbase + mp.dM_hi), al = umma_desc_mn(sbase + mp.dM_lo);
    const uint64_t bh = umma_desc(sbase + mp.fT_hi, mp.rgf, 128), bl = umma_desc(sbase + mp.fT_lo, mp.rgf, 128);
    const uint64_t da = (uint64_t)((2 * MN_SBO) >> 4), db = (uint64_t)((2 * mp.rgf) >> 4);
    uint64_t ao = 0, bo = 0;
#pragma unroll 1
    for (int ks = 0; ks < ROWS / 8; ++ks, ao += da, bo += db) {   // not unrolled: the issuer warp lives on 32 registers
      const uint32_t acc = (first && ks == 0) ? 0u : 1u;
      umma_ss(tmem_base + dst, ah + ao, bh + bo, idesc_g1, acc, leader);
      umma_ss(tmem_base + dst, al + ao, bh + bo, idesc_g1, 1u, leader);
      umma_ss(tmem_base + dst, ah + ao, bl + bo, idesc_g1, 1u, leader);
    }
  };
  if (!worker) {
    // =============================== issuer warpgroup ========================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    // The four warps of this warpgroup share the issue work (one thread sustains only about one tcgen05.mma per 100
    // cycles, scripts/exp/mma_rate.cu): warp 0: U2, D1 (the critical chain) ; 1: G1^T ; 2: G2 ; 3: idle.
    const int iw = warp - ISSUER;
    const uint32_t leader = elect_one();
    int64_t n_done = 0;
    for (int64_t item = chunk; item < n_items; item += n_chunks, ++n_done) {
      if constexpr (L1TC) {
        tile_barrier();                                  // P0 done: features of this tile in TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (iw == 0) {                                   // Z1 = [f|1] [W1; b1] into the U2 columns (free until U2 is issued)
          const uint64_t bh = umma_desc(sbase + mp.w1_hi, H * 16, 128), bl = umma_desc(sbase + mp.w1_lo, H * 16, 128);
#pragma unroll
          for (int ks = 0; ks < K1 / 8; ++ks) {
            const uint64_t bo = (uint64_t)((ks * 2 * H * 16) >> 4);
            const uint32_t ah = tmem_base + C_F1 + 8 * ks, al = tmem_base + C_F1 + K1 + 8 * ks;
            umma_ts(tmem_base + C_U2, ah, bh + bo, ks == 0 ? 0u : 1u, leader);
            umma_ts(tmem_base + C_U2, al, bh + bo, 1u, leader);
            umma_ts(tmem_base + C_U2, ah, bl + bo, 1u, leader);
          }
          umma_commit(bar_l1, leader);
        }
      }
      tile_barrier();                                    // P1 done: A(h1) in TMEM, delta1(t-1) in dM
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (iw == 0) {
        const uint64_t bh = umma_desc(sbase + mp.wa_hi, H * 16, 128), bl = umma_desc(sbase + mp.wa_lo, H * 16, 128);
#pragma unroll 1
        for (int ks = 0; ks < H / 8; ++ks) {             // U2 = h1 W2 (A from tensor memory)
          const uint64_t bo = (uint64_t)((ks * 2 * H * 16) >> 4);
          const uint32_t ah = tmem_base + C_AH + 8 * ks, al = tmem_base + C_AL + 8 * ks;
          umma_ts(tmem_base + C_U2, ah, bh + bo, ks == 0 ? 0u : 1u, leader);
          umma_ts(tmem_base + C_U2, al, bh + bo, 1u, leader);
          umma_ts(tmem_base + C_U2, ah, bl + bo, 1u, leader);
        }
        umma_commit(bar_u, leader);
      } else if (iw == 1 && n_done > 0) {
        issue_df(C_G1, leader, n_done == 1);             // G1^T(t-1) += delta1^T [f|1]
        umma_commit(bar_g, leader);
      }
      tile_barrier();                                    // P2 done: A(delta2) in TMEM
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (iw == 0) {
        const uint64_t bh = umma_desc(sbase + mp.wb_hi, H * 16, 128), bl = umma_desc(sbase + mp.wb_lo, H * 16, 128);
#pragma unroll 1
        for (int ks = 0; ks < H / 8; ++ks) {             // D1 = delta2 W2^T (A from tensor memory)
          const uint64_t bo = (uint64_t)((ks * 2 * H * 16) >> 4);
          const uint32_t ah = tmem_base + C_AH + 8 * ks, al = tmem_base + C_AL + 8 * ks;
          umma_ts(tmem_base + C_D1, ah, bh + bo, ks == 0 ? 0u : 1u, leader);
          umma_ts(tmem_base + C_D1, al, bh + bo, 1u, leader);
          umma_ts(tmem_base + C_D1, ah, bl + bo, 1u, leader);
        }
        umma_commit(bar_d, leader);
      }
      tile_barrier();                                    // P3 done: h1 in hM, delta2 in dM, features in fT
      if (iw == 2) {
        const uint64_t ah = umma_desc_mn(sbase + mp.hM_hi);   // M = 128 reads on into hM_lo, N = 128 into dM_lo (contiguous: SmemMap)
        const uint64_t bh = umma_desc_mn(sbase + mp.dM_hi);
#pragma unroll 1
        for (int ks = 0; ks < ROWS / 8; ++ks) {          // G2 += h1^T delta2   (both operands MN-major, K = 8 rows per UMMA)
          const uint64_t ko = (uint64_t)((ks * 2 * MN_SBO) >> 4);
          const uint32_t acc = (n_done == 0 && ks == 0) ? 0u : 1u;
          umma_ss(tmem_base + C_G2, ah + ko, bh + ko, IDESC_MN128M, acc, leader);  // [a_hi ; a_lo] [b_hi | b_lo]
        }
        umma_commit(bar_g2, leader);
      }
    }
    if (n_done > 0) {                                    // drain: G1^T of the last tile
      tile_barrier();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (iw == 1) {
        issue_df(C_G1, leader, n_done == 1);
        umma_commit(bar_g, leader);
      }
    }
    tile_barrier();                                      // workers have read every accumulator
    return;
  }
  // =============================== worker warps ================================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
  if constexpr (L1TC) {                                // the padding feature columns (4 + 2A .. 15, hi and lo) must read as zero
    if (worker && quarter == 0) {
      const uint32_t z[16] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st16(tmem_base + lane_base + C_F1, z);
      tmem_st16(tmem_base + lane_base + C_F1 + 16u, z);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  int64_t n_done = 0;                                  // tiles this CTA has completed
  for (int64_t item = chunk; item < n_items; item += n_chunks, ++n_done, cur.next()) {
    float h1[UPT], d2[UPT], seed = 0.f;
    // ======== P1 ========
    if (worker) {
      const float x = pf[0 * ROWS + row], y = pf[1 * ROWS + row], z = pf[2 * ROWS + row];
      {
        const float nwt = pf[3 * ROWS + row], ninv = pf[4 * ROWS + row];
        float wt = nwt;
        if (wgt == nullptr && e_l != nullptr) wt = isfinite(nwt) ? nwt - centerf : 0.f;
        seed = wt * ninv;
        if ((int64_t)cur.wb * ROWS + row >= W || !isfinite(seed)) seed = 0.f;
      }
      if constexpr (L1TC) {
        // P0: this thread's share of the row's features -> TMEM (operand order x, y, z, 1, (d_a, e_a) pairs) and -> fstage
        // (reference order, the B operand of G1)
        const uint32_t fa = tmem_base + lane_base + C_F1;
        if (quarter == 0) {
          float hx, lx, hy, ly, hz, lz;
          split_tf32(x, hx, lx);
          split_tf32(y, hy, ly);
          split_tf32(z, hz, lz);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(fa), "r"(__float_as_uint(hx)),
                       "r"(__float_as_uint(hy)), "r"(__float_as_uint(hz)), "r"(0x3f800000u) : "memory");
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(fa + K1), "r"(__float_as_uint(lx)),
                       "r"(__float_as_uint(ly)), "r"(__float_as_uint(lz)), "r"(0u) : "memory");
          fstage[0 * FS_LD + row] = x;
          fstage[1 * FS_LD + row] = y;
          fstage[2 * FS_LD + row] = z;
          fstage[F * FS_LD + row] = 1.0f;                // ones unit -> db1
        }
        if (quarter < A) {                               // A <= 4: nucleus `quarter`
          const int a = quarter;
          const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
          const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
          const float ex = __expf(-dd);
          float hd, ld, he, le;
          split_tf32(dd, hd, ld);
          split_tf32(ex, he, le);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(fa + (uint32_t)(4 + 2 * a)),
                       "r"(__float_as_uint(hd)), "r"(__float_as_uint(he)) : "memory");
          asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(fa + (uint32_t)(K1 + 4 + 2 * a)),
                       "r"(__float_as_uint(ld)), "r"(__float_as_uint(le)) : "memory");
          fstage[(3 + a) * FS_LD + row] = dd;
          fstage[(3 + A + a) * FS_LD + row] = ex;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tile_barrier();                                  // issuer: Z1 = [f|1] [W1; b1]
        if (n_done > 0) {                                // while those UMMAs are in flight:
          mbar_wait(bar_g2, ph_g2);                      // G2(t-1) read delta2^T from the buffer delta1^T goes into
          ph_g2 ^= 1;
          stage_mn(dM_hi, dM_lo, mn_off0, mn_off1, d1, mn_flip);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        mbar_wait(bar_l1, ph_l1);
        ph_l1 ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld16(tmem_base + lane_base + C_U2 + (uint32_t)c0, h1);
#pragma unroll
        for (int j = 0; j < UPT; ++j) h1[j] = nq_tanh(h1[j]);
        stage_rows_tmem(tmem_base + lane_base + C_AH + (uint32_t)c0, tmem_base + lane_base + C_AL + (uint32_t)c0, h1);
      } else {
#pragma unroll
        for (int j = 0; j < UPT; ++j) h1[j] = b1s[c0 + j];
        const float pc[3] = {x, y, z};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
#pragma unroll
          for (int v = 0; v < UPT / 4; ++v) {
            const float4 wv = *reinterpret_cast<const float4*>(W1s + d * H + c0 + 4 * v);
            h1[4 * v + 0] = fmaf(pc[d], wv.x, h1[4 * v + 0]);
            h1[4 * v + 1] = fmaf(pc[d], wv.y, h1[4 * v + 1]);
            h1[4 * v + 2] = fmaf(pc[d], wv.z, h1[4 * v + 2]);
            h1[4 * v + 3] = fmaf(pc[d], wv.w, h1[4 * v + 3]);
          }
        }
        if (quarter == 0) {
          fstage[0 * FS_LD + row] = x;
          fstage[1 * FS_LD + row] = y;
          fstage[2 * FS_LD + row] = z;
          fstage[F * FS_LD + row] = 1.0f;                  // ones unit -> db1
        }
#pragma unroll
        for (int a = 0; a < A; ++a) {
          const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
          const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
          const float ex = __expf(-dd);
          if (quarter == 0) {
            fstage[(3 + a) * FS_LD + row] = dd;
            fstage[(3 + A + a) * FS_LD + row] = ex;
          }
#pragma unroll
          for (int v = 0; v < UPT / 4; ++v) {
            const float4 wd = *reinterpret_cast<const float4*>(W1s + (3 + a) * H + c0 + 4 * v);
            const float4 we = *reinterpret_cast<const float4*>(W1s + (3 + A + a) * H + c0 + 4 * v);
            h1[4 * v + 0] = fmaf(dd, wd.x, fmaf(ex, we.x, h1[4 * v + 0]));
            h1[4 * v + 1] = fmaf(dd, wd.y, fmaf(ex, we.y, h1[4 * v + 1]));
            h1[4 * v + 2] = fmaf(dd, wd.z, fmaf(ex, we.z, h1[4 * v + 2]));
            h1[4 * v + 3] = fmaf(dd, wd.w, fmaf(ex, we.w, h1[4 * v + 3]));
          }
        }
#pragma unroll
        for (int j = 0; j < UPT; ++j) h1[j] = nq_tanh(h1[j]);
        // A operand of U2: straight into tensor memory (its previous reader, D1 of the last tile, has completed)
        stage_rows_tmem(tmem_base + lane_base + C_AH + (uint32_t)c0, tmem_base + lane_base + C_AL + (uint32_t)c0, h1);
        if (n_done > 0) {
          mbar_wait(bar_g2, ph_g2);                      // G2(t-1) read delta2^T from the buffer delta1^T goes into
          ph_g2 ^= 1;
          stage_mn(dM_hi, dM_lo, mn_off0, mn_off1, d1, mn_flip);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tile_barrier();
    // ======== P2 ========
    if (worker) {
      stage_mn(hM_hi, hM_lo, mn_off0, mn_off1, h1, mn_flip);         // G2(t-1), the reader of this buffer, was waited for in P1
      mbar_wait(bar_u, ph_u);
      ph_u ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[16];
      tmem_ld16(tmem_base + lane_base + C_U2 + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        const float h = nq_tanh(v[j] + b2s[c0 + j]);
        gW3[j] = fmaf(seed, h, gW3[j]);
        v[j] = seed * W3s[c0 + j] * fmaf(-h, h, 1.0f);
      }
#pragma unroll
      for (int j = 0; j < UPT; ++j) gb2[j] += v[j];
      if (quarter == 0) gb3 += seed;
      stage_rows_tmem(tmem_base + lane_base + C_AH + (uint32_t)c0, tmem_base + lane_base + C_AL + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < UPT; ++j) d2[j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tile_barrier();
    // ======== P3 ========
    if (worker) {
      const bool more = item + n_chunks < n_items;
      if (more && quarter == 0) fetch();         // next tile's positions / seed (read last in P1, two barriers ago)
      if (n_done > 0) { mbar_wait(bar_g, ph_g); ph_g ^= 1; }   // G1(t-1) reads delta1^T(t-1) and the features of tile t-1
      stage_mn(dM_hi, dM_lo, mn_off0, mn_off1, d2, mn_flip);
      // granule (feature unit f, row group g) <- fstage[f][4g .. 4g+3]
      for (int idx = tid; idx < mp.n_fu * NRG; idx += TCW) {
        const int g = idx / mp.n_fu, f = idx % mp.n_fu;     // consecutive lanes: consecutive 16-byte granules
        const float4 fv = *reinterpret_cast<const float4*>(fstage + f * FS_LD + 4 * g);
        float4 h4, l4;
        split_tf32(fv.x, h4.x, l4.x);
        split_tf32(fv.y, h4.y, l4.y);
        split_tf32(fv.z, h4.z, l4.z);
        split_tf32(fv.w, h4.w, l4.w);
        *reinterpret_cast<float4*>(smem + mp.fT_hi + g * mp.rgf + f * 16) = h4;
        *reinterpret_cast<float4*>(smem + mp.fT_lo + g * mp.rgf + f * 16) = l4;
      }
      if (more && quarter == 0) asm volatile("cp.async.wait_all;" ::: "memory");   // the barrier below publishes the prefetch
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tile_barrier();
    // ======== P4 ========
    if (worker) {
      mbar_wait(bar_d, ph_d);
      ph_d ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_ld16(tmem_base + lane_base + C_D1 + (uint32_t)c0, d1);
#pragma unroll
      for (int j = 0; j < UPT; ++j) d1[j] *= fmaf(-h1[j], h1[j], 1.0f);
    }
  }
  // ---- drain: delta1^T of the last tile, its G1 -----------------------------------------------------------------------
  const bool any = n_done > 0;
  if (any) {
    if (worker) {
      mbar_wait(bar_g2, ph_g2);
      ph_g2 ^= 1;
      stage_mn(dM_hi, dM_lo, mn_off0, mn_off1, d1, mn_flip);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tile_barrier();
  }

  // ---- flush: G2 -> dW2 (+db2), G1^T -> dW1 (+db1), register partials -> dW3, db3 ------------------------------------
  float* slab = part + (int64_t)blockIdx.x * p_mlp;
  const int o_b1 = 0, o_W1 = H, o_b2 = H + F * H, o_W2 = o_b2 + H, o_b3 = o_W2 + H * H, o_W3 = o_b3 + 1;
  if (worker) {
    float v2[16];
    if (any) {
      mbar_wait(bar_g, ph_g);                            // the last commit: every MMA of this CTA has completed
      ph_g ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v2l[16];
      tmem_ld16(tmem_base + lane_base + C_G2 + (uint32_t)c0, v2);   // lane = h1 unit
      tmem_ld16(tmem_base + lane_base + C_G2 + (uint32_t)(H + c0), v2l);   // the a_hi b_lo products
#pragma unroll
      for (int j = 0; j < UPT; ++j) v2[j] += v2l[j];
    } else {
#pragma unroll
      for (int j = 0; j < UPT; ++j) v2[j] = 0.f;          // this CTA had no tile: the slab must still be defined
    }
    // G2 (M = 128): lane r < 64 holds the h1_hi products of unit r, lane 64 + r the h1_lo products: the upper half hands its
    // sums to the lower half through the (now idle) hM buffer
    {
      float* xch = reinterpret_cast<float*>(hM_hi);      // [64][H]
      if (row >= H) {
#pragma unroll
        for (int j = 0; j < UPT; ++j) xch[(row - H) * H + c0 + j] = v2[j];
      }
      asm volatile("bar.sync 2, %0;" ::"n"(TCW) : "memory");      // workers only
      if (row < H) {
#pragma unroll
        for (int j = 0; j < UPT; ++j) slab[o_W2 + row * H + c0 + j] = v2[j] + xch[row * H + c0 + j];
      }
    }
    // G1^T, M = 64 accumulators: lanes 0-15, 32-47, 64-79, 96-111 hold units 0-15, 16-31, 32-47, 48-63
    const bool lane_ok = (row & 16) == 0;
    const int unit = (row >> 5) * 16 + (row & 15);
    // G1^T: lane = delta1 unit, column = feature (column F: ones -> db1); 8 columns per read,
    // column groups dealt over the quarters
    for (int cg = quarter; cg < mp.n_g1 / 8; cg += NQ) {
      float v1[8];
      if (any) {
        tmem_ld8(tmem_base + lane_base + C_G1 + (uint32_t)(8 * cg), v1);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v1[j] = 0.f;
      }
      if (lane_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int f = 8 * cg + j;
          if (f < F) slab[o_W1 + f * H + unit] = v1[j];
          else if (f == F) {
            slab[o_b1 + unit] = v1[j];
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < UPT; ++j) {                     // dW3 / db3: reduce the per-row partials over the 128 rows
      const float t = warp_sum(gW3[j]);
      if ((tid & 31) == 0) atomicAdd(&redS[c0 + j], t);
      const float t2 = warp_sum(gb2[j]);
      if ((tid & 31) == 0) atomicAdd(&redS[H + 8 + c0 + j], t2);
    }
    if (quarter == 0) {
      const float t = warp_sum(gb3);
      if ((tid & 31) == 0) atomicAdd(&redS[H], t);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  tile_barrier();
  for (int q = tid; q < H; q += TCW) {
    slab[o_W3 + q] = redS[q];
    slab[o_b2 + q] = redS[H + 8 + q];
  }
  if (tid == 0) slab[o_b3] = redS[H];
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

template <int NA>
int launch_t(nqmc_ctx* ctx, const SmemMap& mp, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
             int n_chunks, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(orb_bwd_tc_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured = true;
  }
  {
    ProfScope ps(ctx, NQ_PROF_BWD, st);
    if (ctx->prof_on)
      orb_bwd_tc_kernel<NA><<<n_chunks * s.n_mlp, TCT, mp.total, st>>>(s, ctx->theta, r, ctx->inv, wgt, hdr, e_l, W, ctx->part_mlp,
                                                                       n_chunks, ctx->p_mlp);
    else
      NQMC_CUDA(nq_launch_pdl_guarded(ctx, orb_bwd_tc_kernel<NA>, dim3(n_chunks * s.n_mlp), dim3(TCT), (size_t)mp.total, st, s, ctx->theta, r,
                              ctx->inv, wgt, hdr, e_l, W, ctx->part_mlp, n_chunks, ctx->p_mlp));
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

}  // namespace

int nq_launch_orb_bwd_tc(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                         int* n_chunks_out, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H != H) return NQMC_ERR_UNSUPPORTED;
  const SmemMap mp = make_map(s.F, s.A >= 1 && s.A <= 4);   // must match the kernel's L1TC (template NA = s.A for these)
  if (mp.total > SMEM_LIMIT || mp.n_g1 > 32) return NQMC_ERR_UNSUPPORTED;   // too many nuclei for this variant's smem plan
  if (const char* e = getenv("NQMC_TC_BWD"))
    if (atoi(e) == 0) return NQMC_ERR_UNSUPPORTED;
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  if (max_items >= (int64_t)1 << 31) return NQMC_ERR_UNSUPPORTED;   // ItemIter counts items in 32 bits
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks * s.n_mlp > ctx->n_cta_bwd) n_chunks = ctx->n_cta_bwd / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  *n_chunks_out = n_chunks;
  switch (s.A) {
    case 1: return launch_t<1>(ctx, mp, r, wgt, hdr, e_l, W, n_chunks, st);
    case 2: return launch_t<2>(ctx, mp, r, wgt, hdr, e_l, W, n_chunks, st);
    case 3: return launch_t<3>(ctx, mp, r, wgt, hdr, e_l, W, n_chunks, st);
    case 4: return launch_t<4>(ctx, mp, r, wgt, hdr, e_l, W, n_chunks, st);
    case 6: return launch_t<6>(ctx, mp, r, wgt, hdr, e_l, W, n_chunks, st);
    default: return launch_t<0>(ctx, mp, r, wgt, hdr, e_l, W, n_chunks, st);
  }
}
