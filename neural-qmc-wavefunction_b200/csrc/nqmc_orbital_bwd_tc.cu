// nqmc_orbital_bwd_tc.cu -- tensor-core (tcgen05 / TMEM, 3xTF32) reverse pass of the orbital networks for
// hidden width 64.  Same math, interface and slab layout as orb_bwd_kernel<64,...> (nqmc_orbital.cu), which
// stays as the FP32 FFMA parity anchor.
//
// Per 128-row tile (128 consecutive walkers, one electron of the network's spin), seed = weight * inv:
//     h1 = tanh(f W1 + b1)                      FP32 FFMA, thread-per-row (K = 3+2A is tiny)
//     U2 = h1 W2                                UMMA  M=128 rows, N=64, K=64    A: h1 in TENSOR MEMORY, B: W2 smem
//     h2 = tanh(U2 + b2); dW3 += seed h2; delta2 = seed W3 (1-h2^2)
//     D1 = delta2 W2^T                          UMMA  M=128 rows, N=64, K=64    A: delta2 in TMEM,      B: W2^T smem
//     G2 += [h1 | 1]^T delta2                   UMMA  M=128 (64 units + ones -> db2), N=64, K=128 rows  (smem x smem)
//     delta1 = D1 (1-h1^2)
//     G1 += [f | 1]^T delta1                    UMMA  M=128 (F features + ones -> db1), N=64, K=128 rows (smem x smem)
// G2 / G1 accumulate in tensor memory for the whole kernel and are read back once.
//
// Operand layouts.  tf32 operands may only be MN-major in a special swizzled format, so every shared-memory
// operand here is K-major, no swizzle: element (mn, k) at (k/4)*LBO + mn*16 + (k%4)*4 bytes, SBO = 128.
//   * "row" operands of U2 / D1 (A[m=row][k=unit]) are written by their owning thread straight into TMEM
//     (tcgen05.st, lane = row): no shared-memory traffic at all for them.
//   * "K = row" operands of G2 / G1 need 4 consecutive rows of one unit in a 16-byte granule: a thread owns
//     one row, so the 4 lanes of a quad transpose their 4x4 blocks with two butterfly shuffle stages
//     (4 SHFL per block) before the hi/lo split; features go through a small [feature][row] staging array.
#include <cstdlib>

#include "nqmc_internal.cuh"

namespace {

constexpr int H = 64;
constexpr int NQ = 4;                 // threads per row
constexpr int TCW = 128 * NQ;         // worker threads
constexpr int TCT = TCW;              // 16 warps -> 128 registers per thread (a 17th warp would cap them at 96 and spill)
constexpr int ISSUER = TCW / 32 - 1;  // the last worker warp also issues the MMAs (one elected lane, predicated)
constexpr int ROWS = 128;
constexpr int UPT = H / NQ;           // hidden units / output columns per thread = 16
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t C_U2 = 0, C_D1 = 64, C_G2 = 128, C_G1 = 192, C_AH = 256, C_AL = 320;
// Row-group strides of the transposed buffers.  A quarter warp (rows 4g..4g+7, 8 lanes x 16 B) stores to two
// consecutive row groups, so the stride has to be 64 (mod 128) bytes for the two halves to land in disjoint banks.
constexpr int RGH_MIN = (H + 1) * 16; // 64 units + the ones unit = 1040 B (unpadded: 2-way store conflicts)
constexpr int RG_PAD = (H + 4) * 16;  // 1088 B = 64 (mod 128)
constexpr int FS_LD = 128 + 4;        // leading dimension of the [feature][row] staging array (float4 reads, stride 4 banks)
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int NRG = ROWS / 4;         // 32 row groups (granule = 4 consecutive rows of one unit)
constexpr int WB = (H / 4) * H * 16;  // one W2 copy (hi or lo)                                                   = 16384 B

struct SmemMap {      // byte offsets, computed from F at run time (identical on host and device)
  int hT_hi, hT_lo, dT_hi, dT_lo, fT_hi, fT_lo, wa_hi, wa_lo, wb_hi, wb_lo, fstage, misc, total;
  int n_fu, rgf;      // feature units incl. ones, padded to 4 ; row-group stride of the feature buffer
  int n_g1;
  int rgh, rgd;       // row-group strides of the transposed h1 / delta buffers
};
__host__ __device__ inline SmemMap make_map(int F, bool pad_h = true) {
  SmemMap m;
  m.n_fu = 4 * ((F + 1 + 3) / 4);
  m.n_g1 = 8 * ((F + 1 + 7) / 8);   // N of the G1^T MMAs (a multiple of 8): columns >= n_fu read into the next row group, ignored
  m.rgf = m.n_fu * 16;
  m.rgd = RG_PAD;
  m.rgh = pad_h ? RG_PAD : RGH_MIN;
  int o = 0;
  m.fT_hi = o; o += NRG * m.rgf;
  m.fT_lo = o; o += NRG * m.rgf;
  m.hT_hi = o; o += NRG * m.rgh;
  m.hT_lo = o; o += NRG * m.rgh;
  m.dT_hi = o; o += NRG * m.rgd;
  m.dT_lo = o; o += NRG * m.rgd;
  m.wa_hi = o; o += WB;
  m.wa_lo = o; o += WB;
  m.wb_hi = o; o += WB;
  m.wb_lo = o; o += WB;
  m.fstage = o; o += m.n_fu * FS_LD * 4;
  m.misc = o; o += (F * H + 3 * H + H + 8) * 4 + 64;
  m.total = o;
  if (pad_h && m.total > SMEM_LIMIT) return make_map(F, false);   // many nuclei: give up the h1 padding to fit
  return m;
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// D=f32, A=B=tf32, both K-major, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

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

// 4x4 transpose across the 4 lanes of a quad (two butterfly stages), then hi/lo split and one 16-byte store per
// unit: on return the granule (rows row0..row0+3 of unit c0 + 4u + (lane & 3)) is in the transposed buffers.
__device__ __forceinline__ void stage_cols_smem(uint8_t* hi_rg, uint8_t* lo_rg, const int c0, const int t,
                                                const float (&v)[16]) {
  const bool odd = t & 1, upper = t & 2;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float x0 = v[4 * u], x1 = v[4 * u + 1], x2 = v[4 * u + 2], x3 = v[4 * u + 3];
    // stage 1 (partner t^1): 2x2 blocks
    const float s0 = odd ? x0 : x1, s1 = odd ? x2 : x3;
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    const float y0 = odd ? r0 : x0, y1 = odd ? x1 : r0, y2 = odd ? r1 : x2, y3 = odd ? x3 : r1;
    // stage 2 (partner t^2): swap the off-diagonal 2-element blocks
    const float q0 = upper ? y0 : y2, q1 = upper ? y1 : y3;
    const float p0 = __shfl_xor_sync(0xffffffffu, q0, 2), p1 = __shfl_xor_sync(0xffffffffu, q1, 2);
    const float z0 = upper ? p0 : y0, z1 = upper ? p1 : y1, z2 = upper ? y2 : p0, z3 = upper ? y3 : p1;
    float4 h4, l4;
    split_tf32(z0, h4.x, l4.x);
    split_tf32(z1, h4.y, l4.y);
    split_tf32(z2, h4.z, l4.z);
    split_tf32(z3, h4.w, l4.w);
    const int unit = c0 + 4 * u + t;
    *reinterpret_cast<float4*>(hi_rg + unit * 16) = h4;
    *reinterpret_cast<float4*>(lo_rg + unit * 16) = l4;
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
  const SmemMap mp = make_map(F);
  float* misc = reinterpret_cast<float*>(smem + mp.misc);
  float* W1s = misc;                 // [F][H]
  float* b1s = W1s + F * H;
  float* b2s = b1s + H;
  float* W3s = b2s + H;
  float* redS = W3s + H;             // [H + 8]: dW3[H], db3
  uint64_t* bars = reinterpret_cast<uint64_t*>(redS + H + 8);   // 4 mbarriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
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
  for (int idx = tid; idx < mp.n_fu * FS_LD; idx += TCT) fstage[idx] = 0.f;
  for (int idx = tid; idx < F * H; idx += TCT) W1s[idx] = theta[off.W1 + idx];
  for (int q = tid; q < H; q += TCT) {
    b1s[q] = theta[off.b1 + q];
    b2s[q] = theta[off.b2 + q];
    W3s[q] = theta[off.W3 + q];
  }
  for (int q = tid; q < H + 8; q += TCT) redS[q] = 0.f;
  {
    // all loads of a thread in flight at once: with a cold L2 each dependent round trip costs ~1 us
    constexpr int NL = H * H / TCT;
    static_assert(NL * TCT == H * H, "W2 staging assumes the block size divides H*H");
    float wv[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) wv[i] = theta[off.W2 + tid + i * TCT];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int idx = tid + i * TCT;
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
  // the ones unit (index 64) of the transposed h1 buffer -> row 64 of G2 = db2 ; hi = 1, lo stays 0
  for (int g = tid; g < NRG; g += TCT) *reinterpret_cast<float4*>(smem + mp.hT_hi + g * mp.rgh + H * 16) = make_float4(1.f, 1.f, 1.f, 1.f);
  if (tid == 0) {
    for (int b = 0; b < 4; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[b])));
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
  uint32_t ph_u = 0, ph_d = 0, ph_g2 = 0, ph_g = 0;
  nq_pdl_wait();                              // everything above read theta only; header, E_L, inverses come from predecessors
  float centerf = 0.f;
  if (hdr != nullptr) centerf = (float)(hdr[NQMC_HDR_SUM_E] / fmax(hdr[NQMC_HDR_N], 1.0));

  float gW3[UPT];
#pragma unroll
  for (int j = 0; j < UPT; ++j) gW3[j] = 0.f;
  float gb3 = 0.f;
  const int c0 = UPT * quarter;                     // first hidden unit / output column of this thread
  const int t4 = tid & 3;                           // position inside the lane quad == row % 4
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  uint8_t* hT_hi = smem + mp.hT_hi + (row >> 2) * mp.rgh;
  uint8_t* hT_lo = smem + mp.hT_lo + (row >> 2) * mp.rgh;
  uint8_t* dT_hi = smem + mp.dT_hi + (row >> 2) * mp.rgd;
  uint8_t* dT_lo = smem + mp.dT_lo + (row >> 2) * mp.rgd;

  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;
  // G1^T: D=f32, A=B=tf32, K-major, M=128 (64 delta1 units + 64 ignored rows), N = n_fu
  const uint32_t idesc_g1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(mp.n_g1 >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
  // positions and reverse seed of one tile row; the next tile's are fetched while the current one computes
  float nx = 0.f, ny = 0.f, nz = 0.f, nwt = 1.f, ninv = 0.f;     // raw loads; consumed one tile later
  auto fetch = [&](const int64_t item) {
    const int i = (int)(item % ns);
    int64_t w = (item / ns) * ROWS + row;
    if (w >= W) w = W - 1;
    const int e = elec0 + i;
    nx = r[(int64_t)(3 * e + 0) * W + w];
    ny = r[(int64_t)(3 * e + 1) * W + w];
    nz = r[(int64_t)(3 * e + 2) * W + w];
    if (wgt != nullptr) nwt = wgt[w];
    else if (e_l != nullptr) nwt = e_l[w];
    ninv = inv[(int64_t)ent_index(s, spin, i, korb) * W + w];
  };
  if (worker && chunk < n_items) fetch(chunk);
  // Pipeline (one tile = 3 CTA barriers; delta1 of tile t stays in registers until tile t+1's first phase, so the
  // wait for G2(t) -- which still reads the buffer delta1^T goes into -- hides behind layer 1 of tile t+1):
  //   P1  layer 1(t) -> A(h1) in TMEM ; delta1^T(t-1) -> dT                 | issue U2(t), G1(t-1)
  //   P2  h1^T(t) -> hT ; wait U2 ; h2, dW3, delta2 -> A(delta2) in TMEM      | issue D1(t)
  //   P3  wait G1(t-1) ; delta2^T(t) -> dT ; features(t) -> fT              | issue G2(t)
  //   P4  wait D1 ; delta1(t) = D1 (1 - h1^2)  (registers)
  float d1[UPT];
#pragma unroll
  for (int j = 0; j < UPT; ++j) d1[j] = 0.f;
  auto issue_g1 = [&](const uint32_t leader, const bool first) {   // G1^T += delta1 [f|1]  (M = delta1 unit, N = feature)
    const uint64_t ah = umma_desc(sbase + mp.dT_hi, mp.rgd, 128), al = umma_desc(sbase + mp.dT_lo, mp.rgd, 128);
    const uint64_t bh = umma_desc(sbase + mp.fT_hi, mp.rgf, 128), bl = umma_desc(sbase + mp.fT_lo, mp.rgf, 128);
#pragma unroll
    for (int ks = 0; ks < ROWS / 8; ++ks) {
      const uint64_t ao = (uint64_t)((ks * 2 * mp.rgd) >> 4), bo = (uint64_t)((ks * 2 * mp.rgf) >> 4);
      const uint32_t acc = (first && ks == 0) ? 0u : 1u;
      umma_ss(tmem_base + C_G1, ah + ao, bh + bo, idesc_g1, acc, leader);
      umma_ss(tmem_base + C_G1, al + ao, bh + bo, idesc_g1, 1u, leader);
      umma_ss(tmem_base + C_G1, ah + ao, bl + bo, idesc_g1, 1u, leader);
    }
    umma_commit(bar_g, leader);
  };
  int64_t n_done = 0;                                  // tiles this CTA has completed
  for (int64_t item = chunk; item < n_items; item += n_chunks, ++n_done) {
    float h1[UPT], d2[UPT], seed = 0.f;
    // ======== P1 ========
    if (worker) {
      const float x = nx, y = ny, z = nz;
      {
        float wt = nwt;
        if (wgt == nullptr && e_l != nullptr) wt = isfinite(nwt) ? nwt - centerf : 0.f;
        seed = wt * ninv;
        if ((item / ns) * ROWS + row >= W || !isfinite(seed)) seed = 0.f;
      }
      if (item + n_chunks < n_items) fetch(item + n_chunks);
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
        stage_cols_smem(dT_hi, dT_lo, c0, t4, d1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == ISSUER) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t leader = elect_one();
      const uint64_t bh = umma_desc(sbase + mp.wa_hi, H * 16, 128), bl = umma_desc(sbase + mp.wa_lo, H * 16, 128);
#pragma unroll
      for (int ks = 0; ks < H / 8; ++ks) {               // U2 = h1 W2 (A from tensor memory)
        const uint64_t bo = (uint64_t)((ks * 2 * H * 16) >> 4);
        const uint32_t ah = tmem_base + C_AH + 8 * ks, al = tmem_base + C_AL + 8 * ks;
        umma_ts(tmem_base + C_U2, ah, bh + bo, ks == 0 ? 0u : 1u, leader);
        umma_ts(tmem_base + C_U2, al, bh + bo, 1u, leader);
        umma_ts(tmem_base + C_U2, ah, bl + bo, 1u, leader);
      }
      umma_commit(bar_u, leader);
      if (n_done > 0) issue_g1(leader, n_done == 1);
      __syncwarp();
    }
    // ======== P2 ========
    if (worker) {
      stage_cols_smem(hT_hi, hT_lo, c0, t4, h1);         // G2(t-1), the reader of this buffer, was waited for in P1
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
      if (quarter == 0) gb3 += seed;
      stage_rows_tmem(tmem_base + lane_base + C_AH + (uint32_t)c0, tmem_base + lane_base + C_AL + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < UPT; ++j) d2[j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == ISSUER) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t leader = elect_one();
      const uint64_t bh = umma_desc(sbase + mp.wb_hi, H * 16, 128), bl = umma_desc(sbase + mp.wb_lo, H * 16, 128);
#pragma unroll
      for (int ks = 0; ks < H / 8; ++ks) {               // D1 = delta2 W2^T (A from tensor memory)
        const uint64_t bo = (uint64_t)((ks * 2 * H * 16) >> 4);
        const uint32_t ah = tmem_base + C_AH + 8 * ks, al = tmem_base + C_AL + 8 * ks;
        umma_ts(tmem_base + C_D1, ah, bh + bo, ks == 0 ? 0u : 1u, leader);
        umma_ts(tmem_base + C_D1, al, bh + bo, 1u, leader);
        umma_ts(tmem_base + C_D1, ah, bl + bo, 1u, leader);
      }
      umma_commit(bar_d, leader);
      __syncwarp();
    }
    // ======== P3 ========
    if (worker) {
      if (n_done > 0) { mbar_wait(bar_g, ph_g); ph_g ^= 1; }   // G1(t-1) reads delta1^T(t-1) and the features of tile t-1
      stage_cols_smem(dT_hi, dT_lo, c0, t4, d2);
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
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == ISSUER) {
      const uint32_t leader = elect_one();
      const uint64_t ah = umma_desc(sbase + mp.hT_hi, mp.rgh, 128), al = umma_desc(sbase + mp.hT_lo, mp.rgh, 128);
      const uint64_t bh = umma_desc(sbase + mp.dT_hi, mp.rgd, 128), bl = umma_desc(sbase + mp.dT_lo, mp.rgd, 128);
#pragma unroll
      for (int ks = 0; ks < ROWS / 8; ++ks) {            // G2 += [h1|1]^T delta2
        const uint64_t ao = (uint64_t)((ks * 2 * mp.rgh) >> 4), bo = (uint64_t)((ks * 2 * mp.rgd) >> 4);
        const uint32_t acc = (n_done == 0 && ks == 0) ? 0u : 1u;
        umma_ss(tmem_base + C_G2, ah + ao, bh + bo, IDESC, acc, leader);
        umma_ss(tmem_base + C_G2, al + ao, bh + bo, IDESC, 1u, leader);
        umma_ss(tmem_base + C_G2, ah + ao, bl + bo, IDESC, 1u, leader);
      }
      umma_commit(bar_g2, leader);
      __syncwarp();
    }
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
      stage_cols_smem(dT_hi, dT_lo, c0, t4, d1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == ISSUER) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t leader = elect_one();
      issue_g1(leader, n_done == 1);
      __syncwarp();
    }
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
      tmem_ld16(tmem_base + lane_base + C_G2 + (uint32_t)c0, v2);   // lane = h1 unit (row 64: ones -> db2)
    } else {
#pragma unroll
      for (int j = 0; j < UPT; ++j) v2[j] = 0.f;          // this CTA had no tile: the slab must still be defined
    }
    if (row < H) {
#pragma unroll
      for (int j = 0; j < UPT; ++j) slab[o_W2 + row * H + c0 + j] = v2[j];
    } else if (row == H) {
#pragma unroll
      for (int j = 0; j < UPT; ++j) slab[o_b2 + c0 + j] = v2[j];
    }
    // G1^T: lane = delta1 unit, column = feature (column F: ones -> db1); 8 columns per read, groups dealt over quarters
    for (int cg = quarter; cg < mp.n_g1 / 8; cg += NQ) {
      float v1[8];
      if (any) tmem_ld8(tmem_base + lane_base + C_G1 + (uint32_t)(8 * cg), v1);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v1[j] = 0.f;
      }
      if (row < H) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int f = 8 * cg + j;
          if (f < F) slab[o_W1 + f * H + row] = v1[j];
          else if (f == F) slab[o_b1 + row] = v1[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < UPT; ++j) {                     // dW3 / db3: reduce the per-row partials over the 128 rows
      const float t = warp_sum(gW3[j]);
      if ((tid & 31) == 0) atomicAdd(&redS[c0 + j], t);
    }
    if (quarter == 0) {
      const float t = warp_sum(gb3);
      if ((tid & 31) == 0) atomicAdd(&redS[H], t);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  for (int q = tid; q < H; q += TCT) slab[o_W3 + q] = redS[q];
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
      NQMC_CUDA(nq_launch_pdl(orb_bwd_tc_kernel<NA>, dim3(n_chunks * s.n_mlp), dim3(TCT), (size_t)mp.total, st, s, ctx->theta, r,
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
  const SmemMap mp = make_map(s.F);
  if (mp.total > SMEM_LIMIT || s.F + 1 > 64) return NQMC_ERR_UNSUPPORTED;   // too many nuclei for this variant's smem plan
  if (const char* e = getenv("NQMC_TC_BWD"))
    if (atoi(e) == 0) return NQMC_ERR_UNSUPPORTED;
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
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
