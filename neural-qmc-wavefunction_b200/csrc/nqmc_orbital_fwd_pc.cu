// nqmc_orbital_fwd_pc.cu -- orbital forward pass on tcgen05 for hidden width 128 (BASELINE configs 3 and 5): value
// only (CH = 1, the Metropolis proposal) or value + gradient + Laplacian (CH = 5, the local energy).
// Reference lines replaced: sampler/metropolis.py:124 (value), hamiltonian/molecular.py:60-64 (dense nested-autodiff
// Hessian -> forward Laplacian); orbital definition [SPEC] .github/phase2-issue-body.md:78-113.
//
// Same producer / consumer pipeline as nqmc_orbital_lap_pc.cu (hidden width 64) -- one orbital network per CTA,
// 128-row tiles (128 consecutive walkers x one electron), W2 split once per CTA into tf32 hi/lo in shared memory,
// 16 worker warps (thread = (row, quarter)) computing layer 1 + activation rules for 8 hidden units at a time and
// storing them as tf32 hi/lo straight into TMEM (the A operand), a 17th warp issuing
// tcgen05.mma.cta_group::1.kind::tf32 (3 products per fp32-grade product) behind mbarriers -- generalised over the width:
//   * CH = 5: the accumulators of a 128-row tile would need 5 x 128 = 640 TMEM columns (> 512), so the output units
//     are processed in NPASS = H / 64 passes of N = 64.  The layer-1 producer work is repeated per pass (layer 1 is
//     3..7 % of the MACs) and the output-layer partial sums of a row are carried in registers across the passes;
//   * CH = 1: one pass with N = H;
//   * more than 6 nuclei: the per-row nucleus features (d, exp(-d), unit vector, 2/d) no longer fit in registers next
//     to the accumulators.  The four quarter-threads of a row compute them once per tile, each for a quarter of the
//     nuclei, into a shared-memory table that every chunk reads back with 16-byte loads.
// TMEM: CH x NT accumulator columns + 2 x (CH channels x 8 units x {hi, lo}) A columns (CH = 5: 480; CH = 1: 160).
#include "nqmc_internal.cuh"

namespace {

constexpr int NQ = 4;
constexpr int TCW = 128 * NQ;
constexpr int NWW = TCW / 32;          // worker warps
constexpr int TCT = TCW + 128;         // 16 worker warps + the issuer warpgroup (one issuing warp)
constexpr int ROWS = 128;
constexpr int KC = 8;                  // units per K-chunk == one UMMA k-step
constexpr int UPT = KC / NQ;           // 2 units per thread per chunk
constexpr int NISS = 4;                // issuing warps (one thread sustains one tcgen05.mma per ~117 cycles; see nqmc_orbital_lap_pc.cu)

template <int H, int NA, int CH, bool GW = false, int NBUF = 2, bool BW = false>
struct Cfg {
  static constexpr int NT = CH == 1 ? H : 64;          // output units per pass == UMMA N
  static constexpr int NPASS = H / NT;
  static constexpr int NCHUNK = H / KC;
  static constexpr int EPC = NT / NQ;                  // epilogue columns per thread
  static constexpr uint32_t C_A = CH * NT;             // first A-staging column
  static constexpr uint32_t A_LO = CH * KC, A_STRIDE = 2 * CH * KC;
  // CH = 1 (value pass / reverse-pass value sweep): layer 1 on the tensor cores too.  The row's four threads write the
  // features [x, y, z, 1, (d_a, exp(-d_a)) ...] as tf32 hi / lo straight into TMEM (nuclei dealt over the quarters), K1 / 8
  // x 3 UMMAs with N = H form all pre-activations [f|1] [W1; b1] in 128 spare columns, and the chunk producer reads its
  // two units with one tcgen05.ld instead of 2 (3 + 2 NA) FMAs behind as many weight loads (-1.0 of 3.0 ms at 10 nuclei).
  static constexpr bool L1TC = CH == 1 && 4 + 2 * NA <= 24;
  static constexpr int K1 = ((4 + 2 * NA + 7) / 8) * 8;
  static constexpr uint32_t C_F1 = C_A + NBUF * A_STRIDE;        // features: K1 hi columns, then K1 lo columns
  static constexpr uint32_t C_Z1 = 256;                          // pre-activations: H columns
  static constexpr int W1B = (K1 / 4) * H * 16;                  // one [W1; b1] operand copy (hi or lo)
  static constexpr uint32_t TMEM_COLS = L1TC ? 512 : ((C_A + NBUF * A_STRIDE) <= 256 ? 256 : 512);
  static_assert(!L1TC || (C_F1 + 2 * K1 <= C_Z1 && C_Z1 + H <= 512), "TMEM: layer-1 operands do not fit");
  static_assert(C_A + NBUF * A_STRIDE <= 512, "TMEM: accumulators + A staging exceed 512 columns");
  // value pass: only d and exp(-d) per nucleus (2 NA registers); GW: one more value per nucleus + the six G entries
  static constexpr bool FEAT_SMEM = CH == 1 ? (NA > 16) : (NA > 6 || (GW && NA > 4));
  static constexpr int B_K4 = H * 16;                  // bytes between K-groups of 4 (LBO)
  static constexpr int B_MAT = (H / 4) * B_K4;
  static constexpr uint32_t IDESC =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
  static size_t smem_bytes(int F) {
    return 2 * (size_t)B_MAT + (L1TC ? 2 * (size_t)W1B : 0) + ((size_t)F * H + 3 * H + (NQ - 1) * ROWS * CH) * 4 +
           (FEAT_SMEM ? (size_t)NA * 2 * ROWS * 16 : 0) + 160;
  }
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                        uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, l;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 l, %5, 0;\n\t"
      "@l tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void split_tf32(float a, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(a) & 0xFFFFE000u;
  lo = __float_as_uint(a - __uint_as_float(hi));
}

// BW (reverse-pass mode, CH = 1, H = 128; see nqmc_orbital_bwd128.cu): the value pass of the reverse sweep.  Instead of
// phi it leaves, for every row (row id = 128 item + row, item = walker block * ns + electron), unit-major and
// row-contiguous (so the stores are coalesced and the arrays are K-major GEMM operands for the weight-gradient kernel):
//   hT[m][unit][row] = h1, dT[m][unit][row] = delta2 = seed W3 (1 - h2^2), fT[m][0..F][row] = [1 | features],
// and accumulates dW3 += seed h2, db3 += seed in registers (flushed into the CTA's slab at the end).
// The 5-channel pass (CH = 5, BW false) called with hT != nullptr parks the same per-row activations for a reverse pass
// that follows at the same positions (fused walker step): hT = h1, dT = h2 (NOT delta2: the seeds need E_L and the inverse
// matrices, which exist only after this pass -- bwd128_seed_kernel turns h2 into delta2), fT = [1 | features].  The value
// sweep above is then skipped (H10: 2.67 ms) for a 1.46 ms HBM-bound seed kernel and +0.5 ms here.  The stores sit where they
// cost least (measured one by one): h1 in the LAST pass's chunk loop, which only reloads layer 1 and waits for the MMAs
// anyway (+0.12 ms at H10; in pass 0, +0.29), the features after the first pass's chunks while the tensor pipe drains
// (+0.08; ahead of the first chunk, +0.24), h2 in the epilogue that computes it (+0.28).
struct BwdOut {
  const float* inv; const float* wgt; const double* hdr; const float* e_l;   // seed = weight * inverse-matrix entry
  float* hT; float* dT; float* fT; float* part;
  int64_t rows_pad; int p_mlp;
};

// GW: the fifth channel is the G-weighted Laplacian L_G = sum_ab G_ab d_a d_b (backflow, nqmc_backflow.cu) with
// G = Gw[(6 e + c) W + w] (packed xx xy xz yy yz zz) per (electron, walker); GW == false is G = identity.
template <int H, int NA, int CH, bool GW, int NBUF, bool BW>
__global__ void __launch_bounds__(TCT, 1)
    orb_fwd_pc_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r, const int64_t W,
                      float* __restrict__ phi, const int n_chunks, const float* __restrict__ Gw, const BwdOut bw,
                      float* __restrict__ l1cache) {
  using C = Cfg<H, NA, CH, GW, NBUF, BW>;
  static_assert(!BW || (CH == 1 && !GW && !C::FEAT_SMEM && 3 + 2 * NA + 1 <= 32), "reverse-pass mode: value pass, <= 14 nuclei");
  constexpr int NT = C::NT, NPASS = C::NPASS, NCHUNK = C::NCHUNK, EPC = C::EPC;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* Bhi = smem_raw;                                    // [H/4][H n][4]
  uint8_t* Blo = Bhi + C::B_MAT;
  uint8_t* W1hi = Blo + C::B_MAT;                             // L1TC: [K1/4][H n][4] = [W1; b1]^T operand, hi / lo
  uint8_t* W1lo = W1hi + (C::L1TC ? C::W1B : 0);
  float* W1s = reinterpret_cast<float*>(W1lo + (C::L1TC ? C::W1B : 0));      // [F][H]
  float* b1s = W1s + s.F * H;
  float* b2s = b1s + H;
  float* W3s = b2s + H;
  float* outS = W3s + H;                                      // [NQ-1][ROWS][CH]
  float4* featA = reinterpret_cast<float4*>(outS + (NQ - 1) * ROWS * CH);   // [NA][ROWS] {d, e, 2/d, -}   (FEAT_SMEM)
  float4* featB = featA + (C::FEAT_SMEM ? NA * ROWS : 0);                   // [NA][ROWS] {ux, uy, uz, -}
  uint64_t* bars = reinterpret_cast<uint64_t*>(featB + (C::FEAT_SMEM ? NA * ROWS : 0));   // full[NBUF], empty[NBUF], dfull, dempty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NBUF + 4);   // + f_full, z_full (L1TC)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (ROWS - 1), quarter = (tid >> 7) & (NQ - 1);
  const bool worker = tid < TCW;
  const int m = blockIdx.x % s.n_mlp, chunk = blockIdx.x / s.n_mlp;
  const int spin = m < s.n_up ? 0 : 1;
  const int korb = spin == 0 ? m : m - s.n_up;
  const int ns = spin == 0 ? s.n_up : s.n_dn;
  const int elec0 = spin == 0 ? 0 : s.n_up;
  const MlpOff off = s.mlp[m];
  const int F = s.F;

  for (int idx = tid; idx < F * H; idx += TCT) W1s[idx] = theta[off.W1 + idx];
  for (int q = tid; q < H; q += TCT) {
    b1s[q] = theta[off.b1 + q];
    b2s[q] = theta[off.b2 + q];
    W3s[q] = theta[off.W3 + q];
  }
  for (int idx = tid; idx < H * H; idx += TCT) {
    const int k = idx / H, n = idx % H;
    uint32_t hi, lo;
    split_tf32(theta[off.W2 + idx], hi, lo);
    const int o = (k >> 2) * (H * 4) + n * 4 + (k & 3);
    reinterpret_cast<uint32_t*>(Bhi)[o] = hi;
    reinterpret_cast<uint32_t*>(Blo)[o] = lo;
  }
  if constexpr (C::L1TC) {
    for (int idx = tid; idx < C::K1 * H; idx += TCT) {
      const int k = idx / H, n = idx % H;   // operand feature order: x, y, z, 1, then (d_a, exp(-d_a)) pairs, zero padding
      float v = 0.f;
      if (k < 3) v = theta[off.W1 + k * H + n];
      else if (k == 3) v = theta[off.b1 + n];
      else if (k < 4 + 2 * NA) v = theta[off.W1 + (((k - 4) & 1) ? 3 + NA + ((k - 4) >> 1) : 3 + ((k - 4) >> 1)) * H + n];
      uint32_t hi, lo;
      split_tf32(v, hi, lo);
      const int o = (k >> 2) * (H * 4) + n * 4 + (k & 3);
      reinterpret_cast<uint32_t*>(W1hi)[o] = hi;
      reinterpret_cast<uint32_t*>(W1lo)[o] = lo;
    }
  }
  const float b3 = theta[off.b3];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[2 * NBUF + 2])), "r"(NWW));   // f_full
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[2 * NBUF + 3])));              // z_full
    for (int b = 0; b < NBUF; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[b])), "r"(NWW));   // full[b]
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[NBUF + b])), "r"(NISS));   // empty[b]
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[2 * NBUF])), "r"(NISS));  // dfull
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[2 * NBUF + 1])), "r"(NWW));   // dempty
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[NBUF]);
  const uint32_t dfull = smem_u32(&bars[2 * NBUF]), dempty = smem_u32(&bars[2 * NBUF + 1]);
  const uint32_t f_full = smem_u32(&bars[2 * NBUF + 2]), z_full = smem_u32(&bars[2 * NBUF + 3]);

  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;
  nq_pdl_wait();                              // everything above read theta only; the positions come from the predecessor

  if (!worker) {
    // =============================== MMA issuer warp =====================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    const int iw = warp - NWW;                                 // CH = 5: channel c is issued by warp c % NISS; CH = 1: k-steps dealt round-robin
    const uint32_t leader = elect_one();
    const uint64_t dBh = umma_desc(smem_u32(Bhi), C::B_K4, 128), dBl = umma_desc(smem_u32(Blo), C::B_K4, 128);
    uint32_t g = 0, ep = 0;                                    // global chunk counter, global pass counter
    uint32_t tl = 0;                                           // local tile counter
    for (int64_t item = chunk; item < n_items; item += n_chunks, ++tl) {
      if constexpr (C::L1TC) {
        if (iw == 0) {                                         // Z1 = [f|1] [W1; b1]: every worker has read Z1 of the previous tile
          mbar_wait(f_full, tl & 1);                           // (they arrive here only after their chunk loop)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t d1h = umma_desc(smem_u32(W1hi), C::B_K4, 128), d1l = umma_desc(smem_u32(W1lo), C::B_K4, 128);
#pragma unroll
          for (int ks = 0; ks < C::K1 / 8; ++ks) {
            const uint64_t kb = (uint64_t)(((uint32_t)ks * 2 * C::B_K4) >> 4);
            const uint32_t ah = tmem_base + C::C_F1 + 8u * ks, al = ah + (uint32_t)C::K1;
            umma_ts(tmem_base + C::C_Z1, ah, d1h + kb, C::IDESC, ks == 0 ? 0u : 1u, leader);
            umma_ts(tmem_base + C::C_Z1, al, d1h + kb, C::IDESC, 1u, leader);
            umma_ts(tmem_base + C::C_Z1, ah, d1l + kb, C::IDESC, 1u, leader);
          }
          umma_commit(z_full, leader);
        }
      }
#pragma unroll 1
      for (int pass = 0; pass < NPASS; ++pass, ++ep) {
        if (ep > 0) mbar_wait(dempty, (ep - 1) & 1);           // accumulators of the previous pass have been read
        const uint64_t nb = (uint64_t)(((uint32_t)pass * NT * 16) >> 4);   // output units [NT pass, NT pass + NT)
#pragma unroll 4
        for (int kc = 0; kc < NCHUNK; ++kc, ++g) {
          const uint32_t b = g % NBUF, u = g / NBUF;
          mbar_wait(full0 + 8 * b, u & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t kb = (uint64_t)(((uint32_t)kc * 2 * C::B_K4) >> 4) + nb;
          const uint32_t abase = tmem_base + C::C_A + C::A_STRIDE * b;
          for (int c = iw; c < CH; c += NISS) {                // a warp without a channel only commits
            const uint32_t d = tmem_base + (uint32_t)(c * NT);
            const uint32_t ah = abase + 8u * c, al = ah + C::A_LO;
            umma_ts(d, ah, dBh + kb, C::IDESC, kc == 0 ? 0u : 1u, leader);
            umma_ts(d, al, dBh + kb, C::IDESC, 1u, leader);
            umma_ts(d, ah, dBl + kb, C::IDESC, 1u, leader);
          }
          umma_commit(empty0 + 8 * b, leader);
          if (kc == NCHUNK - 1) umma_commit(dfull, leader);
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== worker warps ========================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t g = 0, ep = 0, tl = 0;
    if constexpr (C::L1TC) {                                   // the padding feature columns must read as zero (written once)
      if (quarter == NQ - 1) {
#pragma unroll
        for (int k = 4 + 2 * NA; k < C::K1; ++k) {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem_base + lane_base + C::C_F1 + (uint32_t)k), "r"(0u) : "memory");
          asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem_base + lane_base + C::C_F1 + (uint32_t)(C::K1 + k)), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
    }
    float nx = 0.f, ny = 0.f, nz = 0.f;                        // positions of the next tile, fetched one tile ahead
    float nwt = 1.f, ninv = 0.f;                               // BW: weight and inverse-matrix entry of the row
    float centerf = 0.f;
    if constexpr (BW) {
      if (bw.hdr != nullptr) centerf = (float)(bw.hdr[NQMC_HDR_SUM_E] / fmax(bw.hdr[NQMC_HDR_N], 1.0));
    }
    float gW3[BW ? C::EPC : 1];
    float gb3 = 0.f;
#pragma unroll
    for (int j = 0; j < (BW ? C::EPC : 1); ++j) gW3[j] = 0.f;
    ItemIter cur(chunk, n_chunks, ns), nxt = cur;             // this tile / the next tile to request (in order)
    auto fetch = [&]() {
      int64_t w = (int64_t)nxt.wb * ROWS + row;
      if (w >= W) w = W - 1;
      const int e = elec0 + nxt.e;
      nx = r[(int64_t)(3 * e + 0) * W + w];
      ny = r[(int64_t)(3 * e + 1) * W + w];
      nz = r[(int64_t)(3 * e + 2) * W + w];
      if constexpr (BW) {
        nwt = bw.wgt != nullptr ? bw.wgt[w] : (bw.e_l != nullptr ? bw.e_l[w] : 1.0f);
        ninv = bw.inv[(int64_t)ent_index(s, spin, nxt.e, korb) * W + w];
      }
      nxt.next();
    };
    if (chunk < n_items) fetch();
    for (int64_t item = chunk; item < n_items; item += n_chunks, ++tl, cur.next()) {
      const int i = cur.e;
      const int64_t w0 = (int64_t)cur.wb * ROWS;
      const float x = nx, y = ny, z = nz;
      float seed = 0.f;
      if constexpr (BW) {
        float wt = nwt;
        if (bw.wgt == nullptr && bw.e_l != nullptr) wt = isfinite(nwt) ? nwt - centerf : 0.f;
        seed = wt * ninv;
        if (w0 + row >= W || !isfinite(seed)) seed = 0.f;
      }
      const int64_t rid = item * ROWS + row;                   // BW: row id inside this network's scratch arrays
      if (item + n_chunks < n_items) fetch();
      // per nucleus: d, exp(-d), unit vector, 2/d.  With s = wd - e we the five channels of (d, e) contribute
      //   value: d wd + e we ; gradient: u s ; Laplacian: (2/d) s + e we      (7 FMA per nucleus and unit, not 10)
      // GW: Laplacian-channel inputs become  L_G d = (tr G - u^T G u)/d =: A ,  L_G e^-d = e (q - A) with q = u^T G u
      // (G = I: A = 2/d, q = 1), so the channel gets A s + q e we per nucleus.
      constexpr int NREG = C::FEAT_SMEM ? 1 : NA;
      float fdd[NREG], fex[NREG], fux[NREG], fuy[NREG], fuz[NREG], fti[NREG], fq[GW ? NREG : 1];
      float g6[6] = {1.f, 0.f, 0.f, 1.f, 0.f, 1.f};
      if constexpr (GW) {
        int64_t wg = w0 + row;
        if (wg >= W) wg = W - 1;
#pragma unroll
        for (int c = 0; c < 6; ++c) g6[c] = Gw[(int64_t)(6 * (elec0 + i) + c) * W + wg];
      }
      auto quad = [&](const float a, const float b, const float c) {   // (a,b,c)^T G (a,b,c)
        return fmaf(g6[0] * a, a, fmaf(g6[3] * b, b, g6[5] * c * c)) +
               2.0f * fmaf(g6[1] * a, b, fmaf(g6[2] * a, c, g6[4] * b * c));
      };
      const float trG = g6[0] + g6[3] + g6[5];
      if constexpr (C::L1TC) {
        // this thread's share of the row's features -> TMEM (and, in reverse-pass mode, -> fT); then Z1 from the issuer
        const uint32_t fa = tmem_base + lane_base + C::C_F1;
        float* fr = nullptr;
        if constexpr (BW) fr = bw.fT + (int64_t)m * 32 * bw.rows_pad + rid;
        if (quarter == 0) {
          uint32_t h0, l0, h1, l1, h2, l2;
          split_tf32(x, h0, l0);
          split_tf32(y, h1, l1);
          split_tf32(z, h2, l2);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(fa), "r"(h0), "r"(h1), "r"(h2), "r"(0x3f800000u) : "memory");
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(fa + (uint32_t)C::K1), "r"(l0), "r"(l1), "r"(l2), "r"(0u) : "memory");
          if constexpr (BW) {
            fr[0] = 1.0f;
            fr[bw.rows_pad] = x;
            fr[2 * bw.rows_pad] = y;
            fr[3 * bw.rows_pad] = z;
          }
        }
#pragma unroll
        for (int a0 = 0; a0 < NA; a0 += NQ) {
          const int a = a0 + quarter;
          if (a < NA) {
            const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
            const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
            const float ex = __expf(-dd);
            uint32_t hd, ld, he, le;
            split_tf32(dd, hd, ld);
            split_tf32(ex, he, le);
            tmem_st2(fa + (uint32_t)(4 + 2 * a), hd, he);
            tmem_st2(fa + (uint32_t)(C::K1 + 4 + 2 * a), ld, le);
            if constexpr (BW) {
              fr[(int64_t)(4 + a) * bw.rows_pad] = dd;
              fr[(int64_t)(4 + NA + a) * bw.rows_pad] = ex;
            }
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(f_full);
        mbar_wait(z_full, tl & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      } else if constexpr (C::FEAT_SMEM) {
        // every reader of the previous tile's table has passed that tile's quarter-combine barrier
        for (int a = quarter; a < NA; a += NQ) {
          const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
          const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
          const float inv = 1.0f / dd;
          const float ux = dx * inv, uy = dy * inv, uz = dz * inv;
          const float q = GW ? quad(ux, uy, uz) : 1.0f;
          featA[a * ROWS + row] = make_float4(dd, __expf(-dd), GW ? (trG - q) * inv : 2.0f * inv, q);
          featB[a * ROWS + row] = make_float4(ux, uy, uz, 0.f);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TCW) : "memory");
      } else {
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
          const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
          const float inv = 1.0f / dd;
          fdd[a] = dd; fex[a] = __expf(-dd); fux[a] = dx * inv; fuy[a] = dy * inv; fuz[a] = dz * inv; fti[a] = 2.0f * inv;
          if constexpr (GW) {
            fq[a] = quad(fux[a], fuy[a], fuz[a]);
            fti[a] = (trG - fq[a]) * inv;
          }
        }
      }
      const bool st_act = !BW && CH == 5 && bw.hT != nullptr;     // park h1, h2 and the features for the reverse pass
      if constexpr (BW && !C::L1TC) {                            // [1 | x y z | d_a | e_a] of this row, dealt over its four threads
        float* fr = bw.fT + (int64_t)m * 32 * bw.rows_pad + rid;
        const float fv[4] = {1.0f, x, y, z};
#pragma unroll
        for (int f = 0; f < 4; ++f)
          if ((f & 3) == quarter) fr[(int64_t)f * bw.rows_pad] = fv[f];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          if (((4 + a) & 3) == quarter) fr[(int64_t)(4 + a) * bw.rows_pad] = fdd[a];
          if (((4 + NA + a) & 3) == quarter) fr[(int64_t)(4 + NA + a) * bw.rows_pad] = fex[a];
        }
      }
      float out[CH];
      float2 out2[CH];                                         // even / odd columns of this thread, added at the end
#pragma unroll
      for (int c = 0; c < CH; ++c) out2[c] = make_float2(0.f, 0.f);
      float pre[(NPASS > 1) ? (CH == 1 ? 1 : 5) * UPT : 1];        // pass 1: the next chunk's cached layer-1 outputs
#pragma unroll 1
      for (int pass = 0; pass < NPASS; ++pass, ++ep) {
        if (NPASS > 1 && pass > 0) {
#pragma unroll
          for (int q = 0; q < (CH == 1 ? 1 : 5) * UPT; ++q)
            pre[q] = __ldcg(l1cache + (((int64_t)blockIdx.x * NCHUNK) * ((CH == 1 ? 1 : 5) * UPT) + q) * TCW + tid);
        }
#pragma unroll 2
        for (int kc = 0; kc < NCHUNK; ++kc, ++g) {
          const int j0 = KC * kc + UPT * quarter;
          constexpr int CM = CH == 1 ? 1 : 5;                  // channels computed (CH = 4 exists for timing experiments only)
          float acc[CM][UPT];
          // Two output passes (CH = 5, H = 128) need the same layer-1 / tanh-rule outputs twice.  They are computed in pass 0
          // and parked in a per-CTA scratch that stays L2-resident (10 floats per thread and chunk, thread-major so each
          // warp writes and reads 128 contiguous bytes; .cg on both sides: the lines are rewritten every tile); pass 1 loads
          // them one chunk ahead instead of recomputing (-31 % of this kernel at 10 nuclei).
          float* cslot = nullptr;
          if constexpr (NPASS > 1) cslot = l1cache + (((int64_t)blockIdx.x * NCHUNK + kc) * (CM * UPT)) * TCW + tid;
          if constexpr (C::L1TC) {
            uint32_t z0, z1;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(z0), "=r"(z1) : "r"(tmem_base + lane_base + C::C_Z1 + (uint32_t)j0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float2 h = nq_tanh2(make_float2(__uint_as_float(z0), __uint_as_float(z1)));
            acc[0][0] = h.x; acc[0][1] = h.y;
          } else if (NPASS > 1 && pass > 0) {
#pragma unroll
            for (int c = 0; c < CM; ++c)
#pragma unroll
              for (int u2 = 0; u2 < UPT; ++u2) acc[c][u2] = pre[c * UPT + u2];
            if (kc + 1 < NCHUNK) {
#pragma unroll
              for (int q = 0; q < CM * UPT; ++q) pre[q] = __ldcg(cslot + (int64_t)(CM * UPT) * TCW + (int64_t)q * TCW);
            }
          } else {
            // the thread's two hidden units are one packed pair: every FMA / MUL / ADD below is one FFMA2 / FMUL2 / FADD2
            float2 pa[CM];
            pa[0] = *reinterpret_cast<const float2*>(b1s + j0);
            float2 ew = make_float2(0.f, 0.f);                     // sum_a e_a we_a : part of the value and of the Laplacian
            float2 ewq = make_float2(0.f, 0.f);                    // GW: sum_a q_a e_a we_a (Laplacian channel)
            {
              const float2 w0 = *reinterpret_cast<const float2*>(W1s + 0 * H + j0);
              const float2 w1 = *reinterpret_cast<const float2*>(W1s + 1 * H + j0);
              const float2 w2 = *reinterpret_cast<const float2*>(W1s + 2 * H + j0);
              pa[0] = nq_fma2(nq_bc2(x), w0, pa[0]);
              pa[0] = nq_fma2(nq_bc2(y), w1, pa[0]);
              pa[0] = nq_fma2(nq_bc2(z), w2, pa[0]);
              if constexpr (CM == 5) {
                pa[1] = w0; pa[2] = w1; pa[3] = w2;
                pa[4] = make_float2(0.f, 0.f);
              }
            }
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              const float2 wd = *reinterpret_cast<const float2*>(W1s + (3 + a) * H + j0);
              const float2 we = *reinterpret_cast<const float2*>(W1s + (3 + NA + a) * H + j0);
              float dd, ex, ux = 0.f, uy = 0.f, uz = 0.f, ti = 0.f, qa = 1.f;
              if constexpr (C::FEAT_SMEM) {
                const float4 fa = featA[a * ROWS + row];
                dd = fa.x; ex = fa.y; ti = fa.z; qa = fa.w;
                if constexpr (CM == 5) {
                  const float4 fb = featB[a * ROWS + row];
                  ux = fb.x; uy = fb.y; uz = fb.z;
                }
              } else {
                dd = fdd[a]; ex = fex[a]; ux = fux[a]; uy = fuy[a]; uz = fuz[a]; ti = fti[a];
                if constexpr (GW) qa = fq[a];
              }
              const float2 ewe = nq_mul2(nq_bc2(ex), we);
              ew = nq_add2(ew, ewe);
              if constexpr (GW) ewq = nq_fma2(nq_bc2(qa), ewe, ewq);
              pa[0] = nq_fma2(nq_bc2(dd), wd, pa[0]);
              if constexpr (CM == 5) {
                const float2 sv = nq_fma2(nq_bc2(-ex), we, wd);
                pa[1] = nq_fma2(nq_bc2(ux), sv, pa[1]);
                pa[2] = nq_fma2(nq_bc2(uy), sv, pa[2]);
                pa[3] = nq_fma2(nq_bc2(uz), sv, pa[3]);
                pa[4] = nq_fma2(nq_bc2(ti), sv, pa[4]);
              }
            }
            // tanh rules: lap' = tp (lap - 2 h |grad|^2)
            const float2 h = nq_tanh2(nq_add2(pa[0], ew));
            if constexpr (CM == 5) {
              const float2 tp = nq_fma2(make_float2(-h.x, -h.y), h, nq_bc2(1.0f));
              float2 g2;
              if constexpr (GW) g2 = make_float2(quad(pa[1].x, pa[2].x, pa[3].x), quad(pa[1].y, pa[2].y, pa[3].y));
              else g2 = nq_fma2(pa[1], pa[1], nq_fma2(pa[2], pa[2], nq_mul2(pa[3], pa[3])));
              pa[4] = nq_mul2(tp, nq_fma2(make_float2(-(h.x + h.x), -(h.y + h.y)), g2, nq_add2(pa[4], GW ? ewq : ew)));
              pa[1] = nq_mul2(pa[1], tp);
              pa[2] = nq_mul2(pa[2], tp);
              pa[3] = nq_mul2(pa[3], tp);
            }
            pa[0] = h;
#pragma unroll
            for (int c = 0; c < CM; ++c) { acc[c][0] = pa[c].x; acc[c][1] = pa[c].y; }
            if constexpr (NPASS > 1) {
#pragma unroll
              for (int c = 0; c < CM; ++c)
#pragma unroll
                for (int u2 = 0; u2 < UPT; ++u2) __stcg(cslot + (int64_t)(c * UPT + u2) * TCW, acc[c][u2]);
            }
          }
          if constexpr (BW) {
            float* hr = bw.hT + ((int64_t)m * H + j0) * bw.rows_pad + rid;
            hr[0] = acc[0][0];
            hr[bw.rows_pad] = acc[0][1];
          } else if constexpr (CH == 5) {
            if (st_act && pass == NPASS - 1) {                        // last pass: its chunk loop only reloads layer 1, the stores hide behind the MMAs
              float* hr = bw.hT + ((int64_t)m * H + j0) * bw.rows_pad + rid;
              __stcs(hr, acc[0][0]);
              __stcs(hr + bw.rows_pad, acc[0][1]);
            }
          }
          const uint32_t b = g % NBUF, u = g / NBUF;
          if (u > 0) mbar_wait(empty0 + 8 * b, (u - 1) & 1);      // MMAs of the previous use of this buffer are done
          const uint32_t t_hi = tmem_base + lane_base + C::C_A + C::A_STRIDE * b + (uint32_t)(UPT * quarter);
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            uint32_t h0, l0, h1, l1;
            split_tf32(acc[c][0], h0, l0);
            split_tf32(acc[c][1], h1, l1);
            tmem_st2(t_hi + 8u * c, h0, h1);
            tmem_st2(t_hi + 8u * c + C::A_LO, l0, l1);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(full0 + 8 * b);
        }
        // parked features: stored here, while the tensor pipe drains, not ahead of the tile's first chunk
        if constexpr (!BW && CH == 5) {
          if (st_act && pass == 0) {
            float* fr = bw.fT + (int64_t)m * 32 * bw.rows_pad + rid;
            const float fv[4] = {1.0f, x, y, z};
#pragma unroll
            for (int f = 0; f < 4; ++f)
              if (f == quarter) fr[(int64_t)f * bw.rows_pad] = fv[f];
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              float dd, ex;
              if constexpr (C::FEAT_SMEM) { const float4 fa = featA[a * ROWS + row]; dd = fa.x; ex = fa.y; }
              else { dd = fdd[a]; ex = fex[a]; }
              if (((4 + a) & 3) == quarter) fr[(int64_t)(4 + a) * bw.rows_pad] = dd;
              if (((4 + NA + a) & 3) == quarter) fr[(int64_t)(4 + NA + a) * bw.rows_pad] = ex;
            }
          }
        }
        // ---- epilogue of this pass: output units [NT pass, NT pass + NT) ----------------------------------------
        mbar_wait(dfull, ep & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int sub = 0; sub < EPC / 8; ++sub) {
          const int c0 = EPC * quarter + 8 * sub;
          float v[CH][8];
#pragma unroll
          for (int c = 0; c < CH; ++c) tmem_ld8(tmem_base + lane_base + (uint32_t)(c * NT + c0), v[c]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (sub == EPC / 8 - 1) {                               // all accumulator reads of this warp are done
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(dempty);
          }
          const int jb = pass * NT + c0;
#pragma unroll
          for (int j = 0; j < 8; j += 2) {                      // two columns per packed instruction
            const float2 h = nq_tanh2(nq_add2(make_float2(v[0][j], v[0][j + 1]), *reinterpret_cast<const float2*>(b2s + jb + j)));
            const float2 w3 = *reinterpret_cast<const float2*>(W3s + jb + j);
            const float2 tp = nq_fma2(make_float2(-h.x, -h.y), h, nq_bc2(1.0f));
            if constexpr (BW) {
              const float2 gw = nq_fma2(nq_bc2(seed), h, make_float2(gW3[8 * sub + j], gW3[8 * sub + j + 1]));
              gW3[8 * sub + j] = gw.x; gW3[8 * sub + j + 1] = gw.y;
              const float2 d2 = nq_mul2(nq_mul2(nq_bc2(seed), w3), tp);
              float* dst = bw.dT + ((int64_t)m * H + jb + j) * bw.rows_pad + rid;
              dst[0] = d2.x;
              dst[bw.rows_pad] = d2.y;
              continue;
            }
            if constexpr (!BW && CH == 5) {
              if (st_act) {
                float* dst = bw.dT + ((int64_t)m * H + jb + j) * bw.rows_pad + rid;
                __stcs(dst, h.x);
                __stcs(dst + bw.rows_pad, h.y);
              }
            }
            out2[0] = nq_fma2(w3, h, out2[0]);
            if constexpr (CH >= 4) {
              const float2 v1 = make_float2(v[1][j], v[1][j + 1]), v2 = make_float2(v[2][j], v[2][j + 1]);
              const float2 v3 = make_float2(v[3][j], v[3][j + 1]), vl = make_float2(v[CH - 1][j], v[CH - 1][j + 1]);
              float2 g2;
              if constexpr (GW) g2 = make_float2(quad(v1.x, v2.x, v3.x), quad(v1.y, v2.y, v3.y));
              else g2 = nq_fma2(v1, v1, nq_fma2(v2, v2, nq_mul2(v3, v3)));
              const float2 w3tp = nq_mul2(w3, tp);
              out2[1] = nq_fma2(w3tp, v1, out2[1]);
              out2[2] = nq_fma2(w3tp, v2, out2[2]);
              out2[3] = nq_fma2(w3tp, v3, out2[3]);
              out2[CH - 1] = nq_fma2(w3tp, nq_fma2(make_float2(-(h.x + h.x), -(h.y + h.y)), g2, vl), out2[CH - 1]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) out[c] = out2[c].x + out2[c].y;
      if constexpr (BW) {
        if (quarter == 0) gb3 += seed;
        continue;
      }
      if (quarter != 0) {
#pragma unroll
        for (int c = 0; c < CH; ++c) outS[((quarter - 1) * ROWS + row) * CH + c] = out[c];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TCW) : "memory");    // workers only: partial outputs of the 4 quarters visible
      if (quarter == 0) {
        const int64_t wr = w0 + row;
        if (wr < W) {
          const int ent = ent_index(s, spin, i, korb);
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            float o = out[c] + (c == 0 ? b3 : 0.f);
#pragma unroll
            for (int q = 0; q < NQ - 1; ++q) o += outS[(q * ROWS + row) * CH + c];
            phi[((int64_t)c * s.n_ent + ent) * W + wr] = o;
          }
        }
      }
      // outS (and the feature table) are rewritten only after the next tile's first full[] round / feature barrier,
      // which every worker warp -- including the readers above -- must reach first
    }
    if constexpr (BW) {                                          // dW3 / db3: reduce the per-row partials over the 128 rows
      float* red = outS;                                         // [H + 1], unused in this mode
      for (int q = tid; q < H + 1; q += TCW) red[q] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(TCW) : "memory");
#pragma unroll
      for (int j = 0; j < C::EPC; ++j) {
        const float t = warp_sum(gW3[j]);
        if (lane == 0) atomicAdd(&red[C::EPC * quarter + j], t);
      }
      if (quarter == 0) {
        const float t = warp_sum(gb3);
        if (lane == 0) atomicAdd(&red[H], t);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TCW) : "memory");
      float* slab = bw.part + (int64_t)blockIdx.x * bw.p_mlp;
      const int o_b3 = 2 * H + F * H + H * H, o_W3 = o_b3 + 1;   // slab order: b1[H] W1[F*H] b2[H] W2[H*H] b3 W3[H]
      for (int q = tid; q < H; q += TCW) slab[o_W3 + q] = red[q];
      if (tid == 0) slab[o_b3] = red[H];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("bar.sync 2, %0;" ::"n"(TCT) : "memory");        // workers + the issuing warps
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
}

template <int H, int NA, int CH, bool GW = false, int NBUF = 2, bool BW = false>
int launch_t(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, const float* G, cudaStream_t st, const BwdOut bw = BwdOut{},
             int n_chunks_forced = 0) {
  using C = Cfg<H, NA, CH, GW, NBUF, BW>;
  const DevSys& s = ctx->sys;
  const size_t smem = C::smem_bytes(s.F);
  if (smem > 227 * 1024) return NQMC_ERR_UNSUPPORTED;
  if (C::NPASS > 1 && ctx->l1cache == nullptr) return NQMC_ERR_UNSUPPORTED;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(orb_fwd_pc_kernel<H, NA, CH, GW, NBUF, BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  if (max_items >= (int64_t)1 << 31) return NQMC_ERR_UNSUPPORTED;   // ItemIter counts items in 32 bits
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  if (n_chunks_forced > 0) n_chunks = n_chunks_forced;
  {
    ProfScope ps(ctx, BW ? NQ_PROF_BWD : (CH == 1 ? NQ_PROF_EVAL1 : NQ_PROF_EVAL5), st);
    if (ctx->prof_on)
      orb_fwd_pc_kernel<H, NA, CH, GW, NBUF, BW><<<n_chunks * s.n_mlp, TCT, smem, st>>>(s, ctx->theta, r, W, phi, n_chunks, G, bw, ctx->l1cache);
    else
      NQMC_CUDA(nq_launch_pdl_guarded(ctx, orb_fwd_pc_kernel<H, NA, CH, GW, NBUF, BW>, dim3(n_chunks * s.n_mlp), dim3(TCT), smem, st, s,
                                      ctx->theta, r, W, phi, n_chunks, G, bw, ctx->l1cache));
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

template <int H, int CH, bool GW>
int dispatch_na(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, const float* G, cudaStream_t st, const BwdOut bw = BwdOut{}) {
  switch (ctx->sys.A) {                // the number of nuclei is a template parameter: the feature loop is unrolled
    case 1: return launch_t<H, 1, CH, GW>(ctx, r, W, phi, G, st, bw);
    case 2: return launch_t<H, 2, CH, GW>(ctx, r, W, phi, G, st, bw);
    case 4: return launch_t<H, 4, CH, GW>(ctx, r, W, phi, G, st, bw);
    case 6: return launch_t<H, 6, CH, GW>(ctx, r, W, phi, G, st, bw);
    case 8: return launch_t<H, 8, CH, GW>(ctx, r, W, phi, G, st, bw);
    case 10: return launch_t<H, 10, CH, GW>(ctx, r, W, phi, G, st, bw);
    default: return NQMC_ERR_UNSUPPORTED;
  }
}

}  // namespace

// Reverse-pass value sweep (hidden width 128): see BwdOut.  n_chunks CTAs per network, slab index = blockIdx.x.
int nq_launch_orb_fwd_bw128(nqmc_ctx* ctx, const float* r, const float* wgt, const double* hdr, const float* e_l, int64_t W,
                            float* hT, float* dT, float* fT, int64_t rows_pad, int n_chunks, cudaStream_t st) {
  if (ctx->sys.H != 128) return NQMC_ERR_UNSUPPORTED;
  const BwdOut bw{ctx->inv, wgt, hdr, e_l, hT, dT, fT, ctx->part_mlp, rows_pad, ctx->p_mlp};
  switch (ctx->sys.A) {
    case 1: return launch_t<128, 1, 1, false, 2, true>(ctx, r, W, nullptr, nullptr, st, bw, n_chunks);
    case 2: return launch_t<128, 2, 1, false, 2, true>(ctx, r, W, nullptr, nullptr, st, bw, n_chunks);
    case 4: return launch_t<128, 4, 1, false, 2, true>(ctx, r, W, nullptr, nullptr, st, bw, n_chunks);
    case 6: return launch_t<128, 6, 1, false, 2, true>(ctx, r, W, nullptr, nullptr, st, bw, n_chunks);
    case 8: return launch_t<128, 8, 1, false, 2, true>(ctx, r, W, nullptr, nullptr, st, bw, n_chunks);
    case 10: return launch_t<128, 10, 1, false, 2, true>(ctx, r, W, nullptr, nullptr, st, bw, n_chunks);
    default: return NQMC_ERR_UNSUPPORTED;
  }
}

// Value (channels == 1) or 5-channel pass.  Without G: hidden width 128 only (width 64 has its own kernels).  With G
// (backflow, G-weighted Laplacian channel): width 64 or 128.  Returns NQMC_ERR_UNSUPPORTED otherwise (the caller falls
// back to the FP32 FFMA kernels).
// ctx->want_act (set by the fused walker step around its energy pass; 5-channel pass, width 128, reverse-pass scratch
// reserved): also park h1 / h2 / features (see BwdOut) and report it in ctx->act_stored.
int nq_launch_orb_fwd_pc(nqmc_ctx* ctx, const float* r, int64_t W, int channels, float* phi, const float* G, cudaStream_t st) {
  const int H = ctx->sys.H;
  BwdOut bw{};
  const bool park = ctx->want_act && channels == 5 && H == 128 && ctx->bw_h != nullptr && ctx->bw_h2 != nullptr &&
                    nq_bwd128_rows(ctx, W) <= ctx->bw_rows;
  if (park) { bw.hT = ctx->bw_h; bw.dT = ctx->bw_h2; bw.fT = ctx->bw_f; bw.rows_pad = ctx->bw_rows; bw.part = ctx->part_mlp; }
  int rc = NQMC_ERR_UNSUPPORTED;
  if (G != nullptr) {
    if (channels != 5 || 3 + 2 * ctx->sys.A > H) return NQMC_ERR_UNSUPPORTED;
    if (H == 64 && ctx->sys.A <= 6) return dispatch_na<64, 5, true>(ctx, r, W, phi, G, st);
    if (H == 128) rc = dispatch_na<128, 5, true>(ctx, r, W, phi, G, st, bw);
  } else if (H == 128) {
    if (channels == 1) return dispatch_na<128, 1, false>(ctx, r, W, phi, nullptr, st);
    if (channels == 5) rc = dispatch_na<128, 5, false>(ctx, r, W, phi, nullptr, st, bw);
  }
  if (rc == NQMC_OK && park) ctx->act_stored = true;
  return rc;
}
