// nqmc_orbital_lap_pc.cu -- 5-channel (value, gradient, Laplacian) orbital forward pass on tcgen05, third variant:
// a producer / consumer pipeline with NO block-wide barrier in the steady state.
//
// The two earlier variants (A operand in shared memory; A operand in TMEM with lockstep phases -- both removed in
// round 2) ran all 16 worker warps in lockstep through "layer 1 -> stage -> __syncthreads -> MMA -> wait", so the FMA
// pipe, the TMEM/shared-memory ports and the tensor pipe are each busy only in their own phase (ncu: issue slots
// ~55 %, FMA pipe ~39 %, tensor pipe ~23 %).  Here the warps are decoupled with mbarriers:
//
//   worker warp (16 of them, thread = (row, quarter)), per 8-unit K-chunk (8 chunks per 128-row tile):
//       layer 1 for 2 hidden units x 5 channels (FFMA, features in registers) -> tanh rules -> tf32 hi/lo
//       wait  empty[b]      (the MMAs that read A buffer b two chunks ago have completed)
//       tcgen05.st -> A buffer b (TMEM, lane = row) ; one lane arrives on full[b]
//   issuer warp, per chunk: wait full[b] (16 warp arrivals) ; 5 channels x 3 products = 15 UMMAs (A from TMEM,
//       K = 8) ; tcgen05.commit -> empty[b] ; after the last chunk of a tile also -> dfull
//   epilogue (workers): wait dfull ; tcgen05.ld accumulators ; arrive on dempty ; activation rules, output layer ;
//       the four quarters of a row meet at a named barrier (workers only) to combine their partial outputs.
//   The issuer waits dempty before it overwrites the accumulators with the next tile's first chunk.
//
// TMEM: 5 x 64 accumulator columns + 2 x (5 channels x 8 units x {hi, lo}) A columns = 480.
#include "nqmc_internal.cuh"

namespace {

constexpr int H = 64;
constexpr int NQ = 4;
constexpr int TCW = 128 * NQ;
constexpr int NWW = TCW / 32;          // worker warps
constexpr int TCT = TCW + 128;         // 16 worker warps + the issuer warpgroup (one issuing warp; see setmaxnreg below)
constexpr int ROWS = 128;
constexpr int CH = 5;
constexpr int KC = 8;                  // units per K-chunk == one UMMA k-step
constexpr int NCHUNK = H / KC;         // 8
constexpr int UPT = KC / NQ;           // 2 units per thread per chunk
constexpr int EPC = H / NQ;            // 16 epilogue columns per thread
// One thread gets a tcgen05.mma out only every ~117 cycles whatever its shape (scripts/exp/mma_rate.cu: N = 32 ... 192,
// one or several accumulators), while the tensor pipe needs 32 cycles for an M128 N64 K8 one; several issuing warps
// scale.  With a single issuer the 120 UMMAs of a tile took 14 k cycles -- the whole tile time.  All four warps of the
// issuer warpgroup therefore issue, each for its own accumulators (channel c belongs to warp c % 4: the kc == 0
// overwrite and the accumulating products of one accumulator stay in one thread's program order).
constexpr int NISS = 4;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t C_A = 320, A_STRIDE = 80, A_LO = 40;   // buffer b: hi at C_A + 80 b + 8 c, lo 40 columns later
constexpr int B_K4 = H * 16;
constexpr int B_MAT = (H / 4) * B_K4;

inline size_t smem_bytes(int F) { return 2 * B_MAT + ((size_t)F * H + 3 * H + (NQ - 1) * ROWS * CH) * 4 + 128; }

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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

template <int NA>
__global__ void __launch_bounds__(TCT, 1)
    orb_lap_pc_kernel(const DevSys s, const float* __restrict__ theta, const float* __restrict__ r, const int64_t W,
                      float* __restrict__ phi, const int n_chunks) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* Bhi = smem_raw;                                    // [H/4][H n][4]
  uint8_t* Blo = Bhi + B_MAT;
  float* W1s = reinterpret_cast<float*>(Blo + B_MAT);         // [F][H]
  float* b1s = W1s + s.F * H;
  float* b2s = b1s + H;
  float* W3s = b2s + H;
  float* outS = W3s + H;                                      // [NQ-1][ROWS][CH]
  uint64_t* bars = reinterpret_cast<uint64_t*>(outS + (NQ - 1) * ROWS * CH);   // full[2], empty[2], dfull, dempty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

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
  const float b3 = theta[off.b3];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(NWW));   // full[0]
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[1])), "r"(NWW));   // full[1]
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[2])), "r"(NISS));  // empty[0]: one commit per issuing warp
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[3])), "r"(NISS));  // empty[1]
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[4])), "r"(NISS));  // dfull
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[5])), "r"(NWW));   // dempty
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
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]);
  const uint32_t dfull = smem_u32(&bars[4]), dempty = smem_u32(&bars[5]);

  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t n_items = nwb * ns;
  nq_pdl_wait();                              // everything above read theta only; the positions come from the predecessor

  if (!worker) {
    // =============================== MMA issuer warp =====================================================
    // register split (as in nqmc_orbital_bwd_tc.cu): 20 warps x 96 at launch -> 112 per worker, 32 per issuer thread
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    const int iw = warp - NWW;                                 // channel c is issued by warp c % NISS
    const uint32_t leader = elect_one();
    const uint64_t dBh = umma_desc(smem_u32(Bhi), B_K4, 128), dBl = umma_desc(smem_u32(Blo), B_K4, 128);
    uint32_t t = 0;                                            // local tile counter (global chunk number = 8 t + kc)
    for (int64_t item = chunk; item < n_items; item += n_chunks, ++t) {
      if (t > 0) mbar_wait(dempty, (t - 1) & 1);               // accumulators of the previous tile have been read
#pragma unroll 8
      for (int kc = 0; kc < NCHUNK; ++kc) {
        const uint32_t b = kc & 1;
        mbar_wait(full0 + 8 * b, (kc >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t kb = (uint64_t)(((uint32_t)kc * 2 * B_K4) >> 4);
        const uint32_t abase = tmem_base + C_A + A_STRIDE * b;
        for (int c = iw; c < CH; c += NISS) {
          const uint32_t d = tmem_base + (uint32_t)(c * H);
          const uint32_t ah = abase + 8u * c, al = ah + A_LO;
          umma_ts(d, ah, dBh + kb, kc == 0 ? 0u : 1u, leader);
          umma_ts(d, al, dBh + kb, 1u, leader);
          umma_ts(d, ah, dBl + kb, 1u, leader);
        }
        umma_commit(empty0 + 8 * b, leader);
        if (kc == NCHUNK - 1) umma_commit(dfull, leader);
        __syncwarp();
      }
    }
  } else {
    // =============================== worker warps ========================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_q = tmem_base + lane_base + C_A + (uint32_t)(UPT * quarter);
    const float* const w1q = W1s + UPT * quarter;
    const float* const b1q = b1s + UPT * quarter;
    uint32_t t = 0;
    float nx = 0.f, ny = 0.f, nz = 0.f;                        // positions of the next tile, fetched one tile ahead
    ItemIter cur(chunk, n_chunks, ns), nxt = cur;             // this tile / the next tile whose positions are requested
    auto fetch = [&]() {                                       // tiles are requested in order: chunk, chunk + n_chunks, ...
      int64_t w = (int64_t)nxt.wb * ROWS + row;
      if (w >= W) w = W - 1;
      const float* p = r + (int64_t)(3 * (elec0 + nxt.e)) * W + w;   // one address, then two strides of W
      nx = p[0];
      ny = p[W];
      nz = p[2 * W];
      nxt.next();
    };
    // per nucleus: d, exp(-d), unit vector, 2/d -- in registers for all 8 chunks.  With s = wd - e we the five
    // channels of (d, e) contribute  value: d wd + e we ; gradient: u s ; Laplacian: (2/d) s + e we   (7 FMA, not 10).
    float x = 0.f, y = 0.f, z = 0.f;
    float fdd[NA], fex[NA], fux[NA], fuy[NA], fuz[NA], fti[NA];
    auto features = [&]() {                                    // of the tile whose positions sit in (nx, ny, nz)
      x = nx; y = ny; z = nz;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const float dx = x - s.R[3 * a], dy = y - s.R[3 * a + 1], dz = z - s.R[3 * a + 2];
        const float dd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        const float inv = 1.0f / dd;
        fdd[a] = dd; fex[a] = __expf(-dd); fux[a] = dx * inv; fuy[a] = dy * inv; fuz[a] = dz * inv; fti[a] = 2.0f * inv;
      }
    };
    // AHEAD: the features of tile t+1 are formed while the last UMMAs of tile t are in flight (they stay live across the
    // epilogue: 6 registers per nucleus, which the 112-register budget has room for up to 4 nuclei)
    constexpr bool AHEAD = NA <= 4;
    if (chunk < n_items) {
      fetch();
      if (AHEAD) {
        features();
        if (chunk + n_chunks < n_items) fetch();
      }
    }
    for (int64_t item = chunk; item < n_items; item += n_chunks, ++t, cur.next()) {
      const int i = cur.e;
      const int64_t w0 = (int64_t)cur.wb * ROWS;
      if (!AHEAD) {
        features();
        if (item + n_chunks < n_items) fetch();
      }
      // fully unrolled: the global chunk number is 8 t + kc, so the staging buffer (kc & 1), both mbarrier parities and
      // every shared-memory / TMEM offset of a chunk are immediates (the 2x-unrolled loop spent ~20 of its 126
      // instructions per chunk re-deriving them from threadIdx)
#pragma unroll
      for (int kc = 0; kc < NCHUNK; ++kc) {
        const float* W1s = w1q + KC * kc;                       // shadows the CTA-wide pointers: this thread's two units
        const float* b1s = b1q + KC * kc;
        constexpr int j0 = 0;
        // the thread's two hidden units are one packed pair: every FMA / MUL / ADD below is one FFMA2 / FMUL2 / FADD2
        float2 acc[CH];
        {
          acc[0] = *reinterpret_cast<const float2*>(b1s + j0);
          float2 ew = make_float2(0.f, 0.f);                     // sum_a e_a we_a : part of the value and of the Laplacian
          acc[0] = nq_fma2(nq_bc2(x), *reinterpret_cast<const float2*>(W1s + 0 * H + j0), acc[0]);
          acc[1] = *reinterpret_cast<const float2*>(W1s + 0 * H + j0);
          acc[0] = nq_fma2(nq_bc2(y), *reinterpret_cast<const float2*>(W1s + 1 * H + j0), acc[0]);
          acc[2] = *reinterpret_cast<const float2*>(W1s + 1 * H + j0);
          acc[0] = nq_fma2(nq_bc2(z), *reinterpret_cast<const float2*>(W1s + 2 * H + j0), acc[0]);
          acc[3] = *reinterpret_cast<const float2*>(W1s + 2 * H + j0);
          acc[4] = make_float2(0.f, 0.f);
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            const float2 wd = *reinterpret_cast<const float2*>(W1s + (3 + a) * H + j0);
            const float2 we = *reinterpret_cast<const float2*>(W1s + (3 + NA + a) * H + j0);
            ew = nq_fma2(nq_bc2(fex[a]), we, ew);
            acc[0] = nq_fma2(nq_bc2(fdd[a]), wd, acc[0]);
            const float2 sv = nq_fma2(nq_bc2(-fex[a]), we, wd);
            acc[1] = nq_fma2(nq_bc2(fux[a]), sv, acc[1]);
            acc[2] = nq_fma2(nq_bc2(fuy[a]), sv, acc[2]);
            acc[3] = nq_fma2(nq_bc2(fuz[a]), sv, acc[3]);
            acc[4] = nq_fma2(nq_bc2(fti[a]), sv, acc[4]);
          }
          // tanh rules: lap' = tp (lap - 2 h |grad|^2)
          const float2 h = nq_tanh2(nq_add2(acc[0], ew));
          const float2 tp = nq_fma2(make_float2(-h.x, -h.y), h, nq_bc2(1.0f));
          const float2 g2 = nq_fma2(acc[1], acc[1], nq_fma2(acc[2], acc[2], nq_mul2(acc[3], acc[3])));
          acc[4] = nq_mul2(tp, nq_fma2(make_float2(-(h.x + h.x), -(h.y + h.y)), g2, nq_add2(acc[4], ew)));
          acc[1] = nq_mul2(acc[1], tp);
          acc[2] = nq_mul2(acc[2], tp);
          acc[3] = nq_mul2(acc[3], tp);
          acc[0] = h;
        }
        const uint32_t b = kc & 1;                              // u = (8 t + kc) >> 1 uses of this buffer so far
        if (t > 0 || kc >= 2) mbar_wait(empty0 + 8 * b, ((kc >> 1) + 1) & 1);   // MMAs of its previous use are done
        const uint32_t t_hi = t_q + A_STRIDE * b;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float2 hi = make_float2(__uint_as_float(__float_as_uint(acc[c].x) & 0xFFFFE000u),
                                        __uint_as_float(__float_as_uint(acc[c].y) & 0xFFFFE000u));
          const float2 lo = nq_add2(acc[c], make_float2(-hi.x, -hi.y));
          tmem_st2(t_hi + 8u * c, __float_as_uint(hi.x), __float_as_uint(hi.y));
          tmem_st2(t_hi + 8u * c + A_LO, __float_as_uint(lo.x), __float_as_uint(lo.y));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * b);
      }
      // the last chunks' UMMAs are still in flight: the features of the next tile (the chunk loop is done with this
      // tile's) fill the wait, and the positions of the tile after that are requested
      if (AHEAD && item + n_chunks < n_items) {
        features();
        if (item + 2 * (int64_t)n_chunks < n_items) fetch();
      }
      // ---- epilogue ------------------------------------------------------------------------------------------
      mbar_wait(dfull, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float2 out2[CH];                                         // even / odd columns of this thread, added at the end
#pragma unroll
      for (int c = 0; c < CH; ++c) out2[c] = make_float2(0.f, 0.f);
#pragma unroll
      for (int sub = 0; sub < EPC / 8; ++sub) {
        const int c0 = EPC * quarter + 8 * sub;
        float v[CH][8];
#pragma unroll
        for (int c = 0; c < CH; ++c) tmem_ld8(tmem_base + lane_base + (uint32_t)(c * H + c0), v[c]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (sub == EPC / 8 - 1) {                               // all accumulator reads of this warp are done
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(dempty);
        }
#pragma unroll
        for (int j = 0; j < 8; j += 2) {                        // two columns per packed instruction
          const float2 v0 = make_float2(v[0][j], v[0][j + 1]), v1 = make_float2(v[1][j], v[1][j + 1]);
          const float2 v2 = make_float2(v[2][j], v[2][j + 1]), v3 = make_float2(v[3][j], v[3][j + 1]);
          const float2 v4 = make_float2(v[4][j], v[4][j + 1]);
          const float2 h = nq_tanh2(nq_add2(v0, *reinterpret_cast<const float2*>(b2s + c0 + j)));
          const float2 tp = nq_fma2(make_float2(-h.x, -h.y), h, nq_bc2(1.0f));
          const float2 g2 = nq_fma2(v1, v1, nq_fma2(v2, v2, nq_mul2(v3, v3)));
          const float2 w3 = *reinterpret_cast<const float2*>(W3s + c0 + j), w3tp = nq_mul2(w3, tp);
          out2[0] = nq_fma2(w3, h, out2[0]);
          out2[1] = nq_fma2(w3tp, v1, out2[1]);
          out2[2] = nq_fma2(w3tp, v2, out2[2]);
          out2[3] = nq_fma2(w3tp, v3, out2[3]);
          out2[4] = nq_fma2(w3tp, nq_fma2(make_float2(-(h.x + h.x), -(h.y + h.y)), g2, v4), out2[4]);
        }
      }
      float out[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) out[c] = out2[c].x + out2[c].y;
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
      // outS is rewritten only after 8 more full[] rounds, which need every worker warp (incl. the readers above)
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("bar.sync 2, %0;" ::"n"(TCT) : "memory");        // workers + the issuing warps
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

template <int NA>
int launch_t(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  const size_t smem = smem_bytes(s.F);
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[ctx->device & 63];
  if (!configured) {
    NQMC_CUDA(cudaFuncSetAttribute(orb_lap_pc_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  const int64_t nwb = (W + ROWS - 1) / ROWS;
  const int64_t max_items = nwb * (s.n_up > s.n_dn ? s.n_up : s.n_dn);
  if (max_items >= (int64_t)1 << 31) return NQMC_ERR_UNSUPPORTED;   // ItemIter counts items in 32 bits
  int n_chunks = ctx->sm_count / s.n_mlp;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > max_items) n_chunks = (int)max_items;
  {
    ProfScope ps(ctx, NQ_PROF_EVAL5, st);
    if (ctx->prof_on) orb_lap_pc_kernel<NA><<<n_chunks * s.n_mlp, TCT, smem, st>>>(s, ctx->theta, r, W, phi, n_chunks);
    else NQMC_CUDA(nq_launch_pdl_guarded(ctx, orb_lap_pc_kernel<NA>, dim3(n_chunks * s.n_mlp), dim3(TCT), smem, st, s, ctx->theta, r, W, phi, n_chunks));
  }
  NQMC_LAUNCH_CHECK();
  return NQMC_OK;
}

}  // namespace

// returns NQMC_ERR_UNSUPPORTED when this variant does not apply (the caller falls back to the other variants)
int nq_launch_orb_lap_pc(nqmc_ctx* ctx, const float* r, int64_t W, float* phi, cudaStream_t st) {
  const DevSys& s = ctx->sys;
  if (s.H != H) return NQMC_ERR_UNSUPPORTED;
  switch (s.A) {                       // the number of nuclei is a template parameter: features live in registers
    case 1: return launch_t<1>(ctx, r, W, phi, st);
    case 2: return launch_t<2>(ctx, r, W, phi, st);
    case 3: return launch_t<3>(ctx, r, W, phi, st);
    case 4: return launch_t<4>(ctx, r, W, phi, st);
    case 6: return launch_t<6>(ctx, r, W, phi, st);
    default: return NQMC_ERR_UNSUPPORTED;
  }
}
