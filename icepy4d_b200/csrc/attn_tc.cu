// Flash-style multi-head attention on sm_100a tensor cores (tcgen05 + TMEM + TMA), head_dim 64, bf16 operands,
// f32 softmax statistics, f32 accumulation in TMEM.  The Nq x Nk probability matrix never leaves the SM.
//
//   O[q, h*64 : h*64+64] = softmax_k( scale * Q_h[q] . K_h[k] ) V_h[k]
//
// Q, K, V live in ONE row-major bf16 buffer X [rows, ld] (e.g. the fused QKV projection output): a "problem" z is given by
// row ranges (q_row0, Nq), (k_row0, Nk) and column offsets (q_col, k_col, v_col); head h adds 64*h columns.  One TMA tensor map
// over X serves all three operands.  gridDim = (ceil(maxNq/128), heads, n_problems) so both images of a SuperGlue /
// LightGlue layer (self or cross) run in one launch.
//
// At head_dim 64 the exponentials, not the tensor cores, are the critical resource: a 128 x 128 score block costs 512 clk of
// tcgen05.mma but 1024 clk of MUFU.EX2 (16 / clk / SM, measured).  The kernel is therefore organised around keeping the
// MUFU pipe busy: one CTA per SM (128 queries x 1 head, 320 threads) with TWO softmax warp groups that ping-pong over the
// key blocks, so one group's TMEM loads / row maxima / barrier traffic hide behind the other group's exponentials.
//   warp 0      TMA producer: Q once, then K_j and V_j blocks of 128 keys through two 3-stage rings
//   warp 1      MMA issuer (one elected lane): S_j = Q K_j^T (M128 N128 K64) into TMEM S[j & 1] (two score buffers);
//               O += P_j V_j (M128 N64 K128, V is the MN-major B operand) accumulating in TMEM; issue order
//               ... PV_{j-1}, QK_{j+2}, PV_j ... so the scores of a group's next block are ready when it returns
//   warps 2..5  softmax group 0 (even key blocks), warps 6..9 group 1 (odd key blocks): ONE thread per query row holds the
//               block's 128 scores in registers (read from TMEM once, the buffer is released immediately).  The running row
//               maximum is shared between the two groups through shared memory (block j publishes m_j, block j+1 consumes
//               it: a short handshake before the long exponential phase) and is LAZY: O (in TMEM) and the row sums are
//               rescaled only when the maximum grows by more than 2^8.  p = exp2(s*c - m) with packed f32x2 arithmetic
//               (FFMA2 / FADD2 / FMNMX3), packed to bf16 and stored in the 128B-swizzled K-major layout the MMA reads,
//               into one of two P buffers.
//
// Reference behaviour replaced: `attention()` + MultiHeadedAttention of thirdparty/SuperGlue/models/superglue.py:87-116
// (materialises a 4 x N x M f32 tensor) and Attention/SelfBlock/CrossBlock of thirdparty/LightGlue/lightglue/lightglue.py:92-216.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/icepy4d_b200.h"

#define FA_BM 128
#define FA_BN 128
#define FA_D 64
#define FA_KV_STAGES 3
#define FA_THREADS 640                          // warpgroup 0: TMA producer, MMA issuer, two idle warps; warpgroups 1-4: softmax
#define FA_Q_BYTES (FA_BM * FA_D * 2)           // 16 KB
#define FA_KV_BYTES (FA_BN * FA_D * 2)          // 16 KB each for K and V
#ifndef FA_P_TMEM
#define FA_P_TMEM 1                             // 1: P_j goes to tensor memory (A operand of the P V product read from TMEM)
#endif
#if FA_P_TMEM
#define FA_P_BYTES 0
#else
#define FA_P_BYTES (FA_BM * FA_BN * 2)          // 32 KB (two 16 KB K-halves), two buffers
#endif
#define FA_SMEM_BYTES (FA_Q_BYTES + 2 * FA_KV_STAGES * FA_KV_BYTES + 2 * FA_P_BYTES + 1024)   // Q | K ring | V ring (| P x2) (+ align)
#define FA_TMEM_COLS 512                        // S0: [0,128)  S1: [128,256)  O: [256,320)  P0: [320,384)  P1: [384,448) (bf16 pairs)
#define FA_TMEM_P 320
#define FA_MAX_PROBLEMS 4
#define FA_TAU 8.0f                             // lazy-rescale threshold (log2 units)
// Share of the exponentials evaluated on the FMA pipe instead of MUFU (Cody-Waite range reduction + degree-3 minimax
// polynomial, relative error 7.7e-5 — far below the bf16 rounding of P): bit e of the mask selects pair e of every 8-element
// chunk; even / odd chunks use the low / high nibble.  MUFU.EX2 (16 / clk / SM) is the critical pipe of this kernel.
#ifndef FA_POLY_MASK
#define FA_POLY_MASK 0x00
#endif

struct AttnProblem { int q_row0, nq, k_row0, nk; };
struct AttnParams {
  AttnProblem prob[FA_MAX_PROBLEMS];
  int q_col, k_col, v_col;
  float scale_log2;                 // scale * log2(e)
  __nv_bfloat16* O; int ldo;        // O rows are indexed like Q rows (q_row0 + i)
  // work decomposition: item = (problem * heads + head) * n_qt + q-tile.  CTAs [0, n_whole) run whole items; the items behind
  // them — the ragged last wave of the one-CTA-per-SM schedule — are split in two along the keys (CTA pairs) and merged by
  // whichever half finishes second, so that wave costs half a tile time instead of a whole one.
  int n_qt, heads, n_whole;
  float* part;                      // per half: O [128][64] f32 (unnormalised, relative to m), m [128], l [128]
  int* counters;                    // per split item: arrival counter (zeroed before the launch)
};
#define FA_PART_FLOATS (FA_BM * FA_D + 2 * FA_BM)
#define FA_MAX_SPLITS 128

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a pair on the FMA / ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 polynomial; 2^n by adding n
// to the exponent field (the magic-number addition leaves n in the low mantissa bits of r)
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float magic = 12582912.f;                                   // 1.5 * 2^23
  x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
  const float2 r = __fadd2_rn(x, make_float2(magic, magic));
  const float2 fi = __fadd2_rn(r, make_float2(-magic, -magic));
  const float2 f = __fadd2_rn(x, make_float2(-fi.x, -fi.y));
  float2 pl = __ffma2_rn(f, make_float2(0.05508868f, 0.05508868f), make_float2(0.24260405f, 0.24260405f));
  pl = __ffma2_rn(pl, f, make_float2(0.69327624f, 0.69327624f));
  pl = __ffma2_rn(pl, f, make_float2(0.99992894f, 0.99992894f));
  return make_float2(__int_as_float(__float_as_int(pl.x) + (__float_as_int(r.x) << 23)),
                     __int_as_float(__float_as_int(pl.y) + (__float_as_int(r.y) << 23)));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void tmem_ld32x(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int POLY_MASK>
__global__ void __launch_bounds__(FA_THREADS, 1) attn_tc_kernel(const __grid_constant__ CUtensorMap tmX, AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FA_Q_BYTES;                                   // K ring: stage s at sK + s*16K
  uint8_t* sV = sK + FA_KV_STAGES * FA_KV_BYTES;                   // V ring
  uint8_t* sP = sV + FA_KV_STAGES * FA_KV_BYTES;                   // P buffers: block j -> sP + (j & 1) * 32K
  __shared__ __align__(8) uint64_t q_full, k_full[FA_KV_STAGES], k_empty[FA_KV_STAGES], v_full[FA_KV_STAGES], v_empty[FA_KV_STAGES];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], p_full[2], pv_done[2], m_ready[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int merge_flag_s;
  __shared__ float mrun_s[FA_BM];                                  // running row maximum after the latest block (log2 units)
  __shared__ float lsum_s[4][FA_BM];                               // per (group, column half) row sums for the final exchange
  __shared__ float xch[2][2][2][FA_BM];                            // [group][block parity of the group][column half][row]: row-max exchange

  int item = blockIdx.x, split = -1, half = 0;
  if (item >= p.n_whole) {
    const int k = item - p.n_whole;
    split = k >> 1; half = k & 1;
    item = p.n_whole + split;
  }
  const int qt = item % p.n_qt, h = (item / p.n_qt) % p.heads;
  AttnProblem pr = p.prob[0];                                     // (a dynamic index would put the parameter array on the stack)
  {
    const int z = item / (p.n_qt * p.heads);
#pragma unroll
    for (int i = 1; i < FA_MAX_PROBLEMS; ++i)
      if (z == i) pr = p.prob[i];
  }
  const int q0 = qt * FA_BM;
  if (q0 >= pr.nq) return;                                        // uniform per CTA (both halves of a split item): safe before any barrier
  if (split >= 0) {                                               // my half of the key blocks
    const int nb0 = ((pr.nk + FA_BN - 1) / FA_BN + 1) >> 1;
    if (half == 0) pr.nk = min(pr.nk, nb0 * FA_BN);
    else { pr.k_row0 += nb0 * FA_BN; pr.nk -= nb0 * FA_BN; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (pr.nk + FA_BN - 1) / FA_BN;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX);
    tc::mbar_init(&q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) {
      tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1);
      tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&s_full[b], 1); tc::mbar_init(&s_empty[b], 256); tc::mbar_init(&p_full[b], 256);
      tc::mbar_init(&pv_done[b], 1); tc::mbar_init(&m_ready[b], 256);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, FA_TMEM_COLS);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_O = tmem + 256;

  // Register budget by role (setmaxnreg works on whole warpgroups): the launch gives every thread 96 registers (640 threads =
  // 61440); the four softmax warpgroups grow to 104 (128 * 96 + 512 * 104 = 65536) so that the 64 scores, the packed
  // probabilities and the running statistics of a row half stay in registers (no local-memory traffic in the loop).
  if (warp >= 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(&q_full, FA_Q_BYTES);
      tc::tma_load_2d(sQ, &tmX, &q_full, p.q_col + h * FA_D, pr.q_row0 + q0);
      auto load_k = [&](int j) {
        const int s = j % FA_KV_STAGES;
        tc::mbar_wait(&k_empty[s], ((j / FA_KV_STAGES) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&k_full[s], FA_KV_BYTES);
        tc::tma_load_2d(sK + s * FA_KV_BYTES, &tmX, &k_full[s], p.k_col + h * FA_D, pr.k_row0 + j * FA_BN);
      };
      auto load_v = [&](int j) {
        const int s = j % FA_KV_STAGES;
        tc::mbar_wait(&v_empty[s], ((j / FA_KV_STAGES) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&v_full[s], FA_KV_BYTES);
        tc::tma_load_2d(sV + s * FA_KV_BYTES, &tmX, &v_full[s], p.v_col + h * FA_D, pr.k_row0 + j * FA_BN);
      };
      for (int j = 0; j < FA_KV_STAGES && j < nblk; ++j) load_k(j);
      for (int j = 0; j < nblk; ++j) {
        load_v(j);
        if (j + FA_KV_STAGES < nblk) load_k(j + FA_KV_STAGES);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (one elected lane: operands stay in uniform registers)
    if (tc::elect_one()) {
      constexpr uint32_t idesc_qk = tc::make_idesc(FA_BM, FA_BN, 0, 0, 1);   // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = tc::make_idesc(FA_BM, FA_D, 0, 1, 1);    // A = P (K-major), B = V (MN-major)
      constexpr uint32_t hi_k = tc::desc_hi_sw128(1024);                     // K-major operands and MN-major V: SBO = 1024
      const uint32_t dQ = tc::desc_lo_sw128(tc::smem_u32(sQ)), dP0 = tc::desc_lo_sw128(tc::smem_u32(sP));
      const uint32_t dK0 = tc::desc_lo_sw128(tc::smem_u32(sK));
      // V descriptor: MN-major, 8-key groups 1024 B apart (SBO), one 64-wide N atom: LBO field = 1024 >> 4 as well
      const uint32_t dV0 = ((tc::smem_u32(sV) >> 4) & 0x3FFF) | ((1024u >> 4) << 16);
      auto issue_qk = [&](int j) {                                          // S[j & 1] = Q K_j^T
        const int s = j % FA_KV_STAGES;
        tc::mbar_wait(&k_full[s], (j / FA_KV_STAGES) & 1);
        if (j >= 2) tc::mbar_wait(&s_empty[j & 1], ((j - 2) >> 1) & 1);    // S_{j-2} has been read into registers
        tc::tcgen05_fence_after();
        const uint32_t dK = dK0 + (uint32_t)(s * (FA_KV_BYTES >> 4));
        const uint32_t tS = tmem + (uint32_t)(j & 1) * FA_BN;
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) tc::umma_f16_parts(tS, dQ + k * 2, hi_k, dK + k * 2, hi_k, idesc_qk, k ? 1u : 0u);
        tc::umma_commit(&k_empty[s]);                                       // K_j no longer needed once these retire
        tc::umma_commit(&s_full[j & 1]);
      };
      tc::mbar_wait(&q_full, 0);
      issue_qk(0);
      if (nblk > 1) issue_qk(1);
      for (int j = 0; j < nblk; ++j) {
        // QK_{j+2} goes in front of PV_j: its score buffer S[j & 1] is released early in block j (right after the TMEM load),
        // so the scores are ready when the group that owns blocks j, j + 2, ... returns from its exponentials
        if (j + 2 < nblk) issue_qk(j + 2);
        const int s = j % FA_KV_STAGES;
        tc::mbar_wait(&v_full[s], (j / FA_KV_STAGES) & 1);
        tc::mbar_wait(&p_full[j & 1], (j >> 1) & 1);                        // P_j in shared memory, O rescaled if needed
        tc::tcgen05_fence_after();
        const uint32_t dV = dV0 + (uint32_t)(s * (FA_KV_BYTES >> 4));
#if FA_P_TMEM
        const uint32_t tP = tmem + FA_TMEM_P + (uint32_t)(j & 1) * (FA_BN / 2);
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)      // A: P k-slice = 16 keys = 8 TMEM columns;  B: V rows [16k, 16k+16) x 64 dims
          tc::umma_f16_ts(tmem_O, tP + (uint32_t)(k * 8), dV + (uint32_t)((k * 16 * 128) >> 4), hi_k, idesc_pv, (j | k) ? 1u : 0u);
        (void)dP0;
#else
        const uint32_t dP = dP0 + (uint32_t)((j & 1) * (FA_P_BYTES >> 4));
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          // A: P k-slice = 16 keys = 32 B inside the 128-B swizzle row of K-half (k / 4);  B: V rows [16k, 16k+16) x 64 dims
          tc::umma_f16_parts(tmem_O, dP + (uint32_t)(((k >> 2) * (FA_BM * 128) + (k & 3) * 32) >> 4), hi_k,
                             dV + (uint32_t)((k * 16 * 128) >> 4), hi_k, idesc_pv, (j | k) ? 1u : 0u);
        }
#endif
        tc::umma_commit(&v_empty[s]);                                       // V stage free
        tc::umma_commit(&pv_done[j & 1]);                                   // O includes block j; P buffer free
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------ softmax: group g (8 warps) owns key blocks j = g, g + 2, ...;
    // TWO threads per query row (TMEM lane), 64 keys each: 16 softmax warps per SM keep the issue slots and the MUFU pipe fed
    const int sw = warp - 4;                                                // 0..15
    const int g = sw >> 3;
    const int hf = (sw >> 2) & 1;                                           // column half: keys [64 hf, 64 hf + 64) of the block
    const int quarter = warp & 3;                                           // TMEM lane quarter this warp may access
    const int q = quarter * 32 + lane;                                      // TMEM lane == tile row
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t rsw = (uint32_t)(q & 7);
    const uint32_t prow0 = tc::smem_u32(sP) + (uint32_t)hf * (FA_BM * 128) + (uint32_t)(q >> 3) * 1024 + (uint32_t)(q & 7) * 128;
    const int pair_bar = 2 + g * 4 + quarter;                               // named barrier of the two warps that share my rows
    float m_mine = -INFINITY, l_part = 0.f;                                 // l_part is relative to m_mine
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);

    for (int j = g; j < nblk; j += 2) {
      const uint32_t b = (uint32_t)j & 1u, ph = ((uint32_t)j >> 1) & 1u;
      tc::mbar_wait(&s_full[b], ph);
      tc::tcgen05_fence_after();
      uint32_t v[64];
      const uint32_t tS = tmem + lane_off + b * FA_BN + hf * 64;
      tmem_ld32x(tS, v); tmem_ld32x(tS + 32, v + 32);
      tc::tmem_ld_wait();
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&s_empty[b]);                                         // QK_{j+2} may overwrite this score buffer now
      const int kvalid = pr.nk - j * FA_BN - hf * 64;                       // keys of my half that exist
      if (kvalid < 64) {                                                    // only in the last block
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kvalid) v[i] = 0xff800000u;                              // -inf
      }
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 64; i += 2) mxa[(i >> 1) & 3] = fmax3(mxa[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
      const int xp = (j >> 1) & 1;
      xch[g][xp][hf][q] = mx;
      named_bar_sync(pair_bar, 64);                                         // the two warps that share this lane quarter
      mx = fmaxf(mx, xch[g][xp][hf ^ 1][q]);
      const float m_blk = mx * p.scale_log2;
      // ---- running maximum handshake with the other group (lazy: move only when the block exceeds it by more than 2^TAU)
      float m_prev = -INFINITY, m_new = m_blk;
      bool need = false;
      if (j > 0) {
        tc::mbar_wait(&m_ready[b ^ 1], (((uint32_t)(j - 1)) >> 1) & 1u);
        m_prev = mrun_s[q];
        if (m_blk > m_prev + FA_TAU) need = true; else m_new = m_prev;
        named_bar_sync(pair_bar, 64);                                       // my partner has read m_prev before I overwrite it
      }
      if (hf == 0) mrun_s[q] = m_new;
      tc::mbar_arrive(&m_ready[b]);
      if (m_new != m_mine) {                                                // bring my partial row sum to the new reference
        l_part = (m_mine == -INFINITY) ? 0.f : l_part * ex2_approx(m_mine - m_new);
        m_mine = m_new;
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // O (TMEM) is relative to m_prev: rescale my half of the row once PV_{j-1} has retired (PV_j cannot start before my
        // p_full arrival)
        tc::mbar_wait(&pv_done[b ^ 1], (((uint32_t)(j - 1)) >> 1) & 1u);
        tc::tcgen05_fence_after();
        const float alpha = need ? ex2_approx(m_prev - m_new) : 1.f;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t ov[16];
          tmem_ld16(tmem_O + lane_off + hf * 32 + hh * 16, ov);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st16(tmem_O + lane_off + hf * 32 + hh * 16, ov);
          tmem_st_wait();
        }
      }
      if (j >= 2) tc::mbar_wait(&pv_done[b], (((uint32_t)(j - 2)) >> 1) & 1u);   // P buffer b free (PV_{j-2} retired)
      // p = exp2(s * c - m); f32 row sum; bf16 pack; swizzled store (8 chunks of 16 B = my K-half row)
      float2 rs2 = make_float2(0.f, 0.f);
      const float2 nm2 = make_float2(-m_new, -m_new);
      const uint32_t prow = prow0 + b * FA_P_BYTES;
#if FA_P_TMEM
      uint32_t pk_all[16];
#endif
#pragma unroll
      for (int t = 0; t < 8; ++t) {
#if FA_P_TMEM
        uint32_t* pk = pk_all + (t & 3) * 4;
#else
        uint32_t pk[4];
#endif
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = t * 8 + e * 2;
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2);
          const bool poly = (((t & 1) ? (POLY_MASK >> 4) : POLY_MASK) >> e) & 1;      // compile-time after unrolling
          const float2 ab = (POLY_MASK & 0x100) ? x                                            // timing experiment: no exponential at all
                            : poly ? ex2_poly2(x) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
          rs2 = __fadd2_rn(rs2, ab);
          __nv_bfloat162 pr2 = __floats2bfloat162_rn(ab.x, ab.y);
          pk[e] = *reinterpret_cast<uint32_t*>(&pr2);
        }
#if FA_P_TMEM
        // my 64 keys of row q as bf16 pairs -> TMEM lane q, columns [32 hf, 32 hf + 32) of P[b], 16 columns (32 keys) at a time
        if ((t & 3) == 3) tmem_st16(tmem + lane_off + FA_TMEM_P + b * (FA_BN / 2) + hf * 32 + (t >> 2) * 16, pk_all);
#else
        const uint32_t addr = prow + ((((uint32_t)t) ^ rsw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
#endif
      }
      l_part += rs2.x + rs2.y;
#if FA_P_TMEM
      (void)prow; (void)rsw;
      tmem_st_wait();
      tc::tcgen05_fence_before();
#else
      tc::tcgen05_fence_before();
      tc::fence_proxy_async_smem();                                         // make P_j visible to the tensor-core proxy
#endif
      tc::mbar_arrive(&p_full[b]);
    }
    // ---- final: bring all four partial row sums to the final maximum, exchange, normalise O
    {
      const uint32_t jl = (uint32_t)(nblk - 1);
      tc::mbar_wait(&m_ready[jl & 1], (jl >> 1) & 1u);
      const float m_fin = mrun_s[q];
      lsum_s[g * 2 + hf][q] = (m_mine == -INFINITY) ? 0.f : l_part * ex2_approx(m_mine - m_fin);
      asm volatile("bar.sync 1, 512;" ::: "memory");                        // the sixteen softmax warps
      const float l = (lsum_s[0][q] + lsum_s[1][q]) + (lsum_s[2][q] + lsum_s[3][q]);
      tc::mbar_wait(&pv_done[jl & 1], (jl >> 1) & 1u);
      tc::tcgen05_fence_after();
      const int c0 = (g * 2 + hf) * 16;                                     // four threads per row: 16 of the 64 output dims each
      uint32_t ov[16];
      tmem_ld16(tmem_O + lane_off + c0, ov);
      tc::tmem_ld_wait();
      float inv = 1.f / l;
      bool store = true;
      if (split >= 0) {
        // ---- split item: publish my half (O, m, l); the half that arrives second merges both and writes the output
        float* mine = p.part + (size_t)(split * 2 + half) * FA_PART_FLOATS;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(mine + q * FA_D + c0 + i) =
              make_float4(__uint_as_float(ov[i]), __uint_as_float(ov[i + 1]), __uint_as_float(ov[i + 2]), __uint_as_float(ov[i + 3]));
        if (g == 0 && hf == 0) { mine[FA_BM * FA_D + q] = m_fin; mine[FA_BM * FA_D + FA_BM + q] = l; }
        __threadfence();
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == 128) merge_flag_s = atomicAdd(p.counters + split, 1);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        store = merge_flag_s != 0;                                          // uniform over the CTA
        if (store) {
          __threadfence();
          const float* oth = p.part + (size_t)(split * 2 + (half ^ 1)) * FA_PART_FLOATS;
          const float mo = __ldcg(oth + FA_BM * FA_D + q), lo = __ldcg(oth + FA_BM * FA_D + FA_BM + q);
          const float mm = fmaxf(m_fin, mo);
          const float wa = ex2_approx(m_fin - mm), wb = ex2_approx(mo - mm);
          inv = 1.f / (l * wa + lo * wb);
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 o4 = __ldcg(reinterpret_cast<const float4*>(oth + q * FA_D + c0 + i));
            ov[i] = __float_as_uint(__uint_as_float(ov[i]) * wa + o4.x * wb);
            ov[i + 1] = __float_as_uint(__uint_as_float(ov[i + 1]) * wa + o4.y * wb);
            ov[i + 2] = __float_as_uint(__uint_as_float(ov[i + 2]) * wa + o4.z * wb);
            ov[i + 3] = __float_as_uint(__uint_as_float(ov[i + 3]) * wa + o4.w * wb);
          }
        }
      }
      if (store && q0 + q < pr.nq) {
        __nv_bfloat16* dst = p.O + (size_t)(pr.q_row0 + q0 + q) * p.ldo + h * FA_D + c0;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          uint4 pk;
          __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
          __nv_bfloat162 c2 = __floats2bfloat162_rn(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
          __nv_bfloat162 d = __floats2bfloat162_rn(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
          pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b2);
          pk.z = *reinterpret_cast<uint32_t*>(&c2); pk.w = *reinterpret_cast<uint32_t*>(&d);
          *reinterpret_cast<uint4*>(dst + i) = pk;
        }
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, FA_TMEM_COLS);
}

extern "C" __attribute__((visibility("default"))) size_t i4d_attention_workspace_bytes(void) {
  return (size_t)FA_MAX_SPLITS * 2 * FA_PART_FLOATS * sizeof(float) + FA_MAX_SPLITS * sizeof(int) + 256;
}

extern "C" __attribute__((visibility("default"))) int i4d_attention_bf16_tc(
    const void* X, int rows, int ld, int q_col, int k_col, int v_col, int heads, const int* problems_host, int n_problems,
    float scale, void* O, int ldo, void* workspace, size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(X && O && problems_host, "null pointer");
  I4D_CHECK_ARG(n_problems >= 1 && n_problems <= FA_MAX_PROBLEMS && heads >= 1, "1..4 problems, heads >= 1");
  I4D_CHECK_ARG((ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "O must be 16-byte aligned with ldo % 8 == 0");
  I4D_CHECK_ARG(q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0, "column offsets must be multiples of 8");
  AttnParams p;
  int max_nq = 0, min_nk = 0x7fffffff;
  for (int z = 0; z < FA_MAX_PROBLEMS; ++z) {
    if (z < n_problems) {
      p.prob[z] = AttnProblem{problems_host[4 * z], problems_host[4 * z + 1], problems_host[4 * z + 2], problems_host[4 * z + 3]};
      I4D_CHECK_ARG(p.prob[z].nq >= 0 && p.prob[z].nk >= 1, "every problem needs nk >= 1");
      I4D_CHECK_ARG(p.prob[z].q_row0 >= 0 && p.prob[z].k_row0 >= 0 && p.prob[z].q_row0 + p.prob[z].nq <= rows &&
                    p.prob[z].k_row0 + p.prob[z].nk <= rows, "row ranges outside the buffer");
      if (p.prob[z].nq > max_nq) max_nq = p.prob[z].nq;
      if (p.prob[z].nk < min_nk) min_nk = p.prob[z].nk;
    } else {
      p.prob[z] = AttnProblem{0, 0, 0, 1};
    }
  }
  if (max_nq == 0) return I4D_OK;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.O = reinterpret_cast<__nv_bfloat16*>(O); p.ldo = ldo;
  CUtensorMap tmX;
  if (int rc = i4d_make_tmap_2d_bf16(&tmX, X, (uint64_t)rows, (uint64_t)ld, (uint64_t)ld, FA_BN, FA_D)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // one CTA per SM: n_items CTAs run in ceil(n_items / SMs) waves.  When the last wave fills less than half of the SMs its
  // items are split in two along the keys (needs the caller's workspace and >= 4 key blocks per problem)
  p.n_qt = i4d_cdiv(max_nq, FA_BM); p.heads = heads;
  const int n_items = p.n_qt * heads * n_problems, sms = i4d_num_sms();
  int rem = n_items % sms;
  if (!(workspace && workspace_bytes >= i4d_attention_workspace_bytes() && n_items > sms && rem > 0 && 2 * rem <= sms &&
        rem <= FA_MAX_SPLITS && min_nk >= 4 * FA_BN))
    rem = 0;
  p.n_whole = n_items - rem;
  p.part = nullptr; p.counters = nullptr;
  if (rem) {
    uint8_t* w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    p.counters = reinterpret_cast<int*>(w);
    p.part = reinterpret_cast<float*>(w + ((FA_MAX_SPLITS * sizeof(int) + 255) & ~(size_t)255) - 0);
    I4D_CUDA_CALL(cudaMemsetAsync(p.counters, 0, FA_MAX_SPLITS * sizeof(int), st));
  }
  // share of exponentials on the FMA pipe: none by default (measured: no gain, the kernel is latency-bound around the MUFU pipe,
  // DESIGN.md); I4D_FA_POLY=25 selects the 25 % build for experiments
  static int variant = -1;
  static bool attr_seen[64] = {};
  if (variant < 0) {
    const char* e = getenv("I4D_FA_POLY");
    const int v = e ? atoi(e) : 0;
    variant = v == 25 ? 1 : v == 50 ? 2 : v == 75 ? 3 : v == 100 ? 4 : v == -1 ? 5 : 0;
  }
  if (i4d_first_use_on_device(attr_seen)) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(attn_tc_kernel<0x00>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
    I4D_CUDA_CALL(cudaFuncSetAttribute(attn_tc_kernel<0x88>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
    I4D_CUDA_CALL(cudaFuncSetAttribute(attn_tc_kernel<0xAA>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
    I4D_CUDA_CALL(cudaFuncSetAttribute(attn_tc_kernel<0xEE>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
    I4D_CUDA_CALL(cudaFuncSetAttribute(attn_tc_kernel<0xFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
    I4D_CUDA_CALL(cudaFuncSetAttribute(attn_tc_kernel<0x100>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
  }
  const int grid = n_items + rem;
  switch (variant) {
    case 1: attn_tc_kernel<0x88><<<grid, FA_THREADS, FA_SMEM_BYTES, st>>>(tmX, p); break;
    case 2: attn_tc_kernel<0xAA><<<grid, FA_THREADS, FA_SMEM_BYTES, st>>>(tmX, p); break;
    case 3: attn_tc_kernel<0xEE><<<grid, FA_THREADS, FA_SMEM_BYTES, st>>>(tmX, p); break;
    case 4: attn_tc_kernel<0xFF><<<grid, FA_THREADS, FA_SMEM_BYTES, st>>>(tmX, p); break;
    case 5: attn_tc_kernel<0x100><<<grid, FA_THREADS, FA_SMEM_BYTES, st>>>(tmX, p); break;
    default: attn_tc_kernel<0x00><<<grid, FA_THREADS, FA_SMEM_BYTES, st>>>(tmX, p);
  }
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
