// Flash-style multi-head attention on sm_100a tensor cores (tcgen05 + TMEM + TMA), head_dim 64, bf16 operands,
// f32 softmax statistics, f32 accumulation in TMEM.  Neither the Nq x Nk score matrix nor the probabilities ever leave the SM.
//
//   O[q, h*64 : h*64+64] = softmax_k( scale * Q_h[q] . K_h[k] ) V_h[k]
//
// Q, K, V live in ONE row-major bf16 buffer X [rows, ld] (e.g. the fused QKV projection output): a "problem" z is given by
// row ranges (q_row0, Nq), (k_row0, Nk) and column offsets (q_col, k_col, v_col); head h adds 64*h columns.  One TMA tensor map
// over X serves all three operands; both images of a SuperGlue / LightGlue layer (self or cross) run in one launch.
//
// Organisation (one persistent CTA per SM, 640 threads):
//   work item   = (problem, head, 256-query tile); its unit of work = one block of 128 keys.  The total number of units is cut
//                 into equal contiguous ranges, one per CTA (stream-K): a CTA runs 1-3 segments (item, key-block range); items cut
//                 by a range boundary publish unnormalised partial results (O, m, l) and the part that arrives last merges.
//   warp 0      TMA producer: the two 128-row Q tiles (A, B) of the item, then K_j / V_j blocks through two 4-stage rings —
//               every K / V block is loaded once for 256 queries
//   warps 1, 2  MMA issuers of tile A / tile B (one elected lane each): S_t(j) = Q_t K_j^T (M128 N128 K64) into the tile's score
//               buffer in TMEM; O_t += P_t(j) V_j (M128 N64 K128) with the A operand P_t read FROM TENSOR MEMORY (bf16 pairs),
//               V the MN-major B operand; per tile the order is QK(j+1), PV(j): the next scores are computed while the softmax
//               group is busy with the exponentials of the current block.  One issuer per tile, because a single in-order issuer
//               chains the tiles (QK_B(j+1) behind the wait for P_A(j)): measured 1430 clk per block with the exponentials
//               switched off, against 512 clk of tensor work
//   warps 4-11  softmax group of tile A, warps 12-19 of tile B (warp 3 idle): the two groups are INDEPENDENT — own score
//               buffer, own O accumulator, own running maxima — so nothing couples them and they settle half a block apart:
//               one group's TMEM loads / row maxima hide behind the other's exponentials.  Two threads per query row (64 keys
//               each): scores TMEM -> registers (buffer released at once), row max exchanged with the partner thread, lazy
//               rescale of O (only when the maximum grows by more than 2^8), p = exp2(s*c - m) with packed f32x2 arithmetic,
//               packed to bf16 and written back to TMEM with tcgen05.st.
// TMEM map (512 columns): S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)  P_A [384,448)  P_B [448,512).
// At head_dim 64 a 128 x 128 score block costs 512 clk of tcgen05.mma but 1024 clk of MUFU.EX2 (16 / clk / SM): the exponentials
// bound the kernel at half of the tensor peak; POLY_MASK moves a share of them to the FMA pipe.
//
// Reference behaviour replaced: `attention()` + MultiHeadedAttention of thirdparty/SuperGlue/models/superglue.py:87-116
// (materialises a 4 x N x M f32 tensor) and Attention/SelfBlock/CrossBlock of thirdparty/LightGlue/lightglue/lightglue.py:92-216.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/icepy4d_b200.h"

#define FA_BM 128                               // rows per query tile (one softmax group)
#define FA_BN 128                               // keys per block
#define FA_D 64
#define FA_KV_STAGES 4
#define FA_THREADS 640
#define FA_Q_BYTES (FA_BM * FA_D * 2)           // 16 KB per query tile
#define FA_KV_BYTES (FA_BN * FA_D * 2)          // 16 KB each for K and V
#define FA_SMEM_BYTES (2 * FA_Q_BYTES + 2 * FA_KV_STAGES * FA_KV_BYTES + 1024)   // Q_A Q_B | K ring | V ring = 160 KB (+ align)
#define FA_TMEM_COLS 512
#define FA_TMEM_O 256
#define FA_TMEM_P 384
#define FA_MAX_PROBLEMS 4
#define FA_TAU 8.0f                             // lazy-rescale threshold (log2 units)
#define FA_MIN_SEG 4                            // a CTA's share is at least this many key blocks (bounds the parts per item)
#define FA_MAX_ITEMS 8192                       // merge counters in the workspace
#define FA_PART_FLOATS (2 * FA_BM * FA_D + 4 * FA_BM)   // per part: O [256][64] f32 (unnormalised, relative to m), m [256], l [256]
// Share of the exponentials evaluated on the FMA pipe instead of MUFU (Cody-Waite range reduction + degree-3 minimax
// polynomial, relative error 7.7e-5 — far below the bf16 rounding of P): bit e of the mask selects pair e of every 8-element
// chunk; even / odd chunks use the low / high nibble.  0x100 = no exponential at all (timing experiments only).

// Two measured-and-rejected variants stay behind compile-time switches (numbers in profiles/r2_attn_experiments.txt, session 3):
//   -DFA_WARP_ARRIVE   softmax-side mbarrier arrivals by one lane per warp instead of every thread: the two tiles then run their
//                      exponential phases TOGETHER (and idle the MUFU pipe together): 209 us against 196-203
//   -DFA_TURNS=1       explicit turn-taking of the two tiles on the MUFU pipe (A(n), B(n), A(n+1), ...): the phases alternate as
//                      intended, but the loads / row maxima of one tile now compete with the other's exponentials: 216 us
#ifdef FA_WARP_ARRIVE
#define FA_ARRIVALS 8
#define FA_ARRIVE(addr) do { __syncwarp(); if (lane == 0) tc::mbar_arrive_a(addr); } while (0)
#else
#define FA_ARRIVALS 256
#define FA_ARRIVE(addr) tc::mbar_arrive_a(addr)
#endif
#ifndef FA_SOFTMAX_REGS
#define FA_SOFTMAX_REGS 104             // 0 = no setmaxnreg (every thread keeps the 96 registers of the launch)
#endif
#ifndef FA_ISSUER_REGS
#define FA_ISSUER_REGS 40
#endif
#ifndef FA_TURNS
#define FA_TURNS 0
#endif
struct AttnProblem { int q_row0, nq, k_row0, nk; };
struct AttnParams {
  AttnProblem prob[FA_MAX_PROBLEMS];
  int unit_off[FA_MAX_PROBLEMS + 1];   // first global unit of problem z (unit = one key block of one item); [n] = total
  int item_off[FA_MAX_PROBLEMS + 1];   // first global item of problem z
  int n_qt[FA_MAX_PROBLEMS];           // 256-row query tiles
  int nblk[FA_MAX_PROBLEMS];           // key blocks
  int heads, n_ctas, balanced;         // balanced = 1: unit-balanced ranges (needs the workspace); 0: whole items per CTA
  int q_col, k_col, v_col;
  float scale_log2;                    // scale * log2(e)
  __nv_bfloat16* O; int ldo;           // O rows are indexed like Q rows (q_row0 + i)
  float* part;                         // [2 * n_ctas] part slots: slot 2c = CTA c's first segment, 2c + 1 = its last
  int* counters;                       // per item: parts arrived (zeroed before the launch)
  const int* nk_dev;                   // optional: per problem, how many of the nk keys are real (device memory; the rest is masked)
};

// Timeline instrumentation (scripts/attn_trace.py builds a second library with -DFA_TRACE): CTA 0 records %clock64 at the
// synchronisation points of one softmax warp per tile for its first FA_TRACE_MAX key blocks.
#ifdef FA_TRACE
#define FA_TRACE_MAX 96
__device__ unsigned long long fa_trace_buf[2][FA_TRACE_MAX][8];
#define FA_STAMP(k) do { if (blockIdx.x == 0 && (sw & 7) == 0 && lane == 0 && bb + jj < FA_TRACE_MAX) { unsigned long long c_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); fa_trace_buf[t][bb + jj][k] = c_; } } while (0)
extern "C" __attribute__((visibility("default"))) int i4d_attention_trace_dump(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, fa_trace_buf, sizeof(fa_trace_buf));
}
#else
#define FA_STAMP(k) do { } while (0)
#endif
#define FA_DEFAULT_POLY 25                      // a quarter of the exponentials on the FMA pipe: 1-4 % faster than all-MUFU on every box measured (profiles/r2_attn_experiments.txt)

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


// global unit range [u0, u1) of CTA c
__device__ __forceinline__ long long fa_range_start(const AttnParams& p, int c) {
  const int total = p.unit_off[FA_MAX_PROBLEMS];
  if (p.balanced) return (long long)c * total / p.n_ctas;
  // whole items: item range [c I / G, (c+1) I / G) expressed in units
  const int n_items = p.item_off[FA_MAX_PROBLEMS];
  const int it = (int)((long long)c * n_items / p.n_ctas);
  int z = 0;
#pragma unroll
  for (int i = 1; i < FA_MAX_PROBLEMS; ++i)
    if (it >= p.item_off[i] && p.item_off[i] < n_items) z = i;
  if (it >= n_items) return total;
  return (long long)p.unit_off[z] + (long long)(it - p.item_off[z]) * p.nblk[z];
}
// CTA whose (balanced) range contains unit u
__device__ __forceinline__ int fa_cta_of_unit(const AttnParams& p, long long u) {
  const long long total = p.unit_off[FA_MAX_PROBLEMS];
  return (int)(((u + 1) * p.n_ctas + total - 1) / total) - 1;
}

// ---- one segment of this CTA's unit range: item (z, h, qt), key blocks [kb0, kb1).  Every role runs its own copy of the segment
// loop (the role branches never re-join, so each keeps the register budget setmaxnreg gave it); the loops stay in step through
// the block-wide barrier at the end of every segment.
#define FA_SEGMENT_DECODE \
    int z = 0; \
    _Pragma("unroll") \
    for (int i = 1; i < FA_MAX_PROBLEMS; ++i) \
      if (u >= p.unit_off[i] && p.unit_off[i] < p.unit_off[FA_MAX_PROBLEMS]) z = i; \
    AttnProblem pr = p.prob[0]; \
    int nblk = p.nblk[0], n_qt = p.n_qt[0], uoff = p.unit_off[0], ioff = p.item_off[0]; \
    _Pragma("unroll") \
    for (int i = 1; i < FA_MAX_PROBLEMS; ++i) \
      if (z == i) { pr = p.prob[i]; nblk = p.nblk[i]; n_qt = p.n_qt[i]; uoff = p.unit_off[i]; ioff = p.item_off[i]; } \
    if (p.nk_dev != nullptr) pr.nk = min(pr.nk, __ldg(p.nk_dev + z)); \
    const int r = (int)(u - uoff), it = r / nblk, kb0 = r - it * nblk; \
    const long long item_start = (long long)uoff + (long long)it * nblk; \
    const int kb1 = (int)min((long long)nblk, (long long)kb0 + (u_end - u)); \
    const int nb = kb1 - kb0; \
    const bool whole = kb0 == 0 && kb1 == nblk; \
    const int h = it / n_qt, qt = it - h * n_qt; \
    const int q0 = qt * (2 * FA_BM); \
    const bool validB = q0 + FA_BM < pr.nq; \
    const int slot = (u == u_begin) ? 0 : 1;
// next segment: the softmax groups have O_t in registers (they waited for the last PV), Q / O / P may be overwritten
#define FA_SEGMENT_ADVANCE \
    u = item_start + kb1; \
    ++seg; kv_base += (uint32_t)nb; blk_base[0] += (uint32_t)nb; if (validB) { blk_base[1] += (uint32_t)nb; turn_base += (uint32_t)nb; } \
    tc::tcgen05_fence_before(); \
    __syncthreads(); \
    tc::tcgen05_fence_after();

template <int POLY_MASK>
__global__ void __launch_bounds__(FA_THREADS, 1) attn_tc_kernel(const __grid_constant__ CUtensorMap tmX, AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                              // Q_A | Q_B
  uint8_t* sK = sQ + 2 * FA_Q_BYTES;                               // K ring: stage s at sK + s*16K
  uint8_t* sV = sK + FA_KV_STAGES * FA_KV_BYTES;                   // V ring
  __shared__ __align__(8) uint64_t q_full, k_full[FA_KV_STAGES], k_empty[FA_KV_STAGES], v_full[FA_KV_STAGES], v_empty[FA_KV_STAGES];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], p_full[2], pv_done[2], turn[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int merge_flag_s;
  __shared__ float lsum_s[2][2][FA_BM];                            // [tile][column half][row]: row sums for the final exchange
  __shared__ float xch[2][2][2][FA_BM];                            // [tile][block parity][column half][row]: row-max exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const long long u_begin = fa_range_start(p, cta), u_end = fa_range_start(p, cta + 1);

  // Barriers are initialised ONCE; their phases run on across the segments of this CTA (re-initialising a live mbarrier is
  // undefined).  K / V stages are released by two arrivals: one per tile issuer, or both from tile A's issuer when the item has no
  // rows for tile B.
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX);
    tc::mbar_init(&q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) {
      tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 2);
      tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      tc::mbar_init(&s_full[t], 1); tc::mbar_init(&s_empty[t], FA_ARRIVALS); tc::mbar_init(&p_full[t], FA_ARRIVALS);
      tc::mbar_init(&pv_done[t], 1); tc::mbar_init(&turn[t], FA_ARRIVALS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, FA_TMEM_COLS);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;

  uint32_t seg = 0, kv_base = 0, blk_base[2] = {0u, 0u};           // segments / K-V blocks / blocks per tile processed so far
  uint32_t turn_base = 0;                                          // key blocks of segments in which BOTH tiles ran (MUFU turn-taking)
  // Register re-partition: the producer / issuer warpgroup (warps 0-3) hands registers to the four softmax warpgroups
  // (96 -> FA_SOFTMAX_REGS per thread): the softmax loop keeps its barrier / exchange addresses in registers instead of
  // recomputing them in every key block.
  if (warp < 4) {
#if FA_SOFTMAX_REGS
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FA_ISSUER_REGS));   // (96 - 40) x 128 registers freed >= (104 - 96) x 512 taken
#endif
    if (warp == 0) {
      for (long long u = u_begin; u < u_end;) {
        FA_SEGMENT_DECODE
          // ------------------------------------------------ TMA producer
          if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(&q_full, validB ? 2 * FA_Q_BYTES : FA_Q_BYTES);
            tc::tma_load_2d(sQ, &tmX, &q_full, p.q_col + h * FA_D, pr.q_row0 + q0);
            if (validB) tc::tma_load_2d(sQ + FA_Q_BYTES, &tmX, &q_full, p.q_col + h * FA_D, pr.q_row0 + q0 + FA_BM);
            auto load_k = [&](int jj) {
              const uint32_t g = kv_base + (uint32_t)jj, s = g % FA_KV_STAGES;
              tc::mbar_wait(&k_empty[s], ((g / FA_KV_STAGES) & 1u) ^ 1u);
              tc::mbar_arrive_expect_tx(&k_full[s], FA_KV_BYTES);
              tc::tma_load_2d(sK + s * FA_KV_BYTES, &tmX, &k_full[s], p.k_col + h * FA_D, pr.k_row0 + (kb0 + jj) * FA_BN);
            };
            auto load_v = [&](int jj) {
              const uint32_t g = kv_base + (uint32_t)jj, s = g % FA_KV_STAGES;
              tc::mbar_wait(&v_empty[s], ((g / FA_KV_STAGES) & 1u) ^ 1u);
              tc::mbar_arrive_expect_tx(&v_full[s], FA_KV_BYTES);
              tc::tma_load_2d(sV + s * FA_KV_BYTES, &tmX, &v_full[s], p.v_col + h * FA_D, pr.k_row0 + (kb0 + jj) * FA_BN);
            };
            for (int jj = 0; jj < FA_KV_STAGES && jj < nb; ++jj) load_k(jj);
            for (int jj = 0; jj < nb; ++jj) {
              load_v(jj);
              if (jj + FA_KV_STAGES < nb) load_k(jj + FA_KV_STAGES);
            }
          }
          __syncwarp();
        FA_SEGMENT_ADVANCE
      }
    } else {
      for (long long u = u_begin; u < u_end;) {
        FA_SEGMENT_DECODE
        if (warp == 1 || (warp == 2 && validB)) {
          // ------------------------------------------------ MMA issuers: warp 1 drives tile A, warp 2 tile B (one elected lane each;
          // operands stay in uniform registers).  The tiles share nothing but the K / V stages, so neither issuer ever waits for the
          // other tile's softmax group: S_t(j+1) is issued the moment the group has S_t(j) in registers.
          if (tc::elect_one()) {
            const int t = warp - 1;
            constexpr uint32_t idesc_qk = tc::make_idesc(FA_BM, FA_BN, 0, 0, 1);   // A = Q (K-major), B = K (K-major)
            constexpr uint32_t idesc_pv = tc::make_idesc(FA_BM, FA_D, 0, 1, 1);    // A = P (TMEM, K-major), B = V (MN-major)
            constexpr uint32_t hi_k = tc::desc_hi_sw128(1024);                     // K-major operands and MN-major V: SBO = 1024
            const uint32_t dQ = tc::desc_lo_sw128(tc::smem_u32(sQ)) + (uint32_t)(t * (FA_Q_BYTES >> 4));
            const uint32_t dK0 = tc::desc_lo_sw128(tc::smem_u32(sK));
            // V descriptor: MN-major, 8-key groups 1024 B apart (SBO), one 64-wide N atom: LBO field = 1024 >> 4 as well
            const uint32_t dV0 = ((tc::smem_u32(sV) >> 4) & 0x3FFF) | ((1024u >> 4) << 16);
            const uint32_t tS = tmem + (uint32_t)t * FA_BN;
            const uint32_t tO = tmem + FA_TMEM_O + (uint32_t)t * FA_D, tP = tmem + FA_TMEM_P + (uint32_t)t * (FA_BN / 2);
            const uint32_t bb = blk_base[t];                                      // this tile's block counter at the segment start
            const bool twice = !validB;                                           // tile A alone: both stage-release arrivals are mine
            auto issue_qk = [&](int jj) {                                         // S_t = Q_t K_jj^T
              const uint32_t s = (kv_base + (uint32_t)jj) % FA_KV_STAGES;
              const uint32_t dK = dK0 + s * (FA_KV_BYTES >> 4);
    #pragma unroll
              for (int k = 0; k < FA_D / 16; ++k) tc::umma_f16_parts(tS, dQ + k * 2, hi_k, dK + k * 2, hi_k, idesc_qk, k ? 1u : 0u);
              tc::umma_commit(&s_full[t]);
              tc::umma_commit(&k_empty[s]);
              if (twice) tc::umma_commit(&k_empty[s]);
            };
            tc::mbar_wait(&q_full, seg & 1u);
            tc::mbar_wait(&k_full[kv_base % FA_KV_STAGES], (kv_base / FA_KV_STAGES) & 1u);
            tc::tcgen05_fence_after();
            issue_qk(0);
            for (int jj = 0; jj < nb; ++jj) {
              const uint32_t g = kv_base + (uint32_t)jj, s = g % FA_KV_STAGES;
              if (jj + 1 < nb) {
                // scores of the NEXT block: the buffer is free as soon as the group has block jj in registers
                tc::mbar_wait(&k_full[(g + 1) % FA_KV_STAGES], ((g + 1) / FA_KV_STAGES) & 1u);
                tc::mbar_wait(&s_empty[t], (bb + (uint32_t)jj) & 1u);
                tc::tcgen05_fence_after();
                issue_qk(jj + 1);
              }
              tc::mbar_wait(&v_full[s], (g / FA_KV_STAGES) & 1u);
              tc::mbar_wait(&p_full[t], (bb + (uint32_t)jj) & 1u);                 // P_t(jj) in TMEM, O_t rescaled if needed
              tc::tcgen05_fence_after();
              const uint32_t dV = dV0 + s * (FA_KV_BYTES >> 4);
    #pragma unroll
              for (int k = 0; k < FA_BN / 16; ++k)        // A: P k-slice = 16 keys = 8 TMEM columns;  B: V rows [16k, 16k+16) x 64 dims
                tc::umma_f16_ts(tO, tP + (uint32_t)(k * 8), dV + (uint32_t)((k * 16 * 128) >> 4), hi_k, idesc_pv, (jj | k) ? 1u : 0u);
              tc::umma_commit(&pv_done[t]);                                        // O_t includes block jj; P_t free
              tc::umma_commit(&v_empty[s]);
              if (twice) tc::umma_commit(&v_empty[s]);
            }
          }
          __syncwarp();
        }
        FA_SEGMENT_ADVANCE
      }
    }
  } else {
#if FA_SOFTMAX_REGS
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FA_SOFTMAX_REGS));
#endif
    // ------------------------------------------------ softmax: tile t (8 warps), TWO threads per query row, 64 keys each.
    // Thread constants are computed once and pinned in registers (the block loop is latency-bound: every address the compiler
    // re-derives from %tid in front of a barrier or an exchange sits on the serial chain of the block).
    const int sw = warp - 4;                                                // 0..15
    const int t = sw >> 3;
    const int hf = (sw >> 2) & 1;                                           // column half: keys [64 hf, 64 hf + 64) of the block
    const int quarter = warp & 3;                                           // TMEM lane quarter this warp may access
    const int q = quarter * 32 + lane;                                      // TMEM lane == tile row
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int pair_bar = (int)tc::keep_in_register((uint32_t)(2 + t * 4 + quarter));   // named barrier of the two warps that share my rows
    const uint32_t tS = tc::keep_in_register(tmem + lane_off + (uint32_t)t * FA_BN + (uint32_t)hf * 64);
    const uint32_t tO = tc::keep_in_register(tmem + lane_off + FA_TMEM_O + (uint32_t)t * FA_D + (uint32_t)hf * 32);   // my 32 of the 64 output dims
    const uint32_t tP = tc::keep_in_register(tmem + lane_off + FA_TMEM_P + (uint32_t)t * (FA_BN / 2) + (uint32_t)hf * 32);
    const uint32_t a_s_full = tc::keep_in_register(tc::smem_u32(&s_full[t])), a_s_empty = tc::keep_in_register(tc::smem_u32(&s_empty[t]));
    const uint32_t a_p_full = tc::keep_in_register(tc::smem_u32(&p_full[t])), a_pv_done = tc::keep_in_register(tc::smem_u32(&pv_done[t]));
    const uint32_t a_xch_mine = tc::keep_in_register(tc::smem_u32(&xch[t][0][hf][q]));       // block parity ph adds ph * sizeof(xch[t][0])
    const uint32_t a_xch_part = tc::keep_in_register(tc::smem_u32(&xch[t][0][hf ^ 1][q]));
    for (long long u = u_begin; u < u_end;) {
      FA_SEGMENT_DECODE
        const bool tile_valid = t == 0 || validB;
        float m_run = -INFINITY, l_part = 0.f;                                  // l_part (my 64-key halves) is relative to m_run
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        uint32_t ov[32];
        float l = 1.f;
        const uint32_t bb = blk_base[t];
        const int kv_left = (int)tc::keep_in_register((uint32_t)(pr.nk - kb0 * FA_BN - hf * 64));

        if (tile_valid) {
          for (int jj = 0; jj < nb; ++jj) {
            const uint32_t ph = (bb + (uint32_t)jj) & 1u;
            FA_STAMP(0);
            tc::mbar_wait_a(a_s_full, ph);
            FA_STAMP(1);
            tc::tcgen05_fence_after();
            uint32_t v[64];
            tmem_ld32x(tS, v); tmem_ld32x(tS + 32, v + 32);
            tc::tmem_ld_wait();
            tc::tcgen05_fence_before();
            FA_ARRIVE(a_s_empty);                                       // QK_t(jj + 1) may overwrite the score buffer now
            const int kvalid = kv_left - jj * FA_BN;                            // keys of my half that exist
            if (kvalid < 64) {                                                  // only in the problem's last block
  #pragma unroll
              for (int i = 0; i < 64; ++i)
                if (i >= kvalid) v[i] = 0xff800000u;                            // -inf
            }
            float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  #pragma unroll
            for (int i = 0; i < 64; i += 2) mxa[(i >> 1) & 3] = fmax3(mxa[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
            tc::sts_f32(a_xch_mine + ph * (uint32_t)sizeof(xch[0][0]), mx);
            FA_STAMP(2);
            named_bar_sync(pair_bar, 64);                                       // the two warps that share this lane quarter
            FA_STAMP(3);
            mx = fmaxf(mx, tc::lds_f32(a_xch_part + ph * (uint32_t)sizeof(xch[0][0])));
            const float m_blk = mx * p.scale_log2;
            // lazy running maximum: move only when the block exceeds it by more than 2^TAU (both partner threads decide alike)
            const bool need = jj > 0 && m_blk > m_run + FA_TAU;
            const float m_new = (jj == 0 || need) ? fmaxf(m_blk, -1e30f) : m_run;   // finite even if the block holds no real key
            if (jj > 0) {
              tc::mbar_wait_a(a_pv_done, ph ^ 1u);                              // PV_t(jj-1) retired: O_t complete, P_t free
              FA_STAMP(4);
              if (__any_sync(0xffffffffu, need)) {
                tc::tcgen05_fence_after();
                const float alpha = need ? ex2_approx(m_run - m_new) : 1.f;     // O_t and the row sum are relative to m_run
  #pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                  uint32_t o16[16];
                  tmem_ld16(tO + hh * 16, o16);
                  tc::tmem_ld_wait();
  #pragma unroll
                  for (int i = 0; i < 16; ++i) o16[i] = __float_as_uint(__uint_as_float(o16[i]) * alpha);
                  tmem_st16(tO + hh * 16, o16);
                }
                tmem_st_wait();
                l_part *= alpha;
              }
            }
            m_run = m_new;
            // p = exp2(s * c - m); f32 row sum; bf16 pairs -> TMEM lane q, columns [32 hf, 32 hf + 32) of P_t, 16 columns at a time
            float2 rs2 = make_float2(0.f, 0.f);
            const float2 nm2 = make_float2(-m_new, -m_new);
            uint32_t pk[16];
            // experiment (off): the two tiles take turns on the SM's MUFU pipe, A(n), B(n), A(n+1), ...
            if (FA_TURNS && validB) {
              const uint32_t n = turn_base + (uint32_t)jj;
              if (t == 1) tc::mbar_wait(&turn[0], n & 1u);                      // A finished its n-th exponential phase
              else if (n > 0) tc::mbar_wait(&turn[1], (n - 1u) & 1u);           // B finished its (n-1)-th
            }
  #pragma unroll
            for (int tt = 0; tt < 8; ++tt) {
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = tt * 8 + e * 2;
                const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2);
                const bool poly = (((tt & 1) ? (POLY_MASK >> 4) : POLY_MASK) >> e) & 1;    // compile-time after unrolling
                const float2 ab = (POLY_MASK & 0x100) ? x                                  // timing experiment: no exponential at all
                                  : poly ? ex2_poly2(x) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
                rs2 = __fadd2_rn(rs2, ab);
                __nv_bfloat162 pr2 = __floats2bfloat162_rn(ab.x, ab.y);
                pk[(tt & 3) * 4 + e] = *reinterpret_cast<uint32_t*>(&pr2);
              }
              if ((tt & 3) == 3) tmem_st16(tP + (tt >> 2) * 16, pk);
            }
            l_part += rs2.x + rs2.y;
            if (FA_TURNS && validB) FA_ARRIVE(tc::smem_u32(&turn[t]));                        // the pipe is the other tile's
            FA_STAMP(5);
            tmem_st_wait();
            tc::tcgen05_fence_before();
            FA_ARRIVE(a_p_full);
            FA_STAMP(6);
          }
          // ---- row sum of both halves, O_t (my 32 dims) into registers
          lsum_s[t][hf][q] = l_part;
          named_bar_sync(pair_bar, 64);
          l = l_part + lsum_s[t][hf ^ 1][q];
          tc::mbar_wait_a(a_pv_done, (bb + (uint32_t)(nb - 1)) & 1u);
          tc::tcgen05_fence_after();
          tmem_ld32x(tO, ov);
          tc::tmem_ld_wait();
          tc::tcgen05_fence_before();
        }
        float m_fin = m_run;
        bool store = true;
        if (!whole) {
          // ---- split item: publish my part (O, m, l); the part that arrives last merges all of them and writes the output
          const int row = t * FA_BM + q;
          const long long item_end = item_start + nblk;
          const int c_first = fa_cta_of_unit(p, item_start), c_last = fa_cta_of_unit(p, item_end - 1);
          float* mine = p.part + (size_t)(2 * cta + slot) * FA_PART_FLOATS;
          if (tile_valid) {
  #pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(mine + (size_t)row * FA_D + hf * 32 + i) =
                  make_float4(__uint_as_float(ov[i]), __uint_as_float(ov[i + 1]), __uint_as_float(ov[i + 2]), __uint_as_float(ov[i + 3]));
            if (hf == 0) { mine[2 * FA_BM * FA_D + row] = m_fin; mine[2 * FA_BM * FA_D + 2 * FA_BM + row] = l; }
          }
          __threadfence();
          asm volatile("bar.sync 1, 512;" ::: "memory");                        // the sixteen softmax warps
          if (threadIdx.x == 128) merge_flag_s = atomicAdd(p.counters + ioff + it, 1);
          asm volatile("bar.sync 1, 512;" ::: "memory");
          store = merge_flag_s == c_last - c_first;                             // uniform over the CTA: I am the last part
          if (store && tile_valid) {
            // The parts are folded in CTA order c_first .. c_last WHATEVER part arrived last (the merger re-reads its own part from
            // global memory like everybody else's), so the result does not depend on the arrival order: bit-reproducible output.
            __threadfence();
            for (int c2 = c_first; c2 <= c_last; ++c2) {
              const long long rs = fa_range_start(p, c2);
              const int slot2 = (rs >= item_start) ? 0 : 1;                     // the item is CTA c2's first segment iff its range starts inside it
              const float* oth = p.part + (size_t)(2 * c2 + slot2) * FA_PART_FLOATS;
              const float mo = __ldcg(oth + 2 * FA_BM * FA_D + row), lo = __ldcg(oth + 2 * FA_BM * FA_D + 2 * FA_BM + row);
              if (c2 == c_first) {
                m_fin = mo; l = lo;
  #pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 o4 = __ldcg(reinterpret_cast<const float4*>(oth + (size_t)row * FA_D + hf * 32 + i));
                  ov[i] = __float_as_uint(o4.x); ov[i + 1] = __float_as_uint(o4.y); ov[i + 2] = __float_as_uint(o4.z); ov[i + 3] = __float_as_uint(o4.w);
                }
                continue;
              }
              const float mm = fmaxf(m_fin, mo);
              const float wa = ex2_approx(m_fin - mm), wb = ex2_approx(mo - mm);
              l = l * wa + lo * wb;
              m_fin = mm;
  #pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 o4 = __ldcg(reinterpret_cast<const float4*>(oth + (size_t)row * FA_D + hf * 32 + i));
                ov[i] = __float_as_uint(__uint_as_float(ov[i]) * wa + o4.x * wb);
                ov[i + 1] = __float_as_uint(__uint_as_float(ov[i + 1]) * wa + o4.y * wb);
                ov[i + 2] = __float_as_uint(__uint_as_float(ov[i + 2]) * wa + o4.z * wb);
                ov[i + 3] = __float_as_uint(__uint_as_float(ov[i + 3]) * wa + o4.w * wb);
              }
            }
          }
        }
        if (store && tile_valid && q0 + t * FA_BM + q < pr.nq) {
          const float inv = 1.f / l;
          __nv_bfloat16* dst = p.O + (size_t)(pr.q_row0 + q0 + t * FA_BM + q) * p.ldo + h * FA_D + hf * 32;
  #pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 pk4;
            __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
            __nv_bfloat162 c2 = __floats2bfloat162_rn(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
            __nv_bfloat162 d = __floats2bfloat162_rn(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
            pk4.x = *reinterpret_cast<uint32_t*>(&a); pk4.y = *reinterpret_cast<uint32_t*>(&b2);
            pk4.z = *reinterpret_cast<uint32_t*>(&c2); pk4.w = *reinterpret_cast<uint32_t*>(&d);
            *reinterpret_cast<uint4*>(dst + i) = pk4;
          }
        }
      FA_SEGMENT_ADVANCE
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, FA_TMEM_COLS);
}

extern "C" __attribute__((visibility("default"))) size_t i4d_attention_workspace_bytes(void) {
  return (size_t)2 * i4d_num_sms() * FA_PART_FLOATS * sizeof(float) + FA_MAX_ITEMS * sizeof(int) + 512;
}

extern "C" __attribute__((visibility("default"))) int i4d_attention_bf16_tc(
    const void* X, int rows, int ld, int q_col, int k_col, int v_col, int heads, const int* problems_host, int n_problems,
    const int* key_counts_dev, float scale, void* O, int ldo, void* workspace, size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(X && O && problems_host, "null pointer");
  I4D_CHECK_ARG(n_problems >= 1 && n_problems <= FA_MAX_PROBLEMS && heads >= 1, "1..4 problems, heads >= 1");
  I4D_CHECK_ARG((ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "O must be 16-byte aligned with ldo % 8 == 0");
  I4D_CHECK_ARG(q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0, "column offsets must be multiples of 8");
  AttnParams p;
  long long units = 0;
  int items = 0;
  for (int z = 0; z < FA_MAX_PROBLEMS; ++z) {
    p.unit_off[z] = (int)units; p.item_off[z] = items;
    if (z < n_problems) {
      p.prob[z] = AttnProblem{problems_host[4 * z], problems_host[4 * z + 1], problems_host[4 * z + 2], problems_host[4 * z + 3]};
      I4D_CHECK_ARG(p.prob[z].nq >= 0 && p.prob[z].nk >= 1, "every problem needs nk >= 1");
      I4D_CHECK_ARG(p.prob[z].q_row0 >= 0 && p.prob[z].k_row0 >= 0 && p.prob[z].q_row0 + p.prob[z].nq <= rows &&
                    p.prob[z].k_row0 + p.prob[z].nk <= rows, "row ranges outside the buffer");
      p.n_qt[z] = i4d_cdiv(p.prob[z].nq, 2 * FA_BM); p.nblk[z] = i4d_cdiv(p.prob[z].nk, FA_BN);
      units += (long long)p.n_qt[z] * heads * p.nblk[z];
      items += p.n_qt[z] * heads;
    } else {
      p.prob[z] = AttnProblem{0, 0, 0, 1}; p.n_qt[z] = 0; p.nblk[z] = 1;
    }
  }
  p.unit_off[FA_MAX_PROBLEMS] = (int)units; p.item_off[FA_MAX_PROBLEMS] = items;
  if (items == 0) return I4D_OK;
  I4D_CHECK_ARG(units < (1ll << 30), "problem too large");
  p.heads = heads;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.O = reinterpret_cast<__nv_bfloat16*>(O); p.ldo = ldo;
  p.nk_dev = key_counts_dev;
  CUtensorMap tmX;
  if (int rc = i4d_make_tmap_2d_bf16(&tmX, X, (uint64_t)rows, (uint64_t)ld, (uint64_t)ld, FA_BN, FA_D)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // one persistent CTA per SM.  With the caller's workspace the key blocks of all items are cut into equal contiguous ranges
  // (stream-K: no ragged last wave); without it every CTA runs whole items.
  const int sms = i4d_num_sms();
  p.balanced = (workspace && workspace_bytes >= i4d_attention_workspace_bytes() && items <= FA_MAX_ITEMS) ? 1 : 0;
  p.part = nullptr; p.counters = nullptr;
  if (p.balanced) {
    long long g = units / FA_MIN_SEG;
    p.n_ctas = (int)(g < 1 ? 1 : (g > sms ? sms : g));
    uint8_t* w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    p.counters = reinterpret_cast<int*>(w);
    p.part = reinterpret_cast<float*>(w + ((FA_MAX_ITEMS * sizeof(int) + 255) & ~(size_t)255));
    I4D_CUDA_CALL(cudaMemsetAsync(p.counters, 0, (size_t)items * sizeof(int), st));
  } else {
    p.n_ctas = items < sms ? items : sms;
  }
  // share of exponentials on the FMA pipe (I4D_FA_POLY = 12 / 25 / 37 / 50 / 75 / 100 percent, -1 = none at all; experiments)
  using KernelFn = void (*)(const CUtensorMap, AttnParams);
  static const struct { int pct; KernelFn fn; } table[] = {
      {0, attn_tc_kernel<0x00>},  {12, attn_tc_kernel<0x80>}, {25, attn_tc_kernel<0x88>}, {37, attn_tc_kernel<0x8A>},
      {50, attn_tc_kernel<0xAA>}, {75, attn_tc_kernel<0xEE>}, {100, attn_tc_kernel<0xFF>}, {-1, attn_tc_kernel<0x100>}};
  static int variant = -1;
  static bool attr_seen[64] = {};
  if (variant < 0) {
    const char* e = getenv("I4D_FA_POLY");
    const int v = e ? atoi(e) : FA_DEFAULT_POLY;
    variant = 0;
    for (int i = 0; i < (int)(sizeof(table) / sizeof(table[0])); ++i)
      if (table[i].pct == v) variant = i;
  }
  if (i4d_first_use_on_device(attr_seen)) {
    for (const auto& t : table)
      I4D_CUDA_CALL(cudaFuncSetAttribute(t.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
  }
  table[variant].fn<<<p.n_ctas, FA_THREADS, FA_SMEM_BYTES, st>>>(tmX, p);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
