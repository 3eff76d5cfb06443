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
// CTA = 128 query rows x 1 head, 320 threads, two CTAs per SM:
//   warp 0      TMA producer: Q once, then K_j (2-stage ring) and V_j (single buffer) blocks of 128 keys
//   warp 1      MMA issuer (one lane): S_j = Q K_j^T (M128 N128 K64 -> 4 tcgen05.mma) into TMEM[0,128);
//               O += P_j V_j (M128 N64 K128 -> 8 tcgen05.mma, V is the MN-major B operand) accumulating in TMEM[128,192)
//   warps 2..9  softmax: TWO threads per query row (TMEM lane), 64 keys each.  S is read from TMEM exactly once
//               (TMEM reads, 64 B/clk/SM, and MUFU.EX2, 16/clk/SM, are the two co-limiting units at head_dim 64),
//               released immediately so that QK_{j+1} overlaps the exponentials of block j.  Row max exchanged
//               between the two threads of a row through shared memory; p = exp2(s*c - m) packed to bf16 and stored
//               in the 128B-swizzled K-major layout the MMA reads.  The running max is LAZY: O (in TMEM) and the row
//               sum are only rescaled when the max grows by more than 2^8, so the TMEM read-modify-write of O is rare.
//
// Reference behaviour replaced: `attention()` + MultiHeadedAttention of thirdparty/SuperGlue/models/superglue.py:87-116
// (materialises a 4 x N x M f32 tensor) and Attention/SelfBlock/CrossBlock of thirdparty/LightGlue/lightglue/lightglue.py:92-216.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/icepy4d_b200.h"

#define FA_BM 128
#define FA_BN 128
#define FA_D 64
#define FA_KV_STAGES 2
#define FA_THREADS 320
#define FA_Q_BYTES (FA_BM * FA_D * 2)           // 16 KB
#define FA_KV_BYTES (FA_BN * FA_D * 2)          // 16 KB each for K and V
#define FA_P_BYTES (FA_BM * FA_BN * 2)          // 32 KB (two 16 KB K-halves)
#define FA_SMEM_BYTES (FA_Q_BYTES + FA_KV_STAGES * FA_KV_BYTES + FA_KV_BYTES + FA_P_BYTES + 1024)   // Q | K x2 | V | P = 96 KB (+ align)
#define FA_TMEM_COLS 256                        // S: [0,128)  O: [128,192)
#define FA_MAX_PROBLEMS 4
#define FA_TAU 8.0f                             // lazy-rescale threshold (log2 units)

struct AttnProblem { int q_row0, nq, k_row0, nk; };
struct AttnParams {
  AttnProblem prob[FA_MAX_PROBLEMS];
  int q_col, k_col, v_col;
  float scale_log2;                 // scale * log2(e)
  __nv_bfloat16* O; int ldo;        // O rows are indexed like Q rows (q_row0 + i)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__global__ void __launch_bounds__(FA_THREADS, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap tmX, AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FA_Q_BYTES;                                   // K ring: stage s at sK + s*16K (double buffered)
  uint8_t* sV = sK + FA_KV_STAGES * FA_KV_BYTES;                   // V: single buffer (needed only from P_j to PV_j)
  uint8_t* sP = sV + FA_KV_BYTES;
  __shared__ __align__(8) uint64_t q_full, k_full[FA_KV_STAGES], k_empty[FA_KV_STAGES], v_full, v_empty, s_full, s_empty, p_full, pv_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float xch[2][2][FA_BM];                               // [block parity][column half][row]: row-max exchange

  const AttnProblem pr = p.prob[blockIdx.z];
  const int q0 = blockIdx.x * FA_BM;
  if (q0 >= pr.nq) return;                                        // uniform per CTA: safe before any barrier
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int nblk = (pr.nk + FA_BN - 1) / FA_BN;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX);
    tc::mbar_init(&q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) { tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1); }
    tc::mbar_init(&v_full, 1);
    tc::mbar_init(&v_empty, 1);
    tc::mbar_init(&s_full, 1);
    tc::mbar_init(&s_empty, 256);
    tc::mbar_init(&p_full, 256);
    tc::mbar_init(&pv_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, FA_TMEM_COLS);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_S = tmem, tmem_O = tmem + 128;

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
      load_k(0);
      if (nblk > 1) load_k(1);
      for (int j = 0; j < nblk; ++j) {
        tc::mbar_wait(&v_empty, (j & 1) ^ 1);                                // PV_{j-1} retired
        tc::mbar_arrive_expect_tx(&v_full, FA_KV_BYTES);
        tc::tma_load_2d(sV, &tmX, &v_full, p.v_col + h * FA_D, pr.k_row0 + j * FA_BN);
        if (j + 2 < nblk) load_k(j + 2);                                     // stage freed when QK_j retired
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (one elected lane: operands stay in uniform registers)
    if (tc::elect_one()) {
      constexpr uint32_t idesc_qk = tc::make_idesc(FA_BM, FA_BN, 0, 0, 1);   // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = tc::make_idesc(FA_BM, FA_D, 0, 1, 1);    // A = P (K-major), B = V (MN-major)
      constexpr uint32_t hi_k = tc::desc_hi_sw128(1024);                     // K-major operands and MN-major V: SBO = 1024
      const uint32_t dQ = tc::desc_lo_sw128(tc::smem_u32(sQ)), dP = tc::desc_lo_sw128(tc::smem_u32(sP));
      const uint32_t dK0 = tc::desc_lo_sw128(tc::smem_u32(sK));
      auto issue_qk = [&](int j) {
        const int s = j % FA_KV_STAGES;
        tc::mbar_wait(&k_full[s], (j / FA_KV_STAGES) & 1);
        if (j > 0) tc::mbar_wait(&s_empty, (j - 1) & 1);                    // S_{j-1} has been read into registers
        tc::tcgen05_fence_after();
        const uint32_t dK = dK0 + (uint32_t)(s * (FA_KV_BYTES >> 4));
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) tc::umma_f16_parts(tmem_S, dQ + k * 2, hi_k, dK + k * 2, hi_k, idesc_qk, k ? 1u : 0u);
        tc::umma_commit(&k_empty[s]);                                       // K_j no longer needed once these retire
        tc::umma_commit(&s_full);
      };
      tc::mbar_wait(&q_full, 0);
      issue_qk(0);
      // V descriptor: MN-major, 8-key groups 1024 B apart (SBO), one 64-wide N atom: LBO field = 1024 >> 4 as well
      const uint32_t dV = ((tc::smem_u32(sV) >> 4) & 0x3FFF) | ((1024u >> 4) << 16);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_qk(j + 1);                                  // runs while the softmax warps exponentiate block j
        tc::mbar_wait(&v_full, j & 1);
        tc::mbar_wait(&p_full, j & 1);                                      // P_j in shared memory, O rescaled if needed
        tc::tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          // A: P k-slice = 16 keys = 32 B inside the 128-B swizzle row of K-half (k / 4);  B: V rows [16k, 16k+16) x 64 dims
          tc::umma_f16_parts(tmem_O, dP + (uint32_t)(((k >> 2) * (FA_BM * 128) + (k & 3) * 32) >> 4), hi_k,
                             dV + (uint32_t)((k * 16 * 128) >> 4), hi_k, idesc_pv, (j | k) ? 1u : 0u);
        }
        tc::umma_commit(&v_empty);                                          // V buffer free
        tc::umma_commit(&pv_done);                                          // O includes block j; P buffer free
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ softmax: two threads per query row
    const int quarter = warp & 3;                                           // TMEM lane quarter this warp may access
    const int hf = (warp - 2) >> 2;                                         // column half: keys [64 hf, 64 hf + 64) of the block
    const int q = quarter * 32 + lane;                                      // TMEM lane == tile row
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t rsw = (uint32_t)(q & 7);
    const uint32_t prow = tc::smem_u32(sP) + (uint32_t)hf * (FA_BM * 128) + (uint32_t)(q >> 3) * 1024 + (uint32_t)(q & 7) * 128;
    float m_run = -INFINITY, l_part = 0.f;

    for (int j = 0; j < nblk; ++j) {
      tc::mbar_wait(&s_full, j & 1);
      tc::tcgen05_fence_after();
      uint32_t v0[32], v1[32];
      tc::tmem_ld32(tmem_S + lane_off + hf * 64, v0);
      tc::tmem_ld32(tmem_S + lane_off + hf * 64 + 32, v1);
      tc::tmem_ld_wait();
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&s_empty);                                            // QK_{j+1} may overwrite S now
      const int kvalid = min(64, max(0, pr.nk - j * FA_BN - hf * 64));      // keys of my half that exist
      if (kvalid < 64) {                                                    // only in the last block
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kvalid) v0[i] = 0xff800000u;                             // -inf
          if (32 + i >= kvalid) v1[i] = 0xff800000u;
        }
      }
      // 3-input max (FMNMX3): 32 instructions for the 64 values, four independent chains
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; ++i) mxa[i & 3] = fmax3(mxa[i & 3], __uint_as_float(v0[i]), __uint_as_float(v1[i]));
      float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
      xch[j & 1][hf][q] = mx;
      named_bar_sync(1 + quarter, 64);                                      // the two warps that share this lane quarter
      mx = fmaxf(mx, xch[j & 1][hf ^ 1][q]);
      const float m_blk = mx * p.scale_log2;
      // lazy running max: rescale only when this block's max exceeds the reference max by more than 2^TAU
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_run = m_blk;
      } else if (m_blk > m_run + FA_TAU) {
        alpha = ex2_approx(m_run - m_blk);
        m_run = m_blk;
        need = true;
      }
      if (j > 0) {
        tc::mbar_wait(&pv_done, (j - 1) & 1);                               // P buffer free, O quiescent
        if (__any_sync(0xffffffffu, need)) {
          tc::tcgen05_fence_after();
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {                                  // two 16-column halves: keeps the register peak low
            uint32_t ov[16];
            tmem_ld16(tmem_O + lane_off + hf * 32 + hh * 16, ov);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st16(tmem_O + lane_off + hf * 32 + hh * 16, ov);
            tmem_st_wait();
          }
          l_part *= alpha;
        }
      }
      // p = exp2(s * c - m_run); f32 row sum; bf16 pack; swizzled store (8 chunks of 16 B = my K-half row)
      // packed f32x2 arithmetic (FFMA2 / FADD2): one issue slot per two elements for the scaling and the row sum
      float2 rs2 = make_float2(0.f, 0.f);
      const float2 nm2 = make_float2(-m_run, -m_run), sc2 = make_float2(p.scale_log2, p.scale_log2);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = t * 8 + e * 2;
          const uint32_t ra = (i < 32) ? v0[i & 31] : v1[i & 31];
          const uint32_t rb = (i + 1 < 32) ? v0[(i + 1) & 31] : v1[(i + 1) & 31];
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(ra), __uint_as_float(rb)), sc2, nm2);
          const float2 ab = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          rs2 = __fadd2_rn(rs2, ab);
          __nv_bfloat162 pr2 = __floats2bfloat162_rn(ab.x, ab.y);
          pk[e] = *reinterpret_cast<uint32_t*>(&pr2);
        }
        const uint32_t addr = prow + ((((uint32_t)t) ^ rsw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      }
      l_part += rs2.x + rs2.y;
      tc::tcgen05_fence_before();
      tc::fence_proxy_async_smem();                                         // make P_j visible to the tensor-core proxy
      tc::mbar_arrive(&p_full);
    }
    // epilogue: O / l -> bf16.  Each thread owns 32 of the 64 output dims of its row.
    tc::mbar_wait(&pv_done, (nblk - 1) & 1);
    tc::tcgen05_fence_after();
    xch[nblk & 1][hf][q] = l_part;
    named_bar_sync(1 + quarter, 64);
    const float l = l_part + xch[nblk & 1][hf ^ 1][q];
    uint32_t ov[32];
    tc::tmem_ld32(tmem_O + lane_off + hf * 32, ov);
    tc::tmem_ld_wait();
    if (q0 + q < pr.nq) {
      const float inv = 1.f / l;
      __nv_bfloat16* dst = p.O + (size_t)(pr.q_row0 + q0 + q) * p.ldo + h * FA_D + hf * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 pk;
        __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
        __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
        __nv_bfloat162 c2 = __floats2bfloat162_rn(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
        __nv_bfloat162 d = __floats2bfloat162_rn(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
        pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
        pk.z = *reinterpret_cast<uint32_t*>(&c2); pk.w = *reinterpret_cast<uint32_t*>(&d);
        *reinterpret_cast<uint4*>(dst + i) = pk;
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, FA_TMEM_COLS);
}

extern "C" __attribute__((visibility("default"))) int i4d_attention_bf16_tc(
    const void* X, int rows, int ld, int q_col, int k_col, int v_col, int heads, const int* problems_host, int n_problems,
    float scale, void* O, int ldo, void* stream) {
  I4D_CHECK_ARG(X && O && problems_host, "null pointer");
  I4D_CHECK_ARG(n_problems >= 1 && n_problems <= FA_MAX_PROBLEMS && heads >= 1, "1..4 problems, heads >= 1");
  I4D_CHECK_ARG((ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "O must be 16-byte aligned with ldo % 8 == 0");
  I4D_CHECK_ARG(q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0, "column offsets must be multiples of 8");
  AttnParams p;
  int max_nq = 0;
  for (int z = 0; z < FA_MAX_PROBLEMS; ++z) {
    if (z < n_problems) {
      p.prob[z] = AttnProblem{problems_host[4 * z], problems_host[4 * z + 1], problems_host[4 * z + 2], problems_host[4 * z + 3]};
      I4D_CHECK_ARG(p.prob[z].nq >= 0 && p.prob[z].nk >= 1, "every problem needs nk >= 1");
      I4D_CHECK_ARG(p.prob[z].q_row0 >= 0 && p.prob[z].k_row0 >= 0 && p.prob[z].q_row0 + p.prob[z].nq <= rows &&
                    p.prob[z].k_row0 + p.prob[z].nk <= rows, "row ranges outside the buffer");
      if (p.prob[z].nq > max_nq) max_nq = p.prob[z].nq;
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
  static bool attr_set = false;
  if (!attr_set) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(i4d_cdiv(max_nq, FA_BM), heads, n_problems);
  attn_tc_kernel<<<grid, FA_THREADS, FA_SMEM_BYTES, (cudaStream_t)stream>>>(tmX, p);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
