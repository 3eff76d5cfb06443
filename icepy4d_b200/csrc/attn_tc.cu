// Flash-style multi-head attention on sm_100a tensor cores (tcgen05 + TMEM + TMA), head_dim 64, bf16 operands,
// f32 softmax statistics and f32 output accumulation.  The Nq x Nk probability matrix never leaves the SM.
//
//   O[q, h*64 : h*64+64] = softmax_k( scale * Q_h[q] . K_h[k] ) V_h[k]
//
// Q, K, V live in ONE row-major bf16 buffer X [rows, ld] (e.g. the fused QKV projection output): a "problem" z is given by
// row ranges (q_row0, Nq), (k_row0, Nk) and column offsets (q_col, k_col, v_col); head h adds 64*h columns.  One TMA tensor map
// over X serves all three operands.  gridDim = (ceil(maxNq/128), heads, n_problems) so both images of a SuperGlue /
// LightGlue layer (self or cross) run in one launch.
//
// CTA = 128 query rows x 1 head, 192 threads:
//   warp 0      TMA producer: Q once, then K_j / V_j blocks of 128 keys through a 2-stage mbarrier ring
//   warp 1      MMA issuer (one lane): S = Q K_j^T (M128 N128 K64 -> 4 tcgen05.mma) into TMEM[0,128);
//               PV_j = P_j V_j (M128 N64 K128 -> 8 tcgen05.mma, V is the MN-major B operand) into TMEM[128,192)
//   warps 2..5  softmax: thread = one query row (TMEM lane).  Two passes over S in TMEM (row max, then exp2 + bf16 pack),
//               P_j written to shared memory in the 128B-swizzled K-major layout the MMA reads, running (max, sum) in
//               registers, O accumulated in f32 registers: O = O * alpha_j + PV_j, deferred by one block so that the
//               PV MMA of block j and the QK MMA of block j+1 overlap the softmax of block j+1.
// Two CTAs fit per SM (2 x ~98 KB smem, 2 x 256 TMEM columns), so MMA / MUFU / TMEM traffic of two tiles interleave.
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
#define FA_Q_BYTES (FA_BM * FA_D * 2)           // 16 KB
#define FA_KV_BYTES (FA_BN * FA_D * 2)          // 16 KB each for K and V
#define FA_P_BYTES (FA_BM * FA_BN * 2)          // 32 KB (two 16 KB K-halves)
#define FA_SMEM_BYTES (FA_Q_BYTES + FA_KV_STAGES * 2 * FA_KV_BYTES + FA_P_BYTES + 1024)
#define FA_TMEM_COLS 256                        // S: [0,128)  PV: [128,192)
#define FA_MAX_PROBLEMS 4

struct AttnProblem { int q_row0, nq, k_row0, nk; };
struct AttnParams {
  AttnProblem prob[FA_MAX_PROBLEMS];
  int q_col, k_col, v_col;
  float scale_log2;                 // scale * log2(e)
  __nv_bfloat16* O; int ldo;        // O rows are indexed like Q rows (q_row0 + i)
};

__global__ void __launch_bounds__(192, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap tmX, AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + FA_Q_BYTES;                                  // stage s: K at sKV + s*32K, V at +16K
  uint8_t* sP = sKV + FA_KV_STAGES * 2 * FA_KV_BYTES;
  __shared__ __align__(8) uint64_t q_full, kv_full[FA_KV_STAGES], kv_empty[FA_KV_STAGES], s_full, s_empty, p_full, pv_full, pv_empty;
  __shared__ uint32_t tmem_base_s;

  const AttnProblem pr = p.prob[blockIdx.z];
  const int q0 = blockIdx.x * FA_BM;
  if (q0 >= pr.nq) return;                                        // uniform per CTA: safe before any barrier
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int nblk = (pr.nk + FA_BN - 1) / FA_BN;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmX);
    tc::mbar_init(&q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
    tc::mbar_init(&s_full, 1);
    tc::mbar_init(&s_empty, 128);
    tc::mbar_init(&p_full, 128);
    tc::mbar_init(&pv_full, 1);
    tc::mbar_init(&pv_empty, 128);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, FA_TMEM_COLS);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_S = tmem, tmem_PV = tmem + 128;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&q_full, FA_Q_BYTES);
      tc::tma_load_2d(sQ, &tmX, &q_full, p.q_col + h * FA_D, pr.q_row0 + q0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % FA_KV_STAGES;
        const uint32_t ph = (j / FA_KV_STAGES) & 1;
        tc::mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sK = sKV + s * 2 * FA_KV_BYTES;
        tc::mbar_arrive_expect_tx(&kv_full[s], 2 * FA_KV_BYTES);
        tc::tma_load_2d(sK, &tmX, &kv_full[s], p.k_col + h * FA_D, pr.k_row0 + j * FA_BN);
        tc::tma_load_2d(sK + FA_KV_BYTES, &tmX, &kv_full[s], p.v_col + h * FA_D, pr.k_row0 + j * FA_BN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = tc::make_idesc(FA_BM, FA_BN, 0, 0, 1);   // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = tc::make_idesc(FA_BM, FA_D, 0, 1, 1);    // A = P (K-major), B = V (MN-major)
      const uint32_t aQ = tc::smem_u32(sQ), aP = tc::smem_u32(sP);
      auto issue_qk = [&](int j) {
        const int s = j % FA_KV_STAGES;
        tc::mbar_wait(&kv_full[s], (j / FA_KV_STAGES) & 1);
        if (j > 0) tc::mbar_wait(&s_empty, (j - 1) & 1);                    // softmax finished reading S_{j-1}
        tc::tcgen05_fence_after();
        const uint32_t aK = tc::smem_u32(sKV + s * 2 * FA_KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k)
          tc::umma_f16(tmem_S, tc::make_smem_desc_sw128(aQ + k * 32, 16, 1024), tc::make_smem_desc_sw128(aK + k * 32, 16, 1024),
                       idesc_qk, k ? 1u : 0u);
        tc::umma_commit(&s_full);
      };
      tc::mbar_wait(&q_full, 0);
      issue_qk(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % FA_KV_STAGES;
        tc::mbar_wait(&p_full, j & 1);                                      // P_j is in shared memory (and S_j consumed)
        if (j + 1 < nblk) issue_qk(j + 1);                                  // overlaps softmax of block j+1 with PV_j
        if (j > 0) tc::mbar_wait(&pv_empty, (j - 1) & 1);                   // PV_{j-1} has been read out of TMEM
        tc::tcgen05_fence_after();
        const uint32_t aV = tc::smem_u32(sKV + s * 2 * FA_KV_BYTES + FA_KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          // A: P k-slice = 16 keys = 32 B inside the 128-B swizzle row of K-half (k / 4)
          const uint64_t da = tc::make_smem_desc_sw128(aP + (k >> 2) * (FA_BM * 128) + (k & 3) * 32, 16, 1024);
          // B: V rows [16k, 16k+16) x 64 dims, MN-major: 8-key groups are 1024 B apart (SBO), one 64-wide N atom (LBO unused)
          const uint64_t db = tc::make_smem_desc_sw128(aV + k * 16 * 128, 1024, 1024);
          tc::umma_f16(tmem_PV, da, db, idesc_pv, k ? 1u : 0u);
        }
        tc::umma_commit(&kv_empty[s]);                                      // K_j / V_j stage free
        tc::umma_commit(&pv_full);                                          // PV_j ready (also: P buffer free)
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ softmax / accumulate: thread = one query row
    const int q = (warp & 3) * 32 + lane;                                   // TMEM lane == tile row
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    float o[FA_D];
#pragma unroll
    for (int i = 0; i < FA_D; ++i) o[i] = 0.f;
    const uint32_t sP_u = tc::smem_u32(sP);
    const uint32_t rsw = (uint32_t)(q & 7);
    const uint32_t prow = sP_u + (uint32_t)(q >> 3) * 1024 + (uint32_t)(q & 7) * 128;

    for (int j = 0; j < nblk; ++j) {
      const int kvalid = min(FA_BN, pr.nk - j * FA_BN);                     // keys of this block that exist
      tc::mbar_wait(&s_full, j & 1);
      tc::tcgen05_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < FA_BN / 32; ++c) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_S + lane_off + c * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(v[i]);
          if (c * 32 + i < kvalid) mx = fmaxf(mx, x);
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = exp2f(m_run - m_new);                             // 0 on the first block (m_run = -inf)
      // the previous P buffer must have been consumed by PV_{j-1} before we overwrite it
      if (j > 0) tc::mbar_wait(&pv_full, (j - 1) & 1);
      // pass 2: p = exp2(s * scale_log2 - m_new), bf16 pack, swizzled store
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < FA_BN / 32; ++c) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_S + lane_off + c * 32, v);
        tc::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a = (c * 32 + i < kvalid) ? exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new)) : 0.f;
          float b = (c * 32 + i + 1 < kvalid) ? exp2f(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_new)) : 0.f;
          __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
          // accumulate the row sum from the ROUNDED probabilities, so that sum(P) matches what the MMA multiplies
          float2 tf = __bfloat1622float2(t);
          rs += tf.x + tf.y;
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
        }
        // 32 keys = 64 B = four 16-B chunks: global chunk index cc = c*4 + t -> K-half cc / 8, chunk (cc % 8) ^ (row % 8)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t cc = (uint32_t)(c * 4 + t);
          const uint32_t addr = prow + (cc >> 3) * (FA_BM * 128) + (((cc & 7) ^ rsw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * t]), "r"(pk[4 * t + 1]), "r"(pk[4 * t + 2]),
                       "r"(pk[4 * t + 3])
                       : "memory");
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&s_empty);                                            // S_j fully read: QK_{j+1} may overwrite it
      tc::fence_proxy_async_smem();                                         // make P_j visible to the tensor-core proxy
      tc::mbar_arrive(&p_full);
      // deferred accumulation of block j-1 (its PV has completed: we waited on pv_full above)
      if (j > 0) {
        tc::tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < FA_D / 32; ++c) {
          uint32_t v[32];
          tc::tmem_ld32(tmem_PV + lane_off + c * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
        }
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&pv_empty);
      }
      alpha_prev = alpha;
    }
    // last block
    tc::mbar_wait(&pv_full, (nblk - 1) & 1);
    tc::tcgen05_fence_after();
#pragma unroll
    for (int c = 0; c < FA_D / 32; ++c) {
      uint32_t v[32];
      tc::tmem_ld32(tmem_PV + lane_off + c * 32, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
    }
    if (q0 + q < pr.nq) {
      const float inv = 1.f / l_run;
      __nv_bfloat16* dst = p.O + (size_t)(pr.q_row0 + q0 + q) * p.ldo + h * FA_D;
#pragma unroll
      for (int i = 0; i < FA_D; i += 8) {
        uint4 pk;
        __nv_bfloat162 a = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv), b = __floats2bfloat162_rn(o[i + 2] * inv, o[i + 3] * inv);
        __nv_bfloat162 c2 = __floats2bfloat162_rn(o[i + 4] * inv, o[i + 5] * inv), d = __floats2bfloat162_rn(o[i + 6] * inv, o[i + 7] * inv);
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
  attn_tc_kernel<<<grid, 192, FA_SMEM_BYTES, (cudaStream_t)stream>>>(tmX, p);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
