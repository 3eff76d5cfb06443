// tcgen05 / TMA GEMM on sm_100a:  C[M,N] = alpha * A[M,K] * W[N,K]^T + bias[N]  (+ReLU) (+R[M,N] f32),
// bf16 operands (both K-major), f32 accumulation in TMEM, outputs f32 and/or bf16 from one epilogue.
//
// One 128 x 128 output tile per CTA.  warp 0 / lane 0: TMA producer (A and W tiles of 128 x 64 bf16 = one 128-byte
// swizzle row per matrix row) through a 3-stage mbarrier ring (3 stages); warp 1 / lane 0: issues tcgen05.mma (M128 N128 K16, four
// per stage) and commits each stage back to the producer; all four warps then read the accumulator with tcgen05.ld
// (warp w owns TMEM lanes 32w..32w+31 = output rows) and run the fused epilogue.  Two CTAs fit per SM (2 x 97 KB
// smem, 2 x 128 TMEM columns), so one CTA's epilogue overlaps the other's main loop.
//
// Reference behaviour replaced: the Conv1d(k=1)/Linear layers of thirdparty/SuperGlue/models/superglue.py:51-61,
// 100-128, 276-280 and thirdparty/LightGlue/lightglue/lightglue.py:133-216, 253-287 (cuBLAS sgemm via torch there).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/icepy4d_b200.h"

#define GT_BM 128
#define GT_BN 128
#define GT_BK 64
#define GT_STAGES 3
#define GT_STAGE_BYTES ((GT_BM + GT_BN) * GT_BK * 2)
#define GT_SMEM_BYTES (GT_STAGES * GT_STAGE_BYTES + 1024)

struct GemmTcParams {
  int M, N, K;
  float alpha;
  const float* bias;
  const float* R; int ldr;
  float* C32; int ldc32;
  __nv_bfloat16* C16; int ldc16;
  int relu;
};

__global__ void __launch_bounds__(128) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmW, GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[GT_STAGES], empty_bar[GT_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * GT_BN;
  const int kblocks = p.K / GT_BK;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
    for (int s = 0; s < GT_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&accum_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, GT_BN);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % GT_STAGES;
        const uint32_t ph = (kb / GT_STAGES) & 1;
        tc::mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * GT_STAGE_BYTES;
        uint8_t* sb = sa + GT_BM * GT_BK * 2;
        tc::mbar_arrive_expect_tx(&full_bar[s], GT_STAGE_BYTES);
        tc::tma_load_2d(sa, &tmA, &full_bar[s], kb * GT_BK, m0);
        tc::tma_load_2d(sb, &tmW, &full_bar[s], kb * GT_BK, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc(GT_BM, GT_BN, 0, 0, 1);
      constexpr uint32_t hi = tc::desc_hi_sw128(1024);
      const uint32_t d0 = tc::desc_lo_sw128(tc::smem_u32(smem));
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % GT_STAGES;
        const uint32_t ph = (kb / GT_STAGES) & 1;
        tc::mbar_wait(&full_bar[s], ph);
        tc::tcgen05_fence_after();
        const uint32_t da = d0 + (uint32_t)(s * (GT_STAGE_BYTES >> 4)), db = da + ((GT_BM * GT_BK * 2) >> 4);
#pragma unroll
        for (int k = 0; k < GT_BK / 16; ++k) tc::umma_f16_parts(tmem_d, da + k * 2, hi, db + k * 2, hi, idesc, (kb | k) ? 1u : 0u);
        tc::umma_commit(&empty_bar[s]);       // smem stage reusable once these MMAs retire
      }
      tc::umma_commit(&accum_bar);            // accumulator complete
    }
    __syncwarp();
  }

  // ---- epilogue: all 4 warps ----
  tc::mbar_wait(&accum_bar, 0);
  tc::tcgen05_fence_after();
  const int row = m0 + warp * 32 + lane;
  const bool row_ok = row < p.M;
#pragma unroll 1
  for (int c = 0; c < GT_BN / 32; ++c) {
    uint32_t v[32];
    tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tc::tmem_ld_wait();
    const int nb = n0 + c * 32;
    if (row_ok && nb < p.N) {
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = p.alpha * __uint_as_float(v[j]);
        if (p.bias) x += __ldg(p.bias + min(nb + j, p.N - 1));
        if (p.relu) x = fmaxf(x, 0.f);
        f[j] = x;
      }
      const bool full = nb + 32 <= p.N;
      if (p.R) {
        const float* r = p.R + (size_t)row * p.ldr + nb;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 t = *reinterpret_cast<const float4*>(r + j);
            f[j] += t.x; f[j + 1] += t.y; f[j + 2] += t.z; f[j + 3] += t.w;
          }
        } else {
          for (int j = 0; j < 32 && nb + j < p.N; ++j) f[j] += r[j];
        }
      }
      if (p.C32) {
        float* o = p.C32 + (size_t)row * p.ldc32 + nb;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
          for (int j = 0; j < 32 && nb + j < p.N; ++j) o[j] = f[j];
        }
      }
      if (p.C16) {
        __nv_bfloat16* o = p.C16 + (size_t)row * p.ldc16 + nb;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            __nv_bfloat162 a = __floats2bfloat162_rn(f[j], f[j + 1]), b = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
            __nv_bfloat162 c2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), d = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
            pk.z = *reinterpret_cast<uint32_t*>(&c2); pk.w = *reinterpret_cast<uint32_t*>(&d);
            *reinterpret_cast<uint4*>(o + j) = pk;
          }
        } else {
          for (int j = 0; j < 32 && nb + j < p.N; ++j) o[j] = __float2bfloat16_rn(f[j]);
        }
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_d, GT_BN);
}

// ---- host: tensor-map encoding through the driver entry point (no link-time libcuda dependency) ----------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int i4d_make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                          uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { i4d_set_error("cuTensorMapEncodeTiled entry point not available"); return I4D_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15)) {
    i4d_set_error("TMA needs a 16-byte aligned base and row pitch (base %p, ld %llu)", base, (unsigned long long)ld);
    return I4D_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { i4d_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return I4D_ERR_CUDA; }
  return I4D_OK;
}

int i4d_make_tmap_hwc_bf16(CUtensorMap* out, const void* base, uint64_t H, uint64_t W, uint64_t C, uint32_t box_h, uint32_t box_w) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { i4d_set_error("cuTensorMapEncodeTiled entry point not available"); return I4D_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (C % 64) || box_h > 256 || box_w > 256) {
    i4d_set_error("TMA image map: base %p must be 16-byte aligned, C %llu a multiple of 64, box <= 256", base, (unsigned long long)C);
    return I4D_ERR_INVALID;
  }
  cuuint64_t dims[3] = {C, W, H};
  cuuint64_t strides[2] = {C * 2, W * C * 2};
  cuuint32_t box[3] = {64, box_w, box_h};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { i4d_set_error("cuTensorMapEncodeTiled (3D) failed (%d)", (int)r); return I4D_ERR_CUDA; }
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_gemm_bf16_tc(
    const void* A, int lda, const void* W, int ldw, const float* bias, const float* R, int ldr, float* C32, int ldc32,
    void* C16, int ldc16, int M, int N, int K, float alpha, int relu, void* stream) {
  I4D_CHECK_ARG(A && W && (C32 || C16), "null pointer");
  I4D_CHECK_ARG(M > 0 && N > 0 && K > 0, "bad sizes");
  I4D_CHECK_ARG(K % GT_BK == 0, "K must be a multiple of 64 for the tensor-core GEMM");
  I4D_CHECK_ARG(lda >= K && ldw >= K, "leading dimensions too small");
  I4D_CHECK_ARG(!C32 || ((ldc32 & 3) == 0 && (reinterpret_cast<uintptr_t>(C32) & 15) == 0), "C32 must be 16-byte aligned, ldc32 % 4 == 0");
  I4D_CHECK_ARG(!C16 || ((ldc16 & 7) == 0 && (reinterpret_cast<uintptr_t>(C16) & 15) == 0), "C16 must be 16-byte aligned, ldc16 % 8 == 0");
  I4D_CHECK_ARG(!R || ((ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0), "R must be 16-byte aligned, ldr % 4 == 0");
  CUtensorMap tmA, tmW;
  if (int rc = i4d_make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GT_BM, GT_BK)) return rc;
  if (int rc = i4d_make_tmap_2d_bf16(&tmW, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, GT_BN, GT_BK)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES));
    attr_set = true;
  }
  GemmTcParams p{M, N, K, alpha, bias, R, ldr, C32, ldc32, reinterpret_cast<__nv_bfloat16*>(C16), ldc16, relu};
  dim3 grid(i4d_cdiv(N, GT_BN), i4d_cdiv(M, GT_BM));
  gemm_tc_kernel<<<grid, 128, GT_SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmW, p);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---- f32 -> bf16 row-major conversion with leading dimensions (feeds the tensor-core path) ---------------------------
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ Y,
                                                          int ldy, int rows, int cols4) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols4) return;
  int r = i / cols4, c = (i - r * cols4) * 4;
  float4 v = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + c);
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(Y + (size_t)r * ldy + c) = pk;
}

extern "C" __attribute__((visibility("default"))) int i4d_f32_to_bf16(const float* X, int ldx, void* Y, int ldy, int rows,
                                                                     int cols, void* stream) {
  I4D_CHECK_ARG(X && Y && rows >= 0 && cols > 0, "bad arguments");
  I4D_CHECK_ARG((cols & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0, "cols and leading dimensions must be multiples of 4");
  if (rows == 0) return I4D_OK;
  long long n = (long long)rows * (cols / 4);
  f32_to_bf16_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, reinterpret_cast<__nv_bfloat16*>(Y), ldy, rows,
                                                                         cols / 4);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
