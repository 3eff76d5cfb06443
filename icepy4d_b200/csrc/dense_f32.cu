// Exact-f32 dense kernels (CUDA cores): tiled GEMM with fused bias / ReLU / residual epilogue and a flash-style
// multi-head attention that never materialises the N x M probability matrix.  These are the "precision = f32"
// path (bit-for-bit f32 semantics of the reference's Conv1d/Linear/einsum/softmax stack, used for parity gates and
// for the small MLPs whose K is 3..128); the tensor-core path lives in gemm_tc.cu / attn_tc.cu.
//
// Reference behaviour replaced (paths into /root/reference/src/icepy4d/thirdparty):
//   SuperGlue/models/superglue.py:51-61 (MLP, Conv1d k=1 + BatchNorm folded on the host) :87-116 (attention)
//   LightGlue/lightglue/lightglue.py:108-130 (Attention), :133-216 (Linear layers), :49-57 (rotary), ffn LayerNorm+GELU
#include "common.cuh"
#include "../../include/icepy4d_b200.h"

// ------------------------------------------------------------------------------------------------------------
// C[M,N] = alpha * A[M,K] * W[N,K]^T + bias[N]  (+ReLU) (+ R[M,N]);  row-major with leading dimensions.
// 64x64 tile, BK = 16, 256 threads, 4x4 micro-tile, register-staged double buffering.
// ------------------------------------------------------------------------------------------------------------
#define GB 64
#define GK 16

__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                       int ldw, const float* __restrict__ bias,
                                                       const float* __restrict__ R, int ldr, float* __restrict__ C,
                                                       int ldc, int M, int N, int K, float alpha, int relu) {
  __shared__ float As[GK][GB + 4];
  __shared__ float Ws[GK][GB + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // each thread loads 4 elements of A and 4 of W per k-tile: row = tid / 4, k = (tid % 4) * 4 .. +3
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += GK) {
    float a[4], w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int kk = k0 + lk + e;
      a[e] = (m0 + lr < M && kk < K) ? __ldg(A + (size_t)(m0 + lr) * lda + kk) : 0.f;
      w[e] = (n0 + lr < N && kk < K) ? __ldg(W + (size_t)(n0 + lr) * ldw + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) { As[lk + e][lr] = a[e]; Ws[lk + e][lr] = w[e]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      if (R) v += R[(size_t)m * ldr + n];
      C[(size_t)m * ldc + n] = v;
    }
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias, const float* R,
                            int ldr, float* C, int ldc, int M, int N, int K, float alpha, int relu, void* stream) {
  I4D_CHECK_ARG(A && W && C, "null pointer");
  I4D_CHECK_ARG(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, "bad sizes");
  if (M == 0) return I4D_OK;
  dim3 grid(i4d_cdiv(N, GB), i4d_cdiv(M, GB));
  gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, W, ldw, bias, R, ldr, C, ldc, M, N, K, alpha, relu);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Flash-style attention, head_dim 64, f32.  Q [Nq, H*64] (ldq), K/V [Nk, H*64] (ldk/ldv), O [Nq, H*64] (ldo);
// head h owns columns [64h, 64h+64).  softmax(scale * Q K^T) V with online (running max / sum) normalisation.
// grid = (ceil(Nq/64), H); 256 threads; thread (ty,tx) owns S[4ty..4ty+3][4tx..4tx+3] and O[4ty..][4tx..].
// ------------------------------------------------------------------------------------------------------------
#define AT 64
#define AD 64

__global__ void __launch_bounds__(256) attn_f32_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ Kp,
                                                       int ldk, const float* __restrict__ Vp, int ldv,
                                                       float* __restrict__ O, int ldo, int Nq, int Nk, float scale) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                   // [AD][AT+4]  (transposed: Qs[d][q])
  float* Ks = Qs + AD * (AT + 4);   // [AD][AT+4]  (transposed: Ks[d][k])
  float* Vs = Ks + AD * (AT + 4);   // [AT][AD+4]  (Vs[k][d])
  float* Ps = Vs + AT * (AD + 4);   // [AT][AT+4]  (transposed: Ps[k][q])
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * AT, hoff = blockIdx.y * AD;
  for (int i = tid; i < AT * AD; i += 256) {
    int q = i >> 6, d = i & 63;
    Qs[d * (AT + 4) + q] = (q0 + q < Nq) ? __ldg(Q + (size_t)(q0 + q) * ldq + hoff + d) * scale : 0.f;
  }
  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < Nk; k0 += AT) {
    __syncthreads();
    for (int i = tid; i < AT * AD; i += 256) {
      int k = i >> 6, d = i & 63;
      bool ok = k0 + k < Nk;
      Ks[d * (AT + 4) + k] = ok ? __ldg(Kp + (size_t)(k0 + k) * ldk + hoff + d) : 0.f;
      Vs[k * (AD + 4) + d] = ok ? __ldg(Vp + (size_t)(k0 + k) * ldv + hoff + d) : 0.f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < AD; ++d) {
      float4 qv = *reinterpret_cast<const float4*>(&Qs[d * (AT + 4) + ty * 4]);
      float4 kv = *reinterpret_cast<const float4*>(&Ks[d * (AT + 4) + tx * 4]);
      float qr[4] = {qv.x, qv.y, qv.z, qv.w}, kr[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qr[i], kr[j], s[i][j]);
    }
    // online softmax per query row (a row is spread over the 16 lanes sharing ty)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k0 + tx * 4 + j >= Nk) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float m_new = fmaxf(m_run[i], mx);
      float corr = expf(m_run[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p = expf(s[i][j] - m_new);
        rs += p;
        Ps[(tx * 4 + j) * (AT + 4) + ty * 4 + i] = p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[i] = l_run[i] * corr + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < AT; ++k) {
      float4 pv = *reinterpret_cast<const float4*>(&Ps[k * (AT + 4) + ty * 4]);
      float4 vv = *reinterpret_cast<const float4*>(&Vs[k * (AD + 4) + tx * 4]);
      float pr[4] = {pv.x, pv.y, pv.z, pv.w}, vr[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pr[i], vr[j], o[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int q = q0 + ty * 4 + i;
    if (q >= Nq) continue;
    float inv = 1.f / l_run[i];
    float4 r = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    *reinterpret_cast<float4*>(O + (size_t)q * ldo + hoff + tx * 4) = r;
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_attention_f32(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* O,
                                 int ldo, int Nq, int Nk, int heads, float scale, void* stream) {
  I4D_CHECK_ARG(Q && K && V && O, "null pointer");
  I4D_CHECK_ARG(Nq >= 0 && Nk > 0 && heads > 0, "bad sizes");
  I4D_CHECK_ARG((ldo & 3) == 0, "ldo must be a multiple of 4 floats");
  if (Nq == 0) return I4D_OK;
  size_t smem = (size_t)(2 * AD * (AT + 4) + AT * (AD + 4) + AT * (AT + 4)) * sizeof(float);
  static bool attr_seen[64] = {};
  if (i4d_first_use_on_device(attr_seen)) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid(i4d_cdiv(Nq, AT), heads);
  attn_f32_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(Q, ldq, K, ldk, V, ldv, O, ldo, Nq, Nk, scale);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Row-wise LayerNorm(C) + exact (erf) GELU, in place capable.  One warp per row.  (LightGlue ffn, lightglue.py:144-149)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_gelu_kernel(const float* __restrict__ X, int ldx,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ Y,
                                                             int ldy, int rows, int C, float eps) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = X + (size_t)row * ldx;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += x[c];
  float mean = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = x[c] - mean; v += d * d; }
  float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
  float* y = Y + (size_t)row * ldy;
  for (int c = lane; c < C; c += 32) {
    float t = (x[c] - mean) * rstd * gamma[c] + beta[c];
    y[c] = 0.5f * t * (1.f + erff(t * 0.70710678118654752440f));
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_layernorm_gelu(const float* X, int ldx, const float* gamma, const float* beta, float* Y, int ldy,
                                  int rows, int C, float eps, void* stream) {
  I4D_CHECK_ARG(X && Y && gamma && beta && C > 0 && rows >= 0, "bad arguments");
  if (rows == 0) return I4D_OK;
  layernorm_gelu_kernel<<<i4d_cdiv((long long)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, gamma, beta, Y, ldy,
                                                                                               rows, C, eps);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ------------------------------------------------------------------------------------------------------------
// LightGlue learnable Fourier positional encoding + rotary application (lightglue.py:49-74).
//   enc: cs[n][0..31] = cos(Wr * kn), cs[n][32..63] = sin(Wr * kn)   (one value per rotary PAIR)
//   rotary: for every head h and pair p: (x1, x2) -> (x1 c - x2 s, x2 c + x1 s) on columns 64h + 2p, 64h + 2p + 1
// ------------------------------------------------------------------------------------------------------------
__global__ void lg_posenc_kernel(const float* __restrict__ kpts, int n, float sx, float sy, float inv_scale,
                                 const float* __restrict__ Wr, float* __restrict__ cs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 32) return;
  int p = i & 31, t = i >> 5;
  float x = (kpts[2 * t] - sx) * inv_scale, y = (kpts[2 * t + 1] - sy) * inv_scale;
  float pr = x * Wr[2 * p] + y * Wr[2 * p + 1];
  cs[(size_t)t * 64 + p] = cosf(pr);
  cs[(size_t)t * 64 + 32 + p] = sinf(pr);
}
__global__ void lg_rotary_kernel(float* __restrict__ X, int ldx, int n, int heads, const float* __restrict__ cs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * heads * 32) return;
  int p = i & 31, h = (i >> 5) % heads, t = i / (32 * heads);
  float c = cs[(size_t)t * 64 + p], s = cs[(size_t)t * 64 + 32 + p];
  float2* q = reinterpret_cast<float2*>(X + (size_t)t * ldx + h * 64 + 2 * p);
  float2 v = *q;
  *q = make_float2(v.x * c - v.y * s, v.y * c + v.x * s);
}

extern "C" __attribute__((visibility("default"))) int i4d_lg_posenc(const float* kpts, int n, float width, float height, const float* Wr, float* cs,
                             void* stream) {
  I4D_CHECK_ARG(kpts && Wr && cs && n >= 0, "bad arguments");
  if (n == 0) return I4D_OK;
  float sc = fmaxf(width, height) / 2.f;
  lg_posenc_kernel<<<i4d_cdiv((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(kpts, n, width / 2.f, height / 2.f,
                                                                                        1.f / sc, Wr, cs);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
extern "C" __attribute__((visibility("default"))) int i4d_lg_rotary(float* X, int ldx, int n, int heads, const float* cs, void* stream) {
  I4D_CHECK_ARG(X && cs && n >= 0 && heads > 0 && (ldx & 1) == 0, "bad arguments");
  if (n == 0) return I4D_OK;
  lg_rotary_kernel<<<i4d_cdiv((long long)n * heads * 32, 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, n, heads, cs);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ------------------------------------------------------------------------------------------------------------
// SuperGlue keypoint-encoder input: [ (x - W/2) / (0.7 max(W,H)), (y - H/2) / (0.7 max(W,H)), score ]  (superglue.py:64-71,82-84)
// ------------------------------------------------------------------------------------------------------------
__global__ void sg_kenc_input_kernel(const float* __restrict__ kpts, const float* __restrict__ sc, int n, float cx,
                                     float cy, float scaling, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[3 * i + 0] = (kpts[2 * i] - cx) / scaling;
  out[3 * i + 1] = (kpts[2 * i + 1] - cy) / scaling;
  out[3 * i + 2] = sc[i];
}
extern "C" __attribute__((visibility("default"))) int i4d_sg_kenc_input(const float* kpts, const float* scores, int n, float width, float height, float* out,
                                 void* stream) {
  I4D_CHECK_ARG(kpts && scores && out && n >= 0, "bad arguments");
  if (n == 0) return I4D_OK;
  sg_kenc_input_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(kpts, scores, n, width / 2.f, height / 2.f,
                                                                           fmaxf(width, height) * 0.7f, out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
