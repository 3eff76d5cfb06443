/* icepy4d_b200 — C ABI of the B200-native (sm_100a) hot path of ICEpy4D's per-epoch stereo
 * match -> verify -> triangulate pipeline.
 *
 * The reference (franioli/icepy4d) is pure Python: there is no FFI for this path today.  Each
 * entry point below replaces a block of torch / OpenCV / numpy calls at the cited reference location
 * (paths relative to /root/reference/src/icepy4d).  INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless its name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises unless stated;
 *   - no allocation inside: scratch comes in through `workspace` (+ its size query);
 *   - return value: 0 = ok, <0 = error (I4D_ERR_*), message via i4d_last_error(); never throws.
 *   - matrices are row-major with explicit leading dimensions (in elements) where given.
 */
#ifndef ICEPY4D_B200_H
#define ICEPY4D_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I4D_VERSION 100 /* 0.1.0 */

/* ---- library ------------------------------------------------------------------------------------------ */
int i4d_version(void);
const char* i4d_last_error(void);
/* SM count of the current device (0 if no device) — grids are sized from it. */
int i4d_device_sm_count(void);

/* ---- SuperPoint first layer ---------------------------------------------------------------------------- */
/* superpoint.py:154 — relu(conv1a(image)): 1 -> 64 channels, 3x3, pad 1.  image [H,W] f32, weight [64,9] (= [64,1,3,3]),
 * bias [64]; out [H,W,64] channels-last, out_dtype 0 = f32, 1 = f16, 2 = bf16, 3 = split bf16: out is [2][H,W,64], plane 0
 * = bf16(v), plane 1 = bf16(v - plane 0), the operand format of i4d_conv_bf16x3_tc; out_dtype 4: the same with IEEE halves. */
int i4d_sp_conv1a_relu(const float* image, int H, int W, const float* weight, const float* bias, void* out_nhwc,
                       int out_dtype, void* stream);

/* superpoint.py:156,159,162 — MaxPool2d(2,2) on a channels-last activation [H,W,C] -> [H/2,W/2,C].  elem_bytes 4 (f32) or
 * 2 (f16/bf16, requires nonneg = 1: post-ReLU data, compared as bit patterns). */
int i4d_maxpool2x2_nhwc(const void* in, int H, int W, int C, void* out, int elem_bytes, int nonneg, void* stream);

/* superpoint.py:154-168 (conv1b .. conv4b, convPa, convDa: 3x3, pad 1) and :193-196 (convPb, convDb: 1x1) as tcgen05
 * implicit GEMMs on SPLIT bf16 operands: every f32 value v travels as hi = bf16(v), lo = bf16(v - hi) and the kernel
 * accumulates x_hi*w_hi + x_hi*w_lo + x_lo*w_hi in f32 (relative error ~2^-16; a single TF32/bf16 product is not enough
 * for the >= 99 % match-IoU bar).
 *   x_hi, x_lo : bf16 [H][W][Cin] channels-last planes, Cin % 64 == 0
 *   w_packed   : bf16 [(nt*taps + tap)*Cin/64 + kc][2*n_t][64] with n_t = i4d_conv_tile_cout(cout_pad), taps = ksize^2
 *                (tap = ky*3 + kx); rows 0..n_t-1 of a block hold bf16(w[nt*n_t + r, kc*64 + k, ky, kx]), rows
 *                n_t..2n_t-1 the bf16 remainder; output channels >= cout are zero
 *   bias       : f32 [cout_pad]
 *   outputs    : y_hi / y_lo bf16 planes [Ho][Wo][cout_pad] (requires cout == cout_pad; with pool = 1 the 2x2/stride-2
 *                max-pool of superpoint.py:156,159,162 is fused: Ho = H/2, Wo = W/2) and/or y32 f32, either channels-last
 *                [H][W][ld32] or planar [cout][H][W] (y32_planar = 1, the layout i4d_sp_score_map reads)
 *   operand_format : 1 = bfloat16 pairs as described ("bf16x3", 16 mantissa bits per value); 0 = IEEE-half pairs ("f16x3",
 *                22 mantissa bits, values clamped to +-65504): same three products, same speed, f32-grade results */
int i4d_conv_tile_cout(int cout_pad);
/* superpoint.py:154-156 — relu(conv1a) -> relu(conv1b) (-> MaxPool2d(2,2) when pool = 1) in ONE kernel: conv1a's 64-channel
 * activation (1 GB per 2000 x 2000 tile as split planes) never goes to HBM; warps of the conv1b kernel compute it from the grey
 * image into the shared-memory operand buffers (bit-identical values to i4d_sp_conv1a_relu).  image [H][W] f32, w1a [64][9],
 * b1a [64], w1b_packed / b1b as for i4d_conv_bf16x3_tc (Cin = 64, cout = 64), outputs y_hi / y_lo [Ho][Wo][64] split planes. */
int i4d_sp_conv1ab_tc(const float* image, int H, int W, const float* w1a, const float* b1a, const void* w1b_packed,
                      const float* b1b, int pool, void* y_hi, void* y_lo, int operand_format, void* stream);
int i4d_conv_bf16x3_tc(const void* x_hi, const void* x_lo, int H, int W, int Cin, const void* w_packed, const float* bias,
                       int cout_pad, int cout, int ksize, int relu, int pool, void* y_hi, void* y_lo, float* y32, int ld32,
                       int y32_planar, int operand_format, void* stream);

/* ---- SuperPoint post-processing ----------------------------------------------------------------------- */
/* thirdparty/SuperGlue/models/superpoint.py:169-172 — softmax over 65 channels, drop dustbin, 8x8 pixel shuffle.
 * logits [65,h,w] f32 -> scores [8h,8w] f32. */
int i4d_sp_score_map(const float* logits, int h, int w, float* scores, void* stream);
/* superpoint.py:48-64 (simple_nms, 3 passes, radius <= 4) + :176-190 threshold `> thr` and border removal
 * (LightGlue/lightglue/superpoint.py:176-186 gives the same set).  Candidates are appended (unordered) as 64-bit
 * keys (score bits << 32 | ~linear index) to cand_keys[cand_cap]; *cand_count (device) receives the total found
 * (may exceed cand_cap: caller checks).  nms_out (nullable) receives the dense NMS'd score map [H,W]. */
int i4d_sp_nms_candidates(const float* scores, int H, int W, int nms_radius, float thr, int border,
                          unsigned long long* cand_keys, int cand_cap, int* cand_count, float* nms_out,
                          void* stream);
/* superpoint.py:75-79,193-203 — keep the k best (k < 0: keep all), emit (x, y) as f32 and scores.  Order: score
 * descending (ties: lower linear index first) when k applies, row-major otherwise, like topk / nonzero.
 * kpts [out_cap,2], scores [out_cap], *n_out (device) = number written.  spill: scratch of 2*cand_cap + 4096 keys
 * (cand_cap >= 16384).
 * Synchronises the stream only when k < 0 or k > 16384. */
int i4d_sp_select_topk(const unsigned long long* cand_keys, const int* cand_count, int cand_cap, int k, int W,
                       float* kpts, float* scores, int out_cap, int* n_out, unsigned long long* spill, void* stream);
/* superpoint.py:82-97,208 — dense L2 normalisation, bilinear grid_sample(align_corners=True), L2 normalisation.
 * desc_hwc [h,w,256] f32 (channels-last), kpts [n,2]; n = min(*n_dev, n_max) if n_dev != NULL else n_max.
 * out [n,256] (token-major; the reference's [256,n] is its transpose). */
int i4d_sp_sample_descriptors(const float* desc_hwc, int h, int w, const float* kpts, const int* n_dev, int n_max,
                              float* out, void* stream);

/* ---- dense f32 building blocks ------------------------------------------------------------------------ */
/* C = alpha * A[M,K] * W[N,K]^T + bias[N] (+ReLU) (+R[M,N]).  Replaces Conv1d(k=1)/Linear (+ folded BatchNorm,
 * + residual): superglue.py:51-61,119-128,147-148,276-280; lightglue.py:133-216,253-287. bias, R nullable. */
int i4d_gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias, const float* R, int ldr,
                 float* C, int ldc, int M, int N, int K, float alpha, int relu, void* stream);
/* softmax(scale * Q K^T) V per head, head_dim 64, head h = columns [64h, 64h+64) — superglue.py:87-93 (with the
 * head permutation folded into the weights), lightglue.py:108-130. Never materialises the Nq x Nk matrix. */
int i4d_attention_f32(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* O, int ldo,
                      int Nq, int Nk, int heads, float scale, void* stream);
/* LayerNorm(C) + GELU(erf) row-wise — lightglue.py:144-149 (ffn[1], ffn[2]). */
int i4d_layernorm_gelu(const float* X, int ldx, const float* gamma, const float* beta, float* Y, int ldy, int rows,
                       int C, float eps, void* stream);
/* lightglue.py:23-35,60-74 — normalise keypoints by the image size and evaluate the learnable Fourier encoding.
 * Wr [32,2]; cs [n,64] = (cos[32] | sin[32]) per keypoint. */
int i4d_lg_posenc(const float* kpts, int n, float width, float height, const float* Wr, float* cs, void* stream);
/* lightglue.py:49-57 — apply the cached rotary embedding in place to `heads` heads of 64 columns. */
int i4d_lg_rotary(float* X, int ldx, int n, int heads, const float* cs, void* stream);
/* superglue.py:64-71,82-84 — keypoint-encoder input rows (x_norm, y_norm, score), out [n,3]. */
int i4d_sg_kenc_input(const float* kpts, const float* scores, int n, float width, float height, float* out,
                      void* stream);

/* ---- dense tensor-core building blocks (tcgen05 + TMEM + TMA, bf16 operands, f32 accumulation) ---------- */
/* Same contraction as i4d_gemm_f32 with A [M,K] and W [N,K] in bf16 (K % 64 == 0, 16-byte aligned bases/pitches).
 * Outputs: C32 (f32, nullable) and/or C16 (bf16, nullable) written by one fused epilogue
 * (alpha, bias, ReLU, f32 residual R). */
int i4d_gemm_bf16_tc(const void* A, int lda, const void* W, int ldw, const float* bias, const float* R, int ldr,
                     float* C32, int ldc32, void* C16, int ldc16, int M, int N, int K, float alpha, int relu,
                     void* stream);
/* The same GEMM with LightGlue's rotary embedding (lightglue.py:49-57, 155-159) applied in the epilogue: output columns
 * < rot_cols (a multiple of 128; heads of 64 columns, pair p = (column % 64) / 2) are rotated by the row's angles,
 * cs [M,64] = cos[32] | sin[32] from i4d_lg_posenc; the remaining columns (v) pass through.  bf16 output only. */
int i4d_gemm_bf16_tc_rotary(const void* A, int lda, const void* W, int ldw, const float* bias, const float* cs, int rot_cols,
                            void* C16, int ldc16, int M, int N, int K, void* stream);
/* LightGlue glue of the tensor-core path (lightglue.py:49-57,144-159): fused element-wise passes between the tcgen05 kernels.
 *   i4d_lg_rotary_cast_bf16: qkv [n,768] f32 (heads contiguous: q | k | v, 4 x 64 each), cs [n,64] (cos[32] | sin[32] per rotary
 *     pair, from i4d_lg_posenc) -> out [n,768] bf16 with the rotary embedding applied to q and k.
 *   i4d_layernorm_gelu_bf16: X [n,C] f32 -> GELU(LayerNorm(X) * gamma + beta) as bf16 (C = 512). */
int i4d_lg_rotary_cast_bf16(const float* qkv, int ldx, const float* cs, int n, void* out, int ldy, void* stream);
int i4d_layernorm_gelu_bf16(const float* X, int ldx, const float* gamma, const float* beta, int n, int C, float eps, void* out,
                            int ldy, void* stream);
/* Flash attention, head_dim 64, on one bf16 buffer X [rows, ld] holding Q, K and V as column blocks
 * (q_col/k_col/v_col + 64*head).  problems_host: n_problems x {q_row0, nq, k_row0, nk} (host ints, 1..4 problems run in
 * one launch: both images of a self/cross layer).  O [rows, ldo] bf16, row-indexed like Q.  `workspace` (nullable,
 * i4d_attention_workspace_bytes()) lets the persistent kernel cut the key blocks of all (problem, head, 256-query tile) items into
 * equal per-SM ranges and merge the parts of items a range boundary cuts; without it every CTA runs whole items.
 * key_counts_dev (nullable, DEVICE ints, one per problem): only the first key_counts_dev[z] of the problem's nk keys are real, the
 * rest is masked — the launch geometry stays that of a shape bucket (a captured CUDA graph serves every keypoint count of the
 * bucket), the true counts are read on the device.
 * Replaces superglue.py:87-93 and lightglue.py:108-130 on the throughput path. */
size_t i4d_attention_workspace_bytes(void);
int i4d_attention_bf16_tc(const void* X, int rows, int ld, int q_col, int k_col, int v_col, int heads,
                          const int* problems_host, int n_problems, const int* key_counts_dev, float scale, void* O,
                          int ldo, void* workspace, size_t workspace_bytes, void* stream);
/* row-major f32 -> bf16 with leading dimensions (cols % 4 == 0). */
int i4d_f32_to_bf16(const float* X, int ldx, void* Y, int ldy, int rows, int cols, void* stream);
/* row-major f32 [rows, cols] -> bf16 [rows, 3 cols]: hi = bf16(x), lo = bf16(x - hi) laid out as [hi | hi | lo] (order 0,
 * activations) or [hi | lo | hi] (order 1, weights), so that ONE i4d_gemm_bf16_tc of depth 3 cols computes
 * A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T in f32: the split-precision product for layers that must keep ~f32 accuracy on the tensor
 * cores (SuperGlue's KeypointEncoder MLP, thirdparty/SuperGlue/models/superglue.py:51-61,67-78).  cols % 4 == 0. */
int i4d_f32_split3_bf16(const float* X, int ldx, void* Y, int ldy, int rows, int cols, int order, void* stream);

/* ---- assignment --------------------------------------------------------------------------------------- */
size_t i4d_assignment_workspace_bytes(int M, int N);
/* out_i = LSE_j(scale * S_ij + coloff_j) / out_j = LSE_i(scale * S_ij + rowoff_i); offsets nullable. */
int i4d_row_lse(const float* S, int M, int N, float scale, const float* coloff, float* out, void* stream);
int i4d_col_lse(const float* S, int M, int N, float scale, const float* rowoff, float* out, void* workspace,
                size_t workspace_bytes, void* stream);
/* superglue.py:152-186 — log-domain Sinkhorn potentials u[M+1], v[N+1] of the dustbin-augmented problem.
 * scores [M, N] row-major with a row pitch of `ld` floats (ld >= N).  Buffers: u holds M + 1 floats, v must have room for
 * N + 4 floats (entries past v[N] are scratch).  When ld is a multiple of 4 that covers N rounded up to 4 (and `scores` is
 * 16-byte aligned) any 64 <= N <= 8192 takes the fused kernel; for N % 4 != 0 it OVERWRITES the pad columns
 * scores[:, N .. round_up(N, 4)) with -1e30.  Other shapes run the two-pass row/column kernels. */
int i4d_sinkhorn(float* scores, int M, int N, int ld, float bin_score, int iters, float* u, float* v, void* workspace,
                 size_t workspace_bytes, void* stream);
/* Sinkhorn implementation switch (tests / comparisons): 0 = fused persistent kernel (one HBM read of the score matrix
 * per iteration; see i4d_sinkhorn for the shapes it takes), 1 = always the two-pass row/column kernels, 2 = fused kernel
 * with running maxima in every iteration. */
int i4d_set_sinkhorn_mode(int mode);
/* superglue.py:152-186 + :288-298 — Sinkhorn, then mutual nearest neighbours with threshold.
 * matches0 [M] / matches1 [N] int32 (-1 = unmatched), mscores0/1 f32; u [M+1], v [N+4] scratch/outputs; `ld` and the pad
 * columns as for i4d_sinkhorn. */
int i4d_sg_assign(float* scores, int M, int N, int ld, float bin_score, int iters, float match_threshold,
                  int* matches0, int* matches1, float* mscores0, float* mscores1, float* u, float* v,
                  void* workspace, size_t workspace_bytes, void* stream);
/* lightglue.py:253-266 + :290-306 — sigmoid/log double softmax + mutual NN with threshold, from sim [M,N] and
 * matchability logits z0 [M], z1 [N]; sim has a row pitch of `ld` floats (128-bit loads when ld % 4 == 0).  The
 * (M+1)x(N+1) log-assignment matrix is never written. */
int i4d_lg_assign(const float* sim, int M, int N, int ld, const float* z0, const float* z1, float filter_threshold,
                  int* matches0, int* matches1, float* mscores0, float* mscores1, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ---- two-view geometry -------------------------------------------------------------------------------- */
/* sfm/geometry.py:103-118 — cv2.undistortPoints(pts, K, dist, None, K) cast to f32.  K_host [9] row-major,
 * dist_host [n_dist <= 5] = k1,k2,p1,p2[,k3] (host pointers, copied by value). */
int i4d_undistort_points(const float* pts, int n, const double* K_host, const double* dist_host, int n_dist,
                         float* out, void* stream);
/* thirdparty/triangulation.py:79-177 — iterative re-weighted LS triangulation (f64).  P*_host [12] row-major.
 * X [n,3] f64, status [n] int32 with the reference's codes (1, 0, -1, -2, -3). */
int i4d_triangulate_iterative_ls(const float* u1, const float* u2, int n, const double* P1_host,
                                 const double* P2_host, double tolerance, double* X, int* status, void* stream);
/* sfm/triangulation.py:154-183 — DLT via the null vector of the 6x6 system, dehomogenised. X [n,3] f64. */
int i4d_triangulate_dlt(const float* x1, const float* x2, int n, const double* P1_host, const double* P2_host,
                        double* X, void* stream);

/* sfm/interpolate_colors.py:14-87 + sfm/geometry.py:78-100 — colours of 3-D points from an oriented image: Brown-model
 * projection (cv2.projectPoints arithmetic, f64, rounded to f32) and bilinear interpolation of image / 255 with the
 * reference's clipped-neighbour weights.  X [n,3] f64 and image [H,W,C] u8 on the device; R [9], t [3], K [9], dist on the
 * host.  colors [n,C] f64; projections [n,2] f32 (nullable). */
int i4d_interpolate_point_colors(const double* X, int n, const double* R_host, const double* t_host, const double* K_host,
                                 const double* dist_host, int n_dist, const unsigned char* image, int H, int W, int C,
                                 int convert_bgr2rgb, double* colors, float* projections, void* stream);

/* sfm/absolute_orientation.py:141-154 (Absolute_orientation.estimate_transformation_linear -> thirdparty/transformations.py:
 * 889-1020 affine_matrix_from_points(shear=False, scale=True, usesvd=False)) — the reductions of the closed-form Helmert /
 * similarity fit.  v0, v1 [n,3] f64 on the device.  out [17] f64 (device): mean(v0) [3], mean(v1) [3], the cross-covariance of
 * the centred sets sum c0 c1^T [9, row-major], sum |c0|^2, sum |c1|^2.  The 4x4 quaternion eigenproblem is solved by the host. */
int i4d_helmert_moments(const double* v0, const double* v1, int n, double* out, void* stream);
/* sfm/absolute_orientation.py:269-272 (apply_transformation): out = dehomogenise(T [x; 1]), T [16] row-major 4x4 on the host,
 * X, out [n,3] f64 on the device (out may alias X). */
int i4d_apply_transform(const double* X, int n, const double* T_host, double* out, void* stream);

/* matching/geometric_verification.py:43-102 and sfm/two_view_geometry.py:127-197 — robust fundamental matrix.
 * Batched-hypothesis RANSAC (8-point samples drawn from `seed`) scored with the MAGSAC++ marginalised quality function
 * (sigma-consensus++, 4 degrees of freedom, k = 3.64, noise scale marginalised up to `sigma_max` on the Sampson error — what
 * cv2.findFundamentalMat(USAC_MAGSAC) scores with), then a polisher run to its fixed point from each of the 4 best hypotheses
 * (at most `polish_iters` weighted normalised 8-point solves each, batched in the same launches; the best final quality wins and
 * replaces the RANSAC winner only if it is not worse):
 *   polish_mode 0 = MAGSAC++ weights  w(r^2) = Gamma(3/2, r^2 / (2 sigma_max^2)) - Gamma(3/2, k^2 / 2)   (the MAGSAC branch,
 *                   geometric_verification.py:89-92; sigma_max = 4.5 / 3.64 px reproduces OpenCV 4.13, scripts/magsac_probe.py)
 *   polish_mode 1 = least squares on the inliers at `threshold`, LO-RANSAC's final step (the pydegensac branch, :66-76).
 * Inliers = sqrt(Sampson error) < threshold (OpenCV USAC's and pydegensac's rule).  x0, x1 [n,2] f32 raw pixel coordinates.
 * Outputs (device): F_out [9] f64 row-major (scaled so F[8] = 1 when possible), mask [n] u8, *n_inliers.  If no valid model
 * exists (every minimal sample degenerate: coincident / collinear points) F_out = NaN, mask = all zeros, *n_inliers = 0 — what
 * cv2.findFundamentalMat returns in that case (F None, zero mask), which the reference passes on (geometric_verification.py:89-92). */
size_t i4d_fundamental_workspace_bytes(void);
int i4d_fundamental_ransac(const float* x0, const float* x1, int n, double threshold, double confidence,
                           int max_iters, unsigned int seed, double sigma_max, int polish_iters, int polish_mode,
                           double* F_out, unsigned char* mask, int* n_inliers, void* workspace, size_t workspace_bytes,
                           void* stream);

/* sfm/geometry.py:31-76 (cv2.findEssentialMat inlier rule + cv2.recoverPose) — relative pose from an essential-matrix
 * estimate.  E_in [9] f64 (device, row-major, any scale), xn0 / xn1 [n,2] f32 K-normalised coordinates.  Projects E onto the
 * essential manifold, marks Sampson inliers (error < threshold_norm^2), votes the four (R, t) candidates by cheirality over
 * the inliers (depth in both cameras in (0, distance_threshold)) and keeps the best in OpenCV's candidate order.
 * Outputs (device): E_out [9], R_out [9] row-major, t_out [3] (unit), mask [n] u8 = inlier AND in front of both cameras,
 * n_good [2] = {votes of the chosen candidate, Sampson inliers}. */
/* sfm/geometry.py:63-65 (cv2.findEssentialMat(..., method=cv2.RANSAC) on K-normalised coordinates) — five-point RANSAC.
 * xn0 / xn1 [n,2] f32 K-normalised coordinates (n >= 5), threshold_norm = pixel threshold / focal length.  Seeded batches of
 * 512 minimal samples (Nister's five-point solver, up to 10 solutions each), every solution scored on all correspondences by
 * its Sampson-inlier count, OpenCV's adaptive stopping rule, at most max_iters samples (rounded up to whole batches).
 * Outputs (device): E_out [9] row-major, unit Frobenius norm (NaN when no sample produced a model), n_inliers [1]. */
size_t i4d_essential_workspace_bytes(void);
int i4d_essential_ransac(const float* xn0, const float* xn1, int n, double threshold_norm, double confidence, int max_iters,
                         unsigned int seed, double* E_out, int* n_inliers, void* workspace, size_t workspace_bytes,
                         void* stream);
size_t i4d_pose_workspace_bytes(int n);
int i4d_essential_pose(const double* E_in, const float* xn0, const float* xn1, int n, double threshold_norm,
                       double distance_threshold, double* E_out, double* R_out, double* t_out, unsigned char* mask,
                       int* n_good, void* workspace, size_t workspace_bytes, void* stream);

/* ---- tiler front end ---------------------------------------------------------------------------------- */
/* matching/tiling.py:123-135 (extract_patch) fused with the grey conversion and the /255 tensor conversion:
 *   mode 0 (SuperGlueMatcher, matchers.py:911-917,263-274): cv2.cvtColor(RGB2GRAY) fixed point on u8, then /255.
 *   mode 1 (LightGlueMatcher, matchers.py:1212-1220 + LightGlue/lightglue/utils.py:35-36): /255. per channel, then
 *          0.299 r + 0.587 g + 0.114 b in f32.
 * image [H,W,C] u8 (C = 1 or 3) on the device; tile = rows [y0, y0+th), cols [x0, x0+tw); out [th,tw] f32. */
int i4d_tile_to_gray_f32(const unsigned char* image, int H, int W, int C, int x0, int y0, int tw, int th, int mode,
                         float* out, void* stream);

/* matching/matchers.py:495-499,541-556 — the PRESELECTION decision: for every tile pair (t0, t1) the number of pre-matches whose
 * up-scaled keypoints lie strictly inside both tile rectangles.  kp0 / kp1 [n,2] f32 (row i = one candidate match, kp1 already
 * gathered through matches0), valid [n] u8 or NULL, scale = 2^n_down, lims0 [T0,4] / lims1 [T1,4] f32 (xmin, ymin, xmax, ymax;
 * T <= 64) on the device.  counts [T0*T1] i32 (device) is zeroed and filled; the caller keeps pairs with count > min_matches. */
int i4d_tile_pair_counts(const float* kp0, const float* kp1, const unsigned char* valid, int n, float scale, const float* lims0,
                         int T0, const float* lims1, int T1, int* counts, void* stream);

/* matching/matchers.py:583-610 (Quality resize) and :526-531 (PRESELECTION pass) — cv2.pyrDown / cv2.pyrUp on 8-bit images,
 * bit-exact with OpenCV's integer arithmetic and border rules.  image [H,W,C] u8 (C = 1 or 3).
 * pyr_down: out [(H+1)/2, (W+1)/2, C];  pyr_up: out [2H, 2W, C]. */
int i4d_pyr_down_u8(const unsigned char* in, int H, int W, int C, unsigned char* out, void* stream);
int i4d_pyr_up_u8(const unsigned char* in, int H, int W, int C, unsigned char* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
