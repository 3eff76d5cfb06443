// tcgen05 / TMA GEMM on sm_100a:  C[M,N] = alpha * A[M,K] * W[N,K]^T + bias[N]  (+ReLU) (+R[M,N] f32),
// bf16 operands (both K-major), f32 accumulation in TMEM, outputs f32 and/or bf16 from one epilogue.
//
// Persistent, warp-specialised: one CTA per SM walks 128 x 128 output tiles (n fastest, so neighbouring CTAs share A rows in L2).
//   warp 0      TMA producer: A and W tiles of 128 x 64 bf16 (one 128-byte swizzle row per matrix row) through a 3- to 5-stage
//               mbarrier ring (template parameter); it runs ahead across tile boundaries
//   warp 1      MMA issuer (one elected lane): tcgen05.mma M128 N128 K16, four per stage, into one of FOUR TMEM accumulators,
//               so the main loop of tile i+1 overlaps the epilogue of tile i
//   warps 2..9  epilogue (two warpgroups, each takes every other 32-column chunk), thread = output row: tcgen05.ld 32 columns at a time -> alpha, bias (staged in shared memory), ReLU,
//               + residual (R chunk brought in by TMA) -> results staged in 128B-swizzled shared memory and written with
//               TMA stores (full 128-byte lines; rows/columns beyond M/N are clipped by the tensor map).  The staging boxes
//               form rings (8 boxes of 16 KB shared by the residual loads and the f32 / bf16 stores; the f32 and the bf16
//               stores are issued by two different threads so each has its own bulk-group FIFO): a box is rewritten only
//               several stores later — waiting for the PREVIOUS store to release its box (a ~2500-clk round trip through the
//               TMA unit) cost 80 % of the time of the first version of this epilogue.
//               MERGED = true (bf16-only output, no residual): both 64-column halves of a tile are staged, then ONE proxy fence, ONE
//               barrier and one bulk group of two stores per tile instead of a fence / barrier / store per half plus a barrier at
//               the top of the tile (epilogue thread 0's timeline, scripts/gemm_trace.py: 3450 -> 2290 clk per tile).
// The previous version (one tile per CTA, each thread storing its own row with 16-byte st.global) spent its time in setup
// latency and 12-wavefront stores: 21-38 us per SuperGlue layer GEMM (profiles/).
//
// Reference behaviour replaced: the Conv1d(k=1)/Linear layers of thirdparty/SuperGlue/models/superglue.py:51-61,
// 100-128, 276-280 and thirdparty/LightGlue/lightglue/lightglue.py:133-216, 253-287 (cuBLAS sgemm via torch there).
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/icepy4d_b200.h"

#define GT_BM 128
#define GT_BN 128
#define GT_BK 64
#define GT_THREADS 320
#define GT_STAGE_BYTES ((GT_BM + GT_BN) * GT_BK * 2)          // 32 KB
#define GT_CHUNK_BYTES (GT_BM * 128)                            // 16 KB box: 128 rows x 128 B (32 f32 or 64 bf16 columns)
#define GT_NACC 4                                               // TMEM accumulators (4 x 128 columns = all of TMEM): the MMA warp runs
                                                                // up to four tiles ahead of the epilogue, which hides the hand-off latencies
// GT_STAGES (operand ring) and GT_NBOX (staging pool shared by the R loads and the C32 / C16 stores) are template parameters of the
// kernel: <3, 8> serves every output combination; <5, 4> is for the bf16-only outputs without a residual (three of the four GEMMs of
// a GNN layer), whose epilogue needs four boxes only.  The main loop of these short-K GEMMs is bound by the bytes the ring keeps in
// flight against the L2 -> shared-memory latency (3 x 32 KB per SM and ~1 us: 98 GB/s per SM measured), not by the tensor pipe.
#define GT_OFF_POOL (GT_STAGES * GT_STAGE_BYTES)
#define GT_OFF_BIAS (GT_OFF_POOL + GT_NBOX * GT_CHUNK_BYTES)    // 2 x 128 f32
constexpr int gt_smem_bytes(int stages, int nbox) { return stages * GT_STAGE_BYTES + nbox * GT_CHUNK_BYTES + 2 * GT_BN * 4; }

// Timeline instrumentation (scripts/gemm_trace.py builds a second library with -DGT_TRACE): CTA 0 stamps %clock64 at the
// synchronisation points of the producer lane, the MMA lane and the first epilogue thread for its first GT_TRACE_MAX tiles.
#ifdef GT_TRACE
#define GT_TRACE_MAX 32
__device__ unsigned long long gt_trace_buf[3][GT_TRACE_MAX][8];
#define GT_STAMP(role, tile, k) do { if (blockIdx.x == 0 && (tile) < GT_TRACE_MAX) { unsigned long long c_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)); gt_trace_buf[role][tile][k] = c_; } } while (0)
extern "C" __attribute__((visibility("default"))) int i4d_gemm_trace_dump(unsigned long long* host) {
  return cudaMemcpyFromSymbol(host, gt_trace_buf, sizeof(gt_trace_buf)) == cudaSuccess ? 0 : 1;
}
#else
#define GT_STAMP(role, tile, k) do { } while (0)
#endif

struct GemmTcParams {
  int M, N, K;
  float alpha;
  const float* bias;
  int has_r, has_c32, has_c16, relu;
  int dbg;   // experiments only (I4D_GEMM_DBG): 1 = no TMA stores, 2 = no proxy fence, 4 = no staging writes, 8 = no epilogue barrier
  const float* rot_cs;   // optional rotary tables [M, 64] (cos[32] | sin[32] per row): output columns < rot_cols are rotated in pairs
  int rot_cols;          // (multiple of 128; head_dim 64: column c belongs to pair (c % 64) / 2) — LightGlue's q / k, lightglue.py:49-57
};

__device__ __forceinline__ void gt_tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void gt_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void gt_wait_read(int pending) {          // cp.async.bulk.wait_group.read needs an immediate
  if (pending >= 6) asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory");
  else if (pending >= 4) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
  else if (pending >= 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
  else if (pending >= 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int GT_STAGES, int GT_NBOX, bool MERGED>
__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmW,
                                                                const __grid_constant__ CUtensorMap tmR,
                                                                const __grid_constant__ CUtensorMap tmC32,
                                                                const __grid_constant__ CUtensorMap tmC16, GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[GT_STAGES], empty_bar[GT_STAGES], acc_full[GT_NACC], acc_empty[GT_NACC], r_full[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = p.K / GT_BK;
  const int tiles_n = (p.N + GT_BN - 1) / GT_BN, tiles_m = (p.M + GT_BM - 1) / GT_BM;
  const int n_tiles = tiles_n * tiles_m;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
    if (p.has_r) tc::prefetch_tmap(&tmR);
    if (p.has_c32) tc::prefetch_tmap(&tmC32);
    if (p.has_c16) tc::prefetch_tmap(&tmC16);
    for (int s = 0; s < GT_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < GT_NACC; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_empty[b], 8); }
    for (int b = 0; b < 2; ++b) tc::mbar_init(&r_full[b], 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, GT_NACC * GT_BN);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (tc::elect_one()) {
      uint32_t s = 0, sph = 1;                                       // stage + the phase of its "empty" barrier to wait for
      [[maybe_unused]] int ti = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
        const int m0 = (t / tiles_n) * GT_BM, n0 = (t % tiles_n) * GT_BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (kb == 0) GT_STAMP(0, ti, 0);
          tc::mbar_wait(&empty_bar[s], sph);
          if (kb == 0) GT_STAMP(0, ti, 1);
          if (kb == kblocks - 1) GT_STAMP(0, ti, 2);
          uint8_t* sa = smem + s * GT_STAGE_BYTES;
          tc::mbar_arrive_expect_tx(&full_bar[s], GT_STAGE_BYTES);
          tc::tma_load_2d(sa, &tmA, &full_bar[s], kb * GT_BK, m0);
          tc::tma_load_2d(sa + GT_BM * GT_BK * 2, &tmW, &full_bar[s], kb * GT_BK, n0);
          if (++s == GT_STAGES) { s = 0; sph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer.  The loop is pure latency between two MMA groups: stage / phase
    // are carried as counters (no modulo), barrier addresses are kept in registers.
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc(GT_BM, GT_BN, 0, 0, 1);
      constexpr uint32_t hi = tc::desc_hi_sw128(1024);
      const uint32_t d0 = tc::desc_lo_sw128(tc::smem_u32(smem));
      const uint32_t a_full = tc::keep_in_register(tc::smem_u32(&full_bar[0])), a_empty = tc::keep_in_register(tc::smem_u32(&empty_bar[0]));
      const uint32_t a_acc_full = tc::keep_in_register(tc::smem_u32(&acc_full[0])), a_acc_empty = tc::keep_in_register(tc::smem_u32(&acc_empty[0]));
      uint32_t s = 0, sph = 0, b = 0, bph = 1;                       // smem stage + its phase, accumulator + the phase to wait for
      [[maybe_unused]] int ti = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
        GT_STAMP(1, ti, 0);
        tc::mbar_wait_a(a_acc_empty + b * 8, bph);                   // the epilogue has drained this accumulator
        GT_STAMP(1, ti, 1);
        tc::tcgen05_fence_after();
        const uint32_t acc = tmem_d + b * GT_BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          tc::mbar_wait_a(a_full + s * 8, sph);
          if (kb == 0) GT_STAMP(1, ti, 2);
          if (kb == kblocks - 1) GT_STAMP(1, ti, 3);
          tc::tcgen05_fence_after();
          const uint32_t da = d0 + s * (GT_STAGE_BYTES >> 4), db = da + ((GT_BM * GT_BK * 2) >> 4);
#pragma unroll
          for (int k = 0; k < GT_BK / 16; ++k) tc::umma_f16_parts(acc, da + k * 2, hi, db + k * 2, hi, idesc, (kb | k) ? 1u : 0u);
          tc::umma_commit_a(a_empty + s * 8);     // smem stage reusable once these MMAs retire
          if (++s == GT_STAGES) { s = 0; sph ^= 1u; }
        }
        tc::umma_commit_a(a_acc_full + b * 8);    // accumulator complete
        GT_STAMP(1, ti, 4);
        if (++b == GT_NACC) { b = 0; bph ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue: 8 warps = two warpgroups; thread = output row of the tile,
    // warpgroup wg takes the 32-column chunks c = wg, wg + 2 of every tile (one "step" = two chunks side by side)
    const int e = threadIdx.x - 64;                                  // 0..255
    const int wg = (warp - 2) >> 2;
    const int q = warp & 3;                                          // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                                   // row of the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t rsw = (uint32_t)(row & 7);
    // staging pool of 16 KB boxes: [R x2 if has_r][C32 ring][C16 ring]
    const int nR = p.has_r ? 2 : 0;
    const int n32 = p.has_c32 ? (p.has_c16 ? 4 : GT_NBOX - nR) : 0;
    const int n16 = p.has_c16 ? GT_NBOX - nR - n32 : 0;
    uint8_t* sR = smem + GT_OFF_POOL;
    uint8_t* sC32 = sR + nR * GT_CHUNK_BYTES;
    uint8_t* sC16 = sC32 + n32 * GT_CHUNK_BYTES;
    float* sBias = reinterpret_cast<float*>(smem + GT_OFF_BIAS);
    const bool lead32 = (e == 0), lead16 = (e == 32);                // two leaders: each owns the bulk-group FIFO of its stores
    uint32_t i = 0;
    // R chunk `gc` (counted over all of this CTA's tiles, 4 per tile) -> buffer gc & 1
    auto issue_r = [&](uint32_t gc) {
      const uint32_t ti = gc >> 2, c = gc & 3;
      const long long t = (long long)blockIdx.x + (long long)ti * gridDim.x;
      if (t >= n_tiles) return;
      const int m0 = (int)(t / tiles_n) * GT_BM, n0 = (int)(t % tiles_n) * GT_BN;
      tc::mbar_arrive_expect_tx(&r_full[gc & 1], GT_CHUNK_BYTES);
      tc::tma_load_2d(sR + (gc & 1) * GT_CHUNK_BYTES, &tmR, &r_full[gc & 1], n0 + (int)c * 32, m0);
    };
    if (lead32 && p.has_r) { issue_r(0); issue_r(1); }
    // bias of a tile is fetched one tile ahead (an exposed L2 round trip + barrier per tile was the longest link of the
    // epilogue's dependency chain) and parked in the double-buffered shared-memory row at the end of the previous tile
    // tile coordinates advance by (gridDim.x / tiles_n, gridDim.x % tiles_n) with a carry: no integer division on the per-tile chain
    // (the timeline showed ~900 clk per tile between the stamps around the two divisions of the old loop head)
    const int dq = (int)gridDim.x / tiles_n, dr = (int)gridDim.x % tiles_n;
    int tm = (int)blockIdx.x / tiles_n, tn = (int)blockIdx.x % tiles_n;
    auto bias_at = [&](int tile_n, bool exists) -> float {
      if (!exists || e >= GT_BN || !p.bias) return 0.f;
      return __ldg(p.bias + min(tile_n * GT_BN + e, p.N - 1));
    };
    if (e < GT_BN) sBias[e] = bias_at(tn, (int)blockIdx.x < n_tiles);
    if (MERGED) gt_epi_bar();                                        // bias row of the first tile visible
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
      const int m0 = tm * GT_BM, n0 = tn * GT_BN;
      int tn_next = tn + dr, tm_next = tm + dq;
      if (tn_next >= tiles_n) { tn_next -= tiles_n; ++tm_next; }
      const uint32_t b = i % GT_NACC, bb2 = i & 1;
      float* bs = sBias + bb2 * GT_BN;
      const float bias_next = bias_at(tn_next, t + (int)gridDim.x < n_tiles);   // in flight during this tile
      // rotary tables of my row: both of my chunks start at column 32 wg of a 64-wide head, i.e. they use the same 16 pairs; the
      // loads are in flight while the accumulator is awaited
      const bool rot = p.rot_cs != nullptr && n0 < p.rot_cols;
      float4 rc[4], rs[4];
      if (rot) {
        const float4* cr = reinterpret_cast<const float4*>(p.rot_cs + (size_t)min(m0 + row, p.M - 1) * 64 + wg * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) { rc[j] = __ldg(cr + j); rs[j] = __ldg(cr + 8 + j); }
      }
      if (e == 0) GT_STAMP(2, i, 0);
      if (!MERGED) gt_epi_bar();                                     // bias row of this tile visible (written a tile ago)
      if (e == 0) GT_STAMP(2, i, 1);
      tc::mbar_wait(&acc_full[b], (i / GT_NACC) & 1);
      if (e == 0) GT_STAMP(2, i, 2);
      tc::tcgen05_fence_after();
      // both of my chunks leave TMEM together (two loads in flight), and the accumulator goes back to the MMA warp at once
      uint32_t vv[2][32];
      tc::tmem_ld32(tmem_d + lane_off + b * GT_BN + wg * 32, vv[0]);
      tc::tmem_ld32(tmem_d + lane_off + b * GT_BN + (2 + wg) * 32, vv[1]);
      tc::tmem_ld_wait();
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[b]);
      if (e == 0) GT_STAMP(2, i, 3);
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        const int c = 2 * st + wg;                                   // my chunk of this step
        const uint32_t S = 2 * i + st;                               // global step index
        const uint32_t g = 2 * S + wg;                               // global chunk index
        const uint32_t (&v)[32] = vv[st];
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bs + c * 32 + j);
          f[j] = fmaf(p.alpha, __uint_as_float(v[j]), bb.x); f[j + 1] = fmaf(p.alpha, __uint_as_float(v[j + 1]), bb.y);
          f[j + 2] = fmaf(p.alpha, __uint_as_float(v[j + 2]), bb.z); f[j + 3] = fmaf(p.alpha, __uint_as_float(v[j + 3]), bb.w);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (rot) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {                                // 4 pairs per float4 of cos / sin
            const float cc[4] = {rc[j].x, rc[j].y, rc[j].z, rc[j].w}, ss[4] = {rs[j].x, rs[j].y, rs[j].z, rs[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x = f[8 * j + 2 * e], y = f[8 * j + 2 * e + 1];
              f[8 * j + 2 * e] = x * cc[e] - y * ss[e];
              f[8 * j + 2 * e + 1] = y * cc[e] + x * ss[e];
            }
          }
        }
        if (p.has_r) {
          tc::mbar_wait(&r_full[wg], S & 1);                         // R chunk g lives in buffer g & 1 = wg; its S-th use
          const uint8_t* rb = sR + wg * GT_CHUNK_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r4 = *reinterpret_cast<const float4*>(rb + ((((uint32_t)j) ^ rsw) << 4));
            f[4 * j] += r4.x; f[4 * j + 1] += r4.y; f[4 * j + 2] += r4.z; f[4 * j + 3] += r4.w;
          }
        }
        // The boxes written below were released by their previous stores: the leaders waited for that before the barrier of the
        // PREVIOUS step (C32 box g % n32: store g - n32; C16 box S % n16: store S - n16).
        if (p.has_c32) {
          uint8_t* cb = sC32 + (g % (uint32_t)n32) * GT_CHUNK_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(cb + ((((uint32_t)j) ^ rsw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        if (p.has_c16 && !(p.dbg & 4)) {
          uint8_t* hb = sC16 + (S % (uint32_t)n16) * GT_CHUNK_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 a = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]), b2 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
            __nv_bfloat162 c2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]), d = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b2);
            pk.z = *reinterpret_cast<uint32_t*>(&c2); pk.w = *reinterpret_cast<uint32_t*>(&d);
            *reinterpret_cast<uint4*>(hb + ((((uint32_t)(wg * 4 + j)) ^ rsw) << 4)) = pk;
          }
        }
        if (MERGED) continue;                                        // one fence / barrier / store group per TILE, below
        if (!(p.dbg & 2)) tc::fence_proxy_async_smem();              // my staged results -> visible to the TMA engine
        // ring discipline: before anybody writes the next step's boxes, the stores that last used them must have read them.
        //   C32 (two stores per step): next step writes boxes (g0+2, g0+3) % n32, last used by stores g0+2-n32, g0+3-n32; this
        //        leader has issued stores <= g0-1, so at most n32 - 4 of them may still be pending.
        //   C16 (one store per step): next step writes box (S+1) % n16, last used by store S+1-n16: at most n16 - 2 pending.
        if (e == 0) GT_STAMP(2, i, 4 + 2 * st);
        if (lead32 && p.has_c32) gt_wait_read(n32 - 4);
        if (lead16 && p.has_c16) gt_wait_read(n16 - 2);
        if (!(p.dbg & 8)) gt_epi_bar();
        if (e == 0) GT_STAMP(2, i, 5 + 2 * st);
        if (lead32) {
          if (p.has_c32) {
            const uint32_t g0 = 2 * S;
            gt_tma_store_2d(&tmC32, sC32 + (g0 % (uint32_t)n32) * GT_CHUNK_BYTES, n0 + (2 * st) * 32, m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            gt_tma_store_2d(&tmC32, sC32 + ((g0 + 1) % (uint32_t)n32) * GT_CHUNK_BYTES, n0 + (2 * st + 1) * 32, m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (p.has_r) { issue_r(2 * S + 2); issue_r(2 * S + 3); }   // every reader of both R buffers has passed the barrier
        }
        if (lead16 && p.has_c16 && !(p.dbg & 1)) {
          gt_tma_store_2d(&tmC16, sC16 + (S % (uint32_t)n16) * GT_CHUNK_BYTES, n0 + st * 64, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (e < GT_BN) sBias[(bb2 ^ 1) * GT_BN + e] = bias_next;         // its previous reader (tile i-1) finished a tile ago
      if (MERGED) {
        // bf16-only output without a residual: both 64-column halves of the tile are staged (boxes 2i, 2i + 1 of the ring), then ONE
        // proxy fence, ONE barrier (which also publishes the next tile's bias row) and one bulk group of two stores per tile.
        // Ring discipline: the next tile writes boxes (2i + 2, 2i + 3) % n16, last used by the group of tile i + 1 - n16 / 2; this
        // leader has committed the groups of tiles <= i - 1, so at most n16 / 2 - 2 of them may still be reading.
        tc::fence_proxy_async_smem();
        if (e == 0) GT_STAMP(2, i, 4);
        if (lead16) gt_wait_read(n16 / 2 - 2);
        gt_epi_bar();
        if (e == 0) GT_STAMP(2, i, 5);
        if (lead16) {
          gt_tma_store_2d(&tmC16, sC16 + ((2 * i) % (uint32_t)n16) * GT_CHUNK_BYTES, n0, m0);
          gt_tma_store_2d(&tmC16, sC16 + ((2 * i + 1) % (uint32_t)n16) * GT_CHUNK_BYTES, n0 + 64, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tm = tm_next; tn = tn_next;
    }
    if (lead32 || lead16) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_d, GT_NACC * GT_BN);
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

static int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                        uint32_t box_cols, CUtensorMapDataType dt, uint32_t esize) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { i4d_set_error("cuTensorMapEncodeTiled entry point not available"); return I4D_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * esize) & 15)) {
    i4d_set_error("TMA needs a 16-byte aligned base and row pitch (base %p, ld %llu)", base, (unsigned long long)ld);
    return I4D_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { i4d_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return I4D_ERR_CUDA; }
  return I4D_OK;
}
int i4d_make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                          uint32_t box_cols) {
  return make_tmap_2d(out, base, rows, cols, ld, box_rows, box_cols, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
}
// row-major f32 matrix: box = box_cols x box_rows elements (box_cols * 4 bytes must be 128)
static int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                            uint32_t box_cols) {
  return make_tmap_2d(out, base, rows, cols, ld, box_rows, box_cols, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
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

static int gemm_tc_launch(const void* A, int lda, const void* W, int ldw, const float* bias, const float* R, int ldr, float* C32,
                          int ldc32, void* C16, int ldc16, int M, int N, int K, float alpha, int relu, const float* rot_cs,
                          int rot_cols, void* stream) {
  I4D_CHECK_ARG(A && W && (C32 || C16), "null pointer");
  I4D_CHECK_ARG(M > 0 && N > 0 && K > 0, "bad sizes");
  I4D_CHECK_ARG(K % GT_BK == 0, "K must be a multiple of 64 for the tensor-core GEMM");
  I4D_CHECK_ARG(lda >= K && ldw >= K, "leading dimensions too small");
  I4D_CHECK_ARG(!C32 || ((ldc32 & 3) == 0 && (reinterpret_cast<uintptr_t>(C32) & 15) == 0), "C32 must be 16-byte aligned, ldc32 % 4 == 0");
  I4D_CHECK_ARG(!C16 || ((ldc16 & 7) == 0 && (reinterpret_cast<uintptr_t>(C16) & 15) == 0), "C16 must be 16-byte aligned, ldc16 % 8 == 0");
  I4D_CHECK_ARG(!R || ((ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0), "R must be 16-byte aligned, ldr % 4 == 0");
  CUtensorMap tmA, tmW, tmR, tmC32, tmC16;
  if (int rc = i4d_make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GT_BM, GT_BK)) return rc;
  if (int rc = i4d_make_tmap_2d_bf16(&tmW, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, GT_BN, GT_BK)) return rc;
  tmR = tmA; tmC32 = tmA; tmC16 = tmA;                                // placeholders for the operands that are absent
  if (R) { if (int rc = make_tmap_2d_f32(&tmR, R, (uint64_t)M, (uint64_t)N, (uint64_t)ldr, GT_BM, 32)) return rc; }
  if (C32) { if (int rc = make_tmap_2d_f32(&tmC32, C32, (uint64_t)M, (uint64_t)N, (uint64_t)ldc32, GT_BM, 32)) return rc; }
  if (C16) { if (int rc = i4d_make_tmap_2d_bf16(&tmC16, C16, (uint64_t)M, (uint64_t)N, (uint64_t)ldc16, GT_BM, 64)) return rc; }
  static bool attr_seen[64] = {};
  if (i4d_first_use_on_device(attr_seen)) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<3, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem_bytes(3, 8)));
    I4D_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<5, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem_bytes(5, 4)));
    I4D_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<3, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem_bytes(3, 8)));
    I4D_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<4, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem_bytes(4, 6)));
    I4D_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<5, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem_bytes(5, 4)));
  }
  static int dbg = -1, deep_ring = -1;
  if (dbg < 0) { const char* e = getenv("I4D_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
  if (deep_ring < 0) { const char* e = getenv("I4D_GEMM_DEEP_RING"); deep_ring = e ? atoi(e) : 1; }   // 0: <3, 8> for everything (A/B runs)
  GemmTcParams p{M, N, K, alpha, bias, R ? 1 : 0, C32 ? 1 : 0, C16 ? 1 : 0, relu, dbg, rot_cs, rot_cols};
  const int n_tiles = i4d_cdiv(N, GT_BN) * i4d_cdiv(M, GT_BM);
  const int grid = n_tiles < i4d_num_sms() ? n_tiles : i4d_num_sms();
  // bf16-only outputs without a residual (three of the four GEMMs of a GNN layer) take the merged epilogue protocol: <4, 6> for K < 512,
  // <5, 4> from K = 512 on (16384 x 768: K = 256 12.1 / 12.3 us, K = 512 16.8 / 16.2 us; per-step protocol <3, 8>: 14.8 / 18.4 us).
  // I4D_GEMM_VARIANT (experiments): 0 = <3, 8> per-step protocol for everything, 1 = <5, 4> per-step, 2 = <3, 8> merged, 3 = <4, 6> merged,
  // 4 = <5, 4> merged
  static int variant_env = -2;
  if (variant_env == -2) { const char* e = getenv("I4D_GEMM_VARIANT"); variant_env = e ? atoi(e) : -1; }
  const int variant = variant_env >= 0 ? variant_env : (!deep_ring ? 0 : (K >= 512 ? 4 : 3));
  const bool simple_out = C16 && !C32 && !R;
  const cudaStream_t cs = (cudaStream_t)stream;
  if (simple_out && variant == 1) gemm_tc_kernel<5, 4, false><<<grid, GT_THREADS, gt_smem_bytes(5, 4), cs>>>(tmA, tmW, tmR, tmC32, tmC16, p);
  else if (simple_out && variant == 2) gemm_tc_kernel<3, 8, true><<<grid, GT_THREADS, gt_smem_bytes(3, 8), cs>>>(tmA, tmW, tmR, tmC32, tmC16, p);
  else if (simple_out && variant == 3) gemm_tc_kernel<4, 6, true><<<grid, GT_THREADS, gt_smem_bytes(4, 6), cs>>>(tmA, tmW, tmR, tmC32, tmC16, p);
  else if (simple_out && variant == 4) gemm_tc_kernel<5, 4, true><<<grid, GT_THREADS, gt_smem_bytes(5, 4), cs>>>(tmA, tmW, tmR, tmC32, tmC16, p);
  else gemm_tc_kernel<3, 8, false><<<grid, GT_THREADS, gt_smem_bytes(3, 8), cs>>>(tmA, tmW, tmR, tmC32, tmC16, p);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_gemm_bf16_tc(
    const void* A, int lda, const void* W, int ldw, const float* bias, const float* R, int ldr, float* C32, int ldc32,
    void* C16, int ldc16, int M, int N, int K, float alpha, int relu, void* stream) {
  return gemm_tc_launch(A, lda, W, ldw, bias, R, ldr, C32, ldc32, C16, ldc16, M, N, K, alpha, relu, nullptr, 0, stream);
}

// LightGlue's fused QKV projection with the rotary embedding applied to q and k in the epilogue (K12 of SURVEY.md's kernel list):
// C16 = rotary(A W^T + bias) for the output columns < rot_cols, plain A W^T + bias for the rest (v).
extern "C" __attribute__((visibility("default"))) int i4d_gemm_bf16_tc_rotary(
    const void* A, int lda, const void* W, int ldw, const float* bias, const float* cs, int rot_cols, void* C16, int ldc16,
    int M, int N, int K, void* stream) {
  I4D_CHECK_ARG(cs && C16, "null pointer");
  I4D_CHECK_ARG(rot_cols >= 0 && rot_cols % GT_BN == 0 && (reinterpret_cast<uintptr_t>(cs) & 15) == 0,
                "rot_cols must be a multiple of 128 and cs 16-byte aligned");
  return gemm_tc_launch(A, lda, W, ldw, bias, nullptr, 0, nullptr, 0, C16, ldc16, M, N, K, 1.f, 0, cs, rot_cols, stream);
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

// ---- f32 -> three bf16 column blocks for a split-precision product through the plain bf16 GEMM ------------------------
// x = hi + lo + O(2^-17 |x|) with hi = bf16(x), lo = bf16(x - hi).  Concatenated along K,
//   activations  A' = [hi | hi | lo]   (order 0)        weights  W' = [hi | lo | hi]   (order 1)
// give  A' W'^T = A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T  accumulated in f32 by ONE tensor-core GEMM of depth 3K: the three-product
// scheme of the SuperPoint convolutions (csrc/conv_tc.cu) for the layers that must keep ~f32 accuracy (SuperGlue's keypoint encoder,
// thirdparty/SuperGlue/models/superglue.py:51-61,67-78, whose input carries the keypoint positions).
__global__ void __launch_bounds__(256) f32_split3_bf16_kernel(const float* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ Y,
                                                              int ldy, int rows, int cols, int order) {
  const int cols4 = cols >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols4) return;
  const int r = i / cols4, c = (i - r * cols4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + c);
  const float x[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    hi[k] = __float2bfloat16_rn(x[k]);
    lo[k] = __float2bfloat16_rn(x[k] - __bfloat162float(hi[k]));
  }
  uint2 ph, pl;
  ph.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
  ph.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
  pl.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
  pl.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
  __nv_bfloat16* y = Y + (size_t)r * ldy + c;
  *reinterpret_cast<uint2*>(y) = ph;
  *reinterpret_cast<uint2*>(y + cols) = order == 0 ? ph : pl;
  *reinterpret_cast<uint2*>(y + 2 * cols) = order == 0 ? pl : ph;
}

extern "C" __attribute__((visibility("default"))) int i4d_f32_split3_bf16(const float* X, int ldx, void* Y, int ldy, int rows,
                                                                         int cols, int order, void* stream) {
  I4D_CHECK_ARG(X && Y && rows >= 0 && cols > 0, "bad arguments");
  I4D_CHECK_ARG((cols & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0 && ldy >= 3 * cols, "cols, ldx, ldy must be multiples of 4 and ldy >= 3 cols");
  I4D_CHECK_ARG(order == 0 || order == 1, "order must be 0 (activations: hi|hi|lo) or 1 (weights: hi|lo|hi)");
  if (rows == 0) return I4D_OK;
  const long long n = (long long)rows * (cols / 4);
  f32_split3_bf16_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, reinterpret_cast<__nv_bfloat16*>(Y), ldy, rows,
                                                                            cols, order);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
