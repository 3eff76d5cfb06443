// Fused persistent Sinkhorn kernel and its launcher: included TWICE by assignment.cu, inside namespaces sk512 / sk256, with
// SK_THREADS = 512 (16 columns per thread, 128 registers) and 256 (32 columns per thread, 255 registers: half the per-stage
// bookkeeping per matrix element, shared-memory addresses pinned in registers).  See the design comment in assignment.cu.

#ifndef SK_THREADS
#error "define SK_THREADS (256 or 512) before including sinkhorn_fused.inl"
#endif
#define SK_WARPS (SK_THREADS / 32)
#define SK_ROWS 2                 // rows per stage
#define SK_STAGES 3
#define SK_MAXN 8192
#define SK_GROUPS (SK_MAXN / 4 / SK_THREADS)   // float4 column groups per thread = 4
#define SK_EXACT_ITERS 1
#define SK_MAX_BAND 1024            // rows per CTA the previous-u staging buffer can hold
#define SK_KEEP_PCT_DEFAULT 0
// row pitch of the shared-memory stages for NP (= N rounded up to 4) columns
__host__ __device__ inline int sk_smem_pitch(int NP) { return (NP * 8 >= SK_MAXN * 7) ? SK_MAXN : NP; }

__device__ __forceinline__ float sk_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct L2Acc {   // running (max, sum) in the log2 domain, one exp per update
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void add(float x) {
    float d = x - m;
    float e = sk_ex2(-fabsf(d));
    s = (d > 0.f) ? fmaf(s, e, 1.f) : (s + e);
    m = fmaxf(m, x);
  }
  __device__ __forceinline__ void merge(float om, float os) {
    float nm = fmaxf(m, om);
    if (nm == -INFINITY) return;
    s = s * exp2f(m - nm) + os * exp2f(om - nm);
    m = nm;
  }
  __device__ __forceinline__ float lse_log2() const { return m + log2f(s); }
  __device__ __forceinline__ float lse_ln() const { return (m + log2f(s)) * LN2; }
};

__device__ __forceinline__ void sk_mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void sk_mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
  uint32_t ok = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void sk_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)),
                 "l"(policy)
               : "memory");
}

struct SkCtx {
  const float* S; float* stage_buf; uint64_t* full; uint64_t* bpart;
  float (*part_m)[SK_ROWS][SK_WARPS]; float (*part_s)[SK_ROWS][SK_WARPS];
  float* u; const float* v; float* pm; float* ps; int* flag;
  float* unew_s;         // this iteration's u of the band, log2 domain (shared memory; the next iteration's uold_s)
  int M, N, NP, NS, ld, n4, row0, row1, nst, cta, keep_rows, pf, tid, warp, lane; uint32_t total;   // NP = 4 * n4: N rounded up to whole float4 groups; NS = row pitch of the smem stages
  float norm, c_mu, c_nu, extra_row, kfac;
  const float* uold_s;   // previous-iteration u of this CTA's band, log2 domain (shared memory)
  uint32_t row_bytes;
  // 32-bit shared-memory addresses of everything the fast stage loop touches, derived once per thread (SK_PIN: pinned in registers)
  uint32_t a_stage;      // stage_buf + tid * 16 (this thread's first float4 column group of stage buffer 0)
  uint32_t stage_bytes;  // SK_ROWS * NS * 4
  uint32_t pitch_bytes;  // NS * 4
  uint32_t a_full, a_bpart;
  uint32_t a_part_st;    // &part_s[0][lane >> 4][warp]          (written by lanes 0 / 16)
  uint32_t a_part_ld;    // &part_s[0][lane >> 4][(lane & (SK_WARPS / 2 - 1)) * 2] (read by every lane)
  uint32_t a_uold, a_unew;
};
__device__ __forceinline__ uint32_t sk_pin(uint32_t x) { asm volatile("" : "+r"(x)); return x; }   // used by the 256-thread instantiation only
__device__ __forceinline__ uint32_t sk_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float sk_lds(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 sk_lds2(uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float4 sk_lds4(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sk_sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sk_mbar_arrive_a(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void sk_mbar_wait_a(uint32_t a, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (spin > (1u << 26)) __trap();
  }
}
// ring / barrier phase bookkeeping carried across stages, bands and iterations (one copy per thread, all identical)
// Timeline instrumentation (-DSK_TRACE, scripts/sinkhorn_trace.py): CTA 0 stamps %clock64 at the synchronisation points of its
// fast-mode stages (warps 0 and 9) and of the iteration boundary (thread 0).
#if defined(SK_TRACE) && SK_THREADS == 256
#define SK_TRACE_MAX 512
__device__ unsigned long long sk_trace_buf[2][SK_TRACE_MAX][8];
#define SK_STAMP(fq_, k) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 160) && (fq_) < 440) { unsigned long long c_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); sk_trace_buf[threadIdx.x ? 1 : 0][(fq_)][k] = c_; } } while (0)
#define SK_BSTAMP(it_, k) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (it_) < 64) { unsigned long long c_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); sk_trace_buf[1][448 + (it_)][k] = c_; } } while (0)
extern "C" __attribute__((visibility("default"))) int i4d_sinkhorn_trace_dump(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, sk_trace_buf, sizeof(sk_trace_buf));
}
#else
#define SK_STAMP(fq_, k) do { } while (0)
#define SK_BSTAMP(it_, k) do { } while (0)
#endif
struct SkRing {
  uint32_t seq;        // stages consumed so far (all iterations)
  uint32_t buf;        // seq % SK_STAGES
  uint32_t full_par;   // bit b: parity the next wait on full[b] expects
  uint32_t fq;         // fast-mode stages so far: partial-sum slot fq & 1, duty warp fq % SK_WARPS
  uint32_t bp_par;     // bit p: parity the next wait on bpart[p] expects
  uint32_t total;      // stages to run in total
  __device__ __forceinline__ void advance() { ++seq; buf = (buf == SK_STAGES - 1) ? 0u : buf + 1u; }
};

// Stages are numbered consecutively across iterations (S never changes, so the ring keeps streaming through the grid
// barriers).  Stage `seq` lives in buffer seq % SK_STAGES and covers row block sk_idx(seq) of the band: forward on even
// passes, backward on odd ones.
__device__ __forceinline__ int sk_idx(uint32_t seq, int nst) {
  const uint32_t pass = seq / (uint32_t)nst, r = seq - pass * (uint32_t)nst;
  return (pass & 1u) ? (nst - 1 - (int)r) : (int)r;
}
// one thread: start the bulk loads of stage `seq`.  Invariant kept by the callers: stage seq + SK_STAGES is issued as soon as
// every thread has taken stage seq out of shared memory.
__device__ __forceinline__ void sk_issue(const SkCtx& c, uint32_t seq) {
  const int buf = seq % SK_STAGES;
  const int r = c.row0 + sk_idx(seq, c.nst) * SK_ROWS;
  const int nr = min(SK_ROWS, c.row1 - r);
  // the first `keep_rows` rows of the band are asked to stay in L2 (evict_last), the rest to stream through (evict_first)
  uint64_t policy;
  if (r - c.row0 < c.keep_rows) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  sk_mbar_expect(&c.full[buf], nr * c.row_bytes);
  for (int k = 0; k < nr; ++k)
    sk_bulk_load(c.stage_buf + ((size_t)buf * SK_ROWS + k) * c.NS, c.S + (size_t)(r + k) * c.ld, c.row_bytes, &c.full[buf], policy);
  if (c.pf > 0 && seq + (uint32_t)c.pf < c.total) {      // pull a later stage of the band from HBM into L2 ahead of its smem fill
    const int rp = c.row0 + sk_idx(seq + (uint32_t)c.pf, c.nst) * SK_ROWS;
    const int np = min(SK_ROWS, c.row1 - rp);
    for (int k = 0; k < np; ++k)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c.S + (size_t)(rp + k) * c.ld), "r"(c.row_bytes) : "memory");
  }
}
__device__ __forceinline__ void sk_wait_full(const SkCtx& c, SkRing& rg) {
#if SK_THREADS == 256
  sk_mbar_wait_a(c.a_full + rg.buf * 8u, (rg.full_par >> rg.buf) & 1u);
#else
  sk_mbar_wait(&c.full[rg.buf], (rg.full_par >> rg.buf) & 1u);
#endif
  rg.full_par ^= 1u << rg.buf;
}

// ---- exact mode: one stage (SK_ROWS rows): elements of my columns -> registers, row-pass partials, ONE block barrier, then
// every warp finishes the row reduction itself and runs the column pass from registers.
// FULL = all SK_ROWS rows and all SK_GROUPS column groups are present (straight-line code without predicates).
template <bool FULL>
__device__ __forceinline__ void sk_stage_exact(const SkCtx& c, int idx, SkRing& rg, const float (&vl)[SK_GROUPS][4],
                                               L2Acc (&col)[SK_GROUPS][4]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = c.NS, n4 = c.n4;
  const int pp = rg.seq & 1;
  const int r_base = c.row0 + idx * SK_ROWS;
  const int nr = FULL ? SK_ROWS : min(SK_ROWS, c.row1 - r_base);
  const float* sb = c.stage_buf + (size_t)rg.buf * SK_ROWS * N;
  sk_wait_full(c, rg);
  float x[SK_ROWS][SK_GROUPS][4];
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k)
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int gi = g * SK_THREADS + tid;
      float4 t = (FULL || (k < nr && gi < n4)) ? reinterpret_cast<const float4*>(sb + (size_t)k * N)[gi] : make_float4(0.f, 0.f, 0.f, 0.f);
      x[k][g][0] = t.x * LOG2E; x[k][g][1] = t.y * LOG2E; x[k][g][2] = t.z * LOG2E; x[k][g][3] = t.w * LOG2E;
    }
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) {
    if (FULL || k < nr) {
      L2Acc a, a1; a.init(); a1.init();
#pragma unroll
      for (int g = 0; g < SK_GROUPS; ++g) {
        if (FULL || g * SK_THREADS + tid < n4) {
          a.add(x[k][g][0] + vl[g][0]); a1.add(x[k][g][1] + vl[g][1]); a.add(x[k][g][2] + vl[g][2]); a1.add(x[k][g][3] + vl[g][3]);
        }
      }
      a.merge(a1.m, a1.s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
      if (lane == 0) { c.part_m[pp][k][warp] = a.m; c.part_s[pp][k][warp] = a.s; }
    }
  }
  __syncthreads();              // partials published; every thread has its elements in registers -> the buffer is free
  if (tid == 0 && rg.seq + SK_STAGES < rg.total) sk_issue(c, rg.seq + SK_STAGES);
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) {
    if (FULL || k < nr) {
      L2Acc a; a.init();
      if (lane < SK_WARPS) { a.m = c.part_m[pp][k][lane]; a.s = c.part_s[pp][k][lane]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
      a.add(c.extra_row);
      const float ui = c.norm - a.lse_ln();
      if (tid == 0) c.u[r_base + k] = ui;
      const float ul = ui * LOG2E;
#pragma unroll
      for (int g = 0; g < SK_GROUPS; ++g) {
        if (FULL || g * SK_THREADS + tid < n4) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) col[g][cc].add(x[k][g][cc] + ul);
        }
      }
    }
  }
  rg.advance();
}

__device__ __forceinline__ void sk_band_exact(const SkCtx& c, SkRing& rg, const float4 (&vraw)[SK_GROUPS]) {
  const int tid = threadIdx.x;
  const int N = c.NP, n4 = c.n4;
  float vl[SK_GROUPS][4];                                          // old v of my columns, log2 domain
  L2Acc col[SK_GROUPS][4];
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    const float4 t = vraw[g];
    vl[g][0] = t.x * LOG2E; vl[g][1] = t.y * LOG2E; vl[g][2] = t.z * LOG2E; vl[g][3] = t.w * LOG2E;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) col[g][cc].init();
  }
  const bool full_cols = c.NS == SK_MAXN;
  const bool rev = ((rg.seq / (uint32_t)c.nst) & 1u) != 0;
#pragma unroll 1
  for (int st = 0; st < c.nst; ++st) {
    const int idx = rev ? c.nst - 1 - st : st;
    if (full_cols && c.row0 + (idx + 1) * SK_ROWS <= c.row1) sk_stage_exact<true>(c, idx, rg, vl, col);
    else sk_stage_exact<false>(c, idx, rg, vl, col);
  }
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    const int gi = g * SK_THREADS + tid;
    if (gi < n4) {
      reinterpret_cast<float4*>(c.pm + (size_t)c.cta * N)[gi] = make_float4(col[g][0].m, col[g][1].m, col[g][2].m, col[g][3].m);
      reinterpret_cast<float4*>(c.ps + (size_t)c.cta * N)[gi] = make_float4(col[g][0].s, col[g][1].s, col[g][2].s, col[g][3].s);
    }
  }
}

// ---- fast mode --------------------------------------------------------------------------------------------------------
// Column pass of a stage (row block `idx`, stage number seq, fast-stage number fq) from the registers kept since its
// phase A: cs_j += e_ij * a_i.  Every warp finishes the row sums itself as soon as all partials are published (mbarrier
// bpart): a_i = 2^(u_i + m_i - c_mu) = kfac / rowsum_i, one reciprocal.  The stage's duty warp (round-robin) also writes u_i,
// checks the sums and refills the shared-memory buffer the stage has released; nobody waits for it.
template <bool FULL>
__device__ __forceinline__ void sk_col_accum(const SkCtx& c, int idx, uint32_t seq, uint32_t fq, SkRing& rg,
                                             const float (&e)[SK_ROWS][SK_GROUPS][4], float2 (&cs)[SK_GROUPS][2], float& csN,
                                             bool refill = true) {
  const int warp = c.warp, lane = c.lane;
  const uint32_t pp = fq & 1u, slot = fq & 3u;
  const int row = lane >> 4;
  const bool valid = FULL || (c.row0 + idx * SK_ROWS + row < c.row1);
#if SK_THREADS == 256
  const float mr = valid ? (c.c_nu - sk_lds(c.a_uold + (uint32_t)(idx * SK_ROWS + row) * 4u)) : 0.f;
#else
  const float mr = valid ? (c.c_nu - c.uold_s[idx * SK_ROWS + row]) : 0.f;
#endif
  const float ex = sk_ex2(c.extra_row - mr);                        // dustbin column term
#if SK_THREADS == 256
  sk_mbar_wait_a(c.a_bpart + pp * 8u, (rg.bp_par >> pp) & 1u);      // all warps: partials published, buffer emptied
#else
  sk_mbar_wait(&c.bpart[pp], (rg.bp_par >> pp) & 1u);               // all warps: partials published, buffer emptied
#endif
  SK_STAMP(fq + 1u, 4);
  rg.bp_par ^= 1u << pp;
#if SK_THREADS == 256
  const float2 pr = sk_lds2(c.a_part_ld + slot * (uint32_t)(SK_ROWS * SK_WARPS * 4));                           // SK_WARPS partials, two per lane
#else
  const float2 pr = *reinterpret_cast<const float2*>(&c.part_s[slot][row][(lane & (SK_WARPS / 2 - 1)) * 2]);   // SK_WARPS partials, two per lane
#endif
  float t = pr.x + pr.y;
#pragma unroll
  for (int o = SK_WARPS / 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  const float sm = t + ex;
  const float a = valid ? __fdividef(c.kfac, sm) : 0.f;
  csN = fmaf(ex, a, csN);                                           // dustbin column (every lane; lanes 0 / 16 of warp 0 are read)
  const float a0 = __shfl_sync(0xffffffffu, a, 0), a1 = __shfl_sync(0xffffffffu, a, 16);
  const float2 a0v = make_float2(a0, a0), a1v = make_float2(a1, a1);
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    cs[g][0] = __ffma2_rn(make_float2(e[0][g][0], e[0][g][1]), a0v, cs[g][0]);
    cs[g][1] = __ffma2_rn(make_float2(e[0][g][2], e[0][g][3]), a0v, cs[g][1]);
  }
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    cs[g][0] = __ffma2_rn(make_float2(e[1][g][0], e[1][g][1]), a1v, cs[g][0]);
    cs[g][1] = __ffma2_rn(make_float2(e[1][g][2], e[1][g][3]), a1v, cs[g][1]);
  }
  if (warp == (int)(fq & (SK_WARPS - 1))) {                                               // duty: off everybody's critical path
    if (refill && lane == 0 && seq + SK_STAGES < rg.total) sk_issue(c, seq + SK_STAGES);
    if ((lane & 15) == 0 && valid) {
      if (!(sm > 0.f && sm < INFINITY)) atomicExch(c.flag, 1);
      const float ul = c.norm * LOG2E - (mr + __log2f(sm));
      c.unew_s[idx * SK_ROWS + row] = ul;
      c.u[c.row0 + idx * SK_ROWS + row] = ul * LN2;
    }
  }
}

// One stage: phase A of row block `idx` (results in e_cur), then the column pass of the previous stage (from e_prev): a
// warp only waits for the slowest warp of the PREVIOUS stage after finishing its own share of this one.
// FULL = both rows exist and every thread owns SK_GROUPS complete float4 column groups (N == SK_MAXN).
template <bool FULL, bool HAVE_PREV, bool PREV_FULL>
__device__ __forceinline__ void sk_stage_fast(const SkCtx& c, int idx, int idx_prev, SkRing& rg,
                                              const float2 (&vl)[SK_GROUPS][2], float2 (&cs)[SK_GROUPS][2], float& csN,
                                              float (&e_cur)[SK_ROWS][SK_GROUPS][4], const float (&e_prev)[SK_ROWS][SK_GROUPS][4]) {
  const int tid = c.tid, warp = c.warp, lane = c.lane;
  const int N = c.NS, n4 = c.n4;
  // partial sums go to slot fq % 4: a warp may run a stage ahead of another one that has arrived for stage s+1 but not yet read
  // the partials of stage s, so a slot is only rewritten four stages later (behind the wait on stage s+2's barrier)
  const uint32_t pp = rg.fq & 1u, slot = rg.fq & 3u;
  const int nr = FULL ? SK_ROWS : min(SK_ROWS, c.row1 - (c.row0 + idx * SK_ROWS));
#if SK_THREADS == 256
  const uint32_t sb = c.a_stage + rg.buf * c.stage_bytes;
#else
  const float* sb = c.stage_buf + (size_t)rg.buf * SK_ROWS * N;
#endif
  float mrow[SK_ROWS], rsum[SK_ROWS];
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k)
#if SK_THREADS == 256
    mrow[k] = (FULL || k < nr) ? (c.c_nu - sk_lds(c.a_uold + (uint32_t)(idx * SK_ROWS + k) * 4u)) : 0.f;   // previous u
#else
    mrow[k] = (FULL || k < nr) ? (c.c_nu - c.uold_s[idx * SK_ROWS + k]) : 0.f;   // previous u
#endif
  SK_STAMP(rg.fq, 0);
  sk_wait_full(c, rg);
  SK_STAMP(rg.fq, 1);
  const float2 l2e = make_float2(LOG2E, LOG2E);
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) {
    float2 s2 = make_float2(0.f, 0.f);
    const float2 nm = make_float2(-mrow[k], -mrow[k]);
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int gi = g * SK_THREADS + tid;
      if (FULL || (k < nr && gi < n4)) {
#if SK_THREADS == 256
        const float4 t = sk_lds4(sb + (uint32_t)k * c.pitch_bytes + (uint32_t)(g * SK_THREADS * 16));
#else
        const float4 t = reinterpret_cast<const float4*>(sb + (size_t)k * N)[gi];
#endif
        const float2 a01 = __ffma2_rn(make_float2(t.x, t.y), l2e, __fadd2_rn(vl[g][0], nm));
        const float2 a23 = __ffma2_rn(make_float2(t.z, t.w), l2e, __fadd2_rn(vl[g][1], nm));
        e_cur[k][g][0] = sk_ex2(a01.x); e_cur[k][g][1] = sk_ex2(a01.y);
        e_cur[k][g][2] = sk_ex2(a23.x); e_cur[k][g][3] = sk_ex2(a23.y);
        s2 = __fadd2_rn(s2, make_float2(e_cur[k][g][0], e_cur[k][g][1]));
        s2 = __fadd2_rn(s2, make_float2(e_cur[k][g][2], e_cur[k][g][3]));
      } else {
        e_cur[k][g][0] = e_cur[k][g][1] = e_cur[k][g][2] = e_cur[k][g][3] = 0.f;
      }
    }
    rsum[k] = s2.x + s2.y;
  }
  SK_STAMP(rg.fq, 2);
  // transposed warp reduction of the two row sums: lanes 0-15 end up with row 0, lanes 16-31 with row 1
  const bool hi = (lane & 16) != 0;
  float keep = hi ? rsum[1] : rsum[0];
  const float send = hi ? rsum[0] : rsum[1];
  keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
#if SK_THREADS == 256
  if ((lane & 15) == 0) sk_sts(c.a_part_st + slot * (uint32_t)(SK_ROWS * SK_WARPS * 4), keep);
#else
  if ((lane & 15) == 0) c.part_s[slot][lane >> 4][warp] = keep;
#endif
  __syncwarp();
#if SK_THREADS == 256
  if (lane == 0) sk_mbar_arrive_a(c.a_bpart + pp * 8u);              // release: this warp's partials, and its reads of the buffer
#else
  if (lane == 0) sk_mbar_arrive(&c.bpart[pp]);                       // release: this warp's partials, and its reads of the buffer
#endif
  SK_STAMP(rg.fq, 3);
  if (HAVE_PREV) sk_col_accum<PREV_FULL>(c, idx_prev, rg.seq - 1, rg.fq - 1, rg, e_prev, cs, csN);
  SK_STAMP(rg.fq, 5);
  rg.advance();
  ++rg.fq;
}

// One pass over the band with the potentials vl (log2 domain) of this thread's columns.  Results stay in registers: cs = column
// sums relative to the stabiliser m_j = c_mu - v_j, csN = the dustbin column's (valid in lanes 0 / 16 of warp 0: rows 0 / 1 of
// every stage).  The shared-memory buffer of the band's LAST stage is not refilled here: the caller stages the reduction of the
// column sums through it and refills it afterwards (stage rg.seq - 1 + SK_STAGES).
__device__ __forceinline__ void sk_band_fast(const SkCtx& c, SkRing& rg, const float2 (&vl)[SK_GROUPS][2], float2 (&cs)[SK_GROUPS][2],
                                             float& csN) {
  const int nst = c.nst;
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) cs[g][0] = cs[g][1] = make_float2(0.f, 0.f);
  csN = 0.f;
  const bool rev = ((rg.seq / (uint32_t)nst) & 1u) != 0;
  // everything but (possibly) the band's last row block is complete: the ragged block — the first stage of a backward
  // pass, the last of a forward one — takes the predicated instantiation
  const bool all_full = (c.NS == SK_MAXN) && (c.row0 + nst * SK_ROWS == c.row1);
  const int ragged = all_full ? -1 : ((c.NS == SK_MAXN) ? nst - 1 : -2);                // -2: every stage is predicated
  float eA[SK_ROWS][SK_GROUPS][4], eB[SK_ROWS][SK_GROUPS][4];      // registers rotate between the two (loop unrolled by 2)
#define SK_IDX(st) (rev ? nst - 1 - (st) : (st))
#define SK_ISFULL(idx) (ragged == -1 || (ragged >= 0 && (idx) != ragged))
#define SK_STAGE(st, cur, prev)                                                                              \
  {                                                                                                          \
    const int i_ = SK_IDX(st), ip_ = SK_IDX((st) - 1);                                                       \
    const bool f_ = SK_ISFULL(i_), fp_ = SK_ISFULL(ip_);                                                     \
    if (f_ && fp_) sk_stage_fast<true, true, true>(c, i_, ip_, rg, vl, cs, csN, cur, prev);                 \
    else if (f_) sk_stage_fast<true, true, false>(c, i_, ip_, rg, vl, cs, csN, cur, prev);                  \
    else sk_stage_fast<false, true, false>(c, i_, ip_, rg, vl, cs, csN, cur, prev);                         \
  }
  int st = 1;
  if (SK_ISFULL(SK_IDX(0))) sk_stage_fast<true, false, false>(c, SK_IDX(0), 0, rg, vl, cs, csN, eA, eB);
  else sk_stage_fast<false, false, false>(c, SK_IDX(0), 0, rg, vl, cs, csN, eA, eB);
#pragma unroll 1
  for (; st + 1 < nst; st += 2) {
    SK_STAGE(st, eB, eA);
    SK_STAGE(st + 1, eA, eB);
  }
  if (st < nst) {
    SK_STAGE(st, eB, eA);
    sk_col_accum<false>(c, SK_IDX(nst - 1), rg.seq - 1, rg.fq - 1, rg, eB, cs, csN, false);
  } else {
    sk_col_accum<false>(c, SK_IDX(nst - 1), rg.seq - 1, rg.fq - 1, rg, eA, cs, csN, false);
  }
#undef SK_STAGE
#undef SK_ISFULL
#undef SK_IDX
}

__global__ void __launch_bounds__(SK_THREADS, 1) sinkhorn_fused_kernel(const float* __restrict__ S, int M, int N, int ld, float alpha,
                                                                       int iters, float* u, float* v, float* pm, float* ps,
                                                                       int* flag, unsigned long long* acc_base, int fx_shift, int rows_per_cta, int allow_fast,
                                                                       int keep_pct, int pf_stages) {
  extern __shared__ __align__(128) unsigned char sk_smem[];
  float* stage_buf = reinterpret_cast<float*>(sk_smem);                         // [SK_STAGES][SK_ROWS][N]
  __shared__ __align__(8) uint64_t full[SK_STAGES], bpart[2];
  __shared__ __align__(8) float part_m[2][SK_ROWS][SK_WARPS], part_s[4][SK_ROWS][SK_WARPS];
  __shared__ float red_m[SK_WARPS], red_s[SK_WARPS];
  __shared__ __align__(8) float uold_s[2][SK_MAX_BAND];           // u of the band, log2 domain: previous / this iteration
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int row0 = min(M, cta * rows_per_cta), row1 = min(M, row0 + rows_per_cta);
  const int nst = (row1 - row0 + SK_ROWS - 1) / SK_ROWS;                       // stages per iteration for this CTA (>= 1)
  const float norm = -logf((float)M + (float)N);
  const float log_mu_last = logf((float)N) + norm, log_nu_last = logf((float)M) + norm;
  const float c_mu = fmaxf(norm, log_mu_last) * LOG2E, c_nu = fmaxf(norm, log_nu_last) * LOG2E;   // log2 of the largest marginals
  const int n4 = (N + 3) >> 2, NP = n4 * 4;      // columns [N, NP) hold -1e30 (filled by the launcher): they add 0 to every sum
  // Wide matrices (N >= 7/8 of the maximum) use the maximal smem row pitch: the bulk copies fill the first NP floats of a row,
  // the tail [NP, SK_MAXN) is set to -1e30 once, and every stage takes the straight-line (unpredicated) instantiation.
  const int NS = sk_smem_pitch(NP);

  if (tid == 0) {
    for (int s = 0; s < SK_STAGES; ++s) sk_mbar_init(&full[s], 1);
    for (int s = 0; s < 2; ++s) sk_mbar_init(&bpart[s], SK_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (NS != NP)
    for (int i = tid; i < SK_STAGES * SK_ROWS * (NS - NP); i += SK_THREADS)
      stage_buf[(size_t)(i / (NS - NP)) * NS + NP + i % (NS - NP)] = -1e30f;
  __syncthreads();

  SkCtx ctx;
  ctx.S = S; ctx.stage_buf = stage_buf; ctx.full = full; ctx.bpart = bpart;
  ctx.part_m = part_m; ctx.part_s = part_s; ctx.u = u; ctx.v = v;
  ctx.pm = pm; ctx.ps = ps; ctx.flag = flag; ctx.M = M; ctx.N = N; ctx.NP = NP; ctx.NS = NS; ctx.ld = ld; ctx.n4 = n4; ctx.row0 = row0; ctx.row1 = row1;
  ctx.uold_s = uold_s[0]; ctx.unew_s = uold_s[1];
  ctx.nst = nst; ctx.cta = cta; ctx.norm = norm; ctx.c_mu = c_mu; ctx.c_nu = c_nu; ctx.row_bytes = (uint32_t)NP * 4u;
  ctx.keep_rows = (row1 - row0) * keep_pct / 100;
  ctx.tid = tid; ctx.warp = warp; ctx.lane = lane;
  ctx.pf = pf_stages;
  ctx.total = (uint32_t)iters * (uint32_t)nst;
  ctx.kfac = exp2f(norm * LOG2E - c_mu);                             // a_i = 2^(u_i + m_i - c_mu) = kfac / rowsum_i
#if SK_THREADS == 256
  ctx.a_stage = sk_pin(sk_saddr(stage_buf) + (uint32_t)tid * 16u);
  ctx.stage_bytes = sk_pin((uint32_t)SK_ROWS * (uint32_t)NS * 4u); ctx.pitch_bytes = sk_pin((uint32_t)NS * 4u);
  ctx.a_full = sk_pin(sk_saddr(&full[0])); ctx.a_bpart = sk_pin(sk_saddr(&bpart[0]));
  ctx.a_part_st = sk_pin(sk_saddr(&part_s[0][lane >> 4][warp]));
  ctx.a_part_ld = sk_pin(sk_saddr(&part_s[0][lane >> 4][(lane & (SK_WARPS / 2 - 1)) * 2]));
  ctx.a_uold = sk_saddr(uold_s[0]); ctx.a_unew = sk_saddr(uold_s[1]);
#endif

  // the band is re-streamed every iteration: stages are counted across iterations
  SkRing rg;
  rg.seq = 0; rg.buf = 0; rg.full_par = 0; rg.fq = 0; rg.bp_par = 0;
  rg.total = (uint32_t)iters * (uint32_t)nst;
  if (tid == 0)
    for (uint32_t s = 0; s < rg.total && s < SK_STAGES; ++s) sk_issue(ctx, s);
  bool fast_ok = allow_fast != 0;

  unsigned int* gbar = reinterpret_cast<unsigned int*>(flag) + 1;      // monotonic arrival counter of the grid barrier (zeroed by the host)
  unsigned int gtarget = 0;
  // Grid barrier: one release-add + relaxed polling per CTA.  Cumulativity through the two block barriers makes every
  // thread's earlier writes visible to every thread of the grid afterwards (readers use ld.global.cg).
  auto grid_barrier = [&]() {
    __syncthreads();
    gtarget += (unsigned int)G;
    if (tid == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gbar) : "memory");
      unsigned int seen;
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(gbar) : "memory");
      } while (seen < gtarget);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
  };

  // ---- state of the fast iterations (registers / shared memory, never through HBM) ----
  float2 vl[SK_GROUPS][2];               // v of my 16 columns, log2 domain
  float vN_l = 0.f, uM_l = 0.f;          // dustbin column / dustbin row potentials, log2 domain
  bool v_regs = false;                   // vl / vN_l / uM_l / uold are current (else: reload from u, v in global memory)
  uint32_t fit = 0;                      // fast iterations so far: sums of iteration f go through acc[f % 3]
  float* uold = uold_s[0];
  float* unew = uold_s[1];
  const float alpha_l = alpha * LOG2E, norm_l = norm * LOG2E;
  const float fx_up = __int_as_float((127 + fx_shift) << 23), fx_dn = __int_as_float((127 - fx_shift) << 23);   // 2^shift, 2^-shift
  const size_t acc_stride = (size_t)NP + 2;                           // u64 per accumulation buffer: NP columns, [NP] = dustbin column

  for (int it = 0; it < iters; ++it) {
    const bool fast = fast_ok && it >= SK_EXACT_ITERS;
    if (fast) {
      // ================= fast iteration: ONE grid barrier =================
      // Column sums leave the CTA as 64-bit fixed-point numbers and are added up by the L2 (one bulk reduction per CTA): integer
      // addition is associative, so the result does not depend on the order in which the CTAs arrive (bit-reproducible), and
      // after the barrier EVERY CTA derives the new v of its threads' columns itself — no combine phase, no second barrier.
      float uM_prev;
      if (!v_regs) {
        // first fast iteration (or after exact ones): potentials come from global memory
#pragma unroll
        for (int g = 0; g < SK_GROUPS; ++g) {
          const int gi = g * SK_THREADS + tid;
          const float4 t = (gi < n4) ? __ldcg(reinterpret_cast<const float4*>(v) + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
          vl[g][0] = make_float2(t.x * LOG2E, t.y * LOG2E); vl[g][1] = make_float2(t.z * LOG2E, t.w * LOG2E);
        }
        vN_l = __ldcg(v + N) * LOG2E;
        uM_prev = __ldcg(u + M) * LOG2E;
        for (int i = tid; i < row1 - row0; i += SK_THREADS) uold[i] = __ldcg(u + row0 + i) * LOG2E;
      } else {
        uM_prev = uM_l;
      }
      // dustbin row: u_M = log_mu_last - LSE_{j <= N}(alpha + v_j).  alpha + v_j <= m_M = c_nu - u_M(previous) (same a-priori
      // bound as for the other rows), so plain sums of 2^(alpha + v_j - m_M) are safe; every CTA computes it (same order: same bits)
      const float m_M = c_nu - uM_prev;
      {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < SK_GROUPS; ++g) {
          const int j = (g * SK_THREADS + tid) * 4;
          if (j + 0 < N) t += sk_ex2(alpha_l + vl[g][0].x - m_M);
          if (j + 1 < N) t += sk_ex2(alpha_l + vl[g][0].y - m_M);
          if (j + 2 < N) t += sk_ex2(alpha_l + vl[g][1].x - m_M);
          if (j + 3 < N) t += sk_ex2(alpha_l + vl[g][1].y - m_M);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red_s[warp] = t;
      }
      ctx.extra_row = alpha_l + vN_l;                                 // dustbin column term of every row sum
      ctx.uold_s = uold; ctx.unew_s = unew;
#if SK_THREADS == 256
      ctx.a_uold = sk_pin(sk_saddr(uold)); ctx.a_unew = sk_pin(sk_saddr(unew));
#endif
      __syncthreads();
      bool bad = false;
      {
        float t = sk_ex2(alpha_l + vN_l - m_M);
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) t += red_s[w];
        bad = !(t > 0.f && t < INFINITY);
        uM_l = log_mu_last * LOG2E - (m_M + __log2f(t));
      }
      float2 cs[SK_GROUPS][2];
      float csN;
      SK_BSTAMP(fit, 0);
      sk_band_fast(ctx, rg, vl, cs, csN);
      SK_BSTAMP(fit, 1);
      // ---- column sums -> fixed point -> the shared-memory buffer of the band's last stage (free: every warp has passed the wait
      //      for that stage's partials, i.e. every warp has taken its elements out of it) -> one bulk reduction into acc ----
      unsigned long long* acc = acc_base + (size_t)(fit % 3u) * acc_stride;
      {
        const uint32_t last_buf = (rg.buf == 0 ? SK_STAGES : rg.buf) - 1u;
        // The staging area is the DATA part of the buffer's two rows (16 n4 bytes each = half of the 32 n4 bytes of sums): the
        // -1e30 tails [NP, NS) of the rows must survive.  16-byte unit w (two columns) goes to row w / n4, offset (w % n4) * 16.
        unsigned char* stg0 = reinterpret_cast<unsigned char*>(stage_buf + (size_t)last_buf * SK_ROWS * NS);
        unsigned char* stg1 = stg0 + (size_t)NS * 4;
#pragma unroll
        for (int g = 0; g < SK_GROUPS; ++g) {
          const int gi = g * SK_THREADS + tid;
          if (gi < n4) {
            const int w0 = 2 * gi, w1 = 2 * gi + 1;
            *reinterpret_cast<ulonglong2*>(w0 < n4 ? stg0 + (size_t)w0 * 16 : stg1 + (size_t)(w0 - n4) * 16) =
                make_ulonglong2(__float2ull_rn(cs[g][0].x * fx_up), __float2ull_rn(cs[g][0].y * fx_up));
            *reinterpret_cast<ulonglong2*>(w1 < n4 ? stg0 + (size_t)w1 * 16 : stg1 + (size_t)(w1 - n4) * 16) =
                make_ulonglong2(__float2ull_rn(cs[g][1].x * fx_up), __float2ull_rn(cs[g][1].y * fx_up));
          }
        }
        asm volatile("fence.proxy.async;" ::: "memory");             // my generic-proxy writes (shared and global) before the bulk engine's
        const float csN1 = __shfl_sync(0xffffffffu, csN, 16);
        __syncthreads();
        SK_BSTAMP(fit, 2);
        if (tid == 0) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                       ::"l"(acc), "r"((uint32_t)__cvta_generic_to_shared(stg0)), "r"((uint32_t)n4 * 16u) : "memory");
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                       ::"l"(acc + 2 * (size_t)n4), "r"((uint32_t)__cvta_generic_to_shared(stg1)), "r"((uint32_t)n4 * 16u) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          atomicAdd(acc + NP, __float2ull_rn((csN + csN1) * fx_up));  // dustbin column: one more 64-bit integer add
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging buffer read: hand it back to the ring
          if (rg.seq - 1u + SK_STAGES < rg.total) sk_issue(ctx, rg.seq - 1u + SK_STAGES);
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the reduction is performed before this CTA arrives
        }
      }
      SK_BSTAMP(fit, 3);
      grid_barrier();
      SK_BSTAMP(fit, 4);
      // ---- every CTA: new v of my columns from the reduced sums; recycle the buffer of two iterations ago ----
      {
        unsigned long long* old = acc_base + (size_t)((fit + 2u) % 3u) * acc_stride;
        for (size_t i = (size_t)cta * SK_THREADS + tid; i < acc_stride; i += (size_t)G * SK_THREADS) old[i] = 0ull;
      }
      const float extra_col = alpha_l + uM_l;
#pragma unroll
      for (int g = 0; g < SK_GROUPS; ++g) {
        const int gi = g * SK_THREADS + tid;
        if (gi < n4) {
          const ulonglong2 s01 = __ldcg(reinterpret_cast<const ulonglong2*>(acc) + 2 * gi);
          const ulonglong2 s23 = __ldcg(reinterpret_cast<const ulonglong2*>(acc) + 2 * gi + 1);
          const float fs[4] = {__ull2float_rn(s01.x) * fx_dn, __ull2float_rn(s01.y) * fx_dn, __ull2float_rn(s23.x) * fx_dn,
                               __ull2float_rn(s23.y) * fx_dn};
          float vo[4] = {vl[g][0].x, vl[g][0].y, vl[g][1].x, vl[g][1].y};
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float mj = c_mu - vo[cc];                            // the stabiliser used in the band (old v)
            const float sj = fs[cc] + sk_ex2(extra_col - mj);          // + dustbin row
            if (gi * 4 + cc < N && !(sj > 0.f && sj < INFINITY)) bad = true;
            vo[cc] = norm_l - (mj + __log2f(sj));
          }
          vl[g][0] = make_float2(vo[0], vo[1]); vl[g][1] = make_float2(vo[2], vo[3]);
        }
      }
      {
        const float mN = c_mu - vN_l;
        const float sN = __ull2float_rn(__ldcg(acc + NP)) * fx_dn + sk_ex2(extra_col - mN);
        if (!(sN > 0.f && sN < INFINITY)) bad = true;
        vN_l = log_nu_last * LOG2E - (mN + __log2f(sN));
      }
      { float* t = uold; uold = unew; unew = t; }
      v_regs = true;
      SK_BSTAMP(fit, 5);
      ++fit;
      if (__syncthreads_or((bad || __ldcg(flag) != 0) ? 1 : 0)) {
        // the fast mode tripped (potential jump beyond the f32 exponent range): restart the whole solve in exact mode.  Every
        // CTA sees the same sums and the same flag, so all of them take this branch together.
        fast_ok = false; v_regs = false;
        for (int j = cta * SK_THREADS + tid; j <= N; j += G * SK_THREADS) v[j] = 0.f;
        for (int i = cta * SK_THREADS + tid; i <= M; i += G * SK_THREADS) u[i] = 0.f;
        const uint32_t old_total = rg.total;
        rg.total = rg.seq + (uint32_t)iters * (uint32_t)nst;
        ctx.total = rg.total;
        if (tid == 0)      // stages below min(old_total, seq + SK_STAGES) are already in flight
          for (uint32_t sq = min(old_total, rg.seq + SK_STAGES); sq < rg.total && sq < rg.seq + SK_STAGES; ++sq) sk_issue(ctx, sq);
        it = -1;
        grid_barrier();
      }
      continue;
    }
    // ================= exact iteration: running maxima, column partials in global memory, two grid barriers =================
    v_regs = false;
    float4 vraw[SK_GROUPS];
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int gi = g * SK_THREADS + tid;
      vraw[g] = (gi < n4) ? __ldcg(reinterpret_cast<const float4*>(v) + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ctx.extra_row = (alpha + __ldcg(v + N)) * LOG2E;                  // dustbin column term of every row LSE (old v)
    // dustbin row: u[M] = log_mu_last - LSE_j(alpha + v_j), j in [0, N]   (last CTA: its band is the short one)
    if (cta == G - 1) {
      L2Acc a; a.init();
      for (int j0 = tid; j0 <= N; j0 += 16 * SK_THREADS) {
        float vv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int j = j0 + k * SK_THREADS; vv[k] = (j <= N) ? __ldcg(v + j) : -INFINITY; }
#pragma unroll
        for (int k = 0; k < 16; ++k) if (vv[k] != -INFINITY) a.add((alpha + vv[k]) * LOG2E);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
      if (lane == 0) { red_m[warp] = a.m; red_s[warp] = a.s; }
      __syncthreads();
      if (tid == 0) {
        L2Acc t; t.init();
        for (int w = 0; w < SK_WARPS; ++w) t.merge(red_m[w], red_s[w]);
        u[M] = log_mu_last - t.lse_ln();
      }
    }
    __syncthreads();
    sk_band_exact(ctx, rg, vraw);
    grid_barrier();
    // ---- combine: v_j = log_nu_j - LSE_i(S_ij + u_i) incl. the dustbin row; v[N] from all u ----
    {
      const float extra_col = (alpha + __ldcg(u + M)) * LOG2E;
      // one warp per 4-column tile, lane = 4 * sub + column: 8 lanes share a column
      for (int tile = cta * SK_WARPS + warp; tile < n4; tile += G * SK_WARPS) {
        const int sub = lane >> 2, j = tile * 4 + (lane & 3);
        L2Acc a; a.init();
#pragma unroll 2
        for (int c = sub; c < G; c += 8) a.merge(__ldcg(pm + (size_t)c * NP + j), __ldcg(ps + (size_t)c * NP + j));
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
        if (sub == 0 && j < N) {
          a.add(extra_col);
          v[j] = norm - a.lse_ln();
        }
      }
      if (cta == G - 1) {                                             // v[N] = log_nu_last - LSE_{i <= M}(alpha + u_i)
        L2Acc a; a.init();
        for (int i0 = tid; i0 <= M; i0 += 16 * SK_THREADS) {
          float uu[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) { const int i = i0 + k * SK_THREADS; uu[k] = (i <= M) ? __ldcg(u + i) : -INFINITY; }
#pragma unroll
          for (int k = 0; k < 16; ++k) if (uu[k] != -INFINITY) a.add((alpha + uu[k]) * LOG2E);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
        if (lane == 0) { red_m[warp] = a.m; red_s[warp] = a.s; }
        __syncthreads();
        if (warp == 0) {
          L2Acc t; t.init();
          if (lane < SK_WARPS) { t.m = red_m[lane]; t.s = red_s[lane]; }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t.merge(__shfl_xor_sync(0xffffffffu, t.m, o), __shfl_xor_sync(0xffffffffu, t.s, o));
          if (lane == 0) v[N] = log_nu_last - t.lse_ln();
        }
      }
    }
    grid_barrier();
  }
  // the potentials of the last fast iteration are still in registers: CTA 0 writes them out (u_i, i < M, went out row by row)
  if (v_regs && cta == 0) {
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int j = (g * SK_THREADS + tid) * 4;
      if (j + 0 < N) v[j + 0] = vl[g][0].x * LN2;
      if (j + 1 < N) v[j + 1] = vl[g][0].y * LN2;
      if (j + 2 < N) v[j + 2] = vl[g][1].x * LN2;
      if (j + 3 < N) v[j + 3] = vl[g][1].y * LN2;
    }
    if (tid == 0) { v[N] = vN_l * LN2; u[M] = uM_l * LN2; }
  }
}

// the row pitch must hold whole float4 groups (bulk copies are 16-byte granular); N itself may be anything in [64, SK_MAXN]
static bool sinkhorn_fused_ok(const float* S, int M, int N, int ld) {
  return (ld & 3) == 0 && ld >= ((N + 3) & ~3) && (reinterpret_cast<uintptr_t>(S) & 15) == 0 && N <= SK_MAXN && N >= 64 && M >= 1;
}
// columns [N, round_up(N, 4)) of every row <- -1e30: the fused kernel then treats them as ordinary columns of weight 0
__global__ void sk_pad_fill_kernel(float* S, int M, int N, int ld) {
  const int npad = ((N + 3) & ~3) - N;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * npad) return;
  S[(size_t)(i / npad) * ld + N + i % npad] = -1e30f;
}

// returns 0 when the fused kernel ran, 1 when the shape is unsupported (caller falls back to the two-pass kernels)
static int sinkhorn_fused_launch(float* S, int M, int N, int ld, float alpha, int iters, float* u, float* v, AssignWs& w,
                                 cudaStream_t st) {
  if (!sinkhorn_fused_ok(S, M, N, ld) || iters <= 0) return 1;
  static int coop = -1, sms = 0;   // B200 boxes are homogeneous: queried once
  if (coop < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!coop || sms <= 0 || sms > 256) return 1;
  const int NP = (N + 3) & ~3;
  const size_t smem = (size_t)SK_STAGES * SK_ROWS * sk_smem_pitch(NP) * sizeof(float);
  static bool attr_seen[64] = {};
  if (i4d_first_use_on_device(attr_seen)) {
    if (cudaFuncSetAttribute(sinkhorn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SK_STAGES * SK_ROWS * SK_MAXN * (int)sizeof(float)) != cudaSuccess) { cudaGetLastError(); return 1; }
  }
  int G = sms;
  {   // I4D_SK_G: run on fewer CTAs (experiments: what a band pass costs when two problems share the GPU)
    static int g_env = -1;
    if (g_env < 0) { const char* e = getenv("I4D_SK_G"); g_env = e ? atoi(e) : 0; }
    if (g_env > 0 && g_env < G) G = g_env;
  }
  int rpc = (M + G - 1) / G;
  rpc = (rpc + SK_ROWS - 1) / SK_ROWS * SK_ROWS;
  if (rpc > SK_MAX_BAND) return 1;
  G = (M + rpc - 1) / rpc;                      // CTAs that actually own rows (<= sms <= 256 partial slots)
  cudaMemsetAsync(u, 0, (size_t)(M + 1) * sizeof(float), st);
  cudaMemsetAsync(v, 0, (size_t)(NP + 1) * sizeof(float), st);       // v[N+1 .. NP] are read (never written) as the pad columns' potentials
  if (NP != N) sk_pad_fill_kernel<<<i4d_cdiv(M * (NP - N), 256), 256, 0, st>>>(S, M, N, ld);
  float* pm = w.pm; float* ps = w.ps;
  int* flag = w.pi;
  cudaMemsetAsync(flag, 0, 2 * sizeof(int), st);   // [0] fast-mode trip flag, [1] grid-barrier arrival counter
  // three rotating accumulation buffers of (NP + 2) 64-bit fixed-point column sums (the argmax-index partials of the workspace
  // are free during the solve).  Column sums relative to their stabilisers are bounded by M / N (total row mass over the
  // largest column marginal): scale 2^shift with (M / N) * 2^shift < 2^62.
  unsigned long long* acc_base = reinterpret_cast<unsigned long long*>(w.pi + 4);   // (the first 16 bytes hold flag / barrier counter)
  cudaMemsetAsync(acc_base, 0, 3 * ((size_t)NP + 2) * sizeof(unsigned long long), st);
  int fx_shift = 61;
  for (long long r = 1; r * (long long)N < (long long)M; r *= 2) --fx_shift;
  int allow_fast = g_sinkhorn_fast;
  static int keep_pct = -1;      // share of every band pinned in L2 with evict_last (I4D_SK_KEEP_PCT overrides, for experiments)
  if (keep_pct < 0) {
    const char* e = getenv("I4D_SK_KEEP_PCT");
    keep_pct = e ? atoi(e) : SK_KEEP_PCT_DEFAULT;
    if (keep_pct < 0 || keep_pct > 100) keep_pct = SK_KEEP_PCT_DEFAULT;
  }
  // L2 prefetch distance in stages (cp.async.bulk.prefetch.L2 of a later stage with every shared-memory refill): 1 measured
  // best at 8192^2 (48.8 -> 46.3 us per iteration; 2: 47.3, 3: 48.4).  I4D_SK_PF overrides, for experiments.
  static int pf_stages = -1;
  if (pf_stages < 0) { const char* e = getenv("I4D_SK_PF"); pf_stages = e ? atoi(e) : 1; if (pf_stages < 0 || pf_stages > 16) pf_stages = 1; }
  void* args[] = {(void*)&S, (void*)&M, (void*)&N, (void*)&ld, (void*)&alpha, (void*)&iters, (void*)&u, (void*)&v, (void*)&pm, (void*)&ps,
                  (void*)&flag, (void*)&acc_base, (void*)&fx_shift, (void*)&rpc, (void*)&allow_fast, (void*)&keep_pct, (void*)&pf_stages};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)sinkhorn_fused_kernel, dim3(G), dim3(SK_THREADS), args, smem, st);
  if (e != cudaSuccess) { cudaGetLastError(); return 1; }
  return 0;
}

#undef SK_THREADS
#undef SK_WARPS
#undef SK_ROWS
#undef SK_STAGES
#undef SK_MAXN
#undef SK_GROUPS
#undef SK_EXACT_ITERS
#undef SK_MAX_BAND
#undef SK_KEEP_PCT_DEFAULT
#undef SK_STAMP
#undef SK_BSTAMP
#ifdef SK_TRACE_MAX
#undef SK_TRACE_MAX
#endif
