#!/usr/bin/env python
"""Benchmark of the per-epoch stereo hot path (BASELINE.json: "stereo epochs/sec (6000x4000 tiled SP+LG)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg5|cfg1]

One step = one stereo epoch of the named BASELINE.json config on a synthetic 6000x4000 pair:
  cfg2 (default, configs[1], the config the metric is quoted on): 2x3 GRID tiles (1999x1999), SuperPoint (8192 kp/tile) +
       SuperGlue (outdoor arch, 100 Sinkhorn iterations), tile merge, MAGSAC++ F verification, iterative-LS triangulation;
  cfg5 (configs[4]): 3x4 GRID tiles (1499x1329), SuperPoint (16384 kp/tile) + LightGlue (static depth/width), dual-softmax +
       mutual NN, verification, triangulation;
  cfg1 (configs[0]): Quality.LOW (1500x1000), LightGlue 2048 kp, one tile.
Epochs are independent: `shard_epochs` gives rank r of N its epochs (weak scaling, no data-path collective); after the timed
region the per-epoch results are exchanged once with `gather_results` (NCCL) and that time is reported separately.

Output: ONE JSON line (rank 0).  `value` = epochs/s with the u8 images already resident in HBM; `e2e` = the same through the
public plugin API (`*Matcher.match` + `Triangulate.triangulate_two_views`) from pinned host images to host arrays;
`roofline` = the dominant HBM-bound kernel and `roofline_tensor` = the attention kernel, both timed live with CUDA events;
`cpu_baseline` = the reference's own CPU implementation (oracle/_ref, staged by oracle/stage_ref.py; the CPU port in oracle/ if
that is absent) on this box's host cores over a bounded sample.  `--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

UNIT = "epochs/s"
H, W = 4000, 6000
CONFIGS = {
    "cfg2": {"metric": "stereo epochs/sec (6000x4000, 2x3 tiles, SuperPoint+SuperGlue 8192 kp/tile, 100 Sinkhorn iters)",
             "workload": "cfg2: 6000x4000 stereo pair, 2x3 GRID tiles (1999x1999), SuperPoint+SuperGlue outdoor arch, 8192 kp/tile, "
                         "100 Sinkhorn iters, MAGSAC++ F verification, iterative-LS triangulation",
             "grid": (2, 3), "kp": 8192, "pairs": 6, "tile": (2000, 2000), "matcher": "superglue"},
    "cfg5": {"metric": "stereo epochs/sec (6000x4000, 3x4 tiles, SuperPoint+LightGlue 16384 kp/tile, dual-softmax)",
             "workload": "cfg5: 6000x4000 stereo pair, 3x4 GRID tiles (1499x1329), SuperPoint+LightGlue (static depth/width), "
                         "16384 kp/tile, dual-softmax + mutual NN, MAGSAC++ F verification, iterative-LS triangulation",
             "grid": (3, 4), "kp": 16384, "pairs": 12, "tile": (1330, 1500), "matcher": "lightglue"},
    "cfg1": {"metric": "stereo epochs/sec (6000x4000 at Quality.LOW = 1500x1000, SuperPoint+LightGlue 2048 kp)",
             "workload": "cfg1: 6000x4000 stereo pair, Quality.LOW (2x pyrDown -> 1500x1000), one tile, SuperPoint+LightGlue 2048 kp, "
                         "MAGSAC++ F verification, iterative-LS triangulation",
             "grid": (1, 1), "kp": 2048, "pairs": 1, "tile": (1000, 1500), "matcher": "lightglue", "shift": (64, 32)},
}
SINKHORN_ITERS = 100


def _config_block(cfg_name: str):
    """The `config` object both arms print (identical for `--impl ours` and `--impl reference`)."""
    c = CONFIGS[cfg_name]
    return {"workload": c["workload"], "weights": "seeded structured random init (icepy4d_b200/weights.py)",
            "l2": "inputs_larger_than_l2 (2 x 72 MB images; 268 MB / 1.07 GB score matrices)"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU reference timing
def _reference_matcher(cfg_name: str):
    """The UNMODIFIED reference matcher class (staged under oracle/_ref or mounted at /root/reference) with the seeded weights,
    on CPU.  Returns (callable(image0, image1, **kw) -> n_matches, description) or None when no reference tree is available."""
    from oracle import ref_shims

    if not ref_shims.reference_available():
        return None
    from icepy4d_b200 import weights

    ref_shims.install_shims()
    import icepy4d.matching.matchers as M
    from icepy4d.matching import GeometricVerification, Quality, TileSelection

    c = CONFIGS[cfg_name]
    sp_sd = weights.make_superpoint_state(1)
    if c["matcher"] == "superglue":
        with ref_shims.no_checkpoint_loading():
            m = M.SuperGlueMatcher({"weights": "outdoor", "keypoint_threshold": 1e-4, "max_keypoints": c["kp"], "match_threshold": 0.2,
                                    "force_cpu": True, "sinkhorn_iterations": SINKHORN_ITERS})
        m.matcher.superpoint.load_state_dict(sp_sd)
        m.matcher.superglue.load_state_dict(weights.make_superglue_state(2))

        def run(i0, i1):
            m.match(i0, i1, quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[1, 1], overlap=0,
                    geometric_verification=GeometricVerification.MAGSAC)
            return len(m.mkpts0)
        return run, "icepy4d.matching.SuperGlueMatcher.match (reference source, CPU)"
    lg_sd = weights.make_lightglue_state(3)
    import icepy4d.thirdparty.LightGlue.lightglue as LGpkg
    # LightGlueMatcher re-instantiates both networks (and reloads checkpoints that do not exist here) on every call
    # (matchers.py:1256-1258): the constructors are pointed at the seeded weights; everything else is the reference's code
    LGpkg.SuperPoint = lambda **kw: ref_shims.build_reference_superpoint_lg(sp_sd, **kw)
    LGpkg.LightGlue = lambda features="superpoint", **kw: ref_shims.build_reference_lightglue(lg_sd, **kw)
    lm = M.LightGlueMatcher({"features": "superpoint", "force_cpu": True})

    def run_lg(i0, i1, kp):
        lm.match(i0, i1, quality=Quality.HIGH, tile_selection=TileSelection.GRID, grid=[1, 1], overlap=0, max_keypoints=kp,
                 geometric_verification=GeometricVerification.MAGSAC)
        return len(lm.mkpts0)
    return run_lg, "icepy4d.matching.LightGlueMatcher.match (reference source, CPU)"


def cpu_sample(cfg_name: str, threads: int):
    """Bounded sample of one epoch on the host cores, through the reference's own plugin call: ONE tile pair of the config
    (`match(tile0, tile1, GRID, grid=[1,1])` on a crop of the epoch's images: with grid [1,1] the reference's tiler yields exactly
    the (DX-1)x(DY-1) tile the config's grid yields), times the number of tile pairs.  cfg5's 16384-kp pair does not fit a
    bounded sample (minutes, > 20 GB on the CPU path): it is timed at 4096 kp and scaled by the N^2 attention/assignment work.
    Returns (seconds per epoch [extrapolated], seconds actually timed, kind, description)."""
    from icepy4d_b200 import synthetic

    torch.set_num_threads(threads)
    c = CONFIGS[cfg_name]
    th, tw = c["tile"]
    i0, i1 = synthetic.stereo_pair(th, tw, seed=1000, shift=(16, 24), channels=3)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        ref = _reference_matcher(cfg_name)
    kp, scale = c["kp"], 1.0
    if cfg_name == "cfg5":
        kp, scale = 4096, (c["kp"] / 4096.0) ** 2
    if ref is not None:
        import contextlib
        run, what = ref
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sys.stderr):        # the reference prints progress lines; stdout carries the JSON line only
            n = run(i0, i1) if c["matcher"] == "superglue" else run(i0, i1, kp)
        dt = time.perf_counter() - t0
        kind = "reference"
    else:                                   # the CPU port in oracle/ (same algorithm, restated; used only if oracle/_ref is missing)
        from icepy4d_b200 import weights
        from oracle import lg_oracle, sg_oracle, sp_oracle
        sp_sd = weights.make_superpoint_state(1)
        g = lambda im: torch.tensor(im[: th - 1, : tw - 1, 0] / 255.0, dtype=torch.float)[None, None]
        t0 = time.perf_counter()
        with torch.inference_mode():
            if c["matcher"] == "superglue":
                f0, f1 = sp_oracle.superpoint_sg(g(i0), sp_sd, 3, 1e-4, kp), sp_oracle.superpoint_sg(g(i1), sp_sd, 3, 1e-4, kp)
                out = sg_oracle.superglue(f0["keypoints"], f0["scores"], f0["descriptors"], f1["keypoints"], f1["scores"], f1["descriptors"],
                                          (th - 1, tw - 1), (th - 1, tw - 1), weights.make_superglue_state(2), iters=SINKHORN_ITERS, thr=0.2)
            else:
                f0, f1 = sp_oracle.superpoint_lg(g(i0), sp_sd, k=kp), sp_oracle.superpoint_lg(g(i1), sp_sd, k=kp)
                out = lg_oracle.lightglue(f0["keypoints"], f0["descriptors"], (tw - 1.0, th - 1.0), f1["keypoints"], f1["descriptors"],
                                          (tw - 1.0, th - 1.0), weights.make_lightglue_state(3), depth_conf=-1, width_conf=-1)
        n = int((out["matches0"] > -1).sum())
        dt = time.perf_counter() - t0
        kind, what = "port", "oracle/{sp,sg,lg}_oracle.py (CPU port; oracle/_ref not staged)"
    est = c["pairs"] * dt * scale
    desc = (f"{what}: 1 of {c['pairs']} tile pairs ({tw - 1}x{th - 1}, {kp} kp/tile, {n} verified matches) in {dt:.1f} s on {threads} threads"
            + (f", x{scale:.0f} for the N^2 work at {c['kp']} kp" if scale != 1.0 else "")
            + f", x{c['pairs']} tile pairs = one epoch (extrapolated; triangulation, < 1 % of the CPU time, not included)")
    return est, dt, kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # one sample costs tens of seconds of all host cores: the run stays within minutes by timing min(steps, 2) samples, no warm-up
    n_samples = max(1, min(args.steps, 2 if args.config == "cfg1" else 1))
    ests, dts, kind, desc = [], [], "port", ""
    for _ in range(n_samples):
        est, dt, kind, desc = cpu_sample(args.config, threads)
        ests.append(est); dts.append(dt)
    sec = float(np.median(ests))
    v = 1.0 / sec
    c = CONFIGS[args.config]
    line = {"impl": "reference", "metric": c["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": _config_block(args.config),
            "samples_timed": n_samples, "seconds_timed": float(sum(dts)), "extrapolated": True,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def _event_time(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ncu `--set full` captures of the final kernels (profiles/r2_ncu_*.txt): DRAM bytes per launch unit
# (r2, 256-thread instantiation: 3.421 GB read + 17.9 MB written over a 20-iteration launch, L2 hit rate 51 % incl. the L2 prefetches)
SINKHORN_TRAFFIC_PER_ITER = {"read": 171.0e6, "write": 0.89e6, "source": "profiles/r2_ncu_sinkhorn.txt"}
# ncu dram__bytes_read + dram__bytes_write of the two one-read passes of i4d_lg_assign at 16384^2 (profiles/r2_ncu_rowcol.txt):
# rowcol_lse_kernel 1.0743 GB + 12.5 MB, rowcol_argmax_kernel 1.0740 GB + 12.7 MB (the combine kernels move < 40 MB)
DUALSOFTMAX_TRAFFIC = {(16384, 16384): 1.0743e9 + 12.48e6 + 1.0740e9 + 12.73e6}


def _roofline_hbm(cfg_name: str, peaks):
    """The dominant HBM-bound kernel of the config, timed live with CUDA events on the launch stream.
    cfg2: the fused persistent Sinkhorn (600 iterations over 8192x8192 f32 per epoch).  SURVEY.md §8d counts 2*M*N*4 bytes per
    iteration (one read per LSE pass); the fused algorithm needs ONE read of the matrix per iteration, so `achieved` / `frac` are
    quoted on those one-read algorithmic bytes (M*N*4 per iteration: the floor of any kernel that streams the matrix once per
    iteration).  `traffic` is what ncu saw in DRAM (less than the algorithmic bytes: a third of the read is served by L2) with
    its own fraction `frac_actual_traffic`, and the §8d figure is kept as `algorithmic_2pass`.
    cfg5 / cfg1: the dual-softmax + mutual-NN assignment over the MxN similarity matrix (§8d: 2 reads of sim)."""
    from icepy4d_b200 import ops

    c = CONFIGS[cfg_name]
    M = N = c["kp"]
    S = torch.randn(M, N, device="cuda") * 3
    ws = ops.AssignWorkspace(M, N, S.device)
    if c["matcher"] == "superglue":
        iters = SINKHORN_ITERS
        ms = _event_time(lambda: ops.sinkhorn(S, 1.0, iters, ws))
        alg2 = 2.0 * M * N * 4 * iters
        one_read = 1.0 * M * N * 4 * iters
        tr = SINKHORN_TRAFFIC_PER_ITER
        traffic = (tr["read"] + tr["write"]) * iters if (M, N) == (8192, 8192) else None
        gbs = one_read / (ms * 1e-3) / 1e9
        out = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
               "traffic": traffic, "kernel": f"sinkhorn_fused_kernel ({iters} iterations, {M}x{N} f32, one launch)",
               "ms_per_launch": ms, "us_per_iteration": ms * 1e3 / iters, "peak_source": peaks["source"],
               "bytes_basis": f"one-read algorithmic bytes: M*N*4 per iteration x {iters} (the fused kernel reads the score matrix once "
                              "per iteration; SURVEY.md §8d's two-pass count is in algorithmic_2pass)",
               "algorithmic_bytes": one_read,
               "algorithmic_2pass": {"bytes_per_launch": alg2, "GB/s": alg2 / (ms * 1e-3) / 1e9,
                                     "frac": alg2 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                     "note": "SURVEY.md §8d model (2 reads of the matrix per iteration); the fused kernel reads it once"}}
        if traffic is not None:
            out["frac_actual_traffic"] = traffic / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]
            out["traffic_source"] = (f"ncu dram__bytes_read + dram__bytes_write per iteration x {iters} ({tr['source']}); below the "
                                     "algorithmic bytes because part of every pass is served by L2")
        return out
    z0 = torch.randn(M, device="cuda")
    z1 = torch.randn(N, device="cuda")
    ms = _event_time(lambda: ops.lg_assign(S, z0, z1, 0.1, ws))
    alg = 2.0 * M * N * 4
    gbs = alg / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": DUALSOFTMAX_TRAFFIC.get((M, N)), "kernel": f"i4d_lg_assign (dual softmax + mutual NN, {M}x{N} f32 similarity)",
            "ms_per_launch": ms, "peak_source": peaks["source"],
            "bytes_basis": "SURVEY.md §8d: 2 reads of the similarity matrix (row/column statistics, then row/column argmax)"}


def _roofline_tensor(cfg_name: str, peaks):
    """attn_tc_kernel (tcgen05 flash attention, head_dim 64, 4 heads) on the config's layer shape, timed live with CUDA events."""
    from icepy4d_b200 import ops_tc

    c = CONFIGS[cfg_name]
    n = c["kp"]
    qkv = (torch.randn(2 * n, 768, device="cuda") * 0.5).to(torch.bfloat16)
    out = torch.empty(2 * n, 256, device="cuda", dtype=torch.bfloat16)
    probs = [(0, n, 0, n), (n, n, n, n)]
    ms = _event_time(lambda: ops_tc.attention_tc(qkv, probs, out, 0, 256, 512), warm=3, reps=10)
    flops = 2 * 4.0 * n * n * 64 * 4
    tf = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"],
            "traffic": None, "kernel": f"attn_tc_kernel (2 images x 4 heads x {n}x{n}, head_dim 64, bf16 operands, f32 accumulate)",
            "ms_per_launch": ms, "flops_per_launch": flops, "peak_source": peaks["source"]}


def _make_pipeline(cfg_name: str, args):
    from icepy4d_b200 import epoch

    c = CONFIGS[cfg_name]
    if cfg_name == "cfg2":
        return epoch.make_cfg2_pipeline(c["kp"], SINKHORN_ITERS, precision=args.precision, conv_precision=args.conv_precision, grid=c["grid"])
    if cfg_name == "cfg5":
        return epoch.make_cfg5_pipeline(c["kp"], precision=args.precision, conv_precision=args.conv_precision, grid=c["grid"])
    return epoch.make_cfg1_pipeline(c["kp"], precision=args.precision, conv_precision=args.conv_precision)


def run_ours(args):
    import torch.distributed as dist

    from icepy4d_b200 import _native, synthetic
    from icepy4d_b200.epoch import gather_results_device, shard_epochs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = _peaks()
    cfg = CONFIGS[args.config]
    pipe = _make_pipeline(args.config, args)

    # the job: world * steps epochs, epoch e = synthetic pair with seed 1000 + e; rank r owns shard_epochs(...) (round-robin).
    # Images come from a small pool cycled over the rank's epochs (generating 72 MB of blurred noise per epoch on the host
    # would dominate the run); every rank has its own pool.
    my_epochs = shard_epochs(world * args.steps, rank, world)
    assert len(my_epochs) == args.steps
    pool = 2
    host = []
    for e in range(pool):
        # image 1 = image 0 shifted by a multiple of 8 px AT THE RESOLUTION THE DETECTOR SEES (cfg1 works at 1/4 resolution)
        i0, i1 = synthetic.stereo_pair(H, W, seed=1000 + rank * pool + e, shift=cfg.get("shift", (16, 24)), channels=3)
        host.append((torch.from_numpy(i0).pin_memory(), torch.from_numpy(i1).pin_memory()))
    dev = [(a.cuda(non_blocking=True), b.cuda(non_blocking=True)) for a, b in host]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- kernel-only: inputs resident in HBM ----
    for s in range(args.warmup):
        pipe.run_device(*dev[s % pool])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:                       # one nvidia-smi poller per box (eight of them perturb an 8-GPU run through the driver lock)
        sampler.start()
    n0 = _native.LAUNCHES
    store = torch.empty((args.steps * 65536, 3), device="cuda", dtype=torch.float64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # every epoch's 3-D points are kept for the end-of-run gather in ONE preallocated device buffer (holding the per-epoch
    # tensors themselves would pin blocks of the caching allocator: measured 2x outliers on some epochs)
    results, off = {}, 0
    for s, epoch_id in enumerate(my_epochs):
        r = pipe.run_device(*dev[s % pool])["points3d"]
        n_pts = r.shape[0]
        if off + n_pts > store.shape[0]:
            store = torch.cat([store, torch.empty_like(store)])
        store[off:off + n_pts].copy_(r)
        results[epoch_id] = (off, n_pts)
        off += n_pts
        del r
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _native.LAUNCHES - n0
    clocks = sampler.stop() if rank == 0 else None
    n_matches = int(results[my_epochs[-1]][1])
    results = {e: store[o:o + k] for e, (o, k) in results.items()}

    # ---- the single end-of-run exchange: every rank receives every epoch's 3-D points over NCCL ----
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record()
    allres = gather_results_device(results, world)
    g1.record()
    barrier()
    gather_ms = g0.elapsed_time(g1)
    gathered_epochs = len(allres)
    gathered_bytes = int(sum(t.numel() * t.element_size() for t in allres.values()))
    assert gathered_epochs == world * args.steps, (gathered_epochs, world, args.steps)
    del allres, results, store

    # ---- end to end through the public API: pinned host images -> host arrays ----
    e2e_steps = max(1, min(args.steps, 5))
    pipe.run(host[0][0].numpy(), host[0][1].numpy())
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for s in range(e2e_steps):
        r = pipe.run(host[s % pool][0].numpy(), host[s % pool][1].numpy())
        d2h = r.mkpts0.nbytes + r.mkpts1.nbytes + r.scores0.nbytes * 3 + r.points3d.nbytes + (0 if r.status is None else r.status.nbytes) + 72
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    t = torch.tensor([ms, e2e_ms, gather_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, gather_ms = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * args.steps / (ms * 1e-3)
    e2e_value = world * e2e_steps / (e2e_ms * 1e-3)
    roof = _roofline_hbm(args.config, peaks)
    roof_t = _roofline_tensor(args.config, peaks)
    threads = os.cpu_count() or 1
    cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "skipped (N > 1 or --no-cpu-baseline)"}
    if world == 1 and not args.no_cpu_baseline:
        est, dt, kind, desc = cpu_sample(args.config, threads)
        cpu = {"value": 1.0 / est, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc, "seconds_timed": dt}
    line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "f32" else "bf16 operands / f32 accumulate (matcher), split-f16 x3 / f32 accumulate (backbone), f32 elsewhere",
            "data": "synthetic", "config": _config_block(args.config),
            "details": {"precision": args.precision, "conv_precision": args.conv_precision, "verified_matches_last_epoch": n_matches,
                        "parallelism": f"{world * args.steps} epochs sharded round-robin over {world} GPU(s), no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * H * W * 3, "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
            "gather": {"ms": gather_ms, "epochs": gathered_epochs, "bytes": gathered_bytes,
                       "what": "gather_results_device: one count exchange + one padded all_gather of every epoch's 3-D points (NCCL), outside the timed region"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_tensor": roof_t, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    # bf16 = the tcgen05 tensor-core matcher (bf16 operands, f32 accumulation; tests/test_gpu_defaults.py); f32 = the SIMT f32
    # matcher kept as the on-device cross-check
    ap.add_argument("--precision", default="bf16", choices=["f32", "bf16"])
    ap.add_argument("--conv-precision", dest="conv_precision", default="f16x3", choices=["bf16x3", "f16x3", "f32", "tf32", "f16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
