"""CPU-only tests of the host side: tiler limits vs the reference's golden table, epoch sharding + the single result
gather with world_size 2 over gloo, header/ABI bookkeeping, weight preparation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tiler_limits_match_reference(golden_dir):
    from icepy4d_b200.matching.tiling import Tiler
    t = np.load(os.path.join(golden_dir, "tiler_quality.npz"))["tiler"]
    seen = 0
    for h, w, nr, nc, ov, ox, oy in {tuple(r[:7]) for r in t.tolist()}:
        lims, org = Tiler(grid=[nr, nc], overlap=ov, origin=[ox, oy]).compute_limits_by_shape(h, w)
        rows = [r for r in t.tolist() if tuple(r[:7]) == (h, w, nr, nc, ov, ox, oy)]
        assert len(rows) == len(lims)
        for r in rows:
            assert tuple(lims[r[7]]) == tuple(r[8:12])
            seen += 1
    assert seen == len(t)
    # cfg2 / cfg5 tile shapes quoted in SURVEY.md Appendix C
    lims, _ = Tiler(grid=[2, 3]).compute_limits_by_shape(4000, 6000)
    assert Tiler.patch_shape(lims[0], 4000, 6000) == (0, 0, 1999, 1999) and len(lims) == 6
    lims, _ = Tiler(grid=[3, 4]).compute_limits_by_shape(4000, 6000)
    assert Tiler.patch_shape(lims[5], 4000, 6000)[2:] == (1499, 1329) and len(lims) == 12


def test_enums_match_reference_values():
    from icepy4d_b200.matching import GeometricVerification, Quality, TileSelection
    assert [e.value for e in TileSelection] == [0, 1, 2, 3]
    assert [e.value for e in GeometricVerification] == [1, 2, 3]
    assert [e.value for e in Quality] == [1, 2, 3, 4]


def test_shard_epochs_partition():
    from icepy4d_b200.epoch import shard_epochs
    for n, g in ((256, 8), (10, 4), (3, 8), (0, 2)):
        parts = [shard_epochs(n, r, g) for r in range(g)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from icepy4d_b200.epoch import gather_results, gather_results_device, shard_epochs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_epochs(7, rank, world)
        local = {e: np.full((e + 1, 3), float(e), dtype=np.float64) for e in mine}       # ragged per-epoch results
        allr = gather_results(local, world)
        ok = sorted(allr) == list(range(7)) and all(allr[e].shape == (e + 1, 3) and float(allr[e][0, 0]) == e for e in allr)
        # the tensor path bench.py uses on GPUs (padded all_gather; here on CPU tensors over gloo), incl. an empty epoch
        tl = {e: torch.full(((e * 5) % 4, 3), float(e), dtype=torch.float64) for e in mine}
        tr = gather_results_device(tl, world)
        ok = ok and sorted(tr) == list(range(7)) and all(tr[e].shape == ((e * 5) % 4, 3) and bool((tr[e] == e).all()) for e in tr)
        t = torch.tensor([len(mine)])
        dist.all_reduce(t)
        out.put((rank, ok and int(t) == 7))
    finally:
        dist.destroy_process_group()


def test_epoch_sharding_and_gather_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == [0, 1] and all(ok for _, ok in res)


def test_weight_preparation_matches_oracle_algebra():
    """BatchNorm folding and the head permutation are pure host algebra: check them against the oracle on CPU."""
    from icepy4d_b200 import weights
    from icepy4d_b200.matching.superglue import _fold_bn, _head_perm
    from oracle import sg_oracle
    sd = weights.make_superglue_state(2)
    g = torch.Generator().manual_seed(0)
    # non-trivial BN statistics
    for k in list(sd):
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        if k.endswith("running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    x = torch.randn(37, 512, generator=g)
    w, b = _fold_bn(sd["gnn.layers.0.mlp.0.weight"][:, :, 0], sd["gnn.layers.0.mlp.0.bias"], sd, "gnn.layers.0.mlp.1")
    ref = sg_oracle._bn(sg_oracle._lin(x, sd, "gnn.layers.0.mlp.0"), sd, "gnn.layers.0.mlp.1")
    assert torch.allclose(x @ w.t() + b, ref, atol=1e-5)
    # head permutation: attention with permuted weights on contiguous heads == oracle MHA
    perm = _head_perm()
    xs, src = torch.randn(11, 256, generator=g), torch.randn(13, 256, generator=g)
    p = "gnn.layers.0.attn"
    q = (xs @ sd[f"{p}.proj.0.weight"][:, :, 0][perm].t() + sd[f"{p}.proj.0.bias"][perm]).view(11, 4, 64).permute(1, 0, 2)
    k = (src @ sd[f"{p}.proj.1.weight"][:, :, 0][perm].t() + sd[f"{p}.proj.1.bias"][perm]).view(13, 4, 64).permute(1, 0, 2)
    v = (src @ sd[f"{p}.proj.2.weight"][:, :, 0][perm].t() + sd[f"{p}.proj.2.bias"][perm]).view(13, 4, 64).permute(1, 0, 2)
    o = (torch.softmax(q @ k.transpose(1, 2) / 8, -1) @ v).permute(1, 0, 2).reshape(11, 256)
    out = o @ sd[f"{p}.merge.weight"][:, :, 0][:, perm].t() + sd[f"{p}.merge.bias"]
    assert torch.allclose(out, sg_oracle.mha(xs, src, sd, p), atol=1e-4)


def test_no_product_module_imports_the_oracle():
    """The oracle is test infrastructure: nothing under icepy4d_b200/ may import it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "icepy4d_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("needs a GPU-less host")
    from icepy4d_b200 import weights
    from icepy4d_b200.matching import SuperGlueMatcher
    with pytest.raises(RuntimeError):
        SuperGlueMatcher({"weights": "outdoor", "keypoint_threshold": 0.001, "max_keypoints": 128, "match_threshold": 0.2,
                          "force_cpu": False, "superpoint_state": weights.make_superpoint_state(1),
                          "superglue_state": weights.make_superglue_state(2)})


def test_features_points_containers_and_bundler_vs_reference(golden_dir, tmp_path):
    """the array-backed Features / Points containers and the Bundler exporter against what the reference's own
    classes (core/features.py, core/points.py, io/export2bundler.py) produced for the same seeded inputs
    (oracle/make_golden.py::golden_containers): same track ids (incl. the duplicate-id fallback), same arrays, same filters,
    and a byte-identical .out file."""
    from icepy4d_b200 import synthetic
    from icepy4d_b200.core import Feature, Features, Points
    from icepy4d_b200.io import write_bundler_out

    g = np.load(os.path.join(golden_dir, "containers.npz"))
    x, y, descr, scores = g["x"], g["y"], g["descr"], g["scores"]
    f = Features()
    f.append_features_from_numpy(x[:160], y[:160], descr[:, :160], scores[:160], epoch=3)
    f.append_features_from_numpy(x[160:], y[160:], descr[:, 160:], scores[160:], track_ids=[int(i) for i in range(1000, 1080)], epoch=3)
    assert len(f) == int(g["f_len"]) and int(f.last_track_id) == int(g["f_last"]) and f.num_features == len(f)
    assert np.array_equal(np.array(f.get_track_ids()), g["f_ids"])
    d = f.to_numpy(get_descr=True, get_score=True)
    for k, ref in (("kpts", "f_kpts"), ("descr", "f_descr"), ("scores", "f_scores")):
        assert d[k].dtype == np.float32 and np.array_equal(d[k], g[ref]), k
    assert set(f.to_numpy(get_score=True)) == {"kpts"}                       # the reference returns scores only with the descriptors
    one = f[1005]
    assert isinstance(one, Feature) and one.descr.shape == tuple(g["f_1005_descr_shape"]) and one.xy.shape == (1, 2)
    got = np.concatenate([one.xy.reshape(-1), [one.score], one.descr.reshape(-1)[:4], [one.track_id], [one.epoch]]).astype(np.float64)
    assert np.array_equal(got, g["f_1005"])
    assert isinstance(one.track_id, np.int32) and f[999999] is None
    assert [ft.track_id for _, ft in zip(range(3), f)] == [0, 1, 2]
    f.append_features_from_numpy(x[:5], y[:5], descr[:, :5], scores[:5], track_ids=[0, 1, 2, 3, 4])      # duplicates -> progressive ids
    assert np.array_equal(np.array(f.get_track_ids()), g["f_ids_dup"])
    f.filter_feature_by_mask(list(g["mask"]))
    assert np.array_equal(np.array(f.get_track_ids()), g["f_ids_masked"]) and np.array_equal(f.kpts_to_numpy(), g["f_kpts_masked"])
    keep = [int(i) for i in np.array(f.get_track_ids())[::3]]
    f.filter_feature_by_index(keep)
    assert np.array_equal(np.array(f.get_track_ids()), g["f_ids_indexed"]) and np.array_equal(f.scores_to_numpy(), g["f_scores_indexed"])
    assert np.array_equal(np.array([keep[0] in f, 999999 in f]), g["f_contains"])
    assert f.get_features_as_dict(get_track_id=True)["descriptors0"].shape == (128, len(f))
    with pytest.raises(AssertionError):
        f.append_features_from_numpy(x[:3], y[:3], np.zeros((256, 3), np.float32))                        # descriptor size mismatch
    with pytest.raises(ValueError):
        Features().append_features_from_numpy(x[:3].astype(np.float16), y[:3])
    e = Features()
    e.append_features_from_numpy(np.zeros(4, np.float32), np.zeros(4, np.float32))                        # `if not np.any(x)`: nothing done
    assert len(e) == 0 and e.last_track_id == -1

    m = 200
    fa, fb = Features(), Features()
    fa.append_features_from_numpy(x[:m], y[:m], descr[:, :m], scores[:m])
    kb = g["fb_kpts"]
    fb.append_features_from_numpy(kb[:, 0].copy(), kb[:, 1].copy(), descr[:, :m], scores[:m])
    pts = Points()
    pts.append_points_from_numpy(g["xyz"][:120], colors=g["colors"][:120])
    pts.append_points_from_numpy(g["xyz"][120:], track_ids=[int(i) for i in range(120, m)], colors=g["colors"][120:])
    assert np.array_equal(pts.to_numpy(), g["p_xyz"]) and np.array_equal(pts.colors_to_numpy(), g["p_col"])
    assert np.array_equal(pts.colors_to_numpy(as_uint8=True), g["p_col8"]) and np.array_equal(np.array(pts.get_track_ids()), g["p_ids"])
    assert int(pts.last_track_id) == int(g["p_last"]) and np.array_equal(pts[7].coordinates, g["p_xyz"][7])

    class Cam:
        width, height = 6012, 4008

        def __init__(self, c):
            self.K, self.dist, self.R, self.t = c.K, c.dist, c.R, c.t
    sc = synthetic.two_view_scene(n=8, seed=0, outlier_frac=0.0)["cams"]
    assert write_bundler_out(tmp_path, "epoch", {"cam1": "/data/cam1/a.jpg", "cam2": "/data/cam2/a.jpg"},
                             {"cam1": Cam(sc[0]), "cam2": Cam(sc[1])}, {"cam1": fa, "cam2": fb}, pts) is True
    assert open(tmp_path / "epoch.out", "rb").read() == g["bundler_out"].tobytes()
    assert open(tmp_path / "im_list.txt", "rb").read() == g["bundler_imlist"].tobytes()
    pts.filter_point_by_mask(g["pmask"])
    assert np.array_equal(np.array(pts.get_track_ids()), g["p_ids_masked"]) and int(pts.last_track_id) == int(g["p_last_masked"])
