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
