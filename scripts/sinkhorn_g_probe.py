"""Per-iteration time of the fused Sinkhorn kernel on a subset of the SMs (I4D_SK_G=<CTAs>), alone and with two launches
running concurrently on two streams: what interleaving two tile pairs' solves would buy."""
import os, sys, subprocess

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

if len(sys.argv) > 1:
    from icepy4d_b200 import ops
    M = N = 8192
    torch.manual_seed(0)
    S = [torch.randn(M, N, device="cuda") * 3.0 for _ in range(2)]
    ws = [ops.AssignWorkspace(M, N, torch.device("cuda")) for _ in range(2)]
    st = [torch.cuda.Stream() for _ in range(2)]
    iters = 100

    def one():
        ops.sinkhorn(S[0], 1.0, iters, ws[0])

    def two():
        cur = torch.cuda.current_stream()
        for k in range(2):
            st[k].wait_stream(cur)
            with torch.cuda.stream(st[k]):
                ops.sinkhorn(S[k], 1.0, iters, ws[k])
        for k in range(2):
            cur.wait_stream(st[k])

    for name, fn, n in (("one launch", one, 1), ("two concurrent launches", two, 2)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"I4D_SK_G={os.environ.get('I4D_SK_G', '-'):>4}  {name:26s} {ms:8.3f} ms  = {ms * 1e3 / iters / n:6.2f} us per iteration and problem")
else:
    for g in ("", "74", "73", "49"):
        env = dict(os.environ)
        if g:
            env["I4D_SK_G"] = g
        subprocess.run([sys.executable, __file__, "run"], env=env, check=False)
