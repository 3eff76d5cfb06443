"""Where is the GPU idle inside one cfg2 epoch?  Records one warm epoch with the torch profiler (CUPTI kernel / memcpy activity
records), merges the device intervals and lists the busy time, the idle time and the largest gaps with the kernels around them.
Timings under the profiler are inflated; the point is the split busy / idle and where the gaps are."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from icepy4d_b200 import synthetic
from icepy4d_b200.epoch import make_cfg2_pipeline

pipe = make_cfg2_pipeline(8192, 100, precision="bf16", conv_precision="f16x3")
i0, i1 = synthetic.stereo_pair(4000, 6000, seed=1000, shift=(16, 24), channels=3)
d0, d1 = torch.from_numpy(i0).cuda(), torch.from_numpy(i1).cuda()
for _ in range(3):
    pipe.run_device(d0, d1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pipe.run_device(d0, d1)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
        ev.append((e.time_range.start, e.time_range.end, e.name))
ev.sort()
t0, t1 = ev[0][0], max(e[1] for e in ev)
busy, cur_s, cur_e, gaps, last_name = 0.0, ev[0][0], ev[0][1], [], ev[0][2]
for s, e, name in ev[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e - t0, last_name, name))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
    last_name = name if e >= cur_e else last_name
busy += cur_e - cur_s
span = t1 - t0
print(f"{len(ev)} device activities | span {span / 1e3:.2f} ms | busy {busy / 1e3:.2f} ms | idle {(span - busy) / 1e3:.2f} ms ({100 * (span - busy) / span:.1f} %)")
hist = [(5, 0, 0.0), (20, 0, 0.0), (100, 0, 0.0), (1e9, 0, 0.0)]
for g in gaps:
    for k, (lim, n, tot) in enumerate(hist):
        if g[0] <= lim:
            hist[k] = (lim, n + 1, tot + g[0]); break
lo = 0
for lim, n, tot in hist:
    print(f"  gaps in ({lo}, {lim if lim < 1e9 else 'inf'}] us: {n:5d}, {tot / 1e3:7.3f} ms")
    lo = lim
print("largest gaps (us, at ms, after kernel -> before kernel):")
for g in sorted(gaps, reverse=True)[:25]:
    print(f"  {g[0]:8.1f} us at {g[1] / 1e3:7.2f} ms   {g[2][:60]}  ->  {g[3][:60]}")
