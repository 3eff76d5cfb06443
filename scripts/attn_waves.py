"""How the attention kernel's time scales with the number of work items (one CTA per SM, 148 SMs): T per wave and the effect of
splitting the ragged last wave."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops_tc
NK = 8192
qkv = (torch.randn(2 * NK, 768, device="cuda") * 1.5).bfloat16()
att = torch.empty(2 * NK, 256, device="cuda", dtype=torch.bfloat16)
def t(problems, heads=4, reps=10):
    for _ in range(2): ops_tc.attention_tc(qkv, problems, att, 0, 256, 512, heads=heads)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops_tc.attention_tc(qkv, problems, att, 0, 256, 512, heads=heads)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for nq_tiles, nprob in ((37, 1), (74, 1), (55, 1), (64, 1), (64, 2), (56, 2), (48, 2)):
    nq = nq_tiles * 128
    probs = [(0, min(nq, NK), 0, NK)] if nprob == 1 else [(0, nq, 0, NK), (NK, nq, NK, NK)]
    if nq > NK:
        probs = [(0, nq, 0, NK)]
    items = nq_tiles * 4 * nprob
    print(f"q-tiles {nq_tiles} x 4 heads x {nprob} problems = {items} items ({items / 148:.2f} waves): {t(probs):.1f} us")
