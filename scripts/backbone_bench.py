"""cuDNN backbone variants at one 1999x1999 tile (and batches): which torch configuration is fastest."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from icepy4d_b200 import weights
sd = weights.make_superpoint_state(1)
names = ["conv1a","conv1b","conv2a","conv2b","conv3a","conv3b","conv4a","conv4b","convPa","convPb","convDa","convDb"]
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def make(dtype, cl):
    w = {}
    for n in names:
        W = sd[n+".weight"].cuda().to(dtype); b = sd[n+".bias"].cuda().to(dtype)
        if cl: W = W.contiguous(memory_format=torch.channels_last)
        w[n] = (W, b)
    return w
def run(x, w, fused):
    def cr(x, n, pad, relu=True):
        W, b = w[n]
        if fused and relu:
            return torch.cudnn_convolution_relu(x, W, b, (1,1), (pad,pad), (1,1), 1)
        y = F.conv2d(x, W, b, padding=pad)
        return F.relu_(y) if relu else y
    x = cr(cr(x,"conv1a",1),"conv1b",1); x = F.max_pool2d(x,2,2)
    x = cr(cr(x,"conv2a",1),"conv2b",1); x = F.max_pool2d(x,2,2)
    x = cr(cr(x,"conv3a",1),"conv3b",1); x = F.max_pool2d(x,2,2)
    x = cr(cr(x,"conv4a",1),"conv4b",1)
    l = cr(cr(x,"convPa",1),"convPb",0,False); d = cr(cr(x,"convDa",1),"convDb",0,False)
    return l, d
torch.backends.cudnn.benchmark = True
for B in (1, 2, 6):
    img = torch.rand(B,1,1999,1999, device="cuda")
    for (dtype, cl, fused, tf32) in ((torch.float32, False, False, True), (torch.float32, True, False, True), (torch.float32, True, True, True),
                                     (torch.bfloat16, True, False, True), (torch.bfloat16, True, True, True), (torch.bfloat16, False, True, True),
                                     (torch.float16, True, True, True), (torch.float32, True, False, False)):
        torch.backends.cudnn.allow_tf32 = tf32
        w = make(dtype, cl)
        x = img.to(dtype)
        if cl: x = x.contiguous(memory_format=torch.channels_last)
        try:
            with torch.inference_mode():
                ms = timeit(lambda: run(x, w, fused))
            print(f"B={B} {str(dtype):16s} channels_last={cl!s:5s} fused_relu={fused!s:5s} tf32={tf32!s:5s}: {ms/B:8.2f} ms/image  ({676.4*B/ms:6.1f} TFLOP/s)", flush=True)
        except Exception as e:
            print(f"B={B} {dtype} cl={cl} fused={fused}: FAILED {str(e)[:100]}", flush=True)
