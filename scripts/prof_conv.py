"""ncu driver: one launch of each split-bf16 convolution shape of the SuperPoint backbone (2000x2000 tile)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icepy4d_b200 import ops, weights
from icepy4d_b200.matching.superpoint import SuperPointB200
sp = SuperPointB200(weights.make_superpoint_state(1), conv_precision="f16x3")
img = torch.rand(1, 1, 2000, 2000, device="cuda")
for _ in range(2):
    sp.backbone(img)
torch.cuda.synchronize()
