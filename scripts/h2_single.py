"""Run the fp16-split conv on the 3x3 128->128 @200x176 layer a few times (target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import torch
from sessd_b200 import ops
x = torch.randn(1, 200, 176, 128, device="cuda"); wp = torch.randn(9, 128, 128, device="cuda") * 0.05
planes, inv = ops.pack_weight_h2(wp, 128); out = torch.zeros(1, 200, 176, 128, device="cuda")
t3 = [(dy - 1, dx - 1) for dy in range(3) for dx in range(3)]
d = ops.conv_desc(1, (200, 176), 128, (200, 176), 128, (200, 176), t3, relu=True)
amax = torch.zeros(1, device="cuda"); ops.absmax(x, amax)
sc = inv[:128].contiguous()
for _ in range(4):
    ops.bev_conv_h2(x, planes, sc, None, None, out, d, amax, None)
torch.cuda.synchronize()
print("ok")
