#!/usr/bin/env python
"""One weight-gradient launch at a cfg-2 layer shape (target of ncu captures): CostRegNet conv0 (Cin 32, Cout 8, 64 x 96 x 192,
3x3x3) by default; `red` = the level-0 gate filters of the RED regulariser (Cin 40, Cout 16, 64 planes of 96 x 192, 3x3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from satmvs_b200 import training
red = len(sys.argv) > 1 and sys.argv[1] == "red"
cin, cout, nz = (40, 16, 1) if red else (32, 8, 3)
x = torch.randn(1, cin, 64, 96, 192, device="cuda:0")
dy = torch.randn(1, cout, 64, 96, 192, device="cuda:0")
dw = torch.empty(cout, cin, *( (3, 3) if red else (3, 3, 3)), device="cuda:0")
taps = 9 if red else 27
for _ in range(3):
    training.conv3d_wgrad(x, dy, 1, dw, cin * taps, taps, nz=nz)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); training.conv3d_wgrad(x, dy, 1, dw, cin * taps, taps, nz=nz); e.record(); torch.cuda.synchronize()
gmac = cin * cout * taps * 64 * 96 * 192 / 1e9
print("wgrad %s: %.3f ms, %.1f TFLOP/s" % ("red level 0 gates" if red else "costreg conv0", s.elapsed_time(e), 2 * gmac / s.elapsed_time(e)))
