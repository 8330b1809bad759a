#!/usr/bin/env python
"""One RED_Regularization forward at the cfg-2 stage-1 shape (for ncu captures of the recurrence kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, satmvs_b200
torch.set_grad_enabled(False)
from satmvs_b200 import synth
C, D, H, W = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 64, 96, 192)))
m = satmvs_b200.RED_Regularization(C, 8)
m.load_state_dict(synth.make_red_weights(C, seed=7))
m = m.to("cuda:0")
x = torch.rand(1, C, D, H, W, device="cuda:0")
for _ in range(2):
    y = m(x)
torch.cuda.synchronize()
print(float(y.abs().mean()))
