#!/usr/bin/env python
"""Timeline (CUDA events around every instrumented launch) of one RED_Regularization forward: SATMVS_PROF_DUMP=1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SATMVS_PROF_DUMP"] = "1"
import torch, satmvs_b200
from satmvs_b200 import synth, _lib
torch.set_grad_enabled(False)
m = satmvs_b200.RED_Regularization(32, 8)
m.load_state_dict(synth.make_red_weights(32, seed=7))
m = m.to("cuda:0")
x = torch.rand(1, 32, 64, 96, 192, device="cuda:0")
for _ in range(3):
    m(x)
torch.cuda.synchronize()
with _lib.profile() as prof:
    m(x)
print(prof.ms)
