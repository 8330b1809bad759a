#!/usr/bin/env python
"""Times RED_Regularization.forward at a shape under the environment's SATMVS_RED_* knobs; prints which recurrence ran."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, satmvs_b200
from satmvs_b200 import synth, _lib
torch.set_grad_enabled(False)
C, D, H, W = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 64, 96, 192)))
m = satmvs_b200.RED_Regularization(C, 8)
m.load_state_dict(synth.make_red_weights(C, seed=7))
m = m.to("cuda:0")
x = torch.rand(1, C, D, H, W, device="cuda:0")
for _ in range(5):
    y = m(x)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); y = m(x); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
print({k: v for k, v in os.environ.items() if k.startswith("SATMVS_")}, "path", _lib.lib().satmvs_red_last_path(),
      "median ms %.3f min %.3f" % (ts[len(ts) // 2], ts[0]), "checksum %.6f" % float(y.double().abs().mean()))
