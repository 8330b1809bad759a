#!/usr/bin/env python
"""One FeatureNet forward over 3 views of 768x384 (target of ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, satmvs_b200
from satmvs_b200 import synth
torch.set_grad_enabled(False)
m = satmvs_b200.FeatureNet(8)
m.load_state_dict(synth.make_featurenet_weights(8))
m = m.to("cuda:0").eval()
imgs = [torch.rand(1, 3, 384, 768, device="cuda:0") for _ in range(3)]
for _ in range(3):
    out = m.forward_views(imgs)
torch.cuda.synchronize()
print(float(out[0]["stage1"].abs().mean()))
