import sys, torch
sys.path.insert(0,'.')
import satmvs_b200
from satmvs_b200 import synth
C,D,H,W,V=4,64,96,192,3
fe=[f.cuda() for f in synth.make_features(1,V,C,H,W)]
rp=synth.make_rpc_stack(1,V,H,W); dv=synth.make_depth_planes(1,D,H,W).cuda()
for _ in range(8): satmvs_b200.build_cost_volume(fe[0],fe[1:],rp[:,0],rp[:,1:],dv,"rpc")
torch.cuda.synchronize()
