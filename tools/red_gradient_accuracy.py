#!/usr/bin/env python
"""Accuracy of the RED regulariser's backward at cfg-2 size (C 32, 64 planes, 96x192) on one GPU: gradient of the training loss to
the variance volume against the oracle with everything in fp64 (on CUDA), next to torch's own fp32 autograd of the same oracle,
per plane, and with the exact dlogits fed in (isolates the regulariser from the head).  Output: profiles/r02_red_backward_accuracy.txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, satmvs_b200
from oracle import regnets, regress, volume
from satmvs_b200 import synth, training
dev = "cuda:0"
w = bench.WORKLOADS["cfg2_train_red"]
fe, cams, dv = bench.make_inputs(w)
var = volume.variance_cost_volume(fe, cams, dv, w["geo"]).to(dev)
dv = dv.to(dev)
gt = torch.full((1, w["H"], w["W"]), 500.0, device=dev)
sd = synth.make_red_weights(w["C"])
def oracle(dt):
    sdt = {k: v.to(dev, dt).requires_grad_(True) for k, v in sd.items()}
    v = var.to(dt).detach().requires_grad_(True)
    logits = bench.red_on_device(v, sdt, dev, regnets)
    depth, _ = regress.softargmin_red(logits, dv.to(dt))
    logits.retain_grad()
    torch.nn.functional.smooth_l1_loss(depth, gt.to(dt)).backward()
    return v.grad, logits.grad, depth.detach(), logits.detach()
g64, gl64, d64, l64 = oracle(torch.float64)
g32, gl32, d32, l32 = oracle(torch.float32)
net = satmvs_b200.RED_Regularization(w["C"], 8); net.load_state_dict(sd); net = net.to(dev).train()
v = var.detach().requires_grad_(True)
logits = net(v); logits.retain_grad()
depth, _ = training.softargmin_train(logits, dv, "red")
torch.nn.functional.smooth_l1_loss(depth, gt).backward()
def rep(name, a, b):
    d = (a.double() - b).abs(); s = b.abs().max().item()
    i = d.argmax().item(); idx = []
    for n in reversed(a.shape): idx.append(i % n); i //= n
    print(name, "rel_linf %.3e" % (d.max().item() / s), "at", list(reversed(idx)), "mean rel %.3e" % (d.mean().item() / s))
    return d / s
rep("logits ours", logits.detach(), l64); rep("logits fp32", l32, l64)
rep("dlogits ours", logits.grad, gl64); rep("dlogits fp32", gl32, gl64)
e = rep("dvar ours", v.grad, g64); rep("dvar fp32", g32, g64)
print("per-plane max err ours:", ["%.1e" % e[0, :, d].max().item() for d in range(0, 64, 4)])
e32 = (g32.double() - g64).abs() / g64.abs().max()
print("per-plane max err fp32:", ["%.1e" % e32[0, :, d].max().item() for d in range(0, 64, 4)])
# feed the fp64 dlogits into our backward to isolate the regulariser backward from the head
v2 = var.detach().requires_grad_(True)
lg = net(v2); lg.backward(gl64.float())
rep("dvar ours given exact dlogits", v2.grad, g64)
