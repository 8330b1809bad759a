"""GPU parity of the convolution block operator (C ABI `satmvs_conv_forward`): the tcgen05 tensor-core path (3-way TF32
split, `csrc/umma_conv.cuh`) and the fp32 FFMA path (`csrc/direct_conv.cuh`) against torch's fp64 CPU convolution, on shapes
beyond those the regularisers use.  Tolerances are relative to max|out|: 3xTF32 keeps ~2^-20 per product (asserted at 2e-5),
fp32 FFMA at 1e-5."""
import pytest
import torch
import torch.nn.functional as F

import satmvs_b200

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference(x, w, scale, shift, stride, relu, acc_scale):
    xd, wd = x.double(), w.double()
    if w.dim() == 5:
        y = F.conv3d(xd, wd, stride=stride, padding=1)
    else:
        B, C, D, H, W = xd.shape
        y = F.conv2d(xd.permute(0, 2, 1, 3, 4).reshape(B * D, C, H, W), wd, stride=stride, padding=1)
        y = y.reshape(B, D, *y.shape[1:]).permute(0, 2, 1, 3, 4)
    y = y * acc_scale
    shape = (1, -1, 1, 1, 1)
    if scale is not None:
        y = y * scale.double().view(shape)
    if shift is not None:
        y = y + shift.double().view(shape)
    return torch.relu(y) if relu else y


CASES = [
    # Cin, Cout, D, H, W, nz, stride, relu, engines
    (32, 24, 3, 96, 192, 1, 1, False, ("tcgen05", "ffma")),    # RED level-1 x-halves
    (32, 16, 2, 96, 192, 1, 2, True, ("tcgen05", "ffma")),     # stride-2 encoder as a tensor-core head
    (16, 40, 5, 13, 50, 1, 1, True, ("tcgen05",)),             # ragged plane, odd sizes, Cout not a multiple of 16
    (8, 1, 4, 24, 40, 1, 1, False, ("tcgen05",)),              # single output channel (ragged head)
    (64, 64, 2, 24, 392, 1, 1, False, ("tcgen05",)),           # 8 channel chunks, several strips
    (16, 16, 8, 24, 40, 3, 1, True, ("tcgen05", "ffma")),      # 3x3x3: three plane convs per output plane, z padding
    (8, 8, 8, 16, 24, 3, 2, True, ("ffma",)),                  # stride-2 3-D conv stays on the FFMA kernel
    (32, 8, 8, 96, 192, 3, 1, True, ("tcgen05_tn", "tcgen05", "ffma")),   # CostRegNet conv0: taps in N
    (8, 1, 6, 24, 40, 3, 1, False, ("tcgen05_tn",)),           # CostRegNet prob: one output channel, ragged planes
    (16, 5, 3, 13, 50, 1, 1, True, ("tcgen05_tn",)),           # 2-D, odd sizes, Cout < 8
]


@pytest.mark.parametrize("Cin,Cout,D,H,W,nz,stride,relu,engines", CASES)
def test_conv_block_against_fp64(Cin, Cout, D, H, W, nz, stride, relu, engines):
    g = torch.Generator().manual_seed(Cin * 1000 + Cout)
    x = torch.randn(1, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, *([3] * (3 if nz == 3 else 2)), generator=g) / (Cin * 9 * nz) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g)
    want = reference(x, w, scale, shift, stride, relu, -1.0)
    for eng in engines:
        got = satmvs_b200.conv_block(x.to(DEV), w.to(DEV), scale.to(DEV), shift.to(DEV), stride=stride, relu=relu,
                                     acc_scale=-1.0, engine=eng)
        torch.cuda.synchronize()
        assert got.shape == want.shape
        err = (got.cpu().double() - want).abs().max().item() / want.abs().max().item()
        assert err < (2e-5 if eng.startswith("tcgen05") else 1e-5), (eng, err)


def test_tensor_core_engine_refuses_what_it_cannot_do():
    x = torch.randn(1, 4, 1, 8, 8, device=DEV)           # Cin not a multiple of 8
    w = torch.randn(8, 4, 3, 3, device=DEV)
    with pytest.raises(RuntimeError, match="tcgen05"):
        satmvs_b200.conv_block(x, w, engine="tcgen05")
