"""Mirror of `tools/rpc_tensor.py:RPCModelParameter` (the CuPy localisation -> projection model
used by the geometric-consistency filter, `tools/rpc_filter.py:11-45`) on libsatmvs_b200.so.

Same constructor argument (a float64[170] vector, layout `tools/rpc_tensor.py:11-22`) and the same
two methods taking / returning host numpy arrays; `*_device` variants keep the points on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class RPCModelParameter:
    def __init__(self, data=np.zeros(170, dtype=np.float64)):
        self.data = np.ascontiguousarray(np.asarray(data, dtype=np.float64).reshape(-1))
        if self.data.shape[0] != 170:
            raise ValueError("an RPC vector has 170 entries (10 offsets/scales + 8 x 20 coefficients)")

    def load_dirpc_from_file(self, filepath: str) -> None:
        """170 lines `name value` (`tools/rpc_tensor.py:79-107`)."""
        with open(filepath, "r") as f:
            vals = [line.split(" ")[1] for line in f.read().splitlines()]
        self.__init__(np.array(vals, dtype=np.float64))

    def _run(self, fn, a, b, h):
        a, b, h = (torch.as_tensor(v, dtype=torch.float64) for v in (a, b, h))
        if not (a.shape == b.shape == h.shape):
            raise AssertionError("coordinate arrays must share one shape")      # rpc_tensor.py:110,139
        dev = a.device if a.is_cuda else torch.device("cuda", torch.cuda.current_device())
        a, b, h = (v.to(dev).contiguous() for v in (a, b, h))
        o1, o2 = torch.empty_like(a), torch.empty_like(a)
        with torch.cuda.device(dev):
            _lib.check(fn(self.data.ctypes.data_as(C.c_void_p), a.data_ptr(), b.data_ptr(), h.data_ptr(), a.numel(),
                          o1.data_ptr(), o2.data_ptr(), _lib.stream_ptr(dev)), "rpc point op")
        return o1, o2

    def RPC_PHOTO2OBJ_device(self, insamp, inline, inhei):
        """(samp, line, height) -> (lat, lon), tensors stay on the GPU."""
        return self._run(_lib.lib().satmvs_rpc_localise, insamp, inline, inhei)

    def RPC_OBJ2PHOTO_device(self, inlat, inlon, inhei):
        """(lat, lon, height) -> (samp, line), tensors stay on the GPU."""
        return self._run(_lib.lib().satmvs_rpc_project, inlat, inlon, inhei)

    def RPC_PHOTO2OBJ(self, insamp, inline, inhei):
        """`tools/rpc_tensor.py:138-165`: numpy in, numpy out."""
        lat, lon = self.RPC_PHOTO2OBJ_device(insamp, inline, inhei)
        return lat.cpu().numpy(), lon.cpu().numpy()

    def RPC_OBJ2PHOTO(self, inlat, inlon, inhei):
        """`tools/rpc_tensor.py:109-136`: numpy in, numpy out."""
        samp, line = self.RPC_OBJ2PHOTO_device(inlat, inlon, inhei)
        return samp.cpu().numpy(), line.cpu().numpy()
