"""Wire / on-disk formats on either side of the hot path: the 170-value RPC
text file and PFM height maps.  Host-side numpy only; same function names, arguments, return values
and error behaviour as the reference's `dataset/data_io.py` (`load_pfm` :17-42, `save_pfm` :45-73,
`load_rpc_as_array` :77-92), written independently.  No GDAL / PIL dependency."""
from __future__ import annotations

import os
import sys

import numpy as np

_PFM_CHANNELS = {b"PF": 3, b"Pf": 1}


def load_pfm(fname):
    """PFM -> float32 array, rows top-to-bottom (the file stores them bottom-to-top); [H,W] for 'Pf',
    [H,W,3] for 'PF'.  A negative scale line means little-endian samples (data_io.py:33-37)."""
    with open(fname, "rb") as f:
        magic = f.readline().rstrip()
        if magic not in _PFM_CHANNELS:
            raise Exception("Not a PFM file.")
        dims = f.readline().decode("latin-1").split()
        if len(dims) != 2 or not all(d.isdigit() for d in dims):
            raise Exception("Malformed PFM header.")
        width, height = int(dims[0]), int(dims[1])
        scale = float(f.readline().decode("latin-1").strip())
        raw = np.frombuffer(bytearray(f.read()), dtype="<f4" if scale < 0 else ">f4")   # writable, like np.fromfile
    nc = _PFM_CHANNELS[magic]
    shape = (height, width, 3) if nc == 3 else (height, width)
    return np.ascontiguousarray(raw.reshape(shape)[::-1])


def save_pfm(file, image, scale=1):
    """float32 [H,W], [H,W,1] or [H,W,3] -> PFM, native byte order recorded in the sign of the scale
    line (data_io.py:64-69)."""
    image = np.asarray(image)
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        magic = b"PF\n"
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        magic = b"Pf\n"
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    order = image.dtype.byteorder
    little = order == "<" or (order == "=" and sys.byteorder == "little")
    with open(file, "wb") as f:
        f.write(magic)
        f.write(b"%d %d\n" % (image.shape[1], image.shape[0]))
        f.write(("%f\n" % (-scale if little else scale)).encode("utf8"))
        f.write(np.ascontiguousarray(image[::-1]).tobytes())


def load_rpc_as_array(filepath):
    """RPC text file ('KEY value [unit]' per line, 170 lines in the order of SURVEY.md §8 a5) ->
    (float64[170], h_max, h_min) with h = HEIGHT_OFF +- HEIGHT_SCALE (data_io.py:77-92)."""
    if not os.path.exists(filepath):
        raise Exception("RPC not found! Can not find " + filepath + " in the file system!")
    with open(filepath, "r") as f:
        values = [line.split(" ")[1] for line in f.read().splitlines()]
    data = np.array(values, dtype=np.float64)
    return data, data[4] + data[9], data[4] - data[9]
