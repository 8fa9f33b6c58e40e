"""CPU-only access to the synthetic-input generator and stochasticSubSample: libdsac_synth.so is
host_util.cpp compiled by g++ with no CUDA dependency (dsac_b200.build.build_synth).  For consumers that must not map
the CUDA library, such as the CPU arm of bench.py."""
import ctypes as C
import os

import numpy as np

from . import build as _build

N = 1600
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.SYNTH_LIB
        if not os.path.exists(path):
            _build.build_synth()
        _lib = C.CDLL(path)
        _lib.dsac_synth_frames.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_int64, C.c_int32, C.c_double, C.c_double,
                                           C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
    return _lib


def synth_frames(n_frames, frame0=0, data_seed=20170721, sampler_seed=1305, n_streams=1, rho=0.5, sigma=25.0, traj=False,
                 focal=525.0, cx=320.0, cy=240.0):
    """Same frames as dsac_b200.engine.synth_frames (it is the same function of the same source file)."""
    coords = np.empty((n_frames, N, 3), np.int16)
    pix = np.empty((n_frames, N, 2), np.int32)
    gt_cv = np.empty((n_frames, 6))
    gt_jp = np.empty((n_frames, 12))
    p = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    rc = _load().dsac_synth_frames(data_seed, sampler_seed, n_streams, frame0, n_frames, rho, sigma, int(traj), focal, cx, cy,
                                   p(coords), p(pix), p(gt_cv), p(gt_jp))
    assert rc == 0
    return coords, pix, gt_cv, gt_jp
