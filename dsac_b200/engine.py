"""Python host-side binding of the C ABI (include/dsac_b200.h) via ctypes.

This is a thin mirror for tests and bench.py; the C/C++ surface (include/, dsac_b200/host/)
is the drop-in boundary.  There is no CPU fallback: `Engine(...)` raises if the CUDA
library is missing or no device is present.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

GRID = 40
N = GRID * GRID
STAGE_SAMPLE, STAGE_SCORE, STAGE_REFINE, STAGE_EVAL, STAGE_ALL = 1, 2, 4, 8, 15
ST_SAMPLER_EXHAUSTED, ST_REFINE_ABORTED = 1, 2

# every symbol include/dsac_b200.h declares
EXPORTS = [
    "dsac_default_config", "dsac_engine_create", "dsac_engine_destroy", "dsac_last_error", "dsac_engine_config",
    "dsac_forward", "dsac_forward_submit", "dsac_forward_wait", "dsac_forward_device", "dsac_fetch", "dsac_device_view_get", "dsac_set_stages",
    "dsac_set_tail_split", "dsac_launch_count", "dsac_sampler_profile", "dsac_sampler_profile_read", "dsac_debug_spec_result", "dsac_set_score_hook", "dsac_set_score_backward_hook", "dsac_backward", "dsac_forward_dsac", "dsac_backward_dsac", "dsac_gather_patches_device", "dsac_coords_from_prediction_device", "dsac_kabsch",
    "dsac_stochastic_subsample",
    "dsac_synth_frames", "dsac_version",
]


class Config(C.Structure):
    _fields_ = [
        ("focal", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("n_hyps", C.c_int32), ("thr2d", C.c_int32), ("inlier_count", C.c_int32), ("ref_steps", C.c_int32),
        ("sub_sample", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
        ("seed", C.c_uint32), ("n_streams", C.c_int32), ("stream_skip", C.c_uint32), ("max_candidates", C.c_int32),
        ("fix_q4", C.c_int32), ("grad_clamp", C.c_double),
        ("write_diffmaps", C.c_int32), ("device", C.c_int32), ("max_frames", C.c_int32), ("hyps_per_cta", C.c_int32),
    ]


_OUT_FIELDS = ["hyp_pose", "img_idx", "cand_idx", "scores", "sf", "diffmaps", "entropy", "avg_pose", "ref_pose",
               "inlier_map", "ref_steps_done", "n_perm_steps", "loss", "rot_err", "t_err", "correct", "n_candidates",
               "status"]


class ForwardOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _OUT_FIELDS]


class DeviceView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("diffmaps", "hyp_pose", "scores", "sf", "avg_pose", "ref_pose", "img_idx")]


_BW_FIELDS = ["dloss_dobj", "dloss_dref", "dref_dhyp", "dref_dobj", "score_grads", "dpnp"]


class BackwardOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _BW_FIELDS]


_DSAC_FIELDS = ["hyp_pose", "img_idx", "sf", "entropy", "ref_pose", "losses", "inlier_maps", "steps_done", "expected_loss",
                "hyp_idx", "rot_err", "t_err", "correct", "status"]


class DsacOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _DSAC_FIELDS]


SCORE_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p)
SCORE_BACKWARD_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p)

_lib = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """dlopen libdsac_b200.so (building it with nvcc first if needed)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    alt = os.environ.get("DSAC_B200_LIB")   # development aid (tools/sweep.py): another BUILD of the same CUDA library
    if alt:
        path = alt
    elif build_if_missing:
        _build.build()
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: build it with `python -m dsac_b200.build` (no CPU fallback exists)" % path)
    lib = C.CDLL(path)
    lib.dsac_last_error.restype = C.c_char_p
    lib.dsac_last_error.argtypes = [C.c_void_p]
    lib.dsac_version.restype = C.c_char_p
    lib.dsac_launch_count.restype = C.c_int64
    lib.dsac_launch_count.argtypes = [C.c_void_p]
    lib.dsac_engine_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.dsac_engine_destroy.argtypes = [C.c_void_p]
    lib.dsac_engine_destroy.restype = None
    lib.dsac_forward.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                 C.POINTER(ForwardOut)]
    lib.dsac_forward_submit.argtypes = lib.dsac_forward.argtypes
    lib.dsac_forward_wait.argtypes = [C.c_void_p]
    lib.dsac_forward_device.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                        C.c_void_p]
    lib.dsac_fetch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(ForwardOut), C.c_void_p]
    lib.dsac_device_view_get.argtypes = [C.c_void_p, C.POINTER(DeviceView)]
    lib.dsac_sampler_profile.argtypes = [C.c_void_p, C.c_int32]
    lib.dsac_sampler_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dsac_debug_spec_result.argtypes = [C.c_void_p, C.c_void_p]
    lib.dsac_set_stages.argtypes = [C.c_void_p, C.c_uint32]
    lib.dsac_set_tail_split.argtypes = [C.c_void_p, C.c_int32]
    lib.dsac_set_score_hook.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dsac_set_score_backward_hook.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dsac_backward.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.POINTER(BackwardOut)]
    lib.dsac_forward_dsac.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                      C.POINTER(DsacOut)]
    lib.dsac_backward_dsac.argtypes = [C.c_void_p, C.c_int32, C.POINTER(BackwardDsacOut)]
    lib.dsac_gather_patches_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                               C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dsac_coords_from_prediction_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dsac_kabsch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dsac_stochastic_subsample.argtypes = [C.c_uint32, C.c_int32, C.c_int32, C.c_void_p]
    lib.dsac_synth_frames.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_int64, C.c_int32, C.c_double, C.c_double,
                                      C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_config(**kw):
    c = Config()
    load().dsac_default_config(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def stochastic_subsample(seed=1305, width=640, height=480):
    pix = np.empty((N, 2), np.int32)
    rc = load().dsac_stochastic_subsample(seed, width, height, _p(pix))
    assert rc == 0
    return pix


def synth_frames(n_frames, frame0=0, data_seed=20170721, sampler_seed=1305, n_streams=1, rho=0.5, sigma=25.0,
                 traj=False, focal=525.0, cx=320.0, cy=240.0):
    """Synthetic frames of SURVEY.md section 8(d): (coords int16 [n,N,3], pix int32 [n,N,2], gt_cv [n,6], gt_jp [n,12])."""
    coords = np.empty((n_frames, N, 3), np.int16)
    pix = np.empty((n_frames, N, 2), np.int32)
    gt_cv = np.empty((n_frames, 6))
    gt_jp = np.empty((n_frames, 12))
    rc = load().dsac_synth_frames(data_seed, sampler_seed, n_streams, frame0, n_frames, rho, sigma, int(traj), focal, cx,
                                  cy, _p(coords), _p(pix), _p(gt_cv), _p(gt_jp))
    assert rc == 0
    return coords, pix, gt_cv, gt_jp


class ForwardResult:
    """Host copies of the processImage outputs for a batch of frames."""

    def __init__(self, n, H, want_diffmaps):
        self.hyp_pose = np.zeros((n, H, 6))
        self.img_idx = np.zeros((n, H, 4), np.int32)
        self.cand_idx = np.zeros((n, H), np.int32)
        self.scores = np.zeros((n, H))
        self.sf = np.zeros((n, H))
        self.diffmaps = np.zeros((n, H, N), np.float32) if want_diffmaps else None
        self.entropy = np.zeros(n)
        self.avg_pose = np.zeros((n, 6))
        self.ref_pose = np.zeros((n, 6))
        self.inlier_map = np.zeros((n, N), np.int32)
        self.ref_steps_done = np.zeros(n, np.int32)
        self.n_perm_steps = np.zeros(n, np.int32)
        self.loss = np.zeros(n)
        self.rot_err = np.zeros(n)
        self.t_err = np.zeros(n)
        self.correct = np.zeros(n, np.int32)
        self.n_candidates = np.zeros(n, np.int64)
        self.status = np.zeros(n, np.uint32)
        self.raw = ForwardOut()
        for k in _OUT_FIELDS:
            setattr(self.raw, k, _p(getattr(self, k)))


class DsacResult:
    """Host copies of the DSAC-variant processImage outputs (core/cnn.h:1028-1257)."""

    def __init__(self, n, H, want_inlier_maps):
        self.hyp_pose = np.zeros((n, H, 6)); self.img_idx = np.zeros((n, H, 4), np.int32); self.sf = np.zeros((n, H))
        self.entropy = np.zeros(n); self.ref_pose = np.zeros((n, H, 6)); self.losses = np.zeros((n, H))
        self.inlier_maps = np.zeros((n, H, N), np.int32) if want_inlier_maps else None
        self.steps_done = np.zeros((n, H), np.int32); self.expected_loss = np.zeros(n); self.hyp_idx = np.zeros(n, np.int32)
        self.rot_err = np.zeros(n); self.t_err = np.zeros(n); self.correct = np.zeros(n, np.int32); self.status = np.zeros(n, np.uint32)
        self.raw = DsacOut()
        for k in _DSAC_FIELDS:
            setattr(self.raw, k, _p(getattr(self, k)))


_BWD_FIELDS = ("dloss_dobj", "path1", "path2", "score_grads", "n_selected", "n_refine_jobs")


class BackwardDsacOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _BWD_FIELDS]


class BackwardDsacResult:
    """dsac_backward_dsac outputs (train_ransac.cpp:304-381)."""

    def __init__(self, n, H):
        self.dloss_dobj = np.zeros((n, N, 3)); self.path1 = np.zeros((n, N, 3)); self.path2 = np.zeros((n, N, 3))
        self.score_grads = np.zeros((n, H)); self.n_selected = np.zeros(n, np.int32); self.n_refine_jobs = np.zeros(n, np.int32)
        self.raw = BackwardDsacOut()
        for k in _BWD_FIELDS:
            setattr(self.raw, k, _p(getattr(self, k)))


class BackwardResult:
    def __init__(self, n, H, full=True):
        self.dloss_dobj = np.zeros((n, N, 3))
        self.dloss_dref = np.zeros((n, 6))
        self.dref_dhyp = np.zeros((n, 6, 6))
        self.dref_dobj = np.zeros((n, 6, N * 3)) if full else None
        self.score_grads = np.zeros((n, H))
        self.dpnp = np.zeros((n, H, 6, 12)) if full else None
        self.raw = BackwardOut()
        for k in _BW_FIELDS:
            setattr(self.raw, k, _p(getattr(self, k)))


class Engine:
    """One engine per GPU (dsac_engine_create / dsac_engine_destroy)."""

    def __init__(self, cfg=None, **kw):
        self.lib = load()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        if cfg is not None:
            for k, v in kw.items():
                setattr(self.cfg, k, v)
        h = C.c_void_p()
        rc = self.lib.dsac_engine_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError("dsac_engine_create failed (%d): %s" % (rc, self.lib.dsac_last_error(None).decode()))
        self.h = h
        self._hook = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.dsac_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("dsac call failed (%d): %s" % (rc, self.lib.dsac_last_error(self.h).decode()))

    @property
    def launches(self):
        return int(self.lib.dsac_launch_count(self.h))

    def set_stages(self, mask):
        self._check(self.lib.dsac_set_stages(self.h, mask))

    def sampler_profile(self, enable=True):
        self._check(self.lib.dsac_sampler_profile(self.h, int(enable)))

    def sampler_profile_read(self):
        """(ms[4], counts[4]) of the last profiled forward: generation+selection, filter, solve, tail; candidates, flagged,
        accepted, rounds."""
        ms = (C.c_double * 4)()
        cnt = (C.c_uint64 * 4)()
        self._check(self.lib.dsac_sampler_profile_read(self.h, ms, cnt))
        return list(ms), [int(x) for x in cnt]

    def spec_result(self):
        """Windows stitched per stream by the speculative first round of the last pass (0: abandoned / not used)."""
        out = (C.c_int32 * 8)()
        self._check(self.lib.dsac_debug_spec_result(self.h, out))
        return [int(x) for x in out]

    def set_tail_split(self, mode):
        """0: off, 1: forward_device + blocking forward (default), 2: submitted passes too (scheduling only)."""
        self._check(self.lib.dsac_set_tail_split(self.h, mode))

    def forward(self, coords, pix, gt_jp=None, frame0=0, want_diffmaps=False, out=None):
        """dsac_forward with HOST buffers (numpy arrays)."""
        coords = np.ascontiguousarray(coords, np.int16).reshape(-1, N, 3)
        n = coords.shape[0]
        pix = np.ascontiguousarray(pix, np.int32)
        shared = 1 if pix.size == N * 2 else 0
        assert shared or pix.size == n * N * 2
        gt = np.ascontiguousarray(gt_jp, np.float64).reshape(n, 12) if gt_jp is not None else None
        res = out if out is not None else ForwardResult(n, self.cfg.n_hyps, want_diffmaps)
        self._check(self.lib.dsac_forward(self.h, n, frame0, _p(coords), _p(pix), shared, _p(gt), C.byref(res.raw)))
        return res

    def forward_submit(self, coords, pix, gt_jp=None, frame0=0, out=None, want_diffmaps=False):
        """dsac_forward_submit: enqueue a pass and return; the arrays must stay alive (ideally pinned) until forward_wait()."""
        coords = np.ascontiguousarray(coords, np.int16).reshape(-1, N, 3)
        n = coords.shape[0]
        pix = np.ascontiguousarray(pix, np.int32)
        shared = 1 if pix.size == N * 2 else 0
        gt = np.ascontiguousarray(gt_jp, np.float64).reshape(n, 12) if gt_jp is not None else None
        res = out if out is not None else ForwardResult(n, self.cfg.n_hyps, want_diffmaps)
        self._pending = (coords, pix, gt, res)
        self._check(self.lib.dsac_forward_submit(self.h, n, frame0, _p(coords), _p(pix), shared, _p(gt), C.byref(res.raw)))
        return res

    def forward_wait(self):
        self._check(self.lib.dsac_forward_wait(self.h))
        pending = getattr(self, "_pending", None)
        self._pending = None
        return pending[3] if pending else None

    def forward_device(self, n, d_coords, d_pix, pix_shared=0, d_gt=None, frame0=0, stream=None):
        """dsac_forward_device with raw device pointers (ints)."""
        self._check(self.lib.dsac_forward_device(self.h, n, frame0, C.c_void_p(d_coords), C.c_void_p(d_pix), pix_shared,
                                                 C.c_void_p(d_gt) if d_gt else None, C.c_void_p(stream) if stream else None))

    def fetch(self, n, want_diffmaps=False, stream=None, out=None):
        res = out if out is not None else ForwardResult(n, self.cfg.n_hyps, want_diffmaps)
        self._check(self.lib.dsac_fetch(self.h, n, C.byref(res.raw), C.c_void_p(stream) if stream else None))
        return res

    def device_view(self):
        v = DeviceView()
        self._check(self.lib.dsac_device_view_get(self.h, C.byref(v)))
        return v

    def set_score_hook(self, fn):
        """fn(d_diffmaps:int, n:int, H:int, d_scores:int, stream:int) -> int, called with device pointers."""
        if fn is None:
            self._hook = None
            self._check(self.lib.dsac_set_score_hook(self.h, None, None))
            return

        def tramp(dm, n, H, sc, stream, user):
            return int(fn(dm or 0, n, H, sc or 0, stream or 0))

        self._hook = SCORE_HOOK(tramp)
        self._check(self.lib.dsac_set_score_hook(self.h, C.cast(self._hook, C.c_void_p), None))

    def set_score_backward_hook(self, fn):
        """fn(d_diffmaps:int, d_score_grads:int, n:int, H:int, d_diffmap_grads:int, stream:int) -> int (device pointers):
        the seam's adjoint (lua_calls.h:312-341); must be registered together with the forward hook."""
        if fn is None:
            self._bw_hook = None
            self._check(self.lib.dsac_set_score_backward_hook(self.h, None, None))
            return

        def tramp(dm, sg, n, H, out, stream, user):
            return int(fn(dm or 0, sg or 0, n, H, out or 0, stream or 0))

        self._bw_hook = SCORE_BACKWARD_HOOK(tramp)
        self._check(self.lib.dsac_set_score_backward_hook(self.h, C.cast(self._bw_hook, C.c_void_p), None))

    def forward_dsac(self, coords, pix, gt_jp, random_draw=True, frame0=0, want_inlier_maps=False):
        """dsac_forward_dsac: the DSAC / RANSAC variant (draw + refine all hypotheses + expected loss)."""
        coords = np.ascontiguousarray(coords, np.int16).reshape(-1, N, 3)
        n = coords.shape[0]
        pix = np.ascontiguousarray(pix, np.int32)
        shared = 1 if pix.size == N * 2 else 0
        gt = np.ascontiguousarray(gt_jp, np.float64).reshape(n, 12) if gt_jp is not None else None
        res = DsacResult(n, self.cfg.n_hyps, want_inlier_maps)
        self._check(self.lib.dsac_forward_dsac(self.h, n, frame0, _p(coords), _p(pix), shared, _p(gt), int(random_draw),
                                               C.byref(res.raw)))
        return res

    def backward(self, coords, pix, gt_jp, full=True):
        coords = np.ascontiguousarray(coords, np.int16).reshape(-1, N, 3)
        n = coords.shape[0]
        pix = np.ascontiguousarray(pix, np.int32)
        shared = 1 if pix.size == N * 2 else 0
        gt = np.ascontiguousarray(gt_jp, np.float64).reshape(n, 12)
        res = BackwardResult(n, self.cfg.n_hyps, full)
        self._check(self.lib.dsac_backward(self.h, n, _p(coords), _p(pix), shared, _p(gt), C.byref(res.raw)))
        return res

    def backward_dsac(self, n):
        """dsac_backward_dsac: gradients of the expected loss of the DSAC variant; follows forward_dsac over n frames."""
        res = BackwardDsacResult(n, self.cfg.n_hyps)
        self._check(self.lib.dsac_backward_dsac(self.h, n, C.byref(res.raw)))
        return res

    def gather_patches_device(self, n, d_frames, width, height, d_pix, pix_shared, d_patches, mean=127.0, d_status=None, stream=None):
        """dsac_gather_patches_device (raw device pointers as ints): BGR frames -> normalised 3x42x42 CNN patches."""
        self._check(self.lib.dsac_gather_patches_device(self.h, n, C.c_void_p(d_frames), width, height, C.c_void_p(d_pix), pix_shared,
                                                        C.c_float(mean), C.c_void_p(d_patches), C.c_void_p(d_status) if d_status else None,
                                                        C.c_void_p(stream) if stream else None))

    def coords_from_prediction_device(self, n, d_pred, d_coords, stream=None):
        """dsac_coords_from_prediction_device: float metres -> int16 millimetres (cv::saturate_cast<short>)."""
        self._check(self.lib.dsac_coords_from_prediction_device(self.h, n, C.c_void_p(d_pred), C.c_void_p(d_coords),
                                                                C.c_void_p(stream) if stream else None))

    def kabsch(self, a, b):
        a = np.ascontiguousarray(a, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        n, m = a.shape[0], a.shape[1]
        R = np.zeros((n, 9))
        t = np.zeros((n, 3))
        self._check(self.lib.dsac_kabsch(self.h, n, m, _p(a), _p(b), _p(R), _p(t)))
        return R.reshape(n, 3, 3), t
