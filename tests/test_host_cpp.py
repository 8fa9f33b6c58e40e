"""The C++ host mirror of the reference surface (dsac_b200/host/: Hypothesis, GlobalProperties, ThreadRand,
maxLoss, conventions) and the two drivers (apps/)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APPS = os.path.join(ROOT, "apps")


@pytest.fixture(scope="module")
def apps(engine_mod):
    subprocess.check_call(["make", "-C", APPS, "-s"])
    return APPS


def test_host_selftest(apps):
    out = subprocess.run([os.path.join(apps, "host_selftest")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("ok")


def test_reference_surface_names_present():
    """Same class / function names as the reference's headers."""
    h = open(os.path.join(ROOT, "dsac_b200", "host", "Hypothesis.h")).read()
    for name in ("getRodVecAndTrans", "calcAngularDistance", "calcRigidBodyTransform", "getInvRotation", "invTransform",
                 "getTransformation", "operator*", "operator/", "setRotation", "setTranslation", "refine"):
        assert name in h, name
    p = open(os.path.join(ROOT, "dsac_b200", "host", "properties.h")).read()
    for name in ("ransacIterations", "ransacRefinementIterations", "ransacBatchSize", "ransacSubSample",
                 "ransacInlierThreshold2D", "getCamMat", "parseCmdLine", "parseConfig", "readArguments", "GlobalProperties"):
        assert name in p, name
    c = open(os.path.join(ROOT, "dsac_b200", "host", "cnn_softam.h")).read()
    for name in ("processImage", "refAvgHyp", "sfEntropy", "sampledPoints", "inlierMap", "pixelIdxs", "stochasticSubSample"):
        assert name in c, name


@pytest.mark.gpu
def test_test_driver_writes_reference_logs(apps, tmp_path, oracle, engine_mod):
    """BASELINE config 5 in miniature: the test driver on a 48-frame synthetic trajectory."""
    out = subprocess.run([os.path.join(apps, "test_ransac_softam"), "-frames", "48", "-batch", "16"], cwd=tmp_path,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    errs = np.loadtxt(tmp_path / "ransac_test_errors_obj_model_init.net_rdraw1_softam.txt")
    summ = np.loadtxt(tmp_path / "ransac_test_loss_obj_model_init.net_rdraw1_softam.txt")
    assert errs.shape == (48, 10) and summ.shape == (7,)
    assert summ[0] > 0.9                                   # accuracy (5 cm / 5 deg)
    assert np.isclose(summ[1], errs[:, 0].mean()) and np.isclose(summ[5], np.sort(errs[:, 3])[24])
    # per-frame numbers equal the oracle's on the same synthetic trajectory (frames 0, 17, 47)
    coords, pix, gt_cv, gt_jp = engine_mod.synth_frames(48, traj=True)
    for f in (0, 17, 47):
        cfg = oracle.default_config(seed=1305 + f)
        fw = oracle.forward(cfg, coords[f], pix[f], gt_jp[f, :9], gt_jp[f, 9:])
        assert abs(errs[f, 2] - fw.t_err) < 5e-2 and abs(errs[f, 3] - fw.rot_err) < 1e-3 and abs(errs[f, 0] - fw.loss) < 5e-3


@pytest.mark.gpu
def test_train_driver_runs(apps, tmp_path):
    out = subprocess.run([os.path.join(apps, "train_ransac_softam"), "-frames", "2", "-rI", "32"], cwd=tmp_path,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Max gradient" in out.stdout
    log = np.loadtxt(tmp_path / "ransac_training_loss_train_obj.lua.txt")
    assert log.shape == (3, 3)


@pytest.mark.gpu
def test_dsac_variant_driver(apps, tmp_path):
    """test_ransac (core/test_ransac.cpp) in synthetic mode, arg-max and random draw."""
    for rdraw in ("0", "1"):
        out = subprocess.run([os.path.join(apps, "test_ransac"), "-frames", "12", "-batch", "6", "-rI", "64", "-rdraw", rdraw],
                             cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        errs = np.loadtxt(tmp_path / ("ransac_test_errors_obj_model_init.net_rdraw%s.txt" % rdraw))
        summ = np.loadtxt(tmp_path / ("ransac_test_loss_obj_model_init.net_rdraw%s.txt" % rdraw))
        assert errs.shape == (12, 11) and summ.shape == (7,)
        assert summ[0] > 0.8 and np.isfinite(errs).all()


@pytest.mark.gpu
def test_dsac_variant_train_driver(apps, tmp_path):
    """SURVEY.md section 8(f) N1, backward half: train_ransac (core/train_ransac.cpp) in synthetic mode."""
    out = subprocess.run([os.path.join(apps, "train_ransac"), "-frames", "2", "-rI", "32"], cwd=tmp_path,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Max gradient" in out.stdout and "refinements differentiated" in out.stdout
    log = np.loadtxt(tmp_path / "ransac_training_loss_train_obj.lua.txt")
    assert log.shape == (3, 3) and np.isfinite(log).all()
