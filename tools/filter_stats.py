"""CPU check of the K1 conservative filter (host build of the product headers): over n candidates per frame of the
synthetic benchmark frames, count accepted / flagged-for-full-solve / MISSED (accepted by the exact check but
rejected by the filter -- must be 0)."""
import os, sys, ctypes as C, subprocess, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsac_b200 import engine as E

def main(nf=8, n=250000, extra=()):
    d = os.path.join(ROOT, "tests", "host_shim"); so = os.path.join(d, "libhost_shim_stats.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", *extra, "-o", so, os.path.join(d, "host_math_shim.cpp")])
    shim = C.CDLL(so)
    coords, pix, _, _ = E.synth_frames(nf)
    tot = np.zeros(6, np.int64); t0 = time.time()
    for i in range(nf):
        out = (C.c_longlong * 6)()
        shim.shim_filter_stats(coords[i].ctypes.data_as(C.c_void_p), pix[i].ctypes.data_as(C.c_void_p), C.c_uint32(1305 + i),
                               C.c_uint32(6400 if i == 0 else 0), C.c_int(n), C.c_double(525), C.c_double(320), C.c_double(240), C.c_int(10), out)
        tot += np.array(list(out))
    print(dict(candidates=int(tot[0]), accepted=int(tot[1]), flagged=int(tot[2]), missed=int(tot[4]),
               flag_pct=round(100 * tot[2] / tot[0], 3), accept_pct=round(100 * tot[1] / tot[0], 3), secs=round(time.time() - t0, 1)))
    rs = (C.c_longlong * 24)(); shim.shim_filter_reasons(rs)
    if rs[0] >= 0: print("guards fired:", {k: rs[k] for k in range(24) if rs[k]})
    shim.shim_filter_fp32_maxdev.restype = C.c_double
    if shim.shim_filter_fp32_maxdev() >= 0: print("largest |fp32 - fp64| pixel error of the 4th-point stage (well-conditioned roots): %.3g px" % shim.shim_filter_fp32_maxdev())
    os.remove(so)
    return tot

if __name__ == "__main__":
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
    main(nf, n, tuple(sys.argv[3:]))
