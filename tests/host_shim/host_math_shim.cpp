// Host build of the PRODUCT's device arithmetic headers (dsac_b200/csrc/*.cuh compiled with
// DSAC_HOST_ONLY) so that `-m "not gpu"` tests can compare them with the oracle on a box
// without a GPU.  Test infrastructure only; never shipped or used by the product path.
#define DSAC_HOST_ONLY 1
#include "../../dsac_b200/csrc/pose_math.cuh"

extern "C" {
int shim_minimal_set(const float* obj, const float* img, double f, double cx, double cy, int thr, double* rvec,
                     double* tvec, int* fragile) {
    bool fr = false;
    bool ok = dsac::minimal_set_hypothesis(obj, img, f, cx, cy, thr, rvec, tvec, &fr);
    *fragile = fr ? 1 : 0;
    return ok ? 1 : 0;
}
int shim_p3p_best(const float* obj, const float* img, double f, double cx, double cy, double* R, double* t, double* e2) {
    dsac::P3PProblem pr;
    for (int i = 0; i < 4; i++) {
        pr.mu[i] = dsac::p3p_pixel(img[i * 2], cx, f);
        pr.mv[i] = dsac::p3p_pixel(img[i * 2 + 1], cy, f);
        for (int k = 0; k < 3; k++) pr.X[i][k] = obj[i * 3 + k];
    }
    return dsac::p3p_best(pr, f, cx, cy, R, t, e2);
}
void shim_rodrigues_v2m(const double* r, double* R) { dsac::rodrigues_v2m(r, R); }
void shim_rodrigues_m2v(const double* R, double* r) { dsac::rodrigues_m2v(R, r); }
}

#include <vector>
#include "../../dsac_b200/csrc/sampler.cuh"
extern "C" {
// product MT19937 + Lemire + candidate parser, run sequentially on the host
void shim_mt_raw(uint32_t seed, int n, uint32_t* out) {
    uint32_t mt[dsac::MT_N];
    dsac::mt_seed(mt, seed);
    int i = 0;
    while (i < n) {
        dsac::mt_twist_all_seq(mt);
        for (int k = 0; k < dsac::MT_N && i < n; k++) out[i++] = dsac::mt_temper(mt[k]);
    }
}
void shim_candidates(uint32_t seed, uint32_t skip, int n, int32_t* cells, uint32_t* used) {
    size_t words = (size_t)skip + (size_t)n * 16 + 4096;
    size_t cap = 1;
    while (cap < words) cap <<= 1;
    std::vector<uint32_t> buf(cap);
    shim_mt_raw(seed, (int)words, buf.data());
    dsac::WordRing ring{buf.data(), (uint32_t)(cap - 1)};
    uint32_t pos = skip;
    for (int k = 0; k < n; k++) {
        int c[4];
        uint32_t u = dsac::parse_candidate(ring, pos, (uint32_t)words, c);
        for (int j = 0; j < 4; j++) cells[k * 4 + j] = c[j];
        used[k] = u;
        pos += u;
    }
}
void shim_stream_chunk(int H, int T, int s, int* h0, int* cnt) { dsac::stream_chunk(H, T, s, h0, cnt); }
}
