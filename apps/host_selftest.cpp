// host_selftest -- CPU-only checks of the C++ host mirror (Hypothesis, GlobalProperties, ThreadRand, conventions).
// Run by tests/test_host_cpp.py; prints "ok" and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "../dsac_b200/host/cnn_softam.h"
#include "../dsac_b200/host/read_data.h"
#include "../dsac_b200/host/thread_rand.h"

using namespace cvlite;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, const char* argv[]) {
    // `host_selftest pose <file>`: the ground-truth pose of a 7-Scenes pose file (read_data.cpp:69-133 + Hypothesis(info)),
    // printed with 17 digits for tests/test_oracle_vs_ref.py (translation.txt is looked up in the current directory)
    if (argc == 3 && std::string(argv[1]) == "pose") {
        jp::info_t info;
        if (!jp::readData(argv[2], info)) { std::printf("cannot read %s\n", argv[2]); return 1; }
        Hypothesis h(info);
        for (int i = 0; i < 9; i++) std::printf("%.17g ", h.getRotation()(i / 3, i % 3));
        std::printf("%.17g %.17g %.17g\n", h.getTranslation().x, h.getTranslation().y, h.getTranslation().z);
        return 0;
    }
    // Hypothesis: Rodrigues 6-vector round trip, inverse, composition, angular distance
    Hypothesis h(std::vector<double>{0.3, -0.2, 0.5, 100, -50, 2000});
    std::vector<double> v = h.getRodVecAndTrans();
    CHECK(std::fabs(v[0] - 0.3) < 1e-12 && std::fabs(v[2] - 0.5) < 1e-12 && v[5] == 2000);
    Hypothesis hi = h.getInv();
    Hypothesis id = h * hi;
    CHECK(std::fabs(id.getTranslation().x) < 1e-9 && std::fabs(id.getRotation()(0, 0) - 1) < 1e-12);
    Hypothesis q = h / h;
    CHECK(std::fabs(q.getRotation()(1, 1) - 1) < 1e-12);
    CHECK(h.calcAngularDistance(h) < 1e-5);
    Hypothesis h2(std::vector<double>{0.3, -0.2, 0.5 + 0.01, 100, -50, 2000});
    CHECK(std::fabs(h.calcAngularDistance(h2) - 0.01 * 180 / 3.14159265358979323846) < 0.2);
    Point3d p(10, 20, 30), tp = h.transform(p), back = h.invTransform(tp);
    CHECK(std::fabs(back.x - 10) < 1e-9 && std::fabs(back.z - 30) < 1e-9);
    CHECK(norm(Hypothesis(std::vector<double>{1e-7, 0, 0, 0, 0, 0}).getTranslation()) == 0);   // tiny rotation -> identity (Hypothesis.cpp:92)
    // Kabsch from correspondences
    std::vector<std::pair<Point3d, Point3d>> pts;
    for (int i = 0; i < 6; i++) {
        Point3d a(13.0 * i - 20, 7.0 * ((i * i) % 5), -11.0 * i + 3 * (i % 2));
        pts.push_back({a, h.transform(a)});
    }
    Hypothesis k(pts);
    CHECK(k.calcAngularDistance(h) < 1e-5 && norm(k.getTranslation() - h.getTranslation()) < 1e-6);
    // conventions: cv2our / our2cv round trip, det +1, maxLoss of 10 mm shift
    jp::cv_trans_t cv(Vec3d{0.2, 0.1, -0.3}, Vec3d{30, -40, 2200});
    jp::jp_trans_t jpT = jp::cv2our(cv);
    CHECK(std::fabs(determinant3(jpT.first) - 1) < 1e-12 && jpT.second.y == 40 && jpT.second.z == -2200);
    jp::cv_trans_t cv2 = jp::our2cv(jpT);
    CHECK(std::fabs(cv2.first[0] - 0.2) < 1e-12 && std::fabs(cv2.second[2] - 2200) < 1e-9);
    Hypothesis a(jpT.first, jpT.second);
    CHECK(maxLoss(a, a) < 1e-5);
    // GlobalProperties: defaults and flag parsing
    GlobalProperties* gp = GlobalProperties::getInstance();
    CHECK(gp->pP.ransacIterations == 256 && gp->pP.ransacRefinementIterations == 8 && gp->pP.ransacBatchSize == 100);
    CHECK(gp->getCamMat()(0, 0) == 525.f && gp->getCamMat()(0, 2) == 320.f && gp->getCamMat()(1, 2) == 240.f);
    CHECK(gp->readArguments({"-rI", "64", "-rT2D", "12.7", "-rSS", "0.02", "-iw", "320"}));
    CHECK(gp->pP.ransacIterations == 64 && gp->getCamMat()(0, 2) == 160.f);
    dsac_config c = gp->engineConfig(4);
    CHECK(c.n_hyps == 64 && c.thr2d == 12 && c.max_frames == 4 && std::fabs(c.sub_sample - 0.02) < 1e-7);
    CHECK(!gp->readArguments({"-nonsense", "1"}));
    // ThreadRand: thread-0 stream of seed 1305 reproduces the engine's sampling grid (stochasticSubSample contract)
    ThreadRand::forceInit(1305);
    Mat_<Point2i> s = stochasticSubSample(640, 480, 1305);
    double x0 = drand(21, 21 + 14.95f);   // first draw of the same generator: x of cell (0,0)
    CHECK((int)x0 == s(0, 0).x);
    CHECK(irand(0, 40) >= 0);
    std::printf("ok\n");
    return 0;
}
