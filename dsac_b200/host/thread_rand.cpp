// thread_rand.cpp -- see thread_rand.h (behaviour of /root/reference/core/thread_rand.cpp:31-112).
#include "thread_rand.h"

std::vector<std::mt19937> ThreadRand::generators;
bool ThreadRand::initialised = false;

void ThreadRand::forceInit(unsigned seed, unsigned nStreams) {
    initialised = false;
    generators.clear();
    init(seed, nStreams);
}

void ThreadRand::init(unsigned seed, unsigned nStreams) {
    if (initialised) return;
    for (unsigned i = 0; i < nStreams; i++) {
        generators.push_back(std::mt19937());
        generators[i].seed(i + seed);
    }
    initialised = true;
}

int ThreadRand::irand(int min, int max, int tid) {
    std::uniform_int_distribution<int> dist(min, max);
    if (!initialised) init();
    return dist(generators[tid]);
}

double ThreadRand::drand(double min, double max, int tid) {
    std::uniform_real_distribution<double> dist(min, max);
    if (!initialised) init();
    return dist(generators[tid]);
}

double ThreadRand::dgauss(double mean, double stdDev, int tid) {
    std::normal_distribution<double> dist(mean, stdDev);
    if (!initialised) init();
    return dist(generators[tid]);
}

int irand(int incMin, int excMax, int tid) { return ThreadRand::irand(incMin, excMax - 1, tid); }
double drand(double incMin, double incMax, int tid) { return ThreadRand::drand(incMin, incMax, tid); }
int igauss(int mean, int stdDev, int tid) { return (int)ThreadRand::dgauss(mean, stdDev, tid); }
double dgauss(double mean, double stdDev, int tid) { return ThreadRand::dgauss(mean, stdDev, tid); }
