// thread_rand.h -- ThreadRand with the surface of /root/reference/core/thread_rand.h: one std::mt19937
// per stream seeded seed+i, drawn through fresh libstdc++ distributions.  The reference keys streams by
// OpenMP thread id; here the stream id is explicit (tid), which is what the engine's contract uses.
#pragma once
#include <random>
#include <vector>

class ThreadRand {
public:
    static int irand(int min, int max, int tid = 0);          // [min, max]
    static double drand(double min, double max, int tid = 0);
    static double dgauss(double mean, double stdDev, int tid = 0);
    static void forceInit(unsigned seed, unsigned nStreams = 1);
    static std::vector<std::mt19937> generators;
    static bool initialised;
private:
    static void init(unsigned seed = 1305, unsigned nStreams = 1);
};

int irand(int incMin, int excMax, int tid = 0);               // thread_rand.cpp:95-98
double drand(double incMin, double incMax, int tid = 0);
int igauss(int mean, int stdDev, int tid = 0);
double dgauss(double mean, double stdDev, int tid = 0);
