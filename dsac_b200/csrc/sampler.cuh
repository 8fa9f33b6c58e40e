// sampler.cuh -- the minimal-set sampler's random stream, reproduced on the device.
//
// The reference draws every minimal set from a per-OpenMP-thread std::mt19937
// (thread_rand.cpp:40-69) through fresh std::uniform_int_distribution<int>(0,39) objects
// (irand, thread_rand.cpp:95-98; call sites cnn_softam.h:1024-1025), re-drawing a cell
// that was already chosen (cnn_softam.h:1027-1031).  "Bit-exact sampled indices" therefore
// means reproducing (a) the MT19937 output sequence and (b) libstdc++ 13's Lemire
// down-scaling (bits/uniform_int_dist.h:_S_nd, 32-bit generator -> 64-bit product).
// Both are restated here from their published definitions (Matsumoto & Nishimura 1998;
// Lemire 2019) and pinned against libstdc++ in tests/.
//
// A stream is parsed into candidates: candidate k starts where candidate k-1 stopped.
// Nearly all candidates consume exactly 8 words, so a block parses 256 candidates in
// parallel by iterating "assume previous extras -> parse -> prefix-sum the extras" to a
// fixed point (usually 1-2 passes).
#pragma once
#include "pose_math.cuh"

namespace dsac {

constexpr int MT_N = 624;
constexpr int MT_M = 397;

DSAC_HD uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

DSAC_HD uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((0u - (nxt & 1u)) & 0x9908b0dfu);
}

DSAC_HD void mt_seed(uint32_t* mt, uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < MT_N; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}

// Sequential twist (host tests / reference for the block-parallel version below).
DSAC_HD void mt_twist_all_seq(uint32_t* mt) {
    for (int k = 0; k < MT_N - MT_M; k++) mt[k] = mt_twist(mt[k], mt[k + 1], mt[k + MT_M]);
    for (int k = MT_N - MT_M; k < MT_N - 1; k++) mt[k] = mt_twist(mt[k], mt[k + 1], mt[k + MT_M - MT_N]);
    mt[MT_N - 1] = mt_twist(mt[MT_N - 1], mt[0], mt[MT_M - 1]);
}

// Block-parallel regeneration of the whole state (k_sample): in the in-place order k = 0..623 the word
//   new[k] = (k < 227 ? old[k+397] : new[k-227]) ^ twist(old[k], k < 623 ? old[k+1] : new[0])
// so thread t < 227 owns k = t, t+227, t+454 (the last only for t < 170): its second and third word need its own previous
// result and OLD neighbours only (k = 623 = 169+454 needs new[0], recomputed here from old words).  Returns the number of
// words produced (x[w] is new[t + 227 w]); `so` is the old state, which the caller keeps apart from the new one.
DSAC_HD int mt_regenerate_words(const uint32_t* so, int t, uint32_t x[3]) {
    x[0] = mt_twist(so[t], so[t + 1], so[t + MT_M]);
    x[1] = mt_twist(so[t + (MT_N - MT_M)], so[t + (MT_N - MT_M) + 1], x[0]);
    const int k = t + 2 * (MT_N - MT_M);
    if (k >= MT_N) return 2;
    const uint32_t nxt = (k < MT_N - 1) ? so[k + 1] : mt_twist(so[0], so[1], so[MT_M]);
    x[2] = mt_twist(so[k], nxt, x[1]);
    return 3;
}

// The same regeneration with the thread's own three words of the OLD state in registers (own[s] = old[t + 227 s]: they are
// what the thread produced one regeneration earlier) and without a branch: all loads are issued first, the word after the
// state's last one (needed by thread 169 only) is recomputed by every thread and selected.  A lone CTA twists a block in 230
// cycles this way against 388 with mt_regenerate_words (tools/micro/twist_bench.cu).  x[2] is meaningful for t < 170.
DSAC_HD void mt_twist3(const uint32_t* so, int t, const uint32_t own[3], uint32_t x[3]) {
    const int k3 = (t + 2 * (MT_N - MT_M) < MT_N - 1) ? t + 2 * (MT_N - MT_M) : MT_N - 1;
    const uint32_t a1 = so[t + 1], af = so[t + MT_M];
    const uint32_t b1 = so[t + (MT_N - MT_M) + 1];
    const uint32_t c1 = so[(k3 + 1 < MT_N - 1) ? k3 + 1 : MT_N - 1];
    const uint32_t n0 = mt_twist(so[0], so[1], so[MT_M]);
    x[0] = mt_twist(own[0], a1, af);
    x[1] = mt_twist(own[1], b1, x[0]);
    x[2] = mt_twist(own[2], (k3 == MT_N - 1) ? n0 : c1, x[1]);
}

// A window of tempered stream words addressed by absolute stream position.
struct WordRing {
    const uint32_t* buf;
    uint32_t mask;  // size-1, size a power of two
    DSAC_HD uint32_t at(uint32_t pos) const { return buf[pos & mask]; }
};

// One libstdc++ uniform_int_distribution<int>(0, 39) draw starting at stream position
// *pos; advances *pos past the words it consumed.
DSAC_HD int lemire40(const WordRing& ring, uint32_t* pos) {
    const uint32_t range = DSAC_GRID_CONST;
    uint64_t product = (uint64_t)ring.at((*pos)++) * range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        const uint32_t threshold = (0u - range) % range;  // 2^32 mod 40 = 16
        while (low < threshold) {
            product = (uint64_t)ring.at((*pos)++) * range;
            low = (uint32_t)product;
        }
    }
    return (int)(product >> 32);
}

// Parses the candidate starting at `start`: 4 distinct cells (x,y drawn in that order),
// returns the number of words consumed.  Never reads at or beyond `limit`
// (returns 0 if the candidate does not fit).
DSAC_HD uint32_t parse_candidate(const WordRing& ring, uint32_t start, uint32_t limit, int cell[4]) {
    uint32_t pos = start;
    int n = 0;
    while (n < 4) {
        if (limit - pos < 2 + 2) return 0;  // keep a safety margin for Lemire re-draws
        int x = lemire40(ring, &pos);
        int y = lemire40(ring, &pos);
        int c = y * DSAC_GRID_CONST + x;
        bool dup = false;
        for (int j = 0; j < 4; j++)
            if (j < n && cell[j] == c) dup = true;
        if (dup) continue;
        cell[n++] = c;
    }
    return pos - start;
}

// Length in (x, y) pairs of the candidate that starts at pair sp, for a window without rejected draws (every word is a
// value, so candidates are sequences of pairs): pairs are taken until 4 distinct ones are found.  The first 8 pairs are
// fetched with independent loads (a candidate longer than that needs >= 4 repeats); returns -1 if the window ends first.
DSAC_HD int cand_pairs_len(const unsigned short* pr16, int sp, int limit) {
    if (sp + 8 <= limit) {
        unsigned v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = pr16[sp + i];
        unsigned c0 = v[0], c1 = 0x10000u, c2 = 0x10000u;   // 0x10000: no 16-bit pair value
        int n = 1;
#pragma unroll
        for (int i = 1; i < 8; i++) {
            const bool dup = (v[i] == c0) | (v[i] == c1) | (v[i] == c2);
            if (!dup) {
                if (n == 3) return i + 1;
                if (n == 1) c1 = v[i]; else c2 = v[i];
                n++;
            }
        }
    }
    unsigned c[4];
    int n = 0, q = sp;
    while (n < 4) {
        if (q >= limit) return -1;
        const unsigned v = pr16[q++];
        bool dup = false;
#pragma unroll
        for (int j = 0; j < 3; j++)
            if (j < n && c[j] == v) dup = true;
        if (!dup) { c[n < 3 ? n : 3] = v; n++; }
    }
    return q - sp;
}

// libgomp static schedule of `#pragma omp parallel for` over h (cnn_softam.h:1010):
// stream s of T owns a contiguous chunk; the first H%T streams get one more.
DSAC_HD void stream_chunk(int H, int T, int s, int* h0, int* cnt) {
    int q = H / T, r = H % T;
    *cnt = q + (s < r ? 1 : 0);
    *h0 = s * q + (s < r ? s : r);
}

}  // namespace dsac
