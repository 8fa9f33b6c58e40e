// sampler_split.cuh -- K1 (minimal-set sampling, cnn_softam.h:1010-1060) as a round-based pipeline of flat kernels.
//
// The monolithic k_sample (kernels.cuh) runs one persistent CTA per (frame, stream) that alternates between integer
// work (MT19937 regeneration, candidate boundaries) and fp64 work (filter, full P3P): the fp64 pipe idles during the
// former, block barriers separate the phases and the register budget is the maximum over all of them.  Here every
// phase is its own kernel with its own register budget and its own grid:
//
//   k1_cells   once per call: per-cell record table (scene coordinate, float-rounded normalised pixel, 1/|bearing|)
//   per round r = 0 .. R-1:
//     k1_slot    one CTA per (frame, stream) -- integer only: (a) ordered selection of the previous round's accepted
//                candidates (first `quota` of the stream, in stream order), (b) size of the next round from the
//                observed acceptance rate, (c) MT19937 regeneration + uniform_int_distribution decode + candidate
//                boundaries for that many candidates, written to HBM as 4 cell indices per candidate
//     k1_filter  flat, persistent over (slot, chunk) work items: conservative fp64 filter, one thread per candidate,
//                no barriers in the loop; flagged candidates go to a global queue
//     k1_solve   flat, persistent over the flagged queue: full fp64 P3P with 4 lanes per candidate (one quartic root
//                each), the reference's float-rounded reprojection check, accepted poses to HBM
//   k1_slot (select only), then k_sample in resume mode for streams that still miss hypotheses (normally none).
//
// The candidate stream, the acceptance decisions and the order of the accepted candidates are exactly those of
// k_sample (same device functions), so the sampled indices stay bit-exact; only the scheduling differs.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace dsac {

constexpr int K1S_THREADS = 256;                 // k1_slot
constexpr int K1S_WARPS = K1S_THREADS / 32;
constexpr int K1S_SR = 2048;                     // candidates per generator window (super-round)
constexpr int K1S_WORDS = K1S_SR * 8 + 1024;     // decoded stream words buffered per window
constexpr int K1S_MAX_CAP = 16384;               // candidates per stream and round (flag queue entries hold 14 bits)
constexpr int K1S_ACC_LIST = 1024;               // accepted candidates selected per stream and round (<= quota <= DSAC_MAX_HYPS)
constexpr int K1F_THREADS = 256;                 // k1_filter
constexpr int K1F_MAX_CHUNK = 2048;              // candidates per filter work item (multiple of K1F_THREADS)
constexpr int K1V_THREADS = 128;                 // k1_solve: 32 groups of 4 lanes
constexpr int K1S_MAX_ROUNDS = 16;

struct K1SplitParams {
    SampleParams sp;            // inputs / outputs of k_sample
    K1SlotState* state;         // [slots]
    CellRec* celltab;           // [n][N]
    uint2* cells;               // [slots][cap]   4 x uint16 cell indices per candidate
    uint32_t* endw;             // [slots][cap]   stream position after the candidate
    uint32_t* accbits;          // [slots][cap/32]
    double* pose_out;           // [slots][cap][6] rvec, tvec of accepted candidates
    uint32_t* wq;               // filter work items: slot * 64 + chunk
    uint32_t* fq;               // flagged candidates: slot << 14 | index
    int* wq_n;                  // [K1S_MAX_ROUNDS]
    int* fq_n;                  // [K1S_MAX_ROUNDS]
    unsigned long long* stats_cur;   // [2] candidates, accepted hypotheses of the streams finished in this call
    const unsigned long long* stats_prev;   // the same of the previous call (prior for the first round's size)
    unsigned long long* dbg;    // [K1S_MAX_ROUNDS][4] or null: active streams, candidates, flagged, accepted per round
    int cap;                    // candidates per stream and round (multiple of 256, <= K1S_MAX_CAP)
    int chunk;                  // candidates per filter work item (multiple of 256)
    int n_slots;
    int round;                  // >= 0; select_only: no generation
    int select_only;
};

// ------------------------------------------------------------------ k1_cells
__global__ void __launch_bounds__(256) k1_cells(K1SplitParams q, int n_frames) {
    const SampleParams& p = q.sp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_frames * DSAC_N_CONST) return;
    const int frame = i / DSAC_N_CONST, c = i - frame * DSAC_N_CONST;
    const int16_t* coords = p.coords + (size_t)frame * DSAC_N_CONST * 3;
    const int32_t* pix = p.pix + (size_t)frame * p.pix_stride;
    const double inv_f = 1. / p.f, cx_f = p.cx * inv_f, cy_f = p.cy * inv_f;
    CellRec r;
    r.X = __ldg(coords + c * 3); r.Y = __ldg(coords + c * 3 + 1); r.Z = __ldg(coords + c * 3 + 2); r.pad = 0;
    r.xn = (float)(((double)__ldg(pix + c * 2) - p.cx) * inv_f);       // cv::undistortPoints rounds to float (p3p_pixel)
    r.yn = (float)(((double)__ldg(pix + c * 2 + 1) - p.cy) * inv_f);
    const double mu = r.xn * p.f + p.cx, mv = r.yn * p.f + p.cy;     // float * double
    const double u = inv_f * mu - cx_f, v = inv_f * mv - cy_f;
    r.k = rsqrt(u * u + v * v + 1);
    q.celltab[i] = r;
}

// ------------------------------------------------------------------ k1_slot
struct K1GSmem {
    uint32_t st[2 * MT_N];                       // MT19937 state, double-buffered (old / new generation)
    __align__(16) unsigned char vals[K1S_WORDS];  // Lemire value (0..39) of stream word pos+i; 255 = rejected draw
    unsigned short cand_start[K1S_SR + 512];     // word offset (from pos) where candidate i starts
    unsigned short acc_list[K1S_ACC_LIST];       // selection: accepted candidates of the round, in order
    int warp[2][K1S_WARPS];
    uint32_t newpos;
    int n_sr, any_reject, walk_fail;
    uint32_t ev[K1S_WARPS][K1_EV_CAP];
    int ev_n[K1S_WARPS];
    unsigned short brk_ci[K1_BRK_CAP], brk_cur[K1_BRK_CAP];
    int brk_n;
    int sel_total;
};

// Size of the next round: enough candidates for the hypotheses still missing, from the acceptance rate observed so
// far (cpa = candidates per accepted hypothesis) plus a margin of ~2.5 sigma of the binomial count, so that nearly
// every stream finishes in the round.  The first round has no observation of its own: it takes 80 % of what the
// streams of the previous call needed (prior), or 16 candidates per hypothesis.
__device__ __forceinline__ int k1_round_size(int quota, int acc, long long cand_base, int cap, long long cand_left, double prior_cpa) {
    double n;
    if (cand_base == 0) {
        n = (prior_cpa > 0) ? 0.8 * prior_cpa * quota : 16.0 * quota;
        if (n < 64) n = 64;
    } else if (acc == 0) {
        n = 4.0 * (double)cand_base;
    } else {
        const double need = (double)(quota - acc), cpa = (double)cand_base / (double)acc;
        n = (need + 2.5 * sqrt(need) + 1.0) * cpa;
    }
    if (n > (double)cap) n = (double)cap;
    if (n > (double)cand_left) n = (double)cand_left;
    int r = (int)n;
    if (r < 1 && cand_left > 0) r = 1;
    return r;
}

__global__ void __launch_bounds__(K1S_THREADS) k1_slot(K1SplitParams q) {
    __shared__ K1GSmem sm;
    const SampleParams& p = q.sp;
    const int tid = threadIdx.x, lane = tid & 31, warp_id = tid >> 5;
    const int s = blockIdx.x, frame = blockIdx.y;
    const int slot = frame * p.T + s;
    K1SlotState& S = q.state[slot];
    int h0, quota;
    stream_chunk(p.H, p.T, s, &h0, &quota);
    const size_t cbase = (size_t)slot * q.cap;
    const long long cand_max = p.max_candidates > 0 ? (long long)p.max_candidates : (1ll << 40);

    uint32_t pos, gen;
    int acc;
    long long cand_base;
    if (q.round == 0) {
        if (quota == 0) {
            if (tid == 0) {
                S.done = 1; S.acc = 0; S.n_round = 0; S.cand_base = 0; S.left_n = 0; S.any_reject = 0; S.overflow = 0; S.pos = 0; S.gen = 0;
                p.stream_ncand[slot] = 0;
                p.stream_endpos[slot] = 0;
            }
            return;
        }
        // stream s of global frame g: mt19937(seed + g*T + s)   (thread_rand.cpp:52 for g = 0)
        if (tid == 0) mt_seed(sm.st, p.seed + (uint32_t)((p.frame0 + frame) * (long long)p.T + s));
        pos = (s == 0) ? p.skip : 0u;
        gen = 0;
        acc = 0;
        cand_base = 0;
        if (tid == 0) { sm.any_reject = 0; sm.walk_fail = 0; }
        __syncthreads();
    } else {
        if (S.done) return;
        // ---------------- selection: the first (quota - acc) accepted candidates of the previous round, in order
        acc = S.acc;
        cand_base = S.cand_base;
        const int n_prev = S.n_round;
        const int n_words = (n_prev + 31) >> 5;          // <= 512
        const uint32_t* ab = q.accbits + (size_t)slot * (q.cap >> 5);
        uint32_t w0 = 0, w1 = 0;
        if (2 * tid < n_words) w0 = ab[2 * tid];
        if (2 * tid + 1 < n_words) w1 = ab[2 * tid + 1];
        const int c0 = __popc(w0), c1 = __popc(w1);
        int tot;
        int rank = block_excl_scan<K1S_WARPS>(c0 + c1, &tot, sm.warp[0]);
        const int room = quota - acc;
        {
            uint32_t w = w0;
            int base_i = 2 * tid * 32;
            for (int half = 0; half < 2; half++) {
                while (w) {
                    const int b = __ffs(w) - 1;
                    w &= w - 1;
                    if (rank < room) sm.acc_list[rank] = (unsigned short)(base_i + b);
                    rank++;
                }
                w = w1;
                base_i += 32;
            }
        }
        __syncthreads();
        const int n_emit = min(tot, room);
        for (int k = tid; k < n_emit; k += K1S_THREADS) {
            const int i = sm.acc_list[k];
            const size_t hi = (size_t)frame * p.H + h0 + acc + k;
            const double* po = q.pose_out + (cbase + i) * 6;
            double rvec[3] = {po[0], po[1], po[2]}, tvec[3] = {po[3], po[4], po[5]};
            double* hp = p.hyp_pose + hi * 6;
            hp[0] = rvec[0]; hp[1] = rvec[1]; hp[2] = rvec[2];
            hp[3] = tvec[0]; hp[4] = tvec[1]; hp[5] = tvec[2];
            double Rm[9];
            rodrigues_v2m(rvec, Rm);  // getDiffMap -> cv::projectPoints rebuilds R from rvec
            float4* P = reinterpret_cast<float4*>(p.hyp_P + hi * 12);
            P[0] = make_float4((float)(p.f * Rm[0]), (float)(p.f * Rm[1]), (float)(p.f * Rm[2]), (float)(p.f * tvec[0]));
            P[1] = make_float4((float)(p.f * Rm[3]), (float)(p.f * Rm[4]), (float)(p.f * Rm[5]), (float)(p.f * tvec[1]));
            P[2] = make_float4((float)Rm[6], (float)Rm[7], (float)Rm[8], (float)tvec[2]);
            const uint2 cc = q.cells[cbase + i];
            *reinterpret_cast<int4*>(p.img_idx + hi * 4) =
                make_int4((int)(cc.x & 0xffffu), (int)(cc.x >> 16), (int)(cc.y & 0xffffu), (int)(cc.y >> 16));
            p.cand_idx[hi] = (int32_t)(cand_base + i);
            if (acc + k == quota - 1) {
                p.stream_ncand[slot] = cand_base + i + 1;
                p.stream_endpos[slot] = (unsigned long long)q.endw[cbase + i];
            }
        }
        if (q.dbg && tid == 0) atomicAdd(q.dbg + (size_t)(q.round - 1) * 4 + 3, (unsigned long long)tot);
        if (acc + n_emit >= quota) {
            if (tid == 0) {
                S.done = 1; S.acc = quota; S.cand_base = cand_base + n_prev; S.n_round = 0;
                atomicAdd(q.stats_cur + 0, (unsigned long long)(cand_base + sm.acc_list[room - 1] + 1));
                atomicAdd(q.stats_cur + 1, (unsigned long long)quota);
            }
            return;
        }
        acc += n_emit;
        cand_base += n_prev;
        if (q.select_only || cand_base >= cand_max) {   // the resume pass of k_sample takes over from here
            if (tid == 0) { S.acc = acc; S.cand_base = cand_base; S.n_round = 0; }
            return;
        }
        // ---------------- restore the generator
        pos = S.pos;
        gen = S.gen;
        {
            uint32_t* half = sm.st + ((gen / MT_N) & 1u) * MT_N;
            for (int k = tid; k < MT_N; k += K1S_THREADS) half[k] = S.mt[k];
            const int ln = S.left_n;
            for (int k = tid; k < ln; k += K1S_THREADS) sm.vals[k] = S.left[k];
            if (tid == 0) { sm.any_reject = S.any_reject; sm.walk_fail = 0; }
        }
        __syncthreads();
    }

    // ---------------- size of this round
    double prior = 0;
    if (q.stats_prev && q.stats_prev[1] > 0) prior = (double)q.stats_prev[0] / (double)q.stats_prev[1];
    const int n_round = k1_round_size(quota, acc, cand_base, q.cap, cand_max - cand_base, prior);

    // ---------------- generation: windows of up to K1S_SR candidates (phases A1, A2 and E of k_sample)
    int produced = 0;
    uint32_t st_par = (gen / MT_N) & 1u;   // which half of sm.st holds the current state
    while (produced < n_round) {
        const int n_target = min(K1S_SR, n_round - produced);
        const int w_need = min(K1S_WORDS - 640, n_target * 8 + 768);   // a regeneration may overshoot by 623 words
        // (leftover words [pos, gen) are at vals[0..gen-pos))
        while ((int)(gen - pos) < w_need) {
            const uint32_t* so = sm.st + st_par * MT_N;           // old state
            uint32_t* sn = sm.st + (st_par ^ 1u) * MT_N;            // new state
            st_par ^= 1u;
            if (tid < K1_WAVE) {
                uint32_t x[3];
                const bool has3 = mt_regenerate_words(so, tid, x) == 3;
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    if (w < 2 || has3) {
                        const int k = tid + w * K1_WAVE;
                        sn[k] = x[w];
                        const int off = (int)(gen - pos) + k;
                        if (off >= 0 && off < K1S_WORDS) {
                            const uint64_t prod = (uint64_t)mt_temper(x[w]) * DSAC_GRID_CONST;
                            const bool rej = (uint32_t)prod < ((0u - DSAC_GRID_CONST) % DSAC_GRID_CONST);  // Lemire: low < 2^32 mod 40
                            sm.vals[off] = rej ? (unsigned char)255 : (unsigned char)(prod >> 32);
                            if (rej) sm.any_reject = 1;
                        }
                    }
                }
            }
            gen += MT_N;
            __syncthreads();
        }
        const int w_avail = min((int)(gen - pos), K1S_WORDS);

        // candidate boundaries (see k_sample, phase A2)
        bool a2_done = false;
        if (!sm.any_reject && !p.a2_generic) {
            const unsigned short* pr16 = reinterpret_cast<const unsigned short*>(sm.vals);
            const int scan_end = min(w_avail >> 1, n_target * 4 + 384);
            // Every thread takes 16-byte blocks (8 pairs = 4 words of 2 pairs) and compares each word with the three
            // preceding pairs two pairs at a time (zero-half test on the XOR with the shifted predecessors, ~7
            // instructions per pair); only blocks that contain a repeat (~1 in 60) go through the recording path.
            const uint4* v128 = reinterpret_cast<const uint4*>(sm.vals);
            const int n_blocks = (scan_end + 7) >> 3;
            const int bseg = ((n_blocks + K1S_WARPS - 1) / K1S_WARPS + 31) & ~31;   // blocks per warp, contiguous: events stay ordered
            const int bbeg = warp_id * bseg, bend = min(bbeg + bseg, n_blocks);
            int cnt = 0;
            for (int b0 = bbeg; b0 < bend; b0 += 32) {
                const int b = b0 + lane;
                uint32_t w[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};   // w[0..1]: the two words before the block
                if (b < bend) {
                    const uint4 cur = v128[b];
                    w[2] = cur.x; w[3] = cur.y; w[4] = cur.z; w[5] = cur.w;
                    if (b > 0) {
                        const uint2 prv = *reinterpret_cast<const uint2*>(sm.vals + 16 * b - 8);
                        w[0] = prv.x; w[1] = prv.y;
                    }
                }
                uint32_t any = 0;
#pragma unroll
                for (int j2 = 2; j2 < 6; j2++) {
                    const uint32_t s1 = __funnelshift_l(w[j2 - 1], w[j2], 16);       // (pair 2j-1, pair 2j)
                    const uint32_t s3 = __funnelshift_l(w[j2 - 2], w[j2 - 1], 16);   // (pair 2j-3, pair 2j-2)
                    const uint32_t x1 = w[j2] ^ s1, x2 = w[j2] ^ w[j2 - 1], x3 = w[j2] ^ s3;
                    any |= ((x1 - 0x00010001u) & ~x1) | ((x2 - 0x00010001u) & ~x2) | ((x3 - 0x00010001u) & ~x3);
                }
                const bool has = (b < bend) && (any & 0x80008000u) != 0u;
                if (__any_sync(0xffffffffu, has)) {   // rare: record (pair index, distance) in pair order
                    uint32_t evs[8];
                    int ne = 0;
                    if (has) {
#pragma unroll
                        for (int h2 = 0; h2 < 8; h2++) {
                            const int k = 8 * b + h2;
                            const uint32_t wj = w[2 + (h2 >> 1)], wp1 = w[1 + (h2 >> 1)], wp2 = w[(h2 >> 1)];
                            uint32_t pk, q1, q2, q3;
                            if (h2 & 1) { pk = wj >> 16; q1 = wj & 0xffffu; q2 = wp1 >> 16; q3 = wp1 & 0xffffu; }
                            else { pk = wj & 0xffffu; q1 = wp1 >> 16; q2 = wp1 & 0xffffu; q3 = wp2 >> 16; }
                            const uint32_t d = (pk == q1) ? 1u : (pk == q2) ? 2u : (pk == q3) ? 3u : 0u;
                            if (d && k < scan_end) evs[ne++] = ((uint32_t)k << 2) | d;
                        }
                    }
                    int incl = ne;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += o;
                    }
                    const int tot = __shfl_sync(0xffffffffu, incl, 31);
                    const int at = cnt + incl - ne;
                    for (int e = 0; e < ne; e++)
                        if (at + e < K1_EV_CAP) sm.ev[warp_id][at + e] = evs[e];
                    cnt += tot;
                }
            }
            if (lane == 0) sm.ev_n[warp_id] = cnt;
            __syncthreads();
            if (tid == 0) {
                int cur = 0, ci = 0, nb = 1, stop_at = -1;
                bool fail = false;
                sm.brk_ci[0] = 0; sm.brk_cur[0] = 0;
                for (int w = 0; w < K1S_WARPS && !fail && stop_at < 0; w++) {
                    const int n = sm.ev_n[w];
                    if (n > K1_EV_CAP) { fail = true; break; }
                    for (int e = 0; e < n; e++) {
                        const uint32_t evv = sm.ev[w][e];
                        const int k = (int)(evv >> 2), d = (int)(evv & 3u);
                        if (k < cur) continue;                 // inside a candidate already parsed
                        const int j = (k - cur) >> 2, spp = cur + 4 * j;
                        if (k - d < spp) continue;             // the equal pair belongs to the previous candidate
                        if (ci + j >= n_target) { stop_at = n_target; break; }
                        const int np = cand_pairs_len(pr16, spp, scan_end);
                        if (np < 0) { stop_at = ci + j; break; }   // window ends inside this candidate
                        if (nb >= K1_BRK_CAP) { fail = true; break; }
                        ci += j + 1;
                        cur = spp + np;
                        sm.brk_ci[nb] = (unsigned short)ci;
                        sm.brk_cur[nb] = (unsigned short)cur;
                        nb++;
                    }
                }
                int n_ok = ci + ((scan_end - cur) >> 2);       // clean 4-pair candidates after the last break
                if (stop_at >= 0) n_ok = min(n_ok, stop_at);
                sm.n_sr = min(n_ok, n_target);
                sm.brk_n = nb;
                sm.walk_fail = fail ? 1 : 0;
            }
            __syncthreads();
            if (!sm.walk_fail) {
                const int n_ok = sm.n_sr, nb = sm.brk_n;
                int m = 0;
                for (int i = tid; i <= n_ok; i += K1S_THREADS) {
                    while (m + 1 < nb && (int)sm.brk_ci[m + 1] <= i) m++;
                    sm.cand_start[i] = (unsigned short)(2 * ((int)sm.brk_cur[m] + 4 * (i - (int)sm.brk_ci[m])));
                }
                a2_done = true;
            }
        }
        if (!a2_done) {
            int rp = 0, n_done = 0, par = 0;   // word offset of the chunk, candidates placed so far
            bool out_of_words = false;
            while (n_done < n_target && !out_of_words) {
                const int n_chunk = min(K1S_THREADS, n_target - n_done);
                int extra = 0, start = 0, qn = 0;
                for (;;) {
                    int tot;
                    const int excl = block_excl_scan<K1S_WARPS>(extra, &tot, sm.warp[par]);
                    par ^= 1;
                    start = rp + 8 * tid + excl;
                    int cells[4];
                    qn = (tid < n_chunk) ? cand_parse_fast(sm.vals, start, w_avail, cells) : start + 8;
                    const int ne = (qn < 0) ? 0 : (qn - start) - 8;
                    const int changed = (ne != extra);
                    extra = ne;
                    if (!__syncthreads_or(changed)) break;
                }
                int bad = (tid < n_chunk && qn < 0) ? tid : K1S_THREADS;
#pragma unroll
                for (int off = 16; off; off >>= 1) bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, off));
                if (lane == 0) sm.warp[par][tid >> 5] = bad;
                __syncthreads();
                int n_ok = n_chunk;
#pragma unroll
                for (int w = 0; w < K1S_WARPS; w++) n_ok = min(n_ok, sm.warp[par][w]);
                par ^= 1;
                if (tid < n_ok) sm.cand_start[n_done + tid] = (unsigned short)start;
                if (tid == n_ok - 1) sm.newpos = (uint32_t)qn;          // end of the last complete candidate
                __syncthreads();
                if (n_ok > 0) rp = (int)sm.newpos;
                n_done += n_ok;
                out_of_words = (n_ok < n_chunk);
            }
            if (tid == 0) {
                sm.cand_start[n_done] = (unsigned short)rp;
                sm.n_sr = n_done;
            }
        }
        __syncthreads();
        const int n_sr = sm.n_sr;

        // the window's candidates to HBM
        for (int i = tid; i < n_sr; i += K1S_THREADS) {
            int cells[4];
            cand_parse_fast(sm.vals, sm.cand_start[i], w_avail, cells);
            q.cells[cbase + produced + i] = make_uint2((uint32_t)cells[0] | ((uint32_t)cells[1] << 16), (uint32_t)cells[2] | ((uint32_t)cells[3] << 16));
            q.endw[cbase + produced + i] = pos + (uint32_t)sm.cand_start[i + 1];
        }

        // advance the stream: the unread tail of the window moves to the front
        {
            const int consumed = sm.cand_start[n_sr];
            const int left = w_avail - consumed;
            unsigned char keep[K1S_LEFT_CAP / K1S_THREADS];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K1S_LEFT_CAP / K1S_THREADS; k++) {
                const int i = tid + k * K1S_THREADS;
                keep[k] = (i < left) ? sm.vals[consumed + i] : (unsigned char)0;
            }
            __syncthreads();
            bool rej_left = false;
#pragma unroll
            for (int k = 0; k < K1S_LEFT_CAP / K1S_THREADS; k++) {
                const int i = tid + k * K1S_THREADS;
                if (i < left) {
                    sm.vals[i] = keep[k];
                    rej_left |= (keep[k] == 255);
                }
            }
            if (tid == 0) { sm.any_reject = 0; sm.walk_fail = 0; }
            __syncthreads();
            if (rej_left) sm.any_reject = 1;   // a rejected draw carried over into the next window
            if (left > K1S_LEFT_CAP && tid == 0) S.overflow = 1;
            pos += (uint32_t)consumed;
            produced += n_sr;
            __syncthreads();
        }
        if (n_sr == 0) break;   // (cannot happen: every window holds at least one candidate)
    }

    // ---------------- hand the round over: state, work items, cleared accept bits
    {
        const int left = min((int)(gen - pos), K1S_LEFT_CAP);
        const uint32_t* half = sm.st + ((gen / MT_N) & 1u) * MT_N;
        for (int k = tid; k < MT_N; k += K1S_THREADS) S.mt[k] = half[k];
        for (int k = tid; k < left; k += K1S_THREADS) S.left[k] = sm.vals[k];
        uint32_t* ab = q.accbits + (size_t)slot * (q.cap >> 5);
        for (int k = tid; k < ((produced + 31) >> 5); k += K1S_THREADS) ab[k] = 0u;
        if (tid == 0) {
            S.pos = pos; S.gen = gen; S.acc = acc; S.cand_base = cand_base; S.n_round = produced; S.done = 0;
            S.left_n = left; S.any_reject = sm.any_reject;
            if (q.round == 0) S.overflow = 0;
            const int n_items = (produced + q.chunk - 1) / q.chunk;
            if (n_items > 0) {
                const int at = atomicAdd(q.wq_n + q.round, n_items);
                for (int c = 0; c < n_items; c++) q.wq[at + c] = (uint32_t)slot * 64u + (uint32_t)c;
            }
            if (q.dbg) {
                atomicAdd(q.dbg + (size_t)q.round * 4 + 0, 1ull);
                atomicAdd(q.dbg + (size_t)q.round * 4 + 1, (unsigned long long)produced);
            }
        }
    }
}

// ------------------------------------------------------------------ k1_filter
// The filter's inputs straight from the cell table: bearings of points 0..2 from the stored 1/|(u, v, 1)|.
__device__ __forceinline__ bool k1_filter_candidate(const CellRec* cell, const int cells[4], double f, double cx, double cy,
                                                    double inv_f, double cx_f, double cy_f, double thr) {
    double bear[3][3], X[4][3], mu3 = 0, mv3 = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const CellRec r = cell[cells[j]];
        const double mu = r.xn * f + cx, mv = r.yn * f + cy;
        X[j][0] = (double)r.X; X[j][1] = (double)r.Y; X[j][2] = (double)r.Z;
        if (j < 3) {
            const double u = inv_f * mu - cx_f, v = inv_f * mv - cy_f;
            bear[j][0] = u * r.k; bear[j][1] = v * r.k; bear[j][2] = r.k;
        } else {
            mu3 = mu; mv3 = mv;
        }
    }
    return p3p_quick_core(bear, X, mu3, mv3, f, cx, cy, thr);
}

#ifndef K1F_MIN_BLOCKS
#define K1F_MIN_BLOCKS 2
#endif
struct K1FSmem {
    CellRec cell[DSAC_N_CONST];
    unsigned short wlist[K1F_THREADS / 32][K1F_MAX_CHUNK / (K1F_THREADS / 32)];   // per warp: flagged candidates of its share of the item
};

// One work item = up to `chunk` consecutive candidates of one stream.  The CTA stages the frame's cell table (only when
// the frame changes); after that its warps run independently: a warp filters every 8th group of 32 candidates of the
// item (coalesced 8-byte loads of the cell indices, the next group's prefetched), keeps the flagged ones in its own
// list and appends them to the global queue with one atomic per item.  No block barrier except around the staging.
__global__ void __launch_bounds__(K1F_THREADS, K1F_MIN_BLOCKS) k1_filter(K1SplitParams q) {
    extern __shared__ __align__(16) unsigned char k1f_smem_raw[];
    K1FSmem& sm = *reinterpret_cast<K1FSmem*>(k1f_smem_raw);
    const SampleParams& p = q.sp;
    const int tid = threadIdx.x, lane = tid & 31, warp_id = tid >> 5;
    const int n_items = q.wq_n[q.round];
    const double inv_f = 1. / p.f, cx_f = p.cx * inv_f, cy_f = p.cy * inv_f;
    unsigned short* wl = sm.wlist[warp_id];
    int staged_frame = -1;
    unsigned long long n_flagged = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t it = q.wq[item];
        const int slot = (int)(it >> 6), chunk = (int)(it & 63u);
        const int frame = slot / p.T;
        if (frame != staged_frame) {   // the frame's cell table (38.4 KB) to shared memory
            __syncthreads();
            const uint4* src = reinterpret_cast<const uint4*>(q.celltab + (size_t)frame * DSAC_N_CONST);
            uint4* dst = reinterpret_cast<uint4*>(sm.cell);
            for (int k = tid; k < (int)(sizeof(CellRec) * DSAC_N_CONST / 16); k += K1F_THREADS) dst[k] = __ldg(src + k);
            staged_frame = frame;
            __syncthreads();
        }
        const int n_round = q.state[slot].n_round;
        const int base = chunk * q.chunk;
        const int n_here = min(q.chunk, n_round - base);
        const uint2* cc = q.cells + (size_t)slot * q.cap + base;
        int cnt = 0;
        int i = warp_id * 32 + lane;
        uint2 c = (i < n_here) ? __ldg(cc + i) : make_uint2(0u, 0u);
        for (int i0 = warp_id * 32; i0 < n_here; i0 += K1F_THREADS) {
            i = i0 + lane;
            const int in = i + K1F_THREADS;
            const uint2 cn = (in < n_here) ? __ldg(cc + in) : make_uint2(0u, 0u);   // next group's cell indices
            bool need = false;
            if (i < n_here) {
                const int cells[4] = {(int)(c.x & 0xffffu), (int)(c.x >> 16), (int)(c.y & 0xffffu), (int)(c.y >> 16)};
                need = k1_filter_candidate(sm.cell, cells, p.f, p.cx, p.cy, inv_f, cx_f, cy_f, (double)p.thr);
            }
            const uint32_t bits = __ballot_sync(0xffffffffu, need);
            if (need) wl[cnt + __popc(bits & ((1u << lane) - 1u))] = (unsigned short)(base + i);
            cnt += __popc(bits);
            c = cn;
        }
        __syncwarp();
        if (cnt > 0) {
            int at = 0;
            if (lane == 0) at = atomicAdd(q.fq_n + q.round, cnt);
            at = __shfl_sync(0xffffffffu, at, 0);
            for (int k = lane; k < cnt; k += 32) q.fq[at + k] = ((uint32_t)slot << 14) | (uint32_t)wl[k];
            n_flagged += (unsigned long long)cnt;
        }
        __syncwarp();
    }
    if (q.dbg && lane == 0 && n_flagged) atomicAdd(q.dbg + (size_t)q.round * 4 + 2, n_flagged);
}

// ------------------------------------------------------------------ k1_solve
// Full fp64 P3P + the reference's reprojection check (cnn_softam.h:1041-1059) on the flagged candidates; 4 lanes per
// candidate, lane `sub` handles quartic root `sub` (as phase D of k_sample).
#ifndef K1V_GROUP
#define K1V_GROUP 4   /* lanes per flagged candidate: 4 = one quartic root per lane, 2 = two roots per lane, 1 = one thread per candidate */
#endif
__global__ void __launch_bounds__(K1V_THREADS) k1_solve(K1SplitParams q) {
    const SampleParams& p = q.sp;
    const int tid = threadIdx.x;
    constexpr int GROUP = K1V_GROUP, GROUPS = K1V_THREADS / GROUP, ROOTS = 4 / GROUP;
    const int n_q = q.fq_n[q.round];
    const int sub = tid % GROUP;
    for (int base = blockIdx.x * GROUPS; base < n_q; base += gridDim.x * GROUPS) {
        const int qi = base + tid / GROUP;
        bool ok = false, fragile = false;
        double rvec[3], tvec[3];
        double e2 = 1.7976931348623157e308, R[9], t[3];
        float obj[12], img[8];
        int nsol = 0, slot = 0, ci = 0;
        if (qi < n_q) {
            const uint32_t ent = q.fq[qi];
            slot = (int)(ent >> 14);
            ci = (int)(ent & 16383u);
            const int frame = slot / p.T;
            const uint2 c = q.cells[(size_t)slot * q.cap + ci];
            const int cells[4] = {(int)(c.x & 0xffffu), (int)(c.x >> 16), (int)(c.y & 0xffffu), (int)(c.y >> 16)};
            const CellRec* cell = q.celltab + (size_t)frame * DSAC_N_CONST;
            const int32_t* pix = p.pix + (size_t)frame * p.pix_stride;
            P3PProblem pr;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const CellRec r = cell[cells[j]];
                pr.mu[j] = r.xn * p.f + p.cx;   // float * double
                pr.mv[j] = r.yn * p.f + p.cy;
                pr.X[j][0] = (double)r.X; pr.X[j][1] = (double)r.Y; pr.X[j][2] = (double)r.Z;
                obj[j * 3] = (float)pr.X[j][0]; obj[j * 3 + 1] = (float)pr.X[j][1]; obj[j * 3 + 2] = (float)pr.X[j][2];
                img[j * 2] = (float)__ldg(pix + cells[j] * 2);
                img[j * 2 + 1] = (float)__ldg(pix + cells[j] * 2 + 1);
            }
            P3PFront fr;
            p3p_front(pr, p.f, p.cx, p.cy, fr);
            nsol = p3p_full(pr, fr, p.f, p.cx, p.cy, R, t, &e2, sub * ROOTS, (sub + 1) * ROOTS);
            if (nsol == 0) e2 = 1.7976931348623157e308;
        }
        // minimum over the group; ties go to the lower root index, as the sequential loop would
        double be = e2;
        int bl = sub;
#pragma unroll
        for (int off = 1; off < GROUP; off <<= 1) {
            double oe = __shfl_xor_sync(0xffffffffu, be, off);
            int ol = __shfl_xor_sync(0xffffffffu, bl, off);
            if (oe < be || (oe == be && ol < bl)) { be = oe; bl = ol; }
        }
        if (nsol > 0 && bl == sub)
            ok = minimal_set_accept(obj, img, p.f, p.cx, p.cy, p.thr, R, t, e2, rvec, tvec, &fragile);
        if (fragile) atomicAdd(p.n_fragile, 1ull);
        if (ok) {
            double* po = q.pose_out + ((size_t)slot * q.cap + ci) * 6;
            po[0] = rvec[0]; po[1] = rvec[1]; po[2] = rvec[2];
            po[3] = tvec[0]; po[4] = tvec[1]; po[5] = tvec[2];
            atomicOr(q.accbits + (size_t)slot * (q.cap >> 5) + (ci >> 5), 1u << (ci & 31));
        }
    }
}

}  // namespace dsac
