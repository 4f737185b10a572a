// myers.cuh -- unit-cost Myers/Hyyro bit-vector DP, global and semi-global, as an align_kernel policy.
//
// Replaces the generated align_cpu / align_sse (original/BGSA_CPU/align_core.c:19-148,
// original/BGSA_SSE/align_core.c:19-152) and the semi-global emission
// (generator/.../MyersGenerator.java:56-223).
//
// Formulation (ours, not the reference's): full 32-bit words, hardware carry (IADD3.X chain),
// funnel shifts for the one-position shift across words, no per-column score tracking:
//   global      : Pv = ~0, Mv = 0 (D[j][0] = j), top-row delta +1, score = n + popc(Pv) - popc(Mv)
//                 over the m valid query bits                       (== D[m][n], edit distance)
//   semi-global : reference semantics "whole subject inside the query" (SURVEY.md Appendix A9): in
//                 the transposed matrix the first column is all 0 (Pv = Mv = 0, free start in the
//                 query), the top row still grows by +1 per subject base, and the answer is the
//                 minimum over the last column, i.e. the minimum prefix sum of the final vertical
//                 deltas starting from n.
// 10 ALU-pipe instructions per 32-cell word-column: and, add-with-carry, or, lop3 (D0), lop3 (Ph),
// and (Mh), 2 funnel shifts, lop3 (Pv'), and (Mv').
#pragma once

#include "align_kernel.cuh"

namespace bgsa {

enum { MYERS_GLOBAL = 0, MYERS_SEMIGLOBAL = 1 };

struct MyersParams { int sign; };              // -1: score = -distance (generator -m 0), +1: -m 1

template <int K_, int MODE>
struct MyersAlgo {
    static constexpr int K = K_;
    static constexpr int kOpsPerWord = 10;
    // Wavefront instances: left alone ptxas takes 173-199 registers (2 CTAs = 2 warps per SM sub-partition, ALU pipe 94 %
    // busy, profiles/r02_ncu_metrics.csv myers5k); capped for 3 CTAs they need 141-152 without a single spill.
#ifndef BGSA_MYERS_WF_MINBLOCKS
#define BGSA_MYERS_WF_MINBLOCKS 3
#endif
    static constexpr int kMinBlocksWavefront = BGSA_MYERS_WF_MINBLOCKS;
    using Params = MyersParams;
    struct State { uint32_t pv[K], mv[K]; };
    // carry stream (CarryIn/CarryOut, consumption order): add carry, Ph shift-in, Mh shift-in
    static constexpr uint32_t kBoundary = 0x40000000u;   // top row: no carry, D[0][i] - D[0][i-1] = +1 (Ph), Mh 0

    static BGSA_HD void init(State &s, int, int) {
#pragma unroll
        for (int j = 0; j < K; j++) { s.pv[j] = (MODE == MYERS_GLOBAL) ? 0xffffffffu : 0u; s.mv[j] = 0u; }
    }

    // One DP column.  `row` = this lane's K words of the Peq row of the subject base (16-B aligned).
    template <bool CARRY>
    static BGSA_HD uint32_t column(State &s, const uint32_t *row, uint32_t cin) {
        uint32_t eq[(K + 3) / 4 * 4];
#pragma unroll
        for (int j = 0; j < (K + 3) / 4; j++) {
            const uint4 v = reinterpret_cast<const uint4 *>(row)[j];
            eq[4 * j] = v.x; eq[4 * j + 1] = v.y; eq[4 * j + 2] = v.z; eq[4 * j + 3] = v.w;
        }
        CarryIn in(cin);
        if (CARRY) in.to_cf();                                   // CF := add carry from the lane above
        uint32_t ph_prev = CARRY ? in.top() : 0x80000000u;       // only bit 31 is consumed
        uint32_t mh_prev = CARRY ? in.top() : 0u;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const uint32_t p = s.pv[j], m = s.mv[j], e = eq[j];
            const uint32_t t = e & p;
            const uint32_t sum = (j == 0 && !CARRY) ? add_cc(t, p) : addc_cc(t, p);
            const uint32_t x = e | m;
            const uint32_t d0 = lop3<(LA ^ LB) | LC>(sum, p, x);            // (sum ^ p) | x
            const uint32_t ph = lop3<LA | (0xFF ^ (LB | LC))>(m, d0, p);    // m | ~(d0 | p)
            const uint32_t mh = p & d0;
            // (word 0 without a lane above: the shift-in bits are the constants of the top row -- FMA pipe)
            const uint32_t phs = (j == 0 && !CARRY) ? shl1_const<1>(ph) : shl1_carry(ph_prev, ph);
            const uint32_t mhs = (j == 0 && !CARRY) ? shl1_const<0>(mh) : shl1_carry(mh_prev, mh);
            ph_prev = ph; mh_prev = mh;
            s.pv[j] = lop3<LA | (0xFF ^ (LB | LC))>(mhs, d0, phs);          // mhs | ~(d0 | phs)
            s.mv[j] = phs & d0;
        }
        if (!CARRY) return 0u;
        CarryOut out;
        out.push_cf();
        out.push_top(ph_prev);
        out.push_top(mh_prev);
        return out.finish<3>();
    }

    // pieces of the final vertical delta vector held by this lane (bits first_bit .. first_bit+32K)
    static BGSA_HD Partial partial(const State &s, int first_bit, int qlen) {
        Partial r; r.sum = 0; r.minpre = 0;
        if (MODE == MYERS_GLOBAL) {
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int rem = qlen - first_bit - 32 * j;
                const uint32_t mask = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                r.sum += popc32(s.pv[j] & mask) - popc32(s.mv[j] & mask);
            }
        } else {
            int cur = 0, best = 0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                int rem = qlen - first_bit - 32 * j;
                rem = rem < 32 ? rem : 32;
                uint32_t p = s.pv[j], m = s.mv[j];
                for (int b = 0; b < rem; b++) {
                    cur += (int)(p & 1u) - (int)(m & 1u);
                    p >>= 1; m >>= 1;
                    best = cur < best ? cur : best;
                }
            }
            r.sum = cur; r.minpre = best;
        }
        return r;
    }

    static BGSA_HD int final_score(int sum, int minpre, int qlen, int slen, Params prm) {
        (void)qlen;
        return prm.sign * (slen + (MODE == MYERS_GLOBAL ? sum : minpre));
    }
};

}  // namespace bgsa
