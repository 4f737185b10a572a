// banded.cuh -- banded Myers verification under an error threshold e (Hyyro's diagonal band),
// one subject per thread.  Replaces banded/BGSA_CPU/align_core.c:69-252 (+ the Peq builder
// banded/BGSA_CPU/global.c:25-84).
//
// Unlike the other algorithms the banded result is NOT a symmetric DP value (early-exit value
// 127, minimum over e+1 end cells), so the reference orientation is
// kept: the band is a bit-vector of 2e+1 cells sliding along the SUBJECT, one row per QUERY base.
// What is redesigned:
//   * no per-subject Peq (the reference builds 5 x word_num x 8 B per subject and shifts all five
//     rows every query base, align_core.c:35-62): subjects are stored as two bit-planes (low /
//     high bit of the 2-bit base code, 32 bases per word); the band window of row r is ONE funnel
//     shift per plane and Eq is two LOP3 against per-row uniform masks (the query is the same
//     for every thread).  All edge masking (band width, cells left of the subject start, cells
//     right of its end, query 'N') is folded into the per-row mask table built on the host.
//   * 32-bit band words when 2e+2 <= 32 (the result is independent of the word width as long as
//     the band plus one bit fits -- tests/test_oracle_pins.py::test_banded_word_width), else 64.
//   * the error counter: bit 0 of D0 is shifted into an accumulator (one SHF per row) and
//     popcounted per 32 rows, instead of and/sub/add per row (align_core.c:64-67).
//   * early exit: err is non-decreasing, so "some checkpoint saw err > max_err" (align_core.c:
//     136-140,170-174,199-203,216-220) is equivalent to "the LAST checkpoint, after row
//     C = (q <= 64 ? q : max(64, q - e)), saw it".  Lanes whose err already exceeds max_err
//     before row C are certain to report 127; a warp stops as soon as all its lanes are.
// Restriction: query and subject must have the same length (the reference's own Peq layout is
// inconsistent otherwise, Appendix A7) and e <= 31.
#pragma once

#include "bgsa_common.cuh"
#include "pack.cuh"

namespace bgsa {

constexpr int kBandedMaxError = 127;   // MAX_ERROR, banded/BGSA_CPU/config.h:19

// Per-row uniform masks, built by build_banded_table() (banded_host.h).
struct __align__(16) BandedRow {
    uint32_t nclo, nchi;     // ~0 where the query base's code bit is 0 (so plane ^ mask = match)
    uint32_t bm_lo, bm_hi;   // valid band cells of this row (0 if the query base is N)
    uint32_t cn_lo, cn_hi;   // valid band cells if the query base is N, else 0
    uint32_t pad0, pad1;
};

template <bool WIDE> struct BandWord { using type = uint32_t; };
template <> struct BandWord<true> { using type = uint64_t; };

// planes layout of a 128-bit unit: x = low bits of bases 0..31, y = high bits of bases 0..31,
// z = low bits of bases 32..63, w = high bits of bases 32..63.
// (min-blocks: 10 CTAs of 128 threads = 48 registers for the narrow band -- what the kernel needs; left alone ptxas
//  takes 61 and the occupancy drops to 8)
// MULTI = false (one query): the row-mask table and the result row are kernel-wide constants (uniform-register
// addressing); MULTI = true: they change with the work unit's query.
// FUSED = true: the subjects come as ASCII rows (`ascii`, stride slen+1; ps only carries the geometry): the warp
// encodes its tile into a shared-memory strip first (step 1 of the streaming pack, pack.cuh) and the lanes cut their
// bit-plane words out of the strip instead of loading packed tiles -- no packed round trip through HBM, and the
// memory-bound encode of one warp overlaps the ALU-bound band rows of the others.
template <bool WIDE, bool MULTI, bool FUSED, int THREADS>
__global__ void __launch_bounds__(THREADS, FUSED ? (WIDE ? 7 : 9) : (WIDE ? 8 : 10))
banded_kernel(PackedSubjects ps, const uint8_t *__restrict__ ascii, const BandedRow *__restrict__ g_rows, int n_queries, int qlen,
              int e, int8_t *__restrict__ results, long long result_stride, unsigned long long *__restrict__ counters) {
    using T = typename BandWord<WIDE>::type;
    extern __shared__ __align__(16) uint32_t s_strips[];       // FUSED: one strip per warp
    const int lane = threadIdx.x & 31;
    const int ku = ps.ku;
    const int stride = ps.slen + 1;
    uint32_t *s_c = s_strips + (threadIdx.x >> 5) * (FUSED ? pack_warp_words(stride, 1) : 0);
    uint16_t *s_n = reinterpret_cast<uint16_t *>(s_c + (FUSED ? pack_code_words(stride, 1) : 0));
    const int off = FUSED ? (int)(reinterpret_cast<uintptr_t>(ascii) & 15) : 0;
    const int inc = 512 % stride;
    int m0 = (16 * lane - off) % stride;
    if (m0 < 0) m0 += stride;
    const int base_pos = off + lane * stride;                 // strip position of this lane's own row
    const int wi0 = base_pos >> 4, sub = base_pos & 15;
    const int sh = e + 1;                                     // plane index u = i + e + 1
    const int C = qlen <= 64 ? qlen : max(64, qlen - e);      // rows done at the last checkpoint
    const int max_err = 2 * e + 1;                            // threshold + h_threshold + 1 (:114)

    // work unit = (query, tile), query-major, from ONE counter: the per-row masks of a query are read straight from
    // global memory (uniform loads), so a warp changes query for free and every warp stays busy whatever n_queries is
    const long long nwork = ps.ntiles * n_queries;
    for (long long work = next_tile(counters, lane); work < nwork; work = next_tile(counters, lane)) {
        const int q = MULTI ? (int)(work / ps.ntiles) : 0;
        const long long tile = work - (long long)q * ps.ntiles;
        const BandedRow *rows = MULTI ? g_rows + (size_t)q * qlen : g_rows;
        int8_t *out = MULTI ? results + (long long)q * result_stride : results;
        bool with_n;
        if (FUSED) {
            const long long first = tile * kTileSubjects;
            const int live_rows = (int)min((long long)kTileSubjects, ps.count - first);
            const int npieces = (off + live_rows * stride + 15) >> 4;
            __syncwarp();                                      // the previous tile's strip is no longer read
            const uint32_t any_n = pack_run_to_strip<LAYOUT_PLANES, false>(reinterpret_cast<const uint4 *>(ascii + first * stride - off), npieces,
                                                                    ps.slen, stride, inc, m0, lane, s_c, s_n);
            with_n = __ballot_sync(0xffffffffu, any_n != 0u) != 0u;
            __syncwarp();
        } else {
            with_n = ps.tile_has_n[tile] != 0;
        }
        const uint4 *src = ps.codes + tile * ku * 32 + lane;
        const uint32_t *nsrc = ps.nmask + tile * ps.kn * 32 + lane;
        T VP = 0, VN = 0;
        int ones = 0;            // rows >= e (and < C) whose D0 bit 0 was 1
        int ones_tail = 0;       // same for rows >= C
        bool dead = false;
        // plane words around the current 32-row block: p*[0] = word kb-1, [1] = kb, [2] = kb+1, [3] = kb+2
        uint32_t plo[4] = {0u, 0u, 0u, 0u}, phi[4] = {0u, 0u, 0u, 0u}, pn[4] = {0u, 0u, 0u, 0u};
        uint4 unit = make_uint4(0u, 0u, 0u, 0u);
        auto plane_word = [&](int k, uint32_t &lo, uint32_t &hi, uint32_t &nn) {
            lo = hi = nn = 0u;
            if (FUSED) {
                // (dead lanes of the last tile read stale strip words: harmless, their result is not stored)
                if (k >= 0 && k < 2 * ku) {
                    strip_plane_word(s_c, wi0, sub, k, ps.slen, lo, hi);
                    if (with_n && k < ps.kn) nn = strip_n_word(s_n, wi0, sub, k, ps.slen);
                }
            } else if (k >= 0 && k < 2 * ku) {
                if ((k & 1) == 0) unit = __ldg(src + (long long)(k >> 1) * 32);
                lo = (k & 1) ? unit.z : unit.x;
                hi = (k & 1) ? unit.w : unit.y;
                if (with_n && k < ps.kn) nn = __ldg(nsrc + (long long)k * 32);
            }
        };
        plane_word(0, plo[1], phi[1], pn[1]);
        plane_word(1, plo[2], phi[2], pn[2]);
        plane_word(2, plo[3], phi[3], pn[3]);

        const int nblocks = (qlen + 31) / 32;
        for (int kb = 0; kb < nblocks; kb++) {
            // u-indexed plane words U[kb], U[kb+1] (, U[kb+2]): plane shifted left by e+1 bits
            const uint32_t ulo0 = __funnelshift_lc(plo[0], plo[1], sh), ulo1 = __funnelshift_lc(plo[1], plo[2], sh);
            const uint32_t uhi0 = __funnelshift_lc(phi[0], phi[1], sh), uhi1 = __funnelshift_lc(phi[1], phi[2], sh);
            const uint32_t un0 = __funnelshift_lc(pn[0], pn[1], sh), un1 = __funnelshift_lc(pn[1], pn[2], sh);
            uint32_t ulo2 = 0u, uhi2 = 0u, un2 = 0u;
            if (WIDE) {
                ulo2 = __funnelshift_lc(plo[2], plo[3], sh);
                uhi2 = __funnelshift_lc(phi[2], phi[3], sh);
                un2 = __funnelshift_lc(pn[2], pn[3], sh);
            }
            const int nrows = min(32, qlen - 32 * kb);
            uint32_t acc = 0u;
            auto row = [&](int t) {
                const BandedRow *rw = rows + 32 * kb + t;
                const uint4 m = __ldg(reinterpret_cast<const uint4 *>(rw));
                T wlo, whi, eq;
                if (!WIDE) {
                    wlo = __funnelshift_r(ulo0, ulo1, t);
                    whi = __funnelshift_r(uhi0, uhi1, t);
                    const uint32_t x = lop3<(LA ^ LB) & LC>(wlo, m.x, m.z);       // (lo ^ nclo) & bm
                    eq = lop3<(LA ^ LB) & LC>(whi, m.y, x);                       // (hi ^ nchi) & x
                    if (with_n) {
                        const uint4 m2 = __ldg(reinterpret_cast<const uint4 *>(rw) + 1);
                        const uint32_t wn = __funnelshift_r(un0, un1, t);
                        eq = (eq & ~wn) | (wn & m2.x);
                    }
                } else {
                    const uint32_t l0 = __funnelshift_r(ulo0, ulo1, t), l1 = __funnelshift_r(ulo1, ulo2, t);
                    const uint32_t h0 = __funnelshift_r(uhi0, uhi1, t), h1 = __funnelshift_r(uhi1, uhi2, t);
                    const uint32_t x0 = lop3<(LA ^ LB) & LC>(l0, m.x, m.z), x1 = lop3<(LA ^ LB) & LC>(l1, m.x, m.w);
                    uint32_t e0 = lop3<(LA ^ LB) & LC>(h0, m.y, x0), e1 = lop3<(LA ^ LB) & LC>(h1, m.y, x1);
                    if (with_n) {
                        const uint4 m2 = __ldg(reinterpret_cast<const uint4 *>(rw) + 1);
                        const uint32_t n0 = __funnelshift_r(un0, un1, t), n1 = __funnelshift_r(un1, un2, t);
                        e0 = (e0 & ~n0) | (n0 & m2.x);
                        e1 = (e1 & ~n1) | (n1 & m2.y);
                    }
                    eq = ((uint64_t)e1 << 32) | e0;
                }
                // cal_D0 (banded/BGSA_CPU/align_core.c:19-33)
                const T X = eq | VN;
                const T D0 = (((X & VP) + VP) ^ VP) | X;
                const T HN = D0 & VP;
                const T HP = VN | ~(D0 | VP);
                const T X2 = D0 >> 1;
                VN = X2 & HP;
                VP = HN | ~(HP | X2);
                acc = __funnelshift_r(acc, (uint32_t)D0, 1);   // collect D0 bit 0, row t ends at bit t
            };
            if (nrows == 32) {
#pragma unroll
                for (int t = 0; t < 32; t++) row(t);
            } else {
#pragma unroll 1
                for (int t = 0; t < nrows; t++) row(t);
                acc >>= (32 - nrows);
            }
            // error bookkeeping (align_core.c:113-134): rows < e are not counted
            uint32_t counted = acc;
            if (32 * kb < e) counted &= ~((e - 32 * kb >= 32) ? 0xffffffffu : ((1u << (e - 32 * kb)) - 1u));
            const int cbits = C - 32 * kb;                     // rows of this block before the last checkpoint
            const uint32_t before = cbits >= 32 ? 0xffffffffu : (cbits <= 0 ? 0u : ((1u << cbits) - 1u));
            ones += __popc(counted & before);
            ones_tail += __popc(counted & ~before);
            // early exit: err so far (over rows < min(done, C)) already above max_err => certain 127
            const int done = min(32 * kb + nrows, C);
            const int err_now = e + max(done - e, 0) - ones;
            dead = err_now > max_err;
            if (__all_sync(0xffffffffu, dead)) break;
            // slide the plane words
            plo[0] = plo[1]; plo[1] = plo[2]; plo[2] = plo[3];
            phi[0] = phi[1]; phi[1] = phi[2]; phi[2] = phi[3];
            pn[0] = pn[1]; pn[1] = pn[2]; pn[2] = pn[3];
            plane_word(kb + 3, plo[3], phi[3], pn[3]);
        }
        const long long subject = tile * kTileSubjects + lane;
        const int err_c = e + max(C - e, 0) - ones;            // err at the last checkpoint
        int result;
        if (dead || err_c > max_err) {
            result = kBandedMaxError;
        } else {
            int err = err_c + (qlen - C) - ones_tail;          // all rows (:221-226)
            int best = err;                                    // :230-240, last_bits = h_threshold = e
            for (int i = 0; i <= e; i++) {
                err += (int)((VP >> i) & 1) - (int)((VN >> i) & 1);
                best = min(best, err);
            }
            result = best;
        }
        if (subject < ps.count) out[subject] = (int8_t)result;   // int64 -> int8 (:242-245)
    }
}

}  // namespace bgsa
