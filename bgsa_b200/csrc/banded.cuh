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

#include <type_traits>

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
//
// Survivor compaction (refill_block = P >= 0).  The early exit above is per warp: a tile whose 32 subjects are a mix
// of hopeless ones (dead after the first rows) and near-matches (all rows needed) would run all its rows with most
// lanes idle, so the throughput would depend on the ORDER of the subjects.  Instead, after row block P (the host
// picks the block by which a random subject has certainly exceeded max_err) a tile that has few lanes left alive
// stores 127 for its dead lanes and PARKS the survivors -- VP, VN, error counters, subject id and (FUSED) their
// remaining bit-plane words -- in a per-warp ring in shared memory; whenever 32 survivors have gathered the warp runs
// their remaining blocks with every lane busy.  Results do not depend on the order of evaluation, so parity is
// untouched (tests/test_gpu_parity.py::test_banded_shuffled_*).

// shared-memory words of one parked survivor: VP, VN (2 or 4), counters, id (2), then FUSED plane words P .. 2ku-1
__host__ __device__ inline int banded_entry_words(bool wide, bool fused, int ku, int refill_block) {
    return (wide ? 4 : 2) + 3 + (fused ? 2 * (2 * ku - refill_block) : 0);
}
constexpr int kBandedRing = 64;        // ring slots per warp (at most 31 waiting + 24 new)
__host__ __device__ inline int banded_warp_words(bool wide, bool fused, int slen, int refill_block) {
    const int ku = (slen + kBasesPerUnit - 1) / kBasesPerUnit;
    return (fused ? pack_warp_words(slen + 1, 1) : 0) + (refill_block >= 0 ? banded_entry_words(wide, fused, ku, refill_block) * kBandedRing : 0);
}

template <bool WIDE, bool MULTI, bool FUSED, int THREADS>
__global__ void __launch_bounds__(THREADS, FUSED ? (WIDE ? 7 : 9) : (WIDE ? 8 : 10))
banded_kernel(PackedSubjects ps, const uint8_t *__restrict__ ascii, const BandedRow *__restrict__ g_rows, int n_queries, int qlen,
              int e, int8_t *__restrict__ results, long long result_stride, unsigned long long *__restrict__ counters,
              int refill_block, int refill_max_alive) {
    using T = typename BandWord<WIDE>::type;
    extern __shared__ __align__(16) uint32_t s_dyn[];          // per warp: [FUSED strip][survivor ring]
    constexpr uint32_t kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int ku = ps.ku;
    const int nwords = 2 * ku;                                 // 32-base plane words per subject
    const int stride = ps.slen + 1;
    const int P = refill_block;
    const int strip_words = FUSED ? pack_warp_words(stride, 1) : 0;
    uint32_t *s_c = s_dyn + (threadIdx.x >> 5) * banded_warp_words(WIDE, FUSED, ps.slen, P);
    uint16_t *s_n = reinterpret_cast<uint16_t *>(s_c + (FUSED ? pack_code_words(stride, 1) : 0));
    uint32_t *s_q = s_c + strip_words;                         // ring, structure of arrays: s_q[word * kBandedRing + slot]
    constexpr int QW_STATE = WIDE ? 4 : 2, QW_CNT = QW_STATE, QW_ID = QW_STATE + 1, QW_PLANES = QW_STATE + 3;
    const int off = FUSED ? (int)(reinterpret_cast<uintptr_t>(ascii) & 15) : 0;
    const int inc = 512 % stride;
    int m0 = (16 * lane - off) % stride;
    if (m0 < 0) m0 += stride;
    const int base_pos = off + lane * stride;                 // strip position of this lane's own row
    const int wi0 = base_pos >> 4, sub = base_pos & 15;
    const int sh = e + 1;                                     // plane index u = i + e + 1
    const int C = qlen <= 64 ? qlen : max(64, qlen - e);      // rows done at the last checkpoint
    const int max_err = 2 * e + 1;                            // threshold + h_threshold + 1 (:114)
    const int nblocks = (qlen + 31) / 32;

    // ---- the subject this lane is working on
    T VP = 0, VN = 0;
    int ones = 0;            // rows >= e (and < C) whose D0 bit 0 was 1
    int ones_tail = 0;       // same for rows >= C
    bool dead = false;
    // plane words around the current 32-row block: p*[0] = word kb-1, [1] = kb, [2] = kb+1, [3] = kb+2
    uint32_t plo[4] = {0u, 0u, 0u, 0u}, phi[4] = {0u, 0u, 0u, 0u}, pn[4] = {0u, 0u, 0u, 0u};
    const BandedRow *rows = g_rows;
    int qslot = 0;                                             // ring slot of the survivor this lane took

    // one block of up to 32 band rows; WN: the tile holds an 'N' somewhere (rare -- its own copy of the code, so that
    // the common path carries no predicated-off N instructions: 13 instead of 15 ALU instructions per row)
    auto block = [&](int kb, auto wn_tag) {
        constexpr bool WN = decltype(wn_tag)::value;
        // u-indexed plane words U[kb], U[kb+1] (, U[kb+2]): plane shifted left by e+1 bits
        const uint32_t ulo0 = __funnelshift_lc(plo[0], plo[1], sh), ulo1 = __funnelshift_lc(plo[1], plo[2], sh);
        const uint32_t uhi0 = __funnelshift_lc(phi[0], phi[1], sh), uhi1 = __funnelshift_lc(phi[1], phi[2], sh);
        uint32_t un0 = 0u, un1 = 0u, un2 = 0u;
        if (WN) { un0 = __funnelshift_lc(pn[0], pn[1], sh); un1 = __funnelshift_lc(pn[1], pn[2], sh); }
        uint32_t ulo2 = 0u, uhi2 = 0u;
        if (WIDE) {
            ulo2 = __funnelshift_lc(plo[2], plo[3], sh);
            uhi2 = __funnelshift_lc(phi[2], phi[3], sh);
            if (WN) un2 = __funnelshift_lc(pn[2], pn[3], sh);
        }
        const int nrows = min(32, qlen - 32 * kb);
        uint32_t acc = 0u;
        auto row = [&](int t) {
            const BandedRow *rw = rows + 32 * kb + t;
            const uint4 m = __ldg(reinterpret_cast<const uint4 *>(rw));
            T eq;
            if (!WIDE) {
                const uint32_t wlo = __funnelshift_r(ulo0, ulo1, t);
                const uint32_t whi = __funnelshift_r(uhi0, uhi1, t);
                const uint32_t x = lop3<(LA ^ LB) & LC>(wlo, m.x, m.z);       // (lo ^ nclo) & bm
                eq = lop3<(LA ^ LB) & LC>(whi, m.y, x);                       // (hi ^ nchi) & x
                if (WN) {
                    const uint4 m2 = __ldg(reinterpret_cast<const uint4 *>(rw) + 1);
                    const uint32_t wn = __funnelshift_r(un0, un1, t);
                    eq = (eq & ~wn) | (wn & m2.x);
                }
            } else {
                const uint32_t l0 = __funnelshift_r(ulo0, ulo1, t), l1 = __funnelshift_r(ulo1, ulo2, t);
                const uint32_t h0 = __funnelshift_r(uhi0, uhi1, t), h1 = __funnelshift_r(uhi1, uhi2, t);
                const uint32_t x0 = lop3<(LA ^ LB) & LC>(l0, m.x, m.z), x1 = lop3<(LA ^ LB) & LC>(l1, m.x, m.w);
                uint32_t e0 = lop3<(LA ^ LB) & LC>(h0, m.y, x0), e1 = lop3<(LA ^ LB) & LC>(h1, m.y, x1);
                if (WN) {
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
        if (nrows == 32 && !WN) {
#pragma unroll
            for (int t = 0; t < 32; t++) row(t);
        } else if (nrows == 32) {
#pragma unroll 4
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
    };

    // work unit = (query, tile), query-major, from ONE counter: the per-row masks of a query are read straight from
    // global memory (uniform loads), so a warp changes query for free and every warp stays busy whatever n_queries is
    const long long nwork = ps.ntiles * n_queries;
    long long work = next_tile(counters, lane);
    int qcount = 0, qhead = 0, ring_query = 0;                 // parked survivors (all of one query)

    while (true) {
        const bool tiles_left = work < nwork;
        const int wq = (MULTI && tiles_left) ? (int)(work / ps.ntiles) : 0;
        // a full warp of survivors first; the rest of the ring when the tiles (or the ring's query) have run out
        const bool from_ring = qcount >= 32 || (qcount > 0 && (!tiles_left || (MULTI && wq != ring_query)));
        if (!from_ring && !tiles_left) break;

        int q, kb0;
        long long tile, subject;
        bool store, with_n = false;
        const uint4 *src;
        if (from_ring) {
            const int n = min(32, qcount);
            store = lane < n;
            qslot = (qhead + lane) & (kBandedRing - 1);
            q = ring_query;
            const uint32_t *ent = s_q + qslot;
            VP = ent[0]; VN = ent[kBandedRing];
            if (WIDE) { VP |= (T)ent[2 * kBandedRing] << (WIDE ? 32 : 0); VN |= (T)ent[3 * kBandedRing] << (WIDE ? 32 : 0); }
            const uint32_t cnt = ent[QW_CNT * kBandedRing];
            ones = (int)(cnt & 0xffffu); ones_tail = (int)(cnt >> 16);
            const long long id = (long long)ent[QW_ID * kBandedRing] | ((long long)ent[(QW_ID + 1) * kBandedRing] << 32);
            const long long w = id >> 5;
            tile = MULTI ? w - (long long)q * ps.ntiles : w;
            subject = tile * kTileSubjects + (id & 31);
            if (!store) { tile = 0; subject = 0; }
            src = ps.codes + tile * ku * 32 + (store ? (int)(id & 31) : 0);
            dead = !store;
            kb0 = P + 1;
            qhead = (qhead + n) & (kBandedRing - 1);
            qcount -= n;
        } else {
            q = wq;
            tile = work - (long long)q * ps.ntiles;
            if (FUSED) {
                const long long first = tile * kTileSubjects;
                const int live_rows = (int)min((long long)kTileSubjects, ps.count - first);
                const int npieces = (off + live_rows * stride + 15) >> 4;
                __syncwarp();                                      // the previous tile's strip is no longer read
                const uint32_t any_n = pack_run_to_strip<LAYOUT_PLANES, false>(reinterpret_cast<const uint4 *>(ascii + first * stride - off), npieces,
                                                                        ps.slen, stride, inc, m0, lane, s_c, s_n);
                with_n = __ballot_sync(kFull, any_n != 0u) != 0u;
                __syncwarp();
            } else {
                with_n = ps.tile_has_n[tile] != 0;
            }
            subject = tile * kTileSubjects + lane;
            store = subject < ps.count;
            src = ps.codes + tile * ku * 32 + lane;
            VP = 0; VN = 0; ones = 0; ones_tail = 0;
            dead = false;
            kb0 = 0;
        }
        rows = MULTI ? g_rows + (size_t)q * qlen : g_rows;
        int8_t *out = MULTI ? results + (long long)q * result_stride : results;
        const uint32_t *nsrc = ps.nmask + tile * ps.kn * 32 + lane;
        uint4 unit = make_uint4(0u, 0u, 0u, 0u);
        int unit_k = -1;

        // plane word k (32 bases) of this lane's subject: from the strip / the packed tile, or -- for a parked
        // survivor of the FUSED kernel -- from the words saved in the ring
        auto plane_word = [&](int k, uint32_t &lo, uint32_t &hi, uint32_t &nn) {
            lo = hi = nn = 0u;
            if (k < 0 || k >= nwords) return;
            if (FUSED) {
                // (dead lanes of the last tile read stale strip words: harmless, their result is not stored)
                if (from_ring) {
                    lo = s_q[(QW_PLANES + 2 * (k - P)) * kBandedRing + qslot];
                    hi = s_q[(QW_PLANES + 2 * (k - P) + 1) * kBandedRing + qslot];
                } else {
                    strip_plane_word(s_c, wi0, sub, k, ps.slen, lo, hi);
                    if (with_n && k < ps.kn) nn = strip_n_word(s_n, wi0, sub, k, ps.slen);
                }
            } else {
                if ((k >> 1) != unit_k) { unit = __ldg(src + (long long)(k >> 1) * 32); unit_k = k >> 1; }
                lo = (k & 1) ? unit.z : unit.x;
                hi = (k & 1) ? unit.w : unit.y;
                if (with_n && k < ps.kn) nn = __ldg(nsrc + (long long)k * 32);
            }
        };
        plane_word(kb0 - 1, plo[0], phi[0], pn[0]);
        plane_word(kb0, plo[1], phi[1], pn[1]);
        plane_word(kb0 + 1, plo[2], phi[2], pn[2]);
        plane_word(kb0 + 2, plo[3], phi[3], pn[3]);

        bool parked = false;
        for (int kb = kb0; kb < nblocks; kb++) {
            if (with_n) block(kb, std::true_type{}); else block(kb, std::false_type{});
            if (__all_sync(kFull, dead || !store)) break;
            if (!from_ring && kb == P && !with_n && kb + 1 < nblocks) {
                const bool survivor = store && !dead;
                const uint32_t alive = __ballot_sync(kFull, survivor);
                if (__popc(alive) <= refill_max_alive) {
                    // park the survivors, settle the dead lanes now
                    if (survivor) {
                        const int slot = (qhead + qcount + __popc(alive & ((1u << lane) - 1u))) & (kBandedRing - 1);
                        uint32_t *ent = s_q + slot;
                        ent[0] = (uint32_t)VP; ent[kBandedRing] = (uint32_t)VN;
                        if (WIDE) { ent[2 * kBandedRing] = (uint32_t)((uint64_t)VP >> (WIDE ? 32 : 0)); ent[3 * kBandedRing] = (uint32_t)((uint64_t)VN >> (WIDE ? 32 : 0)); }
                        ent[QW_CNT * kBandedRing] = (uint32_t)ones | ((uint32_t)ones_tail << 16);
                        const long long id = work * 32 + lane;
                        ent[QW_ID * kBandedRing] = (uint32_t)id; ent[(QW_ID + 1) * kBandedRing] = (uint32_t)(id >> 32);
                        if (FUSED) {
                            for (int k = P; k < nwords; k++) {
                                uint32_t lo, hi;
                                strip_plane_word(s_c, wi0, sub, k, ps.slen, lo, hi);
                                ent[(QW_PLANES + 2 * (k - P)) * kBandedRing] = lo;
                                ent[(QW_PLANES + 2 * (k - P) + 1) * kBandedRing] = hi;
                            }
                        }
                    } else if (store) {
                        out[subject] = (int8_t)kBandedMaxError;
                    }
                    qcount += __popc(alive);
                    ring_query = q;
                    __syncwarp();
                    parked = true;
                    break;
                }
            }
            // slide the plane words
            plo[0] = plo[1]; plo[1] = plo[2]; plo[2] = plo[3];
            phi[0] = phi[1]; phi[1] = phi[2]; phi[2] = phi[3];
            pn[0] = pn[1]; pn[1] = pn[2]; pn[2] = pn[3];
            plane_word(kb + 3, plo[3], phi[3], pn[3]);
        }
        if (!parked) {
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
            if (store) out[subject] = (int8_t)result;   // int64 -> int8 (:242-245)
        }
        if (from_ring) __syncwarp();                               // the slots just read may be refilled
        else work = next_tile(counters, lane);
    }
}

}  // namespace bgsa
