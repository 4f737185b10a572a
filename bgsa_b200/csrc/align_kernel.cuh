// align_kernel.cuh -- the persistent alignment kernel shared by the transposed-DP algorithms
// (Myers global / semi-global, BitPAl packed / non-packed).
//
// An algorithm is a policy class:
//     static constexpr int K;                         // 32-bit words of the bit-vector PER LANE
//     struct State;  struct Params;                   // per-lane DP state (registers), constants
//     static void init(State&, int first_bit, int qlen);   // first_bit: where this lane's segment starts
//     template <bool CARRY> static uint32_t column(State&, const uint32_t* peq_row, uint32_t cin);
//         one DP column = one subject base.  With CARRY the low word's carry-ins come from `cin`
//         (a CarryIn stream, top-aligned, bits 0..2 ignored) and the carry-outs are returned as
//         a CarryOut stream with bits 0..2 clear.
//     static constexpr uint32_t kBoundary;            // carry-in bits at the top row (lane 0)
//     static constexpr int kOpsPerWord;               // rough ALU instructions per word-column (unrolling policy)
//     static constexpr int kMinBlocksWavefront;       // resident CTAs per SM the L > 1 instances are compiled for
//     static Partial partial(const State&, int first_bit, int qlen);   // per-lane score pieces
//     static int final_score(sum, min_prefix, qlen, slen, Params);     // 32-bit score
//
// Thread mapping.  L = lanes per subject (1, 2, 4, 8, 16 or 32); the query bit-vector of L*K words
// is split into L consecutive segments of K words, one per lane of a group.
//   L == 1 : one subject per thread (short reads).  No cross-lane traffic at all.
//   L  > 1 : systolic wavefront inside the warp.  At step t lane r works on column t-r, so the
//            carries it needs (add carry, shift-in bits) were produced by lane r-1 one step
//            earlier: ONE __shfl_up_sync per step moves them, together with the subject base,
//            down the group.  No ballot/carry-lookahead is needed and every lane does full
//            K-word work; the price is L-1 idle steps per subject (< 1 % for the long sequences
//            this mode is for).
//
// Launch geometry: grid = persistent CTAs (SMs x occupancy), THREADS threads.  Shared memory holds the
// Peq (5 rows) of the query the CTA is working on and, per warp, a 2-stage buffer of subject tiles
// filled by 1-D bulk async copies (TMA engine) -- WarpStage in bgsa_common.cuh.  Warps take (tile,
// pass) work units from the query's counter (first unit static, then dynamic, claimed one ahead);
// with several queries the CTAs are dealt over them round-robin and move on when a query runs dry.
// HBM traffic per subject and query: slen/4 bytes in, 2 bytes out.
#pragma once

#include "bgsa_common.cuh"

namespace bgsa {

struct Partial { int sum; int minpre; };
template <bool B> struct BoolTag { static constexpr bool value = B; };     // (std::bool_constant, without the host header: NVRTC)   // segment sum of deltas, minimum prefix sum inside it

// Row layout of the query Peq in shared/global memory: lane r's K words start at r * KP(K),
// KP = K rounded up to 4 words so that every lane can use LDS.128.
__host__ __device__ constexpr int peq_kp(int k) { return (k + 3) / 4 * 4; }
__host__ __device__ constexpr int peq_row_stride(int k, int lanes) { return peq_stride(peq_kp(k) * lanes); }

// BGSA_MIN_BLOCKS: second __launch_bounds__ argument.  Without it ptxas caps the big instances at 128 registers
// and spills (K >= 8: -25 % on C5); with 1 it takes what the instance needs (measured: profiles/r01_minblocks_ab_mb{0,1}.log).
#ifndef BGSA_MIN_BLOCKS
#define BGSA_MIN_BLOCKS 1
#endif

template <class Algo, int L, int CH, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS, (L > 1 ? Algo::kMinBlocksWavefront : BGSA_MIN_BLOCKS))
align_kernel(PackedSubjects ps, const uint32_t *__restrict__ g_peq, int n_queries, int qlen, int16_t *__restrict__ results,
             long long result_stride, typename Algo::Params prm, unsigned long long *__restrict__ counters) {
    constexpr int K = Algo::K;
    constexpr int KP = peq_kp(K);
    constexpr int STRIDE = peq_row_stride(K, L);
    constexpr int WARPS = THREADS / 32;
    constexpr int GROUPS = 32 / L;                 // subjects in flight per warp
    constexpr int WUNROLL = (K * Algo::kOpsPerWord <= 100) ? 4 : 1;
    __shared__ __align__(16) uint32_t s_peq[kPeqRows * STRIDE];
    __shared__ __align__(128) uint4 s_stage[WARPS * 2 * CH * 32];
    __shared__ __align__(8) uint64_t s_bar[WARPS * 2];
    __shared__ int s_skip;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = lane % L, group = lane / L;
    WarpStage<CH> st;
    st.init(s_stage + warp * 2 * CH * 32, s_bar + warp * 2, lane);

    const uint32_t *my_peq = s_peq + rank * KP;
    const int slen = ps.slen, ku = ps.ku;
    const int nstages = (ku + CH - 1) / CH;
    // work unit = (tile, pass): the 32/L subjects of a tile that a warp has in flight at once
    const long long nunits = ps.ntiles * L;
    int sb = 0;

    // Many queries (the reference's buckets of up to 100, cal_cpu.c:210-216): the CTAs are dealt over the queries
    // round-robin (CTA c starts on query c % n_queries) and, when their query runs out of work, move on to the next
    // ones -- every resident CTA stays busy until the whole launch is done, whatever n_queries is.
    for (int visit = 0; visit < n_queries; visit++) {
    const int q = (int)((blockIdx.x + (unsigned)visit) % (unsigned)n_queries);
    unsigned long long *counter = counters + q;
    // Every warp of a CTA that STARTS on q has "its own" first unit (static), all further units come from the query's
    // counter: the next unit is claimed one ahead (its first stage is prefetched), and with a purely dynamic start
    // the first warps to arrive would claim two units each while others get none whenever a launch holds about
    // one unit per warp (chunked batches).
    const long long static_units = (long long)(gridDim.x / n_queries + (q < (int)(gridDim.x % n_queries) ? 1 : 0)) * WARPS;
    __syncthreads();                                 // everybody is done with the previous query's masks
    if (threadIdx.x == 0)
        s_skip = visit > 0 && static_units + (long long)*reinterpret_cast<volatile unsigned long long *>(counter) >= nunits;
    __syncthreads();
    if (s_skip) continue;                            // (a stale read only costs a useless visit)
    for (int i = threadIdx.x; i < kPeqRows * STRIDE; i += THREADS)
        s_peq[i] = g_peq[(size_t)q * kPeqRows * STRIDE + i];
    __syncthreads();
    int16_t *out = results + (long long)q * result_stride;
    long long unit = visit == 0 ? (long long)(blockIdx.x / n_queries) * WARPS + warp : static_units + next_tile(counter, lane);
    if (unit < nunits) st.issue(sb, ps.codes + (unit / L) * ku * 32, min(CH, ku), lane);

    while (unit < nunits) {
        const long long nxt = static_units + next_tile(counter, lane);
        const long long tile = unit / L;
        const int pass = (int)(unit % L);
        const bool with_n = ps.tile_has_n[tile] != 0;
        {
            const int sidx = pass * GROUPS + group;          // subject of this group inside the tile
            typename Algo::State state;
            Algo::init(state, rank * K * 32, qlen);
            uint32_t packet = 0u;                            // what this lane hands to lane rank+1
            int t = 0;                                       // wavefront step = column of lane 0 (checked steps only)

            // One wavefront step: receive (base, carries) from the lane above, do one column.  Lane r works
            // on column t - r, which exists for every lane once t >= L-1 and while t < slen: only the L-1
            // steps at either end of a subject need the range check.
            auto step = [&](uint32_t head_base, auto checked) {
                if (L == 1) {
                    (void)Algo::template column<false>(state, my_peq + head_base * STRIDE, 0u);
                } else {
                    uint32_t recv = __shfl_up_sync(0xffffffffu, packet, 1, L);
                    if (rank == 0) recv = Algo::kBoundary | head_base;
                    if (!decltype(checked)::value || (unsigned)(t - rank) < (unsigned)slen)
                        packet = Algo::template column<true>(state, my_peq + (recv & 7u) * STRIDE, recv) | (recv & 7u);
                    if (decltype(checked)::value) t++;
                }
            };
            constexpr BoolTag<true> kChecked{};
            constexpr BoolTag<false> kFree{};

            for (int sg = 0; sg < nstages; sg++) {
                // prefetch the following stage (same unit, or the first stage of the next unit)
                if (sg + 1 < nstages)
                    st.issue(sb ^ 1, ps.codes + (tile * ku + (long long)(sg + 1) * CH) * 32, min(CH, ku - (sg + 1) * CH), lane);
                else if (nxt < nunits)
                    st.issue(sb ^ 1, ps.codes + (nxt / L) * ku * 32, min(CH, ku), lane);
                st.wait(sb);
                const int units = min(CH, ku - sg * CH);
#pragma unroll 1
                for (int u = 0; u < units; u++) {
                    const uint4 v = st.load(sb, u, sidx);
                    const int base0 = (sg * CH + u) * kBasesPerUnit;
                    uint32_t n0 = 0u, n1 = 0u;
                    if (with_n) {
                        const int kk = 2 * (sg * CH + u);
                        n0 = ps.nmask[(tile * ps.kn + kk) * 32 + sidx];
                        if (kk + 1 < ps.kn) n1 = ps.nmask[(tile * ps.kn + kk + 1) * 32 + sidx];
                    }
                    // The body holds UNROLL + 2 copies of the column: unrolled 4x more only where a column is small
                    // (big instances would overflow the instruction cache -- measured -25 % on BitPAl K=10).
#pragma unroll (WUNROLL)
                    for (int w = 0; w < 4; w++) {
                        const int nb = min(16, slen - base0 - 16 * w);
                        uint32_t word = (w == 0) ? v.x : (w == 1) ? v.y : (w == 2) ? v.z : v.w;
                        if (!with_n && (L == 1 || base0 + 16 * w >= L - 1)) {
#pragma unroll (UNROLL)
                            for (int i = 0; i < nb; i++) {
                                const uint32_t c = word & 3u;
                                word >>= 2;
                                step(c, kFree);
                            }
                            t += nb > 0 ? nb : 0;
                        } else if (!with_n) {   // first L-1 columns of a subject: the wavefront is filling
#pragma unroll 1
                            for (int i = 0; i < nb; i++) {
                                const uint32_t c = word & 3u;
                                word >>= 2;
                                step(c, kChecked);
                            }
                        } else {   // rare path: the tile contains at least one 'N'
                            uint32_t nbits = ((w < 2) ? n0 : n1) >> (16 * (w & 1));
#pragma unroll 1
                            for (int i = 0; i < nb; i++) {
                                const uint32_t c = (word & 3u) | ((nbits & 1u) << 2);
                                word >>= 2;
                                nbits >>= 1;
                                step(c, kChecked);
                            }
                        }
                    }
                }
                __syncwarp();   // every lane has consumed stage sb before it is refilled
                sb ^= 1;
            }
            if (L > 1) {
#pragma unroll 1
                for (int i = 0; i < L - 1; i++) step(0u, kChecked);    // drain the wavefront
            }
            // ---- score: combine the per-lane pieces of the final delta vector
            Partial p = Algo::partial(state, rank * K * 32, qlen);
            int total = p.sum, best = p.minpre;
            if (L > 1) {
                // exclusive prefix of segment sums over the group, then min / total reductions
                int incl = p.sum;
#pragma unroll
                for (int o = 1; o < L; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o, L);
                    if (rank >= o) incl += up;
                }
                best = (incl - p.sum) + p.minpre;
#pragma unroll
                for (int o = L >> 1; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o, L));
                total = __shfl_sync(0xffffffffu, incl, L - 1, L);
            }
            const long long subject = tile * kTileSubjects + sidx;
            if (rank == 0 && subject < ps.count) out[subject] = narrow16(Algo::final_score(total, best, qlen, slen, prm));
        }
        unit = nxt;
    }
    }   // visit
}

}  // namespace bgsa
