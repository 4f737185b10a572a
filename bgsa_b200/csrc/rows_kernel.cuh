// rows_kernel.cuh -- thread-per-subject alignment straight from the ASCII rows (short reads): no pack kernel, no
// packed tiles, no 2-bit decode in the column loop.
//
// The boundary hands the library one byte per base (seq_t.content, file.c:44-115).  align_kernel (align_kernel.cuh)
// wants 2-bit tiles, so short reads pay a separate pack launch (2 % of the step for BitPAl at 150 bp, 10 % for Myers)
// plus, per DP column, the extraction of the 2-bit code (LOP3 + SHF on the ALU pipe that bounds the kernel).  Here
//   * a tile = 32 consecutive rows = ONE contiguous run of 32 x (slen+1) bytes: a single 1-D bulk async copy (TMA
//     engine, SASS UBLKCP) brings it into the warp's shared-memory stage exactly as it lies in HBM;
//   * the query's match masks sit in shared memory as a table indexed by the BYTE VALUE (256 rows): rows 'A' 'C' 'G'
//     'T' 'N' hold the masks of codes 0..4, every other byte the masks of code 0 -- the alphabet rule of
//     original/BGSA_CPU/global.c:9-15 ("anything else is an A") becomes a table fill, and validating / encoding the
//     subject bytes costs no instruction at all;
//   * per DP column a lane reads its next byte (LDS.U8), turns it into a row address (IMAD: FMA pipe) and runs the
//     very same column function as align_kernel.
// ALU-pipe cost per column: that of the DP recurrence alone (Myers K=5: 49.3 instead of 51.3 + the pack kernel).
// HBM traffic per subject: slen+1 bytes in, 2 out (instead of slen/4 in after a pack pass that read slen+1 anyway).
//
// Restrictions (host side: rows_kernel_fits): L = 1 instances with a small mask table (K <= 12), a tile that fits the
// stage (rows up to ~400 bases), and a row pitch whose 32 lanes do not pile up on a few shared-memory banks (a pitch
// that is a multiple of 64 bytes would serialise every byte load; such sets take the pack + align path).
#pragma once

#include "align_kernel.cuh"

namespace bgsa {

// bytes of one stage: 32 rows + the misalignment of the run, in 16-byte granules
__host__ __device__ inline int rows_stage_bytes(int stride) { return (32 * stride + 15 + 15) / 16 * 16; }
// worst number of lanes whose byte loads fall into the same shared-memory bank (all lanes read the same column i of
// their own row: address = lane * stride + i)
#ifndef __CUDACC_RTC__
__host__ inline int rows_bank_conflict_degree(int stride) {
    int worst = 0;
    for (int sub = 0; sub < 4; sub++) {
        int hits[32] = {0};
        for (int lane = 0; lane < 32; lane++) hits[((lane * stride + sub) >> 2) & 31]++;
        for (int b = 0; b < 32; b++) worst = hits[b] > worst ? hits[b] : worst;
    }
    return worst;
}
#endif

__host__ __device__ constexpr int byte_code(int b) { return b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : b == 'N' ? 4 : 0; }

template <class Algo, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS, BGSA_MIN_BLOCKS)
align_rows_kernel(const uint8_t *__restrict__ rows, int slen, long long count, const uint32_t *__restrict__ g_peq, int n_queries,
                  int qlen, int16_t *__restrict__ results, long long result_stride, typename Algo::Params prm,
                  unsigned long long *__restrict__ counters, int nstage) {
    constexpr int K = Algo::K;
    constexpr int STRIDE = peq_row_stride(K, 1);
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) uint8_t s_rows_dyn[];           // [256][STRIDE] masks by byte value, then per warp nstage x TB bytes
    __shared__ __align__(8) uint64_t s_bar[WARPS * 2];
    __shared__ int s_skip;
    uint32_t *s_peq = reinterpret_cast<uint32_t *>(s_rows_dyn);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = slen + 1;
    const int TB = rows_stage_bytes(stride);
    uint8_t *stage = s_rows_dyn + 256 * STRIDE * sizeof(uint32_t) + (size_t)warp * nstage * TB;
    uint64_t *bar = s_bar + warp * 2;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncwarp();
    uint32_t phases = 0u;
    // tile starts are multiples of 32 bytes from `rows`: every run has the same misalignment
    const int off = (int)(reinterpret_cast<uintptr_t>(rows) & 15);
    const long long nunits = (count + kTileSubjects - 1) / kTileSubjects;
    // (all lanes call it) A run is fetched in 16-byte granules: rounding up reads into the next tile's rows, except at
    // the very end of the caller's buffer -- there the bulk copy stops at the last whole granule and the lanes bring the
    // remaining < 16 bytes themselves, so nothing past the last row is ever read.
    auto issue = [&](int s, long long tile) {
        const long long first = tile * kTileSubjects;
        const int live = (int)min((long long)kTileSubjects, count - first);
        const int run_end = off + live * stride;
        const bool last = first + live >= count;
        const uint32_t bytes = (uint32_t)(last ? run_end & ~15 : (run_end + 15) & ~15);
        uint8_t *dst = stage + s * TB;
        const uint8_t *src = rows + first * stride - off;
        if (lane == 0) {
            mbar_expect_tx(&bar[s], bytes);
            if (bytes) bulk_g2s(dst, src, bytes, &bar[s]);
        }
        if (last && (int)bytes + lane < run_end) dst[bytes + lane] = __ldg(src + bytes + lane);
    };
    int sb = 0;

    // (queries: the same round-robin walk of the CTAs as align_kernel)
    for (int visit = 0; visit < n_queries; visit++) {
        const int q = (int)((blockIdx.x + (unsigned)visit) % (unsigned)n_queries);
        unsigned long long *counter = counters + q;
        const long long static_units = (long long)(gridDim.x / n_queries + (q < (int)(gridDim.x % n_queries) ? 1 : 0)) * WARPS;
        __syncthreads();
        if (threadIdx.x == 0)
            s_skip = visit > 0 && static_units + (long long)*reinterpret_cast<volatile unsigned long long *>(counter) >= nunits;
        __syncthreads();
        if (s_skip) continue;
        const uint32_t *qp = g_peq + (size_t)q * kPeqRows * STRIDE;
        for (int i = threadIdx.x; i < 256 * STRIDE; i += THREADS) {
            const int b = i / STRIDE, j = i - b * STRIDE;
            s_peq[i] = qp[byte_code(b) * STRIDE + j];
        }
        __syncthreads();
        int16_t *out = results + (long long)q * result_stride;
        long long unit = visit == 0 ? (long long)(blockIdx.x / n_queries) * WARPS + warp : static_units + next_tile(counter, lane);
        if (unit < nunits) issue(sb, unit);

        while (unit < nunits) {
            const long long nxt = static_units + next_tile(counter, lane);
            if (nstage == 2 && nxt < nunits) issue(sb ^ 1, nxt);
            mbar_wait(&bar[sb], (phases >> sb) & 1u);
            phases ^= 1u << sb;
            __syncwarp();                                        // (the tail bytes of the last run were stored by other lanes)
            typename Algo::State state;
            Algo::init(state, 0, qlen);
            const uint8_t *r = stage + sb * TB + off + lane * stride;
#pragma unroll (UNROLL)
            for (int i = 0; i < slen; i++) {
                const uint32_t b = r[i];
                (void)Algo::template column<false>(state, s_peq + b * STRIDE, 0u);
            }
            __syncwarp();                                        // every lane has read its row: the stage may be refilled
            if (nstage == 2) sb ^= 1;
            else if (nxt < nunits) issue(0, nxt);
            const Partial p = Algo::partial(state, 0, qlen);
            const long long subject = unit * kTileSubjects + lane;
            if (subject < count) out[subject] = narrow16(Algo::final_score(p.sum, p.minpre, qlen, slen, prm));
            unit = nxt;
        }
    }
}

}  // namespace bgsa
