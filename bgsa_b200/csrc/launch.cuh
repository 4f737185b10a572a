// launch.cuh -- host-side launch helper shared by the instance files.
#pragma once

#include <atomic>
#include <cstdlib>

#include "align_kernel.cuh"
#include "rows_kernel.cuh"

namespace bgsa {

struct LaunchArgs {
    PackedSubjects ps;
    const uint32_t *d_peq;           // [n_queries][5][row stride]
    int n_queries;
    int qlen;
    void *d_results;
    long long result_stride;
    unsigned long long *d_counters;  // [n_queries], zeroed by the launcher
    int sm_count;
    cudaStream_t stream;
    // dry_run: do not launch, only report in *resident_subjects how many subjects the persistent
    // grid works on at once (resident warps x subjects per warp) -- the host uses it to size chunks
    bool dry_run = false;
    long long *resident_subjects = nullptr;
    // ASCII rows (stride slen+1) instead of packed tiles: only for instances with a rows kernel (rows_kernel_fits);
    // ps then carries the geometry (slen, count) only
    const uint8_t *d_ascii = nullptr;
};

constexpr int kAlignThreads = 128;
constexpr int kAlignCH = 4;          // 128-bit units (256 bases) per lane and stage

template <class Algo, int L, int UNROLL>
cudaError_t launch_align(const LaunchArgs &a, typename Algo::Params prm) {
    auto kern = align_kernel<Algo, L, kAlignCH, kAlignThreads, UNROLL>;
    // resident CTAs per SM of this instance, cached PER DEVICE (contexts may differ in carve-out; first calls may race:
    // the atomics make that benign -- every thread computes the same value)
    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = occ_cache[dev & 63].load(std::memory_order_relaxed);
    if (occ == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kAlignThreads, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        occ_cache[dev & 63].store(occ, std::memory_order_relaxed);
    }
    if (a.dry_run) {
        if (a.resident_subjects) *a.resident_subjects = (long long)a.sm_count * occ * (kAlignThreads / 32) * (32 / L);
        return cudaSuccess;
    }
    cudaError_t e = cudaMemsetAsync(a.d_counters, 0, sizeof(unsigned long long) * a.n_queries, a.stream);
    if (e != cudaSuccess) return e;
    const long long warps_per_cta = kAlignThreads / 32;
    const int nq = a.n_queries > 0 ? a.n_queries : 1;
    long long want = (a.ps.ntiles * L * nq + warps_per_cta - 1) / warps_per_cta;   // one (query, tile, pass) work unit per warp at least
    const long long resident = (long long)a.sm_count * occ;
    if (want > resident) want = resident;
    if (want < 1) want = 1;
    kern<<<(unsigned)want, kAlignThreads, 0, a.stream>>>(a.ps, a.d_peq, nq, a.qlen, static_cast<int16_t *>(a.d_results),
                                                         a.result_stride, prm, a.d_counters);
    return cudaGetLastError();
}

// ---- rows kernel (rows_kernel.cuh): thread-per-subject straight from the ASCII rows ------------------------------------
constexpr int kRowsMaxK = 12;                    // mask table by byte value: 256 x 12 words = 12 KB up to K = 12 (queries up to 384 bases)
constexpr int kRowsMaxStageBytes = 13 * 1024;    // one tile of rows of up to ~400 bases
inline bool rows_kernel_fits(int K, int L, int slen) {
    const bool off = getenv("BGSA_NO_ROWS_KERNEL") != nullptr;               // A/B knob (read per call: the tests flip it): always pack + align
    return !off && L == 1 && K <= kRowsMaxK && rows_stage_bytes(slen + 1) <= kRowsMaxStageBytes && rows_bank_conflict_degree(slen + 1) <= 4;
}

template <class Algo, int UNROLL>
cudaError_t launch_align_rows(const LaunchArgs &a, typename Algo::Params prm) {
    auto kern = align_rows_kernel<Algo, kAlignThreads, UNROLL>;
    constexpr int WARPS = kAlignThreads / 32;
    const size_t table = sizeof(uint32_t) * 256 * peq_row_stride(Algo::K, 1);
    const size_t tb = (size_t)rows_stage_bytes(a.ps.slen + 1);
    static std::atomic<int> attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63].load(std::memory_order_relaxed)) {   // opt-in above 48 KB is a per-device function attribute
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63].store(1, std::memory_order_relaxed);
    }
    // two stages per warp (the next tile arrives while this one is aligned) unless that costs residency the ALU pipe needs:
    // with >= 3 CTAs (12 warps) per SM a warp waiting for its tile is covered by the others either way
    int occ1 = 0, occ2 = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, kAlignThreads, table + WARPS * 2 * tb);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, kern, kAlignThreads, table + WARPS * tb);
    if (e != cudaSuccess) return e;
    const char *force_ns = getenv("BGSA_ROWS_STAGES");          // A/B knob (read per launch)
    int nstage = (occ2 >= occ1 || occ2 >= 3) ? 2 : 1;
    if (force_ns && (atoi(force_ns) == 1 || atoi(force_ns) == 2)) nstage = atoi(force_ns);
    int occ = nstage == 2 ? occ2 : occ1;
    if (occ < 1) return cudaErrorInvalidConfiguration;
    if (a.dry_run) {
        if (a.resident_subjects) *a.resident_subjects = (long long)a.sm_count * occ * WARPS * 32;
        return cudaSuccess;
    }
    e = cudaMemsetAsync(a.d_counters, 0, sizeof(unsigned long long) * a.n_queries, a.stream);
    if (e != cudaSuccess) return e;
    const int nq = a.n_queries > 0 ? a.n_queries : 1;
    long long want = (a.ps.ntiles * nq + WARPS - 1) / WARPS;
    const long long resident = (long long)a.sm_count * occ;
    if (want > resident) want = resident;
    if (want < 1) want = 1;
    kern<<<(unsigned)want, kAlignThreads, table + WARPS * nstage * tb, a.stream>>>(a.d_ascii, a.ps.slen, a.ps.count, a.d_peq, nq, a.qlen,
                                                                                 static_cast<int16_t *>(a.d_results), a.result_stride, prm,
                                                                                 a.d_counters, nstage);
    return cudaGetLastError();
}

// what the instance tables call: the rows kernel when the caller handed ASCII rows to an instance that has one
template <class Algo, int L, int UNROLL>
cudaError_t launch_instance(const LaunchArgs &a, typename Algo::Params prm) {
    if constexpr (L == 1 && Algo::K <= kRowsMaxK) {
        if (a.d_ascii) return launch_align_rows<Algo, UNROLL>(a, prm);
    }
    if (a.d_ascii) return cudaErrorInvalidValue;
    return launch_align<Algo, L, UNROLL>(a, prm);
}

// instance-file entry points (return cudaErrorInvalidValue when (K, L) has no instance)
cudaError_t launch_myers(int mode, int K, int L, const LaunchArgs &a, int sign);
cudaError_t launch_bitpal_packed(int scheme, int K, int L, const LaunchArgs &a);
cudaError_t launch_bitpal_semiglobal(int scheme, int K, int L, const LaunchArgs &a);
cudaError_t launch_bitpal_nonpacked(int scheme, int K, int L, const LaunchArgs &a);
cudaError_t launch_banded(const LaunchArgs &a, const void *d_rows_table, int e);
bool banded_fused_fits(int slen);
cudaError_t launch_banded_fused(const LaunchArgs &a, const void *d_ascii_rows, const void *d_rows_table, int e);
cudaError_t launch_pack(int layout, const void *d_rows, int slen, long long count, void *d_packed, int sm_count,
                        cudaStream_t stream);
cudaError_t launch_unpeq(int layout, int wordbytes, const void *d_peq, int word_num, int usable, int head, int slen,
                         long long count, int vnum, void *d_packed, int sm_count, cudaStream_t stream);
cudaError_t launch_int_peak(int sm_count, int iters, unsigned int *d_sink, cudaStream_t stream);

}  // namespace bgsa
