// launch.cuh -- host-side launch helper shared by the instance files.
#pragma once

#include <atomic>

#include "align_kernel.cuh"

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
