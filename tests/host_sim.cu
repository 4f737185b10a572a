// tests/host_sim.cu -- TEST INFRASTRUCTURE ONLY.
//
// Runs the product's DP column functions (bgsa_b200/csrc/{myers,bitpal}.cuh -- the same source
// the CUDA kernels are instantiated from) on the HOST, where their PTX primitives are replaced
// by bit-exact C++ emulations (bgsa_common.cuh, #ifndef __CUDA_ARCH__).  It lets the CPU-only
// test tier check the recurrences, the carry words of the multi-lane wavefront and the score
// assembly against the oracle without a GPU.  The kernel skeleton, the packing and the banded
// kernel are only exercised on the GPU (tests marked gpu).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bitpal.cuh"
#include "instances.h"
#include "myers.cuh"
#include "query_peq.h"

using namespace bgsa;

static int map_char(unsigned char c) {
    switch (c) { case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4; default: return 0; }
}

// Emulates align_kernel for one (query, subject) pair with L lanes of K words: the wavefront is
// executed lane by lane, step by step, passing the packed carry word exactly like the
// __shfl_up_sync in the kernel does.
template <class Algo, int L>
static int run_pair(const char *qcodes, int qlen, const char *subject, int slen, typename Algo::Params prm) {
    constexpr int K = Algo::K;
    const int stride = h_peq_row_stride(K, L), kp = h_peq_kp(K);
    std::vector<uint32_t> peq_store(5 * stride + 8);
    uint32_t *peq = peq_store.data();
    while (reinterpret_cast<uintptr_t>(peq) & 15) peq++;      // LDS.128-style loads need 16-B alignment
    build_query_peq(qcodes, qlen, K, L, peq);
    std::vector<typename Algo::State> st(L);
    for (int r = 0; r < L; r++) Algo::init(st[r], r * K * 32, qlen);
    std::vector<uint32_t> packet(L, 0u), next(L, 0u);
    for (int t = 0; t < slen + L - 1; t++) {
        for (int r = 0; r < L; r++) {
            const int tcol = t - r;
            next[r] = packet[r];
            if (tcol < 0 || tcol >= slen) continue;
            uint32_t recv;
            if (r == 0) recv = Algo::kBoundary | (uint32_t)map_char((unsigned char)subject[tcol]);
            else recv = packet[r - 1];
            const uint32_t base = recv & 7u;
            if (L == 1) (void)Algo::template column<false>(st[r], peq + base * stride + r * kp, 0u);
            else next[r] = Algo::template column<true>(st[r], peq + base * stride + r * kp, recv) | base;
        }
        packet.swap(next);
    }
    int total = 0, best = 0, run = 0;
    for (int r = 0; r < L; r++) {
        Partial p = Algo::partial(st[r], r * K * 32, qlen);
        const int cand = run + p.minpre;
        if (r == 0 || cand < best) best = cand;
        run += p.sum;
    }
    total = run;
    return Algo::final_score(total, best, qlen, slen, prm);
}

template <class Algo, int L>
static void run_batch(const char *queries, int nq, int qlen, const char *subjects, long long ns, int slen,
                      typename Algo::Params prm, int16_t *out) {
    for (int q = 0; q < nq; q++)
        for (long long s = 0; s < ns; s++)
            out[q * ns + s] = narrow16(run_pair<Algo, L>(queries + (size_t)q * (qlen + 1), qlen,
                                                         subjects + (size_t)s * (slen + 1), slen, prm));
}

extern "C" {

// algo: 0 Myers global, 1 Myers semi-global, 3 BitPAl packed, 4 BitPAl non-packed, 5 BitPAl packed semi-global (bgsa_algo_t).
// (K, L) must be one of the instances of csrc/instances.h; scheme = index into BGSA_SCHEMES.
// Returns 0, or -1 when (algo, scheme, K, L) is not an instance.
int host_sim_align(int algo, int scheme, int K, int L, int sign, const char *queries, int nq, int qlen,
                   const char *subjects, long long ns, int slen, int16_t *out) {
    if (32 * K * L < qlen) return -1;
    if (algo == 0 || algo == 1) {
#define X(k, l)                                                                                              \
        if (K == k && L == l) {                                                                              \
            if (algo == 0) run_batch<MyersAlgo<k, MYERS_GLOBAL>, l>(queries, nq, qlen, subjects, ns, slen, MyersParams{sign}, out); \
            else run_batch<MyersAlgo<k, MYERS_SEMIGLOBAL>, l>(queries, nq, qlen, subjects, ns, slen, MyersParams{sign}, out);      \
            return 0;                                                                                        \
        }
        BGSA_MYERS_INSTANCES(X)
#undef X
        return -1;
    }
#define S(id, m, i, g)                                                                                       \
    if (scheme == id) {                                                                                      \
        using Sch = Scheme<m, i, g>;                                                                         \
        if (algo == 3) {                                                                                     \
            BGSA_BITPAL_PACKED_INSTANCES(XP)                                                                 \
        } else if (algo == 5) {                                                                              \
            BGSA_BITPAL_PACKED_INSTANCES(XS)                                                                 \
        } else if (algo == 4) {                                                                              \
            BGSA_BITPAL_NONPACKED_INSTANCES(XN)                                                              \
        }                                                                                                    \
        return -1;                                                                                           \
    }
#define XP(k, l) if (K == k && L == l) { run_batch<BitpalPacked<Sch, k>, l>(queries, nq, qlen, subjects, ns, slen, BitpalParams{0}, out); return 0; }
#define XS(k, l) if (K == k && L == l) { run_batch<BitpalPacked<Sch, k, BITPAL_SEMIGLOBAL>, l>(queries, nq, qlen, subjects, ns, slen, BitpalParams{0}, out); return 0; }
#define XN(k, l) if (K == k && L == l) { run_batch<BitpalNonPacked<Sch, k>, l>(queries, nq, qlen, subjects, ns, slen, BitpalParams{0}, out); return 0; }
    BGSA_SCHEMES(S)
#undef S
#undef XP
#undef XS
#undef XN
    return -1;
}

}  // extern "C"
