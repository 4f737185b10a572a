// inst_bitpal.cu -- BitPAl kernel instances for ONE scoring scheme and ONE variant (BGSA_PACKED: 0 non-packed,
// 1 packed global, 2 packed semi-global), compiled once per (scheme, variant) so that the heavy instances build
// in parallel.
#include "instances.h"
#include "launch.cuh"
#include "bitpal.cuh"

namespace bgsa {

#if !defined(BGSA_SCHEME_ID) || !defined(BGSA_M) || !defined(BGSA_I) || !defined(BGSA_G) || !defined(BGSA_PACKED)
#error "compile with -DBGSA_SCHEME_ID= -DBGSA_M= -DBGSA_I= -DBGSA_G= -DBGSA_PACKED="
#endif

#define BGSA_CAT_(a, b) a##b
#define BGSA_CAT(a, b) BGSA_CAT_(a, b)

using TheScheme = Scheme<BGSA_M, BGSA_I, BGSA_G>;

#if BGSA_PACKED == 2
cudaError_t BGSA_CAT(launch_bitpal_semiglobal_s, BGSA_SCHEME_ID)(int K, int L, const LaunchArgs &a) {
#define X(k, l) \
    if (K == k && L == l) return launch_instance<BitpalPacked<TheScheme, k, BITPAL_SEMIGLOBAL>, l, 2>(a, BitpalParams{0});
    BGSA_BITPAL_PACKED_INSTANCES(X)
#undef X
    return cudaErrorInvalidValue;
}
#elif BGSA_PACKED
cudaError_t BGSA_CAT(launch_bitpal_packed_s, BGSA_SCHEME_ID)(int K, int L, const LaunchArgs &a) {
#define X(k, l) \
    if (K == k && L == l) return launch_instance<BitpalPacked<TheScheme, k>, l, 2>(a, BitpalParams{0});
    BGSA_BITPAL_PACKED_INSTANCES(X)
#undef X
    return cudaErrorInvalidValue;
}
#else
cudaError_t BGSA_CAT(launch_bitpal_nonpacked_s, BGSA_SCHEME_ID)(int K, int L, const LaunchArgs &a) {
#define X(k, l) \
    if (K == k && L == l) return launch_instance<BitpalNonPacked<TheScheme, k>, l, 1>(a, BitpalParams{0});
    BGSA_BITPAL_NONPACKED_INSTANCES(X)
#undef X
    return cudaErrorInvalidValue;
}
#endif

}  // namespace bgsa
