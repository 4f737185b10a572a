// inst_myers.cu -- Myers global / semi-global kernel instances (see instances.h, myers.cuh).
#include "instances.h"
#include "launch.cuh"
#include "myers.cuh"

namespace bgsa {

#ifndef BGSA_MYERS_MODE
#error "compile with -DBGSA_MYERS_MODE=0 (global) or 1 (semi-global)"
#endif

#if BGSA_MYERS_MODE == 0
cudaError_t launch_myers_global(int K, int L, const LaunchArgs &a, int sign) {
#else
cudaError_t launch_myers_semiglobal(int K, int L, const LaunchArgs &a, int sign) {
#endif
    MyersParams prm{sign};
#define X(k, l) \
    if (K == k && L == l) return launch_instance<MyersAlgo<k, BGSA_MYERS_MODE>, l, (l > 1 && k >= 24 ? 1 : (k <= 8 ? 4 : 2))>(a, prm);   // wavefront with wide lanes: unrolling only costs registers
    BGSA_MYERS_INSTANCES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace bgsa
