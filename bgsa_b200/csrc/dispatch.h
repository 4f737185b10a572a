// dispatch.h -- which kernel instance serves a (algorithm, scoring scheme, query length).
//
// The reference bakes these choices into the generated align_core.c (one file per algorithm x
// scheme x SIMD width, generator/.../Main.java:240-315); here they are template instances picked
// at run time.  K = words per lane, L = lanes per subject; the bit-vector has K*L >= ceil(m/32) words.
#pragma once
#include <stdint.h>

namespace bgsa {

struct Geometry { int K; int L; };

// smallest L (power of two <= 32) for which ceil(W/L) <= kmax; then K = ceil(W/L) rounded up to an
// instantiated value.  Returns {0,0} when the query is too long for the instantiated kernels.
inline Geometry pick_geometry(int qlen, const int *ks, int nks, int kmax_single) {
    const int W = (qlen + 31) / 32;
    for (int L = 1; L <= 32; L <<= 1) {
        const int need = (W + L - 1) / L;
        for (int i = 0; i < nks; i++) {
            const int K = ks[i];
            if (K < need) continue;
            if (L == 1 && K > kmax_single) break;
            if (L > 1 && K > 8) break;
            return Geometry{K, L};
        }
    }
    return Geometry{0, 0};
}

}  // namespace bgsa
