// banded_host.h -- host-side per-row mask table for the banded kernel (see banded.cuh).
#pragma once
#include <stdint.h>

namespace bgsa {

struct BandedRowHost {
    uint32_t nclo, nchi, bm_lo, bm_hi, cn_lo, cn_hi, pad0, pad1;
};

// Row r of the band covers subject indices i = r + b - e - 1 for band bits b = 0..2e (equal
// lengths: h_threshold = e, band_down = 2e; banded/BGSA_CPU/align_core.c:70-72 and the Peq
// placement of banded/BGSA_CPU/global.c:45-82).  Cells with i < 0 or i >= slen never match.
inline void build_banded_table(const char *qcodes, int qlen, int slen, int e, BandedRowHost *out) {
    for (int r = 0; r < qlen; r++) {
        int c = (unsigned char)qcodes[r];
        if (c > 4) c = 0;
        uint64_t geom = 0;
        for (int b = 0; b <= 2 * e; b++) {
            const long long i = (long long)r + b - e - 1;
            if (i >= 0 && i < slen) geom |= 1ULL << b;
        }
        BandedRowHost row;
        row.nclo = (c & 1) ? 0u : 0xffffffffu;
        row.nchi = (c & 2) ? 0u : 0xffffffffu;
        const uint64_t bm = (c == 4) ? 0ULL : geom;
        const uint64_t cn = (c == 4) ? geom : 0ULL;
        row.bm_lo = (uint32_t)bm; row.bm_hi = (uint32_t)(bm >> 32);
        row.cn_lo = (uint32_t)cn; row.cn_hi = (uint32_t)(cn >> 32);
        row.pad0 = row.pad1 = 0u;
        out[r] = row;
    }
}

// The reference's banded Peq builder writes char_index up to 1 + (len-1)/64 but allocates
// word_num = ceil((len - e)/64) + 1 words per class (banded/BGSA_CPU/global.c:67-82 vs
// cal_cpu.c:253-254): when that overflows it corrupts the neighbouring masks and the reference's
// own output is garbage.  Parity is only defined where this predicate holds.
inline bool banded_reference_in_bounds(int len, int e) {
    return (len - 1) / 64 + 1 < (len - e + 63) / 64 + 1;
}

}  // namespace bgsa
