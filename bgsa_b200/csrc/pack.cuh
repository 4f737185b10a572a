// pack.cuh -- ASCII subject rows -> packed, transposed tiles (replaces the per-subject Peq build
// cpu_handle_reads, original/BGSA_CPU/global.c:25-70 and banded/BGSA_CPU/global.c:25-84).
//
// Input : rows of stride slen+1 exactly as get_read_from_file() leaves them in seq_t.content
//         (file.c:44-115), already in device memory.
// Output: the tile layout described in bgsa_common.cuh.  Two encodings of the 128-bit unit:
//   LAYOUT_CODES  (Myers / BitPAl): base i of the unit at bits 2*(i%16) of word i/16
//   LAYOUT_PLANES (banded)        : x,y = low/high code bit of bases 0..31, z,w = of bases 32..63
// plus the N bit-plane and the per-tile "contains N" flag.
// Alphabet (global.c:9-15): A,C,G,T -> 0..3, N -> code 0 with its N bit set, anything else -> 0.
//
// One warp per tile, one lane per subject.  A lane walks its own row in aligned 32-bit words
// (row starts are not 4-byte aligned in general; two words are funnel-shifted together), so the
// global reads are sector-granular rather than fully coalesced -- acceptable because this kernel
// moves 1/100th of the time of the alignment kernels it feeds ("pack").
#pragma once

#include "bgsa_common.cuh"

namespace bgsa {

enum { LAYOUT_CODES = 0, LAYOUT_PLANES = 1 };

// 4 ASCII bytes -> (2-bit codes in bits 0..7, N flags in bits 8..11)
__device__ __forceinline__ uint32_t encode4(uint32_t w) {
    uint32_t out = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t c = (w >> (8 * i)) & 0xffu;
        const uint32_t code = (c == 'C') ? 1u : (c == 'G') ? 2u : (c == 'T') ? 3u : 0u;
        out |= code << (2 * i);
        out |= (c == 'N' ? 1u : 0u) << (8 + i);
    }
    return out;
}

template <int LAYOUT>
__global__ void __launch_bounds__(128)
pack_kernel(const uint8_t *__restrict__ rows, int slen, long long count, uint4 *__restrict__ codes,
            uint32_t *__restrict__ nmask, uint8_t *__restrict__ tile_has_n, long long ntiles, int ku, int kn) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    // last aligned word that still overlaps the buffer (reads beyond it are suppressed)
    const uint32_t *last_word = reinterpret_cast<const uint32_t *>(
        (reinterpret_cast<uintptr_t>(rows) + (uintptr_t)(count * (long long)(slen + 1)) - 1) & ~(uintptr_t)3);
    for (long long tile = warp_global; tile < ntiles; tile += nwarps) {
        const long long subject = tile * kTileSubjects + lane;
        const bool live = subject < count;
        const uint8_t *row = rows + subject * (long long)(slen + 1);
        uint32_t any_n = 0u;
        for (int u = 0; u < ku; u++) {
            uint32_t cw[4] = {0u, 0u, 0u, 0u};     // 4 x 16 bases, 2 bits each
            uint32_t nb[2] = {0u, 0u};             // 2 x 32 N flags
            if (live) {
                const int base0 = u * kBasesPerUnit;
                const int ngroups = min(16, (slen - base0 + 3) / 4);  // 4-base groups in this unit
                const uintptr_t p = reinterpret_cast<uintptr_t>(row + base0);
                const uint32_t *a = reinterpret_cast<const uint32_t *>(p & ~(uintptr_t)3);
                const int boff = (int)(p & 3) * 8;
                auto ld = [&](const uint32_t *q) { return q <= last_word ? __ldg(q) : 0u; };
                uint32_t cur = ld(a);
#pragma unroll
                for (int g = 0; g < 16; g++) {
                    if (g < ngroups) {
                        const uint32_t nxt = ld(a + g + 1);
                        uint32_t w = __funnelshift_r(cur, nxt, boff);     // bytes row[base0+4g .. +3]
                        cur = nxt;
                        const int valid = slen - base0 - 4 * g;           // bases left in the row
                        if (valid < 4) w &= (1u << (8 * valid)) - 1u;     // bytes past the row -> 0 (code A, no N)
                        const uint32_t enc = encode4(w);
                        cw[g >> 2] |= (enc & 0xffu) << (8 * (g & 3));
                        nb[g >> 3] |= ((enc >> 8) & 0xfu) << (4 * (g & 7));
                    }
                }
            }
            any_n |= nb[0] | nb[1];
            uint4 outv;
            if (LAYOUT == LAYOUT_CODES) {
                outv = make_uint4(cw[0], cw[1], cw[2], cw[3]);
            } else {
                // de-interleave: even bits -> low plane, odd bits -> high plane
                auto compact = [](uint32_t x) {
                    x &= 0x55555555u;
                    x = (x | (x >> 1)) & 0x33333333u;
                    x = (x | (x >> 2)) & 0x0f0f0f0fu;
                    x = (x | (x >> 4)) & 0x00ff00ffu;
                    x = (x | (x >> 8)) & 0x0000ffffu;
                    return x;
                };
                const uint32_t lo0 = compact(cw[0]) | (compact(cw[1]) << 16), hi0 = compact(cw[0] >> 1) | (compact(cw[1] >> 1) << 16);
                const uint32_t lo1 = compact(cw[2]) | (compact(cw[3]) << 16), hi1 = compact(cw[2] >> 1) | (compact(cw[3] >> 1) << 16);
                outv = make_uint4(lo0, hi0, lo1, hi1);
            }
            codes[(tile * ku + u) * 32 + lane] = outv;
            if (2 * u < kn) nmask[(tile * kn + 2 * u) * 32 + lane] = nb[0];
            if (2 * u + 1 < kn) nmask[(tile * kn + 2 * u + 1) * 32 + lane] = nb[1];
        }
        const uint32_t warp_any = __ballot_sync(0xffffffffu, any_n != 0u);
        if (lane == 0) tile_has_n[tile] = warp_any ? 1 : 0;
    }
}

// Peq -> tiles, for the per-chunk drop-in entry points (include/align_core.h): recovers the subject
// bases from the reference's match masks.  Reference layout (global.c:25-70 and SIMD twins):
// group g = subject / vnum, lane l = subject % vnum, word (c, j) at peq[((g*5 + c)*word_num + j)*vnum + l],
// `usable` cells per word.  head > 0 selects the banded placement (banded/BGSA_CPU/global.c:45-82):
// subject[0..head) at bits head+1.. of word 0, subject[head + i] at bit i%64 of word 1 + i/64.
template <typename WordT, int LAYOUT>
__global__ void unpeq_kernel(const WordT *__restrict__ peq, int word_num, int usable, int head, int slen, long long count,
                             uint4 *__restrict__ codes, uint32_t *__restrict__ nmask, uint8_t *__restrict__ tile_has_n,
                             long long ntiles, int ku, int kn, int vnum) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long tile = warp_global; tile < ntiles; tile += nwarps) {
        const long long subject = tile * kTileSubjects + lane;
        const bool live = subject < count;
        const long long g = subject / vnum, l = subject % vnum;
        uint32_t any_n = 0u;
        for (int u = 0; u < ku; u++) {
            uint32_t lo[2] = {0u, 0u}, hi[2] = {0u, 0u}, nb[2] = {0u, 0u};
            if (live) {
                for (int i = 0; i < kBasesPerUnit; i++) {
                    const int pos = u * kBasesPerUnit + i;
                    if (pos >= slen) break;
                    int j, bit;
                    if (head > 0) {
                        if (pos < head) { j = 0; bit = head + 1 + pos; }
                        else { j = 1 + (pos - head) / usable; bit = (pos - head) % usable; }
                    } else { j = pos / usable; bit = pos % usable; }
                    uint32_t code = 0u, isn = 0u;
#pragma unroll
                    for (int c = 1; c < 5; c++) {
                        const WordT w = peq[((g * 5 + c) * word_num + j) * vnum + l];
                        if ((w >> bit) & 1) { if (c == 4) isn = 1u; else code = (uint32_t)c; }
                    }
                    lo[i >> 5] |= (code & 1u) << (i & 31);
                    hi[i >> 5] |= (code >> 1) << (i & 31);
                    nb[i >> 5] |= isn << (i & 31);
                }
            }
            any_n |= nb[0] | nb[1];
            uint4 outv;
            if (LAYOUT == LAYOUT_PLANES) {
                outv = make_uint4(lo[0], hi[0], lo[1], hi[1]);
            } else {
                // interleave the planes back into 2-bit codes
                auto spread = [](uint32_t x) {       // 16 low bits -> even bit positions
                    x &= 0xffffu;
                    x = (x | (x << 8)) & 0x00ff00ffu;
                    x = (x | (x << 4)) & 0x0f0f0f0fu;
                    x = (x | (x << 2)) & 0x33333333u;
                    x = (x | (x << 1)) & 0x55555555u;
                    return x;
                };
                outv = make_uint4(spread(lo[0]) | (spread(hi[0]) << 1), spread(lo[0] >> 16) | (spread(hi[0] >> 16) << 1),
                                  spread(lo[1]) | (spread(hi[1]) << 1), spread(lo[1] >> 16) | (spread(hi[1] >> 16) << 1));
            }
            codes[(tile * ku + u) * 32 + lane] = outv;
            if (2 * u < kn) nmask[(tile * kn + 2 * u) * 32 + lane] = nb[0];
            if (2 * u + 1 < kn) nmask[(tile * kn + 2 * u + 1) * 32 + lane] = nb[1];
        }
        const uint32_t warp_any = __ballot_sync(0xffffffffu, any_n != 0u);
        if (lane == 0) tile_has_n[tile] = warp_any ? 1 : 0;
    }
}

}  // namespace bgsa
