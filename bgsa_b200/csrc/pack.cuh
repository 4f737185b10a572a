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

// ---------------------------------------------------------------------------------------------
// pack_stream_kernel -- the HBM-rate version of pack_kernel (same output, bit for bit).
//
// A tile (32 consecutive rows) is ONE contiguous run of 32*(slen+1) bytes.  One warp owns a tile:
//   step 1 (stream order, coalesced): every lane loads 16-byte pieces of the run with LDG.128
//          (4 in flight per lane), encodes 16 ASCII bytes -> one 32-bit word with SWAR logic (no
//          per-byte work, see encode_piece) and stores it to the warp's shared-memory strip;
//   step 2 (subject order): lane L cuts its own row out of the strip -- the row starts at an
//          arbitrary bit offset, so every 128-bit unit is five LDS + four funnel shifts -- and
//          stores it to the tile layout, 512 contiguous bytes per warp store.
// No block-level synchronisation (__syncwarp only).  Traffic: slen+1 bytes in, ~slen/4 out per
// subject; the N plane is written only for tiles that contain an 'N' (the alignment kernels never
// read it otherwise, tile_has_n).
//
// SWAR encoding of one 32-bit word w = 4 ASCII bytes (exhaustively checked in
// tests/test_host_logic.py::test_pack_swar_model):
//   code  = (b1 ^ b2, b2 ^ b3) per byte b                A,C,G,T -> 0,1,2,3
//   valid = byte in {A,C,G,T}  <=>  (b & 0xE8) == 0x40  and  b0 != b4  and  b4 ^ (~b2 | b1)
//   the four 2-bit codes are gathered into one byte by an integer multiply (FMA pipe) whose
//   partial products do not overlap, the bytes of four words are merged with PRMT.
// A piece that holds anything else ('N', lower case, ...; the '\n' that ends a row is excused
// by position) takes the exact per-byte path.  Alphabet: original/BGSA_CPU/global.c:9-15.
// ---------------------------------------------------------------------------------------------
constexpr int kPackUnroll = 4;

__host__ __device__ constexpr int pack_phys(int idx) { return idx + (idx >> 5); }   // one pad word per 32: kills bank conflicts of power-of-two row pitches
// shared-memory words one warp needs for a pass over `g` tiles of rows of `stride` bytes
__host__ __device__ inline int pack_strip_pieces(int stride, int g) { return 2 * stride * g + 3; }   // ceil((15 + 32*g*stride)/16) + 1
__host__ __device__ inline int pack_code_words(int stride, int g) { return pack_phys(pack_strip_pieces(stride, g) + 8) + 1; }
__host__ __device__ inline int pack_warp_words(int stride, int g) { return pack_code_words(stride, g) + (pack_strip_pieces(stride, g) + 8 + 1) / 2 + 1; }
// tiles per pass: short rows are grouped so that a pass holds ~1000 pieces (whole groups of 128 dominate, the
// per-pass overhead is amortised) -- rows of 100 bases: 6 tiles, 150: 4, >= 512: 1
__host__ __device__ inline int pack_tiles_per_pass(int stride) { const int g = (512 + stride - 1) / stride; return g < 1 ? 1 : (g > 8 ? 8 : g); }

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n) {    // PTX shl clamps n > 31 to "all bits out"
    uint32_t d; asm("shl.b32 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(n)); return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
}

// The exact per-byte path (rare: only pieces that hold something besides A,C,G,T).  Kept out of the unrolled
// streaming loop on purpose: a data-dependent branch inside that loop crashes ptxas 12.9 and bloats the hot path.
template <int LAYOUT>
__device__ __noinline__ uint2 encode_piece_exact(const uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t cw = 0u, nbits = 0u;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 0xffu;
        const uint32_t code = (c == 'C') ? 1u : (c == 'G') ? 2u : (c == 'T') ? 3u : 0u;
        if (LAYOUT == LAYOUT_CODES) cw |= code << (2 * i);
        else cw |= ((code & 1u) << i) | ((code >> 1) << (16 + i));
        nbits |= (c == 'N' ? 1u : 0u) << i;
    }
    return make_uint2(cw, nbits);
}

// 16 ASCII bytes -> packed word; returns non-zero when the piece holds a byte outside {A,C,G,T} (then the word is
// meaningless and encode_piece_exact must be used).  e = index of the row-end byte inside the piece (>= 16: none).
// CODES : base i at bits 2i.   PLANES: low code bit of base i at bit i, high code bit at bit 16+i.
template <int LAYOUT>
__device__ __forceinline__ uint32_t encode_piece(const uint4 v, int e, uint32_t &cw) {
    const uint32_t c40 = 0x40404040u, cE8 = 0xE8E8E8E8u, c10 = 0x10101010u;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t x[4], xh[4], accbad = 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        // Left shifts only where possible: they compile to IMAD.SHL (FMA pipe), the ALU pipe is this kernel's bound.
        // code bits (b1^b2, b2^b3) land at bits 1-2 of every byte, the two validity tests at bit 4.
        const uint32_t s1 = w[j] >> 1, l2 = w[j] << 2, l3 = w[j] << 3, l4 = w[j] << 4;
        if (LAYOUT == LAYOUT_CODES) {
            x[j] = lop3<(LA ^ LB) & LC>(w[j], s1, 0x06060606u);
        } else {
            x[j] = lop3<(LA ^ LB) & LC>(w[j], s1, 0x02020202u);              // low code bit
            xh[j] = lop3<(LA ^ LB) & LC>(w[j], s1, 0x04040404u);             // high code bit
        }
        const uint32_t v2 = lop3<LA ^ ((0xFF ^ LB) | LC)>(w[j], l2, l3);     // bit 4: b4 ^ (~b2 | b1)
        const uint32_t g = lop3<(LA ^ LB) & LC>(w[j], l4, v2);               // bit 4: (b4 ^ b0) & v2
        uint32_t bad = lop3<(LA ^ LB) & LC>(w[j], c40, cE8);                 // (b & 0xE8) != 0x40
        bad = lop3<LA | ((0xFF ^ LB) & LC)>(bad, g, c10);
        const uint32_t excuse = shl_clamp(0xffu, (uint32_t)(8 * e - 32 * j));     // the row-end byte may be anything
        accbad = lop3<LA | (LB & (0xFF ^ LC))>(accbad, bad, excuse);
    }
    // Gather by integer multiply (FMA pipe): the partial products of the chosen constants never overlap, the
    // wanted bits end up in the top byte; PRMT merges the top bytes.
    if (LAYOUT == LAYOUT_CODES) {
        const uint32_t K = (1u << 23) | (1u << 17) | (1u << 11) | (1u << 5);         // 4 x 2 bits of one word
        const uint32_t t01 = prmt(x[0] * K, x[1] * K, 0x0073u), t23 = prmt(x[2] * K, x[3] * K, 0x0073u);
        cw = prmt(t01, t23, 0x5410u);
    } else {
        const uint32_t KP = (1u << 23) | (1u << 16) | (1u << 9) | (1u << 2);         // 8 x 1 bit of two words (second << 4)
        const uint32_t lo01 = (x[1] * 16u + x[0]) * KP, lo23 = (x[3] * 16u + x[2]) * KP;
        const uint32_t hi01 = (xh[1] * 16u + xh[0]) * (KP >> 1), hi23 = (xh[3] * 16u + xh[2]) * (KP >> 1);
        cw = prmt(prmt(lo01, lo23, 0x0073u), prmt(hi01, hi23, 0x0073u), 0x5410u);    // lo16 | hi16 << 16
    }
    return accbad;
}

// Step 1 of the streaming pack for one warp (also used by the fused banded kernel): encodes the `npieces` 16-byte
// pieces at `src` into the warp's strip (s_c: one word per piece, s_n: its N bits) and returns the OR of the N bits
// this lane produced.  m0 = position of the lane's first piece inside its row, inc = 512 % stride.
template <int LAYOUT, bool PREFETCH = true>
__device__ __forceinline__ uint32_t pack_run_to_strip(const uint4 *__restrict__ src, int npieces, int slen, int stride, int inc,
                                                      int m0, int lane, uint32_t *s_c, uint16_t *s_n) {
    // stream order.  Whole groups of kPackUnroll x 32 pieces first (no bounds checks, loads in
    // flight before the first is used), then the remainder one piece per lane and trip.
    uint32_t any_n = 0u;
    int m = m0;
    const int full = npieces - npieces % (32 * kPackUnroll);
    int p0 = lane;
    // current / next group (software prefetch; not worth its 16 registers when a run is a single short tile)
    uint4 v[kPackUnroll], nv[PREFETCH ? kPackUnroll : 1];
    if (PREFETCH && p0 < full) {
#pragma unroll
        for (int k = 0; k < kPackUnroll; k++) v[k] = __ldg(src + p0 + 32 * k);
    }
    for (; p0 < full; p0 += 32 * kPackUnroll) {
        if (PREFETCH) {
            if (p0 + 32 * kPackUnroll < full) {                  // warp-uniform
#pragma unroll
                for (int k = 0; k < kPackUnroll; k++) nv[k] = __ldg(src + p0 + 32 * kPackUnroll + 32 * k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kPackUnroll; k++) v[k] = __ldg(src + p0 + 32 * k);
        }
        const int ph = pack_phys(p0);                           // pack_phys(p0 + 32k) = ph + 33k
        uint32_t redo = 0u;                                     // bit k: piece k needs the exact path
#pragma unroll
        for (int k = 0; k < kPackUnroll; k++) {
            uint32_t cw;
            const uint32_t bad = encode_piece<LAYOUT>(v[k], slen - m, cw);
            m += inc;
            if (m >= stride) m -= stride;
            s_c[ph + 33 * k] = cw;
            s_n[p0 + 32 * k] = (uint16_t)0;
            redo |= (bad != 0u ? 1u : 0u) << k;
        }
        while (redo) {                                          // rare: 'N', lower case, ...
            const int k = __ffs(redo) - 1;
            redo &= redo - 1u;
            const int p = p0 + 32 * k;
            const uint2 r = encode_piece_exact<LAYOUT>(__ldg(src + p));
            s_c[pack_phys(p)] = r.x;
            s_n[p] = (uint16_t)r.y;
            any_n |= r.y;
        }
        if (PREFETCH) {
#pragma unroll
            for (int k = 0; k < kPackUnroll; k++) v[k] = nv[k];
        }
    }
    uint32_t redo_tail = 0u;                                    // bit i: tail piece p0 + 32 i needs the exact path
    for (int p = p0, i = 0; p < npieces; p += 32, i++) {
        uint32_t cw;
        const uint32_t bad = encode_piece<LAYOUT>(__ldg(src + p), slen - m, cw);
        m += inc;
        if (m >= stride) m -= stride;
        s_c[pack_phys(p)] = cw;
        s_n[p] = (uint16_t)0;
        redo_tail |= (bad != 0u ? 1u : 0u) << i;
    }
    while (redo_tail) {
        const int p = p0 + 32 * (__ffs(redo_tail) - 1);
        redo_tail &= redo_tail - 1u;
        const uint2 r = encode_piece_exact<LAYOUT>(__ldg(src + p));
        s_c[pack_phys(p)] = r.x;
        s_n[p] = (uint16_t)r.y;
        any_n |= r.y;
    }
    return any_n;
}

template <int LAYOUT>
__global__ void __launch_bounds__(128)
pack_stream_kernel(const uint8_t *__restrict__ rows, int slen, long long count, uint4 *__restrict__ codes,
                   uint32_t *__restrict__ nmask, uint8_t *__restrict__ tile_has_n, long long ntiles, int ku, int kn, int G) {
    extern __shared__ __align__(16) uint32_t s_pack[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int stride = slen + 1;
    uint32_t *s_c = s_pack + warp * pack_warp_words(stride, G);
    uint16_t *s_n = reinterpret_cast<uint16_t *>(s_c + pack_code_words(stride, G));
    // tile starts are multiples of 32 bytes from `rows`: the misalignment of the run is the same for every tile
    const int off = (int)(reinterpret_cast<uintptr_t>(rows) & 15);
    const int inc = 512 % stride;                                   // advance of (position mod stride) per loop trip
    int m0 = (16 * lane - off) % stride;                            // position of this lane's first piece inside its row
    if (m0 < 0) m0 += stride;
    const long long npasses = (ntiles + G - 1) / G;

    // one pass = G consecutive tiles = one contiguous run of up to 32*G rows
    for (long long pass = (long long)blockIdx.x * warps + warp; pass < npasses; pass += (long long)gridDim.x * warps) {
        const long long tile0 = pass * G;
        const long long first = tile0 * kTileSubjects;
        const int pass_rows = (int)min((long long)kTileSubjects * G, count - first);
        const uint4 *src = reinterpret_cast<const uint4 *>(rows + first * stride - off);
        const int npieces = (off + pass_rows * stride + 15) >> 4;
        // ---- step 1: stream order
        const uint32_t any_n = pack_run_to_strip<LAYOUT>(src, npieces, slen, stride, inc, m0, lane, s_c, s_n);
        const bool pass_n = __ballot_sync(0xffffffffu, any_n != 0u) != 0u;
        __syncwarp();
        // ---- step 2: subject order, tile by tile
        for (int g = 0; g < G; g++) {
            const long long tile = tile0 + g;
            if (tile >= ntiles) break;
            const int base_pos = off + (g * kTileSubjects + lane) * stride;   // strip position of this lane's own row
            const int wi0 = base_pos >> 4, sub = base_pos & 15;
            const bool live = g * kTileSubjects + lane < pass_rows;
            for (int u = 0; u < ku; u++) {
                uint4 outv = make_uint4(0u, 0u, 0u, 0u);
                if (live) {
                    uint32_t t[5];
#pragma unroll
                    for (int i = 0; i < 5; i++) t[i] = s_c[pack_phys(wi0 + 4 * u + i)];
                    if (LAYOUT == LAYOUT_CODES) {
                        const int sh = 2 * sub;
                        outv.x = __funnelshift_r(t[0], t[1], sh); outv.y = __funnelshift_r(t[1], t[2], sh);
                        outv.z = __funnelshift_r(t[2], t[3], sh); outv.w = __funnelshift_r(t[3], t[4], sh);
                        const int nb = slen - u * kBasesPerUnit;          // bases of this unit that exist
                        if (nb < kBasesPerUnit) {
                            auto keep = [](int n) { return n >= 16 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << (2 * n)) - 1u)); };
                            outv.x &= keep(nb); outv.y &= keep(nb - 16); outv.z &= keep(nb - 32); outv.w &= keep(nb - 48);
                        }
                    } else {
                        // strip word = lo16 | hi16 << 16 per 16 bases
                        const uint32_t lo01 = prmt(t[0], t[1], 0x5410u), lo23 = prmt(t[2], t[3], 0x5410u);
                        const uint32_t hi01 = prmt(t[0], t[1], 0x7632u), hi23 = prmt(t[2], t[3], 0x7632u);
                        outv.x = __funnelshift_r(lo01, lo23, sub); outv.y = __funnelshift_r(hi01, hi23, sub);
                        outv.z = __funnelshift_r(lo23, t[4] & 0xffffu, sub); outv.w = __funnelshift_r(hi23, t[4] >> 16, sub);
                        const int nb = slen - u * kBasesPerUnit;
                        if (nb < kBasesPerUnit) {
                            auto keep = [](int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); };
                            outv.x &= keep(nb); outv.y &= keep(nb); outv.z &= keep(nb - 32); outv.w &= keep(nb - 32);
                        }
                    }
                }
                codes[(tile * ku + u) * 32 + lane] = outv;
            }
            // N plane: only for tiles that really hold an 'N' (rare: decided per tile, the pass flag only says "look")
            bool flag = false;
            if (pass_n) {
                auto nword = [&](int k) {
                    uint32_t nm = 0u;
                    if (live) {
                        const int q = wi0 + 2 * k;
                        const uint32_t a = (uint32_t)s_n[q] | ((uint32_t)s_n[q + 1] << 16), b = s_n[q + 2];
                        nm = __funnelshift_r(a, b, sub);
                        const int nb = slen - 32 * k;
                        if (nb < 32) nm &= (1u << nb) - 1u;
                    }
                    return nm;
                };
                uint32_t row_n = 0u;
                for (int k = 0; k < kn; k++) row_n |= nword(k);
                flag = __ballot_sync(0xffffffffu, row_n != 0u) != 0u;
                if (flag)
                    for (int k = 0; k < kn; k++) nmask[(tile * kn + k) * 32 + lane] = nword(k);
            }
            if (lane == 0) tile_has_n[tile] = flag ? 1 : 0;
        }
        __syncwarp();      // the strip is rewritten by the next pass
    }
}

// Readers of a PLANES strip (word per 16 bases = low plane | high plane << 16) for a row that starts at strip position
// wi0*16 + sub: the low / high bit-plane word (32 bases) number k of the row, bases beyond slen cleared.
__device__ __forceinline__ void strip_plane_word(const uint32_t *s_c, int wi0, int sub, int k, int slen, uint32_t &lo, uint32_t &hi) {
    const uint32_t t0 = s_c[pack_phys(wi0 + 2 * k)], t1 = s_c[pack_phys(wi0 + 2 * k + 1)], t2 = s_c[pack_phys(wi0 + 2 * k + 2)];
    lo = __funnelshift_r(prmt(t0, t1, 0x5410u), t2 & 0xffffu, sub);
    hi = __funnelshift_r(prmt(t0, t1, 0x7632u), t2 >> 16, sub);
    const int nb = slen - 32 * k;
    if (nb < 32) { const uint32_t keep = nb <= 0 ? 0u : ((1u << nb) - 1u); lo &= keep; hi &= keep; }
}
__device__ __forceinline__ uint32_t strip_n_word(const uint16_t *s_n, int wi0, int sub, int k, int slen) {
    const int q = wi0 + 2 * k;
    const uint32_t a = (uint32_t)s_n[q] | ((uint32_t)s_n[q + 1] << 16), b = s_n[q + 2];
    uint32_t nm = __funnelshift_r(a, b, sub);
    const int nb = slen - 32 * k;
    if (nb < 32) nm &= nb <= 0 ? 0u : ((1u << nb) - 1u);
    return nm;
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
