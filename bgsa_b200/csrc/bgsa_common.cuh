// bgsa_common.cuh -- device-side building blocks shared by the sm_100a alignment kernels.
//
// Conventions used by every kernel in this directory (see DESIGN.md "Data layout in HBM"):
//   * DP orientation is TRANSPOSED with respect to the reference for the symmetric problems
//     (Myers global / semi-global, BitPAl): the bit-vector runs along the QUERY, whose match
//     masks (Peq, 5 rows x W words) are built once per launch and live in shared memory; each
//     subject is streamed one base per DP column.  The reference does the opposite (Peq per
//     subject, original/BGSA_CPU/global.c:25-70).  Scores are identical because the DP matrix is
//     the same matrix.
//   * Subjects live in HBM as TILES of 32 subjects, 2 bits per base, 128-bit units:
//       codes[(tile * KU + k) * 32 + lane]  (uint4)  = bases 64k .. 64k+63 of subject tile*32+lane,
//     base i at bits 2*(i%16) of 32-bit word (i/16)%4.  A warp reads 512 contiguous bytes per k.
//     'N' is carried in a separate 1-bit plane nmask[(tile * KN + k) * 32 + lane] (uint32) and a
//     per-tile flag says whether the plane needs to be looked at at all.
#pragma once

// The kernel headers (this file, align_kernel.cuh, rows_kernel.cuh, myers.cuh, bitpal.cuh) also compile under NVRTC:
// jit.cu instantiates BitPAl kernels for scoring schemes that were not built into the library (the run-time counterpart
// of the reference's Java generator).  NVRTC has no host headers, hence the typedefs.
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace bgsa {

constexpr int kTileSubjects = 32;   // one warp's worth of subjects
constexpr int kPeqRows = 5;         // A C G T N  (CHAR_NUM, original/BGSA_CPU/config.h:17)
constexpr int kBasesPerUnit = 64;   // bases per uint4

// ---------------------------------------------------------------------------------------------
// packed subject set (device pointers into ONE allocation, see bgsa_packed_bytes())
// ---------------------------------------------------------------------------------------------
struct PackedSubjects {
    const uint4 *codes;          // [ntiles][ku][32]
    const uint32_t *nmask;       // [ntiles][kn][32]
    const uint8_t *tile_has_n;   // [ntiles]
    int64_t count;               // subjects
    int64_t ntiles;
    int slen;
    int ku;                      // ceil(slen / 64)
    int kn;                      // ceil(slen / 32)
};

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

#ifndef __CUDACC_RTC__
__host__ inline PackedSubjects make_packed_view(void *base, int slen, int64_t count) {
    PackedSubjects v;
    v.count = count;
    v.slen = slen;
    v.ntiles = (count + kTileSubjects - 1) / kTileSubjects;
    v.ku = (slen + kBasesPerUnit - 1) / kBasesPerUnit;
    v.kn = (slen + 31) / 32;
    char *p = static_cast<char *>(base);
    v.codes = reinterpret_cast<const uint4 *>(p);
    p += align_up(v.ntiles * v.ku * 32 * (int64_t)sizeof(uint4), 256);
    v.nmask = reinterpret_cast<const uint32_t *>(p);
    p += align_up(v.ntiles * v.kn * 32 * (int64_t)sizeof(uint32_t), 256);
    v.tile_has_n = reinterpret_cast<const uint8_t *>(p);
    return v;
}

__host__ inline int64_t packed_bytes(int slen, int64_t count) {
    int64_t ntiles = (count + kTileSubjects - 1) / kTileSubjects;
    int64_t ku = (slen + kBasesPerUnit - 1) / kBasesPerUnit, kn = (slen + 31) / 32;
    return align_up(ntiles * ku * 32 * (int64_t)sizeof(uint4), 256) +
           align_up(ntiles * kn * 32 * (int64_t)sizeof(uint32_t), 256) + align_up(ntiles, 256);
}
#endif

// ---------------------------------------------------------------------------------------------
// integer-pipe primitives
// ---------------------------------------------------------------------------------------------
// Every DP building block is __host__ __device__: the device path is inline PTX, the host path is
// a bit-exact emulation used ONLY by tests/host_sim (algorithm checks in a container without GPU).
#define BGSA_HD __host__ __device__ __forceinline__

// 3-input logic op with an 8-bit truth table (LOP3.LUT).  Table built from 0xF0 (a), 0xCC (b), 0xAA (c).
template <int LUT>
BGSA_HD uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
#else
    uint32_t d = 0;
    for (int i = 0; i < 8; i++)
        if ((LUT >> i) & 1) d |= ((i & 4) ? a : ~a) & ((i & 2) ? b : ~b) & ((i & 1) ? c : ~c);
    return d;
#endif
}
constexpr int LA = 0xF0, LB = 0xCC, LC = 0xAA;

// (cur << 1) | (prev >> 31): the one-position shift of a multi-word bit-vector (SHF.L.W.U32.HI)
BGSA_HD uint32_t shl1_carry(uint32_t prev, uint32_t cur) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(prev, cur, 1);
#else
    return (cur << 1) | (prev >> 31);
#endif
}
// (cur << 1) | BIT with a CONSTANT shift-in bit (word 0 of a vector in the thread-per-subject kernels, where nothing
// comes from a neighbouring lane): an integer multiply-add, i.e. the FMA pipe instead of the ALU pipe that bounds the
// kernels.  Only the plain 32-bit IMAD qualifies: IMAD.WIDE / IMAD.HI do not overlap with ALU work (tools/pipe_probe.cu,
// profiles/r02_pipe_probe.log), so the words above word 0 -- whose shift-in bit is the top bit of the word below -- stay
// funnel shifts.  The multiplier is a run-time value on purpose: with an immediate ptxas strength-reduces the multiply
// to SHF/LEA, which are ALU-pipe instructions again.
#ifndef BGSA_FMA_SHIFT0
#define BGSA_FMA_SHIFT0 1
#endif
#ifdef __CUDACC__
static __constant__ uint32_t c_bgsa_two = 2u;      // (one copy per translation unit; never written)
#endif
template <int BIT>
BGSA_HD uint32_t shl1_const(uint32_t cur) {
#ifdef __CUDA_ARCH__
#if BGSA_FMA_SHIFT0
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(cur), "r"(c_bgsa_two), "n"(BIT));
    return d;
#else
    return __funnelshift_l(BIT ? 0x80000000u : 0u, cur, 1);
#endif
#else
    return (cur << 1) | (uint32_t)BIT;
#endif
}
BGSA_HD int popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// Multi-word add with the hardware carry flag: one IADD3(.X) per word.  The statements are
// volatile so that nothing is scheduled between the links of one chain; nvcc itself never emits
// .cc instructions, so the flag survives from one asm statement to the next.
#ifndef __CUDA_ARCH__
inline uint32_t &host_carry_flag() { static thread_local uint32_t cf = 0; return cf; }
#endif
BGSA_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t d; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
#else
    uint64_t s = (uint64_t)a + b; host_carry_flag() = (uint32_t)(s >> 32); return (uint32_t)s;
#endif
}
BGSA_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t d; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
#else
    uint64_t s = (uint64_t)a + b + host_carry_flag(); host_carry_flag() = (uint32_t)(s >> 32); return (uint32_t)s;
#endif
}
BGSA_HD uint32_t addc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t d; asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
#else
    return a + b + host_carry_flag();
#endif
}

// out = a + b (+ CF) over K words; the carry out stays in the hardware flag.  CF_IN: the caller has
// just loaded the flag (CarryIn::to_cf); otherwise the chain starts with a plain add.  The
// thread-per-subject kernels (no neighbours) pay exactly K instructions.
template <int K, bool CF_IN>
BGSA_HD void add_chain(uint32_t (&out)[K], const uint32_t (&a)[K], const uint32_t (&b)[K]) {
    out[0] = CF_IN ? addc_cc(a[0], b[0]) : add_cc(a[0], b[0]);
#pragma unroll
    for (int j = 1; j < K; j++) out[j] = addc_cc(a[j], b[j]);
}

// Carry stream between neighbouring lanes of a wavefront group (align_kernel.cuh): the bits a lane
// hands down (add carries, shift-in bits) travel in ONE 32-bit word, top-aligned, in the order the
// receiver consumes them, the subject base in bits 0..2.  One instruction per bit on either side:
//   receiver: add.cc w,w   moves the next bit into the hardware carry flag AND advances the word;
//             a shift-in bit is used in place (funnel shifts only look at bit 31), then w += w.
//   sender  : addc w,w     appends the carry flag;  funnel-shift appends the top bit of a vector.
struct CarryIn {
    uint32_t w;
    BGSA_HD explicit CarryIn(uint32_t recv) : w(recv) {}
    BGSA_HD void to_cf() { w = add_cc(w, w); }                      // CF := next bit
    BGSA_HD uint32_t top() { const uint32_t t = w; w += w; return t; }  // next bit at bit 31 of the result
};
struct CarryOut {
    uint32_t w = 0u;
    BGSA_HD void push_cf() { w = addc(w, w); }                      // append CF
    BGSA_HD void push_top(uint32_t v) { w = shl1_carry(v, w); }     // append bit 31 of v
    template <int NBITS> BGSA_HD uint32_t finish() const { static_assert(NBITS <= 29, "carry word overflow"); return w << (32 - NBITS); }
};

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) for staging subject tiles
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion signalled on `bar` (bytes multiple of 16, both sides 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// Per-warp double-buffered stream of subject tiles.
//
// Every warp owns 2 stages x CH x 32 uint4 of shared memory and two mbarriers.  A "stage" is
// CH consecutive 128-bit units (64*CH bases) of the warp's current tile, contiguous in HBM by
// construction of the tile layout, so one bulk copy per stage suffices.  Lane 0 issues, all 32
// lanes wait on the barrier; no block-level synchronisation is involved, which lets warps drift
// apart (banded early exit) and fetch tiles from the global work counter independently.
// ---------------------------------------------------------------------------------------------
template <int CH>
struct WarpStage {
    uint4 *buf;          // [2][CH*32]
    uint64_t *bar;       // [2]
    uint32_t phases;     // bit s = parity to wait for on stage s

    __device__ __forceinline__ void init(uint4 *warp_buf, uint64_t *warp_bar, int lane) {
        buf = warp_buf; bar = warp_bar; phases = 0;
        if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
        __syncwarp();
    }
    // units = number of uint4 per lane in this stage (<= CH)
    __device__ __forceinline__ void issue(int s, const uint4 *src, int units, int lane) {
        if (lane == 0) {
            uint32_t bytes = (uint32_t)units * 32u * (uint32_t)sizeof(uint4);
            mbar_expect_tx(&bar[s], bytes);
            bulk_g2s(buf + s * CH * 32, src, bytes, &bar[s]);
        }
    }
    __device__ __forceinline__ void wait(int s) { mbar_wait(&bar[s], (phases >> s) & 1u); phases ^= 1u << s; }
    __device__ __forceinline__ uint4 load(int s, int unit, int lane) const { return buf[(s * CH + unit) * 32 + lane]; }
};

// next tile from the global work counter (dynamic scheduling; one atomic per 32 subjects)
__device__ __forceinline__ long long next_tile(unsigned long long *counter, int lane) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(counter, 1ULL);
    return (long long)__shfl_sync(0xffffffffu, t, 0);
}

// Shared-memory Peq row stride (in 32-bit words) for a K-word query: multiple of 4 words so rows
// can be read with LDS.128, and an odd number of 16-byte units so the 5 rows start in different
// bank groups (lanes holding different bases read different rows of the same column).
__host__ __device__ constexpr int peq_stride(int k) { return ((k + 3) / 4 | 1) * 4; }

// Narrowing of the reference result store: the kernel computes in a 32-bit lane, the low 32 bits
// are stored to int16_t (original/BGSA_CPU/align_core.c:139-144, config.h:19).
BGSA_HD int16_t narrow16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }

}  // namespace bgsa
