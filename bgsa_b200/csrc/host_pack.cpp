// host_pack.cpp -- see host_pack.h.  Plain C++ (g++), no CUDA: AVX2 encoder with a scalar twin, and the worker pool.
#include "host_pack.h"

#include <immintrin.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace bgsa {

namespace {

constexpr int kTile = 32;
inline int64_t align_up64(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct View {            // make_packed_view (bgsa_common.cuh), host side
    uint32_t *codes;     // [ntiles][ku][32][4]
    uint32_t *nmask;     // [ntiles][kn][32]
    uint8_t *flags;      // [ntiles]
    int ku, kn;
};
View make_view(void *base, int slen, int64_t count) {
    View v;
    const int64_t ntiles = (count + kTile - 1) / kTile;
    v.ku = (slen + 63) / 64;
    v.kn = (slen + 31) / 32;
    char *p = static_cast<char *>(base);
    v.codes = reinterpret_cast<uint32_t *>(p);
    p += align_up64(ntiles * v.ku * 32 * 16, 256);
    v.nmask = reinterpret_cast<uint32_t *>(p);
    p += align_up64(ntiles * v.kn * 32 * 4, 256);
    v.flags = reinterpret_cast<uint8_t *>(p);
    return v;
}

// ---- 32 ASCII bytes -> (codes or planes) + N bits -------------------------------------------------------------------
// Alphabet (original/BGSA_CPU/global.c:9-15): A,C,G,T -> 0..3, N -> code 0 with its N bit set, anything else -> 0.
struct Enc32 { uint32_t w0, w1, n; };      // CODES: w0 = bases 0..15, w1 = bases 16..31 (2 bits each); PLANES: w0 = low bits, w1 = high bits

inline Enc32 encode32_scalar(const uint8_t *b, int valid, bool planes) {
    Enc32 r{0u, 0u, 0u};
    for (int i = 0; i < valid; i++) {
        const uint8_t c = b[i];
        const uint32_t code = (c == 'C') ? 1u : (c == 'G') ? 2u : (c == 'T') ? 3u : 0u;
        if (planes) { r.w0 |= (code & 1u) << i; r.w1 |= (code >> 1) << i; }
        else if (i < 16) r.w0 |= code << (2 * i);
        else r.w1 |= code << (2 * (i - 16));
        if (c == 'N') r.n |= 1u << i;
    }
    return r;
}

// One 64-base unit of a row -> the 16-byte unit of the tile layout + its two N words.
//   CODES : 4 words of 16 bases, 2 bits each;  PLANES: x,y = low/high bit-plane of bases 0..31, z,w = of bases 32..63
inline void unit_scalar(const uint8_t *p, int left, bool planes, const uint8_t *, uint32_t *out, uint32_t *n0, uint32_t *n1) {
    const Enc32 a = encode32_scalar(p, left < 32 ? left : 32, planes);
    Enc32 b{0u, 0u, 0u};
    if (left > 32) b = encode32_scalar(p + 32, left - 32 < 32 ? left - 32 : 32, planes);
    out[0] = a.w0; out[1] = a.w1; out[2] = b.w0; out[3] = b.w1;
    *n0 = a.n; *n1 = b.n;
}

// dst (16-byte aligned: tiles start on 512-byte boundaries of a 256-byte aligned buffer) <- src, non-temporal (SSE2)
inline void stream_out(uint32_t *dst, const uint32_t *src, size_t bytes) {
    if (reinterpret_cast<uintptr_t>(dst) & 15) { memcpy(dst, src, bytes); return; }      // (caller's buffer not aligned: plain copy)
    __m128i *d = reinterpret_cast<__m128i *>(dst);
    const __m128i *s = reinterpret_cast<const __m128i *>(src);
    for (size_t i = 0; i < bytes / 16; i++) _mm_stream_si128(d + i, _mm_load_si128(s + i));
}

// The tile loop, instantiated once per encoder (the AVX2 copy lives inside a `#pragma GCC target("avx2")` region so that
// the intrinsics inline; the library itself is built for baseline x86-64 and picks at run time).
#define BGSA_PACK_RANGE_BODY(UNIT)                                                                                       \
    const View v = make_view(packed, slen, count);                                                                       \
    const bool planes = layout != 0;                                                                                     \
    const int64_t stride = (int64_t)slen + 1;                                                                            \
    const uint8_t *end = rows + count * stride; /* one past the last byte that may be read */                           \
    std::vector<uint32_t> nbuf((size_t)v.kn * 32, 0u); /* N words of the current tile; all zero between tiles */         \
    /* A tile is assembled in a cache-resident buffer and then streamed out with non-temporal stores: the destination  */ \
    /* (pinned staging) is only read by the DMA engine afterwards, and a thread's speed is bound by the memory traffic */ \
    /* it causes -- plain stores would first READ every destination line (write-allocate).                             */ \
    std::vector<uint32_t> tbuf((size_t)v.ku * 32 * 4 + 16);                                                              \
    uint32_t *tile_out = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(tbuf.data()) + 63) & ~(uintptr_t)63); \
    bool any = false;                                                                                                    \
    for (int64_t tile = t0; tile < t1; tile++) {                                                                         \
        uint32_t tile_n = 0u;                                                                                            \
        for (int lane = 0; lane < 32; lane++) {                                                                          \
            const int64_t subject = tile * kTile + lane;                                                                 \
            if (subject >= count) {                                                                                      \
                for (int k = 0; k < v.ku; k++) memset(tile_out + ((size_t)k * 32 + lane) * 4, 0, 16);                    \
                continue;                                                                                                \
            }                                                                                                            \
            const uint8_t *row = rows + subject * stride;                                                                \
            for (int k = 0; k < v.ku; k++) {                                                                             \
                uint32_t n0, n1;                                                                                         \
                UNIT(row + 64 * k, slen - 64 * k, planes, end, tile_out + ((size_t)k * 32 + lane) * 4, &n0, &n1);        \
                if (n0 | n1) { /* rare */                                                                                \
                    nbuf[(size_t)(2 * k) * 32 + lane] = n0;                                                              \
                    if (2 * k + 1 < v.kn) nbuf[(size_t)(2 * k + 1) * 32 + lane] = n1;                                    \
                    tile_n |= n0 | n1;                                                                                   \
                }                                                                                                        \
            }                                                                                                            \
        }                                                                                                                \
        stream_out(v.codes + tile * v.ku * 32 * 4, tile_out, (size_t)v.ku * 32 * 16);                                    \
        v.flags[tile] = tile_n ? 1 : 0;                                                                                  \
        if (tile_n) {                                                                                                    \
            memcpy(v.nmask + tile * v.kn * 32, nbuf.data(), sizeof(uint32_t) * (size_t)v.kn * 32);                       \
            memset(nbuf.data(), 0, sizeof(uint32_t) * (size_t)v.kn * 32);                                                \
            any = true;                                                                                                  \
        }                                                                                                                \
    }                                                                                                                    \
    _mm_sfence(); /* the streamed tiles are globally visible before the caller hands the buffer to the copy engine */    \
    return any;

bool pack_range_scalar(int layout, const uint8_t *rows, int slen, int64_t count, void *packed, int64_t t0, int64_t t1) {
    BGSA_PACK_RANGE_BODY(unit_scalar)
}

#pragma GCC push_options
#pragma GCC target("avx2")
// `valid` (< 32 for a row tail) bytes at b, the rest zero.  The bytes behind a row tail belong to the following rows: they
// are read and blanked; only the last rows of the buffer go through a bounce buffer so that nothing past `end` is touched.
inline __m256i load_valid(const uint8_t *b, int valid, const uint8_t *end) {
    if (valid >= 32) return _mm256_loadu_si256(reinterpret_cast<const __m256i *>(b));
    if (b + 32 <= end) {
        const __m256i iota = _mm256_setr_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25,
                                              26, 27, 28, 29, 30, 31);
        return _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(b)), _mm256_cmpgt_epi8(_mm256_set1_epi8((char)valid), iota));
    }
    alignas(32) uint8_t tmp[32] = {0};
    memcpy(tmp, b, (size_t)(valid > 0 ? valid : 0));
    return _mm256_load_si256(reinterpret_cast<const __m256i *>(tmp));
}
// Classification by nibbles (two PSHUFB): per byte, bits 0-1 = code and bit 2 set for A,C,G (0x41,0x43,0x47); bits 4-5 = 3
// and bit 6 set for T (0x54); bit 3 set for N (0x4E); 0 for everything else (-> code 0 = A, the reference's zero-initialised
// mapping table, global.c:9-15).
inline __m256i classify(__m256i v) {
    const __m256i lut_lo = _mm256_setr_epi8(0, 0x04, 0, 0x05, 0x70, 0, 0, 0x06, 0, 0, 0, 0, 0, 0, 0x08, 0,
                                            0, 0x04, 0, 0x05, 0x70, 0, 0, 0x06, 0, 0, 0, 0, 0, 0, 0x08, 0);
    const __m256i lut_hi = _mm256_setr_epi8(0, 0, 0, 0, 0x0f, 0x70, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                            0, 0, 0, 0, 0x0f, 0x70, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i nib = _mm256_set1_epi8(0x0f);
    const __m256i lo = _mm256_and_si256(v, nib), hi = _mm256_and_si256(_mm256_srli_epi16(v, 4), nib);
    return _mm256_and_si256(_mm256_shuffle_epi8(lut_lo, lo), _mm256_shuffle_epi8(lut_hi, hi));
}
#ifndef BGSA_PREFETCH_AHEAD
#define BGSA_PREFETCH_AHEAD 2048
#endif
// Bytes ahead of the unit being encoded (the rows of a range are one ascending stream).  The hardware prefetcher of these
// (virtualised) hosts leaves a thread at 4-5 GB/s; with a software prefetch 1.5-4 KB ahead it reaches 6-9 (sweep: 2 KB).
constexpr int kPrefetchAhead = BGSA_PREFETCH_AHEAD;
inline void unit_avx2(const uint8_t *p, int left, bool planes, const uint8_t *end, uint32_t *out, uint32_t *n0, uint32_t *n1) {
    _mm_prefetch(reinterpret_cast<const char *>(p) + kPrefetchAhead, _MM_HINT_T0);
    const __m256i A = load_valid(p, left, end);
    const __m256i B = left > 32 ? load_valid(p + 32, left - 32, end) : _mm256_setzero_si256();
    const __m256i xa = classify(A), xb = classify(B);
    const __m256i three = _mm256_set1_epi8(0x03);
    const __m256i ca = _mm256_and_si256(_mm256_or_si256(xa, _mm256_srli_epi16(xa, 4)), three);
    const __m256i cb = _mm256_and_si256(_mm256_or_si256(xb, _mm256_srli_epi16(xb, 4)), three);
    if (planes) {
        out[0] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(ca, 7));
        out[1] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(ca, 6));
        out[2] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(cb, 7));
        out[3] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(cb, 6));
    } else {
        // 4 codes -> one byte: (c0 + 4 c1) + 16 (c2 + 4 c3) per 32-bit lane, then the 16 bytes of the unit in base order
        const __m256i k1 = _mm256_set1_epi16(0x0401), k2 = _mm256_set1_epi32(0x00100001);
        const __m256i ma = _mm256_madd_epi16(_mm256_maddubs_epi16(ca, k1), k2), mb = _mm256_madd_epi16(_mm256_maddubs_epi16(cb, k1), k2);
        const __m256i w16 = _mm256_packus_epi32(ma, mb);              // per 128-bit lane: 4 x a, 4 x b (16-bit)
        const __m256i w8 = _mm256_packus_epi16(w16, w16);             // per lane: a(4 bytes) b(4 bytes) twice
        const __m256i ord = _mm256_permutevar8x32_epi32(w8, _mm256_setr_epi32(0, 4, 1, 5, 0, 0, 0, 0));   // a.lo a.hi b.lo b.hi
        _mm_storeu_si128(reinterpret_cast<__m128i *>(out), _mm256_castsi256_si128(ord));
    }
    *n0 = 0u; *n1 = 0u;
    const __m256i nbit = _mm256_set1_epi8(0x08);
    if (!_mm256_testz_si256(_mm256_or_si256(xa, xb), nbit)) {         // rare: an 'N' in this unit
        *n0 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(xa, 4));
        *n1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(xb, 4));
    }
}
bool pack_range_avx2(int layout, const uint8_t *rows, int slen, int64_t count, void *packed, int64_t t0, int64_t t1) {
    BGSA_PACK_RANGE_BODY(unit_avx2)
}
#pragma GCC pop_options

// (An AVX-512 VBMI encoder -- one VPERMB lookup pair and three mask tests per 64 bases, PDEP interleave -- was written and
//  measured: no faster than this one, 4.7 against 5.1 GB/s per thread.  A thread is bound by the memory traffic of its
//  stream, about 10 GB/s of reads per core on these hosts, not by the encode; tools/host_pack_bench.py scales linearly
//  with the threads up to the host's memory bandwidth.)
// BGSA_HOST_PACK_ISA=scalar (or the older BGSA_HOST_PACK_SCALAR=1) forces the scalar twin (tests).
bool cpu_has_avx2() {
    static const bool has = [] {
        const char *cap = getenv("BGSA_HOST_PACK_ISA");
        return __builtin_cpu_supports("avx2") && !getenv("BGSA_HOST_PACK_SCALAR") && !(cap && !strcmp(cap, "scalar"));
    }();
    return has;
}

}  // namespace

bool host_pack_tiles(int layout, const uint8_t *rows, int slen, int64_t count, void *packed, int64_t tile_begin, int64_t tile_end) {
    if (cpu_has_avx2()) return pack_range_avx2(layout, rows, slen, count, packed, tile_begin, tile_end);
    return pack_range_scalar(layout, rows, slen, count, packed, tile_begin, tile_end);
}

const char *host_pack_isa() { return cpu_has_avx2() ? "avx2" : "scalar"; }

// ---- worker pool ------------------------------------------------------------------------------------------------------
struct Batch {                       // one parallel_for in flight
    int64_t n;
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> done{0};
    int refs = 0;                    // workers that hold a pointer to this batch (guarded by Impl::mu): the batch lives on
                                     // its caller's stack, so parallel_for must not return while a worker can still touch it
    void (*fn)(int64_t, void *);
    void *arg;
};

struct HostPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<Batch *> open;       // batches that still have indices to hand out
    bool stop = false;

    static bool run_one(Batch *b) {
        const int64_t i = b->next.fetch_add(1);
        if (i >= b->n) return false;
        b->fn(i, b->arg);
        b->done.fetch_add(1);
        return true;
    }
    void worker() {
        std::unique_lock<std::mutex> lk(mu);
        while (true) {
            cv_work.wait(lk, [&] { return stop || !open.empty(); });
            if (stop) return;
            Batch *b = open.front();
            b->refs++;
            lk.unlock();
            while (run_one(b)) {}
            lk.lock();
            b->refs--;
            for (size_t k = 0; k < open.size(); k++)
                if (open[k] == b && b->next.load() >= b->n) { open.erase(open.begin() + (long)k); break; }
            if (b->refs == 0 && b->done.load() >= b->n) cv_done.notify_all();
        }
    }
};

static int pool_default_threads() {
    if (const char *e = getenv("BGSA_HOST_THREADS")) {
        const int t = atoi(e);
        if (t >= 1) return t > 256 ? 256 : t;
    }
    // one process per GPU is the usual deployment (bench.py under torchrun, one aligner per device): leave the other
    // ranks of the box their share of the cores.  BGSA_HOST_GPUS = GPUs sharing this host's cores (default: torchrun's LOCAL_WORLD_SIZE, else 1).
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    int share = 1;
    if (const char *g = getenv("BGSA_HOST_GPUS")) share = atoi(g) >= 1 ? atoi(g) : 1;
    else if (const char *w = getenv("LOCAL_WORLD_SIZE")) share = atoi(w) >= 1 ? atoi(w) : 1;     // torchrun: ranks on this host
    int t = (int)hw / share;
    // never more threads than CPUs this thread may run on (taskset, cgroup cpusets, bgsa_bind_thread_to_device: the workers
    // inherit the creating thread's affinity mask, and two workers per allowed CPU only slow each other down)
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
        const int n = CPU_COUNT(&allowed);
        if (n >= 1 && n < (int)hw) t = std::min(t, std::max(1, n));
    }
    if (t < 1) t = 1;
    return t > 64 ? 64 : t;
}

HostPool::HostPool() : impl_(new Impl), nthreads_(pool_default_threads()) {
    for (int i = 0; i + 1 < nthreads_; i++) impl_->workers.emplace_back([this] { impl_->worker(); });   // the caller is the last worker
}
HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv_work.notify_all();
    for (std::thread &t : impl_->workers) t.join();
    delete impl_;
}
HostPool &HostPool::instance() {
    static HostPool *pool = new HostPool();      // never destroyed: worker threads must not be joined from a static destructor
    return *pool;
}
void HostPool::parallel_for(int64_t n, void (*fn)(int64_t, void *), void *arg) {
    if (n <= 0) return;
    Batch b;
    b.n = n; b.fn = fn; b.arg = arg;
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->open.push_back(&b);
    }
    impl_->cv_work.notify_all();
    while (Impl::run_one(&b)) {}
    std::unique_lock<std::mutex> lk(impl_->mu);
    for (size_t k = 0; k < impl_->open.size(); k++)
        if (impl_->open[k] == &b) { impl_->open.erase(impl_->open.begin() + (long)k); break; }
    impl_->cv_done.wait(lk, [&] { return b.refs == 0 && b.done.load() >= b.n; });
}

}  // namespace bgsa
