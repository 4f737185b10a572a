// api.cu -- the C ABI of libbgsa_b200.so (include/bgsa_b200.h): argument checking, per-device
// contexts (streams + grow-only device buffers), host<->device staging and kernel selection.
// No computation happens on the host besides building the query's match masks (5 x W words).
#include <cuda_runtime.h>
#include <sched.h>

#include <atomic>
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bgsa_b200.h"
#include "banded_host.h"
#include "host_pack.h"
#include "bitpal.cuh"
#include "instances.h"
#include "jit.h"
#include "launch.cuh"
#include "pack.cuh"
#include "query_peq.h"

using namespace bgsa;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(e_ == cudaErrorMemoryAllocation ? BGSA_ERR_NOMEM : BGSA_ERR_CUDA,      \
                        "%s failed: %s", #expr, cudaGetErrorString(e_));                       \
    } while (0)

// ---- instance tables ------------------------------------------------------------------------
struct KL { int K, L; };
#define X(k, l) {k, l},
const KL kMyersTable[] = {BGSA_MYERS_INSTANCES(X)};
const KL kPackedTable[] = {BGSA_BITPAL_PACKED_INSTANCES(X)};
const KL kNonPackedTable[] = {BGSA_BITPAL_NONPACKED_INSTANCES(X)};
#undef X
struct SchemeRow { int id, M, I, G; };
#define X(id, m, i, g) {id, m, i, g},
const SchemeRow kSchemes[] = {BGSA_SCHEMES(X)};
#undef X

// Cheapest instance that fits the query: cost of one DP column of one subject in ALU instructions
// = L lanes x (K words x ops per word + wavefront hand-over).  Ties go to fewer lanes.
// BGSA_FORCE_KL="K,L" pins an instance (A/B measurements; must fit the query).
template <size_t N>
bool pick(const KL (&table)[N], int qlen, int ops_per_word, int handover, KL *out) {
    static const char *force = getenv("BGSA_FORCE_KL");
    int fk = 0, fl = 0;
    if (force && sscanf(force, "%d,%d", &fk, &fl) != 2) fk = fl = 0;
    long best = -1;
    for (size_t i = 0; i < N; i++) {
        const KL c = table[i];
        if (32 * c.K * c.L < qlen) continue;
        if (fk && c.K == fk && c.L == fl) { *out = c; return true; }
        const long cost = (long)c.L * (c.K * ops_per_word + (c.L > 1 ? handover : 0));
        if (best < 0 || cost < best) { best = cost; *out = c; }
    }
    return best >= 0;
}
int find_scheme(int M, int I, int G) {
    for (const SchemeRow &s : kSchemes) if (s.M == M && s.I == I && s.G == G) return s.id;
    return -1;
}

constexpr int kSchemeJit = -2;       // Plan::scheme of a scoring scheme instantiated at run time (jit.h)
struct Plan {
    int algo;
    KL kl;           // transposed kernels
    int scheme;      // BitPAl: id of a built-in scheme, or kSchemeJit
    int M, I, G;     // BitPAl scores as given (0 otherwise)
    int sign;        // Myers
    int e;           // banded
    int layout;      // pack layout the kernel consumes
    int result_size;
};

int make_plan(const bgsa_params_t *p, int qlen, int slen, Plan *plan) {
    if (!p) return fail(BGSA_ERR_ARG, "params is NULL");
    if (qlen <= 0 || slen <= 0) return fail(BGSA_ERR_ARG, "sequence lengths must be positive (query %d, subject %d)", qlen, slen);
    plan->algo = p->algo;
    plan->scheme = -1;
    plan->M = plan->I = plan->G = 0;
    plan->sign = p->myers_sign == 0 ? -1 : p->myers_sign;
    plan->e = p->threshold;
    plan->layout = LAYOUT_CODES;
    plan->result_size = 2;
    plan->kl = KL{0, 0};
    switch (p->algo) {
        case BGSA_MYERS_GLOBAL:
        case BGSA_MYERS_SEMIGLOBAL:
            if (plan->sign != -1 && plan->sign != 1) return fail(BGSA_ERR_ARG, "myers_sign must be -1, 0 or +1");
            if (!pick(kMyersTable, qlen, 10, 20, &plan->kl))
                return fail(BGSA_ERR_UNSUPPORTED, "Myers: query length %d exceeds the largest kernel instance (32768)", qlen);
            return BGSA_OK;
        case BGSA_BITPAL_PACKED:
        case BGSA_BITPAL_PACKED_SEMIGLOBAL:
        case BGSA_BITPAL_NONPACKED: {
            plan->scheme = find_scheme(p->match, p->mismatch, p->gap);
            plan->M = p->match; plan->I = p->mismatch; plan->G = p->gap;
            if (plan->scheme < 0) {
                // not one of the schemes built into the library (make SCHEMES=...): instantiate it at run time, the way the
                // reference runs its generator for a new scheme (Main.java:240-315)
                std::string why;
                const int variant = p->algo == BGSA_BITPAL_NONPACKED ? 0 : (p->algo == BGSA_BITPAL_PACKED ? 1 : 2);
                if (!jit_scheme_ok(variant, p->match, p->mismatch, p->gap, &why))
                    return fail(BGSA_ERR_UNSUPPORTED, "BitPAl: scoring scheme (%d,%d,%d): %s", p->match, p->mismatch, p->gap, why.c_str());
                if (!jit_available(&why))
                    return fail(BGSA_ERR_UNSUPPORTED, "BitPAl: scoring scheme (%d,%d,%d) is not built into the library (csrc/instances.h, "
                                "make SCHEMES=...) and cannot be instantiated at run time: %s", p->match, p->mismatch, p->gap, why.c_str());
                plan->scheme = kSchemeJit;
            }
            const bool ok = p->algo != BGSA_BITPAL_NONPACKED ? pick(kPackedTable, qlen, 77, 34, &plan->kl)
                                                          : pick(kNonPackedTable, qlen, 185, 45, &plan->kl);
            if (!ok) return fail(BGSA_ERR_UNSUPPORTED, "BitPAl: query length %d exceeds the largest kernel instance", qlen);
            return BGSA_OK;
        }
        case BGSA_BANDED_MYERS:
            plan->layout = LAYOUT_PLANES;
            plan->result_size = 1;
            if (p->threshold < 1 || p->threshold > 31)
                return fail(BGSA_ERR_UNSUPPORTED, "banded: threshold %d outside 1..31 (band must fit a 64-bit word)", p->threshold);
            if (qlen != slen)
                return fail(BGSA_ERR_UNSUPPORTED, "banded: query (%d) and subject (%d) lengths must be equal", qlen, slen);
            if (qlen < p->threshold)
                return fail(BGSA_ERR_UNSUPPORTED, "banded: sequences shorter than the threshold");
            return BGSA_OK;
        default:
            return fail(BGSA_ERR_ARG, "unknown algorithm %d", p->algo);
    }
}

// ---- per-device context ----------------------------------------------------------------------
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return BGSA_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(BGSA_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        return BGSA_OK;
    }
};
struct QueryCache {          // device copy of the query-side tables of the last call
    Buf d_tab;               // Peq rows, or BandedRow table
    std::vector<char> key;   // (plan, queries) it was built from
    void *pinned = nullptr;
    size_t pinned_cap = 0;
};
// One in-flight chunk: its own stream and device buffers, so that the H2D copy of chunk i+1
// overlaps the kernels of chunk i and the D2H copy of chunk i-1 (three different engines).
struct PinnedBuf {           // grow-only pinned host staging
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return BGSA_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) return fail(BGSA_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        return BGSA_OK;
    }
};
struct Lane {
    cudaStream_t stream = nullptr;
    Buf d_rows, d_packed, d_results, d_counters;
    PinnedBuf h_packed;                  // host-packed tiles of the chunk in flight (host_pack.h)
    cudaEvent_t staged = nullptr;        // the H2D copy out of h_packed has completed
    bool staged_pending = false;
};
constexpr int kLanesPerJob = 3;
// A "job" = one bgsa_align_batch_submit() call: the subject range is cut into chunks that
// rotate over the job's lanes.  Two jobs (slot 0 / 1) can be in flight per device, mirroring the
// reference's a/b ping-pong buffers (cal_cpu.c:258-267, thread.c:35-170).
// BGSA_TRACE=1: per-chunk device timeline (ms since submit) printed to stderr by bgsa_align_batch_wait --
// the GPU-side counterpart of the reference's read/mem/cal/write timers (cal_cpu.c:459-475).
struct ChunkTrace { int64_t off, n; int lane; cudaEvent_t ev[4]; int host_packed; };     // after H2D, pack, align, D2H
struct Job {
    Lane lane[kLanesPerJob];
    QueryCache qc;
    cudaEvent_t tab_ready = nullptr;
    cudaEvent_t t0 = nullptr;            // trace origin
    // Front-end tuner (pinned subjects): the share of the chunks the host threads pack, searched job by job on the
    // measured throughput of the jobs themselves (submit_impl "front end" comment)
    struct Tuner {
        std::vector<char> key;          // workload shape the state belongs to
        double p = 0.0;                 // current best share of host-packed chunks
        double r_best = 0.0;            // rows bytes/s measured at p
        double step = 0.25;
        int dir = -1, rejected = 0, hold = 0, hold_len = 4;
        bool trial = false;             // the job in flight runs at p_trial, not p
        double p_trial = 0.0;
        double p_used = 0.0, bytes = 0.0;
        bool armed = false;             // the job in flight is being timed
    } tuner;
    cudaEvent_t ev_begin = nullptr, ev_lane_end[kLanesPerJob] = {};
    bool lane_used[kLanesPerJob] = {};
    // link feedback of the hybrid front end: the H2D copies of the last chunks, bracketed by timing events
    struct LinkItem { cudaEvent_t begin = nullptr, end = nullptr; double bytes = 0.0; bool pending = false; };
    static constexpr int kLinkItems = 8;
    LinkItem link[kLinkItems];
    double link_rate = 0.0;              // H2D bytes/s the hybrid front end measured on its last job (0: not yet)
    double last_share = 0.0;             // share of host-packed chunks of the last job (bgsa_batch_front_end)
    std::vector<ChunkTrace> trace;
};
// State of the device-resident entry points (bgsa_align_device / bgsa_align_rows_device), ONE PER CALLER STREAM: query
// tables, work counters and the packed scratch are reused call after call, which is only safe in stream order.  Calls on
// different streams (or from different threads on different streams) get different slots; calls on the same stream are
// serialised by the slot's mutex while they enqueue.
struct Resident {
    bool used = false;
    cudaStream_t stream = nullptr;
    unsigned long long last_use = 0;
    QueryCache qc;
    Buf counters;
    Buf packed;              // bgsa_align_rows_device: packed form of the caller's rows (algorithms without a fused kernel)
    std::mutex mu;
};
constexpr int kResidentSlots = 8;
struct DeviceCtx {
    bool ready = false;
    int sm_count = 0;
    Job job[2];
    Resident resident[kResidentSlots];
    std::mutex resident_mu;
    unsigned long long resident_clock = 0;
    Lane chunk_lane;         // bgsa_align_peq_chunk (its callers are serialised)
    QueryCache chunk_qc;
};
constexpr int kMaxDevices = 64;
DeviceCtx g_ctx[kMaxDevices];
std::mutex g_mu;

int get_ctx(int device, DeviceCtx **out) {
    if (device < 0 || device >= kMaxDevices) return fail(BGSA_ERR_ARG, "bad device ordinal %d", device);
    CUDA_TRY(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceCtx &c = g_ctx[device];
    if (!c.ready) {
        CUDA_TRY(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
        for (Job &j : c.job) {
            for (Lane &l : j.lane) CUDA_TRY(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&j.tab_ready, cudaEventDisableTiming));
        }
        CUDA_TRY(cudaStreamCreateWithFlags(&c.chunk_lane.stream, cudaStreamNonBlocking));
        c.ready = true;
    }
    *out = &c;
    return BGSA_OK;
}

// The resident slot of `stream`, locked (lk).  A stream seen for the first time takes a free slot, or the least recently
// used one after its stream has drained.
int get_resident(DeviceCtx *ctx, cudaStream_t stream, Resident **out, std::unique_lock<std::mutex> *lk) {
    Resident *r = nullptr;
    {
        std::lock_guard<std::mutex> g(ctx->resident_mu);
        Resident *lru = nullptr;
        for (Resident &c : ctx->resident) {
            if (c.used && c.stream == stream) { r = &c; break; }
            if (!c.used) { if (!lru || lru->used) lru = &c; }
            else if (!lru || (lru->used && c.last_use < lru->last_use)) lru = &c;
        }
        if (!r) {
            r = lru;
            if (r->used) {                                   // recycle: nothing of the old stream may still read the buffers
                std::lock_guard<std::mutex> busy(r->mu);
                CUDA_TRY(cudaStreamSynchronize(r->stream));
                r->qc.key.clear();
            }
            r->used = true;
            r->stream = stream;
        }
        r->last_use = ++ctx->resident_clock;
    }
    *lk = std::unique_lock<std::mutex>(r->mu);
    *out = r;
    return BGSA_OK;
}

// Pointers the kernels read with 16-byte vector loads / bulk copies must be aligned (a misaligned address is a sticky
// context error, not a status code): packed tiles 256 B (make_packed_view), results to their element size.
int check_device_pointers(const void *d_packed, const void *d_results, size_t esize) {
    if (d_packed && (reinterpret_cast<uintptr_t>(d_packed) & 255u))
        return fail(BGSA_ERR_ARG, "d_packed must be 256-byte aligned (cudaMalloc alignment)");
    if (d_results && (reinterpret_cast<uintptr_t>(d_results) & (esize - 1)))
        return fail(BGSA_ERR_ARG, "d_results must be aligned to the score size (%zu bytes)", esize);
    return BGSA_OK;
}

// Build (or reuse) the query-side tables on the device.  Returns the device pointer in *d_tab.
int stage_queries(QueryCache &qc, const Plan &plan, const char *queries, int nq, int qlen, int slen, cudaStream_t stream,
                  const void **d_tab) {
    if (nq <= 0 || qlen <= 0) return fail(BGSA_ERR_ARG, "stage_queries: no queries");
    const size_t qbytes = (size_t)nq * (size_t)(qlen + 1);
    std::vector<char> key(sizeof(Plan) + sizeof(int) * 3 + qbytes);
    memcpy(key.data(), &plan, sizeof(Plan));
    const int dims[3] = {nq, qlen, slen};
    memcpy(key.data() + sizeof(Plan), dims, sizeof(dims));
    memcpy(key.data() + sizeof(Plan) + sizeof(dims), queries, qbytes);
    int rc;
    if (key == qc.key && qc.d_tab.p) { *d_tab = qc.d_tab.p; return BGSA_OK; }

    size_t bytes;
    if (plan.algo == BGSA_BANDED_MYERS) bytes = sizeof(BandedRowHost) * (size_t)nq * qlen;
    else bytes = sizeof(uint32_t) * (size_t)nq * kPeqRows * h_peq_row_stride(plan.kl.K, plan.kl.L);
    if (bytes > qc.pinned_cap) {
        if (qc.pinned) { cudaStreamSynchronize(stream); cudaFreeHost(qc.pinned); }
        qc.pinned = nullptr; qc.pinned_cap = 0;
        CUDA_TRY(cudaMallocHost(&qc.pinned, bytes + 256));
        qc.pinned_cap = bytes + 256;
    } else {
        // the previous upload from this pinned buffer must have completed before we overwrite it
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    for (int q = 0; q < nq; q++) {
        const char *row = queries + (size_t)q * (qlen + 1);
        if (plan.algo == BGSA_BANDED_MYERS)
            build_banded_table(row, qlen, slen, plan.e, static_cast<BandedRowHost *>(qc.pinned) + (size_t)q * qlen);
        else
            build_query_peq(row, qlen, plan.kl.K, plan.kl.L,
                            static_cast<uint32_t *>(qc.pinned) + (size_t)q * kPeqRows * h_peq_row_stride(plan.kl.K, plan.kl.L));
    }
    qc.key.clear();                                      // whatever follows, the old tables are gone
    rc = qc.d_tab.ensure(bytes);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(qc.d_tab.p, qc.pinned, bytes, cudaMemcpyHostToDevice, stream));
    qc.key.swap(key);
    *d_tab = qc.d_tab.p;
    return BGSA_OK;
}

// ASCII rows in, scores out, ONE kernel: banded Myers with a warp strip that fits shared memory (banded.cuh FUSED), and the
// thread-per-subject instances of the other algorithms on short rows (rows_kernel.cuh)
bool rows_path_fused(const Plan &plan, int slen) {
    if (plan.algo == BGSA_BANDED_MYERS) return banded_fused_fits(slen);
    return rows_kernel_fits(plan.kl.K, plan.kl.L, slen);
}

int run_align(const Plan &plan, int sm_count, const void *d_tab, unsigned long long *d_counters, int nq, int qlen,
              const void *d_packed, int slen, int64_t count, void *d_results, int64_t result_stride, cudaStream_t stream,
              long long *resident_subjects = nullptr, const void *d_ascii_rows = nullptr) {
    if ((count == 0 || nq == 0) && !resident_subjects) return BGSA_OK;
    LaunchArgs a;
    a.dry_run = resident_subjects != nullptr;
    a.resident_subjects = resident_subjects;
    a.ps = make_packed_view(const_cast<void *>(d_packed), slen, count);
    a.d_peq = static_cast<const uint32_t *>(d_tab);
    a.n_queries = nq;
    a.qlen = qlen;
    a.d_results = d_results;
    a.result_stride = result_stride;
    a.d_counters = d_counters;
    a.sm_count = sm_count;
    a.stream = stream;
    a.d_ascii = plan.algo != BGSA_BANDED_MYERS ? static_cast<const uint8_t *>(d_ascii_rows) : nullptr;   // rows kernel (rows_kernel.cuh)
    cudaError_t e;
    if (plan.scheme == kSchemeJit) {
        const JitSpec spec{plan.algo == BGSA_BITPAL_NONPACKED ? 0 : (plan.algo == BGSA_BITPAL_PACKED ? 1 : 2), plan.M, plan.I, plan.G,
                           plan.kl.K, plan.kl.L};
        std::string err;
        e = launch_bitpal_jit(spec, a, &err);
        if (e != cudaSuccess)
            return fail(err.empty() ? BGSA_ERR_CUDA : BGSA_ERR_UNSUPPORTED, "run-time instance of scheme (%d,%d,%d): %s", plan.M, plan.I,
                        plan.G, err.empty() ? cudaGetErrorString(e) : err.c_str());
        if (!a.dry_run) g_launches.fetch_add(1);
        return BGSA_OK;
    }
    switch (plan.algo) {
        case BGSA_MYERS_GLOBAL:     e = launch_myers(0, plan.kl.K, plan.kl.L, a, plan.sign); break;
        case BGSA_MYERS_SEMIGLOBAL: e = launch_myers(1, plan.kl.K, plan.kl.L, a, plan.sign); break;
        case BGSA_BITPAL_PACKED:    e = launch_bitpal_packed(plan.scheme, plan.kl.K, plan.kl.L, a); break;
        case BGSA_BITPAL_NONPACKED: e = launch_bitpal_nonpacked(plan.scheme, plan.kl.K, plan.kl.L, a); break;
        case BGSA_BITPAL_PACKED_SEMIGLOBAL: e = launch_bitpal_semiglobal(plan.scheme, plan.kl.K, plan.kl.L, a); break;
        case BGSA_BANDED_MYERS:     e = d_ascii_rows ? launch_banded_fused(a, d_ascii_rows, d_tab, plan.e) : launch_banded(a, d_tab, plan.e); break;
        default: return fail(BGSA_ERR_ARG, "unknown algorithm %d", plan.algo);
    }
    if (e != cudaSuccess) return fail(BGSA_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    if (!a.dry_run) g_launches.fetch_add(1);
    return BGSA_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

const char *bgsa_version(void) { return "bgsa_b200 0.1 (sm_100a)"; }
const char *bgsa_last_error(void) { return g_err; }

int bgsa_device_count(int *count) {
    if (!count) return fail(BGSA_ERR_ARG, "count is NULL");
    CUDA_TRY(cudaGetDeviceCount(count));
    return BGSA_OK;
}

int bgsa_init_devices(int n_devices) {
    int have = 0;
    CUDA_TRY(cudaGetDeviceCount(&have));
    if (n_devices < 1 || n_devices > have || n_devices > kMaxDevices)
        return fail(BGSA_ERR_ARG, "bgsa_init_devices: %d devices requested, %d present", n_devices, have);
    std::vector<int> rc((size_t)n_devices, BGSA_OK);
    std::vector<std::string> msg((size_t)n_devices);
    std::vector<std::thread> th;
    for (int g = 0; g < n_devices; g++)
        th.emplace_back([g, &rc, &msg] {
            DeviceCtx *ctx;
            if (cudaSetDevice(g) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) {   // creates the primary context
                rc[(size_t)g] = BGSA_ERR_CUDA;
                msg[(size_t)g] = cudaGetErrorString(cudaGetLastError());
                return;
            }
            rc[(size_t)g] = get_ctx(g, &ctx);                  // streams, events (short, under the registry lock)
            if (rc[(size_t)g] != BGSA_OK) msg[(size_t)g] = g_err;   // g_err is thread-local
        });
    for (std::thread &t : th) t.join();
    for (int g = 0; g < n_devices; g++)
        if (rc[(size_t)g] != BGSA_OK) return fail(rc[(size_t)g], "device %d: %s", g, msg[(size_t)g].c_str());
    return BGSA_OK;
}

void bgsa_params_default(bgsa_params_t *p, int algo) {
    if (!p) return;
    p->algo = algo;
    p->match = 2; p->mismatch = -3; p->gap = -5;   // original/BGSA_AVX512/align_core.c:13-15
    if (algo == BGSA_MYERS_GLOBAL || algo == BGSA_MYERS_SEMIGLOBAL || algo == BGSA_BANDED_MYERS) {
        p->match = 0; p->mismatch = -1; p->gap = -1;   // Main.java:253-257
    }
    p->threshold = 31;                              // banded/BGSA_CPU/main.c:43
    p->myers_sign = -1;                             // generator -m 0
}

int bgsa_result_size(int algo) { return algo == BGSA_BANDED_MYERS ? 1 : 2; }

int bgsa_supported(const bgsa_params_t *p, int query_len, int subject_len) {
    Plan plan;
    return make_plan(p, query_len, subject_len, &plan);
}

int bgsa_kernel_name(const bgsa_params_t *p, int query_len, int subject_len, char *buf, int buflen) {
    Plan plan;
    int rc = make_plan(p, query_len, subject_len, &plan);
    if (rc) return rc;
    if (!buf || buflen <= 0) return fail(BGSA_ERR_ARG, "buf is NULL");
    switch (plan.algo) {
        case BGSA_MYERS_GLOBAL: snprintf(buf, buflen, "align_kernel<MyersAlgo<K=%d,global>,L=%d>", plan.kl.K, plan.kl.L); break;
        case BGSA_MYERS_SEMIGLOBAL: snprintf(buf, buflen, "align_kernel<MyersAlgo<K=%d,semiglobal>,L=%d>", plan.kl.K, plan.kl.L); break;
        case BGSA_BITPAL_PACKED: snprintf(buf, buflen, "align_kernel<BitpalPacked<%d,%d,%d,K=%d>,L=%d>", p->match, p->mismatch, p->gap, plan.kl.K, plan.kl.L); break;
        case BGSA_BITPAL_PACKED_SEMIGLOBAL: snprintf(buf, buflen, "align_kernel<BitpalPacked<%d,%d,%d,K=%d,semiglobal>,L=%d>", p->match, p->mismatch, p->gap, plan.kl.K, plan.kl.L); break;
        case BGSA_BITPAL_NONPACKED: snprintf(buf, buflen, "align_kernel<BitpalNonPacked<%d,%d,%d,K=%d>,L=%d>", p->match, p->mismatch, p->gap, plan.kl.K, plan.kl.L); break;
        default: snprintf(buf, buflen, "banded_kernel<%s>", 2 * plan.e + 2 <= 32 ? "u32" : "u64"); break;
    }
    if (plan.scheme == kSchemeJit && strlen(buf) + 9 < (size_t)buflen) strcat(buf, " [NVRTC]");
    return BGSA_OK;
}

int bgsa_rows_kernel_name(const bgsa_params_t *p, int query_len, int subject_len, char *buf, int buflen, int *fused) {
    Plan plan;
    int rc = make_plan(p, query_len, subject_len, &plan);
    if (rc) return rc;
    if (!buf || buflen <= 0) return fail(BGSA_ERR_ARG, "buf is NULL");
    char packed[200];
    if ((rc = bgsa_kernel_name(p, query_len, subject_len, packed, (int)sizeof(packed)))) return rc;
    const bool one = rows_path_fused(plan, subject_len);
    if (fused) *fused = one ? 1 : 0;
    if (!one) { snprintf(buf, buflen, "pack_stream_kernel + %s", packed); return BGSA_OK; }
    if (plan.algo == BGSA_BANDED_MYERS) { snprintf(buf, buflen, "%s (fused: ASCII tile -> shared-memory strip -> band)", packed); return BGSA_OK; }
    std::string name(packed);                                  // align_kernel<Algo,L=1> -> align_rows_kernel<Algo>
    const size_t at = name.find("align_kernel<"), l = name.rfind(",L=1>");
    if (at != std::string::npos && l != std::string::npos) name = "align_rows_kernel<" + name.substr(at + 13, l - at - 13) + ">";
    snprintf(buf, buflen, "%s", name.c_str());
    return BGSA_OK;
}

int bgsa_jit_precompile(const bgsa_params_t *p, int query_len, int subject_len) {
    Plan plan;
    int rc = make_plan(p, query_len, subject_len, &plan);
    if (rc) return rc;
    if (plan.scheme != kSchemeJit) return BGSA_OK;            // built in: nothing to do
    const JitSpec spec{plan.algo == BGSA_BITPAL_NONPACKED ? 0 : (plan.algo == BGSA_BITPAL_PACKED ? 1 : 2), plan.M, plan.I, plan.G,
                       plan.kl.K, plan.kl.L};
    std::string err;
    if (jit_precompile(spec, &err)) return fail(BGSA_ERR_UNSUPPORTED, "%s", err.c_str());
    return BGSA_OK;
}

int64_t bgsa_launch_count(void) { return g_launches.load(); }

void *bgsa_malloc_host(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { fail(BGSA_ERR_NOMEM, "cudaMallocHost(%zu) failed", bytes); return nullptr; }
    return p;
}
void bgsa_free_host(void *p) { if (p) cudaFreeHost(p); }
int bgsa_host_register(void *p, size_t bytes) {
    if (!p || bytes == 0) return fail(BGSA_ERR_ARG, "bgsa_host_register: empty buffer");
    CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return BGSA_OK;
}
int bgsa_bind_thread_to_device(int device, int *numa_node) {
    if (numa_node) *numa_node = -1;
    char bus[32] = "";
    CUDA_TRY(cudaDeviceGetPCIBusId(bus, sizeof(bus), device));
    for (char *c = bus; *c; c++) *c = (char)tolower((unsigned char)*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return BGSA_OK;                       // no NUMA information: leave the thread alone
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return BGSA_OK;
    cpu_set_t set;
    CPU_ZERO(&set);
    int a, b, n = 0;
    while (fscanf(f, "%d", &a) == 1) {                   // "0-31,64-95"
        b = a;
        int ch = fgetc(f);
        if (ch == '-') { if (fscanf(f, "%d", &b) != 1) b = a; ch = fgetc(f); }
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); n++; }
        if (ch != ',') break;
    }
    fclose(f);
    if (n == 0) return BGSA_OK;
    // keep only CPUs this process is allowed to use (containers); give up quietly if none is left
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
        CPU_AND(&set, &set, &allowed);
        if (CPU_COUNT(&set) == 0) return BGSA_OK;
    }
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return fail(BGSA_ERR_ARG, "sched_setaffinity failed for NUMA node %d", node);
    if (numa_node) *numa_node = node;
    return BGSA_OK;
}
int bgsa_host_unregister(void *p) {
    if (!p) return fail(BGSA_ERR_ARG, "bgsa_host_unregister: NULL");
    CUDA_TRY(cudaHostUnregister(p));
    return BGSA_OK;
}

int64_t bgsa_packed_bytes(int subject_len, int64_t count) {
    if (subject_len <= 0 || count < 0) return -1;
    return packed_bytes(subject_len, count);
}

int bgsa_pack_subjects_device(const bgsa_params_t *p, const void *d_rows, int subject_len, int64_t count, void *d_packed,
                              int device, void *stream) {
    if (!p || (!d_rows && count > 0) || (!d_packed && count > 0) || subject_len <= 0 || count < 0)
        return fail(BGSA_ERR_ARG, "bgsa_pack_subjects_device: bad argument");
    DeviceCtx *ctx;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    if (count == 0) return BGSA_OK;
    if ((rc = check_device_pointers(d_packed, nullptr, 1))) return rc;
    const int layout = p->algo == BGSA_BANDED_MYERS ? LAYOUT_PLANES : LAYOUT_CODES;
    cudaError_t e = launch_pack(layout, d_rows, subject_len, count, d_packed, ctx->sm_count, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(BGSA_ERR_CUDA, "pack kernel launch failed: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return BGSA_OK;
}

static bool host_pack_chunk(int layout, const uint8_t *rows, int slen, int64_t n, void *h_packed);

int bgsa_pack_subjects_host(const bgsa_params_t *p, const void *rows, int subject_len, int64_t count, void *packed) {
    if (!p || (!rows && count > 0) || (!packed && count > 0) || subject_len <= 0 || count < 0)
        return fail(BGSA_ERR_ARG, "bgsa_pack_subjects_host: bad argument");
    if (count == 0) return BGSA_OK;
    const int layout = p->algo == BGSA_BANDED_MYERS ? LAYOUT_PLANES : LAYOUT_CODES;
    // (the N plane of tiles without an N is left untouched, exactly as the device pack kernels leave it)
    host_pack_chunk(layout, static_cast<const uint8_t *>(rows), subject_len, count, packed);
    return BGSA_OK;
}

int bgsa_host_pack_info(int *threads, char *isa, int isa_len) {
    if (threads) *threads = HostPool::instance().threads();
    if (isa && isa_len > 0) snprintf(isa, (size_t)isa_len, "%s", host_pack_isa());
    return BGSA_OK;
}

int bgsa_align_device(const bgsa_params_t *p, const char *h_queries, int n_queries, int query_len, const void *d_packed,
                      int subject_len, int64_t count, void *d_results, int64_t result_stride, int device, void *stream) {
    Plan plan;
    int rc = make_plan(p, query_len, subject_len, &plan);
    if (rc) return rc;
    if (!h_queries || n_queries < 0 || count < 0 || (count > 0 && (!d_packed || !d_results)) || result_stride < count)
        return fail(BGSA_ERR_ARG, "bgsa_align_device: bad argument");
    DeviceCtx *ctx;
    rc = get_ctx(device, &ctx);
    if (rc) return rc;
    if (count == 0 || n_queries == 0) return BGSA_OK;
    if ((rc = check_device_pointers(d_packed, d_results, (size_t)plan.result_size))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Resident *res;
    std::unique_lock<std::mutex> lk;
    if ((rc = get_resident(ctx, st, &res, &lk))) return rc;
    const void *d_tab;
    rc = stage_queries(res->qc, plan, h_queries, n_queries, query_len, subject_len, st, &d_tab);
    if (rc) return rc;
    if ((rc = res->counters.ensure(sizeof(unsigned long long) * (size_t)n_queries))) return rc;
    return run_align(plan, ctx->sm_count, d_tab, static_cast<unsigned long long *>(res->counters.p), n_queries,
                     query_len, d_packed, subject_len, count, d_results, result_stride, st);
}

int bgsa_align_rows_device(const bgsa_params_t *p, const char *h_queries, int n_queries, int query_len, const void *d_rows,
                           int subject_len, int64_t count, void *d_results, int64_t result_stride, int device, void *stream) {
    Plan plan;
    int rc = make_plan(p, query_len, subject_len, &plan);
    if (rc) return rc;
    if (!h_queries || n_queries < 0 || count < 0 || (count > 0 && (!d_rows || !d_results)) || result_stride < count)
        return fail(BGSA_ERR_ARG, "bgsa_align_rows_device: bad argument");
    DeviceCtx *ctx;
    rc = get_ctx(device, &ctx);
    if (rc) return rc;
    if (count == 0 || n_queries == 0) return BGSA_OK;
    if ((rc = check_device_pointers(nullptr, d_results, (size_t)plan.result_size))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Resident *res;
    std::unique_lock<std::mutex> lk;
    if ((rc = get_resident(ctx, st, &res, &lk))) return rc;
    const void *d_tab;
    rc = stage_queries(res->qc, plan, h_queries, n_queries, query_len, subject_len, st, &d_tab);
    if (rc) return rc;
    if ((rc = res->counters.ensure(sizeof(unsigned long long) * (size_t)n_queries))) return rc;
    unsigned long long *d_counters = static_cast<unsigned long long *>(res->counters.p);
    if (rows_path_fused(plan, subject_len))                                  // ASCII in, scores out, one kernel
        return run_align(plan, ctx->sm_count, d_tab, d_counters, n_queries, query_len, nullptr, subject_len, count, d_results,
                         result_stride, st, nullptr, d_rows);
    if (packed_bytes(subject_len, count) > (int64_t)res->packed.cap) CUDA_TRY(cudaStreamSynchronize(st));   // the old scratch may still be read
    if ((rc = res->packed.ensure((size_t)packed_bytes(subject_len, count)))) return rc;
    cudaError_t e = launch_pack(plan.layout, d_rows, subject_len, count, res->packed.p, ctx->sm_count, st);
    if (e != cudaSuccess) return fail(BGSA_ERR_CUDA, "pack kernel launch failed: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return run_align(plan, ctx->sm_count, d_tab, d_counters, n_queries, query_len, res->packed.p, subject_len, count,
                     d_results, result_stride, st);
}

// ---- optional host-side pack (host_pack.h) -----------------------------------------------------------------------------
// Per job: NEVER (the kernel, not the link, bounds the batch: long sequences, expensive scoring), ALWAYS (pageable subject
// memory, where the ASCII copy crawls at ~10 GB/s through the driver's staging buffers; or forced), or HYBRID: the PCIe
// link and the host threads work side by side -- a chunk goes over the link as ASCII (device pack) whenever the link would
// otherwise run dry while the threads encode, else the threads encode it and a quarter of the bytes is shipped.
// BGSA_HOST_PACK=0 / 1 forces never / always; =2 forces hybrid.
enum HostPackMode { HP_NEVER = 0, HP_ALWAYS = 1, HP_HYBRID = 2 };
constexpr double kPcieBytesPerS = 52e9;     // measured H2D rate of pinned rows on this platform (55 GB/s peak)
// *tunable: pinned subjects, nothing forced, and a batch whose input path matters -- the model's answer is then only the
// starting point of the per-job search (Job::Tuner).
static HostPackMode decide_host_pack(const Plan &plan, int nq, int qlen, int slen, int64_t count, const void *rows, bool *tunable) {
    *tunable = false;
    if (const char *env = getenv("BGSA_HOST_PACK")) {
        if (env[0] == '0') return HP_NEVER;
        if (env[0] == '1') return HP_ALWAYS;
        if (env[0] == '2') return HP_HYBRID;
    }
    if (count < 8192) return HP_NEVER;
    const double bytes = (double)slen + 1.0;
    const double words = (double)plan.kl.K * (plan.kl.L > 0 ? plan.kl.L : 1);
    double instr;                                        // ALU lane-instructions per (query, subject)
    switch (plan.algo) {
        case BGSA_MYERS_GLOBAL: case BGSA_MYERS_SEMIGLOBAL: instr = (double)slen * words * 10.2; break;
        case BGSA_BITPAL_PACKED: case BGSA_BITPAL_PACKED_SEMIGLOBAL: instr = (double)slen * words * 67.0; break;
        case BGSA_BITPAL_NONPACKED: instr = (double)slen * words * 165.0; break;
        default: instr = (double)qlen * 13.0 * 0.7 + 250.0; break;      // banded: ~2/3 of the rows on average
    }
    const double t_kernel = nq * instr / 18.0e12;
    cudaPointerAttributes attr;
    bool pinned = true;
    if (cudaPointerGetAttributes(&attr, rows) == cudaSuccess) pinned = attr.type != cudaMemoryTypeUnregistered;
    else cudaGetLastError();
    const double t_pack = bytes / (HostPool::instance().threads() * 6.0e9);     // measured: 6-8 GB/s per thread (bound by the thread's memory stream)
    const double t_link = bytes / kPcieBytesPerS;
    if (!pinned) return std::max(t_pack, t_kernel) < 0.9 * std::max(bytes / 9e9, t_kernel) ? HP_ALWAYS : HP_NEVER;
    if (t_kernel > 2.0 * t_link) return HP_NEVER;                                 // the link hides behind the kernel with room to spare: leave the host
                                                                                  // cores alone and keep the submit call asynchronous (C5: 375 ms of kernel per 12 ms of copy)
    *tunable = getenv("BGSA_HOST_PACK_NO_TUNING") == nullptr;
    if (t_pack <= 0.9 * t_kernel) return HP_ALWAYS;                               // the threads stay ahead of the kernel: the link is nearly free
    if (t_kernel > 1.15 * t_link) return HP_NEVER;                                // they cannot, and the link hides behind the kernel anyway
    // The pool must at least come near the link's rate (6 threads and more).  With fewer -- 8 ranks on a 32-core host -- the
    // DMA engines of all the GPUs together already saturate the host's memory (182 GB/s) and every byte the threads touch
    // only adds traffic: any share of host packing LOSES there, tuned or not (profiles/r02_e2e_multi_rank.log).
    if (t_pack >= 1.5 * t_link) { *tunable = false; return HP_NEVER; }
    return HP_HYBRID;
}
// Ranks (processes or devices of this process) that pull subjects through this host at the same time: torchrun's
// LOCAL_WORLD_SIZE / BGSA_HOST_GPUS, or the device contexts this process has brought up.
static int ranks_sharing_host() {
    int n = 1;
    if (const char *g = getenv("BGSA_HOST_GPUS")) n = atoi(g);
    else if (const char *w = getenv("LOCAL_WORLD_SIZE")) n = atoi(w);
    int ready = 0;
    for (int d = 0; d < kMaxDevices; d++) ready += g_ctx[d].ready ? 1 : 0;
    return std::max(std::max(n, ready), 1);
}
static double host_now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct HostPackTask { int layout; const uint8_t *rows; int slen; int64_t count; void *packed; int64_t ntiles, grain; std::atomic<int> any_n; };
static void host_pack_block(int64_t i, void *arg) {
    HostPackTask *t = static_cast<HostPackTask *>(arg);
    const int64_t t0 = i * t->grain, t1 = std::min(t->ntiles, t0 + t->grain);
    if (host_pack_tiles(t->layout, t->rows, t->slen, t->count, t->packed, t0, t1)) t->any_n.store(1);
}
// packs `n` rows into `h_packed` (tile layout) with the pool; returns whether any tile holds an N
static bool host_pack_chunk(int layout, const uint8_t *rows, int slen, int64_t n, void *h_packed) {
    HostPackTask t;
    t.layout = layout; t.rows = rows; t.slen = slen; t.count = n; t.packed = h_packed;
    t.ntiles = (n + kTileSubjects - 1) / kTileSubjects;
    t.grain = std::max<int64_t>(1, (96 << 10) / (kTileSubjects * ((int64_t)slen + 1)));
    t.any_n.store(0);
    HostPool::instance().parallel_for((t.ntiles + t.grain - 1) / t.grain, host_pack_block, &t);
    return t.any_n.load() != 0;
}

static int submit_impl(const bgsa_params_t *p, const char *queries, int n_queries, int query_len,
                       const bgsa_seq_t *subjects, int64_t first, int64_t count, void *results, int64_t result_stride,
                       int device, int slot) {
    if (!subjects) return fail(BGSA_ERR_ARG, "subjects is NULL");
    Plan plan;
    int rc = make_plan(p, query_len, subjects->len, &plan);
    if (rc) return rc;
    if (!queries || n_queries < 0 || first < 0 || count < 0 || first + count > subjects->count || slot < 0 || slot > 1 ||
        (count > 0 && n_queries > 0 && (!results || !subjects->content)) || result_stride < count)
        return fail(BGSA_ERR_ARG, "bgsa_align_batch: bad argument (first %lld count %lld of %lld, stride %lld)",
                    (long long)first, (long long)count, (long long)subjects->count, (long long)result_stride);
    DeviceCtx *ctx;
    rc = get_ctx(device, &ctx);
    if (rc) return rc;
    if (count == 0 || n_queries == 0) return BGSA_OK;
    Job &job = ctx->job[slot];
    const int slen = subjects->len;
    const size_t esize = plan.result_size;
    // query-side tables once per job, on lane 0's stream; the other lanes wait on the event
    const void *d_tab;
    rc = stage_queries(job.qc, plan, queries, n_queries, query_len, slen, job.lane[0].stream, &d_tab);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(job.tab_ready, job.lane[0].stream));
    // Chunking.  The persistent grid works on `quantum` subjects at once (resident warps x subjects per warp) and
    // every warp starts on its own work unit, so a chunk of a whole number of quanta keeps all warps busy for the
    // same number of rounds.  Small chunks let the kernels follow the H2D stream closely (only the first chunk's
    // copy and the last chunk's kernel are exposed): aim at ~16 chunks, never below one quantum or 4 MB of rows.
    // banded Myers on short rows: one fused kernel per chunk (ASCII tile -> shared-memory strip -> band), no pack launch
    // Front end.  The static model (decide_host_pack) knows the nominal link and pack rates of ONE rank on an idle host; with
    // several ranks on a host the link delivers a fraction of that and the pack threads compete with the DMA engines for
    // the host's memory bandwidth (8 GPUs on a 32-core host: any host packing LOSES 15 %; 4 GPUs behind one PCIe switch:
    // it GAINS 20 %; profiles/r02_e2e_multi_rank.log).  With several ranks and pinned subjects the model therefore only seeds a search: every
    // job is timed on the device, and the share p of host-packed chunks moves by +-step whenever a trial job beats the
    // incumbent by 3 % (or ties it with less host work); two failed trials halve the step and double the pause before
    // the next probe.  The state belongs to the (device, slot) and to the workload shape.
    bool tunable = false;
    HostPackMode hp_mode = decide_host_pack(plan, n_queries, query_len, slen, count, subjects->content + (size_t)first * (slen + 1), &tunable);
    if (tunable && ranks_sharing_host() < 2 && !getenv("BGSA_HOST_PACK_TUNING")) tunable = false;   // alone on the host: the model holds
    double share = hp_mode == HP_ALWAYS ? 1.0 : (hp_mode == HP_NEVER ? 0.0 : -1.0);    // -1: live model of the measured rates (BGSA_HOST_PACK=2)
    Job::Tuner &tn = job.tuner;
    tn.armed = false;
    if (tunable) {
        const long long shape[6] = {plan.algo, plan.kl.K, plan.kl.L, slen, n_queries, (long long)(count >> 12)};
        std::vector<char> key(sizeof(shape));
        memcpy(key.data(), shape, sizeof(shape));
        if (key != tn.key) {
            tn = Job::Tuner();
            tn.key = key;
            if (hp_mode == HP_HYBRID) {
                const double tp = 1.0 / (HostPool::instance().threads() * 6.0e9), tl = 1.0 / kPcieBytesPerS;
                tn.p = 1.0 - (tp - 0.25 * tl) / (tp + 0.75 * tl);
            } else {
                tn.p = share;
            }
            share = tn.p;
        } else {
            auto clamp01 = [](double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); };
            tn.trial = false;                           // (a job whose wait never evaluated it leaves no stale trial behind)
            if (tn.hold > 0) {
                tn.hold--;
                share = tn.p;
            } else {
                double c = clamp01(tn.p + tn.dir * tn.step);
                if (c == tn.p) { tn.dir = -tn.dir; c = clamp01(tn.p + tn.dir * tn.step); }
                tn.trial = c != tn.p;
                tn.p_trial = c;
                share = c;
            }
        }
        tn.p_used = share;
        tn.bytes = (double)count * (slen + 1);
        tn.armed = true;
        hp_mode = share >= 1.0 ? HP_ALWAYS : (share <= 0.0 ? HP_NEVER : HP_HYBRID);
    }
    const bool live_model = hp_mode == HP_HYBRID && share < 0.0;
    const bool can_fuse = rows_path_fused(plan, slen);                                     // chunks that arrive as ASCII
    long long quantum = 0;
    rc = run_align(plan, ctx->sm_count, d_tab, nullptr, n_queries, query_len, nullptr, slen, 0, nullptr, 0, nullptr, &quantum,
                   (can_fuse && hp_mode != HP_ALWAYS) ? static_cast<const void *>(&quantum) : nullptr);   // (dry run: the pointer only selects the kernel)
    if (rc) return rc;
    if (quantum < kTileSubjects) quantum = kTileSubjects;
    static const int kChunks = getenv("BGSA_CHUNKS") ? atoi(getenv("BGSA_CHUNKS")) : 16;   // tuning knob
    static const bool kTrace = getenv("BGSA_TRACE") != nullptr;
    const int64_t min_subjects = ((4 << 20) + slen) / (slen + 1);
    int64_t min_rounds = (min_subjects + quantum - 1) / quantum;
    if (min_rounds < 1) min_rounds = 1;
    int64_t rounds = (count / (kChunks > 0 ? kChunks : 1) + quantum / 2) / quantum;
    if (rounds < min_rounds) rounds = min_rounds;
    const int64_t chunk = (rounds * quantum + kTileSubjects - 1) / kTileSubjects * kTileSubjects;
    int64_t first_chunk = (min_rounds * quantum + kTileSubjects - 1) / kTileSubjects * kTileSubjects;
    if (count < 2 * chunk) first_chunk = chunk;
    if (kTrace) {
        if (!job.t0) CUDA_TRY(cudaEventCreate(&job.t0));
        for (ChunkTrace &c : job.trace) for (cudaEvent_t e : c.ev) cudaEventDestroy(e);
        job.trace.clear();
        CUDA_TRY(cudaEventRecord(job.t0, job.lane[0].stream));
    }
    int li = 0;
    if (tn.armed) {
        if (!job.ev_begin) {
            CUDA_TRY(cudaEventCreate(&job.ev_begin));
            for (cudaEvent_t &e : job.ev_lane_end) CUDA_TRY(cudaEventCreate(&e));
        }
        for (bool &u : job.lane_used) u = false;
        CUDA_TRY(cudaEventRecord(job.ev_begin, job.lane[0].stream));
    }
    // hybrid bookkeeping (host clock): when the link will have drained what has been queued on it, and the threads' measured rate
    // hybrid bookkeeping, all MEASURED: the threads' pack rate, and the link's state read back from the lanes' copy events --
    // bytes still queued on the link and the rate at which the finished copies really moved (with several ranks on one
    // host the link delivers a fraction of its nominal rate: 184 GB/s for 8 GPUs together against 55 for one alone)
    double pack_rate = HostPool::instance().threads() * 6.0e9, link_rate = job.link_rate > 0 ? job.link_rate : kPcieBytesPerS;
    // A copy's own duration = from the later of (its begin event, the end of the copy queued before it) to its end: the
    // lanes' copies share one engine, so a copy queued behind another would otherwise look slow.
    auto link_backlog_s = [&]() {
        double queued = 0.0;
        for (int i = 0; i < Job::kLinkItems; i++) {
            Job::LinkItem &x = job.link[i];
            if (!x.pending) continue;
            if (cudaEventQuery(x.end) == cudaSuccess) {
                float own = 0.f, after_prev = 0.f;
                const Job::LinkItem &prev = job.link[(i + Job::kLinkItems - 1) % Job::kLinkItems];
                if (x.bytes >= (1 << 20) && cudaEventElapsedTime(&own, x.begin, x.end) == cudaSuccess && own > 0.f) {
                    if (prev.end && !prev.pending && cudaEventElapsedTime(&after_prev, prev.end, x.end) == cudaSuccess && after_prev > 0.f &&
                        after_prev < own)
                        own = after_prev;
                    link_rate = 0.5 * link_rate + 0.5 * x.bytes / (1e-3 * own);
                }
                cudaGetLastError();
                x.pending = false;
            } else {
                cudaGetLastError();                       // cudaErrorNotReady is not an error
                queued += x.bytes;
            }
        }
        return queued / link_rate;
    };
    int link_slot = 0;
    double ascii_credit = 1.0 - 1e-9;
    auto link_begin = [&](cudaStream_t st) -> int {
        Job::LinkItem &x = job.link[link_slot];
        if (!x.begin && (cudaEventCreate(&x.begin) != cudaSuccess || cudaEventCreate(&x.end) != cudaSuccess)) return BGSA_ERR_CUDA;
        x.pending = false;
        return cudaEventRecord(x.begin, st) == cudaSuccess ? BGSA_OK : BGSA_ERR_CUDA;
    };
    auto link_end = [&](cudaStream_t st, double bytes) -> int {
        Job::LinkItem &x = job.link[link_slot];
        link_slot = (link_slot + 1) % Job::kLinkItems;
        x.bytes = bytes; x.pending = true;
        return cudaEventRecord(x.end, st) == cudaSuccess ? BGSA_OK : BGSA_ERR_CUDA;
    };
    for (int64_t off = 0, step = first_chunk; off < count; off += step, step = chunk, li = (li + 1) % kLanesPerJob) {
        const int64_t n = count - off < step ? count - off : step;
        Lane &l = job.lane[li];
        bool host_pack = hp_mode == HP_ALWAYS;
        if (hp_mode == HP_HYBRID) {
            if (live_model) {
                // One rank on the host: decide chunk by chunk from the link's measured state -- pack while the copies
                // already queued keep the link busy for the time the threads need for this chunk, else feed the link
                // (the first chunk goes over the link: the threads start on the second at once).  Factor swept on the
                // B200 box (profiles/r02_e2e_hybrid_factor.log): 1.0 beats 0.5 / 0.25 / 0.1 on C3 and Myers 150 bp.
                static const double kF = getenv("BGSA_HYBRID_F") ? atof(getenv("BGSA_HYBRID_F")) : 1.0;      // A/B knob
                host_pack = link_backlog_s() > kF * (double)n * (slen + 1) / pack_rate;
            } else {
                // Tuned share (several ranks on the host): 1 - share of the chunks cross the link as ASCII, dealt out by
                // error diffusion.
                ascii_credit += 1.0 - share;
                host_pack = ascii_credit < 1.0;
                if (!host_pack) ascii_credit -= 1.0;
            }
        }
        const bool fused = can_fuse && !host_pack;
        ChunkTrace tr{off, n, li, {nullptr, nullptr, nullptr, nullptr}, host_pack ? 1 : 0};
        auto mark = [&](int i) {
            if (kTrace && cudaEventCreate(&tr.ev[i]) == cudaSuccess) cudaEventRecord(tr.ev[i], l.stream);
        };
        const size_t row_bytes = (size_t)n * (slen + 1);
        if (!host_pack && (rc = l.d_rows.ensure(row_bytes + 16))) return rc;
        if (!fused && (rc = l.d_packed.ensure((size_t)packed_bytes(slen, n)))) return rc;
        if ((rc = l.d_results.ensure(esize * (size_t)n_queries * (size_t)n))) return rc;
        if ((rc = l.d_counters.ensure(sizeof(unsigned long long) * (size_t)n_queries))) return rc;
        if (host_pack) {
            // host threads encode the chunk into pinned staging (the kernels of the previous chunks run meanwhile), then a
            // quarter of the bytes crosses the link: the codes, the per-tile N flags and -- only if some tile has an N -- the N plane
            const size_t pbytes = (size_t)packed_bytes(slen, n);
            if (l.staged_pending) { CUDA_TRY(cudaEventSynchronize(l.staged)); l.staged_pending = false; }
            if ((rc = l.h_packed.ensure(pbytes))) return rc;
            const double t_pack0 = host_now_s();
            const bool any_n = host_pack_chunk(plan.layout, reinterpret_cast<const uint8_t *>(subjects->content) + (size_t)(first + off) * (slen + 1),
                                               slen, n, l.h_packed.p);
            const double t_pack1 = host_now_s();
            if (t_pack1 > t_pack0) pack_rate = 0.5 * pack_rate + 0.5 * (double)row_bytes / (t_pack1 - t_pack0);
            const PackedSubjects hv = make_packed_view(l.h_packed.p, slen, n);
            const char *hb = static_cast<const char *>(l.h_packed.p);
            char *db = static_cast<char *>(l.d_packed.p);
            const size_t codes_bytes = (size_t)hv.ntiles * hv.ku * 32 * sizeof(uint4);
            const size_t nm_off = (size_t)(reinterpret_cast<const char *>(hv.nmask) - hb), fl_off = (size_t)(reinterpret_cast<const char *>(hv.tile_has_n) - hb);
            if (live_model && link_begin(l.stream)) return fail(BGSA_ERR_CUDA, "event record failed");
            CUDA_TRY(cudaMemcpyAsync(db, hb, codes_bytes, cudaMemcpyHostToDevice, l.stream));
            CUDA_TRY(cudaMemcpyAsync(db + fl_off, hb + fl_off, (size_t)hv.ntiles, cudaMemcpyHostToDevice, l.stream));
            if (any_n) CUDA_TRY(cudaMemcpyAsync(db + nm_off, hb + nm_off, (size_t)hv.ntiles * hv.kn * 32 * sizeof(uint32_t), cudaMemcpyHostToDevice, l.stream));
            if (!l.staged) CUDA_TRY(cudaEventCreateWithFlags(&l.staged, cudaEventDisableTiming));
            CUDA_TRY(cudaEventRecord(l.staged, l.stream));
            l.staged_pending = true;
            if (live_model && link_end(l.stream, (double)codes_bytes)) return fail(BGSA_ERR_CUDA, "event record failed");
        } else {
            // host -> device: the ASCII rows exactly as file.c:44-115 left them
            if (live_model && link_begin(l.stream)) return fail(BGSA_ERR_CUDA, "event record failed");
            CUDA_TRY(cudaMemcpyAsync(l.d_rows.p, subjects->content + (size_t)(first + off) * (slen + 1), row_bytes,
                                     cudaMemcpyHostToDevice, l.stream));
            if (live_model && link_end(l.stream, (double)row_bytes)) return fail(BGSA_ERR_CUDA, "event record failed");
        }
        mark(0);
        if (!fused && !host_pack) {
            cudaError_t e = launch_pack(plan.layout, l.d_rows.p, slen, n, l.d_packed.p, ctx->sm_count, l.stream);
            if (e != cudaSuccess) return fail(BGSA_ERR_CUDA, "pack kernel launch failed: %s", cudaGetErrorString(e));
            g_launches.fetch_add(1);
        }
        mark(1);
        if (li != 0) CUDA_TRY(cudaStreamWaitEvent(l.stream, job.tab_ready, 0));
        rc = run_align(plan, ctx->sm_count, d_tab, static_cast<unsigned long long *>(l.d_counters.p), n_queries, query_len,
                       fused ? nullptr : l.d_packed.p, slen, n, l.d_results.p, n, l.stream, nullptr, fused ? l.d_rows.p : nullptr);
        if (rc) return rc;
        mark(2);
        // device -> host: [query][subject] rows into the caller's (possibly wider) result matrix
        CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(results) + esize * (size_t)off, esize * (size_t)result_stride, l.d_results.p,
                                   esize * (size_t)n, esize * (size_t)n, (size_t)n_queries, cudaMemcpyDeviceToHost, l.stream));
        mark(3);
        if (tn.armed) { CUDA_TRY(cudaEventRecord(job.ev_lane_end[li], l.stream)); job.lane_used[li] = true; }
        if (kTrace) job.trace.push_back(tr);
    }
    job.link_rate = link_rate;                          // the next job on this slot starts from what this one measured
    job.last_share = hp_mode == HP_ALWAYS ? 1.0 : (hp_mode == HP_NEVER ? 0.0 : (share >= 0.0 ? share : -1.0));
    if (kTrace && (hp_mode == HP_HYBRID || tunable))
        fprintf(stderr, "[bgsa trace] front end: share of host-packed chunks %.3f%s, link %.1f GB/s, host pack %.1f GB/s (measured)\n",
                share, live_model ? " (live model)" : (tn.armed && tn.trial ? " (trial)" : ""), link_rate / 1e9, pack_rate / 1e9);
    return BGSA_OK;
}

int bgsa_align_batch_submit(const bgsa_params_t *p, const char *queries, int n_queries, int query_len,
                            const bgsa_seq_t *subjects, int64_t first, int64_t count, void *results, int64_t result_stride,
                            int device, int slot) {
    const int rc = submit_impl(p, queries, n_queries, query_len, subjects, first, count, results, result_stride, device, slot);
    if (rc == BGSA_ERR_CUDA || rc == BGSA_ERR_NOMEM) {
        // A failure in the middle of the chunk loop leaves copies and kernels of the earlier chunks in flight, some of
        // them writing into the caller's `results`: nothing may outlive the failed call, so drain the job's streams
        // (the message of the original failure is kept).
        if (device >= 0 && device < kMaxDevices && slot >= 0 && slot <= 1 && g_ctx[device].ready)
            for (Lane &l : g_ctx[device].job[slot].lane) cudaStreamSynchronize(l.stream);
        g_ctx[device >= 0 && device < kMaxDevices ? device : 0].job[slot >= 0 && slot <= 1 ? slot : 0].qc.key.clear();
    }
    return rc;
}

int bgsa_align_batch_wait(int device, int slot) {
    if (slot < 0 || slot > 1) return fail(BGSA_ERR_ARG, "slot must be 0 or 1");
    DeviceCtx *ctx;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    Job &job = ctx->job[slot];
    for (Lane &l : job.lane) CUDA_TRY(cudaStreamSynchronize(l.stream));
    Job::Tuner &tn = job.tuner;
    if (tn.armed) {                                             // the finished job's verdict (front-end tuner, submit_impl)
        tn.armed = false;
        float ms = 0.f, worst = 0.f;
        for (int i = 0; i < kLanesPerJob; i++)
            if (job.lane_used[i] && cudaEventElapsedTime(&ms, job.ev_begin, job.ev_lane_end[i]) == cudaSuccess && ms > worst) worst = ms;
        if (worst > 0.f) {
            const double r = tn.bytes / (1e-3 * worst);
            if (tn.trial) {
                tn.trial = false;
                const bool less_host_work = tn.p_trial < tn.p;
                if (r > tn.r_best * 1.03 || (less_host_work && r >= tn.r_best)) {
                    tn.p = tn.p_trial; tn.r_best = r; tn.rejected = 0;          // accepted: keep walking this way
                } else {
                    const bool at_edge = tn.p <= 0.0 || tn.p >= 1.0;
                    tn.dir = -tn.dir;
                    tn.rejected += at_edge ? 2 : 1;
                    if (tn.rejected >= 2) {                                     // both neighbours are worse: settle for a while
                        tn.rejected = 0;
                        if (tn.step > 0.0625) tn.step *= 0.5;
                        tn.hold = tn.hold_len;
                        if (tn.hold_len < 64) tn.hold_len *= 2;
                    }
                }
            } else {
                tn.r_best = tn.r_best > 0.0 ? 0.5 * tn.r_best + 0.5 * r : r;
            }
        } else {
            cudaGetLastError();
        }
    }
    if (!job.trace.empty()) {
        fprintf(stderr, "[bgsa trace] device %d slot %d: chunk(first,count,lane)  h2d_done pack_done align_done d2h_done [ms since submit]\n",
                device, slot);
        for (ChunkTrace &c : job.trace) {
            float t[4] = {-1.f, -1.f, -1.f, -1.f};
            for (int i = 0; i < 4; i++) {
                if (c.ev[i]) { cudaEventElapsedTime(&t[i], job.t0, c.ev[i]); cudaEventDestroy(c.ev[i]); }
            }
            fprintf(stderr, "[bgsa trace]   %10lld %9lld %d   %8.3f %8.3f %8.3f %8.3f  %s\n", (long long)c.off, (long long)c.n, c.lane,
                    t[0], t[1], t[2], t[3], c.host_packed ? "host-packed" : "ascii");
        }
        job.trace.clear();
    }
    return BGSA_OK;
}

int bgsa_batch_front_end(int device, int slot, double *host_pack_share) {
    if (slot < 0 || slot > 1 || !host_pack_share) return fail(BGSA_ERR_ARG, "bgsa_batch_front_end: bad argument");
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    *host_pack_share = ctx->job[slot].last_share;
    return BGSA_OK;
}

int bgsa_align_batch(const bgsa_params_t *p, const char *queries, int n_queries, int query_len, const bgsa_seq_t *subjects,
                     int64_t first, int64_t count, void *results, int64_t result_stride, int device) {
    int rc = bgsa_align_batch_submit(p, queries, n_queries, query_len, subjects, first, count, results, result_stride, device, 0);
    if (rc) return rc;
    return bgsa_align_batch_wait(device, 0);
}

// ---- per-chunk entry behind the reference's kernel symbols (include/align_core.h) -------------
int bgsa_align_peq_chunk(const bgsa_params_t *p, const char *query, int query_len, const void *peq, int word_bytes,
                         int v_num, int usable_bits, int word_num, int subject_len, int64_t n_subjects, void *results,
                         int device) {
    Plan plan;
    int rc = make_plan(p, query_len, subject_len, &plan);
    if (rc) return rc;
    if (!query || !peq || !results || n_subjects < 0 || (word_bytes != 4 && word_bytes != 8) || v_num < 1 ||
        usable_bits < 1 || usable_bits > 8 * word_bytes || word_num < 1 || n_subjects % v_num != 0)
        return fail(BGSA_ERR_ARG, "bgsa_align_peq_chunk: bad argument");
    DeviceCtx *ctx;
    rc = get_ctx(device, &ctx);
    if (rc) return rc;
    if (n_subjects == 0) return BGSA_OK;
    static std::mutex chunk_mu;                 // the reference calls its kernel from an OpenMP team
    std::lock_guard<std::mutex> lk(chunk_mu);
    Lane &l = ctx->chunk_lane;                  // its own stream and buffers: never collides with a slot-1 batch
    const size_t peq_bytes = (size_t)word_bytes * 5 * word_num * (size_t)n_subjects;
    const size_t esize = plan.result_size;
    if ((rc = l.d_rows.ensure(peq_bytes))) return rc;
    if ((rc = l.d_packed.ensure((size_t)packed_bytes(subject_len, n_subjects)))) return rc;
    if ((rc = l.d_results.ensure(esize * (size_t)n_subjects))) return rc;
    if ((rc = l.d_counters.ensure(sizeof(unsigned long long)))) return rc;
    CUDA_TRY(cudaMemcpyAsync(l.d_rows.p, peq, peq_bytes, cudaMemcpyHostToDevice, l.stream));
    const int head = plan.algo == BGSA_BANDED_MYERS ? plan.e : 0;
    cudaError_t e = launch_unpeq(plan.layout, word_bytes, l.d_rows.p, word_num, usable_bits, head, subject_len, n_subjects, v_num,
                                 l.d_packed.p, ctx->sm_count, l.stream);
    if (e != cudaSuccess) return fail(BGSA_ERR_CUDA, "unpeq kernel launch failed: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1);
    std::vector<char> qrow(query, query + query_len);
    qrow.push_back('\n');
    const void *d_tab;
    rc = stage_queries(ctx->chunk_qc, plan, qrow.data(), 1, query_len, subject_len, l.stream, &d_tab);
    if (rc) return rc;
    rc = run_align(plan, ctx->sm_count, d_tab, static_cast<unsigned long long *>(l.d_counters.p), 1, query_len, l.d_packed.p,
                   subject_len, n_subjects, l.d_results.p, n_subjects, l.stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(results, l.d_results.p, esize * (size_t)n_subjects, cudaMemcpyDeviceToHost, l.stream));
    CUDA_TRY(cudaStreamSynchronize(l.stream));
    return BGSA_OK;
}

int bgsa_int_peak(int device, double *lane_ops_per_s, double *sm_clock_mhz) {
    if (!lane_ops_per_s) return fail(BGSA_ERR_ARG, "lane_ops_per_s is NULL");
    DeviceCtx *ctx;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    unsigned int *d_sink;
    CUDA_TRY(cudaMalloc(&d_sink, 64));
    CUDA_TRY(cudaMemset(d_sink, 0, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    cudaStream_t st = ctx->job[0].lane[0].stream;
    const int iters = 20000;
    double best = 0.0, best_mhz = 0.0;
    for (int rep = 0; rep < 4; rep++) {           // rep 0 = warm-up
        const long long init[2] = {0x7fffffffffffffffLL, 0LL};
        CUDA_TRY(cudaMemcpyAsync(d_sink + 2, init, sizeof(init), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(e0, st));
        cudaError_t e = launch_int_peak(ctx->sm_count, iters, d_sink, st);
        if (e != cudaSuccess) return fail(BGSA_ERR_CUDA, "int_peak launch failed: %s", cudaGetErrorString(e));
        g_launches.fetch_add(1);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        long long span[2] = {0, 0};
        CUDA_TRY(cudaMemcpy(span, d_sink + 2, sizeof(span), cudaMemcpyDeviceToHost));
        const long long cycles = span[1] - span[0];
        const double ops = (double)ctx->sm_count * 8 * 256 * (double)iters * 64.0;
        const double rate = ops / (ms * 1e-3);
        if (rep > 0 && rate > best) { best = rate; best_mhz = (double)cycles / (ms * 1e3); }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_sink);
    *lane_ops_per_s = best;
    if (sm_clock_mhz) *sm_clock_mhz = best_mhz;
    return BGSA_OK;
}

}  // extern "C"
