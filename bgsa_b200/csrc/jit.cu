// jit.cu -- run-time instantiation of BitPAl kernel instances (see jit.h).
//
//   source   : a three-line translation unit that includes bitpal.cuh / rows_kernel.cuh -- the headers the built-in
//              instances were compiled from, embedded as strings by the Makefile (build/jit_sources.inc) and handed to
//              NVRTC as in-memory headers, so a JIT instance cannot drift from the compiled ones;
//   compile  : NVRTC (dlopen, no link-time dependency), --gpu-architecture=sm_100a, straight to a cubin;
//   cache    : $BGSA_JIT_CACHE or ~/.cache/bgsa_b200/<fnv64 of sources + options + spec>.bin (lowered kernel names +
//              cubin), written atomically; one compile per scheme and geometry per machine;
//   launch   : cudaLibraryLoadData / cudaLibraryGetKernel (context-independent), then the same launch geometry as the
//              built-in instances (launch.cuh).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "bitpal.cuh"
#include "jit.h"

namespace bgsa {

namespace {

struct EmbeddedSource { const char *name; const char *text; };
#include "jit_sources.inc"      // static const EmbeddedSource kJitSources[] = {...};

// ---- NVRTC, loaded on first use --------------------------------------------------------------------------------------
struct Nvrtc {
    void *handle = nullptr;
    std::string why;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcAddNameExpression) AddNameExpression = nullptr;
    decltype(&nvrtcGetLoweredName) GetLoweredName = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
    decltype(&nvrtcVersion) Version = nullptr;
};
Nvrtc &nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        if (getenv("BGSA_NO_JIT")) { n.why = "switched off by BGSA_NO_JIT"; return; }
        for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "libnvrtc.so.13"}) {
            n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.handle) break;
        }
        if (!n.handle) { n.why = std::string("libnvrtc not found: ") + dlerror(); return; }
        bool ok = true;
#define SYM(f) ok = ok && (n.f = reinterpret_cast<decltype(n.f)>(dlsym(n.handle, "nvrtc" #f))) != nullptr
        SYM(CreateProgram); SYM(DestroyProgram); SYM(CompileProgram); SYM(AddNameExpression); SYM(GetLoweredName);
        SYM(GetCUBINSize); SYM(GetCUBIN); SYM(GetProgramLogSize); SYM(GetProgramLog); SYM(GetErrorString); SYM(Version);
#undef SYM
        if (!ok) { n.why = "libnvrtc lacks a required entry point"; dlclose(n.handle); n.handle = nullptr; }
    });
    return n;
}

uint64_t fnv64(uint64_t h, const void *data, size_t n) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

std::string algo_expr(const JitSpec &s) {
    char buf[256];
    if (s.variant == 0)
        snprintf(buf, sizeof(buf), "bgsa::BitpalNonPacked<bgsa::Scheme<%d, %d, %d>, %d>", s.M, s.I, s.G, s.K);
    else
        snprintf(buf, sizeof(buf), "bgsa::BitpalPacked<bgsa::Scheme<%d, %d, %d>, %d, %d>", s.M, s.I, s.G, s.K, s.variant == 2 ? 1 : 0);
    return buf;
}
int unroll_of(const JitSpec &s) { return s.variant == 0 ? 1 : 2; }     // as inst_bitpal.cu
bool has_rows(const JitSpec &s) { return s.L == 1 && s.K <= kRowsMaxK; }

// -default-device: the constexpr helpers of bitpal.cuh carry no execution-space annotation (host functions do not exist under NVRTC)
const char *kOptions[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "-DBGSA_JIT=1"};
constexpr int kNumOptions = sizeof(kOptions) / sizeof(kOptions[0]);

std::string cache_dir() {
    if (const char *e = getenv("BGSA_JIT_CACHE")) return e;
    const char *home = getenv("HOME");
    return std::string(home && *home ? home : "/tmp") + "/.cache/bgsa_b200";
}
void mkdirs(const std::string &path) {
    for (size_t i = 1; i <= path.size(); i++)
        if (i == path.size() || path[i] == '/') mkdir(path.substr(0, i).c_str(), 0755);
}

struct Blob { std::string name_align, name_rows; std::vector<char> cubin; };

std::string cache_file(const JitSpec &s) {
    uint64_t h = 1469598103934665603ull;
    for (const EmbeddedSource &e : kJitSources) h = fnv64(h, e.text, strlen(e.text));
    for (const char *o : kOptions) h = fnv64(h, o, strlen(o));
    int major = 0, minor = 0;
    if (nvrtc().handle) nvrtc().Version(&major, &minor);
    const int key[8] = {s.variant, s.M, s.I, s.G, s.K, s.L, major, minor};
    h = fnv64(h, key, sizeof(key));
    char name[64];
    snprintf(name, sizeof(name), "/bgsa_%016llx.bin", (unsigned long long)h);
    return cache_dir() + name;
}
bool read_blob(const std::string &file, Blob *b) {
    FILE *f = fopen(file.c_str(), "rb");
    if (!f) return false;
    auto rd_str = [&](std::string *out) {
        uint32_t n = 0;
        if (fread(&n, 4, 1, f) != 1 || n > 4096) return false;
        out->resize(n);
        return n == 0 || fread(&(*out)[0], 1, n, f) == n;
    };
    uint32_t magic = 0;
    uint64_t size = 0;
    bool ok = fread(&magic, 4, 1, f) == 1 && magic == 0x42475341u && rd_str(&b->name_align) && rd_str(&b->name_rows) &&
              fread(&size, 8, 1, f) == 1 && size > 0 && size < (1ull << 30);
    if (ok) { b->cubin.resize(size); ok = fread(b->cubin.data(), 1, size, f) == size; }
    fclose(f);
    return ok;
}
void write_blob(const std::string &file, const Blob &b) {
    mkdirs(cache_dir());
    char tmp[64];
    snprintf(tmp, sizeof(tmp), ".tmp%d", (int)getpid());
    const std::string t = file + tmp;
    FILE *f = fopen(t.c_str(), "wb");
    if (!f) return;                                     // an unwritable cache only costs the next process a compile
    auto wr_str = [&](const std::string &s) { const uint32_t n = (uint32_t)s.size(); fwrite(&n, 4, 1, f); fwrite(s.data(), 1, n, f); };
    const uint32_t magic = 0x42475341u;
    const uint64_t size = b.cubin.size();
    fwrite(&magic, 4, 1, f); wr_str(b.name_align); wr_str(b.name_rows); fwrite(&size, 8, 1, f);
    const bool ok = fwrite(b.cubin.data(), 1, size, f) == size;
    if (fclose(f) == 0 && ok) rename(t.c_str(), file.c_str()); else remove(t.c_str());
}

int compile(const JitSpec &s, Blob *out, std::string *err) {
    Nvrtc &n = nvrtc();
    if (!n.handle) { *err = n.why; return 1; }
    const std::string file = cache_file(s);
    if (!getenv("BGSA_JIT_NO_CACHE") && read_blob(file, out)) return 0;
    const std::string algo = algo_expr(s);
    char expr[512];
    snprintf(expr, sizeof(expr), "bgsa::align_kernel<%s, %d, %d, %d, %d>", algo.c_str(), s.L, kAlignCH, kAlignThreads, unroll_of(s));
    const std::string e_align = expr;
    snprintf(expr, sizeof(expr), "bgsa::align_rows_kernel<%s, %d, %d>", algo.c_str(), kAlignThreads, unroll_of(s));
    const std::string e_rows = expr;
    const char *src = "#include \"bitpal.cuh\"\n#include \"rows_kernel.cuh\"\n";
    std::vector<const char *> names, texts;
    for (const EmbeddedSource &e : kJitSources) { names.push_back(e.name); texts.push_back(e.text); }
    nvrtcProgram prog;
    nvrtcResult r = n.CreateProgram(&prog, src, "bgsa_jit.cu", (int)names.size(), texts.data(), names.data());
    if (r != NVRTC_SUCCESS) { *err = std::string("nvrtcCreateProgram: ") + n.GetErrorString(r); return 1; }
    n.AddNameExpression(prog, e_align.c_str());
    if (has_rows(s)) n.AddNameExpression(prog, e_rows.c_str());
    r = n.CompileProgram(prog, kNumOptions, kOptions);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.GetProgramLog(prog, &log[0]);
        *err = std::string("NVRTC failed for ") + algo + ": " + n.GetErrorString(r) + "\n" + log.substr(0, 1500);
        n.DestroyProgram(&prog);
        return 1;
    }
    const char *low = nullptr;
    bool ok = n.GetLoweredName(prog, e_align.c_str(), &low) == NVRTC_SUCCESS && low;
    if (ok) out->name_align = low;
    if (ok && has_rows(s)) { ok = n.GetLoweredName(prog, e_rows.c_str(), &low) == NVRTC_SUCCESS && low; if (ok) out->name_rows = low; }
    size_t cs = 0;
    ok = ok && n.GetCUBINSize(prog, &cs) == NVRTC_SUCCESS && cs > 0;
    if (ok) { out->cubin.resize(cs); ok = n.GetCUBIN(prog, out->cubin.data()) == NVRTC_SUCCESS; }
    n.DestroyProgram(&prog);
    if (!ok) { *err = "NVRTC produced no cubin / lowered names"; return 1; }
    if (!getenv("BGSA_JIT_NO_CACHE")) write_blob(file, *out);
    return 0;
}

// ---- loaded instances ------------------------------------------------------------------------------------------------
struct Loaded { cudaLibrary_t lib = nullptr; cudaKernel_t align = nullptr, rows = nullptr; Blob blob; };
std::mutex g_jit_mu;
std::map<std::vector<int>, Loaded *> g_loaded;

int get_loaded(const JitSpec &s, Loaded **out, std::string *err) {
    std::lock_guard<std::mutex> lk(g_jit_mu);
    const std::vector<int> key = {s.variant, s.M, s.I, s.G, s.K, s.L};
    auto it = g_loaded.find(key);
    if (it != g_loaded.end()) { *out = it->second; return 0; }
    Loaded *l = new Loaded;
    if (compile(s, &l->blob, err)) { delete l; return 1; }
    cudaError_t e = cudaLibraryLoadData(&l->lib, l->blob.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&l->align, l->lib, l->blob.name_align.c_str());
    if (e == cudaSuccess && !l->blob.name_rows.empty()) e = cudaLibraryGetKernel(&l->rows, l->lib, l->blob.name_rows.c_str());
    if (e != cudaSuccess) { *err = std::string("loading the JIT cubin failed: ") + cudaGetErrorString(e); delete l; return 1; }
    g_loaded[key] = l;
    *out = l;
    return 0;
}

}  // namespace

bool jit_available(std::string *why) {
    Nvrtc &n = nvrtc();
    if (!n.handle && why) *why = n.why;
    return n.handle != nullptr;
}

bool jit_scheme_ok(int variant, int M, int I, int G, std::string *why) {
    auto no = [&](const char *m) { if (why) *why = m; return false; };
    if (!(G < 0 && M > I && M >= 0)) return no("needs gap < 0, match > mismatch, match >= 0");
    if (M > 64 || I < -128 || G < -128) return no("scores out of range (|score| <= 128, match <= 64)");
    const int F = common_factor(M, I, G);
    const int m = M / F, i = I / F, g = G / F;
    const int A = m - 2 * g, B = (i - 2 * g) > 0 ? (i - 2 * g) : 0;
    if (A < 1 || B >= A) return no("degenerate scheme (match - 2 gap must exceed max(mismatch - 2 gap, 0))");
    // code size and registers grow with the number of delta values: packed keeps ceil(log2(A+1)) planes but one add
    // chain per high class, non-packed one vector per value and O(A^2) terms
    if (variant == 0 ? A > 24 : A > 63) return no("score range too wide for a bit-parallel kernel (match - 2 gap > 63, non-packed > 24)");
    return true;
}

int jit_precompile(const JitSpec &spec, std::string *err) {
    Blob b;
    return compile(spec, &b, err);
}

cudaError_t launch_bitpal_jit(const JitSpec &s, const LaunchArgs &a, std::string *err) {
    Loaded *l = nullptr;
    if (get_loaded(s, &l, err)) return cudaErrorJitCompilationDisabled;
    BitpalParams prm{0};
    PackedSubjects ps = a.ps;
    const uint32_t *peq = a.d_peq;
    int nq = a.n_queries > 0 ? a.n_queries : 1, qlen = a.qlen;
    int16_t *res = static_cast<int16_t *>(a.d_results);
    long long stride = a.result_stride;
    unsigned long long *counters = a.d_counters;
    constexpr int WARPS = kAlignThreads / 32;
    if (a.d_ascii) {
        if (!l->rows) { *err = "no rows kernel for this instance"; return cudaErrorInvalidValue; }
        const void *kern = l->rows;
        const size_t table = sizeof(uint32_t) * 256 * peq_row_stride(s.K, 1);
        const size_t tb = (size_t)rows_stage_bytes(ps.slen + 1);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        int occ1 = 0, occ2 = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, kAlignThreads, table + WARPS * 2 * tb)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, kern, kAlignThreads, table + WARPS * tb)) != cudaSuccess) return e;
        int nstage = (occ2 >= occ1 || occ2 >= 3) ? 2 : 1;
        const int occ = nstage == 2 ? occ2 : occ1;
        if (occ < 1) return cudaErrorInvalidConfiguration;
        if (a.dry_run) {
            if (a.resident_subjects) *a.resident_subjects = (long long)a.sm_count * occ * WARPS * 32;
            return cudaSuccess;
        }
        if ((e = cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * nq, a.stream)) != cudaSuccess) return e;
        long long want = (ps.ntiles * nq + WARPS - 1) / WARPS;
        const long long resident = (long long)a.sm_count * occ;
        if (want > resident) want = resident;
        if (want < 1) want = 1;
        const uint8_t *rows = a.d_ascii;
        int slen = ps.slen;
        long long count = ps.count;
        void *args[] = {&rows, &slen, &count, &peq, &nq, &qlen, &res, &stride, &prm, &counters, &nstage};
        return cudaLaunchKernel(kern, dim3((unsigned)want), dim3(kAlignThreads), args, table + WARPS * nstage * tb, a.stream);
    }
    const void *kern = l->align;
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kAlignThreads, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    if (a.dry_run) {
        if (a.resident_subjects) *a.resident_subjects = (long long)a.sm_count * occ * WARPS * (32 / s.L);
        return cudaSuccess;
    }
    if ((e = cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * nq, a.stream)) != cudaSuccess) return e;
    long long want = (ps.ntiles * s.L * nq + WARPS - 1) / WARPS;
    const long long resident = (long long)a.sm_count * occ;
    if (want > resident) want = resident;
    if (want < 1) want = 1;
    void *args[] = {&ps, &peq, &nq, &qlen, &res, &stride, &prm, &counters};
    return cudaLaunchKernel(kern, dim3((unsigned)want), dim3(kAlignThreads), args, 0, a.stream);
}

}  // namespace bgsa
