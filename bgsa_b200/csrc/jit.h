// jit.h -- BitPAl kernels for scoring schemes that were not compiled into the library, instantiated at run time.
//
// The reference runs its Java generator once per scoring scheme and compiles the emitted align_core.c
// (generator/src/main/java/edu/sdu/hpcl/bgsa/Main.java:240-315, BitPAlGenerator.java:151-534 packed, :1392-1701 non-packed).
// Here a scheme is a template instance: the ones listed at build time (make SCHEMES=...) are in the library, any other
// valid (match, mismatch, gap) is instantiated by NVRTC from the very same kernel headers (embedded in the library),
// cached on disk, and launched through the runtime's cudaLibrary interface.
#pragma once

#include <string>

#include "launch.cuh"

namespace bgsa {

struct JitSpec {
    int variant;      // 0 non-packed, 1 packed global, 2 packed semi-global
    int M, I, G;      // as the caller gave them (the common factor is divided out by Scheme<>)
    int K, L;         // instance geometry (instances.h tables)
};

// NVRTC can be loaded in this process (libnvrtc.so.12 / libnvrtc.so)
bool jit_available(std::string *why);
// The scheme satisfies what Scheme<M, I, G> static_asserts and stays within sane code size; *why says what does not.
bool jit_scheme_ok(int variant, int M, int I, int G, std::string *why);
// Compiles (or finds in the disk cache) the kernels of `spec`; needs no GPU.  0 = ok.
int jit_precompile(const JitSpec &spec, std::string *err);
// Launches the instance on a.stream; a.d_ascii selects the rows kernel (only when rows_kernel_fits).
cudaError_t launch_bitpal_jit(const JitSpec &spec, const LaunchArgs &a, std::string *err);

}  // namespace bgsa
