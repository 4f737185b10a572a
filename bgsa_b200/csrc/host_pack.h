// host_pack.h -- ASCII subject rows -> packed tiles ON THE HOST CORES (optional front end of the batch entry).
//
// The boundary hands the library one byte per base (seq_t.content, file.c:44-115); the kernels want two bits.  For
// short reads the alignment kernel is faster than the PCIe copy of the ASCII rows, and the host cores -- which the
// reference keeps busy with its Peq build (cpu_handle_reads, original/BGSA_CPU/global.c:25-70: byte -> bit scatter under
// OpenMP) -- would sit idle.  bgsa_align_batch_submit can therefore run the SAME encoding as pack_stream_kernel
// (pack.cuh) on a pool of host threads, chunk by chunk into pinned staging buffers, and ship a quarter of the bytes.
// Output is bit-identical to the device pack kernels (tests/test_host_logic.py::test_host_pack_matches_numpy and the
// GPU tier), so the alignment kernels cannot tell the difference.
#pragma once

#include <stddef.h>
#include <stdint.h>

namespace bgsa {

// Packs tiles [tile_begin, tile_end) of `count` rows (stride slen+1) into the tile layout that starts at `packed`
// (make_packed_view(packed, slen, count)); layout: 0 = 2-bit codes, 1 = bit-planes (banded).  The N plane of a tile is
// written only when the tile holds an 'N'; returns true when any tile in the range does.
bool host_pack_tiles(int layout, const uint8_t *rows, int slen, int64_t count, void *packed, int64_t tile_begin, int64_t tile_end);

// "avx2" or "scalar": the encoder the running CPU gets.
const char *host_pack_isa();

// A small fixed pool of worker threads; parallel_for blocks until every index has run (the caller works too).
// Several callers may use the pool at the same time.
class HostPool {
public:
    static HostPool &instance();
    int threads() const { return nthreads_; }
    // fn(i) for i in [0, n): indices are handed out one at a time in increasing order
    void parallel_for(int64_t n, void (*fn)(int64_t, void *), void *arg);
    ~HostPool();

private:
    HostPool();
    struct Impl;
    Impl *impl_;
    int nthreads_;
};

}  // namespace bgsa
