// query_peq.h -- host-side construction of the QUERY match masks (Peq).
//
// Counterpart of cpu_handle_reads (original/BGSA_CPU/global.c:25-70), but for the query instead
// of every subject (the DP is transposed, see bgsa_common.cuh): 5 rows (A C G T N) of
// peq_row_stride(K, L) words; lane r of a group owns words [r*K, (r+1)*K) of the bit-vector and
// finds them at row + r * peq_kp(K).  Query bytes are the codes 0..4 that get_ref_from_file()
// leaves in ref_seq.content (file.c:135-139); anything else is treated as 0 like the reference's
// zero-initialised mapping table would have produced.
#pragma once
#include <stdint.h>
#include <string.h>

namespace bgsa {

constexpr int h_peq_kp(int k) { return (k + 3) / 4 * 4; }
constexpr int h_peq_stride(int k) { return ((k + 3) / 4 | 1) * 4; }
constexpr int h_peq_row_stride(int k, int lanes) { return h_peq_stride(h_peq_kp(k) * lanes); }

inline void build_query_peq(const char *codes, int qlen, int K, int L, uint32_t *out) {
    const int kp = h_peq_kp(K), stride = h_peq_row_stride(K, L);
    memset(out, 0, sizeof(uint32_t) * 5 * stride);
    for (int i = 0; i < qlen; i++) {
        int c = (unsigned char)codes[i];
        if (c > 4) c = 0;
        const int w = i / 32, lane = w / K, j = w % K;
        out[c * stride + lane * kp + j] |= 1u << (i % 32);
    }
}

}  // namespace bgsa
