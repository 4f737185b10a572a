/*
 * include/align_core.h -- drop-in replacement for the reference's generated kernel header
 * (original/BGSA_CPU/align_core.h:8, original/BGSA_SSE/align_core.h, original/BGSA_AVX2/align_core.h,
 *  original/BGSA_AVX512/align_core.h:8, banded/BGSA_CPU/align_core.h:8).
 *
 * The symbols below have the reference's exact signatures and argument meaning, so the
 * reference's unmodified host pipeline (main.c, cal_*.c, thread.c, global.c, file.c) links
 * against libalign_core_<variant>.so instead of the generated align_core.c and produces the same
 * result file, with the DP running on the GPU.  Each shim object also defines the globals the
 * generated file defines (align_core.c:13-17): match_score, mismatch_score, gap_score, dvdh_len,
 * full_bits.
 *
 *   ref            query codes 0..4 (not NUL terminated)
 *   read           Peq block of this chunk, built by <arch>_handle_reads (global.c:25-70)
 *   ref_len        query length            read_len   subject length
 *   word_num       words per match mask    chunk_read_num  vector groups in this chunk
 *   result_index   first group index; results[(result_index + k) * V_NUM + lane] is written
 *   dvdh_bit_mem   scratch of the CPU kernels -- unused here (state lives in GPU registers)
 *
 * Which variant a shim implements is fixed when it is compiled (bgsa_b200/host/align_core_shim.c,
 * -DBGSA_SHIM_<VARIANT>), exactly like the reference fixes it when the generator is run.
 * These per-chunk calls move a few hundred subjects each; pipelines that want GPU throughput
 * call bgsa_align_batch (bgsa_b200.h) once per bucket instead -- see INTEGRATION.md.
 */
#ifndef _ALIGN_CORE_H_
#define _ALIGN_CORE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

extern int match_score, mismatch_score, gap_score, dvdh_len, full_bits;

#if defined(BGSA_SHIM_MYERS_CPU) || defined(BGSA_SHIM_SEMIGLOBAL_CPU)
/* original/BGSA_CPU: CPU_V_NUM 1, 64-bit words, int16 results (config.h:19-27) */
void align_cpu(char *ref, uint64_t *read, int ref_len, int read_len, int word_num, int chunk_read_num,
               int result_index, int16_t *results, uint64_t *dvdh_bit_mem);
#elif defined(BGSA_SHIM_BANDED_CPU)
/* banded/BGSA_CPU: int8 results, reads the global `threshold` set by -k (global.h:43, main.c:43,62-64) */
extern int threshold;
void align_cpu(char *query, uint64_t *read, int query_len, int subject_len, int word_num, int chunk_read_num,
               int result_index, int8_t *score, uint64_t *dvdh_bit_mem);
#elif defined(BGSA_SHIM_MYERS_SSE)
/* original/BGSA_SSE: SSE_V_NUM 4, 32-bit lanes; dvdh_bit_mem is __m128i* in the reference */
void align_sse(char *ref, uint32_t *read, int ref_len, int read_len, int word_num, int chunk_read_num,
               int result_index, int16_t *results, void *dvdh_bit_mem);
#elif defined(BGSA_SHIM_BITPAL_AVX2)
/* original/BGSA_AVX2: AVX_V_NUM 8; dvdh_bit_mem is __m256i* in the reference */
void align_avx(char *ref, uint32_t *read, int ref_len, int read_len, int word_num, int chunk_read_num,
               int result_index, int16_t *results, void *dvdh_bit_mem);
#elif defined(BGSA_SHIM_BITPAL_AVX512)
/* original/BGSA_AVX512: MIC_V_NUM 16; dvdh_bit_mem is __m512i* in the reference */
void align_mic(char *ref, uint32_t *read, int ref_len, int read_len, int word_num, int chunk_read_num,
               int result_index, int16_t *results, void *dvdh_bit_mem);
#endif

#ifdef __cplusplus
}
#endif
#endif
