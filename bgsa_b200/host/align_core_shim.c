/*
 * align_core_shim.c -- the reference's kernel symbols, forwarding to the GPU (libbgsa_b200.so).
 * One object per reference variant (-DBGSA_SHIM_<VARIANT>), see include/align_core.h.
 * Replaces the generated align_core.c of that variant directory.
 */
#include <stdio.h>
#include <stdlib.h>

#include "../../include/align_core.h"
#include "../../include/bgsa_b200.h"

#if defined(BGSA_SHIM_MYERS_CPU)
int match_score = 0, mismatch_score = -1, gap_score = -1, dvdh_len = 16, full_bits = 0;   /* original/BGSA_CPU/align_core.c:13-17 */
#define SHIM_ALGO BGSA_MYERS_GLOBAL
#define SHIM_V 1
#define SHIM_WORD 8
#define SHIM_USABLE 63
#elif defined(BGSA_SHIM_SEMIGLOBAL_CPU)
int match_score = 0, mismatch_score = -1, gap_score = -1, dvdh_len = 16, full_bits = 1;   /* GeneratorUtils.java:66-70 */
#define SHIM_ALGO BGSA_MYERS_SEMIGLOBAL
#define SHIM_V 1
#define SHIM_WORD 8
#define SHIM_USABLE 64
#elif defined(BGSA_SHIM_BANDED_CPU)
int match_score = 0, mismatch_score = -1, gap_score = -1, dvdh_len = 16, full_bits = 0;   /* banded/BGSA_CPU/align_core.c:13-17 */
#define SHIM_ALGO BGSA_BANDED_MYERS
#define SHIM_V 1
#define SHIM_WORD 8
#define SHIM_USABLE 64
#elif defined(BGSA_SHIM_MYERS_SSE)
int match_score = 0, mismatch_score = -1, gap_score = -1, dvdh_len = 16, full_bits = 0;   /* original/BGSA_SSE/align_core.c:13-17 */
#define SHIM_ALGO BGSA_MYERS_GLOBAL
#define SHIM_V 4
#define SHIM_WORD 4
#define SHIM_USABLE 31
#elif defined(BGSA_SHIM_BITPAL_AVX2)
int match_score = 2, mismatch_score = -3, gap_score = -5, dvdh_len = 20, full_bits = 0;   /* original/BGSA_AVX2/align_core.c:13-17 */
#define SHIM_ALGO BGSA_BITPAL_PACKED
#define SHIM_V 8
#define SHIM_WORD 4
#define SHIM_USABLE 31
#elif defined(BGSA_SHIM_BITPAL_AVX512)
int match_score = 2, mismatch_score = -3, gap_score = -5, dvdh_len = 20, full_bits = 0;   /* original/BGSA_AVX512/align_core.c:13-17 */
#define SHIM_ALGO BGSA_BITPAL_PACKED
#define SHIM_V 16
#define SHIM_WORD 4
#define SHIM_USABLE 31
#else
#error "define one BGSA_SHIM_<VARIANT>"
#endif

static void forward(char *ref, void *read, int ref_len, int read_len, int word_num, int chunk_read_num, int result_index,
                    void *results, int esize) {
    bgsa_params_t p;
    bgsa_params_default(&p, SHIM_ALGO);
    p.match = match_score; p.mismatch = mismatch_score; p.gap = gap_score;
#if defined(BGSA_SHIM_BANDED_CPU)
    p.threshold = threshold;
#endif
    /* the reference kernels return void and treat bad input as UB; a failing GPU
     * call has no channel to report through, so it is fatal like the pipeline's own errors (file.c:13-16) */
    int rc = bgsa_align_peq_chunk(&p, ref, ref_len, read, SHIM_WORD, SHIM_V, SHIM_USABLE, word_num, read_len,
                                  (int64_t)chunk_read_num * SHIM_V, (char *)results + (size_t)esize * result_index * SHIM_V, 0);
    if (rc != BGSA_OK) {
        printf("Error - GPU alignment failed: %s\n", bgsa_last_error());
        exit(1);
    }
}

#if defined(BGSA_SHIM_MYERS_CPU) || defined(BGSA_SHIM_SEMIGLOBAL_CPU)
void align_cpu(char *ref, uint64_t *read, int ref_len, int read_len, int word_num, int chunk_read_num, int result_index,
               int16_t *results, uint64_t *dvdh_bit_mem) {
    (void)dvdh_bit_mem;
    forward(ref, read, ref_len, read_len, word_num, chunk_read_num, result_index, results, 2);
}
#elif defined(BGSA_SHIM_BANDED_CPU)
void align_cpu(char *query, uint64_t *read, int query_len, int subject_len, int word_num, int chunk_read_num,
               int result_index, int8_t *score, uint64_t *dvdh_bit_mem) {
    (void)dvdh_bit_mem;
    forward(query, read, query_len, subject_len, word_num, chunk_read_num, result_index, score, 1);
}
#elif defined(BGSA_SHIM_MYERS_SSE)
void align_sse(char *ref, uint32_t *read, int ref_len, int read_len, int word_num, int chunk_read_num, int result_index,
               int16_t *results, void *dvdh_bit_mem) {
    (void)dvdh_bit_mem;
    forward(ref, read, ref_len, read_len, word_num, chunk_read_num, result_index, results, 2);
}
#elif defined(BGSA_SHIM_BITPAL_AVX2)
void align_avx(char *ref, uint32_t *read, int ref_len, int read_len, int word_num, int chunk_read_num, int result_index,
               int16_t *results, void *dvdh_bit_mem) {
    (void)dvdh_bit_mem;
    forward(ref, read, ref_len, read_len, word_num, chunk_read_num, result_index, results, 2);
}
#else
void align_mic(char *ref, uint32_t *read, int ref_len, int read_len, int word_num, int chunk_read_num, int result_index,
               int16_t *results, void *dvdh_bit_mem) {
    (void)dvdh_bit_mem;
    forward(ref, read, ref_len, read_len, word_num, chunk_read_num, result_index, results, 2);
}
#endif
