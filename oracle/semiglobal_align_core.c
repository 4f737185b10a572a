/*
 * oracle/semiglobal_align_core.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference ships no C for semi-global Myers: it only exists as generator source
 * (generator/.../MyersGenerator.java:56-223, emitted with CPUArch.java:23-44, 64-bit element
 * constants from Element64.java:10-14, globals from GeneratorUtils.java:36-72 with
 * full_bits = 1, :66-70) and the JRE needed to run generator.jar is absent.  This file is OUR
 * restatement of what `java -jar generator.jar -m 0 -s -a none -e 64` would emit, written
 * against the reference's own align_core.h ABI so that it can be dropped into the unmodified
 * original/BGSA_CPU host pipeline (oracle/Makefile compiles it with the reference sources in
 * place).  It is pinned by the only golden output the reference checks in:
 * banded/BGSA_KNC/data/result.txt (384 int16 scores on sample-data), see tests/test_oracle_pins.py.
 *
 * Semantics: bit-vector runs along the SUBJECT ("read", Peq built by
 * cpu_handle_reads with 64 bits per word because full_bits = 1), text = QUERY ("ref").  The top
 * row delta is 0 (free start in the query), pv = ~0 / mv = 0 (subject consumed entirely), and
 * the result is the minimum over all query columns (column 0 included) of the bottom-row
 * score, times factor -1.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
#include "cal.h"
#include "align_core.h"

int match_score = 0;
int mismatch_score = -1;
int gap_score = -1;
int dvdh_len = 16;
int full_bits = 1;

void align_cpu(char * ref, cpu_read_t * read, int ref_len, int read_len, int word_num, int chunk_read_num, int result_index, cpu_write_t * results, cpu_data_t * dvdh_bit_mem) {
    const int word_size = CPU_WORD_SIZE;                      /* full-width words (MyersGenerator.java:64) */
    const int last_shift = (read_len - 1) % word_size;        /* :80 */
    const int common_shift = word_size - 1;                   /* :81 */
    int tid = omp_get_thread_num();
    cpu_data_t * pv_arr = &dvdh_bit_mem[(int64_t)tid * word_num * dvdh_len];              /* :95 */
    cpu_data_t * mv_arr = pv_arr + word_num;                                              /* :96 */
    cpu_read_t * read_base = read;

    for (int k = 0; k < chunk_read_num; k++) {
        cpu_read_t * peq = &read_base[(int64_t)k * word_num * CPU_V_NUM * CHAR_NUM];      /* :105 */
        for (int j = 0; j < word_num; j++) { mv_arr[j] = 0; pv_arr[j] = ~(cpu_data_t)0; } /* :107-112 */
        int64_t score = read_len;                                                         /* :115 */
        int64_t min_score = score;                                                        /* :117 */

        for (int i = 0; i < ref_len; i++) {
            cpu_read_t * matchv = &peq[((int)ref[i]) * CPU_V_NUM * word_num];             /* :123 */
            int64_t h_out = 0;                       /* semi-global: top-row delta 0 (:125-126) */
            for (int j = 0; j < word_num; j++) {
                int shift = (j == word_num - 1) ? last_shift : common_shift;  /* :150-154 vs :188-192 */
                int64_t h_in = h_out;                                         /* :133 */
                cpu_data_t h_in_neg = ((cpu_data_t)h_in >> 1) & 1;            /* 1 iff h_in == -1 (:134-135) */
                cpu_data_t pv = pv_arr[j], mv = mv_arr[j];
                cpu_data_t eq = matchv[j * CPU_V_NUM];                        /* :138-139 */
                cpu_data_t xv = eq | mv;                                      /* :140 */
                eq |= h_in_neg;                                               /* :141 */
                cpu_data_t xh = (((eq & pv) + pv) ^ pv) | eq;                 /* :142-145 */
                cpu_data_t ph = mv | ~(xh | pv);                              /* :146-148 */
                cpu_data_t mh = pv & xh;                                      /* :149 */
                h_out = (int64_t)((ph >> shift) & 1) - (int64_t)((mh >> shift) & 1);   /* :150-154 */
                ph <<= 1; mh <<= 1;                                           /* :155-156 */
                mh |= h_in_neg;                                               /* :157 */
                ph |= ((cpu_data_t)(h_in + 1)) >> 1;                          /* 1 iff h_in == +1 (:158-160) */
                pv_arr[j] = mh | ~(xv | ph);                                  /* :161-163 */
                mv_arr[j] = ph & xv;                                          /* :164 */
            }
            score += h_out;                                                   /* :205 */
            if (score < min_score) min_score = score;                         /* :207 */
        }

        score = min_score * -1;                                               /* :40-45, factor -1 */
        int index = result_index * CPU_V_NUM;
        int * vec_dump = (int *)&score;         /* low 32 bits, then narrowed to int16 (:46-52) */
        for (int i = 0; i < CPU_V_NUM; i++) results[index + i] = vec_dump[i];
        result_index++;
    }
}
