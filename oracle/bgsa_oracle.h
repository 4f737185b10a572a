/*
 * oracle/bgsa_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the BGSA one-query-vs-many-subjects hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libbgsa_b200.so) never does and has no CPU fallback.
 *
 * Pinning (tests/test_oracle_pins.py): checked against
 *   - the reference's only checked-in golden output, banded/BGSA_KNC/data/result.txt
 *     (semi-global Myers on sample-data, 384 int16 scores),
 *   - the unmodified reference compiled from /root/reference into oracle/_ref/
 *     (original/BGSA_CPU, BGSA_SSE, BGSA_AVX2, BGSA_AVX512, banded/BGSA_CPU),
 *   - plain O(nm) DP (edit distance, Needleman-Wunsch, semi-global minimum).
 * BitPAl non-packed and (M,I,G) other than (2,-3,-5) have no reference artefact: for those
 * parity is pinned by DP equivalence only ("parity unpinned" by a reference build).
 */
#ifndef BGSA_ORACLE_H
#define BGSA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORACLE_MYERS_GLOBAL = 0,     /* original/BGSA_CPU/align_core.c:19-148                     */
    ORACLE_MYERS_SEMIGLOBAL = 1, /* generator/.../MyersGenerator.java:56-223                  */
    ORACLE_BANDED_MYERS = 2,     /* banded/BGSA_CPU/align_core.c:69-252                       */
    ORACLE_BITPAL_PACKED = 3,    /* original/BGSA_AVX512/align_core.c:19-485                  */
    ORACLE_BITPAL_NONPACKED = 4, /* generator/.../BitPAlGenerator.java:939-1061,1392-1701     */
    ORACLE_BITPAL_PACKED_SEMIGLOBAL = 5 /* generator -s: BitPAlGenerator.java:77-80,112-114,289-308 */
};

/* global.c:9-15 -- A,C,G,T,N -> 0..4, every other byte -> 0 */
int oracle_map_char(int c);

/* One (query, subject) pair.  `q` = query codes 0..4 (file.c:135-139), `s` = subject ASCII. */
int16_t oracle_myers_global(const char *q, int qlen, const char *s, int slen);
int16_t oracle_myers_semiglobal(const char *q, int qlen, const char *s, int slen);
/* `s` must have `e` readable bytes after s[slen-1] (the reference Peq builder reads them,
 * banded/BGSA_CPU/global.c:52-82); batch rows of stride slen+1 satisfy this except the last
 * row, which the caller must pad. */
int8_t oracle_banded_myers(const char *q, int qlen, const char *s, int slen, int e);
int8_t oracle_banded_myers_w(const char *q, int qlen, const char *s, int slen, int e, int wordbits);
int16_t oracle_bitpal_packed(const char *q, int qlen, const char *s, int slen, int M, int I, int G);
int16_t oracle_bitpal_nonpacked(const char *q, int qlen, const char *s, int slen, int M, int I, int G);
int16_t oracle_bitpal_packed_semiglobal(const char *q, int qlen, const char *s, int slen, int M, int I, int G);

/* Plain dynamic programming (independent mathematical cross-check; no narrowing). */
int oracle_dp_edit(const char *q, int qlen, const char *s, int slen);
int oracle_dp_nw(const char *q, int qlen, const char *s, int slen, int M, int I, int G);
/* max over substrings of s of the NW score of the WHOLE of q against it (Appendix A9, BitPAl orientation) */
int oracle_dp_nw_semiglobal(const char *q, int qlen, const char *s, int slen, int M, int I, int G);
/* min over substrings of q of edit distance to the WHOLE of s (Appendix A9, Myers orientation) */
int oracle_dp_semiglobal(const char *q, int qlen, const char *s, int slen);

/*
 * Batch driver = cpu_cal_align_score (original/BGSA_CPU/cal_cpu.c:43-85): results are
 * [query][subject] row-major.  queries: code rows of stride qlen+1; subjects: ASCII rows of
 * stride slen+1 (seq_t.content layout, global.h:9-16).  `out` is int16 except for
 * ORACLE_BANDED_MYERS where it is int8.  The subjects buffer must have `e` readable bytes
 * after its last row for ORACLE_BANDED_MYERS.  Returns 0, or -1 on bad arguments.
 */
int oracle_align_batch(int algo, int M, int I, int G, int e,
                       const char *queries, int n_queries, int qlen,
                       const char *subjects, int64_t n_subjects, int slen,
                       void *out, int threads);

#ifdef __cplusplus
}
#endif
#endif
