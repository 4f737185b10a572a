/*
 * oracle/bgsa_oracle.c -- TEST INFRASTRUCTURE ONLY (see bgsa_oracle.h for the pinning status).
 *
 * Plain-C restatement of the reference's algorithms for the one-query-vs-many-subjects path.
 * Every function keeps the REFERENCE orientation (bit-vector along the subject, one DP column
 * per query character) and the reference's word geometry, and cites the file:line it follows.
 * Paths are relative to /root/reference.
 */
#include "bgsa_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CHAR_NUM 5                       /* original/BGSA_CPU/config.h:17 */
#define W63 63                           /* CPU_WORD_SIZE - 1, align_core.c:22 */
#define MASK63 0x7fffffffffffffffULL     /* carry_bitmask, align_core.c:37 */

/* ------------------------------------------------------------------------------------------
 * Alphabet: original/BGSA_CPU/global.c:6-15.  mapping_table is a zero-initialised global, so
 * every byte other than A,C,G,T,N maps to 0 (= A).  (Bytes >= 128 index out of bounds in the
 * reference; we define them as 0 too.)
 * ---------------------------------------------------------------------------------------- */
int oracle_map_char(int c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case 'N': return 4;
        default:  return 0;
    }
}

/* Peq builder: original/BGSA_CPU/global.c:25-70 for ONE subject (CPU_V_NUM = 1).
 * out[c * word_num + j] has bit (i % wbits) set iff map(s[i]) == c and i / wbits == j. */
static void build_peq(const char *s, int slen, int wbits, int word_num, uint64_t *out) {
    memset(out, 0, sizeof(uint64_t) * CHAR_NUM * word_num);
    for (int i = 0; i < slen; i++) {
        int c = oracle_map_char((unsigned char)s[i]);
        out[c * word_num + i / wbits] |= 1ULL << (i % wbits);
    }
}

/* ------------------------------------------------------------------------------------------
 * a1. Myers global, unit cost: original/BGSA_CPU/align_core.c:19-148.
 * 63 cells per 64-bit word; bit 63 carries the add overflow into the next word (:79-83) and the
 * HP/HN shift carry is taken from bit 63 after the shift (:91-96).  Top row delta = +1
 * (HP_shift = 1, :68).  Score read from bit (read_len-1)%63 of the last word (:43,121-124).
 * ---------------------------------------------------------------------------------------- */
int16_t oracle_myers_global(const char *q, int qlen, const char *s, int slen) {
    int word_num = (slen + W63 - 1) / W63;                         /* cal_cpu.c:255 (full_bits = 0) */
    uint64_t *peq = (uint64_t *)malloc(sizeof(uint64_t) * (CHAR_NUM + 2) * word_num);
    uint64_t *VN = peq + CHAR_NUM * word_num, *VP = VN + word_num;
    build_peq(s, slen, W63, word_num, peq);
    for (int j = 0; j < word_num; j++) { VN[j] = 0; VP[j] = MASK63; }          /* :58-61 */
    uint64_t maskh = 1ULL << ((slen - 1) % W63);                               /* :43 */
    int64_t score = slen;                                                      /* :63 */
    for (int i = 0; i < qlen; i++) {
        const uint64_t *matchv = &peq[(int)q[i] * word_num];                   /* :67 */
        uint64_t hp_in = 1, hn_in = 0, sum = 0;                                /* :68-70 */
        for (int j = 0; j < word_num; j++) {
            uint64_t vn = VN[j], vp = VP[j];
            uint64_t pm = matchv[j] | vn;                                      /* :77 */
            uint64_t cin = sum >> W63;                                         /* :79 */
            sum = (vp & pm) + vp + cin;                                        /* :80-82 */
            uint64_t d0 = ((sum & MASK63) ^ vp) | pm;                          /* :83-85 */
            uint64_t hp = vn | ~(d0 | vp);                                     /* :86-88 */
            uint64_t hn = d0 & vp;                                             /* :89 */
            if (j == word_num - 1) {                                           /* :121-124 */
                if (hn & maskh) score--;
                else if (hp & maskh) score++;
            }
            hp = (hp << 1) | hp_in;  hp_in = hp >> W63;                        /* :91-93 */
            hn = (hn << 1) | hn_in;  hn_in = hn >> W63;                        /* :94-96 */
            VP[j] = (hn | ~(d0 | hp)) & MASK63;                                /* :97-100 */
            VN[j] = (d0 & hp) & MASK63;                                        /* :101-102 */
        }
    }
    score *= -1;                                                               /* factor, :44,137 */
    free(peq);
    return (int16_t)(int32_t)(uint32_t)(uint64_t)score;    /* low 32 bits -> int16 (:139-144) */
}

/* ------------------------------------------------------------------------------------------
 * a4. Myers semi-global: generator/.../MyersGenerator.java:56-223 (CPUArch, 64-bit element).
 * Full 64-bit words (full_bits = 1, GeneratorUtils.java:66-70), block formulation with
 * h_in/h_out in {-1,0,+1}; top-row delta 0; result = min over query columns incl. column 0.
 * ---------------------------------------------------------------------------------------- */
int16_t oracle_myers_semiglobal(const char *q, int qlen, const char *s, int slen) {
    int word_num = (slen + 63) / 64;                               /* cal_cpu.c:252-253 */
    uint64_t *peq = (uint64_t *)malloc(sizeof(uint64_t) * (CHAR_NUM + 2) * word_num);
    uint64_t *MV = peq + CHAR_NUM * word_num, *PV = MV + word_num;
    build_peq(s, slen, 64, word_num, peq);
    for (int j = 0; j < word_num; j++) { MV[j] = 0; PV[j] = ~0ULL; }           /* :107-112 */
    int last_shift = (slen - 1) % 64;                                          /* :80 */
    int64_t score = slen, min_score = slen;                                    /* :115-118 */
    for (int i = 0; i < qlen; i++) {
        const uint64_t *matchv = &peq[(int)q[i] * word_num];                   /* :123 */
        int64_t h_out = 0;                                                     /* :125-126 */
        for (int j = 0; j < word_num; j++) {
            int shift = (j == word_num - 1) ? last_shift : 63;                 /* :150-154 / :188-192 */
            int64_t h_in = h_out;                                              /* :133 */
            uint64_t neg = ((uint64_t)h_in >> 1) & 1;                          /* :134-135 */
            uint64_t pv = PV[j], mv = MV[j], eq = matchv[j];
            uint64_t xv = eq | mv;                                             /* :140 */
            eq |= neg;                                                         /* :141 */
            uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;                        /* :142-145 */
            uint64_t ph = mv | ~(xh | pv);                                     /* :146-148 */
            uint64_t mh = pv & xh;                                             /* :149 */
            h_out = (int64_t)((ph >> shift) & 1) - (int64_t)((mh >> shift) & 1);
            ph = (ph << 1) | (((uint64_t)(h_in + 1)) >> 1);                    /* :155,158-160 */
            mh = (mh << 1) | neg;                                              /* :156-157 */
            PV[j] = mh | ~(xv | ph);                                           /* :161-163 */
            MV[j] = ph & xv;                                                   /* :164 */
        }
        score += h_out;                                                        /* :205 */
        if (score < min_score) min_score = score;                              /* :207 */
    }
    free(peq);
    int64_t out = min_score * -1;                                              /* :40-45 */
    return (int16_t)(int32_t)(uint32_t)(uint64_t)out;
}

/* ------------------------------------------------------------------------------------------
 * a2. Banded Myers verifier: banded/BGSA_CPU/align_core.c:69-252 with the Peq layout of
 * banded/BGSA_CPU/global.c:25-84 and word_num of banded/BGSA_CPU/cal_cpu.c:253-254.
 * The control flow (checkpoints after row min(64,q), every 16 rows, remainder) is replicated
 * statement by statement because the early-exit value 127 depends on it (Appendix A6).
 * ---------------------------------------------------------------------------------------- */
#define BANDED_WORD 64          /* CPU_WORD_SIZE, banded/BGSA_CPU/config.h:24 */
#define BANDED_BATCH 16         /* batch_size,    banded/BGSA_CPU/config.h:20 */
#define BANDED_MAX_ERROR 127    /* MAX_ERROR,     banded/BGSA_CPU/config.h:19 */

static int8_t banded_core(const char *q, int qlen, const char *s, int slen, int e, uint64_t wmask);
int8_t oracle_banded_myers(const char *q, int qlen, const char *s, int slen, int e) {
    return banded_core(q, qlen, s, slen, e, ~0ULL);
}
/* Same recurrence with every band word truncated to `wordbits` bits: used by the tests to show
 * that the result does not depend on the word width as long as the band fits (Appendix A8). */
int8_t oracle_banded_myers_w(const char *q, int qlen, const char *s, int slen, int e, int wordbits) {
    return banded_core(q, qlen, s, slen, e, wordbits >= 64 ? ~0ULL : ((1ULL << wordbits) - 1));
}
static int8_t banded_core(const char *q, int qlen, const char *s, int slen, int e, uint64_t wmask) {
    int h_threshold = e + slen - qlen;                                          /* align_core.c:70 */
    int band_down = e + h_threshold;                                            /* :71-72 */
    int word_num = (slen - h_threshold + BANDED_WORD - 1) / BANDED_WORD + 1;    /* cal_cpu.c:254 */
    /* Peq: word 0 holds subject[0..e) at bits e+1..2e; words 1.. hold subject[e..e+slen)
     * (global.c:45-82; note it reads e bytes past the subject).  We allocate one extra word so
     * that the tmp_peq reloads at :151-155,:180-184 stay in bounds like the reference's
     * bucket-sized allocation does. */
    int alloc_words = word_num + 2;
    uint64_t *dist = (uint64_t *)calloc((size_t)CHAR_NUM * alloc_words, sizeof(uint64_t));
#define DIST(c, w) dist[(c) * alloc_words + (w)]
    for (int i = 0; i < e; i++)                                                 /* global.c:52-62 */
        DIST(oracle_map_char((unsigned char)s[i]), 0) |= 1ULL << (e + 1 + i);
    for (int i = 0; i < slen; i++) {                                            /* global.c:67-82 */
        int w = 1 + i / BANDED_WORD;
        if (w < word_num)   /* the reference writes exactly word_num words per subject */
            DIST(oracle_map_char((unsigned char)s[e + i]), w) |= 1ULL << (i % BANDED_WORD);
    }
    uint64_t peq[CHAR_NUM], tmp_peq[CHAR_NUM];
    for (int c = 0; c < CHAR_NUM; c++) { peq[c] = DIST(c, 0); tmp_peq[c] = DIST(c, 1); }  /* :85-94 */

    uint64_t VN = 0, VP = 0, X, D0 = 0, HN, HP;                                 /* :96-101 */
    int i_bd = h_threshold, last_bits = h_threshold, bit_index = 0, query_index = 0;
    uint64_t err = (uint64_t)e;                                                 /* :113 */
    uint64_t max_err = (uint64_t)(e + last_bits + 1);                           /* :114 */
    int8_t result;

#define CAL_D0(c)                        /* :19-33 */ \
    X = peq[c] | VN; D0 = X & VP; D0 = (D0 + VP) & wmask; D0 = D0 ^ VP; D0 = D0 | X; \
    HN = D0 & VP; HP = D0 | VP; HP = ~HP & wmask; HP = HP | VN; \
    X = D0 >> 1; VN = X & HP; VP = HP | X; VP = ~VP; VP = VP | HN; \
    D0 &= wmask; VN &= wmask; VP &= wmask;
#define CAL_SCORE() err += 1 - (D0 & 1);                                        /* :64-67 */
#define MOVE_PEQ() for (int c_ = 0; c_ < CHAR_NUM; c_++) peq[c_] >>= 1;         /* :35-40 */
#define OR_PEQ()   for (int c_ = 0; c_ < CHAR_NUM; c_++) \
                       peq[c_] |= ((tmp_peq[c_] >> bit_index) & 1ULL) << band_down;   /* :42-62 */
#define LOAD_TMP(w) for (int c_ = 0; c_ < CHAR_NUM; c_++) tmp_peq[c_] = DIST(c_, (w));

    for (; query_index < e; query_index++) {                                    /* :116-123 */
        CAL_D0((int)q[query_index]); MOVE_PEQ(); OR_PEQ(); bit_index++; i_bd++;
    }
    int length = BANDED_WORD < qlen ? BANDED_WORD : qlen;                       /* :125 */
    for (; query_index < length; query_index++) {                               /* :126-134 */
        CAL_D0((int)q[query_index]); CAL_SCORE(); MOVE_PEQ(); OR_PEQ(); bit_index++; i_bd++;
    }
    if (err > max_err) { result = BANDED_MAX_ERROR; goto end; }                 /* :136-140 */

    if (qlen > BANDED_WORD) {                                                   /* :142 */
        bit_index = 0;
        int rest_length = slen - i_bd;                                          /* :144 */
        int batch_count = rest_length / BANDED_BATCH;                           /* :145 */
        int word_count = rest_length / BANDED_WORD;                             /* :146 */
        int word_batch_count = BANDED_WORD / BANDED_BATCH;                      /* :147 */
        int batch_index = 0, word_index = 2;
        LOAD_TMP(word_index);                                                   /* :151-155 */
        for (int i = 0; i < word_count; i++) {                                  /* :157 */
            for (int j = 0; j < word_batch_count; j++) {
                for (int k = 0; k < BANDED_BATCH; k++) {                        /* :159-168 */
                    CAL_D0((int)q[query_index]); CAL_SCORE(); MOVE_PEQ(); OR_PEQ();
                    bit_index++; i_bd++; query_index++;
                }
                if (err > max_err) { result = BANDED_MAX_ERROR; goto end; }     /* :170-174 */
                batch_index++;
            }
            bit_index = 0; word_index++;
            LOAD_TMP(word_index);                                               /* :180-184 */
        }
        for (; batch_index < batch_count; batch_index++) {                      /* :187 */
            for (int k = 0; k < BANDED_BATCH; k++) {
                CAL_D0((int)q[query_index]); CAL_SCORE(); MOVE_PEQ(); OR_PEQ();
                bit_index++; i_bd++; query_index++;
            }
            if (err > max_err) { result = BANDED_MAX_ERROR; goto end; }         /* :199-203 */
        }
        for (; i_bd < slen; i_bd++) {                                           /* :206-214 */
            CAL_D0((int)q[query_index]); CAL_SCORE(); MOVE_PEQ(); OR_PEQ();
            bit_index++; query_index++;
        }
        if (err > max_err) { result = BANDED_MAX_ERROR; goto end; }             /* :216-220 */
        for (; query_index < qlen; query_index++) {                             /* :221-226 */
            CAL_D0((int)q[query_index]); CAL_SCORE(); MOVE_PEQ();
        }
    }
    {
        uint64_t min_err = err;                                                 /* :231 */
        for (int i = 0; i <= last_bits; i++) {                                  /* :232-240 */
            err += (VP >> i) & 1;
            err -= (VN >> i) & 1;
            /* reference compares cpu_data_t (uint64) values */
            min_err = min_err < err ? min_err : err;
        }
        result = (int8_t)(int64_t)min_err;                                      /* :242-245 */
    }
end:
    free(dist);
    return result;
#undef DIST
#undef CAL_D0
#undef CAL_SCORE
#undef MOVE_PEQ
#undef OR_PEQ
#undef LOAD_TMP
}

/* ------------------------------------------------------------------------------------------
 * BitPAl helpers: multi-word bit-vectors of 63 usable bits per uint64 (the reference keeps the
 * top bit of every lane word free for the add overflow, original/BGSA_AVX512/align_core.c:28,
 * 214,222).  Word width does not change the mathematics, only the carry plumbing.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int M, I, G, factor;     /* scores after division by their common factor (Main.java:213-267) */
    int A;                   /* maxLength = max - min = M - 2G   (ScoreMsg.java:23-43) */
    int B;                   /* mid - min   = I - 2G (clamped at 0)                    */
    int nb;                  /* maxBitsNum = ceil(log2(A+1)) + 1 planes incl. sign      */
} bitpal_cfg;

static int common_factor(int M, int I, int G) {       /* Main.java:213-238 */
    int factor = 1;
    int m = abs(I), g = abs(G);
    int mn = M == 0 ? m : M;
    if (m < mn) mn = m;
    if (g < mn) mn = g;
    for (int i = 2; i <= mn; i++)
        if (M % i == 0 && m % i == 0 && g % i == 0) factor = i;
    return factor;
}

static int bitpal_setup(bitpal_cfg *c, int M, int I, int G) {
    if (G >= 0 || M < I || M < 0) return -1;
    c->factor = common_factor(M, I, G);
    c->M = M / c->factor; c->I = I / c->factor; c->G = G / c->factor;
    c->A = c->M - 2 * c->G;
    c->B = c->I - 2 * c->G; if (c->B < 0) c->B = 0;
    if (c->B > c->A) return -1;
    int bits = 0; while ((1 << bits) < c->A + 1) bits++;
    c->nb = bits + 1;
    return 0;
}

static void mw_add(uint64_t *out, const uint64_t *a, const uint64_t *b, int nw) {
    uint64_t carry = 0;                          /* overflowK, align_core.c:216-222 */
    for (int j = 0; j < nw; j++) {
        uint64_t sum = a[j] + b[j] + carry;
        carry = sum >> W63;
        out[j] = sum & MASK63;
    }
}
static void mw_shl1(uint64_t *out, const uint64_t *a, int nw, uint64_t carry_in) {
    uint64_t carry = carry_in;                   /* init_*_prevbit / bitK, align_core.c:226-229,333-360 */
    for (int j = 0; j < nw; j++) {
        uint64_t v = (a[j] << 1) | carry;
        carry = v >> W63;
        out[j] = v & MASK63;
    }
}

/* Final score, genPackedScore / genUnpackedScore (BitPAlGenerator.java:67-148, 939-1061;
 * original/BGSA_AVX512/align_core.c:432-480): G*ref_len + sum over subject positions of the
 * delta, then * factor, low 32 bits -> int16.  `d` = per-position (delta - G). */
static int16_t bitpal_finish(const bitpal_cfg *c, int qlen, int slen, int64_t sum_d) {
    int64_t score = (int64_t)c->G * qlen + sum_d + (int64_t)c->G * slen;
    score *= c->factor;
    return (int16_t)(int32_t)(uint32_t)(uint64_t)score;
}

/* ------------------------------------------------------------------------------------------
 * a3. BitPAl packed, general (M,I,G): original/BGSA_AVX512/align_core.c:19-485 is the
 * (2,-3,-5) instance; BitPAlGenerator.java:151-534,2246-2716 is the general emitter.
 *
 * State per subject position p: d_p = (delta along the subject) - G in [0, A], stored as the
 * two's complement of -d_p in nb bit-planes (dvdh_bit1,2,4,...; all-zero = global start, :168).
 * Per query column:
 *   decode one-hot classes of d (:189-214); class max by one add-carry chain through runs of
 *   d = 0 (:216-224); classes max-1 .. mid+1 by one shift + one add-carry chain each (:225-279);
 *   binary-encode y = max(w, e_{p-1}) (:281-297); bit-sliced add y + (-d), clamp at 0 -> e
 *   (:299-331); shift e one position (:333-366); rewrite -d to -max(w, d) (:368-389);
 *   bit-sliced add + e_{p-1}, keep only negatives -> new -d (:391-428).
 * with w = A on a match and B on a mismatch.
 * ---------------------------------------------------------------------------------------- */
static int16_t bitpal_packed_core(const char *q, int qlen, const char *s, int slen, int M, int I, int G, int semi) {
    bitpal_cfg cfg;
    if (bitpal_setup(&cfg, M, I, G)) return 0;
    const int A = cfg.A, B = cfg.B, nb = cfg.nb;
    const int nw = (slen + W63 - 1) / W63;                   /* cal_mic.c:305-310 geometry */
    const int ndec = (B > A - B - 1 ? B : A - B - 1) + 1;    /* decoded low classes 0..ndec-1 */
    const unsigned full = (1u << nb) - 1u;

    uint64_t *peq = (uint64_t *)malloc(sizeof(uint64_t) * CHAR_NUM * nw);
    build_peq(s, slen, W63, nw, peq);
    /* scratch vectors */
    size_t nvec = (size_t)nb * 5 + ndec + (A + 1) + 8;
    uint64_t *mem = (uint64_t *)calloc(nvec * nw, sizeof(uint64_t));
    uint64_t *p = mem;
#define TAKE(n) (p += (size_t)(n) * nw, p - (size_t)(n) * nw)
    uint64_t *S = TAKE(nb), *yb = TAKE(nb), *e = TAKE(nb), *es = TAKE(nb), *sum = TAKE(nb);
    uint64_t *Dlow = TAKE(ndec), *Y = TAKE(A + 1);
    uint64_t *init = TAKE(1), *shifted = TAKE(1), *acc = TAKE(1), *remain = TAKE(1), *rest = TAKE(1);
    uint64_t *lowmask = TAKE(1), *carry = TAKE(1), *tmp = TAKE(1);
#undef TAKE
#define V(base, k) ((base) + (size_t)(k) * nw)
    uint64_t valid_last = (slen % W63) ? ((1ULL << (slen % W63)) - 1) : MASK63;
    if (semi) {
        /* writeBitInitStr (BitPAlGenerator.java:289-291,2201-2218): every delta along the subject starts at 0,
         * i.e. d = -G, planes = two's complement of G */
        const unsigned pat = ((1u << nb) - (unsigned)(-cfg.G)) & full;
        for (int b = 0; b < nb; b++)
            for (int j = 0; j < nw; j++) V(S, b)[j] = ((pat >> b) & 1) ? MASK63 : 0;
    }

    for (int col = 0; col < qlen; col++) {
        const uint64_t *eq = &peq[(int)q[col] * nw];
        /* decode: Dlow[v] = [d == v]  <=> planes == two's complement of -v   (:189-214) */
        for (int v = 0; v < ndec; v++) {
            unsigned pat = ((1u << nb) - (unsigned)v) & full;
            for (int j = 0; j < nw; j++) {
                uint64_t m = MASK63;
                for (int b = 0; b < nb; b++) m &= ((pat >> b) & 1) ? V(S, b)[j] : ~V(S, b)[j];
                V(Dlow, v)[j] = m;
            }
        }
        /* class A: e_{p-1} == A, shifted form (:216-224) */
        for (int j = 0; j < nw; j++) init[j] = V(Dlow, 0)[j] & eq[j];
        mw_add(acc, init, V(Dlow, 0), nw);
        for (int j = 0; j < nw; j++) {
            uint64_t x = (acc[j] ^ V(Dlow, 0)[j] ^ init[j]) & MASK63;
            remain[j] = V(Dlow, 0)[j] ^ init[j];             /* d == 0 and mismatch */
            V(Y, A)[j] = x | eq[j];                          /* dvpos7shiftormatch */
        }
        /* classes A-1 .. B+1 (:225-279) */
        for (int k = A - 1; k > B; k--) {
            for (int j = 0; j < nw; j++) {
                uint64_t v = 0;
                for (int dlt = 1; k + dlt <= A; dlt++) v |= V(Dlow, dlt)[j] & V(Y, k + dlt)[j];
                init[j] = v;
            }
            mw_shl1(shifted, init, nw, 0);
            mw_add(acc, shifted, remain, nw);
            for (int j = 0; j < nw; j++) V(Y, k)[j] = (acc[j] ^ remain[j]) & ~eq[j] & MASK63;
        }
        /* everything else: y == B (:281-285) */
        for (int j = 0; j < nw; j++) {
            uint64_t any = 0;
            for (int k = A; k > B; k--) any |= V(Y, k)[j];
            rest[j] = ~any & MASK63;
        }
        /* binary planes of y (:286-297) */
        for (int b = 0; b < nb; b++)
            for (int j = 0; j < nw; j++) {
                uint64_t v = ((B >> b) & 1) ? rest[j] : 0;
                for (int k = A; k > B; k--) if ((k >> b) & 1) v |= V(Y, k)[j];
                V(yb, b)[j] = v;
            }
        /* e = max(0, y - d): bit-sliced add of y and -d, clamp by the sign plane (:299-331) */
        for (int j = 0; j < nw; j++) {
            uint64_t c = 0;
            for (int b = 0; b < nb; b++) {
                uint64_t a = V(S, b)[j], y = V(yb, b)[j];
                V(sum, b)[j] = a ^ y ^ c;
                c = (a & y) | ((a ^ y) & c);
            }
            uint64_t neg = V(sum, nb - 1)[j];
            for (int b = 0; b < nb; b++) V(e, b)[j] = V(sum, b)[j] & ~neg;
        }
        /* shift e by one subject position; e_0 = 0 is the global boundary (:333-366) */
        for (int b = 0; b < nb; b++) mw_shl1(V(es, b), V(e, b), nw, 0);
        /* -d := -max(w, d) (:368-389) */
        for (int j = 0; j < nw; j++) {
            uint64_t lm = 0;
            for (int v = 0; v <= B; v++) lm |= V(Dlow, v)[j];
            lowmask[j] = lm & ~eq[j];
        }
        {
            unsigned patB = ((1u << nb) - (unsigned)B) & full, patA = ((1u << nb) - (unsigned)A) & full;
            for (int b = 0; b < nb; b++)
                for (int j = 0; j < nw; j++) {
                    uint64_t v = V(S, b)[j];
                    v = ((patB >> b) & 1) ? (v | lowmask[j]) : (v & ~lowmask[j]);
                    v = ((patA >> b) & 1) ? (v | eq[j]) : (v & ~eq[j]);
                    V(S, b)[j] = v & MASK63;
                }
        }
        /* new -d = min(0, -max(w,d) + e_{p-1}) (:391-428) */
        for (int j = 0; j < nw; j++) {
            uint64_t c = 0;
            for (int b = 0; b < nb; b++) {
                uint64_t a = V(S, b)[j], y = V(es, b)[j];
                V(sum, b)[j] = a ^ y ^ c;
                c = (a & y) | ((a ^ y) & c);
            }
            uint64_t neg = V(sum, nb - 1)[j];
            for (int b = 0; b < nb; b++) V(S, b)[j] = V(sum, b)[j] & neg;
        }
        (void)carry; (void)tmp;
    }
    if (semi) {
        /* genPackedScore with isSemiGlobal (BitPAlGenerator.java:67-148): score = G*ref_len, then one subject
         * position at a time score += delta, max_score = max(max_score, score); max_score * factor -> int16 */
        int64_t score = (int64_t)cfg.G * qlen, best = score;
        for (int i = 0; i < slen; i++) {
            const int j = i / W63, bit = i % W63;
            int64_t negd = 0;
            for (int b = 0; b < nb; b++) {
                const int64_t v = (int64_t)((V(S, b)[j] >> bit) & 1) << b;
                if (b == nb - 1) negd -= v; else negd += v;          /* two's complement value of the planes = -d */
            }
            score += -negd + cfg.G;
            if (score > best) best = score;
        }
        free(mem); free(peq);
        best *= cfg.factor;
        return (int16_t)(int32_t)(uint32_t)(uint64_t)best;
    }
    /* score (:432-471): planes hold -d as nb-bit two's complement */
    int64_t sum_d = 0;
    for (int j = 0; j < nw; j++) {
        uint64_t valid = (j == nw - 1) ? valid_last : MASK63;
        for (int b = 0; b < nb; b++) {
            int64_t cnt = __builtin_popcountll(V(S, b)[j] & valid);
            if (b == nb - 1) sum_d += cnt << b; else sum_d -= cnt << b;
        }
    }
    free(mem); free(peq);
    return bitpal_finish(&cfg, qlen, slen, sum_d);
#undef V
}

int16_t oracle_bitpal_packed(const char *q, int qlen, const char *s, int slen, int M, int I, int G) {
    return bitpal_packed_core(q, qlen, s, slen, M, I, G, 0);
}
/* BitPAl packed, semi-global (generator -s; BitPAlGenerator.java:77-80,112-114,289-308): the whole QUERY aligned
 * to the best substring of the SUBJECT.  No reference artefact exists for it (the
 * generator needs a JRE): "parity unpinned", checked against plain DP only. */
int16_t oracle_bitpal_packed_semiglobal(const char *q, int qlen, const char *s, int slen, int M, int I, int G) {
    return bitpal_packed_core(q, qlen, s, slen, M, I, G, 1);
}

/* ------------------------------------------------------------------------------------------
 * a5. BitPAl non-packed (one bit-vector per delta value): BitPAlGenerator.java:1392-1701,
 * score :939-1061.  D[v] = [d == v], v = 0..A  (dh_min .. dh_max), X[v] = [e_{p-1} == v].
 * ---------------------------------------------------------------------------------------- */
int16_t oracle_bitpal_nonpacked(const char *q, int qlen, const char *s, int slen, int M, int I, int G) {
    bitpal_cfg cfg;
    if (bitpal_setup(&cfg, M, I, G)) return 0;
    const int A = cfg.A, B = cfg.B;
    const int nw = (slen + W63 - 1) / W63;
    uint64_t *peq = (uint64_t *)malloc(sizeof(uint64_t) * CHAR_NUM * nw);
    build_peq(s, slen, W63, nw, peq);
    size_t nvec = (size_t)4 * (A + 1) + 6;
    uint64_t *mem = (uint64_t *)calloc(nvec * nw, sizeof(uint64_t));
    uint64_t *D = mem, *X = D + (size_t)(A + 1) * nw, *Mx = X + (size_t)(A + 1) * nw, *Dn = Mx + (size_t)(A + 1) * nw;
    uint64_t *init = Dn + (size_t)(A + 1) * nw, *shifted = init + nw, *acc = shifted + nw, *remain = acc + nw;
#define V(base, k) ((base) + (size_t)(k) * nw)
    for (int j = 0; j < nw; j++) V(D, 0)[j] = MASK63;        /* dh_min = carry_bitmask (:1494) */

    for (int col = 0; col < qlen; col++) {
        const uint64_t *eq = &peq[(int)q[col] * nw];
        /* X[A] by add-carry through runs of d == 0 (:1530-1539) */
        for (int j = 0; j < nw; j++) init[j] = V(D, 0)[j] & eq[j];
        mw_add(acc, init, V(D, 0), nw);
        for (int j = 0; j < nw; j++) {
            V(X, A)[j] = (acc[j] ^ V(D, 0)[j] ^ init[j]) & MASK63;
            remain[j] = V(D, 0)[j] ^ (init[j] & MASK63);
        }
        /* Y[k] in place of X[k] while building: Y[A] = X[A] | match, Y[k] = X[k] & ~match */
        /* high classes A-1 .. B+1 (:1545-1563) */
        for (int k = A - 1; k > B; k--) {
            for (int j = 0; j < nw; j++) {
                uint64_t v = V(D, A - k)[j] & (V(X, A)[j] | eq[j]);
                for (int x = 1; x <= A - 1 - k; x++) v |= V(D, A - k - x)[j] & (V(X, A - x)[j] & ~eq[j]);
                init[j] = v;
            }
            mw_shl1(shifted, init, nw, 0);
            mw_add(acc, shifted, remain, nw);
            for (int j = 0; j < nw; j++) V(X, k)[j] = (acc[j] ^ remain[j]) & MASK63;
        }
        /* "rest": neither match nor any high class (:1566-1574) */
        for (int j = 0; j < nw; j++) {
            uint64_t any = V(X, A)[j] | eq[j];
            for (int k = A - 1; k > B; k--) any |= V(X, k)[j];
            acc[j] = ~any & MASK63;            /* dvnot<max>to<mid>ormatch */
        }
        /* low classes B .. 1: plain shift, no propagation (:1581-1599) */
        for (int k = B; k >= 1; k--) {
            for (int j = 0; j < nw; j++) {
                uint64_t v = (A - k <= A) ? (V(D, A - k)[j] & (V(X, A)[j] | eq[j])) : 0;
                for (int h = A - 1; h > B; h--)
                    if (h - k >= 0) v |= V(D, h - k)[j] & (V(X, h)[j] & ~eq[j]);
                if (B - k >= 0) v |= V(D, B - k)[j] & acc[j];
                init[j] = v;
            }
            mw_shl1(V(X, k), init, nw, 0);
        }
        /* X[0] = complement (:1601-1611) */
        for (int j = 0; j < nw; j++) {
            uint64_t any = 0;
            for (int k = A; k >= 1; k--) any |= V(X, k)[j];
            V(X, 0)[j] = ~any & MASK63;
        }
        /* Mx[v] = [max(w, d) == v] (:1613-1629) */
        for (int j = 0; j < nw; j++) {
            uint64_t any = 0;
            V(Mx, A)[j] = V(D, A)[j] | eq[j];
            any |= V(Mx, A)[j];
            for (int v = A - 1; v > B; v--) { V(Mx, v)[j] = V(D, v)[j] & ~eq[j]; any |= V(Mx, v)[j]; }
            V(Mx, B)[j] = ~any & MASK63;                     /* dh<min>to<mid> */
        }
        /* new d == k  <=>  max(w,d) - e_{p-1} == k, k >= 1 (:1634-1661); d == 0 = complement (:1664-1671) */
        for (int j = 0; j < nw; j++) {
            uint64_t any = 0;
            for (int k = 1; k <= A; k++) {
                uint64_t v = 0;
                for (int m = (k > B ? k : B); m <= A; m++) v |= V(Mx, m)[j] & V(X, m - k)[j];
                V(Dn, k)[j] = v & MASK63; any |= v;
            }
            V(Dn, 0)[j] = ~any & MASK63;
        }
        memcpy(D, Dn, sizeof(uint64_t) * (size_t)(A + 1) * nw);
    }
    uint64_t valid_last = (slen % W63) ? ((1ULL << (slen % W63)) - 1) : MASK63;
    int64_t sum_d = 0;
    for (int j = 0; j < nw; j++) {
        uint64_t valid = (j == nw - 1) ? valid_last : MASK63;
        for (int v = 1; v <= A; v++) sum_d += (int64_t)v * __builtin_popcountll(V(D, v)[j] & valid);
    }
    free(mem); free(peq);
    return bitpal_finish(&cfg, qlen, slen, sum_d);
#undef V
}

/* ------------------------------------------------------------------------------------------
 * Plain DP cross-checks (no reference line: textbook recurrences).
 * ---------------------------------------------------------------------------------------- */
int oracle_dp_edit(const char *q, int qlen, const char *s, int slen) {
    int *row = (int *)malloc(sizeof(int) * (slen + 1));
    for (int j = 0; j <= slen; j++) row[j] = j;
    for (int i = 1; i <= qlen; i++) {
        int diag = row[0]; row[0] = i;
        for (int j = 1; j <= slen; j++) {
            int up = row[j];
            int best = diag + ((int)q[i - 1] != oracle_map_char((unsigned char)s[j - 1]));
            if (up + 1 < best) best = up + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            diag = up; row[j] = best;
        }
    }
    int r = row[slen]; free(row); return r;
}

int oracle_dp_nw(const char *q, int qlen, const char *s, int slen, int M, int I, int G) {
    int *row = (int *)malloc(sizeof(int) * (slen + 1));
    for (int j = 0; j <= slen; j++) row[j] = j * G;
    for (int i = 1; i <= qlen; i++) {
        int diag = row[0]; row[0] = i * G;
        for (int j = 1; j <= slen; j++) {
            int up = row[j];
            int best = diag + (((int)q[i - 1] == oracle_map_char((unsigned char)s[j - 1])) ? M : I);
            if (up + G > best) best = up + G;
            if (row[j - 1] + G > best) best = row[j - 1] + G;
            diag = up; row[j] = best;
        }
    }
    int r = row[slen]; free(row); return r;
}

/* whole query inside the subject: S[0][j] = 0, S[i][0] = i*G, answer = max_j S[qlen][j] */
int oracle_dp_nw_semiglobal(const char *q, int qlen, const char *s, int slen, int M, int I, int G) {
    int *row = (int *)malloc(sizeof(int) * (slen + 1));
    for (int j = 0; j <= slen; j++) row[j] = 0;
    for (int i = 1; i <= qlen; i++) {
        int diag = row[0]; row[0] = i * G;
        for (int j = 1; j <= slen; j++) {
            int up = row[j];
            int best = diag + (((int)q[i - 1] == oracle_map_char((unsigned char)s[j - 1])) ? M : I);
            if (up + G > best) best = up + G;
            if (row[j - 1] + G > best) best = row[j - 1] + G;
            diag = up; row[j] = best;
        }
    }
    int r = row[0];
    for (int j = 1; j <= slen; j++) if (row[j] > r) r = row[j];
    free(row); return r;
}

int oracle_dp_semiglobal(const char *q, int qlen, const char *s, int slen) {
    /* rows = subject (must be consumed), columns = query (free start and end) */
    int *col = (int *)malloc(sizeof(int) * (slen + 1));
    for (int i = 0; i <= slen; i++) col[i] = i;
    int best = col[slen];
    for (int j = 1; j <= qlen; j++) {
        int diag = col[0]; col[0] = 0;
        for (int i = 1; i <= slen; i++) {
            int left = col[i];
            int v = diag + ((int)q[j - 1] != oracle_map_char((unsigned char)s[i - 1]));
            if (left + 1 < v) v = left + 1;
            if (col[i - 1] + 1 < v) v = col[i - 1] + 1;
            diag = left; col[i] = v;
        }
        if (col[slen] < best) best = col[slen];
    }
    free(col); return best;
}

/* ------------------------------------------------------------------------------------------
 * a7. Batch driver: original/BGSA_CPU/cal_cpu.c:43-85 -- results [query][subject] row-major.
 * ---------------------------------------------------------------------------------------- */
int oracle_align_batch(int algo, int M, int I, int G, int e,
                       const char *queries, int n_queries, int qlen,
                       const char *subjects, int64_t n_subjects, int slen,
                       void *out, int threads) {
    if (!queries || !subjects || !out || qlen <= 0 || slen <= 0 || n_queries < 0 || n_subjects < 0) return -1;
    if (algo < ORACLE_MYERS_GLOBAL || algo > ORACLE_BITPAL_PACKED_SEMIGLOBAL) return -1;
    if (algo >= ORACLE_BITPAL_PACKED) { bitpal_cfg c; if (bitpal_setup(&c, M, I, G)) return -1; }
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    (void)threads;
#endif
    int64_t total = (int64_t)n_queries * n_subjects;
#pragma omp parallel for num_threads(threads) schedule(guided)
    for (int64_t t = 0; t < total; t++) {
        int64_t qi = t / n_subjects, si = t % n_subjects;
        const char *q = queries + qi * (qlen + 1);
        const char *s = subjects + si * (int64_t)(slen + 1);
        switch (algo) {
            case ORACLE_MYERS_GLOBAL:     ((int16_t *)out)[t] = oracle_myers_global(q, qlen, s, slen); break;
            case ORACLE_MYERS_SEMIGLOBAL: ((int16_t *)out)[t] = oracle_myers_semiglobal(q, qlen, s, slen); break;
            case ORACLE_BANDED_MYERS:     ((int8_t *)out)[t]  = oracle_banded_myers(q, qlen, s, slen, e); break;
            case ORACLE_BITPAL_PACKED:    ((int16_t *)out)[t] = oracle_bitpal_packed(q, qlen, s, slen, M, I, G); break;
            case ORACLE_BITPAL_NONPACKED: ((int16_t *)out)[t] = oracle_bitpal_nonpacked(q, qlen, s, slen, M, I, G); break;
            case ORACLE_BITPAL_PACKED_SEMIGLOBAL:
                ((int16_t *)out)[t] = oracle_bitpal_packed_semiglobal(q, qlen, s, slen, M, I, G); break;
        }
    }
    return 0;
}
