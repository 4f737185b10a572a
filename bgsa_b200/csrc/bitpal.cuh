// bitpal.cuh -- BitPAl (Loving, Hernandez, Benson) bit-parallel global alignment with general
// integer scores (match M, mismatch I, gap G), as tps_kernel policies.
//
// Replaces the generated align_mic / align_avx (original/BGSA_AVX512/align_core.c:19-485,
// original/BGSA_AVX2/align_core.c) and what BitPAlGenerator.java emits for other (M,I,G)
// (:151-534 packed, :1392-1701 non-packed).  Scoring schemes are template instances.
//
// The recurrence we implement (derived from the NW recurrence; the reference's generated code
// evaluates the same quantities).  With p the position along the bit-vector
// (here: the QUERY), one DP column per subject base, and after subtracting G from every delta:
//     d_p = (S[p][t] - S[p-1][t]) - G        state, in [0, A]          A = M - 2G
//     e_p = (S[p][t+1] - S[p][t]) - G        within the column, e_0 = 0 (global top row)
//     w_p = A on a match, B on a mismatch                               B = max(I - 2G, 0)
//     y_p = max(w_p, e_{p-1})
//     e_p = max(0, y_p - d_p)
//     T_p = max(y_p, d_p)                    (= diagonal difference - 2G)
//     d'_p = T_p - e_{p-1}
// The serial dependency e_{p-1} -> e_p only matters for the "high" classes e in (B, A]: they
// propagate unchanged through runs of d = 0 on mismatches, which one integer ADD per class
// resolves for 32 positions at once (carry propagation = run propagation), exactly BitPAl's trick.
//
// Semi-global (generator -s, BitPAlGenerator.java:77-80,112-114,289-308): the whole QUERY against the best
// substring of the SUBJECT -- S[0][t] = 0, S[p][0] = p*G, answer = max_t S[m][t].  Transposed: the top row's
// horizontal delta is 0, i.e. e_{-1} = -G enters every column as the boundary value, and the lane that holds
// the last query row adds up its horizontal deltas e_m + G column by column, keeping the maximum.
//
// "packed"     : d kept as NB = ceil(log2(A+1)) binary bit-planes (reference: two's complement of
//                -d in NB+1 planes); y, e, T as binary planes; bit-sliced subtract/compare.
// "non-packed" : d kept one-hot, one bit-vector per value 0..A (BitPAl's original formulation).
// Final score = G*(n+m) + sum_p d_p, times the common factor the scores were divided by
// (Main.java:213-267, BitPAlGenerator.java:124-130), low 32 bits narrowed to int16.
#pragma once

#include "align_kernel.cuh"

namespace bgsa {

constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
constexpr int cabs(int a) { return a < 0 ? -a : a; }
constexpr int cbits(int v) { int b = 0; while ((1 << b) < v + 1) b++; return b; }   // bits to hold 0..v

// Main.java:213-238 picks the largest i <= min(|scores|) dividing all three (0 match is skipped).
constexpr int common_factor(int M, int I, int G) {
    int m = cabs(I), g = cabs(G);
    int mn = M == 0 ? m : M;
    if (m < mn) mn = m;
    if (g < mn) mn = g;
    int f = 1;
    for (int i = 2; i <= mn; i++) if (M % i == 0 && m % i == 0 && g % i == 0) f = i;
    return f;
}

template <int M_, int I_, int G_>
struct Scheme {
    static constexpr int F = common_factor(M_, I_, G_);
    static constexpr int M = M_ / F, I = I_ / F, G = G_ / F;
    static constexpr int A = M - 2 * G;
    static constexpr int B = (I - 2 * G) > 0 ? (I - 2 * G) : 0;
    static constexpr int NH = A - B;          // number of high classes B+1 .. A
    static constexpr int NB = cbits(A);       // binary planes for 0..A
    static_assert(G_ < 0 && M_ > I_ && M_ >= 0 && A >= 1 && B < A, "unsupported scoring scheme");
};

struct BitpalParams { int dummy; };

// ---------------------------------------------------------------------------------------------
// packed
// ---------------------------------------------------------------------------------------------
enum { BITPAL_GLOBAL = 0, BITPAL_SEMIGLOBAL = 1 };

template <class S, int K_, int MODE = BITPAL_GLOBAL>
struct BitpalPacked {
    static constexpr int K = K_;
    static constexpr int A = S::A, B = S::B, NB = S::NB, NH = S::NH;
    static constexpr bool SEMI = MODE == BITPAL_SEMIGLOBAL;
    static constexpr int kOpsPerWord = 65;
    static constexpr int kMinBlocksWavefront = BGSA_MIN_BLOCKS;
    static constexpr int E0 = SEMI ? -S::G : 0;              // e_{-1}: horizontal delta of the top row, minus G
    using Params = BitpalParams;
    // semi-global only: tmask[j] = the bit of word j that is the last query row (0: not in this word),
    // tbit its position, cur/best = running S[m][t] - S[m][0] and its maximum
    struct State { uint32_t d[NB][K]; uint32_t tmask[SEMI ? K : 1]; int cur, best, tbit; bool owner; };

    static BGSA_HD void init(State &s, int first_bit, int qlen) {
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int j = 0; j < K; j++) s.d[b][j] = 0u;      // first column: every vertical delta = G
        const int last = qlen - 1 - first_bit;
        s.owner = SEMI && last >= 0 && last < 32 * K;
        s.tbit = last & 31;
        if (SEMI) {
#pragma unroll
            for (int j = 0; j < K; j++) s.tmask[j] = (s.owner && (last >> 5) == j) ? (1u << s.tbit) : 0u;
        }
        s.cur = 0; s.best = 0;
    }

    // carry stream (CarryIn/CarryOut, consumption order): add carry of class A; then for every lower
    // high class, top down: shift-in bit of its init vector, its add carry; then the NB shift-in
    // bits of the e planes.
    static constexpr int kCarryBits = NH + (NH > 0 ? NH - 1 : 0) + NB;
    // what the top row hands to lane 0: e_{-1} = E0 as a carry stream (a high class enters through the shift-in
    // bit of its init vector, every value through the e planes); global: E0 = 0, nothing propagates in
    static __host__ __device__ constexpr bool init_in(int c) { return E0 > 0 && B + 1 + c == E0; }
    static constexpr uint32_t boundary_word() {
        uint32_t w = 0u;
        if (NH > 0) w = (w << 1);                            // add carry of class A
        for (int c = NH - 2; c >= 0; c--) { w = (w << 1) | (init_in(c) ? 1u : 0u); w = (w << 1); }
        for (int b = 0; b < NB; b++) w = (w << 1) | ((E0 >> b) & 1);
        return w << (32 - kCarryBits);
    }
    static constexpr uint32_t kBoundary = boundary_word();
    static_assert(E0 <= A && E0 != A, "boundary value must be representable below class A");

    template <bool CARRY>
    static BGSA_HD uint32_t column(State &s, const uint32_t *row, uint32_t cin) {
        CarryIn in(cin);
        CarryOut out;
        uint32_t eq[(K + 3) / 4 * 4];
#pragma unroll
        for (int j = 0; j < (K + 3) / 4; j++) {
            const uint4 v = reinterpret_cast<const uint4 *>(row)[j];
            eq[4 * j] = v.x; eq[4 * j + 1] = v.y; eq[4 * j + 2] = v.z; eq[4 * j + 3] = v.w;
        }
        // ---- decode the d classes the chains need: Z = [d == 0], D[v] = [d == v], v = 1..NH-1.
        // Values below 4 share the test of the planes above bit 1 (one LOP3 each after that).
        uint32_t Z[K], remain[K];
        uint32_t D[NH > 1 ? NH : 1][K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            // hin = some plane above bit 1 is set  <=>  d >= 4.  Kept in this polarity: the tables below take its complement
            // for free, whereas "all zero" would cost a NOT that the explicit tables keep the compiler from folding away
            // (one LOP3 per word-column; profiles/r02_sass).
            uint32_t hin = 0u;
#pragma unroll
            for (int b = 2; b < NB; b++) hin |= s.d[b][j];
            constexpr int NA = 0xFF ^ LA;
            const uint32_t d0 = s.d[0][j], d1 = NB > 1 ? s.d[1][j] : 0u;
            Z[j] = lop3<NA & (0xFF ^ LB) & (0xFF ^ LC)>(hin, d1, d0);
            remain[j] = Z[j] & ~eq[j];                       // d == 0 and mismatch: runs that propagate
#pragma unroll
            for (int v = 1; v < NH; v++) {
                uint32_t m;
                if (v == 1) m = lop3<NA & (0xFF ^ LB) & LC>(hin, d1, d0);
                else if (v == 2) m = lop3<NA & LB & (0xFF ^ LC)>(hin, d1, d0);
                else if (v == 3) m = lop3<NA & LB & LC>(hin, d1, d0);
                else {
                    m = 0xffffffffu;
#pragma unroll
                    for (int b = NB - 1; b >= 0; b--) m &= ((v >> b) & 1) ? s.d[b][j] : ~s.d[b][j];
                }
                D[v][j] = m;
            }
        }
        // ---- high classes, top down.  Y[c][j] = [y_p == B+1+c]   (c = NH-1 is class A)
        uint32_t Y[NH > 0 ? NH : 1][K];
        if (NH > 0) {
            uint32_t a0[K], sum[K];
#pragma unroll
            for (int j = 0; j < K; j++) a0[j] = Z[j] & eq[j];
            if (CARRY) in.to_cf();
            add_chain<K, CARRY>(sum, a0, Z);
            if (CARRY) out.push_cf();
#pragma unroll
            for (int j = 0; j < K; j++) Y[NH - 1][j] = lop3<(LA ^ LB) | LC>(sum[j], remain[j], eq[j]);   // e_{p-1} == A, or match
#pragma unroll
            for (int c = NH - 2; c >= 0; c--) {               // class k = B+1+c
                uint32_t init[K], sh[K];
#pragma unroll
                for (int j = 0; j < K; j++) {
                    uint32_t v = D[1][j] & Y[c + 1][j];
#pragma unroll
                    for (int dl = 2; c + dl <= NH - 1; dl++) v = lop3<LA | (LB & LC)>(v, D[dl][j], Y[c + dl][j]);
                    init[j] = v;                              // e_p == k at a position with d != 0
                }
                const uint32_t sin = CARRY ? in.top() : (init_in(c) ? 0x80000000u : 0u);
#pragma unroll
                for (int j = 0; j < K; j++)
                    sh[j] = (j == 0 && !CARRY) ? (init_in(c) ? shl1_const<1>(init[0]) : shl1_const<0>(init[0]))
                                               : shl1_carry(j ? init[j - 1] : sin, init[j]);
                if (CARRY) { out.push_top(init[K - 1]); in.to_cf(); }
                add_chain<K, CARRY>(sum, sh, remain);
                if (CARRY) out.push_cf();
#pragma unroll
                for (int j = 0; j < K; j++) Y[c][j] = lop3<(LA ^ LB) & (0xFF ^ LC)>(sum[j], remain[j], eq[j]);
            }
        }
        // ---- binary planes of y, then e = max(0, y - d), T = max(y, d), d' = T - (e << 1)
        uint32_t e_prev[NB], e_last[NB];                     // e_last: the e planes at the last query row (one bit)
#pragma unroll
        for (int b = 0; b < NB; b++) e_last[b] = 0u;
#pragma unroll
        for (int b = 0; b < NB; b++) e_prev[b] = CARRY ? in.top() : (((E0 >> b) & 1) ? 0x80000000u : 0u);   // top row
#pragma unroll
        for (int j = 0; j < K; j++) {
            uint32_t hi = 0u;
#pragma unroll
            for (int c = 0; c < NH; c++) hi |= Y[c][j];
            uint32_t y[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) {
                uint32_t v = ((B >> b) & 1) ? ~hi : 0u;
#pragma unroll
                for (int c = 0; c < NH; c++) if (((B + 1 + c) >> b) & 1) v |= Y[c][j];
                y[b] = v;
            }
            // borrow chain of y - d
            // (explicit 3-input tables: left to itself the compiler materialises y ^ d and spends a third
            //  instruction per plane)
            constexpr int kBorrow = ((0xFF ^ LA) & LB) | ((0xFF ^ (LA ^ LB)) & LC);   // borrow out of a - b - c
            // (top plane: the difference bit is only ever used masked by "no final borrow", and that borrow is a function of
            //  the same three inputs -- one table gives the masked bit directly: one LOP3 less per word)
            uint32_t diff[NB], br = 0u, e_top = 0u;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const uint32_t yb = y[b], db = s.d[b][j];
                if (b == 0) { diff[b] = yb ^ db; br = ~yb & db; }
                else if (b == NB - 1) { e_top = lop3<(LA ^ LB ^ LC) & (0xFF ^ kBorrow)>(yb, db, br); diff[b] = 0u; br = lop3<kBorrow>(yb, db, br); }
                else { diff[b] = lop3<LA ^ LB ^ LC>(yb, db, br); br = lop3<kBorrow>(yb, db, br); }
            }
            const uint32_t lt = br;                           // y < d
            uint32_t e[NB], T[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) {
                e[b] = (b == NB - 1 && b > 0) ? e_top : (diff[b] & ~lt);
                T[b] = lop3<(LA & LB) | ((0xFF ^ LA) & LC)>(lt, s.d[b][j], y[b]);
            }
            if (SEMI) {                                      // branch-free: one LOP3 per plane and word
#pragma unroll
                for (int b = 0; b < NB; b++) e_last[b] = lop3<LA | (LB & LC)>(e_last[b], e[b], s.tmask[SEMI ? j : 0]);
            }
            // shift e one position up (e_{p-1} aligned with p), then d' = T - es
            uint32_t br2 = 0u;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const uint32_t es = (j == 0 && !CARRY) ? (((E0 >> b) & 1) ? shl1_const<1>(e[b]) : shl1_const<0>(e[b]))
                                                       : shl1_carry(e_prev[b], e[b]);
                e_prev[b] = e[b];
                const uint32_t tb = T[b];
                if (b == 0) { s.d[b][j] = tb ^ es; br2 = ~tb & es; }
                else { s.d[b][j] = lop3<LA ^ LB ^ LC>(tb, es, br2); if (b + 1 < NB) br2 = lop3<kBorrow>(tb, es, br2); }
            }
        }
        if (SEMI) {                                          // horizontal delta of the last query row = e + G
            int v = 0;
#pragma unroll
            for (int b = 0; b < NB; b++) v += (int)(e_last[b] >> s.tbit) << b;
            s.cur += v + S::G;                               // (lanes that do not own the row add G to a value nobody reads)
            s.best = s.cur > s.best ? s.cur : s.best;
        }
        if (CARRY) {
#pragma unroll
            for (int b = 0; b < NB; b++) out.push_top(e_prev[b]);
        }
        return CARRY ? out.template finish<kCarryBits>() : 0u;
    }

    static BGSA_HD Partial partial(const State &s, int first_bit, int qlen) {
        Partial r; r.sum = 0; r.minpre = 0;
        if (SEMI) {                                          // only the lane with the last query row has the answer
            r.minpre = s.owner ? -s.best : 0x3fffffff;
        } else {
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int rem = qlen - first_bit - 32 * j;
                const uint32_t mask = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
#pragma unroll
                for (int b = 0; b < NB; b++) r.sum += popc32(s.d[b][j] & mask) << b;
            }
        }
        return r;
    }
    static BGSA_HD int final_score(int sum, int minpre, int qlen, int slen, Params) {
        if (SEMI) return (S::G * qlen - minpre) * S::F;      // S[m][0] + max_t (S[m][t] - S[m][0])
        return (S::G * (qlen + slen) + sum) * S::F;
    }
};

// ---------------------------------------------------------------------------------------------
// non-packed: one bit-vector per value of d (BitPAlGenerator.java:1392-1701 keeps dh_min..dh_max;
// we keep the A vectors for d = 1..A, d == 0 being their complement).
// ---------------------------------------------------------------------------------------------
template <class S, int K_>
struct BitpalNonPacked {
    static constexpr int K = K_;
    static constexpr int A = S::A, B = S::B, NH = S::NH;
    static constexpr int kOpsPerWord = 185;
    static constexpr int kMinBlocksWavefront = BGSA_MIN_BLOCKS;
    using Params = BitpalParams;
    struct State { uint32_t d[A][K]; };        // d[v-1][j] = [d == v]

    static BGSA_HD void init(State &s, int, int) {
#pragma unroll
        for (int v = 0; v < A; v++)
#pragma unroll
            for (int j = 0; j < K; j++) s.d[v][j] = 0u;
    }

    // carry stream (consumption order): add carry of class A; for k = A-1 .. B+1: shift-in bit of
    // init_k, add carry of class k; for k = B .. 1: shift-in bit of init_k.
    static constexpr int kCarryBits = NH + (A - 1);
    static constexpr uint32_t kBoundary = 0u;

    template <bool CARRY>
    static BGSA_HD uint32_t column(State &s, const uint32_t *row, uint32_t cin) {
        CarryIn in(cin);
        CarryOut out;
        uint32_t eq[(K + 3) / 4 * 4];
#pragma unroll
        for (int j = 0; j < (K + 3) / 4; j++) {
            const uint4 v = reinterpret_cast<const uint4 *>(row)[j];
            eq[4 * j] = v.x; eq[4 * j + 1] = v.y; eq[4 * j + 2] = v.z; eq[4 * j + 3] = v.w;
        }
        // dhi = [d_p > B] is needed twice (for Z here and for max(w, d) == B at the end): OR it once
        uint32_t Z[K], remain[K], dhi[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            uint32_t hi = 0u, lo = 0u;
#pragma unroll
            for (int v = B; v < A; v++) hi |= s.d[v][j];
#pragma unroll
            for (int v = 0; v < B; v++) lo |= s.d[v][j];
            dhi[j] = hi;
            Z[j] = ~(hi | lo);
            remain[j] = Z[j] & ~eq[j];
        }
        // DV(v, j) = [d_p == v]
#define DV(v, j) (((v) == 0) ? Z[j] : s.d[((v) > 0 ? (v) : 1) - 1][j])
        // X[v][j] = [e_{p-1} == v] for v = 1..A (shifted form), built class by class.
        // Yh[v][j] = [y_p == v]: v == A -> X[A] | match ; B < v < A -> X[v] & ~match ; v == B -> rest.
        // (every OR-of-ANDs below is written as one LOP3 per term: a | (b & c))
        uint32_t X[A + 1][K];
        uint32_t Yh[A + 1][K];
        {
            uint32_t a0[K], sum[K];
#pragma unroll
            for (int j = 0; j < K; j++) a0[j] = Z[j] & eq[j];
            if (CARRY) in.to_cf();
            add_chain<K, CARRY>(sum, a0, Z);
            if (CARRY) out.push_cf();
#pragma unroll
            for (int j = 0; j < K; j++) {
                X[A][j] = sum[j] ^ remain[j];
                Yh[A][j] = lop3<(LA ^ LB) | LC>(sum[j], remain[j], eq[j]);
            }
#pragma unroll
            for (int k = A - 1; k > B; k--) {
                uint32_t init[K], sh[K];
#pragma unroll
                for (int j = 0; j < K; j++) {
                    uint32_t v = DV(A - k, j) & Yh[A][j];
#pragma unroll
                    for (int h = A - 1; h > k; h--) v = lop3<LA | (LB & LC)>(v, DV(h - k, j), Yh[h][j]);
                    init[j] = v;
                }
                const uint32_t sin = CARRY ? in.top() : 0u;
#pragma unroll
                for (int j = 0; j < K; j++) sh[j] = (j == 0 && !CARRY) ? shl1_const<0>(init[0]) : shl1_carry(j ? init[j - 1] : sin, init[j]);
                if (CARRY) { out.push_top(init[K - 1]); in.to_cf(); }
                add_chain<K, CARRY>(sum, sh, remain);
                if (CARRY) out.push_cf();
#pragma unroll
                for (int j = 0; j < K; j++) {
                    X[k][j] = sum[j] ^ remain[j];
                    Yh[k][j] = lop3<(LA ^ LB) & (0xFF ^ LC)>(sum[j], remain[j], eq[j]);
                }
            }
        }
        // (rest, X0 and max(w,d) == B below are complements of ORs that feed explicit LOP3 tables: they are kept as the ORs --
        //  nrest, nX0, nmxB -- and the tables take the complement, which saves the three NOTs per word that the compiler
        //  cannot fold across the asm statements)
        uint32_t nrest[K], xhi[K];                            // xhi = [B < e_{p-1} < A], shared with nX0 below
#pragma unroll
        for (int j = 0; j < K; j++) {
            uint32_t any = 0u;
#pragma unroll
            for (int k = A - 1; k > B; k--) any |= X[k][j];
            xhi[j] = any;
            nrest[j] = any | Yh[A][j];
        }
        // low classes 1..B: e_p == k  <=>  y_p - d_p == k ; no propagation, plain shift
#pragma unroll
        for (int k = B; k >= 1; k--) {
            uint32_t init[K];
#pragma unroll
            for (int j = 0; j < K; j++) {
                uint32_t v = DV(A - k, j) & Yh[A][j];
#pragma unroll
                for (int h = A - 1; h > B; h--) v = lop3<LA | (LB & LC)>(v, DV(h - k, j), Yh[h][j]);
                init[j] = lop3<LA | (LB & (0xFF ^ LC))>(v, DV(B - k, j), nrest[j]);
            }
            const uint32_t sin = CARRY ? in.top() : 0u;
            if (CARRY) out.push_top(init[K - 1]);
#pragma unroll
            for (int j = 0; j < K; j++) X[k][j] = (j == 0 && !CARRY) ? shl1_const<0>(init[0]) : shl1_carry(j ? init[j - 1] : sin, init[j]);
        }
#undef DV
        // X0 = [e_{p-1} == 0]
        uint32_t nX0[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            uint32_t any = xhi[j] | X[A][j];
#pragma unroll
            for (int k = 1; k <= B; k++) any |= X[k][j];
            nX0[j] = any;
        }
        // Mx(v) = [max(w, d) == v]: v == A -> d==A | match ; B < v < A -> d==v & ~match ; v == B -> others
        // d'_k = OR_{m >= max(k,B)} Mx(m) & [e_{p-1} == m - k]      (k >= 1)
#pragma unroll
        for (int j = 0; j < K; j++) {
            uint32_t mx[A + 1];
            mx[A] = s.d[A - 1][j] | eq[j];
#pragma unroll
            for (int v = A - 1; v > B; v--) mx[v] = s.d[v - 1][j] & ~eq[j];
            mx[B] = dhi[j] | eq[j];                          // COMPLEMENT of [max(w,d) == B]: a match or d > B
            uint32_t nd[A];
#pragma unroll
            for (int k = 1; k <= A; k++) {
                uint32_t v = 0u;
#pragma unroll
                for (int m = (k > B ? k : B); m <= A; m++)
                {
                    const uint32_t xv = (m - k) == 0 ? nX0[j] : X[(m - k) > 0 ? (m - k) : 1][j];
                    // v | (mx & x), with mx complemented for m == B and x complemented for m == k
                    if (m == B && m == k) v = lop3<LA | ((0xFF ^ LB) & (0xFF ^ LC))>(v, mx[m], xv);
                    else if (m == B) v = lop3<LA | ((0xFF ^ LB) & LC)>(v, mx[m], xv);
                    else if (m == k) v = lop3<LA | (LB & (0xFF ^ LC))>(v, mx[m], xv);
                    else v = lop3<LA | (LB & LC)>(v, mx[m], xv);
                }
                nd[k - 1] = v;
            }
#pragma unroll
            for (int k = 0; k < A; k++) s.d[k][j] = nd[k];
        }
        return CARRY ? out.template finish<kCarryBits>() : 0u;
    }

    static BGSA_HD Partial partial(const State &s, int first_bit, int qlen) {
        Partial r; r.sum = 0; r.minpre = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int rem = qlen - first_bit - 32 * j;
            const uint32_t mask = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
#pragma unroll
            for (int v = 1; v <= A; v++) r.sum += v * popc32(s.d[v - 1][j] & mask);
        }
        return r;
    }
    static BGSA_HD int final_score(int sum, int, int qlen, int slen, Params) {
        return (S::G * (qlen + slen) + sum) * S::F;
    }
};

}  // namespace bgsa
