// instances.h -- the kernel instances compiled into libbgsa_b200.so.
//
// The reference generates one align_core.c per (algorithm, scoring scheme, SIMD width)
// (generator/.../Main.java:240-315).  Here every combination is a C++ template instance selected
// at run time; adding a scoring scheme or a geometry is one line in the tables below.
// X(K, L): K 32-bit words per lane, L lanes per subject; serves queries up to 32*K*L bases.
// Among the instances that fit a query the cheapest is used (api.cu pick(): L * (K * ops per word
// + per-step wavefront overhead)): wide lanes amortise the carry hand-over, few lanes waste less
// of the last word.  (16, 2) is never the cheapest -- (32, 1) serves the same queries without the hand-over; it exists for
// the like-for-like measurement of the wavefront against thread-per-subject on C4 (BGSA_FORCE_KL=16,2).
#pragma once

// Myers global / semi-global
#define BGSA_MYERS_INSTANCES(X)                                                              \
    X(1, 1) X(2, 1) X(3, 1) X(4, 1) X(5, 1) X(6, 1) X(7, 1) X(8, 1) X(10, 1) X(12, 1) X(16, 1) \
    X(20, 1) X(24, 1) X(32, 1) X(16, 2) X(24, 2) X(32, 2) X(24, 4) X(32, 4) X(20, 8) X(24, 8) X(32, 8)  \
    X(24, 16) X(32, 16) X(24, 32) X(32, 32)

// BitPAl packed (global and semi-global)
#define BGSA_BITPAL_PACKED_INSTANCES(X)                                                      \
    X(1, 1) X(2, 1) X(3, 1) X(4, 1) X(5, 1) X(6, 1) X(8, 1) X(10, 1) X(6, 2) X(10, 2)         \
    X(8, 4) X(10, 4) X(8, 8) X(10, 8) X(8, 16) X(10, 16) X(8, 32) X(10, 32)

// BitPAl non-packed (one vector per delta value: register hungry, so few words per lane)
#define BGSA_BITPAL_NONPACKED_INSTANCES(X)                                                   \
    X(1, 1) X(2, 1) X(3, 1) X(4, 1) X(5, 1) X(3, 2) X(2, 2) X(2, 4) X(2, 8) X(2, 16) X(2, 32) X(3, 32) X(5, 32)

// Scoring schemes (match, mismatch, gap) with a kernel instance; index = scheme id.
// Scheme 0 is the one the reference checks in (original/BGSA_AVX512/align_core.c:13-15).
// The Makefile passes the list (make SCHEMES="2,-3,-5 ..."); this is its default.
#ifndef BGSA_SCHEMES
#define BGSA_SCHEMES(X) X(0, 2, -3, -5) X(1, 1, -1, -1) X(2, 1, -3, -2)
#endif
