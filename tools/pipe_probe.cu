// pipe_probe.cu -- issue-rate microbenchmark of the integer instruction mixes the alignment kernels are made of.
//
// Answers, on the box, the questions DESIGN.md section 3 "ALU pipe" asks before a formulation is adopted:
// how many warp-instructions per clock and SM sub-partition does each of these sustain, alone and mixed with LOP3:
//   LOP3 / SHF (ALU pipe), IMAD, IMAD.SHL, IMAD.WIDE.U32, IMAD.HI.U32, IADD3.X chains.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipe_probe tools/pipe_probe.cu
// Output: one line per mix: warp-instructions / clk / SMSP (total and per class).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHAINS 8

enum Mix {
    LOP3_ONLY, SHF_ONLY, IMAD_ONLY, IMADSHL_ONLY, IMADWIDE_ONLY, IMADHI_ONLY, ADDX_ONLY,
    LOP3_IMAD_1_1, LOP3_IMAD_2_1, LOP3_IMAD_4_1, LOP3_WIDE_2_1, LOP3_WIDE_4_1, LOP3_WIDE_8_1, LOP3_WIDE_MOV_4_1, SHF_IMAD_1_1,
    LOP3_ADDX_4_1, NMIX
};
static const char *kNames[NMIX] = {
    "lop3", "shf.l.w", "imad (3 reg)", "imad.shl (x*2)", "imad.wide + imad pairs", "imad.hi.u32", "iadd3.x chain",
    "lop3:imad 1:1", "lop3:imad 2:1", "lop3:imad 4:1", "lop3:imad.wide 2:1", "lop3:imad.wide 4:1", "lop3:imad.wide 8:1",
    "lop3:(wide+imad) 8:2x2", "shf:imad 1:1", "lop3:iadd3.x 4:1"};
// instructions per chain and inner trip: {alu, fma}
static const int kCount[NMIX][2] = {{8, 0}, {8, 0}, {0, 8}, {0, 8}, {0, 8}, {0, 8}, {8, 0}, {4, 4}, {8, 4}, {8, 2}, {8, 4}, {8, 2},
                                    {8, 1}, {8, 4}, {4, 4}, {10, 0}};   // (approximate where ptxas adds moves: read the SASS)

__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t shf(uint32_t lo, uint32_t hi) {
    uint32_t d; asm volatile("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(d) : "r"(lo), "r"(hi)); return d;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
// the multiplier 2 is a RUN-TIME value (kernel parameter): with an immediate ptxas strength-reduces the multiply to
// LEA / LEA.HI / SHF, which are ALU-pipe instructions -- the pipe we want to unload
__device__ uint32_t g_two_dummy;
#define TWO two
__device__ __forceinline__ uint32_t imadshl_(uint32_t a, uint32_t c, uint32_t two) {
    uint32_t d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(two), "r"(c)); return d;
}
__device__ __forceinline__ uint64_t imadwide_(uint32_t a, uint64_t c, uint32_t two) {
    uint64_t d; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(two), "l"(c)); return d;
}
__device__ __forceinline__ uint32_t imadhi_(uint32_t a, uint32_t c, uint32_t two) {
    uint32_t d; asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(two), "r"(c)); return d;
}
#define imadshl(a, c) imadshl_(a, c, two)
#define imadwide(a, c) imadwide_(a, c, two)
#define imadhi(a, c) imadhi_(a, c, two)

template <int MIX>
__global__ void __launch_bounds__(256) probe(int iters, uint32_t seed, uint32_t *sink, long long *clk, uint32_t two) {
    uint32_t a[CHAINS], b[CHAINS];
    uint64_t w[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { a[i] = seed * (threadIdx.x + 17 * i + 1); b[i] = a[i] ^ 0x9e3779b9u; w[i] = a[i]; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MIX == LOP3_ONLY) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[i] = lop(a[i], b[i], a[(i + 1) % CHAINS]);
            } else if (MIX == SHF_ONLY) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[i] = shf(a[i], b[i]);
            } else if (MIX == IMAD_ONLY) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[i] = imad(a[i], b[i], a[(i + 1) % CHAINS]);
            } else if (MIX == IMADSHL_ONLY) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[i] = imadshl(a[i], b[i]);
            } else if (MIX == IMADWIDE_ONLY) {
                // pairs of IMAD.WIDE (both halves used) + IMAD: 4 + 4 FMA-pipe instructions
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint64_t W = imadwide(a[i], 0ull);
                    a[i] = imadshl(b[i], (uint32_t)(W >> 32));
                    b[i] = (uint32_t)W;
                }
            } else if (MIX == IMADHI_ONLY) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[i] = imadhi(a[i], b[i]);
            } else if (MIX == ADDX_ONLY) {
                uint32_t t;
                asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(t) : "r"(a[i]), "r"(b[i]));
#pragma unroll
                for (int k = 0; k < 6; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(t) : "r"(b[i]));
                asm volatile("addc.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(t), "r"(b[i]));
            } else if (MIX == LOP3_IMAD_1_1) {
#pragma unroll
                for (int k = 0; k < 4; k++) { a[i] = lop(a[i], b[i], a[(i + 1) % CHAINS]); b[i] = imad(b[i], b[i], b[(i + 1) % CHAINS]); }
            } else if (MIX == LOP3_IMAD_2_1) {
#pragma unroll
                for (int k = 0; k < 4; k++) { a[i] = lop(a[i], b[i], a[(i + 1) % CHAINS]); a[i] = lop(a[i], b[i], a[(i + 2) % CHAINS]); b[i] = imad(b[i], b[i], b[(i + 1) % CHAINS]); }
            } else if (MIX == LOP3_IMAD_4_1) {
#pragma unroll
                for (int k = 0; k < 2; k++) {
#pragma unroll
                    for (int q = 0; q < 4; q++) a[i] = lop(a[i], b[i], a[(i + 1 + q) % CHAINS]);
                    b[i] = imad(b[i], b[i], b[(i + 1) % CHAINS]);
                }
            } else if (MIX == LOP3_WIDE_2_1) {
#pragma unroll
                for (int k = 0; k < 4; k++) { a[i] = lop(a[i], b[i], a[(i + 1) % CHAINS]); a[i] = lop(a[i], b[i], a[(i + 2) % CHAINS]); w[i] = imadwide((uint32_t)w[i], w[i] >> 32); }
            } else if (MIX == LOP3_WIDE_4_1) {
#pragma unroll
                for (int k = 0; k < 2; k++) {
#pragma unroll
                    for (int q = 0; q < 4; q++) a[i] = lop(a[i], b[i], a[(i + 1 + q) % CHAINS]);
                    w[i] = imadwide((uint32_t)w[i], w[i] >> 32);
                }
            } else if (MIX == LOP3_WIDE_8_1) {
#pragma unroll
                for (int q = 0; q < 8; q++) a[i] = lop(a[i], b[i], a[(i + 1 + q) % CHAINS]);
                w[i] = imadwide((uint32_t)w[i], w[i] >> 32);
            } else if (MIX == LOP3_WIDE_MOV_4_1) {
                // the multi-word one-position shift as the kernels would write it, FMA pipe only:
                //   {lo, hi} = w_j * 2 (IMAD.WIDE, hi = the bit that leaves the word);  out_j = w_j * 2 + hi_{j-1} (IMAD)
                // 8 LOP3 : 2 x (IMAD.WIDE + IMAD)  -- the Myers word-column with both shifts moved off the ALU pipe
#pragma unroll
                for (int q = 0; q < 6; q++) a[i] = lop(a[i], b[i], a[(i + 1 + q) % CHAINS]);
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const uint64_t W = imadwide(b[i], 0ull);
                    b[i] = imadshl(b[(i + 1) % CHAINS], (uint32_t)(W >> 32));
                    a[i] = lop(a[i], (uint32_t)W, b[i]);
                }
            } else if (MIX == SHF_IMAD_1_1) {
#pragma unroll
                for (int k = 0; k < 4; k++) { a[i] = shf(a[i], a[(i + 1) % CHAINS]); b[i] = imad(b[i], b[i], b[(i + 1) % CHAINS]); }
            } else if (MIX == LOP3_ADDX_4_1) {
                uint32_t t;
#pragma unroll
                for (int q = 0; q < 4; q++) a[i] = lop(a[i], b[i], a[(i + 1 + q) % CHAINS]);
                asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(t) : "r"(a[i]), "r"(b[i]));
                asm volatile("addc.u32 %0, %1, %2;" : "=r"(b[i]) : "r"(t), "r"(b[i]));
#pragma unroll
                for (int q = 0; q < 4; q++) a[i] = lop(a[i], b[i], a[(i + 1 + q) % CHAINS]);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int MIX>
void run(int sm, uint32_t *sink, long long *clk) {
    const int iters = 4000;
    const int ctas = sm * 4, threads = 256;      // 8 warps per SMSP
    probe<MIX><<<ctas, threads>>>(10, 3, sink, clk, 2u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MIX><<<ctas, threads>>>(iters, 3, sink, clk, 2u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cycles = 0;
    cudaMemcpy(&cycles, clk, sizeof(cycles), cudaMemcpyDeviceToHost);
    const double warps_per_smsp = (double)ctas * threads / 32 / sm / 4;
    const double alu = (double)kCount[MIX][0] * CHAINS * iters * warps_per_smsp / cycles;
    const double fma = (double)kCount[MIX][1] * CHAINS * iters * warps_per_smsp / cycles;
    printf("%-28s  %.3f ms  %9lld clk  warp-instr/clk/SMSP: total %.3f  alu-class %.3f  fma-class %.3f   (MHz %.0f)\n", kNames[MIX], ms,
           cycles, alu + fma, alu, fma, cycles / (ms * 1e3));
}

template <int MIX>
void run_all(int sm, uint32_t *sink, long long *clk) {
    run<MIX>(sm, sink, clk);
    if constexpr (MIX + 1 < NMIX) run_all<MIX + 1>(sm, sink, clk);
}

int main() {
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *sink; long long *clk;
    cudaMalloc(&sink, 64); cudaMalloc(&clk, 64);
    printf("SMs %d\n", sm);
    run_all<0>(sm, sink, clk);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
