// Micro-benchmark of the sm_100a issue/pipe rates that bound K1 (measurement tool, not product):
// scalar vs packed (f32x2) FP32 add/mul/fma, I2F.U8, F2I, LOP3, and an FADD+FMUL mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes tools/ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float seed, unsigned* iout) {
    float a[CHAINS], b = seed, c = seed * 0.5f;
    unsigned long long p[CHAINS], pb, pc;
    unsigned u[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        a[i] = seed + i + threadIdx.x;
        u[i] = threadIdx.x * 2654435761u + i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 1.0f));
    }
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) a[i] = __fmaf_rn(a[i], b, c);
            if (MODE == 1) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
            if (MODE == 2) a[i] = __fadd_rn(a[i], b);
            if (MODE == 3) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (MODE == 4) a[i] = __fmul_rn(a[i], b);
            if (MODE == 5) asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (MODE == 6) { if (i & 1) a[i] = __fadd_rn(a[i], b); else a[i] = __fmul_rn(a[i], b); }
            if (MODE == 7) a[i] = (float)((u[i] >> 8) & 0xFFu) + a[i];          // I2F.U8 + FADD
            if (MODE == 8) { short s; asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(s) : "f"(a[i])); u[i] += (unsigned)(int)s; a[i] = __fadd_rn(a[i], b); }  // F2I + IADD + FADD
            if (MODE == 9) u[i] = (u[i] & 0x80000000u) | (u[i] >> 1) | 0x3EFFFFFFu;  // LOP3/SHF
            if (MODE == 10) { if (i & 1) a[i] = __fadd_rn(a[i], b); else asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); }
        }
    }
    float s = 0;
    unsigned us = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        s += a[i] + lo + hi;
        us += u[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = us;
}

template <int MODE>
void run(const char* name, double ops_per_inst, float* d, unsigned* di) {
    int sms = 148, ctas = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    bench<MODE><<<ctas, 256>>>(d, 1.0001f, di);
    cudaEventRecord(e0);
    bench<MODE><<<ctas, 256>>>(d, 1.0001f, di);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_inst = (double)ctas * 8 * ITERS * CHAINS;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SM  (%.1f f32-lane-ops/clk/SM at max clock %d MHz)\n", name, ms,
           warp_inst / cycles / sms, warp_inst / cycles / sms * 32 * ops_per_inst, clk / 1000);
}

int main() {
    float* d;
    unsigned* di;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaMalloc(&di, 148 * 8 * 256 * 4);
    run<0>("FFMA", 1, d, di);
    run<1>("FFMA2 (fma.rn.f32x2)", 2, d, di);
    run<2>("FADD", 1, d, di);
    run<3>("FADD2 (add.rn.f32x2)", 2, d, di);
    run<4>("FMUL", 1, d, di);
    run<5>("FMUL2 (mul.rn.f32x2)", 2, d, di);
    run<6>("FADD+FMUL alternating", 1, d, di);
    run<7>("I2F.U8 + FADD", 1, d, di);
    run<8>("F2I.S16 + IADD + FADD", 1, d, di);
    run<9>("LOP3/SHF mix (3 ops)", 1, d, di);
    run<10>("FADD + FADD2 alternating", 1.5, d, di);
    return 0;
}
