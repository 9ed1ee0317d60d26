// Exhaustive proof tool (measurement/test infrastructure): for EVERY f32 bit pattern d and every
// quantiser q in 1..255, the 2-instruction sequence  x = fma(d, rq_hi, d * rq_lo)
// (rq_hi = fl(1/q), rq_lo = fl(1/q - rq_hi)) is compared with the IEEE division d / (float)q that
// the reference performs (quantizer.rs:60), both as floats and after round-half-away + i16 saturation.
// Also checks the normalisation v / max for every v <= max, max in 1..65535.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o exhaustive_div tools/exhaustive_div.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int rha_i16(float x) {
    const float h = __int_as_float((__float_as_int(x) & 0x80000000) | 0x3EFFFFFF);
    short s;
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(s) : "f"(__fadd_rn(x, h)));
    return (int)s;
}
__device__ __forceinline__ int rha_ref(float x) {  // roundf + saturating cast (Rust `as i16`)
    float r = roundf(x);
    if (isnan(r)) return 0;
    if (r > 32767.f) return 32767;
    if (r < -32768.f) return -32768;
    return (int)r;
}

__global__ void check_q(int q, float rq_hi, float rq_lo, unsigned long long* bad_float, unsigned long long* bad_int) {
    unsigned long long bf = 0, bi = 0;
    const float qf = (float)q;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32);
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float d = __uint_as_float((unsigned)i);
        const float ref = __fdiv_rn(d, qf);
        const float x = __fmaf_rn(d, rq_hi, __fmul_rn(d, rq_lo));
        if (__float_as_uint(ref) != __float_as_uint(x) && !(isnan(ref) && isnan(x))) bf++;
        if (rha_ref(ref) != rha_i16(x)) bi++;
    }
    if (bf) atomicAdd(bad_float, bf);
    if (bi) atomicAdd(bad_int, bi);
}

int main() {
    unsigned long long *d_bad, h_bad[2];
    cudaMalloc(&d_bad, 16);
    unsigned long long tot_f = 0, tot_i = 0;
    for (int q = 1; q <= 255; q++) {
        const double r = 1.0 / q;
        const float hi = (float)r, lo = (float)(r - (double)hi);
        cudaMemset(d_bad, 0, 16);
        check_q<<<148 * 16, 256>>>(q, hi, lo, d_bad, d_bad + 1);
        cudaMemcpy(h_bad, d_bad, 16, cudaMemcpyDeviceToHost);
        if (h_bad[0] || h_bad[1]) printf("q=%3d float mismatches %llu, quantised mismatches %llu\n", q, h_bad[0], h_bad[1]);
        tot_f += h_bad[0], tot_i += h_bad[1];
    }
    printf("division: total float mismatches %llu, total quantised-i16 mismatches %llu over 255 x 2^32 cases\n", tot_f, tot_i);
    // normalisation on the host (tiny)
    unsigned long long nb = 0;
    int worst = 0;
    for (int mx = 1; mx <= 65535; mx++) {
        const double r = 1.0 / mx;
        const float hi = (float)r, lo = (float)(r - (double)hi);
        for (int v = 0; v <= mx; v++) {
            const float ref = (float)v / (float)mx;
            const float x = fmaf((float)v, hi, (float)v * lo);
            if (ref != x) nb++, worst = mx;
        }
    }
    printf("normalisation v/max, all max in 1..65535, all v<=max: mismatches %llu (last bad max %d)\n", nb, worst);
    return (tot_i || nb) ? 1 : 0;
}
