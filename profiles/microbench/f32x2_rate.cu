// Microbenchmark: issue/pipe rate of packed fp32 (FMUL2/FFMA2/FADD2) against scalar FMUL/FFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/f32x2_rate f32x2_rate.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long pf2;
#define ITER 2048
template <int MODE> __global__ void __launch_bounds__(256) k(float *out, float seed, long long *cyc) {
    float a[16];
    pf2 p[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) { p[i] = ((pf2)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]); u[i] = threadIdx.x * 7 + i; }
    const float m = seed * 0.5f;
    const pf2 pm = ((pf2)__float_as_uint(m) << 32) | __float_as_uint(m);
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
        if (MODE == 0) {  // 16 scalar FMUL
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fmul_rn(a[i], m);
        } else if (MODE == 1) {  // 8 FMUL2 (same flops as mode 0)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm));
        } else if (MODE == 2) {  // 8 FFMA2
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pm));
        } else if (MODE == 3) {  // 8 FADD2
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm));
        } else if (MODE == 4) {  // 8 FMUL2 + 8 scalar FMUL
#pragma unroll
            for (int i = 0; i < 8; i++) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm)); a[i] = __fmul_rn(a[i], m); }
        } else if (MODE == 5) {  // 8 FMUL2 + 8 LOP3 (alu pipe)
#pragma unroll
            for (int i = 0; i < 8; i++) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(it)); }
        } else if (MODE == 6) {  // 16 scalar FMUL + 8 LOP3
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fmul_rn(a[i], m);
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(it));
        } else if (MODE == 7) {  // 16 scalar FADD
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fadd_rn(a[i], m);
        } else if (MODE == 8) {  // 16 LOP3 only
#pragma unroll
            for (int i = 0; i < 8; i++) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(it)); }
#pragma unroll
            for (int i = 0; i < 8; i++) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 3) & 7]), "r"(it)); }
        } else if (MODE == 9) {  // 8 FMUL2 + 8 MOV-ish (alu FMNMX)
#pragma unroll
            for (int i = 0; i < 8; i++) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm)); a[i] = fmaxf(a[i], a[(i + 1) & 15]); }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char *name, int insts_per_iter) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(float)); cudaMalloc(&cyc, 8);
    k<MODE><<<148 * 4, 256>>>(out, 1.0001f, cyc);  // 4 CTAs x 8 warps per SM = 8 warps per SMSP
    k<MODE><<<148 * 4, 256>>>(out, 1.0001f, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per SMSP: 8 warps x ITER x insts_per_iter instructions in h cycles (all CTAs co-resident)
    printf("%-28s %8lld cycles  %.3f warp-inst/cycle/SMSP\n", name, h, 8.0 * ITER * insts_per_iter / (double)h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("16 FMUL", 16);
    run<7>("16 FADD", 16);
    run<1>("8 FMUL2", 8);
    run<2>("8 FFMA2", 8);
    run<3>("8 FADD2", 8);
    run<4>("8 FMUL2 + 8 FMUL", 16);
    run<5>("8 FMUL2 + 8 LOP3", 16);
    run<6>("16 FMUL + 8 LOP3", 24);
    run<8>("16 LOP3", 16);
    run<9>("8 FMUL2 + 8 FMNMX", 16);
    return 0;
}
