// pipe_rates.cu -- issue-rate microbenchmark for the FP32 pipes of sm_100a (dev tool, not product).
// One CTA of 1024 threads per SM (8 warps per scheduler), 8 independent dependency chains per thread.
// Prints warp-instructions per clock per scheduler for scalar FFMA / FMUL / FADD, the packed
// FFMA2 / FADD2 / FMUL2 (fma.rn.f32x2 ...), and mixes with integer / MUFU work.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 2048
#define CH 8
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, float seed) {
    float a[CH], b = seed, c = seed * 0.5f;
    u64 p[CH];
    int n[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); n[i] = threadIdx.x + i; }
    u64 bb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), cc = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) a[i] = __fmaf_rn(a[i], b, c);
            if (MODE == 1) a[i] = __fmul_rn(a[i], b);
            if (MODE == 2) a[i] = __fadd_rn(a[i], c);
            if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
            if (MODE == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
            if (MODE == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(bb));
            if (MODE == 6) { a[i] = __fmaf_rn(a[i], b, c); n[i] = (n[i] ^ it) + i; }        // FFMA + 2 ALU (LOP3, IADD3)
            if (MODE == 7) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc)); n[i] = (n[i] ^ it) + i; }
            if (MODE == 8) { a[i] = __fmaf_rn(a[i], b, c); a[i] = fminf(a[i], 3.0f); }     // FFMA + FMNMX
            if (MODE == 9) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (MODE == 10) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (MODE == 11) { a[i] = __fmaf_rn(a[i], b, c); a[i] = (n[i] & 1) ? a[i] : c; } // FFMA + select
        }
    }
    long long t1 = clock64();
    float s = 0; int ns = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) { s += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]); ns += n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + ns;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int instr_per_chain_step, float* out, long long* cyc, int sms) {
    k<MODE><<<sms, 1024>>>(out, cyc, 1.0001f);
    k<MODE><<<sms, 1024>>>(out, cyc, 1.0001f);
    cudaDeviceSynchronize();
    long long h[512];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    double winstr_per_sched = 8.0 * ITERS * CH * instr_per_chain_step;   // 8 warps per scheduler
    printf("%-28s %8.0f cycles  %.3f warp-instr/clk/scheduler (counting %d instr per step)\n", name, avg,
           winstr_per_sched / avg, instr_per_chain_step);
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms);
    printf("%s, %d SMs\n", pr.name, sms);
    run<0>("FFMA", 1, out, cyc, sms);
    run<1>("FMUL", 1, out, cyc, sms);
    run<2>("FADD", 1, out, cyc, sms);
    run<3>("FFMA2 (fma.rn.f32x2)", 1, out, cyc, sms);
    run<4>("FADD2", 1, out, cyc, sms);
    run<5>("FMUL2", 1, out, cyc, sms);
    run<6>("FFMA + LOP3 + IADD3", 3, out, cyc, sms);
    run<7>("FFMA2 + LOP3 + IADD3", 3, out, cyc, sms);
    run<8>("FFMA + FMNMX", 2, out, cyc, sms);
    run<9>("MUFU.EX2", 1, out, cyc, sms);
    run<10>("MUFU.RCP", 1, out, cyc, sms);
    run<11>("FFMA + LOP + SEL", 3, out, cyc, sms);
    return 0;
}
