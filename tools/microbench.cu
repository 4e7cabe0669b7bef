// Instruction-throughput microbenchmark for sm_100a (B200).
// Purpose: pick the arithmetic strategy for the FMA-free hydrodynamics kernel
// (scalar vs packed f32x2, cost of cvt.f64.f32, DADD, IEEE div/sqrt).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Output: one line per op: lane-ops per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, double* outd, long long* cycles, float seed)
{
    float a[CHAINS];
    double d[CHAINS];
    unsigned long long pk[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
        d[i] = (double)a[i];
        float lo = a[i], hi = a[i] + 0.5f;
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(pk[i]) : "f"(lo), "f"(hi));
    }
    float b = seed * 0.999f + 1.0f;
    double bd = (double)b;
    unsigned long long bpk;
    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(bpk) : "f"(b), "f"(b));
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pk[i]) : "l"(bpk));
            if (OP == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(pk[i]) : "l"(bpk));
            if (OP == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pk[i]) : "l"(bpk));
            if (OP == 6) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(a[i])); asm volatile("" : "+f"(a[i])); }
            if (OP == 7) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(bd));
            if (OP == 8) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(bd));
            if (OP == 9) asm volatile("div.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == 10) asm volatile("sqrt.rn.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 11) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 12) { // alternate add / mul (do both pipes overlap?)
                if (i & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                else       asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            }
            if (OP == 13) { // cvt + dadd pair, the accumulate idiom
                double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(a[i]));
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(t));
            }
            if (OP == 14) asm volatile("sub.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == 15) { int v = __float_as_int(a[i]); v = __shfl_xor_sync(0xffffffffu, v, 1); a[i] = __int_as_float(v); }
            if (OP == 16) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 17) { // mul then add, dependent (the unfused a*b+c idiom)
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f; double sd = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pk[i]));
        s += a[i] + lo + hi; sd += d[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    outd[blockIdx.x * blockDim.x + threadIdx.x] = sd;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int lanes_per_instr, int instr_per_iter)
{
    int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    int sms = prop.multiProcessorCount;
    int blocks_per_sm = 4, threads = 256;   // 32 warps / SM
    int blocks = sms * blocks_per_sm;
    float* out; double* outd; long long* cyc;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&outd, blocks * threads * sizeof(double));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    bench<OP><<<blocks, threads>>>(out, outd, cyc, 1.0001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<blocks, threads>>>(out, outd, cyc, 1.0001f);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
    double ops_per_sm = (double)blocks_per_sm * threads * ITERS * CHAINS * instr_per_iter * lanes_per_instr;
    printf("%-28s %8.2f lane-ops/clk/SM   (%.3f ms, %.0f cycles/block, err=%d)\n", name, ops_per_sm / avg, ms, avg, (int)err);
    delete[] h; cudaFree(out); cudaFree(outd); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    run<0>("add.f32", 1, 1);
    run<14>("sub.f32", 1, 1);
    run<1>("mul.f32", 1, 1);
    run<2>("fma.f32", 1, 1);
    run<12>("add/mul alternating", 1, 1);
    run<17>("mul->add dependent", 1, 2);
    run<3>("add.f32x2", 2, 1);
    run<4>("mul.f32x2", 2, 1);
    run<5>("fma.f32x2", 2, 1);
    run<6>("cvt.f64.f32", 1, 1);
    run<7>("add.f64", 1, 1);
    run<8>("fma.f64", 1, 1);
    run<13>("cvt.f64.f32+add.f64", 1, 1);
    run<9>("div.rn.f32", 1, 1);
    run<10>("sqrt.rn.f32", 1, 1);
    run<11>("rsqrt.approx", 1, 1);
    run<16>("rcp.approx", 1, 1);
    run<15>("shfl.xor", 1, 1);
    return 0;
}
