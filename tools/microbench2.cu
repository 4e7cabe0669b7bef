// Latency / issue-rate microbenchmark for packed fp32 (FFMA2/FMUL2) vs scalar on sm_100a.
// For each op: W warps per SM sub-partition (blocks of 32*4*W threads, 1 block per SM), C independent chains per
// thread. Reports cycles per warp-instruction per SMSP assuming the SM clock given by nvidia-smi (printed).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 8192

template <int OP, int C>
__global__ void bench(unsigned long long* out, float seed)
{
    unsigned long long a[C], b, c;
    float one = seed, x = seed * 0.999f;
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(one));
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(x));
#pragma unroll
    for (int i = 0; i < C; ++i) { float v = seed + i + threadIdx.x * 1e-3f; asm volatile("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(v)); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
            if (OP == 1) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
            if (OP == 2) { // scalar fma on the low half
                float lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(lo) : "f"(one), "f"(x));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP, int C>
void run(const char* name, int warps_per_smsp, double ghz)
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int threads = 32 * 4 * warps_per_smsp, blocks = prop.multiProcessorCount;
    unsigned long long* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
    bench<OP, C><<<blocks, threads>>>(out, 1.0001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP, C><<<blocks, threads>>>(out, 1.0001f);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cycles = ms * 1e-3 * ghz * 1e9;
    double inst_per_smsp = (double)ITERS * C * warps_per_smsp;
    printf("%-10s warps/SMSP=%d chains=%d : %.2f cycles per warp-instruction per SMSP (%.3f ms)\n", name, warps_per_smsp, C, cycles / inst_per_smsp, ms);
    cudaFree(out);
}

int main(int argc, char** argv)
{
    double ghz = argc > 1 ? atof(argv[1]) : 1.92;
    printf("assuming %.3f GHz\n", ghz);
    run<0, 1>("ffma2", 1, ghz); run<0, 2>("ffma2", 1, ghz); run<0, 4>("ffma2", 1, ghz); run<0, 8>("ffma2", 1, ghz);
    run<0, 1>("ffma2", 4, ghz); run<0, 2>("ffma2", 4, ghz); run<0, 4>("ffma2", 4, ghz); run<0, 1>("ffma2", 8, ghz); run<0, 2>("ffma2", 8, ghz);
    run<1, 1>("fmul2", 1, ghz); run<1, 4>("fmul2", 1, ghz); run<1, 2>("fmul2", 4, ghz);
    run<2, 1>("ffma", 1, ghz); run<2, 4>("ffma", 1, ghz); run<2, 8>("ffma", 1, ghz); run<2, 2>("ffma", 4, ghz); run<2, 4>("ffma", 4, ghz);
    return 0;
}
