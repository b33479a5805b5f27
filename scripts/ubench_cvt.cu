// Micro-benchmark: per-SM throughput of the instructions the decoder epilogue is made of (f32 -> f16x2 pack, f16 -> f32,
// FFMA) with 1..16 warps per SM; prints lane-operations per clock per SM.  Build: nvcc -arch=sm_100a -O3 -o bin/ubench_cvt ubench_cvt.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int OP>
__global__ void k(float *out, long long *clk, int iters)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) {   // pack two f32 -> f16x2
                uint32_t h;
                asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
                acc ^= h;
            } else if (OP == 1) {   // unpack f16x2 -> two f32
                float a, b;
                uint32_t h = __float_as_uint(x[2 * i]);
                asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(a), "=f"(b) : "r"(h));
                x[2 * i] = a; x[2 * i + 1] += b;
            } else if (OP == 2) {   // FFMA
                x[2 * i] = fmaf(x[2 * i], 1.0001f, 0.5f); x[2 * i + 1] = fmaf(x[2 * i + 1], 1.0001f, 0.5f);
            } else {                // the whole hi / lo split of two values
                uint32_t hi, lo;
                float h0, h1;
                asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
                asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
                const float r0 = x[2 * i] - h0, r1 = x[2 * i + 1] - h1;
                asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
                acc ^= hi ^ lo;
                x[2 * i] += 1.0f;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main()
{
    float *out; long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 2000;
    const char *names[4] = {"pack f32x2->f16x2 (per instr = 2 values)", "unpack f16x2->f32 (2 cvt)", "FFMA x2", "hi/lo split of 2 values"};
    for (int op = 0; op < 4; ++op)
        for (int warps = 1; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) k<0><<<148, warps * 32>>>(out, clk, iters);
                if (op == 1) k<1><<<148, warps * 32>>>(out, clk, iters);
                if (op == 2) k<2><<<148, warps * 32>>>(out, clk, iters);
                if (op == 3) k<3><<<148, warps * 32>>>(out, clk, iters);
                cudaDeviceSynchronize();
            }
            long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
            double c = (double)h[0];
            // "units" per iteration per thread: 8 (instr groups); lane-units per clk per SM
            printf("%-45s warps %2d: %.1f clk/iter, %.2f groups(lane)/clk/SM\n", names[op], warps, c / iters, 8.0 * warps * 32 * iters / c);
        }
    return 0;
}
