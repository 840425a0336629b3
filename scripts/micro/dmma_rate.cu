// FP64 tensor-core instruction throughput on sm_100a: mma.sync m8n8k4 vs m16n8k4 / k8 / k16, and plain DFMA.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void __launch_bounds__(256) rate(double* out, int iters) {
    double c[8][4];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x * 1e-3 + i + j;
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = 0.5, a2 = 0.25, a3 = 0.125, a4 = 1.5, a5 = 2.5, a6 = 3.5, a7 = 4.5;
    double b0 = 1.0 - threadIdx.x * 1e-9, b1 = 0.75, b2 = 0.3, b3 = 0.2;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a0), "d"(b0));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b0));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
            } else if (SHAPE == 3) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7), "d"(b0), "d"(b1), "d"(b2), "d"(b3));
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) c[i][j] = fma(c[i][j], a0, b0);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double flops_per_inst_per_warp) {
    double* out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 256);
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        rate<SHAPE><<<148 * 8, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = 148.0 * 8 * 8, insts = warps * iters * 8;
    printf("%-22s %8.3f ms  %7.2f TFLOP/s  (%.1f cycles per instruction per SM sub-partition at 1.9 GHz)%s\n", name, ms,
           insts * flops_per_inst_per_warp / ms / 1e9, ms * 1e-3 * 1.9e9 / (insts / (148.0 * 4)),
           cudaGetLastError() == cudaSuccess ? "" : "  [error]");
    cudaFree(out);
}

int main() {
    run<0>("mma m8n8k4 f64", 2.0 * 8 * 8 * 4);
    run<1>("mma m16n8k4 f64", 2.0 * 16 * 8 * 4);
    run<2>("mma m16n8k8 f64", 2.0 * 16 * 8 * 8);
    run<3>("mma m16n8k16 f64", 2.0 * 16 * 8 * 16);
    run<4>("DFMA (4 per thread)", 2.0 * 32 * 4);
    return 0;
}
