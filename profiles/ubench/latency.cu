#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, double a, double b, int n) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < n; ++i) { 
        x = x + b; x = x + b; x = x + b; x = x + b; x = x + b; x = x + b; x = x + b; x = x + b; }
    long long t1 = clock64();
    double y = a + threadIdx.x;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) { 
        y = y * b; y = y * b; y = y * b; y = y * b; y = y * b; y = y * b; y = y * b; y = y * b; }
    long long t2 = clock64();
    double z = a + threadIdx.x;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        z += __shfl_xor_sync(0xffffffffu, z, 1); z += __shfl_xor_sync(0xffffffffu, z, 2);
        z += __shfl_xor_sync(0xffffffffu, z, 1); z += __shfl_xor_sync(0xffffffffu, z, 2);
        z += __shfl_xor_sync(0xffffffffu, z, 1); z += __shfl_xor_sync(0xffffffffu, z, 2);
        z += __shfl_xor_sync(0xffffffffu, z, 1); z += __shfl_xor_sync(0xffffffffu, z, 2); }
    long long t3 = clock64();
    float f = (float)a + threadIdx.x;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) { 
        f = f + (float)b; f = f * (float)b; f = f + (float)b; f = f * (float)b; f = f + (float)b; f = f * (float)b; f = f + (float)b; f = f * (float)b; }
    long long t4 = clock64();
    __shared__ double sm[256];
    sm[threadIdx.x] = a;
    __syncthreads();
    long long t5 = clock64();
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory"); asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory");
        asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory"); asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory");
        asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory"); asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory");
        asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory"); asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory");
    }
    long long t6 = clock64();
    // store -> barrier -> dependent load chain (what a step does)
    int idx = threadIdx.x;
    double w = a;
    #pragma unroll 1
    for (int i = 0; i < n * 8; ++i) {
        sm[idx] = w + 1.0;
        asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory");
        w = sm[(idx + 33) % blockDim.x];
    }
    long long t7 = clock64();
    if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; clk[4] = t6 - t5; clk[5] = t7 - t6; }
    out[threadIdx.x] = x + y + z + f + w;
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 4096 * 8); cudaMalloc(&clk, 64);
    for (int threads : {32, 128, 256}) {
        int n = 1000;
        k<<<1, threads>>>(out, clk, 1.0, 1.0000001, n); cudaDeviceSynchronize();
        k<<<1, threads>>>(out, clk, 1.0, 1.0000001, n); cudaDeviceSynchronize();
        long long h[6]; cudaMemcpy(h, clk, 48, cudaMemcpyDeviceToHost);
        printf("threads %3d: DADD %.1f  DMUL %.1f  SHFL64+DADD %.1f  FADD/FMUL %.1f  named BAR %.1f  STS+BAR+LDS+DADD %.1f cycles\n", threads,
               h[0] / (8.0 * n), h[1] / (8.0 * n), h[2] / (8.0 * n), h[3] / (8.0 * n), h[4] / (8.0 * n), h[5] / (8.0 * n));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
