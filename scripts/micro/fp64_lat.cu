// fp64 latency / throughput microbenchmark (B200): dependent DFMA/DADD/DMUL chains and independent streams
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, double a, double b, int n, long long *cyc) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) x = __fma_rn(x, b, a);
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void latadd(double *out, double a, double b, int n, long long *cyc) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) { x = __dadd_rn(x, b); x = __dmul_rn(x, a); }
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void thr(double *out, double a, double b, int n, long long *cyc) {
    double x[ILP];
    for (int j = 0; j < ILP; j++) x[j] = a + threadIdx.x + j;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int j = 0; j < ILP; j++) x[j] = __fma_rn(x[j], b, a);
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < ILP; j++) s += x[j];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    int n = 4096;
    lat<<<1, 32>>>(out, 1.0000001, 0.9999999, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)h / (n * 16.0));
    latadd<<<1, 32>>>(out, 1.0000001, 0.9999999, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DADD/DMUL latency: %.2f cycles\n", (double)h / (n * 16.0));
    for (int warps = 1; warps <= 32; warps *= 2) {
        thr<8><<<1, 32 * warps>>>(out, 1.0000001, 0.9999999, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("1 SM, %2d warps x ILP 8: %.2f DFMA lanes/clk/SM\n", warps, (double)n * 4 * 8 * 32 * warps / h);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    thr<8><<<148 * 4, 512>>>(out, 1.0000001, 0.9999999, n, cyc);
    cudaEventRecord(e0);
    thr<8><<<148 * 4, 512>>>(out, 1.0000001, 0.9999999, n, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("full GPU: %.2f TFLOP/s fp64 (DFMA=2 flop)\n", 2.0 * n * 4 * 8 * 512.0 * 148 * 4 / (ms * 1e-3) / 1e12);
    return 0;
}
