// Dev microbenchmark: dependent-issue latency of FP64 ops, conversions and loads on one warp (B200).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int ILP, int OP>
__global__ void fp(double* out, double a, double b, long long* cyc) {
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (OP == 0) x[k] = __dadd_rn(x[k], b);
            if (OP == 1) x[k] = __dmul_rn(x[k], b);
            if (OP == 2) x[k] = __fma_rn(x[k], b, a);
            if (OP == 3) x[k] = (double)(int)(x[k]) + b;     // F2I + I2F + DADD
            if (OP == 4) x[k] = x[k] > b ? x[k] - a : x[k] + a;  // DSETP + select + DADD
            if (OP == 5) x[k] = x[k] / (b + x[k]);
        }
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chase(const int* p, int n, int* out, long long* cyc) {
    int j = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) j = __ldg(p + j);
    long long t1 = clock64();
    out[threadIdx.x] = j;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chase_cg(const int* p, int n, int* out, long long* cyc) {
    int j = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) j = __ldcg(p + j);
    long long t1 = clock64();
    out[threadIdx.x] = j;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    long long h;
    const char* names[] = {"DADD", "DMUL", "DFMA", "F2I+I2F+DADD", "DSETP+SEL+DADD", "DDIV(a/(b+a))"};
#define RUN(ILP, OP, W)                                                              \
    fp<ILP, OP><<<1, 32 * W>>>(out, 1.0, 1.0000001, cyc); cudaDeviceSynchronize(); \
    fp<ILP, OP><<<1, 32 * W>>>(out, 1.0, 1.0000001, cyc); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                  \
    printf("%-16s warps/CTA=%d ILP=%d: %.2f cycles per op-step (%.2f per instr-slot)\n", names[OP], W, ILP, (double)h / N, (double)h / N / ILP);
    RUN(1, 0, 1) RUN(2, 0, 1) RUN(4, 0, 1) RUN(8, 0, 1) RUN(16, 0, 1)
    RUN(1, 1, 1) RUN(1, 2, 1) RUN(4, 2, 1) RUN(8, 2, 1)
    RUN(1, 3, 1) RUN(4, 3, 1) RUN(1, 4, 1) RUN(4, 4, 1) RUN(1, 5, 1) RUN(4, 5, 1)
    RUN(1, 0, 4) RUN(4, 0, 4) RUN(8, 0, 4) RUN(1, 0, 8) RUN(4, 0, 8) RUN(8, 0, 8)
    // pointer chase: L1 hit (small footprint), L2 hit (4 MB, .cg), DRAM-ish (1 GB, .cg)
    for (int pass = 0; pass < 3; ++pass) {
        size_t n = pass == 0 ? 2048 : pass == 1 ? (1u << 20) : (1u << 28);
        int* p; cudaMalloc(&p, n * 4);
        int* hp = (int*)malloc(n * 4);
        size_t stride = pass == 0 ? 33 : 4099 * 32 + 1;
        for (size_t i = 0; i < n; ++i) hp[i] = (int)((i + stride) % n);
        cudaMemcpy(p, hp, n * 4, cudaMemcpyHostToDevice);
        int* o; cudaMalloc(&o, 4096);
        if (pass == 0) { chase<<<1, 32>>>(p, 2000, o, cyc); cudaDeviceSynchronize(); chase<<<1, 32>>>(p, 2000, o, cyc); }
        else { chase_cg<<<1, 32>>>(p, 2000, o, cyc); cudaDeviceSynchronize(); chase_cg<<<1, 32>>>(p, 2000, o, cyc); }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("pointer chase footprint %zu B (%s): %.1f cycles per dependent load\n", n * 4, pass == 0 ? "ld.nc L1" : "ld.cg", (double)h / 2000);
        cudaFree(p); free(hp);
    }
    return 0;
}
