// Dev microbenchmark: DRAM bandwidth of the band kernels' access pattern (one thread = one column, per layer NR
// 8-byte reads and NW 8-byte writes per thread) for two scratch layouts:
//   PLANE : array a at [a][lay][col]           (each stream is its own plane, 150 MB apart)  -- the current layout
//   BLOCK : [col/128][lay][a][col%128]         (a CTA's data of one layer is one contiguous block)
#include <cstdio>
#include <cuda_runtime.h>
template <int NR, int NW, bool BLOCK>
__global__ void __launch_bounds__(128) k(const double* __restrict__ rd, double* __restrict__ wr, int ncol, int nlay) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    const long P = ncol, plane = (long)nlay * P;
    double acc = 0;
    for (int l = nlay - 1; l >= 0; --l) {
        double v[NR];
#pragma unroll
        for (int a = 0; a < NR; ++a)
            v[a] = BLOCK ? rd[(((long)blockIdx.x * nlay + l) * NR + a) * 128 + threadIdx.x] : rd[a * plane + (long)l * P + i];
#pragma unroll
        for (int a = 0; a < NR; ++a) acc += v[a];
#pragma unroll
        for (int a = 0; a < NW; ++a) {
            if (BLOCK) wr[(((long)blockIdx.x * nlay + l) * NW + a) * 128 + threadIdx.x] = acc + a;
            else wr[a * plane + (long)l * P + i] = acc + a;
        }
    }
}
template <int NR, int NW, bool BLOCK>
void run(const double* rd, double* wr, int ncol, int nlay, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NR, NW, BLOCK><<<ncol / 128, 128>>>(rd, wr, ncol, nlay); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) k<NR, NW, BLOCK><<<ncol / 128, 128>>>(rd, wr, ncol, nlay);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    double bytes = (double)(NR + NW) * 8 * ncol * nlay;
    printf("%-6s NR=%2d NW=%2d: %.3f ms  %.2f TB/s\n", name, NR, NW, ms, bytes / ms / 1e9);
}
int main() {
    const int ncol = 262144, nlay = 72;
    double *rd, *wr;
    cudaMalloc(&rd, (size_t)32 * ncol * nlay * 8); cudaMalloc(&wr, (size_t)32 * ncol * nlay * 8);
    cudaMemset(rd, 0, (size_t)32 * ncol * nlay * 8);
    run<19, 32, false>(rd, wr, ncol, nlay, "PLANE"); run<19, 32, true>(rd, wr, ncol, nlay, "BLOCK");
    run<32, 4, false>(rd, wr, ncol, nlay, "PLANE");  run<32, 4, true>(rd, wr, ncol, nlay, "BLOCK");
    run<16, 20, false>(rd, wr, ncol, nlay, "PLANE"); run<16, 20, true>(rd, wr, ncol, nlay, "BLOCK");
    run<8, 8, false>(rd, wr, ncol, nlay, "PLANE");   run<8, 8, true>(rd, wr, ncol, nlay, "BLOCK");
    return 0;
}
