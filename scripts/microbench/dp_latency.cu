// FP64 pipe micro-benchmark for sm_100a: dependent-issue latency of DFMA/DADD/DMUL and the
// throughput reached with W warps per SM sub-partition x C independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_latency dp_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chains(double* out, int iters, double b, double c, long long* cycles) {
    double a[C];
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < C; ++i) a[i] = fma(a[i], b, c);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int C>
void run(int warps_per_sm, double* d_out, long long* d_cyc) {
    const int iters = 2000;
    int threads = 32 * warps_per_sm;      // one block per SM
    int blocks = 148;
    if (threads > 1024) { blocks = 148 * (threads / 1024); threads = 1024; }
    chains<C><<<blocks, threads>>>(d_out, iters, 1.0000001, 1e-9, d_cyc);
    cudaDeviceSynchronize();
    chains<C><<<blocks, threads>>>(d_out, iters, 1.0000001, 1e-9, d_cyc);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double n_inst = (double) iters * 8 * C;             // per warp
    const double per_smsp_warps = warps_per_sm / 4.0;
    // DP warp-instructions issued per cycle per sub-partition (peak = 0.5)
    const double rate = n_inst * per_smsp_warps / (double) cyc;
    printf("warps/SM %3d chains %d : %.2f cycles per dependent step, %.3f DP warp-inst/clk/SMSP (%.0f%% of peak)\n",
           warps_per_sm, C, (double) cyc / (iters * 8.0), rate, 100 * rate / 0.5);
}

int main() {
    double* d_out;  long long* d_cyc;
    cudaMalloc(&d_out, 148 * 2048 * sizeof(double));
    cudaMalloc(&d_cyc, sizeof(long long));
    for (int w : {4, 8, 16, 24, 32, 64}) {
        run<1>(w, d_out, d_cyc);
        run<2>(w, d_out, d_cyc);
        run<4>(w, d_out, d_cyc);
        run<8>(w, d_out, d_cyc);
    }
    return 0;
}
