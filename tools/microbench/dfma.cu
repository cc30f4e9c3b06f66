// DFMA latency / issue microbenchmark: cycles per DFMA per warp for ILP chains x warps per SM
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// mixed: dependent chain of DFMA interleaved with LDS (shared memory round trip)
__global__ void klds(double* out, int iters, long long* cyc) {
    __shared__ double sm[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    int idx = threadIdx.x & 31;
    double x = 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) { x += sm[idx]; idx = (idx + 32 + (int)x) & 2047; }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int wps, double* out, long long* dc) {
    int iters = 2000;
    k<ILP><<<148, wps * 32>>>(out, iters, 1.0000001, 1e-9, dc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("ILP %d warps/SM %2d: %.2f cycles per DFMA per warp, %.2f DFMA-warp-instr per cycle per SM\n", ILP, wps,
           (double)c / (iters * 8.0 * ILP), (iters * 8.0 * ILP * wps) / (double)c);
}
int main() {
    double* out; long long* dc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    for (int wps : {1, 4, 8, 16, 32}) { run<1>(wps, out, dc); run<2>(wps, out, dc); run<4>(wps, out, dc); run<8>(wps, out, dc); }
    klds<<<148, 32>>>(out, 1000, dc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("LDS->DADD->IADD dependent round: %.1f cycles\n", (double)c / 8000.0);
    return 0;
}
