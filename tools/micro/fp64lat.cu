// fp64 latency / throughput micro-benchmark (development tool): nvcc -arch=sm_100a -o fp64lat fp64lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) x[i] = fma(x[i], b, a);
                else if (OP == 1) x[i] = x[i] + a;
                else x[i] = x[i] * b;
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int OP>
void run(int warps, const char* name) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8);
    int iters = 2000;
    k<ILP, OP><<<1, warps * 32>>>(out, cyc, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    k<ILP, OP><<<1, warps * 32>>>(out, cyc, iters, 1.0000001, 0.9999999);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / (iters * 8.0 * ILP);
    printf("%s ILP=%d warps/SM=%d: %.2f cycles per instr per warp; SM rate %.2f warp-instr/cycle\n",
           name, ILP, warps, per, warps / per);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, 0>(1, "DFMA"); run<2, 0>(1, "DFMA"); run<4, 0>(1, "DFMA"); run<8, 0>(1, "DFMA"); run<16, 0>(1, "DFMA");
    run<1, 1>(1, "DADD"); run<4, 1>(1, "DADD"); run<8, 1>(1, "DADD");
    run<1, 2>(1, "DMUL"); run<8, 2>(1, "DMUL");
    run<8, 0>(4, "DFMA"); run<8, 0>(8, "DFMA"); run<8, 0>(16, "DFMA"); run<8, 0>(32, "DFMA");
    run<1, 0>(4, "DFMA"); run<1, 0>(8, "DFMA"); run<1, 0>(16, "DFMA"); run<1, 0>(32, "DFMA");
    run<2, 0>(8, "DFMA"); run<2, 0>(16, "DFMA"); run<4, 0>(8, "DFMA"); run<4, 0>(16, "DFMA");
    return 0;
}
