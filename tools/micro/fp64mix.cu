// Does an fp64 instruction keep the issue port for both cycles of its half-rate pipe?  Development
// micro-benchmark: 8 independent DFMA per iteration mixed with K instructions of another pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o fp64mix fp64mix.cu
#include <cstdio>
#include <cuda_runtime.h>
// OTHER: 0 none, 1 FFMA, 2 integer add/xor, 3 LDS.64, 4 SHFL, 5 STS.64
template <int K, int OTHER>
__global__ void k(double* out, long long* cyc, int iters, double a, double b, float fa, float fb) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double x[8];
    float f[16];
    int n[16];
    double l[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) { f[i] = fa + i; n[i] = threadIdx.x + i; l[i] = 0.0; }
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = fma(x[i], b, a);
#pragma unroll
                for (int j = 0; j < K / 8; ++j) {
                    const int q = (i * (K / 8) + j) & 15;
                    if (OTHER == 1) f[q] = fmaf(f[q], fb, fa);
                    else if (OTHER == 2) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[q]) : "r"(it));
                    else if (OTHER == 3) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + ((lane + q * 32 + it) & 1023)))); l[q] += 0.0 * 0 + (q == 99 ? v : 0.0); n[q] ^= __double2loint(v); }
                    else if (OTHER == 4) n[q] = __shfl_xor_sync(0xffffffffu, n[q], 1);
                    else if (OTHER == 5) asm volatile("st.shared.f64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(sm + ((lane + q * 32) & 1023))), "d"(a) : "memory");
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i] + n[i] + l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int K, int OTHER>
void run(int warps, const char* name) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8);
    int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) k<K, OTHER><<<1, warps * 32>>>(out, cyc, iters, 1.0000001, 0.9999999, 1.0001f, 0.9999f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / (iters * 4.0);
    printf("8 DFMA + %2d %-5s warps/SM=%2d: %6.2f cycles per group per warp; per scheduler %.2f cycles per group (fp64 pipe alone: 16)\n",
           K, name, warps, per, per / (warps / 4.0));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 12}) {
        run<0, 0>(w, "none");
        run<8, 1>(w, "FFMA"); run<16, 1>(w, "FFMA"); run<24, 1>(w, "FFMA");
        run<8, 2>(w, "IADD"); run<16, 2>(w, "IADD");
        run<8, 3>(w, "LDS"); run<16, 3>(w, "LDS");
        run<8, 4>(w, "SHFL"); run<16, 4>(w, "SHFL");
        run<8, 5>(w, "STS"); run<16, 5>(w, "STS");
    }
    return 0;
}
