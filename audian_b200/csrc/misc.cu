// Element-wise kernels: decibel() (specitem.py:36 via thunderlab) and the
// counter-based synthetic recording generator (host twin in
// audian_b200/synth.py).
#include "common.cuh"
#include <cmath>

namespace adn {

__global__ void __launch_bounds__(256)
decibel_kernel(const double* __restrict__ p, int64_t n, double inv_ref, double min_power,
               double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double v = p[i];
        // NaN fails `v > min_power` and `v <= min_power`: numpy leaves the copy untouched
        double r = v;
        if (v > min_power) r = 10.0 * log10(v * inv_ref);
        else if (v <= min_power) r = -INFINITY;
        out[i] = r;
    }
}

int32_t decibel_dev(const double* p, int64_t n, double ref_power, double min_power, double* dst,
                    cudaStream_t st) {
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)ctx().sm_count * 16;
    if (blocks > cap) blocks = cap;
    decibel_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n, 1.0 / ref_power, min_power, dst);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct SynthInc { uint32_t inc[64]; };

__global__ void __launch_bounds__(256)
synth_kernel(double* __restrict__ dst, int64_t t0, int64_t n, int32_t C, uint64_t seed,
             uint32_t period, uint32_t on, const __grid_constant__ SynthInc incs) {
    int64_t total = n * C;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        uint64_t t = (uint64_t)(t0 + i / C);
        uint32_t c = (uint32_t)(i % C);
        uint64_t r = splitmix64(seed ^ (t * (uint64_t)C + c));
        int64_t noise = (int64_t)(int16_t)(uint16_t)(r >> 48);
        uint32_t phase = (uint32_t)(t * (uint64_t)incs.inc[c & 63]);
        int64_t q = (int64_t)(phase >> 16);
        int64_t tri = q < 32768 ? 2 * q - 32767 : 98303 - 2 * q;
        int64_t g = (t % period) < on ? 1 : 0;
        int64_t v = ((13107 * tri * g) >> 15) + (noise >> 5);
        v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
        dst[i] = (double)v * (1.0 / 32768.0);
    }
}

int32_t synth_dev(double* dst, int64_t t0, int64_t n, int32_t C, double rate, uint64_t seed,
                  cudaStream_t st) {
    SynthInc incs;
    for (int c = 0; c < 64; ++c)
        incs.inc[c] = (uint32_t)(uint64_t)llrint(4294967296.0 * (0.05 + 0.005 * c));
    // round-half-even like Python's round() in audian_b200/synth.py
    long long period = llrint(rate / 20.0), on = llrint(rate / 40.0);
    if (period < 2) period = 2;
    if (on < 1) on = 1;
    int64_t total = n * C;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)ctx().sm_count * 16;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(unsigned)blocks, 256, 0, st>>>(dst, t0, n, C, seed, (uint32_t)period,
                                                   (uint32_t)on, incs);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

}  // namespace adn
