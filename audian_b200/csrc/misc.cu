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

// decibel image of one channel of a spectrogram buffer (specitem.py:33-39):
// dst (F, n) = decibel(spec[:, channel, :].T); spec is (n, C, F) or, with C == 1, the
// (n, F) slice of one channel
__global__ void __launch_bounds__(256)
spec_image_kernel(const double* __restrict__ spec, int64_t n, int32_t C, int32_t F, int32_t channel,
                  double* __restrict__ dst) {
    __shared__ double tile[32][33];
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int64_t t = t0 + r;
        const int f = f0 + tx;
        double v = 0.0;
        if (t < n && f < F) v = spec[(t * C + channel) * F + f];
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int f = f0 + r;
        const int64_t t = t0 + tx;
        if (t < n && f < F) {
            double v = tile[tx][r];
            double o = v;
            if (v > 1e-20) o = 10.0 * log10(v);
            else if (v <= 1e-20) o = -INFINITY;
            dst[(int64_t)f * n + t] = o;
        }
    }
}

int32_t spec_image_dev(const double* spec, int64_t n, int32_t C, int32_t F, int32_t channel, double* dst,
                       cudaStream_t st) {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((F + 31) / 32));
    spec_image_kernel<<<grid, 256, 0, st>>>(spec, n, C, F, channel, dst);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

// power spectrum of the visible frames (spectrogramplot.py:158-160):
// dst[f] = max(decibel(mean(spec[i0:i1, channel, f])), floor_db)
__global__ void __launch_bounds__(256)
mean_power_kernel(const double* __restrict__ spec, int32_t C, int32_t F, int32_t channel, int64_t i0,
                  int64_t i1, double floor_db, double* __restrict__ dst) {
    __shared__ double part[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (f < F)
        for (int64_t t = i0 + ty; t < i1; t += 8) s += spec[(t * C + channel) * F + f];
    part[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && f < F) {
        // numpy's mean adds pairwise; the order here differs: ~1e-16 relative
        double a = 0.0;
        for (int k = 0; k < 8; ++k) a += part[k][tx];
        a /= (double)(i1 - i0);
        double o = a;
        if (a > 1e-20) o = 10.0 * log10(a);
        else if (a <= 1e-20) o = -INFINITY;
        if (o < floor_db) o = floor_db;
        dst[f] = o;
    }
}

int32_t mean_power_dev(const double* spec, int32_t C, int32_t F, int32_t channel, int64_t i0, int64_t i1,
                       double floor_db, double* dst, cudaStream_t st) {
    mean_power_kernel<<<(unsigned)((F + 31) / 32), 256, 0, st>>>(spec, C, F, channel, i0, i1, floor_db, dst);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

// PCM samples -> float64 in [-1, 1) times gain, the scaling audio readers apply
// (int16: / 2^15, packed little-endian int24: / 2^23, int32: / 2^31)
// acc[j] += sum_i spec[i, j], j < W: column sums of a (n, W) block of frames, accumulated
// across calls (the mean power spectrum of a whole recording streamed chunk by chunk,
// spectrogramplot.py:158 over all frames).  One block per 256 columns x a slab of rows.
__global__ void __launch_bounds__(256)
colsum_kernel(const double* __restrict__ spec, int64_t n, int64_t W, int64_t rows_per_block,
              double* __restrict__ acc) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= W) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = min(n, r0 + rows_per_block);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int64_t r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += __ldcs(spec + r * W + j);
        s1 += __ldcs(spec + (r + 1) * W + j);
        s2 += __ldcs(spec + (r + 2) * W + j);
        s3 += __ldcs(spec + (r + 3) * W + j);
    }
    for (; r < r1; ++r) s0 += __ldcs(spec + r * W + j);
    atomicAdd(acc + j, (s0 + s1) + (s2 + s3));
}

int32_t colsum_dev(const double* spec, int64_t n, int64_t W, double* acc, cudaStream_t st) {
    const int64_t bx = (W + 255) / 256;
    int64_t by = ((int64_t)ctx().sm_count * 8 + bx - 1) / bx;
    if (by > n) by = n;
    if (by < 1) by = 1;
    if (by > 65535) by = 65535;
    const int64_t rpb = (n + by - 1) / by;
    colsum_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, st>>>(spec, n, W, rpb, acc);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

__global__ void __launch_bounds__(256)
pcm_kernel(const unsigned char* __restrict__ pcm, int64_t n, int32_t bytes, double scale,
           double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const unsigned char* q = pcm + i * bytes;
        int32_t v;
        if (bytes == 2) v = (int16_t)((uint32_t)q[0] | ((uint32_t)q[1] << 8));
        else if (bytes == 3) v = ((int32_t)(((uint32_t)q[0] << 8) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 24))) >> 8;
        else v = (int32_t)((uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24));
        dst[i] = (double)v * scale;
    }
}

int32_t pcm_dev(const void* pcm, int64_t n, int32_t bits, double gain, double* dst, cudaStream_t st) {
    const int bytes = bits / 8;
    const double scale = gain / (double)((int64_t)1 << (bits - 1));
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx().sm_count * 16;
    if (blocks > cap) blocks = cap;
    pcm_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<const unsigned char*>(pcm), n, bytes, scale, dst);
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
             uint32_t period, uint32_t on, int64_t threads, const __grid_constant__ SynthInc incs) {
    // `threads` (a multiple of C) threads stride over the samples: a thread keeps its channel and
    // advances by threads / C rows per round, so the row-dependent quantities (noise counter,
    // carrier phase, position in the gate period) advance by constants -- no division in the loop
    const int64_t total = n * C;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= threads) return;
    const int64_t row0 = i / C;
    const uint32_t c = (uint32_t)(i - row0 * C);
    const int64_t drow = threads / C;
    const uint32_t inc = incs.inc[c & 63];
    uint64_t t = (uint64_t)(t0 + row0);
    uint64_t counter = t * (uint64_t)C + c;                  // argument of the noise hash
    uint32_t phase = (uint32_t)(t * (uint64_t)inc);
    uint32_t tmod = (uint32_t)(t % period);
    const uint32_t dphase = (uint32_t)((uint64_t)drow * (uint64_t)inc);
    const uint32_t dmod = (uint32_t)((uint64_t)drow % period);
    for (; i < total; i += threads) {
        const uint64_t r = splitmix64(seed ^ counter);
        const int32_t noise = (int32_t)(int16_t)(uint16_t)(r >> 48);
        const int32_t q = (int32_t)(phase >> 16);
        const int32_t tri = q < 32768 ? 2 * q - 32767 : 98303 - 2 * q;
        int32_t v = tmod < on ? (int32_t)(((int64_t)13107 * tri) >> 15) : 0;
        v += noise >> 5;
        v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
        dst[i] = (double)v * (1.0 / 32768.0);
        counter += (uint64_t)threads;
        phase += dphase;
        tmod += dmod;
        if (tmod >= period) tmod -= period;
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
    int64_t threads = blocks * 256 / C * C;                  // a multiple of C
    if (threads < C) { threads = C; blocks = (C + 255) / 256; }
    synth_kernel<<<(unsigned)blocks, 256, 0, st>>>(dst, t0, n, C, seed, (uint32_t)period,
                                                   (uint32_t)on, threads, incs);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

}  // namespace adn
