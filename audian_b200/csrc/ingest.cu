// Raw-data ingest and play-back preparation on the device:
//   unwrap          the loader option of Data.open (src/audian/data.py:180 -> audioio unwrap):
//                   data that wrapped around the +-1 range of its file format get the multiples
//                   of 2 back that the wrap took away
//   play_region     the signal DataBrowser.play_region hands to the audio device
//                   (src/audian/databrowser.py:1702-1731): mean over the shown channels (one or
//                   two output columns), heterodyne multiplication, zero-phase low-pass, decimation
//   gather_channel  one column of an interleaved trace as a contiguous vector (the visible-window
//                   min/max of one trace item, src/audian/traceitem.py:58-61)
#include "sos_common.cuh"
#include <cmath>

namespace adn {

namespace {

// ---------------------------------------------------------------- unwrap
constexpr int UW_ROWS = 2048;          // rows per chunk

// net number of +-2 corrections the rows of a chunk add (row r compares with row r - 1)
__global__ void __launch_bounds__(256)
unwrap_count_kernel(const double* __restrict__ src, int64_t n, int32_t C, double thresh,
                    int64_t nchunks, int32_t* __restrict__ counts) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nchunks * C) return;
    const int64_t chunk = id / C;
    const int c = (int)(id - chunk * C);
    const int64_t r0 = chunk * UW_ROWS, r1 = min(n, r0 + UW_ROWS);
    int32_t k = 0;
    double prev = r0 > 0 ? src[(r0 - 1) * C + c] : src[c];
    for (int64_t r = r0; r < r1; ++r) {
        const double v = src[r * C + c];
        const double d = v - prev;
        if (r > 0) {
            if (d < -thresh) ++k;
            else if (d > thresh) --k;
        }
        prev = v;
    }
    counts[id] = k;
}

// exclusive prefix over the chunks of a channel
__global__ void unwrap_prefix_kernel(int32_t* __restrict__ counts, int64_t nchunks, int32_t C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int32_t acc = 0;
    for (int64_t k = 0; k < nchunks; ++k) {
        const int32_t v = counts[k * C + c];
        counts[k * C + c] = acc;
        acc += v;
    }
}

__global__ void __launch_bounds__(256)
unwrap_apply_kernel(const double* __restrict__ src, int64_t n, int32_t C, double thresh, int32_t clips,
                    int64_t nchunks, const int32_t* __restrict__ counts, double* __restrict__ dst) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nchunks * C) return;
    const int64_t chunk = id / C;
    const int c = (int)(id - chunk * C);
    const int64_t r0 = chunk * UW_ROWS, r1 = min(n, r0 + UW_ROWS);
    int32_t k = counts[id];
    double prev = r0 > 0 ? src[(r0 - 1) * C + c] : src[c];
    for (int64_t r = r0; r < r1; ++r) {
        const double v = src[r * C + c];
        const double d = v - prev;
        if (r > 0) {
            if (d < -thresh) ++k;
            else if (d > thresh) --k;
        }
        prev = v;
        double o = v + 2.0 * (double)k;
        if (clips) o = o > 1.0 ? 1.0 : (o < -1.0 ? -1.0 : o);
        dst[r * C + c] = o;
    }
}

// ---------------------------------------------------------------- play_region
struct PlayCols {
    int32_t ncols;
    int32_t count[2];
    int32_t idx[2][64];
};

// numpy's pairwise summation of a contiguous run of n <= 128 values (np.mean over axis 1 of
// the fancy-indexed (rows, k) array), then the division by n
__device__ __forceinline__ double np_mean_small(const double* a, int n) {
    double res;
    if (n < 8) {
        res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
    } else {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
    }
    return res / (double)n;
}

__global__ void __launch_bounds__(256)
play_mix_kernel(const double* __restrict__ src, int64_t n, int32_t C, const __grid_constant__ PlayCols P,
                double w, double rate, int32_t heterodyne, double* __restrict__ dst) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n * P.ncols) return;
    const int64_t r = id / P.ncols;
    const int j = (int)(id - r * P.ncols);
    double a[64];
    const int cnt = P.count[j];
    for (int i = 0; i < cnt; ++i) a[i] = src[r * C + P.idx[j][i]];
    double v = np_mean_small(a, cnt);
    // np.sin(2*np.pi*f*np.arange(n)/rate): w = 2*pi*f evaluated on the host in that order
    if (heterodyne) v *= sin(w * (double)r / rate);
    dst[id] = v;
}

__global__ void __launch_bounds__(256)
decimate_kernel(const double* __restrict__ src, int64_t n_out, int32_t C, int64_t nstep,
                double* __restrict__ dst) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_out * C) return;
    const int64_t r = id / C;
    const int c = (int)(id - r * C);
    dst[id] = src[r * nstep * C + c];
}

__global__ void __launch_bounds__(256)
gather_channel_kernel(const double* __restrict__ src, int64_t n, int32_t C, int32_t channel,
                      double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i * C + channel];
}

unsigned blocks_for(int64_t items) {
    int64_t b = (items + 255) / 256;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

int32_t unwrap_dev(const double* src, int64_t n, int32_t C, double thresh, int32_t clips, double* dst,
                   cudaStream_t st) {
    const int64_t nchunks = (n + UW_ROWS - 1) / UW_ROWS;
    if (nchunks * C > 0x7fffffff) return fail(ADN_ERR_UNSUPPORTED, "unwrap: %lld chunks", (long long)nchunks);
    DevBuf& cb = scratch(SCR_UNWRAP, st);
    int32_t rc = cb.reserve((size_t)nchunks * C * sizeof(int32_t));
    if (rc) return rc;
    int32_t* counts = cb.as<int32_t>();
    unwrap_count_kernel<<<blocks_for(nchunks * C), 256, 0, st>>>(src, n, C, thresh, nchunks, counts);
    unwrap_prefix_kernel<<<(C + 63) / 64, 64, 0, st>>>(counts, nchunks, C);
    unwrap_apply_kernel<<<blocks_for(nchunks * C), 256, 0, st>>>(src, n, C, thresh, clips, nchunks, counts, dst);
    count_launch(3);
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

int32_t gather_channel_dev(const double* src, int64_t n, int32_t C, int32_t channel, double* dst,
                           cudaStream_t st) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx().sm_count * 16;
    if (b > cap) b = cap;
    gather_channel_kernel<<<(unsigned)(b < 1 ? 1 : b), 256, 0, st>>>(src, n, C, channel, dst);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

// playdata[:, j] = mean(src[:, group j], 1) (* sin(2 pi het_freq k / rate) if het_freq > 0);
// dst (n, 1 or 2)
int32_t play_mix_dev(const double* src, int64_t n, int32_t C, const int32_t* left, int32_t nleft,
                     const int32_t* right, int32_t nright, double rate, double het_freq, double* dst,
                     cudaStream_t st) {
    PlayCols P;
    P.ncols = nright > 0 ? 2 : 1;
    P.count[0] = nleft;
    P.count[1] = nright;
    for (int i = 0; i < nleft; ++i) P.idx[0][i] = left[i];
    for (int i = 0; i < nright; ++i) P.idx[1][i] = right[i];
    const double w = 2 * 3.141592653589793 * het_freq;       // 2*np.pi*f, numpy's order of operations
    play_mix_kernel<<<blocks_for(n * P.ncols), 256, 0, st>>>(src, n, C, P, w, rate, het_freq > 0.0 ? 1 : 0, dst);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

int32_t decimate_dev(const double* src, int64_t n_out, int32_t C, int64_t nstep, double* dst, cudaStream_t st) {
    decimate_kernel<<<blocks_for(n_out * C), 256, 0, st>>>(src, n_out, C, nstep, dst);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

}  // namespace adn
