// STFT power spectral density frames: BufferedSpectrogram.process
// (src/audian/bufferedspectrogram.py:45-62) = thunderlab spectrogram ->
// scipy.signal.spectrogram(window='hann', detrend='constant', scaling='density',
// mode='psd') restated as one fused kernel per batch of frames:
//   frame k, channel c = src[k*hop : k*hop+nfft, c]
//   -> minus its mean -> x periodic Hann -> real FFT -> |X|^2 / (rate * sum(w^2))
//   -> x2 for 0 < bin < nfft/2 -> dst[k, c, :]            
// The real FFT of N points is a complex FFT of N/2 points (even samples = real,
// odd = imaginary part) in shared memory, radix-2^2 decimation in frequency
// (two stages per pass, bit-reversed output order), followed by the split step.
// All arithmetic is fp64 (the parity bar is rtol 1e-5 per bin against scipy's
// fp64 pocketfft, which an fp32 FFT cannot give for tonal frames).
// No cuFFT, no tensor cores.
#include "common.cuh"
#include <cmath>
#include <mutex>

namespace adn {

namespace {

constexpr int SP_NT = 256;

struct SpecPlan {
    int nfft = 0;
    double2* tw = nullptr;      // exp(-2 pi i j / nfft), j < nfft/2
    double* win = nullptr;      // periodic Hann
    double sumw2 = 0.0;
};

std::vector<SpecPlan> g_splans;
std::mutex g_splan_mu;

int32_t get_spec_plan(int nfft, cudaStream_t st, SpecPlan* out) {
    std::lock_guard<std::mutex> lk(g_splan_mu);
    for (auto& p : g_splans)
        if (p.nfft == nfft) { *out = p; return ADN_OK; }
    SpecPlan p;
    p.nfft = nfft;
    std::vector<double2> tw(nfft / 2);
    std::vector<double> win(nfft);
    const long double two_pi = 6.283185307179586476925286766559L;
    for (int j = 0; j < nfft / 2; ++j) {
        long double a = two_pi * (long double)j / (long double)nfft;
        tw[j].x = (double)cosl(a);
        tw[j].y = (double)(-sinl(a));
    }
    double s2 = 0.0;
    for (int j = 0; j < nfft; ++j) {
        double a = 2.0 * M_PI * (double)j / (double)nfft;
        win[j] = 0.5 - 0.5 * cos(a);
        s2 += win[j] * win[j];
    }
    p.sumw2 = s2;
    ADN_CK(cudaMalloc(&p.tw, sizeof(double2) * tw.size()));
    ADN_CK(cudaMalloc(&p.win, sizeof(double) * win.size()));
    ADN_CK(cudaMemcpyAsync(p.tw, tw.data(), sizeof(double2) * tw.size(), cudaMemcpyHostToDevice, st));
    ADN_CK(cudaMemcpyAsync(p.win, win.data(), sizeof(double) * win.size(), cudaMemcpyHostToDevice, st));
    ADN_CK(cudaStreamSynchronize(st));
    g_splans.push_back(p);
    *out = p;
    return ADN_OK;
}

__device__ __forceinline__ double2 cmul(double2 a, double2 w) {
    return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}

struct SpecArgs {
    const double* src;
    double* dst;
    const double2* tw;
    const double* win;
    int64_t nframes;
    int32_t C, nfft, hop, logM;
    int32_t FB, CB;           // frames and channels per block
    int32_t detrend, out_db;
    double scale;             // 1 / (rate * sum(w^2))
};

// generic path: any power-of-two nfft in [8, 16384]; a block owns FB frames x CB channels
__global__ void __launch_bounds__(SP_NT)
spectrogram_kernel(const __grid_constant__ SpecArgs P) {
    extern __shared__ __align__(16) double sbuf[];
    const int N = P.nfft, M = N >> 1, C = P.C;
    const int stride = N + 2;                       // doubles per item (bank shift between items)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * P.FB;
    const int c0 = blockIdx.y * P.CB;
    const int FBa = (int)min((int64_t)P.FB, P.nframes - f0);
    const int CBa = min(P.CB, C - c0);
    const int items = FBa * CBa;
    double* means = sbuf + (size_t)P.FB * P.CB * stride;

    // ---- load raw samples: consecutive threads read consecutive channels of a row
    for (int fi = 0; fi < FBa; ++fi) {
        const double* base = P.src + ((f0 + fi) * (int64_t)P.hop) * C + c0;
        const int total = N * CBa;
        for (int q = tid; q < total; q += SP_NT) {
            int j = q / CBa, ci = q - j * CBa;
            sbuf[(size_t)(fi * CBa + ci) * stride + j] = __ldg(base + (int64_t)j * C + ci);
        }
    }
    __syncthreads();
    // ---- frame means (detrend='constant'): one warp per item
    for (int it = warp; it < items; it += SP_NT / 32) {
        double s = 0.0;
        if (P.detrend) {
            const double* b = sbuf + (size_t)it * stride;
            for (int j = lane; j < N; j += 32) s += b[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            s /= (double)N;
        }
        if (lane == 0) means[it] = s;
    }
    __syncthreads();
    // ---- window
    {
        const int total = items * N;
        for (int q = tid; q < total; q += SP_NT) {
            int it = q / N, j = q - it * N;
            double* b = sbuf + (size_t)it * stride;
            b[j] = (b[j] - means[it]) * __ldg(P.win + j);
        }
    }
    __syncthreads();
    // ---- complex FFT of M points per item, DIF, two stages per pass
    const int logM = P.logM;
    int lg = logM;                                  // log2 of the current sub-transform length n
    for (; lg >= 2; lg -= 2) {
        const int n = 1 << lg, q4 = n >> 2;
        const int tstep = N >> lg;                  // W_n^k = tw[k * N / n]
        const int total = items * (M >> 2);
        for (int q = tid; q < total; q += SP_NT) {
            int it = q / (M >> 2), r = q - it * (M >> 2);
            int blk = r >> (lg - 2), k = r & (q4 - 1);
            double2* z = reinterpret_cast<double2*>(sbuf + (size_t)it * stride) + (blk << lg) + k;
            double2 a0 = z[0], a1 = z[q4], a2 = z[2 * q4], a3 = z[3 * q4];
            double2 w1 = __ldg(P.tw + k * tstep);               // W_n^k
            double2 w2 = __ldg(P.tw + 2 * k * tstep);           // W_{n/2}^k
            // stage n
            double2 b0 = make_double2(a0.x + a2.x, a0.y + a2.y);
            double2 b2 = cmul(make_double2(a0.x - a2.x, a0.y - a2.y), w1);
            double2 b1 = make_double2(a1.x + a3.x, a1.y + a3.y);
            double2 d13 = make_double2(a1.x - a3.x, a1.y - a3.y);
            // W_n^(k+n/4) = -i W_n^k : (x + iy)(-i) = y - ix
            double2 b3 = cmul(make_double2(d13.y, -d13.x), w1);
            // stage n/2
            z[0] = make_double2(b0.x + b1.x, b0.y + b1.y);
            z[q4] = cmul(make_double2(b0.x - b1.x, b0.y - b1.y), w2);
            z[2 * q4] = make_double2(b2.x + b3.x, b2.y + b3.y);
            z[3 * q4] = cmul(make_double2(b2.x - b3.x, b2.y - b3.y), w2);
        }
        __syncthreads();
    }
    if (lg == 1) {                                  // last single radix-2 stage (n = 2)
        const int total = items * (M >> 1);
        for (int q = tid; q < total; q += SP_NT) {
            int it = q / (M >> 1), r = q - it * (M >> 1);
            double2* z = reinterpret_cast<double2*>(sbuf + (size_t)it * stride) + 2 * r;
            double2 u = z[0], v = z[1];
            z[0] = make_double2(u.x + v.x, u.y + v.y);
            z[1] = make_double2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
    // ---- split step + power, bins 0..M; Z[k] sits at bit-reversed position
    {
        const int F = M + 1;
        const int total = items * F;
        const int sh = 32 - logM;
        for (int q = tid; q < total; q += SP_NT) {
            int it = q / F, k = q - it * F;
            const double2* z = reinterpret_cast<const double2*>(sbuf + (size_t)it * stride);
            double xr, xi, fac;
            if (k == 0 || k == M) {
                double2 z0 = z[0];
                xr = k == 0 ? z0.x + z0.y : z0.x - z0.y;
                xi = 0.0;
                fac = 1.0;
            } else {
                double2 zk = z[__brev((unsigned)k) >> sh];
                double2 zm = z[__brev((unsigned)(M - k)) >> sh];
                double er = 0.5 * (zk.x + zm.x), ei = 0.5 * (zk.y - zm.y);
                double orr = 0.5 * (zk.y + zm.y), oi = -0.5 * (zk.x - zm.x);
                double2 w = __ldg(P.tw + k);
                xr = er + (orr * w.x - oi * w.y);
                xi = ei + (orr * w.y + oi * w.x);
                fac = 2.0;
            }
            double pw = (xr * xr + xi * xi) * (P.scale * fac);
            if (P.out_db) pw = pw > 1e-20 ? 10.0 * log10(pw) : (pw <= 1e-20 ? -INFINITY : pw);
            int fi = it / CBa, ci = it - fi * CBa;
            P.dst[((f0 + fi) * (int64_t)C + c0 + ci) * F + k] = pw;
        }
    }
}

}  // namespace

int32_t spectrogram_dev(const double* src, int64_t n_src, int32_t C, double rate, int32_t nfft,
                        int32_t hop, int32_t window_id, int32_t detrend_id, double* dst,
                        int64_t n_dst, int32_t out_db, int64_t* n_computed, cudaStream_t st) {
    if (window_id != ADN_WINDOW_HANN)
        return fail(ADN_ERR_UNSUPPORTED, "spectrogram: window_id %d (only ADN_WINDOW_HANN)", window_id);
    if (detrend_id != ADN_DETREND_NONE && detrend_id != ADN_DETREND_CONSTANT)
        return fail(ADN_ERR_INVALID, "spectrogram: detrend_id %d", detrend_id);
    if (nfft < ADN_MIN_NFFT || nfft > ADN_MAX_NFFT || (nfft & (nfft - 1)))
        return fail(ADN_ERR_UNSUPPORTED, "spectrogram: nfft=%d (power of two in [%d, %d] required)",
                    nfft, ADN_MIN_NFFT, ADN_MAX_NFFT);
    const int64_t nf = spectrogram_frames(n_src, n_dst, nfft, hop);
    const size_t F = (size_t)nfft / 2 + 1;
    if (n_computed) *n_computed = nf;
    if (n_dst > nf)
        ADN_CK(cudaMemsetAsync(dst + (size_t)nf * C * F, 0, (size_t)(n_dst - nf) * C * F * 8, st));
    if (nf == 0) return ADN_OK;
    SpecPlan plan;
    int32_t rc = get_spec_plan(nfft, st, &plan);
    if (rc) return rc;
    SpecArgs P;
    P.src = src; P.dst = dst; P.tw = plan.tw; P.win = plan.win;
    P.nframes = nf; P.C = C; P.nfft = nfft; P.hop = hop;
    int logM = 0;
    while ((1 << logM) < nfft / 2) ++logM;
    P.logM = logM;
    P.detrend = detrend_id == ADN_DETREND_CONSTANT;
    P.out_db = out_db;
    P.scale = 1.0 / (rate * plan.sumw2);
    int max_items = 65536 / (nfft * 8);
    if (max_items < 1) max_items = 1;
    if (max_items > 64) max_items = 64;
    P.CB = C < max_items ? C : max_items;
    if (P.CB > 8) P.CB = 8;
    P.FB = max_items / P.CB;
    if (P.FB < 1) P.FB = 1;
    if ((int64_t)P.FB > nf) P.FB = (int32_t)nf;
    const size_t smem = ((size_t)P.FB * P.CB * (nfft + 2) + (size_t)P.FB * P.CB) * 8;
    static bool attr_done = false;
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(spectrogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    200 * 1024));
        attr_done = true;
    }
    int64_t gx = (nf + P.FB - 1) / P.FB;
    int gy = (C + P.CB - 1) / P.CB;
    if (gx > 0x7fffffff || gy > 65535)
        return fail(ADN_ERR_UNSUPPORTED, "spectrogram: grid %lld x %d", (long long)gx, gy);
    spectrogram_kernel<<<dim3((unsigned)gx, (unsigned)gy), SP_NT, smem, st>>>(P);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

}  // namespace adn
