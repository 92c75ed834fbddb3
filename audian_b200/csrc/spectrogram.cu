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
#include <cuda.h>               // CUtensorMap: types only, the encoder comes through cudaGetDriverEntryPoint
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace adn {

namespace {

constexpr int SP_NT = 256;

struct SpecPlan {
    int nfft = 0;
    double2* tw = nullptr;      // exp(-2 pi i j / nfft), j < nfft/2
    double2* twA = nullptr;     // warp path: [16][T] W_M^(t k1), M = nfft/2, T = M/16
    double* win = nullptr;      // periodic Hann
    double sumw2 = 0.0;
    // nfft not a power of two (Bluestein): transforms of length L = 2^logL
    int logL = 0;
    double2* twL = nullptr;     // exp(-2 pi i j / L), j < L/2
    double2* chirp = nullptr;   // exp(+i pi n^2 / nfft), n < nfft
    double2* Bbr = nullptr;     // transform of the wrapped chirp, bit-reversed order
};

std::vector<SpecPlan> g_splans;
std::mutex g_splan_mu;

int32_t big_fft_dif(double2* work, int64_t items, int logL, const double2* tw, int twlog, cudaStream_t st);

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
    if (nfft >= 128 && nfft <= 1024) {
        const int M = nfft / 2, T = M / 16;
        std::vector<double2> twA((size_t)M);
        for (int k1 = 0; k1 < 16; ++k1)
            for (int t = 0; t < T; ++t) {
                long double a = two_pi * (long double)((k1 * t) % M) / (long double)M;
                twA[(size_t)k1 * T + t].x = (double)cosl(a);
                twA[(size_t)k1 * T + t].y = (double)(-sinl(a));
            }
        ADN_CK(cudaMalloc(&p.twA, sizeof(double2) * twA.size()));
        ADN_CK(cudaMemcpy(p.twA, twA.data(), sizeof(double2) * twA.size(), cudaMemcpyHostToDevice));
    }
    ADN_CK(cudaMalloc(&p.tw, sizeof(double2) * (tw.size() + 1)));
    ADN_CK(cudaMalloc(&p.win, sizeof(double) * win.size()));
    ADN_CK(cudaMemcpyAsync(p.tw, tw.data(), sizeof(double2) * tw.size(), cudaMemcpyHostToDevice, st));
    ADN_CK(cudaMemcpyAsync(p.win, win.data(), sizeof(double) * win.size(), cudaMemcpyHostToDevice, st));
    ADN_CK(cudaStreamSynchronize(st));
    if (nfft & (nfft - 1)) {
        // circular length: the lags k - n, k <= nfft/2, n < nfft must not alias
        int logL = 1;
        while (((int64_t)1 << logL) < (int64_t)nfft + nfft / 2 + 1) ++logL;
        const int64_t L = (int64_t)1 << logL;
        p.logL = logL;
        std::vector<double2> twL((size_t)(L / 2)), ch((size_t)nfft), b((size_t)L, make_double2(0.0, 0.0));
        for (int64_t j = 0; j < L / 2; ++j) {
            long double a = two_pi * (long double)j / (long double)L;
            twL[(size_t)j] = make_double2((double)cosl(a), (double)(-sinl(a)));
        }
        for (int64_t n = 0; n < nfft; ++n) {
            const int64_t q = (n * n) % (2 * (int64_t)nfft);          // exact: n < 2^20
            long double a = (two_pi / 2) * (long double)q / (long double)nfft;
            ch[(size_t)n] = make_double2((double)cosl(a), (double)sinl(a));
        }
        for (int64_t j = 0; j <= nfft / 2; ++j) b[(size_t)j] = ch[(size_t)j];
        for (int64_t m = 1; m < nfft; ++m) b[(size_t)(L - m)] = ch[(size_t)m];
        ADN_CK(cudaMalloc(&p.twL, sizeof(double2) * twL.size()));
        ADN_CK(cudaMalloc(&p.chirp, sizeof(double2) * ch.size()));
        ADN_CK(cudaMalloc(&p.Bbr, sizeof(double2) * b.size()));
        ADN_CK(cudaMemcpyAsync(p.twL, twL.data(), sizeof(double2) * twL.size(), cudaMemcpyHostToDevice, st));
        ADN_CK(cudaMemcpyAsync(p.chirp, ch.data(), sizeof(double2) * ch.size(), cudaMemcpyHostToDevice, st));
        ADN_CK(cudaMemcpyAsync(p.Bbr, b.data(), sizeof(double2) * b.size(), cudaMemcpyHostToDevice, st));
        int32_t rc = big_fft_dif(p.Bbr, 1, logL, p.twL, logL, st);
        if (rc) return rc;
        ADN_CK(cudaStreamSynchronize(st));
    }
    g_splans.push_back(p);
    *out = p;
    return ADN_OK;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
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


// ======================================================================================
// Register-resident path for nfft = 128 .. 1024 (the interactive sizes): one frame is
// transformed by T = nfft/32 lanes of a warp (4 .. 32), 16 complex points per lane.
//   M = nfft/2 = 16*T complex points z[j] = (x[2j], x[2j+1]);  lane t holds z[T p + t]
//   pass 1   16-point DFT over p in registers, twiddle W_M^(t k1)
//   exchange through a per-warp shared-memory buffer (conflict-free padded layout)
//   pass 2   T-point DFTs over t in registers (T = 32: 16-point + one xor-shuffle stage)
//   split    Z[k], Z[M-k] -> X[k], X[M-k] of the real transform, |X|^2 scaling, store
// The input rows of FB consecutive frames x CB channels are staged once per block and
// de-interleaved on the way in (8-byte cp.async), so every lane reads its 16-byte pairs
// of consecutive samples without bank conflicts; only __syncwarp between the phases.

constexpr int SW_NWARP = 4;
constexpr int SW_NT = SW_NWARP * 32;

struct SpecWArgs {
    const double* src;
    double* dst;
    const double* win;          // nfft
    const double2* twA;         // [16][T]: W_M^(t k1)
    const double2* twS;         // W_N^k, k <= M/2
    int64_t nframes;
    int32_t C, hop, FB, CB, RP;
    int32_t detrend, out_db;
    double scale;               // 1 / (rate * sum(w^2))
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// a - i b  and  a + i b
__device__ __forceinline__ double2 csub_i(double2 a, double2 b) { return make_double2(a.x + b.y, a.y - b.x); }
__device__ __forceinline__ double2 cadd_i(double2 a, double2 b) { return make_double2(a.x - b.y, a.y + b.x); }

__device__ __forceinline__ void dft4(double2 x0, double2 x1, double2 x2, double2 x3,
                                     double2& y0, double2& y1, double2& y2, double2& y3) {
    double2 s02 = cadd(x0, x2), d02 = csub(x0, x2), s13 = cadd(x1, x3), d13 = csub(x1, x3);
    y0 = cadd(s02, s13);
    y2 = csub(s02, s13);
    y1 = csub_i(d02, d13);
    y3 = cadd_i(d02, d13);
}

#define ADN_C1 0.92387953251128674     /* cos(pi/8) */
#define ADN_S1 0.38268343236508977     /* sin(pi/8) */
#define ADN_R2 0.70710678118654752     /* sqrt(1/2) */

// natural-order 16-point forward DFT, 4 x 4 decomposition
__device__ __forceinline__ void dft16(const double2 (&a)[16], double2 (&X)[16]) {
    double2 u[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2)
        dft4(a[n2], a[n2 + 4], a[n2 + 8], a[n2 + 12], u[n2][0], u[n2][1], u[n2][2], u[n2][3]);
    // u[n2][k1] *= W16^(n2 k1)
    u[1][1] = cmul(u[1][1], make_double2(ADN_C1, -ADN_S1));
    u[1][2] = cmul(u[1][2], make_double2(ADN_R2, -ADN_R2));
    u[1][3] = cmul(u[1][3], make_double2(ADN_S1, -ADN_C1));
    u[2][1] = cmul(u[2][1], make_double2(ADN_R2, -ADN_R2));
    u[2][2] = make_double2(u[2][2].y, -u[2][2].x);                       // W16^4 = -i
    u[2][3] = cmul(u[2][3], make_double2(-ADN_R2, -ADN_R2));
    u[3][1] = cmul(u[3][1], make_double2(ADN_S1, -ADN_C1));
    u[3][2] = cmul(u[3][2], make_double2(-ADN_R2, -ADN_R2));
    u[3][3] = cmul(u[3][3], make_double2(-ADN_C1, ADN_S1));              // W16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
        dft4(u[0][k1], u[1][k1], u[2][k1], u[3][k1], X[k1], X[k1 + 4], X[k1 + 8], X[k1 + 12]);
}

// 8-point forward DFT of a[0..7] (2 x 4)
__device__ __forceinline__ void dft8(const double2* a, double2* X) {
    double2 u0[4], u1[4];
    dft4(a[0], a[2], a[4], a[6], u0[0], u0[1], u0[2], u0[3]);
    dft4(a[1], a[3], a[5], a[7], u1[0], u1[1], u1[2], u1[3]);
    u1[1] = cmul(u1[1], make_double2(ADN_R2, -ADN_R2));
    u1[2] = make_double2(u1[2].y, -u1[2].x);
    u1[3] = cmul(u1[3], make_double2(-ADN_R2, -ADN_R2));
#pragma unroll
    for (int k = 0; k < 4; ++k) { X[k] = cadd(u0[k], u1[k]); X[k + 4] = csub(u0[k], u1[k]); }
}

// W32^k = exp(-2 pi i k / 32), k < 16
__host__ __device__ constexpr double w32c(int k, int im) {
    constexpr double tab[16][2] = {
        {1.0, -0.0},
        {0.98078528040323043, -0.19509032201612825}, {0.92387953251128674, -0.38268343236508977},
        {0.83146961230254524, -0.55557023301960218}, {0.70710678118654752, -0.70710678118654752},
        {0.55557023301960218, -0.83146961230254524}, {0.38268343236508977, -0.92387953251128674},
        {0.19509032201612825, -0.98078528040323043}, {0.0, -1.0},
        {-0.19509032201612825, -0.98078528040323043}, {-0.38268343236508977, -0.92387953251128674},
        {-0.55557023301960218, -0.83146961230254524}, {-0.70710678118654752, -0.70710678118654752},
        {-0.83146961230254524, -0.55557023301960218}, {-0.92387953251128674, -0.38268343236508977},
        {-0.98078528040323043, -0.19509032201612825}};
    return tab[k][im];
}

template <int LOGN> struct SWCfg {
    static constexpr int N = 1 << LOGN, M = N / 2, T = M / 16, FPW = 32 / T;
    // per-frame stride of the per-warp exchange buffer (complex elements), == 4 mod 8
    static constexpr int FS = T == 32 ? 16 * 34 : (T * 17 + ((T * 17) % 8 == 4 ? 0 : 4));
    static constexpr int WB = FS * FPW;
};

template <int LOGN>
__global__ void __launch_bounds__(SW_NT, 3)
spectrogram_warp_kernel(const __grid_constant__ SpecWArgs P) {
    using Cf = SWCfg<LOGN>;
    constexpr int N = Cf::N, M = Cf::M, T = Cf::T, FPW = Cf::FPW, FS = Cf::FS;
    constexpr int F = M + 1;
    extern __shared__ __align__(16) double sbuf[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = P.C, hop = P.hop, RP = P.RP;
    const int ngrp = (C + P.CB - 1) / P.CB;
    const int grp = blockIdx.x % ngrp;
    const int64_t f0 = (int64_t)(blockIdx.x / ngrp) * P.FB;
    const int c0 = grp * P.CB;
    const int FBa = (int)min((int64_t)P.FB, P.nframes - f0);
    const int CBa = min(P.CB, C - c0);
    const int items = FBa * CBa;

    double* xs = sbuf;                                               // [CB][RP]
    double2* wb = reinterpret_cast<double2*>(sbuf + (size_t)P.CB * RP) + (size_t)warp * Cf::WB;

    // ---- stage: rows [f0*hop, f0*hop + (FBa-1)*hop + N) x CBa channels, de-interleaved;
    // a thread keeps its channel, so addresses just advance by a constant
    {
        const int rows = (FBa - 1) * hop + N;
        const int CBs = P.CB;                          // 1 or 2
        const int ci = tid % CBs, rstep = SW_NT / CBs;
        if (ci < CBa) {
            const double* gp = P.src + (f0 * hop) * (int64_t)C + c0 + ci + (int64_t)(tid / CBs) * C;
            double* sp = xs + (size_t)ci * RP + tid / CBs;
            const int64_t gstep = (int64_t)rstep * C;
            int row = tid / CBs;
            for (; row + 3 * rstep < rows; row += 4 * rstep) {
                cp_async8(sp, gp);
                cp_async8(sp + rstep, gp + gstep);
                cp_async8(sp + 2 * rstep, gp + 2 * gstep);
                cp_async8(sp + 3 * rstep, gp + 3 * gstep);
                sp += 4 * rstep;
                gp += 4 * gstep;
            }
            for (; row < rows; row += rstep) {
                cp_async8(sp, gp);
                sp += rstep;
                gp += gstep;
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();

    const int sub = lane / T, t = lane % T;
    double2* wbf = wb + sub * FS;
    // W_M^(t k1), k1 = 1, 2, 4, 8 from the table; the other powers are products of two of them
    const double2 tw1 = __ldg(P.twA + 1 * T + t), tw2 = __ldg(P.twA + 2 * T + t);
    const double2 tw4 = __ldg(P.twA + 4 * T + t), tw8 = __ldg(P.twA + 8 * T + t);
    const int niter = (items + SW_NWARP * FPW - 1) / (SW_NWARP * FPW);
    for (int iter = 0; iter < niter; ++iter) {
        const int it = (warp + SW_NWARP * iter) * FPW + sub;
        const bool live = it < items;
        const int ite = live ? it : 0;
        const int fi = ite / CBa, ci = ite - fi * CBa;
        const double* xr = xs + (size_t)ci * RP + fi * hop;

        // ---- load, detrend, window
        double2 a[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) a[p] = *reinterpret_cast<const double2*>(xr + 2 * (T * p + t));
        double mean = 0.0;
        if (P.detrend) {
            double s = 0.0;
#pragma unroll
            for (int p = 0; p < 16; ++p) s += a[p].x + a[p].y;
#pragma unroll
            for (int o = 1; o < T; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            mean = s * (1.0 / N);
        }
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            double2 w = __ldg(reinterpret_cast<const double2*>(P.win) + (T * p + t));
            a[p].x = (a[p].x - mean) * w.x;
            a[p].y = (a[p].y - mean) * w.y;
        }
        // ---- pass 1 + twiddle, into the exchange buffer at (k1, t)
        double2 b[16];
        dft16(a, b);
        {
            double2 w[16];
            w[1] = tw1; w[2] = tw2; w[4] = tw4; w[8] = tw8;
            w[3] = cmul(tw2, tw1); w[5] = cmul(tw4, tw1); w[6] = cmul(tw4, tw2);
            w[9] = cmul(tw8, tw1); w[10] = cmul(tw8, tw2); w[12] = cmul(tw8, tw4);
            w[7] = cmul(w[6], tw1); w[11] = cmul(w[10], tw1); w[13] = cmul(w[12], tw1);
            w[14] = cmul(w[12], tw2); w[15] = cmul(w[14], tw1);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                double2 v = k1 == 0 ? b[0] : cmul(b[k1], w[k1]);
                const int f = k1 * T + t;
                wbf[T == 32 ? k1 * 34 + t : f + (f >> 4)] = v;
            }
        }
        __syncwarp();
        // ---- pass 2: T-point DFTs over t
        double2 zout[16];
        int kout[16];
        if (T == 32) {
            const int k1 = lane >> 1, tp = lane & 1;
            const double sgn = tp ? -1.0 : 1.0;
#pragma unroll
            for (int pp = 0; pp < 16; ++pp) a[pp] = wbf[k1 * 34 + 2 * pp + tp];
            dft16(a, b);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                double2 e = b[k];
                if (tp) e = cmul(e, make_double2(w32c(k, 0), w32c(k, 1)));
                double2 o;
                o.x = __shfl_xor_sync(0xffffffffu, e.x, 1);
                o.y = __shfl_xor_sync(0xffffffffu, e.y, 1);
                zout[k] = make_double2(fma(sgn, e.x, o.x), fma(sgn, e.y, o.y));   // e0 + e1 | e0 - e1
                kout[k] = k1 + 16 * k + 256 * tp;
            }
        } else {
            constexpr int Q = 16 / (T == 32 ? 16 : T);
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = wbf[t * 17 + i];
            if (T == 16) {
                dft16(a, b);
            } else if (T == 8) {
                dft8(&a[0], &b[0]);
                dft8(&a[8], &b[8]);
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    dft4(a[4 * r], a[4 * r + 1], a[4 * r + 2], a[4 * r + 3],
                         b[4 * r], b[4 * r + 1], b[4 * r + 2], b[4 * r + 3]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int r = i / (16 / Q), k2 = i % (16 / Q);
                zout[i] = b[i];
                kout[i] = (t * Q + r) + 16 * k2;
            }
        }
        __syncwarp();
        // ---- Z in natural order
#pragma unroll
        for (int i = 0; i < 16; ++i) wbf[kout[i] + (kout[i] > 256 ? 4 : 0)] = zout[i];
        __syncwarp();
        // ---- split step and power
        if (live) {
            double* out = P.dst + (((f0 + fi) * (int64_t)C + c0 + ci) * F);
            const double sc = 0.5 * P.scale;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double2 zk[4], zm[4], tw[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {              // k = 1 + t + T i  in [1, M/2]
                    const int k = 1 + t + T * (4 * h + j), km = M - k;
                    zk[j] = wbf[k];
                    zm[j] = wbf[km + (km > 256 ? 4 : 0)];
                    tw[j] = __ldg(P.twS + k);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = 1 + t + T * (4 * h + j), km = M - k;
                    double e_r = zk[j].x + zm[j].x, e_i = zk[j].y - zm[j].y;      // Zk + conj(Zm)
                    double o_r = zk[j].y + zm[j].y, o_i = zm[j].x - zk[j].x;      // -i (Zk - conj(Zm))
                    double t_r = o_r * tw[j].x - o_i * tw[j].y, t_i = o_r * tw[j].y + o_i * tw[j].x;
                    double pr = e_r + t_r, pi = e_i + t_i, qr = e_r - t_r, qi = e_i - t_i;
                    double pk = (pr * pr + pi * pi) * sc, pm = (qr * qr + qi * qi) * sc;
                    if (P.out_db) {
                        pk = pk > 1e-20 ? 10.0 * log10(pk) : (pk <= 1e-20 ? -INFINITY : pk);
                        pm = pm > 1e-20 ? 10.0 * log10(pm) : (pm <= 1e-20 ? -INFINITY : pm);
                    }
                    out[k] = pk;
                    if (km != k) out[km] = pm;
                }
            }
            if (t == 0) {
                double2 z0 = wbf[0];
                double p0 = (z0.x + z0.y) * (z0.x + z0.y) * P.scale;
                double pM = (z0.x - z0.y) * (z0.x - z0.y) * P.scale;
                if (P.out_db) {
                    p0 = p0 > 1e-20 ? 10.0 * log10(p0) : (p0 <= 1e-20 ? -INFINITY : p0);
                    pM = pM > 1e-20 ? 10.0 * log10(pM) : (pM <= 1e-20 ? -INFINITY : pM);
                }
                out[0] = p0;
                out[M] = pM;
            }
        }
        __syncwarp();
    }
}


// ======================================================================================
// Streaming variant of the register-resident path: a block owns a run of consecutive
// frames of one channel group and walks along time.  The rows of the group are kept
// de-interleaved in a shared-memory ring (one array per channel); every step the block
//   - issues the 16-byte global loads of the rows the NEXT step adds (full row segments,
//     coalesced; held in registers while the FFTs run),
//   - transforms the FSTEP x W frames of this step,
//   - scatters the loaded rows into the ring (two 8-byte stores per vector, conflict
//     free because the channel arrays start RS = 4 mod 8 doubles apart), one barrier.
// Every input row is read from global memory exactly once per run (plus nfft - hop rows
// at the start of a run), nothing is staged through LDGSTS, and the loads overlap the
// arithmetic instead of preceding it.  The ring is a power of two rows long (exactly one
// frame when a step is one frame), so the chunk of the next step overwrites the oldest rows
// between two barriers and positions wrap with a mask; three blocks of four warps fit an SM.
// Differences in the arithmetic: the frame mean is removed after the transform (the
// spectrum of the periodic Hann window is N/2 at bin 0 and -N/4 at bins +-1, so only the
// output bins 0 and 1 change), which takes the reduction over the frame off the critical
// path; for T = 32 the lane pair of a 32-point DFT reads all 32 points and does the first
// radix-2 step itself (decimation in frequency) instead of exchanging results by shuffle.
constexpr int SR_MAXNT = 128;
// measured on B200 (nfft 1024, hop 512, 8 ch): 3 resident blocks, window in shared memory 172 us;
// window through L1 instead 183 us; 4 resident blocks at 128 registers (spills) 232 us
#ifndef SR_BLOCKS
#define SR_BLOCKS 3
#endif
// SR_WIN_CALC: the Hann window computed where it is used, w[j] = 1/2 - 1/2 cos(2 pi j / N) with
// j = 2 (T p + t) + s: cos(theta_p + phi) from the sixteen compile-time (cos, sin)(2 pi p / 16) and
// two per-lane (cos, sin) pairs -- two DFMA per value instead of half a 128-bit shared-memory load
// (and 8 KB less shared memory: four blocks per SM fit with SR_ASYNC).  Measured on B200 (8 ch,
// nfft 1024 / hop 512): 175 us against 167 us with the table in shared memory (the eight extra
// live registers spill); with SR_ASYNC and four blocks per SM 176 us.  Off.
#ifndef SR_WIN_CALC
#define SR_WIN_CALC 0
#endif
#ifndef SR_WIN_SMEM
#define SR_WIN_SMEM (SR_WIN_CALC ? 0 : 1)
#endif
// SR_WIN_HALF: only the first half of the window is kept in shared memory: the periodic Hann
// window has w[j + N/2] = 1 - w[j], so the second half of a frame is x - x w[j] (one DFMA in the
// place of the DMUL).  Half the window loads of a frame and 4 KB less shared memory per block; the
// two forms of w[j + N/2] differ by less than 2^-53 absolute.
#ifndef SR_WIN_HALF
#define SR_WIN_HALF 1
#endif
// SR_PAIRMAP: second pass with the two lanes of a row next to each other (T == 32): their identical
// 128-bit reads of the exchange buffer are served as one access.  Measured on B200 (8 ch, nfft
// 1024 / hop 512; ncu): shared-memory load wavefronts per frame 231 -> 164 (with SR_WIN_HALF: 263
// -> 164), 164.9 -> 163.8 us -- the kernel is not bound by the shared-memory pipe.
#ifndef SR_PAIRMAP
#define SR_PAIRMAP 1
#endif
// SR_LDCS: the rows of the next chunk come in with evict-first loads (they are used once by this
// block).  Measured on B200: nothing gained next to the default store policy (157.7 us both ways) and
// DRAM reads grow from 253 to 286 MB (the other channel group's block finds the line evicted).  Off.
#ifndef SR_LDCS
#define SR_LDCS 0
#endif

// Timing ablations (wrong results, build flags of tools/build_alt.sh only): SR_ABL & 1: no output
// stores (the arithmetic stays live), & 2: no loads of the next chunk and no ring fill, & 4: no
// block barriers around the fill
#ifndef SR_ABL
#define SR_ABL 0
#endif
#define SR_ST_OK(v) (!(SR_ABL & 1) || (v) == 1.2345e300)
constexpr int SR_PF = 8;            // 16-byte vectors in flight per thread and step
// SR_ASYNC: the rows of the next step go straight into the ring with 8-byte cp.async, issued as
// soon as every warp holds its frame of this step in registers (one barrier after the frame
// loads): the copies land while the transforms run and no register stages them.  Needs one
// item per warp and step (the launcher sees to it).  Measured on B200 (nfft 1024, hop 512):
// 8 ch 178 us against 172 us with register staging, 1 ch 155 / 145 us, 64 ch 396 / 407 us: the
// barrier in the middle of the step costs what the hidden load latency gains.  With shorter steps
// (hop < 512) it wins 3 - 9 % (measured after the round-2 changes of the small transforms: 64 ch
// 256 / 128 316 -> 289 us, 512 / 128 482 -> 454 us, 1024 / 128 951 -> 899 us; 8 ch 1024 / 128 597 ->
// 560 us, 256 / 32 521 -> 491 us), so it is a template parameter chosen by shape in launch_ring();
// SR_ASYNC = 1 forces it for every launch.
#ifndef SR_ASYNC
#define SR_ASYNC 0
#endif

struct SpecRArgs {
    const double* src;
    double* dst;
    const double* win;          // nfft
    const double2* twA;         // [16][T]: W_M^(t k1)
    const double2* twS;         // W_N^k, k <= M/2
    int64_t nframes;
    int64_t nrows;              // source rows the frames read: (nframes-1)*hop + nfft
    int32_t C, hop, CB, LW, ngrp;   // CB = 1 << LW channels per group
    int32_t FSTEP, FRUN;        // frames per step / per block (FRUN % FSTEP == 0)
    int32_t RC, RS;             // ring length in rows; per-channel stride in doubles
    int32_t detrend;
    int32_t pfsplit;            // L2 prefetch of the next chunk: every group's block its share of the rows
    double scale;               // 1 / (rate * sum(w^2))
};

template <int LOGN> struct SRCfg {
    static constexpr int T = SWCfg<LOGN>::T;
    // per-frame stride of the exchange buffer (complex): T = 32 rows of 33 (16 k1 rows read
    // by 16 lane pairs: 16-byte offsets, two wavefronts per 128-bit load)
    // T == 32: 16 rows of SR_XS doubles; the stride is even, so that a lane pair reads its row as
    // 16-byte vectors (two wavefronts per 128-bit load of 16 rows: the minimum for 256 bytes)
    // T == 8: rows of 8 padded to 9 (lane t owns rows t and t + 8 in the second pass), 16 rows
    static constexpr int FS = T == 32 ? 16 * 34 / 2 + 4 : (T == 8 ? 148 : SWCfg<LOGN>::FS);
    static constexpr int WB = FS * SWCfg<LOGN>::FPW;
};

__device__ __forceinline__ double to_db(double p) {
    return p > 1e-20 ? 10.0 * log10(p) : (p <= 1e-20 ? -INFINITY : p);
}

template <int LOGN, bool DB, bool ASYNC_FILL>
__global__ void __launch_bounds__(SR_MAXNT, SR_BLOCKS)
spectrogram_ring_kernel(const __grid_constant__ SpecRArgs P) {
    constexpr bool ASYNC = ASYNC_FILL || SR_ASYNC;
    using Cf = SWCfg<LOGN>;
    constexpr int N = Cf::N, M = Cf::M, T = Cf::T, FPW = Cf::FPW;
    constexpr int FS = SRCfg<LOGN>::FS;
    constexpr int F = M + 1;
    extern __shared__ __align__(16) double sbuf[];
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));    // read once: never rematerialised
    const int lane = tid & 31, warp = tid >> 5;
    const int NT = blockDim.x, NW = NT >> 5;
    const int C = P.C, hop = P.hop, RM = P.RC - 1, RS = P.RS;
    const int W = P.CB, LW = P.LW;                       // channels of a group = row width
    const int grp = blockIdx.x % P.ngrp;
    const int64_t f0 = (int64_t)(blockIdx.x / P.ngrp) * P.FRUN;
    // the last group of a channel count that is no multiple of W overlaps its neighbour
    const int c0 = min(grp * W, C - W);
    const int FRa = (int)min((int64_t)P.FRUN, P.nframes - f0);
    const int FSTEP = P.FSTEP, CH = FSTEP * hop;
    const int nsteps = (FRa + FSTEP - 1) / FSTEP;
    const int span0 = N + (FSTEP - 1) * hop;             // rows one step reads

    double* xs = sbuf;                                               // [W][RS]
    double* wins = sbuf + (size_t)W * RS;                            // [N] (if SR_WIN_SMEM)
    constexpr int NWIN = SR_WIN_SMEM ? (SR_WIN_HALF ? N / 2 : N) : 0;
    constexpr int NTWS = 0;                                          // split-step twiddles are built in registers
    double2* tws = reinterpret_cast<double2*>(wins + NWIN);          // [NTWS]
    double2* wb = tws + NTWS + (size_t)warp * SRCfg<LOGN>::WB;

    // the j-th 16-byte vector of this thread in a chunk: row r0 + j DR, channels col, col + 1
    // (W == 1: rows r0 + j DR and the next one).  Rows past the end of the source are
    // clamped to its last row(s): what they put into the ring is never read by a live frame.
    const int r0 = (2 * tid) >> LW, col = (2 * tid) & (W - 1);
    const int DR = (2 * NT) >> LW;
    const double* gp0 = P.src + ((f0 * hop) * (int64_t)C + c0 + col);
    const int rmax = (int)min((int64_t)0x3fffffff, P.nrows - f0 * hop - (W == 1 ? 2 : 1));
    double* sp0 = xs + col * RS;

    auto gaddr = [&](int row) {
        return reinterpret_cast<const double2*>(gp0 + (int64_t)min(row, rmax) * C);
    };
    const int64_t gstep = (int64_t)DR * C, gchunk = (int64_t)CH * C;
    const double* gnext = gp0 + (int64_t)(span0 + r0) * C;   // this thread's first vector of the next chunk
    const int64_t rows_run = P.nrows - f0 * hop;             // source rows from the start of the run
    const int rmax_async = (int)min((int64_t)0x3fffffff, rows_run - 1);
    auto put = [&](double2 v, int pos) {
        pos &= RM;
        if (W == 1) {
            *reinterpret_cast<double2*>(sp0 + pos) = v;
        } else {
            sp0[pos] = v.x;
            sp0[pos + RS] = v.y;
        }
    };
    // rows [r_start + r, ...), r = r0 + j DR < nr, j >= jbegin -> ring at pos0 + r
    auto stage_direct = [&](int r_start, int pos0, int nr, int jbegin) {
        for (int r = r0 + jbegin * DR; r < nr; r += 4 * DR) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(gaddr(r_start + r + u * DR));
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r + u * DR < nr) put(v[u], pos0 + r + u * DR);
        }
    };

    stage_direct(0, 0, span0, 0);
    if (SR_WIN_SMEM)
        for (int i = tid; i < NWIN; i += NT) wins[i] = __ldg(P.win + i);
    __syncthreads();

    const int sub = lane / T, t = lane % T;
    double2* wbf = wb + sub * FS;
    // W_M^(t k1), k1 = 1, 2, 4, 8 from the table; the other powers are products of two of them
    const double2 tw1 = __ldg(P.twA + 1 * T + t), tw2 = __ldg(P.twA + 2 * T + t);
    const double2 tw4 = __ldg(P.twA + 4 * T + t), tw8 = __ldg(P.twA + 8 * T + t);
#if SR_PAIRMAP
    const double2 twl = __ldg(P.twS + (T == 32 ? (lane >> 1) + 16 * (lane & 1) : lane % T));
#else
    const double2 twl = __ldg(P.twS + lane % T);         // W_N^t of the split step
#endif
    // SR_WIN_CALC: (cos, sin)(2 pi j / N) of this lane's two window positions j = 2 t, 2 t + 1
    // (from the table of W_N^k = exp(-2 pi i k / N) of the split step: k = 2 t + 1 < M / 8)
    double2 wcs0 = __ldg(P.twS + 2 * t), wcs1 = __ldg(P.twS + 2 * t + 1);
    wcs0.y = -wcs0.y;
    wcs1.y = -wcs1.y;
    const int nitems = FSTEP * W;
    const int niter = (nitems + NW * FPW - 1) / (NW * FPW);
    const double corr = P.detrend ? 0.5 : 0.0;           // (sum x / N) * N/2

    // all SR_PF vectors of all threads lie inside a chunk (then no load or store is predicated)
    const bool allv = (SR_PF - 1) * DR + ((2 * (NT - 1)) >> LW) < CH;
    // frames start at multiples of 64 rows in a ring of exactly one frame: the 16 loads of a
    // lane are a rotation of 16 fixed 64-row slices (immediate offsets, picked by a switch)
    const bool rot_ok = T == 32 && RM == N - 1 && (hop & 63) == 0;
    // the chunk of a step is exactly the SR_PF vectors of every thread and never wraps
    const bool fastfill = !ASYNC && W == 4 && NT == 128 && allv && CH == SR_PF * 64 && ((RM + 1) % CH) == 0 &&
                          (span0 % CH) == 0;
    int ws = 0;                 // ring position of the first row of this step
    int npos = span0;           // ring position / row of the chunk the next step adds
    int nrow = span0;
    for (int s = 0; s < nsteps; ++s) {
        // pull the rows of the next step into L2 while this step computes
        const bool more = s + 1 < nsteps;
        const bool inside = nrow + CH <= rows_run;       // the whole next chunk exists
        if (more && tid == 0 && P.pfsplit != 2) {    // whole rows (every group's block: measured faster)
            int64_t nr = inside ? CH : rows_run - nrow;
            int64_t rlo = 0;
            if (P.pfsplit == 1) {
                const int64_t per = (nr + P.ngrp - 1) / P.ngrp;
                rlo = grp * per;
                nr = nr - rlo < per ? nr - rlo : per;
            }
            if (nr > 0) {
                const double* a0 = P.src + (f0 * hop + nrow + rlo) * (int64_t)C;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0),
                             "r"((uint32_t)(nr * C * 8)) : "memory");
            }
        }

        double2 pf[SR_PF];
        bool loaded = false;
        // the loads of the next chunk are issued late in the last item of the step, when the
        // registers of the first pass are free again (they come from L2 by then)
        auto issue_loads = [&]() {
            if (SR_ABL & 2) return;
            if (inside && allv) {                   // every vector of every thread exists
                const double* gp = gnext;
#pragma unroll
                for (int j = 0; j < SR_PF; ++j) {
#if SR_LDCS
                    pf[j] = __ldcs(reinterpret_cast<const double2*>(gp));
#else
                    pf[j] = __ldg(reinterpret_cast<const double2*>(gp));
#endif
                    gp += gstep;
                }
            } else {
#pragma unroll
                for (int j = 0; j < SR_PF; ++j) {
                    pf[j] = make_double2(0.0, 0.0);
                    if (r0 + j * DR < CH) pf[j] = __ldg(gaddr(nrow + r0 + j * DR));
                }
            }
        };
        for (int iter = 0; iter < niter; ++iter) {
            const int it = (warp + NW * iter) * FPW + sub;
            int fi = it >> LW, ci = it & (W - 1);
            const bool live = it < nitems && s * FSTEP + fi < FRa;
            const bool any_live = __any_sync(0xffffffffu, live);
            if (!ASYNC && !any_live) continue;
            const bool last_iter = !ASYNC && more && iter == niter - 1;
            if (!live) { fi = 0; ci = 0; }
            const int start = ws + fi * hop + 2 * t;
            const double* xr = xs + ci * RS;

            // ---- load, window; the frame sum for the mean goes on in the background
            double2 a[16];
            if (ASYNC && !any_live) {
                // no frame for this warp: it only takes part in the hand-over of the ring
#pragma unroll
                for (int p = 0; p < 16; ++p) a[p] = make_double2(0.0, 0.0);
            } else
            if (T == 32 && rot_ok) {
                const double* xb = xr + 2 * t;
                switch (((ws + fi * hop) >> 6) & 15) {
#define ADN_ROT(Q) case Q: _Pragma("unroll") for (int p = 0; p < 16; ++p) \
                        a[p] = *reinterpret_cast<const double2*>(xb + (((p + Q) & 15) << 6)); break;
                    ADN_ROT(0) ADN_ROT(1) ADN_ROT(2) ADN_ROT(3) ADN_ROT(4) ADN_ROT(5) ADN_ROT(6) ADN_ROT(7)
                    ADN_ROT(8) ADN_ROT(9) ADN_ROT(10) ADN_ROT(11) ADN_ROT(12) ADN_ROT(13) ADN_ROT(14) ADN_ROT(15)
#undef ADN_ROT
                }
            } else {
#pragma unroll
                for (int p = 0; p < 16; ++p)
                    a[p] = *reinterpret_cast<const double2*>(xr + ((start + 2 * T * p) & RM));
            }
            if (ASYNC) {
                // every warp holds its frame: the oldest rows of the ring are free for the chunk
                // of the next step
                __syncthreads();
                if (more) {
                    const int total = CH << LW;                      // doubles of the chunk (CH rows x W)
                    for (int e = tid; e < total; e += NT) {
                        const int r = e >> LW, ch = e & (W - 1);
                        const double* g = P.src + ((f0 * hop + min(nrow + r, rmax_async)) * (int64_t)C + c0 + ch);
                        cp_async8(xs + ch * RS + ((npos + r) & RM), g);
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
                if (!any_live) continue;
            }
            double sm = 0.0;
            {
                double s0 = a[0].x + a[0].y, s1 = a[1].x + a[1].y, s2 = a[2].x + a[2].y, s3 = a[3].x + a[3].y;
#pragma unroll
                for (int p = 4; p < 16; p += 4) {
                    s0 += a[p].x + a[p].y; s1 += a[p + 1].x + a[p + 1].y;
                    s2 += a[p + 2].x + a[p + 2].y; s3 += a[p + 3].x + a[p + 3].y;
                }
                sm = (s0 + s1) + (s2 + s3);
            }
            if (SR_WIN_CALC) {
                // keeps the compiler from hoisting the 32 window values of a lane out of the frame
                // loop (they are the same for every frame: 64 registers)
                asm volatile("" : "+d"(wcs0.x), "+d"(wcs0.y), "+d"(wcs1.x), "+d"(wcs1.y));
            }
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                double2 w;
                if (SR_WIN_CALC) {
                    // 1/2 cos / sin (2 pi p / 16): W32^(2p) for p < 8, its negative beyond
                    const double hc = (p < 8 ? 0.5 : -0.5) * w32c(2 * (p & 7), 0);    // folded: p is unrolled
                    const double hs = (p < 8 ? -0.5 : 0.5) * w32c(2 * (p & 7), 1);
                    w.x = fma(-hc, wcs0.x, fma(hs, wcs0.y, 0.5));
                    w.y = fma(-hc, wcs1.x, fma(hs, wcs1.y, 0.5));
                } else if (SR_WIN_SMEM && SR_WIN_HALF) {
                    // 2 (T p + t) < N / 2 for p < 8; the second half from w[j + N/2] = 1 - w[j]
                    if (p >= 8) continue;
                    w = *reinterpret_cast<const double2*>(wins + 2 * (T * p + t));
                    a[p + 8].x = fma(-a[p + 8].x, w.x, a[p + 8].x);
                    a[p + 8].y = fma(-a[p + 8].y, w.y, a[p + 8].y);
                } else {
                    w = SR_WIN_SMEM ? *reinterpret_cast<const double2*>(wins + 2 * (T * p + t))
                                    : __ldg(reinterpret_cast<const double2*>(P.win) + (T * p + t));
                }
                a[p].x *= w.x;
                a[p].y *= w.y;
            }
            // ---- pass 1 + twiddle, into the exchange buffer at (k1, t)
            double2 b[16];
            dft16(a, b);
            {
                double2 w[16];
                w[1] = tw1; w[2] = tw2; w[4] = tw4; w[8] = tw8;
                w[3] = cmul(tw2, tw1); w[5] = cmul(tw4, tw1); w[6] = cmul(tw4, tw2);
                w[9] = cmul(tw8, tw1); w[10] = cmul(tw8, tw2); w[12] = cmul(tw8, tw4);
                w[7] = cmul(w[6], tw1); w[11] = cmul(w[10], tw1); w[13] = cmul(w[12], tw1);
                w[14] = cmul(w[12], tw2); w[15] = cmul(w[14], tw1);
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) {
                    double2 v = k1 == 0 ? b[0] : cmul(b[k1], w[k1]);
                    if (T == 32) {
                        // real parts now, imaginary parts in a second round through the same
                        // buffer of doubles (half the shared memory; the 64-bit reads of the 16
                        // lane pairs are one 128-byte wavefront each)
                        b[k1] = v;
                        reinterpret_cast<double*>(wbf)[k1 * 34 + t] = v.x;
                    } else if (T == 8) {
                        wbf[k1 * 9 + t] = v;
                    } else if (T == 4) {
                        wbf[k1 * 4 + ((t + k1) & 3)] = v;       // rows of 4, rotated by the row index
                    } else {
                        const int f = k1 * T + t;
                        wbf[f + (f >> 4)] = v;
                    }
                }
            }
#pragma unroll
            for (int o = 1; o < T; o <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
            __syncwarp();
            // ---- pass 2: T-point DFTs over t, split step and power
            double* out = P.dst + (((f0 + s * FSTEP + fi) * (int64_t)C + c0 + ci) * F);
            const double sc = 0.5 * P.scale;
            const double mN2 = sm * corr;                  // mean * N/2
            if constexpr (T == 32) {
                // lane (k1, tp): outputs k2 = 2 kk + tp of the 32-point DFT of row k1, i.e.
                // Z[lane + 32 kk], kk < 16 (first radix-2 step done here, decimation in frequency)
#if SR_PAIRMAP
                // the two lanes of a row are neighbours (same quarter warp: their identical 16-byte
                // reads are one access); kl = the lane's place in frequency order
                const int k1 = lane >> 1, tp = lane & 1;
                const int kl = k1 + 16 * tp;
#else
                const int k1 = lane & 15, tp = lane >> 4;
                const int kl = lane;
#endif
                const double sgn = tp ? -1.0 : 1.0;
                double* wre = reinterpret_cast<double*>(wbf);
                const double2* row = reinterpret_cast<const double2*>(wre + k1 * 34);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const double2 lo = row[m], hi = row[m + 8];
                    a[2 * m].x = fma(sgn, hi.x, lo.x);
                    a[2 * m + 1].x = fma(sgn, hi.y, lo.y);
                }
                __syncwarp();
#pragma unroll
                for (int kq = 0; kq < 16; ++kq) wre[kq * 34 + t] = b[kq].y;
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const double2 lo = row[m], hi = row[m + 8];
                    a[2 * m].y = fma(sgn, hi.x, lo.x);
                    a[2 * m + 1].y = fma(sgn, hi.y, lo.y);
                }
                if (tp) {
#pragma unroll
                    for (int n = 1; n < 16; ++n) a[n] = cmul(a[n], make_double2(w32c(n, 0), w32c(n, 1)));
                }
                dft16(a, b);
                if (last_iter) { issue_loads(); loaded = true; }
                // split step in registers: Z[M - k] of k = lane + 32 kk sits in lane 32 - lane,
                // register 15 - kk (lane 0: its own register 16 - kk), so a lane handles the
                // pairs of its registers kk < 8 and fetches the partner by shuffle
#if SR_PAIRMAP
                const int kp = (32 - kl) & 31;
                const int partner = 2 * (kp & 15) + (kp >> 4);
#else
                const int partner = (32 - lane) & 31;
#endif
                double* outk = out + kl;
                double* outm = out + M - kl;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    double2 zk = b[kk], zm;
                    zm.x = __shfl_sync(0xffffffffu, b[15 - kk].x, partner);
                    zm.y = __shfl_sync(0xffffffffu, b[15 - kk].y, partner);
                    if (kk > 0 && kl == 0) zm = b[16 - kk];
                    double2 tw = kk == 0 ? twl : cmul(twl, make_double2(w32c(kk, 0), w32c(kk, 1)));
                    double e_r = zk.x + zm.x, e_i = zk.y - zm.y;          // Zk + conj(Zm)
                    double o_r = zk.y + zm.y, o_i = zm.x - zk.x;          // -i (Zk - conj(Zm))
                    double t_r = o_r * tw.x - o_i * tw.y, t_i = o_r * tw.y + o_i * tw.x;
                    double pr = e_r + t_r, pi = e_i + t_i, qr = e_r - t_r, qi = e_i - t_i;
                    // p = 2 X[k]; the window's spectrum at bin 1 is -N/4
                    if (kk == 0) pr += kl == 1 ? mN2 : 0.0;
                    double pk = (pr * pr + pi * pi) * sc, pm = (qr * qr + qi * qi) * sc;
                    if (kk == 0 && kl == 0) {                             // bins 0 and M from Z[0]
                        double x0 = zk.x + zk.y - mN2, xM = zk.x - zk.y;
                        pk = x0 * x0 * P.scale;
                        pm = xM * xM * P.scale;
                    }
                    if (DB) { pk = to_db(pk); pm = to_db(pm); }
                    if (live && SR_ST_OK(pk)) {
                        ADN_STORE(outk + 32 * kk, pk);
                        ADN_STORE(outm - 32 * kk, pm);
                    }
                }
                if (kl == 0 && live) {                                    // k = M/2 pairs with itself
                    double2 zk = b[8];
                    double2 tw = make_double2(w32c(8, 0), w32c(8, 1));    // W_N^(M/2) = -i
                    double e_r = zk.x + zk.x, o_r = zk.y + zk.y;
                    double t_r = o_r * tw.x, t_i = o_r * tw.y;
                    double pr = e_r + t_r, pi = t_i;
                    double pk = (pr * pr + pi * pi) * sc;
                    if (DB) pk = to_db(pk);
                    out[M / 2] = pk;
                }
            } else {
                // T == 8: lane t transforms rows t and t + 8 (rows of 9: the lanes of a frame read eight
                // different 16-byte bank groups).  T == 4: rows t + 4 r, rows of 4 rotated by the row
                // index (no padding: the two frames of a quarter warp sit FS == 4 mod 8 apart)
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    a[i] = T == 8 ? wbf[(t + 8 * (i >> 3)) * 9 + (i & 7)]
                         : T == 4 ? wbf[(t + 4 * (i >> 2)) * 4 + ((t + i) & 3)] : wbf[t * 17 + i];
                if (T == 16) {
                    dft16(a, b);
                } else if (T == 8) {
                    dft8(&a[0], &b[0]);
                    dft8(&a[8], &b[8]);
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        dft4(a[4 * r], a[4 * r + 1], a[4 * r + 2], a[4 * r + 3],
                             b[4 * r], b[4 * r + 1], b[4 * r + 2], b[4 * r + 3]);
                }
                if (last_iter) { issue_loads(); loaded = true; }
                // split step in registers, as for T == 32: lane t of a frame holds Z[t + T m], m < 16
                // (row t + T r, output k2 of its DFT: m = r + (16 / T) k2), so Z[M - k] of k = t + T m
                // sits in lane (T - t) % T of the frame, at m' = 15 - m (lane 0: its own m' = 16 - m);
                // a lane handles the pairs of m < 8 and fetches the partner by shuffle.  The twiddle
                // W_N^(t + T m) is W_N^t W_32^m.  (Until round 2 this path wrote Z back to the exchange
                // buffer in natural order and read pairs and twiddles from shared memory: 40 of the
                // 104 shared-memory wavefronts of a 256-point frame.)
#define ADN_RG(m) (T == 16 ? (m) : (T == 8 ? 8 * ((m) & 1) + ((m) >> 1) : 4 * ((m) & 3) + ((m) >> 2)))
                const int partner = (lane & ~(T - 1)) | ((T - t) & (T - 1));
                double* outk = out + t;
                double* outm = out + M - t;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    double2 zk = b[ADN_RG(kk)], zm;
                    zm.x = __shfl_sync(0xffffffffu, b[ADN_RG(15 - kk)].x, partner);
                    zm.y = __shfl_sync(0xffffffffu, b[ADN_RG(15 - kk)].y, partner);
                    if (kk > 0 && t == 0) zm = b[ADN_RG(16 - kk)];
                    double2 tw = kk == 0 ? twl : cmul(twl, make_double2(w32c(kk, 0), w32c(kk, 1)));
                    double e_r = zk.x + zm.x, e_i = zk.y - zm.y;          // Zk + conj(Zm)
                    double o_r = zk.y + zm.y, o_i = zm.x - zk.x;          // -i (Zk - conj(Zm))
                    double t_r = o_r * tw.x - o_i * tw.y, t_i = o_r * tw.y + o_i * tw.x;
                    double pr = e_r + t_r, pi = e_i + t_i, qr = e_r - t_r, qi = e_i - t_i;
                    if (kk == 0) pr += t == 1 ? mN2 : 0.0;                // the window's spectrum at bin 1 is -N/4
                    double pk = (pr * pr + pi * pi) * sc, pm = (qr * qr + qi * qi) * sc;
                    if (kk == 0 && t == 0) {                              // bins 0 and M from Z[0]
                        double x0 = zk.x + zk.y - mN2, xM = zk.x - zk.y;
                        pk = x0 * x0 * P.scale;
                        pm = xM * xM * P.scale;
                    }
                    if (DB) { pk = to_db(pk); pm = to_db(pm); }
                    if (live) {
                        // evict-first here: measured on B200 (64 ch x 250 kHz, nfft 128 .. 512, overlap
                        // <= 50 %) the default policy is 8 - 13 % slower on this path
                        __stcs(outk + T * kk, pk);
                        __stcs(outm - T * kk, pm);
                    }
                }
                if (t == 0 && live) {                                     // k = M/2 pairs with itself
                    double2 zk = b[ADN_RG(8)];
                    double2 tw = make_double2(w32c(8, 0), w32c(8, 1));    // W_N^(M/2) = -i
                    double e_r = zk.x + zk.x, o_r = zk.y + zk.y;
                    double t_r = o_r * tw.x, t_i = o_r * tw.y;
                    double pr = e_r + t_r, pi = t_i;
                    double pk = (pr * pr + pi * pi) * sc;
                    if (DB) pk = to_db(pk);
                    out[M / 2] = pk;
                }
#undef ADN_RG
            }
            __syncwarp();
        }

        if (ASYNC) {
            if (more) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
            }
        } else
        if (more) {
            if (!loaded) issue_loads();     // warps without an item in the last iteration
            if (!(SR_ABL & 4)) __syncthreads();            // every warp is done with the rows the chunk replaces
            if (SR_ABL & 2) {
            } else
            if (fastfill) {
                // groups of four channels, 128 threads, chunks that never wrap around the ring:
                // the SR_PF vectors of a thread land 64 rows apart, two stores each at literal offsets
                double* sp = sp0 + (npos & RM) + r0;
                double* sq = sp + RS;
#pragma unroll
                for (int j = 0; j < SR_PF; ++j) {
                    sp[j * 64] = pf[j].x;
                    sq[j * 64] = pf[j].y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < SR_PF; ++j)
                    if (allv || r0 + j * DR < CH) put(pf[j], npos + r0 + j * DR);
                if (r0 + SR_PF * DR < CH) stage_direct(nrow, npos, CH, SR_PF);
            }
            if (!(SR_ABL & 4)) __syncthreads();
        }
        ws = (ws + CH) & RM;
        npos = (npos + CH) & RM;
        nrow += CH;
        gnext += gchunk;
    }
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

// returns ADN_ERR_UNSUPPORTED (without an error message) when the shape does not fit
template <int LOGN, bool DB, bool ASYNC>
int32_t launch_ring_kernel(SpecRArgs& P, int64_t nf, cudaStream_t st) {
    using Cf = SWCfg<LOGN>;
    const int C = P.C, hop = P.hop;
    int NW = env_int("ADN_SPEC_NW", 4);
    if (NW < 1) NW = 1;
    if (NW > SR_MAXNT / 32) NW = SR_MAXNT / 32;
    int CBmax = env_int("ADN_SPEC_CB", 4);
    P.LW = 0;
    while ((2 << P.LW) <= C && (2 << P.LW) <= CBmax) ++P.LW;
    P.CB = 1 << P.LW;
    P.ngrp = (C + P.CB - 1) / P.CB;
    P.nrows = (nf - 1) * hop + Cf::N;
    const int items = NW * Cf::FPW;
    P.FSTEP = items / P.CB > 1 ? items / P.CB : 1;
    const size_t limit = 227 * 1024 - 1024;
    size_t smem = 0;
    for (;; P.FSTEP = (P.FSTEP + 1) / 2) {
        const int64_t CH = (int64_t)P.FSTEP * hop, span0 = Cf::N + (int64_t)(P.FSTEP - 1) * hop;
        (void)CH;
        P.RC = Cf::N;
        while (P.RC < span0) P.RC <<= 1;                   // power of two >= the rows of a step
        P.RS = P.RC + 4;                                   // == 4 mod 8
        smem = ((size_t)P.CB * P.RS + (SR_WIN_SMEM ? (SR_WIN_HALF ? Cf::N / 2 : Cf::N) : 0)) * 8 +
               ((Cf::T == 32 ? 0 : (size_t)Cf::M / 2 + 2) + (size_t)NW * SRCfg<LOGN>::WB) * 16;
        if (smem <= limit || P.FSTEP == 1) break;
    }
    if (smem > limit) return ADN_ERR_UNSUPPORTED;
    if ((ASYNC || SR_ASYNC) && P.FSTEP * P.CB > items) return ADN_ERR_UNSUPPORTED;    // one item per warp and step
    auto kern = spectrogram_ring_kernel<LOGN, DB, ASYNC>;
    static bool attr_done = false;
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    int bps = 1;
    ADN_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, NW * 32, smem));
    if (bps < 1) return ADN_ERR_UNSUPPORTED;
    // one resident wave of blocks, runs of at least 4 steps
    int64_t nruns = (int64_t)ctx().sm_count * bps / P.ngrp;
    if (nruns < 1) nruns = 1;
    int64_t frun = (nf + nruns - 1) / nruns;
    if (frun < 4 * P.FSTEP) frun = 4 * P.FSTEP;
    frun = (frun + P.FSTEP - 1) / P.FSTEP * P.FSTEP;
    if (frun * hop + Cf::N > 0x3fffffff) return ADN_ERR_UNSUPPORTED;
    P.FRUN = (int32_t)frun;
    const int64_t grid = ((nf + frun - 1) / frun) * P.ngrp;
    if (grid > 0x7fffffff) return ADN_ERR_UNSUPPORTED;
    kern<<<(unsigned)grid, NW * 32, smem, st>>>(P);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

template <int LOGN>
int32_t launch_ring(SpecRArgs& P, int64_t nf, int out_db, cudaStream_t st) {
    if (out_db) return launch_ring_kernel<LOGN, true, false>(P, nf, st);
    // ring fill by cp.async (the copies land while the transforms run) or through registers (8
    // vectors per thread, loaded late in the step): the choice is by shape, see SR_ASYNC above
    int async = env_int("ADN_SPEC_ASYNC", -1);
    // measured on B200 (tools/kbench.py, 8 and 64 channels): async wins 1 - 9 % for every shape but
    // nfft 1024 with hop >= 512 (-3 %) and one or two channels (-3 %)
    if (async < 0) async = P.C >= 4 && (LOGN < 10 || P.hop < 512) ? 1 : 0;
    if (async) {
        int32_t rc = launch_ring_kernel<LOGN, false, true>(P, nf, st);
        if (rc != ADN_ERR_UNSUPPORTED) return rc;
    }
    return launch_ring_kernel<LOGN, false, false>(P, nf, st);
}

template <int LOGN>
int32_t launch_warp_kernel(SpecWArgs& P, int64_t nf, cudaStream_t st) {
    using Cf = SWCfg<LOGN>;
    const int C = P.C;
    // channels per block: 2 when they pair up (16-byte sectors shared by neighbour blocks in L2)
    P.CB = C >= 2 ? 2 : 1;
    int rows_budget = 1024 + Cf::N / 2;
    int FB = (rows_budget - Cf::N) / P.hop + 1;
    if (FB < 1) FB = 1;
    if (FB > 64) FB = 64;
    if ((int64_t)FB > nf) FB = (int)nf;
    P.FB = FB;
    int rows = (FB - 1) * P.hop + Cf::N;
    P.RP = ((rows + 15) / 16) * 16 + 8;               // == 8 mod 16: channel arrays on disjoint banks
    const size_t smem = (size_t)P.CB * P.RP * 8 + (size_t)SW_NWARP * Cf::WB * 16;
    auto kern = spectrogram_warp_kernel<LOGN>;
    static bool attr_done = false;
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    const int ngrp = (C + P.CB - 1) / P.CB;
    int64_t grid = ((nf + FB - 1) / FB) * ngrp;
    if (grid > 0x7fffffff) return fail(ADN_ERR_UNSUPPORTED, "spectrogram: grid %lld", (long long)grid);
    kern<<<(unsigned)grid, SW_NT, smem, st>>>(P);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}


// ======================================================================================
// nfft 2048 .. 16384: one block per frame and channel pair, R complex points per thread in
// registers, Stockham autosort passes (radix R, then the remaining factor) with the exchanges
// through one padded shared-memory buffer.  Pass with accumulated length Ns, radix r,
// butterfly b < M/r:   y[(b / Ns) Ns r + b % Ns + m Ns] = sum_k W_(Ns r)^((b % Ns) k) x[b + k M/r] W_r^(k m)
// Natural order in, natural order out, so the split step reads Z[k] and Z[M - k] directly.
// The raw rows of both channels of a pair come in as 16-byte loads and stay in registers
// while the channels are transformed one after the other; the frame mean is removed after
// the transform (bins 0 and 1, see the ring kernel).
struct SpecMArgs {
    const double* src;
    double* dst;
    const double* win;          // nfft
    const double2* tw;          // exp(-2 pi i j / nfft), j < nfft/2
    int64_t nframes;
    int32_t C, hop, npair;      // channel pairs (the last one may be a single channel)
    int32_t detrend;
    double scale;
};

template <int LOGN>
__device__ __forceinline__ double2 tw_full(const double2* __restrict__ tw, int n) {
    // exp(-2 pi i n / N) for any n from the half table
    constexpr int N = 1 << LOGN;
    n &= N - 1;
    double2 v = __ldg(tw + (n & (N / 2 - 1)));
    return n >= N / 2 ? make_double2(-v.x, -v.y) : v;
}

__device__ __forceinline__ int mp_pad(int i) { return i + (i >> 3); }

template <int r>
__device__ __forceinline__ void mp_dft(const double2* x, double2* y) {
    if (r == 2) {
        y[0] = cadd(x[0], x[1]);
        y[1] = csub(x[0], x[1]);
    } else if (r == 4) {
        dft4(x[0], x[1], x[2], x[3], y[0], y[1], y[2], y[3]);
    } else if (r == 8) {
        dft8(x, y);
    } else {
        double2 a[16], b[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = x[i];
        dft16(a, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = b[i];
    }
}

// one pass over the R points of this thread: R/r butterflies of radix r
// wb: exp(-2 pi i (t mod Ns) / (Ns r)) of this thread (Ns <= T), or for the last pass with
// Ns > T exp(-2 pi i t / (Ns r)): the twiddle of butterfly b = t + q T is then wb W_(Ns r / T)^q
template <int LOGN, int R, int r, bool FROM_REGS, bool FIRST>
__device__ __forceinline__ void mp_pass(double2 (&z)[R], double2* S, double2 wb, int t, int logNs) {
    constexpr int M = 1 << (LOGN - 1), T = M / R, NB = M / r, Q = R / r;
    constexpr int LOGR = r == 2 ? 1 : (r == 4 ? 2 : (r == 8 ? 3 : 4));
    if (!FROM_REGS) {
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int k = 0; k < r; ++k) z[q * r + k] = S[mp_pad(t + q * T + k * NB)];
        __syncthreads();
    }
    const int Ns = 1 << logNs;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int b = t + q * T;
        const int bl = b & (Ns - 1);
        double2 x[r], y[r];
#pragma unroll
        for (int k = 0; k < r; ++k) x[k] = z[q * r + k];
        if (!FIRST) {
            // W_(Ns r)^(bl k), k < r: powers of the thread's base twiddle
            double2 w[r];
            w[1] = wb;
            if (Q > 1 && q > 0) {
                // bl = t + q T: times exp(-2 pi i q T / (Ns r)), a power of W_(2 Q) or finer
                constexpr int LT = LOGN - 1 - (R == 8 ? 3 : 4);       // log2 T
                const int sh = logNs + LOGR - LT;                      // Ns r / T = 2^sh
                // angle = -2 pi q / 2^sh, sh <= 5: from the 32nd roots of unity
                const int idx = (q << (5 - sh)) & 31;
                double2 c = idx < 16 ? make_double2(w32c(idx & 15, 0), w32c(idx & 15, 1))
                                     : make_double2(-w32c(idx & 15, 0), -w32c(idx & 15, 1));
                w[1] = cmul(wb, c);
            }
            if (r > 2) w[2] = cmul(w[1], w[1]);
            if (r > 4) w[4] = cmul(w[2], w[2]);
            if (r > 8) w[8] = cmul(w[4], w[4]);
            if (r > 2) w[3] = cmul(w[2], w[1]);
            if (r > 4) { w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[6], w[1]); }
            if (r > 8) {
                w[9] = cmul(w[8], w[1]); w[10] = cmul(w[8], w[2]); w[12] = cmul(w[8], w[4]);
                w[11] = cmul(w[10], w[1]); w[13] = cmul(w[12], w[1]); w[14] = cmul(w[12], w[2]);
                w[15] = cmul(w[14], w[1]);
            }
#pragma unroll
            for (int k = 1; k < r; ++k) x[k] = cmul(x[k], w[k]);
        }
        mp_dft<r>(x, y);
        const int base = ((b - bl) << LOGR) + bl;
#pragma unroll
        for (int m = 0; m < r; ++m) S[mp_pad(base + (m << logNs))] = y[m];
    }
    __syncthreads();
}

template <int LOGN, int R>
__device__ __forceinline__ void mp_fft(double2 (&z)[R], double2* S, const double2 (&wb)[5], int t) {
    constexpr int LOGM = LOGN - 1;
    constexpr int LOGR = R == 8 ? 3 : 4;
    constexpr int NFULL = LOGM / LOGR;                 // passes of radix R
    constexpr int REM = LOGM - NFULL * LOGR;           // log2 of the last, smaller radix
    mp_pass<LOGN, R, R, true, true>(z, S, make_double2(1.0, 0.0), t, 0);
#pragma unroll
    for (int p = 1; p < NFULL; ++p) mp_pass<LOGN, R, R, false, false>(z, S, wb[p], t, p * LOGR);
    if (REM == 1) mp_pass<LOGN, R, 2, false, false>(z, S, wb[NFULL], t, NFULL * LOGR);
    if (REM == 2) mp_pass<LOGN, R, 4, false, false>(z, S, wb[NFULL], t, NFULL * LOGR);
    if (REM == 3) mp_pass<LOGN, R, 8, false, false>(z, S, wb[NFULL], t, NFULL * LOGR);
}

// base twiddles of thread t for the passes 1 .. (wb[0] unused): exp(-2 pi i (t mod Ns) / (Ns r))
template <int LOGN, int R>
__device__ __forceinline__ void mp_twiddles(double2 (&wb)[5], const double2* __restrict__ tw, int t) {
    constexpr int LOGM = LOGN - 1;
    constexpr int LOGR = R == 8 ? 3 : 4;
    constexpr int NFULL = LOGM / LOGR;
    constexpr int REM = LOGM - NFULL * LOGR;
    constexpr int T = (1 << LOGM) / R;
    wb[0] = make_double2(1.0, 0.0);
#pragma unroll
    for (int p = 1; p < 5; ++p) {
        wb[p] = make_double2(1.0, 0.0);
        if (p < NFULL || (p == NFULL && REM > 0)) {
            const int logNs = p * LOGR;
            const int logr = p < NFULL ? LOGR : REM;
            const int Ns = 1 << logNs;
            const int bl = Ns <= T ? (t & (Ns - 1)) : t;
            wb[p] = tw_full<LOGN>(tw, bl << (LOGN - logNs - logr));
        }
    }
}

// transform of one channel-frame held in za (windowed), split step, power, store of its F bins;
// red: per-warp partial sums of the raw frame (visible after the first barrier of the transform)
template <int LOGN, int R, bool DB>
__device__ __forceinline__ void mp_channel(double2 (&za)[R], double2* S, const double2 (&wb)[5], int t,
                                           const double* red, int detrend, double scale,
                                           const double2* __restrict__ tw, double* __restrict__ out) {
    constexpr int N = 1 << LOGN, M = N / 2, T = M / R;
    const double sc = 0.5 * scale;
    mp_fft<LOGN, R>(za, S, wb, t);                // ends with a barrier: red[] is visible too
    double fsum = 0.0;
    for (int i = 0; i < T / 32; ++i) fsum += red[i];
    const double mN2 = detrend ? fsum * 0.5 : 0.0;          // mean * N/2
#pragma unroll
    for (int q = 0; q < R / 2; ++q) {
        const int k = 1 + t + q * T, km = M - k;              // k in [1, M/2]
        double2 zk = S[mp_pad(k)], zm = S[mp_pad(km)];
        double2 w = __ldg(tw + k);
        double e_r = zk.x + zm.x, e_i = zk.y - zm.y;          // Zk + conj(Zm)
        double o_r = zk.y + zm.y, o_i = zm.x - zk.x;          // -i (Zk - conj(Zm))
        double t_r = o_r * w.x - o_i * w.y, t_i = o_r * w.y + o_i * w.x;
        double pr = e_r + t_r, pi = e_i + t_i, qr = e_r - t_r, qi = e_i - t_i;
        if (k == 1) pr += mN2;                                // the window's spectrum at bin 1 is -N/4
        double pk = (pr * pr + pi * pi) * sc, pm = (qr * qr + qi * qi) * sc;
        if (DB) { pk = to_db(pk); pm = to_db(pm); }
        ADN_STORE(out + k, pk);
        if (km != k) ADN_STORE(out + km, pm);
    }
    if (t == 0) {
        double2 z0 = S[0];
        double x0v = z0.x + z0.y - mN2, xM = z0.x - z0.y;
        double p0 = x0v * x0v * scale, pM = xM * xM * scale;
        if (DB) { p0 = to_db(p0); pM = to_db(pM); }
        out[0] = p0;
        out[M] = pM;
    }
}

template <int LOGN, int R, int CP, bool DB>
__global__ void __launch_bounds__((1 << (LOGN - 1)) / R)
spectrogram_mp_kernel(const __grid_constant__ SpecMArgs P) {
    constexpr int N = 1 << LOGN, M = N / 2, T = M / R, F = M + 1;
    extern __shared__ __align__(16) double sbuf[];
    double2* S = reinterpret_cast<double2*>(sbuf);               // M + M/8 complex
    __shared__ double red[2][32];
    const int t = threadIdx.x;
    const int64_t frame = blockIdx.x / P.npair;
    const int pair = blockIdx.x % P.npair;
    const int C = P.C;
    const int c0 = pair * CP;
    const int nch = min(CP, C - c0);
    const double* x0 = P.src + (frame * P.hop) * (int64_t)C + c0;

    double2 za[R], zb[CP == 2 ? R : 1];
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int64_t row = 2 * (t + k * T);
        if (CP == 2 && nch == 2) {
            double2 v0 = __ldg(reinterpret_cast<const double2*>(x0 + row * C));
            double2 v1 = __ldg(reinterpret_cast<const double2*>(x0 + (row + 1) * C));
            za[k] = make_double2(v0.x, v1.x);
            zb[CP == 2 ? k : 0] = make_double2(v0.y, v1.y);
            sb += v0.y + v1.y;
        } else {
            za[k] = make_double2(__ldg(x0 + row * C), __ldg(x0 + (row + 1) * C));
        }
        sa += za[k].x + za[k].y;
    }
    // frame sums (for the mean): warp shuffle, then one value per warp through shared memory
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if ((t & 31) == 0) { red[0][t >> 5] = sa; red[1][t >> 5] = sb; }
#pragma unroll
    for (int k = 0; k < R; ++k) {
        double2 w = __ldg(reinterpret_cast<const double2*>(P.win) + (t + k * T));
        za[k].x *= w.x; za[k].y *= w.y;
        if (CP == 2) { zb[CP == 2 ? k : 0].x *= w.x; zb[CP == 2 ? k : 0].y *= w.y; }
    }
    double2 wb[5];
    mp_twiddles<LOGN, R>(wb, P.tw, t);
    for (int ch = 0; ch < nch; ++ch) {
        if (CP == 2 && ch == 1) {
#pragma unroll
            for (int k = 0; k < R; ++k) za[k] = zb[CP == 2 ? k : 0];
        }
        mp_channel<LOGN, R, DB>(za, S, wb, t, red[ch], P.detrend, P.scale, P.tw,
                                P.dst + ((frame * C + c0 + ch) * (int64_t)F));
        __syncthreads();                              // S is reused by the second channel
    }
}

// Overlapping frames (hop < nfft): a block owns a run of consecutive frames of a channel pair and
// keeps their rows de-interleaved in a shared-memory ring of exactly one frame; every step only
// the `hop` new rows are loaded (16-byte loads of both channels) and overwrite the oldest ones
// between two barriers.  Each input row is then read once per run instead of nfft/hop times --
// the row-strided loads are what bounds the frame-per-block kernel above.
template <int LOGN, int R, bool DB>
__global__ void __launch_bounds__((1 << (LOGN - 1)) / R)
spectrogram_mpr_kernel(const __grid_constant__ SpecMArgs P, int32_t frun) {
    constexpr int N = 1 << LOGN, M = N / 2, T = M / R, F = M + 1;
    constexpr int RS = N + 4;                                    // channel arrays 4 doubles apart
    extern __shared__ __align__(16) double sbuf[];
    double2* S = reinterpret_cast<double2*>(sbuf);               // M + M/8 (+8) complex
    double* xs = sbuf + 2 * (M + M / 8 + 8);                     // [2][RS]
    __shared__ double red[2][32];
    const int t = threadIdx.x;
    const int pair = blockIdx.x % P.npair;
    const int64_t f0 = (int64_t)(blockIdx.x / P.npair) * frun;
    const int nfr = (int)min((int64_t)frun, P.nframes - f0);
    const int C = P.C, hop = P.hop;
    const int c0 = 2 * pair;
    const double* x0 = P.src + (f0 * hop) * (int64_t)C + c0;

    auto stage = [&](int r_first, int nrows) {                   // rows r_first .. of the run -> ring
        for (int r = t; r < nrows; r += T) {
            double2 v = __ldg(reinterpret_cast<const double2*>(x0 + (int64_t)(r_first + r) * C));
            const int pos = (r_first + r) & (N - 1);
            xs[pos] = v.x;
            xs[RS + pos] = v.y;
        }
    };
    stage(0, N);
    double2 wb[5];
    mp_twiddles<LOGN, R>(wb, P.tw, t);
    __syncthreads();

    for (int s = 0; s < nfr; ++s) {
        if (t == 0 && s + 1 < nfr) {                             // next rows on their way to L2
            const double* a0 = P.src + ((f0 + s) * hop + N) * (int64_t)C;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0),
                         "r"((uint32_t)(hop * C * 8)) : "memory");
        }
        const int start = (s * hop) & (N - 1);
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
            const double* xc = xs + ch * RS;
            double2 za[R];
            double sa = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                za[k] = *reinterpret_cast<const double2*>(xc + ((start + 2 * (t + k * T)) & (N - 1)));
                sa += za[k].x + za[k].y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o);
            if ((t & 31) == 0) red[ch][t >> 5] = sa;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                double2 w = __ldg(reinterpret_cast<const double2*>(P.win) + (t + k * T));
                za[k].x *= w.x;
                za[k].y *= w.y;
            }
            mp_channel<LOGN, R, DB>(za, S, wb, t, red[ch], P.detrend, P.scale, P.tw,
                                    P.dst + (((f0 + s) * C + c0 + ch) * (int64_t)F));
            __syncthreads();                                     // S (and red) free again
        }
        if (s + 1 < nfr) {
            stage(s * hop + N, hop);                             // replaces rows [s hop, (s+1) hop)
            __syncthreads();
        }
    }
}

// Frame-per-block kernel with the rows of the channel pair gathered by the TMA unit (nfft 2048, 4096).
// The frame-per-block kernel above is bound by the L1 data pipe (ncu, 64 ch x 250 kHz: 85 %): a block
// needs 16 bytes of every row, so each of its 128-bit global loads touches 32 lines = 32 wavefronts,
// as many as the shared-memory passes of the transform.  Here one thread issues N / 256 tensor
// copies (a 2-D tensor map over the source: box = 2 channels x 256 rows, 4 KB each) that complete on
// an mbarrier; the rows arrive in shared memory as [row][2] without passing through the L1 data
// pipe, the threads read them as 128-bit vectors, and the exchange buffer of the transform reuses
// the space.
constexpr int MPT_ROWS = 256;

__device__ __forceinline__ uint32_t sp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int LOGN, bool DB>
__global__ void __launch_bounds__((1 << (LOGN - 1)) / 8, LOGN == 11 ? 3 : 2)
spectrogram_mpt_kernel(const __grid_constant__ SpecMArgs P, const __grid_constant__ CUtensorMap tmap) {
    constexpr int R = 8;
    constexpr int N = 1 << LOGN, M = N / 2, T = M / R, F = M + 1;
    extern __shared__ __align__(16) double sbuf[];
    double* sb = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sbuf) + 127) & ~(uintptr_t)127);
    double2* S = reinterpret_cast<double2*>(sb);                   // M + M/8 complex, after the rows are read
    const double2* raw = reinterpret_cast<const double2*>(sb);     // [N]: (channel c0, c0 + 1) of a row
    __shared__ double red[2][32];
    __shared__ __align__(8) uint64_t mbar;
    const int t = threadIdx.x;
    const int64_t frame = blockIdx.x / P.npair;
    const int pair = blockIdx.x % P.npair;
    const int C = P.C;
    const int c0 = pair * 2;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_u32(&mbar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
        asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}"
                     ::"r"(sp_smem_u32(&mbar)), "r"((uint32_t)(N * 16)) : "memory");
        const int row0 = (int)(frame * P.hop);
#pragma unroll 1
        for (int j = 0; j < N / MPT_ROWS; ++j)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
                         "[%0], [%1, {%2, %3}], [%4];"
                         ::"r"(sp_smem_u32(sb + (size_t)j * MPT_ROWS * 2)), "l"(reinterpret_cast<uint64_t>(&tmap)),
                           "r"(c0), "r"(row0 + j * MPT_ROWS), "r"(sp_smem_u32(&mbar)) : "memory");
    }
    double2 wb[5];
    mp_twiddles<LOGN, R>(wb, P.tw, t);
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
        ::"r"(sp_smem_u32(&mbar)), "r"(0) : "memory");

    double2 za[R], zb[R];
    double sa = 0.0, sbsum = 0.0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int row = 2 * (t + k * T);
        const double2 v0 = raw[row], v1 = raw[row + 1];
        za[k] = make_double2(v0.x, v1.x);
        zb[k] = make_double2(v0.y, v1.y);
        sa += v0.x + v1.x;
        sbsum += v0.y + v1.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sbsum += __shfl_xor_sync(0xffffffffu, sbsum, o);
    }
    if ((t & 31) == 0) { red[0][t >> 5] = sa; red[1][t >> 5] = sbsum; }
#pragma unroll
    for (int k = 0; k < R; ++k) {
        double2 w = __ldg(reinterpret_cast<const double2*>(P.win) + (t + k * T));
        za[k].x *= w.x; za[k].y *= w.y;
        zb[k].x *= w.x; zb[k].y *= w.y;
    }
    __syncthreads();                                  // every row has been read: S takes the place
    for (int ch = 0; ch < 2; ++ch) {
        if (ch == 1) {
#pragma unroll
            for (int k = 0; k < R; ++k) za[k] = zb[k];
        }
        mp_channel<LOGN, R, DB>(za, S, wb, t, red[ch], P.detrend, P.scale, P.tw,
                                P.dst + ((frame * C + c0 + ch) * (int64_t)F));
        __syncthreads();                              // S is reused by the second channel
    }
}

// Overlapping frames with the TMA unit: a block owns a run of consecutive frames of a channel pair
// (as spectrogram_mpr_kernel) and keeps one frame of rows in a shared-memory ring, here as [row][2]
// exactly as the tensor copies deliver them.  Every step waits for the rows of its frame (mbarrier,
// one phase per step), reads both channels into registers, and -- one barrier later, when the
// oldest `hop` rows are free -- one thread issues the tensor copies of the next step's rows, which
// land while the two transforms run.  No thread loads a row, no register stages one.
template <int LOGN, bool DB>
__global__ void __launch_bounds__((1 << (LOGN - 1)) / 8, LOGN == 11 ? 3 : 2)
spectrogram_mprt_kernel(const __grid_constant__ SpecMArgs P, const __grid_constant__ CUtensorMap tmap, int32_t frun) {
    constexpr int R = 8;
    constexpr int N = 1 << LOGN, M = N / 2, T = M / R, F = M + 1;
    extern __shared__ __align__(16) double sbuf[];
    double* sb = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sbuf) + 127) & ~(uintptr_t)127);
    const double2* ring = reinterpret_cast<const double2*>(sb);    // [N]: (channel c0, c0 + 1) of a row
    double2* S = reinterpret_cast<double2*>(sb + 2 * N);           // M + M/8 (+8) complex
    __shared__ double red[2][32];
    __shared__ __align__(8) uint64_t mbar;
    const int t = threadIdx.x;
    const int pair = blockIdx.x % P.npair;
    const int64_t f0 = (int64_t)(blockIdx.x / P.npair) * frun;
    const int nfr = (int)min((int64_t)frun, P.nframes - f0);
    const int C = P.C, hop = P.hop;
    const int c0 = 2 * pair;
    const int row_run = (int)(f0 * hop);                           // first source row of the run
    auto copy_rows = [&](int row_first, int nrows) {               // source rows -> ring at row_first mod N
        asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}"
                     ::"r"(sp_smem_u32(&mbar)), "r"((uint32_t)(nrows * 16)) : "memory");
#pragma unroll 1
        for (int j = 0; j < nrows; j += MPT_ROWS)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
                         "[%0], [%1, {%2, %3}], [%4];"
                         ::"r"(sp_smem_u32(sb + (size_t)((row_first + j) & (N - 1)) * 2)),
                           "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(row_run + row_first + j),
                           "r"(sp_smem_u32(&mbar)) : "memory");
    };
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_u32(&mbar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) copy_rows(0, N);
    double2 wb[5];
    mp_twiddles<LOGN, R>(wb, P.tw, t);

    for (int s = 0; s < nfr; ++s) {
        asm volatile(
            "{\n.reg .pred p;\nWAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
            ::"r"(sp_smem_u32(&mbar)), "r"(s & 1) : "memory");
        const int start = (s * hop) & (N - 1);
        double2 za[R], zb[R];
        double sa = 0.0, sbsum = 0.0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int row = (start + 2 * (t + k * T)) & (N - 1);   // even: row + 1 does not wrap
            const double2 v0 = ring[row], v1 = ring[row + 1];
            za[k] = make_double2(v0.x, v1.x);
            zb[k] = make_double2(v0.y, v1.y);
            sa += v0.x + v1.x;
            sbsum += v0.y + v1.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sbsum += __shfl_xor_sync(0xffffffffu, sbsum, o);
        }
        if ((t & 31) == 0) { red[0][t >> 5] = sa; red[1][t >> 5] = sbsum; }
        __syncthreads();                              // every thread holds its rows: the oldest are free
        if (t == 0 && s + 1 < nfr) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            copy_rows(s * hop + N, hop);                  // replaces rows [s hop, (s + 1) hop) of the ring
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            double2 w = __ldg(reinterpret_cast<const double2*>(P.win) + (t + k * T));
            za[k].x *= w.x; za[k].y *= w.y;
            zb[k].x *= w.x; zb[k].y *= w.y;
        }
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
            if (ch == 1) {
#pragma unroll
                for (int k = 0; k < R; ++k) za[k] = zb[k];
            }
            mp_channel<LOGN, R, DB>(za, S, wb, t, red[ch], P.detrend, P.scale, P.tw,
                                    P.dst + (((f0 + s) * C + c0 + ch) * (int64_t)F));
            __syncthreads();                          // S (and, after the second channel, red) free again
        }
    }
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

TmapEncodeFn tmap_encoder() {
    static TmapEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<TmapEncodeFn>(p);
    }();
    return fn;
}

int32_t make_pair_tmap(const double* src, int32_t C, int64_t rows, CUtensorMap* tm) {
    TmapEncodeFn enc = tmap_encoder();
    if (!enc || C % 2 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 || rows >= 0x7fffffff)
        return ADN_ERR_UNSUPPORTED;
    const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)C * 8};
    const cuuint32_t box[2] = {2, MPT_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(src), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return ADN_ERR_UNSUPPORTED;
    return ADN_OK;
}

// ADN_ERR_UNSUPPORTED (no message) when the shape or the driver does not allow it: the caller falls back
template <int LOGN>
int32_t launch_mpt(SpecMArgs& P, const double* src, int64_t nf, int out_db, cudaStream_t st) {
    constexpr int M = 1 << (LOGN - 1), T = M / 8, N = 2 * M;
    CUtensorMap tm;
    if (make_pair_tmap(src, P.C, (nf - 1) * (int64_t)P.hop + N, &tm) != ADN_OK) return ADN_ERR_UNSUPPORTED;
    P.npair = P.C / 2;
    const size_t smem = (size_t)N * 16 + 128;
    const int64_t grid = nf * P.npair;
    if (grid > 0x7fffffff) return ADN_ERR_UNSUPPORTED;
    auto k0 = spectrogram_mpt_kernel<LOGN, false>;
    auto k1 = spectrogram_mpt_kernel<LOGN, true>;
    static bool attr_done = false;
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ADN_CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    if (out_db) k1<<<(unsigned)grid, T, smem, st>>>(P, tm);
    else k0<<<(unsigned)grid, T, smem, st>>>(P, tm);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

template <int LOGN>
int32_t launch_mprt(SpecMArgs& P, const double* src, int64_t nf, int out_db, cudaStream_t st) {
    constexpr int M = 1 << (LOGN - 1), T = M / 8, N = 2 * M;
    if (P.hop % MPT_ROWS != 0 || P.hop > N) return ADN_ERR_UNSUPPORTED;
    CUtensorMap tm;
    if (make_pair_tmap(src, P.C, (nf - 1) * (int64_t)P.hop + N, &tm) != ADN_OK) return ADN_ERR_UNSUPPORTED;
    P.npair = P.C / 2;
    const size_t smem = (size_t)N * 16 + 128 + (size_t)(M + M / 8 + 8) * 16;
    auto k0 = spectrogram_mprt_kernel<LOGN, false>;
    auto k1 = spectrogram_mprt_kernel<LOGN, true>;
    static bool attr_done = false;
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ADN_CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    int bps = 1;
    ADN_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k0, T, smem));
    if (bps < 1) return ADN_ERR_UNSUPPORTED;
    // about two resident waves of blocks, runs of at least 8 frames
    int64_t nruns = (int64_t)ctx().sm_count * bps * 2 / P.npair;
    if (nruns < 1) nruns = 1;
    int64_t frun = (nf + nruns - 1) / nruns;
    if (frun < 8) frun = nf < 8 ? nf : 8;
    const int64_t grid = ((nf + frun - 1) / frun) * P.npair;
    if (grid > 0x7fffffff) return ADN_ERR_UNSUPPORTED;
    if (out_db) k1<<<(unsigned)grid, T, smem, st>>>(P, tm, (int32_t)frun);
    else k0<<<(unsigned)grid, T, smem, st>>>(P, tm, (int32_t)frun);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

template <int LOGN, int R>
int32_t launch_mpr(SpecMArgs& P, int64_t nf, int out_db, cudaStream_t st) {
    constexpr int M = 1 << (LOGN - 1), T = M / R, N = 2 * M;
    P.npair = P.C / 2;
    const size_t smem = (size_t)(M + M / 8 + 8) * 16 + (size_t)2 * (N + 4) * 8;
    auto k0 = spectrogram_mpr_kernel<LOGN, R, false>;
    auto k1 = spectrogram_mpr_kernel<LOGN, R, true>;
    static bool attr_done = false;
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ADN_CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    int bps = 1;
    ADN_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k0, T, smem));
    if (bps < 1) return ADN_ERR_UNSUPPORTED;
    // about two resident waves of blocks, runs of at least 8 frames
    int64_t nruns = (int64_t)ctx().sm_count * bps * 2 / P.npair;
    if (nruns < 1) nruns = 1;
    int64_t frun = (nf + nruns - 1) / nruns;
    if (frun < 8) frun = nf < 8 ? nf : 8;
    if (frun * P.hop + N > 0x3fffffff) return ADN_ERR_UNSUPPORTED;
    const int64_t grid = ((nf + frun - 1) / frun) * P.npair;
    if (grid > 0x7fffffff) return ADN_ERR_UNSUPPORTED;
    if (out_db) k1<<<(unsigned)grid, T, smem, st>>>(P, (int32_t)frun);
    else k0<<<(unsigned)grid, T, smem, st>>>(P, (int32_t)frun);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

template <int LOGN, int R, int CP>
int32_t launch_mp(SpecMArgs& P, int64_t nf, int out_db, cudaStream_t st) {
    constexpr int M = 1 << (LOGN - 1), T = M / R;
    P.npair = (P.C + CP - 1) / CP;
    const size_t smem = (size_t)(M + M / 8 + 8) * 16;
    const int64_t grid = nf * P.npair;
    if (grid > 0x7fffffff) return fail(ADN_ERR_UNSUPPORTED, "spectrogram: grid %lld", (long long)grid);
    auto k0 = spectrogram_mp_kernel<LOGN, R, CP, false>;
    auto k1 = spectrogram_mp_kernel<LOGN, R, CP, true>;
    static bool attr_done = false;
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ADN_CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    if (out_db) k1<<<(unsigned)grid, T, smem, st>>>(P);
    else k0<<<(unsigned)grid, T, smem, st>>>(P);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

// ======================================================================================
// Frames longer than one block's shared memory (nfft 2^15 .. 2^20; the GUI offers up to 2^19,
// databrowser.py:516): the complex transforms live in a work buffer in global memory.
//   pack    frame -> minus mean, x window -> M = nfft/2 complex points (even, odd samples)
//   stages  radix-2^2 decimation-in-frequency passes over the whole buffer while the
//           sub-transforms are longer than what a block holds in shared memory
//   tail    the remaining stages of each 4096-point sub-block in shared memory
//   split   Z (bit-reversed order) -> |X|^2 scaling of the real transform -> dst
constexpr int BIG_LOGSL = 12;                 // sub-block of the tail kernel: 4096 complex, 64 KB
constexpr int BIG_NT = 256;

// two DIF stages (n = 2^lg and n/2) on four points z[0], z[q4], z[2 q4], z[3 q4]
__device__ __forceinline__ void dif4(double2* z, int q4, double2 w1, double2 w2) {
    double2 a0 = z[0], a1 = z[q4], a2 = z[2 * q4], a3 = z[3 * q4];
    double2 b0 = make_double2(a0.x + a2.x, a0.y + a2.y);
    double2 b2 = cmul(make_double2(a0.x - a2.x, a0.y - a2.y), w1);
    double2 b1 = make_double2(a1.x + a3.x, a1.y + a3.y);
    double2 d13 = make_double2(a1.x - a3.x, a1.y - a3.y);
    double2 b3 = cmul(make_double2(d13.y, -d13.x), w1);        // W_n^(k+n/4) = -i W_n^k
    z[0] = make_double2(b0.x + b1.x, b0.y + b1.y);
    z[q4] = cmul(make_double2(b0.x - b1.x, b0.y - b1.y), w2);
    z[2 * q4] = make_double2(b2.x + b3.x, b2.y + b3.y);
    z[3 * q4] = cmul(make_double2(b2.x - b3.x, b2.y - b3.y), w2);
}

// tw: exp(-2 pi i j / 2^twlog), j < 2^(twlog-1);  W_n^k = tw[k << (twlog - lg)]
__global__ void __launch_bounds__(BIG_NT)
big_stage2_kernel(double2* __restrict__ work, const double2* __restrict__ tw, int64_t items,
                  int32_t logL, int32_t twlog, int32_t lg) {
    const int64_t total = items << (logL - 2);
    const int q4 = 1 << (lg - 2), sh = twlog - lg;
    for (int64_t idx = (int64_t)blockIdx.x * BIG_NT + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * BIG_NT) {
        const int64_t it = idx >> (logL - 2);
        const int r = (int)(idx & (((int64_t)1 << (logL - 2)) - 1));
        const int blk = r >> (lg - 2), k = r & (q4 - 1);
        double2* z = work + (it << logL) + ((int64_t)blk << lg) + k;
        dif4(z, q4, __ldg(tw + ((int64_t)k << sh)), __ldg(tw + ((int64_t)(2 * k) << sh)));
    }
}

__global__ void __launch_bounds__(BIG_NT)
big_stage1_kernel(double2* __restrict__ work, const double2* __restrict__ tw, int64_t items,
                  int32_t logL, int32_t twlog, int32_t lg) {
    const int64_t total = items << (logL - 1);
    const int h = 1 << (lg - 1), sh = twlog - lg;
    for (int64_t idx = (int64_t)blockIdx.x * BIG_NT + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * BIG_NT) {
        const int64_t it = idx >> (logL - 1);
        const int r = (int)(idx & (((int64_t)1 << (logL - 1)) - 1));
        const int blk = r >> (lg - 1), k = r & (h - 1);
        double2* z = work + (it << logL) + ((int64_t)blk << lg) + k;
        double2 a = z[0], b = z[h];
        z[0] = make_double2(a.x + b.x, a.y + b.y);
        z[h] = cmul(make_double2(a.x - b.x, a.y - b.y), __ldg(tw + ((int64_t)k << sh)));
    }
}

// the last lgs stages of one 2^lgs-point sub-block per block, in shared memory
__global__ void __launch_bounds__(BIG_NT)
big_tail_kernel(double2* __restrict__ work, const double2* __restrict__ tw, int32_t twlog, int32_t lgs) {
    extern __shared__ __align__(16) double sbuf[];
    double2* zs = reinterpret_cast<double2*>(sbuf);
    const int SL = 1 << lgs;
    double2* g = work + ((int64_t)blockIdx.x << lgs);
    for (int j = threadIdx.x; j < SL; j += BIG_NT) zs[j] = g[j];
    __syncthreads();
    int lg = lgs;
    for (; lg >= 2; lg -= 2) {
        const int q4 = 1 << (lg - 2), sh = twlog - lg;
        for (int r = threadIdx.x; r < (SL >> 2); r += BIG_NT) {
            const int blk = r >> (lg - 2), k = r & (q4 - 1);
            dif4(zs + (blk << lg) + k, q4, __ldg(tw + ((int64_t)k << sh)), __ldg(tw + ((int64_t)(2 * k) << sh)));
        }
        __syncthreads();
    }
    if (lg == 1) {
        for (int r = threadIdx.x; r < (SL >> 1); r += BIG_NT) {
            double2 u = zs[2 * r], v = zs[2 * r + 1];
            zs[2 * r] = make_double2(u.x + v.x, u.y + v.y);
            zs[2 * r + 1] = make_double2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < SL; j += BIG_NT) g[j] = zs[j];
}

// forward transform of `items` arrays of 2^logL complex points, in place, bit-reversed result
int32_t big_fft_dif(double2* work, int64_t items, int logL, const double2* tw, int twlog, cudaStream_t st) {
    const int lgs = logL < BIG_LOGSL ? logL : BIG_LOGSL;
    const int64_t cap = (int64_t)ctx().sm_count * 32;
    int lg = logL;
    while (lg > lgs) {
        if (lg - lgs >= 2) {
            int64_t nb = ((items << (logL - 2)) + BIG_NT - 1) / BIG_NT;
            big_stage2_kernel<<<(unsigned)(nb < cap ? nb : cap), BIG_NT, 0, st>>>(work, tw, items, logL, twlog, lg);
            lg -= 2;
        } else {
            int64_t nb = ((items << (logL - 1)) + BIG_NT - 1) / BIG_NT;
            big_stage1_kernel<<<(unsigned)(nb < cap ? nb : cap), BIG_NT, 0, st>>>(work, tw, items, logL, twlog, lg);
            lg -= 1;
        }
        count_launch();
    }
    static bool attr_done = false;
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(big_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        attr_done = true;
    }
    const int64_t nblk = items << (logL - lgs);
    if (nblk > 0x7fffffff) return fail(ADN_ERR_UNSUPPORTED, "spectrogram: %lld tail blocks", (long long)nblk);
    big_tail_kernel<<<(unsigned)nblk, BIG_NT, (size_t)16 << lgs, st>>>(work, tw, twlog, lgs);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

// frame (f0 + item / C, channel item % C) -> work[item][0 .. M)
__global__ void __launch_bounds__(BIG_NT)
big_pack_kernel(const double* __restrict__ src, int32_t C, int64_t hop, int32_t N, int64_t f0,
                const double* __restrict__ win, int32_t detrend, double2* __restrict__ work) {
    __shared__ double red[BIG_NT / 32];
    __shared__ double s_mean;
    const int64_t item = blockIdx.x;
    const int64_t fi = item / C;
    const int c = (int)(item - fi * C);
    const double* x = src + ((f0 + fi) * hop) * C + c;
    const int M = N >> 1;
    double mean = 0.0;
    if (detrend) {
        double s = 0.0;
        for (int j = threadIdx.x; j < N; j += BIG_NT) s += __ldg(x + (int64_t)j * C);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < BIG_NT / 32; ++i) t += red[i];
            s_mean = t / (double)N;
        }
        __syncthreads();
        mean = s_mean;
    }
    double2* z = work + item * M;
    for (int j = threadIdx.x; j < M; j += BIG_NT) {
        double a = __ldg(x + (int64_t)(2 * j) * C), b = __ldg(x + (int64_t)(2 * j + 1) * C);
        z[j] = make_double2((a - mean) * __ldg(win + 2 * j), (b - mean) * __ldg(win + 2 * j + 1));
    }
}

__global__ void __launch_bounds__(BIG_NT)
big_split_kernel(const double2* __restrict__ work, const double2* __restrict__ tw, int64_t items,
                 int32_t logM, double scale, int32_t out_db, double* __restrict__ dst) {
    const int M = 1 << logM, F = M + 1, sh = 32 - logM;
    const int64_t total = items * F;
    for (int64_t idx = (int64_t)blockIdx.x * BIG_NT + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * BIG_NT) {
        const int64_t it = idx / F;
        const int k = (int)(idx - it * F);
        const double2* z = work + (it << logM);
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
            double2 w = __ldg(tw + k);
            xr = er + (orr * w.x - oi * w.y);
            xi = ei + (orr * w.y + oi * w.x);
            fac = 2.0;
        }
        double pw = (xr * xr + xi * xi) * (scale * fac);
        if (out_db) pw = to_db(pw);
        dst[idx] = pw;
    }
}

int32_t spectrogram_big(const SpecPlan& plan, const double* src, int32_t C, double rate, int32_t nfft,
                        int32_t hop, int32_t detrend, double* dst, int64_t nf, int32_t out_db,
                        cudaStream_t st) {
    int logN = 0;
    while ((1 << logN) < nfft) ++logN;
    const int logM = logN - 1;
    const int64_t M = (int64_t)1 << logM, F = M + 1;
    // frames per pass: the work buffer stays below 1 GiB (at least one frame of all channels)
    int64_t fr = ((int64_t)1 << 30) / (M * 16 * C);
    if (fr < 1) fr = 1;
    if (fr > nf) fr = nf;
    DevBuf& wb = scratch(SCR_SPEC_WORK, st);
    int32_t rc = wb.reserve((size_t)(fr * C * M * 16));
    if (rc) return rc;
    double2* work = wb.as<double2>();
    const double scale = 1.0 / (rate * plan.sumw2);
    const int64_t cap = (int64_t)ctx().sm_count * 32;
    for (int64_t f0 = 0; f0 < nf; f0 += fr) {
        const int64_t fc = f0 + fr < nf ? fr : nf - f0;
        const int64_t items = fc * C;
        if (items > 0x7fffffff) return fail(ADN_ERR_UNSUPPORTED, "spectrogram: %lld items", (long long)items);
        big_pack_kernel<<<(unsigned)items, BIG_NT, 0, st>>>(src, C, hop, nfft, f0, plan.win, detrend, work);
        count_launch();
        if ((rc = big_fft_dif(work, items, logM, plan.tw, logN, st))) return rc;
        int64_t nb = (items * F + BIG_NT - 1) / BIG_NT;
        big_split_kernel<<<(unsigned)(nb < cap ? nb : cap), BIG_NT, 0, st>>>(work, plan.tw, items, logM, scale,
                                                                           out_db, dst + f0 * C * F);
        count_launch();
        ADN_CK(cudaGetLastError());
    }
    return ADN_OK;
}


// ======================================================================================
// nfft that is no power of two (reachable through update(): nfft is clamped to
// len(source)//2, bufferedspectrogram.py:88): Bluestein's chirp-z form of the DFT,
//   Y[k] = conj(c[k]) sum_n (y[n] conj(c[n])) c[k - n],   c[n] = exp(i pi n^2 / N),
// as a circular convolution of length L = 2^logL >= 1.5 N by two power-of-two transforms in
// the work buffer.  Only |Y[k]|^2 is wanted, so the final chirp multiplication drops out.
__global__ void __launch_bounds__(BIG_NT)
blue_pack_kernel(const double* __restrict__ src, int32_t C, int64_t hop, int32_t N, int32_t logL, int64_t f0,
                 const double* __restrict__ win, const double2* __restrict__ chirp, int32_t detrend,
                 double2* __restrict__ work) {
    __shared__ double red[BIG_NT / 32];
    __shared__ double s_mean;
    const int64_t item = blockIdx.x;
    const int64_t fi = item / C;
    const int c = (int)(item - fi * C);
    const double* x = src + ((f0 + fi) * hop) * C + c;
    double mean = 0.0;
    if (detrend) {
        double s = 0.0;
        for (int j = threadIdx.x; j < N; j += BIG_NT) s += __ldg(x + (int64_t)j * C);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < BIG_NT / 32; ++i) t += red[i];
            s_mean = t / (double)N;
        }
        __syncthreads();
        mean = s_mean;
    }
    double2* z = work + (item << logL);
    const int L = 1 << logL;
    for (int j = threadIdx.x; j < L; j += BIG_NT) {
        double2 v = make_double2(0.0, 0.0);
        if (j < N) {
            double y = (__ldg(x + (int64_t)j * C) - mean) * __ldg(win + j);
            double2 ch = __ldg(chirp + j);
            v = make_double2(y * ch.x, -y * ch.y);
        }
        z[j] = v;
    }
}

// b[brev(i)] = conj(a[i] * B[i]): product of the two spectra, back in natural order for the
// second forward transform (ifft(x) = conj(fft(conj(x))) / L)
__global__ void __launch_bounds__(BIG_NT)
blue_mul_kernel(const double2* __restrict__ a, const double2* __restrict__ B, int64_t items, int32_t logL,
                double2* __restrict__ b) {
    const int64_t total = items << logL;
    const int sh = 32 - logL;
    for (int64_t idx = (int64_t)blockIdx.x * BIG_NT + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * BIG_NT) {
        const int64_t it = idx >> logL;
        const unsigned i = (unsigned)(idx & (((int64_t)1 << logL) - 1));
        double2 v = cmul(a[idx], __ldg(B + i));
        b[(it << logL) + (__brev(i) >> sh)] = make_double2(v.x, -v.y);
    }
}

__global__ void __launch_bounds__(BIG_NT)
blue_power_kernel(const double2* __restrict__ r, int64_t items, int32_t N, int32_t logL, double scale,
                  int32_t out_db, double* __restrict__ dst) {
    const int F = N / 2 + 1, sh = 32 - logL;
    const int64_t total = items * F;
    const double inv = 1.0 / (double)((int64_t)1 << logL);
    for (int64_t idx = (int64_t)blockIdx.x * BIG_NT + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * BIG_NT) {
        const int64_t it = idx / F;
        const int k = (int)(idx - it * F);
        double2 v = r[(it << logL) + (__brev((unsigned)k) >> sh)];
        const double xr = v.x * inv, xi = v.y * inv;
        const bool edge = k == 0 || (N % 2 == 0 && k == N / 2);
        double pw = (xr * xr + xi * xi) * (scale * (edge ? 1.0 : 2.0));
        if (out_db) pw = to_db(pw);
        dst[idx] = pw;
    }
}

int32_t spectrogram_bluestein(const SpecPlan& plan, const double* src, int32_t C, double rate, int32_t nfft,
                              int32_t hop, int32_t detrend, double* dst, int64_t nf, int32_t out_db,
                              cudaStream_t st) {
    const int logL = plan.logL;
    const int64_t L = (int64_t)1 << logL, F = nfft / 2 + 1;
    int64_t fr = ((int64_t)1 << 29) / (L * 16 * C);             // two buffers of at most 512 MiB
    if (fr < 1) fr = 1;
    if (fr > nf) fr = nf;
    DevBuf& wb = scratch(SCR_SPEC_WORK, st);
    int32_t rc = wb.reserve((size_t)(2 * fr * C * L * 16));
    if (rc) return rc;
    double2* wa = wb.as<double2>();
    double2* wc = wa + fr * C * L;
    const double scale = 1.0 / (rate * plan.sumw2);
    const int64_t cap = (int64_t)ctx().sm_count * 32;
    for (int64_t f0 = 0; f0 < nf; f0 += fr) {
        const int64_t fc = f0 + fr < nf ? fr : nf - f0;
        const int64_t items = fc * C;
        if (items > 0x7fffffff) return fail(ADN_ERR_UNSUPPORTED, "spectrogram: %lld items", (long long)items);
        blue_pack_kernel<<<(unsigned)items, BIG_NT, 0, st>>>(src, C, hop, nfft, logL, f0, plan.win, plan.chirp,
                                                           detrend, wa);
        count_launch();
        if ((rc = big_fft_dif(wa, items, logL, plan.twL, logL, st))) return rc;
        int64_t nb = ((items << logL) + BIG_NT - 1) / BIG_NT;
        blue_mul_kernel<<<(unsigned)(nb < cap ? nb : cap), BIG_NT, 0, st>>>(wa, plan.Bbr, items, logL, wc);
        count_launch();
        if ((rc = big_fft_dif(wc, items, logL, plan.twL, logL, st))) return rc;
        nb = (items * F + BIG_NT - 1) / BIG_NT;
        blue_power_kernel<<<(unsigned)(nb < cap ? nb : cap), BIG_NT, 0, st>>>(wc, items, nfft, logL, scale, out_db,
                                                                            dst + f0 * C * F);
        count_launch();
        ADN_CK(cudaGetLastError());
    }
    return ADN_OK;
}

}  // namespace

int32_t spectrogram_dev(const double* src, int64_t n_src, int32_t C, double rate, int32_t nfft,
                        int32_t hop, int32_t window_id, int32_t detrend_id, double* dst,
                        int64_t n_dst, int32_t out_db, int64_t* n_computed, cudaStream_t st) {
    if (window_id != ADN_WINDOW_HANN)
        return fail(ADN_ERR_UNSUPPORTED, "spectrogram: window_id %d (only ADN_WINDOW_HANN)", window_id);
    if (detrend_id != ADN_DETREND_NONE && detrend_id != ADN_DETREND_CONSTANT)
        return fail(ADN_ERR_INVALID, "spectrogram: detrend_id %d", detrend_id);
    if (nfft < ADN_MIN_NFFT || nfft > ADN_MAX_NFFT)
        return fail(ADN_ERR_UNSUPPORTED, "spectrogram: nfft=%d (supported: %d .. %d)",
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
    if (nfft & (nfft - 1))
        return spectrogram_bluestein(plan, src, C, rate, nfft, hop, detrend_id == ADN_DETREND_CONSTANT, dst,
                                     nf, out_db, st);
    const bool ring_ok = plan.twA && (hop % 2 == 0) && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                         (C % 2 == 0 || C == 1) && env_int("ADN_SPEC_RING", 1) != 0;
    if (ring_ok) {
        SpecRArgs R;
        R.src = src; R.dst = dst; R.win = plan.win; R.twA = plan.twA; R.twS = plan.tw;
        R.nframes = nf; R.C = C; R.hop = hop;
        R.detrend = detrend_id == ADN_DETREND_CONSTANT;
        // wide arrays: every group's block pulls its share of the next rows into L2 instead of all of
        // them (measured on B200, 64 ch x 250 kHz: nfft 1024 / hop 512 293 -> 276 us, 256 / 128 334 ->
        // 327 us; 512 / 128 508 -> 517 us, 16 and 8 channels +-1 %); 2 = no prefetch (3 - 12 % slower)
        R.pfsplit = env_int("ADN_SPEC_PFSPLIT", C >= 32 && 2 * hop >= nfft ? 1 : 0);
        R.scale = 1.0 / (rate * plan.sumw2);
        int32_t rr = ADN_ERR_UNSUPPORTED;
        switch (nfft) {
            case 128: rr = launch_ring<7>(R, nf, out_db, st); break;
            case 256: rr = launch_ring<8>(R, nf, out_db, st); break;
            case 512: rr = launch_ring<9>(R, nf, out_db, st); break;
            case 1024: rr = launch_ring<10>(R, nf, out_db, st); break;
        }
        if (rr != ADN_ERR_UNSUPPORTED) return rr;
    }
    if (plan.twA && (hop % 2 == 0) && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
        SpecWArgs W;
        W.src = src; W.dst = dst; W.win = plan.win; W.twA = plan.twA; W.twS = plan.tw;
        W.nframes = nf; W.C = C; W.hop = hop;
        W.detrend = detrend_id == ADN_DETREND_CONSTANT; W.out_db = out_db;
        W.scale = 1.0 / (rate * plan.sumw2);
        switch (nfft) {
            case 128: return launch_warp_kernel<7>(W, nf, st);
            case 256: return launch_warp_kernel<8>(W, nf, st);
            case 512: return launch_warp_kernel<9>(W, nf, st);
            case 1024: return launch_warp_kernel<10>(W, nf, st);
        }
    }
    if (nfft > 16384)
        return spectrogram_big(plan, src, C, rate, nfft, hop, detrend_id == ADN_DETREND_CONSTANT, dst, nf,
                               out_db, st);
    if (nfft >= 2048 && env_int("ADN_SPEC_MP", 1) != 0) {
        SpecMArgs Q;
        Q.src = src; Q.dst = dst; Q.win = plan.win; Q.tw = plan.tw;
        Q.nframes = nf; Q.C = C; Q.hop = hop;
        Q.detrend = detrend_id == ADN_DETREND_CONSTANT;
        Q.scale = 1.0 / (rate * plan.sumw2);
        // channel pairs need 16-byte aligned rows: even C and an aligned base
        const bool pairs = C % 2 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
        const bool tma_ok = pairs && C >= 4 && env_int("ADN_SPEC_TMA", 1) != 0;
        // overlapping frames of channel pairs: a ring of one frame of rows per block (nfft 2048, 4096).
        // Filled by the TMA unit while the transforms run (mprt) up to 75 % overlap -- measured on B200
        // against the thread-staged ring (mpr): 64 ch 2048 / 75 % 815 -> 714 us, 4096 / 75 % 928 -> 796 us,
        // 16 ch 4096 / 50 % 532 -> 405 us, 8 ch 2048 / 50 % 268 -> 237 us; at 87.5 % mpr's 16 points per
        // thread win (64 ch 2048: 1349 against 1393 us, 8 ch: 797 against 884 us), and at 75 % on 8
        // channels too (4096: 491 against 513 us)
        if (pairs && hop % 2 == 0 && nf >= 4 && 2 * hop <= nfft && (nfft == 2048 || nfft == 4096)) {
            int mprt = env_int("ADN_SPEC_MPRT", -1);
            // 50 %: always (64 ch 2048: 375 against 397 us of the frame-per-block TMA kernel, 4096 even);
            // 75 %: from 16 channels on; 87.5 %: never
            if (mprt < 0) mprt = 4 * hop > nfft ? 1 : (8 * hop > nfft ? (C >= 16) : 0);
            if (mprt && tma_ok) {
                int32_t rt = ADN_ERR_UNSUPPORTED;
                if (nfft == 2048) rt = launch_mprt<11>(Q, src, nf, out_db, st);
                if (nfft == 4096) rt = launch_mprt<12>(Q, src, nf, out_db, st);
                if (rt != ADN_ERR_UNSUPPORTED) return rt;
            }
        }
        // (measured: the thread-staged ring pays from 75 % overlap on; at 50 % only for few channels, where
        // runs are long)
        if (pairs && hop % 2 == 0 && nf >= 4 && (4 * hop <= nfft || (2 * hop <= nfft && C <= 16)) &&
            env_int("ADN_SPEC_MPR", 1) != 0) {
            // 16 points per thread (three passes instead of four, half the threads): measured on B200
            // 6 - 10 % faster than 8 points at 75 and 87.5 % overlap (8 and 64 channels), although the
            // compiler keeps the twiddle powers of all passes in registers across the frame loop (255)
            int32_t rr = ADN_ERR_UNSUPPORTED;
            if (nfft == 2048) rr = launch_mpr<11, 16>(Q, nf, out_db, st);
            if (nfft == 4096) rr = launch_mpr<12, 16>(Q, nf, out_db, st);
            if (rr != ADN_ERR_UNSUPPORTED) return rr;
        }
        // rows gathered by the TMA unit (measured on B200: 64 ch 2048 / 0 % 270 -> 210 us, 4096 / 50 %
        // 516 -> 421 us, 8 ch 2048 168 -> 141 us; two channels = contiguous rows: 120 -> 142 us, so not there)
        // (8192 points the same way, 128 KB of rows per block: 342 against 335 us -- not taken)
        if (tma_ok) {
            int32_t rr = ADN_ERR_UNSUPPORTED;
            if (nfft == 2048) rr = launch_mpt<11>(Q, src, nf, out_db, st);
            if (nfft == 4096) rr = launch_mpt<12>(Q, src, nf, out_db, st);
            if (rr != ADN_ERR_UNSUPPORTED) return rr;
        }
        switch (nfft) {
            case 2048: return pairs ? launch_mp<11, 8, 2>(Q, nf, out_db, st) : launch_mp<11, 8, 1>(Q, nf, out_db, st);
            case 4096: return pairs ? launch_mp<12, 8, 2>(Q, nf, out_db, st) : launch_mp<12, 8, 1>(Q, nf, out_db, st);
            case 8192: return launch_mp<13, 16, 1>(Q, nf, out_db, st);
            case 16384: return launch_mp<14, 16, 1>(Q, nf, out_db, st);
        }
    }
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
