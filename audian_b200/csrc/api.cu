// C ABI of libaudian_b200.so: context, error state, host-pointer entry points.
// Declarations and the reference functions each entry replaces: include/audian_b200.h
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace adn {

static thread_local std::string g_err;

int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int32_t fail_cuda(cudaError_t e, const char* what) {
    return fail(ADN_ERR_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

int32_t DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return ADN_OK;
    if (p) { ADN_CK(cudaFree(p)); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {             // retry with the exact size before giving up
        cudaGetLastError();
        want = bytes;
        ADN_CK(cudaMalloc(&p, want));
    }
    cap = want;
    return ADN_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static Ctx g_ctx;
static DevBuf g_scratch[SCR_COUNT];
static std::mutex g_mu;

Ctx& ctx() { return g_ctx; }
DevBuf& scratch(int slot) { return g_scratch[slot]; }

static int32_t init_locked(int32_t device) {
    if (g_ctx.ready && (device < 0 || device == g_ctx.device)) return ADN_OK;
    if (g_ctx.ready) return fail(ADN_ERR_INVALID, "adn_init: already initialised on device %d",
                                 g_ctx.device);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(ADN_ERR_CUDA, "no CUDA device available (%s); libaudian_b200 has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0) {
        int cur = 0;
        if (cudaGetDevice(&cur) == cudaSuccess) device = cur; else device = 0;
    }
    if (device >= count) return fail(ADN_ERR_INVALID, "adn_init: device %d of %d", device, count);
    ADN_CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    ADN_CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(ADN_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    g_ctx.sm_count = prop.multiProcessorCount;
    ADN_CK(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    g_ctx.device = device;
    g_ctx.ready = true;
    return ADN_OK;
}

int32_t ensure_init() {
    std::lock_guard<std::mutex> lk(g_mu);
    int32_t rc = init_locked(-1);
    if (rc == ADN_OK) {
        // entry points may be called from a thread whose current device differs
        cudaError_t e = cudaSetDevice(g_ctx.device);
        if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
    }
    return rc;
}

// Stage `bytes` of host memory into the `in` staging buffer.
static int32_t stage_in(const void* host, size_t bytes) {
    Ctx& c = ctx();
    int32_t rc = c.in.reserve(bytes ? bytes : 16);
    if (rc) return rc;
    if (bytes) ADN_CK(cudaMemcpyAsync(c.in.p, host, bytes, cudaMemcpyHostToDevice, c.stream));
    return ADN_OK;
}

static int32_t stage_out(void* host, size_t bytes) {
    Ctx& c = ctx();
    if (bytes) ADN_CK(cudaMemcpyAsync(host, c.out.p, bytes, cudaMemcpyDeviceToHost, c.stream));
    ADN_CK(cudaStreamSynchronize(c.stream));
    return ADN_OK;
}

}  // namespace adn

using namespace adn;

extern "C" {

int32_t adn_init(int32_t device) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_locked(device);
}

int32_t adn_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx.ready) return ADN_OK;
    cudaSetDevice(g_ctx.device);
    cudaStreamSynchronize(g_ctx.stream);
    g_ctx.in.release();
    g_ctx.out.release();
    g_ctx.aux.release();
    for (int i = 0; i < SCR_COUNT; ++i) g_scratch[i].release();
    cudaStreamDestroy(g_ctx.stream);
    g_ctx.stream = nullptr;
    g_ctx.ready = false;
    g_ctx.device = -1;
    return ADN_OK;
}

const char* adn_last_error(void) { return g_err.c_str(); }
int32_t adn_version(void) { return 100; }
int64_t adn_launch_count(void) { return g_ctx.launches.load(); }

int32_t adn_synchronize(void) {
    int32_t rc = ensure_init();
    if (rc) return rc;
    ADN_CK(cudaStreamSynchronize(ctx().stream));
    return ADN_OK;
}

int32_t adn_host_register(void* ptr, int64_t bytes) {
    int32_t rc = ensure_init();
    if (rc) return rc;
    if (!ptr || bytes <= 0) return fail(ADN_ERR_INVALID, "adn_host_register: empty range");
    ADN_CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return ADN_OK;
}

int32_t adn_host_unregister(void* ptr) {
    int32_t rc = ensure_init();
    if (rc) return rc;
    ADN_CK(cudaHostUnregister(ptr));
    return ADN_OK;
}

// ---------------------------------------------------------------- host entry points

int32_t adn_minmax_f64(const double* src, int64_t n, int32_t C, int64_t step, double* dst) {
    if (n < 0 || C < 1 || step < 1) return fail(ADN_ERR_INVALID, "adn_minmax_f64: n=%lld C=%d step=%lld",
                                               (long long)n, C, (long long)step);
    if (n == 0) return ADN_OK;
    if (!src || !dst) return fail(ADN_ERR_INVALID, "adn_minmax_f64: NULL pointer");
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    int64_t nseg = (n + step - 1) / step;
    size_t in_b = (size_t)n * C * 8, out_b = (size_t)nseg * 2 * C * 8;
    if ((rc = stage_in(src, in_b))) return rc;
    if ((rc = c.out.reserve(out_b))) return rc;
    if ((rc = minmax_dev(c.in.as<double>(), n, C, step, c.out.as<double>(), c.stream))) return rc;
    return stage_out(dst, out_b);
}

int32_t adn_sosfilt_f64(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                        int64_t nbefore, double* dst, int64_t n_dst, double* zi_inout) {
    if (S < 0 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 0 || n_dst < 0 || nbefore < 0)
        return fail(ADN_ERR_INVALID, "adn_sosfilt_f64: S=%d C=%d n_src=%lld n_dst=%lld nbefore=%lld",
                    S, C, (long long)n_src, (long long)n_dst, (long long)nbefore);
    if (n_dst > n_src - nbefore)
        return fail(ADN_ERR_INVALID, "adn_sosfilt_f64: n_dst=%lld exceeds n_src-nbefore=%lld",
                    (long long)n_dst, (long long)(n_src - nbefore));
    if ((S > 0 && !sos) || (n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_sosfilt_f64: NULL pointer");
    if (n_src == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    size_t in_b = (size_t)n_src * C * 8, out_b = (size_t)n_dst * C * 8;
    size_t z_b = (size_t)C * (S > 0 ? S : 1) * 2 * 8;
    if ((rc = stage_in(src, in_b))) return rc;
    if ((rc = c.out.reserve(out_b ? out_b : 16))) return rc;
    double* dzi = nullptr;
    if (zi_inout && S > 0) {
        if ((rc = c.aux.reserve(2 * z_b))) return rc;
        dzi = c.aux.as<double>();
        ADN_CK(cudaMemcpyAsync(dzi, zi_inout, z_b, cudaMemcpyHostToDevice, c.stream));
    }
    double* dzf = dzi ? dzi + (size_t)C * S * 2 : nullptr;
    if ((rc = sosfilt_dev(sos, S, c.in.as<double>(), n_src, C, nbefore,
                          n_dst > 0 ? c.out.as<double>() : nullptr, n_dst, dzi, dzf, c.stream)))
        return rc;
    if (dzf) ADN_CK(cudaMemcpyAsync(zi_inout, dzf, z_b, cudaMemcpyDeviceToHost, c.stream));
    return stage_out(dst, out_b);
}

int32_t adn_envelope_f64(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                         int64_t nbefore, double* dst, int64_t n_dst, int32_t clamp_negative) {
    if (S < 0 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 0 || n_dst < 0 || nbefore < 0)
        return fail(ADN_ERR_INVALID, "adn_envelope_f64: S=%d C=%d n_src=%lld n_dst=%lld nbefore=%lld",
                    S, C, (long long)n_src, (long long)n_dst, (long long)nbefore);
    if (n_dst > n_src - nbefore)
        return fail(ADN_ERR_INVALID, "adn_envelope_f64: n_dst=%lld exceeds n_src-nbefore=%lld",
                    (long long)n_dst, (long long)(n_src - nbefore));
    if ((S > 0 && !sos) || (n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_envelope_f64: NULL pointer");
    if (S == 0) {                               // reference: sos is None -> zeros
        if (n_dst > 0) memset(dst, 0, (size_t)n_dst * C * 8);
        return ADN_OK;
    }
    if (n_src <= adn_sosfiltfilt_edge(sos, S))
        return fail(ADN_ERR_SHORT, "adn_envelope_f64: the length of the input (%lld) must be greater "
                    "than the sosfiltfilt pad length %d", (long long)n_src, adn_sosfiltfilt_edge(sos, S));
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    size_t in_b = (size_t)n_src * C * 8, out_b = (size_t)n_dst * C * 8;
    if ((rc = stage_in(src, in_b))) return rc;
    if ((rc = c.out.reserve(out_b ? out_b : 16))) return rc;
    if (n_dst > 0 &&
        (rc = envelope_dev(sos, S, c.in.as<double>(), n_src, C, nbefore, c.out.as<double>(), n_dst,
                           clamp_negative, c.stream)))
        return rc;
    return stage_out(dst, out_b);
}

int32_t adn_spectrogram_f64(const double* src, int64_t n_src, int32_t C, double rate, int32_t nfft,
                            int32_t hop, int32_t window_id, int32_t detrend_id, double* dst,
                            int64_t n_dst, int32_t out_db, int64_t* n_computed) {
    if (C < 1 || n_src < 0 || n_dst < 0 || nfft < 1 || hop < 1 || hop > nfft || !(rate > 0))
        return fail(ADN_ERR_INVALID, "adn_spectrogram_f64: C=%d n_src=%lld n_dst=%lld nfft=%d hop=%d rate=%g",
                    C, (long long)n_src, (long long)n_dst, nfft, hop, rate);
    if ((n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_spectrogram_f64: NULL pointer");
    if (n_computed) *n_computed = 0;
    if (n_dst == 0) return ADN_OK;
    int64_t nf = spectrogram_frames(n_src, n_dst, nfft, hop);
    size_t F = (size_t)nfft / 2 + 1;
    if (nf == 0) {                              // reference: dest[:] = 0
        memset(dst, 0, (size_t)n_dst * C * F * 8);
        return ADN_OK;
    }
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    int64_t nsource = (nf - 1) * (int64_t)hop + nfft;      // rows the frames actually read
    size_t in_b = (size_t)nsource * C * 8, out_b = (size_t)nf * C * F * 8;
    if ((rc = stage_in(src, in_b))) return rc;
    if ((rc = c.out.reserve(out_b))) return rc;
    int64_t got = 0;
    if ((rc = spectrogram_dev(c.in.as<double>(), nsource, C, rate, nfft, hop, window_id, detrend_id,
                              c.out.as<double>(), nf, out_db, &got, c.stream)))
        return rc;
    if ((rc = stage_out(dst, out_b))) return rc;
    if (n_dst > nf) memset(dst + (size_t)nf * C * F, 0, (size_t)(n_dst - nf) * C * F * 8);
    if (n_computed) *n_computed = nf;
    return ADN_OK;
}

int32_t adn_decibel_f64(const double* power, int64_t n, double ref_power, double min_power,
                        double* dst) {
    if (n < 0) return fail(ADN_ERR_INVALID, "adn_decibel_f64: n=%lld", (long long)n);
    if (n == 0) return ADN_OK;
    if (!power || !dst) return fail(ADN_ERR_INVALID, "adn_decibel_f64: NULL pointer");
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    if ((rc = stage_in(power, (size_t)n * 8))) return rc;
    if ((rc = c.out.reserve((size_t)n * 8))) return rc;
    if ((rc = decibel_dev(c.in.as<double>(), n, ref_power, min_power, c.out.as<double>(), c.stream)))
        return rc;
    return stage_out(dst, (size_t)n * 8);
}

// ---------------------------------------------------------------- device entry points

int32_t adn_minmax_f64_dev(const double* src, int64_t n, int32_t C, int64_t step, double* dst,
                           void* stream) {
    if (n < 0 || C < 1 || step < 1) return fail(ADN_ERR_INVALID, "adn_minmax_f64_dev: bad shape");
    if (n == 0) return ADN_OK;
    if (!src || !dst) return fail(ADN_ERR_INVALID, "adn_minmax_f64_dev: NULL pointer");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return minmax_dev(src, n, C, step, dst, pick(stream));
}

int32_t adn_sosfilt_f64_dev(const double* sos, int32_t S, const double* src, int64_t n_src,
                            int32_t C, int64_t nbefore, double* dst, int64_t n_dst,
                            const double* zi, double* zf, void* stream) {
    if (S < 0 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 0 || nbefore < 0 ||
        (dst && (n_dst < 0 || n_dst > n_src - nbefore)))
        return fail(ADN_ERR_INVALID, "adn_sosfilt_f64_dev: bad shape");
    if ((S > 0 && !sos) || (n_src > 0 && !src)) return fail(ADN_ERR_INVALID, "adn_sosfilt_f64_dev: NULL pointer");
    if (n_src == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    return sosfilt_dev(sos, S, src, n_src, C, nbefore, dst, dst ? n_dst : 0, zi, zf, pick(stream));
}

int32_t adn_envelope_f64_dev(const double* sos, int32_t S, const double* src, int64_t n_src,
                             int32_t C, int64_t nbefore, double* dst, int64_t n_dst,
                             int32_t clamp_negative, void* stream) {
    if (S < 0 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 0 || n_dst < 0 || nbefore < 0 ||
        n_dst > n_src - nbefore)
        return fail(ADN_ERR_INVALID, "adn_envelope_f64_dev: bad shape");
    if ((S > 0 && !sos) || (n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_envelope_f64_dev: NULL pointer");
    if (n_dst == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    if (S == 0) {
        ADN_CK(cudaMemsetAsync(dst, 0, (size_t)n_dst * C * 8, pick(stream)));
        return ADN_OK;
    }
    if (n_src <= adn_sosfiltfilt_edge(sos, S))
        return fail(ADN_ERR_SHORT, "adn_envelope_f64_dev: input not longer than the sosfiltfilt pad");
    return envelope_dev(sos, S, src, n_src, C, nbefore, dst, n_dst, clamp_negative, pick(stream));
}

int32_t adn_spectrogram_f64_dev(const double* src, int64_t n_src, int32_t C, double rate,
                                int32_t nfft, int32_t hop, int32_t window_id, int32_t detrend_id,
                                double* dst, int64_t n_dst, int32_t out_db, int64_t* n_computed,
                                void* stream) {
    if (C < 1 || n_src < 0 || n_dst < 0 || nfft < 1 || hop < 1 || hop > nfft || !(rate > 0))
        return fail(ADN_ERR_INVALID, "adn_spectrogram_f64_dev: bad shape");
    if ((n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_spectrogram_f64_dev: NULL pointer");
    if (n_computed) *n_computed = 0;
    if (n_dst == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    return spectrogram_dev(src, n_src, C, rate, nfft, hop, window_id, detrend_id, dst, n_dst,
                           out_db, n_computed, pick(stream));
}

int32_t adn_decibel_f64_dev(const double* power, int64_t n, double ref_power, double min_power,
                            double* dst, void* stream) {
    if (n < 0) return fail(ADN_ERR_INVALID, "adn_decibel_f64_dev: n<0");
    if (n == 0) return ADN_OK;
    if (!power || !dst) return fail(ADN_ERR_INVALID, "adn_decibel_f64_dev: NULL pointer");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return decibel_dev(power, n, ref_power, min_power, dst, pick(stream));
}

int32_t adn_synth_f64_dev(double* dst, int64_t t0, int64_t n, int32_t C, double rate,
                          uint64_t seed, void* stream) {
    if (n < 0 || C < 1 || t0 < 0 || !(rate > 0)) return fail(ADN_ERR_INVALID, "adn_synth_f64_dev: bad shape");
    if (n == 0) return ADN_OK;
    if (!dst) return fail(ADN_ERR_INVALID, "adn_synth_f64_dev: NULL pointer");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return synth_dev(dst, t0, n, C, rate, seed, pick(stream));
}

}  // extern "C"
