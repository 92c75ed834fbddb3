// C ABI of libaudian_b200.so: context, error state, host-pointer entry points.
// Declarations and the reference functions each entry replaces: include/audian_b200.h
#include "sos_common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <map>
#include <array>

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
// scratch memory of the kernels' launchers, one set per stream: work on different streams
// never shares tile records or work buffers (a std::map never moves its elements)
static std::map<cudaStream_t, std::array<DevBuf, SCR_COUNT>> g_scratch;
static std::mutex g_scratch_mu;
static std::mutex g_mu;

Ctx& ctx() { return g_ctx; }
DevBuf& scratch(int slot, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    return g_scratch[st][slot];
}

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

// ---------------------------------------------------------------- host <-> device plumbing
//
// Host-pointer entry points move data on three streams -- uploads, kernels, downloads -- in
// chunks, so that on pinned host memory the two PCIe directions and the kernels overlap.
//
// Device copies of results are handed from a producer to its consumers EXPLICITLY: a caller
// that owns a buffer (a derived trace of audian_b200) creates a mirror (adn_mirror_create) and
// passes its handle as `dst_mirror` to the call that fills the buffer; the result then also
// stays in the mirror's device buffer, tagged with the host range it was copied to.  A consumer
// passes the same handle as `src_mirror` together with its host source pointer: if the source
// range lies inside the mirror's valid range the upload is skipped.  The owner invalidates the
// mirror (adn_mirror_invalidate) whenever the host buffer changes by other means -- the trace
// classes do so in move_buffer / allocate_buffer / reload_buffer.  Calls without a mirror
// (handle 0: every plain entry point) never look at device copies, so an array that was freed
// and reallocated, or edited in place, can never be served from stale device data.  A mirror
// is marked valid only after the computation and the download succeeded.

struct Mirror {
    DevBuf buf;
    const char* host = nullptr;        // host range the device buffer mirrors
    size_t bytes = 0;
    bool valid = false;
};

static std::map<int64_t, Mirror> g_mirrors;
static int64_t g_next_mirror = 1;
static int64_t g_opt[ADN_OPT_COUNT] = {1, 0, (int64_t)32 << 20, (int64_t)1 << 20, (int64_t)16 << 30,
                                       0, 1, 1};
int64_t option(int32_t which) { return g_opt[which]; }
static cudaStream_t g_h2d = nullptr, g_d2h = nullptr;
static std::vector<cudaEvent_t> g_events;
static size_t g_event_next = 0;
static int64_t g_res_hits = 0, g_res_misses = 0;
static int64_t g_bytes_h2d = 0, g_bytes_d2h = 0;   // moved by the host-pointer entry points
// host-pointer entry points share the staging buffers, the three streams and the mirrors:
// one call at a time (ctypes releases the GIL, so Python threads do get here concurrently)
static std::recursive_mutex g_api_mu;
#define ADN_API_LOCK std::lock_guard<std::recursive_mutex> api_lock__(g_api_mu)


static cudaError_t h2d(void* dev, const void* host, size_t bytes, cudaStream_t st) {
    g_bytes_h2d += (int64_t)bytes;
    return cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st);
}
static cudaError_t d2h(void* host, const void* dev, size_t bytes, cudaStream_t st) {
    g_bytes_d2h += (int64_t)bytes;
    return cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st);
}

static int32_t ensure_streams() {
    if (!g_h2d) ADN_CK(cudaStreamCreateWithFlags(&g_h2d, cudaStreamNonBlocking));
    if (!g_d2h) ADN_CK(cudaStreamCreateWithFlags(&g_d2h, cudaStreamNonBlocking));
    g_event_next = 0;
    return ADN_OK;
}

static int32_t next_event(cudaEvent_t* e) {
    if (g_event_next == g_events.size()) {
        cudaEvent_t ev;
        ADN_CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        g_events.push_back(ev);
    }
    *e = g_events[g_event_next++];
    return ADN_OK;
}

// b waits for everything enqueued on a so far
static int32_t chain(cudaStream_t a, cudaStream_t b) {
    cudaEvent_t e;
    int32_t rc = next_event(&e);
    if (rc) return rc;
    ADN_CK(cudaEventRecord(e, a));
    ADN_CK(cudaStreamWaitEvent(b, e, 0));
    return ADN_OK;
}

static Mirror* mirror_of(int64_t handle) {
    if (handle <= 0 || !g_opt[ADN_OPT_RESIDENT]) return nullptr;
    auto it = g_mirrors.find(handle);
    return it == g_mirrors.end() ? nullptr : &it->second;
}

static void release_residents() {
    for (auto& kv : g_mirrors) kv.second.buf.release();
    g_mirrors.clear();
    for (auto e : g_events) cudaEventDestroy(e);
    g_events.clear();
    if (g_h2d) { cudaStreamDestroy(g_h2d); g_h2d = nullptr; }
    if (g_d2h) { cudaStreamDestroy(g_d2h); g_d2h = nullptr; }
}

// Device buffer the result for host range [dst, dst + bytes) is computed into: the mirror's
// (left INVALID until mirror_commit) or the shared staging buffer.
static int32_t out_buffer(int64_t dst_mirror, void* dst, size_t bytes, double** dev) {
    Ctx& c = ctx();
    Mirror* m = mirror_of(dst_mirror);
    if (m) {
        m->valid = false;
        if ((int64_t)bytes >= g_opt[ADN_OPT_RESIDENT_MIN_BYTES]) {
            size_t total = 0;
            for (auto& kv : g_mirrors) if (&kv.second != m) total += kv.second.buf.cap;
            if ((int64_t)(total + bytes) <= g_opt[ADN_OPT_RESIDENT_CAP_BYTES] &&
                m->buf.reserve(bytes) == ADN_OK) {
                *dev = m->buf.as<double>();
                return ADN_OK;
            }
            cudaGetLastError();          // out of device memory for a kept copy: staging buffer
        }
    }
    int32_t rc = c.out.reserve(bytes ? bytes : 16);
    if (rc) return rc;
    *dev = c.out.as<double>();
    return ADN_OK;
}

// the result in `dev` has reached the host range: from now on the mirror may serve it
static void mirror_commit(int64_t dst_mirror, const void* dst, size_t bytes, const double* dev) {
    Mirror* m = mirror_of(dst_mirror);
    if (!m || m->buf.p != dev || bytes == 0) return;
    m->host = static_cast<const char*>(dst);
    m->bytes = bytes;
    m->valid = true;
}

// Device copy of the host range [src, src + bytes) if the mirror holds it.
static int32_t in_resident(int64_t src_mirror, const void* src, size_t bytes, const double** dev) {
    *dev = nullptr;
    Mirror* m = mirror_of(src_mirror);
    if (!m) return ADN_OK;
    const char* lo = static_cast<const char*>(src);
    if (m->valid && lo >= m->host && lo + bytes <= m->host + m->bytes) {
        *dev = reinterpret_cast<const double*>(static_cast<const char*>(m->buf.p) + (lo - m->host));
        ++g_res_hits;
    } else {
        ++g_res_misses;
    }
    return ADN_OK;
}

// Whole upload of `bytes` into the `in` staging buffer on the kernel stream.
static int32_t stage_in(const void* host, size_t bytes) {
    Ctx& c = ctx();
    int32_t rc = c.in.reserve(bytes ? bytes : 16);
    if (rc) return rc;
    if (bytes) ADN_CK(h2d(c.in.p, host, bytes, c.stream));
    return ADN_OK;
}

// Source of a host-pointer call: the resident copy, or a whole upload.
static int32_t source_dev(int64_t src_mirror, const void* host, size_t bytes, const double** dev) {
    int32_t rc = in_resident(src_mirror, host, bytes, dev);
    if (rc || *dev) return rc;
    if ((rc = stage_in(host, bytes))) return rc;
    *dev = ctx().in.as<double>();
    return ADN_OK;
}

static int32_t copy_out(void* host, const double* dev, size_t bytes) {
    Ctx& c = ctx();
    if (bytes) ADN_CK(d2h(host, dev, bytes, c.stream));
    ADN_CK(cudaStreamSynchronize(c.stream));
    return ADN_OK;
}

static int32_t sync_pipeline() {
    ADN_CK(cudaStreamSynchronize(g_h2d));
    ADN_CK(cudaStreamSynchronize(ctx().stream));
    ADN_CK(cudaStreamSynchronize(g_d2h));
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
    ADN_API_LOCK;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx.ready) return ADN_OK;
    cudaSetDevice(g_ctx.device);
    cudaStreamSynchronize(g_ctx.stream);
    g_ctx.in.release();
    g_ctx.out.release();
    g_ctx.aux.release();
    release_residents();
    {
        std::lock_guard<std::mutex> lk2(g_scratch_mu);
        for (auto& kv : g_scratch) for (auto& b : kv.second) b.release();
        g_scratch.clear();
    }
    cudaStreamDestroy(g_ctx.stream);
    g_ctx.stream = nullptr;
    g_ctx.ready = false;
    g_ctx.device = -1;
    return ADN_OK;
}

const char* adn_last_error(void) { return g_err.c_str(); }
int32_t adn_version(void) { return 100; }
int64_t adn_launch_count(void) { return g_ctx.launches.load(); }
int64_t adn_scan_run_count(void) { return adn::scan_run_launches(); }
int64_t adn_fwd_park_count(void) { return adn::fwd_park_launches(); }
int64_t adn_zero_phase_count(void) { return adn::zp_launches(); }

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

int32_t adn_host_alloc(int64_t bytes, void** ptr) {
    if (!ptr) return fail(ADN_ERR_INVALID, "adn_host_alloc: NULL pointer");
    *ptr = nullptr;
    if (bytes <= 0) return fail(ADN_ERR_INVALID, "adn_host_alloc: %lld bytes", (long long)bytes);
    int32_t rc = ensure_init();
    if (rc) return rc;
    void* p = nullptr;
    ADN_CK(cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable));
    *ptr = p;
    return ADN_OK;
}

int32_t adn_host_free(void* ptr) {
    if (!ptr) return ADN_OK;
    ADN_CK(cudaFreeHost(ptr));
    return ADN_OK;
}

// ---------------------------------------------------------------- host entry points

int32_t adn_set_option(int32_t option, int64_t value) {
    if (option < 0 || option >= ADN_OPT_COUNT) return fail(ADN_ERR_INVALID, "adn_set_option: option %d", option);
    ADN_API_LOCK;
    g_opt[option] = value;
    if (option == ADN_OPT_RESIDENT && value == 0)
        for (auto& kv : g_mirrors) kv.second.valid = false;
    return ADN_OK;
}

int64_t adn_get_option(int32_t option) {
    if (option < 0 || option >= ADN_OPT_COUNT) return -1;
    return g_opt[option];
}

int32_t adn_mirror_create(int64_t* handle) {
    if (!handle) return fail(ADN_ERR_INVALID, "adn_mirror_create: NULL pointer");
    ADN_API_LOCK;
    *handle = g_next_mirror++;
    g_mirrors[*handle] = Mirror();
    return ADN_OK;
}

int32_t adn_mirror_release(int64_t handle) {
    ADN_API_LOCK;
    auto it = g_mirrors.find(handle);
    if (it == g_mirrors.end()) return ADN_OK;
    if (it->second.buf.p && g_ctx.ready) {
        cudaSetDevice(g_ctx.device);
        cudaStreamSynchronize(g_ctx.stream);
    }
    it->second.buf.release();
    g_mirrors.erase(it);
    return ADN_OK;
}

int32_t adn_mirror_invalidate(int64_t handle) {
    ADN_API_LOCK;
    auto it = g_mirrors.find(handle);
    if (it != g_mirrors.end()) it->second.valid = false;
    return ADN_OK;
}

int32_t adn_invalidate(const void* host, int64_t bytes) {
    ADN_API_LOCK;
    if (!host || bytes <= 0) return ADN_OK;
    const char* lo = static_cast<const char*>(host);
    const char* hi = lo + bytes;
    for (auto& kv : g_mirrors) {
        Mirror& m = kv.second;
        if (m.valid && m.host < hi && lo < m.host + m.bytes) m.valid = false;
    }
    return ADN_OK;
}

int64_t adn_resident_hits(void) { return g_res_hits; }

int32_t adn_transfer_bytes(int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (h2d_bytes) *h2d_bytes = g_bytes_h2d;
    if (d2h_bytes) *d2h_bytes = g_bytes_d2h;
    return ADN_OK;
}

int32_t adn_minmax_f64_m(const double* src, int64_t n, int32_t C, int64_t step, double* dst,
                         int64_t src_mirror) {
    if (n < 0 || C < 1 || step < 1) return fail(ADN_ERR_INVALID, "adn_minmax_f64: n=%lld C=%d step=%lld",
                                               (long long)n, C, (long long)step);
    if (n == 0) return ADN_OK;
    if (!src || !dst) return fail(ADN_ERR_INVALID, "adn_minmax_f64: NULL pointer");
    ADN_API_LOCK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    int64_t nseg = (n + step - 1) / step;
    size_t in_b = (size_t)n * C * 8, out_b = (size_t)nseg * 2 * C * 8;
    const double* dsrc = nullptr;
    if ((rc = in_resident(src_mirror, src, in_b, &dsrc))) return rc;
    if ((rc = c.out.reserve(out_b))) return rc;
    if (dsrc) {
        if ((rc = minmax_dev(dsrc, n, C, step, c.out.as<double>(), c.stream))) return rc;
        return copy_out(dst, c.out.as<double>(), out_b);
    }
    // upload in chunks of whole segments; the kernel of a chunk runs while the next one arrives
    if ((rc = ensure_streams())) return rc;
    if ((rc = c.in.reserve(in_b))) return rc;
    int64_t rows = g_opt[ADN_OPT_CHUNK_BYTES] / ((int64_t)C * 8);
    int64_t segs = rows / step;
    if (segs < 1) segs = 1;
    for (int64_t s0 = 0; s0 < nseg; s0 += segs) {
        const int64_t a = s0 * step, b = (s0 + segs) * step < n ? (s0 + segs) * step : n;
        ADN_CK(h2d(c.in.as<double>() + a * C, src + a * C, (size_t)(b - a) * C * 8, g_h2d));
        if ((rc = chain(g_h2d, c.stream))) return rc;
        if ((rc = minmax_dev(c.in.as<double>() + a * C, b - a, C, step, c.out.as<double>() + 2 * s0 * C,
                             c.stream)))
            return rc;
    }
    return copy_out(dst, c.out.as<double>(), out_b);
}

int32_t adn_minmax_f64(const double* src, int64_t n, int32_t C, int64_t step, double* dst) {
    return adn_minmax_f64_m(src, n, C, step, dst, 0);
}

int32_t adn_sosfilt_f64_m(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                          int64_t nbefore, double* dst, int64_t n_dst, double* zi_inout,
                          int64_t src_mirror, int64_t dst_mirror) {
    if (S < 0 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 0 || n_dst < 0 || nbefore < 0)
        return fail(ADN_ERR_INVALID, "adn_sosfilt_f64: S=%d C=%d n_src=%lld n_dst=%lld nbefore=%lld",
                    S, C, (long long)n_src, (long long)n_dst, (long long)nbefore);
    if (n_dst > n_src - nbefore)
        return fail(ADN_ERR_INVALID, "adn_sosfilt_f64: n_dst=%lld exceeds n_src-nbefore=%lld",
                    (long long)n_dst, (long long)(n_src - nbefore));
    if ((S > 0 && !sos) || (n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_sosfilt_f64: NULL pointer");
    if (n_src == 0) return ADN_OK;
    ADN_API_LOCK;
    size_t in_b = (size_t)n_src * C * 8, out_b = (size_t)n_dst * C * 8;
    if (S == 0) {                               // reference: sos is None -> dest = source[nbefore:]
        adn_mirror_invalidate(dst_mirror);
        if (n_dst > 0) memmove(dst, src + nbefore * C, out_b);
        return ADN_OK;
    }
    int32_t rc = ensure_init();
    if (rc) return rc;
    if ((rc = ensure_streams())) return rc;
    Ctx& c = ctx();
    const size_t z_b = (size_t)C * S * 2 * 8;
    // aux: [user zi][state A][state B]
    if ((rc = c.aux.reserve(3 * z_b + 4096))) return rc;
    double* d_zi = c.aux.as<double>();
    double* d_state[2] = {d_zi + (size_t)C * S * 2, d_zi + (size_t)C * S * 4};
    if (zi_inout) ADN_CK(h2d(d_zi, zi_inout, z_b, c.stream));
    const double* dsrc = nullptr;
    if ((rc = in_resident(src_mirror, src, in_b, &dsrc))) return rc;
    double* dout = nullptr;
    if ((rc = out_buffer(dst_mirror, dst, out_b, &dout))) return rc;
    if (dsrc) {
        if ((rc = sosfilt_dev(sos, S, dsrc, n_src, C, nbefore, n_dst > 0 ? dout : nullptr, n_dst,
                              zi_inout ? d_zi : nullptr, zi_inout ? d_state[0] : nullptr, c.stream)))
            return rc;
        if (zi_inout) ADN_CK(d2h(zi_inout, d_state[0], z_b, c.stream));
        if ((rc = copy_out(dst, dout, out_b))) return rc;
        mirror_commit(dst_mirror, dst, out_b, dout);
        return ADN_OK;
    }
    // chunks of rows: upload k+1, filter k (state carried from chunk to chunk: equal to one pass,
    // the streamed == one-shot property of the scan), download k-1
    if ((rc = c.in.reserve(in_b))) return rc;
    double* din = c.in.as<double>();
    int64_t R = g_opt[ADN_OPT_CHUNK_BYTES] / ((int64_t)C * 8);
    if (R < 4096) R = 4096;
    const double* zin = zi_inout ? d_zi : nullptr;
    // short first chunks (R/8, R/4, R/2, then R rows), the same schedule as in adn_chain_f64: the
    // first results start their way down early, and both calls cut the trace at the same rows
    int64_t Rk = R / 8 < 4096 ? 4096 : R / 8;
    int64_t a = 0;
    for (int64_t k = 0; a < n_src; ++k) {
        const int64_t b = a + Rk < n_src ? a + Rk : n_src;
        if (Rk < R) Rk = 2 * Rk < R ? 2 * Rk : R;
        ADN_CK(h2d(din + a * C, src + a * C, (size_t)(b - a) * C * 8, g_h2d));
        if ((rc = chain(g_h2d, c.stream))) return rc;
        int64_t nb = nbefore - a;
        if (nb < 0) nb = 0;
        if (nb > b - a) nb = b - a;
        const int64_t o0 = a + nb - nbefore;                      // first output row of the chunk
        int64_t no = b - a - nb;
        if (no > n_dst - o0) no = n_dst - o0;
        if (no < 0) no = 0;
        const bool last = b == n_src;
        double* zout = (last && !zi_inout) ? nullptr : d_state[k & 1];
        if (no > 0 || zout) {
            if ((rc = sosfilt_dev(sos, S, din + a * C, b - a, C, nb, no > 0 ? dout + o0 * C : nullptr, no,
                                  zin, zout, c.stream)))
                return rc;
        }
        zin = zout;
        if (no > 0) {
            if ((rc = chain(c.stream, g_d2h))) return rc;
            ADN_CK(d2h(dst + o0 * C, dout + o0 * C, (size_t)no * C * 8, g_d2h));
        }
        if (last && zi_inout)
            ADN_CK(d2h(zi_inout, zout, z_b, c.stream));
        a = b;
    }
    if ((rc = sync_pipeline())) return rc;
    mirror_commit(dst_mirror, dst, out_b, dout);
    return ADN_OK;
}

int32_t adn_sosfilt_f64(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                        int64_t nbefore, double* dst, int64_t n_dst, double* zi_inout) {
    return adn_sosfilt_f64_m(sos, S, src, n_src, C, nbefore, dst, n_dst, zi_inout, 0, 0);
}

int32_t adn_envelope_f64_m(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                           int64_t nbefore, double* dst, int64_t n_dst, int32_t clamp_negative,
                           int64_t src_mirror, int64_t dst_mirror) {
    if (S < 0 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 0 || n_dst < 0 || nbefore < 0)
        return fail(ADN_ERR_INVALID, "adn_envelope_f64: S=%d C=%d n_src=%lld n_dst=%lld nbefore=%lld",
                    S, C, (long long)n_src, (long long)n_dst, (long long)nbefore);
    if (n_dst > n_src - nbefore)
        return fail(ADN_ERR_INVALID, "adn_envelope_f64: n_dst=%lld exceeds n_src-nbefore=%lld",
                    (long long)n_dst, (long long)(n_src - nbefore));
    if ((S > 0 && !sos) || (n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_envelope_f64: NULL pointer");
    ADN_API_LOCK;
    size_t in_b = (size_t)n_src * C * 8, out_b = (size_t)n_dst * C * 8;
    if (S == 0) {                               // reference: sos is None -> zeros
        adn_mirror_invalidate(dst_mirror);
        if (n_dst > 0) memset(dst, 0, out_b);
        return ADN_OK;
    }
    if (n_src <= adn_sosfiltfilt_edge(sos, S))
        return fail(ADN_ERR_SHORT, "adn_envelope_f64: the length of the input (%lld) must be greater "
                    "than the sosfiltfilt pad length %d", (long long)n_src, adn_sosfiltfilt_edge(sos, S));
    int32_t rc = ensure_init();
    if (rc) return rc;
    const double* dsrc = nullptr;
    if ((rc = source_dev(src_mirror, src, in_b, &dsrc))) return rc;
    double* dout = nullptr;
    if ((rc = out_buffer(dst_mirror, dst, out_b, &dout))) return rc;
    if (n_dst > 0 &&
        (rc = envelope_dev(sos, S, dsrc, n_src, C, nbefore, dout, n_dst, clamp_negative, ctx().stream)))
        return rc;
    if ((rc = copy_out(dst, dout, out_b))) return rc;
    mirror_commit(dst_mirror, dst, out_b, dout);
    return ADN_OK;
}

int32_t adn_envelope_f64(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                         int64_t nbefore, double* dst, int64_t n_dst, int32_t clamp_negative) {
    return adn_envelope_f64_m(sos, S, src, n_src, C, nbefore, dst, n_dst, clamp_negative, 0, 0);
}

int32_t adn_sosfiltfilt_f64(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                            double* dst, int64_t n_dst) {
    if (S < 1 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 0 || n_dst < 0 || n_dst > n_src)
        return fail(ADN_ERR_INVALID, "adn_sosfiltfilt_f64: S=%d C=%d n_src=%lld n_dst=%lld", S, C,
                    (long long)n_src, (long long)n_dst);
    if (!sos || (n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_sosfiltfilt_f64: NULL pointer");
    if (n_src <= adn_sosfiltfilt_edge(sos, S))
        return fail(ADN_ERR_SHORT, "adn_sosfiltfilt_f64: the length of the input (%lld) must be greater "
                    "than the pad length %d", (long long)n_src, adn_sosfiltfilt_edge(sos, S));
    if (n_dst == 0) return ADN_OK;
    ADN_API_LOCK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    size_t in_b = (size_t)n_src * C * 8, out_b = (size_t)n_dst * C * 8;
    const double* dsrc = nullptr;
    if ((rc = source_dev(0, src, in_b, &dsrc))) return rc;
    double* dout = nullptr;
    if ((rc = out_buffer(0, dst, out_b, &dout))) return rc;
    if ((rc = sosfiltfilt_dev(sos, S, dsrc, n_src, C, 0, dout, n_dst, ctx().stream))) return rc;
    return copy_out(dst, dout, out_b);
}

int32_t adn_spectrogram_f64_m(const double* src, int64_t n_src, int32_t C, double rate, int32_t nfft,
                              int32_t hop, int32_t window_id, int32_t detrend_id, double* dst,
                              int64_t n_dst, int32_t out_db, int64_t* n_computed,
                              int64_t src_mirror, int64_t dst_mirror) {
    if (C < 1 || n_src < 0 || n_dst < 0 || nfft < 1 || hop < 1 || hop > nfft || !(rate > 0))
        return fail(ADN_ERR_INVALID, "adn_spectrogram_f64: C=%d n_src=%lld n_dst=%lld nfft=%d hop=%d rate=%g",
                    C, (long long)n_src, (long long)n_dst, nfft, hop, rate);
    if ((n_src > 0 && !src) || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_spectrogram_f64: NULL pointer");
    if (n_computed) *n_computed = 0;
    if (n_dst == 0) return ADN_OK;
    ADN_API_LOCK;
    int64_t nf = spectrogram_frames(n_src, n_dst, nfft, hop);
    size_t F = (size_t)nfft / 2 + 1;
    if (nf == 0) {                              // reference: dest[:] = 0
        adn_mirror_invalidate(dst_mirror);
        memset(dst, 0, (size_t)n_dst * C * F * 8);
        return ADN_OK;
    }
    int32_t rc = ensure_init();
    if (rc) return rc;
    if ((rc = ensure_streams())) return rc;
    Ctx& c = ctx();
    int64_t nsource = (nf - 1) * (int64_t)hop + nfft;      // rows the frames actually read
    size_t in_b = (size_t)nsource * C * 8, out_b = (size_t)nf * C * F * 8;
    const double* dsrc = nullptr;
    if ((rc = in_resident(src_mirror, src, in_b, &dsrc))) return rc;
    double* dout = nullptr;
    // the mirror keeps all n_dst frames (the zero-filled ones too): consumers ask for the whole buffer
    if ((rc = out_buffer(dst_mirror, dst, (size_t)n_dst * C * F * 8, &dout))) return rc;
    // chunks of frames: upload the rows chunk k adds, transform chunk k, download chunk k-1
    const bool upload = dsrc == nullptr;
    if (upload) {
        if ((rc = c.in.reserve(in_b))) return rc;
        dsrc = c.in.as<double>();
    }
    int64_t FR = g_opt[ADN_OPT_CHUNK_BYTES] / ((int64_t)C * (int64_t)F * 8);
    if (FR < 16) FR = 16;
    int64_t up = 0;                                        // rows uploaded so far
    for (int64_t f0 = 0; f0 < nf; f0 += FR) {
        const int64_t fc = f0 + FR < nf ? FR : nf - f0;
        const int64_t r1 = (f0 + fc - 1) * hop + nfft;     // rows [0, r1) are needed
        if (upload && r1 > up) {
            ADN_CK(h2d(c.in.as<double>() + up * C, src + up * C, (size_t)(r1 - up) * C * 8, g_h2d));
            up = r1;
            if ((rc = chain(g_h2d, c.stream))) return rc;
        }
        int64_t got = 0;
        if ((rc = spectrogram_dev(dsrc + f0 * hop * C, (fc - 1) * hop + nfft, C, rate, nfft, hop, window_id,
                                  detrend_id, dout + (size_t)f0 * C * F, fc, out_db, &got, c.stream)))
            return rc;
        if ((rc = chain(c.stream, g_d2h))) return rc;
        ADN_CK(d2h(dst + (size_t)f0 * C * F, dout + (size_t)f0 * C * F, (size_t)fc * C * F * 8, g_d2h));
    }
    if (n_dst > nf && mirror_of(dst_mirror) && dout == mirror_of(dst_mirror)->buf.p)
        ADN_CK(cudaMemsetAsync(dout + (size_t)nf * C * F, 0, (size_t)(n_dst - nf) * C * F * 8, c.stream));
    if ((rc = sync_pipeline())) return rc;
    if (n_dst > nf) memset(dst + (size_t)nf * C * F, 0, (size_t)(n_dst - nf) * C * F * 8);
    mirror_commit(dst_mirror, dst, (size_t)n_dst * C * F * 8, dout);
    if (n_computed) *n_computed = nf;
    return ADN_OK;
}

int32_t adn_spectrogram_f64(const double* src, int64_t n_src, int32_t C, double rate, int32_t nfft,
                            int32_t hop, int32_t window_id, int32_t detrend_id, double* dst,
                            int64_t n_dst, int32_t out_db, int64_t* n_computed) {
    return adn_spectrogram_f64_m(src, n_src, C, rate, nfft, hop, window_id, detrend_id, dst, n_dst, out_db,
                                 n_computed, 0, 0);
}

int32_t adn_decibel_f64(const double* power, int64_t n, double ref_power, double min_power,
                        double* dst) {
    if (n < 0) return fail(ADN_ERR_INVALID, "adn_decibel_f64: n=%lld", (long long)n);
    if (n == 0) return ADN_OK;
    if (!power || !dst) return fail(ADN_ERR_INVALID, "adn_decibel_f64: NULL pointer");
    ADN_API_LOCK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    const double* dsrc = nullptr;
    if ((rc = source_dev(0, power, (size_t)n * 8, &dsrc))) return rc;
    if ((rc = c.out.reserve((size_t)n * 8))) return rc;
    if ((rc = decibel_dev(dsrc, n, ref_power, min_power, c.out.as<double>(), c.stream)))
        return rc;
    return copy_out(dst, c.out.as<double>(), (size_t)n * 8);
}

// one channel of a (n, C, F) spectrogram buffer on the device: the mirror's copy of the whole
// buffer if there is one (then *Cd = C, *chd = channel), else only that channel's rows are
// uploaded with a strided copy (*Cd = 1, *chd = 0)
static int32_t spec_channel_dev(int64_t src_mirror, const double* spec, int64_t n, int32_t C, int32_t F,
                                int32_t channel, const double** dev, int32_t* Cd, int32_t* chd) {
    int32_t rc = in_resident(src_mirror, spec, (size_t)n * C * F * 8, dev);
    if (rc) return rc;
    if (*dev) { *Cd = C; *chd = channel; return ADN_OK; }
    Ctx& c = ctx();
    if ((rc = c.in.reserve((size_t)n * F * 8))) return rc;
    g_bytes_h2d += (int64_t)n * F * 8;
    ADN_CK(cudaMemcpy2DAsync(c.in.p, (size_t)F * 8, spec + (size_t)channel * F, (size_t)C * F * 8,
                             (size_t)F * 8, (size_t)n, cudaMemcpyHostToDevice, c.stream));
    *dev = c.in.as<double>();
    *Cd = 1;
    *chd = 0;
    return ADN_OK;
}

int32_t adn_spec_image_db_f64_m(const double* spec, int64_t n, int32_t C, int32_t F, int32_t channel,
                                double* dst, int64_t src_mirror) {
    if (n < 0 || C < 1 || F < 1 || channel < 0 || channel >= C)
        return fail(ADN_ERR_INVALID, "adn_spec_image_db_f64: n=%lld C=%d F=%d channel=%d", (long long)n, C, F, channel);
    if (n == 0) return ADN_OK;
    if (!spec || !dst) return fail(ADN_ERR_INVALID, "adn_spec_image_db_f64: NULL pointer");
    ADN_API_LOCK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    const double* d = nullptr;
    int32_t Cd = C, chd = channel;
    if ((rc = spec_channel_dev(src_mirror, spec, n, C, F, channel, &d, &Cd, &chd))) return rc;
    if ((rc = c.out.reserve((size_t)n * F * 8))) return rc;
    if ((rc = spec_image_dev(d, n, Cd, F, chd, c.out.as<double>(), c.stream))) return rc;
    return copy_out(dst, c.out.as<double>(), (size_t)n * F * 8);
}

int32_t adn_spec_image_db_f64(const double* spec, int64_t n, int32_t C, int32_t F, int32_t channel,
                              double* dst) {
    return adn_spec_image_db_f64_m(spec, n, C, F, channel, dst, 0);
}

int32_t adn_mean_power_db_f64_m(const double* spec, int64_t n, int32_t C, int32_t F, int32_t channel,
                                int64_t i0, int64_t i1, double floor_db, double* dst, int64_t src_mirror) {
    if (n < 1 || C < 1 || F < 1 || channel < 0 || channel >= C || i0 < 0 || i1 <= i0 || i1 > n)
        return fail(ADN_ERR_INVALID, "adn_mean_power_db_f64: n=%lld C=%d F=%d channel=%d i0=%lld i1=%lld",
                    (long long)n, C, F, channel, (long long)i0, (long long)i1);
    if (!spec || !dst) return fail(ADN_ERR_INVALID, "adn_mean_power_db_f64: NULL pointer");
    ADN_API_LOCK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    const double* d = nullptr;
    int32_t Cd = C, chd = channel;
    // without a device copy only the rows i0..i1 of the channel go up
    rc = in_resident(src_mirror, spec, (size_t)n * C * F * 8, &d);
    if (rc) return rc;
    int64_t a0 = i0, a1 = i1;
    if (!d) {
        if ((rc = spec_channel_dev(0, spec + (size_t)i0 * C * F, i1 - i0, C, F, channel, &d, &Cd, &chd))) return rc;
        a0 = 0;
        a1 = i1 - i0;
    }
    if ((rc = c.out.reserve((size_t)F * 8))) return rc;
    if ((rc = mean_power_dev(d, Cd, F, chd, a0, a1, floor_db, c.out.as<double>(), c.stream))) return rc;
    return copy_out(dst, c.out.as<double>(), (size_t)F * 8);
}

int32_t adn_mean_power_db_f64(const double* spec, int64_t n, int32_t C, int32_t F, int32_t channel,
                              int64_t i0, int64_t i1, double floor_db, double* dst) {
    return adn_mean_power_db_f64_m(spec, n, C, F, channel, i0, i1, floor_db, dst, 0);
}

int32_t adn_pcm_to_f64(const void* pcm, int64_t n, int32_t bits, double gain, double* dst) {
    if (n < 0 || (bits != 16 && bits != 24 && bits != 32))
        return fail(ADN_ERR_INVALID, "adn_pcm_to_f64: n=%lld bits=%d (16, 24 or 32)", (long long)n, bits);
    if (n == 0) return ADN_OK;
    if (!pcm || !dst) return fail(ADN_ERR_INVALID, "adn_pcm_to_f64: NULL pointer");
    ADN_API_LOCK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    const size_t in_b = (size_t)n * (bits / 8), out_b = (size_t)n * 8;
    if ((rc = stage_in(pcm, in_b))) return rc;
    double* dout = nullptr;
    if ((rc = out_buffer(0, dst, out_b, &dout))) return rc;
    if ((rc = pcm_dev(c.in.p, n, bits, gain, dout, c.stream))) return rc;
    return copy_out(dst, dout, out_b);
}

int32_t adn_minmax_channel_f64_m(const double* src, int64_t n, int32_t C, int32_t channel, int64_t step,
                                 double* dst, int64_t src_mirror) {
    if (n < 0 || C < 1 || step < 1 || channel < 0 || channel >= C)
        return fail(ADN_ERR_INVALID, "adn_minmax_channel_f64: n=%lld C=%d channel=%d step=%lld",
                    (long long)n, C, channel, (long long)step);
    if (n == 0) return ADN_OK;
    if (!src || !dst) return fail(ADN_ERR_INVALID, "adn_minmax_channel_f64: NULL pointer");
    ADN_API_LOCK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    const int64_t nseg = (n + step - 1) / step;
    const size_t col_b = (size_t)n * 8, out_b = (size_t)nseg * 2 * 8;
    const double* dsrc = nullptr;
    if ((rc = in_resident(src_mirror, src, (size_t)n * C * 8, &dsrc))) return rc;
    if ((rc = c.in.reserve(col_b))) return rc;
    if ((rc = c.out.reserve(out_b))) return rc;
    if (dsrc) {
        // the trace is on the device: gather the column there
        if ((rc = gather_channel_dev(dsrc, n, C, channel, c.in.as<double>(), c.stream))) return rc;
    } else {
        // only the column goes up (strided copy): n x 8 bytes instead of n x C x 8
        g_bytes_h2d += (int64_t)col_b;
        ADN_CK(cudaMemcpy2DAsync(c.in.p, 8, src + channel, (size_t)C * 8, 8, (size_t)n,
                                 cudaMemcpyHostToDevice, c.stream));
    }
    if ((rc = minmax_dev(c.in.as<double>(), n, 1, step, c.out.as<double>(), c.stream))) return rc;
    return copy_out(dst, c.out.as<double>(), out_b);
}

int32_t adn_unwrap_f64(double* data, int64_t n, int32_t C, double thresh, int32_t clips) {
    if (n < 0 || C < 1) return fail(ADN_ERR_INVALID, "adn_unwrap_f64: n=%lld C=%d", (long long)n, C);
    if (n == 0 || !(thresh > 0.0)) return ADN_OK;          // audioio: thresh <= 0 switches it off
    if (!data) return fail(ADN_ERR_INVALID, "adn_unwrap_f64: NULL pointer");
    ADN_API_LOCK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    Ctx& c = ctx();
    const size_t b = (size_t)n * C * 8;
    if ((rc = stage_in(data, b))) return rc;
    if ((rc = c.out.reserve(b))) return rc;
    if ((rc = unwrap_dev(c.in.as<double>(), n, C, thresh, clips, c.out.as<double>(), c.stream))) return rc;
    return copy_out(data, c.out.as<double>(), b);
}

static int32_t play_region_impl(const double* dsrc, int64_t n, int32_t C, const int32_t* left, int32_t nleft,
                                const int32_t* right, int32_t nright, double rate, double het_freq,
                                const double* sos, int32_t S, int64_t nstep, double* work, double* dout,
                                cudaStream_t st) {
    const int ncols = nright > 0 ? 2 : 1;
    const bool het = het_freq > 0.0;
    int32_t rc;
    if (!het) return play_mix_dev(dsrc, n, C, left, nleft, right, nright, rate, 0.0, dout, st);
    double* mixed = work;
    double* filt = work + (size_t)n * ncols;
    if ((rc = play_mix_dev(dsrc, n, C, left, nleft, right, nright, rate, het_freq, mixed, st))) return rc;
    if ((rc = sosfiltfilt_dev(sos, S, mixed, n, ncols, 0, filt, n, st))) return rc;
    return decimate_dev(filt, (n + nstep - 1) / nstep, ncols, nstep, dout, st);
}

static int32_t play_region_check(int64_t n, int32_t C, const int32_t* left, int32_t nleft, const int32_t* right,
                                 int32_t nright, double rate, double het_freq, const double* sos, int32_t S,
                                 int64_t nstep) {
    if (n < 0 || C < 1 || nleft < 1 || nleft > 64 || nright < 0 || nright > 64 || !left ||
        (nright > 0 && !right) || !(rate > 0) || nstep < 1)
        return fail(ADN_ERR_INVALID, "adn_play_region_f64: n=%lld C=%d nleft=%d nright=%d nstep=%lld",
                    (long long)n, C, nleft, nright, (long long)nstep);
    for (int i = 0; i < nleft; ++i) if (left[i] < 0 || left[i] >= C) return fail(ADN_ERR_INVALID, "adn_play_region_f64: channel %d", left[i]);
    for (int i = 0; i < nright; ++i) if (right[i] < 0 || right[i] >= C) return fail(ADN_ERR_INVALID, "adn_play_region_f64: channel %d", right[i]);
    if (het_freq > 0.0) {
        if (!sos || S < 1 || S > ADN_MAX_SECTIONS) return fail(ADN_ERR_INVALID, "adn_play_region_f64: heterodyne needs the low-pass sos");
        if (n <= adn_sosfiltfilt_edge(sos, S))
            return fail(ADN_ERR_SHORT, "adn_play_region_f64: the length of the input (%lld) must be greater "
                        "than the sosfiltfilt pad length %d", (long long)n, adn_sosfiltfilt_edge(sos, S));
    }
    return ADN_OK;
}

int32_t adn_play_region_f64_m(const double* src, int64_t n, int32_t C, const int32_t* left, int32_t nleft,
                              const int32_t* right, int32_t nright, double rate, double het_freq,
                              const double* sos, int32_t S, int64_t nstep, double* dst, int64_t src_mirror) {
    int32_t rc = play_region_check(n, C, left, nleft, right, nright, rate, het_freq, sos, S, nstep);
    if (rc) return rc;
    if (n == 0) return ADN_OK;
    if (!src || !dst) return fail(ADN_ERR_INVALID, "adn_play_region_f64: NULL pointer");
    ADN_API_LOCK;
    if ((rc = ensure_init())) return rc;
    Ctx& c = ctx();
    const int ncols = nright > 0 ? 2 : 1;
    if (!(het_freq > 0.0)) nstep = 1;
    const int64_t no = (n + nstep - 1) / nstep;
    const double* dsrc = nullptr;
    if ((rc = source_dev(src_mirror, src, (size_t)n * C * 8, &dsrc))) return rc;
    if ((rc = c.out.reserve((size_t)no * ncols * 8))) return rc;
    DevBuf& wb = scratch(SCR_PLAY, c.stream);
    if ((rc = wb.reserve((size_t)2 * n * ncols * 8 + 16))) return rc;
    if ((rc = play_region_impl(dsrc, n, C, left, nleft, right, nright, rate, het_freq, sos, S, nstep,
                               wb.as<double>(), c.out.as<double>(), c.stream)))
        return rc;
    return copy_out(dst, c.out.as<double>(), (size_t)no * ncols * 8);
}

// ---------------------------------------------------------------- the chain
// data -> filtered -> {spectrogram, envelope} (+ min/max of the raw rows) in one call: what a
// parameter change walks through in the reference (BufferedFilter.update -> recompute_all ->
// recompute of filtered, then of its dests: src/audian/bufferedfilter.py:53,
// buffereddata.py:149-153).  The source goes up once, the filtered trace never leaves the
// device between its producer and its consumers, every result starts on its way down as soon
// as its kernel has run, and the host waits once at the end.

static int32_t sosfilt_minmax_any_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                                      int64_t nbefore, double* dst, int64_t n_dst, const double* zi, double* zf,
                                      int64_t mm_step, double* mm_raw, double* mm_filt, cudaStream_t st);

static int32_t chain_check(const adn_chain_t* c, int64_t n_src, int32_t C, double rate, int64_t n_filt,
                           bool want_spec, bool want_env, bool want_mm) {
    if (!c || C < 1 || n_src < 1 || !(rate > 0) || c->S < 0 || c->S > ADN_MAX_SECTIONS || c->nbefore < 0 ||
        n_filt < 1 || n_filt > n_src - c->nbefore || (c->S > 0 && !c->sos))
        return fail(ADN_ERR_INVALID, "adn_chain_f64: filter stage n_src=%lld nbefore=%lld n_filt=%lld",
                    (long long)n_src, c ? (long long)c->nbefore : -1LL, (long long)n_filt);
    if (want_spec && (c->nfft < 1 || c->hop < 1 || c->hop > c->nfft || c->spec_first < 0 || c->spec_rows < 0 ||
                      c->spec_first + c->spec_rows > n_filt || c->n_spec < 0))
        return fail(ADN_ERR_INVALID, "adn_chain_f64: spectrogram stage");
    if (want_env && (c->ES < 1 || c->ES > ADN_MAX_SECTIONS || !c->esos || c->env_first < 0 || c->env_rows < 1 ||
                     c->env_first + c->env_rows > n_filt || c->env_nbefore < 0 || c->n_env < 0 ||
                     c->n_env > c->env_rows - c->env_nbefore))
        return fail(ADN_ERR_INVALID, "adn_chain_f64: envelope stage");
    if (want_env && c->env_rows <= adn_sosfiltfilt_edge(c->esos, c->ES))
        return fail(ADN_ERR_SHORT, "adn_chain_f64: the length of the envelope's input (%lld) must be greater "
                    "than the sosfiltfilt pad length %d", (long long)c->env_rows, adn_sosfiltfilt_edge(c->esos, c->ES));
    if (want_mm && c->mm_step < 1) return fail(ADN_ERR_INVALID, "adn_chain_f64: mm_step");
    return ADN_OK;
}

// the stages after the filter, on device buffers; spectrogram frames land in dspec (zero-filled
// beyond the computed ones), *nf = computed frames
static int32_t chain_consumers(const adn_chain_t* c, const double* dfilt, int32_t C, double rate,
                               double* dspec, double* denv, int64_t* nf, cudaStream_t st) {
    int32_t rc;
    if (dspec && c->n_spec > 0) {
        const size_t F = (size_t)c->nfft / 2 + 1;
        int64_t got = 0;
        if ((rc = spectrogram_dev(dfilt + c->spec_first * C, c->spec_rows, C, rate, c->nfft, c->hop,
                                  ADN_WINDOW_HANN, ADN_DETREND_CONSTANT, dspec, c->n_spec, c->out_db, &got, st)))
            return rc;
        if (got < c->n_spec)
            ADN_CK(cudaMemsetAsync(dspec + (size_t)got * C * F, 0, (size_t)(c->n_spec - got) * C * F * 8, st));
        if (nf) *nf = got;
    }
    if (denv && c->n_env > 0) {
        if ((rc = envelope_dev(c->esos, c->ES, dfilt + c->env_first * C, c->env_rows, C, c->env_nbefore, denv,
                               c->n_env, c->clamp_negative, st)))
            return rc;
    }
    return ADN_OK;
}

int32_t adn_chain_f64_dev(const adn_chain_t* c, const double* src, int64_t n_src, int32_t C, double rate,
                          double* filtered, int64_t n_filt, double* spec, double* env, double* minmax,
                          int64_t* n_computed, void* stream) {
    int32_t rc = chain_check(c, n_src, C, rate, n_filt, spec != nullptr, env != nullptr, minmax != nullptr);
    if (rc) return rc;
    if (!src || !filtered) return fail(ADN_ERR_INVALID, "adn_chain_f64_dev: NULL pointer");
    if ((rc = ensure_init())) return rc;
    cudaStream_t st = pick(stream);
    if (n_computed) *n_computed = 0;
    if (minmax) {
        // the raw rows' min/max in the filter's own pass over them where the kernel allows it
        if ((rc = sosfilt_minmax_any_dev(c->sos, c->S, src, n_src, C, c->nbefore, filtered, n_filt, nullptr, nullptr,
                                         c->mm_step, minmax, nullptr, st)))
            return rc;
    } else if ((rc = sosfilt_dev(c->sos, c->S, src, n_src, C, c->nbefore, filtered, n_filt, nullptr, nullptr, st))) {
        return rc;
    }
    return chain_consumers(c, filtered, C, rate, spec, env, n_computed, st);
}

int32_t adn_chain_f64(const adn_chain_t* c, const double* src, int64_t n_src, int32_t C, double rate,
                      double* filtered, int64_t n_filt, double* spec, double* env, double* minmax,
                      int64_t* n_computed, int64_t src_mirror, int64_t filt_mirror) {
    int32_t rc = chain_check(c, n_src, C, rate, n_filt, spec != nullptr, env != nullptr, minmax != nullptr);
    if (rc) return rc;
    if (!src || !filtered) return fail(ADN_ERR_INVALID, "adn_chain_f64: NULL pointer");
    ADN_API_LOCK;
    if ((rc = ensure_init())) return rc;
    if ((rc = ensure_streams())) return rc;
    Ctx& cx = ctx();
    if (n_computed) *n_computed = 0;
    const size_t in_b = (size_t)n_src * C * 8, filt_b = (size_t)n_filt * C * 8;
    const size_t F = (size_t)c->nfft / 2 + 1;
    const size_t spec_b = spec ? (size_t)c->n_spec * C * F * 8 : 0;
    const size_t env_b = env ? (size_t)c->n_env * C * 8 : 0;
    const int64_t nseg = minmax ? (n_src + c->mm_step - 1) / c->mm_step : 0;
    const size_t mm_b = (size_t)nseg * 2 * C * 8;
    // device buffers: filtered in its mirror (or staging), the other results in the chain's own
    double* dfilt = nullptr;
    if ((rc = out_buffer(filt_mirror, filtered, filt_b, &dfilt))) return rc;
    DevBuf& bs = scratch(SCR_CHAIN_SPEC, cx.stream);
    DevBuf& be = scratch(SCR_CHAIN_ENV, cx.stream);
    if ((rc = bs.reserve(spec_b + mm_b + 16))) return rc;
    if ((rc = be.reserve(env_b + 16))) return rc;
    double* dspec = spec ? bs.as<double>() : nullptr;
    double* dmm = minmax ? reinterpret_cast<double*>(static_cast<char*>(bs.p) + ((spec_b + 15) & ~(size_t)15)) : nullptr;
    double* denv = env ? be.as<double>() : nullptr;
    // ---- the source: its mirror, or uploaded in chunks while the filter follows behind
    const double* dsrc = nullptr;
    if ((rc = in_resident(src_mirror, src, in_b, &dsrc))) return rc;
    if (c->S == 0) {
        // reference: sos is None -> filtered = source[nbefore:]
        if (!dsrc) {
            if ((rc = cx.in.reserve(in_b))) return rc;
            ADN_CK(h2d(cx.in.p, src, in_b, cx.stream));
            dsrc = cx.in.as<double>();
        }
        ADN_CK(cudaMemcpyAsync(dfilt, dsrc + c->nbefore * C, filt_b, cudaMemcpyDeviceToDevice, cx.stream));
        if ((rc = chain(cx.stream, g_d2h))) return rc;
        ADN_CK(d2h(filtered, dfilt, filt_b, g_d2h));
    } else if (dsrc) {
        if ((rc = sosfilt_dev(c->sos, c->S, dsrc, n_src, C, c->nbefore, dfilt, n_filt, nullptr, nullptr, cx.stream)))
            return rc;
        if ((rc = chain(cx.stream, g_d2h))) return rc;
        ADN_CK(d2h(filtered, dfilt, filt_b, g_d2h));
    } else {
        if ((rc = cx.in.reserve(in_b))) return rc;
        double* din = cx.in.as<double>();
        const size_t z_b = (size_t)C * c->S * 2 * 8;
        if ((rc = cx.aux.reserve(3 * z_b + 4096))) return rc;
        double* d_state[2] = {cx.aux.as<double>() + (size_t)C * c->S * 2, cx.aux.as<double>() + (size_t)C * c->S * 4};
        int64_t R = g_opt[ADN_OPT_CHUNK_BYTES] / ((int64_t)C * 8);
        if (R < 4096) R = 4096;
        const double* zin = nullptr;
        // the first chunks are short (R/8, R/4, R/2, then R rows): the first results start their way
        // down after an eighth of a chunk's upload instead of a whole one
        int64_t Rk = R / 8 < 4096 ? 4096 : R / 8;
        int64_t a = 0;
        for (int64_t k = 0; a < n_src; ++k) {
            const int64_t b = a + Rk < n_src ? a + Rk : n_src;
            if (Rk < R) Rk = 2 * Rk < R ? 2 * Rk : R;
            ADN_CK(h2d(din + a * C, src + a * C, (size_t)(b - a) * C * 8, g_h2d));
            if ((rc = chain(g_h2d, cx.stream))) return rc;
            int64_t nb = c->nbefore - a;
            if (nb < 0) nb = 0;
            if (nb > b - a) nb = b - a;
            const int64_t o0 = a + nb - c->nbefore;
            int64_t no = b - a - nb;
            if (no > n_filt - o0) no = n_filt - o0;
            if (no < 0) no = 0;
            double* zout = b == n_src ? nullptr : d_state[k & 1];
            if (no > 0 || zout) {
                if ((rc = sosfilt_dev(c->sos, c->S, din + a * C, b - a, C, nb, no > 0 ? dfilt + o0 * C : nullptr, no,
                                      zin, zout, cx.stream)))
                    return rc;
            }
            zin = zout;
            if (no > 0) {
                if ((rc = chain(cx.stream, g_d2h))) return rc;
                ADN_CK(d2h(filtered + o0 * C, dfilt + o0 * C, (size_t)no * C * 8, g_d2h));
            }
            a = b;
        }
        dsrc = din;
    }
    // ---- the consumers read the filtered trace where it is
    int64_t nf = 0;
    if (dspec && c->n_spec > 0) {
        int64_t got = 0;
        if ((rc = spectrogram_dev(dfilt + c->spec_first * C, c->spec_rows, C, rate, c->nfft, c->hop,
                                  ADN_WINDOW_HANN, ADN_DETREND_CONSTANT, dspec, c->n_spec, c->out_db, &got, cx.stream)))
            return rc;
        nf = got;
        if ((rc = chain(cx.stream, g_d2h))) return rc;
        if (got > 0) ADN_CK(d2h(spec, dspec, (size_t)got * C * F * 8, g_d2h));
    }
    if (denv && c->n_env > 0) {
        if ((rc = envelope_dev(c->esos, c->ES, dfilt + c->env_first * C, c->env_rows, C, c->env_nbefore, denv,
                               c->n_env, c->clamp_negative, cx.stream)))
            return rc;
        if ((rc = chain(cx.stream, g_d2h))) return rc;
        ADN_CK(d2h(env, denv, env_b, g_d2h));
    }
    if (dmm) {
        if ((rc = minmax_dev(dsrc, n_src, C, c->mm_step, dmm, cx.stream))) return rc;
        if ((rc = chain(cx.stream, g_d2h))) return rc;
        ADN_CK(d2h(minmax, dmm, mm_b, g_d2h));
    }
    if ((rc = sync_pipeline())) return rc;
    if (spec && c->n_spec > nf) memset(spec + (size_t)nf * C * F, 0, (size_t)(c->n_spec - nf) * C * F * 8);
    mirror_commit(filt_mirror, filtered, filt_b, dfilt);
    if (n_computed) *n_computed = nf;
    return ADN_OK;
}

// ---------------------------------------------------------------- device entry points

int32_t adn_spec_image_db_f64_dev(const double* spec, int64_t n, int32_t C, int32_t F, int32_t channel,
                                  double* dst, void* stream) {
    if (n < 0 || C < 1 || F < 1 || channel < 0 || channel >= C || (n > 0 && (!spec || !dst)))
        return fail(ADN_ERR_INVALID, "adn_spec_image_db_f64_dev: bad arguments");
    if (n == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    return spec_image_dev(spec, n, C, F, channel, dst, pick(stream));
}

int32_t adn_mean_power_db_f64_dev(const double* spec, int32_t C, int32_t F, int32_t channel, int64_t i0,
                                  int64_t i1, double floor_db, double* dst, void* stream) {
    if (C < 1 || F < 1 || channel < 0 || channel >= C || i0 < 0 || i1 <= i0 || !spec || !dst)
        return fail(ADN_ERR_INVALID, "adn_mean_power_db_f64_dev: bad arguments");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return mean_power_dev(spec, C, F, channel, i0, i1, floor_db, dst, pick(stream));
}

int32_t adn_colsum_f64_dev(const double* spec, int64_t n, int64_t width, double* acc, void* stream) {
    if (n < 0 || width < 1 || (n > 0 && !spec) || !acc)
        return fail(ADN_ERR_INVALID, "adn_colsum_f64_dev: bad arguments");
    if (n == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    return colsum_dev(spec, n, width, acc, pick(stream));
}

int32_t adn_pcm_to_f64_dev(const void* pcm, int64_t n, int32_t bits, double gain, double* dst, void* stream) {
    if (n < 0 || (bits != 16 && bits != 24 && bits != 32) || (n > 0 && (!pcm || !dst)))
        return fail(ADN_ERR_INVALID, "adn_pcm_to_f64_dev: bad arguments");
    if (n == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    return pcm_dev(pcm, n, bits, gain, dst, pick(stream));
}


int32_t adn_unwrap_f64_dev(const double* src, int64_t n, int32_t C, double thresh, int32_t clips, double* dst,
                           void* stream) {
    if (n < 0 || C < 1 || (n > 0 && (!src || !dst)) || src == dst)
        return fail(ADN_ERR_INVALID, "adn_unwrap_f64_dev: bad arguments (out of place only)");
    if (n == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    if (!(thresh > 0.0)) {
        ADN_CK(cudaMemcpyAsync(dst, src, (size_t)n * C * 8, cudaMemcpyDeviceToDevice, pick(stream)));
        return ADN_OK;
    }
    return unwrap_dev(src, n, C, thresh, clips, dst, pick(stream));
}

int32_t adn_play_region_f64_dev(const double* src, int64_t n, int32_t C, const int32_t* left, int32_t nleft,
                                const int32_t* right, int32_t nright, double rate, double het_freq,
                                const double* sos, int32_t S, int64_t nstep, double* dst, void* stream) {
    int32_t rc = play_region_check(n, C, left, nleft, right, nright, rate, het_freq, sos, S, nstep);
    if (rc) return rc;
    if (n == 0) return ADN_OK;
    if (!src || !dst) return fail(ADN_ERR_INVALID, "adn_play_region_f64_dev: NULL pointer");
    if ((rc = ensure_init())) return rc;
    const int ncols = nright > 0 ? 2 : 1;
    if (!(het_freq > 0.0)) nstep = 1;
    DevBuf& wb = scratch(SCR_PLAY, pick(stream));
    if ((rc = wb.reserve((size_t)2 * n * ncols * 8 + 16))) return rc;
    return play_region_impl(src, n, C, left, nleft, right, nright, rate, het_freq, sos, S, nstep,
                            wb.as<double>(), dst, pick(stream));
}

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

// filter + the full-trace min/max rows of the raw source rows and / or of the filtered rows, in one
// pass over the data where the pipelined kernel applies (csrc/sosfwd.cu), else three launches
static int32_t sosfilt_minmax_any_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                                      int64_t nbefore, double* dst, int64_t n_dst, const double* zi, double* zf,
                                      int64_t mm_step, double* mm_raw, double* mm_filt, cudaStream_t st) {
    int32_t rc;
    if (S > 0 && option(ADN_OPT_SCAN_RUNS)) {
        bool handled = false;
        if ((rc = sosfilt_minmax_park_dev(sos, S, src, n_src, C, nbefore, dst, n_dst, zi, zf, mm_step, mm_raw,
                                          mm_filt, &handled, st)))
            return rc;
        if (handled) return ADN_OK;
    }
    if (mm_raw && (rc = minmax_dev(src, n_src, C, mm_step, mm_raw, st))) return rc;
    if ((rc = sosfilt_dev(sos, S, src, n_src, C, nbefore, dst, n_dst, zi, zf, st))) return rc;
    if (mm_filt && (rc = minmax_dev(dst, n_dst, C, mm_step, mm_filt, st))) return rc;
    return ADN_OK;
}

int32_t adn_sosfilt_minmax_f64_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                                   int64_t nbefore, double* dst, int64_t n_dst, const double* zi, double* zf,
                                   int64_t mm_step, double* mm_raw, double* mm_filt, void* stream) {
    if (S < 0 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 1 || nbefore < 0 || n_dst < 1 ||
        n_dst > n_src - nbefore || mm_step < 1 || (S > 0 && !sos) || !src || !dst)
        return fail(ADN_ERR_INVALID, "adn_sosfilt_minmax_f64_dev: bad arguments");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return sosfilt_minmax_any_dev(sos, S, src, n_src, C, nbefore, dst, n_dst, zi, zf, mm_step, mm_raw, mm_filt,
                                  pick(stream));
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

int32_t adn_sosfiltfilt_f64_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                                double* dst, int64_t n_dst, void* stream) {
    if (S < 1 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 1 || n_dst < 0 || n_dst > n_src || !sos || !src ||
        (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_sosfiltfilt_f64_dev: bad arguments");
    if (n_src <= adn_sosfiltfilt_edge(sos, S))
        return fail(ADN_ERR_SHORT, "adn_sosfiltfilt_f64_dev: input not longer than the pad");
    if (n_dst == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    return sosfiltfilt_dev(sos, S, src, n_src, C, 0, dst, n_dst, pick(stream));
}

int32_t adn_zero_phase_range_f64_dev(const double* sos, int32_t S, const double* src, int64_t n_src,
                                     int32_t C, int32_t rectify, int32_t edge_left, int32_t edge_right,
                                     int64_t first, double* dst, int64_t n_dst, int32_t clamp_negative,
                                     void* stream) {
    if (S < 1 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 1 || first < 0 || n_dst < 0 ||
        first + n_dst > n_src || !sos || !src || (n_dst > 0 && !dst))
        return fail(ADN_ERR_INVALID, "adn_zero_phase_range_f64_dev: bad arguments");
    if (n_src <= adn_sosfiltfilt_edge(sos, S))
        return fail(ADN_ERR_SHORT, "adn_zero_phase_range_f64_dev: input not longer than the sosfiltfilt pad");
    if (n_dst == 0) return ADN_OK;
    int32_t rc = ensure_init();
    if (rc) return rc;
    return zero_phase_range_dev(rectify != 0, sos, S, src, n_src, C, edge_left, edge_right, first, dst,
                                n_dst, clamp_negative, pick(stream));
}

int32_t adn_envelope_forward_f64_dev(const double* sos, int32_t S, const double* src, int64_t n_src,
                                     int32_t C, int32_t edge_left, int32_t edge_right,
                                     const double* zi, double* dst, double* zf, void* stream) {
    if (S < 1 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 1 || edge_left < 0 || edge_right < 0 ||
        edge_left >= n_src || edge_right >= n_src)
        return fail(ADN_ERR_INVALID, "adn_envelope_forward_f64_dev: bad shape");
    if (!sos || !src) return fail(ADN_ERR_INVALID, "adn_envelope_forward_f64_dev: NULL pointer");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return envelope_forward_dev(sos, S, src, n_src, C, edge_left, edge_right, zi, dst, zf, pick(stream));
}

int32_t adn_envelope_state0_f64_dev(const double* sos, int32_t S, const double* src, int32_t C,
                                    int32_t edge, int32_t which, double* out, void* stream) {
    if (S < 1 || S > ADN_MAX_SECTIONS || C < 1 || edge < 0 || which < 0 || which > 1 || !sos || !src || !out)
        return fail(ADN_ERR_INVALID, "adn_envelope_state0_f64_dev: bad arguments");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return envelope_state0_dev(sos, S, src, C, edge, which, out, pick(stream));
}

int32_t adn_fold_states_f64_dev(const double* packs, const double* mats, int32_t world, int32_t C,
                                int32_t D, int32_t rank, int32_t backward, double* out, void* stream) {
    if (world < 1 || C < 1 || D < 2 || D > 2 * ADN_MAX_SECTIONS || rank < 0 || rank >= world ||
        !packs || !mats || !out)
        return fail(ADN_ERR_INVALID, "adn_fold_states_f64_dev: bad arguments");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return fold_states_dev(packs, mats, world, C, D, rank, backward, out, pick(stream));
}

int32_t adn_sosfilt_reverse_f64_dev(const double* sos, int32_t S, const double* src, int64_t n_src,
                                    int32_t C, const double* zi, double* dst, int64_t first,
                                    int64_t n_dst, int32_t clamp_negative, double* zf, void* stream) {
    if (S < 1 || S > ADN_MAX_SECTIONS || C < 1 || n_src < 1 ||
        (dst && (first < 0 || n_dst < 0 || first + n_dst > n_src)))
        return fail(ADN_ERR_INVALID, "adn_sosfilt_reverse_f64_dev: bad shape");
    if (!sos || !src) return fail(ADN_ERR_INVALID, "adn_sosfilt_reverse_f64_dev: NULL pointer");
    int32_t rc = ensure_init();
    if (rc) return rc;
    return sosfilt_reverse_dev(sos, S, src, n_src, C, zi, dst, first, n_dst, clamp_negative, zf,
                               pick(stream));
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
