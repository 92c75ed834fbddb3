// Segment-wise min/max decimation of interleaved (n, C) float64 traces.
//
// Replaces the np.minimum/maximum.reduceat idiom of the reference
// (src/audian/compresseddata.py:49-52, :97-100 and src/audian/traceitem.py:58-61):
//   dst[2j, c] = min(src[j*step:(j+1)*step, c]),  dst[2j+1, c] = max(...)
// Bit-exact numpy semantics (probed on numpy 2.3.5 for the 2-D
// axis-0 form the reference uses): applied in time order the reduction is
//   acc = (isnan(v) || (!isnan(acc) && !(acc < v))) ? v : acc      (max: >)
// i.e. NaN propagates (the LAST NaN's payload survives) and among equal values
// (+0.0 / -0.0) the LATER row wins.  The parallel reduction keeps (value, row)
// pairs, which makes that rule commutative, so any combination order gives
// numpy's answer.  (For C == 1 numpy takes its 1-D SIMD path, which returns the
// canonical quiet NaN; signed-zero ties are lane-order dependent there.)
//
// Pure read-bandwidth kernel: 8 B per input sample, roofline = HBM.
#include "common.cuh"

namespace adn {

namespace {

constexpr int MM_THREADS = 256;
constexpr int MM_UNROLL = 4;

struct Best {          // running min or max with the row it came from (-1 = empty)
    double v;
    int32_t row;
};

// b is merged into a; `a` and `b` cover arbitrary (possibly interleaved) row sets
template <bool IS_MIN>
__device__ __forceinline__ void merge(Best& a, const Best& b) {
    if (b.row < 0) return;
    if (a.row < 0) { a = b; return; }
    bool a_nan = a.v != a.v, b_nan = b.v != b.v;
    bool take_b;
    if (a_nan || b_nan) {
        // a NaN beats any number; among NaNs the latest row wins
        take_b = b_nan && (!a_nan || b.row > a.row);
    } else if (a.v == b.v) {
        take_b = b.row > a.row;                     // ties: later row (signed zeros)
    } else {
        take_b = IS_MIN ? (b.v < a.v) : (b.v > a.v);
    }
    if (take_b) a = b;
}

// rows are fed to one accumulator in increasing order: numpy's predicate directly
__device__ __forceinline__ void feed(Best& mn, Best& mx, double v, int32_t row) {
    if (mn.row < 0) { mn.v = v; mn.row = row; mx.v = v; mx.row = row; return; }
    const bool vnan = v != v;
    if (vnan || (mn.v == mn.v && !(mn.v < v))) { mn.v = v; mn.row = row; }
    if (vnan || (mx.v == mx.v && !(mx.v > v))) { mx.v = v; mx.row = row; }
}

struct Fast {
    double mn, mx;         // over the non-NaN elements
    double special;        // value of the latest zero (row zrow) ...
    double nanv;           // ... and of the latest NaN (row nrow)
    int32_t zrow, nrow, any;
    __device__ __forceinline__ void reset() {
        mn = __longlong_as_double(0x7ff0000000000000ll);      // +inf
        mx = __longlong_as_double(0xfff0000000000000ll);      // -inf
        special = 0.0; nanv = 0.0; zrow = -1; nrow = -1; any = -1;
    }
    __device__ __forceinline__ void feed(double v, int32_t row) {
        mn = fmin(mn, v);
        mx = fmax(mx, v);
        any = row;
        if (!(fabs(v) > 0.0)) {                                // zero or NaN: rare
            if (v != v) { nanv = v; nrow = row; }
            else { special = v; zrow = row; }
        }
    }
    // numpy's result for the rows this thread saw, with the row that decides ties in the merge
    __device__ __forceinline__ void resolve(Best& bmn, Best& bmx) const {
        if (any < 0) { bmn.v = 0; bmn.row = -1; bmx.v = 0; bmx.row = -1; return; }
        if (nrow >= 0) { bmn.v = nanv; bmn.row = nrow; bmx.v = nanv; bmx.row = nrow; return; }
        bmn.v = mn; bmn.row = any;
        bmx.v = mx; bmx.row = any;
        if (mn == 0.0) { bmn.v = special; bmn.row = zrow; }
        if (mx == 0.0) { bmx.v = special; bmx.row = zrow; }
    }
};

// numpy's ordered update of a running min / max
__device__ __forceinline__ void upd_min(double& acc, double v) {
    if (v != v || (acc == acc && !(acc < v))) acc = v;
}
__device__ __forceinline__ void upd_max(double& acc, double v) {
    if (v != v || (acc == acc && !(acc > v))) acc = v;
}
__device__ __forceinline__ double canon(double v, int32_t C) {
    return (C == 1 && v != v) ? __longlong_as_double(0x7ff8000000000000ll) : v;
}

template <int VEC> struct VecT;
template <> struct VecT<1> { using type = double; };
template <> struct VecT<2> { using type = double2; };

__device__ __forceinline__ void unpack(double v, double* e) { e[0] = v; }
__device__ __forceinline__ void unpack(double2 v, double* e) { e[0] = v.x; e[1] = v.y; }

// One block reduces rows [row0, row1) of one segment (all channels) to one
// (min, max) pair per channel.  active = threads whose flat stride keeps their
// channel fixed: (active*VEC) % C == 0.
template <int VEC>
__global__ void __launch_bounds__(MM_THREADS)
minmax_split_kernel(const double* __restrict__ src, int64_t n, int32_t C, int64_t step,
                    int64_t rows_per_split, int32_t nsplit, int32_t active,
                    double* __restrict__ dst, double* __restrict__ part) {
    using V = typename VecT<VEC>::type;
    __shared__ double s_mnv[MM_THREADS * VEC], s_mxv[MM_THREADS * VEC];
    __shared__ int32_t s_mni[MM_THREADS * VEC], s_mxi[MM_THREADS * VEC];

    const int64_t seg = blockIdx.x / nsplit;
    const int32_t p = (int32_t)(blockIdx.x % nsplit);
    const int64_t seg0 = seg * step;
    int64_t seg1 = seg0 + step;
    if (seg1 > n) seg1 = n;
    const int64_t row0 = seg0 + (int64_t)p * rows_per_split;
    int64_t row1 = row0 + rows_per_split;
    if (row1 > seg1) row1 = seg1;
    const int tid = threadIdx.x;

    // Fast accumulation: plain fmin / fmax (they skip NaNs) plus, on the rare elements that are
    // a zero or a NaN, the row and value of the latest one.  numpy's ordered rule follows from
    // that: any NaN -> the latest NaN; a zero extremum -> the sign of the latest zero; every
    // other tie is between identical bit patterns.
    Fast acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e].reset();
    Best mn[VEC], mx[VEC];

    if (row0 < row1 && tid < active) {
        const int64_t nflat = (row1 - row0) * C;              // multiple of VEC by construction
        const int64_t nunits = nflat / VEC;
        const V* base = reinterpret_cast<const V*>(src + row0 * C);
        const int32_t rpi = (int32_t)(((int64_t)active * VEC) / C);   // rows advanced per iteration
        int32_t r[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) r[e] = (tid * VEC + e) / C;
        int64_t u = tid;
        // main loop: MM_UNROLL independent loads in flight per thread
        for (; u + (int64_t)(MM_UNROLL - 1) * active < nunits; u += (int64_t)MM_UNROLL * active) {
            V v[MM_UNROLL];
#pragma unroll
            for (int k = 0; k < MM_UNROLL; ++k) v[k] = __ldcs(base + u + (int64_t)k * active);
#pragma unroll
            for (int k = 0; k < MM_UNROLL; ++k) {
                double el[VEC];
                unpack(v[k], el);
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e].feed(el[e], r[e] + k * rpi);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) r[e] += MM_UNROLL * rpi;
        }
        for (; u < nunits; u += active) {
            double el[VEC];
            unpack(__ldcs(base + u), el);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { acc[e].feed(el[e], r[e]); r[e] += rpi; }
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e].resolve(mn[e], mx[e]);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        s_mnv[tid * VEC + e] = mn[e].v; s_mni[tid * VEC + e] = mn[e].row;
        s_mxv[tid * VEC + e] = mx[e].v; s_mxi[tid * VEC + e] = mx[e].row;
    }
    __syncthreads();
    // flat position f = tid*VEC + e belongs to channel f % C; K entries per channel
    const int32_t K = (active * VEC) / C;
    int32_t s = 1;
    while (s < K) s <<= 1;
    for (s >>= 1; s >= 1; s >>= 1) {
        for (int32_t f = tid; f < s * C; f += MM_THREADS) {
            int32_t k = f / C;
            if (k + s < K) {
                int32_t g = f + s * C;
                Best a{s_mnv[f], s_mni[f]}, b{s_mnv[g], s_mni[g]};
                merge<true>(a, b);
                s_mnv[f] = a.v; s_mni[f] = a.row;
                Best c{s_mxv[f], s_mxi[f]}, d{s_mxv[g], s_mxi[g]};
                merge<false>(c, d);
                s_mxv[f] = c.v; s_mxi[f] = c.row;
            }
        }
        __syncthreads();
    }
    for (int32_t c = tid; c < C; c += MM_THREADS) {
        if (nsplit == 1) {
            dst[(2 * seg) * C + c] = canon(s_mnv[c], C);
            dst[(2 * seg + 1) * C + c] = canon(s_mxv[c], C);
        } else {
            int64_t o = ((seg * nsplit + p) * 2) * C + c;
            part[o] = s_mnv[c];
            part[o + C] = s_mxv[c];
        }
    }
}

// splits of one segment are consecutive row ranges: numpy's predicate in order
__global__ void __launch_bounds__(256)
minmax_combine_kernel(const double* __restrict__ part, int64_t nseg, int32_t C, int32_t nsplit,
                      int64_t n, int64_t step, int64_t rows_per_split, double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg * C) return;
    int64_t seg = i / C;
    int32_t c = (int32_t)(i % C);
    int64_t seg0 = seg * step, seg1 = seg0 + step;
    if (seg1 > n) seg1 = n;
    int64_t used = (seg1 - seg0 + rows_per_split - 1) / rows_per_split;   // non-empty splits
    const double* q = part + (seg * nsplit * 2) * C + c;
    double mn = q[0], mx = q[C];
    for (int64_t p = 1; p < used; ++p) {
        upd_min(mn, q[(p * 2) * C]);
        upd_max(mx, q[(p * 2 + 1) * C]);
    }
    dst[(2 * seg) * C + c] = canon(mn, C);
    dst[(2 * seg + 1) * C + c] = canon(mx, C);
}

// short segments: one thread per (segment, channel), rows in order
__global__ void __launch_bounds__(256)
minmax_small_kernel(const double* __restrict__ src, int64_t n, int32_t C, int64_t step,
                    int64_t nseg, double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg * C) return;
    int64_t seg = i / C;
    int32_t c = (int32_t)(i % C);
    int64_t r0 = seg * step, r1 = r0 + step;
    if (r1 > n) r1 = n;
    const double* q = src + r0 * C + c;
    int64_t len = r1 - r0;
    double mn = q[0], mx = mn;
    int64_t r = 1;
    for (; r + 3 < len; r += 4) {
        double v0 = q[r * C], v1 = q[(r + 1) * C], v2 = q[(r + 2) * C], v3 = q[(r + 3) * C];
        upd_min(mn, v0); upd_max(mx, v0);
        upd_min(mn, v1); upd_max(mx, v1);
        upd_min(mn, v2); upd_max(mx, v2);
        upd_min(mn, v3); upd_max(mx, v3);
    }
    for (; r < len; ++r) {
        double v = q[r * C];
        upd_min(mn, v); upd_max(mx, v);
    }
    dst[(2 * seg) * C + c] = canon(mn, C);
    dst[(2 * seg + 1) * C + c] = canon(mx, C);
}

}  // namespace

int32_t minmax_dev(const double* src, int64_t n, int32_t C, int64_t step, double* dst,
                   cudaStream_t st) {
    const int64_t nseg = (n + step - 1) / step;
    const int64_t seg_elems = (step < n ? step : n) * (int64_t)C;
    const bool vec2 = (C % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int VEC = vec2 ? 2 : 1;
    if (seg_elems < 2048 || C > MM_THREADS * VEC) {
        int64_t total = nseg * C;
        minmax_small_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, n, C, step, nseg, dst);
        count_launch();
        ADN_CK(cudaGetLastError());
        return ADN_OK;
    }
    // threads whose per-iteration flat stride is a multiple of C
    int32_t q = C / (C % VEC == 0 ? VEC : 1);
    int32_t active = (MM_THREADS / q) * q;
    // rows per block: ~32K samples, but enough blocks to fill the chip a few times over
    int64_t rows = 32768 / C;
    if (rows < 1) rows = 1;
    int64_t want_blocks = (int64_t)ctx().sm_count * 8;
    int64_t seg_rows = step < n ? step : n;
    while (rows > 4096 / C + 1 && nseg * ((seg_rows + rows - 1) / rows) < want_blocks) rows /= 2;
    if (rows > seg_rows) rows = seg_rows;
    if (rows >= (int64_t)1 << 30) rows = ((int64_t)1 << 30) - 1;          // row index is int32
    int64_t nsplit64 = (seg_rows + rows - 1) / rows;
    if (nsplit64 > 0x7fffffff || nseg * nsplit64 > 0x7fffffff)
        return fail(ADN_ERR_UNSUPPORTED, "adn_minmax: too many blocks (%lld segments x %lld splits)",
                    (long long)nseg, (long long)nsplit64);
    int32_t nsplit = (int32_t)nsplit64;
    double* part = nullptr;
    if (nsplit > 1) {
        DevBuf& sb = scratch(SCR_MINMAX_PART);
        int32_t rc = sb.reserve((size_t)nseg * nsplit * 2 * C * 8);
        if (rc) return rc;
        part = sb.as<double>();
    }
    unsigned grid = (unsigned)(nseg * nsplit);
    if (vec2)
        minmax_split_kernel<2><<<grid, MM_THREADS, 0, st>>>(src, n, C, step, rows, nsplit, active, dst, part);
    else
        minmax_split_kernel<1><<<grid, MM_THREADS, 0, st>>>(src, n, C, step, rows, nsplit, active, dst, part);
    count_launch();
    ADN_CK(cudaGetLastError());
    if (nsplit > 1) {
        int64_t total = nseg * C;
        minmax_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part, nseg, C, nsplit, n,
                                                                           step, rows, dst);
        count_launch();
        ADN_CK(cudaGetLastError());
    }
    return ADN_OK;
}

}  // namespace adn
