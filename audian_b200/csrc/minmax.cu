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
// independent 16-byte loads in flight per thread: 8 for full-size blocks and the warp kernel
// (measured on B200: config 4 73.8 -> 75.6 %, 64 channels 82 -> 85 %, one channel 71 -> 81 % of the
// HBM roofline), 4 where a block only has a few iterations (step 1920 x 8 channels: 65 against 63 %)
constexpr int MM_UNROLL = 8;

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
    // compare-and-select instead of fmin / fmax (a NaN fails both compares and is skipped just
    // the same; which of two equal zeros stays does not matter, resolve() overrides it): three
    // instructions per update where fmin / fmax cost eight on sm_100, which made the kernels
    // issue bound.  `any` (>= 0: the thread saw an element) is set once after the loops.
    __device__ __forceinline__ void feed(double v, int32_t row) {
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
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
template <int VEC, int UN>
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
        // main loop: UN independent loads in flight per thread
        for (; u + (int64_t)(UN - 1) * active < nunits; u += (int64_t)UN * active) {
            V v[UN];
#pragma unroll
            for (int k = 0; k < UN; ++k) v[k] = __ldcs(base + u + (int64_t)k * active);
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                double el[VEC];
                unpack(v[k], el);
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e].feed(el[e], r[e] + k * rpi);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) r[e] += UN * rpi;
        }
        for (; u < nunits; u += active) {
            double el[VEC];
            unpack(__ldcs(base + u), el);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { acc[e].feed(el[e], r[e]); r[e] += rpi; }
        }
        if (tid < nunits) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e].any = r[e] - rpi;        // the last row fed
        }
        // single channel read as pairs of rows: an odd row count leaves one element over
        if (VEC == 2 && (nflat & 1) && tid == (int)(nunits % active)) {
            acc[0].feed(src[row0 * C + nflat - 1], (int32_t)(nflat - 1));
            acc[0].any = (int32_t)(nflat - 1);
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e].resolve(mn[e], mx[e]);
    // Channel counts that tile a warp (slot period C/VEC lanes divides 32; C == 1 as pairs of
    // rows): butterfly over the lanes that hold the same channels, one row of shared memory per
    // warp, one barrier.  Other channel counts: tree over shared memory below.
    {
        const bool single = VEC == 2 && C == 1;
        const int lq = single ? 1 : C / VEC;
        if ((single || (C % VEC == 0 && lq >= 1 && lq <= 32 && (32 % lq) == 0)) && active == MM_THREADS) {
            if (single) { merge<true>(mn[0], mn[VEC - 1]); merge<false>(mx[0], mx[VEC - 1]); }
            const int nslot = single ? 1 : VEC;
            for (int o = lq; o < 32; o <<= 1) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    if (e < nslot) {
                        Best a, b;
                        a.v = __shfl_xor_sync(0xffffffffu, mn[e].v, o);
                        a.row = __shfl_xor_sync(0xffffffffu, mn[e].row, o);
                        b.v = __shfl_xor_sync(0xffffffffu, mx[e].v, o);
                        b.row = __shfl_xor_sync(0xffffffffu, mx[e].row, o);
                        merge<true>(mn[e], a);
                        merge<false>(mx[e], b);
                    }
                }
            }
            const int lane = tid & 31, warp = tid >> 5;
            if (lane < lq) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    if (e < nslot) {
                        const int ch = single ? 0 : lane * VEC + e;
                        s_mnv[warp * C + ch] = mn[e].v; s_mni[warp * C + ch] = mn[e].row;
                        s_mxv[warp * C + ch] = mx[e].v; s_mxi[warp * C + ch] = mx[e].row;
                    }
                }
            }
            __syncthreads();
            if (tid < C) {
                Best a{s_mnv[tid], s_mni[tid]}, b{s_mxv[tid], s_mxi[tid]};
                for (int w = 1; w < MM_THREADS / 32; ++w) {
                    Best c{s_mnv[w * C + tid], s_mni[w * C + tid]}, d{s_mxv[w * C + tid], s_mxi[w * C + tid]};
                    merge<true>(a, c);
                    merge<false>(b, d);
                }
                if (nsplit == 1) {
                    dst[(2 * seg) * C + tid] = canon(a.v, C);
                    dst[(2 * seg + 1) * C + tid] = canon(b.v, C);
                } else {
                    int64_t o = ((seg * nsplit + p) * 2) * C + tid;
                    part[o] = a.v;
                    part[o + C] = b.v;
                }
            }
            return;
        }
    }
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

// Short segments (up to 48 KB; longer ones are faster block-wise): one warp per segment, eight independent warps per
// block, no block barrier at all -- 16-byte loads of consecutive lanes are contiguous, four
// in flight per lane, and the lanes that hold the same channels are folded by shuffles.
// Needs a channel count that tiles a warp (32*VEC % C == 0), or C == 1 read as pairs of rows.
template <int VEC>
__global__ void __launch_bounds__(MM_THREADS)
minmax_warp_kernel(const double* __restrict__ src, int64_t n, int32_t C, int64_t step, int64_t nseg,
                   double* __restrict__ dst) {
    using V = typename VecT<VEC>::type;
    const int lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * (MM_THREADS / 32) + (threadIdx.x >> 5);
    if (seg >= nseg) return;
    const int64_t seg0 = seg * step;
    int64_t seg1 = seg0 + step;
    if (seg1 > n) seg1 = n;
    const int64_t nflat = (seg1 - seg0) * C;
    const int64_t nunits = nflat / VEC;
    const V* base = reinterpret_cast<const V*>(src + seg0 * C);
    const bool single = VEC == 2 && C == 1;
    const int32_t rpi = (32 * VEC) / C;                    // rows advanced per 32 units
    Fast acc[VEC];
    int32_t r[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { acc[e].reset(); r[e] = (lane * VEC + e) / C; }
    int64_t u = lane;
    for (; u + (MM_UNROLL - 1) * 32 < nunits; u += MM_UNROLL * 32) {
        V v[MM_UNROLL];
#pragma unroll
        for (int k = 0; k < MM_UNROLL; ++k) v[k] = __ldcs(base + u + k * 32);
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
    for (; u < nunits; u += 32) {
        double el[VEC];
        unpack(__ldcs(base + u), el);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { acc[e].feed(el[e], r[e]); r[e] += rpi; }
    }
    if (lane < nunits) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e].any = r[e] - rpi;            // the last row fed
    }
    if (VEC == 2 && (nflat & 1) && lane == (int)(nunits & 31)) {
        acc[0].feed(src[seg0 * C + nflat - 1], (int32_t)(nflat - 1));
        acc[0].any = (int32_t)(nflat - 1);
    }
    Best mn[VEC], mx[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e].resolve(mn[e], mx[e]);
    if (single) { merge<true>(mn[0], mn[VEC - 1]); merge<false>(mx[0], mx[VEC - 1]); }
    const int lq = single ? 1 : C / VEC;
    const int nslot = single ? 1 : VEC;
    for (int o = lq; o < 32; o <<= 1) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (e < nslot) {
                Best a, b;
                a.v = __shfl_xor_sync(0xffffffffu, mn[e].v, o);
                a.row = __shfl_xor_sync(0xffffffffu, mn[e].row, o);
                b.v = __shfl_xor_sync(0xffffffffu, mx[e].v, o);
                b.row = __shfl_xor_sync(0xffffffffu, mx[e].row, o);
                merge<true>(mn[e], a);
                merge<false>(mx[e], b);
            }
        }
    }
    if (lane < lq) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (e < nslot) {
                const int ch = single ? 0 : lane * VEC + e;
                dst[(2 * seg) * C + ch] = canon(mn[e].v, C);
                dst[(2 * seg + 1) * C + ch] = canon(mx[e].v, C);
            }
        }
    }
}

// One warp per (segment, channel) folds the partial results of the splits.  Splits are
// consecutive row ranges, so numpy's ordered rule is "the later split wins ties"; with the
// split index as the row of a (value, row) pair the merge is commutative and can run as a
// strided loop per lane followed by a shuffle reduction.
__global__ void __launch_bounds__(256)
minmax_combine_kernel(const double* __restrict__ part, int64_t nseg, int32_t C, int32_t nsplit,
                      int64_t n, int64_t step, int64_t rows_per_split, double* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nseg * C) return;
    const int64_t seg = i / C;
    const int32_t c = (int32_t)(i % C);
    int64_t seg0 = seg * step, seg1 = seg0 + step;
    if (seg1 > n) seg1 = n;
    const int32_t used = (int32_t)((seg1 - seg0 + rows_per_split - 1) / rows_per_split);   // non-empty splits
    const double* q = part + (seg * nsplit * 2) * C + c;
    Best mn{0.0, -1}, mx{0.0, -1};
    for (int32_t p = lane; p < used; p += 32) {
        Best a{q[((int64_t)p * 2) * C], p}, b{q[((int64_t)p * 2 + 1) * C], p};
        merge<true>(mn, a);
        merge<false>(mx, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best a, b;
        a.v = __shfl_xor_sync(0xffffffffu, mn.v, o);
        a.row = __shfl_xor_sync(0xffffffffu, mn.row, o);
        b.v = __shfl_xor_sync(0xffffffffu, mx.v, o);
        b.row = __shfl_xor_sync(0xffffffffu, mx.row, o);
        merge<true>(mn, a);
        merge<false>(mx, b);
    }
    if (lane == 0) {
        dst[(2 * seg) * C + c] = canon(mn.v, C);
        dst[(2 * seg + 1) * C + c] = canon(mx.v, C);
    }
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
    // 16-byte loads: rows of an even channel count, or a single channel read as pairs of rows
    // (every block then starts at an even row: even step, even rows per split)
    const bool vec2 = ((C % 2 == 0) || (C == 1 && step % 2 == 0)) &&
                      ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int VEC = vec2 ? 2 : 1;
    if (seg_elems < 2048 || C > MM_THREADS * VEC) {
        int64_t total = nseg * C;
        minmax_small_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, n, C, step, nseg, dst);
        count_launch();
        ADN_CK(cudaGetLastError());
        return ADN_OK;
    }
    {
        // short segments and enough of them: a warp per segment
        const bool single = vec2 && C == 1;
        const bool tiles = single || (C % VEC == 0 && (32 * VEC) % C == 0);
        const int64_t seg_bytes = seg_elems * 8;
        if (tiles && seg_bytes <= ((int64_t)48 << 10) && nseg >= (int64_t)ctx().sm_count * 8 &&
            seg_elems / C < ((int64_t)1 << 30)) {
            const int64_t blocks = (nseg + MM_THREADS / 32 - 1) / (MM_THREADS / 32);
            if (vec2)
                minmax_warp_kernel<2><<<(unsigned)blocks, MM_THREADS, 0, st>>>(src, n, C, step, nseg, dst);
            else
                minmax_warp_kernel<1><<<(unsigned)blocks, MM_THREADS, 0, st>>>(src, n, C, step, nseg, dst);
            count_launch();
            ADN_CK(cudaGetLastError());
            return ADN_OK;
        }
    }
    // threads whose per-iteration flat stride is a multiple of C
    int32_t q = C / (C % VEC == 0 ? VEC : 1);
    int32_t active = (MM_THREADS / q) * q;
    // rows per block: between ~4K and ~32K samples, chosen so that the blocks come in whole waves
    // of the resident slots (sm_count x 8 blocks of 256 threads): a last wave that fills a fraction
    // of the chip costs as much as a full one.  Measured on B200, 3.84 M rows x 8 ch in segments of
    // 90 000 rows: 938 blocks of 4096 rows 62.6 us, 1161 blocks of 3334 rows (one wave) see profiles/
    int64_t seg_rows = step < n ? step : n;
    const int64_t rows_hi = 32768 / C > 1 ? 32768 / C : 1;
    const int64_t rows_lo = 4096 / C + 1 < rows_hi ? 4096 / C + 1 : rows_hi;
    const int64_t slots = (int64_t)ctx().sm_count * (2048 / MM_THREADS);
    int64_t rows = rows_hi;
    {
        const int64_t k0 = (seg_rows + rows_hi - 1) / rows_hi, k1 = (seg_rows + rows_lo - 1) / rows_lo;
        const int64_t kstep = (k1 - k0) / 4096 + 1;
        double best = 0.0;
        for (int64_t k = k0; k <= k1; k += kstep) {
            int64_t r = (seg_rows + k - 1) / k;
            if (C == 1 && vec2 && (r & 1)) ++r;
            const int64_t nb = nseg * ((seg_rows + r - 1) / r);
            const double cost = (double)((nb + slots - 1) / slots) * ((double)r * C + 2048.0);
            if (k == k0 || cost < best) { best = cost; rows = r; }
        }
    }
    if (C == 1 && vec2 && rows > 1) rows &= ~(int64_t)1;
    if (rows > seg_rows) rows = seg_rows;
    if (rows < 1) rows = 1;
    if (rows >= (int64_t)1 << 30) rows = ((int64_t)1 << 30) - 1;          // row index is int32
    int64_t nsplit64 = (seg_rows + rows - 1) / rows;
    if (nsplit64 > 0x7fffffff || nseg * nsplit64 > 0x7fffffff)
        return fail(ADN_ERR_UNSUPPORTED, "adn_minmax: too many blocks (%lld segments x %lld splits)",
                    (long long)nseg, (long long)nsplit64);
    int32_t nsplit = (int32_t)nsplit64;
    double* part = nullptr;
    if (nsplit > 1) {
        DevBuf& sb = scratch(SCR_MINMAX_PART, st);
        int32_t rc = sb.reserve((size_t)nseg * nsplit * 2 * C * 8);
        if (rc) return rc;
        part = sb.as<double>();
    }
    unsigned grid = (unsigned)(nseg * nsplit);
    // deep unrolling pays when a thread has at least a few rounds of it
    const bool deep = rows * (int64_t)C >= (int64_t)4 * MM_THREADS * VEC * MM_UNROLL;
    if (vec2 && deep)
        minmax_split_kernel<2, MM_UNROLL><<<grid, MM_THREADS, 0, st>>>(src, n, C, step, rows, nsplit, active, dst, part);
    else if (vec2)
        minmax_split_kernel<2, 4><<<grid, MM_THREADS, 0, st>>>(src, n, C, step, rows, nsplit, active, dst, part);
    else if (deep)
        minmax_split_kernel<1, MM_UNROLL><<<grid, MM_THREADS, 0, st>>>(src, n, C, step, rows, nsplit, active, dst, part);
    else
        minmax_split_kernel<1, 4><<<grid, MM_THREADS, 0, st>>>(src, n, C, step, rows, nsplit, active, dst, part);
    count_launch();
    ADN_CK(cudaGetLastError());
    if (nsplit > 1) {
        int64_t total = nseg * C;
        minmax_combine_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(part, nseg, C, nsplit, n,
                                                                       step, rows, dst);
        count_launch();
        ADN_CK(cudaGetLastError());
    }
    return ADN_OK;
}

}  // namespace adn
