// SOS biquad cascades over long interleaved traces as a single-pass chunked
// linear-recurrence scan: BufferedFilter.process (src/audian/bufferedfilter.py:31-36,
// scipy sosfilt) and BufferedEnvelope.process (src/audian/bufferedenvelope.py:34-41,
// scipy sosfiltfilt of (pi/2)|x|).
//
// The cascade of S direct-form-II-transposed biquads is the linear system
//     s[t+1] = A s[t] + B x[t],   y[t] = (last section's output)
// with D = 2S states per channel.  A tile = SOS_NT*SOS_L samples (time x channel
// group, contiguous rows of the interleaved input) is staged in shared memory
// (cp.async, 16-byte granules), each thread owns SOS_L consecutive samples of one
// channel:
//   pass A   zero-state end state of the sub-chunk: v = sum_i A^(L-1-i) B x_i  (dot
//            products against a host-made table in the kernel's constant bank)
//   scan     Kogge-Stone over the sub-chunks of a channel inside a warp with
//            precomputed A^(L 2^k), then a short serial scan over the warps
//   carry    decoupled look-back over preceding tiles (aggregate / inclusive
//            state records in global memory; a tile's index is blockIdx.x, which relies on
//            blocks being dispatched in index order, as CUB's look-back does);
//            contributions are weighted with (A^T)^j tables and the look-back stops
//            where the filter has decayed below 1e-30
//   pass B   the exact DF2T recurrence from the true incoming state, outputs
//            written in place into the staged tile, tile stored coalesced
// HBM traffic: 8 B read + 8 B written per sample; state carried in fp64.
#include "sos_common.cuh"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <atomic>

namespace adn {

namespace {

__device__ __forceinline__ double ld_relaxed(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(double* p, double v) {
    if (__double_as_longlong(v) == (long long)SOS_EMPTY) v = __longlong_as_double(0x7ff8000000000000ll);
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// all D words of a record, or false if any is still unpublished
template <int D>
__device__ __forceinline__ bool read_record(const double* p, double (&v)[D]) {
    bool ok = true;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        v[d] = ld_relaxed(p + d);
        ok = ok && (__double_as_longlong(v[d]) != (long long)SOS_EMPTY);
    }
    return ok;
}
// results go out with a streaming store, except the forward sweep of the envelope: the reversed
// sweep starts reading where that one stopped writing, so its tail is left to live in L2
// (measured on B200: envelope 208.6 against 210.6 us, order 4 262 against 266 us)
#ifndef ENVF_PLAIN_STORE
#define ENVF_PLAIN_STORE 1
#endif
template <int MODE, class T>
__device__ __forceinline__ void st_out(T* p, T v) {
    if (MODE == MODE_ENVF && ENVF_PLAIN_STORE) *p = v; else ADN_STORE(p, v);
}

// ---- write the finished tile back (clamp fused), coalesced
template <int MODE>
__device__ __forceinline__ void sos_store_tile(const SosRun& R, const double* tile_s, int64_t t0,
                                               int c0, int Cw, int tid) {
    const int CG = R.CG, C = R.C, T = R.T;
    const int pad = CG < 16 ? CG : 0;
    const int GS = SOS_L * Cw + pad;
    const int lc = Cw == CG ? R.lc : -1;
    // ---------------------------------------------------------------- store
    // fast path: contiguous tile, every row lands inside dst
    bool fast_out = lc >= 0 && t0 + T <= R.n;
    if (fast_out) {
        const int64_t lo = MODE == MODE_REV ? R.n - t0 - T : t0, hi = lo + T;   // physical rows
        fast_out = lo >= R.out_skip && hi <= R.out_skip + R.n_dst;
    }
    if (fast_out) {
        const int64_t rowbase = (MODE == MODE_REV ? R.n - 1 - t0 : t0) - R.out_skip;
        if (R.vec_out) {
            // the same constant strides as in sos_load_tile
            const int f0 = 2 * tid, r0 = f0 >> lc;
            const int DRk = (2 * SOS_NT) >> lc;
            const int64_t gstr = (MODE == MODE_REV ? -(int64_t)DRk : (int64_t)DRk) * C;
            double* gp = R.dst + (MODE == MODE_REV ? rowbase - r0 : rowbase + r0) * C + c0 + (f0 & (CG - 1));
            const double* sp = tile_s + f0 + (r0 >> 5) * pad;
            const int sstr = 2 * SOS_NT + (pad ? 8 : 0);
            if (R.clamp) {
#pragma unroll
                for (int k = 0; k < SOS_L / 2; ++k) {
                    double2 o = *reinterpret_cast<const double2*>(sp + k * sstr);
                    o.x = o.x < 0.0 ? 0.0 : o.x;
                    o.y = o.y < 0.0 ? 0.0 : o.y;
                    st_out<MODE>(reinterpret_cast<double2*>(gp), o);
                    gp += gstr;
                }
            } else {
#pragma unroll
                for (int k = 0; k < SOS_L / 2; ++k) {
                    st_out<MODE>(reinterpret_cast<double2*>(gp), *reinterpret_cast<const double2*>(sp + k * sstr));
                    gp += gstr;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < SOS_L; ++k) {
                const int f = tid + SOS_NT * k;
                const int r = f >> lc, c = c0 + (f & (CG - 1));
                const int64_t orow = MODE == MODE_REV ? rowbase - r : rowbase + r;
                double o = tile_s[f + (r >> 5) * pad];
                if (R.clamp) o = o < 0.0 ? 0.0 : o;
                st_out<MODE>(R.dst + orow * C + c, o);
            }
        }
        return;
    }
    {
        const int gw = R.vec_out ? 2 : 1;
        const int gpr = Cw / gw;
        const int total = T * gpr;
        int row = tid / gpr, col = tid - row * gpr;
        const int drow = SOS_NT / gpr, dcol = SOS_NT - drow * gpr;
        for (int q = tid; q < total; q += SOS_NT) {
            int64_t tau = t0 + row;
            int64_t phys = MODE == MODE_REV ? R.n - 1 - tau : tau;
            int64_t orow = phys - R.out_skip;
            if (tau < R.n && orow >= 0 && orow < R.n_dst) {
                const double* sp = tile_s + (row / SOS_L) * GS + (row % SOS_L) * Cw + col * gw;
                double* gp = R.dst + orow * C + c0 + col * gw;
                if (gw == 2) {
                    double2 o = *reinterpret_cast<const double2*>(sp);
                    if (R.clamp) { o.x = o.x < 0.0 ? 0.0 : o.x; o.y = o.y < 0.0 ? 0.0 : o.y; }
                    st_out<MODE>(reinterpret_cast<double2*>(gp), o);
                } else {
                    double o = *sp;
                    if (R.clamp) o = o < 0.0 ? 0.0 : o;
                    st_out<MODE>(gp, o);
                }
            }
            row += drow;
            col += dcol;
            if (col >= gpr) { col -= gpr; ++row; }
        }
    }
}

// ---- pass B of one thread: the exact DF2T recurrence over its SOS_L samples from state z,
// outputs written in place; z leaves as the state after the last sample
template <int S, int MODE, bool xform>
__device__ __forceinline__ void sos_recurrence_x(const SosK<S>& K, const SosRun& R, double* xp, int Cw,
                                                 double (&z)[2 * S], bool want_state,
                                                 int ilast_in, int chan) {
    constexpr int D = 2 * S;
        if (!want_state && S > 4) {
            const double* xr = xp;
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) {
                double x = *xr;
                if (MODE == MODE_ENVF && xform) x = pre_x<MODE>(x);
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    double y = fma(K.coef[s][0], x, z[2 * s]);
                    z[2 * s] = fma(K.coef[s][1], x, z[2 * s + 1]) - K.coef[s][3] * y;
                    z[2 * s + 1] = K.coef[s][2] * x - K.coef[s][4] * y;
                    x = y;
                }
                *const_cast<double*>(xr) = x;
                xr += Cw;
            }
        } else if (!want_state) {
            // the sections run skewed by one sample each (section s works on sample j - s in
            // step j): the S recurrences of a step are independent of each other, which hides
            // the latency of the dependent fp64 operations; the arithmetic per sample is unchanged
            double xin[S + 1];
#pragma unroll
            for (int j = 0; j < SOS_L + S - 1; ++j) {
#pragma unroll
                for (int s = S - 1; s >= 0; --s) {
                    const int i = j - s;
                    if (i < 0 || i >= SOS_L) continue;
                    double x;
                    if (s == 0) {
                        x = xp[i * Cw];
                        if (MODE == MODE_ENVF && xform) x = pre_x<MODE>(x);
                    } else {
                        x = xin[s];
                    }
                    double y = fma(K.coef[s][0], x, z[2 * s]);
                    z[2 * s] = fma(K.coef[s][1], x, z[2 * s + 1]) - K.coef[s][3] * y;
                    z[2 * s + 1] = K.coef[s][2] * x - K.coef[s][4] * y;
                    if (s == S - 1) xp[i * Cw] = y; else xin[s + 1] = y;
                }
            }
        } else {
            // the one sub-chunk per channel that contains the last sample: plain loop, state
            // captured right after that sample
            const int ilast = ilast_in;
            for (int i = 0; i < SOS_L; ++i) {
                double x = xp[i * Cw];
                if (MODE == MODE_ENVF && xform) x = pre_x<MODE>(x);
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    double y = fma(K.coef[s][0], x, z[2 * s]);
                    z[2 * s] = fma(K.coef[s][1], x, z[2 * s + 1]) - K.coef[s][3] * y;
                    z[2 * s + 1] = K.coef[s][2] * x - K.coef[s][4] * y;
                    x = y;
                }
                xp[i * Cw] = x;
                if (i == ilast) {
#pragma unroll
                    for (int d = 0; d < D; ++d) R.zf[(size_t)chan * D + d] = z[d];
                }
            }
        }
}

// the rectification flag of a tile is block-uniform: one branch per tile instead of a select per
// sample (only the two edge tiles of an envelope sweep take the second copy)
template <int S, int MODE>
__device__ __forceinline__ void sos_recurrence(const SosK<S>& K, const SosRun& R, double* xp, int Cw,
                                               double (&z)[2 * S], bool xform, bool want_state,
                                               int ilast_in, int chan) {
    if (MODE == MODE_ENVF && !xform) sos_recurrence_x<S, MODE, false>(K, R, xp, Cw, z, want_state, ilast_in, chan);
    else sos_recurrence_x<S, MODE, true>(K, R, xp, Cw, z, want_state, ilast_in, chan);
}

// pass A of one thread: zero-state end state of its SOS_L samples (dot products against K.W)
template <int S, int MODE, bool xform>
__device__ __forceinline__ void sos_pass_a_x(const SosK<S>& K, const double* xp, int Cw, double (&v)[2 * S]) {
    constexpr int D = 2 * S;
#pragma unroll
    for (int i = 0; i < SOS_L; ++i) {
        double x = xp[i * Cw];
        if (MODE == MODE_ENVF && xform) x = pre_x<MODE>(x);
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = fma(K.W[d][i], x, v[d]);
    }
}
template <int S, int MODE>
__device__ __forceinline__ void sos_pass_a(const SosK<S>& K, const double* xp, int Cw, double (&v)[2 * S],
                                           bool xform) {
    if (MODE == MODE_ENVF && !xform) sos_pass_a_x<S, MODE, false>(K, xp, Cw, v);
    else sos_pass_a_x<S, MODE, true>(K, xp, Cw, v);
}

template <int S, int MODE>
__global__ void __launch_bounds__(SOS_NT, S <= 2 ? 6 : (S <= 4 ? 4 : 1))
sos_scan_kernel(const __grid_constant__ SosK<S> K, const __grid_constant__ SosRun R) {
    constexpr int D = 2 * S;
    constexpr int DD = D * D;
    constexpr bool STAGE = S <= 4;           // small tables live in shared memory
    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tiles only wait on lower-numbered tiles; like CUB's decoupled look-back this relies on
    // the hardware dispatching the blocks of a 1-D grid in index order
    const int64_t tile = blockIdx.x;
    const int64_t tt = tile / R.ngroups;
    const int grp = (int)(tile % R.ngroups);
    const int CG = R.CG, C = R.C;
    const int c0 = grp * CG;
    const int Cw = min(CG, C - c0);
    const int GW = 32 / CG;
    const int gl = lane / CG, cw = lane % CG;
    const int g = warp * GW + gl;
    const bool chan_ok = cw < Cw;
    const int T = R.T;
    const int64_t t0 = tt * T;
    const int pad = CG < 16 ? CG : 0;
    const int GS = SOS_L * Cw + pad;
    const int G = SOS_NT / CG;

    double* tile_s = smem;                                   // G * (L*CG + pad)
    double* wagg = smem + (size_t)G * (SOS_L * CG + pad);    // [NW][CG][D]
    double* wtile = wagg + SOS_NW * CG * D;                  // [NW][CG][D]
    double* sin_s = wtile + SOS_NW * CG * D;                 // [CG][D]
    double* tab_s = sin_s + CG * D;                          // n_staged * DD (if STAGE)
    const double* tab = STAGE ? tab_s : R.tab;               // scan / fix / wpow tables
    const double* tab_fix = tab + R.off_fix * DD;
    const double* tab_wpow = tab + R.off_wpow * DD;

    // ---------------------------------------------------------------- load
    // warm L2 for the tile a later block will load: one bulk prefetch (TMA unit) of the
    // contiguous rows `R.pf_tiles` time tiles ahead
    if (tid == 0 && grp == 0 && R.pf_tiles > 0) {
        const int64_t pt0 = t0 + (int64_t)R.pf_tiles * T;
        const int64_t lim = ADN_EXT(MODE) ? R.nx : R.n;
        int64_t p0 = MODE == MODE_REV ? R.n - pt0 - T : (ADN_EXT(MODE) ? pt0 - R.edge : pt0);
        int64_t p1 = p0 + T;
        if (p0 < 0) p0 = 0;
        if (p1 > lim) p1 = lim;
        if (p1 > p0) {
            const char* a = reinterpret_cast<const char*>(R.src + p0 * C);
            int64_t bytes = (p1 - p0) * (int64_t)C * 8;
            int64_t mis = reinterpret_cast<uintptr_t>(a) & 15;
            a -= mis;
            bytes = (bytes + mis + 15) & ~(int64_t)15;
            if (a >= reinterpret_cast<const char*>(R.src))
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((uint32_t)bytes) : "memory");
        }
    }
    if (STAGE) {
        const int ngran = R.n_staged * DD / 2;
        for (int q = tid; q < ngran; q += SOS_NT) cp_async16(tab_s + 2 * q, R.tab + 2 * q, 16);
    }
    const bool xform = sos_load_tile<MODE>(R, tile_s, t0, c0, Cw, tid);
    cp_async_wait_all();
    __syncthreads();

    // ---------------------------------------------------------------- pass A
    double* xp = tile_s + g * GS + cw;
    double v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = 0.0;
    if (chan_ok) {
        sos_pass_a<S, MODE>(K, xp, Cw, v, xform);
    }

    // ---------------------------------------------------------------- warp scan over gl
    {
        int k = 0;
        for (int off = CG; off < 32; off <<= 1, ++k) {
            double u[D];
#pragma unroll
            for (int d = 0; d < D; ++d) u[d] = __shfl_up_sync(0xffffffffu, v[d], off);
            if (lane >= off) matvec_acc<D>(tab + k * DD, u, v);
        }
    }
    double ex[D];                         // exclusive prefix inside the warp
#pragma unroll
    for (int d = 0; d < D; ++d) {
        double t = __shfl_up_sync(0xffffffffu, v[d], CG & 31);
        ex[d] = gl == 0 ? 0.0 : t;
    }
    if (gl == GW - 1) {
#pragma unroll
        for (int d = 0; d < D; ++d) wagg[(warp * CG + cw) * D + d] = v[d];
        // this warp's share of the tile aggregate, A^(L GW (NW-1-warp)) v: computed here, in
        // parallel over the warps, so that the look-back lanes only have to add four vectors
        double wv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) wv[d] = 0.0;
        matvec_acc<D>(tab_wpow + (SOS_NW - 1 - warp) * DD, v, wv);
#pragma unroll
        for (int d = 0; d < D; ++d) wtile[(warp * CG + cw) * D + d] = wv[d];
    }
    __syncthreads();

    // ---------------------------------------------------------------- tile level
    // every thread: state entering its warp's chunk if the tile started from zero,
    //   pre = sum_{j<warp} A^(L GW (warp-1-j)) wagg[j]
    double pre[D];
#pragma unroll
    for (int d = 0; d < D; ++d) pre[d] = 0.0;
    for (int j = 0; j < warp; ++j) {
        double a[D];
#pragma unroll
        for (int d = 0; d < D; ++d) a[d] = wagg[(j * CG + cw) * D + d];
        matvec_acc<D>(tab_wpow + (warp - 1 - j) * DD, a, pre);
    }
    // warp 0: aggregate of the tile and look-back for the incoming state.  Lane (gl, cw) handles
    // predecessor j = base + gl + 1 of channel cw, so that the records of GW predecessors are
    // fetched in ONE round trip to L2 instead of one per predecessor; the weighted contributions
    // are then folded over gl by shuffles.
    if (warp == 0) {
        const double* Pt = R.tab + (size_t)R.off_tile * DD;
        const bool ch_real = c0 + cw < C;
        double acc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] = 0.0;
#pragma unroll
        for (int j = 0; j < SOS_NW; ++j) {
#pragma unroll
            for (int d = 0; d < D; ++d) acc[d] += wtile[(j * CG + cw) * D + d];
        }
        const bool publish = tt + 1 < R.ntt;
        const size_t rec = (size_t)tile * CG * D + (size_t)cw * D;
        if (publish && gl == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) st_relaxed(R.agg + rec + d, acc[d]);
        }
        // ---- incoming state of the tile: sum_j (A^T)^(j-1) aggregate(tile-j), closed by the
        // first inclusive state found, by the initial state, or where the weights have decayed
        double sin[D];
#pragma unroll
        for (int d = 0; d < D; ++d) sin[d] = 0.0;
        unsigned chmask = 0;                       // the lanes of this lane's channel
        for (int k = 0; k < GW; ++k) chmask |= 1u << (cw + CG * k);
        const unsigned mybit = 1u << lane;
        const int jmax = min(R.jdecay, SOS_LOOK + 1);
        bool ch_closed = false;
        for (int base = 0; base < jmax; base += GW) {
            const int j = base + gl + 1;
            const int64_t b = tt - j;
            // 0: record outstanding, 1: aggregate taken (or nothing to add), 2: closes the sum
            int state = 0;
            double vec[D];
#pragma unroll
            for (int d = 0; d < D; ++d) vec[d] = 0.0;
            if (j > jmax || ch_closed || b < -1) {
                state = 1;
            } else if (b < 0) {
#pragma unroll
                for (int d = 0; d < D; ++d)
                    vec[d] = (ch_real && R.s0) ? __ldg(R.s0 + (size_t)(c0 + cw) * D + d) : 0.0;
                state = 2;
            }
            const size_t prec = (size_t)(tile - (int64_t)j * R.ngroups) * CG * D + (size_t)cw * D;
            unsigned needed = chmask;
            unsigned ns = 0;
            while (true) {
                if (state == 0 && (needed & mybit)) {
                    // both records in one round trip: the inclusive state closes the look-back,
                    // else the aggregate is taken (the last slot of the window must be inclusive)
                    double vi[D], va[D];
                    const bool ok_i = read_record<D>(R.incl + prec, vi);
                    const bool ok_a = read_record<D>(R.agg + prec, va);
                    if (ok_i) {
                        state = 2;
#pragma unroll
                        for (int d = 0; d < D; ++d) vec[d] = vi[d];
                    } else if (j <= SOS_LOOK && ok_a) {
                        state = 1;
#pragma unroll
                        for (int d = 0; d < D; ++d) vec[d] = va[d];
                    }
                }
                const unsigned pend = __ballot_sync(0xffffffffu, state == 0) & chmask;
                const unsigned clos = __ballot_sync(0xffffffffu, state == 2) & chmask;
                // predecessors up to the first closing one are needed, the others are not
                const int fc = clos ? __ffs(clos) - 1 : 31;
                needed = chmask & (fc >= 31 ? 0xffffffffu : ((2u << fc) - 1u));
                const bool done = (pend & needed) == 0;
                if (__all_sync(0xffffffffu, done)) {
                    ch_closed = ch_closed || clos != 0;
                    break;
                }
                if (ns) __nanosleep(ns);
                ns = ns ? (ns < 256 ? ns * 2 : ns) : 32;
            }
            double w[D];
#pragma unroll
            for (int d = 0; d < D; ++d) w[d] = 0.0;
            if ((needed & mybit) && state != 0 && j <= jmax)
                matvec_acc<D>(Pt + (size_t)(j - 1) * DD, vec, w);
            for (int off = CG; off < 32; off <<= 1) {
#pragma unroll
                for (int d = 0; d < D; ++d) w[d] += __shfl_xor_sync(0xffffffffu, w[d], off);
            }
#pragma unroll
            for (int d = 0; d < D; ++d) sin[d] += w[d];
            if (__all_sync(0xffffffffu, ch_closed)) break;
        }
        if (gl == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) sin_s[cw * D + d] = sin[d];
            // ---- inclusive state of this tile, for the successors
            if (publish) {
                double inc[D];
#pragma unroll
                for (int d = 0; d < D; ++d) inc[d] = acc[d];
                matvec_acc<D>(Pt + DD, sin, inc);
#pragma unroll
                for (int d = 0; d < D; ++d) st_relaxed(R.incl + rec + d, inc[d]);
            }
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- pass B
    const int64_t tau0 = t0 + (int64_t)g * SOS_L;
    const int64_t last = R.n - 1;
    const bool want_state = R.zf != nullptr && last >= tau0 && last < tau0 + SOS_L;
    if (chan_ok && (R.dst != nullptr || want_state)) {
        double z[D];
        {
            // state entering the warp chunk: pre + A^(L GW warp) sin
            double sv[D];
#pragma unroll
            for (int d = 0; d < D; ++d) sv[d] = sin_s[cw * D + d];
            matvec_acc<D>(tab_wpow + warp * DD, sv, pre);
#pragma unroll
            for (int d = 0; d < D; ++d) z[d] = ex[d];
            if (gl == 0) {
#pragma unroll
                for (int d = 0; d < D; ++d) z[d] += pre[d];
            } else {
                matvec_acc<D>(tab_fix + gl * DD, pre, z);
            }
        }
        sos_recurrence<S, MODE>(K, R, xp, Cw, z, xform, want_state, (int)(last - tau0), c0 + cw);
    }
    if (R.dst == nullptr) return;
    __syncthreads();

    sos_store_tile<MODE>(R, tile_s, t0, c0, Cw, tid);
}

// ======================================================================================
// Run variant of the scan for cascades that forget quickly (the usual case: every cut-off
// audian offers at audio rates decays below 1e-20 within one or two tiles).  A block owns a
// RUN of consecutive time tiles of one channel group and walks along time, the state handed
// from tile to tile through shared memory: no tile records, no look-back, no waiting on
// other blocks.  A run that does not start at row 0 first runs over the `pre_tiles` tiles
// before it from zero state without storing anything: by then the unknown state that entered
// those tiles has decayed below 1e-20 of its size, far under the rounding of the state
// itself (the look-back kernel truncates in the same way at 1e-30).  The tiles of a run are
// prefetched nbuf - 1 ahead with cp.async into a ring of tile buffers, so the loads of the
// next tile overlap both passes over the current one.  Passes A and B and the scans inside
// the tile are those of sos_scan_kernel.
constexpr int SOS_RUN_BLOCKS = 3;

template <int S, int MODE>
__global__ void __launch_bounds__(SOS_NT, SOS_RUN_BLOCKS)
sos_run_kernel(const __grid_constant__ SosK<S> K, const __grid_constant__ SosRun R) {
    constexpr int D = 2 * S;
    constexpr int DD = D * D;
    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = (int)(blockIdx.x % R.ngroups);
    const int64_t run = blockIdx.x / R.ngroups;
    const int CG = R.CG, C = R.C;
    const int c0 = grp * CG;
    const int Cw = min(CG, C - c0);
    const int GW = 32 / CG;
    const int gl = lane / CG, cw = lane % CG;
    const int g = warp * GW + gl;
    const bool chan_ok = cw < Cw;
    const int T = R.T;
    const int pad = CG < 16 ? CG : 0;
    const int GS = SOS_L * Cw + pad;
    const int G = SOS_NT / CG;
    const int nbuf = R.nbuf;
    const size_t TS = (size_t)G * (SOS_L * CG + pad);        // doubles per tile buffer

    double* bufs = smem;                                     // nbuf * TS
    double* wagg = bufs + (size_t)nbuf * TS;                 // [NW][CG][D]
    double* sin_s = wagg + SOS_NW * CG * D;                  // [2][CG][D]
    double* tab_s = sin_s + 2 * CG * D;                      // n_staged * DD
    const double* tab = tab_s;
    const double* tab_fix = tab + R.off_fix * DD;
    const double* tab_wpow = tab + R.off_wpow * DD;

    const int64_t tt_first = run * R.run_tiles;
    const int64_t tt_last = min(tt_first + (int64_t)R.run_tiles, R.ntt);
    const int64_t tt_start = max((int64_t)0, tt_first - R.pre_tiles);
    if (tt_first >= tt_last) return;

    // ---- prologue: tables, the first nbuf - 1 tiles, the state entering the run
    {
        const int ngran = R.n_staged * DD / 2;
        for (int q = tid; q < ngran; q += SOS_NT) cp_async16(tab_s + 2 * q, R.tab + 2 * q, 16);
    }
    for (int k = 0; k < nbuf - 1; ++k) {
        if (tt_start + k < tt_last) sos_load_tile<MODE>(R, bufs + (size_t)k * TS, (tt_start + k) * T, c0, Cw, tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (tid < CG * D) {
        const int ch = tid / D;
        double v0 = 0.0;
        if (tt_start == 0 && R.s0 && c0 + ch < C) v0 = __ldg(R.s0 + (size_t)c0 * D + tid);
        sin_s[tid] = v0;
    }

    const int64_t last = R.n - 1;
    int it = 0;
    for (int64_t tt = tt_start; tt < tt_last; ++tt, ++it) {
        double* tile_s = bufs + (size_t)(it % nbuf) * TS;
        const int64_t t0 = tt * T;
        const bool xform = ADN_EXT(MODE) && t0 >= R.edge && t0 + T <= R.edge + R.nx;
        // the current tile has landed when at most nbuf - 2 younger groups are outstanding
        if (nbuf == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        // the tile nbuf - 1 ahead goes into the buffer the previous iteration stored from (every
        // thread has crossed the barrier above, which it reaches after that store); one group
        // per iteration, empty at the end of the run
        if (tt + nbuf - 1 < tt_last)
            sos_load_tile<MODE>(R, bufs + (size_t)((it + nbuf - 1) % nbuf) * TS, (tt + nbuf - 1) * T, c0, Cw, tid);
        asm volatile("cp.async.commit_group;" ::: "memory");

        // ------------------------------------------------------------ pass A
        double* xp = tile_s + g * GS + cw;
        double v[D];
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = 0.0;
        if (chan_ok) {
            // full groups of 8 channels (the usual case): the row stride is a literal, so the
            // shared-memory offsets of the unrolled passes are immediates
            if (Cw == 8) sos_pass_a<S, MODE>(K, xp, 8, v, xform);
            else sos_pass_a<S, MODE>(K, xp, Cw, v, xform);
        }
        // ------------------------------------------------------------ warp scan over gl
        {
            int k = 0;
            for (int off = CG; off < 32; off <<= 1, ++k) {
                double u[D];
#pragma unroll
                for (int d = 0; d < D; ++d) u[d] = __shfl_up_sync(0xffffffffu, v[d], off);
                if (lane >= off) matvec_acc<D>(tab + k * DD, u, v);
            }
        }
        double ex[D];                         // exclusive prefix inside the warp
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double t = __shfl_up_sync(0xffffffffu, v[d], CG & 31);
            ex[d] = gl == 0 ? 0.0 : t;
        }
        if (gl == GW - 1) {
#pragma unroll
            for (int d = 0; d < D; ++d) wagg[(warp * CG + cw) * D + d] = v[d];
        }
        __syncthreads();

        // ------------------------------------------------------------ state entering the thread
        double pre[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pre[d] = 0.0;
        for (int j = 0; j < warp; ++j) {
            double a[D];
#pragma unroll
            for (int d = 0; d < D; ++d) a[d] = wagg[(j * CG + cw) * D + d];
            matvec_acc<D>(tab_wpow + (warp - 1 - j) * DD, a, pre);
        }
        const int64_t tau0 = t0 + (int64_t)g * SOS_L;
        const bool own = tt >= tt_first;
        const bool want_state = own && R.zf != nullptr && last >= tau0 && last < tau0 + SOS_L;
        if (chan_ok) {
            double z[D];
            {
                double sv[D];
#pragma unroll
                for (int d = 0; d < D; ++d) sv[d] = sin_s[(it & 1) * CG * D + cw * D + d];
                matvec_acc<D>(tab_wpow + warp * DD, sv, pre);
#pragma unroll
                for (int d = 0; d < D; ++d) z[d] = ex[d];
                if (gl == 0) {
#pragma unroll
                    for (int d = 0; d < D; ++d) z[d] += pre[d];
                } else {
                    matvec_acc<D>(tab_fix + gl * DD, pre, z);
                }
            }
            // -------------------------------------------------------- pass B
            if (Cw == 8) sos_recurrence<S, MODE>(K, R, xp, 8, z, xform, want_state, (int)(last - tau0), c0 + cw);
            else sos_recurrence<S, MODE>(K, R, xp, Cw, z, xform, want_state, (int)(last - tau0), c0 + cw);
            if (g == G - 1) {                 // state after the tile's last row: enters the next tile
#pragma unroll
                for (int d = 0; d < D; ++d) sin_s[((it + 1) & 1) * CG * D + cw * D + d] = z[d];
            }
        }
        __syncthreads();
        if (own && R.dst != nullptr) sos_store_tile<MODE>(R, tile_s, t0, c0, Cw, tid);
    }
}

// initial states of the two sosfiltfilt sweeps: s0[c][d] = zi[d] * x0[c]
// which = 0: x0 = ext[0] = 2 r(0) - r(edge) of the raw input;  which = 1: x0 = row[c]

__global__ void env_s0_kernel(int which, const double* __restrict__ src, int32_t C, int32_t D,
                              int64_t edge, const __grid_constant__ ZiK zik, double* __restrict__ s0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * D) return;
    int c = i / D, d = i - c * D;
    const double* zi = zik.z;
    double x0;
    if (which == 0) {
        const double hp = 1.5707963267948966;
        x0 = 2.0 * (hp * fabs(src[c])) - hp * fabs(src[edge * C + c]);
    } else if (which == 2) {
        x0 = 2.0 * src[c] - src[edge * C + c];       // ext[0] of the plain odd extension
    } else {
        x0 = src[c];
    }
    s0[i] = zi[d] * x0;
}

// ------------------------------------------------------------------ host side plan

typedef long double ld;

struct Mat {                      // D x D, row major
    int D;
    std::vector<ld> a;
    explicit Mat(int D_) : D(D_), a((size_t)D_ * D_, 0.0L) {}
    ld& at(int r, int c) { return a[(size_t)r * D + c]; }
    ld at(int r, int c) const { return a[(size_t)r * D + c]; }
    static Mat eye(int D) { Mat m(D); for (int i = 0; i < D; ++i) m.at(i, i) = 1.0L; return m; }
};

Mat mul(const Mat& x, const Mat& y) {
    Mat r(x.D);
    for (int i = 0; i < x.D; ++i)
        for (int k = 0; k < x.D; ++k) {
            ld xv = x.at(i, k);
            if (xv == 0.0L) continue;
            for (int j = 0; j < x.D; ++j) r.at(i, j) += xv * y.at(k, j);
        }
    return r;
}

Mat mpow(Mat base, int64_t e) {
    Mat r = Mat::eye(base.D);
    while (e > 0) {
        if (e & 1) r = mul(r, base);
        e >>= 1;
        if (e) base = mul(base, base);
    }
    return r;
}

// one step of the cascade (the recurrence of scipy's _sosfilt) in long double
void step(const double* sos, int S, std::vector<ld>& z, ld x) {
    for (int s = 0; s < S; ++s) {
        const double* q = sos + 6 * s;
        ld y = (ld)q[0] * x + z[2 * s];
        z[2 * s] = (ld)q[1] * x - (ld)q[4] * y + z[2 * s + 1];
        z[2 * s + 1] = (ld)q[2] * x - (ld)q[5] * y;
        x = y;
    }
}

void state_space(const double* sos, int S, Mat& A, std::vector<ld>& B) {
    const int D = 2 * S;
    for (int j = 0; j < D; ++j) {
        std::vector<ld> z(D, 0.0L);
        z[j] = 1.0L;
        step(sos, S, z, 0.0L);
        for (int r = 0; r < D; ++r) A.at(r, j) = z[r];
    }
    std::vector<ld> z(D, 0.0L);
    step(sos, S, z, 1.0L);
    B = z;
}

}  // namespace

SosPlan::~SosPlan() {
    if (dtab) cudaFree(dtab);         // waits for kernels that still read the tables
}

static std::vector<std::shared_ptr<SosPlan>> g_plans;
static std::mutex g_plan_mu;

int32_t get_sos_plan(const double* sos, int S, int CG, cudaStream_t st, std::shared_ptr<SosPlan>* out) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    for (size_t i = 0; i < g_plans.size(); ++i) {
        auto& q = g_plans[i];
        if (q->S == S && q->CG == CG && memcmp(q->sos.data(), sos, sizeof(double) * 6 * S) == 0) {
            *out = q;
            return ADN_OK;
        }
    }
    if (g_plans.size() >= 64) g_plans.erase(g_plans.begin());     // bounded cache: drop the oldest
    const int D = 2 * S, DD = D * D;
    auto pp = std::make_shared<SosPlan>();
    SosPlan& p = *pp;
    p.sos.assign(sos, sos + 6 * S);
    p.S = S;
    p.CG = CG;
    Mat A(D);
    std::vector<ld> B;
    state_space(sos, S, A, B);
    // W[d][i] = (A^(L-1-i) B)[d]
    p.W.assign((size_t)D * SOS_L, 0.0);
    {
        std::vector<ld> v = B;
        for (int i = SOS_L - 1; i >= 0; --i) {
            for (int d = 0; d < D; ++d) p.W[(size_t)d * SOS_L + i] = (double)v[d];
            std::vector<ld> nv(D, 0.0L);
            for (int r = 0; r < D; ++r)
                for (int c = 0; c < D; ++c) nv[r] += A.at(r, c) * v[c];
            v = nv;
        }
    }
    const int GW = 32 / CG;
    int nscan = 0;
    while ((CG << nscan) < 32) ++nscan;
    p.off_fix = nscan;
    p.off_wpow = p.off_fix + GW;
    p.n_staged = p.off_wpow + SOS_NW + 1;
    p.off_tile = p.n_staged;
    std::vector<double> tab((size_t)(p.off_tile + SOS_LOOK + 1) * DD, 0.0);
    auto put = [&](int slot, const Mat& m) {
        for (int i = 0; i < DD; ++i) tab[(size_t)slot * DD + i] = (double)m.a[i];
    };
    const Mat AL = mpow(A, SOS_L);
    {
        Mat m = AL;
        for (int k = 0; k < nscan; ++k) { put(k, m); m = mul(m, m); }
    }
    {
        Mat m = Mat::eye(D);
        for (int j = 0; j < GW; ++j) { put(p.off_fix + j, m); m = mul(m, AL); }
    }
    {
        const Mat AW = mpow(AL, GW);
        Mat m = Mat::eye(D);
        for (int k = 0; k <= SOS_NW; ++k) { put(p.off_wpow + k, m); m = mul(m, AW); }
    }
    {
        const Mat AT = mpow(AL, SOS_NT / CG);         // one tile = NT/CG sub-chunks per channel
        Mat m = Mat::eye(D);
        p.jdecay = SOS_LOOK + 1;
        for (int j = 0; j <= SOS_LOOK; ++j) {
            put(p.off_tile + j, m);
            ld mx = 0.0L;
            for (auto x : m.a) mx = fmaxl(mx, fabsl(x));
            if (j >= 1 && mx < 1e-30L && p.jdecay > SOS_LOOK) p.jdecay = j;
            if (j >= 1 && mx < 1e-20L && p.jpre > SOS_LOOK) p.jpre = j;
            if (j >= 1 && mx < 1e-17L && p.jzp > SOS_LOOK) p.jzp = j;
            m = mul(m, AT);
        }
    }
    ADN_CK(cudaMalloc(&p.dtab, tab.size() * sizeof(double)));
    ADN_CK(cudaMemcpyAsync(p.dtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    ADN_CK(cudaStreamSynchronize(st));           // `tab` is a stack-lifetime host buffer
    g_plans.push_back(pp);
    *out = pp;
    return ADN_OK;
}

int pick_cg(int C) {
    static int cap = -1;
    if (cap < 0) {
        const char* e = getenv("ADN_SOS_CG");
        cap = e ? atoi(e) : 8;       // measured on B200: groups of 8 channels (64-byte row segments)
        if (cap < 1) cap = 1;
        if (cap > 32) cap = 32;
    }
    int cg = 1;
    while (cg < C && cg < cap) cg <<= 1;
    return cg;
}

namespace {

template <int S, int MODE>
int32_t launch_mode(const SosPlan& plan, SosRun& R, size_t smem_bytes, unsigned grid, cudaStream_t st) {
    SosK<S> K;
    fill_sosk<S>(plan, K);
    auto kern = sos_scan_kernel<S, MODE>;
    static bool attr_done = false;               // per instantiation
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    kern<<<grid, SOS_NT, smem_bytes, st>>>(K, R);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

template <int S>
int32_t launch_S(int mode, const SosPlan& plan, SosRun& R, size_t smem, unsigned grid, cudaStream_t st) {
    switch (mode) {
        case MODE_FWD: return launch_mode<S, MODE_FWD>(plan, R, smem, grid, st);
        case MODE_ENVF: return launch_mode<S, MODE_ENVF>(plan, R, smem, grid, st);
        case MODE_ZPF: return launch_mode<S, MODE_ZPF>(plan, R, smem, grid, st);
        default: return launch_mode<S, MODE_REV>(plan, R, smem, grid, st);
    }
}

std::atomic<int64_t> g_run_launches{0};

template <int S, int MODE>
int32_t launch_run_mode(const SosPlan& plan, SosRun& R, size_t smem_bytes, unsigned grid, cudaStream_t st) {
    SosK<S> K;
    fill_sosk<S>(plan, K);
    auto kern = sos_run_kernel<S, MODE>;
    static bool attr_done = false;               // per instantiation
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    kern<<<grid, SOS_NT, smem_bytes, st>>>(K, R);
    count_launch();
    g_run_launches.fetch_add(1);
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

template <int S>
int32_t launch_run_S(int mode, const SosPlan& plan, SosRun& R, size_t smem, unsigned grid, cudaStream_t st) {
    switch (mode) {
        case MODE_FWD: return launch_run_mode<S, MODE_FWD>(plan, R, smem, grid, st);
        case MODE_ENVF: return launch_run_mode<S, MODE_ENVF>(plan, R, smem, grid, st);
        case MODE_ZPF: return launch_run_mode<S, MODE_ZPF>(plan, R, smem, grid, st);
        default: return launch_run_mode<S, MODE_REV>(plan, R, smem, grid, st);
    }
}

// the run kernel when the cascade forgets fast enough for its run-in to be cheap; false: use the
// look-back kernel
bool plan_runs(const SosPlan& plan, SosRun& R, int S, size_t* smem_out, unsigned* grid_out) {
    static int enabled = -1, nbuf_env = 0;
    if (enabled < 0) {
        const char* e = getenv("ADN_SOS_RUN");           // development switch; the option rules
        enabled = e ? atoi(e) : 1;
        const char* b = getenv("ADN_SOS_NBUF");
        nbuf_env = b ? atoi(b) : 2;
        if (nbuf_env < 2) nbuf_env = 2;
        if (nbuf_env > 3) nbuf_env = 3;
    }
    if (!enabled || !option(ADN_OPT_SCAN_RUNS) || S > 4 || R.dst == nullptr || plan.jpre > SOS_LOOK) return false;
    const int CG = R.CG, D = 2 * S;
    const int pad = CG < 16 ? CG : 0;
    const size_t TS = (size_t)(SOS_NT / CG) * (SOS_L * CG + pad);
    const size_t smem = ((size_t)nbuf_env * TS + (size_t)(SOS_NW + 2) * CG * D +
                         (size_t)plan.n_staged * D * D) * 8;
    int bps = (int)((227 * 1024) / (smem + 1024));
    if (bps > SOS_RUN_BLOCKS) bps = SOS_RUN_BLOCKS;
    if (bps < 1) return false;
    const int64_t resident = (int64_t)ctx().sm_count * bps;
    int64_t runs = resident / R.ngroups;                     // per channel group
    if (runs < 1) runs = 1;
    int64_t run_tiles = (R.ntt + runs - 1) / runs;
    // the run-in may cost a quarter of a run at most
    if (run_tiles < 4 * (int64_t)plan.jpre) run_tiles = 4 * (int64_t)plan.jpre;
    runs = (R.ntt + run_tiles - 1) / run_tiles;
    if (runs * R.ngroups * 2 < resident) return false;       // too short to fill the device this way
    if (run_tiles > 0x3fffffff || runs * R.ngroups > 0x7fffffff) return false;
    R.run_tiles = (int32_t)run_tiles;
    R.pre_tiles = plan.jpre;
    R.nbuf = nbuf_env;
    *smem_out = smem;
    *grid_out = (unsigned)(runs * R.ngroups);
    return true;
}

// one sweep of the scan kernel
int32_t run_scan(int mode, const double* sos, int S, const double* src, int64_t n, int64_t nx,
                 int edge, int32_t C, double* dst, int64_t out_skip, int64_t n_dst, int clamp,
                 const double* s0, double* zf, int tile_slot, cudaStream_t st) {
    if (mode == MODE_FWD && option(ADN_OPT_SCAN_RUNS)) {
        // long traces of cascades that forget fast: the pipelined kernel with the tile in registers
        bool handled = false;
        int32_t rc1 = sosfilt_park_dev(sos, S, src, n, C, out_skip, dst, n_dst, s0, zf, &handled, st);
        if (rc1 || handled) return rc1;
    }
    const int CG = pick_cg(C);
    const int D = 2 * S;
    std::shared_ptr<SosPlan> plan;
    int32_t rc = get_sos_plan(sos, S, CG, st, &plan);
    if (rc) return rc;
    SosRun R;
    memset(&R, 0, sizeof R);
    R.src = src; R.dst = dst; R.tab = plan->dtab; R.s0 = s0; R.zf = zf;
    R.n = n; R.nx = nx; R.out_skip = out_skip; R.n_dst = n_dst;
    R.C = C; R.CG = CG; R.ngroups = (C + CG - 1) / CG;
    R.T = (SOS_NT / CG) * SOS_L;
    R.ntt = (n + R.T - 1) / R.T;
    R.jdecay = plan->jdecay;
    R.edge = edge; R.clamp = clamp;
    const bool even = (C % 2 == 0) && (CG % 2 == 0);
    R.vec_in = even && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    R.vec_out = even && dst && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    R.lc = 0;
    while ((1 << R.lc) < CG) ++R.lc;
    if (C % 2 && CG > 1) R.lc = -1;                   // odd C > 1: rows are not granule aligned
    {
        // measured on B200: no gain at any distance (the loads are not latency bound), so off
        // unless asked for
        const char* e = getenv("ADN_SOS_PREFETCH");
        R.pf_tiles = e ? atoi(e) : 0;
    }
    const int64_t ntiles = R.ntt * R.ngroups;
    if (ntiles > 0x7fffffff) return fail(ADN_ERR_UNSUPPORTED, "sos scan: %lld tiles", (long long)ntiles);
    R.off_fix = plan->off_fix; R.off_wpow = plan->off_wpow; R.off_tile = plan->off_tile;
    R.n_staged = plan->n_staged;
    {
        size_t rsmem = 0;
        unsigned rgrid = 0;
        if (plan_runs(*plan, R, S, &rsmem, &rgrid)) {
            switch (S) {
                case 1: return launch_run_S<1>(mode, *plan, R, rsmem, rgrid, st);
                case 2: return launch_run_S<2>(mode, *plan, R, rsmem, rgrid, st);
                case 3: return launch_run_S<3>(mode, *plan, R, rsmem, rgrid, st);
                case 4: return launch_run_S<4>(mode, *plan, R, rsmem, rgrid, st);
            }
        }
    }
    // tile records: agg | incl, every word SOS_EMPTY until published
    const size_t recs = (size_t)ntiles * CG * D;
    DevBuf& tb = scratch(tile_slot, st);
    if ((rc = tb.reserve(recs * 16))) return rc;
    R.agg = tb.as<double>();
    R.incl = R.agg + recs;
    ADN_CK(cudaMemsetAsync(R.agg, 0xFF, recs * 16, st));
    const int pad = CG < 16 ? CG : 0;
    const size_t smem = ((size_t)(SOS_NT / CG) * (SOS_L * CG + pad) + (size_t)(2 * SOS_NW + 1) * CG * D +
                         (S <= 4 ? (size_t)plan->n_staged * D * D : 0)) * 8;
    const unsigned grid = (unsigned)ntiles;
    switch (S) {
        case 1: return launch_S<1>(mode, *plan, R, smem, grid, st);
        case 2: return launch_S<2>(mode, *plan, R, smem, grid, st);
        case 3: return launch_S<3>(mode, *plan, R, smem, grid, st);
        case 4: return launch_S<4>(mode, *plan, R, smem, grid, st);
        case 5: return launch_S<5>(mode, *plan, R, smem, grid, st);
        case 6: return launch_S<6>(mode, *plan, R, smem, grid, st);
        case 7: return launch_S<7>(mode, *plan, R, smem, grid, st);
        case 8: return launch_S<8>(mode, *plan, R, smem, grid, st);
    }
    return fail(ADN_ERR_INVALID, "sos scan: S=%d", S);
}

}  // namespace

int64_t scan_run_launches() { return g_run_launches.load() + fwd_park_launches(); }

int32_t sosfilt_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                    int64_t nbefore, double* dst, int64_t n_dst, const double* zi, double* zf,
                    cudaStream_t st) {
    if (S == 0) {                                // reference: sos is None -> copy
        if (dst && n_dst > 0)
            ADN_CK(cudaMemcpyAsync(dst, src + nbefore * C, (size_t)n_dst * C * 8,
                                   cudaMemcpyDeviceToDevice, st));
        return ADN_OK;
    }
    // scipy's zi/zf layout (C, S, 2) is the kernel's [C][D] state layout
    return run_scan(MODE_FWD, sos, S, src, n_src, n_src, 0, C, dst, nbefore, n_dst, 0, zi, zf,
                    SCR_SOS_TILES, st);
}

// zi = sosfilt_zi(sos): per section scale * lfilter_zi(b, a), scale *= sum(b)/sum(a)
void sosfilt_zi_host(const double* sos, int S, double* zi) {
    double scale = 1.0;
    for (int s = 0; s < S; ++s) {
        const double* q = sos + 6 * s;
        double b0 = q[0], b1 = q[1], b2 = q[2], a0 = q[3], a1 = q[4], a2 = q[5];
        double B0 = b1 - a1 * b0, B1 = b2 - a2 * b0;
        double det = 1.0 + a1 + a2;
        zi[2 * s] = scale * ((B0 + B1) / det);
        zi[2 * s + 1] = scale * (((1.0 + a1) * B1 - a2 * B0) / det);
        scale *= (b0 + b1 + b2) / (a0 + a1 + a2);
    }
}

// out (C, S, 2) = sosfilt_zi(sos) * x0 per channel: which = 0: x0 = ext[0] = 2 r(0) - r(edge) of
// the raw rows at src (the start of the recording), which = 1: x0 = the row at src
int32_t envelope_state0_dev(const double* sos, int32_t S, const double* src, int32_t C, int32_t edge,
                            int32_t which, double* out, cudaStream_t st) {
    ZiK zik;
    sosfilt_zi_host(sos, S, zik.z);
    const int D = 2 * S;
    env_s0_kernel<<<(C * D + 127) / 128, 128, 0, st>>>(which, src, C, D, edge, zik, out);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

// incoming state of rank `rank` from the gathered boundary records, packs (W, 2, C, D):
// [i][0] = end state of shard i from zero state, [i][1] = state entering the recording (used
// from shard 0 forward, from shard W-1 backward); mats (W, D, D) = A^len(shard i), row major
__global__ void fold_states_kernel(const double* __restrict__ packs, const double* __restrict__ mats,
                                   int32_t W, int32_t C, int32_t D, int32_t rank, int32_t backward,
                                   double* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s[2 * ADN_MAX_SECTIONS], t[2 * ADN_MAX_SECTIONS];
    const int first = backward ? W - 1 : 0;
    const double* init = packs + ((size_t)(first * 2 + 1) * C + c) * D;
    for (int d = 0; d < D; ++d) s[d] = init[d];
    for (int k = 0; k < (backward ? W - 1 - rank : rank); ++k) {
        const int i = backward ? W - 1 - k : k;
        const double* M = mats + (size_t)i * D * D;
        const double* v = packs + ((size_t)(i * 2) * C + c) * D;
        for (int r = 0; r < D; ++r) {
            double a = v[r];
            for (int q = 0; q < D; ++q) a = fma(M[r * D + q], s[q], a);
            t[r] = a;
        }
        for (int d = 0; d < D; ++d) s[d] = t[d];
    }
    for (int d = 0; d < D; ++d) out[(size_t)c * D + d] = s[d];
}

int32_t fold_states_dev(const double* packs, const double* mats, int32_t W, int32_t C, int32_t D,
                        int32_t rank, int32_t backward, double* out, cudaStream_t st) {
    fold_states_kernel<<<(C + 63) / 64, 64, 0, st>>>(packs, mats, W, C, D, rank, backward, out);
    count_launch();
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

static int32_t zero_phase_dev(bool rect, const double* sos, int32_t S, const double* src, int64_t n_src,
                              int32_t C, int64_t nbefore, double* dst, int64_t n_dst,
                              int32_t clamp_negative, cudaStream_t st);

int32_t envelope_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                     int64_t nbefore, double* dst, int64_t n_dst, int32_t clamp_negative,
                     cudaStream_t st) {
    return zero_phase_dev(true, sos, S, src, n_src, C, nbefore, dst, n_dst, clamp_negative, st);
}

// scipy.signal.sosfiltfilt(sos, src, axis=0)[nbefore:][:n_dst] (default odd padding)
int32_t sosfiltfilt_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                        int64_t nbefore, double* dst, int64_t n_dst, cudaStream_t st) {
    return zero_phase_dev(false, sos, S, src, n_src, C, nbefore, dst, n_dst, 0, st);
}

// rect: the envelope ((pi/2)|x| in front of the filter); else the plain zero-phase filter
static int32_t zero_phase_dev(bool rect, const double* sos, int32_t S, const double* src, int64_t n_src,
                              int32_t C, int64_t nbefore, double* dst, int64_t n_dst,
                              int32_t clamp_negative, cudaStream_t st) {
    const int D = 2 * S;
    const int edge = adn_sosfiltfilt_edge(sos, S);
    const int64_t next = n_src + 2 * (int64_t)edge;
    {
        // one pass with the tile in registers when the cascade forgets fast enough (zerophase.cu)
        bool handled = false;
        int32_t rc1 = zero_phase_regs_dev(rect, sos, S, src, n_src, C, edge, edge, edge + nbefore, dst, n_dst,
                                          clamp_negative, &handled, st);
        if (rc1 || handled) return rc1;
    }
    ZiK zik;
    sosfilt_zi_host(sos, S, zik.z);
    DevBuf& fwd = scratch(SCR_ENV_FWD, st);
    DevBuf& misc = scratch(SCR_ENV_MISC, st);
    int32_t rc;
    if ((rc = fwd.reserve((size_t)next * C * 8))) return rc;
    if ((rc = misc.reserve((size_t)(2 * C * D) * 8))) return rc;
    double* d_s0f = misc.as<double>();
    double* d_s0b = d_s0f + (size_t)C * D;
    const int nb = (C * D + 127) / 128;
    env_s0_kernel<<<nb, 128, 0, st>>>(rect ? 0 : 2, src, C, D, edge, zik, d_s0f);
    count_launch();
    ADN_CK(cudaGetLastError());
    double* y1 = fwd.as<double>();
    if ((rc = run_scan(rect ? MODE_ENVF : MODE_ZPF, sos, S, src, next, n_src, edge, C, y1, 0, next, 0, d_s0f, nullptr,
                       SCR_SOS_TILES, st)))
        return rc;
    env_s0_kernel<<<nb, 128, 0, st>>>(1, y1 + (next - 1) * C, C, D, 0, zik, d_s0b);
    count_launch();
    ADN_CK(cudaGetLastError());
    return run_scan(MODE_REV, sos, S, y1, next, next, 0, C, dst, edge + nbefore, n_dst,
                    clamp_negative ? 1 : 0, d_s0b, nullptr, SCR_SOS_MISC, st);
}

// sosfiltfilt (rect: of (pi/2)|src|) over a RANGE of a longer recording: scipy's odd extension and
// sosfilt_zi initial conditions only at the ends that are ends of the recording (edge_left /
// edge_right != 0); at the other ends the sweeps start from zero state, so the caller passes
// enough extra rows there for the cascade to forget it (adn_sos_decay_length) and drops them:
// dst rows = result rows [first, first + n_dst) of the n_src rows.  One pass in registers where
// the cascade allows it, else the two sweeps through memory.
int32_t zero_phase_range_dev(bool rect, const double* sos, int32_t S, const double* src, int64_t n_src,
                             int32_t C, int32_t edge_left, int32_t edge_right, int64_t first, double* dst,
                             int64_t n_dst, int32_t clamp_negative, cudaStream_t st) {
    const int D = 2 * S;
    const int edge = adn_sosfiltfilt_edge(sos, S);
    const int el = edge_left ? edge : 0, er = edge_right ? edge : 0;
    {
        bool handled = false;
        int32_t rc1 = zero_phase_regs_dev(rect, sos, S, src, n_src, C, el, er, el + first, dst, n_dst,
                                          clamp_negative, &handled, st);
        if (rc1 || handled) return rc1;
    }
    const int64_t next = n_src + el + er;
    ZiK zik;
    sosfilt_zi_host(sos, S, zik.z);
    DevBuf& fwd = scratch(SCR_ENV_FWD, st);
    DevBuf& misc = scratch(SCR_ENV_MISC, st);
    int32_t rc;
    if ((rc = fwd.reserve((size_t)next * C * 8))) return rc;
    if ((rc = misc.reserve((size_t)(2 * C * D) * 8))) return rc;
    double* d_s0f = misc.as<double>();
    double* d_s0b = d_s0f + (size_t)C * D;
    const int nb = (C * D + 127) / 128;
    if (el) {
        env_s0_kernel<<<nb, 128, 0, st>>>(rect ? 0 : 2, src, C, D, edge, zik, d_s0f);
        count_launch();
    }
    double* y1 = fwd.as<double>();
    if ((rc = run_scan(rect ? MODE_ENVF : MODE_ZPF, sos, S, src, next, n_src, el, C, y1, 0, next, 0,
                       el ? d_s0f : nullptr, nullptr, SCR_SOS_TILES, st)))
        return rc;
    if (er) {
        env_s0_kernel<<<nb, 128, 0, st>>>(1, y1 + (next - 1) * C, C, D, 0, zik, d_s0b);
        count_launch();
    }
    ADN_CK(cudaGetLastError());
    return run_scan(MODE_REV, sos, S, y1, next, next, 0, C, dst, el + first, n_dst,
                    clamp_negative ? 1 : 0, er ? d_s0b : nullptr, nullptr, SCR_SOS_MISC, st);
}

// The two sweeps of the envelope as separate steps (time-sharded recordings run them with the
// boundary states exchanged in between): forward sosfilt of (pi/2)|src| with scipy's odd
// extension of edge_left / edge_right rows at the respective end (0 = none), from state zi;
// dst (if given) receives all edge_left + n_src + edge_right rows.
int32_t envelope_forward_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                             int32_t edge_left, int32_t edge_right, const double* zi, double* dst,
                             double* zf, cudaStream_t st) {
    const int64_t next = n_src + edge_left + edge_right;
    return run_scan(MODE_ENVF, sos, S, src, next, n_src, edge_left, C, dst, 0, next, 0, zi, zf,
                    SCR_SOS_TILES, st);
}

// Time-reversed sosfilt: rows are filtered from the last to the first starting from state zi;
// dst[i] (if given) = result at row first + i, i < n_dst; zf = state after row 0.
int32_t sosfilt_reverse_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                            const double* zi, double* dst, int64_t first, int64_t n_dst,
                            int32_t clamp_negative, double* zf, cudaStream_t st) {
    return run_scan(MODE_REV, sos, S, src, n_src, n_src, 0, C, dst, first, dst ? n_dst : 0,
                    clamp_negative ? 1 : 0, zi, zf, SCR_SOS_MISC, st);
}

}  // namespace adn

using namespace adn;

extern "C" {

int32_t adn_sos_state_space(const double* sos, int32_t S, double* A, double* B, int64_t power,
                            double* A_pow) {
    if (!sos || S < 1 || S > ADN_MAX_SECTIONS || power < 0)
        return fail(ADN_ERR_INVALID, "adn_sos_state_space: bad arguments");
    const int D = 2 * S;
    Mat Am(D);
    std::vector<ld> Bv;
    state_space(sos, S, Am, Bv);
    if (A) for (int i = 0; i < D * D; ++i) A[i] = (double)Am.a[i];
    if (B) for (int i = 0; i < D; ++i) B[i] = (double)Bv[i];
    if (A_pow) {
        Mat P = mpow(Am, power);
        for (int i = 0; i < D * D; ++i) A_pow[i] = (double)P.a[i];
    }
    return ADN_OK;
}

int64_t adn_sos_decay_length(const double* sos, int32_t S, double tol) {
    if (!sos || S < 1 || S > ADN_MAX_SECTIONS || !(tol > 0)) return -1;
    const int D = 2 * S;
    Mat P(D);
    std::vector<ld> Bv;
    state_space(sos, S, P, Bv);
    for (int j = 0; j <= 40; ++j) {
        ld mx = 0.0L;
        for (auto x : P.a) mx = fmaxl(mx, fabsl(x));
        if (!(mx == mx)) return -1;
        if (mx < (ld)tol) return (int64_t)1 << j;
        P = mul(P, P);
    }
    return -1;
}

int32_t adn_sosfiltfilt_edge(const double* sos, int32_t S) {
    if (!sos || S < 1) return 0;
    int nb = 0, na = 0;
    for (int s = 0; s < S; ++s) {
        if (sos[6 * s + 2] == 0.0) ++nb;
        if (sos[6 * s + 5] == 0.0) ++na;
    }
    int ntaps = 2 * S + 1 - (nb < na ? nb : na);
    return 3 * ntaps;
}

}  // extern "C"
