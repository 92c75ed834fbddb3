// Zero-phase SOS filtering (scipy sosfiltfilt; with the rectification in front of it:
// BufferedEnvelope.process, src/audian/bufferedenvelope.py:34-41; without:
// the play-back low-pass, src/audian/databrowser.py:1725) in ONE pass over the input, the
// tile held in registers: HBM sees every input row once and every output row once
// (16 B per sample instead of the 32 B of a forward and a backward sweep through memory).
//
// Geometry as in sosfilt.cu: a tile = SOS_NT threads x SOS_L consecutive samples, lane (gl, cw)
// of warp w owns sub-chunk g = w GW + gl of channel cw of a group of CG channels.  A block
// owns a run of consecutive tiles and walks BACKWARD along time:
//   - the backward sweep carries its state exactly from tile to tile (the state after the
//     first row of tile T+1 enters the last row of tile T);
//   - the forward state entering tile T comes from the zero-state aggregates of the JJ tiles
//     before it, sum_j (A^T)^(j-1) agg[T-j] (a cascade that forgets its state within JJ - 1
//     tiles to 1e-20: the truncation sosfilt.cu's run kernel makes); those aggregates need
//     pass A only (dot products), computed when a tile is visited JJ iterations ahead of its
//     own turn ("look-ahead visit": tile read from HBM; the second read, JJ tiles later, is
//     an L2 hit);
//   - a main visit loads the tile into registers, runs the exact DF2T recurrence forward from
//     the true incoming state (y1 replaces x in the registers), pass A + scan + DF2T backward
//     over the same registers (y2 replaces y1), clamps and stores from the registers: each
//     warp store covers GW rows x CG channels = full 32-byte sectors for CG >= 4.
// A run starts JO tiles behind its last output tile (run-out of the backward sweep from zero
// state) unless it reaches the end of the sequence, where scipy's zi * y1[last] applies.
// output stores evict-first here (measured on B200, envelope of 8 ch x 48 kHz x 80 s: 120.1 against
// 123.2 us with the default policy; the forward filter and the spectrogram prefer the default)
#ifndef ADN_STORE_CS
#define ADN_STORE_CS 1
#endif
#include "sos_common.cuh"
#include <cstring>
#include <cstdlib>
#include <atomic>

namespace adn {

namespace {

constexpr int ZP_MAX_JJ = 16;

struct ZpArgs {
    const double* src;
    double* dst;
    const double* tab;                 // plan tables in global memory
    int32_t off_fix, off_wpow, off_tile, n_staged;
    int64_t nx;                        // raw rows
    int64_t N;                         // rows of the extended sequence: edgeL + nx + edgeR
    int64_t out_first, n_dst;          // sequence rows [out_first, out_first + n_dst) -> dst rows
    int64_t ntt;                       // tiles of the sequence
    int64_t t_out0, t_out1;            // output tiles [t_out0, t_out1)
    int32_t C, CG, ngroups, T;
    int32_t edgeL, edgeR;
    int32_t zi_left, zi_right;         // sosfilt_zi initial conditions at that end (else zero state)
    int32_t clamp;
    int32_t JJ, JO;                    // look-ahead tiles of the forward state; run-out tiles
    int32_t run_tiles;
    int32_t short0;                    // the first run of the pipelined kernel is this many tiles shorter
    int32_t pf;                        // bulk L2 prefetch of the next look-ahead tile
    int32_t bulk_ok;                   // tiles of full, contiguous channel groups through the TMA unit
    ZiK zi;
};

template <bool RECT> __device__ __forceinline__ double zp_pre(double x) {
    return RECT ? HALF_PI * fabs(x) : x;
}

// value of the extended sequence at row e of the channel whose column starts at xc
template <bool RECT>
__device__ __forceinline__ double zp_ext_value(const ZpArgs& P, const double* __restrict__ xc, int64_t e) {
    if (e < 0 || e >= P.N) return 0.0;
    const int64_t C = P.C;
    if (e < P.edgeL)
        return 2.0 * zp_pre<RECT>(__ldg(xc)) - zp_pre<RECT>(__ldg(xc + (P.edgeL - e) * C));
    e -= P.edgeL;
    if (e < P.nx) return zp_pre<RECT>(__ldg(xc + e * C));
    e -= P.nx;
    return 2.0 * zp_pre<RECT>(__ldg(xc + (P.nx - 1) * C)) - zp_pre<RECT>(__ldg(xc + (P.nx - 2 - e) * C));
}

#ifndef ZP_BLOCKS
#define ZP_BLOCKS 4
#endif

template <int S, bool RECT>
__global__ void __launch_bounds__(SOS_NT, S <= 2 ? ZP_BLOCKS : 3)
sos_zp_kernel(const __grid_constant__ SosK<S> K, const __grid_constant__ ZpArgs P) {
    constexpr int D = 2 * S;
    constexpr int DD = D * D;
    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = (int)(blockIdx.x % P.ngroups);
    const int64_t run = blockIdx.x / P.ngroups;
    const int CG = P.CG, C = P.C;
    const int c0 = grp * CG;
    const int Cw = min(CG, C - c0);
    const int GW = 32 / CG;
    const int gl = lane / CG, cw = lane % CG;
    const int g = warp * GW + gl;
    const bool chan_ok = cw < Cw;
    const int T = P.T;
    const int JJ = P.JJ, NS = JJ + 1;

    double* tab_s = smem;                                   // n_staged * DD
    double* wagg = tab_s + (size_t)P.n_staged * DD;         // [NW][CG][D]
    double* sinf_s = wagg + SOS_NW * CG * D;                // [CG][D]
    double* sinb_s = sinf_s + CG * D;                       // [2][CG][D]
    double* aggr = sinb_s + 2 * CG * D;                     // [NS][CG][D]   zero-state tile aggregates
    double* etot = aggr + (size_t)NS * CG * D;              // [NS][D][NT]   per-thread zero-state offsets
    const double* tab_fix = tab_s + P.off_fix * DD;
    const double* tab_wpow = tab_s + P.off_wpow * DD;
    const double* Pt = P.tab + (size_t)P.off_tile * DD;     // (A^T)^j, global

    const int64_t a = P.t_out0 + run * P.run_tiles;         // output tiles [a, b)
    const int64_t b = min(a + (int64_t)P.run_tiles, P.t_out1);
    if (a >= b) return;
    const int64_t top = min(b + (int64_t)P.JO, P.ntt) - 1;  // first tile of the walk

    for (int q = tid; q < P.n_staged * DD; q += SOS_NT) tab_s[q] = __ldg(P.tab + q);
    if (tid < 2 * CG * D) sinb_s[tid] = 0.0;

    const double* xc = P.src + c0 + (chan_ok ? cw : 0);
    auto slot_of = [&](int64_t t) { return (int)((t + 64 * (int64_t)NS) % NS); };

    auto prefetch = [&](int64_t t) {
        if (!(P.pf & 1) || tid != 0 || grp != 0 || t < 0 || t >= P.ntt) return;
        int64_t p0 = t * T - P.edgeL, p1 = p0 + T;
        if (p0 < 0) p0 = 0;
        if (p1 > P.nx) p1 = P.nx;
        if (p1 <= p0) return;
        const char* adr = reinterpret_cast<const char*>(P.src + p0 * C);
        int64_t bytes = (p1 - p0) * (int64_t)C * 8;
        const int64_t mis = reinterpret_cast<uintptr_t>(adr) & 15;
        adr -= mis;
        bytes = (bytes + mis + 15) & ~(int64_t)15;
        if (adr >= reinterpret_cast<const char*>(P.src))
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(adr), "r"((uint32_t)bytes) : "memory");
    };

    __syncthreads();
    // The walk, one tile visit per iteration: look-ahead visits L(u) of tiles u = top, top-1, ...
    // interleaved with the main visits V(u + JJ) once JJ + 1 look-ahead visits have filled the
    // ring: L(top) .. L(top-JJ) V(top) L(top-JJ-1) V(top-1) ... L(a-JJ) V(a).  Both kinds start
    // with the same tile load (one copy of that code).
    int par = 0;                                            // parity of the backward state slot
    const int64_t nvis = 2 * (top - (a - JJ) + 1);
    for (int64_t it = 0; it < nvis; ++it) {
        const int64_t u = top - (it >> 1);
        const bool is_main = it & 1;
        const int64_t t = is_main ? u + JJ : u;
        if (is_main && t > top) continue;
        const int slot = slot_of(t);
        if (!is_main && t < 0) {
            // before the sequence: the virtual tile -1 carries the initial state of the sweep
            if (tid < CG * D) {
                const int ch = tid / D, d = tid - ch * D;
                double v = 0.0;
                if (t == -1 && P.zi_left && c0 + ch < C)
                    v = P.zi.z[d] * zp_ext_value<RECT>(P, P.src + c0 + ch, 0);
                aggr[(size_t)slot * CG * D + tid] = v;
            }
            __syncthreads();
            continue;
        }
        // ---- rows of tile t of this thread's channel -> registers
        double x[SOS_L];
        {
            const int64_t e0 = t * T + (int64_t)g * SOS_L;  // sequence row of x[0]
            const bool fast = t * T >= P.edgeL && (t + 1) * (int64_t)T <= P.edgeL + P.nx;
            if (!chan_ok) {
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = 0.0;
            } else if (fast) {
                const double* p = xc + (e0 - P.edgeL) * C;
                if (C == 8) {
#pragma unroll
                    for (int i = 0; i < SOS_L; ++i) x[i] = __ldg(p + i * 8);
                } else {
#pragma unroll
                    for (int i = 0; i < SOS_L; ++i) { x[i] = __ldg(p); p += C; }
                }
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = zp_pre<RECT>(x[i]);
            } else {
                // the two edge tiles and the tail: element by element through local memory (a
                // rolled loop: this path is rare and must not cost code size)
                double tmp[SOS_L];
#pragma unroll 1
                for (int i = 0; i < SOS_L; ++i) tmp[i] = zp_ext_value<RECT>(P, xc, e0 + i);
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = tmp[i];
            }
        }
        if (!is_main) {
            // ---- look-ahead visit: zero-state forward aggregate of the tile and of every
            // thread's prefix inside it
            double v[D];
#pragma unroll
            for (int d = 0; d < D; ++d) v[d] = 0.0;
            zp_pass_a<S, false>(K, x, v);
            {
                int k = 0;
                for (int off = CG; off < 32; off <<= 1, ++k) {
                    double w[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) w[d] = __shfl_up_sync(0xffffffffu, v[d], off);
                    if (lane >= off) matvec_acc<D>(tab_s + k * DD, w, v);
                }
            }
            double ex[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double w = __shfl_up_sync(0xffffffffu, v[d], CG & 31);
                ex[d] = gl == 0 ? 0.0 : w;
            }
            if (gl == GW - 1) {
#pragma unroll
                for (int d = 0; d < D; ++d) wagg[(warp * CG + cw) * D + d] = v[d];
            }
            __syncthreads();
            double pre[D];
#pragma unroll
            for (int d = 0; d < D; ++d) pre[d] = 0.0;
            for (int j = 0; j < warp; ++j) {
                double w[D];
#pragma unroll
                for (int d = 0; d < D; ++d) w[d] = wagg[(j * CG + cw) * D + d];
                matvec_acc<D>(tab_wpow + (warp - 1 - j) * DD, w, pre);
            }
            matvec_acc<D>(tab_fix + gl * DD, pre, ex);       // ex + A^(L gl) pre
#pragma unroll
            for (int d = 0; d < D; ++d) etot[((size_t)slot * D + d) * SOS_NT + tid] = ex[d];
            if (tid < CG * D) {
                // aggregate of the tile: sum_w A^(L GW (NW-1-w)) wagg[w]
                const int ch = tid / D, r = tid - ch * D;
                double acc = 0.0;
                for (int w = 0; w < SOS_NW; ++w) {
                    const double* M = tab_wpow + (SOS_NW - 1 - w) * DD + r * D;
                    const double* q = wagg + (w * CG + ch) * D;
                    for (int c = 0; c < D; ++c) acc = fma(M[c], q[c], acc);
                }
                aggr[(size_t)slot * CG * D + tid] = acc;
            }
            __syncthreads();
            continue;
        }
        // ---- main visit of tile t
        const bool store = t < b;
        const bool last = t == P.ntt - 1;
        prefetch(t - 2 - JJ);
        if (tid < CG * D) {
            // forward state entering the tile: sum_{j=1..JJ} (A^T)^(j-1) agg[t-j]
            const int ch = tid / D, r = tid - ch * D;
            double acc = 0.0;
            for (int j = 1; j <= JJ; ++j) {
                const double* M = Pt + (size_t)(j - 1) * DD + r * D;
                const double* q = aggr + (size_t)slot_of(t - j) * CG * D + ch * D;
                for (int c = 0; c < D; ++c) acc = fma(__ldg(M + c), q[c], acc);
            }
            sinf_s[tid] = acc;
        }
        __syncthreads();
        double z[D];
        {
            double sv[D], tmp[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                sv[d] = sinf_s[cw * D + d];
                tmp[d] = 0.0;
                z[d] = etot[((size_t)slot * D + d) * SOS_NT + tid];
            }
            matvec_acc<D>(tab_wpow + warp * DD, sv, tmp);
            matvec_acc<D>(tab_fix + gl * DD, tmp, z);
        }
        zp_df2t<S, false>(K, x, z);                          // x <- y1
        // ---- backward sweep over the same registers
        double vb[D];
#pragma unroll
        for (int d = 0; d < D; ++d) vb[d] = 0.0;
        int lstar = SOS_L;                                   // last tile: index of the last row in its owner
        bool owner = false;
        double sb0[D];
#pragma unroll
        for (int d = 0; d < D; ++d) sb0[d] = 0.0;
        if (last) {
            const int il = (int)(P.N - 1 - t * T);
            const int gstar = il / SOS_L;
            lstar = il - gstar * SOS_L;
            owner = g == gstar;
            double ylast = 0.0;
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) {
                if (owner && i == lstar) ylast = x[i];
                if (g > gstar || (owner && i > lstar)) x[i] = 0.0;
            }
            if (owner && P.zi_right) {
#pragma unroll
                for (int d = 0; d < D; ++d) sb0[d] = P.zi.z[d] * ylast;
            }
        }
        zp_pass_a<S, true>(K, x, vb);
        if (last && owner) {
            // the state zi * y1[last] enters at row lstar: seen from the rows before, it has
            // been carried over lstar + 1 rows
            double h[D];
#pragma unroll
            for (int d = 0; d < D; ++d) h[d] = sb0[d];
            for (int k = 0; k <= lstar; ++k) zp_step0<S>(K, h);
#pragma unroll
            for (int d = 0; d < D; ++d) vb[d] += h[d];
        }
        {
            int k = 0;
            for (int off = CG; off < 32; off <<= 1, ++k) {
                double w[D];
#pragma unroll
                for (int d = 0; d < D; ++d) w[d] = __shfl_down_sync(0xffffffffu, vb[d], off);
                if (lane + off < 32) matvec_acc<D>(tab_s + k * DD, w, vb);
            }
        }
        double zb[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double w = __shfl_down_sync(0xffffffffu, vb[d], CG & 31);
            zb[d] = gl == GW - 1 ? 0.0 : w;
        }
        if (gl == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) wagg[(warp * CG + cw) * D + d] = vb[d];
        }
        __syncthreads();
        {
            double pre[D];
#pragma unroll
            for (int d = 0; d < D; ++d) pre[d] = 0.0;
            for (int j = SOS_NW - 1; j > warp; --j) {
                double w[D];
#pragma unroll
                for (int d = 0; d < D; ++d) w[d] = wagg[(j * CG + cw) * D + d];
                matvec_acc<D>(tab_wpow + (j - 1 - warp) * DD, w, pre);
            }
            double sv[D];
#pragma unroll
            for (int d = 0; d < D; ++d) sv[d] = last ? 0.0 : sinb_s[par * CG * D + cw * D + d];
            matvec_acc<D>(tab_wpow + (SOS_NW - 1 - warp) * DD, sv, pre);
            matvec_acc<D>(tab_fix + (GW - 1 - gl) * DD, pre, zb);
        }
        if (last) {
            // rolled loop through local memory: the owner of the last row starts there from
            // zi * y1[last]; rows behind it are not part of the sequence
            double tmp[SOS_L];
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) tmp[i] = x[i];
#pragma unroll 1
            for (int i = SOS_L - 1; i >= 0; --i) {
                if (owner && i == lstar) {
#pragma unroll
                    for (int d = 0; d < D; ++d) zb[d] = sb0[d];
                }
                double xv = tmp[i];
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    const double y = fma(K.coef[q][0], xv, zb[2 * q]);
                    zb[2 * q] = fma(K.coef[q][1], xv, zb[2 * q + 1]) - K.coef[q][3] * y;
                    zb[2 * q + 1] = K.coef[q][2] * xv - K.coef[q][4] * y;
                    xv = y;
                }
                tmp[i] = xv;
            }
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) x[i] = tmp[i];
        } else {
            zp_df2t<S, true>(K, x, zb);                      // x <- y2
        }
        if (g == 0 && chan_ok) {
#pragma unroll
            for (int d = 0; d < D; ++d) sinb_s[(par ^ 1) * CG * D + cw * D + d] = zb[d];
        }
        par ^= 1;
        if (store && chan_ok) {
            if (P.clamp) {
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = 0.5 * (x[i] + fabs(x[i]));   // max(x, 0), two fp64 ops
            }
            const int64_t e0 = t * T + (int64_t)g * SOS_L;
            const bool fast = t * T >= P.out_first && (t + 1) * (int64_t)T <= P.out_first + P.n_dst;
            double* p = P.dst + (e0 - P.out_first) * C + c0 + cw;
            if (fast && C == 8) {
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) ADN_STORE(p + i * 8, x[i]);
            } else {
                const int64_t o0 = e0 - P.out_first;
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) {
                    if (fast || (o0 + i >= 0 && o0 + i < P.n_dst)) ADN_STORE(p, x[i]);
                    p += C;
                }
            }
        }
        __syncthreads();
    }
}

// ======================================================================================
// Pipelined variant for cascades that forget within a few tiles: nothing is read twice, not even
// from L2, and every warp works all the time.  A block is NTEAM teams of four warps and owns ONE
// run; tiles are staged in a ring of NTEAM + JJ shared-memory slots by cp.async (the loader of
// sosfilt.cu: coalesced 16-byte granules, the odd extension built on the fly) one team
// iteration ahead of their use.  Iteration of a team for walk position w (tile t = top - w):
//   L(t)       tile slot -> registers, zero-state forward aggregates (pass A + scans), the tile
//              aggregate and every thread's prefix are published in shared memory; the samples
//              stay PARKED in their slot
//   M(t + JJ)  tile slot -> registers again (parked JJ positions ago by another team), the
//              slot is handed to the prefetch of the team's next tile; forward state from the JJ
//              aggregates behind it, forward recurrence, backward pass A + scan and backward
//              recurrence in the same registers, clamp, store from the registers.
// The backward state is handed from tile to tile as  sin_b(T) = A^T sin_b(T+1) + agg_b(T)  by one
// warp as soon as agg_b(T) is known, so only that small matrix-vector product is serial along
// the run and the teams overlap freely.
constexpr int ZP_NTEAM_MAX = 4;

#ifdef ADN_ZP_TIMING
// development builds only (tools/build_alt.sh ... -DADN_ZP_TIMING): per block start, end of every
// team and the SM it ran on, read back by adn_debug_zp_times
__device__ unsigned long long zp_times[2 * 1024 * 8];     // [launch parity][block][8]
__device__ unsigned int zp_count;
__device__ unsigned long long zp_tile_times[4 * 64];      // runs 0..3: end of the main visit of tile a + i
__device__ __forceinline__ unsigned long long zp_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

template <int S, int MODE, int NTC>
__global__ void __launch_bounds__(SOS_NT * NTC, 1)
sos_zp_park_kernel(const __grid_constant__ SosK<S> K, const __grid_constant__ ZpArgs P,
                   const __grid_constant__ SosRun R) {
    constexpr int D = 2 * S;
    constexpr int DD = D * D;
    constexpr bool RECT = MODE == MODE_ENVF;
    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x, lane = tid & 31;
    const int team = tid >> 7, ttid = tid & (SOS_NT - 1), warp = ttid >> 5;   // warp inside the team
    constexpr int NTEAM = NTC;
    const int grp = (int)(blockIdx.x % P.ngroups);
#ifdef ADN_ZP_TIMING
    // development: runs rotated over the blocks (is a slow block slow because of its run or its SM?)
    const int64_t nruns_dbg = (P.t_out1 - P.t_out0 + P.run_tiles - 1) / P.run_tiles;
    const int64_t run = (blockIdx.x / P.ngroups + (P.pf >> 8)) % nruns_dbg;
#else
    const int64_t run = blockIdx.x / P.ngroups;
#endif
    const int CG = P.CG, C = P.C;
    const int c0 = grp * CG;
    const int Cw = min(CG, C - c0);
    const int GW = 32 / CG;
    const int gl = lane / CG, cw = lane % CG;
    const int g = warp * GW + gl;
    const bool chan_ok = cw < Cw;
    const int T = P.T;
    const int JJ = P.JJ, NSLOT = NTEAM + JJ;
    const int pad = CG < 16 ? CG : 0;
    const int GS = SOS_L * Cw + pad;
    const size_t TS = (size_t)(SOS_NT / CG) * (SOS_L * CG + pad);   // doubles per tile slot

    double* tab_s = smem;                                   // n_staged * DD
    double* wagg = tab_s + (size_t)P.n_staged * DD + (size_t)team * 2 * SOS_NW * CG * D;   // [NTEAM][2][NW][CG][D]
    double* waggb = wagg + (size_t)SOS_NW * CG * D;
    double* aggr = tab_s + (size_t)P.n_staged * DD + (size_t)NTEAM * 2 * SOS_NW * CG * D;   // [NSLOT][CG][D]
    double* sinb_s = aggr + (size_t)NSLOT * CG * D;         // [NTEAM][CG][D]
    double* etot = sinb_s + (size_t)NTEAM * CG * D;         // [NSLOT][D][NT]
    volatile long long* la_flag = reinterpret_cast<volatile long long*>(etot + (size_t)NSLOT * D * SOS_NT);  // [NSLOT]
    volatile long long* sb_flag = la_flag + NSLOT;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(const_cast<long long*>(la_flag) + NSLOT + 1);         // [NSLOT]
    double* pt_s = reinterpret_cast<double*>(mbar + NSLOT + ((NSLOT + 1 + NSLOT) & 1));   // [(JJ + 2)][DD], 16-byte aligned
    double* tiles = pt_s + (size_t)(JJ + 2) * DD;           // [NSLOT][TS]
    const double* tab_fix = tab_s + P.off_fix * DD;
    const double* tab_wpow = tab_s + P.off_wpow * DD;
    const double* Pt = pt_s;                                // (A^T)^j, j <= JJ + 1
    const int G = SOS_NT / CG;
    // whole tiles come in through the TMA unit when the rows of the group are contiguous
    const bool bulk_group = P.bulk_ok && Cw == CG;

    // output tiles [a, b); the run that builds the left extension tile is shorter by what that costs
    const int64_t a = run == 0 ? P.t_out0 : P.t_out0 + run * P.run_tiles - P.short0;
    const int64_t b = min(P.t_out0 + (run + 1) * P.run_tiles - P.short0, P.t_out1);
    if (a >= b) return;
    const int64_t top = min(b + (int64_t)P.JO, P.ntt) - 1;  // first tile of the walk
#ifdef ADN_ZP_TIMING
    __shared__ unsigned zp_slot_s;
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned long long now = zp_now();
        zp_slot_s = ((atomicAdd(&zp_count, 1u) / gridDim.x) & 1u) * 1024u + blockIdx.x;
        zp_times[zp_slot_s * 8] = now;
        zp_times[zp_slot_s * 8 + 6] = smid;
    }
#endif

    for (int q = tid; q < P.n_staged * DD; q += blockDim.x) tab_s[q] = __ldg(P.tab + q);
    for (int q = tid; q < (JJ + 2) * DD; q += blockDim.x) pt_s[q] = __ldg(P.tab + (size_t)P.off_tile * DD + q);
    for (int q = tid; q < NTEAM * CG * D; q += blockDim.x) sinb_s[q] = 0.0;
    if (tid < NSLOT) {
        la_flag[tid] = (long long)1 << 60;
        zp_mbar_init(mbar + tid, 1);
    }
    if (tid == 0) *sb_flag = top + 1;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
#ifdef ADN_ZP_TIMING
    if (tid == 0) zp_times[zp_slot_s * 8 + 1] = zp_now();
#endif
    auto slot_of = [&](int64_t t) { return (int)((t + 64 * (int64_t)NSLOT) % NSLOT); };
    // how tile t reaches its slot: 0 nothing to load, 1 TMA bulk copies, 2 cp.async granules
    auto load_kind = [&](int64_t t) {
        if (t < 0 || t < a - JJ) return 0;
        if (!bulk_group || t * T < P.edgeL || (t + 1) * (int64_t)T > P.edgeL + P.nx) return 2;
        const uintptr_t g0 = reinterpret_cast<uintptr_t>(P.src + (t * T - P.edgeL) * C + c0);
        return (g0 & 15) == 0 ? 1 : 2;
    };
    // every use of a slot completes exactly one phase of its mbarrier (count 1): the bulk copies'
    // bytes + the arrival that announced them, or a plain arrival for the other two kinds
    auto issue_load = [&](int64_t t, int slot) {
        const int kind = load_kind(t);
        if (kind == 1) {
            if (warp == 0) {
                // the slot was read through the generic proxy until the barrier before this call
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (lane == 0) zp_mbar_expect_tx(mbar + slot, (uint32_t)(T * C * 8));
                __syncwarp();
                const double* gsrc = P.src + (t * T - P.edgeL) * C + c0;
                double* sdst = tiles + (size_t)slot * TS;
                for (int q = lane; q < G; q += 32)
                    zp_bulk_g2s(sdst + (size_t)q * GS, gsrc + (size_t)q * SOS_L * C, (uint32_t)(SOS_L * C * 8), mbar + slot);
            }
        } else if (kind == 2) {
            sos_load_tile<MODE>(R, tiles + (size_t)slot * TS, t * T, c0, Cw, ttid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // the tile has landed in its slot (all threads of the loading team call this)
    auto wait_load = [&](int64_t t, int slot, uint32_t parity) {
        const int kind = load_kind(t);
        if (kind == 1) {
            zp_mbar_wait(mbar + slot, parity);
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            zp_team_bar(team);
            if (ttid == 0) zp_mbar_arrive(mbar + slot);
        }
    };
    // the team's first tile is on its way
    issue_load(top - team, slot_of(top - team));

    // tile slot -> this thread's SOS_L samples
    auto read_tile = [&](int64_t t, int slot, double (&x)[SOS_L]) {
        const double* xp = tiles + (size_t)slot * TS + g * GS + cw;
        const bool xform = RECT && t * T >= P.edgeL && (t + 1) * (int64_t)T <= P.edgeL + P.nx;
        if (Cw == 8) {                                       // full group (block-uniform): no predicates
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) x[i] = xp[i * 8];
        } else if (!chan_ok) {
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) x[i] = 0.0;
        } else {
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) x[i] = xp[i * Cw];
        }
        if (xform) {
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) x[i] = HALF_PI * fabs(x[i]);
        }
    };

    // slots and mbarrier parities advance with the walk (no divisions in the loop): sl = slot of
    // the look-ahead tile, wr / wpar = walk position modulo NSLOT and the parity of its quotient
    int sl = slot_of(top - team), wr = team;
    uint32_t wpar = 0;
    auto advance = [&]() {
        sl -= NTEAM;
        if (sl < 0) sl += NSLOT;
        wr += NTEAM;
        if (wr >= NSLOT) { wr -= NSLOT; wpar ^= 1u; }
    };
    for (int64_t w = team; ; w += NTEAM, advance()) {
        const int64_t tl = top - w;                          // look-ahead tile of this iteration
        const int64_t tm = tl + JJ;                          // main tile of this iteration
        if (tm < a) break;
        const bool do_l = tl >= a - JJ;
        const bool do_m = tm <= top;
        int sm = sl + JJ;                                    // slot of the main tile (and of tile tl - NTEAM)
        if (sm >= NSLOT) sm -= NSLOT;
        // ---------------------------------------------------------------- L(tl)
        if (do_l || tl >= 0) wait_load(tl, sl, wpar);        // the team's tile tl has landed
        if (do_l) {
            const int slot = sl;
            if (tl < 0) {
                // before the sequence: the virtual tile -1 carries the initial state of the sweep
                if (ttid < CG * D) {
                    const int ch = ttid / D, d = ttid - ch * D;
                    double v = 0.0;
                    if (tl == -1 && P.zi_left && c0 + ch < C)
                        v = P.zi.z[d] * zp_ext_value<RECT>(P, P.src + c0 + ch, 0);
                    aggr[(size_t)slot * CG * D + ttid] = v;
                }
                zp_team_bar(team);
                if (ttid == 0) { __threadfence_block(); la_flag[slot] = tl; }
            } else {
                double x[SOS_L];
                read_tile(tl, slot, x);
                double v[D], ex[D];
#pragma unroll
                for (int d = 0; d < D; ++d) v[d] = 0.0;
                zp_pass_a<S, false>(K, x, v);
                {
                    int k = 0;
                    for (int off = CG; off < 32; off <<= 1, ++k) {
                        double u[D];
#pragma unroll
                        for (int d = 0; d < D; ++d) u[d] = __shfl_up_sync(0xffffffffu, v[d], off);
                        if (lane >= off) matvec_acc<D>(tab_s + k * DD, u, v);
                    }
                }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const double u = __shfl_up_sync(0xffffffffu, v[d], CG & 31);
                    ex[d] = gl == 0 ? 0.0 : u;
                }
                if (gl == GW - 1) {
#pragma unroll
                    for (int d = 0; d < D; ++d) wagg[(warp * CG + cw) * D + d] = v[d];
                }
                zp_team_bar(team);
                double pre[D];
#pragma unroll
                for (int d = 0; d < D; ++d) pre[d] = 0.0;
                for (int j = 0; j < warp; ++j) {
                    double u[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) u[d] = wagg[(j * CG + cw) * D + d];
                    matvec_acc<D>(tab_wpow + (warp - 1 - j) * DD, u, pre);
                }
                matvec_acc<D>(tab_fix + gl * DD, pre, ex);   // ex + A^(L gl) pre
#pragma unroll
                for (int d = 0; d < D; ++d) etot[((size_t)slot * D + d) * SOS_NT + ttid] = ex[d];
                if (ttid < CG * D) {
                    const int ch = ttid / D, r = ttid - ch * D;
                    double acc = 0.0;
                    for (int q = 0; q < SOS_NW; ++q) {
                        const double* M = tab_wpow + (SOS_NW - 1 - q) * DD + r * D;
                        const double* u = wagg + (q * CG + ch) * D;
                        for (int c = 0; c < D; ++c) acc = fma(M[c], u[c], acc);
                    }
                    aggr[(size_t)slot * CG * D + ttid] = acc;
                }
                zp_team_bar(team);
                if (ttid == 0) { __threadfence_block(); la_flag[slot] = tl; }
            }
        }
        // the team's next look-ahead tile: into the slot the main visit below frees
        const int64_t tn = tl - NTEAM;
        if (!do_m) {
            issue_load(tn, sm);
            continue;
        }
        // ---------------------------------------------------------------- M(tm)
        const int64_t t = tm;
        const int slot = sm;
        const bool store = t < b;
        const bool last = t == P.ntt - 1;
        zp_wait_le(la_flag + slot, t, lane);                 // parked by its look-ahead visit
        double x[SOS_L];
        read_tile(t, slot, x);
        double z[D];
#pragma unroll
        for (int d = 0; d < D; ++d) z[d] = etot[((size_t)slot * D + d) * SOS_NT + ttid];
        zp_team_bar(team);                                   // every thread of the team has its samples
        issue_load(tn, slot);
        // ---- forward state entering the tile: sum_{j=1..JJ} (A^T)^(j-1) agg[t-j]
        {
            double sf[D];
#pragma unroll
            for (int d = 0; d < D; ++d) sf[d] = 0.0;
            for (int j = 1; j <= JJ; ++j) {
                int sj = slot - j;
                if (sj < 0) sj += NSLOT;
                zp_wait_le(la_flag + sj, t - j, lane);
                double q[D], M[DD];
#pragma unroll
                for (int d = 0; d < D; ++d) q[d] = aggr[(size_t)sj * CG * D + cw * D + d];
#pragma unroll
                for (int e = 0; e < DD; ++e) M[e] = Pt[(size_t)(j - 1) * DD + e];
                matvec_acc<D>(M, q, sf);
            }
            double tmp[D];
#pragma unroll
            for (int d = 0; d < D; ++d) tmp[d] = 0.0;
            matvec_acc<D>(tab_wpow + warp * DD, sf, tmp);
            matvec_acc<D>(tab_fix + gl * DD, tmp, z);
        }
        zp_df2t<S, false>(K, x, z);                          // x <- y1
        // ---- backward sweep over the same registers
        double vb[D];
#pragma unroll
        for (int d = 0; d < D; ++d) vb[d] = 0.0;
        int lstar = SOS_L;                                   // last tile: index of the last row in its owner
        bool owner = false;
        double sb0[D];
#pragma unroll
        for (int d = 0; d < D; ++d) sb0[d] = 0.0;
        if (last) {
            const int il = (int)(P.N - 1 - t * T);
            const int gstar = il / SOS_L;
            lstar = il - gstar * SOS_L;
            owner = g == gstar;
            double ylast = 0.0;
            {
                double tmp[SOS_L];
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) tmp[i] = x[i];
#pragma unroll 1
                for (int i = 0; i < SOS_L; ++i) {
                    if (owner && i == lstar) ylast = tmp[i];
                    if (g > gstar || (owner && i > lstar)) tmp[i] = 0.0;
                }
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = tmp[i];
            }
            if (owner && P.zi_right) {
#pragma unroll
                for (int d = 0; d < D; ++d) sb0[d] = P.zi.z[d] * ylast;
            }
        }
        zp_pass_a<S, true>(K, x, vb);
        if (last && owner) {
            double h[D];
#pragma unroll
            for (int d = 0; d < D; ++d) h[d] = sb0[d];
            for (int k = 0; k <= lstar; ++k) zp_step0<S>(K, h);
#pragma unroll
            for (int d = 0; d < D; ++d) vb[d] += h[d];
        }
        {
            int k = 0;
            for (int off = CG; off < 32; off <<= 1, ++k) {
                double u[D];
#pragma unroll
                for (int d = 0; d < D; ++d) u[d] = __shfl_down_sync(0xffffffffu, vb[d], off);
                if (lane + off < 32) matvec_acc<D>(tab_s + k * DD, u, vb);
            }
        }
        double zb[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double u = __shfl_down_sync(0xffffffffu, vb[d], CG & 31);
            zb[d] = gl == GW - 1 ? 0.0 : u;
        }
        if (gl == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) waggb[(warp * CG + cw) * D + d] = vb[d];
        }
        zp_team_bar(team);
        // ---- the backward state entering this tile (from the team of tile t + 1)
        double sv[D];
        const int tprev = (team + NTEAM - 1) % NTEAM;         // team of the main visit of tile t + 1
        {
            if (t < top) zp_wait_le(sb_flag, t + 1, lane);
#pragma unroll
            for (int d = 0; d < D; ++d) sv[d] = (last || t == top) ? 0.0 : sinb_s[(size_t)tprev * CG * D + cw * D + d];
        }
        zp_team_bar(team);                                   // every warp of the team has read it
        if (warp == 0) {
            // hand the state on: sin_b(t) = A^T sin_b(t+1) + sum_w A^(L GW w) wagg[w]
            for (int e = lane; e < CG * D; e += 32) {
                const int ch = e / D, r = e - ch * D;
                double acc = 0.0;
                for (int q = 0; q < SOS_NW; ++q) {
                    const double* M = tab_wpow + q * DD + r * D;
                    const double* u = waggb + (q * CG + ch) * D;
                    for (int c = 0; c < D; ++c) acc = fma(M[c], u[c], acc);
                }
                if (!(last || t == top)) {
                    const double* M = Pt + DD + r * D;
                    const double* u = sinb_s + (size_t)tprev * CG * D + ch * D;
                    for (int c = 0; c < D; ++c) acc = fma(M[c], u[c], acc);
                }
                sinb_s[(size_t)team * CG * D + e] = acc;
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); *sb_flag = t; }
        }
        {
            double pre[D];
#pragma unroll
            for (int d = 0; d < D; ++d) pre[d] = 0.0;
            for (int j = SOS_NW - 1; j > warp; --j) {
                double u[D];
#pragma unroll
                for (int d = 0; d < D; ++d) u[d] = waggb[(j * CG + cw) * D + d];
                matvec_acc<D>(tab_wpow + (j - 1 - warp) * DD, u, pre);
            }
            matvec_acc<D>(tab_wpow + (SOS_NW - 1 - warp) * DD, sv, pre);
            matvec_acc<D>(tab_fix + (GW - 1 - gl) * DD, pre, zb);
        }
        if (last) {
            double tmp[SOS_L];
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) tmp[i] = x[i];
#pragma unroll 1
            for (int i = SOS_L - 1; i >= 0; --i) {
                if (owner && i == lstar) {
#pragma unroll
                    for (int d = 0; d < D; ++d) zb[d] = sb0[d];
                }
                double xv = tmp[i];
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    const double y = fma(K.coef[q][0], xv, zb[2 * q]);
                    zb[2 * q] = fma(K.coef[q][1], xv, zb[2 * q + 1]) - K.coef[q][3] * y;
                    zb[2 * q + 1] = K.coef[q][2] * xv - K.coef[q][4] * y;
                    xv = y;
                }
                tmp[i] = xv;
            }
#pragma unroll
            for (int i = 0; i < SOS_L; ++i) x[i] = tmp[i];
        } else {
            zp_df2t<S, true>(K, x, zb);                      // x <- y2
        }
        if (store && chan_ok) {
            if (P.clamp) {
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = 0.5 * (x[i] + fabs(x[i]));   // max(x, 0), two fp64 ops
            }
            const int64_t e0 = t * T + (int64_t)g * SOS_L;
            const bool fast = t * T >= P.out_first && (t + 1) * (int64_t)T <= P.out_first + P.n_dst;
            double* p = P.dst + (e0 - P.out_first) * C + c0 + cw;
            if (fast && C == 8) {
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) ADN_STORE(p + i * 8, x[i]);
            } else {
                const int64_t o0 = e0 - P.out_first;
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) {
                    if (fast || (o0 + i >= 0 && o0 + i < P.n_dst)) ADN_STORE(p, x[i]);
                    p += C;
                }
            }
        }
#ifdef ADN_ZP_TIMING
        if (ttid == 0 && run < 4 && t - a < 64 && t >= a) zp_tile_times[run * 64 + (t - a)] = zp_now();
#endif
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef ADN_ZP_TIMING
    if (ttid == 0) zp_times[zp_slot_s * 8 + 2 + team] = zp_now();
#endif
}

std::atomic<int64_t> g_zp_launches{0};
#ifdef ADN_ZP_TIMING
constexpr int ZP_SMEM_MAX = 226 * 1024;     // the timing build keeps a word of static shared memory
#else
constexpr int ZP_SMEM_MAX = 227 * 1024;
#endif

template <int S, bool RECT>
int32_t launch_zp(const SosPlan& plan, const ZpArgs& P, const SosRun& R, size_t smem, unsigned grid, int nteam,
                  cudaStream_t st) {
    SosK<S> K;
    fill_sosk<S>(plan, K);
    if (nteam == 4) {
        auto kern = sos_zp_park_kernel<S, RECT ? MODE_ENVF : MODE_ZPF, 4>;
        static bool attr_done = false;           // per instantiation
        if (!attr_done) {
            ADN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ZP_SMEM_MAX));
            attr_done = true;
        }
        kern<<<grid, SOS_NT * 4, smem, st>>>(K, P, R);
    } else if (nteam == 3) {
        auto kern = sos_zp_park_kernel<S, RECT ? MODE_ENVF : MODE_ZPF, 3>;
        static bool attr_done = false;
        if (!attr_done) {
            ADN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ZP_SMEM_MAX));
            attr_done = true;
        }
        kern<<<grid, SOS_NT * 3, smem, st>>>(K, P, R);
    } else {
        auto kern = sos_zp_kernel<S, RECT>;
        static bool attr_done = false;
        if (!attr_done) {
            ADN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            attr_done = true;
        }
        kern<<<grid, SOS_NT, smem, st>>>(K, P);
    }
    count_launch();
    g_zp_launches.fetch_add(1);
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

template <int S>
int32_t launch_zp_S(bool rect, const SosPlan& plan, const ZpArgs& P, const SosRun& R, size_t smem,
                    unsigned grid, int nteam, cudaStream_t st) {
    return rect ? launch_zp<S, true>(plan, P, R, smem, grid, nteam, st)
                : launch_zp<S, false>(plan, P, R, smem, grid, nteam, st);
}

int zp_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

}  // namespace

int64_t zp_launches() { return g_zp_launches.load(); }

#ifdef ADN_ZP_TIMING
extern "C" int adn_debug_zp_times(unsigned long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, zp_times, sizeof(unsigned long long) * (size_t)n);
}
extern "C" int adn_debug_zp_tile_times(unsigned long long* out) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, zp_tile_times, sizeof(unsigned long long) * 4 * 64);
}
#endif

int32_t zero_phase_regs_dev(bool rect, const double* sos, int32_t S, const double* src, int64_t n_src,
                            int32_t C, int32_t edge_left, int32_t edge_right, int64_t out_first,
                            double* dst, int64_t n_dst, int32_t clamp_negative, bool* handled,
                            cudaStream_t st) {
    *handled = false;
    if (!option(ADN_OPT_ZERO_PHASE_ONEPASS) || S < 1 || S > 4 || n_dst <= 0) return ADN_OK;
    const int D = 2 * S;
    // Channels per group.  A tile is 4096 / CG rows long, and the pipelined kernel holds NTEAM + JJ
    // tiles in shared memory, JJ = the tiles the cascade needs to forget: a slow cascade (500 Hz at
    // 250 kHz: 4400 rows) gets narrower groups = longer tiles until that fits (16-byte row segments
    // at least, as long as the channel count is even).  Measured on B200, 500-Hz envelope at 250 kHz:
    // 8 ch (groups of 2) 0.236 against 0.441 ms for 32 M samples; 64 ch 0.552 against 0.506 ms -- the
    // 16-byte segments of 512-byte rows are scattered over too many lines, so only rows of at most
    // one 128-byte line take narrower groups.
    int CG = pick_cg(C);
    std::shared_ptr<SosPlan> plan;
    int32_t rc = get_sos_plan(sos, S, CG, st, &plan);
    if (rc) return rc;
    auto pipe_need = [&](const SosPlan& pl, int cg, int nt) {
        const size_t TS = (size_t)(SOS_NT / cg) * (SOS_L * cg + (cg < 16 ? cg : 0));
        const int ns = nt + pl.jzp;
        return ((size_t)pl.n_staged * D * D + (size_t)nt * 2 * SOS_NW * cg * D + (size_t)ns * cg * D +
                (size_t)nt * cg * D + (size_t)ns * D * SOS_NT + (size_t)(2 * ns + 2) +
                (size_t)(pl.jzp + 2) * D * D + (size_t)ns * TS) * 8;
    };
    if (zp_env("ADN_ZP_PIPE", 1) && zp_env("ADN_ZP_NARROW", 1) && C <= 16 && pipe_need(*plan, CG, 3) > 227 * 1024) {
        const int cg_min = C % 2 == 0 ? 2 : 1;
        for (int cg = CG / 2; cg >= cg_min; cg /= 2) {
            std::shared_ptr<SosPlan> pl;
            if ((rc = get_sos_plan(sos, S, cg, st, &pl))) return rc;
            if (pipe_need(*pl, cg, 3) <= 227 * 1024) { CG = cg; plan = pl; break; }
        }
    }
    if (plan->jzp > ZP_MAX_JJ) return ADN_OK;               // forgets too slowly: two sweeps with look-back
    ZpArgs P;
    memset(&P, 0, sizeof P);
    P.src = src; P.dst = dst; P.tab = plan->dtab;
    P.off_fix = plan->off_fix; P.off_wpow = plan->off_wpow; P.off_tile = plan->off_tile;
    P.n_staged = plan->n_staged;
    P.nx = n_src;
    P.edgeL = edge_left; P.edgeR = edge_right;
    P.N = n_src + edge_left + edge_right;
    P.out_first = out_first; P.n_dst = n_dst;
    P.C = C; P.CG = CG; P.ngroups = (C + CG - 1) / CG;
    P.T = (SOS_NT / CG) * SOS_L;
    P.ntt = (P.N + P.T - 1) / P.T;
    P.t_out0 = out_first / P.T;
    P.t_out1 = (out_first + n_dst - 1) / P.T + 1;
    P.zi_left = edge_left > 0; P.zi_right = edge_right > 0;
    P.clamp = clamp_negative ? 1 : 0;
    P.JJ = plan->jzp;
    P.JO = plan->jzp;
    P.pf = zp_env("ADN_ZP_PREFETCH", 1);
#ifdef ADN_ZP_TIMING
    P.pf = (P.pf & 1) | (zp_env("ADN_ZP_ROT", 0) << 8);
#endif
    sosfilt_zi_host(sos, S, P.zi.z);
    const int64_t out_tiles = P.t_out1 - P.t_out0;
    // the pipelined kernel (one block of NTEAM teams per SM, tiles parked in shared memory) when
    // NTEAM + JJ tile slots fit; else every tile is read twice (the second time from L2)
    SosRun R;
    memset(&R, 0, sizeof R);
    R.src = src; R.tab = plan->dtab;
    R.n = P.N; R.nx = n_src; R.edge = edge_left;
    R.C = C; R.CG = CG; R.ngroups = P.ngroups; R.T = P.T; R.ntt = P.ntt;
    {
        const bool even = (C % 2 == 0) && (CG % 2 == 0);
        R.vec_in = even && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        R.lc = 0;
        while ((1 << R.lc) < CG) ++R.lc;
        if (C % 2 && CG > 1) R.lc = -1;
    }
    // rows of a full group are contiguous (C == CG) and the padded sub-chunks 16-byte aligned
    // (measured on B200: pays for 64-byte rows, 152 against 164 us; not for narrower groups)
    P.bulk_ok = (C == CG && CG >= 8 && zp_env("ADN_ZP_TMA", 1)) ? 1 : 0;
    int nteam = 0;
    size_t smem = 0;
    int64_t resident;
    if (zp_env("ADN_ZP_PIPE", 1)) {
        // measured on B200 (8 ch x 48 kHz, 80 s): three teams at 168 registers 141 us, four at 128
        // registers (some spills) 149 us for one section; two sections are even (204 / 200 us)
        const int want = zp_env("ADN_ZP_NTEAM", S == 1 ? 3 : ZP_NTEAM_MAX);
        for (int nt = want < ZP_NTEAM_MAX ? want : ZP_NTEAM_MAX; nt >= 3; --nt) {
            const size_t need = pipe_need(*plan, CG, nt);
            if (need <= 227 * 1024) { nteam = nt; smem = need; break; }
        }
    }
    if (nteam > 0) {
        resident = ctx().sm_count;
    } else {
        smem = ((size_t)plan->n_staged * D * D + (size_t)(SOS_NW + 3) * CG * D +
                (size_t)(P.JJ + 1) * (CG * D + SOS_NT * D)) * 8;
        if (smem > 96 * 1024) return ADN_OK;
        int bps = (int)((220 * 1024) / (smem + 1024));
        const int bmax = S <= 2 ? ZP_BLOCKS : 3;
        if (bps > bmax) bps = bmax;
        resident = (int64_t)ctx().sm_count * bps;
    }
    int64_t runs = resident / P.ngroups;
    if (runs < 1) runs = 1;
    // the runs at the two ends of the sequence build their extension tile element by element with
    // blocking loads (measured on B200: some 10 us, five tiles' worth, at the end of the first
    // run's walk): those two runs are that much shorter, so that all blocks end together
    int64_t short0 = 0, short1 = 0;
    if (nteam > 0 && zp_env("ADN_ZP_SHORT_EDGE", 1)) {
        if (edge_left > 0 && P.t_out0 == 0) short0 = 5;
        if (edge_right > 0 && P.t_out1 == P.ntt) short1 = 5;
    }
    int64_t run_tiles = (out_tiles + short0 + short1 + runs - 1) / runs;
    // the run-out and the look-ahead may cost a third of a run at most
    const int64_t min_run = 2 * (int64_t)(P.JJ + P.JO);
    if (run_tiles < min_run) run_tiles = min_run;
    if (run_tiles < 4 * (short0 + short1)) short0 = short1 = 0;
    if (short0 + short1 == 0) run_tiles = (out_tiles + runs - 1) / runs < min_run ? min_run : (out_tiles + runs - 1) / runs;
    P.short0 = (int32_t)short0;
    runs = (out_tiles + short0 + run_tiles - 1) / run_tiles;
    if (out_tiles < 2 * min_run) return ADN_OK;              // short input: the two sweeps
    if (run_tiles > 0x3fffffff || runs * P.ngroups > 0x7fffffff) return ADN_OK;
    P.run_tiles = (int32_t)run_tiles;
    const unsigned grid = (unsigned)(runs * P.ngroups);
    switch (S) {
        case 1: rc = launch_zp_S<1>(rect, *plan, P, R, smem, grid, nteam, st); break;
        case 2: rc = launch_zp_S<2>(rect, *plan, P, R, smem, grid, nteam, st); break;
        case 3: rc = launch_zp_S<3>(rect, *plan, P, R, smem, grid, nteam, st); break;
        default: rc = launch_zp_S<4>(rect, *plan, P, R, smem, grid, nteam, st); break;
    }
    if (rc == ADN_OK) *handled = true;
    return rc;
}

}  // namespace adn
