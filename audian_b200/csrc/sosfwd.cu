// Forward SOS filtering (scipy sosfilt: BufferedFilter.process, src/audian/bufferedfilter.py:31-36)
// of long traces on the pipelined skeleton of zerophase.cu: a block is NTEAM teams of four warps
// and owns one run of consecutive tiles, walking forward along time; team k takes the tiles
// t0 + k, t0 + k + NTEAM, ...  A tile is staged in the team's shared-memory slot (TMA bulk copies
// for 64-byte rows, cp.async granules otherwise) one iteration ahead, read into registers (one
// thread = SOS_L consecutive samples of one channel), and the slot is handed to the next
// prefetch at once.  In registers: pass A (zero-state end state of the thread's samples), the
// scans inside the tile, and the exact DF2T recurrence from the true incoming state; the
// results are stored from the registers (a warp store covers GW rows x CG channels = whole
// 32-byte sectors for CG >= 4).  The state entering a tile is handed from tile to tile as
//     s(t + 1) = A^T s(t) + agg(t)
// by one warp as soon as the tile aggregate agg(t) is known: only that small matrix-vector
// product is serial along the run, the teams overlap freely.  A run that does not start at row 0
// runs over `pre` tiles before it from zero state without storing (the cascade has forgotten
// its state by then: the criterion of sosfilt.cu's run kernel, which this kernel replaces).
#include "sos_common.cuh"
#include <cstring>
#include <cstdlib>
#include <atomic>

namespace adn {

namespace {

struct FwdArgs {
    const double* src;
    double* dst;
    const double* tab;                 // tables of the first sub-cascade
    const double* tab2;                // ... of the second one (3 and 4 sections: 2 + 1, 2 + 2)
    int32_t off_fix, off_wpow, off_tile, n_staged;
    const double* s0;                  // [C][D] initial state or null (D = 2 x all sections)
    double* zf;                        // [C][D] final state or null
    int64_t n;                         // rows of the source
    int64_t out_skip, n_dst;           // source rows [out_skip, out_skip + n_dst) -> dst rows
    int64_t ntt;                       // tiles of the source
    int64_t t_out0, t_out1;            // tiles that hold output rows
    int32_t C, CG, ngroups, T;
    int32_t pre;                       // run-in tiles
    int32_t run_tiles;
    int32_t bulk_ok;
    // full-trace min/max of the same pass (compresseddata.py:49-52): per-tile partial (min, max)
    // of the raw rows / of the filtered rows, [tile][C][2]; folded per segment by mm_fold_kernel
    double* mm_raw;
    double* mm_filt;
};

// ---- min / max with numpy's ordered rule (axis 0 of a 2-D array): a NaN wins and stays (the
// latest one), equal values resolve to the later row (signed zeros).  `a` comes before `b`.
__device__ __forceinline__ double mm_min2(double a, double b) {
    if (b != b) return b;
    if (a != a) return a;
    return a < b ? a : b;
}
__device__ __forceinline__ double mm_max2(double a, double b) {
    if (b != b) return b;
    if (a != a) return a;
    return a > b ? a : b;
}

// (min, max) of the nv <= SOS_L samples a thread holds, in time order
__device__ __forceinline__ void mm_thread(const double (&x)[SOS_L], int nv, double& mn, double& mx) {
    if (nv == SOS_L) {
        // plain compare-and-select (later wins ties); NaNs and infinities are spotted on the
        // exponent bits and sent through the exact rule
        double m = x[0], M = x[0];
        int top = __double2hiint(x[0]) & 0x7fffffff;
#pragma unroll
        for (int i = 1; i < SOS_L; ++i) {
            m = m < x[i] ? m : x[i];
            M = M > x[i] ? M : x[i];
            top = max(top, __double2hiint(x[i]) & 0x7fffffff);
        }
        if (top < 0x7ff00000) { mn = m; mx = M; return; }
    }
    double tmp[SOS_L];
#pragma unroll
    for (int i = 0; i < SOS_L; ++i) tmp[i] = x[i];
    double m = tmp[0], M = tmp[0];
#pragma unroll 1
    for (int i = 1; i < nv; ++i) { m = mm_min2(m, tmp[i]); M = mm_max2(M, tmp[i]); }
    mn = m; mx = M;
}

struct FwdLane;
// fold the threads of a team over time (sub-chunks in order) and write the tile's partial
__device__ __forceinline__ void mm_tile(double mn, double mx, bool valid, int lane, int warp, int gl, int cw,
                                        int CG, int GW, int team, bool chan_ok, double* wmm, double* out) {
    // inside the warp: ordered tree over gl (lanes CG apart); a lane without valid rows passes
    // its later neighbour's value on unchanged
    for (int off = CG; off < 32; off <<= 1) {
        const double bn = __shfl_down_sync(0xffffffffu, mn, off);
        const double bx = __shfl_down_sync(0xffffffffu, mx, off);
        const int bv = __shfl_down_sync(0xffffffffu, (int)valid, off);
        if (lane + off < 32 && bv) {
            mn = valid ? mm_min2(mn, bn) : bn;
            mx = valid ? mm_max2(mx, bx) : bx;
            valid = true;
        }
    }
    if (gl == 0) {
        wmm[(warp * CG + cw) * 3 + 0] = mn;
        wmm[(warp * CG + cw) * 3 + 1] = mx;
        wmm[(warp * CG + cw) * 3 + 2] = valid ? 1.0 : 0.0;
    }
    zp_team_bar(team);
    if (warp == 0 && gl == 0 && chan_ok) {
        double m = 0.0, M = 0.0;
        bool v = false;
        for (int q = 0; q < SOS_NW; ++q) {
            if (wmm[(q * CG + cw) * 3 + 2] != 0.0) {
                const double bn = wmm[(q * CG + cw) * 3 + 0], bx = wmm[(q * CG + cw) * 3 + 1];
                m = v ? mm_min2(m, bn) : bn;
                M = v ? mm_max2(M, bx) : bx;
                v = true;
            }
        }
        out[0] = m;
        out[1] = M;
    }
    zp_team_bar(team);                                       // wmm is free again
}

// rows 2 j / 2 j + 1 of dst = fold of the partials of the tiles of segment j, in time order
__global__ void __launch_bounds__(32)
mm_fold_kernel(const double* __restrict__ part, int64_t ntiles, int32_t C, int64_t tiles_per_seg,
               double* __restrict__ dst) {
    const int64_t seg = blockIdx.x / C;
    const int c = (int)(blockIdx.x % C);
    const int lane = threadIdx.x;
    const int64_t t0 = seg * tiles_per_seg, t1 = min(ntiles, t0 + tiles_per_seg);
    const int64_t per = (t1 - t0 + 31) / 32;
    const int64_t a = t0 + lane * per, b = min(t1, a + per);
    double m = 0.0, M = 0.0;
    bool v = false;
    for (int64_t t = a; t < b; ++t) {
        const double bn = part[(t * C + c) * 2], bx = part[(t * C + c) * 2 + 1];
        m = v ? mm_min2(m, bn) : bn;
        M = v ? mm_max2(M, bx) : bx;
        v = true;
    }
    for (int off = 1; off < 32; off <<= 1) {
        const double bn = __shfl_down_sync(0xffffffffu, m, off);
        const double bx = __shfl_down_sync(0xffffffffu, M, off);
        const int bv = __shfl_down_sync(0xffffffffu, (int)v, off);
        if (lane + off < 32 && bv) {
            m = v ? mm_min2(m, bn) : bn;
            M = v ? mm_max2(M, bx) : bx;
            v = true;
        }
    }
    if (lane == 0) {
        dst[(2 * seg) * C + c] = m;
        dst[(2 * seg + 1) * C + c] = M;
    }
}

// shared-memory state of one sub-cascade (stage) of the kernel
template <int S>
struct FwdStage {
    const double* tab_s;               // staged scan / fix / wpow tables
    const double* tab_fix;
    const double* tab_wpow;
    const double* pt_s;                // A^T (one tile)
    double* wagg;                      // this team's [NW][CG][D]
    double* sin_s;                     // [NTEAM][CG][D]: state entering a team's tile
    volatile long long* flag;          // tile whose incoming state is published
};

struct FwdLane {                       // where a thread sits
    int lane, warp, team, nteam, CG, GW, gl, cw;
    bool chan_ok;
};

// One sub-cascade over the SOS_L samples a thread holds (in place): pass A, the scans inside the
// tile, the hand-over of the state along the run, the exact recurrence.  zf_out (or null): where
// the state right after sample `ilast` goes (the sub-chunk that holds the last row).
template <int S>
__device__ __forceinline__ void fwd_stage(const SosK<S>& K, const FwdStage<S>& st, const FwdLane& L,
                                          double (&x)[SOS_L], int64_t t, double* zf_out, int ilast) {
    constexpr int D = 2 * S;
    constexpr int DD = D * D;
    const int lane = L.lane, warp = L.warp, team = L.team, CG = L.CG, GW = L.GW, gl = L.gl, cw = L.cw;
    double z[D];
    {
        double v[D];
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = 0.0;
        zp_pass_a<S, false>(K, x, v);
        {
            int k = 0;
            for (int off = CG; off < 32; off <<= 1, ++k) {
                double u[D];
#pragma unroll
                for (int d = 0; d < D; ++d) u[d] = __shfl_up_sync(0xffffffffu, v[d], off);
                if (lane >= off) matvec_acc<D>(st.tab_s + k * DD, u, v);
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double u = __shfl_up_sync(0xffffffffu, v[d], CG & 31);
            z[d] = gl == 0 ? 0.0 : u;
        }
        if (gl == GW - 1) {
#pragma unroll
            for (int d = 0; d < D; ++d) st.wagg[(warp * CG + cw) * D + d] = v[d];
        }
        zp_team_bar(team);
        double pre[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pre[d] = 0.0;
        for (int j = 0; j < warp; ++j) {
            double u[D];
#pragma unroll
            for (int d = 0; d < D; ++d) u[d] = st.wagg[(j * CG + cw) * D + d];
            matvec_acc<D>(st.tab_wpow + (warp - 1 - j) * DD, u, pre);
        }
        matvec_acc<D>(st.tab_fix + gl * DD, pre, z);         // z = ex + A^(L gl) pre
    }
    // ---- the state entering this tile (from the team of tile t - 1), and on to tile t + 1
    double sv[D];
    zp_wait_ge(st.flag, t, lane);
#pragma unroll
    for (int d = 0; d < D; ++d) sv[d] = st.sin_s[(size_t)team * CG * D + cw * D + d];
    zp_team_bar(team);                                       // every warp of the team has read it
    if (warp == 0) {
        const int tnext = (team + 1) % L.nteam;
        for (int e = lane; e < CG * D; e += 32) {
            const int ch = e / D, r = e - ch * D;
            double acc = 0.0;
            for (int q = 0; q < SOS_NW; ++q) {
                const double* M = st.tab_wpow + (SOS_NW - 1 - q) * DD + r * D;
                const double* u = st.wagg + (q * CG + ch) * D;
                for (int c = 0; c < D; ++c) acc = fma(M[c], u[c], acc);
            }
            const double* M = st.pt_s + r * D;
            const double* u = st.sin_s + (size_t)team * CG * D + ch * D;
            for (int c = 0; c < D; ++c) acc = fma(M[c], u[c], acc);
            st.sin_s[(size_t)tnext * CG * D + e] = acc;
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); *st.flag = t + 1; }
    }
    {
        double tmp[D];
#pragma unroll
        for (int d = 0; d < D; ++d) tmp[d] = 0.0;
        matvec_acc<D>(st.tab_wpow + warp * DD, sv, tmp);
        matvec_acc<D>(st.tab_fix + gl * DD, tmp, z);
    }
    // ---- the exact recurrence from the true incoming state, in registers
    if (zf_out != nullptr) {
        // the sub-chunk that holds the last row: rolled loop, state captured right after it
        double tmp[SOS_L];
#pragma unroll
        for (int i = 0; i < SOS_L; ++i) tmp[i] = x[i];
#pragma unroll 1
        for (int i = 0; i < SOS_L; ++i) {
            double xv = tmp[i];
#pragma unroll
            for (int q = 0; q < S; ++q) {
                const double y = fma(K.coef[q][0], xv, z[2 * q]);
                z[2 * q] = fma(K.coef[q][1], xv, z[2 * q + 1]) - K.coef[q][3] * y;
                z[2 * q + 1] = K.coef[q][2] * xv - K.coef[q][4] * y;
                xv = y;
            }
            tmp[i] = xv;
            if (i == ilast && L.chan_ok) {
#pragma unroll
                for (int d = 0; d < D; ++d) zf_out[d] = z[d];
            }
        }
#pragma unroll
        for (int i = 0; i < SOS_L; ++i) x[i] = tmp[i];
    } else {
        zp_df2t<S, false>(K, x, z);
    }
}

// S1 sections, then S2 more (S2 == 0: one stage): the cascade of S1 + S2 sections is the second
// sub-cascade applied to the output of the first, sample by sample the same arithmetic
template <int S1, int S2, int NTC>
__global__ void __launch_bounds__(SOS_NT * NTC, 1)
sos_fwd_park_kernel(const __grid_constant__ SosK<S1> K1, const __grid_constant__ SosK<(S2 > 0 ? S2 : 1)> K2,
                    const __grid_constant__ FwdArgs P, const __grid_constant__ SosRun R) {
    constexpr int D1 = 2 * S1, D2 = 2 * S2, DT = D1 + D2;
    constexpr int DD1 = D1 * D1, DD2 = D2 * D2;
    constexpr int NTEAM = NTC;
    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x, lane = tid & 31;
    const int team = tid >> 7, ttid = tid & (SOS_NT - 1), warp = ttid >> 5;
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
    const int pad = CG < 16 ? CG : 0;
    const int GS = SOS_L * Cw + pad;
    const int G = SOS_NT / CG;
    const size_t TS = (size_t)G * (SOS_L * CG + pad);

    // shared memory: per stage [tables | A^T | wagg | sin_s], then flags, mbarriers, tile slots
    double* p = smem;
    double* tab1_s = p;  p += (size_t)P.n_staged * DD1;
    double* pt1_s = p;   p += DD1;
    double* wagg1 = p + (size_t)team * SOS_NW * CG * D1;  p += (size_t)NTEAM * SOS_NW * CG * D1;
    double* sin1_s = p;  p += (size_t)NTEAM * CG * D1;
    double* tab2_s = p;  p += (size_t)P.n_staged * DD2;
    double* pt2_s = p;   p += DD2;
    double* wagg2 = p + (size_t)team * SOS_NW * CG * D2;  p += (size_t)NTEAM * SOS_NW * CG * D2;
    double* sin2_s = p;  p += (size_t)NTEAM * CG * D2;
    double* wmm = p + (size_t)team * SOS_NW * CG * 3;  p += (size_t)NTEAM * SOS_NW * CG * 3 + ((NTEAM * SOS_NW * CG * 3) & 1);
    volatile long long* sf_flag = reinterpret_cast<volatile long long*>(p);        // [2]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(const_cast<long long*>(sf_flag) + 2);      // [NTEAM]
    double* tiles = reinterpret_cast<double*>(mbar + NTEAM + (NTEAM & 1));          // [NTEAM][TS]
    double* slot_s = tiles + (size_t)team * TS;

    const int64_t a = P.t_out0 + run * P.run_tiles;         // output tiles [a, b)
    const int64_t b = min(a + (int64_t)P.run_tiles, P.t_out1);
    if (a >= b) return;
    const int64_t t_first = max((int64_t)0, a - P.pre);     // first tile of the walk

    for (int q = tid; q < P.n_staged * DD1; q += blockDim.x) tab1_s[q] = __ldg(P.tab + q);
    for (int q = tid; q < DD1; q += blockDim.x) pt1_s[q] = __ldg(P.tab + (size_t)(P.off_tile + 1) * DD1 + q);
    if (S2 > 0) {
        for (int q = tid; q < P.n_staged * DD2; q += blockDim.x) tab2_s[q] = __ldg(P.tab2 + q);
        for (int q = tid; q < DD2; q += blockDim.x) pt2_s[q] = __ldg(P.tab2 + (size_t)(P.off_tile + 1) * DD2 + q);
    }
    // state entering the first tile of the walk: the initial state at row 0, else zero (run-in);
    // team 0 takes the first tile
    for (int q = tid; q < CG * DT; q += blockDim.x) {
        const int ch = q / DT, d = q - ch * DT;
        double v = 0.0;
        if (t_first == 0 && P.s0 && c0 + ch < C) v = __ldg(P.s0 + (size_t)(c0 + ch) * DT + d);
        if (d < D1) sin1_s[ch * D1 + d] = v;
        else sin2_s[ch * D2 + (d - D1)] = v;
    }
    if (tid < NTEAM) zp_mbar_init(mbar + tid, 1);
    if (tid < 2) sf_flag[tid] = t_first;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    FwdLane L;
    L.lane = lane; L.warp = warp; L.team = team; L.nteam = NTEAM; L.CG = CG; L.GW = GW; L.gl = gl; L.cw = cw;
    L.chan_ok = chan_ok;
    FwdStage<S1> st1;
    st1.tab_s = tab1_s; st1.tab_fix = tab1_s + P.off_fix * DD1; st1.tab_wpow = tab1_s + P.off_wpow * DD1;
    st1.pt_s = pt1_s; st1.wagg = wagg1; st1.sin_s = sin1_s; st1.flag = sf_flag;
    FwdStage<(S2 > 0 ? S2 : 1)> st2;
    st2.tab_s = tab2_s; st2.tab_fix = tab2_s + P.off_fix * DD2; st2.tab_wpow = tab2_s + P.off_wpow * DD2;
    st2.pt_s = pt2_s; st2.wagg = wagg2; st2.sin_s = sin2_s; st2.flag = sf_flag + 1;

    const bool bulk_group = P.bulk_ok && Cw == CG;
    auto load_kind = [&](int64_t t) {
        if (t >= b) return 0;
        if (!bulk_group || (t + 1) * (int64_t)T > P.n) return 2;
        const uintptr_t g0 = reinterpret_cast<uintptr_t>(P.src + t * T * C + c0);
        return (g0 & 15) == 0 ? 1 : 2;
    };
    auto issue_load = [&](int64_t t) {
        const int kind = load_kind(t);
        if (kind == 1) {
            if (warp == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (lane == 0) zp_mbar_expect_tx(mbar + team, (uint32_t)(T * C * 8));
                __syncwarp();
                const double* gsrc = P.src + t * T * C + c0;
                for (int q = lane; q < G; q += 32)
                    zp_bulk_g2s(slot_s + (size_t)q * GS, gsrc + (size_t)q * SOS_L * C, (uint32_t)(SOS_L * C * 8), mbar + team);
            }
        } else if (kind == 2) {
            sos_load_tile<MODE_FWD>(R, slot_s, t * T, c0, Cw, ttid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    issue_load(t_first + team);
    int64_t use = 0;                                        // uses of this team's slot so far
    for (int64_t t = t_first + team; t < b; t += NTEAM, ++use) {
        // ---- the tile has landed in the team's slot
        if (load_kind(t) == 1) {
            zp_mbar_wait(mbar + team, (uint32_t)(use & 1));
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            zp_team_bar(team);
            if (ttid == 0) zp_mbar_arrive(mbar + team);
        }
        double x[SOS_L];
        {
            const double* xp = slot_s + g * GS + cw;
            if (Cw == 8) {                                   // full group (block-uniform): no predicates
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = xp[i * 8];
            } else if (!chan_ok) {
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = 0.0;
            } else {
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) x[i] = xp[i * Cw];
            }
        }
        zp_team_bar(team);                                   // every thread of the team has its samples
        issue_load(t + NTEAM);                               // the slot goes to the team's next tile
        const int64_t e0 = t * T + (int64_t)g * SOS_L;       // source row of x[0]
        const int64_t lastrow = P.n - 1;
        const bool capture = P.zf != nullptr && lastrow >= e0 && lastrow < e0 + SOS_L && t >= a;
        double* zf_c = capture ? P.zf + (size_t)(c0 + (chan_ok ? cw : 0)) * DT : nullptr;
        // rows of this thread that exist (the last tile is zero-filled behind the recording)
        const int nv = (int)max((int64_t)0, min((int64_t)SOS_L, P.n - e0));
        if (P.mm_raw != nullptr && t >= a) {
            double mn = 0.0, mx = 0.0;
            if (nv > 0) mm_thread(x, nv, mn, mx);
            mm_tile(mn, mx, nv > 0, lane, warp, gl, cw, CG, GW, team, chan_ok, wmm,
                    P.mm_raw + ((size_t)t * C + c0 + (chan_ok ? cw : 0)) * 2);
        }
        fwd_stage<S1>(K1, st1, L, x, t, zf_c, (int)(lastrow - e0));
        if (S2 > 0)
            fwd_stage<(S2 > 0 ? S2 : 1)>(K2, st2, L, x, t, capture ? zf_c + D1 : nullptr, (int)(lastrow - e0));
        if (P.mm_filt != nullptr && t >= a) {
            double mn = 0.0, mx = 0.0;
            if (nv > 0) mm_thread(x, nv, mn, mx);
            mm_tile(mn, mx, nv > 0, lane, warp, gl, cw, CG, GW, team, chan_ok, wmm,
                    P.mm_filt + ((size_t)t * C + c0 + (chan_ok ? cw : 0)) * 2);
        }
        if (t >= a && chan_ok && P.dst != nullptr) {
            const bool fast = t * T >= P.out_skip && (t + 1) * (int64_t)T <= P.out_skip + P.n_dst;
            double* q = P.dst + (e0 - P.out_skip) * C + c0 + cw;
            if (fast && C == 8) {
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) ADN_STORE(q + i * 8, x[i]);
            } else {
                const int64_t o0 = e0 - P.out_skip;
#pragma unroll
                for (int i = 0; i < SOS_L; ++i) {
                    if (fast || (o0 + i >= 0 && o0 + i < P.n_dst)) ADN_STORE(q, x[i]);
                    q += C;
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

std::atomic<int64_t> g_fwd_launches{0};

template <int S1, int S2, int NTC>
int32_t launch_fwd(const SosPlan& plan1, const SosPlan* plan2, const FwdArgs& P, const SosRun& R, size_t smem,
                   unsigned grid, cudaStream_t st) {
    SosK<S1> K1;
    fill_sosk<S1>(plan1, K1);
    SosK<(S2 > 0 ? S2 : 1)> K2;
    memset(&K2, 0, sizeof K2);
    if (S2 > 0) fill_sosk<(S2 > 0 ? S2 : 1)>(*plan2, K2);
    auto kern = sos_fwd_park_kernel<S1, S2, NTC>;
    static bool attr_done = false;               // per instantiation
    if (!attr_done) {
        ADN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    kern<<<grid, SOS_NT * NTC, smem, st>>>(K1, K2, P, R);
    count_launch();
    g_fwd_launches.fetch_add(1);
    ADN_CK(cudaGetLastError());
    return ADN_OK;
}

int fwd_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

}  // namespace

int64_t fwd_park_launches() { return g_fwd_launches.load(); }

// dst = sosfilt(sos, src)[out_skip:][:n_dst] from state s0 (or zero); zf = state after the last row.
// *handled = false (nothing launched): the cascade forgets too slowly or the trace is too short.
int32_t sosfilt_park_dev(const double* sos, int32_t S, const double* src, int64_t n, int32_t C,
                         int64_t out_skip, double* dst, int64_t n_dst, const double* s0, double* zf,
                         bool* handled, cudaStream_t st) {
    return sosfilt_minmax_park_dev(sos, S, src, n, C, out_skip, dst, n_dst, s0, zf, 0, nullptr, nullptr,
                                   handled, st);
}

// the same with the full-trace min/max rows of the raw source rows (mm_raw, 2 ceil(n / step) x C)
// and / or of the filtered rows (mm_filt) computed in the same pass: needs out_skip == 0,
// n_dst == n, whole tiles per segment (step a multiple of the tile's rows) and C >= 2
int32_t sosfilt_minmax_park_dev(const double* sos, int32_t S, const double* src, int64_t n, int32_t C,
                                int64_t out_skip, double* dst, int64_t n_dst, const double* s0, double* zf,
                                int64_t mm_step, double* mm_raw, double* mm_filt, bool* handled,
                                cudaStream_t st) {
    *handled = false;
    const bool want_mm = mm_step > 0 && (mm_raw || mm_filt);
    // measured on B200 (80 s of 8 ch x 48 kHz): one or two sections 96 / 90 us against 97 / 103 us of
    // the run kernel (64 ch: 739 against 779 us); three and four sections run as two chained
    // sub-cascades of at most two sections (all eight states at once are register bound: 187 us)
    const int smax = fwd_env("ADN_SOS_PARK_SMAX", 4);
    if (!fwd_env("ADN_SOS_PARK", 1) || S < 1 || S > 4 || S > smax || dst == nullptr || n_dst <= 0) return ADN_OK;
    const int CG = pick_cg(C);
    const int S1 = S <= 2 ? S : 2, S2 = S - S1;
    std::shared_ptr<SosPlan> plan, plan2;
    int32_t rc = get_sos_plan(sos, S1, CG, st, &plan);
    if (rc) return rc;
    if (S2 > 0 && (rc = get_sos_plan(sos + 6 * S1, S2, CG, st, &plan2))) return rc;
    if (plan->jpre > SOS_LOOK || (S2 > 0 && plan2->jpre > SOS_LOOK)) return ADN_OK;
    const int D1 = 2 * S1, D2 = 2 * S2;
    FwdArgs P;
    memset(&P, 0, sizeof P);
    P.src = src; P.dst = dst; P.tab = plan->dtab; P.tab2 = S2 > 0 ? plan2->dtab : nullptr; P.s0 = s0; P.zf = zf;
    P.off_fix = plan->off_fix; P.off_wpow = plan->off_wpow; P.off_tile = plan->off_tile;
    P.n_staged = plan->n_staged;
    P.n = n; P.out_skip = out_skip; P.n_dst = n_dst;
    P.C = C; P.CG = CG; P.ngroups = (C + CG - 1) / CG;
    P.T = (SOS_NT / CG) * SOS_L;
    P.ntt = (n + P.T - 1) / P.T;
    P.t_out0 = out_skip / P.T;
    if (want_mm) {
        if (out_skip != 0 || n_dst != n || mm_step % P.T != 0 || C < 2) return ADN_OK;
        DevBuf& pb = scratch(SCR_MM_TILES, st);
        const size_t one = (size_t)P.ntt * C * 2 * 8;
        if ((rc = pb.reserve(2 * one))) return rc;
        P.mm_raw = mm_raw ? pb.as<double>() : nullptr;
        P.mm_filt = mm_filt ? pb.as<double>() + (size_t)P.ntt * C * 2 : nullptr;
    }
    // the tile of the last source row always belongs to the walk when the final state is wanted
    P.t_out1 = zf ? P.ntt : (out_skip + n_dst - 1) / P.T + 1;
    // the second sub-cascade forgets what the first one feeds it while that one still converges
    P.pre = plan->jpre + (S2 > 0 ? plan2->jpre : 0);
    P.bulk_ok = (C == CG && CG >= 8 && fwd_env("ADN_ZP_TMA", 1)) ? 1 : 0;
    SosRun R;
    memset(&R, 0, sizeof R);
    R.src = src; R.tab = plan->dtab;
    R.n = n; R.nx = n; R.edge = 0;
    R.C = C; R.CG = CG; R.ngroups = P.ngroups; R.T = P.T; R.ntt = P.ntt;
    {
        const bool even = (C % 2 == 0) && (CG % 2 == 0);
        R.vec_in = even && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        R.lc = 0;
        while ((1 << R.lc) < CG) ++R.lc;
        if (C % 2 && CG > 1) R.lc = -1;
    }
    const int nteam = fwd_env("ADN_SOS_NTEAM", 3) >= 4 ? 4 : 3;
    const size_t TS = (size_t)(SOS_NT / CG) * (SOS_L * CG + (CG < 16 ? CG : 0));
    auto stage_doubles = [&](int D) {
        return (size_t)plan->n_staged * D * D + (size_t)D * D + (size_t)nteam * SOS_NW * CG * D + (size_t)nteam * CG * D;
    };
    const size_t wmm_d = (size_t)nteam * SOS_NW * CG * 3;
    const size_t smem = (stage_doubles(D1) + stage_doubles(D2) + wmm_d + (wmm_d & 1) + 2 + (size_t)nteam +
                         (size_t)(nteam & 1) + (size_t)nteam * TS) * 8;
    if (smem > 227 * 1024) return ADN_OK;
    const int64_t out_tiles = P.t_out1 - P.t_out0;
    int64_t runs = (int64_t)ctx().sm_count / P.ngroups;
    if (runs < 1) runs = 1;
    int64_t run_tiles = (out_tiles + runs - 1) / runs;
    const int64_t min_run = 4 * (int64_t)P.pre > 8 ? 4 * (int64_t)P.pre : 8;   // run-in: a quarter of a run at most
    if (run_tiles < min_run) run_tiles = min_run;
    runs = (out_tiles + run_tiles - 1) / run_tiles;
    if (runs * P.ngroups * 2 < ctx().sm_count) return ADN_OK;     // too short to fill the device this way
    if (run_tiles > 0x3fffffff || runs * P.ngroups > 0x7fffffff) return ADN_OK;
    P.run_tiles = (int32_t)run_tiles;
    const unsigned grid = (unsigned)(runs * P.ngroups);
#define ADN_FWD_CASE(A, B)                                                                            \
    rc = nteam == 4 ? launch_fwd<A, B, 4>(*plan, plan2.get(), P, R, smem, grid, st)                  \
                    : launch_fwd<A, B, 3>(*plan, plan2.get(), P, R, smem, grid, st)
    switch (S) {
        case 1: ADN_FWD_CASE(1, 0); break;
        case 2: ADN_FWD_CASE(2, 0); break;
        case 3: ADN_FWD_CASE(2, 1); break;
        default: ADN_FWD_CASE(2, 2); break;
    }
#undef ADN_FWD_CASE
    if (rc) return rc;
    if (want_mm) {
        const int64_t nseg = (n + mm_step - 1) / mm_step;
        const int64_t tps = mm_step / P.T;
        if (P.mm_raw) mm_fold_kernel<<<(unsigned)(nseg * C), 32, 0, st>>>(P.mm_raw, P.ntt, C, tps, mm_raw);
        if (P.mm_filt) mm_fold_kernel<<<(unsigned)(nseg * C), 32, 0, st>>>(P.mm_filt, P.ntt, C, tps, mm_filt);
        count_launch((P.mm_raw ? 1 : 0) + (P.mm_filt ? 1 : 0));
        ADN_CK(cudaGetLastError());
    }
    *handled = true;
    return ADN_OK;
}

}  // namespace adn
