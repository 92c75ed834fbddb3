// Declarations shared by the SOS scan kernels (sosfilt.cu) and the register-resident
// zero-phase kernel (zerophase.cu): tile geometry, the constant-bank coefficient block, the
// host-made table plan of a cascade.
#pragma once
#include "common.cuh"
#include <memory>

namespace adn {

constexpr int SOS_L = 32;          // samples per thread
constexpr int SOS_NT = 128;        // threads per block
constexpr int SOS_NW = SOS_NT / 32;
constexpr int SOS_LOOK = 32;       // look-back window (tiles)

// MODE_ENVF: forward sweep of the envelope (rectified input, odd extension); MODE_ZPF: the same
// without the rectification = forward sweep of a plain sosfiltfilt
enum { MODE_FWD = 0, MODE_ENVF = 1, MODE_REV = 2, MODE_ZPF = 3 };
#define ADN_EXT(MODE) ((MODE) == MODE_ENVF || (MODE) == MODE_ZPF)

// table layout (D x D row-major matrices, DD = D*D doubles each), packed per channel-group
// width CG (GW = 32/CG sub-chunks per warp):
//   [0, nscan)                 A^(L 2^k),  k < log2(GW)     warp scan
//   [off_fix, off_fix+GW)      A^(L j),    j < GW           fix-up inside the warp
//   [off_wpow, off_wpow+NW+1)  A^(L GW k), k <= NW          warp prefixes
//   -- the slots above are staged in shared memory (n_staged of them) --
//   [off_tile, off_tile+33)    (A^T)^j,    j <= 32          look-back over tiles
// tile records are self-validating: every double of an aggregate / inclusive state is
// published with a plain 8-byte store and read back until it differs from this pattern
// (a NaN payload no arithmetic produces), so no flag, fence or L1 invalidation is needed
constexpr unsigned long long SOS_EMPTY = 0xFFFFFFFFFFFFFFFFull;

template <int S>
struct SosK {                      // lives in the kernel's constant bank
    double coef[S][5];             // b0 b1 b2 a1 a2
    double W[2 * S][SOS_L];        // pass-A weights
};

// v += M u, M block lower triangular (section k only sees states of sections <= k)
template <int D>
__device__ __forceinline__ void matvec_acc(const double* __restrict__ M, const double (&u)[D],
                                           double (&v)[D]) {
#pragma unroll
    for (int r = 0; r < D; ++r) {
        double a = v[r];
#pragma unroll
        for (int c = 0; c <= (r | 1); ++c) a = fma(M[r * D + c], u[c], a);
        v[r] = a;
    }
}

constexpr double HALF_PI = 1.5707963267948966;

// ---- arguments of one sweep of the scan kernels, and the tile loader they share with zerophase.cu
struct SosRun {
    const double* src;
    double* dst;                   // may be null: state only
    const double* tab;
    double* agg;                   // [tile][CG][D], SOS_EMPTY until published
    double* incl;                  // [tile][CG][D]
    int32_t off_fix, off_wpow, off_tile, n_staged;
    const double* s0;              // [C][D] initial state or null
    double* zf;                    // [C][D] final state or null
    int64_t n;                     // logical length (rows the recurrence runs over)
    int64_t nx;                    // MODE_ENVF: rows of the raw input
    int64_t out_skip;              // physical rows dropped before dst row 0
    int64_t n_dst;
    int64_t ntt;                   // time tiles
    int32_t C, CG, ngroups, T;
    int32_t jdecay;                // (A^T)^j ~ 0 for j >= jdecay
    int32_t edge;                  // MODE_ENVF: odd-extension length
    int32_t clamp;                 // negative outputs -> 0
    int32_t vec_in, vec_out;       // 16-byte granules allowed
    int32_t pf_tiles;              // L2 prefetch distance in time tiles (0 = off)
    int32_t lc;                    // log2(CG), or -1: no fast path
    // sos_run_kernel: a block walks over run_tiles consecutive time tiles, after pre_tiles
    // tiles of run-in from zero state; nbuf tile buffers (nbuf - 1 tiles prefetched)
    int32_t run_tiles, pre_tiles, nbuf;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* g, int src_bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g, int src_bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// input transform of the forward sweeps: (pi/2)|x| for the envelope, identity otherwise
template <int MODE> __device__ __forceinline__ double pre_x(double x) {
    return MODE == MODE_ENVF ? HALF_PI * fabs(x) : x;
}

// ---- stage one tile (rows [t0, t0 + T) of channel group c0 .. c0 + Cw) in shared memory with
// cp.async (the caller commits / waits).  Returns true when the tile holds RAW input that the
// passes have to rectify on the fly (MODE_ENVF, tile clear of the odd extension).
template <int MODE>
__device__ __forceinline__ bool sos_load_tile(const SosRun& R, double* tile_s, int64_t t0, int c0,
                                              int Cw, int tid) {
    const int CG = R.CG, C = R.C, T = R.T;
    const int pad = CG < 16 ? CG : 0;
    const int GS = SOS_L * Cw + pad;
    // MODE_ENVF: tiles that do not touch the odd extension are loaded raw and
    // rectified on the fly; the two edge tiles are built element by element
    bool xform = false;
    int64_t shift = 0;
    if (ADN_EXT(MODE)) {
        xform = t0 >= R.edge && t0 + T <= R.edge + R.nx;
        shift = R.edge;
    }
    // fast path (block-uniform): a full channel group (CG = 2^k channels), every row of the
    // tile inside the source: the addresses of the granules a thread copies are shifts and adds
    const int lc = Cw == CG ? R.lc : -1;                   // log2(CG) when the group is full
    const int64_t nlim = ADN_EXT(MODE) ? R.edge + R.nx : R.n;
    const bool fast_in = lc >= 0 && t0 + T <= nlim && !(ADN_EXT(MODE) && !xform);
    if (fast_in) {
        const int64_t rowbase = MODE == MODE_REV ? R.n - 1 - t0 : t0 - shift;   // physical row of tile row 0
        if (R.vec_in) {
            // granule k of a thread: flat index f = 2 (tid + NT k), row r0 + k DRk (DRk = 2 NT / CG
            // rows, a multiple of 32 whenever the tile is padded): both addresses advance by
            // constants
            const int f0 = 2 * tid, r0 = f0 >> lc;
            const int DRk = (2 * SOS_NT) >> lc;
            const int64_t gstr = (MODE == MODE_REV ? -(int64_t)DRk : (int64_t)DRk) * C;
            const double* gp = R.src + (MODE == MODE_REV ? rowbase - r0 : rowbase + r0) * C + c0 + (f0 & (CG - 1));
            double* sp = tile_s + f0 + (r0 >> 5) * pad;
            if (pad) {
#pragma unroll
                for (int k = 0; k < SOS_L / 2; ++k) {
                    cp_async16(sp + k * (2 * SOS_NT + 8), gp, 16);
                    gp += gstr;
                }
            } else {
#pragma unroll
                for (int k = 0; k < SOS_L / 2; ++k) {
                    cp_async16(sp + k * (2 * SOS_NT), gp, 16);
                    gp += gstr;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < SOS_L; ++k) {
                const int f = tid + SOS_NT * k;
                const int r = f >> lc, c = c0 + (f & (CG - 1));
                const int64_t grow = MODE == MODE_REV ? rowbase - r : rowbase + r;
                cp_async8(tile_s + f + (r >> 5) * pad, R.src + grow * C + c, 8);
            }
        }
    } else
    if (ADN_EXT(MODE) && !xform) {
        // element by element, in batches whose loads are all issued before the first use (a run
        // whose first tile is an edge tile would otherwise start some 30 load latencies late)
        const int64_t nx = R.nx, edge = R.edge;
        const int total = T * Cw;
        constexpr int UB = 8;
        for (int q0 = tid; q0 < total; q0 += SOS_NT * UB) {
            double va[UB], vb[UB];
            int kind[UB], sidx[UB];          // 0: beyond the end, 1: plain row, 2: reflected, -1: no element
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int q = q0 + u * SOS_NT;
                kind[u] = -1;
                va[u] = vb[u] = 0.0;
                sidx[u] = 0;
                if (q < total) {
                    const int row = q / Cw, col = q - row * Cw;
                    const int64_t e = t0 + row;
                    sidx[u] = (row / SOS_L) * GS + (row % SOS_L) * Cw + col;
                    const double* xc = R.src + c0 + col;
                    kind[u] = 0;
                    if (e < R.n) {
                        int64_t ra, rb = -1;
                        if (e < edge) { ra = 0; rb = edge - e; }
                        else if (e < edge + nx) { ra = e - edge; }
                        else { ra = nx - 1; rb = nx - 2 - (e - edge - nx); }
                        va[u] = __ldg(xc + ra * C);
                        kind[u] = 1;
                        if (rb >= 0) { vb[u] = __ldg(xc + rb * C); kind[u] = 2; }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                if (kind[u] < 0) continue;
                double val = 0.0;
                if (kind[u] == 1) val = pre_x<MODE>(va[u]);
                else if (kind[u] == 2) val = 2.0 * pre_x<MODE>(va[u]) - pre_x<MODE>(vb[u]);
                tile_s[sidx[u]] = val;
            }
        }
    } else {
        const int gw = R.vec_in ? 2 : 1;                 // doubles per granule
        const int gpr = Cw / gw;                         // granules per row
        const int total = T * gpr;
        int row = tid / gpr, col = tid - row * gpr;
        const int drow = SOS_NT / gpr, dcol = SOS_NT - drow * gpr;
        for (int q = tid; q < total; q += SOS_NT) {
            int64_t tau = t0 + row;
            bool ok = tau < nlim;
            int64_t phys = MODE == MODE_REV ? R.n - 1 - tau : tau - shift;
            const double* gp = ok ? R.src + phys * C + c0 + col * gw : R.src;
            double* sp = tile_s + (row / SOS_L) * GS + (row % SOS_L) * Cw + col * gw;
            if (gw == 2) cp_async16(sp, gp, ok ? 16 : 0);
            else cp_async8(sp, gp, ok ? 8 : 0);
            row += drow;
            col += dcol;
            if (col >= gpr) { col -= gpr; ++row; }
        }
    }
    return xform;
}

// ---- register-resident tiles: the recurrences on SOS_L samples held by a thread, and the
// synchronisation of the pipelined kernels (teams of four warps, TMA bulk copies)
// exact DF2T recurrence over the SOS_L register-resident samples (in time order, or reversed),
// outputs in place; the sections run skewed by one sample each so that the S recurrences of
// a step are independent (as in sosfilt.cu)
template <int S, bool REVERSE>
__device__ __forceinline__ void zp_df2t(const SosK<S>& K, double (&x)[SOS_L], double (&z)[2 * S]) {
    double xin[S + 1];
#pragma unroll
    for (int j = 0; j < SOS_L + S - 1; ++j) {
#pragma unroll
        for (int s = S - 1; s >= 0; --s) {
            const int k = j - s;
            if (k < 0 || k >= SOS_L) continue;
            const int i = REVERSE ? SOS_L - 1 - k : k;
            const double xv = s == 0 ? x[i] : xin[s];
            const double y = fma(K.coef[s][0], xv, z[2 * s]);
            z[2 * s] = fma(K.coef[s][1], xv, z[2 * s + 1]) - K.coef[s][3] * y;
            z[2 * s + 1] = K.coef[s][2] * xv - K.coef[s][4] * y;
            if (s == S - 1) x[i] = y; else xin[s + 1] = y;
        }
    }
}

// zero-state end state of the register-resident samples, processed forward or reversed
template <int S, bool REVERSE>
__device__ __forceinline__ void zp_pass_a(const SosK<S>& K, const double (&x)[SOS_L], double (&v)[2 * S]) {
#pragma unroll
    for (int i = 0; i < SOS_L; ++i) {
#pragma unroll
        for (int d = 0; d < 2 * S; ++d) v[d] = fma(K.W[d][REVERSE ? SOS_L - 1 - i : i], x[i], v[d]);
    }
}

// one homogeneous step of the cascade (input 0)
template <int S>
__device__ __forceinline__ void zp_step0(const SosK<S>& K, double (&z)[2 * S]) {
    double xv = 0.0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const double y = fma(K.coef[s][0], xv, z[2 * s]);
        z[2 * s] = fma(K.coef[s][1], xv, z[2 * s + 1]) - K.coef[s][3] * y;
        z[2 * s + 1] = K.coef[s][2] * xv - K.coef[s][4] * y;
        xv = y;
    }
}

__device__ __forceinline__ void zp_team_bar(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(SOS_NT) : "memory");
}
// lane 0 polls a shared-memory tile counter that counts up along the walk until it is >= v
__device__ __forceinline__ void zp_wait_ge(const volatile long long* f, long long v, int lane) {
    if (lane == 0) {
        while (*f < v) __nanosleep(20);
        __threadfence_block();
    }
    __syncwarp();
}
// lane 0 polls a shared-memory tile counter (counts down along the walk) until it is <= v
__device__ __forceinline__ void zp_wait_le(const volatile long long* f, long long v, int lane) {
    if (lane == 0) {
        while (*f > v) __nanosleep(20);
        __threadfence_block();
    }
    __syncwarp();
}

// ---- TMA unit: 1-D bulk copies global -> shared memory, completion counted in bytes on an mbarrier
__device__ __forceinline__ uint32_t zp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void zp_mbar_init(uint64_t* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(zp_smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void zp_mbar_arrive(uint64_t* mbar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(zp_smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void zp_mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}"
                 ::"r"(zp_smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void zp_mbar_wait(uint64_t* mbar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
        ::"r"(zp_smem_u32(mbar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void zp_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(zp_smem_u32(dst)), "l"(src), "r"(bytes), "r"(zp_smem_u32(mbar)) : "memory");
}

// ---- host side plan of a cascade (tables on the device), cached by (sos, CG)
struct SosPlan {
    std::vector<double> sos;      // key
    int S = 0, CG = 0;
    int jdecay = SOS_LOOK + 1;    // (A^T)^j < 1e-30 from here on
    int jpre = SOS_LOOK + 1;      // tiles after which the cascade has forgotten its state (< 1e-20)
    int jzp = SOS_LOOK + 1;       // the same to 1e-17 (far below the rounding of the state itself)
    int off_fix = 0, off_wpow = 0, off_tile = 0, n_staged = 0;
    std::vector<double> W;        // [D][L]
    double* dtab = nullptr;       // device tables
    ~SosPlan();
};
// plans are shared: a caller keeps its plan alive for the duration of its launch even when the
// cache evicts it
int32_t get_sos_plan(const double* sos, int S, int CG, cudaStream_t st, std::shared_ptr<SosPlan>* out);
int pick_cg(int C);               // channels per group for a channel count

template <int S>
inline void fill_sosk(const SosPlan& plan, SosK<S>& K) {
    for (int s = 0; s < S; ++s) {
        const double* q = plan.sos.data() + 6 * s;
        K.coef[s][0] = q[0]; K.coef[s][1] = q[1]; K.coef[s][2] = q[2];
        K.coef[s][3] = q[4]; K.coef[s][4] = q[5];
    }
    for (int i = 0; i < 2 * S * SOS_L; ++i) (&K.W[0][0])[i] = plan.W[i];
}

struct ZiK { double z[2 * ADN_MAX_SECTIONS]; };      // sosfilt_zi(sos), passed by value
void sosfilt_zi_host(const double* sos, int S, double* zi);

// zerophase.cu: sosfiltfilt (rect: of (pi/2)|x|) in one pass over the input with the tile held
// in registers.  The sequence is [edge_left odd extension | n_src rows | edge_right odd
// extension] (scipy's padding; an edge of 0 rows = that end is not an end of the recording and
// the sweep starts there from zero state instead of sosfilt_zi).  dst rows = sequence rows
// [out_first, out_first + n_dst).  *handled = false (and nothing launched) when the shape or
// the cascade does not suit the kernel: the caller runs the two sweeps instead.
int32_t zero_phase_regs_dev(bool rect, const double* sos, int32_t S, const double* src, int64_t n_src,
                            int32_t C, int32_t edge_left, int32_t edge_right, int64_t out_first,
                            double* dst, int64_t n_dst, int32_t clamp_negative, bool* handled,
                            cudaStream_t st);
int64_t zp_launches();
// sosfwd.cu: forward sosfilt of long traces on the pipelined skeleton (tiles in registers)
int32_t sosfilt_park_dev(const double* sos, int32_t S, const double* src, int64_t n, int32_t C,
                         int64_t out_skip, double* dst, int64_t n_dst, const double* s0, double* zf,
                         bool* handled, cudaStream_t st);
int32_t sosfilt_minmax_park_dev(const double* sos, int32_t S, const double* src, int64_t n, int32_t C,
                                int64_t out_skip, double* dst, int64_t n_dst, const double* s0, double* zf,
                                int64_t mm_step, double* mm_raw, double* mm_filt, bool* handled,
                                cudaStream_t st);
int64_t fwd_park_launches();

}  // namespace adn
