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

}  // namespace adn
