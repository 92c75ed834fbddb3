// Shared declarations of libaudian_b200: context, error handling, scratch memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include <atomic>
#include "../../include/audian_b200.h"

// Output stores of the streaming kernels.  ADN_STORE_CS=1: evict-first (st.global.cs).  Measured on
// B200 (8 ch x 48 kHz x 80 s): the spectrogram's rows of 513 doubles start at odd offsets, the
// partial sectors at the ends of every warp store merge in L1 only with the default policy
// (163.1 -> 157.7 us); the forward filter gains 1.6 % (87.2 -> 85.8 us), the zero-phase kernel loses
// 2.5 % and keeps evict-first stores (zerophase.cu).
#ifndef ADN_STORE_CS
#define ADN_STORE_CS 0
#endif
#if ADN_STORE_CS
#define ADN_STORE(p, v) __stcs((p), (v))
#else
#define ADN_STORE(p, v) (*(p) = (v))
#endif

namespace adn {

int32_t fail(int32_t code, const char* fmt, ...);
int32_t fail_cuda(cudaError_t e, const char* what);

#define ADN_CK(call)                                              \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return adn::fail_cuda(e__, #call);\
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int32_t reserve(size_t bytes);
    void release();
    template <class T> T* as() { return static_cast<T*>(p); }
};

struct Ctx {
    bool ready = false;
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;     // the library's own stream
    DevBuf in, out, aux;               // staging of the host-pointer entry points
    std::atomic<int64_t> launches{0};
};

Ctx& ctx();
int64_t option(int32_t which);         // adn_set_option() values
int64_t scan_run_launches();           // sosfilt.cu: launches of sos_run_kernel
int32_t ensure_init();
// *_dev entry points: the caller's stream as is (NULL = CUDA's default stream, which is
// what torch.cuda.current_stream().cuda_stream is unless a side stream is active)
inline cudaStream_t pick(void* s) { return static_cast<cudaStream_t>(s); }
inline void count_launch(int n = 1) { ctx().launches += n; }

// scratch that must outlive an asynchronous launch: one set of grow-only buffers per
// stream, keyed by purpose.  Calls on different streams never share scratch; calls on one
// stream are ordered by the stream.
DevBuf& scratch(int slot, cudaStream_t st);
enum { SCR_SOS_TILES = 0, SCR_SOS_TABLES, SCR_SOS_MISC, SCR_ENV_FWD, SCR_ENV_MISC,
       SCR_MINMAX_PART, SCR_SPEC_TABLES, SCR_SPEC_WORK, SCR_UNWRAP, SCR_PLAY, SCR_CHAIN_SPEC, SCR_CHAIN_ENV, SCR_MM_TILES, SCR_COUNT };

// ---- kernels' host launchers (device pointers, asynchronous) ----
int32_t minmax_dev(const double* src, int64_t n, int32_t C, int64_t step, double* dst,
                   cudaStream_t st);
int32_t sosfilt_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                    int64_t nbefore, double* dst, int64_t n_dst, const double* zi, double* zf,
                    cudaStream_t st);
int32_t envelope_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                     int64_t nbefore, double* dst, int64_t n_dst, int32_t clamp_negative,
                     cudaStream_t st);
int32_t envelope_forward_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                             int32_t edge_left, int32_t edge_right, const double* zi, double* dst,
                             double* zf, cudaStream_t st);
int32_t envelope_state0_dev(const double* sos, int32_t S, const double* src, int32_t C, int32_t edge,
                            int32_t which, double* out, cudaStream_t st);
int32_t fold_states_dev(const double* packs, const double* mats, int32_t W, int32_t C, int32_t D,
                        int32_t rank, int32_t backward, double* out, cudaStream_t st);
int32_t sosfilt_reverse_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                            const double* zi, double* dst, int64_t first, int64_t n_dst,
                            int32_t clamp_negative, double* zf, cudaStream_t st);
int32_t sosfiltfilt_dev(const double* sos, int32_t S, const double* src, int64_t n_src, int32_t C,
                        int64_t nbefore, double* dst, int64_t n_dst, cudaStream_t st);
int32_t zero_phase_range_dev(bool rect, const double* sos, int32_t S, const double* src, int64_t n_src,
                             int32_t C, int32_t edge_left, int32_t edge_right, int64_t first, double* dst,
                             int64_t n_dst, int32_t clamp_negative, cudaStream_t st);
int32_t spectrogram_dev(const double* src, int64_t n_src, int32_t C, double rate, int32_t nfft,
                        int32_t hop, int32_t window_id, int32_t detrend_id, double* dst,
                        int64_t n_dst, int32_t out_db, int64_t* n_computed, cudaStream_t st);
int32_t decibel_dev(const double* p, int64_t n, double ref_power, double min_power, double* dst,
                    cudaStream_t st);
int32_t spec_image_dev(const double* spec, int64_t n, int32_t C, int32_t F, int32_t channel, double* dst,
                       cudaStream_t st);
int32_t mean_power_dev(const double* spec, int32_t C, int32_t F, int32_t channel, int64_t i0, int64_t i1,
                       double floor_db, double* dst, cudaStream_t st);
int32_t colsum_dev(const double* spec, int64_t n, int64_t W, double* acc, cudaStream_t st);
int32_t pcm_dev(const void* pcm, int64_t n, int32_t bits, double gain, double* dst, cudaStream_t st);
int32_t synth_dev(double* dst, int64_t t0, int64_t n, int32_t C, double rate, uint64_t seed,
                  cudaStream_t st);
// ingest.cu
int32_t unwrap_dev(const double* src, int64_t n, int32_t C, double thresh, int32_t clips, double* dst,
                   cudaStream_t st);
int32_t gather_channel_dev(const double* src, int64_t n, int32_t C, int32_t channel, double* dst,
                           cudaStream_t st);
int32_t play_mix_dev(const double* src, int64_t n, int32_t C, const int32_t* left, int32_t nleft,
                     const int32_t* right, int32_t nright, double rate, double het_freq, double* dst,
                     cudaStream_t st);
int32_t decimate_dev(const double* src, int64_t n_out, int32_t C, int64_t nstep, double* dst, cudaStream_t st);

// number of spectrogram frames the reference computes (bufferedspectrogram.py:46-57)
inline int64_t spectrogram_frames(int64_t n_src, int64_t n_dst, int32_t nfft, int32_t hop) {
    if (n_dst <= 0) return 0;
    int64_t nsource = (n_dst - 1) * (int64_t)hop + nfft;
    if (nsource > n_src) nsource = n_src;
    if (nsource < nfft) return 0;
    return (nsource - (nfft - hop)) / hop;
}

}  // namespace adn
