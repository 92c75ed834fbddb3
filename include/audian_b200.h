/* audian_b200.h -- C ABI of libaudian_b200.so
 *
 * B200 (sm_100a) implementation of the derived-trace DSP path of
 * bendalab/audian.  Every entry point is what a ctypes binding on the
 * reference side would call in place of the numpy/scipy body of one
 * reference function (paths relative to the reference tree):
 *
 *   adn_sosfilt_f64      BufferedFilter.process      src/audian/bufferedfilter.py:31-36
 *   adn_envelope_f64     BufferedEnvelope.process    src/audian/bufferedenvelope.py:34-41
 *   adn_spectrogram_f64  BufferedSpectrogram.process src/audian/bufferedspectrogram.py:45-62
 *   adn_minmax_f64       down_sample_worker / CompressedData.start
 *                                                    src/audian/compresseddata.py:49-52,97-100
 *                        TraceItem.update_plot       src/audian/traceitem.py:58-61
 *   adn_decibel_f64      thunderlab decibel() as used by SpecItem.update_plot
 *                                                    src/audian/specitem.py:36
 *
 * Data layout is the reference's: float64, C-contiguous, time-major, channels
 * interleaved -- traces are (frames, C), spectrograms (frames, C, nfft/2+1).
 * Lengths are int64_t, counts int32_t.  Every function returns a status
 * (ADN_OK == 0); the message of the last failure on the calling thread is
 * returned by adn_last_error().  There is no CPU fallback: without a CUDA
 * device every compute entry point fails with ADN_ERR_CUDA.
 *
 * Host-pointer entry points copy to and from device memory that the library
 * owns, on the library's own streams (upload, kernels and download overlap
 * chunk by chunk), and block until the result is in `dst`.  They are serialised
 * by one library mutex: safe to call from several threads, one runs at a time.
 * The *_dev entry points take device pointers and a cudaStream_t (passed as
 * void*; NULL = CUDA's default stream), only enqueue work and do not
 * synchronise.  Their scratch memory is kept per stream, so calls on different
 * streams are independent; it grows on first use (cudaMalloc), so warm a
 * stream up before capturing it into a CUDA graph.
 */
#ifndef AUDIAN_B200_H
#define AUDIAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADN_OK               0
#define ADN_ERR_INVALID      1   /* bad argument (shape, NULL pointer, ...) */
#define ADN_ERR_CUDA         2   /* CUDA runtime error / no device */
#define ADN_ERR_UNSUPPORTED  3   /* valid in the reference, not built yet */
#define ADN_ERR_SHORT        4   /* envelope input not longer than the sosfiltfilt pad
                                    (scipy raises ValueError) */

#define ADN_MAX_SECTIONS     8   /* biquad sections per cascade */
#define ADN_MIN_NFFT         8
#define ADN_MAX_NFFT         1048576 /* 2^20; up to 16384 one kernel, beyond that the
                                       transforms run in a global-memory work buffer */

/* adn_set_option() */
#define ADN_OPT_RESIDENT            0  /* 1 (default): mirrors (adn_mirror_create) keep results on the
                                          device for the calls that name them; 0: mirror arguments
                                          are ignored, every call uploads its source */
#define ADN_OPT_VERIFY              1  /* reserved (was: sampled verification of the address-keyed
                                          cache of version 1.0, which no longer exists) */
#define ADN_OPT_CHUNK_BYTES         2  /* granularity of the upload/kernel/download pipeline
                                          of the host-pointer entry points (default 32 MiB) */
#define ADN_OPT_RESIDENT_MIN_BYTES  3  /* smaller results are never kept in a mirror (default 1 MiB) */
#define ADN_OPT_RESIDENT_CAP_BYTES  4  /* a result that would take the device memory of all mirrors
                                          beyond this total is not kept (default 16 GiB) */
#define ADN_OPT_ENVELOPE_CHUNK_BYTES 5 /* reserved (was: chunked two-sweep schedule of the envelope;
                                          superseded by the one-pass kernel, option 7) */
#define ADN_OPT_SCAN_RUNS           6  /* 1 (default): SOS cascades that forget their state within
                                          a few tiles are filtered by the run kernel (a block walks
                                          along time, state handed on in shared memory, run-in from
                                          zero state); 0: always the look-back kernel */
#define ADN_OPT_ZERO_PHASE_ONEPASS  7  /* 1 (default): envelope / sosfiltfilt of cascades that forget
                                          their state within a few tiles run as ONE pass with the tile
                                          held in registers (csrc/zerophase.cu: 16 B per sample);
                                          0: always a forward and a backward sweep through memory */
#define ADN_OPT_COUNT               8

#define ADN_WINDOW_HANN      0   /* periodic Hann == scipy get_window('hann', nfft) */
#define ADN_DETREND_NONE     0
#define ADN_DETREND_CONSTANT 1   /* subtract the frame mean */

/* ---- context -------------------------------------------------------- */
int32_t adn_init(int32_t device);      /* idempotent; selects the device */
int32_t adn_shutdown(void);
const char* adn_last_error(void);
int32_t adn_version(void);
int64_t adn_launch_count(void);        /* kernels launched by this library so far */
int64_t adn_scan_run_count(void);      /* of these: launches of the SOS run kernels (ADN_OPT_SCAN_RUNS) */
int64_t adn_fwd_park_count(void);      /* of those: launches of the pipelined forward kernel (csrc/sosfwd.cu) */
int64_t adn_zero_phase_count(void);    /* launches of the one-pass zero-phase kernel (ADN_OPT_ZERO_PHASE_ONEPASS) */
int32_t adn_synchronize(void);         /* waits for the library's stream */
/* page-lock a host range so that the host-pointer entry points copy at full
 * PCIe rate (optional; plain pageable memory works too) */
int32_t adn_host_register(void* ptr, int64_t bytes);
int32_t adn_host_unregister(void* ptr);
/* page-locked host memory from the driver's own allocator (cudaHostAlloc): measured 6-7 % faster end to
 * end than a registered pageable range (15.7 -> 14.6 ms for the 246-MB-up / 738-MB-down step of bench.py);
 * what the trace classes' allocate_buffer() uses for `buffer`.  *ptr = NULL on failure. */
int32_t adn_host_alloc(int64_t bytes, void** ptr);
int32_t adn_host_free(void* ptr);
int32_t adn_set_option(int32_t option, int64_t value);
int64_t adn_get_option(int32_t option);
/* Mirrors: explicit hand-over of device copies from the call that produces a
 * buffer to the calls that consume it (the derived traces of a data graph:
 * reference src/audian/buffereddata.py:149-153 walks filtered -> its dests).
 * The owner of a host buffer creates a mirror and names it as `dst_mirror` in the
 * *_m call that fills the buffer (or part of it): the result then also stays on
 * the device, tagged with the host range it was copied to -- only after the
 * computation and the download succeeded.  A consumer names the same handle as
 * `src_mirror`: if its source range lies inside the mirror's valid range the
 * upload is skipped.  The owner calls adn_mirror_invalidate() whenever the host
 * buffer changes by other means (moved, reallocated, reloaded, edited in place).
 * Handle 0 = no mirror: the plain entry points never read device copies. */
int32_t adn_mirror_create(int64_t* handle);
int32_t adn_mirror_release(int64_t handle);
int32_t adn_mirror_invalidate(int64_t handle);
/* invalidates every mirror whose host range overlaps [host, host + bytes) */
int32_t adn_invalidate(const void* host, int64_t bytes);
int64_t adn_resident_hits(void);       /* sources served from a mirror so far */
/* bytes the host-pointer entry points have copied to / from the device so far */
int32_t adn_transfer_bytes(int64_t* h2d_bytes, int64_t* d2h_bytes);

/* ---- host-pointer entry points (the plugin path) -------------------- */

/* dst (2*ceil(n/step), C): row 2j = min, row 2j+1 = max over source rows
 * [j*step, min((j+1)*step, n)).  numpy semantics bit for bit: NaN propagates
 * (the last NaN in time order; the canonical quiet NaN when C == 1), equal
 * values resolve to the later row (signed zeros). */
int32_t adn_minmax_f64(const double* src, int64_t n, int32_t C, int64_t step,
                       double* dst);

/* dst[i, c] = sosfilt(sos, src[:, c])[nbefore + i], i < n_dst.
 * sos: S rows of (b0, b1, b2, a0, a1, a2), a0 == 1 (scipy layout).
 * n_dst must be <= n_src - nbefore.  zi_inout: NULL = zero initial state,
 * else (C, S, 2) doubles, read as the initial state and overwritten with the
 * state after the last source row (streaming, == scipy's zi/zf per channel).
 * S == 0 copies src[nbefore:] (the reference's `sos is None` branch). */
int32_t adn_sosfilt_f64(const double* sos, int32_t S,
                        const double* src, int64_t n_src, int32_t C,
                        int64_t nbefore, double* dst, int64_t n_dst,
                        double* zi_inout);

/* dst = sosfiltfilt(sos, (pi/2)*|src|, axis=0)[nbefore:][:n_dst] with scipy's
 * default odd padding of 3*ntaps rows and sosfilt_zi initial conditions;
 * clamp_negative != 0 sets negative results to 0.  S == 0 writes zeros. */
int32_t adn_envelope_f64(const double* sos, int32_t S,
                         const double* src, int64_t n_src, int32_t C,
                         int64_t nbefore, double* dst, int64_t n_dst,
                         int32_t clamp_negative);

/* One-sided power spectral density frames (V**2/Hz), scipy 'density' scaling:
 * nsource = min((n_dst-1)*hop + nfft, n_src); n = (nsource - (nfft-hop))/hop
 * frames are computed, dst[n:] is zero-filled (all of dst if nsource < nfft).
 * dst is (n_dst, C, nfft/2+1).  nfft: power of two in [8, 2^20];
 * 1 <= hop <= nfft.  out_db != 0 stores decibel(P) instead of P.
 * n_computed (may be NULL) receives n. */
int32_t adn_spectrogram_f64(const double* src, int64_t n_src, int32_t C,
                            double rate, int32_t nfft, int32_t hop,
                            int32_t window_id, int32_t detrend_id,
                            double* dst, int64_t n_dst, int32_t out_db,
                            int64_t* n_computed);

/* dst[i] = power[i] > min_power ? 10*log10(power[i]/ref_power) : -inf */
int32_t adn_decibel_f64(const double* power, int64_t n, double ref_power,
                        double min_power, double* dst);

/* ---- display side and ingest (SURVEY.md 8f "next" rows) -------------- */

/* Decibel image of one channel, what SpecItem.update_plot hands to setImage()
 * (src/audian/specitem.py:33-39): dst (F, n) = decibel(spec[:, channel, :].T),
 * spec (n, C, F).  With the mirror of the buffer (adn_spec_image_db_f64_m, the
 * buffer was written by adn_spectrogram_f64_m) nothing is uploaded; otherwise only
 * the rows of that channel are (strided copy). */
int32_t adn_spec_image_db_f64(const double* spec, int64_t n, int32_t C, int32_t F,
                              int32_t channel, double* dst);
/* Power spectrum of the visible frames (src/audian/spectrogramplot.py:158-160):
 * dst[f] = max(decibel(mean(spec[i0:i1, channel, f])), floor_db), F values. */
int32_t adn_mean_power_db_f64(const double* spec, int64_t n, int32_t C, int32_t F,
                              int32_t channel, int64_t i0, int64_t i1,
                              double floor_db, double* dst);
/* Raw-data ingest: n little-endian PCM samples of 16, 24 (packed) or 32 bits ->
 * float64 value / 2^(bits-1) * gain, the scaling the audio readers behind
 * thunderlab's DataLoader apply (src/audian/data.py:172-180 loads float64).  The
 * host path moves bits/8 bytes per sample over PCIe instead of 8. */
int32_t adn_pcm_to_f64(const void* pcm, int64_t n, int32_t bits, double gain,
                       double* dst);

/* dst = scipy.signal.sosfiltfilt(sos, src, axis=0)[:n_dst] (default odd padding):
 * the zero-phase low-pass of the play-back path (src/audian/databrowser.py:1725;
 * the caller heterodynes before and decimates after).  Same kernels as the envelope,
 * without the rectification. */
int32_t adn_sosfiltfilt_f64(const double* sos, int32_t S,
                            const double* src, int64_t n_src, int32_t C,
                            double* dst, int64_t n_dst);

/* ---- the same entry points with mirrors (0 = none) ------------------- */
int32_t adn_minmax_f64_m(const double* src, int64_t n, int32_t C, int64_t step,
                         double* dst, int64_t src_mirror);
int32_t adn_sosfilt_f64_m(const double* sos, int32_t S,
                          const double* src, int64_t n_src, int32_t C,
                          int64_t nbefore, double* dst, int64_t n_dst,
                          double* zi_inout, int64_t src_mirror, int64_t dst_mirror);
int32_t adn_envelope_f64_m(const double* sos, int32_t S,
                           const double* src, int64_t n_src, int32_t C,
                           int64_t nbefore, double* dst, int64_t n_dst,
                           int32_t clamp_negative, int64_t src_mirror, int64_t dst_mirror);
int32_t adn_spectrogram_f64_m(const double* src, int64_t n_src, int32_t C,
                              double rate, int32_t nfft, int32_t hop,
                              int32_t window_id, int32_t detrend_id,
                              double* dst, int64_t n_dst, int32_t out_db,
                              int64_t* n_computed, int64_t src_mirror, int64_t dst_mirror);
int32_t adn_spec_image_db_f64_m(const double* spec, int64_t n, int32_t C, int32_t F,
                                int32_t channel, double* dst, int64_t src_mirror);
int32_t adn_mean_power_db_f64_m(const double* spec, int64_t n, int32_t C, int32_t F,
                                int32_t channel, int64_t i0, int64_t i1,
                                double floor_db, double* dst, int64_t src_mirror);

/* One channel of a trace, the reduction TraceItem.update_plot draws
 * (src/audian/traceitem.py:58-61): dst (2*ceil(n/step)) = interleaved min / max of
 * src[:, channel].  With the trace's mirror the column is gathered on the device,
 * otherwise only the column is uploaded (n x 8 bytes, not n x C x 8). */
int32_t adn_minmax_channel_f64_m(const double* src, int64_t n, int32_t C, int32_t channel,
                                 int64_t step, double* dst, int64_t src_mirror);
/* The loader's unwrap option (src/audian/data.py:180 -> audioio unwrap(data, thresh,
 * clips)), in place on a (n, C) host array: a step between consecutive samples of a
 * channel below -thresh adds 2 to everything that follows, a step above +thresh takes
 * 2 away (the file format wrapped the signal around +-1); clips != 0 limits the result to
 * [-1, 1].  thresh <= 0: nothing is done.  audioio is not vendored with the reference:
 * this restates its documented behaviour (parity unpinned for this entry). */
int32_t adn_unwrap_f64(double* data, int64_t n, int32_t C, double thresh, int32_t clips);
/* The signal DataBrowser.play_region plays (src/audian/databrowser.py:1702-1731):
 * dst[:, 0] = mean(src[:, left], 1), dst[:, 1] = mean(src[:, right], 1) (nright == 0: one
 * column); if het_freq > 0 the columns are multiplied by sin(2 pi het_freq k / rate),
 * zero-phase filtered with `sos` (the caller designs butter(2, 20000, 'low', fs=rate)) and
 * every nstep-th row is kept.  dst: (ceil(n/nstep), 1 or 2); nstep is ignored (1) without
 * heterodyne. */
int32_t adn_play_region_f64_m(const double* src, int64_t n, int32_t C,
                              const int32_t* left, int32_t nleft,
                              const int32_t* right, int32_t nright,
                              double rate, double het_freq,
                              const double* sos, int32_t S, int64_t nstep,
                              double* dst, int64_t src_mirror);

/* ---- the chain: data -> filtered -> {spectrogram, envelope} in one call ---------
 * What a parameter change recomputes in the reference: BufferedFilter.update() ->
 * recompute_all() -> filtered, then its dests (src/audian/bufferedfilter.py:53,
 * buffereddata.py:149-153).  One call: the source goes up once, the filtered trace stays on
 * the device between its producer and its consumers, results travel down while the next
 * kernel runs, one wait at the end.  Every stage computes exactly what its own entry point
 * computes (adn_sosfilt_f64, adn_spectrogram_f64, adn_envelope_f64, adn_minmax_f64). */
typedef struct adn_chain {
    /* filtered = sosfilt(sos, src)[nbefore:][:n_filt]; S == 0: copy */
    const double* sos;          /* host, S x 6 */
    int32_t S;
    int32_t out_db;             /* spectrogram in decibel */
    int64_t nbefore;
    /* spectrogram of filtered[spec_first : spec_first + spec_rows], n_spec destination frames */
    int32_t nfft, hop;
    int64_t spec_first, spec_rows, n_spec;
    /* envelope of filtered[env_first : env_first + env_rows], rows [env_nbefore:][:n_env] */
    const double* esos;         /* host, ES x 6 */
    int32_t ES;
    int32_t clamp_negative;
    int64_t env_first, env_rows, env_nbefore, n_env;
    /* min/max rows of the raw source (2*ceil(n_src/mm_step), C) */
    int64_t mm_step;
} adn_chain_t;
/* spec / env / minmax may be NULL: that stage is skipped.  n_computed: spectrogram frames.
 * src_mirror / filt_mirror: see adn_mirror_create (0 = none). */
int32_t adn_chain_f64(const adn_chain_t* chain, const double* src, int64_t n_src, int32_t C,
                      double rate, double* filtered, int64_t n_filt, double* spec, double* env,
                      double* minmax, int64_t* n_computed, int64_t src_mirror, int64_t filt_mirror);

/* ---- device-pointer entry points (bench / multi-GPU path) ----------- */
/* the chain on device buffers (sos / esos stay host pointers); only enqueues */
int32_t adn_chain_f64_dev(const adn_chain_t* chain, const double* src, int64_t n_src, int32_t C,
                          double rate, double* filtered, int64_t n_filt, double* spec, double* env,
                          double* minmax, int64_t* n_computed, void* stream);
/* out of place: dst != src */
int32_t adn_unwrap_f64_dev(const double* src, int64_t n, int32_t C, double thresh,
                           int32_t clips, double* dst, void* stream);
/* left / right / sos are HOST pointers */
int32_t adn_play_region_f64_dev(const double* src, int64_t n, int32_t C,
                                const int32_t* left, int32_t nleft,
                                const int32_t* right, int32_t nright,
                                double rate, double het_freq,
                                const double* sos, int32_t S, int64_t nstep,
                                double* dst, void* stream);
int32_t adn_sosfiltfilt_f64_dev(const double* sos, int32_t S,
                                const double* src, int64_t n_src, int32_t C,
                                double* dst, int64_t n_dst, void* stream);
int32_t adn_spec_image_db_f64_dev(const double* spec, int64_t n, int32_t C, int32_t F,
                                  int32_t channel, double* dst, void* stream);
int32_t adn_mean_power_db_f64_dev(const double* spec, int32_t C, int32_t F,
                                  int32_t channel, int64_t i0, int64_t i1,
                                  double floor_db, double* dst, void* stream);
/* acc[j] += sum_i spec[i, j], j < width: column sums of n rows, accumulated across calls --
 * the mean power spectrum of a streamed recording (src/audian/spectrogramplot.py:158 over all
 * frames; the caller divides by the frame count).  acc must be zeroed by the caller. */
int32_t adn_colsum_f64_dev(const double* spec, int64_t n, int64_t width, double* acc,
                           void* stream);
int32_t adn_pcm_to_f64_dev(const void* pcm, int64_t n, int32_t bits, double gain,
                           double* dst, void* stream);
int32_t adn_minmax_f64_dev(const double* src, int64_t n, int32_t C,
                           int64_t step, double* dst, void* stream);
/* sos is a HOST pointer; zi / zf are device pointers to (C, S, 2) or NULL.
 * dst may be NULL (n_dst ignored): only the final state zf is computed. */
int32_t adn_sosfilt_f64_dev(const double* sos, int32_t S,
                            const double* src, int64_t n_src, int32_t C,
                            int64_t nbefore, double* dst, int64_t n_dst,
                            const double* zi, double* zf, void* stream);
/* adn_sosfilt_f64_dev plus the full-trace min/max rows (compresseddata.py:49-52) of the raw
 * source rows (mm_raw, 2*ceil(n_src/mm_step) x C, or NULL) and / or of the filtered rows
 * (mm_filt, 2*ceil(n_dst/mm_step) x C, or NULL): BASELINE config 4 in ONE pass over the data.
 * Fused into the filter kernel when nbefore == 0, n_dst == n_src, C >= 2 and mm_step is a
 * multiple of the kernel's tile (4096 / min(8, C') rows, C' = C rounded up to a power of two);
 * otherwise the separate kernels run.  Results are those of adn_minmax_f64 either way. */
int32_t adn_sosfilt_minmax_f64_dev(const double* sos, int32_t S,
                                   const double* src, int64_t n_src, int32_t C,
                                   int64_t nbefore, double* dst, int64_t n_dst,
                                   const double* zi, double* zf,
                                   int64_t mm_step, double* mm_raw, double* mm_filt,
                                   void* stream);
int32_t adn_envelope_f64_dev(const double* sos, int32_t S,
                             const double* src, int64_t n_src, int32_t C,
                             int64_t nbefore, double* dst, int64_t n_dst,
                             int32_t clamp_negative, void* stream);
/* Zero-phase filtering of a RANGE of a longer recording (time shards with halo rows, chunks of
 * a streamed file): sosfiltfilt (rectify != 0: of (pi/2)*|src|, the envelope) over the n_src
 * rows with scipy's odd padding and sosfilt_zi initial conditions only at the ends flagged as
 * ends of the recording (edge_left / edge_right != 0); the other ends start from zero state:
 * pass adn_sos_decay_length(sos, S, tol) extra rows there and drop them.  dst rows = result
 * rows [first, first + n_dst). */
int32_t adn_zero_phase_range_f64_dev(const double* sos, int32_t S,
                                     const double* src, int64_t n_src, int32_t C,
                                     int32_t rectify, int32_t edge_left, int32_t edge_right,
                                     int64_t first, double* dst, int64_t n_dst,
                                     int32_t clamp_negative, void* stream);
/* The two sweeps of the envelope as separate steps, for time-sharded recordings
 * (audian_b200/sharded.py exchanges the boundary states in between):
 * forward sosfilt of (pi/2)*|src| extended by scipy's odd padding of edge_left /
 * edge_right rows (0 = that end is not an end of the recording) from state zi
 * (device (C, S, 2) or NULL = zero); dst (or NULL) receives all
 * edge_left + n_src + edge_right rows, zf (or NULL) the final state. */
int32_t adn_envelope_forward_f64_dev(const double* sos, int32_t S,
                                     const double* src, int64_t n_src, int32_t C,
                                     int32_t edge_left, int32_t edge_right,
                                     const double* zi, double* dst, double* zf,
                                     void* stream);
/* out (C, S, 2) = sosfilt_zi(sos) * x0 per channel, the initial state scipy's
 * sosfiltfilt gives a sweep: which = 0: x0 = 2 r[0] - r[edge], r = (pi/2)|src|
 * (src = first rows of the recording); which = 1: x0 = the row at src. */
int32_t adn_envelope_state0_f64_dev(const double* sos, int32_t S,
                                    const double* src, int32_t C, int32_t edge,
                                    int32_t which, double* out, void* stream);
/* State entering shard `rank` of a time-sharded linear recurrence from the
 * all-gathered boundary records: packs (world, 2, C, D), [i][0] = end state of
 * shard i from zero state, [i][1] = state entering the recording (read from
 * shard 0, or from shard world-1 if backward); mats (world, D, D) = A^len_i:
 * s <- mats[i] s + packs[i][0] over the shards before (after) `rank`. */
int32_t adn_fold_states_f64_dev(const double* packs, const double* mats,
                                int32_t world, int32_t C, int32_t D,
                                int32_t rank, int32_t backward, double* out,
                                void* stream);
/* sosfilt over the rows in reversed order from state zi; dst[i] (or NULL) =
 * result at row first + i, i < n_dst; zf = state after row 0. */
int32_t adn_sosfilt_reverse_f64_dev(const double* sos, int32_t S,
                                    const double* src, int64_t n_src, int32_t C,
                                    const double* zi, double* dst, int64_t first,
                                    int64_t n_dst, int32_t clamp_negative,
                                    double* zf, void* stream);
int32_t adn_spectrogram_f64_dev(const double* src, int64_t n_src, int32_t C,
                                double rate, int32_t nfft, int32_t hop,
                                int32_t window_id, int32_t detrend_id,
                                double* dst, int64_t n_dst, int32_t out_db,
                                int64_t* n_computed, void* stream);
int32_t adn_decibel_f64_dev(const double* power, int64_t n, double ref_power,
                            double min_power, double* dst, void* stream);
/* rows t0 .. t0+n of the deterministic synthetic recording,
 * identical to audian_b200.synth.synth() */
int32_t adn_synth_f64_dev(double* dst, int64_t t0, int64_t n, int32_t C,
                          double rate, uint64_t seed, void* stream);

/* ---- host-side plan inspection (no GPU needed; used by the CPU tests) -- */
/* Fills the matrices the scan kernel uses for a cascade: A (D x D state
 * transition, D = 2*S), B (D), and A^power (D x D), all row-major doubles. */
int32_t adn_sos_state_space(const double* sos, int32_t S, double* A, double* B,
                            int64_t power, double* A_pow);
/* Smallest power of two n with max|A^n| < tol (the cascade forgets its state after n
 * samples to within tol), or -1 if that needs more than 2^40 samples. */
int64_t adn_sos_decay_length(const double* sos, int32_t S, double tol);
/* pad length of scipy's sosfiltfilt for this cascade (3*ntaps) */
int32_t adn_sosfiltfilt_edge(const double* sos, int32_t S);

#ifdef __cplusplus
}
#endif
#endif /* AUDIAN_B200_H */
