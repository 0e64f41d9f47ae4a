/* nwcwt.h - C ABI of libnwcwt.so, the B200 (sm_100a) backend of the ninwavelets
 * frequency-domain CWT path.
 *
 * The reference (Hiroki-Maeda/ninwavelets) has no FFI of its own: its GPU branch
 * is a set of cupy calls inside WaveletBase.  These entry points are what a
 * binding for that branch has to provide; each one names the reference code it
 * replaces (paths relative to the reference's ninwavelets/ directory).
 *
 * Conventions
 *   - every function returns 0 on success or a negative NWCWT_ERR_* code;
 *     nwcwt_last_error() gives a message for the calling thread.
 *   - no C++ / torch types cross the boundary: plain pointers, sizes, enums.
 *   - "dev" pointers are caller-owned device memory on the plan's device
 *     (e.g. torch.Tensor.data_ptr()); work is enqueued on the caller's
 *     cudaStream_t (passed as void*, NULL = default stream) and the call
 *     returns without synchronising.  "host" entry points take host buffers,
 *     own their staging memory and return when the result is in the buffer.
 *   - there is no CPU implementation behind this API.
 */
#ifndef NWCWT_H
#define NWCWT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NWCWT_VERSION 100 /* 0.1.0 */

enum nwcwt_error {
    NWCWT_OK = 0,
    NWCWT_ERR_INVALID = -1,     /* bad argument */
    NWCWT_ERR_UNSUPPORTED = -2, /* a request outside what the library implements (message has the detail) */
    NWCWT_ERR_CUDA = -3,        /* CUDA runtime error (message has the detail) */
    NWCWT_ERR_WORKSPACE = -4,   /* workspace too small */
    NWCWT_ERR_ZERO_FREQ = -5    /* freq == 0: base.py:234-235 raises ZeroDivisionError */
};

enum nwcwt_dtype { NWCWT_F32 = 0, NWCWT_F64 = 1 };

/* wavelets.py families.  TABLE is the route for WaveletMode.Normal families
 * (MexicanHat wavelets.py:194-228, Haar :265-280; base.py:250-255) and for user
 * subclasses with a numpy trans_formula: the caller supplies the per-frequency
 * complex spectrum and the library places it with pad_to (base.py:75-82). */
enum nwcwt_family { NWCWT_MORSE = 0, NWCWT_MORLET = 1, NWCWT_SHANNON = 2, NWCWT_TABLE = 3 };

/* base.py:404-407 (cwt, complex), :443 (abs), :425 (power) */
enum nwcwt_output { NWCWT_OUT_CWT = 0, NWCWT_OUT_ABS = 1, NWCWT_OUT_POWER = 2 };

/* Baseline, base.py:52-68, applied to every (signal, frequency) row */
enum nwcwt_baseline {
    NWCWT_BL_NONE = 0, NWCWT_BL_MEAN = 1, NWCWT_BL_RATIO = 2, NWCWT_BL_PERCENT = 3,
    NWCWT_BL_LOG = 4, NWCWT_BL_ZSCORE = 5, NWCWT_BL_ZLOG = 6
};

/* Everything make_fft_wavelets(freqs, N / sfreq) (base.py:258-279) depends on.
 * A plan is the device-side counterpart of the reference's `self.fft_wavelets`
 * cache (base.py:394-395): per-frequency parameters and bands, twiddle tables
 * and the radix plan for length n. */
typedef struct nwcwt_plan_desc {
    int32_t device;       /* CUDA device ordinal */
    int32_t dtype;        /* nwcwt_dtype: arithmetic and I/O precision */
    int32_t family;       /* nwcwt_family */
    int32_t interpolate;  /* WaveletBase.interpolate (base.py:239-242, 276, 400-401) */
    int64_t n;            /* samples per signal, wave.shape[0] */
    int32_t n_freqs;      /* len(freqs) */
    /* Resampled rows (long signals, abs / power output): a row whose non-pruned band holds B << n bins is
     * computed as an inverse transform of length n / D >= B and interpolated back to n samples.
     * 0: library default (on), 1: on, -1: off - every row takes the exact length-n inverse transform. */
    int32_t resample;
    double sfreq;         /* WaveletBase.sfreq */
    const double* freqs;  /* host, [n_freqs] */
    /* Morse (wavelets.py:38-45): p0 = b, p1 = r.
     * Morlet (wavelets.py:110-122): p0 = sigma, p1 = c * pi**-0.25, p2 = k (0 for gabor);
     *   aux[n_freqs] = peak_freq(freq) (wavelets.py:143-144), host. */
    double p0, p1, p2;
    const double* aux;
    /* TABLE: complex128 interleaved (re, im), host, [n_freqs][table_len]: the arrays
     * make_fft_wavelet (base.py:221-256) returns, before pad_to. */
    const double* table;
    int64_t table_len;
    /* optional, host, [n_freqs]: valid length of each table row (<= table_len); the Normal-mode
     * wavelets of base.py:250-253 can differ by a sample between frequencies.  NULL: all table_len. */
    const int64_t* table_lens;
    /* Bins whose spectrum magnitude is below prune_eps * peak are treated as zero
     * (never evaluated, pruned from the transform).  < 0: library default
     * (1e-12 for F32, 1e-24 for F64: far below the rounding noise of the
     * arithmetic itself); 0: evaluate every bin like the reference. */
    double prune_eps;
    /* Bound on the relative error the interpolation of a resampled row may add, for ANY input (worst case over
     * single sinusoids in the band).  <= 0: library default, 5e-7 (F32) / 5e-14 (F64) - a quarter of the
     * arithmetic's own rounding error budget.  The planner reports what it achieved in nwcwt_plan_info. */
    double resample_tol;
} nwcwt_plan_desc;

typedef struct nwcwt_plan nwcwt_plan;

typedef struct nwcwt_plan_info {
    int64_t n;
    int32_t n_freqs;
    int32_t path;          /* 0 = short rows (one CTA per signal), 1 = long rows (two passes, generic
                              kernels), 2 = long rows (two passes, packed in-place kernels), 3 = short rows
                              (packed in-place kernel, one CTA per signal pair), 4 = any length whose radix plan
                              does not exist (a prime factor > 64): Bluestein's chirp-z algorithm on transforms of
                              the smooth length n1 >= 2 n - 1 */
    int32_t n1, n2;        /* long rows: n = n1 * n2 */
    int32_t batch;         /* frequencies (short) / columns (long) interleaved per CTA */
    int32_t n_stages[2];   /* radix stages of the n (short) or n1, n2 (long) point transforms */
    int32_t radices[2][16];
    int64_t band_bins;     /* sum over frequencies of non-pruned bins */
    int64_t smem_bytes;    /* dynamic shared memory of the dominant kernel */
    int32_t threads[2];    /* threads per CTA: short kernel / pass A, pass B */
    int32_t rows_per_launch; /* long rows: (signal, frequency) rows per pass-A / pass-B launch pair */
    int32_t n_groups;      /* resampled rows: groups of frequencies by decimation (0: every row exact) */
    int32_t group_D[32];   /* decimation of each group (1 = exact rows) */
    int32_t group_K[32];   /* interpolation taps */
    int32_t group_rows[32];/* frequencies in the group */
    int32_t group_n1[32], group_n2[32]; /* split of the group's transform length n / D */
    double group_err[32];  /* worst-case relative interpolation error of the group (any input) */
} nwcwt_plan_info;

int nwcwt_version(void);
const char* nwcwt_last_error(void);
/* Number of kernels this library has launched in this process (all plans); bench.py's gpu_launches. */
int64_t nwcwt_launch_count(void);
/* Kernel-class timing for bench.py's roofline: when enabled, every launch of the transform's kernels is
 * bracketed by CUDA events on the launching stream; nwcwt_profile_read synchronises and returns the
 * accumulated milliseconds and launch counts of [0] short fused kernel, [1] forward pass A, [2] forward
 * pass B, [3] inverse pass A, [4] inverse pass B, [5] baseline rows / epoch reductions, [6] interpolation of
 * resampled rows, [7] reserved, and resets them. */
int nwcwt_profile_enable(int32_t on);
/* Measured rate of the FP32 (dtype F32: packed FFMA2 chains) or FP64 (DFMA chains) pipe of `device`, in lane
 * multiply-adds per second: the denominator of bench.py's FLOP roofline (SURVEY.md 8d asks for the HBM and the
 * FLOP fraction of every config). */
int nwcwt_fma_peak(int32_t device, int32_t dtype, double* lane_ops_per_s);
/* Test hook: non-zero makes every transform use the generic (any-length) kernels even where the plan has
 * the packed fast path, so both can be checked against the oracle on the same input. */
int nwcwt_debug_force_generic(int32_t on);
int nwcwt_profile_read(double ms[8], int64_t launches[8]);
/* Test hook: non-zero makes a plan with resampled rows run the exact length-n transform for every row. */
int nwcwt_debug_force_exact(int32_t on);

/* Host-side planning only (no CUDA call): factorisation, bands, tables.  Device
 * resources are created on first use.  Replaces make_fft_wavelets, base.py:258-279. */
int nwcwt_plan_create(nwcwt_plan** plan, const nwcwt_plan_desc* desc);
int nwcwt_plan_destroy(nwcwt_plan* plan);
int nwcwt_plan_get_info(const nwcwt_plan* plan, nwcwt_plan_info* info);
/* Non-pruned band [lo, hi) of every frequency on the n-bin axis, host, [n_freqs] each. */
int nwcwt_plan_get_bands(const nwcwt_plan* plan, int32_t* lo, int32_t* hi);

/* Device workspace needed by the calls below for n_signals signals. */
int nwcwt_workspace_bytes(const nwcwt_plan* plan, int64_t n_signals, size_t* bytes);

/* The spectrum bank itself, as make_fft_wavelets (base.py:258-279) followed by pad_to
 * (base.py:396-397) would hold it: bank_dev [n_freqs][n] complex interleaved, without the
 * 1/n of the inverse transform.  For inspection (WaveletBase.fft_wavelets) and tests; the
 * transform never materialises it. */
int nwcwt_spectrum_bank(nwcwt_plan* plan, void* bank_dev, void* stream);

/* Reductions over the epoch axis for mneutils.EpochsWavelet (mneutils.py:53-55, 68-71).
 *   kind 0: mean over epochs of a real array  in [n_epochs][count] -> out [count]
 *   kind 1: inter-trial coherence |mean(z / |z|)| of a complex array in [n_epochs][count] -> real out [count] */
int nwcwt_reduce_epochs(nwcwt_plan* plan, const void* in_dev, void* out_dev, int64_t n_epochs, int64_t count,
                        int32_t kind, void* stream);

/* Epoch reductions FUSED into the transform (mneutils.py:42-55, 57-71): signals_dev [n_channels][n_epochs][n] real;
 *   kind 0: out_dev [n_channels][n_freqs][n] = mean over epochs of |cwt|^2
 *   kind 1: out_dev [n_channels][n_freqs][n] = |mean over epochs of cwt / |cwt||   (inter-trial coherence)
 * The (n_epochs, n_freqs, n) array of per-epoch rows is never materialised: a CTA loops over the epochs of its channel
 * and accumulates into the result.  kind 1 needs a workspace of n_channels * n_freqs * n complex values.  Rows must fit
 * one CTA (plan path 3); otherwise NWCWT_ERR_UNSUPPORTED - use nwcwt_transform + nwcwt_reduce_epochs. */
int nwcwt_transform_epochs(nwcwt_plan* plan, const void* signals_dev, void* out_dev, int64_t n_channels, int64_t n_epochs,
                           int32_t kind, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Baseline (base.py:46-68) applied in place to n_rows rows of length n (dtype real): the standalone
 * form of the epilogue, for `Baseline(wave, sfreq, start, stop).<mode>()` on data already on the device. */
int nwcwt_baseline_rows(int32_t device, int32_t dtype, void* rows_dev, int64_t n_rows, int64_t n, int32_t baseline,
                        int64_t base_lo, int64_t base_hi, void* stream);

/* Forward FFT of n_signals real signals [n_signals][n] -> complex [n_signals][n]
 * (scipy.fftpack.fft / cp.fft.fft, base.py:399). */
int nwcwt_forward(nwcwt_plan* plan, const void* signals_dev, void* spectra_dev, int64_t n_signals,
                  void* workspace_dev, size_t workspace_bytes, void* stream);

/* The fused path: out[s][f][:] = epilogue(ifft(W_f * fft(signal_s))) for all plan
 * frequencies (base.py:396-406 + :443 + :425 + Baseline :46-68).
 *   out_dev: [n_signals][n_freqs][n], real (ABS/POWER) or complex interleaved (CWT)
 *   baseline != NONE needs a real output; the window is [base_lo, base_hi) in samples,
 *   i.e. wave[int(start*sfreq):int(stop*sfreq)] (base.py:49).
 * The call enqueues on `stream` and returns.  Long rows: the second call with the same arguments (buffers,
 * n_signals, modes, workspace) records its launch sequence as a CUDA graph and later identical calls replay it
 * with one cudaGraphLaunch - same kernels, same results, the buffers' current contents are read; a `stream`
 * that is itself being captured gets the plain launches.  NWCWT_GRAPH=0 in the environment switches this off.
 * A plan serves one stream / one host thread at a time (its workspace and auxiliary streams are shared state). */
int nwcwt_transform(nwcwt_plan* plan, const void* signals_dev, void* out_dev, int64_t n_signals,
                    int32_t output, int32_t baseline, int64_t base_lo, int64_t base_hi,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same computation from/to HOST buffers (what WaveletBase.cwt/power hand back:
 * cp.asarray / cp.asnumpy at base.py:398-404).  Signals are processed in chunks
 * through internally owned device and pinned staging buffers with copies and
 * kernels overlapped on two streams; returns when out_host is complete. */
int nwcwt_transform_host(nwcwt_plan* plan, const void* signals_host, void* out_host, int64_t n_signals,
                         int32_t output, int32_t baseline, int64_t base_lo, int64_t base_hi);

#ifdef __cplusplus
}
#endif
#endif /* NWCWT_H */
