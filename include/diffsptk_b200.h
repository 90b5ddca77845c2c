/*
 * diffsptk_b200 -- C ABI of the B200-native frame-rate analysis path.
 *
 * The reference (sp-nitech/diffsptk v4.0.0) has no FFI: its operator API is the
 * nn.Module classes in diffsptk/modules/ and the delegates in
 * diffsptk/functional.py.  Each entry point below replaces the arithmetic of one
 * reference `_forward` (cited per function, paths relative to the reference
 * root) and is what a binding of that module would call; INTEGRATION.md shows
 * the ctypes stub on the reference side.
 *
 * Conventions
 *  - Every function returns DSB200_OK (0) or a negative dsb200_status; a message
 *    for the last failure on the calling thread is available from
 *    dsb200_last_error().  Nothing throws across the ABI.
 *  - `_f32` / `_f64` suffix = element type of every data pointer (float/double);
 *    complex outputs are interleaved (re, im) pairs of that type.
 *  - All data pointers are DEVICE pointers on CUDA device `device`, contiguous,
 *    row-major, last dimension fastest, owned by the caller.  Leading batch
 *    dimensions are flattened by the caller (`batch` x T waveforms, or `rows` x D
 *    frame-rate vectors).  Only the `*_host` pipeline entry points take host
 *    pointers.
 *  - Work is enqueued asynchronously on `stream` (a cudaStream_t / CUstream
 *    passed as void*; NULL = legacy default stream).  No host synchronisation,
 *    no allocation on the data path.  The only persistent state is a per-device
 *    cache of twiddle tables (created on first use of an FFT length; warm up
 *    before CUDA-graph capture).
 *  - Thread-safe and re-entrant.
 */
#ifndef DIFFSPTK_B200_H_
#define DIFFSPTK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSB200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define DSB200_API __attribute__((visibility("default")))
#else
#define DSB200_API
#endif

typedef enum dsb200_status {
  DSB200_OK = 0,
  DSB200_E_BAD_PARAM = -1,   /* maps to ValueError on the Python side */
  DSB200_E_UNSUPPORTED = -2, /* valid for the reference, not implemented here */
  DSB200_E_ALIGN = -3,
  DSB200_E_CUDA = -4
} dsb200_status;

/* Frame padding, diffsptk/modules/frame.py:130-137 (torch F.pad modes). */
enum { DSB200_PAD_CONSTANT = 0, DSB200_PAD_REFLECT = 1, DSB200_PAD_REPLICATE = 2, DSB200_PAD_CIRCULAR = 3 };
/* fftr output formats, diffsptk/modules/fftr.py:110-121. */
enum { DSB200_FFTR_COMPLEX = 0, DSB200_FFTR_REAL = 1, DSB200_FFTR_IMAG = 2, DSB200_FFTR_AMPLITUDE = 3, DSB200_FFTR_POWER = 4 };
/* spec/stft output formats, diffsptk/modules/spec.py:123-132; COMPLEX is stft-only (stft.py:206-217). */
enum { DSB200_SPEC_DB = 0, DSB200_SPEC_LOGMAG = 1, DSB200_SPEC_MAGNITUDE = 2, DSB200_SPEC_POWER = 3, DSB200_SPEC_COMPLEX = 4 };
/* acorr output formats, diffsptk/modules/acorr.py:95-107. */
enum { DSB200_ACORR_NAIVE = 0, DSB200_ACORR_NORMALIZED = 1, DSB200_ACORR_BIASED = 2, DSB200_ACORR_UNBIASED = 3 };
/* mfcc output packing, diffsptk/modules/mfcc.py:188-197. */
enum { DSB200_MFCC_Y = 0, DSB200_MFCC_YE = 1, DSB200_MFCC_YC = 2, DSB200_MFCC_YCE = 3 };
/* per-row coefficient converters of dsb200_rowconv */
enum { DSB200_CONV_LPC2PAR = 0, DSB200_CONV_PAR2LPC = 1, DSB200_CONV_GNORM = 2, DSB200_CONV_IGNORM = 3, DSB200_CONV_NORM0 = 4 };

typedef struct dsb200_frame_params {
  int32_t frame_length;  /* L >= 1 */
  int32_t frame_period;  /* P >= 1 */
  int32_t center;        /* bool: pad (L/2, (L-1)/2) else (0, L-1) */
  int32_t zmean;         /* bool: subtract the per-frame mean */
  int32_t pad_mode;      /* DSB200_PAD_* */
} dsb200_frame_params;

typedef struct dsb200_spec_params {
  int32_t fft_length;          /* even, >= 2 */
  int32_t out_format;          /* DSB200_SPEC_* */
  int32_t has_relative_floor;  /* bool */
  int32_t reserved;
  double eps;                  /* added to the power spectrum */
  double relative_floor;       /* LINEAR factor 10^(dB/10), spec.py:121-122 */
} dsb200_spec_params;

typedef struct dsb200_stft_params {
  dsb200_frame_params frame;
  dsb200_spec_params spec;
} dsb200_stft_params;

typedef struct dsb200_fbank_params {
  int32_t fft_length;  /* L; input rows have L/2+1 bins */
  int32_t n_channel;   /* C */
  int32_t use_power;   /* bool: feed x instead of sqrt(x) to the filters */
  int32_t want_energy; /* bool: also produce E */
  double floor;        /* > 0 */
  double gamma;        /* generalized-log parameter; 0 = log */
} dsb200_fbank_params;

typedef struct dsb200_mfcc_params {
  dsb200_fbank_params fbank; /* use_power is ignored (always amplitude, mfcc.py:214) */
  int32_t mfcc_order;        /* M < C */
  int32_t out_format;        /* DSB200_MFCC_* */
} dsb200_mfcc_params;

typedef struct dsb200_mcep_params {
  int32_t fft_length; /* L */
  int32_t cep_order;  /* M, L >= 2M */
  int32_t n_iter;     /* Newton iterations */
  int32_t reserved;
} dsb200_mcep_params;

DSB200_API int dsb200_version(void);
/* Message of the last non-OK return on this thread ("" if none). Never NULL. */
DSB200_API const char* dsb200_last_error(void);
/* Kernel launches issued by this library since load (all threads); the bench's `gpu_launches`. */
DSB200_API int64_t dsb200_launch_count(void);
/* Name of the kernel the calling thread launched last ("" if none): lets a test assert WHICH path served a call
 * (the specialised kernel or the general one).  Never NULL; points to a string literal. */
DSB200_API const char* dsb200_last_kernel(void);
/* The persistent kernels of this library launch one CTA per SM.  A collective that should run WHILE they run (the
 * chunked NCCL all-gather of batch-sharded features) finds no SM free until they finish; with a
 * margin of n SMs the kernels launch on (SM count - n) SMs and leave the rest to it.  Process-wide; returns the
 * previous margin (>= 0), or DSB200_E_BAD_PARAM. */
DSB200_API int dsb200_set_sm_margin(int32_t n_sms);
/* Tuning knobs of the kernels (warps per CTA, kernel variants, start-up stagger; the table in README.md).  A knob is an
 * integer looked up at every launch: a value set here wins over the environment variable DSB200_<name>, which wins over
 * the built-in default (the measured winner).  No reference counterpart: measurement plumbing (tools/sweep_knobs.py).
 * dsb200_clear_knobs forgets every value set through dsb200_set_knob. */
DSB200_API int dsb200_set_knob(const char* name, int32_t value);
DSB200_API int dsb200_clear_knobs(void);

/* Number of frames for a waveform of T samples: (T-1)/P + 1 (frame.py:138); 0 if T <= 0. */
DSB200_API int64_t dsb200_num_frames(int64_t T, int32_t frame_period);

#define DSB200_DECL2(name, args) DSB200_API int name##_f32 args; DSB200_API int name##_f64 args;

/* Frame._forward, diffsptk/modules/frame.py:120-141.   x[batch,T] -> y[batch,N,L] */
DSB200_DECL2(dsb200_frame, (const void* x, void* y, int64_t batch, int64_t T,
                            const dsb200_frame_params* p, int device, void* stream))

/* Window._forward, diffsptk/modules/window.py:185-193.  y[r,:] = pad_or_truncate(x[r,:] * w, out_length) */
DSB200_DECL2(dsb200_window, (const void* x, const void* w, void* y, int64_t rows, int32_t in_length,
                             int32_t out_length, int device, void* stream))

/* RealValuedFastFourierTransform._forward (non-learnable), diffsptk/modules/fftr.py:136-151.
 * x[rows,in_length] is zero-padded / truncated to fft_length (torch.fft.rfft(x, n)); y[rows, L/2+1]
 * (interleaved complex for DSB200_FFTR_COMPLEX). */
DSB200_DECL2(dsb200_rfft, (const void* x, void* y, int64_t rows, int32_t in_length, int32_t fft_length,
                           int32_t out_format, int device, void* stream))

/* Spectrum._forward, diffsptk/modules/spec.py:152-178.  b and/or a may be NULL (not both).
 * b[rows,b_length], a[rows,a_length] with a[:,0] the gain K (remove_gain, utils/private.py:200-209). */
DSB200_DECL2(dsb200_spec, (const void* b, int32_t b_length, const void* a, int32_t a_length, void* y,
                           int64_t rows, const dsb200_spec_params* p, int device, void* stream))

/* ShortTimeFourierTransform._forward = spec(window(frame(x))), diffsptk/modules/stft.py:237-241,
 * as ONE kernel: x[batch,T], window[frame_length] -> y[batch,N,L/2+1] (complex interleaved for
 * DSB200_SPEC_COMPLEX).  The window table is the buffer Window._precompute builds (window.py:122-183). */
DSB200_DECL2(dsb200_stft, (const void* x, const void* window, void* y, int64_t batch, int64_t T,
                           const dsb200_stft_params* p, int device, void* stream))

/* Autocorrelation._forward, diffsptk/modules/acorr.py:110-120 (time-domain lag sums; the reference's
 * FFT route computes the same linear autocorrelation).  x[rows,L] -> r[rows,M+1] */
DSB200_DECL2(dsb200_acorr, (const void* x, void* r, int64_t rows, int32_t frame_length, int32_t acr_order,
                            int32_t out_format, int device, void* stream))

/* LevinsonDurbin._forward, diffsptk/modules/levdur.py:113-127: solves (Toeplitz(r_0..r_{M-1}) + eps I) a = -r_{1..M}
 * by the Levinson recursion on r with r_0 + eps; gain from the un-regularised r_0.  r[rows,M+1] -> [K,a_1..a_M]. */
DSB200_DECL2(dsb200_levdur, (const void* r, void* a, int64_t rows, int32_t lpc_order, double eps,
                             int device, void* stream))

/* LinearPredictiveCodingAnalysis._forward = levdur(acorr(x)), diffsptk/modules/lpc.py:137-139, fused.
 * x[rows,L] (framed, windowed) -> [K,a_1..a_M]. */
DSB200_DECL2(dsb200_lpc, (const void* x, void* a, int64_t rows, int32_t frame_length, int32_t lpc_order,
                          double eps, int device, void* stream))

/* Fused Frame -> Window -> LPC from the waveform (README.md:198-201 pipeline; frame.py:120-141,
 * window.py:185-193 with out_length=None, lpc.py:137-139).  x[batch,T] -> a[batch,N,M+1]. */
DSB200_DECL2(dsb200_lpc_wave, (const void* x, const void* window, void* a, int64_t batch, int64_t T,
                               const dsb200_frame_params* fp, int32_t lpc_order, double eps,
                               int device, void* stream))

/* y[rows,out_dim] = x[rows,in_dim] @ W[in_dim,out_dim]: FrequencyTransform._forward
 * (diffsptk/modules/freqt.py:141-143) and DiscreteCosineTransform._forward (dct.py:135-137). */
DSB200_DECL2(dsb200_rowmat, (const void* x, const void* W, void* y, int64_t rows, int32_t in_dim,
                             int32_t out_dim, int device, void* stream))

/* MelCepstralAnalysis._forward, diffsptk/modules/mcep.py:189-224.  x[rows,L/2+1] power spectrum ->
 * mc[rows,M+1].  Tables (built on the host in float64 from the reference's own matrices, then cast):
 *   P0[L/2+1, M+1]  = irfft-cosine * halve(c0,cH) * freqt.A              (mcep.py:203-207)
 *   G [M+1, L/2+1]  = ifreqt.A * real-rfft-cosine                        (mcep.py:210-211)
 *   Hm[L/2+1, 2M+1] = irfft-cosine * rfreqt.A                            (mcep.py:214-215)
 *   alpha_vector[M+1] = (-alpha)^k                                       (mcep.py:179-181) */
DSB200_DECL2(dsb200_mcep, (const void* x, void* mc, int64_t rows, const dsb200_mcep_params* p,
                           const void* P0, const void* G, const void* Hm, const void* alpha_vector,
                           int device, void* stream))

/* MelFilterBankAnalysis._forward, diffsptk/modules/fbank.py:305-321.  x[rows,L/2+1], H[L/2+1,C];
 * col_begin/col_end[C] give the non-zero row range of each column of H (NULL = dense).
 * y[rows,C]; E[rows] (may be NULL unless want_energy). */
DSB200_DECL2(dsb200_fbank, (const void* x, const void* H, const int32_t* col_begin, const int32_t* col_end,
                            void* y, void* E, int64_t rows, const dsb200_fbank_params* p,
                            int device, void* stream))

/* MelFrequencyCepstralCoefficientsAnalysis._forward, diffsptk/modules/mfcc.py:243-256 (fbank -> DCT-II
 * dct.py:135-137 -> lifter -> pack).  W[C,C] DCT basis, lifter[M+1]; y[rows, M (+1) (+1)]. */
DSB200_DECL2(dsb200_mfcc, (const void* x, const void* H, const int32_t* col_begin, const int32_t* col_end,
                           const void* W, const void* lifter, void* y, int64_t rows,
                           const dsb200_mfcc_params* p, int device, void* stream))

/* Fused waveform -> STFT power (eps, no floor) -> MFCC in one kernel (stft.py:237-241 + mfcc.py:243-256).
 * x[batch,T] -> y[batch,N,D]. */
DSB200_DECL2(dsb200_mfcc_wave, (const void* x, const void* window, const void* H, const int32_t* col_begin,
                                const int32_t* col_end, const void* W, const void* lifter, void* y,
                                int64_t batch, int64_t T, const dsb200_stft_params* sp,
                                const dsb200_mfcc_params* mp, int device, void* stream))

/* dsb200_mfcc_wave with two extras of the B200 build.
 *  (1) `plan`: optional device copy of dsb200_mfcc_plan_build()'s output -- where each filter's support is cut and
 *      which lane walks which piece, chosen on the host so that the kernel's shared-memory loads are free of bank
 *      conflicts (NULL: the kernel cuts the supports in order).
 *  (2) `y_dst[0 .. n_dst)` (HOST array of DEVICE pointers) + `row_offset`: every feature row r of this call is
 *      written to y_dst[d] + (row_offset + r) * D for all d.  With n_dst = 1 this is dsb200_mfcc_wave.  With the
 *      output tensor mapped on several GPUs (CUDA IPC / symmetric memory peer pointers, or ONE NVSwitch multicast
 *      pointer) the all-gather of batch-sharded features happens inside the kernel's
 *      epilogue: 128-bit stores over NVLink, quad by quad, overlapped with the remaining math.  The caller owns
 *      the cross-rank synchronisation (a barrier before the buffers are rewritten and after the kernel ends). */
DSB200_DECL2(dsb200_mfcc_wave_ex, (const void* x, const void* window, const void* H, const int32_t* col_begin,
                                   const int32_t* col_end, const void* W, const void* lifter, const int32_t* plan,
                                   void* const* y_dst, int32_t n_dst, int64_t row_offset, int64_t batch, int64_t T,
                                   const dsb200_stft_params* sp, const dsb200_mfcc_params* mp, int device,
                                   void* stream))
/* Host-side planner for (1): col_begin / col_end are HOST copies of the filter supports (fbank.py:233-293 builds
 * H; its non-zero rows per column), n_bins = L/2+1; `plan` receives dsb200_mfcc_plan_ints(n_channel) int32 values
 * (plan[0] = slots used, 0 = "no plan possible", plan[2] = bank conflicts the planner could not avoid). */
DSB200_API int32_t dsb200_mfcc_plan_ints(int32_t n_channel);
DSB200_API int dsb200_mfcc_plan_build(const int32_t* col_begin, const int32_t* col_end, int32_t n_channel,
                                      int32_t n_bins, int32_t* plan);

/* ---- backward (vector-Jacobian products; SURVEY.md section 8f rank 1: the reference is differentiable) ----
 * The forward kernels are fused, so torch autograd cannot see inside them; these are their adjoints.  Nothing
 * is saved by the forward pass: each row's spectrum is recomputed on chip.  gx / gw are overwritten. */

/* d/dx and (optionally, gw != NULL) d/dwindow of dsb200_stft.  gy has the layout of the forward output. */
DSB200_DECL2(dsb200_stft_backward, (const void* x, const void* window, const void* gy, void* gx, void* gw,
                                    int64_t batch, int64_t T, const dsb200_stft_params* p, int device,
                                    void* stream))
/* d/dx of dsb200_rfft. */
DSB200_DECL2(dsb200_rfft_backward, (const void* x, const void* gy, void* gx, int64_t rows, int32_t in_length,
                                    int32_t fft_length, int32_t out_format, int device, void* stream))
/* d/db of dsb200_spec (numerator-only spectra). */
DSB200_DECL2(dsb200_spec_backward, (const void* b, int32_t b_length, const void* gy, void* gb, int64_t rows,
                                    const dsb200_spec_params* p, int device, void* stream))
/* d/dx of dsb200_frame: scatter-add of the frame gradients (adjoint of pad + unfold + mean removal). */
DSB200_DECL2(dsb200_frame_backward, (const void* gy, void* gx, int64_t batch, int64_t T,
                                     const dsb200_frame_params* p, int device, void* stream))
/* d/dx of dsb200_fbank (diffsptk/modules/fbank.py:305-330).  gy[rows, n_channel] is the gradient of the
 * (log / power-law) filter-bank outputs, gE[rows] (or NULL) that of the log energy; gx[rows, K] is overwritten. */
DSB200_DECL2(dsb200_fbank_backward, (const void* x, const void* H, const int32_t* col_begin,
                                     const int32_t* col_end, const void* gy, const void* gE, void* gx,
                                     int64_t rows, const dsb200_fbank_params* p, int device, void* stream))
/* d/dx of dsb200_acorr (diffsptk/modules/acorr.py:112-121), every out_format. */
DSB200_DECL2(dsb200_acorr_backward, (const void* x, const void* gy, void* gx, int64_t rows, int32_t frame_length,
                                     int32_t acr_order, int32_t out_format, int device, void* stream))
/* d/dr of dsb200_levdur (diffsptk/modules/levdur.py:113-127): ga[rows, M+1] = gradient of (K, a_1..a_M). */
DSB200_DECL2(dsb200_levdur_backward, (const void* r, const void* ga, void* gr, int64_t rows, int32_t lpc_order,
                                      double eps, int device, void* stream))

/* ---- the inverse of the path -------------------------------------------
 * Inverse real FFT (diffsptk/modules/ifftr.py:130-143): y[rows, fft_length/2+1] interleaved complex ->
 * x[rows, out_length], out_length <= fft_length; DC / Nyquist imaginary parts are ignored like torch.fft.irfft. */
DSB200_DECL2(dsb200_ifftr, (const void* y, void* x, int64_t rows, int32_t fft_length, int32_t out_length,
                            int device, void* stream))
/* Windowed overlap-add with sum-of-squares normalisation (diffsptk/modules/unframe.py:164-211):
 * frames[batch, n_frames, frame_length] -> out[batch, out_length]; out_length counts from frame_length/2
 * (center) or 0 and must not exceed the overlap-added span. */
DSB200_DECL2(dsb200_unframe, (const void* frames, const void* window, void* out, int64_t batch, int64_t n_frames,
                              int64_t out_length, int32_t frame_length, int32_t frame_period, int32_t center,
                              int device, void* stream))
/* Inverse STFT = unframe(ifftr(Y)[..., :frame_length]) in one kernel (diffsptk/modules/istft.py:186-193):
 * Y[batch, n_frames, fft_length/2+1] interleaved complex -> out[batch, out_length]. */
DSB200_DECL2(dsb200_istft, (const void* Y, const void* window, void* out, int64_t batch, int64_t n_frames,
                            int64_t out_length, int32_t frame_length, int32_t frame_period, int32_t fft_length,
                            int32_t center, int device, void* stream))

/* Delta (regression) features over the frame axis (diffsptk/modules/delta.py:172-194), replicate padding:
 * x[batch, n_frames, dim], window[n_windows, width] (width odd) -> y[batch, n_frames, n_windows * dim];
 * the backward entry takes gy with y's layout and overwrites gx[batch, n_frames, dim]. */
DSB200_DECL2(dsb200_delta, (const void* x, const void* window, void* y, int64_t batch, int64_t n_frames,
                            int32_t dim, int32_t n_windows, int32_t width, int device, void* stream))
DSB200_DECL2(dsb200_delta_backward, (const void* gy, const void* window, void* gx, int64_t batch,
                                     int64_t n_frames, int32_t dim, int32_t n_windows, int32_t width, int device,
                                     void* stream))

/* Per-row converters on coefficient vectors x[rows, dim] -> y[rows, dim] (y may alias x); `param` is gamma:
 *   DSB200_CONV_LPC2PAR  LinearPredictiveCoefficientsToParcorCoefficients._forward, diffsptk/modules/lpc2par.py:104-120
 *   DSB200_CONV_PAR2LPC  ParcorCoefficientsToLinearPredictiveCoefficients._forward, diffsptk/modules/par2lpc.py:100-107
 *   DSB200_CONV_GNORM    GeneralizedCepstrumGainNormalization._forward, diffsptk/modules/gnorm.py:101-112
 *   DSB200_CONV_IGNORM   GeneralizedCepstrumInverseGainNormalization._forward, diffsptk/modules/ignorm.py:98-109
 *   DSB200_CONV_NORM0    AllPoleToAllZeroDigitalFilterCoefficients._forward, diffsptk/modules/norm0.py:88-94 (param unused) */
DSB200_DECL2(dsb200_rowconv, (const void* x, void* y, int64_t rows, int32_t dim, int32_t op, double param,
                              int device, void* stream))

/* LinearPredictiveCoefficientsToLineSpectralPairs._forward, diffsptk/modules/lpc2lsp.py:159-197:
 * a[rows, lpc_order + 1] = [K, a_1..a_M] -> w[rows, lpc_order + 1] = [K or log K, scale * w_1..w_M], the line
 * spectral frequencies in ascending order (radians times `scale`: 1, 1/2pi, sr/2000pi, sr/2pi for the reference's
 * out_format radian / cycle / khz / hz).  Zeros that cannot be isolated (a double zero) are returned as NaN. */
DSB200_DECL2(dsb200_lpc2lsp, (const void* a, void* w, int64_t rows, int32_t lpc_order, int32_t log_gain,
                              double scale, int device, void* stream))

/* GeneralizedCepstrumToGeneralizedCepstrum._forward, diffsptk/modules/mgc2mgc.py:327-364: gamma conversion of
 * gain-normalised generalized cepstra, c1[rows, in_order + 1] -> c2[rows, out_order + 1], both transforms and
 * the pointwise spectrum map in one kernel (any n_fft > max(in_order, out_order) + 1, even or odd). */
DSB200_DECL2(dsb200_gc2gc, (const void* c1, void* c2, int64_t rows, int32_t in_order, int32_t out_order,
                            double in_gamma, double out_gamma, int32_t n_fft, int device, void* stream))

/* Per-row solve of (Toeplitz(t) + Hankel(h)) x = r: t[rows, order], h[rows, 2 * order - 1], r[rows, order] ->
 * x[rows, order].  The Newton step of MelGeneralizedCepstralAnalysis, diffsptk/modules/mgcep.py:219-222
 * (symmetric_toeplitz / hankel, diffsptk/utils/private.py:291-302, then torch.linalg.solve). */
DSB200_DECL2(dsb200_thsolve, (const void* t, const void* h, const void* r, void* x, int64_t rows, int32_t order,
                              int device, void* stream))

/* ---- host-buffer pipeline (the end-to-end path: pinned host -> device -> kernel -> host) -------------
 * One object owns two device staging slots and three streams (H2D, compute, D2H) and runs
 * dsb200_stft on utterance chunks so that copies overlap compute.  x_host[batch,T], y_host[batch,N,K]
 * should be page-locked for the copies to be asynchronous. */
typedef struct dsb200_pipeline dsb200_pipeline;
DSB200_API int dsb200_pipeline_create(dsb200_pipeline** out, int device, int64_t chunk_utterances, int64_t T,
                           const dsb200_stft_params* p, int is_f64);
DSB200_API int dsb200_pipeline_stft_host(dsb200_pipeline* pl, const void* x_host, const void* window_dev,
                              void* y_host, int64_t batch);
DSB200_API int dsb200_pipeline_destroy(dsb200_pipeline* pl);

#undef DSB200_DECL2

#ifdef __cplusplus
}
#endif
#endif /* DIFFSPTK_B200_H_ */
