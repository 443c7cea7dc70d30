/* ssb200.h -- C ABI of libssb200.so: the B200 (sm_100a) implementation of the
 * SoundSpaces per-step audio observation (binaural RIR (*) source -> waveform ->
 * log-magnitude spectrogram).
 *
 * The reference is 100 % Python and has no FFI of its own for this path; the
 * entry points below are what a ctypes binding inside the reference would bind
 * in place of the following reference code (paths relative to the upstream
 * repository, see INTEGRATION.md for the binding stub):
 *
 *   ssb_source_windows      <- the source-side half of scipy.signal.fftconvolve as
 *                              called at soundspaces/simulator.py:630,638,645,661 and
 *                              soundspaces/continuous_simulator.py:436,449 (one clip,
 *                              cached per (sound, sample offset) like
 *                              simulator.py:595-600 caches the decoded clip)
 *   ssb_convolve_batch      <- SoundSpacesSim._compute_audiogoal,
 *                              soundspaces/simulator.py:608-666 (all branches) and
 *                              ContinuousSoundSpacesSim._convolve_with_rir,
 *                              soundspaces/continuous_simulator.py:428-456
 *   ssb_crossfade_batch     <- crossfade(), soundspaces/continuous_simulator.py:47-53
 *   ssb_spectrogram_batch   <- SpectrogramSensor.compute_spectrogram,
 *                              soundspaces/tasks/nav.py:86-100
 *   ssb_render_batch        <- get_current_spectrogram_observation,
 *                              soundspaces/simulator.py:690-701 for a batch of envs
 *   ssb_render_batch_host   <- the same with host buffers (what the per-env numpy API
 *                              of the reference hands over), copies included
 *   ssb_sh_decode_batch     <- scripts/ambisonic_to_binaural.py:14-19 (closed AmbisonicBinauralizer ELF)
 *   ssb_intensity_batch     <- Intensity.get_observation, ss_baselines/av_wan/avwan_sensors.py:91-100
 *   ssb_audio_conv1_batch   <- permute(0, 3, 1, 2) + the first Conv2d + ReLU of AudioCNN,
 *                              ss_baselines/av_nav/models/audio_cnn.py:51-58,86 (rollout-time forward)
 *   ssb_logmel_batch        <- EXTENSION (no reference code): the log-mel front end BASELINE.json configs[2] names;
 *                              log1p(librosa.feature.melspectrogram) on the reference's STFT geometry (nav.py:89-92)
 *   ssb_pcm16_decode/encode <- int16 <-> float32 PCM (librosa.load decode used at
 *                              simulator.py:597; np.int16(audio*32767) at
 *                              scripts/interactive_demo.py:110)
 *
 * Conventions: every function returns 0 on success or a negative SSB_E* code and
 * never throws; ssb_last_error(ctx) describes the last failure.  All pointers
 * named d_* are device pointers on the context's device; h_* are host pointers
 * (pinned for asynchronous copies).  Work is enqueued on the caller's stream
 * (a cudaStream_t passed as void*; NULL = legacy default stream) and is
 * asynchronous; a context is not thread-safe.  No torch types cross this ABI.
 */
#ifndef SSB200_H
#define SSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_OK 0
#define SSB_E_INVALID_ARG (-1)
#define SSB_E_CUDA (-2)
#define SSB_E_OOM (-3)
#define SSB_E_SHAPE (-4)

#define SSB_FLAG_SILENT 1u      /* simulator.py:610-612: exact zeros */

#define SSB_PAD_REFLECT 0       /* librosa < 0.10 default */
#define SSB_PAD_CONSTANT 1      /* librosa >= 0.10 default */
#define SSB_LAYOUT_NCHW 0x100   /* OR into pad_mode: emit (2, 65, T') per env instead of the reference's (65, T', 2) */

#define SSB_N_FFT 512           /* nav.py:89-91 */
#define SSB_HOP 160
#define SSB_WIN 400
#define SSB_POOL 4              /* nav.py:93 */
#define SSB_SPEC_ROWS 65        /* ceil(257 / 4) */

typedef struct ssb_ctx ssb_ctx;

/* One convolution term: out[m] += sum_{k < rir_taps} rir[k] * src_ext[m0 + m - k].
 * The source enters through its overlap-save window spectra (ssb_source_windows). */
typedef struct {
    int64_t rir_offset; /* first tap of the (taps, 2) float32 interleaved RIR, in taps from d_rir_bank */
    int64_t x_offset;   /* first element (float2 units) of the window spectra, from d_xpool */
    int32_t rir_taps;   /* taps to use = min(file taps, m0 + out_samples); 0 => term contributes nothing
                           (zero-RIR fallback simulator.py:617-624, or unused distractor slot) */
    int32_t x_nw;       /* number of windows stored at x_offset */
    int32_t x_wofs;     /* stored window j holds overlap-save block index j - x_wofs */
    int32_t reserved;
} ssb_conv_term;

typedef struct {
    ssb_conv_term term[2]; /* [0] the goal sound, [1] the distractor (simulator.py:649-664) */
    int32_t out_samples;   /* valid samples; [out_samples, sr) is zero (continuous_simulator.py:454) */
    uint32_t flags;        /* SSB_FLAG_* */
} ssb_req;

/* Two convolution plans:
 *  - partitioned overlap-save (log2n 12, 13 or 14): any RIR length, one or two terms; three kernels with the RIR
 *    spectra and the partition sums as L2-visible intermediates;
 *  - single block (log2n 16): ONE 65536-point circular convolution per env by a 4-CTA cluster, RIR FFT x source
 *    spectrum x inverse FFT fused in one kernel (no intermediate but the waveform); one term, RIRs of at most
 *    65536 - sr + 1 effective taps.  The source enters as ONE 65536-point spectrum per (clip, offset):
 *    ssb_source_windows with nw = 1, wofs = 0 writes 65536 float2; x_nw / x_wofs of the request are ignored. */
typedef struct {
    int32_t log2n;      /* 12, 13, 14: FFT size of the overlap-save blocks; 16: single-block plan */
    int32_t block;      /* P = 2^log2n / 2 output samples per block (log2n 16: the shift D = 65536 - sr) */
    int32_t sr;         /* samples per output row (RIR_SAMPLING_RATE) */
    int32_t n_blocks;   /* ceil(sr / P) (log2n 16: 1) */
    int32_t max_parts;  /* RIR partitions the scratch is sized for (log2n 16: 1) */
    int32_t n_terms;    /* 1 or 2 (log2n 16: 1) */
    int64_t h_elems_per_env; /* float2 elements of RIR-spectrum scratch per env (log2n 16: 0, d_hscratch unused) */
} ssb_plan;

int ssb_version(void);
int ssb_create(int device, ssb_ctx** out);
void ssb_destroy(ssb_ctx* ctx);
const char* ssb_last_error(const ssb_ctx* ctx);
/* kernels launched by this context since creation (for bench.py's gpu_launches) */
int64_t ssb_launch_count(const ssb_ctx* ctx);

/* Optional per-kernel timing for bench.py's roofline: when enabled every kernel launch is
 * bracketed by CUDA events on the launching stream.  ssb_get_kernel_timing synchronises, writes
 * the summed milliseconds and launch counts per kernel (index SSB_K_*) and clears the record. */
#define SSB_K_FWD_RIR 0
#define SSB_K_MAC_IFFT 1
#define SSB_K_SPECTROGRAM 2
#define SSB_K_FWD_SRC 3
#define SSB_K_MAC_BINS 4
#define SSB_K_CONV64K 5
#define SSB_N_KERNELS 6
int ssb_set_kernel_timing(ssb_ctx* ctx, int enable);
int ssb_get_kernel_timing(ssb_ctx* ctx, double* ms_sum /*[SSB_N_KERNELS]*/, int64_t* counts /*[SSB_N_KERNELS]*/);

/* Convolution schedule: 0 (default) = per-bin partition sums (mac_bins_kernel) then inverse FFTs;
 * 1 = partition sums fused into the inverse-FFT kernel (re-reads every RIR partition per block). */
int ssb_set_conv_mode(ssb_ctx* ctx, int mode);
/* ssb_render_batch splits the batch over n (1..8) internal streams forked from / joined to the
 * caller's stream, so kernel tails and memory- vs compute-bound kernels of different sub-batches overlap */
int ssb_set_streams(ssb_ctx* ctx, int n);

/* Each internal stream works through n sub-batches one after the other (default 1): sub-batches = streams x n,
 * at least 16 envs each.  Smaller sub-batches keep the per-step intermediates (partition spectra, partition sums,
 * waveform: 1.3 MB per env at 44.1 kHz / 16384 taps) inside the L2. */
int ssb_set_chunks(ssb_ctx* ctx, int n);

/* Profiling only: ablation switches (1 skip the spectrum MAC, 2 skip the inverse FFT, 4 skip the
 * waveform loads of the spectrogram kernel, 8 skip its FFT, 32 force the direct-form SH decode).  Results are wrong when non-zero. */
int ssb_set_debug(ssb_ctx* ctx, int flags);

/* Fill a plan.  log2n = 0 picks the partitioned default (12; 13 when max_taps > 24576); log2n = 16 asks for the
 * single-block plan and fails with SSB_E_INVALID_ARG when n_terms != 1, sr > 61440 or max_taps > 65536 - sr + 1. */
int ssb_make_plan(ssb_ctx* ctx, int sr, int max_taps, int n_terms, int log2n, ssb_plan* plan);

/* spectrogram geometry: frames = 1 + sr/160, cols = ceil(frames/4) */
int ssb_spec_cols(int sr);

/* Overlap-save window spectra of one mono clip.  Window j (0 <= j < nw) covers source
 * samples [m0 + (j - wofs - 1) * P, m0 + (j - wofs + 1) * P); samples < 0 read as 0, samples >= S
 * read as src[n - S] when wrap != 0 (continuous_simulator.py:443-445) else 0.
 * d_x receives nw * 2^log2n float2. */
int ssb_source_windows(ssb_ctx* ctx, const ssb_plan* plan, const float* d_src, int S, int64_t m0, int wrap,
                       int nw, int wofs, void* d_x, void* stream);

/* Convolve a batch: d_wave[env][ear][n], row stride wave_stride floats (>= sr), ear 0 = left.
 * d_hscratch: B * plan->h_elems_per_env float2. */
int ssb_convolve_batch(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs,
                       const float* d_rir_bank, const void* d_xpool, void* d_hscratch,
                       float* d_wave, int64_t wave_stride, void* stream);

/* out[env][:, :n+1] = a*w1 + b*w2 over the first n+1 = int(0.05*sr)+1 samples, rest = b; in place on b. */
int ssb_crossfade_batch(ssb_ctx* ctx, int B, const float* d_prev, float* d_cur, int sr,
                        int64_t wave_stride, const uint8_t* d_enable, void* stream);

/* d_spec[env][65][cols][2] = log1p(mean4x4(|STFT(d_wave[env][ear])|)) */
int ssb_spectrogram_batch(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr,
                          int pad_mode, float* d_spec, void* stream);

/* convolve + spectrogram.  d_wave is required (it is the intermediate). */
int ssb_render_batch(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs,
                     const float* d_rir_bank, const void* d_xpool, void* d_hscratch,
                     float* d_wave, int64_t wave_stride, int pad_mode, float* d_spec, void* stream);

/* Host-buffer entry: copies h_rir (rir_bytes; offsets in h_reqs index it) and h_reqs to the device
 * staging buffers the caller provides, renders, and copies the spectrogram (and the waveform when
 * h_wave != NULL, densely packed [B][2][sr]) back to the host.  n_chunks > 1 splits the batch and
 * overlaps H2D copy / kernels / D2H copy of consecutive chunks on internal copy streams; either way
 * all of it is ordered after prior work on `stream` and `stream` completes when the results have
 * landed; the caller synchronises. */
int ssb_render_batch_host(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* h_reqs,
                          const float* h_rir, int64_t rir_bytes, float* d_rir_staging, ssb_req* d_reqs_staging,
                          const void* d_xpool, void* d_hscratch, float* d_wave, int64_t wave_stride,
                          int pad_mode, float* d_spec, float* h_spec, float* h_wave, int n_chunks, void* stream);

/* bytes the last ssb_render_batch_host call copied host->device and device->host */
int ssb_host_copy_bytes(const ssb_ctx* ctx, int64_t* h2d, int64_t* d2h);

/* Ambisonic -> binaural decode of a batch of 9-channel (ACN, second order) impulse responses:
 * d_out_rir[env][n][ear] = sum_k sum_tau (R(az_env) a)_k[n - 128 - tau] * hbank[k][ear][tau], n < L,
 * i.e. what scripts/ambisonic_to_binaural.py:14-19 obtains from the closed AmbisonicBinauralizer
 * (rotation about the vertical axis, 9 x 2 FIR filters of 256 taps, 128-sample bulk delay, output
 * cut to the input length).  d_amb: [B][L][9] f32; d_az_deg: [B] f32 degrees; d_hbank: [9][2][256] f32;
 * d_filters: scratch of B * 9 * 256 float2; d_out_rir: [B][L][2] f32, directly usable as an RIR bank. */
int ssb_sh_decode_batch(ssb_ctx* ctx, int B, const float* d_amb, int L, const float* d_az_deg, const float* d_hbank,
                        void* d_filters, float* d_out_rir, void* stream);

/* AV-WaN Intensity sensor (ss_baselines/av_wan/avwan_sensors.py:91-100) for a batch of (2, sr)
 * waveforms: d_out[env] = mean(x[:, i:i+num_frame]**2), i = first sample (min over ears) above 10 % of
 * the clip maximum. */
int ssb_intensity_batch(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int num_frame,
                        float* d_out, void* stream);

/* EXTENSION -- the reference has no mel front end; BASELINE.json configs[2] asks for one.
 * d_out[env][j][t][ear] = log1p(sum_k M[j][k] * |STFT(d_wave[env][ear])[k][t]|^power), power 1 or 2, t < 1 + sr/160,
 * STFT as in ssb_spectrogram_batch (n_fft 512, hop 160, Hann(400), centre padding per pad_mode), M the Slaney
 * filterbank of librosa.filters.mel(sr, 512, n_mels) with fmin 0, fmax sr/2 (n_mels <= 64), i.e.
 * log1p(librosa.feature.melspectrogram(y, sr, n_fft=512, hop_length=160, win_length=400, n_mels, power)).
 * ssb_mel_filterbank writes that matrix densely, [n_mels][257], to HOST memory (needs no device). */
int ssb_logmel_frames(int sr);
int ssb_mel_filterbank(int sr, int n_mels, float* h_out);
int ssb_logmel_batch(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int n_mels, int power,
                     int pad_mode, float* d_out, void* stream);

/* first layer of the policy's audio encoder, fused with the layout change.  AudioCNN.forward
 * (ss_baselines/av_nav/models/audio_cnn.py:79-89) permutes the observation to channels-first and applies
 * Conv2d(2 -> OC, KHxKW, stride SHxSW, no padding) [+ ReLU]; this reads d_spec[env][H][W][2] (the layout
 * ssb_spectrogram_batch writes) and writes d_out[env][OC][H1][W1], H1 = (H-KH)/SH + 1, W1 = (W-KW)/SW + 1.
 * d_weight: [OC][2][KH][KW] (torch's Conv2d.weight), d_bias: [OC] or NULL.  Inference only.  SSB_E_SHAPE when
 * W1 * OC > 1024 or the staged rows + weights exceed 48 KB. */
int ssb_audio_conv1_batch(ssb_ctx* ctx, int B, const float* d_spec, int H, int W, const float* d_weight,
                          const float* d_bias, int OC, int KH, int KW, int SH, int SW, int relu, float* d_out, void* stream);

/* PCM helpers.  decode: float32(x) / 32768 (exact).  encode mode 0: round(x*32768) saturated;
 * mode 1: trunc(x*32767) saturated (interactive_demo.py:110). */
int ssb_pcm16_decode(ssb_ctx* ctx, const int16_t* d_in, int64_t n, float* d_out, void* stream);
int ssb_pcm16_encode(ssb_ctx* ctx, const float* d_in, int64_t n, int mode, int16_t* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSB200_H */
