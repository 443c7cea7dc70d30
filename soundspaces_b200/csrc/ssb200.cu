// libssb200.so -- sm_100a kernels + C ABI for the SoundSpaces audio observation.
// See include/ssb200.h for the boundary and DESIGN.md for the data layout.
//
// Kernels (all hand written, FP32 SIMT; no cuFFT / cuBLAS / torch):
//   fwd_rir_kernel      RIR partition -> spectrum (both ears packed as one complex signal)
//   fwd_src_kernel      source overlap-save windows -> spectra (cached per clip)
//   mac_ifft_kernel     sum_p X[b-p] * H[p], inverse FFT, emit the valid half -> waveform
//   spectrogram_kernel  frame + pad + Hann + FFT-512 (ears packed) + |.| + 4x4 mean + log1p
//   crossfade_kernel, pcm16 kernels
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <utility>
#include <vector>

#include "../../include/ssb200.h"
#include "fft16.cuh"
#include "conv64k.cuh"

using namespace ssb;

// Device-side ablation switches (ssb_set_debug bits 1, 2, 4, 8: skip the partition sums / inverse FFT / sample
// loads / STFT) cost branches and a register in the hot kernels, so they only exist in -DSSB_ABLATION builds
// (scratch/ablate.py); in the product build the kernels see a constant 0.
// Read-once operands (the RIR taps, H in the partition sums, Y in the inverse FFT) can be loaded with the
// evict-first policy (ld.global.cs) so they do not push the re-read data (source spectra, waveform) out of L2.
#ifndef SSB_STREAM_HINTS
#define SSB_STREAM_HINTS 1
#endif
#if SSB_STREAM_HINTS
#define LD_ONCE(p) __ldcs(p)
#else
#define LD_ONCE(p) (*(p))
#endif
#ifdef SSB_ABLATION
#define SSB_DBG(x) (x)
#else
#define SSB_DBG(x) 0
#endif

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
enum { K_FWD_RIR = 0, K_MAC_IFFT = 1, K_SPECTROGRAM = 2, K_FWD_SRC = 3, K_MAC_BINS = 4, K_CONV64K = 5, K_COUNT = SSB_N_KERNELS };
#define SSB_MAX_CHUNKS 32
#define SSB_MAX_STREAMS 8
struct TimedLaunch { int kernel; cudaEvent_t a, b; };

struct ssb_ctx {
    int device;
    int sm_count;
    float2* tw[16];      // twiddle tables by log2n (device)
    float2* twm64;       // single-block plan: TWM[r][tau] = w_65536^(tau r), r < 16, tau < 1024
    int c64_groups;      // single-block plan: 256-thread groups per CTA (2 or 4)
    float* window;       // 512-float centre-padded periodic Hann(400)
    int64_t launches;
    // optional per-kernel CUDA-event timing (bench.py roofline); see ssb_set_kernel_timing
    int timing;
    cudaStream_t s_h2d, s_d2h;                     // copy streams of the pipelined host entry
    cudaEvent_t ev[2 * SSB_MAX_CHUNKS + 2];
    // last two staging buffers written by the copy stream and the event after the kernels that read them
    const void* stage_ptr[2];
    cudaEvent_t stage_done[2];
    int stage_next;
    int64_t last_h2d_bytes, last_d2h_bytes;        // bytes moved by the last ssb_render_batch_host call
    float2* gscratch;                              // [env][9][4096] SH-decode filter spectra
    size_t gscratch_elems;
    float2* yscratch;                              // [env][block][N] partition sums (transposed MAC)
    size_t yscratch_elems;
    int conv_mode;       // 0: mac_bins + ifft (default), 1: fused mac_ifft
    int n_streams;       // internal compute streams of ssb_render_batch (1 = caller's stream only)
    int n_chunks;        // sub-batches each internal stream works through in turn (L2 footprint, ssb_set_chunks)
    cudaStream_t s_comp[SSB_MAX_STREAMS];
    cudaEvent_t ev_comp[SSB_MAX_STREAMS], ev_fork;
    int debug;           // ablation switches for profiling (ssb_set_debug); 0 in production
    bool attr_sh, attr_shf;   // cudaFuncSetAttribute is per DEVICE: opt-in flags live in the (per-device) context
    // mel filterbanks by (sr, n_mels), built on first use (ssb_logmel_batch)
    struct MelBank { int sr, n_mels; int2* rows; int* ofs; float* w; };
    std::vector<MelBank>* mel;
    std::vector<TimedLaunch>* timed;
    char err[512];
};

#define SSB_FAIL(ctx, code, ...)                                  \
    do {                                                          \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
        return (code);                                            \
    } while (0)

#define SSB_CUDA(ctx, call)                                                             \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            SSB_FAIL(ctx, e__ == cudaErrorMemoryAllocation ? SSB_E_OOM : SSB_E_CUDA,    \
                     "%s failed: %s", #call, cudaGetErrorString(e__));                  \
    } while (0)

static const int kSupportedLog2[] = {9, 12, 13, 14, 16};

// Every ABI entry runs on the context's device, whatever the caller's current device is, and restores the
// caller's device afterwards (one process may hold contexts on several devices: per-rank trainers do not, but
// the reference's ``GPU_DEVICE_ID`` / ``TORCH_GPU_ID`` split can put the policy on another device than the renderer).
struct DevGuard {
    int prev = -1, dev;
    explicit DevGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

// No C++ exception may cross the C ABI (include/ssb200.h: "never throws"): allocating STL containers in an entry
// are wrapped here and reported as SSB_E_OOM / SSB_E_CUDA with a message.
template <class F>
static int abi_call(ssb_ctx* ctx, F&& body) noexcept {
    if (!ctx) return SSB_E_INVALID_ARG;
    try {
        DevGuard g(ctx->device);
        return body();
    } catch (const std::bad_alloc&) {
        snprintf(ctx->err, sizeof(ctx->err), "out of host memory");
        return SSB_E_OOM;
    } catch (...) {
        snprintf(ctx->err, sizeof(ctx->err), "unexpected C++ exception inside libssb200");
        return SSB_E_CUDA;
    }
}

struct LaunchTimer {          // records an event pair around one kernel launch when timing is on
    ssb_ctx* ctx; cudaStream_t st; TimedLaunch tl; bool on;
    LaunchTimer(ssb_ctx* c, int kernel, cudaStream_t s) : ctx(c), st(s), on(c->timing && c->timed) {
        ctx->launches += 1;
        if (!on) return;
        tl.kernel = kernel;
        if (cudaEventCreate(&tl.a) != cudaSuccess || cudaEventCreate(&tl.b) != cudaSuccess) { on = false; return; }
        cudaEventRecord(tl.a, st);
    }
    ~LaunchTimer() {
        if (!on) return;
        cudaEventRecord(tl.b, st);
        try {                                     // a destructor must not throw (std::terminate): drop the sample instead
            ctx->timed->push_back(tl);
        } catch (...) {
            cudaEventDestroy(tl.a);
            cudaEventDestroy(tl.b);
        }
    }
};

// ---------------------------------------------------------------------------
// forward FFT kernels
// ---------------------------------------------------------------------------
template <int LOG2N>
__device__ __forceinline__ void store_slots(float2* __restrict__ dst, const float2 (&v)[16], int t) {
    constexpr int T = FftPlan<LOG2N>::T;
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i * T + t] = v[i];
}

// copy the twiddle rows of passes >= 1 into shared memory (a few hundred float2)
template <int LOG2N>
__device__ __forceinline__ void stage_small_twiddles(float2* __restrict__ stw, const float2* __restrict__ tw, int t) {
    using P = FftPlan<LOG2N>;
    for (int i = t; i < P::TW_SMALL_ELEMS; i += P::T) stw[i] = __ldg(tw + P::TW_SMALL_OFFSET + i);
}

// grid (max_parts * n_terms, B); block T.  H[env][term][p][N] in slot order.
template <int LOG2N>
__global__ void __launch_bounds__(FftPlan<LOG2N>::T)
fwd_rir_kernel(const ssb_req* __restrict__ reqs, const float2* __restrict__ rir_bank,
               float2* __restrict__ H, int max_parts, long long h_elems_per_env,
               const float2* __restrict__ tw) {
    using P = FftPlan<LOG2N>;
    extern __shared__ float2 smem[];
    float2* stw = smem + P::SMEM_ELEMS;
    constexpr int PART = P::N / 2;
    const int env = blockIdx.y;
    const int term = blockIdx.x / max_parts;
    const int p = blockIdx.x % max_parts;
    const ssb_req& rq = reqs[env];
    if (rq.flags & SSB_FLAG_SILENT) return;
    const int taps = rq.term[term].rir_taps;
    if (p * PART >= taps) return;                      // this partition is never read
    const int t = threadIdx.x;
    const float2* __restrict__ src = rir_bank + rq.term[term].rir_offset;
    float2 v[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        int n = p * PART + t + q * P::T;
        v[q] = n < taps ? (SSB_STREAM_HINTS ? __ldcs(src + n) : __ldg(src + n)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int q = 8; q < 16; ++q) v[q] = make_float2(0.f, 0.f);
    const Tw6 w0 = load_tw6<true>(tw, P::T, t);
    stage_small_twiddles<LOG2N>(stw, tw, t);           // visible after the first exchange barrier
    fft_forward<LOG2N>(v, t, smem, w0, stw);
    float2* dst = H + (long long)env * h_elems_per_env + ((long long)term * max_parts + p) * P::N;
    store_slots<LOG2N>(dst, v, t);
}

// grid (nw); block T.  X[j][N] in slot order; window j covers samples
// [m0 + (j - wofs - 1) P, m0 + (j - wofs + 1) P).
template <int LOG2N>
__global__ void __launch_bounds__(FftPlan<LOG2N>::T)
fwd_src_kernel(const float* __restrict__ src, int S, long long m0, int wrap, int wofs,
               float2* __restrict__ X, const float2* __restrict__ tw) {
    using P = FftPlan<LOG2N>;
    extern __shared__ float2 smem[];
    float2* stw = smem + P::SMEM_ELEMS;
    constexpr int PART = P::N / 2;
    const int j = blockIdx.x;
    const int t = threadIdx.x;
    const Tw6 w0 = load_tw6<true>(tw, P::T, t);
    stage_small_twiddles<LOG2N>(stw, tw, t);
    const long long base = m0 + (long long)(j - wofs - 1) * PART;
    float2 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        long long n = base + t + q * P::T;
        float x = 0.f;
        if (n >= 0) {
            if (n < S) x = __ldg(src + n);
            else if (wrap && n - S < S) x = __ldg(src + (n - S));
        }
        v[q] = make_float2(x, 0.f);
    }
    fft_forward<LOG2N>(v, t, smem, w0, stw);
    store_slots<LOG2N>(X + (long long)j * P::N, v, t);
}

// ---------------------------------------------------------------------------
// "transposed" multiply-accumulate: one thread per (env, spectrum slot) forms ALL the block sums
//   Y[b][s] = sum_p X[b - p + wofs][s] * H[p][s]
// so every H value crosses the L2->SM link once (the fused loop in mac_ifft_kernel re-reads each
// partition for every block: 4.9 MB per env at config 2 against 0.26 + 0.70 MB here); repeated
// touches of the same slot column are L1 hits.  grid (N / 256, B); block 256.
// ---------------------------------------------------------------------------
#ifndef MACB_PREFETCH
#define MACB_PREFETCH 0          // 0: all source windows loaded up front (measured default; 3 / 4 / 6 slower, profiles/ab_r02_prepared.log)
#endif
// Also measured and dropped (profiles/ab_r02_macb_envs.log): one CTA looping over 2 / 4 / 8 envs with the source windows kept
// in registers (X crosses the L2->SM link once per group instead of once per env): 0.0902 / 0.0926 / 0.1007 ms per step against
// 0.0902 -- the smaller grid costs more than the saved L2 traffic gains.
#ifdef MACB_MIN_BLOCKS
#define MACB_BOUNDS __launch_bounds__(256, MACB_MIN_BLOCKS)
#else
#define MACB_BOUNDS __launch_bounds__(256)
#endif
template <int LOG2N, int NPMAX, int NBMAX>
__global__ void MACB_BOUNDS
mac_bins_kernel(const ssb_req* __restrict__ reqs, const float2* __restrict__ xpool, const float2* __restrict__ H,
                int max_parts, int n_terms, long long h_elems_per_env, float2* __restrict__ Y, int n_blocks, int sr) {
    using P = FftPlan<LOG2N>;
    constexpr int PART = P::N / 2;
    const int env = blockIdx.y;
    const int s = blockIdx.x * 256 + threadIdx.x;
    const ssb_req& rq = reqs[env];
    if (rq.flags & SSB_FLAG_SILENT) return;
    const int nvalid = min(rq.out_samples, sr);
    const int nblk = (nvalid + PART - 1) / PART;
    float2* __restrict__ yo = Y + (long long)env * n_blocks * P::N + s;
    bool first = true;                                          // first term stores, later terms accumulate into Y
    for (int term = 0; term < n_terms; ++term) {
        const ssb_conv_term& ct = rq.term[term];
        if (ct.rir_taps <= 0) continue;
        const int nparts = min((ct.rir_taps + PART - 1) / PART, max_parts);
        const float2* __restrict__ hp = H + (long long)env * h_elems_per_env + (long long)term * max_parts * P::N + s;
        const float2* __restrict__ xp = xpool + ct.x_offset + s;
        if (ct.x_wofs == 0 && nparts <= NPMAX && nblk <= NBMAX && ct.x_nw >= nblk) {
            // fast path (clip starts at offset 0): every operand is loaded exactly once, up front
            float2 h[NPMAX], x[NBMAX];
#pragma unroll
            for (int p = 0; p < NPMAX; ++p) h[p] = p < nparts ? LD_ONCE(hp + (long long)p * P::N) : make_float2(0.f, 0.f);
#if MACB_PREFETCH > 0
            // UNMEASURED round-2 candidate: source windows fetched MACB_PREFETCH blocks ahead of their first use
            // instead of all up front, so only 8 + MACB_PREFETCH of them are live (fewer registers, one more CTA per SM)
#pragma unroll
            for (int w = 0; w < MACB_PREFETCH; ++w) x[w] = w < nblk ? __ldg(xp + (long long)w * P::N) : make_float2(0.f, 0.f);
#else
#pragma unroll
            for (int w = 0; w < NBMAX; ++w) x[w] = w < nblk ? __ldg(xp + (long long)w * P::N) : make_float2(0.f, 0.f);
#endif
#pragma unroll
            for (int b = 0; b < NBMAX; ++b) {
#if MACB_PREFETCH > 0
                if (b + MACB_PREFETCH < NBMAX)
                    x[b + MACB_PREFETCH] = b + MACB_PREFETCH < nblk ? __ldg(xp + (long long)(b + MACB_PREFETCH) * P::N) : make_float2(0.f, 0.f);
#endif
                if (b < nblk) {
                    // x * h = x.x * h + i * (x.y * h): two accumulators, ONE swizzled add at the end, so the
                    // inner loop is two FFMA2 per complex multiply-accumulate with no operand shuffling
                    float2 acc_re = first ? make_float2(0.f, 0.f) : yo[(long long)b * P::N];
                    float2 acc_im = make_float2(0.f, 0.f);
#pragma unroll
                    for (int p = 0; p < NPMAX; ++p) {
                        if (p <= b) {                            // window b - p >= 0; h[p] is 0 beyond nparts
                            acc_re = fma2(bcast(x[b - p].x), h[p], acc_re);
                            acc_im = fma2(bcast(x[b - p].y), h[p], acc_im);
                        }
                    }
                    yo[(long long)b * P::N] = add2(acc_re, make_float2(-acc_im.y, acc_im.x));
                }
            }
        } else {
            // general path: arbitrary window offset / partition count; repeated touches are L1 hits
            for (int b = 0; b < nblk; ++b) {
                float2 acc = first ? make_float2(0.f, 0.f) : yo[(long long)b * P::N];
                const int p_lo = max(0, b + ct.x_wofs - (ct.x_nw - 1));
                const int p_hi = min(nparts - 1, b + ct.x_wofs);
                for (int p = p_lo; p <= p_hi; ++p) {
                    const float2 h = hp[(long long)p * P::N];
                    const float2 x = __ldg(xp + (long long)(b - p + ct.x_wofs) * P::N);
                    acc = cfma(x, h, acc);
                }
                yo[(long long)b * P::N] = acc;
            }
        }
        first = false;
    }
    if (first)                                                  // no term had taps: the inverse-FFT kernel writes zeros itself
        return;
}

// ---------------------------------------------------------------------------
// multiply-accumulate over partitions + inverse FFT + emit
// grid (B, n_blocks); block T.
// ---------------------------------------------------------------------------
#ifndef IFFT_TW_GLOBAL
#define IFFT_TW_GLOBAL 1         // 1: later-pass twiddles straight from the L1-resident global table (r02 A/B: mac_ifft 45.2 -> 41.9 us per step); 0: staged in shared memory
#endif
#ifndef IFFT12_MIN_BLOCKS
#define IFFT12_MIN_BLOCKS 5      // resident CTAs per SM asked of ptxas for the N = 4096 instance
#endif
template <int LOG2N, bool FROM_Y>
__global__ void __launch_bounds__(FftPlan<LOG2N>::T, (LOG2N == 12 && FROM_Y) ? IFFT12_MIN_BLOCKS : 1)
mac_ifft_kernel(const ssb_req* __restrict__ reqs, const float2* __restrict__ xpool,
                const float2* __restrict__ H, int max_parts, int n_terms, long long h_elems_per_env,
                float* __restrict__ wave, long long wave_stride, int sr,
                const float2* __restrict__ tw, int dbg_arg) {
    using P = FftPlan<LOG2N>;
    const int dbg = SSB_DBG(dbg_arg);
    extern __shared__ float2 smem[];
    float2* stw = smem + P::SMEM_ELEMS;
    constexpr int PART = P::N / 2;
    // env is the fastest block index: CTAs resident together work on the same block b of different
    // envs and read the same source windows X[b-p] (one L2->L1 fill serves both)
    const int env = blockIdx.x;
    const int b = blockIdx.y;
    const int t = threadIdx.x;
    const ssb_req& rq = reqs[env];
    float* __restrict__ wl = wave + (long long)env * 2 * wave_stride;
    float* __restrict__ wr = wl + wave_stride;
    const int n0 = b * PART;
    const int nvalid = (rq.flags & SSB_FLAG_SILENT) ? 0 : min(rq.out_samples, sr);
    bool any = false;
    if (n0 < nvalid) {
        for (int term = 0; term < n_terms; ++term) any |= rq.term[term].rir_taps > 0;
    }
    if (!any) {                                        // silent / past the rendered window / zero RIR
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int n = n0 + t + q * P::T;
            if (n < sr) { wl[n] = 0.f; wr[n] = 0.f; }
        }
        return;
    }
    float2 acc[16];
    if constexpr (FROM_Y) {
        // the partition sums were formed per bin by mac_bins_kernel; H here is its output Y[env][b][N]
        const float2* __restrict__ yp = H + ((long long)env * gridDim.y + b) * P::N + t;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = (dbg & 1) ? make_float2(0.f, 0.f) : LD_ONCE(yp + i * P::T);
    } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
    for (int term = 0; term < n_terms; ++term) {
        const ssb_conv_term& ct = rq.term[term];
        if (ct.rir_taps <= 0 || (dbg & 1)) continue;
        const int nparts = min((ct.rir_taps + PART - 1) / PART, max_parts);
        const float2* __restrict__ Hb = H + (long long)env * h_elems_per_env + (long long)term * max_parts * P::N;
        const float2* __restrict__ Xb = xpool + ct.x_offset;
        // window index j = b - p + wofs must lie in [0, nw)
        int p_lo = max(0, b + ct.x_wofs - (ct.x_nw - 1));
        int p_hi = min(nparts - 1, b + ct.x_wofs);
        for (int p = p_lo; p <= p_hi; ++p) {
            const float2* __restrict__ hp = Hb + (long long)p * P::N + t;
            const float2* __restrict__ xp = Xb + (long long)(b - p + ct.x_wofs) * P::N + t;
            float2 h[16], x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { h[i] = hp[i * P::T]; x[i] = __ldg(xp + i * P::T); }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = cfma(x[i], h[i], acc[i]);
        }
    }
    }
#if IFFT_TW_GLOBAL
    // UNMEASURED round-2 candidate: read the later passes' twiddles straight from the (L1-resident) global table, as the
    // STFT kernel does, instead of staging them in shared memory behind a block barrier
    if (!(dbg & 2)) fft_inverse<LOG2N>(acc, t, smem, tw, tw + P::TW_SMALL_OFFSET);
#else
    stage_small_twiddles<LOG2N>(stw, tw, t);
    __syncthreads();
    if (!(dbg & 2)) fft_inverse<LOG2N>(acc, t, smem, tw, stw);
#endif
    constexpr float scale = 1.0f / (float)P::N;
#pragma unroll
    for (int q = 8; q < 16; ++q) {
        int n = n0 + t + (q - 8) * P::T;
        if (n < sr) {
            bool ok = n < nvalid;
            wl[n] = ok ? acc[q].x * scale : 0.f;
            wr[n] = ok ? acc[q].y * scale : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------
// spectrogram: one warp per POOLED COLUMN (4 consecutive STFT frames), warps fully independent
// (only __syncwarp): frame + pad + Hann + packed FFT-512 + |.| per ear, magnitudes accumulated
// over the 4 frames in registers, pooled over 4 bins with two shuffles, log1p, store.
// ---------------------------------------------------------------------------
constexpr int SPEC_BUF = 512 + 32;
#ifndef SPEC_UNROLL
#define SPEC_UNROLL 1            // frames of a pooled column unrolled (1, 2 or 4)
#endif
constexpr int kSpecUnroll = SPEC_UNROLL;
#ifndef SPEC_FOLD_R2
#define SPEC_FOLD_R2 1           // 1: fold the STFT's last radix-2 (lane) stage into the magnitude reads
#endif

// sqrt.approx.ftz.f32: one MUFU, max relative error 2^-23 (IEEE sqrtf expands to ~8 instructions, the
// non-ftz approx form to 4: range test + two scalings around the MUFU for denormal inputs)
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ int nat_idx(int k) { return k + ((k >> 8) << 3); }   // de-conflict k and 256+k

// Samples: element idx = lane + 32 q of frame f is sample 160 f + idx - 256, and 160 = 5 * 32, so
// over the 4 frames of a column lane `lane` only ever touches x_j = y[n0 + lane + 32 j], j = 1..29
// (frame fr uses j = q + 5 fr): each frame needs just 5 new samples per ear; they are fetched
// before the previous frame's FFT and shifted into a 14-deep register window.
template <bool REFLECT, bool INTERIOR>
__device__ __forceinline__ float spec_sample(const float* __restrict__ y, int n, int sr) {
    if (INTERIOR) return __ldg(y + n);
    if (REFLECT) {                       // np.pad(y, 256, mode='reflect'): librosa < 0.10
        if (n < 0) n = -n;
        if (n >= sr) n = 2 * (sr - 1) - n;
        n = max(0, min(n, sr - 1));      // only reached by frames past the end (weight 0)
        return __ldg(y + n);
    }
    return (n >= 0 && n < sr) ? __ldg(y + n) : 0.f;   // mode='constant': librosa >= 0.10
}

// One pooled column (4 frames) by one warp.  INTERIOR: every sample in range and all 4 frames real.
template <bool REFLECT, bool INTERIOR>
__device__ __forceinline__ void spec_column(const float* __restrict__ yl, const float* __restrict__ yr, int sr,
                                            int n_frames, int col, int lane, float2* __restrict__ buf,
                                            const float2* __restrict__ stw, const float* __restrict__ swin,
                                            float (&accl)[8], float (&accr)[8], float& acc64l, float& acc64r, int dbg) {
    const Tw6 w0 = load_tw6<true>(stw, 32, lane);
    const int n0 = col * SSB_POOL * SSB_HOP - SSB_N_FFT / 2 + lane;          // sample of x_0 for this lane
    // interior columns: every sample is a fixed offset from these two pointers (immediates once the frame loop is unrolled)
    const float* __restrict__ pl = yl + n0;
    const float* __restrict__ pr = yr + n0;
    float2 x[15];                                                            // window of the current frame: (L, R) pairs, q = 1..14
#pragma unroll
    for (int q = 1; q < 15; ++q) {
        if (INTERIOR) x[q] = make_float2(__ldg(pl + 32 * q), __ldg(pr + 32 * q));
        else x[q] = (dbg & 4) ? make_float2(0.f, 0.f)
                              : make_float2(spec_sample<REFLECT, false>(yl, n0 + 32 * q, sr),
                                            spec_sample<REFLECT, false>(yr, n0 + 32 * q, sr));
    }
#pragma unroll kSpecUnroll
    for (int fr = 0; fr < SSB_POOL; ++fr) {
        // prefetch the 5 new samples of the next frame (j = 15 + 5 fr ..); latency hides behind this frame's FFT
        float2 nx[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int j = 15 + 5 * fr + u;
            const bool need = fr + 1 < SSB_POOL && !(dbg & 4);
            if (INTERIOR) nx[u] = fr + 1 < SSB_POOL ? make_float2(__ldg(pl + 32 * j), __ldg(pr + 32 * j)) : make_float2(0.f, 0.f);
            else nx[u] = need ? make_float2(spec_sample<REFLECT, false>(yl, n0 + 32 * j, sr), spec_sample<REFLECT, false>(yr, n0 + 32 * j, sr))
                              : make_float2(0.f, 0.f);
        }
        const bool f_ok = INTERIOR || (col * SSB_POOL + fr < n_frames);     // frames past the end add 0 (block_reduce pads with 0)
        float2 v[16];
        v[0] = make_float2(0.f, 0.f);                          // idx < 56 and idx >= 456 lie outside the 400-tap window
        v[15] = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 1; q < 15; ++q) {
            const float w = f_ok ? __ldg(swin + lane + 32 * q) : 0.f;  // centre-padded Hann: 0 for idx < 56, idx >= 456
            v[q] = mul2(bcast(w), x[q]);
        }
#if SPEC_FOLD_R2
        // The last radix-2 stage (across lane pairs) is folded into the reads below: lane 2a holds P_a[i], lane
        // 2a+1 holds Q_a[i], Z[kap] = P[kap] + Q[kap] and Z[kap + 256] = P[kap] - Q[kap] for kap = a + 16 i.
        // P goes to the lower half of buf, Q to the upper half -- the same addresses the finished spectrum used.
        fft_forward<9, false>(v, lane, buf, w0, stw + FftPlan<9>::TW_SMALL_OFFSET);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) buf[nat_idx((lane >> 1) + 16 * i + 256 * (lane & 1))] = v[i];
        __syncwarp();
        // Z = FFT(w*(yL + i yR)):  XL[k] = (Z[k] + conj Z[N-k])/2,  XR[k] = (Z[k] - conj Z[N-k])/(2i);
        // for 0 < k < 256: Z[N-k] = Z[(256-k) + 256] = P[256-k] - Q[256-k]; Z[N-0] = Z[0]
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = lane + 32 * m;
            const int kk = (256 - k) & 255;
            const float2 a = add2(buf[nat_idx(k)], buf[nat_idx(256 + k)]);
            float2 bb;
            if (m == 0) bb = fma2(bcast(lane == 0 ? 1.f : -1.f), buf[nat_idx(256 + kk)], buf[nat_idx(kk)]);
            else bb = sub2(buf[nat_idx(kk)], buf[nat_idx(256 + kk)]);
            const float2 cb = make_float2(bb.x, -bb.y);
            const float2 l = add2(a, cb), r = sub2(a, cb);     // A + conj(B), A - conj(B)
            const float2 l2 = mul2(l, l), r2 = mul2(r, r);
            accl[m] += fast_sqrt(l2.x + l2.y);
            accr[m] += fast_sqrt(r2.x + r2.y);
        }
        if (lane == 0) {                                       // bin 256 = P[0] - Q[0] is alone in pooled row 64
            const float2 a = sub2(buf[nat_idx(0)], buf[nat_idx(256)]);
            acc64l += fabsf(a.x);
            acc64r += fabsf(a.y);
        }
#else
        if (!(dbg & 8)) fft_forward<9>(v, lane, buf, w0, stw + FftPlan<9>::TW_SMALL_OFFSET);
        __syncwarp();
        // natural order: k = (lane>>1) + 16 i + 256 (lane&1)
#pragma unroll
        for (int i = 0; i < 16; ++i) buf[nat_idx((lane >> 1) + 16 * i + 256 * (lane & 1))] = v[i];
        __syncwarp();
        // Z = FFT(w*(yL + i yR)):  XL[k] = (Z[k] + conj Z[N-k])/2,  XR[k] = (Z[k] - conj Z[N-k])/(2i)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = lane + 32 * m;
            const float2 a = buf[nat_idx(k)];
            const float2 bb = buf[nat_idx((SSB_N_FFT - k) & (SSB_N_FFT - 1))];
            const float2 cb = make_float2(bb.x, -bb.y);
            const float2 l = add2(a, cb), r = sub2(a, cb);     // A + conj(B), A - conj(B)
            const float2 l2 = mul2(l, l), r2 = mul2(r, r);
            accl[m] += fast_sqrt(l2.x + l2.y);
            accr[m] += fast_sqrt(r2.x + r2.y);
        }
        if (lane == 0) {                                       // bin 256 is alone in pooled row 64
            const float2 a = buf[nat_idx(256)];
            acc64l += fabsf(a.x);
            acc64r += fabsf(a.y);
        }
#endif
        __syncwarp();                                          // buf is rewritten by the next frame's exchange
        // slide the sample window by one hop (5 x 32 samples)
#pragma unroll
        for (int q = 1; q < 10; ++q) x[q] = x[q + 5];
#pragma unroll
        for (int u = 0; u < 5; ++u) x[10 + u] = nx[u];
    }
}

// grid (cols, B); block = ONE warp.  Everything that selects a code path (column index, interior
// test) derives from blockIdx, so the compiler can prove the warp converged and the shuffles and
// __syncwarp()s stay single instructions (with several warps per CTA and a per-warp column they
// were compiled into WARPSYNC.COLLECTIVE sequences).
#ifndef SPEC_MIN_BLOCKS
#define SPEC_MIN_BLOCKS 16
#endif
template <bool REFLECT>
__global__ void __launch_bounds__(32, SPEC_MIN_BLOCKS)
spectrogram_kernel(const float* __restrict__ wave, long long wave_stride, int sr, int n_frames, int cols,
                   float* __restrict__ out, const float2* __restrict__ tw,
                   const float* __restrict__ window, int dbg_arg, int nchw) {
    __shared__ float2 xbuf[SPEC_BUF];
    const int dbg = SSB_DBG(dbg_arg);
    const int lane = threadIdx.x;
    const int col = blockIdx.x;
    const int env = blockIdx.y;
    const float* __restrict__ yl = wave + (long long)env * 2 * wave_stride;
    const float* __restrict__ yr = yl + wave_stride;
    // every x_j, j = 1..29, in range and all four frames real => no padding logic at all
    const int first = col * SSB_POOL * SSB_HOP - SSB_N_FFT / 2;
    const bool interior = first + 32 >= 0 && first + 32 * 30 <= sr && col * SSB_POOL + SSB_POOL <= n_frames;
    float accl[8], accr[8], acc64l = 0.f, acc64r = 0.f;
#pragma unroll
    for (int m = 0; m < 8; ++m) { accl[m] = 0.f; accr[m] = 0.f; }
    if (interior)
        spec_column<REFLECT, true>(yl, yr, sr, n_frames, col, lane, xbuf, tw, window, accl, accr, acc64l, acc64r, dbg);
    else
        spec_column<REFLECT, false>(yl, yr, sr, n_frames, col, lane, xbuf, tw, window, accl, accr, acc64l, acc64r, dbg);
    // pool 4 adjacent bins (lanes): after the two xor-shuffles every lane of a 4-lane group holds the group sum;
    // lane q of the group then finishes rows m = 2q, 2q+1 only (4 log1p per lane instead of 16).
    // scale: 0.5 (ear packing) / 16 (4x4 mean)
    const int q = lane & 3;
    float ml[2] = {0.f, 0.f}, mr[2] = {0.f, 0.f};
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        float l = accl[m], r = accr[m];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        if (q == (m >> 1)) { ml[m & 1] = l; mr[m & 1] = r; }
    }
    const float v64l = log1pf(acc64l * (1.0f / 16.0f)), v64r = log1pf(acc64r * (1.0f / 16.0f));
    if (!nchw) {
        // reference layout (65, T', 2): ears interleaved (nav.py:98 np.stack(axis=-1))
        float* __restrict__ o = out + (((long long)env * SSB_SPEC_ROWS) * cols + col) * 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = (lane >> 2) + 8 * (2 * q + j);
            *reinterpret_cast<float2*>(o + (long long)row * cols * 2) =
                make_float2(log1pf(ml[j] * (0.5f / 16.0f)), log1pf(mr[j] * (0.5f / 16.0f)));
        }
        if (lane == 0) *reinterpret_cast<float2*>(o + (long long)64 * cols * 2) = make_float2(v64l, v64r);
    } else {
        // channels-first (2, 65, T'): what AudioCNN's permute(0, 3, 1, 2) produces (audio_cnn.py:86)
        float* __restrict__ ol = out + ((long long)env * 2 * SSB_SPEC_ROWS) * cols + col;
        float* __restrict__ orr = ol + (long long)SSB_SPEC_ROWS * cols;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = (lane >> 2) + 8 * (2 * q + j);
            ol[(long long)row * cols] = log1pf(ml[j] * (0.5f / 16.0f));
            orr[(long long)row * cols] = log1pf(mr[j] * (0.5f / 16.0f));
        }
        if (lane == 0) { ol[(long long)64 * cols] = v64l; orr[(long long)64 * cols] = v64r; }
    }
}


// ---------------------------------------------------------------------------
// log-mel spectrogram.  EXTENSION: BASELINE.json configs[2] and the north star name a log-mel front end, the
// reference itself has none; defined here as
//     out[env][j][t][ear] = log1p( sum_k M[j][k] * |STFT(wave[env][ear])[k][t]|^power ),   power in {1, 2},
// with the reference's STFT (n_fft 512, hop 160, Hann(400), centre padding: nav.py:89-92) and M the Slaney
// mel filterbank of librosa.filters.mel(sr, 512, n_mels) (area-normalised triangles, fmin 0, fmax sr/2), i.e.
// log1p(librosa.feature.melspectrogram(...)) -- the same log1p compression the reference applies (nav.py:97).
// One warp per group of 4 consecutive frames (the sample window slides exactly as in spec_column).  The
// filterbank is a sparse matrix -- every bin feeds at most two triangles, 2 * 257 non-zeros in all against
// 64 * 257 = 16448 dense entries -- so it is applied as per-mel sparse dot products from shared memory; a dense
// tensor-core GEMM would first have to write the (257, T) power spectra to HBM (2 x 284 KB per env at
// 44.1 kHz, more than every other byte this path moves) to do 32x the arithmetic.
// ---------------------------------------------------------------------------
constexpr int MEL_MAX = 64;                  // two mel rows per lane
constexpr int MEL_PBUF = SSB_N_FFT / 2 + 1;  // 257 bins

template <bool REFLECT>
__global__ void __launch_bounds__(32, 16)
logmel_kernel(const float* __restrict__ wave, long long wave_stride, int sr, int n_frames, int n_mels, int power,
              float* __restrict__ out, const float2* __restrict__ tw, const float* __restrict__ window,
              const int2* __restrict__ mel_rows /* (first bin, count) */, const int* __restrict__ mel_ofs,
              const float* __restrict__ mel_w) {
    __shared__ float2 buf[SPEC_BUF];
    __shared__ float2 pbuf[MEL_PBUF + 1];
    const int lane = threadIdx.x;
    const int t0 = blockIdx.x * SSB_POOL;                       // first frame of this warp
    const int env = blockIdx.y;
    const float* __restrict__ yl = wave + (long long)env * 2 * wave_stride;
    const float* __restrict__ yr = yl + wave_stride;
    const Tw6 w0 = load_tw6<true>(tw, 32, lane);
    const int n0 = t0 * SSB_HOP - SSB_N_FFT / 2 + lane;
    float2 x[15];
#pragma unroll
    for (int q = 1; q < 15; ++q)
        x[q] = make_float2(spec_sample<REFLECT, false>(yl, n0 + 32 * q, sr), spec_sample<REFLECT, false>(yr, n0 + 32 * q, sr));
    // this lane's mel rows j = lane and lane + 32
    int2 row[2];
    const float* __restrict__ wrow[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        row[h] = j < n_mels ? __ldg(mel_rows + j) : make_int2(0, 0);
        wrow[h] = mel_w + (j < n_mels ? __ldg(mel_ofs + j) : 0);
    }
    float2 res[2][SSB_POOL];
#pragma unroll 1
    for (int fr = 0; fr < SSB_POOL; ++fr) {
        float2 nx[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int n = n0 + 32 * (15 + 5 * fr + u);
            nx[u] = fr + 1 < SSB_POOL ? make_float2(spec_sample<REFLECT, false>(yl, n, sr), spec_sample<REFLECT, false>(yr, n, sr))
                                      : make_float2(0.f, 0.f);
        }
        float2 v[16];
        v[0] = make_float2(0.f, 0.f);
        v[15] = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 1; q < 15; ++q) v[q] = mul2(bcast(__ldg(window + lane + 32 * q)), x[q]);
        fft_forward<9>(v, lane, buf, w0, tw + FftPlan<9>::TW_SMALL_OFFSET);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) buf[nat_idx((lane >> 1) + 16 * i + 256 * (lane & 1))] = v[i];
        __syncwarp();
        // XL[k] = (Z[k] + conj Z[N-k]) / 2, XR[k] = (Z[k] - conj Z[N-k]) / (2i)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = lane + 32 * m;
            const float2 a = buf[nat_idx(k)];
            const float2 bb = buf[nat_idx((SSB_N_FFT - k) & (SSB_N_FFT - 1))];
            const float2 cb = make_float2(bb.x, -bb.y);
            const float2 l = add2(a, cb), r = sub2(a, cb);
            const float2 l2 = mul2(l, l), r2 = mul2(r, r);
            const float pl2 = l2.x + l2.y, pr2 = r2.x + r2.y;        // 4 |X|^2
            pbuf[k] = power == 2 ? make_float2(0.25f * pl2, 0.25f * pr2) : make_float2(0.5f * fast_sqrt(pl2), 0.5f * fast_sqrt(pr2));
        }
        if (lane == 0) {                                              // Nyquist bin: Z[256] = XL[256] + i XR[256], both real
            const float2 a = buf[nat_idx(256)];
            pbuf[256] = power == 2 ? make_float2(a.x * a.x, a.y * a.y) : make_float2(fabsf(a.x), fabsf(a.y));
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 acc = make_float2(0.f, 0.f);
            for (int i = 0; i < row[h].y; ++i) acc = fma2(bcast(__ldg(wrow[h] + i)), pbuf[row[h].x + i], acc);
            res[h][fr] = acc;
        }
        __syncwarp();                                                 // buf / pbuf are rewritten by the next frame
#pragma unroll
        for (int q = 1; q < 10; ++q) x[q] = x[q + 5];
#pragma unroll
        for (int u = 0; u < 5; ++u) x[10 + u] = nx[u];
    }
    // out[env][j][t][ear]: 4 consecutive frames of one mel row are 32 contiguous bytes
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        if (j >= n_mels) continue;
        float* __restrict__ o = out + (((long long)env * n_mels + j) * n_frames + t0) * 2;
#pragma unroll
        for (int fr = 0; fr < SSB_POOL; ++fr)
            if (t0 + fr < n_frames)
                *reinterpret_cast<float2*>(o + 2 * fr) = make_float2(log1pf(res[h][fr].x), log1pf(res[h][fr].y));
    }
}

// ---------------------------------------------------------------------------
// ambisonic -> binaural decode (reference: scripts/ambisonic_to_binaural.py:14-19 driving the
// closed AmbisonicBinauralizer): out[n][ear] = sum_k sum_tau (R(az) a)_k[n - 128 - tau] h[k][ear][tau].
// The rotation is folded into per-env filters g[k][tau] = sum_j R[j][k] h[j][.][tau] (sh_filters_kernel),
// then a direct-form FIR (9 channels x 256 taps, both ears as one float2) produces the (L, 2) RIR in the
// interleaved layout the convolution kernels read.
// ---------------------------------------------------------------------------
constexpr int SH_CH = 9, SH_TAPS = 256, SH_DELAY = 128;
constexpr int SH_THREADS = 128, SH_PER_THREAD = 8, SH_TILE = SH_THREADS * SH_PER_THREAD;   // outputs per CTA

// grid (B); block 256: g[env][k][tau] (float2 = L,R)
__global__ void __launch_bounds__(SH_TAPS)
sh_filters_kernel(const float* __restrict__ hbank, const float* __restrict__ az_deg, float2* __restrict__ g) {
    const int env = blockIdx.x, tau = threadIdx.x;
    const float al = az_deg[env] * 0.017453292519943295f;
    float c1, s1, c2, s2;
    sincosf(al, &s1, &c1);
    sincosf(2.f * al, &s2, &c2);
    float2 h[SH_CH];
#pragma unroll
    for (int j = 0; j < SH_CH; ++j)
        h[j] = make_float2(__ldg(hbank + (j * 2 + 0) * SH_TAPS + tau), __ldg(hbank + (j * 2 + 1) * SH_TAPS + tau));
    // (R a)_j = sum_k R[j][k] a_k  =>  g_k = sum_j R[j][k] h_j; pairs (neg, pos, m): (1,3,1), (5,7,1), (4,8,2)
    // R[pos][pos] = c, R[pos][neg] = -s, R[neg][pos] = s, R[neg][neg] = c
    float2 o[SH_CH];
    o[0] = h[0]; o[2] = h[2]; o[6] = h[6];
    auto mix = [&](int neg, int pos, float c, float s) {
        o[neg] = make_float2(c * h[neg].x - s * h[pos].x, c * h[neg].y - s * h[pos].y);     // column neg: R[neg][neg] h_neg + R[pos][neg] h_pos
        o[pos] = make_float2(s * h[neg].x + c * h[pos].x, s * h[neg].y + c * h[pos].y);     // column pos: R[neg][pos] h_neg + R[pos][pos] h_pos
    };
    mix(1, 3, c1, s1);
    mix(5, 7, c1, s1);
    mix(4, 8, c2, s2);
#pragma unroll
    for (int k = 0; k < SH_CH; ++k) g[((long long)env * SH_CH + k) * SH_TAPS + tau] = o[k];
}

// grid (ceil(L / SH_TILE), B); block SH_THREADS.  amb[env][L][9] f32 -> out[env][L][2]
__global__ void __launch_bounds__(SH_THREADS)
sh_decode_kernel(const float* __restrict__ amb, int L, const float2* __restrict__ g, float* __restrict__ out) {
    extern __shared__ float sh_smem[];
    constexpr int SPAN = SH_TILE + SH_DELAY + SH_TAPS - 1;      // input samples feeding one tile
    float* sa = sh_smem;                                          // [9][SPAN]
    float2* sg = reinterpret_cast<float2*>(sh_smem + SH_CH * SPAN + (SH_CH * SPAN & 1));   // [9][256]
    const int env = blockIdx.y;
    const int n0 = blockIdx.x * SH_TILE;
    const int first = n0 - SH_DELAY - (SH_TAPS - 1);              // input index of sa[.][0]
    const float* __restrict__ a = amb + (long long)env * L * SH_CH;
    for (int i = threadIdx.x; i < SPAN * SH_CH; i += SH_THREADS) {
        const int pos = i / SH_CH, k = i - pos * SH_CH;            // coalesced over the interleaved (L, 9) input
        const int n = first + pos;
        sa[k * SPAN + pos] = (n >= 0 && n < L) ? __ldg(a + (long long)n * SH_CH + k) : 0.f;
    }
    for (int i = threadIdx.x; i < SH_CH * SH_TAPS; i += SH_THREADS) sg[i] = __ldg(g + (long long)env * SH_CH * SH_TAPS + i);
    __syncthreads();
    float2 acc[SH_PER_THREAD];
#pragma unroll
    for (int j = 0; j < SH_PER_THREAD; ++j) acc[j] = make_float2(0.f, 0.f);
    // thread t owns outputs n0 + t + j*SH_THREADS (j < 8): out[n] = sum_k sum_tau a_k[n - 128 - tau] g_k[tau];
    // in tile coordinates a_k[n - 128 - tau] = sa[k][(n - n0) + 255 - tau]
    for (int k = 0; k < SH_CH; ++k) {
        const float* __restrict__ ak = sa + k * SPAN + threadIdx.x + (SH_TAPS - 1);
        const float2* __restrict__ gk = sg + k * SH_TAPS;
#pragma unroll 4
        for (int tau = 0; tau < SH_TAPS; ++tau) {
            const float2 w = gk[tau];                             // broadcast
#pragma unroll
            for (int j = 0; j < SH_PER_THREAD; ++j) {
                const float x = ak[j * SH_THREADS - tau];         // conflict-free: consecutive lanes
                acc[j].x = fmaf(x, w.x, acc[j].x);
                acc[j].y = fmaf(x, w.y, acc[j].y);
            }
        }
    }
    float2* __restrict__ o = reinterpret_cast<float2*>(out) + (long long)env * L;
#pragma unroll
    for (int j = 0; j < SH_PER_THREAD; ++j) {
        const int n = n0 + threadIdx.x + j * SH_THREADS;
        if (n < L) o[n] = acc[j];
    }
}

// FFT version of the same decode (default for long responses): overlap-save with N = 4096 blocks.  The filters are
// 384 taps long including their 128-sample delay, so a block yields N - 384 = 3712 valid outputs (91 % efficient,
// against 50 % for the uniformly partitioned RIR convolution).  Per block: 9 forward transforms of the (real)
// ambisonic channels, multiply-accumulate with the 9 filter spectra G_k = FFT(delayed g_k) (both ears packed:
// real input x complex filter = left + i right), one inverse transform.  7x fewer flops than the direct form.
constexpr int SHF_LOG2N = 12;
constexpr int SHF_F = SH_DELAY + SH_TAPS;                          // 384
constexpr int SHF_HOP = (1 << SHF_LOG2N) - SHF_F;                  // 3712

// grid (9, B); block 256: G[env][k][slot]
__global__ void __launch_bounds__(FftPlan<SHF_LOG2N>::T)
sh_filter_fft_kernel(const float2* __restrict__ g, float2* __restrict__ G, const float2* __restrict__ tw) {
    using P = FftPlan<SHF_LOG2N>;
    extern __shared__ float2 smem[];
    float2* stw = smem + P::SMEM_ELEMS;
    const int k = blockIdx.x, env = blockIdx.y, t = threadIdx.x;
    const Tw6 w0 = load_tw6<true>(tw, P::T, t);
    stage_small_twiddles<SHF_LOG2N>(stw, tw, t);
    const float2* __restrict__ gk = g + ((long long)env * SH_CH + k) * SH_TAPS;
    float2 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = t + q * P::T;
        v[q] = (n >= SH_DELAY && n < SHF_F) ? __ldg(gk + n - SH_DELAY) : make_float2(0.f, 0.f);
    }
    fft_forward<SHF_LOG2N>(v, t, smem, w0, stw);
    store_slots<SHF_LOG2N>(G + ((long long)env * SH_CH + k) * P::N, v, t);
}

// grid (ceil(L / SHF_HOP), B); block 256
__global__ void __launch_bounds__(FftPlan<SHF_LOG2N>::T, 2)
sh_ols_kernel(const float* __restrict__ amb, int L, const float2* __restrict__ G, float* __restrict__ out,
              const float2* __restrict__ tw) {
    using P = FftPlan<SHF_LOG2N>;
    extern __shared__ float2 smem[];
    float2* stw = smem + P::SMEM_ELEMS;
    const int b = blockIdx.x, env = blockIdx.y, t = threadIdx.x;
    const Tw6 w0 = load_tw6<true>(tw, P::T, t);
    stage_small_twiddles<SHF_LOG2N>(stw, tw, t);
    const long long base = (long long)b * SHF_HOP - SHF_F;          // first input sample of this block's window
    const float* __restrict__ a = amb + (long long)env * L * SH_CH;
    const float2* __restrict__ Ge = G + (long long)env * SH_CH * P::N + t;
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
    for (int k = 0; k < SH_CH; ++k) {
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const long long n = base + t + q * P::T;
            v[q] = make_float2((n >= 0 && n < L) ? __ldg(a + n * SH_CH + k) : 0.f, 0.f);
        }
        if (k > 0) __syncthreads();                                 // the exchange buffer is reused by the next transform
        fft_forward<SHF_LOG2N>(v, t, smem, w0, stw);
        const float2* __restrict__ gk = Ge + (long long)k * P::N;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = cfma(v[i], __ldg(gk + i * P::T), acc[i]);
    }
    __syncthreads();
    fft_inverse<SHF_LOG2N>(acc, t, smem, tw, stw);
    constexpr float scale = 1.0f / (float)P::N;
    float2* __restrict__ o = reinterpret_cast<float2*>(out) + (long long)env * L;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int nl = t + q * P::T;                                // valid outputs: nl >= 384
        const long long n = (long long)b * SHF_HOP + nl - SHF_F;
        if (nl >= SHF_F && n < L) o[n] = make_float2(acc[q].x * scale, acc[q].y * scale);
    }
}

// ---------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------
__global__ void crossfade_kernel(const float* __restrict__ prev, float* __restrict__ cur, int n_fade,
                                 long long wave_stride, const uint8_t* __restrict__ enable) {
    const int row = blockIdx.y;                  // env*2 + ear
    if (enable && !enable[row >> 1]) return;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m > n_fade) return;
    // continuous_simulator.py:48-51: x2_weight = arange(n+1)/n ; x1_weight = flip(x2_weight)
    const float w2 = (float)m / (float)n_fade;
    const float w1 = (float)(n_fade - m) / (float)n_fade;
    const long long o = (long long)row * wave_stride + m;
    cur[o] = prev[o] * w1 + cur[o] * w2;
}

// Intensity sensor of AV-WaN (ss_baselines/av_wan/avwan_sensors.py:91-100): onset = the first sample (min
// over ears) exceeding 10 % of the clip maximum; result = mean square of the num_frame samples after it.
// grid (B); block 256
__global__ void __launch_bounds__(256)
intensity_kernel(const float* __restrict__ wave, long long wave_stride, int sr, int num_frame, float* __restrict__ out) {
    __shared__ float sf[256];
    __shared__ int si[256];
    const int env = blockIdx.x, t = threadIdx.x;
    const float* __restrict__ yl = wave + (long long)env * 2 * wave_stride;
    const float* __restrict__ yr = yl + wave_stride;
    float m = -INFINITY;
    for (int n = t; n < sr; n += 256) m = fmaxf(m, fmaxf(yl[n], yr[n]));
    sf[t] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (t < s) sf[t] = fmaxf(sf[t], sf[t + s]); __syncthreads(); }
    const float thr = 0.1f * sf[0];
    __syncthreads();
    // first index per ear with value > thr (np.argmax of an all-False row is 0), then min over ears
    int fl = sr, fr = sr;
    for (int n = t; n < sr; n += 256) {
        if (yl[n] > thr && n < fl) fl = n;
        if (yr[n] > thr && n < fr) fr = n;
    }
    si[t] = fl;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (t < s) si[t] = min(si[t], si[t + s]); __syncthreads(); }
    const int first_l = si[0] == sr ? 0 : si[0];
    __syncthreads();
    si[t] = fr;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (t < s) si[t] = min(si[t], si[t + s]); __syncthreads(); }
    const int first_r = si[0] == sr ? 0 : si[0];
    const int idx = min(first_l, first_r);
    const int end = min(idx + num_frame, sr);               // numpy slicing clips at the end of the clip
    float acc = 0.f;
    for (int n = idx + t; n < end; n += 256) acc += yl[n] * yl[n] + yr[n] * yr[n];
    __syncthreads();
    sf[t] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (t < s) sf[t] += sf[t + s]; __syncthreads(); }
    if (t == 0) out[env] = end > idx ? sf[0] / (float)(2 * (end - idx)) : nanf("");
}

// ---------------------------------------------------------------------------
// the first layer of the policy's audio encoder fused with the layout change.  AudioCNN.forward
// (ss_baselines/av_nav/models/audio_cnn.py:79-89) permutes the (N, 65, T', 2) observation to (N, 2, 65, T') and runs
// Conv2d(2 -> 32, 8x8 / 4, or 5x5 / 2 for inputs under 30 pixels) + ReLU; here the convolution reads the observation
// in the layout the spectrogram kernel wrote it (ears interleaved) and writes (N, OC, H1, W1): no permuted copy, one
// launch.  Inference only (rollout collection runs under no_grad, ppo_trainer.py:131-146); the PPO update keeps
// PyTorch's layer.  grid (H1, N); block (W1 * OC) <= 1024; dynamic smem: KH input rows + all weights.
// ---------------------------------------------------------------------------
__global__ void audio_conv1_kernel(const float* __restrict__ spec, int H, int W, const float* __restrict__ weight,
                                   const float* __restrict__ bias, int OC, int KH, int KW, int SH, int SW, int H1, int W1,
                                   int relu, float* __restrict__ out) {
    extern __shared__ float c1_smem[];
    float* rows = c1_smem;                                  // [KH][W][2]
    float* wsm = c1_smem + KH * W * 2;                      // [OC][2][KH][KW]
    const int oh = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
    const float* __restrict__ in = spec + ((long long)n * H + (long long)oh * SH) * W * 2;
    for (int i = tid; i < KH * W * 2; i += blockDim.x) rows[i] = __ldg(in + i);
    for (int i = tid; i < OC * 2 * KH * KW; i += blockDim.x) wsm[i] = __ldg(weight + i);
    __syncthreads();
    const int ow = tid % W1, oc = tid / W1;
    if (oc >= OC) return;
    float acc = bias ? __ldg(bias + oc) : 0.f;
    const float* __restrict__ w0 = wsm + (oc * 2 + 0) * KH * KW;
    const float* __restrict__ w1 = wsm + (oc * 2 + 1) * KH * KW;
    for (int kh = 0; kh < KH; ++kh) {
        const float* __restrict__ r = rows + (kh * W + ow * SW) * 2;
        for (int kw = 0; kw < KW; ++kw) {
            acc = fmaf(r[2 * kw], w0[kh * KW + kw], acc);           // channel 0 = left ear
            acc = fmaf(r[2 * kw + 1], w1[kh * KW + kw], acc);       // channel 1 = right ear
        }
    }
    if (relu) acc = fmaxf(acc, 0.f);
    out[(((long long)n * OC + oc) * H1 + oh) * W1 + ow] = acc;
}

__global__ void pcm16_decode_kernel(const int16_t* __restrict__ in, long long n, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) out[i] = (float)in[i] * (1.0f / 32768.0f);
}

__global__ void pcm16_encode_kernel(const float* __restrict__ in, long long n, int mode, int16_t* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        float x = in[i];
        float v = mode == 0 ? rintf(x * 32768.0f) : truncf(x * 32767.0f);
        v = fminf(fmaxf(v, -32768.0f), 32767.0f);
        out[i] = (int16_t)(int)v;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int log2_supported(int l) {
    for (int s : kSupportedLog2)
        if (s == l) return 1;
    return 0;
}

template <int LOG2N>
static int setup_smem_attrs() {
    using P = FftPlan<LOG2N>;
    const int bytes = (P::SMEM_ELEMS + P::TW_SMALL_ELEMS) * (int)sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(fwd_rir_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fwd_src_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mac_ifft_kernel<LOG2N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mac_ifft_kernel<LOG2N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e == cudaSuccess ? 0 : -1;
}

// twiddle table of FftPlan<LOG2N>: pass p, rows e in {1,2,3,4,8,12}: [row][j] = exp(-2 pi i j e / (16 stride(p))),
// computed in double and rounded once
template <int LOG2N>
static cudaError_t upload_twiddles(ssb_ctx* ctx) {
    using P = FftPlan<LOG2N>;
    static const int kExp[6] = {1, 2, 3, 4, 8, 12};
    std::vector<float2> h(P::TW_ELEMS);
    for (int p = 0; p < P::NPASS; ++p) {
        const int st = P::stride(p);
        for (int row = 0; row < 6; ++row)
            for (int j = 0; j < st; ++j) {
                const double a = -2.0 * M_PI * (double)j * (double)kExp[row] / (16.0 * (double)st);
                h[P::tw_offset(p) + row * st + j] = make_float2((float)cos(a), (float)sin(a));
            }
    }
    cudaError_t e = cudaMalloc(&ctx->tw[LOG2N], h.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->tw[LOG2N], h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
    return e;
}

extern "C" int ssb_version(void) { return 100; }

static int create_impl(ssb_ctx* ctx, int device);

extern "C" int ssb_create(int device, ssb_ctx** out) {
    if (!out) return SSB_E_INVALID_ARG;
    *out = nullptr;
    ssb_ctx* ctx = new (std::nothrow) ssb_ctx();
    if (!ctx) return SSB_E_OOM;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->timed = new (std::nothrow) std::vector<TimedLaunch>();
    ctx->mel = new (std::nothrow) std::vector<ssb_ctx::MelBank>();
    *out = ctx;   // returned even on failure so that ssb_last_error works; caller destroys
    if (!ctx->timed || !ctx->mel) SSB_FAIL(ctx, SSB_E_OOM, "out of host memory");
    return abi_call(ctx, [&] { return create_impl(ctx, device); });   // the caller's current device is restored
}

static int create_impl(ssb_ctx* ctx, int device) {
    {   // abi_call's guard has already selected the device; surface a bad ordinal as an error
        int cur = -1;
        SSB_CUDA(ctx, cudaGetDevice(&cur));
        if (cur != device) SSB_CUDA(ctx, cudaSetDevice(device));
    }
    cudaDeviceProp prop;
    SSB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) SSB_FAIL(ctx, SSB_E_CUDA, "device %d is sm_%d%d; libssb200 is built for sm_100a only", device, prop.major, prop.minor);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->n_streams = 2;
    ctx->n_chunks = 1;
    {
        cudaError_t e = cudaSuccess;
        if (e == cudaSuccess) e = upload_twiddles<9>(ctx);
        if (e == cudaSuccess) e = upload_twiddles<12>(ctx);
        if (e == cudaSuccess) e = upload_twiddles<13>(ctx);
        if (e == cudaSuccess) e = upload_twiddles<14>(ctx);
        SSB_CUDA(ctx, e);
    }
    {
        float hw[SSB_N_FFT];
        const int lpad = (SSB_N_FFT - SSB_WIN) / 2;
        for (int i = 0; i < SSB_N_FFT; ++i) hw[i] = 0.f;
        // scipy.signal.get_window('hann', 400, fftbins=True): 0.5 - 0.5 cos(2 pi n / 400)
        for (int n = 0; n < SSB_WIN; ++n) hw[lpad + n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)SSB_WIN));
        SSB_CUDA(ctx, cudaMalloc(&ctx->window, sizeof(hw)));
        SSB_CUDA(ctx, cudaMemcpy(ctx->window, hw, sizeof(hw), cudaMemcpyHostToDevice));
    }
    {   // single-block plan: w_64 constants (per-device constant memory), the [16][1024] modulation table, smem opt-in
        float2 w64[64];
        for (int k = 0; k < 64; ++k) {
            const double a = -2.0 * M_PI * (double)k / 64.0;
            w64[k] = make_float2((float)cos(a), (float)sin(a));
        }
        SSB_CUDA(ctx, cudaMemcpyToSymbol(kW64, w64, sizeof(w64)));
        std::vector<float2> twm(C64_TWM_ELEMS);
        for (int r = 0; r < 16; ++r)
            for (int tau = 0; tau < 1024; ++tau) {
                const double a = -2.0 * M_PI * (double)((r * tau) % C64_M) / (double)C64_M;
                twm[r * 1024 + tau] = make_float2((float)cos(a), (float)sin(a));
            }
        SSB_CUDA(ctx, cudaMalloc(&ctx->twm64, twm.size() * sizeof(float2)));
        SSB_CUDA(ctx, cudaMemcpy(ctx->twm64, twm.data(), twm.size() * sizeof(float2), cudaMemcpyHostToDevice));
        SSB_CUDA(ctx, cudaFuncSetAttribute(conv64k_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM_BYTES));
        SSB_CUDA(ctx, cudaFuncSetAttribute(conv64k_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM_BYTES));
        // groups of 256 threads per CTA of the single-block kernel: 4 (default) takes the whole SM; 2 leaves half of the
        // SM's registers and warp slots to co-resident kernels of other streams.  Measured (profiles/times_r02b_block64_groups.log):
        // the kernel takes the SAME time with half the threads (96.8 vs 92.8 us per 128 envs) -- it is bound by the chain of
        // phases of the ~33 envs in flight, not by warps -- and the freed half of the SM did not buy overlap (121.7 vs 119.4 us per step)
        const char* g = getenv("SSB200_C64_GROUPS");
        ctx->c64_groups = (g && g[0] == '2') ? 2 : 4;
    }
    if (setup_smem_attrs<12>() || setup_smem_attrs<13>() || setup_smem_attrs<14>())
        SSB_FAIL(ctx, SSB_E_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: %s",
                 cudaGetErrorString(cudaGetLastError()));
    return SSB_OK;
}

extern "C" void ssb_destroy(ssb_ctx* ctx) {
    if (!ctx) return;
    DevGuard guard(ctx->device);
    for (int l = 0; l < 16; ++l)
        if (ctx->tw[l]) cudaFree(ctx->tw[l]);
    if (ctx->window) cudaFree(ctx->window);
    if (ctx->twm64) cudaFree(ctx->twm64);
    if (ctx->yscratch) cudaFree(ctx->yscratch);
    if (ctx->gscratch) cudaFree(ctx->gscratch);
    if (ctx->s_comp[0]) {
        for (int i = 0; i < SSB_MAX_STREAMS; ++i) { cudaStreamDestroy(ctx->s_comp[i]); cudaEventDestroy(ctx->ev_comp[i]); }
        cudaEventDestroy(ctx->ev_fork);
    }
    if (ctx->s_h2d) {
        cudaStreamDestroy(ctx->s_h2d);
        cudaStreamDestroy(ctx->s_d2h);
        for (int i = 0; i < 2 * SSB_MAX_CHUNKS + 2; ++i)
            if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
        for (int i = 0; i < 2; ++i)
            if (ctx->stage_done[i]) cudaEventDestroy(ctx->stage_done[i]);
    }
    if (ctx->timed) {
        for (auto& tl : *ctx->timed) { cudaEventDestroy(tl.a); cudaEventDestroy(tl.b); }
        delete ctx->timed;
    }
    if (ctx->mel) {
        for (auto& mb : *ctx->mel) { cudaFree(mb.rows); cudaFree(mb.ofs); cudaFree(mb.w); }
        delete ctx->mel;
    }
    delete ctx;
}

extern "C" const char* ssb_last_error(const ssb_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" int64_t ssb_launch_count(const ssb_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ssb_set_conv_mode(ssb_ctx* ctx, int mode) {
    if (!ctx || (mode != 0 && mode != 1)) return SSB_E_INVALID_ARG;
    ctx->conv_mode = mode;
    return SSB_OK;
}

extern "C" int ssb_set_streams(ssb_ctx* ctx, int n) {
    if (!ctx || n < 1 || n > SSB_MAX_STREAMS) return SSB_E_INVALID_ARG;
    ctx->n_streams = n;
    return SSB_OK;
}

extern "C" int ssb_set_chunks(ssb_ctx* ctx, int n) {
    if (!ctx || n < 1 || n > 16) return SSB_E_INVALID_ARG;
    ctx->n_chunks = n;
    return SSB_OK;
}

extern "C" int ssb_set_debug(ssb_ctx* ctx, int flags) {
    if (!ctx) return SSB_E_INVALID_ARG;
    ctx->debug = flags;
    return SSB_OK;
}

extern "C" int ssb_set_kernel_timing(ssb_ctx* ctx, int enable) {
    if (!ctx) return SSB_E_INVALID_ARG;
    ctx->timing = enable ? 1 : 0;
    return SSB_OK;
}

static int ssb_get_kernel_timing_impl(ssb_ctx* ctx, double* ms_sum, int64_t* counts) {
    if (!ctx || !ms_sum || !counts || !ctx->timed) return SSB_E_INVALID_ARG;
    for (int k = 0; k < K_COUNT; ++k) { ms_sum[k] = 0.0; counts[k] = 0; }
    for (auto& tl : *ctx->timed) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(tl.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, tl.a, tl.b);
        cudaEventDestroy(tl.a);
        cudaEventDestroy(tl.b);
        if (e != cudaSuccess) { ctx->timed->clear(); SSB_CUDA(ctx, e); }
        ms_sum[tl.kernel] += ms;
        counts[tl.kernel] += 1;
    }
    ctx->timed->clear();
    return SSB_OK;
}

extern "C" int ssb_spec_cols(int sr) {
    const int frames = 1 + sr / SSB_HOP;
    return (frames + SSB_POOL - 1) / SSB_POOL;
}

extern "C" int ssb_make_plan(ssb_ctx* ctx, int sr, int max_taps, int n_terms, int log2n, ssb_plan* plan) {
    if (!ctx || !plan) return SSB_E_INVALID_ARG;
    // default block size: N = 4096 keeps 4 CTAs per SM and the finest tiling; beyond ~12 partitions the per-bin
    // partition sums dominate and N = 8192 is the better plan (measured at config 3)
    if (log2n == 0) log2n = max_taps > 24576 ? 13 : 12;
    if (log2n == C64_LOG2M) {
        // single-block plan: out[m] sits at circular index m + D, D = 65536 - sr; alias-free for taps <= D + 1
        if (n_terms != 1 || sr < SSB_N_FFT || sr > C64_M - C64_NS || max_taps < 0 || max_taps > C64_M - sr + 1)
            SSB_FAIL(ctx, SSB_E_INVALID_ARG, "single-block plan needs n_terms = 1, sr <= %d and max_taps <= 65536 - sr + 1 "
                     "(sr=%d max_taps=%d n_terms=%d)", C64_M - C64_NS, sr, max_taps, n_terms);
        plan->log2n = C64_LOG2M;
        plan->block = C64_M - sr;
        plan->sr = sr;
        plan->n_blocks = 1;
        plan->max_parts = 1;
        plan->n_terms = 1;
        plan->h_elems_per_env = 0;
        return SSB_OK;
    }
    if (log2n < 12 || !log2_supported(log2n)) SSB_FAIL(ctx, SSB_E_INVALID_ARG, "log2n %d unsupported (12, 13, 14, 16)", log2n);
    if (sr < SSB_N_FFT || max_taps < 0 || n_terms < 1 || n_terms > 2)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "bad plan arguments sr=%d max_taps=%d n_terms=%d", sr, max_taps, n_terms);
    const int P = (1 << log2n) / 2;
    plan->log2n = log2n;
    plan->block = P;
    plan->sr = sr;
    plan->n_blocks = (sr + P - 1) / P;
    plan->max_parts = max_taps > 0 ? (max_taps + P - 1) / P : 1;
    plan->n_terms = n_terms;
    plan->h_elems_per_env = (int64_t)n_terms * plan->max_parts * (1 << log2n);
    return SSB_OK;
}

static int check_plan(ssb_ctx* ctx, const ssb_plan* plan) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (plan && plan->log2n == C64_LOG2M) {
        if (plan->n_terms != 1 || plan->max_parts != 1 || plan->n_blocks != 1 || plan->sr < SSB_N_FFT ||
            plan->sr > C64_M - C64_NS || plan->block != C64_M - plan->sr)
            SSB_FAIL(ctx, SSB_E_INVALID_ARG, "invalid single-block plan");
        return SSB_OK;
    }
    if (!plan || plan->log2n < 12 || !log2_supported(plan->log2n) || plan->block != (1 << plan->log2n) / 2 ||
        plan->n_terms < 1 || plan->n_terms > 2 || plan->max_parts < 1 ||
        plan->n_blocks != (plan->sr + plan->block - 1) / plan->block)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "invalid plan");
    return SSB_OK;
}

template <int LOG2N>
static cudaError_t launch_src(ssb_ctx* ctx, const float* d_src, int S, int64_t m0, int wrap, int nw, int wofs,
                              float2* d_x, cudaStream_t st) {
    using P = FftPlan<LOG2N>;
    {
        LaunchTimer lt(ctx, K_FWD_SRC, st);
        fwd_src_kernel<LOG2N><<<nw, P::T, (P::SMEM_ELEMS + P::TW_SMALL_ELEMS) * sizeof(float2), st>>>(d_src, S, (long long)m0, wrap, wofs, d_x, ctx->tw[LOG2N]);
    }
    return cudaGetLastError();
}

static int ssb_source_windows_impl(ssb_ctx* ctx, const ssb_plan* plan, const float* d_src, int S, int64_t m0, int wrap,
                                  int nw, int wofs, void* d_x, void* stream) {
    int rc = check_plan(ctx, plan);
    if (rc) return rc;
    if (!d_src || !d_x || S <= 0 || nw <= 0 || wofs < 0 || wofs >= nw)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_source_windows: bad arguments (S=%d nw=%d wofs=%d)", S, nw, wofs);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (plan->log2n == C64_LOG2M) {
        if (nw != 1 || wofs != 0) SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_source_windows: the single-block plan has one spectrum (nw = 1, wofs = 0)");
        {
            LaunchTimer lt(ctx, K_FWD_SRC, st);
            src64k_kernel<<<16, 256, C64_BUF * sizeof(float2), st>>>(d_src, S, (long long)m0, wrap, plan->block, (float2*)d_x,
                                                                    ctx->tw[12], ctx->twm64);
        }
        SSB_CUDA(ctx, cudaGetLastError());
        return SSB_OK;
    }
    switch (plan->log2n) {
        case 12: e = launch_src<12>(ctx, d_src, S, m0, wrap, nw, wofs, (float2*)d_x, st); break;
        case 13: e = launch_src<13>(ctx, d_src, S, m0, wrap, nw, wofs, (float2*)d_x, st); break;
        default: e = launch_src<14>(ctx, d_src, S, m0, wrap, nw, wofs, (float2*)d_x, st); break;
    }
    SSB_CUDA(ctx, e);
    return SSB_OK;
}

template <int LOG2N>
static cudaError_t launch_conv_sub(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs, const float* d_rir_bank,
                                   const void* d_xpool, void* d_h, float2* y, float* d_wave, int64_t wave_stride,
                                   cudaStream_t st) {
    using P = FftPlan<LOG2N>;
    const size_t smem = (P::SMEM_ELEMS + P::TW_SMALL_ELEMS) * sizeof(float2);
    dim3 g1(plan->max_parts * plan->n_terms, B);
    {
        LaunchTimer lt(ctx, K_FWD_RIR, st);
        fwd_rir_kernel<LOG2N><<<g1, P::T, smem, st>>>(d_reqs, (const float2*)d_rir_bank, (float2*)d_h, plan->max_parts,
                                                      (long long)plan->h_elems_per_env, ctx->tw[LOG2N]);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 g2(B, plan->n_blocks);
    if (ctx->conv_mode == 0) {
        {
            LaunchTimer lt(ctx, K_MAC_BINS, st);
            mac_bins_kernel<LOG2N, 8, 24><<<dim3(P::N / 256, B), 256, 0, st>>>(
                d_reqs, (const float2*)d_xpool, (const float2*)d_h, plan->max_parts, plan->n_terms,
                (long long)plan->h_elems_per_env, y, plan->n_blocks, plan->sr);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        LaunchTimer lt(ctx, K_MAC_IFFT, st);
        mac_ifft_kernel<LOG2N, true><<<g2, P::T, smem, st>>>(d_reqs, (const float2*)d_xpool, y, plan->max_parts,
                                                             plan->n_terms, (long long)plan->h_elems_per_env, d_wave,
                                                             (long long)wave_stride, plan->sr, ctx->tw[LOG2N], ctx->debug);
    } else {
        LaunchTimer lt(ctx, K_MAC_IFFT, st);
        mac_ifft_kernel<LOG2N, false><<<g2, P::T, smem, st>>>(d_reqs, (const float2*)d_xpool, (const float2*)d_h, plan->max_parts,
                                                              plan->n_terms, (long long)plan->h_elems_per_env, d_wave,
                                                              (long long)wave_stride, plan->sr, ctx->tw[LOG2N], ctx->debug);
    }
    return cudaGetLastError();
}

// partition-sum scratch Y[B][n_blocks][N] (mode 0), owned by the context
static cudaError_t ensure_yscratch(ssb_ctx* ctx, const ssb_plan* plan, int B, cudaStream_t st) {
    if (ctx->conv_mode != 0 || plan->log2n == C64_LOG2M) return cudaSuccess;
    const size_t need = (size_t)B * plan->n_blocks * ((size_t)1 << plan->log2n);
    if (need <= ctx->yscratch_elems) return cudaSuccess;
    if (ctx->yscratch) {
        cudaDeviceSynchronize();
        cudaFree(ctx->yscratch);
        ctx->yscratch = nullptr;
        ctx->yscratch_elems = 0;
    }
    (void)st;
    cudaError_t e = cudaMalloc(&ctx->yscratch, need * sizeof(float2));
    if (e == cudaSuccess) ctx->yscratch_elems = need;
    return e;
}

static cudaError_t launch_conv_any(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs, const float* d_rir_bank,
                                   const void* d_xpool, void* d_h, float2* y, float* d_wave, int64_t wave_stride,
                                   cudaStream_t st) {
    if (plan->log2n == C64_LOG2M) {
        LaunchTimer lt(ctx, K_CONV64K, st);
        if (ctx->c64_groups == 2)
            conv64k_kernel<2><<<C64_CL * B, 512, C64_SMEM_BYTES, st>>>(d_reqs, (const float2*)d_rir_bank, (const float2*)d_xpool,
                                                                        d_wave, (long long)wave_stride, plan->sr, ctx->tw[12], ctx->twm64);
        else
            conv64k_kernel<4><<<C64_CL * B, 1024, C64_SMEM_BYTES, st>>>(d_reqs, (const float2*)d_rir_bank, (const float2*)d_xpool,
                                                                         d_wave, (long long)wave_stride, plan->sr, ctx->tw[12], ctx->twm64);
        return cudaGetLastError();
    }
    switch (plan->log2n) {
        case 12: return launch_conv_sub<12>(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_h, y, d_wave, wave_stride, st);
        case 13: return launch_conv_sub<13>(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_h, y, d_wave, wave_stride, st);
        default: return launch_conv_sub<14>(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_h, y, d_wave, wave_stride, st);
    }
}

static int ssb_convolve_batch_impl(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs,
                                  const float* d_rir_bank, const void* d_xpool, void* d_hscratch, float* d_wave,
                                  int64_t wave_stride, void* stream) {
    int rc = check_plan(ctx, plan);
    if (rc) return rc;
    if (B == 0) return SSB_OK;
    if (B < 0 || B > 16383 * 4 || !d_reqs || !d_rir_bank || !d_xpool || (!d_hscratch && plan->h_elems_per_env) || !d_wave || wave_stride < plan->sr)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_convolve_batch: bad arguments (B=%d wave_stride=%lld sr=%d)", B,
                 (long long)wave_stride, plan->sr);
    if (((uintptr_t)d_rir_bank & 7) || ((uintptr_t)d_xpool & 7) || ((uintptr_t)d_hscratch & 7))
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_convolve_batch: rir bank / spectra must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    SSB_CUDA(ctx, ensure_yscratch(ctx, plan, B, st));
    SSB_CUDA(ctx, launch_conv_any(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_hscratch, ctx->yscratch, d_wave, wave_stride, st));
    return SSB_OK;
}

static int ssb_crossfade_batch_impl(ssb_ctx* ctx, int B, const float* d_prev, float* d_cur, int sr, int64_t wave_stride,
                                   const uint8_t* d_enable, void* stream) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (B == 0) return SSB_OK;
    if (B < 0 || B > 32767 || !d_prev || !d_cur || sr < 20 || wave_stride < sr)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_crossfade_batch: bad arguments");
    const int n_fade = (int)(0.05 * (double)sr);       // int(0.05 * sr), continuous_simulator.py:48
    dim3 g((n_fade + 1 + 255) / 256, B * 2);
    crossfade_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(d_prev, d_cur, n_fade, (long long)wave_stride, d_enable);
    ctx->launches += 1;
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}

static int spectrogram_checked(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int pad_mode,
                               float* d_spec, cudaStream_t st);

static int ssb_spectrogram_batch_impl(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int pad_mode,
                                     float* d_spec, void* stream) {
    return spectrogram_checked(ctx, B, d_wave, wave_stride, sr, pad_mode, d_spec, (cudaStream_t)stream);
}

static int spectrogram_checked(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int pad_mode,
                               float* d_spec, cudaStream_t stream) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (B == 0) return SSB_OK;
    if (B < 0 || B > 65535 || !d_wave || !d_spec || sr < SSB_N_FFT || wave_stride < sr ||
        ((pad_mode & ~SSB_LAYOUT_NCHW) != SSB_PAD_REFLECT && (pad_mode & ~SSB_LAYOUT_NCHW) != SSB_PAD_CONSTANT))
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_spectrogram_batch: bad arguments (B=%d sr=%d pad_mode=%d)", B, sr, pad_mode);
    const int nchw = (pad_mode & SSB_LAYOUT_NCHW) ? 1 : 0;
    pad_mode &= ~SSB_LAYOUT_NCHW;
    const int frames = 1 + sr / SSB_HOP;
    const int cols = ssb_spec_cols(sr);
    dim3 g(cols, B);
    {
        LaunchTimer lt(ctx, K_SPECTROGRAM, (cudaStream_t)stream);
        if (pad_mode == SSB_PAD_REFLECT)
            spectrogram_kernel<true><<<g, 32, 0, (cudaStream_t)stream>>>(
                d_wave, (long long)wave_stride, sr, frames, cols, d_spec, ctx->tw[9], ctx->window, ctx->debug, nchw);
        else
            spectrogram_kernel<false><<<g, 32, 0, (cudaStream_t)stream>>>(
                d_wave, (long long)wave_stride, sr, frames, cols, d_spec, ctx->tw[9], ctx->window, ctx->debug, nchw);
    }
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}

// convolve + spectrogram.  With ssb_set_streams(ctx, S > 1) the batch is split into S sub-batches whose
// kernel chains run on S internal streams forked from / joined to the caller's stream: the tail of one
// kernel (partial last wave) overlaps the next sub-batch's work, and memory-bound kernels of one chain
// overlap compute-bound kernels of another.
static int ssb_render_batch_impl(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs, const float* d_rir_bank,
                                const void* d_xpool, void* d_hscratch, float* d_wave, int64_t wave_stride, int pad_mode,
                                float* d_spec, void* stream) {
    int S = ctx->n_streams;
    if (S > SSB_MAX_STREAMS) S = SSB_MAX_STREAMS;
    if (S > B / 32) S = B / 32;                       // sub-batches of at least 32 envs
    if (S <= 1) {
        int rc = ssb_convolve_batch_impl(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_hscratch, d_wave, wave_stride, stream);
        if (rc) return rc;
        return ssb_spectrogram_batch_impl(ctx, B, d_wave, wave_stride, plan->sr, pad_mode, d_spec, stream);
    }
    int rc = check_plan(ctx, plan);
    if (rc) return rc;
    if (B > 65535 || !d_reqs || !d_rir_bank || !d_xpool || (!d_hscratch && plan->h_elems_per_env) || !d_wave || !d_spec || wave_stride < plan->sr)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_render_batch: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (!ctx->s_comp[0]) {
        for (int i = 0; i < SSB_MAX_STREAMS; ++i) {
            SSB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_comp[i], cudaStreamNonBlocking));
            SSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
        }
        SSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    }
    SSB_CUDA(ctx, ensure_yscratch(ctx, plan, B, st));
    SSB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    const size_t spec_row = (size_t)SSB_SPEC_ROWS * ssb_spec_cols(plan->sr) * 2;
    const size_t y_env = (size_t)plan->n_blocks * ((size_t)1 << plan->log2n);
    // S streams x C chunks: sub-batch i runs on stream i % S, so a stream works through its chunks one after the
    // other and only S sub-batches' intermediates (H, Y, waveform: 1.3 MB per env at config 2) are live in L2 at a time
    int C = ctx->n_chunks;
    while (C > 1 && B / (S * C) < 16) --C;            // sub-batches of at least 16 envs
    const int n_sub = S * C;
    const int per = (B + n_sub - 1) / n_sub;
    for (int i = 0; i < S; ++i) SSB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comp[i], ctx->ev_fork, 0));
    int i = 0;
    for (int e0 = 0; e0 < B; e0 += per, ++i) {
        const int nb = (B - e0 < per) ? (B - e0) : per;
        cudaStream_t si = ctx->s_comp[i % S];
        SSB_CUDA(ctx, launch_conv_any(ctx, plan, nb, d_reqs + e0, d_rir_bank, d_xpool,
                                      (float2*)d_hscratch + (size_t)e0 * plan->h_elems_per_env,
                                      ctx->yscratch ? ctx->yscratch + (size_t)e0 * y_env : nullptr,
                                      d_wave + (size_t)e0 * 2 * wave_stride, wave_stride, si));
        rc = spectrogram_checked(ctx, nb, d_wave + (size_t)e0 * 2 * wave_stride, wave_stride, plan->sr, pad_mode,
                                 d_spec + (size_t)e0 * spec_row, si);
        if (rc) return rc;
    }
    for (int k = 0; k < S; ++k) {
        SSB_CUDA(ctx, cudaEventRecord(ctx->ev_comp[k], ctx->s_comp[k]));
        SSB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_comp[k], 0));
    }
    return SSB_OK;
}

// Host-buffer entry.  n_chunks > 1 pipelines the batch: chunk c+1's H2D copy (copy stream) overlaps
// chunk c's kernels (caller's stream) and chunk c-1's D2H copy (second copy stream); PCIe is full
// duplex, so a step costs about max(H2D, kernels, D2H) instead of their sum.
static int ssb_render_batch_host_impl(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* h_reqs, const float* h_rir,
                                     int64_t rir_bytes, float* d_rir_staging, ssb_req* d_reqs_staging, const void* d_xpool,
                                     void* d_hscratch, float* d_wave, int64_t wave_stride, int pad_mode, float* d_spec,
                                     float* h_spec, float* h_wave, int n_chunks, void* stream) {
    int rc = check_plan(ctx, plan);
    if (rc) return rc;
    if (B <= 0 || !h_reqs || !h_rir || rir_bytes <= 0 || !d_rir_staging || !d_reqs_staging || !h_spec || !d_spec ||
        !d_wave || wave_stride < plan->sr)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_render_batch_host: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t spec_row = (size_t)SSB_SPEC_ROWS * ssb_spec_cols(plan->sr) * 2;      // floats per env
    if (n_chunks > B) n_chunks = B;
    if (n_chunks > SSB_MAX_CHUNKS) n_chunks = SSB_MAX_CHUNKS;
    if (n_chunks <= 1) {
        SSB_CUDA(ctx, cudaMemcpyAsync(d_rir_staging, h_rir, (size_t)rir_bytes, cudaMemcpyHostToDevice, st));
        SSB_CUDA(ctx, cudaMemcpyAsync(d_reqs_staging, h_reqs, (size_t)B * sizeof(ssb_req), cudaMemcpyHostToDevice, st));
        ctx->last_h2d_bytes = rir_bytes + (int64_t)B * (int64_t)sizeof(ssb_req);
        ctx->last_d2h_bytes = (int64_t)B * (int64_t)spec_row * 4 + (h_wave ? (int64_t)B * 2 * plan->sr * 4 : 0);
        rc = ssb_render_batch_impl(ctx, plan, B, d_reqs_staging, d_rir_staging, d_xpool, d_hscratch, d_wave, wave_stride,
                              pad_mode, d_spec, stream);
        if (rc) return rc;
        SSB_CUDA(ctx, cudaMemcpyAsync(h_spec, d_spec, (size_t)B * spec_row * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (h_wave)
            SSB_CUDA(ctx, cudaMemcpy2DAsync(h_wave, (size_t)plan->sr * sizeof(float), d_wave,
                                            (size_t)wave_stride * sizeof(float), (size_t)plan->sr * sizeof(float),
                                            (size_t)B * 2, cudaMemcpyDeviceToHost, st));
        return SSB_OK;
    }
    if (!ctx->s_h2d) {
        SSB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        SSB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2 * SSB_MAX_CHUNKS + 2; ++i)
            SSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) SSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_done[i], cudaEventDisableTiming));
    }
    cudaEvent_t ev_done = ctx->ev[2 * SSB_MAX_CHUNKS + 1];
    // The H2D copies only have to wait for the kernels that last read THIS staging buffer.  A caller
    // that alternates two staging buffers (HostSession does) therefore gets the next step's copy
    // overlapped with the current step's kernels; with a single buffer the copy waits for them.
    for (int i = 0; i < 2; ++i)
        if (ctx->stage_ptr[i] == (const void*)d_rir_staging || ctx->stage_ptr[i] == (const void*)d_reqs_staging)
            SSB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_h2d, ctx->stage_done[i], 0));
    SSB_CUDA(ctx, cudaMemcpyAsync(d_reqs_staging, h_reqs, (size_t)B * sizeof(ssb_req), cudaMemcpyHostToDevice, ctx->s_h2d));
    ctx->last_h2d_bytes = (int64_t)B * (int64_t)sizeof(ssb_req);
    ctx->last_d2h_bytes = (int64_t)B * (int64_t)spec_row * 4 + (h_wave ? (int64_t)B * 2 * plan->sr * 4 : 0);
    const int64_t bank_taps = rir_bytes / (int64_t)sizeof(float2);
    const int per = (B + n_chunks - 1) / n_chunks;
    int c = 0;
    for (int e0 = 0; e0 < B; e0 += per, ++c) {
        const int nb = (B - e0 < per) ? (B - e0) : per;
        // tap ranges of the host bank this chunk reads: silent envs and zero-tap terms need no RIR at all
        // (the reference returns zeros before it even opens the file, simulator.py:610-612); neighbouring
        // ranges are merged so a chunk costs a handful of copies
        std::vector<std::pair<int64_t, int64_t>> runs;
        for (int e = e0; e < e0 + nb; ++e) {
            if (h_reqs[e].flags & SSB_FLAG_SILENT) continue;
            for (int term = 0; term < plan->n_terms; ++term) {
                const ssb_conv_term& ct = h_reqs[e].term[term];
                if (ct.rir_taps <= 0) continue;
                if (ct.rir_offset < 0 || ct.rir_offset + ct.rir_taps > bank_taps)
                    SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_render_batch_host: request %d reads outside the host RIR buffer", e);
                runs.emplace_back(ct.rir_offset, ct.rir_offset + ct.rir_taps);
            }
        }
        std::sort(runs.begin(), runs.end());
        const int64_t kMergeGap = 2048;                         // taps (16 KB): cheaper to copy than to split
        size_t i0 = 0;
        while (i0 < runs.size()) {
            int64_t lo = runs[i0].first, hi = runs[i0].second;
            size_t i1 = i0 + 1;
            while (i1 < runs.size() && runs[i1].first <= hi + kMergeGap) { if (runs[i1].second > hi) hi = runs[i1].second; ++i1; }
            SSB_CUDA(ctx, cudaMemcpyAsync(d_rir_staging + 2 * lo, h_rir + 2 * lo, (size_t)(hi - lo) * sizeof(float2),
                                          cudaMemcpyHostToDevice, ctx->s_h2d));
            ctx->last_h2d_bytes += (hi - lo) * (int64_t)sizeof(float2);
            i0 = i1;
        }
        SSB_CUDA(ctx, cudaEventRecord(ctx->ev[2 * c], ctx->s_h2d));
        SSB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev[2 * c], 0));
        rc = ssb_render_batch_impl(ctx, plan, nb, d_reqs_staging + e0, d_rir_staging, d_xpool,
                              (float2*)d_hscratch + (size_t)e0 * plan->h_elems_per_env, d_wave + (size_t)e0 * 2 * wave_stride,
                              wave_stride, pad_mode, d_spec + (size_t)e0 * spec_row, stream);
        if (rc) return rc;
        SSB_CUDA(ctx, cudaEventRecord(ctx->ev[2 * c + 1], st));
        SSB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->ev[2 * c + 1], 0));
        SSB_CUDA(ctx, cudaMemcpyAsync(h_spec + (size_t)e0 * spec_row, d_spec + (size_t)e0 * spec_row,
                                      (size_t)nb * spec_row * sizeof(float), cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (h_wave)
            SSB_CUDA(ctx, cudaMemcpy2DAsync(h_wave + (size_t)e0 * 2 * plan->sr, (size_t)plan->sr * sizeof(float),
                                            d_wave + (size_t)e0 * 2 * wave_stride, (size_t)wave_stride * sizeof(float),
                                            (size_t)plan->sr * sizeof(float), (size_t)nb * 2, cudaMemcpyDeviceToHost,
                                            ctx->s_d2h));
    }
    {   // remember who read this staging buffer last
        const int slot = ctx->stage_next;
        ctx->stage_next ^= 1;
        ctx->stage_ptr[slot] = (const void*)d_rir_staging;
        SSB_CUDA(ctx, cudaEventRecord(ctx->stage_done[slot], st));
    }
    // the caller's stream completes when the last D2H copy has landed
    SSB_CUDA(ctx, cudaEventRecord(ev_done, ctx->s_d2h));
    SSB_CUDA(ctx, cudaStreamWaitEvent(st, ev_done, 0));
    return SSB_OK;
}

static int ssb_sh_decode_batch_impl(ssb_ctx* ctx, int B, const float* d_amb, int L, const float* d_az_deg, const float* d_hbank,
                                   void* d_filters, float* d_out_rir, void* stream) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (B == 0) return SSB_OK;
    if (B < 0 || B > 65535 || L <= 0 || !d_amb || !d_az_deg || !d_hbank || !d_filters || !d_out_rir ||
        ((uintptr_t)d_out_rir & 7) || ((uintptr_t)d_filters & 7))
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_sh_decode_batch: bad arguments (B=%d L=%d)", B, L);
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int SPAN = SH_TILE + SH_DELAY + SH_TAPS - 1;
    const size_t smem = (size_t)(SH_CH * SPAN + 1) * sizeof(float) + (size_t)SH_CH * SH_TAPS * sizeof(float2);
    if (!ctx->attr_sh) {
        SSB_CUDA(ctx, cudaFuncSetAttribute(sh_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_sh = true;
    }
    sh_filters_kernel<<<B, SH_TAPS, 0, st>>>(d_hbank, d_az_deg, (float2*)d_filters);
    SSB_CUDA(ctx, cudaGetLastError());
    if (L >= 2 * SHF_HOP && !(ctx->debug & 32)) {
        // FFT path: filter spectra (context-owned scratch), then overlap-save blocks
        using P = FftPlan<SHF_LOG2N>;
        const size_t need = (size_t)B * SH_CH * P::N;
        if (need > ctx->gscratch_elems) {
            if (ctx->gscratch) { cudaDeviceSynchronize(); cudaFree(ctx->gscratch); ctx->gscratch = nullptr; ctx->gscratch_elems = 0; }
            SSB_CUDA(ctx, cudaMalloc(&ctx->gscratch, need * sizeof(float2)));
            ctx->gscratch_elems = need;
        }
        const size_t fsmem = (P::SMEM_ELEMS + P::TW_SMALL_ELEMS) * sizeof(float2);
        if (!ctx->attr_shf) {
            SSB_CUDA(ctx, cudaFuncSetAttribute(sh_filter_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            SSB_CUDA(ctx, cudaFuncSetAttribute(sh_ols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            ctx->attr_shf = true;
        }
        sh_filter_fft_kernel<<<dim3(SH_CH, B), P::T, fsmem, st>>>((const float2*)d_filters, ctx->gscratch, ctx->tw[SHF_LOG2N]);
        SSB_CUDA(ctx, cudaGetLastError());
        sh_ols_kernel<<<dim3((L + SHF_HOP - 1) / SHF_HOP, B), P::T, fsmem, st>>>(d_amb, L, ctx->gscratch, d_out_rir,
                                                                              ctx->tw[SHF_LOG2N]);
        ctx->launches += 3;
        SSB_CUDA(ctx, cudaGetLastError());
        return SSB_OK;
    }
    dim3 g((L + SH_TILE - 1) / SH_TILE, B);
    sh_decode_kernel<<<g, SH_THREADS, smem, st>>>(d_amb, L, (const float2*)d_filters, d_out_rir);
    ctx->launches += 2;
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}


// ---------------------------------------------------------------------------
// log-mel host side: the Slaney filterbank of librosa.filters.mel(sr, n_fft=512, n_mels, fmin=0, fmax=sr/2,
// htk=False, norm='slaney'), built in double precision, rounded like librosa (float32 triangle, then
// float32(triangle * enorm)) and stored row-sparse.
// ---------------------------------------------------------------------------
static double hz_to_mel_slaney(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz_slaney(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// row-sparse filterbank: rows[j] = (first bin, count), ofs[j] = index of the row's first weight in w
static void build_mel_rows(int sr, int n_mels, std::vector<int2>& rows, std::vector<int>& ofs, std::vector<float>& w) {
    const int n_bins = SSB_N_FFT / 2 + 1;
    std::vector<double> mel_f(n_mels + 2);
    const double m_lo = hz_to_mel_slaney(0.0), m_hi = hz_to_mel_slaney(0.5 * (double)sr);
    const double step = (m_hi - m_lo) / (double)(n_mels + 1);
    for (int i = 0; i < n_mels + 2; ++i)        // np.linspace(m_lo, m_hi, n_mels + 2)
        mel_f[i] = mel_to_hz_slaney(i == n_mels + 1 ? m_hi : (double)i * step + m_lo);
    rows.assign(n_mels, make_int2(0, 0));
    ofs.assign(n_mels, 0);
    w.clear();
    std::vector<float> rw(n_bins);
    for (int j = 0; j < n_mels; ++j) {
        const double enorm = 2.0 / (mel_f[j + 2] - mel_f[j]);
        int first = -1, last = -1;
        for (int k = 0; k < n_bins; ++k) {
            const double f = (double)k * ((double)sr / (double)SSB_N_FFT);       // librosa.fft_frequencies(sr, n_fft)
            const double lower = -(mel_f[j] - f) / (mel_f[j + 1] - mel_f[j]);
            const double upper = (mel_f[j + 2] - f) / (mel_f[j + 2] - mel_f[j + 1]);
            const float tri = (float)fmax(0.0, fmin(lower, upper));             // weights[i] is float32 ...
            rw[k] = (float)((double)tri * enorm);                               // ... `weights *= enorm` rounds once more
            if (rw[k] != 0.f) { if (first < 0) first = k; last = k; }
        }
        if (first >= 0) rows[j] = make_int2(first, last - first + 1);
        ofs[j] = (int)w.size();
        for (int k = 0; k < rows[j].y; ++k) w.push_back(rw[rows[j].x + k]);
    }
    if (w.empty()) w.push_back(0.f);
}

// host only (no device needed): dense [n_mels][257] copy of the filterbank the kernels use
extern "C" int ssb_mel_filterbank(int sr, int n_mels, float* h_out) {
    if (sr < SSB_N_FFT || n_mels < 1 || n_mels > MEL_MAX || !h_out) return SSB_E_INVALID_ARG;
    std::vector<int2> rows; std::vector<int> ofs; std::vector<float> w;
    try { build_mel_rows(sr, n_mels, rows, ofs, w); } catch (...) { return SSB_E_OOM; }
    const int n_bins = SSB_N_FFT / 2 + 1;
    for (int j = 0; j < n_mels; ++j) {
        for (int k = 0; k < n_bins; ++k) h_out[j * n_bins + k] = 0.f;
        for (int i = 0; i < rows[j].y; ++i) h_out[j * n_bins + rows[j].x + i] = w[ofs[j] + i];
    }
    return SSB_OK;
}

static int get_mel_bank(ssb_ctx* ctx, int sr, int n_mels, const ssb_ctx::MelBank** out) {
    for (auto& mb : *ctx->mel)
        if (mb.sr == sr && mb.n_mels == n_mels) { *out = &mb; return SSB_OK; }
    std::vector<int2> rows; std::vector<int> ofs; std::vector<float> w;
    build_mel_rows(sr, n_mels, rows, ofs, w);
    ssb_ctx::MelBank mb{sr, n_mels, nullptr, nullptr, nullptr};
    SSB_CUDA(ctx, cudaMalloc(&mb.rows, rows.size() * sizeof(int2)));
    SSB_CUDA(ctx, cudaMalloc(&mb.ofs, ofs.size() * sizeof(int)));
    SSB_CUDA(ctx, cudaMalloc(&mb.w, w.size() * sizeof(float)));
    SSB_CUDA(ctx, cudaMemcpy(mb.rows, rows.data(), rows.size() * sizeof(int2), cudaMemcpyHostToDevice));
    SSB_CUDA(ctx, cudaMemcpy(mb.ofs, ofs.data(), ofs.size() * sizeof(int), cudaMemcpyHostToDevice));
    SSB_CUDA(ctx, cudaMemcpy(mb.w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->mel->push_back(mb);
    *out = &ctx->mel->back();
    return SSB_OK;
}

extern "C" int ssb_logmel_frames(int sr) { return 1 + sr / SSB_HOP; }

static int ssb_logmel_batch_impl(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int n_mels, int power,
                                int pad_mode, float* d_out, void* stream) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (B == 0) return SSB_OK;
    if (B < 0 || B > 65535 || !d_wave || !d_out || sr < SSB_N_FFT || wave_stride < sr || n_mels < 1 || n_mels > MEL_MAX ||
        (power != 1 && power != 2) || (pad_mode != SSB_PAD_REFLECT && pad_mode != SSB_PAD_CONSTANT) || ((uintptr_t)d_out & 7))
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_logmel_batch: bad arguments (B=%d sr=%d n_mels=%d power=%d pad_mode=%d)", B, sr,
                 n_mels, power, pad_mode);
    const ssb_ctx::MelBank* mb = nullptr;
    int rc = get_mel_bank(ctx, sr, n_mels, &mb);
    if (rc) return rc;
    const int frames = ssb_logmel_frames(sr);
    dim3 g((frames + SSB_POOL - 1) / SSB_POOL, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (pad_mode == SSB_PAD_REFLECT)
        logmel_kernel<true><<<g, 32, 0, st>>>(d_wave, (long long)wave_stride, sr, frames, n_mels, power, d_out, ctx->tw[9],
                                               ctx->window, mb->rows, mb->ofs, mb->w);
    else
        logmel_kernel<false><<<g, 32, 0, st>>>(d_wave, (long long)wave_stride, sr, frames, n_mels, power, d_out, ctx->tw[9],
                                                ctx->window, mb->rows, mb->ofs, mb->w);
    ctx->launches += 1;
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}

static int ssb_intensity_batch_impl(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int num_frame,
                                   float* d_out, void* stream) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (B == 0) return SSB_OK;
    if (B < 0 || !d_wave || !d_out || sr <= 0 || wave_stride < sr || num_frame <= 0)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_intensity_batch: bad arguments");
    intensity_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(d_wave, (long long)wave_stride, sr, num_frame, d_out);
    ctx->launches += 1;
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}

extern "C" int ssb_host_copy_bytes(const ssb_ctx* ctx, int64_t* h2d, int64_t* d2h) {
    if (!ctx || !h2d || !d2h) return SSB_E_INVALID_ARG;
    *h2d = ctx->last_h2d_bytes;
    *d2h = ctx->last_d2h_bytes;
    return SSB_OK;
}

static int ssb_audio_conv1_batch_impl(ssb_ctx* ctx, int B, const float* d_spec, int H, int W, const float* d_weight,
                                      const float* d_bias, int OC, int KH, int KW, int SH, int SW, int relu, float* d_out,
                                      void* stream) {
    if (B == 0) return SSB_OK;
    if (B < 0 || B > 65535 || !d_spec || !d_weight || !d_out || H < 1 || W < 1 || OC < 1 || KH < 1 || KW < 1 || SH < 1 || SW < 1 ||
        KH > H || KW > W)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_audio_conv1_batch: bad arguments (B=%d H=%d W=%d OC=%d K=%dx%d S=%dx%d)", B, H, W, OC, KH,
                 KW, SH, SW);
    const int H1 = (H - KH) / SH + 1, W1 = (W - KW) / SW + 1;
    const size_t smem = ((size_t)KH * W * 2 + (size_t)OC * 2 * KH * KW) * sizeof(float);
    if (W1 * OC > 1024 || smem > 48 * 1024)
        SSB_FAIL(ctx, SSB_E_SHAPE, "ssb_audio_conv1_batch: layer too large for the fused kernel (W1*OC=%d, smem=%zu)", W1 * OC, smem);
    audio_conv1_kernel<<<dim3(H1, B), W1 * OC, smem, (cudaStream_t)stream>>>(d_spec, H, W, d_weight, d_bias, OC, KH, KW, SH, SW, H1, W1,
                                                                          relu, d_out);
    ctx->launches += 1;
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}

static int ssb_pcm16_decode_impl(ssb_ctx* ctx, const int16_t* d_in, int64_t n, float* d_out, void* stream) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (n == 0) return SSB_OK;
    if (n < 0 || !d_in || !d_out) SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_pcm16_decode: bad arguments");
    int blocks = (int)((n + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    pcm16_decode_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_in, (long long)n, d_out);
    ctx->launches += 1;
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}

static int ssb_pcm16_encode_impl(ssb_ctx* ctx, const float* d_in, int64_t n, int mode, int16_t* d_out, void* stream) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (n == 0) return SSB_OK;
    if (n < 0 || !d_in || !d_out || (mode != 0 && mode != 1)) SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_pcm16_encode: bad arguments");
    int blocks = (int)((n + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    pcm16_encode_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_in, (long long)n, mode, d_out);
    ctx->launches += 1;
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}

// ---------------------------------------------------------------------------
// C ABI entries: device guard + exception barrier around the implementations above
// ---------------------------------------------------------------------------
extern "C" int ssb_get_kernel_timing(ssb_ctx* ctx, double* ms_sum, int64_t* counts) {
    return abi_call(ctx, [&] { return ssb_get_kernel_timing_impl(ctx, ms_sum, counts); });
}

extern "C" int ssb_source_windows(ssb_ctx* ctx, const ssb_plan* plan, const float* d_src, int S, int64_t m0, int wrap,
                                  int nw, int wofs, void* d_x, void* stream) {
    return abi_call(ctx, [&] { return ssb_source_windows_impl(ctx, plan, d_src, S, m0, wrap, nw, wofs, d_x, stream); });
}

extern "C" int ssb_convolve_batch(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs,
                                  const float* d_rir_bank, const void* d_xpool, void* d_hscratch, float* d_wave,
                                  int64_t wave_stride, void* stream) {
    return abi_call(ctx, [&] { return ssb_convolve_batch_impl(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_hscratch, d_wave, wave_stride, stream); });
}

extern "C" int ssb_crossfade_batch(ssb_ctx* ctx, int B, const float* d_prev, float* d_cur, int sr, int64_t wave_stride,
                                   const uint8_t* d_enable, void* stream) {
    return abi_call(ctx, [&] { return ssb_crossfade_batch_impl(ctx, B, d_prev, d_cur, sr, wave_stride, d_enable, stream); });
}

extern "C" int ssb_spectrogram_batch(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int pad_mode,
                                     float* d_spec, void* stream) {
    return abi_call(ctx, [&] { return ssb_spectrogram_batch_impl(ctx, B, d_wave, wave_stride, sr, pad_mode, d_spec, stream); });
}

extern "C" int ssb_render_batch(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs, const float* d_rir_bank,
                                const void* d_xpool, void* d_hscratch, float* d_wave, int64_t wave_stride, int pad_mode,
                                float* d_spec, void* stream) {
    return abi_call(ctx, [&] { return ssb_render_batch_impl(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_hscratch, d_wave, wave_stride, pad_mode, d_spec, stream); });
}

extern "C" int ssb_render_batch_host(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* h_reqs, const float* h_rir,
                                     int64_t rir_bytes, float* d_rir_staging, ssb_req* d_reqs_staging, const void* d_xpool,
                                     void* d_hscratch, float* d_wave, int64_t wave_stride, int pad_mode, float* d_spec,
                                     float* h_spec, float* h_wave, int n_chunks, void* stream) {
    return abi_call(ctx, [&] { return ssb_render_batch_host_impl(ctx, plan, B, h_reqs, h_rir, rir_bytes, d_rir_staging, d_reqs_staging, d_xpool, d_hscratch, d_wave, wave_stride, pad_mode, d_spec, h_spec, h_wave, n_chunks, stream); });
}

extern "C" int ssb_sh_decode_batch(ssb_ctx* ctx, int B, const float* d_amb, int L, const float* d_az_deg, const float* d_hbank,
                                   void* d_filters, float* d_out_rir, void* stream) {
    return abi_call(ctx, [&] { return ssb_sh_decode_batch_impl(ctx, B, d_amb, L, d_az_deg, d_hbank, d_filters, d_out_rir, stream); });
}

extern "C" int ssb_logmel_batch(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int n_mels, int power,
                                int pad_mode, float* d_out, void* stream) {
    return abi_call(ctx, [&] { return ssb_logmel_batch_impl(ctx, B, d_wave, wave_stride, sr, n_mels, power, pad_mode, d_out, stream); });
}

extern "C" int ssb_intensity_batch(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int num_frame,
                                   float* d_out, void* stream) {
    return abi_call(ctx, [&] { return ssb_intensity_batch_impl(ctx, B, d_wave, wave_stride, sr, num_frame, d_out, stream); });
}

extern "C" int ssb_pcm16_decode(ssb_ctx* ctx, const int16_t* d_in, int64_t n, float* d_out, void* stream) {
    return abi_call(ctx, [&] { return ssb_pcm16_decode_impl(ctx, d_in, n, d_out, stream); });
}

extern "C" int ssb_pcm16_encode(ssb_ctx* ctx, const float* d_in, int64_t n, int mode, int16_t* d_out, void* stream) {
    return abi_call(ctx, [&] { return ssb_pcm16_encode_impl(ctx, d_in, n, mode, d_out, stream); });
}
extern "C" int ssb_audio_conv1_batch(ssb_ctx* ctx, int B, const float* d_spec, int H, int W, const float* d_weight,
                                     const float* d_bias, int OC, int KH, int KW, int SH, int SW, int relu, float* d_out,
                                     void* stream) {
    return abi_call(ctx, [&] { return ssb_audio_conv1_batch_impl(ctx, B, d_spec, H, W, d_weight, d_bias, OC, KH, KW, SH, SW, relu, d_out, stream); });
}
