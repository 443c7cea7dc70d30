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
#include <string.h>

#include <new>
#include <vector>

#include "../../include/ssb200.h"
#include "fft16.cuh"

using namespace ssb;

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
enum { K_FWD_RIR = 0, K_MAC_IFFT = 1, K_SPECTROGRAM = 2, K_FWD_SRC = 3, K_COUNT = SSB_N_KERNELS };
struct TimedLaunch { int kernel; cudaEvent_t a, b; };

struct ssb_ctx {
    int device;
    int sm_count;
    float2* tw[16];      // twiddle tables by log2n (device)
    float* window;       // 512-float centre-padded periodic Hann(400)
    int64_t launches;
    // optional per-kernel CUDA-event timing (bench.py roofline); see ssb_set_kernel_timing
    int timing;
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

static const int kSupportedLog2[] = {9, 12, 13, 14};

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
        ctx->timed->push_back(tl);
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

// grid (max_parts * n_terms, B); block T.  H[env][term][p][N] in slot order.
template <int LOG2N>
__global__ void __launch_bounds__(FftPlan<LOG2N>::T)
fwd_rir_kernel(const ssb_req* __restrict__ reqs, const float2* __restrict__ rir_bank,
               float2* __restrict__ H, int max_parts, long long h_elems_per_env,
               const float2* __restrict__ tw) {
    using P = FftPlan<LOG2N>;
    extern __shared__ float2 smem[];
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
        v[q] = n < taps ? __ldg(src + n) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int q = 8; q < 16; ++q) v[q] = make_float2(0.f, 0.f);
    fft_forward<LOG2N>(v, t, smem, tw);
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
    constexpr int PART = P::N / 2;
    const int j = blockIdx.x;
    const int t = threadIdx.x;
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
    fft_forward<LOG2N>(v, t, smem, tw);
    store_slots<LOG2N>(X + (long long)j * P::N, v, t);
}

// ---------------------------------------------------------------------------
// multiply-accumulate over partitions + inverse FFT + emit
// grid (n_blocks, B); block T.
// ---------------------------------------------------------------------------
template <int LOG2N>
__global__ void __launch_bounds__(FftPlan<LOG2N>::T)
mac_ifft_kernel(const ssb_req* __restrict__ reqs, const float2* __restrict__ xpool,
                const float2* __restrict__ H, int max_parts, int n_terms, long long h_elems_per_env,
                float* __restrict__ wave, long long wave_stride, int sr,
                const float2* __restrict__ tw) {
    using P = FftPlan<LOG2N>;
    extern __shared__ float2 smem[];
    constexpr int PART = P::N / 2;
    const int b = blockIdx.x;
    const int env = blockIdx.y;
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
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
    for (int term = 0; term < n_terms; ++term) {
        const ssb_conv_term& ct = rq.term[term];
        if (ct.rir_taps <= 0) continue;
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
            for (int i = 0; i < 16; ++i) {
                acc[i].x = fmaf(x[i].x, h[i].x, fmaf(-x[i].y, h[i].y, acc[i].x));
                acc[i].y = fmaf(x[i].x, h[i].y, fmaf(x[i].y, h[i].x, acc[i].y));
            }
        }
    }
    fft_inverse<LOG2N>(acc, t, smem, tw);
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
// spectrogram: one warp per STFT frame, 8 frames (2 pooled columns) per CTA
// grid (ceil(cols/2), B); block 256
// ---------------------------------------------------------------------------
constexpr int SPEC_WARPS = 8;
constexpr int SPEC_COLS_PER_CTA = SPEC_WARPS / SSB_POOL;
constexpr int SPEC_BUF = 512 + 32;

__device__ __forceinline__ int nat_idx(int k) { return k + ((k >> 8) << 3); }   // de-conflict k and 256+k

__global__ void __launch_bounds__(SPEC_WARPS * 32)
spectrogram_kernel(const float* __restrict__ wave, long long wave_stride, int sr, int n_frames, int cols,
                   int pad_mode, float* __restrict__ out, const float2* __restrict__ tw,
                   const float* __restrict__ window) {
    __shared__ float2 xbuf[SPEC_WARPS][SPEC_BUF];
    __shared__ float fsum[SPEC_WARPS][SSB_SPEC_ROWS][2];
    __shared__ float2 stw[FftPlan<9>::TW_ELEMS];
    for (int i = threadIdx.x; i < FftPlan<9>::TW_ELEMS; i += SPEC_WARPS * 32) stw[i] = __ldg(tw + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.y;
    const int col0 = blockIdx.x * SPEC_COLS_PER_CTA;
    const int f = col0 * SSB_POOL + warp;
    const float* __restrict__ yl = wave + (long long)env * 2 * wave_stride;
    const float* __restrict__ yr = yl + wave_stride;
    float2* buf = xbuf[warp];
    if (f < n_frames) {
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int idx = lane + 32 * q;
            float2 z = make_float2(0.f, 0.f);
            if (idx >= (SSB_N_FFT - SSB_WIN) / 2 && idx < (SSB_N_FFT + SSB_WIN) / 2) {
                int n = f * SSB_HOP + idx - SSB_N_FFT / 2;
                bool ok = true;
                if (pad_mode == SSB_PAD_REFLECT) {
                    if (n < 0) n = -n;
                    if (n >= sr) n = 2 * (sr - 1) - n;
                } else {
                    ok = n >= 0 && n < sr;
                }
                if (ok) {
                    float w = __ldg(window + idx);
                    z = make_float2(w * __ldg(yl + n), w * __ldg(yr + n));
                }
            }
            v[q] = z;
        }
        fft_forward<9, false>(v, lane, buf, stw);
        __syncwarp();
        // natural order: k = (lane>>1) + 16 i + 256 (lane&1)
#pragma unroll
        for (int i = 0; i < 16; ++i) buf[nat_idx((lane >> 1) + 16 * i + 256 * (lane & 1))] = v[i];
        __syncwarp();
        // Z = FFT(w*(yL + i yR)):  XL[k] = (Z[k] + conj Z[N-k])/2,  XR[k] = (Z[k] - conj Z[N-k])/(2i)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = lane + 32 * m;
            float2 a = buf[nat_idx(k)];
            float2 bb = buf[nat_idx((SSB_N_FFT - k) & (SSB_N_FFT - 1))];
            float lx = a.x + bb.x, ly = a.y - bb.y;      // A + conj(B)
            float rx = a.x - bb.x, ry = a.y + bb.y;      // A - conj(B)
            float ml = 0.5f * sqrtf(lx * lx + ly * ly);
            float mr = 0.5f * sqrtf(rx * rx + ry * ry);
            ml += __shfl_xor_sync(0xffffffffu, ml, 1);
            mr += __shfl_xor_sync(0xffffffffu, mr, 1);
            ml += __shfl_xor_sync(0xffffffffu, ml, 2);
            mr += __shfl_xor_sync(0xffffffffu, mr, 2);
            if ((lane & 3) == 0) {
                fsum[warp][(lane >> 2) + 8 * m][0] = ml;
                fsum[warp][(lane >> 2) + 8 * m][1] = mr;
            }
        }
        if (lane == 0) {                                   // bin 256 is alone in pooled row 64
            float2 a = buf[nat_idx(256)];
            fsum[warp][64][0] = fabsf(a.x);
            fsum[warp][64][1] = fabsf(a.y);
        }
    } else {
        for (int i = lane; i < SSB_SPEC_ROWS * 2; i += 32) (&fsum[warp][0][0])[i] = 0.f;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < SPEC_COLS_PER_CTA * SSB_SPEC_ROWS * 2; idx += SPEC_WARPS * 32) {
        const int c = idx / (SSB_SPEC_ROWS * 2);
        const int r = (idx % (SSB_SPEC_ROWS * 2)) >> 1;
        const int e = idx & 1;
        const int col = col0 + c;
        if (col < cols) {
            float s = (fsum[4 * c][r][e] + fsum[4 * c + 1][r][e]) + (fsum[4 * c + 2][r][e] + fsum[4 * c + 3][r][e]);
            out[(((long long)env * SSB_SPEC_ROWS + r) * cols + col) * 2 + e] = log1pf(s * (1.0f / 16.0f));
        }
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
    const int bytes = P::SMEM_ELEMS * (int)sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(fwd_rir_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fwd_src_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mac_ifft_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e == cudaSuccess ? 0 : -1;
}

// twiddle table of FftPlan<LOG2N>: pass p, [s-1][j] = exp(-2 pi i j s / (16 stride(p))), rounded from double
template <int LOG2N>
static cudaError_t upload_twiddles(ssb_ctx* ctx) {
    using P = FftPlan<LOG2N>;
    std::vector<float2> h(P::TW_ELEMS);
    for (int p = 0; p < P::NPASS; ++p) {
        const int st = P::stride(p);
        for (int s = 1; s < 16; ++s)
            for (int j = 0; j < st; ++j) {
                const double a = -2.0 * M_PI * (double)j * (double)s / (16.0 * (double)st);
                h[P::tw_offset(p) + (s - 1) * st + j] = make_float2((float)cos(a), (float)sin(a));
            }
    }
    cudaError_t e = cudaMalloc(&ctx->tw[LOG2N], h.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->tw[LOG2N], h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
    return e;
}

extern "C" int ssb_version(void) { return 100; }

extern "C" int ssb_create(int device, ssb_ctx** out) {
    if (!out) return SSB_E_INVALID_ARG;
    *out = nullptr;
    ssb_ctx* ctx = new (std::nothrow) ssb_ctx();
    if (!ctx) return SSB_E_OOM;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->timed = new (std::nothrow) std::vector<TimedLaunch>();
    *out = ctx;   // returned even on failure so that ssb_last_error works; caller destroys
    SSB_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    SSB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) SSB_FAIL(ctx, SSB_E_CUDA, "device %d is sm_%d%d; libssb200 is built for sm_100a only", device, prop.major, prop.minor);
    ctx->sm_count = prop.multiProcessorCount;
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
    if (setup_smem_attrs<12>() || setup_smem_attrs<13>() || setup_smem_attrs<14>())
        SSB_FAIL(ctx, SSB_E_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: %s",
                 cudaGetErrorString(cudaGetLastError()));
    return SSB_OK;
}

extern "C" void ssb_destroy(ssb_ctx* ctx) {
    if (!ctx) return;
    for (int l = 0; l < 16; ++l)
        if (ctx->tw[l]) cudaFree(ctx->tw[l]);
    if (ctx->window) cudaFree(ctx->window);
    if (ctx->timed) {
        for (auto& tl : *ctx->timed) { cudaEventDestroy(tl.a); cudaEventDestroy(tl.b); }
        delete ctx->timed;
    }
    delete ctx;
}

extern "C" const char* ssb_last_error(const ssb_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" int64_t ssb_launch_count(const ssb_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ssb_set_kernel_timing(ssb_ctx* ctx, int enable) {
    if (!ctx) return SSB_E_INVALID_ARG;
    ctx->timing = enable ? 1 : 0;
    return SSB_OK;
}

extern "C" int ssb_get_kernel_timing(ssb_ctx* ctx, double* ms_sum, int64_t* counts) {
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
    if (log2n == 0) log2n = 13;
    if (log2n < 12 || !log2_supported(log2n)) SSB_FAIL(ctx, SSB_E_INVALID_ARG, "log2n %d unsupported (12, 13, 14)", log2n);
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
        fwd_src_kernel<LOG2N><<<nw, P::T, P::SMEM_ELEMS * sizeof(float2), st>>>(d_src, S, (long long)m0, wrap, wofs, d_x, ctx->tw[LOG2N]);
    }
    return cudaGetLastError();
}

extern "C" int ssb_source_windows(ssb_ctx* ctx, const ssb_plan* plan, const float* d_src, int S, int64_t m0, int wrap,
                                  int nw, int wofs, void* d_x, void* stream) {
    int rc = check_plan(ctx, plan);
    if (rc) return rc;
    if (!d_src || !d_x || S <= 0 || nw <= 0 || wofs < 0 || wofs >= nw)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_source_windows: bad arguments (S=%d nw=%d wofs=%d)", S, nw, wofs);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    switch (plan->log2n) {
        case 12: e = launch_src<12>(ctx, d_src, S, m0, wrap, nw, wofs, (float2*)d_x, st); break;
        case 13: e = launch_src<13>(ctx, d_src, S, m0, wrap, nw, wofs, (float2*)d_x, st); break;
        default: e = launch_src<14>(ctx, d_src, S, m0, wrap, nw, wofs, (float2*)d_x, st); break;
    }
    SSB_CUDA(ctx, e);
    return SSB_OK;
}

template <int LOG2N>
static cudaError_t launch_conv(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs, const float* d_rir_bank,
                               const void* d_xpool, void* d_h, float* d_wave, int64_t wave_stride, cudaStream_t st) {
    using P = FftPlan<LOG2N>;
    const size_t smem = P::SMEM_ELEMS * sizeof(float2);
    dim3 g1(plan->max_parts * plan->n_terms, B);
    {
        LaunchTimer lt(ctx, K_FWD_RIR, st);
        fwd_rir_kernel<LOG2N><<<g1, P::T, smem, st>>>(d_reqs, (const float2*)d_rir_bank, (float2*)d_h, plan->max_parts,
                                                      (long long)plan->h_elems_per_env, ctx->tw[LOG2N]);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 g2(plan->n_blocks, B);
    {
        LaunchTimer lt(ctx, K_MAC_IFFT, st);
        mac_ifft_kernel<LOG2N><<<g2, P::T, smem, st>>>(d_reqs, (const float2*)d_xpool, (const float2*)d_h, plan->max_parts,
                                                       plan->n_terms, (long long)plan->h_elems_per_env, d_wave,
                                                       (long long)wave_stride, plan->sr, ctx->tw[LOG2N]);
    }
    return cudaGetLastError();
}

extern "C" int ssb_convolve_batch(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs,
                                  const float* d_rir_bank, const void* d_xpool, void* d_hscratch, float* d_wave,
                                  int64_t wave_stride, void* stream) {
    int rc = check_plan(ctx, plan);
    if (rc) return rc;
    if (B == 0) return SSB_OK;
    if (B < 0 || B > 65535 || !d_reqs || !d_rir_bank || !d_xpool || !d_hscratch || !d_wave || wave_stride < plan->sr)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_convolve_batch: bad arguments (B=%d wave_stride=%lld sr=%d)", B,
                 (long long)wave_stride, plan->sr);
    if (((uintptr_t)d_rir_bank & 7) || ((uintptr_t)d_xpool & 7) || ((uintptr_t)d_hscratch & 7))
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_convolve_batch: rir bank / spectra must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    switch (plan->log2n) {
        case 12: e = launch_conv<12>(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_hscratch, d_wave, wave_stride, st); break;
        case 13: e = launch_conv<13>(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_hscratch, d_wave, wave_stride, st); break;
        default: e = launch_conv<14>(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_hscratch, d_wave, wave_stride, st); break;
    }
    SSB_CUDA(ctx, e);
    return SSB_OK;
}

extern "C" int ssb_crossfade_batch(ssb_ctx* ctx, int B, const float* d_prev, float* d_cur, int sr, int64_t wave_stride,
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

extern "C" int ssb_spectrogram_batch(ssb_ctx* ctx, int B, const float* d_wave, int64_t wave_stride, int sr, int pad_mode,
                                     float* d_spec, void* stream) {
    if (!ctx) return SSB_E_INVALID_ARG;
    if (B == 0) return SSB_OK;
    if (B < 0 || B > 65535 || !d_wave || !d_spec || sr < SSB_N_FFT || wave_stride < sr ||
        (pad_mode != SSB_PAD_REFLECT && pad_mode != SSB_PAD_CONSTANT))
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_spectrogram_batch: bad arguments (B=%d sr=%d pad_mode=%d)", B, sr, pad_mode);
    const int frames = 1 + sr / SSB_HOP;
    const int cols = ssb_spec_cols(sr);
    dim3 g((cols + SPEC_COLS_PER_CTA - 1) / SPEC_COLS_PER_CTA, B);
    {
        LaunchTimer lt(ctx, K_SPECTROGRAM, (cudaStream_t)stream);
        spectrogram_kernel<<<g, SPEC_WARPS * 32, 0, (cudaStream_t)stream>>>(d_wave, (long long)wave_stride, sr, frames, cols,
                                                                           pad_mode, d_spec, ctx->tw[9], ctx->window);
    }
    SSB_CUDA(ctx, cudaGetLastError());
    return SSB_OK;
}

extern "C" int ssb_render_batch(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* d_reqs, const float* d_rir_bank,
                                const void* d_xpool, void* d_hscratch, float* d_wave, int64_t wave_stride, int pad_mode,
                                float* d_spec, void* stream) {
    int rc = ssb_convolve_batch(ctx, plan, B, d_reqs, d_rir_bank, d_xpool, d_hscratch, d_wave, wave_stride, stream);
    if (rc) return rc;
    return ssb_spectrogram_batch(ctx, B, d_wave, wave_stride, plan->sr, pad_mode, d_spec, stream);
}

extern "C" int ssb_render_batch_host(ssb_ctx* ctx, const ssb_plan* plan, int B, const ssb_req* h_reqs, const float* h_rir,
                                     int64_t rir_bytes, float* d_rir_staging, ssb_req* d_reqs_staging, const void* d_xpool,
                                     void* d_hscratch, float* d_wave, int64_t wave_stride, int pad_mode, float* d_spec,
                                     float* h_spec, float* h_wave, void* stream) {
    int rc = check_plan(ctx, plan);
    if (rc) return rc;
    if (B <= 0 || !h_reqs || !h_rir || rir_bytes <= 0 || !d_rir_staging || !d_reqs_staging || !h_spec)
        SSB_FAIL(ctx, SSB_E_INVALID_ARG, "ssb_render_batch_host: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SSB_CUDA(ctx, cudaMemcpyAsync(d_rir_staging, h_rir, (size_t)rir_bytes, cudaMemcpyHostToDevice, st));
    SSB_CUDA(ctx, cudaMemcpyAsync(d_reqs_staging, h_reqs, (size_t)B * sizeof(ssb_req), cudaMemcpyHostToDevice, st));
    rc = ssb_render_batch(ctx, plan, B, d_reqs_staging, d_rir_staging, d_xpool, d_hscratch, d_wave, wave_stride, pad_mode,
                          d_spec, stream);
    if (rc) return rc;
    const size_t spec_bytes = (size_t)B * SSB_SPEC_ROWS * ssb_spec_cols(plan->sr) * 2 * sizeof(float);
    SSB_CUDA(ctx, cudaMemcpyAsync(h_spec, d_spec, spec_bytes, cudaMemcpyDeviceToHost, st));
    if (h_wave)
        SSB_CUDA(ctx, cudaMemcpy2DAsync(h_wave, (size_t)plan->sr * sizeof(float), d_wave, (size_t)wave_stride * sizeof(float),
                                        (size_t)plan->sr * sizeof(float), (size_t)B * 2, cudaMemcpyDeviceToHost, st));
    return SSB_OK;
}

extern "C" int ssb_pcm16_decode(ssb_ctx* ctx, const int16_t* d_in, int64_t n, float* d_out, void* stream) {
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

extern "C" int ssb_pcm16_encode(ssb_ctx* ctx, const float* d_in, int64_t n, int mode, int16_t* d_out, void* stream) {
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
