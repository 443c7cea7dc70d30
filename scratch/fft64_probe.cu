// scratch/fft64_probe.cu -- ROUND-2 PROTOTYPE, not product code, not measured yet.
//
// A 4096-point complex FFT as 64 x 64 with 64 points per thread (64 threads = 2 warps per transform): ONE
// shared-memory exchange (a 64 x 64 transpose) and one 2-warp barrier per transform, against two exchanges and a
// block barrier for the radix-16 schedule of csrc/fft16.cuh (256 threads, 16 points per thread).  Same flop count
// (1186 packed FP32 instructions per thread x 64 threads vs ~330 x 256), half the shared-memory traffic, 4x the ILP
// per thread; ptxas: 168 registers, no spills -> 6 transforms resident per SM (5 for the radix-16 inverse kernel).
// The spectrum lands in NATURAL order (dst[r * 64 + s] = X[s + 64 r]), so the per-bin partition sums are unaffected.
//
// This file is a micro-benchmark: forward/inverse kernels for both schedules behind a tiny C ABI, driven by
// scratch/fft64_probe.py (numerics vs torch.fft, then transforms/s of each).  If the 64 x 64 schedule wins here it
// replaces FftPlan<12> in fwd_rir / fwd_src / mac_ifft ("Next").
//
// build: nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//             -o scratch/libfft64_probe.so scratch/fft64_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <vector>

#include "../soundspaces_b200/csrc/fft16.cuh"
using namespace ssb;

// ------------------------------------------------------------------------------------------------ radix-64 codelet
// 8-point DFT in place, natural order out (forward w8 = exp(-i pi/4); INV conjugates)
template <bool INV>
__device__ __forceinline__ void fft8(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4, float2& a5, float2& a6,
                                     float2& a7) {
    constexpr float R2 = 0.70710678118654752f;
    float2 b0 = cadd(a0, a4), b4 = csub(a0, a4);
    float2 b1 = cadd(a1, a5), b5 = csub(a1, a5);
    float2 b2 = cadd(a2, a6), b6 = csub(a2, a6);
    float2 b3 = cadd(a3, a7), b7 = csub(a3, a7);
    b5 = mul2(cadd(b5, INV ? mul_pi(b5) : mul_mi(b5)), bcast(R2));      // * w8^1
    b6 = INV ? mul_pi(b6) : mul_mi(b6);                                  // * w8^2
    b7 = mul2(csub(INV ? mul_pi(b7) : mul_mi(b7), b7), bcast(R2));      // * w8^3
    bfly4<INV>(b0, b1, b2, b3);                                          // X0, X2, X4, X6
    bfly4<INV>(b4, b5, b6, b7);                                          // X1, X3, X5, X7
    a0 = b0; a1 = b4; a2 = b1; a3 = b5; a4 = b2; a5 = b6; a6 = b3; a7 = b7;
}

// cos / sin of 2 pi k / 64, correctly rounded
__device__ constexpr float W64C[64] = {1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f, 0.92387953251128674f, 0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f, 0.70710678118654757f, 0.63439328416364549f, 0.55557023301960229f, 0.47139673682599781f, 0.38268343236508984f, 0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f, 6.123233995736766e-17f, -0.098017140329560645f, -0.19509032201612819f, -0.29028467725446216f, -0.38268343236508973f, -0.4713967368259977f, -0.55557023301960196f, -0.63439328416364538f, -0.70710678118654746f, -0.77301045336273699f, -0.83146961230254535f, -0.88192126434835494f, -0.92387953251128674f, -0.95694033573220882f, -0.98078528040323043f, -0.99518472667219682f, -1.0f, -0.99518472667219693f, -0.98078528040323043f, -0.95694033573220894f, -0.92387953251128685f, -0.88192126434835505f, -0.83146961230254546f, -0.7730104533627371f, -0.70710678118654768f, -0.63439328416364593f, -0.55557023301960218f, -0.47139673682599786f, -0.38268343236509034f, -0.29028467725446244f, -0.19509032201612866f, -0.098017140329560451f, -1.8369701987210297e-16f, 0.09801714032956009f, 0.1950903220161283f, 0.29028467725446205f, 0.38268343236509f, 0.47139673682599759f, 0.55557023301960184f, 0.6343932841636456f, 0.70710678118654735f, 0.77301045336273666f, 0.83146961230254524f, 0.88192126434835483f, 0.92387953251128652f, 0.95694033573220882f, 0.98078528040323032f, 0.99518472667219693f};
__device__ constexpr float W64S[64] = {0.0f, 0.098017140329560604f, 0.19509032201612825f, 0.29028467725446233f, 0.38268343236508978f, 0.47139673682599764f, 0.55557023301960218f, 0.63439328416364549f, 0.70710678118654746f, 0.77301045336273699f, 0.83146961230254524f, 0.88192126434835494f, 0.92387953251128674f, 0.95694033573220894f, 0.98078528040323043f, 0.99518472667219682f, 1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220894f, 0.92387953251128674f, 0.88192126434835505f, 0.83146961230254546f, 0.7730104533627371f, 0.70710678118654757f, 0.63439328416364549f, 0.55557023301960218f, 0.47139673682599786f, 0.38268343236508989f, 0.29028467725446239f, 0.19509032201612861f, 0.098017140329560826f, 1.2246467991473532e-16f, -0.09801714032956059f, -0.19509032201612836f, -0.29028467725446211f, -0.38268343236508967f, -0.47139673682599764f, -0.55557023301960196f, -0.63439328416364527f, -0.70710678118654746f, -0.77301045336273666f, -0.83146961230254524f, -0.88192126434835494f, -0.92387953251128652f, -0.95694033573220882f, -0.98078528040323032f, -0.99518472667219693f, -1.0f, -0.99518472667219693f, -0.98078528040323043f, -0.95694033573220894f, -0.92387953251128663f, -0.88192126434835505f, -0.83146961230254546f, -0.77301045336273688f, -0.70710678118654768f, -0.63439328416364593f, -0.55557023301960218f, -0.47139673682599792f, -0.38268343236509039f, -0.2902846772544625f, -0.19509032201612872f, -0.098017140329560506f};

// 64-point DFT in registers: in v[q], q = q1 + 8 q2; out X[s1 + 8 s2] left in v[8 s1 + s2]
template <bool INV>
__device__ __forceinline__ void fft64(float2 (&v)[64]) {
#pragma unroll
    for (int q1 = 0; q1 < 8; ++q1)      // A[q1][s1] = sum_q2 v[q1 + 8 q2] w8^(q2 s1), left in v[q1 + 8 s1]
        fft8<INV>(v[q1], v[q1 + 8], v[q1 + 16], v[q1 + 24], v[q1 + 32], v[q1 + 40], v[q1 + 48], v[q1 + 56]);
#pragma unroll
    for (int q1 = 1; q1 < 8; ++q1)
#pragma unroll
        for (int s1 = 1; s1 < 8; ++s1) {   // * w64^(q1 s1)
            const float2 w = make_float2(W64C[(q1 * s1) & 63], INV ? W64S[(q1 * s1) & 63] : -W64S[(q1 * s1) & 63]);
            v[q1 + 8 * s1] = cmul(v[q1 + 8 * s1], w);
        }
#pragma unroll
    for (int s1 = 0; s1 < 8; ++s1)      // X[s1 + 8 s2] = sum_q1 A[q1][s1] w8^(q1 s2)
        fft8<INV>(v[8 * s1], v[8 * s1 + 1], v[8 * s1 + 2], v[8 * s1 + 3], v[8 * s1 + 4], v[8 * s1 + 5], v[8 * s1 + 6],
                  v[8 * s1 + 7]);
}

constexpr int R64_N = 4096, R64_T = 64, R64_LD = 65;      // transpose buffer: row stride 65 float2 (conflict free both ways)

// forward: x[t + 64 q] -> X[s + 64 r] at dst[r * 64 + s] (natural order).  tw[t * 64 + s] = exp(-2 pi i t s / 4096)
__global__ void __launch_bounds__(R64_T, 6)
fwd_r64(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw) {
    __shared__ float2 buf[R64_T * R64_LD];
    const int t = threadIdx.x;
    const float2* __restrict__ src = in + (long long)blockIdx.x * R64_N;
    float2 v[64];
#pragma unroll
    for (int q = 0; q < 64; ++q) v[q] = __ldg(src + t + 64 * q);
    fft64<false>(v);
#pragma unroll
    for (int s1 = 0; s1 < 8; ++s1)
#pragma unroll
        for (int s2 = 0; s2 < 8; ++s2) {
            const int s = s1 + 8 * s2;
            float2 y = v[8 * s1 + s2];
            if (s) y = cmul(y, __ldg(tw + t * 64 + s));
            buf[s * R64_LD + t] = y;
        }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = buf[t * R64_LD + i];            // thread s = t holds y[.][s]
    fft64<false>(v);
    float2* __restrict__ dst = out + (long long)blockIdx.x * R64_N;
#pragma unroll
    for (int r1 = 0; r1 < 8; ++r1)
#pragma unroll
        for (int r2 = 0; r2 < 8; ++r2) dst[(r1 + 8 * r2) * 64 + t] = v[8 * r1 + r2];
}

// inverse (unscaled): X natural order -> N x[n]
__global__ void __launch_bounds__(R64_T, 6)
inv_r64(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw) {
    __shared__ float2 buf[R64_T * R64_LD];
    const int s = threadIdx.x;
    const float2* __restrict__ src = in + (long long)blockIdx.x * R64_N;
    float2 v[64];
#pragma unroll
    for (int r = 0; r < 64; ++r) v[r] = __ldg(src + r * 64 + s);         // X[s + 64 r]
    fft64<true>(v);                                                      // z[s][t] in v[8 t1 + t2], t = t1 + 8 t2
#pragma unroll
    for (int t1 = 0; t1 < 8; ++t1)
#pragma unroll
        for (int t2 = 0; t2 < 8; ++t2) {
            const int t = t1 + 8 * t2;
            float2 y = v[8 * t1 + t2];
            if (t) y = cmulc(y, __ldg(tw + t * 64 + s));                  // * conj(w^(t s))
            buf[t * R64_LD + s] = y;
        }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = buf[s * R64_LD + i];            // thread t = s holds z[.][t]
    fft64<true>(v);                                                      // N x[t + 64 q] in v[8 q1 + q2], q = q1 + 8 q2
    float2* __restrict__ dst = out + (long long)blockIdx.x * R64_N;
#pragma unroll
    for (int q1 = 0; q1 < 8; ++q1)
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2) dst[s + 64 * (q1 + 8 * q2)] = v[8 * q1 + q2];
}

// ------------------------------------------------------------------------------------------------ radix-16 twins
template <int LOG2N>
__global__ void __launch_bounds__(FftPlan<LOG2N>::T)
fwd_r16(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw) {
    using P = FftPlan<LOG2N>;
    extern __shared__ float2 smem[];
    float2* stw = smem + P::SMEM_ELEMS;
    const int t = threadIdx.x;
    const float2* __restrict__ src = in + (long long)blockIdx.x * P::N;
    float2 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = __ldg(src + t + q * P::T);
    const Tw6 w0 = load_tw6<true>(tw, P::T, t);
    for (int i = t; i < P::TW_SMALL_ELEMS; i += P::T) stw[i] = __ldg(tw + P::TW_SMALL_OFFSET + i);
    fft_forward<LOG2N>(v, t, smem, w0, stw);
    float2* __restrict__ dst = out + (long long)blockIdx.x * P::N;
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i * P::T + t] = v[i];              // slot order
}

template <int LOG2N>
__global__ void __launch_bounds__(FftPlan<LOG2N>::T, 5)
inv_r16(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw) {
    using P = FftPlan<LOG2N>;
    extern __shared__ float2 smem[];
    float2* stw = smem + P::SMEM_ELEMS;
    const int t = threadIdx.x;
    const float2* __restrict__ src = in + (long long)blockIdx.x * P::N;
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __ldg(src + i * P::T + t);
    for (int i = t; i < P::TW_SMALL_ELEMS; i += P::T) stw[i] = __ldg(tw + P::TW_SMALL_OFFSET + i);
    __syncthreads();
    fft_inverse<LOG2N>(v, t, smem, tw, stw);
    float2* __restrict__ dst = out + (long long)blockIdx.x * P::N;
#pragma unroll
    for (int q = 0; q < 16; ++q) dst[t + q * P::T] = v[q];
}

// ------------------------------------------------------------------------------------------------ C entry points
static float2* g_tw64 = nullptr;
static float2* g_tw16 = nullptr;

extern "C" int probe_init(void) {
    using P = FftPlan<12>;
    std::vector<float2> h64(64 * 64), h16(P::TW_ELEMS);
    for (int t = 0; t < 64; ++t)
        for (int s = 0; s < 64; ++s) {
            const double a = -2.0 * M_PI * (double)(t * s) / 4096.0;
            h64[t * 64 + s] = make_float2((float)cos(a), (float)sin(a));
        }
    static const int kExp[6] = {1, 2, 3, 4, 8, 12};
    for (int p = 0; p < P::NPASS; ++p) {
        const int st = P::stride(p);
        for (int row = 0; row < 6; ++row)
            for (int j = 0; j < st; ++j) {
                const double a = -2.0 * M_PI * (double)j * (double)kExp[row] / (16.0 * (double)st);
                h16[P::tw_offset(p) + row * st + j] = make_float2((float)cos(a), (float)sin(a));
            }
    }
    if (cudaMalloc(&g_tw64, h64.size() * sizeof(float2)) != cudaSuccess) return -1;
    if (cudaMalloc(&g_tw16, h16.size() * sizeof(float2)) != cudaSuccess) return -1;
    cudaMemcpy(g_tw64, h64.data(), h64.size() * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(g_tw16, h16.data(), h16.size() * sizeof(float2), cudaMemcpyHostToDevice);
    const int bytes = (P::SMEM_ELEMS + P::TW_SMALL_ELEMS) * (int)sizeof(float2);
    cudaFuncSetAttribute(fwd_r16<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(inv_r16<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// which: 0 fwd_r64, 1 inv_r64, 2 fwd_r16 (slot order out), 3 inv_r16 (slot order in); n transforms of 4096 points
extern "C" int probe_run(int which, const void* d_in, void* d_out, int n, void* stream) {
    using P = FftPlan<12>;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (P::SMEM_ELEMS + P::TW_SMALL_ELEMS) * sizeof(float2);
    switch (which) {
        case 0: fwd_r64<<<n, R64_T, 0, st>>>((const float2*)d_in, (float2*)d_out, g_tw64); break;
        case 1: inv_r64<<<n, R64_T, 0, st>>>((const float2*)d_in, (float2*)d_out, g_tw64); break;
        case 2: fwd_r16<12><<<n, P::T, smem, st>>>((const float2*)d_in, (float2*)d_out, g_tw16); break;
        case 3: inv_r16<12><<<n, P::T, smem, st>>>((const float2*)d_in, (float2*)d_out, g_tw16); break;
        default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// frequency of slot (t, i) of the radix-16 schedule, for the harness: out[i * 256 + t] = X[freq]
extern "C" void probe_r16_slot_freq(int* h_out) {
    for (int t = 0; t < 256; ++t)
        for (int i = 0; i < 16; ++i) h_out[i * 256 + t] = freq_of_slot<12>(t, i);
}
