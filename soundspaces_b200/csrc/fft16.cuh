// Register-resident radix-16 FFT for sm_100a.
//
// One "group" of T = N/16 threads transforms N = 2^LOG2N complex points; every
// thread keeps 16 points in registers for the whole transform.  A transform is
//   NPASS = LOG2N/4 radix-16 passes (4x4 butterflies in registers), with one
//   shared-memory exchange between consecutive passes, followed by
//   a radix-M stage (M = N / 16^NPASS in {1,2,4}) done with warp shuffles
//   between M adjacent lanes.
// The forward transform is decimation-in-frequency and leaves the spectrum in
// a fixed digit-permuted order ("slot order": slot = i*T + t for register i of
// thread t); the inverse transform is the exact mirror image (decimation in
// time) and consumes slot order, so the pointwise spectrum product of the
// overlap-save convolution never needs a reordering pass.  freq_of_slot()
// gives the permutation for callers that need natural order (the STFT).
//
// No cuFFT, no library code: this is the hot loop of the SoundSpaces audio
// path (scipy.signal.fftconvolve at soundspaces/simulator.py:630-647 and
// librosa.stft at soundspaces/tasks/nav.py:92 in the reference).
#pragma once
#include <cuda_runtime.h>

namespace ssb {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// a * (-i) and a * (+i)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }
__device__ __forceinline__ float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }

// 4-point DFT, in place.  Forward: w4 = -i.  Inverse: w4 = +i.
template <bool INV>
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    float2 r = INV ? mul_pi(t3) : mul_mi(t3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, r);
    a3 = csub(t1, r);
}

// multiply by w16^k (forward) or conj(w16^k) (inverse), k a compile-time constant
template <bool INV>
__device__ __forceinline__ float2 mul_cs(float2 a, float c, float s) {
    // forward multiplies by (c - i s), inverse by (c + i s)
    const float ss = INV ? -s : s;
    return make_float2(a.x * c + a.y * ss, a.y * c - a.x * ss);
}
template <bool INV, int K>
__device__ __forceinline__ float2 mul_w16(float2 a) {
    constexpr float C1 = 0.92387953251128674f;   // cos(pi/8)
    constexpr float S1 = 0.38268343236508977f;   // sin(pi/8)
    constexpr float R2 = 0.70710678118654752f;
    static_assert(K == 0 || K == 1 || K == 2 || K == 3 || K == 4 || K == 6 || K == 9, "unsupported w16 power");
    if constexpr (K == 0) return a;
    else if constexpr (K == 1) return mul_cs<INV>(a, C1, S1);
    else if constexpr (K == 2)                   // (x+iy)(1 -+ i) R2
        return INV ? make_float2((a.x - a.y) * R2, (a.x + a.y) * R2)
                   : make_float2((a.x + a.y) * R2, (a.y - a.x) * R2);
    else if constexpr (K == 3) return mul_cs<INV>(a, S1, C1);
    else if constexpr (K == 4) return INV ? mul_pi(a) : mul_mi(a);
    else if constexpr (K == 6)                   // w16^6 = -R2 - i R2 (forward), conj = -R2 + i R2
        return INV ? make_float2((-a.x - a.y) * R2, (a.x - a.y) * R2)
                   : make_float2((a.y - a.x) * R2, (-a.x - a.y) * R2);
    else return mul_cs<INV>(a, -C1, -S1);        // K == 9
}

// 16-point DFT in registers: v[s] <- sum_q v[q] * w16^(+-q s)
template <bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
    // stage 1: 4-point DFTs over q = c + 4d (d = 0..3); result a[c][r] left in v[c + 4r]
#pragma unroll
    for (int c = 0; c < 4; ++c) bfly4<INV>(v[c], v[c + 4], v[c + 8], v[c + 12]);
    // twiddles w16^(c r)
    v[5] = mul_w16<INV, 1>(v[5]);
    v[9] = mul_w16<INV, 2>(v[9]);
    v[13] = mul_w16<INV, 3>(v[13]);
    v[6] = mul_w16<INV, 2>(v[6]);
    v[10] = mul_w16<INV, 4>(v[10]);
    v[14] = mul_w16<INV, 6>(v[14]);
    v[7] = mul_w16<INV, 3>(v[7]);
    v[11] = mul_w16<INV, 6>(v[11]);
    v[15] = mul_w16<INV, 9>(v[15]);
    // stage 2: 4-point DFTs over c for each r; output s = r + 4m is left in v[4r + m]
#pragma unroll
    for (int r = 0; r < 4; ++r) bfly4<INV>(v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]);
    // un-transpose: out[s] = v[4*(s%4) + s/4]
    float2 u[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) u[s] = v[4 * (s & 3) + (s >> 2)];
#pragma unroll
    for (int s = 0; s < 16; ++s) v[s] = u[s];
}

// v[s] *= w^(j s) for s = 1..15 with correctly rounded twiddles read from a table laid out
// [s-1][j] (j < st); tp already points at column j.  (A product tree from w^1 costs fewer loads
// but its 3-4 ulp twiddle error leaks the loudest bin into quiet ones: measured 8x the error of
// an exact-twiddle FFT on music, see DESIGN.md "Numerics".)
template <bool INV, bool TW_GLOBAL>
__device__ __forceinline__ void apply_twiddles(float2 (&v)[16], const float2* __restrict__ tp, int st) {
#pragma unroll
    for (int s = 1; s < 16; ++s) {
        float2 w = TW_GLOBAL ? __ldg(tp + (s - 1) * st) : tp[(s - 1) * st];
        if (INV) w.y = -w.y;
        v[s] = cmul(v[s], w);
    }
}

template <int LOG2N>
struct FftPlan {
    static constexpr int N = 1 << LOG2N;
    static constexpr int T = N / 16;                       // threads per transform
    static constexpr int NPASS = LOG2N / 4;                // radix-16 passes
    static constexpr int M = N >> (4 * NPASS);             // leftover radix (lanes)
    static_assert(M == 1 || M == 2 || M == 4, "supported sizes: 2^{8,9,10,12,13,14}");
    static_assert(T >= 32 || T == 16, "group must be whole warps");
    // shared memory (float2 elements) for the exchange buffer, with padding
    static constexpr int SMEM_ELEMS = N + N / 16;
    __host__ __device__ static constexpr int stride(int p) { return N >> (4 * (p + 1)); }
    // twiddle table: pass p occupies 15 * stride(p) float2 at tw_offset(p), laid out [s-1][j]
    // with value exp(-2 pi i * j * s / (16 * stride(p)))
    __host__ __device__ static constexpr int tw_offset(int p) {
        int o = 0;
        for (int q = 0; q < p; ++q) o += 15 * stride(q);
        return o;
    }
    static constexpr int TW_ELEMS = tw_offset(NPASS);
};

// logical element index held by thread t, register i during pass with stride st
__device__ __forceinline__ int pass_pos(int t, int i, int st) {
    return (t / st) * (16 * st) + (t % st) + i * st;
}
// padding that makes the exchange next to a pass of (fine) stride st bank-conflict free:
// st float2 of padding after every 16*st elements (none needed when st >= 16)
__device__ __forceinline__ int pad_idx(int l, int st) {
    return st >= 16 ? l : l + st * (l / (16 * st));
}

template <int T>
__device__ __forceinline__ void group_sync() {
    if constexpr (T <= 32) __syncwarp();
    else __syncthreads();                                  // group == CTA
}

// Forward transform.  In: v[q] = x[t + q*T].  Out: v[i] = spectrum slot (t, i).
// tw: twiddle table (FftPlan::tw_offset layout; global memory when TW_GLOBAL, else shared).
// buf: >= SMEM_ELEMS float2, private to the group.
template <int LOG2N, bool TW_GLOBAL = true>
__device__ __forceinline__ void fft_forward(float2 (&v)[16], int t, float2* __restrict__ buf,
                                            const float2* __restrict__ tw) {
    using P = FftPlan<LOG2N>;
#pragma unroll
    for (int p = 0; p < P::NPASS; ++p) {
        const int st = P::stride(p);
        if (p > 0) {
            // exchange: written with pass p-1 layout, read with pass p layout
            const int stp = P::stride(p - 1);
            if (p > 1) group_sync<P::T>();
#pragma unroll
            for (int i = 0; i < 16; ++i) buf[pad_idx(pass_pos(t, i, stp), st)] = v[i];
            group_sync<P::T>();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = buf[pad_idx(pass_pos(t, i, st), st)];
        }
        fft16<false>(v);
        if (st > 1) apply_twiddles<false, TW_GLOBAL>(v, tw + P::tw_offset(p) + (t % st), st);
    }
    // leftover radix-M across M adjacent lanes (DIF)
    if constexpr (P::M == 2) {
        const bool up = t & 1;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float2 o = make_float2(__shfl_xor_sync(0xffffffffu, v[i].x, 1), __shfl_xor_sync(0xffffffffu, v[i].y, 1));
            v[i] = up ? csub(o, v[i]) : cadd(v[i], o);
        }
    } else if constexpr (P::M == 4) {
        const int j = t & 3;
        const bool up2 = j & 2, up1 = j & 1;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float2 o = make_float2(__shfl_xor_sync(0xffffffffu, v[i].x, 2), __shfl_xor_sync(0xffffffffu, v[i].y, 2));
            float2 r = up2 ? csub(o, v[i]) : cadd(v[i], o);
            if (j == 3) r = mul_mi(r);                     // (a0 - a1) * w4^1
            o = make_float2(__shfl_xor_sync(0xffffffffu, r.x, 1), __shfl_xor_sync(0xffffffffu, r.y, 1));
            v[i] = up1 ? csub(o, r) : cadd(r, o);
        }
    }
}

// Inverse transform (unscaled).  In: v[i] = spectrum slot (t, i).  Out: v[q] = N * x[t + q*T].
template <int LOG2N, bool TW_GLOBAL = true>
__device__ __forceinline__ void fft_inverse(float2 (&v)[16], int t, float2* __restrict__ buf,
                                            const float2* __restrict__ tw) {
    using P = FftPlan<LOG2N>;
    if constexpr (P::M == 2) {
        const bool up = t & 1;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float2 o = make_float2(__shfl_xor_sync(0xffffffffu, v[i].x, 1), __shfl_xor_sync(0xffffffffu, v[i].y, 1));
            v[i] = up ? csub(o, v[i]) : cadd(v[i], o);
        }
    } else if constexpr (P::M == 4) {
        const int j = t & 3;
        const bool up2 = j & 2, up1 = j & 1;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float2 o = make_float2(__shfl_xor_sync(0xffffffffu, v[i].x, 1), __shfl_xor_sync(0xffffffffu, v[i].y, 1));
            float2 r = up1 ? csub(o, v[i]) : cadd(v[i], o);
            if (j == 3) r = mul_pi(r);                     // * conj(w4^1)
            o = make_float2(__shfl_xor_sync(0xffffffffu, r.x, 2), __shfl_xor_sync(0xffffffffu, r.y, 2));
            v[i] = up2 ? csub(o, r) : cadd(r, o);
        }
    }
#pragma unroll
    for (int p = P::NPASS - 1; p >= 0; --p) {
        const int st = P::stride(p);
        if (st > 1) apply_twiddles<true, TW_GLOBAL>(v, tw + P::tw_offset(p) + (t % st), st);
        fft16<true>(v);
        if (p > 0) {
            const int stn = P::stride(p - 1);
            if (p < P::NPASS - 1) group_sync<P::T>();
#pragma unroll
            for (int i = 0; i < 16; ++i) buf[pad_idx(pass_pos(t, i, st), st)] = v[i];
            group_sync<P::T>();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = buf[pad_idx(pass_pos(t, i, stn), st)];
        }
    }
}

// frequency index of spectrum slot (thread t, register i)
template <int LOG2N>
__host__ __device__ inline int freq_of_slot(int t, int i) {
    using P = FftPlan<LOG2N>;
    // position after the last radix-16 pass
    int pos = (t / P::stride(P::NPASS - 1)) * (16 * P::stride(P::NPASS - 1)) + (t % P::stride(P::NPASS - 1)) +
              i * P::stride(P::NPASS - 1);
    int k = 0, mult = 1, rem = pos;
    for (int p = 0; p < P::NPASS; ++p) {
        int st = P::stride(p);
        k += (rem / st) * mult;
        rem %= st;
        mult *= 16;
    }
    if (P::M == 2) k += rem * mult;
    if (P::M == 4) k += (((rem & 1) << 1) | (rem >> 1)) * mult;
    return k;
}

}  // namespace ssb
