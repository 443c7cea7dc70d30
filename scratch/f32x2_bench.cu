// microbenchmark: FADD vs FADD2 / FFMA vs FFMA2 issue throughput on sm_100a
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm volatile("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
template <int MODE>
__global__ void k(float2* out, int iters, float2 seed) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    float2 b = make_float2(seed.y, seed.x);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }            // 2 FADD
            if (MODE == 1) a[i] = add2(a[i], b);                                        // 1 FADD2
            if (MODE == 2) { a[i].x = fmaf(a[i].x, b.x, b.y); a[i].y = fmaf(a[i].y, b.x, b.y); }  // 2 FFMA
            if (MODE == 3) a[i] = fma2(a[i], b, b);                                     // 1 FFMA2
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(const char* name, float2* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(d, 16, make_float2(1.f, 2.f));
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, make_float2(1.f, 2.f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)blocks * threads * iters * 8 * 2 * (MODE >= 2 ? 2 : 1);
    printf("%-8s %.3f ms  %.1f TFLOP/s  (%.2f G pair-ops/s)\n", name, ms, flops / ms / 1e9, (double)blocks * threads * iters * 8 / ms / 1e6);
    return ms;
}
int main() {
    float2* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float2));
    run<0>("2xFADD", d); run<1>("FADD2", d); run<2>("2xFFMA", d); run<3>("FFMA2", d);
    return 0;
}
