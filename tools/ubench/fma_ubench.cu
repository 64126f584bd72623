// fma_ubench.cu -- register-resident FMA throughput microbenchmarks for sm_100a.
// Answers: what is the real FP32 FMA ceiling for (a) immediate-operand FFMA, (b) 3-register FFMA chains,
// (c) an 8x8 outer-product register tile (the SGEMM inner loop, operand reuse), (d) the same with packed
// fma.rn.f32x2 (FFMA2), (e) FP64 DFMA outer product.  Prints TFLOP/s per variant.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void k_imm(float* out, int iters)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++)
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], 1.0000001f, 1e-9f);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 123456789.f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_3reg(float* out, int iters, float x, float y)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++)
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 123456789.f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8x8 outer product, operands rotate in registers (no memory): 64 FFMA per step
__global__ void k_outer(float* out, int iters, float x, float y)
{
    float acc[8][8], a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = x + (float)(threadIdx.x + i); b[i] = y - (float)(i + (threadIdx.x & 7)); }
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[r][c] = 0.f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 8; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
            // perturb operands a little so the compiler cannot hoist (2 extra ops per 64 FMA)
            a[rep] += y; b[rep] += x;
        }
    }
    float s = 0;
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) s += acc[r][c];
    if (s == 123456789.f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b)
{
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// 8x8 outer product with packed FFMA2: acc pairs along c; a duplicated into (a,a) pairs
__global__ void k_outer2(float* out, int iters, float x, float y)
{
    unsigned long long acc[8][4], a2[8], b2[4];
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = x + (float)(threadIdx.x + i); b[i] = y - (float)(i + (threadIdx.x & 7)); }
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0ull;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a2[i] = pack2(a[i], a[i]);
#pragma unroll
            for (int i = 0; i < 4; i++) b2[i] = pack2(b[2 * i], b[2 * i + 1]);
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) fma2(acc[r][c], a2[r], b2[c]);
            a[rep] += y; b[rep] += x;
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) s ^= acc[r][c];
    if (s == 123456789ull) out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

// FFMA2 chains, no packing overhead (pure pipe rate)
__global__ void k_chain2(float* out, int iters, float x, float y)
{
    unsigned long long a[16];
    const unsigned long long xx = pack2(x, x), yy = pack2(y, y);
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = pack2((float)(threadIdx.x + i), (float)i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++)
#pragma unroll
            for (int i = 0; i < 16; i++)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(xx), "l"(yy));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= a[i];
    if (s == 123456789ull) out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

__global__ void k_outer_f64(double* out, int iters, double x, double y)
{
    double acc[8][8], a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = x + (double)(threadIdx.x + i); b[i] = y - (double)(i + (threadIdx.x & 7)); }
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[r][c] = 0.;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 8; c++) acc[r][c] = fma(a[r], b[c], acc[r][c]);
            a[rep] += y; b[rep] += x;
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) s += acc[r][c];
    if (s == 123456789.) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F launch)
{
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CHECK(cudaEventRecord(e0));
        launch();
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main()
{
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float* out; CHECK(cudaMalloc(&out, 1 << 24));
    const int threads = 256;
    for (int bps = 1; bps <= 4; bps *= 2) {   // resident CTAs per SM: 2, 4, 8 warps per scheduler
        const int blocks = nsm * bps;
        const int iters = 20000;
        const double lanes = (double)blocks * threads;
        double ms;
        ms = time_ms([&] { k_imm<<<blocks, threads>>>(out, iters); });
        printf("bps=%d imm-form FFMA chains      : %7.2f TFLOP/s\n", bps, lanes * iters * 64 * 2 / ms / 1e9);
        ms = time_ms([&] { k_3reg<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f); });
        printf("bps=%d 3-reg FFMA chains         : %7.2f TFLOP/s\n", bps, lanes * iters * 64 * 2 / ms / 1e9);
        ms = time_ms([&] { k_chain2<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f); });
        printf("bps=%d FFMA2 chains              : %7.2f TFLOP/s\n", bps, lanes * iters * 64 * 2 * 2 / ms / 1e9);
        if (bps <= 2) {
            ms = time_ms([&] { k_outer<<<blocks, threads>>>(out, iters / 4, 1.0000001f, 1e-9f); });
            printf("bps=%d 8x8 outer product FFMA    : %7.2f TFLOP/s\n", bps, lanes * (iters / 4) * 256 * 2 / ms / 1e9);
            ms = time_ms([&] { k_outer2<<<blocks, threads>>>(out, iters / 4, 1.0000001f, 1e-9f); });
            printf("bps=%d 8x8 outer product FFMA2   : %7.2f TFLOP/s\n", bps, lanes * (iters / 4) * 256 * 2 / ms / 1e9);
            ms = time_ms([&] { k_outer_f64<<<blocks, threads>>>((double*)out, iters / 8, 1.0000001, 1e-9); });
            printf("bps=%d 8x8 outer product DFMA    : %7.2f TFLOP/s\n", bps, lanes * (iters / 8) * 256 * 2 / ms / 1e9);
        }
    }
    return 0;
}
