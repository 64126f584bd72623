// tile_ubench.cu -- shared-memory-fed 8x8 register-tile FMA loops (the scoring micro-kernel) in several
// instruction forms, to find the achievable ceiling of the inner loop on sm_100a before any pipeline,
// barrier or epilogue cost.  256 threads, CTA tile 128x128, operands read from a static smem slab.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int BK = 16, BM = 128, BN = 128;
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void fma2(u64& d, u64 a, u64 b) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

template <int VARIANT, int MINB>
__global__ void __launch_bounds__(256, MINB) k_tile(float* out, const float* in, int iters)
{
    __shared__ __align__(16) float As[BK * BM];
    __shared__ __align__(16) float Bs[BK * BN];
    for (int i = threadIdx.x; i < BK * BM; i += 256) { As[i] = in[i]; Bs[i] = in[BK * BM + i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_base = (warp & 3) * 32 + (lane >> 3) * 4;
    const int col_base = (warp >> 2) * 64 + (lane & 7) * 4;
    float s = 0.f;
    if (VARIANT == 0) {
        float acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) acc[r][c] = 0.f;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int kk = 0; kk < BK; kk++) {
                float a[8], b[8];
                *(float4*)&a[0] = *(const float4*)&As[kk * BM + row_base];
                *(float4*)&a[4] = *(const float4*)&As[kk * BM + row_base + 16];
                *(float4*)&b[0] = *(const float4*)&Bs[kk * BN + col_base];
                *(float4*)&b[4] = *(const float4*)&Bs[kk * BN + col_base + 32];
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
            }
        }
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) s += acc[r][c];
    } else {
        u64 acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = 0ull;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int kk = 0; kk < BK; kk++) {
                float a[8], b[8];
                *(float4*)&a[0] = *(const float4*)&As[kk * BM + row_base];
                *(float4*)&a[4] = *(const float4*)&As[kk * BM + row_base + 16];
                *(float4*)&b[0] = *(const float4*)&Bs[kk * BN + col_base];
                *(float4*)&b[4] = *(const float4*)&Bs[kk * BN + col_base + 32];
                u64 b2[4];
#pragma unroll
                for (int c = 0; c < 4; c++) b2[c] = pack2(b[2 * c], b[2 * c + 1]);
                if (VARIANT == 1) {
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        const u64 a2 = pack2(a[r], a[r]);
#pragma unroll
                        for (int c = 0; c < 4; c++) fma2(acc[r][c], a2, b2[c]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int r = 0; r < 8; r++) fma2(acc[r][c], pack2(a[r], a[r]), b2[c]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) { float lo, hi; unpack2(acc[r][c], lo, hi); s += lo + hi; }
    }
    if (s == 123456789.f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// 16 warps x (4 users x 8 items) per thread: warp owns 8 rows x 128 cols
template <int MINB>
__global__ void __launch_bounds__(512, MINB) k_tile48(float* out, const float* in, int iters)
{
    __shared__ __align__(16) float As[BK * BM];
    __shared__ __align__(16) float Bs[BK * BN];
    for (int i = threadIdx.x; i < BK * BM; i += 512) { As[i] = in[i]; Bs[i] = in[BK * BM + i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_base = warp * 8 + (lane >> 4) * 4;
    const int col_base = (lane & 15) * 4;
    u64 acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0ull;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            const float4 a4 = *(const float4*)&As[kk * BM + row_base];
            const ulonglong2 b0 = *(const ulonglong2*)&Bs[kk * BN + col_base];
            const ulonglong2 b1 = *(const ulonglong2*)&Bs[kk * BN + col_base + 64];
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const u64 b2[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 4; r++) fma2(acc[r][c], pack2(a[r], a[r]), b2[c]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) { float lo, hi; unpack2(acc[r][c], lo, hi); s += lo + hi; }
    if (s == 123456789.f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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

template <int V, int MINB>
void run(const char* name, float* out, const float* in, int nsm)
{
    const int iters = 4000;
    const int blocks = nsm * MINB;
    double ms = time_ms([&] { k_tile<V, MINB><<<blocks, 256>>>(out, in, iters); });
    CHECK(cudaGetLastError());
    printf("%-44s CTAs/SM=%d : %7.2f TFLOP/s\n", name, MINB, (double)blocks * 256 * iters * BK * 64 * 2 / ms / 1e9);
}

int main()
{
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in;
    CHECK(cudaMalloc(&out, 1 << 24)); CHECK(cudaMalloc(&in, 1 << 20)); CHECK(cudaMemset(in, 0, 1 << 20));
    run<0, 1>("FFMA  8x8 (r outer)", out, in, nsm);
    run<0, 2>("FFMA  8x8 (r outer)", out, in, nsm);
    run<1, 1>("FFMA2 8x8 pairs along c, r outer", out, in, nsm);
    run<1, 2>("FFMA2 8x8 pairs along c, r outer", out, in, nsm);
    run<2, 1>("FFMA2 8x8 pairs along c, c outer", out, in, nsm);
    run<2, 2>("FFMA2 8x8 pairs along c, c outer", out, in, nsm);
    {
        const int iters = 4000;
        double ms = time_ms([&] { k_tile48<1><<<nsm, 512>>>(out, in, iters); });
        CHECK(cudaGetLastError());
        printf("%-44s CTAs/SM=%d : %7.2f TFLOP/s\n", "FFMA2 4x8 per thread, 16 warps", 1, (double)nsm * 512 * iters * BK * 32 * 2 / ms / 1e9);
    }
    return 0;
}
