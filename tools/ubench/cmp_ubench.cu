// Throughput of the ways to count "score > threshold" per lane on sm_100a (warp instructions per cycle per SM):
//   0  set.gt.f32.f32 (FSET.BF) + add.f32x2        1  set.gt.u32.f32 (FSET) + subtract the mask
//   2  plain C  c += (s > t)  (FSETP + predicated / select add)   3  sign bit of (t - s): FADD + (x >> 31) + c
//   4  FFMA baseline (one FFMA per element)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cmp_ubench cmp_ubench.cu ; run: ./cmp_ubench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ float gt_one(float a, float p) { float r; asm volatile("set.gt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(p)); return r; }
__device__ __forceinline__ unsigned gt_mask(float a, float p) { unsigned r; asm volatile("set.gt.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(p)); return r; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void k(float* out, const float* in, int iters)
{
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = in[threadIdx.x + 32 * i];
    float t0 = in[1000], t1 = in[1001], t2 = in[1002], t3 = in[1003];
    unsigned c[4] = {0, 0, 0, 0};
    u64 a[4][2] = {};
    float f[4] = {0, 0, 0, 0};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float t = q == 0 ? t0 : (q == 1 ? t1 : (q == 2 ? t2 : t3));
            if (MODE == 0) {
                a[q][0] = add2(a[q][0], pack2(gt_one(v[0], t), gt_one(v[1], t)));
                a[q][1] = add2(a[q][1], pack2(gt_one(v[2], t), gt_one(v[3], t)));
                a[q][0] = add2(a[q][0], pack2(gt_one(v[4], t), gt_one(v[5], t)));
                a[q][1] = add2(a[q][1], pack2(gt_one(v[6], t), gt_one(v[7], t)));
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; i++) c[q] -= gt_mask(v[i], t);
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) c[q] += (v[i] > t) ? 1u : 0u;
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 8; i++) c[q] += __float_as_uint(__fsub_rn(t, v[i])) >> 31;
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) f[q] = fmaf(v[i], t, f[q]);
            }
        }
        t0 += 1e-9f; t1 -= 1e-9f; t2 += 2e-9f; t3 -= 2e-9f;       // keep the loop body from being hoisted
    }
    float s = 0;
    for (int q = 0; q < 4; q++) {
        float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[q][0])); s += lo + hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[q][1])); s += lo + hi;
        s += (float)c[q] + f[q];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, float* out, const float* in, int nsm, double ghz)
{
    const int iters = 20000, blocks = nsm * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, in, 100);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cmp = (double)blocks * (threads / 32) * iters * 32.0;            // warp-level compares (32 per iteration)
    printf("%-46s %8.3f ms  %.2f warp-compares / cycle / SM (at %.2f GHz)\n", name, ms, cmp / (ms * 1e-3) / (ghz * 1e9) / nsm, ghz);
}

int main()
{
    int nsm, khz; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *in, *out; cudaMalloc(&in, 1 << 16); cudaMalloc(&out, 1 << 22); cudaMemset(in, 0, 1 << 16);
    const double ghz = khz * 1e-6;
    run<4>("FFMA baseline", out, in, nsm, ghz);
    run<0>("FSET.BF + FADD2 (set.gt.f32 + add.f32x2)", out, in, nsm, ghz);
    run<1>("FSET mask + IADD (set.gt.u32)", out, in, nsm, ghz);
    run<2>("c += (s > t)  (compiler's choice)", out, in, nsm, ghz);
    run<3>("sign bit of (t - s): FADD + shift-add", out, in, nsm, ghz);
    return 0;
}
