// umma_probe.cu -- pins the tcgen05 (UMMA) descriptor encodings used by filter_select.cuh:
// D[128x128] (fp32, TMEM) = A[128xK] * B[128xK]^T with bf16 operands in the no-swizzle K-major
// "interleaved" shared-memory layout  [k/8][row][8 elements]  (core matrix = 8 rows x 16 bytes contiguous).
// Usage: umma_probe <K> <variant>   variant bit0: swap LBO/SBO, bit1: omit the descriptor version bit
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes, int version)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    if (version) d |= 1ull << 46;
    return d;   // layout_type (bits 61..63) = 0: no swizzle
}

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* __restrict__ Ap, const __nv_bfloat16* __restrict__ Bp,
                                                 float* __restrict__ D, int K, unsigned lbo, unsigned sbo, int version)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* Bs = As + 128 * K;
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < 128 * K / 8; i += 128) {
        reinterpret_cast<uint4*>(As)[i] = reinterpret_cast<const uint4*>(Ap)[i];
        reinterpret_cast<uint4*>(Bs)[i] = reinterpret_cast<const uint4*>(Bp)[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;

    if (tid == 0) {
        // instruction descriptor: D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), K-major both, N=128 (>>3 at 17), M=128 (>>4 at 24)
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < K / 16; ks++) {
            // one MMA consumes K=16 = two 16-byte k-chunks; chunk c of the tile starts at c * (128 rows * 16 B)
            const uint64_t ad = make_desc(smem_u32(As) + ks * 2 * 2048, lbo, sbo, version);
            const uint64_t bd = make_desc(smem_u32(Bs) + ks * 2 * 2048, lbo, sbo, version);
            const unsigned acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs
    {
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        unsigned r[32];
        const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; c++) D[(size_t)(warp * 32 + lane) * 128 + c0 + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
}

int main(int argc, char** argv)
{
    const int K = argc > 1 ? atoi(argv[1]) : 64;
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    std::vector<float> A(128 * K), B(128 * K);
    srand(1);
    for (auto& v : A) v = (float)(rand() % 7 - 3);
    for (auto& v : B) v = (float)(rand() % 5 - 2);
    // pack: [k/8][row][8]
    std::vector<__nv_bfloat16> Ap(128 * K), Bp(128 * K);
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < K; k++) {
            Ap[((size_t)(k / 8) * 128 + r) * 8 + (k % 8)] = __float2bfloat16(A[r * K + k]);
            Bp[((size_t)(k / 8) * 128 + r) * 8 + (k % 8)] = __float2bfloat16(B[r * K + k]);
        }
    __nv_bfloat16 *dA, *dB; float* dD;
    CHECK(cudaMalloc(&dA, Ap.size() * 2)); CHECK(cudaMalloc(&dB, Bp.size() * 2)); CHECK(cudaMalloc(&dD, 128 * 128 * 4));
    CHECK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemset(dD, 0xff, 128 * 128 * 4));
    unsigned lbo = 128 * 16, sbo = 128;     // k-chunk stride, 8-row-group stride
    if (variant & 1) { unsigned t = lbo; lbo = sbo; sbo = t; }
    const int version = (variant & 2) ? 0 : 1;
    const size_t smem = 2 * 128 * K * 2 + 1024;
    CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 128, smem>>>(dA, dB, dD, K, lbo, sbo, version);
    CHECK(cudaGetLastError());
    CHECK(cudaDeviceSynchronize());
    std::vector<float> D(128 * 128);
    CHECK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < 128; j++) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += (double)A[i * K + k] * B[j * K + k];
            const double e = fabs(ref - D[i * 128 + j]);
            if (!(e <= 1e-3)) bad++;
            if (e > maxerr || e != e) maxerr = e;
        }
    printf("K=%d variant=%d (lbo=%u sbo=%u version=%d): mismatches=%d of 16384, max abs err=%g  D[0][0..3]=%g %g %g %g\n",
           K, variant, lbo, sbo, version, bad, maxerr, D[0], D[1], D[2], D[3]);
    return bad ? 1 : 0;
}
