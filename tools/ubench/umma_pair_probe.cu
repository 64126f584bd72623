// umma_pair_probe.cu -- round-2 groundwork.  Run on a B200 at the end of round 1 (profiles/r01w_umma_pair_probe.txt):
// variant 0 reproduces D exactly for K = 64 and K = 128, variant 1 (halves swapped) returns the two column halves
// exchanged -- i.e. everything listed below holds as written, and CTA r of the pair supplies items r*64 .. r*64+63.
// DESIGN.md section 4 "Where the filter kernel's time goes": the filter's pipeline is bound by L2 ->
// shared-memory traffic (every CTA pulls every 32 KB item tile for its 128 users).  A CTA pair (tcgen05 cta_group::2, two
// SMs of one TPC) runs ONE MMA of M = 256 users x N = 128 items per k-step with each CTA holding its own 128 user rows and
// only HALF of the item tile, which halves that traffic while each CTA's accumulators (its 128 rows x 128 columns of TMEM)
// and therefore the whole epilogue stay as they are.  Before filter_select.cuh is rebuilt around that, this probe pins what
// the guides do not spell out for the no-swizzle K-major "interleaved" layout used there:
//   * which half of B (N) each CTA of the pair has to hold (variant bit0: CTA r holds items r*64.. / swapped),
//   * the descriptor strides of a 64-row B half (LBO = 64 rows * 16 B),
//   * that smem descriptors are CTA-relative offsets applied to both CTAs, tcgen05.alloc.cta_group::2 by one warp of each
//     CTA, commit with .multicast::cluster to the same barrier offset in both CTAs, TMEM lanes = the CTA's own 128 rows.
// D[256 x 128] (fp32) = A[256 x K] * B[128 x K]^T, bf16 operands, checked on the host.
// Usage: umma_pair_probe <K> <variant>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
           (1ull << 46);                                   // version 1, no swizzle
}

// Ap: [2][K/8][128][8] (CTA r's 128 user rows), Bp: [2][K/8][64][8] (the two 64-item halves), D: [256][128]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe(const __nv_bfloat16* __restrict__ Ap, const __nv_bfloat16* __restrict__ Bp, float* __restrict__ D, int K, int swap_b)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem);          // 128 x K
    __nv_bfloat16* Bs = As + 128 * K;                                    // 64 x K: this CTA's half of the item tile
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned rank = cluster_ctarank();
    const unsigned bhalf = swap_b ? (rank ^ 1u) : rank;

    for (int i = tid; i < 128 * K / 8; i += 128) reinterpret_cast<uint4*>(As)[i] = reinterpret_cast<const uint4*>(Ap + (size_t)rank * 128 * K)[i];
    for (int i = tid; i < 64 * K / 8; i += 128) reinterpret_cast<uint4*>(Bs)[i] = reinterpret_cast<const uint4*>(Bp + (size_t)bhalf * 64 * K)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the tensor core
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                                     // one warp of EACH CTA of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();                                                      // both CTAs: operands in place, barriers initialised, TMEM allocated
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;

    if (rank == 0 && tid == 0) {                                         // only the leader CTA issues
        // D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), K-major both, N = 128 (>>3 at 17), M = 256 (>>4 at 24)
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
        for (int ks = 0; ks < K / 16; ks++) {
            const uint64_t ad = make_desc(smem_u32(As) + ks * 2 * (128 * 16), 128 * 16, 128);   // A: 128 rows per k-chunk
            const uint64_t bd = make_desc(smem_u32(Bs) + ks * 2 * (64 * 16), 64 * 16, 128);     // B half: 64 rows per k-chunk
            const unsigned acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        // completion of everything issued so far -> the barrier at this offset in BOTH CTAs
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((unsigned short)3) : "memory");
    }
    {
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {                               // this CTA's 128 rows x 128 columns
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
        for (int c = 0; c < 32; c++) D[(size_t)(rank * 128 + warp * 32 + lane) * 128 + c0 + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();                                                      // neither CTA may leave (or free TMEM) while the other still reads
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
}

int main(int argc, char** argv)
{
    const int K = argc > 1 ? atoi(argv[1]) : 64;
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    if (K % 16 || K <= 0 || K > 256) { printf("K must be a multiple of 16 in (0, 256]\n"); return 2; }
    std::vector<float> A(256 * K), B(128 * K);
    srand(1);
    for (auto& v : A) v = (float)(rand() % 7 - 3);
    for (auto& v : B) v = (float)(rand() % 5 - 2);
    std::vector<__nv_bfloat16> Ap(256 * K), Bp(128 * K);
    for (int r = 0; r < 256; r++)                           // [cta][k/8][128][8]
        for (int k = 0; k < K; k++)
            Ap[(size_t)(r / 128) * 128 * K + ((size_t)(k / 8) * 128 + (r % 128)) * 8 + (k % 8)] = __float2bfloat16(A[r * K + k]);
    for (int r = 0; r < 128; r++)                           // [half][k/8][64][8]
        for (int k = 0; k < K; k++)
            Bp[(size_t)(r / 64) * 64 * K + ((size_t)(k / 8) * 64 + (r % 64)) * 8 + (k % 8)] = __float2bfloat16(B[r * K + k]);
    __nv_bfloat16 *dA, *dB; float* dD;
    CHECK(cudaMalloc(&dA, Ap.size() * 2)); CHECK(cudaMalloc(&dB, Bp.size() * 2)); CHECK(cudaMalloc(&dD, 256 * 128 * 4));
    CHECK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemset(dD, 0xff, 256 * 128 * 4));
    const size_t smem = (size_t)(128 + 64) * K * 2 + 1024;
    CHECK(cudaFuncSetAttribute(pair_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pair_probe<<<2, 128, smem>>>(dA, dB, dD, K, variant & 1);
    CHECK(cudaGetLastError());
    CHECK(cudaDeviceSynchronize());
    std::vector<float> D(256 * 128);
    CHECK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < 256; i++)
        for (int j = 0; j < 128; j++) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += (double)A[i * K + k] * B[j * K + k];
            const double e = fabs(ref - D[i * 128 + j]);
            if (!(e <= 1e-3)) bad++;
            if (e > maxerr || e != e) maxerr = e;
        }
    printf("pair probe K=%d variant=%d: mismatches=%d of 32768, max abs err=%g  D[0][0..3]=%g %g %g %g  D[128][0..1]=%g %g  D[0][64..65]=%g %g\n",
           K, variant, bad, maxerr, D[0], D[1], D[2], D[3], D[128 * 128], D[128 * 128 + 1], D[64], D[65]);
    return bad ? 1 : 0;
}
