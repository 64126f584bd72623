// tie_noise.cuh -- the reference's per-user tie-breaking noise (break_ties_with_noise, its API default):
//     std::mt19937 rng_user(seed + (uint64_t)user);
//     std::uniform_real_distribution<real_t> runif((real_t)(-1e-12), (real_t)1e-12);
//     for (ix < move_to) pred[ind[ix]] += runif(rng_user);          /root/reference/src/recometrics.hpp:531-534
// i.e. candidate number ix of the user (candidates in ascending item order) gets draw number ix of the user's stream.
// Restated for the device from the published algorithms: MT19937 (624-word state, seeded with value mod 2^32) and
// libstdc++ 13's uniform_real_distribution = generate_canonical * (b - a) + a (one 32-bit word per float draw, two per
// double draw; the reference's C++ build contracts the last step into one fma).  oracle/ holds the same restatement in
// C, pinned bit-for-bit against the compiled reference on inputs whose ranking only the noise decides (tests/golden).
//
// A draw cannot be reached without stepping the generator through everything before it, so one warp walks the user's
// stream block by block (624 words each: three dependent thirds, computed 32 words at a time in shared memory) and
// picks out the words of the candidates it is interested in.
#pragma once
#include "score_select.cuh"

namespace rmb {

constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ unsigned mt_temper(unsigned z)
{
    z ^= (z >> 11);
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= (z >> 18);
    return z;
}

// one warp: state after std::mt19937(value) is constructed (no output drawn yet), x = 624 words of shared memory
__device__ __forceinline__ void mt_seed_warp(unsigned* x, const unsigned long long value, const int lane)
{
    if (lane == 0) {
        unsigned prev = (unsigned)(value & 0xffffffffull);
        x[0] = prev;
        for (int i = 1; i < MT_N; i++) { prev = 1812433253u * (prev ^ (prev >> 30)) + (unsigned)i; x[i] = prev; }
    }
    __syncwarp();
}

// one warp: the next 624 untempered words, in place.  Word k needs the OLD words k, k+1 and k+397 (k < 227) or the NEW
// word k-227 (k >= 227; the last one also the new word 0): 32 consecutive words at a time never depend on each other.
__device__ __forceinline__ void mt_twist_warp(unsigned* x, const int lane)
{
    for (int base = 0; base < MT_N; base += 32) {
        const int k = base + lane;
        unsigned nv = 0;
        if (k < MT_N) {
            const unsigned xk = x[k], xk1 = x[k + 1 == MT_N ? 0 : k + 1], xm = x[k + MT_M < MT_N ? k + MT_M : k + MT_M - MT_N];
            const unsigned y = (xk & 0x80000000u) | (xk1 & 0x7fffffffu);
            nv = xm ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp();
        if (k < MT_N) x[k] = nv;
        __syncwarp();
    }
}

template <typename T> struct TieNoise;
template <> struct TieNoise<float> {
    static constexpr int WORDS = 1;                    // generate_canonical<float, 24>: one word
    // adding the noise to a float score of this magnitude or more gives the score back (|noise| <= 1e-12 < ulp / 2)
    __device__ __forceinline__ static bool can_change(const float s) { return fabsf(s) < 6.103515625e-05f; }   // 2^-14
    __device__ __forceinline__ static float draw(const unsigned w0, const unsigned)
    {
        float ret = __uint2float_rn(w0) * 2.3283064365386963e-10f;                 // / 2^32
        if (ret >= 1.f) ret = __uint_as_float(0x3f7fffffu);                        // nextafter(1, 0)
        return fmaf(ret, 1e-12f - (-1e-12f), -1e-12f);
    }
};
template <> struct TieNoise<double> {
    static constexpr int WORDS = 2;                    // generate_canonical<double, 53>: two words, low one first
    __device__ __forceinline__ static bool can_change(const double s) { return fabs(s) < 32768.; }
    __device__ __forceinline__ static double draw(const unsigned w0, const unsigned w1)
    {
        const double sum = (double)w0 + (double)w1 * 4294967296.0;
        double ret = sum * 5.421010862427522170e-20;                                 // / 2^64
        if (ret >= 1.) ret = __longlong_as_double(0x3fefffffffffffffll);            // nextafter(1, 0)
        return fma(ret, 1e-12 - (-1e-12), -1e-12);
    }
};

}  // namespace rmb
