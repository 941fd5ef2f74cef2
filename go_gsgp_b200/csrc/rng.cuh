// rng.cuh -- counter-based RNG of the on-device selection / random-tree kernel.
//
// Philox4x32-10 (Salmon et al., SC'11) keyed by the run seed; one stream per (generation,
// individual) so every individual's draws are independent of the others' -- the property the
// reference's sequential math/rand stream (main_cpu.go:1367,1451) lacks.  The derived draws restate
// Go math/rand v1 (Int63, Int31, Float64 with redraw on 1.0, Intn -> Int31n with rejection) so that
// Float64/Intn mean the same thing they mean at main_cpu.go:459,468-473,588,626,832-834,919,1451.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GSGP_HD __host__ __device__ __forceinline__
#else
#define GSGP_HD inline
#endif

namespace gsgp {

GSGP_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                           uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct PhiloxStream {
    uint32_t k0, k1;        // key = seed
    uint32_t s0, s1, s2;    // stream id = counter words 1..3 (generation, individual, 0)
    uint32_t block;         // counter word 0
    uint32_t out[4];
    int have;               // unread 64-bit halves in out[]

    GSGP_HD void init(int64_t seed, uint32_t a, uint32_t b, uint32_t c)
    {
        k0 = (uint32_t)((uint64_t)seed); k1 = (uint32_t)((uint64_t)seed >> 32);
        s0 = a; s1 = b; s2 = c; block = 0; have = 0;
    }
    GSGP_HD uint64_t uint64()
    {
        if (have == 0) { philox4x32_10(block++, s0, s1, s2, k0, k1, out); have = 2; }
        const uint64_t v = (have == 2) ? ((uint64_t)out[0] | ((uint64_t)out[1] << 32))
                                       : ((uint64_t)out[2] | ((uint64_t)out[3] << 32));
        --have;
        return v;
    }
    GSGP_HD int64_t int63() { return (int64_t)(uint64() & 0x7fffffffffffffffull); }
    GSGP_HD int32_t int31() { return (int32_t)(int63() >> 32); }
    GSGP_HD double float64()
    {
        for (;;) {
            const double f = (double)int63() / 9223372036854775808.0;   // round-to-nearest, as Go
            if (f != 1.0) return f;
        }
    }
    GSGP_HD int32_t intn(int32_t n)
    {
        if ((n & (n - 1)) == 0) return int31() & (n - 1);
        const int32_t mx = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n);
        int32_t v = int31();
        while (v > mx) v = int31();
        return v % n;
    }
};

}  // namespace gsgp
