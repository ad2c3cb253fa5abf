/*
 * include/drt_rng.h -- counter-based per-path random streams.
 *
 * Replaces the reference's single global libc stream
 *   f64 rng() { return (f64)rand() / (f64)RAND_MAX; }        (src/rng.c:2-7)
 * with one independent stream per camera path, so that paths can be traced in any
 * order on any number of GPUs and still draw the same numbers:
 *
 *   stream key  = (pixel index y*W+x, global sample index)
 *   counter     = (draw_index / 4, 0, seed_lo, seed_hi)
 *   generator   = Philox4x32-10 (Salmon et al., SC'11), one 32-bit word per draw
 *   r31         = word >> 1                      in [0, 2^31-1]  == glibc rand() range
 *   rng()       = (f64)r31 / 2147483647.0        in [0,1] INCLUSIVE, same lattice as
 *                                                 rand()/RAND_MAX with glibc's RAND_MAX
 *
 * The same header is compiled by gcc (host C, oracle) and nvcc (device), so the oracle
 * and the CUDA path cannot disagree about the bits.
 */
#ifndef DRT_RNG_H
#define DRT_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define DRT_HD __host__ __device__ __forceinline__
#else
#define DRT_HD static inline
#endif

#define DRT_PHILOX_M0 0xD2511F53u
#define DRT_PHILOX_M1 0xCD9E8D57u
#define DRT_PHILOX_W0 0x9E3779B9u
#define DRT_PHILOX_W1 0xBB67AE85u
#define DRT_RAND_MAX  2147483647

typedef struct
{
    uint32_t key0, key1;       /* pixel index, global sample index */
    uint32_t seed_lo, seed_hi; /* render-wide seed */
    uint32_t draws;            /* number of 32-bit words consumed so far */
    uint32_t buf[4];           /* current Philox block */
} drt_rng_stream;

DRT_HD void drt_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#if defined(__CUDA_ARCH__) && defined(DRT_PHILOX_ROLLED)
#pragma unroll 2   /* the render kernels of mixed scenes are bound by instruction fetch: a rolled loop is 1.5 KB less code */
#endif
    for(int round = 0; round < 10; round += 1)
    {
#if defined(__CUDA_ARCH__)
        uint32_t hi0 = __umulhi(DRT_PHILOX_M0, c0), lo0 = DRT_PHILOX_M0 * c0;
        uint32_t hi1 = __umulhi(DRT_PHILOX_M1, c2), lo1 = DRT_PHILOX_M1 * c2;
#else
        uint64_t p0 = (uint64_t)DRT_PHILOX_M0 * c0, p1 = (uint64_t)DRT_PHILOX_M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += DRT_PHILOX_W0; k1 += DRT_PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

DRT_HD void drt_rng_begin(drt_rng_stream *s, uint64_t seed, uint32_t pixel, uint32_t sample)
{
    s->key0 = pixel; s->key1 = sample;
    s->seed_lo = (uint32_t)seed; s->seed_hi = (uint32_t)(seed >> 32);
    s->draws = 0;
    s->buf[0] = s->buf[1] = s->buf[2] = s->buf[3] = 0;
}

/* Next value of the stream in glibc rand()'s range [0, RAND_MAX]. */
DRT_HD uint32_t drt_rng_next31(drt_rng_stream *s)
{
    uint32_t lane = s->draws & 3u;
    if(lane == 0) drt_philox4x32_10(s->draws >> 2, 0u, s->seed_lo, s->seed_hi, s->key0, s->key1, s->buf);
    s->draws += 1;
    uint32_t w = (lane == 0) ? s->buf[0] : (lane == 1) ? s->buf[1] : (lane == 2) ? s->buf[2] : s->buf[3];
    return w >> 1;
}

/* rng() of src/rng.c:2-7 on the per-path stream. */
DRT_HD double drt_rng_f64(drt_rng_stream *s)
{
    return (double)drt_rng_next31(s) / (double)DRT_RAND_MAX;
}

#endif
