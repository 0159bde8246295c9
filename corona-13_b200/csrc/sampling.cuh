// sampling.cuh -- random dimensions for the wavefront integrator (device side).
//
//   pointsampler(path, dim)  src/pointsampler.d/halton.c:69-84 / rand.c:48-55
//   halton_sample            ext/halton/halton.h (Gruenschloss' Faure-permuted Halton points, 256 dimensions;
//                            generator rule in ext/halton/halton_gen.py): dimension d uses the d-th prime b,
//                            processes D = digits*lookups base-b digits of the 32-bit path index, least
//                            significant first, acc = acc*b + perm_b[digit], and returns
//                            (float)acc * (float)(0x1.fffffcp-1 / b^D).  Dimension 0 is a bit reversal.
//   points_rand              src/points.d/sfmt.c:436 -- per-thread SFMT-19937 in the reference.  Its streams depend
//                            on which worker thread picks which path ("RNG reproducibility"),
//                            so only its distribution can be matched: here a counter-based generator keyed by
//                            (frame, path index, draw counter) -- independent of which GPU or wave traces the path.
#pragma once
#include <stdint.h>

#define HALTON_DIMS 256

struct HaltonDev
{
  const uint16_t *perm;      // concatenated digit permutations, one block of `base` entries per dimension
  const uint32_t *perm_off;  // [256] offset of the dimension's block
  const uint16_t *base;      // [256] prime
  const uint8_t  *digits;    // [256] number of base-b digits consumed
  const float    *scale;     // [256] (float)(0x1.fffffcp-1 / base^digits)
};

__device__ __forceinline__ float halton_dim0(uint32_t index)
{
  index = __brev(index);
  return __uint_as_float(0x3f800000u | (index >> 9)) - 1.0f;
}

__device__ __forceinline__ float halton_sample_dev(const HaltonDev &H, uint32_t dim, uint32_t index)
{
  if(dim == 0) return halton_dim0(index);
  const uint32_t b = H.base[dim];
  const uint16_t *perm = H.perm + H.perm_off[dim];
  const int D = H.digits[dim];
  uint32_t acc = 0;
  for(int k=0;k<D;k++)
  {
    const uint32_t q = index / b;
    acc = acc*b + perm[index - q*b];
    index = q;
  }
  return (float)acc * H.scale[dim];   // unsigned -> float (round to nearest even), then one multiply, like the C code
}

// counter-based generator: two rounds of a 64-bit mix (splitmix64 finaliser) over (key, counter)
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float counter_rand(uint64_t key, uint64_t index, uint32_t counter)
{
  const uint64_t h = mix64(mix64(key ^ (index*0x9e3779b97f4a7c15ull)) + counter);
  return (float)(uint32_t)(h >> 40) * (1.0f/16777216.0f);   // 24 bits -> [0,1), like genrand_real2f's open upper end
}

struct PointsDev
{
  HaltonDev halton;
  int32_t mode;        // CB_POINTS_RAND / CB_POINTS_HALTON
  uint64_t key;        // frame mix for the counter generator
};

// pointsampler(): `dim` is already rand_beg + i.  stream 0 = dimensions, stream 1 = points_rand draws
__device__ __forceinline__ float point_dim(const PointsDev &P, uint64_t index, int dim)
{
  if(P.mode == 1 && dim < HALTON_DIMS) return halton_sample_dev(P.halton, (uint32_t)dim, (uint32_t)index);
  return counter_rand(P.key, index, (uint32_t)dim);
}
__device__ __forceinline__ float point_mt(const PointsDev &P, uint64_t index, uint32_t draw)
{
  return counter_rand(P.key ^ 0x5851f42d4c957f2dull, index, draw);
}
