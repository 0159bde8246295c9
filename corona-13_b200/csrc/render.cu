// render.cu -- wavefront pt / ptdl integrator (C ABI in include/corona_b200_render.h).
//
// One progression = path indices [first, first+count).  The pool holds up to `batch_paths` paths; every wave tops it up with
// new paths behind the survivors of the previous wave (streaming: paths that outlive their progression ride along with the
// next one, cb200_render_flush finishes them):
//
//   k_pixel_keys + radix sort + k_path_start
//                    path_extend at length 0 (src/pathspace.c:205-250): lambda, time, thin-lens camera sample
//                    (src/camera.d/thinlens.c:68-128) -> first ray; the wave's new indices are processed in pixel-Morton order
//   k_intersect      closest hit (traverse.cu) = accel_intersect in path_propagate (pathspace.c:763)
//   k_compact_hits   slots whose ray hit something -> one index list per BSDF kind (material-sorted shading); rays that ended
//                    at their sampled free-flight distance inside a medium form a fourth list (volume vertices)
//   k_shade<kind>    rest of path_propagate + path_extend bookkeeping + the sampler (pt.c:40-54 / ptdl.c:112-150):
//                    vertex preparation, emission splat with MIS, next-event sample (nee.h:87-243) -> shadow ray,
//                    bsdf sample (shader.c:577-590) -> next ray, stream compaction of surviving paths
//   k_visible<SHADOW> shadow rays in path_visible's terms (pathspace.c:311-344), any-hit sweep (traverse.cu)
//   k_nee_resolve    splat of the visible next-event contributions
//
// Splat: spectrum_p_to_camera + 4x4 Blackman-Harris footprint (view.c:455-495, blackmanharris.h:43-77) with
// float atomics into the W*H*3 accumulation buffer (the reference uses CAS loops the same way).
// Scope: surfaces, nested dielectrics, homogeneous participating media (interior / exterior, Henyey-Greenstein), geometric
// lights, the built-in `black' and `cloudy' skies and the constant-colour sky module.  Heterogeneous media, volume lights and
// the other sky modules (daylight, environment maps) are SURVEY 8(f).
#include "shading.cuh"
#include <cub/cub.cuh>
#include <vector>
#include <string>
#include <cstring>
#include <cmath>
#include <cstdlib>

#define RB 128   // block size of the integrator kernels
#ifndef CB200_VEC_ATOMICS
#define CB200_VEC_ATOMICS 1
#endif

struct __align__(16) PathState   // 128 bytes, one per path in flight
{
  float x[3];        float time;        // vertex v position (un-offset)
  float omega[3];    float lambda;      // e[v+1].omega as sampled
  float thr, thr_prev, pdf_proj, cos_prev;   // throughput into v+1, throughput into v, bsdf pdf (proj. solid angle), |n_v . omega|
  float pixel_i, pixel_j, scramble, cur_ior;
  uint32_t prim_lo, prim_hi;            // primitive of vertex v (ignore / self-intersection test)
  uint32_t index_lo, index_hi;          // path index
  uint32_t lre;                         // bits 0-5 path->length (finished vertices), 6-15 rand_beg of the vertex the current ray is
                                        // about to create, 16-31 binary exponent of the path pdf (PathPdf)
  float pdf_m;                          // mantissa of the path pdf: the product of v[1..length-1].pdf the reference's
                                        // sampler_mis() forms in double precision (pt.c:30-38, ptdl.c:78-88)
  uint32_t bits;                        // bit0: nee_possible(v) (material_modes & (diffuse|glossy)); bits 8..: mt draw counter
  int32_t med_n;                        // media stack: entry count and the entries' medium ids (media_pack)
  uint32_t med_shape[MED_MAX];
  float med_ior[MED_MAX];
};
static_assert(sizeof(PathState) == 128, "PathState size");

struct NeeRec   // pending next-event contribution, resolved after the shadow wave
{
  float value;       // throughput * mis weight (spectral, one wavelength)
  float lambda;
  float pixel_i, pixel_j;
  float total_dist;
  uint32_t light_lo, light_hi;
  uint32_t len;      // vertices of the completed path (path->length at the splat: the vertex + the light point)
};

// The reference weighs every contribution with sampler_mis(): own pdf and competing pdf are each multiplied by the product of
// ALL previous vertex pdfs (vertex area measure) in double precision and then converted to float for the division
// (pt.c:30-38, ptdl.c:78-88).  The conversion overflows to inf beyond 3.4e38 and flushes to 0 below 1.4e-45, the quotient becomes
// NaN or 0 and view_splat drops the sample: paths whose pdf product leaves the float range do not contribute -- in dense
// participating media (vertex pdfs of 1e7 and more) that is every path after a handful of scattering events, and it halves
// the brightness of 0030_subsurf-style skin.  A drop-in has to do the same, so the product rides along with the path as
// mantissa (float, [0.5,1)) and exponent: exact in range where a double would have overflowed, sticky like a double there.
struct PathPdf
{
  float m; int e;     // value = m * 2^e; e == PP_INF: +inf, e == PP_ZERO: 0, m != m: NaN
};
#define PP_INF 32767
#define PP_ZERO (-32768)
__device__ __forceinline__ PathPdf pp_one() { PathPdf p; p.m = 0.5f; p.e = 1; return p; }
__device__ __forceinline__ PathPdf pp_mul(PathPdf p, float x)     // p *= (double)x with IEEE double range semantics
{
  const uint32_t xb = __float_as_uint(x), xe = (xb >> 23) & 0xffu;
  if(xe != 0u && xe != 0xffu && !(xb >> 31) && p.m == p.m && p.e != PP_INF && p.e != PP_ZERO)
  { // the usual case: a normal positive float times a finite product
    float mm = p.m*__uint_as_float((xb & 0x007fffffu) | 0x3f000000u);   // [0.5,1) x [0.5,1)
    int e = p.e + (int)xe - 126;
    if(mm < 0.5f) { mm *= 2.0f; e--; }
    p.m = mm; p.e = e > 1024 ? PP_INF : (e < -1073 ? PP_ZERO : e);
    return p;
  }
  if(p.m != p.m) return p;
  if(x != x) { p.m = x; return p; }
  if(p.e == PP_INF) { if(x == 0.0f) p.m = __int_as_float(0x7fc00000); return p; }
  if(p.e == PP_ZERO) { if(isinf(x)) p.m = __int_as_float(0x7fc00000); return p; }
  if(x == 0.0f) { p.e = PP_ZERO; return p; }
  if(isinf(x)) { p.e = PP_INF; return p; }     // (a negative pdf does not occur)
  int ex;
  const float mx = frexpf(fabsf(x), &ex);       // denormal x
  float mm = p.m*mx;
  int e = p.e + ex;
  if(mm < 0.5f) { mm *= 2.0f; e--; }
  p.m = mm; p.e = e > 1024 ? PP_INF : (e < -1073 ? PP_ZERO : e);
  return p;
}
// (float)(p * (double)y): 0 = finite and not zero, 1 = inf, -1 = zero, 2 = NaN
__device__ __forceinline__ int pp_class(PathPdf p, float y)
{
  p = pp_mul(p, y);
  if(p.m != p.m) return 2;
  if(p.e == PP_INF) return 1;
  if(p.e == PP_ZERO) return -1;
  if(p.e > 128) return 1;      // >= 2^128 > FLT_MAX
  if(p.e < -149) return -1;    // < 2^-150: rounds to zero
  return 0;
}
// does a sample with own pdf `own` and competing pdf `other` (both still to be multiplied by the path pdf) survive the
// reference's float(own P) / float(own P + other P)?  (NaN and 0 weights are dropped by view_splat, view.c:457-458)
__device__ __forceinline__ bool pp_contributes(PathPdf p, float own, float other)
{
  // nearly always: a product within 2^+-100 times pdfs within 1e-6 .. 1e6 stays far inside the float range
  if(p.m == p.m && p.e > -100 && p.e < 100 && own > 1e-6f && own + other < 1e6f) return true;
  return pp_class(p, own) == 0 && pp_class(p, own + other) == 0;
}
__device__ __forceinline__ uint32_t lre_pack(int length, int rand_beg, int e) { return (uint32_t)length | ((uint32_t)rand_beg << 6) | ((uint32_t)(e & 0xffff) << 16); }
__device__ __forceinline__ int lre_length(uint32_t w) { return (int)(w & 63u); }
__device__ __forceinline__ int lre_rand_beg(uint32_t w) { return (int)((w >> 6) & 1023u); }
__device__ __forceinline__ int lre_exp(uint32_t w) { return (int)(int16_t)(w >> 16); }

struct CameraDev
{
  cb_camera_t c;
  float width, height;
  float fstop, exposure_time;
};

struct EnvDev   // sky_envmap.c's rgbe_t on the device
{
  const float4 *px;          // width*height texels: rgb2spec coefficients + scale
  const float *mip;          // importance hierarchy, level k at mip + mip_off[k], (w2n >> k) x (h2n >> k)
  uint32_t mip_off[16];
  int32_t width, height, w2n, h2n, levels;
  float aspectx, aspecty, mul;
  double sum;
  float world[9], world_inv[9];
};

struct RenderDev
{
  DevAccel accel;
  SceneGeo geo;
  MaterialsDev mats;
  LightsDev lights;
  PointsDev points;
  CameraDev cam;
  float *fb;
  uint32_t fb_w, fb_h;
  int32_t sampler, colour, max_path_len;
  int32_t sky;                     // CB_SKY_*
  float sky_coeff[3], sky_scale;   // CB_SKY_CONST
  float p_sky;                     // lights_pdf_type: probability of connecting to the sky (list.c:44-49,76-88)
  float sky_far;                   // distance of the next-event point on the sky (shader.c:313-316)
  const EnvDev *env;               // CB_SKY_ENVMAP: in device memory (the out-of-line env_* functions take its address)
  uint32_t exterior_medium;        // 1 + index of the medium the camera sits in (shader_exterior_medium), 0 = vacuum
  int32_t has_media;               // any medium in the scene: the free-flight / transmittance code paths are live
  float *dbor;                     // `--dbor n` (view.c:291,339-350): n cascade buffers of fb_w*fb_h*3 floats, level major; NULL = off
  int32_t num_dbors;
  double *stat;                    // view->stat_enery / stat_cnt (view.c:470-471): [32 lanes][33 path lengths]{energy, count}; NULL = off
  // atomic-free accumulation (k_tile_accumulate): splats are appended to this list instead of being added to fb; a == NULL = off
  float4 *tile_a;                  // pixel_i, pixel_j, colour[0], colour[1]
  float *tile_b;                   // colour[2]
  uint32_t *tile_key, *tile_idx;   // 32x32 pixel tile of the sample / the record's own index (payload of the sort)
  unsigned int *tile_count;
  uint32_t tile_cap, tiles_x, tiles_y, tile_shift;
  float force_scramble;            // > 0: path->tangent_frame_scrambling of every new path (known-answer entries only)
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_add_f(float *p, float v) { atomicAdd(p, v); }
// the three channels of a pixel as TWO reductions: a two-float vector one (red.global.add.v2.f32, sm_90+) on the 8-byte aligned pair and
// a scalar one on the channel left over.  A pixel is 12 bytes, so the pair is (0,1) for even pixel indices and (1,2) for odd ones.
// Each float is still added on its own; only the number of requests to the L2 drops.
__device__ __forceinline__ void atomic_add_rgb(float *p, float a, float b, float c)
{
#if CB200_VEC_ATOMICS
  const size_t g = __cvta_generic_to_global(p);
  if((g & 7u) == 0u)
  {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(g), "f"(a), "f"(b) : "memory");
    atomicAdd(p + 2, c);
  }
  else
  {
    atomicAdd(p, a);
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(g + 4), "f"(b), "f"(c) : "memory");
  }
#else
  atomicAdd(p, a); atomicAdd(p + 1, b); atomicAdd(p + 2, c);
#endif
}

__device__ __forceinline__ float bh_w(float n)   // filter_bh_w, blackmanharris.h:28-41
{
  if(n > 3.0f || n < 0.0f) return 0.0f;
  const float N_1 = 1.0f/3.0f;
  const float cos1 = cosf(2.0f*PI_F*n*N_1), cos2 = cosf(4.0f*PI_F*n*N_1), cos3 = cosf(6.0f*PI_F*n*N_1);
  return 0.35875f - 0.48829f*cos1 + 0.14128f*cos2 - 0.01168f*cos3;
}

// view_splat (view.c:455-495)
__device__ bool splat(const RenderDev &R, float pixel_i, float pixel_j, float lambda, float value, int len)
{
  if(!(value > 0.0f)) return false;
  if(!(value < FLT_MAX)) return false;
  if(R.stat)
  { // view_deferred_splat's path statistics (view.c:470-471): energy and count per path length; one copy per lane keeps the
    // atomics of a warp on different addresses, the host sums the copies (cb200_render_path_stats)
    double *p = R.stat + (((threadIdx.x & 31u)*33u + (uint32_t)(len < 0 ? 0 : len > 32 ? 32 : len))*2u);
    atomicAdd(p, (double)value);
    atomicAdd(p + 1, 1.0);
  }
  float col[3];
  spectrum_to_camera(lambda, value, R.colour, col);
  const int wd = (int)R.fb_w, ht = (int)R.fb_h;
  if(R.tile_a && R.num_dbors <= 1)
  { // atomic-free mode: the sample is only recorded here; k_tile_accumulate filters it into its 32x32 tile in shared memory and the
    // tile is added to the framebuffer by its one owner.  Samples whose footprint misses the image are dropped like below.
    const int fx0 = (int)(pixel_i - 1.5f), fy0 = (int)(pixel_j - 1.5f);
    if(fx0 + 4 <= 0 || fy0 + 4 <= 0 || fx0 >= wd || fy0 >= ht) return false;
    const unsigned active = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(active) - 1;
    unsigned base = 0;
    if((int)lane == leader) base = atomicAdd(R.tile_count, (unsigned)__popc(active));
    base = __shfl_sync(active, base, leader);
    const unsigned o = base + __popc(active & ((1u << lane) - 1u));
    if(o < R.tile_cap)
    {
      int px = (int)floorf(pixel_i), py = (int)floorf(pixel_j);
      px = px < 0 ? 0 : px >= wd ? wd - 1 : px;
      py = py < 0 ? 0 : py >= ht ? ht - 1 : py;
      R.tile_a[o] = make_float4(pixel_i, pixel_j, col[0], col[1]);
      R.tile_b[o] = col[2];
      R.tile_key[o] = (uint32_t)(py >> R.tile_shift)*R.tiles_x + (uint32_t)(px >> R.tile_shift);
      R.tile_idx[o] = o;
    }
    return true;
  }
  const int x0 = (int)(pixel_i - 1.5f), y0 = (int)(pixel_j - 1.5f);
  const int u0 = -x0 < 0 ? 0 : -x0, v0 = -y0 < 0 ? 0 : -y0;
  const int u4 = x0 + 4 > wd ? wd - x0 : 4, v4 = y0 + 4 > ht ? ht - y0 : 4;
  float w[16];
  float weight = 0.0f;
  for(int v=v0;v<v4;v++) for(int u=u0;u<u4;u++)
  {
    const float uu = (x0 + u + .5f) - pixel_i, vv = (y0 + v + .5f) - pixel_j;
    const float f = bh_w(sqrtf(uu*uu + vv*vv) + 1.5f);
    w[4*v+u] = f;
    weight += f;
  }
  if(weight <= 0.0f) return false;
  weight = 1.0f/weight;
  // taps further than 1.5 pixels from the sample have weight exactly 0 (filter_bh_w returns 0 beyond its support): on average 9
  // of the 16; adding +-0 leaves a pixel unchanged, so they are skipped (the reference adds them, with the same result)
  for(int v=v0;v<v4;v++) for(int u=u0;u<u4;u++)
  {
    const float f = weight*w[4*v+u];
    if(f == 0.0f) continue;
    float *p = R.fb + 3*((size_t)(x0+u) + (size_t)wd*(y0+v));
    atomic_add_rgb(p, col[0]*f, col[1]*f, col[2]*f);
  }
  if(R.num_dbors > 1)
  { // density based outlier rejection cascade (view_splat_col, view.c:497-522): the sample goes to the two buffers whose
    // brightness range [2^l, 2^(l+1)) brackets its mean colour, split linearly in 1/value, through the same filter taps
    const float value = (col[0] + col[1] + col[2])/3.f;
    if(value > 0.0f)
    {
      const float lg = log2f(value);
      const float logval = 0.0f > lg ? 0.0f : lg;
      const int li = (int)logval;
      const int l = li < 0 ? 0 : li > R.num_dbors-1 ? R.num_dbors-1 : li;
      const int up = l + 1;
      const float lraw = value < 1.f ? 1.f : (((float)(1 << l)/value) - 0.5f)/0.5f;
      const float lv = lraw < 0.0f ? 0.0f : lraw > 1.0f ? 1.0f : lraw;
      const float uv = 1.f - lv;
      const bool last = l == R.num_dbors-1;
      const size_t level = (size_t)wd*ht*3;
      float coll[3], colu[3];
      for(int k=0;k<3;k++) { coll[k] = last ? col[k] : lv*col[k]; colu[k] = uv*col[k]; }
      for(int v=v0;v<v4;v++) for(int u=u0;u<u4;u++)
      {
        const float f = weight*w[4*v+u];
        if(f == 0.0f) continue;
        float *p = R.dbor + level*l + 3*((size_t)(x0+u) + (size_t)wd*(y0+v));
        atomic_add_rgb(p, coll[0]*f, coll[1]*f, coll[2]*f);
        if(up < R.num_dbors)
        {
          p += level;
          atomic_add_rgb(p, colu[0]*f, colu[1]*f, colu[2]*f);
        }
      }
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Atomic-free framebuffer accumulation (BASELINE north_star (3); the reference accumulates with a CAS loop per tap,
// filter/blackmanharris.h:43-77 + corona_common.h:316-329).  The recorded samples are sorted by 32x32 pixel tile; one block per
// tile filters its samples into a 36x36 shared-memory patch (the tile plus the two pixels the 4x4 Blackman-Harris footprint can
// reach beyond it on every side) with the same weights as splat(), then adds the patch to the framebuffer.  Patches of
// neighbouring tiles overlap in that border, so the tiles are processed in four checkerboard phases (even/odd tile column x
// even/odd tile row): within a phase no two blocks touch the same pixel, every pixel has exactly one writer and no global atomic
// is issued.
// ---------------------------------------------------------------------------------------------
// TILE: 32 for large frames; smaller frames take 16 or 8 so that a phase still has a block for every SM (R.tile_shift = log2 TILE)
template<int TILE>
__global__ void __launch_bounds__(256)
k_tile_accumulate(RenderDev R, const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ idx, int phase_x, int phase_y)
{
  constexpr int PATCH = TILE + 4;
  const uint32_t tx = 2u*blockIdx.x + (uint32_t)phase_x, ty = 2u*blockIdx.y + (uint32_t)phase_y;
  if(tx >= R.tiles_x || ty >= R.tiles_y) return;
  const uint32_t t = ty*R.tiles_x + tx;
  const uint32_t begin = __ldg(offsets + t), end = __ldg(offsets + t + 1);   // this tile's samples in the sorted list (k_tile_offsets)
  if(begin >= end) return;
  __shared__ float acc[PATCH*PATCH*3];
  for(int p=threadIdx.x;p<PATCH*PATCH*3;p+=blockDim.x) acc[p] = 0.0f;
  __syncthreads();
  const int wd = (int)R.fb_w, ht = (int)R.fb_h;
  const int ox = (int)tx*TILE - 2, oy = (int)ty*TILE - 2;     // image coordinates of acc[0]
  for(uint32_t k=begin+threadIdx.x; k<end; k+=blockDim.x)
  {
    const uint32_t i = __ldg(idx + k);
    const float4 a = __ldg(R.tile_a + i);
    const float c2 = __ldg(R.tile_b + i);
    const float pixel_i = a.x, pixel_j = a.y;
    // view_splat / filter_blackmanharris_splat, the arithmetic of splat() above
    const int x0 = (int)(pixel_i - 1.5f), y0 = (int)(pixel_j - 1.5f);
    const int u0 = -x0 < 0 ? 0 : -x0, v0 = -y0 < 0 ? 0 : -y0;
    const int u4 = x0 + 4 > wd ? wd - x0 : 4, v4 = y0 + 4 > ht ? ht - y0 : 4;
    float w[16];
    float weight = 0.0f;
    for(int v=v0;v<v4;v++) for(int u=u0;u<u4;u++)
    {
      const float uu = (x0 + u + .5f) - pixel_i, vv = (y0 + v + .5f) - pixel_j;
      const float f = bh_w(sqrtf(uu*uu + vv*vv) + 1.5f);
      w[4*v+u] = f;
      weight += f;
    }
    if(weight <= 0.0f) continue;
    weight = 1.0f/weight;
    for(int v=v0;v<v4;v++) for(int u=u0;u<u4;u++)
    {
      const float f = weight*w[4*v+u];
      if(f == 0.0f) continue;
      const int lx = x0 + u - ox, ly = y0 + v - oy;
      if(lx < 0 || ly < 0 || lx >= PATCH || ly >= PATCH) continue;   // cannot happen for a sample keyed to this tile
      float *p = acc + 3*(lx + PATCH*ly);
      atomicAdd(p+0, a.z*f); atomicAdd(p+1, a.w*f); atomicAdd(p+2, c2*f);   // shared memory
    }
  }
  __syncthreads();
  for(int p=threadIdx.x;p<PATCH*PATCH;p+=blockDim.x)
  {
    const int gx = ox + p % PATCH, gy = oy + p / PATCH;
    if(gx < 0 || gy < 0 || gx >= wd || gy >= ht) continue;
    const float r0 = acc[3*p], r1 = acc[3*p+1], r2 = acc[3*p+2];
    if(r0 == 0.0f && r1 == 0.0f && r2 == 0.0f) continue;
    float *q = R.fb + 3*((size_t)gx + (size_t)wd*gy);
    q[0] += r0; q[1] += r1; q[2] += r2;      // the only writer of this pixel in this phase
  }
}

// offsets[t] = first position of tile t in the sorted key list (the position where a key >= t first appears), for t in [0, tiles];
// offsets[tiles] = number of real records (the unused slots behind them carry the key `tiles`)
__global__ void k_tile_offsets(const uint32_t *__restrict__ key, uint32_t n, uint32_t tiles, uint32_t *__restrict__ offsets)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i > n) return;
  const uint32_t prev = i ? key[i-1] : 0u, cur = i < n ? key[i] : tiles;
  const uint32_t from = i ? prev + 1 : 0u, to = cur < tiles ? cur : tiles;
  for(uint32_t t=from; t<=to; t++) offsets[t] = i;
}

__global__ void k_fill_u32(uint32_t *p, uint32_t v, uint32_t n)
{
  for(uint32_t i=blockIdx.x*blockDim.x + threadIdx.x; i<n; i+=gridDim.x*blockDim.x) p[i] = v;
}

__device__ __forceinline__ void quat_mult(float *in, const float *p)   // quaternion.h:40-47 (w,x,y,z)
{
  const float rw = in[0], r0 = in[1], r1 = in[2], r2 = in[3];
  in[1] = rw*p[1] + r0*p[0] + r1*p[3] - r2*p[2];
  in[2] = rw*p[2] - r0*p[3] + r1*p[0] + r2*p[1];
  in[3] = rw*p[3] + r0*p[2] - r1*p[1] + r2*p[0];
  in[0] = rw*p[0] - r0*p[1] - r1*p[2] - r2*p[3];
}
__device__ __forceinline__ V3 quat_transform(const float *q, V3 p)
{
  float vq[4] = {0.0f, p.x, p.y, p.z};
  float inv[4] = {q[0], -q[1], -q[2], -q[3]};
  float res[4] = {q[0], q[1], q[2], q[3]};
  quat_mult(res, vq);
  quat_mult(res, inv);
  return mk3(res[1], res[2], res[3]);
}
__device__ void quat_slerp(const float *q, const float *p, float t, float *res)   // quaternion.h:78-103
{
  const float c = q[0]*p[0] + (q[1]*p[1] + q[2]*p[2] + q[3]*p[3]);
  if(fabsf(c) >= 1.0f) { for(int k=0;k<4;k++) res[k] = q[k]; return; }
  const float theta = acosf(c);
  const float s = sqrtf(1.0f - c*c);
  if(fabsf(s) < 1e-10f) { for(int k=0;k<4;k++) res[k] = (q[k] + p[k])*.5f; return; }
  const float a = sinf((1.0f - t)*theta)/s, b = sinf(t*theta)/s;
  for(int k=0;k<4;k++) res[k] = q[k]*a + p[k]*b;
}

// view_cam_init_frame (view.c:903-921)
__device__ void camera_frame(const CameraDev &C, float time, V3 &x, V3 &a, V3 &b, V3 &n)
{
  float q[4];
  quat_slerp(C.c.orient, C.c.orient_t1, time, q);
  a = normalise(quat_transform(q, mk3(1.0f, 0.0f, 0.0f)));
  b = normalise(quat_transform(q, mk3(0.0f, 1.0f, 0.0f)));
  n = normalise(quat_transform(q, mk3(0.0f, 0.0f, 1.0f)));
  x = mk3(C.c.pos[0]*(1.0f-time) + C.c.pos_t1[0]*time, C.c.pos[1]*(1.0f-time) + C.c.pos_t1[1]*time, C.c.pos[2]*(1.0f-time) + C.c.pos_t1[2]*time);
}

__device__ __forceinline__ void write_ray(cb_ray_t *rays, uint64_t i, V3 pos, V3 dir, float time, uint32_t ign_lo, uint32_t ign_hi)
{
  float2 *p = reinterpret_cast<float2 *>(rays + i);
  p[0] = make_float2(pos.x, pos.y); p[1] = make_float2(pos.z, dir.x); p[2] = make_float2(dir.y, dir.z);
  p[3] = make_float2(time, 0.0f);
  p[4] = make_float2(__uint_as_float(ign_lo), __uint_as_float(ign_hi));
}

// path_extend at length == 0: sensor vertex + first edge
__device__ void path_start(const RenderDev &R, uint64_t index, PathState &s, V3 &ray_pos)
{
  const PointsDev &P = R.points;
  const CameraDev &C = R.cam;
  s.index_lo = (uint32_t)index; s.index_hi = (uint32_t)(index >> 32);
  uint32_t mt = 0;
  s.scramble = 0.1f + point_mt(P, index, mt++)*(0.9f - 0.1f);
  if(R.force_scramble > 0.0f) s.scramble = R.force_scramble;
  s.lambda = 360.0f + (830.0f - 360.0f)*fmodf(point_dim(P, index, 2) + 0.0f, 1.0f);   // spectrum_sample_lambda
  const float exposure_time = C.exposure_time;
  s.time = point_dim(P, index, 3)*fminf(1.0f, exposure_time/(1.0f/30.0f));             // view_sample_time
  (void)point_dim(P, index, 6);                                                          // camid: one camera
  const float i = point_dim(P, index, 0)*C.width, j = point_dim(P, index, 1)*C.height;
  const float r1 = point_dim(P, index, 4), r2 = point_dim(P, index, 5);
  const float lens_radius = (.5f/C.fstop)*C.c.focal_length;
  const float ang = (float)(2.0*PI_D*(double)r1);
  const float u = cosf(ang)*sqrtf(r2)*lens_radius, v = sinf(ang)*sqrtf(r2)*lens_radius;
  V3 x, a, b, n;
  camera_frame(C, s.time, x, a, b, n);
  const float f = C.c.focus/C.c.focal_length;
  const float f_dir = C.c.focus;
  const float f_rg = -C.c.film_width*f/C.width, f_up = -C.c.film_height*f/C.height;
  const V3 aoff = mk3(u*a.x + v*b.x, u*a.y + v*b.y, u*a.z + v*b.z);
  const float ci = (i - .5f*C.width)*f_rg, cj = (j - .5f*C.height)*f_up;
  V3 om = mk3(f_dir*n.x + (ci*a.x + cj*b.x) - aoff.x, f_dir*n.y + (ci*a.y + cj*b.y) - aoff.y, f_dir*n.z + (ci*a.z + cj*b.z) - aoff.z);
  om = normalise(om);
  const float fs = C.fstop;
  const float A = PI_F*C.c.focal_length*C.c.focal_length/(4.0f*fs*fs);
  const float pdf_a = (float)(1.0/(double)A);
  const float sensor = 106.86535f*100.0f*exposure_time;
  const float dt = dot(om, n);
  const float dot4 = dt*dt*dt*dt;
  s.pixel_i = fminf(fmaxf(i, 0.0f), C.width - 1e-4f);
  s.pixel_j = fminf(fmaxf(j, 0.0f), C.height - 1e-4f);
  const float G = dot4/(C.c.focal_length*C.c.focal_length);
  const float pdf_v = 1.0f/(C.c.film_width*C.c.film_height);
  s.pdf_proj = pdf_v*pdf_a/G;
  const float thr = sensor*G/(pdf_a*pdf_v);
  s.x[0] = x.x + aoff.x; s.x[1] = x.y + aoff.y; s.x[2] = x.z + aoff.z;
  s.omega[0] = om.x; s.omega[1] = om.y; s.omega[2] = om.z;
  s.thr = thr; s.thr_prev = thr;
  s.cos_prev = fabsf(dot(n, om));     // path_lambert at the sensor vertex
  s.cur_ior = 1.0f;
  s.prim_lo = s.prim_hi = 0xffffffffu;
  { const PathPdf one = pp_one(); s.lre = lre_pack(1, 7, one.e); s.pdf_m = one.m; }   // v[0] used 7 dims; v[1] uses one (free path)
                                                                                       // (thinlens.c:101-104)
  s.bits = (mt << 8);
  s.med_n = 0;
  for(int k=0;k<MED_MAX;k++) { s.med_shape[k] = 0; s.med_ior[k] = 1.0f; }
  ray_pos = mk3(s.x[0], s.x[1], s.x[2]);   // sensor vertex has no primitive: no offset (pathspace.c:759-761)
}

// Wave ordering.  The reference draws the pixel of path i from its random dimensions 0/1 over the whole image
// (thinlens.c:117-118), so consecutive path indices start at unrelated pixels and a warp's 32 camera rays share nothing.
// The samples are a pure function of the path index, so the wave is free to process its indices in any order: sort them
// by the Morton code of their pixel (24 bits) and neighbouring lanes trace neighbouring pixels.  Same samples, same image.
__device__ __forceinline__ uint32_t part1by1(uint32_t x)
{
  x &= 0x0000ffffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__global__ void __launch_bounds__(RB)
k_pixel_keys(RenderDev R, uint64_t first_index, uint32_t n, uint32_t *keys, uint32_t *vals)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  const uint64_t index = first_index + i;
  const float pi = point_dim(R.points, index, 0)*R.cam.width, pj = point_dim(R.points, index, 1)*R.cam.height;
  const uint32_t x = (uint32_t)fminf(fmaxf(pi, 0.0f), R.cam.width - 1.0f), y = (uint32_t)fminf(fmaxf(pj, 0.0f), R.cam.height - 1.0f);
  keys[i] = part1by1(x) | (part1by1(y) << 1);
  vals[i] = i;
}

__global__ void __launch_bounds__(RB)
k_path_start(RenderDev R, uint64_t first_index, uint32_t n, const uint32_t *__restrict__ order, PathState *st, cb_ray_t *rays, float *aux,
             float *maxd)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  PathState s;
  V3 pos;
  const uint64_t index = first_index + (order ? order[i] : i);
  path_start(R, index, s, pos);   // st / rays already point at the first free slot of the pool
  if(maxd)
  { // the camera sits in the exterior medium: sample the free flight of the first edge before tracing it (pathspace.c:716-747)
    float clip = FLT_MAX;
    const Vol vol = medium_eval(R.mats, R.exterior_medium, s.lambda);
    if(vol.present && vol.mu_s > 0.0f) clip = vol_free_flight(vol, point_dim(R.points, index, lre_rand_beg(s.lre) + 0));
    maxd[i] = clip;
  }
  write_ray(rays, i, pos, mk3(s.omega[0], s.omega[1], s.omega[2]), s.time, 0xffffffffu, 0xffffffffu);
  if(st) st[i] = s;
  if(aux) { aux[4*i+0] = s.pixel_i; aux[4*i+1] = s.pixel_j; aux[4*i+2] = s.lambda; aux[4*i+3] = s.thr; }
}

__device__ __forceinline__ uint32_t sample_cdf_dev(const float *cdf, uint32_t num, float rand)   // sampler_common.h:206-226
{
  uint32_t mn = 0, mx = num, t = mx/2;
  while(t != mn)
  {
    if(cdf[t] <= rand) mn = t; else mx = t;
    t = (mn + mx)/2;
  }
  if(mx < num && cdf[t] <= rand) t = mx;
  return t;
}

// prims_sample + prims_retime (prims.c:177-252) for triangles, quads and spheres (line primitives have a zero or negative
// area upstream, line.h:55-67, and are never picked)
__device__ void light_point(const SceneGeo &S, uint64_t pid, float r0, float r1, float time, Vtx &h)
{
  const uint32_t vcnt = (uint32_t)(pid >> 61) & 7u;
  h.prim_lo = (uint32_t)pid; h.prim_hi = (uint32_t)(pid >> 32);
  if(vcnt == CB_PRIM_SPHERE)
  { // prims.c:225-230 + geo_sphere_retime (sphere.h:38-49) + sample_sphere (sampler_common.h:136-143)
    h.u = r0;
    h.v = (float)((double)acosf(r1)/PI_D);
    const float x1 = -(cosf((float)((double)h.v*PI_D)) - 1.f)/2.f;
    const float z = 1.f - 2.f*x1;
    const float rr = sqrtf(1.f - z*z);
    const float phi = (float)((double)2.f*PI_D*(double)h.u);
    const float radius = __uint_as_float(geo_vtx(S, pid, 0, 0)->n);
    const V3 c = geo_vertex_time(S, pid, 0, time);
    h.x = mk3(c.x + radius*(rr*cosf(phi)), c.y + radius*(rr*sinf(phi)), c.z + radius*z);
  }
  else if(vcnt == CB_PRIM_QUAD)
  {
    h.u = r0; h.v = r1;
    const V3 v0 = geo_vertex_time(S, pid, 0, time), v2 = geo_vertex_time(S, pid, 2, time);
    V3 p1, p2; float u, v;
    if(h.v >= h.u) { p1 = geo_vertex_time(S, pid, 1, time); p2 = v2; u = h.u; v = h.v - h.u; }
    else           { p1 = v2; p2 = geo_vertex_time(S, pid, 3, time); u = h.u - h.v; v = h.v; }
    const float w = 1.0f - u - v;
    h.x = mk3(w*v0.x + v*p1.x + u*p2.x, w*v0.y + v*p1.y + u*p2.y, w*v0.z + v*p1.z + u*p2.z);
  }
  else
  {
    float a = sqrtf(r0);
    const float b = (1.0f - r1)*a, c = r1*a;
    a = 1.0f - a;
    h.u = c; h.v = b;
    const V3 v0 = geo_vertex_time(S, pid, 0, time), v1 = geo_vertex_time(S, pid, 1, time), v2 = geo_vertex_time(S, pid, 2, time);
    const float w = 1.0f - h.u - h.v;
    h.x = mk3(w*v0.x + h.v*v1.x + h.u*v2.x, w*v0.y + h.v*v1.y + h.u*v2.y, w*v0.z + h.v*v1.z + h.u*v2.z);
  }
}

// accel_intersect's test of ONE primitive addressed through the geometry store (prims_intersect, prims.c:638-672): the same
// vertices and the same arithmetic as the leaf records, used to find where a shadow ray first crosses its own light primitive
__device__ void prim_test_geo(const SceneGeo &S, uint64_t pid, const RayD &r, HitD &h)
{
  const uint32_t vcnt = (uint32_t)(pid >> 61) & 7u;
  if(vcnt == CB_PRIM_SPHERE)
  { // geo_sphere_intersect (sphere.h:112-166): only the distance matters here
    const float t = sphere_t(geo_vertex_time(S, pid, 0, r.time), __uint_as_float(geo_vtx(S, pid, 0, 0)->n), r);
    if(t > r.min_dist && t < h.dist) h.dist = t;
    return;
  }
  if(vcnt != CB_PRIM_TRI && vcnt != CB_PRIM_QUAD) return;
  const uint32_t id_lo = (uint32_t)pid, id_hi = (uint32_t)(pid >> 32);
  const V3 v0 = geo_vertex_time(S, pid, 0, r.time), v1 = geo_vertex_time(S, pid, 1, r.time), v2 = geo_vertex_time(S, pid, 2, r.time);
  if(tri_intersect(v0, v1, v2, id_lo, id_hi, r, h) || vcnt == CB_PRIM_TRI) return;
  tri_intersect(v0, v2, geo_vertex_time(S, pid, 3, r.time), id_lo, id_hi, r, h);
}

__device__ __forceinline__ float max3abs(V3 x) { return fmaxf(fmaxf(.5f, fabsf(x.x)), fmaxf(fabsf(x.y), fabsf(x.z))); }
// prims_offset_ray (src/prims.c:374-388): the origin of a ray leaving the surface point x in direction dir
__device__ __forceinline__ V3 offset_origin(V3 x, V3 dir)
{
  const float eps = max3abs(x)*1e-4f;
  return mk3(x.x + eps*dir.x, x.y + eps*dir.y, x.z + eps*dir.z);
}

// next, nee, hits[kind], em are per wave (the first 8 words are cleared before every wave); splats keeps counting; tile_count: records
// pending for k_tile_accumulate
// next: low word = surviving paths, high word = next-event records of the wave (reserved with one atomic); nee: unused
struct ShadeCounters { unsigned long long next, nee, hits[5], em, splats; unsigned int tile_count, pad_; };

// path_G for the edge between a surface vertex and the sampled light point (pathspace.c:58-69)
__device__ __forceinline__ float cos_lambert(const Vtx &v, const Vtx &l, V3 d, float dist)
{
  return fabsf(dot(v.n, d))*fabsf(dot(l.n, d))/(dist*dist);
}

// ---- skies: the built-in `cloudy' (src/shader.c:268-334: L = 500 (1 + omega_z)/2, sampled with pdf (1 + z)/2 / (2 pi)) and the
//      constant-colour module (src/shaders/sky_const.c: L = scale * rgb2spec(lambda), uniform sphere) -----------------------------
// ---- sky_envmap.c: latitude-longitude map of spectral coefficients with a mip hierarchy for importance sampling ------------
__device__ __forceinline__ float env_sh(const float4 c)   // sky_envmap_sh (:40-45): the spectrum at 660, 560, 480, 400 nm
{
  const float cf[3] = {c.x, c.y, c.z};
  return ((rgb2spec_eval(cf, 660.0f)*c.w + rgb2spec_eval(cf, 560.0f)*c.w) + rgb2spec_eval(cf, 480.0f)*c.w) + rgb2spec_eval(cf, 400.0f)*c.w;
}
__device__ __forceinline__ float4 env_fetch(const EnvDev &E, int i, int j)   // fb_fetchi (framebuffer.h:211-215)
{
  if(i < 0 || i >= E.width || j < 0 || j >= E.height) return E.px[0];
  return E.px[(size_t)E.width*j + i];
}
__device__ __forceinline__ V3 env_mulv(const float *a, V3 v)   // mat3_mulv (matrix3.inc:41-50)
{
  V3 r;
  r.x = ((0.0f + a[0]*v.x) + a[1]*v.y) + a[2]*v.z;
  r.y = ((0.0f + a[3]*v.x) + a[4]*v.y) + a[5]*v.z;
  r.z = ((0.0f + a[6]*v.x) + a[7]*v.y) + a[8]*v.z;
  return r;
}
__device__ __noinline__ float env_eval(const EnvDev &E, V3 omega, float lambda)   // eval (:59-94), sensor paths
{
  const V3 dir = env_mulv(E.world_inv, omega);
  float x, y;
  if(fabsf(dir.z) > 1.0f) { x = 0.0f; y = 0.0f; }
  else
  {
    y = (float)((double)acosf(dir.z)/PI_D*(double)E.height);
    x = (float)(((PI_D + (double)atan2f(dir.x, dir.y))/(2.0*PI_D))*(double)E.width);
  }
  if(!(x < E.width && x >= 0.0f && y < E.height && y >= 0.0f)) x = y = 0.0f;
  const float4 c = env_fetch(E, (int)x, (int)y);
  const float cf[3] = {c.x, c.y, c.z};
  return rgb2spec_eval(cf, lambda)*(E.mul*c.w);
}
__device__ __noinline__ float env_pdf(const EnvDev &E, V3 omega)   // pdf (:194-219), solid angle
{
  const V3 dir = env_mulv(E.world_inv, omega);
  const double yd = (double)acosf(dir.z)/PI_D*(double)E.height, xd = ((PI_D + (double)atan2f(dir.x, dir.y))/(2.0*PI_D))*(double)E.width;
  const float y = (float)fmin(fmax(yd, 0.0), (double)(E.height - 1)), x = (float)fmin(fmax(xd, 0.0), (double)(E.width - 1));
  const float sin_theta = sqrtf(fmaxf(1e-12f, 1.0f - dir.z*dir.z));
  const int i = (int)x, j = (int)y;
  const float qs = sinf((float)(PI_D*(double)(.5f + j)/(double)E.height));
  return (float)((double)(env_sh(env_fetch(E, i, j))*E.mul*qs)/(E.sum*(double)sin_theta*(double)2.0f*PI_D*PI_D));
}
__device__ __noinline__ V3 env_sample(const EnvDev &E, float x1, float x2, float lambda, float &edf, float &pdf)   // sample (:96-192), next event
{
  double wx = x1, wy = x2;
  const float *top = E.mip + E.mip_off[E.levels-1];
  if(wx < top[0]) wx *= .5/(double)top[0];
  else wx = 1.0 - (1.0 - wx)*.5/(1.0 - (double)top[0]);
  wx *= 2.0;
  for(int m=E.levels-2;m>=0;m--)
  {
    const float *mp = E.mip + E.mip_off[m];
    int i = (int)wx, j = (int)wy;
    const int wd = E.w2n >> m;
    if(2*i >= wd) i = (wd-1)/2;
    if(4*j >= wd) j = (wd-1)/4;
    const float up = mp[2*i + 2*j*wd] + mp[2*i+1 + 2*j*wd];
    if(wy < (double)((float)j + up))
    {
      wy = j + (wy - j)*.5/(double)up;
      j = 2*j;
    }
    else if((double)up < 0.999)
    {
      wy = j + 1.0 - (1.0 - (wy - j))*.5/(1.0 - (double)up);
      j = 2*j + 1;
    }
    const double f = (double)(mp[2*i + j*wd]/(mp[2*i + j*wd] + mp[2*i+1 + j*wd]));
    if(wx <= i + f) wx = i + (wx - i)*.5/f;
    else if(wx > i + f) wx = i + 1.0 - (1.0 - wx + i)*.5/(1.0 - f);
    wx *= 2.0; wy *= 2.0;
  }
  const float x = (float)(wx*(double)E.aspectx), y = (float)(wy*(double)E.aspecty);
  const int i = (int)x, j = (int)y;
  const float4 c = env_fetch(E, i, j);
  const float theta = (float)(PI_D*(double)(y/E.height));
  const float phi = (float)(2.0*PI_D*(double)(x/E.width) - PI_D);
  float sin_theta, cos_theta, sin_phi, cos_phi;
  sincosf(theta, &sin_theta, &cos_theta);
  const float qs = sinf((float)(PI_D*(double)(.5f + j)/(double)E.height));
  sincosf(phi, &sin_phi, &cos_phi);
  pdf = (float)((double)(env_sh(c)*E.mul*qs)/(PI_D*PI_D*2.0*E.sum*(double)sin_theta));
  const float cf[3] = {c.x, c.y, c.z};
  const float em = rgb2spec_eval(cf, lambda)*(E.mul*c.w);
  edf = em/pdf;
  return env_mulv(E.world, mk3(sin_phi*sin_theta, cos_phi*sin_theta, cos_theta));
}

__device__ __forceinline__ float sky_eval(const RenderDev &R, V3 omega, float lambda)
{
  if(R.sky == CB_SKY_ENVMAP) return env_eval(*R.env, omega, lambda);
  if(R.sky == CB_SKY_CONST) return rgb2spec_eval(R.sky_coeff, lambda)*R.sky_scale;                      // sky_const.c:41-45
  return (float)((double)(1.0f*500.0f*0.5f)*(1.0 + (double)omega.z));                                    // sky_cloudy, v != 0 (shader.c:276-279)
}
__device__ __forceinline__ float sky_pdf(const RenderDev &R, V3 omega)    // solid angle
{
  if(R.sky == CB_SKY_ENVMAP) return env_pdf(*R.env, omega);
  if(R.sky == CB_SKY_CONST) return (float)(1.0/((double)4.0f*PI_D));                                     // sky_const.c:84-87
  return (float)((double)(0.5f + omega.z*.5f)/(2.0*PI_D));                                               // shader.c:328-331
}
// next-event sample: direction, emission / pdf, pdf
__device__ __forceinline__ V3 sky_sample(const RenderDev &R, float x1, float x2, float lambda, float &edf, float &pdf)
{
  if(R.sky == CB_SKY_ENVMAP) return env_sample(*R.env, x1, x2, lambda, edf, pdf);
  if(R.sky == CB_SKY_CONST)
  { // sample_sphere (sampler_common.h:136-143)
    const float z = 1.f - 2.f*x1;
    const float r = sqrtf(1.f - z*z);
    const float phi = (float)((double)2.f*PI_D*(double)x2);
    pdf = (float)((double)1.0f/((double)4.0f*PI_D));
    edf = (rgb2spec_eval(R.sky_coeff, lambda)*R.sky_scale)/pdf;
    return mk3(r*cosf(phi), r*sinf(phi), z);
  }
  const float z = -(1.0f - 2.0f*sqrtf(1.0f - x1));                                                        // shader.c:297-305
  const float sin_theta = (float)sqrt(1.0 - (double)(z*z));
  const float ang = (float)((double)2.f*PI_D*(double)x2);
  const float em = ((.5f + z*.5f)*1.0f)*500.0f;
  pdf = (float)((double)(.5f + z*.5f)/((double)2.0f*PI_D));
  edf = em/pdf;
  return mk3(sin_theta*cosf(ang), sin_theta*sinf(ang), z);
}

// A path whose ray left the scene under a non-black sky gets an environment vertex: emission with the sampler's weight, then
// the path ends (path_propagate pathspace.c:856-873, path_extend / nee_sample refuse to continue :196 / nee.h:91).
__global__ void __launch_bounds__(RB)
k_sky_miss(RenderDev R, uint32_t n, const PathState *__restrict__ st, const cb_hitrec_t *__restrict__ hits, ShadeCounters *cnt,
           NeeRec *__restrict__ em_recs)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  NeeRec e;
  e.value = 0.0f;
  if(i < n)
  {
    const uint2 p = *reinterpret_cast<const uint2 *>(hits + i);
    if((p.x & p.y) == 0xffffffffu && !(R.has_media && hits[i].dist < FLT_MAX))   // (a clipped miss is a volume vertex)
    {
      const PathState s = st[i];
      const V3 omega = mk3(s.omega[0], s.omega[1], s.omega[2]);
      float em = sky_eval(R, omega, s.lambda);
      if(R.has_media)
      { // an edge of infinite length through a medium has transmittance exp(-FLT_MAX mu_t) = 0 (pathspace.c:839-843, shader.c:57-64)
        Media med; media_unpack(med, s.med_n);
        for(int k=0;k<MED_MAX;k++) med.shape[k] = s.med_shape[k];
        const Vol vol = medium_eval(R.mats, media_medium(med, R.exterior_medium), s.lambda);
        if(vol.present && !(vol.mu_t == 0.0f)) em = 0.0f;
      }
      if(em > 0.0f)
      {
        const float pdf_v = s.pdf_proj*s.cos_prev;      // path_G towards the environment = lambert at the previous vertex (pathspace.c:60-61)
        PathPdf pp; pp.m = s.pdf_m; pp.e = lre_exp(s.lre);
        float w = 1.0f;
        float pdf_nee = 0.0f;
        if(R.sampler == CB_SAMPLER_PTDL)
        {
          if(lre_length(s.lre) + 1 >= 3 && (s.bits & 1u) && R.p_sky > 0.0f) pdf_nee = R.p_sky*sky_pdf(R, omega);   // nee_pdf_nee, nee.h:21-38
          w = pdf_v/(pdf_nee + pdf_v);
        }
        if(!pp_contributes(pp, pdf_v, pdf_nee)) w = 0.0f;   // sampler_mis in float range only (PathPdf)
        if(R.sampler == CB_SAMPLER_PTNEE) w = lre_length(s.lre) + 1 == 2 ? 1.0f : 0.0f;   // ptnee.c:53-58: only the directly visible sky
        // queued like the emission k_shade finds on surfaces: k_nee_resolve is the one kernel that splats
        e.value = (s.thr*em)*w;   // lights_eval_vertex: isotropic for the envmap (list.c:272-273)
        e.lambda = s.lambda; e.pixel_i = s.pixel_i; e.pixel_j = s.pixel_j; e.total_dist = 0.0f;
        e.light_lo = e.light_hi = 0xffffffffu; e.len = (uint32_t)(lre_length(s.lre) + 1);
      }
    }
  }
  const bool have = e.value > 0.0f && e.value < FLT_MAX;      // view_splat's own acceptance test (view.c:457-459)
  const uint32_t m = __ballot_sync(0xffffffffu, have);
  if(m)
  {
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long base = 0;
    if(lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(&cnt->em, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if(have) em_recs[base + __popc(m & ((1u << lane) - 1u))] = e;
  }
}

// Paths whose ray escaped into the (black) sky have nothing left to do (pathspace.c:856-873): the shading kernel only runs
// on the slots that hit something.  Their indices are compacted first -- into one list per BSDF kind of the surface that was
// hit (diffuse / dielectric / metal), so that every shading launch has full warps AND a single material class: the
// "material-sorted" part of the wavefront.  list[kind*n_cap + k] = slot.
__global__ void __launch_bounds__(256)
k_compact_hits(RenderDev R, const cb_hitrec_t *__restrict__ hits, uint32_t n, uint32_t n_cap, uint32_t *__restrict__ list, ShadeCounters *cnt,
               int single_kind)
{
  __shared__ uint32_t warp_count[5][8];
  __shared__ uint32_t block_base[5];
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  int kind = -1;
  if(i < n)
  {
    const uint2 p = *reinterpret_cast<const uint2 *>(hits + i);   // prim id words
    if((p.x & p.y) != 0xffffffffu)
      kind = single_kind >= 0 ? single_kind : R.mats.mat[R.geo.shape_material[p.x >> 3]].bsdf;
    else if(R.has_media && hits[i].dist < FLT_MAX) kind = 4;   // the sampled free flight ended before any surface: volume vertex
  }
  uint32_t m[5];
#pragma unroll
  for(int k=0;k<5;k++) { m[k] = __ballot_sync(0xffffffffu, kind == k); if(lane == 0) warp_count[k][w] = __popc(m[k]); }
  __syncthreads();
  if(threadIdx.x < 5)
  {
    const int k = threadIdx.x;
    uint32_t tot = 0;
    for(int q=0;q<8;q++) { const uint32_t c = warp_count[k][q]; warp_count[k][q] = tot; tot += c; }
    block_base[k] = tot ? (uint32_t)atomicAdd(&cnt->hits[k], (unsigned long long)tot) : 0u;
  }
  __syncthreads();
#pragma unroll
  for(int k=0;k<5;k++)
    if(kind == k) list[(size_t)k*n_cap + block_base[k] + warp_count[k][w] + __popc(m[k] & ((1u << lane) - 1u))] = i;
}

// The vertex-level callbacks of a volume vertex are the medium's (medium_rgb.c:62-103): Henyey-Greenstein around the incoming
// direction, scaled by mu_s.  KINDS == 16 selects them, every other KINDS goes to the surface BSDFs of shading.cuh.
template<int KINDS>
__device__ __forceinline__ float vtx_eval(const MaterialsDev &M, Vtx &v, V3 wi, V3 wo, float lambda, float cur_ior)
{
  if constexpr(KINDS == 16) { v.mode |= M_GLOSSY | M_VOLUME; return v.vol_mu_s*hg_eval(v.vol_g, wi, wo); }
  else return bsdf_eval<KINDS>(M, v, wi, wo, lambda, cur_ior);
}
template<int KINDS>
__device__ __forceinline__ float vtx_pdf(const MaterialsDev &M, const Vtx &v, V3 wi, V3 wo)
{
  if constexpr(KINDS == 16) return (v.mode & M_VOLUME) ? hg_eval(v.vol_g, wi, wo) : 0.0f;
  else return bsdf_pdf<KINDS>(M, v, wi, wo);
}

// one vertex of every live path whose surface has BSDF kind KIND (0 diffuse, 1 dielectric, 2 metal, 3 diffdiel; 4 = volume
// vertex of a homogeneous medium): each launch carries the code of ONE BSDF (the all-in-one kernel was 17 k instructions and lost 24 % of
// its stall samples to instruction-cache misses) and no lane waits for another material's branch.  KINDS = 1 << KIND for the
// templates of shading.cuh; light vertices reached by next-event estimation only need their emission slots, which every
// variant fills.  MEDIA compiles the participating-media terms in (free-flight sampling of the next edge, transmittance and
// distance pdf of the finished one: pathspace.c:716-747,822-843, shader.c:46-155); scenes without media run the MEDIA = false
// variants, which are instruction for instruction the surface-only integrator.
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 1
#endif
template<int KINDS, bool MEDIA>
__global__ void __launch_bounds__(RB, SHADE_MIN_BLOCKS)
k_shade(RenderDev R, uint32_t n, const PathState *__restrict__ st_in, const cb_ray_t *__restrict__ rays_in,
        const cb_hitrec_t *__restrict__ hits, PathState *__restrict__ st_out, cb_ray_t *__restrict__ rays_out,
        cb_ray_t *__restrict__ nee_rays, float *__restrict__ nee_maxdist, uint2 *__restrict__ nee_light, NeeRec *__restrict__ nee_recs,
        ShadeCounters *cnt, const uint32_t *__restrict__ hit_list, int kind, float *__restrict__ maxd_out, NeeRec *__restrict__ em_recs)
{
  constexpr bool VOLV = KINDS == 16;   // this launch shades volume vertices
  const uint32_t n_hits = (uint32_t)*reinterpret_cast<volatile unsigned long long *>(&cnt->hits[kind]);   // written by k_compact_hits
  (void)n;
  // (a resident-wave grid striding over the list, with the next slot's state prefetched, was measured: 3.1 -> 3.2 ms per 4K step
  // with or without the prefetch, profiles/r2n -- the block scheduler's own refill does better)
  const uint32_t t = blockIdx.x*blockDim.x + threadIdx.x;
  if(blockIdx.x*blockDim.x >= n_hits) return;   // whole block beyond the list (the grid is sized for n, the upper bound)
  bool alive = false, have_nee = false;
  // emission found by extension is not splatted here: the contribution is queued (em_recs) and filtered into the framebuffer by
  // k_nee_resolve together with the wave's next-event contributions -- few paths end on an emitter, and the 4x4 Blackman-Harris
  // footprint inlined twice into this kernel cost every path registers and instruction-cache misses (10 % of the kernel's time)
  float em_value = 0.0f; int em_len = 0;
  PathState s;
  V3 next_pos = mk3(0, 0, 0), next_dir = mk3(0, 0, 0);
  float next_clip = FLT_MAX;
  cb_ray_t nray; NeeRec nrec; float nmax = 0.0f;
  if(t < n_hits)
  {
    const uint32_t i = hit_list[t];
    s = st_in[i];
    const cb_hitrec_t h = hits[i];
    const uint64_t index = (uint64_t)s.index_lo | ((uint64_t)s.index_hi << 32);
    const V3 omega = mk3(s.omega[0], s.omega[1], s.omega[2]);
    const bool hit_geo = !((h.prim[0] & h.prim[1]) == 0xffffffffu);
    // black sky: an environment vertex carries no emission and ends the path (pathspace.c:856-873, shader.c:464-470)
    if(hit_geo || VOLV)
    {
      const cb_ray_t ray = rays_in[i];
      Vtx v;
      v.prim_lo = h.prim[0]; v.prim_hi = h.prim[1];
      v.u = h.u; v.v = h.v; v.s = v.t = 0.0f;
      v.flags = 0; v.mode = M_ABSORB; v.material_modes = 0;
      v.x = mk3(ray.pos[0] + h.dist*ray.dir[0], ray.pos[1] + h.dist*ray.dir[1], ray.pos[2] + h.dist*ray.dir[2]);
      Media med; media_unpack(med, s.med_n);
      for(int k=0;k<MED_MAX;k++) { med.shape[k] = s.med_shape[k]; med.ior[k] = s.med_ior[k]; }
      Vol vol; vol.present = false; vol.mu_t = vol.mu_s = vol.g = 0.0f;   // e[v].vol: the medium of the edge that just ended
      if(MEDIA) vol = medium_eval(R.mats, media_medium(med, R.exterior_medium), s.lambda);
      if constexpr(VOLV)
      { // manifold_init + shader_prepare of a vertex without primitive (manifold.h:236-240, shader.c:478-503)
        v.n = v.gn = omega;
        scrambled_onb(s.scramble, v.n, v.a, v.b);
        v.material_modes = M_VOLUME | M_GLOSSY;
        v.rd = v.rs = v.rg = v.em = 0.0f; v.roughness = 1.0f; v.ior = 1.0f; v.eta = 1.0f; v.mat = -1;
        v.vol_mu_s = vol.mu_s; v.vol_g = vol.g;
      }
      else prepare_vertex<KINDS>(R.geo, R.mats, v, omega, s.time, s.lambda, s.scramble, med, s.cur_ior);
      // self intersection (pathspace.c:809-820)
      const uint32_t vcnt = v.prim_hi >> 29;
      const bool self = !VOLV && (vcnt > 2 || h.dist < 1e-4f) && v.prim_lo == s.prim_lo && v.prim_hi == s.prim_hi;
      if(!self)
      {
        if(!VOLV && v.em > 0.0f && !(v.flags & F_INSIDE)) v.material_modes = v.mode = M_EMIT;
        // on-surface pdf of this vertex (pathspace.c:252-253, 45-69); no cosine at a volume vertex
        const float cos_v = VOLV ? 1.0f : fabsf(dot(v.n, omega));
        const float G = s.cos_prev*cos_v/(h.dist*h.dist);
        float pdf_v, thr;
        if(MEDIA)
        { // distance sampling: pdf and transmittance of the edge (shader_vol_sample / shader_vol_transmittance, pathspace.c:822-843)
          float e_pdf = 1.0f, e_T = 1.0f;
          if(vol.present)
          {
            e_T = expf(-h.dist*vol.mu_t);
            if(vol.mu_s > 0.0f) e_pdf = VOLV ? e_T*vol.mu_t : e_T;   // scattering medium: free flight sampled first
          }
          pdf_v = (s.pdf_proj*e_pdf)*G;          // pathspace.c:889-890, then :252
          thr = s.thr*(e_T/e_pdf);               // path_update_throughput (pathspace.c:158)
        }
        else { pdf_v = s.pdf_proj*G; thr = s.thr; }
        const int vi = lre_length(s.lre);        // index of this vertex
        const int len = vi + 1;                  // path->length from here on
        const int rand_beg_v = lre_rand_beg(s.lre);
        PathPdf pp; pp.m = s.pdf_m; pp.e = lre_exp(s.lre);   // product of v[1..vi-1].pdf
        float pdf_v_final = pdf_v;                           // v[vi].pdf as later vertices will see it
        bool stop = false;
        // ---- emission: path_update_throughput + sampler splat
        if((v.mode & M_EMIT) && R.sampler == CB_SAMPLER_PTNEE)
        { // ptnee.c:53-58: emission found by extension only counts where next-event estimation cannot create the path, on
          // directly visible emitters, unweighted.  (Upstream tests `v[2].mode` of the two-vertex path there -- the slot
          // BEHIND its last vertex, i.e. whatever the worker thread's previous path left in it -- so its own result on those
          // pixels depends on thread scheduling; the vertex the comment in the source means, v[1], is used here.)
          if(len == 2) { em_value = thr*light_eval(v, omega); em_len = len; }
        }
        else if(v.mode & M_EMIT)
        {
          const float L = thr*light_eval(v, omega);
          float w = 1.0f;
          float pdf_nee = 0.0f;
          if(R.sampler == CB_SAMPLER_PTDL)
          {
            if(len >= 3 && (s.bits & 1u) && R.lights.p_geo > 0.0f)
              pdf_nee = R.lights.p_geo*R.lights.shape_pdf[v.prim_lo >> 3];
            w = pdf_v/(pdf_nee + pdf_v);        // sampler_mis, ptdl.c:78-88 with one wavelength
          }
          if(!pp_contributes(pp, pdf_v, pdf_nee)) w = 0.0f;   // ... which only exists inside the float range (PathPdf)
          em_value = L*w; em_len = len;
          if(R.sampler == CB_SAMPLER_PT && len > 3)
          { // path_russian_roulette (pathspace.c:273-292)
            const float p_survival = fminf(1.0f, thr/s.thr_prev);
            const float rr = point_dim(R.points, index, rand_beg_v + 4);
            if(rr >= p_survival) stop = true;
            else { thr *= 1.0f/p_survival; pdf_v_final = pdf_v*p_survival; }
          }
        }
        const bool nee_only = R.sampler == CB_SAMPLER_PTNEE;   // ptnee.c: next events without a competing technique, weight 1
        if((R.sampler == CB_SAMPLER_PTDL || nee_only) && len >= R.max_path_len) stop = true;
        const PathPdf pp_v = pp_mul(pp, pdf_v_final);        // product of v[1..vi].pdf
        uint32_t mt = s.bits >> 8;
        int rand_cnt_v = 5;
        if(vi == 1) rand_cnt_v = 1;
        // ---- next event estimation (ptdl.c:136-148, nee.h:87-243)
        if(!stop && (R.sampler == CB_SAMPLER_PTDL || nee_only))
        {
          if(!nee_only) (void)point_mt(R.points, index, mt++);   // ptdl's rr draw against nee_probability() == 1 (ptdl.c:136-137)
          if(len >= 32) stop = true;               // nee_sample refuses at PATHSPACE_MAX_VERTS and the sampler returns
          else
          {
            const int rb = rand_beg_v + rand_cnt_v; // rand_beg of the nee vertex (nee.h:107)
            if((v.material_modes & (M_DIFFUSE | M_GLOSSY)) && (R.lights.num > 0 || R.p_sky > 0.0f))
            {
              const float r0 = point_dim(R.points, index, rb + 0);
              if(r0 < R.p_sky)
              { // connect to the sky: sky_cloudy_sample (shader.c:281-326), then the common tail of nee_sample (nee.h:184-243)
                const float x1 = point_dim(R.points, index, rb + 2), x2 = point_dim(R.points, index, rb + 3);
                float edf, pdf_sky;
                const V3 d = sky_sample(R, x1, x2, s.lambda, edf, pdf_sky);
                edf = edf/R.p_sky;
                if(edf > 0.0f)
                {
                  Vtx vb = v;
                  const float bsdf = vtx_eval<KINDS>(R.mats, vb, omega, d, s.lambda, s.cur_ior);
                  float T_nee = 1.0f, vol_pdf = 1.0f;
                  bool edge_ok = true;
                  if(MEDIA && bsdf > 0.0f)
                  { // path_edge_init_volume for the connection edge (nee.h:192), its transmittance with the environment clamp
                    // (shader.c:57-60) and shader_vol_pdf for the extension the sample competes with (shader.c:107-131)
                    Vol ve = vol;
                    if(!VOLV && (vb.mode & M_TRANSMIT))
                    {
                      Media tm = med;
                      edge_ok = media_transmit(tm, v.prim_lo >> 3, v.ior, v.flags & F_INSIDE, (uint32_t)R.mats.mat[v.mat].medium);
                      ve = medium_eval(R.mats, media_medium(tm, R.exterior_medium), s.lambda);
                    }
                    if(ve.present)
                    {
                      T_nee = expf(-10000.0f*ve.mu_t);
                      if(ve.mu_s > 0.0f) vol_pdf = expf(-R.sky_far*ve.mu_t);
                    }
                  }
                  if(bsdf > 0.0f && edge_ok)
                  {
                    const V3 lx = mk3(v.x.x + R.sky_far*d.x, v.x.y + R.sky_far*d.y, v.x.z + R.sky_far*d.z);
                    const float eps = 1e-4f*max3abs(v.x);                    // prims_get_ray towards a vertex without primitive
                    V3 rd = sub(lx, v.x);
                    const float ilen = 1.0f/sqrtf(dot(rd, rd));
                    rd = mk3(rd.x*ilen, rd.y*ilen, rd.z*ilen);
                    const V3 rp = VOLV ? v.x : mk3(v.x.x + eps*rd.x, v.x.y + eps*rd.y, v.x.z + eps*rd.z);
                    const V3 dv = sub(lx, rp);
                    const float total_dist = sqrtf(dot(dv, dv));
                    const float Gl = VOLV ? 1.0f : fabsf(dot(vb.n, d));      // path_G with an environment end point
                    const float thr_l = ((thr*bsdf)*(T_nee*edf))*Gl;
                    const float pdf_nee = pdf_sky*R.p_sky;
                    const float pdf_ext = (vol_pdf*vtx_pdf<KINDS>(R.mats, vb, omega, d))*Gl;
                    const float w = nee_only ? 1.0f : pp_contributes(pp_v, pdf_nee, pdf_ext) ? pdf_nee/(pdf_ext + pdf_nee) : 0.0f;
                    if(thr_l*w > 0.0f && total_dist > 0.0f)
                    {
                      have_nee = true;
                      for(int k=0;k<3;k++) { nray.pos[k] = (&rp.x)[k]; nray.dir[k] = (&rd.x)[k]; }
                      nray.time = s.time; nray.min_dist = 0.0f;
                      nray.ignore[0] = v.prim_lo; nray.ignore[1] = v.prim_hi;
                      nmax = total_dist;
                      nrec.value = thr_l*w; nrec.lambda = s.lambda; nrec.pixel_i = s.pixel_i; nrec.pixel_j = s.pixel_j;
                      nrec.total_dist = total_dist; nrec.light_lo = nrec.light_hi = 0xffffffffu; nrec.len = (uint32_t)len + 1u;
                    }
                  }
                }
              }
              else if(r0 < R.p_sky + R.lights.p_geo)
              {
                const float rl = point_dim(R.points, index, rb + 1), rx = point_dim(R.points, index, rb + 2), ry = point_dim(R.points, index, rb + 3);
                const uint32_t t = sample_cdf_dev(R.lights.cdf, R.lights.num, rl);
                const uint64_t lpid = R.lights.primid[t];
                Vtx l;
                l.flags = 0; l.mode = 0; l.material_modes = 0; l.s = l.t = 0.0f;
                light_point(R.geo, lpid, rx, ry, s.time, l);
                V3 d = sub(l.x, v.x);
                const float dist = sqrtf(dot(d, d));
                const float il = (float)(1.0/(double)dist);
                d = mk3(d.x*il, d.y*il, d.z*il);
                prepare_vertex<KINDS>(R.geo, R.mats, l, d, s.time, s.lambda, s.scramble, med, s.cur_ior);
                const float pdf_l = R.lights.L[t];
                float edf = l.em/pdf_l;
                if(l.roughness > 1.0f - 1e-4f) edf *= (float)(1.0/PI_D);
                else
                {
                  const float phongexp = 2.0f/(l.roughness*l.roughness) - 2.0f;
                  edf *= (float)((double)(powf(-dot(l.gn, d), phongexp)*(phongexp + 2.0f))/(2.0*PI_D));
                }
                edf = edf/R.lights.p_geo;
                if(edf > 0.0f)
                {
                  Vtx vb = v;   // shader_brdf sets the mode on v; path_pop resets it afterwards
                  const float bsdf = vtx_eval<KINDS>(R.mats, vb, omega, d, s.lambda, s.cur_ior);
                  float T_nee = 1.0f, vol_pdf = 1.0f;
                  bool edge_ok = true;
                  if(MEDIA && bsdf > 0.0f)
                  { // path_edge_init_volume for the connection edge (nee.h:192), shader_vol_transmittance (:205), shader_vol_pdf
                    Vol ve = vol;
                    if(!VOLV && (vb.mode & M_TRANSMIT))
                    {
                      Media tm = med;
                      edge_ok = media_transmit(tm, v.prim_lo >> 3, v.ior, v.flags & F_INSIDE, (uint32_t)R.mats.mat[v.mat].medium);
                      ve = medium_eval(R.mats, media_medium(tm, R.exterior_medium), s.lambda);
                    }
                    if(ve.present)
                    {
                      T_nee = expf(-dist*ve.mu_t);
                      if(ve.mu_s > 0.0f) vol_pdf = T_nee;
                    }
                  }
                  if(bsdf > 0.0f && edge_ok)
                  {
                    // path_visible: prims_get_ray (prims.c:390-492); a vertex without primitive starts at its own position
                    const float eps = 1e-4f*max3abs(v.x);
                    V3 rd = sub(l.x, v.x);
                    const float ilen = 1.0f/sqrtf(dot(rd, rd));
                    rd = mk3(rd.x*ilen, rd.y*ilen, rd.z*ilen);
                    const V3 rp = VOLV ? v.x : mk3(v.x.x + eps*rd.x, v.x.y + eps*rd.y, v.x.z + eps*rd.z);
                    const V3 dv = mk3(l.x.x - eps*rd.x - rp.x, l.x.y - eps*rd.y - rp.y, l.x.z - eps*rd.z - rp.z);
                    const float total_dist = sqrtf(dot(dv, dv));
                    if(!(dot(l.gn, rd) >= 0.0f) && total_dist > 0.0f)
                    {
                      const float Gl = VOLV ? fabsf(dot(l.n, d))/(dist*dist) : cos_lambert(vb, l, d, dist);
                      const float thr_l = ((thr*bsdf)*(T_nee*edf))*Gl;
                      // mis against extending the path into the light (ptdl.c:142-146)
                      const float pdf_nee = R.lights.p_geo*pdf_l;
                      const float pdf_ext = (vol_pdf*vtx_pdf<KINDS>(R.mats, vb, omega, d))*Gl;
                      const float w = nee_only ? 1.0f : pp_contributes(pp_v, pdf_nee, pdf_ext) ? pdf_nee/(pdf_ext + pdf_nee) : 0.0f;
                      if(thr_l*w > 0.0f)
                      {
                        have_nee = true;
                        for(int k=0;k<3;k++) { nray.pos[k] = (&rp.x)[k]; nray.dir[k] = (&rd.x)[k]; }
                        nray.time = s.time; nray.min_dist = 0.0f;
                        nray.ignore[0] = v.prim_lo; nray.ignore[1] = v.prim_hi;
                        // the search ends where the ray first crosses the light primitive itself (if it does before total_dist)
                        RayD tr; HitD th;
                        tr.px = rp.x; tr.py = rp.y; tr.pz = rp.z; tr.dx = rd.x; tr.dy = rd.y; tr.dz = rd.z;
                        tr.time = s.time; tr.min_dist = 0.0f; tr.ign_lo = v.prim_lo; tr.ign_hi = v.prim_hi;
                        th.dist = total_dist; th.u = th.v = 0.0f; th.prim_lo = th.prim_hi = 0xffffffffu;
                        prim_test_geo(R.geo, lpid, tr, th);
                        nmax = th.dist;
                        nrec.value = thr_l*w; nrec.lambda = s.lambda; nrec.pixel_i = s.pixel_i; nrec.pixel_j = s.pixel_j;
                        nrec.total_dist = total_dist; nrec.light_lo = l.prim_lo; nrec.light_hi = l.prim_hi; nrec.len = (uint32_t)len + 1u;
                      }
                    }
                  }
                }
              }
            }
            rand_cnt_v += 4;   // path_pop folds the nee vertex' dimensions into v (pathspace.c:298)
          }
        }
        // ---- extend: sample the bsdf at v (pathspace.c:186-203, shader.c:577-590)
        if(!stop && thr > 0.0f && len < 32)
        {
          const int rb = rand_beg_v + rand_cnt_v;   // rand_beg of vertex vi+1 (pathspace.c:199)
          const float rx = point_dim(R.points, index, rb + 1), ry = point_dim(R.points, index, rb + 2), rm = point_dim(R.points, index, rb + 3);
          V3 wo; float pdf = 1.0f;
          v.mode &= M_EMIT;
          float weight;
          if constexpr(VOLV)
          { // medium_rgb.c:62-73: Henyey-Greenstein around the incoming direction, weight mu_s
            (void)rm;
            float out[3];
            hg_sample(v.vol_g, rx, ry, out, pdf);
            v.mode |= M_GLOSSY | M_VOLUME;
            wo = mk3(v.n.x*out[0] + v.a.x*out[1] + v.b.x*out[2], v.n.y*out[0] + v.a.y*out[1] + v.b.y*out[2], v.n.z*out[0] + v.a.z*out[1] + v.b.z*out[2]);
            weight = v.vol_mu_s;
          }
          else weight = bsdf_sample<KINDS>(R.mats, v, omega, s.lambda, s.cur_ior, rx, ry, rm, wo, pdf);
          wo = normalise(wo);
          const float dt = ((v.flags & F_INSIDE) ? -1.0f : 1.0f)*dot(v.gn, wo);
          if(((v.mode & M_REFLECT) && dt < 0.0f) || ((v.mode & M_TRANSMIT) && dt > 0.0f)) weight = 0.0f;
          const float thr_next = thr*weight;
          bool ok = thr_next > 0.0f;
          Vol vnext = vol;   // medium of the next edge: unchanged unless the path crosses the interface
          if(!VOLV && ok && (v.mode & M_TRANSMIT))
          { // path_edge_init_volume for the next edge
            ok = media_transmit(med, v.prim_lo >> 3, v.ior, v.flags & F_INSIDE, MEDIA ? (uint32_t)R.mats.mat[v.mat].medium : 0u);
            s.cur_ior = media_ior(med);
            if(MEDIA) vnext = medium_eval(R.mats, media_medium(med, R.exterior_medium), s.lambda);
          }
          if(ok)
          {
            alive = true;
            s.thr_prev = thr; s.thr = thr_next; s.pdf_proj = pdf;
            s.cos_prev = VOLV ? 1.0f : fabsf(dot(v.n, wo));
            s.x[0] = v.x.x; s.x[1] = v.x.y; s.x[2] = v.x.z;
            s.omega[0] = wo.x; s.omega[1] = wo.y; s.omega[2] = wo.z;
            s.prim_lo = v.prim_lo; s.prim_hi = v.prim_hi;
            s.lre = lre_pack(len, rb, pp_v.m != pp_v.m ? PP_ZERO : pp_v.e);
            s.pdf_m = pp_v.m;
            s.bits = (mt << 8) | ((v.material_modes & (M_DIFFUSE | M_GLOSSY)) ? 1u : 0u);
            s.med_n = media_pack(med);
            for(int k=0;k<MED_MAX;k++) { s.med_shape[k] = med.shape[k]; s.med_ior[k] = med.ior[k]; }
            next_pos = VOLV ? v.x : offset_origin(v.x, wo);   // prims_offset_ray, only on geometry (pathspace.c:759-761)
            next_dir = wo;
            // a scattering medium ahead: draw the free flight now, it clips the next closest-hit ray (pathspace.c:742-747)
            if(MEDIA && vnext.present && vnext.mu_s > 0.0f) next_clip = vol_free_flight(vnext, point_dim(R.points, index, rb + 0));
          }
        }
      }
    }
  }
  // warp-aggregated compaction of surviving paths and of pending next events
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t ma = __ballot_sync(0xffffffffu, alive);
  const uint32_t mn = __ballot_sync(0xffffffffu, have_nee);
  // one atomic reserves the warp's slots in both lists: surviving paths in the low, next-event records in the high word of cnt->next
  // (the two separate round trips were 9 % of the kernel's stall samples)
  unsigned long long both = 0;
  if(ma | mn)
  {
    if(lane == 0) both = atomicAdd(&cnt->next, (unsigned long long)__popc(ma) | ((unsigned long long)__popc(mn) << 32));
    both = __shfl_sync(0xffffffffu, both, 0);
  }
  if(alive)
  {
    const uint64_t o = (uint32_t)both + __popc(ma & ((1u << lane) - 1u));
    st_out[o] = s;
    write_ray(rays_out, o, next_pos, next_dir, s.time, s.prim_lo, s.prim_hi);
    if(MEDIA) maxd_out[o] = next_clip;
  }
  const bool have_em = em_value > 0.0f && em_value < FLT_MAX;     // view_splat's own acceptance test (view.c:457-459)
  const uint32_t me = __ballot_sync(0xffffffffu, have_em);
  if(me)
  {
    unsigned long long base = 0;
    if(lane == (uint32_t)(__ffs(me) - 1)) base = atomicAdd(&cnt->em, (unsigned long long)__popc(me));
    base = __shfl_sync(0xffffffffu, base, __ffs(me) - 1);
    if(have_em)
    {
      NeeRec e;
      e.value = em_value; e.lambda = s.lambda; e.pixel_i = s.pixel_i; e.pixel_j = s.pixel_j; e.total_dist = 0.0f;
      e.light_lo = e.light_hi = 0xffffffffu; e.len = (uint32_t)em_len;
      em_recs[base + __popc(me & ((1u << lane) - 1u))] = e;
    }
  }
  if(have_nee)
  {
    const uint64_t o = (uint32_t)(both >> 32) + __popc(mn & ((1u << lane) - 1u));
    nee_rays[o] = nray; nee_maxdist[o] = nmax; nee_recs[o] = nrec; nee_light[o] = make_uint2(nrec.light_lo, nrec.light_hi);
  }
}

// path_visible's decision (pathspace.c:325-331) + the splat of ptdl.c:146
__global__ void __launch_bounds__(RB)
k_nee_resolve(RenderDev R, uint32_t n, const NeeRec *__restrict__ recs, const int32_t *__restrict__ vis, ShadeCounters *cnt)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  bool did = false;
  if(i < n && (!vis || vis[i]))
  {
    const NeeRec r = recs[i];
    did = splat(R, r.pixel_i, r.pixel_j, r.lambda, r.value, (int)r.len);
  }
  const uint32_t m = __ballot_sync(0xffffffffu, did);
  if(m && (threadIdx.x & 31u) == 0) atomicAdd(&cnt->splats, (unsigned long long)__popc(m));
}

__global__ void k_offset_ray(const float *__restrict__ x, const float *__restrict__ dir, float *__restrict__ out, uint32_t n)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  const V3 o = offset_origin(mk3(x[3*i], x[3*i+1], x[3*i+2]), mk3(dir[3*i], dir[3*i+1], dir[3*i+2]));
  out[3*i] = o.x; out[3*i+1] = o.y; out[3*i+2] = o.z;
}

__global__ void k_bsdf(MaterialsDev M, int32_t material, const cb_bsdf_query_t *__restrict__ q, cb_bsdf_result_t *__restrict__ out, uint32_t n)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  const cb_bsdf_query_t Q = q[i];
  Vtx v;
  v.x = mk3(0.0f, 0.0f, 0.0f);
  v.n = v.gn = mk3(0.0f, 0.0f, Q.flip ? -1.0f : 1.0f);
  onb(v.n, v.a, v.b);
  v.u = v.v = v.s = v.t = 0.0f;
  v.prim_lo = 0u; v.prim_hi = CB_PRIM_LINE << 29;   // battle-test.c:84-88: shape 0, vcnt 2
  v.flags = 0; v.mode = M_ABSORB; v.material_modes = 0;
  v.rd = Q.rd; v.rs = Q.rs; v.rg = Q.rg; v.em = 0.0f; v.roughness = Q.roughness;
  v.ior = 1.0f; v.eta = 1.0f; v.mat = material;
  Media med; med.n = 0;
  for(int k=0;k<MED_MAX;k++) { med.shape[k] = 0; med.ior[k] = 1.0f; }
  const float cur_ior = 1.0f;
  bsdf_prepare(M, v, Q.lambda, med, cur_ior);
  const V3 wi = mk3(Q.wi[0], Q.wi[1], Q.wi[2]), wo_q = mk3(Q.wo[0], Q.wo[1], Q.wo[2]);
  cb_bsdf_result_t R;
  Vtx vs = v;
  V3 wo = mk3(0.0f, 0.0f, 0.0f);
  float pdf = 1.0f;
  R.s_weight = bsdf_sample(M, vs, wi, Q.lambda, cur_ior, Q.rand[0], Q.rand[1], Q.rand[2], wo, pdf);
  R.s_wo[0] = wo.x; R.s_wo[1] = wo.y; R.s_wo[2] = wo.z; R.s_pdf = pdf; R.s_mode = vs.mode;
  Vtx vb = v;
  R.f = bsdf_eval(M, vb, wi, wo_q, Q.lambda, cur_ior);
  R.f_mode = vb.mode;
  R.pdf = bsdf_pdf(M, vb, wi, wo_q);
  out[i] = R;
}

// cb200_render_medium: the medium helpers of shading.cuh exactly as k_path_start / k_shade<8> call them
__global__ void k_medium(MaterialsDev M, uint32_t medium, const cb_medium_query_t *__restrict__ q, cb_medium_result_t *__restrict__ out, uint32_t n)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  const cb_medium_query_t Q = q[i];
  cb_medium_result_t o;
  const Vol vol = medium_eval(M, medium, Q.lambda);
  o.mu_t = vol.mu_t; o.mu_s = vol.mu_s;
  o.free_dist = FLT_MAX; o.free_pdf = 1.0f;
  o.transmittance = expf(-Q.dist*vol.mu_t);
  o.vol_pdf = 1.0f;
  if(vol.mu_s > 0.0f)
  {
    o.free_dist = vol_free_flight(vol, Q.rand[2]);
    o.free_pdf = expf(-o.free_dist*vol.mu_t)*vol.mu_t;
    o.vol_pdf = o.transmittance*vol.mu_t;
  }
  const V3 wi = mk3(Q.wi[0], Q.wi[1], Q.wi[2]), wo_q = mk3(Q.wo[0], Q.wo[1], Q.wo[2]);
  Vtx v;
  v.n = v.gn = wi;
  scrambled_onb(0.5f, v.n, v.a, v.b);
  v.mode = M_ABSORB; v.material_modes = M_VOLUME | M_GLOSSY;
  v.vol_mu_s = vol.mu_s; v.vol_g = vol.g;
  float out3[3], pdf;
  hg_sample(v.vol_g, Q.rand[0], Q.rand[1], out3, pdf);
  o.s_wo[0] = v.n.x*out3[0] + v.a.x*out3[1] + v.b.x*out3[2];
  o.s_wo[1] = v.n.y*out3[0] + v.a.y*out3[1] + v.b.y*out3[2];
  o.s_wo[2] = v.n.z*out3[0] + v.a.z*out3[1] + v.b.z*out3[2];
  o.s_weight = v.vol_mu_s; o.s_pdf = pdf; o.s_mode = M_GLOSSY | M_VOLUME;
  Vtx vb = v;
  o.f = vtx_eval<16>(M, vb, wi, wo_q, Q.lambda, 1.0f);
  o.f_mode = vb.mode;
  o.pdf = vtx_pdf<16>(M, vb, wi, wo_q);
  out[i] = o;
}

__global__ void k_points(PointsDev P, const uint64_t *index, const int32_t *dim, float *out, uint32_t n)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i < n) out[i] = point_dim(P, index[i], dim[i]);
}

// emissive primitive areas at shutter open (prims_get_area, prims.c:133-153; triangle.h:36-49)
__global__ void k_prim_area(SceneGeo S, const uint64_t *pid, uint32_t n, float *area)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  const uint64_t p = pid[i];
  const uint32_t vcnt = (uint32_t)(p >> 61) & 7u;
  float a = 0.0f;
  if(vcnt == CB_PRIM_TRI || vcnt == CB_PRIM_QUAD)
  {
    const cb_vtx_t *q0 = geo_vtx(S, p, 0, 0), *q1 = geo_vtx(S, p, 1, 0), *q2 = geo_vtx(S, p, 2, 0);
    const V3 v0 = mk3(q0->v[0], q0->v[1], q0->v[2]), v1 = mk3(q1->v[0], q1->v[1], q1->v[2]), v2 = mk3(q2->v[0], q2->v[1], q2->v[2]);
    V3 nn = cross(sub(v1, v0), sub(v2, v0));
    a = sqrtf(dot(nn, nn))*.5f;
    if(vcnt == CB_PRIM_QUAD)
    {
      const cb_vtx_t *q3 = geo_vtx(S, p, 3, 0);
      const V3 v3 = mk3(q3->v[0], q3->v[1], q3->v[2]);
      nn = cross(sub(v2, v0), sub(v3, v0));
      a = a + sqrtf(dot(nn, nn))*.5f;
    }
  }
  else if(vcnt == CB_PRIM_SPHERE)
  {
    const float r = __uint_as_float(geo_vtx(S, p, 0, 0)->n);
    a = (float)(4.0*PI_D*(double)r)*r;
  }
  area[i] = a;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static const float k_fstop[] = { 0.5f, 0.7f, 1.0f, 1.4f, 2, 2.8f, 4, 5.6f, 8, 11, 16, 22, 32, 45, 64, 90, 128 };   // view.c:71-73
static const float k_exposure[] = { 60.0f, 30.0f, 15.0f, 8.0f, 4.0f, 2.0f, 1.0f, 0.5f, 1.0f/4.0f, 1.0f/8.0f, 1.0f/15.0f, 1.0f/30.0f,
  1.0f/60.0f, 1.0f/125.0f, 1.0f/250.0f, 1.0f/500.0f, 1.0f/1000.0f, 1.0f/2000.0f, 1.0f/4000.0f, 1.0f/8000.0f };     // view.c:75-79

#define PATH_STAT_DOUBLES (32*33*2)
struct cb200_render
{
  cb200_accel *accel;
  RenderDev dev;
  cb_render_desc_t desc;         // scalars only after create: the caller's arrays are not kept
  std::vector<char> mat_ok;      // material i describes a surface the path supports (cb200_render_bsdf)
  uint64_t batch;
  // owned device memory
  std::vector<void *> owned;
  PathState *st[2];
  cb_ray_t *rays[2];
  cb_hitrec_t *hits;
  float *maxd[2];                // sampled free-flight distance of every pending ray (scenes with media only), rides with rays[]
  cb_ray_t *nee_rays; float *nee_md; NeeRec *nee_recs; uint2 *nee_light; int32_t *nee_vis;
  NeeRec *em_recs;               // emission found by extension, queued by k_shade for k_nee_resolve
  uint32_t nee_deferred;         // next-event records of small waves parked at the front of the nee arrays: their shadow sweep runs
                                 // once, with the records of the following small waves, at the end of the call (render_collect)
  ShadeCounters *d_cnt, *h_cnt;
  cb_render_stats_t stats;
  // wave ordering by pixel (k_pixel_keys + radix sort)
  uint32_t *keys[2], *order[2];
  uint32_t *hit_list;            // slots of the current wave whose ray hit something (k_compact_hits)
  void *sort_tmp; size_t sort_tmp_bytes;
  // streaming wavefront: paths still alive when a pass has started all of its indices stay in the pool
  // (st[cur] / rays[cur], slots [0, n_alive)) and ride along with the next pass' waves until cb200_render_flush
  uint32_t n_alive; int cur;
  float *own_fb;
  float *own_dbor;               // cb200_render_set_dbor
  double *own_stat;              // cb200_render_path_stats
  // atomic-free accumulation (cb200_render_set_accumulation): the record list lives in dev.tile_*; sort outputs and scratch here
  int tile_mode; uint32_t tile_ub;    // tile_ub: upper bound of the records appended since the last k_tile_accumulate pass
  uint32_t *tile_key2, *tile_idx2, *tile_offsets; void *tile_tmp; size_t tile_tmp_bytes; int tile_bits;
  float4 *tile_a; float *tile_b; uint32_t *tile_key, *tile_idx;   // owned copies of the dev.tile_* pointers (dev's are null while the mode is off)
  // asynchronous snapshots: device-side copy of the accumulation buffer, drained to the host on a stream of its own
  float *snap_stage; cudaStream_t snap_stream; cudaEvent_t snap_ready, snap_done; int snap_pending;
  int bsdf_kinds;   // bit mask of the BSDF kinds referenced by shapes (selects the k_shade variant)
  // instrumentation (cb200_render_instrument): CUDA events around every launch on the pass' own stream, summed per
  // kernel class after the pass; ACCEL_DEBUG-style traversal counters
  int timing, counting;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_class;
  std::vector<uint32_t> ev_n;      // rays / paths the launch worked on (CB200_RENDER_TRACE)
  size_t ev_used;
  unsigned long long *d_trav_cnt;   // [0..3] closest waves, [4..7] shadow waves
};

enum { KC_START = 0, KC_CLOSEST = 1, KC_SHADE = 2, KC_SHADOW = 3, KC_RESOLVE = 4, KC_NUM = 5 };

struct TimeScope   // records an event pair around the launches of one kernel class
{
  cb200_render *r; cudaStream_t st; size_t slot;
  TimeScope(cb200_render *r_, cudaStream_t st_, int cls, uint32_t n = 0) : r(r_), st(st_), slot(0)
  {
    if(!r->timing) return;
    if(r->ev_used + 2 > r->ev_pool.size())
      for(int k=0;k<2;k++) { cudaEvent_t e; cudaEventCreate(&e); r->ev_pool.push_back(e); r->ev_class.push_back(0); r->ev_n.push_back(0); }
    slot = r->ev_used; r->ev_used += 2;
    r->ev_class[slot] = cls; r->ev_n[slot] = n;
    cudaEventRecord(r->ev_pool[slot], st);
  }
  ~TimeScope() { if(r->timing) cudaEventRecord(r->ev_pool[slot+1], st); }
};

template<typename T> static T *dev_upload(cb200_render *r, const T *src, size_t count)
{
  T *p = nullptr;
  if(cudaMalloc(&p, (count ? count : 1)*sizeof(T)) != cudaSuccess) return nullptr;
  r->owned.push_back(p);
  if(count && src && cudaMemcpy(p, src, count*sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return p;
}
template<typename T> static T *dev_alloc(cb200_render *r, size_t count)
{
  T *p = nullptr;
  if(cudaMalloc(&p, (count ? count : 1)*sizeof(T)) != cudaSuccess) return nullptr;
  r->owned.push_back(p);
  return p;
}

// Halton digit permutations exactly as halton_init_random builds them (ext/halton/halton.h:3244-3274): Fisher-Yates per base
// driven by lrand48() after srand48(frame); bases 1..3 keep the identity.  Only the 256 primes' rows are uploaded.
static int build_halton(cb200_render *r, uint64_t frame)
{
  const unsigned max_base = 1619u;
  std::vector<std::vector<uint16_t>> perms(max_base + 1);
  srand48((long)frame);
  for(unsigned base=1; base<=max_base; ++base)
  {
    perms[base].resize(base);
    for(unsigned i=0;i<base;i++) perms[base][i] = (uint16_t)i;
    if(base < 4) continue;
    for(unsigned i=0;i<base-1;i++)
    {
      const size_t j = i + lrand48() / ((1ul<<31) / (base - i) + 1);
      std::swap(perms[base][j], perms[base][i]);
    }
  }
  std::vector<uint16_t> flat, bases(HALTON_DIMS);
  std::vector<uint32_t> off(HALTON_DIMS);
  std::vector<uint8_t> digits(HALTON_DIMS);
  std::vector<float> scale(HALTON_DIMS);
  unsigned cand = 1;
  for(int d=0; d<HALTON_DIMS; d++)
  {
    for(;;) { cand++; bool prime = true; for(unsigned k=2;k<cand;k++) if(cand % k == 0) { prime = false; break; } if(prime) break; }
    bases[d] = (uint16_t)cand;
    off[d] = (uint32_t)flat.size();
    flat.insert(flat.end(), perms[cand].begin(), perms[cand].end());
    // digits per table lookup: largest power of the base <= 500; lookups: largest power of that < 2^32 (halton_gen.py)
    unsigned dg = 1; uint64_t pow_base = cand;
    while(pow_base*cand <= 500) { pow_base *= cand; dg++; }
    uint64_t max_power = pow_base; unsigned lookups = 1;
    while(max_power*pow_base < (1ull << 32)) { max_power *= pow_base; lookups++; }
    digits[d] = (uint8_t)(dg*lookups);
    scale[d] = (float)(0x1.fffffcp-1 / (double)max_power);
  }
  HaltonDev &H = r->dev.points.halton;
  H.perm = dev_upload(r, flat.data(), flat.size());
  H.perm_off = dev_upload(r, off.data(), off.size());
  H.base = dev_upload(r, bases.data(), bases.size());
  H.digits = dev_upload(r, digits.data(), digits.size());
  H.scale = dev_upload(r, scale.data(), scale.size());
  return (H.perm && H.perm_off && H.base && H.digits && H.scale) ? 0 : CB200_ERR_NOMEM;
}

static float host_rgb2spec(const float *c, float lambda)
{
  const float x = fmaf(fmaf(c[0], lambda, c[1]), lambda, c[2]);
  const float y = 1.0f/sqrtf(fmaf(x, x, 1.0f));
  return fmaf(.5f*x, y, .5f);
}

// sky_envmap.c's init() (:303-366): texels to the device, importance mip hierarchy built exactly like there (float arithmetic,
// the four children of a cell normalised by their sum where that is >= 0.001, top level = the two halves)
static float host_env_sh(const float *c)
{
  return ((host_rgb2spec(c, 660.0f)*c[3] + host_rgb2spec(c, 560.0f)*c[3]) + host_rgb2spec(c, 480.0f)*c[3]) + host_rgb2spec(c, 400.0f)*c[3];
}
static int build_envmap(cb200_render *r, const cb_envmap_t *em)
{
  EnvDev E;
  memset(&E, 0, sizeof(E));
  E.width = (int)em->width; E.height = (int)em->height; E.mul = em->mul;
  for(int k=0;k<9;k++) { E.world[k] = em->world[k]; E.world_inv[k] = em->world_inv[k]; }
  E.w2n = 1; E.h2n = 1;
  while(E.w2n <= E.width) E.w2n <<= 1;
  E.w2n >>= 1;
  while(E.h2n <= E.height) E.h2n <<= 1;
  E.h2n >>= 1;
  E.aspectx = E.width/(float)E.w2n;
  E.aspecty = E.height/(float)E.h2n;
  E.levels = 0;
  size_t size = 0;
  for(int h=E.h2n;h>0;h>>=1,E.levels++) size += (size_t)h*h*2;
  if(E.levels > 16) { cb200_set_error("render_create: environment map too large"); return CB200_ERR_ARG; }
  std::vector<float> pix(size, 0.0f);
  std::vector<float *> mip(E.levels);
  mip[0] = pix.data();
  E.mip_off[0] = 0;
  for(int k=0;k<E.levels-1;k++) { mip[k+1] = mip[k] + (size_t)(E.w2n >> k)*(E.h2n >> k); E.mip_off[k+1] = (uint32_t)(mip[k+1] - mip[0]); }
  auto fetch = [&](int i, int j) -> const float * {
    if(i < 0 || i >= E.width || j < 0 || j >= E.height) return em->pixels;
    return em->pixels + 4*((size_t)E.width*j + i);
  };
  const float b = em->mul;
  E.sum = 0.0;
  for(int j=0;j<E.h2n;j++) for(int i=0;i<E.w2n;i++)
  {
    const int ii = (int)(i*E.aspectx + 0.5f), jj = (int)(j*E.aspecty + 0.5f);
    const float sn = sinf((float)(M_PI*jj/(float)E.height));
    mip[0][i + E.w2n*j] = host_env_sh(fetch(ii, jj))*b*sn;
    E.sum += mip[0][i + E.w2n*j];
  }
  E.sum /= E.w2n*E.h2n;
  for(int m=1;m<E.levels;m++)
  {
    const int wp = E.w2n >> (m-1);
    for(int j=0;j<(E.h2n >> m);j++) for(int i=0;i<(E.w2n >> m);i++)
    {
      float &a00 = mip[m-1][i*2 + wp*j*2], &a01 = mip[m-1][i*2 + wp*(j*2+1)], &a10 = mip[m-1][i*2 + 1 + wp*j*2], &a11 = mip[m-1][i*2 + 1 + wp*(j*2+1)];
      const float all = a00 + a01 + a10 + a11;
      mip[m][i + (E.w2n >> m)*j] = .25f*all;
      if(all >= 0.001f) { a00 /= all; a10 /= all; a11 /= all; a01 /= all; }
    }
  }
  {
    const float all = mip[E.levels-1][0] + mip[E.levels-1][1];
    mip[E.levels-1][0] /= all;
    mip[E.levels-1][1] /= all;
  }
  E.px = reinterpret_cast<const float4 *>(dev_upload(r, em->pixels, (size_t)E.width*E.height*4));
  E.mip = dev_upload(r, pix.data(), pix.size());
  r->dev.env = dev_upload(r, &E, 1);
  return (E.px && E.mip && r->dev.env) ? 0 : CB200_ERR_NOMEM;
}

// lights_init_light per emissive shape + lights_prepare_frame (list.c:56-104)
static int build_lights(cb200_render *r)
{
  cb200_scene *s = r->accel->scene;
  const cb_render_desc_t &d = r->desc;
  std::vector<uint64_t> primid(s->num_prims ? s->num_prims : 1);
  if(s->num_prims) if(cudaMemcpy(primid.data(), s->d_primid, s->num_prims*sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess) return CB200_ERR_CUDA;
  std::vector<int32_t> shape_mat(s->num_shapes ? s->num_shapes : 1, 0);
  {
    std::vector<int64_t> tmp(s->num_shapes ? s->num_shapes : 1);
    memcpy(tmp.data(), s->h_material.data(), sizeof(int64_t)*s->num_shapes);
    for(int i=0;i<s->num_shapes;i++)
    {
      if(tmp[i] < 0 || tmp[i] >= d.num_materials) { cb200_set_error("render_create: shape references a material that was not supplied"); return CB200_ERR_ARG; }
      if(d.materials[tmp[i]].num_ops < 0)
      { cb200_set_error("render_create: shape uses a shader outside the pt/ptdl surface path (no CPU fallback for unknown shaders)"); return CB200_ERR_UNSUPPORTED; }
      shape_mat[i] = (int32_t)tmp[i];
    }
  }
  r->dev.geo.shape_material = dev_upload(r, shape_mat.data(), shape_mat.size());
  r->bsdf_kinds = 0;
  for(int i=0;i<s->num_shapes;i++) r->bsdf_kinds |= 1 << d.materials[shape_mat[i]].bsdf;
  r->dev.has_media = d.exterior_medium ? 1 : 0;
  for(int i=0;i<s->num_shapes;i++) if(d.materials[shape_mat[i]].medium) r->dev.has_media = 1;
  if(!r->bsdf_kinds) r->bsdf_kinds = 1;
  std::vector<float> shape_L(s->num_shapes ? s->num_shapes : 1, 0.0f);
  for(int i=0;i<s->num_shapes;i++)
  {
    const cb_material_t &m = d.materials[shape_mat[i]];
    for(int k=0;k<m.num_ops;k++) if(m.ops[k].op == CB_OP_COLOR && m.ops[k].slot == CB_SLOT_EMISSION)
    { // color.c:65-73: mean of the spectrum at 660, 560, 480, 400 nm
      const float *c = m.ops[k].coeff;
      shape_L[i] = m.ops[k].mul*(host_rgb2spec(c, 660.0f) + host_rgb2spec(c, 560.0f) + host_rgb2spec(c, 480.0f) + host_rgb2spec(c, 400.0f))/4.0f;
    }
  }
  std::vector<uint64_t> lp;
  std::vector<float> lL;
  for(uint64_t k=0;k<s->num_prims;k++)
  {
    const uint32_t sh = cb_primid_shapeid(primid[k]);
    if(shape_L[sh] > 0.0f) { lp.push_back(primid[k]); lL.push_back(shape_L[sh]); }
  }
  // the global list is in load order, i.e. grouped by ascending shape id like lights_init_light appends them
  const uint32_t n = (uint32_t)lp.size();
  std::vector<float> area(n ? n : 1), cdf(n ? n : 1), shape_pdf(s->num_shapes ? s->num_shapes : 1, 0.0f);
  LightsDev &L = r->dev.lights;
  L.num = n;
  { // lights_prepare_frame (list.c:76-88)
    float p_sky = r->desc.sky != CB_SKY_BLACK ? 1.0f : 0.0f, p_geo = n ? 1.0f : 0.0f;
    const float p_sum = p_sky + p_geo;
    if(p_sum > 0.0f) { p_sky /= p_sum; p_geo /= p_sum; }
    L.p_geo = p_geo; r->dev.p_sky = p_sky;
  }
  L.primid = dev_upload(r, lp.data(), lp.size());
  if(n)
  {
    float *d_area = dev_alloc<float>(r, n);
    k_prim_area<<<(n + 127)/128, 128>>>(r->dev.geo, L.primid, n, d_area);
    cb200_count_launch();
    if(cudaMemcpy(area.data(), d_area, n*sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return CB200_ERR_CUDA;
    float sum = 0.0f;
    for(uint32_t k=0;k<n;k++) { cdf[k] = area[k]*lL[k]; sum += cdf[k]; }
    for(uint32_t k=0;k<n;k++) lL[k] /= sum;
    for(uint32_t k=1;k<n;k++) cdf[k] += cdf[k-1];
    for(uint32_t k=0;k+1<n;k++) cdf[k] /= cdf[n-1];
    cdf[n-1] = 1.0f;
    for(uint32_t k=0;k<n;k++) shape_pdf[cb_primid_shapeid(lp[k])] = lL[k];
  }
  L.cdf = dev_upload(r, cdf.data(), cdf.size());
  L.L = dev_upload(r, lL.data(), lL.size());
  L.shape_pdf = dev_upload(r, shape_pdf.data(), shape_pdf.size());
  return (L.cdf && L.L && L.shape_pdf && L.primid && r->dev.geo.shape_material) ? 0 : CB200_ERR_NOMEM;
}

extern "C" {

void cb200_render_destroy(cb200_render_t *r)
{
  if(!r) return;
  if(r->snap_stream) { cudaStreamSynchronize(r->snap_stream); cudaStreamDestroy(r->snap_stream); cudaEventDestroy(r->snap_ready); cudaEventDestroy(r->snap_done); }
  for(void *p : r->owned) cudaFree(p);
  if(r->own_dbor) cudaFree(r->own_dbor);
  if(r->own_stat) cudaFree(r->own_stat);
  for(cudaEvent_t e : r->ev_pool) cudaEventDestroy(e);
  if(r->h_cnt) cudaFreeHost(r->h_cnt);
  delete r;
}

cb200_render_t *cb200_render_create(cb200_accel_t *a, const cb_render_desc_t *desc)
{
  if(!a || !desc || !desc->materials || desc->num_materials < 1 || desc->width == 0 || desc->height == 0)
  { cb200_set_error("render_create: bad arguments"); return nullptr; }
  if(desc->camera.aperture_value < 0 || desc->camera.aperture_value >= (int)(sizeof(k_fstop)/sizeof(float)) ||
     desc->camera.exposure_value < 0 || desc->camera.exposure_value >= (int)(sizeof(k_exposure)/sizeof(float)))
  { cb200_set_error("render_create: aperture/exposure index out of range"); return nullptr; }
  for(int i=0;i<desc->num_materials;i++)
  {
    const cb_material_t &m = desc->materials[i];
    if(m.num_ops < 0) continue;   // a shader outside the hot path (media, skies): only an error when a shape references it
    if(m.num_ops > CB_MAX_MATOPS || m.bsdf < 0 || m.bsdf > CB_BSDF_DIFFDIEL)
    { cb200_set_error("render_create: malformed material"); return nullptr; }
    if(m.bsdf == CB_BSDF_METAL && (m.table < 0 || m.table >= desc->num_tables)) { cb200_set_error("render_create: metal without ior table"); return nullptr; }
    for(int k=0;k<m.num_ops;k++) if(m.ops[k].op == CB_OP_CHECKERSG && (m.ops[k].table < 0 || m.ops[k].table >= desc->num_tables))
    { cb200_set_error("render_create: colour checker without table"); return nullptr; }
  }
  if(desc->sky != CB_SKY_BLACK && desc->sky != CB_SKY_CLOUDY && desc->sky != CB_SKY_CONST && desc->sky != CB_SKY_ENVMAP) { cb200_set_error("render_create: unsupported sky (no CPU fallback)"); return nullptr; }
  if(desc->sky == CB_SKY_ENVMAP && (!desc->envmap || !desc->envmap->pixels || desc->envmap->height < 2 || desc->envmap->width != 2*desc->envmap->height))
  { cb200_set_error("render_create: environment map missing or not 2:1 (sky_envmap.c:309)"); return nullptr; }
  if(desc->num_media < 0 || desc->num_media > CB_MAX_MEDIA || (desc->num_media > 0 && !desc->media) ||
     desc->exterior_medium < 0 || desc->exterior_medium > desc->num_media)
  { cb200_set_error("render_create: malformed media list"); return nullptr; }
  for(int i=0;i<desc->num_materials;i++)
    if(desc->materials[i].num_ops >= 0 && (desc->materials[i].medium < 0 || desc->materials[i].medium > desc->num_media))
    { cb200_set_error("render_create: material references a medium that was not supplied"); return nullptr; }
  if(desc->exterior_medium && desc->sky != CB_SKY_BLACK)
  { cb200_set_error("render_create: an exterior medium under a non-black sky is not supported (no CPU fallback)"); return nullptr; }
  if(desc->sampler != CB_SAMPLER_PT && desc->sampler != CB_SAMPLER_PTDL && desc->sampler != CB_SAMPLER_PTNEE)
  { cb200_set_error("render_create: unknown sampler (pt, ptdl and ptnee exist on the gpu; no CPU fallback)"); return nullptr; }
  cb200_render *r = new cb200_render();
  r->accel = a;
  r->desc = *desc;   // the arrays behind its pointers are the caller's and only read inside this function (copied to the device below)
  r->mat_ok.assign(desc->num_materials > 0 ? desc->num_materials : 0, 0);
  for(int i=0;i<desc->num_materials;i++) r->mat_ok[i] = desc->materials[i].num_ops >= 0;
  memset(&r->stats, 0, sizeof(r->stats));
  r->h_cnt = nullptr;
  r->timing = r->counting = 0; r->ev_used = 0; r->d_trav_cnt = nullptr;
  r->n_alive = 0; r->cur = 0; r->nee_deferred = 0;
  r->snap_stage = nullptr; r->snap_stream = nullptr; r->snap_pending = 0;
  r->own_stat = nullptr;
  r->tile_mode = 0; r->tile_ub = 0; r->tile_key2 = r->tile_idx2 = r->tile_offsets = nullptr; r->tile_tmp = nullptr; r->tile_tmp_bytes = 0; r->tile_bits = 0;
  r->tile_a = nullptr; r->tile_b = nullptr; r->tile_key = r->tile_idx = nullptr;
  r->own_dbor = nullptr;
  cb200_scene *s = a->scene;
  RenderDev &D = r->dev;
  memset(&D, 0, sizeof(D));
  D.accel = a->dev;
  D.geo.vtx = s->d_vtx; D.geo.vtxidx = s->d_vtxidx; D.geo.shapes = s->d_shapes;
  D.sampler = desc->sampler; D.colour = desc->colour_camera;
  D.sky = desc->sky;
  for(int k=0;k<3;k++) D.sky_coeff[k] = desc->sky_coeff[k];
  D.sky_scale = desc->sky_scale;
  D.sky_far = (a->aabb[3] + a->aabb[4] + a->aabb[5]) - a->aabb[0] - a->aabb[1] - a->aabb[2];
  D.max_path_len = desc->max_path_len > 0 && desc->max_path_len <= 32 ? desc->max_path_len : 32;
  D.fb_w = desc->width; D.fb_h = desc->height;
  D.cam.c = desc->camera; D.cam.width = (float)desc->width; D.cam.height = (float)desc->height;
  D.cam.fstop = k_fstop[desc->camera.aperture_value];
  D.cam.exposure_time = k_exposure[desc->camera.exposure_value];
  D.points.mode = desc->pointsampler;
  // the key carries the frame only: path indices are unique across the ranks of a sample-split job, so every draw is a function of
  // (frame, path index, dimension) and an N-GPU image equals the 1-GPU image of the same progressions up to fp32 summation order
  D.points.key = (desc->frame + 1)*0x9e3779b97f4a7c15ull;
  bool ok = true;
  // materials and tables
  std::vector<TableDev> tabs(desc->num_tables > 0 ? desc->num_tables : 1);
  std::vector<float> tdata;
  for(int i=0;i<desc->num_tables;i++)
  {
    const cb_table_t &t = desc->tables[i];
    tabs[i].lambda_min = t.lambda_min; tabs[i].lambda_step = t.lambda_step; tabs[i].num_lambda = t.num_lambda; tabs[i].rows = t.rows;
    tabs[i].offset = (uint32_t)tdata.size();
    tdata.insert(tdata.end(), t.data, t.data + (size_t)t.num_lambda*t.rows);
  }
  D.mats.mat = dev_upload(r, desc->materials, desc->num_materials);
  D.mats.tables = dev_upload(r, tabs.data(), tabs.size());
  D.mats.table_data = dev_upload(r, tdata.data(), tdata.size());
  D.mats.media = dev_upload(r, desc->media, (size_t)desc->num_media);
  D.exterior_medium = (uint32_t)desc->exterior_medium;
  ok = ok && D.mats.mat && D.mats.tables && D.mats.table_data && D.mats.media;
  ok = ok && cudaMemcpyToSymbol(c_cie, cie1931_xyz, sizeof(cie1931_xyz)) == cudaSuccess;
  ok = ok && build_halton(r, desc->frame) == 0;
  if(ok && desc->sky == CB_SKY_ENVMAP) ok = build_envmap(r, desc->envmap) == 0;
  if(ok && build_lights(r)) { cb200_render_destroy(r); return nullptr; }
  // wave buffers
  r->batch = desc->batch_paths;
  if(!r->batch) { const char *e = getenv("CB200_POOL_PATHS"); if(e) r->batch = strtoull(e, nullptr, 10); }   // measurement knob (scripts/pool_sweep_named.py)
  if(!r->batch)
  { // default: a pool of FOUR progressions (4 W*H paths, view.c:636-638), at most 2^25.  A streamed progression then is one
    // wave -- its new paths plus the survivors of the previous ones, ~20 M rays at 4K -- instead of two or three pool-sized
    // ones: every persistent traversal launch ends in a tail of half-empty warps and every wave costs a counter read-back, and
    // both are paid per launch, not per ray.  Measured on the 4K bench (profiles/r3a_batch_sweep.log): 20.7 ms per progression
    // with a pool of 2^23, 19.8 with 2^24, 19.6 with 2^25, same image.  ~560 bytes per slot; halved while that would take more than
    // half of the free device memory.
    r->batch = 4ull*desc->width*desc->height;
    if(r->batch > (1ull << 25)) r->batch = 1ull << 25;
    // small frames: room for many progressions per wave (--batch).  1024 x 576, 0011_ptdl through the command line
    // (profiles/r3f_pool_named.log): 1067 spp/s with a pool of 2^21, 1229 with 2^23, 1259 with 2^24
    if(r->batch < (1ull << 23)) r->batch = 1ull << 23;
    size_t mem_free = 0, mem_total = 0;
    if(cudaMemGetInfo(&mem_free, &mem_total) == cudaSuccess)
      while(r->batch > (1ull << 21) && r->batch > (uint64_t)desc->width*desc->height && r->batch*560ull > mem_free/2) r->batch >>= 1;
  }
  const uint64_t N = r->batch;
  D.fb = r->own_fb = dev_alloc<float>(r, (size_t)desc->width*desc->height*3);
  for(int k=0;k<2;k++) { r->st[k] = dev_alloc<PathState>(r, N); r->rays[k] = dev_alloc<cb_ray_t>(r, N); ok = ok && r->st[k] && r->rays[k]; }
  r->hits = dev_alloc<cb_hitrec_t>(r, N);
  r->hit_list = dev_alloc<uint32_t>(r, 5*N); ok = ok && r->hit_list;
  r->maxd[0] = r->maxd[1] = nullptr;
  if(D.has_media) for(int k=0;k<2;k++) { r->maxd[k] = dev_alloc<float>(r, N); ok = ok && r->maxd[k]; }
  r->nee_rays = dev_alloc<cb_ray_t>(r, N); r->nee_md = dev_alloc<float>(r, N); r->nee_recs = dev_alloc<NeeRec>(r, N); r->nee_light = dev_alloc<uint2>(r, N); r->nee_vis = dev_alloc<int32_t>(r, N);
  r->em_recs = dev_alloc<NeeRec>(r, N);
  r->d_cnt = dev_alloc<ShadeCounters>(r, 1);
  for(int k=0;k<2;k++) { r->keys[k] = dev_alloc<uint32_t>(r, N); r->order[k] = dev_alloc<uint32_t>(r, N); ok = ok && r->keys[k] && r->order[k]; }
  r->sort_tmp = nullptr; r->sort_tmp_bytes = 0;
  if(ok)
  {
    cub::DeviceRadixSort::SortPairs(nullptr, r->sort_tmp_bytes, r->keys[0], r->keys[1], r->order[0], r->order[1], (int)N, 0, 24);
    r->sort_tmp = dev_alloc<uint8_t>(r, r->sort_tmp_bytes);
    ok = ok && r->sort_tmp;
  }
  r->d_trav_cnt = dev_alloc<unsigned long long>(r, 8);
  if(r->d_trav_cnt) cudaMemset(r->d_trav_cnt, 0, 8*sizeof(unsigned long long));
  ok = ok && r->d_trav_cnt && D.fb && r->hits && r->nee_rays && r->nee_md && r->nee_recs && r->nee_light && r->nee_vis && r->em_recs && r->d_cnt;
  ok = ok && cudaMallocHost(&r->h_cnt, sizeof(ShadeCounters)) == cudaSuccess;
  if(!ok)
  {
    cb200_set_error(std::string("render_create: device allocation/upload failed: ") + cudaGetErrorString(cudaGetLastError()));
    cb200_render_destroy(r);
    return nullptr;
  }
  cudaMemset(D.fb, 0, (size_t)desc->width*desc->height*3*sizeof(float));
  cudaMemset(r->d_cnt, 0, sizeof(ShadeCounters));
  r->desc.materials = nullptr; r->desc.tables = nullptr; r->desc.media = nullptr; r->desc.envmap = nullptr;   // borrowed: the caller may free them now
  {
    const char *e = getenv("CB200_TILES");     // default accumulation mode of new render objects (A/B measurements)
    if(e && atoi(e) && cb200_render_set_accumulation(r, CB200_ACCUM_TILES)) { cb200_render_destroy(r); return nullptr; }
  }
  return r;
}

int cb200_render_clear(cb200_render_t *r, void *stream)
{
  if(!r) { cb200_set_error("render_clear: null"); return CB200_ERR_ARG; }
  CB_CUDA(cudaMemsetAsync(r->dev.fb, 0, (size_t)r->dev.fb_w*r->dev.fb_h*3*sizeof(float), (cudaStream_t)stream));
  if(r->dev.dbor) CB_CUDA(cudaMemsetAsync(r->dev.dbor, 0, (size_t)r->dev.num_dbors*r->dev.fb_w*r->dev.fb_h*3*sizeof(float), (cudaStream_t)stream));
  CB_CUDA(cudaMemsetAsync(r->d_cnt, 0, sizeof(ShadeCounters), (cudaStream_t)stream));
  CB_CUDA(cudaMemsetAsync(r->d_trav_cnt, 0, 8*sizeof(unsigned long long), (cudaStream_t)stream));
  if(r->own_stat) CB_CUDA(cudaMemsetAsync(r->own_stat, 0, PATH_STAT_DOUBLES*sizeof(double), (cudaStream_t)stream));
  if(r->tile_a && r->tile_ub)
  { // recorded samples of the dropped image
    k_fill_u32<<<cb200_sm_count_cached()*4, 256, 0, (cudaStream_t)stream>>>(r->tile_key, r->dev.tiles_x*r->dev.tiles_y, r->tile_ub < r->dev.tile_cap ? r->tile_ub : r->dev.tile_cap);
    r->tile_ub = 0;
  }
  r->n_alive = 0;   // paths still in flight are dropped with the image they belong to
  r->nee_deferred = 0;
  memset(&r->stats, 0, sizeof(r->stats));
  return 0;
}

void *cb200_render_fb_device(cb200_render_t *r) { return r ? r->dev.fb : nullptr; }

int cb200_render_set_framebuffer(cb200_render_t *r, void *d_fb)
{
  if(!r) { cb200_set_error("render_set_framebuffer: null"); return CB200_ERR_ARG; }
  r->dev.fb = d_fb ? (float *)d_fb : r->own_fb;
  return 0;
}

// an asynchronous snapshot still on its way must land before anything else is written to (possibly) the same host buffer
static int snapshot_drain(cb200_render *r)
{
  if(r->snap_pending) { CB_CUDA(cudaEventSynchronize(r->snap_done)); r->snap_pending = 0; }
  return 0;
}

int cb200_render_snapshot_wait(cb200_render_t *r)
{
  if(!r) { cb200_set_error("render_snapshot_wait: null"); return CB200_ERR_ARG; }
  return snapshot_drain(r);
}

int cb200_render_snapshot_async(cb200_render_t *r, float *fb_host, void *stream_)
{
  if(!r || !fb_host) { cb200_set_error("render_snapshot_async: bad arguments"); return CB200_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream_;
  const size_t bytes = (size_t)r->dev.fb_w*r->dev.fb_h*3*sizeof(float);
  if(!r->snap_stream)
  {
    CB_CUDA(cudaStreamCreateWithFlags(&r->snap_stream, cudaStreamNonBlocking));
    CB_CUDA(cudaEventCreateWithFlags(&r->snap_ready, cudaEventDisableTiming));
    CB_CUDA(cudaEventCreateWithFlags(&r->snap_done, cudaEventDisableTiming));
    r->snap_stage = dev_alloc<float>(r, (size_t)r->dev.fb_w*r->dev.fb_h*3);
    if(!r->snap_stage) { cb200_set_error("render_snapshot_async: out of device memory"); return CB200_ERR_NOMEM; }
  }
  if(r->snap_pending) CB_CUDA(cudaStreamWaitEvent(st, r->snap_done, 0));   // the staging copy is still being drained
  CB_CUDA(cudaMemcpyAsync(r->snap_stage, r->dev.fb, bytes, cudaMemcpyDeviceToDevice, st));
  CB_CUDA(cudaEventRecord(r->snap_ready, st));
  CB_CUDA(cudaStreamWaitEvent(r->snap_stream, r->snap_ready, 0));
  CB_CUDA(cudaMemcpyAsync(fb_host, r->snap_stage, bytes, cudaMemcpyDeviceToHost, r->snap_stream));
  CB_CUDA(cudaEventRecord(r->snap_done, r->snap_stream));
  r->snap_pending = 1;
  return 0;
}

int cb200_render_download(cb200_render_t *r, float *fb_host, void *stream)
{
  if(!r || !fb_host) { cb200_set_error("render_download: bad arguments"); return CB200_ERR_ARG; }
  { const int rc = snapshot_drain(r); if(rc) return rc; }
  if(r->n_alive) { const int rc = cb200_render_flush(r, stream); if(rc) return rc; }
  CB_CUDA(cudaMemcpyAsync(fb_host, r->dev.fb, (size_t)r->dev.fb_w*r->dev.fb_h*3*sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

// `--dbor n` (view.c:291: clamped to [0, 20]; the cascade exists for n > 1, view.c:339)
int cb200_render_set_dbor(cb200_render_t *r, int32_t levels)
{
  if(!r) { cb200_set_error("render_set_dbor: null"); return CB200_ERR_ARG; }
  if(levels < 0) levels = 0;
  if(levels > 20) levels = 20;
  CB_CUDA(cudaDeviceSynchronize());
  if(r->own_dbor) { cudaFree(r->own_dbor); r->own_dbor = nullptr; }
  r->dev.dbor = nullptr; r->dev.num_dbors = 0;
  if(levels > 1)
  {
    const size_t bytes = (size_t)levels*r->dev.fb_w*r->dev.fb_h*3*sizeof(float);
    if(cudaMalloc(&r->own_dbor, bytes) != cudaSuccess) { r->own_dbor = nullptr; cb200_set_error("render_set_dbor: out of device memory"); return CB200_ERR_NOMEM; }
    CB_CUDA(cudaMemset(r->own_dbor, 0, bytes));
    r->dev.dbor = r->own_dbor; r->dev.num_dbors = levels;
  }
  return 0;
}

int cb200_render_num_dbors(cb200_render_t *r) { return r ? r->dev.num_dbors : 0; }

void *cb200_render_dbor_device(cb200_render_t *r, int32_t level)
{
  if(!r || level < 0 || level >= r->dev.num_dbors) return nullptr;
  return r->dev.dbor + (size_t)level*r->dev.fb_w*r->dev.fb_h*3;
}

int cb200_render_download_dbor(cb200_render_t *r, int32_t level, float *fb_host, void *stream)
{
  if(!r || !fb_host || level < 0 || level >= r->dev.num_dbors) { cb200_set_error("render_download_dbor: bad arguments"); return CB200_ERR_ARG; }
  if(r->n_alive) { const int rc = cb200_render_flush(r, stream); if(rc) return rc; }
  const size_t count = (size_t)r->dev.fb_w*r->dev.fb_h*3;
  CB_CUDA(cudaMemcpyAsync(fb_host, r->dev.dbor + count*level, count*sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int cb200_render_snapshot(cb200_render_t *r, float *fb_host, void *stream)
{
  if(!r || !fb_host) { cb200_set_error("render_snapshot: bad arguments"); return CB200_ERR_ARG; }
  { const int rc = snapshot_drain(r); if(rc) return rc; }
  CB_CUDA(cudaMemcpyAsync(fb_host, r->dev.fb, (size_t)r->dev.fb_w*r->dev.fb_h*3*sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int cb200_render_stats(cb200_render_t *r, cb_render_stats_t *out)
{
  if(!r || !out) { cb200_set_error("render_stats: bad arguments"); return CB200_ERR_ARG; }
  *out = r->stats;
  return 0;
}

// view->stat_enery / stat_cnt (src/view.c:46-47,470-471): energy and number of splats per path length, what view_print_info draws
// as the histogram of the sidecar file (view.c:759-790)
int cb200_render_path_stats(cb200_render_t *r, int enable)
{
  if(!r) { cb200_set_error("render_path_stats: null"); return CB200_ERR_ARG; }
  CB_CUDA(cudaDeviceSynchronize());
  if(enable && !r->own_stat)
  {
    if(cudaMalloc(&r->own_stat, PATH_STAT_DOUBLES*sizeof(double)) != cudaSuccess) { r->own_stat = nullptr; cb200_set_error("render_path_stats: out of device memory"); return CB200_ERR_NOMEM; }
    CB_CUDA(cudaMemset(r->own_stat, 0, PATH_STAT_DOUBLES*sizeof(double)));
  }
  r->dev.stat = enable ? r->own_stat : nullptr;
  return 0;
}

int cb200_render_get_path_stats(cb200_render_t *r, double energy[33], uint64_t count[33])
{
  if(!r || !energy || !count) { cb200_set_error("render_get_path_stats: bad arguments"); return CB200_ERR_ARG; }
  for(int k=0;k<33;k++) { energy[k] = 0.0; count[k] = 0; }
  if(!r->own_stat) return 0;
  std::vector<double> h(PATH_STAT_DOUBLES);
  CB_CUDA(cudaMemcpy(h.data(), r->own_stat, PATH_STAT_DOUBLES*sizeof(double), cudaMemcpyDeviceToHost));
  for(int lane=0;lane<32;lane++) for(int k=0;k<33;k++) { energy[k] += h[(lane*33 + k)*2]; count[k] += (uint64_t)h[(lane*33 + k)*2 + 1]; }
  return 0;
}

int cb200_render_instrument(cb200_render_t *r, int timing, int counters)
{
  if(!r) { cb200_set_error("render_instrument: null"); return CB200_ERR_ARG; }
  r->timing = timing ? 1 : 0;
  r->counting = counters ? 1 : 0;
  return 0;
}

// sort the recorded samples by tile and add them to the framebuffer: four checkerboard phases of one block per tile
static int tile_pass(cb200_render *r, cudaStream_t st)
{
  const uint32_t n = r->tile_ub < r->dev.tile_cap ? r->tile_ub : r->dev.tile_cap;
  r->tile_ub = 0;
  if(!n) return 0;
  TimeScope ts(r, st, KC_RESOLVE, n);
  size_t tmp = r->tile_tmp_bytes;
  CB_CUDA(cub::DeviceRadixSort::SortPairs(r->tile_tmp, tmp, r->dev.tile_key, r->tile_key2, r->dev.tile_idx, r->tile_idx2, (int)n, 0, r->tile_bits, st));
  const uint32_t tiles = r->dev.tiles_x*r->dev.tiles_y;
  k_tile_offsets<<<(n + 1 + 255)/256, 256, 0, st>>>(r->tile_key2, n, tiles, r->tile_offsets);
  const dim3 grid((r->dev.tiles_x + 1)/2, (r->dev.tiles_y + 1)/2);
  for(int ph=0;ph<4;ph++)
  {
    if(r->dev.tile_shift == 5)      k_tile_accumulate<32><<<grid, 256, 0, st>>>(r->dev, r->tile_offsets, r->tile_idx2, ph & 1, ph >> 1);
    else if(r->dev.tile_shift == 4) k_tile_accumulate<16><<<grid, 256, 0, st>>>(r->dev, r->tile_offsets, r->tile_idx2, ph & 1, ph >> 1);
    else                            k_tile_accumulate<8><<<grid, 256, 0, st>>>(r->dev, r->tile_offsets, r->tile_idx2, ph & 1, ph >> 1);
  }
  // unused slots carry the key one past the last tile, so that a sort over an upper bound of the count leaves them at the end
  k_fill_u32<<<cb200_sm_count_cached()*4, 256, 0, st>>>(r->dev.tile_key, r->dev.tiles_x*r->dev.tiles_y, n);
  CB_CUDA(cudaMemsetAsync(r->dev.tile_count, 0, sizeof(unsigned int), st));
  cb200_count_launch(9); r->stats.kernel_launches += 9;   // radix sort (counted as 3), offsets, four phases, fill
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb200_render_set_accumulation(cb200_render_t *r, int mode)
{
  if(!r || (mode != CB200_ACCUM_ATOMIC && mode != CB200_ACCUM_TILES)) { cb200_set_error("render_set_accumulation: bad arguments"); return CB200_ERR_ARG; }
  CB_CUDA(cudaDeviceSynchronize());
  if(r->tile_ub) { const int rc = tile_pass(r, 0); if(rc) return rc; CB_CUDA(cudaDeviceSynchronize()); }
  if(mode == CB200_ACCUM_TILES && !r->tile_a)
  {
    const size_t cap = 4*(size_t)r->batch;
    r->tile_a = dev_alloc<float4>(r, cap); r->tile_b = dev_alloc<float>(r, cap);
    r->tile_key = dev_alloc<uint32_t>(r, cap); r->tile_idx = dev_alloc<uint32_t>(r, cap);
    r->tile_key2 = dev_alloc<uint32_t>(r, cap); r->tile_idx2 = dev_alloc<uint32_t>(r, cap);
    // the largest tile that still gives every SM two blocks per checkerboard phase (frames are multiples of 32 pixels)
    r->dev.tile_shift = 5;
    while(r->dev.tile_shift > 3 && (uint64_t)(r->dev.fb_w >> r->dev.tile_shift)*(r->dev.fb_h >> r->dev.tile_shift) < 8ull*cb200_sm_count_cached()) r->dev.tile_shift--;
    const uint32_t T = 1u << r->dev.tile_shift;
    r->dev.tiles_x = (r->dev.fb_w + T - 1)/T; r->dev.tiles_y = (r->dev.fb_h + T - 1)/T;
    r->tile_bits = 1;
    while((1u << r->tile_bits) <= r->dev.tiles_x*r->dev.tiles_y) r->tile_bits++;
    r->tile_offsets = dev_alloc<uint32_t>(r, (size_t)r->dev.tiles_x*r->dev.tiles_y + 2);
    bool ok = r->tile_a && r->tile_b && r->tile_key && r->tile_idx && r->tile_key2 && r->tile_idx2 && r->tile_offsets;
    if(ok)
    {
      cub::DeviceRadixSort::SortPairs(nullptr, r->tile_tmp_bytes, r->tile_key, r->tile_key2, r->tile_idx, r->tile_idx2, (int)cap, 0, r->tile_bits);
      r->tile_tmp = dev_alloc<uint8_t>(r, r->tile_tmp_bytes);
      ok = r->tile_tmp != nullptr;
    }
    if(!ok) { r->tile_a = nullptr; cb200_set_error("render_set_accumulation: out of device memory"); return CB200_ERR_NOMEM; }
    r->dev.tile_cap = (uint32_t)cap;
    k_fill_u32<<<cb200_sm_count_cached()*4, 256>>>(r->tile_key, r->dev.tiles_x*r->dev.tiles_y, (uint32_t)cap);
    CB_CUDA(cudaDeviceSynchronize());
  }
  r->tile_mode = mode == CB200_ACCUM_TILES;
  r->dev.tile_a = r->tile_mode ? r->tile_a : nullptr;
  r->dev.tile_b = r->tile_b; r->dev.tile_key = r->tile_key; r->dev.tile_idx = r->tile_idx;
  r->dev.tile_count = &r->d_cnt->tile_count;
  return 0;
}
extern "C" int cb200_render_accumulation(cb200_render_t *r) { return r && r->tile_mode ? CB200_ACCUM_TILES : CB200_ACCUM_ATOMIC; }

#define DEFER_MAX (1u << 18)   // parked next-event records that trigger their shadow sweep

// shadow sweep + splat of the parked next-event records
static int resolve_deferred(cb200_render *r, cudaStream_t st)
{
  const uint32_t n_nee = r->nee_deferred;
  r->nee_deferred = 0;
  if(!n_nee) return 0;
  int rc;
  {
    TimeScope ts(r, st, KC_SHADOW, n_nee);
    rc = cb200_launch_shadow(r->accel, r->nee_rays, r->nee_md, r->nee_light, r->nee_vis, n_nee, st);
  }
  if(rc) return rc;
  {
    TimeScope ts(r, st, KC_RESOLVE, n_nee);
    k_nee_resolve<<<(n_nee + RB - 1)/RB, RB, 0, st>>>(r->dev, n_nee, r->nee_recs, r->nee_vis, r->d_cnt);
  }
  cb200_count_launch(2); r->stats.kernel_launches += 2;
  r->stats.rays_shadow += n_nee;
  CB_CUDA(cudaGetLastError());
  return 0;
}

// One wave of the pool: trace every path's pending ray, shade, trace and resolve the next-event rays.
static int render_wave(cb200_render *r, uint32_t n, cudaStream_t st)
{
  const int cur = r->cur;
  int rc;
  {
    TimeScope ts(r, st, KC_CLOSEST, n);
    rc = cb200_launch_intersect(r->accel, r->rays[cur], r->maxd[cur], r->hits, n, st, r->counting ? r->d_trav_cnt : nullptr);
  }
  if(rc) return rc;
  r->stats.rays_closest += n; r->stats.kernel_launches++;
  CB_CUDA(cudaMemsetAsync(r->d_cnt, 0, 8*sizeof(unsigned long long), st));   // next, nee, hits[5], em (splats keeps counting)
  // a wave that cannot fill the parked records' room up (every path makes at most one) runs its shadow sweep at once
  if(r->nee_deferred && (uint64_t)r->nee_deferred + n > r->batch) { rc = resolve_deferred(r, st); if(rc) return rc; }
  const uint32_t nd = r->nee_deferred;
  {
    TimeScope ts(r, st, KC_SHADE, n);
    if(r->dev.sky != CB_SKY_BLACK)
    {
      k_sky_miss<<<(n + RB - 1)/RB, RB, 0, st>>>(r->dev, n, r->st[cur], r->hits, r->d_cnt, r->em_recs);
      cb200_count_launch(); r->stats.kernel_launches++;
    }
    const int single = r->dev.has_media ? -1 : (r->bsdf_kinds == 1) ? 0 : (r->bsdf_kinds == 2) ? 1 : (r->bsdf_kinds == 4) ? 2 : (r->bsdf_kinds == 8) ? 3 : -1;
    k_compact_hits<<<(n + 255)/256, 256, 0, st>>>(r->dev, r->hits, n, (uint32_t)r->batch, r->hit_list, r->d_cnt, single);
    cb200_count_launch(); r->stats.kernel_launches++;
#define SHADE_ARGS(K) (r->dev, n, r->st[cur], r->rays[cur], r->hits, r->st[cur^1], r->rays[cur^1], \
      r->nee_rays + nd, r->nee_md + nd, r->nee_light + nd, r->nee_recs + nd, r->d_cnt, r->hit_list + (size_t)K*r->batch, K, r->maxd[cur^1], r->em_recs)
#define SHADE_LAUNCH(K) do { if(r->dev.has_media) k_shade<(1 << K), true><<<(n + RB - 1)/RB, RB, 0, st>>>SHADE_ARGS(K); \
                             else k_shade<(1 << K), false><<<(n + RB - 1)/RB, RB, 0, st>>>SHADE_ARGS(K); \
                             cb200_count_launch(); r->stats.kernel_launches++; } while(0)
    if(r->bsdf_kinds & 1) SHADE_LAUNCH(0);
    if(r->bsdf_kinds & 2) SHADE_LAUNCH(1);
    if(r->bsdf_kinds & 4) SHADE_LAUNCH(2);
    if(r->bsdf_kinds & 8) SHADE_LAUNCH(3);
    if(r->dev.has_media)
    { // volume vertices of scattering media (kind 4)
      k_shade<16, true><<<(n + RB - 1)/RB, RB, 0, st>>>SHADE_ARGS(4);
      cb200_count_launch(); r->stats.kernel_launches++;
    }
#undef SHADE_LAUNCH
#undef SHADE_ARGS
  }
  CB_CUDA(cudaMemcpyAsync(r->h_cnt, r->d_cnt, sizeof(ShadeCounters), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaStreamSynchronize(st));
  const uint32_t n_next = (uint32_t)r->h_cnt->next, n_nee = (uint32_t)(r->h_cnt->next >> 32), n_em = (uint32_t)r->h_cnt->em;   // next: two 32-bit counts
  if(r->tile_mode) r->tile_ub = r->h_cnt->tile_count + nd + n_nee + n_em;   // exact up to this wave's shading + at most one record per queued (or parked) contribution
  if(n_em)
  { // emission found by extension (k_shade's queue): no visibility to wait for
    TimeScope ts(r, st, KC_RESOLVE, n_em);
    k_nee_resolve<<<(n_em + RB - 1)/RB, RB, 0, st>>>(r->dev, n_em, r->em_recs, nullptr, r->d_cnt);
    cb200_count_launch(); r->stats.kernel_launches++;
  }
  // Small waves (the tail of a flush: ~22 waves of fewer than 100 k rays) only PARK their next-event records; one shadow sweep at
  // the end of the call traces them together: a sweep of a few thousand rays costs the latency floor of its launch (~0.15 ms),
  // and nothing waits for its answers but the framebuffer.
  r->nee_deferred = nd + n_nee;
  if(r->nee_deferred >= DEFER_MAX) { rc = resolve_deferred(r, st); if(rc) return rc; }
  r->n_alive = n_next;
  r->cur = cur ^ 1;
  // the list holds 4 waves' worth and the next wave adds at most 2 * batch records (one per path that finds an emitter, one per
  // next-event ray): flush it only when that might not fit -- normally once per call, at the end (render_collect)
  if(r->tile_mode && (uint64_t)r->tile_ub + 2ull*r->batch > r->dev.tile_cap) { rc = tile_pass(r, st); if(rc) return rc; }
  return 0;
}

static int render_collect(cb200_render *r, cudaStream_t st)
{
  if(r->nee_deferred) { const int rc = resolve_deferred(r, st); if(rc) return rc; }
  if(r->tile_mode && r->tile_ub) { const int rc = tile_pass(r, st); if(rc) return rc; }
  CB_CUDA(cudaMemcpyAsync(r->h_cnt, r->d_cnt, sizeof(ShadeCounters), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaStreamSynchronize(st));
  r->stats.splats = r->h_cnt->splats;
  if(r->timing)
    for(size_t k=0;k+1<r->ev_used;k+=2)
    {
      float ms = 0.0f;
      if(cudaEventElapsedTime(&ms, r->ev_pool[k], r->ev_pool[k+1]) == cudaSuccess) r->stats.ms[r->ev_class[k]] += ms;
      if(getenv("CB200_RENDER_TRACE")) fprintf(stderr, "[cb200 trace] class %d n %u ms %.3f -> %.1f M/s\n", r->ev_class[k], r->ev_n[k], ms, r->ev_n[k]/(ms*1e3));
    }
  r->ev_used = 0;
  if(r->counting)
  {
    unsigned long long c[8];
    CB_CUDA(cudaMemcpy(c, r->d_trav_cnt, sizeof(c), cudaMemcpyDeviceToHost));
    for(int k=0;k<4;k++) { r->stats.trav_closest[k] = c[k]; r->stats.trav_shadow[k] = c[4+k]; }
  }
  CB_CUDA(cudaGetLastError());
  return 0;
}

int cb200_render_pass_stream(cb200_render_t *r, uint64_t first_index, uint64_t count, void *stream_)
{
  if(!r) { cb200_set_error("render_pass: null"); return CB200_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream_;
  const uint32_t N = (uint32_t)r->batch;
  uint64_t started = 0;
  while(started < count)
  {
    // top the pool up with new paths behind the survivors of the previous wave
    const uint32_t room = N - r->n_alive;
    const uint32_t n_new = (uint32_t)((count - started) < room ? (count - started) : room);
    if(n_new)
    {
      TimeScope ts(r, st, KC_START, n_new);
      k_pixel_keys<<<(n_new + RB - 1)/RB, RB, 0, st>>>(r->dev, first_index + started, n_new, r->keys[0], r->order[0]);
      size_t tmp = r->sort_tmp_bytes;
      // all 24 Morton bits: sorting only the upper 16 (16 x 16-pixel tiles left in index order inside, one radix pass less) costs the
      // closest-hit kernel 9 % (86.2 -> 93.9 ms per 8 progressions), the upper 12 bits 31 %, the upper 8 44 % (profiles/r3o)
      CB_CUDA(cub::DeviceRadixSort::SortPairs(r->sort_tmp, tmp, r->keys[0], r->keys[1], r->order[0], r->order[1], (int)n_new, 0, 24, st));
      k_path_start<<<(n_new + RB - 1)/RB, RB, 0, st>>>(r->dev, first_index + started, n_new, r->order[1],
                                                        r->st[r->cur] + r->n_alive, r->rays[r->cur] + r->n_alive, nullptr,
                                                        r->maxd[r->cur] ? r->maxd[r->cur] + r->n_alive : nullptr);
      cb200_count_launch(5); r->stats.kernel_launches += 5;   // keys, radix sort (histogram + digit passes, counted as 3), path start
      r->stats.paths += n_new;
      started += n_new;
    }
    const int rc = render_wave(r, r->n_alive + n_new, st);
    if(rc) return rc;
  }
  return render_collect(r, st);
}

int cb200_render_flush(cb200_render_t *r, void *stream_)
{
  if(!r) { cb200_set_error("render_flush: null"); return CB200_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream_;
  for(int guard=0; r->n_alive && guard<64; guard++)
  {
    const int rc = render_wave(r, r->n_alive, st);
    if(rc) return rc;
  }
  return render_collect(r, st);
}

int cb200_render_pass(cb200_render_t *r, uint64_t first_index, uint64_t count, void *stream)
{
  const int rc = cb200_render_pass_stream(r, first_index, count, stream);
  return rc ? rc : cb200_render_flush(r, stream);
}

int cb200_render_point(cb200_render_t *r, const uint64_t *index, const int32_t *dim, float *out, uint64_t n)
{
  if(!r || !index || !dim || !out) { cb200_set_error("render_point: bad arguments"); return CB200_ERR_ARG; }
  uint64_t *d_i = nullptr; int32_t *d_d = nullptr; float *d_o = nullptr;
  CB_CUDA(cudaMalloc(&d_i, n*8 + 8)); CB_CUDA(cudaMalloc(&d_d, n*4 + 4)); CB_CUDA(cudaMalloc(&d_o, n*4 + 4));
  CB_CUDA(cudaMemcpy(d_i, index, n*8, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d_d, dim, n*4, cudaMemcpyHostToDevice));
  if(n) k_points<<<(unsigned)((n + 127)/128), 128>>>(r->dev.points, d_i, d_d, d_o, (uint32_t)n);
  cb200_count_launch();
  CB_CUDA(cudaMemcpy(out, d_o, n*4, cudaMemcpyDeviceToHost));
  cudaFree(d_i); cudaFree(d_d); cudaFree(d_o);
  return 0;
}

int cb200_render_offset_ray(cb200_render_t *r, const float *x, const float *dir, float *out_pos, uint64_t n)
{
  if(!r || !x || !dir || !out_pos) { cb200_set_error("render_offset_ray: bad arguments"); return CB200_ERR_ARG; }
  float *d = nullptr;
  CB_CUDA(cudaMalloc(&d, (n + 1)*9*sizeof(float)));
  CB_CUDA(cudaMemcpy(d, x, n*3*sizeof(float), cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d + 3*n, dir, n*3*sizeof(float), cudaMemcpyHostToDevice));
  if(n) k_offset_ray<<<(unsigned)((n + 127)/128), 128>>>(d, d + 3*n, d + 6*n, (uint32_t)n);
  cb200_count_launch();
  CB_CUDA(cudaMemcpy(out_pos, d + 6*n, n*3*sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

int cb200_render_bsdf(cb200_render_t *r, int32_t material, const cb_bsdf_query_t *queries, cb_bsdf_result_t *results, uint64_t n)
{
  if(!r || !queries || !results || material < 0 || material >= r->desc.num_materials || !r->mat_ok[material])
  { cb200_set_error("render_bsdf: bad arguments"); return CB200_ERR_ARG; }
  cb_bsdf_query_t *d_q = nullptr; cb_bsdf_result_t *d_o = nullptr;
  CB_CUDA(cudaMalloc(&d_q, (n + 1)*sizeof(cb_bsdf_query_t))); CB_CUDA(cudaMalloc(&d_o, (n + 1)*sizeof(cb_bsdf_result_t)));
  CB_CUDA(cudaMemcpy(d_q, queries, n*sizeof(cb_bsdf_query_t), cudaMemcpyHostToDevice));
  if(n) k_bsdf<<<(unsigned)((n + 127)/128), 128>>>(r->dev.mats, material, d_q, d_o, (uint32_t)n);
  cb200_count_launch();
  CB_CUDA(cudaMemcpy(results, d_o, n*sizeof(cb_bsdf_result_t), cudaMemcpyDeviceToHost));
  cudaFree(d_q); cudaFree(d_o);
  return 0;
}

int cb200_render_medium(cb200_render_t *r, int32_t medium, const cb_medium_query_t *queries, cb_medium_result_t *results, uint64_t n)
{
  if(!r || !queries || !results || medium < 0 || medium >= r->desc.num_media)
  { cb200_set_error("render_medium: bad arguments"); return CB200_ERR_ARG; }
  cb_medium_query_t *d_q = nullptr; cb_medium_result_t *d_o = nullptr;
  CB_CUDA(cudaMalloc(&d_q, (n + 1)*sizeof(cb_medium_query_t))); CB_CUDA(cudaMalloc(&d_o, (n + 1)*sizeof(cb_medium_result_t)));
  CB_CUDA(cudaMemcpy(d_q, queries, n*sizeof(cb_medium_query_t), cudaMemcpyHostToDevice));
  if(n) k_medium<<<(unsigned)((n + 127)/128), 128>>>(r->dev.mats, (uint32_t)medium + 1u, d_q, d_o, (uint32_t)n);
  cb200_count_launch();
  CB_CUDA(cudaMemcpy(results, d_o, n*sizeof(cb_medium_result_t), cudaMemcpyDeviceToHost));
  cudaFree(d_q); cudaFree(d_o);
  return 0;
}

int cb200_render_camera_rays(cb200_render_t *r, uint64_t first_index, uint64_t n, cb_ray_t *out_rays, float *out_aux)
{
  if(!r || !out_rays || n > r->batch) { cb200_set_error("render_camera_rays: bad arguments (n must be <= batch_paths)"); return CB200_ERR_ARG; }
  float *d_aux = nullptr;
  CB_CUDA(cudaMalloc(&d_aux, n*16 + 16));
  if(n) k_path_start<<<(unsigned)((n + RB - 1)/RB), RB>>>(r->dev, first_index, (uint32_t)n, nullptr, nullptr, r->rays[0], d_aux, nullptr);
  cb200_count_launch();
  CB_CUDA(cudaMemcpy(out_rays, r->rays[0], n*sizeof(cb_ray_t), cudaMemcpyDeviceToHost));
  if(out_aux) CB_CUDA(cudaMemcpy(out_aux, d_aux, n*16, cudaMemcpyDeviceToHost));
  cudaFree(d_aux);
  return 0;
}

} // extern "C"

// The first wave of path indices [first_index, first_index + n) on an empty pool, for the known-answer entries below: camera sample,
// closest hit, vertex preparation + next-event sample + BSDF sample (k_shade) -- the same kernels as cb200_render_pass; what they
// queue is left in the wave buffers instead of being traced / splatted.  scramble > 0 presets path->tangent_frame_scrambling.
static int first_wave(cb200_render *r, uint64_t first_index, uint32_t m, float scramble, cudaStream_t st, int cur = -1)
{
  if(cur < 0)
  {
    cur = r->cur;
    r->dev.force_scramble = scramble;
    k_path_start<<<(m + RB - 1)/RB, RB, 0, st>>>(r->dev, first_index, m, nullptr, r->st[cur], r->rays[cur], nullptr, r->maxd[cur]);
    r->dev.force_scramble = 0.0f;
    cb200_count_launch();
  }   // else: the wave behind it -- the survivors the previous call left in the buffers of side `cur`
  int rc = cb200_launch_intersect(r->accel, r->rays[cur], r->maxd[cur], r->hits, m, st, nullptr);
  if(rc) return rc;
  CB_CUDA(cudaMemsetAsync(r->d_cnt, 0, 8*sizeof(unsigned long long), st));
  const int single = r->dev.has_media ? -1 : (r->bsdf_kinds == 1) ? 0 : (r->bsdf_kinds == 2) ? 1 : (r->bsdf_kinds == 4) ? 2 : (r->bsdf_kinds == 8) ? 3 : -1;
  if(r->dev.sky != CB_SKY_BLACK) { k_sky_miss<<<(m + RB - 1)/RB, RB, 0, st>>>(r->dev, m, r->st[cur], r->hits, r->d_cnt, r->em_recs); cb200_count_launch(); }
  k_compact_hits<<<(m + 255)/256, 256, 0, st>>>(r->dev, r->hits, m, (uint32_t)r->batch, r->hit_list, r->d_cnt, single);
#define NEE_SHADE(K, MED) k_shade<(1 << K), MED><<<(m + RB - 1)/RB, RB, 0, st>>>(r->dev, m, r->st[cur], r->rays[cur], r->hits, r->st[cur^1], r->rays[cur^1], \
      r->nee_rays, r->nee_md, r->nee_light, r->nee_recs, r->d_cnt, r->hit_list + (size_t)K*r->batch, K, r->maxd[cur^1], r->em_recs)
#define NEE_SHADE_K(K) do { if(r->dev.has_media) NEE_SHADE(K, true); else NEE_SHADE(K, false); cb200_count_launch(); } while(0)
  if(r->bsdf_kinds & 1) NEE_SHADE_K(0);
  if(r->bsdf_kinds & 2) NEE_SHADE_K(1);
  if(r->bsdf_kinds & 4) NEE_SHADE_K(2);
  if(r->bsdf_kinds & 8) NEE_SHADE_K(3);
  if(r->dev.has_media) { NEE_SHADE(4, true); cb200_count_launch(); }
#undef NEE_SHADE_K
#undef NEE_SHADE
  cb200_count_launch(2);
  CB_CUDA(cudaMemcpyAsync(r->h_cnt, r->d_cnt, sizeof(ShadeCounters), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaStreamSynchronize(st));
  return 0;   // the survivors and the queued emission of this wave are dropped by the callers: the pool stays empty, the framebuffer untouched
}

// Next-event samples of the FIRST hit vertex of path indices [first_index, first_index + n), for known-answer tests against the
// reference's own nee_sample (oracle/ref_path.c): out[k][16] = {pixel_i, pixel_j, lambda, value (throughput x mis weight), total_dist,
// light prim (2 words, bit pattern), ray pos[3], ray dir[3], search limit, visible (1 / 0), path length at the splat}; *n_out records.
int cb200_render_nee_records(cb200_render_t *r, uint64_t first_index, uint64_t n, float *out, uint64_t *n_out)
{
  if(!r || !out || !n_out || n == 0 || n > r->batch) { cb200_set_error("render_nee_records: bad arguments (0 < n <= batch_paths)"); return CB200_ERR_ARG; }
  if(r->n_alive || r->nee_deferred) { cb200_set_error("render_nee_records: paths in flight (flush first)"); return CB200_ERR_ARG; }
  cudaStream_t st = 0;
  int rc = first_wave(r, first_index, (uint32_t)n, 0.0f, st);
  if(rc) return rc;
  const uint32_t n_nee = (uint32_t)(r->h_cnt->next >> 32);
  *n_out = n_nee;
  if(!n_nee) return 0;
  rc = cb200_launch_shadow(r->accel, r->nee_rays, r->nee_md, r->nee_light, r->nee_vis, n_nee, st);
  if(rc) return rc;
  std::vector<cb_ray_t> rays(n_nee);
  std::vector<float> md(n_nee);
  std::vector<NeeRec> recs(n_nee);
  std::vector<int32_t> vis(n_nee);
  CB_CUDA(cudaMemcpy(rays.data(), r->nee_rays, n_nee*sizeof(cb_ray_t), cudaMemcpyDeviceToHost));
  CB_CUDA(cudaMemcpy(md.data(), r->nee_md, n_nee*sizeof(float), cudaMemcpyDeviceToHost));
  CB_CUDA(cudaMemcpy(recs.data(), r->nee_recs, n_nee*sizeof(NeeRec), cudaMemcpyDeviceToHost));
  CB_CUDA(cudaMemcpy(vis.data(), r->nee_vis, n_nee*sizeof(int32_t), cudaMemcpyDeviceToHost));
  for(uint32_t k=0;k<n_nee;k++)
  {
    float *o = out + 16*(size_t)k;
    o[0] = recs[k].pixel_i; o[1] = recs[k].pixel_j; o[2] = recs[k].lambda; o[3] = recs[k].value; o[4] = recs[k].total_dist;
    memcpy(o + 5, &recs[k].light_lo, 4); memcpy(o + 6, &recs[k].light_hi, 4);
    for(int c=0;c<3;c++) { o[7+c] = rays[k].pos[c]; o[10+c] = rays[k].dir[c]; }
    o[13] = md[k]; o[14] = (float)vis[k]; o[15] = (float)recs[k].len;
  }
  return 0;
}

// The paths that go on after their first hit vertex, as the first wave leaves them: the BSDF-sampled direction and everything
// path_extend records for the next edge, plus the ray the next wave will trace.  out[k][16] = {pixel_i, pixel_j, lambda, omega[3]
// (e[2].omega), ray pos[3] (prims_offset_ray applied), throughput into the next vertex, throughput into this one, bsdf pdf (projected
// solid angle), |n . omega|, vertex position x[3]}; *n_out records.  tangent_frame_scrambling > 0 presets the path's scrambling
// number (upstream draws it from the worker's twister) so that the reference harness can be run with the same one.
int cb200_render_bounce_records(cb200_render_t *r, uint64_t first_index, uint64_t n, float tangent_frame_scrambling, float *out, uint64_t *n_out)
{
  if(!r || !out || !n_out || n == 0 || n > r->batch) { cb200_set_error("render_bounce_records: bad arguments (0 < n <= batch_paths)"); return CB200_ERR_ARG; }
  if(r->n_alive || r->nee_deferred) { cb200_set_error("render_bounce_records: paths in flight (flush first)"); return CB200_ERR_ARG; }
  cudaStream_t st = 0;
  const int rc = first_wave(r, first_index, (uint32_t)n, tangent_frame_scrambling, st);
  if(rc) return rc;
  const uint32_t n_next = (uint32_t)r->h_cnt->next;
  *n_out = n_next;
  if(!n_next) return 0;
  std::vector<PathState> stv(n_next);
  std::vector<cb_ray_t> rays(n_next);
  CB_CUDA(cudaMemcpy(stv.data(), r->st[r->cur^1], n_next*sizeof(PathState), cudaMemcpyDeviceToHost));
  CB_CUDA(cudaMemcpy(rays.data(), r->rays[r->cur^1], n_next*sizeof(cb_ray_t), cudaMemcpyDeviceToHost));
  for(uint32_t k=0;k<n_next;k++)
  {
    float *o = out + 16*(size_t)k;
    const PathState &s = stv[k];
    o[0] = s.pixel_i; o[1] = s.pixel_j; o[2] = s.lambda;
    for(int c=0;c<3;c++) { o[3+c] = s.omega[c]; o[6+c] = rays[k].pos[c]; o[13+c] = s.x[c]; }
    o[9] = s.thr; o[10] = s.thr_prev; o[11] = s.pdf_proj; o[12] = s.cos_prev;
  }
  return 0;
}

// Emission found by EXTENSION, as the sampler would splat it: wave == 1: emitters (and nothing else) the camera sees directly
// (weight 1); wave == 2: emitters the first BSDF-sampled edge ends on, weighted against next-event estimation (ptdl.c:124-131,
// sampler_mis :78-88; pt.c: weight 1).  out[k][8] = {pixel_i, pixel_j, lambda, value (throughput x emission x mis weight), path
// length at the splat, 0, 0, 0}; *n_out records.  Same preconditions and scrambling preset as cb200_render_bounce_records.
int cb200_render_emission_records(cb200_render_t *r, uint64_t first_index, uint64_t n, float tangent_frame_scrambling, int32_t wave,
                                  float *out, uint64_t *n_out)
{
  if(!r || !out || !n_out || n == 0 || n > r->batch || wave < 1 || wave > 2) { cb200_set_error("render_emission_records: bad arguments (0 < n <= batch_paths, wave 1 or 2)"); return CB200_ERR_ARG; }
  if(r->n_alive || r->nee_deferred) { cb200_set_error("render_emission_records: paths in flight (flush first)"); return CB200_ERR_ARG; }
  cudaStream_t st = 0;
  int rc = first_wave(r, first_index, (uint32_t)n, tangent_frame_scrambling, st);
  if(rc) return rc;
  if(wave == 2)
  {
    const uint32_t n_next = (uint32_t)r->h_cnt->next;
    *n_out = 0;
    if(!n_next) return 0;
    rc = first_wave(r, 0, n_next, 0.0f, st, r->cur ^ 1);
    if(rc) return rc;
  }
  const uint32_t n_em = (uint32_t)r->h_cnt->em;
  *n_out = n_em;
  if(!n_em) return 0;
  std::vector<NeeRec> recs(n_em);
  CB_CUDA(cudaMemcpy(recs.data(), r->em_recs, n_em*sizeof(NeeRec), cudaMemcpyDeviceToHost));
  for(uint32_t k=0;k<n_em;k++)
  {
    float *o = out + 8*(size_t)k;
    o[0] = recs[k].pixel_i; o[1] = recs[k].pixel_j; o[2] = recs[k].lambda; o[3] = recs[k].value; o[4] = (float)recs[k].len;
    o[5] = o[6] = o[7] = 0.0f;
  }
  return 0;
}
