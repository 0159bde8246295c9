// traverse_common.cuh -- pieces shared by the 4-wide kernels (traverse.cu) and the 8-wide compressed ones (traverse8.cu)
#pragma once
#include "prims.cuh"
#include <type_traits>

#define TRACE_BLOCK 128
#ifndef TRACE_MIN_BLOCKS
#define TRACE_MIN_BLOCKS 8   // static triangle scenes: cap at 64 registers -> 32 resident warps per SM
#endif
#define FULL 0xffffffffu
#define STACK_SMALL 64    // stack entries of the common 4-wide kernels (3 per level + 4)
#define STACK_BIG   304   // the reference's MAX_TREE_DEPTH worth of entries
#define STACK_EXACT STACK_SMALL

__device__ __forceinline__ float sse_min(float a, float b) { return a < b ? a : b; }   // _mm_min_ps
__device__ __forceinline__ float sse_max(float a, float b) { return a > b ? a : b; }   // _mm_max_ps

// (a0 - p)*inv and (a1 - p)*inv as two packed IEEE operations (sub.rn.f32x2 / mul.rn.f32x2 -> FADD2 / FMUL2, sm_100): each half is
// rounded exactly like the scalar sub and mul of the reference's slab test (qbvhmp.c:1219-1221), at half the issue slots
__device__ __forceinline__ void slab2(float a0, float a1, float p, float inv, float &t0, float &t1)
{
  asm("{ .reg .b64 v, q, w;\n"
      "mov.b64 v, {%2, %3};\n"
      "mov.b64 q, {%4, %4};\n"
      "mov.b64 w, {%5, %5};\n"
      "sub.rn.f32x2 v, v, q;\n"
      "mul.rn.f32x2 v, v, w;\n"
      "mov.b64 {%0, %1}, v; }" : "=f"(t0), "=f"(t1) : "f"(a0), "f"(a1), "f"(p), "f"(inv));
}
// ((q0*t0 + w0*t1) - p)*inv for two planes at once: the time interpolation of aabb_intersect (qbvhmp.c:1206-1215) followed by the slab
// distances, every operation a packed IEEE one with the scalar form's rounding
__device__ __forceinline__ void lerp_slab2(float q0, float q1, float w0, float w1, float t0, float t1, float p, float inv, float &o0, float &o1)
{
  asm("{ .reg .b64 a, b, c;\n"
      "mov.b64 a, {%2, %3};\n"
      "mov.b64 c, {%6, %6};\n"
      "mul.rn.f32x2 a, a, c;\n"
      "mov.b64 b, {%4, %5};\n"
      "mov.b64 c, {%7, %7};\n"
      "mul.rn.f32x2 b, b, c;\n"
      "add.rn.f32x2 a, a, b;\n"
      "mov.b64 c, {%8, %8};\n"
      "sub.rn.f32x2 a, a, c;\n"
      "mov.b64 c, {%9, %9};\n"
      "mul.rn.f32x2 a, a, c;\n"
      "mov.b64 {%0, %1}, a; }" : "=f"(o0), "=f"(o1) : "f"(q0), "f"(q1), "f"(w0), "f"(w1), "f"(t0), "f"(t1), "f"(p), "f"(inv));
}
// three-input min / max (FMNMX3, sm_100): same result as the nested two-input forms for NaN-free operands
__device__ __forceinline__ float max3f(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min3f(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

struct NodeOut
{
  float tmin[4];
  bool hit[4];
  uint64_t child[4];
  int axis0, axis00, axis01;
};

// EXACT: reference select semantics; otherwise fminf/fmaxf (only valid when no NaN can occur)
template<bool MB, bool EXACT>
__device__ __forceinline__ void node_slabs(const void *nodes, uint64_t idx, float px, float py, float pz,
                                           float ix, float iy, float iz, float t0, float t1, float tmax_init, NodeOut &o,
                                           float tmin_init = 0.0f)
{
  float tmin[4] = {tmin_init, tmin_init, tmin_init, tmin_init};
  float tmax[4] = {tmax_init, tmax_init, tmax_init, tmax_init};
  const float pos[3] = {px, py, pz};
  const float inv[3] = {ix, iy, iz};
  const float4 *a0;
  const ulonglong2 *ch;
  if(MB)
  {
    const Node256 *n = reinterpret_cast<const Node256 *>(nodes) + idx;
    a0 = reinterpret_cast<const float4 *>(n->aabb0);
    ch = reinterpret_cast<const ulonglong2 *>(n->child);
  }
  else
  {
    const Node128 *n = reinterpret_cast<const Node128 *>(nodes) + idx;
    a0 = reinterpret_cast<const float4 *>(n->aabb0);
    ch = reinterpret_cast<const ulonglong2 *>(n->child);
  }
#pragma unroll
  for(int k=0;k<3;k++)
  {
    const float4 m0 = __ldg(a0 + k), M0 = __ldg(a0 + k + 3);
    float mo[4] = {m0.x, m0.y, m0.z, m0.w}, Mo[4] = {M0.x, M0.y, M0.z, M0.w};
    if(MB)
    {
      const float4 m1 = __ldg(a0 + 6 + k), M1 = __ldg(a0 + 6 + k + 3);
      const float mc[4] = {m1.x, m1.y, m1.z, m1.w}, Mc[4] = {M1.x, M1.y, M1.z, M1.w};
#pragma unroll
      for(int c=0;c<4;c++) { mo[c] = mo[c]*t0 + mc[c]*t1; Mo[c] = Mo[c]*t0 + Mc[c]*t1; }
    }
#pragma unroll
    for(int c=0;c<4;c++)
    {
      const float lo = (mo[c] - pos[k]) * inv[k];
      const float hi = (Mo[c] - pos[k]) * inv[k];
      if(EXACT)
      {
        tmin[c] = sse_max(tmin[c], sse_min(lo, hi));
        tmax[c] = sse_min(tmax[c], sse_max(lo, hi));
      }
      else
      {
        tmin[c] = fmaxf(tmin[c], fminf(lo, hi));
        tmax[c] = fminf(tmax[c], fmaxf(lo, hi));
      }
    }
  }
  const ulonglong2 c01 = __ldg(ch), c23 = __ldg(ch + 1);
  if(MB)
  {
    const ulonglong2 pa = __ldg(ch + 2), aa = __ldg(ch + 3);   // {parent, axis0}, {axis00, axis01}
    o.child[0] = c01.x;
    o.axis0 = (int)pa.y; o.axis00 = (int)aa.x; o.axis01 = (int)aa.y;
  }
  else
  {
    const uint32_t ax = (uint32_t)(c01.x >> CB_AXIS_SHIFT) & 63u;
    o.child[0] = c01.x & CB_CHILD_MASK;
    o.axis0 = ax & 3; o.axis00 = (ax >> 2) & 3; o.axis01 = (ax >> 4) & 3;
  }
  o.child[1] = c01.y; o.child[2] = c23.x; o.child[3] = c23.y;
#pragma unroll
  for(int c=0;c<4;c++) { o.tmin[c] = tmin[c]; o.hit[c] = tmin[c] <= tmax[c]; }
}

__device__ __forceinline__ void load_ray(const cb_ray_t *rays, uint32_t i, RayD &r)
{
  const float2 *p = reinterpret_cast<const float2 *>(rays + i);   // 40-byte records are 8-byte aligned
  // (evict-first loads of the ray stream and stores of the hit stream, ld.global.cs / st.global.cs, were measured: closest hit 95.8 ->
  // 95.4 ms per 8 progressions, within noise; profiles/r3d)
  const float2 a = __ldg(p), b = __ldg(p+1), c = __ldg(p+2), d = __ldg(p+3), e = __ldg(p+4);
  r.px = a.x; r.py = a.y; r.pz = b.x; r.dx = b.y; r.dy = c.x; r.dz = c.y;
  r.time = d.x; r.min_dist = d.y;
  r.ign_lo = __float_as_uint(e.x); r.ign_hi = __float_as_uint(e.y);
}

__device__ __forceinline__ bool finite_nonzero(float x) { const float a = fabsf(x); return a > 0.0f && a < __int_as_float(0x7f800000); }
__device__ __forceinline__ bool finite(float x) { return fabsf(x) < __int_as_float(0x7f800000); }

// a leaf reference with count 0 (the reference builder emits them with a non-zero begin, qbvhmp.c:989)
__device__ __forceinline__ bool is_empty_leaf(uint64_t c) { return (c & CB_LEAF_BIT) && !(c & 31ull); }

#define KEY_MISS __int_as_float(0x7fc00000)
#define KEY_HIT(k) ((k) == (k))
enum { ST_IDLE = 0, ST_NODE = 1, ST_PRIM = 2 };
