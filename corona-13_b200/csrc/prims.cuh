// prims.cuh -- ray/primitive tests inlined into traversal (device side).
//
// Same arithmetic, same operation order as the reference so that prim id, u, v and dist are
// bit-exact (compile with -fmad=false; IEEE div/sqrt are nvcc defaults):
//   triangle / quad : include/geo/triangle.h:263-343, src/prims.c:638-701
//   motion blur     : include/geo.h:120-138  ((1-time)*open + time*close, mul mul add)
//   sphere          : include/geo/sphere.h:112-180
//   cylinder / cone : include/geo/line.h:362-592 (mixed float/double where the C source promotes)
// atan2f/acosf (sphere / line u,v only) come from CUDA's libm and may differ from glibc by ulps.
#pragma once
#include "internal.h"
#include <float.h>

struct RayD
{
  float px, py, pz;
  float dx, dy, dz;
  float time, min_dist;
  uint32_t ign_lo, ign_hi;
};

struct HitD
{
  float dist, u, v;
  uint32_t prim_lo, prim_hi;
};

#define CBD __device__ __forceinline__

CBD float dot3(float ax, float ay, float az, float bx, float by, float bz) { return (ax*bx + ay*by) + az*bz; }

struct V3 { float x, y, z; };
CBD V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CBD float dot(V3 a, V3 b) { return (a.x*b.x + a.y*b.y) + a.z*b.z; }
CBD V3 cross(V3 a, V3 b)   // corona_common.h:161-164
{
  return mk3(a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y);
}
CBD V3 sub(V3 a, V3 b) { return mk3(a.x-b.x, a.y-b.y, a.z-b.z); }
CBD V3 normalise(V3 f)     // corona_common.h:173-177
{
  const float len = 1.0f/sqrtf(dot(f, f));
  return mk3(f.x*len, f.y*len, f.z*len);
}
CBD void onb(V3 n, V3 &u, V3 &v)   // get_onb, corona_common.h:179-200
{
  if(fabsf(n.y) < 0.5f) u = cross(n, mk3(0.0f, 1.0f, 0.0f));
  else                  u = cross(n, mk3(1.0f, 0.0f, 0.0f));
  u = normalise(u);
  v = cross(n, u);
}

// closest-hit triangle test; returns 1 and updates hit on acceptance (dist > min_dist && dist <= hit.dist)
CBD int tri_intersect(V3 v0, V3 v1, V3 v2, uint32_t id_lo, uint32_t id_hi, const RayD &r, HitD &h)
{
  if(id_lo == r.ign_lo && id_hi == r.ign_hi) return 0;
  const V3 e1 = sub(v1, v0), e2 = sub(v2, v0);
  const V3 dir = mk3(r.dx, r.dy, r.dz);
  const V3 pvec = cross(dir, e2);
  const float det = dot(e1, pvec);
  const float inv_det = 1.0f / det;
  const V3 tvec = mk3(r.px - v0.x, r.py - v0.y, r.pz - v0.z);
  const float v = dot(tvec, pvec) * inv_det;
  if(v < 0.0f || v > 1.0f) return 0;
  const V3 qvec = cross(tvec, e1);
  const float u = dot(dir, qvec) * inv_det;
  if(u < 0.0f || u + v > 1.0f) return 0;
  const float dist = dot(e2, qvec) * inv_det;
  if(dist > r.min_dist && dist <= h.dist)
  {
    h.dist = dist; h.prim_lo = id_lo; h.prim_hi = id_hi; h.u = u; h.v = v;
    return 1;
  }
  return 0;
}

// any-hit triangle test: no ignore, accepts 0 < dist <= max_dist
CBD int tri_visible(V3 v0, V3 v1, V3 v2, const RayD &r, float max_dist)
{
  const V3 e1 = sub(v1, v0), e2 = sub(v2, v0);
  const V3 dir = mk3(r.dx, r.dy, r.dz);
  const V3 pvec = cross(dir, e2);
  const float det = dot(e1, pvec);
  const float inv_det = 1.0f / det;   // (float)(1.0/(double)det) rounds to the same value
  const V3 tvec = mk3(r.px - v0.x, r.py - v0.y, r.pz - v0.z);
  const float v = dot(tvec, pvec) * inv_det;
  if(v < 0.0f || v > 1.0f) return 0;
  const V3 qvec = cross(tvec, e1);
  const float u = dot(dir, qvec) * inv_det;
  if(u < 0.0f || u + v > 1.0f) return 0;
  const float dist = dot(e2, qvec) * inv_det;
  return (dist > 0.0f && dist <= max_dist) ? 1 : 0;
}

CBD float sphere_t(V3 c, float radius, const RayD &r)
{
  const V3 dir = mk3(r.dx, r.dy, r.dz);
  const float a = dot(dir, dir);
  const V3 o = mk3(r.px - c.x, r.py - c.y, r.pz - c.z);
  const float b = 2.0f*dot(o, dir);
  const float cc = dot(o, o) - radius*radius;
  if(a == 0.0f)
  {
    if(b != 0.0f) return -cc / b;
    return -FLT_MAX;
  }
  const float discrim = b*b - 4.0f*a*cc;
  if(discrim < 0.0f) return -FLT_MAX;
  float temp;
  const float sq = sqrtf(discrim);
  if(b < 0.0f) temp = -0.5f * (b - sq);
  else         temp = -0.5f * (b + sq);
  const float x0 = temp / a;
  const float x1 = cc / temp;
  if(x0 <= 0.0f) return x1;
  else if(x1 <= 0.0f) return x0;
  else return fminf(x0, x1);
}

CBD float cylinder_t(V3 v0, V3 v1, float rad, const RayD &r, float out[3], float &len)
{
  V3 d = sub(v1, v0);
  const V3 dir = mk3(r.dx, r.dy, r.dz);
  if((double)rad < 0.01)
  {
    const V3 a = cross(d, dir);
    const V3 o = mk3(v0.x - r.px, v0.y - r.py, v0.z - r.pz);
    const float dotp = dot(o, a);
    const float ilen = (float)(1.0/(double)sqrtf(dot(a, a)));
    const float dist = fabsf(dotp*ilen);
    if(dist > rad) return -1.0f;
    const float dlen = sqrtf(dot(d, d));
    len = dlen;
    const V3 b = cross(d, a);
    const float t = dot(o, b)/dot(b, dir);
    const V3 w = mk3(t*dir.x - o.x, t*dir.y - o.y, t*dir.z - o.z);
    out[0] = dot(w, d)/dlen;
    out[1] = 0.0f;
    out[2] = 1.0f;
    if(out[0] >= 0.0f && out[0] <= dlen) return t;
    return -1.0f;
  }
  const float dlen = sqrtf(dot(d, d));
  len = dlen;
  const float il = 1.0f/dlen;
  d = mk3(d.x*il, d.y*il, d.z*il);
  V3 a, b;
  onb(d, a, b);
  const float rx = r.px - v0.x, ry = r.py - v0.y, rz = r.pz - v0.z;
  float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
  o0 += rx*d.x; o1 += rx*a.x; o2 += rx*b.x; w0 += r.dx*d.x; w1 += r.dx*a.x; w2 += r.dx*b.x;
  o0 += ry*d.y; o1 += ry*a.y; o2 += ry*b.y; w0 += r.dy*d.y; w1 += r.dy*a.y; w2 += r.dy*b.y;
  o0 += rz*d.z; o1 += rz*a.z; o2 += rz*b.z; w0 += r.dz*d.z; w1 += r.dz*a.z; w2 += r.dz*b.z;
  const float A = w1*w1 + w2*w2;
  const float B = 2.0f*(o1*w1 + o2*w2);
  const float C = o1*o1 + o2*o2 - rad*rad;
  const float discr = (float)__dsub_rn((double)(B*B), __dmul_rn(__dmul_rn(4.0, (double)A), (double)C));
  if(discr < 0.0f) return -1.0f;
  float temp;
  const float sq = sqrtf(discr);
  if(B < 0.0f) temp = -0.5f * (B - sq);
  else         temp = -0.5f * (B + sq);
  const float t0 = temp / A;
  const float t1 = C / temp;
  float t;
  if(t0 <= 0.0f) t = t1;
  else if(t1 <= 0.0f) t = t0;
  else
  {
    t = fminf(t0, t1);
    for(int i=0;i<2;i++)
    {
      out[0] = o0 + t*w0; out[1] = o1 + t*w1; out[2] = o2 + t*w2;
      if(out[0] >= 0.0f && out[0] <= dlen) return t;
      t = fmaxf(t0, t1);
    }
    return -1.0f;
  }
  out[0] = o0 + t*w0; out[1] = o1 + t*w1; out[2] = o2 + t*w2;
  if(out[0] >= 0.0f && out[0] <= dlen) return t;
  return -1.0f;
}

// hit == nullptr: shadow variant (ray direction may be unnormalised)
CBD float cone_t(V3 v0, V3 v1, float r0, float r1, const RayD &r, float dist, HitD *hit)
{
  const V3 dir = mk3(r.dx, r.dy, r.dz);
  float iraylen = 1.0f;
  if(!hit) iraylen = 1.0f/sqrtf(dot(dir, dir));
  V3 d = sub(v1, v0);
  const float d_len = sqrtf(dot(d, d));
  const double idl = 1.0/(double)d_len;
  d = mk3((float)__dmul_rn((double)d.x, idl), (float)__dmul_rn((double)d.y, idl), (float)__dmul_rn((double)d.z, idl));
  const float cos_dr = dot(d, dir)*iraylen;
  const float cos_a2 = d_len*d_len/((r1-r0)*(r1-r0) + d_len*d_len);
  const float tt = -r0*d_len/(r1-r0);
  const V3 tip = mk3(v0.x + tt*d.x, v0.y + tt*d.y, v0.z + tt*d.z);
  const V3 o = mk3(r.px - tip.x, r.py - tip.y, r.pz - tip.z);
  const float cos_do = dot(d, o);
  const float cos_ro = dot(dir, o)*iraylen;
  const float cos_oo = dot(o, o);
  const float c2 = cos_dr*cos_dr - cos_a2;
  const float c1 = cos_dr*cos_do - cos_a2*cos_ro;
  const float c0 = cos_do*cos_do - cos_a2*cos_oo;
  float tmin = -1.0f;
  if(fabsf(c2) > 0.0f)
  {
    const float discr = c1*c1 - c0*c2;
    if(discr < 0.0f) return -1.0f;
    const float root = sqrtf(discr);
    for(int i=-1;i<2;i+=2)
    {
      const float t = (-c1 + (float)i*root)/c2;
      if(t > 0.0f && t < dist/iraylen)
      {
        const V3 x = mk3(r.px + t*r.dx*iraylen - v0.x, r.py + t*r.dy*iraylen - v0.y, r.pz + t*r.dz*iraylen - v0.z);
        const float dt = dot(x, d);
        if(dt >= 0.0f && dt <= d_len)
        {
          if(hit)
          {
            hit->u = dt/d_len;
            V3 a, b;
            onb(d, a, b);
            hit->v = (float)((double)atan2f(dot(a, x), dot(b, x))/(2.0*3.14159265358979323846));
          }
          tmin = dist = t;
        }
      }
    }
  }
  return tmin;
}

// ---------------------------------------------------------------------------------------------
// record access.  rec points at the primitive's first 64-byte unit.
// ---------------------------------------------------------------------------------------------
CBD V3 lerp_v(const float4 &o, const float4 &c, float t0, float t1)
{
  return mk3(t0*o.x + t1*c.x, t0*o.y + t1*c.y, t0*o.z + t1*c.z);
}

// closest hit against one primitive record (prims_intersect, src/prims.c:638-672)
template<bool ANALYTIC = true>
CBD void prim_intersect(const float4 *__restrict__ rec, uint32_t rec_units, const RayD &r, HitD &h)
{
  const float4 r0 = __ldg(rec + 0);
  const float4 r1 = __ldg(rec + 1);
  const uint32_t id_lo = __float_as_uint(r0.w), id_hi = __float_as_uint(r1.w);
  const uint32_t vcnt = id_hi >> 29;
  const bool mb = (id_hi >> 28) & 1u;
  const float t1 = r.time, t0 = 1.0f - r.time;
  if(vcnt == CB_PRIM_TRI || vcnt == CB_PRIM_QUAD)
  {
    const float4 r2 = __ldg(rec + 2);
    V3 v0, v1, v2;
    if(mb)
    {
      v0 = lerp_v(r0, __ldg(rec + 4), t0, t1);
      v1 = lerp_v(r1, __ldg(rec + 5), t0, t1);
      v2 = lerp_v(r2, __ldg(rec + 6), t0, t1);
    }
    else { v0 = mk3(r0.x, r0.y, r0.z); v1 = mk3(r1.x, r1.y, r1.z); v2 = mk3(r2.x, r2.y, r2.z); }
    if(vcnt == CB_PRIM_TRI) tri_intersect(v0, v1, v2, id_lo, id_hi, r, h);
    else
    {
      if(tri_intersect(v0, v1, v2, id_lo, id_hi, r, h)) { h.v += h.u; return; }
      const float4 r3 = __ldg(rec + 3);
      V3 v3;
      if(mb) v3 = lerp_v(r3, __ldg(rec + 7), t0, t1);
      else   v3 = mk3(r3.x, r3.y, r3.z);
      if(tri_intersect(v0, v2, v3, id_lo, id_hi, r, h)) h.u += h.v;
    }
  }
  else if(ANALYTIC && vcnt == CB_PRIM_SPHERE)
  {
    const float4 r2 = __ldg(rec + 2);
    const float radius = r2.w;
    V3 c;
    if(mb) c = lerp_v(r0, __ldg(rec + 4), t0, t1);
    else   c = mk3(r0.x, r0.y, r0.z);
    const float t = sphere_t(c, radius, r);
    if(t > r.min_dist && t < h.dist)
    {
      h.dist = t; h.prim_lo = id_lo; h.prim_hi = id_hi;
      const float x = r.px + t*r.dx, y = r.py + t*r.dy, z = r.pz + t*r.dz;
      h.u = (float)((double)atan2f((y - c.y)/radius, (x - c.x)/radius)/(2.0*3.14159265358979323846));
      const float cz = (z - c.z)/radius;
      const float cl = cz > -1.0f ? cz : -1.0f;
      h.v = (float)((double)acosf(cl < 1.0f ? cl : 1.0f)/3.14159265358979323846);
    }
  }
  else if(ANALYTIC && vcnt == CB_PRIM_LINE)
  {
    const float4 r2 = __ldg(rec + 2);
    const float4 r3 = __ldg(rec + 3);
    const float rad0 = r2.w, rad1 = r3.w;
    const bool linestrip = (rad0 > rad1 ? rad0 : rad1) <= 1e-2f;
    if(linestrip && id_lo == r.ign_lo && id_hi == r.ign_hi) return;
    V3 v0, v1;
    if(mb) { v0 = lerp_v(r0, __ldg(rec + 4), t0, t1); v1 = lerp_v(r1, __ldg(rec + 5), t0, t1); }
    else   { v0 = mk3(r0.x, r0.y, r0.z); v1 = mk3(r1.x, r1.y, r1.z); }
    if((double)fabsf(rad1 - rad0) < 1e-3)
    {
      float out[3], len;
      const float t = cylinder_t(v0, v1, rad0, r, out, len);
      if(t > r.min_dist && t < h.dist)
      {
        h.dist = t; h.prim_lo = id_lo; h.prim_hi = id_hi;
        h.u = out[0]/len;
        h.v = (float)((double)atan2f(out[1], out[2])/(2.0*3.14159265358979323846));
      }
    }
    else
    {
      const float t = cone_t(v0, v1, rad0, rad1, r, h.dist, &h);
      const float lim = r.min_dist > 1e-3f ? r.min_dist : 1e-3f;
      if((linestrip && t > lim) || (!linestrip && t > r.min_dist))
      {
        h.dist = t; h.prim_lo = id_lo; h.prim_hi = id_hi;
      }
    }
  }
}

// any hit against one primitive record (prims_intersect_visible, src/prims.c:674-701)
template<bool ANALYTIC = true>
CBD int prim_visible(const float4 *__restrict__ rec, uint32_t rec_units, const RayD &r, float max_dist)
{
  const float4 r0 = __ldg(rec + 0);
  const float4 r1 = __ldg(rec + 1);
  const uint32_t id_lo = __float_as_uint(r0.w), id_hi = __float_as_uint(r1.w);
  const uint32_t vcnt = id_hi >> 29;
  const bool mb = (id_hi >> 28) & 1u;
  const float t1 = r.time, t0 = 1.0f - r.time;
  if(vcnt == CB_PRIM_TRI || vcnt == CB_PRIM_QUAD)
  {
    const float4 r2 = __ldg(rec + 2);
    V3 v0, v1, v2;
    if(mb)
    {
      v0 = lerp_v(r0, __ldg(rec + 4), t0, t1);
      v1 = lerp_v(r1, __ldg(rec + 5), t0, t1);
      v2 = lerp_v(r2, __ldg(rec + 6), t0, t1);
    }
    else { v0 = mk3(r0.x, r0.y, r0.z); v1 = mk3(r1.x, r1.y, r1.z); v2 = mk3(r2.x, r2.y, r2.z); }
    if(vcnt == CB_PRIM_TRI) return tri_visible(v0, v1, v2, r, max_dist);
    if(tri_visible(v0, v1, v2, r, max_dist)) return 1;
    const float4 r3 = __ldg(rec + 3);
    V3 v3;
    if(mb) v3 = lerp_v(r3, __ldg(rec + 7), t0, t1);
    else   v3 = mk3(r3.x, r3.y, r3.z);
    return tri_visible(v0, v2, v3, r, max_dist);
  }
  else if(ANALYTIC && vcnt == CB_PRIM_SPHERE)
  {
    const float4 r2 = __ldg(rec + 2);
    V3 c;
    if(mb) c = lerp_v(r0, __ldg(rec + 4), t0, t1);
    else   c = mk3(r0.x, r0.y, r0.z);
    const float t = sphere_t(c, r2.w, r);
    return (t > 0.0f && t <= max_dist) ? 1 : 0;
  }
  else if(ANALYTIC && vcnt == CB_PRIM_LINE)
  {
    const float4 r2 = __ldg(rec + 2);
    const float4 r3 = __ldg(rec + 3);
    const float rad0 = r2.w, rad1 = r3.w;
    const bool linestrip = (rad0 > rad1 ? rad0 : rad1) <= 1e-2f;
    if(linestrip && id_lo == r.ign_lo && id_hi == r.ign_hi) return 0;
    V3 v0, v1;
    if(mb) { v0 = lerp_v(r0, __ldg(rec + 4), t0, t1); v1 = lerp_v(r1, __ldg(rec + 5), t0, t1); }
    else   { v0 = mk3(r0.x, r0.y, r0.z); v1 = mk3(r1.x, r1.y, r1.z); }
    if((double)fabsf(rad1 - rad0) < 1e-3)
    {
      float out[3], len;
      const float t = cylinder_t(v0, v1, rad0, r, out, len);
      const float lim = r.min_dist > 1e-3f ? r.min_dist : 1e-3f;
      if(((linestrip && t > lim) || (!linestrip && t > r.min_dist)) && t <= max_dist) return 1;
    }
    else
    {
      const float t = cone_t(v0, v1, rad0, rad1, r, max_dist, nullptr);
      if(t > r.min_dist) return 1;
    }
    return 0;
  }
  return 0;
}
