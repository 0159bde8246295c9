// shading.cuh -- vertex preparation, material chain and BSDFs of the wavefront integrator (device side).
// MF_COUNT == 1 (one wavelength per path, include/mf.h:4-6).  Reference sources followed:
//   geometry at a hit      : src/prims.c:254-366 (prims_get_normal_time), include/geo.h:24-44 (oct normals),
//                            include/geo/triangle.h:63-82, sphere.h:54-64, line.h:123-158
//   manifold_init          : include/pathspace/manifold.h:214-233 (flip towards ray, s_inside, scrambled onb)
//   shader_prepare         : src/shader.c:462-542 (+ mult.c:154-167, color.c:75-81, texture.h:38-84,
//                            colorcheckersg.c:244-261, dielectric.c:67-81, metal.c:71-77)
//   diffuse                : src/shader.c:157-257
//   dielectric / ggx       : src/shaders/dielectric.c:96-541, src/shaders/ggx.h
//   metal                  : src/shaders/metal.c:79-310
//   diffdiel               : src/shaders/diffdiel.c (ggx reflection + diffuse transmission, 0030_subsurf's skin surface)
//   media nesting          : src/pathspace.c:80-146 (_path_edge_medium, path_eta_ratio, path_edge_init_volume)
#pragma once
#include "prims.cuh"
#include "sampling.cuh"
#include "corona_b200_render.h"
#include <cuda_fp16.h>

// vertex_scattermode_t (include/pathspace.h:56-69)
enum { M_REFLECT = 1<<0, M_TRANSMIT = 1<<1, M_VOLUME = 1<<2, M_FIBER = 1<<3, M_EMIT = 1<<4, M_SENSOR = 1<<5,
       M_DIFFUSE = 1<<6, M_GLOSSY = 1<<7, M_SPECULAR = 1<<8, M_ABSORB = 0 };
enum { F_INSIDE = 1, F_ENVIRONMENT = 2 };

#define PI_F 3.14159265358979323846f
#define PI_D 3.14159265358979323846

struct SceneGeo
{
  const cb_vtx_t *vtx;
  const cb_vtxidx_t *vtxidx;
  const ShapeDev *shapes;
  const int32_t *shape_material;   // prims_shader(): shape -> material index
};

struct TableDev { float lambda_min, lambda_step; int32_t num_lambda, rows; uint32_t offset; };

struct MaterialsDev
{
  const cb_material_t *mat;
  const TableDev *tables;
  const float *table_data;
  const cb_medium_t *media;   // homogeneous media (cb_material_t.medium / cb_render_desc_t.exterior_medium are 1 + index)
};

struct LightsDev   // src/lights.d/list.c
{
  const uint64_t *primid;     // emissive primitives, grouped by shape
  const float *cdf;           // normalised cumulative area*L
  const float *L;             // L / sum(area*L) per emissive primitive
  const float *shape_pdf;     // the same value per shape id (lights_pdf_next_event), 0 for non-emissive shapes
  uint32_t num;
  float p_geo;
};

struct Vtx
{
  V3 x, n, gn, a, b;
  float u, v, s, t;
  uint32_t prim_lo, prim_hi;
  uint32_t flags, mode, material_modes;
  float rd, rs, rg, em, roughness;
  float ior;         // interior.ior of the shape's material (vacuum 1)
  float eta;         // cached path_eta_ratio
  int32_t mat;
  float vol_mu_s, vol_g;   // volume vertices: interior.mu_s / mean_cos of the medium the vertex sits in
};

// media the path is currently inside of (nested dielectrics / participating media): the incremental form of _path_edge_medium
#define MED_MAX 4
struct Media
{
  uint32_t shape[MED_MAX];
  float ior[MED_MAX];
  int n;
  uint32_t ids;      // 6 bits per entry: 1 + index of the entry's homogeneous medium (0 = none)
};
// PathState packs {n, ids} into one word: bits 0..2 n, bits 8+6k..13+6k the medium of entry k
CBD void media_unpack(Media &m, int32_t word) { m.n = word & 7; m.ids = (uint32_t)word >> 8; }
CBD int32_t media_pack(const Media &m) { return (int32_t)((uint32_t)m.n | (m.ids << 8)); }
CBD uint32_t media_id(const Media &m, int k) { return (m.ids >> (6*k)) & 63u; }
CBD void media_set_id(Media &m, int k, uint32_t id) { m.ids = (m.ids & ~(63u << (6*k))) | (id << (6*k)); }
// highest priority = smallest shape id; empty = global exterior (pathspace.c:107-113).  Returns the entry or -1.
CBD int media_top(const Media &m)
{
  int top = -1;
  uint32_t best = 0xffffffffu;
  for(int i=0;i<m.n;i++) if(m.shape[i] < best) { best = m.shape[i]; top = i; }
  return top;
}
CBD float media_ior(const Media &m)
{
  float ior = 1.0f;
  uint32_t best = 0xffffffffu;
  for(int i=0;i<m.n;i++) if(m.shape[i] < best) { best = m.shape[i]; ior = m.ior[i]; }
  return ior;
}
// 1 + index of the medium the current edge runs through (0 = vacuum)
CBD uint32_t media_medium(const Media &m, uint32_t exterior)
{
  const int top = media_top(m);
  return top < 0 ? exterior : media_id(m, top);
}
// cross the interface of `shape`: returns false on broken nesting
CBD bool media_transmit(Media &m, uint32_t shape, float ior, bool inside, uint32_t medium = 0)
{
  if(!inside)
  {
    if(m.n >= MED_MAX) return false;
    m.shape[m.n] = shape; m.ior[m.n] = ior; media_set_id(m, m.n, medium); m.n++;
    return true;
  }
  for(int i=m.n-1;i>=0;i--)
    if(m.shape[i] == shape)
    {
      m.shape[i] = m.shape[m.n-1]; m.ior[i] = m.ior[m.n-1]; media_set_id(m, i, media_id(m, m.n-1));
      m.n--;
      return true;
    }
  return false;
}
// path_eta_ratio: n1/n2 with n2 on the other side of vertex v; < 0 on broken nesting
CBD float eta_ratio(const Media &m, float cur_ior, const Vtx &v)
{
  if((v.prim_lo & v.prim_hi) == 0xffffffffu) return 1.0f;
  Media t = m;
  if(!media_transmit(t, (v.prim_lo >> 3), v.ior, v.flags & F_INSIDE)) return -1.0f;
  return cur_ior / media_ior(t);
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
CBD const cb_vtx_t *geo_vtx(const SceneGeo &S, uint64_t pid, int k, int close)
{
  const ShapeDev sh = S.shapes[(uint32_t)(pid >> 3) & 0x1fffffffu];
  const uint32_t mb = (uint32_t)(pid >> 60) & 1u;
  const uint32_t vi = S.vtxidx[sh.vtxidx_off + ((uint32_t)(pid >> 32) & 0x0fffffffu) + k].v;
  return S.vtx + sh.vtx_off + (uint64_t)(mb + 1)*vi + (close ? mb : 0);
}
CBD uint32_t geo_uv(const SceneGeo &S, uint64_t pid, int k)
{
  const ShapeDev sh = S.shapes[(uint32_t)(pid >> 3) & 0x1fffffffu];
  return S.vtxidx[sh.vtxidx_off + ((uint32_t)(pid >> 32) & 0x0fffffffu) + k].uv;
}
CBD V3 geo_vertex_time(const SceneGeo &S, uint64_t pid, int k, float time)
{
  const cb_vtx_t *o = geo_vtx(S, pid, k, 0);
  if((pid >> 60) & 1u)
  {
    const cb_vtx_t *c = o + 1;
    const float t0 = 1.0f - time;
    return mk3(t0*o->v[0] + time*c->v[0], t0*o->v[1] + time*c->v[1], t0*o->v[2] + time*c->v[2]);
  }
  return mk3(o->v[0], o->v[1], o->v[2]);
}
CBD V3 decode_normal(uint32_t enc)   // include/geo.h:24-44
{
  const uint32_t p0 = enc & 0xffffu, p1 = enc >> 16;
  const uint32_t v0 = 0x3f800000u | ((p0 & 0x7fffu) << 8);
  const uint32_t v1 = 0x3f800000u | ((p1 & 0x7fffu) << 8);
  V3 r;
  r.x = __uint_as_float(__float_as_uint(2.0f*__uint_as_float(v0) - 2.0f) | ((p0 & 0x8000u) << 16));
  r.y = __uint_as_float(__float_as_uint(2.0f*__uint_as_float(v1) - 2.0f) | ((p1 & 0x8000u) << 16));
  r.z = 1.0f - (fabsf(r.x) + fabsf(r.y));
  if(r.z < 0.0f)
  {
    const float oldx = r.x;
    r.x = (1.0f - fabsf(r.y)) * ((oldx < 0.0f) ? -1.0f : 1.0f);
    r.y = (1.0f - fabsf(oldx)) * ((r.y < 0.0f) ? -1.0f : 1.0f);
  }
  return normalise(r);
}
CBD V3 geo_normal_time(const SceneGeo &S, uint64_t pid, int k, float time)   // include/geo.h:152-162
{
  const V3 n0 = decode_normal(geo_vtx(S, pid, k, 0)->n);
  if((pid >> 60) & 1u)
  {
    const V3 n1 = decode_normal(geo_vtx(S, pid, k, 1)->n);
    const float t0 = 1.0f - time;
    return mk3(t0*n0.x + time*n1.x, t0*n0.y + time*n1.y, t0*n0.z + time*n1.z);
  }
  return n0;
}
CBD float half_to_float_dev(uint16_t h) { return __half2float(__ushort_as_half(h)); }
CBD void decode_uv(uint32_t enc, float &u, float &v) { u = half_to_float_dev(enc & 0xffffu); v = half_to_float_dev(enc >> 16); }

CBD void tri_normal(V3 v0, V3 v1, V3 v2, V3 n0, V3 n1, V3 n2, float u, float v, V3 &gn, V3 &n)
{ // triangle.h:63-82
  gn = mk3((v1.y - v0.y)*(v2.z - v0.z) - (v1.z - v0.z)*(v2.y - v0.y),
           (v1.z - v0.z)*(v2.x - v0.x) - (v1.x - v0.x)*(v2.z - v0.z),
           (v1.x - v0.x)*(v2.y - v0.y) - (v1.y - v0.y)*(v2.x - v0.x));
  gn = normalise(gn);
  const float w = 1.0f - u - v;
  n = normalise(mk3(u*n2.x + v*n1.x + w*n0.x, u*n2.y + v*n1.y + w*n0.y, u*n2.z + v*n1.z + w*n0.z));
}

// prims_get_normal_time: fills gn, n, s, t of the vertex (x, u, v, prim already set)
CBD void geo_normal_uv(const SceneGeo &S, uint64_t pid, float time, Vtx &h)
{
  const uint32_t vcnt = (uint32_t)(pid >> 61) & 7u;
  if(vcnt == CB_PRIM_SPHERE)
  {
    const V3 c = geo_vertex_time(S, pid, 0, time);
    h.gn = normalise(sub(h.x, c));
    h.n = h.gn;
  }
  else if(vcnt == CB_PRIM_LINE)
  { // line.h:123-158
    const V3 v0 = geo_vertex_time(S, pid, 0, time), v1 = geo_vertex_time(S, pid, 1, time);
    const float r0 = __uint_as_float(geo_vtx(S, pid, 0, 0)->n), r1 = __uint_as_float(geo_vtx(S, pid, 1, 0)->n);
    if(fabsf(r0 - r1) < 1e-3f && r0 < 0.01f) { h.n = h.gn = mk3(0.0f, 0.0f, 0.0f); }
    else
    {
      V3 d = sub(v1, v0);
      const float ilen_d = 1.0f/sqrtf(dot(d, d));
      d = mk3(d.x*ilen_d, d.y*ilen_d, d.z*ilen_d);
      V3 a, b;
      onb(d, a, b);
      const float phi = (float)(2.0*PI_D*(double)h.v);
      float sp, cp;
      sincosf(phi, &sp, &cp);
      const V3 nn = mk3(a.x*sp + b.x*cp, a.y*sp + b.y*cp, a.z*sp + b.z*cp);
      const float rr = r1 - r0;
      if((double)fabsf(rr) < 1e-3) h.n = nn;
      else
      {
        const float k = (r1 - r0)*ilen_d;
        h.n = normalise(mk3(nn.x - d.x*k, nn.y - d.y*k, nn.z - d.z*k));
      }
      h.gn = h.n;
    }
  }
  else
  {
    const V3 v0 = geo_vertex_time(S, pid, 0, time), v2 = geo_vertex_time(S, pid, 2, time);
    const V3 n0 = geo_normal_time(S, pid, 0, time), n2 = geo_normal_time(S, pid, 2, time);
    if(vcnt == CB_PRIM_TRI)
      tri_normal(v0, geo_vertex_time(S, pid, 1, time), v2, n0, geo_normal_time(S, pid, 1, time), n2, h.u, h.v, h.gn, h.n);
    else if(h.v >= h.u)
      tri_normal(v0, geo_vertex_time(S, pid, 1, time), v2, n0, geo_normal_time(S, pid, 1, time), n2, h.u, h.v - h.u, h.gn, h.n);
    else
      tri_normal(v0, v2, geo_vertex_time(S, pid, 3, time), n0, n2, geo_normal_time(S, pid, 3, time), h.u - h.v, h.v, h.gn, h.n);
  }
  // texture coordinates (prims.c:303-365)
  if(geo_uv(S, pid, 0) == 0) { h.s = h.u; h.t = h.v; }
  else
  {
    float a0, b0, a1, b1, a2, b2;
    decode_uv(geo_uv(S, pid, 0), a0, b0);
    if(vcnt == CB_PRIM_SPHERE) { h.s = h.u + a0; h.t = h.v + b0; }
    else if(vcnt == CB_PRIM_LINE)
    { // 11/11/10 fixed point (geo.h:91-101)
      const uint32_t e = geo_uv(S, pid, 0);
      h.s = (float)(e >> 21)/2048.0f; h.t = (float)((e & 0x1ffc00u) >> 10)/2048.0f;
    }
    else
    {
      decode_uv(geo_uv(S, pid, 2), a2, b2);
      if(vcnt == CB_PRIM_TRI)
      {
        decode_uv(geo_uv(S, pid, 1), a1, b1);
        h.s = (1.0f - h.u - h.v)*a0 + h.v*a1 + h.u*a2;
        h.t = (1.0f - h.u - h.v)*b0 + h.v*b1 + h.u*b2;
      }
      else if(h.v >= h.u)
      {
        decode_uv(geo_uv(S, pid, 1), a1, b1);
        const float u = h.u, v = h.v - h.u;
        h.s = (1.0f - u - v)*a0 + v*a1 + u*a2;
        h.t = (1.0f - u - v)*b0 + v*b1 + u*b2;
      }
      else
      {
        float a3, b3;
        decode_uv(geo_uv(S, pid, 3), a3, b3);
        const float u = h.u - h.v, v = h.v;
        h.s = (1.0f - u - v)*a0 + v*a2 + u*a3;
        h.t = (1.0f - u - v)*b0 + v*b2 + u*b3;
      }
    }
  }
}

CBD void scrambled_onb(float scramble, V3 n, V3 &u, V3 &v)   // corona_common.h:202-217
{
  if(fabsf(n.y) < scramble) u = cross(n, mk3(0.0f, 1.0f, 0.0f));
  else                      u = cross(n, mk3(1.0f, 0.0f, 0.0f));
  u = normalise(u);
  v = cross(n, u);
}

// ---------------------------------------------------------------------------------------------
// material chain
// ---------------------------------------------------------------------------------------------
CBD float rgb2spec_eval(const float *c, float lambda)   // include/rgb2spec.h:145-149 (rsqrt is approximate there too)
{
  const float x = fmaf(fmaf(c[0], lambda, c[1]), lambda, c[2]);
  const float y = rsqrtf(fmaf(x, x, 1.0f));
  return fmaf(0.5f*x, y, 0.5f);
}
CBD float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// ---- homogeneous media ------------------------------------------------------------------------
// vertex_volume_t of a medium at one wavelength: what `color v` (texture.h:46-52: mu_s <- clamp(albedo), mu_t <- 1) followed by
// medium_rgb's prepare() (medium_rgb.c:46-60: mu_t <- mul*rgb2spec, mu_s *= mu_t/old mu_t) leave in vertex.interior.
// Without a `color v` the old mu_t is vacuum's 0 and mu_s becomes NaN like upstream: such a medium only absorbs.
struct Vol { float mu_t, mu_s, g; bool present; };
CBD Vol medium_eval(const MaterialsDev &M, uint32_t medium, float lambda)
{
  Vol v; v.present = medium > 0; v.mu_t = v.mu_s = v.g = 0.0f;   // path_volume_vacuum (pathspace.h:348-353)
  if(medium > 0)
  {
    const cb_medium_t &m = M.media[medium - 1];
    float mu_s = 0.0f, old_mu_t = 0.0f;
    if(m.has_albedo) { mu_s = clamp01(m.albedo_mul*rgb2spec_eval(m.albedo_coeff, lambda)); old_mu_t = 1.0f; }
    v.mu_t = m.mu_t_mul*rgb2spec_eval(m.mu_t_coeff, lambda);
    v.mu_s = mu_s*(v.mu_t/old_mu_t);
    v.g = m.g;
  }
  return v;
}
// shader_vol_sample's free-flight distance in a scattering medium (shader.c:96-99)
CBD float vol_free_flight(const Vol &v, float rf)
{
  float dist = -logf(1.0f - rf)/v.mu_t;
  if(!(dist > 0.0f)) dist = 1e-15f;
  return dist;
}
// Henyey-Greenstein: sample_eval_hg / sample_hg (sampler_common.h:286-316,338-356); out[0] is along the incoming direction
CBD float hg_eval(float g, V3 wi, V3 wo)
{
  if(g == 0.0f) return (float)(1.0/(4.0*PI_D));
  const float cos_theta = dot(wi, wo);
  return (float)(1.0/(4.0*PI_D)*(double)(1.0f - g*g)/(double)powf(1.0f + g*g - 2.0f*g*cos_theta, 3.0f/2.0f));
}
CBD void hg_sample(float g, float r1, float r2, float *out, float &pdf)
{
  if(g == 0.0f)
  { // sample_sphere (sampler_common.h:136-143)
    out[2] = 1.f - 2.f*r1;
    const float r = sqrtf(1.f - out[2]*out[2]);
    const float phi = (float)(2.0*PI_D*(double)r2);
    out[0] = r*cosf(phi); out[1] = r*sinf(phi);
    pdf = (float)(1.0/(4.0*PI_D));
    return;
  }
  const float sqr = (1.0f - g*g)/(1.0f + g*(2.0f*r1 - 1.0f));
  const float cos_theta = 1.0f/(2.0f*g)*(1.0f + g*g - sqr*sqr);
  const float phi = (float)(2.0*PI_D*(double)r2);
  const float l = sqrtf(fmaxf(0.0f, 1.0f - cos_theta*cos_theta));
  out[0] = cos_theta; out[1] = cosf(phi)*l; out[2] = sinf(phi)*l;
  pdf = (float)(1.0/(4.0*PI_D)*(double)(1.0f - g*g)/(double)powf(1.0f + g*g - 2.0f*g*cos_theta, 3.0f/2.0f));
}
CBD float table_lookup(const MaterialsDev &M, int table, int row, float lambda, bool clamp_index)
{
  const TableDev t = M.tables[table];
  int l = (int)((lambda - t.lambda_min)/t.lambda_step);
  if(clamp_index) l = l < 0 ? 0 : (l > t.num_lambda-1 ? t.num_lambda-1 : l);
  else if(l < 0 || l >= t.num_lambda) return 0.0f;
  return M.table_data[t.offset + row*t.num_lambda + l];
}
CBD void set_slot(Vtx &v, int slot, float data)   // texture.h:38-64 (sensor paths)
{
  switch(slot)
  {
    case CB_SLOT_DIFFUSE:   v.rd = data; return;
    case CB_SLOT_SPECULAR:  v.rs = data; return;
    case CB_SLOT_GLOSSY:    v.rg = data; return;
    case CB_SLOT_ROUGHNESS: v.roughness = data; return;
    case CB_SLOT_EMISSION:  v.em = data; return;
    case CB_SLOT_TRANSMIT_TO_EYE: if(!(v.flags & F_INSIDE)) v.rg = data; return;
    default: return;   // volume slot: part of a medium's chain (cb_medium_t), never of a surface material
  }
}
CBD float dielectric_ior(float n_d, float V_d, float lambda)   // spectrum.h:40-63
{
  if(V_d == 0.0f) return n_d;
  const float l_C = .6563f, l_F = .4861f, l_D = .587561f;
  const float c = (l_C*l_C * l_F*l_F)/(l_C*l_C - l_F*l_F);
  const float B = (n_d - 1.0f)/V_d * c;
  const float A = n_d - B/(l_D*l_D);
  return A + (B*1e6f)/(lambda*lambda);
}
#define DIEL_GLOSSY_THR 1e-3f
#define METAL_GLOSSY_THR 1e-4f
#define DIFFDIEL_GLOSSY_THR 1e-4f
#define HALFVEC_COS_THR .999f
CBD bool indexmatched(float n1, float n2) { return fabsf(1.0f - n1/n2) < 1e-3f; }

// the host bsdf's own prepare() (diffuse: shader.c:157-162, dielectric.c:67-81, metal.c:71-77) and the cached eta ratio
// (shader.c:538): runs after the material chain has filled the shading slots
template<int KINDS = 15>
CBD void bsdf_prepare(const MaterialsDev &M, Vtx &v, float lambda, const Media &med, float cur_ior)
{
  const cb_material_t &m = M.mat[v.mat];
  if(KINDS == 1 || ((KINDS & 1) && m.bsdf == CB_BSDF_DIFFUSE))
  {
    if(v.rd > 0.0f) v.material_modes = M_REFLECT | M_DIFFUSE;
  }
  else if((KINDS & 2) && m.bsdf == CB_BSDF_DIELECTRIC)
  {
    v.ior = dielectric_ior(m.param[0], m.param[1], lambda);
    v.material_modes = M_REFLECT | M_TRANSMIT;
    const float eta = eta_ratio(med, cur_ior, v);
    if(indexmatched(eta, 1.0f)) v.roughness = 0.0f;
    if(v.roughness > DIEL_GLOSSY_THR) v.material_modes |= M_GLOSSY; else v.material_modes |= M_SPECULAR;
  }
  else if((KINDS & 4) && m.bsdf == CB_BSDF_METAL)
  {
    v.material_modes = M_REFLECT;
    if(v.roughness > METAL_GLOSSY_THR) v.material_modes |= M_GLOSSY; else v.material_modes |= M_SPECULAR;
  }
  else if((KINDS & 8) && m.bsdf == CB_BSDF_DIFFDIEL)
  { // diffdiel.c:71-85
    v.ior = dielectric_ior(m.param[0], m.param[1], lambda);
    v.material_modes = M_REFLECT | M_TRANSMIT;
    const float eta = eta_ratio(med, cur_ior, v);
    if(indexmatched(eta, 1.0f)) v.roughness = 0.0f;
    if(v.roughness > DIFFDIEL_GLOSSY_THR) v.material_modes |= M_GLOSSY; else v.material_modes |= M_SPECULAR;
  }
  v.eta = eta_ratio(med, cur_ior, v);
}

// shader_prepare for a surface vertex whose x, u, v, prim are set and whose incoming direction is `omega`
template<int KINDS = 15>
CBD void prepare_vertex(const SceneGeo &S, const MaterialsDev &M, Vtx &v, V3 omega, float time, float lambda, float scramble,
                        const Media &med, float cur_ior)
{
  const uint64_t pid = (uint64_t)v.prim_lo | ((uint64_t)v.prim_hi << 32);
  geo_normal_uv(S, pid, time, v);
  if(dot(omega, v.gn) > 0.0f) { v.n = mk3(-v.n.x, -v.n.y, -v.n.z); v.flags |= F_INSIDE; }
  else v.flags &= ~F_INSIDE;
  scrambled_onb(scramble, v.n, v.a, v.b);
  v.mat = S.shape_material[(v.prim_lo >> 3)];
  v.rd = v.rs = v.rg = v.em = 0.0f;
  v.roughness = 1.0f;
  v.ior = 1.0f;
  const cb_material_t &m = M.mat[v.mat];
  // `interior`: the medium's chain runs first and its `color v` announces a volume lobe (interior.c:113-118, texture.h:46-52);
  // the surface bsdf's prepare() normally replaces it
  if(m.medium > 0 && M.media[m.medium - 1].has_albedo) v.material_modes = M_VOLUME | M_GLOSSY;
  for(int k=0;k<m.num_ops;k++)
  {
    const cb_matop_t &op = m.ops[k];
    if(op.op == CB_OP_COLOR)
    {
      v.roughness = op.roughness;
      const float val = op.mul * rgb2spec_eval(op.coeff, lambda);
      set_slot(v, op.slot, op.slot == CB_SLOT_EMISSION ? val : clamp01(val));
    }
    else if(op.op == CB_OP_CHECKERSG)
    {
      const int i = (int)(14.0f*v.s) % 14, j = (int)(10.0f*v.t) % 10;
      const float fu = fmodf(14.0f*v.s, 1.0f), fv = fmodf(10.0f*v.t, 1.0f);
      float val;
      if(fu < 0.1f || fu > 0.9f || fv < 0.1f || fv > 0.9f) val = 0.3f;
      else val = table_lookup(M, op.table, 14*j + i, lambda, false);
      set_slot(v, op.slot, val);
    }
  }
  bsdf_prepare<KINDS>(M, v, lambda, med, cur_ior);
}

// ---------------------------------------------------------------------------------------------
// ggx (src/shaders/ggx.h)
// ---------------------------------------------------------------------------------------------
CBD float ggx_G1(V3 w, V3 n, float roughness)
{
  const float r2 = roughness*roughness;
  const float cos_th = fabsf(dot(w, n));
  const float sin_th = sqrtf(fmaxf(0.0f, 1.0f - cos_th*cos_th));
  const float tan_th = sin_th/cos_th;
  return 2.0f/(1.0f + sqrtf(1.0f + r2*tan_th*tan_th));
}
CBD float ggx_G1_cos(float cos_wn, float roughness)
{
  const float r2 = roughness*roughness;
  const float sin_wn = sqrtf(clamp01(1.0f - cos_wn*cos_wn));
  const float tan_th = sin_wn/cos_wn;
  return 2.0f/(1.0f + sqrtf(1.0f + (r2*tan_th)*tan_th));
}
CBD void ggx_sample11(float tan_theta_i, float U1, float U2, float &slope_x, float &slope_y)
{
  if(tan_theta_i < 0.0001f)
  {
    const float r = sqrtf(U1/fmaxf(1e-8f, 1.0f - U1));
    const float phi = (float)(2.0*PI_D*(double)U2);
    slope_x = r*cosf(phi);
    slope_y = r*sinf(phi);
    return;
  }
  const float a = 1.0f/tan_theta_i;
  const float G1 = 2.0f/(1.0f + sqrtf(1.0f + 1.0f/(a*a)));
  const float A = 2.0f*U1/G1 - 1.0f;
  const float tmp = 1.0f/(A*A - 1.0f);
  const float B = tan_theta_i;
  const float D = sqrtf(fmaxf(0.0f, B*B*tmp*tmp - (A*A - B*B)*tmp));
  float sx1 = B*tmp - D, sx2 = B*tmp + D;
  if(!(fabsf(sx1) < FLT_MAX)) sx1 = 0.0f;
  if(!(fabsf(sx2) < FLT_MAX)) sx2 = 0.0f;
  slope_x = (A < 0.0f || sx2*tan_theta_i > 1.0f) ? sx1 : sx2;
  float Sg;
  if(U2 > 0.5f) { Sg = 1.0f; U2 = 2.0f*(U2 - 0.5f); }
  else          { Sg = -1.0f; U2 = 2.0f*(0.5f - U2); }
  const float z = (U2*(U2*(U2*(-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f) /
                  (U2*(U2*(U2*(U2*0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.0f) - 0.539825872510702f);
  slope_y = Sg*z*sqrtf(1.0f + slope_x*slope_x);
}
CBD V3 ggx_sample_h(V3 wi, float rx, float ry, float U1, float U2)
{
  V3 w = normalise(mk3(rx*wi.x, ry*wi.y, fabsf(wi.z)));
  float tan_theta = 0.0f, sin_phi = 0.0f, cos_phi = 1.0f;
  if(w.z < 0.99999f)
  {
    const float len = sqrtf(w.x*w.x + w.y*w.y);
    tan_theta = len/w.z; sin_phi = w.y/len; cos_phi = w.x/len;
  }
  float sx, sy;
  ggx_sample11(tan_theta, U1, U2, sx, sy);
  const float tmp = cos_phi*sx - sin_phi*sy;
  sy = sin_phi*sx + cos_phi*sy;
  sx = rx*tmp;
  sy = ry*sy;
  const float inv_h = sqrtf(sx*sx + sy*sy + 1.0f);
  if(!(inv_h > 0.0f)) return mk3(0.0f, 1.0f, 0.0f);
  return mk3(-sx/inv_h, -sy/inv_h, 1.0f/inv_h);
}
CBD float ggx_pdf_h(V3 wi, V3 h, V3 n, float roughness)
{
  const float r2 = roughness*roughness;
  const float cos_th = fabsf(dot(h, n));
  const float sin_th = sqrtf(fmaxf(0.0f, 1.0f - cos_th*cos_th));
  const float tan_th = sin_th/cos_th;
  const float D_h = r2/(PI_F*cos_th*cos_th*cos_th*cos_th*(r2 + tan_th*tan_th)*(r2 + tan_th*tan_th));
  const float G1 = ggx_G1(wi, n, roughness);
  return fabsf(G1*dot(wi, h)*D_h/dot(wi, n));
}
CBD float ggx_pdf_h_cos(float cosh, float cos_in, float cosr, float roughness)
{
  const float r2 = roughness*roughness;
  const float cosh2 = cosh*cosh;
  const float sin_th = sqrtf(clamp01(1.0f - cosh2));
  const float tan_th = sin_th/fabsf(cosh);
  const float den = fmaf(tan_th, tan_th, r2);
  const float ct4 = cosh2*cosh2;
  const float D_h = r2/((PI_F*ct4)*(den*den));
  const float G1 = ggx_G1_cos(cos_in, roughness);
  return fabsf((G1*cosr)*(D_h/cos_in));
}

// ---------------------------------------------------------------------------------------------
// BSDF interface.  wi = e[v].omega (incoming, pointing at the vertex), wo = e[v+1].omega.
// sample: returns throughput weight, sets wo, pdf (projected solid angle), v.mode
// ---------------------------------------------------------------------------------------------
CBD float fresnel_dielectric(float n1, float n2, float cosr, float cost)
{
  if(cost <= 0.0f) return 1.0f;
  const float r1 = n1*cosr, r2 = n2*cosr, t1 = n1*cost, t2 = n2*cost;
  const float Rs = (r1 - t2)/(r1 + t2);
  const float Rp = (t1 - r2)/(t1 + r2);
  return clamp01((Rs*Rs + Rp*Rp)*.5f);
}
CBD float fresnel_conductor(float n1, float n2, float k2, float cosr)   // metal.c:79-164
{
  const float d = n2*n2 + k2*k2;
  const float etar = (n1*n2)/d, etai = -((n1*k2)/d);
  const float eta2r = etar*etar - etai*etai, eta2i = (2.0f*etar)*etai;
  const float sinr = 1.0f - cosr*cosr;
  const float cost2r = 1.0f - eta2r*sinr, cost2i = eta2i*(-sinr);
  const float len = sqrtf(cost2r*cost2r + cost2i*cost2i);
  const float costr = sqrtf(0.5f*(cost2r + len));
  float costi = sqrtf(0.5f*(len - cost2r));
  if(cost2i < 0.0f) costi = -costi;
  const float n1cosr = n1*cosr, n2cosrr = n2*cosr, n2cosri = k2*cosr;
  const float n1costr = n1*costr, n1costi = n1*costi;
  const float n2costr = n2*costr - k2*costi, n2costi = k2*costr + n2*costi;
  const float Rs2 = ((n1cosr - n2costr)*(n1cosr - n2costr) + n2costi*n2costi) /
                    ((n1cosr + n2costr)*(n1cosr + n2costr) + n2costi*n2costi);
  const float Rp2 = ((n1costr - n2cosrr)*(n1costr - n2cosrr) + (n1costi - n2cosri)*(n1costi - n2cosri)) /
                    ((n1costr + n2cosrr)*(n1costr + n2cosrr) + (n1costi + n2cosri)*(n1costi + n2cosri));
  return clamp01((Rs2 + Rp2)*.5f);
}

template<int KINDS = 15>
CBD float bsdf_sample(const MaterialsDev &M, Vtx &v, V3 wi, float lambda, float cur_ior, float r_x, float r_y, float r_mode,
                      V3 &wo, float &pdf)
{
  const cb_material_t &m = M.mat[v.mat];
  if(KINDS == 1 || ((KINDS & 1) && m.bsdf == CB_BSDF_DIFFUSE))
  { // sample_d, shader.c:165-205
    const float x1 = r_x, x2 = r_y;
    const float s = sqrtf(x1);
    const float c0 = sqrtf((float)(1.0 - (double)x1));
    const float ang = (float)(2.0*PI_D*(double)x2);
    const float c1 = s*cosf(ang), c2 = s*sinf(ang);
    wo = mk3(c0*v.n.x + c1*v.a.x + c2*v.b.x, c0*v.n.y + c1*v.a.y + c2*v.b.y, c0*v.n.z + c1*v.a.z + c2*v.b.z);
    pdf = (float)(1.0/PI_D);
    const float cos_out_ng = dot(v.gn, wo);
    if(v.flags & F_INSIDE) { if(cos_out_ng >= 0.0f) return 0.0f; }
    else if(cos_out_ng <= 0.0f) return 0.0f;
    if(v.rd > 0.0f) v.mode = M_DIFFUSE | M_REFLECT;
    return v.rd;
  }
  if((KINDS & 2) && m.bsdf == CB_BSDF_DIELECTRIC)
  { // dielectric.c:240-381 (MF_COUNT == 1 branch)
    const float eta = v.eta;
    if(eta < 0.0f) return 0.0f;
    if(indexmatched(eta, 1.0f))
    {
      wo = wi; v.mode = M_SPECULAR | M_TRANSMIT; pdf = 1.0f;
      return v.rg;
    }
    V3 h = v.n;
    float pdf_h = 1.0f;
    const float r = v.roughness;
    const float cos_in = -dot(v.n, wi);
    if(r > DIEL_GLOSSY_THR)
    {
      const V3 wit = mk3(-dot(v.a, wi), -dot(v.b, wi), cos_in);
      const V3 ht = ggx_sample_h(wit, r, r, r_x, r_y);
      h = mk3(ht.x*v.a.x + ht.y*v.b.x + ht.z*v.n.x, ht.x*v.a.y + ht.y*v.b.y + ht.z*v.n.y, ht.x*v.a.z + ht.y*v.b.z + ht.z*v.n.z);
      pdf_h = ggx_pdf_h(wi, h, v.n, r);
    }
    float p = pdf_h;
    const float cosr = -dot(wi, h);
    if(cosr <= 0.0f) return 0.0f;
    const float n1 = eta, n2 = 1.0f;
    const float nr = n1/n2;
    const float cost2 = 1.0f - (nr*nr)*(1.0f - cosr*cosr);
    const float cost = cost2 <= 0.0f ? 0.0f : sqrtf(cost2);
    const float R = fresnel_dielectric(n1, n2, cosr, cost);
    if(r_mode <= R)
    {
      v.mode = M_REFLECT;
      wo = mk3(wi.x + 2.0f*cosr*h.x, wi.y + 2.0f*cosr*h.y, wi.z + 2.0f*cosr*h.z);
      if(dot(wo, v.n) <= 0.0f) return 0.0f;
      p *= 1.0f/(4.0f*cosr);
      if(r > DIEL_GLOSSY_THR)
      {
        pdf = R*(p/fabsf(dot(wo, v.n)));
        v.mode |= M_GLOSSY;
        if(dot(wo, v.n)*dot(wo, h) < 0.0f) return 0.0f;
        return v.rg*ggx_G1(wo, v.n, v.roughness);
      }
      pdf = R;
      v.mode = M_REFLECT | M_SPECULAR;
      return v.rg;
    }
    if(cost2 <= 0.0f) return 0.0f;
    const float f = eta*cosr - cost;
    wo = normalise(mk3(wi.x*eta + f*h.x, wi.y*eta + f*h.y, wi.z*eta + f*h.z));
    if(dot(wo, v.n) >= 0.0f) return 0.0f;
    if(r <= DIEL_GLOSSY_THR)
    {
      pdf = 1.0f - R;
      v.mode = M_SPECULAR | M_TRANSMIT;
      return v.rg;
    }
    const float denom = n1*cosr - n2*cost;
    p *= n2*n2*cost/(denom*denom);
    pdf = (p*(1.0f - R))/fabsf(dot(wo, v.n));
    v.mode = M_TRANSMIT | M_GLOSSY;
    return v.rg*ggx_G1(wo, v.n, v.roughness);
  }
  if((KINDS & 8) && m.bsdf == CB_BSDF_DIFFDIEL)
  { // diffdiel.c:223-306 (culled_modes == 0): ggx reflection off the interface, cosine-distributed diffuse transmission
    const float eta = v.eta;
    if(eta < 0.0f) return 0.0f;
    if(indexmatched(eta, 1.0f))
    {
      wo = wi; v.mode = M_SPECULAR | M_TRANSMIT; pdf = 1.0f;
      return v.rg;
    }
    V3 h = v.n;
    float pdf_h = 1.0f;
    const float r = v.roughness;
    const float cos_in = -dot(v.n, wi);
    if(r > DIFFDIEL_GLOSSY_THR)
    {
      const V3 wit = mk3(-dot(v.a, wi), -dot(v.b, wi), cos_in);
      const V3 ht = ggx_sample_h(wit, r, r, r_x, r_y);
      h = mk3(ht.x*v.a.x + ht.y*v.b.x + ht.z*v.n.x, ht.x*v.a.y + ht.y*v.b.y + ht.z*v.n.y, ht.x*v.a.z + ht.y*v.b.z + ht.z*v.n.z);
      pdf_h = ggx_pdf_h(wi, h, v.n, r);
    }
    float p = pdf_h;
    const float cosr = -dot(wi, h);
    if(cosr <= 0.0f) return 0.0f;
    const float n1 = eta, n2 = 1.0f;
    const float nr = n1/n2;
    const float cost2 = 1.0f - (nr*nr)*(1.0f - cos_in*cos_in);   // "non-reciprocal fake fresnel": on the macro normal
    const float cost = cost2 <= 0.0f ? 0.0f : sqrtf(cost2);
    const float R = fresnel_dielectric(n1, n2, cos_in, cost);
    if(r_mode <= R)
    {
      v.mode = M_REFLECT;
      wo = mk3(wi.x + 2.0f*cosr*h.x, wi.y + 2.0f*cosr*h.y, wi.z + 2.0f*cosr*h.z);
      if(dot(wo, v.n) <= 0.0f) return 0.0f;
      p *= 1.0f/(4.0f*cosr);
      if(r > DIFFDIEL_GLOSSY_THR)
      {
        pdf = R*(p/fabsf(dot(wo, v.n)));
        v.mode |= M_GLOSSY;
        if(dot(wo, v.n)*dot(wo, h) < 0.0f) return 0.0f;
        return v.rg*ggx_G1(wo, v.n, v.roughness);
      }
      pdf = R;
      v.mode |= M_SPECULAR;
      return v.rg;
    }
    v.mode = M_GLOSSY | M_TRANSMIT;
    pdf = (1.0f - R)/PI_F;
    // sample_cos (sampler_common.h:156-162)
    const float su = sqrtf(r_x);
    const float ang = (float)((double)2.f*PI_D*(double)r_y);
    const float c0 = su*cosf(ang), c1 = su*sinf(ang), c2 = sqrtf((float)(1.0 - (double)r_x));
    wo = mk3(v.a.x*c0 + v.b.x*c1 - v.n.x*c2, v.a.y*c0 + v.b.y*c1 - v.n.y*c2, v.a.z*c0 + v.b.z*c1 - v.n.z*c2);
    return v.rg;
  }
  // metal.c:208-256
  if((KINDS & 4) && (KINDS == 4 || m.bsdf == CB_BSDF_METAL))
  {
    V3 h = v.n;
    float pdf_h = 1.0f;
    const float r = v.roughness;
    if(r > METAL_GLOSSY_THR)
    {
      const V3 wit = mk3(-dot(v.a, wi), -dot(v.b, wi), -dot(v.n, wi));
      const V3 ht = ggx_sample_h(wit, r, r, r_x, r_y);
      h = mk3(ht.x*v.a.x + ht.y*v.b.x + ht.z*v.n.x, ht.x*v.a.y + ht.y*v.b.y + ht.z*v.n.y, ht.x*v.a.z + ht.y*v.b.z + ht.z*v.n.z);
      pdf_h = ggx_pdf_h(wi, h, v.n, r);
    }
    float p = pdf_h;
    const float cosr = -dot(wi, h);
    if(!(cosr > 0.0f)) return 0.0f;
    const float n2 = table_lookup(M, m.table, 0, lambda, true), k2 = -table_lookup(M, m.table, 1, lambda, true);
    const float R = fresnel_conductor(cur_ior, n2, k2, cosr);
    v.mode = M_REFLECT;
    wo = mk3(wi.x + 2.0f*cosr*h.x, wi.y + 2.0f*cosr*h.y, wi.z + 2.0f*cosr*h.z);
    if(dot(wo, v.n) <= 0.0f) return 0.0f;
    p *= 1.0f/(4.0f*cosr);
    if(r > METAL_GLOSSY_THR)
    {
      pdf = p/fabsf(dot(wo, v.n));
      v.mode |= M_GLOSSY;
      if(dot(wo, v.n)*dot(wo, h) < 0.0f) return 0.0f;
      return R*(v.rg*ggx_G1(wo, v.n, v.roughness));
    }
    pdf = 1.0f;   // metal.c leaves p->v[v+1].pdf at the 1.0 path_extend initialised it with
    v.mode |= M_SPECULAR;
    return R*v.rg;
  }
  return 0.0f;   // a material kind this kernel variant was not compiled for: cb200_render_create never selects such a variant
}

// shader_brdf: evaluates f for (wi -> wo) and sets v.mode
template<int KINDS = 15>
CBD float bsdf_eval(const MaterialsDev &M, Vtx &v, V3 wi, V3 wo, float lambda, float cur_ior)
{
  const cb_material_t &m = M.mat[v.mat];
  if(KINDS == 1 || ((KINDS & 1) && m.bsdf == CB_BSDF_DIFFUSE))
  { // brdf_d (sensor paths), shader.c:207-249
    v.mode = M_DIFFUSE | M_REFLECT;
    const float cos_out_ns = dot(v.n, wo);
    if(cos_out_ns <= 0.0f) return 0.0f;
    const float cos_out_ng = dot(v.gn, wo);
    if(v.flags & F_INSIDE) { if(cos_out_ng >= 0.0f) return 0.0f; }
    else if(cos_out_ng <= 0.0f) return 0.0f;
    return v.rd*(float)(1.0/PI_D);
  }
  if((KINDS & 2) && m.bsdf == CB_BSDF_DIELECTRIC)
  { // dielectric.c:386-510
    const float cos_in = -dot(v.n, wi), cos_out = dot(v.n, wo);
    const float eta = v.eta;
    if(eta < 0.0f) return 0.0f;
    const float n1 = eta, n2 = 1.0f;
    const bool matched = indexmatched(n1, n2);
    if(cos_out == 0.0f || cos_in == 0.0f) return 0.0f;
    if(!matched && (cos_in*cos_out > 0.0f)) v.mode = M_REFLECT; else v.mode = M_TRANSMIT;
    const float r = v.roughness;
    if((r > DIEL_GLOSSY_THR) && !matched) v.mode |= M_GLOSSY; else v.mode |= M_SPECULAR;
    if(matched)
    {
      const float dwn = dot(wo, v.n);
      const V3 h = normalise(mk3(-wi.x + wo.x - 2.0f*dwn*v.n.x, -wi.y + wo.y - 2.0f*dwn*v.n.y, -wi.z + wo.z - 2.0f*dwn*v.n.z));
      const float cosh = dot(h, v.n);
      if(cosh < 0.0f || cosh < HALFVEC_COS_THR) return 0.0f;
      return v.rg;
    }
    if(v.mode & M_REFLECT)
    {
      const V3 h = normalise(mk3(-wi.x + wo.x, -wi.y + wo.y, -wi.z + wo.z));
      const float cosh = dot(h, v.n);
      if(cosh < 0.0f) return 0.0f;
      const float DG1 = (v.mode & M_SPECULAR) ? 1.0f : ggx_pdf_h(wi, h, v.n, v.roughness);
      if(DG1 == 0.0f) return 0.0f;
      const float cosr = -dot(h, wi);
      if(cosr < 0.0f) return 0.0f;
      const float nr = n1/n2;
      const float cost2 = 1.0f - (nr*nr)*(1.0f - cosr*cosr);
      const float cost = cost2 <= 0.0f ? 0.0f : sqrtf(cost2);
      const float R = fresnel_dielectric(n1, n2, cosr, cost);
      const float G1 = ggx_G1(wo, v.n, v.roughness);
      if(v.mode & M_GLOSSY) return (v.rg*R)*(DG1*G1/(4.0f*fabsf(cosr*cos_out)));
      if(cosh < HALFVEC_COS_THR) return 0.0f;
      return v.rg*R;
    }
    // transmit
    bool mask = false;
    float h0 = n1*wi.x - n2*wo.x, h1 = n1*wi.y - n2*wo.y, h2 = n1*wi.z - n2*wo.z;
    const float hil = 1.0f/sqrtf(fmaf(h0, h0, fmaf(h1, h1, h2*h2)));
    h0 *= hil; h1 *= hil; h2 *= hil;
    float cosh2 = fmaf(h0, v.n.x, fmaf(h1, v.n.y, h2*v.n.z));
    const bool lt0 = cosh2 < 0.0f;
    mask |= lt0 && (n1 < n2);
    mask |= (!lt0) && (n2 < n1);
    if(lt0) { cosh2 = -cosh2; h0 = -h0; h1 = -h1; h2 = -h2; }
    const float cosr2 = fmaf(h0, -wi.x, fmaf(h1, -wi.y, h2*(-wi.z)));
    mask |= cosr2 <= 0.0f;
    const float nr = n1/n2;
    const float cost2 = 1.0f - (nr*nr)*(1.0f - cosr2*cosr2);
    const float cost = cost2 <= 0.0f ? 0.0f : sqrtf(cost2);
    const float R2 = fresnel_dielectric(n1, n2, cosr2, cost);
    const float DG1 = ggx_pdf_h_cos(cosh2, cos_in, cosr2, v.roughness);
    const float G1 = ggx_G1_cos(cos_in, v.roughness);
    const float cos_hwo = fmaf(h0, wo.x, fmaf(h1, wo.y, h2*wo.z));
    mask |= cos_hwo >= 0.0f;
    float denom = n1*cosr2 - n2*cost;
    denom = denom*denom;
    if(v.mode & M_GLOSSY)
      return mask ? 0.0f : ((v.rg*(1.0f - R2))*((n2*n2)*(cost*(DG1*(G1*(1.0f/fabsf(cos_out)))))))/denom;
    mask |= cosh2 < HALFVEC_COS_THR;
    return mask ? 0.0f : v.rg*clamp01(1.0f - R2);
  }
  if((KINDS & 8) && m.bsdf == CB_BSDF_DIFFDIEL)
  { // diffdiel.c:309-383
    const float cos_in = -dot(v.n, wi), cos_out = dot(v.n, wo);
    const float eta = v.eta;
    if(eta < 0.0f) return 0.0f;
    const float n1 = eta, n2 = 1.0f;
    const bool matched = indexmatched(n1, n2);
    if(cos_out == 0.0f || cos_in == 0.0f) return 0.0f;
    if(!matched && (cos_in*cos_out > 0.0f)) v.mode = M_REFLECT; else v.mode = M_TRANSMIT;
    const float r = v.roughness;
    if((r > DIFFDIEL_GLOSSY_THR) && !matched) v.mode |= M_GLOSSY; else v.mode |= M_SPECULAR;
    const float nr = n1/n2;
    const float cost2 = 1.0f - (nr*nr)*(1.0f - cos_in*cos_in);
    const float cost = cost2 <= 0.0f ? 0.0f : sqrtf(cost2);
    const float R = fresnel_dielectric(n1, n2, cos_in, cost);
    if(matched)
    {
      const float dwn = dot(wo, v.n);
      const V3 h = normalise(mk3(-wi.x + wo.x - 2.0f*dwn*v.n.x, -wi.y + wo.y - 2.0f*dwn*v.n.y, -wi.z + wo.z - 2.0f*dwn*v.n.z));
      const float cosh = dot(h, v.n);
      if(cosh < 0.0f || cosh < HALFVEC_COS_THR) return 0.0f;
      return v.rg;
    }
    if(v.mode & M_REFLECT)
    {
      const V3 h = normalise(mk3(-wi.x + wo.x, -wi.y + wo.y, -wi.z + wo.z));
      const float cosh = dot(h, v.n);
      if(cosh < 0.0f) return 0.0f;
      const float cosr = dot(h, wo);
      const float DG1 = (v.mode & M_SPECULAR) ? 1.0f : ggx_pdf_h(wi, h, v.n, v.roughness);
      if(DG1 == 0.0f) return 0.0f;
      const float G1 = ggx_G1(wo, v.n, v.roughness);
      if(v.mode & M_GLOSSY) return (v.rg*R)*(DG1*G1/(4.0f*fabsf(cosr*cos_out)));
      if(cosh < HALFVEC_COS_THR) return 0.0f;
      return v.rg*R;
    }
    if(v.mode & M_GLOSSY) return v.rg*(clamp01(1.0f - R)/PI_F);
    return 0.0f;   // "pure specular case": cosh is still 0 there and fails the half vector test
  }
  // metal.c:259-310
  if((KINDS & 4) && (KINDS == 4 || m.bsdf == CB_BSDF_METAL))
  {
    const float cos_in = -dot(v.n, wi), cos_out = dot(v.n, wo);
    if(cos_out <= 0.0f || cos_in <= 0.0f) return 0.0f;
    v.mode = M_REFLECT;
    v.mode |= (v.roughness > METAL_GLOSSY_THR) ? M_GLOSSY : M_SPECULAR;
    const V3 h = normalise(mk3(-wi.x + wo.x, -wi.y + wo.y, -wi.z + wo.z));
    const float cosh = dot(h, v.n);
    if(cosh < 0.0f) return 0.0f;
    const float DG1 = ggx_pdf_h(wi, h, v.n, v.roughness);
    if(DG1 == 0.0f) return 0.0f;
    const float cosr = -dot(h, wi);
    if(cosr < 0.0f) return 0.0f;
    const float n2 = table_lookup(M, m.table, 0, lambda, true), k2 = -table_lookup(M, m.table, 1, lambda, true);
    const float R = fresnel_conductor(cur_ior, n2, k2, cosr);
    const float G1 = ggx_G1(wo, v.n, v.roughness);
    if(v.mode & M_GLOSSY) return (v.rg*R)*(DG1*G1/(4.0f*fabsf(cosr*cos_out)));
    if(cosh < HALFVEC_COS_THR) return 0.0f;
    return v.rg*R;
  }
  return 0.0f;
}

// shader_pdf(p, v) for the forward direction (e1 < e2): projected solid angle pdf of wo given wi, for v.mode
template<int KINDS = 15>
CBD float bsdf_pdf(const MaterialsDev &M, const Vtx &v, V3 wi, V3 wo)
{
  const cb_material_t &m = M.mat[v.mat];
  if(KINDS == 1 || ((KINDS & 1) && m.bsdf == CB_BSDF_DIFFUSE)) return (float)(1.0/PI_D);
  if((KINDS & 2) && m.bsdf == CB_BSDF_DIELECTRIC)
  { // dielectric.c:96-237, culled_modes == 0
    const V3 n = v.n;
    const float cos_in = -dot(n, wi), cos_out = dot(n, wo);
    if(cos_in*cos_out == 0.0f) return 0.0f;
    if(cos_out > 0.0f && !(v.mode & M_REFLECT)) return 0.0f;
    if(cos_out < 0.0f && !(v.mode & M_TRANSMIT)) return 0.0f;
    const float eta = v.eta;
    if(eta < 0.0f) return 0.0f;
    const float n1 = eta, n2 = 1.0f;
    bool mask = false;
    float cosr = 0.0f, cosh = 0.0f;
    V3 h = mk3(0.0f, 0.0f, 0.0f);
    if(indexmatched(n1, n2))
    {
      const float dwn = dot(wo, n);
      h = normalise(mk3(-wi.x + wo.x - 2.0f*dwn*n.x, -wi.y + wo.y - 2.0f*dwn*n.y, -wi.z + wo.z - 2.0f*dwn*n.z));
      if(v.mode != (M_TRANSMIT | M_SPECULAR)) return 0.0f;
      if(dot(h, n) < HALFVEC_COS_THR) return 0.0f;
      return 1.0f;
    }
    else if(v.mode & M_REFLECT)
    {
      h = normalise(sub(wi, wo));
      cosh = fabsf(dot(h, n));
      cosr = fabsf(dot(h, wi));
    }
    else
    {
      float h0 = n1*wi.x - n2*wo.x, h1 = n1*wi.y - n2*wo.y, h2 = n1*wi.z - n2*wo.z;
      const float hil = 1.0f/sqrtf(fmaf(h0, h0, fmaf(h1, h1, h2*h2)));
      h0 *= hil; h1 *= hil; h2 *= hil;
      if(n2 < n1) { h0 = -h0; h1 = -h1; h2 = -h2; }
      cosh = fmaf(h0, n.x, fmaf(h1, n.y, h2*n.z));
      mask |= cosh < 0.0f;
      cosr = fmaf(h0, -wi.x, fmaf(h1, -wi.y, h2*(-wi.z)));
      mask |= cosr <= 0.0f;
      h = mk3(h0, h1, h2);
    }
    const float nr = n1/n2;
    const float cost2 = 1.0f - (nr*nr)*(1.0f - cosr*cosr);
    const float cost = cost2 <= 0.0f ? 0.0f : sqrtf(cost2);
    const float R = fresnel_dielectric(n1, n2, cosr, cost);
    float pdf = 1.0f;
    if(v.mode & M_REFLECT)
    {
      if(v.mode & M_SPECULAR) { mask |= cosh < HALFVEC_COS_THR; return mask ? 0.0f : R; }
      pdf *= 1.0f/(4.0f*fabsf(dot(wo, h)));
      pdf *= R;
    }
    else
    {
      if(v.mode & M_SPECULAR) { mask |= cosh < HALFVEC_COS_THR; return mask ? 0.0f : clamp01(1.0f - R); }
      const float denom = n1*cosr - n2*cost;
      pdf *= ((n2*n2)*cost)/(denom*denom);
      pdf *= clamp01(1.0f - R);
    }
    pdf *= ggx_pdf_h_cos(cosh, cos_in, cosr, v.roughness);
    pdf /= fabsf(cos_out);
    mask |= !(pdf > 0.0f);
    return mask ? 0.0f : pdf;
  }
  if((KINDS & 8) && m.bsdf == CB_BSDF_DIFFDIEL)
  { // diffdiel.c:96-221, culled_modes == 0, forward direction
    const V3 n = v.n;
    const float cos_in = -dot(n, wi), cos_out = dot(n, wo);
    if(cos_in*cos_out == 0.0f) return 0.0f;
    if(cos_out > 0.0f && !(v.mode & M_REFLECT)) return 0.0f;
    if(cos_out < 0.0f && !(v.mode & M_TRANSMIT)) return 0.0f;
    const float eta = v.eta;
    if(eta < 0.0f) return 0.0f;
    const float n1 = eta, n2 = 1.0f;
    bool mask = false;
    float cosr = 0.0f, cosh = 0.0f;
    V3 h = mk3(0.0f, 0.0f, 0.0f);
    if(indexmatched(n1, n2))
    {
      const float dwn = dot(wo, n);
      h = normalise(mk3(-wi.x + wo.x - 2.0f*dwn*n.x, -wi.y + wo.y - 2.0f*dwn*n.y, -wi.z + wo.z - 2.0f*dwn*n.z));
      if(v.mode != (M_TRANSMIT | M_SPECULAR)) return 0.0f;
      if(dot(h, n) < HALFVEC_COS_THR) return 0.0f;
      return 1.0f;
    }
    else if(v.mode & M_REFLECT)
    {
      h = normalise(sub(wi, wo));
      cosh = fabsf(dot(h, n));
      cosr = fabsf(dot(h, wi));
    }
    const float nr = n1/n2;
    const float cost2 = 1.0f - (nr*nr)*(1.0f - cos_in*cos_in);
    const float cost = cost2 <= 0.0f ? 0.0f : sqrtf(cost2);
    const float R = fresnel_dielectric(n1, n2, cos_in, cost);
    float pdf = 1.0f;
    if(v.mode & M_REFLECT)
    {
      if(v.mode & M_SPECULAR) { mask |= cosh < HALFVEC_COS_THR; return mask ? 0.0f : R; }
      pdf *= 1.0f/(4.0f*fabsf(dot(wo, h)));
      pdf *= R;
    }
    else
    {
      if(v.mode & M_SPECULAR) return 0.0f;   // cosh == 0 < HALFVEC_COS_THR masks the specular transmit case out
      pdf = (float)(1.0/PI_D);
      pdf *= clamp01(1.0f - R);
      return (pdf > 0.0f) ? pdf : 0.0f;
    }
    pdf *= ggx_pdf_h_cos(cosh, cos_in, cosr, v.roughness);
    pdf /= fabsf(cos_out);
    mask |= !(pdf > 0.0f);
    return mask ? 0.0f : pdf;
  }
  // metal.c:166-205
  if((KINDS & 4) && (KINDS == 4 || m.bsdf == CB_BSDF_METAL))
  {
    if(!(v.mode & M_REFLECT)) return 0.0f;
    const float cos_in = -dot(v.n, wi), cos_out = dot(v.n, wo);
    if(cos_in < 0.0f || cos_out < 0.0f) return 0.0f;
    const V3 h = normalise(sub(wi, wo));
    if(v.mode & M_SPECULAR) return fabsf(dot(h, v.n)) < HALFVEC_COS_THR ? 0.0f : 1.0f;
    float pdf = 1.0f/(4.0f*fabsf(dot(wo, h)));
    pdf *= ggx_pdf_h(wi, h, v.n, v.roughness);
    pdf /= fabsf(cos_out);
    return pdf > 0.0f ? pdf : 0.0f;
  }
  return 0.0f;
}

// lights_eval_vertex for a path that started at the sensor (list.c:242-275): emitted radiance towards -omega_in
CBD float light_eval(const Vtx &v, V3 omega_in)
{
  if(v.em <= 0.0f) return 0.0f;
  if(dot(v.gn, omega_in) >= 0.0f) return 0.0f;
  float edf;
  if(v.roughness > 1.0f - 1e-4f) edf = (float)(1.0/PI_D);
  else
  {
    const float phongexp = 2.0f/(v.roughness*v.roughness) - 2.0f;
    edf = (float)((double)(powf(fabsf(dot(v.gn, omega_in)), phongexp)*(phongexp + 2.0f))/(2.0*PI_D));
  }
  return edf*v.em;
}

// spectrum_p_to_camera (spectrum.h:172-203): one wavelength -> camera space colour
#include "cie1931.h"
__constant__ float c_cie[CIE_ROWS*3];
CBD void spectrum_to_camera(float lambda, float p, int colour, float col[3])
{
  float f = (lambda - CIE_LAMBDA_MIN)/CIE_STEP;
  const int i = (int)f;
  f -= i;
  float xyz[3];
#pragma unroll
  for(int k=0;k<3;k++) xyz[k] = ((1.0f - f)*c_cie[3*i+k] + f*c_cie[3*(i+1)+k])*p;
  if(colour == CB_COLOUR_REC709)
  { // include/colour/rec709.h
    col[0] =  3.2404542f*xyz[0] - 1.5371385f*xyz[1] - 0.4985314f*xyz[2];
    col[1] = -0.9692660f*xyz[0] + 1.8760108f*xyz[1] + 0.0415560f*xyz[2];
    col[2] =  0.0556434f*xyz[0] - 0.2040259f*xyz[1] + 1.0572252f*xyz[2];
  }
  else { col[0] = xyz[0]; col[1] = xyz[1]; col[2] = xyz[2]; }
}
