/* obj2geo_b200.c -- Wavefront .obj -> corona-13 .geo, the ingestion step in front of the hot path.
 *
 *   obj2geo_b200 input.obj [input_shutter_close.obj] [line radius (default 0.001)]
 *
 * Same command line, same outputs as the reference's tools/geo/obj2geo.c (one <object>.geo per `o` statement, the first one named
 * after the input file), written from the file format down (include/prims.h:20-47, include/geo.h:46-94):
 *
 *   header   { magic 0xc01337, version 2, num_prims, vtxidx_offset, vertex_offset }                   prims_header_t
 *   primid   num_prims x 64 bit  { extra:3 shapeid:29 | vi:28 mb:1 vcnt:3 }                           primid_t, shapeid 0
 *   vtxidx   per primitive corner { v: index into vtx (shared between faces), uv: two halfs }          prims_vtxidx_t
 *   vtx      per vertex (x2 interleaved with motion) { float v[3]; uint32 n: octahedral normal / line radius }   prims_vtx_t
 *
 * Behaviour kept from the reference tool because scenes depend on it: faces are triangles / quads / two-point lines (`l`), indices
 * may be negative (relative to the END of the list), a vertex is shared between faces of one object when its encoded normal agrees,
 * objects without normals get area-weighted vertex normals (geo.h:176-235), the vertex section starts 16-byte aligned and the gap
 * in front of it repeats the first bytes of the vertex array, hair strands get (strand, 0, position) texture coordinates.
 * The converter builds everything in memory in one pass over growing arrays instead of counting first.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "corona_types.h"        /* CB_GEO_MAGIC / CB_GEO_VERSION */

typedef struct { float v[3]; uint32_t n; } vtx_t;               /* prims_vtx_t */
typedef struct { uint32_t v, uv; } vtxidx_t;                    /* prims_vtxidx_t */
typedef struct { uint32_t magic, version; uint64_t num_prims, vtxidx_offset, vertex_offset; } header_t;   /* prims_header_t */

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* geo_encode_normal, include/geo.h:46-74: octahedral map, sign + 15 mantissa bits per component */
static uint32_t encode_normal(const float vec[3])
{
  const float inv = 1.0f/(fabsf(vec[0]) + fabsf(vec[1]) + fabsf(vec[2]));
  float e0, e1;
  if(vec[2] < 0.0f)
  {
    e0 = (1.0f - fabsf(vec[1]*inv))*((vec[0] < 0.0f) ? -1.0f : 1.0f);
    e1 = (1.0f - fabsf(vec[0]*inv))*((vec[1] < 0.0f) ? -1.0f : 1.0f);
  }
  else { e0 = vec[0]*inv; e1 = vec[1]*inv; }
  const uint32_t i0 = f2u((fabsf(e0) + 2.0f)/2.0f), i1 = f2u((fabsf(e1) + 2.0f)/2.0f);
  uint16_t p0 = (uint16_t)(((f2u(e0) & 0x80000000u) >> 16) | ((i0 & 0x7fffffu) >> 8));
  uint16_t p1 = (uint16_t)(((f2u(e1) & 0x80000000u) >> 16) | ((i1 & 0x7fffffu) >> 8));
  if((p0 & 0x7fff) == 0) p0 = 0;
  if((p1 & 0x7fff) == 0) p1 = 0;
  return (uint32_t)p0 | ((uint32_t)p1 << 16);
}

/* float_to_half, include/half.h:31-56 (truncating; values are clamped to 65536 and scaled by 2^-112 before the shift) */
static uint16_t to_half(float x)
{
  uint32_t u = f2u(x);
  const uint32_t sign = u & 0x80000000u;
  u ^= sign;
  float f = u2f(u);
  if(f < 2139095040.0f)
  {
    if(f > 65536.0f) f = 65536.0f;
    f *= u2f(15u << 23);
  }
  return (uint16_t)((f2u(f) >> 13) | (sign >> 16));
}
static uint32_t encode_uv(float u, float v) { return (uint32_t)to_half(u) | ((uint32_t)to_half(v) << 16); }
static int clampi(float x, int lo, int hi) { return x < lo ? lo : x > hi ? hi : (int)x; }
static uint32_t encode_uvw(float u, float v, float w)          /* geo.h:91-94 */
{
  return ((uint32_t)clampi(u*2048, 0, 2047) << 21) | ((uint32_t)clampi(v*2048, 0, 2047) << 10) | (uint32_t)clampi(w*1024, 0, 1023);
}

/* ---- the parsed .obj ------------------------------------------------------------------------------------------------ */
typedef struct { float *v, *n, *t; size_t nv, nn, nt; } lists_t;
typedef struct { char **line; size_t num; char *text; } lines_t;

static int read_lines(const char *name, lines_t *L)
{
  FILE *f = fopen(name, "rb");
  if(!f) return 1;
  fseek(f, 0, SEEK_END);
  const long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  L->text = (char *)malloc((size_t)size + 2);
  if(!L->text || fread(L->text, 1, (size_t)size, f) != (size_t)size) { fclose(f); return 1; }
  fclose(f);
  L->text[size] = '\n'; L->text[size+1] = 0;
  size_t cap = 1024;
  L->line = (char **)malloc(cap*sizeof(char *));
  L->num = 0;
  for(char *p = L->text; p < L->text + size; )
  {
    char *e = strchr(p, '\n');
    *e = 0;
    if(e > p && e[-1] == '\r') e[-1] = 0;
    if(L->num == cap) { cap *= 2; L->line = (char **)realloc(L->line, cap*sizeof(char *)); }
    L->line[L->num++] = p;
    p = e + 1;
  }
  return 0;
}

static void read_lists(const lines_t *L, lists_t *o)
{
  memset(o, 0, sizeof(*o));
  size_t cv = 0, cn = 0, ct = 0;
  for(size_t k=0;k<L->num;k++)
  {
    const char *s = L->line[k];
    float a[3];
    if(!strncmp(s, "vn ", 3))
    {
      if(sscanf(s, "vn %f %f %f", a, a+1, a+2) != 3) { fprintf(stderr, "line %zu: weird normal: `%s'\n", k + 1, s); continue; }
      if(o->nn == cn) { cn = cn ? 2*cn : 1024; o->n = (float *)realloc(o->n, cn*3*sizeof(float)); }
      memcpy(o->n + 3*o->nn++, a, 12);
    }
    else if(!strncmp(s, "vt ", 3))
    {
      if(sscanf(s, "vt %f %f", a, a+1) != 2) { fprintf(stderr, "line %zu: weird uvs: `%s'\n", k + 1, s); continue; }
      if(o->nt == ct) { ct = ct ? 2*ct : 1024; o->t = (float *)realloc(o->t, ct*2*sizeof(float)); }
      memcpy(o->t + 2*o->nt++, a, 8);
    }
    else if(!strncmp(s, "v ", 2))
    {
      if(sscanf(s, "v %f %f %f", a, a+1, a+2) != 3) { fprintf(stderr, "line %zu: weird vertex: `%s'\n", k + 1, s); continue; }
      if(o->nv == cv) { cv = cv ? 2*cv : 1024; o->v = (float *)realloc(o->v, cv*3*sizeof(float)); }
      memcpy(o->v + 3*o->nv++, a, 12);
    }
  }
}

/* one face / line statement: up to four corners of (vertex, texture, normal) references, zero based, -1 = absent.
 * returns the number of corners (2, 3, 4) or 0 */
static int parse_face(const char *s, long nv, long nt, long nn, int vert[4], int uvco[4], int norm[4])
{
  int slashes = 0, dbl = 0;
  for(const char *c = s; *c; c++) if(*c == '/') { slashes++; if(c[1] == '/') dbl = 1; }
  int v[4] = {0}, t[4] = {0}, n[4] = {0}, cnt = 0, per = 1;
  if(slashes >= 6)
  {
    if(dbl) { cnt = sscanf(s + 1, " %d//%d %d//%d %d//%d %d//%d", v, n, v+1, n+1, v+2, n+2, v+3, n+3); per = 2; }
    else    { cnt = sscanf(s + 1, " %d/%d/%d %d/%d/%d %d/%d/%d %d/%d/%d", v, t, n, v+1, t+1, n+1, v+2, t+2, n+2, v+3, t+3, n+3); per = 3; }
  }
  else if(slashes >= 3) { cnt = sscanf(s + 1, " %d/%d %d/%d %d/%d %d/%d", v, t, v+1, t+1, v+2, t+2, v+3, t+3); per = 2; }
  else                  { cnt = sscanf(s + 1, " %d %d %d %d", v, v+1, v+2, v+3); per = 1; }
  for(int k=0;k<4;k++)
  { /* obj indices start at 1, negative ones count back from the end of the list (to_zero_base_idx) */
    vert[k] = v[k] < 0 ? v[k] + (int)nv : v[k] - 1;
    uvco[k] = t[k] < 0 ? t[k] + (int)nt : t[k] - 1;
    norm[k] = n[k] < 0 ? n[k] + (int)nn : n[k] - 1;
  }
  /* the reference keys on the raw sscanf count: 12/8/4 quad, 9/6/3 triangle, 2 line (an `l a b`) */
  if(cnt == 12 || cnt == 8 || cnt == 4) return 4;
  if(cnt == 9 || cnt == 6 || cnt == 3) return 3;
  if(cnt == 2) return 2;
  (void)per;
  return 0;
}

/* ---- one output shape ------------------------------------------------------------------------------------------------ */
typedef struct
{
  uint64_t *primid; size_t num_prims, cap_prims;
  vtxidx_t *vtxidx; size_t num_vtxidx, cap_vtxidx;
  vtx_t *vtx; size_t num_verts, cap_verts;       /* num_verts vertices, `stride` records each */
  int stride;
}
shape_t;

/* area-weighted vertex normals from the face normals around each vertex (geo_get_normal + geo_recompute_normals, geo.h:176-235) */
static void recompute_normals(shape_t *S)
{
  const int mb = S->stride;
  float *acc = (float *)calloc(S->num_verts*mb*3 + 3, sizeof(float));
  for(size_t p=0;p<S->num_prims;p++)
  {
    const uint32_t hi = (uint32_t)(S->primid[p] >> 32);
    const int vcnt = (int)(hi >> 29), vi = (int)(hi & 0x0fffffffu);
    if(vcnt != 3 && vcnt != 4) continue;
    for(int k=0;k<vcnt;k++) for(int m=0;m<mb;m++)
    {
      const int i0 = vcnt == 3 ? 0 : (k == 0 ? 3 : k - 1), i1 = vcnt == 3 ? 1 : k, i2 = vcnt == 3 ? 2 : (k == 3 ? 0 : k + 1);
      const float *v0 = S->vtx[mb*S->vtxidx[vi+i0].v + m].v, *v1 = S->vtx[mb*S->vtxidx[vi+i1].v + m].v, *v2 = S->vtx[mb*S->vtxidx[vi+i2].v + m].v;
      const float n[3] = {(v1[1] - v0[1])*(v2[2] - v0[2]) - (v1[2] - v0[2])*(v2[1] - v0[1]),
                          (v1[2] - v0[2])*(v2[0] - v0[0]) - (v1[0] - v0[0])*(v2[2] - v0[2]),
                          (v1[0] - v0[0])*(v2[1] - v0[1]) - (v1[1] - v0[1])*(v2[0] - v0[0])};
      float *a = acc + 3*(mb*S->vtxidx[vi+k].v + m);
      for(int i=0;i<3;i++) a[i] += n[i];
    }
  }
  for(size_t i=0;i<S->num_verts*mb;i++) S->vtx[i].n = encode_normal(acc + 3*i);
  free(acc);
}

static int write_shape(shape_t *S, const char *name, int recompute)
{
  if(!S->num_prims) return 0;
  if(recompute) recompute_normals(S);
  char file[1100];
  snprintf(file, sizeof(file), "%s.geo", name);
  FILE *o = fopen(file, "wb");
  if(!o) { fprintf(stderr, "could not write `%s'\n", file); return 1; }
  header_t h;
  memset(&h, 0, sizeof(h));
  h.magic = CB_GEO_MAGIC;       /* include/geo.h:3-4 */
  h.version = CB_GEO_VERSION;
  h.num_prims = S->num_prims;
  uint64_t cnt = sizeof(header_t) + 8*S->num_prims;
  h.vtxidx_offset = cnt;
  cnt += sizeof(vtxidx_t)*S->num_vtxidx;
  h.vertex_offset = (cnt + 0xf) & ~(uint64_t)0xf;
  fwrite(&h, sizeof(h), 1, o);
  fwrite(S->primid, 8, S->num_prims, o);
  fwrite(S->vtxidx, sizeof(vtxidx_t), S->num_vtxidx, o);
  if(h.vertex_offset > cnt) fwrite(S->vtx, h.vertex_offset - cnt, 1, o);   /* the gap repeats the head of the vertex array, like upstream */
  fwrite(S->vtx, sizeof(vtx_t)*S->stride, S->num_verts, o);
  fclose(o);
  return 0;
}

int main(int argc, char *argv[])
{
  if(argc < 2) { fprintf(stderr, "usage: %s input.obj [input_shutter_close.obj] [line radius (def 0.001)]\n", argv[0]); return 1; }
  lines_t A, B;
  memset(&B, 0, sizeof(B));
  if(read_lines(argv[1], &A)) { fprintf(stderr, "could not open `%s'\n", argv[1]); return 1; }
  int motion = 0;
  if(argc > 2 && !read_lines(argv[2], &B)) { fprintf(stderr, "given two obj files, interpreting as motion blurred geo!\n"); motion = 1; }
  const int stride = motion ? 2 : 1;
  lists_t a, b;
  read_lists(&A, &a);
  memset(&b, 0, sizeof(b));
  if(motion) read_lists(&B, &b);
  fprintf(stderr, "num verts %zu num normals %zu num texture coords %zu\n", a.nv, a.nn, a.nt);
  const int have_normals = a.nn != 0;
  int recompute = !have_normals;
  const float radius = argc > (2 + motion) ? (float)atof(argv[argc-1]) : 0.001f;
  const uint32_t radius_bits = f2u(radius);

  char shapename[1100];
  snprintf(shapename, sizeof(shapename), "%s", argv[1]);
  for(char *c = shapename + strlen(shapename); c > shapename; c--) if(*c == '.') { *c = 0; break; }

  shape_t S;
  memset(&S, 0, sizeof(S));
  S.stride = stride;
  uint32_t *vmap = (uint32_t *)malloc(sizeof(uint32_t)*(a.nv ? a.nv : 1));
  memset(vmap, 0xff, sizeof(uint32_t)*(a.nv ? a.nv : 1));
  size_t lineB = 0;
  long first_hair = -1, last_hair = -1; uint32_t hair_index = 0;
  int rc = 0;

  for(size_t k=0;k<A.num;k++)
  {
    const char *s = A.line[k];
    if(!strncmp(s, "o ", 2))
    { /* a new object: flush the shape so far, start over with its name */
      rc |= write_shape(&S, shapename, recompute);
      recompute = !have_normals;
      snprintf(shapename, sizeof(shapename), "%s", strchr(s, ' ') + 1);
      fprintf(stderr, "object `%s'\n", shapename);
      S.num_prims = S.num_vtxidx = S.num_verts = 0;
      memset(vmap, 0xff, sizeof(uint32_t)*(a.nv ? a.nv : 1));
      continue;
    }
    if(strncmp(s, "f ", 2) && strncmp(s, "l ", 2)) continue;
    const char *s2 = 0;
    if(motion)
    { /* the matching statement of the shutter-close file */
      while(lineB < B.num && strncmp(B.line[lineB], "f ", 2) && strncmp(B.line[lineB], "l ", 2)) lineB++;
      if(lineB == B.num) { fprintf(stderr, "error: premature end of motion obj!\n"); return 14; }
      s2 = B.line[lineB++];
    }
    int vert[4], uvco[4], norm[4], vert2[4] = {0}, uv2[4], norm2[4] = {0};
    const int vcnt = parse_face(s, (long)a.nv, (long)a.nt, (long)(have_normals ? a.nn : a.nv), vert, uvco, norm);
    if(!vcnt) { fprintf(stderr, "only lines, tris and quads supported so far\n"); continue; }
    if(motion && parse_face(s2, (long)b.nv, (long)b.nt, (long)b.nn, vert2, uv2, norm2) != vcnt)
      fprintf(stderr, "error: motion prim with different vertex count!\n");
    int bad = 0;
    for(int c=0;c<vcnt;c++) if(vert[c] < 0 || (size_t)vert[c] >= a.nv || (motion && (vert2[c] < 0 || (size_t)vert2[c] >= b.nv))) bad = 1;
    if(bad) { fprintf(stderr, "line %zu: vertex index out of range, face dropped: `%s'\n", k + 1, s); continue; }

    if(S.num_prims == S.cap_prims) { S.cap_prims = S.cap_prims ? 2*S.cap_prims : 4096; S.primid = (uint64_t *)realloc(S.primid, 8*S.cap_prims); }
    if(S.num_vtxidx + 4 > S.cap_vtxidx) { S.cap_vtxidx = S.cap_vtxidx ? 2*S.cap_vtxidx : 16384; S.vtxidx = (vtxidx_t *)realloc(S.vtxidx, sizeof(vtxidx_t)*S.cap_vtxidx); }
    /* primid_t: low word extra:3 | shapeid:29 = 0, high word vi:28 | mb:1 | vcnt:3 (include/prims.h:20-30) */
    S.primid[S.num_prims++] = ((uint64_t)(((uint32_t)S.num_vtxidx & 0x0fffffffu) | ((uint32_t)motion << 28) | ((uint32_t)vcnt << 29))) << 32;
    for(int c=0;c<vcnt;c++)
    {
      uint32_t n0 = 0, n1 = 0;
      if(vcnt <= 2) { n0 = radius_bits; if(motion) n1 = radius_bits; }        /* lines carry radii where faces carry normals */
      else if(!recompute && norm[c] >= 0 && (size_t)norm[c] < a.nn)
      {
        n0 = encode_normal(a.n + 3*norm[c]);
        if(motion) n1 = (norm2[c] >= 0 && (size_t)norm2[c] < b.nn) ? encode_normal(b.n + 3*norm2[c]) : 0x8000;
      }
      else { recompute = 1; n0 = 0x8000; if(motion) n1 = 0x8000; }
      vtxidx_t vid;
      const uint32_t seen = vmap[vert[c]];
      if(seen != 0xffffffffu && (recompute || vcnt <= 2 || (n0 == S.vtx[stride*seen].n && (!motion || n1 == S.vtx[stride*seen + 1].n)))) vid.v = seen;
      else
      {
        if(S.num_verts == S.cap_verts) { S.cap_verts = S.cap_verts ? 2*S.cap_verts : 8192; S.vtx = (vtx_t *)realloc(S.vtx, sizeof(vtx_t)*stride*S.cap_verts); }
        vid.v = (uint32_t)S.num_verts++;
        memcpy(S.vtx[stride*vid.v].v, a.v + 3*vert[c], 12);
        S.vtx[stride*vid.v].n = n0;
        if(motion) { memcpy(S.vtx[stride*vid.v + 1].v, b.v + 3*vert2[c], 12); S.vtx[stride*vid.v + 1].n = n1; }
        vmap[vert[c]] = vid.v;
      }
      vid.uv = (uvco[c] >= 0 && (size_t)uvco[c] < a.nt) ? encode_uv(a.t[2*uvco[c]], a.t[2*uvco[c]+1]) : 0;
      S.vtxidx[S.num_vtxidx++] = vid;
    }
    /* hair: consecutive `l` segments that continue each other form a strand with texture coordinates (strand id, 0, position) */
    if(vcnt == 2)
    {
      if(first_hair == -1 || vert[0] != last_hair)
      {
        if(first_hair != -1)
        {
          for(size_t i=(size_t)first_hair;i<S.num_vtxidx;i++)
            S.vtxidx[i].uv = encode_uvw(hair_index/10000.0f, 0, (i - first_hair)/((float)S.num_vtxidx - first_hair - 1.0f));
          hair_index++;
        }
        first_hair = (long)S.num_vtxidx - 2;
        last_hair = vert[1];
      }
      else last_hair = vert[1];
    }
    else if(first_hair != -1)
    {
      for(size_t i=(size_t)first_hair;i<S.num_vtxidx - vcnt;i++)
        S.vtxidx[i].uv = encode_uvw(hair_index/10000.0f, 0, (i - first_hair)/((float)S.num_vtxidx - vcnt - first_hair - 1.0f));
      first_hair = last_hair = -1;
      hair_index++;
    }
  }
  if(first_hair != -1)
    for(size_t i=(size_t)first_hair;i<S.num_vtxidx;i++)
      S.vtxidx[i].uv = encode_uvw(hair_index/10000.0f, 0, (i - first_hair)/((float)S.num_vtxidx - first_hair - 1.0f));
  rc |= write_shape(&S, shapename, recompute);
  return rc ? 2 : 0;
}
