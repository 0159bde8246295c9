/* scene_b200.c -- the reference's scene ingestion for the hot path, in plain C on top of the module layer:
 *
 *   .nra2   sky line, shader list, shape list          src/shader.c:605-788 (shader_init), src/corona_common.c:30-68
 *   .geo    mapped read-only (prims_load, accel_b200.c)  src/prims.c:759-828
 *   .cam    camera_t (104 B "CCAM" v1) / camera_v0_t      include/camera.h:13-35,77-99,153-196, src/view.c:933-952
 *   rgb -> spectrum coefficients                          include/rgb2spec.h:28-128, include/spectrum.h:29-38
 *
 * The reference dlopens one module per shader line and keeps them as a graph of callbacks (`mult` calls its `pre[]` list,
 * then its `host`); here every shader line is flattened ONCE into a cb_material_t (include/corona_b200_render.h): the
 * prepare() steps of color / colorcheckersg in order, then the host BSDF diffuse / dielectric / metal.  A shape that uses
 * any other kind of shader makes cb200_render_create fail: there is no CPU fallback.
 *
 * The measured tables two shader modules carry as C initialisers (ColorChecker SG reflectances, metal n/k) are DATA of the
 * reference; they come from a small table file written by tests/golden/make_tables.py (format at tables_load below).
 */
#include "corona_host.h"
#include "corona_b200.h"
#include "corona_b200_render.h"

#include <math.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>
#include <strings.h>

/* ---------------------------------------------------------------------------------------------- rgb2spec */
typedef struct rgb2spec_b200_t { uint32_t res; float *scale, *data; } rgb2spec_b200_t;

static rgb2spec_b200_t *rgb2spec_b200_load(const char *filename)
{ /* rgb2spec_init, include/rgb2spec.h:28-64 */
  FILE *f = fopen(filename, "rb");
  if(!f) return 0;
  char header[4];
  rgb2spec_b200_t *m = (rgb2spec_b200_t *)calloc(1, sizeof(*m));
  if(fread(header, 4, 1, f) != 1 || memcmp(header, "SPEC", 4) || fread(&m->res, 4, 1, f) != 1 || m->res < 2 || m->res > 256)
  { fclose(f); free(m); return 0; }   /* the fetch interpolates between res-2 cells; a wild res would overflow the table size */
  const size_t ns = m->res, nd = (size_t)m->res*m->res*m->res*3*3;
  m->scale = (float *)malloc(sizeof(float)*ns);
  m->data = (float *)malloc(sizeof(float)*nd);
  if(!m->scale || !m->data || fread(m->scale, sizeof(float)*ns, 1, f) != 1 || fread(m->data, sizeof(float)*nd, 1, f) != 1)
  { fclose(f); free(m->scale); free(m->data); free(m); return 0; }
  fclose(f);
  return m;
}
static void rgb2spec_b200_free(rgb2spec_b200_t *m) { if(m) { free(m->scale); free(m->data); free(m); } }

static void rgb2spec_b200_fetch(const rgb2spec_b200_t *model, const float rgb[3], float out[3])
{ /* rgb2spec_fetch, include/rgb2spec.h:87-128, same operation order */
  int i = 0;
  const int res = (int)model->res;
  for(int j=1;j<3;j++) if(rgb[j] >= rgb[i]) i = j;
  const float z = rgb[i];
  if(!(z > 0.0f)) { out[0] = out[1] = out[2] = NAN; return; }   /* black: 0*inf = NaN there as well; clamps to 0 downstream */
  const float scale = (float)(res - 1)/z, x = rgb[(i+1)%3]*scale, y = rgb[(i+2)%3]*scale;
  uint32_t xi = (uint32_t)x, yi = (uint32_t)y;
  if(xi > (uint32_t)res - 2) xi = res - 2;
  if(yi > (uint32_t)res - 2) yi = res - 2;
  int left = 0, last = res - 2, size = last;                      /* rgb2spec_find_interval */
  while(size > 0)
  {
    const int half = size >> 1, mid = left + half + 1;
    if(model->scale[mid] < z) { left = mid; size -= half + 1; } else size = half;
  }
  const uint32_t zi = left < last ? left : last;
  uint32_t off = (((i*res + zi)*res + yi)*res + xi)*3;
  const uint32_t dx = 3, dy = 3*res, dz = 3*res*res;
  const float x1 = x - xi, x0 = 1.0f - x1, y1 = y - yi, y0 = 1.0f - y1,
              z1 = (z - model->scale[zi])/(model->scale[zi+1] - model->scale[zi]), z0 = 1.0f - z1;
  const float *D = model->data;
  for(int j=0;j<3;j++, off++)
    out[j] = ((D[off] * x0 + D[off+dx] * x1) * y0 + (D[off+dy] * x0 + D[off+dy+dx] * x1) * y1) * z0 +
             ((D[off+dz] * x0 + D[off+dz+dx] * x1) * y0 + (D[off+dz+dy] * x0 + D[off+dz+dy+dx] * x1) * y1) * z1;
}

static float rgb_to_coeff(const rgb2spec_b200_t *m, const float rgb[3], float out[3])
{ /* spectrum_rgb_to_coeff, include/spectrum.h:29-38 */
  float mul = rgb[0] > rgb[1] ? rgb[0] : rgb[1];
  if(rgb[2] > mul) mul = rgb[2];
  if(mul == 0.0f || mul < 1.0f) mul = 1.0f;
  const float col[3] = { rgb[0]/mul, rgb[1]/mul, rgb[2]/mul };
  rgb2spec_b200_fetch(m, col, out);
  return mul;
}

/* ---------------------------------------------------------------------------------------------- measured tables
 * file: "CBT1", uint32 count, then per table { char name[32]; uint32 rows, cols; float lambda_min, lambda_step; float data[rows*cols] } */
typedef struct table_file_t { uint32_t count; char (*name)[32]; cb_table_t *tab; float **data; } table_file_t;

static int tables_load(table_file_t *t, const char *filename)
{
  memset(t, 0, sizeof(*t));
  FILE *f = fopen(filename, "rb");
  if(!f) return 1;
  char magic[4];
  uint32_t count = 0;
  if(fread(magic, 4, 1, f) != 1 || memcmp(magic, "CBT1", 4) || fread(&count, 4, 1, f) != 1 || count > 1024) { fclose(f); return 1; }
  t->count = count;   /* (only once it is known to be sane: tables_free walks `count` entries) */
  t->name = calloc(t->count ? t->count : 1, 32);
  t->tab = calloc(t->count ? t->count : 1, sizeof(cb_table_t));
  t->data = calloc(t->count ? t->count : 1, sizeof(float *));
  if(!t->name || !t->tab || !t->data) { fclose(f); return 1; }
  for(uint32_t k=0;k<t->count;k++)
  {
    uint32_t rc[2]; float lm[2];
    if(fread(t->name[k], 32, 1, f) != 1 || fread(rc, 8, 1, f) != 1 || fread(lm, 8, 1, f) != 1) { fclose(f); return 1; }
    t->name[k][31] = 0;
    if(rc[0] == 0 || rc[1] == 0 || (uint64_t)rc[0]*rc[1] > (1u << 24)) { fclose(f); return 1; }   /* 140 x 36 and 3 x n in practice */
    t->data[k] = malloc(sizeof(float)*(size_t)rc[0]*rc[1]);
    if(!t->data[k] || fread(t->data[k], sizeof(float)*(size_t)rc[0]*rc[1], 1, f) != 1) { fclose(f); return 1; }
    t->tab[k].rows = rc[0]; t->tab[k].num_lambda = rc[1]; t->tab[k].lambda_min = lm[0]; t->tab[k].lambda_step = lm[1];
    t->tab[k].data = t->data[k];
  }
  fclose(f);
  return 0;
}
static void tables_free(table_file_t *t)
{
  for(uint32_t k=0;t->data && k<t->count;k++) free(t->data[k]);
  free(t->data); free(t->tab); free(t->name);
  memset(t, 0, sizeof(*t));
}
static int tables_find(const table_file_t *t, const char *name)
{
  for(uint32_t k=0;k<t->count;k++) if(!strcasecmp(t->name[k], name)) return (int)k;
  return -1;
}

/* ---------------------------------------------------------------------------------------------- shader list */
#define MAX_SHADERS 256
typedef struct shader_line_t { char name[64]; char args[448]; } shader_line_t;

typedef struct nra2_t
{
  int num_shaders;
  shader_line_t line[MAX_SHADERS];
  const rgb2spec_b200_t *rgb2spec;
  const table_file_t *tables;
  /* tables actually referenced, in cb_render_desc_t order */
  cb_table_t used[MAX_SHADERS];
  int used_src[MAX_SHADERS];
  int num_used;
  /* homogeneous media, in cb_render_desc_t order; medium_of[k] = 1 + index for shader k once flattened */
  cb_medium_t media[CB_MAX_MEDIA];
  int medium_of[MAX_SHADERS];
  int num_media;
  int exterior_medium;
  int flatten_calls;          /* budget of flatten() per material */
}
nra2_t;

static int use_table(nra2_t *n, int src)
{
  for(int k=0;k<n->num_used;k++) if(n->used_src[k] == src) return k;
  n->used[n->num_used] = n->tables->tab[src];
  n->used_src[n->num_used] = src;
  return n->num_used++;
}

static int parse_slot(char c)
{ /* tex_parse_slot, src/shaders/texture.h:23-35 */
  switch(c)
  {
    case 'd': return CB_SLOT_DIFFUSE;   case 's': return CB_SLOT_SPECULAR; case 'e': return CB_SLOT_EMISSION;
    case 'v': return CB_SLOT_VOLUME;    case 'g': return CB_SLOT_GLOSSY;   case 'r': return CB_SLOT_ROUGHNESS;
    case 't': return CB_SLOT_TRANSMIT_TO_EYE;
    default:  return -1;
  }
}

/* contribution of shader k to a material: its prepare() steps appended to m->ops, its bsdf (if it has one) to *bsdf.
 * returns 0, or 1 when the shader is outside the pt/ptdl surface path */
static int flatten(nra2_t *n, int k, cb_material_t *m, int *have_bsdf, int depth)
{
  /* calls are budgeted per material: a list of `mult 16 -1 ... -1` lines would otherwise fan out 16^depth times */
  if(k < 0 || k >= n->num_shaders || depth > 16 || ++n->flatten_calls > 4096) return 1;
  const shader_line_t *l = n->line + k;
  if(!strcmp(l->name, "diffuse")) { m->bsdf = CB_BSDF_DIFFUSE; *have_bsdf = 1; return 0; }
  if(!strcmp(l->name, "color"))
  { /* src/shaders/color.c:36-61: " %c %f %f %f %f" */
    char c; float col[3], rough = 1.0f;
    if(sscanf(l->args, " %c %f %f %f %f", &c, col, col+1, col+2, &rough) < 4 || parse_slot(c) < 0 || m->num_ops >= CB_MAX_MATOPS) return 1;
    cb_matop_t *op = m->ops + m->num_ops++;
    memset(op, 0, sizeof(*op));
    op->op = CB_OP_COLOR; op->slot = parse_slot(c);
    op->mul = rgb_to_coeff(n->rgb2spec, col, op->coeff);
    op->roughness = rough;
    return 0;
  }
  if(!strcmp(l->name, "colorcheckersg"))
  { /* src/shaders/colorcheckersg.c:213-228: " %c %f" */
    char c; float rough = 1.0f;
    const int t = n->tables ? tables_find(n->tables, "checker") : -1;
    if(sscanf(l->args, " %c %f", &c, &rough) < 1 || parse_slot(c) < 0 || t < 0 || m->num_ops >= CB_MAX_MATOPS) return 1;
    cb_matop_t *op = m->ops + m->num_ops++;
    memset(op, 0, sizeof(*op));
    op->op = CB_OP_CHECKERSG; op->slot = parse_slot(c); op->roughness = rough; op->table = use_table(n, t);
    return 0;
  }
  if(!strcmp(l->name, "dielectric"))
  { /* src/shaders/dielectric.c: " %f %f", Abbe number defaults to 50 */
    float nd = 1.5f, vd = 50.0f;
    if(sscanf(l->args, " %f %f", &nd, &vd) < 1) return 1;
    m->bsdf = CB_BSDF_DIELECTRIC; m->param[0] = nd; m->param[1] = vd; *have_bsdf = 1;
    return 0;
  }
  if(!strcmp(l->name, "diffdiel"))
  { /* src/shaders/diffdiel.c:39-58: same arguments as dielectric */
    float nd = 1.5f, vd = 50.0f;
    if(sscanf(l->args, " %f %f", &nd, &vd) < 1) return 1;
    m->bsdf = CB_BSDF_DIFFDIEL; m->param[0] = nd; m->param[1] = vd; *have_bsdf = 1;
    return 0;
  }
  if(!strcmp(l->name, "metal"))
  { /* src/shaders/metal.c:43-64: material name looked up in fresnel.h's list; unknown names fall back to the first (Ti) */
    char mat[64] = "", tn[80];
    if(sscanf(l->args, " %63s", mat) < 1 || !n->tables) return 1;
    for(char *p=mat;*p;p++) if(*p >= 'A' && *p <= 'Z') *p += 'a' - 'A';
    snprintf(tn, sizeof(tn), "metal_%s", mat);
    int t = tables_find(n->tables, tn);
    if(t < 0) { fprintf(stderr, "[metal] WARNING: didn't find `%s' in material list!\n", mat); t = tables_find(n->tables, "metal_ti"); }
    if(t < 0) return 1;
    m->bsdf = CB_BSDF_METAL; m->table = use_table(n, t); *have_bsdf = 1;
    return 0;
  }
  if(!strcmp(l->name, "mult"))
  { /* src/shaders/mult.c:90-128: "<n> <pre...> <host>", negative numbers are relative to this shader's own index */
    int cnt = 0, pos = 0, adv = 0;
    if(sscanf(l->args, " %d%n", &cnt, &adv) < 1 || cnt < 0 || cnt > 16) return 1;
    pos = adv;
    int idx[17];
    for(int i=0;i<=cnt;i++)
    {
      if(sscanf(l->args + pos, " %d%n", idx + i, &adv) < 1) return 1;
      pos += adv;
      if(idx[i] < 0) idx[i] += k;
      if(idx[i] < 0 || idx[i] >= k) return 1;   /* upstream resolves references against the shaders loaded so far (mult.c:103-120): earlier lines only */
    }
    for(int i=0;i<cnt;i++)
    { /* pre slots only run prepare(): a BSDF-type shader there contributes no callbacks upstream, so it is refused here rather than
       * allowed to overwrite the host's */
      int hb = 0;
      const int bsdf0 = m->bsdf;
      if(flatten(n, idx[i], m, &hb, depth+1) || hb) return 1;
      m->bsdf = bsdf0;
    }
    return flatten(n, idx[cnt], m, have_bsdf, depth+1);
  }
  return 1;   /* skies, hair, heterogeneous media, ...: SURVEY 2.1 marks them outside the hot path */
}

/* shader k as a homogeneous medium: `medium_rgb r g b g` (src/shaders/medium_rgb.c:112-139) or a mult whose pre steps are
 * `color v` and whose host is one.  Returns 1 + index into n->media, 0 when k is something else. */
static int flatten_medium(nra2_t *n, int k)
{
  if(k < 0 || k >= n->num_shaders) return 0;
  if(n->medium_of[k]) return n->medium_of[k];
  const shader_line_t *l = n->line + k;
  cb_medium_t m;
  memset(&m, 0, sizeof(m));
  int host = k;
  if(!strcmp(l->name, "mult"))
  {
    int cnt = 0, pos = 0, adv = 0, idx[17];
    if(sscanf(l->args, " %d%n", &cnt, &adv) < 1 || cnt < 0 || cnt > 16) return 0;
    pos = adv;
    for(int i=0;i<=cnt;i++)
    {
      if(sscanf(l->args + pos, " %d%n", idx + i, &adv) < 1) return 0;
      pos += adv;
      if(idx[i] < 0) idx[i] += k;
      if(idx[i] < 0 || idx[i] >= n->num_shaders) return 0;
    }
    for(int i=0;i<cnt;i++)
    {
      const shader_line_t *p = n->line + idx[i];
      char c; float col[3];
      if(strcmp(p->name, "color") || sscanf(p->args, " %c %f %f %f", &c, col, col+1, col+2) < 4 || c != 'v') return 0;
      m.has_albedo = 1;
      m.albedo_mul = rgb_to_coeff(n->rgb2spec, col, m.albedo_coeff);
    }
    host = idx[cnt];
  }
  const shader_line_t *h = n->line + host;
  float mfp[3], g = 0.0f;
  if(strcmp(h->name, "medium_rgb") || sscanf(h->args, " %f %f %f %f", mfp, mfp+1, mfp+2, &g) != 4) return 0;
  float mu_t[3];
  for(int i=0;i<3;i++) mu_t[i] = 1.0f/mfp[i];   /* mean free path -> collision coefficient */
  m.mu_t_mul = rgb_to_coeff(n->rgb2spec, mu_t, m.mu_t_coeff);
  m.g = g;
  if(n->num_media >= CB_MAX_MEDIA) return 0;
  n->media[n->num_media++] = m;
  return n->medium_of[k] = n->num_media;
}

/* ---------------------------------------------------------------------------------------------- camera */
typedef struct { float w, x[3]; } quat_file_t;
typedef struct
{
  char magic[4]; int32_t version; float pos[3], pos_t1[3]; quat_file_t orient, orient_t1; float speed, focus_sensor_offset, focus,
  film_width, film_height, crop_factor; int32_t aperture_value, exposure_value; float focal_length, iso;
}
camera_file_t;        /* include/camera.h:13-35 */
typedef struct
{
  int32_t legacy0; float pos[3]; quat_file_t orient; float speed; int32_t legacy1[7]; float iso; quat_file_t orient_t1; float pos_t1[3];
  float focus_sensor_offset, fill[4], focus, legacy2, crop_factor, film_width, film_height; int32_t aperture_value; float focal_length,
  legacy3; int32_t exposure_value;
}
camera_file_v0_t;     /* include/camera.h:77-99 */

#define VIEW_FULL_FRAME_WIDTH 0.35f   /* src/view.c:70 */
#define VIEW_NUM_EXPOSURE 20          /* src/view.c:75-79 */

int scene_b200_read_camera(const char *filename, uint32_t width, uint32_t height, cb_camera_t *out)
{
  _Static_assert(sizeof(camera_file_t) == 104 && sizeof(camera_file_v0_t) == 152, "camera file layouts");
  FILE *f = fopen(filename, "rb");
  if(!f) return 1;
  fseek(f, 0, SEEK_END);
  const long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  float crop = 1.0f;   /* cam_init default; camera_read does not take the legacy file's crop factor over (camera.h:168-181) */
  memset(out, 0, sizeof(*out));
  if(size == (long)sizeof(camera_file_v0_t))
  {
    camera_file_v0_t c;
    if(fread(&c, sizeof(c), 1, f) != 1) { fclose(f); return 1; }
    memcpy(out->pos, c.pos, 12); memcpy(out->pos_t1, c.pos_t1, 12);
    out->orient[0] = c.orient.w; memcpy(out->orient + 1, c.orient.x, 12);
    out->orient_t1[0] = c.orient_t1.w; memcpy(out->orient_t1 + 1, c.orient_t1.x, 12);
    out->focus = c.focus; out->aperture_value = c.aperture_value; out->exposure_value = c.exposure_value;
    out->focal_length = c.focal_length; out->iso = c.iso;
  }
  else if(size == (long)sizeof(camera_file_t))
  {
    camera_file_t c;
    if(fread(&c, sizeof(c), 1, f) != 1 || strncmp(c.magic, "CCAM", 4) || c.version != 1) { fclose(f); return 1; }
    memcpy(out->pos, c.pos, 12); memcpy(out->pos_t1, c.pos_t1, 12);
    out->orient[0] = c.orient.w; memcpy(out->orient + 1, c.orient.x, 12);
    out->orient_t1[0] = c.orient_t1.w; memcpy(out->orient_t1 + 1, c.orient_t1.x, 12);
    out->focus = c.focus; out->aperture_value = c.aperture_value; out->exposure_value = c.exposure_value;
    out->focal_length = c.focal_length; out->iso = c.iso; crop = c.crop_factor;
  }
  else { fclose(f); return 1; }
  fclose(f);
  /* view_cam_read, src/view.c:933-952 */
  if(out->exposure_value < 0 || out->exposure_value > VIEW_NUM_EXPOSURE) out->exposure_value = 13;
  if(width > height)
  {
    out->film_width = VIEW_FULL_FRAME_WIDTH/crop;
    out->film_height = (float)height/(float)width*out->film_width;
  }
  else
  {
    out->film_height = VIEW_FULL_FRAME_WIDTH/crop;
    out->film_width = (float)width/(float)height*out->film_height;
  }
  if(out->iso < 1 || out->iso > 409600) out->iso = 100;
  return 0;
}

/* ---------------------------------------------------------------------------------------------- scene */
struct scene_b200_t
{
  prims_t prims;
  accel_t *accel;
  struct render_t *render;
  cb_render_desc_t desc;
  cb_material_t *materials;
  nra2_t *nra2;
  rgb2spec_b200_t *rgb2spec;
  table_file_t tables;
  char basename[1024], searchpath[1024];
  int sky;
  float sky_coeff[3], sky_scale;
  cb_envmap_t envmap;          /* `sky_envmap`: texels point into env_map (the mapped .fb file) */
  void *env_map; size_t env_size;
};

/* sky_envmap.c:272-311: "<file.fb> [brightness] [rot_x rot_y rot_z]" (degrees); the .fb layout is include/framebuffer.h:26-35 */
static void rotation(const float *axis, float angle, float *res)   /* mat3_rotate, include/matrix3.inc:111-134 */
{
  angle = angle/((float)180)*M_PI;
  const float s = sinf(angle), c = cosf(angle);
  res[0] = axis[0]*axis[0] + (1 - axis[0]*axis[0])*c; res[1] = axis[0]*axis[1]*(1 - c) - axis[2]*s; res[2] = axis[0]*axis[2]*(1 - c) + axis[1]*s;
  res[3] = axis[1]*axis[0]*(1 - c) + axis[2]*s; res[4] = axis[1]*axis[1] + (1 - axis[1]*axis[1])*c; res[5] = axis[1]*axis[2]*(1 - c) - axis[0]*s;
  res[6] = axis[2]*axis[0]*(1 - c) - axis[1]*s; res[7] = axis[2]*axis[1]*(1 - c) + axis[0]*s; res[8] = axis[2]*axis[2] + (1 - axis[2]*axis[2])*c;
}
static void mat3_product(const float *a, const float *b, float *res)   /* mat3_mul, matrix3.inc:29-39 */
{
  for(int k=0;k<9;k++) res[k] = 0.0f;
  for(int j=0;j<3;j++) for(int i=0;i<3;i++) for(int k=0;k<3;k++) res[i+3*j] += a[3*j+k]*b[3*k+i];
}
static int mat3_inverse(const float *a, float *inv)                    /* mat3_invert, matrix3.inc:67-95 */
{
#define A(y, x) a[(y - 1)*3 + (x - 1)]
  const float det = A(1, 1)*(A(3, 3)*A(2, 2) - A(3, 2)*A(2, 3)) - A(2, 1)*(A(3, 3)*A(1, 2) - A(3, 2)*A(1, 3)) + A(3, 1)*(A(2, 3)*A(1, 2) - A(2, 2)*A(1, 3));
  if(!(det != 0.0f)) return 1;
  const float invdet = 1.0f/det;
  inv[0] =  invdet*(A(3, 3)*A(2, 2) - A(3, 2)*A(2, 3)); inv[1] = -invdet*(A(3, 3)*A(1, 2) - A(3, 2)*A(1, 3)); inv[2] =  invdet*(A(2, 3)*A(1, 2) - A(2, 2)*A(1, 3));
  inv[3] = -invdet*(A(3, 3)*A(2, 1) - A(3, 1)*A(2, 3)); inv[4] =  invdet*(A(3, 3)*A(1, 1) - A(3, 1)*A(1, 3)); inv[5] = -invdet*(A(2, 3)*A(1, 1) - A(2, 1)*A(1, 3));
  inv[6] =  invdet*(A(3, 2)*A(2, 1) - A(3, 1)*A(2, 2)); inv[7] = -invdet*(A(3, 2)*A(1, 1) - A(3, 1)*A(1, 2)); inv[8] =  invdet*(A(2, 2)*A(1, 1) - A(2, 1)*A(1, 2));
#undef A
  return 0;
}
typedef struct { uint64_t magic, width, height; uint16_t channels, flags; float gain; } fb_header_b200_t;
static int envmap_open(struct scene_b200_t *s, const char *args)
{
  char name[1024] = "", path[2200];
  float b = 1.0f, rot[3] = {0.0f, 0.0f, 0.0f};
  if(sscanf(args, "%1023s %f %f %f %f", name, &b, rot, rot+1, rot+2) < 1) return 1;
  snprintf(path, sizeof(path), "%s", name);
  FILE *f = fopen(path, "rb");
  if(!f) { snprintf(path, sizeof(path), "%s/%s", s->searchpath, name); f = fopen(path, "rb"); }   /* fb_map: as given, then beside the scene */
  if(!f) return 1;
  fseek(f, 0, SEEK_END);
  const long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  fb_header_b200_t h;
  if(size < (long)sizeof(h) || fread(&h, sizeof(h), 1, f) != 1 || h.magic != 1936686951lu || h.channels != 4 ||
     h.width < 2 || h.width > (1u << 16) || h.height < 1 || h.height > (1u << 15) ||     /* bounded before multiplying: no uint64 wrap */
     (long)(h.width*h.height*4*sizeof(float) + sizeof(h)) != size || h.width != 2*h.height) { fclose(f); return 1; }
  s->env_size = (size_t)size - sizeof(h);
  s->env_map = malloc(s->env_size);
  if(!s->env_map || fread(s->env_map, 1, s->env_size, f) != s->env_size) { fclose(f); return 1; }
  fclose(f);
  cb_envmap_t *e = &s->envmap;
  e->width = (uint32_t)h.width; e->height = (uint32_t)h.height; e->pixels = (const float *)s->env_map; e->mul = b;
  const float ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1};
  float rx[9], ry[9], rz[9], tmp[9];
  rotation(ax, rot[0], rx); rotation(ay, rot[1], ry); rotation(az, rot[2], rz);
  mat3_product(ry, rz, tmp);
  mat3_product(rx, tmp, e->world);
  return mat3_inverse(e->world, e->world_inv);
}

static void chomp_comment(char *s)
{
  char *c = strchr(s, '#');
  if(c) *c = 0;
  size_t n = strlen(s);
  while(n && (s[n-1] == '\n' || s[n-1] == '\r' || s[n-1] == ' ' || s[n-1] == '\t')) s[--n] = 0;
}

void scene_b200_free(struct scene_b200_t *s)
{
  if(!s) return;
  if(s->render) render_cleanup(s->render);
  if(s->accel) accel_cleanup(s->accel);
  if(s->prims.shape) prims_cleanup(&s->prims);
  free(s->materials);
  free(s->env_map);
  free(s->nra2);
  rgb2spec_b200_free(s->rgb2spec);
  tables_free(&s->tables);
  free(s);
}

/* parse the shader list of `nra2` into materials (no GPU needed).  coeff_file: data/ergb2spec.coeff; table_file may be NULL
 * when no shader needs measured data */
static struct scene_b200_t *scene_open(const char *nra2_file, const char *coeff_file, const char *table_file, int load_shapes);
struct scene_b200_t *scene_b200_open(const char *nra2_file, const char *coeff_file, const char *table_file)
{
  return scene_open(nra2_file, coeff_file, table_file, 1);
}
/* sky line + shader list only: what the in-tree render module needs beside the geometry the reference has loaded itself */
struct scene_b200_t *scene_b200_open_shaders(const char *nra2_file, const char *coeff_file, const char *table_file)
{
  return scene_open(nra2_file, coeff_file, table_file, 0);
}

static struct scene_b200_t *scene_open(const char *nra2_file, const char *coeff_file, const char *table_file, int load_shapes)
{
  FILE *f = fopen(nra2_file, "rb");
  if(!f) { fprintf(stderr, "[scene b200] can't open %s for reading!\n", nra2_file); return 0; }
  struct scene_b200_t *s = calloc(1, sizeof(*s));
  if(s) s->nra2 = calloc(1, sizeof(nra2_t));
  if(!s || !s->nra2) { fprintf(stderr, "[scene b200] out of memory\n"); fclose(f); free(s); return 0; }
  snprintf(s->searchpath, sizeof(s->searchpath), "%s", nra2_file);
  char *c = s->searchpath + strlen(s->searchpath);
  for(;*c != '/' && c != s->searchpath;c--);
  *c = 0;
  snprintf(s->basename, sizeof(s->basename), "%s", nra2_file);
  c = s->basename + strlen(s->basename);
  for(;*c != '.' && c != s->basename;c--);
  if(c != s->basename) *c = 0;

  s->rgb2spec = rgb2spec_b200_load(coeff_file);
  if(!s->rgb2spec) { fprintf(stderr, "[scene b200] could not load `%s', expect trouble!\n", coeff_file); fclose(f); scene_b200_free(s); return 0; }
  if(table_file && tables_load(&s->tables, table_file)) { fprintf(stderr, "[scene b200] could not load table file `%s'\n", table_file); fclose(f); scene_b200_free(s); return 0; }
  s->nra2->rgb2spec = s->rgb2spec;
  s->nra2->tables = table_file ? &s->tables : 0;

  char line[2048];
  if(!fgets(line, sizeof(line), f)) { fclose(f); scene_b200_free(s); return 0; }
  chomp_comment(line);
  char sky[64] = "";
  sscanf(line, " %63s", sky);
  if(!strncmp(sky, "black", 5)) s->sky = CB_SKY_BLACK;          /* src/shader.c:633-641: prefix matches, like there */
  else if(!strncmp(sky, "cloudy_sky", 10) || !strncmp(sky, "cloudy", 6) || !strncmp(sky, "clear_sky", 9)) s->sky = CB_SKY_CLOUDY;
  else if(!strcmp(sky, "sky_const"))
  { /* src/shaders/sky_const.c:89-101: "r g b [scale]" */
    float col[3] = {1.0f, 1.0f, 1.0f}, scale = 1.0f;
    sscanf(strstr(line, "sky_const") + 9, "%f %f %f %f", col, col+1, col+2, &scale);
    s->sky = CB_SKY_CONST;
    s->sky_scale = scale*rgb_to_coeff(s->rgb2spec, col, s->sky_coeff);
  }
  else if(!strcmp(sky, "sky_envmap"))
  {
    if(envmap_open(s, strstr(line, "sky_envmap") + 10))
    { fprintf(stderr, "[envmap] could not read hdri file (w = 2h, four channels of rgb2spec coefficients + scale)!\n"); fclose(f); scene_b200_free(s); return 0; }
    s->sky = CB_SKY_ENVMAP;
  }
  else if(!strncmp(sky, "daylight", 8))
  { /* the reference's analytic daylight model (src/shaders/daylight.h) is SURVEY 8f rank 3 */
    fprintf(stderr, "[scene b200] sky `%s' is not supported by the gpu path (only `black', `cloudy' and `sky_const'); no cpu fallback\n", sky);
    fclose(f); scene_b200_free(s); return 0;
  }
  else
  { /* any other name: upstream tries dlopen("lib<name>.so"), which fails for everything that is not one of its modules, and
     * keeps the default sky -- the cloudy one (src/shader.c:612-614,643-675; regression/0090_vstack says `const 1 1 1 2000') */
    printf("[shader_init] failed to load sky shader `lib%s.so'\n", sky);
    s->sky = CB_SKY_CLOUDY;
  }
  if(!fgets(line, sizeof(line), f) || sscanf(line, "%d", &s->nra2->num_shaders) != 1 || s->nra2->num_shaders < 0 || s->nra2->num_shaders > MAX_SHADERS)
  { fprintf(stderr, "[scene b200] corrupt model file: could not read number of shaders!\n"); fclose(f); scene_b200_free(s); return 0; }
  for(int k=0;k<s->nra2->num_shaders;k++)
  {
    if(!fgets(line, sizeof(line), f)) { fclose(f); scene_b200_free(s); return 0; }
    chomp_comment(line);
    int adv = 0;
    s->nra2->line[k].name[0] = 0;
    sscanf(line, " %63s%n", s->nra2->line[k].name, &adv);
    snprintf(s->nra2->line[k].args, sizeof(s->nra2->line[k].args), "%s", line + adv);
  }
  s->materials = calloc(s->nra2->num_shaders ? s->nra2->num_shaders : 1, sizeof(cb_material_t));
  for(int k=0;k<s->nra2->num_shaders;k++)
  {
    cb_material_t *m = s->materials + k;
    int have_bsdf = 0;
    m->table = -1;
    const shader_line_t *l = s->nra2->line + k;
    if(!strcmp(l->name, "exterior"))
    { /* src/shader.c:699-716: "<medium shader> [volume light]"; the line itself is no material */
      int id = -1, light = 0;
      sscanf(l->args, " %d %d", &id, &light);
      if(light) { fprintf(stderr, "[scene b200] volume lights are not supported by the gpu path; no cpu fallback\n"); fclose(f); scene_b200_free(s); return 0; }
      if(id >= 0 && !(s->nra2->exterior_medium = flatten_medium(s->nra2, id)))
      { fprintf(stderr, "[scene b200] exterior medium %d is not a homogeneous medium_rgb chain; no cpu fallback\n", id); fclose(f); scene_b200_free(s); return 0; }
      m->num_ops = -1; m->bsdf = -1;
      continue;
    }
    if(!strcmp(l->name, "interior"))
    { /* src/shaders/interior.c:53-72: "<surface id> <interior id>", negative = relative */
      int surf = 0, inner = 0, medium = 0;
      if(sscanf(l->args, " %d %d", &surf, &inner) == 2)
      {
        if(surf < 0) surf += k;
        if(inner < 0) inner += k;
        medium = flatten_medium(s->nra2, inner);
      }
      s->nra2->flatten_calls = 0;
      if(!medium || flatten(s->nra2, surf, m, &have_bsdf, 0)) { memset(m, 0, sizeof(*m)); m->num_ops = -1; m->bsdf = -1; continue; }
      if(!have_bsdf) m->bsdf = CB_BSDF_DIFFUSE;
      m->medium = medium;
      continue;
    }
    s->nra2->flatten_calls = 0;
    if(flatten(s->nra2, k, m, &have_bsdf, 0)) { memset(m, 0, sizeof(*m)); m->num_ops = -1; m->bsdf = -1; continue; }
    if(!have_bsdf) m->bsdf = CB_BSDF_DIFFUSE;   /* a prepare-only shader on a shape gets the default diffuse callbacks (shader.c:761-787) */
  }

  if(!load_shapes) { fclose(f); return s; }
  /* shapes: common_load_scene, src/corona_common.c:30-68 */
  int num_shapes = 0;
  if(!fgets(line, sizeof(line), f) || sscanf(line, "%d", &num_shapes) != 1 || num_shapes < 0 || num_shapes > (1 << 24))
  { fprintf(stderr, "[common_load_scene] corrupt model file: could not read number of shapes!\n"); fclose(f); scene_b200_free(s); return 0; }
  prims_init(&s->prims);
  prims_allocate(&s->prims, num_shapes);
  if(!s->prims.shape) { fprintf(stderr, "[common_load_scene] out of memory for %d shapes\n", num_shapes); fclose(f); scene_b200_free(s); return 0; }
  for(int shape=0;shape<num_shapes;shape++)
  {
    if(!fgets(line, sizeof(line), f)) break;
    int shader = 0; char name[512], tex[512] = "none", path[2048];
    if(sscanf(line, "%d %511s %511s", &shader, name, tex) < 2)
    { fprintf(stderr, "[common_load_scene] WARN: malformed line (%d): %s\n", shape + s->nra2->num_shaders + 1, line); continue; }
    if(shader >= s->nra2->num_shaders) fprintf(stderr, "[common_load_scene] WARN: shader %d in line %d (%s) out of bounds!\n", shader, shape, name);
    if(shader < 0 || shader >= s->nra2->num_shaders) shader = 0;
    snprintf(path, sizeof(path), "%s", name);
    FILE *t = 0;
    { char g[2100]; snprintf(g, sizeof(g), "%s.geo", path); t = fopen(g, "rb"); }
    if(t) fclose(t); else snprintf(path, sizeof(path), "%s/%s", s->searchpath, name);   /* prims.c:775-781 */
    prims_load(&s->prims, path, tex, shader);
  }
  fclose(f);
  prims_allocate_index(&s->prims);
  return s;
}

const cb_material_t *scene_b200_materials(const struct scene_b200_t *s, int *num) { if(num) *num = s->nra2->num_shaders; return s->materials; }
const cb_medium_t *scene_b200_media(const struct scene_b200_t *s, int *num, int *exterior)
{ if(num) *num = s->nra2->num_media; if(exterior) *exterior = s->nra2->exterior_medium; return s->nra2->media; }
const char *scene_b200_basename(const struct scene_b200_t *s) { return s->basename; }
uint64_t scene_b200_num_prims(const struct scene_b200_t *s) { return s->prims.num_prims; }
const float *scene_b200_aabb(const struct scene_b200_t *s) { static const float empty[6] = {1, 1, 1, -1, -1, -1}; return s->accel ? accel_aabb(s->accel) : empty; }

/* the shader-list half of a render description: materials, measured tables, media, sky (borrowed pointers into s) */
void scene_b200_fill_desc(const struct scene_b200_t *s, cb_render_desc_t *d)
{
  d->materials = s->materials; d->num_materials = s->nra2->num_shaders;
  d->tables = s->nra2->used; d->num_tables = s->nra2->num_used;
  d->sky = s->sky;
  for(int k=0;k<3;k++) d->sky_coeff[k] = s->sky_coeff[k];
  d->sky_scale = s->sky_scale;
  d->envmap = s->sky == CB_SKY_ENVMAP ? &s->envmap : 0;
  d->media = s->nra2->media; d->num_media = s->nra2->num_media; d->exterior_medium = s->nra2->exterior_medium;
}

/* accel_init + accel_build + camera + render_b200_init: everything main.c's init() does for the hot path (src/main.c:250-359) */
int scene_b200_prepare(struct scene_b200_t *s, uint32_t width, uint32_t height, int sampler, int pointsampler, int colour, uint64_t frame,
                       const char *cam_file)
{
  while(width & 0x1f) width++;      /* src/view.c:295-296 */
  while(height & 0x1f) height++;
  const int timing = getenv("CB200_TIMING") != 0;
  struct timespec ts0, ts1, ts2;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  s->accel = accel_init(&s->prims);
  if(!s->accel) return 1;
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  accel_build(s->accel, s->basename);
  if(!accel_b200_handle(s->accel)) return 1;
  clock_gettime(CLOCK_MONOTONIC, &ts2);
  if(timing) fprintf(stderr, "[scene b200] accel_init (device / context) %.3f s, accel_build (upload + gpu build + primid read-back) %.3f s\n",
                     (ts1.tv_sec - ts0.tv_sec) + 1e-9*(ts1.tv_nsec - ts0.tv_nsec), (ts2.tv_sec - ts1.tv_sec) + 1e-9*(ts2.tv_nsec - ts1.tv_nsec));
  char cam[1100];
  if(cam_file) snprintf(cam, sizeof(cam), "%s", cam_file);
  else snprintf(cam, sizeof(cam), "%s01.cam", s->basename);        /* src/view.c:301-312 */
  cb_render_desc_t *d = &s->desc;
  memset(d, 0, sizeof(*d));
  if(scene_b200_read_camera(cam, width, height, &d->camera)) { fprintf(stderr, "[scene b200] could not read camera `%s'\n", cam); return 1; }
  d->width = width; d->height = height;
  scene_b200_fill_desc(s, d);
  d->sampler = sampler; d->pointsampler = pointsampler; d->colour_camera = colour;
  d->max_path_len = 32; d->frame = frame; d->rank = 0; d->world = 1; d->batch_paths = 0;
  s->render = render_b200_init(s->accel, d);
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  if(timing) fprintf(stderr, "[scene b200] camera + render_b200_init (materials, lights, halton tables, path pool) %.3f s\n",
                     (ts0.tv_sec - ts2.tv_sec) + 1e-9*(ts0.tv_nsec - ts2.tv_nsec));
  return s->render ? 0 : 1;
}

struct render_t *scene_b200_render(struct scene_b200_t *s) { return s->render; }
const cb_render_desc_t *scene_b200_desc(const struct scene_b200_t *s) { return &s->desc; }

/* fb_export, include/framebuffer.h:142-175: PFM with the header padded to 16 bytes, gain applied */
int scene_b200_write_pfm(const char *filename, const float *fb, uint32_t width, uint32_t height, float gain)
{
  FILE *f = fopen(filename, "wb");
  if(!f) return 1;
  char header[1024];
  snprintf(header, sizeof(header), "PF\n%lu %lu\n-1.0", (unsigned long)width, (unsigned long)height);
  const size_t len = strlen(header);
  fputs(header, f);
  long off = 0;
  while((len + 1 + off) & 0xf) off++;
  while(off-- > 0) fputc('0', f);
  fputc('\n', f);
  for(uint64_t k=0;k<(uint64_t)width*height*3;k++)
  {
    const float v = fb[k]*gain;
    fwrite(&v, sizeof(float), 1, f);
  }
  fclose(f);
  return 0;
}
