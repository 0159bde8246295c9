/* corona_b200_render.h -- C ABI of the wavefront pt / ptdl integrator in libcorona_b200.so.
 *
 * Replaces the reference's per-thread recursive loop
 *     view_render -> work_sample -> render_sample_path -> sampler_create_path      (src/view.c:618-645,
 *     src/render.d/gi.c:81, src/sampler.d/pt.c:40 / ptdl.c:112)
 * by kernels over batches of path indices: camera-ray generation (src/camera.d/thinlens.c:68-128),
 * closest-hit traversal, vertex preparation + material chain (src/shader.c:462-542), emission with MIS,
 * next-event estimation (include/pathspace/nee.h:87-243, src/lights.d/list.c) with shadow rays in
 * path_visible semantics (src/pathspace.c:311-344), BSDF sampling, splatting (src/view.c:455-495,
 * include/filter/blackmanharris.h:43-77).
 *
 * The scene description arrives flattened: the host layer (or a test) walks the reference's shader list
 * (`.nra2`: mult / color / colorcheckersg / dielectric / diffdiel / metal / diffuse / interior / medium_rgb / exterior) once at
 * load time and hands over one cb_material_t per shader index a shape may reference.  Unknown shader kinds are a hard error upstream --
 * there is no CPU fallback.
 */
#ifndef CORONA_B200_RENDER_H
#define CORONA_B200_RENDER_H

#include "corona_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- camera: the fields of camera_t the thin lens model reads (include/camera.h:13-35) ------------------- */
typedef struct cb_camera_t
{
  float pos[3], pos_t1[3];
  float orient[4], orient_t1[4];   /* quaternion w,x,y,z (include/quaternion.h) */
  float focus;
  float film_width, film_height;
  int32_t aperture_value;          /* index into the f-stop table (src/view.c:71-73) */
  int32_t exposure_value;          /* index into the exposure-time table (src/view.c:75-79) */
  float focal_length;
  float iso;
}
cb_camera_t;

/* ---- materials ------------------------------------------------------------------------------------------
 * A material = the reference's `mult <n> <pre...> <host>` flattened: up to CB_MAX_MATOPS prepare() steps that
 * fill vertex_shading_t slots, then the host BSDF.                                                          */
#define CB_MAX_MATOPS 6
enum { CB_OP_NONE = 0,
       CB_OP_COLOR = 1,          /* src/shaders/color.c:75-81: slot <- mul * rgb2spec(coeff, lambda), roughness */
       CB_OP_CHECKERSG = 2 };    /* src/shaders/colorcheckersg.c:244-261: 14x10 ColorChecker SG by uv */
enum { CB_SLOT_DIFFUSE = 0, CB_SLOT_SPECULAR = 1, CB_SLOT_EMISSION = 2, CB_SLOT_VOLUME = 3, CB_SLOT_GLOSSY = 4,
       CB_SLOT_ROUGHNESS = 5, CB_SLOT_TRANSMIT_TO_EYE = 6 };   /* src/shaders/texture.h:8-21 */
enum { CB_BSDF_DIFFUSE = 0,      /* src/shader.c:157-257 */
       CB_BSDF_DIELECTRIC = 1,   /* src/shaders/dielectric.c */
       CB_BSDF_METAL = 2,        /* src/shaders/metal.c */
       CB_BSDF_DIFFDIEL = 3 };   /* src/shaders/diffdiel.c: dielectric reflection, diffuse transmission */

typedef struct cb_matop_t
{
  int32_t op, slot;
  float coeff[3];                /* rgb2spec sigmoid-polynomial coefficients (include/rgb2spec.h:87-128) */
  float mul;
  float roughness;
  int32_t table;                 /* CB_OP_CHECKERSG: index into cb_render_desc_t.tables (140 rows x 36 wavelengths) */
}
cb_matop_t;

/* wavelength-indexed lookup table, nearest sample like the reference's measured data:
 * value(row, lambda) = data[row*num_lambda + (int)((lambda - lambda_min)/lambda_step)]
 * (ColorChecker SG reflectances, src/shaders/colorcheckersg.c:168-179; complex IORs of metals as two rows n, k,
 * src/shaders/fresnel.h:519-531).  The data stay with the caller's scene description; the library copies them. */
typedef struct cb_table_t
{
  float lambda_min, lambda_step;
  int32_t num_lambda, rows;
  const float *data;
}
cb_table_t;

typedef struct cb_material_t
{
  int32_t num_ops;
  int32_t bsdf;
  float param[4];                /* dielectric, diffdiel: n_d, abbe */
  int32_t table;                 /* metal: index of its (n,k) table */
  int32_t medium;                /* `interior <surface> <medium>` (src/shaders/interior.c): 1 + index into cb_render_desc_t.media of
                                    the homogeneous medium behind this surface, 0 = vacuum */
  cb_matop_t ops[CB_MAX_MATOPS];
}
cb_material_t;

/* homogeneous participating medium = the reference's `mult 1 <color v ...> <medium_rgb ...>` chain flattened
 * (src/shaders/medium_rgb.c:46-60,112-139, src/shaders/texture.h:46-52): mu_t(lambda) = mu_t_mul * rgb2spec(mu_t_coeff),
 * single-scattering albedo from the `color v` step (mu_s = albedo * mu_t; without one the medium only absorbs),
 * Henyey-Greenstein phase function with mean cosine g. */
#define CB_MAX_MEDIA 63
typedef struct cb_medium_t
{
  float mu_t_coeff[3];
  float mu_t_mul;
  float g;
  int32_t has_albedo;
  float albedo_coeff[3];
  float albedo_mul;
}
cb_medium_t;

/* ---- render description ---------------------------------------------------------------------------------- */
enum { CB_SAMPLER_PT = 0, CB_SAMPLER_PTDL = 1,            /* src/sampler.d/pt.c, ptdl.c */
       CB_SAMPLER_PTNEE = 2 };                            /* src/sampler.d/ptnee.c: next-event estimation only, no mis */
enum { CB_POINTS_RAND = 0, CB_POINTS_HALTON = 1 };        /* src/pointsampler.d/rand.c, halton.c */
enum { CB_COLOUR_XYZ = 0, CB_COLOUR_REC709 = 1 };         /* COL_camera (Makefile:122-136) */
enum { CB_SKY_BLACK = 0, CB_SKY_CLOUDY = 1,               /* line 1 of the .nra2: built-in skies of src/shader.c:262-334,633-660 */
       CB_SKY_CONST = 2,                                  /* `sky_const r g b [scale]`: src/shaders/sky_const.c */
       CB_SKY_ENVMAP = 3 };                               /* `sky_envmap file.fb brightness [rot_x rot_y rot_z]`: src/shaders/sky_envmap.c */

/* latitude-longitude environment map as sky_envmap.c maps it from its `.fb` file (include/framebuffer.h:26-35: four floats per
 * texel = rgb2spec coefficients + scale, width == 2*height), brightness and the rotation built from the three angles
 * (sky_envmap.c:286-300).  The library builds the importance-sampling mip hierarchy of sky_envmap.c:329-361 itself. */
typedef struct cb_envmap_t
{
  uint32_t width, height;
  const float *pixels;
  float mul;
  float world[9], world_inv[9];  /* row major, dir_world = world * dir_map */
}
cb_envmap_t;

typedef struct cb_render_desc_t
{
  uint32_t width, height;        /* already padded to multiples of 32 by the caller (src/view.c:295-296) */
  cb_camera_t camera;
  const cb_material_t *materials;
  int32_t num_materials;
  const cb_table_t *tables;
  int32_t num_tables;
  int32_t sampler;               /* CB_SAMPLER_* */
  int32_t pointsampler;          /* CB_POINTS_* */
  int32_t colour_camera;         /* CB_COLOUR_* */
  int32_t max_path_len;          /* ptdl: rt.sampler->max_path_len, <= 32 (PATHSPACE_MAX_VERTS) */
  uint64_t frame;                /* rt.anim_frame: seeds the Halton permutations / the counter RNG */
  uint32_t rank, world;          /* sample-space split (informational): every random draw is a function of (frame, path index,
                                    dimension) only, so ranks that render disjoint index ranges sum to the single-GPU image */
  uint64_t batch_paths;          /* paths in flight per wave (0 = default) */
  int32_t sky;                   /* CB_SKY_* */
  float sky_coeff[3];            /* CB_SKY_CONST: rgb2spec coefficients of the colour and scale * mul (sky_const.c:89-101) */
  float sky_scale;
  int32_t exterior_medium;       /* `exterior <id>` (src/shader.c:544-566,699-716): 1 + index into media of the medium the camera
                                    sits in, 0 = vacuum */
  const cb_medium_t *media;
  int32_t num_media;
  int32_t pad;
  const cb_envmap_t *envmap;     /* CB_SKY_ENVMAP */
}
cb_render_desc_t;

typedef struct cb200_render cb200_render_t;

typedef struct cb_render_stats_t
{
  uint64_t paths;                /* render_sample_path calls */
  uint64_t rays_closest;         /* accel_intersect calls from path_propagate */
  uint64_t rays_shadow;          /* accel_intersect calls from path_visible */
  uint64_t splats;
  uint64_t kernel_launches;
  /* filled while cb200_render_instrument(timing) is on: milliseconds between CUDA events recorded on the pass' own stream
   * around the launches of each kernel class: {path start, closest-hit traversal, shade, shadow traversal, nee resolve} */
  double ms[5];
  /* filled while cb200_render_instrument(counters) is on: the reference's ACCEL_DEBUG counters (qbvhmp.c:83-90) of the
   * closest-hit and of the shadow waves: {accel_intersect, aabb_intersect, aabb_true, prims_intersect} */
  uint64_t trav_closest[4], trav_shadow[4];
}
cb_render_stats_t;

cb200_render_t *cb200_render_create(cb200_accel_t *a, const cb_render_desc_t *desc);
void cb200_render_destroy(cb200_render_t *r);
/* one call = the work of view_render()'s fan-out for path indices [first, first+count) (src/view.c:636-645);
 * accumulates into the device framebuffer (W*H*3 floats), asynchronous on `stream` */
int  cb200_render_pass(cb200_render_t *r, uint64_t first_index, uint64_t count, void *stream);
/* streaming form for back-to-back progressions: returns as soon as every index of the range has been STARTED; paths that are
 * still bouncing stay in the pool and ride along with the next call's waves (so no launch ever runs on a nearly empty wave),
 * their contributions arrive in the framebuffer later.  cb200_render_flush traces the stragglers to the end; it is implied by
 * cb200_render_download and must precede any use of cb200_render_fb_device as a finished image.
 * cb200_render_pass == cb200_render_pass_stream + cb200_render_flush.                                                        */
int  cb200_render_pass_stream(cb200_render_t *r, uint64_t first_index, uint64_t count, void *stream);
int  cb200_render_flush(cb200_render_t *r, void *stream);
int  cb200_render_clear(cb200_render_t *r, void *stream);
/* per-kernel-class CUDA event timing and/or traversal work counters for the following passes (bench.py's roofline) */
int  cb200_render_instrument(cb200_render_t *r, int timing, int counters);
/* device pointer of the accumulation buffer (for ncclReduce across ranks) and its download.  The image the
 * reference writes is fb * gain, gain = iso / (100 * spp) (src/view.c:656) */
void *cb200_render_fb_device(cb200_render_t *r);
/* accumulate into a caller-owned device buffer (W*H*3 floats) from now on, NULL = back to the library's own.  Lets the
 * caller double-buffer: reduce / read one buffer while the next progression renders into the other. */
int  cb200_render_set_framebuffer(cb200_render_t *r, void *d_fb);
int  cb200_render_download(cb200_render_t *r, float *fb_host, void *stream);   /* flushes first: the finished image */
/* How splats reach the framebuffer.  CB200_ACCUM_ATOMIC: like the reference (view_splat -> filter_blackmanharris_splat ->
 * common_atomic_add, a CAS loop per tap; here fp32 atomic adds).  CB200_ACCUM_TILES: atomic-free -- samples are recorded, sorted by
 * 32x32 pixel tile, filtered into a shared-memory patch by one block per tile and added to the framebuffer by that block alone
 * (four checkerboard phases, so that the patches' two-pixel borders never overlap within a launch).  Same weights, same per-pixel
 * terms; only the fp32 summation order differs.  With a `--dbor` cascade the atomic path is used regardless.  Call between passes. */
#define CB200_ACCUM_ATOMIC 0
#define CB200_ACCUM_TILES  1
int  cb200_render_set_accumulation(cb200_render_t *r, int mode);
int  cb200_render_accumulation(cb200_render_t *r);
/* `--dbor n` (src/view.c:291,339-350): the density based outlier rejection cascade of view_splat_col (view.c:497-522).  With
 * n > 1 every splat also goes, through the same Blackman-Harris taps, into the two of n extra buffers (W*H*3 floats each)
 * whose brightness range brackets the sample; n <= 1 switches the cascade off.  Call between progressions (it synchronises the
 * device); the buffers start at zero and are cleared by cb200_render_clear.  Level buffers are scaled by the same gain as the
 * framebuffer (view.c:659-664) and written as `<basename><suffix>_dbor%02d.pfm` by the reference (view.c:553-556). */
int  cb200_render_set_dbor(cb200_render_t *r, int32_t levels);
int  cb200_render_num_dbors(cb200_render_t *r);
void *cb200_render_dbor_device(cb200_render_t *r, int32_t level);              /* for the reduce across ranks */
int  cb200_render_download_dbor(cb200_render_t *r, int32_t level, float *fb_host, void *stream);   /* flushes first */
/* the accumulation buffer as it stands, WITHOUT flushing: what a progressive display shows between streamed progressions
 * (the stragglers' contributions arrive with a later snapshot; the reference's display reads its framebuffer mid-flight too) */
int  cb200_render_snapshot(cb200_render_t *r, float *fb_host, void *stream);
/* the same without waiting: the buffer is copied on the device behind the work queued on `stream` and drained to `fb_host`
 * (pinned memory, or the call degrades to a synchronous copy) on a stream of the library's own, so the transfer overlaps the
 * next progression's kernels.  `fb_host` must stay valid until cb200_render_snapshot_wait / _snapshot / _download / _destroy
 * returns; those wait for the transfer. */
int  cb200_render_snapshot_async(cb200_render_t *r, float *fb_host, void *stream);
int  cb200_render_snapshot_wait(cb200_render_t *r);
int  cb200_render_stats(cb200_render_t *r, cb_render_stats_t *out);
/* the view's path statistics (src/view.c:46-47,470-471: stat_enery / stat_cnt): summed contribution and number of splats per path
 * length (vertices, 2..32), what view_print_info draws as the histogram of the sidecar file (view.c:759-790).  Off by default;
 * enable before rendering (synchronises the device), cleared by cb200_render_clear. */
int  cb200_render_path_stats(cb200_render_t *r, int enable);
int  cb200_render_get_path_stats(cb200_render_t *r, double energy[33], uint64_t count[33]);

/* component entry points for parity tests (device work, host buffers): */
/* Halton / counter RNG value for (path index, dimension) as pointsampler() returns it (halton.c:69-84) */
int  cb200_render_point(cb200_render_t *r, const uint64_t *index, const int32_t *dim, float *out, uint64_t n);
/* primary rays for path indices: camera_sample (thinlens.c:115-128) with the path's own random dims;
 * out_aux[n][4] = {pixel_i, pixel_j, lambda, throughput} */
int  cb200_render_camera_rays(cb200_render_t *r, uint64_t first_index, uint64_t n, cb_ray_t *out_rays, float *out_aux);
/* next-event samples at the first hit vertex of path indices [first_index, first_index + n) -- nee_sample (include/pathspace/
 * nee.h:87-243: lights_pdf_type, sample_cdf over the light list, prims_sample, shader_brdf, path_G, path_visible) weighted as
 * ptdl.c:139-147 does -- handed back instead of splatted; only on a render object without paths in flight, framebuffer untouched.
 * out[k][16] = {pixel_i, pixel_j, lambda, throughput x mis weight, |light point - ray origin|, light prim (2 words, bit
 * pattern), shadow ray pos[3], dir[3], search limit, visible (1/0), path length at the splat}; *n_out = records written (<= n) */
int  cb200_render_nee_records(cb200_render_t *r, uint64_t first_index, uint64_t n, float *out, uint64_t *n_out);
/* the paths that go on behind their first hit vertex, as path_extend leaves them (src/pathspace.c:186-260: shader_sample of the
 * BSDF with the vertex' own random dimensions, prims_offset_ray): out[k][16] = {pixel_i, pixel_j, lambda, e[2].omega[3], origin of
 * the next ray[3], throughput into the next vertex, throughput into this one, bsdf pdf (projected solid angle), |n . omega|, vertex
 * position[3]}.  tangent_frame_scrambling > 0 presets path->tangent_frame_scrambling (upstream draws it from the worker thread's
 * twister, pathspace.c:212-213) so that a reference harness can run with the same number.  Same preconditions as above. */
int  cb200_render_bounce_records(cb200_render_t *r, uint64_t first_index, uint64_t n, float tangent_frame_scrambling, float *out, uint64_t *n_out);
/* emission found by extension as the sampler splats it (pt.c:44-50, ptdl.c:124-131 with sampler_mis :78-88): wave 1 = emitters the
 * camera sees directly, wave 2 = emitters the first BSDF-sampled edge ends on, weighted against next-event estimation;
 * out[k][8] = {pixel_i, pixel_j, lambda, throughput x emission x mis weight, path length at the splat, 0, 0, 0} */
int  cb200_render_emission_records(cb200_render_t *r, uint64_t first_index, uint64_t n, float tangent_frame_scrambling, int32_t wave,
                                   float *out, uint64_t *n_out);

/* the origin of the ray that leaves surface point x[i] in direction dir[i], as the integrator computes it: prims_offset_ray
 * (src/prims.c:374-388; 3 floats per point in, 3 out) */
int  cb200_render_offset_ray(cb200_render_t *r, const float *x, const float *dir, float *out_pos, uint64_t n);

/* one BSDF in isolation, the protocol of the reference's tools/battle-test.c:57-236 (regression/0052_dielectric, 0053): a
 * surface vertex at the origin with n = gn = +z (flip: -z), tangent frame get_onb(n), vacuum outside, the shading slots preset
 * by the caller, then the host bsdf's prepare(), sample() with the three given random dimensions, and brdf()/pdf() for the
 * given outgoing direction.  Replaces the dlopen'ed sample/brdf/pdf callbacks (src/shader.h:34-111) for `material`.            */
typedef struct cb_bsdf_query_t
{
  float wi[3];                     /* e[v].omega: unit, pointing AT the surface */
  float wo[3];                     /* e[v+1].omega for brdf() / pdf() */
  float lambda;
  float rand[3];                   /* s_dim_omega_x, s_dim_omega_y, s_dim_scatter_mode (include/pathspace.h:39-45) */
  float rd, rs, rg, roughness;     /* vertex_shading_t */
  int32_t flip;
}
cb_bsdf_query_t;
typedef struct cb_bsdf_result_t
{
  float s_wo[3], s_weight, s_pdf;  /* sample(): direction, throughput weight f*cos/p, pdf in projected solid angle */
  uint32_t s_mode;                 /* vertex_scattermode_t bits after sample() */
  float f;                         /* brdf() */
  uint32_t f_mode;
  float pdf;                       /* pdf(p, e1 = v, v, e2 = v+1) */
}
cb_bsdf_result_t;
int  cb200_render_bsdf(cb200_render_t *r, int32_t material, const cb_bsdf_query_t *queries, cb_bsdf_result_t *results, uint64_t n);

/* The same for a homogeneous medium (index into cb_render_desc_t.media), evaluated the way the integrator's kernels do: the
 * medium's coefficients at lambda (medium_rgb.c:46-60 behind its `color v`), the sampled free-flight distance and its pdf for
 * rand[2] (shader_vol_sample, src/shader.c:76-104), transmittance and distance pdf of an edge of length `dist` that ends on a
 * volume vertex (shader.c:46-72,107-131), and the phase-function callbacks at that vertex (medium_rgb.c:62-103) with incoming
 * direction wi, tangent frame scrambling 0.5, random dimensions rand[0..1] and outgoing direction wo for brdf()/pdf(). */
typedef struct cb_medium_query_t
{
  float wi[3], wo[3];
  float lambda;
  float rand[3];                   /* s_dim_omega_x, s_dim_omega_y, s_dim_free_path */
  float dist;
}
cb_medium_query_t;
typedef struct cb_medium_result_t
{
  float mu_t, mu_s;
  float free_dist, free_pdf;
  float transmittance, vol_pdf;
  float s_wo[3], s_weight, s_pdf;
  uint32_t s_mode;
  float f;
  uint32_t f_mode;
  float pdf;
}
cb_medium_result_t;
int  cb200_render_medium(cb200_render_t *r, int32_t medium, const cb_medium_query_t *queries, cb_medium_result_t *results, uint64_t n);

/* ---- multi-GPU: samples per pixel split over the GPUs of one box ---------------------------------------------------
 * One rank (a process, or a thread with its own device) per GPU: rank g renders progressions g, g+N, ... -- the path-index ranges a
 * 1-GPU run uses for those progressions, so the union over ranks is the same set of paths -- and after every progression ONE reduce
 * (ncclReduce, sum, root 0, W*H*3 floats over NVLink) adds the rank's accumulation buffer into rank 0's running sum.  The reducer owns
 * the rank's two accumulation buffers (double buffered), the root's sum, a side stream and the NCCL communicator; reduce, accumulate,
 * clear and the optional copy of the running sum to the host are queued on the side stream and overlap the next progression.
 * Replaces nothing in the reference (it has no multi-device mode): it partitions view_render()'s index ranges (src/view.c:636-638).
 * NCCL is dlopen'ed at the first call.
 *     id     : rank 0 calls cb200_comm_unique_id and hands the CB200_COMM_ID_BYTES bytes to the other ranks (pipe, file, MPI ...)
 *     step   : the rank's local progression counter, 0, 1, 2, ...; every rank must run the same number of begin/end rounds (the reduce
 *              is a collective), a rank without work in a round still calls both
 *     stream : the stream the progression is rendered on (NULL = default stream)                                                    */
#define CB200_COMM_ID_BYTES 128
typedef struct cb200_reducer cb200_reducer_t;
int  cb200_comm_unique_id(void *id_bytes);
cb200_reducer_t *cb200_reducer_create(cb200_render_t *r, const void *id_bytes, int rank, int world, uint32_t width, uint32_t height);
void cb200_reducer_destroy(cb200_reducer_t *q);
int  cb200_reducer_begin(cb200_reducer_t *q, uint64_t step, void *stream);    /* before cb200_render_pass[_stream] / _flush of this round */
int  cb200_reducer_end(cb200_reducer_t *q, uint64_t step, float *fb_host, void *stream);   /* fb_host: rank 0's progressive host image (pinned) or NULL */
int  cb200_reducer_finish(cb200_reducer_t *q, float *fb_host);               /* waits; rank 0 receives the sum of all ranks */
int  cb200_reducer_clear(cb200_reducer_t *q);

#ifdef __cplusplus
}
#endif
#endif
