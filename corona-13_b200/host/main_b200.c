/* main_b200.c -- `corona_b200`: the reference's offline command line for the hot path, on the GPU modules.
 *
 *   corona_b200 <scene.nra2> [-s spp] [-w width] [-h height] [--frame n] [-x [name]] [-c camfile] [--batch n]
 *               [--sampler pt|ptdl|ptnee] [--points rand|halton] [--colour xyz|rec709] [--coeff file] [--tables file]
 *               [--dump-materials file] [--dbor n] [--gpus n] [--retain-framebuffer] [-q]
 *
 * Same arguments and defaults as the reference binary where they exist there (src/main.c:250-282,415-437,
 * src/view.c:262-297, src/display.d/null.c:47-56): -s samples per pixel then write the image and quit, -w/-h frame size
 * (padded to multiples of 32), --frame = rt.anim_frame (seeds the point sampler, default 1), -x output name (default
 * `render'), data/ergb2spec.coeff relative to the working directory.  --sampler / --points / --colour stand in for the
 * reference's compile-time MOD_sampler / MOD_pointsampler / COL_camera.  Writes <basename><name>_fb00.pfm like
 * view_write_images (src/view.c:549); --dbor n (src/view.c:291) adds the outlier rejection cascade of view_splat_col and its
 * <basename><name>_dbor%02d.pfm files (view.c:553-556).  Next to the image it writes the reference's sidecar <image>.pfm.txt
 * (common_write_sidecar, corona_common.c:70-97: module info, spp, s/prog, mean image intensity, path-length energy histogram), a
 * machine-readable <image>.pfm.json (rays/s, spp/s, rays per path) and, with --retain-framebuffer (view.c:279), the frame buffer file
 * <basename>_<name>_fb00.fb (framebuffer.h).  Without a CUDA device it refuses: there is no CPU path in this binary.
 *
 * --gpus n: the samples per pixel are split over n GPUs of the box, one process per GPU (forked before any device
 * work; this process is rank 0).  Rank g renders the progression groups g, g+n, ... -- the same path-index ranges a 1-GPU run
 * uses, so the image is the 1-GPU image up to fp32 summation order -- and every group ends with one NCCL reduce of the
 * framebuffer to rank 0 (cb200_reducer_*), overlapped with the next group.  Rank 0 writes the image.
 */
#include "corona_host.h"
#include "corona_b200.h"
#include "corona_b200_render.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9*t.tv_nsec; }

static const float k_fstop[] = {0.5, 0.7, 1.0, 1.4, 2, 2.8, 4, 5.6, 8, 11, 16, 22, 32, 45, 64, 90, 128};   /* src/view.c:71-73 */
static const float k_exposure[] = {60.0f, 30.0f, 15.0f, 8.0f, 4.0f, 2.0f, 1.0f, 0.5f, 1.0/4.0f, 1.0/8.0f, 1.0/15.0f, 1.0/30.0f, 1.0/60.0f,
                                   1.0/125.0f, 1.0/250.0f, 1.0/500.0f, 1.0/1000.0f, 1.0/2000.0f, 1.0/4000.0f, 1.0/8000.0f};   /* src/view.c:75-79 */

typedef struct run_info_t
{
  const char *basename, *sampler, *points, *colour;
  uint64_t spp, prims, rays_closest, rays_shadow, paths;
  uint32_t width, height;
  int gpus;
  double seconds;          /* wall clock of the progressions */
  const cb_camera_t *cam;
  const float *aabb;
  double energy[33]; uint64_t count[33];
}
run_info_t;

/* the reference's sidecar (common_write_sidecar, src/corona_common.c:70-97: every module's print_info) for the modules of this
 * binary; the view block follows view_print_info (src/view.c:726-790) including the path-length energy histogram */
static void write_sidecar(const char *pfm, const run_info_t *R, const float *fb, float gain)
{
  char filename[1500];
  snprintf(filename, sizeof(filename), "%s.txt", pfm);
  FILE *f = fopen(filename, "wb");
  if(!f) return;
  fprintf(f, "corona-13: %s\n", cb200_version());
  fprintf(f, "file     : %s\n", R->basename);
  if(R->aabb[3] < R->aabb[0]) fprintf(f, "aabb     : empty\n");
  else fprintf(f, "aabb     : (%.3f, %.3f)x(%.3f, %.3f)x(%.3f, %.3f) dm^3\n", R->aabb[0], R->aabb[3], R->aabb[1], R->aabb[4], R->aabb[2], R->aabb[5]);
  fprintf(f, "points   : counter generator keyed by (frame, path index, dimension)\n");
  fprintf(f, "prims    : %lu primitives\n", (unsigned long)R->prims);
  accel_print_info(f);
  render_print_info(f);
  fprintf(f, "view     : samples per pixel: %lu (%.2f s/prog) max path vertices %d\n", (unsigned long)R->spp, R->spp ? R->seconds/R->spp : 0.0, 32);
  fprintf(f, "           res %lux%lu\n", (unsigned long)R->width, (unsigned long)R->height);
  fprintf(f, "           elapsed wallclock prog %.2fs on %d gpu%s\n", R->seconds, R->gpus, R->gpus > 1 ? "s" : "");
  fprintf(f, "           active cam 0\n");
  const cb_camera_t *c = R->cam;
  fprintf(f, "camera   : thin lens model\n  focus  : %f\n  film   : %dmm x %dmm\n", c->focus, (int)(c->film_width*100.0f+0.5f), (int)(c->film_height*100.0f + 0.5f));
  if(c->exposure_value > 6) fprintf(f, "         : 1/%.0f f/%.1f %.0fmm iso %d\n", roundf(1.0/k_exposure[c->exposure_value]), k_fstop[c->aperture_value], 100.0*c->focal_length, (int)c->iso);
  else fprintf(f, "         : %.1f\" f/%.1f %.0fmm iso %d\n", k_exposure[c->exposure_value], k_fstop[c->aperture_value], 100.0*c->focal_length, (int)c->iso);
  double sum[3] = {0, 0, 0};
  const uint64_t px = (uint64_t)R->width*R->height;
  for(uint64_t k=0;k<px;k++) for(int i=0;i<3;i++) sum[i] += fb[3*k+i];
  fprintf(f, "           cam 0 average image intensity (rgb): (%f %f %f)\n", gain*sum[0]/px, gain*sum[1]/px, gain*sum[2]/px);
  /* energy per path length, log-scaled over three text lines (view.c:759-790) */
  const int lines = 3;
  const double base = 64.0;
  double max = 0.0;
  for(int k=2;k<=32;k++) if(R->count[k] && R->energy[k] > max) max = R->energy[k];
  fprintf(f, "           ");
  for(int k=2;k<=32;k++) if(k % 10 == 0) fprintf(f, "%d", (k/10)%10); else if(k%5==0) fprintf(f, "|"); else fprintf(f, ".");
  fprintf(f, "\n");
  static const char *bar[] = {" ", "\u2581", "\u2582", "\u2583", "\u2584", "\u2585", "\u2586", "\u2587", "\u2588"};
  for(int h=lines-1;h>=0;h--)
  {
    fprintf(f, "           ");
    for(int k=2;k<=32;k++)
    {
      double level = max > 0.0 ? R->energy[k]/max : 0.0;
      level = log(1.0 + (base-1.0)*level)/log(base);
      const double fill = level*lines - h;
      const int step = fill <= 0 ? 0 : fill >= 1.0 ? 8 : (int)ceil(fill*8.0);
      fputs(bar[step], f);
    }
    fprintf(f, "\n");
  }
  fprintf(f, "filter   : blackman harris\n");
  fprintf(f, "sampler  : %s\n", R->sampler);
  fprintf(f, "mutations: %s\n", R->points);
  fprintf(f, "camera   : %s\n", R->colour);
  fprintf(f, "input    : eRGB via rgb2spec coefficients\n");
  fclose(f);
}

/* machine-readable twin of the sidecar: throughput of this run */
static void write_json(const char *pfm, const run_info_t *R)
{
  char filename[1500];
  snprintf(filename, sizeof(filename), "%s.json", pfm);
  FILE *f = fopen(filename, "wb");
  if(!f) return;
  const double rays = (double)R->rays_closest + (double)R->rays_shadow, s = R->seconds > 0 ? R->seconds : 1e-9;
  fprintf(f, "{\"scene\": \"%s\", \"width\": %u, \"height\": %u, \"spp\": %lu, \"gpus\": %d, \"seconds\": %.6f, \"spp_per_s\": %.3f, "
             "\"paths\": %lu, \"paths_per_s\": %.1f, \"rays_closest\": %lu, \"rays_shadow\": %lu, \"rays_per_s\": %.1f, \"rays_per_path\": %.4f, "
             "\"primitives\": %lu, \"sampler\": \"%s\", \"points\": \"%s\", \"counts_cover\": \"rank 0\", \"path_length_energy\": [",
          R->basename, R->width, R->height, (unsigned long)R->spp, R->gpus, R->seconds, R->spp/s, (unsigned long)R->paths, R->paths/s,
          (unsigned long)R->rays_closest, (unsigned long)R->rays_shadow, rays/s, R->paths ? rays/R->paths : 0.0, (unsigned long)R->prims, R->sampler, R->points);
  for(int k=0;k<=32;k++) fprintf(f, "%s%.9g", k ? ", " : "", R->energy[k]);
  fprintf(f, "], \"path_length_count\": [");
  for(int k=0;k<=32;k++) fprintf(f, "%s%lu", k ? ", " : "", (unsigned long)R->count[k]);
  fprintf(f, "]}\n");
  fclose(f);
}

/* the reference's frame buffer file (include/framebuffer.h:19-36,76-112): header {magic, width, height, channels, flags, gain} and the
 * UN-gained floats; what `--retain-framebuffer` leaves behind as <basename>_<name>_fb00.fb (src/view.c:333-337) */
static int write_fb_file(const char *filename, const float *fb, uint32_t width, uint32_t height, float gain)
{
  struct { uint64_t magic, width, height; uint16_t channels, flags; float gain; } h = {1936686951lu, width, height, 3, 0, gain};
  FILE *f = fopen(filename, "wb");
  if(!f) return 1;
  const size_t n = (size_t)width*height*3;
  const int bad = fwrite(&h, sizeof(h), 1, f) != 1 || fwrite(fb, sizeof(float), n, f) != n;
  fclose(f);
  return bad;
}

int main(int argc, char *argv[])
{
  if(argc < 2)
  {
    fprintf(stderr, "usage: %s <scene.nra2> [-s spp] [-w width] [-h height] [--frame n] [-x [name]] [-c camfile] [--batch n]\n"
                    "          [--sampler pt|ptdl|ptnee] [--points rand|halton] [--colour xyz|rec709] [--coeff file] [--tables file] [--dbor n] [-q]\n", argv[0]);
    return 1;
  }
  const char *scene = argv[1], *coeff = "data/ergb2spec.coeff", *tables = getenv("CORONA_B200_TABLES"), *camfile = 0, *dump = 0;
  char outname[256] = "render";
  uint64_t spp = 0, frame = 1, batch = 0;
  uint32_t width = 1024, height = 576;       /* src/view.c:261-262 */
  int sampler = CB_SAMPLER_PTDL, points = CB_POINTS_RAND, colour = CB_COLOUR_XYZ, quiet = 0, dbor = 0, gpus = 1, retain = 0;
  for(int i=2;i<argc;i++)
  {
    if     (!strcmp(argv[i], "-s") && i+1 < argc) spp = strtoull(argv[++i], 0, 10);
    else if(!strcmp(argv[i], "-w") && i+1 < argc) width = (uint32_t)atol(argv[++i]);
    else if(!strcmp(argv[i], "-h") && i+1 < argc) height = (uint32_t)atol(argv[++i]);
    else if(!strcmp(argv[i], "-c") && i+1 < argc) camfile = argv[++i];
    else if(!strcmp(argv[i], "--frame") && i+1 < argc) frame = strtoull(argv[++i], 0, 10);
    else if(!strcmp(argv[i], "--batch") && i+1 < argc) batch = strtoull(argv[++i], 0, 10);
    else if(!strcmp(argv[i], "-x")) { if(i+1 < argc && argv[i+1][0] != '-') snprintf(outname, sizeof(outname), "%s", argv[++i]); }
    else if(!strcmp(argv[i], "--sampler") && i+1 < argc) { ++i; sampler = !strcmp(argv[i], "pt") ? CB_SAMPLER_PT : !strcmp(argv[i], "ptnee") ? CB_SAMPLER_PTNEE : CB_SAMPLER_PTDL; }
    else if(!strcmp(argv[i], "--points") && i+1 < argc) { ++i; points = !strcmp(argv[i], "halton") ? CB_POINTS_HALTON : CB_POINTS_RAND; }
    else if(!strcmp(argv[i], "--colour") && i+1 < argc) { ++i; colour = !strcmp(argv[i], "rec709") ? CB_COLOUR_REC709 : CB_COLOUR_XYZ; }
    else if(!strcmp(argv[i], "--coeff") && i+1 < argc) coeff = argv[++i];
    else if(!strcmp(argv[i], "--tables") && i+1 < argc) tables = argv[++i];
    else if(!strcmp(argv[i], "--dump-materials") && i+1 < argc) dump = argv[++i];
    else if(!strcmp(argv[i], "--dbor") && i+1 < argc) { dbor = atoi(argv[++i]); dbor = dbor < 0 ? 0 : dbor > 20 ? 20 : dbor; }   /* view.c:291 */
    else if(!strcmp(argv[i], "--gpus") && i+1 < argc) { gpus = atoi(argv[++i]); gpus = gpus < 1 ? 1 : gpus > 64 ? 64 : gpus; }
    else if(!strcmp(argv[i], "--retain-framebuffer")) retain = 1;
    else if(!strcmp(argv[i], "-q")) quiet = 1;
    else if((!strcmp(argv[i], "-t") || !strcmp(argv[i], "-b") || !strcmp(argv[i], "-o")) && i+1 < argc) ++i;   /* cpu threads / backups / timeout: n/a */
  }
  /* ---- multi-GPU: fork the other ranks before anything touches the device; the communicator id travels through a pipe */
  int rank = 0, idpipe[2] = {-1, -1};
  pid_t kids[64];
  if(gpus > 1 && spp && !dump)
  {
    if(dbor > 1) { fprintf(stderr, "[main] --dbor with --gpus > 1 is not supported\n"); return 1; }
    if(pipe(idpipe)) { perror("[main] pipe"); return 1; }
    fflush(stdout); fflush(stderr);
    for(int g=1;g<gpus;g++)
    {
      const pid_t pid = fork();
      if(pid < 0) { perror("[main] fork"); return 1; }
      if(pid == 0) { rank = g; quiet = 1; break; }
      kids[g] = pid;
    }
    if(cb200_set_device(rank)) { fprintf(stderr, "[main] rank %d: %s\n", rank, cb200_last_error()); return 2; }
  }
  else gpus = 1;
  const double t_open = now();
  struct scene_b200_t *s = scene_b200_open(scene, coeff, tables);
  if(!s) { fprintf(stderr, "[main] could not load nra2 file!\n"); return 2; }
  if(!quiet) printf("[main] shader list and geometry files mapped in %.3f seconds\n", now() - t_open);
  if(dump)
  { /* the flattened shader list, for the parser tests: no GPU needed */
    int n = 0;
    const cb_material_t *m = scene_b200_materials(s, &n);
    FILE *f = fopen(dump, "wb");
    if(!f || fwrite(m, sizeof(cb_material_t), n, f) != (size_t)n) { fprintf(stderr, "[main] could not write %s\n", dump); return 2; }
    fclose(f);
    int32_t nm[2] = {0, 0};
    const cb_medium_t *med = scene_b200_media(s, nm, nm + 1);
    if(nm[0] || nm[1])
    { /* <file>.media: count, exterior medium, cb_medium_t[] */
      char name[1100];
      snprintf(name, sizeof(name), "%s.media", dump);
      f = fopen(name, "wb");
      if(!f || fwrite(nm, sizeof(nm), 1, f) != 1 || fwrite(med, sizeof(cb_medium_t), nm[0], f) != (size_t)nm[0]) { fprintf(stderr, "[main] could not write %s\n", name); return 2; }
      fclose(f);
    }
    if(!spp) { scene_b200_free(s); return 0; }
  }
  if(!quiet) { accel_print_info(stdout); render_print_info(stdout); }
  double t0 = now();
  if(scene_b200_prepare(s, width, height, sampler, points, colour, frame, camfile)) { fprintf(stderr, "[main] could not initialise the gpu modules\n"); scene_b200_free(s); return 2; }
  const cb_render_desc_t *d = scene_b200_desc(s);
  if(!quiet) printf("[main] %lu primitives, accel + upload took %.3f seconds\n", (unsigned long)scene_b200_num_prims(s), now() - t0);
  if(!quiet) printf("[display] simulating %lu samples per pixel\n", (unsigned long)spp);
  const uint64_t per_frame = (uint64_t)d->width*d->height;
  if(batch < 1)
  { /* no --batch given: hand the device as many progressions per call as fill its path pool (the image does not depend on
     * the grouping: a path is a function of its index; rt.batch_frames only groups work upstream too, src/view.c:630-638) */
    batch = ((1ull << 23) + per_frame - 1)/per_frame;     /* the library's pool holds max(4 frames, 2^23 paths) */
    if(batch > 64) batch = 64;
  }
  float *fb = (float *)malloc(sizeof(float)*per_frame*3);
  struct render_t *r = scene_b200_render(s);
  if(dbor > 1 && render_b200_set_dbor(r, dbor)) { free(fb); scene_b200_free(s); return 3; }
  if(!getenv("CB200_NO_PATH_STATS")) cb200_render_path_stats((cb200_render_t *)render_b200_handle(r), 1);   /* view->stat_enery / stat_cnt for the sidecar */
  t0 = now();
  if(gpus > 1)
  { /* groups of `batch` progressions, group b rendered by rank b % gpus; every rank runs the same number of rounds (the reduce is
     * a collective) plus one for the paths still in flight at the end */
    char id[CB200_COMM_ID_BYTES];
    if(rank == 0)
    {
      if(cb200_comm_unique_id(id)) { fprintf(stderr, "[main] %s\n", cb200_last_error()); return 3; }
      for(int g=1;g<gpus;g++) if(write(idpipe[1], id, sizeof(id)) != (ssize_t)sizeof(id)) { perror("[main] write"); return 3; }
    }
    else if(read(idpipe[0], id, sizeof(id)) != (ssize_t)sizeof(id)) { fprintf(stderr, "[main] rank %d: no communicator id\n", rank); return 3; }
    cb200_render_t *dev = (cb200_render_t *)render_b200_handle(r);
    cb200_reducer_t *q = cb200_reducer_create(dev, id, rank, gpus, d->width, d->height);
    if(!q) { fprintf(stderr, "[main] rank %d: %s\n", rank, cb200_last_error()); return 3; }
    const uint64_t groups = (spp + batch - 1)/batch, rounds = (groups + gpus - 1)/gpus;
    int rc = 0;
    for(uint64_t k=0;k<rounds && !rc;k++)
    {
      const uint64_t b = k*gpus + rank;
      rc = cb200_reducer_begin(q, k, 0);
      if(!rc && b < groups)
      {
        const uint64_t first = b*batch, n = (spp - first) < batch ? (spp - first) : batch;
        rc = render_b200_pass(r, first*per_frame, n*per_frame, 0);
      }
      if(!rc) rc = cb200_reducer_end(q, k, 0, 0);
    }
    if(!rc) rc = cb200_reducer_begin(q, rounds, 0);
    if(!rc) rc = cb200_render_flush(dev, 0);
    if(!rc) rc = cb200_reducer_end(q, rounds, 0, 0);
    if(!rc) rc = cb200_reducer_finish(q, rank == 0 ? fb : 0);
    if(rc) fprintf(stderr, "[main] rank %d: %s\n", rank, cb200_last_error());
    cb200_reducer_destroy(q);
    if(rank != 0) { free(fb); scene_b200_free(s); return rc ? 3 : 0; }
    int failed = rc;
    for(int g=1;g<gpus;g++) { int st = 0; if(waitpid(kids[g], &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st)) failed = 1; }
    if(failed) { fprintf(stderr, "[main] a rank failed\n"); free(fb); scene_b200_free(s); return 3; }
  }
  else
  {
    uint64_t done = 0;
    while(done < spp)
    { /* run(): view_render() per progression (src/main.c:388-412, src/view.c:630-645) */
      const uint64_t n = (spp - done) < batch ? (spp - done) : batch;
      if(render_b200_pass(r, done*per_frame, n*per_frame, 0)) { free(fb); scene_b200_free(s); return 3; }
      done += n;
    }
    if(render_b200_finish(r, fb)) { free(fb); scene_b200_free(s); return 3; }
  }
  const double dt = now() - t0;
  if(!quiet && spp) printf("[main] rendered %lu frames in an average of %.6f s/frame\n", (unsigned long)spp, dt/spp);
  char filename[1400];
  snprintf(filename, sizeof(filename), "%s%s_fb00.pfm", scene_b200_basename(s), outname);
  const float gain = spp ? d->camera.iso/(100.0f*(float)spp) : 0.0f;    /* src/view.c:656 */
  if(scene_b200_write_pfm(filename, fb, d->width, d->height, gain)) { fprintf(stderr, "[main] could not write %s\n", filename); free(fb); scene_b200_free(s); return 4; }
  {
    run_info_t R;
    memset(&R, 0, sizeof(R));
    cb_render_stats_t st;
    memset(&st, 0, sizeof(st));
    cb200_render_t *dev = (cb200_render_t *)render_b200_handle(r);
    cb200_render_stats(dev, &st);
    cb200_render_get_path_stats(dev, R.energy, R.count);
    R.basename = scene_b200_basename(s);
    R.sampler = sampler == CB_SAMPLER_PT ? "pathtracer" : sampler == CB_SAMPLER_PTNEE ? "pathtracer with next event estimation only" : "pathtracer with next event estimation and mis";
    R.points = points == CB_POINTS_HALTON ? "halton points" : "none";
    R.colour = colour == CB_COLOUR_REC709 ? "linear rec709" : "CIE XYZ";
    R.spp = spp; R.prims = scene_b200_num_prims(s); R.width = d->width; R.height = d->height; R.gpus = gpus; R.seconds = dt;
    R.rays_closest = st.rays_closest; R.rays_shadow = st.rays_shadow; R.paths = st.paths;
    R.cam = &d->camera; R.aabb = scene_b200_aabb(s);
    write_sidecar(filename, &R, fb, gain);
    write_json(filename, &R);
    if(retain)
    {
      char fbname[1400];
      snprintf(fbname, sizeof(fbname), "%s_%s_fb00.fb", scene_b200_basename(s), outname);
      if(write_fb_file(fbname, fb, d->width, d->height, gain)) fprintf(stderr, "[main] could not write %s\n", fbname);
    }
  }
  for(int l=0;dbor>1&&l<dbor;l++)
  { /* view_write_images, src/view.c:553-556 (the gcov buffers next to them are never written to upstream and are not produced) */
    snprintf(filename, sizeof(filename), "%s%s_dbor%02d.pfm", scene_b200_basename(s), outname, l);
    if(render_b200_dbor(r, l, fb) || scene_b200_write_pfm(filename, fb, d->width, d->height, gain)) { fprintf(stderr, "[main] could not write %s\n", filename); free(fb); scene_b200_free(s); return 4; }
  }
  if(!quiet) printf("[main] saving framebuffers to %s%s\n", scene_b200_basename(s), outname);
  free(fb);
  scene_b200_free(s);
  return 0;
}
