/* main_b200.c -- `corona_b200`: the reference's offline command line for the hot path, on the GPU modules.
 *
 *   corona_b200 <scene.nra2> [-s spp] [-w width] [-h height] [--frame n] [-x [name]] [-c camfile] [--batch n]
 *               [--sampler pt|ptdl|ptnee] [--points rand|halton] [--colour xyz|rec709] [--coeff file] [--tables file]
 *               [--dump-materials file] [--dbor n] [--gpus n] [-q]
 *
 * Same arguments and defaults as the reference binary where they exist there (src/main.c:250-282,415-437,
 * src/view.c:262-297, src/display.d/null.c:47-56): -s samples per pixel then write the image and quit, -w/-h frame size
 * (padded to multiples of 32), --frame = rt.anim_frame (seeds the point sampler, default 1), -x output name (default
 * `render'), data/ergb2spec.coeff relative to the working directory.  --sampler / --points / --colour stand in for the
 * reference's compile-time MOD_sampler / MOD_pointsampler / COL_camera.  Writes <basename><name>_fb00.pfm like
 * view_write_images (src/view.c:549); --dbor n (src/view.c:291) adds the outlier rejection cascade of view_splat_col and its
 * <basename><name>_dbor%02d.pfm files (view.c:553-556).  Without a CUDA device it refuses: there is no CPU path in this binary.
 *
 * --gpus n: the samples per pixel are split over n GPUs of the box, one process per GPU (forked before any device
 * work; this process is rank 0).  Rank g renders the progression groups g, g+n, ... -- the same path-index ranges a 1-GPU run
 * uses, so the image is the 1-GPU image up to fp32 summation order -- and every group ends with one NCCL reduce of the
 * framebuffer to rank 0 (cb200_reducer_*), overlapped with the next group.  Rank 0 writes the image.
 */
#include "corona_host.h"
#include "corona_b200.h"
#include "corona_b200_render.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9*t.tv_nsec; }

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
  int sampler = CB_SAMPLER_PTDL, points = CB_POINTS_RAND, colour = CB_COLOUR_XYZ, quiet = 0, dbor = 0, gpus = 1;
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
    batch = ((1ull << 22) + per_frame - 1)/per_frame;
    if(batch > 64) batch = 64;
  }
  float *fb = (float *)malloc(sizeof(float)*per_frame*3);
  struct render_t *r = scene_b200_render(s);
  if(dbor > 1 && render_b200_set_dbor(r, dbor)) { free(fb); scene_b200_free(s); return 3; }
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
