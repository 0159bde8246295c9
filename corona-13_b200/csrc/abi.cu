// abi.cu -- extern "C" entry points of libcorona_b200.so (declared in include/corona_b200.h)
#include "internal.h"
#include <atomic>
#include <mutex>
#include <vector>
#include <cstring>
#include <cstdio>
#include <float.h>

static thread_local std::string g_error;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_sm_count[64];   // per device (the current device can change between calls: multi-GPU hosts)

void cb200_set_error(const std::string &msg) { g_error = msg; }
int cb200_cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
  char buf[512];
  snprintf(buf, sizeof(buf), "%s:%d: %s -> %s", file, line, what, cudaGetErrorString(e));
  g_error = buf;
  if(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return CB200_ERR_NO_DEVICE;
  if(e == cudaErrorMemoryAllocation) return CB200_ERR_NOMEM;
  return CB200_ERR_CUDA;
}
void cb200_count_launch(uint64_t n) { g_launches += n; }
int cb200_sm_count_cached()
{
  int dev = 0;
  cudaGetDevice(&dev);
  if(dev < 0 || dev >= 64) dev = 0;
  int n = g_sm_count[dev].load();
  if(!n)
  {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if(n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return n;
}

extern "C" {

const char *cb200_version(void) { return "corona-13_b200 0.1 (sm_100a)"; }
const char *cb200_last_error(void) { return g_error.c_str(); }
uint64_t cb200_launch_count(void) { return g_launches.load(); }

int cb200_device_count(void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if(e != cudaSuccess) { cb200_cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); return 0; }
  return n;
}

int cb200_set_device(int device)
{
  if(device < 0 || device >= cb200_device_count()) { if(g_error.empty()) g_error = "no such CUDA device"; return CB200_ERR_NO_DEVICE; }
  CB_CUDA(cudaSetDevice(device));
  return 0;
}
int cb200_sm_count(void) { if(cb200_device_count() < 1) return 0; return cb200_sm_count_cached(); }

void *cb200_malloc(size_t bytes) { void *p = nullptr; CB_CUDA_NULL(cudaMalloc(&p, bytes ? bytes : 1)); return p; }
int cb200_free(void *p) { CB_CUDA(cudaFree(p)); return 0; }
void *cb200_malloc_host(size_t bytes) { void *p = nullptr; CB_CUDA_NULL(cudaMallocHost(&p, bytes ? bytes : 1)); return p; }
int cb200_free_host(void *p) { CB_CUDA(cudaFreeHost(p)); return 0; }
int cb200_memcpy_h2d(void *d, const void *h, size_t bytes, void *stream)
{ CB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream)); return 0; }
int cb200_memcpy_d2h(void *h, const void *d, size_t bytes, void *stream)
{ CB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream)); return 0; }
int cb200_stream_sync(void *stream) { CB_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return 0; }

// ---------------------------------------------------------------------------------------------
cb200_scene_t *cb200_scene_create(const cb_shape_t *shapes, int num_shapes)
{
  if(num_shapes < 0 || (num_shapes > 0 && !shapes)) { g_error = "scene_create: bad arguments"; return nullptr; }
  if(cb200_device_count() < 1) { if(g_error.empty()) g_error = "no CUDA device"; return nullptr; }
  cb200_scene *s = new cb200_scene();
  s->num_shapes = 0; s->num_prims = s->num_vtx = s->num_vtxidx = 0; s->any_mb = 0; s->any_analytic = 0;
  s->d_vtx = nullptr; s->d_vtxidx = nullptr; s->d_shapes = nullptr; s->d_primid = nullptr;
  cudaGetDevice(&s->device);
  for(int i=0;i<num_shapes;i++) s->h_material.push_back(shapes[i].material);
  s->num_shapes = num_shapes;
  std::vector<ShapeDev> sd(num_shapes > 0 ? num_shapes : 1);
  for(int i=0;i<num_shapes;i++)
  {
    sd[i].vtx_off = s->num_vtx; sd[i].vtxidx_off = s->num_vtxidx;
    s->num_vtx += shapes[i].num_vtx; s->num_vtxidx += shapes[i].num_vtxidx; s->num_prims += shapes[i].num_prims;
    if(shapes[i].num_prims && (!shapes[i].primid || !shapes[i].vtxidx || !shapes[i].vtx))
    { g_error = "scene_create: shape with null arrays"; delete s; return nullptr; }
  }
  // global primid list with the shape id patched in (prims_allocate_index, src/prims.c:741-757)
  std::vector<uint64_t> primid(s->num_prims ? s->num_prims : 1);
  uint64_t k = 0;
  for(int i=0;i<num_shapes;i++)
    for(uint64_t j=0;j<shapes[i].num_prims;j++)
    {
      const uint64_t p = cb_primid_with_shapeid(shapes[i].primid[j], (uint32_t)i);
      const uint32_t vcnt = cb_primid_vcnt(p);
      if(vcnt < 1 || vcnt > 4) { g_error = "scene_create: unsupported primitive type (shells are out of scope)"; delete s; return nullptr; }
      if(cb_primid_vi(p) + vcnt > shapes[i].num_vtxidx) { g_error = "scene_create: vertex index out of range"; delete s; return nullptr; }
      { // ... and the vertices those indices name: vtx[(mb+1)*v (+mb)] (include/geo.h:108-138); an index outside the array would
        // make the build and shading kernels read outside their buffers
        const uint64_t mb = cb_primid_mb(p), nv = shapes[i].num_vtx;
        const cb_vtxidx_t *ix = shapes[i].vtxidx + cb_primid_vi(p);
        for(uint32_t c=0;c<vcnt;c++)
          if((mb + 1)*(uint64_t)ix[c].v + mb >= nv) { g_error = "scene_create: vertex index out of range"; delete s; return nullptr; }
      }
      if(cb_primid_mb(p)) s->any_mb = 1;
      if(cb_primid_vcnt(p) < CB_PRIM_TRI) s->any_analytic = 1;
      primid[k++] = p;
    }
#define SC(call) do { cudaError_t e__ = (call); if(e__ != cudaSuccess) { cb200_cuda_fail(e__, #call, __FILE__, __LINE__); cb200_scene_destroy(s); return nullptr; } } while(0)
  SC(cudaMalloc(&s->d_vtx, (s->num_vtx ? s->num_vtx : 1)*sizeof(cb_vtx_t)));
  SC(cudaMalloc(&s->d_vtxidx, (s->num_vtxidx ? s->num_vtxidx : 1)*sizeof(cb_vtxidx_t)));
  SC(cudaMalloc(&s->d_shapes, sd.size()*sizeof(ShapeDev)));
  SC(cudaMalloc(&s->d_primid, primid.size()*sizeof(uint64_t)));
  for(int i=0;i<num_shapes;i++)
  {
    if(shapes[i].num_vtx) SC(cudaMemcpy(s->d_vtx + sd[i].vtx_off, shapes[i].vtx, shapes[i].num_vtx*sizeof(cb_vtx_t), cudaMemcpyHostToDevice));
    if(shapes[i].num_vtxidx) SC(cudaMemcpy(s->d_vtxidx + sd[i].vtxidx_off, shapes[i].vtxidx, shapes[i].num_vtxidx*sizeof(cb_vtxidx_t), cudaMemcpyHostToDevice));
  }
  SC(cudaMemcpy(s->d_shapes, sd.data(), sd.size()*sizeof(ShapeDev), cudaMemcpyHostToDevice));
  SC(cudaMemcpy(s->d_primid, primid.data(), primid.size()*sizeof(uint64_t), cudaMemcpyHostToDevice));
#undef SC
  return s;
}

void cb200_scene_destroy(cb200_scene_t *s)
{
  if(!s) return;
  cudaFree(s->d_vtx); cudaFree(s->d_vtxidx); cudaFree(s->d_shapes); cudaFree(s->d_primid);
  delete s;
}
uint64_t cb200_scene_num_prims(const cb200_scene_t *s) { return s ? s->num_prims : 0; }

// ---------------------------------------------------------------------------------------------
static cb200_accel *accel_new(cb200_scene *s)
{
  cb200_accel *a = new cb200_accel();
  memset(a, 0, sizeof(*a));
  a->scene = s;
  return a;
}

cb200_accel_t *cb200_accel_build(cb200_scene_t *s, const float *ghost_aabb, uint64_t *primid_out)
{
  if(!s) { g_error = "accel_build: null scene"; return nullptr; }
  cb200_accel *a = accel_new(s);
  if(cb200_build_lbvh(a, ghost_aabb)) { cb200_accel_destroy(a); return nullptr; }
  if(primid_out && s->num_prims)
  {
    cudaError_t e = cudaMemcpy(primid_out, a->d_primid, s->num_prims*sizeof(uint64_t), cudaMemcpyDeviceToHost);
    if(e != cudaSuccess) { cb200_cuda_fail(e, "copy primid", __FILE__, __LINE__); cb200_accel_destroy(a); return nullptr; }
  }
  return a;
}

static int tree_depth(const cb_qbvh_node_t *nodes, uint64_t num_nodes, uint64_t num_prims)
{ // iterative depth of an imported tree, with bounds validation; -1 when malformed
  if(num_nodes == 0) return -1;
  std::vector<std::pair<uint64_t,int>> st;
  st.push_back({0, 1});
  int depth = 0;
  uint64_t visited = 0;
  while(!st.empty())
  {
    auto [ni, d] = st.back(); st.pop_back();
    if(++visited > num_nodes) return -1;   // cycle or shared node
    if(d > depth) depth = d;
    const cb_qbvh_node_t &n = nodes[ni];
    if(n.axis0 < 0 || n.axis0 > 2 || n.axis00 < 0 || n.axis00 > 2 || n.axis01 < 0 || n.axis01 > 2) return -1;
    for(int c=0;c<4;c++)
    {
      const uint64_t ch = n.child[c];
      if(ch & CB_LEAF_BIT)
      {
        const uint64_t beg = (ch ^ CB_LEAF_BIT) >> 5, cnt = ch & 31;
        if(cnt && beg + cnt > num_prims) return -1;
      }
      else
      {
        if(ch >= num_nodes) return -1;
        st.push_back({ch, d+1});
      }
    }
  }
  return depth;
}

cb200_accel_t *cb200_accel_import_qbvh(cb200_scene_t *s, const cb_qbvh_node_t *nodes, uint64_t num_nodes,
                                       const uint64_t *primid_permuted, const float aabb[6])
{
  if(!s || !nodes || num_nodes == 0 || (s->num_prims && !primid_permuted)) { g_error = "accel_import_qbvh: bad arguments"; return nullptr; }
  const int depth = tree_depth(nodes, num_nodes, s->num_prims);
  if(depth < 0 || depth > 101) { g_error = "accel_import_qbvh: malformed tree"; return nullptr; }
  cb200_accel *a = accel_new(s);
  a->imported = 1;
  a->depth = depth;
#define AC(call) do { cudaError_t e__ = (call); if(e__ != cudaSuccess) { cb200_cuda_fail(e__, #call, __FILE__, __LINE__); cb200_accel_destroy(a); return nullptr; } } while(0)
  AC(cudaMalloc(&a->d_nodes, num_nodes*sizeof(Node256)));
  AC(cudaMemcpy(a->d_nodes, nodes, num_nodes*sizeof(Node256), cudaMemcpyHostToDevice));
  AC(cudaMalloc(&a->d_primid, (s->num_prims ? s->num_prims : 1)*sizeof(uint64_t)));
  if(s->num_prims) AC(cudaMemcpy(a->d_primid, primid_permuted, s->num_prims*sizeof(uint64_t), cudaMemcpyHostToDevice));
#undef AC
  a->dev.nodes = a->d_nodes;
  a->dev.num_nodes = num_nodes;
  a->dev.mb = 1;    // reference layout: always interpolates the two box sets, even for static scenes
  a->dev.imported = 1;
  if(aabb) memcpy(a->aabb, aabb, sizeof(float)*6);
  if(cb200_build_records(a, 0)) { cb200_accel_destroy(a); return nullptr; }
  if(cudaStreamSynchronize(0) != cudaSuccess) { g_error = "accel_import_qbvh: sync failed"; cb200_accel_destroy(a); return nullptr; }
  return a;
}

void cb200_accel_destroy(cb200_accel_t *a)
{
  if(!a) return;
  cudaFree(a->d_nodes); cudaFree(a->d_recs); cudaFree(a->d_primid); cudaFree(a->d_nodes8);
  delete a;
}
uint64_t cb200_accel_num_nodes(const cb200_accel_t *a) { return a ? a->dev.num_nodes : 0; }
int cb200_accel_depth(const cb200_accel_t *a) { return a ? a->depth : 0; }
int cb200_accel_aabb(const cb200_accel_t *a, float aabb[6])
{
  if(!a || !aabb) { g_error = "accel_aabb: bad arguments"; return CB200_ERR_ARG; }
  memcpy(aabb, a->aabb, sizeof(float)*6);
  return 0;
}
int cb200_accel_layout(const cb200_accel_t *a, uint32_t *node_bytes, uint32_t *prim_bytes)
{
  if(!a) { g_error = "accel_layout: null accel"; return CB200_ERR_ARG; }
  if(node_bytes) *node_bytes = cb200_use_wide8(a) ? (uint32_t)sizeof(Node8) : a->dev.mb ? 256 : 128;
  if(prim_bytes) *prim_bytes = a->dev.rec_units*64;
  return 0;
}

int cb200_accel_set_traversal(cb200_accel_t *a, int mode)
{
  if(!a || (mode != CB200_TRAVERSAL_EXACT4 && mode != CB200_TRAVERSAL_WIDE8)) { g_error = "accel_set_traversal: bad arguments"; return CB200_ERR_ARG; }
  if(mode == CB200_TRAVERSAL_WIDE8 && !a->dev.nodes8)
  { g_error = "accel_set_traversal: this accel has no 8-wide tree (motion blur, imported reference tree, or built with CB200_BUILD_WIDE8=0)"; return CB200_ERR_UNSUPPORTED; }
  a->traversal = mode;
  return 0;
}
int cb200_accel_traversal(const cb200_accel_t *a) { return a && cb200_use_wide8(a) ? CB200_TRAVERSAL_WIDE8 : CB200_TRAVERSAL_EXACT4; }

int cb200_accel_export_qbvh(const cb200_accel_t *a, cb_qbvh_node_t *nodes, uint64_t cap, uint64_t *primid_out)
{
  if(!a || !nodes || cap < a->dev.num_nodes) { g_error = "accel_export_qbvh: bad arguments"; return CB200_ERR_ARG; }
  const uint64_t n = a->dev.num_nodes;
  if(a->dev.mb) CB_CUDA(cudaMemcpy(nodes, a->d_nodes, n*sizeof(Node256), cudaMemcpyDeviceToHost));
  else
  {
    std::vector<Node128> tmp(n);
    CB_CUDA(cudaMemcpy(tmp.data(), a->d_nodes, n*sizeof(Node128), cudaMemcpyDeviceToHost));
    for(uint64_t i=0;i<n;i++)
    {
      memcpy(nodes[i].aabb0, tmp[i].aabb0, sizeof(tmp[i].aabb0));
      memcpy(nodes[i].aabb1, tmp[i].aabb0, sizeof(tmp[i].aabb0));
      const uint32_t ax = (uint32_t)(tmp[i].child[0] >> CB_AXIS_SHIFT) & 63u;
      nodes[i].child[0] = tmp[i].child[0] & CB_CHILD_MASK;
      for(int c=1;c<4;c++) nodes[i].child[c] = tmp[i].child[c];
      nodes[i].axis0 = ax & 3; nodes[i].axis00 = (ax >> 2) & 3; nodes[i].axis01 = (ax >> 4) & 3;
      nodes[i].parent = ~0ull;
    }
    for(uint64_t i=0;i<n;i++)
      for(int c=0;c<4;c++) if(!(nodes[i].child[c] & CB_LEAF_BIT)) nodes[nodes[i].child[c]].parent = i;
  }
  if(primid_out && a->dev.num_prims)
    CB_CUDA(cudaMemcpy(primid_out, a->d_primid, a->dev.num_prims*sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return 0;
}

// ---------------------------------------------------------------------------------------------
int cb200_accel_intersect_dev(const cb200_accel_t *a, const void *d_rays, const void *d_max_dist, void *d_out, uint64_t n, void *stream)
{
  if(!a || (n && (!d_rays || !d_out))) { g_error = "accel_intersect_dev: bad arguments"; return CB200_ERR_ARG; }
  return cb200_launch_intersect(a, (const cb_ray_t *)d_rays, (const float *)d_max_dist, (cb_hitrec_t *)d_out, n, (cudaStream_t)stream, nullptr);
}

int cb200_accel_visible_dev(const cb200_accel_t *a, const void *d_rays, const void *d_max_dist, void *d_out, uint64_t n, void *stream)
{
  if(!a || (n && (!d_rays || !d_out || !d_max_dist))) { g_error = "accel_visible_dev: bad arguments"; return CB200_ERR_ARG; }
  return cb200_launch_visible(a, (const cb_ray_t *)d_rays, (const float *)d_max_dist, (int32_t *)d_out, n, (cudaStream_t)stream);
}

int cb200_accel_intersect_counted(const cb200_accel_t *a, const void *d_rays, const void *d_max_dist, void *d_out, uint64_t n, uint64_t counters[4])
{
  if(!a || !counters || (n && (!d_rays || !d_out))) { g_error = "accel_intersect_counted: bad arguments"; return CB200_ERR_ARG; }
  unsigned long long *d_cnt = nullptr;
  CB_CUDA(cudaMalloc(&d_cnt, 4*sizeof(unsigned long long)));
  CB_CUDA(cudaMemset(d_cnt, 0, 4*sizeof(unsigned long long)));
  int rc = cb200_launch_intersect(a, (const cb_ray_t *)d_rays, (const float *)d_max_dist, (cb_hitrec_t *)d_out, n, 0, d_cnt);
  if(!rc)
  {
    cudaError_t e = cudaMemcpy(counters, d_cnt, 4*sizeof(uint64_t), cudaMemcpyDeviceToHost);
    if(e != cudaSuccess) rc = cb200_cuda_fail(e, "copy counters", __FILE__, __LINE__);
  }
  cudaFree(d_cnt);
  return rc;
}

} // extern "C"

// host-buffer entry points: stage through device buffers in chunks so that arbitrarily large batches work and the upload of
// chunk k+1, the traversal of chunk k and the download of chunk k-1 overlap (four streams; pinned host memory is the caller's
// choice -- pageable buffers still work, the copies then run one after the other).  The staging buffers and streams are
// allocated once per process and kept: a cudaMalloc / cudaFree pair per call cost as much as the traversal of a 4 Mi-ray batch.
// Calls are serialised by a mutex (the pinned-worker callers of the single-ray entry points each hand in a batch of one).
namespace {
struct Staging
{
  static const int NS = 4;
  static const uint64_t CH = 1ull << 19;          // rays per chunk: 21 MB up, 12.6 MB down
  cudaStream_t st[NS] = {nullptr, nullptr, nullptr, nullptr};
  cb_ray_t *d_rays[NS] = {nullptr, nullptr, nullptr, nullptr};
  float *d_md[NS] = {nullptr, nullptr, nullptr, nullptr};
  void *d_out[NS] = {nullptr, nullptr, nullptr, nullptr};
  int device = -1;
  std::mutex mutex;
  void release()
  {
    for(int k=0;k<NS;k++)
    {
      if(st[k]) cudaStreamDestroy(st[k]);
      cudaFree(d_rays[k]); cudaFree(d_md[k]); cudaFree(d_out[k]);
      st[k] = nullptr; d_rays[k] = nullptr; d_md[k] = nullptr; d_out[k] = nullptr;
    }
    device = -1;
  }
  int prepare()
  {
    int dev = 0;
    if(cudaGetDevice(&dev) != cudaSuccess) return CB200_ERR_CUDA;
    if(device == dev) return 0;
    if(device >= 0) release();   // (buffers of another device: start over on this one)
    for(int k=0;k<NS;k++)
    {
      if(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking) != cudaSuccess) { st[k] = nullptr; release(); return CB200_ERR_CUDA; }
      if(cudaMalloc(&d_rays[k], CH*sizeof(cb_ray_t)) != cudaSuccess || cudaMalloc(&d_md[k], CH*sizeof(float)) != cudaSuccess ||
         cudaMalloc(&d_out[k], CH*sizeof(cb_hitrec_t)) != cudaSuccess) { release(); return CB200_ERR_NOMEM; }
    }
    device = dev;
    return 0;
  }
};
Staging g_staging;
}

template<typename OUT, typename LAUNCH>
static int run_chunked(const cb_ray_t *rays, const float *max_dist, OUT *out, uint64_t n, LAUNCH launch)
{
  static_assert(sizeof(OUT) <= sizeof(cb_hitrec_t), "staging output slot");
  std::lock_guard<std::mutex> lock(g_staging.mutex);
  int rc = g_staging.prepare();
  if(rc) { g_error = "intersect_n: staging allocation failed"; return rc; }
  const uint64_t CH = Staging::CH;
  uint64_t k = 0;
  for(uint64_t off=0; !rc && off<n; off+=CH, k=(k+1)%Staging::NS)
  {
    const uint64_t m = (n - off) < CH ? (n - off) : CH;
    cudaStream_t st = g_staging.st[k];
    OUT *d_out = reinterpret_cast<OUT *>(g_staging.d_out[k]);
    cudaError_t e = cudaMemcpyAsync(g_staging.d_rays[k], rays + off, m*sizeof(cb_ray_t), cudaMemcpyHostToDevice, st);
    if(e == cudaSuccess && max_dist) e = cudaMemcpyAsync(g_staging.d_md[k], max_dist + off, m*sizeof(float), cudaMemcpyHostToDevice, st);
    if(e != cudaSuccess) { rc = cb200_cuda_fail(e, "h2d", __FILE__, __LINE__); break; }
    rc = launch(g_staging.d_rays[k], max_dist ? g_staging.d_md[k] : nullptr, d_out, m, st);
    if(rc) break;
    e = cudaMemcpyAsync(out + off, d_out, m*sizeof(OUT), cudaMemcpyDeviceToHost, st);
    if(e != cudaSuccess) { rc = cb200_cuda_fail(e, "d2h", __FILE__, __LINE__); break; }
  }
  for(int j=0;j<Staging::NS;j++)
  {
    cudaError_t e = cudaStreamSynchronize(g_staging.st[j]);
    if(e != cudaSuccess && !rc) rc = cb200_cuda_fail(e, "sync", __FILE__, __LINE__);
  }
  return rc;
}

// Small batches -- above all the batch of ONE that accel_intersect / accel_visible of an unpatched reference renderer hands in from
// each of its pinned worker threads (src/pathspace.c:763, :325) -- skip the staging buffers, their mutex and the three copies: every
// host thread owns a stream and a block of mapped pinned memory; the rays are written there, the traversal kernel reads them and
// writes its answers over the bus, and the call costs one launch and one stream synchronisation.  Calls from different threads
// run concurrently.  (Never freed: the reference's workers live as long as the process.)
namespace {
struct SmallSlot
{
  static const uint64_t CAP = 128;
  int device = -1;
  cudaStream_t st = nullptr;
  unsigned char *h = nullptr, *d = nullptr;
  cb_ray_t *rays(unsigned char *b) const { return reinterpret_cast<cb_ray_t *>(b); }
  float *md(unsigned char *b) const { return reinterpret_cast<float *>(b + CAP*sizeof(cb_ray_t)); }
  void *out(unsigned char *b) const { return b + CAP*(sizeof(cb_ray_t) + sizeof(float)); }
  int prepare()
  {
    int dev = 0;
    if(cudaGetDevice(&dev) != cudaSuccess) return CB200_ERR_CUDA;
    if(device == dev) return 0;
    if(h) { cudaFreeHost(h); h = nullptr; }
    if(st) { cudaStreamDestroy(st); st = nullptr; }
    device = -1;
    if(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { st = nullptr; return CB200_ERR_CUDA; }
    void *p = nullptr, *q = nullptr;
    if(cudaHostAlloc(&p, CAP*(sizeof(cb_ray_t) + sizeof(float) + sizeof(cb_hitrec_t)), cudaHostAllocMapped) != cudaSuccess) return CB200_ERR_NOMEM;
    if(cudaHostGetDevicePointer(&q, p, 0) != cudaSuccess) { cudaFreeHost(p); return CB200_ERR_CUDA; }
    h = static_cast<unsigned char *>(p); d = static_cast<unsigned char *>(q);
    device = dev;
    return 0;
  }
};
thread_local SmallSlot t_small;
}

template<typename OUT, typename LAUNCH>
static int run_small(int device, const cb_ray_t *rays, const float *max_dist, OUT *out, uint64_t n, LAUNCH launch)
{
  // a worker thread that has never touched CUDA starts on device 0: follow the accel to the device it lives on
  int cur = -1;
  if(cudaGetDevice(&cur) != cudaSuccess) return CB200_ERR_CUDA;
  if(cur != device && cudaSetDevice(device) != cudaSuccess) { g_error = "intersect_n: cannot select the accel's device"; return CB200_ERR_CUDA; }
  SmallSlot &s = t_small;
  int rc = s.prepare();
  if(rc) { g_error = "intersect_n: pinned slot allocation failed"; return rc; }
  memcpy(s.rays(s.h), rays, n*sizeof(cb_ray_t));
  if(max_dist) memcpy(s.md(s.h), max_dist, n*sizeof(float));
  rc = launch(s.rays(s.d), max_dist ? s.md(s.d) : nullptr, static_cast<OUT *>(s.out(s.d)), n, s.st);
  if(rc) return rc;
  cudaError_t e = cudaStreamSynchronize(s.st);
  if(e != cudaSuccess) return cb200_cuda_fail(e, "sync", __FILE__, __LINE__);
  memcpy(out, s.out(s.h), n*sizeof(OUT));
  return 0;
}

extern "C" {

int cb200_accel_intersect_n(const cb200_accel_t *a, const cb_ray_t *rays, const float *max_dist, cb_hitrec_t *out, uint64_t n)
{
  if(!a || (n && (!rays || !out))) { g_error = "accel_intersect_n: bad arguments"; return CB200_ERR_ARG; }
  if(n == 0) return 0;
  if(n <= SmallSlot::CAP)
    return run_small<cb_hitrec_t>(a->scene->device, rays, max_dist, out, n,
      [a](cb_ray_t *dr, float *dm, cb_hitrec_t *dout, uint64_t m, cudaStream_t s) { return cb200_launch_intersect(a, dr, dm, dout, m, s, nullptr); });
  return run_chunked<cb_hitrec_t>(rays, max_dist, out, n,
    [a](cb_ray_t *dr, float *dm, cb_hitrec_t *dout, uint64_t m, cudaStream_t s) { return cb200_launch_intersect(a, dr, dm, dout, m, s, nullptr); });
}

int cb200_accel_closest_n(const cb200_accel_t *a, cb_ray_t *rays, cb_hitrec_t *io, const float *centre, uint64_t n)
{
  if(!a || (n && (!rays || !io || !centre))) { g_error = "accel_closest_n: bad arguments"; return CB200_ERR_ARG; }
  if(n == 0) return 0;
  cb_ray_t *d_r = nullptr; cb_hitrec_t *d_io = nullptr; float *d_c = nullptr;
  CB_CUDA(cudaMalloc(&d_r, n*sizeof(cb_ray_t))); CB_CUDA(cudaMalloc(&d_io, n*sizeof(cb_hitrec_t))); CB_CUDA(cudaMalloc(&d_c, n*sizeof(float)));
  CB_CUDA(cudaMemcpy(d_r, rays, n*sizeof(cb_ray_t), cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d_io, io, n*sizeof(cb_hitrec_t), cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(d_c, centre, n*sizeof(float), cudaMemcpyHostToDevice));
  int rc = cb200_launch_closest(a, d_r, d_io, d_c, n, 0);
  if(!rc)
  {
    CB_CUDA(cudaMemcpy(rays, d_r, n*sizeof(cb_ray_t), cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemcpy(io, d_io, n*sizeof(cb_hitrec_t), cudaMemcpyDeviceToHost));
  }
  cudaFree(d_r); cudaFree(d_io); cudaFree(d_c);
  return rc;
}

int cb200_accel_visible_n(const cb200_accel_t *a, const cb_ray_t *rays, const float *max_dist, int32_t *out, uint64_t n)
{
  if(!a || (n && (!rays || !out || !max_dist))) { g_error = "accel_visible_n: bad arguments"; return CB200_ERR_ARG; }
  if(n == 0) return 0;
  if(n <= SmallSlot::CAP)
    return run_small<int32_t>(a->scene->device, rays, max_dist, out, n,
      [a](cb_ray_t *dr, float *dm, int32_t *dout, uint64_t m, cudaStream_t s) { return cb200_launch_visible(a, dr, dm, dout, m, s); });
  return run_chunked<int32_t>(rays, max_dist, out, n,
    [a](cb_ray_t *dr, float *dm, int32_t *dout, uint64_t m, cudaStream_t s) { return cb200_launch_visible(a, dr, dm, dout, m, s); });
}

} // extern "C"
