// comm.cu -- the framebuffer exchange of a sample-split multi-GPU job, for hosts without torch: one rank (process or
// thread) per GPU renders its own progressions, and after every progression ONE ncclReduce sums the rank's W*H*3 fp32 accumulation
// buffer into rank 0 over NVLink.  Nothing else is exchanged: paths are independent given their index (render_sample_path(i),
// src/render.d/gi.c:81-88) and the framebuffer is additive (view_splat, src/view.c:455-495).
//
// Everything that follows a progression -- the reduce, the root's accumulate, clearing the buffer for its next use and the copy of the
// root's running sum to the host framebuffer -- is queued on one side stream in that order; the render stream only waits for the event
// that marks its next buffer as cleared, so a progression's epilogue overlaps the next progression's kernels (buffers are double
// buffered).  This is corona-13_b200/progressive.py's FramebufferReducer in the library, for the plain-C host (host/main_b200.c --gpus N).
//
// NCCL is loaded with dlopen at the first use, not linked: processes that never run multi-GPU through this file (bench.py's ranks
// exchange through torch.distributed, which carries its own NCCL) never load a second copy.
#include "internal.h"
#include "corona_b200_render.h"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>

namespace {

struct NcclApi
{
  void *lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclReduce) Reduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi g_nccl;

int nccl_load()
{
  if(g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for(const char *n : names) if((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if(!lib) { cb200_set_error(std::string("multi-GPU: cannot load NCCL: ") + dlerror()); return CB200_ERR_UNSUPPORTED; }
#define SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); \
  if(!g_nccl.field) { cb200_set_error("multi-GPU: NCCL lacks " name); dlclose(lib); return CB200_ERR_UNSUPPORTED; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(Reduce, "ncclReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.lib = lib;
  return 0;
}

int nccl_fail(ncclResult_t e, const char *what)
{
  cb200_set_error(std::string("NCCL: ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?"));
  return CB200_ERR_CUDA;
}
#define CB_NCCL(call) do { ncclResult_t e__ = (call); if(e__ != ncclSuccess) return nccl_fail(e__, #call); } while(0)

__global__ void k_accumulate(float *__restrict__ acc, const float *__restrict__ src, uint64_t n)
{
  const uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x*blockDim.x;
  float4 *a4 = reinterpret_cast<float4 *>(acc);
  const float4 *s4 = reinterpret_cast<const float4 *>(src);
  for(uint64_t k=i; k<n/4; k+=stride)
  {
    float4 a = a4[k];
    const float4 s = s4[k];
    a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
    a4[k] = a;
  }
  if(i < (n & 3)) acc[(n & ~3ull) + i] += src[(n & ~3ull) + i];
}

} // namespace

struct cb200_reducer
{
  cb200_render *render;
  int rank, world;
  ncclComm_t comm;
  uint64_t count;                 // floats per framebuffer
  float *buf[2];                  // this rank's accumulation buffers (the render object is pointed at one of them per progression)
  float *accum;                   // rank 0: running sum over all ranks and progressions
  cudaStream_t side;
  cudaEvent_t rendered, cleared[2], summed;
  bool in_use[2];
};

extern "C" {

int cb200_comm_unique_id(void *id128)
{
  if(!id128) { cb200_set_error("comm_unique_id: null buffer"); return CB200_ERR_ARG; }
  if(int rc = nccl_load()) return rc;
  static_assert(sizeof(ncclUniqueId) == CB200_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  CB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

cb200_reducer_t *cb200_reducer_create(cb200_render_t *r, const void *id128, int rank, int world, uint32_t width, uint32_t height)
{
  if(!r || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128) || !width || !height)
  { cb200_set_error("reducer_create: bad arguments"); return nullptr; }
  cb200_reducer *q = new cb200_reducer();
  memset(q, 0, sizeof(*q));
  q->render = r; q->rank = rank; q->world = world;
  q->count = (uint64_t)width*height*3;
  auto fail = [&](const char *what) -> cb200_reducer_t * { if(what) cb200_set_error(what); cb200_reducer_destroy(q); return nullptr; };
  if(world > 1)
  {
    if(nccl_load()) return fail(nullptr);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t e = g_nccl.CommInitRank(&q->comm, world, id, rank);
    if(e != ncclSuccess) { nccl_fail(e, "ncclCommInitRank"); return fail(nullptr); }
  }
  const int nbuf = world > 1 ? 2 : 1;
  for(int k=0;k<nbuf;k++)
  {
    if(cudaMalloc(&q->buf[k], q->count*sizeof(float)) != cudaSuccess) return fail("reducer_create: out of device memory");
    cudaMemset(q->buf[k], 0, q->count*sizeof(float));
  }
  if(world > 1 && rank == 0)
  {
    if(cudaMalloc(&q->accum, q->count*sizeof(float)) != cudaSuccess) return fail("reducer_create: out of device memory");
    cudaMemset(q->accum, 0, q->count*sizeof(float));
  }
  if(cudaStreamCreateWithFlags(&q->side, cudaStreamNonBlocking) != cudaSuccess) return fail("reducer_create: no stream");
  cudaEventCreateWithFlags(&q->rendered, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&q->summed, cudaEventDisableTiming);
  for(int k=0;k<2;k++) cudaEventCreateWithFlags(&q->cleared[k], cudaEventDisableTiming);
  if(cudaDeviceSynchronize() != cudaSuccess) return fail("reducer_create: device error");
  return q;
}

void cb200_reducer_destroy(cb200_reducer_t *q)
{
  if(!q) return;
  if(q->side) { cudaStreamSynchronize(q->side); cudaStreamDestroy(q->side); }
  if(q->render) cb200_render_set_framebuffer(q->render, nullptr);
  if(q->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(q->comm);
  for(int k=0;k<2;k++) { cudaFree(q->buf[k]); if(q->cleared[k]) cudaEventDestroy(q->cleared[k]); }
  cudaFree(q->accum);
  if(q->rendered) cudaEventDestroy(q->rendered);
  if(q->summed) cudaEventDestroy(q->summed);
  delete q;
}

// point the render object at the buffer of local progression `step`, ordered behind the epilogue that last used that buffer
int cb200_reducer_begin(cb200_reducer_t *q, uint64_t step, void *stream)
{
  if(!q) { cb200_set_error("reducer_begin: null reducer"); return CB200_ERR_ARG; }
  const int k = q->world > 1 ? (int)(step & 1) : 0;
  if(q->in_use[k]) { CB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, q->cleared[k], 0)); q->in_use[k] = false; }
  return cb200_render_set_framebuffer(q->render, q->buf[k]);
}

// the progression's work is queued on `stream`: reduce its buffer into rank 0, accumulate there, clear it, and (rank 0, fb_host != NULL)
// copy the running sum to the host -- all on the side stream.  fb_host should be pinned and must stay valid until cb200_reducer_finish.
int cb200_reducer_end(cb200_reducer_t *q, uint64_t step, float *fb_host, void *stream)
{
  if(!q) { cb200_set_error("reducer_end: null reducer"); return CB200_ERR_ARG; }
  const int k = q->world > 1 ? (int)(step & 1) : 0;
  CB_CUDA(cudaEventRecord(q->rendered, (cudaStream_t)stream));
  CB_CUDA(cudaStreamWaitEvent(q->side, q->rendered, 0));
  if(q->world == 1)
  { // a single rank accumulates in place: only the optional host mirror
    if(fb_host) CB_CUDA(cudaMemcpyAsync(fb_host, q->buf[0], q->count*sizeof(float), cudaMemcpyDeviceToHost, q->side));
    return 0;
  }
  CB_NCCL(g_nccl.Reduce(q->buf[k], q->buf[k], q->count, ncclFloat, ncclSum, 0, q->comm, q->side));
  cb200_count_launch();
  if(q->rank == 0)
  {
    k_accumulate<<<cb200_sm_count_cached()*4, 256, 0, q->side>>>(q->accum, q->buf[k], q->count);
    cb200_count_launch();
    CB_CUDA(cudaGetLastError());
    if(fb_host) CB_CUDA(cudaMemcpyAsync(fb_host, q->accum, q->count*sizeof(float), cudaMemcpyDeviceToHost, q->side));
  }
  CB_CUDA(cudaMemsetAsync(q->buf[k], 0, q->count*sizeof(float), q->side));
  CB_CUDA(cudaEventRecord(q->cleared[k], q->side));
  q->in_use[k] = true;
  return 0;
}

// wait for every queued epilogue; rank 0 receives the sum over all ranks in fb_host (may be NULL), the other ranks nothing
int cb200_reducer_finish(cb200_reducer_t *q, float *fb_host)
{
  if(!q) { cb200_set_error("reducer_finish: null reducer"); return CB200_ERR_ARG; }
  CB_CUDA(cudaStreamSynchronize(q->side));
  q->in_use[0] = q->in_use[1] = false;
  if(fb_host && q->rank == 0)
    CB_CUDA(cudaMemcpy(fb_host, q->world > 1 ? q->accum : q->buf[0], q->count*sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int cb200_reducer_clear(cb200_reducer_t *q)
{
  if(!q) { cb200_set_error("reducer_clear: null reducer"); return CB200_ERR_ARG; }
  CB_CUDA(cudaStreamSynchronize(q->side));
  for(int k=0;k<2;k++) if(q->buf[k]) CB_CUDA(cudaMemset(q->buf[k], 0, q->count*sizeof(float)));
  if(q->accum) CB_CUDA(cudaMemset(q->accum, 0, q->count*sizeof(float)));
  q->in_use[0] = q->in_use[1] = false;
  return 0;
}

} // extern "C"
