// traverse.cu -- closest-hit and any-hit traversal of the 4-wide BVH (sm_100a).
//
// Semantics follow accel_intersect / accel_visible (src/accel.d/qbvhmp.c:1262-1490):
//   * four slab tests per node on time-interpolated boxes, clipped to [0, hit.dist]
//     (tmin starts at 0, not ray.min_dist), SSE min/max select semantics spelled out so
//     NaN/inf cases agree (second operand wins), qbvhmp.c:1188-1246;
//   * children visited in the reference's topological order from axis0/axis00/axis01 and the ray's
//     sign bits, the others pushed far->near together with their entry distance, qbvhmp.c:1313-1354;
//   * popped entries whose entry distance exceeds the current hit distance are skipped, :1357-1386;
//   * leaf primitives tested in primid[] order, "dist <= hit.dist" so the last tested wins ties.
//
// Mapping: one ray per thread, while-while loop, persistent warps pulling 32-ray batches from a
// global ticket counter; per-thread traversal stack in local memory (L1-resident).
#include "prims.cuh"

#define TRACE_BLOCK 128

__device__ __forceinline__ float sse_min(float a, float b) { return a < b ? a : b; }   // _mm_min_ps
__device__ __forceinline__ float sse_max(float a, float b) { return a > b ? a : b; }   // _mm_max_ps

template<bool MB>
__device__ __forceinline__ void node_slabs(const void *nodes, uint64_t idx, const RayD &r, float ix, float iy, float iz,
                                           float t0, float t1, float tmax_init, float tmin[4], float tmax[4],
                                           uint64_t child[4], int &axis0, int &axis00, int &axis01)
{
  tmin[0] = tmin[1] = tmin[2] = tmin[3] = 0.0f;
  tmax[0] = tmax[1] = tmax[2] = tmax[3] = tmax_init;
  const float pos[3] = {r.px, r.py, r.pz};
  const float inv[3] = {ix, iy, iz};
  if(MB)
  {
    const Node256 *n = reinterpret_cast<const Node256 *>(nodes) + idx;
    const float4 *a0 = reinterpret_cast<const float4 *>(n->aabb0);
    const float4 *a1 = reinterpret_cast<const float4 *>(n->aabb1);
#pragma unroll
    for(int k=0;k<3;k++)
    {
      const float4 m0 = __ldg(a0 + k), M0 = __ldg(a0 + k + 3);
      const float4 m1 = __ldg(a1 + k), M1 = __ldg(a1 + k + 3);
      const float mo[4] = {m0.x, m0.y, m0.z, m0.w}, Mo[4] = {M0.x, M0.y, M0.z, M0.w};
      const float mc[4] = {m1.x, m1.y, m1.z, m1.w}, Mc[4] = {M1.x, M1.y, M1.z, M1.w};
#pragma unroll
      for(int c=0;c<4;c++)
      {
        const float lo = ((mo[c]*t0 + mc[c]*t1) - pos[k]) * inv[k];
        const float hi = ((Mo[c]*t0 + Mc[c]*t1) - pos[k]) * inv[k];
        tmin[c] = sse_max(tmin[c], sse_min(lo, hi));
        tmax[c] = sse_min(tmax[c], sse_max(lo, hi));
      }
    }
    const ulonglong2 c01 = __ldg(reinterpret_cast<const ulonglong2 *>(n->child));
    const ulonglong2 c23 = __ldg(reinterpret_cast<const ulonglong2 *>(n->child) + 1);
    child[0] = c01.x; child[1] = c01.y; child[2] = c23.x; child[3] = c23.y;
    const ulonglong2 pa = __ldg(reinterpret_cast<const ulonglong2 *>(&n->parent));
    const ulonglong2 aa = __ldg(reinterpret_cast<const ulonglong2 *>(&n->axis00));
    axis0 = (int)pa.y; axis00 = (int)aa.x; axis01 = (int)aa.y;
  }
  else
  {
    const Node128 *n = reinterpret_cast<const Node128 *>(nodes) + idx;
    const float4 *a0 = reinterpret_cast<const float4 *>(n->aabb0);
#pragma unroll
    for(int k=0;k<3;k++)
    {
      const float4 m0 = __ldg(a0 + k), M0 = __ldg(a0 + k + 3);
      const float mo[4] = {m0.x, m0.y, m0.z, m0.w}, Mo[4] = {M0.x, M0.y, M0.z, M0.w};
#pragma unroll
      for(int c=0;c<4;c++)
      {
        const float lo = (mo[c] - pos[k]) * inv[k];
        const float hi = (Mo[c] - pos[k]) * inv[k];
        tmin[c] = sse_max(tmin[c], sse_min(lo, hi));
        tmax[c] = sse_min(tmax[c], sse_max(lo, hi));
      }
    }
    const ulonglong2 c01 = __ldg(reinterpret_cast<const ulonglong2 *>(n->child));
    const ulonglong2 c23 = __ldg(reinterpret_cast<const ulonglong2 *>(n->child) + 1);
    const uint32_t ax = (uint32_t)(c01.x >> CB_AXIS_SHIFT) & 63u;
    child[0] = c01.x & CB_CHILD_MASK; child[1] = c01.y; child[2] = c23.x; child[3] = c23.y;
    axis0 = ax & 3; axis00 = (ax >> 2) & 3; axis01 = (ax >> 4) & 3;
  }
}

__device__ __forceinline__ void load_ray(const cb_ray_t *rays, uint64_t i, RayD &r)
{
  const float2 *p = reinterpret_cast<const float2 *>(rays + i);   // 40-byte records are 8-byte aligned
  const float2 a = __ldg(p), b = __ldg(p+1), c = __ldg(p+2), d = __ldg(p+3), e = __ldg(p+4);
  r.px = a.x; r.py = a.y; r.pz = b.x; r.dx = b.y; r.dy = c.x; r.dz = c.y;
  r.time = d.x; r.min_dist = d.y;
  r.ign_lo = __float_as_uint(e.x); r.ign_hi = __float_as_uint(e.y);
}

// one closest-hit traversal.  CNT adds the reference's ACCEL_DEBUG counters (qbvhmp.c:83-90).
template<bool MB, bool CNT, int STACK>
__device__ __forceinline__ void trace_closest(const DevAccel &A, const RayD &r, HitD &h, unsigned long long cnt[4])
{
  const uint32_t nearx = __float_as_uint(r.dx) >> 31, neary = __float_as_uint(r.dy) >> 31, nearz = __float_as_uint(r.dz) >> 31;
  const uint32_t nearbits = nearx | (neary << 1) | (nearz << 2);
  const float ix = 1.0f/r.dx, iy = 1.0f/r.dy, iz = 1.0f/r.dz;
  const float t1 = r.time, t0 = 1.0f - r.time;
  uint64_t stack[STACK];
  float stack_dist[STACK];
  int sp = 0;
  uint64_t node = 0;
  if(CNT) cnt[0]++;
  while(true)
  {
    float tmin[4], tmax[4];
    uint64_t child[4];
    int axis0, axis00, axis01;
    node_slabs<MB>(A.nodes, node, r, ix, iy, iz, t0, t1, h.dist, tmin, tmax, child, axis0, axis00, axis01);
    bool hitc[4];
    bool any = false;
#pragma unroll
    for(int c=0;c<4;c++) { hitc[c] = tmin[c] <= tmax[c]; any |= hitc[c]; }
    uint64_t current = 0;
    bool have = false;
    if(any)
    {
      if(CNT) { cnt[1]++; for(int c=0;c<4;c++) cnt[2] += hitc[c] ? 1 : 0; }
      // empty leaves (count 0) can only be popped and dropped again: never visit them
#pragma unroll
      for(int c=0;c<4;c++) if(child[c] == CB_LEAF_BIT) hitc[c] = false;
      const uint32_t n0 = (nearbits >> axis0) & 1u;
      const int axis1n = n0 ? axis01 : axis00;
      const int axis1f = n0 ? axis00 : axis01;
      const uint32_t n1n = (nearbits >> axis1n) & 1u, n1f = (nearbits >> axis1f) & 1u;
      const uint32_t f0 = n0 ^ 1u;
      const uint32_t n11 = (f0 << 1) | (n1f ^ 1u);
      const uint32_t n10 = (f0 << 1) | n1f;
      const uint32_t n01 = (n0 << 1) | (n1n ^ 1u);
      const uint32_t n00 = (n0 << 1) | n1n;
      // select by dynamic index without local-memory arrays
#define SEL4(arr, i) ((i) == 0 ? arr[0] : (i) == 1 ? arr[1] : (i) == 2 ? arr[2] : arr[3])
      const bool h00 = SEL4(hitc, n00), h01 = SEL4(hitc, n01), h10 = SEL4(hitc, n10), h11 = SEL4(hitc, n11);
      // far -> near push order: n11, n10, n01; the nearest hit child becomes current
      const int first = h00 ? 0 : h01 ? 1 : h10 ? 2 : h11 ? 3 : 4;
      if(first < 4)
      {
        have = true;
        if(h11 && first < 3) { stack_dist[sp] = SEL4(tmin, n11); stack[sp++] = SEL4(child, n11); }
        if(h10 && first < 2) { stack_dist[sp] = SEL4(tmin, n10); stack[sp++] = SEL4(child, n10); }
        if(h01 && first < 1) { stack_dist[sp] = SEL4(tmin, n01); stack[sp++] = SEL4(child, n01); }
        const uint32_t nf = first == 0 ? n00 : first == 1 ? n01 : first == 2 ? n10 : n11;
        current = SEL4(child, nf);
      }
#undef SEL4
    }
    if(!have)
    {
      do
      {
        if(sp == 0) return;
        --sp;
        current = stack[sp];
      }
      while(stack_dist[sp] > h.dist);
    }
    while(current & CB_LEAF_BIT)
    {
      const uint64_t begin = (current ^ CB_LEAF_BIT) >> 5;
      const uint32_t num = (uint32_t)current & 31u;
      const float4 *rec = A.recs + begin*(uint64_t)(A.rec_units*4);
      for(uint32_t k=0;k<num;k++)
      {
        if(CNT) cnt[3]++;
        prim_intersect(rec, A.rec_units, r, h);
        rec += A.rec_units*4;
      }
      do
      {
        if(sp == 0) return;
        --sp;
        current = stack[sp];
      }
      while(stack_dist[sp] > h.dist);
    }
    node = current;
  }
}

// any-hit sweep; returns 1 when nothing blocks the ray up to max_dist (accel_visible semantics).
template<bool MB, int STACK>
__device__ __forceinline__ int trace_visible(const DevAccel &A, const RayD &r, float max_dist)
{
  const float ix = 1.0f/r.dx, iy = 1.0f/r.dy, iz = 1.0f/r.dz;
  const float t1 = r.time, t0 = 1.0f - r.time;
  uint64_t stack[STACK];
  int sp = 0;
  uint64_t node = 0;
  while(true)
  {
    float tmin[4], tmax[4];
    uint64_t child[4];
    int axis0, axis00, axis01;
    node_slabs<MB>(A.nodes, node, r, ix, iy, iz, t0, t1, max_dist, tmin, tmax, child, axis0, axis00, axis01);
#pragma unroll
    for(int c=0;c<4;c++) if(tmin[c] <= tmax[c] && child[c] != CB_LEAF_BIT) stack[sp++] = child[c];
    uint64_t current;
    while(true)
    {
      if(sp == 0) return 1;
      current = stack[--sp];
      if(!(current & CB_LEAF_BIT)) break;
      const uint64_t begin = (current ^ CB_LEAF_BIT) >> 5;
      const uint32_t num = (uint32_t)current & 31u;
      const float4 *rec = A.recs + begin*(uint64_t)(A.rec_units*4);
      for(uint32_t k=0;k<num;k++, rec += A.rec_units*4)
        if(prim_visible(rec, A.rec_units, r, max_dist)) return 0;
    }
    node = current;
  }
}

// ---------------------------------------------------------------------------------------------
// kernels: persistent warps, each pulls 32 consecutive rays per ticket
// ---------------------------------------------------------------------------------------------
template<bool MB, bool CNT, int STACK>
__global__ void __launch_bounds__(TRACE_BLOCK)
k_intersect(DevAccel A, const cb_ray_t *__restrict__ rays, const float *__restrict__ max_dist,
            cb_hitrec_t *__restrict__ out, uint64_t n, unsigned long long *ticket, unsigned long long *counters)
{
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long cnt[4] = {0, 0, 0, 0};
  while(true)
  {
    unsigned long long base = 0;
    if(lane == 0) base = atomicAdd(ticket, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if(base >= n) break;
    const uint64_t i = base + lane;
    if(i < n)
    {
      RayD r;
      load_ray(rays, i, r);
      HitD h;
      h.dist = max_dist ? __ldg(max_dist + i) : FLT_MAX;
      h.u = 0.0f; h.v = 0.0f;
      h.prim_lo = 0xffffffffu; h.prim_hi = 0xffffffffu;
      trace_closest<MB, CNT, STACK>(A, r, h, cnt);
      // 24-byte record: three 8-byte stores
      uint2 *o = reinterpret_cast<uint2 *>(out + i);
      o[0] = make_uint2(h.prim_lo, h.prim_hi);
      o[1] = make_uint2(__float_as_uint(h.u), __float_as_uint(h.v));
      o[2] = make_uint2(__float_as_uint(h.dist), 0u);
    }
  }
  if(CNT)
    for(int k=0;k<4;k++) if(cnt[k]) atomicAdd(counters + k, cnt[k]);
}

template<bool MB, int STACK>
__global__ void __launch_bounds__(TRACE_BLOCK)
k_visible(DevAccel A, const cb_ray_t *__restrict__ rays, const float *__restrict__ max_dist,
          int32_t *__restrict__ out, uint64_t n, unsigned long long *ticket)
{
  const uint32_t lane = threadIdx.x & 31u;
  while(true)
  {
    unsigned long long base = 0;
    if(lane == 0) base = atomicAdd(ticket, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if(base >= n) break;
    const uint64_t i = base + lane;
    if(i < n)
    {
      RayD r;
      load_ray(rays, i, r);
      out[i] = trace_visible<MB, STACK>(A, r, __ldg(max_dist + i));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
#include <atomic>
#include <mutex>
static unsigned long long *g_tickets = nullptr;   // ring of ticket counters, zeroed asynchronously
static std::atomic<unsigned> g_ticket_next{0};
static std::mutex g_ticket_mutex;
#define NUM_TICKETS 256

static int get_ticket(cudaStream_t stream, unsigned long long **t)
{
  {
    std::lock_guard<std::mutex> lock(g_ticket_mutex);
    if(!g_tickets) CB_CUDA(cudaMalloc(&g_tickets, sizeof(unsigned long long)*NUM_TICKETS));
  }
  unsigned long long *p = g_tickets + (g_ticket_next++ % NUM_TICKETS);
  CB_CUDA(cudaMemsetAsync(p, 0, sizeof(unsigned long long), stream));
  *t = p;
  return 0;
}

static int grid_for(uint64_t n, const void *kernel)
{
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRACE_BLOCK, 0);
  if(per_sm < 1) per_sm = 1;
  const uint64_t want = (n + TRACE_BLOCK - 1)/TRACE_BLOCK;
  const uint64_t full = (uint64_t)cb200_sm_count_cached()*per_sm;
  return (int)(want < full ? (want ? want : 1) : full);
}

// the stack must hold 3 entries per tree level (qbvhmp.c:1277); pick the smallest variant that fits
#define STACK_SMALL 48
#define STACK_MID   96
#define STACK_BIG   304

template<bool MB, bool CNT>
static int launch_intersect_t(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                              uint64_t n, cudaStream_t stream, unsigned long long *d_counters)
{
  unsigned long long *ticket;
  if(get_ticket(stream, &ticket)) return CB200_ERR_CUDA;
  const int need = 3*a->depth + 1;
  if(need <= STACK_SMALL)
  {
    auto k = k_intersect<MB, CNT, STACK_SMALL>;
    k<<<grid_for(n, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays, d_max_dist, d_out, n, ticket, d_counters);
  }
  else if(need <= STACK_MID)
  {
    auto k = k_intersect<MB, CNT, STACK_MID>;
    k<<<grid_for(n, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays, d_max_dist, d_out, n, ticket, d_counters);
  }
  else if(need <= STACK_BIG)
  {
    auto k = k_intersect<MB, CNT, STACK_BIG>;
    k<<<grid_for(n, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays, d_max_dist, d_out, n, ticket, d_counters);
  }
  else { cb200_set_error("tree deeper than the reference's MAX_TREE_DEPTH"); return CB200_ERR_UNSUPPORTED; }
  cb200_count_launch();
  CB_CUDA(cudaGetLastError());
  return 0;
}

int cb200_launch_intersect(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                           uint64_t n, cudaStream_t stream, unsigned long long *d_counters)
{
  if(n == 0) return 0;
  if(a->dev.mb) return d_counters ? launch_intersect_t<true,  true>(a, d_rays, d_max_dist, d_out, n, stream, d_counters)
                                  : launch_intersect_t<true,  false>(a, d_rays, d_max_dist, d_out, n, stream, nullptr);
  else          return d_counters ? launch_intersect_t<false, true>(a, d_rays, d_max_dist, d_out, n, stream, d_counters)
                                  : launch_intersect_t<false, false>(a, d_rays, d_max_dist, d_out, n, stream, nullptr);
}

template<bool MB>
static int launch_visible_t(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, int32_t *d_out,
                            uint64_t n, cudaStream_t stream)
{
  unsigned long long *ticket;
  if(get_ticket(stream, &ticket)) return CB200_ERR_CUDA;
  const int need = 4*a->depth + 4;   // the any-hit sweep pushes up to 4 children per level
  if(need <= STACK_MID)
  {
    auto k = k_visible<MB, STACK_MID>;
    k<<<grid_for(n, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays, d_max_dist, d_out, n, ticket);
  }
  else if(need <= 4*101 + 4)
  {
    auto k = k_visible<MB, 4*101 + 4>;
    k<<<grid_for(n, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays, d_max_dist, d_out, n, ticket);
  }
  else { cb200_set_error("tree deeper than the reference's MAX_TREE_DEPTH"); return CB200_ERR_UNSUPPORTED; }
  cb200_count_launch();
  CB_CUDA(cudaGetLastError());
  return 0;
}

int cb200_launch_visible(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, int32_t *d_out,
                         uint64_t n, cudaStream_t stream)
{
  if(n == 0) return 0;
  if(!d_max_dist) { cb200_set_error("visible: max_dist is required"); return CB200_ERR_ARG; }
  return a->dev.mb ? launch_visible_t<true>(a, d_rays, d_max_dist, d_out, n, stream)
                   : launch_visible_t<false>(a, d_rays, d_max_dist, d_out, n, stream);
}
