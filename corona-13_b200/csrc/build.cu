// build.cu -- GPU construction of the 4-wide BVH (replaces accel_build, src/accel.d/qbvhmp.c:1034-1186)
//
//   1. k_prim_bounds   per-primitive shutter-open / shutter-close boxes (prims.c:20-60, triangle.h:7-33,
//                      sphere.h:16-30, line.h:24-55) + scene box (shutter-open boxes only, like
//                      compute_aabb, qbvhmp.c:1034-1065)
//   2. k_morton        63-bit Morton code of the open-box centroid (21 bits per axis)
//   3. cub radix sort  (code, prim index)
//   4. k_radix_tree    Karras' binary radix tree over the sorted codes (duplicates split by index)
//   5. k_refit         bottom-up boxes (open and close) and subtree primitive counts
//   6. k_collapse      level-by-level collapse of every other binary level into the reference's
//                      4-wide node: children [0,1] = the two halves of the lower side along axis0,
//                      [2,3] = the upper side, axis00/axis01 = split axes of those halves; subtrees
//                      with <= 6 primitives become leaves (1<<63)|(begin<<5)|count (qbvhmp.c:44,989)
//   7. k_records       gather vertices into leaf-ordered 64-byte primitive records
//
// Like the reference build this permutes the global primid list (leaves index the permuted list).
#include "internal.h"
#include <cub/cub.cuh>
#include <float.h>
#include <vector>

#define BUILD_BLOCK 256
#define LEAF_MAX_REF 6   // NUM_TRIS_PER_LEAF, qbvhmp.c:44
#define LEAF_MAX_DEFAULT 3

struct Box { float lo[3], hi[3]; };

__device__ __forceinline__ uint32_t pid_shape(uint64_t p) { return (uint32_t)(p >> 3) & 0x1fffffffu; }
__device__ __forceinline__ uint32_t pid_vi(uint64_t p)    { return (uint32_t)(p >> 32) & 0x0fffffffu; }
__device__ __forceinline__ uint32_t pid_mb(uint64_t p)    { return (uint32_t)(p >> 60) & 1u; }
__device__ __forceinline__ uint32_t pid_vcnt(uint64_t p)  { return (uint32_t)(p >> 61) & 7u; }

struct SceneView
{
  const cb_vtx_t *vtx;
  const cb_vtxidx_t *vtxidx;
  const ShapeDev *shapes;
};

__device__ __forceinline__ const cb_vtx_t *vertex_ptr(const SceneView &S, uint64_t pid, int v, int close)
{
  const ShapeDev sh = S.shapes[pid_shape(pid)];
  const uint32_t mb = pid_mb(pid);
  const uint32_t vi = S.vtxidx[sh.vtxidx_off + pid_vi(pid) + v].v;
  return S.vtx + sh.vtx_off + (uint64_t)(mb + 1)*vi + (close ? mb : 0);
}

// float atomics through the usual order-preserving integer trick
__device__ __forceinline__ void atomic_min_f(float *addr, float v)
{
  if(v >= 0.0f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
  else          atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float *addr, float v)
{
  if(v >= 0.0f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else          atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__device__ void line_bounds(const float *v0, const float *v1, float r0, float r1, float *lo, float *hi)
{
  float d[3] = {v1[0]-v0[0], v1[1]-v0[1], v1[2]-v0[2]};
  const float il = 1.0f/sqrtf((d[0]*d[0] + d[1]*d[1]) + d[2]*d[2]);
  d[0] *= il; d[1] *= il; d[2] *= il;
  for(int k=0;k<3;k++)
  {
    // half extent of a disc of radius r perpendicular to d along axis k is r*sqrt(1-d_k^2); the reference
    // gets the same number through atan2f/sinf/cosf (line.h:33-36).  grow by 2 ulp-ish to stay conservative.
    const float m = sqrtf(fmaxf(0.0f, 1.0f - d[k]*d[k])) * 1.000001f + 1e-7f;
    lo[k] = fminf(v0[k] - r0*m, v1[k] - r1*m);
    hi[k] = fmaxf(v0[k] + r0*m, v1[k] + r1*m);
  }
}

__device__ void prim_box(const SceneView &S, uint64_t pid, int close, float *lo, float *hi)
{
  const uint32_t vcnt = pid_vcnt(pid);
  const cb_vtx_t *p0 = vertex_ptr(S, pid, 0, close);
  if(vcnt == CB_PRIM_SPHERE)
  {
    const float radius = __uint_as_float(vertex_ptr(S, pid, 0, 0)->n);
    for(int k=0;k<3;k++) { lo[k] = p0->v[k] - radius; hi[k] = p0->v[k] + radius; }
  }
  else if(vcnt == CB_PRIM_LINE)
  {
    const cb_vtx_t *p1 = vertex_ptr(S, pid, 1, close);
    const float r0 = __uint_as_float(vertex_ptr(S, pid, 0, 0)->n);
    const float r1 = __uint_as_float(vertex_ptr(S, pid, 1, 0)->n);
    line_bounds(p0->v, p1->v, r0, r1, lo, hi);
  }
  else
  {
    for(int k=0;k<3;k++) lo[k] = hi[k] = p0->v[k];
    for(uint32_t v=1;v<vcnt && v<4;v++)
    {
      const cb_vtx_t *p = vertex_ptr(S, pid, v, close);
      for(int k=0;k<3;k++) { lo[k] = fminf(p->v[k], lo[k]); hi[k] = fmaxf(p->v[k], hi[k]); }
    }
  }
}

__global__ void k_prim_bounds(SceneView S, const uint64_t *__restrict__ primid, uint32_t n, int any_mb,
                              Box *__restrict__ box0, Box *__restrict__ box1, float *scene_box)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if(i < n)
  {
    const uint64_t pid = primid[i];
    prim_box(S, pid, 0, lo, hi);
    Box b; for(int k=0;k<3;k++) { b.lo[k] = lo[k]; b.hi[k] = hi[k]; }
    box0[i] = b;
    if(any_mb)
    {
      float l1[3], h1[3];
      prim_box(S, pid, 1, l1, h1);
      Box c; for(int k=0;k<3;k++) { c.lo[k] = l1[k]; c.hi[k] = h1[k]; }
      box1[i] = c;
    }
  }
  // block reduce the shutter-open scene box
  typedef cub::BlockReduce<float, BUILD_BLOCK> BR;
  __shared__ typename BR::TempStorage tmp;
  for(int k=0;k<3;k++)
  {
    const float m = BR(tmp).Reduce(lo[k], cub::Min()); __syncthreads();
    const float M = BR(tmp).Reduce(hi[k], cub::Max()); __syncthreads();
    if(threadIdx.x == 0) { atomic_min_f(scene_box + k, m); atomic_max_f(scene_box + 3 + k, M); }
  }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x)
{ // spread the low 21 bits so that two zero bits separate consecutive bits
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8)  & 0x100f00f00f00f00full;
  x = (x | x << 4)  & 0x10c30c30c30c30c3ull;
  x = (x | x << 2)  & 0x1249249249249249ull;
  return x;
}

__global__ void k_morton(const Box *__restrict__ box0, uint32_t n, const float *__restrict__ scene_box,
                         uint64_t *__restrict__ codes, uint32_t *__restrict__ index)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  const Box b = box0[i];
  uint64_t q[3];
  for(int k=0;k<3;k++)
  {
    const float ext = scene_box[3+k] - scene_box[k];
    const float c = 0.5f*(b.lo[k] + b.hi[k]);
    float f = ext > 0.0f ? (c - scene_box[k])/ext : 0.0f;
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    q[k] = (uint64_t)fminf(f*2097152.0f, 2097151.0f);
  }
  codes[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);   // x is the most significant axis
  index[i] = i;
}

// binary radix tree ------------------------------------------------------------------------------
// references: bit 31 set = leaf (sorted position), else internal node index
#define BLEAF 0x80000000u
#define BNONE 0xffffffffu
struct BNode
{
  uint32_t left, right;   // child references
  uint32_t first, last;   // sorted range covered (inclusive)
  uint32_t axis;          // split axis 0..2
};

__device__ __forceinline__ int delta(const uint64_t *__restrict__ codes, int n, int i, int j)
{
  if(j < 0 || j >= n) return -1;
  const uint64_t a = codes[i], b = codes[j];
  if(a == b) return 64 + __clz(i ^ j);
  return __clzll(a ^ b);
}

__global__ void k_radix_tree(const uint64_t *__restrict__ codes, int n, BNode *__restrict__ nodes,
                             uint32_t *__restrict__ node_parent, uint32_t *__restrict__ leaf_parent)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n-1) return;
  const int d = (delta(codes, n, i, i+1) - delta(codes, n, i, i-1)) >= 0 ? 1 : -1;
  const int dmin = delta(codes, n, i, i-d);
  int lmax = 2;
  while(delta(codes, n, i, i + lmax*d) > dmin) lmax <<= 1;
  int l = 0;
  for(int t=lmax>>1;t>=1;t>>=1)
    if(delta(codes, n, i, i + (l+t)*d) > dmin) l += t;
  const int j = i + l*d;
  const int dnode = delta(codes, n, i, j);
  int s = 0;
  for(int div=2;;div<<=1)
  {
    const int t = (l + div - 1)/div;
    if(delta(codes, n, i, i + (s+t)*d) > dnode) s += t;
    if(t <= 1) break;
  }
  const int gamma = i + s*d + (d < 0 ? -1 : 0);
  const int lo = min(i, j), hi = max(i, j);
  BNode nd;
  nd.first = lo; nd.last = hi;
  nd.left  = (lo == gamma)   ? (BLEAF | (uint32_t)gamma)     : (uint32_t)gamma;
  nd.right = (hi == gamma+1) ? (BLEAF | (uint32_t)(gamma+1)) : (uint32_t)(gamma+1);
  // split axis from the first differing Morton bit (x lives at bits 3k+2, y at 3k+1, z at 3k)
  nd.axis = 0;
  if(dnode < 64) { const int bit = 63 - dnode; nd.axis = 2 - (bit % 3); }
  nodes[i] = nd;
  if(nd.left & BLEAF)  leaf_parent[nd.left & ~BLEAF] = i;  else node_parent[nd.left] = i;
  if(nd.right & BLEAF) leaf_parent[nd.right & ~BLEAF] = i; else node_parent[nd.right] = i;
}

__device__ __forceinline__ void box_load(const float *b, uint32_t i, float *o)
{
  for(int k=0;k<6;k++) o[k] = __ldcg(b + (uint64_t)i*6 + k);   // L2: written by other blocks in this launch
}

// bottom-up boxes: every leaf walks to the root, the second arrival at a node does the union.
// pbox*: per-primitive boxes in load order (Box = 6 floats), index: sorted position -> load order.
__global__ void k_refit(const BNode *__restrict__ nodes, const uint32_t *__restrict__ node_parent,
                        const uint32_t *__restrict__ leaf_parent, const uint32_t *__restrict__ index,
                        const float *__restrict__ pbox0, const float *__restrict__ pbox1, int n, int any_mb,
                        float *nbox0, float *nbox1, uint32_t *flags)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  uint32_t cur = leaf_parent[i];
  while(cur != BNONE)
  {
    if(atomicAdd(flags + cur, 1u) == 0) return;   // first child to arrive leaves
    __threadfence();
    const BNode nd = nodes[cur];
    float a[6], b[6];
    if(nd.left  & BLEAF) box_load(pbox0, index[nd.left  & ~BLEAF], a); else box_load(nbox0, nd.left,  a);
    if(nd.right & BLEAF) box_load(pbox0, index[nd.right & ~BLEAF], b); else box_load(nbox0, nd.right, b);
    for(int k=0;k<3;k++) { nbox0[(uint64_t)cur*6+k] = fminf(a[k], b[k]); nbox0[(uint64_t)cur*6+3+k] = fmaxf(a[3+k], b[3+k]); }
    if(any_mb)
    {
      if(nd.left  & BLEAF) box_load(pbox1, index[nd.left  & ~BLEAF], a); else box_load(nbox1, nd.left,  a);
      if(nd.right & BLEAF) box_load(pbox1, index[nd.right & ~BLEAF], b); else box_load(nbox1, nd.right, b);
      for(int k=0;k<3;k++) { nbox1[(uint64_t)cur*6+k] = fminf(a[k], b[k]); nbox1[(uint64_t)cur*6+3+k] = fmaxf(a[3+k], b[3+k]); }
    }
    __threadfence();
    cur = node_parent[cur];
  }
}

// collapse -----------------------------------------------------------------------------------------
struct CollapseItem { uint32_t bref; uint32_t wide; };

struct CollapseArgs
{
  const BNode *nodes;
  const uint32_t *index;
  const float *pbox0, *pbox1, *nbox0, *nbox1;
  int n, any_mb;
  Node256 *out256;   // exactly one of out256 / out128 is set
  Node128 *out128;
  uint32_t *num_wide;          // allocation counter
  uint32_t max_wide;           // capacity of the node buffer and of the queues
  uint32_t leaf_max;           // subtrees with at most this many primitives become leaves
};

__device__ __forceinline__ uint32_t bref_count(const BNode *nodes, uint32_t r)
{
  return (r & BLEAF) ? 1u : nodes[r].last - nodes[r].first + 1u;
}
__device__ __forceinline__ uint32_t bref_first(const BNode *nodes, uint32_t r)
{
  return (r & BLEAF) ? (r & ~BLEAF) : nodes[r].first;
}
__device__ __forceinline__ void bref_box(const CollapseArgs &A, uint32_t r, int close, float *o)
{
  const float *pb = close ? A.pbox1 : A.pbox0;
  const float *nb = close ? A.nbox1 : A.nbox0;
  if(r & BLEAF)
  { // a single primitive
    const uint32_t p = A.index[r & ~BLEAF];
    for(int k=0;k<6;k++) o[k] = pb[(uint64_t)p*6+k];
  }
  else for(int k=0;k<6;k++) o[k] = nb[(uint64_t)r*6+k];
}

__global__ void k_collapse(CollapseArgs A, const CollapseItem *__restrict__ in, uint32_t num_in,
                           CollapseItem *__restrict__ out, uint32_t *num_out, uint64_t parent_unused)
{
  const uint32_t t = blockIdx.x*blockDim.x + threadIdx.x;
  if(t >= num_in) return;
  const CollapseItem it = in[t];
  if(it.wide >= A.max_wide) return;   // overflow is reported by the host from the counter
  // slots: bref or BNONE (empty)
  uint32_t slot[4] = {BNONE, BNONE, BNONE, BNONE};
  int axis0 = 0, axis00 = 0, axis01 = 0;
  const bool root_is_leaf = (it.bref & BLEAF) || bref_count(A.nodes, it.bref) <= A.leaf_max;
  if(root_is_leaf) slot[0] = it.bref;   // only possible for the root of a tiny scene
  else
  {
    const BNode b = A.nodes[it.bref];
    axis0 = b.axis;
    const uint32_t side[2] = {b.left, b.right};
    for(int s=0;s<2;s++)
    {
      const uint32_t r = side[s];
      if(r & BLEAF) slot[2*s] = r;   // a single primitive: the pair's second slot stays empty
      else
      { // also when the whole side would fit one leaf: its two halves fill both slots at no extra node, with tighter boxes
        const BNode c = A.nodes[r];
        slot[2*s] = c.left; slot[2*s+1] = c.right;
        if(s == 0) axis00 = c.axis; else axis01 = c.axis;
      }
    }
  }
  uint64_t child[4];
  float b0[4][6], b1[4][6];
  for(int c=0;c<4;c++)
  {
    const uint32_t r = slot[c];
    if(r == BNONE)
    { // empty leaf, inverted box like the reference's empty scene root (qbvhmp.c:1085-1094)
      child[c] = CB_LEAF_BIT;
      for(int k=0;k<3;k++) { b0[c][k] = b1[c][k] = FLT_MAX; b0[c][3+k] = b1[c][3+k] = -FLT_MAX; }
      continue;
    }
    bref_box(A, r, 0, b0[c]);
    if(A.any_mb) bref_box(A, r, 1, b1[c]); else for(int k=0;k<6;k++) b1[c][k] = b0[c][k];
    const uint32_t cnt = bref_count(A.nodes, r);
    if((r & BLEAF) || cnt <= A.leaf_max)
      child[c] = CB_LEAF_BIT | ((uint64_t)bref_first(A.nodes, r) << 5) | cnt;
    else
    {
      const uint32_t w = atomicAdd(A.num_wide, 1u);
      child[c] = w;
      const uint32_t o = atomicAdd(num_out, 1u);
      if(o < A.max_wide) { out[o].bref = r; out[o].wide = w; }
    }
  }
  if(A.out256)
  {
    Node256 *n = A.out256 + it.wide;
    for(int k=0;k<6;k++) for(int c=0;c<4;c++) { n->aabb0[k][c] = b0[c][k]; n->aabb1[k][c] = b1[c][k]; }
    for(int c=0;c<4;c++) n->child[c] = child[c];
    n->parent = parent_unused;   // patched by k_parents
    n->axis0 = axis0; n->axis00 = axis00; n->axis01 = axis01;
  }
  else
  {
    Node128 *n = A.out128 + it.wide;
    for(int k=0;k<6;k++) for(int c=0;c<4;c++) n->aabb0[k][c] = b0[c][k];
    n->child[0] = child[0] | ((uint64_t)(axis0 | (axis00 << 2) | (axis01 << 4)) << CB_AXIS_SHIFT);
    for(int c=1;c<4;c++) n->child[c] = child[c];
  }
}


// 8-wide compressed tree (Node8, internal.h) over the SAME binary radix tree and primitive order -------------------------
struct Collapse8Args
{
  const BNode *nodes;
  const uint32_t *index;
  const float *pbox0, *nbox0;
  Node8 *out;
  uint32_t *num_wide;
  uint32_t max_wide;
  uint32_t leaf_max;
};

__device__ __forceinline__ void bref_box0(const Collapse8Args &A, uint32_t r, float *o)
{
  if(r & BLEAF) { const uint32_t p = A.index[r & ~BLEAF]; for(int k=0;k<6;k++) o[k] = A.pbox0[(uint64_t)p*6+k]; }
  else for(int k=0;k<6;k++) o[k] = A.nbox0[(uint64_t)r*6+k];
}

// one wide node per thread: open the binary subtree below `bref` greedily by surface area until it has 8 children (or nothing
// left to open), assign the children to slots by octant, quantise their boxes outwards on the node's grid
__global__ void k_collapse8(Collapse8Args A, const CollapseItem *__restrict__ in, uint32_t num_in,
                            CollapseItem *__restrict__ out, uint32_t *num_out)
{
  const uint32_t t = blockIdx.x*blockDim.x + threadIdx.x;
  if(t >= num_in) return;
  const CollapseItem it = in[t];
  if(it.wide >= A.max_wide) return;
  uint32_t ch[8];
  float cb[8][6];
  int nch = 0;
  if((it.bref & BLEAF) || bref_count(A.nodes, it.bref) <= A.leaf_max) ch[nch++] = it.bref;   // a scene that fits one leaf
  else
  {
    const BNode b = A.nodes[it.bref];
    ch[0] = b.left; ch[1] = b.right; nch = 2;
    bref_box0(A, ch[0], cb[0]); bref_box0(A, ch[1], cb[1]);
    while(nch < 8)
    {
      int best = -1;
      float best_area = -1.0f;
      for(int i=0;i<nch;i++)
      {
        if((ch[i] & BLEAF) || bref_count(A.nodes, ch[i]) <= A.leaf_max) continue;
        const float dx = cb[i][3] - cb[i][0], dy = cb[i][4] - cb[i][1], dz = cb[i][5] - cb[i][2];
        const float area = dx*dy + dy*dz + dz*dx;
        if(area > best_area) { best_area = area; best = i; }
      }
      if(best < 0) break;
      const BNode c = A.nodes[ch[best]];
      ch[best] = c.left; ch[nch] = c.right;
      bref_box0(A, ch[best], cb[best]); bref_box0(A, ch[nch], cb[nch]);
      nch++;
    }
  }
  if(nch == 1) bref_box0(A, ch[0], cb[0]);
  float lo[3], hi[3];
  for(int k=0;k<3;k++) { lo[k] = cb[0][k]; hi[k] = cb[0][3+k]; }
  for(int i=1;i<nch;i++) for(int k=0;k<3;k++) { lo[k] = fminf(lo[k], cb[i][k]); hi[k] = fmaxf(hi[k], cb[i][3+k]); }
  // slots by octant: child i wants the slot whose sign pattern matches (child centre - node centre); greedy by that score
  int slot_of[8], taken = 0, assigned = 0;
  for(int i=0;i<8;i++) slot_of[i] = -1;
  for(int round=0;round<nch;round++)
  {
    float best = -FLT_MAX; int bi = 0, bs = 0;
    for(int i=0;i<nch;i++)
    {
      if(assigned & (1 << i)) continue;
      float d[3];
      for(int k=0;k<3;k++) d[k] = (cb[i][k] + cb[i][3+k]) - (lo[k] + hi[k]);
      for(int sl=0;sl<8;sl++)
      {
        if(taken & (1 << sl)) continue;
        const float score = ((sl & 1) ? d[0] : -d[0]) + ((sl & 2) ? d[1] : -d[1]) + ((sl & 4) ? d[2] : -d[2]);
        if(score > best) { best = score; bi = i; bs = sl; }
      }
    }
    slot_of[bi] = bs; assigned |= 1 << bi; taken |= 1 << bs;
  }
  // per-axis grid: plane = lo + q * 2^e, 2^e >= extent / 255
  int e[3];
  for(int k=0;k<3;k++)
  {
    const float ext = hi[k] - lo[k];
    int ex = -100;
    if(ext > 0.0f && ext < FLT_MAX) { frexpf(ext/255.0f, &ex); if(ldexp(1.0, ex)*255.0 < (double)hi[k] - (double)lo[k]) ex++; }
    e[k] = ex < -100 ? -100 : ex > 100 ? 100 : ex;
  }
  uint8_t qlo[3][8], qhi[3][8];
  uint32_t child[8];
  for(int sl=0;sl<8;sl++) { child[sl] = 0u; for(int k=0;k<3;k++) { qlo[k][sl] = 255; qhi[k][sl] = 0; } }
  for(int i=0;i<nch;i++)
  {
    const int sl = slot_of[i];
    for(int k=0;k<3;k++)
    {
      const double scale = ldexp(1.0, e[k]);
      double a = floor(((double)cb[i][k] - (double)lo[k])/scale);
      double b = ceil(((double)cb[i][3+k] - (double)lo[k])/scale);
      a = a < 0.0 ? 0.0 : a > 255.0 ? 255.0 : a;
      b = b < 0.0 ? 0.0 : b > 255.0 ? 255.0 : b;
      while(a > 0.0 && (double)lo[k] + a*scale > (double)cb[i][k]) a -= 1.0;
      while(b < 255.0 && (double)lo[k] + b*scale < (double)cb[i][3+k]) b += 1.0;
      qlo[k][sl] = (uint8_t)a; qhi[k][sl] = (uint8_t)b;
    }
    const uint32_t r = ch[i], cnt = bref_count(A.nodes, r);
    if((r & BLEAF) || cnt <= A.leaf_max) child[sl] = CB8_LEAF | (bref_first(A.nodes, r) << 3) | cnt;
    else
    {
      const uint32_t w = atomicAdd(A.num_wide, 1u);
      child[sl] = w;
      const uint32_t o = atomicAdd(num_out, 1u);
      if(o < A.max_wide) { out[o].bref = r; out[o].wide = w; }
    }
  }
  Node8 nd;
  for(int k=0;k<3;k++) nd.origin[k] = lo[k];
  nd.exps = (uint32_t)(e[0] + 7 + 127) | ((uint32_t)(e[1] + 7 + 127) << 8) | ((uint32_t)(e[2] + 7 + 127) << 16);
  for(int k=0;k<3;k++)
    for(int h=0;h<2;h++)
    {
      nd.planes[4*k + h]     = (uint32_t)qlo[k][4*h] | ((uint32_t)qlo[k][4*h+1] << 8) | ((uint32_t)qlo[k][4*h+2] << 16) | ((uint32_t)qlo[k][4*h+3] << 24);
      nd.planes[4*k + 2 + h] = (uint32_t)qhi[k][4*h] | ((uint32_t)qhi[k][4*h+1] << 8) | ((uint32_t)qhi[k][4*h+2] << 16) | ((uint32_t)qhi[k][4*h+3] << 24);
    }
  for(int sl=0;sl<8;sl++) nd.child[sl] = child[sl];
  A.out[it.wide] = nd;
}

__global__ void k_parents(Node256 *nodes, uint32_t num)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= num) return;
  if(i == 0) nodes[0].parent = ~0ull;
  for(int c=0;c<4;c++)
  {
    const uint64_t ch = nodes[i].child[c];
    if(!(ch & CB_LEAF_BIT)) nodes[ch].parent = i;
  }
}

__global__ void k_permute_primid(const uint64_t *__restrict__ primid, const uint32_t *__restrict__ index, uint32_t n,
                                 uint64_t *__restrict__ out)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i < n) out[i] = primid[index[i]];
}

// leaf-ordered primitive records (layout in internal.h)
__global__ void k_records(SceneView S, const uint64_t *__restrict__ primid_perm, uint32_t n, uint32_t rec_units,
                          float4 *__restrict__ recs)
{
  const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  const uint64_t pid = primid_perm[i];
  const uint32_t vcnt = pid_vcnt(pid);
  const uint32_t nv = vcnt == CB_PRIM_SPHERE ? 1 : vcnt == CB_PRIM_LINE ? 2 : vcnt == CB_PRIM_TRI ? 3 : vcnt == CB_PRIM_QUAD ? 4 : 0;
  float4 r[8];
  for(int k=0;k<8;k++) r[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  for(uint32_t v=0;v<nv;v++)
  {
    const cb_vtx_t *o = vertex_ptr(S, pid, v, 0);
    r[v] = make_float4(o->v[0], o->v[1], o->v[2], 0.0f);
    if(rec_units == 2)
    {
      const cb_vtx_t *c = vertex_ptr(S, pid, v, 1);   // static prims: same vertex again
      r[4+v] = make_float4(c->v[0], c->v[1], c->v[2], 0.0f);
    }
  }
  r[0].w = __uint_as_float((uint32_t)pid);
  r[1].w = __uint_as_float((uint32_t)(pid >> 32));
  if(vcnt == CB_PRIM_SPHERE) r[2].w = __uint_as_float(vertex_ptr(S, pid, 0, 0)->n);
  if(vcnt == CB_PRIM_LINE) { r[2].w = __uint_as_float(vertex_ptr(S, pid, 0, 0)->n); r[3].w = __uint_as_float(vertex_ptr(S, pid, 1, 0)->n); }
  float4 *dst = recs + (uint64_t)i*rec_units*4;
  for(uint32_t k=0;k<rec_units*4;k++) dst[k] = r[k];
}

// ---------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------
static inline int nblocks(uint64_t n) { return (int)((n + BUILD_BLOCK - 1)/BUILD_BLOCK); }

int cb200_build_records(cb200_accel *a, cudaStream_t stream)
{
  cb200_scene *s = a->scene;
  const uint32_t units = s->any_mb ? 2 : 1;
  const uint64_t n = s->num_prims;
  CB_CUDA(cudaMalloc(&a->d_recs, (n ? n : 1)*units*64));
  if(n)
  {
    SceneView S = { s->d_vtx, s->d_vtxidx, s->d_shapes };
    k_records<<<nblocks(n), BUILD_BLOCK, 0, stream>>>(S, a->d_primid, (uint32_t)n, units, a->d_recs);
    cb200_count_launch();
    CB_CUDA(cudaGetLastError());
  }
  a->dev.recs = a->d_recs;
  a->dev.rec_units = units;
  a->dev.num_prims = n;
  return 0;
}

template<typename T> struct DevBuf
{
  T *p = nullptr;
  ~DevBuf() { if(p) cudaFree(p); }
  cudaError_t alloc(uint64_t count) { return cudaMalloc(&p, (count ? count : 1)*sizeof(T)); }
};

int cb200_build_lbvh(cb200_accel *a, const float *ghost_aabb)
{
  cb200_scene *s = a->scene;
  const uint64_t n64 = s->num_prims;
  if(n64 >= 0x7fffffffull) { cb200_set_error("more than 2^31 primitives"); return CB200_ERR_UNSUPPORTED; }
  const uint32_t n = (uint32_t)n64;
  const int any_mb = s->any_mb;
  cudaStream_t st = 0;
  a->dev.mb = any_mb;

  float h_box[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  CB_CUDA(cudaMalloc(&a->d_primid, (n ? n : 1)*sizeof(uint64_t)));

  if(n == 0)
  { // root of four empty leaves (qbvhmp.c:1081-1099)
    if(any_mb) { cb200_set_error("internal: empty scene cannot have motion blur"); return CB200_ERR_ARG; }
    Node128 h;
    for(int c=0;c<4;c++) { for(int k=0;k<3;k++) { h.aabb0[k][c] = FLT_MAX; h.aabb0[k+3][c] = -FLT_MAX; } h.child[c] = CB_LEAF_BIT; }
    h.child[0] |= (uint64_t)(0 | (1 << 2) | (1 << 4)) << CB_AXIS_SHIFT;
    CB_CUDA(cudaMalloc(&a->d_nodes, sizeof(Node128)));
    CB_CUDA(cudaMemcpy(a->d_nodes, &h, sizeof(h), cudaMemcpyHostToDevice));
    a->dev.nodes = a->d_nodes; a->dev.num_nodes = 1; a->depth = 1;
    for(int k=0;k<6;k++) a->aabb[k] = h_box[k];
    if(ghost_aabb) for(int k=0;k<3;k++) { a->aabb[k] = fminf(ghost_aabb[k], a->aabb[k]); a->aabb[3+k] = fmaxf(ghost_aabb[3+k], a->aabb[3+k]); }
    return cb200_build_records(a, st);
  }

  SceneView S = { s->d_vtx, s->d_vtxidx, s->d_shapes };
  DevBuf<float> pbox0, pbox1, nbox0, nbox1, scene_box;
  DevBuf<uint64_t> codes, codes_sorted;
  DevBuf<uint32_t> index, index_sorted, node_parent, leaf_parent, flags, counters;
  DevBuf<BNode> bnodes;
  DevBuf<CollapseItem> q0, q1;
  DevBuf<uint8_t> cub_tmp;
  CB_CUDA(pbox0.alloc((uint64_t)n*6));
  CB_CUDA(pbox1.alloc(any_mb ? (uint64_t)n*6 : 1));
  CB_CUDA(scene_box.alloc(6));
  CB_CUDA(codes.alloc(n)); CB_CUDA(codes_sorted.alloc(n));
  CB_CUDA(index.alloc(n)); CB_CUDA(index_sorted.alloc(n));
  CB_CUDA(cudaMemcpyAsync(scene_box.p, h_box, sizeof(h_box), cudaMemcpyHostToDevice, st));

  k_prim_bounds<<<nblocks(n), BUILD_BLOCK, 0, st>>>(S, s->d_primid, n, any_mb, (Box *)pbox0.p, (Box *)pbox1.p, scene_box.p);
  k_morton<<<nblocks(n), BUILD_BLOCK, 0, st>>>((const Box *)pbox0.p, n, scene_box.p, codes.p, index.p);
  cb200_count_launch(2);
  CB_CUDA(cudaGetLastError());

  size_t tmp_bytes = 0;
  CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes.p, codes_sorted.p, index.p, index_sorted.p, (int)n, 0, 63, st));
  CB_CUDA(cub_tmp.alloc(tmp_bytes));
  CB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, codes.p, codes_sorted.p, index.p, index_sorted.p, (int)n, 0, 63, st));
  cb200_count_launch(4);

  k_permute_primid<<<nblocks(n), BUILD_BLOCK, 0, st>>>(s->d_primid, index_sorted.p, n, a->d_primid);
  cb200_count_launch();

  const uint32_t nin = n > 1 ? n - 1 : 1;
  CB_CUDA(bnodes.alloc(nin));
  CB_CUDA(node_parent.alloc(nin)); CB_CUDA(leaf_parent.alloc(n)); CB_CUDA(flags.alloc(nin));
  CB_CUDA(nbox0.alloc((uint64_t)nin*6)); CB_CUDA(nbox1.alloc(any_mb ? (uint64_t)nin*6 : 1));
  CB_CUDA(cudaMemsetAsync(node_parent.p, 0xff, sizeof(uint32_t)*nin, st));
  CB_CUDA(cudaMemsetAsync(leaf_parent.p, 0xff, sizeof(uint32_t)*n, st));
  CB_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(uint32_t)*nin, st));
  if(n > 1)
  {
    k_radix_tree<<<nblocks(n-1), BUILD_BLOCK, 0, st>>>(codes_sorted.p, (int)n, bnodes.p, node_parent.p, leaf_parent.p);
    k_refit<<<nblocks(n), BUILD_BLOCK, 0, st>>>(bnodes.p, node_parent.p, leaf_parent.p, index_sorted.p,
                                                pbox0.p, pbox1.p, (int)n, any_mb, nbox0.p, nbox1.p, flags.p);
    cb200_count_launch(2);
    CB_CUDA(cudaGetLastError());
  }

  // collapse, level by level.  a 4-wide node consumes >= 2 binary nodes except at the fringes; n wide nodes
  // always suffice (the reference sizes its buffer the same way, qbvhmp.c:299).
  // leaf size: the reference builds to <= 6 primitives with a SAH (qbvhmp.c:44); Morton-ordered leaves overlap more, and a
  // primitive test costs this kernel more than a node test (whole-leaf loops run at a third of the lanes), so smaller is better
  uint32_t leaf_max = LEAF_MAX_DEFAULT;
  if(const char *e = getenv("CB200_LEAF_MAX")) { const int v = atoi(e); if(v >= 1 && v <= 31) leaf_max = (uint32_t)v; }
  const uint64_t max_wide = (uint64_t)n + 8;
  if(any_mb) CB_CUDA(cudaMalloc(&a->d_nodes, max_wide*sizeof(Node256)));
  else       CB_CUDA(cudaMalloc(&a->d_nodes, max_wide*sizeof(Node128)));
  CB_CUDA(q0.alloc(max_wide)); CB_CUDA(q1.alloc(max_wide));
  CB_CUDA(counters.alloc(2));   // [0] = wide nodes allocated, [1] = next-level queue length
  uint32_t h_cnt[2] = {1, 0};
  CB_CUDA(cudaMemcpyAsync(counters.p, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, st));
  CollapseItem root = { n > 1 ? 0u : (BLEAF | 0u), 0u };
  CB_CUDA(cudaMemcpyAsync(q0.p, &root, sizeof(root), cudaMemcpyHostToDevice, st));
  CollapseArgs CA;
  CA.nodes = bnodes.p; CA.index = index_sorted.p;
  CA.pbox0 = pbox0.p; CA.pbox1 = pbox1.p; CA.nbox0 = nbox0.p; CA.nbox1 = nbox1.p;
  CA.n = (int)n; CA.any_mb = any_mb;
  CA.out256 = any_mb ? (Node256 *)a->d_nodes : nullptr;
  CA.out128 = any_mb ? nullptr : (Node128 *)a->d_nodes;
  CA.num_wide = counters.p;
  CA.max_wide = (uint32_t)max_wide;
  CA.leaf_max = leaf_max;
  uint32_t num_in = 1;
  int depth = 0;
  CollapseItem *qin = q0.p, *qout = q1.p;
  while(num_in)
  {
    depth++;
    if(depth > 100) { cb200_set_error("collapsed tree deeper than 100 levels"); return CB200_ERR_UNSUPPORTED; }
    k_collapse<<<nblocks(num_in), BUILD_BLOCK, 0, st>>>(CA, qin, num_in, qout, counters.p + 1, 0);
    cb200_count_launch();
    CB_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    num_in = h_cnt[1];
    if(h_cnt[0] > max_wide) { cb200_set_error("internal: wide node buffer overflow"); return CB200_ERR_NOMEM; }
    const uint32_t zero = 0;
    CB_CUDA(cudaMemcpyAsync(counters.p + 1, &zero, sizeof(zero), cudaMemcpyHostToDevice, st));
    CollapseItem *tq = qin; qin = qout; qout = tq;
  }
  a->depth = depth;
  { // right-size the node buffer
    const size_t nb = (size_t)h_cnt[0]*(any_mb ? sizeof(Node256) : sizeof(Node128));
    void *small = nullptr;
    CB_CUDA(cudaMalloc(&small, nb));
    CB_CUDA(cudaMemcpyAsync(small, a->d_nodes, nb, cudaMemcpyDeviceToDevice, st));
    CB_CUDA(cudaStreamSynchronize(st));
    cudaFree(a->d_nodes);
    a->d_nodes = small;
  }

  // ---- the 8-wide compressed tree over the same binary tree and primitive order (static scenes: the throughput path)
  a->d_nodes8 = nullptr; a->dev.nodes8 = nullptr; a->dev.num_nodes8 = 0; a->dev.depth8 = 0;
  {
    const char *e8 = getenv("CB200_BUILD_WIDE8");
    bool small_coords = true;   // the byte slab test keeps 2^e/d and (origin - p)/d finite for coordinates below 2^40 (traverse8.cu:ray8_ok)
    CB_CUDA(cudaMemcpyAsync(h_box, scene_box.p, sizeof(h_box), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    for(int k=0;k<6;k++) if(!(fabsf(h_box[k]) < 1099511627776.0f)) small_coords = false;
    if(!any_mb && n > 0 && small_coords && n < (1u << 28) && !(e8 && !atoi(e8)))
    {
      uint32_t leaf_max8 = LEAF_MAX_DEFAULT;
      if(const char *e = getenv("CB200_LEAF_MAX8")) { const int v = atoi(e); if(v >= 1 && v <= 7) leaf_max8 = (uint32_t)v; }
      Node8 *big = nullptr;
      CB_CUDA(cudaMalloc(&big, max_wide*sizeof(Node8)));
      uint32_t h8[2] = {1, 0};
      CB_CUDA(cudaMemcpyAsync(counters.p, h8, sizeof(h8), cudaMemcpyHostToDevice, st));
      CB_CUDA(cudaMemcpyAsync(q0.p, &root, sizeof(root), cudaMemcpyHostToDevice, st));
      Collapse8Args C8;
      C8.nodes = bnodes.p; C8.index = index_sorted.p; C8.pbox0 = pbox0.p; C8.nbox0 = nbox0.p;
      C8.out = big; C8.num_wide = counters.p; C8.max_wide = (uint32_t)max_wide; C8.leaf_max = leaf_max8;
      uint32_t num8 = 1;
      int depth8 = 0;
      qin = q0.p; qout = q1.p;
      while(num8)
      {
        if(++depth8 > 100) { cudaFree(big); cb200_set_error("collapsed 8-wide tree deeper than 100 levels"); return CB200_ERR_UNSUPPORTED; }
        k_collapse8<<<nblocks(num8), BUILD_BLOCK, 0, st>>>(C8, qin, num8, qout, counters.p + 1);
        cb200_count_launch();
        CB_CUDA(cudaMemcpyAsync(h8, counters.p, sizeof(h8), cudaMemcpyDeviceToHost, st));
        CB_CUDA(cudaStreamSynchronize(st));
        num8 = h8[1];
        if(h8[0] > max_wide) { cudaFree(big); cb200_set_error("internal: 8-wide node buffer overflow"); return CB200_ERR_NOMEM; }
        const uint32_t zero = 0;
        CB_CUDA(cudaMemcpyAsync(counters.p + 1, &zero, sizeof(zero), cudaMemcpyHostToDevice, st));
        CollapseItem *tq = qin; qin = qout; qout = tq;
      }
      if(h8[0] < (1u << 24))
      { // a stack word holds the node index in 24 bits
        CB_CUDA(cudaMalloc(&a->d_nodes8, (size_t)h8[0]*sizeof(Node8)));
        CB_CUDA(cudaMemcpyAsync(a->d_nodes8, big, (size_t)h8[0]*sizeof(Node8), cudaMemcpyDeviceToDevice, st));
        CB_CUDA(cudaStreamSynchronize(st));
        a->dev.nodes8 = a->d_nodes8; a->dev.num_nodes8 = h8[0]; a->dev.depth8 = (uint32_t)depth8;
        const char *ew = getenv("CB200_WIDE8");
        a->traversal = (ew && atoi(ew)) ? CB200_TRAVERSAL_WIDE8 : CB200_TRAVERSAL_EXACT4;   // measured slower on the bench (profiles/README.md r2b): opt-in
      }
      cudaFree(big);
    }
  }
  a->dev.nodes = a->d_nodes;
  a->dev.num_nodes = h_cnt[0];
  if(any_mb)
  {
    k_parents<<<nblocks(h_cnt[0]), BUILD_BLOCK, 0, st>>>((Node256 *)a->d_nodes, h_cnt[0]);
    cb200_count_launch();
  }
  CB_CUDA(cudaMemcpyAsync(h_box, scene_box.p, sizeof(h_box), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaStreamSynchronize(st));
  for(int k=0;k<6;k++) a->aabb[k] = h_box[k];
  if(ghost_aabb) for(int k=0;k<3;k++) { a->aabb[k] = fminf(ghost_aabb[k], a->aabb[k]); a->aabb[3+k] = fmaxf(ghost_aabb[3+k], a->aabb[3+k]); }
  int rc = cb200_build_records(a, st);
  if(rc) return rc;
  CB_CUDA(cudaStreamSynchronize(st));
  return 0;
}
