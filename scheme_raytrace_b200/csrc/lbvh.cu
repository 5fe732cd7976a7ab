// lbvh.cu — GPU LBVH build: primitive bounds -> Morton keys -> LSD radix sort -> Karras radix
// tree emit -> bottom-up refit.  Supersedes the reference's CPU builders make-bvh-node
// (geometry.scm:226-260) and make-bvh-with-sah (geometry.scm:294-371); tree topology is not a
// parity target, the build is instead bit-exact against the sequential host reference
// oracle/lbvh_ref.cpp over the same AABBs (tests/test_gpu_parity.py: test_lbvh_bit_exact).
//
// Algorithm specification (DESIGN.md "LBVH"); every fp32 op below is an explicitly rounded
// intrinsic so that the host reference (compiled with -ffp-contract=off) produces the same bits:
//   centroid = 0.5*(min+max); g = min(uint(q * 2^21), 2^21-1), q = (c - cmin)/(cmax - cmin)
//   key = 63-bit Morton (x highest); stable sort; Karras 2012 with index tie-break for equal keys
//   primitives whose AABB extent is >= half the scene's (at most SRT_MAX_GLOBAL, largest first)
//   are kept OUT of the tree and tested linearly before traversal (like the reference's own
//   (list ground bvh-node) scenes, main.scm:215-235)
//   stored child boxes = union of primitive AABBs as centre c = 0.5*(min+max) and half extent
//   e = 0.5*(max-min) + S * 2^-21 (S = max |coordinate|; the pad covers fp32 rounding in c, e and in
//   the traversal's FMA slab test)
#include "srt_device.cuh"
#include "srt_host.h"

namespace {

constexpr float BIG = 3.0e38f;

// ---- 1. primitive bounds (geometry.scm bbox closures; instances get correct boxes, unlike Q7) ---
__device__ void prim_box(const DScene& sc, int i, float cam_t0, float cam_t1, float3& mn, float3& mx) {
  int4 hdr = sc.prim_hdr[i];
  int type = hdr.x & 0xff;
  float4 a = sc.prim_a[i];
  if (type == SRT_PRIM_SPHERE) {                                   // geometry.scm:172-174
    float r = fabsf(a.w);
    mn = v3(a.x - r, a.y - r, a.z - r); mx = v3(a.x + r, a.y + r, a.z + r);
  } else if (type == SRT_PRIM_MOVING_SPHERE) {                     // geometry.scm:209-214, over the shutter AND time 0 (Q6, Q13)
    float4 b = sc.prim_b[i], c = sc.prim_c[i];
    float r = fabsf(a.w);
    float times[3] = {cam_t0, cam_t1, 0.0f};
    mn = v3(BIG, BIG, BIG); mx = v3(-BIG, -BIG, -BIG);
    for (int k = 0; k < 3; ++k) {
      float3 cc = moving_center(a, b, c, times[k]);
      mn = v3(fminf(mn.x, cc.x - r), fminf(mn.y, cc.y - r), fminf(mn.z, cc.z - r));
      mx = v3(fmaxf(mx.x, cc.x + r), fmaxf(mx.y, cc.y + r), fmaxf(mx.z, cc.z + r));
    }
  } else if (type <= SRT_PRIM_YZ_RECT) {                           // geometry.scm:390-392, 409-411, 428-430
    float k = sc.prim_b[i].x;
    if (type == SRT_PRIM_XY_RECT) { mn = v3(a.x, a.z, k - 0.0001f); mx = v3(a.y, a.w, k + 0.0001f); }
    else if (type == SRT_PRIM_XZ_RECT) { mn = v3(a.x, k - 0.0001f, a.z); mx = v3(a.y, k + 0.0001f, a.w); }
    else { mn = v3(k - 0.0001f, a.x, a.z); mx = v3(k + 0.0001f, a.y, a.w); }
  } else if (type == SRT_PRIM_KLEIN) {                             // no bounding box upstream (geometry.scm:662-663): never enters the tree
    mn = v3(a.x, a.y, a.z); mx = mn;
  } else if (type == SRT_PRIM_PATCH) {                             // convex hull of the 4x4 control net
    const float4* cp = sc.patch_cp + 16 * hdr.w;
    mn = v3(BIG, BIG, BIG); mx = v3(-BIG, -BIG, -BIG);
    for (int k = 0; k < 16; ++k) { float4 q = cp[k]; mn = v3(fminf(mn.x, q.x), fminf(mn.y, q.y), fminf(mn.z, q.z)); mx = v3(fmaxf(mx.x, q.x), fmaxf(mx.y, q.y), fmaxf(mx.z, q.z)); }
  } else {                                                         // bezier.scm:88-98
    float w1 = 0.5f * a.w;
    float4 b = sc.prim_b[i], c = sc.prim_c[i], d = sc.prim_d[i];
    mn = v3(fminf(fminf(a.x, b.x), fminf(c.x, d.x)) - w1, fminf(fminf(a.y, b.y), fminf(c.y, d.y)) - w1, fminf(fminf(a.z, b.z), fminf(c.z, d.z)) - w1);
    mx = v3(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)) + w1, fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)) + w1, fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)) + w1);
  }
  if (hdr.z >= 0) {                                                // geometry.scm:479-480, 490-509 (8 corners)
    Xf x = load_xf(sc, hdr.z);
    float3 wmn = v3(BIG, BIG, BIG), wmx = v3(-BIG, -BIG, -BIG);
    for (int c = 0; c < 8; ++c) {
      float3 p = xf_point_to_world(x, v3((c & 1) ? mx.x : mn.x, (c & 2) ? mx.y : mn.y, (c & 4) ? mx.z : mn.z));
      wmn = v3(fminf(wmn.x, p.x), fminf(wmn.y, p.y), fminf(wmn.z, p.z));
      wmx = v3(fmaxf(wmx.x, p.x), fmaxf(wmx.y, p.y), fmaxf(wmx.z, p.z));
    }
    // the rotation is evaluated in fp32: widen by a relative epsilon so the box stays conservative
    float e = 4e-7f * fmaxf(fmaxf(fabsf(wmn.x), fabsf(wmx.x)), fmaxf(fmaxf(fabsf(wmn.y), fabsf(wmx.y)), fmaxf(fabsf(wmn.z), fabsf(wmx.z))));
    mn = v3(wmn.x - e, wmn.y - e, wmn.z - e); mx = v3(wmx.x + e, wmx.y + e, wmx.z + e);
  }
}
__global__ void k_prim_bounds(DScene sc, float cam_t0, float cam_t1, float* __restrict__ aabb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sc.n_surf) return;
  float3 mn, mx;
  if ((sc.prim_hdr[i].x & 0xff) == SRT_PRIM_CONSTANT_MEDIUM) {     // geometry.scm:576-577: the boundary's box
    float4 a = sc.prim_a[i];
    mn = v3(BIG, BIG, BIG); mx = v3(-BIG, -BIG, -BIG);
    for (int j = (int)a.y; j < (int)a.y + (int)a.z; ++j) {
      float3 bm, bx; prim_box(sc, j, cam_t0, cam_t1, bm, bx);
      mn = v3(fminf(mn.x, bm.x), fminf(mn.y, bm.y), fminf(mn.z, bm.z)); mx = v3(fmaxf(mx.x, bx.x), fmaxf(mx.y, bx.y), fmaxf(mx.z, bx.z));
    }
  } else prim_box(sc, i, cam_t0, cam_t1, mn, mx);
  float* o = aabb + 6 * (size_t)i;
  o[0] = mn.x; o[1] = mn.y; o[2] = mn.z; o[3] = mx.x; o[4] = mx.y; o[5] = mx.z;
}

// ---- 2. centroid bounds + S (order-independent: min/max are exact) ------------------------------
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__global__ void k_bounds_init(int* b) { if (threadIdx.x < 3) b[threadIdx.x] = f2ord(BIG); else if (threadIdx.x < 6) b[threadIdx.x] = f2ord(-BIG); else if (threadIdx.x == 6) b[6] = f2ord(0.0f); }
__global__ void k_bounds_reduce(int n, const int* __restrict__ item_prim, const float* __restrict__ aabb, int* __restrict__ b) {
  float cmin[3] = {BIG, BIG, BIG}, cmax[3] = {-BIG, -BIG, -BIG}, S = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* q = aabb + 6 * (size_t)item_prim[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float c = __fmul_rn(0.5f, __fadd_rn(q[k], q[3 + k]));
      cmin[k] = fminf(cmin[k], c); cmax[k] = fmaxf(cmax[k], c);
      S = fmaxf(S, fmaxf(fabsf(q[k]), fabsf(q[3 + k])));
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o; o >>= 1) { cmin[k] = fminf(cmin[k], __shfl_xor_sync(0xffffffffu, cmin[k], o)); cmax[k] = fmaxf(cmax[k], __shfl_xor_sync(0xffffffffu, cmax[k], o)); }
  }
  for (int o = 16; o; o >>= 1) S = fmaxf(S, __shfl_xor_sync(0xffffffffu, S, o));
  if ((threadIdx.x & 31) == 0) {
    for (int k = 0; k < 3; ++k) { atomicMin(&b[k], f2ord(cmin[k])); atomicMax(&b[3 + k], f2ord(cmax[k])); }
    atomicMax(&b[6], f2ord(S));
  }
}

// ---- 3. Morton keys ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long expand21(unsigned int v) {
  unsigned long long x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ unsigned long long morton_key(const float* __restrict__ q, const int* b) {
  unsigned int g[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float cmin = ord2f(b[k]), cmax = ord2f(b[3 + k]);
    float c = __fmul_rn(0.5f, __fadd_rn(q[k], q[3 + k]));
    float ext = __fsub_rn(cmax, cmin);
    float qq = ext > 0.0f ? __fdiv_rn(__fsub_rn(c, cmin), ext) : 0.0f;
    float scv = __fmul_rn(qq, 2097152.0f);
    unsigned int gi = (unsigned int)scv;
    g[k] = gi < 2097151u ? gi : 2097151u;
  }
  return (expand21(g[0]) << 2) | (expand21(g[1]) << 1) | expand21(g[2]);
}
__global__ void k_morton(int n, const int* __restrict__ item_prim, const float* __restrict__ aabb, const int* __restrict__ b, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = morton_key(aabb + 6 * (size_t)item_prim[i], b);
  vals[i] = i;
}

// ---- 4. stable LSD radix sort, 8-bit digits, 64-bit keys + 32-bit payload ------------------------
constexpr int RS_BLOCK = 256;
__global__ void k_rs_hist(const unsigned long long* __restrict__ keys, int n, int shift, int* __restrict__ hist, int nblk) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int i = blockIdx.x * RS_BLOCK + threadIdx.x;
  if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & 255ull)], 1);
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}
__global__ void k_rs_scan(int* __restrict__ hist, int total) {   // exclusive scan, single CTA of 1024
  __shared__ int wsum[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    int idx = base + threadIdx.x;
    int v = idx < total ? hist[idx] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    int carry = carry_s;
    int incl = x + ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0);
    if (idx < total) hist[idx] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
}
__global__ void k_rs_scatter(const unsigned long long* __restrict__ kin, const int* __restrict__ vin,
                             unsigned long long* __restrict__ kout, int* __restrict__ vout, int n, int shift,
                             const int* __restrict__ hist, int nblk) {
  __shared__ int wc[RS_BLOCK / 32][256];
  for (int j = threadIdx.x; j < (RS_BLOCK / 32) * 256; j += RS_BLOCK) (&wc[0][0])[j] = 0;
  __syncthreads();
  int i = blockIdx.x * RS_BLOCK + threadIdx.x;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bool active = i < n;
  unsigned long long key = 0; int val = 0, d = 0, rank = 0;
  unsigned int amask = __ballot_sync(0xffffffffu, active);
  if (active) {
    key = kin[i]; val = vin[i];
    d = (int)((key >> shift) & 255ull);
    unsigned int peers = __match_any_sync(amask, d);
    rank = __popc(peers & ((1u << lane) - 1u));
    if (rank == 0) wc[warp][d] = __popc(peers);
  }
  __syncthreads();
  if (active) {
    int off = 0;
    for (int w = 0; w < warp; ++w) off += wc[w][d];
    int pos = hist[d * nblk + blockIdx.x] + off + rank;
    if (SRT_BOUNDS_OK(pos >= 0 && pos < n, 301)) { kout[pos] = key; vout[pos] = val; }
  }
}

// ---- 5. Karras 2012 radix tree ------------------------------------------------------------------
__device__ __forceinline__ int delta(const unsigned long long* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long a = keys[i], b = keys[j];
  if (a != b) return __clzll((long long)(a ^ b));
  return 64 + __clz(i ^ j);
}
__device__ __forceinline__ void karras_node(int i, int n, const unsigned long long* keys, const int* order, const int* __restrict__ item_prim,
                                            int4* links /* left right parent sibling per node */, int* leaf_parent) {
  int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(keys, n, i, j);
  int s = 0, t = l;
  do { t = (t + 1) >> 1; if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int a = min(i, j), b = max(i, j);
  if (!SRT_BOUNDS_OK(j >= 0 && j < n && gamma >= 0 && gamma + 1 < n, 311)) return;
  int left = (a == gamma) ? ~item_prim[order[gamma]] : gamma;          // leaf reference = ~primitive id
  int right = (b == gamma + 1) ? ~item_prim[order[gamma + 1]] : gamma + 1;
  links[i].x = left; links[i].y = right;
  if (left >= 0) { links[left].z = i; links[left].w = right; } else leaf_parent[gamma] = i;
  if (right >= 0) { links[right].z = i; links[right].w = left; } else leaf_parent[gamma + 1] = i;
  if (i == 0) { links[0].z = -1; links[0].w = -1; }
}
__global__ void k_karras(int n, const unsigned long long* __restrict__ keys, const int* __restrict__ order, const int* __restrict__ item_prim,
                         int4* __restrict__ links, int* __restrict__ leaf_parent) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  karras_node(i, n, keys, order, item_prim, links, leaf_parent);
}

// ---- 6. bottom-up refit with atomic visit counters ----------------------------------------------
// nbox[i] = unpadded box of internal node i (scratch); stored child boxes are padded on write.
// CTA_SCOPE: boxes and counters live in shared memory (single-CTA build): a block-scope fence orders them; the device-scope
// fence of the multi-CTA build would also wait for the node stores of every level to reach L2 (~1 us per level).
template <bool CTA_SCOPE = false>
__device__ __forceinline__ void refit_from_leaf(int pos, int n, const float* __restrict__ aabb, const int4* links,
                                                const int* leaf_parent, const int* b, float* nbox, int* visit,
                                                float4* nodes, int* depth_out) {
  const float pad = __fmul_rn(ord2f(b[6]), 1.0f / 2097152.0f);
  int node = leaf_parent[pos];
  int depth = 0;   // internal ancestors of this leaf
  {  // count the depth first (read-only walk), used for the stackless-trail bound
    int p = node; while (p >= 0) { ++depth; p = links[p].z; }
    atomicMax(depth_out, depth);
  }
  while (node >= 0) {
    if (!SRT_BOUNDS_OK(node < n - 1, 321)) return;
    if (CTA_SCOPE) __threadfence_block(); else __threadfence();
    if (atomicAdd(&visit[node], 1) == 0) return;      // first arrival: the sibling subtree finishes this node
    int4 lk = links[node];
    float cb[2][6];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      int ch = c ? lk.y : lk.x;
      if (ch < 0) { const float* q = aabb + 6 * (size_t)(~ch); for (int k = 0; k < 6; ++k) cb[c][k] = q[k]; }
      else { const volatile float* q = nbox + 6 * (size_t)ch; for (int k = 0; k < 6; ++k) cb[c][k] = q[k]; }
    }
    float* o = nbox + 6 * (size_t)node;
    for (int k = 0; k < 3; ++k) { o[k] = fminf(cb[0][k], cb[1][k]); o[3 + k] = fmaxf(cb[0][3 + k], cb[1][3 + k]); }
    float l[6], r[6];   // centre (0..2), padded half extent (3..5)
    for (int k = 0; k < 3; ++k) {
      l[k] = __fmul_rn(0.5f, __fadd_rn(cb[0][k], cb[0][3 + k])); l[3 + k] = __fadd_rn(__fmul_rn(0.5f, __fsub_rn(cb[0][3 + k], cb[0][k])), pad);
      r[k] = __fmul_rn(0.5f, __fadd_rn(cb[1][k], cb[1][3 + k])); r[3 + k] = __fadd_rn(__fmul_rn(0.5f, __fsub_rn(cb[1][3 + k], cb[1][k])), pad);
    }
    nodes[4 * node + 0] = make_float4(l[0], l[1], l[2], l[3]);
    nodes[4 * node + 1] = make_float4(l[4], l[5], r[0], r[1]);
    nodes[4 * node + 2] = make_float4(r[2], r[3], r[4], r[5]);
    nodes[4 * node + 3] = make_float4(__int_as_float(lk.x), __int_as_float(lk.y), __int_as_float(lk.z), __int_as_float(lk.w));
    node = lk.z;
  }
}
__global__ void k_refit(int n, const int* __restrict__ order, const float* __restrict__ aabb, const int4* __restrict__ links,
                        const int* __restrict__ leaf_parent, const int* __restrict__ b, float* nbox, int* visit,
                        float4* __restrict__ nodes, int* __restrict__ depth_out) {
  int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  (void)order;
  refit_from_leaf(pos, n, aabb, links, leaf_parent, b, nbox, visit, nodes, depth_out);
}

// n <= 1: a single node.
__global__ void k_single_node(int n, const int* __restrict__ item_prim, const float* __restrict__ aabb_all, const int* __restrict__ b, float4* __restrict__ nodes,
                              unsigned long long* keys, int* order) {
  if (threadIdx.x != 0) return;
  float l[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int prim = 0;
  if (n == 1) {
    prim = item_prim[0];
    const float* aabb = aabb_all + 6 * (size_t)prim;
    const float pad = __fmul_rn(ord2f(b[6]), 1.0f / 2097152.0f);
    for (int k = 0; k < 3; ++k) { l[k] = __fmul_rn(0.5f, __fadd_rn(aabb[k], aabb[3 + k])); l[3 + k] = __fadd_rn(__fmul_rn(0.5f, __fsub_rn(aabb[3 + k], aabb[k])), pad); }
    order[0] = 0;
  }
  // both children reference the same leaf with the same box; n == 0 is never traversed
  nodes[0] = make_float4(l[0], l[1], l[2], l[3]);
  nodes[1] = make_float4(l[4], l[5], l[0], l[1]);
  nodes[2] = make_float4(l[2], l[3], l[4], l[5]);
  nodes[3] = make_float4(__int_as_float(~prim), __int_as_float(~prim), __int_as_float(-1), __int_as_float(-1));
}

// Surface-area sums of a built tree: out[0] = sum over internal nodes, out[1] = sum over leaf
// (child) boxes, out[2] = root.  One CTA, per-thread strided partial sums in double and a fixed
// shared-memory tree: the result does not depend on scheduling.
__device__ __forceinline__ void tree_area_256(int n_items, const float4* nodes, double* __restrict__ out, double (*sh)[256]) {
  const int n_nodes = n_items > 1 ? n_items - 1 : 1;
  double a_int = 0.0, a_leaf = 0.0, a_root = 0.0;
  for (int i = threadIdx.x < 256 ? (int)threadIdx.x : n_nodes; i < n_nodes; i += 256) {   // (the single-CTA build calls this with 1024 threads: 256 of them sum)
    const float4 f0 = nodes[4 * i], f1 = nodes[4 * i + 1], f2 = nodes[4 * i + 2], f3 = nodes[4 * i + 3];
    const double le[3] = {f0.w, f1.x, f1.y}, re[3] = {f2.y, f2.z, f2.w};
    const double al = 8.0 * (le[0] * le[1] + le[1] * le[2] + le[0] * le[2]), ar = 8.0 * (re[0] * re[1] + re[1] * re[2] + re[0] * re[2]);
    if (__float_as_int(f3.x) >= 0) a_int += al; else a_leaf += al;
    if (n_items > 1) { if (__float_as_int(f3.y) >= 0) a_int += ar; else a_leaf += ar; }
    if (i == 0) {
      const double lc[3] = {f0.x, f0.y, f0.z}, rc[3] = {f1.z, f1.w, f2.x};
      double d[3];
      for (int k = 0; k < 3; ++k) d[k] = fmax(lc[k] + le[k], rc[k] + re[k]) - fmin(lc[k] - le[k], rc[k] - re[k]);
      a_root = 2.0 * (d[0] * d[1] + d[1] * d[2] + d[0] * d[2]);
    }
  }
  if (threadIdx.x < 256) { sh[0][threadIdx.x] = a_int; sh[1][threadIdx.x] = a_leaf; }
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sh[0][threadIdx.x] += sh[0][threadIdx.x + s]; sh[1][threadIdx.x] += sh[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh[0][0] + a_root; out[1] = sh[1][0]; out[2] = a_root; }
}
__global__ void __launch_bounds__(256) k_tree_area(int n_items, const float4* __restrict__ nodes, double* __restrict__ out) {
  __shared__ double sh[2][256];
  tree_area_256(n_items, nodes, out, sh);
}

// ---- single-CTA build for small scenes -----------------------------------------------------------
// Every scene of the reference has at most a few hundred primitives; the multi-kernel build above is then ~29
// launches of one or two CTAs each, and the commit-time search for the primitives kept outside the tree builds up
// to nine candidate trees (srt_api.cu) - 0.7 ms of launch latency for cfg2.  Here ONE CTA builds one candidate from
// bounds to surface-area sums, and the candidates of a commit run side by side as the CTAs of one launch.  Same
// arithmetic (the device functions above), same tree: the sort is a bitonic network on (key, index) pairs, which is
// the order the stable LSD radix sort produces; the refit uses the same visit-counter protocol.
constexpr int SMALL_MAX_ITEMS = 2048, SMALL_THREADS = 1024;
// Everything the build chases pointers through lives in shared memory (keys, order, links, leaf parents, the unpadded node
// boxes and the visit counters: 120 KB): the refit walks ~depth dependent steps per leaf, each of which was an L2 round trip
// (~0.5 us) when links / boxes were in global memory.  Only the results (keys, order, nodes, area, depth) are written out.
constexpr size_t SMALL_SMEM = (size_t)SMALL_MAX_ITEMS * (8 + 16 + 24 + 4 + 4 + 4);
__global__ void __launch_bounds__(SMALL_THREADS) k_lbvh_small(const float* __restrict__ aabb, const SrtSmallJob* __restrict__ jobs) {
  const SrtSmallJob J = jobs[blockIdx.x];
  const int n = J.n, tid = threadIdx.x;
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned long long* s_key = (unsigned long long*)s_raw;
  int4* s_links = (int4*)(s_raw + (size_t)SMALL_MAX_ITEMS * 8);
  float* s_nbox = (float*)(s_raw + (size_t)SMALL_MAX_ITEMS * 24);
  int* s_val = (int*)(s_raw + (size_t)SMALL_MAX_ITEMS * 48);
  int* s_visit = s_val + SMALL_MAX_ITEMS;
  int* s_lp = s_visit + SMALL_MAX_ITEMS;
  __shared__ int s_b[8];
  __shared__ int s_depth;
  __shared__ double s_red[2][256];
  if (tid < 3) s_b[tid] = f2ord(BIG); else if (tid < 6) s_b[tid] = f2ord(-BIG); else if (tid == 6) s_b[6] = f2ord(0.0f);
  if (tid == 0) s_depth = 0;
  __syncthreads();
  {   // centroid bounds + S (exact min / max: independent of the order)
    float cmin[3] = {BIG, BIG, BIG}, cmax[3] = {-BIG, -BIG, -BIG}, S = 0.0f;
    for (int i = tid; i < n; i += SMALL_THREADS) {
      const float* q = aabb + 6 * (size_t)J.item_prim[i];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float c = __fmul_rn(0.5f, __fadd_rn(q[k], q[3 + k]));
        cmin[k] = fminf(cmin[k], c); cmax[k] = fmaxf(cmax[k], c);
        S = fmaxf(S, fmaxf(fabsf(q[k]), fabsf(q[3 + k])));
      }
    }
    // warp reduction first: 7 shared atomics per warp instead of per thread
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { cmin[k] = fminf(cmin[k], __shfl_xor_sync(0xffffffffu, cmin[k], o)); cmax[k] = fmaxf(cmax[k], __shfl_xor_sync(0xffffffffu, cmax[k], o)); }
      S = fmaxf(S, __shfl_xor_sync(0xffffffffu, S, o));
    }
    if ((tid & 31) == 0 && tid < ((n + 31) & ~31)) {
      for (int k = 0; k < 3; ++k) { atomicMin(&s_b[k], f2ord(cmin[k])); atomicMax(&s_b[3 + k], f2ord(cmax[k])); }
      atomicMax(&s_b[6], f2ord(S));
    }
  }
  __syncthreads();
  int npad = 2; while (npad < n) npad <<= 1;
  for (int i = tid; i < npad; i += SMALL_THREADS) {
    if (i < n) { s_key[i] = morton_key(aabb + 6 * (size_t)J.item_prim[i], s_b); s_val[i] = i; }
    else { s_key[i] = ~0ull; s_val[i] = 0x7fffffff; }
  }
  // bitonic network; element i belongs to thread i % SMALL_THREADS, so partners at distance j < 32 are exchanged inside
  // one warp: those steps need a warp barrier only (35 of the 45 steps for 512 items)
  int jprev = 1;
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32 || jprev >= 32) __syncthreads(); else __syncwarp();
      jprev = j;
      for (int i = tid; i < npad; i += SMALL_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long ka = s_key[i], kb = s_key[l]; const int va = s_val[i], vb = s_val[l];
          const bool gt = ka > kb || (ka == kb && va > vb);
          if (gt == ((i & k) == 0)) { s_key[i] = kb; s_key[l] = ka; s_val[i] = vb; s_val[l] = va; }
        }
      }
    }
  __syncthreads();
  for (int i = tid; i < n; i += SMALL_THREADS) { J.keys[i] = s_key[i]; J.order[i] = s_val[i]; s_visit[i] = 0; }
  for (int i = tid; i < n - 1; i += SMALL_THREADS) karras_node(i, n, s_key, s_val, J.item_prim, s_links, s_lp);
  __syncthreads();
  for (int pos = tid; pos < n; pos += SMALL_THREADS) refit_from_leaf<true>(pos, n, aabb, s_links, s_lp, s_b, s_nbox, s_visit, J.nodes, &s_depth);
  __syncthreads();
  tree_area_256(n, J.nodes, J.area, s_red);
  if (tid == 0) { *J.depth = s_depth; for (int k = 0; k < 7; ++k) J.bounds[k] = s_b[k]; }
}

}  // namespace

unsigned long long srt_bounds_violations_lbvh(int* first) {
#ifdef SRT_BOUNDS_CHECK
  unsigned long long v = 0ull; int f = 0;
  cudaMemcpyFromSymbol(&v, d_srt_violations, sizeof(v)); cudaMemcpyFromSymbol(&f, d_srt_first_violation, sizeof(f));
  if (first) *first = f;
  return v;
#else
  if (first) *first = 0;
  return 0ull;
#endif
}

// The candidate trees of one commit (n_jobs <= 16, each 2 <= n <= SRT_SMALL_MAX_ITEMS items), one CTA each.
int srt_lbvh_build_small(const float* d_aabb, const SrtSmallJob* d_jobs, int n_jobs, cudaStream_t stream) {
  static bool attr_set[SRT_MAX_DEVICES] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev >= 0 && dev < SRT_MAX_DEVICES && !attr_set[dev]) {
    if (cudaFuncSetAttribute(k_lbvh_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM) != cudaSuccess) return -1;
    attr_set[dev] = true;
  }
  k_lbvh_small<<<n_jobs, SMALL_THREADS, SMALL_SMEM, stream>>>(d_aabb, d_jobs);
  return 1;
}

// Surface-area sums of the tree just built (see k_tree_area); d_out = 3 doubles on the device.
int srt_lbvh_tree_area(int n_items, LbvhBuffers& B, double* d_out, cudaStream_t stream) {
  k_tree_area<<<1, 256, 0, stream>>>(n_items, B.d_nodes, d_out);
  return 1;
}

// Phase A: primitive AABBs of every scene surface (by primitive id).
int srt_lbvh_bounds(const DScene& sc, float cam_t0, float cam_t1, LbvhBuffers& B, cudaStream_t stream) {
  if (sc.n_surf > 0) k_prim_bounds<<<(sc.n_surf + 127) / 128, 128, 0, stream>>>(sc, cam_t0, cam_t1, B.d_aabb);
  return 1;
}

// Phase B: LBVH over the n_items primitives listed in B.d_item_prim (the "huge" primitives that
// are tested linearly before traversal are left out).  All launches on `stream`.
int srt_lbvh_build(int n_items, LbvhBuffers& B, cudaStream_t stream) {
  int n = n_items, launches = 0;
  k_bounds_init<<<1, 32, 0, stream>>>(B.d_bounds); ++launches;
  if (n > 0) {
    int rb = min((n + 255) / 256, 1024);
    k_bounds_reduce<<<rb, 256, 0, stream>>>(n, B.d_item_prim, B.d_aabb, B.d_bounds); ++launches;
  }
  if (n <= 1) {
    k_single_node<<<1, 32, 0, stream>>>(n, B.d_item_prim, B.d_aabb, B.d_bounds, B.d_nodes, B.d_keys[0], B.d_order[0]); ++launches;
    B.sorted = 0;
    int one = n;   // depth = number of internal nodes on the path
    cudaMemcpyAsync(B.d_depth, &one, sizeof(int), cudaMemcpyHostToDevice, stream);
    return launches;
  }
  k_morton<<<(n + 255) / 256, 256, 0, stream>>>(n, B.d_item_prim, B.d_aabb, B.d_bounds, B.d_keys[0], B.d_order[0]); ++launches;
  int nblk = (n + RS_BLOCK - 1) / RS_BLOCK;
  int cur = 0;
  for (int pass = 0; pass < 8; ++pass) {
    int shift = 8 * pass;
    k_rs_hist<<<nblk, RS_BLOCK, 0, stream>>>(B.d_keys[cur], n, shift, B.d_hist, nblk);
    k_rs_scan<<<1, 1024, 0, stream>>>(B.d_hist, 256 * nblk);
    k_rs_scatter<<<nblk, RS_BLOCK, 0, stream>>>(B.d_keys[cur], B.d_order[cur], B.d_keys[cur ^ 1], B.d_order[cur ^ 1], n, shift, B.d_hist, nblk);
    launches += 3; cur ^= 1;
  }
  B.sorted = cur;
  cudaMemsetAsync(B.d_visit, 0, sizeof(int) * (size_t)(n - 1), stream);
  cudaMemsetAsync(B.d_depth, 0, sizeof(int), stream);
  k_karras<<<(n - 1 + 127) / 128, 128, 0, stream>>>(n, B.d_keys[cur], B.d_order[cur], B.d_item_prim, B.d_links, B.d_leaf_parent); ++launches;
  k_refit<<<(n + 127) / 128, 128, 0, stream>>>(n, B.d_order[cur], B.d_aabb, B.d_links, B.d_leaf_parent, B.d_bounds, B.d_nbox, B.d_visit, B.d_nodes, B.d_depth); ++launches;
  return launches;
}
