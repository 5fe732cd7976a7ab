// lbvh_ref.cpp — sequential HOST REFERENCE BUILD of the LBVH, the bit-exact checker for the GPU
// build kernels (north star: "GPU LBVH build: Morton codes, radix sort and Karras tree emit,
// bit-exact against a host reference build over the same AABBs").
//
// *** TEST INFRASTRUCTURE ONLY *** — never linked into the product.
//
// This replaces the reference's CPU BVH builders (geometry.scm:226-260 make-bvh-node,
// geometry.scm:294-371 make-bvh-with-sah), whose tree topology is NOT a parity target
// (SURVEY.md §2 row 10, §8a row G10).  The algorithm is specified in DESIGN.md "LBVH" and is
// restated here independently of the CUDA sources:
//   1. centroid = 0.5f*(min+max); centroid bounds cmin/cmax
//   2. 21-bit grid coordinate per axis, 63-bit Morton key (x highest)
//   3. stable sort by key (ties keep primitive order)
//   4. Karras 2012 radix-tree emit with delta(i,j) = clz64(ki^kj), or 64+clz32(i^j) for equal keys
//   5. child boxes = union of primitive AABBs, stored as centre 0.5*(min+max) and half extent
//      0.5*(max-min) + S*2^-21 (S = largest |coordinate|)
// Compile with -ffp-contract=off (see Makefile) so every fp32 op is separately rounded, matching
// the __f*_rn intrinsics of the device code.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>

namespace {

struct Node64 {   // == SrtBvhNode in include/srt.h (64 bytes)
  float lc[3], le[3], rc[3], re[3];   // child boxes as centre / padded half extent
  int32_t left, right, parent, sibling;
};

inline uint64_t expand21(uint32_t v) {
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }

struct Builder {
  int n; const uint64_t* keys;
  int delta(int i, int j) const {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a != b) return clz64(a ^ b);
    return 64 + clz32((uint32_t)i ^ (uint32_t)j);
  }
};

}  // namespace

extern "C" {

// aabbs: n x [minx miny minz maxx maxy maxz] (fp32).  Outputs: keys_sorted[n], order[n]
// (sorted position -> primitive index), nodes[max(n-1,1)].  Returns the number of nodes.
int orc_lbvh_build(int n, const float* aabbs, uint64_t* keys_sorted, int32_t* order, void* nodes_out) {
  Node64* nodes = (Node64*)nodes_out;
  const float BIG = 3.0e38f;
  auto set_empty = [&](float* mn, float* mx) { for (int k = 0; k < 3; ++k) { mn[k] = BIG; mx[k] = -BIG; } };
  if (n <= 0) {
    Node64& nd = nodes[0]; for (int k = 0; k < 3; ++k) nd.lc[k] = nd.le[k] = nd.rc[k] = nd.re[k] = 0.f;
    nd.left = nd.right = ~0; nd.parent = nd.sibling = -1; return 1;
  }
  // 1. centroids + bounds, S
  std::vector<float> cen(3 * (size_t)n);
  float cmin[3] = {BIG, BIG, BIG}, cmax[3] = {-BIG, -BIG, -BIG}, S = 0.f;
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) {
    float mn = aabbs[6 * i + k], mx = aabbs[6 * i + 3 + k];
    float c = 0.5f * (mn + mx);
    cen[3 * i + k] = c;
    cmin[k] = std::min(cmin[k], c); cmax[k] = std::max(cmax[k], c);
    S = std::max(S, std::max(std::fabs(mn), std::fabs(mx)));
  }
  float pad = S * (1.0f / 2097152.0f);
  // 2. keys
  std::vector<uint64_t> keys(n);
  for (int i = 0; i < n; ++i) {
    uint32_t g[3];
    for (int k = 0; k < 3; ++k) {
      float ext = cmax[k] - cmin[k];
      float q = ext > 0.f ? (cen[3 * i + k] - cmin[k]) / ext : 0.f;
      float sc = q * 2097152.0f;
      uint32_t gi = (uint32_t)sc;
      g[k] = std::min(gi, 2097151u);
    }
    keys[i] = (expand21(g[0]) << 2) | (expand21(g[1]) << 1) | expand21(g[2]);
  }
  // 3. stable sort
  std::vector<int32_t> ord(n); std::iota(ord.begin(), ord.end(), 0);
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return keys[a] < keys[b]; });
  for (int i = 0; i < n; ++i) { keys_sorted[i] = keys[ord[i]]; order[i] = ord[i]; }
  auto leaf_box = [&](int pos, float* mn, float* mx) {
    const float* b = aabbs + 6 * (size_t)ord[pos];
    for (int k = 0; k < 3; ++k) { mn[k] = b[k]; mx[k] = b[3 + k]; }
  };
  if (n == 1) {
    Node64& nd = nodes[0];
    float mn[3], mx[3]; leaf_box(0, mn, mx);
    for (int k = 0; k < 3; ++k) { nd.lc[k] = 0.5f * (mn[k] + mx[k]); nd.le[k] = 0.5f * (mx[k] - mn[k]) + pad; }
    for (int k = 0; k < 3; ++k) { nd.rc[k] = nd.lc[k]; nd.re[k] = nd.le[k]; }   // same leaf, same box on both sides
    nd.left = nd.right = ~ord[0]; nd.parent = nd.sibling = -1; return 1;
  }
  // 4. Karras emit
  Builder B{n, keys_sorted};
  int nint = n - 1;
  std::vector<int> lo(nint), hi(nint);
  for (int i = 0; i < nint; ++i) {
    int d = (B.delta(i, i + 1) - B.delta(i, i - 1)) >= 0 ? 1 : -1;
    int dmin = B.delta(i, i - d);
    int lmax = 2;
    while (B.delta(i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) if (B.delta(i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = B.delta(i, j);
    int s = 0, t = l;
    do { t = (t + 1) >> 1; if (B.delta(i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
    int gamma = i + s * d + std::min(d, 0);
    int a = std::min(i, j), b = std::max(i, j);
    lo[i] = a; hi[i] = b;
    nodes[i].left = (a == gamma) ? ~ord[gamma] : gamma;
    nodes[i].right = (b == gamma + 1) ? ~ord[gamma + 1] : gamma + 1;
  }
  nodes[0].parent = -1; nodes[0].sibling = -1;
  for (int i = 0; i < nint; ++i) {
    int L = nodes[i].left, R = nodes[i].right;
    if (L >= 0) { nodes[L].parent = i; nodes[L].sibling = R; }
    if (R >= 0) { nodes[R].parent = i; nodes[R].sibling = L; }
  }
  // 5. boxes: unpadded union per node range [lo,hi] of sorted leaves, then pad on store.
  //    (min/max are exact, so any evaluation order gives identical bits.)
  auto range_box = [&](int a, int b, float* cen, float* ext) {
    float mn[3], mx[3]; set_empty(mn, mx);
    for (int p = a; p <= b; ++p) {
      float lm[3], lx[3]; leaf_box(p, lm, lx);
      for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], lm[k]); mx[k] = std::max(mx[k], lx[k]); }
    }
    for (int k = 0; k < 3; ++k) { cen[k] = 0.5f * (mn[k] + mx[k]); ext[k] = 0.5f * (mx[k] - mn[k]) + pad; }
  };
  // position of a leaf in sorted order is needed for leaf children: recover from gamma again
  for (int i = 0; i < nint; ++i) {
    int L = nodes[i].left, R = nodes[i].right;
    // the split position gamma: left covers [lo, gamma], right covers [gamma+1, hi]
    int gamma;
    if (L >= 0) gamma = hi[L]; else if (R >= 0) gamma = lo[R] - 1; else gamma = lo[i];
    range_box(lo[i], gamma, nodes[i].lc, nodes[i].le);
    range_box(gamma + 1, hi[i], nodes[i].rc, nodes[i].re);
  }
  return nint;
}

// Maximum number of internal nodes on any root-to-leaf path (the stackless trail needs <= 63).
int orc_lbvh_depth(int n_nodes, const void* nodes_in) {
  const Node64* nodes = (const Node64*)nodes_in;
  int best = 0;
  std::vector<std::pair<int, int>> st; st.push_back({0, 1});
  while (!st.empty()) {
    auto [i, d] = st.back(); st.pop_back();
    best = std::max(best, d);
    if (nodes[i].left >= 0) st.push_back({nodes[i].left, d + 1});
    if (nodes[i].right >= 0 && nodes[i].right != nodes[i].left) st.push_back({nodes[i].right, d + 1});
  }
  (void)n_nodes;
  return best;
}

}  // extern "C"
