// srt_oracle.cpp — CPU restatement of scheme-raytrace's per-sample radiance loop.
//
// *** TEST INFRASTRUCTURE ONLY ***  Nothing in the product path (scheme_raytrace_b200/)
// may include, link or call this file.  It is used by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs as the CHECKER and the CPU baseline.
//
// PARITY STATUS: pinned against outputs of the reference's own code.  The reference ships no tests or
// golden vectors and no Gauche runtime exists in the build container, so its .scm files are executed,
// unmodified, by the minimal interpreter oracle/minischeme.py; tests/golden/make_reference_golden.py
// freezes g:hit records of every constructor and of main.scm's scenes, camera rays, Perlin tables /
// noise / textures, the material helper functions and whole trace-all frames (random-real scripted with
// this file's Philox draws) into tests/golden/ref_*.json, and tests/test_reference_golden.py requires
// this file (f64, un-quantised scene) to reproduce them to 1e-12 (observed: bit-identical hit records;
// radiance sums to 1e-9; 8-bit image exact).  Caveat: the interpreter is not Gauche itself.  Parts the
// reference does not contain (metal / dielectric under `color`, isotropic, the hittable pdf, patches)
// remain "parity unpinned" and say so where they are defined.  Each function cites the file:line under
// /root/reference it follows; hand-derived known-answer vectors (KAT1-10 of SURVEY.md §8c and more) are
// in tests/test_oracle_kat.py.
//
// The whole engine is templated on the arithmetic type: <double> is the truth the GPU path is
// compared against (the reference computes in IEEE f64); <float> exists to flag rays on which the
// reference algorithm itself is unstable under fp32 rounding (used as a near-tie filter).
//
// Scene representation: a tree of nodes that mirrors the reference's closure-vector objects
// one to one (geometry.scm): LIST = hit-obj-list / make-scene / make-box / make-bvh-node
// (pass-through grouping), FLIP = flip-normals, TRANSLATE, ROTATE_Y, and the leaf primitives.
// RNG: the north star replaces srfi-27 MT19937 with counter-based Philox4x32-10 keyed by
// (pixel, sample, bounce); RNG streams are not a parity target, only distributions.
// The draw-slot table below is shared (by specification, not by code) with the CUDA path so the
// two can be compared path by path.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>
#include <limits>

namespace orc {

// ---------------------------------------------------------------------------------------------
// constant.scm:6
static const double MAX_FLOAT = 999999999999.0;
static const double PI = 3.141592653589793;  // math.const pi

enum NodeKind { N_SPHERE = 0, N_MOVING_SPHERE = 1, N_XY_RECT = 2, N_XZ_RECT = 3, N_YZ_RECT = 4,
                N_BEZIER = 5, N_CONSTANT_MEDIUM = 6, N_PATCH = 7, N_KLEIN = 8, N_FLIP = 16, N_LIST = 17, N_TRANSLATE = 18, N_ROTATE_Y = 19 };
enum MatKind { M_LAMBERTIAN = 0, M_METAL = 1, M_DIELECTRIC = 2, M_DIFFUSE_LIGHT = 3, M_ISOTROPIC = 4 };
enum TexKind { T_CONSTANT = 0, T_CHECKER = 1, T_NOISE = 2, T_MARBLE = 3, T_IMAGE = 4 };
enum SkyKind { SKY_GRADIENT = 0, SKY_BLACK = 1 };
// quirk bits (SURVEY §8a "Q" rows); REFERENCE = all on
enum Quirks { Q1_COSINE_X2 = 1, Q4_PERLIN_ALIAS = 2, Q6_SCATTER_TIME0 = 4, Q10_DIELECTRIC_UNNORM = 8,
              Q15_LOCAL_TRIPLE_EVAL = 16, Q_REFERENCE = 31 };

struct Tex { int kind; double rgb[3]; double scale; int even, odd; };
struct Mat { int kind; int tex; double param; };
struct Image { int nx, ny; std::vector<unsigned char> texels; };   // image-texture data, top row first
struct Node {
  int kind; int material; int leaf_id;
  int child_begin, child_count;
  double p[16];
};
struct Scene {
  std::vector<Image> images;
  std::vector<Tex> tex; std::vector<Mat> mat; std::vector<Node> nodes; std::vector<int> children;
  int root = -1;
  double cam[24];   // llc(3) horiz(3) vert(3) origin(3) w(3) u(3) v(3) lens t0 t1   camera.scm:33-61
  int sky = SKY_GRADIENT;
  double ranvec[256 * 3]; int perm_x[256], perm_y[256], perm_z[256];   // perlin.scm:33-36
  int exclude_leaf = -1;   // test hook: skip one leaf (second-best-hit query)
  std::vector<int> lights; // node ids of the light shapes sampled by the hittable pdf (S7)
  std::vector<int> leaf_node;   // leaf id -> node id
  std::vector<double> patches;  // 48 doubles per bicubic patch: P[i][j], i along u, j along v
};

// ---------------------------------------------------------------------------------------------
// vec.scm
template <class T> struct V3 { T x, y, z; };
template <class T> static inline V3<T> mk(T x, T y, T z) { return V3<T>{x, y, z}; }
template <class T> static inline V3<T> add(V3<T> a, V3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }   // vec.scm:20 (sum, pairwise f64vector-add)
template <class T> static inline V3<T> sub(V3<T> a, V3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }   // vec.scm:26
template <class T> static inline V3<T> mul(V3<T> a, V3<T> b) { return mk<T>(a.x * b.x, a.y * b.y, a.z * b.z); }   // vec.scm:35 prod
template <class T> static inline V3<T> scale(V3<T> a, T k) { return mk<T>(a.x * k, a.y * k, a.z * k); }            // vec.scm:41
template <class T> static inline T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }             // vec.scm:52
template <class T> static inline T length(V3<T> a) { return std::sqrt(dot(a, a)); }                               // vec.scm:54
template <class T> static inline V3<T> unit(V3<T> a) { T k = T(1) / length(a); return scale(a, k); }              // vec.scm:60-62
template <class T> static inline V3<T> cross(V3<T> a, V3<T> b) {                                                   // vec.scm:64-70
  return mk<T>(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
template <class T> static inline V3<T> ld3(const double* p) { return mk<T>(T(p[0]), T(p[1]), T(p[2])); }
template <class T> static inline T cmp(V3<T> a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ray.scm:8-54
template <class T> struct Ray { V3<T> o, d; T time; };
template <class T> struct HitRec { T t; V3<T> p, n; int mat; T u, v; int leaf; };
template <class T> static inline V3<T> point_at(const Ray<T>& r, T t) { return add(r.o, scale(r.d, t)); }  // ray.scm:23

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), the north-star RNG.  Known-answer vectors from the
// Random123 distribution are checked in tests/test_oracle_kat.py.
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// Draw-slot specification (DESIGN.md "RNG"): key = (pixel_index, seed); counter = (sample, bounce,
// block, 0); uniform = ((x >> 9) + 0.5) * 2^-23, a 23-bit grid that lies strictly inside (0,1) like
// srfi-27's random-real and is exactly representable in both f32 and f64 (a 24-bit grid is not:
// its top point rounds to 1.0 in fp32).
struct RngAddr { uint32_t seed, pixel, sample, bounce; };
template <class T> static inline void rng_block(const RngAddr& a, uint32_t block, T u[4]) {
  uint32_t ctr[4] = {a.sample, a.bounce, block, 0u}, key[2] = {a.pixel, a.seed}, o[4];
  philox4x32_10(ctr, key, o);
  for (int i = 0; i < 4; ++i) u[i] = T(((double)(o[i] >> 9) + 0.5) * (1.0 / 8388608.0));
}

// ---------------------------------------------------------------------------------------------
// util.scm — samplers.  Rejection loops take block j for iteration j.
template <class T> static V3<T> random_in_unit_sphere(const RngAddr& a, uint32_t first_block) {  // util.scm:9-15
  for (uint32_t j = 0;; ++j) {
    T u[4]; rng_block<T>(a, first_block + j, u);
    V3<T> p = sub(scale(mk<T>(u[0], u[1], u[2]), T(2)), mk<T>(1, 1, 1));
    if (dot(p, p) < T(1)) return p;
  }
}
template <class T> static V3<T> random_in_unit_disk(const RngAddr& a, uint32_t first_block) {    // util.scm:17-23
  for (uint32_t j = 0;; ++j) {
    T u[4]; rng_block<T>(a, first_block + j, u);   // two candidates per block
    for (int h = 0; h < 2; ++h) {
      V3<T> p = sub(scale(mk<T>(u[2 * h], u[2 * h + 1], T(0)), T(2)), mk<T>(1, 1, 0));
      if (dot(p, p) < T(1)) return p;
    }
  }
}
template <class T> static V3<T> random_cosine_direction(T r1, T r2, int quirks) {               // util.scm:37-44
  T z = std::sqrt(T(1) - r2);
  T phi = T(2) * T(PI) * r1;
  T k = (quirks & Q1_COSINE_X2) ? T(2) : T(1);                                                    // Q1: util.scm:42-43
  T x = std::cos(phi) * k * std::sqrt(r2);
  T y = std::sin(phi) * k * std::sqrt(r2);
  return mk<T>(x, y, z);
}

// onb.scm:27-36 — Q15.  `local` is a syntax-rules macro and its two-operand form mentions the operand
// `a` three times, so (local uvw (random-cosine-direction)) at material.scm:27 and pdf.scm:26 EVALUATES
// (random-cosine-direction) three times (6 random-real draws): x comes from the first direction, y from
// the second and z from the third.  (Found by executing the reference, tests/golden/ref_materials.json.)
// Draw slots: first evaluation = block 0 (x, y) as without the quirk; second = block 2 (x, y); third =
// block 2 (z, w), of which only r2 = w is used (its r1 only feeds the discarded x, y).
template <class T> static V3<T> cosine_direction_triple(const T* r6, int quirks) {     // r6: the six draws in call order
  V3<T> a = random_cosine_direction<T>(r6[0], r6[1], quirks), b = random_cosine_direction<T>(r6[2], r6[3], quirks),
        c = random_cosine_direction<T>(r6[4], r6[5], quirks);
  return mk<T>(a.x, b.y, c.z);
}
template <class T> static V3<T> cosine_direction_for_local(const RngAddr& addr, const T* block0, int quirks) {
  if (!(quirks & Q15_LOCAL_TRIPLE_EVAL)) return random_cosine_direction<T>(block0[0], block0[1], quirks);
  T w4[4]; rng_block<T>(addr, 2, w4);
  T r6[6] = {block0[0], block0[1], w4[0], w4[1], w4[2], w4[3]};
  return cosine_direction_triple<T>(r6, quirks);
}

template <class T> static V3<T> random_to_sphere(T radius, T distance_sq, T r1, T r2) {             // util.scm:46-54
  T z = T(1) + r2 * (std::sqrt(T(1) - (radius * radius) / distance_sq) - T(1));
  T phi = T(2) * T(PI) * r1;
  T x = std::cos(phi) * std::sqrt(T(1) - z * z);
  T y = std::sin(phi) * std::sqrt(T(1) - z * z);
  return mk<T>(x, y, z);
}

// onb.scm:8-16, 27-36
template <class T> struct Onb { V3<T> u, v, w; };
template <class T> static Onb<T> make_onb_from_w(V3<T> n) {
  V3<T> axis2 = unit(n);
  V3<T> a = (std::fabs(axis2.x) > T(0.9)) ? mk<T>(0, 1, 0) : mk<T>(1, 0, 0);
  V3<T> axis1 = unit(cross(axis2, a));
  V3<T> axis0 = cross(axis2, axis1);
  return Onb<T>{axis0, axis1, axis2};
}
template <class T> static V3<T> onb_local(const Onb<T>& o, V3<T> a) {
  return add(add(scale(o.u, a.x), scale(o.v, a.y)), scale(o.w, a.z));
}

// ---------------------------------------------------------------------------------------------
// perlin.scm:51-103
static inline long floor_exact(double x) { return (long)std::floor(x); }
template <class T> static T perlin_noise(const Scene& sc, V3<T> p, int quirks) {   // perlin.scm:69-90
  long i = floor_exact((double)p.x), j = floor_exact((double)p.y), k = floor_exact((double)p.z);
  T u = p.x - T(i), v = p.y - T(j), w = p.z - T(k);
  V3<T> c[2][2][2];
  for (int di = 0; di < 2; ++di) for (int dj = 0; dj < 2; ++dj) for (int dk = 0; dk < 2; ++dk) {
    // Q4 (perlin.scm:76): (make-vector 2 (make-vector 2 (make-vector 2))) shares ONE inner vector,
    // so after all writes c[i][j][k] holds the value written for di=1,dj=1,dk=k.
    int ei = (quirks & Q4_PERLIN_ALIAS) ? 1 : di, ej = (quirks & Q4_PERLIN_ALIAS) ? 1 : dj;
    int idx = sc.perm_x[(i + ei) & 255] ^ sc.perm_y[(j + ej) & 255] ^ sc.perm_z[(k + dk) & 255];
    c[di][dj][dk] = ld3<T>(&sc.ranvec[3 * idx]);
  }
  // perlin-interp perlin.scm:51-67
  T uu = u * u * (T(3) - T(2) * u), vv = v * v * (T(3) - T(2) * v), ww = w * w * (T(3) - T(2) * w);
  T acc = 0;
  for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int d = 0; d < 2; ++d) {
    T wa = T(a) * uu + T(1 - a) * (T(1) - uu);
    T wb = T(b) * vv + T(1 - b) * (T(1) - vv);
    T wd = T(d) * ww + T(1 - d) * (T(1) - ww);
    acc += wa * wb * wd * dot(mk<T>(u - T(a), v - T(b), w - T(d)), c[a][b][d]);
  }
  return acc;
}
template <class T> static T perlin_turb(const Scene& sc, V3<T> p, int quirks, int max_depth = 7) {  // perlin.scm:92-103
  T acc = 0, weight = 1;
  for (int d = 0; d < max_depth; ++d) {
    acc = acc + weight * perlin_noise(sc, p, quirks);
    p = scale(p, T(2));
    weight = weight * T(0.5);
  }
  return std::fabs(acc);
}

// texture.scm:9-50
template <class T> static V3<T> tex_value(const Scene& sc, int tex, T u, T v, V3<T> p, int quirks) {
  for (;;) {
    const Tex& t = sc.tex[tex];
    switch (t.kind) {
      case T_CONSTANT: return mk<T>(T(t.rgb[0]), T(t.rgb[1]), T(t.rgb[2]));                       // texture.scm:12-14
      case T_CHECKER: {                                                                           // texture.scm:16-23
        T sines = std::sin(T(10) * p.x) * std::sin(T(10) * p.y) * std::sin(T(10) * p.z);
        tex = (sines < T(0)) ? t.odd : t.even;
        continue;
      }
      case T_NOISE: {                                                                             // texture.scm:25-28
        T n = perlin_noise(sc, scale(p, T(t.scale)), quirks);
        return scale(mk<T>(1, 1, 1), n);
      }
      case T_MARBLE: {                                                                            // texture.scm:30-34
        T s = T(0.5) * (T(1) + std::sin(T(t.scale) * p.z + T(10) * perlin_turb(sc, p, quirks)));
        return scale(mk<T>(1, 1, 1), s);
      }
      case T_IMAGE: {                                                                             // texture.scm:36-50
        // i = u*nx; j = (1-v)*ny - 0.001; clamp both to [0, n-1] (:39-44).  Upstream then passes
        // the still-fractional i, j to vector-ref (an error in Gauche: the constructor is never
        // instantiated); the texel meant is the floor, as in the book this follows.  t.even = image.
        const Image& im = sc.images[t.even];
        T i = u * T(im.nx), j = (T(1) - v) * T(im.ny) - T(0.001);
        if (!(i >= T(0))) i = T(0);      // also sends NaN (Q5 sphere uv) to texel 0, on both sides
        if (!(j >= T(0))) j = T(0);
        if (i > T(im.nx - 1)) i = T(im.nx - 1);
        if (j > T(im.ny - 1)) j = T(im.ny - 1);
        const unsigned char* px = im.texels.data() + 3 * ((size_t)(int)i + (size_t)im.nx * (size_t)(int)j);
        return mk<T>(T(px[0]) / T(255), T(px[1]) / T(255), T(px[2]) / T(255));
      }
    }
    return mk<T>(0, 0, 0);
  }
}

// ---------------------------------------------------------------------------------------------
// geometry.scm:73-105 — AABB slab test, per-axis independent (Q11).  Only used through
// orc_aabb_hit (KATs); the oracle's closest-hit is the reference's linear hit-obj-list.
template <class T> static bool aabb_hit(V3<T> bmin, V3<T> bmax, const Ray<T>& r, T tmin, T tmax) {
  for (int ax = 0; ax < 3; ++ax) {
    T inv = T(1) / cmp(r.d, ax), o = cmp(r.o, ax);
    T a = (cmp(bmin, ax) - o) * inv, b = (cmp(bmax, ax) - o) * inv;
    T t0 = std::min(a, b), t1 = std::max(a, b);
    T lo = std::max(t0, tmin), hi = std::min(t1, tmax);
    if (hi <= lo) return false;
  }
  return true;
}

// geometry.scm:138-144 (Q5: atan(z,z), asin of the un-normalised world point)
template <class T> static void get_sphere_uv(V3<T> p, T& u, T& v) {
  T phi = std::atan2(p.z, p.z);
  T theta = std::asin(p.y);
  u = T(1) - (phi + T(PI)) / (T(2) * T(PI));
  v = (theta + T(PI) / T(2)) / T(PI);
}

// geometry.scm:146-175 (sphere), 177-215 (moving sphere: centre evaluated at ray time)
template <class T> static bool hit_sphere_at(V3<T> center, T radius, int material, int leaf, const Ray<T>& r,
                                             T tmin, T tmax, HitRec<T>& rec) {
  V3<T> oc = sub(r.o, center);
  T a = dot(r.d, r.d);
  T b = dot(oc, r.d);
  T c = dot(oc, oc) - radius * radius;
  T disc = b * b - a * c;
  if (disc <= T(0)) return false;
  T temp = (-b - std::sqrt(disc)) / a;
  if (!(tmin < temp && temp < tmax)) {
    temp = (-b + std::sqrt(disc)) / a;
    if (!(tmin < temp && temp < tmax)) return false;
  }
  rec.t = temp;
  rec.p = point_at(r, temp);
  rec.n = scale(sub(rec.p, center), T(1) / radius);
  rec.mat = material; rec.leaf = leaf;
  get_sphere_uv(rec.p, rec.u, rec.v);
  return true;
}

// geometry.scm:376-431.  axis = the thin axis (2: xy-rect, 1: xz-rect, 0: yz-rect).
// G5 documented divergence: a ray lying IN the plane gives t = NaN which the
// reference accepts; oracle and GPU both reject NaN t explicitly.
template <class T> static bool hit_rect(const Node& nd, int axis, const Ray<T>& r, T tmin, T tmax, HitRec<T>& rec) {
  T a0 = T(nd.p[0]), a1 = T(nd.p[1]), b0 = T(nd.p[2]), b1 = T(nd.p[3]), k = T(nd.p[4]);
  int ia, ib;   // in-plane axes in the reference's argument order
  if (axis == 2) { ia = 0; ib = 1; } else if (axis == 1) { ia = 0; ib = 2; } else { ia = 1; ib = 2; }
  T t = (k - cmp(r.o, axis)) / cmp(r.d, axis);
  if (t != t) return false;
  if (t < tmin || t > tmax) return false;
  T a = cmp(r.o, ia) + t * cmp(r.d, ia);
  T b = cmp(r.o, ib) + t * cmp(r.d, ib);
  if (a < a0 || a > a1 || b < b0 || b > b1) return false;
  rec.t = t; rec.p = point_at(r, t);
  rec.n = mk<T>(axis == 0 ? 1 : 0, axis == 1 ? 1 : 0, axis == 2 ? 1 : 0);
  rec.mat = nd.material; rec.leaf = nd.leaf_id;
  rec.u = (a - a0) / (a1 - a0);
  rec.v = (b - b0) / (b1 - b0);
  return true;
}

// ---------------------------------------------------------------------------------------------
// bezier.scm — cubic Bezier curve with circular width, recursive subdivision in ray space.
template <class T> struct Bez { V3<T> cp[4]; };
struct BezStats { int converge_calls; int max_depth; };

template <class T> static V3<T> bez_p(const Bez<T>& c, T t) {                       // bezier.scm:67-77
  T t2 = t * t, t3 = t2 * t, mt = T(1) - t, mt2 = mt * mt, mt3 = mt2 * mt;
  return add(add(add(scale(c.cp[0], mt3), scale(c.cp[1], T(3) * mt2 * t)), scale(c.cp[2], T(3) * mt * t2)), scale(c.cp[3], t3));
}
template <class T> static V3<T> idiv(V3<T> a, V3<T> b, T t) { return add(scale(a, T(1) - t), scale(b, t)); }  // bezier.scm:45-47
template <class T> static void bez_split(const Bez<T>& c, T t, Bez<T>& l, Bez<T>& r) {  // bezier.scm:78-87
  V3<T> sp = bez_p(c, t);
  V3<T> nbc = idiv(c.cp[1], c.cp[2], t);
  V3<T> lb = idiv(c.cp[0], c.cp[1], t);
  V3<T> lc = idiv(lb, nbc, t);
  V3<T> rc = idiv(c.cp[2], c.cp[3], t);
  V3<T> rb = idiv(nbc, rc, t);
  l.cp[0] = c.cp[0]; l.cp[1] = lb; l.cp[2] = lc; l.cp[3] = sp;
  r.cp[0] = sp; r.cp[1] = rb; r.cp[2] = rc; r.cp[3] = c.cp[3];
}
template <class T> static void bez_bbox(const Bez<T>& c, T width1, V3<T>& bmin, V3<T>& bmax) {  // bezier.scm:88-98
  T mn[3] = {T(MAX_FLOAT), T(MAX_FLOAT), T(MAX_FLOAT)}, mx[3] = {T(-MAX_FLOAT), T(-MAX_FLOAT), T(-MAX_FLOAT)};
  for (int i = 0; i < 4; ++i) for (int ax = 0; ax < 3; ++ax) {
    mn[ax] = std::min(cmp(c.cp[i], ax) - width1, mn[ax]);
    mx[ax] = std::max(cmp(c.cp[i], ax) + width1, mx[ax]);
  }
  bmin = mk<T>(mn[0], mn[1], mn[2]); bmax = mk<T>(mx[0], mx[1], mx[2]);
}
template <class T> static V3<T> bez_tan_vec(const Bez<T>& c, T t) {                 // bezier.scm:106-117
  T t2 = t * t;
  V3<T> a = c.cp[0], b = c.cp[1], cc = c.cp[2], d = c.cp[3];
  V3<T> coef_a = add(add(add(scale(b, T(3)), d), scale(cc, T(-3))), scale(a, T(-1)));
  V3<T> coef_b = scale(add(add(a, scale(b, T(-2))), cc), T(3));
  V3<T> coef_c = scale(sub(b, a), T(3));
  return unit(add(add(scale(coef_a, T(3) * t2), scale(coef_b, T(2) * t)), coef_c));
}
template <class T> static T dot2d(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + T(0) * T(0); }  // bezier.scm:57-59

// bezier.scm:13-43 get-projection-mat: M = Translate * Rotate, row-vector convention.
template <class T> static void projection_mat(const Ray<T>& r, T M[4][4]) {
  T ox = -r.o.x, oy = -(-r.o.z), oz = -r.o.y;
  V3<T> dir = unit(r.d);
  T lx = dir.x, ly = -dir.z, lz = dir.y;
  T d = std::sqrt(lx * lx + lz * lz);
  T R[4][4];
  if (d == T(0)) {
    T angle = (ly >= T(0)) ? -T(PI) / T(2) : T(PI) / T(2);
    T Rr[4][4] = {{1, 0, 0, 0}, {0, std::cos(angle), -std::sin(angle), 0}, {0, std::sin(angle), std::cos(angle), 0}, {0, 0, 0, 1}};
    std::memcpy(R, Rr, sizeof(R));
  } else {
    T Rr[4][4] = {{lz / d, (T(-1) * lx * ly) / d, lx, 0}, {0, d, ly, 0}, {(-lx) / d, (T(-1) * ly * lz) / d, lz, 0}, {0, 0, 0, 1}};
    std::memcpy(R, Rr, sizeof(R));
  }
  T Tm[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {ox, oy, oz, 1}};
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {   // gauche.array array-mul
    T s = 0; for (int k = 0; k < 4; ++k) s += Tm[i][k] * R[k][j];
    M[i][j] = s;
  }
}
template <class T> static V3<T> bez_transform_pt(V3<T> p, const T M[4][4]) {        // bezier.scm:49-55
  T row[4] = {p.x, -p.z, p.y, T(1)}, o[3];
  for (int j = 0; j < 3; ++j) { T s = 0; for (int k = 0; k < 4; ++k) s += row[k] * M[k][j]; o[j] = s; }
  return mk<T>(o[0], o[1], o[2]);
}

// bezier.scm:121-175
template <class T> static bool bez_converge(int depth, const Bez<T>& c, T v0, T vn, T t, T width1, T width2, T& t_out, BezStats* st) {
  if (st) st->converge_calls++;
  V3<T> bmin, bmax; bez_bbox(c, width1, bmin, bmax);
  if (bmin.z >= t || bmax.z <= T(0.000001) || bmin.x >= width1 || bmax.x <= -width1 || bmin.y >= width1 || bmax.y <= -width1)
    return false;
  if (depth < 0) {
    V3<T> dir = sub(c.cp[3], c.cp[0]);
    V3<T> dp0 = bez_tan_vec(c, T(0));
    if (dot2d(dir, dp0) < T(0)) dp0 = scale(dp0, T(-1));
    if (dot2d(dp0, scale(c.cp[0], T(-1))) < T(0)) return false;
    V3<T> dpn = bez_tan_vec(c, T(1));
    if (dot2d(dir, dpn) < T(0)) dpn = scale(dpn, T(-1));
    if (dot2d(dpn, c.cp[3]) < T(0)) return false;
    T w = dir.x * dir.x + dir.y * dir.y;
    if (w == T(0)) return false;
    w = (c.cp[0].x * dir.x + c.cp[0].y * dir.y) / (-w);
    w = std::min(std::max(w, T(0)), T(1));                 // clamp w 0 1
    T v = v0 * (T(1) - w) + vn * w;
    V3<T> p = bez_p(c, v);                                  // Q8: sub-curve evaluated at the GLOBAL parameter
    if ((p.x * p.x + p.y * p.y) >= width2 || p.z <= T(0.0001) || t < p.z) return false;
    t_out = p.z;
    return true;
  }
  T vm = (v0 + vn) / T(2);
  Bez<T> cl, cr; bez_split(c, T(0.5), cl, cr);
  T tl = 0, tr = 0;
  bool hl = bez_converge(depth - 1, cl, v0, vm, t, width1, width2, tl, st);
  bool hr = bez_converge(depth - 1, cr, vm, vn, t, width1, width2, tr, st);
  if (hl && tl < t) t = tl;
  if (hr && tr < t) t = tr;
  t_out = t;
  return hl || hr;
}
// bezier.scm:176-214
template <class T> static bool hit_bezier(const Node& nd, const Ray<T>& r, T tmin, T tmax, HitRec<T>& rec, BezStats* st) {
  Bez<T> c; for (int i = 0; i < 4; ++i) c.cp[i] = ld3<T>(&nd.p[3 * i]);
  T width = T(nd.p[12]);
  T width1 = width / T(2), width2 = width1 * width1, eps = width / T(20);     // bezier.scm:64-66
  const int n = 4;
  T M[4][4]; projection_mat(r, M);
  Bez<T> tr; for (int i = 0; i < 4; ++i) tr.cp[i] = bez_transform_pt(c.cp[i], M);
  T l0 = T(-MAX_FLOAT);
  for (int i = 0; i < n - 2; ++i) {
    T x = std::fabs(tr.cp[i].x + T(-2) * tr.cp[i + 1].x + tr.cp[i + 2].x);
    T y = std::fabs(tr.cp[i].y + T(-2) * tr.cp[i + 1].y + tr.cp[i + 2].y);
    l0 = std::max(std::max(x, y), l0);
  }
  T md = std::log((std::sqrt(T(2)) * T(n) * T(n - 1) * l0) / (T(8) * eps)) / std::log(T(4));
  int max_depth = (md == -std::numeric_limits<T>::infinity()) ? 0 : (int)std::ceil(md);
  if (st) st->max_depth = max_depth;
  T t = 0;
  bool hit = bez_converge(max_depth, tr, T(0), T(1), tmax, width1, width2, t, st);
  if (hit && tmin < t) {
    rec.t = t; rec.p = point_at(r, t);                    // Q9: raw (un-normalised) direction
    rec.n = scale(r.d, T(-1));
    rec.mat = nd.material; rec.leaf = nd.leaf_id; rec.u = 0; rec.v = 0;
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// Bicubic Bezier PATCH — north-star extension, ABSENT from the reference (which only has the
// curve above): PARITY UNPINNED.  Specification shared with the CUDA path (DESIGN.md "Patches"):
//   1. control net projected into ray space with bezier.scm's projection (ray = +z axis)
//   2. fixed-depth (2) quadtree subdivision by de Casteljau at 0.5 (u then v); a cell is culled
//      when the hull of its net misses the z axis or the (tmin, tbest) range
//   3. Newton on (S.x, S.y) = 0 from the cell centre, <= 8 iterations; accepted inside the cell
//      (+-1e-3); t = S.z / |d| (true parameter along the raw direction), strict both sides
//   4. normal = Su x Sv at the global (u, v) in world space, flipped to face the ray
template <class T> struct Net { V3<T> q[4][4]; };
template <class T> static void bern(T s, T b[4], T db[4]) {
  T m = T(1) - s;
  b[0] = m * m * m; b[1] = T(3) * s * m * m; b[2] = T(3) * s * s * m; b[3] = s * s * s;
  db[0] = T(-3) * m * m; db[1] = T(3) * m * m - T(6) * s * m; db[2] = T(6) * s * m - T(3) * s * s; db[3] = T(3) * s * s;
}
template <class T> static void patch_eval(const Net<T>& N, T s, T t, V3<T>& S, V3<T>& Su, V3<T>& Sv) {
  T bs[4], dbs[4], bt[4], dbt[4]; bern(s, bs, dbs); bern(t, bt, dbt);
  S = Su = Sv = mk<T>(0, 0, 0);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
    S = add(S, scale(N.q[i][j], bs[i] * bt[j]));
    Su = add(Su, scale(N.q[i][j], dbs[i] * bt[j]));
    Sv = add(Sv, scale(N.q[i][j], bs[i] * dbt[j]));
  }
}
template <class T> static void cubic_split(const V3<T> c[4], V3<T> l[4], V3<T> r[4]) {   // de Casteljau at 0.5
  V3<T> ab = scale(add(c[0], c[1]), T(0.5)), bc = scale(add(c[1], c[2]), T(0.5)), cd = scale(add(c[2], c[3]), T(0.5));
  V3<T> abc = scale(add(ab, bc), T(0.5)), bcd = scale(add(bc, cd), T(0.5)), m = scale(add(abc, bcd), T(0.5));
  l[0] = c[0]; l[1] = ab; l[2] = abc; l[3] = m; r[0] = m; r[1] = bcd; r[2] = cd; r[3] = c[3];
}
template <class T> static void net_split_u(const Net<T>& N, Net<T>& lo, Net<T>& hi) {
  for (int j = 0; j < 4; ++j) { V3<T> c[4] = {N.q[0][j], N.q[1][j], N.q[2][j], N.q[3][j]}, l[4], r[4]; cubic_split(c, l, r);
    for (int i = 0; i < 4; ++i) { lo.q[i][j] = l[i]; hi.q[i][j] = r[i]; } }
}
template <class T> static void net_split_v(const Net<T>& N, Net<T>& lo, Net<T>& hi) {
  for (int i = 0; i < 4; ++i) { V3<T> l[4], r[4]; cubic_split(N.q[i], l, r);
    for (int j = 0; j < 4; ++j) { lo.q[i][j] = l[j]; hi.q[i][j] = r[j]; } }
}
template <class T> static bool net_cull(const Net<T>& N, T zmin, T zmax) {
  T mnx = N.q[0][0].x, mxx = mnx, mny = N.q[0][0].y, mxy = mny, mnz = N.q[0][0].z, mxz = mnz;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { const V3<T>& q = N.q[i][j];
    mnx = std::min(mnx, q.x); mxx = std::max(mxx, q.x); mny = std::min(mny, q.y); mxy = std::max(mxy, q.y); mnz = std::min(mnz, q.z); mxz = std::max(mxz, q.z); }
  return mnx > T(0) || mxx < T(0) || mny > T(0) || mxy < T(0) || mxz < zmin || mnz > zmax;
}
template <class T> static void patch_recurse(const Net<T>& N, int depth, T u0, T v0, T size, T zmin, T& zbest, T& ub, T& vb, bool& any) {
  if (net_cull(N, zmin, zbest)) return;
  if (depth == 0) {
    T s = T(0.5), t = T(0.5); bool conv = false; V3<T> S, Su, Sv;
    for (int it = 0; it < 8; ++it) {
      patch_eval(N, s, t, S, Su, Sv);
      T det = Su.x * Sv.y - Sv.x * Su.y;
      if (!(std::fabs(det) > T(1e-30))) break;
      T ds = (-S.x * Sv.y + S.y * Sv.x) / det, dt = (-Su.x * S.y + Su.y * S.x) / det;
      s += ds; t += dt;
      if (!(std::fabs(s) < T(4)) || !(std::fabs(t) < T(4))) break;
      if (std::max(std::fabs(ds), std::fabs(dt)) < T(1e-5)) { conv = true; break; }
    }
    if (!conv || s < T(-1e-3) || s > T(1) + T(1e-3) || t < T(-1e-3) || t > T(1) + T(1e-3)) return;
    s = std::min(std::max(s, T(0)), T(1)); t = std::min(std::max(t, T(0)), T(1));
    patch_eval(N, s, t, S, Su, Sv);
    if (S.z > zmin && S.z < zbest) { zbest = S.z; ub = u0 + s * size; vb = v0 + t * size; any = true; }
    return;
  }
  Net<T> lo, hi, a, b; net_split_u(N, lo, hi);
  T h = size * T(0.5);
  net_split_v(lo, a, b); patch_recurse(a, depth - 1, u0, v0, h, zmin, zbest, ub, vb, any); patch_recurse(b, depth - 1, u0, v0 + h, h, zmin, zbest, ub, vb, any);
  net_split_v(hi, a, b); patch_recurse(a, depth - 1, u0 + h, v0, h, zmin, zbest, ub, vb, any); patch_recurse(b, depth - 1, u0 + h, v0 + h, h, zmin, zbest, ub, vb, any);
}
template <class T> static bool hit_patch(const Scene& sc, const Node& nd, const Ray<T>& r, T tmin, T tmax, HitRec<T>& rec) {
  const double* cp = &sc.patches[48 * (size_t)nd.p[0]];
  T M[4][4]; projection_mat(r, M);
  Net<T> N, W;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { W.q[i][j] = ld3<T>(cp + 3 * (4 * i + j)); N.q[i][j] = bez_transform_pt(W.q[i][j], M); }
  T len = length(r.d);
  T zbest = tmax * len, ub = 0, vb = 0; bool any = false;
  patch_recurse<T>(N, 2, T(0), T(0), T(1), tmin * len, zbest, ub, vb, any);
  if (!any) return false;
  rec.t = zbest / len; rec.p = point_at(r, rec.t);
  V3<T> S, Su, Sv; patch_eval(W, ub, vb, S, Su, Sv);
  V3<T> n = unit(cross(Su, Sv));
  if (dot(n, r.d) > T(0)) n = scale(n, T(-1));
  rec.n = n; rec.mat = nd.material; rec.leaf = nd.leaf_id; rec.u = ub; rec.v = vb;
  return true;
}

// ---------------------------------------------------------------------------------------------
// geometry.scm:590-664 — Klein / IIS fractal: a distance field made by inverting the point in six
// spheres (at most 10 times), sphere-traced for at most 100 steps; normal by central differences.
static const double KLEIN_POS[6][3] = {{300, 300, 0}, {300, -300, 0}, {-300, 300, 0}, {-300, -300, 0}, {0, 0, 424.26}, {0, 0, -424.26}};   // :590-595
template <class T> static T klein_dist(V3<T> center, V3<T> pos) {              // dist-func geometry.scm:602-624
  const T R = T(300), R2 = R * R, KR = T(125);                                 // :596-598
  pos = sub(pos, center);
  T dr = 1;
  for (int iter = 0;;) {
    if (iter >= 10) return T(0.7) * ((length(pos) - KR) / std::fabs(dr));      // +max-klein-loop+ :600
    bool inverted = false;
    for (int k = 0; k < 6; ++k) {
      V3<T> sp = ld3<T>(KLEIN_POS[k]);
      if (length(sub(pos, sp)) < R) {
        V3<T> diff = sub(pos, sp);
        dr = dr * (R2 / dot(diff, diff));
        pos = add(scale(scale(diff, R2), T(1) / (length(diff) * length(diff))), sp);
        ++iter; inverted = true;
        break;
      }
    }
    if (!inverted) return T(0.7) * ((length(pos) - KR) / std::fabs(dr));
  }
}
template <class T> static V3<T> klein_normal(V3<T> center, V3<T> p) {          // get-normal geometry.scm:626-632
  const T e = T(0.01);
  return unit(mk<T>(klein_dist(center, add(p, mk<T>(e, 0, 0))) - klein_dist(center, sub(p, mk<T>(e, 0, 0))),
                    klein_dist(center, add(p, mk<T>(0, e, 0))) - klein_dist(center, sub(p, mk<T>(0, e, 0))),
                    klein_dist(center, add(p, mk<T>(0, 0, e))) - klein_dist(center, sub(p, mk<T>(0, 0, e)))));
}
template <class T> static bool hit_klein(const Node& nd, const Ray<T>& r, T tmin, T tmax, HitRec<T>& rec) {   // make-klein :644-664
  V3<T> center = ld3<T>(nd.p);
  V3<T> ray_pos = r.o; T ray_length = 0;
  for (int iter = 0; iter < 100; ++iter) {                                     // +max-marching-loop+ :635
    T dist = klein_dist(center, ray_pos);
    ray_length = ray_length + dist;
    ray_pos = add(r.o, scale(r.d, ray_length));
    if (dist < T(0.001) && tmin < ray_length && ray_length < tmax) {
      rec.t = ray_length; rec.p = point_at(r, ray_length); rec.n = klein_normal(center, ray_pos);
      rec.mat = nd.material; rec.leaf = nd.leaf_id; rec.u = 0; rec.v = 0;
      return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// geometry.scm:14-15 — (hit obj r t-min t-max), dispatch over the object kinds.
template <class T> static bool hit_node(const Scene& sc, int id, const Ray<T>& r, T tmin, T tmax, HitRec<T>& rec, const RngAddr* rng = nullptr) {
  const Node& nd = sc.nodes[id];
  switch (nd.kind) {
    case N_LIST: {                                         // geometry.scm:33-50 hit-obj-list
      bool any = false; T closest = tmax; HitRec<T> tmp;
      for (int i = 0; i < nd.child_count; ++i) {
        if (hit_node(sc, sc.children[nd.child_begin + i], r, tmin, closest, tmp, rng)) { any = true; closest = tmp.t; rec = tmp; }
      }
      return any;
    }
    case N_SPHERE:
      if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false;
      return hit_sphere_at(ld3<T>(nd.p), T(nd.p[3]), nd.material, nd.leaf_id, r, tmin, tmax, rec);
    case N_MOVING_SPHERE: {                                // geometry.scm:178-184
      if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false;
      V3<T> c0 = ld3<T>(nd.p), c1 = ld3<T>(nd.p + 4);
      T time0 = T(nd.p[7]), time1 = T(nd.p[8]);
      V3<T> cc = add(c0, scale(sub(c1, c0), (r.time - time0) / (time1 - time0)));
      return hit_sphere_at(cc, T(nd.p[3]), nd.material, nd.leaf_id, r, tmin, tmax, rec);
    }
    case N_XY_RECT: if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false; return hit_rect(nd, 2, r, tmin, tmax, rec);
    case N_XZ_RECT: if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false; return hit_rect(nd, 1, r, tmin, tmax, rec);
    case N_YZ_RECT: if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false; return hit_rect(nd, 0, r, tmin, tmax, rec);
    case N_BEZIER: if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false; return hit_bezier(nd, r, tmin, tmax, rec, (BezStats*)nullptr);
    case N_KLEIN: if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false; return hit_klein(nd, r, tmin, tmax, rec);
    case N_PATCH: if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false; return hit_patch(sc, nd, r, tmin, tmax, rec);
    case N_CONSTANT_MEDIUM: {                              // geometry.scm:545-578
      // phase function = lambertian (isotropic is commented out upstream, geometry.scm:546).
      // The free-flight draw replaces (random-real) by block 16 + leaf id of the ray's
      // (pixel, sample, bounce) Philox stream, so the result does not depend on visit order.
      if (sc.exclude_leaf >= 0 && nd.leaf_id == sc.exclude_leaf) return false;
      int obj = sc.children[nd.child_begin];
      HitRec<T> rec1, rec2;
      if (!hit_node(sc, obj, r, T(-MAX_FLOAT), T(MAX_FLOAT), rec1)) return false;
      if (!hit_node(sc, obj, r, rec1.t + T(0.0001), T(MAX_FLOAT), rec2)) return false;
      T t1 = (rec1.t < tmin) ? tmin : rec1.t;
      T t2 = (rec2.t > tmax) ? tmax : rec2.t;
      if (t1 >= t2) return false;
      if (t1 < T(0)) t1 = 0;
      T len = length(r.d);
      T distance_inside = (t2 - t1) * len;
      T xi = T(0.5);
      if (rng) { T u4[4]; rng_block<T>(*rng, 16u + (uint32_t)nd.leaf_id, u4); xi = u4[0]; }
      T hit_distance = (-(T(1) / T(nd.p[0]))) * std::log(xi);
      if (!(hit_distance < distance_inside)) return false;
      T nt = t1 + hit_distance / len;
      rec.t = nt; rec.p = point_at(r, nt); rec.n = mk<T>(1, 0, 0);
      rec.mat = nd.material; rec.leaf = nd.leaf_id; rec.u = 0; rec.v = 0;
      return true;
    }
    case N_FLIP: {                                         // geometry.scm:433-442
      if (!hit_node(sc, sc.children[nd.child_begin], r, tmin, tmax, rec, rng)) return false;
      rec.n = scale(rec.n, T(-1));
      return true;
    }
    case N_TRANSLATE: {                                    // geometry.scm:465-481
      V3<T> off = ld3<T>(nd.p);
      Ray<T> moved{sub(r.o, off), r.d, r.time};
      if (!hit_node(sc, sc.children[nd.child_begin], moved, tmin, tmax, rec, rng)) return false;
      rec.p = add(rec.p, off);
      return true;
    }
    case N_ROTATE_Y: {                                     // geometry.scm:483-543 (p[0]=sin, p[1]=cos)
      T s = T(nd.p[0]), c = T(nd.p[1]);
      Ray<T> rot{mk<T>(c * r.o.x - s * r.o.z, r.o.y, s * r.o.x + c * r.o.z),
                 mk<T>(c * r.d.x - s * r.d.z, r.d.y, s * r.d.x + c * r.d.z), r.time};
      int ch = sc.children[nd.child_begin];
      if (!hit_node(sc, ch, rot, tmin, tmax, rec, rng)) return false;
      V3<T> p = mk<T>(c * rec.p.x + s * rec.p.z, rec.p.y, (-s) * rec.p.x + c * rec.p.z);
      V3<T> n = mk<T>(c * rec.n.x + s * rec.n.z, rec.n.y, (-s) * rec.n.x + c * rec.n.z);
      rec.p = p; rec.n = n;
      if (sc.nodes[ch].material >= 0) rec.mat = sc.nodes[ch].material;   // (material obj), geometry.scm:537
      return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// material.scm
template <class T> static V3<T> reflect(V3<T> v, V3<T> n) { return sub(v, scale(n, T(2) * dot(v, n))); }   // material.scm:41-43
template <class T> static bool refract(V3<T> v, V3<T> n, T ni_over_nt, V3<T>& out, int quirks) {           // material.scm:59-67
  V3<T> uv = unit(v);
  T dt = dot(uv, n);
  T disc = T(1) - ni_over_nt * ni_over_nt * (T(1) - dt * dt);
  if (disc > T(0)) {
    V3<T> vv = (quirks & Q10_DIELECTRIC_UNNORM) ? v : uv;   // Q10: tangential term uses the raw v
    out = sub(scale(sub(vv, scale(n, dt)), ni_over_nt), scale(n, std::sqrt(disc)));
    return true;
  }
  return false;
}
template <class T> static T schlick(T cosine, T ref_idx) {                                                  // material.scm:69-74
  T r0 = (T(1) - ref_idx) / (T(1) + ref_idx);
  r0 = r0 * r0;
  return r0 + (T(1) - r0) * std::pow(T(1) - cosine, T(5));
}

// material.scm:45-53 — metal: reflect the unit direction, add fuzz * random-in-unit-sphere (rejection iteration j draws
// block j of the bounce), valid iff the result leaves the surface.
template <class T> static bool metal_scatter(V3<T> d, V3<T> n, T fuzz, const RngAddr& addr, V3<T>& out) {
  V3<T> reflected = reflect(unit(d), n);
  V3<T> fuzzv = random_in_unit_sphere<T>(addr, 0);
  out = add(reflected, scale(fuzzv, fuzz));
  return dot(out, n) > T(0);
}
// material.scm:76-98 — dielectric: xi < reflect-prob picks the reflected direction (Q10: raw d throughout).
template <class T> static V3<T> dielectric_scatter(V3<T> d, V3<T> n, T ref_idx, T xi, int quirks) {
  V3<T> din = (quirks & Q10_DIELECTRIC_UNNORM) ? d : unit(d);
  V3<T> reflected = reflect(din, n);
  T dd = dot(din, n);
  V3<T> outward = (dd > T(0)) ? scale(n, T(-1)) : n;
  T ni_over_nt = (dd > T(0)) ? ref_idx : T(1) / ref_idx;
  T cosine = (dd > T(0)) ? (dd * ref_idx) / length(din) : (-dd) / length(din);
  V3<T> refracted;
  bool ok = refract(din, outward, ni_over_nt, refracted, quirks);
  T reflect_prob = ok ? schlick(cosine, ref_idx) : T(1);
  return (xi < reflect_prob) ? reflected : refracted;
}

// main.scm:91-98
template <class T> static V3<T> sky_value(const Scene& sc, const Ray<T>& r) {
  if (sc.sky == SKY_BLACK) return mk<T>(0, 0, 0);
  V3<T> ud = unit(r.d);
  T t = T(0.5) * (T(1.0) + ud.y);
  return add(scale(mk<T>(1, 1, 1), T(1) - t), scale(mk<T>(T(0.5), T(0.7), T(1.0)), t));
}

// ---------------------------------------------------------------------------------------------
// pdf.scm.  make-cosine-pdf (:18-26) and make-mixture-pdf (:34-41) are pinned by source;
// make-hitable-pdf (:28-32) delegates to g:pdf-value / g:random, which do NOT exist upstream
//  — PARITY UNPINNED; the per-shape pdf-value / random below follow "Ray Tracing:
// The Rest of Your Life" (rect: dist^2 / (|cos| area); sphere: 1 / (2 pi (1 - cos_theta_max)),
// random = onb.local(random-to-sphere), util.scm:46-54).  Lights must be un-instanced rects or
// spheres.
template <class T> static T cosine_pdf_value(V3<T> w_unit, V3<T> direction) {                    // pdf.scm:19-23
  T cosine = dot(unit(direction), w_unit);
  return cosine > T(0) ? cosine / T(PI) : T(0);
}
template <class T> static T light_pdf_value(const Scene& sc, int node, V3<T> o, V3<T> v) {
  const Node& nd = sc.nodes[node];
  Ray<T> r{o, v, T(0)}; HitRec<T> rec;
  if (nd.kind == N_SPHERE) {
    if (!hit_sphere_at(ld3<T>(nd.p), T(nd.p[3]), nd.material, nd.leaf_id, r, T(0.001), T(MAX_FLOAT), rec)) return T(0);
    V3<T> dc = sub(ld3<T>(nd.p), o);
    T cos_theta_max = std::sqrt(T(1) - T(nd.p[3]) * T(nd.p[3]) / dot(dc, dc));
    return T(1) / (T(2) * T(PI) * (T(1) - cos_theta_max));
  }
  int axis = nd.kind == N_XY_RECT ? 2 : (nd.kind == N_XZ_RECT ? 1 : 0);
  if (!hit_rect(nd, axis, r, T(0.001), T(MAX_FLOAT), rec)) return T(0);
  T area = (T(nd.p[1]) - T(nd.p[0])) * (T(nd.p[3]) - T(nd.p[2]));
  T dist2 = rec.t * rec.t * dot(v, v);
  T cosine = std::fabs(cmp(v, axis)) / length(v);
  return dist2 / (cosine * area);
}
template <class T> static V3<T> light_random(const Scene& sc, int node, V3<T> o, T xa, T xb) {
  const Node& nd = sc.nodes[node];
  if (nd.kind == N_SPHERE) {
    V3<T> dc = sub(ld3<T>(nd.p), o);
    Onb<T> uvw = make_onb_from_w(dc);
    return onb_local(uvw, random_to_sphere<T>(T(nd.p[3]), dot(dc, dc), xa, xb));
  }
  T a = T(nd.p[0]) + xa * (T(nd.p[1]) - T(nd.p[0])), b = T(nd.p[2]) + xb * (T(nd.p[3]) - T(nd.p[2])), k = T(nd.p[4]);
  V3<T> pt = nd.kind == N_XY_RECT ? mk<T>(a, b, k) : (nd.kind == N_XZ_RECT ? mk<T>(a, k, b) : mk<T>(k, a, b));
  return sub(pt, o);
}
template <class T> static T lights_pdf_value(const Scene& sc, V3<T> o, V3<T> v) {   // hittable_list::pdf_value: plain average
  T sum = 0;
  for (size_t j = 0; j < sc.lights.size(); ++j) sum += light_pdf_value<T>(sc, sc.lights[j], o, v);
  return sum / T(sc.lights.size());
}

enum Estimator { EST_REFERENCE = 0, EST_MIXTURE = 1 };
struct RenderCtx { int max_depth; int quirks; uint32_t seed; std::atomic<uint64_t>* nrays; int estimator; };

// main.scm:100-121 — recursive radiance estimator.  RNG: scatter at depth k draws from bounce k+1.
template <class T> static V3<T> color(const Scene& sc, const Ray<T>& r, int depth, RngAddr addr, const RenderCtx& cx, uint64_t& nrays) {
  HitRec<T> rec;
  nrays++;
  addr.bounce = (uint32_t)depth + 1;
  if (!hit_node(sc, sc.root, r, T(0.001), T(MAX_FLOAT), rec, &addr)) return sky_value(sc, r);
  const Mat& m = sc.mat[rec.mat];
  T time0 = (cx.quirks & Q6_SCATTER_TIME0) ? T(0) : r.time;         // Q6: make-ray forces time 0
  switch (m.kind) {
    case M_LAMBERTIAN: {                                            // material.scm:24-39
      Onb<T> uvw = make_onb_from_w(rec.n);
      T u4[4]; rng_block<T>(addr, 0, u4);
      if (cx.estimator == EST_MIXTURE && !sc.lights.empty()) {
        // Rest-of-Life estimator: mixture(hittable(lights), cosine) — pdf.scm:34-41.  Draw slots:
        // block 0 = (r1, r2, xi_choice, xi_light), block 1 = (xa, xb).
        T v4[4]; rng_block<T>(addr, 1, v4);
        V3<T> dir;
        if (u4[2] < T(0.5)) {                                       // (generate p0): hittable pdf
          int li = std::min((int)(u4[3] * T(sc.lights.size())), (int)sc.lights.size() - 1);
          dir = light_random<T>(sc, sc.lights[li], rec.p, v4[0], v4[1]);
        } else {
          dir = onb_local(uvw, cosine_direction_for_local<T>(addr, u4, cx.quirks));   // (generate p1): cosine pdf
        }
        Ray<T> scattered{rec.p, dir, time0};
        T pdf_val = T(0.5) * lights_pdf_value<T>(sc, rec.p, dir) + T(0.5) * cosine_pdf_value<T>(uvw.w, dir);
        T cosine = dot(rec.n, unit(dir)); if (cosine < T(0)) cosine = 0;
        T spdf = cosine / T(PI);
        V3<T> atten = tex_value<T>(sc, m.tex, T(0), T(0), rec.p, cx.quirks);
        if (depth < cx.max_depth && spdf > T(0) && pdf_val > T(0))  // zero-weight paths end here (0 * color = 0)
          return scale(mul(scale(atten, spdf), color(sc, scattered, depth + 1, addr, cx, nrays)), T(1) / pdf_val);
        return mk<T>(0, 0, 0);
      }
      V3<T> target = onb_local(uvw, cosine_direction_for_local<T>(addr, u4, cx.quirks));
      Ray<T> scattered{rec.p, unit(target), time0};
      V3<T> atten = tex_value<T>(sc, m.tex, T(0), T(0), rec.p, cx.quirks);
      T pdf = dot(uvw.w, scattered.d) / T(PI);
      T cosine = dot(rec.n, unit(scattered.d));                     // scattering-pdf material.scm:33-36
      if (cosine < T(0)) cosine = 0;
      T spdf = cosine / T(PI);
      V3<T> emitted = mk<T>(0, 0, 0);
      if (depth < cx.max_depth)                                     // main.scm:112-119
        return add(emitted, scale(mul(scale(atten, spdf), color(sc, scattered, depth + 1, addr, cx, nrays)), T(1) / pdf));
      return emitted;
    }
    case M_METAL: {                                                 // material.scm:45-57
      // HEAD's `color` cannot run metal/dielectric (3 values into a 4-value receive, SURVEY §8a
      // M2/M3); intended Weekend semantics: specular, weight = attenuation, emitted = 0.
      // IMAGE-LEVEL PARITY UNPINNED; the scatter arithmetic below is pinned by source.
      V3<T> sdir;
      bool valid = metal_scatter<T>(r.d, rec.n, T(m.param), addr, sdir);
      Ray<T> scattered{rec.p, sdir, time0};
      V3<T> atten = tex_value<T>(sc, m.tex, T(0), T(0), rec.p, cx.quirks);
      if (depth < cx.max_depth && valid) return mul(atten, color(sc, scattered, depth + 1, addr, cx, nrays));
      return mk<T>(0, 0, 0);
    }
    case M_DIELECTRIC: {                                            // material.scm:76-101
      T u4[4]; rng_block<T>(addr, 0, u4);
      Ray<T> scattered{rec.p, dielectric_scatter<T>(r.d, rec.n, T(m.param), u4[0], cx.quirks), time0};
      if (depth < cx.max_depth) return color(sc, scattered, depth + 1, addr, cx, nrays);   // attenuation (1,1,1)
      return mk<T>(0, 0, 0);
    }
    case M_DIFFUSE_LIGHT: {                                         // material.scm:103-111
      if (dot(rec.n, r.d) < T(0.0)) return tex_value<T>(sc, m.tex, rec.u, rec.v, rec.p, cx.quirks);
      return mk<T>(0, 0, 0);
    }
    case M_ISOTROPIC: {
      // ABSENT from the reference (geometry.scm:546 comments it out) — PARITY UNPINNED.
      // Book ("The Next Week") semantics: scattered dir = random-in-unit-sphere, weight = albedo.
      V3<T> dirv = random_in_unit_sphere<T>(addr, 0);
      Ray<T> scattered{rec.p, dirv, time0};
      V3<T> atten = tex_value<T>(sc, m.tex, rec.u, rec.v, rec.p, cx.quirks);
      if (depth < cx.max_depth) return mul(atten, color(sc, scattered, depth + 1, addr, cx, nrays));
      return mk<T>(0, 0, 0);
    }
  }
  return mk<T>(0, 0, 0);
}

// camera.scm:80-92.  Draw slots (bounce 0): block 0 = [xi_u, xi_v, xi_time, -]; lens disk
// rejection from block 1 on (skipped when lens-radius = 0: the product is exactly zero).
template <class T> static Ray<T> get_ray(const Scene& sc, T s, T t, T xi_time, const RngAddr& addr) {
  const double* c = sc.cam;
  V3<T> llc = ld3<T>(c), horiz = ld3<T>(c + 3), vert = ld3<T>(c + 6), origin = ld3<T>(c + 9), cu = ld3<T>(c + 15), cv = ld3<T>(c + 18);
  T lens = T(c[21]), time0 = T(c[22]), time1 = T(c[23]);
  V3<T> rd = mk<T>(0, 0, 0);
  if (lens != T(0)) rd = scale(random_in_unit_disk<T>(addr, 1), lens);
  V3<T> offset = add(scale(cu, rd.x), scale(cv, rd.y));
  T time = time0 + xi_time * (time1 - time0);
  Ray<T> r;
  r.o = add(origin, offset);
  r.d = sub(sub(add(add(llc, scale(horiz, s)), scale(vert, t)), origin), offset);
  r.time = time;
  return r;
}

// main.scm:471-491 trace-all, one sample of one pixel (without the gamma/quantise tail).
template <class T> static V3<T> sample_pixel(const Scene& sc, int x, int y, int w, int h, uint32_t sample, const RenderCtx& cx, uint64_t& nrays) {
  RngAddr addr{cx.seed, (uint32_t)(y * w + x), sample, 0u};
  T u4[4]; rng_block<T>(addr, 0, u4);
  T u = (T(x) + u4[0]) / T(w);
  T v = (T(y) + u4[1]) / T(h);
  Ray<T> r = get_ray<T>(sc, u, v, u4[2], addr);
  return color<T>(sc, r, 0, addr, cx, nrays);
}

template <class T> static void render(const Scene& sc, int w, int h, int spp_begin, int spp_end, int max_depth, uint32_t seed,
                                      int quirks, int nthreads, double* rgb_sum, uint64_t* nrays_out, int estimator) {
  std::atomic<int> next_row{0};
  std::atomic<uint64_t> total{0};
  RenderCtx cx{max_depth, quirks, seed, &total, estimator};
  auto work = [&]() {
    uint64_t nr = 0;
    for (;;) {
      int y = next_row.fetch_add(1);
      if (y >= h) break;
      for (int x = 0; x < w; ++x) {
        double acc[3] = {rgb_sum[3 * (y * w + x)], rgb_sum[3 * (y * w + x) + 1], rgb_sum[3 * (y * w + x) + 2]};
        for (int s = spp_begin; s < spp_end; ++s) {
          V3<T> c = sample_pixel<T>(sc, x, y, w, h, (uint32_t)s, cx, nr);
          acc[0] += (double)c.x; acc[1] += (double)c.y; acc[2] += (double)c.z;   // main.scm:480 running sum
        }
        rgb_sum[3 * (y * w + x)] = acc[0]; rgb_sum[3 * (y * w + x) + 1] = acc[1]; rgb_sum[3 * (y * w + x) + 2] = acc[2];
      }
    }
    total += nr;
  };
  if (nthreads <= 1) work();
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; ++i) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  if (nrays_out) *nrays_out = total.load();
}

}  // namespace orc

// =================================================================================================
// C API (ctypes).  All arrays are caller-owned.
using namespace orc;
extern "C" {

void* orc_create() { Scene* s = new Scene(); std::memset(s->cam, 0, sizeof(s->cam)); std::memset(s->ranvec, 0, sizeof(s->ranvec));
  for (int i = 0; i < 256; ++i) s->perm_x[i] = s->perm_y[i] = s->perm_z[i] = i; return s; }
void orc_destroy(void* h) { delete (Scene*)h; }

int orc_add_texture(void* h, int kind, double r, double g, double b, double scale, int even, int odd) {
  Scene* s = (Scene*)h; s->tex.push_back(Tex{kind, {r, g, b}, scale, even, odd}); return (int)s->tex.size() - 1; }
int orc_add_image(void* h, const unsigned char* texels, int nx, int ny) {
  Scene* s = (Scene*)h; s->images.push_back(Image{nx, ny, std::vector<unsigned char>(texels, texels + 3 * (size_t)nx * ny)}); return (int)s->images.size() - 1; }
int orc_add_material(void* h, int kind, int tex, double param) {
  Scene* s = (Scene*)h; s->mat.push_back(Mat{kind, tex, param}); return (int)s->mat.size() - 1; }
int orc_add_node(void* h, int kind, int material, int leaf_id, const double* params, int nparams, const int* children, int nchildren) {
  Scene* s = (Scene*)h; Node n; std::memset(&n, 0, sizeof(n));
  n.kind = kind; n.material = material; n.leaf_id = leaf_id;
  for (int i = 0; i < nparams && i < 16; ++i) n.p[i] = params[i];
  n.child_begin = (int)s->children.size(); n.child_count = nchildren;
  for (int i = 0; i < nchildren; ++i) s->children.push_back(children[i]);
  s->nodes.push_back(n);
  if (leaf_id >= 0) { if ((int)s->leaf_node.size() <= leaf_id) s->leaf_node.resize(leaf_id + 1, -1); s->leaf_node[leaf_id] = (int)s->nodes.size() - 1; }
  return (int)s->nodes.size() - 1; }
// lights by LEAF id (== primitive id of the flattened scene); only un-instanced spheres / rects
int orc_set_lights(void* h, const int* leaf_ids, int n) {
  Scene* s = (Scene*)h; s->lights.clear();
  for (int i = 0; i < n; ++i) {
    int lf = leaf_ids[i]; if (lf < 0 || lf >= (int)s->leaf_node.size() || s->leaf_node[lf] < 0) return -1;
    int k = s->nodes[s->leaf_node[lf]].kind; if (k != N_SPHERE && k != N_XY_RECT && k != N_XZ_RECT && k != N_YZ_RECT) return -2;
    s->lights.push_back(s->leaf_node[lf]);
  }
  return 0; }
int orc_add_patch(void* h, const double* cp48) { Scene* s = (Scene*)h; s->patches.insert(s->patches.end(), cp48, cp48 + 48); return (int)(s->patches.size() / 48) - 1; }
void orc_set_root(void* h, int node) { ((Scene*)h)->root = node; }
void orc_set_camera(void* h, const double* cam24) { std::memcpy(((Scene*)h)->cam, cam24, 24 * sizeof(double)); }
void orc_set_sky(void* h, int kind) { ((Scene*)h)->sky = kind; }
void orc_set_perlin(void* h, const double* ranvec, const int* px, const int* py, const int* pz) {
  Scene* s = (Scene*)h; std::memcpy(s->ranvec, ranvec, sizeof(s->ranvec));
  std::memcpy(s->perm_x, px, 256 * sizeof(int)); std::memcpy(s->perm_y, py, 256 * sizeof(int)); std::memcpy(s->perm_z, pz, 256 * sizeof(int)); }
void orc_set_exclude_leaf(void* h, int leaf) { ((Scene*)h)->exclude_leaf = leaf; }

// camera.scm:63-78 make-camera → the 10 slots as 24 doubles.
void orc_make_camera(const double* lookfrom, const double* lookat, const double* vup, double vfov, double aspect,
                     double aperture, double focus_dist, double time0, double time1, double* out24) {
  typedef double T;
  T theta = vfov * (PI / 180.0);
  T half_height = std::tan(theta / 2);
  T half_width = aspect * half_height;
  V3<T> from = ld3<T>(lookfrom), at = ld3<T>(lookat), up = ld3<T>(vup);
  V3<T> w = unit(sub(from, at));
  V3<T> u = unit(cross(up, w));
  V3<T> v = cross(w, u);
  V3<T> llc = sub(sub(sub(from, scale(u, half_width * focus_dist)), scale(v, half_height * focus_dist)), scale(w, focus_dist));
  V3<T> horiz = scale(u, 2 * half_width * focus_dist);
  V3<T> vert = scale(v, 2 * half_height * focus_dist);
  V3<T> slots[7] = {llc, horiz, vert, from, w, u, v};
  for (int i = 0; i < 7; ++i) { out24[3 * i] = slots[i].x; out24[3 * i + 1] = slots[i].y; out24[3 * i + 2] = slots[i].z; }
  out24[21] = aperture / 2; out24[22] = time0; out24[23] = time1;
}

// Fixed-ray-batch closest hit: rays = n x [ox oy oz dx dy dz time]; out arrays sized n (t, u, v),
// 3n (p, nrm); leaf = -1 on miss.  precision: 64 or 32.
int orc_trace_batch(void* h, int n, const double* rays, double tmin, double tmax, int precision,
                    int* leaf, double* t, double* p, double* nrm, double* uv, int* mat) {
  Scene* s = (Scene*)h;
  for (int i = 0; i < n; ++i) {
    const double* q = rays + 7 * i;
    bool hit; double rt = 0, rp[3] = {0, 0, 0}, rn[3] = {0, 0, 0}, ru = 0, rv = 0; int rl = -1, rm = -1;
    if (precision == 32) {
      Ray<float> r{ld3<float>(q), ld3<float>(q + 3), (float)q[6]}; HitRec<float> rec;
      RngAddr ra{0u, (uint32_t)i, 0u, 1u};
      hit = hit_node<float>(*s, s->root, r, (float)tmin, (float)tmax, rec, &ra);
      if (hit) { rt = rec.t; rp[0] = rec.p.x; rp[1] = rec.p.y; rp[2] = rec.p.z; rn[0] = rec.n.x; rn[1] = rec.n.y; rn[2] = rec.n.z; ru = rec.u; rv = rec.v; rl = rec.leaf; rm = rec.mat; }
    } else {
      Ray<double> r{ld3<double>(q), ld3<double>(q + 3), q[6]}; HitRec<double> rec;
      RngAddr ra{0u, (uint32_t)i, 0u, 1u};
      hit = hit_node<double>(*s, s->root, r, tmin, tmax, rec, &ra);
      if (hit) { rt = rec.t; rp[0] = rec.p.x; rp[1] = rec.p.y; rp[2] = rec.p.z; rn[0] = rec.n.x; rn[1] = rec.n.y; rn[2] = rec.n.z; ru = rec.u; rv = rec.v; rl = rec.leaf; rm = rec.mat; }
    }
    leaf[i] = hit ? rl : -1; t[i] = rt;
    if (p) { p[3 * i] = rp[0]; p[3 * i + 1] = rp[1]; p[3 * i + 2] = rp[2]; }
    if (nrm) { nrm[3 * i] = rn[0]; nrm[3 * i + 1] = rn[1]; nrm[3 * i + 2] = rn[2]; }
    if (uv) { uv[2 * i] = ru; uv[2 * i + 1] = rv; }
    if (mat) mat[i] = hit ? rm : -1;
  }
  return 0;
}

// Bezier diagnostics for KAT7-10: returns hit flag, fills t, p, n, max_depth, converge calls.
int orc_bezier_hit(const double* cp12, double width, const double* ray7, double tmin, double tmax,
                   double* t, double* p, double* nrm, int* max_depth, int* converge_calls) {
  Node nd; std::memset(&nd, 0, sizeof(nd)); nd.kind = N_BEZIER; nd.material = 0; nd.leaf_id = 0;
  for (int i = 0; i < 12; ++i) nd.p[i] = cp12[i]; nd.p[12] = width;
  Ray<double> r{ld3<double>(ray7), ld3<double>(ray7 + 3), ray7[6]}; HitRec<double> rec; BezStats st{0, 0};
  bool hit = hit_bezier<double>(nd, r, tmin, tmax, rec, &st);
  *max_depth = st.max_depth; *converge_calls = st.converge_calls;
  if (hit) { *t = rec.t; p[0] = rec.p.x; p[1] = rec.p.y; p[2] = rec.p.z; nrm[0] = rec.n.x; nrm[1] = rec.n.y; nrm[2] = rec.n.z; }
  return hit ? 1 : 0;
}

int orc_aabb_hit(const double* bmin, const double* bmax, const double* ray7, double tmin, double tmax) {
  Ray<double> r{ld3<double>(ray7), ld3<double>(ray7 + 3), ray7[6]};
  return aabb_hit<double>(ld3<double>(bmin), ld3<double>(bmax), r, tmin, tmax) ? 1 : 0;
}

void orc_philox4x32_10(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { philox4x32_10(ctr4, key2, out4); }
void orc_rng_block(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, double* out4) {
  RngAddr a{seed, pixel, sample, bounce}; rng_block<double>(a, block, out4); }

// Shading-unit hooks (texture / sampler / sky / camera) for unit-level parity with the CUDA path.
void orc_tex_value(void* h, int tex, int n, const double* uvp5, int quirks, double* rgb) {
  Scene* s = (Scene*)h;
  for (int i = 0; i < n; ++i) { const double* q = uvp5 + 5 * i; V3<double> c = tex_value<double>(*s, tex, q[0], q[1], ld3<double>(q + 2), quirks);
    rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z; } }
void orc_noise(void* h, int n, const double* p3, int quirks, int turb, double* out) {
  Scene* s = (Scene*)h;
  for (int i = 0; i < n; ++i) out[i] = turb ? perlin_turb<double>(*s, ld3<double>(p3 + 3 * i), quirks) : perlin_noise<double>(*s, ld3<double>(p3 + 3 * i), quirks); }
void orc_get_ray(void* h, double s_, double t_, double xi_time, uint32_t seed, uint32_t pixel, uint32_t sample, double* ray7) {
  Scene* s = (Scene*)h; RngAddr a{seed, pixel, sample, 0u};
  Ray<double> r = get_ray<double>(*s, s_, t_, xi_time, a);
  ray7[0] = r.o.x; ray7[1] = r.o.y; ray7[2] = r.o.z; ray7[3] = r.d.x; ray7[4] = r.d.y; ray7[5] = r.d.z; ray7[6] = r.time; }
// material.scm:41-43, 59-67, 69-74 helper functions, for hand-derived known-answer tests
void orc_reflect(const double* v3, const double* n3, double* out3) { V3<double> r = reflect<double>(ld3<double>(v3), ld3<double>(n3)); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; }
int orc_refract(const double* v3, const double* n3, double ni_over_nt, int quirks, double* out3) {
  V3<double> r = mk<double>(0, 0, 0);
  bool ok = refract<double>(ld3<double>(v3), ld3<double>(n3), ni_over_nt, r, quirks);
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; return ok ? 1 : 0;
}
double orc_schlick(double cosine, double ref_idx) { return schlick<double>(cosine, ref_idx); }
void orc_sky(void* h, const double* d3, double* rgb) {
  Scene* s = (Scene*)h; Ray<double> r{mk<double>(0, 0, 0), ld3<double>(d3), 0.0}; V3<double> c = sky_value<double>(*s, r); rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z; }
void orc_onb_cosine(const double* n3, double r1, double r2, int quirks, double* target3) {
  Onb<double> o = make_onb_from_w(ld3<double>(n3)); V3<double> t = onb_local(o, random_cosine_direction<double>(r1, r2, quirks));
  target3[0] = t.x; target3[1] = t.y; target3[2] = t.z; }
void orc_onb_local(const double* n3, const double* a3, double* out3) {                     // onb.scm:8-16, 27-36
  Onb<double> o = make_onb_from_w(ld3<double>(n3)); V3<double> t = onb_local(o, ld3<double>(a3));
  out3[0] = t.x; out3[1] = t.y; out3[2] = t.z; }
// (local uvw (random-cosine-direction)) as the reference evaluates it (Q15): r6 = its six random-real draws in call order
void orc_onb_local_cosine(const double* n3, const double* r6, int quirks, double* target3) {
  Onb<double> o = make_onb_from_w(ld3<double>(n3));
  V3<double> t = onb_local(o, (quirks & Q15_LOCAL_TRIPLE_EVAL) ? cosine_direction_triple<double>(r6, quirks) : random_cosine_direction<double>(r6[0], r6[1], quirks));
  target3[0] = t.x; target3[1] = t.y; target3[2] = t.z; }

// test hooks for the reference-executed goldens (tests/golden/ref_scatter.json)
int orc_metal_scatter(const double* d3, const double* n3, double fuzz, uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, double* out3) {
  RngAddr a{seed, pixel, sample, bounce}; V3<double> o;
  bool valid = metal_scatter<double>(ld3<double>(d3), ld3<double>(n3), fuzz, a, o);
  out3[0] = o.x; out3[1] = o.y; out3[2] = o.z; return valid ? 1 : 0; }
void orc_dielectric_scatter(const double* d3, const double* n3, double ref_idx, double xi, int quirks, double* out3) {
  V3<double> o = dielectric_scatter<double>(ld3<double>(d3), ld3<double>(n3), ref_idx, xi, quirks);
  out3[0] = o.x; out3[1] = o.y; out3[2] = o.z; }
void orc_random_in_unit_sphere(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t first_block, double* out3) {
  RngAddr a{seed, pixel, sample, bounce}; V3<double> o = random_in_unit_sphere<double>(a, first_block); out3[0] = o.x; out3[1] = o.y; out3[2] = o.z; }
void orc_random_in_unit_disk(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t first_block, double* out3) {
  RngAddr a{seed, pixel, sample, bounce}; V3<double> o = random_in_unit_disk<double>(a, first_block); out3[0] = o.x; out3[1] = o.y; out3[2] = o.z; }
void orc_random_to_sphere(double radius, double distance_sq, double r1, double r2, double* out3) {
  V3<double> o = random_to_sphere<double>(radius, distance_sq, r1, r2); out3[0] = o.x; out3[1] = o.y; out3[2] = o.z; }
double orc_cosine_pdf_value(const double* w3, const double* dir3) {                        // pdf.scm:18-23 (w normalised by make-onb-from-w)
  Onb<double> o = make_onb_from_w(ld3<double>(w3)); return cosine_pdf_value<double>(o.w, ld3<double>(dir3)); }

// Render: accumulates samples [spp_begin, spp_end) into rgb_sum (w*h*3 doubles, y=0 bottom row).
int orc_render(void* h, int w, int hh, int spp_begin, int spp_end, int max_depth, uint32_t seed, int quirks,
               int nthreads, int precision, double* rgb_sum, uint64_t* nrays, int estimator) {
  Scene* s = (Scene*)h;
  if (precision == 32) render<float>(*s, w, hh, spp_begin, spp_end, max_depth, seed, quirks, nthreads, rgb_sum, nrays, estimator);
  else render<double>(*s, w, hh, spp_begin, spp_end, max_depth, seed, quirks, nthreads, rgb_sum, nrays, estimator);
  return 0;
}

// main.scm:123-124, 481-487: correct-gamma + quantise.  Negative sums are out of the reference's
// domain (sqrt of a negative is complex in Gauche); clamp them to 0 here.
void orc_resolve(const double* rgb_sum, int w, int hh, int spp, uint8_t* image) {
  for (int i = 0; i < w * hh * 3; ++i) {
    double c = rgb_sum[i] / (double)spp;
    c = std::sqrt(c < 0 ? 0 : c);
    double q = std::floor(255.99 * std::min(1.0, c));
    image[i] = (uint8_t)q;
  }
}
// main.scm:439-450 save-as-ppm: "P3\n W H\n255\n" (note the space before W), rows top to bottom.
int orc_save_ppm(const char* path, const uint8_t* image, int nx, int ny) {
  FILE* f = std::fopen(path, "w"); if (!f) return -1;
  std::fprintf(f, "P3\n %d %d\n255\n", nx, ny);
  for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) {
    int i = ((ny - y - 1) * nx + x) * 3;
    std::fprintf(f, "%d %d %d\n", image[i], image[i + 1], image[i + 2]);
  }
  std::fclose(f); return 0;
}

}  // extern "C"
