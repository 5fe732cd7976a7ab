// srt_device.cuh — device-side building blocks of the radiance loop (fp32 register math).
//
// Each function cites the reference Scheme it replaces (paths under /root/reference).  The
// reference computes in f64 on boxed f64vectors (vec.scm); here vec3 is three registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/srt.h"

// ---- bounds-checked build (-DSRT_BOUNDS_CHECK; tools/r2_bounds.sh) ---------------------------------------------------
// compute-sanitizer is not available on the B200 pool (profiles/r2_sanitizer_unavailable.txt), so the indices the
// kernels trust - queue slots, compaction positions, node / primitive / material / texture ids, traversal-stack
// pointers, pixel ids, radix-sort scatter positions - are checked by the kernels themselves in this build: a
// violation is counted (per translation unit), the access is skipped, and the entry point fails with the code of
// the first violation.  The release build compiles the checks away.
#ifdef SRT_BOUNDS_CHECK
static __device__ unsigned long long d_srt_violations = 0ull;
static __device__ int d_srt_first_violation = 0;
#define SRT_BOUNDS_OK(cond, code) ((cond) ? true : (atomicAdd(&d_srt_violations, 1ull), atomicCAS(&d_srt_first_violation, 0, (code)), false))
#else
#define SRT_BOUNDS_OK(cond, code) true
#endif

#define SRT_PI 3.14159265358979323846f
#define SRT_MAX_FLOAT 999999999999.0f  // constant.scm:6
#define SRT_MAX_GLOBAL 8

// ------------------------------------------------------------------------------------------------
// Device view of a committed scene ("primitive SoA, material and texture tables, Perlin tables").
struct DScene {
  int n_prims, n_surf, n_nodes, n_xforms, n_mats, n_tex, bvh_depth;   // n_surf = scene surfaces; the rest are medium boundaries
  int n_items;              // surfaces inside the LBVH
  int n_global; int global_prims[SRT_MAX_GLOBAL];   // huge surfaces tested linearly before traversal
  int global_small;         // bit g: global_prims[g] is a sphere that is small against the scene (an "outlier" taken out of the
                            // tree by the cost search, e.g. the three r = 1 spheres of random-scene): fp32 sphere test
  const int4* prim_hdr;     // x = type | flags << 8, y = material, z = xform, w = 0
  const float4* prim_a;     // sphere: c.xyz r | rect: a0 a1 b0 b1 | bezier: A.xyz width
  const float4* prim_b;     // moving: c1.xyz time0 | rect: k | bezier: B.xyz
  const float4* prim_c;     // moving: time1 | bezier: C.xyz
  const float4* prim_d;     // bezier: D.xyz
  const float4* xf;         // 2 per xform: (sin cos off.x off.y) (off.z 0 0 0)
  const float4* nodes;      // 4 per LBVH node == SrtBvhNode
  const int4* mats;         // kind, tex, float_as_int(param), 0
  const float4* tex;        // 2 per texture: (kind even odd scale as int bits / float) (r g b 0)
  const float4* ranvec;     // 256 unit gradient vectors (perlin.scm:33)
  const uint8_t* perm;      // 3 x 256: perm-x, perm-y, perm-z (perlin.scm:34-36)
  const float4* prim_shade; // 2 per primitive: (albedo.rgb if its texture is constant, material param) and
                            // (material kind, texture id, texture-is-constant, 0) as int bits: flattens the
                            // prim -> material -> texture -> colour chain of dependent loads for the shade kernel
  const unsigned char* img_texels;   // image-texture texels (8-bit RGB), all images concatenated
  const int4* imgs;                  // per image: (nx, ny, byte offset, 0)
  const float4* patch_cp;   // 16 control points per (sub-)patch (north-star extension)
  const int* prim_logical;  // logical primitive id reported by the parity hook (sub-patches share their parent's)
  const int* lights;        // primitive ids sampled by the hittable pdf (pdf.scm:28-32)
  int n_lights;
};

struct DCamera { float3 llc, horiz, vert, origin, w, u, v; float lens_radius, time0, time1; };

// ------------------------------------------------------------------------------------------------
// vec.scm
__device__ __forceinline__ float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }   // vec.scm:20
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }   // vec.scm:26
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }   // vec.scm:35
__device__ __forceinline__ float3 operator*(float3 a, float k) { return v3(a.x * k, a.y * k, a.z * k); }          // vec.scm:41
__device__ __forceinline__ float3 operator-(float3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }    // vec.scm:52
__device__ __forceinline__ float length(float3 a) { return sqrtf(dot(a, a)); }                                    // vec.scm:54
__device__ __forceinline__ float3 unit(float3 a) { float k = 1.0f / length(a); return a * k; }                    // vec.scm:60
__device__ __forceinline__ float3 cross(float3 a, float3 b) {                                                      // vec.scm:64
  return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float3 madd(float3 a, float k, float3 b) { return v3(fmaf(a.x, k, b.x), fmaf(a.y, k, b.y), fmaf(a.z, k, b.z)); }  // a*k + b
__device__ __forceinline__ float3 xyz(float4 a) { return v3(a.x, a.y, a.z); }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10, keyed by (pixel, seed), counter (sample, bounce, block, 0)  — replaces srfi-27
// random-real.  uniform = ((x >> 9) + 0.5) * 2^-23: a 23-bit grid whose every point (including
// the +0.5) is exactly representable in fp32, strictly inside (0,1) like srfi-27's random-real.
// (A 24-bit grid is NOT: 16777215.5f rounds to 2^24, i.e. u = 1.0, and sqrt(1 - r2) = 0 then
// turns the lambertian weight into 0 * inf.)
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
struct RngAddr { uint32_t seed, pixel, sample, bounce; };
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 9) + 0.5f) * (1.0f / 8388608.0f); }
__device__ __forceinline__ float4 rng_block(const RngAddr& a, uint32_t block) {
  uint4 r = philox4x32_10(a.sample, a.bounce, block, 0u, a.pixel, a.seed);
  return make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
}

// util.scm:9-15 random-in-unit-sphere (iteration j draws block first+j)
__device__ __forceinline__ float3 random_in_unit_sphere(const RngAddr& a, uint32_t first) {
  for (uint32_t j = 0;; ++j) {
    float4 u = rng_block(a, first + j);
    float3 p = v3(2.0f * u.x - 1.0f, 2.0f * u.y - 1.0f, 2.0f * u.z - 1.0f);
    if (dot(p, p) < 1.0f) return p;
  }
}
// util.scm:17-23 random-in-unit-disk (two candidates per block)
__device__ __forceinline__ float3 random_in_unit_disk(const RngAddr& a, uint32_t first) {
  for (uint32_t j = 0;; ++j) {
    float4 u = rng_block(a, first + j);
    float3 p = v3(2.0f * u.x - 1.0f, 2.0f * u.y - 1.0f, 0.0f);
    if (dot(p, p) < 1.0f) return p;
    p = v3(2.0f * u.z - 1.0f, 2.0f * u.w - 1.0f, 0.0f);
    if (dot(p, p) < 1.0f) return p;
  }
}
// util.scm:37-44 random-cosine-direction (Q1: x,y scaled by 2)
__device__ __forceinline__ float3 random_cosine_direction(float r1, float r2, int quirks) {
  float z = sqrtf(1.0f - r2);
  float s, c;
  sincospif(2.0f * r1, &s, &c);                       // phi = 2*pi*r1
  float k = (quirks & SRT_Q1_COSINE_X2) ? 2.0f : 1.0f;
  float q = k * sqrtf(r2);
  return v3(c * q, s * q, z);
}

// ------------------------------------------------------------------------------------------------
// Instance transform (geometry.scm:465-543 translate / rotate-y, composed on the host).
struct Xf { float s, c; float3 off; };
__device__ __forceinline__ Xf load_xf(const DScene& sc, int id) {
  float4 a = __ldg(&sc.xf[2 * id]), b = __ldg(&sc.xf[2 * id + 1]);
  Xf x; x.s = a.x; x.c = a.y; x.off = v3(a.z, a.w, b.x); return x;
}
// world -> object (geometry.scm:467, 512-521)
__device__ __forceinline__ float3 xf_point_to_obj(const Xf& x, float3 p) {
  float3 q = p - x.off;
  return v3(x.c * q.x - x.s * q.z, q.y, x.s * q.x + x.c * q.z);
}
__device__ __forceinline__ float3 xf_vec_to_obj(const Xf& x, float3 d) { return v3(x.c * d.x - x.s * d.z, d.y, x.s * d.x + x.c * d.z); }
// object -> world (geometry.scm:473, 526-535)
__device__ __forceinline__ float3 xf_vec_to_world(const Xf& x, float3 n) { return v3(x.c * n.x + x.s * n.z, n.y, -x.s * n.x + x.c * n.z); }
__device__ __forceinline__ float3 xf_point_to_world(const Xf& x, float3 p) { return xf_vec_to_world(x, p) + x.off; }

// ------------------------------------------------------------------------------------------------
// Primitive intersectors.  All return the candidate t with only the LOWER bound applied; the
// caller applies the upper bound / exact-tie rule against the current best.

// geometry.scm:146-175 sphere.  The reference evaluates oc, a, b, c and the discriminant
// b^2 - a*c in f64.  For the r = 100/1000 ground spheres (main.scm:38,160,319) those terms cancel
// catastrophically in fp32 (t off by 1e-2 relative for origins near the surface), so exactly
// these terms are evaluated in FP64 here — B200 issues DFMA at half the FFMA rate, and sphere
// tests are a small share of the extend kernel — with the reference's own formula; the square
// root and the two roots are then finished in fp32 in the cancellation-free form
// q = -(b + sign(b) sqrt(disc)), roots {q/a, c/q}.  Miss iff disc <= 0, like the reference.
__device__ __forceinline__ bool isect_sphere(float3 c, float r, float3 o, float3 d, float inv_a, float tmin, float& t) {
  double ocx = (double)o.x - (double)c.x, ocy = (double)o.y - (double)c.y, ocz = (double)o.z - (double)c.z;
  double dx = d.x, dy = d.y, dz = d.z;
  double a = dx * dx + dy * dy + dz * dz;
  double b = ocx * dx + ocy * dy + ocz * dz;
  double cc = ocx * ocx + ocy * ocy + ocz * ocz - (double)r * (double)r;
  double disc = b * b - a * cc;
  if (!(disc > 0.0)) return false;
  float bf = (float)b, ccf = (float)cc;
  // b, c, disc carry the FP64 accuracy; the fp32 finish uses the 1-instruction approximate sqrt and
  // divide (<= 2 ulp each: ~2e-7 relative on t, far inside the 1e-4 parity tolerance and the 1e-5
  // near-tie filter) instead of the ~8-instruction IEEE sequences: this runs for every ray x
  // every primitive kept outside the tree, and at ~5/32 lanes for leaf tests
  float sq; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"((float)disc));
  float q = -(bf + copysignf(sq, bf));
  float tb; asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(tb) : "f"(ccf), "f"(q));
  float ta = q * inv_a;                          // the two roots (-b -/+ sqrt(disc))/a in some order
  float t1 = fminf(ta, tb), t2 = fmaxf(ta, tb);
  if (t1 > tmin) { t = t1; return true; }        // (< t-min temp ...) strict
  if (t2 > tmin) { t = t2; return true; }
  return false;
}
// The same test in fp32 only, for spheres that are SMALL against the scene (|r| < extent / 64, decided
// at commit: SRT_MASK_LEAF32) - the many small spheres of random-scene, never the r = 100 / 1000 grounds.
// Conditioning comes from the formulation instead of from FP64 (Haines / Guenther / Akenine-Moeller,
// "Precision improvements for ray / sphere intersection"): the discriminant is taken from the
// distance of the centre to the ray, disc / a = r^2 - |oc - (b / a) d|^2, which does not cancel for far
// origins (b^2 - a c does: both terms ~ |oc|^2 |d|^2); c = oc.oc - r^2 only cancels for origins within
// ~r / 300 of the surface (relative error of t ~ eps r / 2h), and the root that uses it is then ~0 and
// rejected by t-min.  Same acceptance rule as above (miss iff disc <= 0, strict t-min).  No F2F / DFMA:
// this runs at ~5 of 32 lanes in the leaf loop, where the FP64 form was 10 conversions on the
// quarter-rate XU pipe plus ~17 half-rate FP64 instructions per test.
__device__ __forceinline__ bool isect_sphere32(float3 c, float r, float3 o, float3 d, float inv_a, float tmin, float& t) {
  const float3 oc = o - c;
  const float b = dot(oc, d);
  const float k = b * inv_a;
  const float3 l = v3(fmaf(-k, d.x, oc.x), fmaf(-k, d.y, oc.y), fmaf(-k, d.z, oc.z));   // centre -> closest point of the ray
  const float discr = fmaf(r, r, -dot(l, l));                                            // = disc / a
  if (!(discr > 0.0f)) return false;
  const float a = dot(d, d);
  // approximate sqrt / divide, flush-to-zero forms: one MUFU each, no denormal rescue code (<= 2 ulp; see isect_sphere)
  float sq; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(a * discr));
  const float q = -(b + copysignf(sq, b));
  const float cc = fmaf(-r, r, dot(oc, oc));
  float tb; asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(tb) : "f"(cc), "f"(q));
  const float ta = q * inv_a;
  const float t1 = fminf(ta, tb), t2 = fmaxf(ta, tb);
  if (t1 > tmin) { t = t1; return true; }
  if (t2 > tmin) { t = t2; return true; }
  return false;
}
// An INSTANCED sphere (translate / rotate-y above a sphere leaf, geometry.scm:465-543): the ray goes to object space,
// the rigid transform keeps t.  Out of line on purpose (see intersect_prim).
static __device__ __noinline__ bool isect_sphere_instanced(const DScene& sc, int xform, float3 c, float r, float3 o, float3 d, float inv_a, float tmin, float& t);
// geometry.scm:178-182 moving-sphere centre at the ray's time
__device__ __forceinline__ float3 moving_center(float4 a, float4 b, float4 c, float time) {
  float3 c0 = xyz(a), c1 = xyz(b);
  float f = (time - b.w) / (c.x - b.w);
  return madd(c1 - c0, f, c0);
}
// geometry.scm:376-431 rects.  axis = thin axis; (ia, ib) in-plane axes in argument order.
// Inclusive bounds (t < t-min rejects).  A NaN t (ray in the plane) is rejected.
__device__ __forceinline__ float cmp3(float3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__device__ __forceinline__ bool isect_rect(int type, float4 a, float k, float3 o, float3 d, float tmin, float t_inst, bool have_t,
                                           float& t, float& u, float& v) {
  int axis = (type == SRT_PRIM_XY_RECT) ? 2 : (type == SRT_PRIM_XZ_RECT ? 1 : 0);
  int ia = (type == SRT_PRIM_YZ_RECT) ? 1 : 0, ib = (type == SRT_PRIM_XY_RECT) ? 1 : 2;
  float tt = have_t ? t_inst : __fdiv_rn(k - cmp3(o, axis), cmp3(d, axis));     // (/ (- k (v:z (origin ray))) (v:z (dir ray)))
  if (!(tt >= tmin)) return false;                       // also rejects NaN
  float pa = fmaf(tt, cmp3(d, ia), cmp3(o, ia));
  float pb = fmaf(tt, cmp3(d, ib), cmp3(o, ib));
  if (pa < a.x || pa > a.y || pb < a.z || pb > a.w) return false;
  t = tt;
  u = __fdividef(pa - a.x, a.y - a.x);                   // uv feeds texture lookups only: 2-ulp divide
  v = __fdividef(pb - a.z, a.w - a.z);
  return true;
}
// t = (k - o'[axis]) / d'[axis] for an instanced rect, with the rotate-y / translate of origin and
// direction (geometry.scm:467, 512-521) carried in FP64: at Cornell coordinates (~555) the fp32
// rotation loses ~6e-5 absolute in o' (> 1e-4 relative on t for origins within ~0.5 of a face)
// and cancels in d' for rays grazing a rotated face.
__device__ __forceinline__ float rect_t_f64(const Xf& x, int type, float k, float3 o, float3 d) {
  if (type == SRT_PRIM_XZ_RECT) return __fdiv_rn(k - (o.y - x.off.y), d.y);
  double qx = (double)o.x - (double)x.off.x, qz = (double)o.z - (double)x.off.z;
  double num, den;
  if (type == SRT_PRIM_XY_RECT) { num = (double)k - ((double)x.s * qx + (double)x.c * qz); den = (double)x.s * (double)d.x + (double)x.c * (double)d.z; }
  else { num = (double)k - ((double)x.c * qx - (double)x.s * qz); den = (double)x.c * (double)d.x - (double)x.s * (double)d.z; }
  return __fdiv_rn((float)num, (float)den);
}

static __device__ __noinline__ bool isect_sphere_instanced(const DScene& sc, int xform, float3 c, float r, float3 o, float3 d, float inv_a, float tmin, float& t) {
  const Xf x = load_xf(sc, xform);
  return isect_sphere(c, r, xf_point_to_obj(x, o), xf_vec_to_obj(x, d), inv_a, tmin, t);
}

// bezier.scm — cubic Bezier curve with width; recursive subdivision with an explicit stack.
struct Bez { float3 a, b, c, d; };
__device__ __forceinline__ float3 bez_p(const Bez& c, float t) {                    // bezier.scm:67-77
  float t2 = t * t, t3 = t2 * t, mt = 1.0f - t, mt2 = mt * mt, mt3 = mt2 * mt;
  return c.a * mt3 + c.b * (3.0f * mt2 * t) + c.c * (3.0f * mt * t2) + c.d * t3;
}
__device__ __forceinline__ float3 idiv(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }   // bezier.scm:45
__device__ __forceinline__ float3 bez_tan(const Bez& c, float t) {                  // bezier.scm:106-117
  float3 ca = c.b * 3.0f + c.d + c.c * -3.0f + c.a * -1.0f;
  float3 cb = (c.a + c.b * -2.0f + c.c) * 3.0f;
  float3 cc = (c.b - c.a) * 3.0f;
  return unit(ca * (3.0f * t * t) + cb * (2.0f * t) + cc);
}
__device__ __forceinline__ float dot2d(float3 a, float3 b) { return a.x * b.x + a.y * b.y; }
#define SRT_BEZ_MAX_DEPTH 20
// bezier.scm:176-214 hit + :121-175 converge.  tbest plays the role of t-max (closest so far).
static __device__ __noinline__ bool isect_bezier(float4 pa, float4 pb, float4 pc, float4 pd, float3 o, float3 dir, float tmin, float tbest, float& tout) {
  float width = pa.w, width1 = width * 0.5f, width2 = width1 * width1, eps = width / 20.0f;
  // get-projection-mat bezier.scm:13-43 (row-vector convention, swizzle (x,y,z)->(x,-z,y))
  float3 ud = unit(dir);
  float lx = ud.x, ly = -ud.z, lz = ud.y;
  float dd = sqrtf(lx * lx + lz * lz);
  float R[3][3];
  if (dd == 0.0f) {
    float ang = (ly >= 0.0f) ? -0.5f * SRT_PI : 0.5f * SRT_PI;
    float ca = cosf(ang), sa = sinf(ang);
    R[0][0] = 1; R[0][1] = 0; R[0][2] = 0; R[1][0] = 0; R[1][1] = ca; R[1][2] = -sa; R[2][0] = 0; R[2][1] = sa; R[2][2] = ca;
  } else {
    R[0][0] = lz / dd; R[0][1] = (-lx * ly) / dd; R[0][2] = lx;
    R[1][0] = 0.0f;    R[1][1] = dd;              R[1][2] = ly;
    R[2][0] = -lx / dd; R[2][1] = (-ly * lz) / dd; R[2][2] = lz;
  }
  float3 so = v3(o.x, -o.z, o.y);
  auto xform = [&](float4 q) {                        // bezier.scm:49-55 transform
    float3 s = v3(q.x, -q.z, q.y) - so;
    return v3(s.x * R[0][0] + s.y * R[1][0] + s.z * R[2][0], s.x * R[0][1] + s.y * R[1][1] + s.z * R[2][1],
              s.x * R[0][2] + s.y * R[1][2] + s.z * R[2][2]);
  };
  Bez cur; cur.a = xform(pa); cur.b = xform(pb); cur.c = xform(pc); cur.d = xform(pd);
  float l0 = fmaxf(fmaxf(fabsf(cur.a.x - 2.0f * cur.b.x + cur.c.x), fabsf(cur.a.y - 2.0f * cur.b.y + cur.c.y)),
                   fmaxf(fabsf(cur.b.x - 2.0f * cur.c.x + cur.d.x), fabsf(cur.b.y - 2.0f * cur.c.y + cur.d.y)));
  int max_depth = 0;                                   // bezier.scm:179-192
  if (l0 > 0.0f) max_depth = (int)ceilf(logf((1.41421356237f * 4.0f * 3.0f * l0) / (8.0f * eps)) * (1.0f / 1.38629436112f));
  if (max_depth > SRT_BEZ_MAX_DEPTH) max_depth = SRT_BEZ_MAX_DEPTH;
  // depth-first over the binary subdivision tree; the stack keeps the pending right halves
  Bez stk[SRT_BEZ_MAX_DEPTH + 1]; float sv0[SRT_BEZ_MAX_DEPTH + 1], svn[SRT_BEZ_MAX_DEPTH + 1]; int sdepth[SRT_BEZ_MAX_DEPTH + 1];
  int sp = 0; float v0 = 0.0f, vn = 1.0f; int depth = max_depth;
  const float t = tbest;                               // both halves see the same incoming t (let*-values)
  bool any = false; float tmin_found = tbest;
  for (;;) {
    bool descend = false;
    // bbox reject bezier.scm:126-129 (bbox = cps -/+ width1)
    float bminx = fminf(fminf(cur.a.x, cur.b.x), fminf(cur.c.x, cur.d.x)) - width1, bmaxx = fmaxf(fmaxf(cur.a.x, cur.b.x), fmaxf(cur.c.x, cur.d.x)) + width1;
    float bminy = fminf(fminf(cur.a.y, cur.b.y), fminf(cur.c.y, cur.d.y)) - width1, bmaxy = fmaxf(fmaxf(cur.a.y, cur.b.y), fmaxf(cur.c.y, cur.d.y)) + width1;
    float bminz = fminf(fminf(cur.a.z, cur.b.z), fminf(cur.c.z, cur.d.z)) - width1, bmaxz = fmaxf(fmaxf(cur.a.z, cur.b.z), fmaxf(cur.c.z, cur.d.z)) + width1;
    if (!(bminz >= t || bmaxz <= 0.000001f || bminx >= width1 || bmaxx <= -width1 || bminy >= width1 || bmaxy <= -width1)) {
      if (depth < 0) {                                 // leaf bezier.scm:130-166
        float3 dirv = cur.d - cur.a;
        float3 dp0 = bez_tan(cur, 0.0f);
        if (dot2d(dirv, dp0) < 0.0f) dp0 = -dp0;
        if (!(dot2d(dp0, -cur.a) < 0.0f)) {
          float3 dpn = bez_tan(cur, 1.0f);
          if (dot2d(dirv, dpn) < 0.0f) dpn = -dpn;
          if (!(dot2d(dpn, cur.d) < 0.0f)) {
            float w = dirv.x * dirv.x + dirv.y * dirv.y;
            if (w != 0.0f) {
              w = (cur.a.x * dirv.x + cur.a.y * dirv.y) / (-w);
              w = fminf(fmaxf(w, 0.0f), 1.0f);
              float v = v0 * (1.0f - w) + vn * w;
              float3 p = bez_p(cur, v);                // Q8: sub-curve at the global parameter
              if (!((p.x * p.x + p.y * p.y) >= width2 || p.z <= 0.0001f || t < p.z)) {
                any = true; if (p.z < tmin_found) tmin_found = p.z;
              }
            }
          }
        }
      } else {
        descend = true;
      }
    }
    if (descend) {                                     // split at 0.5 bezier.scm:78-87, left first
      float3 sp_ = bez_p(cur, 0.5f);
      float3 nbc = idiv(cur.b, cur.c, 0.5f), lb = idiv(cur.a, cur.b, 0.5f), lc = idiv(lb, nbc, 0.5f);
      float3 rc = idiv(cur.c, cur.d, 0.5f), rb = idiv(nbc, rc, 0.5f);
      float vm = (v0 + vn) / 2.0f;
      stk[sp].a = sp_; stk[sp].b = rb; stk[sp].c = rc; stk[sp].d = cur.d; sv0[sp] = vm; svn[sp] = vn; sdepth[sp] = depth - 1; ++sp;
      cur.b = lb; cur.c = lc; cur.d = sp_; vn = vm; depth -= 1;
    } else {
      if (sp == 0) break;
      --sp; cur = stk[sp]; v0 = sv0[sp]; vn = svn[sp]; depth = sdepth[sp];
    }
  }
  if (any && tmin < tmin_found) { tout = tmin_found; return true; }   // (and hit? (< t-min t))
  return false;
}

// ------------------------------------------------------------------------------------------------
// Bicubic Bezier PATCH (north-star extension; the reference only has the curve).  Same
// specification as the oracle (DESIGN.md "Patches"): ray-space projection of bezier.scm, depth-2
// quadtree subdivision by de Casteljau (done once on the host: 16 leaves per patch), hull culling,
// Newton on (S.x, S.y) = 0 from each leaf's centre, t = S.z / |d|.
__device__ __forceinline__ void bern(float s, float b[4], float db[4]) {   // cubic Bernstein weights and derivatives
  float m = 1.0f - s;
  b[0] = m * m * m; b[1] = 3.0f * s * m * m; b[2] = 3.0f * s * s * m; b[3] = s * s * s;
  db[0] = -3.0f * m * m; db[1] = 3.0f * m * m - 6.0f * s * m; db[2] = 6.0f * s * m - 3.0f * s * s; db[3] = 3.0f * s * s;
}
// One LEAF sub-patch: the host pre-splits every patch SRT_PATCH_LEVELS = 2 levels (16 leaves,
// separate LBVH leaves, so the BVH does the subdivision culling with cheap box tests); here only
// hull cull + Newton from the leaf's centre, with the projected net held in registers (no local
// memory: the smem-staged BVH leaves almost no L1; inlined: as a call the caller-saved registers
// went through local memory around every test, +9 % on cfg5_teapot).  dom = (u0, v0, size) of the leaf in the
// parent patch's domain; u/v out are GLOBAL.
static __device__ __forceinline__ bool isect_patch(const float4* __restrict__ cp, float4 dom, float3 o, float3 dir, float tmin, float tbest,
                                                float& tout, float& uout, float& vout) {
  // ray-space projection (bezier.scm:13-55)
  float3 ud = unit(dir);
  float lx = ud.x, ly = -ud.z, lz = ud.y;
  float dd = sqrtf(lx * lx + lz * lz);
  float R00, R01, R02, R10, R11, R12, R20, R21, R22;
  if (dd == 0.0f) {
    float ang = (ly >= 0.0f) ? -0.5f * SRT_PI : 0.5f * SRT_PI;
    float ca = cosf(ang), sa = sinf(ang);
    R00 = 1; R01 = 0; R02 = 0; R10 = 0; R11 = ca; R12 = -sa; R20 = 0; R21 = sa; R22 = ca;
  } else {
    R00 = lz / dd; R01 = (-lx * ly) / dd; R02 = lx;
    R10 = 0.0f;    R11 = dd;              R12 = ly;
    R20 = -lx / dd; R21 = (-ly * lz) / dd; R22 = lz;
  }
  const float3 so = v3(o.x, -o.z, o.y);
  float X[16], Y[16], Z[16];
  float mnx = 3e38f, mxx = -3e38f, mny = 3e38f, mxy = -3e38f, mnz = 3e38f, mxz = -3e38f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float4 q = __ldg(&cp[k]);
    float3 s = v3(q.x, -q.z, q.y) - so;
    X[k] = s.x * R00 + s.y * R10 + s.z * R20; Y[k] = s.x * R01 + s.y * R11 + s.z * R21; Z[k] = s.x * R02 + s.y * R12 + s.z * R22;
    mnx = fminf(mnx, X[k]); mxx = fmaxf(mxx, X[k]); mny = fminf(mny, Y[k]); mxy = fmaxf(mxy, Y[k]); mnz = fminf(mnz, Z[k]); mxz = fmaxf(mxz, Z[k]);
  }
  const float len = length(dir);
  const float zmin = tmin * len, zmax = tbest * len;
  if (mnx > 0.f || mxx < 0.f || mny > 0.f || mxy < 0.f || mxz < zmin || mnz > zmax) return false;   // hull misses the ray
  float s = 0.5f, t = 0.5f; bool conv = false;
  float bs[4], dbs[4], bt[4], dbt[4];
  for (int it = 0; it < 8; ++it) {
    bern(s, bs, dbs); bern(t, bt, dbt);
    float Sx = 0.f, Sy = 0.f, Sux = 0.f, Suy = 0.f, Svx = 0.f, Svy = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float rx = X[4 * i] * bt[0] + X[4 * i + 1] * bt[1] + X[4 * i + 2] * bt[2] + X[4 * i + 3] * bt[3];
      float ry = Y[4 * i] * bt[0] + Y[4 * i + 1] * bt[1] + Y[4 * i + 2] * bt[2] + Y[4 * i + 3] * bt[3];
      float dx = X[4 * i] * dbt[0] + X[4 * i + 1] * dbt[1] + X[4 * i + 2] * dbt[2] + X[4 * i + 3] * dbt[3];
      float dy = Y[4 * i] * dbt[0] + Y[4 * i + 1] * dbt[1] + Y[4 * i + 2] * dbt[2] + Y[4 * i + 3] * dbt[3];
      Sx = fmaf(bs[i], rx, Sx); Sy = fmaf(bs[i], ry, Sy); Sux = fmaf(dbs[i], rx, Sux); Suy = fmaf(dbs[i], ry, Suy); Svx = fmaf(bs[i], dx, Svx); Svy = fmaf(bs[i], dy, Svy);
    }
    float det = Sux * Svy - Svx * Suy;
    if (!(fabsf(det) > 1e-30f)) break;
    // one correctly rounded reciprocal instead of two IEEE divisions: Newton is self-correcting, the
    // converged (s, t) is the same root (cfg5_teapot +8 %, cfg5 +6.5 %)
    const float idet = __frcp_rn(det);
    float ds = (-Sx * Svy + Sy * Svx) * idet, dt = (-Sux * Sy + Suy * Sx) * idet;
    s += ds; t += dt;
    if (!(fabsf(s) < 4.0f) || !(fabsf(t) < 4.0f)) break;
    if (fmaxf(fabsf(ds), fabsf(dt)) < 1e-5f) { conv = true; break; }
  }
  if (!conv || s < -1e-3f || s > 1.0f + 1e-3f || t < -1e-3f || t > 1.0f + 1e-3f) return false;
  s = fminf(fmaxf(s, 0.f), 1.f); t = fminf(fmaxf(t, 0.f), 1.f);
  bern(s, bs, dbs); bern(t, bt, dbt);
  float Sz = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) Sz = fmaf(bs[i], Z[4 * i] * bt[0] + Z[4 * i + 1] * bt[1] + Z[4 * i + 2] * bt[2] + Z[4 * i + 3] * bt[3], Sz);
  if (!(Sz > zmin && Sz < zmax)) return false;
  tout = Sz / len; uout = dom.x + s * dom.z; vout = dom.y + t * dom.z;
  return true;
}

// Bounding slab of a leaf sub-patch (srt_api.cu fills prim_a = (n', d') with |n'.p - d'| <= 1 on the whole
// convex hull of its control net): true iff the ray stays outside the slab over the leaf's box interval
// [tb0, tb1], i.e. the full test (projection, hull cull, Newton) cannot find a hit.  b = 0 (ray parallel to
// the slab) gives +-inf bounds of the right sign; a degenerate slab (n' = 0) never culls.
__device__ __forceinline__ bool patch_slab_culled(float4 sl, float3 o, float3 d, float tb0, float tb1) {
  const float a = dot(xyz(sl), o) - sl.w, b = dot(xyz(sl), d);
  const float ib = __frcp_rn(b);
  const float t0 = (-1.0f - a) * ib, t1 = (1.0f - a) * ib;
  return fmaxf(fminf(t0, t1), tb0) > fminf(fmaxf(t0, t1), tb1);
}

// ------------------------------------------------------------------------------------------------
// geometry.scm:590-664 Klein / IIS fractal: inversion in six spheres (<= 10 times), sphere-traced
// <= 100 steps, central-difference normal.  Evaluated in FP64: the finite-difference normal of a
// fractal distance field is ill-conditioned in fp32, the primitive is rare and B200 runs DFMA at
// half rate.
__device__ __forceinline__ double klein_dist(double cx, double cy, double cz, double px, double py, double pz) {   // dist-func :602-624
  const double KX[6] = {300, 300, -300, -300, 0, 0}, KY[6] = {300, -300, 300, -300, 0, 0}, KZ[6] = {0, 0, 0, 0, 424.26, -424.26};
  px -= cx; py -= cy; pz -= cz;
  double dr = 1.0;
  for (int iter = 0;;) {
    if (iter >= 10) break;
    bool inverted = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double dx = px - KX[k], dy = py - KY[k], dz = pz - KZ[k];
      double d2 = dx * dx + dy * dy + dz * dz;
      if (!inverted && sqrt(d2) < 300.0) {
        dr = dr * (90000.0 / d2);
        double len = sqrt(d2), f = 1.0 / (len * len);
        px = dx * 90000.0 * f + KX[k]; py = dy * 90000.0 * f + KY[k]; pz = dz * 90000.0 * f + KZ[k];
        ++iter; inverted = true;
      }
    }
    if (!inverted) break;
  }
  return 0.7 * ((sqrt(px * px + py * py + pz * pz) - 125.0) / fabs(dr));
}
static __device__ __noinline__ bool isect_klein(float4 a, float3 o, float3 d, float tmin, float tbest, float& tout) {   // make-klein :644-661
  const double cx = a.x, cy = a.y, cz = a.z, ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z;
  double px = ox, py = oy, pz = oz, ray_length = 0.0;
  for (int iter = 0; iter < 100; ++iter) {
    double dist = klein_dist(cx, cy, cz, px, py, pz);
    ray_length += dist;
    px = ox + dx * ray_length; py = oy + dy * ray_length; pz = oz + dz * ray_length;
    if (dist < 0.001 && (double)tmin < ray_length && ray_length < (double)tbest) { tout = (float)ray_length; return true; }
  }
  return false;
}
static __device__ __noinline__ float3 klein_normal(float4 a, float3 o, float3 d, float t) {   // get-normal :626-632 at origin + dir * t
  const double cx = a.x, cy = a.y, cz = a.z, e = 0.01;
  const double px = (double)o.x + (double)d.x * (double)t, py = (double)o.y + (double)d.y * (double)t, pz = (double)o.z + (double)d.z * (double)t;
  double nx = klein_dist(cx, cy, cz, px + e, py, pz) - klein_dist(cx, cy, cz, px - e, py, pz);
  double ny = klein_dist(cx, cy, cz, px, py + e, pz) - klein_dist(cx, cy, cz, px, py - e, pz);
  double nz = klein_dist(cx, cy, cz, px, py, pz + e) - klein_dist(cx, cy, cz, px, py, pz - e);
  double k = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
  return v3((float)(nx * k), (float)(ny * k), (float)(nz * k));
}

// ------------------------------------------------------------------------------------------------
// Exact-tie rule: the order-independent restatement of hit-obj-list's
// sequential "later object replaces the best iff t < best (sphere-type) or t <= best (rect-type,
// curve)".  ids are positions in the reference's flattened object list.
__device__ __forceinline__ bool prim_inclusive(int type) { return type >= SRT_PRIM_XY_RECT && type <= SRT_PRIM_BEZIER; }
__device__ __forceinline__ bool accept_hit(float t, int id, bool incl, float best_t, int best_id, bool best_incl) {
  if (t < best_t) return true;
  if (!(t == best_t)) return false;
  if (id > best_id) return incl;
  if (id < best_id) return !best_incl;
  return false;
}

struct Hit { float t; int prim; float u, v; bool incl; };

// One leaf primitive against the ray (world space in, candidate merged into `h`).  MASK is the
// set of primitive kinds present in the scene (bit = SRT_PRIM_*): the extend kernel is compiled
// per mask so that e.g. sphere-only scenes carry no rect / instance / Bezier code or registers.
#ifndef SRT_MASK_ALL
#define SRT_MASK_ALL 0x1ff
#endif
#define SRT_MASK_LEAF32 0x400   // not a primitive kind: "the spheres inside the LBVH are small against the scene" (fp32 sphere test)
template <int MASK, class PrimSrc>
__device__ __forceinline__ void intersect_prim(const DScene& sc, const PrimSrc& ps, int id, float3 o, float3 d, float time, float inv_a, float tmin,
                                               const RngAddr& ra, Hit& h) {
  constexpr bool HAS_SPHERE = MASK & 1, HAS_MOVING = MASK & 2, HAS_RECT = MASK & 0x1c, HAS_BEZIER = MASK & 0x20, HAS_MEDIUM = MASK & 0x40, HAS_PATCH = MASK & 0x80, HAS_KLEIN = MASK & 0x100;
  constexpr bool LEAF32 = MASK & SRT_MASK_LEAF32;           // small spheres: the fp32 formulation (isect_sphere32)
  if (!SRT_BOUNDS_OK(id >= 0 && id < sc.n_prims, 101)) return;
  constexpr bool SINGLE_KIND = ((MASK & SRT_MASK_ALL) & ((MASK & SRT_MASK_ALL) - 1)) == 0;
  float4 a = ps.a(id);
  int type, xform = -1, aux = 0;
  if (SINGLE_KIND && (MASK & 0x23)) type = HAS_SPHERE ? SRT_PRIM_SPHERE : (HAS_MOVING ? SRT_PRIM_MOVING_SPHERE : SRT_PRIM_BEZIER);
  else { int4 hdr = ps.hdr(id); type = hdr.x & 0xff; xform = hdr.z; aux = hdr.w; if (!SRT_BOUNDS_OK(xform < sc.n_xforms, 102)) return; }
  float t = 0.f, u = 0.f, v = 0.f; bool ok = false;
  if (HAS_SPHERE && type == SRT_PRIM_SPHERE) {
    // an instanced sphere (geometry.scm:465-543 above a sphere leaf): the rigid transform keeps t, so
    // the ray goes to object space and t comes back unchanged.  Sphere-only kernels (SINGLE_KIND) never
    // see one: such scenes are routed to a variant that reads the header (variant_of, wavefront.cu).
    // (the instanced case is a rare out-of-line call: inlined, its transform code cost the curve / patch variant
    // 13 % on cfg5, where no sphere is instanced)
    if (!SINGLE_KIND && xform >= 0) ok = isect_sphere_instanced(sc, xform, xyz(a), a.w, o, d, inv_a, tmin, t);
    else ok = LEAF32 ? isect_sphere32(xyz(a), a.w, o, d, inv_a, tmin, t) : isect_sphere(xyz(a), a.w, o, d, inv_a, tmin, t);
  } else if (HAS_MOVING && type == SRT_PRIM_MOVING_SPHERE) {
    float4 b = __ldg(&sc.prim_b[id]), c = __ldg(&sc.prim_c[id]);
    const float3 ctr = moving_center(a, b, c, time);
    if (!SINGLE_KIND && xform >= 0) ok = isect_sphere_instanced(sc, xform, ctr, a.w, o, d, inv_a, tmin, t);
    else ok = LEAF32 ? isect_sphere32(ctr, a.w, o, d, inv_a, tmin, t) : isect_sphere(ctr, a.w, o, d, inv_a, tmin, t);
  } else if (HAS_RECT && type <= SRT_PRIM_YZ_RECT) {
    float k = __int_as_float(aux);                       // plane constant, carried in the header (srt_api.cu)
    float3 oo = o, dd = d; float ti = 0.f; bool have_t = false;
    if (xform >= 0) { Xf x = load_xf(sc, xform); oo = xf_point_to_obj(x, o); dd = xf_vec_to_obj(x, d); ti = rect_t_f64(x, type, k, o, d); have_t = true; }
    ok = isect_rect(type, a, k, oo, dd, tmin, ti, have_t, t, u, v);
  } else if (HAS_BEZIER && type == SRT_PRIM_BEZIER) {
    ok = isect_bezier(a, __ldg(&sc.prim_b[id]), __ldg(&sc.prim_c[id]), __ldg(&sc.prim_d[id]), o, d, tmin, h.t, t);
  } else if (HAS_KLEIN && type == SRT_PRIM_KLEIN) {
    ok = isect_klein(a, o, d, tmin, h.t, t);
  } else if (HAS_PATCH && type == SRT_PRIM_PATCH) {
    ok = isect_patch(sc.patch_cp + 16 * aux, __ldg(&sc.prim_b[id]), o, d, tmin, h.t, t, u, v);
  } else if (HAS_MEDIUM && type == SRT_PRIM_CONSTANT_MEDIUM) {
    // geometry.scm:545-578.  Two closest-hit queries on the boundary shapes, then the free flight
    // -log(xi)/density.  The reference clamps the exit to t-max = closest-so-far; accepting
    // nt < exit here and nt < best in accept_hit below is the same condition, order-independent.
    // xi = block (16 + id) of the ray's (pixel, sample, bounce) stream (upstream: (random-real)).
    const int first = (int)a.y, cnt = (int)a.z;
    Hit h1; h1.t = SRT_MAX_FLOAT; h1.prim = -1; h1.u = h1.v = 0.f; h1.incl = false;
    for (int j = first; j < first + cnt; ++j) intersect_prim<0x1d>(sc, ps, j, o, d, time, inv_a, -SRT_MAX_FLOAT, ra, h1);   // boundaries: spheres / rects (commit checks)
    if (h1.prim >= 0) {
      Hit h2; h2.t = SRT_MAX_FLOAT; h2.prim = -1; h2.u = h2.v = 0.f; h2.incl = false;
      for (int j = first; j < first + cnt; ++j) intersect_prim<0x1d>(sc, ps, j, o, d, time, inv_a, h1.t + 0.0001f, ra, h2);
      if (h2.prim >= 0) {
        float t1 = fmaxf(fmaxf(h1.t, tmin), 0.0f);
        float len = length(d);
        float xi = rng_block(ra, 16u + (uint32_t)id).x;
        float hit_distance = -(1.0f / a.x) * logf(xi);
        float nt = t1 + hit_distance / len;
        if (t1 < h2.t && hit_distance < (h2.t - t1) * len) { t = nt; ok = true; }
      }
    }
  }
  (void)aux;
  bool incl = prim_inclusive(type);
  if (ok && accept_hit(t, id, incl, h.t, h.prim, h.incl)) { h.t = t; h.prim = id; h.u = u; h.v = v; h.incl = incl; }
}

// world-space normal of a (sub-)patch at the global (u, v): Su x Sv accumulated row by row so
// that the shade kernel never holds the 4x4 net in registers
__device__ __forceinline__ float3 patch_normal(const float4* __restrict__ cp, float4 dom, float hu, float hv, float3 d) {
  const float s = (hu - dom.x) / dom.z, t = (hv - dom.y) / dom.z;
  float bt[4], dbt[4]; bern(t, bt, dbt);
  const float m = 1.0f - s;
  float3 Su = v3(0.f, 0.f, 0.f), Sv = v3(0.f, 0.f, 0.f);
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const float bi = i == 0 ? m * m * m : (i == 1 ? 3.0f * s * m * m : (i == 2 ? 3.0f * s * s * m : s * s * s));
    const float dbi = i == 0 ? -3.0f * m * m : (i == 1 ? 3.0f * m * m - 6.0f * s * m : (i == 2 ? 6.0f * s * m - 3.0f * s * s : 3.0f * s * s));
    const float3 q0 = xyz(__ldg(&cp[4 * i])), q1 = xyz(__ldg(&cp[4 * i + 1])), q2 = xyz(__ldg(&cp[4 * i + 2])), q3 = xyz(__ldg(&cp[4 * i + 3]));
    Su = Su + (q0 * bt[0] + q1 * bt[1] + q2 * bt[2] + q3 * bt[3]) * dbi;
    Sv = Sv + (q0 * dbt[0] + q1 * dbt[1] + q2 * dbt[2] + q3 * dbt[3]) * bi;
  }
  float3 n = unit(cross(Su, Sv));
  return dot(n, d) > 0.0f ? -n : n;
}

// Hit-record completion: p and normal in world space (ray.scm:27 make-hit-record fields).
// geometry.scm:158-160 (sphere), :386-387 (rects), :438 (flip), :473/:526-535 (instances),
// bezier.scm:209-211 (Q9: p along the raw direction, normal = -dir).
__device__ __forceinline__ void complete_hit(const DScene& sc, int prim, float t, float hu, float hv, float3 o, float3 d, float time, float3& p, float3& n, int& material) {
  if (!SRT_BOUNDS_OK(prim >= 0 && prim < sc.n_prims, 111)) { p = o; n = v3(0.f, 1.f, 0.f); material = 0; return; }
  int4 hdr = __ldg(&sc.prim_hdr[prim]);
  int type = hdr.x & 0xff;
  material = hdr.y;
  (void)SRT_BOUNDS_OK(material >= 0 && material < sc.n_mats && hdr.z < sc.n_xforms, 112);
  float4 a = __ldg(&sc.prim_a[prim]);
  if (type <= SRT_PRIM_MOVING_SPHERE) {
    float3 c = xyz(a);
    if (type == SRT_PRIM_MOVING_SPHERE) c = moving_center(a, __ldg(&sc.prim_b[prim]), __ldg(&sc.prim_c[prim]), time);
    p = madd(d, t, o);
    if (hdr.z >= 0) {                                  // instanced: normal in object space, rotated back (geometry.scm:473, 526-535)
      Xf x = load_xf(sc, hdr.z);
      n = xf_vec_to_world(x, (xf_point_to_obj(x, p) - c) * (1.0f / a.w));
    } else n = (p - c) * (1.0f / a.w);
  } else if (type <= SRT_PRIM_YZ_RECT) {
    n = v3(type == SRT_PRIM_YZ_RECT ? 1.f : 0.f, type == SRT_PRIM_XZ_RECT ? 1.f : 0.f, type == SRT_PRIM_XY_RECT ? 1.f : 0.f);
    if (hdr.z >= 0) {
      Xf x = load_xf(sc, hdr.z);
      float3 po = madd(xf_vec_to_obj(x, d), t, xf_point_to_obj(x, o));
      p = xf_point_to_world(x, po);
      n = xf_vec_to_world(x, n);
    } else {
      p = madd(d, t, o);
    }
  } else if (type == SRT_PRIM_KLEIN) {                // geometry.scm:657-658
    p = madd(d, t, o);
    n = klein_normal(a, o, d, t);
  } else if (type == SRT_PRIM_PATCH) {                // normal = Su x Sv at (u, v), facing the ray
    p = madd(d, t, o);
    n = patch_normal(sc.patch_cp + 16 * hdr.w, __ldg(&sc.prim_b[prim]), hu, hv, d);
  } else if (type == SRT_PRIM_CONSTANT_MEDIUM) {      // geometry.scm:567-571: normal (1,0,0), phase-function material
    p = madd(d, t, o);
    n = v3(1.f, 0.f, 0.f);
  } else {
    p = madd(d, t, o);
    n = -d;
  }
  if ((hdr.x >> 8) & SRT_PRIM_FLAG_FLIP) n = -n;
}
// geometry.scm:138-144 get-sphere-uv (Q5) — only the parity hook needs it (dead in shading).
__device__ __forceinline__ void sphere_uv(float3 p, float& u, float& v) {
  float phi = atan2f(p.z, p.z), theta = asinf(p.y);
  u = 1.0f - (phi + SRT_PI) / (2.0f * SRT_PI);
  v = (theta + 0.5f * SRT_PI) / SRT_PI;
}

// ------------------------------------------------------------------------------------------------
// perlin.scm:51-103 / texture.scm:9-34
__device__ __forceinline__ float perlin_noise(const DScene& sc, float3 p, int quirks) {
  float fi = floorf(p.x), fj = floorf(p.y), fk = floorf(p.z);
  float u = p.x - fi, v = p.y - fj, w = p.z - fk;
  int i = (int)fi, j = (int)fj, k = (int)fk;
  float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
  const bool alias = quirks & SRT_Q4_PERLIN_ALIAS;   // Q4 perlin.scm:76: c[i][j][k] = grad(i+1, j+1, k+dk)
  float acc = 0.0f;
  if (alias) {
    // Under the reference's aliasing the eight corners hold only TWO distinct gradients g0, g1 (dk = 0, 1): 5 table
    // loads per call instead of 32 - the marble texture (7 octaves) goes from 224 to 35 loads per hit - and the sum
    // over the four (a, b) corners collapses: with sum_a w_a = 1 and sum_a w_a a = uu,
    //   sum_{a,b} w_a w_b ((u - a) g.x + (v - b) g.y + (w - c) g.z) = (u - uu) g.x + (v - vv) g.y + (w - c) g.z,
    // so noise = (1 - ww) (..g0, c = 0) + ww (..g1, c = 1): ~15 flops instead of ~80.  Algebraically the reference's
    // trilinear sum (perlin.scm:51-67); the rounding differs at the 1e-7 level (parity bar on textures: 2e-4).
    const int ixy = __ldg(&sc.perm[(i + 1) & 255]) ^ __ldg(&sc.perm[256 + ((j + 1) & 255)]);
    const float3 g0 = xyz(__ldg(&sc.ranvec[ixy ^ __ldg(&sc.perm[512 + (k & 255)])]));
    const float3 g1 = xyz(__ldg(&sc.ranvec[ixy ^ __ldg(&sc.perm[512 + ((k + 1) & 255)])]));
    // every operation explicitly rounded: the function is instantiated in k_shade and in k_tail, whose frames must
    // stay bit-identical - a contraction the compiler is free to make in one of them (u - A * B -> fma) would show
    const float hu = __fmul_rn(__fmul_rn(u, u), __fsub_rn(3.0f, __fmul_rn(2.0f, u)));
    const float hv = __fmul_rn(__fmul_rn(v, v), __fsub_rn(3.0f, __fmul_rn(2.0f, v)));
    const float hw = __fmul_rn(__fmul_rn(w, w), __fsub_rn(3.0f, __fmul_rn(2.0f, w)));
    const float du = __fsub_rn(u, hu), dv = __fsub_rn(v, hv);
    const float n0 = fmaf(w, g0.z, fmaf(dv, g0.y, __fmul_rn(du, g0.x)));
    const float n1 = fmaf(__fsub_rn(w, 1.0f), g1.z, fmaf(dv, g1.y, __fmul_rn(du, g1.x)));
    return fmaf(hw, __fsub_rn(n1, n0), n0);
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        int ei = alias ? 1 : a, ej = alias ? 1 : b;
        int idx = __ldg(&sc.perm[(i + ei) & 255]) ^ __ldg(&sc.perm[256 + ((j + ej) & 255)]) ^ __ldg(&sc.perm[512 + ((k + c) & 255)]);
        float3 g = xyz(__ldg(&sc.ranvec[idx]));
        float wa = a ? uu : 1.0f - uu, wb = b ? vv : 1.0f - vv, wc = c ? ww : 1.0f - ww;
        acc += wa * wb * wc * dot(v3(u - a, v - b, w - c), g);
      }
  return acc;
}
__device__ __forceinline__ float perlin_turb(const DScene& sc, float3 p, int quirks) {   // perlin.scm:92-103, depth 7
  float acc = 0.0f, weight = 1.0f;
  for (int d = 0; d < 7; ++d) { acc = fmaf(weight, perlin_noise(sc, p, quirks), acc); p = p * 2.0f; weight *= 0.5f; }
  return fabsf(acc);
}
__device__ __forceinline__ float3 tex_value(const DScene& sc, int tex, float u, float v, float3 p, int quirks) {
  for (int guard = 0; guard < 64; ++guard) {
    if (!SRT_BOUNDS_OK(tex >= 0 && tex < sc.n_tex, 121)) return v3(0.f, 0.f, 0.f);
    float4 t0 = __ldg(&sc.tex[2 * tex]);
    int kind = __float_as_int(t0.x);
    if (kind == SRT_TEX_CONSTANT) return xyz(__ldg(&sc.tex[2 * tex + 1]));                    // texture.scm:12
    if (kind == SRT_TEX_CHECKER) {                                                           // texture.scm:16-23
      float sines = sinf(10.0f * p.x) * sinf(10.0f * p.y) * sinf(10.0f * p.z);
      tex = (sines < 0.0f) ? __float_as_int(t0.z) : __float_as_int(t0.y);
      continue;
    }
    if (kind == SRT_TEX_NOISE) { float n = perlin_noise(sc, p * t0.w, quirks); return v3(n, n, n); }   // texture.scm:25
    if (kind == SRT_TEX_IMAGE) {                                                             // texture.scm:36-50
      // i = u*nx, j = (1-v)*ny - 0.001, both clamped to [0, n-1]; upstream then indexes its data
      // vector with the (non-integer) result, which Gauche rejects - the texel is the floor.
      const int4 im = __ldg(&sc.imgs[__float_as_int(t0.y)]);
      float fi = u * (float)im.x, fj = (1.0f - v) * (float)im.y - 0.001f;
      fi = fi >= 0.0f ? fi : 0.0f; fj = fj >= 0.0f ? fj : 0.0f;          // also sends NaN (Q5 uv) to texel 0
      fi = fi > (float)(im.x - 1) ? (float)(im.x - 1) : fi; fj = fj > (float)(im.y - 1) ? (float)(im.y - 1) : fj;
      const unsigned char* px = sc.img_texels + (size_t)im.z + 3 * ((size_t)(int)fi + (size_t)im.x * (size_t)(int)fj);
      return v3((float)__ldg(px) / 255.0f, (float)__ldg(px + 1) / 255.0f, (float)__ldg(px + 2) / 255.0f);
    }
    float s = 0.5f * (1.0f + sinf(fmaf(t0.w, p.z, 10.0f * perlin_turb(sc, p, quirks))));                // texture.scm:30
    return v3(s, s, s);
  }
  return v3(0.f, 0.f, 0.f);
}

// main.scm:91-98 sky functions
__device__ __forceinline__ float3 sky_value(int sky, float3 d) {
  if (sky == SRT_SKY_BLACK) return v3(0.f, 0.f, 0.f);
  float3 ud = unit(d);
  float t = 0.5f * (1.0f + ud.y);
  return v3(1.f, 1.f, 1.f) * (1.0f - t) + v3(0.5f, 0.7f, 1.0f) * t;
}

// camera.scm:80-92 get-ray.  Draw slots (bounce 0): block 0 = [xi_u, xi_v, xi_time, -], lens
// disk rejection from block 1 (skipped when lens-radius == 0: the offset is exactly zero).
__device__ __forceinline__ void get_ray(const DCamera& cam, float s, float t, float xi_time, const RngAddr& addr, float3& o, float3& d, float& time) {
  float3 rd = v3(0.f, 0.f, 0.f);
  if (cam.lens_radius != 0.0f) rd = random_in_unit_disk(addr, 1) * cam.lens_radius;
  float3 offset = cam.u * rd.x + cam.v * rd.y;
  time = cam.time0 + xi_time * (cam.time1 - cam.time0);
  o = cam.origin + offset;
  d = cam.llc + cam.horiz * s + cam.vert * t - cam.origin - offset;
}

// ------------------------------------------------------------------------------------------------
// pdf.scm — importance sampling.  make-cosine-pdf (:18-26) and make-mixture-pdf (:34-41) follow
// the source; make-hitable-pdf (:28-32) calls g:pdf-value / g:random which do not exist upstream,
// so the per-shape functions follow "The Rest of Your Life" (PARITY UNPINNED, validated against
// the oracle's identical definition).  Lights are un-instanced rects or spheres.
__device__ __forceinline__ float3 random_to_sphere(float radius, float distance_sq, float r1, float r2) {   // util.scm:46-54
  float z = 1.0f + r2 * (sqrtf(1.0f - (radius * radius) / distance_sq) - 1.0f);
  float s, c; sincospif(2.0f * r1, &s, &c);
  float q = sqrtf(fmaxf(0.0f, 1.0f - z * z));
  return v3(c * q, s * q, z);
}
__device__ __forceinline__ float light_pdf_value(const DScene& sc, int prim, float3 o, float3 v) {
  int type = __ldg(&sc.prim_hdr[prim]).x & 0xff;
  float4 a = __ldg(&sc.prim_a[prim]);
  float t, uu, vv;
  if (type == SRT_PRIM_SPHERE) {
    if (!isect_sphere(xyz(a), a.w, o, v, 1.0f / dot(v, v), 0.001f, t) || !(t < SRT_MAX_FLOAT)) return 0.0f;
    float3 dc = xyz(a) - o;
    float cos_theta_max = sqrtf(1.0f - a.w * a.w / dot(dc, dc));
    return 1.0f / (2.0f * SRT_PI * (1.0f - cos_theta_max));
  }
  float k = __ldg(&sc.prim_b[prim]).x;
  if (!isect_rect(type, a, k, o, v, 0.001f, 0.f, false, t, uu, vv) || !(t <= SRT_MAX_FLOAT)) return 0.0f;
  int axis = (type == SRT_PRIM_XY_RECT) ? 2 : (type == SRT_PRIM_XZ_RECT ? 1 : 0);
  float area = (a.y - a.x) * (a.w - a.z);
  float dist2 = t * t * dot(v, v);
  float cosine = fabsf(cmp3(v, axis)) / length(v);
  return dist2 / (cosine * area);
}
__device__ __forceinline__ float3 light_random(const DScene& sc, int prim, float3 o, float xa, float xb) {
  int type = __ldg(&sc.prim_hdr[prim]).x & 0xff;
  float4 a = __ldg(&sc.prim_a[prim]);
  if (type == SRT_PRIM_SPHERE) {
    float3 dc = xyz(a) - o;
    float3 w = unit(dc);
    float3 ax = (fabsf(w.x) > 0.9f) ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
    float3 vv = unit(cross(w, ax)), uu = cross(w, vv);
    float3 r = random_to_sphere(a.w, dot(dc, dc), xa, xb);
    return uu * r.x + vv * r.y + w * r.z;
  }
  float k = __ldg(&sc.prim_b[prim]).x;
  float pa = a.x + xa * (a.y - a.x), pb = a.z + xb * (a.w - a.z);
  float3 pt = type == SRT_PRIM_XY_RECT ? v3(pa, pb, k) : (type == SRT_PRIM_XZ_RECT ? v3(pa, k, pb) : v3(k, pa, pb));
  return pt - o;
}
__device__ __forceinline__ float lights_pdf_value(const DScene& sc, float3 o, float3 v) {   // plain average over the lights
  float sum = 0.0f;
  for (int j = 0; j < sc.n_lights; ++j) sum += light_pdf_value(sc, __ldg(&sc.lights[j]), o, v);
  return sum / (float)sc.n_lights;
}

// ------------------------------------------------------------------------------------------------
// material.scm — scatter.  Returns true when the path continues; `weight` multiplies throughput.
__device__ __forceinline__ float3 reflect(float3 v, float3 n) { return v - n * (2.0f * dot(v, n)); }   // material.scm:41
__device__ __forceinline__ float schlick(float cosine, float ref_idx) {                                 // material.scm:69
  float r0 = (1.0f - ref_idx) / (1.0f + ref_idx); r0 = r0 * r0;
  float m = 1.0f - cosine, m2 = m * m;
  return r0 + (1.0f - r0) * (m2 * m2 * m);
}
struct Scatter { float3 dir; float3 weight; float3 emitted; bool valid; };

template <int EST>
__device__ __forceinline__ Scatter scatter(const DScene& sc, int prim, float3 d_in, float3 p, float3 n, float u, float v,
                                           const RngAddr& addr, int quirks) {
  Scatter r; r.valid = false; r.emitted = v3(0.f, 0.f, 0.f); r.weight = v3(1.f, 1.f, 1.f); r.dir = v3(0.f, 1.f, 0.f);
  const float4 sh0 = __ldg(&sc.prim_shade[2 * prim]), sh1 = __ldg(&sc.prim_shade[2 * prim + 1]);
  const int4 m = make_int4(__float_as_int(sh1.x), __float_as_int(sh1.y), 0, 0);     // (material kind, texture id)
  const float param = sh0.w;
  // Shared by every material, so done ONCE before the switch with all hit lanes active instead of
  // once per divergent branch: the first Philox block of this bounce and the texture lookup
  // ((t:value albedo 0 0 p) for lambertian / metal, (t:value emit u v p) for lights).
  const float4 xi = rng_block(addr, 0);
  const bool uv_tex = m.x == SRT_MAT_DIFFUSE_LIGHT || m.x == SRT_MAT_ISOTROPIC;
  float3 tex = xyz(sh0);                                        // constant textures: colour already in the record
  if (m.x != SRT_MAT_DIELECTRIC && __float_as_int(sh1.z) == 0) tex = tex_value(sc, m.y, uv_tex ? u : 0.0f, uv_tex ? v : 0.0f, p, quirks);
  switch (m.x) {
    case SRT_MAT_LAMBERTIAN: {                                   // material.scm:24-39 + onb.scm:8-16
      float len_n = length(n);
      float3 w = n * (1.0f / len_n);
      float3 a = (fabsf(w.x) > 0.9f) ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
      float3 vv = unit(cross(w, a));
      float3 uu = cross(w, vv);
      float3 rc = random_cosine_direction(xi.x, xi.y, quirks);
      if (quirks & SRT_Q15_LOCAL_TRIPLE_EVAL) {
        // Q15, onb.scm:27-36: `local` expands its operand three times, so the reference draws three cosine
        // directions and keeps x of the first (above), y of the second and z of the third.  Draw slots:
        // block 2 = (r1', r2', -, r2'') — the third evaluation's r1 only feeds its discarded x, y.
        const float4 xk = rng_block(addr, 2);
        float s2, c2;
        sincospif(2.0f * xk.x, &s2, &c2);
        rc.y = s2 * (((quirks & SRT_Q1_COSINE_X2) ? 2.0f : 1.0f) * sqrtf(xk.y));
        rc.z = sqrtf(1.0f - xk.w);
      }
      if (EST == SRT_EST_MIXTURE && sc.n_lights > 0) {
        // mixture(hittable(lights), cosine) pdf.scm:34-41; block 0 = (r1, r2, xi_choice, xi_light), block 1 = (xa, xb)
        float4 xj = rng_block(addr, 1);
        float3 dir;
        if (xi.z < 0.5f) {
          int li = min((int)(xi.w * (float)sc.n_lights), sc.n_lights - 1);
          dir = light_random(sc, __ldg(&sc.lights[li]), p, xj.x, xj.y);
        } else {
          dir = uu * rc.x + vv * rc.y + w * rc.z;
        }
        float cosw = dot(unit(dir), w);
        float pdf_val = 0.5f * lights_pdf_value(sc, p, dir) + 0.5f * (cosw > 0.0f ? cosw / SRT_PI : 0.0f);
        float spdf = fmaxf(0.0f, dot(n, unit(dir))) / SRT_PI;
        r.dir = dir;
        r.valid = spdf > 0.0f && pdf_val > 0.0f;             // zero-weight paths end here
        if (r.valid) r.weight = tex * spdf * (1.0f / pdf_val);
        break;
      }
      float3 target = uu * rc.x + vv * rc.y + w * rc.z;
      r.dir = unit(target);
      float pdf_cos = dot(w, r.dir);                              // pdf  = (w . dir)/pi
      float spdf_cos = fmaxf(0.0f, dot(n, r.dir));                // spdf = max(0, n . dir)/pi   (material.scm:33-36)
      r.weight = tex * spdf_cos * (1.0f / pdf_cos);               // main.scm:113-118
      r.valid = true;
      break;
    }
    case SRT_MAT_METAL: {                                        // material.scm:45-57 (specular: weight = albedo)
      float3 refl = reflect(unit(d_in), n);
      float3 f = v3(2.0f * xi.x - 1.0f, 2.0f * xi.y - 1.0f, 2.0f * xi.z - 1.0f);   // util.scm:9-15, iteration 0 = block 0
      if (!(dot(f, f) < 1.0f)) f = random_in_unit_sphere(addr, 1);
      r.dir = refl + f * param;
      r.valid = dot(r.dir, n) > 0.0f;
      r.weight = tex;
      break;
    }
    case SRT_MAT_DIELECTRIC: {                                   // material.scm:76-101 (Q10)
      float ref_idx = param;
      float3 din = (quirks & SRT_Q10_DIELECTRIC_UNNORM) ? d_in : unit(d_in);
      float3 refl = reflect(din, n);
      float dd = dot(din, n);
      float3 outward = (dd > 0.0f) ? -n : n;
      float ni_over_nt = (dd > 0.0f) ? ref_idx : 1.0f / ref_idx;
      float len_d = length(din);
      float cosine = (dd > 0.0f) ? (dd * ref_idx) / len_d : (-dd) / len_d;
      float3 uv_ = din * (1.0f / len_d);                          // refract material.scm:59-67
      float dt = dot(uv_, outward);
      float disc = 1.0f - ni_over_nt * ni_over_nt * (1.0f - dt * dt);
      float reflect_prob = 1.0f; float3 refr = refl;
      if (disc > 0.0f) {
        float3 vv = (quirks & SRT_Q10_DIELECTRIC_UNNORM) ? din : uv_;
        refr = (vv - outward * dt) * ni_over_nt - outward * sqrtf(disc);
        reflect_prob = schlick(cosine, ref_idx);
      }
      r.dir = (xi.x < reflect_prob) ? refl : refr;
      r.valid = true;
      break;
    }
    case SRT_MAT_DIFFUSE_LIGHT: {                                // material.scm:103-111
      if (dot(n, d_in) < 0.0f) r.emitted = tex;
      break;
    }
    case SRT_MAT_ISOTROPIC: {                                    // absent upstream; book semantics
      float3 f = v3(2.0f * xi.x - 1.0f, 2.0f * xi.y - 1.0f, 2.0f * xi.z - 1.0f);
      if (!(dot(f, f) < 1.0f)) f = random_in_unit_sphere(addr, 1);
      r.dir = f;
      r.weight = tex;
      r.valid = true;
      break;
    }
  }
  return r;
}
