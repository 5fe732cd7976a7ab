// wavefront.cu — the per-sample radiance loop as a wavefront path tracer.
//
// Replaces trace-all -> get-ray -> color -> g:hit / m:scatter / m:emitted (main.scm:100-121,
// 471-491; camera.scm:80-92; geometry.scm:14-15; material.scm:15-22).  STREAMING wavefront: one
// queue of up to `capacity` paths is kept full; every iteration runs
//   k_extend  closest hit: LBVH traversal (near child first, per-thread stack of 16-bit node ids in
//             shared memory - 32-bit for trees of >= 65536 nodes), the whole node array (as four
//             child-interleaved planes: FFMA2 slab test of both children) + primitive headers staged
//             in shared memory
//   k_shade   hit-record completion, material scatter / emitted, sky on miss; terminated paths add
//             their radiance into a 64-bit fixed-point accumulator; survivors are compacted
//             (warp ballot + one atomic per CTA) into the other queue generation
//   k_regen   fresh camera paths (Philox bounce slot 0) are appended behind the survivors
// so the long thin tail of deep paths (depth <= max_depth) is paid once per frame, not once per
// batch of samples.  Iterations are enqueued in batches of 8 that are captured once into a CUDA
// graph (cached on the scene) and replayed; the host polls the control block one batch behind.
//   k_tail    once every path is generated and the queue is small (or has stopped shrinking: the
//             long-lived glass paths), ONE launch runs the remaining paths to their end
// The recursive estimator `color` is run in its iterative form
// L = sum_k (prod_{j<k} w_j) e_k; paths carry (throughput, pixel, sample, depth).
#include <mutex>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include "srt_device.cuh"
#include "srt_host.h"

namespace {

#ifndef SRT_HEAVY_THREADS
#define SRT_HEAVY_THREADS 512
#endif
constexpr int EXT_THREADS = 256, EXT_THREADS_HEAVY = SRT_HEAVY_THREADS;
#ifndef SRT_SHD_THREADS
#define SRT_SHD_THREADS 256
#endif
constexpr int SHD_THREADS = SRT_SHD_THREADS;

// ------------------------------------------------------------------------------------------------
// Queue control block (device memory).  Fields are double-buffered by queue generation /
// iteration parity so that no kernel reads a field another CTA of the same launch writes.
//   qcount[g]     live length of queue generation g (read by extend / shade)
//   survivors[g]  compaction cursor: shade(g^1) atomically appends survivors into generation g
//   next_path[k]  first camera path not yet generated, as seen by the regen of parity k
//   spp_begin     first sample of the range this call renders (kept here, not in the kernel
//                 arguments, so that the captured iteration graph is reusable across calls)
struct WaveCtrl {
  int qcount[2]; int survivors[2];
  unsigned long long next_path[2];
  unsigned long long total_paths;
  unsigned long long rays;          // sum of qcount over iterations = closest-hit queries
  unsigned long long iterations;
  unsigned long long nonfinite;     // radiance contributions dropped because they were NaN / Inf
  unsigned long long bounce_hist[8];// closest-hit queries by bounce (0..6, >= 7), profile mode only
  int spp_begin; int tail_runs;     // tail_runs: how many times the one-launch tail kernel took over
};
#define SRT_ACC_SCALE 68719476736.0f   // 2^36: radiance accumulates in 64-bit fixed point
#define SRT_ACC_MAX 6.7e7f             // per-contribution clamp: 2^26 * 2^36 = 2^62 stays inside the signed 64-bit sum

// Order-independent (integer) accumulation: WITHIN ONE CALL the image is bit-identical for any
// scheduling, queue size, number of GPUs (the multi-GPU reduce adds these integers) and sample-range
// split; across calls the float running sum rgb_sum += ... rounds once per call.  main.scm:480 running sum.
// A NaN / Inf contribution (outside the reference's domain: Gauche would carry the NaN into the
// pixel) is dropped and counted in SrtStats.nonfinite instead of poisoning the integer sum; a finite
// one is clamped to SRT_ACC_MAX, so a call's per-pixel sum is exact up to 2^27 (spp x radiance).
__device__ __forceinline__ void accumulate_fixed(unsigned long long* __restrict__ acc, int pixel, float3 L, unsigned long long* __restrict__ nonfinite) {
  if (!(fabsf(L.x) <= 3.0e38f) || !(fabsf(L.y) <= 3.0e38f) || !(fabsf(L.z) <= 3.0e38f)) { atomicAdd(nonfinite, 1ull); return; }
  L.x = fminf(fmaxf(L.x, -SRT_ACC_MAX), SRT_ACC_MAX); L.y = fminf(fmaxf(L.y, -SRT_ACC_MAX), SRT_ACC_MAX); L.z = fminf(fmaxf(L.z, -SRT_ACC_MAX), SRT_ACC_MAX);
  if (L.x != 0.f) atomicAdd(&acc[3 * (size_t)pixel + 0], (unsigned long long)__float2ll_rn(L.x * SRT_ACC_SCALE));
  if (L.y != 0.f) atomicAdd(&acc[3 * (size_t)pixel + 1], (unsigned long long)__float2ll_rn(L.y * SRT_ACC_SCALE));
  if (L.z != 0.f) atomicAdd(&acc[3 * (size_t)pixel + 2], (unsigned long long)__float2ll_rn(L.z * SRT_ACC_SCALE));
}

// ------------------------------------------------------------------------------------------------
#ifndef SRT_TILE_LOG_W
#define SRT_TILE_LOG_W 4
#define SRT_TILE_LOG_H 4
#endif
// regen: tops queue generation g up with fresh camera paths (main.scm:476-478 + camera.scm:80-92).
// Path id -> (pixel, sample): see "Path order" below.  Path state: ray_o = (o, time), ray_d = (d, sample << 12 | depth),
// state = (throughput, pixel).
__global__ void __launch_bounds__(256) k_regen(DCamera cam, SrtRenderParams p, int npix, int capacity, int g, int parity,
                                                float4* __restrict__ ray_o, float4* __restrict__ ray_d, float4* __restrict__ state,
                                                WaveCtrl* __restrict__ ctrl) {
  const int surv = ctrl->survivors[g];
  const unsigned long long next = ctrl->next_path[parity], total = ctrl->total_paths;
  const int spp_begin = ctrl->spp_begin;
  unsigned long long room = (unsigned long long)(capacity - surv), left = total - next;
  const int n_new = (int)(room < left ? room : left);
  // Path order.  Frames that divide into 16 x 16 pixel tiles: TILE-major - all samples of a tile before the next tile, inside a
  // sample the tile's 256 pixels as eight 8 x 4 blocks - so a warp of fresh paths is an 8 x 4 block (not a 32 x 1 run) and,
  // more important, what survives many bounces next to each other in the queue still comes from the same tile (with sample-major
  // order the 32 survivors of a deep-bounce warp were spread over several tiles).  Other frames: sample-major scanlines.  The
  // image does not depend on the order (Philox is keyed by (pixel, sample, bounce), the sums are integers).
  // path id -> (outer, inner) = (tile, sample * 256 + pixel in tile) or (sample, pixel) needs a 64-bit division (ids pass 2^32 on
  // cfg4 / cfg5): done ONCE per thread, the grid-stride loop then advances by the precomputed quotient / remainder of its stride
  constexpr int TLW = SRT_TILE_LOG_W, TLH = SRT_TILE_LOG_H, TLP = TLW + TLH;          // tile = 2^TLW x 2^TLH pixels
  const bool tiled = ((p.width & ((1 << TLW) - 1)) | (p.height & ((1 << TLH) - 1))) == 0;
  const unsigned int n_samples = (unsigned int)(total / (unsigned long long)npix);
  const unsigned int span = max(tiled ? n_samples << TLP : (unsigned int)npix, 1u);          // paths per outer unit
  const int stride = gridDim.x * blockDim.x;
  const unsigned int s_q = (unsigned int)stride / span, s_r = (unsigned int)stride % span;
  const int nbx = p.width >> TLW;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int outer = 0u, inner = 0u;
  if (j < n_new) {
    const unsigned long long id = next + (unsigned long long)j;
    outer = (unsigned int)(id / (unsigned long long)span);
    inner = (unsigned int)(id - (unsigned long long)outer * (unsigned long long)span);
  }
  for (; j < n_new; j += stride, outer += s_q, inner += s_r) {
    if (inner >= span) { inner -= span; ++outer; }
    int x, y; unsigned int sl;
    if (tiled) {
      const int r = (int)(inner & ((1u << TLP) - 1u)), by = (int)outer / nbx, bx = (int)outer - by * nbx;
      const int blk = r >> 5, l = r & 31, bxi = blk & ((1 << (TLW - 3)) - 1), byi = blk >> (TLW - 3);     // 8 x 4 blocks, row-major inside the tile
      sl = inner >> TLP;
      x = (bx << TLW) + bxi * 8 + (l & 7); y = (by << TLH) + byi * 4 + (l >> 3);
    } else { sl = outer; y = (int)inner / p.width; x = (int)inner - y * p.width; }
    const int pixel = y * p.width + x;
    unsigned int sample = (unsigned int)spp_begin + sl;
    RngAddr addr{p.seed, (uint32_t)pixel, sample, 0u};
    float4 xi = rng_block(addr, 0);
    float u = ((float)x + xi.x) / (float)p.width;       // y = 0 is the bottom row
    float v = ((float)y + xi.y) / (float)p.height;
    float3 o, d; float time;
    get_ray(cam, u, v, xi.z, addr, o, d, time);
    int slot = surv + j;
    if (!SRT_BOUNDS_OK(slot >= 0 && slot < capacity, 211)) continue;
    ray_o[slot] = make_float4(o.x, o.y, o.z, time);
    ray_d[slot] = make_float4(d.x, d.y, d.z, __int_as_float((int)(sample << 12)));
    state[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(pixel));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int q = surv + n_new;
    ctrl->qcount[g] = q;
    ctrl->next_path[parity ^ 1] = next + (unsigned long long)n_new;
    ctrl->survivors[g ^ 1] = 0;          // cursor of the generation the next shade appends into
    ctrl->rays += (unsigned long long)q;
    ctrl->iterations += q ? 1ull : 0ull;
  }
}

// ------------------------------------------------------------------------------------------------
// extend: closest-hit traversal.  Node = 64 B: left box, right box, (left, right, parent,
// sibling).  Pending far children: see the TRAV_* modes at node_step (shared-memory stack by
// default; the stackless fallback keeps a bit trail - bit k from the LSB = "the far child at the
// k-th level above is still pending" - and follows parent links up to the lowest set bit, then
// enters the sibling).  No local memory in either mode.
// staged primitive tables, addressed through 32-bit shared-space addresses (through generic pointers
// the compiler re-derives the shared window - S2UR SR_CgaCtaId + LEA - at every leaf test)
struct PrimShared {
  uint32_t h, pa;
  __device__ __forceinline__ PrimShared(const int4* hp, const float4* ap) : h((uint32_t)__cvta_generic_to_shared(hp)), pa((uint32_t)__cvta_generic_to_shared(ap)) {
    asm volatile("" : "+r"(h), "+r"(pa));    // opaque: held in registers, not re-derived per test
  }
  __device__ __forceinline__ int4 hdr(int i) const { int4 r; asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(h + 16u * (uint32_t)i)); return r; }
  __device__ __forceinline__ float4 a(int i) const { float4 r; asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(pa + 16u * (uint32_t)i)); return r; }
};
struct PrimGlobal {
  const int4* h; const float4* pa;
  __device__ __forceinline__ int4 hdr(int i) const { return __ldg(&h[i]); }
  __device__ __forceinline__ float4 a(int i) const { return __ldg(&pa[i]); }
};

// Per-lane traversal state of the persistent extend kernel.
struct Trav {
  float3 o, d, inv, oi, ainv; float time, inv_a;
  Hit h; int node; unsigned long long trail; int ray;
  RngAddr ra;                                // only read by constant-medium primitives
  unsigned long long s0, s1; int nstk;      // register-packed cache of the 8 most recent pending far children (16-bit ids)
  uint32_t sp, sp0;                         // TRAV_STACK: shared-space address of the next free / first stack slot
  int n_nodes;                              // plane stride of the staged node layout
#ifdef SRT_BOUNDS_CHECK
  int stack_slots;                          // halfwords of this thread's stack (bounds-checked build)
#endif
#ifdef SRT_COUNT_STEPS
  int nsteps, ntests, nmiss, maxsp;         // instrumented build only (tools/step_stats.py); maxsp = deepest stack use
#endif
};
__device__ __forceinline__ void trav_init(Trav& T, float4 o4, float4 d4, float tmax, int ray) {
  T.o = xyz(o4); T.d = xyz(d4); T.time = o4.w; T.ray = ray;
  // reciprocal direction; (near-)zero components become +-1e18 instead of +-inf so that the FMA
  // slab form never produces inf - inf (boxes are padded, see lbvh.cu, so the sign of
  // (b - o) * 1e18 is exact for every ray that can reach a primitive inside the box)
  // (__frcp_rn is the correctly rounded reciprocal, i.e. the value of 1.0f / x, without the division's slow-path check)
  T.inv = v3(fabsf(T.d.x) > 1e-18f ? __frcp_rn(T.d.x) : copysignf(1e18f, T.d.x), fabsf(T.d.y) > 1e-18f ? __frcp_rn(T.d.y) : copysignf(1e18f, T.d.y),
             fabsf(T.d.z) > 1e-18f ? __frcp_rn(T.d.z) : copysignf(1e18f, T.d.z));
  T.oi = v3(T.o.x * T.inv.x, T.o.y * T.inv.y, T.o.z * T.inv.z);
  T.ainv = v3(fabsf(T.inv.x), fabsf(T.inv.y), fabsf(T.inv.z));
  T.inv_a = __frcp_rn(dot(T.d, T.d));
  T.h.t = tmax; T.h.prim = -1; T.h.u = 0.f; T.h.v = 0.f; T.h.incl = false;
  T.node = 0; T.trail = 0ull; T.s0 = T.s1 = 0ull; T.nstk = 0;      // T.sp: back above the sentinel after every finished traversal
#ifdef SRT_COUNT_STEPS
  T.nsteps = 0; T.ntests = 0; T.nmiss = 0; T.maxsp = 0;
#endif
}

// One node step: slab-test both children, stash hit leaf children in (pend0, pend1), descend into
// the near internal child or backtrack.  Returns false when the traversal is finished.
//
// Slab test on centre / half-extent boxes: t_c = c*inv - o*inv, t_near/far = t_c -/+ e*|inv| —
// 9 FFMA + 4 min/max per box (the min/max-heavy 6 FFMA + 10 FMNMX form saturates the ALU pipe while
// the FMA pipe idles).
//
// Backtracking is stackless in the sense of Barringer & Akenine-Moller: the bit trail says at
// which level above a far child is still pending, parent links lead there and the sibling link
// enters it.  Because that climb is a serial chain of dependent loads executed by few lanes, the
// 8 most recent far children are additionally cached in two 64-bit registers (16-bit ids,
// CACHE = true when the tree has < 65536 nodes); the climb only runs when the cache has overflowed.
// 16-byte read of LBVH word `idx` (4 per node).  Shared-memory trees are addressed through a
// 32-bit shared-space base computed once per thread: through the generic pointer the compiler
// re-derives the shared window (S2R SR_CgaCtaId + LEA) in every iteration of the node loop.
template <bool SMEM>
__device__ __forceinline__ float4 ld_node(const float4* __restrict__ nodes, uint32_t sbase, int idx) {
  if (SMEM) {
    float4 r;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sbase + 16u * (uint32_t)idx));
    return r;
  }
  return __ldg(&nodes[idx]);
}

// STAGED node layout (shared memory only; the global array keeps the 64-byte SrtBvhNode the bit-exact
// LBVH check reads back).  Word-major ("SoA") and child-interleaved:
//   plane 0 [node] = (lc.x rc.x lc.y rc.y)   plane 1 = (lc.z rc.z le.x re.x)
//   plane 2 [node] = (le.y re.y le.z re.z)   plane 3 = (left right parent sibling)
// Word-major: the lanes of a quarter-warp read 16-byte words of DIFFERENT nodes; as 64-byte records all
// of them fall on 2 of the 8 16-byte bank groups (ncu r1: 31 % of the shared wavefronts were conflict
// replays), as planes they spread over all 8 by node id.  Child-interleaved: (left, right) of one
// coordinate sit in an aligned register pair, which is what the packed FFMA2 of sm_100 takes.
#ifndef SRT_NODE_AOS
#define SRT_NODE_PLANES 1
#else
#define SRT_NODE_PLANES 0
#endif
__device__ __forceinline__ void stage_nodes(float4* __restrict__ smem, const float4* __restrict__ nodes, int n_nodes) {
#if SRT_NODE_PLANES
  for (int i = threadIdx.x; i < n_nodes; i += blockDim.x) {
    const float4 a = nodes[4 * i], b = nodes[4 * i + 1], c = nodes[4 * i + 2], d = nodes[4 * i + 3];
    // global: a = (lc.x lc.y lc.z le.x) b = (le.y le.z rc.x rc.y) c = (rc.z re.x re.y re.z)
    smem[i] = make_float4(a.x, b.z, a.y, b.w);
    smem[n_nodes + i] = make_float4(a.z, c.x, a.w, c.y);
    smem[2 * n_nodes + i] = make_float4(b.x, c.z, b.y, c.w);
    smem[3 * n_nodes + i] = d;
  }
#else
  for (int i = threadIdx.x; i < 4 * n_nodes; i += blockDim.x) smem[i] = nodes[i];
#endif
}
// packed fp32x2 FMA of sm_100 (FFMA2): (a.x, a.y) * b + (c.x, c.y), b broadcast; one issue slot for two FMAs
__device__ __forceinline__ float2 ffma2(float2 a, float b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
  return r;
}
__device__ __forceinline__ float2 ffma2s(float2 a, float b, float c) { return ffma2(a, b, make_float2(c, c)); }

// Backtracking modes (template parameter TRAV):
//  TRAV_STACK  per-thread stack of 16-bit node ids in shared memory, bvh_depth + 1 entries (a thread
//              never has more far children pending than internal nodes on its root path): push =
//              STS.U16, pop = LDS.U16.  Used whenever the tree has < 65536 nodes and the stack fits.
//  TRAV_CACHE  stackless bit trail + parent/sibling climb, with the 8 most recent far children in
//              two 64-bit registers.
//  TRAV_STACK32 the same stack with 32-bit entries, for trees with >= 65536 nodes (never staged, so the stack
//              always fits): replaced the bare bit trail, whose parent / sibling climb is a chain of dependent
//              GLOBAL loads on such trees (70,000-sphere cloud: profiles/README.md, round 2).
enum { TRAV_STACK32 = 0, TRAV_CACHE = 1, TRAV_STACK = 2 };
// halfwords per thread, odd (spreads banks): bvh_depth + 1 pending far children + the SENTINEL 0xffff at
// the bottom, written once per thread: "pop" needs no compare against the stack base, the empty stack
// answers with the sentinel (node ids are < 65535 in this mode)
__host__ __device__ __forceinline__ int trav_stack_stride(int bvh_depth) { return (bvh_depth + 2) | 1; }
#define SRT_STACK_SENTINEL 0xffffu
__device__ __forceinline__ uint32_t trav_stack_init(uint32_t base) {
  asm volatile("st.shared.u16 [%0], %1;" :: "r"(base), "h"((unsigned short)SRT_STACK_SENTINEL) : "memory");
  return base + 2u;
}
__device__ __forceinline__ uint32_t trav_stack32_init(uint32_t base) {       // TRAV_STACK32: words, sentinel 0xffffffff
  asm volatile("st.shared.u32 [%0], %1;" :: "r"(base), "r"(0xffffffffu) : "memory");
  return base + 4u;
}

// WANT_T: also returns the box intervals [t0, t1] of the stashed leaf children (ivl = {pend0's t0, t1,
// pend1's t0, t1}; already clipped to [tmin, closest hit so far]) for the patch variants' slab cull.
template <bool SMEM, int TRAV, bool WANT_T = false>
__device__ __forceinline__ bool node_step(Trav& T, const float4* __restrict__ nodes, uint32_t sbase, float tmin, int& pend0, int& pend1, float4* ivl = nullptr) {
  const int node = T.node;
  if (!SRT_BOUNDS_OK(node >= 0 && node < T.n_nodes, 231)) return false;
  if (TRAV != TRAV_CACHE && !SRT_BOUNDS_OK(T.sp >= T.sp0 && T.sp <= T.sp0 + (TRAV == TRAV_STACK32 ? 4u : 2u) * (uint32_t)(T.stack_slots - 1), 232)) return false;
#ifdef SRT_COUNT_STEPS
  T.nsteps++;
#endif
  const float3 inv = T.inv, oi = T.oi, ai = T.ainv;
  float lt0, lt1, rt0, rt1; float4 n3;
  if (SMEM && SRT_NODE_PLANES) {
    // staged planes: both children of a coordinate in one register pair, 9 FFMA2 for the two slab tests
    const int N = T.n_nodes;
    const float4 w0 = ld_node<true>(nodes, sbase, node), w1 = ld_node<true>(nodes, sbase, N + node), w2 = ld_node<true>(nodes, sbase, 2 * N + node);
    n3 = ld_node<true>(nodes, sbase, 3 * N + node);
    const float2 cx = ffma2s(make_float2(w0.x, w0.y), inv.x, -oi.x), cy = ffma2s(make_float2(w0.z, w0.w), inv.y, -oi.y), cz = ffma2s(make_float2(w1.x, w1.y), inv.z, -oi.z);
    const float2 ex = make_float2(w1.z, w1.w), ey = make_float2(w2.x, w2.y), ez = make_float2(w2.z, w2.w);
    const float2 nx = ffma2(ex, -ai.x, cx), ny = ffma2(ey, -ai.y, cy), nz = ffma2(ez, -ai.z, cz);
    const float2 fx = ffma2(ex, ai.x, cx), fy = ffma2(ey, ai.y, cy), fz = ffma2(ez, ai.z, cz);
    lt0 = fmaxf(fmaxf(nx.x, ny.x), fmaxf(nz.x, tmin)); lt1 = fminf(fminf(fx.x, fy.x), fminf(fz.x, T.h.t));
    rt0 = fmaxf(fmaxf(nx.y, ny.y), fmaxf(nz.y, tmin)); rt1 = fminf(fminf(fx.y, fy.y), fminf(fz.y, T.h.t));
  } else {
    const float4 n0 = ld_node<SMEM>(nodes, sbase, 4 * node), n1 = ld_node<SMEM>(nodes, sbase, 4 * node + 1), n2 = ld_node<SMEM>(nodes, sbase, 4 * node + 2);
    n3 = ld_node<SMEM>(nodes, sbase, 4 * node + 3);
    // left: c = (n0.x n0.y n0.z) e = (n0.w n1.x n1.y); right: c = (n1.z n1.w n2.x) e = (n2.y n2.z n2.w)
    float lcx = fmaf(n0.x, inv.x, -oi.x), lcy = fmaf(n0.y, inv.y, -oi.y), lcz = fmaf(n0.z, inv.z, -oi.z);
    float rcx = fmaf(n1.z, inv.x, -oi.x), rcy = fmaf(n1.w, inv.y, -oi.y), rcz = fmaf(n2.x, inv.z, -oi.z);
    lt0 = fmaxf(fmaxf(fmaf(-n0.w, ai.x, lcx), fmaf(-n1.x, ai.y, lcy)), fmaxf(fmaf(-n1.y, ai.z, lcz), tmin));
    lt1 = fminf(fminf(fmaf(n0.w, ai.x, lcx), fmaf(n1.x, ai.y, lcy)), fminf(fmaf(n1.y, ai.z, lcz), T.h.t));
    rt0 = fmaxf(fmaxf(fmaf(-n2.y, ai.x, rcx), fmaf(-n2.z, ai.y, rcy)), fmaxf(fmaf(-n2.w, ai.z, rcz), tmin));
    rt1 = fminf(fminf(fmaf(n2.y, ai.x, rcx), fmaf(n2.z, ai.y, rcy)), fminf(fmaf(n2.w, ai.z, rcz), T.h.t));
  }
  const int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
  const bool hl0 = lt0 <= lt1, hr0 = rt0 <= rt1;
  // hit leaf children go to (pend0, pend1) as RAW leaf references (negative: ~primitive id; 0 = none) - select
  // form: as branches this was a divergent region plus byte-packed booleans carried across it.  The node loop
  // is bound by the ALU pipe (compares, selects, min / max: ncu r2), so the logic below is written for
  // predicates: no boolean is materialised in a register, the leaf reference is decoded at the leaf test.
  const bool ll = hl0 & (left < 0), rl = hr0 & (right < 0);
  pend0 = ll ? left : (rl ? right : 0);
  pend1 = (ll & rl) ? right : 0;
  if (WANT_T) *ivl = make_float4(ll ? lt0 : rt0, ll ? lt1 : rt1, rt0, rt1);
  const bool hl = hl0 & (left >= 0), hr = hr0 & (right >= 0);
  if (hl | hr) {
    const bool both = hl & hr;
    const bool right_first = !hl | (hr & (rt0 < lt0));          // near child first; a single hit child is "near"
    const int nearc = right_first ? right : left, farc = right_first ? left : right;
    T.node = nearc;
    if (TRAV == TRAV_STACK32) {
      if (both) {
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(T.sp), "r"((uint32_t)farc) : "memory");
        T.sp += 4u;
      }
      return true;
    }
    if (TRAV == TRAV_STACK) {
      if (both) {
        asm volatile("st.shared.u16 [%0], %1;" :: "r"(T.sp), "h"((unsigned short)farc) : "memory");
        T.sp += 2u;
#ifdef SRT_COUNT_STEPS
        T.maxsp = max(T.maxsp, (int)((T.sp - T.sp0) >> 1));
#endif
      }
      return true;
    }
    const bool go_left = !right_first;
    T.trail = (T.trail << 1) | (both ? 1ull : 0ull);
    if (TRAV == TRAV_CACHE && both) {
      unsigned long long far_id = (unsigned long long)(unsigned)(go_left ? right : left);
      T.s1 = (T.s1 << 16) | (T.s0 >> 48);
      T.s0 = (T.s0 << 16) | far_id;
      T.nstk = min(T.nstk + 1, 8);
    }
    return true;
  }
  if (TRAV == TRAV_STACK32) {
    uint32_t id;
    asm volatile("ld.shared.u32 %0, [%1+-4];" : "=r"(id) : "r"(T.sp) : "memory");
    if (id == 0xffffffffu) return false;
    T.sp -= 4u;
    T.node = (int)id;
    return true;
  }
  if (TRAV == TRAV_STACK) {
    unsigned short id;
    asm volatile("ld.shared.u16 %0, [%1+-2];" : "=h"(id) : "r"(T.sp) : "memory");
    if (id == SRT_STACK_SENTINEL) return false;              // empty: sp stays just above the sentinel, ready for the next ray
    T.sp -= 2u;
    T.node = (int)id;
    return true;
  }
  if (T.trail == 0ull) return false;
  int up = __ffsll((long long)T.trail) - 1;                   // levels up to the pending far child
  T.trail = (T.trail >> up) ^ 1ull;
  if (TRAV == TRAV_CACHE && T.nstk > 0) {
    T.node = (int)(T.s0 & 0xffffull);
    T.s0 = (T.s0 >> 16) | (T.s1 << 48);
    T.s1 >>= 16;
    T.nstk -= 1;
    return true;
  }
  int par = __float_as_int(n3.z), sib = __float_as_int(n3.w);
  for (int k = 0; k < up; ++k) {
    const float4 m = (SMEM && SRT_NODE_PLANES) ? ld_node<true>(nodes, sbase, 3 * T.n_nodes + par) : ld_node<SMEM>(nodes, sbase, 4 * par + 3);
    par = __float_as_int(m.z); sib = __float_as_int(m.w);
  }
  T.node = sib;
  return true;
}

// One ray per lane, 32 consecutive rays per warp (static assignment: dynamic ray fetch and a
// while-while loop were both measured slower at 32 resident warps/SM, see profiles/README.md).
template <bool SMEM, int MASK, int TRAV, class PrimSrc>
__device__ __forceinline__ void extend_loop(const DScene& sc, const float4* __restrict__ nodes, const PrimSrc& ps,
                                            const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ state,
                                            float4* __restrict__ hit, int count, float tmin, float tmax, uint32_t seed, uint32_t stack_base) {
  if (sc.n_surf == 0) {   // empty scene: every ray misses
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) hit[i] = make_float4(tmax, __int_as_float(-1), 0.f, 0.f);
    return;
  }
  Trav T;
  T.n_nodes = sc.n_nodes;
#ifdef SRT_BOUNDS_CHECK
  T.stack_slots = trav_stack_stride(sc.bvh_depth) - 1;     // slots above the sentinel
#endif
  T.sp0 = stack_base + (TRAV == TRAV_STACK32 ? 4u : 2u) * (uint32_t)trav_stack_stride(sc.bvh_depth) * threadIdx.x;
  T.sp = T.sp0;
  if (TRAV == TRAV_STACK) T.sp = T.sp0 = trav_stack_init(T.sp0);
  if (TRAV == TRAV_STACK32) T.sp = T.sp0 = trav_stack32_init(T.sp0);
  uint32_t sbase = SMEM ? (uint32_t)__cvta_generic_to_shared(nodes) : 0u;
  asm volatile("" : "+r"(sbase));            // opaque: keep it in a register instead of re-deriving it per iteration
  // the next ray of this thread is fetched while the current one is traversed
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float4 o4n = make_float4(0.f, 0.f, 0.f, 0.f), d4n = o4n;
  if (i < count) { o4n = ray_o[i]; d4n = ray_d[i]; }
  for (; i < count; i += stride) {
    const float4 o4 = o4n, d4 = d4n;
    if (i + stride < count) { o4n = ray_o[i + stride]; d4n = ray_d[i + stride]; }
    trav_init(T, o4, d4, tmax, i);
    if (MASK & 0x40) {   // Philox address of this ray: (pixel, sample, bounce = depth + 1)
      const int sd = __float_as_int(d4.w);
      T.ra.seed = seed; T.ra.pixel = state ? (uint32_t)__float_as_int(state[i].w) : (uint32_t)i;
      T.ra.sample = (uint32_t)sd >> 12; T.ra.bounce = (uint32_t)(sd & 0xfff) + 1u;
    }
    for (int g = 0; g < sc.n_global; ++g) {    // primitives kept outside the tree first, all lanes together (uniform control flow)
      // FP64 sphere terms for the huge ones (ground spheres); the small outliers take the fp32 test like the tree's leaves
      if ((MASK & SRT_MASK_LEAF32) && ((sc.global_small >> g) & 1)) intersect_prim<MASK>(sc, ps, sc.global_prims[g], T.o, T.d, T.time, T.inv_a, tmin, T.ra, T.h);
      else intersect_prim<MASK & ~SRT_MASK_LEAF32>(sc, ps, sc.global_prims[g], T.o, T.d, T.time, T.inv_a, tmin, T.ra, T.h);
    }
    bool more = sc.n_items > 0;
    while (more) {
      int pend0 = 0, pend1 = 0;                  // raw leaf references (negative), 0 = none
      more = node_step<SMEM, TRAV>(T, nodes, sbase, tmin, pend0, pend1);
      while (pend0 < 0) {
#ifdef SRT_COUNT_STEPS
        T.ntests++;
#endif
        intersect_prim<MASK>(sc, ps, ~pend0, T.o, T.d, T.time, T.inv_a, tmin, T.ra, T.h); pend0 = pend1; pend1 = 0;
      }
    }
#ifdef SRT_COUNT_STEPS
    T.h.u = (float)T.nsteps; T.h.v = (float)(T.ntests + 1000 * T.maxsp);   // decoded by tools/step_stats.py
#endif
    hit[i] = make_float4(T.h.t, __int_as_float(T.h.prim), T.h.u, T.h.v);
  }
}

// Variant for scenes with EXPENSIVE primitives (Bezier curve / patch, constant medium: thousands of
// instructions per test).  Interleaving such a test with traversal serialises it lane by lane
// (each lane reaches its leaf in a different iteration).  Here a lane that reaches an expensive
// leaf PARKS; a warp vote runs the test block only when >= `vote` lanes are parked or no lane can
// make progress otherwise, so the block executes with many active lanes ("deferred candidate
// queue", one or two entries per lane, in registers).  Cheap primitives (spheres, rects) are still
// intersected immediately.
//
// Lanes are REFILLED: a lane whose ray is finished writes its hit record and becomes idle; as soon
// as >= `refill` lanes are idle they take the next rays of the warp's current 32-ray chunk (idle
// lanes in lane order take consecutive rays, so the fetch stays a contiguous segment).  Without the
// refill every 32-ray batch ended with a drain in which the few lanes that still had a parked test
// ran it almost alone (ncu source page, cfg5_teapot: 8 of 32 lanes at the entry of the patch test,
// 5.4 in its Newton loop, 9.3 in the node loop).  For the cheap sphere / rect variants the same
// refill was measured slower (bookkeeping > idle lanes recovered, profiles/README.md); here one test
// is worth hundreds of node steps.
// node steps per bookkeeping round of the deferred loop (measured, round 2, Grays/s on cfg5 / cfg5_teapot / cfg5_curves:
// 1: 2.73 / 3.21 / 3.39, 2: 2.83 / 3.36 / 3.49, 4: 2.97 / 3.55 / 3.53, 6: 3.00 / 3.59 / 3.51, 8: 3.01 / 3.58 / 3.52,
// 12: 2.99 / 3.54 / 3.52, unbounded: 2.93 / 3.34 / 3.52)
#ifndef SRT_DEFER_STEPS
#define SRT_DEFER_STEPS 6
#endif
constexpr int EXT_PARK_VOTE = 16, EXT_REFILL_MIN = 16;   // measured sweep: profiles/README.md (tools/sweep_deferred.sh)
template <bool SMEM, int MASK, int TRAV, class PrimSrc>
__device__ __forceinline__ void extend_loop_deferred(const DScene& sc, const float4* __restrict__ nodes, const PrimSrc& ps,
                                                     const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ state,
                                                     float4* __restrict__ hit, int count, float tmin, float tmax, uint32_t seed, uint32_t stack_base, int tune) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  if (sc.n_surf == 0) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) hit[i] = make_float4(tmax, __int_as_float(-1), 0.f, 0.f);
    return;
  }
  const int vote = tune & 0xff, refill = (tune >> 8) & 0xff;
  Trav T;
  T.n_nodes = sc.n_nodes;
#ifdef SRT_BOUNDS_CHECK
  T.stack_slots = trav_stack_stride(sc.bvh_depth) - 1;     // slots above the sentinel
#endif
  T.sp0 = stack_base + (TRAV == TRAV_STACK32 ? 4u : 2u) * (uint32_t)trav_stack_stride(sc.bvh_depth) * threadIdx.x;
  T.sp = T.sp0;
  if (TRAV == TRAV_STACK) T.sp = T.sp0 = trav_stack_init(T.sp0);
  if (TRAV == TRAV_STACK32) T.sp = T.sp0 = trav_stack32_init(T.sp0);
  uint32_t sbase = SMEM ? (uint32_t)__cvta_generic_to_shared(nodes) : 0u;
  asm volatile("" : "+r"(sbase));            // opaque: keep it in a register instead of re-deriving it per iteration
  const int stride = gridDim.x * blockDim.x;
  int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;   // warp-uniform: the warp's current 32-ray chunk
  int cur = 0;                                                // rays of that chunk already handed out (warp-uniform)
  int ray = -1;                                               // this lane's ray, -1 = idle
  bool more = false;
  int park0 = -1, park1 = -1;
  for (;;) {
    if (ray >= 0 && !more && park0 < 0) {                     // finished: retire
#ifdef SRT_COUNT_STEPS
      T.h.u = (float)T.nsteps; T.h.v = (float)(T.ntests + 1000 * T.maxsp);   // decoded by tools/step_stats.py
#endif
      hit[ray] = make_float4(T.h.t, __int_as_float(T.h.prim), T.h.u, T.h.v);
      ray = -1;
    }
    const unsigned idle = __ballot_sync(full, ray < 0);
    if (base < count && __popc(idle) >= refill) {             // hand the next rays of the chunk to the idle lanes
      const int take = min(__popc(idle), 32 - cur);
      const int i = base + cur + __popc(idle & ((1u << lane) - 1u));
      if (ray < 0 && i < base + cur + take && i < count) {
        const float4 d4 = ray_d[i];
        trav_init(T, ray_o[i], d4, tmax, i);
        const int sd = __float_as_int(d4.w);
        T.ra.seed = seed; T.ra.pixel = state ? (uint32_t)__float_as_int(state[i].w) : (uint32_t)i;
        T.ra.sample = (uint32_t)sd >> 12; T.ra.bounce = (uint32_t)(sd & 0xfff) + 1u;
        for (int g = 0; g < sc.n_global; ++g)
          intersect_prim<MASK>(sc, ps, sc.global_prims[g], T.o, T.d, T.time, T.inv_a, tmin, T.ra, T.h);
        ray = i; more = sc.n_items > 0;
      }
      cur += take;
      if (cur == 32) { cur = 0; base += stride; }
    }
    const bool parked = park0 >= 0;
    const unsigned pm = __ballot_sync(full, parked);
    const unsigned rm = __ballot_sync(full, ray >= 0 && more && !parked);
    if ((pm | rm) == 0u) {                                    // nothing to step or test: lanes are idle or about to retire
      if (base >= count && __ballot_sync(full, ray >= 0) == 0u) break;
      continue;
    }
    if (__popc(pm) >= vote || rm == 0u) {
      if (parked) {
#ifdef SRT_COUNT_STEPS
        T.ntests++;
#endif
        intersect_prim<MASK & 0x1e0>(sc, ps, park0, T.o, T.d, T.time, T.inv_a, tmin, T.ra, T.h);   // only the expensive kinds are ever parked
        park0 = park1; park1 = -1;
      }
      continue;
    }
    // up to SRT_DEFER_STEPS node steps per round of the bookkeeping above (three ballots, the refill and vote logic:
    // ~40 instructions, as much as a node step itself); a lane that parks or finishes early waits for the round to end
#pragma unroll 1
    for (int k = 0; k < SRT_DEFER_STEPS && ray >= 0 && more && park0 < 0; ++k) {
      int pend0 = 0, pend1 = 0;                  // raw leaf references (negative), 0 = none
      float4 ivl = make_float4(0.f, 0.f, 0.f, 0.f);
      more = node_step<SMEM, TRAV, (MASK & 0x80) != 0>(T, nodes, sbase, tmin, pend0, pend1, &ivl);
      while (pend0 < 0) {
        const int leaf = ~pend0;
        const int type = ps.hdr(leaf).x & 0xff;
        if (type >= SRT_PRIM_BEZIER) {
          // a patch leaf is parked only if the ray meets its bounding slab inside the leaf's box interval
          // (15 instructions here against a ~300-instruction projection + hull cull inside the parked test)
          const bool culled = (MASK & 0x80) && type == SRT_PRIM_PATCH && patch_slab_culled(ps.a(leaf), T.o, T.d, ivl.x, ivl.y);
          if (!culled) { if (park0 < 0) park0 = leaf; else park1 = leaf; }
        } else {
#ifdef SRT_COUNT_STEPS
          T.ntests++;
#endif
          intersect_prim<MASK & 0x1f>(sc, ps, leaf, T.o, T.d, T.time, T.inv_a, tmin, T.ra, T.h);
        }
        pend0 = pend1; pend1 = 0; ivl.x = ivl.z; ivl.y = ivl.w;
      }
    }
  }
}

// CTA size: 256 threads, 4 CTAs per SM for the sphere and the rect / instance variants.  The variant with the
// expensive primitives needs 128 registers, i.e. at most 512 threads per SM; it runs as ONE 512-thread
// CTA per SM so that those 16 warps share one staged copy of the scene (as 2 x 256 the teapot scene's
// 147 KB tree allowed a single 256-thread CTA: 8 warps/SM, issue slots 39 % busy).
// (round 2: the rect / instance variant also runs 4 CTAs per SM - 62 registers with 56 B of spills beat 80 registers
// at 3 CTAs: cfg4 8.96 -> 9.36, cfg3 6.79 -> 6.96 Grays/s)
#ifndef SRT_RECT_CTAS
#define SRT_RECT_CTAS 4
#endif
template <bool SMEM, int MASK, int TRAV>
__global__ void __launch_bounds__((MASK & 0x1e0) ? EXT_THREADS_HEAVY : EXT_THREADS, (MASK & 0x1e0) ? 1 : ((MASK & 0x1c) ? SRT_RECT_CTAS : 4))
k_extend(DScene sc, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ state,
         float4* __restrict__ hit, const int* __restrict__ count_ptr, int count_fixed, float tmin, float tmax, uint32_t seed, int tune) {
  extern __shared__ float4 smem[];
  const int count = count_ptr ? *count_ptr : count_fixed;
  if ((int)(blockIdx.x * blockDim.x) >= count) return;      // no ray for this CTA (drain tail): skip the staging too
  if (SMEM) {
    // stage the whole LBVH + primitive headers ("shared-memory staging of BVH top levels":
    // for the <= few-thousand-primitive scenes of the reference the top levels are all levels)
    int nn = 4 * sc.n_nodes, np = sc.n_prims;
    stage_nodes(smem, sc.nodes, sc.n_nodes);
    int4* sh = (int4*)(smem + nn);
    float4* sa = smem + nn + np;
    for (int i = threadIdx.x; i < np; i += blockDim.x) { sh[i] = sc.prim_hdr[i]; sa[i] = sc.prim_a[i]; }
    __syncthreads();
    PrimShared ps(sh, sa);
    const uint32_t stack_base = (uint32_t)__cvta_generic_to_shared(smem + nn + 2 * np);      // TRAV_STACK: after the staged scene
    if (MASK & 0x1e0) extend_loop_deferred<true, MASK, TRAV>(sc, smem, ps, ray_o, ray_d, state, hit, count, tmin, tmax, seed, stack_base, tune);
    else extend_loop<true, MASK, TRAV>(sc, smem, ps, ray_o, ray_d, state, hit, count, tmin, tmax, seed, stack_base);
  } else {
    PrimGlobal ps{sc.prim_hdr, sc.prim_a};
    const uint32_t stack_base = (uint32_t)__cvta_generic_to_shared(smem);
    if (MASK & 0x1e0) extend_loop_deferred<false, MASK, TRAV>(sc, sc.nodes, ps, ray_o, ray_d, state, hit, count, tmin, tmax, seed, stack_base, tune);
    else extend_loop<false, MASK, TRAV>(sc, sc.nodes, ps, ray_o, ray_d, state, hit, count, tmin, tmax, seed, stack_base);
  }
}

// ------------------------------------------------------------------------------------------------
// One bounce of `color` (main.scm:100-121) for one path: hit-record completion, emitted / sky into
// the accumulator, scatter.  Returns true when the path continues; the new ray / state are written
// to (no4, nd4, ns4) in the queue layout.  Shared by k_shade (wavefront) and k_tail (drain).
template <int EST>
__device__ __forceinline__ bool shade_path(const DScene& sc, const SrtRenderParams& p, float4 h4, float4 o4, float4 d4, float4 s4,
                                           unsigned long long* __restrict__ accum, WaveCtrl* __restrict__ ctrl,
                                           float4& no4, float4& nd4, float4& ns4) {
  const int prim = __float_as_int(h4.y);
  float3 thr = xyz(s4); const int pixel = __float_as_int(s4.w); const int sd = __float_as_int(d4.w);
  const int depth = sd & 0xfff; const unsigned int sample = (unsigned int)sd >> 12;
  const float3 o = xyz(o4), d = xyz(d4);
  if (!SRT_BOUNDS_OK(pixel >= 0 && pixel < p.width * p.height && depth <= p.max_depth, 201)) return false;
  if (prim < 0) {                                              // main.scm:120 sky
    accumulate_fixed(accum, pixel, thr * sky_value(p.sky, d), &ctrl->nonfinite);
    return false;
  }
  float3 pt, n; int material;
  complete_hit(sc, prim, h4.x, h4.z, h4.w, o, d, o4.w, pt, n, material);
  RngAddr addr{p.seed, (uint32_t)pixel, sample, (uint32_t)(depth + 1)};
  Scatter s = scatter<EST>(sc, prim, d, pt, n, h4.z, h4.w, addr, p.quirks);
  if (s.emitted.x != 0.f || s.emitted.y != 0.f || s.emitted.z != 0.f)    // main.scm:113/119 emitted
    accumulate_fixed(accum, pixel, thr * s.emitted, &ctrl->nonfinite);
  if (!(s.valid && depth < p.max_depth)) return false;       // main.scm:112
  thr = thr * s.weight;
  const float ntime = (p.quirks & SRT_Q6_SCATTER_TIME0) ? 0.0f : o4.w;   // Q6: make-ray forces time 0
  no4 = make_float4(pt.x, pt.y, pt.z, ntime);
  nd4 = make_float4(s.dir.x, s.dir.y, s.dir.z, __int_as_float(sd + 1));
  ns4 = make_float4(thr.x, thr.y, thr.z, __int_as_float(pixel));
  return true;
}

// shade: one bounce for every live path + compaction of survivors into the other queue generation
// (warp ballot -> per-warp count -> one atomic per CTA).  Measured and NOT adopted (profiles/README.md,
// round 2): the per-primitive tables staged in shared memory (+1 % cfg2, -2 % cfg3) and a software
// prefetch of the next tile's queue entries (-9 % at 4 CTAs/SM with spills, -24 % at 3 CTAs/SM); one atomic per WARP
// instead of the two barriers + one atomic per CTA (shade 2x slower: 16 M single-address atomics per launch serialise in L2).
struct ShadeShared { int warp[2][SHD_THREADS / 32]; int base[2]; unsigned hist[8]; };    // [2]: tiles alternate buffers, see shade_tiles
// One pass over queue generation g: shade every path, compact the survivors behind *next_count (shared by the
// wavefront's k_shade and the persistent drain kernel).  blockDim.x == SHD_THREADS.
template <int EST>
__device__ __forceinline__ void shade_tiles(const DScene& sc, const SrtRenderParams& p, int count,
                                            const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ state, const float4* __restrict__ hit,
                                            float4* __restrict__ ray_o_next, float4* __restrict__ ray_d_next, float4* __restrict__ state_next,
                                            unsigned long long* __restrict__ accum, WaveCtrl* __restrict__ ctrl, int* next_count, ShadeShared& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (p.reserved[0] == 1) { if (threadIdx.x < 8) S.hist[threadIdx.x] = 0u; __syncthreads(); }
  // Compaction: per-warp survivor counts -> barrier -> thread 0 turns them into offsets and reserves the tile's slots with one
  // global atomic -> barrier -> scatter.  Consecutive tiles alternate between two sets of counters (a warp can only be two tiles
  // ahead of another after both passed the two barriers in between), which saves round 1's third barrier: shade +4.8 % on cfg2.
  // Measured and rejected: one barrier, the LAST warp to arrive reserving the slots (offsets from a shared-memory atomic, i.e.
  // in arrival order): shade -7 % and extend -4 % - the survivors of neighbouring warps are no longer neighbours in the queue.
  int par = 0;
  for (int base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x, par ^= 1) {   // block-uniform trip count
    const int i = base + threadIdx.x;
    bool alive = false;
    float4 no4 = make_float4(0.f, 0.f, 0.f, 0.f), nd4 = no4, ns4 = no4, d4 = no4;
    if (i < count) { d4 = ray_d[i]; alive = shade_path<EST>(sc, p, hit[i], ray_o[i], d4, state[i], accum, ctrl, no4, nd4, ns4); }
    if (p.reserved[0] == 1) {                     // profile mode: closest-hit queries per bounce (per-CTA histogram in shared memory)
      const int bucket = i < count ? min(__float_as_int(d4.w) & 0xfff, 7) : 8;
      const unsigned peers = __match_any_sync(0xffffffffu, bucket);
      if (bucket < 8 && lane == __ffs(peers) - 1) atomicAdd(&S.hist[bucket], (unsigned)__popc(peers));
    }
    unsigned ballot = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) S.warp[par][warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < SHD_THREADS / 32; ++w) { int c = S.warp[par][w]; S.warp[par][w] = tot; tot += c; }
      S.base[par] = tot ? atomicAdd(next_count, tot) : 0;
    }
    __syncthreads();
    if (alive) {
      int pos = S.base[par] + S.warp[par][warp] + __popc(ballot & ((1u << lane) - 1u));
      if (SRT_BOUNDS_OK(pos >= 0 && pos < count, 221)) {       // survivors of a generation never outnumber it
        ray_o_next[pos] = no4; ray_d_next[pos] = nd4; state_next[pos] = ns4;
      }
    }
  }
  if (p.reserved[0] == 1 && threadIdx.x < 8 && S.hist[threadIdx.x]) atomicAdd(&ctrl->bounce_hist[threadIdx.x], (unsigned long long)S.hist[threadIdx.x]);
}
#ifndef SRT_SHADE_CTAS
#define SRT_SHADE_CTAS 4
#endif
template <int EST>
__global__ void __launch_bounds__(SHD_THREADS, SRT_SHADE_CTAS * (256 / SHD_THREADS))
k_shade(DScene sc, SrtRenderParams p, int g,
        const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ state, const float4* __restrict__ hit,
        float4* __restrict__ ray_o_next, float4* __restrict__ ray_d_next, float4* __restrict__ state_next,
        unsigned long long* __restrict__ accum, WaveCtrl* __restrict__ ctrl) {
  __shared__ ShadeShared S;
  shade_tiles<EST>(sc, p, ctrl->qcount[g], ray_o, ray_d, state, hit, ray_o_next, ray_d_next, state_next, accum, ctrl, &ctrl->survivors[g ^ 1], S);
}

// ------------------------------------------------------------------------------------------------
// tail: the drain of the streaming wavefront in ONE launch.  Once every camera path has been
// generated and the queue holds at most `tail_max` paths (about one wave of resident threads), the
// remaining iterations of the wavefront are ~max_depth launch-bound, nearly empty extend / shade /
// regen triples (main.scm:26 +max-depth+ is what makes the tail: a path may live 50 bounces).  Here
// every thread takes one queued path and runs extend -> shade -> extend ... until the path ends
// (the same device functions, the same Philox addresses (pixel, sample, bounce) and the same
// integer accumulator, so the image is bit-identical to the wavefront's).  The kernel is part of
// every iteration batch and returns at once while its condition does not hold; k_tail_done (one
// thread) then empties the queue, so the remaining launches of the batch find nothing to do.
// (Also measured, round 2, and rejected: a PERSISTENT drain for larger queues - one launch that loops
// extend | grid barrier | shade | grid barrier per bounce with the tree staged once.  Bit-identical, fewer
// launches (63 instead of 227 for one GPU's share of cfg2 on an 8-GPU box), but slower: cfg2 at 63 spp 24.0 vs
// 22.7 ms, full frame 176.5 vs 173.5 ms, cfg3 29.6 vs 30.0, cfg4 equal - the fused kernel's shade phase runs
// with the extend phase's shared-memory carve-out (211 KB of tree copies per SM, ~17 KB of L1 left for the
// material tables) and without the oversubscribed grid that balances k_shade's tiles.)
// Two rules: a queue of <= tail_max paths always goes to the tail kernel; one of <= tail_slow (16 Mi) paths when it has
// stopped shrinking (>= SRT_TAIL_RATIO % of the previous iteration's length): those are the long-lived paths - in
// random-scene 3.5 % of the paths bounce inside glass until the depth limit, ~2 M paths per 62.5 spp for 35-40 more
// iterations - which the kernel carries in registers, while a queue that still decays fast (cfg3: -30 % per iteration) is
// cheaper to finish through the wavefront.  The RATIO, not the size, finds the right iteration for every share of a frame:
// a fixed 2 Mi threshold only fitted one rank's share of an 8-GPU cfg2 frame (profiles/r2_sweep22_tail_rule.txt).
// 4 CTAs per SM (64 registers, 88 B of spills): the kernel is latency-bound, 2 CTAs at 108 registers were 17 % slower.
#ifndef SRT_TAIL_RATIO
#define SRT_TAIL_RATIO 96      // percent
#endif
__device__ __forceinline__ bool tail_condition(const WaveCtrl* ctrl, int g, int parity, int tail_max, int tail_slow) {
  const int c = ctrl->qcount[g], prev = ctrl->qcount[g ^ 1];      // prev: the queue one iteration ago
  if (c <= 0 || ctrl->next_path[parity] < ctrl->total_paths) return false;
  return c <= tail_max || (c <= tail_slow && 100ll * c >= (long long)SRT_TAIL_RATIO * prev);
}
template <int MASK, int EST>
#ifndef SRT_TAIL_CTAS
#define SRT_TAIL_CTAS 4
#endif
__global__ void __launch_bounds__(EXT_THREADS, SRT_TAIL_CTAS)
k_tail(DScene sc, SrtRenderParams p, int g, int parity, int tail_max, int tail_slow,
       const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ state,
       unsigned long long* __restrict__ accum, WaveCtrl* __restrict__ ctrl) {
  extern __shared__ float4 smem[];
  if (!tail_condition(ctrl, g, parity, tail_max, tail_slow)) return;
  const int count = ctrl->qcount[g];
  if ((int)(blockIdx.x * blockDim.x) >= count) return;
  const int nn = 4 * sc.n_nodes, np = sc.n_prims;
  stage_nodes(smem, sc.nodes, sc.n_nodes);
  int4* sh = (int4*)(smem + nn);
  float4* sa = smem + nn + np;
  for (int i = threadIdx.x; i < np; i += blockDim.x) { sh[i] = sc.prim_hdr[i]; sa[i] = sc.prim_a[i]; }
  __syncthreads();
  const PrimShared ps(sh, sa);
  Trav T;
  T.n_nodes = sc.n_nodes;
#ifdef SRT_BOUNDS_CHECK
  T.stack_slots = trav_stack_stride(sc.bvh_depth) - 1;     // slots above the sentinel
#endif
  T.sp0 = (uint32_t)__cvta_generic_to_shared(smem + nn + 2 * np) + 2u * (uint32_t)trav_stack_stride(sc.bvh_depth) * threadIdx.x;
  T.sp = T.sp0 = trav_stack_init(T.sp0);
  uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("" : "+r"(sbase));
  T.ra.seed = p.seed; T.ra.pixel = 0u; T.ra.sample = 0u; T.ra.bounce = 0u;
  unsigned int nrays = 0;
  // Every LANE runs its own sequence of paths (grid-stride) and takes the next one the moment its
  // current path ends: one trip of the loop is one bounce for every lane that still has work, so the
  // lanes of a warp stay busy although their paths have different lengths (no queue, no compaction,
  // no HBM traffic between bounces - the path state lives in registers).
  int next = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  bool alive = false;
  float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = o4, s4 = o4;
  for (;;) {
    if (!alive) {
      if (next >= count) break;
      o4 = ray_o[next]; d4 = ray_d[next]; s4 = state[next];
      next += stride; alive = true;
    }
    trav_init(T, o4, d4, SRT_MAX_FLOAT, 0);
    ++nrays;
    if (sc.n_surf > 0) {
      for (int k = 0; k < sc.n_global; ++k) {
        if ((MASK & SRT_MASK_LEAF32) && ((sc.global_small >> k) & 1)) intersect_prim<MASK>(sc, ps, sc.global_prims[k], T.o, T.d, T.time, T.inv_a, p.t_min, T.ra, T.h);
        else intersect_prim<MASK & ~SRT_MASK_LEAF32>(sc, ps, sc.global_prims[k], T.o, T.d, T.time, T.inv_a, p.t_min, T.ra, T.h);
      }
      bool more = sc.n_items > 0;
      while (more) {
        int pend0 = 0, pend1 = 0;
        more = node_step<true, TRAV_STACK>(T, smem, sbase, p.t_min, pend0, pend1);
        while (pend0 < 0) { intersect_prim<MASK>(sc, ps, ~pend0, T.o, T.d, T.time, T.inv_a, p.t_min, T.ra, T.h); pend0 = pend1; pend1 = 0; }
      }
    }
    float4 no4, nd4, ns4;
    alive = shade_path<EST>(sc, p, make_float4(T.h.t, __int_as_float(T.h.prim), T.h.u, T.h.v), o4, d4, s4, accum, ctrl, no4, nd4, ns4);
    if (alive) { o4 = no4; d4 = nd4; s4 = ns4; }
  }
  for (int o = 16; o; o >>= 1) nrays += __shfl_xor_sync(0xffffffffu, nrays, o);
  if ((threadIdx.x & 31) == 0 && nrays) atomicAdd(&ctrl->rays, (unsigned long long)nrays);
}
__global__ void k_tail_done(int g, int parity, int tail_max, int tail_slow, WaveCtrl* ctrl) {
  if (!tail_condition(ctrl, g, parity, tail_max, tail_slow)) return;
  ctrl->rays -= (unsigned long long)ctrl->qcount[g];    // the regen that filled this generation already counted its first query
  ctrl->qcount[g] = 0; ctrl->iterations += 1ull; ctrl->tail_runs += 1;
}

// end of render: rgb_sum += fixed-point accumulator (main.scm:480 running sum, *raw-data*)
__global__ void k_accum_to_float(int n3, const unsigned long long* __restrict__ accum, float* __restrict__ rgb_sum) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x)
    rgb_sum[i] += (float)((double)(long long)accum[i] * (1.0 / 68719476736.0));
}
__global__ void k_ctrl_init(WaveCtrl* ctrl, unsigned long long total_paths, int spp_begin) {
  ctrl->qcount[0] = ctrl->qcount[1] = 0; ctrl->survivors[0] = ctrl->survivors[1] = 0;
  ctrl->next_path[0] = ctrl->next_path[1] = 0ull; ctrl->total_paths = total_paths; ctrl->rays = 0ull; ctrl->iterations = 0ull;
  ctrl->nonfinite = 0ull; ctrl->spp_begin = spp_begin; ctrl->tail_runs = 0;
  for (int k = 0; k < 8; ++k) ctrl->bounce_hist[k] = 0ull;
}

// main.scm:123-124, 481-487: correct-gamma (sqrt) + floor(255.99 * min(1, c)); negative sums
// (outside the reference's domain) clamp to 0.
__global__ void k_resolve(const float* __restrict__ rgb_sum, int n3, float spp, uint8_t* __restrict__ image) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x) {
    float c = __fdiv_rn(rgb_sum[i], spp);
    c = sqrtf(fmaxf(c, 0.0f));
    image[i] = (uint8_t)floorf(255.99f * fminf(1.0f, c));
  }
}

// Multi-GPU combine on the root GPU, one kernel: frame = own accumulator + the peers' accumulators
// read directly over NVLink peer access (integer adds: exact, order-independent), then
// rgb_sum += frame * 2^-36 (main.scm:480) and, optionally, gamma + 8-bit (main.scm:481-487) for the
// `spp` samples the running sum now holds.  With no peers it is the single-GPU epilogue.
struct PeerPtrs { const unsigned long long* p[SRT_MAX_DEVICES]; };
__global__ void __launch_bounds__(256) k_reduce_peers(int n3, unsigned long long* __restrict__ accum, PeerPtrs peers, int n_peers,
                                                      float* __restrict__ rgb_sum, float spp, uint8_t* __restrict__ image) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x) {
    unsigned long long a = accum[i];
    for (int k = 0; k < n_peers; ++k) a += peers.p[k][i];
    if (n_peers) accum[i] = a;
    const float v = rgb_sum[i] + (float)((double)(long long)a * (1.0 / 68719476736.0));
    rgb_sum[i] = v;
    if (image) { float c = sqrtf(fmaxf(__fdiv_rn(v, spp), 0.0f)); image[i] = (uint8_t)floorf(255.99f * fminf(1.0f, c)); }
  }
}

// ---- parity hooks --------------------------------------------------------------------------------
__global__ void k_upload_rays(const SrtRay* __restrict__ rays, int n, float4* __restrict__ ray_o, float4* __restrict__ ray_d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SrtRay r = rays[i];
  ray_o[i] = make_float4(r.o[0], r.o[1], r.o[2], r.time);
  ray_d[i] = make_float4(r.d[0], r.d[1], r.d[2], 0.f);
}
__global__ void k_complete_hits(DScene sc, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ hit, int n, SrtHit* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 h4 = hit[i], o4 = ray_o[i], d4 = ray_d[i];
  SrtHit r; r.prim = __float_as_int(h4.y); r.material = -1; r.t = 0.f; r.u = r.v = 0.f;
  r.p[0] = r.p[1] = r.p[2] = 0.f; r.n[0] = r.n[1] = r.n[2] = 0.f;
#ifdef SRT_COUNT_STEPS
  r.u = h4.z; r.v = h4.w;
#endif
  if (r.prim >= 0) {
    float3 p, nn; int m;
    complete_hit(sc, r.prim, h4.x, h4.z, h4.w, xyz(o4), xyz(d4), o4.w, p, nn, m);
    r.t = h4.x; r.u = h4.z; r.v = h4.w; r.material = m;
    int type = sc.prim_hdr[r.prim].x & 0xff;
#ifndef SRT_COUNT_STEPS
    if (type <= SRT_PRIM_MOVING_SPHERE) sphere_uv(p, r.u, r.v);     // Q5 (dead in shading)
#else
    (void)type;
#endif
    r.p[0] = p.x; r.p[1] = p.y; r.p[2] = p.z; r.n[0] = nn.x; r.n[1] = nn.y; r.n[2] = nn.z;
    r.prim = sc.prim_logical[r.prim];
  }
  out[i] = r;
}
__global__ void k_eval_texture(DScene sc, int tex, const float* __restrict__ uvp5, int n, int quirks, float* __restrict__ rgb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = uvp5 + 5 * (size_t)i;
  float3 c = tex_value(sc, tex, q[0], q[1], v3(q[2], q[3], q[4]), quirks);
  rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
}
__global__ void k_eval_raygen(DCamera cam, SrtRenderParams p, int n, const int* __restrict__ pixel, const int* __restrict__ sample, SrtRay* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int px = pixel[i], y = px / p.width, x = px - y * p.width;
  RngAddr addr{p.seed, (uint32_t)px, (uint32_t)sample[i], 0u};
  float4 xi = rng_block(addr, 0);
  float3 o, d; float time;
  get_ray(cam, ((float)x + xi.x) / (float)p.width, ((float)y + xi.y) / (float)p.height, xi.z, addr, o, d, time);
  SrtRay r; r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z; r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z; r.time = time;
  out[i] = r;
}

// FP32 roofline denominator: dependent-chain-free FFMA loop, 8 accumulators per thread.
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace

// Measures sustained FFMA throughput (TFLOP/s, 2 flops per FFMA) on the current device.
float srt_measure_fma_tflops(int sm_count, cudaStream_t stream) {
  const int grid = sm_count * 8, iters = 4096;
  float* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(float) * (size_t)grid * 256) != cudaSuccess) return 0.f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_fma_peak<<<grid, 256, 0, stream>>>(d, 64, 0.999f, 0.001f);            // warm-up
  float best = 0.f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, stream);
    k_fma_peak<<<grid, 256, 0, stream>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1, stream); cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * 256.0;
    float tf = (float)(flops / (ms * 1e-3) / 1e12);
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  return best;
}

// bounds-checked build: violations counted by the kernels of this translation unit (0 in the release build)
unsigned long long srt_bounds_violations_wavefront(int* first) {
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

// =================================================================================================
size_t srt_extend_smem_bytes(const DScene& sc) { return (size_t)64 * sc.n_nodes + (size_t)32 * sc.n_prims; }

// Kernel variants by primitive mix.  The cache (function attribute + occupancy query) is kept per
// DEVICE: scenes on different GPUs of one process (srt_render_multi) each see their own entry.
typedef void (*ExtendFn)(DScene, const float4*, const float4*, const float4*, float4*, const int*, int, float, float, uint32_t, int);
typedef void (*TailFn)(DScene, SrtRenderParams, int, int, int, int, const float4*, const float4*, const float4*, unsigned long long*, WaveCtrl*);

struct ExtendVariant { ExtendFn fn; int bps; size_t smem; int threads; };
struct TailVariant { TailFn fn; size_t smem; int bps; };
static ExtendVariant g_variants[SRT_MAX_DEVICES][2][6][3][2];
static TailVariant g_tail_variants[SRT_MAX_DEVICES][3][2][2];
static std::mutex g_variant_mu;

// spheres | + moving spheres | + rects / instances | + bicubic patches | everything (curves, media, Klein) |
// patches + curves.  The patch variant exists because the curve's subdivision stack (1.3 KB of local memory per
// thread) and the Klein / medium code cost the patch scenes 14 % when merely compiled in (cfg5_teapot 2.61 -> 2.96
// Grays/s); variant 5 is the same argument for cfg5 (patches AND the reference's curves, no medium / Klein).
// SRT_MASK_XF_SPHERE (an instanced sphere: translate / rotate-y above a sphere leaf, geometry.scm:465-543)
// needs the primitive header, which the sphere-only variant 0 does not read: such scenes use variant 1.
static int variant_of(int mask) {
  const bool xf_sphere = mask & SRT_MASK_XF_SPHERE;
  mask &= SRT_MASK_ALL;
  if ((mask & ~0x01) == 0 && !xf_sphere) return 0;
  if ((mask & ~0x03) == 0) return 1;                 // two kinds: the header (and its xform) is read
  if ((mask & 0x1e0) == 0) return 2;
  if ((mask & 0x160) == 0) return 3;
  if ((mask & 0x140) == 0) return 5;
  return 4;
}
template <bool SMEM, int TRAV> static ExtendFn variant_fn_m(int v, bool leaf32) {
  switch (v) {
    case 0: return leaf32 ? k_extend<SMEM, 0x01 | SRT_MASK_LEAF32, TRAV> : k_extend<SMEM, 0x01, TRAV>;
    case 1: return leaf32 ? k_extend<SMEM, 0x03 | SRT_MASK_LEAF32, TRAV> : k_extend<SMEM, 0x03, TRAV>;
    case 2: return leaf32 ? k_extend<SMEM, 0x1f | SRT_MASK_LEAF32, TRAV> : k_extend<SMEM, 0x1f, TRAV>;
    case 3: return k_extend<SMEM, 0x9f, TRAV>;
    case 5: return k_extend<SMEM, 0xbf, TRAV>;
    default: return k_extend<SMEM, SRT_MASK_ALL, TRAV>;
  }
}
static ExtendFn variant_fn(bool smem, int v, int trav, bool leaf32) {
  if (trav == TRAV_STACK) return smem ? variant_fn_m<true, TRAV_STACK>(v, leaf32) : variant_fn_m<false, TRAV_STACK>(v, leaf32);
  if (trav == TRAV_CACHE) return smem ? variant_fn_m<true, TRAV_CACHE>(v, leaf32) : variant_fn_m<false, TRAV_CACHE>(v, leaf32);
  return smem ? variant_fn_m<true, TRAV_STACK32>(v, leaf32) : variant_fn_m<false, TRAV_STACK32>(v, leaf32);
}
static bool leaf32_of(const RenderLaunch& L, int v) {
  static const bool off = getenv("SRT_NO_LEAF32") != nullptr;                        // A/B switch
  return !off && v <= 2 && (L.prim_mask & SRT_MASK_LEAF32);
}
static int trav_mode(const RenderLaunch& L, int v, size_t* stack_bytes) {
  const int which = L.bvh_in_smem ? 1 : 0;
  // 16-bit node ids: shared-memory stack when it fits beside the staged scene, else the register cache
  const int threads = v >= 3 ? EXT_THREADS_HEAVY : EXT_THREADS;
  const size_t stack = (size_t)2 * trav_stack_stride(L.sc.bvh_depth) * threads;
  const bool small_ids = L.sc.n_nodes < 65536;
  static const bool force_cache = getenv("SRT_TRAV_CACHE") != nullptr;             // A/B switch for profiling
  *stack_bytes = stack;
  if (!small_ids) { *stack_bytes = 2 * stack; return TRAV_STACK32; }      // never staged (>= 4 MB of nodes): the 32-bit stack always fits
  return (!force_cache && (which ? L.extend_smem : 0) + stack <= (size_t)200 * 1024) ? TRAV_STACK : TRAV_CACHE;
}
// persistent grid: SM count x resident CTAs per SM (queried; depends on the staged-BVH size)
static ExtendVariant extend_variant(const RenderLaunch& L) {
  const int which = L.bvh_in_smem ? 1 : 0, v = variant_of(L.prim_mask);
  size_t stack = 0;
  const int trav = trav_mode(L, v, &stack);
  const int threads = v >= 3 ? EXT_THREADS_HEAVY : EXT_THREADS;
  const bool leaf32 = leaf32_of(L, v);
  std::lock_guard<std::mutex> lock(g_variant_mu);
  ExtendVariant& e = g_variants[L.device][which][v][trav][leaf32 ? 1 : 0];
  size_t smem = (which ? L.extend_smem : 0) + (trav != TRAV_CACHE ? stack : 0);
  if (!e.fn || e.smem != smem) {
    e.fn = variant_fn(which, v, trav, leaf32); e.smem = smem; e.threads = threads;
    if (smem) cudaFuncSetAttribute(e.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, e.fn, threads, smem);
    e.bps = bps < 1 ? 1 : bps;
  }
  return e;
}
// The drain kernel exists for the cheap variants with the tree staged in shared memory and the
// shared-memory traversal stack (every reference scene); other scenes drain through the wavefront.
static bool tail_variant(const RenderLaunch& L, TailVariant* out) {
  const int v = variant_of(L.prim_mask);
  size_t stack = 0;
  if (v > 2 || !L.bvh_in_smem || trav_mode(L, v, &stack) != TRAV_STACK) return false;
  static const bool off = getenv("SRT_NO_TAIL") != nullptr;                          // A/B switch
  if (off) return false;
  const int est = L.p.estimator == SRT_EST_MIXTURE ? 1 : 0;
  const bool leaf32 = leaf32_of(L, v);
  std::lock_guard<std::mutex> lock(g_variant_mu);
  TailVariant& t = g_tail_variants[L.device][v][est][leaf32 ? 1 : 0];
  const size_t smem = L.extend_smem + stack;
  if (!t.fn || t.smem != smem) {
    constexpr int L32 = SRT_MASK_LEAF32;
    if (est && leaf32) t.fn = v == 0 ? k_tail<0x01 | L32, SRT_EST_MIXTURE> : (v == 1 ? k_tail<0x03 | L32, SRT_EST_MIXTURE> : k_tail<0x1f | L32, SRT_EST_MIXTURE>);
    else if (est) t.fn = v == 0 ? k_tail<0x01, SRT_EST_MIXTURE> : (v == 1 ? k_tail<0x03, SRT_EST_MIXTURE> : k_tail<0x1f, SRT_EST_MIXTURE>);
    else if (leaf32) t.fn = v == 0 ? k_tail<0x01 | L32, SRT_EST_REFERENCE> : (v == 1 ? k_tail<0x03 | L32, SRT_EST_REFERENCE> : k_tail<0x1f | L32, SRT_EST_REFERENCE>);
    else t.fn = v == 0 ? k_tail<0x01, SRT_EST_REFERENCE> : (v == 1 ? k_tail<0x03, SRT_EST_REFERENCE> : k_tail<0x1f, SRT_EST_REFERENCE>);
    t.smem = smem;
    if (smem) cudaFuncSetAttribute(t.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, t.fn, EXT_THREADS, smem);
    t.bps = bps < 1 ? 1 : bps;
  }
  *out = t;
  return true;
}

void srt_extend_prepare(const RenderLaunch& L) { (void)extend_variant(L); }

int srt_launch_extend(const RenderLaunch& L, const float4* ray_o, const float4* ray_d, const float4* state, float4* hit, const int* d_count, int count,
                      float tmin, float tmax, uint32_t seed, cudaStream_t stream) {
  const ExtendVariant e = extend_variant(L);
  // deferred-test tuning (heavy variant only): park vote | refill threshold << 8; env overrides are for A/B runs
  static const int tune = [] {
    const char* a = getenv("SRT_PARK_VOTE"); const char* b = getenv("SRT_REFILL_MIN");
    int vote = a ? atoi(a) : EXT_PARK_VOTE, refill = b ? atoi(b) : EXT_REFILL_MIN;
    vote = vote < 1 ? 1 : (vote > 32 ? 32 : vote); refill = refill < 1 ? 1 : (refill > 32 ? 32 : refill);
    return vote | (refill << 8);
  }();
  const int div = L.grid_div > 1 ? L.grid_div : 1;
  e.fn<<<L.sm_count * std::max(1, e.bps / div), e.threads, e.smem, stream>>>(L.sc, ray_o, ray_d, state, hit, d_count, count, tmin, tmax, seed, tune);
  return 1;
}

size_t srt_wave_ctrl_bytes() { return sizeof(WaveCtrl); }

void srt_graph_cache_release(GraphCache& c) {
  if (c.exec) cudaGraphExecDestroy((cudaGraphExec_t)c.exec);
  if (c.graph) cudaGraphDestroy((cudaGraph_t)c.graph);
  c.exec = nullptr; c.graph = nullptr; c.key.clear();
}

// The streaming wavefront: iterate extend -> shade -> regen until every camera path of the sample
// range has been generated and the queue has drained.  The host enqueues iterations in batches
// and polls the control block (pinned copy) one batch behind, so the GPU never waits on the host.
// Returns the number of kernels launched, or a negative SrtError: every CUDA call of the loop is
// checked, and the loop is bounded, so a sticky device error cannot leave the host spinning.
// d_rgb_sum == nullptr leaves the frame in the 64-bit accumulator (multi-GPU: reduced first).
#define WCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = e_; goto fail; } } while (0)
int srt_wavefront_render(const RenderLaunch& L, WaveBuffers& W, float* d_rgb_sum, cudaStream_t stream, SrtStats* stats, bool profile, cudaError_t* cuda_err) {
  const SrtRenderParams& p = L.p;
  const int npix = p.width * p.height;
  const unsigned long long total = (unsigned long long)npix * (unsigned long long)(p.spp_end - p.spp_begin);
  const int cap = (int)W.capacity;
  WaveCtrl* ctrl = (WaveCtrl*)W.ctrl;
  int launches = 0;
  const int div = L.grid_div > 1 ? L.grid_div : 1;
  const int shade_grid = L.sm_count * std::max(1, 8 * (256 / SHD_THREADS) / div), regen_grid = L.sm_count * std::max(1, 8 / div);
  typedef void (*ShadeFn)(DScene, SrtRenderParams, int, const float4*, const float4*, const float4*, const float4*, float4*, float4*, float4*, unsigned long long*, WaveCtrl*);
  const ShadeFn shade_fn = p.estimator == SRT_EST_MIXTURE ? (ShadeFn)k_shade<SRT_EST_MIXTURE> : (ShadeFn)k_shade<SRT_EST_REFERENCE>;
  cudaError_t err = cudaSuccess;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  float acc_ext = 0.f, acc_shd = 0.f; int n_ext = 0;
  WaveCtrl* h = (WaveCtrl*)W.h_ctrl;            // pinned, 2 slots
  cudaEvent_t* ev = (cudaEvent_t*)W.poll_events;
  // kernels read the sample range from the control block, so the captured graph (and its key) do
  // not depend on it: progressive passes replay the same executable graph
  SrtRenderParams pk = p; pk.spp_begin = 0; pk.spp_end = 0; pk.wave_spp = 0; pk.reserved[1] = pk.reserved[2] = pk.reserved[3] = 0;
  TailVariant tv; const bool have_tail = p.reserved[3] != 1 && tail_variant(L, &tv);
  const int tail_grid = have_tail ? L.sm_count * std::max(1, tv.bps / div) : 0;
  // The drain takes over below 384 Ki queued paths (measured sweep, profiles/README.md, round 2): per ray the
  // kernel is SLOWER than the wavefront (a whole 60 M-path fill through it: 35.0 vs 23.2 ms on cfg2), so it only
  // pays where the wavefront is launch-bound - small frames (cfg1: 0.79 -> 0.36 ms) and the last, nearly
  // empty iterations of a drain; larger thresholds lose (cfg3 at 2 Mi: +7 %).  SRT_TAIL_MAX overrides for A/B runs.
  static const long long tail_env = getenv("SRT_TAIL_MAX") ? atoll(getenv("SRT_TAIL_MAX")) : -1;
  const int tail_max = !have_tail ? 0 : (tail_env >= 0 ? (int)std::min<long long>(tail_env, 1ll << 30) : (384 << 10));
  static const long long slow_env = getenv("SRT_TAIL_SLOW") ? atoll(getenv("SRT_TAIL_SLOW")) : -1;
  const int tail_slow = !have_tail ? 0 : (slow_env >= 0 ? (int)std::min<long long>(slow_env, 1ll << 30) : (16 << 20));
  auto launch_tail = [&](int g, int parity) {
    tv.fn<<<tail_grid, EXT_THREADS, tv.smem, stream>>>(L.sc, pk, g, parity, tail_max, tail_slow, W.ray_o[g], W.ray_d[g], W.state[g], W.accum64, ctrl);
    k_tail_done<<<1, 1, 0, stream>>>(g, parity, tail_max, tail_slow, ctrl);
    launches += 2;
  };
  if (W.own_accum) WCK(cudaMemsetAsync(W.accum64, 0, sizeof(unsigned long long) * 3 * (size_t)npix, stream));
  k_ctrl_init<<<1, 1, 0, stream>>>(ctrl, total, p.spp_begin); ++launches;
  k_regen<<<regen_grid, 256, 0, stream>>>(L.cam, pk, npix, cap, 0, 0, W.ray_o[0], W.ray_d[0], W.state[0], ctrl); ++launches;
  if (have_tail && !profile && total <= (unsigned long long)tail_max && total <= (unsigned long long)cap) {
    // small frame (cfg1: 320k paths): the whole frame is one wave - regen + tail, no iteration graph, no polling
    launch_tail(0, 1);
  } else {
    if (profile) { WCK(cudaEventCreate(&e0)); WCK(cudaEventCreate(&e1)); WCK(cudaEventCreate(&e2)); }
    const int BATCH = 8;                          // even, so every batch has identical launch parameters
    int g = 0, parity = 1;
    auto enqueue_batch = [&]() {
      for (int k = 0; k < BATCH; ++k) {
        if (have_tail && !profile && (k & 3) == 0) launch_tail(g, parity);     // twice per batch: takes over as soon as the queue is small
        if (profile) cudaEventRecord(e0, stream);
        launches += srt_launch_extend(L, W.ray_o[g], W.ray_d[g], W.state[g], W.hit, &ctrl->qcount[g], 0, p.t_min, SRT_MAX_FLOAT, p.seed, stream);
        if (profile) cudaEventRecord(e1, stream);
        shade_fn<<<shade_grid, SHD_THREADS, 0, stream>>>(L.sc, pk, g, W.ray_o[g], W.ray_d[g], W.state[g], W.hit,
                                                                  W.ray_o[g ^ 1], W.ray_d[g ^ 1], W.state[g ^ 1], W.accum64, ctrl);
        if (profile) { cudaEventRecord(e2, stream); cudaEventSynchronize(e2); float a = 0.f, b = 0.f; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2); acc_ext += a; acc_shd += b; ++n_ext; }
        k_regen<<<regen_grid, 256, 0, stream>>>(L.cam, pk, npix, cap, g ^ 1, parity, W.ray_o[g ^ 1], W.ray_d[g ^ 1], W.state[g ^ 1], ctrl);
        launches += 2;
        g ^= 1; parity ^= 1;
      }
    };
    // The batch of 8 iterations (2 + 24 launches) is captured once into a CUDA graph and replayed: the
    // deep-path drain tail and small frames are launch-bound otherwise.  The executable graph is
    // CACHED on the scene (W.graph) and reused by later calls with the same scene tables, buffers
    // and parameters; only a re-commit or a change of size / estimator / quirks / seed re-captures.
    cudaGraphExec_t gexec = nullptr;
    int launches_per_batch = 0;
    if (!profile && W.use_graph && W.graph) {
      const ExtendVariant e = extend_variant(L);  // function attributes / occupancy query outside the capture
      struct Key { DScene sc; DCamera cam; SrtRenderParams p; WaveBuffers w; void* fn; void* sfn; size_t smem; int bps, device, tail_max, grid_div; } key;
      std::memset(&key, 0, sizeof(key));
      key.sc = L.sc; key.cam = L.cam; key.p = pk; key.fn = (void*)e.fn; key.sfn = (void*)shade_fn; key.smem = e.smem; key.bps = e.bps; key.device = L.device; key.tail_max = tail_max + 7 * tail_slow; key.grid_div = div;
      key.w.capacity = W.capacity; key.w.hit = W.hit; key.w.accum64 = W.accum64; key.w.ctrl = W.ctrl;
      for (int k = 0; k < 2; ++k) { key.w.ray_o[k] = W.ray_o[k]; key.w.ray_d[k] = W.ray_d[k]; key.w.state[k] = W.state[k]; }
      GraphCache& C = *W.graph;
      if (!C.exec || C.key.size() != sizeof(key) || std::memcmp(C.key.data(), &key, sizeof(key)) != 0) {
        srt_graph_cache_release(C);
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          const int l0 = launches;
          enqueue_batch();
          C.launches_per_batch = launches - l0; launches = l0;
          cudaGraph_t graph = nullptr; cudaGraphExec_t ex = nullptr;
          if (cudaStreamEndCapture(stream, &graph) == cudaSuccess && cudaGraphInstantiate(&ex, graph, 0) == cudaSuccess) {
            C.graph = graph; C.exec = ex; C.key.assign((const unsigned char*)&key, (const unsigned char*)&key + sizeof(key));
          } else { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); }
        } else cudaGetLastError();
      }
      gexec = (cudaGraphExec_t)C.exec; launches_per_batch = C.launches_per_batch;
    }
    // bound on the number of iterations: a queue fill lives at most max_depth + 1 iterations
    const unsigned long long max_iter = ((total + (unsigned long long)cap - 1ull) / (unsigned long long)cap + 1ull) * (unsigned long long)(p.max_depth + 2) + 64ull;
    bool done = false; int batch = 0;
    while (!done) {
      if ((unsigned long long)batch * BATCH > max_iter + 2ull * BATCH) { err = cudaErrorLaunchTimeout; goto fail; }   // the queue does not drain: give up loudly
      if (gexec) { WCK(cudaGraphLaunch(gexec, stream)); launches += launches_per_batch; }
      else { enqueue_batch(); WCK(cudaPeekAtLastError()); }
      WCK(cudaMemcpyAsync(&h[batch & 1], ctrl, sizeof(WaveCtrl), cudaMemcpyDeviceToHost, stream));
      WCK(cudaEventRecord(ev[batch & 1], stream));
      if (batch >= 1) {                             // look at the PREVIOUS batch while this one runs
        WCK(cudaEventSynchronize(ev[(batch - 1) & 1]));
        const WaveCtrl& c = h[(batch - 1) & 1];
        if (c.qcount[0] == 0 && c.qcount[1] == 0 && c.next_path[0] >= total && c.next_path[1] >= total) done = true;
      }
      ++batch;
    }
  }
  if (d_rgb_sum && W.own_accum) { k_accum_to_float<<<L.sm_count * 4, 256, 0, stream>>>(3 * npix, W.accum64, d_rgb_sum); ++launches; }
  WCK(cudaMemcpyAsync(&h[0], ctrl, sizeof(WaveCtrl), cudaMemcpyDeviceToHost, stream));
  WCK(cudaStreamSynchronize(stream));
  WCK(cudaGetLastError());
  if (profile) { cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); }
  if (stats) {
    stats->rays = h[0].rays; stats->waves = (int)h[0].iterations; stats->kernel_launches = launches;
    stats->ms_extend = acc_ext; stats->ms_shade = acc_shd; stats->extend_launches = n_ext;
    stats->nonfinite = h[0].nonfinite; stats->tail_runs = h[0].tail_runs;
    for (int k = 0; k < 8; ++k) stats->rays_per_bounce[k] = h[0].bounce_hist[k];
  }
  return launches;
fail:
  if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (e2) cudaEventDestroy(e2);
  { cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone; if (cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) { cudaGraph_t gdead = nullptr; cudaStreamEndCapture(stream, &gdead); if (gdead) cudaGraphDestroy(gdead); } }
  if (cuda_err) *cuda_err = err;
  return SRT_ERR_CUDA;
}
#undef WCK

// Adds the 64-bit accumulator of a finished render into the float running sum (multi-GPU: after the reduce).
int srt_launch_accum_to_float(int sm_count, int n3, const unsigned long long* accum64, float* d_rgb_sum, cudaStream_t stream) {
  k_accum_to_float<<<sm_count * 4, 256, 0, stream>>>(n3, accum64, d_rgb_sum);
  return 1;
}

int srt_launch_reduce_peers(int sm_count, int n3, unsigned long long* accum64, const unsigned long long* const* peer_ptrs, int n_peers,
                            float* d_rgb_sum, float spp_total, uint8_t* d_image, cudaStream_t stream) {
  PeerPtrs pp; std::memset(&pp, 0, sizeof(pp));
  for (int k = 0; k < n_peers && k < SRT_MAX_DEVICES; ++k) pp.p[k] = peer_ptrs[k];
  k_reduce_peers<<<sm_count * 8, 256, 0, stream>>>(n3, accum64, pp, n_peers, d_rgb_sum, spp_total, d_image);
  return 1;
}
int srt_launch_complete_hits(const DScene& sc, const float4* ray_o, const float4* ray_d, const float4* hit, int n, SrtHit* d_out, cudaStream_t stream) {
  if (n > 0) k_complete_hits<<<(n + 127) / 128, 128, 0, stream>>>(sc, ray_o, ray_d, hit, n, d_out);
  return 1;
}
int srt_launch_upload_rays(const SrtRay* d_rays, int n, float4* ray_o, float4* ray_d, cudaStream_t stream) {
  if (n > 0) k_upload_rays<<<(n + 255) / 256, 256, 0, stream>>>(d_rays, n, ray_o, ray_d);
  return 1;
}
int srt_launch_resolve(const float* d_rgb_sum, int n3, int spp, uint8_t* d_image, cudaStream_t stream) {
  if (n3 > 0) k_resolve<<<(n3 + 255) / 256 > 4096 ? 4096 : (n3 + 255) / 256, 256, 0, stream>>>(d_rgb_sum, n3, (float)spp, d_image);
  return 1;
}
int srt_launch_eval_texture(const DScene& sc, int tex, const float* d_uvp5, int n, int quirks, float* d_rgb, cudaStream_t stream) {
  if (n > 0) k_eval_texture<<<(n + 127) / 128, 128, 0, stream>>>(sc, tex, d_uvp5, n, quirks, d_rgb);
  return 1;
}
int srt_launch_eval_raygen(const RenderLaunch& L, int n, const int* d_pixel, const int* d_sample, SrtRay* d_out, cudaStream_t stream) {
  if (n > 0) k_eval_raygen<<<(n + 127) / 128, 128, 0, stream>>>(L.cam, L.p, n, d_pixel, d_sample, d_out);
  return 1;
}
