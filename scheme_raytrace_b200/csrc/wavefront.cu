// wavefront.cu — the per-sample radiance loop as a wavefront path tracer.
//
// Replaces trace-all -> get-ray -> color -> g:hit / m:scatter / m:emitted (main.scm:100-121,
// 471-491; camera.scm:80-92; geometry.scm:14-15; material.scm:15-22).  STREAMING wavefront: one
// queue of up to `capacity` paths is kept full; every iteration runs
//   k_extend  closest hit: stackless LBVH traversal (bit-trail, near child first), the whole
//             node array + primitive headers staged in shared memory
//   k_shade   hit-record completion, material scatter / emitted, sky on miss; terminated paths add
//             their radiance into a 64-bit fixed-point accumulator; survivors are compacted
//             (warp ballot + one atomic per CTA) into the other queue generation
//   k_regen   fresh camera paths (Philox bounce slot 0) are appended behind the survivors
// so the long thin tail of deep paths (depth <= max_depth) is paid once per frame, not once per
// batch of samples.  The recursive estimator `color` is run in its iterative form
// L = sum_k (prod_{j<k} w_j) e_k; paths carry (throughput, pixel, sample, depth).
#include "srt_device.cuh"
#include "srt_host.h"

namespace {

constexpr int EXT_THREADS = 256;
constexpr int SHD_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// Queue control block (device memory).  Fields are double-buffered by queue generation /
// iteration parity so that no kernel reads a field another CTA of the same launch writes.
//   qcount[g]     live length of queue generation g (read by extend / shade)
//   survivors[g]  compaction cursor: shade(g^1) atomically appends survivors into generation g
//   next_path[k]  first camera path not yet generated, as seen by the regen of parity k
struct WaveCtrl {
  int qcount[2]; int survivors[2];
  unsigned long long next_path[2];
  unsigned long long total_paths;
  unsigned long long rays;          // sum of qcount over iterations = closest-hit queries
  unsigned long long iterations;
};
#define SRT_ACC_SCALE 68719476736.0f   // 2^36: radiance accumulates in 64-bit fixed point

__device__ __forceinline__ void accumulate_fixed(unsigned long long* __restrict__ acc, int pixel, float3 L) {
  // order-independent (integer) accumulation: the image is bit-identical for any scheduling,
  // queue size or sample-range split.  main.scm:480 running sum.
  if (L.x != 0.f) atomicAdd(&acc[3 * (size_t)pixel + 0], (unsigned long long)__float2ll_rn(L.x * SRT_ACC_SCALE));
  if (L.y != 0.f) atomicAdd(&acc[3 * (size_t)pixel + 1], (unsigned long long)__float2ll_rn(L.y * SRT_ACC_SCALE));
  if (L.z != 0.f) atomicAdd(&acc[3 * (size_t)pixel + 2], (unsigned long long)__float2ll_rn(L.z * SRT_ACC_SCALE));
}

// ------------------------------------------------------------------------------------------------
// regen: tops queue generation g up with fresh camera paths (main.scm:476-478 + camera.scm:80-92).
// Path id -> (pixel, sample) with consecutive ids on consecutive pixels, so the appended block of
// primary rays is coherent.  Path state: ray_o = (o, time), ray_d = (d, sample << 12 | depth),
// state = (throughput, pixel).
__global__ void __launch_bounds__(256) k_regen(DCamera cam, SrtRenderParams p, int npix, int capacity, int g, int parity,
                                                float4* __restrict__ ray_o, float4* __restrict__ ray_d, float4* __restrict__ state,
                                                WaveCtrl* __restrict__ ctrl) {
  const int surv = ctrl->survivors[g];
  const unsigned long long next = ctrl->next_path[parity], total = ctrl->total_paths;
  unsigned long long room = (unsigned long long)(capacity - surv), left = total - next;
  const int n_new = (int)(room < left ? room : left);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_new; j += gridDim.x * blockDim.x) {
    unsigned long long id = next + (unsigned long long)j;
    unsigned int sl = (unsigned int)(id / (unsigned long long)npix);
    int pixel = (int)(id - (unsigned long long)sl * (unsigned long long)npix);
    int y = pixel / p.width, x = pixel - y * p.width;
    unsigned int sample = (unsigned int)p.spp_begin + sl;
    RngAddr addr{p.seed, (uint32_t)pixel, sample, 0u};
    float4 xi = rng_block(addr, 0);
    float u = ((float)x + xi.x) / (float)p.width;       // y = 0 is the bottom row
    float v = ((float)y + xi.y) / (float)p.height;
    float3 o, d; float time;
    get_ray(cam, u, v, xi.z, addr, o, d, time);
    int slot = surv + j;
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
// extend: stackless closest-hit traversal.  Node = 64 B: left box, right box, (left, right,
// parent, sibling).  Trail bit k (from the LSB) = "the far child at the k-th level above is
// still pending".  Backtracking follows parent links up to the lowest set bit and enters the
// sibling.  No per-thread stack, no local memory.
struct PrimShared {
  const int4* h; const float4* pa;
  __device__ __forceinline__ int4 hdr(int i) const { return h[i]; }
  __device__ __forceinline__ float4 a(int i) const { return pa[i]; }
};
struct PrimGlobal {
  const int4* h; const float4* pa;
  __device__ __forceinline__ int4 hdr(int i) const { return __ldg(&h[i]); }
  __device__ __forceinline__ float4 a(int i) const { return __ldg(&pa[i]); }
};

template <bool SMEM, int MASK, class PrimSrc>
__device__ __forceinline__ Hit traverse(const DScene& sc, const float4* __restrict__ nodes, const PrimSrc& ps,
                                        float3 o, float3 d, float time, float tmin, float tmax) {
  Hit h; h.t = tmax; h.prim = -1; h.u = 0.f; h.v = 0.f; h.incl = false;
  if (sc.n_prims == 0) return h;
  // reciprocal direction; (near-)zero components become +-1e18 instead of +-inf so that the FMA
  // slab form below never produces inf - inf (boxes are padded, see lbvh.cu, so the sign of
  // (b - o) * 1e18 is exact for every ray that can reach a primitive inside the box)
  const float3 inv = v3(fabsf(d.x) > 1e-18f ? 1.0f / d.x : copysignf(1e18f, d.x), fabsf(d.y) > 1e-18f ? 1.0f / d.y : copysignf(1e18f, d.y),
                        fabsf(d.z) > 1e-18f ? 1.0f / d.z : copysignf(1e18f, d.z));
  const float3 oi = v3(o.x * inv.x, o.y * inv.y, o.z * inv.z);
  const float inv_a = 1.0f / dot(d, d);
  int node = 0;
  unsigned long long trail = 0ull;
  for (;;) {
    float4 n0, n1, n2, n3;
    if (SMEM) { n0 = nodes[4 * node]; n1 = nodes[4 * node + 1]; n2 = nodes[4 * node + 2]; n3 = nodes[4 * node + 3]; }
    else { n0 = __ldg(&nodes[4 * node]); n1 = __ldg(&nodes[4 * node + 1]); n2 = __ldg(&nodes[4 * node + 2]); n3 = __ldg(&nodes[4 * node + 3]); }
    // slabs, FMA form t = b*inv - o*inv (6 FFMA per box).  Boxes are padded at build time (lbvh.cu).
    float lx0 = fmaf(n0.x, inv.x, -oi.x), lx1 = fmaf(n0.w, inv.x, -oi.x);
    float ly0 = fmaf(n0.y, inv.y, -oi.y), ly1 = fmaf(n1.x, inv.y, -oi.y);
    float lz0 = fmaf(n0.z, inv.z, -oi.z), lz1 = fmaf(n1.y, inv.z, -oi.z);
    float rx0 = fmaf(n1.z, inv.x, -oi.x), rx1 = fmaf(n2.y, inv.x, -oi.x);
    float ry0 = fmaf(n1.w, inv.y, -oi.y), ry1 = fmaf(n2.z, inv.y, -oi.y);
    float rz0 = fmaf(n2.x, inv.z, -oi.z), rz1 = fmaf(n2.w, inv.z, -oi.z);
    float lt0 = fmaxf(fmaxf(fminf(lx0, lx1), fminf(ly0, ly1)), fmaxf(fminf(lz0, lz1), tmin));
    float lt1 = fminf(fminf(fmaxf(lx0, lx1), fmaxf(ly0, ly1)), fminf(fmaxf(lz0, lz1), h.t));
    float rt0 = fmaxf(fmaxf(fminf(rx0, rx1), fminf(ry0, ry1)), fmaxf(fminf(rz0, rz1), tmin));
    float rt1 = fminf(fminf(fmaxf(rx0, rx1), fmaxf(ry0, ry1)), fminf(fmaxf(rz0, rz1), h.t));
    bool hl = lt0 <= lt1, hr = rt0 <= rt1;
    int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
    // leaves are intersected right away (one primitive per leaf)
    int pend0 = -1, pend1 = -1;
    if (hl && left < 0) { pend0 = ~left; hl = false; }
    if (hr && right < 0) { if (pend0 < 0) pend0 = ~right; else pend1 = ~right; hr = false; }
    while (pend0 >= 0) {
      intersect_prim<MASK>(sc, ps, pend0, o, d, time, inv_a, tmin, h);
      pend0 = pend1; pend1 = -1;
    }
    if (hl | hr) {
      bool both = hl & hr;
      node = (both ? (lt0 <= rt0) : hl) ? left : right;     // near child first
      trail = (trail << 1) | (both ? 1ull : 0ull);
      continue;
    }
    if (trail == 0ull) break;
    int up = __ffsll((long long)trail) - 1;
    trail >>= up;
    int par = __float_as_int(n3.z), sib = __float_as_int(n3.w);
    for (int k = 0; k < up; ++k) {
      float4 m = SMEM ? nodes[4 * par + 3] : __ldg(&nodes[4 * par + 3]);
      par = __float_as_int(m.z); sib = __float_as_int(m.w);
    }
    node = sib;
    trail ^= 1ull;
  }
  return h;
}

template <bool SMEM, int MASK>
__global__ void __launch_bounds__(EXT_THREADS, (MASK & 0x20) ? 2 : ((MASK & 0x1c) ? 3 : 4))
k_extend(DScene sc, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, float4* __restrict__ hit,
         const int* __restrict__ count_ptr, int count_fixed, float tmin, float tmax) {
  extern __shared__ float4 smem[];
  const int count = count_ptr ? *count_ptr : count_fixed;
  if (count == 0) return;
  const float4* nodes = sc.nodes;
  if (SMEM) {
    // stage the whole LBVH + primitive headers ("shared-memory staging of BVH top levels":
    // for the <= few-thousand-primitive scenes of the reference the top levels are all levels)
    int nn = 4 * sc.n_nodes, np = sc.n_prims;
    for (int i = threadIdx.x; i < nn; i += blockDim.x) smem[i] = sc.nodes[i];
    int4* sh = (int4*)(smem + nn);
    float4* sa = smem + nn + np;
    for (int i = threadIdx.x; i < np; i += blockDim.x) { sh[i] = sc.prim_hdr[i]; sa[i] = sc.prim_a[i]; }
    __syncthreads();
    nodes = smem;
    PrimShared ps{sh, sa};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
      float4 o4 = ray_o[i], d4 = ray_d[i];
      Hit h = traverse<true, MASK>(sc, nodes, ps, xyz(o4), xyz(d4), o4.w, tmin, tmax);
      hit[i] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
    }
  } else {
    PrimGlobal ps{sc.prim_hdr, sc.prim_a};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
      float4 o4 = ray_o[i], d4 = ray_d[i];
      Hit h = traverse<false, MASK>(sc, nodes, ps, xyz(o4), xyz(d4), o4.w, tmin, tmax);
      hit[i] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// shade: one bounce of `color` (main.scm:100-121) for every live path + compaction of survivors
// into the other queue generation (warp ballot -> per-warp count -> one atomic per CTA).
__global__ void __launch_bounds__(SHD_THREADS)
k_shade(DScene sc, SrtRenderParams p, int g,
        const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ state, const float4* __restrict__ hit,
        float4* __restrict__ ray_o_next, float4* __restrict__ ray_d_next, float4* __restrict__ state_next,
        unsigned long long* __restrict__ accum, WaveCtrl* __restrict__ ctrl) {
  __shared__ int s_warp[SHD_THREADS / 32];
  __shared__ int s_base;
  const int count = ctrl->qcount[g];
  int* next_count = &ctrl->survivors[g ^ 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x) {   // block-uniform trip count
    int i = base + threadIdx.x;
    bool alive = false;
    float3 no = v3(0, 0, 0), nd = v3(0, 0, 0), thr = v3(0, 0, 0); float ntime = 0.f; int pixel = 0, sd = 0;
    if (i < count) {
      float4 h4 = hit[i], o4 = ray_o[i], d4 = ray_d[i], s4 = state[i];
      int prim = __float_as_int(h4.y);
      thr = xyz(s4); pixel = __float_as_int(s4.w); sd = __float_as_int(d4.w);
      const int depth = sd & 0xfff; const unsigned int sample = (unsigned int)sd >> 12;
      float3 o = xyz(o4), d = xyz(d4);
      if (prim < 0) {                                              // main.scm:120 sky
        accumulate_fixed(accum, pixel, thr * sky_value(p.sky, d));
      } else {
        float3 pt, n; int material;
        complete_hit(sc, prim, h4.x, o, d, o4.w, pt, n, material);
        RngAddr addr{p.seed, (uint32_t)pixel, sample, (uint32_t)(depth + 1)};
        Scatter s = scatter(sc, material, d, pt, n, h4.z, h4.w, addr, p.quirks);
        if (s.emitted.x != 0.f || s.emitted.y != 0.f || s.emitted.z != 0.f)    // main.scm:113/119 emitted
          accumulate_fixed(accum, pixel, thr * s.emitted);
        if (s.valid && depth < p.max_depth) {                      // main.scm:112
          alive = true;
          thr = thr * s.weight;
          no = pt; nd = s.dir; sd += 1;
          ntime = (p.quirks & SRT_Q6_SCATTER_TIME0) ? 0.0f : o4.w;   // Q6: make-ray forces time 0
        }
      }
    }
    unsigned ballot = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < SHD_THREADS / 32; ++w) { int c = s_warp[w]; s_warp[w] = tot; tot += c; }
      s_base = tot ? atomicAdd(next_count, tot) : 0;
    }
    __syncthreads();
    if (alive) {
      int pos = s_base + s_warp[warp] + __popc(ballot & ((1u << lane) - 1u));
      ray_o_next[pos] = make_float4(no.x, no.y, no.z, ntime);
      ray_d_next[pos] = make_float4(nd.x, nd.y, nd.z, __int_as_float(sd));
      state_next[pos] = make_float4(thr.x, thr.y, thr.z, __int_as_float(pixel));
    }
    __syncthreads();
  }
}

// end of render: rgb_sum += fixed-point accumulator (main.scm:480 running sum, *raw-data*)
__global__ void k_accum_to_float(int n3, const unsigned long long* __restrict__ accum, float* __restrict__ rgb_sum) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x)
    rgb_sum[i] += (float)((double)(long long)accum[i] * (1.0 / 68719476736.0));
}
__global__ void k_ctrl_init(WaveCtrl* ctrl, unsigned long long total_paths) {
  ctrl->qcount[0] = ctrl->qcount[1] = 0; ctrl->survivors[0] = ctrl->survivors[1] = 0;
  ctrl->next_path[0] = ctrl->next_path[1] = 0ull; ctrl->total_paths = total_paths; ctrl->rays = 0ull; ctrl->iterations = 0ull;
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
  if (r.prim >= 0) {
    float3 p, nn; int m;
    complete_hit(sc, r.prim, h4.x, xyz(o4), xyz(d4), o4.w, p, nn, m);
    r.t = h4.x; r.u = h4.z; r.v = h4.w; r.material = m;
    int type = sc.prim_hdr[r.prim].x & 0xff;
    if (type <= SRT_PRIM_MOVING_SPHERE) sphere_uv(p, r.u, r.v);     // Q5 (dead in shading)
    r.p[0] = p.x; r.p[1] = p.y; r.p[2] = p.z; r.n[0] = nn.x; r.n[1] = nn.y; r.n[2] = nn.z;
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

}  // namespace

// =================================================================================================
size_t srt_extend_smem_bytes(const DScene& sc) { return (size_t)64 * sc.n_nodes + (size_t)32 * sc.n_prims; }

// Kernel variants by primitive mix: spheres only | spheres + moving spheres | no Bezier | all.
typedef void (*ExtendFn)(DScene, const float4*, const float4*, float4*, const int*, int, float, float);
struct ExtendVariant { ExtendFn fn; int bps; size_t smem; };
static ExtendVariant g_variants[2][4];

static int variant_of(int mask) {
  if ((mask & ~0x01) == 0) return 0;
  if ((mask & ~0x03) == 0) return 1;
  if ((mask & 0x20) == 0) return 2;
  return 3;
}
static ExtendFn variant_fn(bool smem, int v) {
  switch (v) {
    case 0: return smem ? k_extend<true, 0x01> : k_extend<false, 0x01>;
    case 1: return smem ? k_extend<true, 0x03> : k_extend<false, 0x03>;
    case 2: return smem ? k_extend<true, 0x1f> : k_extend<false, 0x1f>;
    default: return smem ? k_extend<true, SRT_MASK_ALL> : k_extend<false, SRT_MASK_ALL>;
  }
}
// persistent grid: SM count x resident CTAs per SM (queried; depends on the staged-BVH size)
static const ExtendVariant& extend_variant(const RenderLaunch& L) {
  int which = L.bvh_in_smem ? 1 : 0, v = variant_of(L.prim_mask);
  ExtendVariant& e = g_variants[which][v];
  size_t smem = which ? L.extend_smem : 0;
  if (!e.fn || e.smem != smem) {
    e.fn = variant_fn(which, v); e.smem = smem;
    if (which) cudaFuncSetAttribute(e.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, e.fn, EXT_THREADS, smem);
    e.bps = bps < 1 ? 1 : bps;
  }
  return e;
}

int srt_launch_extend(const RenderLaunch& L, const float4* ray_o, const float4* ray_d, float4* hit, const int* d_count, int count,
                      float tmin, float tmax, cudaStream_t stream) {
  const ExtendVariant& e = extend_variant(L);
  e.fn<<<L.sm_count * e.bps, EXT_THREADS, e.smem, stream>>>(L.sc, ray_o, ray_d, hit, d_count, count, tmin, tmax);
  return 1;
}

size_t srt_wave_ctrl_bytes() { return sizeof(WaveCtrl); }

// The streaming wavefront: iterate extend -> shade -> regen until every camera path of the sample
// range has been generated and the queue has drained.  The host enqueues iterations in batches
// and polls the control block (pinned copy) one batch behind, so the GPU never waits on the host.
int srt_wavefront_render(const RenderLaunch& L, WaveBuffers& W, float* d_rgb_sum, cudaStream_t stream, SrtStats* stats, bool profile) {
  const SrtRenderParams& p = L.p;
  const int npix = p.width * p.height;
  const unsigned long long total = (unsigned long long)npix * (unsigned long long)(p.spp_end - p.spp_begin);
  const int cap = (int)W.capacity;
  WaveCtrl* ctrl = (WaveCtrl*)W.ctrl;
  int launches = 0;
  const int shade_grid = L.sm_count * 8, regen_grid = L.sm_count * 8;
  cudaMemsetAsync(W.accum64, 0, sizeof(unsigned long long) * 3 * (size_t)npix, stream);
  k_ctrl_init<<<1, 1, 0, stream>>>(ctrl, total); ++launches;
  k_regen<<<regen_grid, 256, 0, stream>>>(L.cam, p, npix, cap, 0, 0, W.ray_o[0], W.ray_d[0], W.state[0], ctrl); ++launches;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  float acc_ext = 0.f, acc_shd = 0.f; int n_ext = 0;
  if (profile) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); }
  const int BATCH = 8;
  WaveCtrl* h = (WaveCtrl*)W.h_ctrl;            // pinned, 2 slots
  cudaEvent_t* ev = (cudaEvent_t*)W.poll_events;
  int g = 0, parity = 1, batch = 0;
  bool done = false;
  while (!done) {
    for (int k = 0; k < BATCH; ++k) {
      if (profile) cudaEventRecord(e0, stream);
      launches += srt_launch_extend(L, W.ray_o[g], W.ray_d[g], W.hit, &ctrl->qcount[g], 0, p.t_min, SRT_MAX_FLOAT, stream);
      if (profile) cudaEventRecord(e1, stream);
      k_shade<<<shade_grid, SHD_THREADS, 0, stream>>>(L.sc, p, g, W.ray_o[g], W.ray_d[g], W.state[g], W.hit,
                                                       W.ray_o[g ^ 1], W.ray_d[g ^ 1], W.state[g ^ 1], W.accum64, ctrl);
      if (profile) { cudaEventRecord(e2, stream); cudaEventSynchronize(e2); float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2); acc_ext += a; acc_shd += b; ++n_ext; }
      k_regen<<<regen_grid, 256, 0, stream>>>(L.cam, p, npix, cap, g ^ 1, parity, W.ray_o[g ^ 1], W.ray_d[g ^ 1], W.state[g ^ 1], ctrl);
      launches += 2;
      g ^= 1; parity ^= 1;
    }
    cudaMemcpyAsync(&h[batch & 1], ctrl, sizeof(WaveCtrl), cudaMemcpyDeviceToHost, stream);
    cudaEventRecord(ev[batch & 1], stream);
    if (batch >= 1) {                             // look at the PREVIOUS batch while this one runs
      cudaEventSynchronize(ev[(batch - 1) & 1]);
      const WaveCtrl& c = h[(batch - 1) & 1];
      if (c.qcount[0] == 0 && c.qcount[1] == 0 && c.next_path[0] >= total && c.next_path[1] >= total) done = true;
    }
    ++batch;
  }
  k_accum_to_float<<<L.sm_count * 4, 256, 0, stream>>>(3 * npix, W.accum64, d_rgb_sum); ++launches;
  cudaMemcpyAsync(&h[0], ctrl, sizeof(WaveCtrl), cudaMemcpyDeviceToHost, stream);
  cudaStreamSynchronize(stream);
  if (profile) { cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); }
  if (stats) {
    stats->rays = h[0].rays; stats->waves = (int)h[0].iterations; stats->kernel_launches = launches;
    stats->ms_extend = acc_ext; stats->ms_shade = acc_shd; stats->extend_launches = n_ext;
  }
  return launches;
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
