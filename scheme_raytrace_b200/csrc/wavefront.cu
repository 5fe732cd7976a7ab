// wavefront.cu — the per-sample radiance loop as a wavefront path tracer.
//
// Replaces trace-all -> get-ray -> color -> g:hit / m:scatter / m:emitted (main.scm:100-121,
// 471-491; camera.scm:80-92; geometry.scm:14-15; material.scm:15-22).  Stages per wave of
// W*H*wave_spp paths:
//   k_raygen     camera rays (Philox bounce slot 0)
//   k_extend     closest hit: stackless LBVH traversal (bit-trail, near child first), the whole
//                node array + primitive headers staged in shared memory
//   k_shade      hit-record completion, material scatter / emitted, sky on miss; survivors are
//                compacted (warp ballot + one atomic per CTA) into the other queue generation
//   k_accumulate per-pixel sum of the wave's path radiances into the running sum (*raw-data*)
// The recursive estimator `color` is run in its iterative form L = sum_k (prod_{j<k} w_j) e_k;
// in the reference every emission event (light, sky, depth cut) ends the path, so each path
// writes its radiance exactly once.
#include "srt_device.cuh"
#include "srt_host.h"

namespace {

constexpr int EXT_THREADS = 256;
constexpr int SHD_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// ray generation: path id -> (pixel, sample); main.scm:476-478 + camera.scm:80-92
__global__ void __launch_bounds__(256) k_raygen(DCamera cam, SrtRenderParams p, int npaths, int npix, int sample_base,
                                                 float4* __restrict__ ray_o, float4* __restrict__ ray_d, float4* __restrict__ state,
                                                 float4* __restrict__ path_L) {
  for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < npaths; id += gridDim.x * blockDim.x) {
    int sl = id / npix, pixel = id - sl * npix;
    int y = pixel / p.width, x = pixel - y * p.width;
    RngAddr addr{p.seed, (uint32_t)pixel, (uint32_t)(sample_base + sl), 0u};
    float4 xi = rng_block(addr, 0);
    float u = ((float)x + xi.x) / (float)p.width;       // y = 0 is the bottom row
    float v = ((float)y + xi.y) / (float)p.height;
    float3 o, d; float time;
    get_ray(cam, u, v, xi.z, addr, o, d, time);
    ray_o[id] = make_float4(o.x, o.y, o.z, time);
    ray_d[id] = make_float4(d.x, d.y, d.z, 0.0f);
    state[id] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(id));
    path_L[id] = make_float4(0.f, 0.f, 0.f, 0.f);
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

template <bool SMEM, class PrimSrc>
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
      intersect_prim(sc, ps, pend0, o, d, time, inv_a, tmin, h);
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

template <bool SMEM>
__global__ void __launch_bounds__(EXT_THREADS, 2)
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
      Hit h = traverse<true>(sc, nodes, ps, xyz(o4), xyz(d4), o4.w, tmin, tmax);
      hit[i] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
    }
  } else {
    PrimGlobal ps{sc.prim_hdr, sc.prim_a};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
      float4 o4 = ray_o[i], d4 = ray_d[i];
      Hit h = traverse<false>(sc, nodes, ps, xyz(o4), xyz(d4), o4.w, tmin, tmax);
      hit[i] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// shade: one bounce of `color` (main.scm:100-121) for every live path + compaction.
__global__ void __launch_bounds__(SHD_THREADS)
k_shade(DScene sc, SrtRenderParams p, int depth, int npix, int sample_base,
        const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ state, const float4* __restrict__ hit,
        float4* __restrict__ ray_o_next, float4* __restrict__ ray_d_next, float4* __restrict__ state_next,
        float4* __restrict__ path_L, const int* __restrict__ count_ptr, int* __restrict__ next_count) {
  __shared__ int s_warp[SHD_THREADS / 32];
  __shared__ int s_base;
  const int count = *count_ptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x) {   // block-uniform trip count
    int i = base + threadIdx.x;
    bool alive = false;
    float3 no = v3(0, 0, 0), nd = v3(0, 0, 0), thr = v3(0, 0, 0); float ntime = 0.f; int id = 0;
    if (i < count) {
      float4 h4 = hit[i], o4 = ray_o[i], d4 = ray_d[i], s4 = state[i];
      int prim = __float_as_int(h4.y);
      thr = xyz(s4); id = __float_as_int(s4.w);
      float3 o = xyz(o4), d = xyz(d4);
      if (prim < 0) {                                              // main.scm:120 sky
        float3 L = thr * sky_value(p.sky, d);
        path_L[id] = make_float4(L.x, L.y, L.z, 0.f);
      } else {
        float3 pt, n; int material;
        complete_hit(sc, prim, h4.x, o, d, o4.w, pt, n, material);
        int sl = id / npix, pixel = id - sl * npix;
        RngAddr addr{p.seed, (uint32_t)pixel, (uint32_t)(sample_base + sl), (uint32_t)(depth + 1)};
        Scatter s = scatter(sc, material, d, pt, n, h4.z, h4.w, addr, p.quirks);
        if (s.emitted.x != 0.f || s.emitted.y != 0.f || s.emitted.z != 0.f) {   // main.scm:113/119 emitted
          float3 L = thr * s.emitted;
          path_L[id] = make_float4(L.x, L.y, L.z, 0.f);
        }
        if (s.valid && depth < p.max_depth) {                      // main.scm:112
          alive = true;
          thr = thr * s.weight;
          no = pt; nd = s.dir;
          ntime = (p.quirks & SRT_Q6_SCATTER_TIME0) ? 0.0f : o4.w;   // Q6: make-ray forces time 0
        }
      }
    }
    // compaction: warp ballot -> per-warp count -> one atomic per CTA
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
      ray_d_next[pos] = make_float4(nd.x, nd.y, nd.z, 0.f);
      state_next[pos] = make_float4(thr.x, thr.y, thr.z, __int_as_float(id));
    }
    __syncthreads();
  }
}

// main.scm:480 running sum: rgb_sum[pixel] += sum over the wave's samples, in sample order
__global__ void k_accumulate(int npix, int wave_spp, const float4* __restrict__ path_L, float* __restrict__ rgb_sum) {
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    float r = 0.f, g = 0.f, b = 0.f;
    for (int s = 0; s < wave_spp; ++s) { float4 L = path_L[(size_t)s * npix + pix]; r += L.x; g += L.y; b += L.z; }
    rgb_sum[3 * (size_t)pix] += r; rgb_sum[3 * (size_t)pix + 1] += g; rgb_sum[3 * (size_t)pix + 2] += b;
  }
}

__global__ void k_wave_begin(int* counts, int ncounts, int npaths) {
  for (int i = threadIdx.x; i < ncounts; i += blockDim.x) counts[i] = (i == 0) ? npaths : 0;
}
__global__ void k_wave_end(const int* counts, int ncounts, unsigned long long* totals) {
  unsigned long long s = 0;
  for (int i = threadIdx.x; i < ncounts; i += blockDim.x) { s += (unsigned long long)counts[i]; if (i < 8) atomicAdd(&totals[1 + i], (unsigned long long)counts[i]); }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(&totals[0], s);
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

static int extend_grid(const RenderLaunch& L, int* blocks_per_sm_out) {
  // persistent grid: SM count x resident CTAs per SM (queried, depends on the staged-BVH size)
  static int cached_bps[2] = {0, 0}; static size_t cached_smem = ~(size_t)0;
  int which = L.bvh_in_smem ? 1 : 0;
  if (!cached_bps[which] || (which && cached_smem != L.extend_smem)) {
    int bps = 0;
    if (which) {
      cudaFuncSetAttribute(k_extend<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.extend_smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_extend<true>, EXT_THREADS, L.extend_smem);
      cached_smem = L.extend_smem;
    } else {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_extend<false>, EXT_THREADS, 0);
    }
    cached_bps[which] = bps < 1 ? 1 : bps;
  }
  if (blocks_per_sm_out) *blocks_per_sm_out = cached_bps[which];
  return L.sm_count * cached_bps[which];
}

int srt_launch_extend(const RenderLaunch& L, const float4* ray_o, const float4* ray_d, float4* hit, const int* d_count, int count,
                      float tmin, float tmax, cudaStream_t stream) {
  int grid = extend_grid(L, nullptr);
  if (L.bvh_in_smem) {
    k_extend<true><<<grid, EXT_THREADS, L.extend_smem, stream>>>(L.sc, ray_o, ray_d, hit, d_count, count, tmin, tmax);
  } else {
    k_extend<false><<<grid, EXT_THREADS, 0, stream>>>(L.sc, ray_o, ray_d, hit, d_count, count, tmin, tmax);
  }
  return 1;
}

int srt_wavefront_render(const RenderLaunch& L, WaveBuffers& W, float* d_rgb_sum, cudaStream_t stream,
                         int* waves_out, bool profile, float* ms_extend, float* ms_shade, int* n_extend) {
  const SrtRenderParams& p = L.p;
  const int npix = p.width * p.height;
  const int spp = p.spp_end - p.spp_begin;
  int wave_spp = (int)(W.capacity / (size_t)npix);
  if (wave_spp > spp) wave_spp = spp;
  int launches = 0, waves = 0;
  const int ncounts = p.max_depth + 2;
  const int shade_grid = L.sm_count * 8;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float acc_ext = 0.f, acc_shd = 0.f; int next_launches = 0;
  if (profile) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  for (int s0 = p.spp_begin; s0 < p.spp_end; s0 += wave_spp, ++waves) {
    int ws = (p.spp_end - s0 < wave_spp) ? (p.spp_end - s0) : wave_spp;
    int npaths = npix * ws;
    k_wave_begin<<<1, 128, 0, stream>>>(W.counts, ncounts, npaths); ++launches;
    k_raygen<<<L.sm_count * 8, 256, 0, stream>>>(L.cam, p, npaths, npix, s0, W.ray_o[0], W.ray_d[0], W.state[0], W.path_L); ++launches;
    int g = 0;
    for (int depth = 0; depth <= p.max_depth; ++depth) {
      if (profile) cudaEventRecord(e0, stream);
      launches += srt_launch_extend(L, W.ray_o[g], W.ray_d[g], W.hit, W.counts + depth, 0, p.t_min, SRT_MAX_FLOAT, stream);
      if (profile) { cudaEventRecord(e1, stream); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); acc_ext += ms; ++next_launches; cudaEventRecord(e0, stream); }
      k_shade<<<shade_grid, SHD_THREADS, 0, stream>>>(L.sc, p, depth, npix, s0, W.ray_o[g], W.ray_d[g], W.state[g], W.hit,
                                                       W.ray_o[g ^ 1], W.ray_d[g ^ 1], W.state[g ^ 1], W.path_L, W.counts + depth, W.counts + depth + 1);
      ++launches;
      if (profile) { cudaEventRecord(e1, stream); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); acc_shd += ms; }
      g ^= 1;
    }
    k_accumulate<<<L.sm_count * 4, 256, 0, stream>>>(npix, ws, W.path_L, d_rgb_sum); ++launches;
    k_wave_end<<<1, 128, 0, stream>>>(W.counts, ncounts, W.totals); ++launches;
  }
  if (profile) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  if (waves_out) *waves_out = waves;
  if (ms_extend) *ms_extend = acc_ext;
  if (ms_shade) *ms_shade = acc_shd;
  if (n_extend) *n_extend = next_launches;
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
