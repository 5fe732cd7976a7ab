// srt_host.h — internal host-side declarations shared by lbvh.cu, wavefront.cu and srt_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "srt_device.cuh"

#define SRT_MAX_DEVICES 16
#define SRT_MAX_PIPES 2                     // concurrent streaming pipelines per render call (srt_api.cu render_impl)
#define SRT_MASK_XF_SPHERE 0x200            // prim_mask bit: a sphere / moving sphere sits under a translate / rotate-y chain

struct LbvhBuffers {
  float* d_aabb = nullptr;                 // 6 per primitive (by primitive id)
  int* d_item_prim = nullptr;              // LBVH item -> primitive id (huge primitives are left out of the tree)
  int* d_bounds = nullptr;                 // 7 ordered-int floats: cmin[3] cmax[3] S
  unsigned long long* d_keys[2] = {nullptr, nullptr};
  int* d_order[2] = {nullptr, nullptr};
  int* d_hist = nullptr;                   // 256 * nblk
  int4* d_links = nullptr;                 // n-1
  int* d_leaf_parent = nullptr;            // n
  float* d_nbox = nullptr;                 // 6 * (n-1)
  int* d_visit = nullptr;                  // n-1
  int* d_depth = nullptr;
  float4* d_nodes = nullptr;               // 4 * max(n-1, 1)
  int sorted = 0;                          // which of d_keys/d_order holds the sorted result
};

// One candidate tree of the single-CTA build (device pointers; see k_lbvh_small in lbvh.cu)
#define SRT_SMALL_MAX_ITEMS 2048
struct SrtSmallJob {
  const int* item_prim; int n;
  unsigned long long* keys; int* order; float4* nodes; int* depth; double* area; int* bounds;
};
int srt_lbvh_build_small(const float* d_aabb, const SrtSmallJob* d_jobs, int n_jobs, cudaStream_t stream);
int srt_lbvh_bounds(const DScene& sc, float cam_t0, float cam_t1, LbvhBuffers& B, cudaStream_t stream);
int srt_lbvh_build(int n_items, LbvhBuffers& B, cudaStream_t stream);
int srt_lbvh_tree_area(int n_items, LbvhBuffers& B, double* d_out3, cudaStream_t stream);

// Wavefront queues (SoA, 16-byte vectorised).  Two generations (ping-pong) of the ray/state
// arrays: shade reads generation g and writes the compacted survivors into generation g^1, regen
// appends fresh camera paths behind them.
struct WaveBuffers {
  size_t capacity = 0;                     // paths in flight
  float4* ray_o[2] = {nullptr, nullptr};   // o.xyz, time
  float4* ray_d[2] = {nullptr, nullptr};   // d.xyz, (sample << 12 | depth) as int bits
  float4* state[2] = {nullptr, nullptr};   // throughput.rgb, pixel (int bits)
  float4* hit = nullptr;                   // t, prim (int bits), u, v
  unsigned long long* accum64 = nullptr;   // W*H*3 fixed-point (2^-36) radiance sums
  void* ctrl = nullptr;                    // WaveCtrl (device)
  void* h_ctrl = nullptr;                  // 2 x WaveCtrl (pinned host)
  void* poll_events = nullptr;             // 2 x cudaEvent_t
  bool use_graph = true;                   // replay the iteration batches as a CUDA graph
  bool own_accum = true;                   // false: the caller zeroes accum64 before and converts it after (several pipelines share it)
  struct GraphCache* graph = nullptr;      // executable graph of one iteration batch, cached on the scene
};

// The captured batch of wavefront iterations.  `key` = the bytes of everything baked into the
// graph's kernel nodes (scene tables, camera, parameters without the sample range, buffers, kernel
// variant); a later render with an equal key replays `exec` instead of capturing again.
struct GraphCache {
  cudaGraphExec_t exec = nullptr; cudaGraph_t graph = nullptr;
  std::vector<unsigned char> key; int launches_per_batch = 0;
};
void srt_graph_cache_release(GraphCache& c);

struct RenderLaunch {
  DScene sc; DCamera cam; SrtRenderParams p;
  int sm_count; bool bvh_in_smem; size_t extend_smem;
  int prim_mask;                            // bit k set = primitive kind k present (extend kernel variant); SRT_MASK_XF_SPHERE
  int device;                               // CUDA device the scene lives on (index of the per-device variant cache)
  int grid_div;                             // this launch's share of the persistent grids is 1 / grid_div (concurrent pipelines)
};

// returns the number of kernel launches or a negative SrtError (*cuda_err = the failing CUDA status);
// d_rgb_sum accumulates W*H*3 floats, nullptr = leave the frame in W.accum64 (multi-GPU reduce first)
int srt_wavefront_render(const RenderLaunch& L, WaveBuffers& W, float* d_rgb_sum, cudaStream_t stream, SrtStats* stats, bool profile, cudaError_t* cuda_err);
int srt_launch_accum_to_float(int sm_count, int n3, const unsigned long long* accum64, float* d_rgb_sum, cudaStream_t stream);
// multi-GPU combine on the root device: accum64 += sum of the peers' accumulators (read over NVLink
// peer access), then rgb_sum += accum64 / 2^36 and, if d_image != nullptr, the gamma-corrected 8-bit image
int srt_launch_reduce_peers(int sm_count, int n3, unsigned long long* accum64, const unsigned long long* const* peer_ptrs /* host array */, int n_peers,
                            float* d_rgb_sum, float spp_total, uint8_t* d_image, cudaStream_t stream);
size_t srt_wave_ctrl_bytes();
void srt_extend_prepare(const RenderLaunch& L);
int srt_launch_extend(const RenderLaunch& L, const float4* ray_o, const float4* ray_d, const float4* state, float4* hit, const int* d_count, int count,
                      float tmin, float tmax, uint32_t seed, cudaStream_t stream);
int srt_launch_complete_hits(const DScene& sc, const float4* ray_o, const float4* ray_d, const float4* hit, int n, SrtHit* d_out, cudaStream_t stream);
int srt_launch_upload_rays(const SrtRay* d_rays, int n, float4* ray_o, float4* ray_d, cudaStream_t stream);
int srt_launch_resolve(const float* d_rgb_sum, int n3, int spp, uint8_t* d_image, cudaStream_t stream);
int srt_launch_eval_texture(const DScene& sc, int tex, const float* d_uvp5, int n, int quirks, float* d_rgb, cudaStream_t stream);
int srt_launch_eval_raygen(const RenderLaunch& L, int n, const int* d_pixel, const int* d_sample, SrtRay* d_out, cudaStream_t stream);
size_t srt_extend_smem_bytes(const DScene& sc);
float srt_measure_fma_tflops(int sm_count, cudaStream_t stream);
// bounds-checked build (-DSRT_BOUNDS_CHECK): violations counted by the kernels of each translation unit
unsigned long long srt_bounds_violations_wavefront(int* first);
unsigned long long srt_bounds_violations_lbvh(int* first);
