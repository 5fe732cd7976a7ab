// srt_api.cu — the C-ABI of include/srt.h: scene tables -> device SoA, commit (H2D + LBVH build),
// render / trace_batch entry points.  Host-side plumbing only; all arithmetic of the radiance loop
// lives in the kernels (wavefront.cu, lbvh.cu).  No CPU fallback anywhere.
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <vector>
#include <chrono>
#include <string>
#include <algorithm>
#include <cmath>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <memory>
#include <dlfcn.h>
#include "srt_host.h"

namespace {

thread_local std::string g_err;
// g_device = the device srt_init selected: new scenes are created on it.  Every scene remembers its
// own device and every entry point makes that device current, so several scenes on several GPUs can
// live in one process (srt_render_multi drives one replica per GPU from one thread each).
int g_device = -1;
bool g_create_multi = false;   // the last init call was srt_init_multi: scenes created now are replicated on every GPU
struct DeviceInfo { bool ready = false; int sm_count = 0; };
DeviceInfo g_dev[SRT_MAX_DEVICES];
std::mutex g_dev_mu;

int fail(int code, const char* fmt, ...) {
  char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  g_err = buf; return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(SRT_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

int ensure_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(SRT_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)"); }
  if (device < 0 || device >= n || device >= SRT_MAX_DEVICES) return fail(SRT_ERR_ARG, "device %d out of range (%d visible)", device, n);
  std::lock_guard<std::mutex> lock(g_dev_mu);
  if (!g_dev[device].ready) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SRT_ERR_NO_DEVICE, "device %d is sm_%d%d; libsrt is built for sm_100a only", device, prop.major, prop.minor);
    g_dev[device].sm_count = prop.multiProcessorCount; g_dev[device].ready = true;
  }
  return 0;
}
struct EventPair {     // timing events that do not leak on the early-return error paths
  cudaEvent_t a = nullptr, b = nullptr;
  cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

template <class T> struct DevBuf {
  T* p = nullptr; size_t n = 0;
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) n = count ? count : 1;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};


// One persistent host thread per replica (srt_render_multi / multi-GPU commit): a frame on an 8-GPU box is ~23 ms, and
// creating + joining seven threads twice per frame was a measurable part of the one-process overhead.
class Worker {
 public:
  Worker() : th_([this] { loop(); }) {}
  ~Worker() { { std::lock_guard<std::mutex> l(mu_); quit_ = true; } cv_.notify_all(); th_.join(); }
  void start(std::function<void()> job) { { std::lock_guard<std::mutex> l(mu_); job_ = std::move(job); busy_ = true; } cv_.notify_all(); }
  void wait() { std::unique_lock<std::mutex> l(mu_); done_.wait(l, [this] { return !busy_; }); }
 private:
  void loop() {
    for (;;) {
      std::function<void()> job;
      { std::unique_lock<std::mutex> l(mu_); cv_.wait(l, [this] { return quit_ || (busy_ && job_); }); if (quit_) return; job = std::move(job_); job_ = nullptr; }
      job();
      { std::lock_guard<std::mutex> l(mu_); busy_ = false; }
      done_.notify_all();
    }
  }
  std::mutex mu_; std::condition_variable cv_, done_; std::function<void()> job_; bool busy_ = false, quit_ = false;
  std::thread th_;
};

// ---- multi-GPU state (srt_init_multi) -------------------------------------------------------------
// One process drives n GPUs (the caller the path replaces, (trace-all scene k) main.scm:471-491, is
// one process).  NCCL is loaded with dlopen so that libsrt.so has no link-time dependency on it and
// single-GPU hosts never need it; types / enums below mirror nccl.h (ncclUint64 = 5, ncclSum = 0).
typedef struct ncclComm* NcclComm;
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GroupStart)() = nullptr; int (*GroupEnd)() = nullptr;
};
struct Multi {
  int n = 0; int devs[SRT_MAX_DEVICES];
  bool peer[SRT_MAX_DEVICES];            // root (devs[0]) can read device r's memory directly
  NcclApi nccl; NcclComm comms[SRT_MAX_DEVICES]; bool have_nccl = false; int nccl_version = 0;
  int reduce_mode = 0;                   // 0 = NCCL reduce of the integer accumulators, 1 = fused peer-read kernel on the root
};
Multi g_multi;
std::mutex g_multi_mu;

bool nccl_load(NcclApi& a) {
  if (a.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) { a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
  if (!a.lib) return false;
  a.CommInitAll = (int (*)(NcclComm*, int, const int*))dlsym(a.lib, "ncclCommInitAll");
  a.CommDestroy = (int (*)(NcclComm))dlsym(a.lib, "ncclCommDestroy");
  a.Reduce = (int (*)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t))dlsym(a.lib, "ncclReduce");
  a.GetErrorString = (const char* (*)(int))dlsym(a.lib, "ncclGetErrorString");
  a.GetVersion = (int (*)(int*))dlsym(a.lib, "ncclGetVersion");
  a.GroupStart = (int (*)())dlsym(a.lib, "ncclGroupStart"); a.GroupEnd = (int (*)())dlsym(a.lib, "ncclGroupEnd");
  if (!a.CommInitAll || !a.CommDestroy || !a.Reduce || !a.GroupStart || !a.GroupEnd) { dlclose(a.lib); a = NcclApi(); return false; }
  return true;
}
void multi_shutdown() {
  std::lock_guard<std::mutex> lock(g_multi_mu);
  if (g_multi.have_nccl) for (int r = 0; r < g_multi.n; ++r) if (g_multi.comms[r]) g_multi.nccl.CommDestroy(g_multi.comms[r]);
  g_multi.have_nccl = false; g_multi.n = 0;
}

}  // namespace

struct SrtScene {
  std::vector<SrtPrim> prims; std::vector<SrtXform> xforms; std::vector<SrtMaterial> mats; std::vector<SrtTexture> texs;
  std::vector<uint8_t> img_texels; std::vector<int4> imgs;   // image-texture data (srt_scene_set_images)
  float ranvec[768]; int32_t perm[3][256]; bool has_perlin = false;
  SrtCamera cam; bool has_cam = false;
  std::vector<int32_t> lights; std::vector<float> patches; std::vector<float4> patch_slabs;   // slab (n', d') per patch-table entry, see patch_slab()
  bool committed = false;
  // device tables
  // device tables: views into ONE allocation (d_tables), uploaded from ONE pinned staging blob (h_tables) per commit
  template <class T> struct View { T* p = nullptr; };
  View<int> d_lights, d_logical; View<float4> d_patches;
  View<int4> d_hdr; View<float4> d_a, d_b, d_c, d_d, d_xf, d_tex, d_ranvec, d_shade; View<int4> d_mats, d_imgs; View<uint8_t> d_perm, d_img_texels;
  DevBuf<unsigned char> d_tables; void* h_tables = nullptr; size_t h_tables_bytes = 0;
  void* h_scratch = nullptr; size_t h_scratch_bytes = 0;      // pinned: bounds read-back, candidate item lists + jobs, per-candidate results
  DevBuf<double> d_tree_area;
  // LBVH
  DevBuf<unsigned char> d_small;   // slots of the single-CTA candidate builds (small scenes)
  LbvhBuffers lb; DevBuf<float> d_aabb, d_nbox; DevBuf<int> d_bounds, d_order0, d_order1, d_hist, d_leaf_parent, d_visit, d_depth, d_item_prim;
  std::vector<int> item_prim, global_prims;
  DevBuf<unsigned long long> d_keys0, d_keys1; DevBuf<int4> d_links; DevBuf<float4> d_nodes;
  int n_nodes = 0, n_surf = 0, n_items = 0, bvh_depth = 0; float ms_commit = 0.f; int commit_launches = 0;
  // wavefront: SRT_MAX_PIPES independent streaming pipelines (queues, control block, stream, cached graph each) that
  // share the frame's 64-bit accumulator.  A render splits its sample range over them and runs them CONCURRENTLY on
  // half-size persistent grids, so that one pipeline's shade (waiting on HBM) and the other's extend (ALU-bound)
  // share every SM ("two pipelines").
  struct Pipe {
    WaveBuffers wb; DevBuf<float4> w_ro[2], w_rd[2], w_st[2], w_hit; DevBuf<unsigned char> w_ctrl;
    void* h_ctrl = nullptr; cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    cudaStream_t work = nullptr; cudaEvent_t ev_out = nullptr;   // graph capture needs a non-legacy stream
    GraphCache gcache;
  };
  Pipe pipe[SRT_MAX_PIPES];
  DevBuf<unsigned long long> w_accum;
  cudaEvent_t ev_in = nullptr;
  DevBuf<float> d_accum;     // staging accumulation buffer for srt_render_host / srt_render_multi
  DevBuf<uint8_t> d_image;   // 8-bit frame of srt_render_multi / srt_progressive_step
  DevBuf<float> d_prog; int prog_w = 0, prog_h = 0, prog_spp = 0;   // device-resident running sum of the progressive path (*raw-data*, main.scm:430)
  DevBuf<unsigned long long> d_stage;   // multi-GPU: peers' accumulators staged on the root when peer access is unavailable
  int device = 0;            // the CUDA device this scene lives on
  int global_small = 0;      // bit g: global_prims[g] is a small sphere (fp32 test)
  bool leaf32_ok = false;    // every sphere inside the LBVH is small against the scene (|r| < extent / 64): fp32 sphere test
  bool multi = false;        // created under srt_init_multi: commit keeps one replica per GPU, srt_render_multi uses them
  unsigned long long version = 0;       // bumped by every set_*; replicas re-commit when it differs
  std::vector<SrtScene*> replicas;      // srt_render_multi: one copy of the scene per extra GPU (owned)
  std::vector<std::unique_ptr<Worker>> workers;   // one persistent host thread per replica
  unsigned long long replica_version = ~0ull;
  cudaEvent_t ev_done = nullptr;        // multi-GPU: "this replica's accumulator is complete"
  DScene ds; DCamera dcam;
};

static void fill_dscene(SrtScene* s) {
  DScene& d = s->ds;
  std::memset(&d, 0, sizeof(d));             // padding too: the bytes are part of the graph-cache key
  d.n_prims = (int)s->prims.size(); d.n_surf = s->n_surf; d.n_nodes = s->n_nodes; d.n_items = s->n_items;
  d.global_small = s->global_small;
  d.n_global = (int)s->global_prims.size(); for (int i = 0; i < SRT_MAX_GLOBAL; ++i) d.global_prims[i] = i < d.n_global ? s->global_prims[i] : 0; d.n_xforms = (int)s->xforms.size();
  d.n_mats = (int)s->mats.size(); d.n_tex = (int)s->texs.size(); d.bvh_depth = s->bvh_depth;
  d.prim_hdr = s->d_hdr.p; d.prim_a = s->d_a.p; d.prim_b = s->d_b.p; d.prim_c = s->d_c.p; d.prim_d = s->d_d.p;
  d.xf = s->d_xf.p; d.prim_shade = s->d_shade.p; d.img_texels = s->d_img_texels.p; d.imgs = s->d_imgs.p; d.nodes = s->lb.d_nodes; d.mats = s->d_mats.p; d.tex = s->d_tex.p; d.ranvec = s->d_ranvec.p; d.perm = s->d_perm.p; d.lights = s->d_lights.p; d.n_lights = (int)s->lights.size(); d.patch_cp = s->d_patches.p; d.prim_logical = s->d_logical.p;
}

static int ensure_wave(SrtScene* s, int k, size_t paths, size_t npix) {
  SrtScene::Pipe& P = s->pipe[k];
  WaveBuffers& W = P.wb;
  if (paths > W.capacity) {
    for (int g = 0; g < 2; ++g) { CK(P.w_ro[g].ensure(paths)); CK(P.w_rd[g].ensure(paths)); CK(P.w_st[g].ensure(paths)); }
    CK(P.w_hit.ensure(paths));
    for (int g = 0; g < 2; ++g) { W.ray_o[g] = P.w_ro[g].p; W.ray_d[g] = P.w_rd[g].p; W.state[g] = P.w_st[g].p; }
    W.hit = P.w_hit.p; W.capacity = paths;
  }
  W.graph = &P.gcache;
  if (npix) CK(s->w_accum.ensure(3 * npix));
  W.accum64 = s->w_accum.p;
  if (!W.ctrl) { CK(P.w_ctrl.ensure(srt_wave_ctrl_bytes())); W.ctrl = P.w_ctrl.p; }
  if (!P.h_ctrl) {
    CK(cudaMallocHost(&P.h_ctrl, 2 * srt_wave_ctrl_bytes()));
    CK(cudaEventCreateWithFlags(&P.poll_ev[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&P.poll_ev[1], cudaEventDisableTiming));
    W.h_ctrl = P.h_ctrl; W.poll_events = P.poll_ev;
    CK(cudaStreamCreateWithFlags(&P.work, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&P.ev_out, cudaEventDisableTiming));
  }
  if (!s->ev_in) CK(cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming));
  return 0;
}

static RenderLaunch make_launch(SrtScene* s, const SrtRenderParams* p) {
  RenderLaunch L; std::memset(&L, 0, sizeof(L));
  L.sc = s->ds; L.cam = s->dcam; if (p) L.p = *p;
  L.device = s->device; L.sm_count = g_dev[s->device].sm_count; L.extend_smem = srt_extend_smem_bytes(s->ds);
  L.bvh_in_smem = L.extend_smem <= (size_t)200 * 1024;
  L.prim_mask = 0;
  for (const SrtPrim& q : s->prims) {
    L.prim_mask |= 1 << q.type;
    if (q.xform >= 0 && q.type <= SRT_PRIM_MOVING_SPHERE) L.prim_mask |= SRT_MASK_XF_SPHERE;
  }
  if (s->leaf32_ok) L.prim_mask |= SRT_MASK_LEAF32;
  return L;
}
#define USE_DEVICE(s) CK(cudaSetDevice((s)->device))
// bounds-checked build: every entry point that launched kernels ends with this
static int bounds_verdict(const char* where) {
#ifdef SRT_BOUNDS_CHECK
  int f1 = 0, f2 = 0;
  const unsigned long long v1 = srt_bounds_violations_wavefront(&f1), v2 = srt_bounds_violations_lbvh(&f2);
  if (v1 + v2) return fail(SRT_ERR_CUDA, "%s: bounds-checked build counted %llu violations (first codes %d / %d)", where, v1 + v2, f1, f2);
#endif
  (void)where;
  return 0;
}

extern "C" {

const char* srt_last_error(void) { return g_err.c_str(); }

int srt_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

int srt_init(int device) {
  if (int rc = ensure_device(device)) return rc;
  CK(cudaSetDevice(device));
  g_device = device; g_create_multi = false;
  return 0;
}
void srt_shutdown(void) { multi_shutdown(); g_device = -1; }

int srt_measure_fp32_peak(float* tflops) {
  if (!tflops) return fail(SRT_ERR_ARG, "measure_fp32_peak: null argument");
  if (g_device < 0) return fail(SRT_ERR_NO_DEVICE, "srt_init() has not succeeded");
  CK(cudaSetDevice(g_device));
  *tflops = srt_measure_fma_tflops(g_dev[g_device].sm_count, 0);
  CK(cudaGetLastError());
  return 0;
}

static SrtScene* scene_create_on(int device) {
  SrtScene* s = new (std::nothrow) SrtScene();
  if (!s) { g_err = "out of host memory"; return nullptr; }
  s->device = device;
  return s;
}
SrtScene* srt_scene_create(void) {
  if (g_device < 0) { g_err = "srt_init() has not succeeded (no CPU fallback)"; return nullptr; }
  SrtScene* s = scene_create_on(g_device);
  if (s) s->multi = g_create_multi && g_multi.n > 1;
  return s;
}

void srt_scene_destroy(SrtScene* s) {
  if (!s) return;
  s->workers.clear();                      // joins the replica threads
  for (SrtScene* r : s->replicas) srt_scene_destroy(r);
  s->replicas.clear();
  cudaSetDevice(s->device);
  s->d_image.release(); s->d_prog.release(); s->d_stage.release();
  if (s->ev_done) cudaEventDestroy(s->ev_done);
  s->d_small.release(); s->d_tree_area.release(); s->d_tables.release();
  if (s->h_tables) cudaFreeHost(s->h_tables);
  if (s->h_scratch) cudaFreeHost(s->h_scratch);
  s->d_aabb.release(); s->d_nbox.release(); s->d_bounds.release();
  s->d_order0.release(); s->d_order1.release(); s->d_hist.release(); s->d_leaf_parent.release(); s->d_item_prim.release(); s->d_visit.release(); s->d_depth.release();
  s->d_keys0.release(); s->d_keys1.release(); s->d_links.release(); s->d_nodes.release();
  for (SrtScene::Pipe& P : s->pipe) {
    srt_graph_cache_release(P.gcache);
    for (int g = 0; g < 2; ++g) { P.w_ro[g].release(); P.w_rd[g].release(); P.w_st[g].release(); }
    P.w_hit.release(); P.w_ctrl.release();
    if (P.h_ctrl) cudaFreeHost(P.h_ctrl);
    for (int i = 0; i < 2; ++i) if (P.poll_ev[i]) cudaEventDestroy(P.poll_ev[i]);
    if (P.ev_out) cudaEventDestroy(P.ev_out);
    if (P.work) cudaStreamDestroy(P.work);
  }
  s->w_accum.release(); s->d_accum.release();
  if (s->ev_in) cudaEventDestroy(s->ev_in);
  delete s;
}

int srt_scene_set_prims(SrtScene* s, const SrtPrim* p, int n) {
  if (!s || n < 0 || (n && !p)) return fail(SRT_ERR_ARG, "set_prims: bad argument");
  for (int i = 0; i < n; ++i) if (p[i].type < SRT_PRIM_SPHERE || p[i].type > SRT_PRIM_KLEIN) return fail(SRT_ERR_ARG, "prim %d: unknown type %d", i, p[i].type);
  s->prims.assign(p, p + n); s->committed = false; ++s->version; return 0;
}
int srt_scene_set_xforms(SrtScene* s, const SrtXform* p, int n) {
  if (!s || n < 0 || (n && !p)) return fail(SRT_ERR_ARG, "set_xforms: bad argument");
  s->xforms.assign(p, p + n); s->committed = false; ++s->version; return 0;
}
int srt_scene_set_materials(SrtScene* s, const SrtMaterial* p, int n) {
  if (!s || n < 0 || (n && !p)) return fail(SRT_ERR_ARG, "set_materials: bad argument");
  s->mats.assign(p, p + n); s->committed = false; ++s->version; return 0;
}
int srt_scene_set_textures(SrtScene* s, const SrtTexture* p, int n) {
  if (!s || n < 0 || (n && !p)) return fail(SRT_ERR_ARG, "set_textures: bad argument");
  s->texs.assign(p, p + n); s->committed = false; ++s->version; return 0;
}
int srt_scene_set_images(SrtScene* s, const uint8_t* texels, const int32_t* dims, int n) {
  if (!s || n < 0 || (n && (!texels || !dims))) return fail(SRT_ERR_ARG, "set_images: bad argument");
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    const int32_t nx = dims[3 * i], ny = dims[3 * i + 1], off = dims[3 * i + 2];
    if (nx < 1 || ny < 1 || off < 0) return fail(SRT_ERR_ARG, "set_images: image %d has dims %d x %d at offset %d", i, nx, ny, off);
    total = std::max(total, (size_t)off + 3 * (size_t)nx * (size_t)ny);
  }
  s->imgs.resize(n);
  for (int i = 0; i < n; ++i) s->imgs[i] = make_int4(dims[3 * i], dims[3 * i + 1], dims[3 * i + 2], 0);
  s->img_texels.assign(texels, texels + total); s->committed = false; ++s->version; return 0;
}
int srt_scene_set_perlin(SrtScene* s, const float* ranvec768, const int32_t* px, const int32_t* py, const int32_t* pz) {
  if (!s || !ranvec768 || !px || !py || !pz) return fail(SRT_ERR_ARG, "set_perlin: bad argument");
  std::memcpy(s->ranvec, ranvec768, sizeof(s->ranvec));
  std::memcpy(s->perm[0], px, 1024); std::memcpy(s->perm[1], py, 1024); std::memcpy(s->perm[2], pz, 1024);
  s->has_perlin = true; s->committed = false; ++s->version; return 0;
}
int srt_scene_set_camera(SrtScene* s, const SrtCamera* c) {
  if (!s || !c) return fail(SRT_ERR_ARG, "set_camera: bad argument");
  s->cam = *c; s->has_cam = true; s->committed = false; ++s->version; return 0;
}

static float4 patch_slab(const float* cp);
int srt_scene_set_patches(SrtScene* s, const float* cp48, int n) {
  if (!s || n < 0 || (n && !cp48)) return fail(SRT_ERR_ARG, "set_patches: bad argument");
  s->patches.assign(cp48, cp48 + 48 * (size_t)n);
  s->patch_slabs.resize((size_t)n);
  for (int i = 0; i < n; ++i) s->patch_slabs[i] = patch_slab(cp48 + 48 * (size_t)i);     // a property of the table, not of a commit
  s->committed = false; ++s->version; return 0;
}

int srt_scene_set_lights(SrtScene* s, const int32_t* prim_ids, int n) {
  if (!s || n < 0 || (n && !prim_ids)) return fail(SRT_ERR_ARG, "set_lights: bad argument");
  s->lights.assign(prim_ids, prim_ids + n); s->committed = false; ++s->version; return 0;
}

// Bounding SLAB of a leaf sub-patch, |n'.p - d'| <= 1 (returned as (n', d')): n = normal of the plane through the
// corner diagonals, [dmin, dmax] = extent of the 16 control points along it (the surface lies in their convex hull),
// padded for fp32 and for Newton's +-1e-3 domain slack.  The extend kernel intersects the ray with it inside the
// leaf's box interval before it parks the full test: 38 % of the box hits are culled (profiles/README.md, round 2).
// Degenerate normal: n' = 0, d' = 0 (never culls).
static float4 patch_slab(const float* cp) {
  auto P = [&](int k, int c) { return (double)cp[3 * k + c]; };
  double e1[3], e2[3], nn[3];
  for (int c = 0; c < 3; ++c) { e1[c] = P(15, c) - P(0, c); e2[c] = P(12, c) - P(3, c); }
  nn[0] = e1[1] * e2[2] - e1[2] * e2[1]; nn[1] = e1[2] * e2[0] - e1[0] * e2[2]; nn[2] = e1[0] * e2[1] - e1[1] * e2[0];
  const double len = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
  float4 slab = make_float4(0.f, 0.f, 0.f, 0.f);
  if (len > 1e-20) {
    const double un[3] = {nn[0] / len, nn[1] / len, nn[2] / len};
    double dmin = 1e300, dmax = -1e300, lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, amax = 0.0;
    for (int k = 0; k < 16; ++k) {
      double dk = 0.0;
      for (int c = 0; c < 3; ++c) { dk += un[c] * P(k, c); lo[c] = std::min(lo[c], P(k, c)); hi[c] = std::max(hi[c], P(k, c)); amax = std::max(amax, std::fabs(P(k, c))); }
      dmin = std::min(dmin, dk); dmax = std::max(dmax, dk);
    }
    const double diag = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
    const double half = 0.5 * (dmax - dmin) * 1.001 + 4e-3 * diag + 4e-6 * amax;
    slab = make_float4((float)(nn[0] / len / half), (float)(nn[1] / len / half), (float)(nn[2] / len / half), (float)(0.5 * (dmin + dmax) / half));
  }
  return slab;
}

static int commit_impl(SrtScene* s) {
  if (!s) return fail(SRT_ERR_ARG, "commit: null scene");
  const auto t_enter = std::chrono::steady_clock::now();     // ms_commit = host wall clock of this (synchronous) call
  if (int rc = ensure_device(s->device)) return rc;
  USE_DEVICE(s);
  const int n = (int)s->prims.size();
  // validate indices (the kernels trust the tables)
  for (int i = 0; i < n; ++i) {
    const SrtPrim& p = s->prims[i];
    if (p.material < 0 || p.material >= (int)s->mats.size()) return fail(SRT_ERR_ARG, "prim %d: material %d out of range", i, p.material);
    if (p.xform >= (int)s->xforms.size()) return fail(SRT_ERR_ARG, "prim %d: xform %d out of range", i, p.xform);
    if (p.type == SRT_PRIM_PATCH && ((int)p.p[0] < 0 || (size_t)p.p[0] >= s->patches.size() / 48 || p.xform >= 0)) return fail(SRT_ERR_ARG, "prim %d: bad patch index / instanced patch", i);
    // instance transforms (geometry.scm:465-543) are applied to rects (box faces) and spheres; a curve,
    // Klein set or medium under translate / rotate-y would be intersected untransformed: refuse it
    if (p.xform >= 0 && p.type > SRT_PRIM_YZ_RECT) return fail(SRT_ERR_ARG, "prim %d: type %d cannot be instanced (translate / rotate-y apply to spheres, rects and boxes)", i, p.type);
  }
  // surfaces first, medium boundaries (SRT_PRIM_FLAG_BOUNDARY) as a suffix
  int ns = 0; while (ns < n && !(s->prims[ns].flags & SRT_PRIM_FLAG_BOUNDARY)) ++ns;
  for (int i = ns; i < n; ++i) {
    const SrtPrim& p = s->prims[i];
    if (!(p.flags & SRT_PRIM_FLAG_BOUNDARY)) return fail(SRT_ERR_ARG, "prim %d: boundary primitives must form a suffix", i);
    if (p.type != SRT_PRIM_SPHERE && (p.type < SRT_PRIM_XY_RECT || p.type > SRT_PRIM_YZ_RECT)) return fail(SRT_ERR_ARG, "prim %d: a medium boundary must be a sphere or rect", i);
  }
  for (int i = 0; i < ns; ++i) {
    const SrtPrim& p = s->prims[i];
    if (p.type != SRT_PRIM_CONSTANT_MEDIUM) continue;
    int first = (int)p.p[1], cnt = (int)p.p[2];
    if (!(p.p[0] > 0.f) || first < ns || cnt < 1 || first + cnt > n) return fail(SRT_ERR_ARG, "prim %d: bad constant-medium parameters", i);
  }
  s->n_surf = ns;
  for (size_t i = 0; i < s->mats.size(); ++i) {
    const SrtMaterial& m = s->mats[i];
    if (m.kind != SRT_MAT_DIELECTRIC && (m.tex < 0 || m.tex >= (int)s->texs.size())) return fail(SRT_ERR_ARG, "material %zu: texture %d out of range", i, m.tex);
  }
  for (size_t i = 0; i < s->texs.size(); ++i) {
    const SrtTexture& t = s->texs[i];
    if (t.kind == SRT_TEX_IMAGE && (t.even < 0 || t.even >= (int)s->imgs.size())) return fail(SRT_ERR_ARG, "texture %zu: image %d out of range (set_images)", i, t.even);
    if (t.kind < SRT_TEX_CONSTANT || t.kind > SRT_TEX_IMAGE) return fail(SRT_ERR_ARG, "texture %zu: unknown kind %d", i, t.kind);
    if (t.kind == SRT_TEX_CHECKER && (t.even < 0 || t.odd < 0 || t.even >= (int)s->texs.size() || t.odd >= (int)s->texs.size()))
      return fail(SRT_ERR_ARG, "texture %zu: checker children out of range", i);
  }
  for (size_t i = 0; i < s->lights.size(); ++i) {
    int id = s->lights[i];
    if (id < 0 || id >= n) return fail(SRT_ERR_ARG, "light %zu: primitive %d out of range", i, id);
    const SrtPrim& p = s->prims[id];
    const bool shape_ok = p.type == SRT_PRIM_SPHERE || (p.type >= SRT_PRIM_XY_RECT && p.type <= SRT_PRIM_YZ_RECT);
    if (!shape_ok || p.xform >= 0 || (p.flags & SRT_PRIM_FLAG_BOUNDARY)) return fail(SRT_ERR_ARG, "light %zu: only un-instanced spheres and rects can be sampled (pdf.scm:28-32)", i);
  }
  cudaStream_t stream = 0;
  static const bool dbg_time = getenv("SRT_DEBUG_COMMIT_TIME") != nullptr;
  double tp[8] = {0}; int ntp = 0;
  auto stamp = [&]() { if (dbg_time && ntp < 8) tp[ntp++] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  stamp();
  // ---- host SoA staging: every table is written straight into ONE pinned blob and goes up in ONE H2D copy ------
  // (sixteen separate small copies were 0.1 ms of the commit; the device tables are views into s->d_tables)
  const size_t np16 = s->patches.size() / 3;
  const size_t nx = s->xforms.size(), nm = s->mats.size(), nt = s->texs.size();
  size_t total = 0;
  auto place = [&](size_t bytes) { const size_t off = total; total += bytes ? (bytes + 255) & ~(size_t)255 : 256; return off; };
  const size_t o_patches = place(sizeof(float4) * np16), o_lights = place(sizeof(int) * s->lights.size()), o_logical = place(sizeof(int) * (size_t)n),
               o_shade = place(sizeof(float4) * 2 * (size_t)n), o_hdr = place(sizeof(int4) * (size_t)n), o_a = place(sizeof(float4) * (size_t)n),
               o_b = place(sizeof(float4) * (size_t)n), o_c = place(sizeof(float4) * (size_t)n), o_d = place(sizeof(float4) * (size_t)n),
               o_xf = place(sizeof(float4) * 2 * nx), o_mats = place(sizeof(int4) * nm), o_tex = place(sizeof(float4) * 2 * nt),
               o_imgs = place(sizeof(int4) * s->imgs.size()), o_texels = place(s->img_texels.size()), o_ranvec = place(sizeof(float4) * 256), o_perm = place(768);
  CK(s->d_tables.ensure(total));
  if (total > s->h_tables_bytes) {
    if (s->h_tables) cudaFreeHost(s->h_tables);
    s->h_tables = nullptr; s->h_tables_bytes = 0;
    CK(cudaMallocHost(&s->h_tables, total)); s->h_tables_bytes = total;
  }
  unsigned char* H = (unsigned char*)s->h_tables;
  unsigned char* Dv = s->d_tables.p;
  s->d_patches.p = (float4*)(Dv + o_patches); s->d_lights.p = (int*)(Dv + o_lights); s->d_logical.p = (int*)(Dv + o_logical); s->d_shade.p = (float4*)(Dv + o_shade);
  s->d_hdr.p = (int4*)(Dv + o_hdr); s->d_a.p = (float4*)(Dv + o_a); s->d_b.p = (float4*)(Dv + o_b); s->d_c.p = (float4*)(Dv + o_c); s->d_d.p = (float4*)(Dv + o_d);
  s->d_xf.p = (float4*)(Dv + o_xf); s->d_mats.p = (int4*)(Dv + o_mats); s->d_tex.p = (float4*)(Dv + o_tex); s->d_imgs.p = (int4*)(Dv + o_imgs);
  s->d_img_texels.p = (unsigned char*)(Dv + o_texels); s->d_ranvec.p = (float4*)(Dv + o_ranvec); s->d_perm.p = (unsigned char*)(Dv + o_perm);
  {
    float4* pc = (float4*)(H + o_patches);
    const float* src = s->patches.data();
    for (size_t k = 0; k < np16; ++k) pc[k] = make_float4(src[3 * k], src[3 * k + 1], src[3 * k + 2], 0.f);
    if (!s->lights.empty()) std::memcpy(H + o_lights, s->lights.data(), sizeof(int) * s->lights.size());
    int4* hdr = (int4*)(H + o_hdr); float4 *a = (float4*)(H + o_a), *b = (float4*)(H + o_b), *c = (float4*)(H + o_c), *d = (float4*)(H + o_d);
    int* logical = (int*)(H + o_logical); float4* sh = (float4*)(H + o_shade);     // sh: per-primitive shading record (see DScene::prim_shade)
    const float4 zero = make_float4(0, 0, 0, 0);
    for (int i = 0; i < n; ++i) {
      const SrtPrim& p = s->prims[i]; const float* q = p.p;
      // aux: patch index for patches; the bits of the plane constant k for rects (saves the extend
      // kernel a global load per rect test: the header is staged in shared memory)
      int aux = 0;
      if (p.type == SRT_PRIM_PATCH) aux = (int)q[0];
      else if (p.type >= SRT_PRIM_XY_RECT && p.type <= SRT_PRIM_YZ_RECT) std::memcpy(&aux, &q[4], 4);
      hdr[i] = make_int4(p.type | (p.flags << 8), p.material, p.xform, aux);
      a[i] = b[i] = c[i] = d[i] = zero;
      switch (p.type) {
        case SRT_PRIM_SPHERE: a[i] = make_float4(q[0], q[1], q[2], q[3]); break;
        case SRT_PRIM_MOVING_SPHERE: a[i] = make_float4(q[0], q[1], q[2], q[3]); b[i] = make_float4(q[4], q[5], q[6], q[7]); c[i] = make_float4(q[8], 0, 0, 0); break;
        case SRT_PRIM_CONSTANT_MEDIUM: a[i] = make_float4(q[0], q[1], q[2], 0); break;
        case SRT_PRIM_PATCH: { a[i] = s->patch_slabs[(size_t)q[0]]; b[i] = make_float4(q[1], q[2], q[3] > 0.f ? q[3] : 1.0f, 0); break; }
        case SRT_PRIM_BEZIER: a[i] = make_float4(q[0], q[1], q[2], q[12]); b[i] = make_float4(q[3], q[4], q[5], 0); c[i] = make_float4(q[6], q[7], q[8], 0); d[i] = make_float4(q[9], q[10], q[11], 0); break;
        default: a[i] = make_float4(q[0], q[1], q[2], q[3]); b[i] = make_float4(q[4], 0, 0, 0); break;
      }
      logical[i] = q[15] > 0.f ? (int)q[15] - 1 : i;                 // p[15] = logical id + 1, 0 = array index
      const SrtMaterial& m = s->mats[p.material];
      float r = 1.f, g = 1.f, bl = 1.f; int is_const = 0;
      if (m.kind != SRT_MAT_DIELECTRIC && s->texs[m.tex].kind == SRT_TEX_CONSTANT) { const SrtTexture& t = s->texs[m.tex]; r = t.rgb[0]; g = t.rgb[1]; bl = t.rgb[2]; is_const = 1; }
      if (m.kind == SRT_MAT_DIELECTRIC) is_const = 1;
      float fk, ft, fc; int tex = m.tex; std::memcpy(&fk, &m.kind, 4); std::memcpy(&ft, &tex, 4); std::memcpy(&fc, &is_const, 4);
      sh[2 * i] = make_float4(r, g, bl, m.param); sh[2 * i + 1] = make_float4(fk, ft, fc, 0.f);
    }
    float4 *xf = (float4*)(H + o_xf), *tx = (float4*)(H + o_tex); int4* mt = (int4*)(H + o_mats);
    for (size_t i = 0; i < nx; ++i) { const SrtXform& x = s->xforms[i]; xf[2 * i] = make_float4(x.sin_t, x.cos_t, x.off[0], x.off[1]); xf[2 * i + 1] = make_float4(x.off[2], 0, 0, 0); }
    for (size_t i = 0; i < nm; ++i) { const SrtMaterial& m = s->mats[i]; int pb; std::memcpy(&pb, &m.param, 4); mt[i] = make_int4(m.kind, m.tex, pb, 0); }
    for (size_t i = 0; i < nt; ++i) {
      const SrtTexture& t = s->texs[i]; float fk, fe, fo; std::memcpy(&fk, &t.kind, 4); std::memcpy(&fe, &t.even, 4); std::memcpy(&fo, &t.odd, 4);
      tx[2 * i] = make_float4(fk, fe, fo, t.scale); tx[2 * i + 1] = make_float4(t.rgb[0], t.rgb[1], t.rgb[2], 0);
    }
    if (!s->imgs.empty()) std::memcpy(H + o_imgs, s->imgs.data(), sizeof(int4) * s->imgs.size());
    if (!s->img_texels.empty()) std::memcpy(H + o_texels, s->img_texels.data(), s->img_texels.size());
    float4* rv = (float4*)(H + o_ranvec); uint8_t* pm = (uint8_t*)(H + o_perm);
    for (int i = 0; i < 256; ++i) rv[i] = s->has_perlin ? make_float4(s->ranvec[3 * i], s->ranvec[3 * i + 1], s->ranvec[3 * i + 2], 0) : zero;
    for (int k = 0; k < 3; ++k) for (int i = 0; i < 256; ++i) pm[256 * k + i] = s->has_perlin ? (uint8_t)(s->perm[k][i] & 255) : (uint8_t)i;
    stamp();
    CK(cudaMemcpyAsync(s->d_tables.p, s->h_tables, total, cudaMemcpyHostToDevice, stream));
  }
  // camera (by value in kernel params)
  if (s->has_cam) {
    const SrtCamera& cm = s->cam; DCamera& dc = s->dcam;
    dc.llc = make_float3(cm.llc[0], cm.llc[1], cm.llc[2]); dc.horiz = make_float3(cm.horiz[0], cm.horiz[1], cm.horiz[2]);
    dc.vert = make_float3(cm.vert[0], cm.vert[1], cm.vert[2]); dc.origin = make_float3(cm.origin[0], cm.origin[1], cm.origin[2]);
    dc.w = make_float3(cm.w[0], cm.w[1], cm.w[2]); dc.u = make_float3(cm.u[0], cm.u[1], cm.u[2]); dc.v = make_float3(cm.v[0], cm.v[1], cm.v[2]);
    dc.lens_radius = cm.lens_radius; dc.time0 = cm.time0; dc.time1 = cm.time1;
  } else std::memset(&s->dcam, 0, sizeof(s->dcam));
  // ---- LBVH ------------------------------------------------------------------------------------
  const int nn = n ? n : 1;
  CK(s->d_aabb.ensure(6 * (size_t)nn)); CK(s->d_bounds.ensure(8)); CK(s->d_keys0.ensure(nn)); CK(s->d_keys1.ensure(nn));
  CK(s->d_order0.ensure(nn)); CK(s->d_order1.ensure(nn)); CK(s->d_hist.ensure(256 * (size_t)((nn + 255) / 256)));
  CK(s->d_links.ensure(nn)); CK(s->d_leaf_parent.ensure(nn)); CK(s->d_nbox.ensure(6 * (size_t)nn)); CK(s->d_visit.ensure(nn));
  CK(s->d_depth.ensure(1)); CK(s->d_nodes.ensure(4 * (size_t)nn)); CK(s->d_item_prim.ensure(nn));
  LbvhBuffers& B = s->lb;
  B.d_aabb = s->d_aabb.p; B.d_bounds = s->d_bounds.p; B.d_keys[0] = s->d_keys0.p; B.d_keys[1] = s->d_keys1.p;
  B.d_order[0] = s->d_order0.p; B.d_order[1] = s->d_order1.p; B.d_hist = s->d_hist.p; B.d_links = s->d_links.p;
  B.d_leaf_parent = s->d_leaf_parent.p; B.d_nbox = s->d_nbox.p; B.d_visit = s->d_visit.p; B.d_depth = s->d_depth.p; B.d_nodes = s->d_nodes.p;
  B.d_item_prim = s->d_item_prim.p;
  s->n_nodes = 1; s->n_items = 0;
  fill_dscene(s);
  // phase A: primitive AABBs on the device, read back (n x 24 B) to pick the huge primitives
  s->commit_launches = srt_lbvh_bounds(s->ds, s->dcam.time0, s->dcam.time1, B, stream);
  // pinned scratch: [bounds 24 ns | candidate item lists + job table | per-candidate results] (sized for the worst case)
  const size_t hs_bounds = (sizeof(float) * 6 * (size_t)(ns ? ns : 1) + 255) & ~(size_t)255;
  const size_t hs_cand = (size_t)(SRT_MAX_GLOBAL + 1) * (((size_t)4 * (ns ? ns : 1) + 15 & ~(size_t)15) + sizeof(SrtSmallJob) + 64) + 256;
  if (hs_bounds + hs_cand > s->h_scratch_bytes) {
    if (s->h_scratch) cudaFreeHost(s->h_scratch);
    s->h_scratch = nullptr; s->h_scratch_bytes = 0;
    CK(cudaMallocHost(&s->h_scratch, hs_bounds + hs_cand)); s->h_scratch_bytes = hs_bounds + hs_cand;
  }
  float* hb = (float*)s->h_scratch;
  unsigned char* hcand = (unsigned char*)s->h_scratch + hs_bounds;
  if (ns) CK(cudaMemcpyAsync(hb, B.d_aabb, sizeof(float) * 6 * (size_t)ns, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  stamp();
  // Which surfaces stay OUT of the tree ("global": tested by all lanes before traversal).
  //  1. Klein primitives (no bounding box upstream) - forced.
  //  2. huge ones: extent e_i (max side of AABB i) >= 0.5 E, E = max side of the union of all boxes;
  //     largest first, ties lower id.  These are what the reference's own scenes keep beside the
  //     BVH, `(list ground bvh-node)` main.scm:215-235.
  //  3. outliers among the rest (spheres / rects only, e_i >= 1.5 x the median extent of the rest,
  //     largest first), taken as a PREFIX of length k chosen by the surface-area cost of the
  //     resulting tree:  cost(k) = (sum internal SA + C_LEAF sum leaf SA) / A_ref + C_GLOBAL k,
  //     A_ref = root SA at k = 0.  A few big objects among many small ones (the three r = 1 spheres
  //     of random-scene, main.scm:84-88) inflate every ancestor box on their root path; LBVH
  //     quality is also not monotone in k (the centroid grid moves), hence measured per k on the
  //     GPU-built tree rather than guessed.  C_LEAF = 2 node steps (a divergent leaf test),
  //     C_GLOBAL = 0.25 (a coherent test by all lanes): fitted to cfg2 / cfg3 timings.
  // At most SRT_MAX_GLOBAL in all; the search only runs when >= 16 items remain.
  std::vector<int> base, extra;
  {
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    std::vector<float> ext(ns ? ns : 1);
    for (int i = 0; i < ns; ++i) { float e = 0.f; if (s->prims[i].type == SRT_PRIM_KLEIN) { ext[i] = 0.f; continue; } for (int k = 0; k < 3; ++k) { float a0 = hb[6 * i + k], a1 = hb[6 * i + 3 + k]; lo[k] = a0 < lo[k] ? a0 : lo[k]; hi[k] = a1 > hi[k] ? a1 : hi[k]; e = (a1 - a0) > e ? (a1 - a0) : e; } ext[i] = e; }
    float E = 0.f; for (int k = 0; k < 3; ++k) E = (hi[k] - lo[k]) > E ? (hi[k] - lo[k]) : E;
    std::vector<int> forced;
    for (int i = 0; i < ns; ++i) if (s->prims[i].type == SRT_PRIM_KLEIN) forced.push_back(i);
    if ((int)forced.size() > SRT_MAX_GLOBAL) return fail(SRT_ERR_ARG, "at most %d Klein primitives per scene", SRT_MAX_GLOBAL);
    for (int i = 0; i < ns; ++i) if (s->prims[i].type != SRT_PRIM_KLEIN && ns > 2 && ext[i] >= 0.5f * E) base.push_back(i);
    std::stable_sort(base.begin(), base.end(), [&](int x, int y) { return ext[x] > ext[y]; });
    if ((int)(base.size() + forced.size()) > SRT_MAX_GLOBAL) base.resize(SRT_MAX_GLOBAL - forced.size());
    base.insert(base.end(), forced.begin(), forced.end());
    std::vector<char> taken(ns ? ns : 1, 0);
    for (int i : base) taken[i] = 1;
    std::vector<int> rest;
    for (int i = 0; i < ns; ++i) if (!taken[i]) rest.push_back(i);
    if (rest.size() >= 16 && (int)base.size() < SRT_MAX_GLOBAL) {
      std::vector<float> es; for (int i : rest) es.push_back(ext[i]);
      std::nth_element(es.begin(), es.begin() + es.size() / 2, es.end());
      const float med = es[es.size() / 2];
      for (int i : rest) if (s->prims[i].type <= SRT_PRIM_YZ_RECT && ext[i] >= 1.5f * med) extra.push_back(i);
      std::stable_sort(extra.begin(), extra.end(), [&](int x, int y) { return ext[x] > ext[y]; });
      if (extra.size() > SRT_MAX_GLOBAL - base.size()) extra.resize(SRT_MAX_GLOBAL - base.size());
    }
  }
  auto build_with = [&](int k) -> int {          // tree over everything but base + extra[0..k)
    std::vector<int> g(base); g.insert(g.end(), extra.begin(), extra.begin() + k);
    std::sort(g.begin(), g.end());
    s->global_prims = g; s->item_prim.clear();
    size_t gi = 0;
    for (int i = 0; i < ns; ++i) { if (gi < g.size() && g[gi] == i) { ++gi; continue; } s->item_prim.push_back(i); }
    const int nitems = (int)s->item_prim.size();
    if (nitems && cudaMemcpyAsync(B.d_item_prim, s->item_prim.data(), sizeof(int) * (size_t)nitems, cudaMemcpyHostToDevice, stream) != cudaSuccess) return -1;
    s->n_items = nitems; s->n_nodes = nitems > 1 ? nitems - 1 : 1;
    s->commit_launches += srt_lbvh_build(nitems, B, stream);
    return nitems;
  };
  int chosen = 0, small_depth = -1;        // small_depth: the chosen tree's depth when the one-launch path already read it back
  const double C_LEAF = 2.0, C_GLOBAL = 0.25;
  const int nk = (int)extra.size() + 1;
  auto pick = [&](const std::vector<double>& area) {
    double best = 0.0;
    for (int k = 0; k < nk; ++k) {
      const double c = (area[3 * k] + C_LEAF * area[3 * k + 1]) / area[2] + C_GLOBAL * k;
      if (getenv("SRT_DEBUG_COMMIT")) std::fprintf(stderr, "[srt commit] k=%d prim=%d internal %.3f leaf %.3f cost %.3f\n", k, k ? extra[k - 1] : -1, area[3 * k] / area[2], area[3 * k + 1] / area[2], c);
      if (k == 0 || c < best) { best = c; chosen = k; }
    }
  };
  const int n0 = ns - (int)base.size();                    // items of candidate k = 0; candidate k has n0 - k
  static const bool small_off = getenv("SRT_NO_SMALL_LBVH") != nullptr;              // A/B switch
  if (!small_off && n0 <= SRT_SMALL_MAX_ITEMS && n0 - (nk - 1) >= 2) {
    // Small scene (every scene of the reference): all nk candidate trees are built by ONE launch, one CTA each
    // (k_lbvh_small), into per-candidate slots of one scratch allocation; the chosen slot IS the committed tree.
    const size_t N = (size_t)n0;
    auto up = [](size_t x) { return (x + 15) & ~(size_t)15; };
    // device scratch: [item lists of all candidates | job table | results (64 B per candidate: area[3] f64, depth, bounds[8])] then
    // one slot of build buffers per candidate; the front part mirrors the pinned host block, so it goes up in ONE copy and the
    // results come back in ONE copy
    const size_t o_keys = 0, o_order = up(o_keys + 8 * N), o_nodes = up(o_order + 4 * N), slot = up(o_nodes + 64 * N);   // (links, leaf parents, node boxes: shared memory)
    const size_t items_stride = up(4 * N), o_jobs = items_stride * (size_t)nk, o_res = up(o_jobs + sizeof(SrtSmallJob) * (size_t)nk),
                 o_slots = up(o_res + 64 * (size_t)nk), bytes = o_slots + slot * (size_t)nk;
    CK(s->d_small.ensure(bytes));
    unsigned char* D = s->d_small.p;
    std::vector<std::vector<int>> globs(nk);
    std::vector<int> n_of(nk);
    for (int k = 0; k < nk; ++k) {
      std::vector<int> g(base); g.insert(g.end(), extra.begin(), extra.begin() + k);
      std::sort(g.begin(), g.end());
      int* items = (int*)(hcand + items_stride * (size_t)k);
      size_t gi = 0; int m = 0;
      for (int i = 0; i < ns; ++i) { if (gi < g.size() && g[gi] == i) { ++gi; continue; } items[m++] = i; }
      globs[k] = g; n_of[k] = m;
      unsigned char* S = D + o_slots + slot * (size_t)k;
      unsigned char* R = D + o_res + 64 * (size_t)k;
      SrtSmallJob J;
      J.item_prim = (const int*)(D + items_stride * (size_t)k); J.n = m;
      J.keys = (unsigned long long*)(S + o_keys); J.order = (int*)(S + o_order);
      J.nodes = (float4*)(S + o_nodes); J.area = (double*)R; J.depth = (int*)(R + 24); J.bounds = (int*)(R + 32);
      std::memcpy(hcand + o_jobs + sizeof(SrtSmallJob) * (size_t)k, &J, sizeof(J));
    }
    stamp();
    CK(cudaMemcpyAsync(D, hcand, o_jobs + sizeof(SrtSmallJob) * (size_t)nk, cudaMemcpyHostToDevice, stream));
    if (srt_lbvh_build_small(B.d_aabb, (const SrtSmallJob*)(D + o_jobs), nk, stream) < 0) return fail(SRT_ERR_CUDA, "commit: the single-CTA LBVH build needs %d KB of shared memory", 120);
    s->commit_launches += 1;
    unsigned char* res = hcand + o_res;
    CK(cudaMemcpyAsync(res, D + o_res, 64 * (size_t)nk, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    stamp();
    if (nk > 1) {
      std::vector<double> area(3 * (size_t)nk);
      for (int k = 0; k < nk; ++k) std::memcpy(&area[3 * k], res + 64 * (size_t)k, 24);
      pick(area);
    }
    unsigned char* S = D + o_slots + slot * (size_t)chosen;
    unsigned char* R = D + o_res + 64 * (size_t)chosen;
    const int* items = (const int*)(hcand + items_stride * (size_t)chosen);
    s->global_prims = globs[chosen]; s->item_prim.assign(items, items + n_of[chosen]);
    s->n_items = n_of[chosen]; s->n_nodes = s->n_items - 1;
    B.d_item_prim = (int*)(D + items_stride * (size_t)chosen); B.d_keys[0] = (unsigned long long*)(S + o_keys); B.d_order[0] = (int*)(S + o_order); B.sorted = 0;
    B.d_nodes = (float4*)(S + o_nodes);
    B.d_depth = (int*)(R + 24); B.d_bounds = (int*)(R + 32);
    std::memcpy(&small_depth, res + 64 * (size_t)chosen + 24, sizeof(int));
  } else {
    if (!extra.empty()) {
      CK(s->d_tree_area.ensure(3 * (size_t)nk));
      for (int k = 0; k < nk; ++k) {
        const int nitems = build_with(k);
        if (nitems < 0) return fail(SRT_ERR_CUDA, "commit: item upload failed");
        s->commit_launches += srt_lbvh_tree_area(nitems, B, s->d_tree_area.p + 3 * k, stream);
      }
      std::vector<double> area(3 * (size_t)nk);
      CK(cudaMemcpyAsync(area.data(), s->d_tree_area.p, sizeof(double) * area.size(), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      pick(area);
    }
    if (extra.empty() || chosen != (int)extra.size()) { if (build_with(chosen) < 0) return fail(SRT_ERR_CUDA, "commit: item upload failed"); }
  }
  {   // fp32 sphere test for the tree's leaves iff every sphere in the tree is small against the scene (srt_device.cuh, isect_sphere32)
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = 0; i < ns; ++i) { if (s->prims[i].type == SRT_PRIM_KLEIN) continue; for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], hb[6 * i + k]); hi[k] = std::max(hi[k], hb[6 * i + 3 + k]); } }
    float E = 0.f; for (int k = 0; k < 3; ++k) E = std::max(E, hi[k] - lo[k]);
    s->leaf32_ok = ns > 0;
    for (int i : s->item_prim) if (s->prims[i].type <= SRT_PRIM_MOVING_SPHERE && !(std::fabs(s->prims[i].p[3]) < E * (1.0f / 64.0f))) s->leaf32_ok = false;
    s->global_small = 0;
    for (size_t g = 0; g < s->global_prims.size(); ++g) {
      const SrtPrim& q = s->prims[s->global_prims[g]];
      if (q.type <= SRT_PRIM_MOVING_SPHERE && q.xform < 0 && std::fabs(q.p[3]) < E * (1.0f / 64.0f)) s->global_small |= 1 << g;
    }
  }
  fill_dscene(s);
  CK(cudaGetLastError());
  int depth = small_depth;
  if (depth < 0) CK(cudaMemcpyAsync(&depth, B.d_depth, sizeof(int), cudaMemcpyDeviceToHost, stream));
  if (small_depth < 0) CK(cudaStreamSynchronize(stream));
  s->ms_commit = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
  stamp();
  if (dbg_time) std::fprintf(stderr, "[srt commit] us: validate+event %.1f | host staging %.1f | H2D + bounds + D2H %.1f | selection %.1f | H2D + build + D2H %.1f | finish %.1f (n = %d)\n",
                             0.0, tp[1] - tp[0], tp[2] - tp[1], tp[3] - tp[2], tp[4] - tp[3], tp[5] - tp[4], n);
  s->bvh_depth = depth; s->ds.bvh_depth = depth;
  if (depth > 64) return fail(SRT_ERR_BVH_DEPTH, "LBVH depth %d exceeds the 64-level traversal bound", depth);
  if (int rc = bounds_verdict("commit")) return rc;
  s->committed = true;
  return 0;
}

static int check_params(const SrtRenderParams* p);
static int render_impl(SrtScene* s, const SrtRenderParams* p, float* d_rgb_sum, SrtStats* stats);

// host tables of `src` -> `dst` (a replica of the scene on another GPU)
static void copy_tables(SrtScene* dst, const SrtScene* src) {
  dst->prims = src->prims; dst->xforms = src->xforms; dst->mats = src->mats; dst->texs = src->texs;
  dst->img_texels = src->img_texels; dst->imgs = src->imgs; dst->lights = src->lights; dst->patches = src->patches;
  std::memcpy(dst->ranvec, src->ranvec, sizeof(src->ranvec)); std::memcpy(dst->perm, src->perm, sizeof(src->perm));
  dst->has_perlin = src->has_perlin; dst->cam = src->cam; dst->has_cam = src->has_cam;
  dst->committed = false;
}

// H2D + GPU LBVH build.  After srt_init_multi(n) a scene on the first device is committed on every
// GPU: the replicas are built concurrently (one host thread per device) while the caller's thread
// builds the primary.
int srt_scene_commit(SrtScene* s) {
  if (!s) return fail(SRT_ERR_ARG, "commit: null scene");
  const int n = g_multi.n;
  if (!s->multi || n <= 1 || s->device != g_multi.devs[0]) return commit_impl(s);
  while ((int)s->replicas.size() < n - 1) {
    SrtScene* r = scene_create_on(g_multi.devs[s->replicas.size() + 1]);
    if (!r) return fail(SRT_ERR_CUDA, "commit: cannot create the replica for device %d", g_multi.devs[s->replicas.size() + 1]);
    s->replicas.push_back(r);
  }
  while ((int)s->workers.size() < n - 1) s->workers.emplace_back(new Worker());
  std::vector<int> rc(n, 0); std::vector<std::string> msg(n);
  for (int r = 1; r < n; ++r) {
    copy_tables(s->replicas[r - 1], s);
    s->workers[r - 1]->start([&, r] { rc[r] = commit_impl(s->replicas[r - 1]); if (rc[r]) msg[r] = g_err; });
  }
  rc[0] = commit_impl(s);
  for (int r = 1; r < n; ++r) s->workers[r - 1]->wait();
  if (rc[0]) return rc[0];
  for (int r = 1; r < n; ++r) if (rc[r]) return fail(rc[r], "commit on device %d: %s", g_multi.devs[r], msg[r].c_str());
  s->replica_version = s->version;
  return 0;
}

// ---- multi-GPU behind the boundary -----------------------------------------------
int srt_init_multi(int n_gpus) {
  const int avail = srt_device_count();
  if (avail <= 0) return fail(SRT_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
  const int n = n_gpus <= 0 ? avail : n_gpus;
  if (n > avail || n > SRT_MAX_DEVICES) return fail(SRT_ERR_ARG, "init_multi: %d GPUs requested, %d visible (max %d)", n, avail, SRT_MAX_DEVICES);
  if (g_multi.n == n) { CK(cudaSetDevice(g_multi.devs[0])); g_device = g_multi.devs[0]; g_create_multi = true; return 0; }
  multi_shutdown();
  std::lock_guard<std::mutex> lock(g_multi_mu);
  Multi& M = g_multi;
  for (int r = 0; r < n; ++r) { M.devs[r] = r; M.peer[r] = false; M.comms[r] = nullptr; if (int rc = ensure_device(r)) return rc; }
  CK(cudaSetDevice(M.devs[0]));
  for (int r = 1; r < n; ++r) {               // the root reads the peers' accumulators over NVLink
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, M.devs[0], M.devs[r]) == cudaSuccess && can) {
      cudaError_t e = cudaDeviceEnablePeerAccess(M.devs[r], 0);
      if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) M.peer[r] = true;
    }
    cudaGetLastError();
  }
  const char* mode = getenv("SRT_MULTI_REDUCE");
  M.reduce_mode = (mode && std::strcmp(mode, "p2p") == 0) ? 1 : 0;
  M.have_nccl = false;
  if (n > 1 && M.reduce_mode == 0 && nccl_load(M.nccl)) {
    int e = M.nccl.CommInitAll(M.comms, n, M.devs);
    if (e == 0) { M.have_nccl = true; if (M.nccl.GetVersion) M.nccl.GetVersion(&M.nccl_version); }
    else { for (int r = 0; r < n; ++r) M.comms[r] = nullptr; cudaGetLastError(); }
  }
  if (!M.have_nccl) M.reduce_mode = 1;        // NCCL absent: the root's fused peer-read kernel does the reduce
  CK(cudaSetDevice(M.devs[0]));
  g_device = M.devs[0]; M.n = n; g_create_multi = true;
  return 0;
}
int srt_multi_device_count(void) { return g_multi.n; }
/* 0 = one NCCL reduce of the accumulators, 1 = the root's fused peer-read kernel; NCCL version as an int (e.g. 22703) */
int srt_multi_reduce_mode(int32_t* nccl_version) { if (nccl_version) *nccl_version = g_multi.have_nccl ? g_multi.nccl_version : 0; return g_multi.reduce_mode; }

// (trace-all scene k) over n GPUs from ONE process: the frame's samples [spp_begin, spp_end) are
// split into n contiguous sample ranges, one per GPU (one host thread each drives that GPU's
// wavefront loop); the per-GPU 64-bit fixed-point accumulators are combined on the first GPU - one
// ncclReduce(uint64, sum) over NVLink, or the root's own peer-read kernel - which is exact, so the
// frame is BIT-IDENTICAL to the single-GPU frame; then rgb_sum += frame, gamma + 8-bit (main.scm:481-487).
// rgb_sum / image are HOST buffers (either may be NULL); reserved[2] == 1: rgb_sum is write-only.
int srt_render_multi(SrtScene* s, const SrtRenderParams* p, float* rgb_sum, uint8_t* image, SrtStats* stats) {
  Multi& M = g_multi;
  if (M.n < 1) return fail(SRT_ERR_ARG, "render_multi: srt_init_multi() has not succeeded");
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  if (!p) return fail(SRT_ERR_ARG, "render_multi: null argument");
  if (int rc = check_params(p)) return rc;
  if (p->spp_end <= 0) return fail(SRT_ERR_ARG, "render_multi: spp_end must be positive");
  if (s->multi && s->device != M.devs[0]) return fail(SRT_ERR_ARG, "render_multi: the scene lives on device %d, not on the first device of srt_init_multi (create it after that call)", s->device);
  const int n = s->multi ? M.n : 1;                 // a scene created under plain srt_init renders on its own GPU only
  if (n > 1 && ((int)s->replicas.size() != n - 1 || s->replica_version != s->version)) return fail(SRT_ERR_NOT_COMMITTED, "render_multi: commit the scene after srt_init_multi()");
  USE_DEVICE(s);
  const size_t npix = (size_t)p->width * p->height, n3 = 3 * npix;
  const int spp = p->spp_end - p->spp_begin;
  CK(s->d_accum.ensure(n3));
  if (!rgb_sum || p->reserved[2] == 1) CK(cudaMemsetAsync(s->d_accum.p, 0, sizeof(float) * n3, 0));
  else CK(cudaMemcpyAsync(s->d_accum.p, rgb_sum, sizeof(float) * n3, cudaMemcpyHostToDevice, 0));
  std::vector<SrtStats> st(n); std::vector<int> rc(n, 0); std::vector<std::string> msg(n);
  auto work = [&](int r) {
    SrtScene* sc = r == 0 ? s : s->replicas[r - 1];
    SrtRenderParams q = *p;
    q.spp_begin = p->spp_begin + (int)((long long)r * spp / n); q.spp_end = p->spp_begin + (int)((long long)(r + 1) * spp / n);
    rc[r] = render_impl(sc, &q, nullptr, &st[r]);
    if (rc[r]) msg[r] = g_err;
  };
  for (int r = 1; r < n; ++r) s->workers[r - 1]->start([&work, r] { work(r); });
  work(0);
  for (int r = 1; r < n; ++r) s->workers[r - 1]->wait();
  for (int r = 0; r < n; ++r) if (rc[r]) return fail(rc[r], "render on device %d: %s", M.devs[r], msg[r].c_str());
  // every rank's accumulator is complete (render_impl returns synchronised).  Combine on the root.
  USE_DEVICE(s);
  int launches = 0;
  if (image) CK(s->d_image.ensure(n3));
  if (n > 1 && M.reduce_mode == 0) {
    if (M.nccl.GroupStart() != 0) return fail(SRT_ERR_CUDA, "render_multi: ncclGroupStart failed");
    int e = 0;
    for (int r = 0; r < n && e == 0; ++r) {
      SrtScene* sc = r == 0 ? s : s->replicas[r - 1];
      e = M.nccl.Reduce(sc->w_accum.p, sc->w_accum.p, n3, 5 /* ncclUint64 */, 0 /* ncclSum */, 0, M.comms[r], sc->pipe[0].work);
    }
    const int e2 = M.nccl.GroupEnd();
    if (e || e2) return fail(SRT_ERR_CUDA, "render_multi: ncclReduce failed: %s", M.nccl.GetErrorString ? M.nccl.GetErrorString(e ? e : e2) : "?");
    USE_DEVICE(s);
    CK(cudaStreamSynchronize(s->pipe[0].work));
    launches += n;
    srt_launch_reduce_peers(g_dev[s->device].sm_count, (int)n3, s->w_accum.p, nullptr, 0, s->d_accum.p, (float)p->spp_end, image ? s->d_image.p : nullptr, 0); ++launches;
  } else {
    const unsigned long long* peers[SRT_MAX_DEVICES]; int np = 0;
    for (int r = 1; r < n; ++r) {
      SrtScene* sc = s->replicas[r - 1];
      if (M.peer[r]) peers[np++] = sc->w_accum.p;
      else {                                   // no peer access: stage the accumulator on the root first
        CK(s->d_stage.ensure((size_t)(n - 1) * n3));
        CK(cudaMemcpyPeerAsync(s->d_stage.p + (size_t)(r - 1) * n3, s->device, sc->w_accum.p, sc->device, sizeof(unsigned long long) * n3, 0));
        peers[np++] = s->d_stage.p + (size_t)(r - 1) * n3;
      }
    }
    srt_launch_reduce_peers(g_dev[s->device].sm_count, (int)n3, s->w_accum.p, peers, np, s->d_accum.p, (float)p->spp_end, image ? s->d_image.p : nullptr, 0); ++launches;
  }
  CK(cudaGetLastError());
  if (rgb_sum) CK(cudaMemcpyAsync(rgb_sum, s->d_accum.p, sizeof(float) * n3, cudaMemcpyDeviceToHost, 0));
  if (image) CK(cudaMemcpyAsync(image, s->d_image.p, n3, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    for (int r = 0; r < n; ++r) {
      stats->rays += st[r].rays; stats->paths += st[r].paths; stats->kernel_launches += st[r].kernel_launches; stats->nonfinite += st[r].nonfinite;
      stats->tail_runs += st[r].tail_runs;
      stats->ms_total = std::max(stats->ms_total, st[r].ms_total); stats->waves = std::max(stats->waves, st[r].waves);
      for (int k = 0; k < 8; ++k) stats->rays_per_bounce[k] += st[r].rays_per_bounce[k];
    }
    stats->kernel_launches += launches; stats->ms_commit = s->ms_commit; stats->bvh_nodes = s->n_nodes; stats->bvh_depth = s->bvh_depth;
  }
  return 0;
}

int srt_bvh_node_count(SrtScene* s) { return (s && s->committed) ? s->n_nodes : 0; }
int srt_bvh_readback(SrtScene* s, SrtBvhNode* nodes, int cap) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  USE_DEVICE(s);
  if (cap < s->n_nodes || !nodes) return fail(SRT_ERR_ARG, "bvh_readback: capacity %d < %d nodes", cap, s->n_nodes);
  CK(cudaMemcpy(nodes, s->lb.d_nodes, sizeof(SrtBvhNode) * (size_t)s->n_nodes, cudaMemcpyDeviceToHost));
  return 0;
}
int srt_bvh_items_readback(SrtScene* s, int32_t* item_prim, int cap, int32_t* global_prims8, int32_t* n_global) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  if (cap < s->n_items || !item_prim || !global_prims8 || !n_global) return fail(SRT_ERR_ARG, "bvh_items_readback: bad argument");
  for (int i = 0; i < s->n_items; ++i) item_prim[i] = s->item_prim[i];
  *n_global = (int)s->global_prims.size();
  for (size_t i = 0; i < s->global_prims.size(); ++i) global_prims8[i] = s->global_prims[i];
  return s->n_items;
}
int srt_bvh_keys_readback(SrtScene* s, uint64_t* keys, int32_t* order, int cap) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  USE_DEVICE(s);
  int n = s->n_items;
  if (cap < n) return fail(SRT_ERR_ARG, "bvh_keys_readback: capacity too small");
  if (n) {
    CK(cudaMemcpy(keys, s->lb.d_keys[s->lb.sorted], sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(order, s->lb.d_order[s->lb.sorted], sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  }
  return 0;
}
int srt_prim_bounds_readback(SrtScene* s, float* aabbs6, int cap) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  USE_DEVICE(s);
  int n = s->n_surf;
  if (cap < n) return fail(SRT_ERR_ARG, "prim_bounds_readback: capacity too small");
  if (n) CK(cudaMemcpy(aabbs6, s->d_aabb.p, sizeof(float) * 6 * (size_t)n, cudaMemcpyDeviceToHost));
  return 0;
}

int srt_trace_batch(SrtScene* s, const SrtRay* rays, int n, float t_min, float t_max, SrtHit* out) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  USE_DEVICE(s);
  if (n < 0 || (n && (!rays || !out))) return fail(SRT_ERR_ARG, "trace_batch: bad argument");
  if (n == 0) return 0;
  if (int rc = ensure_wave(s, 0, (size_t)n, 0)) return rc;
  const WaveBuffers& wb0 = s->pipe[0].wb;
  cudaStream_t stream = 0;
  DevBuf<SrtRay> d_rays; DevBuf<SrtHit> d_out;
  CK(d_rays.ensure(n)); CK(d_out.ensure(n));
  CK(cudaMemcpyAsync(d_rays.p, rays, sizeof(SrtRay) * (size_t)n, cudaMemcpyHostToDevice, stream));
  RenderLaunch L = make_launch(s, nullptr);
  srt_launch_upload_rays(d_rays.p, n, wb0.ray_o[0], wb0.ray_d[0], stream);
  srt_launch_extend(L, wb0.ray_o[0], wb0.ray_d[0], nullptr, wb0.hit, nullptr, n, t_min, t_max, 0u, stream);   // the renderer's extend kernel
  srt_launch_complete_hits(s->ds, wb0.ray_o[0], wb0.ray_d[0], wb0.hit, n, d_out.p, stream);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d_out.p, sizeof(SrtHit) * (size_t)n, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  d_rays.release(); d_out.release();
  return bounds_verdict("trace_batch");
}

static int check_params(const SrtRenderParams* p) {
  if (p->width <= 0 || p->height <= 0 || p->spp_begin < 0 || p->spp_end < p->spp_begin || p->max_depth < 0 || p->max_depth > 4096)
    return fail(SRT_ERR_ARG, "render: bad parameters (%dx%d, spp [%d,%d), depth %d)", p->width, p->height, p->spp_begin, p->spp_end, p->max_depth);
  if (p->max_depth > 4095 || p->spp_end > (1 << 20)) return fail(SRT_ERR_ARG, "render: max_depth <= 4095 and spp_end <= 2^20 (packed path state)");
  if (p->estimator != SRT_EST_REFERENCE && p->estimator != SRT_EST_MIXTURE) return fail(SRT_ERR_ARG, "render: unknown estimator %d", p->estimator);
  if (p->sky != SRT_SKY_GRADIENT && p->sky != SRT_SKY_BLACK) return fail(SRT_ERR_ARG, "render: unknown sky %d", p->sky);
  if ((size_t)p->width * (size_t)p->height > ((size_t)1 << 30)) return fail(SRT_ERR_ARG, "render: frame of %d x %d pixels too large", p->width, p->height);
  return 0;
}

// Renders samples [spp_begin, spp_end) on the scene's device.  d_rgb_sum != nullptr: the frame is
// added to that float buffer; nullptr: it is left in the scene's 64-bit accumulator (multi-GPU).
static int render_impl(SrtScene* s, const SrtRenderParams* p, float* d_rgb_sum, SrtStats* stats) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  if (!p) return fail(SRT_ERR_ARG, "render: null argument");
  if (int rc = check_params(p)) return rc;
  if (!s->has_cam) return fail(SRT_ERR_ARG, "render: no camera set");
  USE_DEVICE(s);
  const size_t npix = (size_t)p->width * p->height;
  const int spp = p->spp_end - p->spp_begin;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  // Pipelines: the sample range CAN be split over SRT_MAX_PIPES concurrent streaming wavefronts (own queues, stream,
  // cached graph; shared integer accumulator), each on a 1/n share of the persistent grids, so that one pipeline's
  // shade (waiting on HBM) overlaps the other's extend (ALU-bound) on every SM.  Measured on B200 (profiles/README.md,
  // round 2): cfg2 +2.5 %, but cfg3 -15 % and cfg4 -20 % (kernels of the two streams co-reside poorly: 3 resident
  // extend CTAs do not halve, and the shared-memory carve-out differs between extend and shade) - so ONE pipeline is
  // the default; SRT_PIPES=2 or params.reserved[4] == 2 selects two (kept bit-identical by tests/test_gpu_round2.py).
  RenderLaunch L = make_launch(s, p);
  static const int pipes_env = getenv("SRT_PIPES") ? atoi(getenv("SRT_PIPES")) : 1;
  const size_t total = npix * (size_t)spp;
  int npipes = 1;
  const int want = p->reserved[4] == 2 ? 2 : (p->reserved[4] == 1 ? 1 : pipes_env);
  if (!(L.prim_mask & 0x1e0) && p->reserved[0] != 1 && spp >= 2 && total >= ((size_t)16 << 20))
    npipes = std::max(1, std::min(std::min(want, SRT_MAX_PIPES), spp));
  L.grid_div = npipes;
  // queue sizing: 64 Mi paths in flight in all (SoA ray/state/hit queues = 112 B/path = 7.5 GB of the 180 GB
  // HBM), never more than the job.  Measured on cfg2: 8 Mi -> 6.94, 64 Mi -> 7.31 Grays/s (fewer
  // iterations, and the depth-50 drain tail weighs less).
  size_t caps[SRT_MAX_PIPES]; SrtRenderParams pp[SRT_MAX_PIPES];
  for (int k = 0; k < npipes; ++k) {
    pp[k] = *p;
    pp[k].spp_begin = p->spp_begin + (int)((long long)k * spp / npipes); pp[k].spp_end = p->spp_begin + (int)((long long)(k + 1) * spp / npipes);
    const size_t tk = npix * (size_t)(pp[k].spp_end - pp[k].spp_begin);
    size_t cap = p->wave_spp > 0 ? npix * (size_t)p->wave_spp : (size_t)(64u << 20);
    cap = std::max<size_t>(1, std::min(cap / npipes, tk));
    if (cap > ((size_t)1 << 30)) return fail(SRT_ERR_ARG, "render: queue of %zu paths too large", cap);
    caps[k] = cap;
    if (int rc = ensure_wave(s, k, cap, npix)) return rc;
  }
  if (spp == 0) {                              // nothing to trace; a multi-GPU rank still owes a zero accumulator
    if (!d_rgb_sum) { CK(cudaMemsetAsync(s->w_accum.p, 0, sizeof(unsigned long long) * 3 * npix, 0)); CK(cudaStreamSynchronize(0)); }
    return 0;
  }
  // The loops run on the pipes' own non-blocking streams (CUDA-graph capture is not allowed on the legacy
  // default stream) but are ordered inside stream 0 by events on both sides, so callers that bracket the call
  // with events / work on the default stream (torch's current stream) stay correct.
  const bool profile = p->reserved[0] == 1;
  EventPair tm; CK(tm.create());
  CK(cudaEventRecord(tm.a, 0));
  CK(cudaMemsetAsync(s->w_accum.p, 0, sizeof(unsigned long long) * 3 * npix, 0));
  CK(cudaEventRecord(s->ev_in, 0));
  SrtStats st[SRT_MAX_PIPES]; int rcs[SRT_MAX_PIPES]; cudaError_t werr[SRT_MAX_PIPES];
  auto run = [&](int k) {
    cudaSetDevice(s->device);
    SrtScene::Pipe& P = s->pipe[k];
    std::memset(&st[k], 0, sizeof(SrtStats)); werr[k] = cudaSuccess;
    if (cudaStreamWaitEvent(P.work, s->ev_in, 0) != cudaSuccess) { rcs[k] = SRT_ERR_CUDA; werr[k] = cudaGetLastError(); return; }
    RenderLaunch Lk = L; Lk.p = pp[k];
    WaveBuffers W = P.wb; W.capacity = caps[k]; W.use_graph = p->reserved[1] != 1; W.own_accum = false;
    rcs[k] = srt_wavefront_render(Lk, W, nullptr, P.work, &st[k], profile, &werr[k]);
    if (rcs[k] >= 0) cudaEventRecord(P.ev_out, P.work);
  };
  {
    std::vector<std::thread> th;
    for (int k = 1; k < npipes; ++k) th.emplace_back(run, k);
    run(0);
    for (std::thread& t : th) t.join();
  }
  for (int k = 0; k < npipes; ++k)
    if (rcs[k] < 0) return fail(SRT_ERR_CUDA, "render: the wavefront loop failed: %s", werr[k] == cudaErrorLaunchTimeout ? "the path queue did not drain within its iteration bound" : cudaGetErrorString(werr[k]));
  for (int k = 0; k < npipes; ++k) CK(cudaStreamWaitEvent(0, s->pipe[k].ev_out, 0));
  int extra = 0;
  if (d_rgb_sum) extra += srt_launch_accum_to_float(L.sm_count, (int)(3 * npix), s->w_accum.p, d_rgb_sum, 0);
  CK(cudaEventRecord(tm.b, 0));
  CK(cudaEventSynchronize(tm.b));
  CK(cudaGetLastError());
  float ms = 0.f; CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
  if (stats) {
    for (int k = 0; k < npipes; ++k) {
      stats->rays += st[k].rays; stats->kernel_launches += st[k].kernel_launches; stats->nonfinite += st[k].nonfinite; stats->tail_runs += st[k].tail_runs;
      stats->waves = std::max(stats->waves, st[k].waves); stats->ms_extend += st[k].ms_extend; stats->ms_shade += st[k].ms_shade; stats->extend_launches += st[k].extend_launches;
      for (int b = 0; b < 8; ++b) stats->rays_per_bounce[b] += st[k].rays_per_bounce[b];
    }
    stats->kernel_launches += extra;
    stats->paths = (uint64_t)total; stats->ms_total = ms; stats->ms_commit = s->ms_commit;
    stats->bvh_nodes = s->n_nodes; stats->bvh_depth = s->bvh_depth; stats->pipes = npipes;
  }
  return bounds_verdict("render");
}

int srt_render_device(SrtScene* s, const SrtRenderParams* p, float* d_rgb_sum, SrtStats* stats) {
  if (!d_rgb_sum) return fail(SRT_ERR_ARG, "render_device: null accumulation buffer");
  return render_impl(s, p, d_rgb_sum, stats);
}

// params.reserved[2] == 1: rgb_sum is WRITE-ONLY (the frame starts from zero, main.scm:430), which
// saves the upload of the running sum.
int srt_render_host(SrtScene* s, const SrtRenderParams* p, float* rgb_sum, SrtStats* stats) {
  if (!s || !p || !rgb_sum) return fail(SRT_ERR_ARG, "render_host: null argument");
  if (p->width <= 0 || p->height <= 0) return fail(SRT_ERR_ARG, "render_host: bad size");
  USE_DEVICE(s);
  size_t n3 = (size_t)p->width * p->height * 3;
  CK(s->d_accum.ensure(n3));
  if (p->reserved[2] == 1) CK(cudaMemsetAsync(s->d_accum.p, 0, sizeof(float) * n3, 0));
  else CK(cudaMemcpyAsync(s->d_accum.p, rgb_sum, sizeof(float) * n3, cudaMemcpyHostToDevice, 0));   // running sum in (*raw-data*)
  if (int rc = render_impl(s, p, s->d_accum.p, stats)) return rc;
  CK(cudaMemcpy(rgb_sum, s->d_accum.p, sizeof(float) * n3, cudaMemcpyDeviceToHost));
  return 0;
}

// Progressive path (main.scm:452-469 trace-line passes, :493-503 the displayed image): the running
// sum *raw-data* stays RESIDENT on the device between calls; one call renders samples
// [spp_begin, spp_end) into it and returns only the 8-bit frame for the spp_end samples so far.
// spp_begin == 0 (or a change of size) starts a new accumulation.
int srt_progressive_step(SrtScene* s, const SrtRenderParams* p, uint8_t* image, SrtStats* stats) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  if (!p || !image) return fail(SRT_ERR_ARG, "progressive_step: null argument");
  if (int rc = check_params(p)) return rc;
  if (p->spp_end <= 0) return fail(SRT_ERR_ARG, "progressive_step: spp_end must be positive");
  USE_DEVICE(s);
  const size_t n3 = (size_t)p->width * p->height * 3;
  if (p->spp_begin == 0 || s->prog_w != p->width || s->prog_h != p->height || !s->d_prog.p) {
    if (p->spp_begin != 0) return fail(SRT_ERR_ARG, "progressive_step: no running sum of %d x %d to continue from sample %d", p->width, p->height, p->spp_begin);
    CK(s->d_prog.ensure(n3)); CK(cudaMemsetAsync(s->d_prog.p, 0, sizeof(float) * n3, 0));
    s->prog_w = p->width; s->prog_h = p->height; s->prog_spp = 0;
  }
  if (p->spp_begin != s->prog_spp) return fail(SRT_ERR_ARG, "progressive_step: the running sum holds %d samples, the pass starts at %d", s->prog_spp, p->spp_begin);
  if (int rc = render_impl(s, p, s->d_prog.p, stats)) return rc;
  s->prog_spp = p->spp_end;
  CK(s->d_image.ensure(n3));
  srt_launch_resolve(s->d_prog.p, (int)n3, p->spp_end, s->d_image.p, 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(image, s->d_image.p, n3, cudaMemcpyDeviceToHost));
  return 0;
}
int srt_progressive_read(SrtScene* s, float* rgb_sum, int32_t* spp) {
  if (!s || !rgb_sum || !s->d_prog.p) return fail(SRT_ERR_ARG, "progressive_read: no running sum");
  USE_DEVICE(s);
  CK(cudaMemcpy(rgb_sum, s->d_prog.p, sizeof(float) * 3 * (size_t)s->prog_w * s->prog_h, cudaMemcpyDeviceToHost));
  if (spp) *spp = s->prog_spp;
  return 0;
}

int srt_resolve_device(const float* d_rgb_sum, int width, int height, int spp, uint8_t* d_image) {
  if (!d_rgb_sum || !d_image || width <= 0 || height <= 0 || spp <= 0) return fail(SRT_ERR_ARG, "resolve: bad argument");
  srt_launch_resolve(d_rgb_sum, width * height * 3, spp, d_image, 0);
  CK(cudaGetLastError()); CK(cudaStreamSynchronize(0));
  return 0;
}
int srt_resolve_host(const float* rgb_sum, int width, int height, int spp, uint8_t* image) {
  if (!rgb_sum || !image || width <= 0 || height <= 0 || spp <= 0) return fail(SRT_ERR_ARG, "resolve: bad argument");
  if (g_device < 0) return fail(SRT_ERR_NO_DEVICE, "srt_init() has not succeeded (no CPU fallback)");
  CK(cudaSetDevice(g_device));
  size_t n3 = (size_t)width * height * 3;
  DevBuf<float> d_in; DevBuf<uint8_t> d_out;
  CK(d_in.ensure(n3)); CK(d_out.ensure(n3));
  CK(cudaMemcpy(d_in.p, rgb_sum, sizeof(float) * n3, cudaMemcpyHostToDevice));
  int rc = srt_resolve_device(d_in.p, width, height, spp, d_out.p);
  if (!rc) { cudaError_t e = cudaMemcpy(image, d_out.p, n3, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = fail(SRT_ERR_CUDA, "resolve D2H: %s", cudaGetErrorString(e)); }
  d_in.release(); d_out.release();
  return rc;
}

// main.scm:439-450 save-as-ppm: ASCII P3, header "P3\n W H\n255\n" (space before W), rows from
// y = H-1 down to 0, one "r g b\n" per pixel.
int srt_save_ppm(const char* path, const uint8_t* image, int width, int height) {
  if (!path || !image || width <= 0 || height <= 0) return fail(SRT_ERR_ARG, "save_ppm: bad argument");
  FILE* f = std::fopen(path, "w");
  if (!f) return fail(SRT_ERR_IO, "save_ppm: cannot open %s", path);
  std::fprintf(f, "P3\n %d %d\n255\n", width, height);
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      size_t i = ((size_t)(height - y - 1) * width + x) * 3;
      std::fprintf(f, "%d %d %d\n", image[i], image[i + 1], image[i + 2]);
    }
  std::fclose(f);
  return 0;
}

int srt_eval_texture(SrtScene* s, int tex, const float* uvp5, int n, int quirks, float* rgb) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  USE_DEVICE(s);
  if (tex < 0 || tex >= (int)s->texs.size() || n < 0 || (n && (!uvp5 || !rgb))) return fail(SRT_ERR_ARG, "eval_texture: bad argument");
  if (n == 0) return 0;
  DevBuf<float> d_in, d_out; CK(d_in.ensure(5 * (size_t)n)); CK(d_out.ensure(3 * (size_t)n));
  CK(cudaMemcpy(d_in.p, uvp5, sizeof(float) * 5 * (size_t)n, cudaMemcpyHostToDevice));
  srt_launch_eval_texture(s->ds, tex, d_in.p, n, quirks, d_out.p, 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(rgb, d_out.p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
  d_in.release(); d_out.release();
  return 0;
}
int srt_eval_raygen(SrtScene* s, const SrtRenderParams* p, int n, const int32_t* pixel, const int32_t* sample, SrtRay* out) {
  if (!s || !s->committed) return fail(SRT_ERR_NOT_COMMITTED, "scene not committed");
  USE_DEVICE(s);
  if (!p || n < 0 || (n && (!pixel || !sample || !out))) return fail(SRT_ERR_ARG, "eval_raygen: bad argument");
  if (n == 0) return 0;
  DevBuf<int> d_px, d_sm; DevBuf<SrtRay> d_out; CK(d_px.ensure(n)); CK(d_sm.ensure(n)); CK(d_out.ensure(n));
  CK(cudaMemcpy(d_px.p, pixel, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_sm.p, sample, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
  RenderLaunch L = make_launch(s, p);
  srt_launch_eval_raygen(L, n, d_px.p, d_sm.p, d_out.p, 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, d_out.p, sizeof(SrtRay) * (size_t)n, cudaMemcpyDeviceToHost));
  d_px.release(); d_sm.release(); d_out.release();
  return 0;
}

}  // extern "C"
