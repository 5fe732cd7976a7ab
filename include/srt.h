/* srt.h — C ABI of libsrt.so: the B200-native per-sample radiance loop of scheme-raytrace.
 *
 * This is the drop-in boundary.  The reference has no FFI layer; its de-facto
 * operator interface is the closure-vector protocol evaluated by trace-all / color
 * (main.scm:100-121, 471-491).  The host (Scheme scene scripts, or the Python mirror in
 * scheme_raytrace_b200/host) keeps the reference's constructor API and flattens the scene into
 * the POD tables below; everything behind these entry points is hand-written sm_100a CUDA.
 *
 * Conventions: every function returns 0 on success and a negative SrtError otherwise; nothing
 * throws or aborts across the ABI; srt_last_error() gives a thread-local message.  There is NO
 * CPU fallback: without an sm_100 device srt_init() fails.  The caller owns all input and output
 * arrays (inputs are copied during set_ / commit); the library owns device memory behind the
 * opaque handle.  A handle is not thread-safe.  Calls are synchronous on return unless the name
 * ends in _async.
 */
#ifndef SRT_H
#define SRT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SRT_OK = 0, SRT_ERR_NO_DEVICE = -1, SRT_ERR_CUDA = -2, SRT_ERR_ARG = -3, SRT_ERR_NOT_COMMITTED = -4,
  SRT_ERR_BVH_DEPTH = -5, SRT_ERR_IO = -6
} SrtError;

/* primitive kinds — replace the hit closures of geometry.scm / bezier.scm */
enum {
  SRT_PRIM_SPHERE = 0,        /* geometry.scm:146 make-sphere         p = cx cy cz r                       */
  SRT_PRIM_MOVING_SPHERE = 1, /* geometry.scm:177 make-moving-sphere  p = c0.xyz r c1.xyz time0 time1      */
  SRT_PRIM_XY_RECT = 2,       /* geometry.scm:376 make-xy-rect        p = x0 x1 y0 y1 k                    */
  SRT_PRIM_XZ_RECT = 3,       /* geometry.scm:395 make-xz-rect        p = x0 x1 z0 z1 k                    */
  SRT_PRIM_YZ_RECT = 4,       /* geometry.scm:414 make-yz-rect        p = y0 y1 z0 z1 k                    */
  SRT_PRIM_BEZIER = 5,        /* bezier.scm:61   make-bezier          p = a.xyz b.xyz c.xyz d.xyz width    */
  SRT_PRIM_CONSTANT_MEDIUM = 6, /* geometry.scm:545 make-constant-medium p = density, first boundary prim, #boundary prims */
  SRT_PRIM_KLEIN = 8,         /* geometry.scm:644 make-klein (sphere-traced IIS fractal, no bounding box: always
                               * tested before traversal, at most 8 per scene)   p = centre.xyz */
  SRT_PRIM_PATCH = 7          /* bicubic Bezier (sub-)patch (north-star extension, absent upstream):
                               * p = index into the patch table, u0, v0, size of this sub-patch in its parent's
                               * (u,v) domain; the host pre-splits each patch 2 levels into 16 of these */
};
#define SRT_PRIM_FLAG_FLIP 1  /* geometry.scm:433 flip-normals (parity of the flips above the leaf)        */
#define SRT_PRIM_FLAG_BOUNDARY 2 /* boundary shape of a constant medium: not a scene surface, not in the LBVH;
                                  * boundary primitives must form a suffix of the primitive array */

/* One flattened leaf primitive.  Position in the array = primitive id = position in the
 * reference's flattened top-level object list (box faces in make-box order, geometry.scm:446-457);
 * the id drives the exact-tie rule. */
typedef struct {
  int32_t type;      /* SRT_PRIM_*            */
  int32_t flags;     /* SRT_PRIM_FLAG_*       */
  int32_t material;  /* index into materials  */
  int32_t xform;     /* index into xforms, -1 = none */
  float p[16];       /* p[15]: logical primitive id + 1 reported by srt_trace_batch (0 = array position);
                      * the 16 sub-patches of one patch share their parent's logical id */
} SrtPrim;

/* Rigid instance transform = the composition of a translate / rotate-y chain
 * (geometry.scm:465 translate, :483 rotate-y):  world = Ry * object + offset with
 * Ry*(x,y,z) = (cos*x + sin*z, y, -sin*x + cos*z)  (geometry.scm:526-530). */
typedef struct { float sin_t, cos_t, off[3]; } SrtXform;

enum { SRT_MAT_LAMBERTIAN = 0, SRT_MAT_METAL = 1, SRT_MAT_DIELECTRIC = 2, SRT_MAT_DIFFUSE_LIGHT = 3, SRT_MAT_ISOTROPIC = 4 };
/* material.scm:24 make-lambertian, :45 make-metal (param = fuzz), :76 make-dielectric
 * (param = ref-idx), :103 make-diffuse-light; isotropic is the book's (absent upstream). */
typedef struct { int32_t kind; int32_t tex; float param; float pad; } SrtMaterial;

enum { SRT_TEX_CONSTANT = 0, SRT_TEX_CHECKER = 1, SRT_TEX_NOISE = 2, SRT_TEX_MARBLE = 3, SRT_TEX_IMAGE = 4 };
/* texture.scm:12 constant-texture (rgb), :16 checker-texture (even, odd = texture indices),
 * :25 noise-texture (scale), :30 marble-texture (scale), :36 image-texture (even = image index,
 * see srt_scene_set_images). */
typedef struct { int32_t kind; int32_t even, odd; float scale; float rgb[3]; float pad; } SrtTexture;

/* the 10 slots of camera.scm:70-78 */
typedef struct { float llc[3], horiz[3], vert[3], origin[3], w[3], u[3], v[3]; float lens_radius, time0, time1; } SrtCamera;

enum { SRT_SKY_GRADIENT = 0 /* main.scm:91 sky-color */, SRT_SKY_BLACK = 1 /* main.scm:97 black */ };

/* quirk bits.  SRT_QUIRKS_REFERENCE reproduces upstream HEAD. */
#define SRT_Q1_COSINE_X2 1          /* util.scm:42-43     */
#define SRT_Q4_PERLIN_ALIAS 2       /* perlin.scm:76      */
#define SRT_Q6_SCATTER_TIME0 4      /* ray.scm:8-9        */
#define SRT_Q10_DIELECTRIC_UNNORM 8 /* material.scm:59-67 */
#define SRT_Q15_LOCAL_TRIPLE_EVAL 16 /* onb.scm:27-36: the `local` macro mentions its operand three times, so
                                        (local uvw (random-cosine-direction)) (material.scm:27, pdf.scm:26) draws three
                                        directions and takes x, y, z from the 1st, 2nd, 3rd */
#define SRT_QUIRKS_REFERENCE 31

/* radiance estimator: the reference's `color` (main.scm:100-121, cosine sampling only), or the
 * Rest-of-Life mixture(hittable(lights), cosine) pdf (pdf.scm:18-41; hittable part unpinned) */
enum { SRT_EST_REFERENCE = 0, SRT_EST_MIXTURE = 1 };

typedef struct {
  int32_t width, height;       /* main.scm:126-127 *size-x* *size-y*                     */
  int32_t spp_begin, spp_end;  /* sample range [begin,end) rendered by this call         */
  int32_t max_depth;           /* main.scm:26 +max-depth+                                */
  int32_t sky;                 /* SRT_SKY_*                                              */
  uint32_t seed;               /* Philox key word 1                                      */
  int32_t quirks;              /* SRT_Q* bits                                            */
  float t_min;                 /* main.scm:104: 0.001                                    */
  int32_t wave_spp;            /* path-queue capacity in samples/pixel; 0 = auto (64 Mi paths) */
  int32_t estimator;           /* SRT_EST_*                                              */
  int32_t reserved[5];         /* [0] = 1: time extend/shade launches separately and count rays per bounce (slower);
                                * [1] = 1: no CUDA graph; [2] = 1: rgb_sum of srt_render_host / _multi is write-only
                                * (the frame starts from zero: skips the upload of the running sum);
                                * [3] = 1: drain through the wavefront instead of the one-launch tail kernel (A/B, tests);
                                * [4] = 2: split the sample range over two concurrent streaming pipelines (A/B, tests; default one) */
} SrtRenderParams;

typedef struct {
  uint64_t rays;               /* closest-hit queries (primary + every bounce)           */
  uint64_t paths;              /* camera samples                                         */
  float ms_total;              /* device time, first ray-gen launch .. accumulation done */
  float ms_commit;             /* last commit: host wall clock of srt_scene_commit (staging + H2D + LBVH build) */
  int32_t kernel_launches;     /* kernels launched by this call                          */
  int32_t waves;               /* extend/shade/regen iterations of the streaming wavefront */
  int32_t bvh_nodes, bvh_depth;
  uint64_t rays_per_bounce[8]; /* closest-hit queries by bounce 0..6 and >= 7; only when params.reserved[0]==1 */
  float ms_extend, ms_shade;   /* per-kernel device time, only when params.reserved[0]==1 */
  int32_t extend_launches;     /* extend launches timed for ms_extend                    */
  int32_t tail_runs;           /* times the one-launch drain kernel finished the queue   */
  uint64_t nonfinite;          /* NaN / Inf radiance contributions dropped (0 inside the reference's domain) */
  int32_t pipes;               /* concurrent streaming pipelines the call used on its GPU (1 or 2) */
  int32_t pad;
} SrtStats;

/* 64-byte node of the LBVH as the traversal kernel reads it: the two child boxes as centre and
 * padded half extent; child < 0 is leaf ~child. */
typedef struct { float lc[3], le[3], rc[3], re[3]; int32_t left, right, parent, sibling; } SrtBvhNode;

/* ray / hit records of the parity hook (replaces (g:hit scene r t-min t-max), geometry.scm:14) */
typedef struct { float o[3], d[3], time; } SrtRay;
typedef struct { int32_t prim; int32_t material; float t, u, v; float p[3], n[3]; } SrtHit;

typedef struct SrtScene SrtScene;

int srt_device_count(void);
int srt_init(int device);                         /* selects the device new scenes are created on; fails without sm_100.
                                                   * Every scene remembers its device; several scenes on several GPUs may coexist. */
/* One process, n GPUs (the caller this path replaces, (trace-all scene k) main.scm:471-491, is one
 * process): devices 0..n-1 (0 = all visible), peer access from the first device, one NCCL communicator
 * per device (ncclCommInitAll; libnccl is dlopen'ed - absent NCCL, or SRT_MULTI_REDUCE=p2p, selects the
 * library's own peer-read reduce kernel).  Scenes created afterwards (until the next plain srt_init) live
 * on the first device; srt_scene_commit() commits one replica of them per GPU and srt_render_multi() uses them all. */
int srt_init_multi(int n_gpus);
int srt_multi_device_count(void);
int srt_multi_reduce_mode(int32_t* nccl_version); /* 0 = ncclReduce, 1 = peer-read kernel */
const char* srt_last_error(void);
void srt_shutdown(void);
int srt_measure_fp32_peak(float* tflops);         /* FFMA microbenchmark: the FP32 roofline denominator */

SrtScene* srt_scene_create(void);
void srt_scene_destroy(SrtScene*);
int srt_scene_set_prims(SrtScene*, const SrtPrim*, int n);
int srt_scene_set_xforms(SrtScene*, const SrtXform*, int n);
int srt_scene_set_patches(SrtScene*, const float* cp48, int n);      /* n x 16 control points (xyz), P[i][j] at 3*(4i+j) */
int srt_scene_set_materials(SrtScene*, const SrtMaterial*, int n);
int srt_scene_set_textures(SrtScene*, const SrtTexture*, int n);
/* image-texture data (texture.scm:36-50): n images of 8-bit RGB texels, row-major from the top row,
 * concatenated in `texels`; dims = n x (nx, ny, byte offset of the image in `texels`). */
int srt_scene_set_images(SrtScene*, const uint8_t* texels, const int32_t* dims, int n);
int srt_scene_set_perlin(SrtScene*, const float* ranvec768, const int32_t* perm_x, const int32_t* perm_y, const int32_t* perm_z);
int srt_scene_set_camera(SrtScene*, const SrtCamera*);
int srt_scene_set_lights(SrtScene*, const int32_t* prim_ids, int n);   /* shapes sampled by the hittable pdf */
int srt_scene_commit(SrtScene*);                  /* H2D + GPU LBVH build                      */

/* LBVH inspection (bit-exact check against the host reference build) */
int srt_bvh_node_count(SrtScene*);
int srt_bvh_readback(SrtScene*, SrtBvhNode* nodes, int cap);
int srt_bvh_keys_readback(SrtScene*, uint64_t* keys_sorted, int32_t* order, int cap);   /* over the LBVH items */
/* LBVH item -> primitive id, and the (<= 8) huge primitives kept out of the tree; returns #items */
int srt_bvh_items_readback(SrtScene*, int32_t* item_prim, int cap, int32_t* global_prims8, int32_t* n_global);
int srt_prim_bounds_readback(SrtScene*, float* aabbs6, int cap);

/* fixed-ray-batch closest hit through the SAME extend kernel the renderer uses */
int srt_trace_batch(SrtScene*, const SrtRay* rays, int n, float t_min, float t_max, SrtHit* out);

/* Render samples [spp_begin, spp_end) and ADD them to rgb_sum (W*H*3 floats, y = 0 bottom row,
 * main.scm:471-491 trace-all).  _host: rgb_sum in host memory (D2H inside); _device: rgb_sum is
 * a device pointer on the scene's device (multi-GPU reduce is done by the caller on it). */
int srt_render_host(SrtScene*, const SrtRenderParams*, float* rgb_sum, SrtStats* stats);
int srt_render_device(SrtScene*, const SrtRenderParams*, float* d_rgb_sum, SrtStats* stats);

/* The same over the n GPUs of srt_init_multi from this one process: the sample range is split into n
 * contiguous ranges, the per-GPU 64-bit fixed-point accumulators are combined on the first GPU with ONE
 * reduce over NVLink (exact integer sum: the frame is bit-identical to the single-GPU frame), then
 * rgb_sum += frame and image = gamma + 8-bit of rgb_sum / spp_end.  rgb_sum (W*H*3 floats) and image
 * (W*H*3 bytes) are HOST buffers, either may be NULL.  params.reserved[2] == 1: rgb_sum is write-only. */
int srt_render_multi(SrtScene*, const SrtRenderParams*, float* rgb_sum, uint8_t* image, SrtStats* stats);

/* Progressive passes (main.scm:452-469 trace-line, :493-503 display): the running sum *raw-data* stays
 * RESIDENT on the device between calls.  One call adds samples [spp_begin, spp_end) and returns only the
 * 8-bit frame (host, W*H*3) of the spp_end samples so far; spp_begin == 0 starts a new accumulation,
 * otherwise spp_begin must equal the previous call's spp_end.  _read copies the running sum out. */
int srt_progressive_step(SrtScene*, const SrtRenderParams*, uint8_t* image, SrtStats* stats);
int srt_progressive_read(SrtScene*, float* rgb_sum, int32_t* spp);

/* main.scm:123-124,481-487 correct-gamma + quantise; :439-450 save-as-ppm */
int srt_resolve_device(const float* d_rgb_sum, int width, int height, int spp, uint8_t* d_image);
int srt_resolve_host(const float* rgb_sum, int width, int height, int spp, uint8_t* image);
int srt_save_ppm(const char* path, const uint8_t* image, int width, int height);

/* unit hooks used by the parity tests (each runs the device function the shade kernel uses) */
int srt_eval_texture(SrtScene*, int tex, const float* uvp5, int n, int quirks, float* rgb);
int srt_eval_raygen(SrtScene*, const SrtRenderParams*, int n, const int32_t* pixel, const int32_t* sample, SrtRay* out);

#ifdef __cplusplus
}
#endif
#endif
