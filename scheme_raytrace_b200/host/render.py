"""The reference-facing render API: trace-all / save-as-ppm (main.scm:439-491) on top of libsrt.so."""
import ctypes as C
import numpy as np
from . import ffi
from .flatten import flatten_scene, FlatScene
from .perlin import perlin_generate

RAY_DTYPE = np.dtype([("o", "<f4", (3,)), ("d", "<f4", (3,)), ("time", "<f4")])
HIT_DTYPE = np.dtype([("prim", "<i4"), ("material", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4"),
                      ("p", "<f4", (3,)), ("n", "<f4", (3,))])
BVH_NODE_DTYPE = np.dtype([("lc", "<f4", (3,)), ("le", "<f4", (3,)), ("rc", "<f4", (3,)), ("re", "<f4", (3,)),
                           ("left", "<i4"), ("right", "<i4"), ("parent", "<i4"), ("sibling", "<i4")])
assert RAY_DTYPE.itemsize == 28 and HIT_DTYPE.itemsize == 44 and BVH_NODE_DTYPE.itemsize == 64


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Renderer:
    """Owns one committed scene on one GPU.  Replaces the in-process closure evaluation of
    trace-all -> color -> g:hit / m:scatter (main.scm:100-121, 471-491)."""

    def __init__(self, scene, device=0, perlin_seed=3, lights=(), perlin=None, gpus=1):
        """gpus > 1 (0 = every visible GPU): the scene is committed on that many GPUs of this process
        (srt_init_multi) and `render_multi` shards a frame's samples over them."""
        self.lib = ffi.load()
        self.gpus = int(gpus)
        if self.gpus != 1:
            ffi.check(self.lib.srt_init_multi(self.gpus), "srt_init_multi")
            self.gpus = self.lib.srt_multi_device_count()
        else:
            ffi.check(self.lib.srt_init(int(device)), "srt_init")
        self.flat = scene if isinstance(scene, FlatScene) else flatten_scene(scene)
        self.perlin = perlin if perlin is not None else perlin_generate(perlin_seed)   # (ranvec, perm-x, perm-y, perm-z), perlin.scm:10-30
        self.lights = np.ascontiguousarray(list(lights), dtype=np.int32)   # primitive ids for the hittable pdf (make-hitable-pdf)
        self.h = self.lib.srt_scene_create()
        if not self.h:
            raise ffi.SrtError("srt_scene_create failed")
        self.commit()

    def commit(self):
        f, lib, h = self.flat, self.lib, self.h
        ffi.check(lib.srt_scene_set_prims(h, _ptr(f.prims), len(f.prims)), "set_prims")
        ffi.check(lib.srt_scene_set_xforms(h, _ptr(f.xforms), len(f.xforms)), "set_xforms")
        ffi.check(lib.srt_scene_set_patches(h, _ptr(f.patches), len(f.patches)), "set_patches")
        ffi.check(lib.srt_scene_set_materials(h, _ptr(f.materials), len(f.materials)), "set_materials")
        ffi.check(lib.srt_scene_set_textures(h, _ptr(f.textures), len(f.textures)), "set_textures")
        ffi.check(lib.srt_scene_set_images(h, _ptr(f.image_texels), _ptr(f.image_dims), len(f.image_dims)), "set_images")
        rv, px, py, pz = self.perlin
        self._rv32 = np.ascontiguousarray(rv, dtype=np.float32)
        ffi.check(lib.srt_scene_set_perlin(h, _ptr(self._rv32), _ptr(px), _ptr(py), _ptr(pz)), "set_perlin")
        ffi.check(lib.srt_scene_set_camera(h, _ptr(f.camera)), "set_camera")
        ffi.check(lib.srt_scene_set_lights(h, _ptr(self.lights), len(self.lights)), "set_lights")
        ffi.check(lib.srt_scene_commit(h), "commit")

    def close(self):
        if getattr(self, "h", None):
            self.lib.srt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- LBVH inspection --------------------------------------------------------------------
    def bvh_nodes(self):
        n = self.lib.srt_bvh_node_count(self.h)
        out = np.zeros(n, dtype=BVH_NODE_DTYPE)
        ffi.check(self.lib.srt_bvh_readback(self.h, _ptr(out), n), "bvh_readback")
        return out

    def bvh_items(self):
        """(item -> primitive id of the LBVH leaves, ids of the huge primitives tested before traversal)"""
        n = int(np.sum((self.flat.prims["flags"] & 2) == 0))
        items, glob, ng = np.zeros(max(n, 1), dtype=np.int32), np.zeros(8, dtype=np.int32), C.c_int32(0)
        k = self.lib.srt_bvh_items_readback(self.h, _ptr(items), n, _ptr(glob), C.byref(ng))
        if k < 0:
            ffi.check(k, "bvh_items_readback")
        return items[:k], glob[:ng.value]

    def bvh_keys(self):
        n = len(self.bvh_items()[0])
        keys, order = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.int32)
        ffi.check(self.lib.srt_bvh_keys_readback(self.h, _ptr(keys), _ptr(order), n), "bvh_keys_readback")
        return keys, order

    def prim_bounds(self):
        n = int(np.sum((self.flat.prims["flags"] & 2) == 0))
        out = np.zeros((n, 6), dtype=np.float32)
        ffi.check(self.lib.srt_prim_bounds_readback(self.h, _ptr(out), n), "prim_bounds_readback")
        return out

    # -- parity hooks -----------------------------------------------------------------------
    def trace_batch(self, rays, t_min=0.001, t_max=999999999999.0):
        """rays: (n,7) float array [o d time] -> HIT_DTYPE array (prim = -1 on miss)."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 7)
        out = np.zeros(len(rays), dtype=HIT_DTYPE)
        ffi.check(self.lib.srt_trace_batch(self.h, _ptr(rays), len(rays), t_min, t_max, _ptr(out)), "trace_batch")
        return out

    def eval_texture(self, tex, uvp, quirks=ffi.QUIRKS_REFERENCE):
        uvp = np.ascontiguousarray(uvp, dtype=np.float32).reshape(-1, 5)
        out = np.zeros((len(uvp), 3), dtype=np.float32)
        ffi.check(self.lib.srt_eval_texture(self.h, tex, _ptr(uvp), len(uvp), quirks, _ptr(out)), "eval_texture")
        return out

    def eval_raygen(self, params, pixel, sample):
        pixel = np.ascontiguousarray(pixel, dtype=np.int32)
        sample = np.ascontiguousarray(sample, dtype=np.int32)
        out = np.zeros((len(pixel), 7), dtype=np.float32)
        ffi.check(self.lib.srt_eval_raygen(self.h, C.byref(params), len(pixel), _ptr(pixel), _ptr(sample), _ptr(out)), "eval_raygen")
        return out

    # -- rendering --------------------------------------------------------------------------
    def params(self, width, height, spp_begin, spp_end, max_depth=50, seed=1, quirks=ffi.QUIRKS_REFERENCE, t_min=0.001, wave_spp=0, sky=None, estimator=0):
        p = ffi.RenderParams()
        p.width, p.height, p.spp_begin, p.spp_end = width, height, spp_begin, spp_end
        p.max_depth, p.seed, p.quirks, p.t_min, p.wave_spp = max_depth, seed, quirks, t_min, wave_spp
        p.sky = self.flat.sky if sky is None else sky
        p.estimator = estimator
        return p

    def render(self, width, height, spp, max_depth=50, seed=1, quirks=ffi.QUIRKS_REFERENCE, spp_begin=0, rgb_sum=None, wave_spp=0, estimator=0):
        """Adds samples [spp_begin, spp_begin+spp) into rgb_sum (H,W,3) float32 (row 0 = bottom).
        Host buffers in and out (D2H inside the call)."""
        if rgb_sum is None:
            rgb_sum = np.zeros((height, width, 3), dtype=np.float32)
        p = self.params(width, height, spp_begin, spp_begin + spp, max_depth, seed, quirks, wave_spp=wave_spp, estimator=estimator)
        st = ffi.Stats()
        ffi.check(self.lib.srt_render_host(self.h, C.byref(p), _ptr(rgb_sum), C.byref(st)), "render_host")
        return rgb_sum, st

    def render_device(self, d_ptr, width, height, spp, max_depth=50, seed=1, quirks=ffi.QUIRKS_REFERENCE, spp_begin=0, wave_spp=0, estimator=0):
        """Same, accumulating into a DEVICE buffer (e.g. a torch tensor's data_ptr())."""
        p = self.params(width, height, spp_begin, spp_begin + spp, max_depth, seed, quirks, wave_spp=wave_spp, estimator=estimator)
        st = ffi.Stats()
        ffi.check(self.lib.srt_render_device(self.h, C.byref(p), C.c_void_p(int(d_ptr)), C.byref(st)), "render_device")
        return st


    def render_multi(self, width, height, spp, max_depth=50, seed=1, quirks=ffi.QUIRKS_REFERENCE, spp_begin=0, rgb_sum=None, estimator=0, want_image=True,
                     image=None, write_only=False, want_sum=True):
        """(trace-all scene k) over every GPU of srt_init_multi from this one process: samples
        [spp_begin, spp_begin+spp) are split into one contiguous range per GPU, the integer accumulators are
        combined with one reduce over NVLink.  Returns (rgb_sum, image8 or None, stats); host buffers.
        `rgb_sum` given: the running sum the frame is added to (write_only=True: an output buffer only, e.g. pinned
        memory reused across frames); `image`: optional preallocated (H, W, 3) uint8 output."""
        fresh = rgb_sum is None
        if fresh and want_sum:
            rgb_sum = np.empty((height, width, 3), dtype=np.float32)
        if want_image and image is None:
            image = np.empty((height, width, 3), dtype=np.uint8)
        if not want_image:
            image = None
        p = self.params(width, height, spp_begin, spp_begin + spp, max_depth, seed, quirks, estimator=estimator)
        p.reserved[2] = 1 if (fresh or write_only) else 0
        st = ffi.Stats()
        ffi.check(self.lib.srt_render_multi(self.h, C.byref(p), _ptr(rgb_sum) if rgb_sum is not None else None, _ptr(image) if want_image else None, C.byref(st)), "render_multi")
        return rgb_sum, image, st

    def progressive_step(self, width, height, spp_begin, spp_end, max_depth=100, seed=1, quirks=ffi.QUIRKS_REFERENCE):
        """One progressive pass with the running sum resident on the device: returns the 8-bit frame only."""
        image = np.zeros((height, width, 3), dtype=np.uint8)
        p = self.params(width, height, spp_begin, spp_end, max_depth, seed, quirks)
        st = ffi.Stats()
        ffi.check(self.lib.srt_progressive_step(self.h, C.byref(p), _ptr(image), C.byref(st)), "progressive_step")
        return image, st

    def progressive_read(self, width, height):
        out = np.zeros((height, width, 3), dtype=np.float32)
        n = C.c_int32(0)
        ffi.check(self.lib.srt_progressive_read(self.h, _ptr(out), C.byref(n)), "progressive_read")
        return out, n.value


class ProgressiveRenderer:
    """The reference's progressive viewer semantics (main.scm:452-469, 533-544: one new sample per
    pixel per pass, running sum, 8-bit image re-derived after every pass) without the GLUT window:
    `step()` == one full-frame pass of `trace-line` over all rows."""

    def __init__(self, scene, width=200, height=200, max_depth=100, seed=1, quirks=ffi.QUIRKS_REFERENCE, device=0):
        self.r = Renderer(scene, device=device)
        self.width, self.height, self.max_depth, self.seed, self.quirks = width, height, max_depth, seed, quirks
        self.sample_count = 0                                              # *sample-count*
        self.image = np.zeros((height, width, 3), dtype=np.uint8)          # *image* main.scm:429

    def step(self, samples=1):
        """One pass: the running sum (*raw-data*, main.scm:430) never leaves the device; only the 8-bit
        frame comes back (srt_progressive_step)."""
        self.image, self.stats = self.r.progressive_step(self.width, self.height, self.sample_count, self.sample_count + samples,
                                                         max_depth=self.max_depth, seed=self.seed, quirks=self.quirks)
        self.sample_count += samples
        return self.image

    @property
    def raw_data(self):
        """*raw-data* read back from the device on demand (inspection / tests)."""
        if self.sample_count == 0:
            return np.zeros((self.height, self.width, 3), dtype=np.float32)
        return self.r.progressive_read(self.width, self.height)[0]

    def save(self, path="test.ppm"):                                       # key 'S' main.scm:551-552
        save_as_ppm(path, self.image)

    def close(self):
        self.r.close()


def correct_gamma_quantise(rgb_sum, spp):
    """main.scm:123-124, 481-487 through the library's device resolve kernel."""
    lib = ffi.load()
    rgb_sum = np.ascontiguousarray(rgb_sum, dtype=np.float32)
    h, w = rgb_sum.shape[:2]
    img = np.zeros((h, w, 3), dtype=np.uint8)
    ffi.check(lib.srt_resolve_host(_ptr(rgb_sum), w, h, int(spp), _ptr(img)), "resolve_host")
    return img


def save_as_ppm(path, image):
    """main.scm:439-450 (ASCII P3, header 'P3\\n W H\\n255\\n', rows top to bottom)."""
    lib = ffi.load()
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w = image.shape[:2]
    ffi.check(lib.srt_save_ppm(str(path).encode(), _ptr(image), w, h), "save_ppm")


def trace_all(scene, sample_count, width=200, height=200, max_depth=100, seed=1, quirks=ffi.QUIRKS_REFERENCE, device=0):
    """(trace-all scene k) for k = 1..sample_count in one call (main.scm:471-491): returns the
    running sum *raw-data* and the 8-bit *image*."""
    r = Renderer(scene, device=device)
    try:
        rgb_sum, st = r.render(width, height, sample_count, max_depth=max_depth, seed=seed, quirks=quirks)
    finally:
        r.close()
    return rgb_sum, correct_gamma_quantise(rgb_sum, sample_count), st
