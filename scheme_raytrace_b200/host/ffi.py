"""ctypes binding of libsrt.so (include/srt.h).  No CPU fallback: a missing library raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SRT_LIB selects an experimental build of the same library for A/B runs (tools/ab.py); the default is the in-tree build
LIB_PATH = os.environ.get("SRT_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "libsrt.so")


# quirk bits of include/srt.h (SURVEY 8a Q rows + Q15); QUIRKS_REFERENCE reproduces upstream HEAD
Q1_COSINE_X2, Q4_PERLIN_ALIAS, Q6_SCATTER_TIME0, Q10_DIELECTRIC_UNNORM, Q15_LOCAL_TRIPLE_EVAL = 1, 2, 4, 8, 16
QUIRKS_REFERENCE = 31


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp_begin", C.c_int32), ("spp_end", C.c_int32),
                ("max_depth", C.c_int32), ("sky", C.c_int32), ("seed", C.c_uint32), ("quirks", C.c_int32),
                ("t_min", C.c_float), ("wave_spp", C.c_int32), ("estimator", C.c_int32), ("reserved", C.c_int32 * 5)]


class Stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("paths", C.c_uint64), ("ms_total", C.c_float), ("ms_commit", C.c_float),
                ("kernel_launches", C.c_int32), ("waves", C.c_int32), ("bvh_nodes", C.c_int32), ("bvh_depth", C.c_int32),
                ("rays_per_bounce", C.c_uint64 * 8), ("ms_extend", C.c_float), ("ms_shade", C.c_float),
                ("extend_launches", C.c_int32), ("tail_runs", C.c_int32), ("nonfinite", C.c_uint64),
                ("pipes", C.c_int32), ("pad", C.c_int32)]


# every symbol include/srt.h declares (tests check the library exports all of them)
SYMBOLS = [
    "srt_device_count", "srt_init", "srt_last_error", "srt_shutdown", "srt_measure_fp32_peak", "srt_scene_create", "srt_scene_destroy",
    "srt_scene_set_prims", "srt_scene_set_xforms", "srt_scene_set_patches", "srt_scene_set_materials", "srt_scene_set_textures",
    "srt_scene_set_images", "srt_scene_set_perlin", "srt_scene_set_camera", "srt_scene_set_lights", "srt_scene_commit", "srt_bvh_node_count", "srt_bvh_readback",
    "srt_bvh_keys_readback", "srt_bvh_items_readback", "srt_prim_bounds_readback", "srt_trace_batch", "srt_render_host", "srt_render_device",
    "srt_resolve_device", "srt_resolve_host", "srt_save_ppm", "srt_eval_texture", "srt_eval_raygen",
    "srt_init_multi", "srt_multi_device_count", "srt_multi_reduce_mode", "srt_render_multi", "srt_progressive_step", "srt_progressive_read",
]

_lib = None


def load():
    """Load libsrt.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32p = C.c_void_p, C.c_int32, C.POINTER(C.c_float)
    lib.srt_device_count.restype = i32
    lib.srt_init.argtypes = [i32]
    lib.srt_last_error.restype = C.c_char_p
    lib.srt_measure_fp32_peak.argtypes = [C.POINTER(C.c_float)]
    lib.srt_scene_create.restype = vp
    lib.srt_scene_destroy.argtypes = [vp]
    for name in ("srt_scene_set_prims", "srt_scene_set_xforms", "srt_scene_set_patches", "srt_scene_set_materials", "srt_scene_set_textures"):
        getattr(lib, name).argtypes = [vp, vp, i32]
    lib.srt_scene_set_patches.argtypes = [vp, vp, i32]
    lib.srt_scene_set_images.argtypes = [vp, vp, vp, i32]
    lib.srt_scene_set_perlin.argtypes = [vp, vp, vp, vp, vp]
    lib.srt_scene_set_camera.argtypes = [vp, vp]
    lib.srt_scene_set_lights.argtypes = [vp, vp, i32]
    lib.srt_scene_commit.argtypes = [vp]
    lib.srt_bvh_node_count.argtypes = [vp]
    lib.srt_bvh_readback.argtypes = [vp, vp, i32]
    lib.srt_bvh_keys_readback.argtypes = [vp, vp, vp, i32]
    lib.srt_bvh_items_readback.argtypes = [vp, vp, i32, vp, vp]
    lib.srt_prim_bounds_readback.argtypes = [vp, vp, i32]
    lib.srt_trace_batch.argtypes = [vp, vp, i32, C.c_float, C.c_float, vp]
    lib.srt_render_host.argtypes = [vp, C.POINTER(RenderParams), vp, C.POINTER(Stats)]
    lib.srt_render_device.argtypes = [vp, C.POINTER(RenderParams), vp, C.POINTER(Stats)]
    if hasattr(lib, "srt_init_multi"):              # (absent only from older experimental builds loaded through SRT_LIB)
        lib.srt_init_multi.argtypes = [i32]
        lib.srt_multi_device_count.restype = i32
        lib.srt_multi_reduce_mode.argtypes = [C.POINTER(i32)]
        lib.srt_render_multi.argtypes = [vp, C.POINTER(RenderParams), vp, vp, C.POINTER(Stats)]
        lib.srt_progressive_step.argtypes = [vp, C.POINTER(RenderParams), vp, C.POINTER(Stats)]
        lib.srt_progressive_read.argtypes = [vp, vp, C.POINTER(i32)]
    lib.srt_resolve_device.argtypes = [vp, i32, i32, i32, vp]
    lib.srt_resolve_host.argtypes = [vp, i32, i32, i32, vp]
    lib.srt_save_ppm.argtypes = [C.c_char_p, vp, i32, i32]
    lib.srt_eval_texture.argtypes = [vp, i32, vp, i32, i32, vp]
    lib.srt_eval_raygen.argtypes = [vp, C.POINTER(RenderParams), i32, vp, vp, vp]
    _lib = lib
    return lib


class SrtError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = load().srt_last_error()
        raise SrtError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
