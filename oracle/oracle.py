"""ctypes wrapper of liborc.so (oracle/srt_oracle.cpp, oracle/lbvh_ref.cpp).

TEST INFRASTRUCTURE ONLY — the product path (scheme_raytrace_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborc.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("srt_oracle.cpp", "lbvh_ref.cpp", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liborc.so"])
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        vp, i32, dbl, u32 = C.c_void_p, C.c_int32, C.c_double, C.c_uint32
        lib.orc_create.restype = vp
        lib.orc_destroy.argtypes = [vp]
        lib.orc_add_texture.argtypes = [vp, i32, dbl, dbl, dbl, dbl, i32, i32]
        lib.orc_add_image.argtypes = [vp, vp, i32, i32]
        lib.orc_add_material.argtypes = [vp, i32, i32, dbl]
        lib.orc_add_node.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32]
        lib.orc_set_root.argtypes = [vp, i32]
        lib.orc_add_patch.argtypes = [vp, vp]
        lib.orc_set_camera.argtypes = [vp, vp]
        lib.orc_set_sky.argtypes = [vp, i32]
        lib.orc_set_perlin.argtypes = [vp, vp, vp, vp, vp]
        lib.orc_set_exclude_leaf.argtypes = [vp, i32]
        lib.orc_make_camera.argtypes = [vp, vp, vp, dbl, dbl, dbl, dbl, dbl, dbl, vp]
        lib.orc_trace_batch.argtypes = [vp, i32, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp]
        lib.orc_bezier_hit.argtypes = [vp, dbl, vp, dbl, dbl, vp, vp, vp, vp, vp]
        lib.orc_aabb_hit.argtypes = [vp, vp, vp, dbl, dbl]
        lib.orc_philox4x32_10.argtypes = [vp, vp, vp]
        lib.orc_rng_block.argtypes = [u32, u32, u32, u32, u32, vp]
        lib.orc_tex_value.argtypes = [vp, i32, i32, vp, i32, vp]
        lib.orc_noise.argtypes = [vp, i32, vp, i32, i32, vp]
        lib.orc_get_ray.argtypes = [vp, dbl, dbl, dbl, u32, u32, u32, vp]
        lib.orc_sky.argtypes = [vp, vp, vp]
        lib.orc_onb_cosine.argtypes = [vp, dbl, dbl, i32, vp]
        lib.orc_render.argtypes = [vp, i32, i32, i32, i32, i32, u32, i32, i32, i32, vp, vp, i32]
        lib.orc_set_lights.argtypes = [vp, vp, i32]
        lib.orc_resolve.argtypes = [vp, i32, i32, i32, vp]
        lib.orc_save_ppm.argtypes = [C.c_char_p, vp, i32, i32]
        lib.orc_lbvh_build.argtypes = [i32, vp, vp, vp, vp]
        lib.orc_lbvh_depth.argtypes = [i32, vp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d(seq):
    return np.ascontiguousarray(seq, dtype=np.float64)


class OracleScene:
    """Builds the oracle's node tree from the host object tree (scheme_raytrace_b200.host.geometry),
    mirroring the reference's closure nesting one to one.  Every scene parameter is first rounded
    to fp32 — the precision of the C-ABI tables — so the f64 oracle and the GPU evaluate the SAME
    scene (like the fixed ray batches, which are generated in fp32 and widened)."""

    def __init__(self, scene, perlin=None, flat=None, quantise=True):
        from scheme_raytrace_b200.host import geometry as g
        from scheme_raytrace_b200.host import texture as t
        from scheme_raytrace_b200.host.camera import camera_to_floats
        from scheme_raytrace_b200.host.flatten import flatten_scene, sky_kind
        from scheme_raytrace_b200.host.perlin import perlin_generate
        self.lib = load()
        self._q = (lambda a: np.asarray(a, dtype=np.float32)) if quantise else (lambda a: np.asarray(a, dtype=np.float64))
        self.h = self.lib.orc_create()
        self.flat = flat if flat is not None else flatten_scene(scene)
        self._tex, self._mat = {}, {}
        self._leaf = 0
        # materials/textures are added in the flat scene's table order so ids agree with the GPU's
        for tx in self.flat.texture_objs:
            self._add_tex(tx, t)
        for m in self.flat.material_objs:
            self._add_mat(m)
        root_children = [self._add(o, g) for o in scene.obj_list]
        root = self._node(g.LIST, -1, -1, (), root_children)
        self.lib.orc_set_root(self.h, root)
        self.n_leaves = self._leaf
        if scene.camera is not None:
            cam = _d(self._q(camera_to_floats(scene.camera)))
            self.lib.orc_set_camera(self.h, _p(cam))
        self.lib.orc_set_sky(self.h, sky_kind(scene.sky_function))
        rv, px, py, pz = perlin if perlin is not None else perlin_generate(3)
        rv = _d(self._q(rv))
        self.lib.orc_set_perlin(self.h, _p(rv), _p(px), _p(py), _p(pz))

    def _add_tex(self, tx, t):
        if id(tx) in self._tex:
            return self._tex[id(tx)]
        even = odd = -1
        if tx.kind == t.CHECKER:
            even, odd = self._add_tex(tx.even, t), self._add_tex(tx.odd, t)
        if tx.kind == t.IMAGE:
            tex = np.ascontiguousarray(tx.image, dtype=np.uint8)
            even = self.lib.orc_add_image(self.h, _p(tex), tex.shape[1], tex.shape[0])
        r, g_, b, sc = (float(x) for x in self._q([tx.rgb[0], tx.rgb[1], tx.rgb[2], tx.scale]))
        i = self.lib.orc_add_texture(self.h, tx.kind, r, g_, b, sc, even, odd)
        self._tex[id(tx)] = i
        return i

    def _add_mat(self, m):
        if m is None:
            return -1
        if id(m) not in self._mat:
            tx = self._tex[id(m.tex)] if m.tex is not None else -1
            self._mat[id(m)] = self.lib.orc_add_material(self.h, m.kind, tx, float(self._q([m.param])[0]))
        return self._mat[id(m)]

    def _node(self, kind, material, leaf, params, children):
        prm = _d(self._q(list(params) if len(params) else [0.0]))
        ch = np.ascontiguousarray(children if len(children) else [0], dtype=np.int32)
        return self.lib.orc_add_node(self.h, kind, material, leaf, _p(prm), len(params), _p(ch), len(children))

    def _add(self, obj, g, boundary=False):
        leaf = -1
        if obj.kind in g.LEAF_KINDS and not boundary:      # boundary shapes of a medium are not scene primitives
            leaf = self._leaf
            self._leaf += 1
        children = [self._add(c, g, boundary or obj.kind == g.CONSTANT_MEDIUM) for c in obj.children]
        mat = self._add_mat(obj.material)
        params = obj.params
        if obj.kind == g.PATCH:                            # control net goes to the patch table
            cp = _d(self._q(obj.params))
            params = (float(self.lib.orc_add_patch(self.h, _p(cp))),)
            return self._node_raw(obj.kind, mat, leaf, params, children)
        return self._node(obj.kind, mat, leaf, params, children)

    def _node_raw(self, kind, material, leaf, params, children):
        prm = _d(list(params))
        ch = np.ascontiguousarray(children if len(children) else [0], dtype=np.int32)
        return self.lib.orc_add_node(self.h, kind, material, leaf, _p(prm), len(params), _p(ch), len(children))

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trace_batch(self, rays, t_min=0.001, t_max=999999999999.0, precision=64, exclude_leaf=None):
        rays = _d(rays).reshape(-1, 7)
        n = len(rays)
        leaf, mat = np.zeros(n, np.int32), np.zeros(n, np.int32)
        t, p, nrm, uv = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 2))
        self.lib.orc_set_exclude_leaf(self.h, -1 if exclude_leaf is None else int(exclude_leaf))
        self.lib.orc_trace_batch(self.h, n, _p(rays), t_min, t_max, precision, _p(leaf), _p(t), _p(p), _p(nrm), _p(uv), _p(mat))
        self.lib.orc_set_exclude_leaf(self.h, -1)
        return dict(prim=leaf, t=t, p=p, n=nrm, uv=uv, material=mat)

    def second_best_t(self, rays, best_leaf, t_min=0.001, t_max=999999999999.0):
        """t of the closest hit with the best leaf excluded (near-tie filter); inf where none."""
        rays = _d(rays).reshape(-1, 7)
        out = np.full(len(rays), np.inf)
        for lf in np.unique(best_leaf):
            if lf < 0:
                continue
            idx = np.nonzero(best_leaf == lf)[0]
            r = self.trace_batch(rays[idx], t_min, t_max, 64, exclude_leaf=int(lf))
            out[idx] = np.where(r["prim"] >= 0, r["t"], np.inf)
        return out

    def set_lights(self, prim_ids):
        ids = np.ascontiguousarray(prim_ids, dtype=np.int32)
        rc = self.lib.orc_set_lights(self.h, _p(ids), len(ids))
        if rc != 0:
            raise ValueError(f"orc_set_lights failed ({rc})")

    def render(self, width, height, spp, max_depth=50, seed=1, quirks=31, spp_begin=0, nthreads=0, precision=64, rgb_sum=None, estimator=0):
        if rgb_sum is None:
            rgb_sum = np.zeros((height, width, 3), dtype=np.float64)
        if nthreads <= 0:
            nthreads = os.cpu_count() or 1
        nrays = C.c_uint64(0)
        self.lib.orc_render(self.h, width, height, spp_begin, spp_begin + spp, max_depth, seed, quirks, nthreads, precision,
                            _p(rgb_sum), C.byref(nrays), estimator)
        return rgb_sum, int(nrays.value)

    def tex_value(self, tex, uvp, quirks=31):
        uvp = _d(uvp).reshape(-1, 5)
        out = np.zeros((len(uvp), 3))
        self.lib.orc_tex_value(self.h, tex, len(uvp), _p(uvp), quirks, _p(out))
        return out

    def noise(self, p, quirks=31, turb=False):
        p = _d(p).reshape(-1, 3)
        out = np.zeros(len(p))
        self.lib.orc_noise(self.h, len(p), _p(p), quirks, int(turb), _p(out))
        return out

    def get_ray(self, s, t, xi_time, seed, pixel, sample):
        out = np.zeros(7)
        self.lib.orc_get_ray(self.h, s, t, xi_time, seed, pixel, sample, _p(out))
        return out

    def sky(self, d):
        d, out = _d(d), np.zeros(3)
        self.lib.orc_sky(self.h, _p(d), _p(out))
        return out


def make_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, time0, time1):
    out = np.zeros(24)
    a, b, c = _d(lookfrom), _d(lookat), _d(vup)
    load().orc_make_camera(_p(a), _p(b), _p(c), vfov, aspect, aperture, focus_dist, time0, time1, _p(out))
    return out


def bezier_hit(cps, width, ray7, t_min=0.001, t_max=999999999999.0):
    cps, ray7 = _d(cps).reshape(12), _d(ray7)
    t, p, n = C.c_double(0), np.zeros(3), np.zeros(3)
    md, calls = C.c_int32(0), C.c_int32(0)
    hit = load().orc_bezier_hit(_p(cps), width, _p(ray7), t_min, t_max, C.byref(t), _p(p), _p(n), C.byref(md), C.byref(calls))
    return dict(hit=bool(hit), t=t.value, p=p, n=n, max_depth=md.value, converge_calls=calls.value)


def reflect(v, n):                             # material.scm:41-43
    a, b, o = _d(v), _d(n), np.zeros(3)
    load().orc_reflect(_p(a), _p(b), _p(o))
    return o


def refract(v, n, ni_over_nt, quirks=31):      # material.scm:59-67 -> (ok, refracted)
    a, b, o = _d(v), _d(n), np.zeros(3)
    lib = load()
    lib.orc_refract.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p]
    ok = lib.orc_refract(_p(a), _p(b), float(ni_over_nt), int(quirks), _p(o))
    return bool(ok), o


def schlick(cosine, ref_idx):                  # material.scm:69-74
    lib = load()
    lib.orc_schlick.argtypes = [C.c_double, C.c_double]
    lib.orc_schlick.restype = C.c_double
    return lib.orc_schlick(float(cosine), float(ref_idx))


def aabb_hit(bmin, bmax, ray7, t_min, t_max):
    a, b, r = _d(bmin), _d(bmax), _d(ray7)
    return bool(load().orc_aabb_hit(_p(a), _p(b), _p(r), t_min, t_max))


def philox4x32_10(ctr, key):
    c, k, o = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    load().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def rng_block(seed, pixel, sample, bounce, block):
    o = np.zeros(4)
    load().orc_rng_block(seed, pixel, sample, bounce, block, _p(o))
    return o


def resolve(rgb_sum, spp):
    rgb_sum = _d(rgb_sum)
    h, w = rgb_sum.shape[:2]
    img = np.zeros((h, w, 3), np.uint8)
    load().orc_resolve(_p(rgb_sum), w, h, spp, _p(img))
    return img


def save_ppm(path, image):
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape[:2]
    return load().orc_save_ppm(str(path).encode(), _p(image), w, h)


NODE64 = np.dtype([("lc", "<f4", (3,)), ("le", "<f4", (3,)), ("rc", "<f4", (3,)), ("re", "<f4", (3,)),
                   ("left", "<i4"), ("right", "<i4"), ("parent", "<i4"), ("sibling", "<i4")])


def lbvh_build(aabbs):
    aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
    n = len(aabbs)
    keys, order = np.zeros(max(n, 1), np.uint64), np.zeros(max(n, 1), np.int32)
    nodes = np.zeros(max(n - 1, 1), dtype=NODE64)
    nn = load().orc_lbvh_build(n, _p(aabbs), _p(keys), _p(order), _p(nodes))
    return keys[:n], order[:n], nodes[:nn]


def lbvh_depth(nodes):
    nodes = np.ascontiguousarray(nodes)
    return load().orc_lbvh_depth(len(nodes), _p(nodes))
