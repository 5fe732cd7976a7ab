"""Scene flattening: object tree (geometry.py) -> the POD tables of include/srt.h.

Primitive id = position in the depth-first, list-order walk of the reference's object list
(box faces in make-box order) — the order `hit-obj-list` (geometry.scm:33-50) visits them, which
is what the exact-tie rule is defined on.
"""
import math
from dataclasses import dataclass
import numpy as np
from . import geometry as g
from . import texture as t
from .camera import camera_to_floats

PRIM_DTYPE = np.dtype([("type", "<i4"), ("flags", "<i4"), ("material", "<i4"), ("xform", "<i4"), ("p", "<f4", (16,))])
XFORM_DTYPE = np.dtype([("sin_t", "<f4"), ("cos_t", "<f4"), ("off", "<f4", (3,))])
MATERIAL_DTYPE = np.dtype([("kind", "<i4"), ("tex", "<i4"), ("param", "<f4"), ("pad", "<f4")])
TEXTURE_DTYPE = np.dtype([("kind", "<i4"), ("even", "<i4"), ("odd", "<i4"), ("scale", "<f4"), ("rgb", "<f4", (3,)), ("pad", "<f4")])
CAMERA_DTYPE = np.dtype([("llc", "<f4", (3,)), ("horiz", "<f4", (3,)), ("vert", "<f4", (3,)), ("origin", "<f4", (3,)),
                         ("w", "<f4", (3,)), ("u", "<f4", (3,)), ("v", "<f4", (3,)),
                         ("lens_radius", "<f4"), ("time0", "<f4"), ("time1", "<f4")])
assert PRIM_DTYPE.itemsize == 80 and XFORM_DTYPE.itemsize == 20 and MATERIAL_DTYPE.itemsize == 16
assert TEXTURE_DTYPE.itemsize == 32 and CAMERA_DTYPE.itemsize == 96

SKY_GRADIENT, SKY_BLACK = 0, 1
PATCH_HOST_LEVELS = 2          # each bicubic patch becomes 4**2 = 16 sub-patch primitives (separate LBVH leaves)


def _split_cubic(c):
    ab, bc, cd = 0.5 * (c[0] + c[1]), 0.5 * (c[1] + c[2]), 0.5 * (c[2] + c[3])
    abc, bcd = 0.5 * (ab + bc), 0.5 * (bc + cd)
    m_ = 0.5 * (abc + bcd)
    return np.stack([c[0], ab, abc, m_]), np.stack([m_, bcd, cd, c[3]])


def split_patch(net, u0=0.0, v0=0.0, size=1.0, levels=PATCH_HOST_LEVELS):
    """de Casteljau split at 0.5 in u then v, `levels` times: [(net(4,4,3), u0, v0, size), ...]."""
    if levels == 0:
        return [(net, u0, v0, size)]
    lo, hi = np.zeros_like(net), np.zeros_like(net)
    for j in range(4):
        lo[:, j], hi[:, j] = _split_cubic(net[:, j])
    out, h = [], size * 0.5
    for half, (du, sub) in enumerate(((0.0, lo), (h, hi))):
        a, b = np.zeros_like(net), np.zeros_like(net)
        for i in range(4):
            a[i], b[i] = _split_cubic(sub[i])
        out += split_patch(a, u0 + du, v0, h, levels - 1) + split_patch(b, u0 + du, v0 + h, h, levels - 1)
    return out


@dataclass
class FlatScene:
    prims: np.ndarray
    patches: np.ndarray     # (n_sub_patches, 48) float32 control nets of the pre-split bicubic patches
    logical: np.ndarray     # array position -> logical primitive id (== oracle leaf id)
    first_of_logical: np.ndarray   # logical id -> first array position
    xforms: np.ndarray
    materials: np.ndarray
    textures: np.ndarray
    camera: np.ndarray
    sky: int
    leaves: list            # leaf Obj per primitive id (host bookkeeping)
    material_objs: list
    texture_objs: list
    image_texels: np.ndarray = None   # uint8, every image-texture's texels concatenated
    image_dims: np.ndarray = None     # (n_images, 3) int32: nx, ny, byte offset

    def h2d_bytes(self):
        return int(self.prims.nbytes + self.patches.nbytes + self.xforms.nbytes + self.materials.nbytes + self.textures.nbytes + self.camera.nbytes + self.image_texels.nbytes)


def _compose(chain):
    """Compose a translate / rotate-y chain (outermost first) into world = Ry*obj + off, in f64.
    rotate-y object->world: (c*x + s*z, y, -s*x + c*z)  (geometry.scm:526-530)."""
    c, s, off = 1.0, 0.0, (0.0, 0.0, 0.0)
    for kind, prm in chain:        # A = current (outer), B = next (inner):  A∘B
        if kind == g.TRANSLATE:
            bc, bs, boff = 1.0, 0.0, prm
        else:
            bc, bs, boff = prm[1], prm[0], (0.0, 0.0, 0.0)
        roff = (c * boff[0] + s * boff[2], boff[1], -s * boff[0] + c * boff[2])
        off = (off[0] + roff[0], off[1] + roff[1], off[2] + roff[2])
        c, s = c * bc - s * bs, s * bc + c * bs
    return s, c, off


def sky_kind(sky_function):
    """Arbitrary sky lambdas cannot cross the FFI; the two shipped functions map to an enum."""
    if sky_function is None:
        return SKY_BLACK
    if isinstance(sky_function, int):
        return sky_function
    name = getattr(sky_function, "__name__", "")
    if name in ("sky_color", "sky-color"):
        return SKY_GRADIENT
    if name == "black":
        return SKY_BLACK
    raise ValueError("sky function must be scenes.sky_color or scenes.black (enum across the C-ABI)")


def flatten_scene(scene):
    prims, xforms, leaves = [], [], []
    patches = []
    boundary = []          # boundary primitives of constant media: appended after the surface primitives
    xform_ids = {}
    mats, mat_ids, texs, tex_ids = [], {}, [], {}
    images = []            # (nx, ny, byte offset, texels) per image-texture

    def tex_id(tx):
        if id(tx) in tex_ids:
            return tex_ids[id(tx)]
        even = odd = -1
        if tx.kind == t.CHECKER:
            even, odd = tex_id(tx.even), tex_id(tx.odd)
        if tx.kind == t.IMAGE:
            even = len(images)
            off = sum(im[3].size for im in images)
            images.append((tx.image.shape[1], tx.image.shape[0], off, tx.image.ravel()))
        tex_ids[id(tx)] = len(texs)
        texs.append((tx, even, odd))
        return tex_ids[id(tx)]

    def mat_id(m):
        if m is None:
            raise ValueError("primitive without material")
        if id(m) not in mat_ids:
            mat_ids[id(m)] = len(mats)
            mats.append((m, tex_id(m.tex) if m.tex is not None else -1))
        return mat_ids[id(m)]

    def walk(obj, chain, flip, out=None, out_leaves=None):
        out = prims if out is None else out
        out_leaves = leaves if out_leaves is None else out_leaves
        if obj.kind == g.LIST:
            for c in obj.children:
                walk(c, chain, flip, out, out_leaves)
        elif obj.kind == g.CONSTANT_MEDIUM:
            if out is not prims:
                raise ValueError("a constant medium cannot bound another constant medium")
            mine = []
            walk(obj.children[0], chain, 0, mine, [])          # boundary shapes (geometry.scm:549-553)
            for b in mine:
                if b[0] not in (g.SPHERE, g.XY_RECT, g.XZ_RECT, g.YZ_RECT):
                    raise ValueError("constant-medium boundaries must be spheres / rects / boxes")
            prims.append((g.CONSTANT_MEDIUM, 0, mat_id(obj.material), -1, [obj.params[0], ("boundary", len(boundary)), len(mine)], len(leaves)))
            boundary.extend((k, fl | 2, mat_id(obj.material), xf, prm, -1) for (k, fl, _m, xf, prm, _l) in mine)
            leaves.append(obj)
        elif obj.kind == g.FLIP:
            walk(obj.children[0], chain, flip ^ 1, out, out_leaves)
        elif obj.kind in (g.TRANSLATE, g.ROTATE_Y):
            walk(obj.children[0], chain + ((obj.kind, obj.params),), flip, out, out_leaves)
        else:
            xf = -1
            if chain:
                if chain not in xform_ids:
                    xform_ids[chain] = len(xforms)
                    xforms.append(_compose(chain))
                xf = xform_ids[chain]
            if obj.kind == g.PATCH:
                if xf >= 0:
                    raise ValueError("bicubic patches cannot be instanced (transform the control net instead)")
                net = np.asarray(obj.params, dtype=np.float64).reshape(4, 4, 3)
                for sub, su0, sv0, ssize in split_patch(net):       # 16 sub-patches share one logical id
                    patches.append(sub.reshape(48))
                    out.append((obj.kind, flip, mat_id(obj.material), xf, (len(patches) - 1, su0, sv0, ssize), len(out_leaves)))
                out_leaves.append(obj)
                return
            out.append((obj.kind, flip, mat_id(obj.material), xf, obj.params, len(out_leaves)))
            out_leaves.append(obj)

    for o in scene.obj_list:
        walk(o, (), 0)

    n_surface = len(prims)
    allp = prims + boundary                  # boundary primitives (flag 2) form a suffix, outside the LBVH
    P = np.zeros(len(allp), dtype=PRIM_DTYPE)
    logical = np.zeros(len(allp), dtype=np.int32)
    for i, (kind, flip, m, xf, prm, lid) in enumerate(allp):
        P[i]["type"], P[i]["flags"], P[i]["material"], P[i]["xform"] = kind, flip, m, xf
        prm = [n_surface + q[1] if isinstance(q, tuple) else q for q in prm]
        P[i]["p"][:len(prm)] = prm
        logical[i] = lid if lid >= 0 else len(leaves) + (i - n_surface)
        P[i]["p"][15] = logical[i] + 1
    first_of_logical = np.full(int(logical.max()) + 1 if len(logical) else 0, -1, dtype=np.int32)
    for i in range(len(allp) - 1, -1, -1):
        first_of_logical[logical[i]] = i
    X = np.zeros(len(xforms), dtype=XFORM_DTYPE)
    for i, (s, c, off) in enumerate(xforms):
        X[i]["sin_t"], X[i]["cos_t"], X[i]["off"] = s, c, off
    M = np.zeros(len(mats), dtype=MATERIAL_DTYPE)
    for i, (m, tx) in enumerate(mats):
        M[i]["kind"], M[i]["tex"], M[i]["param"] = m.kind, tx, m.param
    T = np.zeros(len(texs), dtype=TEXTURE_DTYPE)
    for i, (tx, even, odd) in enumerate(texs):
        T[i]["kind"], T[i]["even"], T[i]["odd"], T[i]["scale"], T[i]["rgb"] = tx.kind, even, odd, tx.scale, tx.rgb
    C = np.zeros(1, dtype=CAMERA_DTYPE)
    if scene.camera is not None:
        C.view("<f4")[:] = np.asarray(camera_to_floats(scene.camera), dtype=np.float32)
    PT = np.asarray(patches, dtype=np.float32).reshape(-1, 48)
    return FlatScene(P, PT, logical, first_of_logical, X, M, T, C, sky_kind(scene.sky_function), leaves, [m for m, _ in mats], [tx for tx, _, _ in texs],
                     np.concatenate([im[3] for im in images]).astype(np.uint8) if images else np.zeros(0, dtype=np.uint8),
                     np.asarray([im[:3] for im in images], dtype=np.int32).reshape(-1, 3))
