"""geometry.scm — object CONSTRUCTORS with the reference's names, arity and argument order.

The reference returns closure vectors `#(hit-fn bbox-fn material ...)`; here every constructor
records a node of a small object tree (same nesting as the closures).  `flatten.flatten_scene`
turns the tree into the flat primitive SoA / transform table the C-ABI takes; hit testing is the
CUDA extend kernel's job.  `make-bvh-node` / `make-bvh-with-sah` are pass-through grouping hints:
the GPU LBVH is built over the flattened leaves.
"""
import math
from dataclasses import dataclass, field
from typing import Any, List, Optional, Tuple

# node kinds (leaf kinds equal the SRT_PRIM_* codes of include/srt.h)
SPHERE, MOVING_SPHERE, XY_RECT, XZ_RECT, YZ_RECT, BEZIER, CONSTANT_MEDIUM, PATCH, KLEIN = 0, 1, 2, 3, 4, 5, 6, 7, 8
FLIP, LIST, TRANSLATE, ROTATE_Y = 16, 17, 18, 19
LEAF_KINDS = (SPHERE, MOVING_SPHERE, XY_RECT, XZ_RECT, YZ_RECT, BEZIER, CONSTANT_MEDIUM, PATCH, KLEIN)


@dataclass(eq=False)
class Obj:
    kind: int
    material: Any = None
    params: Tuple[float, ...] = ()
    children: List["Obj"] = field(default_factory=list)


@dataclass(eq=False)
class Scene:                                    # geometry.scm:52-56  #(hit-fn #f obj-list camera sky-fn)
    obj_list: List[Obj]
    camera: Any
    sky_function: Any


def make_scene(obj_list, camera=None, sky_function=None):      # geometry.scm:52
    return Scene(list(obj_list), camera, sky_function)


def scene_obj_list(scene): return scene.obj_list               # geometry.scm:17
def scene_num_obj(scene): return len(scene.obj_list)           # geometry.scm:24
def scene_camera(scene): return scene.camera                   # geometry.scm:27
def scene_sky_function(scene): return scene.sky_function       # geometry.scm:30
def material(obj): return obj.material                         # geometry.scm:61


def make_sphere(center, radius, material):                     # geometry.scm:146 (negative radius allowed)
    return Obj(SPHERE, material, (*map(float, center), float(radius)))


def make_moving_sphere(center0, center1, time0, time1, radius, material):   # geometry.scm:177
    return Obj(MOVING_SPHERE, material,
               (*map(float, center0), float(radius), *map(float, center1), float(time0), float(time1)))


def make_xy_rect(x0, x1, y0, y1, k, material):                 # geometry.scm:376
    return Obj(XY_RECT, material, tuple(map(float, (x0, x1, y0, y1, k))))


def make_xz_rect(x0, x1, z0, z1, k, material):                 # geometry.scm:395
    return Obj(XZ_RECT, material, tuple(map(float, (x0, x1, z0, z1, k))))


def make_yz_rect(y0, y1, z0, z1, k, material):                 # geometry.scm:414
    return Obj(YZ_RECT, material, tuple(map(float, (y0, y1, z0, z1, k))))


def flip_normals(obj):                                         # geometry.scm:433
    return Obj(FLIP, obj.material, (), [obj])


def make_box(p0, p1, material):                                # geometry.scm:444-463: 6 rects, this order
    faces = [
        make_xy_rect(p0[0], p1[0], p0[1], p1[1], p1[2], material),
        flip_normals(make_xy_rect(p0[0], p1[0], p0[1], p1[1], p0[2], material)),
        make_xz_rect(p0[0], p1[0], p0[2], p1[2], p1[1], material),
        flip_normals(make_xz_rect(p0[0], p1[0], p0[2], p1[2], p0[1], material)),
        make_yz_rect(p0[1], p1[1], p0[2], p1[2], p1[0], material),
        flip_normals(make_yz_rect(p0[1], p1[1], p0[2], p1[2], p0[0], material)),
    ]
    return Obj(LIST, material, (), faces)


def translate(obj, offset):                                    # geometry.scm:465 (no material slot upstream)
    return Obj(TRANSLATE, None, tuple(map(float, offset)), [obj])


def rotate_y(obj, angle):                                      # geometry.scm:483 (angle in degrees)
    radians = (math.pi / 180.0) * angle
    return Obj(ROTATE_Y, obj.material, (math.sin(radians), math.cos(radians), float(angle)), [obj])


def make_bvh_node(obj_list, time0=0, time1=0):                 # geometry.scm:226 -> grouping only
    return Obj(LIST, None, (), list(obj_list))


def make_bvh_with_sah(obj_list, time0=0, time1=0):             # geometry.scm:294 -> grouping only
    return Obj(LIST, None, (), list(obj_list))


def make_constant_medium(obj, density, a):                     # geometry.scm:545-578
    """Volume bounded by `obj` (any object tree); phase function = lambertian(a) exactly like
    upstream (make-isotropic is commented out at geometry.scm:546)."""
    from .material import make_lambertian
    return Obj(CONSTANT_MEDIUM, make_lambertian(a), (float(density),), [obj])


def make_klein(center, material):                              # geometry.scm:644-664
    """Klein / IIS sphere-traced fractal.  It has no bounding box upstream (geometry.scm:662-663),
    so it stays out of the LBVH and is tested before traversal (at most 8 such primitives)."""
    return Obj(KLEIN, material, tuple(map(float, center)))
