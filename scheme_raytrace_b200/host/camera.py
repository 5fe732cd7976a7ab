"""camera.scm — make-camera (host side; get-ray is the ray-generation kernel)."""
import math
from . import vec as v


def make_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, time0, time1):
    """camera.scm:63-78 -> the 10-slot vector #(llc horiz vert origin w u v lens-radius time0 time1)."""
    theta = vfov * (math.pi / 180.0)
    half_height = math.tan(theta / 2)
    half_width = aspect * half_height
    w = v.unit(v.diff(lookfrom, lookat))
    u = v.unit(v.cross(vup, w))
    vv = v.cross(w, u)
    return (
        v.diff(lookfrom, v.scale(u, half_width * focus_dist), v.scale(vv, half_height * focus_dist), v.scale(w, focus_dist)),
        v.scale(u, 2 * half_width * focus_dist),
        v.scale(vv, 2 * half_height * focus_dist),
        tuple(float(c) for c in lookfrom), w, u, vv,
        aperture / 2, float(time0), float(time1),
    )


def lower_left_corner(c): return c[0]     # camera.scm:33-61
def horizontal(c): return c[1]
def vertical(c): return c[2]
def origin(c): return c[3]
def w(c): return c[4]
def u(c): return c[5]
def v_(c): return c[6]
def lens_radius(c): return c[7]
def time0(c): return c[8]
def time1(c): return c[9]


def camera_to_floats(c):
    """24 floats in srt.h SrtCamera order."""
    out = []
    for i in range(7):
        out.extend(c[i])
    out.extend([c[7], c[8], c[9]])
    return out
