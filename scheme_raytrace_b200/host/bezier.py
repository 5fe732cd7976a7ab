"""bezier.scm — make-bezier: cubic Bezier CURVE with circular width (constructor only)."""
from .geometry import Obj, BEZIER, PATCH


def make_bezier(a, b, c, d, width, material):                  # bezier.scm:61
    return Obj(BEZIER, material, (*map(float, a), *map(float, b), *map(float, c), *map(float, d), float(width)))


def bezier_cp(bez, index):                                     # bezier.scm:228
    return tuple(bez.params[3 * index:3 * index + 3])


def make_bezier_patch(control_points, material):
    """Bicubic Bezier PATCH: 4x4 control points P[i][j] (i along u, j along v).  North-star
    extension — the reference only has the curve above (bezier.scm:61); parity unpinned."""
    cps = [tuple(map(float, control_points[i][j])) for i in range(4) for j in range(4)]
    return Obj(PATCH, material, tuple(c for p in cps for c in p))
