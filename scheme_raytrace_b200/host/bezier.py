"""bezier.scm — make-bezier: cubic Bezier CURVE with circular width (constructor only)."""
from .geometry import Obj, BEZIER


def make_bezier(a, b, c, d, width, material):                  # bezier.scm:61
    return Obj(BEZIER, material, (*map(float, a), *map(float, b), *map(float, c), *map(float, d), float(width)))


def bezier_cp(bez, index):                                     # bezier.scm:228
    return tuple(bez.params[3 * index:3 * index + 3])
