"""points.scm — CSV points -> Catmull-Rom-style (tightness 0.5) cubic Bezier control points."""
from . import vec as v
from .bezier import make_bezier


def load_points(file_name, magnitude):                         # points.scm:10-20
    pts = []
    with open(file_name) as f:
        for line in f:
            line = line.strip()
            if line:
                pts.append(v.vec3(*[magnitude * float(p) for p in line.split(",")]))
    return pts


def calc_bezier_cp(pt, p1, p2, p3):                            # points.scm:23-26
    d1 = v.scale(v.diff(p2, pt), 1 / 6)
    d2 = v.scale(v.diff(p3, p1), 1 / 6)
    return [p1, v.sum(p1, d1), v.diff(p2, d2), p2]


def points_to_bezier(points):                                  # points.scm:28-41 points->bezier
    last = len(points) - 2
    return [calc_bezier_cp(points[i - 1], points[i], points[i + 1], points[i + 2]) for i in range(1, last)]


def bezier_to_objs(beziers, width, material):                  # points.scm:43-50 bezier->objs
    return [make_bezier(b[0], b[1], b[2], b[3], width, material) for b in beziers]
