"""vec.scm — vec3 helpers for the HOST side (scene construction only; f64 like the reference's
f64vector).  On the device vec3 is register math in fp32 (csrc/srt_math.cuh)."""
import math


def vec3(x, y, z):            # vec.scm:7
    return (float(x), float(y), float(z))


def x(v): return v[0]         # vec.scm:9-11
def y(v): return v[1]
def z(v): return v[2]


def sum(*vs):                 # vec.scm:20-24  (shadows the builtin on purpose: reference name)
    a = vs[0]
    for b in vs[1:]:
        a = (a[0] + b[0], a[1] + b[1], a[2] + b[2])
    return a


def diff(*vs):                # vec.scm:26-33
    a = vs[0]
    for b in vs[1:]:
        a = (a[0] - b[0], a[1] - b[1], a[2] - b[2])
    return a


def prod(*vs):                # vec.scm:35-39
    a = vs[-1]
    for b in reversed(vs[:-1]):
        a = (b[0] * a[0], b[1] * a[1], b[2] * a[2])
    return a


def scale(v, k):              # vec.scm:41
    return (v[0] * k, v[1] * k, v[2] * k)


def dot(a, b):                # vec.scm:52
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def length(v):                # vec.scm:54
    return math.sqrt(dot(v, v))


def sq_len(v):                # vec.scm:57
    return dot(v, v)


def unit(v):                  # vec.scm:60-62
    k = 1.0 / length(v)
    return scale(v, k)


def cross(a, b):              # vec.scm:64-70
    return (a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1])
