"""Flat scene file (text) = the POD tables of include/srt.h, one record per line.  Written by this
module and by scheme/srt-scene.scm (the Gauche host), read by cli/srt_render.cpp.

    srt-scene 1
    sky <kind>
    camera <24 floats>
    perlin-ranvec <768 floats>
    perlin-perm <768 ints>            (perm-x, perm-y, perm-z)
    textures <n>   then n lines:  kind even odd scale r g b
    materials <n>  then n lines:  kind tex param
    xforms <n>     then n lines:  sin cos ox oy oz
    patches <n>    then n lines:  48 floats
    prims <n>      then n lines:  type flags material xform p0..p15
    lights <n>     then 1 line :  ids
    images <n>     then n lines:  nx ny followed by 3*nx*ny texel values 0..255 (optional section)
"""
import numpy as np


def _f(a):
    return " ".join(repr(float(x)) for x in np.asarray(a, dtype=np.float32).ravel())


def write_scene_file(path, flat, perlin, lights=()):
    rv, px, py, pz = perlin
    with open(path, "w") as f:
        f.write("srt-scene 1\n")
        f.write(f"sky {int(flat.sky)}\n")
        f.write("camera " + _f(flat.camera.view("<f4")) + "\n")
        f.write("perlin-ranvec " + _f(rv) + "\n")
        f.write("perlin-perm " + " ".join(str(int(x)) for x in np.concatenate([px, py, pz])) + "\n")
        f.write(f"textures {len(flat.textures)}\n")
        for t in flat.textures:
            f.write(f"{int(t['kind'])} {int(t['even'])} {int(t['odd'])} {_f([t['scale']])} {_f(t['rgb'])}\n")
        f.write(f"materials {len(flat.materials)}\n")
        for m in flat.materials:
            f.write(f"{int(m['kind'])} {int(m['tex'])} {_f([m['param']])}\n")
        f.write(f"xforms {len(flat.xforms)}\n")
        for x in flat.xforms:
            f.write(f"{_f([x['sin_t'], x['cos_t']])} {_f(x['off'])}\n")
        f.write(f"patches {len(flat.patches)}\n")
        for p in flat.patches:
            f.write(_f(p) + "\n")
        f.write(f"prims {len(flat.prims)}\n")
        for p in flat.prims:
            f.write(f"{int(p['type'])} {int(p['flags'])} {int(p['material'])} {int(p['xform'])} {_f(p['p'])}\n")
        f.write(f"lights {len(lights)}\n")
        f.write(" ".join(str(int(i)) for i in lights) + "\n")
        f.write(f"images {len(flat.image_dims)}\n")
        for nx, ny, off in flat.image_dims:
            f.write(f"{int(nx)} {int(ny)} " + " ".join(str(int(x)) for x in flat.image_texels[off:off + 3 * nx * ny]) + "\n")
