"""texture.scm — texture CONSTRUCTORS.  The reference returns `#(value-fn)` closures; here each
constructor records a table row instead (evaluation is the CUDA shade kernel's job)."""
from dataclasses import dataclass
from typing import Optional, Tuple

CONSTANT, CHECKER, NOISE, MARBLE, IMAGE = 0, 1, 2, 3, 4


@dataclass(frozen=True, eq=False)
class Texture:
    kind: int
    rgb: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    scale: float = 0.0
    even: Optional["Texture"] = None
    odd: Optional["Texture"] = None
    image: Optional[object] = None        # (ny, nx, 3) uint8 array, top row first (image-texture)


def constant_texture(color):                 # texture.scm:12-14
    return Texture(CONSTANT, rgb=tuple(float(c) for c in color))


def checker_texture(even_tex, odd_tex):      # texture.scm:16-23 (nestable)
    return Texture(CHECKER, even=even_tex, odd=odd_tex)


def noise_texture(sc):                       # texture.scm:25-28
    return Texture(NOISE, scale=float(sc))


def marble_texture(sc):                      # texture.scm:30-34
    return Texture(MARBLE, scale=float(sc))


def image_texture(data, nx, ny):             # texture.scm:36-50
    """`data` = nx*ny*3 numbers in 0..255, row-major from the top row (the reference applies
    `floor->exact` to each element at lookup time; here once, when the texels are recorded)."""
    import numpy as np
    nx, ny = int(nx), int(ny)
    a = np.floor(np.asarray(data, dtype=np.float64).ravel())
    if nx < 1 or ny < 1 or a.size != 3 * nx * ny:
        raise ValueError(f"image-texture: expected {3 * nx * ny} values for {nx}x{ny}, got {a.size}")
    if a.min() < 0 or a.max() > 255:
        raise ValueError("image-texture: texel values must lie in 0..255")
    return Texture(IMAGE, image=np.ascontiguousarray(a.astype(np.uint8).reshape(ny, nx, 3)))
