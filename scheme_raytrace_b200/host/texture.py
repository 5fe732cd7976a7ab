"""texture.scm — texture CONSTRUCTORS.  The reference returns `#(value-fn)` closures; here each
constructor records a table row instead (evaluation is the CUDA shade kernel's job)."""
from dataclasses import dataclass
from typing import Optional, Tuple

CONSTANT, CHECKER, NOISE, MARBLE = 0, 1, 2, 3


@dataclass(frozen=True, eq=False)
class Texture:
    kind: int
    rgb: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    scale: float = 0.0
    even: Optional["Texture"] = None
    odd: Optional["Texture"] = None


def constant_texture(color):                 # texture.scm:12-14
    return Texture(CONSTANT, rgb=tuple(float(c) for c in color))


def checker_texture(even_tex, odd_tex):      # texture.scm:16-23 (nestable)
    return Texture(CHECKER, even=even_tex, odd=odd_tex)


def noise_texture(sc):                       # texture.scm:25-28
    return Texture(NOISE, scale=float(sc))


def marble_texture(sc):                      # texture.scm:30-34
    return Texture(MARBLE, scale=float(sc))


def image_texture(data, nx, ny):             # texture.scm:36-50
    raise NotImplementedError("image-texture is out of scope (SURVEY.md §2 row 7: never instantiated upstream)")
