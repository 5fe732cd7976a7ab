"""material.scm — material CONSTRUCTORS (table rows; scatter/emitted run in the shade kernel)."""
from dataclasses import dataclass
from typing import Optional
from .texture import Texture

LAMBERTIAN, METAL, DIELECTRIC, DIFFUSE_LIGHT, ISOTROPIC = 0, 1, 2, 3, 4


@dataclass(frozen=True, eq=False)
class Material:
    kind: int
    tex: Optional[Texture] = None
    param: float = 0.0


def make_lambertian(albedo):                 # material.scm:24-39
    return Material(LAMBERTIAN, albedo)


def make_metal(albedo, fuzz):                # material.scm:45-57
    return Material(METAL, albedo, float(fuzz))


def make_dielectric(ref_idx):                # material.scm:76-101
    return Material(DIELECTRIC, None, float(ref_idx))


def make_diffuse_light(emit):                # material.scm:103-111
    return Material(DIFFUSE_LIGHT, emit)


def make_isotropic(albedo):                  # absent upstream (geometry.scm:546 comment); book semantics
    return Material(ISOTROPIC, albedo)
