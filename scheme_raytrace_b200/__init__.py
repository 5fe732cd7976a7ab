"""scheme_raytrace_b200 — B200-native per-sample radiance loop of soma-arc/scheme-raytrace.

Only what the hot path needs lives here: ``csrc/`` holds the hand-written sm_100a
CUDA kernels and the C-ABI (``include/srt.h``) built into ``libsrt.so``; ``host/`` mirrors the
reference's scene-construction API (geometry.scm, material.scm, texture.scm, camera.scm,
bezier.scm) and flattens scenes into the POD tables the C-ABI takes.  There is no CPU fallback:
rendering without the CUDA library fails loudly.
"""
from .host import vec, constant, texture, material, geometry, bezier, camera, perlin, points, pdf  # noqa: F401
from .host.flatten import flatten_scene, FlatScene  # noqa: F401
from .host.render import Renderer, ProgressiveRenderer, trace_all, save_as_ppm, correct_gamma_quantise  # noqa: F401
from .host import scenes  # noqa: F401

from .host.ffi import QUIRKS_REFERENCE  # noqa: F401,E402
