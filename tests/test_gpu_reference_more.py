"""GPU (-m gpu): the CUDA path against the SECOND set of reference-executed goldens (ref_prims2.json,
ref_scenes2.json: constant media, the Klein primitive, more scenes of main.scm including the ones that go through
the reference's own BVH builders).  Same bars as tests/test_gpu_reference.py (Klein normals 2e-2).  The medium's
free-flight draw is the one trace_batch assigns to (ray index, leaf) - the generator scripted the reference's
random-real with exactly those Philox uniforms."""
import pytest
from scheme_raytrace_b200.host import geometry as g, scenes
from tests.refspec import build_host, host_scene
from tests.test_gpu_reference import _check, load

pytestmark = pytest.mark.gpu
PRIMS2, SCENES2 = load("ref_prims2.json"), load("ref_scenes2.json")


@pytest.mark.parametrize("idx", range(len(PRIMS2["cases"])), ids=[c["name"] for c in PRIMS2["cases"]])
def test_reference_medium_and_klein_hits(idx):
    case = PRIMS2["cases"][idx]
    scene = g.make_scene([build_host(case["spec"])], scenes.default_camera(), scenes.sky_color)
    _check(scene, case, case["name"], PRIMS2["t_min"], PRIMS2["t_max"])


@pytest.mark.parametrize("idx", range(len(SCENES2["scenes"])), ids=[c["name"] for c in SCENES2["scenes"]])
def test_reference_more_scene_hits(idx):
    case = SCENES2["scenes"][idx]
    _check(host_scene(case["name"]), case, case["name"])
