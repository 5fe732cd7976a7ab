"""Print the surface-area cost table srt_scene_commit evaluates when it decides which outlier
primitives stay outside the LBVH (csrc/srt_api.cu), for a few scenes.  Needs a GPU.
    SRT_DEBUG_COMMIT=1 python tools/commit_costs.py [scene ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SRT_DEBUG_COMMIT", "1")
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import scenes

SCENES = {"cfg2": scenes.cfg2_random_spheres, "cfg3": scenes.cfg3_next_week, "cfg5": scenes.cfg5_patches,
          "bvh100": scenes.test_scene_bvh, "cfg5_curves": scenes.test_bezier}
for name in (sys.argv[1:] or list(SCENES)):
    print(name, file=sys.stderr)
    r = srt.Renderer(SCENES[name](64, 64), device=0)
    items, glob = r.bvh_items()
    print("  outside the tree:", [(int(g), int(r.flat.prims["type"][g])) for g in glob], "(prim id, type);", len(items), "items in the tree", file=sys.stderr)
    r.close()
