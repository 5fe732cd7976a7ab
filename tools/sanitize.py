#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck) at CI size: commit (GPU LBVH build with the
outlier search), fixed ray batch, batch render (graph + drain kernel), progressive steps and the no-graph / no-tail paths, on
the scenes that cover every kernel variant: cfg1 (spheres), cfg3 (moving spheres, Perlin), cfg4 (rects, instances, mixture
estimator), cornell_smoke (media), teapot (patches), test_bezier (curves), a 3000-sphere cloud (global-memory tree).

    compute-sanitizer --tool racecheck python tools/sanitize.py [scene ...]
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scheme_raytrace_b200 as srt                     # noqa: E402
from scheme_raytrace_b200.host import scenes           # noqa: E402

SCENES = {"cfg1": scenes.cfg1_weekend, "cfg3": scenes.cfg3_next_week, "cfg4": scenes.cfg4_cornell_box, "smoke": scenes.cornell_smoke,
          "teapot": scenes.teapot_scene, "bezier": scenes.test_bezier, "cloud": lambda w, h: scenes.sphere_cloud(3000, w, h)}


def main():
    names = sys.argv[1:] or list(SCENES)
    w, h = 24, 16
    for name in names:
        r = srt.Renderer(SCENES[name](w, h), device=0, lights=[5] if name == "cfg4" else ())
        rs = np.random.RandomState(1)
        rays = np.concatenate([rs.normal(size=(512, 3)) * 3, rs.normal(size=(512, 3)), rs.random_sample((512, 1))], axis=1).astype(np.float32)
        r.trace_batch(rays)
        img, st = r.render(w, h, 3, max_depth=12, seed=2, wave_spp=1, estimator=1 if name == "cfg4" else 0)
        p = r.params(w, h, 0, 2, 12, 2)
        for flags in ((0, 1, 1, 0), (0, 0, 1, 1), (1, 0, 1, 0)):          # no graph / no drain kernel / profile mode
            import ctypes as C
            from scheme_raytrace_b200.host import ffi
            p.reserved[0], p.reserved[1], p.reserved[2], p.reserved[3] = flags
            out = np.zeros((h, w, 3), np.float32)
            ffi.check(r.lib.srt_render_host(r.h, C.byref(p), out.ctypes.data_as(C.c_void_p), None), "render")
        r.progressive_step(w, h, 0, 1, max_depth=12)
        r.progressive_step(w, h, 1, 2, max_depth=12)
        r.commit()
        print(f"{name}: ok, {st.rays} rays, {st.kernel_launches} launches, finite={bool(np.all(np.isfinite(img)))}", flush=True)
        r.close()


if __name__ == "__main__":
    main()
