#!/usr/bin/env python
"""Commit (H2D + LBVH build) time of every named config: best of 20 re-commits, device events and host wall clock."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import scenes
for name in sys.argv[1:] or ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "cfg5_teapot"):
    cfg = scenes.CONFIGS[name]
    r = srt.Renderer(cfg["scene"](cfg["width"], cfg["height"]), device=0)
    best_dev, best_wall = 1e9, 1e9
    for _ in range(20):
        t0 = time.perf_counter(); r.commit(); dt = (time.perf_counter() - t0) * 1e3
        img, st = r.render(64, 64, 1)
        best_dev, best_wall = min(best_dev, st.ms_commit), min(best_wall, dt)
    print(f"commit {name}: device {best_dev:.3f} ms, wall (set_* + commit) {best_wall:.3f} ms, {len(r.flat.prims)} prims, {st.bvh_nodes} nodes")
    r.close()
