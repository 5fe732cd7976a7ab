"""Generator of the committed golden fixtures in this directory.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.json, *.npz

The reference (Gauche Scheme) cannot run in the build container and ships no test vectors of its
own, so the fixtures have two sources, stated per file:

* `kat_survey.json` — the known-answer vectors KAT1-10 of SURVEY.md §8c, typed in from the survey
  (derived there from the reference sources by an independent transliteration; KAT1-6 can be
  checked by hand).  NOT produced by the oracle: they are what pins it.
* `rays_<scene>.npz`, `image_<scene>.npz` — outputs of the CPU oracle (`oracle/srt_oracle.cpp`,
  f64) on seeded inputs, frozen here so that (a) the oracle cannot drift unnoticed
  (`tests/test_golden.py`, CPU) and (b) the CUDA path is compared with files that do not change
  when the oracle is rebuilt (`tests/test_gpu_golden.py`, through the C-ABI).

Nothing here reads /root/reference; the GPU box only uses the committed files.
"""
import hashlib
import json
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from scheme_raytrace_b200.host import scenes                       # noqa: E402
from scheme_raytrace_b200.host.flatten import flatten_scene        # noqa: E402
from tests import raybatch                                          # noqa: E402

N_RAYS = 2048
RAY_SCENES = {"cfg1": (scenes.cfg1_weekend, 200, 100), "cfg2": (scenes.cfg2_random_spheres, 120, 80),
              "cfg3": (scenes.cfg3_next_week, 80, 80), "cfg4": (scenes.cfg4_cornell_box, 64, 64),
              "bezier": (scenes.test_bezier, 64, 64), "smoke": (scenes.cornell_smoke, 64, 64),
              "teapot": (scenes.teapot_scene, 96, 54)}
IMAGE_SCENES = {"cfg1": (scenes.cfg1_weekend, 100, 50, 16, 1), "cfg2": (scenes.cfg2_random_spheres, 60, 40, 8, 2),
                "cfg3": (scenes.cfg3_next_week, 40, 40, 8, 3), "cfg4": (scenes.cfg4_cornell_box, 32, 32, 16, 4)}

KAT_SURVEY = {
    "source": "SURVEY.md §8c (independent f64 transliteration of the reference sources; KAT1-6 hand-verifiable)",
    "t_min": 0.001, "t_max": 999999999999.0,
    "sphere": [   # geometry.scm:146-175
        {"kat": 1, "center": [0, 0, -1], "radius": 0.5, "ray": [0, 0, 0, 0, 0, -1, 0], "t": 0.5, "p": [0, 0, -0.5], "n": [0, 0, 1]},
        {"kat": 2, "center": [0, 0, -1], "radius": 0.5, "ray": [0, 0, 0, 0, 0, -2, 0], "t": 0.25, "p": [0, 0, -0.5], "n": [0, 0, 1]},
        {"kat": 3, "center": [0, 0, -1], "radius": 0.5, "ray": [0, 0, -1, 0, 1, 0, 0], "t": 0.5, "p": [0, 0.5, -1], "n": [0, 1, 0]},
        {"kat": 4, "center": [-1, 0, -1], "radius": -0.45, "ray": [-1, 0, 0, 0, 0, -1, 0], "t": 0.55, "p": [-1, 0, -0.55], "n": [0, 0, -1]},
    ],
    "camera": [   # camera.scm:63-78
        {"kat": "5", "args": [[278, 278, -800], [278, 278, 0], [0, 1, 0], 40, 1, 0, 1, 0, 1],
         "llc": [278.3639702342662, 277.6360297657338, -799], "horiz": [-0.7279404685324047, 0, 0], "vert": [0, 0.7279404685324047, 0],
         "w": [0, 0, -1], "u": [-1, 0, 0], "v": [0, 1, 0], "centre_dir": [0, 0, 1]},
        {"kat": "5b", "args": [[0, 5, 5], [0, 0, 0], [0, 1, 0], 40, 1, 0, 1, 0, 1],
         "llc": [-0.36397023426620234, 4.035527398013764, 4.55025903961314], "horiz": [0.7279404685324047, 0, 0],
         "vert": [0, 0.5147316415993759, -0.5147316415993759], "centre_dir": [0, -0.7071067811865479, -0.7071067811865479]},
    ],
    "xz_rect": [  # geometry.scm:395-412
        {"kat": 6, "rect": [213, 343, 227, 332, 554], "ray": [278, 0, 279.5, 0, 1, 0, 0], "t": 554, "p": [278, 554, 279.5], "n": [0, 1, 0], "uv": [0.5, 0.5]},
    ],
    "bezier": {   # bezier.scm:61-223, curve of main.scm:259-263
        "cp": [[-1, 0, -1], [-0.8, 1, 1], [0.8, -1, 1], [1, 0, -1]], "width": 0.1, "origin": [0, 5, 5], "max_depth": 6,
        "cases": [
            {"kat": 7, "dir": "unit(curve(0.5)-o)", "aim": [0, 0, 0.5], "normalise": True, "hit": True, "converge_calls": 39,
             "t": 6.731228242402701, "p": [0, -0.003282549631530, 0.497045705331623], "n": [0, 0.7432941462471664, 0.6689647316224497]},
            {"kat": 8, "dir": "curve(0.5)-o", "aim": [0, 0, 0.5], "normalise": False, "hit": True,
             "t": 6.731228242402701, "p": [0, -28.65614121201351, -25.290527090812155], "n": [0, 5, 4.5]},
            {"kat": 9, "dir": "camera 5b centre ray", "raw_dir": [0, -0.7071067811865479, -0.7071067811865479], "hit": False, "converge_calls": 9},
            {"kat": 10, "dir": "unit(curve(0.1)-o)", "aim_param": 0.1, "normalise": True, "hit": True,
             "t": 7.340737709682441, "p": [-0.9039654664425595, 0.19918873061589348, -0.4791867748405565]},
        ]},
}


def flat_digest(flat):
    """sha256 over the flat scene tables: a fixture is only valid for the scene it was made from."""
    h = hashlib.sha256()
    for a in (flat.prims, flat.patches, flat.xforms, flat.materials, flat.textures, flat.camera):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def golden_rays(name):
    """The fixed batch of a scene: N_RAYS seeded random rays (fp32) inside the scene's interest
    bounds; the same call regenerates them in the tests."""
    fn, w, h = RAY_SCENES[name]
    scene = fn(w, h)
    flat = flatten_scene(scene)
    rays = raybatch.random_rays(raybatch.interest_bounds(flat), N_RAYS, 1234)
    return scene, flat, rays


def main():
    from oracle import oracle as O
    O.build()
    with open(os.path.join(HERE, "kat_survey.json"), "w") as f:
        json.dump(KAT_SURVEY, f, indent=1)
    for name in RAY_SCENES:
        scene, flat, rays = golden_rays(name)
        S = O.OracleScene(scene, flat=flat)
        r64 = rays.astype(np.float64)
        o = S.trace_batch(r64)
        o32 = S.trace_batch(r64, precision=32)
        t2 = S.second_best_t(r64, o["prim"])
        near_tie = (o["prim"] >= 0) & (np.abs(t2 - o["t"]) < 1e-5 * np.abs(o["t"])) & (t2 != o["t"])
        unstable = o32["prim"] != o["prim"]
        np.savez_compressed(os.path.join(HERE, f"rays_{name}.npz"), digest=flat_digest(flat), rays=rays,
                            prim=o["prim"].astype(np.int32), t=o["t"], p=o["p"], n=o["n"], uv=o["uv"],
                            filtered=(near_tie | unstable))
        print(f"rays_{name}: hits {np.mean(o['prim'] >= 0):.3f} filtered {int((near_tie | unstable).sum())}")
    for name, (fn, w, h, spp, seed) in IMAGE_SCENES.items():
        scene = fn(w, h)
        flat = flatten_scene(scene)
        S = O.OracleScene(scene, flat=flat)
        img, nrays = S.render(w, h, spp, max_depth=50, seed=seed)
        np.savez_compressed(os.path.join(HERE, f"image_{name}.npz"), digest=flat_digest(flat), rgb_sum=img.astype(np.float32),
                            width=w, height=h, spp=spp, seed=seed, max_depth=50, rays=nrays)
        print(f"image_{name}: {w}x{h}@{spp} rays {nrays}")


if __name__ == "__main__":
    main()
