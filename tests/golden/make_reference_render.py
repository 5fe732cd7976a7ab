"""Converged images rendered BY THE REFERENCE ITSELF (build container only; ~15 min on 8 cores).

    python tests/golden/make_reference_render.py            # rewrites tests/golden/ref_render.npz

The north star's second correctness criterion is "each scene converged at high spp must match the
reference's CPU render".  This script produces that CPU render: /root/reference/*.scm, unmodified, run by
oracle/minischeme.py; per pixel and sample it makes exactly the calls main.scm's trace-all makes
(main.scm:474-480): u = (x + random-real) / W, v = (y + random-real) / H, (cam:get-ray camera u v),
(color ray scene) - with +max-depth+ = 100 as upstream.  `random-real` here is an INDEPENDENT numpy
stream per worker (nothing to do with the Philox draws of the oracle / the CUDA path), so agreement with
those is a statement about distributions, not about replaying one path.

Stored per scene: the per-pixel sum and sum of squares of the radiance samples (so the tests can judge a
difference against the render's own Monte-Carlo standard error) and the sample count.
"""
import multiprocessing as mp
import os
import sys
import threading
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# scenes of main.scm that HEAD's `color` can run (lambertian / diffuse-light only) and that are lit
SCENES = {"cornell-box": dict(width=20, height=20, spp=256), "test-bezier": dict(width=24, height=24, spp=96),
          "cornell-smoke": dict(width=16, height=16, spp=256), "test-scene2": dict(width=16, height=16, spp=256)}
WORKERS = int(os.environ.get("SRT_WORKERS", "8"))
# a second batch with other seeds can be ADDED to an existing file (sums and sums of squares accumulate):
#   SRT_RENDER_BATCH=1 python tests/golden/make_reference_render.py
BATCH = int(os.environ.get("SRT_RENDER_BATCH", "0"))
ONLY = [x for x in os.environ.get("SRT_RENDER_SCENES", "").split(",") if x]          # e.g. SRT_RENDER_SCENES=test-bezier
# round 2: larger thumbnails into a second file, e.g.
#   SRT_RENDER_SIZE=32 SRT_RENDER_SPP=384 SRT_WORKERS=6 SRT_RENDER_OUT=ref_render32.npz SRT_RENDER_SCENES=cornell-box,test-bezier python ...
OUT = os.environ.get("SRT_RENDER_OUT", "ref_render.npz")
if os.environ.get("SRT_RENDER_SIZE"):
    for _c in SCENES.values():
        _c["width"] = _c["height"] = int(os.environ["SRT_RENDER_SIZE"])
if os.environ.get("SRT_RENDER_SPP"):
    for _c in SCENES.values():
        _c["spp"] = int(os.environ["SRT_RENDER_SPP"])


def _worker(args):
    name, width, height, spp, seed = args
    out = {}

    def work():
        sys.setrecursionlimit(200000)
        from oracle.minischeme import Sym
        from tests.golden import make_reference_golden as mk
        rng = mk.ScriptedRng()
        rng.script = [float(x) for x in np.random.RandomState(3).random_sample(256 + 3 * 256 + 3 * 255)]   # perlin.scm's load-time draws
        ref = mk.Ref(rng)
        main = ref.load_main(mk.MAIN_NAMES)
        main.vars[Sym("*size-x*")] = width
        main.vars[Sym("*size-y*")] = height
        rs = np.random.RandomState(seed)
        ref.it.random_real = lambda: float(rs.random_sample())
        scene = main.lookup(Sym(name))
        cam = ref.call("geometry", "scene-camera", scene)
        color = main.lookup(Sym("color"))
        s1, s2 = np.zeros((height, width, 3)), np.zeros((height, width, 3))
        for _ in range(spp):
            for y in range(height):
                for x in range(width):
                    u = (x + ref.it.random_real()) / width                  # main.scm:476-477
                    v = (y + ref.it.random_real()) / height
                    ray = ref.call("camera", "get-ray", cam, u, v)
                    c = np.asarray(ref.it.apply(color, [ray, scene]), np.float64)
                    s1[y, x] += c
                    s2[y, x] += c * c
        out["r"] = (s1, s2)
    threading.stack_size(512 * 1024 * 1024)
    th = threading.Thread(target=work)
    th.start()
    th.join()
    return out["r"]


def main():
    res = {}
    for si, (name, cfg) in enumerate(SCENES.items()):
        if ONLY and name not in ONLY:
            continue
        t0 = time.time()
        per = cfg["spp"] // WORKERS
        assert per * WORKERS == cfg["spp"]
        with mp.Pool(WORKERS) as pool:
            parts = pool.map(_worker, [(name, cfg["width"], cfg["height"], per, 100000 * BATCH + 1000 * (si + 1) + w) for w in range(WORKERS)])
        s1, s2 = sum(p[0] for p in parts), sum(p[1] for p in parts)
        key = name.replace("-", "_")
        res[key + "_sum"], res[key + "_sumsq"] = s1, s2
        res[key + "_meta"] = np.array([cfg["width"], cfg["height"], cfg["spp"]])
        print(f"{name}: {cfg['width']}x{cfg['height']} @ {cfg['spp']} spp, mean radiance {s1.mean() / cfg['spp']:.4f}, {time.time() - t0:.0f} s", flush=True)
    path = os.path.join(HERE, OUT)
    if os.path.exists(path) and (BATCH > 0 or ONLY):             # accumulate into / keep the other scenes of an existing file
        old = np.load(path)
        for k in list(res):
            if BATCH > 0 and k in old.files:
                res[k] = res[k] + old[k] if not k.endswith("_meta") else np.array([old[k][0], old[k][1], old[k][2] + res[k][2]])
        for k in old.files:
            res.setdefault(k, old[k])
    np.savez_compressed(path, **res)
    print("wrote", OUT, {k: int(res[k][2]) for k in res if k.endswith("_meta")})


if __name__ == "__main__":
    main()
