"""Fixed ray batches for GPU<->oracle parity (SURVEY.md §8c "fixed-ray-batch design").  All rays
are generated in fp32 and widened to f64 for the oracle, so both sides see identical inputs."""
import numpy as np


def camera_grid(renderer, w=64, h=64):
    """(i) all primary rays of a w x h camera grid (jitter from the shared Philox stream)."""
    p = renderer.params(w, h, 0, 1)
    pix = np.arange(w * h, dtype=np.int32)
    return renderer.eval_raygen(p, pix, np.zeros_like(pix))


def random_rays(bounds, n, seed):
    """(ii) random rays: origins inside the (slightly enlarged) scene bbox, uniform directions."""
    rs = np.random.RandomState(seed)
    lo, hi = np.asarray(bounds[0], np.float64), np.asarray(bounds[1], np.float64)
    ext = hi - lo
    o = lo - 0.1 * ext + rs.random_sample((n, 3)) * 1.2 * ext
    d = rs.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rs.uniform(0.5, 2.0, size=(n, 1))          # un-normalised directions, like camera rays
    t = rs.random_sample((n, 1))
    return np.concatenate([o, d, t], axis=1).astype(np.float32)


def interest_bounds(flat, clip=30.0):
    """bbox of the primitives' anchor points, huge ground spheres clipped to +-clip."""
    pts = []
    for p in flat.prims:
        q = p["p"]
        if p["type"] == 6 or (p["flags"] & 2 and p["xform"] >= 0):
            continue
        if p["type"] == 8:
            continue
        if p["type"] == 7:
            pts += [c for c in flat.patches[int(q[0])].reshape(16, 3)]
            continue            # media / instanced boundary shapes: covered by the other primitives' extent
        if p["type"] in (0, 1):
            pts.append(q[0:3]);
        elif p["type"] == 2:
            pts += [(q[0], q[2], q[4]), (q[1], q[3], q[4])]
        elif p["type"] == 3:
            pts += [(q[0], q[4], q[2]), (q[1], q[4], q[3])]
        elif p["type"] == 4:
            pts += [(q[4], q[0], q[2]), (q[4], q[1], q[3])]
        else:
            pts += [q[0:3], q[3:6], q[6:9], q[9:12]]
    pts = np.asarray(pts, np.float64)
    med = np.median(pts, axis=0)
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    span = max(np.max(np.percentile(pts, 90, axis=0) - np.percentile(pts, 10, axis=0)), 1.0)
    lo = np.maximum(lo, med - clip * span)
    hi = np.minimum(hi, med + clip * span)
    return lo - 1.0, hi + 1.0


def compare(gpu, orc, orc32, second_t, rel=1e-4):
    """Parity verdict of one batch.  Near-ties (0 < |t2 - t1| < 1e-5 * t1 in the f64 oracle) and rays on
    which the reference algorithm itself flips under fp32 rounding (f32 oracle != f64 oracle) are
    filtered and counted; on every other ray the primitive id must match exactly and t / normal /
    uv must agree within `rel` relative error."""
    n = len(gpu)
    t1 = orc["t"]
    near_tie = (orc["prim"] >= 0) & (np.abs(second_t - t1) < 1e-5 * np.abs(t1)) & (second_t != t1)   # exact ties are kept: the tie rule decides them
    unstable = orc32["prim"] != orc["prim"]
    keep = ~(near_tie | unstable)
    id_ok = gpu["prim"] == orc["prim"]
    hit = keep & id_ok & (orc["prim"] >= 0)
    t_err = np.abs(gpu["t"][hit] - t1[hit]) / np.maximum(np.abs(t1[hit]), 1e-30)
    nlen = np.maximum(np.linalg.norm(orc["n"][hit], axis=1), 1e-30)
    n_err = np.linalg.norm(gpu["n"][hit] - orc["n"][hit], axis=1) / nlen
    p_err = np.linalg.norm(gpu["p"][hit] - orc["p"][hit], axis=1) / np.maximum(np.linalg.norm(orc["p"][hit], axis=1), 1.0)
    return dict(n=n, filtered_near_tie=int(near_tie.sum()), filtered_unstable=int((unstable & ~near_tie).sum()),
                id_mismatch=int((keep & ~id_ok).sum()), id_mismatch_idx=np.nonzero(keep & ~id_ok)[0],
                t_err_max=float(t_err.max()) if len(t_err) else 0.0, t_bad=int((t_err > rel).sum()),
                n_err_max=float(n_err.max()) if len(n_err) else 0.0, n_bad=int((n_err > rel).sum()),
                p_err_max=float(p_err.max()) if len(p_err) else 0.0, hit_mask=hit, keep=keep)
