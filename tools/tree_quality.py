"""Offline (CPU) estimate of what a better tree than the Morton LBVH would buy the extend kernel: node steps and
primitive tests per ray of cfg2-like sphere scenes for (a) the LBVH of oracle/lbvh_ref.cpp, (b) PLOC (bottom-up
nearest-neighbour merging over the Morton order), (c) a top-down full-sweep SAH tree.  Traversal = the kernel's
(both child boxes tested per step, near child first, prune against the best hit).
Usage: python tools/tree_quality.py [cfg2] [n_rays]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as orc
from scheme_raytrace_b200.host import scenes, flatten


def area(mn, mx):
    d = np.maximum(mx - mn, 0)
    return 2 * (d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0])


class Tree:          # children: >=0 internal node, <0 ~leaf item; box per node
    def __init__(self):
        self.left, self.right, self.mn, self.mx = [], [], [], []

    def add(self, l, r, mn, mx):
        self.left.append(l); self.right.append(r); self.mn.append(mn); self.mx.append(mx)
        return len(self.left) - 1


def from_lbvh(aabbs):
    keys, order, nodes = orc.lbvh_build(aabbs)
    t = Tree()
    def box(ref):
        if ref < 0:
            b = aabbs[order[~ref]] if False else aabbs[~ref]
            return b[:3], b[3:]
        return t.mn[ref], t.mx[ref]
    # rebuild boxes bottom-up by recursion from root 0; leaves in lbvh_ref are ~primitive index
    n = len(nodes)
    t.left = [int(x) for x in nodes["left"]]; t.right = [int(x) for x in nodes["right"]]
    t.mn = [None] * n; t.mx = [None] * n
    sys.setrecursionlimit(10000)
    def rec(i):
        bs = []
        for c in (t.left[i], t.right[i]):
            if c < 0:
                b = aabbs[~c]; bs.append((b[:3], b[3:]))
            else:
                rec(c); bs.append((t.mn[c], t.mx[c]))
        t.mn[i] = np.minimum(bs[0][0], bs[1][0]); t.mx[i] = np.maximum(bs[0][1], bs[1][1])
    rec(0)
    return t, 0


def ploc(aabbs, radius=16):
    keys, order, _ = orc.lbvh_build(aabbs)
    t = Tree()
    cl = [(~int(p), aabbs[p][:3].copy(), aabbs[p][3:].copy()) for p in order]      # (ref, mn, mx) in Morton order
    while len(cl) > 1:
        n = len(cl)
        mn = np.array([c[1] for c in cl]); mx = np.array([c[2] for c in cl])
        nn = np.zeros(n, int)
        for i in range(n):
            lo, hi = max(0, i - radius), min(n, i + radius + 1)
            a = area(np.minimum(mn[i], mn[lo:hi]), np.maximum(mx[i], mx[lo:hi]))
            a[i - lo] = np.inf
            nn[i] = lo + int(np.argmin(a))
        out = []
        for i in range(n):
            j = nn[i]
            if nn[j] == i:
                if i < j:
                    bmn, bmx = np.minimum(mn[i], mn[j]), np.maximum(mx[i], mx[j])
                    out.append((t.add(cl[i][0], cl[j][0], bmn, bmx), bmn, bmx))
            else:
                out.append(cl[i])
        cl = out
    return t, cl[0][0]


def sah(aabbs):
    t = Tree()
    cen = 0.5 * (aabbs[:, :3] + aabbs[:, 3:])
    def rec(idx):
        if len(idx) == 1:
            return ~int(idx[0])
        best = (np.inf, None, None)
        for ax in range(3):
            o = idx[np.argsort(cen[idx, ax], kind="stable")]
            lmn = np.minimum.accumulate(aabbs[o, :3], 0); lmx = np.maximum.accumulate(aabbs[o, 3:], 0)
            rmn = np.minimum.accumulate(aabbs[o[::-1], :3], 0)[::-1]; rmx = np.maximum.accumulate(aabbs[o[::-1], 3:], 0)[::-1]
            k = np.arange(1, len(o))
            c = area(lmn[:-1], lmx[:-1]) * k + area(rmn[1:], rmx[1:]) * (len(o) - k)
            j = int(np.argmin(c))
            if c[j] < best[0]:
                best = (c[j], o, j + 1)
        _, o, s = best
        l = rec(o[:s]); r = rec(o[s:])
        return t.add(l, r, aabbs[o, :3].min(0), aabbs[o, 3:].max(0))
    sys.setrecursionlimit(10000)
    root = rec(np.arange(len(aabbs)))
    return t, root


def sah_cost(t, root, aabbs):
    ra = area(t.mn[root], t.mx[root]); c = 0.0
    for i in range(len(t.left)):
        for ch in (t.left[i], t.right[i]):
            b = (aabbs[~ch][:3], aabbs[~ch][3:]) if ch < 0 else (t.mn[ch], t.mx[ch])
            c += area(*b) / ra          # every child box is tested once per visit of its parent
    return c


def depth(t, root):
    def rec(i): return 0 if i < 0 else 1 + max(rec(t.left[i]), rec(t.right[i]))
    return rec(root)


def slab(o, inv, mn, mx, tmax):
    t0 = (mn - o) * inv; t1 = (mx - o) * inv
    lo = np.minimum(t0, t1).max(); hi = np.maximum(t0, t1).min()
    return (lo, True) if (lo <= hi and hi > 1e-3 and lo < tmax) else (np.inf, False)


def trace(t, root, aabbs, cen, rad, o, d):
    inv = 1.0 / d; best = np.inf; steps = tests = 0; hitp = -1
    stack = [root]
    while stack:
        i = stack.pop()
        if i < 0:
            p = ~i; tests += 1
            oc = o - cen[p]; b = oc @ d; c = oc @ oc - rad[p] ** 2; a = d @ d
            disc = b * b - a * c
            if disc > 0:
                s = np.sqrt(disc)
                for tt in ((-b - s) / a, (-b + s) / a):
                    if 1e-3 < tt < best:
                        best = tt; hitp = p; break
            continue
        steps += 1
        ch = []
        for c in (t.left[i], t.right[i]):
            b = (aabbs[~c][:3], aabbs[~c][3:]) if c < 0 else (t.mn[c], t.mx[c])
            lo, ok = slab(o, inv, b[0], b[1], best)
            if ok: ch.append((lo, c))
        ch.sort(key=lambda x: -x[0])          # far pushed first
        for _, c in ch: stack.append(c)
    return steps, tests, best, hitp


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    nrays = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    cfg = scenes.CONFIGS[name]
    sc = cfg["scene"](cfg["width"], cfg["height"])
    fs = flatten.flatten_scene(sc)
    pr = fs.prims
    sel = [i for i in range(len(pr)) if pr["type"][i] == 0 and abs(pr["p"][i][3]) < 100]
    cen = np.array([pr["p"][i][:3] for i in sel], np.float64); rad = np.array([abs(pr["p"][i][3]) for i in sel], np.float64)
    aabbs = np.concatenate([cen - rad[:, None], cen + rad[:, None]], 1).astype(np.float32)
    print(name, len(sel), "spheres in the tree")
    cam = fs.camera
    rs = np.random.RandomState(1)
    trees = {"lbvh": from_lbvh(aabbs), "ploc r=8": ploc(aabbs, 8), "ploc r=16": ploc(aabbs, 16), "ploc r=32": ploc(aabbs, 32), "sah sweep": sah(aabbs)}
    # rays: camera rays, then diffuse-ish bounce rays from their hit points (ground plane y=0 included as a hit surface)
    llc, hz, vt, og = (np.array(cam[k][0] if cam[k].ndim > 1 else cam[k], np.float64) for k in ("llc", "horiz", "vert", "origin"))
    rays = []
    for _ in range(nrays):
        s, v = rs.rand(2)
        rays.append((og, llc + s * hz + v * vt - og))
    for gen in range(3):
        res = {k: [] for k in trees}
        nxt = []
        for (o, d) in rays:
            for k, (t, root) in trees.items():
                st, te, best, hp = trace(t, root, aabbs.astype(np.float64), cen, rad, o, d)
                res[k].append((st, te))
            # ground plane
            tg = -o[1] / d[1] if d[1] < 0 else np.inf
            if min(best, tg) < np.inf:
                tt = min(best, tg); p = o + tt * d
                n = np.array([0, 1.0, 0]) if tg < best else (p - cen[hp]) / rad[hp]
                dd = n + rs.normal(size=3) * 0.7
                nxt.append((p + 1e-4 * n, dd))
        print(f"generation {gen}: {len(rays)} rays")
        base = np.array(res["lbvh"]).mean(0)
        for k in trees:
            a = np.array(res[k]); m = a.mean(0)
            # warp-level: max node steps over groups of 32 consecutive rays
            g = a[: len(a) // 32 * 32, 0].reshape(-1, 32).max(1).mean() if len(a) >= 32 else 0
            print(f"  {k:10s} node steps {m[0]:6.2f} ({m[0] / base[0]:.2f}x)  prim tests {m[1]:5.2f} ({m[1] / base[1]:.2f}x)  warp-max steps {g:6.2f}")
        rays = nxt
    for k, (t, root) in trees.items():
        print(f"{k:10s} SAH cost {sah_cost(t, root, aabbs):7.2f}  depth {depth(t, root)}")


if __name__ == "__main__":
    main()
