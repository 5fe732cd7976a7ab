"""perlin.scm:10-36 — table generation (host side).  The reference draws the tables from the
default MT19937 stream at module load; here the stream is numpy's MT19937 with an explicit seed
(the reference never seeds, so its tables are not reproducible — SURVEY.md §3.1)."""
import math
import numpy as np


def perlin_generate(seed=3):
    rs = np.random.RandomState(seed)
    rnd = rs.random_sample
    ranfloat = [rnd() for _ in range(256)]                      # perlin.scm:10-12 (+ranfloat+, unused)
    ranvec = []
    for _ in range(256):                                        # perlin.scm:14-18
        x, y, z = -1 + 2 * rnd(), -1 + 2 * rnd(), -1 + 2 * rnd()
        k = 1.0 / math.sqrt(x * x + y * y + z * z)
        ranvec.append((x * k, y * k, z * k))

    def perm():                                                 # perlin.scm:20-30
        p = list(range(256))
        for i in range(255, 0, -1):
            target = int(math.floor(rnd() * (i + 1)))
            p[i], p[target] = p[target], p[i]
        return p
    px, py, pz = perm(), perm(), perm()
    del ranfloat
    return (np.asarray(ranvec, dtype=np.float64), np.asarray(px, dtype=np.int32),
            np.asarray(py, dtype=np.int32), np.asarray(pz, dtype=np.int32))
