"""Scene scripts: the reference's scenes (main.scm:31-426) written against the mirrored
constructor API, plus the five BASELINE.json configs as deterministic generators."""
import numpy as np
from . import vec as v
from . import geometry as g
from . import material as m
from . import texture as t
from . import bezier as b
from . import camera as cam


def sky_color(ray=None):        # main.scm:91-95 (evaluated on the device; this is the enum tag)
    raise RuntimeError("sky functions are evaluated by the CUDA shade kernel")


def black(ray=None):            # main.scm:97-98
    raise RuntimeError("sky functions are evaluated by the CUDA shade kernel")


def cornell_camera(size_x=200, size_y=200):      # main.scm:129-139
    return cam.make_camera(v.vec3(278, 278, -800), v.vec3(278, 278, 0), v.vec3(0, 1, 0), 40, size_x / size_y, 0, 1, 0, 1)


def default_camera(size_x=200, size_y=200):      # main.scm:141-153
    return cam.make_camera(v.vec3(0, 5, 5), v.vec3(0, 0, 0), v.vec3(0, 1, 0), 40, size_x / size_y, 0, 1, 0, 1)


def _checker():
    return t.checker_texture(t.constant_texture(v.vec3(0.2, 0.3, 0.1)), t.constant_texture(v.vec3(0.9, 0.9, 0.9)))


def test_scene_objects():                        # main.scm:155-172
    return [
        g.make_sphere(v.vec3(0, 0, -1), 0.5, m.make_lambertian(t.constant_texture(v.vec3(0.1, 0.2, 0.5)))),
        g.make_sphere(v.vec3(0, -100.5, -1), 100, m.make_lambertian(_checker())),
        g.make_sphere(v.vec3(1, 0, -1), 0.5, m.make_metal(t.constant_texture(v.vec3(0.8, 0.6, 0.2)), 0.3)),
        g.make_sphere(v.vec3(-1, 0, -1), 0.5, m.make_dielectric(1.5)),
        g.make_sphere(v.vec3(-1, 0, -1), -0.45, m.make_dielectric(1.5)),
    ]


def cfg1_weekend(size_x=200, size_y=100):
    """configs[0]: main.scm test-scene with the 2:1 Weekend camera and sky-color (SURVEY §8d cfg1:
    upstream's `black` sky renders an all-black image)."""
    c = cam.make_camera(v.vec3(0, 0, 0), v.vec3(0, 0, -1), v.vec3(0, 1, 0), 90, size_x / size_y, 0, 1, 0, 1)
    return g.make_scene(test_scene_objects(), c, sky_color)


def random_scene(seed=2, lo=-11, hi=11, moving=False, checker_ground=False, big_spheres=True):
    """main.scm:31-89 random-scene, generalised (upstream grid is [-5,10), moving lambertians,
    checker ground).  Draw order follows the reference loop; `push!` prepends, so the final list
    is in reverse creation order exactly like upstream."""
    rnd = np.random.RandomState(seed).random_sample
    obj_list = []
    ground_tex = _checker() if checker_ground else t.constant_texture(v.vec3(0.5, 0.5, 0.5))
    obj_list.insert(0, g.make_sphere(v.vec3(0, -1000, 0), 1000, m.make_lambertian(ground_tex)))
    for a in range(lo, hi):
        for bb in range(lo, hi):
            choose_mat = rnd()
            center = v.vec3(a + 0.9 * rnd(), 0.2, bb + 0.9 * rnd())
            if v.length(v.diff(center, v.vec3(4, 0.2, 0))) > 0.9:
                if choose_mat < 0.8:
                    if moving:
                        c1 = v.sum(center, v.vec3(0, 0.5 * rnd(), 0))
                    alb = t.constant_texture(v.vec3(rnd() * rnd(), rnd() * rnd(), rnd() * rnd()))
                    if moving:
                        obj_list.insert(0, g.make_moving_sphere(center, c1, 0, 1, 0.2, m.make_lambertian(alb)))
                    else:
                        obj_list.insert(0, g.make_sphere(center, 0.2, m.make_lambertian(alb)))
                elif choose_mat < 0.95:
                    alb = t.constant_texture(v.vec3(0.5 * (1 + rnd()), 0.5 * (1 + rnd()), 0.5 * (1 + rnd())))
                    obj_list.insert(0, g.make_sphere(center, 0.2, m.make_metal(alb, 0.5 * rnd())))
                else:
                    obj_list.insert(0, g.make_sphere(center, 0.2, m.make_dielectric(1.5)))
    if big_spheres:
        obj_list.insert(0, g.make_sphere(v.vec3(0, 1, 0), 1, m.make_dielectric(1.5)))
        obj_list.insert(0, g.make_sphere(v.vec3(-4, 1, 0), 1, m.make_lambertian(t.constant_texture(v.vec3(0.4, 0.2, 0.1)))))
        obj_list.insert(0, g.make_sphere(v.vec3(4, 1, 0), 1, m.make_metal(t.constant_texture(v.vec3(0.7, 0.6, 0.5)), 0)))
    return obj_list


def cfg2_random_spheres(size_x=1200, size_y=800, seed=2):
    """configs[1]: Weekend final scene, static spheres, constant ground, ~488 primitives; book
    camera (13,2,3)->origin, vfov 20, aperture 0.1, focus 10 (not in the reference)."""
    c = cam.make_camera(v.vec3(13, 2, 3), v.vec3(0, 0, 0), v.vec3(0, 1, 0), 20, size_x / size_y, 0.1, 10, 0, 1)
    return g.make_scene(random_scene(seed, -11, 11), c, sky_color)


def cfg3_next_week(size_x=800, size_y=800, seed=3):
    """configs[2]: random-scene in its reference form (moving spheres, checker ground,
    main.scm:31-89) plus test-scene2's marble sphere, light sphere and light rect
    (main.scm:316-328) and a noise-texture sphere."""
    objs = random_scene(seed, -5, 10, moving=True, checker_ground=True, big_spheres=False)
    per_tex = t.marble_texture(1)
    light = m.make_diffuse_light(t.constant_texture(v.vec3(4, 4, 4)))
    extra = [
        g.make_sphere(v.vec3(0, 2, 0), 2, m.make_lambertian(per_tex)),
        g.make_sphere(v.vec3(0, 7, 0), 2, light),
        g.make_xy_rect(3, 5, 1, 3, -2, light),
        g.make_sphere(v.vec3(-4, 1, 0), 1, m.make_lambertian(t.noise_texture(4))),
        g.make_sphere(v.vec3(4, 1, 0), 1, m.make_metal(t.constant_texture(v.vec3(0.7, 0.6, 0.5)), 0)),
    ]
    c = cam.make_camera(v.vec3(13, 2, 3), v.vec3(0, 1, 0), v.vec3(0, 1, 0), 30, size_x / size_y, 0.1, 10, 0, 1)
    return g.make_scene(extra + objs, c, sky_color)


def _cornell_walls():
    red = m.make_lambertian(t.constant_texture(v.vec3(0.65, 0.05, 0.05)))
    white = m.make_lambertian(t.constant_texture(v.vec3(0.73, 0.73, 0.73)))
    green = m.make_lambertian(t.constant_texture(v.vec3(0.12, 0.45, 0.15)))
    light = m.make_diffuse_light(t.constant_texture(v.vec3(3, 3, 3)))
    walls = [
        g.flip_normals(g.make_yz_rect(0, 555, 0, 555, 555, green)),
        g.make_yz_rect(0, 555, 0, 555, 0, red),
        g.flip_normals(g.make_xz_rect(213, 343, 227, 332, 554, light)),
        g.flip_normals(g.make_xz_rect(0, 555, 0, 555, 555, white)),
        g.make_xz_rect(0, 555, 0, 555, 0, white),
        g.flip_normals(g.make_xy_rect(0, 555, 0, 555, 555, white)),
    ]
    return walls, red, white, green, light


def cfg4_cornell_box(size_x=1024, size_y=1024):
    """configs[3]: main.scm:330-351 cornell-box (sky-color background, light (3,3,3): Q14)."""
    walls, red, white, green, light = _cornell_walls()
    objs = walls + [
        g.translate(g.rotate_y(g.make_box(v.vec3(0, 0, 0), v.vec3(165, 165, 165), white), -18), v.vec3(130, 0, 65)),
        g.translate(g.rotate_y(g.make_box(v.vec3(0, 0, 0), v.vec3(165, 330, 165), white), 15), v.vec3(265, 0, 295)),
    ]
    return g.make_scene(objs, cornell_camera(size_x, size_y), sky_color)


def cornell_smoke(size_x=200, size_y=200):
    """main.scm:375-398 cornell-smoke: two constant media bounded by the rotated boxes, big light,
    black sky."""
    red = m.make_lambertian(t.constant_texture(v.vec3(0.65, 0.05, 0.05)))
    white = m.make_lambertian(t.constant_texture(v.vec3(0.73, 0.73, 0.73)))
    green = m.make_lambertian(t.constant_texture(v.vec3(0.12, 0.45, 0.15)))
    light = m.make_diffuse_light(t.constant_texture(v.vec3(3, 3, 3)))
    b1 = g.translate(g.rotate_y(g.make_box(v.vec3(0, 0, 0), v.vec3(165, 165, 165), white), -18), v.vec3(130, 0, 65))
    b2 = g.translate(g.rotate_y(g.make_box(v.vec3(0, 0, 0), v.vec3(165, 330, 165), white), 15), v.vec3(265, 0, 295))
    objs = [
        g.flip_normals(g.make_yz_rect(0, 555, 0, 555, 555, green)),
        g.make_yz_rect(0, 555, 0, 555, 0, red),
        g.flip_normals(g.make_xz_rect(113, 443, 127, 432, 554, light)),
        g.flip_normals(g.make_xz_rect(0, 555, 0, 555, 555, white)),
        g.make_xz_rect(0, 555, 0, 555, 0, white),
        g.flip_normals(g.make_xy_rect(0, 555, 0, 555, 555, white)),
        g.make_constant_medium(b1, 0.01, t.constant_texture(v.vec3(1, 1, 1))),
        g.make_constant_medium(b2, 0.01, t.constant_texture(v.vec3(0, 0, 0))),
    ]
    return g.make_scene(objs, cornell_camera(size_x, size_y), black)


def klein_scene(size_x=200, size_y=200):
    """main.scm:400-407 klein-scene."""
    white = m.make_lambertian(t.constant_texture(v.vec3(0.73, 0.73, 0.73)))
    red = m.make_lambertian(t.constant_texture(v.vec3(0.65, 0.05, 0.05)))
    objs = [g.make_sphere(v.vec3(0, -1003, -1), 1000, white), g.make_klein(v.vec3(0, 2, 0), red)]
    return g.make_scene(objs, default_camera(size_x, size_y), sky_color)


def cornell_klein(size_x=200, size_y=200):
    """main.scm:409-426 cornell-klein."""
    walls, red, white, green, light = _cornell_walls()
    blue = m.make_lambertian(t.constant_texture(v.vec3(0.05, 0.65, 0.65)))
    walls[2] = g.flip_normals(g.make_xz_rect(113, 443, 127, 432, 554, light))
    return g.make_scene(walls + [g.make_klein(v.vec3(250, 200, 280), blue)], cornell_camera(size_x, size_y), sky_color)


def test_bezier(size_x=200, size_y=200):
    """main.scm:237-277 test-bezier (3 curves + 6 spheres + checker ground)."""
    red = m.make_lambertian(t.constant_texture(v.vec3(0.65, 0.05, 0.05)))
    green = m.make_lambertian(t.constant_texture(v.vec3(0.12, 0.45, 0.15)))
    blue = m.make_lambertian(t.constant_texture(v.vec3(0.12, 0.15, 0.45)))
    objs = [
        g.make_sphere(v.vec3(0, -100.5, -1), 100, m.make_lambertian(_checker())),
        g.make_bvh_node([
            g.make_sphere(v.vec3(2, 0, 2), 0.5, red),
            g.make_sphere(v.vec3(-2, 0, -2), 0.5, green),
            g.make_sphere(v.vec3(-1, 0, -1), 0.1, blue),
            g.make_sphere(v.vec3(-0.8, 1, 1), 0.1, blue),
            g.make_sphere(v.vec3(0.8, -1, 1), 0.1, blue),
            g.make_sphere(v.vec3(1, 0, -1), 0.1, blue),
            b.make_bezier(v.vec3(-1, 0, -1), v.vec3(-0.8, 1, 1), v.vec3(0.8, -1, 1), v.vec3(1, 0, -1), 0.1, red),
            b.make_bezier(v.vec3(-1, 0, 1), v.vec3(-0.8, 1, -1), v.vec3(0.8, -1, -1), v.vec3(1, 0, 1), 0.1, red),
            b.make_bezier(v.vec3(-1, 0, 2), v.vec3(-0.8, 1, -2), v.vec3(0.8, -1, -2), v.vec3(1, 0, 2), 0.1, red),
        ], 0, 0),
    ]
    return g.make_scene(objs, default_camera(size_x, size_y), sky_color)


def cornell_bezier(size_x=200, size_y=200):
    """main.scm:353-373 cornell-bezier."""
    walls, red, white, green, light = _cornell_walls()
    objs = walls + [b.make_bezier(v.vec3(130, 0, 65), v.vec3(150, 0, 190), v.vec3(130, 0, 190), v.vec3(265, 0, 295), 10, red)]
    return g.make_scene(objs, cornell_camera(size_x, size_y), sky_color)


_K = 0.5522847498307936      # cubic Bezier circle constant

# Profile curves (radius, height) of the rotationally symmetric parts of the Utah teapot (rim,
# upper body, lower body, lid knob, lid, bottom), on exact quarter circles (_K).  The full Newell
# data set, handle and spout included, is `utah_teapot` below.
TEAPOT_PROFILES = [
    [(1.4, 2.25), (1.3375, 2.38125), (1.4375, 2.38125), (1.5, 2.25)],
    [(1.5, 2.25), (1.75, 1.725), (2.0, 1.2), (2.0, 0.75)],
    [(2.0, 0.75), (2.0, 0.3), (1.5, 0.075), (1.5, 0.0)],
    [(0.0, 3.15), (0.8, 3.15), (0.0, 2.7), (0.2, 2.55)],
    [(0.2, 2.55), (0.4, 2.4), (1.3, 2.4), (1.3, 2.25)],
    [(1.5, 0.0), (1.5, -0.075), (1.425, -0.15), (0.0, -0.15)],
]


def revolve_profile(profile, material, center=(0.0, 0.0, 0.0), scale=1.0):
    """Four bicubic patches = one cubic profile curve revolved about the y axis."""
    quads = [((1, 0), (1, _K), (_K, 1), (0, 1)), ((0, 1), (-_K, 1), (-1, _K), (-1, 0)),
             ((-1, 0), (-1, -_K), (-_K, -1), (0, -1)), ((0, -1), (_K, -1), (1, -_K), (1, 0))]
    out = []
    for q in quads:
        cp = [[(center[0] + scale * r * cx, center[1] + scale * y, center[2] + scale * r * cz) for (cx, cz) in q] for (r, y) in profile]
        out.append(b.make_bezier_patch(cp, material))
    return out


def teapot_patches(material, center=(0.0, 0.0, 0.0), scale=1.0):
    return [pt for prof in TEAPOT_PROFILES for pt in revolve_profile(prof, material, center, scale)]


# The Utah teapot (Newell 1975) in its compact public-domain form: 127 control points (teapot
# frame: z up, spout towards +x, the data cover the y <= 0 side) and 10 bicubic patches.  The six
# rotationally symmetric patches (rim, body x2, lid x2, bottom) are replicated into the four
# quadrants by mirroring x and y, handle and spout (two patches each) into the two halves by
# mirroring y: 6 x 4 + 4 x 2 = 32 patches.
TEAPOT_CP = [
    (0.2, 0, 2.7), (0.2, -0.112, 2.7), (0.112, -0.2, 2.7), (0, -0.2, 2.7),
    (1.3375, 0, 2.53125), (1.3375, -0.749, 2.53125), (0.749, -1.3375, 2.53125), (0, -1.3375, 2.53125),
    (1.4375, 0, 2.53125), (1.4375, -0.805, 2.53125), (0.805, -1.4375, 2.53125), (0, -1.4375, 2.53125),
    (1.5, 0, 2.4), (1.5, -0.84, 2.4), (0.84, -1.5, 2.4), (0, -1.5, 2.4),
    (1.75, 0, 1.875), (1.75, -0.98, 1.875), (0.98, -1.75, 1.875), (0, -1.75, 1.875),
    (2, 0, 1.35), (2, -1.12, 1.35), (1.12, -2, 1.35), (0, -2, 1.35),
    (2, 0, 0.9), (2, -1.12, 0.9), (1.12, -2, 0.9), (0, -2, 0.9),
    (-2, 0, 0.9),
    (2, 0, 0.45), (2, -1.12, 0.45), (1.12, -2, 0.45), (0, -2, 0.45),
    (1.5, 0, 0.225), (1.5, -0.84, 0.225), (0.84, -1.5, 0.225), (0, -1.5, 0.225),
    (1.5, 0, 0.15), (1.5, -0.84, 0.15), (0.84, -1.5, 0.15), (0, -1.5, 0.15),
    (-1.6, 0, 2.025), (-1.6, -0.3, 2.025), (-1.5, -0.3, 2.25), (-1.5, 0, 2.25),
    (-2.3, 0, 2.025), (-2.3, -0.3, 2.025), (-2.5, -0.3, 2.25), (-2.5, 0, 2.25),
    (-2.7, 0, 2.025), (-2.7, -0.3, 2.025), (-3, -0.3, 2.25), (-3, 0, 2.25),
    (-2.7, 0, 1.8), (-2.7, -0.3, 1.8), (-3, -0.3, 1.8), (-3, 0, 1.8),
    (-2.7, 0, 1.575), (-2.7, -0.3, 1.575), (-3, -0.3, 1.35), (-3, 0, 1.35),
    (-2.5, 0, 1.125), (-2.5, -0.3, 1.125), (-2.65, -0.3, 0.9375), (-2.65, 0, 0.9375),
    (-2, -0.3, 0.9), (-1.9, -0.3, 0.6), (-1.9, 0, 0.6),
    (1.7, 0, 1.425), (1.7, -0.66, 1.425), (1.7, -0.66, 0.6), (1.7, 0, 0.6),
    (2.6, 0, 1.425), (2.6, -0.66, 1.425), (3.1, -0.66, 0.825), (3.1, 0, 0.825),
    (2.3, 0, 2.1), (2.3, -0.25, 2.1), (2.4, -0.25, 2.025), (2.4, 0, 2.025),
    (2.7, 0, 2.4), (2.7, -0.25, 2.4), (3.3, -0.25, 2.4), (3.3, 0, 2.4),
    (2.8, 0, 2.475), (2.8, -0.25, 2.475), (3.525, -0.25, 2.49375), (3.525, 0, 2.49375),
    (2.9, 0, 2.475), (2.9, -0.15, 2.475), (3.45, -0.15, 2.5125), (3.45, 0, 2.5125),
    (2.8, 0, 2.4), (2.8, -0.15, 2.4), (3.2, -0.15, 2.4), (3.2, 0, 2.4),
    (0, 0, 3.15), (0.8, 0, 3.15), (0.8, -0.45, 3.15), (0.45, -0.8, 3.15), (0, -0.8, 3.15),
    (0, 0, 2.85),
    (1.4, 0, 2.4), (1.4, -0.784, 2.4), (0.784, -1.4, 2.4), (0, -1.4, 2.4),
    (0.4, 0, 2.55), (0.4, -0.224, 2.55), (0.224, -0.4, 2.55), (0, -0.4, 2.55),
    (1.3, 0, 2.55), (1.3, -0.728, 2.55), (0.728, -1.3, 2.55), (0, -1.3, 2.55),
    (1.3, 0, 2.4), (1.3, -0.728, 2.4), (0.728, -1.3, 2.4), (0, -1.3, 2.4),
    (0, 0, 0), (1.425, -0.798, 0), (1.5, 0, 0.075), (1.425, 0, 0), (0.798, -1.425, 0),
    (0, -1.5, 0.075), (0, -1.425, 0), (1.5, -0.84, 0.075), (0.84, -1.5, 0.075),
]
TEAPOT_PATCHES = [   # (part, mirrored copies, 16 control-point indices, row = u, column = v)
    ("rim", 4, (102, 103, 104, 105, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)),
    ("body", 4, (12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27)),
    ("body", 4, (24, 25, 26, 27, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40)),
    ("lid", 4, (96, 96, 96, 96, 97, 98, 99, 100, 101, 101, 101, 101, 0, 1, 2, 3)),
    ("lid", 4, (0, 1, 2, 3, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 116, 117)),
    ("bottom", 4, (118, 118, 118, 118, 124, 122, 119, 121, 123, 126, 125, 120, 40, 39, 38, 37)),
    ("handle", 2, (41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56)),
    ("handle", 2, (53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 28, 65, 66, 67)),
    ("spout", 2, (68, 69, 70, 71, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83)),
    ("spout", 2, (80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 93, 94, 95)),
]


def utah_teapot(material, center=(0.0, 0.0, 0.0), scale=1.0, parts=None):
    """All 32 bicubic patches of the Utah teapot as `make-bezier-patch` objects.  Teapot frame
    (x, y, z-up) -> world (x, z, -y), base of the pot at `center`.  `parts` restricts the result to
    some of rim / body / lid / bottom / handle / spout."""
    out = []
    for part, copies, idx in TEAPOT_PATCHES:
        if parts is not None and part not in parts:
            continue
        for sx, sy in ((1, 1), (1, -1), (-1, 1), (-1, -1))[:copies]:
            cp = [[(center[0] + scale * sx * TEAPOT_CP[idx[4 * i + j]][0], center[1] + scale * TEAPOT_CP[idx[4 * i + j]][2],
                    center[2] - scale * sy * TEAPOT_CP[idx[4 * i + j]][1]) for j in range(4)] for i in range(4)]
            out.append(b.make_bezier_patch(cp, material))
    return out


def teapot_scene(size_x=3840, size_y=2160):
    """configs[4] with the complete teapot: one full 32-patch Utah teapot (lambertian) flanked by a
    fuzzy-metal and a red lambertian copy at 2/3 scale on a checker ground, under the gradient sky.
    Patches are a north-star extension (absent upstream): parity vs the oracle's identical
    definition only."""
    white = m.make_lambertian(t.constant_texture(v.vec3(0.73, 0.73, 0.73)))
    red = m.make_lambertian(t.constant_texture(v.vec3(0.65, 0.05, 0.05)))
    gold = m.make_metal(t.constant_texture(v.vec3(0.8, 0.6, 0.2)), 0.1)
    objs = [g.make_sphere(v.vec3(0, -1000, 0), 1000, m.make_lambertian(_checker()))]
    objs += utah_teapot(white, (0.0, 0.0, 0.0), 1.0)
    objs += utah_teapot(gold, (-4.6, 0.0, -1.0), 0.66)
    objs += utah_teapot(red, (4.4, 0.0, 0.0), 0.66)
    c = cam.make_camera(v.vec3(0, 5, 11), v.vec3(0, 1.3, 0), v.vec3(0, 1, 0), 35, size_x / size_y, 0, 1, 0, 1)
    return g.make_scene(objs, c, sky_color)


def cfg5_patches(size_x=3840, size_y=2160):
    """configs[4]: bicubic Bezier-patch scene — the rotationally symmetric 24 patches of the Utah
    teapot (body, lid, bottom) in three materials' worth of copies on a checker ground, plus the
    reference's three Bezier curves (main.scm:259-273).  Patches are a north-star extension
    (absent upstream): parity vs the oracle's identical definition only."""
    white = m.make_lambertian(t.constant_texture(v.vec3(0.73, 0.73, 0.73)))
    red = m.make_lambertian(t.constant_texture(v.vec3(0.65, 0.05, 0.05)))
    gold = m.make_metal(t.constant_texture(v.vec3(0.8, 0.6, 0.2)), 0.1)
    objs = [g.make_sphere(v.vec3(0, -1000, 0), 1000, m.make_lambertian(_checker()))]
    objs += teapot_patches(white, (0.0, 0.15, 0.0), 1.0)
    objs += teapot_patches(gold, (-4.5, 0.1, -1.0), 0.66)
    objs += teapot_patches(red, (4.2, 0.1, 1.0), 0.66)
    objs += [b.make_bezier(v.vec3(-3, 0.3, 3), v.vec3(-1, 2.0, 2.5), v.vec3(1, -1.0, 3.5), v.vec3(3, 0.3, 3), 0.1, red),
             g.make_sphere(v.vec3(0, 4.2, 0), 0.5, m.make_dielectric(1.5))]
    c = cam.make_camera(v.vec3(0, 5, 11), v.vec3(0, 1.3, 0), v.vec3(0, 1, 0), 35, size_x / size_y, 0, 1, 0, 1)
    return g.make_scene(objs, c, sky_color)


def test_scene2(size_x=200, size_y=200):
    """main.scm:316-328 test-scene2 (marble + lights, black sky)."""
    per_tex = t.marble_texture(1)
    objs = [
        g.make_sphere(v.vec3(0, -1000, -1), 1000, m.make_lambertian(per_tex)),
        g.make_sphere(v.vec3(0, 2, 0), 2, m.make_lambertian(per_tex)),
        g.make_sphere(v.vec3(0, 7, 0), 2, m.make_diffuse_light(t.constant_texture(v.vec3(4, 4, 4)))),
        g.make_xy_rect(3, 5, 1, 3, -2, m.make_diffuse_light(t.constant_texture(v.vec3(4, 4, 4)))),
    ]
    return g.make_scene(objs, default_camera(size_x, size_y), black)


def image_scene(size_x=200, size_y=200, nx=16, ny=8, seed=11):
    """image-texture (texture.scm:36-50; never instantiated upstream) on an emissive screen, a
    lambertian ground (lambertian passes u = v = 0: one texel) and an emissive unit sphere (Q5 uv)."""
    texels = np.random.RandomState(seed).randint(0, 256, size=nx * ny * 3)
    img = t.image_texture(texels, nx, ny)
    objs = [
        g.make_sphere(v.vec3(0, -1000, -1), 1000, m.make_lambertian(img)),
        g.make_sphere(v.vec3(0, 2, 0), 2, m.make_lambertian(t.constant_texture(v.vec3(0.7, 0.7, 0.7)))),
        g.make_xy_rect(-5, 5, 0, 6, -3, m.make_diffuse_light(img)),
        g.make_sphere(v.vec3(3.2, 0.5, 1.5), 0.5, m.make_diffuse_light(img)),
    ]
    return g.make_scene(objs, default_camera(size_x, size_y), black)


def sphere_cloud(n, size_x=64, size_y=64, seed=13):
    """n small spheres in a 40^3 cube over a ground sphere: sizes the shared-memory staging
    cannot hold (trees read through L1/L2; >= 65536 nodes also drops the 16-bit far-child cache)."""
    rs = np.random.RandomState(seed)
    c = rs.uniform(-20, 20, (n, 3)); c[:, 1] += 20.5
    r = rs.uniform(0.15, 0.4, n)       # not smaller: the fp32 normal (p - c)/r loses |c|/r ulps
    mats = [m.make_lambertian(t.constant_texture(v.vec3(0.2 + 0.1 * k, 0.5, 0.8 - 0.1 * k))) for k in range(6)]
    mats += [m.make_metal(t.constant_texture(v.vec3(0.8, 0.8, 0.8)), 0.1), m.make_dielectric(1.5)]
    objs = [g.make_sphere(v.vec3(0, -1000, 0), 1000, m.make_lambertian(_checker()))]
    objs += [g.make_sphere(v.vec3(*c[i]), float(r[i]), mats[i % len(mats)]) for i in range(n)]
    camera = cam.make_camera(v.vec3(0, 20, 70), v.vec3(0, 20, 0), v.vec3(0, 1, 0), 40, size_x / size_y, 0.0, 70.0, 0.0, 1.0)
    return g.make_scene(objs, camera, sky_color)


def line_upped_spheres(nx=10, ny=10, seed=7):
    """main.scm:177-191 + 204-213 test-scene-non-bvh: the reference's own (commented) benchmark."""
    rnd = np.random.RandomState(seed).random_sample
    objs = []
    for x in range(nx):
        for y in range(ny):
            objs.insert(0, g.make_sphere(v.vec3(x, 0, y), 0.5, m.make_lambertian(t.constant_texture(v.vec3(rnd(), rnd(), rnd())))))
    return objs


def test_scene_bvh(size_x=200, size_y=200):
    """main.scm:215-235 test-scene-bvh / -bvh-sah (make-bvh-* are grouping hints here)."""
    objs = [g.make_sphere(v.vec3(0, -100.5, -1), 100, m.make_lambertian(_checker())),
            g.make_bvh_with_sah(line_upped_spheres(10, 10), 0, 0)]
    return g.make_scene(objs, default_camera(size_x, size_y), sky_color)


CONFIGS = {
    "cfg1": dict(scene=cfg1_weekend, width=200, height=100, spp=16, max_depth=50, seed=1),
    "cfg2": dict(scene=cfg2_random_spheres, width=1200, height=800, spp=500, max_depth=50, seed=2),
    "cfg3": dict(scene=cfg3_next_week, width=800, height=800, spp=1000, max_depth=50, seed=3),
    "cfg4": dict(scene=cfg4_cornell_box, width=1024, height=1024, spp=4096, max_depth=50, seed=4),
    "cfg5": dict(scene=cfg5_patches, width=3840, height=2160, spp=1024, max_depth=50, seed=5),
    "cfg5_teapot": dict(scene=teapot_scene, width=3840, height=2160, spp=1024, max_depth=50, seed=5),
    "cfg5_curves": dict(scene=test_bezier, width=3840, height=2160, spp=1024, max_depth=50, seed=5),
}
