"""Object specs shared by the reference-golden generator (tests/golden/make_reference_golden.py, which
builds them with the REFERENCE's constructors through oracle/minischeme.py) and by the tests (which
build the same objects with the host mirror of the constructor API).  A spec is a nested list:
["sphere", c, r] ["moving-sphere", c0, c1, t0, t1, r] ["xy-rect"|"xz-rect"|"yz-rect", a0, a1, b0, b1, k]
["flip", obj] ["box", p0, p1] ["translate", obj, offset] ["rotate-y", obj, degrees] ["bezier", a, b, c, d, width]
["medium", boundary-obj, density] (phase texture = constant (1,1,1)) ["klein", center]."""
from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, bezier as bz, vec as v


def build_host(spec, mat=None):
    mat = mat or m.make_lambertian(t.constant_texture(v.vec3(0.5, 0.5, 0.5)))
    k = spec[0]
    if k == "sphere":
        return g.make_sphere(v.vec3(*spec[1]), spec[2], mat)
    if k == "moving-sphere":
        return g.make_moving_sphere(v.vec3(*spec[1]), v.vec3(*spec[2]), spec[3], spec[4], spec[5], mat)
    if k in ("xy-rect", "xz-rect", "yz-rect"):
        return {"xy-rect": g.make_xy_rect, "xz-rect": g.make_xz_rect, "yz-rect": g.make_yz_rect}[k](*spec[1:6], mat)
    if k == "flip":
        return g.flip_normals(build_host(spec[1], mat))
    if k == "box":
        return g.make_box(v.vec3(*spec[1]), v.vec3(*spec[2]), mat)
    if k == "translate":
        return g.translate(build_host(spec[1], mat), v.vec3(*spec[2]))
    if k == "rotate-y":
        return g.rotate_y(build_host(spec[1], mat), spec[2])
    if k == "bezier":
        return bz.make_bezier(v.vec3(*spec[1]), v.vec3(*spec[2]), v.vec3(*spec[3]), v.vec3(*spec[4]), spec[5], mat)
    if k == "medium":
        return g.make_constant_medium(build_host(spec[1], mat), spec[2], t.constant_texture(v.vec3(1, 1, 1)))
    if k == "klein":
        return g.make_klein(v.vec3(*spec[1]), mat)
    raise ValueError(k)


def ref_nextweek_scene(size_x=200, size_y=200):
    """Mirror of NEXTWEEK_SCENE in tests/golden/make_reference_golden.py."""
    from scheme_raytrace_b200.host import scenes, camera as cam
    objs = [g.make_sphere(v.vec3(0, -1000, 0), 1000, m.make_lambertian(t.checker_texture(t.constant_texture(v.vec3(0.2, 0.3, 0.1)), t.constant_texture(v.vec3(0.9, 0.9, 0.9))))),
            g.make_moving_sphere(v.vec3(0, 1, 0), v.vec3(0, 1.75, 0), 0, 1, 0.75, m.make_lambertian(t.constant_texture(v.vec3(0.7, 0.3, 0.1)))),
            g.make_sphere(v.vec3(-2.5, 1, 0), 1, m.make_lambertian(t.noise_texture(4))),
            g.make_sphere(v.vec3(2.5, 1, 0), 1, m.make_lambertian(t.marble_texture(1))),
            g.flip_normals(g.make_xz_rect(-1, 1, -1, 1, 3.5, m.make_diffuse_light(t.constant_texture(v.vec3(4, 4, 4)))))]
    c = cam.make_camera(v.vec3(13, 2, 3), v.vec3(0, 0, 0), v.vec3(0, 1, 0), 20, 1, 0.5, 10, 0, 1)
    return g.make_scene(objs, c, scenes.sky_color)


def host_scene(name, size_x=200, size_y=200):
    """The host mirror of a scene that the generator takes from the reference's main.scm."""
    from scheme_raytrace_b200.host import scenes
    non_bvh = lambda sx, sy: g.make_scene([g.make_sphere(v.vec3(0, -100.5, -1), 100, m.make_lambertian(scenes._checker()))] + scenes.line_upped_spheres(10, 10),
                                          scenes.default_camera(sx, sy), scenes.sky_color)
    return {"cornell-box": scenes.cfg4_cornell_box, "test-scene2": scenes.test_scene2, "cornell-bezier": scenes.cornell_bezier,
            "cornell-smoke": scenes.cornell_smoke, "klein-scene": scenes.klein_scene, "cornell-klein": scenes.cornell_klein,
            "test-bezier": scenes.test_bezier, "test-scene-bvh": scenes.test_scene_bvh, "test-scene-bvh-sah": scenes.test_scene_bvh,
            "test-scene-non-bvh": non_bvh, "ref-nextweek-scene": ref_nextweek_scene,
            # main.scm:31-89 in its reference form (grid [-5, 10), moving lambertians, checker ground); 41 = RANDOM_SCENE_SEED of the generator
            "random-scene": lambda sx, sy: g.make_scene(scenes.random_scene(41, -5, 10, moving=True, checker_ground=True), scenes.default_camera(sx, sy), scenes.sky_color),
            "test-scene": lambda sx, sy: g.make_scene(scenes.test_scene_objects(), scenes.default_camera(sx, sy), scenes.black)}[name](size_x, size_y)
