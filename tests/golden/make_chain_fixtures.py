"""Build-container only: freezes the FIRST half of the drop-in chain as fixtures.

    reference scene text (/root/reference/main.scm, unmodified)
      -> evaluated by oracle/minischeme.py on top of the repo's Gauche host modules (scheme_raytrace_b200/scheme/*.scm)
      -> (srt:write-scene scene "chain_<scene>.srt")          [this script stops here: tests/golden/chain_<scene>.srt]
      -> cli/srt_render chain_<scene>.srt -> test.ppm          [tests/test_gpu_chain.py, on the GPU box]
      -> compared with the PPM the REFERENCE wrote for the same frame (tests/golden/ref_color.json, save-as-ppm main.scm:439-450)

/root/reference does not exist on the GPU box, so the scene files (flattened tables = OUTPUT of running the reference's scene
definitions through the repo's Scheme host; no reference source text) are committed; tests/test_scheme_host.py re-creates them from
the reference whenever it is present and requires byte equality, so they cannot drift.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
CHAIN_SCENES = ["cornell-box", "test-scene2", "cornell-smoke", "test-bezier"]      # the main.scm scenes of ref_color.json


def write_chain_files(out_dir, names=CHAIN_SCENES):
    from tests.test_scheme_host import interpreter, in_big_stack, write_scene, HOST, REFERENCE
    from oracle.minischeme import Sym, read_all
    want = {"+black+", "+white+", "sky-color", "black", "*size-x*", "*size-y*", "*cornell-camera*", "*camera*"} | set(names)

    def work():
        it = interpreter([HOST, REFERENCE])
        it.require("srt-scene")
        main = None
        for form in read_all(open(os.path.join(REFERENCE, "main.scm")).read()):
            if not isinstance(form, list) or not form:
                continue
            if form[0] == "define-module":
                it.eval(form, it.user)
                main = it.modules["main"]
            elif form[0] in ("define", "define-inline") and main is not None:
                name = form[1][0] if isinstance(form[1], list) else form[1]
                if name in want:
                    it.eval(form, main)
        sky = {id(main.lookup(Sym("sky-color"))): "sky-color", id(main.lookup(Sym("black"))): "black"}
        for name in names:
            scene = main.lookup(Sym(name))
            scene[4] = it.modules["srt-scene"].lookup(Sym(sky[id(scene[4])]))
            write_scene(it, scene, os.path.join(out_dir, f"chain_{name}.srt"))
    in_big_stack(work)


if __name__ == "__main__":
    write_chain_files(HERE)
    for n in CHAIN_SCENES:
        print("wrote", f"chain_{n}.srt", os.path.getsize(os.path.join(HERE, f"chain_{n}.srt")) // 1024, "KB")
