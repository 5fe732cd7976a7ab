// srt_render — command-line host over the C-ABI (include/srt.h): reads a flat scene file
// (scheme_raytrace_b200/host/scenefile.py documents the format; scheme/srt-scene.scm writes it from
// Gauche), commits it, renders and writes the PPM exactly like save-as-ppm (main.scm:439-450).
//
//   srt_render scene.srt [--width W] [--height H] [--spp S] [--depth D] [--seed X] [--quirks Q]
//                        [--estimator E] [--device N] [--gpus N] [--passes P] [--out test.ppm]
//
//   --gpus N    render on N GPUs of this box from this one process (0 = all): srt_init_multi +
//               srt_render_multi, the frame's samples sharded by sample range (bit-identical image)
//   --passes P  progressive mode (main.scm:452-469): P passes of spp/P samples through
//               srt_progressive_step, the running sum resident on the device; the PPM is the last frame
//
// This is the route by which an unmodified Gauche scene script drives the GPU path without a
// compiled Gauche extension: (use srt-scene) ... (srt:write-scene scene "scene.srt"), then this.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include "../../include/srt.h"

static bool expect(std::istream& in, const char* word) { std::string w; in >> w; return (bool)in && w == word; }
// a table count read from the file: must be a sane non-negative number before it sizes a vector
static bool read_count(std::istream& in, const char* word, int& n, int max_n = 1 << 26) {
  if (!expect(in, word) || !(in >> n)) { std::fprintf(stderr, "srt_render: expected section '%s <count>'\n", word); return false; }
  if (n < 0 || n > max_n) { std::fprintf(stderr, "srt_render: section '%s' has a bad count %d\n", word, n); return false; }
  return true;
}
#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { std::fprintf(stderr, "srt_render: %s failed (%d): %s\n", #call, rc_, srt_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: srt_render scene.srt [--width W --height H --spp S --depth D --seed X --quirks Q --estimator E --device N --out file.ppm]\n"); return 2; }
  int width = 200, height = 200, spp = 1, depth = 100, quirks = SRT_QUIRKS_REFERENCE, estimator = 0, device = 0, gpus = 1, passes = 0;   // main.scm:26,126-127 defaults
  unsigned seed = 1; std::string out = "test.ppm";
  for (int i = 2; i + 1 < argc; i += 2) {
    std::string k = argv[i]; const char* v = argv[i + 1];
    if (k == "--width") width = std::atoi(v); else if (k == "--height") height = std::atoi(v); else if (k == "--spp") spp = std::atoi(v);
    else if (k == "--depth") depth = std::atoi(v); else if (k == "--seed") seed = (unsigned)std::strtoul(v, nullptr, 10);
    else if (k == "--quirks") quirks = std::atoi(v); else if (k == "--estimator") estimator = std::atoi(v);
    else if (k == "--device") device = std::atoi(v); else if (k == "--out") out = v;
    else if (k == "--gpus") gpus = std::atoi(v); else if (k == "--passes") passes = std::atoi(v);
    else { std::fprintf(stderr, "srt_render: unknown option %s\n", k.c_str()); return 2; }
  }
  if (width <= 0 || height <= 0 || spp <= 0 || depth < 0 || gpus < 0 || passes < 0 || (long long)width * height > (1ll << 30)) {
    std::fprintf(stderr, "srt_render: bad --width / --height / --spp / --depth / --gpus / --passes\n"); return 2;
  }
  std::ifstream in(argv[1]);
  if (!in) { std::fprintf(stderr, "srt_render: cannot open %s\n", argv[1]); return 1; }
  int version = 0, sky = 0, n = 0;
  if (!expect(in, "srt-scene") || !(in >> version) || version != 1) { std::fprintf(stderr, "srt_render: not an srt-scene 1 file\n"); return 1; }
  if (!expect(in, "sky") || !(in >> sky)) return 1;
  SrtCamera cam; float* cf = (float*)&cam;
  if (!expect(in, "camera")) return 1;
  for (int i = 0; i < 24; ++i) in >> cf[i];
  std::vector<float> ranvec(768); std::vector<int32_t> perm(768);
  if (!expect(in, "perlin-ranvec")) return 1;
  for (auto& x : ranvec) in >> x;
  if (!expect(in, "perlin-perm")) return 1;
  for (auto& x : perm) in >> x;
  if (!read_count(in, "textures", n)) return 1;
  std::vector<SrtTexture> tex(n);
  for (auto& t : tex) { std::memset(&t, 0, sizeof(t)); in >> t.kind >> t.even >> t.odd >> t.scale >> t.rgb[0] >> t.rgb[1] >> t.rgb[2]; }
  if (!read_count(in, "materials", n)) return 1;
  std::vector<SrtMaterial> mat(n);
  for (auto& m : mat) { std::memset(&m, 0, sizeof(m)); in >> m.kind >> m.tex >> m.param; }
  if (!read_count(in, "xforms", n)) return 1;
  std::vector<SrtXform> xf(n);
  for (auto& x : xf) in >> x.sin_t >> x.cos_t >> x.off[0] >> x.off[1] >> x.off[2];
  if (!read_count(in, "patches", n, 1 << 22)) return 1;
  std::vector<float> patches(48 * (size_t)n); int npatch = n;
  for (auto& x : patches) in >> x;
  if (!read_count(in, "prims", n)) return 1;
  std::vector<SrtPrim> prims(n);
  for (auto& p : prims) { in >> p.type >> p.flags >> p.material >> p.xform; for (int i = 0; i < 16; ++i) in >> p.p[i]; }
  if (!read_count(in, "lights", n)) return 1;
  std::vector<int32_t> lights(n);
  for (auto& l : lights) in >> l;
  if (!in) { std::fprintf(stderr, "srt_render: truncated scene file\n"); return 1; }
  std::vector<uint8_t> texels; std::vector<int32_t> dims;          // optional trailing section
  if (expect(in, "images") && (in >> n)) {
    if (n < 0 || n > (1 << 16)) { std::fprintf(stderr, "srt_render: section 'images' has a bad count %d\n", n); return 1; }
    for (int i = 0; i < n; ++i) {
      int nx = 0, ny = 0; in >> nx >> ny;
      if (!in || nx < 1 || ny < 1 || (long long)nx * ny > (1ll << 28)) { std::fprintf(stderr, "srt_render: bad image header\n"); return 1; }
      dims.push_back(nx); dims.push_back(ny); dims.push_back((int32_t)texels.size());
      for (size_t k = 0; k < 3 * (size_t)nx * ny; ++k) { int v = 0; in >> v; texels.push_back((uint8_t)v); }
    }
    if (!in) { std::fprintf(stderr, "srt_render: truncated images section\n"); return 1; }
  }

  if (gpus != 1) CHECK(srt_init_multi(gpus)); else CHECK(srt_init(device));
  SrtScene* sc = srt_scene_create();
  if (!sc) { std::fprintf(stderr, "srt_render: %s\n", srt_last_error()); return 1; }
  CHECK(srt_scene_set_prims(sc, prims.data(), (int)prims.size()));
  CHECK(srt_scene_set_xforms(sc, xf.data(), (int)xf.size()));
  CHECK(srt_scene_set_patches(sc, patches.data(), npatch));
  CHECK(srt_scene_set_materials(sc, mat.data(), (int)mat.size()));
  CHECK(srt_scene_set_textures(sc, tex.data(), (int)tex.size()));
  CHECK(srt_scene_set_images(sc, texels.data(), dims.data(), (int)dims.size() / 3));
  CHECK(srt_scene_set_perlin(sc, ranvec.data(), perm.data(), perm.data() + 256, perm.data() + 512));
  CHECK(srt_scene_set_camera(sc, &cam));
  CHECK(srt_scene_set_lights(sc, lights.data(), (int)lights.size()));
  CHECK(srt_scene_commit(sc));
  SrtRenderParams p; std::memset(&p, 0, sizeof(p));
  p.width = width; p.height = height; p.spp_begin = 0; p.spp_end = spp; p.max_depth = depth; p.sky = sky; p.seed = seed;
  p.quirks = quirks; p.t_min = 0.001f; p.estimator = estimator;
  std::vector<float> rgb((size_t)width * height * 3, 0.0f);
  std::vector<uint8_t> img(rgb.size());
  SrtStats st; std::memset(&st, 0, sizeof(st));
  if (passes > 0) {                                          // the viewer's loop without the window (main.scm:533-544)
    unsigned long long rays = 0; float ms = 0.f; int launches = 0;
    for (int k = 0; k < passes; ++k) {
      p.spp_begin = (int)((long long)k * spp / passes); p.spp_end = (int)((long long)(k + 1) * spp / passes);
      if (p.spp_end == p.spp_begin) continue;
      CHECK(srt_progressive_step(sc, &p, img.data(), &st));
      rays += st.rays; ms += st.ms_total; launches += st.kernel_launches;
    }
    st.rays = rays; st.ms_total = ms; st.kernel_launches = launches;
  } else if (gpus != 1) {
    p.reserved[2] = 1;
    CHECK(srt_render_multi(sc, &p, rgb.data(), img.data(), &st));   // (trace-all scene 1..spp) on every GPU
  } else {
    p.reserved[2] = 1;
    CHECK(srt_render_host(sc, &p, rgb.data(), &st));          // (trace-all scene 1..spp)
    CHECK(srt_resolve_host(rgb.data(), width, height, spp, img.data()));
  }
  CHECK(srt_save_ppm(out.c_str(), img.data(), width, height));
  std::fprintf(stderr, "srt_render: %llu rays in %.3f ms (%.1f Mrays/s) on %d GPU(s), %d kernel launches, LBVH %d nodes depth %d -> %s\n",
               (unsigned long long)st.rays, st.ms_total, st.rays / (st.ms_total * 1e3), gpus != 1 ? srt_multi_device_count() : 1, st.kernel_launches, st.bvh_nodes, st.bvh_depth, out.c_str());
  srt_scene_destroy(sc); srt_shutdown();
  return 0;
}
