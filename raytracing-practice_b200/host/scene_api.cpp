// scene_api.cpp -> librtb200_scenes.so : builds a named scene with the host API (scenes.hpp)
// and hands out its flattened description.  Pure host code, no CUDA: it exists so that tests,
// bench.py and tools can feed the SAME POD scene to rt_upload_scene (GPU) and to the oracle.
#define RTB200_NO_RENDER_IMPL 1
#include "rtb200_host.hpp"
#include "scenes.hpp"

struct rth_scene {
  std::unique_ptr<rtb200::scene_builder> flat;
  rt_scene_desc desc;
  rt_camera_desc cam;
};

extern "C" {

int rth_scene_count() {
  int n = 0;
  rtb200_scenes::scene_table(&n);
  return n;
}
const char* rth_scene_name(int i) {
  int n = 0;
  const rtb200_scenes::scene_entry* t = rtb200_scenes::scene_table(&n);
  return (i >= 0 && i < n) ? t[i].name : nullptr;
}

// rand_seed: the reference never calls srand, i.e. behaves like srand(1) in a fresh process.
// Pass 1 to reproduce "one scene per fresh process"; < 0 leaves the global stream untouched.
rth_scene* rth_scene_build(const char* name, long rand_seed) {
  if (rand_seed >= 0) std::srand(unsigned(rand_seed));
  rtb200_scenes::scene_setup s;
  if (!rtb200_scenes::build_scene(name, s)) return nullptr;
  rth_scene* out = new rth_scene;
  out->flat = rtb200::flatten_world(s.world);
  out->desc = out->flat->desc();
  out->cam = s.cam.desc();
  return out;
}
void rth_scene_free(rth_scene* s) { delete s; }
const rt_scene_desc* rth_scene_desc(rth_scene* s) { return &s->desc; }
rt_camera_desc* rth_scene_camera(rth_scene* s) { return &s->cam; }

// Decode an image file the way rtw_image does (JPEG / P6 -> RGB8 -> gamma-2.2 -> bytes).
// Returns 0 on success; *rgb is malloc'd (free with rth_free).
int rth_load_texture(const char* path, int linearise, uint8_t** rgb, int* w, int* h) {
  std::vector<uint8_t> raw;
  if (linearise) {
    rtw_image im;
    if (!im.load(path)) return -1;
    raw = im.bytes();
    *w = im.width(), *h = im.height();
  } else if (!rtb200::load_image_rgb8(path, raw, *w, *h)) {
    return -1;
  }
  *rgb = static_cast<uint8_t*>(std::malloc(raw.size()));
  std::memcpy(*rgb, raw.data(), raw.size());
  return 0;
}
void rth_free(void* p) { std::free(p); }
}
