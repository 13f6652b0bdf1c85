// ref_harness.cpp — drives the UNMODIFIED reference (headers + main.cpp under
// /root/reference/src, compiled where they lie; nothing is copied) as oracle/_ref/ref_harness.
// TEST INFRASTRUCTURE ONLY.  Recipe: oracle/Makefile.  Techniques:
//   * all std headers first, then `#define private public` around the reference includes, to
//     reach camera::initialize/get_ray/ray_color and the containers' children;
//   * `#define main ref_main` + `#define render(...)` around `#include "main.cpp"` so the
//     shipped scene functions run unchanged but hand (cam, world) to this harness.
// Scenes: "shipped:<name>" = the reference's own src/main.cpp function; "<name>" = this
// repo's scenes.hpp builder compiled against the REFERENCE classes (+ ref_ext.hpp).
//
//   ref_harness ppm     <scene> <out.ppm>  [--width W] [--spp S] [--depth D] [--seed K]
//   ref_harness primary <scene> <out.bin>  [--width W]
//   ref_harness linear  <scene> <out.bin>  [--width W] [--spp S] [--depth D] [--seed K]
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "stb_shim/stb_image.h"

#define private public
#include "common/rtweekend.hpp"
#include "accelerator/bvh_node.hpp"
#include "core/camera.hpp"
#include "core/material.hpp"
#include "core/texture.hpp"
#include "hittable/hittable.hpp"
#include "hittable/hittable_list.hpp"
#include "hittable/sphere.hpp"
#include "hittable/quad.hpp"
#include "ref_ext.hpp"
#undef private

#include "../raytracing-practice_b200/host/scenes.hpp"

namespace {

struct options {
  std::string mode, scene, out;
  int width = -1, spp = -1, depth = -1;
  long seed = -1;
} g_opt;

struct counting_world : hittable {
  const hittable& inner;
  mutable unsigned long long rays = 0;
  explicit counting_world(const hittable& w) : inner(w) {}
  bool hit(const ray& r, interval t, hit_record& rec) const override {
    rays++;
    return inner.hit(r, t, rec);
  }
  aabb bounding_box() const override { return inner.bounding_box(); }
};

// ---- DFS leaf numbering + replay of the containers' traversal -----
std::map<const hittable*, int> g_leaf_id;
bool g_has_medium = false;

void number_leaves(const hittable* h) {
  if (auto l = dynamic_cast<const hittable_list*>(h)) {
    for (auto& o : l->objects) number_leaves(o.get());
  } else if (auto b = dynamic_cast<const bvh_node*>(h)) {
    number_leaves(b->left.get());
    if (b->right.get() != b->left.get()) number_leaves(b->right.get());
  } else if (auto t = dynamic_cast<const translate*>(h)) {
    number_leaves(t->object.get());
  } else if (auto r = dynamic_cast<const rotate_y*>(h)) {
    number_leaves(r->object.get());
  } else if (auto m = dynamic_cast<const constant_medium*>(h)) {
    g_has_medium = true;
    number_leaves(m->boundary.get());
  } else if (!g_leaf_id.count(h)) {
    int id = int(g_leaf_id.size());
    g_leaf_id[h] = id;
  }
}

// Same control flow as hittable_list::hit / bvh_node::hit / translate::hit / rotate_y::hit,
// leaves call the real hit(); media are transparent for the primary pass.
bool replay(const hittable* h, const ray& r, interval ray_t, hit_record& rec, int& id) {
  if (auto l = dynamic_cast<const hittable_list*>(h)) {
    hit_record tmp;
    int tmp_id = -1;
    bool any = false;
    auto closest = ray_t.max;
    for (auto& o : l->objects)
      if (replay(o.get(), r, interval(ray_t.min, closest), tmp, tmp_id)) {
        any = true;
        closest = tmp.t;
        rec = tmp;
        id = tmp_id;
      }
    return any;
  }
  if (auto b = dynamic_cast<const bvh_node*>(h)) {
    if (!b->bbox.hit(r, ray_t)) return false;
    bool hl = replay(b->left.get(), r, ray_t, rec, id);
    bool hr = replay(b->right.get(), r, interval(ray_t.min, hl ? rec.t : ray_t.max), rec, id);
    return hl || hr;
  }
  if (auto t = dynamic_cast<const translate*>(h)) {
    ray moved(r.origin() - t->offset, r.direction(), r.time());
    if (!replay(t->object.get(), moved, ray_t, rec, id)) return false;
    rec.p += t->offset;
    return true;
  }
  if (auto ry = dynamic_cast<const rotate_y*>(h)) {
    const double c = ry->cos_theta, s = ry->sin_theta;
    point3 o((c * r.origin().x()) - (s * r.origin().z()), r.origin().y(), (s * r.origin().x()) + (c * r.origin().z()));
    vec3 d((c * r.direction().x()) - (s * r.direction().z()), r.direction().y(),
           (s * r.direction().x()) + (c * r.direction().z()));
    ray rot(o, d, r.time());
    if (!replay(ry->object.get(), rot, ray_t, rec, id)) return false;
    rec.p = point3((c * rec.p.x()) + (s * rec.p.z()), rec.p.y(), (-s * rec.p.x()) + (c * rec.p.z()));
    rec.normal = vec3((c * rec.normal.x()) + (s * rec.normal.z()), rec.normal.y(), (-s * rec.normal.x()) + (c * rec.normal.z()));
    return true;
  }
  if (dynamic_cast<const constant_medium*>(h)) return false;
  if (h->hit(r, ray_t, rec)) {
    id = g_leaf_id[h];
    return true;
  }
  return false;
}

void apply_overrides(camera& cam) {
  if (g_opt.width > 0) cam.image_width = g_opt.width;
  if (g_opt.spp > 0) cam.samples_per_pixel = g_opt.spp;
  if (g_opt.depth > 0) cam.max_depth = g_opt.depth;
}

void run_ppm(camera& cam, const hittable& world) {
  std::ofstream os(g_opt.out);
  counting_world cw(world);
  auto t0 = std::chrono::steady_clock::now();
  cam.render(os, cw);  // the reference's own camera::render, untouched
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("JSON {\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"max_depth\": %d, \"rays\": %llu, \"seconds\": %.6f}\n",
              g_opt.scene.c_str(), cam.image_width, cam.image_height, cam.samples_per_pixel, cam.max_depth, cw.rays, sec);
}

void run_primary(camera& cam, const hittable& world) {
  cam.initialize();
  g_leaf_id.clear();
  number_leaves(&world);
  const int W = cam.image_width, H = cam.image_height;
  std::vector<int32_t> ids(size_t(W) * H);
  std::vector<double> ts(size_t(W) * H), ns(size_t(W) * H * 3);
  long mismatch = 0;
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      // pixel-centre ray: get_ray (camera.hpp:147,156) with zero jitter, no defocus, time 0
      auto pixel_sample = cam.pixel00_loc + ((i + 0.0) * cam.pixel_delta_u) + ((j + 0.0) * cam.pixel_delta_v);
      ray r(cam.camera_center, pixel_sample - cam.camera_center, 0.0);
      hit_record rec;
      int id = -1;
      bool ok = replay(&world, r, interval(0.001, infinity), rec, id);
      size_t p = size_t(j) * W + i;
      ids[p] = ok ? id : -1;
      ts[p] = ok ? rec.t : infinity;
      for (int c = 0; c < 3; c++) ns[3 * p + c] = ok ? rec.normal[c] : 0.0;
      if (!g_has_medium) {  // cross-check the replay against the reference's own world.hit
        hit_record ref;
        bool rok = world.hit(r, interval(0.001, infinity), ref);
        if (rok != ok || (ok && (ref.t != rec.t || ref.normal[0] != rec.normal[0]))) mismatch++;
      }
    }
  FILE* f = std::fopen(g_opt.out.c_str(), "wb");
  int32_t hdr[4] = {W, H, int32_t(g_leaf_id.size()), int32_t(mismatch)};
  std::fwrite(hdr, 4, 4, f);
  std::fwrite(ids.data(), 4, ids.size(), f);
  std::fwrite(ts.data(), 8, ts.size(), f);
  std::fwrite(ns.data(), 8, ns.size(), f);
  std::fclose(f);
  std::printf("JSON {\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"prims\": %d, \"replay_mismatch\": %ld}\n",
              g_opt.scene.c_str(), W, H, int(g_leaf_id.size()), mismatch);
}

void run_linear(camera& cam, const hittable& world) {
  cam.initialize();
  const int W = cam.image_width, H = cam.image_height, spp = cam.samples_per_pixel;
  std::vector<double> sum(size_t(W) * H * 3, 0.0), sq(size_t(W) * H * 3, 0.0);
  counting_world cw(world);
  auto t0 = std::chrono::steady_clock::now();
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      size_t p = (size_t(j) * W + i) * 3;
      for (int s = 0; s < spp; s++) {
        ray r = cam.get_ray(i, j);
        color c = cam.ray_color(r, cam.max_depth, cw);
        for (int k = 0; k < 3; k++) sum[p + k] += c[k], sq[p + k] += c[k] * c[k];
      }
    }
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  FILE* f = std::fopen(g_opt.out.c_str(), "wb");
  int32_t hdr[4] = {W, H, spp, 0};
  std::fwrite(hdr, 4, 4, f);
  std::fwrite(sum.data(), 8, sum.size(), f);
  std::fwrite(sq.data(), 8, sq.size(), f);
  std::fclose(f);
  std::printf("JSON {\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"max_depth\": %d, \"rays\": %llu, \"seconds\": %.6f}\n",
              g_opt.scene.c_str(), W, H, spp, cam.max_depth, cw.rays, sec);
}

void dispatch(camera& cam, std::ostream&, const hittable& world) {
  apply_overrides(cam);
  if (g_opt.seed >= 0) std::srand(unsigned(g_opt.seed));  // after scene construction
  if (g_opt.mode == "ppm") run_ppm(cam, world);
  else if (g_opt.mode == "primary") run_primary(cam, world);
  else if (g_opt.mode == "linear") run_linear(cam, world);
  else std::fprintf(stderr, "unknown mode %s\n", g_opt.mode.c_str()), std::exit(2);
}

}  // namespace

// ---- the shipped scene functions, unchanged; their cam.render(...) lands in dispatch() ----
#define main ref_main
#define render(os, w) image_width += 0, dispatch(cam, os, w)
#include "main.cpp"
#undef render
#undef main

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: ref_harness ppm|primary|linear <scene> <out> [--width W] [--spp S] [--depth D] [--seed K]\n");
    return 2;
  }
  g_opt.mode = argv[1];
  g_opt.scene = argv[2];
  g_opt.out = argv[3];
  for (int i = 4; i + 1 < argc; i += 2) {
    std::string k = argv[i];
    long v = std::atol(argv[i + 1]);
    if (k == "--width") g_opt.width = int(v);
    else if (k == "--spp") g_opt.spp = int(v);
    else if (k == "--depth") g_opt.depth = int(v);
    else if (k == "--seed") g_opt.seed = v;
    else { std::fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
  }
  std::ofstream sink("/dev/null");
  const std::string prefix = "shipped:";
  if (g_opt.scene.compare(0, prefix.size(), prefix) == 0) {
    std::string n = g_opt.scene.substr(prefix.size());
    if (n == "bouncing_spheres") bouncing_spheres(sink);
    else if (n == "checkered_spheres") checkered_spheres(sink);
    else if (n == "earth") earth(sink);
    else if (n == "perlin_sphere") perlin_sphere(sink);
    else if (n == "quads") quads(sink);
    else if (n == "simple_light") simple_light(sink);
    else if (n == "cornell_box") cornell_box(sink);
    else { std::fprintf(stderr, "unknown shipped scene %s\n", n.c_str()); return 2; }
    return 0;
  }
  rtb200_scenes::scene_setup s;
  if (!rtb200_scenes::build_scene(g_opt.scene, s)) {
    std::fprintf(stderr, "unknown scene %s\n", g_opt.scene.c_str());
    return 2;
  }
  dispatch(s.cam, sink, s.world);
  return 0;
}
