// scenes.hpp — the named scenes of BASELINE.json, written against the PUBLIC scene API only
// (hittable_list::add, sphere/quad/box/translate/rotate_y/constant_medium/bvh_node, the
// materials and textures, the camera fields).  Include it AFTER either header set:
//   * raytracing-practice_b200/host/rtb200_host.hpp  -> flattened and rendered on the GPU
//   * the reference's own headers + oracle/ref_ext.hpp -> rendered by the reference (oracle/_ref)
// so both sides consume the same rand() stream and see the identical scene.
//
// The seven shipped scenes restate src/main.cpp (line ranges cited per function); the
// draws from rand() happen in the same order and inside the same expression shapes, since
// g++'s argument evaluation order decides which coordinate gets which draw.
// The three scenes the reference does not ship follow SURVEY.md Appendix B.4-B.6.
#ifndef RTB200_SCENES_HPP
#define RTB200_SCENES_HPP

#include <memory>
#include <string>

namespace rtb200_scenes {

struct scene_setup {
  hittable_list world;
  camera cam;
};

using std::make_shared;
using std::shared_ptr;

inline void look(camera& cam, double vfov, const point3& from, const point3& at) {
  cam.vfov = vfov;
  cam.lookfrom = from;
  cam.lookat = at;
  cam.vup = vec3(0.0f, 1.0f, 0.0f);
}

inline void film(camera& cam, int width, double aspect, int spp, int depth, const color& bg) {
  cam.image_width = width;
  cam.aspect_ratio = aspect;
  cam.samples_per_pixel = spp;
  cam.max_depth = depth;
  cam.background = bg;
}

// The 22x22 grid of small spheres shared by bouncing_spheres (main.cpp:24-63) and the
// Book-1 final scene (Appendix B.4: same grid, but the diffuse ones do not move).
inline void scatter_small_spheres(hittable_list& world, bool moving) {
  for (int a = -11; a < 11; a++) {
    for (int b = -11; b < 11; b++) {
      auto choose_mat = random_double();
      point3 center(a + 0.9f * random_double(), 0.2f, b + 0.9f * random_double());
      if ((center - point3(4.0f, 0.2f, 0.0f)).length() > 0.9f) {
        shared_ptr<material> m;
        if (choose_mat < 0.8f) {
          auto albedo = color::random() * color::random();
          m = make_shared<lambertian>(albedo);
          if (moving) {
            auto center2 = center + vec3(0.0f, random_double(0.0f, 0.5f), 0.0f);
            world.add(make_shared<sphere>(center, center2, 0.2f, m));
          } else {
            world.add(make_shared<sphere>(center, 0.2f, m));
          }
        } else if (choose_mat < 0.95f) {
          auto albedo = color::random(0.5f, 1.0f);
          auto fuzz = random_double(0.0f, 0.5f);
          m = make_shared<metal>(albedo, fuzz);
          world.add(make_shared<sphere>(center, 0.2f, m));
        } else {
          m = make_shared<dielectric>(1.5f);
          world.add(make_shared<sphere>(center, 0.2f, m));
        }
      }
    }
  }
}

inline void three_big_spheres(hittable_list& world) {
  world.add(make_shared<sphere>(point3(0.0f, 1.0f, 0.0f), 1.0f, make_shared<dielectric>(1.5f)));
  world.add(make_shared<sphere>(point3(-4.0f, 1.0f, 0.0f), 1.0f, make_shared<lambertian>(color(0.4f, 0.2f, 0.1f))));
  world.add(make_shared<sphere>(point3(4.0f, 1.0f, 0.0f), 1.0f, make_shared<metal>(color(0.7f, 0.6f, 0.5f), 0.0f)));
}

// src/main.cpp:12-101
inline void bouncing_spheres(scene_setup& s) {
  auto checker = make_shared<checker_texture>(0.32f, color(0.2f, 0.3f, 0.1f), color(0.9f, 0.9f, 0.9f));
  s.world.add(make_shared<sphere>(point3(0.0f, -1000.0f, -1.0f), 1000.0f, make_shared<lambertian>(checker)));
  scatter_small_spheres(s.world, true);
  three_big_spheres(s.world);
  s.world = hittable_list(make_shared<bvh_node>(s.world));
  film(s.cam, 400, 16.0f / 9.0f, 50, 20, color(0.7f, 0.8f, 1.0f));
  look(s.cam, 20.0f, point3(13.0f, 2.0f, 3.0f), point3(0.0f, 0.0f, 0.0f));
  s.cam.defocus_angle = 0.6f;
  s.cam.focus_dist = 10.0f;
}

// SURVEY.md Appendix B.4 (config C1): grey ground at y=-1000, static spheres, BVH, 1200 wide.
inline void book1_final(scene_setup& s) {
  s.world.add(make_shared<sphere>(point3(0.0f, -1000.0f, 0.0f), 1000.0f, make_shared<lambertian>(color(0.5f, 0.5f, 0.5f))));
  scatter_small_spheres(s.world, false);
  three_big_spheres(s.world);
  s.world = hittable_list(make_shared<bvh_node>(s.world));
  film(s.cam, 1200, 16.0f / 9.0f, 10, 50, color(0.7f, 0.8f, 1.0f));
  look(s.cam, 20.0f, point3(13.0f, 2.0f, 3.0f), point3(0.0f, 0.0f, 0.0f));
  s.cam.defocus_angle = 0.6f;
  s.cam.focus_dist = 10.0f;
}

// src/main.cpp:104-138
inline void checkered_spheres(scene_setup& s) {
  auto checker = make_shared<checker_texture>(0.32f, color(0.2f, 0.3f, 0.1f), color(0.9f, 0.9f, 0.9f));
  s.world.add(make_shared<sphere>(point3(0.0f, -10.0f, 0.0f), 10.0f, make_shared<lambertian>(checker)));
  s.world.add(make_shared<sphere>(point3(0.0f, 10.0f, 0.0f), 10.0f, make_shared<lambertian>(checker)));
  film(s.cam, 400, 16.0f / 9.0f, 50, 20, color(0.7f, 0.8f, 1.0f));
  look(s.cam, 20.0f, point3(13.0f, 2.0f, 3.0f), point3(0.0f, 0.0f, 0.0f));
  s.cam.defocus_angle = 0.0f;
}

// src/main.cpp:141-171
inline void earth(scene_setup& s) {
  auto surface = make_shared<lambertian>(make_shared<image_texture>("earthmap.jpg"));
  s.world.add(make_shared<sphere>(point3(0.0f, 0.0f, 0.0f), 2.0f, surface));
  film(s.cam, 400, 16.0f / 9.0f, 100, 50, color(0.7f, 0.8f, 1.0f));
  look(s.cam, 20.0f, point3(0.0f, 0.0f, 12.0f), point3(0.0f, 0.0f, 0.0f));
  s.cam.defocus_angle = 0.0f;
}

// src/main.cpp:174-207
inline void perlin_sphere(scene_setup& s) {
  auto marble = make_shared<noise_texture>(4);
  s.world.add(make_shared<sphere>(point3(0.0f, -1000.0f, 0.0f), 1000.0f, make_shared<lambertian>(marble)));
  s.world.add(make_shared<sphere>(point3(0.0f, 2.0f, 0.0f), 2.0f, make_shared<lambertian>(marble)));
  film(s.cam, 400, 16.0f / 9.0f, 100, 50, color(0.7f, 0.8f, 1.0f));
  look(s.cam, 20.0f, point3(13.0f, 2.0f, 3.0f), point3(0.0f, 0.0f, 0.0f));
  s.cam.defocus_angle = 0.0f;
}

// src/main.cpp:210-251
inline void quads(scene_setup& s) {
  struct row { float q[3], u[3], v[3], rgb[3]; };
  static const row rows[5] = {
      {{-3.0f, -2.0f, 5.0f}, {0.0f, 0.0f, -4.0f}, {0.0f, 4.0f, 0.0f}, {1.0f, 0.2f, 0.2f}},   // left, red
      {{-2.0f, -2.0f, 0.0f}, {4.0f, 0.0f, 0.0f}, {0.0f, 4.0f, 0.0f}, {0.2f, 1.0f, 0.2f}},    // back, green
      {{3.0f, -2.0f, 1.0f}, {0.0f, 0.0f, 4.0f}, {0.0f, 4.0f, 0.0f}, {0.2f, 0.2f, 1.0f}},     // right, blue
      {{-2.0f, 3.0f, 1.0f}, {4.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 4.0f}, {1.0f, 0.5f, 0.0f}},     // upper, orange
      {{-2.0f, -3.0f, 5.0f}, {4.0f, 0.0f, 0.0f}, {0.0f, 0.0f, -4.0f}, {0.2f, 0.8f, 0.8f}}};  // lower, teal
  // materials are created first, then the quads, as in main.cpp:216-227 (no rand() involved)
  shared_ptr<material> mats[5];
  for (int i = 0; i < 5; i++) mats[i] = make_shared<lambertian>(color(rows[i].rgb[0], rows[i].rgb[1], rows[i].rgb[2]));
  for (int i = 0; i < 5; i++)
    s.world.add(make_shared<quad>(point3(rows[i].q[0], rows[i].q[1], rows[i].q[2]), vec3(rows[i].u[0], rows[i].u[1], rows[i].u[2]),
                                  vec3(rows[i].v[0], rows[i].v[1], rows[i].v[2]), mats[i]));
  film(s.cam, 400, 1.0f, 100, 50, color(0.7f, 0.8f, 1.0f));
  look(s.cam, 80.0f, point3(0.0f, 0.0f, 9.0f), point3(0.0f, 0.0f, 0.0f));
  s.cam.defocus_angle = 0.0f;
}

// src/main.cpp:254-298
inline void simple_light(scene_setup& s) {
  auto marble = make_shared<noise_texture>(4);
  s.world.add(make_shared<sphere>(point3(0.0f, -1000.0f, 0.0f), 1000.0f, make_shared<lambertian>(marble)));
  s.world.add(make_shared<sphere>(point3(0.0f, 2.0f, 0.0f), 2.0f, make_shared<lambertian>(marble)));
  auto lamp = make_shared<diffuse_light>(color(4.0f, 4.0f, 4.0f));
  s.world.add(make_shared<sphere>(point3(0.0f, 7.0f, 0.0f), 2.0f, lamp));
  s.world.add(make_shared<quad>(point3(3.0f, 1.0f, -2.0f), vec3(2.0f, 0.0f, 0.0f), vec3(0.0f, 2.0f, 0.0f), lamp));
  film(s.cam, 400, 16.0f / 9.0f, 100, 50, color(0.0f, 0.0f, 0.0f));
  look(s.cam, 20.0f, point3(26.0f, 3.0f, 6.0f), point3(0.0f, 2.0f, 0.0f));
  s.cam.defocus_angle = 0.0f;
}

// The five walls + ceiling lamp of every Cornell variant (main.cpp:307-318).
struct cornell_materials {
  shared_ptr<material> red, white, green;
};
inline cornell_materials cornell_room(hittable_list& world, const point3& lamp_q, const vec3& lamp_u, const vec3& lamp_v,
                                      double lamp_power) {
  cornell_materials m;
  m.red = make_shared<lambertian>(color(0.65f, 0.05f, 0.05f));
  m.white = make_shared<lambertian>(color(0.73f, 0.73f, 0.73f));
  m.green = make_shared<lambertian>(color(0.12f, 0.45f, 0.15f));
  auto lamp = make_shared<diffuse_light>(color(lamp_power, lamp_power, lamp_power));
  const double L = 555.0f;
  world.add(make_shared<quad>(point3(L, 0.0f, 0.0f), vec3(0.0f, L, 0.0f), vec3(0.0f, 0.0f, L), m.green));
  world.add(make_shared<quad>(point3(0.0f, 0.0f, 0.0f), vec3(0.0f, L, 0.0f), vec3(0.0f, 0.0f, L), m.red));
  world.add(make_shared<quad>(lamp_q, lamp_u, lamp_v, lamp));
  world.add(make_shared<quad>(point3(0.0f, 0.0f, 0.0f), vec3(L, 0.0f, 0.0f), vec3(0.0f, 0.0f, L), m.white));
  world.add(make_shared<quad>(point3(L, L, L), vec3(-L, 0.0f, 0.0f), vec3(0.0f, 0.0f, -L), m.white));
  world.add(make_shared<quad>(point3(0.0f, 0.0f, L), vec3(L, 0.0f, 0.0f), vec3(0.0f, L, 0.0f), m.white));
  return m;
}
inline void cornell_camera(camera& cam, int spp) {
  film(cam, 600, 1.0f, spp, 50, color(0.0f, 0.0f, 0.0f));
  look(cam, 40.0f, point3(278.0f, 278.0f, -800.0f), point3(278.0f, 278.0f, 0.0f));
  cam.defocus_angle = 0.0f;
}

// src/main.cpp:301-346 — the shipped variant: two axis-aligned boxes, no instances.
inline void cornell_box(scene_setup& s) {
  auto m = cornell_room(s.world, point3(343.0f, 554.0f, 332.0f), vec3(-130.0f, 0.0f, 0.0f), vec3(0.0f, 0.0f, -105.0f), 15.0f);
  s.world.add(box(point3(130.0f, 0.0f, 65.0f), point3(295.0f, 165.0f, 230.0f), m.white));
  s.world.add(box(point3(265.0f, 0.0f, 295.0f), point3(430.0f, 330.0f, 460.0f), m.white));
  cornell_camera(s.cam, 100);
}

inline shared_ptr<hittable> placed_box(double w, double h, double d, double angle, const vec3& at, shared_ptr<material> m) {
  shared_ptr<hittable> b = box(point3(0, 0, 0), point3(w, h, d), m);
  b = make_shared<rotate_y>(b, angle);
  return make_shared<translate>(b, at);
}

// Intermediate test scene: the book's Cornell box with rotated + translated box instances.
inline void cornell_rotated(scene_setup& s) {
  auto m = cornell_room(s.world, point3(343.0f, 554.0f, 332.0f), vec3(-130.0f, 0.0f, 0.0f), vec3(0.0f, 0.0f, -105.0f), 15.0f);
  s.world.add(placed_box(165, 330, 165, 15, vec3(265, 0, 295), m.white));
  s.world.add(placed_box(165, 165, 165, -18, vec3(130, 0, 65), m.white));
  cornell_camera(s.cam, 100);
}

// SURVEY.md Appendix B.5 (config C4): bigger, dimmer lamp; both boxes become smoke volumes.
inline void cornell_smoke(scene_setup& s) {
  auto m = cornell_room(s.world, point3(113, 554, 127), vec3(330, 0, 0), vec3(0, 0, 305), 7);
  auto box1 = placed_box(165, 330, 165, 15, vec3(265, 0, 295), m.white);
  auto box2 = placed_box(165, 165, 165, -18, vec3(130, 0, 65), m.white);
  s.world.add(make_shared<constant_medium>(box1, 0.01, color(0, 0, 0)));
  s.world.add(make_shared<constant_medium>(box2, 0.01, color(1, 1, 1)));
  cornell_camera(s.cam, 200);
}

// SURVEY.md Appendix B.6 (config C5, the headline scene).  rand() order: 400 box heights,
// the noise_texture's perlin tables, then 1000 sphere centres.
inline void book2_final(scene_setup& s) {
  hittable_list boxes1;
  auto ground = make_shared<lambertian>(color(0.48, 0.83, 0.53));
  const int boxes_per_side = 20;
  for (int i = 0; i < boxes_per_side; i++) {
    for (int j = 0; j < boxes_per_side; j++) {
      auto w = 100.0;
      auto x0 = -1000.0 + i * w;
      auto z0 = -1000.0 + j * w;
      auto y0 = 0.0;
      auto x1 = x0 + w;
      auto y1 = random_double(1, 101);
      auto z1 = z0 + w;
      boxes1.add(box(point3(x0, y0, z0), point3(x1, y1, z1), ground));
    }
  }
  s.world.add(make_shared<bvh_node>(boxes1));

  auto lamp = make_shared<diffuse_light>(color(7, 7, 7));
  s.world.add(make_shared<quad>(point3(123, 554, 147), vec3(300, 0, 0), vec3(0, 0, 265), lamp));

  auto center1 = point3(400, 400, 200);
  auto center2 = center1 + vec3(30, 0, 0);
  s.world.add(make_shared<sphere>(center1, center2, 50, make_shared<lambertian>(color(0.7, 0.3, 0.1))));

  s.world.add(make_shared<sphere>(point3(260, 150, 45), 50, make_shared<dielectric>(1.5)));
  s.world.add(make_shared<sphere>(point3(0, 150, 145), 50, make_shared<metal>(color(0.8, 0.8, 0.9), 1.0)));

  auto boundary = make_shared<sphere>(point3(360, 150, 145), 70, make_shared<dielectric>(1.5));
  s.world.add(boundary);
  s.world.add(make_shared<constant_medium>(boundary, 0.2, color(0.2, 0.4, 0.9)));
  boundary = make_shared<sphere>(point3(0, 0, 0), 5000, make_shared<dielectric>(1.5));
  s.world.add(make_shared<constant_medium>(boundary, .0001, color(1, 1, 1)));

  auto emat = make_shared<lambertian>(make_shared<image_texture>("earthmap.jpg"));
  s.world.add(make_shared<sphere>(point3(400, 200, 400), 100, emat));
  auto pertext = make_shared<noise_texture>(0.2);
  s.world.add(make_shared<sphere>(point3(220, 280, 300), 80, make_shared<lambertian>(pertext)));

  hittable_list boxes2;
  auto white = make_shared<lambertian>(color(.73, .73, .73));
  const int ns = 1000;
  for (int j = 0; j < ns; j++) boxes2.add(make_shared<sphere>(point3::random(0, 165), 10, white));
  s.world.add(make_shared<translate>(make_shared<rotate_y>(make_shared<bvh_node>(boxes2), 15), vec3(-100, 270, 395)));

  film(s.cam, 800, 1.0, 10000, 40, color(0, 0, 0));
  look(s.cam, 40, point3(478, 278, -600), point3(278, 278, 0));
  s.cam.defocus_angle = 0;
}

typedef void (*scene_fn)(scene_setup&);
struct scene_entry {
  const char* name;
  scene_fn build;
};
inline const scene_entry* scene_table(int* count) {
  static const scene_entry table[] = {
      {"bouncing_spheres", bouncing_spheres}, {"checkered_spheres", checkered_spheres}, {"earth", earth},
      {"perlin_sphere", perlin_sphere},       {"quads", quads},                         {"simple_light", simple_light},
      {"cornell_box", cornell_box},           {"book1_final", book1_final},             {"cornell_rotated", cornell_rotated},
      {"cornell_smoke", cornell_smoke},       {"book2_final", book2_final}};
  *count = int(sizeof(table) / sizeof(table[0]));
  return table;
}
inline bool build_scene(const std::string& name, scene_setup& s) {
  int n = 0;
  const scene_entry* t = scene_table(&n);
  for (int i = 0; i < n; i++)
    if (name == t[i].name) {
      t[i].build(s);
      return true;
    }
  return false;
}

}  // namespace rtb200_scenes
#endif
