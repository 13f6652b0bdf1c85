// A scene program written the way the reference's src/main.cpp is (same classes, same
// camera fields, cam.render(std::ofstream, world)), built against the host API.  Used by the GPU
// test of the drop-in boundary; on the build container the reference's own main.cpp is compiled too.
#include "common/rtweekend.hpp"
#include "accelerator/bvh_node.hpp"
#include "core/camera.hpp"
#include "core/material.hpp"
#include "core/texture.hpp"
#include "hittable/hittable.hpp"
#include "hittable/hittable_list.hpp"
#include "hittable/sphere.hpp"
#include "hittable/quad.hpp"
#include "scenes.hpp"

int main(int argc, char* argv[]) {
  if (argc < 5) {
    fprintf(stderr, "usage: dropin <scene> <out.ppm> <width> <spp> [renders in this process = 1]\n");
    return 2;
  }
  const int repeats = argc > 5 ? atoi(argv[5]) : 1;
  for (int rep = 0; rep < repeats; rep++) {  // > 1: a program that renders frame after frame (the contexts are kept)
    std::ofstream output_file(argv[2]);
    if (!output_file) {
      fprintf(stderr, "Error: could not open file %s for writing.\n", argv[2]);
      return 1;
    }
    rtb200_scenes::scene_setup s;
    if (!rtb200_scenes::build_scene(argv[1], s)) return 2;
    if (atoi(argv[3]) > 0) s.cam.image_width = atoi(argv[3]);
    if (atoi(argv[4]) > 0) s.cam.samples_per_pixel = atoi(argv[4]);
    s.cam.render(output_file, s.world);
    output_file.close();
  }
  return 0;
}
