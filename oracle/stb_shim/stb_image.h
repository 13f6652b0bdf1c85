// stb_image.h SHIM for building the unmodified reference as oracle/_ref.
// The reference includes <stb_image.h> (src/core/rtw_stb_image.hpp:11) but does not vendor
// it (_cmake/stb.cmake:6-9 downloads nothings/stb @ f4a71b13...).  There is no network, so
// this shim provides the one entry point the reference calls — stbi_loadf(path,&w,&h,&n,3)
// (rtw_stb_image.hpp:79) — on top of this repo's own baseline-JPEG decoder, followed by
// stb's documented LDR->float conversion (gamma 2.2, scale 1.0).  Test infrastructure only.
#ifndef RTB200_STB_SHIM_H
#define RTB200_STB_SHIM_H
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../raytracing-practice_b200/host/rtb200_jpeg.hpp"

#define STBI_FREE(p) free(p)

static inline float* stbi_loadf(const char* filename, int* x, int* y, int* comp, int req_comp) {
  (void)req_comp;
  std::vector<unsigned char> rgb;
  int w = 0, h = 0;
  if (!rtb200::load_image_rgb8(filename, rgb, w, h)) return nullptr;
  float* out = static_cast<float*>(malloc(rgb.size() * sizeof(float)));
  for (size_t i = 0; i < rgb.size(); i++) out[i] = float(std::pow(rgb[i] / 255.0f, 2.2f));
  *x = w, *y = h;
  if (comp) *comp = 3;
  return out;
}
#endif
