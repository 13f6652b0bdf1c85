/* oracle.h — C entry points of the CPU oracle (oracle/oracle.cpp).
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker; never by the product path. */
#ifndef RTB200_ORACLE_H
#define RTB200_ORACLE_H
#include <stdint.h>

#include "../include/rt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

int orc_camera_initialize(const rt_camera_desc* cam, rt_camera_frame* out);
int orc_render_ppm(const rt_scene_desc* scene, const rt_camera_desc* cam, const char* path, uint64_t* rays_out);
int orc_render_linear(const rt_scene_desc* scene, const rt_camera_desc* cam, int spp, unsigned seed, int threads,
                      double* sum, double* sumsq, uint64_t* rays_out);
int orc_hit_rays(const rt_scene_desc* scene, int64_t n, const double* origin, const double* direction,
                 const double* time, double tmin, double tmax, int skip_media, int32_t* prim_id, double* t,
                 double* normal, uint8_t* front_face, double* uv, uint64_t* census);
int orc_primary(const rt_scene_desc* scene, const rt_camera_desc* cam, int skip_media, int32_t* prim_id, double* t,
                double* normal);
int orc_medium_spans(const rt_scene_desc* scene, int medium_index, int64_t n, const double* origin,
                     const double* direction, const double* time, double* t1, double* t2);
int orc_texture_value(const rt_scene_desc* scene, int texture, int64_t n, const double* uvp, double* rgb);
int orc_scatter(const rt_scene_desc* scene, int material, int64_t n, unsigned seed, const double* dir_in,
                const double* normal, const uint8_t* front_face, double* dir_out, double* attenuation,
                uint8_t* scattered);
void orc_write_color(int64_t npix, const double* linear_rgb, uint8_t* bytes);

#ifdef __cplusplus
}
#endif
#endif
