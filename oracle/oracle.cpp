// oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY: nothing in the product path may link,
// import or execute this file (only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do).
//
// A double-precision restatement of the reference's hot path (camera::render ->
// ray_color -> hittable::hit -> material::scatter -> texture::value), operating on the
// same flattened scene description (include/rt_b200.h) the CUDA library consumes.  Every
// function cites the reference file:line it follows; operation ORDER is kept (division is
// "multiply by 1/t", float literals are floats, turb accumulates in float, ...) because the
// oracle is pinned by requiring BIT-IDENTICAL PPM output and primary-hit t/normal against
// the unmodified reference compiled as oracle/_ref (tests/golden/, oracle/README.md).
// rand() draws are taken in the order g++ evaluates the reference's expressions
// (right-to-left function arguments).
//
// rotate_y / constant_medium / isotropic are not in the reference; they follow the book
// semantics of SURVEY.md Appendix B ("parity unpinned" by the reference for those three).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../include/rt_b200.h"
#include "oracle.h"

namespace {

const double kInf = std::numeric_limits<double>::infinity();
const double kPi = 3.1415926535897932385;  // rtweekend.hpp:15

// ---- vec3 (common/vec3.hpp) ----------------------------------------------------------
struct V {
  double x, y, z;
};
inline V mk(double a, double b, double c) { return V{a, b, c}; }
inline V operator+(V a, V b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V operator-(V a, V b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V operator-(V a) { return mk(-a.x, -a.y, -a.z); }
inline V operator*(V a, V b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V operator*(double t, V v) { return mk(t * v.x, t * v.y, t * v.z); }  // vec3.hpp:120-123
inline V operator*(V v, double t) { return t * v; }
inline V operator/(V v, double t) { return (1 / t) * v; }  // vec3.hpp:131-134
inline double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V cross(V a, V b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline double len2(V a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double len(V a) { return std::sqrt(len2(a)); }
inline V unit(V a) { return a / len(a); }  // vec3.hpp:152-155
inline V ld3(const double* p) { return mk(p[0], p[1], p[2]); }
inline double comp(V v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

struct Ray {
  V o, d;
  double tm;
  V at(double t) const { return o + t * d; }  // ray.hpp:22-26
};

// ---- RNG: glibc rand() stream (rtweekend.hpp:23-39) ------------------------------------
// Global mode draws from ::rand() (so a render continues the stream the scene builder
// left, like the reference).  Thread mode replays the same TYPE_3 additive-feedback
// generator through random_r with a private state, i.e. srand(seed) per thread.
struct Rng {
  bool global = true;
  bool xoshiro = false;  // diagnostic mode: a high-quality generator instead of glibc's
                         // lagged-Fibonacci rand() (which has 3-point correlations), to tell
                         // "reference RNG artefact" from "algorithm" in statistical comparisons
  struct random_data rd;
  char state[128];
  uint64_t xs[4];
  void seed_private(unsigned s) {
    global = false;
    std::memset(&rd, 0, sizeof rd);
    std::memset(state, 0, sizeof state);
    initstate_r(s, state, sizeof state, &rd);
    uint64_t z = 0x9E3779B97F4A7C15ull * (uint64_t(s) + 1);
    for (int i = 0; i < 4; i++) {  // splitmix64
      z += 0x9E3779B97F4A7C15ull;
      uint64_t x = z;
      x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
      x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
      xs[i] = x ^ (x >> 31);
    }
  }
  int next() {
    if (global) return std::rand();
    if (xoshiro) {  // xoshiro256**, top 31 bits
      auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
      uint64_t r = rotl(xs[1] * 5, 7) * 9, t = xs[1] << 17;
      xs[2] ^= xs[0], xs[3] ^= xs[1], xs[1] ^= xs[2], xs[0] ^= xs[3], xs[2] ^= t, xs[3] = rotl(xs[3], 45);
      return int(r >> 33);
    }
    int32_t v;
    random_r(&rd, &v);
    return v;
  }
  double uniform() { return next() / (RAND_MAX + 1.0f); }                    // :26  int / float
  double uniform(double lo, double hi) { return lo + (hi - lo) * uniform(); }  // :32
  // vec3::random(min,max) (vec3.hpp:85-88): g++ evaluates the three arguments right to left.
  V vec(double lo, double hi) {
    double z = uniform(lo, hi);
    double y = uniform(lo, hi);
    double x = uniform(lo, hi);
    return mk(x, y, z);
  }
  V unit_vector() {  // vec3.hpp:172-184
    for (;;) {
      V p = vec(-1, 1);
      double l2 = len2(p);
      if (1e-160 < l2 && l2 <= 1) return p / std::sqrt(l2);
    }
  }
  V in_unit_disk() {  // vec3.hpp:158-169 (second argument drawn first)
    for (;;) {
      double y = uniform(-1.0f, 1.0f);
      double x = uniform(-1.0f, 1.0f);
      V p = mk(x, y, 0.0f);
      if (len2(p) < 1.0f) return p;
    }
  }
};

struct Rec {  // hittable.hpp:16-36
  V p, n;
  int mat = -1;
  double t = 0, u = 0, v = 0;
  bool front = false;
  int prim = -1;  // harness addition: DFS leaf id of the primitive that produced the record
};
inline void set_face_normal(Rec& rec, const Ray& r, V outward) {  // hittable.hpp:29-35
  rec.front = dot(r.d, outward) < 0;
  rec.n = rec.front ? outward : -outward;
}

struct QuadDerived {
  V normal, w;
  double D;
};

struct Scene {
  const rt_scene_desc* d;
  std::vector<QuadDerived> quad;  // indexed by hittable index
  bool skip_media = false;
  uint64_t* counters = nullptr;  // optional census: [0] aabb tests, [1] aabb passes, [2] sphere tests, [3] quad tests
};

void derive(Scene& s) {
  s.quad.assign(size_t(s.d->n_hittables), QuadDerived());
  for (int i = 0; i < s.d->n_hittables; i++) {
    const rt_hittable& h = s.d->hittables[i];
    if (h.kind != RT_H_QUAD) continue;
    V Q = ld3(h.p), u = ld3(h.p + 3), v = ld3(h.p + 6);  // quad.hpp:12-27
    V n = cross(u, v);
    QuadDerived q;
    q.normal = unit(n);
    q.D = dot(q.normal, Q);
    q.w = n / dot(n, n);
    s.quad[size_t(i)] = q;
  }
}

// ---- aabb::hit (accelerator/aabb.hpp:61-112) -------------------------------------------
bool box_hit(const Scene& s, const double* b, const Ray& r, double tmin, double tmax) {
  if (s.counters) s.counters[0]++;
  for (int axis = 0; axis < 3; axis++) {
    const double adinv = 1.0f / comp(r.d, axis);
    double t0 = (b[2 * axis] - comp(r.o, axis)) * adinv;
    double t1 = (b[2 * axis + 1] - comp(r.o, axis)) * adinv;
    if (t0 < t1) {
      if (t0 > tmin) tmin = t0;
      if (t1 < tmax) tmax = t1;
    } else {
      if (t1 > tmin) tmin = t1;
      if (t0 < tmax) tmax = t0;
    }
    if (tmax <= tmin) return false;
  }
  if (s.counters) s.counters[1]++;
  return true;
}

bool hit(const Scene& s, int idx, const Ray& r, double tmin, double tmax, Rec& rec, Rng& rng);

// sphere::hit (hittable/sphere.hpp:47-93) + get_sphere_uv (:100-111)
bool sphere_hit(const Scene& s, const rt_hittable& h, const Ray& r, double tmin, double tmax, Rec& rec) {
  if (s.counters) s.counters[2]++;
  Ray center{ld3(h.p), ld3(h.p + 3), 0};
  const double radius = h.p[6];
  V cc = center.at(r.tm);
  V oc = r.o - cc;
  double a = len2(r.d);
  double half_b = dot(oc, r.d);
  double c = len2(oc) - radius * radius;
  double disc = half_b * half_b - a * c;
  if (disc < 0) return false;
  double sq = std::sqrt(disc);
  double root = (-half_b - sq) / a;
  if (!(tmin < root && root < tmax)) {  // interval::surrounds, interval.hpp:32
    root = (-half_b + sq) / a;
    if (!(tmin < root && root < tmax)) return false;
  }
  rec.t = root;
  rec.p = r.at(rec.t);
  V outward = (rec.p - cc) / radius;
  set_face_normal(rec, r, outward);
  double theta = std::acos(-outward.y);
  double phi = std::atan2(-outward.z, outward.x) + kPi;
  rec.u = phi / (2.0f * kPi);
  rec.v = theta / kPi;
  rec.mat = h.material;
  rec.prim = h.prim_id;
  return true;
}

// quad::hit (hittable/quad.hpp:44-94) + is_interior (:97-114)
bool quad_hit(const Scene& s, int idx, const rt_hittable& h, const Ray& r, double tmin, double tmax, Rec& rec) {
  if (s.counters) s.counters[3]++;
  const QuadDerived& q = s.quad[size_t(idx)];
  V Q = ld3(h.p), u = ld3(h.p + 3), v = ld3(h.p + 6);
  double denom = dot(q.normal, r.d);
  if (std::fabs(denom) < 1e-8) return false;
  double t = (q.D - dot(q.normal, r.o)) / denom;
  if (!(tmin <= t && t <= tmax)) return false;  // interval::contains, interval.hpp:29
  V ip = r.at(t);
  V hp = ip - Q;
  double alpha = dot(q.w, cross(hp, v));
  double beta = dot(q.w, cross(u, hp));
  if (!(0.0f <= alpha && alpha <= 1.0f) || !(0.0f <= beta && beta <= 1.0f)) return false;
  rec.u = alpha;
  rec.v = beta;
  rec.t = t;
  rec.p = ip;
  rec.mat = h.material;
  set_face_normal(rec, r, q.normal);
  rec.prim = h.prim_id;
  return true;
}

// constant_medium::hit — SURVEY.md Appendix B.2 (book semantics; not in the reference)
bool medium_hit(const Scene& s, const rt_hittable& h, const Ray& r, double tmin, double tmax, Rec& rec, Rng& rng) {
  if (s.skip_media) return false;
  Rec rec1, rec2;
  if (!hit(s, h.child0, r, -kInf, kInf, rec1, rng)) return false;
  if (!hit(s, h.child0, r, rec1.t + 0.0001, kInf, rec2, rng)) return false;
  if (rec1.t < tmin) rec1.t = tmin;
  if (rec2.t > tmax) rec2.t = tmax;
  if (rec1.t >= rec2.t) return false;
  if (rec1.t < 0) rec1.t = 0;
  double ray_length = len(r.d);
  double inside = (rec2.t - rec1.t) * ray_length;
  double hit_distance = h.p[1] * std::log(rng.uniform());
  if (hit_distance > inside) return false;
  rec.t = rec1.t + hit_distance / ray_length;
  rec.p = r.at(rec.t);
  rec.n = mk(1, 0, 0);
  rec.front = true;
  rec.mat = h.material;
  rec.u = 0;  // the book leaves u, v unset (the record is an uninitialised local of ray_color): defined as 0, 0, as in
  rec.v = 0;  // ref_ext.hpp and in the device's surface_at — only an image-textured phase function can tell
  rec.prim = -1;
  return true;
}

bool hit(const Scene& s, int idx, const Ray& r, double tmin, double tmax, Rec& rec, Rng& rng) {
  const rt_hittable& h = s.d->hittables[idx];
  switch (h.kind) {
    case RT_H_SPHERE: return sphere_hit(s, h, r, tmin, tmax, rec);
    case RT_H_QUAD: return quad_hit(s, idx, h, r, tmin, tmax, rec);
    case RT_H_LIST: {  // hittable_list::hit (hittable/hittable_list.hpp:40-64)
      Rec tmp;
      bool any = false;
      double closest = tmax;
      for (int k = 0; k < h.child1; k++) {
        if (hit(s, s.d->child_index[h.child0 + k], r, tmin, closest, tmp, rng)) {
          any = true;
          closest = tmp.t;
          rec = tmp;
        }
      }
      return any;
    }
    case RT_H_BVH: {  // bvh_node::hit (accelerator/bvh_node.hpp:80-94)
      if (!box_hit(s, h.bbox, r, tmin, tmax)) return false;
      bool hl = hit(s, h.child0, r, tmin, tmax, rec, rng);
      bool hr = hit(s, h.child1, r, tmin, hl ? rec.t : tmax, rec, rng);
      return hl || hr;
    }
    case RT_H_TRANSLATE: {  // translate::hit (hittable/hittable.hpp:86-104)
      Ray moved{r.o - ld3(h.p), r.d, r.tm};
      if (!hit(s, h.child0, moved, tmin, tmax, rec, rng)) return false;
      rec.p = rec.p + ld3(h.p);
      return true;
    }
    case RT_H_ROTATE_Y: {
      const double sn = h.p[1], cs = h.p[2];
      V o = mk(cs * r.o.x - sn * r.o.z, r.o.y, sn * r.o.x + cs * r.o.z);
      V d = mk(cs * r.d.x - sn * r.d.z, r.d.y, sn * r.d.x + cs * r.d.z);
      Ray rot{o, d, r.tm};
      if (!hit(s, h.child0, rot, tmin, tmax, rec, rng)) return false;
      rec.p = mk(cs * rec.p.x + sn * rec.p.z, rec.p.y, -sn * rec.p.x + cs * rec.p.z);
      rec.n = mk(cs * rec.n.x + sn * rec.n.z, rec.n.y, -sn * rec.n.x + cs * rec.n.z);
      return true;
    }
    case RT_H_MEDIUM: return medium_hit(s, h, r, tmin, tmax, rec, rng);
  }
  return false;
}

// ---- perlin (core/perlin.hpp) ----------------------------------------------------------
double perlin_interp(const V c[2][2][2], double u, double v, double w) {  // :219-255
  double uu = u * u * (3 - 2 * u);
  double vv = v * v * (3 - 2 * v);
  double ww = w * w * (3 - 2 * w);
  float accum = 0.0f;  // `auto accum = 0.0f` : float accumulator
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++) {
        V wv = mk(u - i, v - j, w - k);
        accum += (i * uu + (1 - i) * (1 - uu)) * (j * vv + (1 - j) * (1 - vv)) * (k * ww + (1 - k) * (1 - ww)) *
                 dot(c[i][j][k], wv);
      }
  return accum;
}
double noise_perlin(const rt_perlin& t, V p) {  // :95-132
  double u = p.x - std::floor(p.x);
  double v = p.y - std::floor(p.y);
  double w = p.z - std::floor(p.z);
  int i = int(std::floor(p.x));
  int j = int(std::floor(p.y));
  int k = int(std::floor(p.z));
  V c[2][2][2];
  for (int di = 0; di < 2; di++)
    for (int dj = 0; dj < 2; dj++)
      for (int dk = 0; dk < 2; dk++)
        c[di][dj][dk] = ld3(t.randvec[t.perm_x[(i + di) & 255] ^ t.perm_y[(j + dj) & 255] ^ t.perm_z[(k + dk) & 255]]);
  return perlin_interp(c, u, v, w);
}
double turb(const rt_perlin& t, V p, int depth) {  // :135-158
  float accum = 0.0f;
  V tp = p;
  float weight = 1.0f;
  for (int i = 0; i < depth; i++) {
    accum += weight * noise_perlin(t, tp);
    weight *= 0.5f;
    tp = mk(tp.x * 2.0f, tp.y * 2.0f, tp.z * 2.0f);  // vec3::operator*=
  }
  return std::fabs(accum);
}

// ---- texture::value (core/texture.hpp) -------------------------------------------------
V texture_value(const rt_scene_desc* d, int idx, double u, double v, V p) {
  const rt_texture& t = d->textures[idx];
  switch (t.kind) {
    case RT_T_SOLID: return ld3(t.color);  // :34-37
    case RT_T_CHECKER: {                   // :57-79
      int xi = int(std::floor(t.scale * p.x));
      int yi = int(std::floor(t.scale * p.y));
      int zi = int(std::floor(t.scale * p.z));
      bool even = (xi + yi + zi) % 2 == 0;
      return texture_value(d, even ? t.even : t.odd, u, v, p);
    }
    case RT_T_IMAGE: {  // :97-118 + rtw_stb_image.hpp:104-134
      const rt_image& im = d->images[t.image];
      if (im.height <= 0 || im.rgb == nullptr) return mk(0.0f, 1.0f, 1.0f);
      u = u < 0.0f ? 0.0f : (u > 1.0f ? 1.0f : u);
      v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
      v = 1.0f - v;
      int i = int(u * im.width);
      int j = int(v * im.height);
      i = i < 0 ? 0 : (i < im.width ? i : im.width - 1);
      j = j < 0 ? 0 : (j < im.height ? j : im.height - 1);
      const uint8_t* px = im.rgb + (size_t(j) * im.width + i) * 3;
      float cs = 1.0f / 255.0f;  // float * int -> float product (:116-117)
      return mk(cs * px[0], cs * px[1], cs * px[2]);
    }
    case RT_T_NOISE: {  // :133-151
      double s = 1.0f + std::sin(t.scale * p.z + 10.0f * turb(d->perlins[t.perlin], p, 7));
      return mk(0.5f, 0.5f, 0.5f) * s;
    }
  }
  return mk(0, 0, 0);
}

// ---- material (core/material.hpp) ------------------------------------------------------
V reflect(V v, V n) { return v - 2.0f * dot(v, n) * n; }  // vec3.hpp:207-213
V refract(V uv, V n, double eta) {                         // vec3.hpp:216-226
  double c = std::fmin(dot(-uv, n), 1.0f);
  V perp = eta * (uv + c * n);
  V par = -std::sqrt(std::fabs(1.0f - len2(perp))) * n;
  return perp + par;
}
bool near_zero(V e) {  // vec3.hpp:70-77, including the parenthesis slip on y
  double s = 1e-8;
  return (std::fabs(e.x) < s) && (std::fabs(double(e.y < s))) && (std::fabs(e.z) < s);
}

V emitted(const rt_scene_desc* d, int mat, double u, double v, V p) {
  const rt_material& m = d->materials[mat];
  if (m.kind == RT_M_DIFFUSE_LIGHT) return texture_value(d, m.texture, u, v, p);  // :233-236
  return mk(0.0f, 0.0f, 0.0f);                                                     // :29-33
}

bool scatter(const rt_scene_desc* d, int mat, const Ray& rin, const Rec& rec, V& att, Ray& out, Rng& rng) {
  const rt_material& m = d->materials[mat];
  switch (m.kind) {
    case RT_M_LAMBERTIAN: {  // :51-71
      V dir = rec.n + rng.unit_vector();
      if (near_zero(dir)) dir = rec.n;
      out = Ray{rec.p, dir, rin.tm};
      att = texture_value(d, m.texture, rec.u, rec.v, rec.p);
      return true;
    }
    case RT_M_METAL: {  // :86-106
      V refl = reflect(rin.d, rec.n);
      refl = unit(refl) + (m.fuzz * rng.unit_vector());
      out = Ray{rec.p, refl, rin.tm};
      att = ld3(m.albedo);
      return dot(out.d, rec.n) > 0;
    }
    case RT_M_DIELECTRIC: {  // :128-179, reflectance :198-206
      att = mk(1.0f, 1.0f, 1.0f);
      double ri = rec.front ? (1.0f / m.ior) : m.ior;
      V ud = unit(rin.d);
      double cos_t = std::fmin(dot(-ud, rec.n), 1.0f);
      double sin_t = std::sqrt(1.0f - cos_t * cos_t);
      bool cannot = ri * sin_t > 1.0f;
      bool do_reflect = cannot;
      if (!do_reflect) {
        double r0 = (1.0f - ri) / (1.0f + ri);
        r0 = r0 * r0;
        double refl = r0 + (1.0f - r0) * std::pow((1.0f - cos_t), 5);
        do_reflect = refl > rng.uniform();
      }
      V dir = do_reflect ? reflect(ud, rec.n) : refract(ud, rec.n, ri);
      out = Ray{rec.p, dir, rin.tm};
      return true;
    }
    case RT_M_DIFFUSE_LIGHT: return false;  // base material::scatter, :36
    case RT_M_ISOTROPIC: {
      out = Ray{rec.p, rng.unit_vector(), rin.tm};
      att = texture_value(d, m.texture, rec.u, rec.v, rec.p);
      return true;
    }
  }
  return false;
}

// ---- camera (core/camera.hpp) ----------------------------------------------------------
struct Cam {
  int W, H, spp, max_depth;
  double scale;
  V center, p00, du, dv, u, v, w, ddu, ddv, bg;
  double defocus_angle;
};

void cam_init(const rt_camera_desc* c, Cam& k) {  // camera::initialize :76-136
  k.W = c->image_width;
  k.H = static_cast<int>(c->image_width / c->aspect_ratio);
  k.H = k.H < 1 ? 1 : k.H;
  k.spp = c->samples_per_pixel;
  k.max_depth = c->max_depth;
  k.scale = 1.0f / c->samples_per_pixel;  // float / int -> float (:83)
  double theta = c->vfov * kPi / 180.0f;
  double h = std::tan(theta / 2);
  double vh = 2 * h * c->focus_dist;
  double vw = vh * (static_cast<double>(k.W) / k.H);
  V from = ld3(c->lookfrom), at = ld3(c->lookat), vup = ld3(c->vup);
  k.center = from;
  k.w = unit(from - at);
  k.u = unit(cross(vup, k.w));
  k.v = cross(k.w, k.u);
  V vu = vw * k.u;
  V vv = vh * (-k.v);
  k.du = vu / k.W;
  k.dv = vv / k.H;
  V ul = k.center - (c->focus_dist * k.w) - vu / 2 - vv / 2;
  k.p00 = ul + 0.5 * (k.du + k.dv);
  double dr = c->focus_dist * std::tan((c->defocus_angle * kPi / 180.0f) / 2.0f);
  k.ddu = k.u * dr;
  k.ddv = k.v * dr;
  k.bg = ld3(c->background);
  k.defocus_angle = c->defocus_angle;
}

Ray get_ray(const Cam& k, int i, int j, Rng& rng) {  // :139-162
  double oy = rng.uniform() - 0.5f;  // sample_square :165-168, second argument first
  double ox = rng.uniform() - 0.5f;
  V ps = k.p00 + ((i + ox) * k.du) + ((j + oy) * k.dv);
  V origin = k.center;
  if (!(k.defocus_angle <= 0.0f)) {
    V p = rng.in_unit_disk();  // defocus_disk_sample :171-177
    origin = k.center + (p.x * k.ddu) + (p.y * k.ddv);
  }
  V dir = ps - origin;
  double tm = rng.uniform();
  return Ray{origin, dir, tm};
}

Ray center_ray(const Cam& k, int i, int j) {  // SURVEY.md §8(c) primary-pass convention
  V ps = k.p00 + ((i + 0.0) * k.du) + ((j + 0.0) * k.dv);
  return Ray{k.center, ps - k.center, 0.0};
}

V ray_color(const Scene& s, const Cam& k, const Ray& r, int depth, Rng& rng, uint64_t& rays) {  // :180-232
  if (depth <= 0) return mk(0.0f, 0.0f, 0.0f);
  Rec rec;
  rays++;
  if (!hit(s, s.d->root, r, 0.001, kInf, rec, rng)) return k.bg;
  V emission = emitted(s.d, rec.mat, rec.u, rec.v, rec.p);
  Ray out;
  V att;
  if (!scatter(s.d, rec.mat, r, rec, att, out, rng)) return emission;
  V from_scatter = att * ray_color(s, k, out, depth - 1, rng, rays);
  return emission + from_scatter;
}

void to_bytes(V c, int* out) {  // write_color, common/color.hpp:26-58
  double ch[3] = {c.x, c.y, c.z};
  const double lo = 0.000f, hi = 0.999f;
  for (int a = 0; a < 3; a++) {
    double g = ch[a] > 0.0f ? std::sqrt(ch[a]) : 0.0f;
    g = g < lo ? lo : (g > hi ? hi : g);
    out[a] = int(256 * g);
  }
}

int check(const rt_scene_desc* d) {
  if (!d || d->abi_version != RT_B200_ABI_VERSION || d->root < 0 || d->root >= d->n_hittables) return -1;
  return 0;
}

}  // namespace

extern "C" {

int orc_camera_initialize(const rt_camera_desc* cam, rt_camera_frame* f) {
  Cam k;
  cam_init(cam, k);
  f->image_width = k.W;
  f->image_height = k.H;
  f->pixel_samples_scale = k.scale;
  const V* src[] = {&k.center, &k.p00, &k.du, &k.dv, &k.u, &k.v, &k.w, &k.ddu, &k.ddv};
  double* dst[] = {f->center, f->pixel00_loc, f->pixel_delta_u, f->pixel_delta_v, f->u, f->v, f->w, f->defocus_disk_u, f->defocus_disk_v};
  for (int i = 0; i < 9; i++) dst[i][0] = src[i]->x, dst[i][1] = src[i]->y, dst[i][2] = src[i]->z;
  return 0;
}

// camera::render as shipped (camera.hpp:29-72): single thread, global ::rand(), P3 text.
int orc_render_ppm(const rt_scene_desc* d, const rt_camera_desc* cam, const char* path, uint64_t* rays_out) {
  if (check(d)) return -1;
  Scene s{d, {}, false, nullptr};
  derive(s);
  Cam k;
  cam_init(cam, k);
  FILE* f = std::fopen(path, "w");
  if (!f) return -2;
  std::fprintf(f, "P3\n%d %d\n255\n", k.W, k.H);
  Rng rng;
  uint64_t rays = 0;
  for (int j = 0; j < k.H; j++)
    for (int i = 0; i < k.W; i++) {
      V px = mk(0.0f, 0.0f, 0.0f);
      for (int sidx = 0; sidx < k.spp; sidx++) {
        Ray r = get_ray(k, i, j, rng);
        V c = ray_color(s, k, r, k.max_depth, rng, rays);
        px = px + c;  // vec3::operator+=
      }
      int b[3];
      to_bytes(k.scale * px, b);
      std::fprintf(f, "%d %d %d\n", b[0], b[1], b[2]);
    }
  std::fclose(f);
  if (rays_out) *rays_out = rays;
  return 0;
}

// Linear-radiance statistics for converged-image tests and the multi-core CPU baseline:
// `spp` samples per pixel, rows interleaved over `threads` threads, thread t draws from a
// private glibc stream seeded with seed*4099 + t.  sum / sumsq are W*H*3 doubles.
int orc_render_linear(const rt_scene_desc* d, const rt_camera_desc* cam, int spp, unsigned seed, int threads,
                      double* sum, double* sumsq, uint64_t* rays_out) {
  if (check(d)) return -1;
  Scene s{d, {}, false, nullptr};
  derive(s);
  Cam k;
  cam_init(cam, k);
  if (threads < 1) threads = 1;
  std::vector<uint64_t> rays(size_t(threads), 0);
  auto work = [&](int tid) {
    Rng rng;
    rng.seed_private((seed & 0x7FFFFFFFu) * 4099u + unsigned(tid));
    rng.xoshiro = (seed & 0x80000000u) != 0;  // top bit of the seed selects the diagnostic generator
    uint64_t nr = 0;
    for (int j = tid; j < k.H; j += threads)
      for (int i = 0; i < k.W; i++) {
        double acc[3] = {0, 0, 0}, acc2[3] = {0, 0, 0};
        for (int sidx = 0; sidx < spp; sidx++) {
          Ray r = get_ray(k, i, j, rng);
          V c = ray_color(s, k, r, k.max_depth, rng, nr);
          acc[0] += c.x, acc[1] += c.y, acc[2] += c.z;
          acc2[0] += c.x * c.x, acc2[1] += c.y * c.y, acc2[2] += c.z * c.z;
        }
        size_t o = (size_t(j) * k.W + i) * 3;
        for (int a = 0; a < 3; a++) {
          sum[o + a] = acc[a];
          if (sumsq) sumsq[o + a] = acc2[a];
        }
      }
    rays[size_t(tid)] = nr;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  uint64_t total = 0;
  for (auto v : rays) total += v;
  if (rays_out) *rays_out = total;
  return 0;
}

// world.hit(r, interval(tmin, tmax), rec) for n arbitrary rays.
int orc_hit_rays(const rt_scene_desc* d, int64_t n, const double* origin, const double* direction, const double* time,
                 double tmin, double tmax, int skip_media, int32_t* prim_id, double* t, double* normal,
                 uint8_t* front_face, double* uv, uint64_t* census) {
  if (check(d)) return -1;
  Scene s{d, {}, skip_media != 0, census};
  derive(s);
  Rng rng;
  for (int64_t i = 0; i < n; i++) {
    Ray r{ld3(origin + 3 * i), ld3(direction + 3 * i), time ? time[i] : 0.0};
    Rec rec;
    bool ok = hit(s, d->root, r, tmin, tmax, rec, rng);
    if (prim_id) prim_id[i] = ok ? rec.prim : -1;
    if (t) t[i] = ok ? rec.t : kInf;
    if (normal) {
      normal[3 * i] = ok ? rec.n.x : 0;
      normal[3 * i + 1] = ok ? rec.n.y : 0;
      normal[3 * i + 2] = ok ? rec.n.z : 0;
    }
    if (front_face) front_face[i] = ok && rec.front;
    if (uv) uv[2 * i] = ok ? rec.u : 0, uv[2 * i + 1] = ok ? rec.v : 0;
  }
  return 0;
}

// Pixel-centre primary pass (no jitter, no defocus, time 0, interval (0.001, inf)).
int orc_primary(const rt_scene_desc* d, const rt_camera_desc* cam, int skip_media, int32_t* prim_id, double* t,
                double* normal) {
  if (check(d)) return -1;
  Cam k;
  cam_init(cam, k);
  const int64_t n = int64_t(k.W) * k.H;
  std::vector<double> o(size_t(3 * n)), dir(size_t(3 * n));
  for (int j = 0; j < k.H; j++)
    for (int i = 0; i < k.W; i++) {
      Ray r = center_ray(k, i, j);
      size_t p = (size_t(j) * k.W + i) * 3;
      o[p] = r.o.x, o[p + 1] = r.o.y, o[p + 2] = r.o.z;
      dir[p] = r.d.x, dir[p + 1] = r.d.y, dir[p + 2] = r.d.z;
    }
  return orc_hit_rays(d, n, o.data(), dir.data(), nullptr, 0.001, kInf, skip_media, prim_id, t, normal, nullptr, nullptr,
                      nullptr);
}

// Boundary entry / exit of the medium_index-th RT_H_MEDIUM (rec1.t, rec2.t before clamping).
int orc_medium_spans(const rt_scene_desc* d, int medium_index, int64_t n, const double* origin, const double* direction,
                     const double* time, double* t1, double* t2) {
  if (check(d)) return -1;
  Scene s{d, {}, false, nullptr};
  derive(s);
  int found = -1, seen = 0;
  for (int i = 0; i < d->n_hittables; i++)
    if (d->hittables[i].kind == RT_H_MEDIUM && seen++ == medium_index) found = i;
  if (found < 0) return -1;
  const rt_hittable& h = d->hittables[found];
  Rng rng;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (int64_t i = 0; i < n; i++) {
    Ray r{ld3(origin + 3 * i), ld3(direction + 3 * i), time ? time[i] : 0.0};
    Rec a, b;
    t1[i] = t2[i] = nan;
    if (!hit(s, h.child0, r, -kInf, kInf, a, rng)) continue;
    t1[i] = a.t;
    if (!hit(s, h.child0, r, a.t + 0.0001, kInf, b, rng)) continue;
    t2[i] = b.t;
  }
  return 0;
}

// texture::value(u, v, p): uvp = n x (u, v, px, py, pz), rgb = n x 3 doubles.
int orc_texture_value(const rt_scene_desc* d, int texture, int64_t n, const double* uvp, double* rgb) {
  if (!d || texture < 0 || texture >= d->n_textures) return -1;
  for (int64_t i = 0; i < n; i++) {
    const double* q = uvp + 5 * i;
    V c = texture_value(d, texture, q[0], q[1], mk(q[2], q[3], q[4]));
    rgb[3 * i] = c.x, rgb[3 * i + 1] = c.y, rgb[3 * i + 2] = c.z;
  }
  return 0;
}

// material::scatter for n synthetic hits (distribution tests), private glibc stream.
int orc_scatter(const rt_scene_desc* d, int material, int64_t n, unsigned seed, const double* dir_in,
                const double* normal, const uint8_t* front_face, double* dir_out, double* attenuation,
                uint8_t* scattered) {
  if (!d || material < 0 || material >= d->n_materials) return -1;
  Rng rng;
  rng.seed_private(seed);
  for (int64_t i = 0; i < n; i++) {
    Ray rin{mk(0, 0, 0), ld3(dir_in + 3 * i), 0.0};
    Rec rec;
    rec.p = mk(0, 0, 0);
    rec.n = ld3(normal + 3 * i);
    rec.front = front_face[i] != 0;
    rec.mat = material;
    V att = mk(0, 0, 0);
    Ray out{mk(0, 0, 0), mk(0, 0, 0), 0};
    bool ok = scatter(d, material, rin, rec, att, out, rng);
    scattered[i] = ok;
    dir_out[3 * i] = out.d.x, dir_out[3 * i + 1] = out.d.y, dir_out[3 * i + 2] = out.d.z;
    attenuation[3 * i] = att.x, attenuation[3 * i + 1] = att.y, attenuation[3 * i + 2] = att.z;
  }
  return 0;
}

// write_color's gamma / clamp / int(256 x) on a linear fp64 mean image.
void orc_write_color(int64_t npix, const double* linear_rgb, uint8_t* bytes) {
  for (int64_t i = 0; i < npix; i++) {
    int b[3];
    to_bytes(mk(linear_rgb[3 * i], linear_rgb[3 * i + 1], linear_rgb[3 * i + 2]), b);
    bytes[3 * i] = uint8_t(b[0]), bytes[3 * i + 1] = uint8_t(b[1]), bytes[3 * i + 2] = uint8_t(b[2]);
  }
}

}  // extern "C"
