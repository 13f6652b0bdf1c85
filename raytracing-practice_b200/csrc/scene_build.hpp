// scene_build.hpp — host side of rt_upload_scene: rt_scene_desc (the reference's hittable
// graph as POD) -> flat device arrays.
//
//   * translate / rotate_y instances (src/hittable/hittable.hpp:74-117) are BAKED
//     into world-space primitives: a rotated+translated sphere is a sphere, a rotated quad is
//     a quad.  One single-level BVH, no per-ray instance transform on the fp32 path.
//   * hittable_list / bvh_node containers (hittable_list.hpp:40-64, bvh_node.hpp:21-94) vanish:
//     their leaves feed one SAH-built BVH2.  The reference's median-split tree is poor (the
//     radius-1000 ground sphere sits inside it: 45 box tests per ray); tree shape
//     never changes which primitive is closest, only exact-tie order, which the exact
//     predicate resolves by the reference's visit order (`order`).
//   * constant_medium becomes one BVH leaf item bounded by its boundary; the
//     boundary's own primitives live in the primitive arrays but not in the BVH (unless the
//     scene also adds them as surfaces).
#ifndef RTB200_SCENE_BUILD_HPP
#define RTB200_SCENE_BUILD_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "device_scene.h"

namespace rtb200 {

struct HostScene {
  std::vector<float4> nodes;
  std::vector<uint32_t> leaf_refs;
  std::vector<float4> spheres;
  std::vector<int2> sph_meta;
  std::vector<float4> quads;
  std::vector<int> quad_mat;
  std::vector<float4> boxes;
  std::vector<int4> box_meta;
  std::vector<DMedium> media;
  std::vector<uint32_t> medium_brefs;
  std::vector<float4> materials;
  std::vector<float4> textures;
  std::vector<uchar4> texels;
  std::vector<int4> images;
  std::vector<float4> perlin_vec;
  std::vector<uint8_t> perlin_perm;
  std::vector<float2> rotations;
  std::vector<XSphere> xspheres;
  std::vector<XQuad> xquads;
  std::vector<XOp> xops;
  std::vector<int2> xchains;
  std::vector<int> global_media;  // media enclosing every other item (at most 4)
  float scene_abs_max = 1.0f;
  float bounds_lo[3] = {std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity()};
  float bounds_hi[3] = {-std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity()};
  int bvh_depth = 0;
  double sah_cost = 0;
  std::string error;
};

namespace build_detail {

struct d3 {
  double x, y, z;
};
inline d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline d3 operator*(double t, d3 v) { return {t * v.x, t * v.y, t * v.z}; }
inline double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline d3 cross(d3 a, d3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline d3 ld(const double* p) { return {p[0], p[1], p[2]}; }

// world = R * local + t, R = rotation about y with (x,z) -> (c x + s z, -s x + c z)
struct Xform {
  double c = 1, s = 0;
  d3 t{0, 0, 0};
  bool rotated = false;
  d3 vec(d3 v) const { return {c * v.x + s * v.z, v.y, -s * v.x + c * v.z}; }
  d3 point(d3 p) const { return vec(p) + t; }
};

struct Box {
  float lo[3], hi[3];
  void reset() {
    for (int a = 0; a < 3; a++) lo[a] = std::numeric_limits<float>::infinity(), hi[a] = -lo[a];
  }
  void grow(const Box& b) {
    for (int a = 0; a < 3; a++) lo[a] = std::min(lo[a], b.lo[a]), hi[a] = std::max(hi[a], b.hi[a]);
  }
  double area() const {
    double dx = double(hi[0]) - lo[0], dy = double(hi[1]) - lo[1], dz = double(hi[2]) - lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0;
    return 2 * (dx * dy + dy * dz + dz * dx);
  }
};
inline float round_down(double v) {
  float f = float(v);
  return double(f) > v ? std::nextafter(f, -std::numeric_limits<float>::infinity()) : f;
}
inline float round_up(double v) {
  float f = float(v);
  return double(f) < v ? std::nextafter(f, std::numeric_limits<float>::infinity()) : f;
}
inline Box box_from(const double lo[3], const double hi[3]) {
  Box b;
  for (int a = 0; a < 3; a++) b.lo[a] = round_down(lo[a]), b.hi[a] = round_up(hi[a]);
  return b;
}

struct Item {
  uint32_t ref;
  Box box;
  float cost;
};

struct Builder {
  const rt_scene_desc* d;
  HostScene* out;
  std::vector<Item> items;          // BVH leaves
  std::vector<XOp> chain;           // current instance wrappers, outermost first
  std::map<std::vector<double>, int> chain_ids;
  int order = 0;
  int collecting_medium = -1;       // >= 0: primitives go to that medium's boundary list
  std::vector<uint32_t>* boundary = nullptr;
  Box boundary_box;
  int depth_guard = 0;

  bool fail(const std::string& why) {
    if (out->error.empty()) out->error = why;
    return false;
  }

  int current_chain() {
    std::vector<double> key;
    for (const XOp& o : chain) {
      key.push_back(o.kind);
      key.push_back(o.a[0]), key.push_back(o.a[1]), key.push_back(o.a[2]);
    }
    auto it = chain_ids.find(key);
    if (it != chain_ids.end()) return it->second;
    int id = int(out->xchains.size());
    out->xchains.push_back(int2{int(out->xops.size()), int(chain.size())});
    for (const XOp& o : chain) out->xops.push_back(o);
    chain_ids[key] = id;
    return id;
  }

  void emit(uint32_t ref, const Box& b, float cost) {
    if (boundary) {
      boundary->push_back(ref);
      boundary_box.grow(b);
    } else {
      items.push_back(Item{ref, b, cost});
    }
  }

  bool add_sphere(const rt_hittable& h, const Xform& X) {
    if (h.material < 0 || h.material >= d->n_materials) return fail("sphere: bad material index");
    d3 c0 = X.point(ld(h.p)), dc = X.vec(ld(h.p + 3));
    double r = h.p[6];
    int idx = int(out->spheres.size() / 2);
    out->spheres.push_back(float4{float(c0.x), float(c0.y), float(c0.z), float(r)});
    out->spheres.push_back(float4{float(dc.x), float(dc.y), float(dc.z), 0.0f});
    int rot = -1;
    if (X.rotated) {
      rot = int(out->rotations.size());
      out->rotations.push_back(float2{float(X.s), float(X.c)});
    }
    out->sph_meta.push_back(int2{h.material, rot});
    XSphere xs;
    std::memset(&xs, 0, sizeof xs);
    for (int a = 0; a < 3; a++) xs.c[a] = h.p[a], xs.dc[a] = h.p[3 + a];
    xs.r = r;
    xs.chain = current_chain();
    xs.order = order++;
    xs.pid = h.prim_id;
    out->xspheres.push_back(xs);
    // bounds over time in [0,1] (sphere.hpp:39-43), radius padded by the fp32 rounding of c and r
    d3 c1 = c0 + dc;
    double ar = std::fabs(r);
    double pad = 4e-7 * (std::fabs(c0.x) + std::fabs(c0.y) + std::fabs(c0.z) + std::fabs(dc.x) + std::fabs(dc.y) + std::fabs(dc.z) + ar);
    double lo[3] = {std::min(c0.x, c1.x) - ar - pad, std::min(c0.y, c1.y) - ar - pad, std::min(c0.z, c1.z) - ar - pad};
    double hi[3] = {std::max(c0.x, c1.x) + ar + pad, std::max(c0.y, c1.y) + ar + pad, std::max(c0.z, c1.z) + ar + pad};
    emit(make_ref(REF_SPHERE, uint32_t(idx)), box_from(lo, hi), 1.2f);
    return true;
  }

  bool add_quad(const rt_hittable& h, const Xform& X, bool as_item = true) {
    if (h.material < 0 || h.material >= d->n_materials) return fail("quad: bad material index");
    // exact data first: object space, reference operation order (quad.hpp:17-23)
    d3 Qo = ld(h.p), uo = ld(h.p + 3), vo = ld(h.p + 6);
    XQuad xq;
    std::memset(&xq, 0, sizeof xq);
    {
      d3 n = cross(uo, vo);
      double len = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
      d3 normal = (1 / len) * n;  // unit_vector: v / length = (1/length) * v
      double D = dot(normal, Qo);
      d3 w = (1 / dot(n, n)) * n;
      const d3* src[5] = {&Qo, &uo, &vo, &normal, &w};
      double* dst[5] = {xq.Q, xq.u, xq.v, xq.n, xq.w};
      for (int k = 0; k < 5; k++) dst[k][0] = src[k]->x, dst[k][1] = src[k]->y, dst[k][2] = src[k]->z;
      xq.D = D;
    }
    xq.chain = current_chain();
    xq.order = order++;
    xq.pid = h.prim_id;
    out->xquads.push_back(xq);
    // baked fp32 record, world space
    d3 Q = X.point(Qo), u = X.vec(uo), v = X.vec(vo);
    d3 n = cross(u, v);
    double nn = dot(n, n);
    if (!(nn > 0)) return fail("quad: degenerate (u x v == 0)");
    d3 normal = (1 / std::sqrt(nn)) * n;
    double D = dot(normal, Q);
    d3 w = (1 / nn) * n;
    d3 A = cross(v, w);  // alpha = w . (hp x v) = hp . (v x w)
    d3 B = cross(w, u);  // beta  = w . (u x hp) = hp . (w x u)
    int idx = int(out->quads.size() / 3);
    out->quads.push_back(float4{float(normal.x), float(normal.y), float(normal.z), float(D)});
    out->quads.push_back(float4{float(A.x), float(A.y), float(A.z), float(-dot(A, Q))});
    out->quads.push_back(float4{float(B.x), float(B.y), float(B.z), float(-dot(B, Q))});
    out->quad_mat.push_back(h.material);
    d3 corner[4] = {Q, Q + u, Q + v, Q + u + v};
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, amax = 0;
    for (const d3& p : corner) {
      double c[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; a++) lo[a] = std::min(lo[a], c[a]), hi[a] = std::max(hi[a], c[a]), amax = std::max(amax, std::fabs(c[a]));
    }
    double pad = 5e-5 + 4e-7 * amax;  // thin-axis padding as aabb.hpp:135-154, plus fp32 slack
    for (int a = 0; a < 3; a++) lo[a] -= pad, hi[a] += pad;
    if (as_item) emit(make_ref(REF_QUAD, uint32_t(idx)), box_from(lo, hi), 1.0f);
    return true;
  }

  // box(a, b, mat) (quad.hpp:129-159) is a hittable_list of six quads.  When a list IS such a box — six
  // axis-aligned rectangles of one material that tile the surface of [lo, hi] in the list's own frame — it
  // becomes ONE slab-test primitive (REF_BOX): 1 BVH item instead of 6, 1 intersection instead of up to 6.
  // The six quad records are still written (uv, primitive ids, the fp64 exact predicate), just not as BVH items.
  // Returns false (nothing emitted) when the list is anything else.
  bool try_box(const rt_hittable& list, const Xform& X, bool& ok) {
    if (std::getenv("RT_B200_NO_BOXES")) return false;
    if (list.child1 != 6) return false;
    const rt_hittable* q[6];
    for (int k = 0; k < 6; k++) {
      int ci = d->child_index[list.child0 + k];
      if (ci < 0 || ci >= d->n_hittables) return false;
      q[k] = &d->hittables[ci];
      if (q[k]->kind != RT_H_QUAD || q[k]->material != q[0]->material) return false;
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < 6; k++)
      for (int c = 0; c < 4; c++)
        for (int a = 0; a < 3; a++) {
          double v = q[k]->p[a] + ((c & 1) ? q[k]->p[3 + a] : 0.0) + ((c & 2) ? q[k]->p[6 + a] : 0.0);
          lo[a] = std::min(lo[a], v), hi[a] = std::max(hi[a], v);
        }
    double tol = 0;  // box() builds its corners as min + (max - min): equal to max only up to double rounding
    for (int a = 0; a < 3; a++) {
      if (!(hi[a] > lo[a])) return false;
      tol = std::max(tol, 1e-12 * std::max(std::fabs(lo[a]), std::fabs(hi[a])));
    }
    auto near = [tol](double x, double y) { return std::fabs(x - y) <= tol; };
    int face_quad[6] = {-1, -1, -1, -1, -1, -1}, face_flip[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6; k++) {
      const double* Q = q[k]->p;
      const double* u = q[k]->p + 3;
      const double* v = q[k]->p + 6;
      int au = -1, av = -1;
      for (int a = 0; a < 3; a++) {
        if (u[a] != 0) { if (au >= 0) return false; au = a; }
        if (v[a] != 0) { if (av >= 0) return false; av = a; }
      }
      if (au < 0 || av < 0 || au == av) return false;
      const int an = 3 - au - av;  // the face's constant axis
      int side;
      if (near(Q[an], lo[an])) side = 0; else if (near(Q[an], hi[an])) side = 1; else return false;
      // the rectangle must span the whole face
      for (int a : {au, av}) {
        const double e = a == au ? u[a] : v[a];
        const double x0 = std::min(Q[a], Q[a] + e), x1 = std::max(Q[a], Q[a] + e);
        if (!near(x0, lo[a]) || !near(x1, hi[a])) return false;
      }
      const int f = an * 2 + side;
      if (face_quad[f] >= 0) return false;
      face_quad[f] = k;
      d3 n = cross(ld(u), ld(v));
      const double nn = an == 0 ? n.x : (an == 1 ? n.y : n.z);
      face_flip[f] = (nn > 0) == (side == 1) ? 0 : 1;  // 1: the quad's normal points INTO the box
    }
    const int first_quad = int(out->quads.size() / 3);
    for (int k = 0; k < 6 && ok; k++) ok = add_quad(*q[k], X, false);
    if (!ok) return true;
    int map = 0;
    for (int f = 0; f < 6; f++) map |= (face_quad[f] | (face_flip[f] << 3)) << (4 * f);
    const int idx = int(out->boxes.size() / 3);
    const bool xf = X.rotated;
    // an unrotated box takes the translation into its bounds and needs no ray transform at all
    const d3 t = xf ? X.t : d3{0, 0, 0};
    const d3 sh = xf ? d3{0, 0, 0} : X.t;
    out->boxes.push_back(float4{float(lo[0] + sh.x), float(lo[1] + sh.y), float(lo[2] + sh.z), float(xf ? X.c : 1.0)});
    out->boxes.push_back(float4{float(hi[0] + sh.x), float(hi[1] + sh.y), float(hi[2] + sh.z), float(xf ? X.s : 0.0)});
    out->boxes.push_back(float4{float(t.x), float(t.y), float(t.z), xf ? 1.0f : 0.0f});
    out->box_meta.push_back(int4{q[0]->material, first_quad, map, 0});
    double wlo[3] = {1e300, 1e300, 1e300}, whi[3] = {-1e300, -1e300, -1e300}, amax = 0;
    for (int c = 0; c < 8; c++) {
      d3 p = X.point(d3{(c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2]});
      const double pc[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; a++) wlo[a] = std::min(wlo[a], pc[a]), whi[a] = std::max(whi[a], pc[a]), amax = std::max(amax, std::fabs(pc[a]));
    }
    const double pad = 5e-5 + 4e-7 * amax;
    for (int a = 0; a < 3; a++) wlo[a] -= pad, whi[a] += pad;
    emit(make_ref(REF_BOX, (uint32_t(idx) << 3) | 7u), box_from(wlo, whi), 1.5f);
    return true;
  }

  bool visit(int idx, const Xform& X) {
    if (idx < 0 || idx >= d->n_hittables) return fail("hittable index out of range");
    if (++depth_guard > 4096) return fail("hittable graph too deep (cycle?)");
    const rt_hittable& h = d->hittables[idx];
    bool ok = true;
    switch (h.kind) {
      case RT_H_SPHERE: ok = add_sphere(h, X); break;
      case RT_H_QUAD: ok = add_quad(h, X); break;
      case RT_H_LIST:
        if (h.child0 < 0 || h.child1 < 0 || h.child0 + h.child1 > d->n_child_index) { ok = fail("list: bad child range"); break; }
        if (try_box(h, X, ok)) break;
        for (int k = 0; k < h.child1 && ok; k++) ok = visit(d->child_index[h.child0 + k], X);
        break;
      case RT_H_BVH:
        ok = visit(h.child0, X);
        if (ok && h.child1 != h.child0) ok = visit(h.child1, X);  // span-1 node: same child twice (bvh_node.hpp:57)
        break;
      case RT_H_TRANSLATE: {
        Xform Y = X;
        Y.t = X.vec(ld(h.p)) + X.t;
        XOp op{0, 0, {h.p[0], h.p[1], h.p[2]}};
        chain.push_back(op);
        ok = visit(h.child0, Y);
        chain.pop_back();
        break;
      }
      case RT_H_ROTATE_Y: {
        const double s = h.p[1], c = h.p[2];
        Xform Y = X;  // world = X.R * (Rtheta * local) + X.t ; y-rotations commute and add angles
        Y.c = X.c * c - X.s * s;
        Y.s = X.s * c + X.c * s;
        Y.rotated = true;
        XOp op{1, 0, {s, c, 0}};
        chain.push_back(op);
        ok = visit(h.child0, Y);
        chain.pop_back();
        break;
      }
      case RT_H_MEDIUM: {
        if (boundary) { ok = fail("constant_medium nested inside another medium's boundary is not supported"); break; }
        if (h.material < 0 || h.material >= d->n_materials) { ok = fail("medium: bad material index"); break; }
        std::vector<uint32_t> refs;
        boundary = &refs;
        boundary_box.reset();
        ok = visit(h.child0, X);
        boundary = nullptr;
        if (!ok) break;
        if (refs.empty()) break;  // empty boundary: never hit
        DMedium m;
        m.neg_inv_density = float(h.p[1]);
        m.material = h.material;
        m.first_bref = int(out->medium_brefs.size());
        m.n_bref = int(refs.size());
        out->medium_brefs.insert(out->medium_brefs.end(), refs.begin(), refs.end());
        int mi = int(out->media.size());
        out->media.push_back(m);
        items.push_back(Item{make_ref(REF_MEDIUM, uint32_t(mi)), boundary_box, 1.5f + 1.0f * refs.size()});
        break;
      }
      default: ok = fail("unknown hittable kind " + std::to_string(h.kind));
    }
    --depth_guard;
    return ok;
  }
};

// ---- SAH BVH2 over the leaf items --------------------------------------------------------
struct TmpNode {
  Box box;
  int left = -1, right = -1;  // children (TmpNode indices) or -1
  int first = 0, count = 0;   // leaf range in the ordered item list
};

struct SahBuilder {
  std::vector<Item>& items;
  std::vector<TmpNode> nodes;
  std::vector<Item> ordered;
  // SAH constants in units of one node visit (two slab tests), fitted on the GPU (profiles/r08_sah_constants.md):
  // 0.5 while a node step cost ~75 instructions, 1.0 since the traversal diet brought it down to 52.
  double prim_scale = 1.0;
  int max_leaf = kMaxLeaf;
  explicit SahBuilder(std::vector<Item>& it) : items(it) {
    if (const char* e = std::getenv("RT_B200_SAH_PRIM")) prim_scale = std::max(0.01, std::atof(e));
    if (const char* e = std::getenv("RT_B200_MAX_LEAF")) max_leaf = std::min(8, std::max(1, std::atoi(e)));
  }

  int build(int begin, int end) {
    TmpNode n;
    n.box.reset();
    float leaf_cost = 0;
    for (int i = begin; i < end; i++) n.box.grow(items[size_t(i)].box), leaf_cost += float(prim_scale) * items[size_t(i)].cost;
    const int count = end - begin;
    int best_axis = -1, best_split = -1;
    double best = std::numeric_limits<double>::infinity();
    if (count > 1) {
      const double parent_area = std::max(n.box.area(), 1e-30);
      std::vector<double> right_area(static_cast<size_t>(count), 0.0), right_cost(static_cast<size_t>(count), 0.0);
      for (int axis = 0; axis < 3; axis++) {
        std::sort(items.begin() + begin, items.begin() + end, [axis](const Item& a, const Item& b) {
          double ca = double(a.box.lo[axis]) + a.box.hi[axis], cb = double(b.box.lo[axis]) + b.box.hi[axis];
          return ca < cb || (ca == cb && a.ref < b.ref);
        });
        Box acc;
        acc.reset();
        double c = 0;
        for (int i = count - 1; i > 0; i--) {
          acc.grow(items[size_t(begin + i)].box);
          c += prim_scale * items[size_t(begin + i)].cost;
          right_area[size_t(i)] = acc.area();
          right_cost[size_t(i)] = c;
        }
        acc.reset();
        c = 0;
        for (int i = 1; i < count; i++) {
          acc.grow(items[size_t(begin + i - 1)].box);
          c += prim_scale * items[size_t(begin + i - 1)].cost;
          double sah = 1.0 + (acc.area() * c + right_area[size_t(i)] * right_cost[size_t(i)]) / parent_area;
          if (sah < best) best = sah, best_axis = axis, best_split = i;
        }
      }
    }
    const bool make_leaf = count <= 1 || (count <= max_leaf && double(leaf_cost) <= best);
    if (make_leaf) {
      n.first = int(ordered.size());
      n.count = count;
      for (int i = begin; i < end; i++) ordered.push_back(items[size_t(i)]);
      nodes.push_back(n);
      return int(nodes.size()) - 1;
    }
    const int axis = best_axis;
    std::sort(items.begin() + begin, items.begin() + end, [axis](const Item& a, const Item& b) {
      double ca = double(a.box.lo[axis]) + a.box.hi[axis], cb = double(b.box.lo[axis]) + b.box.hi[axis];
      return ca < cb || (ca == cb && a.ref < b.ref);
    });
    int me = int(nodes.size());
    nodes.push_back(n);
    int l = build(begin, begin + best_split);
    int r = build(begin + best_split, end);
    nodes[size_t(me)].left = l;
    nodes[size_t(me)].right = r;
    return me;
  }
};

inline int encode_leaf(int first, int count) { return ~((first << 3) | (count - 1)); }

inline void pack_node(HostScene& out, const Box& b0, int c0, const Box& b1, int c1) {
  float4 a{b0.lo[0], b0.lo[1], b0.lo[2], b0.hi[0]};
  float4 b{b0.hi[1], b0.hi[2], b1.lo[0], b1.lo[1]};
  float4 c{b1.lo[2], b1.hi[0], b1.hi[1], b1.hi[2]};
  float4 dd;
  std::memcpy(&dd.x, &c0, 4);
  std::memcpy(&dd.y, &c1, 4);
  dd.z = dd.w = 0;
  out.nodes.push_back(a), out.nodes.push_back(b), out.nodes.push_back(c), out.nodes.push_back(dd);
}

}  // namespace build_detail

// Does the texture tree rooted at t contain an image texture (the only consumer of u,v)?
inline bool texture_needs_uv(const rt_scene_desc* d, int t, int guard = 0) {
  if (t < 0 || t >= d->n_textures || guard > 16) return false;
  const rt_texture& x = d->textures[t];
  if (x.kind == RT_T_IMAGE) return true;
  if (x.kind == RT_T_CHECKER) return texture_needs_uv(d, x.even, guard + 1) || texture_needs_uv(d, x.odd, guard + 1);
  return false;
}

inline bool build_host_scene(const rt_scene_desc* d, HostScene& out) {
  using namespace build_detail;
  if (!d || d->abi_version != RT_B200_ABI_VERSION) { out.error = "bad scene description / ABI version"; return false; }
  if (d->root < 0 || d->root >= d->n_hittables) { out.error = "root index out of range"; return false; }

  // ---- tables -----------------------------------------------------------------------
  for (int i = 0; i < d->n_textures; i++) {
    const rt_texture& t = d->textures[i];
    int a = -1, b = -1, kind = 0;
    switch (t.kind) {
      case RT_T_SOLID: kind = TEX_SOLID; break;
      case RT_T_CHECKER:
        kind = TEX_CHECKER, a = t.even, b = t.odd;
        if (a < 0 || a >= d->n_textures || b < 0 || b >= d->n_textures) { out.error = "checker: bad child texture"; return false; }
        break;
      case RT_T_IMAGE:
        kind = TEX_IMAGE, a = t.image;
        if (a < 0 || a >= d->n_images) { out.error = "image texture: bad image index"; return false; }
        break;
      case RT_T_NOISE:
        kind = TEX_NOISE, a = t.perlin;
        if (a < 0 || a >= d->n_perlins) { out.error = "noise texture: bad perlin index"; return false; }
        break;
      default: out.error = "unknown texture kind " + std::to_string(t.kind); return false;
    }
    out.textures.push_back(float4{float(t.color[0]), float(t.color[1]), float(t.color[2]), float(t.scale)});
    float4 meta;
    std::memcpy(&meta.x, &kind, 4), std::memcpy(&meta.y, &a, 4), std::memcpy(&meta.z, &b, 4);
    meta.w = 0;
    out.textures.push_back(meta);
  }
  for (int i = 0; i < d->n_materials; i++) {
    const rt_material& m = d->materials[i];
    int kind = 0, tex = m.texture, flags = 0;
    float param = 0;
    switch (m.kind) {
      case RT_M_LAMBERTIAN: kind = MAT_LAMBERTIAN; break;
      case RT_M_METAL: kind = MAT_METAL, param = float(m.fuzz), tex = -1; break;
      case RT_M_DIELECTRIC: kind = MAT_DIELECTRIC, param = float(m.ior), tex = -1; break;
      case RT_M_DIFFUSE_LIGHT: kind = MAT_LIGHT; break;
      case RT_M_ISOTROPIC: kind = MAT_ISOTROPIC; break;
      default: out.error = "unknown material kind " + std::to_string(m.kind); return false;
    }
    if (kind == MAT_LAMBERTIAN || kind == MAT_LIGHT || kind == MAT_ISOTROPIC) {
      if (tex < 0 || tex >= d->n_textures) { out.error = "material: bad texture index"; return false; }
      if (texture_needs_uv(d, tex)) flags |= MATF_NEEDS_UV;
    }
    out.materials.push_back(float4{float(m.albedo[0]), float(m.albedo[1]), float(m.albedo[2]), param});
    float4 meta;
    std::memcpy(&meta.x, &kind, 4), std::memcpy(&meta.y, &tex, 4), std::memcpy(&meta.z, &flags, 4);
    meta.w = 0;
    out.materials.push_back(meta);
  }
  for (int i = 0; i < d->n_images; i++) {
    const rt_image& im = d->images[i];
    int4 rec{int(out.texels.size()), im.rgb ? im.width : 0, im.rgb ? im.height : 0, 0};
    if (im.rgb && im.width > 0 && im.height > 0)
      for (size_t p = 0; p < size_t(im.width) * im.height; p++)
        out.texels.push_back(uchar4{im.rgb[3 * p], im.rgb[3 * p + 1], im.rgb[3 * p + 2], 255});
    out.images.push_back(rec);
  }
  for (int i = 0; i < d->n_perlins; i++) {
    const rt_perlin& p = d->perlins[i];
    for (int k = 0; k < 256; k++) out.perlin_vec.push_back(float4{float(p.randvec[k][0]), float(p.randvec[k][1]), float(p.randvec[k][2]), 0});
    for (int k = 0; k < 256; k++) out.perlin_perm.push_back(uint8_t(p.perm_x[k]));
    for (int k = 0; k < 256; k++) out.perlin_perm.push_back(uint8_t(p.perm_y[k]));
    for (int k = 0; k < 256; k++) out.perlin_perm.push_back(uint8_t(p.perm_z[k]));
  }

  // ---- geometry ---------------------------------------------------------------------
  Builder b;
  b.d = d;
  b.out = &out;
  if (!b.visit(d->root, Xform())) return false;

  // A medium whose boundary box contains every other item (the Book-2 scene's r=5000 fog) is met by
  // every ray: take it out of the BVH and let the kernel sample it once per ray, warp-converged.
  if (b.items.size() > 1) {
    for (size_t i = 0; i < b.items.size() && out.global_media.size() < 4;) {
      const Item& it = b.items[i];
      bool encloses = (it.ref >> 30) == REF_MEDIUM;
      for (size_t j = 0; j < b.items.size() && encloses; j++) {
        if (j == i) continue;
        for (int a = 0; a < 3; a++)
          if (b.items[j].box.lo[a] < it.box.lo[a] || b.items[j].box.hi[a] > it.box.hi[a]) encloses = false;
      }
      if (encloses) {
        out.global_media.push_back(int(it.ref & 0x3FFFFFFFu));
        b.items.erase(b.items.begin() + long(i));
      } else {
        i++;
      }
    }
  }

  float amax = 1.0f;
  for (const Item& it : b.items)
    for (int a = 0; a < 3; a++) {
      if (std::isfinite(it.box.lo[a])) amax = std::max(amax, std::fabs(it.box.lo[a]));
      if (std::isfinite(it.box.hi[a])) amax = std::max(amax, std::fabs(it.box.hi[a]));
    }
  out.scene_abs_max = amax;
  for (const Item& it : b.items)
    for (int a = 0; a < 3; a++) out.bounds_lo[a] = std::min(out.bounds_lo[a], it.box.lo[a]), out.bounds_hi[a] = std::max(out.bounds_hi[a], it.box.hi[a]);

  // ---- BVH ----------------------------------------------------------------------------
  Box empty;
  empty.reset();
  if (b.items.empty()) {
    pack_node(out, empty, encode_leaf(0, 1), empty, encode_leaf(0, 1));
    out.leaf_refs.push_back(REF_NONE);
    return true;
  }
  SahBuilder sah(b.items);
  int root = sah.build(0, int(b.items.size()));
  for (const Item& it : sah.ordered) out.leaf_refs.push_back(it.ref);
  const std::vector<TmpNode>& tn = sah.nodes;
  if (tn[size_t(root)].left < 0) {  // the whole scene is one leaf
    pack_node(out, tn[size_t(root)].box, encode_leaf(tn[size_t(root)].first, tn[size_t(root)].count), empty, encode_leaf(0, 1));
    out.bvh_depth = 1;
    return true;
  }
  // breadth-first numbering of the internal nodes: a prefix of the array = the top levels
  std::vector<int> bfs{root}, index_of(tn.size(), -1);
  for (size_t q = 0; q < bfs.size(); q++) {
    const TmpNode& n = tn[size_t(bfs[q])];
    index_of[size_t(bfs[q])] = int(q);
    if (tn[size_t(n.left)].left >= 0) bfs.push_back(n.left);
    if (tn[size_t(n.right)].left >= 0) bfs.push_back(n.right);
  }
  // (indices were assigned in queue order; children pushed after parents keep BFS order)
  for (size_t q = 0; q < bfs.size(); q++) index_of[size_t(bfs[q])] = int(q);
  for (size_t q = 0; q < bfs.size(); q++) {
    const TmpNode& n = tn[size_t(bfs[q])];
    const TmpNode& l = tn[size_t(n.left)];
    const TmpNode& r = tn[size_t(n.right)];
    int cl = l.left >= 0 ? index_of[size_t(n.left)] : encode_leaf(l.first, l.count);
    int cr = r.left >= 0 ? index_of[size_t(n.right)] : encode_leaf(r.first, r.count);
    pack_node(out, l.box, cl, r.box, cr);
  }
  // depth + SAH cost, for stats
  struct Rec { int n, depth; };
  std::vector<Rec> st{{root, 1}};
  const double root_area = std::max(tn[size_t(root)].box.area(), 1e-30);
  while (!st.empty()) {
    Rec r = st.back();
    st.pop_back();
    const TmpNode& n = tn[size_t(r.n)];
    out.bvh_depth = std::max(out.bvh_depth, r.depth);
    if (n.left >= 0) {
      out.sah_cost += n.box.area() / root_area;
      st.push_back({n.left, r.depth + 1});
      st.push_back({n.right, r.depth + 1});
    } else {
      out.sah_cost += n.box.area() / root_area * n.count;
    }
  }
  return true;
}

}  // namespace rtb200
#endif
