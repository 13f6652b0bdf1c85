// device_scene.h — POD layouts shared by the host scene converter and the CUDA kernels.
//
// Data layout in HBM (everything is tiny and L2/L1/shared resident):
//   nodes      : BVH2, one 64-byte record per INTERNAL node holding BOTH children's boxes
//                (4 x float4, 64-byte aligned), breadth-first order so that a prefix of the
//                array = the top levels of the tree (staged into shared memory per CTA).
//   leaf_refs  : uint32 per leaf slot: (type << 30) | index,  type 0 sphere / 1 quad / 2 medium
//   spheres    : 2 x float4 {cx,cy,cz,r} {dcx,dcy,dcz,-}           (world space, baked)
//   quads      : 3 x float4 {n,D} {A,a0} {B,b0}: t=(D-n.o)/(n.d), alpha=A.p+a0, beta=B.p+b0
//   boxes      : 3 x float4 {lo,cos} {hi,sin} {t,flag}: the six quads of box() (quad.hpp:129-159) as ONE slab-test
//                primitive in the box's own frame (a translate/rotate_y instance is one 2x2 rotation of the ray);
//                box_meta {material, first quad, face map, -}: face -> which of its quad records (uv, id) it is
//   materials  : 2 x float4; textures: 2 x float4; texels RGBA8; perlin tables float4[256]+perm
// The fp64 "exact" arrays (X*) keep the reference's OBJECT-space doubles and the instance
// transform chain, so the parity harness can re-evaluate a candidate exactly as
// sphere::hit / quad::hit / translate::hit would (no baking, reference operation order).
#ifndef RTB200_DEVICE_SCENE_H
#define RTB200_DEVICE_SCENE_H

#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#else
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
#endif

namespace rtb200 {

// a box reference carries (box index << 3) | face in its 30 index bits; face = axis * 2 + (hi side), 7 = "the box"
enum : uint32_t { REF_SPHERE = 0u, REF_QUAD = 1u, REF_MEDIUM = 2u, REF_BOX = 3u, REF_NONE = 0xFFFFFFFFu };
inline
#ifdef __CUDACC__
    __host__ __device__
#endif
    uint32_t make_ref(uint32_t type, uint32_t index) { return (type << 30) | index; }

// child encoding inside a node: >= 0 internal node index; < 0 leaf: ~((first << 3) | (count-1))
// an EMPTY child has an inverted box (min=+inf, max=-inf) and is never entered.
constexpr int kMaxLeaf = 6;  // <= 8 (3 bits in the leaf code)

struct DMedium {
  float neg_inv_density;
  int material;   // isotropic phase function
  int first_bref; // into medium_brefs
  int n_bref;
};

enum : int { MAT_LAMBERTIAN = 1, MAT_METAL = 2, MAT_DIELECTRIC = 3, MAT_LIGHT = 4, MAT_ISOTROPIC = 5 };
enum : int { TEX_SOLID = 1, TEX_CHECKER = 2, TEX_IMAGE = 3, TEX_NOISE = 4 };
enum : int { MATF_NEEDS_UV = 1 };

// ---- fp64 exact-predicate data (reference object space) --------------------------------
struct XSphere {
  double c[3], dc[3], r;
  int chain, order, pid, pad;
};
struct XQuad {
  double Q[3], u[3], v[3], n[3], w[3], D;  // n, w, D derived on the host as quad.hpp:17-23
  int chain, order, pid, pad;
};
struct XOp {  // one instance wrapper, outermost first
  int kind;   // 0 translate (a = offset), 1 rotate_y (a[0] = sin, a[1] = cos)
  int pad;
  double a[3];
};

struct DeviceScene {
  const float4* nodes;
  const uint32_t* leaf_refs;
  const float4* spheres;
  const int2* sph_meta;  // {material, rotation index or -1}
  const float4* quads;
  const int* quad_mat;
  const float4* boxes;
  const int4* box_meta;  // {material, first quad record, face map: 4 bits per face = quad offset | inward-normal << 3, 0}
  const DMedium* media;
  const uint32_t* medium_brefs;
  const float4* materials;
  const float4* textures;
  const uchar4* texels;
  const int4* images;  // {texel offset, width, height, 0}
  const float4* perlin_vec;
  const uint8_t* perlin_perm;
  const float2* rotations;  // {sin, cos} of the accumulated rotate_y of an instance
  const XSphere* xspheres;
  const XQuad* xquads;
  const XOp* xops;
  const int2* xchains;  // {first op, n ops}
  int n_nodes, n_spheres, n_quads, n_media, n_materials, n_textures, n_boxes, n_leaf_refs;
  int n_global_media;   // media that enclose the whole scene: sampled once per ray, not via the BVH
  int n_noise;          // noise_texture tables (perlin_vec / perlin_perm): 0 = the scene has no Perlin texture
  int global_media[4];
  float scene_abs_max;  // max |coordinate| of any finite bound (conservative-cull epsilon scale)
  float bounds_lo[3], bounds_hi[3];  // union of every BVH item's box (= what node 0 covers); empty scene: +inf / -inf
};

}  // namespace rtb200
#endif
