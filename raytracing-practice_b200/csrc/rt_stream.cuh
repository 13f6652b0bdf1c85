// rt_stream.cuh — the render path as a STREAMING persistent kernel: lanes are traversal engines fed from CTA-wide ray
// queues in shared memory; finished rays are shaded in dense batches of 32.
//
// Why (profiles/r15_render_lean.md, tools/simt_sim): in the megakernel a lane is welded to one path, so a warp's
// while-while traversal lasts as long as its LONGEST ray — 6.8 of 32 lanes active in node_step on the Book-2 final
// scene, 70 % of all issued instructions.  The CPU SIMT simulator (tools/simt_sim, same per-lane code, same BVH)
// says the cure is to refill idle lanes from a queue and to prefer the node step while enough lanes want one:
// 21 lanes in node_step, 1.4x fewer warp instructions per ray.  camera::render's loop body (camera.hpp:55-62,
// 180-232) is split into
//   TRACE  = world.hit (camera.hpp:192).  A lane holds ONE ray (origin, direction, closest hit so far, stack) in
//            registers; when >= RT_STREAM_REFILL lanes of the warp are idle, the finished ones publish their hit
//            into the path's slot, queue the slot for shading, and all idle lanes take the next rays from the
//            CTA's trace queue.  The warp takes a node step while >= RT_STREAM_NODE_THR lanes want one, else a
//            leaf step.
//   SHADE  = emitted + scatter + texture (camera.hpp:199-231), regeneration of ended paths from the global work
//            counter (camera.hpp:139-162) and the scene-enclosing media of the NEXT ray, for 32 slots at a time.
// Every warp does both: at a refill point it first looks at the shade queue and, if a full batch waits (or the
// tracers are starving), reserves it, suspends its unfinished traversals (their stacks stay where they are, in
// shared memory) and shades.  No block or grid barrier after start-up; queues are rings of 16-bit slot numbers
// with counters updated by one shared-memory atomic per warp-level operation.
//   pool   : n_slots paths per CTA, 64 B each, four float4 planes
//              {o.xyz, time} {d.xyz, start primitive} {hit t, hit ref, pixel, next sample} {beta.xyz, depth left}
//   stacks : one 32-bit entry per level and thread, [level][thread] (bank = lane: conflict-free):
//              (entry distance rounded DOWN to bf16) << 16 | 16-bit child code
//   BVH    : node records as three float4 planes + one (child, child) plane — a lane's 16-byte reads of random
//            nodes then spread over all banks (the 64-byte AoS record puts them on two 16-byte bank groups);
//            spheres as 16 B {centre, radius with the sign bit = "moving"}, boxes 48 B, leaf references 4 B
// RNG keys, sample order within a pixel and the fixed-point accumulation are exactly the megakernel's, so the
// accumulator is BIT-IDENTICAL to the megakernel's for the same (seed, sample range): tested.
#pragma once

namespace rtb200 {

#ifndef RT_STREAM_THREADS
#define RT_STREAM_THREADS 768
#endif
#ifndef RT_STREAM_NODE_THR
#define RT_STREAM_NODE_THR 10  // node step while at least this many lanes want one (tools/simt_sim: flat optimum 10..14)
#endif
#ifndef RT_STREAM_REFILL
#define RT_STREAM_REFILL 8  // publish / refill once at least this many lanes are idle
#endif
#ifndef RT_STREAM_SLOTS
#define RT_STREAM_SLOTS 1024  // paths per CTA (the simulator sees no difference between 832 and 1280 at 768 threads)
#endif
constexpr int kStreamThreads = RT_STREAM_THREADS;
constexpr uint32_t kStackStride = 4u * kStreamThreads;
constexpr uint32_t kNoSlot = 0xFFFFu;
constexpr uint32_t kRingEmpty = 0xFFFFu;
constexpr int kStreamMaxCode = 32767;  // child codes must fit 16 bits (signed): < 32768 nodes, < 4096 leaf references

// control words (int32) at StreamLayout::off_ctl
enum : uint32_t { SC_TQ_AVAIL = 0, SC_SQ_AVAIL = 4, SC_TQ_HEAD = 8, SC_TQ_TAIL = 12, SC_SQ_HEAD = 16, SC_SQ_TAIL = 20, SC_DEAD = 24, SC_BYTES = 64 };

// ---- shared-memory access through 32-bit shared-window addresses (see node_source() in rt_device.cuh) ----
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t x) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)x) : "memory"); }
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u16_volatile(uint32_t a) {
  unsigned short v;
  asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int lds_s32_volatile(uint32_t a) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int atoms_add(uint32_t a, int v) {
  int old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}

// Debug builds (-DRT_STREAM_WATCHDOG=<cycles>): every loop that waits on another warp gives up after that many cycles and
// leaves a record in counters[19..31] (printed by rt_get_stats under RT_B200_DEBUG) instead of hanging the GPU.
#ifndef RT_STREAM_WATCHDOG
#define RT_STREAM_WATCHDOG 0
#endif
struct StreamWatch {
  unsigned long long* counters;
  long long t0;
  uint32_t ctl;
};
__device__ __noinline__ void stream_watch_report(const StreamWatch& w, int where, int a, int b) {
  const unsigned long long k = atomicAdd(w.counters + 19, 1ull);
  if (k < 3) {
    unsigned long long* o = w.counters + 20 + 4 * k;
    int c[7];
    for (int i = 0; i < 7; i++) asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(c[i]) : "r"(w.ctl + 4u * i) : "memory");
    o[0] = (unsigned long long)(unsigned)where | ((unsigned long long)(threadIdx.x) << 8) | ((unsigned long long)blockIdx.x << 24) | ((unsigned long long)(unsigned)a << 40);
    o[1] = (unsigned long long)(unsigned)c[0] | ((unsigned long long)(unsigned)c[1] << 32);
    o[2] = (unsigned long long)(unsigned)(c[2] & 0xFFFF) | ((unsigned long long)(unsigned)(c[3] & 0xFFFF) << 16) | ((unsigned long long)(unsigned)(c[4] & 0xFFFF) << 32) |
           ((unsigned long long)(unsigned)(c[5] & 0xFFFF) << 48);
    o[3] = (unsigned long long)(unsigned)c[6] | ((unsigned long long)(unsigned)b << 32);
  }
}
#if RT_STREAM_WATCHDOG
#define RT_WATCH(w, where, a, b) (clock64() - (w).t0 > (long long)(RT_STREAM_WATCHDOG) ? (stream_watch_report(w, where, a, b), true) : false)
#else
#define RT_WATCH(w, where, a, b) false
#endif

// Adds the slots of the lanes in `m` to a ring.  Converged warp.  The payload a consumer will read (the slot's record)
// must have been written before the call: the fence orders it before the ring entry, which is what a consumer waits for.
__device__ __forceinline__ void ring_push(uint32_t ring, uint32_t ring_mask, uint32_t tail_addr, uint32_t avail_addr, unsigned m, bool has, uint32_t slot,
                                          unsigned lane) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int n = __popc(m), leader = __ffs(m) - 1;
  unsigned base = 0;
  if (int(lane) == leader) base = unsigned(atoms_add(tail_addr, n));
  base = __shfl_sync(FULL, base, leader);
  __threadfence_block();
  if (has) sts_u16(ring + 2u * ((base + __popc(m & ((1u << lane) - 1u))) & ring_mask), slot);
  if (int(lane) == leader) atoms_add(avail_addr, n);
}
// Reserves up to `want` entries (ONE lane calls this): returns how many, and the ring position of the first.  The counter
// may dip below zero for a moment (another consumer then simply gets nothing this time); what is taken is always given back.
__device__ __forceinline__ int ring_reserve(uint32_t avail_addr, uint32_t head_addr, int want, unsigned& base) {
  const int old = atoms_add(avail_addr, -want);
  const int got = old >= want ? want : max(old, 0);
  if (got < want) atoms_add(avail_addr, want - got);
  base = got ? unsigned(atoms_add(head_addr, got)) : 0u;
  return got;
}
// The slot number at ring position `pos` (reserved by this lane).  Its producer may still be between bumping the tail and
// writing the entry: wait for it (a few cycles; the producer never waits for anything).
__device__ __forceinline__ uint32_t ring_take(uint32_t ring, uint32_t ring_mask, unsigned pos, const StreamWatch& wd) {
  const uint32_t a = ring + 2u * (pos & ring_mask);
  uint32_t v;
  do {
    v = lds_u16_volatile(a);
    if (RT_WATCH(wd, 1, int(pos), int(ring))) return 0u;
  } while (v == kRingEmpty);
  sts_u16(a, kRingEmpty);
  __threadfence_block();  // the slot's record is read after its number
  return v;
}

// per-lane traversal engine of the streaming kernel
struct StreamRay {
  float3 o, d, inv, ood;
  Hit best;
};

// one BVH2 node for one lane: node_step (rt_device.cuh) on the planar node layout, with the far child going to the
// lane's shared-memory stack as (bf16 entry distance, 16-bit child code)
template <bool COUNT>
__device__ __forceinline__ void stream_pop(int& cur, uint32_t& sp, uint32_t my_stack, float best_t) {
  while (sp != my_stack) {
    sp -= kStackStride;
    const uint32_t e = lds_u32(sp);
    if (__uint_as_float(e & 0xFFFF0000u) <= best_t) {  // the stored distance is a lower bound of the real one: never culls a subtree the megakernel keeps
      cur = int(e << 16) >> 16;
      return;
    }
  }
  cur = kTravDone;
}
template <bool COUNT>
__device__ __forceinline__ void stream_node_step(const StreamRay& r, int& cur, uint32_t& sp, uint32_t my_stack, uint32_t nodes, uint32_t plane, unsigned int* cn) {
  const uint32_t p = nodes + 16u * uint32_t(cur);
  const float4 a = lds_f4(p), b = lds_f4(p + plane), c = lds_f4(p + 2u * plane);
  const uint2 ch = lds_v2(nodes + 3u * plane + 8u * uint32_t(cur));
  const int c0 = int(ch.x), c1 = int(ch.y);
  if (COUNT) cn[CN_NODE]++;
  const float3 inv = r.inv, ood = r.ood;
  const float tmin = 0.001f;  // camera.hpp:192
  float x0 = fmaf(a.x, inv.x, -ood.x), x1 = fmaf(a.w, inv.x, -ood.x);
  float y0 = fmaf(a.y, inv.y, -ood.y), y1 = fmaf(b.x, inv.y, -ood.y);
  float z0 = fmaf(a.z, inv.z, -ood.z), z1 = fmaf(b.y, inv.z, -ood.z);
  const float n0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
  const float f0 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), r.best.t));
  x0 = fmaf(b.z, inv.x, -ood.x), x1 = fmaf(c.y, inv.x, -ood.x);
  y0 = fmaf(b.w, inv.y, -ood.y), y1 = fmaf(c.z, inv.y, -ood.y);
  z0 = fmaf(c.x, inv.z, -ood.z), z1 = fmaf(c.w, inv.z, -ood.z);
  const float n1 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
  const float f1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), r.best.t));
  const bool h0 = n0 <= f0, h1 = n1 <= f1;
  if (h0 && h1) {
    const bool first0 = n0 <= n1;
    const uint32_t far = uint32_t(first0 ? c1 : c0), tf = __float_as_uint(first0 ? n1 : n0);
    sts_u32(sp, (tf & 0xFFFF0000u) | (far & 0xFFFFu));  // entry distances are >= tmin > 0: truncation rounds down
    sp += kStackStride;
    cur = first0 ? c0 : c1;
    return;
  }
  if (h0 || h1) {
    cur = h0 ? c0 : c1;
    return;
  }
  stream_pop<COUNT>(cur, sp, my_stack, r.best.t);
}

// leaf_body (rt_device.cuh) on the staged primitive copies; `aux` = {time, start primitive} of the ray, `key_of` its
// Philox counter (fetched only when a medium is sampled)
template <bool COUNT, typename KeyFn>
__device__ __forceinline__ void stream_leaf(StreamRay& r, int leaf, const DeviceScene& sc, uint32_t s_sph, uint32_t s_box, uint32_t s_refs, float time,
                                            uint32_t skip, KeyFn key_of, unsigned int* cn) {
  const int code = ~leaf;
  const int first = code >> 3, count = (code & 7) + 1;
  const float3 o = r.o, d = r.d;
  const float tmin = 0.001f;
  for (int k = 0; k < count; k++) {
    uint32_t ref = lds_u32(s_refs + 4u * uint32_t(first + k));
    const uint32_t type = ref >> 30, idx = ref & 0x3FFFFFFFu;
    float t = -1.0f;
    if (type == REF_SPHERE) {
      const float4 g0 = lds_f4(s_sph + 16u * idx);
      float4 g1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (__float_as_int(g0.w) < 0) g1 = __ldg(sc.spheres + 2 * idx + 1);  // a moving sphere (sign bit of the staged radius; only r^2 is used here)
      t = hit_sphere(g0, g1, o, d, time, tmin, r.best.t, ref == skip);
      if (COUNT) cn[CN_SPH]++, cn[CN_SPH_HIT] += t != -1.0f;
    } else if (type == REF_QUAD) {
      if (ref != skip) {
        t = hit_quad(__ldg(sc.quads + 3 * idx), __ldg(sc.quads + 3 * idx + 1), __ldg(sc.quads + 3 * idx + 2), o, d, tmin, r.best.t);
        if (COUNT) {
          float4 nD = __ldg(sc.quads + 3 * idx);
          float den = dot(xyz(nD), d), tq = (nD.w - dot(xyz(nD), o)) / den;
          cn[CN_QUAD]++, cn[CN_QUAD_FULL] += (fabsf(den) >= 1e-8f && tq >= tmin && tq <= r.best.t) || t != -1.0f;
        }
      }
    } else if (type == REF_BOX) {
      if (ref != REF_NONE) {
        const uint32_t b = idx >> 3;
        const int self_face = ((skip >> 30) == REF_BOX && skip != REF_NONE && ((skip & 0x3FFFFFFFu) >> 3) == b) ? int(skip & 7u) : -1;
        int face = 0;
        const float4 b0 = lds_f4(s_box + 48u * b), b1 = lds_f4(s_box + 48u * b + 16u), b2 = lds_f4(s_box + 48u * b + 32u);
        t = hit_box(b0, b1, b2, o, d, r.inv, r.ood, tmin, r.best.t, self_face, face);
        if (COUNT) cn[CN_BOX]++;
        ref = make_ref(REF_BOX, (b << 3) | uint32_t(face));
      }
    } else {
      const DMedium m = sc.media[idx];
      PathKey key;
      uint32_t bounce;
      key_of(key, bounce);
      t = medium_sample(sc, m, int(idx), o, d, time, tmin, r.best.t, key, bounce);
      if (COUNT) cn[CN_MEDIUM]++;
    }
    if (t != -1.0f) r.best = Hit{t, ref};
  }
}

// SHADE + REGENERATE for one slot (outlined: its registers must not compete with the traversal loop's): the tail of one
// ray_color level for the finished query in the slot, then — if the path ended, or the slot is fresh — the next camera
// sample, then the scene-enclosing media of the new ray (the seed of its closest hit).  Returns bit 0 = the slot holds a
// ray to trace (else the image has no samples left for it), bit 1 = a query was shaded (one ray of rt_stats.rays).
// `Pp` = the CTA's shared-memory copy of the parameters.
template <bool COUNT>
__device__ __noinline__ int stream_shade_slot(const RenderParams* __restrict__ Pp, uint32_t rec, uint32_t plane, unsigned int* cn) {
  const RenderParams& P = *Pp;
  const DeviceScene& sc = P.sc;
  const float INF = __int_as_float(0x7f800000);
  const float4 R2 = lds_f4(rec + 2u * plane), R3 = lds_f4(rec + 3u * plane);
  int pixel = __float_as_int(R2.z), s = __float_as_int(R2.w), depth = __float_as_int(R3.w);
  float3 o, d, beta;
  float time;
  uint32_t skip = REF_NONE;
  bool alive = false;
  const int shaded = depth > 0 ? 2 : 0;
  if (depth > 0) {  // ---- one segment of ray_color (camera.hpp:180-232), as in render_kernel ----
    const float4 R0 = lds_f4(rec), R1 = lds_f4(rec + plane);
    o = f3(R0.x, R0.y, R0.z), d = f3(R1.x, R1.y, R1.z), time = R0.w;
    beta = f3(R3.x, R3.y, R3.z);
    const Hit h{R2.x, __float_as_uint(R2.y)};
    const PathKey key{P.key, uint32_t(pixel), uint32_t(s - 1)};
    const uint32_t bounce = uint32_t(P.cam.max_depth - depth) + 1u;
    float3 L = f3(0.0f, 0.0f, 0.0f);
    if (h.ref == REF_NONE) {
      L = L + beta * P.cam.bg;
    } else {
      const uint4 rnd = rng_block(key, bounce, 0u);
      Surface sf = surface_at(sc, h, o, d, time);
      float3 emit, atten, d_out;
      const bool cont = scatter_ray<COUNT>(sc, sf, d, rnd, emit, atten, d_out, cn);
      L = L + beta * emit;
      if (cont) {
        beta = beta * atten;
        o = sf.p;
        d = d_out;
        skip = (h.ref >> 30) == REF_MEDIUM ? REF_NONE : h.ref;
        alive = --depth > 0;
      }
    }
    if (!alive) {
      RT_CHECK(pixel >= 0 && (unsigned long long)pixel * 3ull + 2ull < P.n_values, CHK_PIXEL);
      unsigned long long* dst = P.accum + 3ull * (unsigned long long)pixel;
      const long long fr = to_fixed(L.x), fg = to_fixed(L.y), fb = to_fixed(L.z);
      if (fr) atomicAdd(dst + 0, (unsigned long long)fr);
      if (fg) atomicAdd(dst + 1, (unsigned long long)fg);
      if (fb) atomicAdd(dst + 2, (unsigned long long)fb);
    }
  }
  if (!alive) {  // ---- the next sample of this slot's work item, or the next item (render_kernel's regeneration) ----
    const int s_last = P.sample_begin + P.sample_count;
    if ((((unsigned)(s - P.sample_begin)) & (unsigned)(P.chunk - 1)) == 0u || s >= s_last) {
      bool have = false;
      for (;;) {
        const unsigned long long it = atomicAdd(P.counters, 1ull);
        if (it >= (unsigned long long)P.n_items) break;
        const unsigned int item = (unsigned int)it;
        const unsigned int chunk = item / P.per_chunk, q = item - chunk * P.per_chunk;
        const unsigned int tile = q >> 5, l = q & 31u;
        const int px = int(tile % (unsigned)P.tiles_x) * 8 + int(l & 7u);
        const int py = int(tile / (unsigned)P.tiles_x) * 4 + int(l >> 3);
        if (px < P.cam.W && py < P.cam.H) {
          s = P.sample_begin + int(chunk) * P.chunk;
          if (s < s_last) {
            pixel = py * P.cam.W + px;
            have = true;
            break;
          }
        }
      }
      if (!have) {
        sts_f4(rec + 3u * plane, make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0)));
        return shaded;
      }
    }
    // camera::get_ray (camera.hpp:139-162): jitter, defocus disk, shutter time
    const PathKey key{P.key, uint32_t(pixel), uint32_t(s++)};
    const int py = pixel / P.cam.W, px = pixel - py * P.cam.W;
    const uint4 r0 = rng_block(key, 0u, 0u);
    const float ox = u01(r0.x) - 0.5f, oy = u01(r0.y) - 0.5f;
    time = u01(r0.z);
    float3 dir = fma3(float(px) + ox, P.cam.du, fma3(float(py) + oy, P.cam.dv, P.cam.p00c));
    o = P.cam.center;
    if (P.cam.defocus) {
      const uint4 r1 = rng_block(key, 0u, 1u);
      float rr = sqrtf(u01(r1.x)), sn, cs;
      sincos_2pi(u01(r1.y), sn, cs);
      const float3 off = fma3(rr * cs, P.cam.ddu, (rr * sn) * P.cam.ddv);
      o = o + off;
      dir = dir - off;
    }
    d = dir;
    beta = f3(1.0f, 1.0f, 1.0f);
    depth = P.cam.max_depth;
    skip = REF_NONE;
  }
  // world.hit, part 1: the scene-enclosing media (met by every ray) are sampled here, 32 slots wide
  Hit best{INF, REF_NONE};
  if (sc.n_global_media) {
    const PathKey key{P.key, uint32_t(pixel), uint32_t(s - 1)};
    const uint32_t bounce = uint32_t(P.cam.max_depth - depth) + 1u;
    best = sample_global_media<COUNT>(sc, o, d, time, 0.001f, INF, key, bounce, cn);
  }
  sts_f4(rec, make_float4(o.x, o.y, o.z, time));
  sts_f4(rec + plane, make_float4(d.x, d.y, d.z, __uint_as_float(skip)));
  sts_f4(rec + 2u * plane, make_float4(best.t, __uint_as_float(best.ref), __int_as_float(pixel), __int_as_float(s)));
  sts_f4(rec + 3u * plane, make_float4(beta.x, beta.y, beta.z, __int_as_float(depth)));
  return shaded | 1;
}

template <bool COUNT>
__global__ void __launch_bounds__(kStreamThreads, 1) stream_kernel(const __grid_constant__ RenderParams P) {
  extern __shared__ float4 s_dyn[];
  __shared__ RenderParams sP;  // the outlined shade function cannot address the kernel's constant bank
  for (int i = threadIdx.x; i < int(sizeof(RenderParams) / 4); i += blockDim.x) reinterpret_cast<int*>(&sP)[i] = reinterpret_cast<const int*>(&P)[i];
  const StreamLayout& SL = P.sl;
  const unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t s_base = uint32_t(__cvta_generic_to_shared(s_dyn));
  // ---- stage the BVH: node planes, spheres (16 B), boxes, leaf references ----
  {
    const int n = P.sc.n_nodes;
    float4* pa = s_dyn;
    float4* pb = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_dyn) + SL.node_plane);
    float4* pc = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_dyn) + 2u * SL.node_plane);
    uint2* pd = reinterpret_cast<uint2*>(reinterpret_cast<char*>(s_dyn) + 3u * SL.node_plane);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      pa[i] = P.sc.nodes[4 * i], pb[i] = P.sc.nodes[4 * i + 1], pc[i] = P.sc.nodes[4 * i + 2];
      const float4 dd = P.sc.nodes[4 * i + 3];
      pd[i] = make_uint2(__float_as_uint(dd.x), __float_as_uint(dd.y));
    }
    float4* s_sph = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_dyn) + SL.off_sph);
    for (int i = threadIdx.x; i < P.sc.n_spheres; i += blockDim.x) {
      float4 g0 = P.sc.spheres[2 * i];
      const float4 g1 = P.sc.spheres[2 * i + 1];
      g0.w = fabsf(g0.w);  // the hit test only uses r^2; the sign bit marks a moving sphere
      if (g1.x != 0.0f || g1.y != 0.0f || g1.z != 0.0f) g0.w = -g0.w;
      s_sph[i] = g0;
    }
    float4* s_box = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_dyn) + SL.off_box);
    for (int i = threadIdx.x; i < 3 * P.sc.n_boxes; i += blockDim.x) s_box[i] = P.sc.boxes[i];
    uint32_t* s_ref = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(s_dyn) + SL.off_refs);
    for (int i = threadIdx.x; i < P.sc.n_leaf_refs; i += blockDim.x) s_ref[i] = P.sc.leaf_refs[i];
    // ---- queues: every slot starts in the shade queue as a fresh path (depth 0: its first shade is a regeneration) ----
    unsigned short* tq = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(s_dyn) + SL.off_tq);
    unsigned short* sq = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(s_dyn) + SL.off_sq);
    for (unsigned i = threadIdx.x; i <= SL.ring_mask; i += blockDim.x) tq[i] = (unsigned short)kRingEmpty, sq[i] = (unsigned short)(i < SL.n_slots ? i : kRingEmpty);
    float4* r2 = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_dyn) + SL.off_slots + 2u * SL.slot_plane);
    float4* r3 = reinterpret_cast<float4*>(reinterpret_cast<char*>(s_dyn) + SL.off_slots + 3u * SL.slot_plane);
    for (unsigned i = threadIdx.x; i < SL.n_slots; i += blockDim.x) {
      r2[i] = make_float4(0.0f, __uint_as_float(REF_NONE), __int_as_float(-1), __int_as_float(P.sample_begin));
      r3[i] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0));
    }
    int* ctl = reinterpret_cast<int*>(reinterpret_cast<char*>(s_dyn) + SL.off_ctl);
    if (threadIdx.x < SC_BYTES / 4) ctl[threadIdx.x] = (threadIdx.x == SC_SQ_AVAIL / 4 || threadIdx.x == SC_SQ_TAIL / 4) ? int(SL.n_slots) : 0;
  }
  __syncthreads();
  // shared-window addresses, laundered so that they live in registers (see node_source())
  const uint32_t s_nodes = opaque_u32(s_base), node_plane = SL.node_plane;
  const uint32_t s_sph = opaque_u32(s_base + SL.off_sph), s_box = opaque_u32(s_base + SL.off_box), s_refs = opaque_u32(s_base + SL.off_refs);
  const uint32_t s_slots = opaque_u32(s_base + SL.off_slots), slot_plane = SL.slot_plane;
  const uint32_t s_tq = s_base + SL.off_tq, s_sq = s_base + SL.off_sq, ring_mask = SL.ring_mask, s_ctl = opaque_u32(s_base + SL.off_ctl);
  const uint32_t my_stack = s_base + SL.off_stack + 4u * threadIdx.x;
  const DeviceScene& sc = P.sc;

  unsigned int cn[COUNT ? CN_COUNT : 1];
  if (COUNT)
    for (int i = 0; i < CN_COUNT; i++) cn[i] = 0;
  unsigned int n_rays = 0;  // warp-uniform
  // the lane's traversal engine; (cur, sp, slot) survive a SHADE round, the ray itself is re-read from the slot
  int cur = kTravDone;
  uint32_t sp = my_stack, slot = kNoSlot;
  StreamRay r;
  r.o = r.d = r.inv = r.ood = f3(0.0f, 0.0f, 1.0f);
  r.best = Hit{0.0f, REF_NONE};

  const StreamWatch wd{P.counters, clock64(), s_ctl};
  for (;;) {
    // =========================== TRACE ===========================
    unsigned batch_base = 0;
    int batch_n = 0;
    bool out_of_work = false;
    for (;;) {
      if (RT_WATCH(wd, 2, cur, int(slot))) return;
      const unsigned bn = __ballot_sync(FULL, cur >= 0);
      const int nn = __popc(bn);
      if (nn >= RT_STREAM_NODE_THR) {
        if (cur >= 0) stream_node_step<COUNT>(r, cur, sp, my_stack, s_nodes, node_plane, cn);
        continue;
      }
      unsigned busy = __ballot_sync(FULL, cur != kTravDone);
      if (busy & ~bn) {  // some lanes sit on a leaf and too few want a node step: intersect the leaves
        if (cur < 0 && cur != kTravDone) {
          const float time = __uint_as_float(lds_u32(s_slots + 16u * slot + 12u));
          const uint32_t skip = lds_u32(s_slots + slot_plane + 16u * slot + 12u);
          stream_leaf<COUNT>(r, cur, sc, s_sph, s_box, s_refs, time, skip,
                             [&](PathKey& k, uint32_t& b) {
                               const uint2 ps = lds_v2(s_slots + 2u * slot_plane + 16u * slot + 8u);
                               k = PathKey{P.key, ps.x, ps.y - 1u};
                               b = uint32_t(P.cam.max_depth - int(lds_u32(s_slots + 3u * slot_plane + 16u * slot + 12u))) + 1u;
                             },
                             cn);
          stream_pop<COUNT>(cur, sp, my_stack, r.best.t);
        }
        busy = __ballot_sync(FULL, cur != kTravDone);
      } else if (nn) {  // only node work around
        if (cur >= 0) stream_node_step<COUNT>(r, cur, sp, my_stack, s_nodes, node_plane, cn);
        continue;
      }
      const int n_idle = 32 - __popc(busy);
      if (n_idle < RT_STREAM_REFILL && busy != 0u) continue;
      // ---- refill point: publish finished rays, look at the shade queue, take new rays ----
      const bool fin = cur == kTravDone && slot != kNoSlot;
      const unsigned bf = __ballot_sync(FULL, fin);
      if (bf) {
        if (fin) sts_v2(s_slots + 2u * slot_plane + 16u * slot, __float_as_uint(r.best.t), r.best.ref);
        ring_push(s_sq, ring_mask, s_ctl + SC_SQ_TAIL, s_ctl + SC_SQ_AVAIL, bf, fin, slot, lane);
        if (fin) slot = kNoSlot;
      }
      // lane 0 decides and reserves — a shade batch if a full one waits or the tracers are starving, else rays for the idle
      // lanes — and tells the others: (kind << 8 | count, first ring position)
      unsigned q_code = 0u, q_base = 0u;
      if (lane == 0) {
        const int tq_avail = lds_s32_volatile(s_ctl + SC_TQ_AVAIL), sq_avail = lds_s32_volatile(s_ctl + SC_SQ_AVAIL);
        int n = 0;
        if (sq_avail >= 32 || (sq_avail > 0 && tq_avail <= 0)) {
          n = ring_reserve(s_ctl + SC_SQ_AVAIL, s_ctl + SC_SQ_HEAD, min(sq_avail, 32), q_base);
          if (n) q_code = 0x100u | unsigned(n);
        }
        if (n == 0 && tq_avail > 0) {
          n = ring_reserve(s_ctl + SC_TQ_AVAIL, s_ctl + SC_TQ_HEAD, min(tq_avail, n_idle), q_base);
          q_code = unsigned(n);
        }
      }
      q_code = __shfl_sync(FULL, q_code, 0);
      q_base = __shfl_sync(FULL, q_base, 0);
      if (q_code & 0x100u) {
        batch_n = int(q_code & 0xFFu), batch_base = q_base;
        break;
      }
      const int got = int(q_code);
      {
        const unsigned base = q_base;
        const int rank = __popc(~busy & ((1u << lane) - 1u));
        if (cur == kTravDone && rank < got) {
          slot = ring_take(s_tq, ring_mask, base + unsigned(rank), wd);
          const float4 R0 = lds_f4(s_slots + 16u * slot), R1 = lds_f4(s_slots + slot_plane + 16u * slot);
          const uint2 R2 = lds_v2(s_slots + 2u * slot_plane + 16u * slot);
          r.o = f3(R0.x, R0.y, R0.z), r.d = f3(R1.x, R1.y, R1.z);
          r.inv = f3(fabsf(r.d.x) > 1e-30f ? rcp_fast(r.d.x) : copysignf(1e30f, r.d.x), fabsf(r.d.y) > 1e-30f ? rcp_fast(r.d.y) : copysignf(1e30f, r.d.y),
                     fabsf(r.d.z) > 1e-30f ? rcp_fast(r.d.z) : copysignf(1e30f, r.d.z));
          r.ood = r.o * r.inv;
          r.best = Hit{__uint_as_float(R2.x), R2.y};
          sp = my_stack;
          cur = 0;
        }
      }
      if (busy == 0u && got == 0) {  // nothing in flight in this warp and nothing to take
        out_of_work = true;
        break;
      }
    }
    if (batch_n) {
      // =========================== SHADE ===========================
      // suspend: the closest hits so far go to the slots, the stacks stay in place, (cur, sp, slot) stay in registers
      if (cur != kTravDone) sts_v2(s_slots + 2u * slot_plane + 16u * slot, __float_as_uint(r.best.t), r.best.ref);
      int code = -1;  // -1: this lane had no slot
      uint32_t bslot = 0;
      if (int(lane) < batch_n) {
        bslot = ring_take(s_sq, ring_mask, batch_base + lane, wd);
        code = stream_shade_slot<COUNT>(&sP, s_slots + 16u * bslot, slot_plane, cn);
      }
      const bool has_ray = code >= 0 && (code & 1);
      const unsigned bR = __ballot_sync(FULL, has_ray), bD = __ballot_sync(FULL, code >= 0 && !(code & 1));
      n_rays += __popc(__ballot_sync(FULL, code >= 0 && (code & 2)));
      if (bR) ring_push(s_tq, ring_mask, s_ctl + SC_TQ_TAIL, s_ctl + SC_TQ_AVAIL, bR, has_ray, bslot, lane);
      if (bD && lane == 0) atoms_add(s_ctl + SC_DEAD, __popc(bD));
      // resume: the suspended rays come back from their slots
      if (cur != kTravDone) {
        const float4 R0 = lds_f4(s_slots + 16u * slot), R1 = lds_f4(s_slots + slot_plane + 16u * slot);
        const uint2 R2 = lds_v2(s_slots + 2u * slot_plane + 16u * slot);
        r.o = f3(R0.x, R0.y, R0.z), r.d = f3(R1.x, R1.y, R1.z);
        r.inv = f3(fabsf(r.d.x) > 1e-30f ? rcp_fast(r.d.x) : copysignf(1e30f, r.d.x), fabsf(r.d.y) > 1e-30f ? rcp_fast(r.d.y) : copysignf(1e30f, r.d.y),
                   fabsf(r.d.z) > 1e-30f ? rcp_fast(r.d.z) : copysignf(1e30f, r.d.z));
        r.ood = r.o * r.inv;
        r.best = Hit{__uint_as_float(R2.x), R2.y};
      }
      continue;
    }
    if (out_of_work) {
      if (RT_WATCH(wd, 3, 0, 0)) return;
      if (lds_s32_volatile(s_ctl + SC_DEAD) >= int(SL.n_slots)) break;  // every path of this CTA has run out of samples
      __nanosleep(256);
    }
  }
  // ---- counters: one atomic per warp ----
  if (lane == 0 && n_rays) atomicAdd(P.counters + 1, (unsigned long long)n_rays);
  if (COUNT)
    for (int i = 0; i < CN_COUNT; i++)
      if (cn[i]) atomicAdd(P.counters + 4 + i, (unsigned long long)cn[i]);
}

}  // namespace rtb200
