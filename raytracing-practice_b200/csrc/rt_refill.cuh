// rt_refill.cuh — the render megakernel with IN-PLACE LANE REFILL.
//
// render_kernel (rt_b200.cu) runs "regenerate + trace one segment + shade" in lock step: the warp's while-while
// traversal lasts as long as its LONGEST ray (6.8 of 32 lanes active in node_step on the Book-2 final scene, 70 % of
// the issued instructions, profiles/r15_render_lean.md).  Here a lane still owns its path — no queues, no atomics, no
// path records moving between lanes — but the warp LEAVES the traversal as soon as RT_REFILL_SHADE_THR of its lanes
// have their answer: those lanes shade, regenerate and seed their next ray while the unfinished traversals stay
// suspended (node index and stack pointer in two registers, the stack in local memory where it already was), and then
// everybody traverses again.  Inside the traversal the warp takes a node step while at least RT_REFILL_NODE_THR lanes
// want one and a leaf step otherwise.  tools/simt_sim (policy `inplace`) predicted 16-18 lanes in node_step and 1.3x
// fewer warp instructions per ray on book2_final for thresholds 16..24 / 8.
//
// MEASURED (profiles/r2_session_experiments.md, gpurun_out/refill_ab1.log, prof_rf1): bit-identical to the megakernel,
// node_step at 13.9 lanes (6.8 there), traversal instructions -28 % — and 0.83x the megakernel's speed (0.85x with node
// threshold 1, 0.90x with shade threshold 32 = the megakernel's own schedule in this loop shape): the shade round costs
// about the same whether 20 or 31 lanes take it and now runs 1.45x as often (+70 % shade instructions), leaf phases — the
// expensive kind of traversal step — multiply, and issue utilisation falls from 75 % to 62 %.  Kept as an opt-in,
// tested experiment (RT_RENDER_REFILL); the production kernel is render_kernel.
//
// The register rule of the megakernel (anything live across the traversal or across the shade is paid for in occupancy)
// is kept by PARKING: a ray's origin, direction and closest hit so far live in two more float4 records per thread in
// shared memory ({o, best.t}, {d, best.ref}) next to the megakernel's {beta, depth} and {pixel, s, time, skip}; every
// lane stores best.t / best.ref when the warp leaves the traversal and reloads the ray (and rebuilds 1/d, o/d) when it
// re-enters, so that the shade code sees none of the traversal's registers and vice versa.
//
// Per-ray arithmetic, RNG counters and the fixed-point accumulation are the megakernel's, and a lane's sequence of node
// and leaf steps does not depend on what the other lanes do, so the accumulator is BIT-IDENTICAL to render_kernel's
// for the same (seed, sample range): tested (tests/test_gpu_render.py).
// Reference: camera::render / ray_color (src/core/camera.hpp:29-72, 180-232), bvh_node::hit (accelerator/bvh_node.hpp:80-94).
#pragma once

namespace rtb200 {

#ifndef RT_REFILL_SHADE_THR
#define RT_REFILL_SHADE_THR 20  // leave the traversal once this many lanes wait for a shade
#endif
#ifndef RT_REFILL_NODE_THR
#define RT_REFILL_NODE_THR 1  // node step while at least this many lanes want one, else the lanes on a leaf go first (8 measured slower than 1: leaf steps are the expensive ones, gpurun_out/refill_ab1.log, ab_thr.log)
#endif
#ifndef RT_REFILL_THREADS
#define RT_REFILL_THREADS RT_THREADS
#endif
constexpr int kRefillThreads = RT_REFILL_THREADS;
constexpr size_t kRefillStateBytes = size_t(64) * kRefillThreads;  // four float4 records per thread

template <bool COUNT, bool ALL_SMEM>
__global__ void __launch_bounds__(kRefillThreads, 1) refill_kernel(const __grid_constant__ RenderParams P) {
  extern __shared__ float4 s_nodes[];
  stage_nodes<ALL_SMEM>(s_nodes, P.sc.nodes, P.smem_nodes);
  LeafSource ls{0u, 0u, 0u};
  if (ALL_SMEM) {  // [nodes][spheres 2 x float4][boxes 3 x float4][leaf refs u32], as render_kernel
    float4* s_sph = s_nodes + 4 * P.smem_nodes;
    float4* s_box = s_sph + 2 * P.sc.n_spheres;
    uint32_t* s_ref = reinterpret_cast<uint32_t*>(s_box + 3 * P.sc.n_boxes);
    for (int i = threadIdx.x; i < 2 * P.sc.n_spheres; i += blockDim.x) s_sph[i] = P.sc.spheres[i];
    for (int i = threadIdx.x; i < 3 * P.sc.n_boxes; i += blockDim.x) s_box[i] = P.sc.boxes[i];
    for (int i = threadIdx.x; i < P.sc.n_leaf_refs; i += blockDim.x) s_ref[i] = P.sc.leaf_refs[i];
    ls.spheres = opaque_u32(uint32_t(__cvta_generic_to_shared(s_sph)));
    ls.boxes = opaque_u32(uint32_t(__cvta_generic_to_shared(s_box)));
    ls.refs = opaque_u32(uint32_t(__cvta_generic_to_shared(s_ref)));
  }
  __syncthreads();
  const NodeSource ns = node_source(s_nodes, P.sc.nodes, P.smem_nodes, P.sc.n_nodes);
  const DeviceScene& sc = P.sc;
  const float INF = __int_as_float(0x7f800000);
  const unsigned FULL = 0xFFFFFFFFu;

  // per-thread records in shared memory: A {beta.xyz, depth}  B {pixel, next sample, time, start primitive}
  //                                       C {o.xyz, best.t}    D {d.xyz, best.ref}
  const uint32_t st_a = opaque_u32(uint32_t(__cvta_generic_to_shared(s_nodes)) + P.state_off) + 16u * threadIdx.x;
  constexpr uint32_t kStB = 16u * kRefillThreads, kStC = 2u * kStB, kStD = 3u * kStB;
  auto sts_f4 = [](uint32_t addr, float4 v) { asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); };
  auto sts_b32 = [](uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); };
  auto ldv_f4 = [](uint32_t addr) {  // volatile + memory clobber: never merged with an earlier load, never carried in registers
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
  };
  sts_f4(st_a, make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0)));
  sts_f4(st_a + kStB, make_float4(__int_as_float(-1), __int_as_float(P.sample_begin), 0.0f, __uint_as_float(REF_NONE)));
  sts_f4(st_a + kStC, make_float4(0.0f, 0.0f, 0.0f, INF));
  sts_f4(st_a + kStD, make_float4(0.0f, 0.0f, 1.0f, __uint_as_float(REF_NONE)));
  const int s_last = P.sample_begin + P.sample_count;
  const bool media = sc.n_media != 0;
  unsigned int n_rays = 0;  // warp-uniform
  unsigned int cn[COUNT ? CN_COUNT : 1];
  if (COUNT)
    for (int i = 0; i < CN_COUNT; i++) cn[i] = 0;

  // what a lane carries in registers across BOTH phases: the traversal's position and two flags
  TravState ts;
  TravStack st;
  ts.cur = kTravDone;
  trav_reset(ts, st);
  bool alive = false;  // the lane's path has a ray (being traced, or traced and waiting for its shade)
  bool done = false;   // the image has no samples left for this lane

  auto key_of = [&](PathKey& k, uint32_t& b) {
    const float4 A = lds_f4(st_a), B = lds_f4(st_a + kStB);
    k = PathKey{P.key, uint32_t(__float_as_int(B.x)), uint32_t(__float_as_int(B.y) - 1)};
    b = uint32_t(P.cam.max_depth - __float_as_int(A.w)) + 1u;
  };
  auto aux_of = [&](float& t, uint32_t& sk) {
    float tt, ss;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(tt), "=f"(ss) : "r"(st_a + kStB + 8u));
    t = tt, sk = __float_as_uint(ss);
  };

  for (;;) {
    // =============================== SHADE + REGENERATE + SEED ===============================
    // lanes whose traversal is over (ts.cur == kTravDone) and that are not out of work
    const bool mine = ts.cur == kTravDone && !done;
    n_rays += __popc(__ballot_sync(FULL, mine && alive));
    if (mine) {
      float3 o, d;
      float time;
      if (alive) {  // ---- one segment of ray_color (camera.hpp:180-232): render_kernel's shade ----
        const float4 A = ldv_f4(st_a), B = ldv_f4(st_a + kStB), Cc = ldv_f4(st_a + kStC), Dd = ldv_f4(st_a + kStD);
        float3 beta = f3(A.x, A.y, A.z);
        int depth = __float_as_int(A.w);
        const int pixel = __float_as_int(B.x);
        time = B.z;
        o = f3(Cc.x, Cc.y, Cc.z), d = f3(Dd.x, Dd.y, Dd.z);
        const Hit h{Cc.w, __float_as_uint(Dd.w)};
        const PathKey key{P.key, uint32_t(pixel), uint32_t(__float_as_int(B.y) - 1)};
        const uint32_t bounce = uint32_t(P.cam.max_depth - depth) + 1u;
        float3 L = f3(0.0f, 0.0f, 0.0f);
        if (h.ref == REF_NONE) {
          L = L + beta * P.cam.bg;
          alive = false;
        } else {
          const uint4 rnd = rng_block(key, bounce, 0u);
          Surface sf = surface_at(sc, h, o, d, time);
          float3 emit, atten, d_out;
          bool cont = scatter_ray<COUNT>(sc, sf, d, rnd, emit, atten, d_out, cn);
          L = L + beta * emit;
          if (cont) {
            beta = beta * atten;
            o = sf.p;
            d = d_out;
            const uint32_t skip = (h.ref >> 30) == REF_MEDIUM ? REF_NONE : h.ref;
            alive = --depth > 0;
            sts_f4(st_a, make_float4(beta.x, beta.y, beta.z, __int_as_float(depth)));
            sts_b32(st_a + kStB + 12u, skip);
          } else {
            alive = false;
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
      if (!alive) {  // ---- the next sample of the lane's work item, or the next item (render_kernel's regeneration) ----
        const float4 B0 = ldv_f4(st_a + kStB);
        int pixel = __float_as_int(B0.x), s = __float_as_int(B0.y);
        PathKey key{P.key, uint32_t(pixel), 0u};
        if ((((unsigned)(s - P.sample_begin)) & (unsigned)(P.chunk - 1)) == 0u || s >= s_last) {
          done = true;
          for (;;) {
            const unsigned long long it = atomicAdd(P.counters, 1ull);
            if (it >= (unsigned long long)P.n_items) break;
            const unsigned int item = (unsigned int)it;
            const unsigned int chunk = item / P.per_chunk, q = item - chunk * P.per_chunk;
            const unsigned int tile = q >> 5, lane = q & 31u;
            const int px = int(tile % (unsigned)P.tiles_x) * 8 + int(lane & 7u);
            const int py = int(tile / (unsigned)P.tiles_x) * 4 + int(lane >> 3);
            if (px < P.cam.W && py < P.cam.H) {
              s = P.sample_begin + int(chunk) * P.chunk;
              if (s < s_last) {
                pixel = py * P.cam.W + px;
                key.pixel = uint32_t(pixel);
                done = false;
                break;
              }
            }
          }
        }
        if (!done) {  // camera::get_ray (camera.hpp:139-162)
          key.sample = uint32_t(s++);
          const int py = pixel / P.cam.W, px = pixel - py * P.cam.W;
          uint4 r0 = rng_block(key, 0u, 0u);
          float ox = u01(r0.x) - 0.5f, oy = u01(r0.y) - 0.5f;
          time = u01(r0.z);
          float3 dir = fma3(float(px) + ox, P.cam.du, fma3(float(py) + oy, P.cam.dv, P.cam.p00c));
          o = P.cam.center;
          if (P.cam.defocus) {
            uint4 r1 = rng_block(key, 0u, 1u);
            float rr = sqrtf(u01(r1.x)), sn, cs;
            sincos_2pi(u01(r1.y), sn, cs);
            float3 off = fma3(rr * cs, P.cam.ddu, (rr * sn) * P.cam.ddv);
            o = o + off;
            dir = dir - off;
          }
          d = dir;
          sts_f4(st_a, make_float4(1.0f, 1.0f, 1.0f, __int_as_float(P.cam.max_depth)));
          sts_f4(st_a + kStB, make_float4(__int_as_float(pixel), __int_as_float(s), time, __uint_as_float(REF_NONE)));
          alive = P.cam.max_depth > 0;
        }
      }
      if (alive) {  // ---- world.hit, part 1: the scene-enclosing media seed the closest hit; the ray goes to its records ----
        Hit best{INF, REF_NONE};
        if (media && sc.n_global_media) {
          PathKey k;
          uint32_t b;
          key_of(k, b);
          best = sample_global_media<COUNT>(sc, o, d, time, 0.001f, INF, k, b, cn);
        }
        sts_f4(st_a + kStC, make_float4(o.x, o.y, o.z, best.t));
        sts_f4(st_a + kStD, make_float4(d.x, d.y, d.z, __uint_as_float(best.ref)));
        ts.cur = 0;
        trav_reset(ts, st);
      }
    }
    const unsigned out_of_work = __ballot_sync(FULL, done);
    if (!__any_sync(FULL, ts.cur != kTravDone)) {
      if (out_of_work == FULL) break;
      continue;  // (lanes that regenerated into max_depth <= 0 never get here: the host does not launch then)
    }
    // =============================== TRACE ===============================
    {  // every lane: its ray comes back from the records
      const float4 Cc = ldv_f4(st_a + kStC), Dd = ldv_f4(st_a + kStD);
      trav_set_ray(ts, f3(Cc.x, Cc.y, Cc.z), f3(Dd.x, Dd.y, Dd.z), 0.0f, 0.001f, REF_NONE);
      ts.best = Hit{Cc.w, __float_as_uint(Dd.w)};
    }
    for (;;) {
      const unsigned bn = __ballot_sync(FULL, ts.cur >= 0);
      if (__popc(bn) >= RT_REFILL_NODE_THR) {
#pragma unroll
        for (int u = 0; u < RT_NODE_UNROLL; u++)
          if (ts.cur >= 0) node_step<COUNT, ALL_SMEM>(ts, st, ns, cn);
        continue;
      }
      // few lanes want a node step: leave once enough lanes wait for a shade (or nobody traces any more) ...
      const unsigned busy = __ballot_sync(FULL, ts.cur != kTravDone);
      if (busy == 0u || __popc(~busy & ~out_of_work) >= RT_REFILL_SHADE_THR) break;
      if (busy & ~bn) {  // ... else the lanes on a leaf go first ...
        if (ts.cur < 0 && ts.cur != kTravDone) {
          aux_of(ts.time, ts.skip);
          leaf_step<COUNT, false, ALL_SMEM>(ts, st, sc, media, key_of, cn, ls);
        }
      } else {  // ... or, with only node work left, the stragglers
        if (ts.cur >= 0) node_step<COUNT, ALL_SMEM>(ts, st, ns, cn);
      }
    }
    // park: the closest hit so far (a finished lane's answer, a suspended lane's bound)
    sts_b32(st_a + kStC + 12u, __float_as_uint(ts.best.t));
    sts_b32(st_a + kStD + 12u, ts.best.ref);
  }
  if ((threadIdx.x & 31) == 0) atomicAdd(P.counters + 1, (unsigned long long)n_rays);
  if (COUNT)
    for (int i = 0; i < CN_COUNT; i++)
      if (cn[i]) atomicAdd(P.counters + 4 + i, (unsigned long long)cn[i]);
}

}  // namespace rtb200
