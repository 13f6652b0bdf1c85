// rt_pool.cuh — the render path as a persistent kernel with a PER-WARP PATH POOL in shared memory.
//
// Why (profiles/r05_render_mega_phases.md): in the megakernel a lane is welded to one path, so a warp's
// while-while traversal runs until its LONGEST ray is done.  Ray lengths on the Book-2 final scene are
// extremely skewed (a fog-scattered ray ends after 1-2 BVH nodes, a ray into the sphere cluster takes 40+),
// and ncu measured 5.5 of 32 lanes active in node_step and 6.8 in leaf_step — 59 % of all issued
// warp instructions.  Here camera::render's loop body (camera.hpp:55-62, 180-232) is split in two phases
// that a warp alternates between, over a pool of RT_POOL_NP (> 32) paths it owns:
//
//   TRACE  = world.hit.  A lane is a traversal engine, not a path: when >= RT_POOL_REFILL lanes are idle
//            they pop the next pending rays from the warp's trace queue (a few LDS), so node_step /
//            leaf_step always run on a nearly full warp.  When the queue is empty and lanes go idle, the
//            warp leaves for SHADE with the unfinished traversals SUSPENDED in their lanes' registers.
//   SHADE  = emitted + scatter + texture (+ the scene-enclosing media of the NEXT ray) for 32 finished
//            queries at a time, and REGENERATION: a finished path's slot starts the next camera sample at
//            once, so the pool stays full until the image runs out of samples.
//
// Everything is warp-synchronous: queues are ring buffers of slot numbers in shared memory whose
// head / count are warp-uniform registers updated from ballots — no atomics, no block or grid barrier.
// (A grid-wide wavefront with the pool in L2 and cooperative-groups barriers was built first and ran
// 2.2x SLOWER than the megakernel: profiles/r05_wavefront_rejected.md.)
// RNG keys, sample order within a pixel and the fixed-point accumulation are exactly the megakernel's,
// so the accumulator is BIT-IDENTICAL to the megakernel's for the same (seed, sample range): tested.
#pragma once

namespace rtb200 {

#ifndef RT_POOL_NP
#define RT_POOL_NP 64  // paths per warp (power of two, 64..256)
#endif
#ifndef RT_POOL_REFILL
#define RT_POOL_REFILL 12  // refill idle trace lanes once at least this many are idle
#endif

#ifndef RT_POOL_SHADE_MISSES
#define RT_POOL_SHADE_MISSES 0  // 1: SHADE finishes rays that miss the scene's bounding box itself (measured: -28 %, a divergent loop)
#endif
#ifndef RT_POOL_NODE_LOOP
#define RT_POOL_NODE_LOOP 1  // TRACE keeps taking node steps while NODE lanes stay the majority (1 vote per step)
#endif
#ifndef RT_POOL_COLD_GLOBAL
#define RT_POOL_COLD_GLOBAL 0  // 1: the shade-only words of a path live in global memory (L1/L2), not shared
#endif

// one path = 16 words, SoA per warp: hot[field * NP + slot] is what TRACE touches (always shared memory),
// cold[field * NP + slot] is only touched by SHADE (and by a medium's lazy RNG key)
enum : int { PF_OX = 0, PF_OY, PF_OZ, PF_DX, PF_DY, PF_DZ, PF_TIME, PF_SKIP, PF_HT, PF_HREF, PF_WORDS };
enum : int { PC_BX = 0, PC_BY, PC_BZ, PC_PIXEL, PC_SAMPLE, PC_DEPTH, PC_WORDS };

constexpr size_t pool_smem_bytes(int threads) {
  return size_t(threads / 32) * ((PF_WORDS + (RT_POOL_COLD_GLOBAL ? 0 : PC_WORDS)) * RT_POOL_NP * 4 + 2 * RT_POOL_NP + 8 * 4 + 32 * 32);
}
constexpr size_t pool_cold_global_bytes(int ctas, int threads) { return RT_POOL_COLD_GLOBAL ? size_t(ctas) * (threads / 32) * PC_WORDS * RT_POOL_NP * 4 : 0; }

#define PF(f, s) pool[(f) * NP + (s)]
#define PI(f, s) reinterpret_cast<int*>(pool)[(f) * NP + (s)]
#define CF(f, s) cold[(f) * NP + (s)]
#define CI(f, s) reinterpret_cast<int*>(cold)[(f) * NP + (s)]

// SHADE + REGENERATE for one pool slot: the tail of one ray_color level (camera.hpp:192-231) for the finished
// query in `slot`, then — if the path ended — the next camera sample (camera.hpp:139-162).  Returns 1 when the
// slot holds a new ray to trace, 0 when the image has no samples left for it (the slot is dead).
// Deliberately NOT inlined: its ~60 live registers must not compete with the traversal loop's (which then
// runs without local-memory traffic at 64 registers / 1024 threads); `Pp` points at the CTA's shared-memory
// copy of the launch parameters, because an outlined function cannot address the kernel's constant bank.
template <bool COUNT>
__device__ __noinline__ int shade_slot(const RenderParams* __restrict__ Pp, float* __restrict__ pool, float* __restrict__ cold, unsigned slot,
                                       unsigned int& rays_here, unsigned int* cn) {
  constexpr int NP = RT_POOL_NP;
  const RenderParams& P = *Pp;
  const DeviceScene& sc = P.sc;
  const float INF = __int_as_float(0x7f800000);
  int depth = CI(PC_DEPTH, slot), pixel = CI(PC_PIXEL, slot), smp = CI(PC_SAMPLE, slot);
  float3 o, d, beta;
  float time;
  uint32_t skip;
  // radiance of a path = beta * (emission | background) at its LAST vertex: no material here both emits and
  // scatters (diffuse_light::scatter is false, material.hpp:36), so no running sum is kept in the pool
  auto deposit = [&](float3 L) {  // the finished sample, quantised to 2^-32, straight into the int64 accumulator
    RT_CHECK(pixel >= 0 && (unsigned long long)pixel * 3ull + 2ull < P.n_values, CHK_PIXEL);
    unsigned long long* dst = P.accum + 3ull * (unsigned long long)pixel;
    const long long fr = to_fixed(L.x), fg = to_fixed(L.y), fb = to_fixed(L.z);
    if (fr) atomicAdd(dst + 0, (unsigned long long)fr);
    if (fg) atomicAdd(dst + 1, (unsigned long long)fg);
    if (fb) atomicAdd(dst + 2, (unsigned long long)fb);
  };
  bool alive = false;
  if (depth > 0) {  // ---- the tail of one ray_color level (camera.hpp:192-231) for the query TRACE finished ----
    o = f3(PF(PF_OX, slot), PF(PF_OY, slot), PF(PF_OZ, slot));
    d = f3(PF(PF_DX, slot), PF(PF_DY, slot), PF(PF_DZ, slot));
    time = PF(PF_TIME, slot);
    beta = f3(CF(PC_BX, slot), CF(PC_BY, slot), CF(PC_BZ, slot));
    const Hit h{PF(PF_HT, slot), (uint32_t)PI(PF_HREF, slot)};
    float3 L = f3(0.0f, 0.0f, 0.0f);
    if (h.ref == REF_NONE) {
      L = L + beta * P.cam.bg;
    } else {
      const PathKey key{P.key, (uint32_t)pixel, (uint32_t)smp};
      const uint32_t bounce = uint32_t(P.cam.max_depth - depth) + 1u;
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
    if (!alive) deposit(L);
  }
  // ---- until this slot holds a ray that needs the BVH (or the image has no samples left for it) ----
  Hit best;
#pragma unroll 1
  for (;;) {
    if (!alive) {  // next sample of this slot's work item, or the next item
      const int s_last = P.sample_begin + P.sample_count;
      int s_next = smp + 1;
      bool have = pixel >= 0 && ((s_next - P.sample_begin) & (P.chunk - 1)) != 0 && s_next < s_last;
      if (!have) {
        for (;;) {
          const unsigned long long it = atomicAdd(P.counters, 1ull);
          if (it >= (unsigned long long)P.n_items) break;
          const unsigned int item = (unsigned int)it;
          const unsigned int chunk = item / P.per_chunk, q = item - chunk * P.per_chunk;
          const unsigned int tile = q >> 5, l = q & 31u;
          const int px = int(tile % (unsigned)P.tiles_x) * 8 + int(l & 7u);
          const int py = int(tile / (unsigned)P.tiles_x) * 4 + int(l >> 3);
          s_next = P.sample_begin + int(chunk) * P.chunk;
          if (px < P.cam.W && py < P.cam.H && s_next < s_last) {
            pixel = py * P.cam.W + px;
            have = true;
            break;
          }
        }
      }
      if (!have) {
        CI(PC_DEPTH, slot) = 0;
        return 0;
      }
      // camera::get_ray (camera.hpp:139-162): jitter, defocus disk, shutter time
      smp = s_next;
      const PathKey key{P.key, (uint32_t)pixel, (uint32_t)smp};
      const int py = pixel / P.cam.W, px = pixel - py * P.cam.W;
      const uint4 r0 = rng_block(key, 0u, 0u);
      const float ox = u01(r0.x) - 0.5f, oy = u01(r0.y) - 0.5f;
      time = u01(r0.z);
      float3 dir = fma3(float(px) + ox, P.cam.du, fma3(float(py) + oy, P.cam.dv, P.cam.p00c));
      o = P.cam.center;
      if (P.cam.defocus) {  // uniform disk: r = sqrt(u), phi = 2 pi v (== rejection sampling in law)
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
      alive = true;
    }
    // world.hit, part 1: the scene-enclosing media (met by every ray: the r=5000 fog of the Book-2 final scene)
    // are sampled here, for the NEXT query, while the warp is converged; the result seeds the closest hit
    const PathKey key{P.key, (uint32_t)pixel, (uint32_t)smp};
    const uint32_t bounce = uint32_t(P.cam.max_depth - depth) + 1u;
    best = Hit{INF, REF_NONE};
    if (sc.n_global_media) best = sample_global_media<COUNT>(sc, o, d, time, 0.001f, INF, key, bounce, cn);
#if RT_POOL_SHADE_MISSES
    // ... and a ray that misses the box of all geometry (or is stopped by the medium first) needs no BVH at
    // all: finish it here, where the cost of doing so is shared by the 32 lanes of the SHADE round
    TravState ts;
    trav_set_ray(ts, o, d, time, 0.001f, skip);
    if (ray_meets_scene(sc, ts, best.t)) break;
    if (best.ref == REF_NONE) {
      rays_here++;
      deposit(beta * P.cam.bg);
      alive = false;
    } else {
      const int mat = sc.media[best.ref & 0x3FFFFFFFu].material;
      const float4 m1 = __ldg(sc.materials + 2 * mat + 1);
      if (__float_as_int(m1.x) != MAT_ISOTROPIC) break;  // not a phase function this shortcut knows: general path
      rays_here++;
      if (COUNT) cn[CN_ISO]++;
      const uint4 rnd = rng_block(key, bounce, 0u);
      const float3 p = fma3(best.t, d, o);
      const float3 tv = texture_value<COUNT>(sc, __float_as_int(m1.y), 0.0f, 0.0f, p, cn);
      beta = beta * tv;  // isotropic::scatter: attenuation = texture, direction uniform on the sphere
      o = p;
      d = unit_vector_from(u01(rnd.x), u01(rnd.y));
      skip = REF_NONE;
      alive = --depth > 0;  // a path cut by max_depth contributes black: nothing to deposit
    }
#else
    break;
#endif
  }
  CI(PC_PIXEL, slot) = pixel;
  CI(PC_SAMPLE, slot) = smp;
  PF(PF_TIME, slot) = time;
  PF(PF_OX, slot) = o.x, PF(PF_OY, slot) = o.y, PF(PF_OZ, slot) = o.z;
  PF(PF_DX, slot) = d.x, PF(PF_DY, slot) = d.y, PF(PF_DZ, slot) = d.z;
  PI(PF_SKIP, slot) = (int)skip;
  PF(PF_HT, slot) = best.t, PI(PF_HREF, slot) = (int)best.ref;
  CF(PC_BX, slot) = beta.x, CF(PC_BY, slot) = beta.y, CF(PC_BZ, slot) = beta.z;
  CI(PC_DEPTH, slot) = depth;
  return 1;
}

// per-warp control block (shared memory, warp-uniform values; ring buffers over NP entries)
enum : int { QC_TQ_HEAD = 0, QC_TQ_N, QC_SQ_HEAD, QC_SQ_N, QC_DEAD, QC_RAYS, QC_BUSY, QC_WORDS = 8 };
// per-warp block: [SW x NP words of path state][NP bytes trace queue][NP bytes shade queue][QC_WORDS control]
//                 [2 x 32 float4 lane scratch: {1/d.xyz, (o/d).x} {(o/d).yz, cur, sp | tslot << 8 | mode << 16}]
constexpr int kPoolSW = PF_WORDS + (RT_POOL_COLD_GLOBAL ? 0 : PC_WORDS);
constexpr int kPoolQueueOff = kPoolSW * RT_POOL_NP;                 // in words
constexpr int kPoolCtlOff = kPoolQueueOff + 2 * RT_POOL_NP / 4;
constexpr int kPoolLaneOff = kPoolCtlOff + QC_WORDS;
constexpr int kPoolWarpWords = kPoolLaneOff + 2 * 32 * 4;
static_assert(kPoolLaneOff % 4 == 0, "lane scratch must be 16-byte aligned");

// SHADE phase of one warp: shades finished queries 32 at a time while that fills the trace queue; returns
// false when every slot of the pool is dead (the image has no samples left).  Outlined, see shade_slot.
template <bool COUNT>
__device__ __noinline__ bool shade_phase(const RenderParams* __restrict__ Pp, float* __restrict__ pool, float* __restrict__ cold, unsigned int* cn) {
  constexpr int NP = RT_POOL_NP;
  constexpr unsigned QM = NP - 1;
  const unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
  unsigned char* const tq = reinterpret_cast<unsigned char*>(pool + kPoolQueueOff);
  unsigned char* const sq = tq + NP;
  unsigned int* const qc = reinterpret_cast<unsigned int*>(pool + kPoolCtlOff);
  unsigned tq_head = qc[QC_TQ_HEAD], tq_n = qc[QC_TQ_N], sq_head = qc[QC_SQ_HEAD], sq_n = qc[QC_SQ_N], n_dead = qc[QC_DEAD];
  const unsigned n_busy = qc[QC_BUSY];
  unsigned int rays_here = 0;
  __syncwarp();
  while (sq_n >= 32u || (sq_n > 0u && tq_n < 32u - n_busy)) {
    const unsigned take = min(sq_n, 32u);
    int has_ray = -1;  // -1: this lane had no slot
    unsigned slot = 0;
    if (lane < take) {
      slot = sq[(sq_head + lane) & QM];
      has_ray = shade_slot<COUNT>(Pp, pool, cold, slot, rays_here, cn);
    }
    sq_head = (sq_head + take) & QM;
    sq_n -= take;
    const unsigned bR = __ballot_sync(FULL, has_ray == 1);
    if (has_ray == 1) tq[(tq_head + tq_n + __popc(bR & lt_mask)) & QM] = (unsigned char)slot;
    tq_n += __popc(bR);
    n_dead += __popc(__ballot_sync(FULL, has_ray == 0));
  }
  rays_here = __reduce_add_sync(FULL, rays_here);
  if (lane == 0) qc[QC_TQ_HEAD] = tq_head, qc[QC_TQ_N] = tq_n, qc[QC_SQ_HEAD] = sq_head, qc[QC_SQ_N] = sq_n, qc[QC_DEAD] = n_dead, qc[QC_RAYS] += rays_here;
  __syncwarp();
  return n_dead != NP;
}

// TRACE phase of one warp.  Outlined and CALL-FREE on purpose: with any call in the same function ptxas homes
// every value that lives across it (node index, stack pointer, 1/d ...) in local memory and each node step
// pays for it (profiles/r06_pool_kernel.md); here the whole phase gets its own register allocation.
template <bool COUNT>
__device__ __noinline__ void trace_phase(const RenderParams* __restrict__ Pp, float* __restrict__ pool, float* __restrict__ cold, const float4* s_nodes,
                                         TravStack& st, unsigned int* cn) {
  constexpr int NP = RT_POOL_NP;
  constexpr unsigned QM = NP - 1;
  const RenderParams& P = *Pp;
  const DeviceScene& sc = P.sc;
  const unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
  const NodeSource ns = node_source(s_nodes, sc.nodes, P.smem_nodes);
  const bool media = sc.n_media != 0;
  unsigned char* const tq = reinterpret_cast<unsigned char*>(pool + kPoolQueueOff);
  unsigned char* const sq = tq + NP;
  unsigned int* const qc = reinterpret_cast<unsigned int*>(pool + kPoolCtlOff);
  float4* const lsc = reinterpret_cast<float4*>(pool + kPoolLaneOff) + lane;
  unsigned tq_head = qc[QC_TQ_HEAD], tq_n = qc[QC_TQ_N], sq_n = qc[QC_SQ_N], n_rays = 0, n_busy = 0;
  const unsigned sq_head = qc[QC_SQ_HEAD];
  // this lane's traversal engine: a suspended traversal comes back from the lane scratch
  TravState ts;
  ts.tmin = 0.001f;  // camera.hpp:192
  int mode;
  unsigned tslot;
  {
    const float4 s1 = lsc[32];
    const unsigned w = (unsigned)__float_as_int(s1.w);
    ts.cur = __float_as_int(s1.z);
    ts.sp = int(w & 0xFFu), tslot = (w >> 8) & 0xFFu, mode = int(w >> 16);
    ts.best = Hit{PF(PF_HT, tslot), (uint32_t)PI(PF_HREF, tslot)};
  }
  __syncwarp();
  for (;;) {
    const unsigned bN = __ballot_sync(FULL, mode == MODE_NODE), bL = __ballot_sync(FULL, mode == MODE_LEAF);
    const unsigned busy = bN | bL;
    n_busy = __popc(busy);
    if (32u - n_busy >= RT_POOL_REFILL || busy == 0u) {
      if (tq_n > 0u) {  // idle lanes pop the next pending rays
        const unsigned rank = __popc(~busy & lt_mask);
        if (mode == MODE_DONE && rank < tq_n) {
          tslot = tq[(tq_head + rank) & QM];
          const float3 o = f3(PF(PF_OX, tslot), PF(PF_OY, tslot), PF(PF_OZ, tslot));
          const float3 d = f3(PF(PF_DX, tslot), PF(PF_DY, tslot), PF(PF_DZ, tslot));
          const float3 inv = f3(fabsf(d.x) > 1e-30f ? rcp_fast(d.x) : copysignf(1e30f, d.x), fabsf(d.y) > 1e-30f ? rcp_fast(d.y) : copysignf(1e30f, d.y),
                                fabsf(d.z) > 1e-30f ? rcp_fast(d.z) : copysignf(1e30f, d.z));
          const float3 ood = o * inv;
          lsc[0] = make_float4(inv.x, inv.y, inv.z, ood.x);
          lsc[32] = make_float4(ood.y, ood.z, 0.0f, 0.0f);
          ts.best = Hit{PF(PF_HT, tslot), (uint32_t)PI(PF_HREF, tslot)};
          ts.sp = 0;
          ts.cur = 0;
          mode = MODE_NODE;
        }
        const unsigned took = min(32u - n_busy, tq_n);
        tq_head = (tq_head + took) & QM;
        tq_n -= took;
        n_rays += took;
        continue;
      }
      if (sq_n > 0u || busy == 0u) break;  // finished queries wait for shading (or nothing is left at all)
    }
    if (__popc(bN) >= __popc(bL)) {
#if RT_POOL_NODE_LOOP
      const unsigned thr = max(1u, (n_busy + 1u) >> 1);
      do {
        if (mode == MODE_NODE) {
          const float4 s0 = lsc[0], s1 = lsc[32];
          ts.inv = f3(s0.x, s0.y, s0.z);
          ts.ood = f3(s0.w, s1.x, s1.y);
          mode = node_step<COUNT>(ts, st, ns, cn);
        }
      } while ((unsigned)__popc(__ballot_sync(FULL, mode == MODE_NODE)) >= thr);
#else
      if (mode == MODE_NODE) {
        const float4 s0 = lsc[0], s1 = lsc[32];
        ts.inv = f3(s0.x, s0.y, s0.z);
        ts.ood = f3(s0.w, s1.x, s1.y);
        mode = node_step<COUNT>(ts, st, ns, cn);
      }
#endif
    } else {
      if (mode == MODE_LEAF) {
        ts.o = f3(PF(PF_OX, tslot), PF(PF_OY, tslot), PF(PF_OZ, tslot));
        ts.d = f3(PF(PF_DX, tslot), PF(PF_DY, tslot), PF(PF_DZ, tslot));
        ts.time = PF(PF_TIME, tslot);
        ts.skip = (uint32_t)PI(PF_SKIP, tslot);
        {
          const float4 s0 = lsc[0], s1 = lsc[32];
          ts.inv = f3(s0.x, s0.y, s0.z);
          ts.ood = f3(s0.w, s1.x, s1.y);
        }
        mode = leaf_step<COUNT, true>(ts, st, sc, media,
                                [&](PathKey& k, uint32_t& b) {
                                  k = PathKey{P.key, (uint32_t)CI(PC_PIXEL, tslot), (uint32_t)CI(PC_SAMPLE, tslot)};
                                  b = uint32_t(P.cam.max_depth - CI(PC_DEPTH, tslot)) + 1u;
                                },
                                cn);
      }
    }
    const unsigned bF = __ballot_sync(FULL, mode == MODE_SHADE);
    if (bF) {  // finished queries: publish the hit, queue the slot for shading, the lane goes idle
      if (mode == MODE_SHADE) {
        PF(PF_HT, tslot) = ts.best.t, PI(PF_HREF, tslot) = (int)ts.best.ref;
        sq[(sq_head + sq_n + __popc(bF & lt_mask)) & QM] = (unsigned char)tslot;
        mode = MODE_DONE;
      }
      sq_n += __popc(bF);
    }
  }
  // suspend the unfinished traversals: the closest hit so far goes to the slot, the rest to the lane scratch
  if (mode != MODE_DONE) PF(PF_HT, tslot) = ts.best.t, PI(PF_HREF, tslot) = (int)ts.best.ref;
  reinterpret_cast<int2*>(lsc + 32)[1] = make_int2(ts.cur, int(unsigned(ts.sp) | (tslot << 8) | (unsigned(mode) << 16)));
  if (lane == 0) {
    qc[QC_TQ_HEAD] = tq_head, qc[QC_TQ_N] = tq_n, qc[QC_SQ_N] = sq_n, qc[QC_BUSY] = n_busy;
    qc[QC_RAYS] += n_rays;
  }
  __syncwarp();
}

template <bool COUNT>
__global__ void __launch_bounds__(kRenderThreads, 1) pool_kernel(const __grid_constant__ RenderParams P) {
  constexpr int NP = RT_POOL_NP;
  static_assert((NP & (NP - 1)) == 0 && NP >= 32 && NP <= 256, "RT_POOL_NP must be a power of two in [32, 256]");
  extern __shared__ float4 s_dyn[];
  __shared__ RenderParams sP;  // the outlined phases cannot address the kernel's constant bank
  for (int i = threadIdx.x; i < int(sizeof(RenderParams) / 4); i += blockDim.x) reinterpret_cast<int*>(&sP)[i] = reinterpret_cast<const int*>(&P)[i];
  float4* s_nodes = s_dyn;
  for (int i = threadIdx.x; i < 4 * P.smem_nodes; i += blockDim.x) s_nodes[i] = P.sc.nodes[i];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  float* const pool = reinterpret_cast<float*>(s_dyn + 4 * P.smem_nodes) + warp * kPoolWarpWords;
#if RT_POOL_COLD_GLOBAL
  float* const cold = P.pool_cold + (size_t(blockIdx.x) * (blockDim.x >> 5) + warp) * (PC_WORDS * NP);
#else
  float* const cold = pool + PF_WORDS * NP;
#endif
  // every slot starts "fresh" (depth 0, no pixel) in the shade queue: its first shade is a regeneration
  {
    unsigned char* const sq = reinterpret_cast<unsigned char*>(pool + kPoolQueueOff) + NP;
    unsigned int* const qc = reinterpret_cast<unsigned int*>(pool + kPoolCtlOff);
    for (unsigned s = lane; s < NP; s += 32) {
      sq[s] = (unsigned char)s;
      CI(PC_DEPTH, s) = 0;
      CI(PC_PIXEL, s) = -1;
      CI(PC_SAMPLE, s) = 0;
    }
    if (lane < QC_WORDS) qc[lane] = lane == QC_SQ_N ? NP : 0u;
    float4* const lsc = reinterpret_cast<float4*>(pool + kPoolLaneOff) + lane;
    lsc[32] = make_float4(0.0f, 0.0f, __int_as_float(0), __int_as_float(MODE_DONE << 16));
  }
  __syncthreads();
  unsigned int cn[COUNT ? CN_COUNT : 1];
  if (COUNT)
    for (int i = 0; i < CN_COUNT; i++) cn[i] = 0;

  TravStack st;  // the traversal stacks outlive a TRACE phase (suspended traversals): they belong to the kernel's frame
  while (shade_phase<COUNT>(&sP, pool, cold, cn)) trace_phase<COUNT>(&sP, pool, cold, s_nodes, st, cn);

  // ---- counters: one atomic per warp ------------------------------------------------------
  if (lane == 0) {
    const unsigned int rays = reinterpret_cast<unsigned int*>(pool + kPoolCtlOff)[QC_RAYS];
    if (rays) atomicAdd(P.counters + 1, (unsigned long long)rays);
  }
  if (COUNT)
    for (int i = 0; i < CN_COUNT; i++)
      if (cn[i]) atomicAdd(P.counters + 4 + i, (unsigned long long)cn[i]);
}
#undef PF
#undef PI
#undef CF
#undef CI

}  // namespace rtb200
