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

// one path = 16 words, SoA per warp: pool[field * NP + slot]
enum : int { PF_OX = 0, PF_OY, PF_OZ, PF_DX, PF_DY, PF_DZ, PF_TIME, PF_SKIP, PF_HT, PF_HREF, PF_BX, PF_BY, PF_BZ, PF_PIXEL, PF_SAMPLE, PF_DEPTH, PF_WORDS };

constexpr size_t pool_smem_bytes(int threads) { return size_t(threads / 32) * (PF_WORDS * RT_POOL_NP * 4 + 2 * RT_POOL_NP); }

template <bool COUNT>
__global__ void __launch_bounds__(kRenderThreads, 1) pool_kernel(const __grid_constant__ RenderParams P) {
  constexpr int NP = RT_POOL_NP;
  constexpr unsigned QM = NP - 1;
  static_assert((NP & (NP - 1)) == 0 && NP >= 32 && NP <= 256, "RT_POOL_NP must be a power of two in [32, 256]");
  extern __shared__ float4 s_dyn[];
  float4* s_nodes = s_dyn;
  for (int i = threadIdx.x; i < 4 * P.smem_nodes; i += blockDim.x) s_nodes[i] = P.sc.nodes[i];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  float* const pool = reinterpret_cast<float*>(s_dyn + 4 * P.smem_nodes) + warp * (PF_WORDS * NP);
  unsigned char* const tq = reinterpret_cast<unsigned char*>(reinterpret_cast<float*>(s_dyn + 4 * P.smem_nodes) + n_warps * (PF_WORDS * NP)) + warp * (2 * NP);
  unsigned char* const sq = tq + NP;
#define PF(f, s) pool[(f) * NP + (s)]
#define PI(f, s) reinterpret_cast<int*>(pool)[(f) * NP + (s)]
  // every slot starts "fresh" (depth 0, no pixel) in the shade queue: its first shade is a regeneration
  for (unsigned s = lane; s < NP; s += 32) {
    sq[s] = (unsigned char)s;
    PI(PF_DEPTH, s) = 0;
    PI(PF_PIXEL, s) = -1;
    PI(PF_SAMPLE, s) = 0;
  }
  __syncthreads();

  const NodeSource ns{s_nodes, P.sc.nodes, P.smem_nodes};
  const DeviceScene& sc = P.sc;
  const float INF = __int_as_float(0x7f800000);
  const unsigned FULL = 0xFFFFFFFFu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const bool media = sc.n_media != 0;
  const int s_last = P.sample_begin + P.sample_count;
  unsigned int n_rays = 0;
  unsigned int cn[COUNT ? CN_COUNT : 1];
  if (COUNT)
    for (int i = 0; i < CN_COUNT; i++) cn[i] = 0;

  // warp-uniform queue state (ring buffers over NP entries)
  unsigned tq_head = 0, tq_n = 0, sq_head = 0, sq_n = NP, n_dead = 0;
  // this lane's traversal engine
  TravState ts;
  TravStack st;
  ts.tmin = 0.001f;  // camera.hpp:192
  ts.best = Hit{INF, REF_NONE};
  ts.cur = 0, ts.sp = 0;
  int mode = MODE_DONE;  // idle
  unsigned tslot = 0;
  unsigned n_busy = 0;  // lanes holding a (suspended) traversal: warp-uniform

  for (;;) {
    // =========================== SHADE + REGENERATE =================================================
    while (sq_n >= 32u || (sq_n > 0u && tq_n < 32u - n_busy)) {
      const unsigned take = min(sq_n, 32u);
      bool has_ray = false, dead = false;
      unsigned slot = 0;
      if (lane < take) {
        slot = sq[(sq_head + lane) & QM];
        int depth = PI(PF_DEPTH, slot), pixel = PI(PF_PIXEL, slot), smp = PI(PF_SAMPLE, slot);
        float3 o, d, beta;
        float time;
        uint32_t skip;
        bool regen = true;
        if (depth > 0) {  // ---- the tail of one ray_color level (camera.hpp:192-231) ----
          o = f3(PF(PF_OX, slot), PF(PF_OY, slot), PF(PF_OZ, slot));
          d = f3(PF(PF_DX, slot), PF(PF_DY, slot), PF(PF_DZ, slot));
          time = PF(PF_TIME, slot);
          beta = f3(PF(PF_BX, slot), PF(PF_BY, slot), PF(PF_BZ, slot));
          const Hit h{PF(PF_HT, slot), (uint32_t)PI(PF_HREF, slot)};
          // radiance of a path = beta * (emission | background) at its LAST vertex: no material here both
          // emits and scatters (diffuse_light::scatter is false, material.hpp:36), so no running sum is kept
          float3 L = f3(0.0f, 0.0f, 0.0f);
          bool alive = false;
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
          if (!alive) {
            unsigned long long* dst = P.accum + 3ull * (unsigned long long)pixel;
            const long long fr = to_fixed(L.x), fg = to_fixed(L.y), fb = to_fixed(L.z);
            if (fr) atomicAdd(dst + 0, (unsigned long long)fr);
            if (fg) atomicAdd(dst + 1, (unsigned long long)fg);
            if (fb) atomicAdd(dst + 2, (unsigned long long)fb);
          } else {
            regen = false;
          }
        }
        if (regen) {  // ---- next sample of this slot's work item, or the next item ----
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
          if (have) {  // camera::get_ray (camera.hpp:139-162): jitter, defocus disk, shutter time
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
              sincospif(2.0f * u01(r1.y), &sn, &cs);
              const float3 off = fma3(rr * cs, P.cam.ddu, (rr * sn) * P.cam.ddv);
              o = o + off;
              dir = dir - off;
            }
            d = dir;
            beta = f3(1.0f, 1.0f, 1.0f);
            depth = P.cam.max_depth;
            skip = REF_NONE;
            PI(PF_PIXEL, slot) = pixel;
            PI(PF_SAMPLE, slot) = smp;
            PF(PF_TIME, slot) = time;
          } else {
            dead = true;
            PI(PF_DEPTH, slot) = 0;
          }
        }
        if (!dead) {
          // the scene-enclosing media (met by every ray: the r=5000 fog of the Book-2 final scene) are sampled
          // here, for the NEXT query, while the warp is converged; the result seeds the traversal's closest hit
          Hit best{INF, REF_NONE};
          if (media && sc.n_global_media) {
            const PathKey key{P.key, (uint32_t)pixel, (uint32_t)smp};
            const uint32_t bounce = uint32_t(P.cam.max_depth - depth) + 1u;
            for (int g = 0; g < sc.n_global_media; g++) {
              const int mi = sc.global_media[g];
              const DMedium m = sc.media[mi];
              const float t = medium_sample(sc, m, o, d, time, 0.001f, best.t, medium_uniform(key, bounce, mi));
              if (COUNT) cn[CN_MEDIUM]++;
              if (t != -1.0f) best = Hit{t, make_ref(REF_MEDIUM, uint32_t(mi))};
            }
          }
          PF(PF_OX, slot) = o.x, PF(PF_OY, slot) = o.y, PF(PF_OZ, slot) = o.z;
          PF(PF_DX, slot) = d.x, PF(PF_DY, slot) = d.y, PF(PF_DZ, slot) = d.z;
          PI(PF_SKIP, slot) = (int)skip;
          PF(PF_HT, slot) = best.t, PI(PF_HREF, slot) = (int)best.ref;
          PF(PF_BX, slot) = beta.x, PF(PF_BY, slot) = beta.y, PF(PF_BZ, slot) = beta.z;
          PI(PF_DEPTH, slot) = depth;
          has_ray = true;
        }
      }
      sq_head = (sq_head + take) & QM;
      sq_n -= take;
      const unsigned bR = __ballot_sync(FULL, has_ray);
      if (has_ray) tq[(tq_head + tq_n + __popc(bR & lt_mask)) & QM] = (unsigned char)slot;
      tq_n += __popc(bR);
      n_dead += __popc(__ballot_sync(FULL, dead));
      __syncwarp();
    }
    if (n_dead == NP) break;

    // =========================== TRACE ==============================================================
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
            ts.inv = f3(fabsf(d.x) > 1e-30f ? __frcp_rn(d.x) : copysignf(1e30f, d.x), fabsf(d.y) > 1e-30f ? __frcp_rn(d.y) : copysignf(1e30f, d.y),
                        fabsf(d.z) > 1e-30f ? __frcp_rn(d.z) : copysignf(1e30f, d.z));
            ts.ood = o * ts.inv;
            ts.best = Hit{PF(PF_HT, tslot), (uint32_t)PI(PF_HREF, tslot)};
            ts.sp = 0;
            ts.cur = 0;
            mode = MODE_NODE;
            n_rays++;
          }
          const unsigned took = min(32u - n_busy, tq_n);
          tq_head = (tq_head + took) & QM;
          tq_n -= took;
          continue;
        }
        if (sq_n > 0u || busy == 0u) break;  // finished queries wait for shading (or nothing is left at all)
      }
      if (__popc(bN) >= __popc(bL)) {
        if (mode == MODE_NODE) mode = node_step<COUNT>(ts, st, ns, cn);
      } else {
        if (mode == MODE_LEAF) {
          ts.o = f3(PF(PF_OX, tslot), PF(PF_OY, tslot), PF(PF_OZ, tslot));
          ts.d = f3(PF(PF_DX, tslot), PF(PF_DY, tslot), PF(PF_DZ, tslot));
          ts.time = PF(PF_TIME, tslot);
          ts.skip = (uint32_t)PI(PF_SKIP, tslot);
          const PathKey key{P.key, (uint32_t)PI(PF_PIXEL, tslot), (uint32_t)PI(PF_SAMPLE, tslot)};
          const uint32_t bounce = uint32_t(P.cam.max_depth - PI(PF_DEPTH, tslot)) + 1u;
          mode = leaf_step<COUNT>(ts, st, sc, media, key, bounce, cn);
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
    __syncwarp();
  }
#undef PF
#undef PI
  // ---- counters: warp-reduce, one atomic per warp ---------------------------------------
  unsigned int rays = n_rays;
  for (int off = 16; off > 0; off >>= 1) rays += __shfl_down_sync(FULL, rays, off);
  if (lane == 0 && rays) atomicAdd(P.counters + 1, (unsigned long long)rays);
  if (COUNT)
    for (int i = 0; i < CN_COUNT; i++)
      if (cn[i]) atomicAdd(P.counters + 4 + i, (unsigned long long)cn[i]);
}

}  // namespace rtb200
