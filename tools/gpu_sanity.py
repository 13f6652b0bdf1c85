"""GPU bring-up script (run under gpurun): parity vs the CPU oracle + quick timings for every
named scene; writes gpurun_out/sanity.json and small PNG previews.  Not part of the product."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-practice_b200")
from oracle import orc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def main():
    names = sys.argv[1:] or rtb.scene_names()
    ctx = rtb.Context(0)
    report = {}
    for name in names:
        r = {}
        sc = rtb.Scene(name, rand_seed=1)
        t0 = time.time()
        ctx.upload_scene(sc.desc)
        r["upload_s"] = time.time() - t0
        st = ctx.stats()
        r.update(nodes=st.n_nodes, spheres=st.n_spheres, quads=st.n_quads, media=st.n_media, smem_nodes=st.bvh_nodes_in_smem)
        # ---- primary visibility (full resolution) ----
        cam = sc.camera_copy()
        oids, ot, onrm = orc.primary(sc.desc, cam)
        ids, t, nrm = ctx.primary_visibility(cam)
        hit = (oids >= 0) & (ids == oids)
        r["primary_px"] = int(oids.size)
        r["exact_id_mismatch"] = int((ids != oids).sum())
        r["exact_t_maxrel"] = float(np.max(np.abs(t[hit] - ot[hit]) / np.abs(ot[hit]))) if hit.any() else 0.0
        r["exact_n_maxabs"] = float(np.max(np.abs(nrm[hit] - onrm[hit]))) if hit.any() else 0.0
        ids32, t32, n32 = ctx.primary_visibility(cam, flags=rtb.RT_TRACE_FP32 | rtb.RT_TRACE_SKIP_MEDIA)
        hit32 = (oids >= 0) & (ids32 == oids)
        r["fp32_id_mismatch"] = int((ids32 != oids).sum())
        r["fp32_t_maxrel"] = float(np.max(np.abs(t32[hit32] - ot[hit32]) / np.abs(ot[hit32]))) if hit32.any() else 0.0
        r["fp32_n_maxabs"] = float(np.max(np.abs(n32[hit32] - onrm[hit32]))) if hit32.any() else 0.0
        # ---- small converged comparison ----
        w = 160 if cam.aspect_ratio > 1.2 else 120
        small = sc.camera_copy(image_width=w, samples_per_pixel=256)
        t0 = time.time()
        mean, var, orays = orc.render_linear(sc.desc, small, spp=64, seed=5)
        r["oracle_s"] = time.time() - t0
        ctx.render(small, seed=11)
        img = ctx.download_radiance(small.samples_per_pixel).astype(np.float64)
        st = ctx.stats()
        r["small_rays_per_sample_gpu"] = st.rays / max(st.samples, 1)
        r["small_rays_per_sample_oracle"] = orays / (mean.shape[0] * mean.shape[1] * 64)
        g_gpu = orc.write_color(img).astype(np.float64) / 255.0
        g_cpu = orc.write_color(mean).astype(np.float64) / 255.0
        r["small_rmse_gamma"] = float(np.sqrt(np.mean((g_gpu - g_cpu) ** 2)))
        r["small_mean_gpu"] = [float(x) for x in img.reshape(-1, 3).mean(0)]
        r["small_mean_oracle"] = [float(x) for x in mean.reshape(-1, 3).mean(0)]
        r["small_mean_stderr"] = [float(x) for x in np.sqrt(var.reshape(-1, 3).mean(0) / (var.size / 3))]
        try:
            from PIL import Image

            Image.fromarray(np.concatenate([orc.write_color(img), orc.write_color(mean)], axis=1)).save(os.path.join(OUT, f"prev_{name}.png"))
        except Exception as e:  # noqa: BLE001
            r["png_error"] = str(e)
        # ---- timing at the scene's own configuration (capped so bring-up stays short) ----
        full = sc.camera_copy()
        full.samples_per_pixel = min(full.samples_per_pixel, 256)
        ctx.render(full, seed=1)
        ctx.synchronize()
        ctx.render(full, seed=2)
        st = ctx.stats()
        r["full"] = dict(w=st.image_width, h=st.image_height, spp=full.samples_per_pixel, ms=st.last_render_ms, rays=int(st.rays),
                         mrays_s=st.rays / st.last_render_ms / 1e3, msamples_s=st.samples / st.last_render_ms / 1e3)
        print(name, json.dumps(r), flush=True)
        report[name] = r
        sc.close()
    with open(os.path.join(OUT, "sanity.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
