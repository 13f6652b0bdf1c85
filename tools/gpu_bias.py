"""Bias hunt (run under gpurun): GPU at high spp vs the CPU oracle at moderate spp, z-scores of the
mean radiance per channel and per group of pixels bucketed by the primitive the pixel centre sees."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-practice_b200")
from oracle import orc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def main():
    name = sys.argv[1]
    width = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    cpu_spp = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    gpu_spp = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
    depth = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    sc = rtb.Scene(name, rand_seed=1)
    over = dict(image_width=width, samples_per_pixel=gpu_spp)
    if depth:
        over["max_depth"] = depth
    cam = sc.camera_copy(**over)
    ctx = rtb.Context(0)
    ctx.upload_scene(sc.desc)
    ctx.render(cam, seed=3)
    img = ctx.download_radiance(gpu_spp).astype(np.float64)
    st = ctx.stats()
    mean, var, orays = orc.render_linear(sc.desc, cam, spp=cpu_spp, seed=9)
    ids, _, _ = orc.primary(sc.desc, cam, skip_media=True)
    rps_gpu = st.rays / st.samples
    rps_cpu = orays / (mean.shape[0] * mean.shape[1] * cpu_spp)
    var_tot = var * (1.0 + cpu_spp / gpu_spp)
    rep = dict(scene=name, width=width, cpu_spp=cpu_spp, gpu_spp=gpu_spp, rays_per_sample_gpu=rps_gpu, rays_per_sample_cpu=rps_cpu)
    d = img - mean
    rep["global_z"] = [float(d[..., c].sum() / np.sqrt(var_tot[..., c].sum())) for c in range(3)]
    rep["global_rel"] = [float(d[..., c].sum() / mean[..., c].sum()) for c in range(3)]
    groups = {}
    for pid in np.unique(ids):
        m = ids == pid
        if m.sum() < 30:
            continue
        z = [float(d[m][:, c].sum() / np.sqrt(var_tot[m][:, c].sum() + 1e-30)) for c in range(3)]
        rel = float(d[m].sum() / max(mean[m].sum(), 1e-30))
        groups[int(pid)] = dict(n=int(m.sum()), z=z, rel=rel)
    # merge the many small prims into "other" for readability: report the 12 largest groups
    top = sorted(groups.items(), key=lambda kv: -kv[1]["n"])[:14]
    rep["groups"] = {k: v for k, v in top}
    print(json.dumps(rep, indent=1))
    with open(os.path.join(OUT, f"bias_{name}.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
