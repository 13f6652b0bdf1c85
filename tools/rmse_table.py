"""Converged-image parity, measured and written down: for every scene at its FULL size, RMSE after write_color's gamma between
the GPU render and (a) the oracle driven by xoshiro256** and (b) the oracle driven by the reference's own glibc rand()
stream, next to the RMSE the two estimators' Monte-Carlo noise predicts, and the channel-mean z score (bias).
  python tools/rmse_table.py [gpu_spp=2048] [cpu_spp=64]      (under gpurun; uses all host cores for the oracle)"""
import importlib, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
rtb = importlib.import_module("raytracing-practice_b200")
from oracle import orc  # noqa: E402
from conftest import ALL_SCENES  # noqa: E402

GPU_SPP = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
CPU_SPP = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = rtb.Context(0)
rows = []
for name in ALL_SCENES:
    sc = rtb.Scene(name, rand_seed=1)
    cam = sc.camera_copy(samples_per_pixel=GPU_SPP)
    ctx.upload_scene(sc.desc)
    ctx.render(cam, seed=17)
    img = ctx.download_radiance(GPU_SPP).astype(np.float64)
    st = ctx.stats()
    row = dict(scene=name, width=cam.image_width, height=rtb.image_height(cam), gpu_spp=GPU_SPP, cpu_spp=CPU_SPP)
    for rng in ("xoshiro", "glibc"):
        t0 = time.time()
        mean, var, orays = orc.render_linear(sc.desc, cam, spp=CPU_SPP, seed=23, rng=rng)
        var_tot = var * (1.0 + CPU_SPP / GPU_SPP)
        z = [float((img[..., c].sum() - mean[..., c].sum()) / np.sqrt(var_tot[..., c].sum() + 1e-30)) for c in range(3)]
        g_gpu, g_cpu = np.sqrt(np.clip(img, 0, 0.999 ** 2)), np.sqrt(np.clip(mean, 0, 0.999 ** 2))
        rmse = np.sqrt(np.mean((g_gpu - g_cpu) ** 2, axis=(0, 1)))
        lin = np.maximum(0.5 * (img + mean), 1e-4)
        predicted = np.sqrt(np.mean(var_tot / (4.0 * lin), axis=(0, 1)))
        row[rng] = dict(rmse=[float(x) for x in rmse], predicted=[float(x) for x in predicted], ratio=float((rmse / np.maximum(predicted, 1e-12)).max()), z=z,
                        rays_per_sample_cpu=orays / (mean.shape[0] * mean.shape[1] * CPU_SPP), seconds=round(time.time() - t0, 1))
    row["rays_per_sample_gpu"] = st.rays / st.samples
    rows.append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
    sc.close()
with open(os.path.join(ROOT, "gpurun_out", "rmse_table.json"), "w") as f:
    json.dump(rows, f, indent=1)
print(f"| scene | size | RMSE after gamma vs oracle (xoshiro), max channel | predicted from both estimators' variance | ratio | mean z (r, g, b) | RMSE vs reference RNG (glibc) | ratio | mean z | rays/sample GPU | oracle |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    x, g = r["xoshiro"], r["glibc"]
    print(f"| {r['scene']} | {r['width']}x{r['height']} | {max(x['rmse']):.5f} | {max(x['predicted']):.5f} | {x['ratio']:.3f} | {', '.join(f'{v:+.2f}' for v in x['z'])} | "
          f"{max(g['rmse']):.5f} | {g['ratio']:.3f} | {', '.join(f'{v:+.2f}' for v in g['z'])} | {r['rays_per_sample_gpu']:.4f} | {x['rays_per_sample_cpu']:.4f} |")
