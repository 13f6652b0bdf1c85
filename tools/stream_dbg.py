"""Smallest possible run of the streaming kernel (debug builds with -DRT_STREAM_WATCHDOG): scene width spp [lib]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RT_B200_DEBUG", "1")
if len(sys.argv) > 4:
    os.environ["RT_B200_LIB"] = sys.argv[4]
import numpy as np
rtb = importlib.import_module("raytracing-practice_b200")
ctx = rtb.Context(0)
sc = rtb.Scene(sys.argv[1], 1)
cam = sc.camera_copy(image_width=int(sys.argv[2]), samples_per_pixel=int(sys.argv[3]))
ctx.upload_scene(sc.desc)
ctx.render(cam, seed=5, flags=rtb.RT_RENDER_MEGAKERNEL)
st = ctx.stats(); a0 = ctx.download_accum(); r0 = st.rays
print("mega  :", st.rays, "rays", f"{st.last_render_ms:.3f} ms", flush=True)
ctx.render(cam, seed=5, flags=rtb.RT_RENDER_STREAM)
st = ctx.stats(); a1 = ctx.download_accum()
print("stream:", st.rays, "rays", f"{st.last_render_ms:.3f} ms", "same" if np.array_equal(a0, a1) and r0 == st.rays else f"DIFF px={int((a0 != a1).any(axis=2).sum())} of {a0.shape[0] * a0.shape[1]}, sums {int(a0.sum())} vs {int(a1.sum())}", flush=True)
