"""Time one kernel under several environment settings (read by rt_render at every call).
  python tools/ab_env.py <scene> <spp> <mega|stream|refill> VAR=v1,v2,v3 [lib.so]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 5:
    os.environ["RT_B200_LIB"] = sys.argv[5]
rtb = importlib.import_module("raytracing-practice_b200")
name, spp, kern = sys.argv[1], int(sys.argv[2]), sys.argv[3]
var, vals = sys.argv[4].split("=")
flags = {"mega": rtb.RT_RENDER_MEGAKERNEL, "stream": rtb.RT_RENDER_STREAM, "refill": rtb.RT_RENDER_REFILL}[kern]
ctx = rtb.Context(0)
sc = rtb.Scene(name, 1)
cam = sc.camera_copy(samples_per_pixel=spp)
ctx.upload_scene(sc.desc)
for v in vals.split(","):
    os.environ[var] = v
    best = 1e30
    for rep in range(3):
        ctx.render(cam, seed=5, flags=flags)
        st = ctx.stats()
        best = min(best, st.last_render_ms)
    print(f"{name} {kern} {var}={v}: {st.samples / best / 1e3:.1f} Msamples/s", flush=True)
