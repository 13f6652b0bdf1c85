"""One render of one scene (for ncu captures): python tools/prof_one.py <scene> <spp> [mega|wf] [width]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-practice_b200")
name, spp = sys.argv[1], int(sys.argv[2])
kern = sys.argv[3] if len(sys.argv) > 3 else "default"
sc = rtb.Scene(name, 1)
over = dict(samples_per_pixel=spp)
if len(sys.argv) > 4:
    over["image_width"] = int(sys.argv[4])
cam = sc.camera_copy(**over)
ctx = rtb.Context(0)
ctx.upload_scene(sc.desc)
flags = {"mega": rtb.RT_RENDER_MEGAKERNEL, "count": rtb.RT_RENDER_COUNTERS, "stream": rtb.RT_RENDER_STREAM, "refill": rtb.RT_RENDER_REFILL}.get(kern, 0)
ctx.render(cam, seed=5, flags=flags)
st = ctx.stats()
print(f"{name} {kern}: {st.samples / st.last_render_ms / 1e3:.1f} Msamples/s, {st.rays / st.last_render_ms / 1e3:.1f} Mrays/s, {st.last_render_ms:.2f} ms")
