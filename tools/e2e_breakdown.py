"""Wall-clock breakdown of the host-buffer call sequence bench.py times as e2e."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-practice_b200")
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = rtb.Context(0)
sc = rtb.Scene("book2_final", 1)
cam = sc.camera_copy(samples_per_pixel=spp)
for rep in range(3):
    t = [time.time()]
    ctx.upload_scene(sc.desc); t.append(time.time())
    ctx.render(cam, seed=rep); t.append(time.time())
    ctx.synchronize(); t.append(time.time())
    rgb = ctx.download_rgb8(spp); t.append(time.time())
    st = ctx.stats()
    print(f"upload {1e3*(t[1]-t[0]):.1f} ms, render call {1e3*(t[2]-t[1]):.1f} ms, sync {1e3*(t[3]-t[2]):.1f} ms (device {st.last_render_ms:.1f} ms), download {1e3*(t[4]-t[3]):.1f} ms", flush=True)
