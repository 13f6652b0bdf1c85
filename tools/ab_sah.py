"""Scene-build constants (read by rt_upload_scene from the environment): time the megakernel for each setting.
  python tools/ab_sah.py <scene,...> <spp> VAR=v1,v2 [VAR2=w1,w2]"""
import importlib, itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-practice_b200")
names, spp = sys.argv[1].split(","), int(sys.argv[2])
axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[3:]]
ctx = rtb.Context(0)
for combo in itertools.product(*[v for _, v in axes]):
    for (k, _), v in zip(axes, combo):
        os.environ[k] = v
    out = []
    for name in names:
        sc = rtb.Scene(name, 1)
        cam = sc.camera_copy(samples_per_pixel=spp)
        ctx.upload_scene(sc.desc)
        best = 1e30
        for rep in range(3):
            ctx.render(cam, seed=5, flags=rtb.RT_RENDER_MEGAKERNEL)
            st = ctx.stats()
            best = min(best, st.last_render_ms)
        out.append(f"{name} {st.samples / best / 1e3:7.1f} (nodes {st.n_nodes})")
    print(" ".join(f"{k}={v}" for (k, _), v in zip(axes, combo)), "|", "  ".join(out), flush=True)
