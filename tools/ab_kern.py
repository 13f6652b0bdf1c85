"""Another render kernel vs the megakernel on the GPU: bit-exact accumulator check + timing, for the in-tree library and every
variant build (build/variants/*.so from tools/ab_variants.py).
  python tools/ab_kern.py refill|stream scene1,scene2 spp [width]        (under gpurun)"""
import glob, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KERN = sys.argv.pop(1)
if sys.argv[-1] == "child":  # one library
    import numpy as np
    rtb = importlib.import_module("raytracing-practice_b200")
    ctx = rtb.Context(0)
    out = {}
    width = int(sys.argv[3]) if len(sys.argv) > 4 and sys.argv[3].isdigit() else 0
    for name in sys.argv[1].split(","):
        sc = rtb.Scene(name, 1)
        kw = dict(samples_per_pixel=int(sys.argv[2]))
        if width:
            kw["image_width"] = width
        cam = sc.camera_copy(**kw)
        ctx.upload_scene(sc.desc)
        def run(flags):
            best = 1e30
            for rep in range(3):
                ctx.render(cam, seed=5, flags=flags)
                st = ctx.stats()
                best = min(best, st.last_render_ms)
            return ctx.download_accum(), st.rays, round(st.samples / best / 1e3, 1)
        a0, r0, v0 = run(rtb.RT_RENDER_MEGAKERNEL)
        try:
            a1, r1, v1 = run({"stream": rtb.RT_RENDER_STREAM, "refill": rtb.RT_RENDER_REFILL}[KERN])
            same = bool(np.array_equal(a0, a1)) and r0 == r1
            out[name] = [v0, v1, "same" if same else f"DIFF px={int((a0 != a1).any(axis=2).sum())} rays {r0} vs {r1} sum {int(a0.sum())} vs {int(a1.sum())}"]
        except rtb.RtError as e:
            out[name] = [v0, 0.0, f"{KERN} refused: {e}"]
    print("RESULT", json.dumps(out))
else:
    libs = [os.path.join(ROOT, "raytracing-practice_b200", "librt_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
    for so in libs:
        env = dict(os.environ, RT_B200_LIB=so)
        try:
            r = subprocess.run([sys.executable, __file__, KERN] + sys.argv[1:] + ["child"], env=env, capture_output=True, text=True, timeout=150)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            res = json.loads(line[0][7:]) if line else {"error": r.stderr[-600:]}
        except subprocess.TimeoutExpired:
            res = {"error": "TIMEOUT (hung kernel?)"}
        print(f"{os.path.basename(so):24s}", "  ".join(f"{k}: mega {v[0]:7.1f} {KERN} {v[1]:7.1f} ({v[1] / max(v[0], 1e-9):.2f}x) {v[2]}" if isinstance(v, list) else f"{k}:{v}" for k, v in res.items()), flush=True)
