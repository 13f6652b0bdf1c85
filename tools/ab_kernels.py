"""Pool kernel vs megakernel on the GPU: bit-exact accumulator check + timing, for every variant build.
  python tools/ab_kernels.py scene1,scene2 spp            (under gpurun; variants from build/variants/*.so
                                                           built by tools/ab_variants.py, plus the in-tree lib)"""
import glob, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 3:  # child: one library
    import numpy as np
    rtb = importlib.import_module("raytracing-practice_b200")
    ctx = rtb.Context(0)
    out = {}
    for name in sys.argv[1].split(","):
        sc = rtb.Scene(name, 1)
        cam = sc.camera_copy(samples_per_pixel=int(sys.argv[2]))
        ctx.upload_scene(sc.desc)
        def run(flags):
            best = 1e30
            for rep in range(3):
                ctx.render(cam, seed=5, flags=flags)
                st = ctx.stats()
                best = min(best, st.last_render_ms)
            return ctx.download_accum(), st.rays, round(st.samples / best / 1e3, 1)
        a0, r0, v0 = run(rtb.RT_RENDER_MEGAKERNEL)
        a1, r1, v1 = run(rtb.RT_RENDER_POOL)
        same = bool(np.array_equal(a0, a1)) and r0 == r1
        out[name] = [v0, v1, "same" if same else f"DIFF px={int((a0 != a1).any(axis=2).sum())} rays {r0} vs {r1}"]
    print("RESULT", json.dumps(out))
else:
    libs = [os.path.join(ROOT, "raytracing-practice_b200", "librt_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
    for so in libs:
        env = dict(os.environ, RT_B200_LIB=so)
        r = subprocess.run([sys.executable, __file__, sys.argv[1], sys.argv[2], "child"], env=env, capture_output=True, text=True, timeout=100)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        res = json.loads(line[0][7:]) if line else {"error": r.stderr[-400:]}
        print(f"{os.path.basename(so):24s}", "  ".join(f"{k}: mega {v[0]:7.1f} pool {v[1]:7.1f} {v[2]}" if isinstance(v, list) else f"{k}:{v}" for k, v in res.items()), flush=True)
