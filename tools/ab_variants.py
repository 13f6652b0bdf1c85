"""A/B kernel variants: build each -D combination into build/variants/, time them on the GPU.
  python tools/ab_variants.py build  "tag:-DRT_THREADS=768" "tag2:-DRT_NODE_THR=8" ...   (here, CPU)
  python tools/ab_variants.py run scene1,scene2 spp                                      (under gpurun)"""
import glob, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "build", "variants")
if sys.argv[1] == "build":
    os.makedirs(VDIR, exist_ok=True)
    for f in glob.glob(os.path.join(VDIR, "*.so")): os.remove(f)
    procs = []
    for spec in sys.argv[2:]:
        tag, _, flags = spec.partition(":")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-use_fast_math", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
               "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "raytracing-practice_b200", "csrc", "rt_b200.cu"), "-o", os.path.join(VDIR, f"{tag}.so")] + flags.split()
        procs.append((tag, subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True)))
    for tag, p in procs:
        err = p.communicate()[1]
        regs = [l for l in err.splitlines() if "Used" in l]
        k = [i for i, l in enumerate(err.splitlines()) if "render_kernelILb0ELb1" in l and "Compiling" in l]
        info = err.splitlines()[k[0] + 2].strip() if k else "?"
        spill = err.splitlines()[k[0] + 1 + 1 - 1].strip() if k else ""
        print(tag, p.returncode, info, "|", [l.strip() for l in err.splitlines()[k[0]+1:k[0]+3]][0] if k else "")
        ks = [i for i, l in enumerate(err.splitlines()) if "stream_kernelILb0" in l and "Compiling" in l]
        if ks: print("   stream:", err.splitlines()[ks[0] + 3].strip(), "|", err.splitlines()[ks[0] + 2].strip())
elif sys.argv[1] == "run":
    if len(sys.argv) > 4:  # child: one variant
        sys.path.insert(0, ROOT)
        rtb = importlib.import_module("raytracing-practice_b200")
        ctx = rtb.Context(0); out = {}
        for name in sys.argv[2].split(","):
            sc = rtb.Scene(name, 1); cam = sc.camera_copy(samples_per_pixel=int(sys.argv[3])); ctx.upload_scene(sc.desc)
            best = 1e30
            for rep in range(4):
                ctx.render(cam, seed=5); st = ctx.stats(); best = min(best, st.last_render_ms)
            out[name] = st.samples / best / 1e3
            if os.environ.get("AB_HASH"):
                import hashlib
                out[name + "#"] = hashlib.sha1(ctx.download_accum().tobytes()).hexdigest()[:8]
        print("RESULT", json.dumps(out))
    else:
        for so in sorted(glob.glob(os.path.join(VDIR, "*.so"))):
            env = dict(os.environ, RT_B200_LIB=so)
            r = subprocess.run([sys.executable, __file__, "run", sys.argv[2], sys.argv[3], "child"], env=env, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            res = json.loads(line[0][7:]) if line else {"error": r.stderr[-300:]}
            print(f"{os.path.basename(so):28s}", "  ".join(f"{k}:{v:8.1f}" if isinstance(v, float) else f"{k}:{v}" for k, v in res.items()), flush=True)
