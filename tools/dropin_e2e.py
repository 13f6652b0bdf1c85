"""End to end through the REAL drop-in call: process start -> camera::render(std::ofstream, world) -> PPM closed, for every
BASELINE configuration at its own size, next to the unmodified reference (oracle/_ref/ref_harness ppm: process start ->
camera::render -> PPM, on a bounded spp and scaled).  RT_B200_TIMING gives camera::render's own breakdown.
  python tools/dropin_e2e.py [devices, e.g. 0 or 0,1,2,3,4,5,6,7]       (under gpurun; build/dropin from __graft_entry__.build())"""
import importlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-practice_b200")
CONFIGS = [("C1", "book1_final", 1200, 10, 1), ("C2", "bouncing_spheres", 400, 100, 4), ("C3a", "earth", 400, 100, 8), ("C3b", "perlin_sphere", 400, 100, 8),
           ("C4", "cornell_smoke", 600, 200, 2), ("C5", "book2_final", 800, 10000, 1)]
devices = sys.argv[1] if len(sys.argv) > 1 else "0"
exe = os.path.join(ROOT, "build", "dropin")
ref = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
env = dict(os.environ, RT_B200_DEVICES=devices, RT_B200_TIMING="1")
if rtb.default_image_dir():
    env["RTW_IMAGES"] = rtb.default_image_dir()
out = "/dev/shm/dropin_e2e.ppm" if os.path.isdir("/dev/shm") else "/tmp/dropin_e2e.ppm"
subprocess.run([exe, "quads", out, "64", "4"], env=env, capture_output=True)  # page the binaries in
for tag, scene, width, spp, ref_spp in CONFIGS:
    row = dict(config=tag, scene=scene, width=width, spp=spp, devices=devices)
    for label, repeats in (("one_render", 1), ("three_renders", 3)):
        t0 = time.time()
        r = subprocess.run([exe, scene, out, str(width), str(spp), str(repeats)], env=env, capture_output=True, text=True)
        wall = time.time() - t0
        if r.returncode != 0:
            row[label] = {"error": r.stderr[-300:]}
            continue
        parts = [json.loads(l[len("RTB200_TIMING "):]) for l in r.stderr.splitlines() if l.startswith("RTB200_TIMING ")]
        row[label] = dict(process_wall_ms=round(wall * 1e3, 1), renders=parts)
    if os.path.exists(ref) and devices == "0":
        t0 = time.time()
        r = subprocess.run([ref, "ppm", scene, out, "--spp", str(ref_spp), "--seed", "1", "--width", str(width)], env=env, capture_output=True, text=True)
        wall = time.time() - t0
        j = [json.loads(l[5:]) for l in r.stdout.splitlines() if l.startswith("JSON ")]
        if j:
            row["reference"] = dict(spp=ref_spp, process_wall_ms=round(wall * 1e3, 1), render_seconds=j[-1]["seconds"],
                                    full_job_seconds_extrapolated=round(j[-1]["seconds"] * spp / ref_spp + (wall - j[-1]["seconds"]), 1))
    print(json.dumps(row), flush=True)
