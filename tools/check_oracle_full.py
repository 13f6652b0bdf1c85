"""The oracle against the reference at SHIPPED size: renders all seven shipped scenes (src/main.cpp:12-346) with
oracle/liboracle.so exactly as main would (glibc rand() stream continuing after scene construction, shipped width / spp /
depth) and compares the P3 file's sha256 and the ray count with the pins the UNMODIFIED reference produced
(tests/golden/ref_pins.json "shipped", made by tests/golden/make_golden.py --full from oracle/_ref/ref_harness).
CPU only, serial by nature (one rand() stream): ~5 minutes, cornell_box alone is 237 M rays.
  python tools/check_oracle_full.py [scene ...]  > profiles/logs/check_oracle_full.log"""
import hashlib, importlib, json, os, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-practice_b200")
from oracle import orc  # noqa: E402

pins = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_pins.json")))["shipped"]
names = sys.argv[1:] or list(pins)
bad = 0
with tempfile.TemporaryDirectory() as tmp:
    for name in names:
        want = pins[name]
        sc = rtb.Scene(name, rand_seed=1)
        path = os.path.join(tmp, "o.ppm")
        t0 = time.time()
        rays = orc.render_ppm(sc.desc, sc.cam.contents, path)
        got = hashlib.sha256(open(path, "rb").read()).hexdigest()
        ok = got == want["ppm_sha256"] and rays == want["rays"]
        bad += not ok
        print(f"{name:18s} {want['width']}x{want['height']} x {want['spp']} spp: sha256 {got[:16]}... {'==' if got == want['ppm_sha256'] else '!='} reference {want['ppm_sha256'][:16]}..., "
              f"rays {rays} {'==' if rays == want['rays'] else '!='} {want['rays']}  ({time.time() - t0:.1f} s)  {'OK' if ok else 'MISMATCH'}", flush=True)
        sc.close()
print("all seven shipped-size renders of the oracle equal the reference's" if bad == 0 and len(names) == len(pins) else f"{bad} mismatches / {len(names)} scenes checked")
sys.exit(1 if bad else 0)
