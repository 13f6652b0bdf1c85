"""Generate tests/golden/ref_pins.json from the UNMODIFIED reference compiled as
oracle/_ref/ref_harness (recipe: oracle/Makefile; needs /root/reference, so run it in the build
container, not on the GPU box).  The committed JSON is what the tests read.

Every pin is an output of the reference's own code (camera::render, hittable::hit) on a scene
built through its own classes:
  * ppm      : sha256 + ray count of camera::render's P3 output at a reduced width/spp/depth
               (fresh process, unseeded rand() exactly like `main`)
  * primary  : pixel-centre primary pass: FNV-1a64 of the leaf ids,
               sha256 of the raw little-endian doubles of t and of the normals
  * shipped  : the SURVEY.md §4 pins of the seven scenes at their shipped size (re-measured)
  * earthmap : sha256 of the RGB8 texels PIL (libjpeg-turbo) decodes from images/earthmap.jpg
"""
import hashlib
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
ENV = dict(os.environ, RTW_IMAGES="/root/reference/images")

SMALL = dict(width=64, spp=2, depth=6)
SCENES = ["bouncing_spheres", "checkered_spheres", "earth", "perlin_sphere", "quads", "simple_light", "cornell_box",
          "book1_final", "cornell_rotated", "cornell_smoke", "book2_final"]
SHIPPED = ["bouncing_spheres", "checkered_spheres", "earth", "perlin_sphere", "quads", "simple_light", "cornell_box"]
FULL_PPM = {"earth", "checkered_spheres", "quads", "simple_light", "perlin_sphere", "bouncing_spheres", "cornell_box"}


def run(mode, scene, out, **opts):
    cmd = [HARNESS, mode, scene, out]
    for k, v in opts.items():
        cmd += [f"--{k}", str(v)]
    p = subprocess.run(cmd, env=ENV, capture_output=True, text=True, check=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("JSON ")][-1]
    return json.loads(line[5:])


def sha(path_or_bytes):
    data = open(path_or_bytes, "rb").read() if isinstance(path_or_bytes, str) else path_or_bytes
    return hashlib.sha256(data).hexdigest()


def read_primary(path):
    raw = open(path, "rb").read()
    w, h, prims, mismatch = struct.unpack("<4i", raw[:16])
    n = w * h
    ids = np.frombuffer(raw, "<i4", n, 16)
    t = np.frombuffer(raw, "<f8", n, 16 + 4 * n)
    nrm = np.frombuffer(raw, "<f8", 3 * n, 16 + 12 * n)
    return dict(width=w, height=h, prims=prims, replay_mismatch=mismatch, ids_fnv=orc.fnv1a64_ids(ids), t_sha256=sha(t.tobytes()),
                n_sha256=sha(nrm.tobytes()), n_background=int((ids < 0).sum()), n_distinct=int(len(set(ids[ids >= 0].tolist()))))


def main():
    full = "--full" in sys.argv
    pins = {"small": SMALL, "scenes": {}, "shipped": {}}
    old = {}
    out_path = os.path.join(HERE, "ref_pins.json")
    if os.path.exists(out_path):
        old = json.load(open(out_path))
    with tempfile.TemporaryDirectory() as tmp:
        for s in SCENES:
            e = {}
            j = run("ppm", s, f"{tmp}/a.ppm", **SMALL)
            e["ppm"] = dict(sha256=sha(f"{tmp}/a.ppm"), rays=j["rays"], width=j["width"], height=j["height"])
            run("primary", s, f"{tmp}/p.bin", width=SMALL["width"])
            e["primary_small"] = read_primary(f"{tmp}/p.bin")
            run("primary", s, f"{tmp}/p.bin")
            e["primary_full"] = read_primary(f"{tmp}/p.bin")
            pins["scenes"][s] = e
            print(s, e["ppm"]["sha256"][:16], e["primary_full"]["ids_fnv"], flush=True)
        for s in SHIPPED:
            # the reference's own src/main.cpp scene function, shipped width/spp/depth
            prev = old.get("shipped", {}).get(s)
            if prev and not full:
                pins["shipped"][s] = prev
                continue
            j = run("ppm", "shipped:" + s, f"{tmp}/s.ppm")
            run("primary", "shipped:" + s, f"{tmp}/p.bin")
            pins["shipped"][s] = dict(ppm_sha256=sha(f"{tmp}/s.ppm"), rays=j["rays"], width=j["width"], height=j["height"], spp=j["spp"],
                                      max_depth=j["max_depth"], seconds_o2=j["seconds"], primary=read_primary(f"{tmp}/p.bin"))
            print("shipped", s, pins["shipped"][s]["ppm_sha256"][:16], j["rays"], f"{j['seconds']:.1f}s", flush=True)
    from PIL import Image

    im = np.asarray(Image.open("/root/reference/images/earthmap.jpg").convert("RGB"))
    pins["earthmap"] = dict(width=int(im.shape[1]), height=int(im.shape[0]), rgb8_sha256=sha(im.tobytes()),
                            jpeg_sha256=sha("/root/reference/images/earthmap.jpg"))
    json.dump(pins, open(out_path, "w"), indent=1, sort_keys=True)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
