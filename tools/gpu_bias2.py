"""Bias dissection on Cornell-smoke variants (run under gpurun)."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rtb = importlib.import_module("raytracing-practice_b200")
from oracle import orc
import scene_util as su

def room(s, lamp=7.0):
    red, white, green = (s.lambertian(s.solid(.65,.05,.05)), s.lambertian(s.solid(.73,.73,.73)), s.lambertian(s.solid(.12,.45,.15)))
    lt = s.light(s.solid(lamp, lamp, lamp))
    L = 555.0
    k = [s.quad((L,0,0),(0,L,0),(0,0,L),green), s.quad((0,0,0),(0,L,0),(0,0,L),red), s.quad((113,554,127),(330,0,0),(0,0,305),lt),
         s.quad((0,0,0),(L,0,0),(0,0,L),white), s.quad((L,L,L),(-L,0,0),(0,0,-L),white), s.quad((0,0,L),(L,0,0),(0,L,0),white)]
    return k, white

def variant(name):
    s = su.SceneDesc()
    k, white = room(s)
    def placed(w,h,d,ang,at):
        b = s.box((0,0,0),(w,h,d),white)
        if ang is not None: b = s.rotate_y(b, ang)
        return s.translate(b, at)
    if name == "room": pass
    elif name == "black_rot": k.append(s.medium(placed(165,330,165,15,(265,0,295)), 0.01, s.isotropic(s.solid(0,0,0))))
    elif name == "white_rot": k.append(s.medium(placed(165,165,165,-18,(130,0,65)), 0.01, s.isotropic(s.solid(1,1,1))))
    elif name == "white_axis": k.append(s.medium(placed(165,165,165,None,(130,0,65)), 0.01, s.isotropic(s.solid(1,1,1))))
    elif name == "white_sphere": k.append(s.medium(s.sphere((212,90,150),90,white), 0.01, s.isotropic(s.solid(1,1,1))))
    elif name == "solid_rot": k.append(placed(165,165,165,-18,(130,0,65)))
    elif name == "dense_axis": k.append(s.medium(placed(165,165,165,None,(130,0,65)), 0.1, s.isotropic(s.solid(1,1,1))))
    else: raise SystemExit(name)
    return s, s.finish(s.list(k))

def main():
    ctx = rtb.Context(0)
    out = {}
    for name in sys.argv[1].split(","):
        for depth in [int(x) for x in sys.argv[2].split(",")]:
            s, desc = variant(name)
            cam = su.camera(width=80, aspect=1.0, spp=16384, depth=depth, bg=(0,0,0), vfov=40.0, lookfrom=(278,278,-800), lookat=(278,278,0))
            ctx.upload_scene(desc)
            ctx.render(cam, seed=5)
            img = ctx.download_radiance(cam.samples_per_pixel).astype(np.float64)
            st = ctx.stats()
            mean, var, orays = orc.render_linear(desc, cam, spp=2048, seed=13)
            vt = var * (1 + 2048/16384)
            d = img - mean
            z = [float(d[...,c].sum()/np.sqrt(vt[...,c].sum())) for c in range(3)]
            rel = [float(d[...,c].sum()/mean[...,c].sum()) for c in range(3)]
            r = dict(z=[round(x,2) for x in z], rel=[round(x,5) for x in rel], rps_gpu=round(st.rays/st.samples,4), rps_cpu=round(orays/(80*80*2048),4))
            print(name, depth, json.dumps(r), flush=True)
            out[f"{name}_{depth}"] = r
    json.dump(out, open(os.path.join(ROOT,"gpurun_out","bias2.json"),"w"), indent=1)
main()
