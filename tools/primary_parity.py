"""Pixel-centre parity of every traversal the library has, on all scenes at their FULL size, against the CPU oracle
(test infrastructure): mismatching ids, relative t error and normal error (max and p99.9) for
  exact   rt_primary_visibility(RT_TRACE_EXACT)          fp32 conservative traversal + fp64 re-evaluation (parity kernel)
  fp32    rt_primary_visibility(RT_TRACE_FP32)           trace_kernel -> closest_hit<false> (global-memory leaves)
  render  rt_primary_visibility(RT_TRACE_RENDER_KERNEL)  render_kernel<.., AOV>: the kernel, staging and node/leaf steps rt_render uses
Writes gpurun_out/primary_parity.json and a markdown table on stdout.   python tools/primary_parity.py   (under gpurun)"""
import importlib, json, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
rtb = importlib.import_module("raytracing-practice_b200")
from oracle import orc  # noqa: E402
from conftest import ALL_SCENES  # noqa: E402


def stats(ids, t, nrm, oids, ot, onrm):
    mism = ids != oids
    ok = (~mism) & (oids >= 0)
    rel = np.abs(t[ok] - ot[ok]) / np.abs(ot[ok]) if ok.any() else np.zeros(1)
    dn = np.abs(nrm[ok] - onrm[ok]).max(axis=1) if ok.any() else np.zeros(1)
    return dict(px=int(ids.size), id_mismatch=int(mism.sum()), t_rel_max=float(rel.max()), t_rel_p999=float(np.quantile(rel, 0.999)),
                t_rel_over_1e5=int((rel > 1e-5).sum()), n_abs_max=float(dn.max()), n_abs_p999=float(np.quantile(dn, 0.999)), n_over_1e5=int((dn > 1e-5).sum()))


def main():
    ctx = rtb.Context(0)
    out = {}
    modes = (("exact", rtb.RT_TRACE_EXACT | rtb.RT_TRACE_SKIP_MEDIA), ("fp32", rtb.RT_TRACE_FP32 | rtb.RT_TRACE_SKIP_MEDIA),
             ("render", rtb.RT_TRACE_RENDER_KERNEL | rtb.RT_TRACE_SKIP_MEDIA))
    for name in ALL_SCENES:
        sc = rtb.Scene(name, rand_seed=1)
        ctx.upload_scene(sc.desc)
        cam = sc.camera_copy()
        oids, ot, onrm = orc.primary(sc.desc, cam, skip_media=True)
        out[name] = {}
        res = {}
        for tag, flags in modes:
            ids, t, nrm = ctx.primary_visibility(cam, flags)
            res[tag] = ids
            out[name][tag] = stats(ids, t, nrm, oids, ot, onrm)
        out[name]["render_vs_fp32_id_mismatch"] = int((res["render"] != res["fp32"]).sum())
        print(name, json.dumps(out[name]), file=sys.stderr, flush=True)
        sc.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "primary_parity.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("| scene | pixels | mode | id mismatches | t rel. err max | t p99.9 | px with t err > 1e-5 | normal abs. err max | normal p99.9 | px with normal err > 1e-5 |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for name, r in out.items():
        for tag in ("exact", "fp32", "render"):
            s = r[tag]
            print(f"| {name} | {s['px']} | {tag} | {s['id_mismatch']} | {s['t_rel_max']:.2e} | {s['t_rel_p999']:.2e} | {s['t_rel_over_1e5']} | {s['n_abs_max']:.2e} | {s['n_abs_p999']:.2e} | {s['n_over_1e5']} |")
    print("\nrender vs fp32 id mismatches:", {k: v["render_vs_fp32_id_mismatch"] for k, v in out.items()})


if __name__ == "__main__":
    main()
