"""T2 — primary-visibility parity: a jitter-free pass at pixel centres must give
the SAME hit object id as the reference's hittable::hit for every pixel (bit-exact), with t and
normal within 1e-5 relative.  Checked against the oracle AND directly against the reference's
own output hashes in tests/golden/ref_pins.json."""
import numpy as np
import pytest

from conftest import ALL_SCENES

pytestmark = pytest.mark.gpu

T_RTOL = 1e-5  # north_star's tolerance for t and normal (fp32 device vs fp64 reference)


@pytest.mark.parametrize("name", ALL_SCENES)
def test_exact_primary_ids_match_reference(rtb, orc, pins, gpu_ctx, name):
    sc = rtb.Scene(name, rand_seed=1)
    gpu_ctx.upload_scene(sc.desc)
    cam = sc.camera_copy()
    ids, t, nrm = gpu_ctx.primary_visibility(cam, rtb.RT_TRACE_EXACT | rtb.RT_TRACE_SKIP_MEDIA)
    want = pins["scenes"][name]["primary_full"]
    assert ids.shape == (want["height"], want["width"])
    # 1) the reference's own id map (hash of the harness output of the unmodified reference)
    assert orc.fnv1a64_ids(ids) == want["ids_fnv"]
    assert int((ids < 0).sum()) == want["n_background"]
    # 2) the oracle, value by value
    oids, ot, onrm = orc.primary(sc.desc, cam, skip_media=True)
    assert np.array_equal(ids, oids)
    hit = oids >= 0
    assert np.all(np.abs(t[hit] - ot[hit]) <= T_RTOL * np.abs(ot[hit]))
    assert np.all(np.abs(nrm[hit] - onrm[hit]) <= T_RTOL)
    assert np.all(np.isinf(t[~hit]))


@pytest.mark.parametrize("name", ALL_SCENES)
def test_fp32_production_traversal_agrees(rtb, orc, gpu_ctx, name):
    """The fp32 traversal used by rt_render (no fp64 refinement): ids may flip only on a
    handful of silhouette/edge pixels, t stays close."""
    sc = rtb.Scene(name, rand_seed=1)
    gpu_ctx.upload_scene(sc.desc)
    cam = sc.camera_copy()
    ids, t, nrm = gpu_ctx.primary_visibility(cam, rtb.RT_TRACE_FP32 | rtb.RT_TRACE_SKIP_MEDIA)
    oids, ot, onrm = orc.primary(sc.desc, cam, skip_media=True)
    mism = ids != oids
    assert mism.mean() <= 1e-3, f"{mism.sum()} of {ids.size} pixels differ"
    ok = (~mism) & (oids >= 0)
    rel = np.abs(t[ok] - ot[ok]) / np.abs(ot[ok])
    assert np.quantile(rel, 0.999) <= 2e-4 and rel.max() <= 1e-2
    # normals: ignore the few grazing hits on the radius-1000 spheres where fp32 p drifts
    dn = np.abs(nrm[ok] - onrm[ok]).max(axis=1)
    assert np.quantile(dn, 0.999) <= 2e-3


def test_primary_pass_small_widths_and_aspect(rtb, orc, gpu_ctx):
    """Ragged image sizes: width 1, odd widths, height clamped to >= 1 (camera.hpp:79-80)."""
    sc = rtb.Scene("quads", rand_seed=1)
    gpu_ctx.upload_scene(sc.desc)
    for width, aspect in [(1, 1.0), (7, 16.0 / 9.0), (33, 0.5), (3, 100.0)]:
        cam = sc.camera_copy(image_width=width, aspect_ratio=aspect)
        ids, t, nrm = gpu_ctx.primary_visibility(cam)
        oids, ot, _ = orc.primary(sc.desc, cam)
        assert ids.shape == oids.shape and ids.shape[0] >= 1
        assert np.array_equal(ids, oids)
