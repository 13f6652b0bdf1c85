"""material::scatter on the device vs the oracle (core/material.hpp): RNG streams differ by
design (Philox vs glibc rand), so parity is distributional; deterministic parts are exact."""
import numpy as np
import pytest

import scene_util as su

pytestmark = pytest.mark.gpu
N = 400_000


def _setup(gpu_ctx):
    s = su.SceneDesc()
    lam = s.lambertian(s.solid(0.3, 0.6, 0.9))
    met0 = s.metal((0.7, 0.6, 0.5), 0.0)
    met = s.metal((0.8, 0.8, 0.9), 0.6)
    die = s.dielectric(1.5)
    lig = s.light(s.solid(4, 5, 6))
    iso = s.isotropic(s.solid(0.2, 0.4, 0.9))
    desc = s.finish(s.sphere((0, 0, 0), 1, lam))
    gpu_ctx.upload_scene(desc)
    return desc, dict(lam=lam, met0=met0, met=met, die=die, lig=lig, iso=iso)


def _inputs(rng, n, grazing=False):
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    d = rng.normal(size=(n, 3))
    d -= np.sum(d * nrm, 1, keepdims=True) * nrm * (1.0 if grazing else 0.0)
    # make the incoming direction face the (already flipped) normal: dot(d, n) < 0
    flip = np.sum(d * nrm, 1) > 0
    d[flip] *= -1
    d *= rng.uniform(0.2, 3.0, (n, 1))  # un-normalised
    return d, nrm


def test_lambertian_and_isotropic_distributions(rtb, orc, gpu_ctx):
    desc, m = _setup(gpu_ctx)
    rng = np.random.default_rng(1)
    n_fixed = np.tile([[0.0, 0.6, 0.8]], (N, 1))
    d_in = np.tile([[0.3, -1.0, -0.2]], (N, 1))
    ff = np.ones(N, np.uint8)
    for key, mean_expected in (("lam", n_fixed[0]), ("iso", np.zeros(3))):
        g_dir, g_att, g_ok = gpu_ctx.eval_scatter(m[key], d_in, n_fixed, ff, seed=5)
        o_dir, o_att, o_ok = orc.scatter(desc, m[key], d_in[:100_000], n_fixed[:100_000], ff[:100_000])
        assert g_ok.all() and o_ok.all()
        assert np.allclose(g_att, o_att[0], atol=1e-6)
        # first and second moments: E[dir] = n (+0 for isotropic), E[dir dir^T] = n n^T + I/3
        assert np.all(np.abs(g_dir.mean(0) - mean_expected) < 5 * 0.58 / np.sqrt(N))
        for a in range(3):
            for b in range(3):
                want = mean_expected[a] * mean_expected[b] + (1 / 3 if a == b else 0)
                assert abs((g_dir[:, a] * g_dir[:, b]).mean() - want) < 6 / np.sqrt(N) * 1.5
                assert abs((o_dir[:, a] * o_dir[:, b]).mean() - want) < 6 / np.sqrt(100_000) * 1.5
        # the random part is a UNIT vector
        assert np.allclose(np.linalg.norm(g_dir - mean_expected, axis=1), 1.0, atol=1e-5)


def test_metal(rtb, orc, gpu_ctx):
    desc, m = _setup(gpu_ctx)
    rng = np.random.default_rng(2)
    d, nrm = _inputs(rng, N)
    ff = np.ones(N, np.uint8)
    # fuzz 0: deterministic mirror direction, unit length (material.hpp:89-92)
    g_dir, g_att, g_ok = gpu_ctx.eval_scatter(m["met0"], d, nrm, ff)
    refl = d - 2 * np.sum(d * nrm, 1, keepdims=True) * nrm
    refl /= np.linalg.norm(refl, axis=1, keepdims=True)
    assert np.allclose(g_dir, refl, atol=2e-6)
    assert g_ok.all() and np.allclose(g_att, [0.7, 0.6, 0.5], atol=1e-6)
    # fuzz 0.6 at grazing incidence: the absorbed fraction must match the oracle's
    d, nrm = _inputs(rng, N, grazing=True)
    d -= 0.15 * np.linalg.norm(d, axis=1, keepdims=True) * nrm
    g_dir, _, g_ok = gpu_ctx.eval_scatter(m["met"], d, nrm, ff)
    o_dir, _, o_ok = orc.scatter(desc, m["met"], d[:100_000], nrm[:100_000], ff[:100_000])
    p_g, p_o = g_ok.mean(), o_ok.mean()
    assert 0.55 < p_o < 0.95
    assert abs(p_g - p_o) < 5 * np.sqrt(p_o * (1 - p_o) * (1 / N + 1 / 100_000))
    assert np.all(np.sum(g_dir[g_ok == 1] * nrm[g_ok == 1], 1) > 0)


def test_dielectric(rtb, orc, gpu_ctx):
    desc, m = _setup(gpu_ctx)
    rng = np.random.default_rng(3)
    d, nrm = _inputs(rng, N)
    for front in (1, 0):
        ff = np.full(N, front, np.uint8)
        g_dir, g_att, g_ok = gpu_ctx.eval_scatter(m["die"], d, nrm, ff, seed=9)
        o_dir, _, _ = orc.scatter(desc, m["die"], d[:100_000], nrm[:100_000], ff[:100_000])
        assert g_ok.all() and np.allclose(g_att, 1.0)
        ri = 1 / 1.5 if front else 1.5
        ud = d / np.linalg.norm(d, axis=1, keepdims=True)
        cos_t = np.minimum(-np.sum(ud * nrm, 1), 1.0)
        sin_t = np.sqrt(1 - cos_t ** 2)
        refl = ud - 2 * np.sum(ud * nrm, 1, keepdims=True) * nrm
        perp = ri * (ud + cos_t[:, None] * nrm)
        refr = perp - np.sqrt(np.abs(1 - np.sum(perp * perp, 1)))[:, None] * nrm
        is_refl = np.linalg.norm(g_dir - refl, axis=1) < 1e-4
        is_refr = np.linalg.norm(g_dir - refr, axis=1) < 1e-4
        assert np.all(is_refl | is_refr)
        tir = ri * sin_t > 1.0
        assert np.all(is_refl[tir & ~is_refr])
        assert not np.any(is_refr[tir] & ~is_refl[tir])
        # Schlick reflect probability (material.hpp:198-206) over the non-TIR rays
        r0 = ((1 - ri) / (1 + ri)) ** 2
        p = r0 + (1 - r0) * (1 - cos_t) ** 5
        sel = ~tir & (np.linalg.norm(refl - refr, axis=1) > 1e-3)
        expected = p[sel].sum()
        sigma = np.sqrt((p[sel] * (1 - p[sel])).sum())
        assert abs(is_refl[sel].sum() - expected) < 5 * sigma
        o_is_refl = np.linalg.norm(o_dir - refl[:100_000], axis=1) < 1e-9
        sel_o = sel[:100_000]
        assert abs(o_is_refl[sel_o].sum() - p[:100_000][sel_o].sum()) < 5 * np.sqrt((p[:100_000][sel_o] * (1 - p[:100_000][sel_o])).sum())


def test_diffuse_light_never_scatters(rtb, orc, gpu_ctx):
    desc, m = _setup(gpu_ctx)
    rng = np.random.default_rng(4)
    d, nrm = _inputs(rng, 1000)
    _, _, ok = gpu_ctx.eval_scatter(m["lig"], d, nrm, np.ones(1000, np.uint8))
    assert not ok.any()
