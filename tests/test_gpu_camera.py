"""camera::get_ray (camera.hpp:139-177) on the device, tested for what it SAMPLES rather than through converged images of
the book scenes: the defocus disk is a uniform disk of radius focus_dist * tan(defocus_angle / 2) (the closed-form
r = sqrt(u) sampling that replaces the reference's rejection loop, identical in law), the pixel jitter is uniform on the
pixel's square, and the shutter time is uniform on [0, 1).  Each is read off the image of a small emitter, against the
analytic footprint of a thin lens / a box filter / a uniform streak."""
import math

import numpy as np
import pytest

import scene_util as su

pytestmark = pytest.mark.gpu


def _moments(img):
    w = img.sum(axis=2).astype(np.float64)
    ys, xs = np.mgrid[0:w.shape[0], 0:w.shape[1]]
    tot = w.sum()
    cx, cy = (w * xs).sum() / tot, (w * ys).sum() / tot
    dx, dy = xs - cx, ys - cy
    m2x, m2y = (w * dx * dx).sum() / tot, (w * dy * dy).sum() / tot
    rho2 = dx * dx + dy * dy
    return cx, cy, m2x, m2y, (w * rho2 * rho2).sum() / tot / ((w * rho2).sum() / tot) ** 2


def test_defocus_disk_is_a_uniform_disk_of_the_right_radius(rtb, gpu_ctx):
    W, vfov, focus, angle, s_dist, r_e = 201, 20.0, 10.0, 6.0, 5.0, 0.02
    s = su.SceneDesc()
    desc = s.finish(s.list([s.sphere((0, 0, -s_dist), r_e, s.light(s.solid(50, 50, 50)))]))
    cam = su.camera(width=W, spp=4096, depth=2, bg=(0, 0, 0), vfov=vfov, lookfrom=(0, 0, 0), lookat=(0, 0, -1), defocus=angle, focus=focus)
    gpu_ctx.upload_scene(desc)
    gpu_ctx.render(cam, seed=5)
    img = gpu_ctx.download_radiance(cam.samples_per_pixel)
    px = 2.0 * math.tan(math.radians(vfov) / 2) * focus / W  # pixel size on the focus plane (camera.hpp:93-101)
    R = focus * math.tan(math.radians(angle) / 2)  # defocus radius (camera.hpp:128-130)
    # a lens point q sees the on-axis emitter at distance s through the focus-plane point q (1 - f/s): a uniform disk
    rc = R * (focus / s_dist - 1.0) / px
    re = r_e * focus / s_dist / px
    cx, cy, m2x, m2y, k = _moments(img)
    assert abs(cx - (W - 1) / 2) < 0.05 and abs(cy - (W - 1) / 2) < 0.05
    want = rc * rc / 4 + re * re / 4 + 1.0 / 12.0  # uniform disk (+ the emitter's own disk, + the pixel box)
    assert abs(m2x - want) / want < 0.02 and abs(m2y - want) / want < 0.02, (m2x, m2y, want)
    assert abs(k - 4.0 / 3.0) < 0.04, k  # <rho^4> / <rho^2>^2: 4/3 for a uniform disk (2 for a Gaussian, 1.5 for r = u sampling)
    # nothing outside the circle of confusion (+ emitter + a pixel)
    ys, xs = np.mgrid[0:W, 0:W]
    far = np.hypot(xs - cx, ys - cy) > rc + re + 1.5
    assert img[far].sum() == 0.0


def test_pixel_jitter_is_uniform_on_the_pixel_square(rtb, gpu_ctx):
    """A bright half-plane whose edge cuts through one pixel column / row at fraction f: that column's value is f of the
    full value only if the sample offsets are uniform on [-0.5, 0.5) (camera.hpp:165-177)."""
    W, vfov = 64, 40.0
    half = math.tan(math.radians(vfov) / 2) * 10.0  # half extent of the view at distance 10 (focus_dist)
    px = 2 * half / W
    for f in (0.25, 0.5, 0.8):
        for axis in (0, 1):
            edge = -half + (20 + f) * px  # the edge sits at fraction f inside column / row 20
            s = su.SceneDesc()
            light = s.light(s.solid(1, 1, 1))
            if axis == 0:  # lit for x < edge
                q = s.quad((-100, -100, -10), (100 + edge, 0, 0), (0, 200, 0), light)
            else:  # lit for y > -edge, i.e. rows above: image row j grows downwards
                q = s.quad((-100, -edge, -10), (200, 0, 0), (0, 100 + edge, 0), light)
            desc = s.finish(s.list([q]))
            cam = su.camera(width=W, spp=8192, depth=2, bg=(0, 0, 0), vfov=vfov, lookfrom=(0, 0, 0), lookat=(0, 0, -1))
            gpu_ctx.upload_scene(desc)
            gpu_ctx.render(cam, seed=2)
            img = gpu_ctx.download_radiance(cam.samples_per_pixel)[..., 0]
            line = img[32, :] if axis == 0 else img[:, 32]
            assert np.allclose(line[:20], 1.0) and np.allclose(line[21:], 0.0), (f, axis)
            assert abs(line[20] - f) < 0.02, (f, axis, line[20])  # 8192 samples: sigma = 0.005


def test_shutter_time_is_uniform(rtb, gpu_ctx):
    """A small emitter moving from x = -2 to x = +2 during the shutter (sphere.hpp moving sphere, ray time = random_double(),
    camera.hpp:161): its image is a streak of uniform brightness between the two end positions."""
    W, vfov = 200, 40.0
    s = su.SceneDesc()
    desc = s.finish(s.list([s.sphere((-2, 0, -10), 0.05, s.light(s.solid(20, 20, 20)), c2=(2, 0, -10))]))
    cam = su.camera(width=W, spp=16384, depth=2, bg=(0, 0, 0), vfov=vfov, lookfrom=(0, 0, 0), lookat=(0, 0, -1))
    gpu_ctx.upload_scene(desc)
    gpu_ctx.render(cam, seed=3)
    col = gpu_ctx.download_radiance(cam.samples_per_pixel).sum(axis=(0, 2)).astype(np.float64)  # energy per image column
    px = 2 * math.tan(math.radians(vfov) / 2) * 10.0 / W
    x0, x1 = (W - 1) / 2 - 2 / px, (W - 1) / 2 + 2 / px
    inside = col[int(x0) + 3:int(x1) - 2]
    assert col[:int(x0) - 2].sum() == 0 and col[int(x1) + 3:].sum() == 0
    assert inside.std() / inside.mean() < 0.05  # flat: per-column Monte-Carlo noise only (~1,200 hits per column: 2.9 %)
    q = len(inside) // 4
    quarters = [inside[i * q:(i + 1) * q].mean() for i in range(4)]
    assert max(quarters) / min(quarters) < 1.03, quarters
