"""T4/T5/T6 — the rendered image: converged-image parity against the CPU oracle (RNG streams
differ, so parity is statistical with a variance-derived tolerance), sharding invariance
(bit-exact), determinism, the write_color tail and the C-ABI's error behaviour."""
import ctypes as C

import numpy as np
import pytest

import scene_util as su
from conftest import ALL_SCENES

pytestmark = pytest.mark.gpu

GPU_SPP, CPU_SPP = 4096, 256


@pytest.mark.parametrize("name", ALL_SCENES)
def test_converged_image_matches_oracle(rtb, orc, gpu_ctx, name):
    """Gate 1 (north_star): per-channel RMSE after write_color's gamma <= tolerance, where the
    tolerance is 1.3 x the RMSE the two estimators' own Monte-Carlo noise predicts.
    Gate 2 (bias): channel means within 5 standard errors.  Oracle = the reference algorithm
    in fp64 driven by a good generator (see orc.render_linear about glibc rand())."""
    sc = rtb.Scene(name, rand_seed=1)
    width = 96 if sc.cam.contents.aspect_ratio > 1.2 else 72
    cam = sc.camera_copy(image_width=width, samples_per_pixel=GPU_SPP)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=17)
    img = gpu_ctx.download_radiance(GPU_SPP).astype(np.float64)
    st = gpu_ctx.stats()
    mean, var, orays = orc.render_linear(sc.desc, cam, spp=CPU_SPP, seed=23)
    var_tot = var * (1.0 + CPU_SPP / GPU_SPP)
    # gate 2: bias
    for c in range(3):
        z = (img[..., c].sum() - mean[..., c].sum()) / np.sqrt(var_tot[..., c].sum() + 1e-30)
        assert abs(z) < 5.0, f"channel {c}: z = {z:.2f}"
    # gate 1: RMSE after gamma, tolerance propagated through d sqrt(x)/dx = 1 / (2 sqrt(x))
    g_gpu = np.sqrt(np.clip(img, 0, 0.999 ** 2))
    g_cpu = np.sqrt(np.clip(mean, 0, 0.999 ** 2))
    rmse = np.sqrt(np.mean((g_gpu - g_cpu) ** 2, axis=(0, 1)))
    lin = np.maximum(0.5 * (img + mean), 1e-4)
    predicted = np.sqrt(np.mean(var_tot / (4.0 * lin), axis=(0, 1)))
    assert np.all(rmse <= 1.3 * predicted + 2e-3), (rmse, predicted)
    # same path-length statistics (rays per sample)
    rps_gpu, rps_cpu = st.rays / st.samples, orays / (mean.shape[0] * mean.shape[1] * CPU_SPP)
    assert abs(rps_gpu - rps_cpu) / rps_cpu < 0.01
    assert st.samples == mean.shape[0] * mean.shape[1] * GPU_SPP


def test_image_is_independent_of_sample_sharding(rtb, gpu_ctx):
    """T5: Philox keyed on (pixel, sample, bounce) + int64 fixed-point sums => the accumulator is
    BIT-identical whether 64 spp are rendered in one launch, in 4 launches of 16, or as 8
    uneven 'rank' shards — i.e. independent of GPU count."""
    sc = rtb.Scene("cornell_smoke", rand_seed=1)
    cam = sc.camera_copy(image_width=80, samples_per_pixel=64, max_depth=12)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=3)
    whole = gpu_ctx.download_accum()
    for shards in ([16, 16, 16, 16], [1, 7, 8, 13, 3, 20, 5, 7]):
        begin = 0
        for i, n in enumerate(shards):
            gpu_ctx.render(cam, seed=3, sample_begin=begin, sample_count=n, clear=(i == 0))
            begin += n
        assert begin == 64
        assert np.array_equal(gpu_ctx.download_accum(), whole)
    gpu_ctx.render(cam, seed=4)
    assert not np.array_equal(gpu_ctx.download_accum(), whole)  # the seed matters
    gpu_ctx.render(cam, seed=3)
    assert np.array_equal(gpu_ctx.download_accum(), whole)  # and the render is deterministic


def test_peer_accumulation_two_contexts(rtb, gpu_ctx):
    """Two contexts ('ranks') render disjoint sample shards; rank 1 adds straight into rank 0's
    accumulator (rt_render_opts.peer_accum, the fused reduce).  Same bits as one context."""
    sc = rtb.Scene("simple_light", rand_seed=1)
    cam = sc.camera_copy(image_width=96, samples_per_pixel=48)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=8)
    whole = gpu_ctx.download_accum()
    other = rtb.Context(0)
    other.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=8, sample_begin=0, sample_count=20)
    gpu_ctx.synchronize()
    ptr, nbytes = gpu_ctx.accum_ptr()
    assert nbytes == whole.nbytes
    other.render(cam, seed=8, sample_begin=20, sample_count=28, peer_accum=ptr)
    other.synchronize()
    assert np.array_equal(gpu_ctx.download_accum(), whole)
    # the target must be a live context's accumulator of this image size (rt_b200.h): anything else is refused, not written to
    with pytest.raises(rtb.RtError):
        other.render(cam, seed=8, peer_accum=ptr + 64)
    small = sc.camera_copy(image_width=48, samples_per_pixel=4)
    with pytest.raises(rtb.RtError):
        other.render(small, seed=8, peer_accum=ptr)
    assert np.array_equal(gpu_ctx.download_accum(), whole)
    other.close()


def test_write_color_tail_and_ppm(rtb, orc, gpu_ctx, tmp_path):
    """RT_BUF_RGB8 is write_color (color.hpp:26-58) applied to the radiance; P3 text format."""
    sc = rtb.Scene("quads", rand_seed=1)
    cam = sc.camera_copy(image_width=50, samples_per_pixel=30)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=1)
    acc = gpu_ctx.download_accum()
    rgb = gpu_ctx.download_rgb8(30)
    lin = acc.astype(np.float64) / 2.0 ** 32 * np.float64(np.float32(1.0) / np.float32(30))
    assert np.array_equal(rgb, orc.write_color(lin))
    rad = gpu_ctx.download_radiance(30)
    assert np.allclose(rad, lin, rtol=1e-6)
    path = tmp_path / "o.ppm"
    rtb.write_ppm_p3(str(path), rgb)
    lines = open(path).read().split("\n")
    assert lines[0] == "P3" and lines[1] == "50 50" and lines[2] == "255" and len(lines) == 3 + 2500 + 1
    assert all(0 <= int(x) <= 255 for x in lines[3].split())


def test_background_only_and_zero_depth(rtb, gpu_ctx):
    s = su.SceneDesc()
    desc = s.finish(s.list([]))
    gpu_ctx.upload_scene(desc)
    cam = su.camera(width=33, aspect=16 / 9, spp=5, depth=10, bg=(0.25, 0.5, 1.0))
    gpu_ctx.render(cam)
    acc = gpu_ctx.download_accum()
    assert acc.shape == (18, 33, 3)
    assert np.all(acc == (np.array([0.25, 0.5, 1.0]) * 5 * 2 ** 32).astype(np.int64))
    st = gpu_ctx.stats()
    assert st.rays == 33 * 18 * 5 and st.samples == 33 * 18 * 5
    cam.max_depth = 0  # ray_color returns black immediately (camera.hpp:183-186)
    gpu_ctx.render(cam)
    assert not gpu_ctx.download_accum().any()
    assert gpu_ctx.stats().rays == 0


def test_full_size_headline_scene_properties(rtb, gpu_ctx):
    """BASELINE config 5 at its full 800x800 size (reduced spp): shard additivity (a checksum of
    checksums) and sample accounting — size-independent properties, the oracle is too slow here."""
    sc = rtb.Scene("book2_final", rand_seed=1)
    cam = sc.camera_copy(samples_per_pixel=32)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=0)
    whole = gpu_ctx.download_accum()
    st = gpu_ctx.stats()
    assert whole.shape == (800, 800, 3) and st.samples == 800 * 800 * 32
    assert 3.5 < st.rays / st.samples < 6.5
    total = np.zeros_like(whole)
    for r in range(8):  # 8 'ranks', each in its own cleared accumulator, summed on the host
        gpu_ctx.render(cam, seed=0, sample_begin=4 * r, sample_count=4, clear=True)
        total += gpu_ctx.download_accum()
    assert np.array_equal(total, whole)
    assert whole.min() >= 0


def test_error_behaviour(rtb, built):
    lib = rtb.cuda_lib()
    ctx = rtb.Context(0)
    cam = su.camera()
    with pytest.raises(rtb.RtError, match="rt_upload_scene"):
        ctx.render(cam)
    s = su.SceneDesc()
    k = s.sphere((0, 0, 0), 1, s.lambertian(s.solid(1, 1, 1)))
    s.h[k].kind = 99  # a hittable subclass the flattener does not know
    desc = s.finish(k)
    assert lib.rt_upload_scene(ctx._h, desc) == rtb.RT_ERR_UNSUPPORTED
    assert b"unknown hittable" in lib.rt_last_error(ctx._h)
    s = su.SceneDesc()
    desc = s.finish(s.sphere((0, 0, 0), 1, 5))  # material index out of range
    assert lib.rt_upload_scene(ctx._h, desc) == rtb.RT_ERR_INVALID
    s = su.SceneDesc()
    desc = s.finish(s.sphere((0, 0, 0), 1, s.lambertian(s.solid(1, 1, 1))))
    ctx.upload_scene(desc)
    bad = su.camera(width=0)
    with pytest.raises(rtb.RtError):
        ctx.render(bad)
    buf = np.zeros(4, np.uint8)
    ctx.render(su.camera(width=8, spp=1))
    assert lib.rt_download(ctx._h, rtb.RT_BUF_RGB8, 1, buf.ctypes.data_as(C.c_void_p), buf.nbytes) == rtb.RT_ERR_INVALID
    ctx.close()


def test_cpp_drop_in_camera_render(rtb, gpu_ctx, tmp_path):
    """The C++ host path a reference user takes: scene classes -> camera::render(ostream, world)
    -> flatten -> rt_upload_scene / rt_render / rt_download -> P3 text.  Must equal the image
    the Python binding produces from the same scene (same seed, same kernel)."""
    import os
    import subprocess

    host = os.path.join(rtb.REPO_ROOT, "raytracing-practice_b200", "host")
    exe = str(tmp_path / "dropin")
    libdir = os.path.dirname(rtb.CUDA_LIB_PATH)
    cmd = ["g++", "-std=c++11", "-O1", "-I", host, "-I", os.path.join(rtb.REPO_ROOT, "include"),
           os.path.join(rtb.REPO_ROOT, "tests", "cpp", "dropin_main.cpp"), "-L", libdir, "-lrt_b200", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    for devices in ("0", "0,0,0"):  # one context, then three 'ranks' on the same device summed on the host
        out = str(tmp_path / f"img_{len(devices)}.ppm")
        env = dict(os.environ, RT_B200_DEVICES=devices)
        r = subprocess.run([exe, "cornell_rotated", out, "90", "40"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "Scanlines remaining" in r.stdout and "Done." in r.stdout  # camera.hpp:47,70
        tok = open(out).read().split()
        assert tok[:4] == ["P3", "90", "90", "255"]
        img = np.array(tok[4:], dtype=np.int64).reshape(90, 90, 3)
        sc = rtb.Scene("cornell_rotated", rand_seed=1)
        cam = sc.camera_copy(image_width=90, samples_per_pixel=40)
        gpu_ctx.upload_scene(sc.desc)
        gpu_ctx.render(cam, seed=0)
        assert np.array_equal(img, gpu_ctx.download_rgb8(40).astype(np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["stream", "refill"])
@pytest.mark.parametrize("name,width,spp", [("book2_final", 160, 48), ("cornell_smoke", 96, 40), ("book1_final", 200, 24), ("perlin_sphere", 120, 16)])
def test_scheduling_variants_are_bit_identical_to_the_megakernel(rtb, gpu_ctx, kernel, name, width, spp):
    """The streaming kernel (csrc/rt_stream.cuh: CTA-wide ray queues) and the in-place-refill kernel (csrc/rt_refill.cuh:
    the warp leaves the traversal to shade as soon as enough lanes are finished) only change WHEN a ray is traced and
    shaded, not what is computed for it: same accumulator bits and ray count as the megakernel, also on a sample shard."""
    flag = {"stream": rtb.RT_RENDER_STREAM, "refill": rtb.RT_RENDER_REFILL}[kernel]
    sc = rtb.Scene(name, rand_seed=1)
    cam = sc.camera_copy(image_width=width, samples_per_pixel=spp)
    gpu_ctx.upload_scene(sc.desc)
    out = {}
    for tag, flags in (("mega", rtb.RT_RENDER_MEGAKERNEL), ("alt", flag)):
        gpu_ctx.render(cam, seed=13, flags=flags)
        out[tag] = (gpu_ctx.download_accum(), gpu_ctx.stats().rays)
        gpu_ctx.render(cam, seed=13, flags=flags, sample_begin=3, sample_count=9)
        out[tag + "_shard"] = (gpu_ctx.download_accum(), gpu_ctx.stats().rays)
    assert out["mega"][1] == out["alt"][1] and np.array_equal(out["mega"][0], out["alt"][0])
    assert out["mega_shard"][1] == out["alt_shard"][1] and np.array_equal(out["mega_shard"][0], out["alt_shard"][0])
    cam.max_depth = 0  # ray_color returns black at once (camera.hpp:183-186): nothing is launched
    gpu_ctx.render(cam, seed=13, flags=flag)
    assert not gpu_ctx.download_accum().any() and gpu_ctx.stats().rays == 0
    with pytest.raises(rtb.RtError):  # the bit of the removed path-pool kernel
        gpu_ctx.render(cam, seed=13, flags=4)


def test_box_primitive_equals_its_six_quads(rtb, gpu_ctx, monkeypatch):
    """box() lists (quad.hpp:129-159) are flattened to ONE slab-test primitive.  With RT_B200_NO_BOXES the
    same lists stay six quads: the exact primary pass must agree on every pixel (ids, t, normal — the exact
    predicate always evaluates the six reference quads), the fp32 production traversal on all but grazing
    pixels, and the rendered means within Monte-Carlo noise."""
    res = {}
    for tag in ("boxes", "quads"):
        if tag == "quads":
            monkeypatch.setenv("RT_B200_NO_BOXES", "1")
        for name in ("cornell_rotated", "book2_final"):
            sc = rtb.Scene(name, rand_seed=1)
            cam = sc.camera_copy(image_width=128, samples_per_pixel=256)
            gpu_ctx.upload_scene(sc.desc)
            st0 = gpu_ctx.stats()
            exact = gpu_ctx.primary_visibility(cam)
            fp32 = gpu_ctx.primary_visibility(cam, flags=rtb.RT_TRACE_FP32 | rtb.RT_TRACE_SKIP_MEDIA)
            gpu_ctx.render(cam, seed=5)
            res[tag, name] = (exact, fp32, gpu_ctx.download_radiance(256).astype(np.float64), st0.n_boxes, gpu_ctx.stats().rays)
    monkeypatch.delenv("RT_B200_NO_BOXES")
    for name, n_boxes in (("cornell_rotated", 2), ("book2_final", 400)):
        b, q = res["boxes", name], res["quads", name]
        assert b[3] == n_boxes and q[3] == 0
        assert np.array_equal(b[0][0], q[0][0]) and np.array_equal(b[0][1], q[0][1]) and np.array_equal(b[0][2], q[0][2])
        hit = b[0][0] >= 0
        assert (b[1][0] != b[0][0]).mean() <= 1e-3  # fp32 vs exact ids
        same = hit & (b[1][0] == b[0][0])
        assert np.allclose(b[1][1][same], b[0][1][same], rtol=2e-4) and np.allclose(b[1][2][same], b[0][2][same], atol=2e-4)
        # two 256-spp estimates of the same image with DIFFERENT fp32 roundings at the hit points
        assert abs(b[2].mean() - q[2].mean()) < 0.02 * q[2].mean() + 1e-3
        assert abs(b[4] - q[4]) / q[4] < 0.01


def test_peer_reduce_push_accum(rtb, gpu_ctx):
    """The multi-GPU exchange step without a collective (rt_render_opts.push_accum): three 'ranks' (contexts) render
    disjoint sample shards, behind each render a push kernel adds the rank's accumulator into rank 0's reduce buffer;
    the adopted image has the bits of a single render.  Two of the render kernels; the buffer is reusable."""
    sc = rtb.Scene("cornell_smoke", rand_seed=1)
    cam = sc.camera_copy(image_width=120, samples_per_pixel=36, max_depth=10)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=4)
    whole, rays = gpu_ctx.download_accum(), gpu_ctx.stats().rays
    others = [rtb.Context(0), rtb.Context(0)]
    for o in others:
        o.upload_scene(sc.desc)
    for flags in (rtb.RT_RENDER_MEGAKERNEL, rtb.RT_RENDER_REFILL):
        ptr, handle = gpu_ctx.reduce_buffer(cam)
        assert len(handle) == 64 and any(handle)
        for ctx, (b, n) in zip([gpu_ctx] + others, [(0, 10), (10, 7), (17, 19)]):
            ctx.render(cam, seed=4, sample_begin=b, sample_count=n, push_accum=ptr, flags=flags)
        for ctx in [gpu_ctx] + others:
            ctx.synchronize()
        assert not np.array_equal(gpu_ctx.download_accum(), whole)  # rank 0's own accumulator holds only its shard
        gpu_ctx.adopt_reduce_buffer()
        assert np.array_equal(gpu_ctx.download_accum(), whole)
        assert sum(c.stats().rays for c in [gpu_ctx] + others) == rays
    for o in others:
        o.close()


def test_oversized_render_is_split_into_launches(rtb, gpu_ctx, monkeypatch):
    """A request with more than 2^32 work items is rendered as several launches over consecutive sample ranges
    (rt_render); RT_B200_MAX_CHUNKS lowers the per-launch limit so the split can be exercised at test size.  Same
    accumulator bits, ray and sample counts, and the peer push happens once, after the last piece."""
    sc = rtb.Scene("bouncing_spheres", rand_seed=1)
    cam = sc.camera_copy(image_width=120, samples_per_pixel=150)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=6)
    whole, st = gpu_ctx.download_accum(), gpu_ctx.stats()
    monkeypatch.setenv("RT_B200_MAX_CHUNKS", "2")  # 64 samples per launch -> 3 launches; knobs are read when a context is created
    small = rtb.Context(0)
    small.upload_scene(sc.desc)
    launches0 = small.stats().kernel_launches
    small.render(cam, seed=6)
    st2 = small.stats()
    assert np.array_equal(small.download_accum(), whole) and st2.rays == st.rays and st2.samples == st.samples
    assert st2.kernel_launches - launches0 == 3
    ptr, _ = small.reduce_buffer(cam)
    small.render(cam, seed=6, push_accum=ptr, flags=rtb.RT_RENDER_REFILL)
    small.adopt_reduce_buffer()
    assert np.array_equal(small.download_accum(), whole)
    small.close()
    launches1 = gpu_ctx.stats().kernel_launches  # the session's context was created without the knob: one launch
    gpu_ctx.render(cam, seed=6)
    assert gpu_ctx.stats().kernel_launches - launches1 == 1


@pytest.mark.parametrize("width,aspect,spp", [(1, 1.0, 1), (7, 7 / 5, 3), (33, 16 / 9, 1), (257, 4.0, 2)])
def test_ragged_image_sizes_and_tiny_jobs(rtb, gpu_ctx, width, aspect, spp):
    """Images that are not a multiple of the 8x4 work tile, a single pixel, a single sample: every (pixel, sample)
    is rendered exactly once by either kernel (sample accounting, bit-identical accumulators), rows x columns follow
    camera::initialize (int(W / aspect) clamped to >= 1, camera.hpp:79-80)."""
    sc = rtb.Scene("cornell_box", rand_seed=1)
    cam = sc.camera_copy(image_width=width, samples_per_pixel=spp, max_depth=6)
    cam.aspect_ratio = aspect
    gpu_ctx.upload_scene(sc.desc)
    h = max(1, int(width / aspect))
    gpu_ctx.render(cam, seed=2, flags=rtb.RT_RENDER_MEGAKERNEL)
    a, st = gpu_ctx.download_accum(), gpu_ctx.stats()
    assert a.shape == (h, width, 3) and st.samples == width * h * spp and st.rays >= st.samples
    gpu_ctx.render(cam, seed=2, flags=rtb.RT_RENDER_REFILL)
    assert np.array_equal(gpu_ctx.download_accum(), a) and gpu_ctx.stats().rays == st.rays
    # one sample at a time, accumulated: same bits
    for s in range(spp):
        gpu_ctx.render(cam, seed=2, sample_begin=s, sample_count=1, clear=(s == 0))
    assert np.array_equal(gpu_ctx.download_accum(), a)
    assert gpu_ctx.download_rgb8(spp).shape == (h, width, 3)


def test_render_option_errors(rtb, gpu_ctx):
    sc = rtb.Scene("quads", rand_seed=1)
    cam = sc.camera_copy(image_width=32, samples_per_pixel=4)
    gpu_ctx.upload_scene(sc.desc)
    ptr, _ = gpu_ctx.reduce_buffer(cam)
    with pytest.raises(rtb.RtError, match="mutually exclusive"):
        gpu_ctx.render(cam, peer_accum=ptr, push_accum=ptr)
    with pytest.raises(rtb.RtError, match="empty sample range"):
        gpu_ctx.render(cam, sample_begin=4)
    with pytest.raises(rtb.RtError, match="empty sample range"):
        gpu_ctx.render(cam, sample_begin=-1, sample_count=2)
    big = sc.camera_copy(image_width=64, samples_per_pixel=4)
    gpu_ctx.render(big)
    with pytest.raises(rtb.RtError, match="differ in size"):
        gpu_ctx.adopt_reduce_buffer()  # the reduce buffer was made for the 32-wide camera
    gpu_ctx.render(cam, seed=1)  # still usable afterwards
    assert gpu_ctx.stats().samples == 32 * 32 * 4


def test_upload_accum_resumes_a_render(rtb, gpu_ctx):
    """rt_upload_accum is the inverse of rt_download(ACCUM_I64): sums saved after samples [0, 20), restored into a
    fresh accumulator and continued with [20, 40) give the bits of the uninterrupted 40-sample render
    (camera.hpp:55-61 is a plain sum; the device keeps it in integers)."""
    sc = rtb.Scene("cornell_smoke", rand_seed=1)
    cam = sc.camera_copy(image_width=72, samples_per_pixel=40)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=5)
    whole = gpu_ctx.download_accum().copy()
    rgb = gpu_ctx.download_rgb8(40).copy()
    gpu_ctx.render(cam, seed=5, sample_begin=0, sample_count=20)
    saved = gpu_ctx.download_accum().copy()
    gpu_ctx.render(cam, seed=9, sample_begin=0, sample_count=3)  # the accumulator now holds something else
    gpu_ctx.upload_accum(cam, saved)
    assert np.array_equal(gpu_ctx.download_accum(), saved)
    gpu_ctx.render(cam, seed=5, sample_begin=20, sample_count=20, clear=False)
    assert np.array_equal(gpu_ctx.download_accum(), whole)
    assert np.array_equal(gpu_ctx.download_rgb8(40), rgb)
    with pytest.raises(rtb.RtError):  # wrong size
        gpu_ctx.upload_accum(cam, saved[:-1])
    other = rtb.Context(0)  # a context that never rendered can be restored into and finalised
    other.upload_accum(cam, whole)
    assert np.array_equal(other.download_rgb8(40), rgb)
    other.close()


def test_cpp_checkpointed_render_and_p6(rtb, gpu_ctx, tmp_path):
    """The C++ host's progressive mode (RT_B200_CHECKPOINT*): passes of 7 spp, an interruption after 14, a resume on
    another number of 'devices' — the P3 text equals the uninterrupted render's, and RT_B200_P6 writes the same bytes
    in binary.  A checkpoint of another camera is refused."""
    import os
    import subprocess

    host = os.path.join(rtb.REPO_ROOT, "raytracing-practice_b200", "host")
    exe = str(tmp_path / "dropin")
    libdir = os.path.dirname(rtb.CUDA_LIB_PATH)
    cmd = ["g++", "-std=c++11", "-O1", "-I", host, "-I", os.path.join(rtb.REPO_ROOT, "include"),
           os.path.join(rtb.REPO_ROOT, "tests", "cpp", "dropin_main.cpp"), "-L", libdir, "-lrt_b200", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    base = {k: v for k, v in os.environ.items() if not k.startswith("RT_B200_")}

    def run(out, env, spp="33"):
        return subprocess.run([exe, "simple_light", out, "80", spp], capture_output=True, text=True, env=dict(base, **env))

    one, p6 = str(tmp_path / "one.ppm"), str(tmp_path / "one.p6")
    r = run(one, {"RT_B200_P6": p6})
    assert r.returncode == 0, r.stderr[-2000:]
    tok = open(one).read().split()
    w, h = int(tok[1]), int(tok[2])
    raw = open(p6, "rb").read()
    head = b"P6\n%d %d\n255\n" % (w, h)
    assert raw.startswith(head) and len(raw) == len(head) + w * h * 3
    assert np.array_equal(np.frombuffer(raw[len(head):], np.uint8), np.array(tok[4:], dtype=np.uint8))

    ck, part, two = str(tmp_path / "render.ckpt"), str(tmp_path / "part.ppm"), str(tmp_path / "two.ppm")
    r = run(part, {"RT_B200_CHECKPOINT": ck, "RT_B200_CHECKPOINT_SPP": "7", "RT_B200_STOP_AFTER_SPP": "14"})
    assert r.returncode == 3 and "Stopped after 14 of 33" in r.stdout, (r.returncode, r.stdout, r.stderr[-2000:])
    assert os.path.getsize(ck) == 48 + w * h * 3 * 8
    r = run(two, {"RT_B200_CHECKPOINT": ck, "RT_B200_CHECKPOINT_SPP": "7"}, spp="32")  # another sample count: refused
    assert r.returncode == 1 and "another scene, camera or sample count" in r.stderr
    r = run(two, {"RT_B200_CHECKPOINT": ck, "RT_B200_CHECKPOINT_SPP": "8", "RT_B200_DEVICES": "0,0,0"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(two).read() == open(one).read()
    r = run(two, {"RT_B200_CHECKPOINT": ck})  # a finished checkpoint: nothing left to render, same image
    assert r.returncode == 0 and open(two).read() == open(one).read()


def test_very_bright_emitter_saturates_instead_of_wrapping(rtb, gpu_ctx):
    """to_fixed clamps one sample at 2^16 before the 2^32 fixed-point scale (rt_b200.cu): an emitter of 1e5 seen directly
    at 512 spp sums to 512 * 2^48 = 2^57 — inside the signed 64-bit range — instead of wrapping into garbage or negative
    pixels (with the old 1e6 clamp 2,100 such samples overflowed).  write_color clips the mean at 0.999 either way: 255."""
    import scene_util as su

    s = su.SceneDesc()
    sun = s.sphere((0, 0, 0), 1.0, s.light(s.solid(1e5, 2e5, 5e4)))
    desc = s.finish(s.list([sun]))
    cam = su.camera(width=48, spp=512, depth=4, bg=(0, 0, 0), lookfrom=(0, 0, 6), vfov=30.0)
    gpu_ctx.upload_scene(desc)
    gpu_ctx.render(cam, seed=1)
    acc = gpu_ctx.download_accum()
    assert (acc >= 0).all()
    centre = acc[24, 24]
    assert np.array_equal(centre, np.array([512 * 65536 << 32] * 2 + [512 * 50000 << 32]))  # r, g at the clamp; b = 5e4 exactly
    assert np.array_equal(gpu_ctx.download_rgb8(512)[24, 24], [255, 255, 255])
    assert not acc[0, 0].any()  # the black background
