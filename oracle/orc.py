"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never from the product package."""
import ctypes as C
import importlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_HARNESS = os.path.join(_HERE, "_ref", "ref_harness")
_lib = None


def abi():
    return importlib.import_module("raytracing-practice_b200._abi")


def lib():
    global _lib
    if _lib is None:
        a = abi()
        l = C.CDLL(LIB_PATH)
        dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        sp, cp, u64p = C.POINTER(a.rt_scene_desc), C.POINTER(a.rt_camera_desc), C.POINTER(C.c_uint64)
        l.orc_camera_initialize.argtypes = [cp, C.POINTER(a.rt_camera_frame)]
        l.orc_render_ppm.argtypes = [sp, cp, C.c_char_p, u64p]
        l.orc_render_linear.argtypes = [sp, cp, C.c_int, C.c_uint, C.c_int, dp, dp, u64p]
        l.orc_hit_rays.argtypes = [sp, C.c_int64, dp, dp, dp, C.c_double, C.c_double, C.c_int, ip, dp, dp, bp, dp, u64p]
        l.orc_primary.argtypes = [sp, cp, C.c_int, ip, dp, dp]
        l.orc_medium_spans.argtypes = [sp, C.c_int, C.c_int64, dp, dp, dp, dp, dp]
        l.orc_texture_value.argtypes = [sp, C.c_int, C.c_int64, dp, dp]
        l.orc_scatter.argtypes = [sp, C.c_int, C.c_int64, C.c_uint, dp, dp, bp, dp, dp, bp]
        l.orc_write_color.argtypes = [C.c_int64, dp, bp]
        l.orc_write_color.restype = None
        _lib = l
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def image_height(cam):
    return max(int(cam.image_width / cam.aspect_ratio), 1)


def camera_frame(cam):
    f = abi().rt_camera_frame()
    lib().orc_camera_initialize(C.byref(cam), C.byref(f))
    return f


def render_ppm(desc, cam, path):
    rays = C.c_uint64()
    rc = lib().orc_render_ppm(desc, C.byref(cam), path.encode(), C.byref(rays))
    assert rc == 0, rc
    return rays.value


def render_linear(desc, cam, spp, seed=1, threads=None, want_sq=True, rng="xoshiro"):
    """Returns (mean, var_of_mean, rays): linear fp64 radiance statistics per pixel/channel.

    rng="glibc" replays the reference's own generator (rand(), a lagged-Fibonacci r[i]=r[i-3]+r[i-31]);
    rng="xoshiro" (default for statistical gates) drives the SAME algorithm with a high-quality
    generator.  Measured here (Cornell room, 2 x 26 M samples): the glibc-driven estimate is
    0.15 % darker, z = -3.4, because rand()'s 3-point correlation meets the 3-draws-per-try
    rejection sampler; see DESIGN.md "Reference RNG artefact"."""
    threads = threads or os.cpu_count() or 1
    seed = (seed & 0x7FFFFFFF) | (0x80000000 if rng == "xoshiro" else 0)
    h, w = image_height(cam), cam.image_width
    s = np.zeros((h, w, 3))
    q = np.zeros((h, w, 3)) if want_sq else None
    rays = C.c_uint64()
    rc = lib().orc_render_linear(desc, C.byref(cam), spp, seed, threads, _d(s), _d(q), C.byref(rays))
    assert rc == 0, rc
    mean = s / spp
    var = None
    if want_sq and spp > 1:
        var = np.maximum(q / spp - mean * mean, 0.0) / (spp - 1)
    return mean, var, rays.value


def hit_rays(desc, origin, direction, time=None, tmin=0.001, tmax=float("inf"), skip_media=False, census=False):
    origin = np.ascontiguousarray(origin, np.float64)
    direction = np.ascontiguousarray(direction, np.float64)
    n = origin.shape[0]
    time = np.zeros(n) if time is None else np.ascontiguousarray(time, np.float64)
    ids = np.empty(n, np.int32)
    t = np.empty(n)
    nrm = np.empty((n, 3))
    ff = np.empty(n, np.uint8)
    uv = np.empty((n, 2))
    cen = np.zeros(4, np.uint64)
    rc = lib().orc_hit_rays(desc, n, _d(origin), _d(direction), _d(time), tmin, tmax, int(skip_media),
                            ids.ctypes.data_as(C.POINTER(C.c_int32)), _d(t), _d(nrm), ff.ctypes.data_as(C.POINTER(C.c_uint8)), _d(uv),
                            cen.ctypes.data_as(C.POINTER(C.c_uint64)) if census else None)
    assert rc == 0, rc
    return (ids, t, nrm, ff, uv, cen) if census else (ids, t, nrm, ff, uv)


def primary(desc, cam, skip_media=True):
    h, w = image_height(cam), cam.image_width
    ids = np.empty((h, w), np.int32)
    t = np.empty((h, w))
    nrm = np.empty((h, w, 3))
    rc = lib().orc_primary(desc, C.byref(cam), int(skip_media), ids.ctypes.data_as(C.POINTER(C.c_int32)), _d(t), _d(nrm))
    assert rc == 0, rc
    return ids, t, nrm


def medium_spans(desc, medium_index, origin, direction, time=None):
    origin = np.ascontiguousarray(origin, np.float64)
    direction = np.ascontiguousarray(direction, np.float64)
    n = origin.shape[0]
    time = np.zeros(n) if time is None else np.ascontiguousarray(time, np.float64)
    t1, t2 = np.empty(n), np.empty(n)
    rc = lib().orc_medium_spans(desc, medium_index, n, _d(origin), _d(direction), _d(time), _d(t1), _d(t2))
    assert rc == 0, rc
    return t1, t2


def texture_value(desc, texture, uvp):
    uvp = np.ascontiguousarray(uvp, np.float64)
    out = np.empty((uvp.shape[0], 3))
    rc = lib().orc_texture_value(desc, texture, uvp.shape[0], _d(uvp), _d(out))
    assert rc == 0, rc
    return out


def scatter(desc, material, dir_in, normal, front_face, seed=1):
    dir_in = np.ascontiguousarray(dir_in, np.float64)
    normal = np.ascontiguousarray(normal, np.float64)
    front_face = np.ascontiguousarray(front_face, np.uint8)
    n = dir_in.shape[0]
    d, a, s = np.empty((n, 3)), np.empty((n, 3)), np.empty(n, np.uint8)
    bp = C.POINTER(C.c_uint8)
    rc = lib().orc_scatter(desc, material, n, seed, _d(dir_in), _d(normal), front_face.ctypes.data_as(bp), _d(d), _d(a), s.ctypes.data_as(bp))
    assert rc == 0, rc
    return d, a, s


def write_color(linear_rgb):
    """The reference's write_color transform (gamma, clamp, int(256x)) -> uint8 image."""
    lin = np.ascontiguousarray(linear_rgb, np.float64)
    out = np.empty(lin.shape, np.uint8)
    lib().orc_write_color(lin.size // 3, _d(lin), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def fnv1a64_ids(ids):
    """FNV-1a 64 over little-endian int32 ids."""
    h = 0xCBF29CE484222325
    for b in np.ascontiguousarray(ids, "<i4").tobytes():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"
