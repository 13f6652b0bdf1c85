"""The C-ABI library loads, exports every symbol include/rt_b200.h declares, and the ctypes
mirror of its structs matches the compiler's layout.  No compute calls (runs without a GPU)."""
import ctypes as C
import os
import re

import pytest


def test_library_exports_every_declared_symbol(rtb, built):
    header = open(os.path.join(rtb.REPO_ROOT, "include", "rt_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*)\s+(rt_[a-z_0-9]+)\s*\(", header, re.M))
    assert declared == set(rtb._abi.C_ABI_SYMBOLS), declared ^ set(rtb._abi.C_ABI_SYMBOLS)
    lib = C.CDLL(rtb.CUDA_LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"librt_b200.so does not export {sym}"


def test_struct_layouts_match_the_compiler(rtb, built):
    lib = C.CDLL(rtb.CUDA_LIB_PATH)
    a = rtb._abi
    for which, st in enumerate([a.rt_hittable, a.rt_material, a.rt_texture, a.rt_image, a.rt_perlin, a.rt_scene_desc, a.rt_camera_desc,
                                a.rt_camera_frame, a.rt_render_opts, a.rt_stats]):
        assert lib.rt_abi_sizeof(which) == C.sizeof(st), st.__name__


def test_no_torch_types_in_the_boundary(rtb):
    header = open(os.path.join(rtb.REPO_ROOT, "include", "rt_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # strip comments
    assert "torch" not in code and "at::" not in code and "std::" not in code and "#include <" in code


def test_init_fails_loudly_without_a_device(rtb, built):
    """There is no CPU fallback: without a GPU rt_init returns RT_ERR_NO_DEVICE and says why."""
    lib = rtb.cuda_lib()
    h = C.c_void_p()
    rc = lib.rt_init(0, C.byref(h))
    if rc == rtb.RT_OK:  # a GPU is present (the -m gpu box): the other branch cannot be exercised
        lib.rt_shutdown(h)
        rc = lib.rt_init(4096, C.byref(h))
    assert rc == rtb.RT_ERR_NO_DEVICE
    assert lib.rt_last_error(None)
    with pytest.raises(rtb.RtError):
        rtb.Context(4096)


def test_camera_initialize_matches_the_oracle_bit_for_bit(rtb, orc, built):
    """camera::initialize (camera.hpp:76-136): product host arithmetic == oracle restatement."""
    for name in ["bouncing_spheres", "quads", "book2_final"]:
        sc = rtb.Scene(name)
        f1 = rtb.camera_frame(sc.cam.contents)
        f2 = orc.camera_frame(sc.cam.contents)
        assert bytes(f1) == bytes(f2), name
    # image_height = int(width / aspect) with the reference's float-literal aspect
    assert rtb.image_height(rtb.Scene("bouncing_spheres").cam.contents) == 224
    assert rtb.image_height(rtb.Scene("book1_final").cam.contents) == 674
    assert rtb.image_height(rtb.Scene("cornell_box").cam.contents) == 600
