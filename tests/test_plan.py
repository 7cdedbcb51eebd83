"""CPU suite: argument checking of the plan-level C entry (sdk_plan_*, include/sdb200.h) -- no device work."""
import ctypes as C

import pytest

from stable_diffusion_pytorch_b200 import _lib


@pytest.fixture()
def plan():
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.sdk_plan_create(C.byref(h)) == 0
    yield lib, h
    assert lib.sdk_plan_destroy(h) == 0


def _args(*v):
    return (C.c_uint64 * max(len(v), 1))(*v)


def test_add_launch_checks_name_and_arity(plan):
    lib, h = plan
    assert lib.sdk_plan_num_launches(h, 0) == 0
    # sdk_layernorm(x, gamma, beta, eps, out, out_dtype, rows, C, stream): 8 arguments before the stream
    assert lib.sdk_plan_add_launch(h, 0, b"sdk_layernorm", _args(1, 2, 3, 0x3727C5AC, 4, 1, 128, 320), 8) == 0
    assert lib.sdk_plan_num_launches(h, 0) == 1
    assert lib.sdk_plan_add_launch(h, 0, b"sdk_layernorm", _args(1, 2, 3), 3) == -1
    assert b"8 arguments" in lib.sdk_last_error()
    assert lib.sdk_plan_add_launch(h, 0, b"sdk_tc_gemm_create", _args(1, 2), 2) == -3          # not a launch-type entry point
    assert lib.sdk_plan_add_launch(h, 99, b"sdk_layernorm", _args(1, 2, 3, 0, 4, 1, 128, 320), 8) == -1
    # a handle launch needs an adopted handle
    assert lib.sdk_plan_add_launch(h, 0, b"sdk_tc_gemm_launch", _args(0xdead0), 1) == -1
    assert b"adopt" in lib.sdk_last_error()
    assert lib.sdk_plan_num_launches(h, 0) == 1 and lib.sdk_plan_num_launches(h, 1) == 0 and lib.sdk_plan_num_launches(h, 99) == -1


def test_every_recorded_entry_point_is_in_the_binding_table():
    """The Python transcription looks the argument types up in _lib.SIGNATURES: every launch-type entry point must be there."""
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.sdk_plan_create(C.byref(h)) == 0
    launchers = [n for n, sig in _lib.SIGNATURES.items()
                 if sig and sig[-1] is _lib.P and not n.startswith(("sdk_plan_", "sdk_stream_")) and "create" not in n
                 and "destroy" not in n and n not in ("sdk_tc_gemm_set_workspace", "sdk_tc_gemm_info", "sdk_tc_gemm_set_debug", "sdk_tc_gemm_set_stats",
                                                      "sdk_tc_gemm_workspace_bytes", "sdk_linear_ln_info", "sdk_device_info")]
    known = 0
    for n in launchers:
        nargs = len(_lib.SIGNATURES[n]) - 1
        rc = lib.sdk_plan_add_launch(h, 1, n.encode(), _args(*([0] * nargs)), nargs)
        if n in ("sdk_tc_gemm_launch", "sdk_attention_tc_launch", "sdk_linear_ln_launch", "sdk_conv_gemm_f32"):
            assert rc == -1, n                                     # known, but needs a handle / struct
        else:
            assert rc == 0, (n, lib.sdk_last_error())
            known += 1
    assert known >= 25
    lib.sdk_plan_destroy(h)


def test_regions_and_adopt_checks(plan):
    lib, h = plan
    assert lib.sdk_plan_add_region(h, 0x10000, 4096, 0, b"") == 0
    assert lib.sdk_plan_add_region(h, 0x10000, 4096, 0, b"again") == 0               # same buffer twice: harmless
    assert lib.sdk_plan_add_region(h, 0x10800, 4096, 1, b"") == -1                   # partial overlap
    assert lib.sdk_plan_add_region(h, 0x20000, 0, 1, b"") == -1
    p, n = C.c_void_p(), C.c_int64()
    assert lib.sdk_plan_region(h, b"again", C.byref(p), C.byref(n)) == 0 and p.value == 0x10000 and n.value == 4096
    assert lib.sdk_plan_region(h, b"nope", C.byref(p), C.byref(n)) == -1
    d = _lib.LinearLnDesc()
    assert lib.sdk_plan_adopt(h, 3, 0, C.byref(d), C.sizeof(d), _args(0, 0), 2) == -1         # null handle
    assert lib.sdk_plan_adopt(h, 7, 0x1234, C.byref(d), C.sizeof(d), _args(0, 0), 2) == -1    # unknown kind
    assert lib.sdk_plan_adopt(h, 1, 0x1234, C.byref(d), C.sizeof(d), _args(0, 0), 2) == -1    # descriptor size of another kind


def test_load_rejects_missing_and_foreign_files(tmp_path):
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.sdk_plan_load(str(tmp_path / "missing.engine").encode(), C.byref(h)) == -1
    f = tmp_path / "junk.engine"
    f.write_bytes(b"not an engine file at all" * 10)
    assert lib.sdk_plan_load(str(f).encode(), C.byref(h)) == -1
    assert b"engine file" in lib.sdk_last_error()


def test_slots_encode_floats_and_negative_ints():
    from stable_diffusion_pytorch_b200.unet import StepProgram
    s = StepProgram._slots("sdk_gather_row", (0x1000, 20160, 50, 0x2000, -1, 0x3000))
    assert list(s) == [0x1000, 20160, 50, 0x2000, 0xFFFFFFFFFFFFFFFF, 0x3000]
    s = StepProgram._slots("sdk_layernorm", (1, 2, 3, 1e-5, 4, 1, 128, 320))
    assert s[3] == 0x3727C5AC
    with pytest.raises(RuntimeError):
        StepProgram._slots("sdk_layernorm", (1, 2, 3))
