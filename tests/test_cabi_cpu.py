"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/mp3b.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mp3_b200 import _build
    import mp3_b200
    _build.build_lib()
    return mp3_b200.load_library()


def test_exports_match_header(lib):
    import mp3_b200
    hdr = open(os.path.join(ROOT, "include", "mp3b.h")).read()
    declared = set(re.findall(r"\b(mp3b_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(mp3_b200.EXPORTS), declared ^ set(mp3_b200.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_strerror(lib):
    assert lib.mp3b_abi_version() == 1
    assert lib.mp3b_strerror(0) == b"ok"
    assert b"CUDA" in lib.mp3b_strerror(-5)


def test_bad_options_are_rejected_before_any_device_is_touched(lib):
    import mp3_b200
    for field, val in (("pcm_format", 2), ("pcm_format", -1), ("indexer", 5), ("pipeline", 9), ("host_threads", -3),
                       ("struct_size", 4)):
        o = mp3_b200.Opts()
        lib.mp3b_opts_default(ctypes.byref(o))
        setattr(o, field, val)
        ctx = ctypes.c_void_p()
        assert lib.mp3b_ctx_create(0, ctypes.byref(o), ctypes.byref(ctx)) == -1, field
        assert not ctx.value
    assert b"Layer I" not in lib.mp3b_strerror(-4) and b"2.5" not in lib.mp3b_strerror(-4)  # both are decoded


def test_no_cpu_fallback(lib):
    import mp3_b200
    if lib.mp3b_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mp3_b200.Mp3bError) as e:
        mp3_b200.Decoder(device=0)
    assert e.value.status == -5


def test_struct_sizes_match_header(lib):
    import mp3_b200
    assert ctypes.sizeof(mp3_b200.Opts) == 36
    assert ctypes.sizeof(mp3_b200.TagInfo) == 40
    assert ctypes.sizeof(mp3_b200.StreamInfo) == 56
    assert ctypes.sizeof(mp3_b200.Stats) == 96
    o = mp3_b200.Opts()
    lib.mp3b_opts_default(ctypes.byref(o))
    assert o.struct_size == 36


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mp3_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".c")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "l3_oracle" not in txt.replace("oracle/l3_oracle.c", "") and "from oracle" not in txt \
                    and "import oracle" not in txt, os.path.join(dp, f)
