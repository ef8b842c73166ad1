"""include/mp3b.h is a C header and libmp3b.so a C-ABI library: a C99 program (tests/c/cabi_check.c) is
compiled with gcc -pedantic -Werror, linked against the library and run -- on the CPU for what needs no GPU,
on the GPU for a decode."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    import mp3_b200
    mp3_b200.load_library()
    exe = str(tmp_path / "cabi_check")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "cabi_check.c"), "-L" + os.path.join(ROOT, "mp3_b200"), "-lmp3b",
           "-Wl,-rpath," + os.path.join(ROOT, "mp3_b200"), "-lm", "-o", exe]
    cuda_lib = "/usr/local/cuda/lib64"
    if os.path.isdir(cuda_lib):
        cmd += ["-L" + cuda_lib, "-Wl,-rpath," + cuda_lib]
    subprocess.check_call(cmd)
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not present")
def test_c99_caller_without_gpu(tmp_path):
    out = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "cabi_check: ok" in out.stdout, out.stderr


@pytest.mark.gpu
def test_c99_caller_decodes_on_gpu(tmp_path):
    out = subprocess.run([build(tmp_path), "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "gpu ok" in out.stdout, out.stderr
