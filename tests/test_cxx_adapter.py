"""The C++ host adapters (include/microscopes_b200/gpu_models.hpp) implementing the reference's
models/base.hpp interface over the C ABI.  CPU: they compile -- against our declaration of the
interface and, where /root/reference exists, against the reference's real headers.  GPU: the
reference's own driver loop (bin/perf_group.cpp:76-125) runs on them."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX_DIR = os.path.join(ROOT, "tests", "cxx")
LIBDIR = os.path.join(ROOT, "common_b200", "csrc")
REF = "/root/reference"


def _build(out, extra):
    cmd = ["g++", "-std=c++14", "-O1"] + extra + ["-I" + os.path.join(ROOT, "include"), "test_adapter.cpp"]
    if extra:
        cmd.append(os.path.join(REF, "src", "common", "runtime_type.cpp"))
    cmd += ["-L" + LIBDIR, "-lmscope_b200", "-Wl,-rpath," + LIBDIR, "-o", out]
    subprocess.check_call(cmd, cwd=CXX_DIR)


def _fresh(path):
    srcs = [os.path.join(CXX_DIR, "test_adapter.cpp"), os.path.join(ROOT, "include", "microscopes_b200", "gpu_models.hpp"),
            os.path.join(ROOT, "include", "microscopes_b200", "plugin_api.hpp"), os.path.join(ROOT, "include", "mscope_b200.h")]
    return os.path.exists(path) and all(os.path.getmtime(path) >= os.path.getmtime(s) for s in srcs)


def test_adapters_compile_against_our_interface():
    out = os.path.join(CXX_DIR, "test_adapter")
    if not _fresh(out):
        _build(out, [])
    assert subprocess.call([out, "compile-only"]) == 0


def test_adapters_compile_against_the_reference_headers():
    if not os.path.isdir(os.path.join(REF, "include", "microscopes")):
        pytest.skip("reference tree not present (GPU box): the prebuilt binary is used there")
    out = os.path.join(CXX_DIR, "test_adapter_refhdr")
    if not _fresh(out):
        _build(out, ["-DMSB_USE_REFERENCE_HEADERS", "-include", "cstdint", "-include", "functional", "-include", "sys/types.h",
                     "-I" + os.path.join(REF, "include"), "-I" + os.path.join(ROOT, "oracle", "ref_shim")])
    assert subprocess.call([out, "compile-only"]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("binary", ["test_adapter", "test_adapter_refhdr"])
def test_perf_group_loop_on_gpu_adapters(binary):
    path = os.path.join(CXX_DIR, binary)
    if not os.path.exists(path):
        if binary == "test_adapter":
            _build(path, [])
        else:
            pytest.skip("built only where the reference tree exists")
    res = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all ok" in res.stdout
