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
            os.path.join(ROOT, "include", "microscopes_b200", "wire.hpp"),
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


def test_adapter_bags_are_the_reference_wire_messages():
    """gpu_group::get_ss / gpu_hypers::get_hp emit the serialized Group / Shared messages of distributions.hpp:300-314,
    355-369 -- byte for byte what common_b200/wire.py writes (held equal to the protobuf runtime in tests/test_wire.py) --
    and set_ss / set_hp parse bags written by the Python host, packed repeated fields included."""
    import struct
    from common_b200 import wire
    out = os.path.join(CXX_DIR, "test_adapter")
    if not _fresh(out):
        _build(out, [])
    names = {0: "bb", 1: "bnb", 2: "gp", 3: "nich", 4: "dd", 5: "niw", 6: "bbnc", 7: "dm"}
    hp_keys = {"bb": ["alpha", "beta"], "bnb": ["alpha", "beta", "r"], "gp": ["alpha", "inv_beta"], "nich": ["mu", "kappa", "sigmasq", "nu"],
               "dd": ["alphas"], "niw": ["mu", "kappa", "psi", "nu"], "bbnc": ["alpha", "beta"], "dm": ["alphas"]}
    ss_keys = {"bb": ["heads", "tails"], "bnb": ["count", "sum"], "gp": ["count", "sum", "log_prod"], "nich": ["count", "mean", "count_times_variance"],
               "dd": ["counts"], "niw": ["count", "sum_x", "sum_xxT"], "bbnc": ["p", "heads", "tails"], "dm": ["counts", "ratio"]}

    def count(key, dim):
        return {"alphas": dim, "counts": dim, "mu": dim, "psi": dim * dim, "sum_x": dim, "sum_xxT": dim * dim}.get(key, 1)

    def expected(nm, dim):
        hp, base = {}, 1.5
        for k in hp_keys[nm]:
            n = count(k, dim) if nm in ("dd", "dm", "niw") else 1
            b = 21.0 if k == "r" else base
            hp[k] = [b + i for i in range(n)] if nm in ("dd", "dm", "niw") and k in ("alphas", "mu", "psi") else b
            base += 10.0
        ss, base = {}, 3.0
        for k in ss_keys[nm]:
            n = count(k, dim) if nm in ("dd", "dm", "niw") else 1
            b = 0.25 if k == "p" else base
            ss[k] = [b + i for i in range(n)] if k in ("counts", "sum_x", "sum_xxT") else b
            base += 7.0
        return wire.encode(nm + ".Shared", hp), wire.encode(nm + ".Group", ss)

    # bags written by the Python host, the repeated fields of dd in PACKED form (protobuf accepts both on input)
    feed = []
    for fam, dim in [(0, 0), (3, 0), (5, 3), (7, 4)]:
        hp, ss = expected(names[fam], dim)
        feed.append("%d %d %s %s" % (fam, dim, hp.hex() or "-", ss.hex() or "-"))
    packed_alphas = bytes([0x0A, 20]) + b"".join(struct.pack("<f", 1.5 + i) for i in range(5))
    packed_counts = bytes([0x0A, 5, 3, 4, 5, 6, 7])
    feed.append("4 5 %s %s" % (packed_alphas.hex(), packed_counts.hex()))
    feed.append("4 5 %s %s" % (packed_alphas.hex(), bytes([0x0A, 2, 3, 4]).hex()))      # 2 counts for a 5-category group
    res = subprocess.run([out, "bags"], input="\n".join(feed) + "\n", capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stderr
    first, second = res.stdout.split("--\n")
    for line in first.strip().splitlines():
        fam, dim, hp, ss = line.split()
        want_hp, want_ss = expected(names[int(fam)], int(dim))
        assert hp == (want_hp.hex() or "-"), names[int(fam)]
        assert ss == (want_ss.hex() or "-"), names[int(fam)]
    lines = second.strip().splitlines()
    for got, fed in zip(lines[:4], feed[:4]):
        assert got == fed                                   # what went in through set_* comes back out of get_*
    want_hp, want_ss = expected("dd", 5)
    assert lines[4] == "4 5 %s %s" % (want_hp.hex(), want_ss.hex())   # packed input, canonical (unpacked) output
    assert "error wrong dimension" in lines[5]              # distributions.hpp:436


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
