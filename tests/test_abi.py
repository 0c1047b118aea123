"""The C-ABI library: builds, loads, exports every symbol include/wenet_b200.h declares, and refuses to run
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from wenet_b200 import engine
    return engine.load_library()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "wenet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree(lib):
    from wenet_b200 import engine
    decl = declared_symbols()
    bound = sorted(n for n, _, _ in engine.ABI)
    assert decl == bound, (set(decl) ^ set(bound))
    for name in decl:
        assert hasattr(lib, name), name
    assert lib.wb_abi_version() == 1


def test_struct_layouts_match_header(lib):
    from wenet_b200 import engine
    assert C.sizeof(engine.WbConfig) == 64
    assert engine.CODEWORD_DTYPE.itemsize == 280
    assert C.sizeof(engine.WbStats) == 4 * 9 + 4 * 3 + 4 * 8 * 160 + 4 + 4 * 512 + 8 + 8 + 4   # incl. alignment pad


def test_no_gpu_means_no_engine(lib):
    """on a machine without CUDA the engine must fail loudly (WB_ENODEV), never fall back to a CPU path"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from wenet_b200 import engine
    with pytest.raises(engine.WbError) as ei:
        engine.Engine(1, chunk_samples=4096)
    assert ei.value.code == engine.WB_ENODEV
    assert "no CPU fallback" in str(ei.value)


def test_bad_arguments_are_errors_not_aborts(lib):
    from wenet_b200 import engine
    cfg = engine.WbConfig()
    h = C.c_void_p()
    assert lib.wb_create(C.byref(cfg), C.byref(h)) == engine.WB_EINVAL        # struct_size = 0
    assert b"struct_size" in lib.wb_last_error()
    assert lib.wb_create(None, C.byref(h)) == engine.WB_EINVAL
    assert lib.wb_process(None) == engine.WB_EINVAL


def test_product_never_imports_the_oracle():
    """the product package must not reference oracle/ (parity claims depend on it)"""
    pkg = os.path.join(ROOT, "wenet_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")) and f != "wb_tables.h":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_c_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/decode_file.c (a plain C99 caller of include/wenet_b200.h, the shape of a program that links the
    reference's fsk.c / mpdecode_core.c today) compiles with -Wall -Wextra, links against the product library and,
    where no CUDA device is visible, stops with the library's own message instead of decoding on the CPU"""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = str(tmp_path / "decode_file")
    libdir = os.path.join(root, "wenet_b200")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
                        os.path.join(root, "examples", "decode_file.c"), "-L" + libdir, "-lwenet_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    assert subprocess.run([exe], stderr=subprocess.PIPE).returncode == 1            # usage
    try:
        import torch
        if torch.cuda.is_available():
            return                                                                    # the GPU suite covers the rest
    except Exception:
        pass
    src = tmp_path / "in.cu8"
    src.write_bytes(bytes(4096))
    r = subprocess.run([exe, str(src), str(tmp_path / "out.bin")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 2 and b"no CPU fallback" in r.stderr


def test_no_contracted_packed_arithmetic_in_the_library():
    """The kernels use mul.rn.f32x2 / add.rn.f32x2 where that is the same floats as the reference's separate products and
    sums.  ptxas 12.9 contracts a packed product that feeds a packed sum into ONE FFMA2 -- a single rounding -- even for
    the explicit .rn forms and with --fmad false (found in round 2: an oscillator step written as two packed products +
    one packed sum came out as FMUL2 + FFMA2).  The shipped code never feeds a packed product into a packed sum, so the
    library must not contain a single FFMA2; this catches a build where an edit re-opened the door."""
    import shutil
    import subprocess
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wenet_b200", "libwenet_b200.so")
    if not os.path.exists(so) or shutil.which("cuobjdump") is None:
        pytest.skip("no built library or no cuobjdump here")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, timeout=600).stdout
    assert sass.count("FMUL2") > 1000 and sass.count("FADD2") > 100      # the packed forms are there ...
    assert sass.count("FFMA2") == 0                                       # ... and none of them was contracted
