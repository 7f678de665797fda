"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol that
include/hfnet_b200.h declares (and nothing in the Python binding is missing from the header), and fails loudly --
never silently falls back -- when there is no B200."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "hfnet_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(hfb_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_are_exported(native_lib):
    from hfnet_slam_b200 import lib
    declared = _declared()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(native_lib, name), f"{name} declared in include/hfnet_b200.h but not exported"
    assert set(lib.SIGNATURES) == declared, (set(lib.SIGNATURES) ^ declared)
    assert native_lib.hfb_version() == 100


def test_struct_layouts_match_header():
    from hfnet_slam_b200 import lib
    assert C.sizeof(lib.hfb_config) == 32
    assert C.sizeof(lib.hfb_features) == 6 * 8 + 8 * 4 + 4 + 4      # 6 pointers, n_per_level[8], n_total, padding
    assert lib.hfb_lba_problem.K.offset % 4 == 0 and C.sizeof(lib.hfb_lba_stats) == 40


def test_sass_is_blackwell_native():
    """The shipped cubin must contain tcgen05 / TMA instructions (UTCHMMA, UTMALDG, LDTM); warp-level HMMA is allowed in
    the stem kernel (its K = 9 / K = 24 products read an im2col gather out of the shared image patch) and in the global
    head (NetVLAD on a 64-pixel group: 32 x 64 x 240 tiles; the FC is a skinny <= 16-row product streamed once from HBM
    with operands loaded in fragment order) -- every other GEMM-shaped kernel is tcgen05."""
    import shutil
    import subprocess
    from hfnet_slam_b200 import lib
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    for fn in sass.split("Function : ")[1:]:
        name = fn.split("\n", 1)[0]
        if re.search(r"\bHMMA\b", fn):
            assert any(k in name for k in ("stem_kernel", "vlad_kernel", "fc_mma_kernel")), f"legacy HMMA in {name}"


def test_no_gpu_is_a_loud_error(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hfnet_slam_b200.lib import Context, HfbError
    with pytest.raises(HfbError):
        Context(height=64, width=64)


def test_bad_config_rejected_without_touching_the_device(native_lib):
    from hfnet_slam_b200 import lib
    h = C.c_void_p()
    cfg = lib.hfb_config(0, 8, 8, 1, 1.2, 100, 1, 1)          # too small
    assert native_lib.hfb_create(C.byref(cfg), C.byref(h)) == 1 and not h
    cfg = lib.hfb_config(0, 480, 752, 9, 1.2, 100, 1, 1)      # too many levels
    assert native_lib.hfb_create(C.byref(cfg), C.byref(h)) == 1
    assert native_lib.hfb_last_error(None) == b"null context"
    assert native_lib.hfb_sync(None) == 1 and native_lib.hfb_kfdb_size(None) == 0


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under hfnet_slam_b200/ may import it."""
    for p in (ROOT / "hfnet_slam_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p


def test_reference_side_shim_compiles_and_links(native_lib, tmp_path):
    """include/HFNetB200Model.h (the BaseModel subclass a reference maintainer adds) compiles as C++14 against the
    reference's interface (stand-ins for OpenCV) and links against the C-ABI library."""
    import shutil
    import subprocess
    from hfnet_slam_b200 import lib
    if not shutil.which("g++"):
        pytest.skip("g++ not on PATH")
    exe = tmp_path / "shim_check"
    for src in ("shim_compile_check.cpp", "shim_run.cpp"):        # the second one is executed by tests/test_shim_gpu.py
        r = subprocess.run(["g++", "-std=c++14", "-Wall", f"-I{ROOT / 'include'}", f"-I{ROOT / 'tests' / 'native'}", "-o", str(exe),
                            str(ROOT / "tests" / "native" / src), str(lib.LIB_PATH),
                            f"-Wl,-rpath,{lib.LIB_PATH.parent}"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        if src == "shim_compile_check.cpp":
            assert subprocess.run([str(exe)]).returncode == 0


def test_header_is_plain_c():
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not on PATH")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-fsyntax-only", "-x", "c",
                        str(ROOT / "include" / "hfnet_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
