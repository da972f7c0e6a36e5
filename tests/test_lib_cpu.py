"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/m3l_b200.h
declares (no compute without a GPU); the product package never imports the oracle."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from m3l_b200 import build
    path = build.build(verbose=False)
    return ctypes.CDLL(str(path))


def test_header_symbols_are_exported(lib):
    header = (ROOT / "include" / "m3l_b200.h").read_text()
    declared = set(re.findall(r"\b(m3l_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    from m3l_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_invalid_arguments_fail_without_gpu(lib):
    lib.m3l_last_error.restype = ctypes.c_char_p
    assert lib.m3l_gemm_bf16(None, None) != 0
    assert lib.m3l_attention_fwd(None, 1, 10, 4, 64, ctypes.c_float(0.125), None, None, None) != 0
    assert b"null" in lib.m3l_last_error()


def test_product_never_imports_oracle():
    for py in (ROOT / "m3l_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py


def test_sass_is_blackwell_native(lib):
    import subprocess
    from m3l_b200 import build
    sass = subprocess.run(["cuobjdump", "-sass", str(build.LIB)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_single_thread_issue_sites_use_the_uniform_datapath(lib):
    """Regression guard for the `elect.sync` rule (DESIGN.md §4 'Measured'): no tcgen05.mma / TMA load / TMA store in
    the built objects may sit inside a VOTEU / ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop, which is what the
    compiler emits when the issuing thread is selected with `lane == 0` (100-150 clk per instruction)."""
    import shutil
    import subprocess
    from m3l_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    for name in ("gemm", "gemm_gelu", "attention", "rowblock"):
        obj = build.OBJ / f"{name}.o"
        assert obj.exists(), obj
        sass = subprocess.run([cuobjdump, "-sass", str(obj)], capture_output=True, text=True).stdout.splitlines()
        issue = [i for i, l in enumerate(sass) if re.search(r"\b(UTCHMMA|UTMALDG|UTMASTG|UTMAREDG)\b", l)]
        assert issue, f"{name}: no tcgen05 / TMA instructions found"
        bad = [i for i in issue if any("BRA.U.ANY" in l for l in sass[i + 1:i + 4])]
        assert not bad, f"{name}: {len(bad)} of {len(issue)} MMA / TMA instructions are issued from a waterfall loop"
