"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/thunder_b200.h declares, refuses to run without a GPU, and its host-side integer code
(the pixel list) equals the oracle.  No compute calls are made here."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "thunder_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(thb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from thunder_b200 import capi
    lib = capi.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/thunder_b200.h but not exported"


def test_header_is_plain_c():
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = Path(d) / "t.c"
        src.write_text('#include "thunder_b200.h"\nint main(void){ thb_pf_params p; (void)p; return thb_version() > 0 ? 0 : 1; }\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(Path(d) / "t.o")])


def test_no_gpu_fails_loudly():
    from thunder_b200 import capi
    lib = capi.load()
    if lib.thb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.ThbError) as e:
        capi.Context(0)
    assert "no CPU path" in str(e.value) or "no CUDA device" in str(e.value)


@pytest.mark.parametrize("N,rU,rL", [(16, 7, 1), (16, 7, 0), (16, 5.5, 1.5), (128, 63, 0), (256, 127, 1), (200, 99, 1)])
def test_pixel_list_matches_oracle(port, N, rU, rL):
    from thunder_b200 import capi
    a, b = capi.pixel_list(N, 2, rU, rL), port.pixel_list(N, 2, rU, rL)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_product_does_not_touch_oracle():
    """the product package must not import, link or execute anything under oracle/"""
    for p in (ROOT / "thunder_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".h", ".cpp"}:
            t = p.read_text()
            assert "oracle" not in t, p
    import subprocess
    out = subprocess.check_output(["ldd", str(ROOT / "thunder_b200" / "lib" / "libthunder_b200.so")]).decode()
    assert "oracle" not in out and "thunder_ref" not in out


def test_ctypes_binding_matches_the_header():
    """every prototype of include/thunder_b200.h against the ctypes signature thunder_b200/capi.py declares for it: same number
    of parameters, pointer parameters bound as pointers, int / float / double / uint64 as such (catches binding drift)"""
    import ctypes as C
    import re
    from thunder_b200 import capi
    lib = capi.load()
    text = (ROOT / "include" / "thunder_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = re.findall(r"\b(?:int|void|const char\*)\s+(thb_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 40
    checked = 0
    for name, args in protos:
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue                                     # not used from Python
        params = [a.strip() for a in args.replace("\n", " ").split(",")] if args.strip() not in ("", "void") else []
        assert len(params) == len(fn.argtypes), (name, len(params), len(fn.argtypes))
        for prm, ct in zip(params, fn.argtypes):
            is_ptr = "*" in prm or "[" in prm
            if is_ptr:
                assert ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "_type_"), (name, prm, ct)
            elif re.match(r"^(const\s+)?double\b", prm):
                assert ct is C.c_double, (name, prm, ct)
            elif re.match(r"^(const\s+)?float\b", prm):
                assert ct is C.c_float, (name, prm, ct)
            elif re.match(r"^(const\s+)?uint64_t\b", prm):
                assert ct in (C.c_uint64, C.c_ulonglong, C.c_ulong), (name, prm, ct)
            elif re.match(r"^(const\s+)?int\b", prm):
                assert ct is C.c_int, (name, prm, ct)
        checked += 1
    assert checked >= 35
