"""CPU tests of the drop-in boundary: libnwb.so loads, exports every symbol
include/nwb.h declares, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "nwb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nwb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib(nwb):
    if not os.path.exists(nwb.LIB_PATH):
        nwb.build()
    return nwb.load_library()


def test_header_symbols_are_exported(lib, nwb):
    names = declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nwb.h but not exported by libnwb.so"
    # the Python mirror binds exactly the header
    assert sorted(nwb.EXPORTED_SYMBOLS) == names


def test_header_is_c99(tmp_path):
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "nwb.h"\nint main(void){return NWB_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(c), "-o", str(tmp_path / "t.o")], check=True)


def test_no_cpu_fallback(lib, nwb):
    if lib.nwb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nwb.NwbError) as e:
        nwb.fill(b"GCATGCU", b"GATTACA", 1, 1, 1)
    assert e.value.code == -4  # NWB_ERR_NO_DEVICE
    with pytest.raises(nwb.NwbError):
        nwb.Plan(100, 100)


def test_error_strings(lib):
    assert lib.nwb_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.nwb_strerror(-4)
    assert lib.nwb_abi_version() >= 1
    # the queue entry points refuse a missing plan like every other plan call (no device needed to say so)
    assert lib.nwb_plan_run_pipelined(None, 1, 1, 1, None) == -1
    assert lib.nwb_plan_attach_right(None, None) == -1
