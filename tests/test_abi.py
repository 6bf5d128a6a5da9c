"""CPU test: the C-ABI shared library loads without a GPU and exports every symbol the headers declare."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADERS = ["micropp_c.h", "micropp_b200_ext.h", "mgpu.h", "material_base.h"]
DECL = re.compile(r"^\s*(?:const\s+)?(?:unsigned\s+long\s+long|struct\s+\w+\s*\*?|[A-Za-z_]\w*)\s*\**\s*"
                  r"((?:micropp3x?|mgpu|material)_\w+)\s*\(", re.M)


def declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return sorted(set(DECL.findall(text)))


@pytest.fixture(scope="module")
def lib():
    import micropp_b200
    assert micropp_b200.LIB_PATH.exists(), "build first: python -m micropp_b200.build"
    return ctypes.CDLL(str(micropp_b200.LIB_PATH))


@pytest.mark.parametrize("header", HEADERS)
def test_every_declared_symbol_is_exported(lib, header):
    names = declared(header)
    assert names, header
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"{header}: not exported: {missing}"


def test_reference_c_api_surface_is_complete():
    # include/micropp_c.h:37-70 of the reference (+ micropp3_output2, defined in src/micropp_c.cpp:113)
    want = ["micropp3_new", "micropp3_free", "micropp3_set_strain", "micropp3_get_stress", "micropp3_get_ctan",
            "micropp3_homogenize", "micropp3_homogenize_linear", "micropp3_get_cost", "micropp3_has_converged",
            "micropp3_has_subiterated", "micropp3_output", "micropp3_output2", "micropp3_update_vars",
            "micropp3_print_info", "micropp3_is_non_linear", "micropp3_get_non_linear_gps", "micropp3_write_restart",
            "micropp3_read_restart", "material_set", "material_print"]
    got = set(declared("micropp_c.h")) | set(declared("material_base.h")) | {"micropp3_output2"}
    assert set(want) <= got, sorted(set(want) - got)


def test_no_gpu_means_loud_failure_not_fallback():
    # device_count is the only compute-free entry point; constructing a solver without a device aborts the process
    import micropp_b200
    micropp_b200.load()
    n = micropp_b200.device_count()
    assert n >= 0
    src = (ROOT / "micropp_b200" / "csrc" / "mgpu_kernels.cu").read_text()
    assert "this library has no CPU path" in src
    pkg = "".join(p.read_text() for p in (ROOT / "micropp_b200").glob("*.py"))
    assert "oracle" not in pkg.replace("# oracle", ""), "the product package must never import the oracle"


def test_ctypes_mirrors_match_the_c_structs():
    """Every ctypes.Structure the bindings hand to the library as an OUTPUT buffer has the size the C side writes
    (a 16-byte mismatch of SlotState once corrupted the heap of the test process)."""
    import ctypes as C
    import micropp_b200
    from micropp_b200 import slab
    lib = C.CDLL(str(micropp_b200.LIB_PATH))
    lib.mgpu_slot_state_size.restype = C.c_int
    assert lib.mgpu_slot_state_size() == C.sizeof(slab.SlotState)
    assert C.sizeof(slab.SlabHandle) == 72          # struct micropp3x_slab_handle {char mail[64]; int op, pad;}
