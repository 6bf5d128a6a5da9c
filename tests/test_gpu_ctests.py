"""-m gpu: the reference's OWN test programs (test/*.cpp of gagiuntoli/Micropp, unmodified, compiled where they lie by
`make -C oracle ctests`) linked against THIS repository's include/ + libmicropp_b200.so, run on the GPU.

The first block is every command the reference registers with ctest (test/CMakeLists.txt:53-64); the programs carry
their own assertions (golden stress tables of benchmark-{elastic,plastic,damage}.cpp:40-51, ELL tables of
test_ell_1.cpp:61-111, invariants of test3d_4.cpp:109-115 / test3d_5.cpp:107-115), so exit code 0 = the reference's
own acceptance criterion.  The second block is the unregistered programs that exercise the drop-in boundary further
(C API, use_A0, restart files, VTU output, the protected FE stages through subclassing).
"""
import os
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
CT = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "ctests"

REGISTERED = [  # test/CMakeLists.txt:53-64
    ("test3d_1", ["5", "0", "10"]),
    ("test3d_4", ["5", "3", "10"]),
    ("test3d_5", ["5", "5", "5", "2", "10"]),
    ("test_ell_1", []),
    ("test_ell_2", []),
    ("test_util_1", []),
    ("test_material", ["5"]),
    ("benchmark-elastic", []),
    ("benchmark-plastic", []),
    ("benchmark-damage", []),
    ("test_damage", ["10"]),
]
EXTRA = [
    ("test3d_2", ["5", "4", "5"]),
    # plain-C driver of include/micropp_c.h (tests/capi/c_api_drive.c: test/test3d_6.c restated with valid arguments)
    ("c_api_drive", ["5", "0", "10"]),
    ("benchmark-mic-2", ["6", "0", "5"]),
    ("test_cg", ["8", "10"]),                          # subclasses micropp<3>: protected FE stages + free ELL functions
    ("test_A0", ["6", "1", "5"]),
    ("test_restart", ["5", "12"]),
    ("test_get_elem_nodes", []),
    ("test_MIC3D_8", ["6", "0", "3"]),
    ("test_print_vtu_1", ["5", "1"]),
    ("test_omp", ["5", "3", "2"]),
]


def run(name, args, cwd):
    exe = CT / name
    if not exe.exists():
        pytest.fail(f"{exe} missing: run `make -C oracle ctests` where /root/reference exists "
                    "(the binaries travel with the gpurun snapshot)")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run([str(exe), *args], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)


@pytest.mark.parametrize("name,args", REGISTERED, ids=[n for n, _ in REGISTERED])
def test_reference_registered_ctest(name, args, tmp_path):
    r = run(name, args, tmp_path)
    assert r.returncode == 0, r.stdout[-3000:]


@pytest.mark.parametrize("name,args", EXTRA, ids=[n for n, _ in EXTRA])
def test_reference_extra_program(name, args, tmp_path):
    r = run(name, args, tmp_path)
    assert r.returncode == 0, r.stdout[-3000:]
