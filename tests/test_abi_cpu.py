"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/mcfost_b200.h declares, its struct layout matches the ctypes
mirror, and it fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from mcfost_b200 import abi, api
from mcfost_b200 import build as mcb_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcfost_b200.h")


@pytest.fixture(scope="module")
def lib():
    mcb_build.build()          # nvcc cross-compiles sm_100a without a GPU
    return C.CDLL(api.LIB_PATH)


def test_library_exports_every_declared_symbol(lib):
    src = open(HEADER).read()
    declared = sorted(set(re.findall(r"\b(mcfost_b200_\w+)\s*\(", src)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(api.EXPORTS) == declared


def test_struct_layout_matches_ctypes_mirror():
    """sizeof/offsetof from the real header (compiled with gcc) vs the ctypes Structures."""
    structs = {"mcb_grid": abi.mcb_grid, "mcb_opacity": abi.mcb_opacity, "mcb_emission": abi.mcb_emission,
               "mcb_run_params": abi.mcb_run_params, "mcb_tallies": abi.mcb_tallies,
               "mcb_grains": abi.mcb_grains}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for sname, st in structs.items():
        lines.append(f'printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        cfile, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(cfile, "w").write("\n".join(lines))
        subprocess.run(["/usr/bin/gcc", "-o", exe, cfile], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for sname, st in structs.items():
        assert int(got[sname]) == C.sizeof(st), sname
        for fname, _ in st._fields_:
            assert int(got[f"{sname}.{fname}"]) == getattr(st, fname).offset, f"{sname}.{fname}"


def test_no_gpu_means_loud_failure_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mcfost_b200 import synthetic as S
    P = S.ref41_like(n_photons_eq_th=10, dark_zone=False, n_rad=10, nz=5, n_rad_in=2)
    with pytest.raises(api.McfostB200Error) as e:
        api.PhotonLoop(P)
    assert e.value.code == abi.MCB_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "mcfost_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle.binding" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
                assert "liboracle" not in txt, f
