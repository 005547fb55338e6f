"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol declared in
include/adept_b200.h, the ctypes signature table matches the header, and compute calls fail loudly without CUDA."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "adept_b200.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    decls = re.findall(r"(?:long long|int|const char\*)\s+(adept_b200_\w+)\s*\(([^;]*?)\)\s*;", body, flags=re.S)
    return {name: [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"] for name, args in decls}


@pytest.fixture(scope="module")
def lib():
    from adept_b200 import _lib, build

    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    decl = declared_functions()
    assert len(decl) >= 12
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/adept_b200.h but not exported"


def test_signature_table_matches_header(lib):
    from adept_b200 import _lib

    decl = declared_functions()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, args in decl.items():
        assert len(args) == len(_lib.SIGNATURES[name]), (name, args)
        for a, ct in zip(args, _lib.SIGNATURES[name]):
            if "*" in a or a.startswith("void*"):
                assert ct in (ctypes.c_void_p, ctypes.c_char_p) or ct.__name__.startswith("LP_"), (name, a, ct)
            elif a.startswith("double"):
                assert ct is ctypes.c_double, (name, a)
            elif a.startswith("long long"):
                assert ct is ctypes.c_longlong, (name, a)
            elif a.startswith("int"):
                assert ct is ctypes.c_int, (name, a)


def test_version_and_error_string(lib):
    assert lib.adept_b200_version() >= 100
    assert isinstance(lib.adept_b200_last_error(), bytes)
    assert lib.adept_b200_prepare(4096) in (0, -3)  # -3 = no CUDA device in this container
    assert lib.adept_b200_prepare(48) == -2
    assert b"power of two" in lib.adept_b200_last_error()


def test_no_cpu_fallback():
    from adept_b200 import ops
    from adept_b200._lib import AdeptB200Error

    f = torch.zeros(8, 8, dtype=torch.float64)
    with pytest.raises(AdeptB200Error, match="no CPU path"):
        ops.vdfdx(f, torch.zeros(8, dtype=torch.float64), 0.1, 1.0)
    with pytest.raises(AdeptB200Error, match="no CPU path"):
        ops.collide(f, torch.zeros(8, dtype=torch.float64), 0.1, 0.1)


def test_product_never_imports_oracle():
    for py in (ROOT / "adept_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{py} imports the oracle"


def test_step_struct_layout_matches_header(tmp_path):
    """ctypes mirror of struct adept_b200_step == the C layout (compiled from include/adept_b200.h with gcc)."""
    import subprocess

    from adept_b200._lib import Species, Step, StepBwd

    src = tmp_path / "layout.c"
    fields = ["e_in", "dex", "a_out", "wave_on", "pond", "n_ex", "ex_w", "ex_tenv", "ex_wt", "fp_on", "sg_m",
              "nu_fp_space", "nu_fp_time", "f_mx", "poisson_green", "hou_li_filt", "time_row"]
    prints = "".join(f'printf("%zu\\n", offsetof(adept_b200_step, {f}));' for f in fields)
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{ROOT}/include/adept_b200.h"\n'
                   f'int main(void){{printf("%zu\\n%zu\\n%zu\\n", sizeof(adept_b200_step), sizeof(adept_b200_species), '
                   f'sizeof(adept_b200_step_bwd));{prints}return 0;}}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", str(src), "-o", str(exe)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(Step) and out[1] == ctypes.sizeof(Species) and out[2] == ctypes.sizeof(StepBwd)
    for f, off in zip(fields, out[3:]):
        assert getattr(Step, f).offset == off, f
