"""The C-ABI library loads and exports every symbol include/fs2_b200.h declares; the ctypes signature table
matches the header (no compute calls: runs without a GPU)."""
import ctypes
import importlib
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fs2_b200.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return importlib.import_module("fine-grained-emotional-control-of-tts_b200._lib")


def _decls():
    h = open(HEADER).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:int|long long|const char\*)\s+(fs2_\w+)\s*\((.*?)\)\s*;", h, flags=re.S):
        out[m.group(1)] = [a.strip() for a in m.group(2).split(",")]
    return out


def test_every_declared_symbol_is_exported(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    decls = _decls()
    assert len(decls) >= 30
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/fs2_b200.h but not exported"
    assert lib.fs2_abi_version() == 1


def test_signature_table_matches_header(built):
    decls = _decls()
    for name, sig in built.SIGNATURES.items():
        codes = ""
        for a in decls[name]:
            if "*" in a:
                codes += "p"
            elif a.startswith("unsigned long long"):
                codes += "Q"
            elif a.startswith("long long"):
                codes += "q"
            elif a.startswith("float"):
                codes += "f"
            elif a.startswith("int"):
                codes += "i"
            else:
                raise AssertionError((name, a))
        assert codes == sig, f"{name}: header {codes} vs table {sig}"


def test_struct_sizes_match_c(built, tmp_path):
    """ctypes mirrors of the descriptor structs have the C compiler's size."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "fs2_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(Fs2Gemm), '
                   'sizeof(Fs2LnFwd), sizeof(Fs2LnBwd), sizeof(Fs2PackItem), sizeof(Fs2GemmLn));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(built.Fs2Gemm), ctypes.sizeof(built.Fs2LnFwd), ctypes.sizeof(built.Fs2LnBwd),
                     ctypes.sizeof(built.Fs2PackItem), ctypes.sizeof(built.Fs2GemmLn)]


def test_struct_field_offsets_match_c(built, tmp_path):
    """Every field of every ctypes descriptor mirror sits at the offset the C compiler gives it in include/fs2_b200.h
    (a size check alone would miss two swapped fields of the same type)."""
    structs = {"Fs2Gemm": built.Fs2Gemm, "Fs2LnFwd": built.Fs2LnFwd, "Fs2LnBwd": built.Fs2LnBwd, "Fs2PackItem": built.Fs2PackItem,
               "Fs2GemmLn": built.Fs2GemmLn}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fs2_b200.h"', 'int main(){']
    names = []
    for sn, cls in structs.items():
        for fn, _ in cls._fields_:
            lines.append(f'printf("%zu\\n", offsetof({sn}, {fn}));')
            names.append((sn, fn, getattr(cls, fn).offset))
    lines.append('return 0;}')
    src = tmp_path / "off.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "off"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)   # unknown field = compile error
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert len(got) == len(names)
    for (sn, fn, off), g in zip(names, got):
        assert off == g, f"{sn}.{fn}: ctypes offset {off}, C offset {g}"


def test_no_cpu_fallback(built):
    import torch
    pkg = importlib.import_module("fine-grained-emotional-control-of-tts_b200")
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4)
    tok = torch.randint(1, 90, (2, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(tok, torch.zeros(2, dtype=torch.long), torch.ones(2, 8, dtype=torch.long), torch.zeros(2, 8), torch.zeros(2, 8),
          intensity=torch.zeros(2, 8, 5))


def test_measurement_hooks_validate_their_arguments(built):
    """The kernel-selection hooks are host-only state setters: bad values return a non-zero status, good ones 0."""
    lib = ctypes.CDLL(built.LIB_PATH)
    for fn in (lib.fs2_lr_bulk_rows, lib.fs2_lr_tune, lib.fs2_ln_tune):
        fn.argtypes = [ctypes.c_int]
        fn.restype = ctypes.c_int
    assert lib.fs2_lr_bulk_rows(7) != 0 and lib.fs2_lr_bulk_rows(256) != 0
    for rows in (0, 16, 32, 64, 128, 8):             # ends on the default
        assert lib.fs2_lr_bulk_rows(rows) == 0
    assert lib.fs2_lr_tune(3) != 0 and lib.fs2_lr_tune(4) == 0
    assert lib.fs2_ln_tune(0) == 0 and lib.fs2_ln_tune(1) == 0
