"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/wavetorch_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "wavetorch_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    for n in ("wt_abi_version", "wt_last_error", "wt_query_plan", "wt_forward", "wt_backward", "wt_step_forward",
              "wt_step_backward"):
        assert n in names


def test_library_exports_every_declared_symbol():
    from wavetorch_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in _declared_functions():
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.EXPORTS) == _declared_functions()
    lib.wt_abi_version.restype = ctypes.c_int
    assert lib.wt_abi_version() == 1


def test_struct_layout_matches_header():
    """ctypes mirrors of wt_problem / wt_plan have the C sizes (8-byte aligned doubles / uint64)."""
    from wavetorch_b200 import _lib
    assert ctypes.sizeof(_lib.WtProblem) == 8 * 4 + 5 * 8 + 8 * 4
    assert ctypes.sizeof(_lib.WtPlan) == 16 * 4 + 3 * 8


def test_bad_arguments_are_rejected_without_a_gpu():
    from wavetorch_b200 import _lib
    lib = _lib.load()
    p = _lib.make_problem(0, 10, 1, 1, 0, 0, 1.0, 1.0)
    plan = _lib.WtPlan()
    assert lib.wt_query_plan(ctypes.byref(p), ctypes.byref(plan)) == -1
    assert b"bad grid" in lib.wt_last_error()
    with pytest.raises(RuntimeError, match="bad grid"):
        _lib.query_plan(p)


def test_struct_offsets_match_the_header_as_compiled_by_gcc(tmp_path):
    """Compile include/wavetorch_b200.h as plain C and compare sizeof/offsetof of every field with the ctypes mirrors."""
    import shutil
    import subprocess
    from wavetorch_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from wavetorch_b200.domain import WtSlab
    fields_p = [f[0] for f in _lib.WtProblem._fields_]
    fields_q = [f[0] for f in _lib.WtPlan._fields_]
    fields_s = [f[0] for f in WtSlab._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "wavetorch_b200.h"', 'int main(void) {',
           '  printf("%zu %zu %zu\\n", sizeof(wt_problem), sizeof(wt_plan), sizeof(wt_slab));']
    src += ['  printf("%%zu\\n", offsetof(wt_problem, %s));' % f for f in fields_p]
    src += ['  printf("%%zu\\n", offsetof(wt_plan, %s));' % f for f in fields_q]
    src += ['  printf("%%zu\\n", offsetof(wt_slab, %s));' % f for f in fields_s]
    src += ['  return 0;', '}']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.WtProblem) and int(out[1]) == ctypes.sizeof(_lib.WtPlan)
    assert int(out[2]) == ctypes.sizeof(WtSlab)
    offs = [int(v) for v in out[3:]]
    expect = [getattr(_lib.WtProblem, f).offset for f in fields_p] + [getattr(_lib.WtPlan, f).offset for f in fields_q] + \
             [getattr(WtSlab, f).offset for f in fields_s]
    assert offs == expect
