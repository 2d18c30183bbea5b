"""The C-ABI shared library loads and exports every symbol include/xitorch_b200.h declares (no compute)."""
import ctypes
import os
import re

from xitorch_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "xitorch_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(xt_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_all_declared_symbols():
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 10
    for nm in names:
        assert hasattr(L, nm), "missing export %s" % nm
    assert sorted(_lib.EXPORTS) == names
    assert L.xt_version() >= 100


def test_struct_layouts_match_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "xitorch_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for cname, cls in (("xt_matvec_args", _lib.MatvecArgs), ("xt_solve_args", _lib.SolveArgs),
                       ("xt_symeig_args", _lib.SymeigArgs), ("xt_hermcheck_args", _lib.HermCheckArgs)):
        end = hdr.index("} %s;" % cname)
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            parts = [p.strip() for p in decl.split(",")]
            first = parts[0].split()[-1].lstrip("*")
            fields.append(first)
            fields.extend(p.lstrip("*").strip() for p in parts[1:])
        assert fields == [f[0] for f in cls._fields_], cname


def test_workspace_queries_need_no_gpu():
    L = _lib.lib()
    assert L.xt_solve_workspace_bytes(b"cg", _lib.XT_F64, 256, 1, 3, 384, 0) > 256 * 3 * 8 * 5
    assert L.xt_solve_workspace_bytes(b"bicgstab", _lib.XT_F32, 4096, 64, 1, 100, 0) > 64 * 4096 * 4 * 8
    assert L.xt_solve_workspace_bytes(b"gmres", _lib.XT_F32, 100, 1, 2, 50, 0) > 51 * 100 * 2 * 4
    assert L.xt_solve_workspace_bytes(b"nope", 0, 10, 1, 1, 1, 0) == 0
    assert L.xt_symeig_workspace_bytes(_lib.XT_F32, 16384, 8, 128, 1) > 2 * 16384 * 128 * 4
    assert L.xt_symeig_workspace_bytes(_lib.XT_F32, 10, 8, 128, 1) == 0


def test_invalid_arguments_return_status_not_crash():
    L = _lib.lib()
    assert L.xt_block_matvec(None) != 0
    g = _lib.SolveArgs()
    assert L.xt_cg(ctypes.byref(g)) != 0
    assert b"solve" in L.xt_last_error()


def test_struct_sizes_and_offsets_match_the_c_compiler(tmp_path):
    """the ctypes mirrors against what a C compiler makes of include/xitorch_b200.h: total size and the offset of every
    field (catches a wrong field TYPE, which the name-order test above cannot)"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    structs = (("xt_matvec_args", _lib.MatvecArgs), ("xt_solve_args", _lib.SolveArgs),
               ("xt_symeig_args", _lib.SymeigArgs), ("xt_hermcheck_args", _lib.HermCheckArgs))
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "xitorch_b200.h"', 'int main(void) {']
    for cname, cls in structs:
        lines.append('  printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    for (cname, cls), line in zip(structs, out):
        parts = line.split()
        assert parts[0] == cname
        assert int(parts[1]) == ctypes.sizeof(cls), cname
        for (fname, _), off in zip(cls._fields_, parts[2:]):
            assert int(off) == getattr(cls, fname).offset, (cname, fname)
