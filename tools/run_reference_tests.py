"""
Run the reference's OWN test files against this package (build container only: needs /root/reference).

    python tools/run_reference_tests.py [--standin | --emulated] [pytest args]        e.g.  -k "not methods and not large"

The reference's test files are copied to a scratch directory (never into the repo), next to a conftest.py that calls
`xitorch_b200.install_as_xitorch()` before anything imports `xitorch`, and that loads the reference's test helpers
(`xitorch/_tests/utils.py`) under their original module name.  Everything the tests import as `xitorch.*` is then this
package.  Cases that run a Krylov method on CPU tensors fail by construction (there is no CPU path): on a box without
a GPU use  -k "not methods and not large"  for test_linop_fcns.py -- or `--standin`, which replaces the CUDA library
by tests/standin_engine.py (numpy on the argument structs) so that the Krylov `method=` cases exercise this package's
host logic end to end on CPU tensors (the numerics inside the engine are then numpy's, not the kernels').
`--emulated` goes one step further: the engines' own sources (csrc/{symeig,solve,gmres}.cu) as a host build
(tools/emu_engine; only the block matvec is a stand-in), so the reference's method tests run the shipped algorithms.
"""
import os
import shutil
import subprocess
import sys
import tempfile

REF = os.environ.get("XITORCH_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["test_linop.py", "test_linop_fcns.py", "test_editable_module.py", "test_pure_function.py", "test_jac.py",
         "test_optimize.py", "test_debug.py", "test_wrap_nnmodule.py", "test_memleak.py"]

CONFTEST = '''
import importlib.util, sys, types
sys.path.insert(0, %r)
import xitorch_b200
xitorch_b200.install_as_xitorch()
pkg = types.ModuleType("xitorch._tests"); pkg.__path__ = []; sys.modules["xitorch._tests"] = pkg
spec = importlib.util.spec_from_file_location("xitorch._tests.utils", %r)
mod = importlib.util.module_from_spec(spec); sys.modules["xitorch._tests.utils"] = mod; spec.loader.exec_module(mod)
mode = %r
if mode:
    sys.path.insert(0, %r)

    class _Patch(object):
        def setattr(self, obj, name, value):
            setattr(obj, name, value)

    if mode == "standin":
        import standin_engine
        standin_engine.install(_Patch())
    else:
        import tempfile, emu_engine_lib
        emu_engine_lib.install(_Patch(), emu_engine_lib.build(tempfile.mkdtemp(prefix="xt_emu_")))
'''


def main():
    scratch = tempfile.mkdtemp(prefix="xt_reftests_")
    try:
        for f in FILES:
            shutil.copy(os.path.join(REF, "xitorch", "_tests", f), scratch)
        args = sys.argv[1:]
        standin = ""
        for flag in ("--standin", "--emulated"):
            if flag in args:
                args.remove(flag)
                standin = flag[2:]
        with open(os.path.join(scratch, "conftest.py"), "w") as fh:
            fh.write(CONFTEST % (ROOT, os.path.join(REF, "xitorch", "_tests", "utils.py"), standin,
                                 os.path.join(ROOT, "tests")))
        args = args or ["-q"]
        return subprocess.call([sys.executable, "-m", "pytest", "-p", "no:cacheprovider", *args], cwd=scratch)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
