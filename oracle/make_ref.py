"""ORACLE support (test infrastructure): build recipe of ``oracle/_ref`` -- the UNMODIFIED reference, byte-compiled.

The reference's hot path is pure Python (``backend/app/pipeline.py``, ``chain.py``, ``modules/*``): there is nothing to link, so
"building" it means byte-compiling the modules the path imports, from the sources WHERE THEY LIE under ``/root/reference``, into
compiled-code files under ``oracle/_ref/backend/`` (``<module>.bin``: the ``.pyc`` layout -- 16-byte header + marshalled code object --
under another extension, because the GPU-box snapshot drops ``*.pyc``; ``oracle/ref_harness.py`` imports them through a small
finder).  Git-ignored, NOT gpurun-ignored: like our own ``.so`` the directory travels to the GPU box, where ``/root/reference`` does
not exist.  No reference source text is copied into the repository.

What gets compiled is decided by importing the path through ``oracle/ref_harness.py`` and listing the modules that came from
the reference tree, plus the two numeric helpers of ``app/routers/tools.py`` (``_true_peak_dbfs``, ``_loudness_range_lu``), whose
function definitions alone are compiled into ``app/_tools_numeric.bin`` (the router module itself drags FastAPI in).

Users: ``bench.py --impl reference`` and the ``cpu_baseline`` leg (``kind: "reference"``), ``tests/test_oracle_vs_reference.py``
(which prefers the source tree when it exists).  ``__graft_entry__.build()`` runs this where ``/root/reference`` is present.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(_HERE, "_ref", "backend")
SRC = "/root/reference/backend"
EXT = ".bin"


def _write_pyc(code, dst: str, mtime: int = 0, size: int = 0) -> None:
    from importlib._bootstrap_external import _code_to_timestamp_pyc
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "wb") as f:
        f.write(_code_to_timestamp_pyc(code, mtime, size))


def build(verbose: bool = True) -> str | None:
    """Byte-compile the reference modules of the mastering path into oracle/_ref/backend; returns the directory (None when the
    reference tree is absent: the GPU box uses the prebuilt files)."""
    if not os.path.isfile(os.path.join(SRC, "app", "pipeline.py")):
        return None
    if os.path.dirname(_HERE) not in sys.path:
        sys.path.insert(0, os.path.dirname(_HERE))
    os.environ["MM_REFERENCE_BACKEND"] = SRC
    from oracle import ref_harness
    if ref_harness.REFERENCE_BACKEND != SRC:            # imported earlier against the prebuilt copy: list modules from a fresh process
        import subprocess
        subprocess.run([sys.executable, os.path.abspath(__file__)], check=True, env={**os.environ, "MM_REFERENCE_BACKEND": SRC})
        return OUT
    ref_harness.load()
    mods = sorted({m.__file__ for m in list(sys.modules.values())
                   if getattr(m, "__file__", None) and os.path.abspath(m.__file__).startswith(SRC + os.sep) and m.__file__.endswith(".py")})
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    for path in mods:
        rel = os.path.relpath(path, SRC)
        with open(path, "rb") as f:
            data = f.read()
        # The code objects name the compiled file itself as their file: numba's `cache=True` (pipeline.py:25-26) stats the function's
        # file for its cache index and refuses to decorate when it does not exist -- /root/reference does not on the GPU box,
        # /root/repo/oracle/_ref/... does (there /root/repo is a link to the snapshot).
        dst = os.path.join(OUT, rel[:-3] + EXT)
        code = compile(data, dst, "exec", dont_inherit=True, optimize=0)
        _write_pyc(code, dst, int(os.stat(path).st_mtime), len(data))
    # the two numeric helpers of routers/tools.py, as oracle/ref_harness.py extracts them from the source tree
    tools = os.path.join(SRC, "app", "routers", "tools.py")
    tree = ast.parse(open(tools, encoding="utf-8").read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("_true_peak_dbfs", "_loudness_range_lu")]
    _write_pyc(compile(ast.Module(body=keep, type_ignores=[]), tools, "exec"), os.path.join(OUT, "app", "_tools_numeric" + EXT))
    if verbose:
        print(f"oracle/_ref: {len(mods)} reference modules byte-compiled into {OUT}")
    return OUT


if __name__ == "__main__":
    build()
