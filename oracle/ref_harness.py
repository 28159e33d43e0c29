"""ORACLE support (test infrastructure): import the UNMODIFIED reference in this container.

``/root/reference/backend/app/pipeline.py`` imports three I/O packages that are not installed
in this image (``pyloudnorm``, ``soundfile``, ``pydub``; ``backend/app/pipeline.py:13-15``).
This module injects small stand-ins for exactly those three names into ``sys.modules`` and then
imports ``app.pipeline`` / ``app.chain`` / ``app.routers.tools`` helpers from the read-only
reference tree.  No reference source is copied into the repository.  Where ``/root/reference`` does not exist
(the GPU box) the same modules are imported from ``oracle/_ref/backend``: compiled-code files (``*.bin``) that
``oracle/make_ref.py`` byte-compiled from the tree (the "built reference" of this pure-Python path).  Used by
``tests/golden/make_golden*.py``, by ``tests/test_oracle_vs_reference.py`` (pins ``oracle/chain.py`` against the
reference's own arithmetic) and by ``bench.py``'s CPU legs (``--impl reference``, ``cpu_baseline``).

Stand-ins:
* ``pyloudnorm.Meter``  -> ``oracle.bs1770.Meter`` (restated BS.1770-4 / pyloudnorm algorithm)
* ``soundfile.write/read`` -> canonical PCM WAV packing (``mm_b200.wavio``)
* ``pydub.AudioSegment``   -> empty class (only MP3/OPUS/AAC paths touch it)
``pedalboard`` is absent too, so ``apply_multiband_dynamics`` takes its numpy fallback branch
(``backend/app/pipeline.py:442-446``, ``:466-474``) -- the only branch that can be pinned here.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The source tree where it exists (the build container); else the byte-compiled copy oracle/make_ref.py built from it
# (oracle/_ref/backend: compiled-code files, git-ignored, shipped to the GPU box like a built .so).
PREBUILT_BACKEND = os.path.join(_REPO, "oracle", "_ref", "backend")
REFERENCE_BACKEND = os.environ.get("MM_REFERENCE_BACKEND") or (
    "/root/reference/backend" if os.path.isfile("/root/reference/backend/app/pipeline.py") else PREBUILT_BACKEND)


PREBUILT_EXT = ".bin"      # oracle/make_ref.py: .pyc layout under another name (the GPU-box snapshot drops *.pyc)


def _has(backend: str) -> bool:
    return any(os.path.isfile(os.path.join(backend, "app", "pipeline" + ext)) for ext in (".py", PREBUILT_EXT))


def _load_code(path: str):
    import marshal
    with open(path, "rb") as f:
        return marshal.loads(f.read()[16:])


class _PrebuiltFinder:
    """Imports ``app`` / ``app.*`` from the compiled-code files of oracle/_ref/backend (``<module>.bin``, ``<package>/__init__.bin``)."""

    def __init__(self, root: str):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "app" and not fullname.startswith("app."):
            return None
        from importlib.util import spec_from_loader
        rel = os.path.join(self.root, *fullname.split("."))
        if os.path.isfile(os.path.join(rel, "__init__" + PREBUILT_EXT)):
            spec = spec_from_loader(fullname, self, origin=os.path.join(rel, "__init__" + PREBUILT_EXT), is_package=True)
            spec.submodule_search_locations = [rel]
            return spec
        if os.path.isfile(rel + PREBUILT_EXT):
            return spec_from_loader(fullname, self, origin=rel + PREBUILT_EXT)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        origin = module.__spec__.origin
        module.__file__ = origin
        exec(_load_code(origin), module.__dict__)  # noqa: S102 - the reference's own byte-compiled module


def available() -> bool:
    return _has(REFERENCE_BACKEND)


def is_source_tree() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_BACKEND, "app", "pipeline.py"))


def _install_standins():
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    pkg = os.path.join(_REPO, "audio-mastering-web_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from oracle import bs1770
    from mm_b200 import wavio
    import numpy as np

    if "pyloudnorm" not in sys.modules:
        m = types.ModuleType("pyloudnorm")
        m.Meter = bs1770.Meter
        m.__standin__ = True
        sys.modules["pyloudnorm"] = m
    if "soundfile" not in sys.modules:
        m = types.ModuleType("soundfile")

        def write(file, data, samplerate, format=None, subtype=None):  # noqa: A002
            arr = np.asarray(data)
            if arr.dtype != np.int16:
                arr = np.clip(np.round(np.asarray(arr, dtype=np.float64) * 32768.0), -32768, 32767).astype(np.int16)
            file.write(wavio.pack_wav_pcm16(arr, samplerate))

        def read(file, dtype="float32", always_2d=False):
            x, sr = wavio.unpack_wav(file.read())
            if not always_2d and x.shape[1] == 1:
                x = x[:, 0]
            return x.astype(dtype), sr

        m.write, m.read = write, read
        m.__standin__ = True
        sys.modules["soundfile"] = m
    if "pydub" not in sys.modules:
        m = types.ModuleType("pydub")

        class AudioSegment:  # pragma: no cover - compressed formats are out of scope
            pass

        m.AudioSegment = AudioSegment
        m.__standin__ = True
        sys.modules["pydub"] = m


_cache = {}


def _numba_cache_guard():
    """Prebuilt copy only.  The reference decorates its two sample loops with numba.njit(cache=True) (pipeline.py:25-26); numba
    stats the file named in the function's code object for its cache index and raises at decoration time when it is missing.
    oracle/make_ref.py records the .pyc's own path under /root/repo there; if the copy was moved somewhere else, JIT-compile
    without the on-disk cache instead of failing (same machine code, compiled once per process)."""
    recorded = _load_code(os.path.join(REFERENCE_BACKEND, "app", "pipeline" + PREBUILT_EXT)).co_filename
    if os.path.exists(recorded):
        return
    try:
        import numba
    except ImportError:
        return
    orig = numba.njit

    def njit_nocache(*a, **k):
        k.pop("cache", None)
        return orig(*a, **k)

    numba.njit = njit_nocache


def load():
    """Return a namespace with ``pipeline``, ``chain`` (modules) and ``true_peak_dbfs``."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_BACKEND}")
    _install_standins()
    # numba's on-disk cache would try to write next to the read-only reference file
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/mm_numba_cache")
    if is_source_tree():
        if REFERENCE_BACKEND not in sys.path:
            sys.path.insert(0, REFERENCE_BACKEND)
    else:
        sys.meta_path.insert(0, _PrebuiltFinder(REFERENCE_BACKEND))
        _numba_cache_guard()
    pipeline = importlib.import_module("app.pipeline")
    chain = importlib.import_module("app.chain")
    ns = types.SimpleNamespace(pipeline=pipeline, chain=chain)

    # routers/tools.py drags FastAPI + helpers in; its two numeric helpers only need scipy.
    # Execute just those function bodies from the reference source text, unmodified.
    import ast
    import numpy as np
    from scipy.signal import resample_poly

    src_path = os.path.join(REFERENCE_BACKEND, "app", "routers", "tools.py")
    if os.path.isfile(src_path):
        tree = ast.parse(open(src_path, encoding="utf-8").read())
        keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("_true_peak_dbfs", "_loudness_range_lu")]
        code = compile(ast.Module(body=keep, type_ignores=[]), src_path, "exec")
    else:                                           # prebuilt copy: the same two definitions, compiled by oracle/make_ref.py
        code = _load_code(os.path.join(REFERENCE_BACKEND, "app", "_tools_numeric" + PREBUILT_EXT))
    g = {"np": np, "resample_poly": resample_poly, "compute_lufs_timeline": pipeline.compute_lufs_timeline}
    exec(code, g)  # noqa: S102 - reference code, run as the oracle's oracle
    ns.true_peak_dbfs = g["_true_peak_dbfs"]
    ns.loudness_range_lu = g["_loudness_range_lu"]
    _cache["ns"] = ns
    return ns
