"""
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the UNMODIFIED reference (read-only at /root/reference) in-process so that golden
vectors can be generated from it (tests/golden/make_golden.py) and the torch restatement in
oracle/pixelnerf_oracle.py can be pinned against it.  The reference's live
``src/model/models.py`` cannot be constructed with any shipped conf; the
functional model is ``src/model/models.py.backup2``, so ``model.models`` is
loaded from that file, in place, without copying it.

/root/reference exists only in the build container.  On the GPU box the byte-identical staged copy under
the git-ignored ``baseline/_ref`` (oracle/stage_reference.py) is used instead: that is what
``bench.py --impl reference`` / ``cpu_baseline`` time and what the caller tests execute.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.environ.get("PIXELNERF_REFERENCE", "/root/reference"),
               os.path.join(os.path.dirname(_HERE), "baseline", "_ref"))


def _find_root():
    for r in _CANDIDATES:
        if os.path.isfile(os.path.join(r, "src", "model", "models.py.backup2")):
            return r
    return None


REF_ROOT = _find_root() or _CANDIDATES[0]


def available():
    return _find_root() is not None


def load():
    """Returns (make_model, NeRFRenderer, util_module) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "model" in sys.modules and getattr(sys.modules["model"], "_pnr_ref", False):
        import render as _r
        import util as _u

        return sys.modules["model"].make_model, _r.NeRFRenderer, _u
    repo = os.path.dirname(_HERE)
    for p in (repo, os.path.join(_HERE, "_shims"), os.path.join(REF_ROOT, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    src_model = os.path.join(REF_ROOT, "src", "model")
    pkg = types.ModuleType("model")
    pkg.__path__ = [src_model]
    pkg.__package__ = "model"
    pkg._pnr_ref = True
    sys.modules["model"] = pkg
    loader = importlib.machinery.SourceFileLoader(
        "model.models", os.path.join(src_model, "models.py.backup2")
    )
    spec = importlib.util.spec_from_loader("model.models", loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["model.models"] = mod
    loader.exec_module(mod)
    with open(os.path.join(src_model, "__init__.py"), "r", encoding="utf-8") as fh:
        code = compile(fh.read(), os.path.join(src_model, "__init__.py"), "exec")
    exec(code, pkg.__dict__)
    import render as _r
    import util as _u

    return pkg.make_model, _r.NeRFRenderer, _u
