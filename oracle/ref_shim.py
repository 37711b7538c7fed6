"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference sources for golden-vector generation.

``import scdrs`` fails in this image (anndata, scanpy, skmisc, statsmodels ... are absent).
The two hot-path source files, /root/reference/scdrs/pp.py and /root/reference/scdrs/method.py,
nevertheless execute unmodified once those third-party modules are stubbed in ``sys.modules``.
This file does exactly that and nothing else; no reference code is copied.

The sources are looked for under ``$SCDRS_REFERENCE_ROOT``, ``/root/reference`` (the build container) and
``oracle/_ref`` (the git-ignored copy ``oracle/stage_ref.py`` stages so that it travels to the GPU box).  It is
used by ``tests/golden/make_golden.py`` to produce the committed fixtures, by the container-only tests that
compare ``oracle/scdrs_oracle.py`` with the live reference, and by ``bench.py``'s CPU legs (``cpu_baseline`` /
``--impl reference``), which time the reference's own ``score_cell`` where the staged copy exists and fall back to
the numpy port otherwise.  Nothing under ``scdrs_b200/`` imports this module.
"""
import importlib.util
import os
import sys
import types

def _find_root():
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("SCDRS_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "_ref")):
        if cand and os.path.exists(os.path.join(cand, "scdrs", "method.py")) and \
                os.path.exists(os.path.join(cand, "scdrs", "pp.py")):
            return cand
    return os.environ.get("SCDRS_REFERENCE_ROOT", "/root/reference")


REF_ROOT = _find_root()


def available():
    return os.path.exists(os.path.join(REF_ROOT, "scdrs", "method.py"))


def load(loess_impl=None):
    """Return (ref_pp, ref_method) modules executing the reference's own source files.

    loess_impl: optional callable ``loess(x, y, span=, degree=)`` standing in for
    ``skmisc.loess.loess`` (absent here); needed only for ``compute_stats``/``preprocess``.
    """
    if not available():
        raise RuntimeError("reference sources not found under %s" % REF_ROOT)
    if "scdrs_ref_shim.method" in sys.modules:
        pp, method = sys.modules["scdrs_ref_shim.pp"], sys.modules["scdrs_ref_shim.method"]
        if loess_impl is not None:
            pp.loess = loess_impl
        return pp, method

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from scdrs_b200.anndata_lite import AnnData, read_h5ad

    saved = {}

    def stub(name, **attrs):
        saved[name] = sys.modules.get(name)
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def _missing(*a, **k):
        raise NotImplementedError("stubbed third-party function")

    stub("scanpy")
    stub("anndata", AnnData=AnnData, read_h5ad=read_h5ad)
    stub("statsmodels")
    stub("statsmodels.stats")
    stub("statsmodels.stats.multitest", multipletests=_missing)
    stub("skmisc")
    stub("skmisc.loess", loess=loess_impl or _missing)
    pkg = stub("scdrs")
    pkg.__path__ = [os.path.join(REF_ROOT, "scdrs")]

    mods = {}
    for name in ("pp", "method"):
        spec = importlib.util.spec_from_file_location(
            "scdrs_ref_shim." + name, os.path.join(REF_ROOT, "scdrs", name + ".py")
        )
        m = importlib.util.module_from_spec(spec)
        sys.modules["scdrs_ref_shim." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    # undo the stubs so that nothing else in the process sees fake third-party modules
    for name, old in saved.items():
        if old is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = old
    return mods["pp"], mods["method"]
