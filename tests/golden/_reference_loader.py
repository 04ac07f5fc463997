"""Import the reference (buqeye/gsum, read-only at /root/reference) by path for golden-vector generation.

Build-container only: /root/reference does not exist on the GPU box, so nothing in the test-suite
imports this module — only `make_golden.py` (run by hand here; its outputs are the committed .npz).
Import-only stubs for packages missing from this image (docrep, statsmodels, seaborn, matplotlib,
cycler); none of them touch arithmetic on the df=None path.
"""
import importlib.util
import sys
import types

REF = "/root/reference/gsum"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _DocstringProcessor:  # docrep is docstring templating only
    def get_sectionsf(self, *a, **k):
        return lambda f: f

    def dedent(self, f):
        return f


def load_reference():
    if "gsum.models" in sys.modules and getattr(sys.modules["gsum"], "__graft_ref__", False):
        g = sys.modules
        return g["gsum.helpers"], g["gsum.models"], g["gsum.datasets"], g["gsum.diagnostics"]
    _stub("docrep", DocstringProcessor=_DocstringProcessor)
    for n in ("statsmodels", "statsmodels.sandbox", "statsmodels.sandbox.distributions"):
        _stub(n)
    _stub("statsmodels.sandbox.distributions.mv_normal", MVT=object)
    _stub("seaborn")
    _stub("cycler", cycler=lambda *a, **k: None)
    mpl = _stub("matplotlib", rcParams={"axes.prop_cycle": []})
    mpl.pyplot = _stub("matplotlib.pyplot")
    pkg = types.ModuleType("gsum")
    pkg.__path__ = [REF]
    pkg.__graft_ref__ = True
    sys.modules["gsum"] = pkg

    def _load(name):
        spec = importlib.util.spec_from_file_location("gsum." + name, f"{REF}/{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules["gsum." + name] = m
        spec.loader.exec_module(m)
        return m

    helpers = _load("helpers")
    for n in helpers.__all__:
        setattr(pkg, n, getattr(helpers, n))
    models, datasets, diagnostics = _load("models"), _load("datasets"), _load("diagnostics")
    return helpers, models, datasets, diagnostics
