"""Stub loader for the UNMODIFIED reference package (authoring container only).

TEST INFRASTRUCTURE - not part of the product.  `/root/reference` does not
exist on the GPU box, so nothing under `tests/ -m gpu`, `smoke()` or
`bench.py` may import this module.  It is used only by
`oracle/gen_golden.py` (to write `tests/golden/*.npz`) and by the CPU-only
test that cross-checks `oracle/cwt_oracle.py` against the live reference when
the reference tree is present.

The reference cannot be imported as-is: `ninwavelets/base.py:2,3,7` and
`ninwavelets/wavelets.py:4` import cupy / matplotlib / mpl_toolkits
unconditionally and none of them is installed.  We pre-seed `sys.modules`
with inert stand-ins (the CPU path never touches them), register an empty
package whose `__path__` points at the reference and exec `base.py`,
`wavelets.py`, `mneutils.py` from where they lie.  `__init__.py` / `init.py`
are never executed (the latter would chdir and spawn a build).
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NINWAVELETS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ninwavelets", "base.py"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def load():
    """Return a namespace with the reference's base / wavelets / mneutils modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference dir is read-only
    if "cupy" not in sys.modules:
        _stub("cupy", ndarray=type("ndarray", (), {}))
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot", figure=object, Axes=object)
        mpl.pyplot = plt
    if "mpl_toolkits" not in sys.modules:
        tk = _stub("mpl_toolkits")
        ax = _stub("mpl_toolkits.axes_grid1", make_axes_locatable=lambda *a, **k: None)
        tk.axes_grid1 = ax
    pkg_name = "ninwavelets"
    pkg_dir = os.path.join(REFERENCE_ROOT, "ninwavelets")
    if pkg_name not in sys.modules or not hasattr(sys.modules[pkg_name], "_nw_ref_loaded"):
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [pkg_dir]
        pkg._nw_ref_loaded = True
        sys.modules[pkg_name] = pkg
        for sub in ("base", "wavelets", "mneutils"):
            full = "%s.%s" % (pkg_name, sub)
            spec = importlib.util.spec_from_file_location(full, os.path.join(pkg_dir, sub + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
            setattr(pkg, sub, mod)
    pkg = sys.modules[pkg_name]
    return types.SimpleNamespace(base=pkg.base, wavelets=pkg.wavelets, mneutils=pkg.mneutils)
