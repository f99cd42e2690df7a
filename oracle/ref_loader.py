"""Import the *unmodified* reference (``/root/reference``) on CPU.  TEST INFRASTRUCTURE ONLY.

The reference factories do ``from yacs.config import CfgNode`` at call time
(monoport/lib/modeling/MonoPortNet.py:164,188; backbones/HGFilters.py:208;
normalizers/DepthNormalizer.py:37).  yacs is not installed in this image, so a minimal
in-memory attribute-dict stand-in is registered under ``sys.modules['yacs.config']``
before the reference is imported.  Nothing of the reference is copied.

``/root/reference`` only exists in the build container; on the GPU box the committed
golden vectors under ``tests/golden/`` stand in for it.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MONOPORT_REFERENCE", "/root/reference")


class _CfgNode(dict):
    """6-line stand-in for yacs.config.CfgNode: attribute get/set on a dict + clone()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "monoport", "lib", "modeling"))


def _install_yacs_shim():
    if "yacs.config" in sys.modules:
        return
    try:
        import yacs.config  # noqa: F401
        return
    except Exception:
        pass
    yacs = types.ModuleType("yacs")
    cfg = types.ModuleType("yacs.config")
    cfg.CfgNode = _CfgNode
    yacs.config = cfg
    sys.modules["yacs"] = yacs
    sys.modules["yacs.config"] = cfg


def _load_by_path(name, path):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_reference():
    """Returns a namespace with the reference's PIFuNetG, PIFuNetC, orthogonal, perspective, index,
    pifu_calib, forward_vertices -- loaded from the reference tree under private module names so
    that they can never shadow (or be shadowed by) this repo's ``monoport`` drop-in package."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    _install_yacs_shim()
    import importlib.util
    base = os.path.join(REF_ROOT, "monoport", "lib", "modeling")
    # Build a private package tree  _mpref.{geometry,normalizers,backbones.{HGFilters,ResBlkFilters},heads}
    pkg = types.ModuleType("_mpref")
    pkg.__path__ = [base]
    sys.modules["_mpref"] = pkg

    def sub(name, relpath, is_pkg=False):
        full = "_mpref." + name
        path = os.path.join(base, relpath)
        spec = importlib.util.spec_from_file_location(
            full, path, submodule_search_locations=[os.path.dirname(path)] if is_pkg else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        return mod

    geometry = sub("geometry", "geometry.py")
    normalizers = sub("normalizers", "normalizers/__init__.py", True)
    heads = sub("heads", "heads/__init__.py", True)
    # backbones/__init__.py imports Yolov4/HRNet too (pure torch) -- import only the two on the path
    bb = types.ModuleType("_mpref.backbones")
    bb.__path__ = [os.path.join(base, "backbones")]
    sys.modules["_mpref.backbones"] = bb
    hg = sub("backbones.HGFilters", "backbones/HGFilters.py")
    rb = sub("backbones.ResBlkFilters", "backbones/ResBlkFilters.py")
    bb.HGFilter, bb.PIFuHGFilters = hg.HGFilter, hg.PIFuHGFilters
    bb.ResnetFilter, bb.PIFuResBlkFilters = rb.ResnetFilter, rb.PIFuResBlkFilters
    bb.__all__ = ["HGFilter", "PIFuHGFilters", "ResnetFilter", "PIFuResBlkFilters"]
    mpn = sub("MonoPortNet", "MonoPortNet.py")
    recon = _load_by_path("_mpref_recon", os.path.join(REF_ROOT, "RTL", "recon.py"))

    ns = types.SimpleNamespace(
        MonoPortNet=mpn.MonoPortNet, PIFuNetG=mpn.PIFuNetG, PIFuNetC=mpn.PIFuNetC,
        orthogonal=geometry.orthogonal, perspective=geometry.perspective, index=geometry.index,
        pifu_calib=recon.pifu_calib, forward_vertices=recon.forward_vertices,
        SurfaceClassifier=heads.SurfaceClassifier)
    _cache["ns"] = ns
    return ns
