"""Import the UNMODIFIED reference from /root/reference (build container only) -- TEST INFRASTRUCTURE.

/root/reference does not exist on the GPU box; nothing that runs there may call this.  It is used by
``oracle/make_golden.py`` (fixture generation), by CPU tests that skip when the tree is absent, and by
``bench.py --impl reference`` when available.

Work-arounds for upstream import defects:
  * r2plus1d.py:10 does ``import resnet3D`` (absolute)      -> alias ``sys.modules['resnet3D']``
  * trn.py:8 does ``import pretrainedmodels`` (missing dep)  -> alias to the ``pretorched`` package
The reference tree is read-only, so bytecode writing is disabled around the import.
"""
import importlib
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
# the source tree (build container) or, where that does not exist (GPU box), its byte-compiled image under oracle/_ref
# (oracle/build_ref.py): the same unmodified modules either way
_CANDIDATES = [os.environ.get("B2_REFERENCE_ROOT", "/root/reference"), os.path.join(_HERE, "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if os.path.isdir(os.path.join(c, "pretorched"))), _CANDIDATES[0])


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pretorched"))


def is_source_tree():
    """True when the unmodified SOURCE tree is mounted (fixture generation needs data/ files that oracle/_ref does not carry)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pretorched", "__init__.py"))


def load():
    """Returns the reference's top-level ``pretorched`` package."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return importlib.import_module("pretorched")
    finally:
        sys.dont_write_bytecode = old


def load_r2plus1d():
    """pretorched.models.r2plus1d.  NOTE: R2Plus1D.forward breaks once any resnet3d*
    factory has run in the process (class-level patch by modify_resnets); build R(2+1)D first or use
    ``ResNet3D.forward`` explicitly as ``reference_forward`` below does."""
    pt = load()
    sys.modules.setdefault("resnet3D", pt.models.resnet3D)
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        return importlib.import_module("pretorched.models.r2plus1d")
    finally:
        sys.dont_write_bytecode = old


def load_preact():
    """pretorched.models.pre_act_resnet3D (same absolute ``import resnet3D`` defect as r2plus1d.py: pre_act_resnet3D.py:8)."""
    pt = load()
    sys.modules.setdefault("resnet3D", pt.models.resnet3D)
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        return importlib.import_module("pretorched.models.pre_act_resnet3D")
    finally:
        sys.dont_write_bytecode = old


def load_transforms():
    """pretorched.transforms.utils.  transforms/utils.py:7 imports ``munchify`` from the ``munch`` package, which is not
    installed here; it is only used to turn a settings dict into an attribute bag (utils.py:39-40), so a two-line stand-in
    module is registered before the import.  Nothing else is patched."""
    import types
    load()
    if "munch" not in sys.modules:
        stub = types.ModuleType("munch")
        stub.munchify = lambda d: types.SimpleNamespace(**d)
        sys.modules["munch"] = stub
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        return importlib.import_module("pretorched.transforms.utils")
    finally:
        sys.dont_write_bytecode = old


def load_trn():
    pt = load()
    sys.modules.setdefault("pretrainedmodels", pt)
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return importlib.import_module("pretorched.models.trn")
    finally:
        sys.dont_write_bytecode = old


def build(arch, **kwargs):
    """Instantiate a reference model with ``pretrained=None`` (URLs are unreachable offline)."""
    pt = load()
    if arch.startswith("r2plus1d"):
        return getattr(load_r2plus1d(), arch)(**kwargs)
    if arch.startswith("preact_"):
        return getattr(load_preact(), arch)(**kwargs)
    if arch.startswith("resnext3d"):                     # resnext3D.py:224-252: plain **kwargs factories, no `pretrained`
        return getattr(pt, arch)(**kwargs)
    if arch.startswith("nonlocal"):
        return getattr(pt.models.nonlocalnet, arch)(pretrained=None, **kwargs)
    return getattr(pt, arch)(pretrained=None, **kwargs)
