"""TEST-ONLY environment for executing the UNMODIFIED reference callers (``/root/reference/release_server.py``,
``pipeline/causal_inference.py``) in the build container, on the CPU.  Nothing here is product code and nothing
here touches the reference sources: it only supplies what the container lacks.

* absent third-party packages: ``omegaconf``, ``easydict``, ``ftfy`` and the ``diffusers`` tree
  get inert stand-ins (the hot path uses none of them; they are imported by module headers and by the
  bidirectional / training pipelines that ``pipeline/__init__.py`` pulls in);
* ``wan`` is pre-registered as a bare package so ``wan/__init__.py``'s import of the I2V / T2V sampling scripts
  (out of scope, need the real diffusers schedulers) is skipped; ``wan.modules`` and everything below import normally;
* no GPU here: ``torch.cuda.current_device`` -> "cpu", ``torch.cuda.Stream`` / ``Event`` inert, ``Tensor.cuda()``
  identity, ``demo_utils.memory`` (builds a ``cuda:N`` device at import) replaced by its four names.
"""
from __future__ import annotations

import functools
import importlib.abc
import importlib.machinery
import inspect
import sys
import types
from pathlib import Path

import torch

REF = Path("/root/reference")
_REF_TOPLEVEL = ("wan", "pipeline", "utils", "demo_utils", "settings", "release_server", "v2v", "sample")


def available() -> bool:
    return (REF / "release_server.py").exists()


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    __setattr__ = dict.__setitem__


class _ConfigMixin:
    pass


def _register_to_config(init):
    """diffusers' decorator: records the constructor arguments as ``self.config`` (read by
    pipeline/causal_inference.py:289-291 as ``model.config.num_heads`` / ``.dim``)."""
    @functools.wraps(init)
    def wrapped(self, *a, **k):
        bound = inspect.signature(init).bind(self, *a, **k)
        bound.apply_defaults()
        self.config = types.SimpleNamespace(**{k_: v for k_, v in bound.arguments.items() if k_ != "self"})
        init(self, *a, **k)
    return wrapped


class _ModelMixin(torch.nn.Module):
    pass


_REAL = {"ConfigMixin": _ConfigMixin, "register_to_config": _register_to_config, "ModelMixin": _ModelMixin}


class _InertMeta(type):
    def __iter__(cls):
        return iter(())


class _Inert(metaclass=_InertMeta):
    """Placeholder for any name of an absent package: constructible, falsy, usable as a decorator."""

    def __init__(self, *a, **k):
        pass

    def __bool__(self):
        return False

    def __call__(self, *a, **k):
        return a[0] if len(a) == 1 and callable(a[0]) else self


class _AbsentLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []

        def getattr_(name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _REAL.get(name) or type(name, (_Inert,), {})
        m.__getattr__ = getattr_
        return m

    def exec_module(self, module):
        return None


class _AbsentFinder(importlib.abc.MetaPathFinder):
    ROOTS = ("diffusers",)

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, _AbsentLoader(), is_package=True)
        return None


class _NoStream:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def wait_stream(self, *a):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass


_saved = {}


def purge() -> None:
    for k in [k for k in sys.modules if k.split(".")[0] in _REF_TOPLEVEL or k.split(".")[0] in
              ("omegaconf", "easydict", "ftfy", "diffusers")]:
        del sys.modules[k]


def setup(dropin: bool = True):
    """Prepare the process, optionally install the drop-in finder, and return the imported unmodified
    ``release_server`` module.  With ``dropin=False`` the caller is expected to have prepared the reference
    modules itself (tests/golden/ref_shim.py) before calling."""
    if not available():
        raise RuntimeError("/root/reference is not available")
    if dropin:
        purge()
    if str(REF) not in sys.path:
        sys.path.append(str(REF))
    _stub("easydict", EasyDict=_EasyDict)
    _stub("ftfy", fix_text=lambda s: s)
    _stub("omegaconf", OmegaConf=type("OmegaConf", (), {}))
    if not any(isinstance(f, _AbsentFinder) for f in sys.meta_path):
        sys.meta_path.append(_AbsentFinder())
    if not torch.cuda.is_available():
        for name, val in (("current_device", lambda: "cpu"), ("Stream", _NoStream), ("Event", _NoStream)):
            _saved.setdefault(name, getattr(torch.cuda, name))
            setattr(torch.cuda, name, val)
        _saved.setdefault("Tensor.cuda", torch.Tensor.cuda)
        torch.Tensor.cuda = lambda self, *a, **k: self
        _stub("demo_utils.memory", gpu="cpu", get_cuda_free_memory_gb=lambda *_: 0.0, DynamicSwapInstaller=None,
              move_model_to_device_with_memory_preservation=lambda *a, **k: None)
    if dropin:
        pkg = types.ModuleType("wan")
        pkg.__path__ = [str(REF / "wan")]
        sys.modules["wan"] = pkg
        import realtime_video_b200.dropin as d
        d.install()
    import release_server
    return release_server


def teardown() -> None:
    import realtime_video_b200.dropin as d
    d.uninstall()
    for name, val in _saved.items():
        if name == "Tensor.cuda":
            torch.Tensor.cuda = val
        else:
            setattr(torch.cuda, name, val)
    _saved.clear()
    purge()
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _AbsentFinder)]
    if str(REF) in sys.path:
        sys.path.remove(str(REF))
