"""Import the UNMODIFIED reference modules from /root/reference on CPU (build container only).

Used only by the golden-vector generators in this directory; nothing in ``tests/`` proper,
``bench.py`` or the product imports it (``/root/reference`` does not exist on the GPU box).

What the shim does:
  * registers empty package objects for ``wan``, ``wan.modules``, ``pipeline``, ``utils``,
    ``demo_utils`` whose ``__path__`` points at the reference dirs, which skips the
    ``__init__`` fan-out into bidirectional / training code;
  * stubs the absent third-party imports: easydict, diffusers.{configuration_utils,
    models.modeling_utils}, ftfy, demo_utils.memory (touches CUDA at import);
  * on CPU: ``torch.cuda.current_device`` -> 'cpu' (wan/modules/model.py:22 builds the
    sinusoid there) and forces the SDPA fallbacks (FLASH_ATTN_2_AVAILABLE = False);
  * optional fp32 mode: the two SDPA fallbacks hard-cast to bf16 (attention.py:204-212,
    model.py:217-223); ``patch_fp32_sdpa`` replaces them with an fp32 SDPA so an fp32 model runs;
  * optional ``frame_seqlen``: executes causal_model.py / causal_inference.py with the literal
    1560 replaced (the reference hard-codes 1560 tokens per frame, causal_model.py:192,351;
    pipeline/causal_inference.py:35) so small latent grids give self-consistent caches.  Cases
    generated that way are tagged "patched-constant" in the fixture metadata.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
import types
from pathlib import Path

import torch

REF = Path("/root/reference")


def _pkg(name: str, path: Path) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [str(path)]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install(frame_seqlen: int | None = None):
    """Returns a namespace with the reference modules."""
    if not REF.exists():
        raise RuntimeError("/root/reference is not available (golden generation runs in the build container)")
    for k in [k for k in sys.modules if k.split(".")[0] in ("wan", "pipeline", "utils", "demo_utils", "settings")]:
        del sys.modules[k]
    sys.path.insert(0, str(REF))
    _pkg("wan", REF / "wan")
    _pkg("wan.modules", REF / "wan" / "modules")
    _pkg("pipeline", REF / "pipeline")
    _pkg("utils", REF / "utils")
    _pkg("demo_utils", REF / "demo_utils")

    class EasyDict(dict):
        def __getattr__(self, k):
            return self[k]

        def __setattr__(self, k, v):
            self[k] = v

    _stub("easydict", EasyDict=EasyDict)

    class ConfigMixin:
        pass

    def register_to_config(init):
        import functools
        import inspect

        @functools.wraps(init)
        def wrapped(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            self.config = types.SimpleNamespace(**cfg)
            init(self, *args, **kwargs)
        return wrapped

    class ModelMixin(torch.nn.Module):
        pass

    _stub("diffusers")
    _stub("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    _stub("diffusers.models")
    _stub("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    _stub("ftfy", fix_text=lambda s: s)
    _stub("demo_utils.memory", gpu="cpu", get_cuda_free_memory_gb=lambda *_: 0.0,
          DynamicSwapInstaller=None, move_model_to_device_with_memory_preservation=lambda *a, **k: None)
    _stub("settings", MODEL_FOLDER="/nonexistent")

    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: "cpu"

    ns = types.SimpleNamespace()
    ns.attention = importlib.import_module("wan.modules.attention")
    ns.model = importlib.import_module("wan.modules.model")
    ns.attention.FLASH_ATTN_2_AVAILABLE = False
    ns.model.FLASH_ATTN_2_AVAILABLE = False
    ns.attention.FLASH_ATTN_3_AVAILABLE = False
    ns.model.FLASH_ATTN_3_AVAILABLE = False

    def _load_patched(modname: str, file: Path):
        src = file.read_text()
        if frame_seqlen is not None:
            src = "_FRAME_SEQLEN_ = %d\n" % frame_seqlen + src.replace("1560", "_FRAME_SEQLEN_")
        spec = importlib.util.spec_from_loader(modname, loader=None, origin=str(file))
        mod = importlib.util.module_from_spec(spec)
        mod.__file__ = str(file)
        sys.modules[modname] = mod
        exec(compile(src, str(file), "exec"), mod.__dict__)
        return mod

    ns.causal_model = _load_patched("wan.modules.causal_model", REF / "wan" / "modules" / "causal_model.py")
    ns.scheduler = importlib.import_module("utils.scheduler")
    # utils.wan_wrapper imports tokenizers / t5 / vae; they import fine with the stubs above
    ns.wan_wrapper = importlib.import_module("utils.wan_wrapper")
    ns.causal_inference = _load_patched("pipeline.causal_inference", REF / "pipeline" / "causal_inference.py")
    ns.vae = importlib.import_module("wan.modules.vae")
    ns.vae_block3 = importlib.import_module("demo_utils.vae_block3")
    ns.vae_single = importlib.import_module("demo_utils.vae")
    return ns


def patch_fp32_sdpa(ns) -> None:
    """fp32 SDPA for the two bf16-casting fallbacks (attention.py:204-212, model.py:217-223)."""
    import torch.nn.functional as F

    def attention_fp32(q, k, v, *a, **kw):
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2).contiguous()

    ns.causal_model.attention = attention_fp32

    def cross_forward(self, x, context, context_lens, crossattn_cache=None):
        b, n, d = x.size(0), self.num_heads, self.head_dim
        q = self.norm_q(self.q(x)).view(b, -1, n, d)
        if crossattn_cache is not None:
            if not crossattn_cache["is_init"]:
                crossattn_cache["is_init"] = True
                k = self.norm_k(self.k(context)).view(b, -1, n, d)
                v = self.v(context).view(b, -1, n, d)
                crossattn_cache["k"] = k
                crossattn_cache["v"] = v
            else:
                k = crossattn_cache["k"]
                v = crossattn_cache["v"]
        else:
            k = self.norm_k(self.k(context)).view(b, -1, n, d)
            v = self.v(context).view(b, -1, n, d)
        x = attention_fp32(q, k, v).flatten(2)
        return self.o(x)

    # same statements as model.py:171-228 with the dtype-preserving attention call
    ns.model.WanT2VCrossAttention.forward = cross_forward


def patch_flex_dense(ns) -> None:
    """CPU stand-in for the compiled FlexAttention call (causal_model.py:339-348): SDPA with the
    dense mask of the reference's own get_sdpa_mask (causal_model.py:41-106).  The caller sets
    ``model.block_mask`` to that dense bool mask."""
    import torch.nn.functional as F

    def flex(query, key, value, block_mask=None, kernel_options=None):
        return F.scaled_dot_product_attention(query, key, value, attn_mask=block_mask)

    ns.causal_model.flex_attention = flex
