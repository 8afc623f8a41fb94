"""Drop-in for the reference's ``pipeline`` package: only the hot-path pipeline is provided
(reference pipeline/__init__.py also exports training / bidirectional pipelines, which are out
of scope — SURVEY.md §2 row 14)."""
from .causal_inference import CausalInferencePipeline

__all__ = ["CausalInferencePipeline"]
