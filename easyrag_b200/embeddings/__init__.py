"""Drop-in embedding classes (import surface of ``easyrag.custom.embeddings``, pipeline.py:19).

The two classes are resolved on first attribute access, so importing the package does not pull torch-side
modules until one of them is actually used.
"""
import importlib

_EXPORTS = {"GTEEmbedding": "gte_embeddings", "HuggingFaceEmbedding": "hf_embeddings"}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    module = _EXPORTS.get(name)
    if module is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    value = getattr(importlib.import_module(f"{__name__}.{module}"), name)
    globals()[name] = value
    return value
