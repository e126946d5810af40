"""Drop-in for the reference's ``easyrag/custom/embeddings`` package (imported at pipeline.py:15)."""
from easyrag_b200.embeddings import GTEEmbedding, HuggingFaceEmbedding          # noqa: F401
