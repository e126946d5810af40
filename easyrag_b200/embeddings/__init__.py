from .gte_embeddings import GTEEmbedding
from .hf_embeddings import HuggingFaceEmbedding

__all__ = ["GTEEmbedding", "HuggingFaceEmbedding"]
