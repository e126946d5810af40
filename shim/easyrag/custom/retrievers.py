"""Drop-in for the reference's ``easyrag/custom/retrievers.py`` (imported at pipeline.py:19)."""
from easyrag_b200.retrievers import *                                            # noqa: F401,F403
from easyrag_b200.retrievers import (B200VectorStore, BM25Retriever, HybridRetriever, QdrantRetriever,      # noqa: F401
                                     get_node_content, tokenize_and_remove_stopwords)
