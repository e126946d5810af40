"""CPU: the ``shim/easyrag`` overlay resolves ``easyrag.custom.retrievers`` / ``.embeddings`` to easyrag_b200 while the
rest of the reference's ``easyrag`` package (here: a stand-in tree with the reference's relative imports,
pipeline.py:15-26) keeps resolving from its own directory."""
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_pipeline_imports_resolve_unchanged_through_the_overlay(tmp_path):
    ref = tmp_path / "src" / "easyrag"
    (ref / "custom" / "embeddings").mkdir(parents=True)
    (ref / "pipeline").mkdir()
    (ref / "utils").mkdir()
    for d in (ref, ref / "custom", ref / "pipeline", ref / "utils"):
        (d / "__init__.py").write_text("")
    # what the reference would provide itself: these must NOT be picked for retrievers / embeddings
    (ref / "custom" / "retrievers.py").write_text("WHO = 'reference'\nclass BM25Retriever: pass\n")
    (ref / "custom" / "embeddings" / "__init__.py").write_text("WHO = 'reference'\n")
    (ref / "custom" / "rerankers.py").write_text("WHO = 'reference'\nclass LLMRerank: pass\n")
    (ref / "custom" / "template.py").write_text("QA_TEMPLATE = 'qa'\n")
    (ref / "pipeline" / "ingestion.py").write_text("def get_node_content(node, embed_type=0):\n    return 'ref'\n")
    (ref / "utils" / "llm_utils.py").write_text("def local_llm_generate():\n    return 1\n")
    # the import block of pipeline.py:15-26, verbatim relative imports
    (ref / "pipeline" / "pipeline.py").write_text(textwrap.dedent('''
        from ..custom.embeddings import GTEEmbedding, HuggingFaceEmbedding
        from .ingestion import get_node_content as _get_node_content
        from ..custom.rerankers import LLMRerank
        from ..custom.retrievers import QdrantRetriever, BM25Retriever, HybridRetriever
        from ..custom.template import QA_TEMPLATE
        from ..utils.llm_utils import local_llm_generate as _local_llm_generate
    '''))
    code = textwrap.dedent(f'''
        import sys
        sys.path[:0] = [{str(ROOT / "shim")!r}, {str(tmp_path / "src")!r}, {str(ROOT)!r}]
        import easyrag.pipeline.pipeline as p
        import easyrag_b200.retrievers as ours
        import easyrag.custom.rerankers as rr
        assert p.BM25Retriever is ours.BM25Retriever and p.HybridRetriever is ours.HybridRetriever
        assert p.QdrantRetriever is ours.QdrantRetriever
        assert p.GTEEmbedding.__module__ == "easyrag_b200.embeddings.gte_embeddings"
        assert p.HuggingFaceEmbedding.__module__ == "easyrag_b200.embeddings.hf_embeddings"
        assert rr.WHO == "reference" and p.QA_TEMPLATE == "qa" and p._get_node_content(None) == "ref"
        assert callable(p.HybridRetriever.fusion) and callable(p.HybridRetriever.reciprocal_rank_fusion)
        print("overlay-ok")
    ''')
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "overlay-ok" in r.stdout, r.stderr[-2000:]


def test_embedding_classes_construct_on_a_pydantic_base_embedding(tmp_path):
    """With llama_index installed ``BaseEmbedding`` is a pydantic model: only declared fields / private attributes
    may be set.  A minimal stand-in ``llama_index`` with pydantic classes exercises that path."""
    li = tmp_path / "llama_index" / "core"
    (li / "base" / "embeddings").mkdir(parents=True)
    (li / "bridge").mkdir()
    for d in (tmp_path / "llama_index", li, li / "base", li / "base" / "embeddings", li / "bridge"):
        (d / "__init__.py").write_text("")
    (li / "__init__.py").write_text("class QueryBundle:\n    def __init__(self, query_str, custom_embedding_strs=None, embedding=None):\n"
                                    "        self.query_str, self.custom_embedding_strs, self.embedding = query_str, custom_embedding_strs, embedding\n")
    (li / "bridge" / "pydantic.py").write_text("from pydantic import BaseModel, Field, PrivateAttr, ConfigDict\n")
    (li / "base" / "base_retriever.py").write_text(textwrap.dedent('''
        class BaseRetriever:
            def __init__(self, callback_manager=None, object_map=None, objects=None, verbose=False):
                self.callback_manager = callback_manager
    '''))
    (li / "base" / "embeddings" / "base.py").write_text(textwrap.dedent('''
        from typing import Any, List, Optional
        from pydantic import BaseModel, ConfigDict, Field
        class BaseEmbedding(BaseModel):
            model_config = ConfigDict(arbitrary_types_allowed=True, protected_namespaces=())
            model_name: str = Field(default="unknown")
            embed_batch_size: int = Field(default=10, gt=0)
            callback_manager: Optional[Any] = Field(default=None, exclude=True)
            def get_text_embedding_batch(self, texts: List[str], **kw):
                out = []
                for i in range(0, len(texts), self.embed_batch_size):
                    out.extend(self._get_text_embeddings(texts[i:i + self.embed_batch_size]))
                return out
    '''))
    (li / "schema.py").write_text(textwrap.dedent('''
        class BaseNode: pass
        class TextNode(BaseNode):
            def __init__(self, text="", id_=None, metadata=None, embedding=None):
                self.text, self.id_, self.metadata, self.embedding = text, id_ or "n", metadata or {}, embedding
            node_id = property(lambda self: self.id_)
            def get_content(self, metadata_mode=None): return self.text
        class NodeWithScore:
            def __init__(self, node, score=None): self.node, self.score = node, score
    '''))
    code = textwrap.dedent(f'''
        import sys
        sys.path[:0] = [{str(tmp_path)!r}, {str(ROOT)!r}]
        from types import SimpleNamespace as NS
        import pydantic
        from easyrag_b200 import schema
        assert schema.HAVE_LLAMA_INDEX and issubclass(schema.BaseEmbedding, pydantic.BaseModel)
        from easyrag_b200.embeddings import GTEEmbedding, HuggingFaceEmbedding
        enc = NS(cfg=NS(max_position_embeddings=512), device="cpu")
        hf = HuggingFaceEmbedding(model_name="bge-large-zh", embed_batch_size=128, embed_type=1, encoder=enc,
                                  hf_tokenizer=object(), max_length=256, cache_folder="/tmp/x")
        assert hf.max_length == 256 and hf.normalize is True and hf.embed_batch_size == 128 and hf.cache_folder == "/tmp/x"
        assert hf._embed_type == 1 and hf._model is enc and hf._prompts["query"].startswith("为这个句子")
        assert hf.model_name == "bge-large-zh" and HuggingFaceEmbedding.class_name() == "HuggingFaceEmbedding"
        try:
            HuggingFaceEmbedding(model_name="m", encoder=enc, hf_tokenizer=object(), max_length=4096)
            raise SystemExit("max_length beyond the position table must be rejected")
        except ValueError:
            pass
        try:
            HuggingFaceEmbedding(model_name="m", encoder=enc, hf_tokenizer=object(), pooling="cls")
            raise SystemExit("deprecated argument accepted")
        except ValueError:
            pass
        gte = GTEEmbedding(model_name="gte-qwen2", embed_batch_size=64, embed_type=2, encoder=enc, tokenizer=object())
        assert gte._embed_type == 2 and gte._model is enc and gte.embed_batch_size == 64 and gte._device == "cpu"
        assert "Instruct: Given a web search query" in gte.get_detailed_instruct("q")
        print("pydantic-ok")
    ''')
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "pydantic-ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
