"""Document / base-class shims.

The reference builds on LangChain (`langchain_core.documents.Document`,
`BaseDocumentCompressor`, `Embeddings`, `VectorStore`, `BaseRetriever`; imports at
``server/ScoredCrossEncoderReranker.py:6-9`` and ``server/RAGHelper.py:10-33``).  When
``langchain_core`` is importable the real classes are used so the objects plug straight into
LangChain chains; when it is absent (this build image) minimal stand-ins with the same duck-type
surface are used instead.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

try:  # pragma: no cover - not installed in the build image
    from langchain_core.documents import Document  # type: ignore
    HAVE_LANGCHAIN = True
except Exception:  # pragma: no cover - exercised in the build image
    HAVE_LANGCHAIN = False

    class Document:  # type: ignore[no-redef]
        """Minimal stand-in for langchain_core.documents.Document."""

        def __init__(self, page_content: str, metadata: Optional[Dict[str, Any]] = None, **kwargs: Any):
            self.page_content = page_content
            self.metadata = dict(metadata) if metadata is not None else {}
            self.id = kwargs.get("id")

        def copy(self, update: Optional[Dict[str, Any]] = None, **_: Any) -> "Document":
            d = Document(self.page_content, dict(self.metadata), id=self.id)
            for k, v in (update or {}).items():
                setattr(d, k, v)
            return d

        model_copy = copy

        def __eq__(self, other: object) -> bool:
            return (isinstance(other, Document) and other.page_content == self.page_content
                    and other.metadata == self.metadata)

        def __repr__(self) -> str:
            return f"Document(page_content={self.page_content!r}, metadata={self.metadata!r})"


class Runnable:
    """The slice of LangChain's Runnable protocol the reference's chains use on retrievers:
    ``retriever.invoke(q)`` and ``retriever | fn`` (``server/RAGHelper_local.py:157-159,255-256``)."""

    def invoke(self, input: Any, config: Any = None, **kwargs: Any) -> Any:  # noqa: A002
        raise NotImplementedError

    def __or__(self, other: Any) -> "Runnable":
        return _Sequence(self, other)


class _Sequence(Runnable):
    def __init__(self, first: Any, second: Any):
        self.first, self.second = first, second

    def invoke(self, input: Any, config: Any = None, **kwargs: Any) -> Any:  # noqa: A002
        x = self.first.invoke(input, config, **kwargs) if hasattr(self.first, "invoke") else self.first(input)
        return self.second.invoke(x, config, **kwargs) if hasattr(self.second, "invoke") else self.second(x)


def copy_document(doc: Any, metadata: Dict[str, Any]) -> Any:
    """``doc.copy(update={"metadata": ...})`` across pydantic v1/v2 Documents and the stand-in."""
    if hasattr(doc, "model_copy") and HAVE_LANGCHAIN:
        return doc.model_copy(update={"metadata": metadata})
    return doc.copy(update={"metadata": metadata})
