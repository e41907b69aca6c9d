"""CPU oracle for the RAGMeUp dense-retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``ragmeup_b200/`` (the product) may
import this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only
as the checker or as the CPU arm that is timed beside the GPU number.

PARITY STATUS: **parity unpinned by the reference's own repository.**  The
reference (/root/reference, AI-Commandos/RAGMeUp @ 684d2938) holds no tests,
golden vectors or fixtures, and the arithmetic of this path lives in pinned
third-party wheels that are absent from this image (sentence-transformers
2.6.1, langchain-milvus 0.1.3 / milvus-lite 2.4.7, langchain-community
0.2.10, transformers 4.43.1; ``server/requirements.txt``).  The restatements
here follow the published algorithms of those versions
and the reference's call sites.  What *is* pinned:

* the transformer arithmetic (``bert_ref``) is cross-checked against the
  ``transformers`` ``BertModel`` / ``BertForSequenceClassification`` that is
  installed in this image (eager attention) by ``tests/golden/make_golden.py``;
  the resulting vectors are committed under ``tests/golden/``;
* ``ScoredCrossEncoderReranker.compress_documents`` semantics follow the only
  in-repo hot-path file, ``server/ScoredCrossEncoderReranker.py:25-45``.
"""
