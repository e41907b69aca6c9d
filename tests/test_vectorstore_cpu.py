"""CPU: host logic of the vector-store drop-ins with the CUDA index replaced by a numpy stand-in —
persistence at ``connection_args["uri"]`` / ``drop_old`` (server/RAGHelper.py:388-394, .env.template:33),
PGVector's upsert-on-id, locking between add and search, and the LangChain re-basing of install.py against a
stub ``langchain_core`` whose ``VectorStore.embeddings`` is a read-only property (as the real one's is)."""
import os
import sys
import threading
import types

import numpy as np
import pytest

from oracle import flat_ref
from ragmeup_b200 import vectorstore as vs
from ragmeup_b200.documents import Document


class NumpyIndex:
    """FlatIndex look-alike (the product index needs a GPU; only the host logic is under test here)."""

    def __init__(self, dim, metric, device=None):
        self.dim, self.metric = dim, metric
        self.x = np.zeros((0, dim), np.float32)

    def __len__(self):
        return len(self.x)

    def add(self, v):
        self.x = np.concatenate([self.x, np.asarray(v, np.float32)])

    def set_rows(self, rows, v):
        self.x[np.asarray(rows)] = np.asarray(v, np.float32)


class Emb:
    def embed_documents(self, texts):
        return [[float(len(t)), float(sum(map(ord, t)) % 7), 1.0] for t in texts]

    def embed_query(self, t):
        return self.embed_documents([t])[0]


@pytest.fixture()
def numpy_index(monkeypatch):
    monkeypatch.setattr(vs, "FlatIndex", NumpyIndex)


def _docs(n, off=0):
    return [Document(f"text {i + off}" * (1 + i % 3), {"source": f"s{i + off}"}) for i in range(n)]


def test_milvus_uri_persists_and_drop_old_clears(numpy_index, tmp_path):
    uri = str(tmp_path / "data.db")
    a = vs.Milvus.from_documents([], Emb(), drop_old=True, connection_args={"uri": uri}, collection_name="LangChainCollection")
    assert len(a) == 0
    a.add_documents(_docs(5), ids=[f"id{i}" for i in range(5)])
    a.add_documents(_docs(3, 5), ids=[f"id{i}" for i in range(5, 8)])
    # a new process with vector_store_initial_load=False expects the rows to be there (RAGHelper.py:406-409)
    b = vs.Milvus.from_documents([], Emb(), drop_old=False, connection_args={"uri": uri}, collection_name="LangChainCollection")
    assert len(b) == 8 and b._pks == a._pks and b._texts == a._texts and b._metas == a._metas
    assert np.array_equal(b.index.x, a.index.x)
    b.add_documents(_docs(2, 8), ids=["id8", "id9"])
    c = vs.Milvus(Emb(), connection_args={"uri": uri})
    assert len(c) == 10
    other = vs.Milvus(Emb(), connection_args={"uri": uri}, collection_name="another")
    assert len(other) == 0                                   # collections are separate
    d = vs.Milvus.from_documents([], Emb(), drop_old=True, connection_args={"uri": uri})
    assert len(d) == 0
    assert len(vs.Milvus(Emb(), connection_args={"uri": uri})) == 0
    mem = vs.Milvus(Emb())                                   # no uri: memory only, nothing written
    mem.add_documents(_docs(2))
    assert mem._storage_dir is None


def test_pgvector_upserts_on_id_milvus_does_not(numpy_index, tmp_path, monkeypatch):
    monkeypatch.setenv("RMU_STORE_DIR", str(tmp_path / "pg"))
    pg = vs.PGVector(embeddings=Emb(), collection_name="c", connection="postgresql://x", use_jsonb=True)
    pg.add_documents(_docs(4), ids=["a", "b", "c", "d"])
    pg.add_documents([Document("replaced", {"source": "new"})], ids=["b"])
    assert len(pg) == 4 and pg._texts[1] == "replaced" and pg._metas[1] == {"source": "new"}
    assert np.allclose(pg.index.x[1], Emb().embed_documents(["replaced"])[0])
    again = vs.PGVector(embeddings=Emb(), collection_name="c", connection="postgresql://x")
    assert len(again) == 4 and again._texts[1] == "replaced"          # the upsert survives a reload
    mv = vs.Milvus(Emb())
    mv.add_documents(_docs(2), ids=["a", "b"])
    mv.add_documents(_docs(1), ids=["b"])
    assert len(mv) == 3                                       # Milvus inserts a second row with the same pk


def test_saved_metric_is_checked(numpy_index, tmp_path):
    m = vs.Milvus(Emb())
    m.add_documents(_docs(3))
    m.index.data = lambda: types.SimpleNamespace(cpu=lambda: types.SimpleNamespace(numpy=lambda: m.index.x))
    m.save(str(tmp_path / "one.npz"))
    assert len(vs.Milvus.load(str(tmp_path / "one.npz"), Emb())) == 3
    with pytest.raises(ValueError):
        vs.PGVector.load(str(tmp_path / "one.npz"), Emb())   # an L2 collection is not a cosine collection


def test_search_waits_for_a_concurrent_insert(numpy_index):
    """a row the index already returns but the host tables do not hold yet must not raise:
    _docs takes the store lock, which add_embeddings holds across index.add and the table append"""
    st = vs.Milvus(Emb())
    st.add_documents(_docs(2))
    in_add, go = threading.Event(), threading.Event()
    real_add = st.index.add

    def slow_add(v):
        real_add(v)
        in_add.set()
        go.wait(5)
    st.index.add = slow_add
    t = threading.Thread(target=lambda: st.add_documents(_docs(1, 2)))
    t.start()
    assert in_add.wait(5)
    out = []
    reader = threading.Thread(target=lambda: out.append(st._docs([2])))       # the new row, visible in the index only
    reader.start()
    reader.join(0.2)
    assert reader.is_alive()                                  # blocked on the lock, not an IndexError
    go.set()
    t.join(5)
    reader.join(5)
    assert out and out[0][0].page_content == st._texts[2]


def test_install_rebases_on_langchain_without_touching_readonly_embeddings(numpy_index, monkeypatch):
    """stub langchain_core with the real library's shape: VectorStore.embeddings is a setter-less property"""
    pydantic = pytest.importorskip("pydantic")

    class Embeddings:
        pass

    class VectorStore:
        @property
        def embeddings(self):
            return None

        def as_retriever(self, **kw):
            return ("lc-retriever", self, kw)

    class BaseDocumentCompressor(pydantic.BaseModel):
        pass

    class BaseRetriever(pydantic.BaseModel):
        pass

    mods = {
        "langchain_core": types.ModuleType("langchain_core"),
        "langchain_core.embeddings": types.SimpleNamespace(Embeddings=Embeddings),
        "langchain_core.vectorstores": types.SimpleNamespace(VectorStore=VectorStore),
        "langchain_core.documents": types.SimpleNamespace(BaseDocumentCompressor=BaseDocumentCompressor, Document=Document),
        "langchain_core.retrievers": types.SimpleNamespace(BaseRetriever=BaseRetriever),
    }
    for k, v in mods.items():
        monkeypatch.setitem(sys.modules, k, v)
    from ragmeup_b200 import install
    c = install.build_classes()
    assert issubclass(c["Milvus"], VectorStore) and issubclass(c["PGVector"], VectorStore)
    e = Emb()
    m = c["Milvus"].from_documents([], e, drop_old=True, connection_args={"uri": ""}, collection_name="x")
    assert m.embeddings is e                                  # no AttributeError: nothing assigns the property
    p = c["PGVector"](embeddings=e, collection_name="y", connection="postgresql://", use_jsonb=True)
    assert p.embeddings is e
    assert m.as_retriever(search_type="mmr")[0] == "lc-retriever"
    m.add_documents(_docs(2), ids=["a", "b"])
    assert len(m) == 2


def test_background_segment_writer_flush_and_sync_mode(numpy_index, tmp_path, monkeypatch):
    """segments are written by a background thread in insertion order; flush() makes them durable; RMU_STORE_SYNC=1 writes inline"""
    import glob
    uri = str(tmp_path / "bg.db")
    a = vs.Milvus(Emb(), connection_args={"uri": uri})
    for i in range(12):                                      # more batches than the writer queue holds
        a.add_documents(_docs(7, 7 * i), ids=[f"k{7 * i + j}" for j in range(7)])
    a.flush()
    segs = sorted(glob.glob(os.path.join(a._storage_dir, "seg_*.npz")))
    assert [os.path.basename(f) for f in segs] == [f"seg_{i:08d}.npz" for i in range(12)]
    b = vs.Milvus(Emb(), connection_args={"uri": uri})
    assert b._pks == a._pks and b._texts == a._texts and np.array_equal(b.index.x, a.index.x)
    # a writer failure surfaces on the next flush instead of being lost
    os.chmod(a._storage_dir, 0o500)
    try:
        a.add_documents(_docs(1, 100), ids=["late"])
        if os.geteuid() != 0:                                # root ignores directory permissions
            with pytest.raises(RuntimeError):
                a.flush()
    finally:
        os.chmod(a._storage_dir, 0o700)
    monkeypatch.setenv("RMU_STORE_SYNC", "1")
    c = vs.Milvus(Emb(), connection_args={"uri": str(tmp_path / "sync.db")})
    c.add_documents(_docs(3))
    assert c._writer is None and len(glob.glob(os.path.join(c._storage_dir, "seg_*.npz"))) == 1
