"""CPU: the C-ABI library builds, loads, and exports every symbol include/ragmeup_b200.h declares
(no compute calls here — there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from ragmeup_b200 import build
    path = build.build()
    assert os.path.exists(path)
    return path


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ragmeup_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rmu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("rmu_index_create", "rmu_index_add", "rmu_index_search", "rmu_index_search_host", "rmu_topk_merge",
                 "rmu_mmr_select", "rmu_index_gather", "rmu_encoder_create", "rmu_encoder_embed",
                 "rmu_encoder_classify", "rmu_encoder_embed_host", "rmu_encoder_classify_host", "rmu_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_loader_signatures_cover_header(built_lib):
    from ragmeup_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    l = _lib.lib()
    assert l.rmu_version() >= 100
    assert isinstance(_lib.launch_count(), int)


def test_product_fails_loudly_without_cuda(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ragmeup_b200 import _lib
    from ragmeup_b200.index import FlatIndex
    with pytest.raises(_lib.RmuError):
        FlatIndex(384, "l2")
    from ragmeup_b200.embeddings import HuggingFaceEmbeddings
    with pytest.raises(_lib.RmuError):
        HuggingFaceEmbeddings(model_name="synthetic:tiny", model_kwargs={"device": "cpu"})
    with pytest.raises(_lib.RmuError):
        HuggingFaceEmbeddings(model_name="synthetic:tiny", model_kwargs={"device": "cuda"})


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: no product module may import it"""
    pkg = os.path.join(ROOT, "ragmeup_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
