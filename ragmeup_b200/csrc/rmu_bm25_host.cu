// Host side of the sparse leg: building the inverted index from raw texts (no GPU work in this file).
//
// langchain-community's BM25Retriever.from_texts (server/RAGHelper.py:436-443 calls it, and again on every add:
// :531-533) tokenises with `text.split()` and hands the token lists to rank_bm25, whose `_initialize` walks every
// token in Python.  This is that walk in C++: Python `str.split()` semantics on UTF-8 (runs of Unicode whitespace as
// `str.isspace` defines it), vocabulary ids in first-seen order (the order rank_bm25's `nd` dict — and therefore its
// sequential idf sum — sees the words), per-document term frequencies, and the postings CSR by term with documents
// ascending.  The float64 statistics (idf, epsilon floor, length normalisation) stay in Python/numpy so that they are
// computed by the very operations rank_bm25 uses (ragmeup_b200/bm25.py).
#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "rmu_common.h"

namespace {

// str.isspace() code points (Unicode White_Space plus the four ASCII separators 0x1C-0x1F, which Python also splits on)
inline bool py_isspace(uint32_t c) {
    if (c <= 0x20) return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x20);
    if (c < 0x85) return false;
    return c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 ||
           c == 0x202F || c == 0x205F || c == 0x3000;
}

// decode one UTF-8 sequence (input is valid UTF-8 produced by Python's encoder); returns its length
inline int utf8_next(const unsigned char* s, int64_t left, uint32_t* cp) {
    const unsigned char b = s[0];
    if (b < 0x80) { *cp = b; return 1; }
    if ((b >> 5) == 0x6 && left >= 2) { *cp = ((b & 0x1Fu) << 6) | (s[1] & 0x3Fu); return 2; }
    if ((b >> 4) == 0xE && left >= 3) { *cp = ((b & 0x0Fu) << 12) | ((s[1] & 0x3Fu) << 6) | (s[2] & 0x3Fu); return 3; }
    if ((b >> 3) == 0x1E && left >= 4) {
        *cp = ((b & 0x07u) << 18) | ((s[1] & 0x3Fu) << 12) | ((s[2] & 0x3Fu) << 6) | (s[3] & 0x3Fu);
        return 4;
    }
    *cp = 0xFFFD;
    return 1;
}

struct SvHash {
    size_t operator()(const std::pair<const char*, size_t>& k) const {
        uint64_t h = 1469598103934665603ull;                 // FNV-1a
        for (size_t i = 0; i < k.second; ++i) { h ^= static_cast<unsigned char>(k.first[i]); h *= 1099511628211ull; }
        return static_cast<size_t>(h ^ (h >> 29));
    }
};
struct SvEq {
    bool operator()(const std::pair<const char*, size_t>& a, const std::pair<const char*, size_t>& b) const {
        return a.second == b.second && (a.second == 0 || std::memcmp(a.first, b.first, a.second) == 0);
    }
};

}  // namespace

struct rmu_bm25_csr {
    int64_t n_docs = 0;
    std::vector<int64_t> doc_len;                  // tokens per document
    std::vector<int64_t> post_ptr;                 // [n_terms + 1]
    std::vector<int32_t> post_doc, post_tf;        // [nnz]
    std::vector<int64_t> vocab_off;                // [n_terms + 1] byte offsets into vocab_bytes (first-seen order)
    std::string vocab_bytes;                       // UTF-8 terms, concatenated
};

extern "C" {

int rmu_bm25_csr_build(const char* const* texts_utf8, const int64_t* text_bytes, int64_t n_docs, rmu_bm25_csr** out) {
    using rmu::set_error;
    if (!out || n_docs < 0 || (n_docs > 0 && (!texts_utf8 || !text_bytes))) { set_error("rmu_bm25_csr_build: bad argument"); return RMU_ERR_ARG; }
    if (n_docs > 0x7FFFFFF0ll) { set_error("rmu_bm25_csr_build: more than 2^31 documents"); return RMU_ERR_UNSUPPORTED; }
    rmu_bm25_csr* c = new rmu_bm25_csr();
    c->n_docs = n_docs;
    c->doc_len.assign(static_cast<size_t>(n_docs), 0);
    // terms are owned by `arena` chunks so the string views in the map stay valid
    std::vector<std::string*> arena;
    std::unordered_map<std::pair<const char*, size_t>, int32_t, SvHash, SvEq> vocab;
    vocab.reserve(1 << 16);
    std::vector<int64_t> term_off{0};
    std::vector<int32_t> ent_term, ent_doc, ent_tf;    // (term, doc, tf), documents ascending
    std::vector<int32_t> ids;                          // term ids of the current document, token order
    std::string* cur = new std::string();
    cur->reserve(1 << 20);
    arena.push_back(cur);
    for (int64_t d = 0; d < n_docs; ++d) {
        const unsigned char* s = reinterpret_cast<const unsigned char*>(texts_utf8[d]);
        const int64_t n = text_bytes[d];
        ids.clear();
        int64_t i = 0;
        while (i < n) {
            uint32_t cp;
            int l = utf8_next(s + i, n - i, &cp);
            if (py_isspace(cp)) { i += l; continue; }
            const int64_t start = i;
            while (i < n) {
                l = utf8_next(s + i, n - i, &cp);
                if (py_isspace(cp)) break;
                i += l;
            }
            const size_t len = static_cast<size_t>(i - start);
            std::pair<const char*, size_t> key(reinterpret_cast<const char*>(s + start), len);
            auto it = vocab.find(key);
            int32_t tid;
            if (it == vocab.end()) {
                if (cur->size() + len > cur->capacity()) {      // never reallocate a chunk: views point into it
                    cur = new std::string();
                    cur->reserve(std::max<size_t>(1 << 20, len));
                    arena.push_back(cur);
                }
                const size_t pos = cur->size();
                cur->append(key.first, len);
                tid = static_cast<int32_t>(vocab.size());
                vocab.emplace(std::make_pair(cur->data() + pos, len), tid);
                c->vocab_bytes.append(key.first, len);
                term_off.push_back(static_cast<int64_t>(c->vocab_bytes.size()));
            } else {
                tid = it->second;
            }
            ids.push_back(tid);
        }
        c->doc_len[static_cast<size_t>(d)] = static_cast<int64_t>(ids.size());
        std::sort(ids.begin(), ids.end());
        for (size_t a = 0; a < ids.size();) {
            size_t b = a + 1;
            while (b < ids.size() && ids[b] == ids[a]) ++b;
            ent_term.push_back(ids[a]);
            ent_doc.push_back(static_cast<int32_t>(d));
            ent_tf.push_back(static_cast<int32_t>(b - a));
            a = b;
        }
    }
    for (std::string* p : arena) delete p;
    const size_t n_terms = term_off.size() - 1, nnz = ent_term.size();
    c->vocab_off = std::move(term_off);
    // counting sort by term (stable: documents stay ascending inside a term)
    c->post_ptr.assign(n_terms + 1, 0);
    for (size_t e = 0; e < nnz; ++e) ++c->post_ptr[static_cast<size_t>(ent_term[e]) + 1];
    for (size_t t = 0; t < n_terms; ++t) c->post_ptr[t + 1] += c->post_ptr[t];
    c->post_doc.resize(nnz);
    c->post_tf.resize(nnz);
    std::vector<int64_t> fill(c->post_ptr.begin(), c->post_ptr.end() - 1);
    for (size_t e = 0; e < nnz; ++e) {
        const int64_t pos = fill[static_cast<size_t>(ent_term[e])]++;
        c->post_doc[static_cast<size_t>(pos)] = ent_doc[e];
        c->post_tf[static_cast<size_t>(pos)] = ent_tf[e];
    }
    *out = c;
    return RMU_OK;
}

int rmu_bm25_csr_sizes(const rmu_bm25_csr* c, int64_t* n_terms, int64_t* nnz, int64_t* vocab_bytes) {
    if (!c || !n_terms || !nnz || !vocab_bytes) { rmu::set_error("rmu_bm25_csr_sizes: bad argument"); return RMU_ERR_ARG; }
    *n_terms = static_cast<int64_t>(c->post_ptr.size()) - 1;
    *nnz = static_cast<int64_t>(c->post_doc.size());
    *vocab_bytes = static_cast<int64_t>(c->vocab_bytes.size());
    return RMU_OK;
}

int rmu_bm25_csr_export(const rmu_bm25_csr* c, int64_t* doc_len, int64_t* post_ptr, int32_t* post_doc, int32_t* post_tf,
                        int64_t* vocab_off, char* vocab_bytes) {
    if (!c || !doc_len || !post_ptr || !vocab_off) { rmu::set_error("rmu_bm25_csr_export: bad argument"); return RMU_ERR_ARG; }
    std::copy(c->doc_len.begin(), c->doc_len.end(), doc_len);
    std::copy(c->post_ptr.begin(), c->post_ptr.end(), post_ptr);
    if (post_doc) std::copy(c->post_doc.begin(), c->post_doc.end(), post_doc);
    if (post_tf) std::copy(c->post_tf.begin(), c->post_tf.end(), post_tf);
    std::copy(c->vocab_off.begin(), c->vocab_off.end(), vocab_off);
    if (vocab_bytes && !c->vocab_bytes.empty()) std::memcpy(vocab_bytes, c->vocab_bytes.data(), c->vocab_bytes.size());
    return RMU_OK;
}

int rmu_bm25_csr_free(rmu_bm25_csr* c) {
    delete c;
    return RMU_OK;
}

}  // extern "C"
