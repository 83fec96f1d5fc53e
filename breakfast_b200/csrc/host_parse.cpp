// host_parse.cpp — native host side of the parse path (libbfhost.so, plain C ABI, no CUDA).
//
// Does the per-occurrence work of the reference's filter_features + collapse_duplicates +
// sparse_feature_matrix (src/breakfast/breakfast.py:116-190, 72-79, 193-215) in two passes:
//   1. bfh_tokenise: split every profile on the separator, intern the tokens (ids by first appearance).
//      The caller classifies the DISTINCT tokens with the reference's regular expressions in Python, so
//      the classification semantics (unicode digits, '$' before a trailing newline, ...) stay exactly
//      Python's; only ~1e5 distinct tokens exist for ~1e8 occurrences.
//   2. bfh_build: apply the verdicts (keep / drop / invalid), deduplicate the filtered profiles in
//      first-appearance order (equal filtered strings <=> equal kept-token sequences, because the
//      reference re-joins the kept tokens; with no filter active the raw strings are compared), and emit
//      for the unique profiles: the token CSR with the vocabulary in first-appearance order, the
//      strictly binary CSR the device wants (repeats thermometer-coded, columns ascending), and the
//      filtered strings.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

inline uint64_t hash_bytes(const char* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 0x100000001b3ull; }
    h ^= h >> 32;
    return h;
}

struct Interner {  // open addressing over (offset, length) into one byte arena
    std::vector<int32_t> slots;
    std::vector<int64_t> off;
    std::vector<int32_t> len;
    std::string arena;
    uint64_t mask = 0;
    void init(size_t cap) { size_t c = 1024; while (c < cap * 2) c <<= 1; slots.assign(c, -1); mask = c - 1; }
    void grow() {
        std::vector<int32_t> old; old.swap(slots);
        slots.assign(old.size() * 2, -1); mask = slots.size() - 1;
        for (int32_t id = 0; id < (int32_t)off.size(); ++id) {
            uint64_t i = hash_bytes(arena.data() + off[id], (size_t)len[id]) & mask;
            while (slots[i] >= 0) i = (i + 1) & mask;
            slots[i] = id;
        }
    }
    int32_t intern(const char* p, size_t n) {
        uint64_t i = hash_bytes(p, n) & mask;
        while (slots[i] >= 0) {
            int32_t id = slots[i];
            if ((size_t)len[id] == n && memcmp(arena.data() + off[id], p, n) == 0) return id;
            i = (i + 1) & mask;
        }
        int32_t id = (int32_t)off.size();
        off.push_back((int64_t)arena.size());
        len.push_back((int32_t)n);
        arena.append(p, n);
        slots[i] = id;
        if (off.size() * 2 > slots.size()) grow();
        return id;
    }
};

struct State {
    int64_t n_seq = 0;
    std::string sep;
    const char* buf = nullptr;             // caller's buffer (valid until bfh_build returns)
    std::vector<int64_t> rec_off;          // n_seq + 1 record offsets into buf (records separated by rec_sep)
    Interner tokens;
    std::vector<int64_t> tok_ptr;          // n_seq + 1
    std::vector<int32_t> tok_ids;          // raw token occurrences (empty tokens included)
    // results of bfh_build
    std::vector<int32_t> codes, first_seq, invalid;
    std::vector<int64_t> u_ptr, b_ptr, s_off;
    std::vector<int32_t> u_idx, b_idx;
    std::string s_bytes;
    int32_t n_vocab = 0, n_cols = 0;
};

struct VecHash {  // word-wise mix over the token ids of a profile
    size_t operator()(const std::pair<const int32_t*, size_t>& k) const {
        uint64_t h = 0x9e3779b97f4a7c15ull ^ k.second;
        for (size_t i = 0; i < k.second; ++i) {
            h ^= (uint64_t)(uint32_t)k.first[i] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
            h *= 0xff51afd7ed558ccdull;
        }
        return (size_t)(h ^ (h >> 32));
    }
};
struct VecEq {
    bool operator()(const std::pair<const int32_t*, size_t>& a, const std::pair<const int32_t*, size_t>& b) const {
        return a.second == b.second && (a.second == 0 || memcmp(a.first, b.first, a.second * sizeof(int32_t)) == 0);
    }
};

}  // namespace

extern "C" {

void* bfh_tokenise(const char* buf, int64_t buf_len, char rec_sep, const char* sep, int32_t sep_len, int64_t n_seq) {
    if (!buf || n_seq < 0 || sep_len <= 0) return nullptr;
    State* st = new State();
    st->n_seq = n_seq;
    st->sep.assign(sep, (size_t)sep_len);
    st->buf = buf;
    st->rec_off.reserve((size_t)n_seq + 1);
    st->tok_ptr.reserve((size_t)n_seq + 1);
    st->tok_ids.reserve((size_t)(buf_len / 6 + 16));
    st->tokens.init(1 << 16);
    st->tok_ptr.push_back(0);
    int64_t pos = 0;
    for (int64_t r = 0; r < n_seq; ++r) {
        st->rec_off.push_back(pos);
        const char* e = (const char*)memchr(buf + pos, rec_sep, (size_t)(buf_len - pos));
        const int64_t end = e ? (int64_t)(e - buf) : buf_len;
        // Python's str.split(sep): n separators give n + 1 tokens, empty ones included
        int64_t p = pos;
        while (true) {
            int64_t q = end;
            if (sep_len == 1) {
                const char* f = (const char*)memchr(buf + p, sep[0], (size_t)(end - p));
                if (f) q = (int64_t)(f - buf);
            } else {
                for (int64_t k = p; k + sep_len <= end; ++k)
                    if (memcmp(buf + k, sep, (size_t)sep_len) == 0) { q = k; break; }
            }
            st->tok_ids.push_back(st->tokens.intern(buf + p, (size_t)(q - p)));
            if (q >= end) break;
            p = q + sep_len;
        }
        st->tok_ptr.push_back((int64_t)st->tok_ids.size());
        pos = end + 1;
    }
    st->rec_off.push_back(pos);
    return st;
}

int64_t bfh_n_tokens(void* h) { return (int64_t)static_cast<State*>(h)->tok_ids.size(); }
int32_t bfh_n_distinct(void* h) { return (int32_t)static_cast<State*>(h)->tokens.off.size(); }
int64_t bfh_distinct_bytes(void* h) { return (int64_t)static_cast<State*>(h)->tokens.arena.size(); }
void bfh_get_distinct(void* h, char* bytes, int64_t* offsets) {
    State* st = static_cast<State*>(h);
    memcpy(bytes, st->tokens.arena.data(), st->tokens.arena.size());
    const size_t n = st->tokens.off.size();
    for (size_t i = 0; i < n; ++i) offsets[i] = st->tokens.off[i];
    offsets[n] = (int64_t)st->tokens.arena.size();
}

// verdict[distinct token]: 0 keep, 1 drop silently, 2 invalid (dropped and reported).
// filter_active = 0: profiles are compared (and returned) as raw strings, tokens are only needed for the
// matrix; empty tokens never enter the matrix (breakfast.py:208-209).
int bfh_build(void* h, const uint8_t* verdict, int filter_active) {
    State* st = static_cast<State*>(h);
    const bool timing = getenv("BFH_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[bfh_build] %s %.3f s\n", what, std::chrono::duration<double>(now - t_prev).count());
        t_prev = now;
    };
    const int64_t n = st->n_seq;
    const int32_t n_distinct = (int32_t)st->tokens.off.size();
    std::vector<uint8_t> is_empty((size_t)n_distinct);
    for (int32_t t = 0; t < n_distinct; ++t) is_empty[t] = st->tokens.len[t] == 0;
    st->codes.assign((size_t)n, -1);
    std::vector<int32_t> kept;               // kept token ids of all sequences, concatenated
    std::vector<int64_t> kept_ptr((size_t)n + 1, 0);
    kept.reserve(st->tok_ids.size());
    for (int64_t r = 0; r < n; ++r) {
        for (int64_t k = st->tok_ptr[r]; k < st->tok_ptr[r + 1]; ++k) {
            const int32_t t = st->tok_ids[k];
            if (filter_active) {
                if (verdict[t] == 2) { st->invalid.push_back(t); continue; }
                if (verdict[t] == 1) continue;
            }
            if (is_empty[t]) continue;
            kept.push_back(t);
        }
        kept_ptr[r + 1] = (int64_t)kept.size();
    }
    lap("filter");
    // dedup in first-appearance order
    std::vector<int64_t> uniq_rows;
    if (filter_active) {
        std::unordered_map<std::pair<const int32_t*, size_t>, int32_t, VecHash, VecEq> seen;
        seen.reserve((size_t)n);
        for (int64_t r = 0; r < n; ++r) {
            auto key = std::make_pair(kept.data() + kept_ptr[r], (size_t)(kept_ptr[r + 1] - kept_ptr[r]));
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(key, (int32_t)uniq_rows.size()).first;
                uniq_rows.push_back(r);
            }
            st->codes[r] = it->second;
        }
    } else {
        Interner raw;
        raw.init((size_t)n);
        for (int64_t r = 0; r < n; ++r) {
            const int64_t a = st->rec_off[r], b = st->rec_off[r + 1] - 1;
            const int32_t before = (int32_t)raw.off.size();
            const int32_t id = raw.intern(st->buf + a, (size_t)(b - a));
            if (id == before) uniq_rows.push_back(r);
            st->codes[r] = id;
        }
    }
    lap("dedup");
    const int64_t nu = (int64_t)uniq_rows.size();
    st->first_seq.resize((size_t)nu);
    // token CSR of the unique profiles, vocabulary by first appearance (breakfast.py:199-213)
    std::vector<int32_t> vocab_of((size_t)n_distinct, -1);
    st->u_ptr.assign((size_t)nu + 1, 0);
    st->s_off.assign((size_t)nu + 1, 0);
    st->u_idx.reserve(kept.size());
    st->s_bytes.reserve(kept.size() * 8);
    for (int64_t u = 0; u < nu; ++u) {
        const int64_t r = uniq_rows[u];
        st->first_seq[u] = (int32_t)r;
        for (int64_t k = kept_ptr[r]; k < kept_ptr[r + 1]; ++k) {
            const int32_t t = kept[k];
            if (vocab_of[t] < 0) vocab_of[t] = st->n_vocab++;
            st->u_idx.push_back(vocab_of[t]);
        }
        st->u_ptr[u + 1] = (int64_t)st->u_idx.size();
        // the profile string the reference keeps in meta["feature"]
        if (filter_active) {
            for (int64_t k = kept_ptr[r]; k < kept_ptr[r + 1]; ++k) {
                if (k > kept_ptr[r]) st->s_bytes += st->sep;
                const int32_t t = kept[k];
                st->s_bytes.append(st->tokens.arena.data() + st->tokens.off[t], (size_t)st->tokens.len[t]);
            }
        } else {
            st->s_bytes.append(st->buf + st->rec_off[r], (size_t)(st->rec_off[r + 1] - 1 - st->rec_off[r]));
        }
        st->s_off[u + 1] = (int64_t)st->s_bytes.size();
    }
    lap("token csr + strings");
    // strictly binary rows: k-th repeat (k >= 1) of a token inside a profile becomes its own column
    std::unordered_map<int64_t, int32_t> extra;
    st->n_cols = st->n_vocab;
    st->b_ptr.assign((size_t)nu + 1, 0);
    st->b_idx.reserve(st->u_idx.size());
    std::vector<int32_t> row;
    for (int64_t u = 0; u < nu; ++u) {
        row.assign(st->u_idx.begin() + st->u_ptr[u], st->u_idx.begin() + st->u_ptr[u + 1]);
        if (!std::is_sorted(row.begin(), row.end())) std::sort(row.begin(), row.end());
        const size_t base = st->b_idx.size();
        bool has_extra = false;
        for (size_t k = 0; k < row.size();) {
            size_t e = k;
            while (e < row.size() && row[e] == row[k]) ++e;
            st->b_idx.push_back(row[k]);
            for (size_t rep = 1; rep < e - k; ++rep) {
                const int64_t key = ((int64_t)row[k] << 32) | (int64_t)std::min<size_t>(rep, 0x7fffffff);
                auto it = extra.find(key);
                if (it == extra.end()) it = extra.emplace(key, st->n_cols++).first;
                st->b_idx.push_back(it->second);
                has_extra = true;
            }
            k = e;
        }
        if (has_extra) std::sort(st->b_idx.begin() + (int64_t)base, st->b_idx.end());
        st->b_ptr[u + 1] = (int64_t)st->b_idx.size();
    }
    lap("binary csr");
    st->buf = nullptr;
    return 0;
}

int64_t bfh_n_unique(void* h) { return (int64_t)static_cast<State*>(h)->first_seq.size(); }
int64_t bfh_n_invalid(void* h) { return (int64_t)static_cast<State*>(h)->invalid.size(); }
int32_t bfh_n_vocab(void* h) { return static_cast<State*>(h)->n_vocab; }
int32_t bfh_n_cols(void* h) { return static_cast<State*>(h)->n_cols; }
int64_t bfh_token_nnz(void* h) { return (int64_t)static_cast<State*>(h)->u_idx.size(); }
int64_t bfh_binary_nnz(void* h) { return (int64_t)static_cast<State*>(h)->b_idx.size(); }
int64_t bfh_string_bytes(void* h) { return (int64_t)static_cast<State*>(h)->s_bytes.size(); }

void bfh_get_results(void* h, int32_t* codes, int32_t* first_seq, int32_t* invalid, int64_t* u_ptr, int32_t* u_idx,
                     int64_t* b_ptr, int32_t* b_idx, int64_t* s_off, char* s_bytes) {
    State* st = static_cast<State*>(h);
    auto cp = [](void* dst, const void* src, size_t bytes) { if (dst && bytes) memcpy(dst, src, bytes); };
    cp(codes, st->codes.data(), st->codes.size() * 4);
    cp(first_seq, st->first_seq.data(), st->first_seq.size() * 4);
    cp(invalid, st->invalid.data(), st->invalid.size() * 4);
    cp(u_ptr, st->u_ptr.data(), st->u_ptr.size() * 8);
    cp(u_idx, st->u_idx.data(), st->u_idx.size() * 4);
    cp(b_ptr, st->b_ptr.data(), st->b_ptr.size() * 8);
    cp(b_idx, st->b_idx.data(), st->b_idx.size() * 4);
    cp(s_off, st->s_off.data(), st->s_off.size() * 8);
    cp(s_bytes, st->s_bytes.data(), st->s_bytes.size());
}

void bfh_free(void* h) { delete static_cast<State*>(h); }

}  // extern "C"
