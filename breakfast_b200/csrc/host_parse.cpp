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
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// run fn(t, lo, hi) over [0, n) cut into n_threads contiguous parts (part t = records lo .. hi - 1)
template <typename F>
void parallel_ranges(int64_t n, int n_threads, F fn) {
    n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n / 2048 + 1));
    if (n_threads == 1) { fn(0, (int64_t)0, n); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
        pool.emplace_back([=]() { fn(t, lo, hi); });
    }
    for (auto& th : pool) th.join();
}

inline uint64_t hash_bytes(const char* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 0x100000001b3ull; }
    h ^= h >> 32;
    return h;
}

// A big array that is written completely by the threads that fill it: plain malloc, no value-initialisation (a
// std::vector would zero hundreds of MB on one thread first, and take the page faults there too)
template <typename T>
struct RawVec {
    T* p = nullptr;
    size_t n = 0;
    RawVec() = default;
    RawVec(const RawVec&) = delete;
    RawVec& operator=(const RawVec&) = delete;
    ~RawVec() { free(p); }
    void resize(size_t m) {
        free(p);
        p = static_cast<T*>(malloc(std::max<size_t>(m, 1) * sizeof(T)));
        n = p ? m : 0;
    }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

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
    int n_threads = 1;
    int rec_gap = 1;                       // bytes between two records in buf (1: a separator byte, 0: Arrow string buffer)
    std::string sep;
    const char* buf = nullptr;             // caller's buffer (valid until bfh_build returns)
    std::vector<int64_t> rec_off;          // n_seq + 1 record offsets into buf: record r = [rec_off[r], rec_off[r + 1] - rec_gap)
    std::vector<int64_t> rec_end;          // chunked Arrow input only (several data buffers): record r = [rec_off[r], rec_end[r]),
                                           // both relative to buf = the lowest of the buffers' addresses
    Interner tokens;
    std::vector<int64_t> tok_ptr;          // n_seq + 1
    RawVec<int32_t> tok_ids;               // raw token occurrences (empty tokens included)
    // results of bfh_build
    std::vector<int32_t> codes, first_seq, invalid;
    std::vector<int64_t> u_ptr, b_ptr, s_off;
    RawVec<int32_t> u_idx, b_idx;
    RawVec<char> s_bytes;
    int32_t n_vocab = 0, n_cols = 0;
};

inline int64_t rec_a(const State* st, int64_t r) { return st->rec_off[(size_t)r]; }
inline int64_t rec_e(const State* st, int64_t r) {
    return st->rec_end.empty() ? st->rec_off[(size_t)r + 1] - st->rec_gap : st->rec_end[(size_t)r];
}

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

// Python's str.split(sep) of one record: n separators give n + 1 tokens, empty ones included
static inline void split_record(const char* buf, int64_t p, int64_t end, const char* sep, int32_t sep_len, Interner& tokens,
                                std::vector<int32_t>& out) {
    while (true) {
        int64_t q = end;
        if (sep_len == 1) {
            const char* f = (const char*)memchr(buf + p, sep[0], (size_t)(end - p));
            if (f) q = (int64_t)(f - buf);
        } else {
            for (int64_t k = p; k + sep_len <= end; ++k)
                if (memcmp(buf + k, sep, (size_t)sep_len) == 0) { q = k; break; }
        }
        out.push_back(tokens.intern(buf + p, (size_t)(q - p)));
        if (q >= end) break;
        p = q + sep_len;
    }
}

// tokenise all records (rec_off / rec_gap / buf set): every thread interns into its own table, then the tables are
// merged (token ids are internal: nothing depends on their order) and the occurrences renumbered in parallel
static void tokenise_records(State* st) {
    const int64_t n = st->n_seq;
    const int T = std::max(1, st->n_threads);
    struct Piece { Interner tokens; std::vector<int32_t> ids; std::vector<int64_t> cnt; };
    std::vector<Piece> pieces((size_t)T);
    std::vector<int64_t> lo_of((size_t)T + 1, n);
    const char* sep = st->sep.data();
    const int32_t sep_len = (int32_t)st->sep.size();
    int used = 0;
    {
        std::vector<int> seen((size_t)T, 0);
        parallel_ranges(n, T, [&](int t, int64_t lo, int64_t hi) {
            Piece& pc = pieces[(size_t)t];
            seen[(size_t)t] = 1;
            lo_of[(size_t)t] = lo;
            pc.tokens.init(1 << 16);
            pc.cnt.reserve((size_t)(hi - lo));
            int64_t bytes = 0;
            if (st->rec_end.empty()) bytes = st->rec_off[hi] - st->rec_off[lo];
            else for (int64_t r = lo; r < hi; ++r) bytes += rec_e(st, r) - rec_a(st, r);
            pc.ids.reserve((size_t)(bytes / 6 + 16));
            for (int64_t r = lo; r < hi; ++r) {
                const size_t before = pc.ids.size();
                split_record(st->buf, rec_a(st, r), rec_e(st, r), sep, sep_len, pc.tokens, pc.ids);
                pc.cnt.push_back((int64_t)(pc.ids.size() - before));
            }
        });
        for (int t = 0; t < T; ++t) used += seen[(size_t)t];
    }
    // merge the vocabularies
    st->tokens.init(1 << 16);
    std::vector<std::vector<int32_t>> remap((size_t)used);
    for (int t = 0; t < used; ++t) {
        const Interner& loc = pieces[(size_t)t].tokens;
        remap[(size_t)t].resize(loc.off.size());
        for (size_t i = 0; i < loc.off.size(); ++i)
            remap[(size_t)t][i] = st->tokens.intern(loc.arena.data() + loc.off[i], (size_t)loc.len[i]);
    }
    // token pointers + renumbered occurrences
    st->tok_ptr.assign((size_t)n + 1, 0);
    std::vector<int64_t> piece_base((size_t)used + 1, 0);
    for (int t = 0; t < used; ++t) piece_base[(size_t)t + 1] = piece_base[(size_t)t] + (int64_t)pieces[(size_t)t].ids.size();
    st->tok_ids.resize((size_t)piece_base[(size_t)used]);
    std::vector<std::thread> pool;
    for (int t = 0; t < used; ++t)
        pool.emplace_back([&, t]() {
            const Piece& pc = pieces[(size_t)t];
            int64_t pos = piece_base[(size_t)t];
            const int64_t lo = lo_of[(size_t)t];
            for (size_t k = 0; k < pc.cnt.size(); ++k) {
                st->tok_ptr[(size_t)(lo + (int64_t)k)] = pos;
                pos += pc.cnt[k];
            }
            const std::vector<int32_t>& m = remap[(size_t)t];
            int32_t* dst = st->tok_ids.data() + piece_base[(size_t)t];
            for (size_t k = 0; k < pc.ids.size(); ++k) dst[k] = m[(size_t)pc.ids[k]];
        });
    for (auto& th : pool) th.join();
    st->tok_ptr[(size_t)n] = piece_base[(size_t)used];
}

void* bfh_tokenise(const char* buf, int64_t buf_len, char rec_sep, const char* sep, int32_t sep_len, int64_t n_seq) {
    if (!buf || n_seq < 0 || sep_len <= 0) return nullptr;
    State* st = new State();
    st->n_seq = n_seq;
    st->sep.assign(sep, (size_t)sep_len);
    st->buf = buf;
    st->rec_gap = 1;
    st->rec_off.reserve((size_t)n_seq + 1);
    int64_t pos = 0;
    for (int64_t r = 0; r < n_seq; ++r) {
        st->rec_off.push_back(pos);
        const char* e = (const char*)memchr(buf + pos, rec_sep, (size_t)(buf_len - pos));
        pos = (e ? (int64_t)(e - buf) : buf_len) + 1;
    }
    st->rec_off.push_back(pos);
    tokenise_records(st);
    return st;
}

// records given as an Arrow string array: `data` + n_seq + 1 offsets (32- or 64-bit); n_threads host threads
void* bfh_tokenise_arrow(const char* data, const void* offsets, int32_t offset_bytes, int64_t n_seq, const char* sep, int32_t sep_len,
                         int32_t n_threads) {
    if ((!data && n_seq > 0) || !offsets || n_seq < 0 || sep_len <= 0 || (offset_bytes != 4 && offset_bytes != 8)) return nullptr;
    State* st = new State();
    st->n_seq = n_seq;
    st->n_threads = std::max(1, (int)n_threads);
    st->sep.assign(sep, (size_t)sep_len);
    st->buf = data;
    st->rec_gap = 0;
    st->rec_off.resize((size_t)n_seq + 1);
    for (int64_t r = 0; r <= n_seq; ++r)
        st->rec_off[(size_t)r] = offset_bytes == 4 ? (int64_t)static_cast<const int32_t*>(offsets)[r] : static_cast<const int64_t*>(offsets)[r];
    tokenise_records(st);
    return st;
}

// the same for a CHUNKED Arrow string column (what the Arrow CSV reader hands over: hundreds of chunks per million
// lines), read in place: no concatenation of the chunks (660 MB at 10^6 profiles).  Chunk c: data buffer data[c], row
// offsets offsets[c] (n_rows[c] + 1 entries of offset_bytes each, already advanced to the chunk's first row).
void* bfh_tokenise_arrow_chunks(int32_t n_chunks, const char* const* data, const void* const* offsets, const int64_t* n_rows,
                                int32_t offset_bytes, const char* sep, int32_t sep_len, int32_t n_threads) {
    if (n_chunks < 0 || sep_len <= 0 || (offset_bytes != 4 && offset_bytes != 8) || (n_chunks > 0 && (!data || !offsets || !n_rows)))
        return nullptr;
    int64_t n_seq = 0;
    const char* base = nullptr;
    for (int32_t c = 0; c < n_chunks; ++c) {
        if (n_rows[c] < 0 || !offsets[c]) return nullptr;
        n_seq += n_rows[c];
        if (data[c] && (!base || data[c] < base)) base = data[c];
    }
    State* st = new State();
    st->n_seq = n_seq;
    st->n_threads = std::max(1, (int)n_threads);
    st->sep.assign(sep, (size_t)sep_len);
    static const char empty_buffer[1] = {0};
    st->buf = base ? base : empty_buffer;
    st->rec_gap = 0;
    st->rec_off.resize((size_t)n_seq + 1);
    st->rec_end.resize((size_t)n_seq + 1);   // never empty, so that rec_e takes this branch even without rows
    int64_t r = 0;
    for (int32_t c = 0; c < n_chunks; ++c) {
        const int64_t shift = data[c] ? (int64_t)(data[c] - st->buf) : 0;
        for (int64_t k = 0; k < n_rows[c]; ++k, ++r) {
            const int64_t a = offset_bytes == 4 ? (int64_t)static_cast<const int32_t*>(offsets[c])[k] : static_cast<const int64_t*>(offsets[c])[k];
            const int64_t e = offset_bytes == 4 ? (int64_t)static_cast<const int32_t*>(offsets[c])[k + 1] : static_cast<const int64_t*>(offsets[c])[k + 1];
            st->rec_off[(size_t)r] = shift + a;
            st->rec_end[(size_t)r] = shift + e;
        }
    }
    st->rec_off[(size_t)n_seq] = n_seq ? st->rec_end[(size_t)n_seq - 1] : 0;
    tokenise_records(st);
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

// ---- native verdicts for the two DNA notations (reference breakfast.py:135-184, the regular expressions restated as
// byte scanners; classification order substitution, insertion, deletion; trimming applies to substitutions only and
// is inclusive at both ends).  Only pure-ASCII tokens without a line break are judged here - for those Python's
// `re` (\d, [A-Z], ., $) and these scanners agree by construction; every other token gets BFH_ASK_PYTHON and the
// caller runs the reference's own expressions on it.
enum { BFH_KEEP = 0, BFH_DROP = 1, BFH_INVALID = 2, BFH_ASK_PYTHON = 255 };
namespace {
inline bool is_up(char c) { return c >= 'A' && c <= 'Z'; }
inline bool is_dig(char c) { return c >= '0' && c <= '9'; }
// ^[A-Z](\d+)[A-Z]$ -> position (saturated) in *pos
inline bool match_substitution(const char* p, size_t n, int64_t* pos) {
    if (n < 3 || !is_up(p[0]) || !is_up(p[n - 1])) return false;
    int64_t v = 0;
    for (size_t i = 1; i + 1 < n; ++i) {
        if (!is_dig(p[i])) return false;
        v = v > (INT64_MAX - 9) / 10 ? INT64_MAX : v * 10 + (p[i] - '0');
    }
    *pos = v;
    return true;
}
// \d+ starting at i; returns the index after the digits (i itself if there is none)
inline size_t skip_digits(const char* p, size_t n, size_t i) { while (i < n && is_dig(p[i])) ++i; return i; }
inline bool match_insertion(int type, const char* p, size_t n) {
    if (type == 0) return n >= 2 && is_up(p[n - 1]) && is_up(p[n - 2]);          // covsonar_dna  ^.*[A-Z][A-Z]$
    size_t i = skip_digits(p, n, 0);                                               // nextclade_dna ^\d+:[A-Z]+$
    if (i == 0 || i >= n || p[i] != ':' || i + 1 >= n) return false;
    for (size_t k = i + 1; k < n; ++k) if (!is_up(p[k])) return false;
    return true;
}
inline bool match_deletion(int type, const char* p, size_t n) {
    if (type == 0) {                                                               // covsonar_dna  ^del:\d+:\d+$
        if (n < 7 || memcmp(p, "del:", 4) != 0) return false;
        size_t i = skip_digits(p, n, 4);
        if (i == 4 || i >= n || p[i] != ':') return false;
        const size_t j = skip_digits(p, n, i + 1);
        return j > i + 1 && j == n;
    }
    size_t i = skip_digits(p, n, 0);                                               // nextclade_dna ^\d+(-\d+)?$
    if (i == 0) return false;
    if (i == n) return true;
    if (p[i] != '-') return false;
    const size_t j = skip_digits(p, n, i + 1);
    return j > i + 1 && j == n;
}
}  // namespace

// var_type: 0 = covsonar_dna, 1 = nextclade_dna.  upper_cut = reference_length - trim_end.  verdict[n_distinct] out.
// Returns the number of tokens left to Python (BFH_ASK_PYTHON).
int32_t bfh_classify_dna(void* h, int32_t var_type, int32_t skip_ins, int32_t skip_del, int64_t trim_start, int64_t upper_cut,
                         uint8_t* verdict) {
    State* st = static_cast<State*>(h);
    const int32_t n = (int32_t)st->tokens.off.size();
    int32_t ask = 0;
    for (int32_t t = 0; t < n; ++t) {
        const char* p = st->tokens.arena.data() + st->tokens.off[t];
        const size_t len = (size_t)st->tokens.len[t];
        bool plain = true;
        for (size_t i = 0; i < len; ++i) plain &= ((unsigned char)p[i] < 0x80) && p[i] != '\n';
        if (!plain) { verdict[t] = BFH_ASK_PYTHON; ++ask; continue; }
        int64_t pos = 0;
        if (match_substitution(p, len, &pos)) verdict[t] = (pos <= trim_start || pos >= upper_cut) ? BFH_DROP : BFH_KEEP;
        else if (match_insertion(var_type, p, len)) verdict[t] = skip_ins ? BFH_DROP : BFH_KEEP;
        else if (match_deletion(var_type, p, len)) verdict[t] = skip_del ? BFH_DROP : BFH_KEEP;
        else verdict[t] = BFH_INVALID;
    }
    return ask;
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
    const int T = std::max(1, st->n_threads);
    const int32_t n_distinct = (int32_t)st->tokens.off.size();
    std::vector<uint8_t> is_empty((size_t)n_distinct);
    for (int32_t t = 0; t < n_distinct; ++t) is_empty[t] = st->tokens.len[t] == 0;
    st->codes.assign((size_t)n, -1);
    // ---- filter: kept token ids of all sequences, concatenated (every thread filters a contiguous range of records)
    RawVec<int32_t> kept;
    std::vector<int64_t> kept_ptr((size_t)n + 1, 0);
    {
        struct Part { std::vector<int32_t> kept, invalid; int64_t lo = 0, hi = 0; };
        std::vector<Part> parts((size_t)T);
        int used = 0;
        std::vector<int> seen((size_t)T, 0);
        parallel_ranges(n, T, [&](int t, int64_t lo, int64_t hi) {
            Part& pt = parts[(size_t)t];
            seen[(size_t)t] = 1;
            pt.lo = lo;
            pt.hi = hi;
            pt.kept.reserve((size_t)(st->tok_ptr[hi] - st->tok_ptr[lo]));
            for (int64_t r = lo; r < hi; ++r) {
                const size_t before = pt.kept.size();
                for (int64_t k = st->tok_ptr[r]; k < st->tok_ptr[r + 1]; ++k) {
                    const int32_t tk = st->tok_ids[k];
                    if (filter_active) {
                        if (verdict[tk] == 2) { pt.invalid.push_back(tk); continue; }
                        if (verdict[tk] == 1) continue;
                    }
                    if (is_empty[tk]) continue;
                    pt.kept.push_back(tk);
                }
                kept_ptr[(size_t)r + 1] = (int64_t)(pt.kept.size() - before);   // length for now
            }
        });
        for (int t = 0; t < T; ++t) used += seen[(size_t)t];
        for (int64_t r = 0; r < n; ++r) kept_ptr[(size_t)r + 1] += kept_ptr[(size_t)r];
        kept.resize((size_t)kept_ptr[(size_t)n]);
        std::vector<std::thread> pool;
        for (int t = 0; t < used; ++t)
            pool.emplace_back([&, t]() {
                const Part& pt = parts[(size_t)t];
                if (!pt.kept.empty()) memcpy(kept.data() + kept_ptr[(size_t)pt.lo], pt.kept.data(), pt.kept.size() * sizeof(int32_t));
            });
        for (auto& th : pool) th.join();
        for (int t = 0; t < used; ++t) st->invalid.insert(st->invalid.end(), parts[(size_t)t].invalid.begin(), parts[(size_t)t].invalid.end());
    }
    lap("filter");
    // ---- dedup in first-appearance order: hashes in parallel, one sequential pass over an open-addressing table
    std::vector<int64_t> uniq_rows;
    {
        std::vector<uint64_t> hashes((size_t)n);
        parallel_ranges(n, T, [&](int, int64_t lo, int64_t hi) {
            for (int64_t r = lo; r < hi; ++r) {
                if (filter_active) {
                    hashes[(size_t)r] = (uint64_t)VecHash()(std::make_pair((const int32_t*)kept.data() + kept_ptr[(size_t)r],
                                                                           (size_t)(kept_ptr[(size_t)r + 1] - kept_ptr[(size_t)r])));
                } else {
                    const int64_t a = rec_a(st, r), e = rec_e(st, r);
                    hashes[(size_t)r] = hash_bytes(st->buf + a, (size_t)(e - a));
                }
            }
        });
        size_t cap = 1024;
        while (cap < (size_t)n * 2) cap <<= 1;
        std::vector<int32_t> slots(cap, -1);   // unique-profile number
        const uint64_t mask = cap - 1;
        auto same = [&](int64_t r1, int64_t r2) {
            if (filter_active) {
                const size_t l1 = (size_t)(kept_ptr[(size_t)r1 + 1] - kept_ptr[(size_t)r1]), l2 = (size_t)(kept_ptr[(size_t)r2 + 1] - kept_ptr[(size_t)r2]);
                return l1 == l2 && (l1 == 0 || memcmp(kept.data() + kept_ptr[(size_t)r1], kept.data() + kept_ptr[(size_t)r2], l1 * sizeof(int32_t)) == 0);
            }
            const int64_t a1 = rec_a(st, r1), e1 = rec_e(st, r1), a2 = rec_a(st, r2), e2 = rec_e(st, r2);
            return e1 - a1 == e2 - a2 && (e1 == a1 || memcmp(st->buf + a1, st->buf + a2, (size_t)(e1 - a1)) == 0);
        };
        for (int64_t r = 0; r < n; ++r) {
            uint64_t i = hashes[(size_t)r] & mask;
            int32_t id = -1;
            while (slots[i] >= 0) {
                const int32_t u = slots[i];
                if (hashes[(size_t)uniq_rows[(size_t)u]] == hashes[(size_t)r] && same(uniq_rows[(size_t)u], r)) { id = u; break; }
                i = (i + 1) & mask;
            }
            if (id < 0) {
                id = (int32_t)uniq_rows.size();
                slots[i] = id;
                uniq_rows.push_back(r);
            }
            st->codes[(size_t)r] = id;
        }
    }
    lap("dedup");
    const int64_t nu = (int64_t)uniq_rows.size();
    st->first_seq.resize((size_t)nu);
    // ---- token CSR of the unique profiles (breakfast.py:199-213).  The reference numbers its vocabulary by first
    // appearance; nothing downstream depends on that order (distances are symmetric in the columns), so the kept tokens
    // are numbered densely in interning order instead, which needs no sequential sweep over the 10^8 occurrences:
    // pass 1 (parallel) marks the tokens in use and measures every row, pass 2 (parallel, below) fills the arrays
    std::vector<int32_t> vocab_of((size_t)n_distinct, -1);
    st->u_ptr.assign((size_t)nu + 1, 0);
    st->s_off.assign((size_t)nu + 1, 0);
    const int64_t sep_len = (int64_t)st->sep.size();
    {
        std::vector<std::vector<uint8_t>> used_by((size_t)T);
        parallel_ranges(nu, T, [&](int t, int64_t lo, int64_t hi) {
            std::vector<uint8_t>& used = used_by[(size_t)t];
            used.assign((size_t)n_distinct, 0);
            for (int64_t u = lo; u < hi; ++u) {
                const int64_t r = uniq_rows[(size_t)u];
                st->first_seq[(size_t)u] = (int32_t)r;
                int64_t bytes = 0;
                for (int64_t k = kept_ptr[(size_t)r]; k < kept_ptr[(size_t)r + 1]; ++k) {
                    const int32_t tk = kept[(size_t)k];
                    used[(size_t)tk] = 1;
                    bytes += st->tokens.len[(size_t)tk];
                }
                const int64_t cnt = kept_ptr[(size_t)r + 1] - kept_ptr[(size_t)r];
                st->u_ptr[(size_t)u + 1] = cnt;                                  // lengths for now
                if (filter_active) bytes += cnt > 1 ? (cnt - 1) * sep_len : 0;
                else bytes = rec_e(st, r) - rec_a(st, r);
                st->s_off[(size_t)u + 1] = bytes;
            }
        });
        for (int32_t tk = 0; tk < n_distinct; ++tk) {
            bool any = false;
            for (int t = 0; t < T && !any; ++t) any = !used_by[(size_t)t].empty() && used_by[(size_t)t][(size_t)tk];
            if (any) vocab_of[(size_t)tk] = st->n_vocab++;
        }
        for (int64_t u = 0; u < nu; ++u) {
            st->u_ptr[(size_t)u + 1] += st->u_ptr[(size_t)u];
            st->s_off[(size_t)u + 1] += st->s_off[(size_t)u];
        }
    }
    st->u_idx.resize((size_t)st->u_ptr[(size_t)nu]);
    st->s_bytes.resize((size_t)st->s_off[(size_t)nu]);
    parallel_ranges(nu, T, [&](int, int64_t lo, int64_t hi) {
        for (int64_t u = lo; u < hi; ++u) {
            const int64_t r = uniq_rows[(size_t)u];
            int32_t* out = st->u_idx.data() + st->u_ptr[(size_t)u];
            char* sp = st->s_bytes.data() + st->s_off[(size_t)u];
            for (int64_t k = kept_ptr[(size_t)r]; k < kept_ptr[(size_t)r + 1]; ++k) {
                const int32_t t = kept[(size_t)k];
                *out++ = vocab_of[(size_t)t];
                if (filter_active) {   // the profile string the reference keeps in meta["feature"]: the kept tokens re-joined
                    if (k > kept_ptr[(size_t)r]) { memcpy(sp, st->sep.data(), (size_t)sep_len); sp += sep_len; }
                    memcpy(sp, st->tokens.arena.data() + st->tokens.off[(size_t)t], (size_t)st->tokens.len[(size_t)t]);
                    sp += st->tokens.len[(size_t)t];
                }
            }
            if (!filter_active) memcpy(sp, st->buf + rec_a(st, r), (size_t)(rec_e(st, r) - rec_a(st, r)));
        }
    });
    lap("token csr + strings");
    // ---- strictly binary rows: the k-th repeat (k >= 1) of a token inside a profile becomes its own column.  Every
    // occurrence gives exactly one column, so b_ptr = u_ptr: rows are sorted in parallel, and only the (rare) rows with a
    // repeated token go through the sequential pass that numbers the extra columns in order of first appearance
    st->n_cols = st->n_vocab;
    st->b_ptr = st->u_ptr;
    st->b_idx.resize(st->u_idx.size());
    std::vector<uint8_t> has_repeat((size_t)nu, 0);
    parallel_ranges(nu, T, [&](int, int64_t lo, int64_t hi) {
        for (int64_t u = lo; u < hi; ++u) {
            int32_t* row = st->b_idx.data() + st->b_ptr[(size_t)u];
            const int64_t len = st->b_ptr[(size_t)u + 1] - st->b_ptr[(size_t)u];
            if (len) memcpy(row, st->u_idx.data() + st->u_ptr[(size_t)u], (size_t)len * sizeof(int32_t));
            if (!std::is_sorted(row, row + len)) std::sort(row, row + len);
            for (int64_t k = 1; k < len; ++k)
                if (row[k] == row[k - 1]) { has_repeat[(size_t)u] = 1; break; }
        }
    });
    std::unordered_map<int64_t, int32_t> extra;
    std::vector<int32_t> row;
    for (int64_t u = 0; u < nu; ++u) {
        if (!has_repeat[(size_t)u]) continue;
        int32_t* dst = st->b_idx.data() + st->b_ptr[(size_t)u];
        const int64_t len = st->b_ptr[(size_t)u + 1] - st->b_ptr[(size_t)u];
        row.assign(dst, dst + len);   // sorted, with repeats
        int64_t w = 0;
        for (size_t k = 0; k < row.size();) {
            size_t e = k;
            while (e < row.size() && row[e] == row[k]) ++e;
            dst[w++] = row[k];
            for (size_t rep = 1; rep < e - k; ++rep) {
                const int64_t key = ((int64_t)row[k] << 32) | (int64_t)std::min<size_t>(rep, 0x7fffffff);
                auto it = extra.find(key);
                if (it == extra.end()) it = extra.emplace(key, st->n_cols++).first;
                dst[w++] = it->second;
            }
            k = e;
        }
        std::sort(dst, dst + len);
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
    // the big arrays (token and column indices, profile strings: hundreds of MB at 10^6 profiles) are copied by all
    // host threads, each a contiguous slice
    const int T = std::max(1, st->n_threads);
    auto cp = [T](void* dst, const void* src, size_t bytes) {
        if (!dst || !bytes) return;
        if (bytes < ((size_t)8 << 20) || T == 1) { memcpy(dst, src, bytes); return; }
        parallel_ranges((int64_t)bytes, T, [=](int, int64_t lo, int64_t hi) {
            memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, (size_t)(hi - lo));
        });
    };
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
