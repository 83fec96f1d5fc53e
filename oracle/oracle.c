/*
 * oracle.c — CPU restatement of the reference's distance-and-clustering core.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this; the product (breakfast_b200) never does.
 *
 * What it restates (reference file:line, /root/reference = rki-mf1/breakfast v0.4.6):
 *   - L1 distance of two sparse rows by a two-pointer merge over sorted column indices:
 *     scikit-learn 1.9.0 sklearn/metrics/_pairwise_fast.pyx:76-107 (_sparse_manhattan), called from
 *     src/breakfast/breakfast.py:261-267.  Rows here are strictly binary (the host thermometer-codes
 *     repeated tokens), so |x_k - y_k| is 1 exactly where the rows differ.
 *   - the cardinality pre-filter |card_a - card_b| <= max_dist: breakfast.py:250-254
 *     (np.isclose(card, q, atol=max_dist) for every distinct q, breakfast.py:314-318); it can never
 *     change the result because |A xor B| >= ||A| - |B||.
 *   - threshold d <= max_dist: breakfast.py:226-228 (_reduce_func).
 *   - graph + connected components: breakfast.py:93-113 (_to_graph/_to_edges) and networkx 3.6.1
 *     connected_components (networkx/algorithms/components/connected.py:18-90), restated as a
 *     union-find whose label is the smallest row index of the component.
 *
 * Plain, obviously-correct code; OpenMP only parallelises the outer row loop.
 * Parity pin: tests/test_oracle_golden.py checks it against the reference's own expected_clusters_*
 * fixtures and against outputs of the reference itself (tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* |A xor B| for two ascending index lists; stops counting once it exceeds `limit`
 * (returns limit + 1 then).  _pairwise_fast.pyx:83-105 without the early exit. */
static int64_t symdiff(const int32_t* a, int64_t na, const int32_t* b, int64_t nb, int64_t limit) {
    int64_t i = 0, j = 0, d = 0;
    while (i < na && j < nb) {
        if (a[i] == b[j]) { ++i; ++j; }
        else if (a[i] < b[j]) { ++i; if (++d > limit) return d; }
        else { ++j; if (++d > limit) return d; }
    }
    d += (na - i) + (nb - j);
    return d > limit ? limit + 1 : d;
}

int64_t orc_distance(const int64_t* indptr, const int32_t* indices, int64_t a, int64_t b) {
    return symdiff(indices + indptr[a], indptr[a + 1] - indptr[a], indices + indptr[b],
                   indptr[b + 1] - indptr[b], INT64_MAX - 1);
}

static int32_t uf_find(int32_t* p, int32_t v) {
    while (p[v] != v) { p[v] = p[p[v]]; v = p[v]; }
    return v;
}
static void uf_union(int32_t* p, int32_t a, int32_t b) {
    a = uf_find(p, a); b = uf_find(p, b);
    if (a == b) return;
    if (a < b) p[b] = a; else p[a] = b;
}

typedef struct { int32_t a, b; } edge_t;
typedef struct { edge_t* e; int64_t n, cap; } edge_vec;
static int push(edge_vec* v, int32_t a, int32_t b) {
    if (v->n == v->cap) {
        int64_t nc = v->cap ? v->cap * 2 : 1024;
        edge_t* ne = (edge_t*)realloc(v->e, (size_t)nc * sizeof *ne);
        if (!ne) return -1;
        v->e = ne; v->cap = nc;
    }
    v->e[v->n].a = a; v->e[v->n].b = b; ++v->n;
    return 0;
}

static int cmp_edge(const void* x, const void* y) {
    const edge_t* a = (const edge_t*)x; const edge_t* b = (const edge_t*)y;
    if (a->a != b->a) return a->a < b->a ? -1 : 1;
    return (a->b > b->b) - (a->b < b->b);
}

/* All unordered pairs {i<j} with |A_i xor A_j| <= max_dist, where i ranges over `queries`
 * (NULL = all rows) and j over all rows; each unordered pair once, sorted.
 * Returns the edge count (or -1 on allocation failure); *src_out / *dst_out are malloc'ed. */
int64_t orc_edges(const int64_t* indptr, const int32_t* indices, int64_t n, int32_t max_dist,
                  const int32_t* queries, int64_t n_queries, int32_t** src_out, int32_t** dst_out) {
    const int64_t nq = queries ? n_queries : n;
    char* is_query = (char*)calloc((size_t)(n > 0 ? n : 1), 1);
    if (!is_query) return -1;
    for (int64_t q = 0; q < nq; ++q) is_query[queries ? queries[q] : q] = 1;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    edge_vec* vecs = (edge_vec*)calloc((size_t)nthreads, sizeof *vecs);
    int failed = 0;
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        edge_vec* v = &vecs[t];
#pragma omp for schedule(dynamic, 16)
        for (int64_t q = 0; q < nq; ++q) {
            const int64_t i = queries ? queries[q] : q;
            const int64_t ci = indptr[i + 1] - indptr[i];
            for (int64_t j = 0; j < n; ++j) {
                if (j == i) continue;
                if (is_query[j] && j < i) continue; /* that pair is reported from j's side */
                const int64_t cj = indptr[j + 1] - indptr[j];
                const int64_t gap = ci > cj ? ci - cj : cj - ci;
                if (gap > max_dist) continue; /* breakfast.py:250-254 */
                if (symdiff(indices + indptr[i], ci, indices + indptr[j], cj, max_dist) <= max_dist) {
                    int32_t a = (int32_t)(i < j ? i : j), b = (int32_t)(i < j ? j : i);
                    if (push(v, a, b)) failed = 1;
                }
            }
        }
    }
    free(is_query);
    int64_t total = 0;
    for (int t = 0; t < nthreads; ++t) total += vecs[t].n;
    edge_t* all = (edge_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof *all);
    if (!all) failed = 1;
    if (!failed) {
        int64_t off = 0;
        for (int t = 0; t < nthreads; ++t) {
            memcpy(all + off, vecs[t].e, (size_t)vecs[t].n * sizeof *all);
            off += vecs[t].n;
        }
        qsort(all, (size_t)total, sizeof *all, cmp_edge);
    }
    for (int t = 0; t < nthreads; ++t) free(vecs[t].e);
    free(vecs);
    if (failed) { free(all); return -1; }
    int32_t* s = (int32_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof *s);
    int32_t* d = (int32_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof *d);
    if (!s || !d) { free(all); free(s); free(d); return -1; }
    for (int64_t k = 0; k < total; ++k) { s[k] = all[k].a; d[k] = all[k].b; }
    free(all);
    *src_out = s; *dst_out = d;
    return total;
}

void orc_free(void* p) { free(p); }

/* Connected components over explicit edges and member lists (every list is chained like
 * breakfast.py:103-113); labels_out[i] = smallest row index in i's component. */
int orc_components(int64_t n, const int32_t* src, const int32_t* dst, int64_t n_edges,
                   const int64_t* list_indptr, const int32_t* list_members, int64_t n_lists,
                   int32_t* labels_out) {
    int32_t* p = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof *p);
    if (!p) return -1;
    for (int64_t i = 0; i < n; ++i) p[i] = (int32_t)i;
    for (int64_t e = 0; e < n_edges; ++e) uf_union(p, src[e], dst[e]);
    for (int64_t l = 0; l < n_lists; ++l)
        for (int64_t k = list_indptr[l] + 1; k < list_indptr[l + 1]; ++k)
            uf_union(p, list_members[k - 1], list_members[k]); /* path: last -- current */
    for (int64_t i = 0; i < n; ++i) labels_out[i] = uf_find(p, (int32_t)i);
    free(p);
    return 0;
}

/* Full run: labels of the graph {d <= max_dist} over all rows; returns edge count or -1. */
int64_t orc_cluster(const int64_t* indptr, const int32_t* indices, int64_t n, int32_t max_dist,
                    int32_t* labels_out) {
    int32_t *s = NULL, *d = NULL;
    int64_t ne = orc_edges(indptr, indices, n, max_dist, NULL, 0, &s, &d);
    if (ne < 0) return -1;
    int rc = orc_components(n, s, d, ne, NULL, NULL, 0, labels_out);
    free(s); free(d);
    return rc ? -1 : ne;
}

/* ---- helpers of oracle/hashjoin.py (the second, hash-join oracle) ------------------------------------------- */

/* ok_out[k] = (|A_src[k] xor A_dst[k]| <= limit), for m pairs */
void orc_within_batch(const int64_t* indptr, const int32_t* indices, const int32_t* src, const int32_t* dst, int64_t m,
                      int64_t limit, unsigned char* ok_out) {
#pragma omp parallel for schedule(static, 1024)
    for (int64_t k = 0; k < m; ++k) {
        const int64_t a = src[k], b = dst[k];
        ok_out[k] = symdiff(indices + indptr[a], indptr[a + 1] - indptr[a], indices + indptr[b],
                            indptr[b + 1] - indptr[b], limit) <= limit;
    }
}

/* Candidates of the "superset by two columns" case: all (A, B) with H[A] - g[x] - g[y] == H[B] for some columns x < y
 * of A and |B| == |A| - 2.  g[] = per-entry column hashes (aligned with indices), row_hash[] = their row sums (mod 2^64).
 * sorted_hash[] / sorted_row[] = the rows ordered by hash; table = bitset over the low `table_bits` bits of every row
 * hash (rejects almost every key with one lookup).  Hash collisions may add candidates, never lose one; the caller
 * verifies.  Returns the candidate count (-1 on allocation failure); *a_out / *b_out are malloc'ed. */
int64_t orc_join_two_deletions(const int64_t* indptr, const uint64_t* g, const uint64_t* row_hash, int64_t n,
                               const uint64_t* sorted_hash, const int32_t* sorted_row, const uint64_t* table,
                               int32_t table_bits, int32_t** a_out, int32_t** b_out) {
    const uint64_t mask = (table_bits >= 64) ? ~0ull : ((1ull << table_bits) - 1ull);
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    edge_vec* vecs = (edge_vec*)calloc((size_t)nthreads, sizeof *vecs);
    if (!vecs) return -1;
    int failed = 0;
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        edge_vec* v = &vecs[t];
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n; ++r) {
            const int64_t b0 = indptr[r], len = indptr[r + 1] - b0;
            for (int64_t i = 0; i + 1 < len; ++i) {
                const uint64_t hi = row_hash[r] - g[b0 + i];
                for (int64_t j = i + 1; j < len; ++j) {
                    const uint64_t key = hi - g[b0 + j];
                    const uint64_t slot = key & mask;
                    if (!((table[slot >> 6] >> (slot & 63)) & 1ull)) continue;
                    int64_t lo = 0, up = n; /* first sorted position with hash >= key */
                    while (lo < up) {
                        const int64_t mid = (lo + up) >> 1;
                        if (sorted_hash[mid] < key) lo = mid + 1; else up = mid;
                    }
                    for (int64_t p = lo; p < n && sorted_hash[p] == key; ++p) {
                        const int64_t c = sorted_row[p];
                        if (indptr[c + 1] - indptr[c] == len - 2 && push(v, (int32_t)r, (int32_t)c)) failed = 1;
                    }
                }
            }
        }
    }
    int64_t total = 0;
    for (int t = 0; t < nthreads; ++t) total += vecs[t].n;
    int32_t* a = (int32_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof *a);
    int32_t* b = (int32_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof *b);
    if (!a || !b) failed = 1;
    if (!failed) {
        int64_t off = 0;
        for (int t = 0; t < nthreads; ++t)
            for (int64_t k = 0; k < vecs[t].n; ++k, ++off) { a[off] = vecs[t].e[k].a; b[off] = vecs[t].e[k].b; }
    }
    for (int t = 0; t < nthreads; ++t) free(vecs[t].e);
    free(vecs);
    if (failed) { free(a); free(b); return -1; }
    *a_out = a; *b_out = b;
    return total;
}
