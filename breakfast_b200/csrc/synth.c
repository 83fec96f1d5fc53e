/*
 * synth.c — deterministic generator of SARS-CoV-2-shaped, lineage-structured mutation profiles
 * (SURVEY.md section 8(d)).  Test/benchmark input only; not part of the clustering path.
 *
 * Genome length 29903.  Eight clade roots with 40,50,...,110 substitutions at untrimmed positions
 * 265..29674.  Growth: pick a random clade, a uniformly random existing profile of it as parent,
 * apply k ~ Geometric(0.7) events: 88 % new substitution, 5 % reversion, 4 % deletion, 3 %
 * insertion.  A child is rejected when its substitution set (= the filtered profile under the
 * default --skip-del --skip-ins) was already emitted, so "unique" means unique after filtering.
 * Sequence multiplicity ~ Geometric(0.6) capped at 50 (or 1).  Output rows are shuffled.
 *
 * An event is one int32 code = pos * 1024 + slot:
 *   slot 0..3     substitution to "ACGT"[slot]   (slot != reference base)
 *   slot 4..33    deletion of length slot-3      (1..30)
 *   slot 34..117  insertion; slot-34 enumerates the 84 inserted strings of length 1..3 over ACGT
 * Codes of one profile are sorted ascending = sorted by position.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GENOME_LEN 29903
#define POS_LO 265
#define POS_HI 29674
#define N_CLADES 8

typedef struct {
    uint64_t s[4];
} rng_t;

static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t* r) { /* xoshiro256** */
    uint64_t* s = r->s;
    const uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static void rng_seed(rng_t* r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed);
}
static inline uint64_t rng_below(rng_t* r, uint64_t n) { /* unbiased enough for n << 2^64 */
    return (uint64_t)(((__uint128_t)rng_next(r) * n) >> 64);
}
static inline double rng_unit(rng_t* r) { return ((rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline int rng_geometric(rng_t* r, double p) { /* support 1,2,... */
    return 1 + (int)floor(log(rng_unit(r)) / log(1.0 - p));
}

int bfsynth_ref_base(int pos) { return (int)((((uint32_t)pos * 2654435761u) >> 7) & 3u); }

typedef struct {
    int64_t n;
    int64_t* indptr; /* n+1 */
    int32_t* codes;  /* nnz */
    int32_t* mult;   /* n */
} synth_t;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
/* commutative hash of the substitution set */
static uint64_t subs_hash(const int32_t* c, int len, int all_events) {
    uint64_t h = 0x1234567ull + (uint64_t)0;
    for (int i = 0; i < len; ++i)
        if (all_events || (c[i] & 1023) < 4) h += mix64((uint64_t)(uint32_t)c[i] + 0x9E3779B97F4A7C15ull);
    return h ? h : 1;
}

static int set_insert(uint64_t* table, uint64_t mask, uint64_t h) { /* 1 = new */
    uint64_t i = mix64(h) & mask;
    while (table[i]) {
        if (table[i] == h) return 0;
        i = (i + 1) & mask;
    }
    table[i] = h;
    return 1;
}

static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

static int has_sub_at(const int32_t* c, int len, int pos) {
    for (int i = 0; i < len; ++i)
        if ((c[i] >> 10) == pos && (c[i] & 1023) < 4) return 1;
    return 0;
}

/* flags: bit0 = multiplicities on, bit1 = uniqueness over all events (not just substitutions) */
void* bfsynth_create(int64_t n, uint64_t seed, int flags) {
    if (n < 0) return NULL;
    const int with_mult = flags & 1, all_events = (flags >> 1) & 1;
    rng_t rng;
    rng_seed(&rng, seed * 0x2545F4914F6CDD1Dull + 0x1234);
    synth_t* s = (synth_t*)calloc(1, sizeof *s);
    if (!s) return NULL;
    int64_t cap_codes = (n > 16 ? n : 16) * 96;
    int64_t* start = (int64_t*)malloc((size_t)(n + 1) * sizeof *start); /* generation-order offsets */
    int32_t* arena = (int32_t*)malloc((size_t)cap_codes * sizeof *arena);
    uint64_t tsize = 16;
    while (tsize < (uint64_t)n * 3 + 16) tsize <<= 1;
    uint64_t* table = (uint64_t*)calloc(tsize, sizeof *table);
    int32_t* clade_members[N_CLADES];
    int64_t clade_n[N_CLADES], clade_cap[N_CLADES];
    for (int c = 0; c < N_CLADES; ++c) {
        clade_cap[c] = n / N_CLADES + 64;
        clade_members[c] = (int32_t*)malloc((size_t)clade_cap[c] * sizeof(int32_t));
        clade_n[c] = 0;
    }
    int32_t scratch[4096];
    if (!start || !arena || !table) goto fail;
    for (int c = 0; c < N_CLADES; ++c) if (!clade_members[c]) goto fail;

    int64_t made = 0, used = 0;
    start[0] = 0;
    /* roots */
    for (int c = 0; c < N_CLADES && made < n; ++c) {
        int want = 40 + 10 * c, len = 0;
        while (len < want) {
            int pos = POS_LO + (int)rng_below(&rng, POS_HI - POS_LO + 1);
            if (has_sub_at(scratch, len, pos)) continue;
            int alt = (bfsynth_ref_base(pos) + 1 + (int)rng_below(&rng, 3)) & 3;
            scratch[len++] = pos * 1024 + alt;
        }
        qsort(scratch, (size_t)len, sizeof(int32_t), cmp_i32);
        if (!set_insert(table, tsize - 1, subs_hash(scratch, len, all_events))) { --c; continue; }
        memcpy(arena + used, scratch, (size_t)len * sizeof(int32_t));
        used += len;
        clade_members[c][clade_n[c]++] = (int32_t)made;
        start[++made] = used;
    }
    /* growth */
    while (made < n) {
        int c = (int)rng_below(&rng, N_CLADES);
        if (clade_n[c] == 0) continue;
        int64_t par = clade_members[c][rng_below(&rng, (uint64_t)clade_n[c])];
        int len = (int)(start[par + 1] - start[par]);
        if (len > 3900) continue;
        memcpy(scratch, arena + start[par], (size_t)len * sizeof(int32_t));
        int k = rng_geometric(&rng, 0.7);
        for (int e = 0; e < k; ++e) {
            double u = rng_unit(&rng);
            if (u < 0.88) { /* new substitution */
                int pos = POS_LO + (int)rng_below(&rng, POS_HI - POS_LO + 1);
                if (has_sub_at(scratch, len, pos)) { --e; continue; }
                int alt = (bfsynth_ref_base(pos) + 1 + (int)rng_below(&rng, 3)) & 3;
                scratch[len++] = pos * 1024 + alt;
            } else if (u < 0.93) { /* reversion: drop one substitution */
                int nsub = 0;
                for (int i = 0; i < len; ++i) nsub += (scratch[i] & 1023) < 4;
                if (!nsub) continue;
                int pick = (int)rng_below(&rng, (uint64_t)nsub);
                for (int i = 0; i < len; ++i)
                    if ((scratch[i] & 1023) < 4 && pick-- == 0) { scratch[i] = scratch[--len]; break; }
            } else if (u < 0.97) { /* deletion */
                int pos = POS_LO + (int)rng_below(&rng, POS_HI - POS_LO + 1);
                int32_t code = pos * 1024 + 4 + (int)rng_below(&rng, 30);
                int dup = 0;
                for (int i = 0; i < len; ++i) dup |= scratch[i] == code;
                if (!dup) scratch[len++] = code;
            } else { /* insertion */
                int pos = POS_LO + (int)rng_below(&rng, POS_HI - POS_LO + 1);
                int32_t code = pos * 1024 + 34 + (int)rng_below(&rng, 84);
                int dup = 0;
                for (int i = 0; i < len; ++i) dup |= scratch[i] == code;
                if (!dup) scratch[len++] = code;
            }
        }
        if (!set_insert(table, tsize - 1, subs_hash(scratch, len, all_events))) continue;
        qsort(scratch, (size_t)len, sizeof(int32_t), cmp_i32);
        if (used + len > cap_codes) {
            cap_codes = cap_codes * 2 + len;
            int32_t* na = (int32_t*)realloc(arena, (size_t)cap_codes * sizeof *arena);
            if (!na) goto fail;
            arena = na;
        }
        memcpy(arena + used, scratch, (size_t)len * sizeof(int32_t));
        used += len;
        if (clade_n[c] == clade_cap[c]) {
            clade_cap[c] *= 2;
            int32_t* nm = (int32_t*)realloc(clade_members[c], (size_t)clade_cap[c] * sizeof(int32_t));
            if (!nm) goto fail;
            clade_members[c] = nm;
        }
        clade_members[c][clade_n[c]++] = (int32_t)made;
        start[++made] = used;
    }
    /* shuffle row order (Fisher-Yates) and emit */
    {
        int64_t* order = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof *order);
        s->indptr = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
        s->codes = (int32_t*)malloc((size_t)(used > 0 ? used : 1) * sizeof(int32_t));
        s->mult = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
        if (!order || !s->indptr || !s->codes || !s->mult) { free(order); goto fail; }
        for (int64_t i = 0; i < n; ++i) order[i] = i;
        for (int64_t i = n - 1; i > 0; --i) {
            int64_t j = (int64_t)rng_below(&rng, (uint64_t)i + 1);
            int64_t t = order[i]; order[i] = order[j]; order[j] = t;
        }
        int64_t off = 0;
        s->indptr[0] = 0;
        for (int64_t i = 0; i < n; ++i) {
            int64_t g = order[i], len = start[g + 1] - start[g];
            memcpy(s->codes + off, arena + start[g], (size_t)len * sizeof(int32_t));
            off += len;
            s->indptr[i + 1] = off;
            int m = 1;
            if (with_mult) { m = rng_geometric(&rng, 0.6); if (m > 50) m = 50; }
            s->mult[i] = m;
        }
        free(order);
    }
    s->n = n;
    free(start); free(arena); free(table);
    for (int c = 0; c < N_CLADES; ++c) free(clade_members[c]);
    return s;
fail:
    free(start); free(arena); free(table);
    for (int c = 0; c < N_CLADES; ++c) free(clade_members[c]);
    if (s) { free(s->indptr); free(s->codes); free(s->mult); free(s); }
    return NULL;
}

int64_t bfsynth_nnz(const void* h) { const synth_t* s = (const synth_t*)h; return s->indptr[s->n]; }

void bfsynth_copy(const void* h, int64_t* indptr, int32_t* codes, int32_t* mult) {
    const synth_t* s = (const synth_t*)h;
    memcpy(indptr, s->indptr, (size_t)(s->n + 1) * sizeof(int64_t));
    memcpy(codes, s->codes, (size_t)s->indptr[s->n] * sizeof(int32_t));
    memcpy(mult, s->mult, (size_t)s->n * sizeof(int32_t));
}

void bfsynth_free(void* h) {
    synth_t* s = (synth_t*)h;
    if (!s) return;
    free(s->indptr); free(s->codes); free(s->mult); free(s);
}
