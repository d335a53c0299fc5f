/*
 * storm_oracle.c -- CPU restatement of the StormBitmaps all-vs-all
 * intersection-cardinality path.  TEST INFRASTRUCTURE ONLY (see storm_oracle.h).
 *
 * Parity status: PINNED against the unmodified reference compiled into
 * oracle/_ref/libstorm_ref.so and the committed fixtures in tests/golden/.
 *
 * Citations are file:line relative to /root/reference (StormBitmaps @ 2eae567,
 * libalgebra @ bff182e).  Deliberate differences from the reference, all of them
 * listed in SURVEY.md section 7.4:
 *   D2  position lists are stored correctly (reference corrupts them on growth)
 *   D5  all offsets are 64-bit
 *   D1  the bitmap x list probe is exact unless emulate_d1 is requested
 *   D8  free really frees
 */
#include "storm_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* per-pair kernel                                                            */
/* ------------------------------------------------------------------------- */

static inline uint64_t popcnt64(uint64_t x) { return (uint64_t)__builtin_popcountll(x); }

/* libalgebra.h:499-519 (STORM_intersect_count_unrolled) via :2985-2991. */
uint64_t orc_intersect_count(const uint64_t* a, const uint64_t* b, size_t n_words) {
    uint64_t cnt = 0;
    for (size_t k = 0; k < n_words; ++k) cnt += popcnt64(a[k] & b[k]);
    return cnt;
}

/* ------------------------------------------------------------------------- */
/* raw-buffer loops                                                           */
/* ------------------------------------------------------------------------- */

/* storm.c:132-150. */
uint64_t orc_wrapper_diag(uint64_t n_vectors, const uint64_t* vals, uint64_t n_words) {
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_vectors; ++i) {
        const uint64_t* ri = vals + i * n_words;
        for (uint64_t j = i + 1; j < n_vectors; ++j)
            total += orc_intersect_count(ri, vals + j * n_words, n_words);
    }
    return total;
}

uint64_t orc_rect_total(const uint64_t* vals, uint64_t n_words,
                        uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1) {
    uint64_t total = 0;
    for (uint64_t i = i0; i < i1; ++i) {
        uint64_t js = j0 > i + 1 ? j0 : i + 1;
        for (uint64_t j = js; j < j1; ++j)
            total += orc_intersect_count(vals + i * n_words, vals + j * n_words, n_words);
    }
    return total;
}

void orc_rect_counts(const uint64_t* vals, uint64_t n_words,
                     uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, uint32_t* out) {
    const uint64_t nj = j1 - j0;
    for (uint64_t i = i0; i < i1; ++i)
        for (uint64_t j = j0; j < j1; ++j)
            out[(i - i0) * nj + (j - j0)] = (j > i)
                ? (uint32_t)orc_intersect_count(vals + i * n_words, vals + j * n_words, n_words)
                : 0u;
}

/* libalgebra.h:2994-3000 / 3002-3008: the union and diff per-pair kernels. */
uint64_t orc_union_count(const uint64_t* a, const uint64_t* b, size_t n_words) {
    uint64_t count = 0;
    for (size_t k = 0; k < n_words; ++k) count += (uint64_t)__builtin_popcountll(a[k] | b[k]);
    return count;
}
uint64_t orc_diff_count(const uint64_t* a, const uint64_t* b, size_t n_words) {
    uint64_t count = 0;
    for (size_t k = 0; k < n_words; ++k) count += (uint64_t)__builtin_popcountll(a[k] ^ b[k]);
    return count;
}
static uint64_t orc_count_op(const uint64_t* a, const uint64_t* b, size_t n_words, int op) {
    return op == 1 ? orc_union_count(a, b, n_words) : op == 2 ? orc_diff_count(a, b, n_words)
                                                               : orc_intersect_count(a, b, n_words);
}
/* storm.c:132-150 with the kernel pointer f chosen by op. */
uint64_t orc_wrapper_diag_op(uint64_t n_vectors, const uint64_t* vals, uint64_t n_words, int op) {
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_vectors; ++i)
        for (uint64_t j = i + 1; j < n_vectors; ++j)
            total += orc_count_op(vals + i * n_words, vals + j * n_words, n_words, op);
    return total;
}
void orc_rect_counts_op(const uint64_t* vals, uint64_t n_words,
                        uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, int op, uint32_t* out) {
    const uint64_t nj = j1 - j0;
    for (uint64_t i = i0; i < i1; ++i)
        for (uint64_t j = j0; j < j1; ++j)
            out[(i - i0) * nj + (j - j0)] = (j > i)
                ? (uint32_t)orc_count_op(vals + i * n_words, vals + j * n_words, n_words, op) : 0u;
}
uint64_t orc_wrapper_square_op(uint64_t n1, const uint64_t* vals1, uint64_t n2,
                               const uint64_t* vals2, uint64_t n_words, int op) {
    uint64_t total = 0;
    for (uint64_t i = 0; i < n1; ++i)
        for (uint64_t j = 0; j < n2; ++j)
            total += orc_count_op(vals1 + i * n_words, vals2 + j * n_words, n_words, op);
    return total;
}

/* Documented intent of storm.c:153-171 (storm.h:72-76). */
uint64_t orc_wrapper_square(uint64_t n1, const uint64_t* vals1, uint64_t n2,
                            const uint64_t* vals2, uint64_t n_words) {
    uint64_t total = 0;
    for (uint64_t i = 0; i < n1; ++i)
        for (uint64_t j = 0; j < n2; ++j)
            total += orc_intersect_count(vals1 + i * n_words, vals2 + j * n_words, n_words);
    return total;
}

/* ------------------------------------------------------------------------- */
/* closed-form checksum (independent of the pair loop)                        */
/* ------------------------------------------------------------------------- */

static void column_counts(const uint64_t* vals, uint64_t n_words, uint64_t r0, uint64_t r1,
                          uint32_t* col /* n_words*64, zeroed */) {
    for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t* row = vals + r * n_words;
        for (uint64_t w = 0; w < n_words; ++w) {
            uint64_t x = row[w];
            while (x) {
                col[w * 64 + (uint64_t)__builtin_ctzll(x)]++;
                x &= x - 1;
            }
        }
    }
}

uint64_t orc_colcount_total(uint64_t n_vectors, const uint64_t* vals, uint64_t n_words) {
    uint32_t* col = (uint32_t*)calloc(n_words * 64, sizeof(uint32_t));
    if (!col) return UINT64_MAX;
    column_counts(vals, n_words, 0, n_vectors, col);
    uint64_t total = 0;
    for (uint64_t k = 0; k < n_words * 64; ++k) {
        uint64_t c = col[k];
        total += c * (c - 1) / 2;
    }
    free(col);
    return total;
}

uint64_t orc_colcount_rect(const uint64_t* vals, uint64_t n_words,
                           uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1) {
    uint32_t* ci = (uint32_t*)calloc(n_words * 64, sizeof(uint32_t));
    uint32_t* cj = (uint32_t*)calloc(n_words * 64, sizeof(uint32_t));
    if (!ci || !cj) { free(ci); free(cj); return UINT64_MAX; }
    column_counts(vals, n_words, i0, i1, ci);
    column_counts(vals, n_words, j0, j1, cj);
    uint64_t total = 0;
    for (uint64_t k = 0; k < n_words * 64; ++k) total += (uint64_t)ci[k] * cj[k];
    free(ci); free(cj);
    return total;
}

/* ------------------------------------------------------------------------- */
/* contiguous model                                                           */
/* ------------------------------------------------------------------------- */

/* storm.c:1001-1018. */
orc_contig_t* orc_contig_new(size_t vector_length) {
    orc_contig_t* c = (orc_contig_t*)calloc(1, sizeof(orc_contig_t));
    if (!c) return NULL;
    c->vector_length = vector_length;
    c->n_words = (uint32_t)((vector_length + 63) / 64);            /* :1013 ceil(M/64.0) */
    uint64_t q = vector_length / 200;                              /* :1016 */
    c->scalar_cutoff = (uint32_t)(q > 200 ? 200 : q);
    return c;
}

void orc_contig_free(orc_contig_t* c) {
    if (!c) return;
    free(c->data); free(c->positions); free(c->pos_offset); free(c->n_set); free(c);
}

/* storm.c:1031-1137. */
int orc_contig_add(orc_contig_t* c, const uint32_t* values, uint32_t n_values) {
    if (c == NULL) return -1;                                      /* :1032 */
    if (values == NULL) return -2;                                 /* :1033 */
    if (n_values == 0) return 0;                                   /* :1034 -- no row appended (D7) */

    if (c->n_rows == c->cap_rows) {                                /* :1045-1056,1078-1100 (+512 rows) */
        uint64_t cap = c->cap_rows + 512;
        uint64_t* d = (uint64_t*)realloc(c->data, cap * c->n_words * sizeof(uint64_t));
        uint64_t* o = (uint64_t*)realloc(c->pos_offset, cap * sizeof(uint64_t));
        uint32_t* n = (uint32_t*)realloc(c->n_set, cap * sizeof(uint32_t));
        if (!d || !o || !n) return -3;
        memset(d + c->cap_rows * c->n_words, 0, (cap - c->cap_rows) * c->n_words * sizeof(uint64_t));
        c->data = d; c->pos_offset = o; c->n_set = n; c->cap_rows = cap;
    }
    if (c->n_pos + n_values > c->cap_pos) {                        /* :1060-1073 (done right, D2) */
        uint64_t cap = c->cap_pos + (5ull * n_values < 65535 ? 65535 : 5ull * n_values);
        uint32_t* p = (uint32_t*)realloc(c->positions, cap * sizeof(uint32_t));
        if (!p) return -3;
        c->positions = p; c->cap_pos = cap;
    }

    uint64_t* row = c->data + c->n_rows * c->n_words;
    uint32_t used = n_values;
    for (uint32_t i = 0; i < n_values; ++i) {                      /* :1103-1115 */
        if (i != 0 && values[i] == values[i - 1]) { --used; continue; }
        row[values[i] / 64] |= 1ull << (values[i] % 64);
    }
    c->pos_offset[c->n_rows] = c->n_pos;
    if (used < c->scalar_cutoff) {                                 /* :1119-1129 */
        uint64_t w = c->n_pos;
        for (uint32_t i = 0; i < n_values; ++i) {
            if (i != 0 && values[i] == values[i - 1]) continue;
            c->positions[w++] = values[i];
        }
        c->n_pos = w;
    }
    c->n_set[c->n_rows] = used;                                    /* :1132-1133 */
    ++c->n_rows;                                                   /* :1134 */
    return (int)n_values;                                          /* :1136 */
}

/* storm.c:1139-1147. */
int orc_contig_clear(orc_contig_t* c) {
    if (c == NULL) return -1;
    if (c->data == NULL) return 0;
    memset(c->data, 0, c->cap_rows * c->n_words * sizeof(uint64_t));
    c->n_rows = 0; c->n_pos = 0;
    return 1;
}

/* storm.c:108-129 (MOD(x) behaves as x & 63 on x86-64, D10). */
uint64_t orc_probe_list(const uint64_t* b1, const uint64_t* b2,
                        const uint32_t* l1, const uint32_t* l2, uint32_t n1, uint32_t n2) {
    uint64_t count = 0;
    if (n1 < n2) {
        for (uint32_t i = 0; i < n1; ++i) count += (b2[l1[i] >> 6] >> (l1[i] & 63)) & 1u;
    } else {
        for (uint32_t i = 0; i < n2; ++i) count += (b1[l2[i] >> 6] >> (l2[i] & 63)) & 1u;
    }
    return count;
}

static int contig_has_sparse_row(const orc_contig_t* c) {          /* :1151-1162 */
    for (uint64_t i = 0; i < c->n_rows; ++i)
        if (c->n_set[i] < c->scalar_cutoff) return 1;
    return 0;
}

static inline uint64_t contig_pair_dense(const orc_contig_t* c, uint64_t i, uint64_t j) {
    return orc_intersect_count(c->data + i * c->n_words, c->data + j * c->n_words, c->n_words);
}

static inline uint64_t contig_pair_list(const orc_contig_t* c, uint64_t i, uint64_t j) { /* :1253-1258 */
    if (c->n_set[i] < c->scalar_cutoff || c->n_set[j] < c->scalar_cutoff)
        return orc_probe_list(c->data + i * c->n_words, c->data + j * c->n_words,
                              c->positions + c->pos_offset[i], c->positions + c->pos_offset[j],
                              c->n_set[i], c->n_set[j]);
    return contig_pair_dense(c, i, j);
}

typedef uint64_t (*contig_pair_fn)(const orc_contig_t*, uint64_t, uint64_t);

static uint64_t contig_loop_plain(const orc_contig_t* c, contig_pair_fn f) {     /* :1164-1170 */
    uint64_t total = 0;
    for (uint64_t i = 0; i < c->n_rows; ++i)
        for (uint64_t j = i + 1; j < c->n_rows; ++j) total += f(c, i, j);
    return total;
}

/* The diag / square / residual / tail walk of storm.c:1199-1238. */
static uint64_t contig_loop_blocked(const orc_contig_t* c, uint32_t bsize, contig_pair_fn f) {
    uint64_t count = 0, i = 0;
    const uint64_t n = c->n_rows;
    for (; i + bsize <= n; i += bsize) {
        for (uint64_t j = 0; j < bsize; ++j)                       /* diagonal triangle */
            for (uint64_t jj = j + 1; jj < bsize; ++jj) count += f(c, i + j, i + jj);
        uint64_t j = i + bsize;
        for (; j + bsize <= n; j += bsize)                         /* full squares */
            for (uint64_t ii = 0; ii < bsize; ++ii)
                for (uint64_t jj = 0; jj < bsize; ++jj) count += f(c, i + ii, j + jj);
        for (; j < n; ++j)                                         /* residual columns */
            for (uint64_t jj = 0; jj < bsize; ++jj) count += f(c, i + jj, j);
    }
    for (; i < n; ++i)                                             /* tail rows */
        for (uint64_t j = i + 1; j < n; ++j) count += f(c, i, j);
    return count;
}

uint64_t orc_contig_pairw_list(const orc_contig_t* c) {           /* :1243-1263 */
    if (c == NULL) return (uint64_t)-1;
    if (c->positions == NULL && c->n_rows) return (uint64_t)-2;
    return contig_loop_plain(c, contig_pair_list);
}

uint64_t orc_contig_pairw_blocked_list(const orc_contig_t* c, uint32_t bsize) { /* :1265-1347 */
    if (c == NULL) return (uint64_t)-1;
    if (c->positions == NULL && c->n_rows) return (uint64_t)-2;
    if (bsize <= 2) return orc_contig_pairw_list(c);
    return contig_loop_blocked(c, bsize, contig_pair_list);
}

uint64_t orc_contig_pairw(const orc_contig_t* c) {                /* :1149-1173 */
    if (c == NULL) return (uint64_t)-1;
    if (contig_has_sparse_row(c)) return orc_contig_pairw_list(c);
    return contig_loop_plain(c, contig_pair_dense);
}

uint64_t orc_contig_pairw_blocked(const orc_contig_t* c, uint32_t bsize) { /* :1175-1241 */
    if (c == NULL) return (uint64_t)-1;
    if (contig_has_sparse_row(c)) return orc_contig_pairw_blocked_list(c, bsize);
    if (bsize <= 2) return orc_contig_pairw(c);
    return contig_loop_blocked(c, bsize, contig_pair_dense);
}

/* ------------------------------------------------------------------------- */
/* sparse STORM_t model                                                       */
/* ------------------------------------------------------------------------- */

orc_storm_t* orc_storm_new(void) { return (orc_storm_t*)calloc(1, sizeof(orc_storm_t)); }

static void row_release(orc_row_t* r) {
    for (uint32_t b = 0; b < r->n_blocks; ++b) { free(r->blocks[b].words); free(r->blocks[b].list); }
    free(r->blocks);
    r->blocks = NULL; r->n_blocks = r->cap_blocks = 0;
}

void orc_storm_free(orc_storm_t* s) {
    if (!s) return;
    for (uint32_t i = 0; i < s->n_rows; ++i) row_release(&s->rows[i]);
    free(s->rows); free(s);
}

int orc_storm_clear(orc_storm_t* s) {                              /* storm.c:868-875 */
    if (s == NULL) return -1;
    for (uint32_t i = 0; i < s->n_rows; ++i) row_release(&s->rows[i]);
    s->n_rows = 0;
    return 1;
}

/* storm.c:692-758: split a sorted row at multiples of 65536; a block with fewer
 * than 4096 values becomes a u16 list (:745-746, builder :521-558, no de-dup),
 * otherwise an 8 KiB bitmap (:747-748, builder :442-465). */
static int row_build(orc_row_t* r, const uint32_t* values, uint32_t n_values) {
    uint32_t start = 0;
    while (start < n_values) {
        const uint32_t id = values[start] / ORC_BLOCK_BITS;
        uint32_t stop = start;
        while (stop < n_values && values[stop] / ORC_BLOCK_BITS == id) ++stop;
        if (r->n_blocks == r->cap_blocks) {
            uint32_t cap = r->cap_blocks ? r->cap_blocks + 8 : 2;  /* :699,:728-735 */
            orc_block_t* nb = (orc_block_t*)realloc(r->blocks, cap * sizeof(orc_block_t));
            if (!nb) return -3;
            r->blocks = nb; r->cap_blocks = cap;
        }
        orc_block_t* blk = &r->blocks[r->n_blocks++];
        memset(blk, 0, sizeof(*blk));
        blk->id = id;
        blk->n_values = stop - start;
        const uint32_t adjust = id * ORC_BLOCK_BITS;
        if (stop - start < ORC_LIST_THRESHOLD) {
            blk->list = (uint16_t*)malloc((stop - start) * sizeof(uint16_t));
            if (!blk->list) return -3;
            for (uint32_t i = start; i < stop; ++i) blk->list[i - start] = (uint16_t)(values[i] - adjust);
        } else {
            blk->is_bitmap = 1;
            blk->words = (uint64_t*)calloc(ORC_BLOCK_WORDS, sizeof(uint64_t));
            if (!blk->words) return -3;
            for (uint32_t i = start; i < stop; ++i) {
                uint32_t v = values[i] - adjust;
                blk->words[v / 64] |= 1ull << (v % 64);
            }
        }
        start = stop;
    }
    return 1;
}

/* storm.c:844-866: a row is appended even when n_values == 0 (D7). */
int orc_storm_add(orc_storm_t* s, const uint32_t* values, uint32_t n_values) {
    if (s == NULL) return -1;
    if (s->n_rows == s->cap_rows) {
        uint32_t cap = s->cap_rows + 1024;                         /* :847-861 */
        orc_row_t* nr = (orc_row_t*)realloc(s->rows, cap * sizeof(orc_row_t));
        if (!nr) return -3;
        memset(nr + s->cap_rows, 0, (cap - s->cap_rows) * sizeof(orc_row_t));
        s->rows = nr; s->cap_rows = cap;
    }
    orc_row_t* r = &s->rows[s->n_rows++];
    memset(r, 0, sizeof(*r));
    if (values != NULL && n_values != 0) row_build(r, values, n_values);
    return 1;                                                      /* :865 */
}

/* storm.c:59-71 is the scalar form; the pcmpestrm body (:15-57) counts the same
 * matches for sorted unique inputs. */
uint64_t orc_intersect_u16(const uint16_t* v1, const uint16_t* v2, uint32_t n1, uint32_t n2) {
    uint64_t count = 0;
    uint32_t a = 0, b = 0;
    while (a < n1 && b < n2) {
        if (v1[a] < v2[b]) ++a;
        else if (v2[b] < v1[a]) ++b;
        else { ++count; ++a; ++b; }
    }
    return count;
}

/* storm.c:75-106. */
uint64_t orc_intersect_u32_pairs(const uint32_t* v1, const uint32_t* v2,
                                 uint32_t n1, uint32_t n2, uint32_t* out) {
    if (!out || !v1 || !v2 || n1 == 0 || n2 == 0) return 0;
    uint64_t answer = 0;
    uint32_t a = 0, b = 0;
    while (a < n1 && b < n2) {
        if (v1[a] < v2[b]) ++a;
        else if (v1[a] > v2[b]) ++b;
        else { out[answer++] = a++; out[answer++] = b++; }
    }
    return answer;
}

/* storm.c:618-656: the 4-way block x block dispatch. */
static uint64_t block_pair(const orc_block_t* x, const orc_block_t* y, int emulate_d1) {
    if (x->id != y->id) return 0;                                  /* :625-626 */
    if (!x->is_bitmap && !y->is_bitmap)                            /* :628-630 */
        return orc_intersect_u16(x->list, y->list, x->n_values, y->n_values);
    if (x->is_bitmap && y->is_bitmap)                              /* :648-650 */
        return orc_intersect_count(x->words, y->words, ORC_BLOCK_WORDS);
    const orc_block_t* bm = x->is_bitmap ? x : y;                  /* :632-646 */
    const orc_block_t* ls = x->is_bitmap ? y : x;
    uint64_t count = 0;
    for (uint32_t i = 0; i < ls->n_values; ++i) {
        const uint64_t word = bm->words[ls->list[i] / 64];
        /* D1: `word & (1ULL << k) != 0` parses as `word & ((1ULL << k) != 0)` == word & 1 */
        count += emulate_d1 ? (word & 1u) : ((word >> (ls->list[i] % 64)) & 1u);
    }
    return count;
}

/* storm.c:790-814: merge the two block-id lists, then sum the shared blocks. */
uint64_t orc_storm_row_pair(const orc_row_t* a, const orc_row_t* b, int emulate_d1) {
    if (!a || !b || a->n_blocks == 0 || b->n_blocks == 0) return 0;
    uint64_t count = 0;
    uint32_t x = 0, y = 0;
    while (x < a->n_blocks && y < b->n_blocks) {
        if (a->blocks[x].id < b->blocks[y].id) ++x;
        else if (a->blocks[x].id > b->blocks[y].id) ++y;
        else { count += block_pair(&a->blocks[x], &b->blocks[y], emulate_d1); ++x; ++y; }
    }
    return count;
}

uint64_t orc_storm_pairw(const orc_storm_t* s, int emulate_d1) {  /* storm.c:877-895 */
    if (s == NULL) return (uint64_t)-1;
    uint64_t total = 0;
    for (uint32_t i = 0; i < s->n_rows; ++i)
        for (uint32_t j = i + 1; j < s->n_rows; ++j)
            total += orc_storm_row_pair(&s->rows[i], &s->rows[j], emulate_d1);
    return total;
}

/* storm.c:372-381 per block, :384-394 per row, :963-973 total. */
static uint32_t row_serialized_size(const orc_row_t* r) {
    uint32_t total = 0;
    for (uint32_t b = 0; b < r->n_blocks; ++b) {
        const orc_block_t* k = &r->blocks[b];
        total += k->is_bitmap ? 8u * ORC_BLOCK_WORDS : 2u * k->n_values;
        total += 16u;
    }
    return total + 4u * r->n_blocks + 12u;
}

uint64_t orc_storm_serialized_size(const orc_storm_t* s) {
    if (s == NULL) return 0;
    uint64_t tot = 0;
    for (uint32_t i = 0; i < s->n_rows; ++i) tot += row_serialized_size(&s->rows[i]);
    return tot + 8u;
}

uint32_t orc_storm_auto_bsize(const orc_storm_t* s) {             /* storm.c:903-914 */
    if (s == NULL || s->n_rows == 0) return 5;
    uint64_t tot = 0;
    for (uint32_t i = 0; i < s->n_rows; ++i) tot += row_serialized_size(&s->rows[i]);
    uint32_t average = (uint32_t)(tot / s->n_rows);
    double q = ORC_CACHE_BLOCK_BYTES / (double)average;
    uint32_t b = (uint32_t)q;
    if ((double)b < q) ++b;                                        /* ceil */
    return b < 5 ? 5 : b;
}

uint64_t orc_storm_pairw_blocked(const orc_storm_t* s, uint32_t bsize, int emulate_d1) { /* :897-961 */
    if (s == NULL) return (uint64_t)-1;
    if (bsize == 0) bsize = orc_storm_auto_bsize(s);
    if (bsize < 5) bsize = 5;
    uint64_t count = 0;
    uint32_t i = 0;
    const uint32_t n = s->n_rows;
    for (; i + bsize <= n; i += bsize) {
        for (uint32_t j = 0; j < bsize; ++j)
            for (uint32_t jj = j + 1; jj < bsize; ++jj)
                count += orc_storm_row_pair(&s->rows[i + j], &s->rows[i + jj], emulate_d1);
        uint32_t j = i + bsize;
        for (; j + bsize <= n; j += bsize)
            for (uint32_t ii = 0; ii < bsize; ++ii)
                for (uint32_t jj = 0; jj < bsize; ++jj)
                    count += orc_storm_row_pair(&s->rows[i + ii], &s->rows[j + jj], emulate_d1);
        for (; j < n; ++j)
            for (uint32_t jj = 0; jj < bsize; ++jj)
                count += orc_storm_row_pair(&s->rows[i + jj], &s->rows[j], emulate_d1);
    }
    for (; i < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j)
            count += orc_storm_row_pair(&s->rows[i], &s->rows[j], emulate_d1);
    return count;
}

/* ------------------------------------------------------------------------- */
/* deterministic synthetic inputs                                             */
/* ------------------------------------------------------------------------- */

uint64_t orc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline uint64_t row_key(uint64_t seed, uint64_t row) {
    return orc_splitmix64(orc_splitmix64(seed) ^ (row * 0xD1342543DE82EF95ull));
}

/* benchmark.cpp:753,772: uniform position in [0, M-1].  Multiply-high range
 * reduction of a 32-bit hash instead of std::uniform_int_distribution, which is
 * implementation defined. */
uint32_t orc_draw_position(uint64_t seed, uint64_t row, uint64_t draw, uint32_t M) {
    uint64_t z = orc_splitmix64(row_key(seed, row) + draw * 0x9E3779B97F4A7C15ull);
    return (uint32_t)(((z >> 32) * (uint64_t)M) >> 32);
}

static int cmp_u32(const void* a, const void* b) {
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return (x > y) - (x < y);
}

/* benchmark.cpp:770-781 (draw with replacement, keep unique, sort). */
uint32_t orc_gen_row_positions(uint64_t seed, uint64_t row, uint32_t n_draws, uint32_t M, uint32_t* out) {
    for (uint32_t t = 0; t < n_draws; ++t) out[t] = orc_draw_position(seed, row, t, M);
    qsort(out, n_draws, sizeof(uint32_t), cmp_u32);
    uint32_t n = 0;
    for (uint32_t t = 0; t < n_draws; ++t)
        if (n == 0 || out[t] != out[n - 1]) out[n++] = out[t];
    return n;
}

void orc_gen_dense_uniform(uint64_t seed, uint64_t row0, uint64_t n_rows, uint32_t n_draws,
                           uint32_t M, uint64_t* vals, uint64_t n_words) {
    for (uint64_t r = 0; r < n_rows; ++r) {
        uint64_t* row = vals + r * n_words;
        for (uint32_t t = 0; t < n_draws; ++t) {
            uint32_t p = orc_draw_position(seed, row0 + r, t, M);
            row[p >> 6] |= 1ull << (p & 63);
        }
    }
}

static inline uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

uint32_t orc_geno_threshold(uint64_t seed, uint64_t row) {
    uint64_t u24 = row_key(seed ^ 0x47454E4Full, row) >> 40;       /* 24 uniform bits */
    uint64_t thr = (u24 * u24) >> 17;                              /* 0.5 * u^2 * 2^32 */
    return (uint32_t)(thr < 21474836ull ? 21474836ull : thr);      /* floor at p = 0.005 */
}

void orc_gen_dense_geno(uint64_t seed, uint64_t row0, uint64_t n_rows, uint32_t M,
                        uint64_t* vals, uint64_t n_words) {
    for (uint64_t r = 0; r < n_rows; ++r) {
        const uint32_t thr = orc_geno_threshold(seed, row0 + r);
        const uint32_t key = (uint32_t)row_key(seed, row0 + r);
        uint64_t* row = vals + r * n_words;
        for (uint32_t k = 0; k < M; ++k)
            if (fmix32(k * 0x9E3779B1u + key) < thr) row[k >> 6] |= 1ull << (k & 63);
    }
}
