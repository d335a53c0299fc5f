// host_pairs.cu -- the reference's per-pair CPU helpers (storm.h:56-61, 207, 209-210, 220-221), restated as plain
// host code.  They answer ONE pair of lists / blocks / rows and need no device; the all-vs-all queries (contig.cu,
// sparse.cu) never call them.  They exist so that code written against the reference header compiles and links
// unchanged.  Every function returns the exact |a AND b| for sorted unique input: the reference's bitmap x list
// probe (storm.c:636,644: `a & b != 0` parses as `a & (b != 0)`, defect D1) is not reproduced.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr uint32_t BLOCK_BITS = STORM_DEFAULT_BLOCK_SIZE;
constexpr uint32_t BLOCK_WORDS = BLOCK_BITS / 64;

inline uint64_t probe_list_into_bitmap(const uint64_t* words, const uint16_t* list, uint32_t n) {
    uint64_t count = 0;
    for (uint32_t i = 0; i < n; ++i) count += (words[list[i] >> 6] >> (list[i] & 63)) & 1ull;
    return count;
}

uint64_t block_pair(const STORM_bitmap_t* a, const STORM_bitmap_t* b, const STORM_compute_func func) {
    if (a == nullptr || b == nullptr || a->id != b->id) return 0;                       // storm.c:575-581
    if (a->n_bitmap == 0 && b->n_bitmap == 0)                                            // list x list, :583-589
        return STORM_intersect_vector16_cardinality(a->scalar, b->scalar, a->n_scalar, b->n_scalar);
    if (a->n_bitmap && b->n_bitmap == 0) return probe_list_into_bitmap(a->data, b->scalar, b->n_scalar);   // :591-597
    if (a->n_bitmap == 0 && b->n_bitmap) return probe_list_into_bitmap(b->data, a->scalar, a->n_scalar);   // :599-605
    if (func) return (*func)(a->data, b->data, a->n_bitmap);                             // bitmap x bitmap, :607-611 / :650
    uint64_t count = 0;
    for (uint32_t k = 0; k < BLOCK_WORDS; ++k) count += (uint64_t)__builtin_popcountll(a->data[k] & b->data[k]);
    return count;
}

}  // namespace

extern "C" {

uint64_t STORM_intersect_vector16_cardinality(const uint16_t* STORM_RESTRICT v1, const uint16_t* STORM_RESTRICT v2,
                                              const uint32_t len1, const uint32_t len2) {
    // storm.c:4-73 compares 8 x 8 values per SSE4.2 string instruction and finishes with a scalar merge; the value
    // is that of the merge alone.  A short list against a long one is searched instead (galloping lower bound).
    if (v1 == nullptr || v2 == nullptr || len1 == 0 || len2 == 0) return 0;
    const uint16_t* a = v1; const uint16_t* b = v2;
    uint32_t na = len1, nb = len2;
    if (na > nb) { std::swap(a, b); std::swap(na, nb); }
    uint64_t count = 0;
    if ((uint64_t)na * 16 < nb) {
        const uint16_t* lo = b;
        const uint16_t* end = b + nb;
        for (uint32_t i = 0; i < na && lo < end; ++i) {
            lo = std::lower_bound(lo, end, a[i]);
            if (lo < end && *lo == a[i]) { ++count; ++lo; }
        }
        return count;
    }
    uint32_t i = 0, j = 0;
    while (i < na && j < nb) {
        const uint16_t x = a[i], y = b[j];
        count += x == y;
        i += x <= y;
        j += y <= x;
    }
    return count;
}

uint64_t STORM_intersect_vector32_unsafe(const uint32_t* STORM_RESTRICT v1, const uint32_t* STORM_RESTRICT v2,
                                         const uint32_t len1, const uint32_t len2, uint32_t* STORM_RESTRICT out) {
    if (out == nullptr || v1 == nullptr || v2 == nullptr || len1 == 0 || len2 == 0) return 0;   // storm.c:81-84
    uint64_t n = 0;
    uint32_t i = 0, j = 0;
    while (i < len1 && j < len2) {
        if (v1[i] < v2[j]) ++i;
        else if (v1[i] > v2[j]) ++j;
        else { out[n++] = i++; out[n++] = j++; }                                              // :97-99: index pairs
    }
    return n;
}

uint64_t STORM_intersect_bitmaps_scalar_list(const uint64_t* STORM_RESTRICT b1, const uint64_t* STORM_RESTRICT b2,
                                             const uint32_t* l1, const uint32_t* l2, const uint32_t n1, const uint32_t n2) {
    uint64_t count = 0;
    if (n1 < n2) { for (uint32_t i = 0; i < n1; ++i) count += (b2[l1[i] >> 6] >> (l1[i] & 63)) & 1ull; }   // storm.c:116-120
    else         { for (uint32_t i = 0; i < n2; ++i) count += (b1[l2[i] >> 6] >> (l2[i] & 63)) & 1ull; }   // :121-126
    return count;
}

int STORM_bitmap_add_with_scalar(STORM_bitmap_t* b, const uint32_t* values, const uint32_t n_values) {   // storm.c:467-519
    if (b == nullptr) return -1;
    if (values == nullptr) return -3;
    if (n_values == 0) return -4;
    if (b->data == nullptr) {
        void* p = nullptr;
        if (posix_memalign(&p, 64, BLOCK_WORDS * sizeof(uint64_t))) return -5;
        memset(p, 0, BLOCK_WORDS * sizeof(uint64_t));
        b->data = (uint64_t*)p;
        b->own_data = 1;
    }
    const uint32_t need = b->n_scalar + n_values;                        // room for every value being new (the reference checks
    if (b->scalar == nullptr || need > b->m_scalar) {                    // capacity before the loop only, and not against the need)
        const uint32_t cap = std::max<uint32_t>(256, need + (need >> 2));
        uint16_t* p = (uint16_t*)realloc(b->own_scalar ? b->scalar : nullptr, cap * sizeof(uint16_t));
        if (p == nullptr) return -5;
        if (!b->own_scalar && b->scalar) memcpy(p, b->scalar, b->n_scalar * sizeof(uint16_t));
        b->scalar = p; b->m_scalar = cap; b->own_scalar = 1;
    }
    b->n_bitmap = BLOCK_WORDS;
    b->n_scalar_set = 1;
    const uint32_t adjust = b->id * BLOCK_BITS;
    uint32_t n = b->n_scalar;
    for (uint32_t i = 0; i < n_values; ++i) {
        const uint32_t v = values[i] - adjust;
        if (v >= BLOCK_BITS) { b->n_scalar = n; return -5; }            // reference: assert, compiled out
        const uint64_t bit = 1ull << (v & 63);
        if ((b->data[v >> 6] & bit) == 0) {
            b->data[v >> 6] |= bit;
            b->scalar[n++] = (uint16_t)v;
            ++b->n_bits_set;
        }
    }
    b->n_scalar = n;
    return (int)n_values;
}

uint64_t STORM_bitmap_intersect_cardinality(STORM_bitmap_t* STORM_RESTRICT bitmap1, STORM_bitmap_t* STORM_RESTRICT bitmap2) {
    return block_pair(bitmap1, bitmap2, nullptr);                        // storm.c:571-614
}

uint64_t STORM_bitmap_intersect_cardinality_func(STORM_bitmap_t* STORM_RESTRICT bitmap1, STORM_bitmap_t* STORM_RESTRICT bitmap2,
                                                 const STORM_compute_func func) {
    return block_pair(bitmap1, bitmap2, func);                           // storm.c:618-656
}

uint64_t STORM_bitmap_cont_intersect_cardinality_premade(const STORM_bitmap_cont_t* STORM_RESTRICT r1,
                                                         const STORM_bitmap_cont_t* STORM_RESTRICT r2,
                                                         const STORM_compute_func func, uint32_t* out) {   // storm.c:790-814
    if (r1 == nullptr || r2 == nullptr || out == nullptr || r1->n_bitmaps == 0 || r2->n_bitmaps == 0) return 0;
    const uint64_t n = STORM_intersect_vector32_unsafe(r1->block_ids, r2->block_ids, r1->n_bitmaps, r2->n_bitmaps, out);
    uint64_t count = 0;
    for (uint64_t k = 0; k < n; k += 2) count += block_pair(&r1->bitmaps[out[k]], &r2->bitmaps[out[k + 1]], func);
    return count;
}

uint64_t STORM_bitmap_cont_intersect_cardinality(const STORM_bitmap_cont_t* STORM_RESTRICT r1,
                                                 const STORM_bitmap_cont_t* STORM_RESTRICT r2) {           // storm.c:761-788
    if (r1 == nullptr || r2 == nullptr || r1->n_bitmaps == 0 || r2->n_bitmaps == 0) return 0;
    const uint32_t cap = 2 * std::min(r1->n_bitmaps, r2->n_bitmaps);
    uint32_t* out = (uint32_t*)malloc(cap * sizeof(uint32_t));
    if (out == nullptr) return 0;
    const uint64_t count = STORM_bitmap_cont_intersect_cardinality_premade(r1, r2, nullptr, out);
    free(out);
    return count;
}

}  // extern "C"
