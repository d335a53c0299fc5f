/*
 * ref_shim.c -- thin exports around the UNMODIFIED reference so that tests and
 * bench.py can drive it through ctypes.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is ours; it is compiled together with /root/reference/storm.c (the
 * sources stay where they lie, nothing is copied into the repo) by
 * oracle/Makefile into oracle/_ref/libstorm_ref.so.  It exists because every
 * libalgebra kernel and the kernel chooser are `static` in a header
 * (libalgebra.h:3094-3140), so they have no linkable symbol of their own.
 */
#include "storm.h"

/* The reference's per-pair kernel as chosen by its own run-time dispatch
 * (libalgebra.h:3094-3140) for rows of n_words words. */
uint64_t REF_intersect_count(const uint64_t* a, const uint64_t* b, size_t n_words) {
    return (*STORM_get_intersect_count_func(n_words))(a, b, n_words);
}

/* storm.c:132-150 with the reference's own kernel choice. */
uint64_t REF_wrapper_diag(uint32_t n_vectors, const uint64_t* vals, uint32_t n_words) {
    return STORM_wrapper_diag(n_vectors, vals, n_words, STORM_get_intersect_count_func(n_words));
}

/* storm.c:132-150 driven with the reference's own union / diff kernel choosers
 * (libalgebra.h:3142-3188, 3190-3236): op 0 intersect, 1 union, 2 diff. */
uint64_t REF_wrapper_diag_op(uint32_t n_vectors, const uint64_t* vals, uint32_t n_words, int op) {
    const STORM_compute_func f = op == 1 ? STORM_get_union_count_func(n_words)
                               : op == 2 ? STORM_get_diff_count_func(n_words)
                                         : STORM_get_intersect_count_func(n_words);
    return STORM_wrapper_diag(n_vectors, vals, n_words, f);
}
uint64_t REF_count_op(const uint64_t* a, const uint64_t* b, size_t n_words, int op) {
    const STORM_compute_func f = op == 1 ? STORM_get_union_count_func(n_words)
                               : op == 2 ? STORM_get_diff_count_func(n_words)
                                         : STORM_get_intersect_count_func(n_words);
    return (*f)(a, b, n_words);
}

/* storm.c:222-279. */
uint64_t REF_wrapper_diag_blocked(uint32_t n_vectors, const uint64_t* vals, uint32_t n_words, uint32_t bsize) {
    return STORM_wrapper_diag_blocked(n_vectors, vals, n_words, STORM_get_intersect_count_func(n_words), bsize);
}

/* Row-range slice of the upper triangle using the reference kernel per pair:
 * rows [i0,i1) against rows (i, n_vectors).  Lets bench.py time a bounded,
 * multi-threaded sample of a large workload with the reference's own kernel
 * (the row partition is ours -- the reference has no threading). */
uint64_t REF_diag_rows(uint64_t n_vectors, const uint64_t* vals, uint64_t n_words, uint64_t i0, uint64_t i1) {
    const STORM_compute_func f = STORM_get_intersect_count_func(n_words);
    uint64_t total = 0;
    for (uint64_t i = i0; i < i1; ++i)
        for (uint64_t j = i + 1; j < n_vectors; ++j)
            total += (*f)(vals + i * n_words, vals + j * n_words, n_words);
    return total;
}

/* Blocked rectangle rows [i0,i1) x [j0,j1) (all pairs, caller keeps i1<=j0),
 * walking bsize x bsize squares like storm.c:1212-1220. */
uint64_t REF_rect_blocked(const uint64_t* vals, uint64_t n_words,
                          uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, uint32_t bsize) {
    const STORM_compute_func f = STORM_get_intersect_count_func(n_words);
    uint64_t total = 0;
    if (bsize == 0) bsize = 1;
    for (uint64_t ib = i0; ib < i1; ib += bsize)
        for (uint64_t jb = j0; jb < j1; jb += bsize) {
            const uint64_t ie = ib + bsize < i1 ? ib + bsize : i1;
            const uint64_t je = jb + bsize < j1 ? jb + bsize : j1;
            for (uint64_t i = ib; i < ie; ++i)
                for (uint64_t j = jb; j < je; ++j)
                    total += (*f)(vals + i * n_words, vals + j * n_words, n_words);
        }
    return total;
}

/* libalgebra.h:378-426 feature word and the SIMD tier the chooser lands on. */
int REF_cpuid(void) { return STORM_get_cpuid(); }

const char* REF_kernel_name(size_t n_words) {
    const STORM_compute_func f = STORM_get_intersect_count_func(n_words);
#if defined(STORM_HAVE_AVX512)
    if (f == &STORM_intersect_count_avx512) return "avx512";
#endif
#if defined(STORM_HAVE_AVX2)
    if (f == &STORM_intersect_count_avx2) return "avx2";
    if (f == &STORM_intersect_count_lookup_avx2) return "avx2-lookup";
#endif
#if defined(STORM_HAVE_SSE42)
    if (f == &STORM_intersect_count_sse4) return "sse4";
#endif
    return "scalar";
}

uint32_t REF_contig_scalar_cutoff(const STORM_contiguous_t* c) { return c->scalar_cutoff; }
uint64_t REF_contig_n_rows(const STORM_contiguous_t* c) { return c->n_data; }
const uint64_t* REF_contig_data(const STORM_contiguous_t* c) { return c->data; }
uint32_t REF_storm_n_rows(const STORM_t* s) { return s->n_conts; }
