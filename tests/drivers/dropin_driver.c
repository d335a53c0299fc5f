/* dropin_driver.c -- a C99 caller of storm.h, written the way the reference's only caller uses the API
 * (benchmark.cpp:710-711 new, :579-580 / :794-795 add with a reused buffer, :898 / :910 queries and their list
 * variants :835 / :845, :738-739 clear and reuse, :1055-1056 free).  It includes nothing but <storm.h> and links against whatever provides the symbols:
 *     gcc -std=c99 -I include tests/drivers/dropin_driver.c -L stormbitmaps_b200 -lstorm_b200      (this repo)
 *     gcc -std=c99 -I /root/reference tests/drivers/dropin_driver.c /root/reference/storm.c        (the reference)
 * Usage: dropin_driver M N draws seed   -> one line "key=value ..." with every total it computed.
 * Rows are generated with a private LCG (positions sorted and unique, as the callers guarantee, benchmark.cpp:571-572). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "storm.h"

static uint64_t lcg(uint64_t* s) { *s = *s * 6364136223846793005ull + 1442695040888963407ull; return *s >> 33; }
static int cmp_u32(const void* a, const void* b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return (x > y) - (x < y); }

static uint32_t gen_row(uint32_t* buf, uint32_t M, uint32_t draws, uint64_t* state) {
    uint32_t n = 0, i;
    for (i = 0; i < draws; ++i) buf[i] = (uint32_t)(lcg(state) % M);
    qsort(buf, draws, sizeof(uint32_t), cmp_u32);
    for (i = 0; i < draws; ++i) if (n == 0 || buf[n - 1] != buf[i]) buf[n++] = buf[i];
    return n;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s M N draws seed\n", argv[0]); return 2; }
    const uint32_t M = (uint32_t)strtoul(argv[1], 0, 10), N = (uint32_t)strtoul(argv[2], 0, 10);
    const uint32_t draws = (uint32_t)strtoul(argv[3], 0, 10);
    const uint64_t seed = strtoull(argv[4], 0, 10);
    /* one buffer, reused per row; one spare element because the reference reads values[n_values] (storm.c:723, D9) */
    uint32_t* buf = (uint32_t*)calloc((size_t)draws + 1, sizeof(uint32_t));
    uint64_t state, naive = 0;
    uint32_t i, j, k;

    STORM_contiguous_t* c = STORM_contig_new(M);
    STORM_t* s = STORM_new();
    if (!c || !s || !buf) { fprintf(stderr, "allocation failed\n"); return 3; }
    state = seed;
    for (i = 0; i < N; ++i) {
        const uint32_t n = gen_row(buf, M, draws, &state);
        const int rc = STORM_contig_add(c, buf, n);
        if (rc != (int)n) { fprintf(stderr, "STORM_contig_add returned %d for %u values\n", rc, n); return 4; }
        if (n && STORM_add(s, buf, n) != 1) { fprintf(stderr, "STORM_add failed\n"); return 4; }
        memset(buf, 0xFF, (size_t)draws * sizeof(uint32_t));          /* the library must have copied the values */
    }
    /* public fields, read the way external code may (storm.h:188-200) */
    {
        const uint32_t W = c->n_bitmaps_vector;
        for (i = 0; i < c->n_data; ++i)
            for (j = i + 1; j < c->n_data; ++j)
                for (k = 0; k < W; ++k)
                    naive += (uint64_t)__builtin_popcountll(c->data[(size_t)i * W + k] & c->data[(size_t)j * W + k]);
        printf("rows=%llu words=%u cutoff=%u naive=%llu", (unsigned long long)c->n_data, W, c->scalar_cutoff, (unsigned long long)naive);
        /* the kept function pointer is a valid host kernel (libalgebra.h:3035) */
        if (c->n_data >= 2) printf(" fptr01=%llu", (unsigned long long)c->intsec_func(c->data, c->data + W, W));
    }
    printf(" contig=%llu", (unsigned long long)STORM_contig_pairw_intersect_cardinality(c));
    printf(" contig_blocked=%llu", (unsigned long long)STORM_contig_pairw_intersect_cardinality_blocked(c, 31));
    printf(" contig_list=%llu", (unsigned long long)STORM_contig_pairw_intersect_cardinality_list(c));
    printf(" contig_blocked_list=%llu", (unsigned long long)STORM_contig_pairw_intersect_cardinality_blocked_list(c, 31));
    printf(" storm=%llu", (unsigned long long)STORM_pairw_intersect_cardinality(s));
    printf(" storm_blocked=%llu", (unsigned long long)STORM_pairw_intersect_cardinality_blocked(s, 0));
    printf(" wrapper=%llu", (unsigned long long)STORM_wrapper_diag(c->n_data, c->data, c->n_bitmaps_vector, c->intsec_func));
    printf(" wrapper_blocked=%llu", (unsigned long long)STORM_wrapper_diag_blocked(c->n_data, c->data, c->n_bitmaps_vector, c->intsec_func, 9));
    printf(" serialized=%llu", (unsigned long long)STORM_serialized_size(s));
    /* clear keeps the objects usable: second, sparser round on the same containers (benchmark.cpp:738-739) */
    STORM_contig_clear(c);
    STORM_clear(s);
    state = seed + 1;
    for (i = 0; i < N / 2; ++i) {
        const uint32_t n = gen_row(buf, M, draws / 4 + 1, &state);
        STORM_contig_add(c, buf, n);
        STORM_add(s, buf, n);
    }
    printf(" round2_rows=%llu round2_contig=%llu round2_storm=%llu", (unsigned long long)c->n_data,
           (unsigned long long)STORM_contig_pairw_intersect_cardinality_blocked(c, 7),
           (unsigned long long)STORM_pairw_intersect_cardinality(s));
    printf(" null_query=%llu null_add=%d\n", (unsigned long long)STORM_contig_pairw_intersect_cardinality(NULL),
           STORM_contig_add(NULL, buf, 1));
    STORM_contig_free(c);
    STORM_free(s);
    free(buf);
    return 0;
}
