/*
 * contig_api_bench.c -- the north-star struct API end to end, as benchmark.cpp drives it
 * (benchmark.cpp:711 STORM_contig_new, :795 STORM_contig_add per row, :910 the blocked query), in plain C99
 * against <storm.h> only (+ storm_b200.h for the bulk-ingest variant).  Prints one JSON object:
 *
 *   ingest_s        seconds inside N x STORM_contig_add (row generation excluded)
 *   first_query_s   first STORM_contig_pairw_intersect_cardinality_blocked after the ingest (uploads whatever the
 *                   background copies have not pushed yet)
 *   steady_query_s  best of `reps` further queries (rows resident)
 *   bulk            the same with STORM_b200_contig_add_bulk in chunks of 4096 rows
 *   closed_form     sum_k C(c_k, 2) over the column counts of the generated rows (independent checksum)
 *
 * Rows: per row one of five densities (about 3, 12.5, 25, 50, 75 % ones: AND / OR of xorshift words), positions
 * extracted in ascending order -- sorted and duplicate-free, as the driver hands them over (benchmark.cpp:765-767).
 *
 *   cc -std=c99 -O2 -I include tools/contig_api_bench.c -L stormbitmaps_b200 -lstorm_b200 -Wl,-rpath,... -o contig_api_bench
 *   contig_api_bench <bits M> <rows N> [reps=3] [bulk=1]
 */
#define _POSIX_C_SOURCE 199309L
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "storm_b200.h"

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) {
    uint64_t x = rng_state;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return rng_state = x;
}

/* one row of W words -> ascending positions; returns their number */
static uint32_t gen_row(uint32_t M, uint32_t W, uint32_t cls, uint32_t* pos, uint64_t* col) {
    uint32_t n = 0;
    for (uint32_t k = 0; k < W; ++k) {
        uint64_t w;
        switch (cls) {
            case 0: w = rnd() & rnd() & rnd() & rnd() & rnd(); break;
            case 1: w = rnd() & rnd() & rnd(); break;
            case 2: w = rnd() & rnd(); break;
            case 3: w = rnd(); break;
            default: w = rnd() | rnd(); break;
        }
        if ((uint64_t)k * 64 + 64 > M) w &= (~0ull) >> (64 - (M - k * 64));
        while (w) {
            const uint32_t p = k * 64 + (uint32_t)__builtin_ctzll(w);
            pos[n++] = p;
            if (col) ++col[p];
            w &= w - 1;
        }
    }
    if (n == 0) { pos[n++] = 0; if (col) ++col[0]; }     /* an empty list would append no row (D7) */
    return n;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <bits> <rows> [reps] [bulk]\n", argv[0]); return 2; }
    const uint32_t M = (uint32_t)strtoul(argv[1], NULL, 10);
    const uint64_t N = strtoull(argv[2], NULL, 10);
    const int reps = argc > 3 ? atoi(argv[3]) : 3;
    const int do_bulk = argc > 4 ? atoi(argv[4]) : 1;
    const uint32_t W = (M + 63) / 64;
    uint32_t* pos = (uint32_t*)malloc((size_t)W * 64 * sizeof(uint32_t));
    uint64_t* col = (uint64_t*)calloc((size_t)W * 64, sizeof(uint64_t));
    if (!pos || !col) return 3;

    /* ---- row by row ------------------------------------------------------------------------------------ */
    STORM_contiguous_t* c = STORM_contig_new(M);
    if (!c) return 3;
    double ingest_s = 0, gen_s = 0;
    uint64_t n_pos = 0;
    for (uint64_t r = 0; r < N; ++r) {
        double t0 = now_s();
        const uint32_t n = gen_row(M, W, (uint32_t)(r % 5), pos, col);
        double t1 = now_s();
        if (STORM_contig_add(c, pos, n) != (int)n) { fprintf(stderr, "STORM_contig_add failed: %s\n", STORM_b200_last_error()); return 4; }
        double t2 = now_s();
        gen_s += t1 - t0; ingest_s += t2 - t1; n_pos += n;
    }
    uint64_t closed = 0;
    for (uint64_t k = 0; k < (uint64_t)W * 64; ++k) closed += col[k] * (col[k] - (col[k] ? 1 : 0)) / 2;
    const uint32_t bsize = (uint32_t)(256e3 / (W * 8.0)) > 5 ? (uint32_t)(256e3 / (W * 8.0)) : 5;   /* benchmark.cpp:823-824 */
    double t0 = now_s();
    const uint64_t first = STORM_contig_pairw_intersect_cardinality_blocked(c, bsize);
    const double first_s = now_s() - t0;
    double steady_s = 1e30;
    uint64_t steady = first;
    for (int i = 0; i < reps; ++i) {
        t0 = now_s();
        steady = STORM_contig_pairw_intersect_cardinality_blocked(c, bsize);
        const double dt = now_s() - t0;
        if (dt < steady_s) steady_s = dt;
    }
    const int devices = STORM_b200_contig_device_count(c);
    STORM_contig_free(c);

    /* ---- bulk ingest in chunks ---------------------------------------------------------------------------- */
    double bulk_ingest_s = 0, bulk_first_s = 0;
    uint64_t bulk_total = 0;
    if (do_bulk) {
        rng_state = 0x9E3779B97F4A7C15ull;               /* the same rows again */
        const uint64_t CH = 4096;
        uint32_t* big = (uint32_t*)malloc((size_t)CH * W * 64 * sizeof(uint32_t));
        uint64_t* offs = (uint64_t*)malloc((CH + 1) * sizeof(uint64_t));
        if (!big || !offs) return 3;
        c = STORM_contig_new(M);
        for (uint64_t r0 = 0; r0 < N; r0 += CH) {
            const uint64_t n_rows = N - r0 < CH ? N - r0 : CH;
            offs[0] = 0;
            for (uint64_t r = 0; r < n_rows; ++r) offs[r + 1] = offs[r] + gen_row(M, W, (uint32_t)((r0 + r) % 5), big + offs[r], NULL);
            t0 = now_s();
            if (STORM_b200_contig_add_bulk(c, big, offs, n_rows) != 0) { fprintf(stderr, "add_bulk failed: %s\n", STORM_b200_last_error()); return 4; }
            bulk_ingest_s += now_s() - t0;
        }
        t0 = now_s();
        bulk_total = STORM_contig_pairw_intersect_cardinality_blocked(c, bsize);
        bulk_first_s = now_s() - t0;
        STORM_contig_free(c);
        free(big); free(offs);
    }
    const double wp = (double)N * (double)(N - 1) / 2.0 * (double)W;
    printf("{\"bits\": %u, \"rows\": %llu, \"positions\": %llu, \"devices\": %d, \"gen_s\": %.4f, \"ingest_s\": %.4f, \"ingest_call\": \"STORM_contig_add x rows\", "
           "\"first_query_s\": %.6f, \"steady_query_s\": %.6f, \"first_over_steady\": %.4f, \"steady_wp_per_s\": %.6e, "
           "\"total\": %llu, \"steady_total\": %llu, \"closed_form\": %llu, \"match\": %s, "
           "\"bulk\": {\"ingest_s\": %.4f, \"ingest_call\": \"STORM_b200_contig_add_bulk x chunks of 4096 rows\", \"first_query_s\": %.6f, \"total\": %llu, \"match\": %s}}\n",
           M, (unsigned long long)N, (unsigned long long)n_pos, devices, gen_s, ingest_s, first_s, steady_s, first_s / steady_s, wp / steady_s,
           (unsigned long long)first, (unsigned long long)steady, (unsigned long long)closed,
           (first == closed && steady == closed) ? "true" : "false",
           bulk_ingest_s, bulk_first_s, (unsigned long long)bulk_total, (!do_bulk || bulk_total == closed) ? "true" : "false");
    free(pos); free(col);
    return (first == closed && steady == closed && (!do_bulk || bulk_total == closed)) ? 0 : 1;
}
