/*
 * storm_oracle.h -- CPU restatement of the StormBitmaps all-vs-all
 * intersection-cardinality path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker.  The shipped
 * library (stormbitmaps_b200/csrc -> libstorm_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  The reference has no golden vectors of its own
 * (SURVEY.md section 4), so the restatement is pinned against the unmodified
 * reference compiled here into oracle/_ref/libstorm_ref.so (oracle/Makefile)
 * on seeded inputs; the agreed values are committed under tests/golden/ by
 * tools/make_golden.py and re-checked on CPU in tests/test_oracle.py.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  The code is written in plain scalar C on
 * purpose: no SIMD, no blocking, 64-bit offsets everywhere.
 */
#ifndef STORM_ORACLE_H_
#define STORM_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants that define container semantics (storm.h:37-47) ---------- */
#define ORC_BLOCK_BITS        65536u   /* STORM_DEFAULT_BLOCK_SIZE        */
#define ORC_BLOCK_WORDS       1024u    /* ceil(65536 / 64)                */
#define ORC_LIST_THRESHOLD    4096u    /* STORM_DEFAULT_SCALAR_THRESHOLD  */
#define ORC_CACHE_BLOCK_BYTES 256e3    /* STORM_CACHE_BLOCK_SIZE          */

/* ---- per-pair kernel ---------------------------------------------------- */
/* sum_k popcount(a[k] & b[k])  -- libalgebra.h:499-519,2985-2991 (scalar form
 * of the CSA kernels at :2684-2744,2872-2890, which compute the same value). */
uint64_t orc_intersect_count(const uint64_t* a, const uint64_t* b, size_t n_words);

/* sum_k popcount(a[k] | b[k])  -- libalgebra.h:2994-3000 (scalar), 521-540 (unrolled).
 * sum_k popcount(a[k] ^ b[k])  -- libalgebra.h:3002-3008 (scalar), 543-563 (unrolled). */
uint64_t orc_union_count(const uint64_t* a, const uint64_t* b, size_t n_words);
uint64_t orc_diff_count(const uint64_t* a, const uint64_t* b, size_t n_words);
/* storm.c:132-150 with f = the intersect (op 0), union (1) or diff (2) kernel. */
uint64_t orc_wrapper_diag_op(uint64_t n_vectors, const uint64_t* vals, uint64_t n_words, int op);
/* orc_rect_counts / orc_wrapper_square under the same choice of f. */
void orc_rect_counts_op(const uint64_t* vals, uint64_t n_words,
                        uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, int op, uint32_t* out);
uint64_t orc_wrapper_square_op(uint64_t n1, const uint64_t* vals1, uint64_t n2,
                               const uint64_t* vals2, uint64_t n_words, int op);

/* ---- raw-buffer loops --------------------------------------------------- */
/* storm.c:132-150 (STORM_wrapper_diag) with 64-bit offsets (defect D5). */
uint64_t orc_wrapper_diag(uint64_t n_vectors, const uint64_t* vals, uint64_t n_words);
/* Same total restricted to rows [i0,i1) x [j0,j1), only pairs with i<j. */
uint64_t orc_rect_total(const uint64_t* vals, uint64_t n_words,
                        uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1);
/* Per-pair counts of the rectangle, row-major (i1-i0) x (j1-j0); entries with
 * j<=i are written as 0 (the strict upper triangle is what the path sums). */
void orc_rect_counts(const uint64_t* vals, uint64_t n_words,
                     uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, uint32_t* out);
/* XY^T total over two buffers: what STORM_wrapper_square (storm.c:153-171) is
 * documented to compute (storm.h:72-76); the reference body is broken (D4). */
uint64_t orc_wrapper_square(uint64_t n1, const uint64_t* vals1, uint64_t n2,
                            const uint64_t* vals2, uint64_t n_words);

/* ---- independent closed form: sum_k C(c_k,2) ------ */
uint64_t orc_colcount_total(uint64_t n_vectors, const uint64_t* vals, uint64_t n_words);
/* sum_k cI_k * cJ_k for disjoint row ranges I=[i0,i1), J=[j0,j1). */
uint64_t orc_colcount_rect(const uint64_t* vals, uint64_t n_words,
                           uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1);

/* ---- contiguous model (storm.h:181-200, storm.c:1001-1347) -------------- */
typedef struct orc_contig_s {
    uint64_t* data;        /* n_rows x n_words, row-major, stride == n_words      */
    uint32_t* positions;   /* concatenated unique positions of sparse rows        */
    uint64_t* pos_offset;  /* per row: offset into positions (valid if sparse)    */
    uint32_t* n_set;       /* per row: unique set-bit count (reference n_scalar)  */
    uint64_t  n_rows, cap_rows;
    uint64_t  n_pos, cap_pos;
    uint64_t  vector_length;
    uint32_t  n_words;
    uint32_t  scalar_cutoff;
} orc_contig_t;

orc_contig_t* orc_contig_new(size_t vector_length);          /* storm.c:1001-1018 */
void          orc_contig_free(orc_contig_t* c);               /* storm.c:1020-1029 */
int           orc_contig_add(orc_contig_t* c, const uint32_t* values, uint32_t n_values); /* :1031-1137 */
int           orc_contig_clear(orc_contig_t* c);              /* storm.c:1139-1147 */
uint64_t      orc_contig_pairw(const orc_contig_t* c);        /* storm.c:1149-1173 */
uint64_t      orc_contig_pairw_blocked(const orc_contig_t* c, uint32_t bsize); /* :1175-1241 */
uint64_t      orc_contig_pairw_list(const orc_contig_t* c);   /* storm.c:1243-1263 */
uint64_t      orc_contig_pairw_blocked_list(const orc_contig_t* c, uint32_t bsize); /* :1265-1347 */
/* storm.c:108-129: probe the shorter position list into the other bitmap. */
uint64_t      orc_probe_list(const uint64_t* b1, const uint64_t* b2,
                             const uint32_t* l1, const uint32_t* l2, uint32_t n1, uint32_t n2);

/* ---- sparse STORM_t model (storm.h:158-178, storm.c:398-973) ------------ */
typedef struct orc_block_s {
    uint32_t  id;          /* block index = value / 65536                     */
    uint32_t  is_bitmap;   /* reference: n_bitmap != 0                        */
    uint32_t  n_values;    /* values handed to the builder for this block     */
    uint64_t* words;       /* 1024 words when is_bitmap                       */
    uint16_t* list;        /* block-relative values when !is_bitmap           */
} orc_block_t;

typedef struct orc_row_s {
    orc_block_t* blocks;
    uint32_t     n_blocks, cap_blocks;
} orc_row_t;

typedef struct orc_storm_s {
    orc_row_t* rows;
    uint32_t   n_rows, cap_rows;
} orc_storm_t;

orc_storm_t* orc_storm_new(void);                                   /* storm.c:827-834  */
void         orc_storm_free(orc_storm_t* s);                        /* storm.c:836-842  */
int          orc_storm_add(orc_storm_t* s, const uint32_t* values, uint32_t n_values); /* :844-866, :692-758 */
int          orc_storm_clear(orc_storm_t* s);                       /* storm.c:868-875  */
/* storm.c:877-895.  emulate_d1 != 0 reproduces defect D1 (storm.c:636,644: the
 * bitmap x list probe adds `word & 1` instead of the probed bit) so that the
 * restatement can be pinned against the compiled reference on mixed inputs;
 * emulate_d1 == 0 is the mathematically exact value the GPU must match. */
uint64_t     orc_storm_pairw(const orc_storm_t* s, int emulate_d1);
uint64_t     orc_storm_pairw_blocked(const orc_storm_t* s, uint32_t bsize, int emulate_d1); /* :897-961 */
uint64_t     orc_storm_row_pair(const orc_row_t* a, const orc_row_t* b, int emulate_d1);    /* :790-814, :618-656 */
uint64_t     orc_storm_serialized_size(const orc_storm_t* s);       /* storm.c:372-394,963-973 */
uint32_t     orc_storm_auto_bsize(const orc_storm_t* s);            /* storm.c:903-914 */
/* storm.c:4-73: |v1 ∩ v2| for sorted u16 lists (scalar merge; the SSE4.2 body
 * computes the same cardinality for unique sorted inputs). */
uint64_t     orc_intersect_u16(const uint16_t* v1, const uint16_t* v2, uint32_t n1, uint32_t n2);
/* storm.c:75-106: merge two sorted id lists, emit index pairs; returns 2*matches. */
uint64_t     orc_intersect_u32_pairs(const uint32_t* v1, const uint32_t* v2,
                                     uint32_t n1, uint32_t n2, uint32_t* out);

/* ---- deterministic synthetic inputs (benchmark.cpp:749-797 recipe) ------ */
/* Counter-based generator shared bit-for-bit with the CUDA generator
 * (stormbitmaps_b200/csrc/synth.cuh).  Row `row` gets `n_draws` positions drawn
 * uniformly WITH replacement from [0, M); returns the number of unique sorted
 * positions written to out (capacity n_draws). */
uint64_t orc_splitmix64(uint64_t x);
uint32_t orc_draw_position(uint64_t seed, uint64_t row, uint64_t draw, uint32_t M);
uint32_t orc_gen_row_positions(uint64_t seed, uint64_t row, uint32_t n_draws, uint32_t M, uint32_t* out);
/* Fill a dense n_rows x n_words buffer (must be zeroed by the caller). */
void     orc_gen_dense_uniform(uint64_t seed, uint64_t row0, uint64_t n_rows, uint32_t n_draws,
                               uint32_t M, uint64_t* vals, uint64_t n_words);
/* Genotype-like rows, integer-only so that the
 * CPU and the CUDA generator agree bit for bit: per-row minor-allele frequency
 * p = max(0.005, 0.5 * u^2), u uniform in [0,1) (rare-variant-skewed spectrum);
 * as a 32-bit threshold thr = max(21474836, (u24*u24) >> 17).  Bit k of row r
 * is set iff fmix32(k * 0x9E3779B1 + rowkey(seed,r)) < thr. */
uint32_t orc_geno_threshold(uint64_t seed, uint64_t row);
void     orc_gen_dense_geno(uint64_t seed, uint64_t row0, uint64_t n_rows, uint32_t M,
                            uint64_t* vals, uint64_t n_words);

#ifdef __cplusplus
}
#endif
#endif /* STORM_ORACLE_H_ */
