/*
 * storm.h -- B200-native drop-in for the StormBitmaps C API.
 *
 * Source-compatible with the reference header (/root/reference/storm.h): same
 * entry-point names, argument meaning, return conventions, tunable macros and
 * public struct layout prefixes.  What differs is behind the boundary: the
 * pairwise-cardinality queries run as hand-written sm_100a CUDA kernels on a
 * device-resident mirror of the container; there is no CPU
 * implementation of a query in this library and every query fails loudly
 * (UINT64_MAX + STORM_b200_last_error()) when no CUDA device is usable.
 *
 * Layout contract (verified with offsetof in
 * tests/test_abi.py): every field the reference declares keeps its offset; new
 * state is reachable only through the trailing `b200` pointer, which is appended
 * AFTER the last reference field.  Objects are created only by the library's
 * own *_new functions, so growing them is ABI-safe for callers that hold
 * pointers.
 *
 * Each declaration cites the reference interface it replaces as
 * (storm.h:LINE -> storm.c:LINES).
 */
#ifndef STORM_B200_DROPIN_STORM_H_
#define STORM_B200_DROPIN_STORM_H_

#include <stddef.h>
#include <stdint.h>

/* ---- tunables that define container semantics (storm.h:37-47) ------------ */
#ifndef STORM_CACHE_BLOCK_SIZE
#define STORM_CACHE_BLOCK_SIZE 256e3            /* only feeds the bsize==0 heuristic */
#endif
#ifndef STORM_DEFAULT_BLOCK_SIZE
#define STORM_DEFAULT_BLOCK_SIZE 65536          /* bits per STORM_t block */
#endif
#ifndef STORM_DEFAULT_SCALAR_THRESHOLD
#define STORM_DEFAULT_SCALAR_THRESHOLD 4096     /* values below which a block is a u16 list */
#endif

/* ---- names the reference header gets from libalgebra.h ------------------- */
#ifndef STORM_ALIGN
#  if defined(__cplusplus)
#    define STORM_ALIGN(n) alignas(n)
#  elif defined(__STDC_VERSION__) && (__STDC_VERSION__ >= 201112L)
#    include <stdalign.h>
#    define STORM_ALIGN(n) alignas(n)
#  else
#    define STORM_ALIGN(n) __attribute__((aligned(n)))
#  endif
#endif
#ifndef STORM_RESTRICT
#  if defined(__cplusplus)
#    define STORM_RESTRICT __restrict__
#  else
#    define STORM_RESTRICT restrict
#  endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Per-pair kernel signature (libalgebra.h:3035).  Kept so that callers which
 * store or pass such pointers still compile; the GPU path ignores them. */
typedef uint64_t (*STORM_compute_func)(const uint64_t*, const uint64_t*, const size_t);
/* The kernel choosers of libalgebra.h:3094-3140 (intersect), 3142-3188 (union) and 3190-3236
 * (diff), which the reference's storm.h pulls in.  Here they return exported host functions of
 * this library (plain popcount loops with the same results).  A raw-buffer wrapper answers with the
 * set operation its `f` stands for: NULL means intersect, these three are known by address, and any
 * other function (e.g. libalgebra's static kernels compiled into the caller) is identified by what it
 * returns on three small probe vectors; a function that is none of the three gets UINT64_MAX. */
STORM_compute_func STORM_get_intersect_count_func(const size_t n_bitmaps_vector);
STORM_compute_func STORM_get_union_count_func(const size_t n_bitmaps_vector);
STORM_compute_func STORM_get_diff_count_func(const size_t n_bitmaps_vector);
/* Sparse-aware per-pair signature (storm.h:66-67). */
typedef uint64_t (*STORM_compute_lfunc)(const uint64_t*, const uint64_t*,
                                        const uint32_t*, const uint32_t*,
                                        const size_t, const size_t);

/* ========================================================================== *
 *  Public structs (storm.h:151-200).  Field order and types are load-bearing.
 * ========================================================================== */
typedef struct STORM_bitmap_s            STORM_bitmap_t;
typedef struct STORM_bitmap_cont_s       STORM_bitmap_cont_t;
typedef struct STORM_s                   STORM_t;
typedef struct STORM_contiguous_bitmap_s STORM_contiguous_bitmap_t;
typedef struct STORM_contiguous_s        STORM_contiguous_t;

/* One 65536-bit block of one row: a 1024-word bitmap or a sorted u16 list.
 * sizeof == 128, alignof == 64 (storm.h:158-166). */
struct STORM_bitmap_s {
    STORM_ALIGN(64) uint64_t* data;        /* @0   bitmap words (n_bitmap of them)        */
    STORM_ALIGN(64) uint16_t* scalar;      /* @64  block-relative values                  */
    uint32_t n_bitmap : 30, own_data : 1, own_scalar : 1;   /* @72 */
    uint32_t n_bits_set;                   /* @76 */
    uint32_t n_scalar : 31, n_scalar_set : 1, n_missing;    /* @80, @84 */
    uint32_t m_scalar;                     /* @88 */
    uint32_t id;                           /* @92  block index = value / 65536            */
};

/* One row of a STORM_t: its blocks, ascending by id (storm.h:168-173). */
struct STORM_bitmap_cont_s {
    STORM_bitmap_t* bitmaps;               /* @0  */
    uint32_t*       block_ids;             /* @8  ids duplicated for the merge loop       */
    uint32_t        n_bitmaps, m_bitmaps;  /* @16, @20 */
    uint32_t        prev_inserted_value;   /* @24 */
};

/* The sparse model: an array of rows (storm.h:175-178). */
struct STORM_s {
    STORM_bitmap_cont_t* conts;            /* @0  */
    uint32_t n_conts, m_conts;             /* @8, @12 */
    /* ---- appended by this implementation (not in the reference) ---- */
    void* b200;                            /* @16 device mirror + bookkeeping, opaque     */
};

/* Per-row view into the contiguous arena (storm.h:181-186). */
struct STORM_contiguous_bitmap_s {
    uint64_t* data;                        /* @0  host row (not owned)                    */
    uint32_t* scalar;                      /* @8  host position list (not owned)          */
    uint32_t  n_scalar;                    /* @16 unique set bits in the row              */
};

/* The dense model (storm.h:188-200). */
struct STORM_contiguous_s {
    uint64_t* data;                        /* @0  host mirror, n_data x n_bitmaps_vector  */
    uint32_t* scalar;                      /* @8  concatenated positions of sparse rows   */
    uint32_t* n_scalar;                    /* @16 per-row unique set-bit count            */
    STORM_contiguous_bitmap_t* bitmaps;    /* @24 */
    uint64_t  n_data, m_data;              /* @32, @40 rows used / allocated              */
    uint64_t  tot_scalar, m_scalar;        /* @48, @56 */
    uint64_t  vector_length;               /* @64 M, bits per row                         */
    uint32_t  n_bitmaps_vector;            /* @72 W = ceil(M/64)                          */
    STORM_compute_func intsec_func;        /* @80 kept non-NULL for callers that call it  */
    uint32_t  alignment;                   /* @88 host arena alignment in bytes           */
    uint32_t  scalar_cutoff;               /* @92 min(200, M/200)                         */
    /* ---- appended by this implementation (not in the reference) ---- */
    void* b200;                            /* @96 device arena + bookkeeping, opaque      */
};

/* ========================================================================== *
 *  Dense model: STORM_contiguous_t
 * ========================================================================== */

/* storm.h:233 -> storm.c:1001-1018.  NULL on allocation failure. */
STORM_contiguous_t* STORM_contig_new(size_t vector_length);

/* storm.h:234 -> storm.c:1020-1029.  Releases host and device arenas and (unlike
 * the reference, defect D8) the object itself. */
void STORM_contig_free(STORM_contiguous_t* bitmap);

/* storm.h:235 -> storm.c:1031-1137.  Appends one row from a sorted ascending
 * position list (adjacent duplicates are skipped).  Returns n_values; -1 NULL
 * object; -2 NULL values; 0 for an empty list, which appends NO row; -3 on
 * allocation failure or a position >= vector_length (the reference writes out
 * of bounds there). */
int STORM_contig_add(STORM_contiguous_t* bitmap, const uint32_t* values, const uint32_t n_values);

/* storm.h:236 -> storm.c:1139-1147.  Keeps capacity.  1 ok, 0 nothing allocated, -1 NULL. */
int STORM_contig_clear(STORM_contiguous_t* bitmap);

/* storm.h:237 -> storm.c:1149-1173.  sum_{i<j} popcount(row_i & row_j).
 * UINT64_MAX ((uint64_t)-1) for a NULL object or a CUDA failure. */
uint64_t STORM_contig_pairw_intersect_cardinality(STORM_contiguous_t* bitmap);

/* storm.h:238 -> storm.c:1175-1241.  Same value; `bsize` is a CPU cache-blocking
 * hint and does not change the result (integer adds commute). */
uint64_t STORM_contig_pairw_intersect_cardinality_blocked(STORM_contiguous_t* bitmap, uint32_t bsize);

/* storm.h:239 -> storm.c:1243-1263.  Same value; pairs with a sparse row
 * (< scalar_cutoff bits) are answered by the position-probe kernel. */
uint64_t STORM_contig_pairw_intersect_cardinality_list(STORM_contiguous_t* bitmap);

/* storm.h:240 -> storm.c:1265-1347. */
uint64_t STORM_contig_pairw_intersect_cardinality_blocked_list(STORM_contiguous_t* bitmap, uint32_t bsize);

/* ========================================================================== *
 *  Sparse model: STORM_t
 * ========================================================================== */

/* storm.h:223 -> storm.c:827-834. */
STORM_t* STORM_new(void);
/* storm.h:224 -> storm.c:836-842 (frees everything, D8). */
void STORM_free(STORM_t* bitmap);
/* storm.h:225 -> storm.c:844-866.  Appends one row (an empty list appends an
 * empty row).  Returns 1; -1 NULL object. */
int STORM_add(STORM_t* bitmap, const uint32_t* values, const uint32_t n_values);
/* storm.h:226 -> storm.c:868-875. */
int STORM_clear(STORM_t* bitmap);
/* storm.h:227 -> storm.c:877-895.  Exact sum over row pairs; the reference's
 * bitmap x list probe defect (D1, storm.c:636,644) is NOT reproduced. */
uint64_t STORM_pairw_intersect_cardinality(STORM_t* bitmap);
/* storm.h:228 -> storm.c:897-961.  bsize is a hint (0 = auto). */
uint64_t STORM_pairw_intersect_cardinality_blocked(STORM_t* bitmap, uint32_t bsize);
/* storm.h:229 -> declared in the reference but never defined (storm.c:975).
 * Here: sum over all (row of bitmap1, row of bitmap2) pairs. */
uint64_t STORM_intersect_cardinality_square(const STORM_t* STORM_RESTRICT bitmap1,
                                            const STORM_t* STORM_RESTRICT bitmap2);
/* storm.h:230 -> storm.c:963-973. */
uint64_t STORM_serialized_size(const STORM_t* bitmap);

/* Row containers (storm.h:214-222 -> storm.c:659-824): host-side builders. */
STORM_bitmap_cont_t* STORM_bitmap_cont_new(void);
void     STORM_bitmap_cont_init(STORM_bitmap_cont_t* bitmap);
void     STORM_bitmap_cont_free(STORM_bitmap_cont_t* bitmap);
int      STORM_bitmap_cont_add(STORM_bitmap_cont_t* bitmap, const uint32_t* values, const uint32_t n_values);
int      STORM_bitmap_cont_clear(STORM_bitmap_cont_t* bitmap);
uint32_t STORM_bitmap_cont_serialized_size(STORM_bitmap_cont_t* bitmap);

/* Block containers (storm.h:203-212 -> storm.c:398-569): host-side builders. */
STORM_bitmap_t* STORM_bitmap_new(void);
void     STORM_bitmap_init(STORM_bitmap_t* all);
void     STORM_bitmap_free(STORM_bitmap_t* bitmap);
int      STORM_bitmap_add(STORM_bitmap_t* bitmap, const uint32_t* values, const uint32_t n_values);
int      STORM_bitmap_add_scalar_only(STORM_bitmap_t* bitmap, const uint32_t* values, const uint32_t n_values);
int      STORM_bitmap_clear(STORM_bitmap_t* bitmap);
uint32_t STORM_bitmap_serialized_size(STORM_bitmap_t* bitmap);
/* storm.h:207 -> storm.c:467-519.  Sets the bits AND appends every value not seen before to the block's list. */
int      STORM_bitmap_add_with_scalar(STORM_bitmap_t* bitmap, const uint32_t* values, const uint32_t n_values);

/* ========================================================================== *
 *  Per-pair host helpers (storm.h:56-61, 209-210, 220-221 -> storm.c:4-129, 571-656, 761-814)
 *
 *  The reference's CPU building blocks for ONE pair of lists / blocks / rows.  They need no device and are
 *  plain host code here too (host_pairs.cu); the all-vs-all queries above never go through them.  All return
 *  the exact |a AND b|: the bitmap x list probe defect of storm.c:636,644 (D1) is not reproduced.
 * ========================================================================== */
/* storm.c:4-73.  Common values of two sorted unique u16 lists. */
uint64_t STORM_intersect_vector16_cardinality(const uint16_t* STORM_RESTRICT v1, const uint16_t* STORM_RESTRICT v2,
                                              const uint32_t len1, const uint32_t len2);
/* storm.c:75-106.  Merge of two sorted unique u32 lists: out receives (index in v1, index in v2) pairs of the common
 * values; returns the number of u32 written (2 per match).  `out` must hold 2 * min(len1, len2) entries. */
uint64_t STORM_intersect_vector32_unsafe(const uint32_t* STORM_RESTRICT v1, const uint32_t* STORM_RESTRICT v2,
                                         const uint32_t len1, const uint32_t len2, uint32_t* STORM_RESTRICT out);
/* storm.c:108-129.  The shorter position list probed into the other row's bitmap (n1 < n2: l1 into b2, else l2 into b1). */
uint64_t STORM_intersect_bitmaps_scalar_list(const uint64_t* STORM_RESTRICT b1, const uint64_t* STORM_RESTRICT b2,
                                             const uint32_t* l1, const uint32_t* l2, const uint32_t n1, const uint32_t n2);
/* storm.c:571-614 / 618-656.  Two blocks with the same id: list x list, bitmap x list, list x bitmap or
 * bitmap x bitmap (through `func`, or a popcount loop); 0 for different ids or NULL. */
uint64_t STORM_bitmap_intersect_cardinality(STORM_bitmap_t* STORM_RESTRICT bitmap1, STORM_bitmap_t* STORM_RESTRICT bitmap2);
uint64_t STORM_bitmap_intersect_cardinality_func(STORM_bitmap_t* STORM_RESTRICT bitmap1, STORM_bitmap_t* STORM_RESTRICT bitmap2,
                                                 const STORM_compute_func func);
/* storm.c:761-788 / 790-814.  Two rows: merge of the block ids, then the per-block function.  `out` is caller
 * scratch for the merge (2 * min(n_bitmaps) u32). */
uint64_t STORM_bitmap_cont_intersect_cardinality(const STORM_bitmap_cont_t* STORM_RESTRICT bitmap1,
                                                 const STORM_bitmap_cont_t* STORM_RESTRICT bitmap2);
uint64_t STORM_bitmap_cont_intersect_cardinality_premade(const STORM_bitmap_cont_t* STORM_RESTRICT bitmap1,
                                                         const STORM_bitmap_cont_t* STORM_RESTRICT bitmap2,
                                                         const STORM_compute_func func, uint32_t* out);

/* ========================================================================== *
 *  Raw-buffer wrappers (storm.h:95-148 -> storm.c:132-369)
 *
 *  `vals` is a caller-owned HOST buffer of n_vectors x n_ints words; the call
 *  uploads it (over every device of the device set, storm_b200.h), runs the tile
 *  kernels and returns the total.  `f` selects the set operation (intersect, or sum
 *  of popcount(a|b) resp. popcount(a^b) over the same pairs; see STORM_compute_func
 *  above); fl and block_size only steer CPU code in the reference and are ignored.
 * ========================================================================== */
uint64_t STORM_wrapper_diag(const uint32_t n_vectors, const uint64_t* vals,
                            const uint32_t n_ints, const STORM_compute_func f);
uint64_t STORM_wrapper_diag_blocked(const uint32_t n_vectors, const uint64_t* vals,
                                    const uint32_t n_ints, const STORM_compute_func f,
                                    uint32_t block_size);
/* XY^T total over two buffers -- the documented intent (storm.h:72-76); the
 * reference body never resets its inner offset (defect D4). */
uint64_t STORM_wrapper_square(const uint32_t n_vectors1, const uint64_t* STORM_RESTRICT vals1,
                              const uint32_t n_vectors2, const uint64_t* STORM_RESTRICT vals2,
                              const uint32_t n_ints, const STORM_compute_func f);
uint64_t STORM_wrapper_diag_list(const uint32_t n_vectors, const uint64_t* STORM_RESTRICT vals,
                                 const uint32_t n_ints, const uint32_t* STORM_RESTRICT n_alts,
                                 const uint32_t* STORM_RESTRICT alt_positions,
                                 const uint32_t* STORM_RESTRICT alt_offsets,
                                 const STORM_compute_func f, const STORM_compute_lfunc fl,
                                 const uint32_t cutoff);
uint64_t STORM_wrapper_diag_list_blocked(const uint32_t n_vectors, const uint64_t* STORM_RESTRICT vals,
                                         const uint32_t n_ints, const uint32_t* STORM_RESTRICT n_alts,
                                         const uint32_t* STORM_RESTRICT alt_positions,
                                         const uint32_t* STORM_RESTRICT alt_offsets,
                                         const STORM_compute_func f, const STORM_compute_lfunc fl,
                                         const uint32_t cutoff, uint32_t block_size);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* STORM_B200_DROPIN_STORM_H_ */
