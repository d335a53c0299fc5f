/*
 * storm_b200.h -- additive C-ABI of libstorm_b200.so (nothing here exists in
 * the reference; everything the reference declares is in storm.h).
 *
 * Plain pointers and sizes only: no torch / CUDA types in any signature.  A
 * `stream` argument is a CUDA stream handle passed as void* (0 / NULL = the
 * legacy default stream); a `d_` pointer is a device pointer on the current
 * CUDA device of the calling thread.
 *
 * Conventions: int functions return 0 on success and a negative STORM_B200_E*
 * code on failure; STORM_b200_last_error() returns a thread-local message for
 * the most recent failure on the calling thread.
 */
#ifndef STORM_B200_H_
#define STORM_B200_H_

#include "storm.h"

#ifdef __cplusplus
extern "C" {
#endif

#define STORM_B200_OK          0
#define STORM_B200_EINVAL     -1   /* bad argument                                  */
#define STORM_B200_ECUDA      -2   /* CUDA runtime / driver error (see last_error)  */
#define STORM_B200_ENODEV     -3   /* no usable sm_100 device                       */
#define STORM_B200_ENOMEM     -4   /* host or device allocation failed              */

/* Which dense tile kernel answers a query.  AUTO picks
 * UMMA when the shape fills its tiles and POPC otherwise. */
#define STORM_B200_KERNEL_AUTO 0
#define STORM_B200_KERNEL_POPC 1   /* CUDA cores: LOP3 + POPC register tiles          */
#define STORM_B200_KERNEL_UMMA 2   /* tcgen05.mma kind::i8 on bits unpacked on the fly */
#define STORM_B200_KERNEL_CSA  3   /* CUDA cores: carry-save adders feeding POPC        */
#define STORM_B200_KERNEL_B1   5   /* mma.sync.m16n8k256 .b1 AND + POPC: the pre-Blackwell one-bit tensor path, which sm_100a
                                      emulates with IMMA + ALU glue; kept for the record, never chosen by AUTO        */
#define STORM_B200_KERNEL_FP4  4   /* tcgen05.mma kind::mxf4 on bits unpacked to E2M1 nibbles, fp32 accumulators
                                      (exact below 2^24 bits per row; twice the rate of kind::i8) */

/* Set operation of a pairwise query (libalgebra's three per-pair kernel families,
 * libalgebra.h:2985-3008): |a & b|, |a | b| = |a| + |b| - |a & b|, |a ^ b| = |a| + |b| - 2 |a & b|. */
#define STORM_B200_OP_INTERSECT 0
#define STORM_B200_OP_UNION     1
#define STORM_B200_OP_DIFF      2

/* ---- library / device ---------------------------------------------------- */
const char* STORM_b200_last_error(void);
const char* STORM_b200_version(void);
int  STORM_b200_device_count(void);
/* Name, SM count, compute capability major*10+minor of device `dev`. */
int  STORM_b200_device_info(int dev, char* name, size_t name_len, int* sm_count, int* cc);
/* Process-wide default kernel for the storm.h entry points (they have no
 * parameter for it).  Returns the previous value. */
int  STORM_b200_set_default_kernel(int kernel);

/* ---- devices ---------------------------------------------------------------
 *
 * Queries behind storm.h run on a SET of devices (one process, one calling thread): every device holds all
 * rows, device g of G computes shard g of the tile raster, the host adds G totals.  Rows reach the devices in
 * bands -- 1/G of a band over each device's own PCIe link, the other slices from the peers over NVLink -- and the
 * tiles of a band start as soon as it is complete on a device.  The set is read when a container first touches a
 * device (and per call for the raw-buffer wrappers):
 *   environment STORM_B200_DEVICES = all | <k> | <id>,<id>,...   (read once, before the first query)
 *   STORM_b200_set_devices(n): n = 0 all visible devices, n >= 1 devices 0 .. n-1; returns the previous count
 *   STORM_b200_set_device_list(ids, n): explicit ordinals (one may repeat: several replicas on one device, which
 *     is how the multi-device logic is tested on a one-GPU box); n = 0 restores the default
 * Default: the calling thread's current device alone (what a multi-process caller with one rank per GPU wants).
 * Whole-container STORM_t queries run on the set as well (the block mirror is uploaded to every device, each densifies or
 * probes its own copy and answers its shard); per-pair rectangles and XY^T totals use the first device of the set. */
int STORM_b200_set_devices(int n);
int STORM_b200_set_device_list(const int* ids, int n);
/* The devices a query made now would use: fills ids[0 .. min(cap, count)), returns the count (negative on error). */
int STORM_b200_get_devices(int* ids, int cap);
/* A query on G devices makes ~12 us of driver calls per device (launch, read-back, waits).  By default they go out from
 * G host threads at once -- the caller takes the first device, G - 1 worker threads created at the first multi-device
 * query take the others -- which is what lets 0.3 ms queries scale over 8 devices.  0 = issue everything from the
 * calling thread.  Returns the previous value. */
int STORM_b200_set_device_threads(int on);
/* Test hook for that pool (no device needed): n jobs, job `fail_at` (if >= 0) fails on purpose; returns the pool's return
 * code (the failing job's, with its message as the caller's last error) or -100 if a job did not run exactly once. */
int STORM_b200_selftest_device_threads(int n, int fail_at);

/* ---- dense path on device-resident rows ----------------------------------
 *
 * Rows are row-major 64-bit words, row r at d_rows + r*row_stride_words; bit v
 * of a row is (word[v/64] >> (v%64)) & 1 (storm.c:1114).  d_rows must be 16-byte
 * aligned and row_stride_words even.  Words [n_words, row_stride_words) of each
 * row are never read.
 *
 * The strict upper triangle is cut into row-block tiles; tiles are numbered in a
 * fixed raster and shard s of n_shards owns a contiguous, equally sized range of
 * tile indices (multi-GPU: one shard per rank, no data-path collective; the
 * caller adds the partial totals).  *d_total (device, 8 bytes) is ACCUMULATED
 * into with one 64-bit atomic add per CTA: zero it first.
 *
 * Replaces the loop nest of storm.c:1149-1173 / 1175-1241 (and 132-150, 222-279). */
int STORM_b200_pairw_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                            uint64_t row_stride_words, uint32_t shard, uint32_t n_shards,
                            int kernel, uint64_t* d_total, void* stream);

/* The same total over a matrix that is ALREADY resident on several devices (d_rows[g] on device device_ids[g]; same
 * n_rows, n_words and row stride everywhere): device g computes shard g of n_devices, the host adds the totals.  One
 * host thread, one launch per device, no collective.  UINT64_MAX on error. */
uint64_t STORM_b200_pairw_devices(const uint64_t* const* d_rows, const int* device_ids, int n_devices, uint64_t n_rows,
                                  uint32_t n_words, uint64_t row_stride_words, int kernel);

/* Per-pair counts of the rectangle rows [i0,i1) x [j0,j1): d_out[(i-i0)*ld + (j-j0)]
 * = popcount(row_i & row_j); with strict_upper != 0 entries with j <= i are
 * written as 0 and excluded from the total.  d_total may be NULL.  d_out may be
 * NULL (total only).  The per-pair expression is storm.c:1167. */
int STORM_b200_pairw_rect_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                 uint64_t row_stride_words,
                                 uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1,
                                 int strict_upper, int kernel,
                                 uint32_t* d_out, uint64_t ld, uint64_t* d_total, void* stream);

/* XY^T over two device matrices with the same n_words (all n1 x n2 pairs). */
int STORM_b200_square_device(const uint64_t* d_rows1, uint64_t n1, uint64_t stride1,
                             const uint64_t* d_rows2, uint64_t n2, uint64_t stride2,
                             uint32_t n_words, int kernel,
                             uint32_t* d_out, uint64_t ld, uint64_t* d_total, void* stream);

/* Set bits per row: d_counts[r] = popcount(row r) (uint32, device). */
int STORM_b200_row_popcounts_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                    uint64_t row_stride_words, uint32_t* d_counts, void* stream);
/* Upper-triangle total under a set operation `op` (STORM_B200_OP_*), ACCUMULATED into *d_total:
 * row popcounts (one HBM pass) + the intersection tiles.  What STORM_wrapper_diag computes in the
 * reference when handed STORM_get_union_count_func / STORM_get_diff_count_func (storm.c:132-150). */
int STORM_b200_pairw_op_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                               uint64_t row_stride_words, int op, int kernel, uint64_t* d_total, void* stream);
/* STORM_b200_pairw_rect_device under a set operation (at most 65535 rows i1 - i0 when d_out is given). */
int STORM_b200_pairw_rect_op_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                    uint64_t row_stride_words,
                                    uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1,
                                    int strict_upper, int op, int kernel,
                                    uint32_t* d_out, uint64_t ld, uint64_t* d_total, void* stream);
/* The host functions behind STORM_get_{intersect,union,diff}_count_func (storm.h). */
uint64_t STORM_b200_host_intersect_count(const uint64_t* a, const uint64_t* b, const size_t n_words);
uint64_t STORM_b200_host_union_count(const uint64_t* a, const uint64_t* b, const size_t n_words);
uint64_t STORM_b200_host_diff_count(const uint64_t* a, const uint64_t* b, const size_t n_words);

/* Number of tiles the triangle raster of `kernel` has for n_rows (so callers can
 * reason about shard balance), and the tile edge lengths it uses. */
uint64_t STORM_b200_tile_count(uint64_t n_rows, int kernel, uint32_t* tile_rows, uint32_t* tile_cols);

/* Host-only bookkeeping of the same raster (no device needed): the tile range
 * [*tile_begin, *tile_end) that shard `shard` of `n_shards` owns, and the row
 * rectangle [i0,i1) x [j0,j1) (clipped to n_rows) a tile index covers; a tile
 * contributes its pairs with j > i.  Multi-GPU callers and the CPU tests use
 * these to check that the shards partition the triangle. */
int STORM_b200_shard_tiles(uint64_t n_rows, int kernel, uint32_t shard, uint32_t n_shards,
                           uint64_t* tile_begin, uint64_t* tile_end);
int STORM_b200_tile_rect(uint64_t n_rows, int kernel, uint64_t tile,
                         uint64_t* i0, uint64_t* i1, uint64_t* j0, uint64_t* j1);

/* The same query over an explicit range [tile_begin, tile_end) of the triangle raster (any
 * partition of [0, STORM_b200_tile_count) into ranges adds up to the full total).  With
 * STORM_b200_tiles_below_row this lets a caller start on the tiles whose rows are already
 * resident while later rows are still arriving (host upload, NVLink all-gather). */
int STORM_b200_pairw_tiles_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                  uint64_t row_stride_words, uint64_t tile_begin, uint64_t tile_end,
                                  int kernel, uint64_t* d_total, void* stream);
/* The same with the number of SMs the persistent tensor kernel leaves free passed per launch (reserved_sms >= 0; a
 * negative value means the process default of STORM_b200_set_umma_reserved_sms): a multi-process caller that overlaps
 * an NVLink collective with the tile kernel no longer has to flip a process-wide knob around its launches. */
int STORM_b200_pairw_tiles_device_ex(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                     uint64_t row_stride_words, uint64_t tile_begin, uint64_t tile_end,
                                     int kernel, int reserved_sms, uint64_t* d_total, void* stream);
/* Host-only: the raster is monotone in the largest row a tile reads.  *tile_end = number of
 * leading tiles that read only rows below row_limit (all tiles once row_limit >= n_rows);
 * *band_rows (optional) = the row granularity at which that count grows. */
int STORM_b200_tiles_below_row(uint64_t n_rows, int kernel, uint64_t row_limit,
                               uint64_t* tile_end, uint64_t* band_rows);

/* Which kernel id AUTO (or the process default) resolves to for rows of n_words. */
int STORM_b200_resolve_kernel(int kernel, uint32_t n_words);

/* Sharded form of STORM_wrapper_diag (storm.h:95-98): HOST buffer of n_vectors x
 * n_ints words (row pitch n_ints) -> upload -> this shard's partial total.
 * UINT64_MAX on error. */
uint64_t STORM_b200_wrapper_diag_shard(uint64_t n_vectors, const uint64_t* vals, uint64_t n_ints,
                                       uint32_t shard, uint32_t n_shards, int kernel);

/* ---- container extensions ------------------------------------------------ */
/* Device replicas the container's queries run on (0 before it first used a device). */
int STORM_b200_contig_device_count(const STORM_contiguous_t* bitmap);
/* Partial total of shard `shard` of `n_shards` (see above); UINT64_MAX on error. */
uint64_t STORM_b200_contig_pairw_shard(STORM_contiguous_t* bitmap, uint32_t shard, uint32_t n_shards, int kernel);
/* Per-pair counts into a HOST buffer out[(i-i0)*(j1-j0) + (j-j0)], strict upper triangle. */
int STORM_b200_contig_pairw_rect(STORM_contiguous_t* bitmap, uint64_t i0, uint64_t i1,
                                 uint64_t j0, uint64_t j1, uint32_t* out);
/* Device arena of a contiguous container after making it current (uploads dirty
 * rows).  Returns the device pointer and the row stride in words. */
const uint64_t* STORM_b200_contig_device_rows(STORM_contiguous_t* bitmap, uint64_t* row_stride_words);
/* Drop the device copy so that the next query uploads again (used to time the
 * host-buffer path end to end). */
int STORM_b200_contig_invalidate_device(STORM_contiguous_t* bitmap);
/* Bulk ingest: n_rows rows given as one concatenated sorted position array and
 * an offsets array of n_rows+1 entries.  Positions are uploaded once and the
 * bits are scattered on the device; same row semantics as STORM_contig_add
 * (rows with zero positions are skipped, D7). */
int STORM_b200_contig_add_bulk(STORM_contiguous_t* bitmap, const uint32_t* positions,
                               const uint64_t* offsets, uint64_t n_rows);
/* Rows handed over as bitmaps (n_rows x n_bitmaps_vector words at `pitch_words`; bit v of a row = word[v/64] >> (v%64) & 1,
 * storm.c:1114) instead of position lists: same row semantics as STORM_contig_add on the row's sorted positions (an
 * all-zero row appends nothing; rows below the cutoff also get their position list).  Bits >= vector_length must be 0. */
int STORM_b200_contig_add_dense(STORM_contiguous_t* bitmap, const uint64_t* rows, uint64_t n_rows, uint64_t pitch_words);
/* Move the container's device replicas to the device set in force now; they are rebuilt from the host mirror lazily. */
int STORM_b200_contig_rehome(STORM_contiguous_t* bitmap);
/* Seconds spent in the last query of this object: [0] upload (H2D), [1] kernels, [2] total. */
int STORM_b200_contig_last_timing(STORM_contiguous_t* bitmap, double out_seconds[3]);

/* How the *_list queries of the contiguous model are answered (storm.c:1253-1258 switches per pair between the
 * probe and the bitmap kernel; both give |i AND j|): 0 = cost model (default), 1 = every pair through the tile
 * kernel, 2 = dense x dense pairs through the tile kernel and every pair with a sparse row through the probe
 * kernel, 3 = the row-group stream kernel over the position lists when every row is sparse (else as 2).
 * Results are identical.  Returns the previous value; ..._last_list_route tells which one the last query took. */
int STORM_b200_set_contig_list_route(int route);
int STORM_b200_contig_last_list_route(const STORM_contiguous_t* bitmap);

uint64_t STORM_b200_storm_pairw_shard(STORM_t* bitmap, uint32_t shard, uint32_t n_shards);
/* How whole-container STORM_t queries are answered: 0 = cost model (default), 1 = the sparse
 * merge/probe kernels, 2 = rows densified on the device + the dense tile kernel (falls back to
 * 1 if the dense form does not fit in device memory), 3 = the split route where it applies (containers that hold
 * heavy rows -- a bitmap block, or more than 8 192 values -- among light ones: the light rows among themselves through
 * the row-group stream kernel, every pair with a heavy row through the block merge/probe kernel; else 1).  Results are
 * identical.  Returns the previous value; STORM_b200_storm_last_route tells which one the last query of `bitmap` took. */
int STORM_b200_set_storm_route(int route);
int STORM_b200_storm_last_route(const STORM_t* bitmap);   /* 1 sparse kernels, 2 densified rows + tile kernel, 3 the same in row bands, 4 split */
/* The split route's side of the cost model (pure arithmetic): seconds for n_rows rows of which n_heavy are heavy, the
 * light ones holding light_nnz values together, all rows total_nnz, at most max_blocks blocks per row, n_bitmap_blocks
 * bitmap blocks in the container.  Negative on a bad argument. */
double STORM_b200_storm_split_model(uint64_t n_rows, uint64_t n_heavy, double light_nnz, double total_nnz, double max_blocks,
                                    double n_bitmap_blocks);
/* A container whose dense form exceeds 48 GiB is densified in row bands of 12 GiB (two arenas; triangle of a band, then its
 * rectangles with the later bands) instead of falling back to the block merge/probe kernel.  This knob forces the banded
 * form with bands of `rows` rows for any size (0 = the size rule again); results are identical.  Returns the previous value. */
uint64_t STORM_b200_set_storm_band_rows(uint64_t rows);
/* The route cost model as a function (pure arithmetic, no device): expected seconds of a whole-container query of
 * n_rows rows of n_words 64-bit words holding avg_nnz values in avg_blocks blocks each, out[0] through the
 * densified rows + tile kernel, out[1] through the sparse kernels.  The query takes the smaller one. */
int STORM_b200_storm_route_model(uint64_t n_rows, uint32_t n_words, double avg_nnz, double avg_blocks, uint32_t max_row_nnz,
                                 uint64_t n_bitmap_blocks, int fp4, int dense_resident, double out_seconds[2]);
/* Kernels of the sparse route for containers without bitmap blocks (their rows are mirrored as flat position
 * lists on the device).  2 (default): totals through the row-group stream kernel (32 rows i per CTA in a
 * shared-memory hash position -> row mask, all later rows' positions streamed through it), per-pair rectangles
 * through the flat probe kernel (positions of a partner row probed into the whole-row shared bitmap of row i,
 * rows up to 1 310 720 bits); 1: the flat probe kernel for both; 0: always the block merge/probe kernel.
 * Results are identical.  Returns the previous value. */
int STORM_b200_set_sparse_flat(int mode);
/* Per-pair counts of a STORM_t rectangle into a HOST buffer (strict upper). */
int STORM_b200_storm_pairw_rect(STORM_t* bitmap, uint64_t i0, uint64_t i1,
                                uint64_t j0, uint64_t j1, uint32_t* out);

/* ---- synthetic inputs on the device (bit-identical to oracle/storm_oracle.c) */
/* benchmark.cpp:749-797 recipe: n_draws uniform positions with replacement per row. */
int STORM_b200_synth_uniform_device(uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                    uint64_t row_stride_words, uint32_t M, uint32_t n_draws,
                                    uint64_t seed, uint64_t row0, void* stream);
/* Genotype-like rows. */
int STORM_b200_synth_geno_device(uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                 uint64_t row_stride_words, uint32_t M,
                                 uint64_t seed, uint64_t row0, void* stream);

/* ---- measurement helpers -------------------------------------------------- */
/* Issue-rate micro-benchmarks used as roofline denominators for the CUDA-core
 * kernels (MEASURED_PEAKS.json only has HBM and bf16).  kind: 0 POPC.32,
 * 1 LOP3.32, 2 IADD3, 3 POPC+LOP3 mixed: thread-instructions per second over the
 * whole device in *rate (and a clock64-derived SM clock in *sm_mhz, indicative
 * only).  kind 4 / 5: the UMMA kernel's own tcgen05.mma kind::i8 instruction
 * (cta_group 1 / 2) issued back to back with no operand production: int8 ops per
 * second (2 per MAC) in *rate -- the tensor-pipe ceiling of dense_umma_kernel -- and the clock64 ticks per microsecond of
 * the issue loop in *sm_mhz, so that MACs per clock and SM follow without knowing the clock.  kind 8: clock64 ticks per
 * second (*rate) and per microsecond (*sm_mhz) over a 250 ms spin. */
int STORM_b200_microbench(int kind, double* rate, double* sm_mhz);
/* kind 6 / 7 of STORM_b200_microbench: tcgen05.mma kind::mxf4 (block-scaled E2M1, K = 64) at
 * cta_group 1 / 2, ops per second (2 per MAC).
 *
 * Exactness probe of that instruction for bit counting (fp4_probe.cu): each case is three
 * uint32 {n_full, n_single, pattern}: n_full instructions that add 64 to each of the 128 x 256
 * fp32 accumulators, then n_single that add 1; pattern 0..3 picks the operand encodings
 * (1.0 x 1.0, 0.5 x 2.0, 2.0 x 0.5, alternating).  Per case four 32-bit results: expected value,
 * smallest and largest accumulator (floats) and the number of accumulators != expected (uint32). */
int STORM_b200_fp4_probe(const uint32_t* cases, uint32_t n_cases, float* results);
/* The same question with data-dependent increments, as the tile kernel produces them: n_steps (<= 262143) instructions
 * at cta_group cg (1 or 2) over 4 x 4 pseudo-random operand combinations in the production nibble encoding, every
 * accumulator element adding 0 .. 64 per instruction (the all-ones element reaches 64 n_steps).  results[0..2]: largest
 * expected element, smallest and largest (got - expected) as floats; results[3]: number of wrong elements (u32 bits). */
int STORM_b200_fp4_probe_random(int cg, uint32_t n_steps, uint32_t seed, float* results);
/* The one-time per-device check behind KERNEL_AUTO's choice of the FP4 form: 1 if accumulators driven
 * to 2^24 - 1 in steps of 64 and of 1 came out exact for every operand encoding (cta_group 1), and driven to
 * 2^24 - 64 by data-dependent increments at cta_group 1 and 2 (the probe above), else 0. */
int STORM_b200_fp4_selftest(void);
/* cta_group of the UMMA kernel: 2 (default) = CTA pair per 256 x 256 tile, 1 = one CTA per
 * 128 x 256 tile.  Returns the previous value. */
int STORM_b200_set_umma_cta_group(int cg);
/* Code variant of the UMMA kernel (bit 0: hardware-suspended mbarrier waits; bit 1: scaled
 * bit expansion, see dense_umma.cu).  Results are identical for every value.  Returns the
 * previous value. */
int STORM_b200_set_umma_variant(int variant);
/* 1 (default): the persistent UMMA kernel keeps the tiles of a wave in step through K so that
 * they share their row blocks in L2 (a bounded wait on a device counter, results never depend
 * on it); 0: free-running CTAs.  Returns the previous value. */
int STORM_b200_set_umma_wave_sync(int on);
/* 1 (default): total-only queries with few tiles per SM split (tile, K chunk) units evenly over the
 * persistent CTAs (stream-K: no tail wave, no idle SMs on small row counts); 0: whole tiles only.
 * Results are identical.  Returns the previous value. */
int STORM_b200_set_umma_stream_k(int on);
/* 1 (default): total-only queries keep accumulating consecutive interior tiles of a CTA in the same
 * tensor-memory accumulator and drain it once per run (as many tiles as the accumulator's exact range
 * allows: 2^24 / (32 M) for the FP4 form) instead of once per tile; 0: one drain per tile.  Results are
 * identical.  Returns the previous value. */
int STORM_b200_set_umma_chain(int on);
/* SMs the persistent UMMA kernel leaves free (default 0): a multi-GPU caller that overlaps an NVLink
 * all-gather with the tile kernel sets this to a small number so that the collective's CTAs find a place
 * to run beside it.  Clamped to [0, SM count - 2].  Returns the previous value. */
int STORM_b200_set_umma_reserved_sms(int n);
/* L2 eviction hints on the packed-row TMA loads of triangle queries whose rows do not fit in L2 (the 8 column blocks of a
 * raster group are shared by every wave of the group, the row blocks change from wave to wave): 0 none, 1 column blocks
 * evict_last, 2 (default) + row blocks evict_first -- C3: DRAM reads 279 -> 176 GB per query, L2 hit rate 79.8 -> 87.1 %,
 * +1 % throughput (profiles/r02_ab_l2_hints_c3.jsonl).  Results are identical.  Returns the previous value. */
int STORM_b200_set_umma_l2_hints(int mode);
/* Clock probe of the tensor kernels: with it on, every launch records per CTA the clock64 and %globaltimer deltas around
 * its main loop; STORM_b200_last_kernel_clock waits for the device and returns the clock the last probed launch on the
 * current device ran at, in clock64 ticks per microsecond (mean over CTAs; optional min / max).  STORM_b200_microbench(8)
 * gives the same figure for a 250 ms spin on an otherwise idle device (to be read beside nvidia-smi's clocks.sm: is one
 * clock64 tick one SM cycle?).  The roofline of bench.py divides by this clock, not by nvidia-smi's. */
int STORM_b200_set_clock_probe(int on);
int STORM_b200_last_kernel_clock(double* mhz, double* min_mhz, double* max_mhz);
/* Number of kernel launches issued by this library since load (for bench.py). */
uint64_t STORM_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* STORM_B200_H_ */
