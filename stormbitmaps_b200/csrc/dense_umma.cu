// dense_umma.cu -- tensor-core tile kernel: tcgen05.mma kind::i8 over bits that are
// unpacked to {0,1} bytes on the fly.
//
// XX^T over a binary matrix is an integer GEMM: popcount(row_i & row_j) =
// sum_k bit(i,k) * bit(j,k).  Replaces the same loop nest as dense_popc.cu
// (storm.c:1165-1169 / 1199-1238 + libalgebra.h:2684-2744) with UMMA tiles.
// Accumulation is exact: u8 x u8 products into s32 accumulators in tensor memory,
// and a pair count is at most M < 2^31.
//
// Structure.  Persistent CTAs (one per SM, or one CTA pair
// per SM pair) walk the tile list; per CTA:
//   TMA warp    one thread streams PACKED rows into shared memory with
//               cp.async.bulk.tensor (SWIZZLE_128B boxes of 128 rows x 128 bytes =
//               8 k-blocks of 128 bits), a ring of 2-3 boxes, completion on an mbarrier.  It also decides
//               where the accumulator runs of total-only jobs end (run flag ring).
//               Out-of-range rows / columns are zero-filled by the TMA unit.
//   warps 0-3   "A expanders": thread = one A row = one TMEM lane.  Each k-block is
//               read back from the staged box (16 B, conflict-free thanks to the
//               swizzle), expanded with (w >> j) & 0x01010101 into 32 registers and
//               written to tensor memory with tcgen05.st (the A operand lives in TMEM:
//               no shared-memory write + read for it).  After the K loop of a tile the
//               same warps run its epilogue (tcgen05.ld, mask, sum / per-pair store).
//   warps 4..   "B expanders": thread = one B row.  Same expansion, written with
//               st.shared.v4 into the canonical K-major SWIZZLE_128B layout the UMMA
//               shared-memory descriptor expects (16-byte chunk c of row r lands at
//               chunk c ^ (r & 7) of its 128-byte line; 8-row groups are 1024 B apart).
//   MMA warp    TMEM allocation and, in the leader CTA, the single thread that issues
//               tcgen05.mma (4 per k-block, K = 32 bytes each) and tcgen05.commit.
//   Hand-over is by mbarriers only: raw_full/raw_empty (TMA <-> expanders),
//   full[s]/empty[s] (expanders <-> MMA, tcgen05.commit), acc_full/acc_empty
//   (MMA <-> epilogue).
//
// Bit order: within a 32-bit word, output register j holds bits j, j+8, j+16, j+24 as
// its four bytes.  A and B use the same permutation of K, and a dot product is
// invariant under a common permutation of its terms, so no un-shuffling is needed.
//
// CG = 2 (cta_group::2): two CTAs of a cluster share one 256 x 256 tile; each holds
// 128 A rows in its own TMEM and supplies 128 of the 256 B rows from its own shared
// memory, which halves the per-SM expansion work and shared-memory reads per MMA.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace storm {
namespace {

constexpr int UM_N = 256;               // B rows (accumulator columns) per tile
constexpr int UM_ACC_COL = 0;           // accumulator: columns [0, 256)
constexpr int UM_A_COL = 256;           // A stage s: columns [256 + 32 s, 256 + 32 s + 32)
constexpr int UM_SF_COL = 480;          // FP4 form: 32 columns of UE8M0 scale factors, all 127 (x 1.0)
constexpr int UM_SF_COLS = 32;
// k-block = the packed bits of one row that expand to one 128-byte swizzle line (4 MMAs):
// 128 bits (16 B) as bytes for kind::i8, 256 bits (32 B) as nibbles for kind::mxf4.  A TMA box is
// 128 packed bytes wide, i.e. 8 resp. 4 k-blocks.
constexpr int UM_CHUNK_KB = 8;
constexpr int UM_CHUNK_KB_FP4 = 4;
#ifndef STORM_RAW_BUFS
#define STORM_RAW_BUFS 3
#endif

// XW = expander warps per 32 rows: 1 = a thread expands its row's whole k-block (4 K steps), 2 = two
// threads of different warps expand K steps {0, 1} and {2, 3} of it, which halves the time from "stage free"
// to "stage full" (the FP4 form expands twice the bits per MMA and fell 15 % short of the pipe with XW = 1).
//
// PAIRS = the per-pair output form (DenseJob::out set: every tile is drained, nothing can be chained or split along K),
// a separate instantiation so that the total-only kernel carries none of its code.  Its drain does not store from
// registers (a thread owns one accumulator ROW, so a warp-wide store touched 32 different lines -- the drain of a tile
// was bound by the LSU at ~2 clocks per 32-byte sector, ~9000 clocks per 256 x 256 tile): each 32 x 32 chunk of counts
// goes through the shared-memory slices of the B stages -- idle while a tile is drained, and already laid out in
// SWIZZLE_128B atoms of 8 rows x 128 bytes -- and leaves as ONE cp.async.bulk.tensor store of a 32 x 32 box, clipped
// by the TMA unit at the edges of the output matrix.
// (Tried first, profiles/r02_rect_output_n128_*.jsonl: 128-column tiles with two accumulators and dedicated epilogue
// warps, so that the drain overlaps the next tile.  The drain did vanish, but an expander has only 256 clocks per
// k-block then and its serial chain -- box read, barrier wait, tcgen05.st, wait::st, fence, arrive -- takes ~600: the
// tensor pipe ran 64 % busy, 7.0 ms for 12288^2 counts at 131072 bits where this form takes 4.6 ms.)
template <int CG, int XW = 1, bool FP4 = false, bool PAIRS = false>
struct Cfg {
    static constexpr bool FP4_FORM = FP4;
    static constexpr int TN = UM_N;                          // B rows (accumulator columns) per tile
    static constexpr int ACCS = 1;                           // accumulators of TN columns
    static constexpr int A_ROW_WARPS = 4;                    // 128 A rows = TMEM lanes
    static constexpr int B_ROWS = TN / CG;                   // B rows expanded by this CTA
    static constexpr int B_ROW_WARPS = B_ROWS / 32;
    static constexpr int A_WARPS = A_ROW_WARPS * XW;
    static constexpr int B_WARPS = B_ROW_WARPS * XW;
    static constexpr int MMA_WARP = A_WARPS + B_WARPS;
    static constexpr int TMA_WARP = MMA_WARP + 1;
    static constexpr int N_WARPS = A_WARPS + B_WARPS + 2;
    static constexpr int THREADS = N_WARPS * 32;
    // expanded k-blocks in flight between the expanders and the MMA thread (A: 32 TMEM columns each,
    // next to the 256 accumulator columns and, in the FP4 form, 32 scale-factor columns)
    static constexpr int STAGES = CG == 2 ? (FP4 && !PAIRS ? 7 : 6) : 3;   // (the per-pair form gives one stage's memory to its staging slices)
    static constexpr int STAGE_BYTES = B_ROWS * 128;         // expanded B rows of one k-block
    // per-pair form: counts leave through TMA stores when each TMEM lane quarter has an A warp and a B warp to share
    // the B-stage slices of that quarter (the default cta_group::2 form with one expander warp per 32 rows)
    static constexpr bool TMA_DRAIN = PAIRS && CG == 2 && XW == 1 && STAGES >= 6;

    static constexpr int RAW_A_BYTES = 128 * 128;            // one box of packed A rows
    static constexpr int RAW_B_BYTES = B_ROWS * 128;
    static constexpr int RAW_BYTES = RAW_A_BYTES + RAW_B_BYTES;
    static constexpr int TM = 128 * CG;
    static constexpr int EXPANDER_WARPS = A_WARPS + B_WARPS;
    static constexpr uint32_t OFF_RAW = (STAGES * STAGE_BYTES + 1023) / 1024 * 1024;
    // packed-row boxes in flight between the TMA thread and the expanders (how far the loads run ahead of the
    // expansion: one box = 4 (FP4) or 8 k-blocks); three fit beside the stages of the CTA-pair form
    static constexpr int RAW_BUFS = CG == 2 ? (TMA_DRAIN ? 2 : STORM_RAW_BUFS) : 2;
    // per-pair form: two 4 KiB staging slices per expander warp for the counts on their way to a TMA store (the third
    // packed-row box and the seventh stage make room: the box is worth < 2 %, profiles/r01_ab_raw_box_ring.jsonl)
    static constexpr uint32_t OFF_STG = OFF_RAW + RAW_BUFS * RAW_BYTES;
    static constexpr uint32_t STG_BYTES = TMA_DRAIN ? (uint32_t)EXPANDER_WARPS * 2u * 4096u : 0u;   // two slices per warp
    static constexpr uint32_t OFF_BAR = OFF_STG + STG_BYTES;
    static constexpr uint32_t SMEM_BYTES = 1024 /*align slack*/ + OFF_BAR + 512;
    static_assert(SMEM_BYTES <= 232448, "shared memory of one CTA");
    static_assert(TN * ACCS <= UM_A_COL, "accumulators overflow their tensor-memory columns");
    static_assert(UM_A_COL + 32 * STAGES <= (FP4 ? UM_SF_COL : 512), "A stages overflow tensor memory");
    static_assert(XW == 1 || XW == 2, "one or two expander warps per 32 rows");
    static_assert(N_WARPS <= 24, "per-warp reduction slots");
    // kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor bit layout):
    //   [4,6) c_format = 2 (S32); [7,10) a_format = 0 (u8); [10,13) b_format = 0 (u8);
    //   [15] a_major = 0 (K); [16] b_major = 0 (K); [17,23) N >> 3; [24,29) M >> 4
    static constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    // kind::mxf4 block-scaled descriptor (cute::UMMA::InstrDescriptorBlockScaled): [7,10) a_format = 1 (E2M1);
    //   [10,13) b_format = 1; K-major; [17,23) N >> 3; [23] scale format 1 (UE8M0); [24,29) M >> 4; [31] 0 = K 64
    static constexpr uint32_t IDESC_FP4 = (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | (1u << 23) |
                                          ((uint32_t)((128 * CG) >> 4) << 24);
};

// 32 bits -> 32 bytes of {0,1}: register j holds bits j, j+8, j+16, j+24.
__device__ __forceinline__ void expand32(uint32_t w, uint32_t (&r)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (w >> j) & 0x01010101u;
}
// Scaled expansion (VAR_SCALED): the shift is dropped on the A side.  A byte is 2^j where the plain
// form has 1 (one LOP3 per register), the matching B byte is 2^(7-j), so every matching bit
// contributes exactly 128 to the s32 accumulator and the epilogue shifts the count right by 7.
// Exact while 128 * M < 2^31.  The B side gets bit 7-j of each byte by reversing the bits of the
// word (BREV) and restoring the byte order (PRMT): two extra instructions per eight registers.
__device__ __forceinline__ void expand32_a_scaled(uint32_t w, uint32_t (&r)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = w & (0x01010101u << j);
}
__device__ __forceinline__ void expand32_b_scaled(uint32_t w, uint32_t (&r)[8]) {
    const uint32_t wr = __byte_perm(__brev(w), 0u, 0x0123);                // byte b keeps its place, bits reversed inside it
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = wr & (0x01010101u << (7 - j));
}

// FP4 form (VAR_FP4): 32 bits -> 32 E2M1 nibbles in four registers; register j holds bit j of every
// nibble of the word, moved to a nibble position whose E2M1 reading (0001 = 0.5, 0010 = 1, 0100 = 2)
// multiplies to exactly 1.0 with its partner on the other side:
//     A: 0.5, 1, 2, 1      B: 2, 1, 0.5, 1
// Bit 3 of a nibble is the E2M1 sign and cannot be used in place, so bit 3 has to move on both sides; with
// positions p (value 2^(p-1)) the pair (pA, pB) of a bit must satisfy pA + pB = 2, and the assignment above
// is the one with the fewest distinct shift amounts: one on the A side (>> 2), two on the B side (<< 2 and a
// >> 2 that serves two registers) -- 5 + 6 instructions per 32 bits.  (Round 1 used A 0.5, 1, 2, 2 x
// B 2, 1, 0.5, 0.5: three different shifts on the B side, 5 + 7.)  Every matching bit adds exactly 1.0f to an
// fp32 accumulator; integers below 2^24 are exact in fp32 and the tensor core's accumulation of them was
// measured to be exact (fp4_probe.cu, which encodes its operands the same way).
__device__ __forceinline__ void expand32_a_fp4(uint32_t w, uint32_t* r) {
    r[0] = w & 0x11111111u;
    r[1] = w & 0x22222222u;
    r[2] = w & 0x44444444u;
    r[3] = (w >> 2) & 0x22222222u;
}
__device__ __forceinline__ void expand32_b_fp4(uint32_t w, uint32_t* r) {
    const uint32_t d = w >> 2;
    r[0] = (w << 2) & 0x44444444u;
    r[1] = w & 0x22222222u;
    r[2] = d & 0x11111111u;
    r[3] = d & 0x22222222u;
}

// D[tmem] (+)= A[tmem] * B[smem]^T with E2M1 operands, UE8M0 scale factors per 32 elements from tensor
// memory (all 1.0 here), fp32 accumulation, K = 64 per instruction: twice the bits of kind::i8.
template <int CG>
__device__ __forceinline__ void umma_mxf4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}

// Work of one persistent CTA (pair) as a sequence of segments: K chunks [c0, c1) of one tile.  Default:
// whole tiles, dealt round-robin (cluster k takes tiles begin + k, begin + k + n, ...).  Stream-K
// (DenseJob::stream_k): only the full waves are dealt that way; the n_tiles mod n_clusters tiles of the tail
// wave (all tiles, if there are fewer tiles than clusters) are cut along K: their tiles x n_chunks units
// are split into one contiguous range per cluster, sizes differing by at most one chunk, so a cluster's
// first and last tail segments may cover part of a tile's K.  Every role of the kernel walks the same sequence.
struct Seg { uint64_t tile; uint32_t c0, c1; bool tail; };
struct SegWalk {
    uint64_t cur, full_end, step;          // whole tiles: cur, cur + step, ... below full_end
    uint64_t tail_cur, tail_end;           // tail units (tile-major, chunk-minor) of this cluster
    uint32_t n_chunks;
    __device__ __forceinline__ SegWalk(const DenseJob& job, uint64_t cluster_id, uint64_t n_clusters, uint32_t n_chunks_)
        : step(n_clusters), n_chunks(n_chunks_) {
        const uint64_t n_tiles = job.tile_end - job.tile_begin;
        const uint64_t n_full = job.stream_k ? n_tiles / n_clusters * n_clusters : n_tiles;
        cur = job.tile_begin + cluster_id;
        full_end = job.tile_begin + n_full;
        const uint64_t units = (n_tiles - n_full) * n_chunks;
        tail_cur = units / n_clusters * cluster_id + min(units % n_clusters, cluster_id);
        tail_end = tail_cur + units / n_clusters + (cluster_id < units % n_clusters ? 1 : 0);
    }
    __device__ __forceinline__ bool next(Seg& s) {
        if (cur < full_end) { s.tile = cur; s.c0 = 0; s.c1 = n_chunks; s.tail = false; cur += step; return true; }
        if (tail_cur >= tail_end) return false;
        s.tail = true;
        s.tile = full_end + tail_cur / n_chunks;
        s.c0 = (uint32_t)(tail_cur % n_chunks);
        const uint64_t left = tail_end - tail_cur;
        s.c1 = left < (uint64_t)(n_chunks - s.c0) ? s.c0 + (uint32_t)left : n_chunks;
        tail_cur += s.c1 - s.c0;
        return true;
    }
};

// Run flags handed from the expanders to the MMA warp, one ring entry per segment (the expanders are at most
// STAGES k-blocks, hence STAGES segments, ahead of the MMA warp).
constexpr uint32_t RUN_FIRST = 1, RUN_LAST = 2, RUN_RING = 16;

constexpr int VAR_SUSPEND = 1;          // hardware-suspended mbarrier waits
constexpr int VAR_SCALED = 2;           // scaled expansion (counts accumulate x128); kind::i8 only
constexpr int VAR_FP4 = 4;              // bits -> E2M1 nibbles, tcgen05.mma kind::mxf4, fp32 accumulators
constexpr int VAR_WIDE = 8;             // two expander warps per 32 rows (Cfg::XW = 2)
constexpr int VAR_PAIRS = 16;           // per-pair output form: 128-column tiles, two accumulators, epilogue warps (Cfg::PAIRS)

template <int CG, int VAR>
__global__ void __launch_bounds__((Cfg<CG, (VAR & VAR_WIDE) ? 2 : 1, (VAR & VAR_FP4) != 0, (VAR & VAR_PAIRS) != 0>::THREADS), 1)
dense_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_out, const DenseJob job) {
    using C = Cfg<CG, (VAR & VAR_WIDE) ? 2 : 1, (VAR & VAR_FP4) != 0, (VAR & VAR_PAIRS) != 0>;
    constexpr int XW = (VAR & VAR_WIDE) ? 2 : 1;
    constexpr bool PAIRS = (VAR & VAR_PAIRS) != 0;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B needs 1024-byte alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));        // generic pointer to the aligned base
    const uint32_t raw_base = smem_base + C::OFF_RAW;                       // [RAW_BUFS][A box | B box]
    const uint32_t bar_base = smem_base + C::OFF_BAR;
    const uint32_t full_bar = bar_base;                                     // STAGES x 8 B
    const uint32_t empty_bar = full_bar + 8 * C::STAGES;
    const uint32_t raw_full_bar = empty_bar + 8 * C::STAGES;                // RAW_BUFS x 8 B
    const uint32_t raw_empty_bar = raw_full_bar + 8 * C::RAW_BUFS;          // RAW_BUFS x 8 B
    const uint32_t acc_full_bar = raw_empty_bar + 8 * C::RAW_BUFS;          // ACCS x 8 B
    const uint32_t acc_empty_bar = acc_full_bar + 8 * C::ACCS;              // ACCS x 8 B
    const uint32_t tmem_slot = acc_empty_bar + 8 * C::ACCS;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    unsigned long long* red = reinterpret_cast<unsigned long long*>(smem_gen + (tmem_slot + 8 - smem_base));   // one slot per warp (<= 24)
    const uint32_t run_ring = tmem_slot + 8 + 8 * 24;                       // RUN_RING x 4 B: run flags, expanders -> MMA warp
    static_assert(8 * (2 * C::STAGES + 2 * C::RAW_BUFS + 2 * C::ACCS) + 8 + 8 * 24 + 4 * RUN_RING <= 512, "barrier block");

    auto wait = [](uint32_t bar, uint32_t parity) { mbar_wait_t<(VAR & VAR_SUSPEND) != 0>(bar, parity); };
    constexpr bool FP4 = (VAR & VAR_FP4) != 0;
    constexpr bool SCALED = !FP4 && (VAR & VAR_SCALED) != 0;
    constexpr uint32_t CHUNK_KB = FP4 ? UM_CHUNK_KB_FP4 : UM_CHUNK_KB;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const uint64_t cluster_id = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
    const uint64_t n_clusters = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;
    const uint32_t n_kb = FP4 ? (job.n_words + 3) / 4 : (job.n_words + 1) / 2;
    const uint32_t n_chunks = (n_kb + CHUNK_KB - 1) / CHUNK_KB;

    // ---- setup ----------------------------------------------------------------
    if (warp == C::MMA_WARP) tmem_alloc<CG>(tmem_slot);
    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar + 8 * s, CG * C::EXPANDER_WARPS);           // every expander warp of the pair
            mbar_init(empty_bar + 8 * s, 1);                               // tcgen05.commit
        }
        for (int b = 0; b < C::RAW_BUFS; ++b) {
            mbar_init(raw_full_bar + 8 * b, 1);                            // expect_tx arrive + TMA bytes
            mbar_init(raw_empty_bar + 8 * b, C::EXPANDER_WARPS);
        }
        for (int a = 0; a < C::ACCS; ++a) {
            mbar_init(acc_full_bar + 8 * a, 1);                            // tcgen05.commit
            mbar_init(acc_empty_bar + 8 * a, CG * C::EXPANDER_WARPS);      // every epilogue warp of the pair
        }
        fence_mbar_init();
    }
    if (warp == C::TMA_WARP && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (C::TMA_DRAIN && job.out_tma) tma_prefetch_desc(&map_out);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // clock probe: the (mostly idle) TMA thread brackets the kernel's main part with clock64 / %globaltimer
    const bool clk_thread = job.clk != nullptr && tid == (uint32_t)C::TMA_WARP * 32u;
    long long clk0 = 0;
    uint64_t gt0 = 0;
    if (clk_thread) { clk0 = clock64(); gt0 = global_timer_ns(); }

    unsigned long long sum = 0;
    // interior tile: every row and column valid, nothing on or below the diagonal, no per-pair output
    auto is_interior = [&](uint32_t bi_, uint32_t bj_) {
        const uint64_t a0 = (uint64_t)bi_ * C::TM, b0 = (uint64_t)bj_ * C::TN;
        return job.out == nullptr && a0 + C::TM <= job.nA && b0 + C::TN <= job.nB &&
               (!job.strict_upper || job.j_off + b0 >= job.i_off + a0 + C::TM);
    };

    // ---- epilogue of one run: TMEM -> registers -> masked sum / per-pair store ----------------------------------
    // Warp w may read TMEM lanes 32 (w % 4) .. +31 (`quarter`); the `n_sharers` warps of a quarter split the accumulator's
    // columns between them in 32-column chunks dealt round-robin (`sharer` = this warp's turn).  Total-only form: every
    // expander warp takes part (the A warp and the B warp(s) of a quarter), so the drain takes half (a third) as long.
    // Per-pair form: the four dedicated epilogue warps, one per quarter, on the accumulator half `acc_col`.
    auto drain = [&](uint32_t acc_col, uint32_t quarter, uint32_t sharer, uint32_t n_sharers, uint32_t bi, uint32_t bj,
                     bool interior, bool fp4_sum_exact) {
        const uint64_t rowA0 = (uint64_t)bi * C::TM, rowB0 = (uint64_t)bj * C::TN;
        const uint64_t li = rowA0 + rank * 128u + quarter * 32u + lane;      // accumulator row of this thread (= its TMEM lane)
        const uint64_t gi = job.i_off + li;
        const bool row_ok = li < job.nA;
        const uint32_t acc_lane = tmem_base + ((quarter * 32u) << 16) + UM_ACC_COL + acc_col;
        const bool out_vec = ((reinterpret_cast<uintptr_t>(job.out) | (job.ld * 4)) & 31) == 0;   // 32-byte stores possible
#pragma unroll 1
        for (uint32_t c0 = sharer * 32u; c0 < (uint32_t)C::TN; c0 += 32u * n_sharers) {
            uint32_t v[32];
            tmem_ld32(acc_lane + c0, v);
            tc_wait_ld();
            if constexpr (C::TMA_DRAIN) {
                if (job.out_tma) {
                    // Per-pair output through the TMA unit.  The columns of this chunk that count for this row are
                    // [lo, lo + span) (below nB, right of the diagonal); everything else is stored as 0.  The warp's
                    // 32 rows x 32 counts go into its 4 KiB staging slice -- 32 lines of 128 bytes in SWIZZLE_128B atoms --
                    // and one lane issues a single 32 x 32 box store.
                    const long long cb = (long long)(rowB0 + c0);
                    const long long h = (long long)job.nB - cb;
                    const int hi = row_ok ? (h < 0 ? 0 : h > 32 ? 32 : (int)h) : 0;
                    int lo = 0;
                    if (job.strict_upper) {
                        const long long l = (long long)gi - (long long)job.j_off - cb + 1;
                        lo = l < 0 ? 0 : l > 32 ? 32 : (int)l;
                    }
                    const uint32_t span = (uint32_t)(hi > lo ? hi - lo : 0);
                    // With the stores out of the way the conversions showed (F2I runs at 16 per clock and SM: 2048 clocks
                    // per tile; 64-bit adds and per-element masks as much again, profiles/r02_pairs_tma_4096_ncu.md): an
                    // fp32 accumulator holding an integer below 2^23 is converted by one FADD with 2^23 and one integer
                    // subtract (the integer sits in the mantissa); chunks that lie wholly inside the matrix and right of
                    // the diagonal -- all but the edge chunks -- skip the masks; a chunk's 32 counts are summed in 32 bits.
                    const bool whole = __all_sync(0xffffffffu, lo == 0 && span == 32u);
                    const bool magic = FP4 && job.n_words < (1u << 17);            // every count below 2^23
                    uint32_t part = 0;
                    if (whole && magic) {
#pragma unroll
                        for (int cc = 0; cc < 32; ++cc) {
                            v[cc] = __float_as_uint(__uint_as_float(v[cc]) + 8388608.0f) - 0x4B000000u;
                            part += v[cc];
                        }
                        sum += part;
                    } else {
#pragma unroll
                        for (int cc = 0; cc < 32; ++cc) {
                            const uint32_t x = FP4 ? __float2uint_rn(__uint_as_float(v[cc])) : SCALED ? (v[cc] >> 7) : v[cc];
                            v[cc] = ((uint32_t)(cc - lo) < span) ? x : 0u;
                            if (FP4) part += v[cc]; else sum += v[cc];
                        }
                        sum += part;
                    }
                    // this warp's two staging slices in turn: free again once the store before the last has read it (with
                    // one slice every chunk waited ~400 clocks for the TMA unit, profiles/r02_pairs_ahead_4096_ncu.md)
                    if (lane == 0) bulk_wait_read<1>();
                    __syncwarp();
                    const uint32_t slice = smem_base + C::OFF_STG + (warp * 2u + ((c0 / (32u * n_sharers)) & 1u)) * 4096u;
#pragma unroll
                    for (uint32_t j = 0; j < 8; ++j)
                        st_shared_v4(slice + lane * 128u + ((j ^ (lane & 7u)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async_smem();                              // generic writes -> visible to the TMA unit (async proxy)
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_out, slice, (uint32_t)(rowB0 + c0), (uint32_t)(rowA0 + rank * 128u + quarter * 32u));
                        bulk_commit();
                    }
                    continue;
                }
            }
            if (!PAIRS && interior) {
                if constexpr (FP4) {
                    // fp32 accumulators holding exact integers.  While 32 counts cannot exceed 2^24 their
                    // float sum is exact too (one conversion per 32 columns); otherwise convert one by one.
                    if (fp4_sum_exact) {
                        float part = 0.0f;
#pragma unroll
                        for (int cc = 0; cc < 32; ++cc) part += __uint_as_float(v[cc]);
                        sum += __float2uint_rn(part);
                    } else {
#pragma unroll
                        for (int cc = 0; cc < 32; ++cc) sum += __float2uint_rn(__uint_as_float(v[cc]));
                    }
                } else if (SCALED) {                               // 32 counts of at most 2^24 each fit 32 bits
                    uint32_t part = 0;
#pragma unroll
                    for (int cc = 0; cc < 32; ++cc) part += v[cc] >> 7;
                    sum += part;
                } else {
#pragma unroll
                    for (int cc = 0; cc < 32; ++cc) sum += v[cc];
                }
            } else if (!PAIRS || job.out == nullptr) {
                // Edge or diagonal tile, total only: the columns of this 32-column chunk that count are
                // [lo, lo + span) -- those below nB and right of the diagonal for this row (one range
                // compare per column instead of 64-bit index arithmetic: stream-K can hand a run of
                // diagonal tiles to one CTA, and the slow path made that CTA the last to finish).
                const long long cb = (long long)(rowB0 + c0);
                const long long h = (long long)job.nB - cb;
                const int hi = row_ok ? (h < 0 ? 0 : h > 32 ? 32 : (int)h) : 0;
                int lo = 0;
                if (job.strict_upper) {
                    const long long l = (long long)gi - (long long)job.j_off - cb + 1;
                    lo = l < 0 ? 0 : l > 32 ? 32 : (int)l;
                }
                const uint32_t span = (uint32_t)(hi > lo ? hi - lo : 0);
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                    const uint32_t x = FP4 ? __float2uint_rn(__uint_as_float(v[cc])) : SCALED ? (v[cc] >> 7) : v[cc];
                    sum += ((uint32_t)(cc - lo) < span) ? x : 0u;
                }
            } else if (row_ok && out_vec && rowB0 + c0 + 32 <= job.nB &&
                       (!job.strict_upper || job.j_off + rowB0 + c0 > gi)) {
                // Per-pair output, all 32 columns of the chunk valid and right of the diagonal: this thread's
                // 128 consecutive bytes of its output row go out as four 32-byte stores (whole sectors; the
                // scalar form below issues 32 stores that each touch 32 different lines across the warp
                // and held the tensor pipe up for a third of a 2048-word tile).
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                    v[cc] = FP4 ? __float2uint_rn(__uint_as_float(v[cc])) : SCALED ? (v[cc] >> 7) : v[cc];
                    sum += v[cc];
                }
                uint32_t* dst = job.out + li * job.ld + (rowB0 + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) st_global_v8(dst + 8 * q, &v[8 * q]);
            } else if (row_ok) {
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                    const uint64_t lj = rowB0 + c0 + cc;
                    if (lj < job.nB) {
                        uint32_t x = FP4 ? __float2uint_rn(__uint_as_float(v[cc])) : SCALED ? (v[cc] >> 7) : v[cc];
                        if (job.strict_upper && job.j_off + lj <= gi) x = 0;
                        sum += x;
                        job.out[li * job.ld + lj] = x;
                    }
                }
            }
        }
    };

    if (FP4 && warp < C::A_ROW_WARPS) {
        // Scale factors of the block-scaled MMA: every byte of the region is UE8M0 127 = 2^0, so whatever
        // lane / column / byte the hardware layout of SFA and SFB assigns to a (row, 32-element block),
        // it reads 1.0.  Written once; the MMA thread first touches it after the full-barrier hand-over.
        uint32_t sf[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sf[j] = 0x7F7F7F7Fu;
        const uint32_t sf_lane = tmem_base + ((warp * 32u) << 16) + UM_SF_COL;
#pragma unroll
        for (int c = 0; c < UM_SF_COLS; c += 8) tmem_st8(sf_lane + c, sf);
        tc_wait_st();
        tc_fence_before();
    }

    if (warp == C::TMA_WARP) {
        // ===== TMA producer: packed rows -> shared memory ==============================
        if (lane == 0) {
            uint32_t buf = 0, buf_phase = 0;                               // ring of packed-row boxes and its parity
            const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
            uint32_t t_iter = 0, run_pos = 0;                              // segments done; segments already in the open run
            bool in_step = job.wave_sync != nullptr;
            SegWalk walk(job, cluster_id, n_clusters, n_chunks);
            TileCursor cursor;
            Seg seg, seg_next;
            uint32_t bi = 0, bj = 0, bi_next = 0, bj_next = 0;
            bool interior = false, interior_next = false, have = false, primed = false;
            for (;;) {
                // This thread runs ahead of everyone else and is idle most of the time, so it also decides where
                // the accumulator runs of a total-only job end (DenseJob::chain_max): with one segment of
                // look-ahead, a run goes on while this segment and the next are interior and the accumulator has
                // room.  {first, last} go into the flag ring before the segment's first box is requested; the
                // expanders read them after that box has landed, the MMA warp after the segment's first stage is
                // full (release / acquire through the mbarriers in between).  The first pass only fills the look-ahead.
                const bool have_next = walk.next(seg_next);
                if (have_next) { cursor.coords(job, seg_next.tile, C::TM, C::TN, bi_next, bj_next); interior_next = is_interior(bi_next, bj_next); }
                if (primed) {
                    const bool run_last = !(have_next && interior && interior_next && run_pos + 1 < job.chain_max);
                    st_shared_u32(run_ring + 4 * (t_iter & (RUN_RING - 1)), (run_pos == 0 ? RUN_FIRST : 0u) | (run_last ? RUN_LAST : 0u));
                    run_pos = run_last ? 0 : run_pos + 1;
                    if (seg.tail) in_step = false;                         // stream-K tail: uneven segments, nothing to keep in step
                    if (in_step && t_iter > 0) {
                        // Wave barrier.  The tiles of one wave share 8 A and ~9 B row blocks; they only find each
                        // other's lines in L2 if they walk K in step, and without this the CTAs drift apart over
                        // the thousands of tiles of a large query (ncu: 50 % L2 hits, 1.2 TB of DRAM reads on C3).
                        // Everything staged so far keeps the MMA busy while this thread waits.  Bounded: if some
                        // CTA is not resident (SMs taken by another kernel) the hint is dropped, not the query.
                        const uint32_t target = t_iter * gridDim.x;
                        uint32_t spins = 0;
                        while (*reinterpret_cast<volatile unsigned int*>(job.wave_sync) < target) {
                            if (++spins > 8192u) { in_step = false; break; }
                            __nanosleep(64);
                        }
                    }
                    const uint32_t ya = bi * C::TM + rank * 128u;
                    const uint32_t yb = bj * C::TN + rank * C::B_ROWS;
                    for (uint32_t c = seg.c0; c < seg.c1; ++c) {
                        wait(raw_empty_bar + 8 * buf, buf_phase ^ 1);
                        mbar_expect_tx(raw_full_bar + 8 * buf, C::RAW_BYTES);
                        const uint32_t dst = raw_base + buf * C::RAW_BYTES;
                        if (job.l2_hints) {
                            if (job.l2_hints > 1) tma_load_2d_hint(dst, &map_a, c * 128u, ya, raw_full_bar + 8 * buf, pol_a);
                            else tma_load_2d(dst, &map_a, c * 128u, ya, raw_full_bar + 8 * buf);
                            tma_load_2d_hint(dst + C::RAW_A_BYTES, &map_b, c * 128u, yb, raw_full_bar + 8 * buf, pol_b);
                        } else {
                            tma_load_2d(dst, &map_a, c * 128u, ya, raw_full_bar + 8 * buf);
                            tma_load_2d(dst + C::RAW_A_BYTES, &map_b, c * 128u, yb, raw_full_bar + 8 * buf);
                        }
                        if (++buf == (uint32_t)C::RAW_BUFS) { buf = 0; buf_phase ^= 1; }
                    }
                    if (job.wave_sync && !seg.tail) atomicAdd(job.wave_sync, 1u);   // this CTA's loads of the wave are in flight
                    ++t_iter;
                }
                primed = true;
                seg = seg_next; bi = bi_next; bj = bj_next; interior = interior_next; have = have_next;
                if (!have) break;
            }
        }
    } else if (warp == C::MMA_WARP) {
        // ===== MMA issuer: the MMA warp of the leader CTA ==============================
        // The whole warp walks the loop (waits included) and one elected lane issues: with warp-uniform
        // control flow the stage counters, descriptors and tensor-memory addresses live in uniform
        // registers, where tcgen05.mma wants them.  With a single thread in a divergent branch every
        // operand took an ELECT + R2UR.BROADCAST detour and the loop needed ~90 % of a k-block's MMA time
        // (ncu source view, profiles/r01_fp4_c3_ncu_full.md): any hiccup starved the tensor pipe.
        if (rank == 0) {
            const bool leader = elect_one();
            // K-major SWIZZLE_128B shared-memory descriptor (cute::UMMA::SmemDescriptor):
            //   [0,14) addr >> 4; [16,30) LBO >> 4 = 1 (unused for swizzled K-major); [32,46) SBO >> 4 = 64
            //   (1024 B between 8-row groups); [46,48) version = 1; [61,64) layout = 2 (SWIZZLE_128B)
            const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            uint32_t s = 0, phase = 0, t_iter = 0, run_iter = 0;           // stage ring position and its parity; runs done
            // Everything the four MMAs of a k-block need is carried from iteration to iteration instead of being
            // rebuilt from s after the wait: low descriptor word of the stage's B lines (LBO | addr >> 4, + 2 per K step
            // of 32 bytes; shared-memory addresses stay below 2^18, so the 14-bit field cannot carry), tensor-memory
            // column of its A operand, its two barriers.  The issue loop of this warp runs next to the expander
            // warps of its scheduler and its length is what the tensor pipe's duty cycle hangs on: three extra
            // instructions in it cost 2 % on every shape (profiles/r01_ab_chain_v2.jsonl).
            const uint32_t b_lo0 = ((smem_base >> 4) & 0x3FFFu) | (uint32_t)desc_hi, a_col0 = tmem_u + UM_A_COL;
            uint32_t b_lo = b_lo0, a_col = a_col0, full_s = full_bar, empty_s = empty_bar;
            SegWalk walk(job, cluster_id, n_clusters, n_chunks);
            Seg seg;
            for (; walk.next(seg); ++t_iter) {
                const uint32_t kb0 = seg.c0 * CHUNK_KB, kb1 = min(n_kb, seg.c1 * CHUNK_KB);
                // Run flags of this segment, published by the TMA thread before it requested the segment's first box
                // (hence the wait for the first stage here; the loop's own wait on it then passes at once).
                // A run = consecutive interior segments of a total-only job that share the accumulator: only its
                // first segment waits for the drain of the previous run and overwrites, only its last one hands
                // the accumulator to the epilogue.  Read outside the k loop.
                wait(full_s, phase);
                const uint32_t flags = ld_shared_u32(run_ring + 4 * (t_iter & (RUN_RING - 1)));
                const bool run_first = __any_sync(0xffffffffu, (flags & RUN_FIRST) != 0);      // (warp-uniform anyway)
                // accumulator of this run: the only one, or (per-pair form) the two 128-column halves in turn
                const uint32_t acc_sel = C::ACCS == 2 ? (run_iter & 1u) : 0u;
                const uint32_t acc_use = C::ACCS == 2 ? (run_iter >> 1) : run_iter;      // how often this accumulator has been used
                const uint32_t acc_col = tmem_u + UM_ACC_COL + acc_sel * (uint32_t)C::TN;
                if (run_first) wait(acc_empty_bar + 8 * acc_sel, (acc_use & 1) ^ 1);     // the epilogue of its previous run has drained it
                tc_fence_after();
                uint32_t fresh = run_first ? 0u : 1u;                      // accumulate flag of the segment's first MMA
                for (uint32_t kb = kb0; kb < kb1; ++kb) {
                    wait(full_s, phase);
                    tc_fence_after();
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t b_desc = (desc_hi & 0xFFFFFFFF00000000ull) | (uint64_t)(b_lo + 2 * k);
                            const uint32_t accumulate = k == 0 ? fresh : 1u;
                            if (FP4)
                                umma_mxf4_ts<CG>(acc_col, a_col + k * 8, b_desc, C::IDESC_FP4,
                                                 tmem_u + UM_SF_COL, tmem_u + UM_SF_COL + UM_SF_COLS / 2, accumulate);
                            else
                                umma_i8_ts<CG>(acc_col, a_col + k * 8, b_desc, C::IDESC, accumulate);
                        }
                        umma_commit<CG>(empty_s);                          // frees the stage when these MMAs are done
                    }
                    __syncwarp();
                    fresh = 1u;
                    if (++s == (uint32_t)C::STAGES) {
                        s = 0; phase ^= 1;
                        b_lo = b_lo0; a_col = a_col0; full_s = full_bar; empty_s = empty_bar;
                    } else {
                        b_lo += C::STAGE_BYTES >> 4; a_col += 32; full_s += 8; empty_s += 8;
                    }
                }
                if (flags & RUN_LAST) {
                    if (leader) umma_commit<CG>(acc_full_bar + 8 * acc_sel);   // accumulator of this run complete
                    ++run_iter;
                }
                __syncwarp();
            }
        }
    } else if (warp < (uint32_t)C::EXPANDER_WARPS) {
        // ===== expanders (A: warps 0-3 -> TMEM, B: warps 4.. -> shared memory) ==========
        // warp -> (side, K half, 32-row group): A warps [0, 4 XW), then B warps; within a side the row group
        // varies fastest, so that an A warp's TMEM lane quarter (warp % 4) is its row group
        const bool is_a = warp < C::A_WARPS;
        const uint32_t side_warp = is_a ? warp : warp - C::A_WARPS;
        const uint32_t row_warps = is_a ? (uint32_t)C::A_ROW_WARPS : (uint32_t)C::B_ROW_WARPS;
        const uint32_t half = side_warp / row_warps;                       // which K steps of a k-block (always 0 for XW = 1)
        const uint32_t idx = (side_warp % row_warps) * 32u + lane;         // row within this CTA's A / B slice
        const uint32_t raw_row = (is_a ? 0u : (uint32_t)C::RAW_A_BYTES) + idx * 128u;
        const uint32_t sw = idx & 7u;
        const uint32_t b_line = (idx >> 3) * 1024u + (idx & 7u) * 128u;    // 8-row groups are 1024 B apart
        const uint32_t a_lane = tmem_base + (((warp & 3u) * 32u) << 16);
        uint32_t s = 0, phase = 0, t_iter = 0;                             // stage ring position and its parity
        uint32_t buf = 0, buf_phase = 0;                                   // ring of packed-row boxes and its parity
        uint32_t run_iter = 0;                                             // runs drained
        SegWalk walk(job, cluster_id, n_clusters, n_chunks);
        TileCursor cursor;
        Seg seg;
        // Per-pair form (TMA drain): the drain of a tile is postponed until the first K_AHEAD k-blocks of the NEXT tile
        // are expanded.  An expander is done with a tile ~STAGES k-blocks before the MMA warp is (the stages it filled are
        // still queued); instead of idling until the accumulator is complete and only then restarting the pipeline from
        // empty stages, it refills the stages as the last MMAs free them and drains afterwards: when the accumulator is
        // handed back, the next tile's MMAs find their operands waiting.  The MMA warp's order is unchanged (it waits
        // for `acc_empty` of the tile before), the stages it needs to finish that tile are all filled before any of this.
        constexpr uint32_t K_AHEAD = C::TMA_DRAIN ? (uint32_t)C::STAGES - 2u : 0u;
        bool drain_pending = false;
        uint64_t pending_tile = 0;
        auto drain_tile = [&](uint64_t tile) {                             // every tile of a per-pair job is a run of its own
            uint32_t bi = 0, bj = 0;
            cursor.coords(job, tile, C::TM, C::TN, bi, bj);
            wait(acc_full_bar, run_iter & 1);
            tc_fence_after();
            drain(0u, warp & 3u, warp >> 2, (uint32_t)(C::EXPANDER_WARPS / 4), bi, bj, false, false);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader<CG>(acc_empty_bar);          // the MMA thread may overwrite the accumulator
            ++run_iter;
        };
        for (; walk.next(seg); ++t_iter) {
            uint32_t flags = RUN_FIRST | RUN_LAST;
            uint32_t kb_done = 0;                                          // k-blocks of this tile expanded so far
            for (uint32_t c = seg.c0; c < seg.c1; ++c) {
                wait(raw_full_bar + 8 * buf, buf_phase);
                if (c == seg.c0) flags = ld_shared_u32(run_ring + 4 * (t_iter & (RUN_RING - 1)));   // written before this box was requested
                const uint32_t src = raw_base + buf * C::RAW_BYTES + raw_row;
                const uint32_t nq = min(CHUNK_KB, n_kb - c * CHUNK_KB);
                for (uint32_t q = 0; q < nq; ++q) {
                    // the packed bits of this row that this thread expands: the whole k-block (16 B in the i8 form,
                    // 32 B in the FP4 form) or, with two warps per row group that split it, the half that feeds its two K steps
                    constexpr int SPLIT = XW;
                    constexpr int NW = (FP4 ? 8 : 4) / SPLIT;              // 32-bit words per thread and k-block
                    constexpr int KS = 4 / SPLIT;                          // K steps (MMAs) per thread and k-block
                    static_assert(!(XW == 2 && !FP4), "two expander warps per row group: FP4 form only");
                    uint32_t ws[NW];
                    if constexpr (FP4 && SPLIT == 1) {
                        const uint4 w0 = ld_shared_v4(src + (((2 * q) ^ sw) << 4));
                        const uint4 w1 = ld_shared_v4(src + (((2 * q + 1) ^ sw) << 4));
                        ws[0] = w0.x; ws[1] = w0.y; ws[2] = w0.z; ws[3] = w0.w;
                        ws[4] = w1.x; ws[5] = w1.y; ws[6] = w1.z; ws[7] = w1.w;
                    } else {
                        const uint32_t chunk = FP4 ? 2 * q + half : q;
                        const uint4 w = ld_shared_v4(src + ((chunk ^ sw) << 4));
                        ws[0] = w.x; ws[1] = w.y; ws[2] = w.z; ws[3] = w.w;
                    }
                    wait(empty_bar + 8 * s, phase ^ 1);
                    uint32_t e[8];
                    const uint32_t k0 = half * KS;                         // first K step of this thread
                    if (is_a) {
                        tc_fence_after();
                        const uint32_t t = a_lane + UM_A_COL + s * 32;
#pragma unroll
                        for (int k = 0; k < KS; ++k) {                     // K step k0 + k = TMEM columns [8 (k0 + k), +8) of the stage
                            if constexpr (FP4) { expand32_a_fp4(ws[2 * k], e); expand32_a_fp4(ws[2 * k + 1], e + 4); }
                            else if (SCALED) expand32_a_scaled(ws[k], e);
                            else expand32(ws[k], e);
                            tmem_st8(t + 8 * (k0 + k), e);
                        }
                        tc_wait_st();
                        tc_fence_before();
                    } else {
                        const uint32_t dst = smem_base + s * C::STAGE_BYTES + b_line;
#pragma unroll
                        for (int k = 0; k < KS; ++k) {                     // K step k0 + k = bytes [32 (k0 + k), +32) of the line
                            if constexpr (FP4) { expand32_b_fp4(ws[2 * k], e); expand32_b_fp4(ws[2 * k + 1], e + 4); }
                            else if (SCALED) expand32_b_scaled(ws[k], e);
                            else expand32(ws[k], e);
                            st_shared_v4(dst + (((2 * (k0 + k)) ^ sw) << 4), e[0], e[1], e[2], e[3]);
                            st_shared_v4(dst + (((2 * (k0 + k) + 1) ^ sw) << 4), e[4], e[5], e[6], e[7]);
                        }
                        fence_proxy_async_smem();                          // generic writes -> visible to the UMMA (async proxy)
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader<CG>(full_bar + 8 * s);
                    if (++s == (uint32_t)C::STAGES) { s = 0; phase ^= 1; }
                    if constexpr (C::TMA_DRAIN) {
                        if (drain_pending && ++kb_done == K_AHEAD) { drain_tile(pending_tile); drain_pending = false; }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_local(raw_empty_bar + 8 * buf); // this warp is done with the box
                if (++buf == (uint32_t)C::RAW_BUFS) { buf = 0; buf_phase ^= 1; }
            }
            if constexpr (C::TMA_DRAIN) {
                if (job.out_tma) {
                    if (drain_pending) drain_tile(pending_tile);           // (a tile shorter than K_AHEAD k-blocks)
                    drain_pending = true;
                    pending_tile = seg.tile;
                    continue;
                }
            }
            if (flags & RUN_LAST) {
                // a run of several segments is interior by construction; a run of one may be a diagonal or edge tile
                uint32_t bi = 0, bj = 0;
                bool interior = true;
                if (flags & RUN_FIRST) { cursor.coords(job, seg.tile, C::TM, C::TN, bi, bj); interior = is_interior(bi, bj); }
                const uint64_t run_cap = job.chain_max > 1 ? job.chain_max : 1;    // an accumulator element is at most run_cap x M
                const bool fp4_sum_exact = (uint64_t)job.n_words * 64 * 32 * run_cap <= (1ull << 24);
                wait(acc_full_bar, run_iter & 1);
                tc_fence_after();
                drain(0u, warp & 3u, warp >> 2, (uint32_t)(C::EXPANDER_WARPS / 4), bi, bj, interior, fp4_sum_exact);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader<CG>(acc_empty_bar);      // the MMA thread may overwrite the accumulator
                ++run_iter;
            }
        }
        if constexpr (C::TMA_DRAIN) {
            if (drain_pending) drain_tile(pending_tile);                   // the last tile
        }
    }
    __syncwarp();                                                          // re-converge (aligned ops follow)

    // ---- one atomic per CTA, then teardown --------------------------------------------
    if (job.total) {
        sum = warp_sum(sum);
        if (lane == 0) red[warp] = sum;                                    // (zero for the roles that drain nothing)
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();                 // everyone is done with TMEM / smem
    if (job.total && tid == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < C::N_WARPS; ++w) t += red[w];
        if (t) atomicAdd(job.total, t);
    }
    if constexpr (C::TMA_DRAIN) {
        if (job.out_tma && lane == 0 && warp < (uint32_t)C::EXPANDER_WARPS) bulk_wait_all();   // (the stores were issued before the sync above)
    }
    if (clk_thread) {
        job.clk[2 * blockIdx.x] = (unsigned long long)(clock64() - clk0);
        job.clk[2 * blockIdx.x + 1] = global_timer_ns() - gt0;
    }
    if (warp == C::MMA_WARP) tmem_free<CG>(tmem_base);
}

// ---- tensor-pipe peak probe ----------------------------------------------------------
// Issues the production kernel's exact instruction (kind::i8, M = 128 * CG, N = 256, K = 32, A from
// tensor memory, B from SWIZZLE_128B shared memory) back to back with no operand production at
// all, so that its rate is the ceiling the tile kernel can reach on this device at its clocks.
// Operands are whatever the (zeroed) shared memory and tensor memory hold; results are discarded.
template <int CG>
__global__ void __launch_bounds__(128, 1) umma_peak_kernel(uint32_t iters, unsigned long long* cycles) {
    using C = Cfg<CG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr uint32_t RING = 8;                                             // commits in flight
    constexpr uint32_t BATCH = 8;                                            // MMAs per commit (two k-blocks of the tile kernel)
    const uint32_t bar_base = smem_base + C::STAGE_BYTES;                   // RING barriers, then the TMEM slot
    const uint32_t tmem_slot = bar_base + 8 * RING;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;

    for (uint32_t i = tid; i < C::STAGE_BYTES / 16; i += blockDim.x) st_shared_v4(smem_base + i * 16, 0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<CG>(tmem_slot);
    if (tid == 0) {
        for (uint32_t b = 0; b < RING; ++b) mbar_init(bar_base + 8 * b, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // Issue loop as in the tile kernel: the whole warp walks it (waits included), one elected lane issues, so that
    // descriptors and addresses stay in uniform registers.  (First version: one thread in a divergent branch, four
    // MMAs per commit, ring of four -- it reached 93.8 % of the pipe where the tile kernel itself reaches 99 %.)
    if (rank == 0 && warp == 0) {
        const bool leader = elect_one();
        const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const long long c0 = clock64();
        for (uint32_t it = 0; it < iters; ++it) {
            if (it >= RING) mbar_wait_t<true>(bar_base + 8 * (it % RING), ((it / RING) - 1) & 1);
            if (leader) {
#pragma unroll
                for (int k = 0; k < (int)BATCH; ++k) {
                    const uint64_t b_desc = desc_hi | (uint64_t)(((smem_base + (k & 3) * 32) >> 4) & 0x3FFF);
                    umma_i8_ts<CG>(tmem_u + UM_ACC_COL, tmem_u + UM_A_COL + (k & 3) * 8, b_desc, C::IDESC, 1u);
                }
                umma_commit<CG>(bar_base + 8 * (it % RING));
            }
            __syncwarp();
        }
        for (uint32_t it = iters > RING ? iters - RING : 0; it < iters; ++it)   // drain
            mbar_wait_t<true>(bar_base + 8 * (it % RING), (it / RING) & 1);
        const long long c1 = clock64();
        if (leader && cycles) cycles[blockIdx.x / CG] = (unsigned long long)(c1 - c0);
    }
    __syncwarp();
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) tmem_free<CG>(tmem_base);
}

template <int CG>
int run_umma_peak(double* ops_per_s, double* clock64_mhz) {
    using C = Cfg<CG>;
    int dev = 0, sms = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    STORM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int smem_bytes = 1024 + C::STAGE_BYTES + 256;
    STORM_CUDA_TRY(cudaFuncSetAttribute(umma_peak_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(sms / CG * CG));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    STORM_CUDA_TRY(cudaEventCreate(&e0));
    STORM_CUDA_TRY(cudaEventCreate(&e1));
    const unsigned n_clusters = cfg.gridDim.x / CG;
    unsigned long long* d_cyc = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&d_cyc, n_clusters * sizeof(unsigned long long)));
    std::vector<unsigned long long> h_cyc(n_clusters);
    const uint32_t iters = 50000;                                         // x 8 MMAs: ~50 M clocks, tens of milliseconds
    double best = 0, best_mhz = 0;
    for (int rep = 0; rep < 4; ++rep) {                                   // rep 0 is the warm-up
        STORM_CUDA_TRY(cudaEventRecord(e0));
        STORM_CUDA_TRY(cudaLaunchKernelEx(&cfg, umma_peak_kernel<CG>, iters, d_cyc));
        STORM_CUDA_TRY(cudaEventRecord(e1));
        STORM_CUDA_TRY(cudaEventSynchronize(e1));
        count_launch();
        float ms = 0;
        STORM_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        STORM_CUDA_TRY(cudaMemcpy(h_cyc.data(), d_cyc, n_clusters * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double cyc = 0;
        for (unsigned long long c : h_cyc) cyc += (double)c;
        cyc /= n_clusters;
        // per SM and instruction: 128 x 256 x 32 MACs = 2 ops each
        const double ops = (double)cfg.gridDim.x * iters * 8.0 * 128.0 * 256.0 * 32.0 * 2.0;
        if (rep > 0 && ops / (ms * 1e-3) > best) { best = ops / (ms * 1e-3); best_mhz = cyc / (ms * 1e-3) / 1e6; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_cyc);
    *ops_per_s = best;
    if (clock64_mhz) *clock64_mhz = best_mhz;                             // issue-loop clock64 ticks per second of the launch
    return STORM_B200_OK;
}

// ---- host side --------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// Packed rows as a 2-D byte tensor: inner = n_words * 8 bytes, outer = rows; box = 128 bytes x box_rows.
// Encoding is a pure function of the arguments and costs a driver call; a query on a resident matrix asks for the same
// map again and again (and for A and B of a triangle job twice per launch), so the last few are kept per host thread.
int make_row_map(CUtensorMap* map, const uint64_t* base, uint64_t n_rows, uint64_t stride_words, uint32_t n_words, uint32_t box_rows) {
    struct Entry { const uint64_t* base; uint64_t n_rows, stride; uint32_t n_words, box_rows; CUtensorMap map; };
    constexpr int SLOTS = 16;
    thread_local Entry cache[SLOTS];
    thread_local int used = 0, next = 0;
    for (int k = 0; k < used; ++k) {
        const Entry& e = cache[k];
        if (e.base == base && e.n_rows == n_rows && e.stride == stride_words && e.n_words == n_words && e.box_rows == box_rows) { *map = e.map; return STORM_B200_OK; }
    }
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return STORM_B200_ECUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)n_words * 8, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)stride_words * 8};
    const cuuint32_t box[2] = {128, box_rows};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint64_t*>(base), dims, strides, box, elem,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return STORM_B200_ECUDA; }
    Entry& slot = cache[next];
    slot = Entry{base, n_rows, stride_words, n_words, box_rows, *map};
    next = (next + 1) % SLOTS;
    if (used < SLOTS) ++used;
    return STORM_B200_OK;
}

// Per-pair counts as a 2-D uint32 tensor: inner = the nB valid columns, outer = the nA rows, pitch ld; box = 32 x 32
// (128 bytes x 32 rows, SWIZZLE_128B: what a warp of the epilogue stages in one 4 KiB slice).
int make_out_map(CUtensorMap* map, uint32_t* out, uint64_t nA, uint64_t nB, uint64_t ld) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return STORM_B200_ECUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)nB, (cuuint64_t)nA};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, out, dims, strides, box, elem,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)r); return STORM_B200_ECUDA; }
    return STORM_B200_OK;
}

// Wave counters: a small ring per device, one slot per launch, zeroed on the launch's stream.
int wave_counter(cudaStream_t stream, unsigned int** slot) {
    constexpr int MAX_DEV = 16, RING = 256;
    static unsigned int* pool[MAX_DEV] = {};
    static std::atomic<unsigned> next[MAX_DEV];
    static std::mutex mu;
    int dev = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= MAX_DEV) { *slot = nullptr; return STORM_B200_OK; }
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!pool[dev]) STORM_CUDA_TRY(cudaMalloc(&pool[dev], RING * sizeof(unsigned int)));
    }
    *slot = pool[dev] + (next[dev].fetch_add(1) % RING);
    STORM_CUDA_TRY(cudaMemsetAsync(*slot, 0, sizeof(unsigned int), stream));
    return STORM_B200_OK;
}

// Process-wide development / measurement knobs.  Atomics: a launch reads each of them once, so a caller on another
// thread that flips one never tears a launch (what it cannot have is a per-object value: see DenseJob::reserved_sms
// for the one knob a multi-process caller used to flip mid-query).
std::atomic<int> g_umma_wave_sync{1};      // STORM_b200_set_umma_wave_sync
std::atomic<int> g_umma_reserved_sms{0};   // STORM_b200_set_umma_reserved_sms
std::atomic<int> g_umma_stream_k{1};       // STORM_b200_set_umma_stream_k
std::atomic<int> g_umma_chain{1};          // STORM_b200_set_umma_chain
std::atomic<int> g_clock_probe{0};         // STORM_b200_set_clock_probe
std::atomic<int> g_umma_l2_hints{2};       // STORM_b200_set_umma_variant bit 5: L2 eviction hints on the packed-row loads of triangle jobs
std::atomic<int> g_umma_out_tma{1};        // STORM_b200_set_umma_variant bit 4: per-pair counts leave through TMA stores

// Clock-probe buffer of the current device (2 x u64 per CTA of the last probed launch) and the grid of that launch.
struct ClockProbe { unsigned long long* d = nullptr; unsigned grid = 0; };
ClockProbe g_clk[16];
std::mutex g_clk_mu;

template <int CG, int VAR>
int launch_cg(const DenseJob& job_in, cudaStream_t stream) {
    using C = Cfg<CG, (VAR & VAR_WIDE) ? 2 : 1, (VAR & VAR_FP4) != 0, (VAR & VAR_PAIRS) != 0>;
    DenseJob job = job_in;
    alignas(64) CUtensorMap map_a, map_b, map_out;
    int rc = make_row_map(&map_a, job.A, job.nA, job.strideA, job.n_words, 128);
    if (!rc) rc = make_row_map(&map_b, job.B, job.nB, job.strideB, job.n_words, C::B_ROWS);
    if (rc) return rc;
    map_out = map_a;                                                         // (a valid descriptor when none is needed)
    job.out_tma = 0;
    if (C::TMA_DRAIN && job.out && g_umma_out_tma.load() && ((reinterpret_cast<uintptr_t>(job.out) & 15) == 0) && (job.ld % 4 == 0) &&
        job.nA < (1ull << 32) && job.nB < (1ull << 32) && job.ld < (1ull << 38)) {
        if ((rc = make_out_map(&map_out, job.out, job.nA, job.nB, job.ld))) return rc;
        job.out_tma = 1;
    }
    int dev = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    // once per device and instantiation: the shared-memory opt-in of this kernel, the SM count
    static std::atomic<int> sms_of[64];
    int sms = dev < 64 ? sms_of[dev].load(std::memory_order_relaxed) : 0;
    if (sms == 0) {
        STORM_CUDA_TRY(cudaFuncSetAttribute(dense_umma_kernel<CG, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        STORM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (dev < 64) sms_of[dev].store(sms, std::memory_order_relaxed);
    }
    const uint64_t n_tiles = job.tile_end - job.tile_begin;
    const int want_reserved = job.reserved_sms >= 0 ? job.reserved_sms : g_umma_reserved_sms.load();
    const int reserved = want_reserved < sms - 2 ? want_reserved : sms - 2;
    const uint64_t max_clusters = (uint64_t)((sms - reserved) / CG);                               // persistent: one per SM (pair)
    uint64_t clusters = n_tiles < max_clusters ? n_tiles : max_clusters;
    job.wave_sync = nullptr;
    job.stream_k = 0;
    // (worth it where the rows do not fit in L2 anyway: the same condition as the wave barrier below)
    job.l2_hints = 0;
    // Worth it when the rows do not fit in L2 anyway and a tile lasts long enough to hide the barrier: below
    // that it costs up to 10 % (32768 x 4096: 0.79 vs 0.71 ms) and there is no DRAM traffic to save.
    const uint64_t matrix_bytes = (job.nA + (job.A == job.B ? 0 : job.nB)) * (uint64_t)job.n_words * 8;
    if (g_umma_l2_hints.load() && job.triangle && n_tiles > clusters && matrix_bytes >= (96ull << 20)) job.l2_hints = g_umma_l2_hints.load();
    if (g_umma_wave_sync.load() && n_tiles > clusters && matrix_bytes >= (96ull << 20) && job.n_words >= 512) {
        int rc2 = wave_counter(stream, &job.wave_sync);
        if (rc2) return rc2;
    }
    // Stream-K for total-only jobs (a total is a sum, so a tile's K range may be split between CTAs): the tail
    // wave is cut along K (10000 x 65536: 820 tiles on 74 clusters = 11.08 waves, run as 12 without it), and
    // with fewer tiles than clusters the idle SMs get K slices of the tiles (at least STREAMK_MIN_CHUNKS TMA
    // boxes each, so that a segment's epilogue stays small beside its MMAs).
    constexpr uint64_t STREAMK_MIN_CHUNKS = 8;
    if (g_umma_stream_k.load() && !job.out) {
        if (n_tiles >= max_clusters) job.stream_k = (n_tiles % max_clusters) != 0;   // (the wave barrier covers the full waves only)
        else if (!job.wave_sync) {
            const uint32_t n_kb = C::FP4_FORM ? (job.n_words + 3) / 4 : (job.n_words + 1) / 2;
            const uint32_t chunk_kb = C::FP4_FORM ? UM_CHUNK_KB_FP4 : UM_CHUNK_KB;
            const uint64_t units = n_tiles * ((n_kb + chunk_kb - 1) / chunk_kb);
            uint64_t want = units / STREAMK_MIN_CHUNKS;
            if (want > max_clusters) want = max_clusters;
            if (want > n_tiles) { clusters = want; job.stream_k = 1; }
        }
    }
    // Accumulator chaining for total-only jobs: consecutive interior segments of a CTA share the accumulator
    // and are drained once per run.  An element then holds at most chain_max x M, which must stay inside the
    // exact range of the form: fp32 integers below 2^24 for kind::mxf4 (and 32 of them must still add up
    // exactly in the epilogue's float sum), s32 for kind::i8 (x 128 in the scaled form).
    job.chain_max = 1;
    if (g_umma_chain.load() && !job.out && job.total) {
        const uint64_t M = (uint64_t)job.n_words * 64;
        const uint64_t room = C::FP4_FORM ? (1ull << 24) / (M * 32)
                            : (VAR & VAR_SCALED) ? 0x7FFFFFFFull / (M * 128) : 0x7FFFFFFFull / M;
        job.chain_max = (uint32_t)(room < 1 ? 1 : room > 4096 ? 4096 : room);
    }
    job.clk = nullptr;
    if (g_clock_probe.load() && dev < 16) {
        std::lock_guard<std::mutex> lock(g_clk_mu);
        if (!g_clk[dev].d) STORM_CUDA_TRY(cudaMalloc(&g_clk[dev].d, 2 * 512 * sizeof(unsigned long long)));
        if (clusters * CG <= 512) { job.clk = g_clk[dev].d; g_clk[dev].grid = (unsigned)(clusters * CG); }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * CG));
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    STORM_CUDA_TRY(cudaLaunchKernelEx(&cfg, dense_umma_kernel<CG, VAR>, map_a, map_b, map_out, job));
    count_launch();
    return STORM_B200_OK;
}

std::atomic<int> g_umma_cg{2};        // cta_group used by launch_dense_umma (1 or 2); see STORM_b200_set_umma_cta_group
// FP4 form, cta_group 2: two expander warps per 32 rows (STORM_b200_set_umma_variant bit 3).  Off by default:
// once the MMA issue loop was made warp-uniform the narrow form reached 95.8 % of the pipe on C3 and the
// wide one 92.5 % (it only wins by a few percent below 16 Ki bits per row), profiles/r01_fp4_tune.jsonl.
std::atomic<int> g_umma_fp4_wide{0};
std::atomic<int> g_umma_variant{3};   // VAR_* bits (both on: 4.27 vs 3.60 POP/s on 30k x 131072); see STORM_b200_set_umma_variant

template <int CG>
int launch_var(const DenseJob& job, cudaStream_t stream, bool fp4) {
    int var = g_umma_variant.load() & 3;
    const bool pairs = job.out != nullptr;                                  // per-pair output: the two-accumulator form
    if (fp4) {
        if (!umma_fp4_supports(job)) {
            set_error("FP4 kernel: a pair count must stay below 2^24 for exact fp32 accumulation (n_words %u)", job.n_words);
            return STORM_B200_EINVAL;
        }
        if (pairs) return launch_cg<CG, VAR_FP4 | VAR_SUSPEND | VAR_PAIRS>(job, stream);
        if constexpr (CG == 2) {
            if (g_umma_fp4_wide.load()) return launch_cg<CG, VAR_FP4 | VAR_SUSPEND | VAR_WIDE>(job, stream);
        }
        return launch_cg<CG, VAR_FP4 | VAR_SUSPEND>(job, stream);
    }
    if ((uint64_t)job.n_words * 64 * 128 >= (1ull << 31)) var &= ~VAR_SCALED;   // x128 counts must stay below 2^31
    if (pairs) return (var & VAR_SCALED) ? launch_cg<CG, VAR_SUSPEND | VAR_SCALED | VAR_PAIRS>(job, stream)
                                         : launch_cg<CG, VAR_SUSPEND | VAR_PAIRS>(job, stream);
    switch (var) {
        case 0: return launch_cg<CG, 0>(job, stream);
        case 1: return launch_cg<CG, 1>(job, stream);
        case 2: return launch_cg<CG, 2>(job, stream);
        default: return launch_cg<CG, 3>(job, stream);
    }
}

}  // namespace

TileShape umma_tile_shape() { return {(uint32_t)(128 * g_umma_cg.load()), (uint32_t)UM_N}; }
// Tiles of a per-pair job (DenseJob::out set): the same as the total-only form's.
TileShape umma_pairs_tile_shape() { return umma_tile_shape(); }

bool umma_supports(const DenseJob& job) {
    if (job.n_words == 0 || job.n_words >= (1u << 25)) return false;        // counts stay below 2^31
    if ((job.strideA & 1) || (job.strideB & 1)) return false;               // TMA: 16-byte row pitch and base
    if (job.A && ((uintptr_t)job.A & 15)) return false;
    if (job.B && ((uintptr_t)job.B & 15)) return false;
    if (job.nA >= (1ull << 31) || job.nB >= (1ull << 31)) return false;
    return true;
}

// The FP4 form accumulates in fp32: exact while a pair count (at most 64 * n_words) stays below 2^24.
bool umma_fp4_supports(const DenseJob& job) {
    return umma_supports(job) && (uint64_t)job.n_words * 64 <= (1ull << 24);
}

int umma_peak_ops(int cg, double* ops_per_s, double* clock64_mhz) {
    return cg == 1 ? run_umma_peak<1>(ops_per_s, clock64_mhz) : run_umma_peak<2>(ops_per_s, clock64_mhz);
}

int launch_dense_umma(const DenseJob& job, cudaStream_t stream) {
    if (job.tile_end <= job.tile_begin) return STORM_B200_OK;
    return g_umma_cg.load() == 2 ? launch_var<2>(job, stream, false) : launch_var<1>(job, stream, false);
}

int launch_dense_fp4(const DenseJob& job, cudaStream_t stream) {
    if (job.tile_end <= job.tile_begin) return STORM_B200_OK;
    return g_umma_cg.load() == 2 ? launch_var<2>(job, stream, true) : launch_var<1>(job, stream, true);
}

}  // namespace storm

// Development / measurement knob: cta_group of the UMMA kernel (1 = one CTA per 128 x 256 tile,
// 2 = CTA pair per 256 x 256 tile).  Returns the previous value.
extern "C" int STORM_b200_set_umma_cta_group(int cg) {
    const int prev = storm::g_umma_cg.load();
    if (cg == 1 || cg == 2) storm::g_umma_cg.store(cg);
    return prev;
}

// Development / measurement knob: 1 (default) = the CTAs of the persistent UMMA kernel keep their tile
// waves in step (L2 reuse of the shared row blocks), 0 = free-running.  Returns the previous value.
extern "C" int STORM_b200_set_umma_wave_sync(int on) {
    return storm::g_umma_wave_sync.exchange(on ? 1 : 0);
}

// SMs the persistent kernel leaves to a collective running beside it (multi-GPU host queries).  Returns the previous value.
extern "C" int STORM_b200_set_umma_reserved_sms(int n) {
    return storm::g_umma_reserved_sms.exchange(n < 0 ? 0 : n);
}

// Development / measurement knob: 1 (default) = total-only jobs with few tiles per CTA split (tile, K chunk)
// units evenly over the persistent CTAs, 0 = whole tiles only.  Returns the previous value.
extern "C" int STORM_b200_set_umma_stream_k(int on) {
    return storm::g_umma_stream_k.exchange(on ? 1 : 0);
}

// Development / measurement knob: 1 (default) = total-only jobs drain the accumulator once per run of interior
// segments (DenseJob::chain_max), 0 = once per segment.  Returns the previous value.
extern "C" int STORM_b200_set_umma_chain(int on) {
    return storm::g_umma_chain.exchange(on ? 1 : 0);
}

// Development / measurement knob: bit 0 = hardware-suspended mbarrier waits, bit 1 = scaled expansion.
// Returns the previous value.
extern "C" int STORM_b200_set_umma_variant(int variant) {
    const int hints = storm::g_umma_l2_hints.load();
    const int prev = storm::g_umma_variant.load() | (storm::g_umma_fp4_wide.load() ? 8 : 0) | (storm::g_umma_out_tma.load() ? 16 : 0) |
                     (hints ? 32 : 0);
    if (variant >= 0 && variant <= 63) {
        storm::g_umma_variant.store(variant & 3); storm::g_umma_fp4_wide.store((variant >> 3) & 1); storm::g_umma_out_tma.store((variant >> 4) & 1);
        // bit 5 switches the hints on or off; which of the two forms is STORM_b200_set_umma_l2_hints' business (a caller
        // that restores a previous value must not turn mode 2 into mode 1)
        storm::g_umma_l2_hints.store(((variant >> 5) & 1) ? (hints ? hints : 2) : 0);
    }
    return prev;
}

// L2 eviction hints on the packed-row loads of triangle jobs that do not fit in L2: 0 none, 1 column blocks evict_last,
// 2 column blocks evict_last + row blocks evict_first.  Results are identical.  Returns the previous value.
extern "C" int STORM_b200_set_umma_l2_hints(int mode) { return storm::g_umma_l2_hints.exchange(mode < 0 ? 0 : mode > 2 ? 2 : mode); }

// 1: every tensor-kernel launch records, per CTA, the clock64 and %globaltimer deltas around its main loop (one thread,
// a handful of instructions); 0 (default): off.  Returns the previous value.
extern "C" int STORM_b200_set_clock_probe(int on) { return storm::g_clock_probe.exchange(on ? 1 : 0); }

// SM clock of the most recent probed tensor-kernel launch on the current device, in clock64 ticks per microsecond
// (= MHz if clock64 ticks once per SM cycle; STORM_b200_microbench(8) calibrates that): waits for the device to go
// idle, then averages the per-CTA records.  *min_mhz / *max_mhz (optional): the spread over CTAs.
extern "C" int STORM_b200_last_kernel_clock(double* mhz, double* min_mhz, double* max_mhz) {
    using namespace storm;
    if (!mhz) { set_error("mhz is NULL"); return STORM_B200_EINVAL; }
    int dev = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16 || !g_clk[dev].d || g_clk[dev].grid == 0) { set_error("no probed launch on device %d (STORM_b200_set_clock_probe)", dev); return STORM_B200_EINVAL; }
    STORM_CUDA_TRY(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(2 * g_clk[dev].grid);
    STORM_CUDA_TRY(cudaMemcpy(h.data(), g_clk[dev].d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double c = 0, t = 0, lo = 1e30, hi = 0;
    for (unsigned i = 0; i < g_clk[dev].grid; ++i) {
        const double ci = (double)h[2 * i], ti = (double)h[2 * i + 1];
        if (ti <= 0) continue;
        c += ci; t += ti;
        lo = std::min(lo, ci / ti * 1e3); hi = std::max(hi, ci / ti * 1e3);
    }
    if (t <= 0) { set_error("clock probe holds no record"); return STORM_B200_EINVAL; }
    *mhz = c / t * 1e3;
    if (min_mhz) *min_mhz = lo;
    if (max_mhz) *max_mhz = hi;
    return STORM_B200_OK;
}
