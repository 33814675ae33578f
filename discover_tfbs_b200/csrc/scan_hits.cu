// scan_hits.cu — (2b) thresholded both-strand PWM scan ("hits mode"), the headline kernel.
//
// Replaces Biopython PositionSpecificScoringMatrix.search (Bio/motifs/matrix.py, _pwm.c;
// SURVEY.md A.1) for a whole motif library in one pass, and generalises the reference's
// exact-match BLASTn candidate stage (utils.py:340-398, process_blast.sh:32).
//
// Design (one warp lane = one motif — two in narrow blocks — so shared-memory LUT reads are
// bank-conflict free):
//   * motifs are sorted by length and grouped 32 to a (wide) block.  The columns of a block are cut into
//     A groups of 4 columns followed by B groups of 3 (A, B chosen per block to minimise A + B under
//     the shared-memory budget: a 4-mer group costs 32 KB, a 3-mer group 8 KB).  The block table
//     lives in shared memory as T[g][x][lane] (x = 8- or 6-bit k-mer code, 4 B per entry): the 32
//     lanes of a warp read 32 consecutive words for the same (g, x).
//   * an entry packs two 16-bit fixed-point DEFICITS (how far the k-mer is below the best k-mer
//     of that column group, in units chosen from the threshold), forward strand in the high
//     half, reverse-complement strand in the low half, so one LDS.32 + one IADD scores both
//     strands.  Deficits are rounded DOWN and capped, fields are sized so sums never carry, and
//     a window is a candidate iff bit 15 of its field is still clear — a proven superset of
//     {score >= threshold}.
//   * the warp walks its segment base by base (the k-mer code is warp-uniform, sliced from the
//     2-bit packed words staged in shared memory by TMA); A + B loads feed A + B of the 32 ring
//     accumulators held in registers; finalised accumulators are tested up to 8 at a time (their AND
//     keeps a flag bit set unless one of them cleared it), with a warp vote.
//   * candidates (rare) go to a per-warp shared-memory queue; a full queue is appended to a global
//     candidate list (one atomic per batch, coalesced 8-byte records).  A second kernel
//     (verify_candidates_kernel, full occupancy, exact matrices served by L1/L2) re-scores every
//     candidate with the oracle's arithmetic — invalid-base mask check, left-to-right double sum cast
//     to float, compared with >= — so hit sets and scores are bit-exact; hits are compacted with
//     ballot/popc and one global atomic per CTA batch.  Verifying inside the scan kernel (round 1)
//     stalled its 16 warps per SM on L2 latency: ~0.2 ns per candidate, 13 % of the scan at a 1e-4 hit
//     rate and most of it at 1e-3.  The in-kernel verifier is kept as the overflow path (candidate
//     list or queue full: flood thresholds), so no threshold vector can lose a hit.
//   * NARROW blocks: 64 motifs per block, two per lane, four 8-bit fields per entry (motif lane: fwd,
//     rev; motif lane + 32: fwd, rev).  The same LDS.32 + IADD now scores two motifs, halving
//     shared-memory traffic per motif — the resource this kernel is bound by.  Only
//     (128 - 2G) / (G - 1) quantisation steps fit without carries, so the filter is coarser (1.1x the
//     candidates at G = 3, 2.4x at G = 5, 78x at G = 9; profiles/field8_simulation.py) and the exact
//     verifier, which keeps the hit set bit-exact, has more to do.
//   * STICKY narrow blocks (5 or more groups): flag bits, once set, are OR-ed back after every add
//     (acc = (acc + v) | (acc & 0x80808080), one more ALU op per load), so a dead field may wrap
//     without looking alive again.  A dead field counts modulo 128 below its restored flag and sends
//     +1 into the field above it on every wrap, at most (2 + 127 G + c_in) / 128 - 1 <= 8 times per
//     window; a field with a neighbour below therefore starts that many steps lower, which keeps
//     {candidates} a superset of {hits}, and >= 117 quantisation steps are usable whatever G is: the
//     8-bit filter is then almost as sharp as the 16-bit one.
//     Wide, narrow or sticky is chosen per block and per threshold vector by a cost model
//     (tfbs_build_hits_tables).
//   * persistent CTAs (1 per SM, 16 warps) pull (motif block, tile) items from an atomic counter in
//     block-major order, so a CTA re-loads its table (up to 192 KB) only when the block changes.
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cooperative_groups.h>
#include <thread>
#include <cub/device/device_radix_sort.cuh>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kRing = 32;                 // ring accumulators: >= 3*(G-1)+1 for G <= 11, multiple of 16
constexpr int kSeg = 2016;                // window starts per warp per tile (63 ring turns of 32)
constexpr int kTilePos = kWarps * kSeg;   // 32256 starts per tile (multiple of 128)
constexpr int kTileBases = kTilePos + 128;
constexpr int kTileWords = kTileBases / 16;  // 2024 words  = 8096 B (multiple of 16)
constexpr int kTileMask = kTileBases / 32;   // 1012 words  = 4048 B (multiple of 16)
constexpr int kQueueCap = 128;            // candidate queue entries per warp
constexpr int kQueueWarpsA = 13;          // warps whose queue sits before the table (the others: behind it)
constexpr int kTileBufs = 2;              // tile ring: warps run up to one tile ahead of the slowest warp
constexpr int kMaxG = (TFBS_MAX_MOTIF_LEN + 2) / 3;  // 11
constexpr uint32_t kPoison = 0x80008000u;        // flag bits of the two 16-bit fields (wide blocks)
constexpr uint32_t kPoisonNarrow = 0x80808080u;  // flag bits of the four 8-bit fields (narrow blocks)
constexpr int kLutBudget = 192 * 1024;    // shared memory for one block table (6 x 32 KB)
// before the table: tile ring, motif/len of the block, mbarriers + ring bookkeeping; after it: the queues
constexpr int kBookBytes = kTileBufs * (kTileWords * 4 + kTileMask * 4) + 2 * TFBS_BLOCK_SLOTS * 4 + 64;
constexpr int kSmallBytes = kBookBytes + kQueueWarpsA * kQueueCap * 4;   // 31 520 B: fits below the first 32 KB boundary
constexpr int kQueueBytes = (kWarps - kQueueWarpsA) * kQueueCap * 4;       // behind the table

struct HitsParams {
    const uint32_t *packed;
    const uint32_t *mask;
    int64_t n_tiles;   // tiles scanned by this launch
    int64_t tile0;     // first tile (chunk-pipelined launches scan a sub-range)
    int32_t n_blocks;
    const uint32_t *qlut;        // per block [A][256][32] then [B][64][32] packed deficits
    const int32_t *blk_lut_off;  // [n_blocks] offset into qlut (uint32 units)
    const int32_t *blk_AB;       // [n_blocks] (mode << 16) | (A << 8) | B: 4- and 3-column groups of the block
    uint32_t lut_bytes;          // largest block table of the library: the warp queues sit right behind it
    const uint32_t *kinit;       // [n_blocks][32] initial field values per lane (0x7FFF - Dq or 0x7F - Dq each)
    const int32_t *blk_motif;    // [n_blocks][64] original motif index or -1 (slot = lane + 32 * half)
    const int32_t *lens;         // [M]
    const float *thr;            // [M]
    const double *exact;         // [M][2][32][4]
    const double *exact2;        // [M][2][16][16] column-pair sums (verify pass, fast path)
    const double *guard_eps;     // [M] rounding guard of the fast path
    tfbs_hit *hits;
    unsigned long long *hit_count;   // [0] hits, [1] candidates handed to a verifier, [2] error flag, [4] list length
    unsigned long long cap;
    unsigned int *work_counter;
    unsigned long long *cand;        // global candidate list: (pos << 24) | (motif << 1) | strand
    unsigned long long cand_cap;
    int64_t owned_end;               // window starts >= owned_end belong to the next shard (halo): not reported
    int64_t pos_offset;              // added to every reported position (shard start in genome coordinates)
};
constexpr int kCandCountSlot = 4;                      // index into hit_count of the candidate-list length
constexpr unsigned long long kCandEmpty = ~0ull;       // list entry to skip
constexpr int kCandMotifBits = 23;

struct Smem {
    // tile ring and bookkeeping first, then the block table on the next 32 KB boundary of the shared
    // window, then the warps' candidate queues
    uint32_t *lut;       // [A][256][32] then [B][64][32]
    uint32_t *words;     // [kTileBufs][kTileWords]
    uint32_t *mask;      // [kTileBufs][kTileMask]
    uint32_t *queue;     // [kWarps][kQueueCap]
    int32_t *motif;      // [64]
    int32_t *len;        // [64]
    uint64_t *full;      // [kTileBufs] mbarriers: tile landed (and its item slot is valid)
    uint64_t *tab;       // mbarrier: block table landed
    uint32_t *item;      // [kTileBufs] work item staged in each buffer
    uint32_t *done;      // [kTileBufs] warps that have finished with the buffer's current tile
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    uint32_t spins = 0;
    do {
        if (++spins == (1u << 28)) __trap();  // a lost arrival must fail the launch (after tens of seconds), not hang the GPU
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

struct VerifyCtx {
    const uint32_t *s_words;
    const uint32_t *s_mask;
    const int32_t *s_motif;
    const int32_t *s_len;
    const float *thr;
    const double *exact;
    int64_t tile_start;
    int64_t owned_end, pos_offset;
};

// Oracle arithmetic for one candidate (tile-local start p, motif slot ml, strand).
__device__ __forceinline__ bool verify(const VerifyCtx &V, int p, int ml, int strand, int *motif_out, float *score_out)
{
    const int m = V.s_motif[ml];
    if (m < 0 || V.tile_start + p >= V.owned_end) return false;
    const int L = V.s_len[ml];
    const uint32_t lmask = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    const uint32_t inv = __funnelshift_r(V.s_mask[p >> 5], V.s_mask[(p >> 5) + 1], p & 31) & lmask;
    if (inv) return false;  // window touches an invalid base: NaN in the oracle, never a hit
    const double *mat = V.exact + ((size_t)(m * 2 + strand) * TFBS_MAX_MOTIF_LEN) * 4;
    // left-to-right double sum (the oracle's order); the table loads of 8 columns are issued
    // together so their L2 latency overlaps — the adds stay sequential
    double s = 0.0;
    for (int j0 = 0; j0 < L; j0 += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int j = j0 + k, i = p + j;
            const uint32_t code = (V.s_words[i >> 4] >> (2 * (i & 15))) & 3u;
            v[k] = j < L ? __ldg(mat + j * 4 + code) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (j0 + k < L) s += v[k];
    }
    const float f = (float)s;
    *motif_out = m;
    *score_out = f;
    return f >= __ldg(V.thr + m);
}

__device__ __forceinline__ void store_hit(tfbs_hit *hits, unsigned long long slot, unsigned long long cap,
                                          int64_t gpos, int strand, int m, float f)
{
    if (slot < cap) {
        tfbs_hit h;
        h.pos = strand ? ~gpos : gpos;
        h.motif = m;
        h.score = f;
        hits[slot] = h;
    }
}

struct PushCtx {
    uint32_t *queue;   // this warp's queue
    int seg_begin, seg_end;
    int lane;
    const VerifyCtx *V;
    tfbs_hit *hits;
    unsigned long long *hit_count;
    unsigned long long cap;
    unsigned long long *cand;
    unsigned long long cand_cap;
};

// Warp-synchronous hand-over of the queued candidates to the global candidate list (one atomic per
// call, coalesced 8-byte stores).  Entries that do not fit the list are verified here, 32 at a time
// (hits compacted with ballot/popc and one atomic per batch): the overflow path of flood thresholds.
// `qn` = entries pushed since the last drain (warp-uniform register; entries beyond kQueueCap were
// verified inline by push_group).
__device__ __noinline__ void drain_queue(const PushCtx &C, uint32_t qn)
{
    __syncwarp();
    const uint32_t cnt = min(qn, (uint32_t)kQueueCap);
    if (cnt == 0) return;
    unsigned long long list0 = 0;
    if (C.lane == 0) {
        list0 = atomicAdd(C.hit_count + kCandCountSlot, (unsigned long long)cnt);
        atomicAdd(C.hit_count + 1, (unsigned long long)cnt);  // statistics: candidates handed to a verifier
    }
    list0 = __shfl_sync(0xFFFFFFFFu, list0, 0);
    const bool overflow = list0 + cnt > C.cand_cap;  // warp-uniform
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t i = base + C.lane;
        bool hit = false;
        int m = 0, p = 0, strand = 0;
        float f = 0.f;
        if (i < cnt) {
            const uint32_t e = C.queue[i];
            p = (int)(e & 0xFFFFu);
            strand = (int)(e >> 31);
            const int ml = (int)((e >> 16) & 63u);
            const unsigned long long slot = list0 + i;
            if (slot < C.cand_cap) {
                const int mo = C.V->s_motif[ml];
                C.cand[slot] = mo < 0 ? kCandEmpty
                                      : (((unsigned long long)(C.V->tile_start + p) << (kCandMotifBits + 1)) |
                                         ((unsigned long long)mo << 1) | (unsigned long long)strand);
            } else {
                hit = verify(*C.V, p, ml, strand, &m, &f);
            }
        }
        if (overflow) {
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
            if (ballot) {
                unsigned long long slot0 = 0;
                if (C.lane == 0) slot0 = atomicAdd(C.hit_count, (unsigned long long)__popc(ballot));
                slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
                if (hit)
                    store_hit(C.hits, slot0 + __popc(ballot & ((1u << C.lane) - 1u)), C.cap, C.V->tile_start + p + C.V->pos_offset, strand, m, f);
            }
        }
    }
    __syncwarp();
}

// Queue full (flood thresholds): verify one candidate right here, divergently.  Out of line on
// purpose: the call sits in every candidate-test site of the unrolled scan loop, and inlining the
// verifier there made each instantiation several times larger than its hot loop.
__device__ __noinline__ void verify_now(const PushCtx &C, int p, int ml, int strand)
{
    int m;
    float sc;
    if (verify(*C.V, p, ml, strand, &m, &sc)) {
        cg::coalesced_group g = cg::coalesced_threads();
        unsigned long long hb = 0;
        if (g.thread_rank() == 0) hb = atomicAdd(C.hit_count, (unsigned long long)g.size());
        hb = g.shfl(hb, 0);
        store_hit(C.hits, hb + g.thread_rank(), C.cap, C.V->tile_start + p + C.V->pos_offset, strand, m, sc);
    }
}

// Cold path, entered by the WHOLE warp when any lane of a check group has a candidate field (its flag
// bit clear).  NF = fields per accumulator word: 2 (wide: bit 31 forward, bit 15 reverse) or 4 (narrow:
// bits 31 / 23 = motif slot `lane` forward / reverse, bits 15 / 7 = motif slot `lane + 32`).
// a[0..7] are the finalised accumulators of window starts p_first .. p_first+7 (unused slots: all flag
// bits set).  One warp-wide OR tells which (start, field) pairs have candidates at all (usually one); each of those is compacted into the warp's queue with
// ballot/popc — the queue is warp-private and every access is warp-synchronous, so no atomics and
// no lane-by-lane serialisation when hits are dense.
template <int NF>
__device__ __forceinline__ void push_group(const PushCtx &C, const uint32_t (&a)[8], int p_first, uint32_t &qn)
{
    constexpr uint32_t PM = NF == 4 ? kPoisonNarrow : kPoison;
    // bit 8b + k of `mine`: this lane has a candidate at start p_first + k in the field whose flag is
    // bit 8b + 7 of the accumulator (b = 3: slot `lane` forward; narrow b = 2 / 1 / 0: slot `lane`
    // reverse, slot `lane + 32` forward / reverse; wide b = 1: slot `lane` reverse)
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) mine |= (~a[k] & PM) >> (7 - k);  // unused slots hold PM: no bits
    {
        // starts outside [seg_begin, seg_end): the ring lead-in (never a candidate, masked for good
        // measure) and the overshoot of the last turn (windows of the next warp's segment)
        const int lo = C.seg_begin - p_first, hi = C.seg_end - p_first;
        uint32_t kmask = 0xFFu;
        if (lo > 0) kmask &= 0xFFu << min(lo, 8);
        if (hi < 8) kmask &= (1u << max(hi, 0)) - 1u;
        mine &= (kmask & 0xFFu) * 0x01010101u;
    }
    uint32_t any = __reduce_or_sync(0xFFFFFFFFu, mine);
    const uint32_t lt = (1u << C.lane) - 1u;
    while (any) {
        const int bit = __ffs(any) - 1;
        any &= any - 1u;
        const bool c = (mine >> bit) & 1u;
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, c);
        const uint32_t base = qn;
        qn += (uint32_t)__popc(b);
        if (c) {
            const int p = p_first + (bit & 7), byte = bit >> 3;
            const int strand = NF == 4 ? (~byte & 1) : (byte == 1);
            const int ml = C.lane + (NF == 4 && byte < 2 ? 32 : 0);  // motif slot of the block
            const uint32_t slot = base + (uint32_t)__popc(b & lt);
            if (slot < (uint32_t)kQueueCap) {
                C.queue[slot] = (uint32_t)p | ((uint32_t)ml << 16) | ((uint32_t)strand << 31);
            } else {
                verify_now(C, p, ml, strand);  // queue full (flood thresholds)
            }
        }
        __syncwarp();
    }
}

// One warp scans window starts [seg_begin, seg_end) of the staged tile for the 32 (wide) or 64
// (NARROW) motifs of the resident block.  A, B = 4-column and 3-column groups of the block (compile
// time, so the 32-entry accumulator ring and all LUT offsets are static).  Group g < A covers columns 4g..4g+3, group
// A + j covers columns 4A+3j..4A+3j+2.
template <int A, int B, int MODE>
__device__ __forceinline__ void scan_segment(const PushCtx &C, const uint32_t *s_words, uint32_t lut_lane_addr,
                                             uint32_t kinit)
{
    constexpr bool NARROW = MODE != 0, STICKY = MODE == 2;
    constexpr uint32_t PM = NARROW ? kPoisonNarrow : kPoison;
    uint32_t acc[kRing];
#pragma unroll
    for (int i = 0; i < kRing; i++) acc[i] = PM;
    uint32_t qn = 0;  // candidates queued by this warp since its last drain (warp-uniform)
    const uint32_t *wptr = s_words + (C.seg_begin >> 4);
    uint32_t wlo = wptr[0], whi = wptr[1];
    int wk = 2;
    constexpr int kTail = B > 0 ? 4 * A + 3 * (B - 1) : 4 * (A - 1);  // column offset of the last group
    constexpr int kChunks = (kSeg + kTail + kRing - 1) / kRing;
    constexpr int kCheck = (kRing - kTail >= 8) ? 8 : ((kRing - kTail >= 4) ? 4 : ((kRing - kTail >= 2) ? 2 : 1));
    const uint32_t lut3_lane_addr = lut_lane_addr + A * 32768;
    for (int c = 0; c < kChunks; c++) {
        const int q0 = C.seg_begin + c * kRing;
#pragma unroll
        for (int j = 0; j < kRing; j++) {
            const int jj = j & 15;
            // bits [7,15) of t7 = 4-mer code at base q0 + j, bits [7,13) = its 3-mer prefix
            const uint32_t t7 = (2 * jj >= 7) ? __funnelshift_r(wlo, whi, 2 * jj - 7) : (wlo << (7 - 2 * jj));
            if (A > 0) {
                const uint32_t addr4 = (t7 & 0x7F80u) | lut_lane_addr;
#pragma unroll
                for (int g = 0; g < A; g++) {
                    const uint32_t v = lds32(addr4 + g * 32768);
                    const int slot = (j - 4 * g + 2 * kRing) % kRing;
                    if (g == 0) acc[slot] = kinit + v;
                    else if (STICKY) acc[slot] = (acc[slot] + v) | (acc[slot] & PM);
                    else acc[slot] += v;
                }
            }
            if (B > 0) {
                const uint32_t addr3 = (t7 & 0x1F80u) | lut3_lane_addr;
#pragma unroll
                for (int g = 0; g < B; g++) {
                    const uint32_t v = lds32(addr3 + g * 8192);
                    const int slot = (j - 4 * A - 3 * g + 2 * kRing) % kRing;
                    if (A == 0 && g == 0) acc[slot] = kinit + v;
                    else if (STICKY) acc[slot] = (acc[slot] + v) | (acc[slot] & PM);
                    else acc[slot] += v;
                }
            }
            // Candidate test, amortised over kCheck (up to 8) consecutive steps: the flag bit of a field stays SET
            // in the AND of the finalised accumulators unless one of them cleared it.  (A finalised
            // slot is re-opened 32 - kTail steps later, which bounds kCheck.)  The vote keeps the
            // branch warp-uniform, so the cold path can compact candidates cooperatively.
            if ((j % kCheck) == kCheck - 1) {
                uint32_t fin[8] = {PM, PM, PM, PM, PM, PM, PM, PM};
#pragma unroll
                for (int k = 0; k < kCheck; k++) fin[k] = acc[(j - (kCheck - 1 - k) - kTail + 2 * kRing) % kRing];
                const uint32_t all = (fin[0] & fin[1] & fin[2] & fin[3]) & (fin[4] & fin[5] & fin[6] & fin[7]);
                if (__any_sync(0xFFFFFFFFu, (~all & PM) != 0u))
                    push_group<NARROW ? 4 : 2>(C, fin, q0 + j - (kCheck - 1) - kTail, qn);
            }
            if (jj == 15) {
                wlo = whi;
                whi = wptr[wk++];
            }
        }
        // all lanes are converged again here: verify queued candidates before the queue can fill up
        if (qn >= (uint32_t)kQueueCap / 2) {
            drain_queue(C, qn);
            qn = 0;
        }
    }
    drain_queue(C, qn);
}

// Offsets of the bookkeeping words inside the dynamic shared window (the tile ring comes first).
constexpr int kOffMask = kTileBufs * kTileWords * 4;
constexpr int kOffMotif = kOffMask + kTileBufs * kTileMask * 4;
constexpr int kOffLen = kOffMotif + TFBS_BLOCK_SLOTS * 4;
constexpr int kOffFull = kOffLen + TFBS_BLOCK_SLOTS * 4;
constexpr int kOffTab = kOffFull + kTileBufs * 8;
constexpr int kOffItem = kOffTab + 8;
constexpr int kOffDone = kOffItem + kTileBufs * 4;
static_assert(kOffDone + kTileBufs * 4 <= kBookBytes, "bookkeeping does not fit kBookBytes");
static_assert(kSmallBytes <= 32768 - 1024, "small buffers must end below the first 32 KB boundary of the shared window");
static_assert(kOffFull % 8 == 0, "mbarriers must be 8-byte aligned");

// Stage the tile of work item `item` in ring buffer `buf` and publish the item (called by ONE thread).
// An item past the end carries no data: the barrier is completed by a plain arrival so that the warps
// waiting for the buffer wake up and see it.
__device__ __noinline__ void stage_tile(const HitsParams &P, uint8_t *smem_raw, int buf, uint32_t item)
{
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kOffFull) + buf;
    reinterpret_cast<uint32_t *>(smem_raw + kOffItem)[buf] = item;
    if (item < (uint32_t)(P.n_tiles * P.n_blocks)) {
        const int64_t tile_start = (P.tile0 + (int64_t)(item % (uint32_t)P.n_tiles)) * kTilePos;
        // order earlier generic-proxy reads of the buffer before the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(full, kTileWords * 4 + kTileMask * 4);
        tma_load_1d(smem_raw + buf * kTileWords * 4, P.packed + (tile_start >> 4), kTileWords * 4, full);
        tma_load_1d(smem_raw + kOffMask + buf * kTileMask * 4, P.mask + (tile_start >> 5), kTileMask * 4, full);
    } else {
        mbar_arrive(full);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
scan_hits_kernel(const HitsParams P, uint32_t *smem_base_probe)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    if (smem_base_probe) {  // host asks where the dynamic shared window starts (see hits_smem_plan)
        if (threadIdx.x == 0) *smem_base_probe = raw;
        return;
    }
    // Shared memory: the tile ring and bookkeeping first, then the block table on the next 32 KB boundary
    // of the shared window so that  (g << 15) | (x << 7) | (lane << 2)  can be OR-ed into the address,
    // then the candidate queues.
    Smem S;
    S.words = reinterpret_cast<uint32_t *>(smem_raw);
    S.mask = reinterpret_cast<uint32_t *>(smem_raw + kOffMask);
    S.motif = reinterpret_cast<int32_t *>(smem_raw + kOffMotif);
    S.len = reinterpret_cast<int32_t *>(smem_raw + kOffLen);
    S.full = reinterpret_cast<uint64_t *>(smem_raw + kOffFull);
    S.tab = reinterpret_cast<uint64_t *>(smem_raw + kOffTab);
    S.item = reinterpret_cast<uint32_t *>(smem_raw + kOffItem);
    S.done = reinterpret_cast<uint32_t *>(smem_raw + kOffDone);
    const uint32_t lut_addr = (raw + (uint32_t)kSmallBytes + 32767u) & ~32767u;
    S.lut = reinterpret_cast<uint32_t *>(smem_raw + (lut_addr - raw));
    // queues of warps 0 .. kQueueWarpsA-1 in front of the table, the rest behind the largest table
    uint32_t *queue_a = reinterpret_cast<uint32_t *>(smem_raw + kBookBytes);
    uint32_t *queue_b = reinterpret_cast<uint32_t *>(smem_raw + (lut_addr - raw) + P.lut_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t total_items = (uint32_t)(P.n_tiles * P.n_blocks);

    if (tid == 0) {
        for (int i = 0; i < kTileBufs; i++) {
            mbar_init(S.full + i, 1);
            S.done[i] = 0;
        }
        mbar_init(S.tab, 1);
        uint32_t first[kTileBufs];
        for (int i = 0; i < kTileBufs; i++) first[i] = atomicAdd(P.work_counter, 1u);
        for (int i = 0; i < kTileBufs; i++) stage_tile(P, smem_raw, i, first[i]);
    }
    __syncthreads();

    // Every warp walks the same item sequence (round k lives in buffer k % kTileBufs) at its own pace: a
    // warp that is done with a tile moves on to the next one, already staged, instead of waiting for the
    // slowest warp at a CTA barrier; the LAST warp to leave a buffer refills it with round k + kTileBufs.
    // The CTA only synchronises when the motif block (the table) changes.
    int cur_block = -1;
    uint32_t tab_phase = 0;
    uint32_t kinit = kPoison;
    int AB = 1;
    for (uint32_t k = 0;; k++) {
        const int buf = (int)(k % kTileBufs);
        mbar_wait(S.full + buf, (k / kTileBufs) & 1u);
        const uint32_t item = S.item[buf];
        if (item >= total_items) break;
        const int b = (int)(item / (uint32_t)P.n_tiles);
        const int64_t tile_start = (P.tile0 + (int64_t)(item % (uint32_t)P.n_tiles)) * kTilePos;
        if (b != cur_block) {
            __syncthreads();  // every warp is done with the previous block's table (and S.motif / S.len)
            AB = __ldg(P.blk_AB + b);
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const int chunks = 4 * ((AB >> 8) & 255) + (AB & 255);  // 8 KB pieces of the table
                mbar_expect_tx(S.tab, (uint32_t)chunks * 8192u);
                const uint32_t *src = P.qlut + __ldg(P.blk_lut_off + b);
                for (int g = 0; g < chunks; g++) tma_load_1d(S.lut + g * 2048, src + g * 2048, 8192, S.tab);
            }
            if (tid < TFBS_BLOCK_SLOTS) {
                const int m = __ldg(P.blk_motif + b * TFBS_BLOCK_SLOTS + tid);
                S.motif[tid] = m;
                S.len[tid] = m >= 0 ? __ldg(P.lens + m) : 0;
            }
            kinit = __ldg(P.kinit + b * 32 + lane);
            cur_block = b;
            mbar_wait(S.tab, tab_phase);
            tab_phase ^= 1u;
            __syncthreads();  // S.motif / S.len visible
        }

        VerifyCtx V;
        V.s_words = S.words + buf * kTileWords;
        V.s_mask = S.mask + buf * kTileMask;
        V.s_motif = S.motif;
        V.s_len = S.len;
        V.thr = P.thr;
        V.exact = P.exact;
        V.tile_start = tile_start;
        V.owned_end = P.owned_end;
        V.pos_offset = P.pos_offset;
        PushCtx C;
        C.queue = warp < kQueueWarpsA ? queue_a + warp * kQueueCap : queue_b + (warp - kQueueWarpsA) * kQueueCap;
        C.seg_begin = warp * kSeg;
        C.seg_end = C.seg_begin + kSeg;
        C.lane = lane;
        C.V = &V;
        C.hits = P.hits;
        C.hit_count = P.hit_count;
        C.cap = P.cap;
        C.cand = P.cand;
        C.cand_cap = P.cand_cap;
        const uint32_t lut_lane_addr = lut_addr + (uint32_t)lane * 4u;
        switch (AB) {
#define TFBS_CASE(a, b, n) \
    case ((n << 16) | (a << 8) | b): scan_segment<a, b, n>(C, V.s_words, lut_lane_addr, kinit); break;
            // every (A, B) tfbs_hits_group_plan can return for block lengths 1..32 ...
            TFBS_CASE(1, 0, 0) TFBS_CASE(2, 0, 0) TFBS_CASE(3, 0, 0) TFBS_CASE(4, 0, 0) TFBS_CASE(5, 0, 0)
            TFBS_CASE(6, 0, 0) TFBS_CASE(5, 2, 0) TFBS_CASE(5, 3, 0) TFBS_CASE(5, 4, 0)
            // ... and the narrow form of those with at most TFBS_NARROW_MAX_GROUPS groups
            TFBS_CASE(1, 0, 1) TFBS_CASE(2, 0, 1) TFBS_CASE(3, 0, 1) TFBS_CASE(4, 0, 1) TFBS_CASE(5, 0, 1)
            TFBS_CASE(6, 0, 1) TFBS_CASE(5, 2, 1) TFBS_CASE(5, 3, 1) TFBS_CASE(5, 4, 1)
            // ... and the sticky narrow form (flag bits kept by masking, >= 117 quantisation steps) of
            // those with at least TFBS_STICKY_MIN_GROUPS groups
            TFBS_CASE(5, 0, 2) TFBS_CASE(6, 0, 2) TFBS_CASE(5, 2, 2) TFBS_CASE(5, 3, 2) TFBS_CASE(5, 4, 2)
#undef TFBS_CASE
            default:  // no instantiation for this grouping: report it instead of silently skipping
                if (tid == 0) atomicExch(P.hit_count + 2, (unsigned long long)(0x100000000ull | (unsigned)(AB & 0xFFFF)));
                break;
        }
        // leave the buffer; the last warp out refills it with the item kTileBufs rounds ahead
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();  // this warp's reads of the tile precede its departure
            if (atomicAdd(S.done + buf, 1u) == (uint32_t)kWarps - 1u) {
                __threadfence_block();
                S.done[buf] = 0;
                stage_tile(P, smem_raw, buf, atomicAdd(P.work_counter, 1u));
            }
        }
    }
}

// Second pass: exact re-scoring of the candidate list with the oracle's arithmetic.  One thread per
// candidate, plain grid-stride kernel at full occupancy: the list is ordered by (motif block, tile), so
// the exact matrices of 64 motifs and the bases of one tile stay L1-resident.  Hits of a CTA batch are
// compacted with ballot/popc and ONE global atomic.
constexpr int kVerifyThreads = 256;
struct VerifyParams {
    const uint32_t *packed;
    const uint32_t *mask;
    const int32_t *lens;
    const float *thr;
    const double *exact;
    const double *exact2;
    const double *guard_eps;
    double guard_scale;      // 1.0; TFBS_VERIFY_GUARD_SCALE inflates the guard so that tests can push every
                             // candidate through the oracle-order slow path
    const unsigned long long *cand;
    unsigned long long cand_cap;
    tfbs_hit *hits;
    unsigned long long *hit_count;
    unsigned long long cap;
    int64_t owned_end, pos_offset;
};

// Exact score of one candidate record with the oracle's arithmetic; true when it is a hit.
__device__ __forceinline__ bool verify_record(const VerifyParams &P, unsigned long long rec, int &m, int &strand,
                                              int64_t &gpos, float &f)
{
    if (rec == kCandEmpty || (int64_t)(rec >> (kCandMotifBits + 1)) >= P.owned_end) return false;
    strand = (int)(rec & 1u);
    m = (int)((rec >> 1) & ((1u << kCandMotifBits) - 1u));
    gpos = (int64_t)(rec >> (kCandMotifBits + 1));
    const int L = __ldg(P.lens + m);
    const uint32_t lmask = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    const uint32_t *mp = P.mask + (gpos >> 5);
    const uint32_t inv = __funnelshift_r(__ldg(mp), __ldg(mp + 1), (int)(gpos & 31)) & lmask;
    if (inv) return false;  // a window touching an invalid base is NaN in the oracle, never a hit
    const uint32_t *wp = P.packed + (gpos >> 4);
    const uint32_t a0 = __ldg(wp), a1 = __ldg(wp + 1), a2 = __ldg(wp + 2);
    const int sh = (int)(gpos & 15) * 2;
    const uint64_t bases = (uint64_t)__funnelshift_r(a0, a1, sh) | ((uint64_t)__funnelshift_r(a1, a2, sh) << 32);
    // Fast path: sum the column-PAIR table (half the loads; the first pair is exactly the oracle's
    // first two steps).  The pair order differs from the oracle's left-to-right order by at most
    // guard_eps[m] >= 2 L u sum|entries| in double, so whenever the double sum lies farther than
    // that from both ends of the rounding interval of its float, the oracle's sum rounds to the
    // SAME float — the result is bit-exact by construction.  Otherwise (about one candidate in a
    // million), and for zero / denormal / power-of-two floats, the oracle's loop is replayed.
    const double *mat2 = P.exact2 + (size_t)(m * 2 + strand) * 256;
    double s = 0.0;
    {
        uint64_t b2 = bases;
        const int npairs = (L + 1) >> 1;
        for (int j0 = 0; j0 < npairs; j0 += 8) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = j0 + k < npairs ? __ldg(mat2 + (j0 + k) * 16 + (int)((b2 >> (4 * k)) & 15u)) : 0.0;
            b2 >>= 32;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (j0 + k < npairs) s += v[k];
        }
    }
    f = (float)s;
    bool exact_enough = isinf(s);   // a -inf entry gives -inf in any order (+inf / NaN entries are rejected at upload)
    if (!exact_enough) {
        const uint32_t fb = __float_as_uint(f);
        const uint32_t ex = fb & 0x7F800000u;
        if (ex != 0u && ex != 0x7F800000u && (fb & 0x007FFFFFu) != 0u) {
            const double half_ulp = (double)__uint_as_float(ex) * 5.9604644775390625e-8;  // 2^-24 * 2^e
            exact_enough = half_ulp - fabs(s - (double)f) > __ldg(P.guard_eps + m) * P.guard_scale;
        }
    }
    if (!exact_enough) {
        const double *mat = P.exact + ((size_t)(m * 2 + strand) * TFBS_MAX_MOTIF_LEN) * 4;
        s = 0.0;  // left-to-right double sum (the oracle's order)
        for (int j = 0; j < L; j++) s += __ldg(mat + j * 4 + (int)((bases >> (2 * j)) & 3u));
        f = (float)s;
    }
    return f >= __ldg(P.thr + m);
}

__global__ void __launch_bounds__(kVerifyThreads) verify_candidates_kernel(const VerifyParams P)
{
    __shared__ uint32_t s_warp_hits[kVerifyThreads / 32];
    __shared__ unsigned long long s_base;
    const unsigned long long n = min(P.hit_count[kCandCountSlot], P.cand_cap);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned long long base = (unsigned long long)blockIdx.x * kVerifyThreads; base < n;
         base += (unsigned long long)gridDim.x * kVerifyThreads) {
        const unsigned long long i = base + threadIdx.x;
        int m = 0, strand = 0;
        int64_t gpos = 0;
        float f = 0.f;
        const bool hit = verify_record(P, i < n ? P.cand[i] : kCandEmpty, m, strand, gpos, f);
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
        if (lane == 0) s_warp_hits[warp] = (uint32_t)__popc(ballot);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < kVerifyThreads / 32; w++) {
                const uint32_t c = s_warp_hits[w];
                s_warp_hits[w] = total;
                total += c;
            }
            s_base = total ? atomicAdd(P.hit_count, (unsigned long long)total) : 0ull;
        }
        __syncthreads();
        if (hit)
            store_hit(P.hits, s_base + s_warp_hits[warp] + __popc(ballot & ((1u << lane) - 1u)), P.cap, gpos + P.pos_offset, strand, m, f);
        __syncthreads();  // s_warp_hits / s_base are rewritten by the next batch
    }
}

// ---- direct path: motifs of at most 12 columns at sparse thresholds -----------------------------------
// For a short motif the set of L-mers that reach the threshold is small (rate x 4^L: 1 678 twelve-mers at a
// 1e-4 hit rate) and is enumerated on the host, exactly (plan_direct).  Every enumerated L-mer, extended to
// 12 bases with all possible suffixes, is a key of a hash table in global memory (value = motif, strand);
// a bitmap over the first 10 bases (2^20 bits = 128 KB, shared memory) screens the positions.  Lanes are
// POSITIONS here: a thread slices its 12-mer out of the packed words, tests the bitmap (a random, mostly
// conflict-free 4-byte shared load for 32 positions instead of G warp-wide table loads per base for 128
// fields) and only on a set bit walks the hash chain.  Matches are candidates like any other: they go to
// the candidate list and the verify pass re-scores them (they are all hits up to windows with invalid
// bases).  Takes the three or four cheapest blocks — 8 of 67 table loads per base on the 800-PWM library —
// off the shared-memory pipe.
constexpr int kDirectKeyBases = 12;          // hash key: 12 bases = 24 bits
constexpr int kDirectOffBits = 2;            // window offset inside the motif (motifs of 13 / 14 columns)
constexpr int kDirectThreads = 1024;
constexpr int kDirectPerThread = 8;          // CONSECUTIVE positions per thread: two packed words cover their 19 bases
constexpr int kDirectQueue = 4096;
constexpr int kDirectSpansPerCheck = 4;      // spans between two looks at the queue (one CTA barrier each)
constexpr int kDirectBitmapWords = 1 << 15;  // 2^20 bits
struct DirectParams {
    const uint32_t *packed;
    int64_t pos0, pos1;          // window starts scanned by this launch
    const uint32_t *bitmap;      // [2^15] words: first 10 bases
    const uint2 *hash;           // {12-mer key, ((motif << 1) | strand) + 1}; .y == 0: empty slot
    uint32_t hash_mask, hash_shift;
    VerifyParams V;              // candidate list / hits / counters (V.cand is written here)
    unsigned long long *cand;
};

__device__ __noinline__ void direct_overflow(const DirectParams &P, unsigned long long rec)
{
    // candidate list full (flood thresholds never take the direct path, so this is a safety net): verify here
    int m, strand;
    int64_t gpos;
    float f;
    if (verify_record(P.V, rec, m, strand, gpos, f))
        store_hit(P.V.hits, atomicAdd(P.V.hit_count, 1ull), P.V.cap, gpos + P.V.pos_offset, strand, m, f);
}

__global__ void __launch_bounds__(kDirectThreads, 1) scan_direct_kernel(const DirectParams P)
{
    extern __shared__ uint32_t s_direct[];
    uint32_t *s_bm = s_direct;                                                                      // [2^15]
    unsigned long long *s_q = reinterpret_cast<unsigned long long *>(s_direct + kDirectBitmapWords);  // [kDirectQueue]
    __shared__ uint32_t s_qn;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x;
    for (int i = tid; i < kDirectBitmapWords / 4; i += kDirectThreads)
        reinterpret_cast<uint4 *>(s_bm)[i] = __ldg(reinterpret_cast<const uint4 *>(P.bitmap) + i);
    if (tid == 0) s_qn = 0;
    __syncthreads();
    constexpr int kSpan = kDirectThreads * kDirectPerThread;
    static_assert(kDirectPerThread == 8, "a thread's 8 positions + 11 bases must fit two packed words");
    const int64_t base0 = P.pos0 & ~(int64_t)7;   // spans start on a multiple of 8 bases (pos0 is a tile boundary anyway)
    // keys sit up to 3 bases after the window start they stand for: look 8 positions past the range
    const int64_t q_end = P.pos1 + 8;
    const int64_t n_spans = (q_end - base0 + kSpan - 1) / kSpan;
    int since_check = 0;
    for (int64_t span = blockIdx.x; span < n_spans; span += gridDim.x) {
        const int64_t p0 = base0 + span * kSpan + (int64_t)tid * kDirectPerThread;   // multiple of 8
        if (p0 < q_end) {
            const uint32_t *wp = P.packed + (p0 >> 4);
            const uint32_t a0 = __ldg(wp), a1 = __ldg(wp + 1);
            const uint64_t w = (((uint64_t)a1 << 32) | a0) >> ((int)(p0 & 15) * 2);   // bases p0 .. p0+23 (or +31)
            uint32_t pass = 0;
#pragma unroll
            for (int r = 0; r < kDirectPerThread; r++) {
                const uint32_t k10 = (uint32_t)(w >> (2 * r)) & 0xFFFFFu;
                pass |= ((s_bm[k10 >> 5] >> (k10 & 31u)) & 1u) << r;
            }
            while (pass) {
                const int r = __ffs(pass) - 1;
                pass &= pass - 1u;
                const int64_t q = p0 + r;
                const uint32_t key = (uint32_t)(w >> (2 * r)) & 0xFFFFFFu;  // 12 bases
                for (uint32_t i = (key * 0x9E3779B1u) >> P.hash_shift;; i = (i + 1u) & P.hash_mask) {
                    const uint2 e = __ldg(P.hash + i);
                    if (e.y == 0u) break;
                    if (e.x == key) {
                        const uint32_t v = e.y - 1u;
                        const int64_t p = q - (int64_t)(v & ((1u << kDirectOffBits) - 1u));   // window start
                        if (p < P.pos0 || p >= P.pos1) continue;
                        const unsigned long long rec = ((unsigned long long)p << (kCandMotifBits + 1)) | (unsigned long long)(v >> kDirectOffBits);
                        const uint32_t slot = atomicAdd(&s_qn, 1u);
                        if (slot < (uint32_t)kDirectQueue) s_q[slot] = rec;
                        else direct_overflow(P, rec);
                    }
                }
            }
        }
        // the queue is looked at every kDirectSpansPerCheck spans and after the last one (CTA-uniform)
        const bool last = span + gridDim.x >= n_spans;
        if (++since_check < kDirectSpansPerCheck && !last) continue;
        since_check = 0;
        __syncthreads();
        const uint32_t qn = min(s_qn, (uint32_t)kDirectQueue);
        if (qn >= (uint32_t)kDirectQueue / 2 || last) {
            if (tid == 0 && qn) {
                s_base = atomicAdd(P.V.hit_count + kCandCountSlot, (unsigned long long)qn);
                atomicAdd(P.V.hit_count + 1, (unsigned long long)qn);
            }
            __syncthreads();
            for (uint32_t i = tid; i < qn; i += kDirectThreads) {
                const unsigned long long slot = s_base + i;
                if (slot < P.V.cand_cap) P.cand[slot] = s_q[i];
                else direct_overflow(P, s_q[i]);
            }
            __syncthreads();
            if (tid == 0) s_qn = 0;
        }
        __syncthreads();
    }
}

struct HitLess {
    bool operator()(const tfbs_hit &a, const tfbs_hit &b) const
    {
        if (a.motif != b.motif) return a.motif < b.motif;
        const int64_t pa = a.pos < 0 ? ~a.pos : a.pos, pb = b.pos < 0 ? ~b.pos : b.pos;
        if (pa != pb) return pa < pb;
        return (a.pos < 0) < (b.pos < 0);
    }
};

// key = (motif, window start, strand): the order PositionSpecificScoringMatrix.search yields hits in
// (by position, forward before reverse), motif-major
__global__ void __launch_bounds__(256) hit_keys_kernel(const tfbs_hit *__restrict__ hits, unsigned long long n,
                                                       unsigned long long *__restrict__ keys)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const tfbs_hit h = hits[i];
        const unsigned long long rev = h.pos < 0;
        const unsigned long long pos = (unsigned long long)(rev ? ~h.pos : h.pos);
        keys[i] = ((unsigned long long)(uint32_t)h.motif << 43) | (pos << 1) | rev;
    }
}

// Device-side ordering of the first n hits in ctx->hits (CUB radix sort of 64-bit keys carrying the
// 16-byte records).  The sorted records end up back in ctx->hits.
int sort_hits_device(tfbs_ctx *ctx, int64_t n, int M)
{
    if (n <= 1) return TFBS_OK;
    int motif_bits = 1;
    while ((1 << motif_bits) < M) motif_bits++;
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const tfbs_hit *)nullptr, (tfbs_hit *)nullptr, (int64_t)n, 0, 43 + motif_bits, ctx->stream);
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t need = al((size_t)n * 8) * 2 + al((size_t)n * sizeof(tfbs_hit)) + al(temp_bytes) + 256;
    int rc = tfbs_scratch_reserve(ctx->sort, need);
    if (rc) return rc;
    char *p = (char *)ctx->sort.ptr;
    unsigned long long *keys = (unsigned long long *)p;
    p += al((size_t)n * 8);
    unsigned long long *keys_out = (unsigned long long *)p;
    p += al((size_t)n * 8);
    tfbs_hit *hits_out = (tfbs_hit *)p;
    p += al((size_t)n * sizeof(tfbs_hit));
    void *temp = p;
    tfbs_hit *hits = (tfbs_hit *)ctx->hits.ptr;
    int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    hit_keys_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(hits, (unsigned long long)n, keys);
    tfbs_launched(ctx);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_out, hits, hits_out, (int64_t)n, 0,
                                                    43 + motif_bits, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hits, hits_out, (size_t)n * sizeof(tfbs_hit), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) {
        tfbs_set_error("sorting hits on the device failed: %s", cudaGetErrorString(e));
        return TFBS_ERR_CUDA;
    }
    return TFBS_OK;
}

// Capacity of the candidate list for a call that may return `cap` hits: twice as many entries (the
// filters pass 1.1 - 2x the hits at the rates they are chosen for), at least 1 Mi, at most 2 GiB of
// records.  What does not fit is verified inside the scan kernel, so this only tunes speed.
int reserve_candidates(tfbs_ctx *ctx, int64_t cap, unsigned long long *cand_cap)
{
    const int64_t want = std::min<int64_t>(std::max<int64_t>(2 * cap, (int64_t)1 << 20), (int64_t)1 << 28);
    int rc = tfbs_scratch_reserve(ctx->cand, (size_t)want * 8);
    if (rc) return rc;
    *cand_cap = ctx->cand.bytes / 8;
    // TFBS_CAND_CAP=<entries> shrinks the list (tests of the overflow path: what does not fit is verified
    // inside the scan kernel)
    const char *env = getenv("TFBS_CAND_CAP");
    if (env && env[0] >= '0' && env[0] <= '9') *cand_cap = std::min<unsigned long long>(*cand_cap, strtoull(env, nullptr, 10));
    return TFBS_OK;
}

void fill_verify_params(VerifyParams &V, const HitsParams &P)
{
    V.packed = P.packed;
    V.mask = P.mask;
    V.lens = P.lens;
    V.thr = P.thr;
    V.exact = P.exact;
    V.exact2 = P.exact2;
    V.guard_eps = P.guard_eps;
    V.guard_scale = 1.0;
    if (const char *env = getenv("TFBS_VERIFY_GUARD_SCALE")) {
        const double g = atof(env);
        if (g >= 1.0) V.guard_scale = g;
    }
    V.cand = P.cand;
    V.cand_cap = P.cand_cap;
    V.hits = P.hits;
    V.hit_count = P.hit_count;
    V.cap = P.cap;
    V.owned_end = P.owned_end;
    V.pos_offset = P.pos_offset;
}

void launch_direct(tfbs_ctx *ctx, const tfbs_pwms *pw, const HitsParams &P, int64_t pos0, int64_t pos1)
{
    if (pw->n_direct <= 0 || pos1 <= pos0) return;
    DirectParams D;
    D.packed = P.packed;
    D.pos0 = pos0;
    D.pos1 = pos1;
    D.bitmap = pw->direct_bitmap_dev;
    D.hash = pw->direct_hash_dev;
    D.hash_mask = pw->direct_hash_mask;
    D.hash_shift = pw->direct_hash_shift;
    fill_verify_params(D.V, P);
    D.cand = P.cand;
    const size_t smem = (size_t)kDirectBitmapWords * 4 + (size_t)kDirectQueue * 8;
    cudaFuncSetAttribute(scan_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_direct_kernel<<<(unsigned)ctx->sm_count, kDirectThreads, smem, ctx->stream>>>(D);
    tfbs_launched(ctx);
}

void launch_verify(tfbs_ctx *ctx, const HitsParams &P)
{
    VerifyParams V;
    fill_verify_params(V, P);
    verify_candidates_kernel<<<(unsigned)(ctx->sm_count * 8), kVerifyThreads, 0, ctx->stream>>>(V);
    tfbs_launched(ctx);
}

// Dynamic shared memory to request: small buffers + padding up to the next 32 KB boundary of the
// shared window + the largest block table.  The window's base offset is a property of the kernel
// (reserved + static shared memory), asked once from the kernel itself.
int hits_smem_plan(tfbs_ctx *ctx, size_t lut_bytes, size_t *smem_out)
{
    static thread_local int64_t cached_base = -1;
    if (cached_base < 0) {
        int rc = tfbs_scratch_reserve(ctx->misc, 64);
        if (rc) return rc;
        HitsParams dummy;
        memset(&dummy, 0, sizeof(dummy));
        scan_hits_kernel<<<1, 32, 1024, ctx->stream>>>(dummy, (uint32_t *)ctx->misc.ptr);
        uint32_t base = 0;
        cudaError_t e = cudaMemcpyAsync(&base, ctx->misc.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            tfbs_set_error("scan_hits: shared-memory probe failed: %s", cudaGetErrorString(e));
            return TFBS_ERR_CUDA;
        }
        cached_base = base;
    }
    const size_t lut_at = ((size_t)cached_base + kSmallBytes + 32767) / 32768 * 32768;
    *smem_out = lut_at - (size_t)cached_base + lut_bytes + kQueueBytes;
    if (*smem_out > 227 * 1024) {
        tfbs_set_error("scan_hits: %zu bytes of shared memory needed, 232448 available", *smem_out);
        return TFBS_ERR_INVALID;
    }
    return TFBS_OK;
}

}  // namespace

// Column grouping of a block whose longest motif has L columns: A groups of 4 columns followed by
// B groups of 3, minimising A + B (shared-memory loads per window) under the table budget
// (32 KB per 4-mer group, 8 KB per 3-mer group).
void tfbs_hits_group_plan(int L, int *A_out, int *B_out)
{
    // minimise A + B; among the minimal groupings take the one with the most 4-column groups, which
    // leaves only 9 distinct (A, B) pairs for L = 1..32 (kernel instantiations, compile time)
    int best_a = 0, best_b = (L + 2) / 3, best_groups = 1 << 30;
    for (int a = 0; a <= 6; a++)
        for (int b = 0; b <= kMaxG; b++) {
            if (a + b < 1 || 4 * a + 3 * b < L) continue;
            if (a * 32768 + b * 8192 > kLutBudget) continue;
            if (a + b < best_groups || (a + b == best_groups && a > best_a)) {
                best_groups = a + b;
                best_a = a;
                best_b = b;
            }
        }
    *A_out = best_a;
    *B_out = best_b;
}

// ---- block plan + filter tables (host) ------------------------------------------------------------
// Cost model of the block form, in warp-wide shared-memory loads per window start, fitted to B200
// measurements at 100 Mb x 800 PWMs and hit rates 1e-6 .. 1e-2 (profiles/r02/SUMMARY.md): one load
// costs 3.75 ps per window; a candidate (queue + list hand-over in the scan, exact re-score in the
// verify kernel) costs 0.08 - 0.10 ns = 21 / 24 / 26 loads in a wide / narrow / sticky block (0.16 -
// 0.22 ns while the scan kernel verified in place, round 1); a load of the sticky form costs
// kStickyLoad plain loads (one more ALU op each).  A narrow block of G groups replaces two wide blocks
// of G and G2 groups.
constexpr double kCandidateCost[3] = {21.0, 24.0, 26.0};
constexpr double kStickyLoad = 1.08;
constexpr int kSaturatingSteps[] = {48, 64};  // step counts tried for saturating plain 8-bit fields (32, 96 and
                                              // 124 were almost never the best on the 800-PWM library)
enum { kWide = 0, kNarrow = 1, kSticky = 2 };

// Block form forced from the environment (A/B measurements and tests); -1 = cost model (default).
//   TFBS_HITS_NARROW=<g>  plain narrow blocks wherever the block has at most g groups (0 = never)
//   TFBS_HITS_STICKY=<g>  sticky narrow blocks wherever the block has at least g groups (>= 5)
static int env_int(const char *name)
{
    const char *env = getenv(name);
    if (!env || env[0] < '0' || env[0] > '9') return -1;
    return atoi(env);
}

// Cut `M` motifs (lengths in descending order) into blocks.  At each step the block is sized for its
// longest motif; if that needs at most TFBS_NARROW_MAX_GROUPS groups and more than 32 motifs are left,
// accept(first, count, A, B) decides whether the next (up to) 64 motifs form one narrow block
// (returns kNarrow or kSticky) or a wide one (kWide, the first 32 of them).
// emit(block) is called for every block before the next one is considered.
template <class Accept, class Emit>
static void cut_blocks(const int32_t *lens_desc, int M, Accept accept, Emit emit)
{
    for (int i = 0; i < M;) {
        tfbs_hits_block K;
        tfbs_hits_group_plan(lens_desc[i], &K.A, &K.B);
        K.first = i;
        K.narrow = kWide;
        if (K.A + K.B <= TFBS_NARROW_MAX_GROUPS && M - i > 32) K.narrow = accept(i, std::min(M - i, 64), K.A, K.B);
        K.count = std::min(M - i, K.narrow ? 64 : 32);
        emit(K);
        i += K.count;
    }
}

// Fill the filter table and the initial field values of one block.  `lut` (A * 8192 + B * 2048 words)
// and kinit (32 words) are overwritten.  Everything rounds in the direction that keeps {candidates} a
// superset of {hits}.  Returns the expected number of candidates per window start (all fields of the
// block, uniform i.i.d. bases) for a narrow or sticky block, 0 for a wide one (not needed by the model).
//
// fill_block_lanes does the work for the motif slots on lanes [lane0, lane1) only — a lane's table words and
// its initial value belong to that lane alone, so disjoint lane ranges of one block can be filled by
// different host threads — and expects `lut` zeroed and `kinit` poisoned by the caller; field_expected
// (may be nullptr) receives the expected candidates of every field, [slot][strand].
static void fill_block_lanes(const tfbs_pwms *pw, const float *thr, const int32_t *motifs, int count, int A, int B,
                             int mode, uint32_t *lut, uint32_t *kinit, int lane0, int lane1, double *field_expected)
{
    const int G = A + B;
    const bool narrow = mode != kWide;
    const int FB = narrow ? 8 : 16;                   // field width
    const int fmax = (1 << (FB - 1)) - 1;             // 0x7F / 0x7FFF: largest value with the flag bit clear
    // group g: first column, width, table offset (words) inside the block table
    int gcol[2 * kMaxG], gk[2 * kMaxG], goff[2 * kMaxG];
    for (int g = 0; g < G; g++) {
        gk[g] = g < A ? 4 : 3;
        gcol[g] = g < A ? 4 * g : 4 * A + 3 * (g - A);
        goff[g] = g < A ? g * 256 * 32 : A * 256 * 32 + (g - A) * 64 * 32;
    }
    std::vector<double> ent((size_t)G * 256);  // double k-mer partial sums of one (motif, strand)
    for (int l = 0; l < count; l++) {
        const int m = motifs[l];
        const int lane = l & 31, half = l >> 5;
        if (lane < lane0 || lane >= lane1) continue;
        const int L = pw->lens[m];
        const float t = thr[m];
        for (int s = 0; s < 2; s++) {
            const int shift = 32 - FB * (2 * half + s + 1);  // wide: fwd 16, rev 0; narrow: 24, 16, 8, 0
            const double *mat = pw->exact_host.data() + ((size_t)(m * 2 + s) * TFBS_MAX_MOTIF_LEN) * 4;
            double gmax[2 * kMaxG];
            double best = 0.0;
            bool reachable = true;
            int Gm = 0;  // groups that contain at least one column of this motif
            for (int g = 0; g < G && gcol[g] < L; g++, Gm++) {
                const int E = 1 << (2 * gk[g]);
                double mx = -INFINITY;
                for (int x = 0; x < E; x++) {
                    double v = 0.0;
                    for (int j = 0; j < gk[g] && gcol[g] + j < L; j++) v += mat[(gcol[g] + j) * 4 + ((x >> (2 * j)) & 3)];
                    ent[(size_t)g * 256 + x] = v;
                    mx = std::max(mx, v);
                }
                gmax[g] = mx;
                best += mx;
                if (!std::isfinite(mx)) reachable = false;  // a column group of -inf only
            }
            int slack = 0;
            if (mode == kSticky) {
                // Carries the field receives from the byte below it.  Bytes from the top: slot l forward,
                // slot l reverse, slot l + 32 forward, slot l + 32 reverse.  A field starts at most at 2
                // (127 - Dq - slack with Dq >= 125 - slack), adds one entry <= 127 per group of its motif
                // and the carries it receives itself.  Its first crossing of 128 only sets the flag; from
                // then on the OR puts the flag back after every wrap, so the low 7 bits count modulo 128
                // and every further 128 sends one carry up: at most (2 + 127 Gl + c_in) / 128 - 1 carries.
                // (Round 1 divided by 256 — the flag-restoring OR was forgotten — and lost hits of sharp
                // motifs at tight thresholds: tests/test_host.py::test_sticky_carry_slack_adversarial.)
                // An unused or unreachable field keeps zero entries and never carries.
                auto groups_of = [&](int slot) {
                    int n = 0;
                    if (slot < count)
                        for (int g = 0; g < G && gcol[g] < pw->lens[motifs[slot]]; g++) n++;
                    return n;
                };
                auto carries_out = [](int groups, int c_in) { return groups ? std::max(0, (2 + 127 * groups + c_in) / 128 - 1) : 0; };
                const int lo = l & 31, hi = lo + 32;
                const int c0 = carries_out(groups_of(hi), 0);   // out of slot hi reverse  -> into slot hi forward
                const int c1 = carries_out(groups_of(hi), c0);  // out of slot hi forward  -> into slot lo reverse
                const int c2 = carries_out(groups_of(lo), c1);  // out of slot lo reverse  -> into slot lo forward
                slack = half == 0 ? (s == 0 ? c2 : c1) : (s == 0 ? c0 : 0);
            }
            if (std::isnan(t)) reachable = false;
            // D: how much a hit may lose against the best possible score (plus slack for the
            // float rounding of the oracle's comparison and double summation order)
            double D = reachable ? (best - (double)t) : -1.0;
            if (reachable && std::isfinite(D)) D += std::fabs((double)t) * std::ldexp(1.0, -22) + std::fabs(best) * 1e-12 + 1e-9;
            if (!reachable || D < 0.0) continue;  // field stays poisoned: never a candidate

            // Quantise the field with dq_max steps for the budget D and entries capped at capq; with
            // `commit` the table and the initial value are written.  Returns the probability that a
            // window of uniform random bases passes (narrow forms only).
            auto quantise = [&](int dq_max, int capq, bool commit) -> double {
                double scale;
                int Dq;
                if (std::isinf(D)) {
                    scale = 0.0;  // threshold -inf: everything is a candidate
                    Dq = 1;
                } else {
                    scale = (double)dq_max / std::max(D, 1e-300);
                    if (!std::isfinite(scale)) scale = 1e300;
                    const double dqv = std::floor(D * scale) + 1.0;
                    Dq = (int)std::min(dqv, (double)(dq_max + 1));
                    Dq = std::max(Dq, dq_max);  // D * scale == dq_max up to rounding; keeps the no-carry bound
                }
                if (commit)
                    kinit[lane] = (kinit[lane] & ~((uint32_t)(2 * fmax + 1) << shift)) | ((uint32_t)(fmax - Dq - slack) << shift);
                const int Dp = Dq + slack;  // what the field lets pass when no carry comes in
                // pass[v]: probability that the quantised deficits of the groups so far sum to v <= Dp
                double pass[128], next[128];
                if (narrow) {
                    std::fill(pass, pass + 128, 0.0);
                    pass[0] = 1.0;
                }
                for (int g = 0; g < Gm; g++) {
                    const int E = 1 << (2 * gk[g]);
                    double hist[128];
                    int qmax = 0;  // largest quantised deficit of the group that can still pass
                    if (narrow) std::fill(hist, hist + 128, 0.0);
                    for (int x = 0; x < E; x++) {
                        const double d = gmax[g] - ent[(size_t)g * 256 + x];  // >= 0, +inf for -inf entries
                        int q;
                        if (scale == 0.0) q = 0;
                        else {
                            const double v = std::floor(d * scale);
                            q = (v >= (double)capq || std::isnan(v)) ? capq : (int)v;
                        }
                        if (commit) lut[(size_t)goff[g] + (size_t)x * 32 + lane] |= (uint32_t)q << shift;
                        if (narrow && q <= Dp) {
                            hist[q] += 1.0 / E;
                            qmax = std::max(qmax, q);
                        }
                    }
                    if (narrow) {
                        std::fill(next, next + Dp + 1, 0.0);
                        for (int v = 0; v <= Dp; v++)
                            if (pass[v] > 0.0)
                                for (int q = 0; q <= qmax && v + q <= Dp; q++) next[v + q] += pass[v] * hist[q];
                        std::copy(next, next + Dp + 1, pass);
                    }
                }
                double p = 0.0;
                if (narrow)
                    for (int v = 0; v <= Dp; v++) p += pass[v];
                return p;
            };

            // Steps and entry cap.  No carry out of a field needs
            //   (fmax - Dq - slack) + Gm * capq <= 2 * fmax + 1   with   Dq >= dq_max.
            double expected = 0.0;   // of this field
            if (mode == kWide) {
                const int dq_max = 0x7FFF / G - 3;  // block-wide bound; an entry above the budget kills alone
                quantise(dq_max, dq_max + 2, true);
            } else if (mode == kSticky) {
                // sticky fields may wrap (their flag is kept by the OR): entries <= 127, and `slack` steps
                // of the field absorb the carries a wrapping neighbour sends in
                expected = quantise(125 - slack, 127, true);  // slack <= 8: >= 117 steps at any G
            } else {
                // Plain 8-bit fields must not wrap.  Either every entry above the budget kills alone
                // (capq = dq_max + 2, which leaves only (128 - 2 Gm) / (Gm - 1) steps), or the steps are
                // finer and entries SATURATE at capq <= (128 + dq_max) / Gm — a window then needs several
                // bad groups to be rejected.  Both keep {candidates} a superset of {hits}; the one that
                // passes fewer random windows (exact, by the convolution above) is taken.
                int best_dq = Gm <= 1 ? 100 : std::min(100, (128 - 2 * Gm) / (Gm - 1));
                int best_cap = best_dq + 2;
                if (Gm >= 3 && std::isfinite(D)) {
                    double best_p = quantise(best_dq, best_cap, false);
                    for (int dq : kSaturatingSteps) {
                        const int cap = (128 + dq) / Gm;
                        if (dq <= best_dq || cap < 2 || cap >= dq + 2) continue;
                        const double p = quantise(dq, cap, false);
                        if (p < best_p) {
                            best_p = p;
                            best_dq = dq;
                            best_cap = cap;
                        }
                    }
                }
                expected = quantise(best_dq, best_cap, true);
            }
            if (field_expected) field_expected[l * 2 + s] = expected;
        }
    }
}

// Expected candidates per window start of a block from its fields' values, summed in slot / strand order
// (the order is fixed so that the cost model decides the same whichever threads filled the lanes).
static double sum_field_expected(const double *field_expected, int count)
{
    double expected = 0.0;
    for (int i = 0; i < 2 * count; i++) expected += field_expected[i];
    return expected;
}

static double fill_block(const tfbs_pwms *pw, const float *thr, const int32_t *motifs, int count, int A, int B,
                         int mode, uint32_t *lut, uint32_t *kinit)
{
    std::fill(lut, lut + (size_t)A * 8192 + (size_t)B * 2048, 0u);
    for (int lane = 0; lane < 32; lane++) kinit[lane] = mode != kWide ? kPoisonNarrow : kPoison;
    double field_expected[2 * TFBS_BLOCK_SLOTS] = {0.0};
    fill_block_lanes(pw, thr, motifs, count, A, B, mode, lut, kinit, 0, 32, field_expected);
    return mode != kWide ? sum_field_expected(field_expected, count) : 0.0;
}

// The cost model's choice for a block of G groups that would replace two wide blocks of G and G2 groups,
// given the candidates per window start its sticky and its plain narrow tables are expected to pass.  The
// sticky filter (>= 117 steps) also estimates the candidates of the wide form: measured, it passes about
// 1.5x as many.
static int cheapest_form(int G, int G2, double c_sticky, double c_narrow)
{
    const bool can_stick = G >= TFBS_STICKY_MIN_GROUPS;
    const double cost_wide = (double)(G + G2) + kCandidateCost[kWide] * c_sticky / 1.5;
    const double cost_narrow = (double)G + kCandidateCost[kNarrow] * c_narrow;
    const double cost_sticky = can_stick ? kStickyLoad * G + kCandidateCost[kSticky] * c_sticky : 1e300;
    if (cost_narrow <= cost_wide && cost_narrow <= cost_sticky) return kNarrow;
    if (cost_sticky <= cost_wide) return kSticky;
    return kWide;
}

// Form of the block made of motifs order[first, first + count) (count > 32; `order` = the motifs on the table
// filter for this threshold vector, longest first, lens_desc their lengths): forced from the
// environment, else the cheapest by the cost model.  The candidate forms are built in place in
// lut / kinit; on return they hold the tables of the chosen form unless that is kWide (the caller
// builds the wide table of the first 32 motifs).  *expected = candidates per window start the chosen
// narrow / sticky form is expected to pass (0 for kWide).
static int choose_form(const tfbs_pwms *pw, const float *thr, const int32_t *order, const int32_t *lens_desc, int first,
                       int count, int A, int B, uint32_t *lut, uint32_t *kinit, double *expected)
{
    const int G = A + B;
    const int32_t *motifs = order + first;
    const int force_narrow = env_int("TFBS_HITS_NARROW"), force_sticky = env_int("TFBS_HITS_STICKY");
    *expected = 0.0;
    if (force_narrow >= 0 || force_sticky >= 0) {
        int mode = kWide;
        if (force_sticky >= 0 && G >= std::max(force_sticky, TFBS_STICKY_MIN_GROUPS)) mode = kSticky;
        else if (force_narrow >= 0 && G <= force_narrow) mode = kNarrow;
        if (mode != kWide) *expected = fill_block(pw, thr, motifs, count, A, B, mode, lut, kinit);
        return mode;
    }
    // groups of the second of the two wide blocks a narrow block replaces
    int A2 = 0, B2 = 1;
    tfbs_hits_group_plan(lens_desc[first + 32], &A2, &B2);
    // the sticky filter (>= 117 steps) also estimates the candidates of the wide form
    const double c_sticky = fill_block(pw, thr, motifs, count, A, B, kSticky, lut, kinit);
    const double c_narrow = fill_block(pw, thr, motifs, count, A, B, kNarrow, lut, kinit);
    const int form = cheapest_form(G, A2 + B2, c_sticky, c_narrow);
    if (form == kNarrow) *expected = c_narrow;  // tables are in place
    if (form == kSticky) *expected = fill_block(pw, thr, motifs, count, A, B, kSticky, lut, kinit);
    return form;
}

// ---- direct path planning (host) ------------------------------------------------------------------------
constexpr int kDirectMaxLen = 14;            // longest motif the path may take (windows of 12 of its columns)
constexpr int kDirectPrefixBases = 10;       // bitmap: first 10 bases
constexpr size_t kDirectMaxEntries = (size_t)3 << 20;   // 3 Mi keys -> 8 Mi slots = 64 MB of hash
constexpr size_t kDirectMotifMaxEntries = kDirectMaxEntries / 32;   // keys of ONE motif (both strands); more: it stays on the tables
constexpr size_t kDirectSeedBudget = kDirectMaxEntries;             // unexpanded W-mers kept from the count pass (8 B each)
constexpr double kDirectMaxDensity = 0.5;    // bitmap bits set: beyond this the screen stops screening (A/B on B200: 0.35 -> 0.5 gains 0.65 ms)

// Keys of one (motif, strand) field.  L <= 12: every L-mer whose oracle score reaches the threshold —
// depth-first over the columns with the best-possible-remainder bound (slack as in fill_block), partial sums
// in the oracle's left-to-right order so that the leaf test (float)sum >= t IS the oracle's test; each hit
// L-mer becomes 4^(12-L) keys.  L = 13 or 14: the 12-mers at columns off .. off+11 that can still reach the
// threshold when every column outside the window scores its best (a superset of the hits: the verify pass
// rejects the rest); the key is then taken `off` bases after the window start.
// *count = keys of the field (4^(12-W) per enumerated W-mer, W = min(L, 12)).  `seeds` (may be nullptr)
// receives the enumerated W-mers UNEXPANDED as {W-mer, value}, as long as they number at most `seed_cap`
// (*seeds_complete tells whether all of them were recorded); the 12-mer keys are made from them when the
// hash is filled (insert_direct_seeds).  Returns false when the field has more than `cap` keys.
static bool enumerate_direct_field(const double *mat, int L, int off, float t, uint32_t value, std::vector<uint2> *seeds,
                                   size_t seed_cap, bool *seeds_complete, size_t cap, size_t *count)
{
    *count = 0;
    if (seeds_complete) *seeds_complete = true;
    if (std::isnan(t)) return true;
    const int W = std::min(L, kDirectKeyBases);   // window columns off .. off+W-1
    double colmax[TFBS_MAX_MOTIF_LEN], best = 0.0, outside = 0.0;
    for (int j = 0; j < L; j++) {
        colmax[j] = -INFINITY;
        for (int b = 0; b < 4; b++) colmax[j] = std::max(colmax[j], mat[j * 4 + b]);
        best += colmax[j];
        if (j < off || j >= off + W) outside += colmax[j];
    }
    if (!std::isfinite(best)) return true;  // a column of -inf only: no window can score
    double sufbest[kDirectKeyBases + 1];    // best of window columns j.. plus everything outside the window
    sufbest[W] = outside;
    for (int j = W - 1; j >= 0; j--) sufbest[j] = sufbest[j + 1] + colmax[off + j];
    const double need = std::isinf(t) ? (double)t
                                      : (double)t - (std::fabs((double)t) * std::ldexp(1.0, -22) + std::fabs(best) * 1e-12 + 1e-9);
    const bool exact_leaf = W == L;
    int code[kDirectKeyBases];
    double part[kDirectKeyBases + 1];
    part[0] = 0.0;
    int j = 0;
    code[0] = 0;
    const uint32_t n_suffix = 1u << (2 * (kDirectKeyBases - W));
    for (;;) {
        if (code[j] == 4) {
            if (--j < 0) break;
            code[j]++;
            continue;
        }
        const double sc = part[j] + mat[(off + j) * 4 + code[j]];
        if (!(sc + sufbest[j + 1] >= need)) {
            code[j]++;
            continue;
        }
        if (j == W - 1) {
            if (!exact_leaf || (float)sc >= t) {
                uint32_t key = 0;
                for (int i = 0; i < W; i++) key |= (uint32_t)code[i] << (2 * i);
                *count += n_suffix;
                if (*count > cap) return false;
                if (seeds) {
                    if (seeds->size() < seed_cap) seeds->push_back(make_uint2(key, value));
                    else if (seeds_complete) *seeds_complete = false;
                }
            }
            code[j]++;
        } else {
            part[j + 1] = sc;
            code[++j] = 0;
        }
    }
    return true;
}

// Expand seeds (W-mers of one field) to their 4^(12-W) twelve-mer keys and enter them into the bitmap (first
// 10 bases) and the open-addressing hash ({key, value}, linear probing, .y == 0 = empty).  Safe to call from
// several threads at once on different seeds: a slot is claimed by a 64-bit compare-and-swap, a bitmap bit
// set by an atomic OR.  (The position of equal keys in a probe chain may differ from run to run; a look-up
// walks the chain to its first empty slot and reports every match, so results do not.)
static void insert_direct_seeds(const uint2 *seeds, size_t n, int W, uint32_t *bitmap, uint2 *hash, uint32_t bits)
{
    static_assert(sizeof(uint2) == 8, "hash slots are claimed as 64-bit words");
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(hash);
    const uint32_t mask = (uint32_t)(((size_t)1 << bits) - 1);
    const uint32_t pmask = (1u << (2 * kDirectPrefixBases)) - 1u;
    const uint32_t n_suffix = 1u << (2 * (kDirectKeyBases - W));
    for (size_t k = 0; k < n; k++)
        for (uint32_t u = 0; u < n_suffix; u++) {
            const uint32_t key = seeds[k].x | (u << (2 * W));
            const uint32_t k10 = key & pmask;
            const uint32_t bit = 1u << (k10 & 31u);
            if (!(__atomic_load_n(&bitmap[k10 >> 5], __ATOMIC_RELAXED) & bit)) __atomic_fetch_or(&bitmap[k10 >> 5], bit, __ATOMIC_RELAXED);
            const unsigned long long rec = ((unsigned long long)seeds[k].y << 32) | key;   // uint2 {x = key, y = value}
            for (uint32_t i = (key * 0x9E3779B1u) >> (32 - bits);; i = (i + 1u) & mask) {
                unsigned long long expected = 0ull;
                if (__atomic_load_n(&slots[i], __ATOMIC_RELAXED) == 0ull &&
                    __atomic_compare_exchange_n(&slots[i], &expected, rec, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
                    break;
            }
        }
}

// Decide which motifs take the direct path for this threshold vector and build their bitmap + hash.
// Candidates are the motifs of at most Ld columns, Ld = 14, 12, 10, 8: the first set whose keys number at most
// kDirectMaxEntries and whose bitmap stays sparse enough is taken whole.  If no whole set fits, the choice is
// made PER MOTIF: a motif whose own keys number more than kDirectMotifMaxEntries (a loose, flood or -inf
// threshold) stays on the table filter without taking the rest of the library with it — with "80 % of the
// best score" thresholds the 6-column motifs pass a percent of all windows while the 12-column ones pass 1e-5
// of them.  is_direct[m] = 1 for the motifs taken.  Returns their number (0: none).
// TFBS_HITS_DIRECT=0 switches the path off; forcing a block form (TFBS_HITS_NARROW / _STICKY, A/B runs and
// tests of the table filter) does too.
static int plan_direct(const tfbs_pwms *pw, const float *thr, std::vector<uint32_t> &bitmap, std::vector<uint2> &hash,
                       uint32_t *hash_shift, int *direct_len, int64_t *n_entries, std::vector<uint8_t> &is_direct)
{
    const int M = pw->M;
    is_direct.assign((size_t)M, 0);
    if (env_int("TFBS_HITS_DIRECT") == 0 || env_int("TFBS_HITS_NARROW") >= 0 || env_int("TFBS_HITS_STICKY") >= 0) return 0;
    int n_max = 0;   // motifs of at most kDirectMaxLen columns: the tail of the length-sorted order
    while (n_max < M && pw->lens_desc[M - 1 - n_max] <= kDirectMaxLen) n_max++;
    // worth a separate pass only if it takes at least one whole table block off the scan
    if (n_max < 32) return 0;
    cpu_set_t set;
    const unsigned cores = sched_getaffinity(0, sizeof(set), &set) == 0 ? (unsigned)CPU_COUNT(&set) : std::thread::hardware_concurrency();
    unsigned T = std::max(1u, std::min<unsigned>(std::min(cores, 16u), (unsigned)n_max));
    const int forced = env_int("TFBS_PLAN_THREADS");
    if (forced >= 1) T = std::min<unsigned>(T, (unsigned)forced);
    constexpr size_t kTooMany = SIZE_MAX;
    // Count pass, on the host's cores: keys of every candidate field (for the motifs longer than the key, of
    // the 12-column window that passes the fewest 12-mers).  The enumerated W-mers are kept as unexpanded
    // seeds while they are few (a field of at most kDirectMotifMaxEntries seeds, kDirectSeedBudget in all),
    // so that a plan that fits is filled from them without a second enumeration.
    struct Field {
        uint8_t off = 0;          // window offset inside the motif
        bool recorded = false;    // seeds[thread][begin, begin + n) hold every W-mer of the field
        uint32_t thread = 0;
        size_t begin = 0, n = 0;
    };
    std::vector<Field> field((size_t)M * 2);
    std::vector<size_t> keys((size_t)M, kTooMany);     // keys of both strands of a motif
    std::vector<std::vector<uint2>> seeds(T);
    std::atomic<size_t> seeds_total{0};
    {
        std::atomic<int> next{M - n_max};
        auto worker = [&](unsigned t) {
            std::vector<uint2> cur, best;
            for (int i = next.fetch_add(1); i < M; i = next.fetch_add(1)) {
                const int m = pw->order[i], L = pw->lens[m];
                size_t km = 0;
                for (int s2 = 0; s2 < 2 && km != kTooMany; s2++) {
                    const double *mat = pw->exact_host.data() + ((size_t)(m * 2 + s2) * TFBS_MAX_MOTIF_LEN) * 4;
                    Field &F = field[(size_t)m * 2 + s2];
                    size_t best_c = kTooMany;
                    bool best_complete = false;
                    for (int off = 0; off + std::min(L, kDirectKeyBases) <= L; off++) {
                        const uint32_t value = ((((uint32_t)m << 1) | (uint32_t)s2) << kDirectOffBits | (uint32_t)off) + 1u;
                        size_t c2 = 0;
                        bool complete = false;
                        cur.clear();
                        if (enumerate_direct_field(mat, L, off, thr[m], value, &cur, kDirectMotifMaxEntries, &complete, kDirectMaxEntries, &c2) &&
                            c2 < best_c) {
                            best_c = c2;
                            best_complete = complete;
                            F.off = (uint8_t)off;
                            best.swap(cur);
                        }
                    }
                    km = best_c == kTooMany ? kTooMany : km + best_c;
                    if (best_c != kTooMany && best_complete) {
                        if (seeds_total.fetch_add(best.size()) + best.size() <= kDirectSeedBudget) {
                            F.recorded = true;
                            F.thread = t;
                            F.begin = seeds[t].size();
                            F.n = best.size();
                            seeds[t].insert(seeds[t].end(), best.begin(), best.end());
                        } else {
                            seeds_total.fetch_sub(best.size());
                        }
                    }
                }
                keys[(size_t)m] = km;
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < T; t++) pool.emplace_back(worker, t);
        worker(0);
        for (auto &th : pool) th.join();
    }
    std::vector<int32_t> taken;   // motifs of the set being tried
    // One attempt: the motifs of at most Ld columns — all of them, or (per_motif) those with few enough keys.
    auto attempt = [&](int Ld, bool per_motif) -> int {
        int n = 0;
        while (n < M && pw->lens_desc[M - 1 - n] <= Ld) n++;
        taken.clear();
        size_t total = 0;
        for (int i = M - n; i < M; i++) {
            const int m = pw->order[i];
            const size_t km = keys[(size_t)m];
            if (per_motif ? km > kDirectMotifMaxEntries : km == kTooMany) {
                if (per_motif) continue;
                return 0;
            }
            taken.push_back(m);
            total += km;
        }
        if (per_motif && (int)taken.size() == n) return 0;   // nothing left out: the whole set was tried already
        if ((int)taken.size() < 32 || total > kDirectMaxEntries) return 0;
        // fill the bitmap and the hash from the seeds, field by field on the host's cores (lock-free: a slot
        // is claimed with one compare-and-swap, a bitmap bit set with an atomic OR); a field whose seeds
        // were not kept is enumerated again
        uint32_t bits = 10;
        while (((size_t)1 << bits) < 2 * total + 16) bits++;
        hash.assign((size_t)1 << bits, make_uint2(0u, 0u));
        bitmap.assign(kDirectBitmapWords, 0u);
        {
            std::atomic<size_t> next{0};
            auto worker = [&]() {
                std::vector<uint2> again;
                for (size_t k = next.fetch_add(1); k < 2 * taken.size(); k = next.fetch_add(1)) {
                    const int m = taken[k >> 1], s2 = (int)(k & 1), L = pw->lens[m];
                    const Field &F = field[(size_t)m * 2 + s2];
                    const uint2 *src = F.recorded ? seeds[F.thread].data() + F.begin : nullptr;
                    size_t n_src = F.n;
                    if (!F.recorded) {
                        const double *mat = pw->exact_host.data() + ((size_t)(m * 2 + s2) * TFBS_MAX_MOTIF_LEN) * 4;
                        const uint32_t value = ((((uint32_t)m << 1) | (uint32_t)s2) << kDirectOffBits | (uint32_t)F.off) + 1u;
                        size_t c = 0;
                        again.clear();
                        enumerate_direct_field(mat, L, F.off, thr[m], value, &again, SIZE_MAX, nullptr, kDirectMaxEntries, &c);
                        src = again.data();
                        n_src = again.size();
                    }
                    insert_direct_seeds(src, n_src, std::min(L, kDirectKeyBases), bitmap.data(), hash.data(), bits);
                }
            };
            std::vector<std::thread> pool;
            for (unsigned t = 1; t < T; t++) pool.emplace_back(worker);
            worker();
            for (auto &th : pool) th.join();
        }
        size_t bits_set = 0;
        for (uint32_t w : bitmap) bits_set += (size_t)__builtin_popcount(w);
        double max_density = kDirectMaxDensity;
        if (const char *env = getenv("TFBS_DIRECT_DENSITY")) max_density = atof(env);   // A/B measurements
        if ((double)bits_set > max_density * (double)(1u << (2 * kDirectPrefixBases))) return 0;
        *hash_shift = 32 - bits;
        *direct_len = 0;
        for (int m : taken) {
            is_direct[(size_t)m] = 1;
            *direct_len = std::max(*direct_len, (int)pw->lens[m]);
        }
        *n_entries = (int64_t)total;
        return (int)taken.size();
    };
    for (bool per_motif : {false, true})
        for (int Ld : {kDirectMaxLen, 12, 10, 8})
            if (const int n = attempt(Ld, per_motif)) return n;
    return 0;
}

// The motifs the table filter keeps for a threshold vector: the length-sorted order without the motifs on
// the direct path (still longest first, so block i of this order is no longer than block i of the whole
// library and the device arrays sized at upload are large enough).
static void table_order(const tfbs_pwms *pw, const std::vector<uint8_t> &is_direct, std::vector<int32_t> &order,
                        std::vector<int32_t> &lens_desc)
{
    order.clear();
    lens_desc.clear();
    for (int i = 0; i < pw->M; i++) {
        const int m = pw->order[i];
        if (is_direct[(size_t)m]) continue;
        order.push_back(m);
        lens_desc.push_back(pw->lens[m]);
    }
}

// Everything a hits scan needs for one threshold vector, planned on the host (no CUDA): which motifs take
// the direct path (bitmap + hash), how the others are cut into blocks, each block's form and filter table.
struct HitsPlan {
    int n_direct = 0, direct_len = 0;
    int64_t direct_entries = 0;
    uint32_t direct_shift = 0;
    std::vector<uint32_t> direct_bitmap;
    std::vector<uint2> direct_hash;
    int nb = 0;
    int64_t words = 0;                   // table words of all blocks
    std::vector<uint32_t> qlut, kinit;   // [words], [nb][32]
    std::vector<int32_t> blk_A, blk_B, blk_narrow, blk_count, blk_lut_off, blk_AB, blk_motif;  // [nb] x6, [n_blocks_max][64]
};

// Plan the blocks and quantise their tables for a threshold vector.  Reads only the host fields of `pw`
// (lens, exact_host, order, lens_desc, hq_lut_words, n_blocks_max).
static int plan_hits_tables(const tfbs_pwms *pw, const float *thr, HitsPlan &H)
{
    const int Mall = pw->M;
    // short motifs may bypass the table filter altogether (direct path); the blocks below are cut from the
    // remaining M motifs
    std::vector<uint8_t> is_direct;
    H.n_direct = plan_direct(pw, thr, H.direct_bitmap, H.direct_hash, &H.direct_shift, &H.direct_len, &H.direct_entries, is_direct);
    const int M = Mall - H.n_direct;
    std::vector<int32_t> t_order, t_lens;   // the motifs left on the table filter, longest first
    table_order(pw, is_direct, t_order, t_lens);
    H.qlut.assign((size_t)pw->hq_lut_words, 0u);
    H.kinit.assign((size_t)pw->n_blocks_max * 32, kPoison);
    int64_t words = 0;   // table words of the blocks emitted so far
    int nb = 0;
    const int32_t *order = t_order.data();
    // The form of a block depends only on where it starts, and every block starts at a multiple of 32
    // motifs: the forms (and tables) of all possible starts are worked out speculatively on the host's
    // cores, then the sequential cut below picks the ones it reaches.  (Planning the 800-PWM library
    // one block after the other took 0.2 - 0.3 s per new threshold vector, ten times the scan.)
    struct Spec {
        int form = -1, A = 0, B = 0, count = 0;
        std::vector<uint32_t> lut, kinit;     // tables of the chosen form (narrow or sticky)
        std::vector<uint32_t> lut2, kinit2;   // the sticky candidate while both are being filled
        double fe[2][2 * TFBS_BLOCK_SLOTS];   // expected candidates of every field: [narrow, sticky][slot][strand]
    };
    std::vector<Spec> spec((size_t)(M + 31) / 32);
    std::vector<int> starts;
    for (int i = 0; i < M; i += 32) {
        Spec &S = spec[(size_t)i / 32];
        tfbs_hits_group_plan(t_lens[(size_t)i], &S.A, &S.B);
        S.count = std::min(M - i, 64);
        if (S.A + S.B <= TFBS_NARROW_MAX_GROUPS && M - i > 32) starts.push_back(i);
    }
    {
        cpu_set_t set;
        unsigned cores = sched_getaffinity(0, sizeof(set), &set) == 0 ? (unsigned)CPU_COUNT(&set) : std::thread::hardware_concurrency();
        const bool forced_form = env_int("TFBS_HITS_NARROW") >= 0 || env_int("TFBS_HITS_STICKY") >= 0;
        // work items: a whole start when the form is forced from the environment (choose_form), else one
        // quarter of the lanes of one candidate form of one start — a lane's table words are its own, so
        // the pieces of a block are filled independently and 16 or 32 cores stay busy with 10 - 20 starts
        constexpr int kLaneChunks = 4, kLanesPerChunk = 32 / kLaneChunks;
        const size_t n_items = forced_form ? starts.size() : starts.size() * 2 * kLaneChunks;
        unsigned T = std::max<size_t>(1, std::min<size_t>(std::min(cores, 32u), n_items));
        const int forced = env_int("TFBS_PLAN_THREADS");  // 1 = plan in the calling thread
        if (forced >= 1) T = std::min<unsigned>(T, (unsigned)forced);
        for (int i : starts) {
            Spec &S = spec[(size_t)i / 32];
            const size_t n_words = (size_t)S.A * 8192 + (size_t)S.B * 2048;
            S.lut.assign(n_words, 0u);
            S.kinit.assign(32, forced_form ? kPoison : kPoisonNarrow);
            if (!forced_form) {
                S.lut2.assign(n_words, 0u);
                S.kinit2.assign(32, kPoisonNarrow);
                std::fill(&S.fe[0][0], &S.fe[0][0] + 2 * 2 * TFBS_BLOCK_SLOTS, 0.0);
            }
        }
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (size_t k = next.fetch_add(1); k < n_items; k = next.fetch_add(1)) {
                if (forced_form) {
                    Spec &S = spec[(size_t)starts[k] / 32];
                    double expected;
                    S.form = choose_form(pw, thr, order, t_lens.data(), starts[k], S.count, S.A, S.B, S.lut.data(), S.kinit.data(), &expected);
                    continue;
                }
                const int first = starts[k / (2 * kLaneChunks)];
                const int sticky = (int)(k / kLaneChunks) & 1, chunk = (int)(k % kLaneChunks);
                Spec &S = spec[(size_t)first / 32];
                fill_block_lanes(pw, thr, order + first, S.count, S.A, S.B, sticky ? kSticky : kNarrow,
                                 sticky ? S.lut2.data() : S.lut.data(), sticky ? S.kinit2.data() : S.kinit.data(),
                                 chunk * kLanesPerChunk, (chunk + 1) * kLanesPerChunk, S.fe[sticky]);
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < T; t++) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
        if (!forced_form)
            for (int i : starts) {
                Spec &S = spec[(size_t)i / 32];
                int A2 = 0, B2 = 1;   // groups of the second of the two wide blocks a narrow block replaces
                tfbs_hits_group_plan(t_lens[(size_t)i + 32], &A2, &B2);
                S.form = cheapest_form(S.A + S.B, A2 + B2, sum_field_expected(S.fe[1], S.count), sum_field_expected(S.fe[0], S.count));
                if (S.form == kSticky) {
                    S.lut.swap(S.lut2);
                    S.kinit.swap(S.kinit2);
                }
                std::vector<uint32_t>().swap(S.lut2);
            }
    }
    // a narrow / sticky block's tables were built above, a wide one is built by emit below
    bool fits = true;   // the device arrays were sized at upload for any plan (scan.cu); checked all the same
    auto accept = [&](int first, int count, int A, int B) -> int {
        const Spec &S = spec[(size_t)first / 32];
        (void)count; (void)A; (void)B;
        fits = fits && nb < pw->n_blocks_max && words + (int64_t)S.A * 8192 + (int64_t)S.B * 2048 <= pw->hq_lut_words;
        if (!fits) return (int)kWide;
        if (S.form != kWide) {
            std::copy(S.lut.begin(), S.lut.end(), H.qlut.begin() + words);
            std::copy(S.kinit.begin(), S.kinit.end(), H.kinit.begin() + (size_t)nb * 32);
        }
        return S.form;
    };
    H.blk_motif.assign((size_t)pw->n_blocks_max * TFBS_BLOCK_SLOTS, -1);
    cut_blocks(t_lens.data(), M, accept, [&](const tfbs_hits_block &K) {
        fits = fits && nb < pw->n_blocks_max && words + (int64_t)K.A * 8192 + (int64_t)K.B * 2048 <= pw->hq_lut_words;
        if (!fits) return;
        if (K.narrow == kWide)
            fill_block(pw, thr, order + K.first, K.count, K.A, K.B, kWide, H.qlut.data() + words, H.kinit.data() + (size_t)nb * 32);
        for (int l = 0; l < K.count; l++) H.blk_motif[(size_t)nb * TFBS_BLOCK_SLOTS + l] = order[K.first + l];
        H.blk_A.push_back(K.A);
        H.blk_B.push_back(K.B);
        H.blk_narrow.push_back(K.narrow);
        H.blk_count.push_back(K.count);
        H.blk_lut_off.push_back((int32_t)words);
        H.blk_AB.push_back((K.narrow << 16) | (K.A << 8) | K.B);
        words += (int64_t)K.A * 8192 + (int64_t)K.B * 2048;
        nb++;
    });
    H.nb = nb;
    H.words = words;
    if (!fits) {
        tfbs_set_error("tfbs_scan_hits: the block plan does not fit the tables sized at upload (%d blocks, %lld words)", nb, (long long)words);
        return TFBS_ERR_INVALID;
    }
    return TFBS_OK;
}

// Plan (host) + upload of the filter tables of a threshold vector; cached per threshold vector.
int tfbs_build_hits_tables(tfbs_pwms *pw, const float *thr)
{
    const int Mall = pw->M;
    if (pw->hq_valid && pw->hq_thr.size() == (size_t)Mall && std::memcmp(pw->hq_thr.data(), thr, sizeof(float) * Mall) == 0)
        return TFBS_OK;
    pw->hq_valid = false;   // the device tables are overwritten from here on
    HitsPlan H;
    int rc = plan_hits_tables(pw, thr, H);
    if (rc) return rc;
    const int nb = H.nb;
    pw->blk_A = H.blk_A;
    pw->blk_B = H.blk_B;
    pw->blk_narrow = H.blk_narrow;
    pw->blk_count = H.blk_count;
    pw->blk_lut_off = H.blk_lut_off;
    pw->blk_motif = H.blk_motif;
    pw->n_blocks = nb;
    cudaError_t e = cudaSuccess;
    if (nb > 0) {  // (no table block at all when every motif took the direct path)
        e = cudaMemcpy(pw->hq_lut_dev, H.qlut.data(), (size_t)H.words * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(pw->hq_kinit_dev, H.kinit.data(), (size_t)nb * 32 * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(pw->blk_AB_dev, H.blk_AB.data(), (size_t)nb * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(pw->blk_lut_off_dev, pw->blk_lut_off.data(), (size_t)nb * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(pw->blk_motif_dev, pw->blk_motif.data(), (size_t)nb * TFBS_BLOCK_SLOTS * 4, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        tfbs_set_error("tfbs_scan_hits: uploading filter tables failed: %s", cudaGetErrorString(e));
        return TFBS_ERR_CUDA;
    }
    pw->n_direct = H.n_direct;
    pw->direct_len = H.direct_len;
    pw->direct_entries = H.direct_entries;
    if (H.n_direct > 0) {
        const std::vector<uint2> &d_hash = H.direct_hash;
        if (!pw->direct_bitmap_dev) e = cudaMalloc(&pw->direct_bitmap_dev, (size_t)kDirectBitmapWords * 4);
        if (e == cudaSuccess && pw->direct_hash_slots_dev < d_hash.size()) {
            if (pw->direct_hash_dev) cudaFree(pw->direct_hash_dev);
            pw->direct_hash_dev = nullptr;
            pw->direct_hash_slots_dev = 0;
            e = cudaMalloc(&pw->direct_hash_dev, d_hash.size() * sizeof(uint2));
            if (e == cudaSuccess) pw->direct_hash_slots_dev = d_hash.size();
        }
        if (e == cudaSuccess) e = cudaMemcpy(pw->direct_bitmap_dev, H.direct_bitmap.data(), (size_t)kDirectBitmapWords * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(pw->direct_hash_dev, d_hash.data(), d_hash.size() * sizeof(uint2), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            tfbs_set_error("tfbs_scan_hits: uploading the direct-path tables failed: %s", cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? TFBS_ERR_NOMEM : TFBS_ERR_CUDA;
        }
        pw->direct_hash_mask = (uint32_t)(d_hash.size() - 1);
        pw->direct_hash_shift = H.direct_shift;
    }
    pw->hq_thr.assign(thr, thr + Mall);
    pw->hq_valid = true;
    return TFBS_OK;
}

extern "C" int tfbs_scan_hits(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_pwms *pwms_c, const float *thr,
                              tfbs_hit *out_host, int64_t cap, int32_t sorted, int64_t *n_hits)
{
    TFBS_REQUIRE(ctx && seq && pwms_c && thr && n_hits && cap >= 0, "tfbs_scan_hits: bad argument");
    TFBS_REQUIRE(seq->ctx == ctx && pwms_c->ctx == ctx, "tfbs_scan_hits: handle from another context");
    tfbs_pwms *pwms = const_cast<tfbs_pwms *>(pwms_c);  // the threshold-specific tables are a cache
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    *n_hits = 0;
    const int64_t n = seq->n;
    if (n == 0) return TFBS_OK;
    const int M = pwms->M;
    int rc = tfbs_scratch_reserve(ctx->hits, (size_t)std::max<int64_t>(cap, 1) * sizeof(tfbs_hit));
    if (rc) return rc;
    rc = tfbs_scratch_reserve(ctx->hit_count, 64);
    if (rc) return rc;
    rc = tfbs_scratch_reserve(ctx->thr, (size_t)M * sizeof(float));
    if (rc) return rc;
    rc = tfbs_build_hits_tables(pwms, thr);
    if (rc) return rc;
    TFBS_REQUIRE(M < (1 << kCandMotifBits) && n < ((int64_t)1 << (63 - kCandMotifBits)), "tfbs_scan_hits: library or sequence too large");
    TFBS_CHECK_CUDA(cudaMemcpyAsync(ctx->thr.ptr, thr, (size_t)M * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemsetAsync(ctx->hit_count.ptr, 0, 64, ctx->stream));

    HitsParams P;
    rc = reserve_candidates(ctx, cap, &P.cand_cap);
    if (rc) return rc;
    P.cand = (unsigned long long *)ctx->cand.ptr;
    P.packed = seq->packed;
    P.mask = seq->mask;
    P.n_tiles = (n + kTilePos - 1) / kTilePos;
    P.tile0 = 0;
    P.n_blocks = pwms->n_blocks;
    P.qlut = pwms->hq_lut_dev;
    P.lut_bytes = (uint32_t)pwms->hits_lut_bytes;
    P.blk_lut_off = pwms->blk_lut_off_dev;
    P.blk_AB = pwms->blk_AB_dev;
    P.kinit = pwms->hq_kinit_dev;
    P.blk_motif = pwms->blk_motif_dev;
    P.lens = pwms->lens_dev;
    P.thr = (const float *)ctx->thr.ptr;
    P.exact = pwms->exact;
    P.exact2 = pwms->exact2;
    P.guard_eps = pwms->guard_eps;
    P.hits = (tfbs_hit *)ctx->hits.ptr;
    P.hit_count = (unsigned long long *)ctx->hit_count.ptr;
    P.cap = (unsigned long long)cap;
    P.work_counter = (unsigned int *)((char *)ctx->hit_count.ptr + 24);
    P.owned_end = n;
    P.pos_offset = 0;
    TFBS_REQUIRE(P.n_tiles * P.n_blocks < (int64_t)0xFFFF0000ll, "tfbs_scan_hits: too many work items");  // + one prefetch per CTA
    TFBS_REQUIRE((P.n_tiles * kTilePos + 128) <= seq->n_words * 16, "tfbs_scan_hits: sequence padding too small");

    size_t smem = 0;
    rc = hits_smem_plan(ctx, pwms->hits_lut_bytes, &smem);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaFuncSetAttribute(scan_hits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = std::min<int64_t>((int64_t)ctx->sm_count, P.n_tiles * P.n_blocks);
    tfbs_begin(ctx);
    launch_direct(ctx, pwms, P, 0, n);   // the shortest motifs, when their hit L-mers are few enough to enumerate
    cudaEventRecord(ctx->ev_direct, ctx->stream);
    if (grid > 0) {
        scan_hits_kernel<<<(unsigned)grid, kThreads, smem, ctx->stream>>>(P, nullptr);
        tfbs_launched(ctx);
    }
    cudaEventRecord(ctx->ev_mid, ctx->stream);
    launch_verify(ctx, P);
    rc = tfbs_end(ctx);
    if (rc) return rc;
    cudaEventElapsedTime(&ctx->last_scan_ms, ctx->ev0, ctx->ev_mid);
    cudaEventElapsedTime(&ctx->last_direct_ms, ctx->ev0, ctx->ev_direct);
    cudaEventElapsedTime(&ctx->last_verify_ms, ctx->ev_mid, ctx->ev1);
    unsigned long long cnt2[3] = {0, 0, 0};
    TFBS_CHECK_CUDA(cudaMemcpyAsync(cnt2, ctx->hit_count.ptr, sizeof(cnt2), cudaMemcpyDeviceToHost, ctx->stream));
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    TFBS_REQUIRE(cnt2[2] == 0, "tfbs_scan_hits: no kernel instantiation for column grouping A=%u B=%u",
                 (unsigned)((cnt2[2] >> 8) & 255), (unsigned)(cnt2[2] & 255));
    const unsigned long long cnt = cnt2[0];
    *n_hits = (int64_t)cnt;
    ctx->last_candidates = (int64_t)cnt2[1];
    const int64_t take = std::min<int64_t>((int64_t)cnt, cap);
    if (sorted && take > 1) {
        rc = sort_hits_device(ctx, take, M);  // also orders the device-resident copy (out_host == NULL)
        if (rc) return rc;
    }
    if (out_host && take > 0) {
        TFBS_CHECK_CUDA(cudaMemcpyAsync(out_host, ctx->hits.ptr, (size_t)take * sizeof(tfbs_hit),
                                        cudaMemcpyDeviceToHost, ctx->stream));
    }
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((int64_t)cnt > cap) {
        tfbs_set_error("tfbs_scan_hits: %llu hits exceed capacity %lld", cnt, (long long)cap);
        return TFBS_ERR_OVERFLOW;
    }
    return TFBS_OK;
}


// End-to-end entry point for a HOST-resident sequence: the buffer is cut into `chunks` ranges on
// tile boundaries; H2D copies (stream_in), encode + scan (compute stream) and the D2H copies of
// every chunk's hits (stream_out) overlap, so the PCIe time hides behind the scan.  Results are the
// same as tfbs_encode + tfbs_scan_hits (hit order is unspecified unless `sorted`).
//
// Shard form (tfbs_scan_hits_host_shard): the buffer is one shard [shard_start, shard_start + n) of a
// longer sequence, scanned with its right halo; only windows starting in its first `owned` bases are
// reported (the halo's windows belong to the next shard) and `shard_start` is added to every position,
// both on the device, so per-shard results concatenate without a host-side filter.
// With `sorted` the hits are ordered on the device (radix sort) after the last chunk and copied back in
// one piece; without it every chunk's hits are copied while the next chunk is scanned.
static int scan_hits_host_impl(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq *seq,
                               const tfbs_pwms *pwms_c, const float *thr, tfbs_hit *out_host, int64_t cap,
                               int32_t chunks, int32_t sorted, int64_t owned, int64_t shard_start, int64_t *n_hits)
{
    TFBS_REQUIRE(ctx && ascii_host && seq && pwms_c && thr && out_host && n_hits && cap >= 0 && n > 0,
                 "tfbs_scan_hits_host: bad argument");
    TFBS_REQUIRE(owned >= 0 && owned <= n && shard_start >= 0, "tfbs_scan_hits_host: bad shard bounds");
    TFBS_REQUIRE(seq->ctx == ctx && pwms_c->ctx == ctx && seq->n == n, "tfbs_scan_hits_host: handle mismatch");
    tfbs_pwms *pwms = const_cast<tfbs_pwms *>(pwms_c);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    *n_hits = 0;
    const int M = pwms->M;
    int rc = tfbs_scratch_reserve(ctx->hits, (size_t)std::max<int64_t>(cap, 1) * sizeof(tfbs_hit));
    if (rc) return rc;
    rc = tfbs_scratch_reserve(ctx->hit_count, 64);
    if (rc) return rc;
    rc = tfbs_scratch_reserve(ctx->thr, (size_t)M * sizeof(float));
    if (rc) return rc;
    rc = tfbs_build_hits_tables(pwms, thr);
    if (rc) return rc;
    TFBS_REQUIRE(M < (1 << kCandMotifBits) && n < ((int64_t)1 << (63 - kCandMotifBits)), "tfbs_scan_hits_host: library or sequence too large");
    if (seq->staging_bytes < (size_t)n) {
        if (seq->staging) cudaFree(seq->staging);
        seq->staging = nullptr;
        seq->staging_bytes = 0;
        TFBS_CHECK_CUDA(cudaMalloc(&seq->staging, (size_t)n + 256));
        seq->staging_bytes = (size_t)n;
    }
    const int64_t total_tiles = (n + kTilePos - 1) / kTilePos;
    TFBS_REQUIRE((total_tiles * kTilePos + 128) <= seq->n_words * 16, "tfbs_scan_hits_host: sequence padding too small");
    int64_t C = std::max<int64_t>(1, std::min<int64_t>(chunks, std::min<int64_t>(total_tiles, 64)));
    // Chunk boundaries (in tiles).  Only two transfers are exposed: the H2D copy of the first chunk and the
    // D2H copy of the last chunk's hits, so with 4 or more chunks those two are made small (with 6 or more:
    // 0.15 and 0.12 of a middle chunk, their neighbours 0.5) while the middle chunks keep the launches long.
    std::vector<int64_t> t_edge(1, 0);
    {
        std::vector<double> w((size_t)C, 1.0);
        if (C >= 6 && total_tiles >= 40 * C) {
            w[0] = 0.15;
            w[1] = 0.5;
            w[(size_t)C - 2] = 0.5;
            w[(size_t)C - 1] = 0.12;
        } else if (C >= 4 && total_tiles >= 8 * C) {
            w[0] = 0.4;
            w[(size_t)C - 2] = 0.6;
            w[(size_t)C - 1] = 0.25;
        }
        double sum = 0.0, acc = 0.0;
        for (double v : w) sum += v;
        for (int64_t c = 0; c < C; c++) {
            acc += w[(size_t)c];
            const int64_t e = c == C - 1 ? total_tiles : std::min<int64_t>(total_tiles, (int64_t)std::llround(acc / sum * (double)total_tiles));
            if (e > t_edge.back()) t_edge.push_back(e);
        }
        if (t_edge.back() != total_tiles) t_edge.push_back(total_tiles);
        C = (int64_t)t_edge.size() - 1;
    }
    while ((int64_t)ctx->pipe_events.size() < 2 * C + 2) {
        cudaEvent_t ev;
        TFBS_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->pipe_events.push_back(ev);
    }
    cudaEvent_t ev_start = ctx->pipe_events[2 * C], ev_out = ctx->pipe_events[2 * C + 1];
    const int64_t n_pad_total = seq->n_words * 16;

    TFBS_CHECK_CUDA(cudaMemcpyAsync(ctx->thr.ptr, thr, (size_t)M * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemsetAsync(ctx->hit_count.ptr, 0, 64, ctx->stream));
    // copy streams start after everything already queued on the compute stream (and the timer)
    TFBS_CHECK_CUDA(cudaEventRecord(ev_start, ctx->stream));
    TFBS_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream_in, ev_start, 0));
    TFBS_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream_out, ev_start, 0));

    HitsParams P;
    rc = reserve_candidates(ctx, cap, &P.cand_cap);
    if (rc) return rc;
    P.cand = (unsigned long long *)ctx->cand.ptr;
    P.packed = seq->packed;
    P.mask = seq->mask;
    P.n_blocks = pwms->n_blocks;
    P.qlut = pwms->hq_lut_dev;
    P.lut_bytes = (uint32_t)pwms->hits_lut_bytes;
    P.blk_lut_off = pwms->blk_lut_off_dev;
    P.blk_AB = pwms->blk_AB_dev;
    P.kinit = pwms->hq_kinit_dev;
    P.blk_motif = pwms->blk_motif_dev;
    P.lens = pwms->lens_dev;
    P.thr = (const float *)ctx->thr.ptr;
    P.exact = pwms->exact;
    P.exact2 = pwms->exact2;
    P.guard_eps = pwms->guard_eps;
    P.hits = (tfbs_hit *)ctx->hits.ptr;
    P.hit_count = (unsigned long long *)ctx->hit_count.ptr;
    P.cap = (unsigned long long)cap;
    P.owned_end = owned;
    P.pos_offset = shard_start;
    size_t smem = 0;
    rc = hits_smem_plan(ctx, pwms->hits_lut_bytes, &smem);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaFuncSetAttribute(scan_hits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    tfbs_begin(ctx);
    for (int64_t c = 0; c < C; c++) {
        const int64_t t0 = t_edge[(size_t)c], t1 = t_edge[(size_t)c + 1];
        const int64_t b0 = t0 * kTilePos;
        // bytes this chunk's scan can touch: its tiles + the 128-base halo (re-copied by the next chunk)
        const int64_t b1 = std::min<int64_t>(n, t1 * kTilePos + 128);
        const int64_t cover = (c == C - 1) ? (n_pad_total - b0) : (((t1 * kTilePos + 128 - b0) + 2047) / 2048 * 2048);
        TFBS_CHECK_CUDA(cudaMemcpyAsync((char *)seq->staging + b0, ascii_host + b0, (size_t)(b1 - b0),
                                        cudaMemcpyHostToDevice, ctx->stream_in));
        TFBS_CHECK_CUDA(cudaEventRecord(ctx->pipe_events[2 * c], ctx->stream_in));
        TFBS_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pipe_events[2 * c], 0));
        tfbs_encode_range(ctx, (char *)seq->staging + b0, b1 - b0, seq, b0, cover, ctx->stream);
        P.tile0 = t0;
        P.n_tiles = t1 - t0;
        P.work_counter = (unsigned int *)((char *)ctx->hit_count.ptr + 24);
        // work counter (offset 24) and candidate-list length (offset 32) restart for every chunk
        TFBS_CHECK_CUDA(cudaMemsetAsync(P.work_counter, 0, 16, ctx->stream));
        const int64_t grid = std::min<int64_t>((int64_t)ctx->sm_count, P.n_tiles * P.n_blocks);
        launch_direct(ctx, pwms, P, b0, std::min<int64_t>(n, t1 * kTilePos));
        if (grid > 0) {
            scan_hits_kernel<<<(unsigned)grid, kThreads, smem, ctx->stream>>>(P, nullptr);
            tfbs_launched(ctx);
        }
        launch_verify(ctx, P);
        TFBS_CHECK_CUDA(cudaMemcpyAsync(ctx->pinned_counts + c, ctx->hit_count.ptr, sizeof(unsigned long long),
                                        cudaMemcpyDeviceToHost, ctx->stream));
        TFBS_CHECK_CUDA(cudaEventRecord(ctx->pipe_events[2 * c + 1], ctx->stream));
    }
    unsigned long long done = 0;
    if (sorted) {
        // ordered output: radix sort on the device once every chunk is scanned, then one copy
        TFBS_CHECK_CUDA(cudaEventSynchronize(ctx->pipe_events[2 * (C - 1) + 1]));
        done = std::min<unsigned long long>(ctx->pinned_counts[C - 1], (unsigned long long)cap);
        if (done > 1) {
            rc = sort_hits_device(ctx, (int64_t)done, M);
            if (rc) return rc;
        }
        if (done > 0)
            TFBS_CHECK_CUDA(cudaMemcpyAsync(out_host, ctx->hits.ptr, (size_t)done * sizeof(tfbs_hit), cudaMemcpyDeviceToHost,
                                            ctx->stream));
    } else {
        // drain: as soon as a chunk's scan is done its hits go home while the next chunk is scanned
        for (int64_t c = 0; c < C; c++) {
            TFBS_CHECK_CUDA(cudaEventSynchronize(ctx->pipe_events[2 * c + 1]));
            const unsigned long long upto = std::min<unsigned long long>(ctx->pinned_counts[c], (unsigned long long)cap);
            if (upto > done) {
                TFBS_CHECK_CUDA(cudaMemcpyAsync(out_host + done, (tfbs_hit *)ctx->hits.ptr + done,
                                                (size_t)(upto - done) * sizeof(tfbs_hit), cudaMemcpyDeviceToHost,
                                                ctx->stream_out));
                done = upto;
            }
        }
        TFBS_CHECK_CUDA(cudaEventRecord(ev_out, ctx->stream_out));
        TFBS_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ev_out, 0));  // a region timer on the compute stream sees the copies
    }
    rc = tfbs_end(ctx);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream_out));
    {
        unsigned long long flag = 0;
        TFBS_CHECK_CUDA(cudaMemcpy(&flag, (char *)ctx->hit_count.ptr + 16, sizeof(flag), cudaMemcpyDeviceToHost));
        TFBS_REQUIRE(flag == 0, "tfbs_scan_hits_host: no kernel instantiation for column grouping A=%u B=%u",
                     (unsigned)((flag >> 8) & 255), (unsigned)(flag & 255));
        unsigned long long cand = 0;
        TFBS_CHECK_CUDA(cudaMemcpy(&cand, (char *)ctx->hit_count.ptr + 8, sizeof(cand), cudaMemcpyDeviceToHost));
        ctx->last_candidates = (int64_t)cand;
    }
    const unsigned long long cnt = ctx->pinned_counts[C - 1];
    *n_hits = (int64_t)cnt;
    if ((int64_t)cnt > cap) {
        tfbs_set_error("tfbs_scan_hits_host: %llu hits exceed capacity %lld", cnt, (long long)cap);
        return TFBS_ERR_OVERFLOW;
    }
    return TFBS_OK;
}


extern "C" int tfbs_scan_hits_host(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq *seq,
                                   const tfbs_pwms *pwms, const float *thr, tfbs_hit *out_host, int64_t cap,
                                   int32_t chunks, int32_t sorted, int64_t *n_hits)
{
    return scan_hits_host_impl(ctx, ascii_host, n, seq, pwms, thr, out_host, cap, chunks, sorted, n, 0, n_hits);
}

extern "C" int tfbs_scan_hits_host_shard(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq *seq,
                                         const tfbs_pwms *pwms, const float *thr, tfbs_hit *out_host, int64_t cap,
                                         int32_t chunks, int32_t sorted, int64_t owned, int64_t shard_start,
                                         int64_t *n_hits)
{
    return scan_hits_host_impl(ctx, ascii_host, n, seq, pwms, thr, out_host, cap, chunks, sorted, owned, shard_start, n_hits);
}

// Host-only library for the test hooks below: the fields the planning code reads, no device memory.
static int host_library(tfbs_pwms &pw, const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax)
{
    pw.M = M;
    pw.Lmax = Lmax;
    pw.lens.assign(lens, lens + M);
    pw.exact_host.assign((size_t)M * 2 * TFBS_MAX_MOTIF_LEN * 4, 0.0);
    for (int m = 0; m < M; m++) {
        if (lens[m] < 1 || lens[m] > Lmax) return TFBS_ERR_INVALID;
        tfbs_fill_exact(logodds, Lmax, m, lens[m], &pw.exact_host[((size_t)m * 2 + 0) * TFBS_MAX_MOTIF_LEN * 4],
                        &pw.exact_host[((size_t)m * 2 + 1) * TFBS_MAX_MOTIF_LEN * 4]);
    }
    tfbs_hits_order_and_capacity(&pw, lens);
    return TFBS_OK;
}

// Test hook (host only, no CUDA): run the candidate FILTER of the hits kernel on the host, with the
// kernel's own tables (fill_block) and its 32-bit field arithmetic, so the CPU test-suite can check
// that {candidates} is a superset of {hits} for every block form — in particular that fields never
// carry (wide, narrow) and that the carry slack of sticky fields is sufficient.
//   form 0: wide blocks only; 1: narrow wherever more than 32 motifs are left; 2: as 1, but sticky
//   from TFBS_STICKY_MIN_GROUPS groups on; 3: what a scan would use (cost model / environment) with every
//   motif on the tables; 4: the table side of the plan a scan really uses (the motifs on the direct path
//   are left out and get no candidates here).
//   plan (optional): int32[blocks][4] = {A, B, form, motifs}, n_blocks; expected (optional):
//   double[blocks] = candidates per window start the builder expects of a narrow / sticky block.
//   codes: n bases as 2-bit codes, one per byte (bases past the end count as 0, as in the padded device buffer)
//   cand:  uint8[n][M][2], set to 1 where the filter passes window start p of motif m on strand s
extern "C" int tfbs_debug_hits_filter(const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax,
                                      const float *thr, int32_t form, const uint8_t *codes, int64_t n, uint8_t *cand,
                                      int32_t *plan, double *expected, int32_t *n_blocks)
{
    if (!logodds || !lens || !thr || !codes || !cand || M < 1 || Lmax < 1 || Lmax > TFBS_MAX_MOTIF_LEN || n < 1 ||
        form < 0 || form > 4)
        return TFBS_ERR_INVALID;
    tfbs_pwms pw;  // host fields only
    if (host_library(pw, logodds, lens, M, Lmax)) return TFBS_ERR_INVALID;
    memset(cand, 0, (size_t)n * M * 2);
    std::vector<uint32_t> lut((size_t)6 * 8192 + (size_t)kMaxG * 2048), kinit(32);
    auto base = [&](int64_t i) -> uint32_t { return i < n ? (codes[i] & 3u) : 0u; };
    // form 4: the table side of the plan a scan really uses — the motifs the direct path took are left out
    std::vector<int32_t> t_order = pw.order, t_lens = pw.lens_desc;
    if (form == 4) {
        std::vector<uint32_t> bitmap;
        std::vector<uint2> hash;
        std::vector<uint8_t> is_direct;
        uint32_t shift = 0;
        int dlen = 0;
        int64_t entries = 0;
        plan_direct(&pw, thr, bitmap, hash, &shift, &dlen, &entries, is_direct);
        table_order(&pw, is_direct, t_order, t_lens);
    }
    int nb = 0;
    double exp_block = 0.0;
    cut_blocks(
        t_lens.data(), (int)t_order.size(),
        [&](int first, int count, int A, int B) {
            exp_block = 0.0;
            if (form >= 3) return choose_form(&pw, thr, t_order.data(), t_lens.data(), first, count, A, B, lut.data(), kinit.data(), &exp_block);
            return form == 0 ? (int)kWide : (form == 2 && A + B >= TFBS_STICKY_MIN_GROUPS ? (int)kSticky : (int)kNarrow);
        },
        [&](const tfbs_hits_block &K) {
            const int G = K.A + K.B;
            const int32_t *motifs = t_order.data() + K.first;
            const double e = fill_block(&pw, thr, motifs, K.count, K.A, K.B, K.narrow, lut.data(), kinit.data());
            if (plan) {
                plan[4 * nb + 0] = K.A;
                plan[4 * nb + 1] = K.B;
                plan[4 * nb + 2] = K.narrow;
                plan[4 * nb + 3] = K.count;
            }
            if (expected) expected[nb] = K.narrow == kWide ? 0.0 : e;
            nb++;
            const uint32_t PM = K.narrow ? kPoisonNarrow : kPoison;
            for (int64_t p = 0; p < n; p++)
                for (int lane = 0; lane < 32; lane++) {
                    uint32_t acc = 0;
                    for (int g = 0; g < G; g++) {
                        const int k = g < K.A ? 4 : 3;
                        const int col = g < K.A ? 4 * g : 4 * K.A + 3 * (g - K.A);
                        const size_t off = g < K.A ? (size_t)g * 8192 : (size_t)K.A * 8192 + (size_t)(g - K.A) * 2048;
                        uint32_t x = 0;
                        for (int j = 0; j < k; j++) x |= base(p + col + j) << (2 * j);
                        const uint32_t v = lut[off + (size_t)x * 32 + lane];
                        if (g == 0) acc = kinit[lane] + v;                         // the kernel's three
                        else if (K.narrow == kSticky) acc = (acc + v) | (acc & PM);  // accumulate forms
                        else acc += v;
                    }
                    const int nf = K.narrow ? 4 : 2;
                    for (int f = 0; f < nf; f++) {
                        const int flag = 31 - (32 / nf) * f;
                        if ((acc >> flag) & 1u) continue;
                        const int slot = lane + 32 * (K.narrow ? (f >> 1) : 0);
                        if (slot >= K.count) continue;
                        cand[((size_t)p * M + motifs[slot]) * 2 + (f & 1)] = 1;
                    }
                }
        });
    if (n_blocks) *n_blocks = nb;
    return TFBS_OK;
}

// Test hook (host only, no CUDA): plan the DIRECT path for a library and threshold vector and look a
// sequence up in its bitmap + hash exactly as scan_direct_kernel does.  codes = n bases as 2-bit codes, one
// per byte (bases past the end count as 0, like the padded device buffer); cand = uint8[n][M][2], set to 1
// where the 12-mer at position p is a key of (motif m, strand s).  info = {motifs on the direct path, their
// longest length, keys in the hash, bitmap bits set}.
extern "C" int tfbs_debug_direct_lookup(const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax,
                                        const float *thr, const uint8_t *codes, int64_t n, uint8_t *cand, int64_t info[4])
{
    if (!logodds || !lens || !thr || !codes || !cand || !info || M < 1 || Lmax < 1 || Lmax > TFBS_MAX_MOTIF_LEN || n < 1)
        return TFBS_ERR_INVALID;
    tfbs_pwms pw;  // host fields only
    if (host_library(pw, logodds, lens, M, Lmax)) return TFBS_ERR_INVALID;
    std::vector<uint32_t> bitmap;
    std::vector<uint2> hash;
    uint32_t shift = 0;
    int dlen = 0;
    int64_t entries = 0;
    std::vector<uint8_t> is_direct;
    const int nd = plan_direct(&pw, thr, bitmap, hash, &shift, &dlen, &entries, is_direct);
    memset(cand, 0, (size_t)n * M * 2);
    info[0] = nd;
    info[1] = dlen;
    info[2] = entries;
    info[3] = 0;
    if (nd == 0) return TFBS_OK;
    for (uint32_t w : bitmap) info[3] += __builtin_popcount(w);
    const uint32_t mask = (uint32_t)(hash.size() - 1);
    for (int64_t p = 0; p < n; p++) {
        uint32_t key = 0;
        for (int i = 0; i < kDirectKeyBases; i++) key |= (uint32_t)(p + i < n ? (codes[p + i] & 3u) : 0u) << (2 * i);
        const uint32_t k10 = key & 0xFFFFFu;
        if (!((bitmap[k10 >> 5] >> (k10 & 31u)) & 1u)) continue;
        for (uint32_t i = (key * 0x9E3779B1u) >> shift;; i = (i + 1u) & mask) {
            if (hash[i].y == 0u) break;
            if (hash[i].x == key) {
                const uint32_t v = hash[i].y - 1u;
                const int64_t start = p - (int64_t)(v & ((1u << kDirectOffBits) - 1u));
                if (start >= 0) cand[((size_t)start * M + (v >> (kDirectOffBits + 1))) * 2 + ((v >> kDirectOffBits) & 1u)] = 1;
            }
        }
    }
    return TFBS_OK;
}

// Test hook (host only, no CUDA): which motifs the direct path takes for a threshold vector (is_direct[M]);
// the choice is made per motif, so one loose threshold does not keep the rest of the library on the tables.
extern "C" int tfbs_debug_direct_motifs(const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax,
                                        const float *thr, uint8_t *is_direct)
{
    if (!logodds || !lens || !thr || !is_direct || M < 1 || Lmax < 1 || Lmax > TFBS_MAX_MOTIF_LEN) return TFBS_ERR_INVALID;
    tfbs_pwms pw;  // host fields only
    if (host_library(pw, logodds, lens, M, Lmax)) return TFBS_ERR_INVALID;
    std::vector<uint32_t> bitmap;
    std::vector<uint2> hash;
    std::vector<uint8_t> flags;
    uint32_t shift = 0;
    int dlen = 0;
    int64_t entries = 0;
    plan_direct(&pw, thr, bitmap, hash, &shift, &dlen, &entries, flags);
    std::copy(flags.begin(), flags.end(), is_direct);
    return TFBS_OK;
}

// Test hook (host only, no CUDA): the COMPLETE plan a scan makes for a threshold vector (plan_hits_tables:
// direct-path choice, speculative parallel block planning, the cut, every filter table) reduced to
// checksums, so that the CPU suite can pin the plan across refactorings and thread counts:
// out = {blocks, table words, FNV-1a of the filter tables + initial field values, FNV-1a of the block
// descriptors (A, B, form, count, table offset, motif slots), motifs on the direct path, direct keys}.
// plan (optional, room for M blocks): int32[blocks][4] = {A, B, form, motifs}.
extern "C" int tfbs_debug_hits_plan_checksum(const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax,
                                             const float *thr, uint64_t out[6], int32_t *plan)
{
    if (!logodds || !lens || !thr || !out || M < 1 || Lmax < 1 || Lmax > TFBS_MAX_MOTIF_LEN) return TFBS_ERR_INVALID;
    tfbs_pwms pw;  // host fields only
    if (host_library(pw, logodds, lens, M, Lmax)) return TFBS_ERR_INVALID;
    HitsPlan H;
    const int rc = plan_hits_tables(&pw, thr, H);
    if (rc) return rc;
    auto fnv = [](uint64_t h, const void *data, size_t bytes) {
        const uint8_t *p = static_cast<const uint8_t *>(data);
        for (size_t i = 0; i < bytes; i++) h = (h ^ p[i]) * 0x100000001B3ull;
        return h;
    };
    uint64_t ht = 0xCBF29CE484222325ull, hb = 0xCBF29CE484222325ull;
    ht = fnv(ht, H.qlut.data(), (size_t)H.words * 4);
    ht = fnv(ht, H.kinit.data(), (size_t)H.nb * 32 * 4);
    for (const std::vector<int32_t> *v : {&H.blk_A, &H.blk_B, &H.blk_narrow, &H.blk_count, &H.blk_lut_off, &H.blk_AB})
        hb = fnv(hb, v->data(), v->size() * 4);
    hb = fnv(hb, H.blk_motif.data(), (size_t)H.nb * TFBS_BLOCK_SLOTS * 4);
    out[0] = (uint64_t)H.nb;
    out[1] = (uint64_t)H.words;
    out[2] = ht;
    out[3] = hb;
    out[4] = (uint64_t)H.n_direct;
    out[5] = (uint64_t)H.direct_entries;
    for (int b = 0; plan && b < H.nb; b++) {
        plan[4 * b + 0] = H.blk_A[(size_t)b];
        plan[4 * b + 1] = H.blk_B[(size_t)b];
        plan[4 * b + 2] = H.blk_narrow[(size_t)b];
        plan[4 * b + 3] = H.blk_count[(size_t)b];
    }
    return TFBS_OK;
}

// Test hook (host only): the blocks a library of the given motif lengths (any order) is cut into when
// every block of at most `narrow_max_groups` column groups is taken in narrow form (the scan itself
// decides per block from the thresholds, see tfbs_pwms_hits_blocks).
// out[b] = {A, B, narrow, motifs in the block}; at most M blocks are written.
extern "C" int tfbs_debug_hits_block_plan(const int32_t *lens, int32_t M, int32_t narrow_max_groups, int32_t *out,
                                          int32_t *n_blocks)
{
    if (!lens || !out || !n_blocks || M < 1) return TFBS_ERR_INVALID;
    std::vector<int32_t> d(lens, lens + M);
    for (int32_t L : d)
        if (L < 1 || L > TFBS_MAX_MOTIF_LEN) return TFBS_ERR_INVALID;
    std::stable_sort(d.begin(), d.end(), [](int a, int b) { return a > b; });
    int nb = 0;
    cut_blocks(d.data(), M, [&](int, int, int A, int B) { return A + B <= narrow_max_groups ? kNarrow : kWide; },
               [&](const tfbs_hits_block &K) {
                   out[4 * nb + 0] = K.A;
                   out[4 * nb + 1] = K.B;
                   out[4 * nb + 2] = K.narrow;
                   out[4 * nb + 3] = K.count;
                   nb++;
               });
    *n_blocks = nb;
    return TFBS_OK;
}

// The blocks the last tfbs_scan_hits / tfbs_scan_hits_host call on this library used (chosen per
// threshold vector by the cost model): out[b] = {A, B, narrow, motifs}; `cap` = blocks `out` can hold.
extern "C" int tfbs_pwms_hits_blocks(const tfbs_pwms *pwms, int32_t *out, int32_t cap, int32_t *n_blocks)
{
    if (!pwms || !n_blocks) return TFBS_ERR_INVALID;
    *n_blocks = pwms->hq_valid ? pwms->n_blocks : 0;
    for (int b = 0; out && b < *n_blocks && b < cap; b++) {
        out[4 * b + 0] = pwms->blk_A[b];
        out[4 * b + 1] = pwms->blk_B[b];
        out[4 * b + 2] = pwms->blk_narrow[b];
        out[4 * b + 3] = pwms->blk_count[b];
    }
    return TFBS_OK;
}

// The direct path of the last scan of this library: how many motifs (the shortest ones) bypassed the table
// filter, the longest of them, and how many 12-mer keys their hash holds; all 0 when the path was not taken.
extern "C" int tfbs_pwms_hits_direct(const tfbs_pwms *pwms, int32_t *n_motifs, int32_t *max_len, int64_t *keys)
{
    if (!pwms || !n_motifs || !max_len || !keys) return TFBS_ERR_INVALID;
    const bool on = pwms->hq_valid && pwms->n_direct > 0;
    *n_motifs = on ? pwms->n_direct : 0;
    *max_len = on ? pwms->direct_len : 0;
    *keys = on ? pwms->direct_entries : 0;
    return TFBS_OK;
}

// Test hook (host only): the column grouping the hits kernel uses for a block whose longest motif
// has L columns.
extern "C" int tfbs_debug_hits_group_plan(int32_t L, int32_t *A, int32_t *B)
{
    if (!A || !B || L < 1 || L > TFBS_MAX_MOTIF_LEN) return TFBS_ERR_INVALID;
    int a = 0, b = 0;
    tfbs_hits_group_plan(L, &a, &b);
    *A = a;
    *B = b;
    return TFBS_OK;
}
