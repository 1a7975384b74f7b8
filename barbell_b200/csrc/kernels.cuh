// sm_100a kernels of the annotate hot path.
//
//  K1 flank_scan    reference searcher.rs:438 (overhang_searcher.search over the WHOLE read, both strands):
//                   bit-vector edit distance of the N-masked flank against every text position; emits every end
//                   position whose cost is <= k ("sub-threshold entries").  Text tiles arrive in shared memory by one
//                   TMA bulk copy per CTA; one lane owns one text chunk; match masks are staged once in shared memory.
//  K1f filter       the same search as a LOSSLESS pre-filter: only the flank's longest N-free run (<= 15 rows) of both
//                   strands is advanced per base (one 32-bit word, branch-free); the candidate runs are re-scored
//                   together with the second N-free run on the text tile still in shared memory, and
//  K1v verify       the surviving windows (+ the read ends, where the overhang rule applies) are verified with the
//                   exact full-length DP.  Default whenever 3k <= rows of the run.
//  K2a resolve      sassy's local-minimum reporting rule applied to the sorted entries.
//  K2b trace        traceback of every reported flank match -> text_start and the barcode text region
//                   (reference cigar_parse.rs:71-82, searcher.rs:442-456).
//  K3  barcode_rows reference searcher.rs:267-426: all barcodes of the group against the region (one warp per flank
//                   match, barcodes across lanes, the DP rows laid along the text: barcode_rows.cuh), fallback pass,
//                   traceback, Lodhi score, thresholds, row assembly.
//  K4  collapse     reference interval.rs:4-79, one thread per read.
//
// The arithmetic mirrors the CPU oracle's policies S1-S7 bit for bit (the oracle is test infrastructure: nothing here
// includes, links or calls it).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "../../include/barbell_b200.h"
#include "device_types.cuh"
#include "barcode_rows.cuh"

namespace bb {

// ---------------------------------------------------------------------------------------------------------------
// bit-vector column step (Myers 1999 / Hyyro 2003), NW 64-bit words, semi-global (top row = 0)
// ---------------------------------------------------------------------------------------------------------------
template <int NW>
struct Col {
    uint64_t pv[NW], mv[NW];
};

template <int NW>
__device__ __forceinline__ int col_step(Col<NW>& c, const uint64_t* __restrict__ eq, int last_bit) {
    if constexpr (NW == 1) {
        const uint64_t e = eq[0], pv = c.pv[0], mv = c.mv[0];
        const uint64_t xv = e | mv;
        const uint64_t xh = (((e & pv) + pv) ^ pv) | e;
        uint64_t ph = mv | ~(xh | pv);
        uint64_t mh = pv & xh;
        const int d = static_cast<int>((ph >> last_bit) & 1) - static_cast<int>((mh >> last_bit) & 1);
        ph <<= 1; mh <<= 1;
        c.pv[0] = mh | ~(xv | ph);
        c.mv[0] = ph & xv;
        return d;
    } else {
        static_assert(NW == 2, "flank patterns up to 128 characters");
        const uint64_t e0 = eq[0], e1 = eq[1], pv0 = c.pv[0], pv1 = c.pv[1], mv0 = c.mv[0], mv1 = c.mv[1];
        const uint64_t xv0 = e0 | mv0, xv1 = e1 | mv1;
        const uint64_t t0 = e0 & pv0, t1 = e1 & pv1;
        const uint64_t s0 = t0 + pv0;
        const uint64_t s1 = t1 + pv1 + (s0 < t0 ? 1ull : 0ull);
        const uint64_t xh0 = (s0 ^ pv0) | e0, xh1 = (s1 ^ pv1) | e1;
        uint64_t ph0 = mv0 | ~(xh0 | pv0), ph1 = mv1 | ~(xh1 | pv1);
        uint64_t mh0 = pv0 & xh0, mh1 = pv1 & xh1;
        const int d = static_cast<int>((ph1 >> last_bit) & 1) - static_cast<int>((mh1 >> last_bit) & 1);
        ph1 = (ph1 << 1) | (ph0 >> 63); ph0 <<= 1;
        mh1 = (mh1 << 1) | (mh0 >> 63); mh0 <<= 1;
        c.pv[0] = mh0 | ~(xv0 | ph0); c.pv[1] = mh1 | ~(xv1 | ph1);
        c.mv[0] = ph0 & xv0;          c.mv[1] = ph1 & xv1;
        return d;
    }
}

// D[i] of a column from its vertical deltas
template <int NW>
__device__ __forceinline__ int col_val(const uint64_t* pv, const uint64_t* mv, int i) {
    int v = 0;
#pragma unroll
    for (int b = 0; b < NW; b++) {
        const int r = i - 64 * b;
        if (r <= 0) break;
        const uint64_t msk = r >= 64 ? ~0ull : ((1ull << r) - 1ull);
        v += __popcll(pv[b] & msk) - __popcll(mv[b] & msk);
    }
    return v;
}

__device__ __forceinline__ uint64_t make_key(uint32_t read, int group, int strand, uint32_t pos, int cost) {
    return (static_cast<uint64_t>(read) << kKeyReadShift) | (static_cast<uint64_t>(group) << kKeyGroupShift) |
           (static_cast<uint64_t>(strand) << kKeyStrandShift) | (static_cast<uint64_t>(pos) << kKeyPosShift) |
           static_cast<uint64_t>(cost & 0xff);
}

// ---------------------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk copy (1-D) helpers -- SASS: SYNCS.*, UBLKCP
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BB_DONE_%=;\n"
        "bra BB_WAIT_%=;\n"
        "BB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// K1: flank scan
//
// Work unit = one CHUNK of kChunk consecutive bases of ONE read (chunks never straddle reads, so no lane ever switches
// reads; a tiny pre-pass turns the read lengths into a chunk index).  A CTA owns 256 consecutive chunks; their bytes are
// contiguous in the batch buffer, so the CTA's text (plus a warm-up halo on both sides) arrives in shared memory by ONE
// TMA bulk copy (cp.async.bulk -> mbarrier).  Lane t scans its chunk twice: ascending with the forward masks, descending
// with the complemented masks (= forward search of the reverse-complemented read, oracle policy S4), each time after
// `warm` >= m+k warm-up columns inside the same read (the bottom-row cost at j depends on text[j-(m+k), j) only).
// The pattern sits at the TOP of the bit-vector (last flank row = bit 63 of the last word); the unused low bits are
// wildcard rows with zero vertical deltas, which is exactly the semi-global top boundary, so the bottom-row delta of a
// column is the sign bit of Ph / Mh.  Per character: 1 LDS.U8 (text) + 1 LDS.64/128 (match mask) + ~21 ALU-pipe ops.
// Control flow is uniform across the warp: groups of kGroup characters run branch-free and only track the minimum
// score; a group whose minimum drops to <= k is replayed character by character to emit the sub-threshold positions.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kGroup = 20;           // characters per unrolled group; chunk and warm-up are multiples of it
constexpr int kChunk = 300;          // bases per lane: 15 groups; 300/4 = 75 is odd (LDS.U8 of a warp is conflict-free)

struct ScanArgs {
    const uint8_t* bases;        // concatenated read bytes, 16-byte aligned
    const uint64_t* offsets;     // n_reads + 1
    const uint32_t* chunk_base;  // n_reads + 1: exclusive prefix sum of ceil(len / kChunk)
    const uint32_t* tile_first;  // per CTA: read owning the CTA's first chunk
    const uint64_t* tile_span;   // per CTA: [2t] first / [2t+1] one-past-last byte of its chunks in `bases` (no halo)
    uint32_t n_reads;
    uint64_t total16;            // readable bytes of `bases` (total rounded up to 16)
    int group;
    int strand_xor;              // policy S6: 1 = the key's strand bit is inverted so that Rc matches sort before forward ones
    uint64_t* entries;           // global entry list (sorted path) ...
    uint32_t* n_entries;
    uint32_t cap;
    uint64_t* slots;             // ... or, when non-null, per-read entry slots [n_reads][slot_cap] + per-read counts (slot path)
    uint32_t* slot_cnt;
    uint32_t slot_cap, slot_stride;   // slots usable per read (<= stride; a test knob lowers it) / distance between two reads' slots
    uint32_t* slot_overflow;     // set when a read has more entries than slots (the batch is then re-run on the sorted path)
};

// chunks per read (thread per read; entry n_reads is 0 so the exclusive scan yields the total)
__global__ void k_chunk_count(const uint64_t* __restrict__ offsets, uint32_t n_reads, uint32_t* __restrict__ nch) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    nch[r] = r < n_reads ? static_cast<uint32_t>((offsets[r + 1] - offsets[r] + kChunk - 1) / kChunk) : 0u;
}

// largest r in [lo, n_reads] with chunk_base[r] <= c and the read non-empty-for-c (upper_bound - 1)
__device__ __forceinline__ uint32_t find_chunk_read(const uint32_t* __restrict__ chunk_base, uint32_t n_reads, uint32_t lo, uint32_t c) {
    // gallop from lo, then bisect: first index with chunk_base[idx] > c
    uint32_t step = 1, hi = lo + 1;
    while (hi <= n_reads && __ldg(chunk_base + hi) <= c) { lo = hi; step <<= 1; hi = lo + step; }
    if (hi > n_reads + 1) hi = n_reads + 1;
    uint32_t a = lo + 1, b = hi;
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (__ldg(chunk_base + mid) <= c) a = mid + 1; else b = mid;
    }
    return a - 1;
}

// per CTA: the read that owns chunk cta*256 and the byte span of the CTA's chunks (thread per CTA), so that the scan kernels can
// issue their TMA copy before any chunk lookup
__global__ void k_tile_index(const uint32_t* __restrict__ chunk_base, const uint64_t* __restrict__ offsets, uint32_t n_reads, uint32_t n_tiles,
                             uint32_t* __restrict__ tile_first, uint64_t* __restrict__ tile_span) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint32_t total_chunks = chunk_base[n_reads];
    auto owner = [&](uint32_t c) {                      // read owning chunk c: last index with chunk_base[idx] <= c
        uint32_t a = 0, b = n_reads + 1;
        while (a < b) {
            const uint32_t mid = (a + b) >> 1;
            if (chunk_base[mid] <= c) a = mid + 1; else b = mid;
        }
        return a - 1;
    };
    const uint32_t c_first = t * kScanThreads;
    const uint32_t r_first = owner(c_first);
    tile_first[t] = r_first;
    if (c_first >= total_chunks) { tile_span[2 * t] = 0; tile_span[2 * t + 1] = 0; return; }
    const uint32_t c_last = min(c_first + kScanThreads, total_chunks) - 1;
    const uint32_t r_last = owner(c_last);
    const uint64_t g_lo = offsets[r_first] + static_cast<uint64_t>(c_first - chunk_base[r_first]) * kChunk;
    uint64_t g_hi = offsets[r_last] + static_cast<uint64_t>(c_last - chunk_base[r_last] + 1) * kChunk;
    if (g_hi > offsets[r_last + 1]) g_hi = offsets[r_last + 1];
    tile_span[2 * t] = g_lo; tile_span[2 * t + 1] = g_hi;
}

// top-aligned column step: returns the bottom-row delta (+1 / 0 / -1)
template <int NW>
__device__ __forceinline__ int col_step_top(Col<NW>& c, const uint64_t* __restrict__ eq) {
    if constexpr (NW == 1) {
        const uint64_t e = eq[0], pv = c.pv[0], mv = c.mv[0];
        const uint64_t sum = (e & pv) + pv;
        uint64_t ph = mv | ~(sum | pv | e);
        uint64_t mh = pv & ((sum ^ pv) | e);
        const int d = static_cast<int>(ph >> 63) - static_cast<int>(mh >> 63);
        ph <<= 1; mh <<= 1;
        c.pv[0] = mh | ~(e | mv | ph);
        c.mv[0] = ph & (e | mv);
        return d;
    } else {
        const uint64_t e0 = eq[0], e1 = eq[1], pv0 = c.pv[0], pv1 = c.pv[1], mv0 = c.mv[0], mv1 = c.mv[1];
        const uint64_t t0 = e0 & pv0, t1 = e1 & pv1;
        const uint64_t s0 = t0 + pv0;
        const uint64_t s1 = t1 + pv1 + (s0 < t0 ? 1ull : 0ull);
        uint64_t ph0 = mv0 | ~(s0 | pv0 | e0), ph1 = mv1 | ~(s1 | pv1 | e1);
        uint64_t mh0 = pv0 & ((s0 ^ pv0) | e0), mh1 = pv1 & ((s1 ^ pv1) | e1);
        const int d = static_cast<int>(ph1 >> 63) - static_cast<int>(mh1 >> 63);
        ph1 = (ph1 << 1) | (ph0 >> 63); ph0 <<= 1;
        mh1 = (mh1 << 1) | (mh0 >> 63); mh0 <<= 1;
        c.pv[0] = mh0 | ~(e0 | mv0 | ph0); c.pv[1] = mh1 | ~(e1 | mv1 | ph1);
        c.mv[0] = ph0 & (e0 | mv0);        c.mv[1] = ph1 & (e1 | mv1);
        return d;
    }
}

__device__ __forceinline__ void scan_emit(const ScanArgs& A, uint32_t r, int strand, uint32_t pos, int cost) {
    const uint64_t key = make_key(r, A.group, strand ^ A.strand_xor, pos, cost);
    if (A.slots) {
        const uint32_t idx = atomicAdd(A.slot_cnt + r, 1u);
        if (idx < A.slot_cap) A.slots[static_cast<size_t>(r) * A.slot_stride + idx] = key;
        else atomicExch(A.slot_overflow, 1u);
        return;
    }
    const uint32_t idx = atomicAdd(A.n_entries, 1u);
    if (idx < A.cap) A.entries[idx] = key;
}

template <int NW>
__global__ void __launch_bounds__(kScanThreads, 2) k_flank_scan(const ScanArgs A, const DevGroup G) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    int64_t* s_origin = reinterpret_cast<int64_t*>(smem + 16);
    uint64_t* s_eq = reinterpret_cast<uint64_t*>(smem + 128);                       // [2][256][NW]
    unsigned char* s_text = smem + 128 + 2 * 256 * NW * sizeof(uint64_t);           // warm-up + 256 chunks + warm-up

    const int tid = threadIdx.x;
    const int W = G.warm;                                                           // multiple of kGroup, >= m + k
    const uint32_t total_chunks = __ldg(A.chunk_base + A.n_reads);
    const uint32_t c_first = blockIdx.x * kScanThreads;
    if (c_first >= total_chunks) return;
    const uint32_t r_first = __ldg(A.tile_first + blockIdx.x);

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint64_t g_lo = __ldg(A.tile_span + 2 * blockIdx.x), g_hi = __ldg(A.tile_span + 2 * blockIdx.x + 1);
        uint64_t lo = g_lo >= static_cast<uint64_t>(W) ? g_lo - W : 0;
        lo &= ~15ull;
        uint64_t hi = (g_hi + W + 15) & ~15ull;
        if (hi > A.total16) hi = A.total16;
        *s_origin = static_cast<int64_t>(lo);
        const uint32_t bytes = static_cast<uint32_t>(hi - lo);
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(s_text, A.bases + lo, bytes, bar);
    }
    // stage the match masks while the bulk copy is in flight
    for (int i = tid; i < 2 * 256 * NW; i += kScanThreads) s_eq[i] = __ldg(G.eq_top + i);

    // this lane's chunk
    const uint32_t c = c_first + tid;
    const bool active = c < total_chunks;
    uint32_t r = 0;
    int n = 0, a = 0, b = 0;
    uint64_t rs_g = 0;
    if (active) {
        r = find_chunk_read(A.chunk_base, A.n_reads, r_first, c);
        rs_g = __ldg(A.offsets + r);
        n = static_cast<int>(__ldg(A.offsets + r + 1) - rs_g);
        a = static_cast<int>(c - __ldg(A.chunk_base + r)) * kChunk;
        b = min(a + kChunk, n);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    if (!active) return;
    const unsigned char* text = s_text + (static_cast<int64_t>(rs_g) - *s_origin);  // text[x] = base x of read r
    const int m = G.m, k = G.k, shift = 64 * NW - G.m;
    const int* __restrict__ ov = G.ov;

    // ---------------- forward strand: ascending text; reports end positions x+1 for x in [a, b) ----------------
    {
        const uint64_t* eqp = s_eq;
        Col<NW> col;
        const bool fresh = a - W > 0;                       // warm-up starts inside the read
#pragma unroll
        for (int w = 0; w < NW; w++) { col.pv[w] = fresh ? G.pv_plain_top[w] : G.pv_over_top[w]; col.mv[w] = 0; }
        int score = fresh ? m : G.ov_m;
        if (a == 0 && score <= k) scan_emit(A, r, BB_FWD, 0u, score);
#pragma unroll 1
        for (int x = a - W; x < b; x += kGroup) {
            if (x < 0) continue;
            int redo_to = x;                                 // characters [x, redo_to) need the per-character path
            if (x + kGroup <= b) {
                const Col<NW> saved = col;
                const int saved_score = score;
                int mn = 1 << 20;
#pragma unroll
                for (int t = 0; t < kGroup; t++) {
                    score += col_step_top<NW>(col, eqp + static_cast<uint32_t>(text[x + t]) * NW);
                    mn = min(mn, score);
                }
                if (mn <= k && x + kGroup > a) { col = saved; score = saved_score; redo_to = x + kGroup; }
            } else {
                redo_to = b;
            }
#pragma unroll 1
            for (int xx = x; xx < redo_to; xx++) {
                score += col_step_top<NW>(col, eqp + static_cast<uint32_t>(text[xx]) * NW);
                if (score <= k && xx >= a) scan_emit(A, r, BB_FWD, static_cast<uint32_t>(xx + 1), score);
            }
        }
        if (b == n) {                                        // virtual end positions past the text end (oracle policy S3)
            for (int t = 1; t <= m; t++) {
                const int v = col_val<NW>(col.pv, col.mv, shift + m - t) + __ldg(ov + t);
                if (v <= k) scan_emit(A, r, BB_FWD, static_cast<uint32_t>(n + t), v);
            }
        }
    }
    // ---------------- reverse-complement strand: descending text, complemented masks; frame position n - x ----------------
    {
        const uint64_t* eqp = s_eq + 256 * NW;
        Col<NW> col;
        const bool fresh = b + W < n;                        // warm-up starts inside the read
#pragma unroll
        for (int w = 0; w < NW; w++) { col.pv[w] = fresh ? G.pv_plain_top[w] : G.pv_over_top[w]; col.mv[w] = 0; }
        int score = fresh ? m : G.ov_m;
        if (b == n && score <= k) scan_emit(A, r, BB_RC, 0u, score);
        // same group grid as the forward pass (multiples of kGroup from the read start), walked from the top down
#pragma unroll 1
        for (int x = a + kChunk + W - kGroup; x >= a; x -= kGroup) {
            if (x >= n) continue;
            int redo_from = x + kGroup;                      // characters [redo_from, x + kGroup) -> per-character path
            if (x + kGroup <= n) {
                const Col<NW> saved = col;
                const int saved_score = score;
                int mn = 1 << 20;
#pragma unroll
                for (int t = kGroup - 1; t >= 0; t--) {
                    score += col_step_top<NW>(col, eqp + static_cast<uint32_t>(text[x + t]) * NW);
                    mn = min(mn, score);
                }
                if (mn <= k && x < b) { col = saved; score = saved_score; redo_from = x; }
            } else {
                redo_from = x;
            }
            const int top = min(x + kGroup, n);
#pragma unroll 1
            for (int xx = top - 1; xx >= redo_from; xx--) {
                score += col_step_top<NW>(col, eqp + static_cast<uint32_t>(text[xx]) * NW);
                if (score <= k && xx < b) scan_emit(A, r, BB_RC, static_cast<uint32_t>(n - xx), score);
            }
        }
        if (a == 0) {
            for (int t = 1; t <= m; t++) {
                const int v = col_val<NW>(col.pv, col.mv, shift + m - t) + __ldg(ov + t);
                if (v <= k) scan_emit(A, r, BB_RC, static_cast<uint32_t>(n + t), v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K1f/K1v: lossless pre-filter + exact window verification (used instead of k_flank_scan when DevGroup::f_on)
//
// If the flank matches with <= k edits ending at j, then its N-free run Q = rows [q0, q0+q) matches with e <= k edits
// ending at some p, and the D = m-q0-q rows after Q are aligned to text (p, j] with <= k-e edits, so
// |j - p - D| <= k - e.  K1f therefore runs ONLY the q <= 15 rows of Q for the forward strand and the q rows of rc(Q)
// for the reverse-complement strand -- both blocks side by side in ONE 32-bit word, one ascending pass, ~19 ALU ops per
// base for both strands -- and every position where a block's cost drops to <= k becomes a candidate WINDOW of end
// positions that K1v verifies with the exact full-length bit-vector DP (same arithmetic as k_flank_scan, incl. warm-up
// and overhang).  Matches that touch a read end may have Q itself hanging over the end (overhang cost alpha < 1 per row,
// which the filter does not model), so the first and last m+k end positions of every read and strand (and the virtual
// positions past the end) are ALWAYS verified.  Windows may overlap: duplicates are removed after the sort.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRunQueue = 2048;      // candidate runs a CTA can stage in shared memory (~850 per CTA on random text)

struct FilterArgs {
    ScanArgs S;
    uint64_t* windows;           // global queue of windows to verify
    uint32_t* n_windows;
    uint32_t win_cap;
    uint32_t* overflow;          // set when the global window queue overflowed (host falls back to the exact scan)
    int halo_l, halo_r;          // text kept in shared memory before / after the CTA's chunks (pre-check of the second run)
};

__device__ __forceinline__ uint64_t make_window(uint32_t read, int strand, int lo, int len) {
    return (static_cast<uint64_t>(read) << kWinReadShift) | (static_cast<uint64_t>(strand) << kWinStrandShift) |
           (static_cast<uint64_t>(static_cast<uint32_t>(lo)) << kWinLoShift) | static_cast<uint64_t>(len);
}

constexpr int kFiltGroups = kChunk / kGroup;     // candidate-bitmap groups per chunk (15)
constexpr int kFiltStage = kFiltGroups * kScanThreads * 5 / 8;   // windows a CTA can stage (the bitmap area is reused)

// Shared-memory halos of k_flank_filter: the filter's own warm-up, and the reach of the pre-check (the second run S ends
// d rows after/before the first, is scanned over +-k positions and needs its own rows + k warm-up columns).
__host__ inline void filter_halos(const DevGroup& G, int& halo_l, int& halo_r) {
    const int W = ((G.f_q + G.k + kGroup - 1) / kGroup) * kGroup;
    const int d_f = (G.f_s0 + G.f_qs) - (G.f_q0 + G.f_q), d_r = G.f_q0 - G.f_s0;
    const int back = G.f_qs ? std::max(0, std::max(-d_f, -d_r)) + 2 * G.k + G.f_qs + 2 : 0;
    const int fwd = G.f_qs ? std::max(0, std::max(d_f, d_r)) + G.k + 1 : 0;
    halo_l = (std::max(W, back) + 15) & ~15;
    halo_r = (fwd + 15) & ~15;
}
__host__ inline size_t filter_smem_bytes(int halo_l, int halo_r) {
    return 128 + 2 * 1024 + kRunQueue * sizeof(uint32_t) + 3 * kScanThreads * sizeof(int32_t) +
           static_cast<size_t>(kFiltGroups) * kScanThreads * 5 + static_cast<size_t>(kScanThreads) * kChunk + halo_l + halo_r + 64;
}

// min over end positions [lo, hi] (1-based, <= n) of the semi-global cost of a <=15-row block against text[0, n)
__device__ __forceinline__ int block_min_cost(const uint8_t* __restrict__ text, const uint32_t* __restrict__ eq, int shiftbits, int rows,
                                              int k, int lo, int hi) {
    const uint32_t blk = (1u << rows) - 1u;
    int c0 = lo - 1 - (rows + k);                            // the cost at p depends on text[p-(rows+k), p) only
    if (c0 < 0) c0 = 0;
    uint32_t pv = blk, mv = 0;
    int score = rows, best = rows;
    for (int c = c0; c < hi; c++) {
        const uint32_t e = (eq[text[c]] >> shiftbits) & blk;
        const uint32_t sum = (e & pv) + pv;
        uint32_t ph = mv | ~(sum | pv | e);
        uint32_t mh = pv & ((sum ^ pv) | e);
        score += static_cast<int>((ph >> (rows - 1)) & 1u) - static_cast<int>((mh >> (rows - 1)) & 1u);
        ph <<= 1; mh <<= 1;
        pv = (mh | ~(e | mv | ph)) & blk;
        mv = ph & (e | mv) & blk;
        if (c + 1 >= lo) best = min(best, score);
    }
    return best;
}

// Phase 1 (scan): branch-free; per base it advances both 15-row blocks (one 32-bit word), updates the two packed costs and
// shifts the two "cost > k" flags into per-strand bit registers; after every group of 20 bases the 2 x 20 flags go to a
// per-lane bitmap in shared memory.
// Phase 2 (runs): each lane turns its bitmap into runs of consecutive candidate positions per strand (CTA queue).
// Phase 3 (pre-check, all threads over the CTA queue): a run [ps, pe] says "the run Q ends here with <= k edits".  In a real
// match the second N-free run S of the flank ends at pS with |pS - p - d| <= k (d = signed row distance of the two run ends)
// and cost(Q) + cost(S) <= k (disjoint rows of one alignment), so Q is re-scored over [ps, pe] and S over
// [ps + d - k, pe + d + k] with single 32-bit block DPs on the text ALREADY IN SHARED MEMORY, and the run is dropped when
// min cost(Q) + min cost(S) > k: ~95 % of the random candidates disappear before the ~4000-instruction exact verification
// without a second pass over HBM.  Survivors become windows of end positions (one global atomic per CTA):
//   forward: the Df rows after Q are aligned to (p, j] with <= k edits               -> j   in [ps + Df - k, pe + Df + k]
//   rc     : the match starts at s = p - Dr +- k (the run's own indels shift its end) -> n-s in [n - pe + Dr - k, n - ps + Dr + k]
__global__ void __launch_bounds__(kScanThreads, 2) k_flank_filter(const FilterArgs F, const DevGroup G) {
    extern __shared__ __align__(128) unsigned char smem[];
    const ScanArgs& A = F.S;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    int64_t* s_origin = reinterpret_cast<int64_t*>(smem + 16);
    uint32_t* s_qn = reinterpret_cast<uint32_t*>(smem + 32);
    uint32_t* s_wn = reinterpret_cast<uint32_t*>(smem + 36);
    uint32_t* s_wbase = reinterpret_cast<uint32_t*>(smem + 40);
    uint32_t* s_eq = reinterpret_cast<uint32_t*>(smem + 128);                       // [256] first run Q (both strands)
    uint32_t* s_seq = s_eq + 256;                                                   // [256] second run S
    uint32_t* s_runs = s_seq + 256;                                                 // [kRunQueue]
    int32_t* s_lane_r = reinterpret_cast<int32_t*>(s_runs + kRunQueue);             // per lane: read, chunk start, text offset of the read
    int32_t* s_lane_a = s_lane_r + kScanThreads;
    int32_t* s_lane_t = s_lane_a + kScanThreads;
    uint32_t* s_bm32 = reinterpret_cast<uint32_t*>(s_lane_t + kScanThreads);        // [group][lane]
    uint8_t* s_bm8 = reinterpret_cast<uint8_t*>(s_bm32 + kFiltGroups * kScanThreads);   // [group][lane]
    uint64_t* s_stage = reinterpret_cast<uint64_t*>(s_bm32);                        // phase 3: surviving windows (bitmaps are dead)
    unsigned char* s_text = s_bm8 + kFiltGroups * kScanThreads;

    const int tid = threadIdx.x;
    const int q = G.f_q, k = G.k, m = G.m;
    const int W = ((q + k + kGroup - 1) / kGroup) * kGroup;                         // warm-up columns of the filter
    const uint32_t total_chunks = __ldg(A.chunk_base + A.n_reads);
    const uint32_t c_first = blockIdx.x * kScanThreads;
    if (c_first >= total_chunks) return;
    const uint32_t r_first = __ldg(A.tile_first + blockIdx.x);

    if (tid == 0) {
        *s_qn = 0; *s_wn = 0;
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint64_t g_lo = __ldg(A.tile_span + 2 * blockIdx.x), g_hi = __ldg(A.tile_span + 2 * blockIdx.x + 1);
        uint64_t lo = g_lo >= static_cast<uint64_t>(F.halo_l) ? g_lo - F.halo_l : 0;
        lo &= ~15ull;
        uint64_t hi = (g_hi + F.halo_r + 15) & ~15ull;
        if (hi > A.total16) hi = A.total16;
        *s_origin = static_cast<int64_t>(lo);
        const uint32_t bytes = static_cast<uint32_t>(hi - lo);
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(s_text, A.bases + lo, bytes, bar);
    }
    for (int i = tid; i < 256; i += kScanThreads) { s_eq[i] = __ldg(G.f_eq + i); s_seq[i] = G.f_qs ? __ldg(G.f_seq + i) : 0u; }

    const uint32_t c = c_first + tid;
    const bool active = c < total_chunks;
    uint32_t r = 0;
    int n = 0, a = 0, b = 0;
    uint64_t rs_g = 0;
    if (active) {
        r = find_chunk_read(A.chunk_base, A.n_reads, r_first, c);
        rs_g = __ldg(A.offsets + r);
        n = static_cast<int>(__ldg(A.offsets + r + 1) - rs_g);
        a = static_cast<int>(c - __ldg(A.chunk_base + r)) * kChunk;
        b = min(a + kChunk, n);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    if (active) {
        const int tb = static_cast<int>(static_cast<int64_t>(rs_g) - *s_origin);    // text[x] of this lane's read = s_text[tb + x]
        s_lane_r[tid] = static_cast<int32_t>(r); s_lane_a[tid] = a; s_lane_t[tid] = tb;
        const unsigned char* text = s_text + tb;
        const uint32_t blk = (1u << q) - 1u;
        const uint32_t keep = ~(1u << 16);                      // the upper block's row 0 gets no horizontal input
        const uint32_t last2 = 0x00010001u;
        const int sb = q - 1;
        const uint32_t bias = (0x8000u - static_cast<uint32_t>(k + 1)) * 0x00010001u;
        const uint32_t pvmask = blk | (blk << 16);
        uint32_t pv = pvmask, mv = 0;
        uint32_t packed = static_cast<uint32_t>(q) * 0x00010001u + bias;       // bit 15 / 31 set <=> the block's cost >= k+1
        uint32_t fm = ~0u, rm = ~0u;                             // per base: 1 = "cost > k" (forward block / rc block)
#define BB_FSTEP(XX)                                                             \
        {                                                                        \
            const uint32_t e = s_eq[text[(XX)]];                                 \
            const uint32_t sum = (e & pv) + pv;                                  \
            uint32_t ph = mv | ~(sum | pv | e);                                  \
            uint32_t mh = pv & ((sum ^ pv) | e);                                 \
            packed += ((ph >> sb) & last2) - ((mh >> sb) & last2);               \
            ph = (ph << 1) & keep; mh <<= 1;                                     \
            pv = (mh | ~(e | mv | ph)) & pvmask;   /* spacer bits stay 0: no carry ripples from the lower block into the upper */ \
            mv = ph & (e | mv);                                                  \
            rm = __funnelshift_l(packed, rm, 1);                                 \
            fm = __funnelshift_l(packed << 16, fm, 1);                           \
        }
#pragma unroll 1
        for (int x0 = a - W; x0 < b; x0 += kGroup) {
            if (x0 < 0) continue;                               // W and the chunk are multiples of kGroup: groups never straddle 0
            if (x0 + kGroup <= b) {
#pragma unroll
                for (int t = 0; t < kGroup; t++) BB_FSTEP(x0 + t)
            } else {
#pragma unroll 1
                for (int x = x0; x < b; x++) BB_FSTEP(x)
                const int missing = x0 + kGroup - b;            // pad the partial group with "no candidate" flags
                fm = (fm << missing) | ((1u << missing) - 1u); rm = (rm << missing) | ((1u << missing) - 1u);
            }
            if (x0 >= a) {                                      // bit 19-t of the group's masks = base x0+t is a candidate
                const uint32_t cf = ~fm & 0xfffffu, cr = ~rm & 0xfffffu;
                const int gi = (x0 - a) / kGroup;
                s_bm32[gi * kScanThreads + tid] = cf | (cr << 20);
                s_bm8[gi * kScanThreads + tid] = static_cast<uint8_t>(cr >> 12);
            }
        }
#undef BB_FSTEP
        // ---- phase 2: candidate positions -> runs per strand, queued as {lane:8 | strand:1 | first - a:9 | length - 1:9} ----
        auto push = [&](int strand, int ps, int pe) {
            const uint32_t idx = atomicAdd(s_qn, 1u);
            if (idx < kRunQueue) s_runs[idx] = (static_cast<uint32_t>(tid) << 19) | (static_cast<uint32_t>(strand) << 18) |
                                               (static_cast<uint32_t>(ps - a) << 9) | static_cast<uint32_t>(pe - ps);
        };
        const int ng = (b - a + kGroup - 1) / kGroup;
        int f_s = 0, f_e = -2, r_s = 0, r_e = -2;               // run start / last position per strand (empty: e = -2)
#pragma unroll 1
        for (int gi = 0; gi < ng; gi++) {
            const uint32_t w32 = s_bm32[gi * kScanThreads + tid];
            const uint32_t w8 = s_bm8[gi * kScanThreads + tid];
            uint32_t cf = w32 & 0xfffffu, cr = (w32 >> 20) | (w8 << 12);
            const int p0 = a + gi * kGroup + 1;                  // position after the group's first base
            while (cf) {
                const int t = 19 - (31 - __clz(cf));             // candidates in ascending position: highest bit first
                cf &= ~(1u << (19 - t));
                const int p = p0 + t;
                if (p != f_e + 1) { if (f_e >= 0) push(BB_FWD, f_s, f_e); f_s = p; }
                f_e = p;
            }
            while (cr) {
                const int t = 19 - (31 - __clz(cr));
                cr &= ~(1u << (19 - t));
                const int p = p0 + t;
                if (p != r_e + 1) { if (r_e >= 0) push(BB_RC, r_s, r_e); r_s = p; }
                r_e = p;
            }
        }
        if (f_e >= 0) push(BB_FWD, f_s, f_e);
        if (r_e >= 0) push(BB_RC, r_s, r_e);
    }
    __syncthreads();
    uint32_t nq = *s_qn;
    if (nq > kRunQueue) {
        // More candidate runs than the CTA can queue (low-complexity text that keeps matching the run): every chunk of this tile
        // becomes ONE candidate run per strand (all of its positions), the pre-check is skipped, and the windows that follow from
        // those runs are verified exactly.
        nq = 0;
        if (active && b > a) {
            const int Df = m - (G.f_q0 + q), Dr = m - G.f_q0, ps = a + 1, pe = b;
            const int lo_f = max(ps + Df - k, 1), hi_f = min(pe + Df + k, n);
            const int lo_r = max(n - pe + Dr - k, 1), hi_r = min(n - ps + Dr + k, n);
            if (lo_f <= hi_f) s_stage[atomicAdd(s_wn, 1u)] = make_window(r, BB_FWD, lo_f, hi_f - lo_f);
            if (lo_r <= hi_r) s_stage[atomicAdd(s_wn, 1u)] = make_window(r, BB_RC, lo_r, hi_r - lo_r);
        }
    }
    // ---- phase 3: pre-check of the CTA's runs on the shared text, survivors -> windows ----
    {
        const int qs = G.f_qs;
        const int Df = m - (G.f_q0 + q), Dr = m - G.f_q0;
        const int d_f = (G.f_s0 + qs) - (G.f_q0 + q);            // end(S) - end(Q) in the flank (forward strand)
        const int d_r = G.f_q0 - G.f_s0;                         // end(rc S) - end(rc Q) in rc(flank)
        for (uint32_t it = tid; it < nq; it += kScanThreads) {
            const uint32_t e = s_runs[it];
            const int lt = static_cast<int>(e >> 19), strand = static_cast<int>((e >> 18) & 1u);
            const int ps = s_lane_a[lt] + static_cast<int>((e >> 9) & 0x1ffu), pe = ps + static_cast<int>(e & 0x1ffu);
            const uint32_t rr = static_cast<uint32_t>(s_lane_r[lt]);
            const int nn = static_cast<int>(__ldg(A.offsets + rr + 1) - __ldg(A.offsets + rr));
            const uint8_t* text = s_text + s_lane_t[lt];
            bool keep_run = true;
            if (qs > 0) {
                const int d = strand == BB_FWD ? d_f : d_r;
                const int slo = ps + d - k, shi = pe + d + k;
                if (slo >= 1 && shi <= nn) {                     // S entirely inside the read (else the read-end windows decide)
                    const int sbits = strand == BB_FWD ? 0 : 16;
                    const int cq = block_min_cost(text, s_eq, sbits, q, k, ps, pe);
                    if (cq > k) keep_run = false;                // cannot happen for a genuine candidate; cheap guard
                    else keep_run = cq + block_min_cost(text, s_seq, sbits, qs, k, slo, shi) <= k;
                }
            }
            if (!keep_run) continue;
            int lo, hi;
            if (strand == BB_FWD) { lo = ps + Df - k; hi = pe + Df + k; }
            else { lo = nn - pe + Dr - k; hi = nn - ps + Dr + k; }
            lo = max(lo, 1); hi = min(hi, nn);
            if (lo > hi) continue;
            const uint32_t idx = atomicAdd(s_wn, 1u);
            if (idx < kFiltStage) s_stage[idx] = make_window(rr, strand, lo, hi - lo);
        }
    }
    __syncthreads();
    // flush the CTA's windows with one global atomic
    const uint32_t nw = *s_wn;
    if (nw > kFiltStage) { if (tid == 0) atomicExch(F.overflow, 1u); return; }
    if (tid == 0) *s_wbase = nw ? atomicAdd(F.n_windows, nw) : 0u;
    __syncthreads();
    const uint32_t base = *s_wbase;
    if (base + nw > F.win_cap) { if (tid == 0) atomicExch(F.overflow, 1u); return; }
    for (uint32_t i = tid; i < nw; i += kScanThreads) F.windows[base + i] = s_stage[i];
}

// K1v: exact verification of windows; items [0, 4*n_reads) are the read-end windows, the rest come from the queue.
struct VerifyArgs {
    ScanArgs S;
    const uint64_t* windows;
    const uint32_t* n_windows;
    const uint32_t* overflow;    // set by the filter when a window queue overflowed: the queue has holes, the batch is re-run exactly
    uint32_t win_cap;
};

template <int NW>
__device__ void verify_window(const ScanArgs& A, const DevGroup& G, const uint64_t* __restrict__ s_eq, uint32_t r, int strand, int lo, int hi) {
    const uint64_t rs_g = __ldg(A.offsets + r);
    const int n = static_cast<int>(__ldg(A.offsets + r + 1) - rs_g);
    const int m = G.m, k = G.k, shift = 64 * NW - G.m;
    if (lo > hi) return;
    const uint8_t* text = A.bases + rs_g;
    const uint64_t* eq = s_eq + static_cast<size_t>(strand) * 256 * NW;
    // frame character c of this strand: forward text[c] / reverse-complement text[n-1-c] (masks are complemented)
    const int c_end = min(hi, n);                               // characters [c0, c_end) are consumed
    int c0 = max(lo, 1) - 1 - (m + k);                         // the cost at j depends on text[j-(m+k), j) only
    const bool fresh = c0 > 0;
    if (c0 < 0) c0 = 0;
    Col<NW> col;
#pragma unroll
    for (int w = 0; w < NW; w++) { col.pv[w] = fresh ? G.pv_plain_top[w] : G.pv_over_top[w]; col.mv[w] = 0; }
    int score = fresh ? m : G.ov_m;
    if (lo == 0 && score <= k) scan_emit(A, r, strand, 0u, score);
    const int64_t base = strand == BB_FWD ? 0 : static_cast<int64_t>(n) - 1, step = strand == BB_FWD ? 1 : -1;
    for (int c = c0; c < c_end; c += 8) {                        // eight independent byte loads in flight per round
        uint32_t chs[8];
#pragma unroll
        for (int t = 0; t < 8; t++) chs[t] = c + t < c_end ? __ldg(text + base + step * (c + t)) : 0u;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (c + t >= c_end) break;
            score += col_step_top<NW>(col, eq + chs[t] * NW);
            if (score <= k && c + t + 1 >= lo) scan_emit(A, r, strand, static_cast<uint32_t>(c + t + 1), score);
        }
    }
    if (hi > n && c_end == n) {                                  // virtual end positions past the text end (oracle policy S3)
        const int tmax = min(hi - n, m);
        for (int t = 1; t <= tmax; t++) {
            const int v = col_val<NW>(col.pv, col.mv, shift + m - t) + __ldg(G.ov + t);
            if (v <= k) scan_emit(A, r, strand, static_cast<uint32_t>(n + t), v);
        }
    }
}

template <int NW>
__global__ void __launch_bounds__(128) k_flank_verify(const VerifyArgs V, const DevGroup G) {
    __shared__ uint64_t s_eq[2 * 256 * NW];
    for (int i = threadIdx.x; i < 2 * 256 * NW; i += blockDim.x) s_eq[i] = __ldg(G.eq_top + i);
    __syncthreads();
    const ScanArgs& A = V.S;
    // an overflowed queue holds unwritten slots (a CTA that could not reserve returns without writing): nothing of it may be
    // decoded, the host re-runs the batch with the exact scan
    if (__ldg(V.overflow)) return;
    const uint64_t n_end = 4ull * A.n_reads;
    const uint64_t total = n_end + min(__ldg(V.n_windows), V.win_cap);
    const int span = G.m + G.k;
    for (uint64_t it = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; it < total; it += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        if (it < n_end) {
            // all head windows first, then all tail windows (about three times as long): the threads of a warp do the same kind
            const int which = it >= n_end / 2 ? 1 : 0;
            const uint64_t q = which ? it - n_end / 2 : it;
            const uint32_t r = static_cast<uint32_t>(q >> 1);
            const int strand = static_cast<int>(q & 1);
            const int n = static_cast<int>(__ldg(A.offsets + r + 1) - __ldg(A.offsets + r));
            if (n == 0) continue;
            // head: end positions [0, m+k]; tail: [n-m-k, n+m]; one window when they touch
            const bool joined = n - span <= span + 1;
            if (which == 0) verify_window<NW>(A, G, s_eq, r, strand, 0, joined ? n + G.m : span);
            else if (!joined) verify_window<NW>(A, G, s_eq, r, strand, n - span, n + G.m);
        } else {
            const uint64_t w = V.windows[it - n_end];
            const uint32_t r = static_cast<uint32_t>(w >> kWinReadShift);
            const int strand = static_cast<int>((w >> kWinStrandShift) & 1);
            const int lo = static_cast<int>((w >> kWinLoShift) & ((1u << 28) - 1)), len = static_cast<int>(w & ((1u << kWinLoShift) - 1));
            verify_window<NW>(A, G, s_eq, r, strand, lo, lo + len);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2a: local-minimum rule on the sorted entries (oracle policy S1)
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_resolve(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, const uint64_t* __restrict__ offsets,
                          const DevGroup* __restrict__ groups, uint8_t* __restrict__ flags, int pol) {
    const uint32_t n = __ldg(n_ptr);
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint64_t key = keys[e];
        const uint64_t id = key >> kKeyPosShift;
        const int cost = static_cast<int>(key & 0xff);
        const uint32_t r = static_cast<uint32_t>(key >> kKeyReadShift);
        const int g = static_cast<int>((key >> kKeyGroupShift) & 7);
        const uint32_t pos = static_cast<uint32_t>((key >> kKeyPosShift) & ((1u << 28) - 1));
        const uint32_t len = static_cast<uint32_t>(offsets[r + 1] - offsets[r]);
        const uint32_t last = len + groups[g].m;   // the flank searcher always has the overhang extension
        // A plateau of equal costs is a reported minimum when the last strict change before it was a decrease (or there is none)
        // and the first strict change after it is an increase (or the row ends); missing neighbours are > k >= cost.
        // walk left over the plateau: `dec` = it is entered by a decrease; `first` = this entry is its left end
        bool dec = true, first = true;
        uint64_t cur = id;
        for (uint32_t q = e; q > 0; q--) {
            const uint64_t pk = keys[q - 1];
            if ((pk >> kKeyPosShift) != cur - 1) break;
            const int pc = static_cast<int>(pk & 0xff);
            if (pc > cost) break;
            if (pc < cost) { dec = false; break; }
            cur--; first = false;
        }
        // walk right: `up` = it is left by an increase; `lastp` = this entry is its right end
        bool up = true, lastp = true;
        cur = id;
        uint32_t p = pos;
        for (uint32_t q = e; p != last && q + 1 < n; q++) {
            const uint64_t nk = keys[q + 1];
            if ((nk >> kKeyPosShift) != cur + 1) break;
            const int nc = static_cast<int>(nk & 0xff);
            if (nc > cost) break;
            if (nc < cost) { up = false; break; }
            cur++; p++; lastp = false;
            if (!(pol & kPolS1Left)) break;                    // right-end reporting only needs the next entry
        }
        flags[e] = (up && dec && ((pol & kPolS1Left) ? first : lastp)) ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2a, slot path: one WARP per read does what the global radix sort + unique + k_resolve + select do on the sorted path --
// the read's sub-threshold entries (a few dozen at most, duplicates from overlapping verification windows included) are
// ranked in shared memory (distinct keys in ascending order = sassy's order: group, forward before reverse-complement,
// ascending end position), the local-minimum rule (policy S1) is applied to the sorted run, and the reported matches are
// written back to the front of the read's slots.  No host round trip: the match counts are prefix-summed on the device.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSlotCap = 128;        // entry slots per read (k_read_resolve holds them in shared memory)
constexpr int kResolveWarps = 4;

__global__ void __launch_bounds__(kResolveWarps * 32) k_read_resolve(uint64_t* __restrict__ slots, const uint32_t* __restrict__ slot_cnt, uint32_t n_reads,
                                                                      const uint64_t* __restrict__ offsets, const DevGroup* __restrict__ groups,
                                                                      uint32_t* __restrict__ n_hits_read, int pol, uint32_t slot_cap) {
    __shared__ uint64_t s_raw[kResolveWarps][kSlotCap], s_sorted[kResolveWarps][kSlotCap];
    __shared__ uint8_t s_first[kResolveWarps][kSlotCap];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kResolveWarps + wib;
    if (r >= n_reads) return;
    const uint32_t n = min(__ldg(slot_cnt + r), slot_cap);     // (a read that overflowed its slots: the batch is re-run, only written slots are read)
    if (n == 0) { if (lane == 0) n_hits_read[r] = 0; return; }
    uint64_t* mine = slots + static_cast<size_t>(r) * kSlotCap;
    uint64_t* raw = s_raw[wib]; uint64_t* sorted = s_sorted[wib]; uint8_t* first = s_first[wib];
    for (uint32_t q = lane; q < n; q += 32) raw[q] = mine[q];
    __syncwarp();
    // first occurrence of every distinct key
    for (uint32_t q = lane; q < n; q += 32) {
        const uint64_t k = raw[q];
        bool f = true;
        for (uint32_t j = 0; j < q; j++) if (raw[j] == k) { f = false; break; }
        first[q] = f ? 1 : 0;
    }
    __syncwarp();
    // rank among the distinct keys
    uint32_t nd_local = 0;
    for (uint32_t q = lane; q < n; q += 32) {
        if (!first[q]) continue;
        const uint64_t k = raw[q];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; j++) rank += (first[j] && raw[j] < k) ? 1u : 0u;
        sorted[rank] = k;
        nd_local++;
    }
    uint32_t nd = nd_local;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, off);
    __syncwarp();
    // local-minimum rule on the sorted run (same walk as k_resolve), then ordered compaction of the reported matches
    const uint32_t len = static_cast<uint32_t>(offsets[r + 1] - offsets[r]);
    uint32_t out = 0;
    for (uint32_t base = 0; base < nd; base += 32) {
        const uint32_t e = base + lane;
        bool rep = false;
        uint64_t key = 0;
        if (e < nd) {
            key = sorted[e];
            const uint64_t id = key >> kKeyPosShift;
            const int cost = static_cast<int>(key & 0xff);
            const int g = static_cast<int>((key >> kKeyGroupShift) & 7);
            const uint32_t pos = static_cast<uint32_t>((key >> kKeyPosShift) & ((1u << 28) - 1));
            const uint32_t last = len + groups[g].m;
            bool dec = true, firstp = true;
            uint64_t cur = id;
            for (uint32_t q = e; q > 0; q--) {
                const uint64_t pk = sorted[q - 1];
                if ((pk >> kKeyPosShift) != cur - 1) break;
                const int pc = static_cast<int>(pk & 0xff);
                if (pc > cost) break;
                if (pc < cost) { dec = false; break; }
                cur--; firstp = false;
            }
            bool up = true, lastp = true;
            cur = id;
            uint32_t p = pos;
            for (uint32_t q = e; p != last && q + 1 < nd; q++) {
                const uint64_t nk = sorted[q + 1];
                if ((nk >> kKeyPosShift) != cur + 1) break;
                const int nc = static_cast<int>(nk & 0xff);
                if (nc > cost) break;
                if (nc < cost) { up = false; break; }
                cur++; p++; lastp = false;
                if (!(pol & kPolS1Left)) break;
            }
            rep = up && dec && ((pol & kPolS1Left) ? firstp : lastp);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, rep);
        if (rep) mine[out + __popc(m & ((1u << lane) - 1u))] = key;
        out += __popc(m);
    }
    if (lane == 0) n_hits_read[r] = out;
}

// the reads' reported matches (front of their slots) -> one list in read order; thread per read
__global__ void k_hits_gather(const uint64_t* __restrict__ slots, const uint32_t* __restrict__ n_hits_read, const uint32_t* __restrict__ hit_base,
                              uint32_t n_reads, uint32_t hits_cap, uint64_t* __restrict__ hit_keys, uint32_t* __restrict__ n_hits_out, uint32_t* __restrict__ overflow) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) {
        const uint32_t total = hit_base[n_reads];
        if (total > hits_cap) { atomicExch(overflow, 1u); *n_hits_out = 0; } else *n_hits_out = total;
    }
    if (r >= n_reads) return;
    const uint32_t n = n_hits_read[r], b = hit_base[r];
    if (b + n > hits_cap) return;
    const uint64_t* src = slots + static_cast<size_t>(r) * kSlotCap;
    for (uint32_t q = 0; q < n; q++) hit_keys[b + q] = src[q];
}

// ---------------------------------------------------------------------------------------------------------------
// K2b: traceback of the reported flank matches (oracle policies S2-S4) + barcode region
// ---------------------------------------------------------------------------------------------------------------
struct TraceArgs {
    const uint8_t* bases;
    const uint64_t* offsets;
    const uint64_t* hit_keys;
    const uint32_t* n_hits;  // device-side count
    const DevGroup* groups;
    uint64_t* hist;          // [col][2*NW][slot]
    uint32_t n_slots;
    Hit* hits;
    int pol;                 // kPolS2PatFirst, kPolS6RcFirst
};

template <int NW>
__device__ void trace_one(const TraceArgs& A, const DevGroup& G, uint32_t h, uint32_t slot) {
    const uint64_t key = A.hit_keys[h];
    const uint32_t r = static_cast<uint32_t>(key >> kKeyReadShift);
    const int g = static_cast<int>((key >> kKeyGroupShift) & 7);
    const int strand = static_cast<int>((key >> kKeyStrandShift) & 1) ^ ((A.pol & kPolS6RcFirst) ? 1 : 0);
    const int pos = static_cast<int>((key >> kKeyPosShift) & ((1u << 28) - 1));
    const int cost = static_cast<int>(key & 0xff);
    const uint64_t rs0 = A.offsets[r];
    const int n = static_cast<int>(A.offsets[r + 1] - rs0);
    const uint8_t* text = A.bases + rs0;
    const int m = G.m;
    const uint64_t* eq = G.eq + static_cast<size_t>(strand) * 256 * NW;
    const int jend = pos <= n ? pos : n, iend = pos <= n ? m : m - (pos - n);
    int s0 = jend - G.trace_cols; if (s0 < 0) s0 = 0;
    const int edge = s0 > 0 ? s0 : -1;
    const size_t S = A.n_slots;
    uint64_t* hist = A.hist + slot;
#define BB_H(col, w) hist[(static_cast<size_t>(col) * (2 * NW) + (w)) * S]
#define BB_FRAME_CHAR(j) (strand == BB_FWD ? text[(j)] : text[n - 1 - (j)])
    Col<NW> col;
#pragma unroll
    for (int w = 0; w < NW; w++) { col.pv[w] = (s0 == 0) ? G.pv_over[w] : G.pv_plain[w]; col.mv[w] = 0; }
#pragma unroll
    for (int w = 0; w < NW; w++) { BB_H(0, w) = col.pv[w]; BB_H(0, NW + w) = 0; }
    for (int j = s0; j < jend; j++) {
        const uint32_t ch = BB_FRAME_CHAR(j);
        uint64_t e[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) e[w] = __ldg(eq + ch * NW + w);
        col_step<NW>(col, e, G.last_bit);
#pragma unroll
        for (int w = 0; w < NW; w++) { BB_H(j - s0 + 1, w) = col.pv[w]; BB_H(j - s0 + 1, NW + w) = col.mv[w]; }
    }
    auto cell = [&](int i, int j) -> int {
        uint64_t pv[NW], mv[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) { pv[w] = BB_H(j - s0, w); mv[w] = BB_H(j - s0, NW + w); }
        return col_val<NW>(pv, mv, i);
    };
    int i = iend, j = jend;
    int cnt = 0, j_first = 0, j_last = 0;   // path entries with bar0 <= i <= bar1: first/last in PATH order
    while (i > 0) {
        if (j == 0) break;                                    // S3: left overhang (the flank searcher always has alpha)
        int di = 1, dj = 0;
        if (j != edge) {
            const int gcur = cell(i, j), d = cell(i - 1, j - 1);
            const uint32_t ch = BB_FRAME_CHAR(j - 1);
            const bool match = (__ldg(eq + ch * NW + ((i - 1) >> 6)) >> ((i - 1) & 63)) & 1ull;
            if (match && d == gcur) { di = 1; dj = 1; }
            else if (d + 1 == gcur) { di = 1; dj = 1; }
            else if (!(A.pol & kPolS2PatFirst)) {
                if (cell(i, j - 1) + 1 == gcur) { di = 0; dj = 1; } else { di = 1; dj = 0; }
            } else {
                if (cell(i - 1, j) + 1 == gcur) { di = 1; dj = 0; } else { di = 0; dj = 1; }
            }
        }
        i -= di; j -= dj;
        if (i >= G.bar0 && i <= G.bar1) {                     // pre-op position (i, j) of this op
            if (cnt == 0) j_last = j;
            j_first = j;
            cnt++;
        }
    }
#undef BB_H
#undef BB_FRAME_CHAR
    Hit out;
    out.read = r; out.group = g; out.strand = strand; out.cost = cost;
    if (strand == BB_FWD) { out.text_start = j; out.text_end = jend; }
    else { out.text_start = n - jend; out.text_end = n - j; }
    out.has_region = cnt >= 2;
    int64_t p0 = strand == BB_FWD ? j_first : static_cast<int64_t>(n) - 1 - j_first;
    int64_t p1 = strand == BB_FWD ? j_last : static_cast<int64_t>(n) - 1 - j_last;
    if (p0 < 0) p0 = 0;
    if (p1 < 0) p1 = 0;
    int64_t lo = p0 < p1 ? p0 : p1, hi = p0 < p1 ? p1 : p0;
    lo = lo > kPadding ? lo - kPadding : 0;                   // searcher.rs:453
    hi = hi + kPadding < n ? hi + kPadding : n;               // searcher.rs:454
    if (hi < lo) hi = lo;
    out.rs = static_cast<int32_t>(lo); out.re = static_cast<int32_t>(hi);
    A.hits[h] = out;
}

__global__ void __launch_bounds__(64) k_trace(const TraceArgs A) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_hits = __ldg(A.n_hits);
    for (uint32_t h = slot; h < n_hits; h += A.n_slots) {
        const int g = static_cast<int>((A.hit_keys[h] >> kKeyGroupShift) & 7);
        const DevGroup& G = A.groups[g];
        if (G.nw == 1) trace_one<1>(A, G, h, slot); else trace_one<2>(A, G, h, slot);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: barcode stage, one warp per flank match
// ---------------------------------------------------------------------------------------------------------------
struct BarArgs {
    const uint8_t* bases;
    const uint64_t* offsets;
    const Hit* hits;
    const uint32_t* n_hits;   // device-side count of flank matches
    const DevGroup* groups;
    const uint8_t* code;      // [256] byte -> 4-bit IUPAC set
    Params prm;
    bb_row* rows;             // one slot per hit
    uint8_t* row_valid;
    int rn_lo, rn_hi;         // this launch takes the flank matches whose region has rn_lo < bases <= rn_hi
    int sh_rows;              // shared-memory rows reserved for the records of the shared leading rows / the per-row traceback records
    int pol;                  // kPolS1Left | kPolS5Last (S2 is a template parameter)
};

constexpr int kMaxBarRounds = 128;  // up to 4096 barcodes per group (a bound on the pattern table, 256 KB per strand at that size; the kernels loop over rounds)
#ifndef BB_K3_MIN_CTAS
#define BB_K3_MIN_CTAS 18            // resident warps (= CTAs) per SM the register allocation of k_barcode_rows aims at
#endif

__device__ __forceinline__ int64_t rel_dist_to_end(int64_t pos, int64_t read_len) {   // searcher.rs:183-199
    if (pos < 0) return 1;
    if (pos <= read_len / 2) return pos == 0 ? 1 : pos;
    if (pos == read_len) return -1;
    return -(read_len - pos);
}

__device__ __forceinline__ void fill_flank_row(bb_row& row, const Hit& H, const DevGroup& G, int n) {   // searcher.rs:241-265
    row.read_idx = H.read; row.read_len = static_cast<uint32_t>(n);
    row.rel_dist_to_end = rel_dist_to_end(H.text_start, n);
    row.read_start_bar = H.text_start; row.read_end_bar = H.text_end;
    row.read_start_flank = H.text_start; row.read_end_flank = H.text_end;
    row.bar_start = 0; row.bar_end = 0;
    row.flank_cost = H.cost; row.barcode_cost = G.bar_len; row.label_idx = -1; row.group_idx = H.group;
    row.match_type = static_cast<uint8_t>(G.match_type == BB_FTAG ? BB_FFLANK : BB_RFLANK);
    row.strand = static_cast<uint8_t>(H.strand);
    for (int q = 0; q < 6; q++) row.pad_[q] = 0;
}

// running top-two of the candidates a lane has scored, under (normalised score desc, barcode index asc)
struct TopTwo {
    double top_s = -1.0, sec_s = -1.0;
    int top_b = 1 << 30;
    int ok = 0, pi = 0, ei = 0, pj = 0, ej = 0, cost = 0, ts = 0, te = 0;   // map_pat_to_text_with_cost of the top
};

// Shared memory of one k_barcode_rows CTA (= one warp): the scan table, the 16 text masks, the records of the shared leading rows,
// one traceback record byte per row and lane, and the lanes' own-row records.
#ifndef BB_K3_LUT_GLOBAL
#define BB_K3_LUT_GLOBAL 1
#endif
#if BB_K3_LUT_GLOBAL
__device__ uint32_t g_k3_scan_lut[256];      // scan_lut_entry(), filled by Engine::init (read through L1: frees 1 KB of every CTA's shared memory)
constexpr size_t kK3LutBytes = 0;
#else
constexpr size_t kK3LutBytes = 1024;
#endif
template <int NWT, bool PACKED, bool MITM>
__host__ __device__ inline size_t barcode_rows_smem(int sh_rows, int own_rows) {
    return kK3LutBytes + 16 * NWT * 8 + ((static_cast<size_t>(sh_rows) * 3 * NWT * 8 + 15) & ~static_cast<size_t>(15)) + 64 + static_cast<size_t>(sh_rows) * 32 +
           static_cast<size_t>(sh_rows) * kOffStride + row_hist_bytes<NWT, PACKED>(resident_rows(own_rows, MITM)) + 16;
}

// One warp (= one CTA) per flank match; lane = barcode pattern (rounds of 32).  The region's text masks and the pattern rows
// all barcodes share are computed once per flank match, the per-pattern work is rows_lane() (barcode_rows.cuh: forward pass
// over the pattern's own rows with one record per row, row-per-iteration traceback, Lodhi score).
// The per-pattern best minimum is the same with k = floor(0.4*len) and with the fallback k = len (the first
// lowest-cost minimum); the threshold only decides WHICH patterns are candidates, so both candidate sets are reduced
// side by side and the fallback rule (searcher.rs:303-306) picks one at the end.
// Launches: regions of <= 48 bases with 12-byte records (PACKED), 49..64 with 16-byte records, longer ones with three text words.
// MITM = only half of the own rows' records are resident (more warps in flight for some replayed forward rows).
template <int NWT, bool PACKED, bool S2PAT, bool MITM>
__global__ void __launch_bounds__(32, NWT == 1 ? BB_K3_MIN_CTAS : 1) k_barcode_rows(const BarArgs A) {
    extern __shared__ __align__(16) unsigned char bar_smem[];
    const int lane = threadIdx.x;
#if BB_K3_LUT_GLOBAL
    const uint32_t* lut = g_k3_scan_lut;                                                   // [256] bottom-row scan table
#else
    uint32_t* lut = reinterpret_cast<uint32_t*>(bar_smem);
    for (int q = threadIdx.x; q < 256; q += 32) lut[q] = scan_lut_entry(q);
#endif
    uint64_t* tm = reinterpret_cast<uint64_t*>(bar_smem + kK3LutBytes);                    // [16][NWT]
    uint64_t* sh = tm + 16 * NWT;                                                          // [sh_rows][3][NWT]
    uint8_t* s_shoff = reinterpret_cast<uint8_t*>(sh) + ((static_cast<size_t>(A.sh_rows) * 3 * NWT * 8 + 15) & ~static_cast<size_t>(15));   // [64] codes of the shared rows
    uint8_t* rec = s_shoff + 64;                                                           // [sh_rows][32]
    uint8_t* s_off = rec + static_cast<size_t>(A.sh_rows) * 32;                            // [sh_rows][32 lanes] pattern codes of the round
    const RowHist<NWT, PACKED> hist{reinterpret_cast<uint32_t*>(s_off + static_cast<size_t>(A.sh_rows) * kOffStride), lane};
    const uint32_t n_hits = __ldg(A.n_hits);
    for (uint32_t h = blockIdx.x; h < n_hits; h += gridDim.x) {
        const Hit H = A.hits[h];
        if (!H.has_region) { if (A.rn_lo < 0 && lane == 0) A.row_valid[h] = 0; continue; }    // searcher.rs:445-449
        const int rn = H.re - H.rs;
        if (rn <= A.rn_lo || rn > A.rn_hi) continue;            // another launch's flank matches
        const DevGroup& G = A.groups[H.group];
        const uint64_t rs0 = A.offsets[H.read];
        const int n = static_cast<int>(A.offsets[H.read + 1] - rs0);
        const int L = G.bar_len, nb = G.n_barcodes, k1 = G.k_bar;
        const int P = H.strand == BB_FWD ? G.sh_p[0] : G.sh_p[1];
        // ---- text masks: B[a] = region bases whose set holds base a; tm[c] = bases that match a pattern character with set c ----
        __syncwarp();
        {
            const uint8_t* tp = A.bases + rs0 + H.rs;
            uint64_t B[4][NWT];
#pragma unroll
            for (int w = 0; w < NWT; w++) {
                const int q0 = 64 * w + lane, q1 = q0 + 32;
                const uint32_t c0 = q0 < rn ? __ldg(A.code + tp[q0]) : 0u, c1 = q1 < rn ? __ldg(A.code + tp[q1]) : 0u;
#pragma unroll
                for (int a = 0; a < 4; a++)
                    B[a][w] = __ballot_sync(0xffffffffu, (c0 >> a) & 1u) | (static_cast<uint64_t>(__ballot_sync(0xffffffffu, (c1 >> a) & 1u)) << 32);
            }
            if (lane < 16) {
#pragma unroll
                for (int w = 0; w < NWT; w++) {
                    uint64_t v = 0;
#pragma unroll
                    for (int a = 0; a < 4; a++) v |= ((lane >> a) & 1) ? B[a][w] : 0ull;
                    tm[lane * NWT + w] = v;
                }
                reinterpret_cast<uint32_t*>(s_shoff)[lane] = __ldg(reinterpret_cast<const uint32_t*>(G.sh_off + 64 * H.strand) + lane);
            }
        }
        __syncwarp();
        // ---- the leading rows all barcodes of this strand share: once per flank match ----
        uint64_t ph0[NWT], mh0[NWT];
        rows_prefix<NWT, S2PAT>(tm, s_shoff, P, lane == 0, sh, ph0, mh0);
        const uint8_t* offs_g = G.bar_off + static_cast<size_t>(H.strand) * ((nb + 31) / 32) * (64 * 32);   // [round][row][lane]

        TopTwo all, strict;          // candidates under the fallback k = len / under k1
        int matched = 0;
#pragma unroll 1
        for (int rd = 0; rd * 32 < nb; rd++) {
            const int b = rd * 32 + lane;
            bool has1 = false;
            __syncwarp();
            {   // the round's codes: one contiguous [row][lane] block of the table -> shared memory, 16 bytes per lane and step
                const uint4* src = reinterpret_cast<const uint4*>(offs_g + static_cast<size_t>(rd) * (64 * 32));
                uint4* dst = reinterpret_cast<uint4*>(s_off);
                for (int q = lane; q < 2 * L; q += 32) dst[q] = __ldg(src + q);
            }
            __syncwarp();
            if (b < nb) {
                LaneAlign R;
                rows_lane<NWT, PACKED, S2PAT, MITM>(tm, s_off + lane, rn, L, P, ph0, mh0, sh, hist, rec + lane, lut, G.pbar0, G.pbar1, A.pol, R);
                has1 = R.cbest <= k1;
                const double sn = G.perfect > 0.0 ? R.s / G.perfect : 0.0;
#pragma unroll
                for (int set = 0; set < 2; set++) {
                    TopTwo& T = set == 0 ? all : strict;
                    if (set == 1 && !has1) continue;
                    if (sn > T.top_s) {                       // ascending b within a lane: strict > keeps the lower index
                        T.sec_s = T.top_s;
                        T.top_s = sn; T.top_b = b;
                        T.ok = R.cnt > 0; T.pi = R.i_first; T.ei = R.i_last; T.pj = R.j_first; T.ej = R.j_last; T.cost = R.sub_cost;
                        T.ts = R.ts; T.te = R.jend;
                    } else if (sn > T.sec_s) T.sec_s = sn;
                }
            }
            matched += __popc(__ballot_sync(0xffffffffu, has1));
        }
        const bool fallback = matched <= 1 && k1 < L;          // searcher.rs:303-306
        const TopTwo& T = fallback ? all : strict;
        const int total_cand = fallback ? nb : matched;
        // warp reduction: global top (score desc, index asc), then the best of the rest
        double g_s = T.top_s; int g_b = T.top_b;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, g_s, off);
            const int ob = __shfl_xor_sync(0xffffffffu, g_b, off);
            if (os > g_s || (os == g_s && ob < g_b)) { g_s = os; g_b = ob; }
        }
        const bool owner = (g_b == T.top_b) && T.top_b != (1 << 30);
        double rest = owner ? T.sec_s : T.top_s;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, rest, off);
            if (os > rest) rest = os;
        }
        if (total_cand == 0) {
            if (lane == 0) { bb_row row; fill_flank_row(row, H, G, n); A.rows[h] = row; A.row_valid[h] = 1; }
            continue;
        }
        if (owner) {
            bool valid = g_s >= A.prm.min_score;                                   // searcher.rs:391-396
            if (total_cand > 1) valid = valid && (g_s - rest) >= A.prm.min_score_diff;
            bb_row row;
            if (valid && T.ok) {
                // to_path of the candidate with its strand overwritten by the flank's (S4, searcher.rs:333)
                int64_t pj, ej;
                if (H.strand == BB_FWD) { pj = T.pj; ej = T.ej; }
                else { pj = static_cast<int64_t>(T.te) - 1 - (T.pj - T.ts); ej = static_cast<int64_t>(T.te) - 1 - (T.ej - T.ts); }
                row.read_idx = H.read; row.read_len = static_cast<uint32_t>(n);
                row.rel_dist_to_end = rel_dist_to_end(H.text_start, n);
                row.read_start_bar = H.rs + pj; row.read_end_bar = H.rs + ej + 1;
                row.read_start_flank = H.text_start; row.read_end_flank = H.text_end;
                row.bar_start = H.rs + T.pi; row.bar_end = H.rs + T.ei + 1;
                row.flank_cost = H.cost; row.barcode_cost = T.cost; row.label_idx = g_b; row.group_idx = H.group;
                row.match_type = static_cast<uint8_t>(G.match_type); row.strand = static_cast<uint8_t>(H.strand);
                for (int q = 0; q < 6; q++) row.pad_[q] = 0;
            } else {
                fill_flank_row(row, H, G, n);
            }
            A.rows[h] = row; A.row_valid[h] = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K4: collapse_overlapping_matches (interval.rs:4-79), one thread per read (the thread of the read's first hit)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool rows_overlap(const bb_row& a, const bb_row& b, float thr) {
    const int64_t s = a.read_start_flank > b.read_start_flank ? a.read_start_flank : b.read_start_flank;
    const int64_t e = a.read_end_flank < b.read_end_flank ? a.read_end_flank : b.read_end_flank;
    if (e <= s) return false;
    const int64_t la = a.read_end_flank - a.read_start_flank, lb = b.read_end_flank - b.read_start_flank;
    const int64_t mn = la < lb ? la : lb;
    return (static_cast<float>(e - s) / static_cast<float>(mn)) >= thr;
}
__device__ __forceinline__ bool row_better(const bb_row& a, const bb_row& b) {
    const int pa = a.match_type <= BB_RTAG ? 1 : 2, pb = b.match_type <= BB_RTAG ? 1 : 2;
    if (pa != pb) return pa < pb;
    if (pa == 1) {
        if (a.barcode_cost != b.barcode_cost) return a.barcode_cost < b.barcode_cost;
        return a.flank_cost < b.flank_cost;
    }
    return (a.read_end_flank - a.read_start_flank) > (b.read_end_flank - b.read_start_flank);
}

__global__ void k_collapse(const Hit* __restrict__ hits, const uint32_t* __restrict__ n_hits_p, bb_row* rows, uint8_t* row_valid,
                           unsigned long long* kept_reads) {
    const uint32_t n_hits = __ldg(n_hits_p);
    const uint32_t h0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (h0 >= n_hits) return;
    const uint32_t read = hits[h0].read;
    if (h0 > 0 && hits[h0 - 1].read == read) return;
    uint32_t h1 = h0 + 1;
    while (h1 < n_hits && hits[h1].read == read) h1++;
    // compact the valid rows to the front of the read's segment (stable)
    uint32_t nr = 0;
    for (uint32_t h = h0; h < h1; h++)
        if (row_valid[h]) { if (nr != h - h0) rows[h0 + nr] = rows[h]; nr++; }
    bb_row* R = rows + h0;
    // stable insertion sort by read_start_flank
    for (uint32_t i = 1; i < nr; i++) {
        const bb_row x = R[i];
        int j = static_cast<int>(i) - 1;
        while (j >= 0 && R[j].read_start_flank > x.read_start_flank) { R[j + 1] = R[j]; j--; }
        R[j + 1] = x;
    }
    uint32_t out = 0, g0 = 0;
    for (uint32_t i = 1; i <= nr; i++) {
        bool joins = false;
        if (i < nr) for (uint32_t q = g0; q < i; q++) if (rows_overlap(R[q], R[i], 0.8f)) { joins = true; break; }
        if (!joins) {
            uint32_t best = g0;
            for (uint32_t q = g0 + 1; q < i; q++) if (row_better(R[q], R[best])) best = q;
            const bb_row bsel = R[best];
            R[out++] = bsel;
            g0 = i;
        }
    }
    for (uint32_t h = h0; h < h1; h++) row_valid[h] = (h - h0) < out ? 1 : 0;
    if (out > 0) atomicAdd(kept_reads, 1ull);
}

// Expands the nibble-packed wire format of the host->device copy (host/pack.cpp) back to one representative letter per
// 4-bit IUPAC base set; 16 bases per thread (8 bytes in, 16 bytes out).  The search only sees a text byte through its
// base set, so the round trip is lossless for this path.
__global__ void k_unpack_nibbles(const uint8_t* __restrict__ packed, uint8_t* __restrict__ bases, uint64_t n_bases) {
    const uint64_t g = (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x) * 16;
    if (g >= n_bases) return;
    // code -> letter: A=1 C=2 G=4 T=8 and their unions; the empty set (non-IUPAC byte) -> 'X' (matches nothing)
    const uint64_t lut_lo = 0x565352474d434158ull;   // codes 0..7 : X A C M G R S V   (little endian bytes)
    const uint64_t lut_hi = 0x4e42444b48595754ull;   // codes 8..15: T W Y H K D B N
    const uint2 in = *reinterpret_cast<const uint2*>(packed + (g >> 1));
    uint32_t out[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t src = (w < 2 ? in.x : in.y) >> ((w & 1) * 16);
        uint32_t v = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint32_t c = (src >> (4 * t)) & 15u;
            const uint32_t ch = static_cast<uint32_t>(((c & 8u) ? lut_hi : lut_lo) >> ((c & 7u) * 8)) & 0xffu;
            v |= ch << (8 * t);
        }
        out[w] = v;
    }
    *reinterpret_cast<uint4*>(bases + g) = make_uint4(out[0], out[1], out[2], out[3]);
}

// The denser wire format (host/pack.cpp pack_crumbs): 2 bits per base for A C G T, 16 bases per thread (4 bytes in, 16 bytes out) ...
__global__ void k_unpack_crumbs(const uint8_t* __restrict__ packed, uint8_t* __restrict__ bases, uint64_t n_bases) {
    const uint64_t g = (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x) * 16;
    if (g >= n_bases) return;
    const uint32_t lut = 0x54474341u;                 // crumbs 0..3 : A C G T (little endian bytes)
    const uint32_t in = *reinterpret_cast<const uint32_t*>(packed + (g >> 2));
    uint32_t out[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint32_t v = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) v |= ((lut >> (((in >> (8 * w + 2 * t)) & 3u) * 8)) & 0xffu) << (8 * t);
        out[w] = v;
    }
    *reinterpret_cast<uint4*>(bases + g) = make_uint4(out[0], out[1], out[2], out[3]);
}
// ... and every byte that is not a single base arrives as (position << 4 | base set) and is written over the placeholder
// (same set -> letter table as k_unpack_nibbles; ~0 = unused entry of a partly filled block)
__global__ void k_patch_exceptions(const uint64_t* __restrict__ exc, uint64_t n_exc, uint8_t* __restrict__ bases, uint64_t n_bases) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_exc) return;
    const uint64_t v = exc[i];
    if (v == ~0ull || (v >> 4) >= n_bases) return;
    const uint64_t lut_lo = 0x565352474d434158ull, lut_hi = 0x4e42444b48595754ull;
    const uint32_t c = static_cast<uint32_t>(v) & 15u;
    bases[v >> 4] = static_cast<uint8_t>(((c & 8u) ? lut_hi : lut_lo) >> ((c & 7u) * 8));
}

// flank hit list for parity checks of the flank stage alone
__global__ void k_export_hits(const Hit* __restrict__ hits, uint32_t n_hits, int32_t* __restrict__ out6) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hits) return;
    const Hit H = hits[h];
    int32_t* o = out6 + 6 * static_cast<size_t>(h);
    o[0] = static_cast<int32_t>(H.read); o[1] = H.group; o[2] = H.strand; o[3] = H.text_start; o[4] = H.text_end; o[5] = H.cost;
}

}  // namespace bb
