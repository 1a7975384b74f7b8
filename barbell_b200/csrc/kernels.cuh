// sm_100a kernels of the annotate hot path.
//
//  K1 flank_scan    reference searcher.rs:438 (overhang_searcher.search over the WHOLE read, both strands):
//                   bit-vector edit distance of the N-masked flank against every text position; emits every end
//                   position whose cost is <= k ("sub-threshold entries").  Text tiles arrive in shared memory by one
//                   TMA bulk copy per CTA; one lane owns one text chunk; match masks are staged once in shared memory.
//  K2a resolve      sassy's local-minimum reporting rule applied to the sorted entries.
//  K2b trace        traceback of every reported flank match -> text_start and the barcode text region
//                   (reference cigar_parse.rs:71-82, searcher.rs:442-456).
//  K3  barcode      reference searcher.rs:267-426: all barcodes of the group against the region (one warp per flank
//                   match, barcodes across lanes), fallback pass, traceback, Lodhi score, thresholds, row assembly.
//  K4  collapse     reference interval.rs:4-79, one thread per read.
//
// The arithmetic mirrors oracle/barbell_oracle.c bit for bit (policies S1-S7 there).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/barbell_b200.h"
#include "device_types.cuh"

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

__device__ __noinline__ void emit_entry(uint64_t* entries, uint32_t* n_entries, uint32_t cap, uint64_t key) {
    const uint32_t idx = atomicAdd(n_entries, 1u);
    if (idx < cap) entries[idx] = key;
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
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;

struct ScanArgs {
    const uint8_t* bases;        // concatenated read bytes, 16-byte aligned
    const uint64_t* offsets;     // n_reads + 1
    uint32_t n_reads;
    uint64_t total;              // bytes in `bases`
    uint64_t total16;            // readable bytes (total rounded up to 16)
    int group;
    int chunk;                   // bytes of text per lane (multiple of 16; odd multiple keeps LDS.128 conflict-free)
    uint64_t* entries;
    uint32_t* n_entries;
    uint32_t cap;
};

// index of the read containing global byte g: largest r with offsets[r] <= g (skipping empty reads lands on the
// non-empty one because upper_bound returns the first offset > g)
__device__ __forceinline__ uint32_t find_read(const uint64_t* __restrict__ offsets, uint32_t n_reads, uint64_t g) {
    uint32_t lo = 0, hi = n_reads + 1;   // first index with offsets[idx] > g
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= g) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

template <int NW>
__global__ void __launch_bounds__(kScanThreads, 2) k_flank_scan(const ScanArgs A, const DevGroup G) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* s_eq = reinterpret_cast<uint64_t*>(smem + 128);                       // [2][256][NW]
    unsigned char* s_text = smem + 128 + 2 * 256 * NW * sizeof(uint64_t);           // halo + tile + halo

    const int tid = threadIdx.x;
    const int halo = G.halo;
    const uint64_t tile_bytes = static_cast<uint64_t>(kScanThreads) * A.chunk;
    const uint64_t tile_base = static_cast<uint64_t>(blockIdx.x) * tile_bytes;
    const int64_t s_origin = static_cast<int64_t>(tile_base) - halo;                // global byte at s_text[0]

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint64_t lo = tile_base >= static_cast<uint64_t>(halo) ? tile_base - halo : 0;
        uint64_t hi = tile_base + tile_bytes + halo;
        if (hi > A.total16) hi = A.total16;
        const uint32_t bytes = static_cast<uint32_t>(hi - lo);
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(s_text + (static_cast<int64_t>(lo) - s_origin), A.bases + lo, bytes, bar);
    }
    // stage the match masks while the bulk copy is in flight
    for (int i = tid; i < 2 * 256 * NW; i += kScanThreads) s_eq[i] = __ldg(G.eq + i);
    __syncthreads();
    mbar_wait(bar, 0);

    const uint64_t g0 = tile_base + static_cast<uint64_t>(tid) * A.chunk;
    if (g0 >= A.total) return;
    const uint64_t g1 = (g0 + A.chunk < A.total) ? g0 + A.chunk : A.total;
    const int m = G.m, k = G.k, last_bit = G.last_bit, W = G.halo;
    const int* __restrict__ ov = G.ov;
    const uint64_t* eq_f = s_eq;
    const uint64_t* eq_r = s_eq + 256 * NW;

#define BB_TEXT(x) (s_text[static_cast<int64_t>(x) - s_origin])
#define BB_STEP(EQ, CH, EMITCOND, POS)                                                    \
    {                                                                                     \
        score += col_step<NW>(col, (EQ) + static_cast<uint32_t>(CH) * NW, last_bit);      \
        if (score <= k) {                                                                 \
            if (EMITCOND) emit_entry(A.entries, A.n_entries, A.cap, make_key(r, A.group, strand, static_cast<uint32_t>(POS), score)); \
        }                                                                                 \
    }

    uint32_t r = find_read(A.offsets, A.n_reads, g0);
    for (; r < A.n_reads; r++) {
        const uint64_t rs = __ldg(A.offsets + r), re = __ldg(A.offsets + r + 1);
        if (rs >= g1) break;
        if (re <= rs || re <= g0) continue;
        const uint64_t a = rs > g0 ? rs : g0, b = re < g1 ? re : g1;   // this lane reports end positions in (a, b]
        const uint32_t n = static_cast<uint32_t>(re - rs);

        // ---------------- forward strand: ascending text ----------------
        {
            const int strand = BB_FWD;
            const uint64_t ws = (a - rs > static_cast<uint64_t>(W)) ? a - W : rs;
            Col<NW> col;
            int score;
#pragma unroll
            for (int w = 0; w < NW; w++) { col.pv[w] = (ws == rs) ? G.pv_over[w] : G.pv_plain[w]; col.mv[w] = 0; }
            score = (ws == rs) ? G.ov_m : m;
            if (a == rs && score <= k) emit_entry(A.entries, A.n_entries, A.cap, make_key(r, A.group, strand, 0u, score));
            uint64_t x = ws;
            while (x < b && (x & 15)) { BB_STEP(eq_f, BB_TEXT(x), x >= a, x - rs + 1); x++; }
            while (x + 16 <= b) {
                const uint4 w4 = *reinterpret_cast<const uint4*>(&BB_TEXT(x));
                const uint32_t ws4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint32_t ch = (ws4[q] >> (8 * t)) & 0xffu;
                        BB_STEP(eq_f, ch, x + (4 * q + t) >= a, x + (4 * q + t) - rs + 1);
                    }
                }
                x += 16;
            }
            while (x < b) { BB_STEP(eq_f, BB_TEXT(x), x >= a, x - rs + 1); x++; }
            if (b == re) {   // virtual end positions past the text end (oracle policy S3)
                for (int t = 1; t <= m; t++) {
                    const int v = col_val<NW>(col.pv, col.mv, m - t) + __ldg(ov + t);
                    if (v <= k) emit_entry(A.entries, A.n_entries, A.cap, make_key(r, A.group, strand, n + t, v));
                }
            }
        }
        // ---------------- reverse-complement strand: descending text, complemented masks ----------------
        {
            const int strand = BB_RC;
            const uint64_t we = (re - b > static_cast<uint64_t>(W)) ? b + W : re;
            Col<NW> col;
            int score;
#pragma unroll
            for (int w = 0; w < NW; w++) { col.pv[w] = (we == re) ? G.pv_over[w] : G.pv_plain[w]; col.mv[w] = 0; }
            score = (we == re) ? G.ov_m : m;
            if (b == re && score <= k) emit_entry(A.entries, A.n_entries, A.cap, make_key(r, A.group, strand, 0u, score));
            uint64_t x = we;   // next char to consume is x-1
            while (x > a && (x & 15)) { x--; BB_STEP(eq_r, BB_TEXT(x), x < b, re - x); }
            while (x >= a + 16) {
                x -= 16;
                const uint4 w4 = *reinterpret_cast<const uint4*>(&BB_TEXT(x));
                const uint32_t ws4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int q = 3; q >= 0; q--) {
#pragma unroll
                    for (int t = 3; t >= 0; t--) {
                        const uint32_t ch = (ws4[q] >> (8 * t)) & 0xffu;
                        BB_STEP(eq_r, ch, x + (4 * q + t) < b, re - (x + (4 * q + t)));
                    }
                }
            }
            while (x > a) { x--; BB_STEP(eq_r, BB_TEXT(x), x < b, re - x); }
            if (a == rs) {
                for (int t = 1; t <= m; t++) {
                    const int v = col_val<NW>(col.pv, col.mv, m - t) + __ldg(ov + t);
                    if (v <= k) emit_entry(A.entries, A.n_entries, A.cap, make_key(r, A.group, strand, n + t, v));
                }
            }
        }
    }
#undef BB_STEP
#undef BB_TEXT
}

// ---------------------------------------------------------------------------------------------------------------
// K2a: local-minimum rule on the sorted entries (oracle policy S1)
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_resolve(const uint64_t* __restrict__ keys, uint32_t n, const uint64_t* __restrict__ offsets,
                          const DevGroup* __restrict__ groups, uint8_t* __restrict__ flags) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint64_t key = keys[e];
    const uint64_t id = key >> kKeyPosShift;
    const int cost = static_cast<int>(key & 0xff);
    const uint32_t r = static_cast<uint32_t>(key >> kKeyReadShift);
    const int g = static_cast<int>((key >> kKeyGroupShift) & 7);
    const uint32_t pos = static_cast<uint32_t>((key >> kKeyPosShift) & ((1u << 28) - 1));
    const uint32_t len = static_cast<uint32_t>(offsets[r + 1] - offsets[r]);
    const uint32_t last = len + groups[g].m;   // the flank searcher always has the overhang extension
    // does the cost go up after this position?
    bool up = true;
    if (pos == last) up = true;
    else if (e + 1 < n && (keys[e + 1] >> kKeyPosShift) == id + 1) up = static_cast<int>(keys[e + 1] & 0xff) > cost;
    // was the last strict change before this position a decrease?  (missing neighbours are > k >= cost)
    bool dec = true;
    uint64_t cur = id;
    for (uint32_t q = e; q > 0; q--) {
        const uint64_t pk = keys[q - 1];
        if ((pk >> kKeyPosShift) != cur - 1) break;
        const int pc = static_cast<int>(pk & 0xff);
        if (pc > cost) break;
        if (pc < cost) { dec = false; break; }
        cur--;
    }
    flags[e] = (up && dec) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// K2b: traceback of the reported flank matches (oracle policies S2-S4) + barcode region
// ---------------------------------------------------------------------------------------------------------------
struct TraceArgs {
    const uint8_t* bases;
    const uint64_t* offsets;
    const uint64_t* hit_keys;
    uint32_t n_hits;
    const DevGroup* groups;
    uint64_t* hist;          // [col][2*NW][slot]
    uint32_t n_slots;
    Hit* hits;
};

template <int NW>
__device__ void trace_one(const TraceArgs& A, const DevGroup& G, uint32_t h, uint32_t slot) {
    const uint64_t key = A.hit_keys[h];
    const uint32_t r = static_cast<uint32_t>(key >> kKeyReadShift);
    const int g = static_cast<int>((key >> kKeyGroupShift) & 7);
    const int strand = static_cast<int>((key >> kKeyStrandShift) & 1);
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
            else if (cell(i, j - 1) + 1 == gcur) { di = 0; dj = 1; }
            else { di = 1; dj = 0; }
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
    for (uint32_t h = slot; h < A.n_hits; h += A.n_slots) {
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
    uint32_t n_hits;
    const DevGroup* groups;
    const uint8_t* code;      // [256] byte -> 4-bit IUPAC set
    Params prm;
    bb_row* rows;             // one slot per hit
    uint8_t* row_valid;
};

constexpr int kBarWarps = 4;
constexpr int kMaxBarRounds = 16;   // up to 512 barcodes per group

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

__global__ void __launch_bounds__(kBarWarps * 32) k_barcode(const BarArgs A) {
    __shared__ uint8_t s_codes[kBarWarps][kRegionMax];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t n_warps = gridDim.x * kBarWarps;
    for (uint32_t h = blockIdx.x * kBarWarps + wib; h < A.n_hits; h += n_warps) {
        const Hit H = A.hits[h];
        if (!H.has_region) { if (lane == 0) A.row_valid[h] = 0; continue; }   // searcher.rs:445-449
        const DevGroup& G = A.groups[H.group];
        const uint64_t rs0 = A.offsets[H.read];
        const int n = static_cast<int>(A.offsets[H.read + 1] - rs0);
        const int rn = H.re - H.rs;
        const int L = G.bar_len, nb = G.n_barcodes, k1 = G.k_bar, lb = L - 1;
        __syncwarp();
        for (int q = lane; q < rn; q += 32) s_codes[wib][q] = __ldg(A.code + A.bases[rs0 + H.rs + q]);
        __syncwarp();
        const uint8_t* codes = s_codes[wib];
        const uint64_t* eqs = G.bar_eq + static_cast<size_t>(H.strand) * nb * 16;
        const uint64_t pv_init = L >= 64 ? ~0ull : ((1ull << L) - 1ull);

        // pass A: bottom rows + S1 walk; best local minimum under k1 and under k = L (fallback)
        int16_t best1_pos[kMaxBarRounds], bestL_pos[kMaxBarRounds];
        int matched = 0;
#pragma unroll 1
        for (int rd = 0; rd * 32 < nb; rd++) {
            const int b = rd * 32 + lane;
            int p1 = -1, c1 = 1 << 20, pL = -1, cL = 1 << 20;
            if (b < nb) {
                const uint64_t* eq = eqs + static_cast<size_t>(b) * 16;
                Col<1> col; col.pv[0] = pv_init; col.mv[0] = 0;
                int prev = L, dec = 1;
                for (int p = 1; p <= rn; p++) {
                    const uint64_t e = __ldg(eq + codes[p - 1]);
                    const int cur = prev + col_step<1>(col, &e, lb);
                    if (cur > prev && dec) {
                        if (prev <= k1 && prev < c1) { c1 = prev; p1 = p - 1; }
                        if (prev < cL) { cL = prev; pL = p - 1; }
                    }
                    if (cur < prev) dec = 1; else if (cur > prev) dec = 0;
                    prev = cur;
                }
                if (dec) {
                    if (prev <= k1 && prev < c1) { c1 = prev; p1 = rn; }
                    if (prev < cL) { cL = prev; pL = rn; }
                }
            }
            best1_pos[rd] = static_cast<int16_t>(p1); bestL_pos[rd] = static_cast<int16_t>(pL);
            matched += __popc(__ballot_sync(0xffffffffu, p1 >= 0));
        }
        const bool fallback = matched <= 1 && k1 < L;          // searcher.rs:303-306

        // pass B: traceback + Lodhi of every candidate; per-lane top two under (score desc, index asc)
        double top_s = -1.0, sec_s = -1.0;
        int top_b = 1 << 30;
        int t_ok = 0, t_pi = 0, t_ei = 0, t_pj = 0, t_ej = 0, t_cost = 0, t_ts = 0, t_te = 0;
        int n_cand = 0;
#pragma unroll 1
        for (int rd = 0; rd * 32 < nb; rd++) {
            const int b = rd * 32 + lane;
            const int jend = b < nb ? (fallback ? bestL_pos[rd] : best1_pos[rd]) : -1;
            if (jend < 0) continue;
            n_cand++;
            const uint64_t* eq = eqs + static_cast<size_t>(b) * 16;
            uint64_t hpv[kRegionMax + 1], hmv[kRegionMax + 1];
            Col<1> col; col.pv[0] = pv_init; col.mv[0] = 0;
            hpv[0] = pv_init; hmv[0] = 0;
            for (int j = 0; j < jend; j++) {
                const uint64_t e = __ldg(eq + codes[j]);
                col_step<1>(col, &e, lb);
                hpv[j + 1] = col.pv[0]; hmv[j + 1] = col.mv[0];
            }
            // traceback (S2), no overhang: column 0 is walked with pattern-only steps
            uint64_t mbits[4] = {0, 0, 0, 0};                  // is-match bit of op q counted from the END of the path
            int n_ops = 0, i = L, j = jend;
            int cnt = 0, i_first = 0, i_last = 0, j_first = 0, j_last = 0, sub_cost = 0;
            while (i > 0) {
                int di = 1, dj = 0, is_match = 0;
                if (j > 0) {
                    const int gcur = col_val<1>(&hpv[j], &hmv[j], i), d = col_val<1>(&hpv[j - 1], &hmv[j - 1], i - 1);
                    const bool match = (__ldg(eq + codes[j - 1]) >> (i - 1)) & 1ull;
                    if (match && d == gcur) { dj = 1; is_match = 1; }
                    else if (d + 1 == gcur) { dj = 1; }
                    else if (col_val<1>(&hpv[j - 1], &hmv[j - 1], i) + 1 == gcur) { di = 0; dj = 1; }
                }
                i -= di; j -= dj;
                if (is_match && n_ops < 256) mbits[n_ops >> 6] |= 1ull << (n_ops & 63);
                n_ops++;
                if (i >= G.pbar0 && i < G.pbar1) {             // map_pat_to_text_with_cost range (cigar_parse.rs:22-30)
                    if (cnt == 0) { i_last = i; j_last = j; }
                    i_first = i; j_first = j;
                    sub_cost += !is_match;
                    cnt++;
                }
            }
            const int ts = j;
            // Lodhi S_3(C, 1/2), forward over the ops (same recurrence and order as orc_lodhi)
            double a1 = 0.0, a2 = 0.0, s = 0.0;
            for (int q = n_ops - 1; q >= 0; q--) {
                const bool mt = q < 256 && ((mbits[q >> 6] >> (q & 63)) & 1ull);
                if (mt) { s = s + 0.5 * a2; a2 = 0.5 * (a2 + a1); a1 = 0.5 * (a1 + 1.0); }
                else { a2 = 0.5 * a2; a1 = 0.5 * a1; }
            }
            const double sn = G.perfect > 0.0 ? s / G.perfect : 0.0;
            if (sn > top_s) {                                   // ascending b within a lane: strict > keeps the lower index
                sec_s = top_s;
                top_s = sn; top_b = b;
                t_ok = cnt > 0; t_pi = i_first; t_ei = i_last; t_pj = j_first; t_ej = j_last; t_cost = sub_cost;
                t_ts = ts; t_te = jend;
            } else if (sn > sec_s) sec_s = sn;
        }
        // warp reduction: global top (score desc, index asc), then the best of the rest
        double g_s = top_s; int g_b = top_b;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, g_s, off);
            const int ob = __shfl_xor_sync(0xffffffffu, g_b, off);
            if (os > g_s || (os == g_s && ob < g_b)) { g_s = os; g_b = ob; }
        }
        const bool owner = (g_b == top_b) && top_b != (1 << 30);
        double rest = owner ? sec_s : top_s;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, rest, off);
            if (os > rest) rest = os;
        }
        int total_cand = n_cand;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) total_cand += __shfl_xor_sync(0xffffffffu, total_cand, off);

        if (total_cand == 0) {
            if (lane == 0) { bb_row row; fill_flank_row(row, H, G, n); A.rows[h] = row; A.row_valid[h] = 1; }
            continue;
        }
        if (owner) {
            bool valid = g_s >= A.prm.min_score;                                   // searcher.rs:391-396
            if (total_cand > 1) valid = valid && (g_s - rest) >= A.prm.min_score_diff;
            bb_row row;
            if (valid && t_ok) {
                // to_path of the candidate with its strand overwritten by the flank's (S4, searcher.rs:333)
                int64_t pj, ej;
                if (H.strand == BB_FWD) { pj = t_pj; ej = t_ej; }
                else { pj = static_cast<int64_t>(t_te) - 1 - (t_pj - t_ts); ej = static_cast<int64_t>(t_te) - 1 - (t_ej - t_ts); }
                row.read_idx = H.read; row.read_len = static_cast<uint32_t>(n);
                row.rel_dist_to_end = rel_dist_to_end(H.text_start, n);
                row.read_start_bar = H.rs + pj; row.read_end_bar = H.rs + ej + 1;
                row.read_start_flank = H.text_start; row.read_end_flank = H.text_end;
                row.bar_start = H.rs + t_pi; row.bar_end = H.rs + t_ei + 1;
                row.flank_cost = H.cost; row.barcode_cost = t_cost; row.label_idx = g_b; row.group_idx = H.group;
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

__global__ void k_collapse(const Hit* __restrict__ hits, uint32_t n_hits, bb_row* rows, uint8_t* row_valid,
                           unsigned long long* kept_reads) {
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

// flank hit list for parity checks of the flank stage alone
__global__ void k_export_hits(const Hit* __restrict__ hits, uint32_t n_hits, int32_t* __restrict__ out6) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hits) return;
    const Hit H = hits[h];
    int32_t* o = out6 + 6 * static_cast<size_t>(h);
    o[0] = static_cast<int32_t>(H.read); o[1] = H.group; o[2] = H.strand; o[3] = H.text_start; o[4] = H.text_end; o[5] = H.cost;
}

}  // namespace bb
