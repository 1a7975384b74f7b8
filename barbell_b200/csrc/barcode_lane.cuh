// Per-lane body of the barcode stage (K3): ONE padded barcode pattern against the barcode text region of one flank
// match -- reference searcher.rs:282-301 (best match per pattern), sassy's traceback, cigar-lodhi-rs' score and
// cigar_parse.rs:6-68 (map_pat_to_text_with_cost).  The function is __host__ __device__ so that the CPU test-suite can
// run the very same code lane by lane against the oracle (tests/test_barcode_lane.py); the product only runs it on the GPU.
//
//  forward pass  one top-aligned 64-bit bit-vector column step per region base; walks sassy's local-minimum rule (S1)
//                online and stores, for EVERY column, the two bit-vectors the traceback needs:
//                   diag[i] = the path may leave cell (i, j) diagonally (match, or D[i-1][j-1] + 1 == D[i][j])
//                   stop[i] = diag[i] or it may leave horizontally (D[i][j-1] + 1 == D[i][j])
//                ([column][word][lane] in shared memory, 12 bytes per column when the pattern has <= 48 rows).
//  traceback     (S2: match > substitution > text-only > pattern-only) ONE COLUMN per iteration and branch-free: the
//                highest set bit of stop at or below the current row (count-leading-zeros) is the row at which the path
//                leaves the column, the rows skipped on the way are pattern-only steps; no cost value is materialised
//                and nothing is recomputed -- three LDS for the column plus the match mask of its base.
//  history       shared memory per alignment bounds the warps in flight, so only HALF of the columns are resident at a time
//                (meet in the middle: the upper half is stored by the forward pass, the lower half by a replay of the forward
//                recurrence when the traceback arrives there).
//  Lodhi score   S_3(C, 1/2) is accumulated INSIDE the traceback loop in reversed op order.  The score is a sum of
//                powers of two over triples of match ops, symmetric under reversal; while n_ops + log2(score) <= 53 every
//                partial sum of either order is exactly representable in f64 (lodhi_exact), so the reversed accumulation
//                returns the same bits as the reference's forward recurrence.  Longer paths (many inserted bases; rare)
//                replay the per-column records in path order with the forward recurrence.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD inline
#endif

namespace bb {

BB_HD int bb_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll(static_cast<long long>(x));
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
BB_HD double bb_bits_to_double(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(static_cast<long long>(x));
#else
    double d; std::memcpy(&d, &x, 8); return d;
#endif
}
BB_HD double bb_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// The score's partial sums (in either op order) are multiples of 2^-n_ops below the final score: with s < 2^b they need b + n_ops
// mantissa bits, so the computation is exact in f64 whenever n_ops + b <= 53.
BB_HD bool lodhi_exact(double s, int n_ops) {
    const double infl = s * 1.0000000000009095;              // (1 + 2^-40): rounding, had it happened, could not hide a power of two
    uint64_t bits;
#if defined(__CUDA_ARCH__)
    bits = static_cast<uint64_t>(__double_as_longlong(infl));
#else
    std::memcpy(&bits, &infl, 8);
#endif
    int b = static_cast<int>((bits >> 52) & 0x7ff) - 1022;   // s < 2^b
    if (b < 1) b = 1;
    return n_ops <= 52 && n_ops + b <= 53;
}

template <bool PACKED>
struct ColHist {                        // [column][word][lane] so that a warp's accesses are conflict-free
    uint32_t* w;
    int lane;
    BB_HD void store(int col, uint64_t a, uint64_t b) const {
        if constexpr (PACKED) {        // rows live in bits [16, 64): the low 16 bits are wildcard rows and never read
            uint32_t* q = w + static_cast<size_t>(col) * 96 + lane;
            q[0] = static_cast<uint32_t>(a >> 32); q[32] = static_cast<uint32_t>(b >> 32);
            q[64] = (static_cast<uint32_t>(a) >> 16) | (static_cast<uint32_t>(b) & 0xffff0000u);
        } else {
            uint64_t* q = reinterpret_cast<uint64_t*>(w) + static_cast<size_t>(col) * 64 + lane;
            q[0] = a; q[32] = b;
        }
    }
    BB_HD void load(int col, uint64_t& a, uint64_t& b) const {
        if constexpr (PACKED) {
            const uint32_t* q = w + static_cast<size_t>(col) * 96 + lane;
            const uint32_t lo = q[64];
            a = (static_cast<uint64_t>(q[0]) << 32) | (lo << 16);
            b = (static_cast<uint64_t>(q[32]) << 32) | (lo & 0xffff0000u);
        } else {
            const uint64_t* q = reinterpret_cast<const uint64_t*>(w) + static_cast<size_t>(col) * 64 + lane;
            a = q[0]; b = q[32];
        }
    }
};

struct LaneAlign {
    double s;                           // Lodhi S_3(path, 1/2), not normalised
    int cbest, jend, ts;                // cost / end column of the first lowest-cost minimum; first text column of the path
    int cnt, i_first, i_last, j_first, j_last, sub_cost;   // map_pat_to_text_with_cost over pattern rows [pb0, pb1)
    int n_ops;
};

// Region bases are staged as one byte each: low nibble = 4-bit IUPAC base set, high nibble = slot of the set in the lane's
// match-mask table (A, C, G, T, N -> 0..4; every other set -> kEqOther, resolved by OR-ing the masks of its bases).
constexpr int kEqSlots = 5, kEqOther = 5;
BB_HD uint8_t region_byte(uint8_t code) {
    const int slot = code == 1 ? 0 : code == 2 ? 1 : code == 4 ? 2 : code == 8 ? 3 : code == 15 ? 4 : kEqOther;
    return static_cast<uint8_t>(code | (slot << 4));
}
BB_HD uint64_t region_mask(const uint64_t* eq, uint32_t v, uint64_t wild) {
    const uint32_t slot = v >> 4;
    if (slot < static_cast<uint32_t>(kEqOther)) return eq[slot * 32];
    uint64_t e = wild;                                       // rare: an ambiguity code other than N, or a non-IUPAC byte (empty set)
    for (int b = 0; b < 4; b++) if ((v >> b) & 1u) e |= eq[b * 32];
    return e;
}

// eq      : this lane's match masks of the sets {A}, {C}, {G}, {T}, {ACGT}, top-aligned with wildcard rows below: eq[slot * 32]
// txt     : the region's bases as region_byte() values (shared by the warp), rn of them
// hist    : this lane's column history, (rn + 1) / 2 columns: shared memory per alignment is what bounds the warps in flight,
//           so only HALF of the columns are ever resident (see below)
// rec     : this lane's per-column traceback records, rec[q * 32]
//
// FAST variant (the common case): every base of the region is A, C, G, T or N -- no per-column check of the base set.
//
// Meet-in-the-middle history.  With H = rn / 2: the forward pass runs over all columns (it needs the whole bottom row for the
// minima) but stores (diag, stop) for the columns > H only; the traceback walks those; when it arrives at column H the forward
// recurrence is replayed over columns 1..H -- this time storing -- and the traceback continues.  Half the shared memory for
// ~0.5 extra forward columns per column.
//
// map_pat_to_text_with_cost (cigar_parse.rs:6-68) over pattern rows [pb0, pb1): a path visits every pattern row, rows never
// increase along the traceback, so the first entry of the range is the first visit of row pb1-1 and the last entry of the range
// is the entry just before the first visit of row pb0-1 (or the path's last entry): two events, each fires once, and the
// edits in the range are the difference of the running non-match count at the two events.
template <bool PACKED, bool FAST>
BB_HD void barcode_lane(const uint64_t* eq, const uint8_t* txt, int rn, int L, int pb0, int pb1,
                        const ColHist<PACKED>& hist, uint8_t* rec, LaneAlign& O) {
    const int sh = 64 - L;                                   // row i of the pattern is bit i + sh
    const uint64_t wild = sh ? ((1ull << sh) - 1ull) : 0ull;
    const uint64_t pv0 = (L >= 64 ? ~0ull : ((1ull << L) - 1ull)) << sh;
    const int H = rn >> 1;
    auto mask_of = [&](int col) -> uint64_t {                // match mask of region base `col`
        if constexpr (FAST) return eq[static_cast<uint32_t>(txt[col] >> 4) * 32];
        else return region_mask(eq, txt[col], wild);
    };
    // ---- forward pass: the minima of the bottom row (S1) + column history of the upper half ----
    // S1 with every minimum reported (k = len) and "lowest cost, first seen" (searcher.rs:294-300) picks the right end of
    // the FIRST plateau that reaches the global minimum of the bottom row: `open` = still on that plateau.
    uint64_t pv = pv0, mv = 0;
    int cur = L, cbest = L, jend = 0, open = 1;
    uint64_t e_next = rn > 0 ? mask_of(0) : 0;
#define BB_COLUMN(P, STORE, SLOT, MINIMA)                                                                  \
    {                                                                                                      \
        const uint64_t e = e_next;                                                                         \
        if ((P) < rn) e_next = mask_of(P);                                                                 \
        const uint64_t sum = (e & pv) + pv;                                                                \
        uint64_t ph = mv | ~(sum | pv | e);                  /* horizontal deltas between columns P-1 and P */ \
        uint64_t mh = pv & ((sum ^ pv) | e);                                                               \
        if (STORE) {                                                                                       \
            const uint64_t diag = e | (ph & ~(pv | mv)) | (pv & ~(ph | mh));   /* match, or D[i-1][P-1] + 1 == D[i][P] */ \
            hist.store((SLOT), diag, diag | ph);                               /* ... else text-only if D[i][P-1] + 1 == D[i][P] */ \
        }                                                                                                  \
        if (MINIMA) cur += static_cast<int>(ph >> 63) - static_cast<int>(mh >> 63);                        \
        ph <<= 1; mh <<= 1;                                                                                \
        pv = mh | ~(e | mv | ph);                                                                          \
        mv = ph & (e | mv);                                                                                \
        if (MINIMA) {                                                                                      \
            open = cur < cbest ? 1 : (cur == cbest ? open : 0);                                            \
            cbest = cur < cbest ? cur : cbest;                                                             \
            jend = open ? (P) : jend;                                                                      \
        }                                                                                                  \
    }
    for (int p = 1; p <= H; p++) BB_COLUMN(p, false, 0, true)
    for (int p = H + 1; p <= rn; p++) BB_COLUMN(p, true, p - 1 - H, true)
    // ---- traceback (S2) from (L, jend), one column per iteration; column 0 is walked with pattern-only steps ----
    const int R1 = pb1 - 1, Rm = pb0 - 1;                    // first row of the range from above / first row below the range
    int i = L, j = jend, nrec = 0, n_ops = 0, nm = 0;        // nm: non-match ops so far
    int j_first = 0, j_last = 0, nm_a = 0, nm_b = 0, seen_b = 0;
    double a1 = 0.0, a2 = 0.0, s = 0.0;                      // Lodhi accumulators over the REVERSED op sequence
    bool to_row0 = false;                                    // the rest of the path is pattern-only steps in column j
    // walks the columns j > j_lo whose vectors sit in history slots (column - 1 - off)
    auto trace = [&](int j_lo, int off) {
        while (i > 0 && j > j_lo) {
            uint64_t diag, stop;
            hist.load(j - 1 - off, diag, stop);
            const uint64_t e = mask_of(j - 1);
            // move row i (bit i-1+sh) to bit 63: the leading zeros of stop are the pattern-only steps taken in this column
            const int d = bb_clz64(stop << (L - i));
            if (d >= i) { to_row0 = true; break; }           // no row at or below i lets the path out: pattern-only to row 0
            const int il = i - d;                            // the path leaves the column at row il ...
            const int t = il - 1 + sh;                       // ... whose bit this is
            const int is_diag = static_cast<int>((diag >> t) & 1ull), is_match = static_cast<int>((e >> t) & 1ull);
            const int ia = il - is_diag;                     // pre-op row of the leaving op; rows [ia, i-1] are visited here
            if (i > R1 && ia <= R1) { j_last = il <= R1 ? j : j - 1; const int a = i - 1 - R1; nm_a = nm + (a < d ? a : d); }
            if (i > Rm && ia <= Rm) { j_first = j; const int a = i - 1 - Rm; nm_b = nm + (a < d ? a : d); seen_b = 1; }
            nm += d + 1 - is_match;
            n_ops += d + 1;
            rec[nrec * 32] = static_cast<uint8_t>((d << 1) | is_match);
            nrec++;
            i = ia; j = j - 1;
            // reversed op order: d non-match ops, then the leaving op; g = 2^-(d+1)
            const double g = bb_bits_to_double(static_cast<uint64_t>(1022 - d) << 52);
            const double mm = is_match ? 1.0 : 0.0;
            s = bb_fma(is_match ? g : 0.0, a2, s);
            a2 = g * bb_fma(mm, a1, a2);
            a1 = bb_fma(g, a1, 0.5 * mm);
        }
    };
    trace(H, H);
    if (!to_row0 && i > 0 && j > 0) {                        // arrived at column j <= H: replay the forward recurrence, storing
        pv = pv0; mv = 0;
        e_next = mask_of(0);
        const int jj = j;
        for (int p = 1; p <= jj; p++) BB_COLUMN(p, true, p - 1, false)
        trace(0, 0);
    }
#undef BB_COLUMN
    if (i > 0) {                                             // leading pattern-only steps at column j: rows i-1 .. 0, all non-match
        if (i > R1 && R1 >= 0) { j_last = j; nm_a = nm + (i - 1 - R1); }
        if (i > Rm && Rm >= 0) { j_first = j; nm_b = nm + (i - 1 - Rm); seen_b = 1; }
        nm += i;
        n_ops += i;
    }
    if (!seen_b) { j_first = j; nm_b = nm; }                 // the range reaches row 0: its last entry is the path's last entry
    if (!lodhi_exact(s, n_ops)) {
        // same recurrence and op order as the reference's forward pass (leading non-match ops act on zeros)
        a1 = 0.0; a2 = 0.0; s = 0.0;
        for (int q = nrec - 1; q >= 0; q--) {
            const int r = rec[q * 32];
            if (r & 1) { s = s + 0.5 * a2; a2 = 0.5 * (a2 + a1); a1 = 0.5 * (a1 + 1.0); }
            else { a2 = 0.5 * a2; a1 = 0.5 * a1; }
            const int d = r >> 1;
            if (d) {                                         // d non-match ops = exact scaling by 2^-d
                const double f = bb_bits_to_double(static_cast<uint64_t>(1023 - d) << 52);
                a2 = a2 * f; a1 = a1 * f;
            }
        }
    }
    const bool in_range = pb1 > pb0 && pb0 >= 0 && pb1 <= L;
    O.s = s; O.cbest = cbest; O.jend = jend; O.ts = j;
    O.cnt = in_range ? 1 : 0; O.i_first = pb0; O.i_last = R1; O.j_first = j_first; O.j_last = j_last; O.sub_cost = nm_b - nm_a;
    O.n_ops = n_ops;
}

}  // namespace bb
