// Barcode stage (K3), TEXT-ALONG-WORD formulation: the barcode text region of one flank match is short (<= 64 bases for
// every shipped kit, <= 160 by construction), so NWT = 1..3 64-bit words hold a whole DP ROW (bit j-1 = text column j) and
// the recurrence steps over the PATTERN rows -- reference searcher.rs:282-301 (best match per pattern), sassy's traceback,
// cigar-lodhi-rs' score and cigar_parse.rs:6-68 (map_pat_to_text_with_cost).  What the layout buys over pattern-along-word:
//   * the match masks are TEXT-position masks, one per pattern code, built ONCE per flank match for the whole warp
//     (tm[16]); a lane only carries its pattern's codes (one byte per row = offset into tm);
//   * the leading pattern rows that all barcodes of a strand share (the left pad: 10 rows forward, 8 reverse-complement
//     for the native kits) are computed once per flank match -- rows_prefix() -- and every barcode starts from that row
//     state; their traceback records are stored once per warp;
//   * all loops run over pattern rows, a warp-uniform count: no lane ever waits for another lane's longer loop, and the
//     two events of map_pat_to_text_with_cost are plain tests of the loop counter;
//   * one traceback record per row: diag = the path may leave cell (i, j) diagonally, stop = diag or it may NOT stay in the row;
//   * the pattern length is not tied to the word size.
// The routines are __host__ __device__: the CPU suite runs this very source lane by lane against the oracle
// (tests/test_barcode_rows.py); the product only runs it on the GPU.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD inline
#endif

namespace bb {

// search policies that the reference's own tests do not pin; 0 = the defaults.
// bb_opts.policy carries them across the ABI (BB_POL_* in include/barbell_b200.h).
constexpr int kPolS1Left = 1;       // S1: report the LEFT end of a cost plateau instead of the right end
constexpr int kPolS2PatFirst = 2;   // S2: traceback prefers a pattern-only step over a text-only step
constexpr int kPolS5Last = 4;       // S5: among equal lowest-cost minima of one pattern keep the LAST instead of the first
constexpr int kPolS6RcFirst = 8;    // S6: reverse-complement flank matches are listed before forward ones
constexpr int kPolS3Round = 16;     // S3: overhang cost of t hanging rows = round-to-nearest(t*alpha) instead of floor
constexpr int kPolS3Ceil = 32;      // S3: ... = ceil(t*alpha)
constexpr int kPolMask = 63;

// Pattern codes live as [row][lane] bytes (element r of a lane's pattern at offs[r * kOffStride]): the global table is
// [strand][round of 32 barcodes][row][lane], so a round's codes are one contiguous block that the warp copies into shared
// memory with a few 16-byte loads, and a warp's read of one row is 32 consecutive bytes (conflict-free).
constexpr int kOffStride = 32;
#ifndef BB_K3_UNROLL_F
#define BB_K3_UNROLL_F 4             // rows per iteration of the forward loop / of the traceback over a lane's own rows
#endif
#ifndef BB_K3_UNROLL_T
#define BB_K3_UNROLL_T 2
#endif
constexpr int kK3UnrollF = BB_K3_UNROLL_F, kK3UnrollT = BB_K3_UNROLL_T;
BB_HD uint32_t ld_code(const uint8_t* p) { return *p; }

BB_HD int bb_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll(static_cast<long long>(x));
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
BB_HD int bb_ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll(static_cast<long long>(x)) - 1;
#else
    return x ? __builtin_ctzll(x) : -1;
#endif
}
// (a & 0xffff) | (b << 16) in one byte-permute on the device
BB_HD uint32_t bb_pack16(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, 0x5410);
#else
    return (a & 0xffffu) | (b << 16);
#endif
}
BB_HD double bb_bits_to_double(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(static_cast<long long>(x));
#else
    double d; std::memcpy(&d, &x, 8); return d;
#endif
}
// double whose high word is h and whose low word is 0 (powers of two and 0.0 are built with integer multiplies: FMA pipe, not ALU)
BB_HD double bb_hi_to_double(uint32_t h) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(static_cast<int>(h), 0);
#else
    return bb_bits_to_double(static_cast<uint64_t>(h) << 32);
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

struct LaneAlign {
    double s;                           // Lodhi S_3(path, 1/2), not normalised
    int cbest, jend, ts;                // cost / end column of the chosen lowest-cost minimum; first text column of the path
    int cnt, i_first, i_last, j_first, j_last, sub_cost;   // map_pat_to_text_with_cost over pattern rows [pb0, pb1)
    int n_ops;
};

// One pattern row: (ph, mh) are the horizontal deltas of the previous row over the text columns; e = text columns that match
// the row's pattern character.  Transposed Myers/Hyyro step with vertical input +1 at column 0 (D[i][0] = i: the barcode
// searcher has no overhang).  Also returns the row's traceback record:
//   diag = match, or D[i-1][j-1] + 1 == D[i][j]  (<=> no match, the row above does not step down into column j and the column
//          to the left does not step down into row i)
//   stop = diag, or the preferred non-diagonal step leaves the row
//          (default S2, text-only before pattern-only: stop = diag | ~(D[i][j-1] + 1 == D[i][j]);
//           pattern-only first:                         stop = diag |  (D[i-1][j] + 1 == D[i][j]))
template <int NWT, bool S2PAT>
BB_HD void row_step(const uint64_t* e, uint64_t* ph, uint64_t* mh, uint64_t* diag, uint64_t* stop) {
    uint64_t carry = 0, pv_in = 1, mv_in = 0;
#pragma unroll
    for (int w = 0; w < NWT; w++) {
        const uint64_t E = e[w], Ph = ph[w], Mh = mh[w];
        const uint64_t xh = E | Mh;
        const uint64_t t = E & Ph;
        uint64_t sum = t + Ph;
        if constexpr (NWT > 1) {
            const uint64_t c1 = sum < t ? 1ull : 0ull;
            sum += carry;
            carry = c1 | (sum < carry ? 1ull : 0ull);
        }
        const uint64_t xv = (sum ^ Ph) | E;
        const uint64_t pv = Mh | ~(xv | Ph);
        const uint64_t mv = Ph & xv;
        const uint64_t pvs = pv + pv + pv_in, mvs = mv + mv + mv_in;   // shift by one column; the sum form lets one 3-input add do it
        if constexpr (NWT > 1) { pv_in = pv >> 63; mv_in = mv >> 63; }
        const uint64_t nph = mvs | ~(xh | pvs);
        const uint64_t d = E | ~(xh | mvs);
        diag[w] = d;
        if constexpr (S2PAT) stop[w] = d | pv;
        else stop[w] = d | ~nph;
        mh[w] = pvs & xh;
        ph[w] = nph;
    }
}

// traceback records of a lane's own rows: [row][32-bit word][lane] (conflict-free); one text word and a region of <= 48 bases
// (PACKED) needs 12 bytes per row and lane, otherwise 16 * NWT
template <int NWT, bool PACKED>
struct RowHist {
    static_assert(!PACKED || NWT == 1, "the packed record holds one text word");
    static constexpr int kWords = PACKED ? 3 : 4 * NWT;
    uint32_t* w;
    int lane;
    BB_HD void store(int row, const uint64_t* a, const uint64_t* b) const {
        uint32_t* q = w + row * (32 * kWords) + lane;
        if constexpr (PACKED) {
            q[0] = static_cast<uint32_t>(a[0]); q[32] = static_cast<uint32_t>(b[0]);
            q[64] = bb_pack16(static_cast<uint32_t>(a[0] >> 32), static_cast<uint32_t>(b[0] >> 32));
        } else {
#pragma unroll
            for (int t = 0; t < NWT; t++) {
                q[(4 * t) * 32] = static_cast<uint32_t>(a[t]); q[(4 * t + 1) * 32] = static_cast<uint32_t>(b[t]);
                q[(4 * t + 2) * 32] = static_cast<uint32_t>(a[t] >> 32); q[(4 * t + 3) * 32] = static_cast<uint32_t>(b[t] >> 32);
            }
        }
    }
    BB_HD void load(int row, uint64_t* a, uint64_t* b) const {
        const uint32_t* q = w + row * (32 * kWords) + lane;
        if constexpr (PACKED) {
            const uint32_t hi = q[64];
            a[0] = q[0] | (static_cast<uint64_t>(hi & 0xffffu) << 32);
            b[0] = q[32] | (static_cast<uint64_t>(hi >> 16) << 32);
        } else {
#pragma unroll
            for (int t = 0; t < NWT; t++) {
                a[t] = q[(4 * t) * 32] | (static_cast<uint64_t>(q[(4 * t + 2) * 32]) << 32);
                b[t] = q[(4 * t + 1) * 32] | (static_cast<uint64_t>(q[(4 * t + 3) * 32]) << 32);
            }
        }
    }
};
template <int NWT, bool PACKED>
BB_HD constexpr size_t row_hist_bytes(int rows) { return static_cast<size_t>(rows) * 32 * 4 * RowHist<NWT, PACKED>::kWords; }
// own-row records resident at a time: all of them, or the larger half with the meet-in-the-middle traceback
BB_HD constexpr int resident_rows(int own_rows, bool mitm) { return mitm ? (own_rows + 1) / 2 : own_rows; }

// match mask of a pattern row: offs = 8 * code, tm = [16][NWT]
template <int NWT>
BB_HD const uint64_t* row_mask(const uint64_t* tm, uint32_t off) {
    return reinterpret_cast<const uint64_t*>(reinterpret_cast<const unsigned char*>(tm) + off * NWT);
}

// The P leading rows every barcode of the strand shares, from the zero row: leaves (ph, mh) after row P and the rows'
// records in sh = [P][3][NWT] (diag, stop, e).  Every lane of the warp computes the same values; `write` selects the lane that stores.
template <int NWT, bool S2PAT>
BB_HD void rows_prefix(const uint64_t* tm, const uint8_t* offs, int P, bool write, uint64_t* sh, uint64_t* ph, uint64_t* mh) {
#pragma unroll
    for (int w = 0; w < NWT; w++) { ph[w] = 0; mh[w] = 0; }
    for (int r = 0; r < P; r++) {
        const uint64_t* e = row_mask<NWT>(tm, offs[r]);
        uint64_t ev[NWT], diag[NWT], stop[NWT];
#pragma unroll
        for (int w = 0; w < NWT; w++) ev[w] = e[w];
        row_step<NWT, S2PAT>(ev, ph, mh, diag, stop);
        if (write) {
#pragma unroll
            for (int w = 0; w < NWT; w++) { sh[(3 * r) * NWT + w] = diag[w]; sh[(3 * r + 1) * NWT + w] = stop[w]; sh[(3 * r + 2) * NWT + w] = ev[w]; }
        }
    }
}

// 64-bit shifts whose amount may reach (or, as an unsigned number, exceed) 64: the result is then 0 -- PTX semantics on the
// device, spelled out on the host.  They make column 0 of the traceback branch-free.
BB_HD uint64_t shl_clamp(uint64_t x, int s) {
#if defined(__CUDA_ARCH__)
    uint64_t r; asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(s)); return r;
#else
    return static_cast<unsigned>(s) > 63u ? 0ull : x << s;
#endif
}
BB_HD uint64_t shr_clamp(uint64_t x, int s) {
#if defined(__CUDA_ARCH__)
    uint64_t r; asm("shr.u64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(s)); return r;
#else
    return static_cast<unsigned>(s) > 63u ? 0ull : x >> s;
#endif
}

// Bottom-row scan table: entry (p | m << 4) describes four columns whose horizontal deltas are +1 where p, -1 where m:
//   low half  = 128 * (lowest partial sum) + (first column 1..4 that reaches it)      -- relative to the nibble's start
//   high half = 128 * (sum of the four deltas) + 4
// so that with keys K = 128 * cost + column the first lowest-cost column is a running signed minimum.
BB_HD uint32_t scan_lut_entry(int idx) {
    const int p = idx & 15, m = idx >> 4;
    int run = 0, mn = 1 << 20, arg = 0;
    for (int b = 0; b < 4; b++) {
        run += ((p >> b) & 1) - ((m >> b) & 1);
        if (run < mn) { mn = run; arg = b + 1; }
    }
    return (static_cast<uint32_t>(mn * 128 + arg) & 0xffffu) | (static_cast<uint32_t>(run * 128 + 4) << 16);
}

// tm    : [16][NWT] text-column masks indexed by the 4-bit IUPAC set of a PATTERN character (bit j-1 = region base j matches)
// offs  : this lane's pattern, one byte per row = 8 * code, row r at offs[r * kOffStride]
// rn, L : region bases (<= 64 * NWT), pattern rows; P = rows already done by rows_prefix(), (ph, mh) = its row state
// sh    : the records of the shared rows; hist: this lane's records of rows P+1..L; rec: per-row traceback records, rec[row * 32]
// lut   : [256] scan_lut_entry()
// pol   : kPolS1Left | kPolS5Last (S2 is the template parameter: it changes what the forward pass stores)
// MITM : meet in the middle -- shared memory per alignment is what bounds the warps in flight, so only HALF of the own rows'
//         records are resident at a time: the forward pass stores rows > H, the traceback walks those, then the forward
//         recurrence is replayed from the shared row state over rows P+1..H -- this time storing -- and the traceback goes on.
template <int NWT, bool PACKED, bool S2PAT, bool MITM>
BB_HD void rows_lane(const uint64_t* tm, const uint8_t* offs, int rn, int L, int P, const uint64_t* ph0, const uint64_t* mh0, const uint64_t* sh,
                     const RowHist<NWT, PACKED>& hist, uint8_t* rec, const uint32_t* lut, int pb0, int pb1, int pol, LaneAlign& O) {
    constexpr int kRowWords = 32 * RowHist<NWT, PACKED>::kWords;
    uint64_t ph[NWT], mh[NWT];
#pragma unroll
    for (int w = 0; w < NWT; w++) { ph[w] = ph0[w]; mh[w] = mh0[w]; }
    // ---- forward pass over the lane's own rows, storing the records of rows > H (all own rows without MITM) ----
    const int H = MITM ? P + ((L - P) >> 1) : P;
    auto forward = [&](int r0, int r1, bool store, int slot0) {    // rows r0+1 .. r1 of the forward recurrence
        RowHist<NWT, PACKED> hp = hist;
        hp.w += slot0 * kRowWords;
        const uint8_t* cp = offs + r0 * kOffStride;
#pragma unroll kK3UnrollF
        for (int r = r0; r < r1; r++, cp += kOffStride) {
            const uint64_t* e = row_mask<NWT>(tm, ld_code(cp));
            uint64_t ev[NWT], diag[NWT], stop[NWT];
#pragma unroll
            for (int w = 0; w < NWT; w++) ev[w] = e[w];
            row_step<NWT, S2PAT>(ev, ph, mh, diag, stop);
            if (store) { hp.store(0, diag, stop); hp.w += kRowWords; }
        }
    };
    if (H > P) forward(P, H, false, 0);
    forward(H, L, true, 0);
    // ---- bottom row: D[L][j] = L + sum of the horizontal deltas.  S1 reports the plateaus that follow a decrease and are followed
    //      by an increase (or the end); "lowest cost, first seen" (searcher.rs:294-300) = the FIRST plateau at the global minimum ----
    int cbest, jend;
    if (NWT == 1 && !(pol & kPolS5Last)) {
        // four columns per step through the table: keys 128 * cost + column, running minimum = first lowest-cost column
        const uint64_t keep = rn < 64 ? ((1ull << rn) - 1ull) : ~0ull;
        const uint64_t pk = ph[0] & keep, mk = mh[0] & keep;
        int K = L * 128, best = K;
        const int nq = (rn + 3) >> 2;
        uint32_t pw = static_cast<uint32_t>(pk), mw = static_cast<uint32_t>(mk);
        for (int q = 0; q < nq; q++) {
            if (q == 8) { pw = static_cast<uint32_t>(pk >> 32); mw = static_cast<uint32_t>(mk >> 32); }
            const uint32_t v = lut[(pw & 15u) | ((mw & 15u) << 4)];
            pw >>= 4; mw >>= 4;
            const int cand = K + static_cast<int16_t>(v & 0xffffu);
            best = cand < best ? cand : best;
            K += static_cast<int32_t>(v) >> 16;
        }
        cbest = best >> 7; jend = best & 127;
    } else {
        int cur = L, pf = 0, pl = 0;
        cbest = L;
#pragma unroll
        for (int w = 0; w < NWT; w++) {
            const int lim = rn - 64 * w < 64 ? rn - 64 * w : 64;
            for (int b = 0; b < lim; b++) {
                const int dn = static_cast<int>((mh[w] >> b) & 1ull);
                cur += static_cast<int>((ph[w] >> b) & 1ull) - dn;
                const int j = 64 * w + b + 1;
                if (cur < cbest) { cbest = cur; pf = j; pl = j; }
                else if (dn && cur == cbest) pl = j;
            }
        }
        jend = (pol & kPolS5Last) ? pl : pf;                 // left end of the chosen plateau
    }
    if (!(pol & kPolS1Left)) {                               // ... its right end: the columns that follow with a zero delta
        if constexpr (NWT == 1) {
            const uint64_t nz = shr_clamp(ph[0] | mh[0], jend);
            const int run = nz ? bb_ctz64(nz) : 64;
            jend += run < rn - jend ? run : rn - jend;       // the region ends at column rn
        } else {
            while (jend < rn && !(((ph[jend >> 6] | mh[jend >> 6]) >> (jend & 63)) & 1ull)) jend++;
        }
    }
    // ---- traceback from (L, jend), one ROW per iteration ----
    // In row i the path walks left over t text-only steps to the nearest column whose stop bit is set and leaves the row
    // there, diagonally (diag bit: match or substitution) or upwards (pattern-only); at column 0 only "up" is left.
    // Path entries (to_path: position BEFORE the op) with pattern index r are the leaving op of row r+1 and the text-only
    // steps of row r, so over [pb0, pb1):
    //   last entry  = the leaving op of row pb1,
    //   first entry = the last text-only step of row pb0 if there is one (pb0 >= 1), else the leaving op of row pb0+1,
    // and with T = text-only steps so far, M = matches so far the non-match ops between two points of the traceback are
    // (T2 - T1) + (rows between) - (M2 - M1).  The loop is cut at rows pb1, pb0+1 and pb0, so the common rows carry no event tests.
    int j = jend, T = 0, M = 0;
    int j_first = 0, j_last = 0, T_a = 0, M_a = 0, T_b = 0, M_b = 0;
    double a1 = 0.0, a2 = 0.0, s = 0.0;                      // Lodhi accumulators over the REVERSED op sequence
    auto row = [&](int i, const uint64_t* diag, const uint64_t* stop, const uint64_t* e, bool ev_a, bool ev_b1, bool ev_b2) {
        int t, is_diag, is_match;
        if constexpr (NWT == 1) {
            const int lz = bb_clz64(shl_clamp(stop[0], 64 - j));          // bit 63 = column j; j == 0 gives 0
            t = lz < j ? lz : j;                                          // no stop bit at or below column j: walk to column 0
            j -= t;
            is_diag = static_cast<int>(shr_clamp(diag[0], j - 1) & 1ull);   // j == 0: no diagonal, no match
            is_match = static_cast<int>(shr_clamp(e[0], j - 1) & 1ull);     // a matching cell always has its diag bit set
        } else {
            int jj = j;
            while (jj > 0 && !((stop[(jj - 1) >> 6] >> ((jj - 1) & 63)) & 1ull)) jj--;
            t = j - jj; j = jj; is_diag = 0; is_match = 0;
            if (j > 0) {
                is_diag = static_cast<int>((diag[(j - 1) >> 6] >> ((j - 1) & 63)) & 1ull);
                is_match = static_cast<int>((e[(j - 1) >> 6] >> ((j - 1) & 63)) & 1ull);
            }
        }
        T += t;
        if (ev_a) { T_a = T; M_a = M; }
        if (ev_b2 && t > 0) { T_b = T; M_b = M; j_first = j; }
        M += is_match;
        j -= is_diag;
        if (ev_a) j_last = j;
        if (ev_b1) { T_b = T; M_b = M; j_first = j; }
        rec[(i - 1) * 32] = static_cast<uint8_t>((t << 1) | is_match);   // t <= path cost <= L <= 64
        // reversed op order: t non-match ops, then the leaving op; g = 2^-(t+1)
        const uint32_t gh = static_cast<uint32_t>(1022 - t) << 20, um = static_cast<uint32_t>(is_match);
        const double g = bb_hi_to_double(gh);
        s = bb_fma(bb_hi_to_double(gh * um), a2, s);                      // is_match ? g : 0
        a2 = g * bb_fma(bb_hi_to_double(0x3ff00000u * um), a1, a2);       // is_match ? 1 : 0
        a1 = bb_fma(g, a1, bb_hi_to_double(0x3fe00000u * um));            // is_match ? 1/2 : 0
    };
    auto shared_row = [&](int i, bool ev_a, bool ev_b1, bool ev_b2) {
        const uint64_t* q = sh + static_cast<size_t>(3 * (i - 1)) * NWT;
        row(i, q, q + NWT, q + 2 * NWT, ev_a, ev_b1, ev_b2);
    };
    const bool in_range = pb1 > pb0 && pb0 >= 0 && pb1 <= L;
    int base = H;                                            // own row i sits in slot i - 1 - base
    for (int i = L; i >= 1;) {
        if (MITM && i == H && base == H && H > P) {          // the traceback arrives at the lower half: replay it, storing
#pragma unroll
            for (int w = 0; w < NWT; w++) { ph[w] = ph0[w]; mh[w] = mh0[w]; }
            forward(P, H, true, 0);
            base = P;
        }
        RowHist<NWT, PACKED> hp = hist;
        hp.w += (i - 1 - base) * kRowWords;
        const uint8_t* cp = offs + (i - 1) * kOffStride;
        if (in_range && (i == pb1 || i == pb0 + 1 || i == pb0)) {        // the three rows with an event (warp-uniform)
            if (i > P) {
                uint64_t diag[NWT], stop[NWT];
                hp.load(0, diag, stop);
                row(i, diag, stop, row_mask<NWT>(tm, ld_code(cp)), i == pb1, i == pb0 + 1, i == pb0);
            } else shared_row(i, i == pb1, i == pb0 + 1, i == pb0);
            i--;
            continue;
        }
        int lo = 1;                                          // plain rows down to the next event row / the next change of storage
        if (in_range) lo = i > pb1 ? pb1 + 1 : i > pb0 + 1 ? pb0 + 2 : 1;
        if (MITM && i > H && lo <= H) lo = H + 1;
        if (i > P) {
            if (lo <= P) lo = P + 1;
#pragma unroll kK3UnrollT
            for (; i >= lo; i--, cp -= kOffStride, hp.w -= kRowWords) {
                uint64_t diag[NWT], stop[NWT];
                hp.load(0, diag, stop);
                row(i, diag, stop, row_mask<NWT>(tm, ld_code(cp)), false, false, false);
            }
        } else {
#pragma unroll 2
            for (; i >= lo; i--) shared_row(i, false, false, false);
        }
    }
    const int n_ops = L + T;
    if (!lodhi_exact(s, n_ops)) {
        // same recurrence and op order as the reference's forward pass (a row's record in path order: the op that enters the
        // row, then its t text-only ops)
        a1 = 0.0; a2 = 0.0; s = 0.0;
        for (int q = 0; q < L; q++) {
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
    O.s = s; O.cbest = cbest; O.jend = jend; O.ts = j;
    O.cnt = in_range ? 1 : 0; O.i_first = pb0; O.i_last = pb1 - 1; O.j_first = j_first; O.j_last = j_last;
    O.sub_cost = (T_b - T_a) + (pb1 - pb0) - (M_b - M_a);
    O.n_ops = n_ops;
}

}  // namespace bb
