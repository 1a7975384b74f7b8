// Device-side view of the pattern sets and the intermediate records that flow between the kernels.
#pragma once
#include <cstdint>

namespace bb {

constexpr int kMaxGroups = 8;        // 3 key bits
constexpr int kMaxFlankWords = 2;    // flank patterns up to 128 characters
constexpr int kMaxBarLen = 64;       // padded barcode patterns fit one 64-bit word
constexpr int kRegionMax = 160;      // longest barcode text region (mask + k_flank + 2*PADDING)
constexpr int kPadding = 10;         // reference src/lib.rs:10

// sort key of one sub-threshold end position produced by the flank scan
//   [63:40] read (24 bit) | [39:37] group | [36] strand | [35:8] end position in the strand's frame | [7:0] cost
constexpr int kKeyReadShift = 40, kKeyGroupShift = 37, kKeyStrandShift = 36, kKeyPosShift = 8;
constexpr uint32_t kMaxBatchReads = 1u << 24;
constexpr uint32_t kMaxReadLen = (1u << 28) - 512;

struct DevGroup {
    int m, nw, last_bit, k;              // flank length, 64-bit words, bit of the last row in the last word, k_flank
    int bar0, bar1, pad0, pad1;          // reference barcodes.rs:160-192
    int bar_len, n_barcodes, match_type, k_bar;   // k_bar = (int)(bar_len * 0.4f)   searcher.rs:460
    int pbar0, pbar1;                    // bar0 - pad0, bar1 - pad0                   searcher.rs:379-382
    int ov_m, halo;                      // floor(m*alpha); shared-memory halo of the scan tile (multiple of 16, >= warm)
    int warm, pad2_;                     // warm-up columns of a scan chunk: m + k rounded up to the scan group size
    int trace_cols, pad_;                // m + 2k + 8 (+1 columns of history)
    double perfect;                      // Lodhi of pad1-pad0 matches                 searcher.rs:229-239
    uint64_t pv_plain[kMaxFlankWords];   // first column D[i] = i
    uint64_t pv_over[kMaxFlankWords];    // first column D[i] = floor(i*alpha)
    uint64_t pv_plain_top[kMaxFlankWords];   // the same two columns with the pattern moved to the top of the bit-vector
    uint64_t pv_over_top[kMaxFlankWords];    //   (row i at bit i + 64*nw - m; the low bits are wildcard rows, zero deltas)
    const uint64_t* eq;                  // [2 strands][256 bytes][nw]  flank match masks indexed by the raw text byte
    const uint64_t* eq_top;              // same, top-aligned, wildcard rows below (used by the scan kernel)
    const int* ov;                       // [m+1] floor(t*alpha)
    const uint8_t* bar_off;              // [2 strands][rounds of 32 barcodes][64 rows][32 lanes]: 8 * (4-bit IUPAC set) of every pattern row (Rc: of the reverse complement)
    const uint8_t* sh_off;               // [2 strands][64]: the same for the leading rows all barcodes of the strand share
    int sh_p[2], pol, pad3_;             // number of those rows per strand; search policy bits (barcode_rows.cuh kPol*)
    // lossless pre-filter (0 = off): rows [f_q0, f_q0 + f_q) of the flank are an N-free run with f_q <= 15 and 3k <= f_q
    int f_on, f_q, f_q0, f_pad;
    const uint32_t* f_eq;                // [256] bits [0,q): the run; bits [16,16+q): its reverse complement (both vs the forward text)
    // second N-free run S = rows [f_s0, f_s0 + f_qs) (0 rows = none): pre-check of the filter's candidates
    int f_qs, f_s0, f_pad2, f_pad3;
    const uint32_t* f_seq;               // [256] same layout for S
};

// candidate window of the pre-filter / read-end window: end positions [lo, lo+len] of one strand's frame to verify exactly
//   [63:40] read | [39] strand | [38:11] lo (28 bit) | [10:0] len
constexpr int kWinReadShift = 40, kWinStrandShift = 39, kWinLoShift = 11;

struct Params {
    double min_score, min_score_diff;
    int n_groups;
};

// one reported flank match (sassy Match of the overhang searcher) + the barcode text region
struct Hit {
    uint32_t read;
    int32_t group, strand;
    int32_t text_start, text_end, cost;  // forward text coordinates
    int32_t rs, re;                      // padded barcode region [rs, re) in forward text coordinates
    int32_t has_region;                  // get_matching_region returned Some
};

}  // namespace bb
