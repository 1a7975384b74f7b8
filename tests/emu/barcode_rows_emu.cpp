// Host build of the text-along-word barcode routine (barbell_b200/csrc/barcode_rows.cuh) for the CPU test-suite: the SAME
// source the GPU kernel compiles, run for one lane of an emulated warp.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../barbell_b200/csrc/barcode_rows.cuh"

namespace {
uint8_t g_code[256];
bool g_ready = false;
void init_codes() {
    if (g_ready) return;
    const char* L = "ACGTURYSWKMBDHVN";
    const uint8_t V[] = {1, 2, 4, 8, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15};
    for (int i = 0; L[i]; i++) { g_code[static_cast<uint8_t>(L[i])] = V[i]; g_code[static_cast<uint8_t>(L[i] | 0x20)] = V[i]; }
    g_ready = true;
}
template <int NWT, bool PACKED, bool S2PAT, bool MITM>
bool run(const uint8_t* region, const uint8_t* offs1, int rn, int L, int P, int lane, int pb0, int pb1, int pol, bb::LaneAlign& R) {
    uint64_t tm[16 * NWT];
    for (int c = 0; c < 16; c++)
        for (int w = 0; w < NWT; w++) {
            uint64_t v = 0;
            for (int q = 64 * w; q < rn && q < 64 * (w + 1); q++) if (g_code[region[q]] & c) v |= 1ull << (q & 63);
            tm[c * NWT + w] = v;
        }
    std::vector<uint8_t> offs_t(66 * bb::kOffStride, 0x00);                 // (+ 2 rows: the forward pass loads two rows ahead)                 // the kernel's table layout: [row][lane]
    for (int r = 0; r < 64; r++) offs_t[r * bb::kOffStride + lane] = offs1[r];
    const uint8_t* offs = offs_t.data() + lane;
    std::vector<uint64_t> sh(3 * 64 * NWT + 8, 0x5a5a5a5a5a5a5a5aull);
    uint64_t ph[NWT], mh[NWT];
    bb::rows_prefix<NWT, S2PAT>(tm, offs1, P, true, sh.data(), ph, mh);
    const size_t words = bb::row_hist_bytes<NWT, PACKED>(bb::resident_rows(L - P, MITM)) / 4;     // exactly the rows the kernel reserves
    const size_t front = 64;
    std::vector<uint32_t> hist(front + words + 256, 0xabababab);          // guard zones around the reserved rows that must stay untouched
    std::vector<uint8_t> rec(64 * 32 + 64, 0xcd);
    bb::RowHist<NWT, PACKED> H{hist.data() + front, lane};
    uint32_t lut[256];
    for (int q = 0; q < 256; q++) lut[q] = bb::scan_lut_entry(q);
    bb::rows_lane<NWT, PACKED, S2PAT, MITM>(tm, offs, rn, L, P, ph, mh, sh.data(), H, rec.data() + lane, lut, pb0, pb1, pol, R);
    for (size_t q = 0; q < hist.size(); q++) if ((q < front || q >= front + words) && hist[q] != 0xabababab) return false;
    for (size_t q = 3 * static_cast<size_t>(P) * NWT; q < sh.size(); q++) if (sh[q] != 0x5a5a5a5a5a5a5a5aull) return false;
    return true;
}
}  // namespace

extern "C" {
// The reversed-order accumulation of the traceback loop over an op string (1 = match, 0 = non-match, path order), with runs of
// non-match ops folded into one exact scaling by a power of two like the kernel's per-row records; returns whether
// lodhi_exact() vouches for the score.
int emu_lodhi_reversed(const uint8_t* ops, int n, double* score) {
    double a1 = 0.0, a2 = 0.0, s = 0.0;
    int q = n - 1;
    while (q >= 0) {
        if (ops[q]) {                                        // a match: g = 2^-1
            const double g = bb::bb_bits_to_double(static_cast<uint64_t>(1022) << 52);
            s = bb::bb_fma(g, a2, s);
            a2 = g * bb::bb_fma(1.0, a1, a2);
            a1 = bb::bb_fma(g, a1, 0.5);
            q--;
        } else {                                             // a run of non-match ops: g = 2^-run
            int run = 0;
            while (q >= 0 && !ops[q] && run < 40) { run++; q--; }
            const double g = bb::bb_bits_to_double(static_cast<uint64_t>(1023 - run) << 52);
            a2 = g * a2; a1 = g * a1;
        }
    }
    *score = s;
    return bb::lodhi_exact(s, n) ? 1 : 0;
}
// out[12] = {cbest, jend, ts, cnt, i_first, i_last, j_first, j_last, sub_cost, n_ops, variant, replayed}; score = Lodhi S_3
// P = leading pattern rows computed by rows_prefix() (the rows a warp shares); pol = kPolS1Left | kPolS2PatFirst | kPolS5Last;
// words = 0: like the kernels (one text word when the region has <= 64 bases, three otherwise), or force 1 / 3; + 16 = meet-in-the-middle records
int emu_barcode_rows(const uint8_t* pattern, int L, const uint8_t* region, int rn, int pb0, int pb1, int lane, int P, int pol, int words,
                     int32_t* out, double* score) {
    init_codes();
    if (L < 1 || L > 64 || rn < 0 || rn > 192 || lane < 0 || lane > 31 || P < 0 || P > L) return -1;
    uint8_t offs[64];
    std::memset(offs, 0, sizeof offs);
    for (int i = 0; i < L; i++) offs[i] = static_cast<uint8_t>(g_code[pattern[i]] << 3);
    const bool mitm = (words & 16) != 0;
    words &= 15;
    if (words == 0) words = rn <= 64 ? 1 : 3;
    if (words == 1 && rn > 64) return -1;
    const bool packed = words == 1 && rn <= 48;
    const bool s2 = (pol & bb::kPolS2PatFirst) != 0;
    bb::LaneAlign R;
    bool ok;
#define BB_RUN(N, PK, S2, MM) run<N, PK, S2, MM>(region, offs, rn, L, P, lane, pb0, pb1, pol, R)
#define BB_RUN2(N, PK) (s2 ? (mitm ? BB_RUN(N, PK, true, true) : BB_RUN(N, PK, true, false)) : (mitm ? BB_RUN(N, PK, false, true) : BB_RUN(N, PK, false, false)))
    if (words == 3) ok = BB_RUN2(3, false);
    else if (packed) ok = BB_RUN2(1, true);
    else ok = BB_RUN2(1, false);
    if (!ok) return -2;
    out[0] = R.cbest; out[1] = R.jend; out[2] = R.ts; out[3] = R.cnt; out[4] = R.i_first; out[5] = R.i_last;
    out[6] = R.j_first; out[7] = R.j_last; out[8] = R.sub_cost; out[9] = R.n_ops; out[10] = words * 2 + (packed ? 1 : 0);
    out[11] = bb::lodhi_exact(R.s, R.n_ops) ? 0 : 1;
    *score = R.s;
    return 0;
}
}
