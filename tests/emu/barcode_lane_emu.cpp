// Host build of the barcode stage's per-lane routine (barbell_b200/csrc/barcode_lane.cuh) for the CPU test-suite:
// the SAME source the GPU kernel compiles, run for one lane of an emulated warp, so that its results can be compared
// with the oracle without a GPU.  Test infrastructure only.
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "../../barbell_b200/csrc/barcode_lane.cuh"

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
}  // namespace

extern "C" {
// The reversed-order accumulation of the traceback loop over an op string (1 = match, 0 = non-match, path order), with runs of
// non-match ops folded into one exact scaling by a power of two like the kernel's per-column records; returns whether
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
// out[12] = {cbest, jend, ts, cnt, i_first, i_last, j_first, j_last, sub_cost, n_ops, packed_used, 0}; score = Lodhi S_3
int emu_barcode_lane(const uint8_t* pattern, int L, const uint8_t* region, int rn, int pb0, int pb1, int lane, int hist_cols,
                     int32_t* out, double* score) {
    init_codes();
    if (L < 1 || L > 64 || rn < 0 || rn > hist_cols || lane < 0 || lane > 31) return -1;
    const int sh = 64 - L;
    const uint64_t wild = sh ? ((1ull << sh) - 1ull) : 0ull;
    std::vector<uint64_t> eqs(bb::kEqSlots * 32, 0xdeadbeefdeadbeefull);
    for (int sl = 0; sl < bb::kEqSlots; sl++) {
        const int code = sl < 4 ? (1 << sl) : 15;
        uint64_t v = 0;
        for (int i = 0; i < L; i++) if (g_code[pattern[i]] & code) v |= 1ull << i;
        eqs[sl * 32 + lane] = (v << sh) | wild;
    }
    std::vector<uint8_t> txt(rn + 16, 0);
    for (int q = 0; q < rn; q++) txt[q] = bb::region_byte(g_code[region[q]]);
    const int half = (hist_cols + 1) / 2;                     // the kernel keeps half of the columns resident
    const bool packed = L <= 48;
    const size_t hist_words = static_cast<size_t>(half) * 32 * (packed ? 3 : 4);   // exactly what the kernel reserves per warp
    std::vector<uint32_t> hist(hist_words + 256, 0xabababab);                        // + a guard zone that must stay untouched
    std::vector<uint8_t> rec(static_cast<size_t>(hist_cols) * 32 + 64, 0xcd);
    bb::LaneAlign R;
    // like the kernel: the FAST variant when every base is A/C/G/T/N, the general one otherwise or when FAST gives up
    bool plain = true;
    for (int q = 0; q < rn; q++) plain = plain && (txt[q] >> 4) < bb::kEqOther;
    int variant = 0;
    auto run = [&](auto packed_tag) {
        constexpr bool P = decltype(packed_tag)::value;
        bb::ColHist<P> H{hist.data(), lane};
        if (plain) { bb::barcode_lane<P, true>(eqs.data() + lane, txt.data(), rn, L, pb0, pb1, H, rec.data() + lane, R); variant = 1; }
        else { bb::barcode_lane<P, false>(eqs.data() + lane, txt.data(), rn, L, pb0, pb1, H, rec.data() + lane, R); variant = 2; }
    };
    if (packed) run(std::true_type{}); else run(std::false_type{});
    for (size_t q = hist_words; q < hist.size(); q++) if (hist[q] != 0xabababab) return -2;   // wrote past the reserved columns
    out[0] = R.cbest; out[1] = R.jend; out[2] = R.ts; out[3] = R.cnt; out[4] = R.i_first; out[5] = R.i_last;
    out[6] = R.j_first; out[7] = R.j_last; out[8] = R.sub_cost; out[9] = R.n_ops; out[10] = packed; out[11] = variant;   // 1 = FAST (plain bases), 2 = general
    out[11] |= bb::lodhi_exact(R.s, R.n_ops) ? 0 : 4;          // 4: the records were replayed with the forward recurrence
    *score = R.s;
    return 0;
}
}
