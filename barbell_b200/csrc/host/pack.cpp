// Nibble packing of the read bases on the host, so that the PCIe copy moves half the bytes.
// The search depends on a text byte only through its 4-bit IUPAC base set (the match-mask tables are functions of it),
// so two bases per byte is a lossless wire format for this path; the device expands it back to one representative
// letter per set (k_unpack_nibbles).  AVX2 when the CPU has it (like the reference, which requires AVX2:
// bin/main.rs:268 ensure_simd), scalar otherwise; a small persistent thread pool splits the batch.
#include "pack.hpp"

#include <immintrin.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace bb {
namespace {

void pack_scalar(const uint8_t* src, size_t n, uint8_t* dst, const uint8_t* code) {
    size_t i = 0;
    for (; i + 1 < n; i += 2) dst[i >> 1] = static_cast<uint8_t>(code[src[i]] | (code[src[i + 1]] << 4));
    if (i < n) dst[i >> 1] = code[src[i]];
}

// 32 bytes -> their 4-bit codes.  Only letters carry a code: idx = c & 15 selects from the two 16-entry halves of the
// table for '@'..'_' (pshufb), bit 4 of c picks the half, and everything that is not a letter is masked to 0.
__attribute__((target("avx2"))) static inline __m256i codes32(__m256i c, __m256i TL, __m256i TH) {
    const __m256i idx = _mm256_and_si256(c, _mm256_set1_epi8(0x0f));
    const __m256i lo = _mm256_shuffle_epi8(TL, idx), hi = _mm256_shuffle_epi8(TH, idx);
    const __m256i v = _mm256_blendv_epi8(lo, hi, _mm256_slli_epi16(c, 3));       // bit 4 of c -> bit 7 (blendv selector)
    const __m256i biased = _mm256_sub_epi8(_mm256_or_si256(c, _mm256_set1_epi8(0x20)), _mm256_set1_epi8('a'));
    const __m256i is_letter = _mm256_cmpeq_epi8(_mm256_min_epu8(biased, _mm256_set1_epi8(25)), biased);   // 0..25 unsigned
    return _mm256_and_si256(v, is_letter);
}

// 64 input bytes -> 32 output bytes per iteration.
__attribute__((target("avx2"))) void pack_avx2(const uint8_t* src, size_t n, uint8_t* dst, const uint8_t* code) {
    alignas(32) uint8_t tl[32], th[32];
    for (int i = 0; i < 16; i++) { tl[i] = tl[i + 16] = code[0x40 + i]; th[i] = th[i + 16] = code[0x50 + i]; }   // '@'..'O', 'P'..'_'
    const __m256i TL = _mm256_load_si256(reinterpret_cast<const __m256i*>(tl));
    const __m256i TH = _mm256_load_si256(reinterpret_cast<const __m256i*>(th));
    const __m256i mul = _mm256_set1_epi16(0x1001);               // bytes (1, 16): lo + 16 * hi per 16-bit lane
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m256i a = codes32(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)), TL, TH);
        const __m256i b = codes32(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32)), TL, TH);
        const __m256i pa = _mm256_maddubs_epi16(a, mul), pb = _mm256_maddubs_epi16(b, mul);   // 16 x u16 each, values < 256
        const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi16(pa, pb), 0xD8);       // undo the per-lane interleave
        if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + (i >> 1)), pk);   // no read-for-ownership of the output
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + (i >> 1)), pk);
    }
    if (aligned) _mm_sfence();
    pack_scalar(src + i, n - i, dst + (i >> 1), code);
}

// ---- two bits per base + an exception list ------------------------------------------------------------------------------
// Reads are almost all A/C/G/T, so the densest lossless wire format for this path is 2 bits per base (A C G T = 0 1 2 3, any
// case, U as T) with every other byte sent as an exception: (position << 4 | 4-bit base set).  The device expands the crumbs
// to letters and then patches the exceptions (k_unpack_crumbs, k_patch_exceptions).
constexpr uint64_t kExcEmpty = ~0ull;                             // filler of a partly used exception block (the device skips it)
constexpr size_t kExcBlock = 256;                                 // entries a work item takes from the shared list at a time

struct ExcWriter {
    uint64_t* buf; size_t cap; std::atomic<uint64_t>* next; std::atomic<bool>* overflow;   // next == nullptr: one writer, entries [cur, end) are its own
    size_t cur = 0, end = 0;
    bool dead = false;                                            // the list is full: the batch will be re-packed as nibbles
    void push(uint64_t v) {
        if (dead) return;
        if (cur == end) {
            if (!next) { dead = true; return; }
            const uint64_t b = next->fetch_add(kExcBlock);
            if (b + kExcBlock > cap) { overflow->store(true); next->fetch_sub(kExcBlock); dead = true; return; }
            cur = static_cast<size_t>(b); end = cur + kExcBlock;
        }
        buf[cur++] = v;
    }
    void finish() { if (next) while (cur < end) buf[cur++] = kExcEmpty; }
};

// base set -> crumb; 0x80 marks "not a single base": goes to the exception list
alignas(32) const uint8_t kCrumbOfSet[32] = {0x80, 0, 1, 0x80, 2, 0x80, 0x80, 0x80, 3, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80,
                                             0x80, 0, 1, 0x80, 2, 0x80, 0x80, 0x80, 3, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80};

void crumbs_scalar(const uint8_t* src, size_t pos0, size_t n, uint8_t* dst, const uint8_t* code, ExcWriter& w) {
    for (size_t i = 0; i < n; i += 4) {
        uint8_t b = 0;
        for (size_t t = 0; t < 4 && i + t < n; t++) {
            const uint8_t v = code[src[i + t]], cr = kCrumbOfSet[v];
            if (cr & 0x80) w.push(static_cast<uint64_t>(pos0 + i + t) << 4 | v);
            else b = static_cast<uint8_t>(b | cr << (2 * t));
        }
        dst[i >> 2] = b;
    }
}

// 128 input bytes -> 32 output bytes per step.  A C G T U (any case) have distinct low nibbles (1 3 7 4 5), so one pshufb gives the
// crumb and a second one the lower-case letter that nibble stands for; a byte is a single base iff (c | 0x20) equals that letter
// (bytes >= 0x80 look up 0 and fail).  Everything else gets crumb 0 and an exception entry with its 4-bit base set.  Four bytes of
// crumbs are folded into one with two multiply-adds (1, 4 per byte pair; 1, 16 per word pair), 4 x 8 dwords are narrowed to 32 bytes.
// Only blocks with a byte that is not a single base leave the straight path (one movemask per 128 bytes) -- for its exception entries;
// the crumbs are already right (such a byte has crumb 0).
// LINE: the input is a text line of unknown length <= n: stop in front of the first '\n' (*found), which -- not being a base -- can
// only sit in such a block.  The block that holds the end of the input is packed from a copy (or in place when 128 bytes are
// readable) with the crumbs past the end forced to 0: the next read ORs its first bases into the partly used last byte.
struct CrumbConsts { __m256i L, CR, lower, mul4, mul16, perm, iota; };
__attribute__((target("avx2"))) static inline CrumbConsts crumb_consts() {
    alignas(32) static const uint8_t kLetter[32] = {0, 'a', 0, 'c', 't', 'u', 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0,
                                                    0, 'a', 0, 'c', 't', 'u', 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0};
    alignas(32) static const uint8_t kCrumb[32] = {0, 0, 0, 1, 3, 3, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0,
                                                   0, 0, 0, 1, 3, 3, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0};
    alignas(32) static const uint8_t kIota[32] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31};
    CrumbConsts k;
    k.L = _mm256_load_si256(reinterpret_cast<const __m256i*>(kLetter));
    k.CR = _mm256_load_si256(reinterpret_cast<const __m256i*>(kCrumb));
    k.iota = _mm256_load_si256(reinterpret_cast<const __m256i*>(kIota));
    k.lower = _mm256_set1_epi8(0x20);
    k.mul4 = _mm256_set1_epi16(0x0401);
    k.mul16 = _mm256_set1_epi32(0x00100001);
    k.perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    return k;
}
// the 32 output bytes of four vectors of masked crumbs
__attribute__((target("avx2"))) static inline __m256i crumbs_fold(const __m256i c[4], const CrumbConsts& k) {
    __m256i d[4];
    for (int q = 0; q < 4; q++) d[q] = _mm256_madd_epi16(_mm256_maddubs_epi16(c[q], k.mul4), k.mul16);      // 8 dwords, each one output byte
    const __m256i pk = _mm256_packus_epi16(_mm256_packus_epi32(d[0], d[1]), _mm256_packus_epi32(d[2], d[3]));
    return _mm256_permutevar8x32_epi32(pk, k.perm);                                                           // undo the per-lane interleave
}
// one block of `len` <= 128 valid bytes at s (128 readable): exceptions + 32 output bytes (crumbs at >= len are 0)
__attribute__((target("avx2"))) static inline void crumbs_block_masked(const uint8_t* s, size_t len, size_t pos, uint8_t* out, const uint8_t* code,
                                                                       ExcWriter& w, const CrumbConsts& k) {
    __m256i c[4];
    for (int q = 0; q < 4; q++) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32 * q));
        const int left = static_cast<int>(len) - 32 * q;
        const __m256i in = _mm256_cmpgt_epi8(_mm256_set1_epi8(static_cast<char>(std::max(0, std::min(127, left)))), k.iota);
        const __m256i ok = _mm256_cmpeq_epi8(_mm256_or_si256(v, k.lower), _mm256_shuffle_epi8(k.L, v));
        uint32_t m = ~static_cast<uint32_t>(_mm256_movemask_epi8(ok)) & static_cast<uint32_t>(_mm256_movemask_epi8(in));
        while (m) {
            const int t = __builtin_ctz(m); m &= m - 1;
            w.push(static_cast<uint64_t>(pos + 32 * q + t) << 4 | code[s[32 * q + t]]);
        }
        c[q] = _mm256_and_si256(_mm256_shuffle_epi8(k.CR, v), _mm256_and_si256(ok, in));
    }
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out), crumbs_fold(c, k));
}
template <bool LINE>
__attribute__((target("avx2"))) size_t crumbs_avx2_t(const uint8_t* src, size_t pos0, size_t n, uint8_t* dst, const uint8_t* code, ExcWriter& w, bool* found) {
    const CrumbConsts k = crumb_consts();
    const bool aligned = !LINE && (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    static const size_t kAhead = [] { const char* e = std::getenv("BB_PACK_AHEAD"); return e ? static_cast<size_t>(std::atoi(e)) : size_t(1024); }();
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        __m256i c[4], all = _mm256_set1_epi8(-1);
        _mm_prefetch(reinterpret_cast<const char*>(src + i + kAhead), _MM_HINT_T0);
        _mm_prefetch(reinterpret_cast<const char*>(src + i + kAhead + 64), _MM_HINT_T0);
        for (int q = 0; q < 4; q++) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32 * q));
            const __m256i ok = _mm256_cmpeq_epi8(_mm256_or_si256(v, k.lower), _mm256_shuffle_epi8(k.L, v));
            c[q] = _mm256_and_si256(_mm256_shuffle_epi8(k.CR, v), ok);
            all = _mm256_and_si256(all, ok);
        }
        if (__builtin_expect(_mm256_movemask_epi8(all) != -1, 0)) {           // a byte that is not a single base (or the end of the line)
            size_t len = 128;
            for (int q = 0; q < 4 && len == 128; q++) {
                const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32 * q));   // (not kept live: the straight path needs the registers)
                const __m256i ok = _mm256_cmpeq_epi8(_mm256_or_si256(v, k.lower), _mm256_shuffle_epi8(k.L, v));
                uint32_t m = ~static_cast<uint32_t>(_mm256_movemask_epi8(ok));
                while (m) {
                    const int t = __builtin_ctz(m); m &= m - 1;
                    const size_t at = static_cast<size_t>(32 * q + t);
                    if (LINE && src[i + at] == '\n') { len = at; break; }
                    w.push(static_cast<uint64_t>(pos0 + i + at) << 4 | code[src[i + at]]);
                }
            }
            if (LINE && len < 128) {                                          // the line ends in this block: crumbs past its end are 0
                for (int q = 0; q < 4; q++) {
                    const int left = static_cast<int>(len) - 32 * q;
                    c[q] = _mm256_and_si256(c[q], _mm256_cmpgt_epi8(_mm256_set1_epi8(static_cast<char>(std::max(0, std::min(127, left)))), k.iota));
                }
                _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + (i >> 2)), crumbs_fold(c, k));
                *found = true;
                return i + len;
            }
        }
        const __m256i pk = crumbs_fold(c, k);
        if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + (i >> 2)), pk);   // no read-for-ownership of the output
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + (i >> 2)), pk);
    }
    if (aligned) _mm_sfence();
    if (i < n) {                                                              // fewer than 128 readable bytes: pack a padded copy
        alignas(32) uint8_t tmp[128];
        size_t len = n - i;
        std::memcpy(tmp, src + i, len);
        std::memset(tmp + len, 0, 128 - len);
        if (LINE) {
            const void* nl = std::memchr(tmp, '\n', len);
            if (nl) { len = static_cast<size_t>(static_cast<const uint8_t*>(nl) - tmp); *found = true; }
        }
        alignas(32) uint8_t out[32];
        crumbs_block_masked(tmp, len, pos0 + i, out, code, w, k);
        std::memcpy(dst + (i >> 2), out, (len + 3) >> 2);
        i += len;
    }
    return i;
}
void crumbs_avx2(const uint8_t* src, size_t pos0, size_t n, uint8_t* dst, const uint8_t* code, ExcWriter& w) {
    bool f = false;
    crumbs_avx2_t<false>(src, pos0, n, dst, code, w, &f);
}

class Pool {
  public:
    explicit Pool(int n) {
        for (int t = 0; t < n; t++) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    // run fn(chunk) for chunk in [0, n_chunks) on the pool + the calling thread; returns the seconds the region itself took
    double run(int n_chunks, const std::function<void(int)>& fn) {
        std::unique_lock<std::mutex> call(call_mu_);              // one parallel region at a time
        const auto t0 = std::chrono::steady_clock::now();
        // every region has its OWN job object: a worker that drew its last (out-of-range) index from an earlier region and was
        // preempted before looking at it can only ever touch that earlier region's counters
        auto job = std::make_shared<Job>();
        job->fn = &fn; job->total = n_chunks; job->pending.store(n_chunks);
        {
            std::lock_guard<std::mutex> lk(mu_);
            cur_ = job; gen_++;
        }
        cv_.notify_all();
        work(*job);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return job->pending.load() == 0; });
        cur_.reset();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    int size() const { return static_cast<int>(workers_.size()) + 1; }

  private:
    struct Job { const std::function<void(int)>* fn = nullptr; std::atomic<int> next{0}; int total = 0; std::atomic<int> pending{0}; };
    void work(Job& j) {
        for (;;) {
            const int c = j.next.fetch_add(1);
            if (c >= j.total) break;
            (*j.fn)(c);
            if (j.pending.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(mu_); done_cv_.notify_all(); }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                j = cur_;
            }
            if (j) work(*j);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    std::shared_ptr<Job> cur_;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

Pool& pool(int threads) {
    static Pool p(std::max(0, threads - 1));
    return p;
}
}  // namespace

// shared body of crumbs_append / crumbs_append_line
template <bool LINE>
static bool append_impl(const uint8_t* src, size_t n, uint8_t* dst, uint64_t* pos, uint64_t* exc, size_t exc_cap, size_t* n_exc, const uint8_t* code,
                        size_t* line_len, bool* found) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    std::atomic<bool> over{false};
    ExcWriter w{exc, exc_cap, nullptr, &over};
    w.cur = *n_exc; w.end = exc_cap;
    size_t p = static_cast<size_t>(*pos), i = 0;
    bool nl = false;
    // head: fill up the byte the previous read left partly used (its missing crumbs are still zero)
    for (; i < n && (p & 3); i++, p++) {
        if (LINE && src[i] == '\n') { nl = true; break; }
        const uint8_t v = code[src[i]], cr = kCrumbOfSet[v];
        if (cr & 0x80) w.push(static_cast<uint64_t>(p) << 4 | v);
        else dst[p >> 2] = static_cast<uint8_t>(dst[p >> 2] | cr << (2 * (p & 3)));
    }
    // body: from here on the stream is byte aligned; the tail of this read leaves the last byte partly used (upper crumbs zero)
    if (i < n && !nl) {
        size_t took;
        if (avx2) {
            took = crumbs_avx2_t<LINE>(src + i, p, n - i, dst + (p >> 2), code, w, &nl);
        } else {
            took = n - i;
            if (LINE) { const void* e = std::memchr(src + i, '\n', n - i); if (e) { took = static_cast<size_t>(static_cast<const uint8_t*>(e) - (src + i)); nl = true; } }
            crumbs_scalar(src + i, p, took, dst + (p >> 2), code, w);
        }
        p += took; i += took;
    }
    if (LINE) {
        if (nl && i > 0 && src[i - 1] == '\r' && !w.dead) {     // CRLF: the '\r' went in as an exception (crumb 0, base set 0) -- take it back
            p--; i--; w.cur--;
        }
        *line_len = i; *found = nl;
    }
    *pos = p; *n_exc = w.cur;
    return !w.dead;
}
bool crumbs_append(const uint8_t* src, size_t n, uint8_t* dst, uint64_t* pos, uint64_t* exc, size_t exc_cap, size_t* n_exc, const uint8_t* code) {
    return append_impl<false>(src, n, dst, pos, exc, exc_cap, n_exc, code, nullptr, nullptr);
}
bool crumbs_append_line(const uint8_t* src, size_t n, uint8_t* dst, uint64_t* pos, uint64_t* exc, size_t exc_cap, size_t* n_exc, const uint8_t* code,
                        size_t* line_len, bool* found) {
    return append_impl<true>(src, n, dst, pos, exc, exc_cap, n_exc, code, line_len, found);
}

int pack_default_threads() {
    if (const char* e = std::getenv("BB_PACK_THREADS")) { const int v = std::atoi(e); if (v >= 1) return std::min(v, 64); }   // tuning knob
    const unsigned hc = std::thread::hardware_concurrency();
    return static_cast<int>(std::min(64u, std::max(1u, hc)));
}

double pack_nibbles(const uint8_t* src, size_t n, uint8_t* dst, const uint8_t* code, int threads) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    // (an AVX-512 VBMI form -- one vpermi2b as the 128-entry table -- packs no faster per core, the loop is bound by the core's
    //  streaming bandwidth, and end to end it was SLOWER on the B200 host: 4.5 vs 6.8 M reads/s, so it is not built)
    auto one = [&](const uint8_t* s, size_t len, uint8_t* d) { if (avx2) pack_avx2(s, len, d, code); else pack_scalar(s, len, d, code); };
    const size_t kChunk = 4u << 20;                              // bases per work item (even, so chunks start on a byte boundary)
    const int n_chunks = static_cast<int>((n + kChunk - 1) / kChunk);
    if (threads <= 1 || n_chunks <= 1) {
        const auto t0 = std::chrono::steady_clock::now();
        one(src, n, dst);
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return pool(threads).run(n_chunks, [&](int c) {
        const size_t lo = static_cast<size_t>(c) * kChunk, len = std::min(kChunk, n - lo);
        one(src + lo, len, dst + (lo >> 1));
    });
}

double pack_crumbs(const uint8_t* src, size_t n, uint8_t* dst, uint64_t* exc, size_t exc_cap, size_t* n_exc, bool* overflow,
                   const uint8_t* code, int threads) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    std::atomic<uint64_t> next{0};
    std::atomic<bool> over{false};
    auto one = [&](const uint8_t* s, size_t pos0, size_t len, uint8_t* d) {
        ExcWriter w{exc, exc_cap, &next, &over};
        if (avx2) crumbs_avx2(s, pos0, len, d, code, w); else crumbs_scalar(s, pos0, len, d, code, w);
        w.finish();
    };
    const size_t kChunk = 4u << 20;                              // bases per work item (a multiple of 4: chunks start on a byte boundary)
    const int n_chunks = static_cast<int>((n + kChunk - 1) / kChunk);
    double secs;
    if (threads <= 1 || n_chunks <= 1) {
        const auto t0 = std::chrono::steady_clock::now();
        one(src, 0, n, dst);
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } else {
        secs = pool(threads).run(n_chunks, [&](int c) {
            const size_t lo = static_cast<size_t>(c) * kChunk, len = std::min(kChunk, n - lo);
            one(src + lo, lo, len, dst + (lo >> 2));
        });
    }
    *n_exc = static_cast<size_t>(next.load());
    *overflow = over.load();
    return secs;
}
}  // namespace bb
