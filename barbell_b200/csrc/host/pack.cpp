// Nibble packing of the read bases on the host, so that the PCIe copy moves half the bytes.
// The search depends on a text byte only through its 4-bit IUPAC base set (the match-mask tables are functions of it),
// so two bases per byte is a lossless wire format for this path; the device expands it back to one representative
// letter per set (k_unpack_nibbles).  AVX2 when the CPU has it (like the reference, which requires AVX2:
// bin/main.rs:268 ensure_simd), scalar otherwise; a small persistent thread pool splits the batch.
#include "pack.hpp"

#include <immintrin.h>

#include <algorithm>
#include <cstdlib>
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

// 128 input bytes -> 32 output bytes per iteration.  A C G T (any case; U as T) are told apart arithmetically: with x = c >> 1 and
// y = c >> 2 the two low bits of x ^ y are 0 1 2 3 for A C G T; a byte is one of those letters iff (c | 0x20) equals the letter its
// low nibble stands for (1 a, 3 c, 4 t, 5 u, 7 g).  Everything else gets crumb 0 and an exception entry with its 4-bit base set.
__attribute__((target("avx2"))) void crumbs_avx2(const uint8_t* src, size_t pos0, size_t n, uint8_t* dst, const uint8_t* code, ExcWriter& w) {
    alignas(32) static const uint8_t kLetter[32] = {0, 'a', 0, 'c', 't', 'u', 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0,
                                                    0, 'a', 0, 'c', 't', 'u', 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0};
    const __m256i L = _mm256_load_si256(reinterpret_cast<const __m256i*>(kLetter));
    const __m256i three = _mm256_set1_epi8(3), low4 = _mm256_set1_epi8(0x0f), lower = _mm256_set1_epi8(0x20);
    const __m256i mul4 = _mm256_set1_epi16(0x0401), mul16 = _mm256_set1_epi16(0x1001);
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        __m256i c[4];
        for (int q = 0; q < 4; q++) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32 * q));
            const __m256i ok = _mm256_cmpeq_epi8(_mm256_or_si256(v, lower), _mm256_shuffle_epi8(L, _mm256_and_si256(v, low4)));
            uint32_t m = ~static_cast<uint32_t>(_mm256_movemask_epi8(ok));
            while (m) {                                                     // rare: bytes that are not a single base
                const int t = __builtin_ctz(m); m &= m - 1;
                const size_t at = i + 32 * q + t;
                w.push(static_cast<uint64_t>(pos0 + at) << 4 | code[src[at]]);
            }
            const __m256i x = _mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2));   // (bits shifted in from the neighbour byte land above bit 1)
            c[q] = _mm256_and_si256(_mm256_and_si256(x, three), ok);
        }
        const __m256i n01 = _mm256_permute4x64_epi64(_mm256_packus_epi16(_mm256_maddubs_epi16(c[0], mul4), _mm256_maddubs_epi16(c[1], mul4)), 0xD8);
        const __m256i n23 = _mm256_permute4x64_epi64(_mm256_packus_epi16(_mm256_maddubs_epi16(c[2], mul4), _mm256_maddubs_epi16(c[3], mul4)), 0xD8);
        const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi16(_mm256_maddubs_epi16(n01, mul16), _mm256_maddubs_epi16(n23, mul16)), 0xD8);
        if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + (i >> 2)), pk);
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + (i >> 2)), pk);
    }
    if (aligned) _mm_sfence();
    crumbs_scalar(src + i, pos0 + i, n - i, dst + (i >> 2), code, w);
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

bool crumbs_append(const uint8_t* src, size_t n, uint8_t* dst, uint64_t* pos, uint64_t* exc, size_t exc_cap, size_t* n_exc, const uint8_t* code) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    std::atomic<bool> over{false};
    ExcWriter w{exc, exc_cap, nullptr, &over};
    w.cur = *n_exc; w.end = exc_cap;
    size_t p = static_cast<size_t>(*pos), i = 0;
    // head: fill up the byte the previous read left partly used (its missing crumbs are still zero)
    for (; i < n && (p & 3); i++, p++) {
        const uint8_t v = code[src[i]], cr = kCrumbOfSet[v];
        if (cr & 0x80) w.push(static_cast<uint64_t>(p) << 4 | v);
        else dst[p >> 2] = static_cast<uint8_t>(dst[p >> 2] | cr << (2 * (p & 3)));
    }
    // body: from here on the stream is byte aligned; the tail of this read leaves the last byte partly used (upper crumbs zero)
    if (i < n) {
        if (avx2) crumbs_avx2(src + i, p, n - i, dst + (p >> 2), code, w); else crumbs_scalar(src + i, p, n - i, dst + (p >> 2), code, w);
        p += n - i;
    }
    *pos = p; *n_exc = w.cur;
    return !w.dead;
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
