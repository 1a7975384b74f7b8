// `barbell` command line: the reference's `annotate` and `kit` subcommands (bin/main.rs:61-112, 211-263, 274-339) on top
// of the C ABI.  FASTQ(.gz) records are parsed into page-locked batch buffers and pushed through bb_submit/bb_collect
// (two batches in flight per GPU); rows are written as annotation.tsv in input order (reference column order,
// src/annotate/searcher.rs:31-64; the header is written with the first row, so a run without hits leaves an empty
// file exactly like the reference's csv writer, annotator.rs:20-24).
// `filter`, `inspect` and `trim` (bin/main.rs:113-209, 340-395) are thin wrappers over bb_filter / bb_inspect / bb_trim, and
// `kit` chains annotate -> inspect -> filter -> trim with the reference's fixed file names (src/kits/use_kit.rs:11-109).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/barbell_b200.h"
#include "fastq.hpp"

namespace {

struct Args {
    std::string cmd, output = "output.tsv", kit;
    std::vector<std::string> input, queries, barcode_types{"Ftag"};
    int threads = 10, flank_max_errors = -1, gpus = 1;
    bool verbose = false, use_extended = false, maximize = false, gzip = false;
    // filter / trim / inspect
    std::vector<std::string> pattern_files, reads;
    std::string dropped, failed_out, only_side, read_pattern_out;
    bool no_label = false, no_orientation = false, no_flanks = false, sort_labels = false, skip_trim = false, flip = false;
    int top_n = 10, bucket_size = 250;
    bool output_given = false, single_reader = false, no_pack = false, count_only = false;
    size_t chunk_kb = 0;
    double min_score = 0.2, min_score_diff = 0.1;
    float alpha = 0.4f;
    size_t batch_mb = 128;
    unsigned policy = BB_POL_DEFAULT;    // bb_opts.policy (BB_POL_*): the sassy choices the reference's tests do not pin
};

[[noreturn]] void usage(const char* msg) {
    if (msg) std::fprintf(stderr, "error: %s\n", msg);
    std::fprintf(stderr,
        "barbell (B200 build of the annotate path)\n"
        "  barbell annotate -i <fastq>... [-o output.tsv] (--kit <KIT> | -q <fasta>... [-b Ftag|Rtag ...])\n"
        "                   [-t N] [--flank-max-errors INT] [--min-score F] [--min-score-diff F] [--alpha F]\n"
        "                   [--use-extended] [--verbose] [--gpus N] [--batch-mb MB] [--no-pack] [--policy BITS]\n"
        "  barbell kit -k <KIT> -i <fastq>... -o <folder> [--maximize] [--failed-out FILE] [--gzip] [annotate options]\n"
        "  barbell filter -i annotation.tsv -o filtered.tsv -f <pattern file>... [--dropped FILE]\n"
        "  barbell trim -i filtered.tsv -r <fastq>... -o <folder> [--no-label] [--no-orientation] [--no-flanks] [--sort-labels]\n"
        "               [--only-side left|right] [--failed-out FILE] [--skip-trim] [--flip] [--gzip]\n"
        "  barbell inspect -i annotation.tsv [-n 10] [-o read_patterns.tsv] [-s 250]\n");
    std::exit(msg ? 2 : 0);
}

bool is_flag(const std::string& s) { return s.size() > 1 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9'); }

Args parse(int argc, char** argv) {
    Args a;
    if (argc < 2) usage(nullptr);
    a.cmd = argv[1];
    if (a.cmd == "-h" || a.cmd == "--help") usage(nullptr);
    if (a.cmd != "annotate" && a.cmd != "kit" && a.cmd != "filter" && a.cmd != "trim" && a.cmd != "inspect" && a.cmd != "fastq-stats")
        usage("subcommands: annotate, kit, filter, trim, inspect");
    bool types_given = false;
    for (int i = 2; i < argc; i++) {
        std::string f = argv[i];
        auto many = [&](std::vector<std::string>& dst) { while (i + 1 < argc && !is_flag(argv[i + 1])) dst.push_back(argv[++i]); };
        auto one = [&]() -> std::string { if (i + 1 >= argc) usage(("missing value for " + f).c_str()); return argv[++i]; };
        if (f == "-i" || f == "--input") many(a.input);
        else if (f == "-q" || f == "--queries") many(a.queries);
        else if (f == "-b" || f == "--barcode-types") { if (!types_given) a.barcode_types.clear(); types_given = true; many(a.barcode_types); }
        else if ((f == "-o" || f == "--read-pattern-out") && a.cmd == "inspect") a.read_pattern_out = one();
        else if (f == "-o" || f == "--output") { a.output = one(); a.output_given = true; }
        else if (f == "-f" || f == "--file") many(a.pattern_files);
        else if (f == "-r" || f == "--reads") many(a.reads);
        else if (f == "--dropped") a.dropped = one();
        else if (f == "--no-label") a.no_label = true;
        else if (f == "--no-orientation") a.no_orientation = true;
        else if (f == "--no-flanks") a.no_flanks = true;
        else if (f == "--sort-labels") a.sort_labels = true;
        else if (f == "--only-side") a.only_side = one();
        else if (f == "--skip-trim") a.skip_trim = true;
        else if (f == "--flip") a.flip = true;
        else if ((f == "-n" || f == "--top-n") && a.cmd == "inspect") a.top_n = std::atoi(one().c_str());
        else if ((f == "-s" || f == "--bucket-size") && a.cmd == "inspect") a.bucket_size = std::atoi(one().c_str());
        else if (f == "-t" || f == "--threads") a.threads = std::atoi(one().c_str());
        else if (f == "--kit" || f == "-k") a.kit = one();
        else if (f == "--flank-max-errors") a.flank_max_errors = std::atoi(one().c_str());
        else if (f == "--min-score") a.min_score = std::atof(one().c_str());
        else if (f == "--min-score-diff") a.min_score_diff = std::atof(one().c_str());
        else if (f == "--alpha") a.alpha = static_cast<float>(std::atof(one().c_str()));
        else if (f == "--gpus") a.gpus = std::atoi(one().c_str());
        else if (f == "--batch-mb") a.batch_mb = static_cast<size_t>(std::atol(one().c_str()));
        else if (f == "--failed-out") a.failed_out = one();
        else if (f == "--verbose") a.verbose = true;
        else if (f == "--single-reader") a.single_reader = true;
        else if (f == "--no-pack") a.no_pack = true;
        else if (f == "--count-only") a.count_only = true;
        else if (f == "--policy") a.policy = static_cast<unsigned>(std::atoi(one().c_str()));
        else if (f == "--chunk-kb") a.chunk_kb = static_cast<size_t>(std::atol(one().c_str()));
        else if (f == "--use-extended") a.use_extended = true;
        else if (f == "--maximize") a.maximize = true;
        else if (f == "--gzip") a.gzip = true;
        else if (f == "-h" || f == "--help") usage(nullptr);
        else usage(("unknown argument " + f).c_str());
    }
    return a;
}

using bb::FastqReader;

// One batch of reads in page-locked memory, filled by the FASTQ parsers.  Default form: the bases as the library's 2-bit wire
// format, built read by read WHILE PARSING (bb_pack_crumbs_append: a sequence line is packed while it is still in the cache,
// no host core touches the bases again, a quarter of the bytes cross PCIe) + the exception list for every byte that is not
// A/C/G/T (it grows when a batch holds more than ~3 % such bytes).  --no-pack keeps one byte per base.
struct Batch {
    uint8_t* bases = nullptr;            // packed: crumbs; plain: bytes
    uint64_t* exc = nullptr;             // packed only
    uint64_t* offsets = nullptr;
    size_t cap_bytes = 0, cap_reads = 0, cap_exc = 0, bytes = 0;   // bytes = bases in the batch
    uint64_t n_exc = 0;
    uint32_t n_reads = 0;
    std::vector<bb_row> rows;            // the batch's result rows (copied out of the engine's buffer for the writer thread)
    std::vector<char> id_chars;          // read ids back to back
    std::vector<uint32_t> id_off;        // n_reads + 1
    bool pinned = true, packed = true, exc_overflow = false;
    bool alloc(size_t cb, size_t cr, bool pin, bool pack) {
        cap_bytes = cb; cap_reads = cr; pinned = pin; packed = pack;
        auto get = [&](size_t n) { return pin ? bb_host_alloc(n) : std::malloc(n); };
        bases = static_cast<uint8_t*>(get(pack ? cb / 4 + 128 : cb + 64));
        offsets = static_cast<uint64_t*>(get((cr + 1) * sizeof(uint64_t)));
        if (pack) { cap_exc = cb / 128 + 4096; exc = static_cast<uint64_t*>(get(cap_exc * sizeof(uint64_t))); }   // ~0.8 % of the bases; grows
        return bases && offsets && (!pack || exc);
    }
    // room for one more read in the offsets (they start sized for reads of >= 256 bases and grow for shorter ones)
    bool reserve_read() {
        if (n_reads + 1 <= cap_reads) return true;
        const size_t cap2 = cap_reads * 4;
        uint64_t* o2 = static_cast<uint64_t*>(pinned ? bb_host_alloc((cap2 + 1) * sizeof(uint64_t)) : std::malloc((cap2 + 1) * sizeof(uint64_t)));
        if (!o2) return false;
        std::memcpy(o2, offsets, (static_cast<size_t>(n_reads) + 1) * sizeof(uint64_t));
        if (pinned) bb_host_free(offsets); else std::free(offsets);
        offsets = o2; cap_reads = cap2;
        return true;
    }
    void clear() { bytes = 0; n_reads = 0; n_exc = 0; exc_overflow = false; id_chars.clear(); id_off.assign(1, 0); if (offsets) offsets[0] = 0; }
    bool grow_exc() {                    // N-rich input: the exception list grows (re-appending the read rewrites the same crumbs)
        const size_t cap2 = cap_exc * 4;
        uint64_t* e2 = static_cast<uint64_t*>(pinned ? bb_host_alloc(cap2 * sizeof(uint64_t)) : std::malloc(cap2 * sizeof(uint64_t)));
        if (!e2) { exc_overflow = true; return false; }
        std::memcpy(e2, exc, n_exc * sizeof(uint64_t));
        if (pinned) bb_host_free(exc); else std::free(exc);
        exc = e2; cap_exc = cap2;
        return true;
    }
    void push_read(const char* id, size_t id_len) {
        offsets[++n_reads] = bytes;
        id_chars.insert(id_chars.end(), id, id + id_len);
        id_off.push_back(static_cast<uint32_t>(id_chars.size()));
    }
    void append(const char* id, size_t id_len, const char* seq, size_t seq_len) {
        if (packed) {
            uint64_t pos = bytes, ne = n_exc;
            while (!exc_overflow && bb_pack_crumbs_append(reinterpret_cast<const uint8_t*>(seq), seq_len, bases, &pos, exc, cap_exc, &ne) != BB_OK) {
                if (!grow_exc()) break;
                pos = bytes; ne = n_exc;
            }
            n_exc = ne;
        } else {
            std::memcpy(bases + bytes, seq, seq_len);
        }
        bytes += seq_len;
        push_read(id, id_len);
    }
    // packed form, for a reader that has not looked for the end of the sequence line yet: the line at s (avail readable bytes) is
    // packed WHILE its end is searched -- one pass over the bases.  line_len = bases, term = bytes of its terminator (0 at the end
    // of the input); the read is entered with push_read once the rest of the record has been checked.
    bool append_seq_line(const char* s, size_t avail, size_t& line_len, size_t& term, const char*& what) {
        const size_t n = std::min(avail, cap_bytes - bytes);
        uint64_t pos = bytes, ne = n_exc, ll = 0; int found = 0;
        while (!exc_overflow && bb_pack_crumbs_append_line(reinterpret_cast<const uint8_t*>(s), n, bases, &pos, exc, cap_exc, &ne, &ll, &found) != BB_OK) {
            if (!grow_exc()) break;
            pos = bytes; ne = n_exc;
        }
        if (!found && n < avail) { what = "read longer than the batch buffer (raise --batch-mb)"; return false; }
        n_exc = ne; bytes += ll; line_len = ll;
        term = found ? (s[ll] == '\r' ? 2 : 1) : 0;
        return true;
    }
    void release() {
        auto put = [&](void* q) { if (pinned) bb_host_free(q); else std::free(q); };
        put(bases); put(offsets); if (exc) put(exc);
        bases = nullptr; offsets = nullptr; exc = nullptr;
    }
};

// hand-off between the reader thread and the GPU/writer thread
template <class T>
class Channel {
  public:
    void push(T v) { { std::lock_guard<std::mutex> lk(mu_); q_.push_back(v); } cv_.notify_one(); }
    T pop() { std::unique_lock<std::mutex> lk(mu_); cv_.wait(lk, [&] { return !q_.empty(); }); T v = q_.front(); q_.pop_front(); return v; }
  private:
    std::mutex mu_; std::condition_variable cv_; std::deque<T> q_;
};

// ---------------------------------------------------------------------------------------------------------------
// Batch sources: FASTQ -> page-locked batches, delivered to the GPU/writer thread IN INPUT ORDER.
//   SequentialSource  one reader thread over FastqReader (gzip or plain), like paraseq's reader thread (annotator.rs:278-280)
//   ParallelSource    plain (uncompressed) files only: the files are mapped, cut into chunks at record boundaries and parsed
//                     by several threads; chunk i always lands in slot i % n_slots, so batches come out in file order and a
//                     parser only ever waits for EARLIER chunks to be collected (no deadlock).
// ---------------------------------------------------------------------------------------------------------------
class BatchSource {
  public:
    virtual ~BatchSource() {}
    virtual int next_filled() = 0;          // slot index; -1 = end of input; -2 = error (see error())
    virtual void release(int slot) = 0;     // the batch in `slot` has been collected
    virtual void abort() = 0;               // consumer gives up: unblock the producers
    virtual void join() = 0;
    virtual const std::string& error() const = 0;
};

class SequentialSource : public BatchSource {
  public:
    SequentialSource(const std::vector<std::string>& paths, std::vector<Batch>& slots) : slots_(slots) {
        for (size_t s = 0; s < slots.size(); s++) free_.push(static_cast<int>(s));
        thread_ = std::thread([this, paths] { run(paths); });
    }
    int next_filled() override { return filled_.pop(); }
    void release(int slot) override { free_.push(slot); }
    void abort() override { for (size_t s = 0; s < slots_.size() + 2; s++) free_.push(-1); }
    void join() override { if (thread_.joinable()) thread_.join(); }
    const std::string& error() const override { return err_; }

  private:
    void run(const std::vector<std::string>& paths) {
        FastqReader reader(paths);
        FastqReader::View v;
        int cur = free_.pop();
        if (cur < 0) return;
        slots_[cur].clear();
        for (;;) {
            std::string err;
            const bool more = reader.next(v, err);
            if (!more && !err.empty()) { err_ = err; filled_.push(-2); return; }
            Batch* B = &slots_[cur];
            if (more && v.seq_len > B->cap_bytes) { err_ = "read longer than the batch buffer (raise --batch-mb)"; filled_.push(-2); return; }
            const bool full = more && (B->bytes + v.seq_len > B->cap_bytes || !B->reserve_read());
            if ((full || !more) && B->n_reads > 0) {
                filled_.push(cur);
                if (!more) break;
                cur = free_.pop();
                if (cur < 0) return;                              // consumer aborted
                B = &slots_[cur];
                B->clear();
            }
            if (!more) break;
            B->append(v.id, v.id_len, v.seq, v.seq_len);
        }
        filled_.push(-1);
    }
    std::vector<Batch>& slots_;
    Channel<int> free_, filled_;
    std::thread thread_;
    std::string err_;
};

class ParallelSource : public BatchSource {
  public:
    // every path must be a plain (not gzip) regular file; chunk_bytes <= batch capacity
    static bool usable(const std::vector<std::string>& paths) {
        for (const auto& p : paths) {
            FILE* f = std::fopen(p.c_str(), "rb");
            if (!f) return false;
            unsigned char m[2] = {0, 0};
            const size_t n = std::fread(m, 1, 2, f);
            std::fclose(f);
            struct stat st;
            if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return false;
            if (n == 2 && m[0] == 0x1f && m[1] == 0x8b) return false;
        }
        return !paths.empty();
    }
    ParallelSource(const std::vector<std::string>& paths, std::vector<Batch>& slots, size_t chunk_bytes, int n_threads)
        : slots_(slots), chunk_(chunk_bytes), round_(slots.size(), 0) {
        for (const auto& p : paths) {
            File f; f.path = p;
            f.fd = ::open(p.c_str(), O_RDONLY);
            struct stat st;
            if (f.fd < 0 || fstat(f.fd, &st) != 0) { err_ = "Failed to open FASTQ input: " + p; failed_ = true; break; }
            f.size = static_cast<size_t>(st.st_size);
            if (f.size) {
                void* m = mmap(nullptr, f.size, PROT_READ, MAP_PRIVATE, f.fd, 0);
                if (m == MAP_FAILED) { err_ = "mmap failed: " + p; failed_ = true; ::close(f.fd); break; }
                madvise(m, f.size, MADV_SEQUENTIAL);
                f.map = static_cast<const char*>(m);
            }
            f.first_chunk = n_chunks_;
            n_chunks_ += (f.size + chunk_ - 1) / chunk_;
            files_.push_back(f);
        }
        done_.assign(n_chunks_, 0);
        if (failed_) return;
        for (int t = 0; t < std::max(1, n_threads); t++) workers_.emplace_back([this] { work(); });
    }
    ~ParallelSource() override {
        join();
        for (auto& f : files_) { if (f.map) munmap(const_cast<char*>(f.map), f.size); if (f.fd >= 0) ::close(f.fd); }
    }
    int next_filled() override {
        if (failed_) return -2;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            if (deliver_ >= n_chunks_) return -1;
            cv_.wait(lk, [&] { return done_[deliver_] != 0; });
            if (done_[deliver_] == 2) return -2;
            const int slot = static_cast<int>(deliver_ % slots_.size());
            deliver_++;
            if (slots_[slot].n_reads == 0) { round_[slot]++; cv_.notify_all(); continue; }   // chunk without a record start
            return slot;
        }
    }
    void release(int slot) override { { std::lock_guard<std::mutex> lk(mu_); round_[slot]++; } cv_.notify_all(); }
    void abort() override { { std::lock_guard<std::mutex> lk(mu_); aborted_ = true; } cv_.notify_all(); }
    void join() override { for (auto& w : workers_) if (w.joinable()) w.join(); workers_.clear(); }
    const std::string& error() const override { return err_; }
    // seconds the parser threads spent parsing / waiting for a free batch slot, summed over the threads (after join())
    void thread_seconds(double& parse, double& wait) const { parse = parse_secs_; wait = wait_secs_; }

  private:
    struct File { std::string path; int fd = -1; size_t size = 0; const char* map = nullptr; size_t first_chunk = 0; };

    bool parse_chunk(const File& f, size_t k, Batch& B, std::string& err) {
        const char* m = f.map;
        const size_t begin = bb::fastq_record_start(m, f.size, k * chunk_), end = bb::fastq_record_start(m, f.size, std::min(f.size, (k + 1) * chunk_));
        size_t p = begin;
        while (p < end) {
            bb::FastqRec r; const char* what = nullptr;
            size_t idl, doff;
            if (B.packed) {
                const int st = bb::fastq_record_at(m, f.size, p, r, what, [&](const char*, size_t, size_t start, size_t& e, size_t& next) {
                    if (!B.reserve_read()) { what = "out of memory for the read offsets"; return false; }
                    size_t ll = 0, term = 0;
                    if (!B.append_seq_line(m + start, f.size - start, ll, term, what)) return false;
                    e = start + ll; next = e + term;
                    return true;
                });
                if (st == 1) continue;                             // blank line between records
                if (st < 0) { err = std::string(what) + " in " + f.path; return false; }
                bb::fastq_split_header(r.head, r.head_len, idl, doff);
                B.push_read(r.head, idl);
                continue;
            }
            const int st = bb::fastq_record_at(m, f.size, p, r, what);
            if (st == 1) continue;                                 // blank line between records
            if (st < 0) { err = std::string(what) + " in " + f.path; return false; }
            if (B.bytes + r.seq_len > B.cap_bytes || !B.reserve_read()) { err = "read longer than the batch buffer (raise --batch-mb)"; return false; }
            bb::fastq_split_header(r.head, r.head_len, idl, doff);
            B.append(r.head, idl, r.seq, r.seq_len);
        }
        return true;
    }
    void work() {
        // the parsers take every core they get; the few threads that feed the GPU and write the rows (short bursts, latency bound)
        // must not queue behind them: parsers run at a lower priority (a thread may always lower its own)
        if (!std::getenv("BB_PARSER_NICE_OFF")) setpriority(PRIO_PROCESS, static_cast<id_t>(syscall(SYS_gettid)), 10);
        for (;;) {
            const size_t i = claim_.fetch_add(1);
            if (i >= n_chunks_) return;
            const size_t slot = i % slots_.size(), need = i / slots_.size();
            const auto tw = std::chrono::steady_clock::now();
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return aborted_ || round_[slot] == need; });
                if (aborted_) return;
            }
            const auto tp = std::chrono::steady_clock::now();
            const File* f = &files_[0];
            for (const auto& ff : files_) if (i >= ff.first_chunk) f = &ff;
            Batch& B = slots_[slot];
            B.clear();
            std::string err;
            const bool ok = parse_chunk(*f, i - f->first_chunk, B, err);
            const auto te = std::chrono::steady_clock::now();
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!ok && err_.empty()) err_ = err;
                done_[i] = ok ? 1 : 2;
                wait_secs_ += std::chrono::duration<double>(tp - tw).count();
                parse_secs_ += std::chrono::duration<double>(te - tp).count();
            }
            cv_.notify_all();
            if (!ok) return;
        }
    }
    std::vector<Batch>& slots_;
    size_t chunk_, n_chunks_ = 0, deliver_ = 0;
    std::vector<File> files_;
    std::vector<uint64_t> round_;            // how often each slot has been released
    std::vector<uint8_t> done_;              // per chunk: 0 pending, 1 parsed, 2 error
    std::atomic<size_t> claim_{0};
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<std::thread> workers_;
    std::string err_;
    double parse_secs_ = 0, wait_secs_ = 0;
    bool failed_ = false, aborted_ = false;
};

// Several input files, some of them gzip (the usual shape of a nanopore run: thousands of small .fastq.gz): every worker inflates
// and parses ONE whole file into heap buffers; an assembler thread copies the parsed files, in input order, into the
// page-locked batches.  A window bounds the files parsed ahead of the assembler.
class MultiFileSource : public BatchSource {
  public:
    static bool usable(const std::vector<std::string>& paths) {
        if (paths.size() < 2) return false;
        for (const auto& p : paths) {
            struct stat st;
            if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode) || st.st_size > (1ll << 30)) return false;   // a parsed file lives on the heap
        }
        return true;
    }
    MultiFileSource(const std::vector<std::string>& paths, std::vector<Batch>& slots, int n_threads)
        : paths_(paths), slots_(slots), files_(paths.size()), window_(static_cast<size_t>(std::max(2, 2 * n_threads))) {
        for (size_t s = 0; s < slots.size(); s++) free_.push(static_cast<int>(s));
        for (int t = 0; t < std::max(1, n_threads); t++) workers_.emplace_back([this] { work(); });
        assembler_ = std::thread([this] { assemble(); });
    }
    ~MultiFileSource() override { join(); }
    int next_filled() override { return filled_.pop(); }
    void release(int slot) override { free_.push(slot); }
    void abort() override {
        { std::lock_guard<std::mutex> lk(mu_); aborted_ = true; }
        cv_.notify_all();
        for (size_t s = 0; s < slots_.size() + 2; s++) free_.push(-1);
    }
    void join() override {
        for (auto& w : workers_) if (w.joinable()) w.join();
        workers_.clear();
        if (assembler_.joinable()) assembler_.join();
    }
    const std::string& error() const override { return err_; }

  private:
    struct Parsed {
        std::vector<uint8_t> bases; std::vector<uint64_t> ends; std::vector<char> ids; std::vector<uint32_t> id_ends;
        std::string err; bool done = false;
    };
    void work() {
        for (;;) {
            const size_t f = claim_.fetch_add(1);
            if (f >= files_.size()) return;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return aborted_ || f < assembled_ + window_; });
                if (aborted_) return;
            }
            Parsed P;
            FastqReader reader(std::vector<std::string>{paths_[f]});
            FastqReader::View v;
            while (reader.next(v, P.err)) {
                P.bases.insert(P.bases.end(), v.seq, v.seq + v.seq_len);
                P.ends.push_back(P.bases.size());
                P.ids.insert(P.ids.end(), v.id, v.id + v.id_len);
                P.id_ends.push_back(static_cast<uint32_t>(P.ids.size()));
            }
            P.done = true;
            { std::lock_guard<std::mutex> lk(mu_); files_[f] = std::move(P); }
            cv_.notify_all();
        }
    }
    void assemble() {
        int cur = free_.pop();
        if (cur < 0) return;
        slots_[cur].clear();
        for (size_t f = 0; f < files_.size(); f++) {
            Parsed P;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return aborted_ || files_[f].done; });
                if (aborted_) return;
                P = std::move(files_[f]);
            }
            if (!P.err.empty()) { err_ = P.err; filled_.push(-2); return; }
            uint64_t b0 = 0; uint32_t i0 = 0;
            for (size_t r = 0; r < P.ends.size(); r++) {
                const size_t len = P.ends[r] - b0;
                Batch* B = &slots_[cur];
                if (len > B->cap_bytes) { err_ = "read longer than the batch buffer (raise --batch-mb)"; filled_.push(-2); return; }
                if (B->bytes + len > B->cap_bytes || !B->reserve_read()) {
                    filled_.push(cur);
                    cur = free_.pop();
                    if (cur < 0) return;
                    B = &slots_[cur];
                    B->clear();
                }
                B->append(P.ids.data() + i0, P.id_ends[r] - i0, reinterpret_cast<const char*>(P.bases.data()) + b0, len);
                b0 = P.ends[r]; i0 = P.id_ends[r];
            }
            { std::lock_guard<std::mutex> lk(mu_); assembled_ = f + 1; }
            cv_.notify_all();
        }
        if (slots_[cur].n_reads > 0) filled_.push(cur);
        filled_.push(-1);
    }
    std::vector<std::string> paths_;
    std::vector<Batch>& slots_;
    std::vector<Parsed> files_;
    size_t window_, assembled_ = 0;
    std::atomic<size_t> claim_{0};
    std::mutex mu_;
    std::condition_variable cv_;
    Channel<int> free_, filled_;
    std::vector<std::thread> workers_;
    std::thread assembler_;
    std::string err_;
    bool aborted_ = false;
};

// Slots + source for a run: plain files are cut into chunks of `chunk` FASTQ bytes (<= chunk/2 bases each) and parsed by the
// -t threads (at most 32); anything else (gzip, pipes, -t 1, --single-reader) goes through one reader thread.
struct Ingest {
    std::vector<Batch> slots;
    std::unique_ptr<BatchSource> source;
    bool parallel = false, multifile = false;
    size_t input_bytes = 0, chunk_bytes = 0;                // parallel source only
    std::chrono::steady_clock::time_point t_start;          // when the parsers were started (after the slots were allocated)
    // before_start (optional) runs once the plan is known (parallel, input_bytes, chunk_bytes) and the slots are allocated, right
    // before the parser threads start
    bool open(const Args& a, int in_flight, bool pinned, std::string& err, const std::function<bool(std::string&)>& before_start = nullptr) {
        const size_t cap_bytes = a.batch_mb << 20;
        const int parse_threads = std::min(32, std::max(1, a.threads));
        parallel = parse_threads > 1 && !a.single_reader && ParallelSource::usable(a.input);
        size_t chunk = a.chunk_kb ? a.chunk_kb << 10 : cap_bytes, n_chunks = 0;
        if (parallel) {
            for (const auto& p : a.input) { struct stat st; if (stat(p.c_str(), &st) == 0) { n_chunks += (static_cast<size_t>(st.st_size) + chunk - 1) / chunk; input_bytes += static_cast<size_t>(st.st_size); } }
            chunk_bytes = chunk;
        }
        // sequential: in flight + one filled + the one being filled; parallel: in flight + one per parser (fewer for small inputs)
        // slots beyond one per parser: the batches on the GPU(s), the ones queued for the writer thread, and slack against the
        // in-order hand-over (a parser that finishes early must not wait for the slot of a slower one)
        const size_t slack = std::getenv("BB_SLOT_SLACK") ? static_cast<size_t>(std::atoi(std::getenv("BB_SLOT_SLACK"))) : 4;
        const size_t n_slots = parallel ? std::max<size_t>(2, std::min<size_t>(in_flight + parse_threads + slack, n_chunks + 1)) : static_cast<size_t>(in_flight + 2);
        // a chunk holds at most chunk/2 bases plus the tail of the record that straddles its end (16 MB covers the longest reads)
        const size_t slot_bytes = parallel ? std::min(cap_bytes, chunk / 2) + (16u << 20) : cap_bytes;
        slots.resize(n_slots);
        const size_t cap_reads = std::max<size_t>(4096, slot_bytes / 256);         // Batch::reserve_read grows it for shorter reads
        for (auto& b : slots) if (!b.alloc(slot_bytes, cap_reads, pinned, !a.no_pack)) { err = pinned ? "pinned host allocation failed" : "host allocation failed"; return false; }
        multifile = !parallel && parse_threads > 1 && !a.single_reader && MultiFileSource::usable(a.input);
        if (before_start && !before_start(err)) return false;
        t_start = std::chrono::steady_clock::now();
        if (parallel) source.reset(new ParallelSource(a.input, slots, chunk, parse_threads));
        else if (multifile) source.reset(new MultiFileSource(a.input, slots, parse_threads));
        else source.reset(new SequentialSource(a.input, slots));
        return true;
    }
    void close() { if (source) { source->join(); source.reset(); } for (auto& b : slots) b.release(); slots.clear(); }
};

// --verbose: `<dir>/<step>.<unix ms>.log` with the step's final counters, like ProgressTracker::finish (progress.rs:102-144, 188-201)
void write_progress_log(const std::string& dir, const char* step, const std::vector<std::pair<const char*, unsigned long long>>& counts) {
    const long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
    const std::string path = (dir.empty() ? std::string(".") : dir) + "/" + step + "." + std::to_string(ms) + ".log";
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) { std::printf("Failed to create log file '%s'\n", path.c_str()); return; }
    std::fprintf(f, "step\tmetric\tcount\n");
    for (const auto& c : counts) std::fprintf(f, "%s\t%s\t%llu\n", step, c.first, c.second);
    std::fclose(f);
}
std::string parent_dir(const std::string& path) {
    const size_t p = path.find_last_of('/');
    return p == std::string::npos ? std::string(".") : (p == 0 ? std::string("/") : path.substr(0, p));
}

const char* kTypeNames[] = {"Ftag", "Rtag", "Fflank", "Rflank"};

// annotation.tsv rows are formatted by hand into a large buffer (a row is ~100 bytes; printf per row would cap the writer
// thread at ~2 M rows/s, below what one GPU produces)
struct RowWriter {
    FILE* f; std::vector<char> buf; size_t n = 0; bool failed = false;
    explicit RowWriter(FILE* file) : f(file), buf(8u << 20) {}
    void flush() { if (n && std::fwrite(buf.data(), 1, n, f) != n) failed = true; n = 0; }
    void room(size_t need) { if (n + need > buf.size()) flush(); if (need > buf.size()) buf.resize(need * 2); }
    void raw(const char* p, size_t len) { room(len); std::memcpy(buf.data() + n, p, len); n += len; }
    void raw_nocheck(const char* p, size_t len) { std::memcpy(buf.data() + n, p, len); n += len; }   // after room()
    void str(const char* p) { raw(p, std::strlen(p)); }
    void ch(char c) { buf[n++] = c; }                      // after room()
    void num(long long v) {                                // after room(24)
        char tmp[24]; int k = 0;
        unsigned long long u = v < 0 ? 0ull - static_cast<unsigned long long>(v) : static_cast<unsigned long long>(v);
        do { tmp[k++] = static_cast<char>('0' + u % 10); u /= 10; } while (u);
        if (v < 0) buf[n++] = '-';
        while (k) buf[n++] = tmp[--k];
    }
};

// teardown (optional): the release of the pinned slots and of the GPU contexts is handed to this thread, so that `barbell kit` can
// start the stages that consume annotation.tsv while the driver unpins ~1 GB and destroys the contexts
int run_annotate(const Args& a, const std::string& out_path, std::thread* teardown = nullptr) {
    char err[512] = {0};
    bb_groupset* gs = nullptr;
    int rc;
    if (!a.kit.empty()) {
        rc = bb_groups_from_kit(a.kit.c_str(), a.use_extended, &gs, err, sizeof err);
    } else {
        if (a.queries.empty()) { std::snprintf(err, sizeof err, "--queries is required unless --kit is provided"); rc = BB_ERR_INVALID; }
        else if (a.queries.size() != a.barcode_types.size()) { std::snprintf(err, sizeof err, "--queries and --barcode-types must have the same number of values"); rc = BB_ERR_INVALID; }
        else {
            std::vector<const char*> paths; std::vector<int32_t> types;
            rc = BB_OK;
            for (size_t i = 0; i < a.queries.size(); i++) {
                paths.push_back(a.queries[i].c_str());
                if (a.barcode_types[i] == "Ftag") types.push_back(BB_FTAG);
                else if (a.barcode_types[i] == "Rtag") types.push_back(BB_RTAG);
                else { std::snprintf(err, sizeof err, "Unknown barcode type: %s, use one of: Ftag, Rtag", a.barcode_types[i].c_str()); rc = BB_ERR_INVALID; }
            }
            if (rc == BB_OK) rc = bb_groups_from_fasta(paths.data(), types.data(), static_cast<int32_t>(paths.size()), &gs, err, sizeof err);
        }
    }
    if (rc != BB_OK) { std::printf("Error during processing: %s\n", err); return rc; }
    bb_groups_set_flank_threshold(gs, a.flank_max_errors);
    const int n_groups = bb_groups_count(gs);
    const bb_group* groups = bb_groups_data(gs);
    if (a.flank_max_errors < 0)
        for (int g = 0; g < n_groups; g++) std::printf("Auto edit flank cut off: %d\n", groups[g].k_flank);   // annotator.rs:224
    if (a.input.empty()) { std::printf("Error during processing: No FASTQ input files provided\n"); bb_groups_free(gs); return BB_ERR_IO; }

    const int n_gpus = a.gpus < 1 ? 1 : a.gpus;
    std::vector<bb_ctx*> ctx(n_gpus, nullptr);
    const auto t_setup = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
    for (int d = 0; d < n_gpus; d++) {
        bb_opts o{};
        o.device = d; o.alpha = a.alpha; o.min_score = a.min_score; o.min_score_diff = a.min_score_diff; o.policy = a.policy;
        rc = bb_create(&o, &ctx[d], err, sizeof err);
        if (rc == BB_OK) { rc = bb_set_groups(ctx[d], groups, n_groups); if (rc != BB_OK) std::snprintf(err, sizeof err, "%s", bb_last_error(ctx[d])); }
        if (rc != BB_OK) { std::printf("Error during processing: %s\n", err); for (auto* c : ctx) bb_destroy(c); bb_groups_free(gs); return rc; }
    }

    const double ctx_secs = since(t_setup);
    FILE* out = std::fopen(out_path.c_str(), "w");
    if (!out) { std::printf("Error during processing: cannot open %s\n", out_path.c_str()); for (auto* c : ctx) bb_destroy(c); bb_groups_free(gs); return BB_ERR_IO; }
    RowWriter W(out);

    Ingest ingest;
    double reserve_secs = 0;
    {
        // plain files of some size: the engines allocate their buffers and load the kernels for batches of one chunk now -- all engines
        // of all GPUs at once, and before the parser threads start (an allocation changes the address space, which waits for every
        // page fault of 16 threads streaming through a mapped file) -- instead of one after the other on the first batches
        auto reserve = [&](std::string& e) {
            if (!ingest.parallel || ingest.input_bytes < (64u << 20)) return true;
            const auto tr = std::chrono::steady_clock::now();
            const uint64_t max_bases = std::min<uint64_t>(ingest.chunk_bytes / 2 + ingest.chunk_bytes / 32, ingest.input_bytes / 2 + (1u << 20));
            std::vector<std::thread> th;
            std::vector<int> rcs(n_gpus, BB_OK);
            for (int d = 0; d < n_gpus; d++) th.emplace_back([&, d] { rcs[d] = bb_reserve(ctx[d], static_cast<uint32_t>(max_bases / 1024 + 1), max_bases); });
            for (auto& t : th) t.join();
            for (int d = 0; d < n_gpus; d++) if (rcs[d] != BB_OK) { e = bb_last_error(ctx[d]); return false; }
            reserve_secs = since(tr);
            return true;
        };
        std::string ierr;
        if (!ingest.open(a, BB_MAX_INFLIGHT * n_gpus, true, ierr, reserve)) {
            std::printf("Error during processing: %s\n", ierr.c_str());
            ingest.close(); std::fclose(out); for (auto* c : ctx) bb_destroy(c); bb_groups_free(gs);
            return BB_ERR_CUDA;
        }
    }
    std::vector<Batch>& slots = ingest.slots;
    const size_t n_slots_used = slots.size();
    BatchSource* source = ingest.source.get();
    const auto t0 = ingest.t_start;                        // the stream phase starts when the parsers start
    struct rusage ru0; getrusage(RUSAGE_SELF, &ru0);
    const double setup_secs = std::chrono::duration<double>(t0 - t_setup).count();

    struct Flight { int slot, dev; };
    std::deque<Flight> flight;
    uint64_t total_reads = 0, total_rows = 0, kept = 0, submitted = 0;
    bool header_written = false;
    // labels are looked up once per (group, barcode), not per row
    std::vector<std::vector<const char*>> labels(n_groups);
    for (int g = 0; g < n_groups; g++) { labels[g].resize(groups[g].n_barcodes); for (int b = 0; b < groups[g].n_barcodes; b++) labels[g][b] = bb_groups_label(gs, g, b); }
    double t_collect = 0, t_format = 0, t_submit = 0, t_source = 0;    // where the main / writer thread's time goes (--verbose)
    std::vector<std::vector<size_t>> label_len(n_groups);
    for (int g = 0; g < n_groups; g++) for (const char* l : labels[g]) label_len[g].push_back(std::strlen(l));
    // Rows are written by their own thread: the main thread only moves batches between the parsers and the GPU(s), and formatting
    // overlaps the waits for the GPU.  (The rows are copied out of the engine's buffer into the batch slot, ~1 MB per batch: the
    // engine is free for its next batch whatever the writer's backlog.)
    struct WriteJob { int slot; const bb_row* rows; uint64_t n_rows; };
    Channel<WriteJob> write_q;
    // batches in flight per GPU: 2 keep it busy (measured: 4 are no faster, the parsers set the pace); BB_CLI_INFLIGHT = tuning knob
    const int per_gpu = std::getenv("BB_CLI_INFLIGHT") ? std::max(1, std::min(BB_MAX_INFLIGHT, std::atoi(std::getenv("BB_CLI_INFLIGHT")))) : 2;
    auto write_rows = [&](const WriteJob& j) {
        const auto tf0 = std::chrono::steady_clock::now();
        const Batch& B = slots[j.slot];
        uint32_t last = UINT32_MAX;
        if (!header_written && j.n_rows) {
            W.str("read_id\tread_len\trel_dist_to_end\tread_start_bar\tread_end_bar\tread_start_flank\tread_end_flank\t"
                  "bar_start\tbar_end\tmatch_type\tflank_cost\tbarcode_cost\tlabel\tstrand\tcuts\n");
            header_written = true;
        }
        for (uint64_t i = 0; i < j.n_rows; i++) {
            const bb_row& w = j.rows[i];
            if (w.read_idx != last) { kept++; last = w.read_idx; }
            const size_t idl = B.id_off[w.read_idx + 1] - B.id_off[w.read_idx];
            const char* label = w.label_idx < 0 ? "flank" : labels[w.group_idx][w.label_idx];
            const size_t ll = w.label_idx < 0 ? 5 : label_len[w.group_idx][w.label_idx];
            W.room(idl + ll + 320);
            W.raw_nocheck(B.id_chars.data() + B.id_off[w.read_idx], idl);
            W.ch('\t'); W.num(w.read_len);
            W.ch('\t'); W.num(w.rel_dist_to_end);
            W.ch('\t'); W.num(w.read_start_bar); W.ch('\t'); W.num(w.read_end_bar);
            W.ch('\t'); W.num(w.read_start_flank); W.ch('\t'); W.num(w.read_end_flank);
            W.ch('\t'); W.num(w.bar_start); W.ch('\t'); W.num(w.bar_end);
            W.ch('\t'); W.raw_nocheck(kTypeNames[w.match_type & 3], (w.match_type & 2) ? 6 : 4);
            W.ch('\t'); W.num(w.flank_cost); W.ch('\t'); W.num(w.barcode_cost);
            W.ch('\t'); W.raw_nocheck(label, ll);
            if (w.strand) W.raw_nocheck("\tRc\t\n", 5); else W.raw_nocheck("\tFwd\t\n", 6);
        }
        total_rows += j.n_rows;
        t_format += since(tf0);
    };
    std::thread writer([&] {
        for (;;) {
            const WriteJob j = write_q.pop();
            if (j.slot < 0) return;
            write_rows(j);
            source->release(j.slot);
        }
    });
    bool writer_joined = false;
    auto join_writer = [&] { if (!writer_joined) { write_q.push({-1, nullptr, 0}); writer.join(); writer_joined = true; } };
    auto collect_one = [&](bool write) -> int {
        const Flight f = flight.front(); flight.pop_front();
        uint64_t tag = 0, n_rows = 0; const bb_row* rows = nullptr;
        const auto tc0 = std::chrono::steady_clock::now();
        int r = bb_collect(ctx[f.dev], &tag, &rows, &n_rows);
        t_collect += since(tc0);
        if (r != BB_OK) { if (write) std::printf("Error during processing: %s\n", bb_last_error(ctx[f.dev])); source->release(f.slot); return r; }
        if (!write) { source->release(f.slot); return BB_OK; }
        Batch& B = slots[f.slot];
        B.rows.assign(rows, rows + n_rows);
        write_q.push({f.slot, B.rows.data(), n_rows});
        return BB_OK;
    };

    rc = BB_OK;
    for (;;) {
        const auto ts0 = std::chrono::steady_clock::now();
        const int cur = source->next_filled();
        t_source += since(ts0);
        if (cur == -1) break;
        if (cur == -2) { std::printf("Error during processing: %s\n", source->error().c_str()); rc = BB_ERR_IO; break; }
        Batch& B = slots[cur];
        if (B.exc_overflow) {
            std::printf("Error during processing: out of memory for the list of bases other than A/C/G/T: re-run with --no-pack\n");
            rc = BB_ERR_INVALID; source->release(cur); break;
        }
        const int dev = static_cast<int>(submitted % n_gpus);
        size_t on_dev = 0; for (const auto& f : flight) on_dev += f.dev == dev;
        while (on_dev >= static_cast<size_t>(per_gpu) && rc == BB_OK) { const int d0 = flight.front().dev; rc = collect_one(true); if (d0 == dev) on_dev--; }
        if (rc != BB_OK) { source->release(cur); break; }
        const auto tb0 = std::chrono::steady_clock::now();
        rc = B.packed ? bb_submit_packed(ctx[dev], B.bases, B.bytes, B.exc, B.n_exc, B.offsets, B.n_reads, submitted)
                      : bb_submit(ctx[dev], B.bases, B.offsets, B.n_reads, submitted);
        t_submit += since(tb0);
        if (rc != BB_OK) { std::printf("Error during processing: %s\n", bb_last_error(ctx[dev])); source->release(cur); break; }
        flight.push_back({cur, dev});
        submitted++;
        total_reads += B.n_reads;
    }
    while (!flight.empty() && rc == BB_OK) rc = collect_one(true);
    if (rc != BB_OK) {
        source->abort();                             // unblock the producers ...
        while (!flight.empty()) collect_one(false);  // ... and wait for every batch still in flight: its buffers are being read by the workers / the DMA engine
    }
    join_writer();
    source->join();
    double parse_secs = 0, wait_secs = 0;
    if (ingest.parallel) static_cast<ParallelSource*>(source)->thread_seconds(parse_secs, wait_secs);
    W.flush();
    const bool close_failed = std::fclose(out) != 0;
    if ((W.failed || close_failed) && rc == BB_OK) { std::printf("Error during processing: write to %s failed\n", out_path.c_str()); rc = BB_ERR_IO; }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    struct rusage ru1; getrusage(RUSAGE_SELF, &ru1);
    if (rc == BB_OK && a.verbose)
        write_progress_log(parent_dir(out_path), "annotate", {{"Total:", total_reads}, {"Kept:", kept}, {"Dropped:", total_reads - kept}});
    if (rc == BB_OK) {
        std::printf("Total: %llu  Kept: %llu  Dropped: %llu  (rows: %llu, %.3f s, %.0f reads/s)\n", static_cast<unsigned long long>(total_reads),
                    static_cast<unsigned long long>(kept), static_cast<unsigned long long>(total_reads - kept),
                    static_cast<unsigned long long>(total_rows), secs, secs > 0 ? total_reads / secs : 0.0);
    }
    const auto t_down = std::chrono::steady_clock::now();
    if (teardown) {
        ingest.source->join(); ingest.source.reset();        // (the source refers to the slot vector: it goes first, here)
        auto held = std::make_shared<std::vector<Batch>>(std::move(ingest.slots));
        *teardown = std::thread([held, ctx, gs]() { for (auto& b : *held) b.release(); for (auto* c : ctx) bb_destroy(c); bb_groups_free(gs); });
    } else {
        ingest.close();
        for (auto* c : ctx) bb_destroy(c);
        bb_groups_free(gs);
    }
    // (pinning the slots on a second thread beside the context set-up was tried: the driver serialises the two, no gain)
    if (a.verbose) {
        auto tv = [](const timeval& x) { return static_cast<double>(x.tv_sec) + 1e-6 * static_cast<double>(x.tv_usec); };
        std::fprintf(stderr, "[timing] setup %.2f s (CUDA context + engine %.2f s, %zu pinned batch slots + engine buffers %.2f s of which reserve %.3f s), stream %.2f s (all threads: user %.2f s, "
                             "kernel %.2f s, %ld minor faults; parser threads: %.2f s parsing, %.2f s waiting for a slot; main thread: %.3f s waiting for a parsed "
                             "batch, %.3f s submitting, %.3f s waiting for the GPU; writer thread: %.3f s writing rows), teardown %.2f s%s\n",
                     setup_secs, ctx_secs, n_slots_used, setup_secs - ctx_secs, reserve_secs, secs, tv(ru1.ru_utime) - tv(ru0.ru_utime), tv(ru1.ru_stime) - tv(ru0.ru_stime),
                     ru1.ru_minflt - ru0.ru_minflt, parse_secs, wait_secs, t_source, t_submit, t_collect, t_format, since(t_down),
                     teardown ? " (continues in the background)" : "");
    }
    return rc;
}



// `barbell fastq-stats -i files...`: parse only (no GPU): records, bases, FNV-1a of ids and sequences (reader self-check)
int run_fastq_stats(const Args& a) {
    Ingest ingest;
    std::string err;
    if (!ingest.open(a, 1, false, err)) { std::printf("Error during processing: %s\n", err.c_str()); return 1; }
    uint64_t n = 0, bases = 0, batches = 0, h = 1469598103934665603ull;
    auto mix = [&](const char* p, size_t len) { for (size_t i = 0; i < len; i++) { h ^= static_cast<unsigned char>(p[i]); h *= 1099511628211ull; } h ^= 0xff; h *= 1099511628211ull; };
    int rc = 0;
    for (;;) {
        const int cur = ingest.source->next_filled();
        if (cur == -1) break;
        if (cur == -2) { std::printf("Error during processing: %s\n", ingest.source->error().c_str()); rc = 1; ingest.source->abort(); break; }
        const Batch& B = ingest.slots[cur];
        if (B.exc_overflow) { std::printf("Error during processing: exception list overflow (re-run with --no-pack)\n"); rc = 1; ingest.source->abort(); break; }
        std::vector<char> text;
        if (a.count_only) { n += B.n_reads; bases += B.bytes; batches++; ingest.source->release(cur); continue; }   // parser throughput only
        if (B.packed) {
            // the packed form decoded the way the device does it (k_unpack_crumbs + k_patch_exceptions): one letter per base set
            text.resize(B.bytes);
            for (size_t i = 0; i < B.bytes; i++) text[i] = "ACGT"[(B.bases[i >> 2] >> (2 * (i & 3))) & 3];
            for (uint64_t q = 0; q < B.n_exc; q++) text[B.exc[q] >> 4] = "XACMGRSVTWYHKDBN"[B.exc[q] & 15];
        }
        const char* seqs = B.packed ? text.data() : reinterpret_cast<const char*>(B.bases);
        for (uint32_t r = 0; r < B.n_reads; r++) {
            mix(B.id_chars.data() + B.id_off[r], B.id_off[r + 1] - B.id_off[r]);
            mix(seqs + B.offsets[r], B.offsets[r + 1] - B.offsets[r]);
        }
        n += B.n_reads; bases += B.bytes; batches++;
        ingest.source->release(cur);
    }
    const char* kind = ingest.parallel ? "parallel" : ingest.multifile ? "multifile" : "sequential";
    if (a.count_only && ingest.parallel) {
        ingest.source->join();
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - ingest.t_start).count();
        double ps = 0, ws = 0; static_cast<ParallelSource*>(ingest.source.get())->thread_seconds(ps, ws);
        std::fprintf(stderr, "[timing] %.3f s, %.0f reads/s; parser threads: %.2f s parsing, %.2f s waiting for a slot; %zu slots\n", secs, static_cast<double>(n) / secs, ps, ws, ingest.slots.size());
    }
    ingest.close();
    if (rc == 0) std::printf("records=%llu bases=%llu fnv=%016llx batches=%llu reader=%s form=%s\n", static_cast<unsigned long long>(n), static_cast<unsigned long long>(bases),
                             static_cast<unsigned long long>(h), static_cast<unsigned long long>(batches), kind, a.no_pack ? "bytes" : "packed");
    return rc;
}

// `barbell filter` (bin/main.rs:340-354, filter_from_text_file filter.rs:137-181)
int run_filter_cmd(const Args& a) {
    std::printf("Starting filtering...\n");
    char err[1024] = {0};
    std::string msg;
    std::vector<std::string> pats;
    if (a.input.size() != 1 || !a.output_given || a.pattern_files.empty()) usage("filter needs -i <annotation.tsv> -o <output> -f <pattern file>...");
    for (const auto& pf : a.pattern_files) {
        FILE* f = std::fopen(pf.c_str(), "r");
        if (!f) { msg = pf + ": cannot read pattern file"; break; }
        char line[8192];
        while (std::fgets(line, sizeof line, f)) {
            std::string l(line);
            size_t b = 0, e = l.size();
            while (b < e && std::isspace(static_cast<unsigned char>(l[b]))) b++;
            while (e > b && std::isspace(static_cast<unsigned char>(l[e - 1]))) e--;
            if (e > b) pats.push_back(l.substr(b, e - b));
        }
        std::fclose(f);
    }
    if (msg.empty() && pats.empty()) msg = "No filter patterns found";
    if (msg.empty()) {
        std::vector<const char*> pp;
        for (const auto& s : pats) pp.push_back(s.c_str());
        uint64_t counts[3] = {0, 0, 0};
        const int rc = bb_filter(a.input[0].c_str(), a.output.c_str(), a.dropped.empty() ? nullptr : a.dropped.c_str(), pp.data(),
                                 static_cast<int32_t>(pp.size()), counts, err, sizeof err);
        if (rc == BB_OK) {
            if (a.verbose) write_progress_log(parent_dir(a.output), "filter", {{"Total:", counts[0]}, {"Kept:", counts[1]}, {"Dropped:", counts[2]}});
            std::printf("Total: %llu  Kept: %llu  Dropped: %llu reads\n", static_cast<unsigned long long>(counts[0]),
                        static_cast<unsigned long long>(counts[1]), static_cast<unsigned long long>(counts[2]));
            std::printf("Filtering successful!\n");
            return 0;
        }
        msg = err;
        if (msg.rfind("Pattern parse error", 0) == 0 || msg.rfind("Flank is not valid", 0) == 0) { std::fprintf(stderr, "%s\n", msg.c_str()); return 101; }   // the reference panics here
    }
    std::printf("Filtering failed: %s\n", msg.c_str());
    return 0;
}

bb_trim_opts trim_opts_from(const Args& a, bool kit_defaults) {
    bb_trim_opts o{};
    if (kit_defaults) {                                   // use_kit.rs:87-99
        o.add_labels = 1; o.add_orientation = 0; o.add_flank = 0; o.sort_labels = 0; o.only_side = 1;
        o.write_full_header = 1; o.skip_trim = 0; o.flip = 0;
    } else {                                              // bin/main.rs:372-384
        o.add_labels = !a.no_label; o.add_orientation = !a.no_orientation; o.add_flank = !a.no_flanks; o.sort_labels = a.sort_labels;
        o.only_side = a.only_side == "left" ? 1 : a.only_side == "right" ? 2 : 0;
        o.write_full_header = 1; o.skip_trim = a.skip_trim; o.flip = a.flip;
    }
    o.gzip = a.gzip;
    o.failed_out = a.failed_out.empty() ? nullptr : a.failed_out.c_str();
    o.threads = std::min(16, std::max(1, a.threads));     // the reference's trim reads on one thread (trim.rs:375-383); the output is the same
    return o;
}

int run_trim(const std::string& filtered, const std::vector<std::string>& reads, const std::string& out_dir, const bb_trim_opts& o, std::string& msg,
             bool verbose = false) {
    std::vector<const char*> rp;
    for (const auto& r : reads) rp.push_back(r.c_str());
    uint64_t counts[4] = {0, 0, 0, 0};
    char err[1024] = {0};
    const int rc = bb_trim(filtered.c_str(), rp.data(), static_cast<int32_t>(rp.size()), out_dir.c_str(), &o, counts, err, sizeof err);
    if (rc != BB_OK) { msg = err; return rc; }
    if (verbose) write_progress_log(out_dir, "trim", {{"Total:", counts[0]}, {"Kept:", counts[1]}, {"Kept split:", counts[2]}, {"Failed:", counts[3]}});
    std::printf("Total: %llu  Trimmed: %llu  Trimmed split: %llu  Failed trims: %llu reads\n", static_cast<unsigned long long>(counts[0]),
                static_cast<unsigned long long>(counts[1]), static_cast<unsigned long long>(counts[2]), static_cast<unsigned long long>(counts[3]));
    return BB_OK;
}

// `barbell kit` (use_kit.rs:11-109): annotate -> inspect -> filter -> trim with fixed file names in the output folder
int run_kit(const Args& a) {
    char err[1024] = {0}, name[64] = {0}, ranges[256] = {0};
    int dbl = 0;
    ::mkdir(a.output.c_str(), 0755);
    if (bb_kit_info(a.kit.c_str(), name, sizeof name, ranges, sizeof ranges, &dbl, err, sizeof err) != BB_OK) {
        std::fprintf(stderr, "%s\n", err);               // get_kit_info panics on an unknown kit (kits.rs:704)
        return 101;
    }
    std::printf("\nKit info\nKit name: %s\nKit type: %s\n", name, a.maximize ? "Maximize" : "Safe");
    std::string r(ranges);
    for (size_t p = 0; p < r.size();) { size_t q = r.find("; ", p); if (q == std::string::npos) q = r.size(); std::printf("Barcodes: %s\n", r.substr(p, q - p).c_str()); p = q + 2; }
    std::printf("\nAnnotating reads...\n");
    const std::string anno = a.output + "/annotation.tsv", filtered = a.output + "/filtered.tsv";
    std::thread teardown;
    struct Join { std::thread& t; ~Join() { if (t.joinable()) t.join(); } } join_teardown{teardown};
    int rc = run_annotate(a, anno, &teardown);
    if (rc != BB_OK) { std::printf("Demultiplexing failed: annotate stage\n"); return 0; }
    // inspect and filter both only read annotation.tsv: the filter runs beside the inspection (its messages are printed in the reference's order)
    const char* const* pats = nullptr; int32_t n_pats = 0;
    bb_kit_filter_patterns(dbl, a.maximize, &pats, &n_pats);
    uint64_t counts[3] = {0, 0, 0};
    char ferr[1024] = {0};
    int frc = BB_OK;
    std::thread filt([&] { frc = bb_filter(anno.c_str(), filtered.c_str(), nullptr, pats, n_pats, counts, ferr, sizeof ferr); });
    std::printf("\nTop 10 most common patterns\n");
    rc = bb_inspect(anno.c_str(), 10, (a.output + "/pattern_per_read.tsv").c_str(), 250, err, sizeof err);
    filt.join();
    if (rc != BB_OK) { std::printf("Demultiplexing failed: %s\n", err); return 0; }
    std::printf("Want to see more patterns? Run: `barbell inspect %s/annotation.tsv -n 100`\n", a.output.c_str());
    std::printf("\nFiltering reads...\n");
    if (frc != BB_OK) { std::printf("Demultiplexing failed: %s\n", ferr); return 0; }
    if (a.verbose) write_progress_log(a.output, "filter", {{"Total:", counts[0]}, {"Kept:", counts[1]}, {"Dropped:", counts[2]}});
    std::printf("Total: %llu  Kept: %llu  Dropped: %llu reads\n", static_cast<unsigned long long>(counts[0]),
                static_cast<unsigned long long>(counts[1]), static_cast<unsigned long long>(counts[2]));
    std::printf("\nTrimming reads...\n");
    std::string msg;
    rc = run_trim(filtered, a.input, a.output, trim_opts_from(a, true), msg, a.verbose);
    if (rc != BB_OK) { std::printf("Demultiplexing failed: %s\n", msg.c_str()); return 0; }
    std::printf("\nDone!\n");
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    const Args a = parse(argc, argv);
    if (a.cmd == "fastq-stats") return run_fastq_stats(a);
    if (a.cmd == "annotate") {
        std::printf("Starting annotation...\n");
        if (run_annotate(a, a.output) == BB_OK) std::printf("Annotation complete!\n");
        return 0;                                                   // the reference exits 0 even on errors (bin/main.rs:301-304)
    }
    if (a.cmd == "filter") return run_filter_cmd(a);
    if (a.cmd == "trim") {
        std::printf("Starting trimming...\n");
        if (a.input.size() != 1 || !a.output_given) usage("trim needs -i <filtered.tsv> -r <fastq>... -o <folder>");
        if (!a.only_side.empty() && a.only_side != "left" && a.only_side != "right") usage("--only-side takes left or right");
        if (!a.only_side.empty() && a.sort_labels) usage("--only-side cannot be used with --sort-labels");
        std::string msg;
        if (run_trim(a.input[0], a.reads, a.output, trim_opts_from(a, false), msg, a.verbose) == BB_OK) std::printf("Trimming complete!\n");
        else std::printf("Trimming failed: %s\n", msg.c_str());
        return 0;
    }
    if (a.cmd == "inspect") {
        std::printf("Inspecting...\n");
        if (a.input.size() != 1) usage("inspect needs -i <annotation.tsv>");
        char err[1024] = {0};
        if (bb_inspect(a.input[0].c_str(), a.top_n, a.read_pattern_out.empty() ? nullptr : a.read_pattern_out.c_str(), a.bucket_size, err, sizeof err) == BB_OK)
            std::printf("Inspection complete!\n");
        else std::printf("Inspection failed: %s\n", err);
        return 0;
    }
    if (a.kit.empty()) usage("kit needs -k <KIT>");
    return run_kit(a);
}
