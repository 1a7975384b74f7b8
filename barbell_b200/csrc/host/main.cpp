// `barbell` command line: the reference's `annotate` and `kit` subcommands (bin/main.rs:61-112, 211-263, 274-339) on top
// of the C ABI.  FASTQ(.gz) records are parsed into page-locked batch buffers and pushed through bb_submit/bb_collect
// (two batches in flight per GPU); rows are written as annotation.tsv in input order (reference column order,
// src/annotate/searcher.rs:31-64; the header is written with the first row, so a run without hits leaves an empty
// file exactly like the reference's csv writer, annotator.rs:20-24).
// Not part of this build: inspect / filter / trim; `kit` runs the annotate stage only.
#include <sys/stat.h>
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../../include/barbell_b200.h"

namespace {

struct Args {
    std::string cmd, output = "output.tsv", kit;
    std::vector<std::string> input, queries, barcode_types{"Ftag"};
    int threads = 10, flank_max_errors = -1, gpus = 1;
    bool verbose = false, use_extended = false, maximize = false, gzip = false;
    double min_score = 0.2, min_score_diff = 0.1;
    float alpha = 0.4f;
    size_t batch_mb = 256;
};

[[noreturn]] void usage(const char* msg) {
    if (msg) std::fprintf(stderr, "error: %s\n", msg);
    std::fprintf(stderr,
        "barbell (B200 build of the annotate path)\n"
        "  barbell annotate -i <fastq>... [-o output.tsv] (--kit <KIT> | -q <fasta>... [-b Ftag|Rtag ...])\n"
        "                   [-t N] [--flank-max-errors INT] [--min-score F] [--min-score-diff F] [--alpha F]\n"
        "                   [--use-extended] [--verbose] [--gpus N] [--batch-mb MB]\n"
        "  barbell kit -k <KIT> -i <fastq>... -o <folder> [same options]   (annotate stage only)\n");
    std::exit(msg ? 2 : 0);
}

bool is_flag(const std::string& s) { return s.size() > 1 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9'); }

Args parse(int argc, char** argv) {
    Args a;
    if (argc < 2) usage(nullptr);
    a.cmd = argv[1];
    if (a.cmd == "-h" || a.cmd == "--help") usage(nullptr);
    if (a.cmd != "annotate" && a.cmd != "kit") usage("only the `annotate` and `kit` subcommands exist in this build");
    bool types_given = false;
    for (int i = 2; i < argc; i++) {
        std::string f = argv[i];
        auto many = [&](std::vector<std::string>& dst) { while (i + 1 < argc && !is_flag(argv[i + 1])) dst.push_back(argv[++i]); };
        auto one = [&]() -> std::string { if (i + 1 >= argc) usage(("missing value for " + f).c_str()); return argv[++i]; };
        if (f == "-i" || f == "--input") many(a.input);
        else if (f == "-q" || f == "--queries") many(a.queries);
        else if (f == "-b" || f == "--barcode-types") { if (!types_given) a.barcode_types.clear(); types_given = true; many(a.barcode_types); }
        else if (f == "-o" || f == "--output") a.output = one();
        else if (f == "-t" || f == "--threads") a.threads = std::atoi(one().c_str());
        else if (f == "--kit" || f == "-k") a.kit = one();
        else if (f == "--flank-max-errors") a.flank_max_errors = std::atoi(one().c_str());
        else if (f == "--min-score") a.min_score = std::atof(one().c_str());
        else if (f == "--min-score-diff") a.min_score_diff = std::atof(one().c_str());
        else if (f == "--alpha") a.alpha = static_cast<float>(std::atof(one().c_str()));
        else if (f == "--gpus") a.gpus = std::atoi(one().c_str());
        else if (f == "--batch-mb") a.batch_mb = static_cast<size_t>(std::atol(one().c_str()));
        else if (f == "--failed-out") (void)one();
        else if (f == "--verbose") a.verbose = true;
        else if (f == "--use-extended") a.use_extended = true;
        else if (f == "--maximize") a.maximize = true;
        else if (f == "--gzip") a.gzip = true;
        else if (f == "-h" || f == "--help") usage(nullptr);
        else usage(("unknown argument " + f).c_str());
    }
    return a;
}

// FASTQ(.gz) reader over several files (reference io.rs:27-32: one paraseq Collection over all paths)
class FastqReader {
  public:
    explicit FastqReader(std::vector<std::string> paths) : paths_(std::move(paths)) {}
    ~FastqReader() { if (gz_) gzclose(gz_); }
    // next record: header (without '@'), sequence; false at the end of the last file
    bool next(std::string& header, std::string& seq, std::string& err) {
        for (;;) {
            if (!gz_) {
                if (file_ >= paths_.size()) return false;
                gz_ = gzopen(paths_[file_].c_str(), "rb");
                if (!gz_) { err = "Failed to open FASTQ input: " + paths_[file_]; return false; }
                gzbuffer(gz_, 1 << 20);
                pos_ = len_ = 0;
            }
            if (!line(header)) { gzclose(gz_); gz_ = nullptr; file_++; continue; }
            if (header.empty()) continue;
            if (header[0] != '@') { err = "malformed FASTQ record in " + paths_[file_]; return false; }
            header.erase(0, 1);
            std::string plus, qual;
            if (!line(seq) || !line(plus) || !line(qual)) { err = "truncated FASTQ record in " + paths_[file_]; return false; }
            return true;
        }
    }

  private:
    bool line(std::string& out) {
        out.clear();
        for (;;) {
            if (pos_ == len_) {
                const int n = gzread(gz_, buf_, sizeof buf_);
                if (n <= 0) return !out.empty();
                pos_ = 0; len_ = static_cast<size_t>(n);
            }
            const char* p = static_cast<const char*>(std::memchr(buf_ + pos_, '\n', len_ - pos_));
            if (p) {
                out.append(buf_ + pos_, p - (buf_ + pos_));
                pos_ = static_cast<size_t>(p - buf_) + 1;
                if (!out.empty() && out.back() == '\r') out.pop_back();
                return true;
            }
            out.append(buf_ + pos_, len_ - pos_);
            pos_ = len_;
        }
    }
    std::vector<std::string> paths_;
    size_t file_ = 0;
    gzFile gz_ = nullptr;
    char buf_[1 << 16];
    size_t pos_ = 0, len_ = 0;
};

struct Batch {
    uint8_t* bases = nullptr;
    uint64_t* offsets = nullptr;
    size_t cap_bytes = 0, cap_reads = 0, bytes = 0;
    uint32_t n_reads = 0;
    std::vector<std::string> ids;
    bool alloc(size_t cb, size_t cr) {
        cap_bytes = cb; cap_reads = cr;
        bases = static_cast<uint8_t*>(bb_host_alloc(cb + 64));
        offsets = static_cast<uint64_t*>(bb_host_alloc((cr + 1) * sizeof(uint64_t)));
        return bases && offsets;
    }
    void clear() { bytes = 0; n_reads = 0; ids.clear(); if (offsets) offsets[0] = 0; }
    void release() { bb_host_free(bases); bb_host_free(offsets); bases = nullptr; offsets = nullptr; }
};

const char* kTypeNames[] = {"Ftag", "Rtag", "Fflank", "Rflank"};

int run_annotate(const Args& a, const std::string& out_path) {
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
    for (int d = 0; d < n_gpus; d++) {
        bb_opts o{};
        o.device = d; o.alpha = a.alpha; o.min_score = a.min_score; o.min_score_diff = a.min_score_diff;
        rc = bb_create(&o, &ctx[d], err, sizeof err);
        if (rc == BB_OK) { rc = bb_set_groups(ctx[d], groups, n_groups); if (rc != BB_OK) std::snprintf(err, sizeof err, "%s", bb_last_error(ctx[d])); }
        if (rc != BB_OK) { std::printf("Error during processing: %s\n", err); for (auto* c : ctx) bb_destroy(c); bb_groups_free(gs); return rc; }
    }

    FILE* out = std::fopen(out_path.c_str(), "w");
    if (!out) { std::printf("Error during processing: cannot open %s\n", out_path.c_str()); for (auto* c : ctx) bb_destroy(c); bb_groups_free(gs); return BB_ERR_IO; }
    static char outbuf[1 << 22];
    std::setvbuf(out, outbuf, _IOFBF, sizeof outbuf);

    const size_t cap_bytes = a.batch_mb << 20, cap_reads = 1u << 22;
    const int n_slots = 2 * n_gpus + 1;                 // 2 in flight per GPU + the one being filled
    std::vector<Batch> slots(n_slots);
    for (auto& b : slots) if (!b.alloc(cap_bytes, cap_reads)) { std::printf("Error during processing: pinned host allocation failed\n"); return BB_ERR_CUDA; }

    struct Flight { int slot, dev; };
    std::deque<Flight> flight;
    uint64_t total_reads = 0, total_rows = 0, kept = 0, submitted = 0;
    bool header_written = false;
    auto t0 = std::chrono::steady_clock::now();

    auto collect_one = [&]() -> int {
        const Flight f = flight.front(); flight.pop_front();
        uint64_t tag = 0, n_rows = 0; const bb_row* rows = nullptr;
        int r = bb_collect(ctx[f.dev], &tag, &rows, &n_rows);
        if (r != BB_OK) { std::printf("Error during processing: %s\n", bb_last_error(ctx[f.dev])); return r; }
        const Batch& B = slots[f.slot];
        uint32_t last = UINT32_MAX;
        for (uint64_t i = 0; i < n_rows; i++) {
            const bb_row& w = rows[i];
            if (!header_written) {
                std::fputs("read_id\tread_len\trel_dist_to_end\tread_start_bar\tread_end_bar\tread_start_flank\tread_end_flank\t"
                           "bar_start\tbar_end\tmatch_type\tflank_cost\tbarcode_cost\tlabel\tstrand\tcuts\n", out);
                header_written = true;
            }
            if (w.read_idx != last) { kept++; last = w.read_idx; }
            std::fprintf(out, "%s\t%u\t%lld\t%lld\t%lld\t%lld\t%lld\t%lld\t%lld\t%s\t%d\t%d\t%s\t%s\t\n", B.ids[w.read_idx].c_str(), w.read_len,
                         static_cast<long long>(w.rel_dist_to_end), static_cast<long long>(w.read_start_bar), static_cast<long long>(w.read_end_bar),
                         static_cast<long long>(w.read_start_flank), static_cast<long long>(w.read_end_flank), static_cast<long long>(w.bar_start),
                         static_cast<long long>(w.bar_end), kTypeNames[w.match_type & 3], w.flank_cost, w.barcode_cost,
                         bb_groups_label(gs, w.group_idx, w.label_idx), w.strand ? "Rc" : "Fwd");
        }
        total_rows += n_rows;
        return BB_OK;
    };

    FastqReader reader(a.input);
    std::string header, seq, rerr;
    int cur = 0;
    slots[cur].clear();
    bool more = true;
    rc = BB_OK;
    while (more && rc == BB_OK) {
        more = reader.next(header, seq, rerr);
        if (!more && !rerr.empty()) { std::printf("Error during processing: %s\n", rerr.c_str()); rc = BB_ERR_IO; break; }
        Batch& B = slots[cur];
        const bool full = more && (B.bytes + seq.size() > B.cap_bytes || B.n_reads + 1 > B.cap_reads);
        if ((full || !more) && B.n_reads > 0) {
            const int dev = static_cast<int>(submitted % n_gpus);
            // at most 2 batches in flight per GPU
            size_t on_dev = 0; for (const auto& f : flight) on_dev += f.dev == dev;
            while (on_dev >= 2 && rc == BB_OK) { const int d0 = flight.front().dev; rc = collect_one(); if (d0 == dev) on_dev--; }
            if (rc != BB_OK) break;
            rc = bb_submit(ctx[dev], B.bases, B.offsets, B.n_reads, submitted);
            if (rc != BB_OK) { std::printf("Error during processing: %s\n", bb_last_error(ctx[dev])); break; }
            flight.push_back({cur, dev});
            submitted++;
            total_reads += B.n_reads;
            // next free slot: one that is not in flight
            while (static_cast<int>(flight.size()) >= n_slots && rc == BB_OK) rc = collect_one();
            std::vector<char> busy(n_slots, 0); for (const auto& f : flight) busy[f.slot] = 1;
            for (int s = 0; s < n_slots; s++) if (!busy[s]) { cur = s; break; }
            slots[cur].clear();
        }
        if (more) {
            Batch& C = slots[cur];
            if (seq.size() > C.cap_bytes) { std::printf("Error during processing: read longer than the batch buffer (raise --batch-mb)\n"); rc = BB_ERR_INVALID; break; }
            const size_t sp = header.find_first_of(" \t");                   // split_fastq_header, io.rs:5-16
            C.ids.emplace_back(sp == std::string::npos ? header : header.substr(0, sp));
            std::memcpy(C.bases + C.bytes, seq.data(), seq.size());
            C.bytes += seq.size();
            C.offsets[++C.n_reads] = C.bytes;
        }
    }
    while (!flight.empty() && rc == BB_OK) rc = collect_one();
    std::fclose(out);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc == BB_OK) {
        std::printf("Total: %llu  Kept: %llu  Dropped: %llu  (rows: %llu, %.2f s, %.0f reads/s)\n", static_cast<unsigned long long>(total_reads),
                    static_cast<unsigned long long>(kept), static_cast<unsigned long long>(total_reads - kept),
                    static_cast<unsigned long long>(total_rows), secs, secs > 0 ? total_reads / secs : 0.0);
    }
    for (auto& b : slots) b.release();
    for (auto* c : ctx) bb_destroy(c);
    bb_groups_free(gs);
    return rc;
}

}  // namespace

int main(int argc, char** argv) {
    const Args a = parse(argc, argv);
    if (a.cmd == "annotate") {
        std::printf("Starting annotation...\n");
        if (run_annotate(a, a.output) == BB_OK) std::printf("Annotation complete!\n");
        return 0;                                                   // the reference exits 0 even on errors (bin/main.rs:301-304)
    }
    // kit: annotate -> <out>/annotation.tsv (use_kit.rs:43-48); the later stages are not part of this build
    if (a.kit.empty()) usage("kit needs -k <KIT>");
    ::mkdir(a.output.c_str(), 0755);
    std::printf("Running annotate...\n");
    if (run_annotate(a, a.output + "/annotation.tsv") == BB_OK)
        std::printf("Annotation complete: %s/annotation.tsv (inspect / filter / trim are not part of the B200 build)\n", a.output.c_str());
    return 0;
}
