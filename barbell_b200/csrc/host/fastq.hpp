// FASTQ(.gz) reader over several files (reference io.rs:27-32: one paraseq Collection over all paths).  Records are
// parsed in place from a large read buffer (memchr per line) and handed out as views; the annotate driver copies the
// bases straight into the page-locked batch buffer, so every base is copied exactly once on the host.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace bb {

class FastqReader {
  public:
    explicit FastqReader(std::vector<std::string> paths) : paths_(std::move(paths)), buf_(kBuf) {}
    ~FastqReader() { if (gz_) gzclose(gz_); }
    FastqReader(const FastqReader&) = delete;
    FastqReader& operator=(const FastqReader&) = delete;
    // id = header up to the first whitespace, desc = the rest without leading whitespace (split_fastq_header, io.rs:5-16)
    struct View { const char* id; size_t id_len; const char* desc; size_t desc_len; const char* seq; size_t seq_len; const char* qual; };
    // next record; false at the end of the last file or on error (err non-empty)
    bool next(View& v, std::string& err) {
        for (;;) {
            if (!gz_) {
                if (file_ >= paths_.size()) return false;
                gz_ = gzopen(paths_[file_].c_str(), "rb");
                if (!gz_) { err = "Failed to open FASTQ input: " + paths_[file_]; return false; }
                gzbuffer(gz_, 1 << 20);
                pos_ = len_ = 0; eof_ = false;
            }
            size_t p = pos_;
            const char *l0, *l1, *l2, *l3; size_t n0, n1, n2, n3;
            if (line(p, l0, n0) && n0 == 0) { pos_ = p; continue; }   // a blank line between records: skip it alone (like the chunk parser)
            p = pos_;
            if (line(p, l0, n0) && line(p, l1, n1) && line(p, l2, n2) && line(p, l3, n3)) {
                pos_ = p;
                if (l0[0] != '@' || n2 == 0 || l2[0] != '+') { err = "malformed FASTQ record in " + paths_[file_]; return false; }
                if (n3 != n1) { err = "truncated FASTQ record (quality length differs from sequence length) in " + paths_[file_]; return false; }
                auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f' || c == '\r'; };
                size_t idl = 0;
                while (idl < n0 - 1 && !ws(l0[1 + idl])) idl++;
                size_t d0 = 1 + idl;
                while (d0 < n0 && ws(l0[d0])) d0++;
                v.id = l0 + 1; v.id_len = idl; v.desc = l0 + d0; v.desc_len = n0 - d0; v.seq = l1; v.seq_len = n1; v.qual = l3;
                return true;
            }
            // incomplete record in the buffer: compact and refill
            if (eof_) {
                bool only_ws = true;
                for (size_t i = pos_; i < len_; i++) if (buf_[i] != '\n' && buf_[i] != '\r') { only_ws = false; break; }
                if (!only_ws) {
                    // last record without a trailing newline: terminate it and parse once more
                    if (len_ < buf_.size() && !patched_) { buf_[len_++] = '\n'; patched_ = true; continue; }
                    err = "truncated FASTQ record in " + paths_[file_]; return false;
                }
                gzclose(gz_); gz_ = nullptr; file_++; patched_ = false;
                continue;
            }
            if (pos_ > 0) { std::memmove(buf_.data(), buf_.data() + pos_, len_ - pos_); len_ -= pos_; pos_ = 0; }
            if (len_ + 1 >= buf_.size()) buf_.resize(buf_.size() * 2);       // a single record longer than the buffer
            const int n = gzread(gz_, buf_.data() + len_, static_cast<unsigned>(std::min<size_t>(buf_.size() - 1 - len_, 1u << 30)));
            if (n < 0) { err = "read error in " + paths_[file_]; return false; }
            if (n == 0) eof_ = true;
            len_ += static_cast<size_t>(n);
        }
    }

  private:
    static constexpr size_t kBuf = 8u << 20;
    // one line starting at p (without the terminator); false if the buffer holds no complete line
    bool line(size_t& p, const char*& s, size_t& n) {
        if (p >= len_) return false;
        const char* e = static_cast<const char*>(std::memchr(buf_.data() + p, '\n', len_ - p));
        if (!e) return false;
        s = buf_.data() + p; n = static_cast<size_t>(e - s);
        p += n + 1;
        if (n && s[n - 1] == '\r') n--;
        return true;
    }
    std::vector<std::string> paths_;
    size_t file_ = 0;
    gzFile gz_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
    bool eof_ = false, patched_ = false;
};

// ---- helpers for readers that map a plain FASTQ file and parse it in chunks on several threads (the CLI's ParallelSource, bb_trim) ----
// one line [p, e) without its terminator; next = start of the following line
inline void fastq_line_at(const char* m, size_t size, size_t p, size_t& e, size_t& next) {
    const char* nl = static_cast<const char*>(std::memchr(m + p, '\n', size - p));
    e = nl ? static_cast<size_t>(nl - m) : size;
    next = nl ? e + 1 : size;
    if (e > p && m[e - 1] == '\r') e--;
}
// first record start at or after pos: a line "@..." whose third line starts with '+' and whose fourth line is as long as
// its second (a quality line may start with '@' too, but then the line two below is a sequence, never a '+' line)
inline size_t fastq_record_start(const char* m, size_t size, size_t pos) {
    if (pos == 0) return 0;
    if (pos >= size) return size;
    size_t p = pos;
    if (m[pos - 1] != '\n') {
        const char* nl = static_cast<const char*>(std::memchr(m + pos, '\n', size - pos));
        if (!nl) return size;
        p = static_cast<size_t>(nl - m) + 1;
    }
    while (p < size) {
        size_t e0, n0, e1, n1, e2, n2, e3, n3;
        fastq_line_at(m, size, p, e0, n0);
        if (m[p] == '@' && n0 < size) {
            fastq_line_at(m, size, n0, e1, n1);
            if (n1 < size) {
                fastq_line_at(m, size, n1, e2, n2);
                if (e2 > n1 && m[n1] == '+' && n2 <= size) {
                    if (n2 < size) fastq_line_at(m, size, n2, e3, n3); else { e3 = n2; n3 = n2; }
                    if (e3 - n2 == e1 - n0) return p;
                }
            }
        }
        p = n0;
    }
    return size;
}
// One record at p (p < end of the mapped file): views into the map; the quality line is stepped over by its length when it has
// the sequence's length (half of the file is then never read).  Returns 0 = ok (p advanced), 1 = blank line skipped, -1 = error.
// seq_line(head, head_len, start, e, next) scans the sequence line that starts at `start` (e = its end without the terminator,
// next = start of the following line; false + what = error): the plain form is a memchr; a reader that packs the bases while
// it looks for the end of the line (the CLI's) passes its own and touches every base once.
struct FastqRec { const char* head; size_t head_len; const char* seq; size_t seq_len; const char* qual; };
template <class SeqLine>
inline int fastq_record_at(const char* m, size_t size, size_t& p, FastqRec& r, const char*& what, SeqLine&& seq_line) {
    size_t e0, n0, e1, n1, e2, n2, e3, n3;
    fastq_line_at(m, size, p, e0, n0);
    if (e0 == p) { p = n0; return 1; }
    if (n0 >= size) { what = "truncated FASTQ record"; return -1; }
    if (m[p] != '@') { what = "malformed FASTQ record"; return -1; }
    if (!seq_line(m + p + 1, e0 - p - 1, n0, e1, n1)) return -1;
    if (n1 >= size) { what = "truncated FASTQ record"; return -1; }
    fastq_line_at(m, size, n1, e2, n2);
    const size_t seq_len = e1 - n0, q_end = n2 + seq_len;
    if (q_end == size) { e3 = q_end; n3 = q_end; }
    else if (q_end < size && m[q_end] == '\n') { e3 = q_end; n3 = q_end + 1; }
    else if (q_end + 1 < size && m[q_end] == '\r' && m[q_end + 1] == '\n') { e3 = q_end; n3 = q_end + 2; }
    else if (n2 < size) fastq_line_at(m, size, n2, e3, n3); else { e3 = n2; n3 = n2; }
    if (e2 == n1 || m[n1] != '+') { what = "malformed FASTQ record"; return -1; }
    if (e3 - n2 != seq_len) { what = "truncated FASTQ record (quality length differs from sequence length)"; return -1; }
    r.head = m + p + 1; r.head_len = e0 - p - 1; r.seq = m + n0; r.seq_len = seq_len; r.qual = m + n2;
    p = n3;
    return 0;
}
inline int fastq_record_at(const char* m, size_t size, size_t& p, FastqRec& r, const char*& what) {
    return fastq_record_at(m, size, p, r, what, [&](const char*, size_t, size_t start, size_t& e, size_t& next) { fastq_line_at(m, size, start, e, next); return true; });
}
// id = header up to the first whitespace, desc = the rest without leading whitespace (split_fastq_header, io.rs:5-16)
inline void fastq_split_header(const char* head, size_t n, size_t& id_len, size_t& desc_off) {
    auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f' || c == '\r'; };
    id_len = 0;
    while (id_len < n && !ws(head[id_len])) id_len++;
    desc_off = id_len;
    while (desc_off < n && ws(head[desc_off])) desc_off++;
}

}  // namespace bb
