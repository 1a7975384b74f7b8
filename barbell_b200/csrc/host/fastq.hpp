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

}  // namespace bb
