// The stages that consume annotation.tsv: `filter`, `inspect` and `trim`, host-side C++ behind the
// C ABI so that `barbell kit` runs the reference's whole pipeline (src/kits/use_kit.rs:11-109).  They are sequential
// text/IO stages in the reference as well (single-threaded csv + serde); nothing here touches the GPU.
//
//   pattern DSL + matching   src/filter/pattern.rs:8-383      parse_pattern(), match_pattern()
//   filter                   src/filter/filter.rs:10-214       bb_filter()
//   inspect                  src/inspect/inspect.rs:9-208      bb_inspect()
//   trim                     src/trim/trim.rs:31-480           bb_trim()
//   kit pattern sets         src/kits/kits.rs:175-236          bb_kit_filter_patterns()
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/barbell_b200.h"
#include "fastq.hpp"

namespace {

enum { FTAG = 0, RTAG = 1, FFLANK = 2, RFLANK = 3 };
const char* const kTypeNames[] = {"Ftag", "Rtag", "Fflank", "Rflank"};

struct Cut { uint64_t group_id; bool after; };   // pattern.rs:8-19: After = cut at match end (">>"), Before = at match start ("<<")

// One annotation.tsv row (BarbellMatch, searcher.rs:31-64)
struct Row {
    std::string read_id, label;
    uint64_t read_len = 0, read_start_bar = 0, read_end_bar = 0, read_start_flank = 0, read_end_flank = 0, bar_start = 0, bar_end = 0;
    int64_t rel_dist_to_end = 0;
    int match_type = FTAG, strand = 0;
    int32_t flank_cost = 0, barcode_cost = 0;
    std::vector<std::pair<Cut, uint64_t>> cuts;    // (cut, index of the annotation inside its read)
};

const char* const kHeader[15] = {"read_id", "read_len", "rel_dist_to_end", "read_start_bar", "read_end_bar", "read_start_flank", "read_end_flank",
                                 "bar_start", "bar_end", "match_type", "flank_cost", "barcode_cost", "label", "strand", "cuts"};

void set_err(char* err, size_t errlen, const std::string& msg) { if (err && errlen) std::snprintf(err, errlen, "%s", msg.c_str()); }

std::string trim_ws(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) a++;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) b--;
    return s.substr(a, b - a);
}

bool parse_u64(const std::string& s, uint64_t& v) {
    if (s.empty()) return false;
    size_t i = s[0] == '+' ? 1 : 0;
    if (i >= s.size()) return false;
    v = 0;
    for (; i < s.size(); i++) { if (s[i] < '0' || s[i] > '9') return false; v = v * 10 + static_cast<uint64_t>(s[i] - '0'); }
    return true;
}
bool parse_i64(const std::string& s, int64_t& v) {
    if (s.empty()) return false;
    const bool neg = s[0] == '-';
    uint64_t u;
    if (!parse_u64(neg ? s.substr(1) : s, u)) return false;
    v = neg ? -static_cast<int64_t>(u) : static_cast<int64_t>(u);
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// TSV reader / writer (csv crate, tab delimiter, header row; fields of this format never need quoting)
// ---------------------------------------------------------------------------------------------------------------
// The file is mapped and every line is parsed in place (field = pointer + length; no per-field strings): annotation.tsv of a large
// run holds tens of millions of rows, and filter, inspect and trim each read all of it.
class TsvReader {
  public:
    bool open(const std::string& path, std::string& err) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd_ < 0 || fstat(fd_, &st) != 0) { err = path + ": " + std::strerror(errno); return false; }
        size_ = static_cast<size_t>(st.st_size);
        if (size_ == 0) { empty_ = true; return true; }               // 0-byte file: a run without hits (annotator.rs:20-24)
        void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) {                                       // not mappable (a pipe): read it all
            owned_.clear();
            char buf[1 << 16];
            ssize_t n;
            while ((n = ::read(fd_, buf, sizeof buf)) > 0) owned_.append(buf, static_cast<size_t>(n));
            data_ = owned_.data(); size_ = owned_.size();
            if (size_ == 0) { empty_ = true; return true; }
        } else {
            data_ = static_cast<const char*>(m); mapped_ = true;
            madvise(m, size_, MADV_SEQUENTIAL);
        }
        Field f[kMaxCols]; int nf;
        if (!next_line(f, nf)) { empty_ = true; return true; }
        for (int c = 0; c < 15; c++) {
            col_[c] = -1;
            const size_t hl = std::strlen(kHeader[c]);
            for (int i = 0; i < nf; i++) if (f[i].n == hl && std::memcmp(f[i].p, kHeader[c], hl) == 0) col_[c] = i;
            if (col_[c] < 0) { err = "CSV deserialize error: missing field `" + std::string(kHeader[c]) + "`"; return false; }
        }
        return true;
    }
    ~TsvReader() { if (mapped_) munmap(const_cast<char*>(data_), size_); if (fd_ >= 0) ::close(fd_); }
    // 1 = row, 0 = end, -1 = error
    int next(Row& r, std::string& err) {
        if (empty_) return 0;
        Field f[kMaxCols]; int nf;
        for (;;) {
            if (!next_line(f, nf)) return 0;
            if (!(nf == 1 && f[0].n == 0)) break;                    // blank line
        }
        lineno_++;
        static const Field none{"", 0};
        auto get = [&](int c) -> const Field& { return col_[c] < nf ? f[col_[c]] : none; };
        auto bad = [&](int c) { err = "CSV deserialize error: record " + std::to_string(lineno_) + ": field `" + kHeader[c] + "`: invalid value `" + std::string(get(c).p, get(c).n) + "`"; return -1; };
        auto u64 = [](const Field& x, uint64_t& v) {
            if (x.n == 0) return false;
            size_t i = x.p[0] == '+' ? 1 : 0;
            if (i >= x.n) return false;
            v = 0;
            for (; i < x.n; i++) { const unsigned d = static_cast<unsigned char>(x.p[i]) - '0'; if (d > 9) return false; v = v * 10 + d; }
            return true;
        };
        auto i64 = [&](const Field& x, int64_t& v) {
            if (x.n == 0) return false;
            const bool neg = x.p[0] == '-';
            uint64_t u;
            if (!u64(neg ? Field{x.p + 1, x.n - 1} : x, u)) return false;
            v = neg ? -static_cast<int64_t>(u) : static_cast<int64_t>(u);
            return true;
        };
        auto is = [](const Field& x, const char* lit) { const size_t n = std::strlen(lit); return x.n == n && std::memcmp(x.p, lit, n) == 0; };
        r.read_id.assign(get(0).p, get(0).n);
        r.cuts.clear();
        if (!u64(get(1), r.read_len)) return bad(1);
        if (!i64(get(2), r.rel_dist_to_end)) return bad(2);
        uint64_t* u[] = {&r.read_start_bar, &r.read_end_bar, &r.read_start_flank, &r.read_end_flank, &r.bar_start, &r.bar_end};
        for (int c = 3; c <= 8; c++) if (!u64(get(c), *u[c - 3])) return bad(c);
        r.match_type = -1;
        for (int t = 0; t < 4; t++) if (is(get(9), kTypeNames[t])) r.match_type = t;
        if (r.match_type < 0) return bad(9);
        int64_t v;
        if (!i64(get(10), v)) return bad(10);
        r.flank_cost = static_cast<int32_t>(v);
        if (!i64(get(11), v)) return bad(11);
        r.barcode_cost = static_cast<int32_t>(v);
        r.label.assign(get(12).p, get(12).n);
        if (is(get(13), "Fwd")) r.strand = 0; else if (is(get(13), "Rc")) r.strand = 1; else { err = "Invalid strand: " + std::string(get(13).p, get(13).n); return -1; }
        // cuts: "After(0):1,Before(0):2" (searcher.rs:105-142)
        const std::string cs(get(14).p, get(14).n);
        size_t p = 0;
        while (p < cs.size()) {
            size_t q = cs.find(',', p);
            if (q == std::string::npos) q = cs.size();
            const std::string part = cs.substr(p, q - p);
            const size_t colon = part.find(':');
            if (colon == std::string::npos) { err = "Invalid cut format: missing position part"; return -1; }
            std::string cut_str = trim_ws(part.substr(0, colon)), pos_str = part.substr(colon + 1);
            const size_t colon2 = pos_str.find(':');
            if (colon2 != std::string::npos) pos_str = pos_str.substr(0, colon2);
            Cut cut{0, false};
            bool ok = false;
            if (cut_str.rfind("Before(", 0) == 0 && cut_str.back() == ')') { cut.after = false; ok = parse_u64(cut_str.substr(7, cut_str.size() - 8), cut.group_id); }
            else if (cut_str.rfind("After(", 0) == 0 && cut_str.back() == ')') { cut.after = true; ok = parse_u64(cut_str.substr(6, cut_str.size() - 7), cut.group_id); }
            if (!ok) { err = "Invalid cut string: " + cut_str; return -1; }
            uint64_t pos;
            if (!parse_u64(pos_str, pos)) { err = "Invalid position: " + pos_str; return -1; }
            r.cuts.push_back({cut, pos});
            p = q + 1;
        }
        return 1;
    }

  private:
    struct Field { const char* p; size_t n; };
    static constexpr int kMaxCols = 64;
    // the fields of the next line (quotes around a whole field are dropped, like the csv crate does); false at the end of the file
    bool next_line(Field* f, int& nf) {
        if (pos_ >= size_) return false;
        const char* s = data_ + pos_;
        const char* nl = static_cast<const char*>(std::memchr(s, '\n', size_ - pos_));
        const char* e = nl ? nl : data_ + size_;
        pos_ = static_cast<size_t>(e - data_) + (nl ? 1 : 0);
        while (e > s && (e[-1] == '\r' || e[-1] == '\n')) e--;
        nf = 0;
        const char* p = s;
        for (;;) {
            const char* t = static_cast<const char*>(std::memchr(p, '\t', static_cast<size_t>(e - p)));
            const char* fe = t ? t : e;
            Field x{p, static_cast<size_t>(fe - p)};
            if (x.n >= 2 && x.p[0] == '"' && x.p[x.n - 1] == '"') { x.p++; x.n -= 2; }
            if (nf < kMaxCols) f[nf++] = x;
            if (!t) break;
            p = t + 1;
        }
        return true;
    }
    int fd_ = -1;
    const char* data_ = nullptr;
    size_t size_ = 0, pos_ = 0;
    bool mapped_ = false;
    std::string owned_;
    int col_[15];
    bool empty_ = false;
    uint64_t lineno_ = 0;
};

class TsvWriter {                                        // rows are formatted by hand into a large buffer
  public:
    bool open(const std::string& path, std::string& err) {
        f_ = std::fopen(path.c_str(), "w");
        if (!f_) { err = path + ": " + std::strerror(errno); return false; }
        buf_.reserve(4u << 20);
        return true;
    }
    ~TsvWriter() { close(); }
    void write(const Row& r) {
        if (!header_) {               // the csv writer emits the header with the first serialised row
            for (int c = 0; c < 15; c++) { buf_ += kHeader[c]; buf_ += c == 14 ? '\n' : '\t'; }
            header_ = true;
        }
        buf_ += r.read_id; tab(); u(r.read_len); tab(); i(r.rel_dist_to_end); tab();
        u(r.read_start_bar); tab(); u(r.read_end_bar); tab(); u(r.read_start_flank); tab(); u(r.read_end_flank); tab(); u(r.bar_start); tab(); u(r.bar_end); tab();
        buf_ += kTypeNames[r.match_type]; tab(); i(r.flank_cost); tab(); i(r.barcode_cost); tab(); buf_ += r.label; tab(); buf_ += r.strand ? "Rc" : "Fwd"; tab();
        for (size_t k = 0; k < r.cuts.size(); k++) {
            if (k) buf_ += ',';
            buf_ += r.cuts[k].first.after ? "After(" : "Before("; u(r.cuts[k].first.group_id); buf_ += "):"; u(r.cuts[k].second);
        }
        buf_ += '\n';
        if (buf_.size() > (3u << 20)) flush();
    }
    bool close() {
        if (!f_) return ok_;
        flush();
        if (std::fclose(f_) != 0) ok_ = false;
        f_ = nullptr;
        return ok_;
    }

  private:
    void tab() { buf_ += '\t'; }
    void u(uint64_t v) { char t[24]; int k = 0; do { t[k++] = static_cast<char>('0' + v % 10); v /= 10; } while (v); while (k) buf_ += t[--k]; }
    void i(int64_t v) { if (v < 0) { buf_ += '-'; u(0ull - static_cast<uint64_t>(v)); } else u(static_cast<uint64_t>(v)); }
    void flush() { if (!buf_.empty() && std::fwrite(buf_.data(), 1, buf_.size(), f_) != buf_.size()) ok_ = false; buf_.clear(); }
    FILE* f_ = nullptr;
    std::string buf_;
    bool header_ = false, ok_ = true;
};

// ---------------------------------------------------------------------------------------------------------------
// pattern DSL (pattern.rs:242-383) and matching (pattern.rs:100-240)
// ---------------------------------------------------------------------------------------------------------------
enum RelPos { REL_NONE = 0, REL_LEFT, REL_RIGHT, REL_PREV_LEFT };

struct PatternElement {
    int match_type = FTAG;
    int orientation = -1;                 // -1 any, 0 Fwd, 1 Rc
    bool has_label = false; std::string label;
    bool has_placeholder = false; uint64_t placeholder = 0;
    int64_t range0 = 0, range1 = 0;
    RelPos relative_to = REL_NONE;
    std::vector<Cut> cuts;
};
struct Pattern { std::vector<PatternElement> elements; };

bool parse_range(std::string s, int64_t& a, int64_t& b) {
    size_t i = 0, j = s.size();
    while (i < j && (s[i] == '(' || s[i] == ')')) i++;
    while (j > i && (s[j - 1] == '(' || s[j - 1] == ')')) j--;
    s = s.substr(i, j - i);
    std::vector<std::string> parts;
    size_t p = 0;
    for (;;) {
        const size_t q = s.find("..", p);
        parts.push_back(s.substr(p, q == std::string::npos ? std::string::npos : q - p));
        if (q == std::string::npos) break;
        p = q + 2;
    }
    if (parts.size() != 2) return false;
    return parse_i64(trim_ws(parts[0]), a) && parse_i64(trim_ws(parts[1]), b);
}

bool parse_position(const std::string& pos_str, RelPos& rel, int64_t& a, int64_t& b) {
    if (std::count(pos_str.begin(), pos_str.end(), '(') != 1) return false;
    const size_t paren = pos_str.find('(');
    std::string name = pos_str.substr(0, paren);
    size_t i = 0;
    while (i < name.size() && name[i] == '@') i++;
    name = name.substr(i);
    if (name == "left") rel = REL_LEFT; else if (name == "right") rel = REL_RIGHT; else if (name == "prev_left") rel = REL_PREV_LEFT; else return false;
    return parse_range(trim_ws(pos_str.substr(paren)), a, b);
}

// 1 = parsed, 0 = not an element (dropped, like the reference's filter_map), -1 = hard error (the reference panics)
int parse_element(const std::string& element_str, PatternElement& el, std::string& err) {
    const size_t br = element_str.find('[');
    if (br == std::string::npos) return 0;
    const std::string type = trim_ws(element_str.substr(0, br));
    if (type == "Ftag") el.match_type = FTAG; else if (type == "Rtag") el.match_type = RTAG;
    else if (type == "Fflank") el.match_type = FFLANK; else if (type == "Rflank") el.match_type = RFLANK;
    else if (type == "Flank" || type == "flank") { err = "Flank is not valid, use Fflank or Rflank"; return -1; }
    else return 0;
    std::string params = element_str.substr(br + 1);
    while (!params.empty() && params.back() == ']') params.pop_back();
    size_t p = 0;
    for (;;) {
        const size_t q = params.find(',', p);
        const std::string param = trim_ws(params.substr(p, q == std::string::npos ? std::string::npos : q - p));
        if (param == "fw") el.orientation = 0;
        else if (param == "rc") el.orientation = 1;
        else if (!param.empty() && param[0] == '@') {
            RelPos rel; int64_t a, b;
            if (parse_position(param, rel, a, b)) { el.relative_to = rel; el.range0 = a; el.range1 = b; }
        } else if (!param.empty() && param[0] == '?') {
            uint64_t n;
            if (parse_u64(param.substr(1), n)) { el.has_placeholder = true; el.placeholder = n; }
        } else if (!param.empty() && (param[0] == '>' || param[0] == '<')) {
            if (param.size() < 2) { err = "cut marker `" + param + "` is too short (use >> or <<)"; return -1; }   // the reference panics on the slice
            const std::string pre = param.substr(0, 2);
            uint64_t id = 0;
            const bool id_ok = param.size() == 2 || parse_u64(param.substr(2), id);
            if (id_ok && pre == ">>") el.cuts.push_back({id, true});
            else if (id_ok && pre == "<<") el.cuts.push_back({id, false});
        } else if (param == "*") {
        } else {
            std::string l = param;
            while (!l.empty() && l.front() == '"') l.erase(l.begin());
            while (!l.empty() && l.back() == '"') l.pop_back();
            el.has_label = true; el.label = l;
        }
        if (q == std::string::npos) break;
        p = q + 1;
    }
    return 1;
}

bool parse_pattern(const std::string& pattern_str, Pattern& pat, std::string& err) {
    pat.elements.clear();
    size_t p = 0, user_elems = 0;
    for (;;) {
        const size_t q = pattern_str.find("__", p);
        PatternElement el;
        const int rc = parse_element(trim_ws(pattern_str.substr(p, q == std::string::npos ? std::string::npos : q - p)), el, err);
        if (rc < 0) return false;
        if (rc > 0) pat.elements.push_back(el);
        user_elems++;
        if (q == std::string::npos) break;
        p = q + 2;
    }
    if (user_elems != pat.elements.size()) { err = "Pattern parse error for: \"" + pattern_str + "\""; return false; }   // basic_verify
    return true;
}

std::string describe_pattern(const Pattern& pat) {         // canonical text form (tests compare it with the reference's structs)
    std::string s;
    for (size_t i = 0; i < pat.elements.size(); i++) {
        const PatternElement& e = pat.elements[i];
        if (i) s += "__";
        s += kTypeNames[e.match_type];
        s += "[ori=" + std::string(e.orientation < 0 ? "any" : e.orientation ? "rc" : "fw");
        s += ",label=" + (e.has_label ? e.label : std::string("*"));
        s += ",ph=" + (e.has_placeholder ? std::to_string(e.placeholder) : std::string("-"));
        s += ",rel=" + std::string(e.relative_to == REL_LEFT ? "left" : e.relative_to == REL_RIGHT ? "right" : e.relative_to == REL_PREV_LEFT ? "prev_left" : "none");
        s += ",range=" + std::to_string(e.range0) + ".." + std::to_string(e.range1);
        s += ",cuts=";
        for (size_t c = 0; c < e.cuts.size(); c++) s += (c ? "|" : "") + std::string(e.cuts[c].after ? "After(" : "Before(") + std::to_string(e.cuts[c].group_id) + ")";
        s += "]";
    }
    return s;
}

bool matches_element(const Row& m, const PatternElement& el, bool has_prev, int64_t prev_end, std::map<uint64_t, std::string>& labels) {
    // check_match_type_and_label
    if (m.match_type != el.match_type) return false;
    if ((el.match_type == FTAG || el.match_type == RTAG) && el.has_label) {
        if (!el.label.empty() && el.label[0] == '~') { if (m.label.find(el.label.substr(1)) == std::string::npos) return false; }
        else if (el.label != m.label) return false;
    }
    // check_placeholder
    if (el.has_placeholder) {
        auto it = labels.find(el.placeholder);
        if (it != labels.end()) { if (m.label != it->second) return false; }
        else labels[el.placeholder] = m.label;
    }
    // check_orientation
    if (el.orientation >= 0 && el.orientation != m.strand) return false;
    // check_relative_position
    const int64_t m_start = static_cast<int64_t>(m.read_start_bar), m_end = static_cast<int64_t>(m.read_end_bar), seq_len = static_cast<int64_t>(m.read_len);
    switch (el.relative_to) {
        case REL_LEFT: if (m_start < el.range0 || m_start > el.range1) return false; break;
        case REL_RIGHT: if (m_end < seq_len - el.range1 || m_end > seq_len - el.range0) return false; break;
        case REL_PREV_LEFT: if (has_prev && (m_start < prev_end + el.range0 || m_start > prev_end + el.range1)) return false; break;
        case REL_NONE: break;
    }
    return true;
}

// pattern.rs:205-240: element q must match annotation q; returns the (annotation index, cut) list of the pattern
bool match_pattern(const std::vector<Row>& matches, const Pattern& pat, std::vector<std::pair<size_t, Cut>>& cut_positions) {
    cut_positions.clear();
    if (matches.size() < pat.elements.size()) return false;
    bool has_prev = false; int64_t prev_end = 0;
    std::map<uint64_t, std::string> labels;
    size_t idx = 0;
    for (const PatternElement& el : pat.elements) {
        if (idx >= matches.size()) { cut_positions.clear(); return false; }
        const Row& m = matches[idx];
        if (!matches_element(m, el, has_prev, prev_end, labels)) { cut_positions.clear(); return false; }
        for (const Cut& c : el.cuts) cut_positions.push_back({idx, c});
        has_prev = true; prev_end = static_cast<int64_t>(m.read_end_bar);
        idx++;
    }
    return true;
}

// filter.rs:183-214: the longest matching pattern supplies the cuts; the read passes when that pattern covers every annotation
bool check_filter_pass(std::vector<Row>& annotations, const std::vector<Pattern>& patterns) {
    size_t max_matches = 0;
    std::vector<std::pair<size_t, Cut>> best, cur;
    for (const Pattern& p : patterns)
        if (match_pattern(annotations, p, cur) && p.elements.size() > max_matches) { max_matches = p.elements.size(); best = cur; }
    if (max_matches > 0)
        for (const auto& pc : best) annotations[pc.first].cuts.push_back({pc.second, pc.first});
    return max_matches == annotations.size();
}

// kits.rs:175-236
const char* const kSingleSafe[] = {
    "Ftag[fw, *, @left(0..250), >>]",
    "Ftag[fw, ?1, @left(0..250)]__Ftag[fw, ?1, @prev_left(0..250), >>]",
};
const char* const kSingleMax[] = {
    "Ftag[fw, *, @left(0..250), >>]",
    "Ftag[fw, ?1, @left(0..250)]__Ftag[fw, ?1, @prev_left(0..250), >>]",
    "Ftag[fw, *, @left(0..250)]__Ftag[fw, *, @prev_left(0..250), >>]",
    "Ftag[fw, *, @left(0..250), >>]__Ftag[<<, fw, *, @right(0..250)]",
    "Ftag[fw, *, @left(0..250)]__Ftag[fw, *, @prev_left(0..250), >>]__Ftag[<<, fw, *, @right(0..250)]",
};
const char* const kDoubleSafe[] = {
    "Ftag[fw, *, @left(0..250), >>]",
    "Ftag[<<, rc, *, @right(0..250)]",
    "Ftag[fw, ?1, @left(0..250), >>]__Ftag[<<, rc, ?1, @right(0..250)]",
};
const char* const kDoubleMax[] = {
    "Ftag[fw, *, @left(0..250), >>]",
    "Ftag[<<, rc, *, @right(0..250)]",
    "Ftag[fw, ?1, @left(0..250), >>]__Ftag[<<, rc, ?1, @right(0..250)]",
    "Ftag[fw, *, @left(0..250)]__Ftag[fw, ?1, @prev_left(0..250), >>]__Ftag[<<, rc, ?1, @right(0..250)]",
    "Ftag[fw, *, @left(0..250), >>]__Fflank[<<, rc, *, @right(0..250)]",
    "Fflank[fw, *, @left(0..250), >>]__Ftag[<<, rc, *, @right(0..250)]",
    "Ftag[fw, *, @left(0..250)]__Ftag[fw, *, @prev_left(0..250), >>]",
    "Ftag[fw, ?1, @left(0..250), >>]__Ftag[<<, fw, ?1, @right(0..250)]__Ftag[rc, *, @right(0..250)]",
    "Ftag[fw, *, @left(0..250)]__Ftag[rc, *, @prev_left(0..250)]__Ftag[fw, ?1, @prev_left(0..250), >>]__Ftag[<<, rc, ?1, @right(0..250)]",
};

// groups of consecutive rows with the same read_id (filter.rs:50-82, inspect.rs:159-178)
template <class F>
int for_each_read(const char* path, F&& fn, std::string& err) {
    TsvReader rd;
    if (!rd.open(path, err)) return BB_ERR_IO;
    std::vector<Row> group;
    Row r;
    for (;;) {
        const int rc = rd.next(r, err);
        if (rc < 0) return BB_ERR_INVALID;
        if (rc == 0) break;
        if (!group.empty() && group.front().read_id != r.read_id) { fn(group); group.clear(); }
        group.push_back(std::move(r));
    }
    if (!group.empty()) fn(group);
    return BB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// inspect (inspect.rs:9-117)
// ---------------------------------------------------------------------------------------------------------------
uint64_t bucket_position(uint64_t pos, uint64_t bucket) { return ((pos ? pos - 1 : 0) / bucket) * bucket; }
uint64_t sat_sub(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }

std::string group_structure(const std::vector<Row>& group, uint64_t bucket) {
    std::string out;
    bool has_prev = false; uint64_t prev_end = 0;
    for (const Row& a : group) {
        const uint64_t start = a.read_start_bar, end = a.read_end_bar;
        char tag[96];
        auto right_tag = [&] {
            std::snprintf(tag, sizeof tag, "@right(%llu..%llu)", static_cast<unsigned long long>(bucket_position(sat_sub(a.read_len, end), bucket)),
                          static_cast<unsigned long long>(bucket_position(sat_sub(a.read_len, start), bucket) + bucket));
        };
        if (has_prev) {
            const uint64_t d_prev = sat_sub(start, prev_end), d_right = sat_sub(a.read_len, end);
            if (d_prev <= d_right) {
                const uint64_t g0 = bucket_position(d_prev, bucket);
                std::snprintf(tag, sizeof tag, "@prev_left(%llu..%llu)", static_cast<unsigned long long>(g0), static_cast<unsigned long long>(g0 + bucket));
            } else right_tag();
        } else if (a.rel_dist_to_end > 0) {
            const uint64_t s0 = bucket_position(start, bucket);
            std::snprintf(tag, sizeof tag, "@left(%llu..%llu)", static_cast<unsigned long long>(s0), static_cast<unsigned long long>(s0 + bucket));
        } else right_tag();
        const char* cut = a.cuts.empty() ? "" : (a.strand == 0 ? ", <<" : ", >>");
        if (!out.empty()) out += "__";
        out += std::string(kTypeNames[a.match_type]) + "[" + (a.strand == 0 ? "fw" : "rc") + ", *" + cut + ", " + tag + "]";
        has_prev = true; prev_end = end;
    }
    return out;
}

std::string colorize(std::string s) {                       // inspect.rs:120-131 (truecolor escapes of the `colored` crate)
    const struct { const char* word; int r, g, b; } C[] = {{"Fflank", 255, 182, 193}, {"Ftag", 231, 84, 128}, {"Rflank", 173, 216, 230}, {"Rtag", 0, 0, 139}};
    for (const auto& c : C) {
        char rep[64];
        std::snprintf(rep, sizeof rep, "\x1b[38;2;%d;%d;%dm%s\x1b[0m", c.r, c.g, c.b, c.word);
        size_t p = 0;
        while ((p = s.find(c.word, p)) != std::string::npos) { s.replace(p, std::strlen(c.word), rep); p += std::strlen(rep); }
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// trim (trim.rs:31-315)
// ---------------------------------------------------------------------------------------------------------------
std::string create_label(const bb_trim_opts& o, const std::vector<const Row*>& annos) {       // trim.rs:56-106
    if (!o.add_labels) return "none";
    std::vector<std::string> parts;
    for (const Row* m : annos) {
        if (!o.add_flank && m->label.find("flank") != std::string::npos) continue;
        std::string s = m->label;
        if (o.add_orientation) s += m->strand == 0 ? "_fw" : "_rc";
        parts.push_back(s);
    }
    if (parts.empty()) return "none";
    if (o.sort_labels) std::sort(parts.begin(), parts.end());
    else if (o.only_side == 1) return parts.front();
    else if (o.only_side == 2) return parts.back();
    std::string out;
    for (size_t i = 0; i < parts.size(); i++) out += (i ? "__" : "") + parts[i];
    return out;
}

struct Slice { uint64_t start, end; std::vector<const Row*> annos; };

// trim.rs:127-248.  The reference collects the cut groups in a HashMap and orders them with a stable sort on the first
// member's flank start; groups that tie on that key keep HashMap order there (arbitrary) and group-id order here.
std::vector<Slice> preprocess_cuts(const std::vector<Row>& annotations, uint64_t seq_len) {
    struct Member { uint64_t start, end; const Cut* cut; const Row* anno; };
    std::map<uint64_t, std::vector<Member>> groups;
    for (const Row& a : annotations)
        for (const auto& c : a.cuts) groups[c.first.group_id].push_back({a.read_start_flank, a.read_end_flank, &c.first, &a});
    std::vector<const std::vector<Member>*> sorted;
    for (const auto& kv : groups) sorted.push_back(&kv.second);
    std::stable_sort(sorted.begin(), sorted.end(), [](const std::vector<Member>* a, const std::vector<Member>* b) { return a->front().start < b->front().start; });
    std::vector<Slice> slices;
    for (size_t i = 0; i < sorted.size(); i++) {
        const std::vector<Member>& g = *sorted[i];
        if (g.size() == 2) {
            const uint64_t start = g[0].cut->after ? g[0].end : g[0].start;
            const uint64_t end = g[1].cut->after ? g[1].end : g[1].start;
            slices.push_back({start, end, {g[0].anno, g[1].anno}});
        } else if (g.size() == 1) {
            const Member& m = g[0];
            if (!m.cut->after) {                       // Before: look left for the start
                uint64_t s = 0; const Row* left = nullptr;
                if (i > 0) {
                    const std::vector<Member>& pg = *sorted[i - 1];
                    size_t best = 0;                    // max_by_key keeps the LAST maximum
                    for (size_t q = 0; q < pg.size(); q++) if (pg[q].end >= pg[best].end) best = q;
                    s = pg[best].end; left = pg[best].anno;
                }
                Slice sl{s, m.start, {}};
                if (left) sl.annos.push_back(left);
                sl.annos.push_back(m.anno);
                slices.push_back(sl);
            } else {                                   // After: look right for the end
                uint64_t e = seq_len; const Row* right = nullptr;
                if (i + 1 < sorted.size()) {
                    const std::vector<Member>& ng = *sorted[i + 1];
                    size_t best = 0;                    // min_by_key keeps the FIRST minimum
                    for (size_t q = 1; q < ng.size(); q++) if (ng[q].start < ng[best].start) best = q;
                    e = ng[best].start; right = ng[best].anno;
                }
                Slice sl{m.end, e, {m.anno}};
                if (right) sl.annos.push_back(right);
                slices.push_back(sl);
            }
        }
    }
    return slices;
}

struct RcTable {                                        // trim.rs:486-530
    uint8_t t[256];
    RcTable() {
        for (int i = 0; i < 256; i++) t[i] = static_cast<uint8_t>(i);
        const char *a = "ACTGRYSWKMBDHVNX", *b = "TGACYRSWMKVHDBNX";
        for (int i = 0; a[i]; i++) { t[static_cast<uint8_t>(a[i])] = static_cast<uint8_t>(b[i]); t[static_cast<uint8_t>(a[i] | 0x20)] = static_cast<uint8_t>(b[i] | 0x20); }
    }
};
const RcTable kRc;

struct TrimmedRead { std::string seq, qual, label, suffix; };

// trim.rs:250-300
std::vector<TrimmedRead> process_read_and_anno(const char* seq, const char* qual, uint64_t seq_len, const std::vector<Row>& annotations, const bb_trim_opts& o) {
    std::vector<TrimmedRead> out;
    const std::vector<Slice> slices = preprocess_cuts(annotations, seq_len);
    for (size_t sc = 0; sc < slices.size(); sc++) {
        const Slice& s = slices[sc];
        if (s.start >= s.end) continue;
        TrimmedRead t;
        if (o.skip_trim) { t.seq.assign(seq, seq_len); t.qual.assign(qual, seq_len); }
        else {
            const uint64_t e = std::min<uint64_t>(s.end, seq_len), b = std::min<uint64_t>(s.start, e);   // the reference would panic on an out-of-range slice
            t.seq.assign(seq + b, e - b); t.qual.assign(qual + b, e - b);
        }
        bool flip = false;
        if (o.flip) for (const Row* a : s.annos) if (a->match_type == FTAG && a->strand == 1) flip = true;      // should_flip
        if (flip) {
            std::reverse(t.seq.begin(), t.seq.end());
            for (char& c : t.seq) c = static_cast<char>(kRc.t[static_cast<uint8_t>(c)]);
            std::reverse(t.qual.begin(), t.qual.end());
        }
        t.label = create_label(o, s.annos);
        t.suffix = sc == 0 ? "" : "_" + std::to_string(sc);
        out.push_back(std::move(t));
    }
    return out;
}

class FastqOut {                                         // plain or gzip writer per label (trim.rs:420-445)
  public:
    bool open(const std::string& path, bool gz, std::string& err) {
        gz_mode_ = gz;
        if (gz) { g_ = gzopen(path.c_str(), "wb"); if (!g_) { err = "Failed to create output file '" + path + "': " + std::strerror(errno); return false; } gzbuffer(g_, 1 << 18); }
        else { f_ = std::fopen(path.c_str(), "w"); if (!f_) { err = "Failed to create output file '" + path + "': " + std::strerror(errno) + (errno == EMFILE ? "\nTry setting ulimit higher: \"ulimit -n 65000\"" : ""); return false; } std::setvbuf(f_, nullptr, _IOFBF, 1 << 18); }
        return true;
    }
    // the reference aborts on a failed write (`expect("Failed to write ...")`, trim.rs:420-445): errors are remembered and reported by close()
    void write(const std::string& s) {
        if (s.empty()) return;
        if (gz_mode_) { if (gzwrite(g_, s.data(), static_cast<unsigned>(s.size())) != static_cast<int>(s.size())) bad_ = true; }
        else if (std::fwrite(s.data(), 1, s.size(), f_) != s.size()) bad_ = true;
    }
    bool close() {                                       // false when any write or the close itself failed (disk full, I/O error)
        if (g_) { if (gzclose(g_) != Z_OK) bad_ = true; g_ = nullptr; }
        if (f_) { if (std::fclose(f_) != 0) bad_ = true; f_ = nullptr; }
        return !bad_;
    }
    ~FastqOut() { close(); }
  private:
    bool gz_mode_ = false, bad_ = false; gzFile g_ = nullptr; FILE* f_ = nullptr;
};

}  // namespace

extern "C" {

int bb_kit_filter_patterns(int double_label, int maximize, const char* const** patterns, int32_t* n) {
    if (!patterns || !n) return BB_ERR_INVALID;
    if (double_label) { *patterns = maximize ? kDoubleMax : kDoubleSafe; *n = maximize ? 9 : 3; }
    else { *patterns = maximize ? kSingleMax : kSingleSafe; *n = maximize ? 5 : 2; }
    return BB_OK;
}

int bb_pattern_parse(const char* pattern, char* canonical, size_t canonical_len, char* err, size_t errlen) {
    if (!pattern) return BB_ERR_INVALID;
    Pattern p; std::string e;
    if (!parse_pattern(pattern, p, e)) { set_err(err, errlen, e); return BB_ERR_INVALID; }
    if (canonical && canonical_len) std::snprintf(canonical, canonical_len, "%s", describe_pattern(p).c_str());
    return BB_OK;
}

int bb_filter(const char* annotated, const char* output, const char* dropped, const char* const* patterns, int32_t n_patterns,
              uint64_t counts[3], char* err, size_t errlen) {
    if (!annotated || !output || (n_patterns > 0 && !patterns)) return BB_ERR_INVALID;
    std::string e;
    std::vector<Pattern> pats(static_cast<size_t>(std::max(0, n_patterns)));
    for (int i = 0; i < n_patterns; i++) if (!parse_pattern(patterns[i], pats[i], e)) { set_err(err, errlen, e); return BB_ERR_INVALID; }
    TsvWriter out, drop;
    if (!out.open(output, e)) { set_err(err, errlen, e); return BB_ERR_IO; }
    if (dropped && !drop.open(dropped, e)) { set_err(err, errlen, e); return BB_ERR_IO; }
    uint64_t total = 0, kept = 0, dropped_n = 0;
    const int rc = for_each_read(annotated, [&](std::vector<Row>& group) {
        total++;
        if (check_filter_pass(group, pats)) { kept++; for (const Row& r : group) out.write(r); }
        else { dropped_n++; if (dropped) for (const Row& r : group) drop.write(r); }
    }, e);
    if (rc != BB_OK) { set_err(err, errlen, e); return rc; }
    if (!out.close() || !drop.close()) { set_err(err, errlen, "write error"); return BB_ERR_IO; }
    if (counts) { counts[0] = total; counts[1] = kept; counts[2] = dropped_n; }
    return BB_OK;
}

int bb_inspect(const char* annotated, int32_t top_n, const char* read_pattern_out, int32_t bucket_size, char* err, size_t errlen) {
    if (!annotated || bucket_size <= 0) return BB_ERR_INVALID;
    std::string e;
    FILE* rp = nullptr;
    if (read_pattern_out) { rp = std::fopen(read_pattern_out, "w"); if (!rp) { set_err(err, errlen, std::string(read_pattern_out) + ": " + std::strerror(errno)); return BB_ERR_IO; } }
    std::unordered_map<std::string, size_t> index;
    std::vector<std::pair<std::string, uint64_t>> counts;          // first-seen order (the reference's HashMap order is arbitrary)
    const int rc = for_each_read(annotated, [&](std::vector<Row>& group) {
        const std::string label = group_structure(group, static_cast<uint64_t>(bucket_size));
        if (rp) std::fprintf(rp, "%s\t%s\n", group.front().read_id.c_str(), label.c_str());
        auto it = index.find(label);
        if (it == index.end()) { index[label] = counts.size(); counts.push_back({label, 1}); } else counts[it->second].second++;
    }, e);
    if (rp) std::fclose(rp);
    if (rc != BB_OK) { set_err(err, errlen, e); return rc; }
    std::printf("Found %zu unique patterns\n", counts.size());
    std::stable_sort(counts.begin(), counts.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    const bool color = isatty(1) && !std::getenv("NO_COLOR");
    for (size_t i = 0; i < counts.size() && i < static_cast<size_t>(std::max(0, top_n)); i++) {
        std::printf("\tPattern %zu: %llu occurrences\n", i + 1, static_cast<unsigned long long>(counts[i].second));
        std::printf("\t\t%s\n", (color ? colorize(counts[i].first) : counts[i].first).c_str());
    }
    std::printf("Showed %d / %zu patterns\n", top_n, counts.size());
    return BB_OK;
}

int bb_trim(const char* filtered, const char* const* fastq, int32_t n_fastq, const char* out_dir, const bb_trim_opts* opts, uint64_t counts[4],
            char* err, size_t errlen) {
    if (!filtered || !out_dir || !opts) return BB_ERR_INVALID;
    const bb_trim_opts& o = *opts;
    struct stat st;
    if (stat(out_dir, &st) != 0) {                         // create_dir_all
        std::string path(out_dir);
        for (size_t p = 1; p <= path.size(); p++) if (p == path.size() || path[p] == '/') ::mkdir(path.substr(0, p).c_str(), 0755);
    }
    if (o.sort_labels && o.only_side) { set_err(err, errlen, "Cannot enable only keeping left/right label and sorting; this is ambiguous"); return BB_ERR_INVALID; }
    std::string e;
    std::unordered_map<std::string, std::vector<Row>> by_read;      // trim.rs:335-355: all annotations, grouped by read id
    {
        TsvReader rd;
        if (!rd.open(filtered, e)) { set_err(err, errlen, e); return BB_ERR_IO; }
        Row r;
        for (;;) {
            const int rc = rd.next(r, e);
            if (rc < 0) { set_err(err, errlen, "Failed to parse annotation line: " + e); return BB_ERR_INVALID; }
            if (rc == 0) break;
            by_read[r.read_id].push_back(r);
        }
    }
    FILE* failed = nullptr;
    if (o.failed_out) { failed = std::fopen(o.failed_out, "w"); if (!failed) { set_err(err, errlen, std::string(o.failed_out) + ": " + std::strerror(errno)); return BB_ERR_IO; } }
    if (n_fastq <= 0 || !fastq) { if (failed) std::fclose(failed); set_err(err, errlen, "No FASTQ input files provided"); return BB_ERR_IO; }
    std::map<std::string, std::unique_ptr<FastqOut>> writers;
    uint64_t total = 0, trimmed = 0, split = 0, failed_n = 0;
    std::vector<std::string> paths(fastq, fastq + n_fastq);
    int rc = BB_OK;
    // one record -> its trimmed reads appended to per-label text buffers (the reference's per-record body, trim.rs:375-417)
    struct Out { std::vector<std::pair<std::string, std::string>> by_label; std::string failed_ids; uint64_t total = 0, trimmed = 0, split = 0, failed = 0; };
    auto one_record = [&](const char* id, size_t id_len, const char* desc, size_t desc_len, const char* seq, const char* qual, size_t seq_len, Out& O) {
        O.total++;
        const std::string read_id(id, id_len);
        auto it = by_read.find(read_id);
        if (it == by_read.end()) return;
        const std::vector<TrimmedRead> results = process_read_and_anno(seq, qual, seq_len, it->second, o);
        if (!results.empty()) O.trimmed++;
        else { O.failed++; O.failed_ids += read_id; O.failed_ids += '\n'; }
        if (results.size() > 1) O.split++;
        for (const TrimmedRead& t : results) {
            std::string* dst = nullptr;
            for (auto& kv : O.by_label) if (kv.first == t.label) { dst = &kv.second; break; }
            if (!dst) { O.by_label.emplace_back(t.label, std::string()); dst = &O.by_label.back().second; }
            *dst += '@'; *dst += read_id; *dst += t.suffix;
            if (o.write_full_header && desc_len) { *dst += ' '; dst->append(desc, desc_len); }
            *dst += '\n'; *dst += t.seq; *dst += "\n+\n"; *dst += t.qual; *dst += '\n';
        }
    };
    // buffers of one chunk / stretch of records -> the per-label files, in input order (files are created in first-use order)
    auto flush = [&](Out& O) {
        total += O.total; trimmed += O.trimmed; split += O.split; failed_n += O.failed;
        if (failed && !O.failed_ids.empty() && std::fwrite(O.failed_ids.data(), 1, O.failed_ids.size(), failed) != O.failed_ids.size()) { e = std::string("Failed to write ") + o.failed_out; rc = BB_ERR_IO; }
        for (auto& kv : O.by_label) {
            auto w = writers.find(kv.first);
            if (w == writers.end()) {
                auto fo = std::make_unique<FastqOut>();
                const std::string path = std::string(out_dir) + "/" + kv.first + (o.gzip ? ".trimmed.fastq.gz" : ".trimmed.fastq");
                if (!fo->open(path, o.gzip != 0, e)) { rc = BB_ERR_IO; return; }
                w = writers.emplace(kv.first, std::move(fo)).first;
            }
            w->second->write(kv.second);
        }
        O = Out();
    };
    // Plain regular files are mapped, cut into chunks at record boundaries and cut/trimmed by several threads (the annotations are
    // read-only by now); the calling thread appends the chunks' buffers in input order, so every output file holds exactly what
    // the reference's single-threaded loop writes.  Anything else (gzip, pipes) goes through one reader.
    int n_threads = o.threads > 0 ? o.threads : static_cast<int>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
    for (const std::string& path : paths) {
        if (rc != BB_OK) break;
        int fd = -1; const char* map = nullptr; size_t size = 0;
        if (n_threads > 1) {
            fd = ::open(path.c_str(), O_RDONLY);
            struct stat st;
            unsigned char magic[2] = {0, 0};
            if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 && pread(fd, magic, 2, 0) == 2 && !(magic[0] == 0x1f && magic[1] == 0x8b)) {
                void* mm = mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
                if (mm != MAP_FAILED) { map = static_cast<const char*>(mm); size = static_cast<size_t>(st.st_size); madvise(mm, size, MADV_SEQUENTIAL); }
            }
        }
        if (map) {
            size_t chunk = 16u << 20;
            if (const char* ck = std::getenv("BB_TRIM_CHUNK_KB")) chunk = static_cast<size_t>(std::max(1, std::atoi(ck))) << 10;   // test knob
            const size_t n_chunks = (size + chunk - 1) / chunk, window = static_cast<size_t>(2 * n_threads);
            std::vector<Out> outs(n_chunks);
            std::vector<uint8_t> done(n_chunks, 0);              // 1 = parsed, 2 = error
            std::vector<std::string> errs(n_chunks);
            std::atomic<size_t> claim{0};
            std::mutex mu; std::condition_variable cv;
            size_t consumed = 0; bool stop = false;
            auto worker = [&]() {
                for (;;) {
                    const size_t c = claim.fetch_add(1);
                    if (c >= n_chunks) return;
                    { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return stop || c < consumed + window; }); if (stop) return; }
                    const size_t begin = bb::fastq_record_start(map, size, c * chunk), end = bb::fastq_record_start(map, size, std::min(size, (c + 1) * chunk));
                    size_t p = begin; bool ok = true;
                    while (p < end) {
                        bb::FastqRec r; const char* what = nullptr;
                        const int st = bb::fastq_record_at(map, size, p, r, what);
                        if (st == 1) continue;
                        if (st < 0) { errs[c] = std::string(what) + " in " + path; ok = false; break; }
                        size_t idl, doff;
                        bb::fastq_split_header(r.head, r.head_len, idl, doff);
                        one_record(r.head, idl, r.head + doff, r.head_len - doff, r.seq, r.qual, r.seq_len, outs[c]);
                    }
                    { std::lock_guard<std::mutex> lk(mu); done[c] = ok ? 1 : 2; }
                    cv.notify_all();
                }
            };
            std::vector<std::thread> pool;
            for (int t = 0; t < n_threads; t++) pool.emplace_back(worker);
            for (size_t c = 0; c < n_chunks && rc == BB_OK; c++) {
                { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done[c] != 0; }); }
                if (done[c] == 2) { e = errs[c]; rc = BB_ERR_IO; }
                else flush(outs[c]);
                { std::lock_guard<std::mutex> lk(mu); consumed = c + 1; }
                cv.notify_all();
            }
            { std::lock_guard<std::mutex> lk(mu); stop = true; }
            cv.notify_all();
            for (auto& t : pool) t.join();
            munmap(const_cast<char*>(map), size);
            ::close(fd);
            continue;
        }
        if (fd >= 0) ::close(fd);
        bb::FastqReader reader(std::vector<std::string>{path});
        bb::FastqReader::View v;
        Out O;
        std::string re;
        while (rc == BB_OK && reader.next(v, re)) {
            one_record(v.id, v.id_len, v.desc, v.desc_len, v.seq, v.qual, v.seq_len, O);
            if (O.total >= 4096) flush(O);
        }
        if (rc == BB_OK) flush(O);
        if (rc == BB_OK && !re.empty()) { e = re; rc = BB_ERR_IO; }
    }
    if (failed && std::fclose(failed) != 0 && rc == BB_OK) { e = std::string("Failed to write ") + o.failed_out; rc = BB_ERR_IO; }
    if (rc == BB_OK && !e.empty()) rc = BB_ERR_IO;
    for (auto& w : writers)                              // every per-label file is closed here, so that a truncated output is an error, not a count
        if (!w.second->close() && rc == BB_OK) { e = "Failed to write trimmed reads of label '" + w.first + "' in " + out_dir; rc = BB_ERR_IO; }
    if (rc != BB_OK) { set_err(err, errlen, e); return rc; }
    if (counts) { counts[0] = total; counts[1] = trimmed; counts[2] = split; counts[3] = failed_n; }
    return BB_OK;
}

}  // extern "C"
