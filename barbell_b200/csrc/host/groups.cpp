// Pattern-set construction on the host: kit presets -> query sequences -> BarcodeGroup geometry.
// Re-statement (not a translation) of
//   BarcodeGroup::new            reference src/annotate/barcodes.rs:106-197
//   BarcodeGroup::new_from_kit   reference src/annotate/barcodes.rs:251-299
//   BarcodeGroup::new_from_fasta reference src/annotate/barcodes.rs:302-315
//   get_kit_info/get_barcodes/lookup_barcode_seq  reference src/kits/kits.rs:635-816, 1074-1103
//   get_edit_cut_off             reference src/annotate/edit_model.rs:2-11
// Where the reference panics, these functions return BB_ERR_KIT with the reason in `err`.
#include "groups.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace bb {
namespace {
#include "../kit_tables.inc"

constexpr int kPadding = 10;   // reference src/lib.rs:10

bool iupac_ok(char c) {
    static const char* ok = "ACGTURYSWKMBDHVNX";
    return c != 0 && std::strchr(ok, std::toupper(static_cast<unsigned char>(c))) != nullptr;
}

struct Label { std::string prefix; int number = 0; bool a_flag = false; bool ok = false; };

Label parse_label(const std::string& s) {   // kits.rs:710-739
    Label l; size_t i = 0;
    while (i < s.size() && std::isalpha(static_cast<unsigned char>(s[i]))) l.prefix.push_back(static_cast<char>(std::toupper(s[i++])));
    std::string num;
    while (i < s.size() && std::isdigit(static_cast<unsigned char>(s[i]))) num.push_back(s[i++]);
    l.a_flag = i < s.size() && std::toupper(static_cast<unsigned char>(s[i])) == 'A';
    if (num.empty()) return l;
    l.number = std::atoi(num.c_str()); l.ok = true;
    return l;
}

std::string two_digit(const char* pfx, int n) { char b[16]; std::snprintf(b, sizeof b, "%s%02d", pfx, n); return b; }

// kits.rs:741-816
bool label_range(const std::string& from, const std::string& to, bool use_12a_flag, std::vector<std::string>& out, std::string& err) {
    Label a = parse_label(from), b = parse_label(to);
    if (!a.ok || !b.ok) { err = "Invalid numeric part in label"; return false; }
    if (a.prefix != b.prefix) { err = "Mismatched label prefixes: " + a.prefix + " vs " + b.prefix; return false; }
    int start = std::min(a.number, b.number), end = std::max(a.number, b.number);
    bool amplicon = a.prefix == "AB";
    int limit = amplicon ? 24 : 96;
    if (start < 1 || end > limit) { err = "label range out of bounds"; return false; }
    out.clear();
    for (int n = start; n <= end; n++) out.push_back(two_digit(amplicon ? "AB" : "BC", n));
    bool use_12a = use_12a_flag || ((a.a_flag || b.a_flag) && start <= 12 && 12 <= end);
    if (use_12a) for (auto& s : out) if (s == "BC12") s = "BC12A";
    if (a.prefix == "NB") for (auto& s : out) if (s.rfind("BC", 0) == 0) s.replace(0, 2, "NB");
    if (a.prefix == "RBK") {
        static const int special[] = {26, 39, 40, 48, 54, 60};
        for (auto& s : out)
            if (s.rfind("BC", 0) == 0 && s.size() >= 4) {
                int n = std::atoi(s.substr(2, 2).c_str());
                if (std::find(std::begin(special), std::end(special), n) != std::end(special)) s.replace(0, 2, "RBK");
            }
    }
    return true;
}

// kits.rs:1074-1103
const char* barcode_seq(const std::string& label) {
    Label l = parse_label(label);
    if (!l.ok) return nullptr;
    int idx = l.number > 0 ? l.number - 1 : 0;
    if (l.prefix == "BC") { if (l.a_flag && l.number == 12) return BC12A_SEQ; return idx < 96 ? BC_SEQS[idx] : nullptr; }
    if (l.prefix == "NB") { if (l.a_flag && l.number == 12) return BC12A_SEQ; return idx < 96 ? NB_SEQS[idx] : nullptr; }
    if (l.prefix == "AB") return idx < 24 ? AB_SEQS[idx] : nullptr;
    if (l.prefix == "BP") return idx < 24 ? BP_SEQS[idx] : nullptr;
    if (l.prefix == "RBK") {
        for (const auto& s : RBK_SPECIAL) if (s.number == l.number) return s.seq;
        return idx < 96 ? BC_SEQS[idx] : nullptr;
    }
    return nullptr;
}

const KitPreset* find_kit(std::string kit, std::string& note) {   // kits.rs:635-708
    for (int attempt = 0; attempt < 2; attempt++) {
        for (const auto& kn : KIT_NAMES)
            if (kit == kn.kit)
                for (const auto& p : KIT_PRESETS) if (std::strcmp(p.id, kn.preset_id) == 0) return &p;
        if (kit.find('.') == std::string::npos) break;
        std::string nk = kit; std::replace(nk.begin(), nk.end(), '.', '-');
        note = "Your kit name used '.' (" + kit + ") instead of '-' replaced it with " + nk + " and trying again";
        kit = nk;
    }
    return nullptr;
}

void set_err(char* err, size_t errlen, const std::string& msg) {
    if (err && errlen) { std::snprintf(err, errlen, "%s", msg.c_str()); }
}
}  // namespace

int edit_cut_off(int l) {   // edit_model.rs:2-11
    double a = static_cast<double>(l);
    double v = std::ceil(0.5100 * a - 1.7312 * std::sqrt(a));
    return v > 0.0 ? static_cast<int>(v) : 0;
}

// barcodes.rs:106-197
bool Group::build(const std::vector<std::string>& seqs, const std::vector<std::string>& labels_in, int type, std::string& err) {
    if (seqs.size() != labels_in.size() || seqs.empty()) { err = "empty query group"; return false; }
    if (seqs.size() == 1) { err = "For now we only support 'groups': add a second query with the same flanks and a different barcode"; return false; }
    const size_t len = seqs[0].size();
    for (const auto& s : seqs) {
        if (s.size() != len) { err = "All sequences per group must be equally long"; return false; }
        for (char c : s) if (!iupac_ok(c)) { err = "Sequence contains character not supported by IUPAC"; return false; }
    }
    size_t pre = len, suf = len;   // longest common prefix / suffix (barcodes.rs:336-385)
    for (size_t q = 1; q < seqs.size(); q++) {
        size_t a = 0; while (a < len && seqs[0][a] == seqs[q][a]) a++;
        size_t b = 0; while (b < len && seqs[0][len - 1 - b] == seqs[q][len - 1 - b]) b++;
        pre = std::min(pre, a); suf = std::min(suf, b);
    }
    if (pre + suf >= len) { err = "No barcode region found, are you sure the input is unique sequences of <prefix><barcode><suffix>?"; return false; }
    if (pre == 0 && suf == 0) { err = "No prefix or suffix found, we can't search without having 'anchors'"; return false; }
    const size_t mask = len - pre - suf;
    flank = seqs[0].substr(0, pre) + std::string(mask, 'N') + seqs[0].substr(len - suf);
    prefix_len = static_cast<int>(pre); suffix_len = static_cast<int>(suf);
    pad0 = static_cast<int>(pre > static_cast<size_t>(kPadding) ? pre - kPadding : 0);
    pad1 = static_cast<int>(pre + mask + kPadding);             // NOT clamped (barcodes.rs:160-163)
    const size_t end = std::min(static_cast<size_t>(pad1), len);
    bar_len = static_cast<int>(end - pad0);
    barcodes.clear();
    for (const auto& s : seqs) barcodes += s.substr(pad0, end - pad0);
    labels = labels_in;
    bar0 = static_cast<int>(pre); bar1 = static_cast<int>(pre + mask - 1);   // inclusive end (barcodes.rs:192)
    match_type = type;
    k_flank = 0;                                                 // k_cutoff.unwrap_or(0), searcher.rs:435
    return true;
}

bb_group Group::view() const {
    bb_group g{};
    g.flank = flank.c_str(); g.flank_len = static_cast<int32_t>(flank.size()); g.k_flank = k_flank;
    g.bar0 = bar0; g.bar1 = bar1; g.pad0 = pad0; g.pad1 = pad1; g.match_type = match_type;
    g.n_barcodes = static_cast<int32_t>(labels.size()); g.bar_len = bar_len; g.barcodes = barcodes.c_str();
    return g;
}

void GroupSet::refresh() { views.clear(); for (const auto& g : groups) views.push_back(g.view()); }

bool GroupSet::from_kit(const std::string& kit, bool use_extended, std::string& err, std::string* note) {
    std::string n;
    const KitPreset* p = find_kit(kit, n);
    if (note) *note = n;
    if (!p) { err = "Unknown or unsupported kit: " + kit + ", please raise an issue"; return false; }
    for (int t = 0; t < p->n_templates; t++) {
        const TemplateSpec& ts = p->templates[t];
        if (ts.extended && !use_extended) continue;              // barcodes.rs:259-262
        std::vector<std::string> labels, seqs;
        if (!label_range(ts.label_from, ts.label_to, ts.use_12a, labels, err)) return false;
        for (const auto& l : labels) {
            const char* bs = barcode_seq(l);
            if (!bs) { err = "Barcode not found - odd - raise issue: " + l; return false; }
            seqs.push_back(std::string(ts.front) + bs + ts.rear);
        }
        Group g;
        if (!g.build(seqs, labels, ts.right_side ? BB_RTAG : BB_FTAG, err)) return false;
        groups.push_back(std::move(g));
    }
    refresh();
    return true;
}

bool read_fasta(const std::string& path, std::vector<std::string>& seqs, std::vector<std::string>& labels, std::string& err) {
    std::ifstream in(path);
    if (!in) { err = "Query file not found: " + path; return false; }
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            // needletail id() is the whole header line after '>' (barcodes.rs:311)
            labels.push_back(line.substr(1)); seqs.emplace_back();
        } else if (!seqs.empty()) {
            for (char c : line) {                               // normalize(true): upper-case, keep IUPAC
                if (std::isspace(static_cast<unsigned char>(c))) continue;
                char u = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
                if (u == '.' || u == '~' || u == '-') u = 'N';
                seqs.back().push_back(u);
            }
        }
    }
    if (seqs.empty()) { err = "invalid record: no sequences in " + path; return false; }
    return true;
}

bool GroupSet::add_fasta(const std::string& path, int type, std::string& err) {
    std::vector<std::string> seqs, labels;
    if (!read_fasta(path, seqs, labels, err)) return false;
    return add(seqs, labels, type, err);
}

bool GroupSet::add(const std::vector<std::string>& seqs, const std::vector<std::string>& labels, int type, std::string& err) {
    Group g;
    if (!g.build(seqs, labels, type, err)) return false;
    groups.push_back(std::move(g));
    refresh();
    return true;
}

void GroupSet::set_flank_threshold(int max_flank_errors, std::vector<int>* chosen) {   // annotator.rs:216-229
    for (auto& g : groups) {
        g.k_flank = max_flank_errors >= 0 ? max_flank_errors : edit_cut_off(g.prefix_len + g.suffix_len);
        if (chosen) chosen->push_back(g.k_flank);
    }
    refresh();
}
}  // namespace bb

// ---------------- C ABI ----------------
struct bb_groupset { bb::GroupSet gs; };

extern "C" {
int bb_groups_from_kit(const char* kit, int use_extended, bb_groupset** out, char* err, size_t errlen) {
    if (!kit || !out) return BB_ERR_INVALID;
    auto* h = new bb_groupset();
    std::string e;
    if (!h->gs.from_kit(kit, use_extended != 0, e, nullptr)) { bb::set_err(err, errlen, e); delete h; return BB_ERR_KIT; }
    *out = h;
    return BB_OK;
}
int bb_groups_from_fasta(const char* const* paths, const int32_t* types, int32_t n, bb_groupset** out, char* err, size_t errlen) {
    if (!paths || !types || !out || n <= 0) return BB_ERR_INVALID;
    auto* h = new bb_groupset();
    std::string e;
    for (int i = 0; i < n; i++)
        if (!h->gs.add_fasta(paths[i], types[i], e)) { bb::set_err(err, errlen, e); delete h; return BB_ERR_KIT; }
    *out = h;
    return BB_OK;
}
int bb_groups_add(bb_groupset** out, const char* const* seqs, const char* const* labels, int32_t n, int32_t type, char* err, size_t errlen) {
    if (!out || !seqs || !labels || n <= 0) return BB_ERR_INVALID;
    std::vector<std::string> s, l;
    for (int i = 0; i < n; i++) {
        std::string q = seqs[i];
        for (auto& c : q) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
        s.push_back(q); l.push_back(labels[i]);
    }
    bool fresh = *out == nullptr;
    bb_groupset* h = fresh ? new bb_groupset() : *out;
    std::string e;
    if (!h->gs.add(s, l, type, e)) { bb::set_err(err, errlen, e); if (fresh) delete h; return BB_ERR_KIT; }
    *out = h;
    return BB_OK;
}
int bb_groups_set_flank_threshold(bb_groupset* gs, int32_t max_flank_errors) {
    if (!gs) return BB_ERR_INVALID;
    gs->gs.set_flank_threshold(max_flank_errors, nullptr);
    return BB_OK;
}
int32_t bb_groups_count(const bb_groupset* gs) { return gs ? static_cast<int32_t>(gs->gs.groups.size()) : 0; }
const bb_group* bb_groups_data(const bb_groupset* gs) { return gs && !gs->gs.views.empty() ? gs->gs.views.data() : nullptr; }
const char* bb_groups_label(const bb_groupset* gs, int32_t g, int32_t idx) {
    if (!gs || g < 0 || g >= static_cast<int32_t>(gs->gs.groups.size())) return nullptr;
    if (idx < 0) return "flank";
    const auto& L = gs->gs.groups[g].labels;
    return idx < static_cast<int32_t>(L.size()) ? L[idx].c_str() : nullptr;
}
void bb_groups_free(bb_groupset* gs) { delete gs; }
int32_t bb_edit_cut_off(int32_t l) { return bb::edit_cut_off(l); }
int bb_label_range(const char* from, const char* to, int use_12a, char* out, size_t outlen) {
    if (!from || !to || !out || !outlen) return BB_ERR_INVALID;
    std::vector<std::string> labels; std::string e;
    if (!bb::label_range(from, to, use_12a != 0, labels, e)) { std::snprintf(out, outlen, "%s", e.c_str()); return BB_ERR_KIT; }
    std::string joined;
    for (size_t i = 0; i < labels.size(); i++) { if (i) joined += ','; joined += labels[i]; }
    if (joined.size() + 1 > outlen) return BB_ERR_OVERFLOW;
    std::snprintf(out, outlen, "%s", joined.c_str());
    return static_cast<int>(labels.size());
}
const char* bb_lookup_barcode_seq(const char* label) { return label ? bb::barcode_seq(label) : nullptr; }
int bb_kit_info(const char* kit, char* name, size_t namelen, char* ranges, size_t rangeslen, int* double_label, char* err, size_t errlen) {
    if (!kit) return BB_ERR_INVALID;
    std::string note;
    const bb::KitPreset* p = bb::find_kit(kit, note);
    if (!p) { bb::set_err(err, errlen, std::string("Unknown or unsupported kit: ") + kit + ", please raise an issue"); return BB_ERR_KIT; }
    if (name && namelen) std::snprintf(name, namelen, "%s", p->name);
    if (ranges && rangeslen) {
        std::string r;
        for (int t = 0; t < p->n_templates; t++) r += std::string(t ? "; " : "") + p->templates[t].label_from + " - " + p->templates[t].label_to;
        std::snprintf(ranges, rangeslen, "%s", r.c_str());
    }
    if (double_label) *double_label = p->double_label_patterns ? 1 : 0;
    return BB_OK;
}
int bb_abi_version(void) { return BB_ABI_VERSION; }
}
