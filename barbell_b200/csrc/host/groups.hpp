// Host-side pattern sets (see groups.cpp for the reference citations).
#pragma once
#include <string>
#include <vector>

#include "../../../include/barbell_b200.h"

namespace bb {
int edit_cut_off(int effective_len);

struct Group {
    std::string flank;                 // prefix + N*mask + suffix
    int prefix_len = 0, suffix_len = 0;
    int bar0 = 0, bar1 = 0, pad0 = 0, pad1 = 0, bar_len = 0;
    int match_type = BB_FTAG, k_flank = 0;
    std::string barcodes;              // n * bar_len
    std::vector<std::string> labels;
    bool build(const std::vector<std::string>& seqs, const std::vector<std::string>& labels, int type, std::string& err);
    bb_group view() const;
};

struct GroupSet {
    std::vector<Group> groups;
    std::vector<bb_group> views;       // pointers into `groups`
    void refresh();
    bool from_kit(const std::string& kit, bool use_extended, std::string& err, std::string* note);
    bool add_fasta(const std::string& path, int type, std::string& err);
    bool add(const std::vector<std::string>& seqs, const std::vector<std::string>& labels, int type, std::string& err);
    void set_flank_threshold(int max_flank_errors, std::vector<int>* chosen);
};
bool read_fasta(const std::string& path, std::vector<std::string>& seqs, std::vector<std::string>& labels, std::string& err);
}  // namespace bb
