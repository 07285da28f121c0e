// hostlogic.hpp -- the pieces of the tidk path that stay on the host by design
// (SURVEY.md 8b): they touch a handful of k-byte labels per run, never the sequence.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace tidk_host {

// utils.rs:49-58 switch_base / utils.rs:37-46 reverse_complement
inline char switch_base(char c) {
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'T': return 'A';
    case 'G': return 'C';
    default: return 'N';
    }
}
inline std::string reverse_complement(const std::string &s) {
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); i++) r[s.size() - 1 - i] = switch_base(s[i]);
    return r;
}
inline std::string ascii_upper(std::string s) {
    for (char &c : s) if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return s;
}
// explore.rs:438-445 check_repeats: shortest period
inline size_t period(const std::string &s) {
    const size_t n = s.size();
    for (size_t d = 1; d < n; d++)
        if (s.compare(d, n - d, s, 0, n - d) == 0) return d;
    return n;
}
inline std::string min_rotation(const std::string &s) {
    if (s.empty()) return s;
    const std::string d = s + s;
    size_t best = 0;
    for (size_t i = 1; i < s.size(); i++)
        if (d.compare(i, s.size(), d, best, s.size()) < 0) best = i;
    return d.substr(best, s.size());
}
// utils.rs:147-166 lex_min
inline std::string lex_min(const std::string &s) {
    return std::min(min_rotation(s), min_rotation(reverse_complement(s)));
}
inline bool is_rotation(const std::string &a, const std::string &b) { // utils.rs:87-94
    return a.size() == b.size() && (b + b).find(a) != std::string::npos;
}
inline bool same_class(const std::string &a, const std::string &b) { // explore.rs:359-371
    return is_rotation(a, b) || is_rotation(reverse_complement(a), b) || is_rotation(a, reverse_complement(b));
}

// explore.rs:377-431 get_telomeric_repeat_estimates + filter_count_vec (explore.rs:455-464).
// Runs in arrival order -> rows sorted (count desc, key asc).  The reference's pairwise loop has a
// closed form per rotation/revcomp class (first two members m1 < m2 in arrival order contribute
// c[m1] + c[m2], doubled iff (m1, m2) == (0, 1); single-member classes of a group of >= 2 vanish;
// a group of one inserts its own count) -- used when every label of the group is over A/C/G/T/N,
// where revcomp is an involution; other alphabets take the literal loop.
inline std::vector<std::pair<std::string, long long>> estimates(const std::vector<std::string> &labels,
                                                                const std::vector<long long> &counts) {
    std::map<size_t, std::vector<size_t>> groups;
    for (size_t i = 0; i < labels.size(); i++) groups[labels[i].size()].push_back(i);
    std::map<std::string, long long> out;
    for (auto &kv : groups) {
        const std::vector<size_t> &idx = kv.second;
        if (idx.size() == 1) { out[lex_min(labels[idx[0]])] = (int32_t)counts[idx[0]]; continue; }
        bool involutive = true;
        for (size_t i : idx)
            for (char c : labels[i])
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') involutive = false;
        if (involutive) {
            std::unordered_map<std::string, std::string> canon;
            std::unordered_map<std::string, std::pair<size_t, long long>> first;
            std::unordered_set<std::string> done;
            for (size_t g = 0; g < idx.size(); g++) {
                const std::string &lab = labels[idx[g]];
                auto it = canon.find(lab);
                if (it == canon.end()) it = canon.emplace(lab, lex_min(lab)).first;
                const std::string &key = it->second;
                if (done.count(key)) continue;
                auto f = first.find(key);
                if (f == first.end()) first.emplace(key, std::make_pair(g, counts[idx[g]]));
                else {
                    int32_t v = (int32_t)f->second.second + (int32_t)counts[idx[g]];
                    if (f->second.first == 0 && g == 1) v = (int32_t)(v * 2);
                    out[key] = v;
                    done.insert(key);
                }
            }
        } else {
            std::vector<char> tracker(idx.size(), 0);
            for (size_t a = 0; a < idx.size(); a++)
                for (size_t b = a + 1; b < idx.size(); b++) {
                    if (same_class(labels[idx[a]], labels[idx[b]])) {
                        const std::string key = lex_min(labels[idx[a]]);
                        const int32_t add = (int32_t)counts[idx[a]] + (int32_t)counts[idx[b]];
                        auto it = out.find(key);
                        if (it == out.end()) it = out.emplace(key, add).first;
                        if (!tracker[a] && !tracker[b]) it->second = (int32_t)(it->second + add);
                    }
                    tracker[a] = tracker[b] = 1;
                }
        }
    }
    std::vector<std::pair<std::string, long long>> rows(out.begin(), out.end());
    std::stable_sort(rows.begin(), rows.end(), [](const auto &x, const auto &y) {
        if (x.second != y.second) return x.second > y.second;
        return x.first < y.first;
    });
    std::vector<std::pair<std::string, long long>> kept;
    for (auto &r : rows)
        if (period(r.first) > 3) kept.push_back(r);
    return kept;
}

} // namespace tidk_host
