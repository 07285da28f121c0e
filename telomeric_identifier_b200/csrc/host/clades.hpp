// clades.hpp -- clade table lookup (stays on the host, clades.rs:96-141).
#pragma once
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace tidk_host {

inline std::string database_path() { // build.rs:30-40: dirs::data_dir()/tidk/tidk_database.csv
    const char *xdg = getenv("XDG_DATA_HOME");
    std::string base = (xdg && *xdg) ? xdg : std::string(getenv("HOME") ? getenv("HOME") : ".") + "/.local/share";
    return base + "/tidk/tidk_database.csv";
}

// minimal RFC-4180 reader: quoted fields, doubled quotes, embedded newlines
inline std::vector<std::vector<std::string>> read_csv(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open clade database " + path + " (run `tidk build`, or pass --database)");
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string d = ss.str();
    std::vector<std::vector<std::string>> rows;
    std::vector<std::string> row;
    std::string field;
    bool quoted = false, any = false;
    for (size_t i = 0; i < d.size(); i++) {
        char c = d[i];
        if (quoted) {
            if (c == '"') {
                if (i + 1 < d.size() && d[i + 1] == '"') { field += '"'; i++; }
                else quoted = false;
            } else field += c;
        } else if (c == '"') { quoted = true; any = true; }
        else if (c == ',') { row.push_back(field); field.clear(); any = true; }
        else if (c == '\n' || c == '\r') {
            if (c == '\r' && i + 1 < d.size() && d[i + 1] == '\n') i++;
            if (any || !field.empty()) { row.push_back(field); rows.push_back(row); }
            row.clear(); field.clear(); any = false;
        } else { field += c; any = true; }
    }
    if (any || !field.empty()) { row.push_back(field); rows.push_back(row); }
    return rows;
}

struct CladeTable {
    std::vector<std::string> order, repeat;
};

inline CladeTable load_clades(const std::string &path) {
    auto rows = read_csv(path);
    if (rows.empty()) throw std::runtime_error("empty clade database");
    int io = -1, ir = -1;
    for (size_t i = 0; i < rows[0].size(); i++) {
        if (rows[0][i] == "Order") io = (int)i;
        if (rows[0][i] == "Telomeric repeat") ir = (int)i;
    }
    if (io < 0 || ir < 0) throw std::runtime_error("clade database lacks the Order / Telomeric repeat columns");
    CladeTable t;
    for (size_t r = 1; r < rows.size(); r++) {
        const auto &row = rows[r];
        t.order.push_back((size_t)io < row.size() ? row[io] : "");
        t.repeat.push_back((size_t)ir < row.size() ? row[ir] : "");
    }
    return t;
}

// clades.rs:96-115
inline std::vector<std::string> get_clades(const CladeTable &t) {
    std::vector<std::string> out;
    for (const auto &o : t.order)
        if (out.empty() || out.back() != o) out.push_back(o);
    std::vector<std::string> kept;
    for (auto &o : out) if (!o.empty()) kept.push_back(o);
    return kept;
}

// clades.rs:119-141
inline std::vector<std::string> return_telomere_sequence(const CladeTable &t, const std::string &clade) {
    std::vector<std::string> seqs;
    for (size_t i = 0; i < t.order.size(); i++)
        if (t.order[i] == clade) {
            bool seen = false;
            for (auto &s : seqs) if (s == t.repeat[i]) seen = true;
            if (!seen) seqs.push_back(t.repeat[i]);
        }
    return seqs;
}

} // namespace tidk_host
