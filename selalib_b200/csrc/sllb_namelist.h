// Host-only Fortran namelist reader shared by the simulation front-ends (the 6D simulation's C interface and the
// 2D2V / 1D1V namelist runners): &group  key = v1 v2, v3 ... /  with ! comments, quoted strings, .true./.false.,
// d-exponents.  Keys and group names are case-insensitive like in Fortran.
#pragma once
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace sllb {
namespace namelist {

typedef std::map<std::string, std::vector<std::string>> Group;
typedef std::map<std::string, Group> Namelist;

inline std::string lower(std::string s) {
    for (auto &c : s) c = (char)tolower((unsigned char)c);
    return s;
}

inline bool parse_namelist(const char *path, Namelist &out, std::string &err) {
    FILE *fp = fopen(path, "r");
    if (!fp) { err = std::string("failed to open file ") + path; return false; }
    std::string text;
    char buf[4096];
    while (fgets(buf, sizeof(buf), fp)) {
        std::string line(buf);
        bool inq = false; char qc = 0;
        for (size_t i = 0; i < line.size(); ++i) { // strip ! comments outside quotes
            char c = line[i];
            if (inq) { if (c == qc) inq = false; }
            else if (c == '"' || c == '\'') { inq = true; qc = c; }
            else if (c == '!') { line.erase(i); break; }
        }
        text += line;
        text += '\n';
    }
    fclose(fp);
    // tokenise
    std::vector<std::string> tok;
    for (size_t i = 0; i < text.size();) {
        char c = text[i];
        if (isspace((unsigned char)c) || c == ',') { ++i; continue; }
        if (c == '"' || c == '\'') {
            size_t j = text.find(c, i + 1);
            if (j == std::string::npos) { err = "unterminated string"; return false; }
            tok.push_back(std::string("\"") + text.substr(i + 1, j - i - 1));
            i = j + 1;
            continue;
        }
        if (c == '=' || c == '/') { tok.push_back(std::string(1, c)); ++i; continue; }
        size_t j = i;
        while (j < text.size() && !isspace((unsigned char)text[j]) && text[j] != ',' && text[j] != '=' && text[j] != '/' &&
               text[j] != '"' && text[j] != '\'')
            ++j;
        tok.push_back(text.substr(i, j - i));
        i = j;
    }
    Group *g = nullptr;
    std::vector<std::string> *cur = nullptr;
    for (size_t i = 0; i < tok.size(); ++i) {
        const std::string &t = tok[i];
        if (t[0] == '&') { g = &out[lower(t.substr(1))]; cur = nullptr; continue; }
        if (t == "/") { g = nullptr; cur = nullptr; continue; }
        if (!g) continue;
        if (i + 1 < tok.size() && tok[i + 1] == "=") { cur = &(*g)[lower(t)]; cur->clear(); ++i; continue; }
        if (cur) cur->push_back(t);
    }
    return true;
}
inline const std::vector<std::string> *find(const Namelist &n, const char *grp, const char *key) {
    auto g = n.find(grp);
    if (g == n.end()) return nullptr;
    auto k = g->second.find(key);
    if (k == g->second.end() || k->second.empty()) return nullptr;
    return &k->second;
}
inline double to_double(std::string s) {
    for (auto &c : s) if (c == 'd' || c == 'D') c = 'e'; // Fortran exponents
    return atof(s.c_str());
}
inline double get_real(const Namelist &n, const char *g, const char *k, double def, int idx = 0) {
    auto v = find(n, g, k);
    return (v && (int)v->size() > idx) ? to_double((*v)[idx]) : def;
}
inline int get_int(const Namelist &n, const char *g, const char *k, int def, int idx = 0) {
    auto v = find(n, g, k);
    return (v && (int)v->size() > idx) ? atoi((*v)[idx].c_str()) : def;
}
inline std::string get_str(const Namelist &n, const char *g, const char *k, const char *def) {
    auto v = find(n, g, k);
    if (!v) return def;
    std::string s = (*v)[0];
    if (!s.empty() && s[0] == '"') s = s.substr(1);
    return s;
}
inline bool get_bool(const Namelist &n, const char *g, const char *k, bool def) {
    auto v = find(n, g, k);
    if (!v) return def;
    std::string s = lower((*v)[0]);
    return s.find('t') != std::string::npos;
}

} // namespace namelist
} // namespace sllb
