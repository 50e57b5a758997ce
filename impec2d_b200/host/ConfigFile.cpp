#include "ConfigFile.h"
#include <fstream>
#include <sstream>
#include <cctype>
#include <cstdlib>

namespace {
struct Parser {
    const std::string &s; std::string name; size_t p = 0; int line = 1;
    Parser(const std::string &text, const std::string &n) : s(text), name(n) {}
    [[noreturn]] void fail(const std::string &m) { std::ostringstream o; o << name << ":" << line << ": " << m; throw std::runtime_error(o.str()); }
    void skip() {
        for (;;) {
            while (p < s.size() && isspace((unsigned char)s[p])) { if (s[p] == '\n') line++; p++; }
            if (p >= s.size()) return;
            if (s[p] == '#' || (s[p] == '/' && p + 1 < s.size() && s[p + 1] == '/')) { while (p < s.size() && s[p] != '\n') p++; continue; }
            if (s[p] == '/' && p + 1 < s.size() && s[p + 1] == '*') {
                p += 2;
                while (p + 1 < s.size() && !(s[p] == '*' && s[p + 1] == '/')) { if (s[p] == '\n') line++; p++; }
                if (p + 1 >= s.size()) fail("unterminated comment");
                p += 2; continue;
            }
            return;
        }
    }
    bool eat(char c) { skip(); if (p < s.size() && s[p] == c) { p++; return true; } return false; }
    std::string ident() {
        skip(); size_t b = p;
        while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '_' || s[p] == '-' || s[p] == '*')) p++;
        if (b == p) fail("expected a setting name");
        return s.substr(b, p - b);
    }
    CfgSetting value() {
        skip();
        CfgSetting v;
        if (p >= s.size()) fail("unexpected end of file");
        const char c = s[p];
        if (c == '{') { p++; v.type = CfgSetting::GROUP; members(v, '}'); return v; }
        if (c == '[' || c == '(') {
            const char close = c == '[' ? ']' : ')'; p++; v.type = CfgSetting::LIST;
            for (;;) { skip(); if (eat(close)) break; v.items.push_back(value()); skip(); eat(','); }
            return v;
        }
        if (c == '"') {
            v.type = CfgSetting::STRING;
            while (eat('"')) {                                        // adjacent strings concatenate
                while (p < s.size() && s[p] != '"') { if (s[p] == '\\' && p + 1 < s.size()) p++; v.sval += s[p++]; }
                if (p >= s.size()) fail("unterminated string");
                p++;
            }
            return v;
        }
        size_t b = p;
        while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '.' || s[p] == '+' || s[p] == '-')) p++;
        std::string tok = s.substr(b, p - b);
        if (tok.empty()) fail(std::string("unexpected character '") + c + "'");
        std::string low; for (char ch : tok) low += (char)tolower((unsigned char)ch);
        if (low == "true" || low == "false") { v.type = CfgSetting::BOOL; v.bval = low == "true"; return v; }
        const bool is_float = tok.find_first_of(".eE") != std::string::npos && !(tok.size() > 2 && tok[0] == '0' && (tok[1] == 'x' || tok[1] == 'X'));
        char *end = nullptr;
        if (is_float) { v.type = CfgSetting::FLOAT; v.fval = strtod(tok.c_str(), &end); }
        else { if (!tok.empty() && (tok.back() == 'L')) tok.pop_back(); v.type = CfgSetting::INT; v.ival = strtoll(tok.c_str(), &end, 0); }
        if (!end || (*end != 0 && *end != 'L')) fail("cannot parse value '" + tok + "'");
        return v;
    }
    void members(CfgSetting &g, char close) {
        for (;;) {
            skip();
            if (close && eat(close)) return;
            if (p >= s.size()) { if (close) fail("missing '}'"); return; }
            const std::string k = ident();
            skip();
            if (!(eat(':') || eat('='))) fail("expected ':' or '=' after '" + k + "'");
            g.members.push_back(std::make_pair(k, value()));
            skip();
            while (eat(';') || eat(',')) {}
        }
    }
};
}  // namespace

bool CfgSetting::exists(const std::string &k) const { for (auto &m : members) if (m.first == k) return true; return false; }
const CfgSetting &CfgSetting::operator[](const std::string &k) const {
    for (auto &m : members) if (m.first == k) return m.second;
    throw std::runtime_error("setting '" + k + "' not found");
}
bool CfgSetting::lookupValue(const std::string &k, int &v) const { if (!exists(k)) return false; const CfgSetting &s = (*this)[k]; if (s.type != INT) return false; v = (int)s.ival; return true; }
bool CfgSetting::lookupValue(const std::string &k, double &v) const { if (!exists(k)) return false; const CfgSetting &s = (*this)[k]; if (s.type == FLOAT) v = s.fval; else if (s.type == INT) v = (double)s.ival; else return false; return true; }
bool CfgSetting::lookupValue(const std::string &k, bool &v) const { if (!exists(k)) return false; const CfgSetting &s = (*this)[k]; if (s.type != BOOL) return false; v = s.bval; return true; }
bool CfgSetting::lookupValue(const std::string &k, std::string &v) const { if (!exists(k)) return false; const CfgSetting &s = (*this)[k]; if (s.type != STRING) return false; v = s.sval; return true; }

void ConfigFile::readString(const std::string &text, const std::string &name) {
    root_ = CfgSetting(); root_.type = CfgSetting::GROUP;
    Parser ps(text, name);
    ps.members(root_, 0);
}
void ConfigFile::readFile(const std::string &path) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) throw std::runtime_error("I/O error while reading file " + path);
    std::stringstream ss; ss << f.rdbuf();
    readString(ss.str(), path);
}
