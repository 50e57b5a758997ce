// libconfig.h++ -- SHIM, test infrastructure only.  libconfig++ is not installed in this image; the reference's SimParams.cpp
// (src/SimParams.cpp:42-471, compiled UNMODIFIED by oracle/Makefile.ref) uses Config::readFile/getRoot, Setting::operator[],
// lookupValue, getLength, the int/double conversions and two exception types.  They are mapped onto this repo's own reader of the
// cfg syntax (impec2d_b200/host/ConfigFile.h).  Difference from libconfig: lookupValue(name, double&) also accepts an integer
// literal (libconfig does only with auto-convert on); the shipped examples/*.cfg write every real with a decimal point.
#ifndef IMPEC_REF_SHIM_LIBCONFIG_HPP
#define IMPEC_REF_SHIM_LIBCONFIG_HPP
#include <string>
#include <exception>
#include <cstdlib>
#include "../../impec2d_b200/host/ConfigFile.h"

namespace libconfig {

class ConfigException : public std::exception {};
class FileIOException : public ConfigException {
public:
    const char *what() const noexcept override { return "FileIOException"; }
};
class ParseException : public ConfigException {
    std::string file_, err_; int line_;
public:
    ParseException(const std::string &f, int l, const std::string &e) : file_(f), err_(e), line_(l) {}
    const char *getFile() const { return file_.c_str(); }
    int getLine() const { return line_; }
    const char *getError() const { return err_.c_str(); }
    const char *what() const noexcept override { return "ParseException"; }
};
class SettingException : public ConfigException {
    std::string path_;
public:
    explicit SettingException(const std::string &p) : path_(p) {}
    const char *getPath() const { return path_.c_str(); }
    const char *what() const noexcept override { return "SettingException"; }
};
class SettingNotFoundException : public SettingException {
public:
    explicit SettingNotFoundException(const std::string &p) : SettingException(p) {}
    const char *what() const noexcept override { return "SettingNotFoundException"; }
};
class SettingTypeException : public SettingException {
public:
    explicit SettingTypeException(const std::string &p) : SettingException(p) {}
    const char *what() const noexcept override { return "SettingTypeException"; }
};

// A Setting is a view of one node of the parsed tree (the tree is owned by the Config).
class Setting {
    const CfgSetting *s_;
public:
    explicit Setting(const CfgSetting *s) : s_(s) {}
    // children are materialised on demand and kept alive for the life of the process: the reference holds `const Setting &`
    // to the results of operator[] (src/SimParams.cpp:52-55), exactly as libconfig allows
    const Setting &child(const CfgSetting *c) const { return *new Setting(c); }
    const Setting &operator[](const char *k) const {
        if (!s_->exists(k)) throw SettingNotFoundException(k);
        return child(&(*s_)[std::string(k)]);
    }
    const Setting &operator[](const std::string &k) const { return (*this)[k.c_str()]; }
    const Setting &operator[](int i) const {
        if (i < 0 || i >= s_->getLength()) throw SettingNotFoundException("[index]");
        return child(&(*s_)[i]);
    }
    int getLength() const { return s_->getLength(); }
    bool exists(const char *k) const { return s_->exists(k); }
    bool exists(const std::string &k) const { return s_->exists(k); }
    bool lookupValue(const char *k, int &v) const { return s_->lookupValue(k, v); }
    bool lookupValue(const char *k, unsigned int &v) const { int t; if (!s_->lookupValue(k, t)) return false; v = (unsigned)t; return true; }
    bool lookupValue(const char *k, double &v) const { return s_->lookupValue(k, v); }
    bool lookupValue(const char *k, float &v) const { double t; if (!s_->lookupValue(k, t)) return false; v = (float)t; return true; }
    bool lookupValue(const char *k, bool &v) const { return s_->lookupValue(k, v); }
    bool lookupValue(const char *k, std::string &v) const { return s_->lookupValue(k, v); }
    template <typename V> bool lookupValue(const std::string &k, V &v) const { return lookupValue(k.c_str(), v); }
    operator int() const { if (s_->type != CfgSetting::INT) throw SettingTypeException("int"); return (int)s_->ival; }
    operator double() const {
        if (s_->type == CfgSetting::INT) return (double)s_->ival;
        if (s_->type != CfgSetting::FLOAT) throw SettingTypeException("float");
        return s_->fval;
    }
    operator bool() const { if (s_->type != CfgSetting::BOOL) throw SettingTypeException("bool"); return s_->bval; }
    operator std::string() const { if (s_->type != CfgSetting::STRING) throw SettingTypeException("string"); return s_->sval; }
};

class Config {
    ConfigFile f_;
    Setting *root_ = nullptr;
public:
    Config() {}
    Config(const Config &) = delete;
    Config &operator=(const Config &) = delete;
    ~Config() { delete root_; }
    void readFile(const char *path) {
        FILE *t = fopen(path, "r");
        if (!t) throw FileIOException();
        fclose(t);
        try { f_.readFile(path); }
        catch (const std::runtime_error &e) {                       // "file:line: message"
            const std::string m = e.what();
            const size_t a = m.find(':'), b = a == std::string::npos ? a : m.find(':', a + 1);
            int line = 0; std::string err = m;
            if (b != std::string::npos) { line = atoi(m.substr(a + 1, b - a - 1).c_str()); err = m.substr(b + 1); }
            throw ParseException(path, line, err);
        }
        delete root_; root_ = new Setting(&f_.getRoot());
    }
    const Setting &getRoot() const { return *root_; }
    const Setting &lookup(const std::string &path) const {
        const Setting *s = root_;
        size_t p = 0;
        while (p < path.size()) { size_t d = path.find('.', p); if (d == std::string::npos) d = path.size(); s = &(*s)[path.substr(p, d - p)]; p = d + 1; }
        return *s;
    }
};

}  // namespace libconfig
#endif
