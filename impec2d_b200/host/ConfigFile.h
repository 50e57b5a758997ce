// ConfigFile.h -- reader for the libconfig subset the shipped examples/*.cfg use (libconfig++ itself is not a
// dependency of this build): groups `name: { ... };`, settings `key = value;`, values int / float / bool / "string",
// arrays `[a, b]`, comments `#`, `//`, `/* */`.  Lookup semantics follow libconfig's Setting::lookupValue:
// false when the key is missing or has the wrong type (ints are accepted where a double is asked for).
#ifndef CONFIGFILE_H_
#define CONFIGFILE_H_
#include <string>
#include <vector>
#include <map>
#include <stdexcept>

class CfgSetting {
public:
    enum Type { GROUP, INT, FLOAT, BOOL, STRING, LIST };
    Type type = GROUP;
    long long ival = 0; double fval = 0; bool bval = false; std::string sval;
    std::vector<CfgSetting> items;                       // LIST
    std::vector<std::pair<std::string, CfgSetting> > members;   // GROUP (file order)
    bool exists(const std::string &k) const;
    const CfgSetting &operator[](const std::string &k) const;   // throws std::runtime_error if missing
    const CfgSetting &operator[](const char *k) const { return (*this)[std::string(k)]; }
    const CfgSetting &operator[](int i) const { return items.at(i); }
    int getLength() const { return type == LIST ? (int)items.size() : (int)members.size(); }
    bool lookupValue(const std::string &k, int &v) const;
    bool lookupValue(const std::string &k, double &v) const;
    bool lookupValue(const std::string &k, bool &v) const;
    bool lookupValue(const std::string &k, std::string &v) const;
    operator int() const { return (int)ival; }
    operator double() const { return type == INT ? (double)ival : fval; }
};

class ConfigFile {
    CfgSetting root_;
public:
    void readFile(const std::string &path);              // throws std::runtime_error("file:line: message")
    void readString(const std::string &text, const std::string &name = "<string>");
    const CfgSetting &getRoot() const { return root_; }
};
#endif
