// Configuration surface of wass_stereo: every key of the reference, the incfg
// file syntax and its error behaviour (ext/incfg/incfg.cpp:86-164, incfg.hpp:436-450), re-implemented
// as a plain registry (no static-initialisation macros).
#pragma once
#include <iosfwd>
#include <map>
#include <stdexcept>
#include <string>

namespace wasshost {

struct ConfigError : std::runtime_error { using std::runtime_error::runtime_error; };

class Config {
public:
    enum Type { INT, DOUBLE, BOOL, STRING };
    struct Option {
        bool is_def = true;      // incfg's sticky default flag, see Option::parse
        Type type; std::string desc;
        long long i = 0, i0 = 0; double d = 0, d0 = 0; bool b = false, b0 = false; std::string s, s0;
        bool is_default() const;
        std::string value_str() const;
        void parse(const std::string& v);
    };
    Config();                                   // registers all wass_stereo keys with the reference defaults
    void load(std::istream& is);                // throws ConfigError (unknown key, parse error)
    std::string to_config_string() const;       // generator format of incfg.hpp:436-450
    int geti(const char* k) const { return (int)opt(k).i; }
    double getd(const char* k) const { return opt(k).d; }
    bool getb(const char* k) const { return opt(k).b; }
    const std::string& gets(const char* k) const { return opt(k).s; }
    size_t size() const { return options_.size(); }
private:
    void addi(const char* k, int v, const char* d);
    void addd(const char* k, double v, const char* d);
    void addb(const char* k, bool v, const char* d);
    void adds(const char* k, const char* v, const char* d);
    const Option& opt(const char* k) const;
    std::map<std::string, Option> options_;     // std::map: keys come out sorted, like incfg's
};

}  // namespace wasshost
