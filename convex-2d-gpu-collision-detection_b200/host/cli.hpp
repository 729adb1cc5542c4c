// cli.hpp -- command-line parser with the boost::program_options behaviour the reference relies on
// (boost headers are not available in this image; parse_args of generate_dataset.cu:66-169,
// ztest.cu:49-101, compute_collision_probability.cu:44-85, argparser.h:24-119):
//   --name value | --name=value | unambiguous long-option prefixes | short aliases (-n -b -s -w -h)
//   multitoken float lists run to the next option (a token starting with '-' followed by a non-digit)
//   bool values 1/0/true/false/on/off/yes/no; switches take no value
//   unknown / ambiguous option or missing value -> std::runtime_error (boost throws, the reference aborts)
#pragma once
#include <cerrno>
#include <climits>
#include <cstdlib>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace cli {

enum class Kind { Switch, String, Int, Float, Bool, FloatList };

struct Option {
    std::string name;
    char short_name;
    Kind kind;
    std::string help;
};

class Parser {
public:
    explicit Parser(std::string caption) : caption_(std::move(caption)) {}
    Parser& add(const std::string& name, Kind kind, const std::string& help, char short_name = 0) {
        opts_.push_back({name, short_name, kind, help});
        return *this;
    }

    void parse(int argc, char** argv) {
        for (int i = 1; i < argc; i++) {
            std::string tok = argv[i];
            const Option* opt = nullptr;
            std::string inline_val;
            bool has_inline = false;
            if (tok.rfind("--", 0) == 0) {
                std::string name = tok.substr(2);
                size_t eq = name.find('=');
                if (eq != std::string::npos) { inline_val = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
                opt = find_long(name);
            } else if (tok.size() >= 2 && tok[0] == '-' && !looks_numeric(tok)) {
                opt = find_short(tok[1]);
                if (tok.size() > 2) { inline_val = tok.substr(2); has_inline = true; }
            } else {
                throw std::runtime_error("too many positional options have been specified on the command line: " + tok);
            }
            std::vector<std::string>& vals = values_[opt->name];
            seen_[opt->name]++;
            if (opt->kind == Kind::Switch) {
                if (has_inline) throw std::runtime_error("option '--" + opt->name + "' does not take any arguments");
                continue;
            }
            if (has_inline) vals.push_back(inline_val);
            if (opt->kind == Kind::FloatList) {
                while (i + 1 < argc && !is_option(argv[i + 1])) vals.push_back(argv[++i]);
                if (vals.empty()) throw std::runtime_error("the required argument for option '--" + opt->name + "' is missing");
            } else if (!has_inline) {
                if (i + 1 >= argc || is_option(argv[i + 1]))
                    throw std::runtime_error("the required argument for option '--" + opt->name + "' is missing");
                vals.assign(1, argv[++i]);
            } else {
                vals.assign(1, inline_val);
            }
        }
    }

    int count(const std::string& name) const { auto it = seen_.find(name); return it == seen_.end() ? 0 : it->second; }
    std::string str(const std::string& name) const { return values_.at(name).back(); }
    int integer(const std::string& name) const {
        const std::string v = str(name);
        char* end = nullptr;
        errno = 0;
        long long x = std::strtoll(v.c_str(), &end, 10);
        if (end == v.c_str() || *end || errno == ERANGE || x < INT_MIN || x > INT_MAX)      // boost: bad_lexical_cast
            throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid");
        return (int)x;
    }
    // unsigned 64-bit value (--seed): the whole range is accepted, anything else is rejected rather than truncated
    unsigned long long unsigned64(const std::string& name) const {
        const std::string v = str(name);
        char* end = nullptr;
        errno = 0;
        if (v.empty() || v[0] == '-' || v[0] == '+' || v[0] == ' ')
            throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid");
        unsigned long long x = std::strtoull(v.c_str(), &end, 10);
        if (end == v.c_str() || *end || errno == ERANGE)
            throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid");
        return x;
    }
    float real(const std::string& name) const { return to_float(str(name), name); }
    bool boolean(const std::string& name) const {
        std::string v = str(name);
        for (char& c : v) c = (char)std::tolower((unsigned char)c);
        if (v == "1" || v == "true" || v == "on" || v == "yes") return true;
        if (v == "0" || v == "false" || v == "off" || v == "no") return false;
        throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid. Valid choices are 'on|off', 'yes|no', '1|0' and 'true|false'");
    }
    std::vector<float> reals(const std::string& name) const {
        std::vector<float> out;
        for (const std::string& v : values_.at(name)) out.push_back(to_float(v, name));
        return out;
    }

    void print_help(std::ostream& os) const {
        os << caption_ << ":\n";
        for (const Option& o : opts_) {
            std::string left = "  ";
            if (o.short_name) left += std::string("-") + o.short_name + " [ --" + o.name + " ]";
            else left += "--" + o.name;
            if (o.kind != Kind::Switch) left += " arg";
            if (left.size() < 34) left.append(34 - left.size(), ' '); else left += " ";
            os << left << o.help << "\n";
        }
    }

private:
    static bool looks_numeric(const std::string& t) {
        return t.size() >= 2 && t[0] == '-' && ((t[1] >= '0' && t[1] <= '9') || t[1] == '.');
    }
    static bool is_option(const std::string& t) { return t.size() >= 2 && t[0] == '-' && !looks_numeric(t); }
    static float to_float(const std::string& v, const std::string& name) {
        char* end = nullptr;
        float x = std::strtof(v.c_str(), &end);
        if (end == v.c_str() || *end) throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid");
        return x;
    }
    const Option* find_long(const std::string& name) const {
        const Option* hit = nullptr;
        int n = 0;
        for (const Option& o : opts_) {
            if (o.name == name) return &o;
            if (o.name.rfind(name, 0) == 0) { hit = &o; n++; }
        }
        if (n == 1) return hit;
        if (n > 1) throw std::runtime_error("option '--" + name + "' is ambiguous");
        throw std::runtime_error("unrecognised option '--" + name + "'");
    }
    const Option* find_short(char c) const {
        for (const Option& o : opts_) if (o.short_name == c) return &o;
        throw std::runtime_error(std::string("unrecognised option '-") + c + "'");
    }

    std::string caption_;
    std::vector<Option> opts_;
    std::map<std::string, std::vector<std::string>> values_;
    std::map<std::string, int> seen_;
};

}  // namespace cli
