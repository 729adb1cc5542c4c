// Build-harness stub (oracle/_ref only), compile-only: lets the reference's parse_args
// (ztest.cu:65-101) type-check.  The oracle harness never calls the reference's main() or
// parse_args(), so nothing here parses anything.
#pragma once
#include <map>
#include <ostream>
#include <string>
#include <vector>
namespace boost { namespace program_options {
struct value_semantic { value_semantic* multitoken() { return this; } };
template <class T> inline value_semantic* value() { static value_semantic v; return &v; }
struct options_description;
struct options_adder {
    options_adder& operator()(const char*, const char*) { return *this; }
    options_adder& operator()(const char*, value_semantic*, const char*) { return *this; }
};
struct options_description {
    explicit options_description(const char*) {}
    options_adder add_options() { return options_adder(); }
};
inline std::ostream& operator<<(std::ostream& os, const options_description&) { return os; }
struct variable_value { template <class T> T as() const { return T(); } };
struct variables_map {
    int count(const std::string&) const { return 0; }
    variable_value operator[](const std::string&) const { return variable_value(); }
};
struct parsed_options {};
inline parsed_options parse_command_line(int, char**, const options_description&) { return parsed_options(); }
inline void store(const parsed_options&, variables_map&) {}
inline void notify(variables_map&) {}
}}
