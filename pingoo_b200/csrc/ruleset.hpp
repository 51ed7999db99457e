// Host-side ruleset assembly shared by the CUDA C-ABI (capi.cu) and the
// test-only program simulator: parse rules, collect lists / GeoIP, compile.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "compile.hpp"

namespace pgw {

struct RuleSource {
    std::string name;
    bool has_expression = false;
    std::string expression;
    std::vector<uint8_t> actions;
    ExprP ast;
};

class RulesetBuilder {
  public:
    CompileOptions options;

    // rules::compile_expression (rules/rules.rs:45-53): syntax check only
    static bool compile_expression(const std::string& src, std::string& err, ExprP* ast = nullptr);
    // rules::validate_expression (rules/rules.rs:55-77)
    static bool validate_expression(const std::string& src, std::string& err);

    // config.rs:255-269: a rule whose expression does not compile is a fatal configuration error
    bool add_rule(const char* name, const char* expression, const uint8_t* actions, uint32_t n_actions, std::string& err);
    // config_file.rs:257-265: a service's `route` is compiled like a rule expression (a failure is a fatal config error);
    // services are tried in order (http_listener.rs:266-270)
    bool add_service(const char* name, const char* route, std::string& err);
    bool add_list(const char* name, int type, const uint8_t* csv, size_t len, std::string& err);
    bool load_geoip(const uint8_t* mmdb, size_t len, std::string& err);
    bool finalize(HostProgram* out, std::string& err);

    size_t n_rules() const { return rules_.size(); }
    size_t n_services() const { return services_.size(); }

  private:
    std::vector<RuleSource> rules_;
    std::vector<RuleSource> services_;
    Model model_;
    std::vector<uint8_t> mmdb_;
    bool finalized_ = false;
};

// config_dir.cpp: fill `builder` from a Pingoo configuration directory (rules, the services an HTTP listener offers,
// lists, GeoIP database).  `listener` null: the default service set of an http / https listener.
bool load_config_dir(const std::string& folder, const char* listener, const std::vector<std::string>& geoip_dirs, RulesetBuilder* builder,
                     std::string* geoip_path, std::string& err);

}  // namespace pgw
