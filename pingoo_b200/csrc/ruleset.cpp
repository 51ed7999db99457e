#include "ruleset.hpp"

#include <algorithm>

namespace pgw {

bool RulesetBuilder::compile_expression(const std::string& src, std::string& err, ExprP* ast) {
    std::string perr;
    ExprP e = parse_expression(src, perr);
    if (!e) {
        err = "Expression is not valid: " + perr;  // rules::Error::ExpressionIsNotValid
        return false;
    }
    if (ast) *ast = std::move(e);
    return true;
}

bool RulesetBuilder::validate_expression(const std::string& src, std::string& err) {
    if (src.empty()) {
        err = "Expression is not valid: expression is empty";
        return false;
    }
    ExprP ast;
    if (!compile_expression(src, err, &ast)) return false;
    std::vector<std::string> fns;
    collect_functions(*ast, fns);
    if (std::find(fns.begin(), fns.end(), "@in") != fns.end()) {
        err = "Expression is not valid: unknown operator: in";
        return false;
    }
    return true;
}

bool RulesetBuilder::add_rule(const char* name, const char* expression, const uint8_t* actions, uint32_t n_actions, std::string& err) {
    if (finalized_) { err = "ruleset already finalized"; return false; }
    RuleSource r;
    r.name = name ? name : "";
    if (n_actions && !actions) { err = "error parsing rules: rule '" + r.name + "': actions is null"; return false; }
    for (uint32_t i = 0; i < n_actions; ++i) {
        if (actions[i] != ACT_BLOCK && actions[i] != ACT_CAPTCHA) {
            err = "error parsing rules: rule '" + r.name + "': unknown action code " + std::to_string((int)actions[i]);
            return false;
        }
        r.actions.push_back(actions[i]);
    }
    if (expression) {
        r.has_expression = true;
        r.expression = expression;
        std::string cerr;
        if (!compile_expression(r.expression, cerr, &r.ast)) {
            err = "error parsing rules: " + cerr;  // config.rs:268
            return false;
        }
    }
    rules_.push_back(std::move(r));
    return true;
}

bool RulesetBuilder::add_service(const char* name, const char* route, std::string& err) {
    if (finalized_) { err = "ruleset already finalized"; return false; }
    RuleSource r;
    r.name = name ? name : "";
    if (route) {
        r.has_expression = true;
        r.expression = route;
        std::string cerr;
        if (!compile_expression(r.expression, cerr, &r.ast)) {
            err = "error parsing route for service " + r.name + ": " + cerr;
            return false;
        }
    }
    services_.push_back(std::move(r));
    return true;
}

bool RulesetBuilder::add_list(const char* name, int type, const uint8_t* csv, size_t len, std::string& err) {
    if (finalized_) { err = "ruleset already finalized"; return false; }
    if (type < 0 || type > 2) { err = std::to_string(type) + " is not a valid ListType"; return false; }
    ListData L;
    if (!parse_list_csv(name, (ListType)type, csv, len, &L, err)) return false;
    model_.lists[name] = std::move(L);
    return true;
}

bool RulesetBuilder::load_geoip(const uint8_t* mmdb, size_t len, std::string& err) {
    if (finalized_) { err = "ruleset already finalized"; return false; }
    if (len == 0) { err = "mmdb file is not valid: empty"; return false; }
    mmdb_.assign(mmdb, mmdb + len);
    return true;
}

bool RulesetBuilder::finalize(HostProgram* out, std::string& err) {
    if (finalized_) { err = "ruleset already finalized"; return false; }
    try {
        for (auto& r : rules_) {
            RuleModel rm;
            rm.name = r.name;
            rm.has_expression = r.has_expression;
            rm.actions = r.actions;
            // pingoo/rules.rs:36-52: no expression => the rule matches every request
            rm.formula = r.has_expression ? lower_rule_expression(model_, *r.ast, r.name) : model_.pool.constant(true);
            model_.rules.push_back(std::move(rm));
        }
        // service routes ride in the same program after the WAF rules: same atoms, same scan, no actions
        model_.n_waf_rules = (uint32_t)model_.rules.size();
        for (auto& r : services_) {
            RuleModel rm;
            rm.name = r.name;
            rm.has_expression = r.has_expression;
            rm.is_service = true;
            // http_proxy_service.rs:84-95: no route => the service matches every request
            rm.formula = r.has_expression ? lower_rule_expression(model_, *r.ast, r.name) : model_.pool.constant(true);
            model_.rules.push_back(std::move(rm));
        }
    } catch (const LowerError& e) {
        err = e.msg;
        return false;
    }
    if (!compile_program(model_, options, mmdb_, out, err)) return false;
    finalized_ = true;
    return true;
}

}  // namespace pgw
