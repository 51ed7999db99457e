// Host-side IR of a compiled ruleset: atoms (per-request predicates the GPU
// evaluates) and rules as pure boolean formulas over atoms.
//
// Mirrors the runtime objects of the reference:
//   pingoo::rules::Rule{name, expression, actions}      pingoo/rules.rs:9-14
//   rules::Action{Block, Captcha}                        rules/rules.rs:30-35
//   lists (name -> String/Int/Ip list)                   pingoo/lists.rs:11-15,115-125
//   variables http_request{..}, client{..}               pingoo/rules.rs:16-34
#pragma once
#include <bitset>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "expr.hpp"
#include "regex.hpp"

namespace pgw {

enum Field : int { F_HOST = 0, F_URL = 1, F_PATH = 2, F_METHOD = 3, F_USER_AGENT = 4, N_FIELDS = 5 };
extern const char* const kFieldNames[N_FIELDS];

enum IntFeat : int { IF_PORT = 0, IF_ASN = 1, IF_LEN0 = 2 /* IF_LEN0 + field */, N_INT_FEATS = 2 + N_FIELDS };

enum CmpOp : int { CMP_EQ = 0, CMP_NE, CMP_LT, CMP_LE, CMP_GT, CMP_GE };

enum ActionCode : uint8_t { ACT_BLOCK = 1, ACT_CAPTCHA = 2 };

struct IpNet {
    bool v6 = false;
    uint8_t addr[16] = {0};  // v4 in addr[0..4)
    int prefix = 0;
};

enum ListType : int { LT_STRING = 0, LT_INT = 1, LT_IP = 2 };

struct ListData {
    ListType type = LT_STRING;
    std::vector<std::string> strs;
    std::vector<int64_t> ints;
    std::vector<IpNet> nets;
};

// ---- boolean formula pool (hash-consed, constant-folded) --------------------
struct BoolNode {
    enum Kind : uint8_t { CONST, ATOM, NOT, AND, OR } kind;
    int a = -1, b = -1;  // children / atom index
    bool v = false;
};

class BoolPool {
  public:
    BoolPool();
    int constant(bool v) const { return v ? 1 : 0; }
    int atom(int idx);
    int mk_not(int x);
    int mk_and(int x, int y);
    int mk_or(int x, int y);
    // rebuild `root` with every ATOM(a) for which repl[a] >= 0 replaced by node repl[a]
    int substitute(int root, const std::vector<int>& repl);
    const BoolNode& at(int i) const { return nodes_[i]; }
    bool is_const(int i) const { return nodes_[i].kind == BoolNode::CONST; }
    bool const_value(int i) const { return nodes_[i].v; }
    size_t size() const { return nodes_.size(); }
    bool eval(int root, const std::vector<uint8_t>& atom_values) const;

  private:
    std::vector<BoolNode> nodes_;
    std::map<std::tuple<int, int, int>, int> index_;
    int intern(BoolNode::Kind k, int a, int b);
};

// ---- atoms --------------------------------------------------------------------
struct AtomDesc {
    enum Kind : uint8_t {
        STR_PATTERN,   // field matches a pattern (regex or anchored literal); scanned by the DFA kernel
        INT_CMP,       // int feature <op> constant
        INT_SET,       // int feature in a sorted constant set
        IP_SET,        // client.ip contained in any network of a set
        COUNTRY_SET,   // client.country in a 26x26 bitmap
        INT_EXPR,      // comparison of two integer expressions over request variables (prog), or "the expression errors" (op 6)
        FIELD_CMP      // one http_request field against another: op 0 ==, 1 starts_with, 2 ends_with, 3 contains, 4 <, 5 <=, 6 >, 7 >= (field, feat = second field)
    } kind;
    int field = -1;        // STR_PATTERN
    std::vector<int> nfa_starts;  // STR_PATTERN: start node(s) in Model::nfa[field]; pattern id of part k = event_base + k
    int event_base = -1;          // STR_PATTERN: first index into Model::events
    bool has_latch = false;       // STR_PATTERN: gap-split pattern (RegexParts) needing one latch bit in its scan unit
    int lit_kind = 0;             // STR_PATTERN from a literal anchored at the start: 1 starts_with(lit), 2 == lit
    std::string lit;
    int feat = -1;         // INT_*
    int op = 0;            // INT_CMP
    int64_t cval = 0;      // INT_CMP
    int set_id = -1;       // INT_SET / IP_SET / COUNTRY_SET; INT_EXPR: index into Model::int_progs
    int pos_refs = 0, neg_refs = 0;  // polarity statistics -> expected value heuristic
    std::string key;       // dedupe key / debug description
};

// What the DFA reports for NFA pattern id i (= index into Model::events)
struct PatternEvent {
    uint8_t kind;  // EventKind
    int atom;
};

struct RuleModel {
    std::string name;
    bool has_expression = false;
    int formula = 1;                 // BoolPool node: "expression evaluates to Bool(true)"
    bool is_service = false;         // a service route (HttpService::match_request), not a WAF rule
    std::vector<uint8_t> actions;    // ActionCode sequence, reference order
};

struct GeoRecord {
    uint32_t asn = 0;
    char country[2] = {'X', 'X'};
};

struct Model {
    BoolPool pool;
    std::vector<AtomDesc> atoms;
    std::vector<PatternEvent> events;
    std::unordered_map<std::string, int> atom_index;
    Nfa nfa[N_FIELDS];
    std::vector<std::vector<int64_t>> int_sets;
    // INT_EXPR programs (postfix, one int64 token each; program.hpp IntTok): two operand expressions, compared by the atom's op
    std::vector<std::vector<int64_t>> int_progs;
    std::vector<std::vector<IpNet>> ip_sets;
    std::vector<std::bitset<676>> country_sets;
    std::vector<RuleModel> rules;    // WAF rules first, then service routes
    uint32_t n_waf_rules = 0xFFFFFFFFu;  // unset: every rule is a WAF rule
    std::map<std::string, ListData> lists;
    std::vector<std::string> warnings;  // e.g. invalid regex folded to runtime error
};

struct LowerError {
    std::string msg;
};

// Lower one parsed expression into `model` (adding atoms) and return the
// BoolPool node for "rule matches".  Throws LowerError for constructs that are
// valid in the reference language but not implemented by this engine.
int lower_rule_expression(Model& model, const Expr& e, const std::string& rule_name);

// host-side regex evaluation used for constant folding
bool host_regex_is_match(const std::string& pattern, const std::string& hay, RegexStatus* st, std::string& err);

}  // namespace pgw
