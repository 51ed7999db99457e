// Typed symbolic evaluation of a rule expression -> atoms + boolean formula.
//
// Restates, at compile time, what `bel::Program::execute` does per request
// (reference call site pingoo/rules.rs:36-52): every sub-expression is either a
// constant, a request variable, or a boolean that is a pure function of atoms.
// Runtime errors (SEMANTICS.md A4-A6: unknown variable/function/list key,
// cross-type comparison, non-bool operand) are compile-time-known here, so the
// three-valued {false,true,error} logic is folded into two boolean formulas
// (is_true, is_error) and only is_true(root) is shipped to the GPU.
#include <algorithm>
#include <cstring>
#include <functional>
#include <tuple>

#include "dfa.hpp"
#include "model.hpp"
#include "program.hpp"

namespace pgw {

const char* const kFieldNames[N_FIELDS] = {"host", "url", "path", "method", "user_agent"};

// ---- BoolPool -------------------------------------------------------------------
BoolPool::BoolPool() {
    BoolNode f; f.kind = BoolNode::CONST; f.v = false;
    BoolNode t; t.kind = BoolNode::CONST; t.v = true;
    nodes_.push_back(f);
    nodes_.push_back(t);
}
int BoolPool::intern(BoolNode::Kind k, int a, int b) {
    auto key = std::make_tuple((int)k, a, b);
    auto it = index_.find(key);
    if (it != index_.end()) return it->second;
    BoolNode n; n.kind = k; n.a = a; n.b = b;
    nodes_.push_back(n);
    index_[key] = (int)nodes_.size() - 1;
    return (int)nodes_.size() - 1;
}
int BoolPool::atom(int idx) { return intern(BoolNode::ATOM, idx, -1); }
int BoolPool::mk_not(int x) {
    if (is_const(x)) return constant(!const_value(x));
    if (nodes_[x].kind == BoolNode::NOT) return nodes_[x].a;
    return intern(BoolNode::NOT, x, -1);
}
int BoolPool::mk_and(int x, int y) {
    if (is_const(x)) return const_value(x) ? y : 0;
    if (is_const(y)) return const_value(y) ? x : 0;
    if (x == y) return x;
    if (mk_not(x) == y) return 0;
    if (x > y) std::swap(x, y);
    return intern(BoolNode::AND, x, y);
}
int BoolPool::mk_or(int x, int y) {
    if (is_const(x)) return const_value(x) ? 1 : y;
    if (is_const(y)) return const_value(y) ? 1 : x;
    if (x == y) return x;
    if (mk_not(x) == y) return 1;
    if (x > y) std::swap(x, y);
    return intern(BoolNode::OR, x, y);
}
int BoolPool::substitute(int root, const std::vector<int>& repl) {
    const BoolNode n = nodes_[root];
    switch (n.kind) {
        case BoolNode::CONST: return root;
        case BoolNode::ATOM: return (n.a < (int)repl.size() && repl[n.a] >= 0) ? repl[n.a] : root;
        case BoolNode::NOT: return mk_not(substitute(n.a, repl));
        case BoolNode::AND: return mk_and(substitute(n.a, repl), substitute(n.b, repl));
        case BoolNode::OR: return mk_or(substitute(n.a, repl), substitute(n.b, repl));
    }
    return root;
}
bool BoolPool::eval(int root, const std::vector<uint8_t>& av) const {
    const BoolNode& n = nodes_[root];
    switch (n.kind) {
        case BoolNode::CONST: return n.v;
        case BoolNode::ATOM: return av[n.a] != 0;
        case BoolNode::NOT: return !eval(n.a, av);
        case BoolNode::AND: return eval(n.a, av) && eval(n.b, av);
        case BoolNode::OR: return eval(n.a, av) || eval(n.b, av);
    }
    return false;
}

bool host_regex_is_match(const std::string& pattern, const std::string& hay, RegexStatus* st, std::string& err) {
    Nfa nfa;
    RegexParts parts;
    RegexInfo info;
    *st = regex_compile(pattern, 0, nfa, &parts, &info, err, /*allow_split=*/false);
    if (*st != RX_OK) return false;
    if (info.always_true) return true;
    Dfa d;
    if (!build_dfa(nfa, {parts.start[0]}, 1 << 20, &d)) {
        *st = RX_TOO_BIG;
        err = "pattern too complex for constant folding";
        return false;
    }
    int s = d.start;
    bool hit = false;
    for (unsigned char c : hay) {
        s = d.trans[(size_t)s * d.n_classes + d.classmap[c]];
        if (s >= d.acc_lo) hit = true;
    }
    return hit || !d.endacc[s].empty();
}

namespace {

struct BV {
    int t = 0, e = 0;
};

struct Sym {
    enum K : uint8_t {
        ERR, NUL, BOOL, INT_C, UINT_C, FLOAT_C, STR_C, BYTES_C, LIST_C, LIST_REF, MAP_C,
        MAP_LISTS, MAP_HTTP, MAP_CLIENT, STR_FIELD, INT_FEAT, INT_EXPR, IP_VAR, COUNTRY_VAR,
        CHOICE,  // c ? x : y with non-boolean branches on a request-dependent condition: bv = the condition, items = {x, y}
        CONCAT   // a + b + ... on strings with at least one http_request field among them: items = the parts (STR_C / STR_FIELD, no
                 // empty and no two adjacent constants)
    } k = ERR;
    std::vector<int64_t> prog;   // INT_EXPR: postfix tokens (program.hpp IntTok)
    BV bv;
    int64_t i = 0;
    double f = 0;
    std::string s;
    std::vector<Sym> items;
    const ListData* list = nullptr;
    std::string list_name;
    int field = -1;
    int feat = -1;
};

struct Lowerer {
    Model& M;
    BoolPool& P;
    std::string rule;
    std::map<std::string, int> ipset_of_list;

    explicit Lowerer(Model& m, const std::string& r) : M(m), P(m.pool), rule(r) {}

    [[noreturn]] void unsupported(const Expr& e, const std::string& what) {
        throw LowerError{"rule '" + rule + "': " + what + " (offset " + std::to_string(e.pos) +
                         ") is valid in the rule language but not supported by the GPU engine"};
    }

    static Sym err() { return Sym(); }
    Sym boolean(int t, int e = 0) {
        Sym s; s.k = Sym::BOOL; s.bv.t = t; s.bv.e = e;
        return s;
    }
    Sym const_bool(bool v) { return boolean(P.constant(v)); }
    static Sym const_int(int64_t v) { Sym s; s.k = Sym::INT_C; s.i = v; return s; }
    static Sym const_str(const std::string& v) { Sym s; s.k = Sym::STR_C; s.s = v; return s; }

    bool is_const(const Sym& s) {
        switch (s.k) {
            case Sym::NUL: case Sym::INT_C: case Sym::UINT_C: case Sym::FLOAT_C: case Sym::STR_C: case Sym::BYTES_C:
            case Sym::LIST_REF: return true;
            case Sym::BOOL: return P.is_const(s.bv.t) && P.is_const(s.bv.e);
            case Sym::LIST_C: case Sym::MAP_C:
                for (const auto& it : s.items) if (!is_const(it)) return false;
                return true;
            default: return false;
        }
    }

    // as a boolean operand of ! && || ?: : non-bool -> error value
    BV as_bool(const Sym& s) {
        if (s.k == Sym::BOOL) return s.bv;
        BV r; r.t = 0; r.e = 1;
        return r;
    }

    // ---- atoms -----------------------------------------------------------------
    int add_atom(AtomDesc&& a) {
        auto it = M.atom_index.find(a.key);
        if (it != M.atom_index.end()) return it->second;
        int id = (int)M.atoms.size();
        M.atom_index[a.key] = id;
        M.atoms.push_back(std::move(a));
        return id;
    }

    static std::string lenpfx(const std::string& s) { return std::to_string(s.size()) + ":" + s; }

    Sym str_literal_atom(int field, const std::string& lit, bool a_start, bool a_end) {
        if (lit.empty() && !(a_start && a_end)) return const_bool(true);
        std::string key = std::string("S|") + std::to_string(field) + "|" + (a_start ? "^" : "") + (a_end ? "$" : "") + "|" + lenpfx(lit);
        auto it = M.atom_index.find(key);
        if (it != M.atom_index.end()) return boolean(P.atom(it->second));
        AtomDesc a;
        a.kind = AtomDesc::STR_PATTERN;
        a.field = field;
        a.key = key;
        int id = (int)M.atoms.size();
        a.event_base = (int)M.events.size();
        a.nfa_starts.push_back(nfa_literal(M.nfa[field], lit, a_start, a_end, a.event_base));
        M.events.push_back(PatternEvent{EV_FIRE, id});
        if (a_start) { a.lit_kind = a_end ? 2 : 1; a.lit = lit; }
        return boolean(P.atom(add_atom(std::move(a))));
    }

    Sym str_regex_atom(const Expr& e, int field, const std::string& pattern) {
        std::string key = std::string("S|") + std::to_string(field) + "|re|" + lenpfx(pattern);
        auto it = M.atom_index.find(key);
        if (it != M.atom_index.end()) return boolean(P.atom(it->second));
        AtomDesc a;
        a.kind = AtomDesc::STR_PATTERN;
        a.field = field;
        a.key = key;
        int id = (int)M.atoms.size();
        RegexInfo info;
        RegexParts parts;
        std::string msg;
        a.event_base = (int)M.events.size();
        RegexStatus st = regex_compile(pattern, a.event_base, M.nfa[field], &parts, &info, msg);
        if (st == RX_UNSUPPORTED) unsupported(e, "regex feature: " + msg);
        if (st != RX_OK) {
            M.warnings.push_back("rule '" + rule + "': regex does not compile (" + msg + "); evaluation is a runtime error -> no match");
            return err();
        }
        if (info.always_true) return const_bool(true);
        for (int k = 0; k < parts.n; ++k) {
            a.nfa_starts.push_back(parts.start[k]);
            M.events.push_back(PatternEvent{parts.kind[k], id});
        }
        a.has_latch = parts.n > 1;
        return boolean(P.atom(add_atom(std::move(a))));
    }

    Sym str_set_atom(const Expr& e, int field, const std::vector<std::string>& strs) {
        if (strs.empty()) return const_bool(false);
        std::vector<std::string> u(strs);
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        size_t total = 0;
        for (auto& s : u) total += s.size();
        if (total > 200000) unsupported(e, "String list membership on an http_request field with more than 200000 bytes of entries");
        std::string key = std::string("S|") + std::to_string(field) + "|set|";
        for (auto& s : u) key += lenpfx(s) + ",";
        auto it = M.atom_index.find(key);
        if (it != M.atom_index.end()) return boolean(P.atom(it->second));
        AtomDesc a;
        a.kind = AtomDesc::STR_PATTERN;
        a.field = field;
        a.key = key;
        int id = (int)M.atoms.size();
        Nfa& nfa = M.nfa[field];
        auto add = [&](NfaKind k) { NfaNode n; n.kind = k; nfa.nodes.push_back(n); return (int)nfa.nodes.size() - 1; };
        a.event_base = (int)M.events.size();
        M.events.push_back(PatternEvent{EV_FIRE, id});
        int m = add(N_MATCH);
        nfa.nodes[m].pattern = a.event_base;
        int eol = add(N_ASSERT);
        nfa.nodes[eol].assert_kind = A_EOL_TEXT;
        nfa.nodes[eol].out = m;
        // trie over the (sorted) strings keeps the NFA and its closure small
        struct TrieNode { std::map<unsigned char, int> next; bool end = false; };
        std::vector<TrieNode> trie(1);
        for (auto& s : u) {
            int cur = 0;
            for (unsigned char c : s) {
                auto f = trie[cur].next.find(c);
                if (f == trie[cur].next.end()) {
                    trie.emplace_back();
                    int nn = (int)trie.size() - 1;
                    trie[cur].next[c] = nn;
                    cur = nn;
                } else cur = f->second;
            }
            trie[cur].end = true;
        }
        std::function<int(int)> emit = [&](int t) -> int {
            // returns NFA entry for trie node t
            std::vector<int> alts;
            if (trie[t].end) alts.push_back(eol);
            for (auto& kv : trie[t].next) {
                int child = emit(kv.second);
                ByteSet bs;
                bs.set(kv.first);
                int c = add(N_CHAR);
                nfa.nodes[c].set = nfa.add_set(bs);
                nfa.nodes[c].out = child;
                alts.push_back(c);
            }
            int entry = alts.back();
            for (size_t k = alts.size() - 1; k-- > 0;) {
                int sp = add(N_SPLIT);
                nfa.nodes[sp].out = alts[k];
                nfa.nodes[sp].out1 = entry;
                entry = sp;
            }
            return entry;
        };
        int body = emit(0);
        int bol = add(N_ASSERT);
        nfa.nodes[bol].assert_kind = A_BOL_TEXT;
        nfa.nodes[bol].out = body;
        a.nfa_starts.push_back(bol);
        return boolean(P.atom(add_atom(std::move(a))));
    }

    Sym int_cmp_atom(int feat, int op, int64_t c) {
        if (op == CMP_NE) {
            Sym eq = int_cmp_atom(feat, CMP_EQ, c);
            return boolean(P.mk_not(eq.bv.t));
        }
        AtomDesc a;
        a.kind = AtomDesc::INT_CMP;
        a.feat = feat;
        a.op = op;
        a.cval = c;
        a.key = "I|" + std::to_string(feat) + "|" + std::to_string(op) + "|" + std::to_string(c);
        return boolean(P.atom(add_atom(std::move(a))));
    }

    Sym int_set_atom(int feat, std::vector<int64_t> v) {
        if (v.empty()) return const_bool(false);
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        int sid = -1;
        for (size_t k = 0; k < M.int_sets.size(); ++k)
            if (M.int_sets[k] == v) sid = (int)k;
        if (sid < 0) { M.int_sets.push_back(v); sid = (int)M.int_sets.size() - 1; }
        AtomDesc a;
        a.kind = AtomDesc::INT_SET;
        a.feat = feat;
        a.set_id = sid;
        a.key = "IS|" + std::to_string(feat) + "|" + std::to_string(sid);
        return boolean(P.atom(add_atom(std::move(a))));
    }

    // ---- integer expressions over request variables (client.remote_port + 1, path.length() * 2, asn % 10 ...) --------
    static bool is_int_value(const Sym& s) { return s.k == Sym::INT_C || s.k == Sym::INT_FEAT || s.k == Sym::INT_EXPR; }
    static int64_t tok(uint32_t op, int64_t v = 0) { return (int64_t)(((uint64_t)op << 56) | ((uint64_t)v & 0x00FFFFFFFFFFFFFFull)); }
    static void append_int(std::vector<int64_t>& out, const Sym& s) {
        if (s.k == Sym::INT_C) {
            if (s.i >= -(1ll << 55) && s.i < (1ll << 55)) out.push_back(tok(IT_CONST, s.i));
            else { out.push_back(tok(IT_CONST64)); out.push_back(s.i); }
        } else if (s.k == Sym::INT_FEAT) out.push_back(tok(IT_FEAT, s.feat));
        else out.insert(out.end(), s.prog.begin(), s.prog.end());
    }
    static int prog_depth(const std::vector<int64_t>& p) {
        int d = 0, mx = 0;
        for (size_t i = 0; i < p.size(); ++i) {
            const uint32_t op = (uint32_t)((uint64_t)p[i] >> 56);
            if (op == IT_CONST || op == IT_FEAT) ++d;
            else if (op == IT_CONST64) { ++d; ++i; }
            else if (op == IT_NEG) {}
            else --d;
            mx = std::max(mx, d);
        }
        return mx;
    }
    static bool prog_can_error(const std::vector<int64_t>& p) {
        for (size_t i = 0; i < p.size(); ++i) {
            const uint32_t op = (uint32_t)((uint64_t)p[i] >> 56);
            if (op == IT_CONST64) { ++i; continue; }
            if (op >= IT_ADD) return true;
        }
        return false;
    }
    Sym int_arith(const Expr& e, uint32_t op, const Sym& a, const Sym& b) {
        Sym r;
        r.k = Sym::INT_EXPR;
        append_int(r.prog, a);
        append_int(r.prog, b);
        r.prog.push_back(tok(op));
        if (prog_depth(r.prog) > (int)kIntExprStack) unsupported(e, "integer expression nested deeper than the evaluator's operand stack");
        return r;
    }
    // (a <op> b) on two integer values at least one of which depends on the request: a true-atom and, if the arithmetic can
    // fail, an error-atom on the same program
    Sym int_expr_cmp(const Expr& e, int op, const Sym& a, const Sym& b) {
        std::vector<int64_t> prog;
        append_int(prog, a);
        append_int(prog, b);
        prog.push_back(tok(IT_END));
        if (prog_depth(prog) > (int)kIntExprStack) unsupported(e, "integer expression nested deeper than the evaluator's operand stack");
        int pid = -1;
        for (size_t k = 0; k < M.int_progs.size(); ++k)
            if (M.int_progs[k] == prog) pid = (int)k;
        if (pid < 0) { M.int_progs.push_back(prog); pid = (int)M.int_progs.size() - 1; }
        auto mk = [&](int o) {
            AtomDesc d;
            d.kind = AtomDesc::INT_EXPR;
            d.op = o;
            d.set_id = pid;
            d.key = "IX|" + std::to_string(pid) + "|" + std::to_string(o);
            return P.atom(add_atom(std::move(d)));
        };
        const int er = prog_can_error(prog) ? mk((int)kIntExprIsError) : 0;
        int t;
        if (op == CMP_NE) t = P.mk_and(P.mk_not(mk(CMP_EQ)), P.mk_not(er));   // the "==" atom is false on an error as well
        else t = mk(op);
        return boolean(t, er);
    }

    // one http_request field against another: 0 ==, 1 starts_with, 2 ends_with, 3 contains, 4 <, 5 <=, 6 >, 7 >= (byte-wise order)
    Sym field_cmp_atom(int f1, int f2, int op) {
        AtomDesc d;
        d.kind = AtomDesc::FIELD_CMP;
        d.field = f1;
        d.feat = f2;
        d.op = op;
        d.key = "FC|" + std::to_string(f1) + "|" + std::to_string(f2) + "|" + std::to_string(op);
        return boolean(P.atom(add_atom(std::move(d))));
    }

    // http_request.<field> <op> "literal" for the ordering operators: byte-wise lexicographic order (Rust str::cmp), a regular
    // language anchored at the start of the field: after the common prefix lit[0..i) either the field ends (less), or its next
    // byte is smaller / greater than lit[i]
    Sym str_order_atom(int field, const std::string& lit, int op) {
        if (lit.empty()) {   // against "": only the length matters
            if (op == CMP_GE) return const_bool(true);
            if (op == CMP_LT) return const_bool(false);
            return int_cmp_atom(IF_LEN0 + field, op == CMP_GT ? CMP_GT : CMP_EQ, 0);
        }
        std::string key = std::string("S|") + std::to_string(field) + "|ord" + std::to_string(op) + "|" + lenpfx(lit);
        auto it = M.atom_index.find(key);
        if (it != M.atom_index.end()) return boolean(P.atom(it->second));
        AtomDesc a;
        a.kind = AtomDesc::STR_PATTERN;
        a.field = field;
        a.key = key;
        const int id = (int)M.atoms.size();
        a.event_base = (int)M.events.size();
        Nfa& nfa = M.nfa[field];
        auto add = [&](NfaKind k) { NfaNode n; n.kind = k; nfa.nodes.push_back(n); return (int)nfa.nodes.size() - 1; };
        const int m = add(N_MATCH);
        nfa.nodes[m].pattern = a.event_base;
        const bool less = op == CMP_LT || op == CMP_LE, or_equal = op == CMP_LE || op == CMP_GE;
        auto eol_to_match = [&]() { int x = add(N_ASSERT); nfa.nodes[x].assert_kind = A_EOL_TEXT; nfa.nodes[x].out = m; return x; };
        auto alt = [&](int x, int y) { if (x < 0) return y; if (y < 0) return x; int sp = add(N_SPLIT); nfa.nodes[sp].out = x; nfa.nodes[sp].out1 = y; return sp; };
        // node for "the first |lit| bytes all matched": == decides <= / >=, a longer field is greater
        int next = -1;
        if (or_equal) next = less ? eol_to_match() : m;   // >=: anything from here on (end or more bytes) is >= lit
        else if (!less) {                                  // >: at least one more byte
            ByteSet any;
            any.negate();
            next = add(N_CHAR);
            nfa.nodes[next].set = nfa.add_set(any);
            nfa.nodes[next].out = m;
        }
        for (size_t k = lit.size(); k-- > 0;) {
            const unsigned c = (unsigned char)lit[k];
            int here = -1;
            if (less) here = eol_to_match();   // the field ends inside the literal: a proper prefix is smaller
            ByteSet diff;
            for (unsigned v = 0; v < 256; ++v)
                if (less ? v < c : v > c) diff.set(v);
            if (!diff.empty()) {
                int d = add(N_CHAR);
                nfa.nodes[d].set = nfa.add_set(diff);
                nfa.nodes[d].out = m;
                here = alt(here, d);
            }
            if (next >= 0) {
                ByteSet same;
                same.set(c);
                int sm = add(N_CHAR);
                nfa.nodes[sm].set = nfa.add_set(same);
                nfa.nodes[sm].out = next;
                here = alt(here, sm);
            }
            next = here;
        }
        if (next < 0) return const_bool(false);   // e.g. field < "": nothing is smaller than the empty string
        const int bol = add(N_ASSERT);
        nfa.nodes[bol].assert_kind = A_BOL_TEXT;
        nfa.nodes[bol].out = next;
        a.nfa_starts.push_back(bol);
        M.events.push_back(PatternEvent{EV_FIRE, id});
        return boolean(P.atom(add_atom(std::move(a))));
    }

    Sym ip_set_atom(const Sym& list) {
        // list is LIST_REF of type Ip
        if (list.list->nets.empty()) return const_bool(false);
        int sid;
        auto it = ipset_of_list.find(list.list_name);
        if (it != ipset_of_list.end()) sid = it->second;
        else {
            // the model may already hold this list's set from a previous rule
            sid = -1;
            for (size_t k = 0; k < M.atoms.size(); ++k)
                if (M.atoms[k].kind == AtomDesc::IP_SET && M.atoms[k].key == "IP|" + list.list_name) sid = M.atoms[k].set_id;
            if (sid < 0) { M.ip_sets.push_back(list.list->nets); sid = (int)M.ip_sets.size() - 1; }
            ipset_of_list[list.list_name] = sid;
        }
        AtomDesc a;
        a.kind = AtomDesc::IP_SET;
        a.set_id = sid;
        a.key = "IP|" + list.list_name;
        return boolean(P.atom(add_atom(std::move(a))));
    }

    Sym country_atom(const std::function<bool(const std::string&)>& pred) {
        std::bitset<676> bs;
        for (int a = 0; a < 26; ++a)
            for (int b = 0; b < 26; ++b) {
                std::string code;
                code.push_back((char)('A' + a));
                code.push_back((char)('A' + b));
                if (pred(code)) bs.set(a * 26 + b);
            }
        if (bs.none()) return const_bool(false);
        if (bs.all()) return const_bool(true);
        int sid = -1;
        for (size_t k = 0; k < M.country_sets.size(); ++k)
            if (M.country_sets[k] == bs) sid = (int)k;
        if (sid < 0) { M.country_sets.push_back(bs); sid = (int)M.country_sets.size() - 1; }
        AtomDesc a;
        a.kind = AtomDesc::COUNTRY_SET;
        a.set_id = sid;
        a.key = "C|" + std::to_string(sid);
        return boolean(P.atom(add_atom(std::move(a))));
    }

    // ---- constants ---------------------------------------------------------------
    // equality of two constants: 1 equal, 0 different, -1 cross-type (error for ==, "different" inside contains)
    int const_equal(const Sym& a, const Sym& b) {
        if (a.k != b.k) return -1;
        switch (a.k) {
            case Sym::NUL: return 1;
            case Sym::BOOL: return P.const_value(a.bv.t) == P.const_value(b.bv.t);
            case Sym::INT_C: case Sym::UINT_C: return a.i == b.i;
            case Sym::FLOAT_C: return a.f == b.f;
            case Sym::STR_C: case Sym::BYTES_C: return a.s == b.s;
            case Sym::LIST_C: {
                if (a.items.size() != b.items.size()) return 0;
                for (size_t k = 0; k < a.items.size(); ++k) {
                    int r = const_equal(a.items[k], b.items[k]);
                    if (r != 1) return r < 0 ? -1 : 0;
                }
                return 1;
            }
            default: return -1;
        }
    }

    static size_t utf8_len(const std::string& s) {
        size_t n = 0;
        for (unsigned char c : s) if ((c & 0xC0) != 0x80) ++n;
        return n;
    }

    // ---- expression walk ---------------------------------------------------------
    Sym lower(const Expr& e) {
        switch (e.kind) {
            case Expr::LIT_NULL: { Sym s; s.k = Sym::NUL; return s; }
            case Expr::LIT_BOOL: return const_bool(e.bval);
            case Expr::LIT_INT: return const_int(e.ival);
            case Expr::LIT_UINT: { Sym s; s.k = Sym::UINT_C; s.i = e.ival; return s; }
            case Expr::LIT_FLOAT: { Sym s; s.k = Sym::FLOAT_C; s.f = e.fval; return s; }
            case Expr::LIT_STR: return const_str(e.name);
            case Expr::LIT_BYTES: { Sym s; s.k = Sym::BYTES_C; s.s = e.name; return s; }
            case Expr::IDENT: {
                Sym s;
                if (e.name == "http_request") s.k = Sym::MAP_HTTP;
                else if (e.name == "client") s.k = Sym::MAP_CLIENT;
                else if (e.name == "lists") s.k = Sym::MAP_LISTS;
                else return err();  // undeclared reference -> runtime error
                return s;
            }
            case Expr::MEMBER: {
                Sym b0 = lower(*e.kids[0]);
                return lift1(b0, [&](const Sym& b) { return member(b, e.name); });
            }
            case Expr::INDEX: {
                Sym base0 = lower(*e.kids[0]);
                Sym ix0 = lower(*e.kids[1]);
                return lift2(base0, ix0, [&](const Sym& base, const Sym& ix) { return index_on(e, base, ix); });
            }
            case Expr::CALL: {
                for (auto& k : e.kids) (void)lower(*k);  // surface unsupported constructs in arguments
                return err();  // no global functions in the documented language (docs/rules.md:71-76)
            }
            case Expr::METHOD: return method(e);
            case Expr::UNARY: {
                Sym x0 = lower(*e.kids[0]);
                if (x0.k == Sym::CHOICE) return lift1(x0, [&](const Sym& v) { return unary_on(e, v); });
                return unary_on(e, x0);
            }
            case Expr::BINARY: return binary(e);
            case Expr::TERNARY: {
                Sym c = lower(*e.kids[0]);
                Sym x = lower(*e.kids[1]);
                Sym y = lower(*e.kids[2]);
                return lift1(c, [&](const Sym& cc) { return select(as_bool(cc), x, y); });
            }
            case Expr::LIST: {
                Sym l;
                l.k = Sym::LIST_C;
                for (auto& k : e.kids) {
                    Sym it = lower(*k);
                    if (it.k == Sym::ERR) return err();
                    // request variables may be elements: [http_request.method, "x"].contains("GET"), [..][0], .length() work on such a
                    // list; what cannot (whole-list comparison) says so where it is attempted
                    if (it.k == Sym::CHOICE) unsupported(e, "conditional (?:) value as a list element");
                    l.items.push_back(std::move(it));
                }
                return l;
            }
            case Expr::MAP: {
                Sym m;
                m.k = Sym::MAP_C;
                for (auto& k : e.kids) {
                    Sym it = lower(*k);
                    if (it.k == Sym::ERR) return err();
                    if (!is_const(it)) unsupported(e, "map literal containing request variables");
                    m.items.push_back(std::move(it));
                }
                return m;
            }
        }
        return err();
    }

    Sym index_on(const Expr& e, const Sym& base, const Sym& ix) {
        if (base.k == Sym::ERR || ix.k == Sym::ERR) return err();
        if (base.k == Sym::MAP_HTTP || base.k == Sym::MAP_CLIENT || base.k == Sym::MAP_LISTS) {
            if (ix.k != Sym::STR_C) {
                if (!is_const(ix)) unsupported(e, "map index by a request variable");
                return err();
            }
            return member(base, ix.s);
        }
        if (base.k == Sym::LIST_C || base.k == Sym::LIST_REF) {
            if (ix.k != Sym::INT_C) {
                if (!is_const(ix)) unsupported(e, "list index by a request variable");
                return err();
            }
            return list_elem(base, ix.i);
        }
        if (base.k == Sym::MAP_C) {
            if (!is_const(ix)) unsupported(e, "map index by a request variable");
            for (size_t k = 0; k + 1 < base.items.size(); k += 2)
                if (const_equal(base.items[k], ix) == 1) return base.items[k + 1];
            return err();
        }
        return err();
    }

    Sym unary_on(const Expr& e, Sym x) {
        if (e.op == Expr::OP_NOT) {
            BV b = as_bool(x);
            return boolean(P.mk_and(P.mk_not(b.t), P.mk_not(b.e)), b.e);
        }
        // negation
        if (x.k == Sym::INT_C) {
            if (x.i == INT64_MIN) return err();
            return const_int(-x.i);
        }
        if (x.k == Sym::FLOAT_C) { x.f = -x.f; return x; }
        if (x.k == Sym::INT_FEAT || x.k == Sym::INT_EXPR) {
            Sym r;
            r.k = Sym::INT_EXPR;
            append_int(r.prog, x);
            r.prog.push_back(tok(IT_NEG));
            return r;
        }
        return err();
    }

    // c ? x : y.  Boolean (or erroneous) branches fold into formulas at once; other values stay a CHOICE until the operation
    // that consumes them has been applied to both branches (lift1 / lift2): the condition decides which branch is evaluated,
    // an error in the branch not taken does not count (SEMANTICS.md A5 / A8).
    Sym select(const BV& cb, const Sym& x, const Sym& y) {
        if (P.is_const(cb.t) && P.is_const(cb.e)) {
            if (P.const_value(cb.e)) return err();
            return P.const_value(cb.t) ? x : y;
        }
        if ((x.k == Sym::BOOL || x.k == Sym::ERR) && (y.k == Sym::BOOL || y.k == Sym::ERR)) {
            BV xb = as_bool(x), yb = as_bool(y);
            if (x.k == Sym::ERR) { xb.t = 0; xb.e = 1; }
            if (y.k == Sym::ERR) { yb.t = 0; yb.e = 1; }
            const int cf = P.mk_and(P.mk_not(cb.t), P.mk_not(cb.e));
            const int t = P.mk_or(P.mk_and(cb.t, xb.t), P.mk_and(cf, yb.t));
            const int er = P.mk_or(cb.e, P.mk_or(P.mk_and(cb.t, xb.e), P.mk_and(cf, yb.e)));
            return boolean(t, er);
        }
        Sym c;
        c.k = Sym::CHOICE;
        c.bv = cb;
        c.items.push_back(x);
        c.items.push_back(y);
        return c;
    }
    template <class F>
    Sym lift1(const Sym& a, F&& f) {
        if (a.k != Sym::CHOICE) return f(a);
        return select(a.bv, lift1(a.items[0], f), lift1(a.items[1], f));
    }
    template <class F>
    Sym lift2(const Sym& a, const Sym& b, F&& f) {
        if (a.k == Sym::CHOICE) return select(a.bv, lift2(a.items[0], b, f), lift2(a.items[1], b, f));
        if (b.k == Sym::CHOICE) return select(b.bv, lift2(a, b.items[0], f), lift2(a, b.items[1], f));
        return f(a, b);
    }

    Sym member(const Sym& base, const std::string& name) {
        Sym s;
        switch (base.k) {
            case Sym::MAP_HTTP:
                for (int f = 0; f < N_FIELDS; ++f)
                    if (name == kFieldNames[f]) { s.k = Sym::STR_FIELD; s.field = f; return s; }
                return err();
            case Sym::MAP_CLIENT:
                if (name == "ip") { s.k = Sym::IP_VAR; return s; }
                if (name == "remote_port") { s.k = Sym::INT_FEAT; s.feat = IF_PORT; return s; }
                if (name == "asn") { s.k = Sym::INT_FEAT; s.feat = IF_ASN; return s; }
                if (name == "country") { s.k = Sym::COUNTRY_VAR; return s; }
                return err();
            case Sym::MAP_LISTS: {
                auto it = M.lists.find(name);
                if (it == M.lists.end()) return err();  // A6: missing key -> runtime error
                s.k = Sym::LIST_REF;
                s.list = &it->second;
                s.list_name = name;
                return s;
            }
            case Sym::MAP_C:
                for (size_t k = 0; k + 1 < base.items.size(); k += 2)
                    if (base.items[k].k == Sym::STR_C && base.items[k].s == name) return base.items[k + 1];
                return err();
            default: return err();
        }
    }

    size_t list_size(const Sym& l) {
        if (l.k == Sym::LIST_C) return l.items.size();
        switch (l.list->type) {
            case LT_STRING: return l.list->strs.size();
            case LT_INT: return l.list->ints.size();
            case LT_IP: return l.list->nets.size();
        }
        return 0;
    }

    Sym list_elem(const Sym& l, int64_t i) {
        if (i < 0 || (uint64_t)i >= list_size(l)) return err();
        if (l.k == Sym::LIST_C) return l.items[i];
        if (l.list->type == LT_STRING) return const_str(l.list->strs[i]);
        if (l.list->type == LT_INT) return const_int(l.list->ints[i]);
        return err();  // Ip element as a free-standing value: nothing in the language can consume it
    }

    // strings of a constant list (only String elements)
    std::vector<std::string> list_strings(const Sym& l) {
        std::vector<std::string> v;
        if (l.k == Sym::LIST_C) { for (auto& it : l.items) if (it.k == Sym::STR_C) v.push_back(it.s); }
        else if (l.list->type == LT_STRING) v = l.list->strs;
        return v;
    }
    std::vector<int64_t> list_ints(const Sym& l) {
        std::vector<int64_t> v;
        if (l.k == Sym::LIST_C) { for (auto& it : l.items) if (it.k == Sym::INT_C) v.push_back(it.i); }
        else if (l.list->type == LT_INT) v = l.list->ints;
        return v;
    }

    // ---- string concatenation with request fields ------------------------------------------------------------------------------
    // `http_request.host + http_request.path == "example.com/admin"`, `(method + " " + path).starts_with("POST /api")`: the value
    // stays symbolic (the parts); a comparison with a CONSTANT is decided by the finitely many ways the constant can be cut along
    // the parts -- each cut is a conjunction of `field == piece` / `field.starts_with(piece)` / ... atoms the engine already has.
    static bool is_stringy(const Sym& s) { return s.k == Sym::STR_C || s.k == Sym::STR_FIELD || s.k == Sym::CONCAT; }
    Sym concat(const Sym& a, const Sym& b) {
        std::vector<Sym> parts;
        auto take = [&](const Sym& x) {
            if (x.k == Sym::CONCAT) { for (const Sym& y : x.items) parts.push_back(y); }
            else parts.push_back(x);
        };
        take(a);
        take(b);
        std::vector<Sym> merged;
        for (Sym& x : parts) {
            if (x.k == Sym::STR_C && x.s.empty()) continue;
            if (x.k == Sym::STR_C && !merged.empty() && merged.back().k == Sym::STR_C) merged.back().s += x.s;
            else merged.push_back(std::move(x));
        }
        if (merged.empty()) return const_str("");
        if (merged.size() == 1) return merged[0];
        Sym r;
        r.k = Sym::CONCAT;
        r.items = std::move(merged);
        return r;
    }
    struct ConcatCuts {
        Lowerer& L;
        const Expr& e;
        const std::vector<Sym>& parts;
        const std::string& C;
        size_t leaves = 0;
        std::map<std::pair<size_t, size_t>, int> memo_eq, memo_sw, memo_ew;
        int leaf(const Sym& r) {
            if (++leaves > 2048) L.unsupported(e, "concatenation compared with a constant that can be cut in too many ways");
            return r.bv.t;
        }
        // parts[i..] == C[pos..]
        int eq(size_t i, size_t pos) {
            if (i == parts.size()) return L.P.constant(pos == C.size());
            auto key = std::make_pair(i, pos);
            auto it = memo_eq.find(key);
            if (it != memo_eq.end()) return it->second;
            int r = L.P.constant(false);
            const Sym& p = parts[i];
            if (p.k == Sym::STR_C) {
                if (pos + p.s.size() <= C.size() && C.compare(pos, p.s.size(), p.s) == 0) r = eq(i + 1, pos + p.s.size());
            } else {
                for (size_t l = 0; pos + l <= C.size(); ++l) {
                    const int rest = eq(i + 1, pos + l);
                    if (L.P.is_const(rest) && !L.P.const_value(rest)) continue;
                    r = L.P.mk_or(r, L.P.mk_and(leaf(L.str_literal_atom(p.field, C.substr(pos, l), true, true)), rest));
                }
            }
            return memo_eq[key] = r;
        }
        // parts[i..] starts with C[pos..]
        int sw(size_t i, size_t pos) {
            if (pos == C.size()) return L.P.constant(true);
            if (i == parts.size()) return L.P.constant(false);
            auto key = std::make_pair(i, pos);
            auto it = memo_sw.find(key);
            if (it != memo_sw.end()) return it->second;
            int r = L.P.constant(false);
            const Sym& p = parts[i];
            const size_t left = C.size() - pos;
            if (p.k == Sym::STR_C) {
                const size_t m = std::min(p.s.size(), left);
                if (C.compare(pos, m, p.s, 0, m) == 0) r = m == left ? L.P.constant(true) : sw(i + 1, pos + p.s.size());
            } else {
                r = leaf(L.str_literal_atom(p.field, C.substr(pos), true, false));   // the rest of C lies inside this field
                for (size_t l = 0; l < left; ++l) {
                    const int rest = sw(i + 1, pos + l);
                    if (L.P.is_const(rest) && !L.P.const_value(rest)) continue;
                    r = L.P.mk_or(r, L.P.mk_and(leaf(L.str_literal_atom(p.field, C.substr(pos, l), true, true)), rest));
                }
            }
            return memo_sw[key] = r;
        }
        // parts[..n) ends with C[..end)
        int ew(size_t n, size_t end) {
            if (end == 0) return L.P.constant(true);
            if (n == 0) return L.P.constant(false);
            auto key = std::make_pair(n, end);
            auto it = memo_ew.find(key);
            if (it != memo_ew.end()) return it->second;
            int r = L.P.constant(false);
            const Sym& p = parts[n - 1];
            if (p.k == Sym::STR_C) {
                const size_t m = std::min(p.s.size(), end);
                if (C.compare(end - m, m, p.s, p.s.size() - m, m) == 0) r = m == end ? L.P.constant(true) : ew(n - 1, end - p.s.size());
            } else {
                r = leaf(L.str_literal_atom(p.field, C.substr(0, end), false, true));   // C[..end) lies inside this field
                for (size_t l = 0; l < end; ++l) {
                    const int rest = ew(n - 1, end - l);
                    if (L.P.is_const(rest) && !L.P.const_value(rest)) continue;
                    r = L.P.mk_or(r, L.P.mk_and(leaf(L.str_literal_atom(p.field, C.substr(end - l, l), true, true)), rest));
                }
            }
            return memo_ew[key] = r;
        }
        int contains() {
            if (C.empty()) return L.P.constant(true);
            int r = L.P.constant(false);
            for (const Sym& p : parts) {   // inside one part
                if (p.k == Sym::STR_C) { if (p.s.find(C) != std::string::npos) return L.P.constant(true); }
                else r = L.P.mk_or(r, leaf(L.str_literal_atom(p.field, C, false, false)));
            }
            for (size_t b = 1; b < parts.size(); ++b)   // across the boundary in front of part b: C[..l) ends there, C[l..) starts there
                for (size_t l = 1; l < C.size(); ++l) {
                    const int left = ew(b, l);
                    if (L.P.is_const(left) && !L.P.const_value(left)) continue;
                    // sw() works on C from `pos`: the right-hand piece is C[l..)
                    r = L.P.mk_or(r, L.P.mk_and(left, sw(b, l)));
                }
            return r;
        }
    };
    Sym concat_length(const Expr& e, const Sym& c) {
        Sym acc = const_int(0);
        bool first = true;
        for (const Sym& p : c.items) {
            Sym term;
            if (p.k == Sym::STR_C) term = const_int((int64_t)utf8_len(p.s));
            else { term.k = Sym::INT_FEAT; term.feat = IF_LEN0 + p.field; }
            if (first) { acc = term; first = false; }
            else if (acc.k == Sym::INT_C && term.k == Sym::INT_C) acc = const_int(acc.i + term.i);
            else acc = int_arith(e, IT_ADD, acc, term);
        }
        return acc;
    }
    // CONCAT <fn> constant, fn in {"==", "contains", "starts_with", "ends_with"}
    Sym concat_pred(const Expr& e, const Sym& c, const std::string& fn, const std::string& C) {
        ConcatCuts cc{*this, e, c.items, C};
        if (fn == "==") return boolean(cc.eq(0, 0));
        if (fn == "starts_with") return boolean(cc.sw(0, 0));
        if (fn == "ends_with") return boolean(cc.ew(c.items.size(), C.size()));
        return boolean(cc.contains());
    }

    Sym method(const Expr& e) {
        Sym recv0 = lower(*e.kids[0]);
        std::vector<Sym> args0;
        for (size_t k = 1; k < e.kids.size(); ++k) args0.push_back(lower(*e.kids[k]));
        // a conditional value as receiver or (single) argument: the method is applied to both branches
        if (recv0.k == Sym::CHOICE || (args0.size() == 1 && args0[0].k == Sym::CHOICE)) {
            if (args0.size() > 1) unsupported(e, "conditional (?:) value passed to a method with several arguments");
            if (args0.empty()) return lift1(recv0, [&](const Sym& r) { return method_on(e, r, {}); });
            return lift2(recv0, args0[0], [&](const Sym& r, const Sym& a) { return method_on(e, r, std::vector<Sym>(1, a)); });
        }
        for (auto& a : args0)
            if (a.k == Sym::CHOICE) unsupported(e, "conditional (?:) value passed to a method with several arguments");
        return method_on(e, recv0, args0);
    }

    Sym method_on(const Expr& e, const Sym& recv, const std::vector<Sym>& args) {
        if (recv.k == Sym::ERR) return err();
        for (auto& a : args) if (a.k == Sym::ERR) return err();
        const std::string& fn = e.name;

        if (fn == "length") {
            if (!args.empty()) return err();
            switch (recv.k) {
                case Sym::STR_C: return const_int((int64_t)utf8_len(recv.s));
                case Sym::STR_FIELD: { Sym s; s.k = Sym::INT_FEAT; s.feat = IF_LEN0 + recv.field; return s; }
                case Sym::CONCAT: return concat_length(e, recv);
                case Sym::COUNTRY_VAR: return const_int(2);
                case Sym::LIST_C: case Sym::LIST_REF: return const_int((int64_t)list_size(recv));
                case Sym::MAP_C: return const_int((int64_t)recv.items.size() / 2);
                case Sym::MAP_LISTS: return const_int((int64_t)M.lists.size());
                case Sym::MAP_HTTP: return const_int(5);
                case Sym::MAP_CLIENT: return const_int(4);
                default: return err();
            }
        }

        if (fn == "contains" || fn == "starts_with" || fn == "ends_with" || fn == "matches") {
            if (args.size() != 1) return err();
            const Sym& a = args[0];
            // list / map receivers (contains only)
            if (recv.k == Sym::LIST_C || recv.k == Sym::LIST_REF) {
                if (fn != "contains") return err();
                return list_contains(e, recv, a);
            }
            if (recv.k == Sym::MAP_LISTS || recv.k == Sym::MAP_HTTP || recv.k == Sym::MAP_CLIENT || recv.k == Sym::MAP_C) {
                if (fn != "contains") return err();
                if (a.k == Sym::STR_FIELD || a.k == Sym::COUNTRY_VAR) {
                    // key presence with a request string: membership in the (constant) key set
                    std::vector<std::string> keys;
                    if (recv.k == Sym::MAP_HTTP) for (int f = 0; f < N_FIELDS; ++f) keys.push_back(kFieldNames[f]);
                    else if (recv.k == Sym::MAP_CLIENT) keys = {"ip", "remote_port", "asn", "country"};
                    else if (recv.k == Sym::MAP_LISTS) for (auto& kv : M.lists) keys.push_back(kv.first);
                    else for (size_t k = 0; k + 1 < recv.items.size(); k += 2) if (recv.items[k].k == Sym::STR_C) keys.push_back(recv.items[k].s);
                    if (a.k == Sym::COUNTRY_VAR) return country_atom([&](const std::string& code) { return std::find(keys.begin(), keys.end(), code) != keys.end(); });
                    return str_set_atom(e, a.field, keys);
                }
                if (a.k != Sym::STR_C) {
                    if (!is_const(a)) return const_bool(false);   // an integer / address value is never a key
                    return const_bool(false);
                }
                return const_bool(member(recv, a.s).k != Sym::ERR);
            }
            // a concatenation on either side
            if (recv.k == Sym::CONCAT) {
                if (a.k == Sym::STR_C) {
                    if (fn == "matches") unsupported(e, "matches() on a concatenation of request variables");
                    return concat_pred(e, recv, fn, a.s);
                }
                if (is_stringy(a) || a.k == Sym::COUNTRY_VAR) unsupported(e, fn + "() between a concatenation and a request variable");
                return err();
            }
            if (a.k == Sym::CONCAT) {
                if (recv.k == Sym::STR_C || recv.k == Sym::STR_FIELD || recv.k == Sym::COUNTRY_VAR) unsupported(e, fn + "() with a concatenation of request variables as argument");
                return err();
            }
            // string receivers
            auto str_pred = [&](const std::string& hay, const std::string& arg, bool* is_err) -> bool {
                *is_err = false;
                if (fn == "contains") return hay.find(arg) != std::string::npos;
                if (fn == "starts_with") return hay.size() >= arg.size() && hay.compare(0, arg.size(), arg) == 0;
                if (fn == "ends_with") return hay.size() >= arg.size() && hay.compare(hay.size() - arg.size(), arg.size(), arg) == 0;
                RegexStatus st;
                std::string msg;
                bool r = host_regex_is_match(arg, hay, &st, msg);
                if (st == RX_UNSUPPORTED) unsupported(e, "regex feature: " + msg);
                if (st != RX_OK) *is_err = true;
                return r;
            };
            if (recv.k == Sym::STR_C) {
                if (a.k == Sym::STR_C) {
                    bool er;
                    bool r = str_pred(recv.s, a.s, &er);
                    return er ? err() : const_bool(r);
                }
                if (a.k == Sym::STR_FIELD || a.k == Sym::COUNTRY_VAR) {
                    // "GET POST".contains(http_request.method): the variable must be one of the constant's substrings / prefixes /
                    // suffixes -- a finite set of strings, i.e. a membership test (the empty string is a member of all three)
                    if (fn == "matches") unsupported(e, "matches() with a variable pattern");
                    const std::string& c = recv.s;
                    if (a.k == Sym::COUNTRY_VAR) {
                        bool er;
                        return country_atom([&](const std::string& code) { return str_pred(c, code, &er); });
                    }
                    if (fn == "contains" && c.size() > 96) unsupported(e, "constant.contains(request variable) on a constant longer than 96 bytes");
                    std::vector<std::string> set(1, std::string());
                    if (fn == "starts_with") for (size_t n = 1; n <= c.size(); ++n) set.push_back(c.substr(0, n));
                    else if (fn == "ends_with") for (size_t n = 1; n <= c.size(); ++n) set.push_back(c.substr(c.size() - n));
                    else for (size_t i = 0; i < c.size(); ++i) for (size_t n = 1; i + n <= c.size(); ++n) set.push_back(c.substr(i, n));
                    return str_set_atom(e, a.field, set);
                }
                return err();
            }
            if (recv.k == Sym::STR_FIELD) {
                if (a.k == Sym::STR_FIELD || a.k == Sym::COUNTRY_VAR) {
                    if (fn != "matches" && a.k == Sym::STR_FIELD && a.field == recv.field) return const_bool(true);
                    if (fn != "matches" && a.k == Sym::STR_FIELD) return field_cmp_atom(recv.field, a.field, fn == "starts_with" ? 1 : fn == "ends_with" ? 2 : 3);
                    unsupported(e, fn + "() between two request variables");
                }
                if (a.k != Sym::STR_C) return err();
                if (fn == "contains") return str_literal_atom(recv.field, a.s, false, false);
                if (fn == "starts_with") return str_literal_atom(recv.field, a.s, true, false);
                if (fn == "ends_with") return str_literal_atom(recv.field, a.s, false, true);
                return str_regex_atom(e, recv.field, a.s);
            }
            if (recv.k == Sym::COUNTRY_VAR) {
                if (a.k == Sym::STR_FIELD) unsupported(e, fn + "() between two request variables");
                if (a.k == Sym::COUNTRY_VAR) return fn == "matches" ? (unsupported(e, "matches() with a variable pattern"), err()) : const_bool(true);
                if (a.k != Sym::STR_C) return err();
                bool any_err = false;
                Sym r = country_atom([&](const std::string& code) {
                    bool er;
                    bool v = str_pred(code, a.s, &er);
                    any_err |= er;
                    return v;
                });
                return any_err ? err() : r;
            }
            return err();
        }
        return err();  // unknown method -> runtime error
    }

    Sym list_contains(const Expr& e, const Sym& list, const Sym& x) {
        if (list.k == Sym::LIST_C && !is_const(list)) {
            // a list literal with request variables among its elements: one equality per element; an element of another type is
            // simply not equal (SEMANTICS.md A2), errors can only come from evaluating the elements or the argument
            if (x.k == Sym::CHOICE) unsupported(e, "conditional (?:) value looked up in a list of request variables");
            int t = P.constant(false), er = P.constant(false);
            for (const Sym& it : list.items) {
                const bool it_int = is_int_value(it), x_int = is_int_value(x);
                if ((it.k == Sym::INT_EXPR && prog_can_error(it.prog) && !x_int) || (x.k == Sym::INT_EXPR && prog_can_error(x.prog) && !it_int))
                    unsupported(e, "list.contains() mixing a fallible integer expression with values of another type");
                Sym c = (is_const(it) && is_const(x)) ? const_bool(const_equal(it, x) == 1) : compare(e, CMP_EQ, it, x);
                if (c.k != Sym::BOOL) continue;
                t = P.mk_or(t, c.bv.t);
                er = P.mk_or(er, c.bv.e);
            }
            return boolean(P.mk_and(t, P.mk_not(er)), er);
        }
        switch (x.k) {
            case Sym::IP_VAR:
                if (list.k == Sym::LIST_REF && list.list->type == LT_IP) return ip_set_atom(list);
                return const_bool(false);
            case Sym::STR_FIELD: return str_set_atom(e, x.field, list_strings(list));
            case Sym::COUNTRY_VAR: {
                std::vector<std::string> v = list_strings(list);
                return country_atom([&](const std::string& code) { return std::find(v.begin(), v.end(), code) != v.end(); });
            }
            case Sym::INT_FEAT: return int_set_atom(x.feat, list_ints(list));
            default: break;
        }
        if (x.k == Sym::INT_EXPR) {
            // [80, 443].contains(client.remote_port + 1): one comparison of the expression per integer element; all of them
            // share the expression's error condition.  (An empty or non-integer list still evaluates the expression: `x != x`
            // is false and carries the error.)
            const std::vector<int64_t> ints = list_ints(list);
            if (ints.size() > 16) unsupported(e, "list.contains(<integer expression>) on a list with more than 16 integers");
            if (ints.empty()) return int_expr_cmp(e, CMP_NE, x, x);
            Sym acc = int_expr_cmp(e, CMP_EQ, x, const_int(ints[0]));
            for (size_t k = 1; k < ints.size(); ++k) {
                Sym c = int_expr_cmp(e, CMP_EQ, x, const_int(ints[k]));
                acc = boolean(P.mk_or(acc.bv.t, c.bv.t), P.mk_or(acc.bv.e, c.bv.e));
            }
            return acc;
        }
        if (!is_const(x)) unsupported(e, "list.contains() of this value");
        if (list.k == Sym::LIST_C) {
            for (auto& it : list.items) if (const_equal(it, x) == 1) return const_bool(true);
            return const_bool(false);
        }
        if (x.k == Sym::STR_C && list.list->type == LT_STRING)
            return const_bool(std::find(list.list->strs.begin(), list.list->strs.end(), x.s) != list.list->strs.end());
        if (x.k == Sym::INT_C && list.list->type == LT_INT)
            return const_bool(std::find(list.list->ints.begin(), list.list->ints.end(), x.i) != list.list->ints.end());
        return const_bool(false);
    }

    static int flip_cmp(int op) {
        switch (op) {
            case CMP_LT: return CMP_GT;
            case CMP_LE: return CMP_GE;
            case CMP_GT: return CMP_LT;
            case CMP_GE: return CMP_LE;
            default: return op;
        }
    }

    template <class T>
    static bool cmp_values(int op, const T& a, const T& b) {
        switch (op) {
            case CMP_EQ: return a == b;
            case CMP_NE: return a != b;
            case CMP_LT: return a < b;
            case CMP_LE: return a <= b;
            case CMP_GT: return a > b;
            case CMP_GE: return a >= b;
        }
        return false;
    }

    Sym compare(const Expr& e, int op, const Sym& a, const Sym& b) {
        bool ordering = op != CMP_EQ && op != CMP_NE;
        // constants
        if (is_const(a) && is_const(b)) {
            if (!ordering) {
                int r = const_equal(a, b);
                if (r < 0) return err();
                return const_bool(op == CMP_EQ ? r == 1 : r == 0);
            }
            if (a.k != b.k) return err();
            if (a.k == Sym::INT_C || a.k == Sym::UINT_C) return const_bool(a.k == Sym::INT_C ? cmp_values(op, a.i, b.i) : cmp_values(op, (uint64_t)a.i, (uint64_t)b.i));
            if (a.k == Sym::FLOAT_C) return const_bool(cmp_values(op, a.f, b.f));
            if (a.k == Sym::STR_C || a.k == Sym::BYTES_C) return const_bool(cmp_values(op, a.s, b.s));
            return err();
        }
        // variable on the right: swap
        if (is_const(a) && !is_const(b)) return compare(e, flip_cmp(op), b, a);
        if (a.k == Sym::CONCAT || b.k == Sym::CONCAT) {
            const Sym& c = a.k == Sym::CONCAT ? a : b;
            const Sym& o = a.k == Sym::CONCAT ? b : a;
            if (o.k == Sym::STR_C) {
                if (ordering) unsupported(e, "lexicographic ordering of a concatenation of request variables");
                Sym eq = concat_pred(e, c, "==", o.s);
                return op == CMP_EQ ? eq : boolean(P.mk_not(eq.bv.t));
            }
            if (is_stringy(o) || o.k == Sym::COUNTRY_VAR) unsupported(e, "comparison of a concatenation with a request variable");
            return err();   // a String against a value of another type
        }
        // a is a variable
        switch (a.k) {
            case Sym::STR_FIELD:
                if (b.k == Sym::STR_C) {
                    if (ordering) return str_order_atom(a.field, b.s, op);
                    Sym eq = str_literal_atom(a.field, b.s, true, true);
                    return op == CMP_EQ ? eq : boolean(P.mk_not(eq.bv.t));
                }
                if (b.k == Sym::STR_FIELD) {
                    if (b.field == a.field) return const_bool(op == CMP_EQ || op == CMP_LE || op == CMP_GE);
                    if (ordering) return field_cmp_atom(a.field, b.field, op == CMP_LT ? 4 : op == CMP_LE ? 5 : op == CMP_GT ? 6 : 7);   // byte-wise, as Rust's str::cmp
                    Sym eq = field_cmp_atom(a.field, b.field, 0);
                    return op == CMP_EQ ? eq : boolean(P.mk_not(eq.bv.t));
                }
                if (b.k == Sym::COUNTRY_VAR) unsupported(e, "comparison between two request variables");
                return err();
            case Sym::COUNTRY_VAR:
                if (b.k == Sym::STR_C) return country_atom([&](const std::string& code) { return cmp_values(op, code, b.s); });
                if (b.k == Sym::COUNTRY_VAR) return const_bool(op == CMP_EQ || op == CMP_LE || op == CMP_GE);
                if (b.k == Sym::STR_FIELD) unsupported(e, "comparison between two request variables");
                return err();
            case Sym::INT_FEAT:
                if (b.k == Sym::INT_C) return int_cmp_atom(a.feat, op, b.i);
                if (b.k == Sym::INT_FEAT && b.feat == a.feat) return const_bool(op == CMP_EQ || op == CMP_LE || op == CMP_GE);
                if (is_int_value(b)) return int_expr_cmp(e, op, a, b);
                return err();
            case Sym::INT_EXPR:
                if (is_int_value(b)) return int_expr_cmp(e, op, a, b);
                return err();
            case Sym::IP_VAR:
                if (b.k == Sym::IP_VAR && !ordering) return const_bool(op == CMP_EQ);
                return err();  // no Ip literal syntax: every other comparison is cross-type
            case Sym::BOOL:
                if (b.k == Sym::BOOL && !ordering) {
                    // bool == bool on request-dependent values
                    BV x = a.bv, y = b.bv;
                    int both = P.mk_or(P.mk_and(x.t, y.t), P.mk_and(P.mk_not(x.t), P.mk_not(y.t)));
                    int t = op == CMP_EQ ? both : P.mk_not(both);
                    int er = P.mk_or(x.e, y.e);
                    return boolean(P.mk_and(t, P.mk_not(er)), er);
                }
                return err();
            case Sym::MAP_HTTP: case Sym::MAP_CLIENT: case Sym::MAP_LISTS:
                unsupported(e, "comparison of whole maps");
            case Sym::LIST_C:
                unsupported(e, "comparison of a whole list that holds request variables");
            default: return err();
        }
    }

    Sym binary(const Expr& e) {
        Sym a0 = lower(*e.kids[0]);
        Sym b0 = lower(*e.kids[1]);
        if (a0.k == Sym::CHOICE || b0.k == Sym::CHOICE) return lift2(a0, b0, [&](const Sym& x, const Sym& y) { return binary_on(e, x, y); });
        return binary_on(e, a0, b0);
    }

    Sym binary_on(const Expr& e, const Sym& a, const Sym& b) {
        switch (e.op) {
            case Expr::OP_AND: {
                BV x = as_bool(a), y = as_bool(b);
                return boolean(P.mk_and(x.t, y.t), P.mk_or(x.e, P.mk_and(x.t, y.e)));
            }
            case Expr::OP_OR: {
                BV x = as_bool(a), y = as_bool(b);
                int xf = P.mk_and(P.mk_not(x.t), P.mk_not(x.e));
                return boolean(P.mk_or(x.t, P.mk_and(xf, y.t)), P.mk_or(x.e, P.mk_and(xf, y.e)));
            }
            case Expr::OP_IN: return err();  // "@in" is not provided (reference rules/rules.rs:67-71)
            case Expr::OP_EQ: case Expr::OP_NE: case Expr::OP_LT: case Expr::OP_LE: case Expr::OP_GT: case Expr::OP_GE: {
                if (a.k == Sym::ERR || b.k == Sym::ERR) return err();
                int op = e.op == Expr::OP_EQ ? CMP_EQ : e.op == Expr::OP_NE ? CMP_NE : e.op == Expr::OP_LT ? CMP_LT
                         : e.op == Expr::OP_LE ? CMP_LE : e.op == Expr::OP_GT ? CMP_GT : CMP_GE;
                return compare(e, op, a, b);
            }
            default: break;
        }
        // arithmetic
        if (a.k == Sym::ERR || b.k == Sym::ERR) return err();
        if ((a.k == Sym::INT_FEAT || a.k == Sym::INT_EXPR || b.k == Sym::INT_FEAT || b.k == Sym::INT_EXPR)) {
            if (!is_int_value(a) || !is_int_value(b)) return err();   // no implicit conversions: Int with Float / String is an error
            switch (e.op) {
                case Expr::OP_ADD: return int_arith(e, IT_ADD, a, b);
                case Expr::OP_SUB: return int_arith(e, IT_SUB, a, b);
                case Expr::OP_MUL: return int_arith(e, IT_MUL, a, b);
                case Expr::OP_DIV: return int_arith(e, IT_DIV, a, b);
                case Expr::OP_MOD: return int_arith(e, IT_MOD, a, b);
                default: return err();
            }
        }
        if (e.op == Expr::OP_ADD && (a.k == Sym::STR_FIELD || b.k == Sym::STR_FIELD || a.k == Sym::CONCAT || b.k == Sym::CONCAT)) {
            if (is_stringy(a) && is_stringy(b)) return concat(a, b);
            if (a.k == Sym::COUNTRY_VAR || b.k == Sym::COUNTRY_VAR) unsupported(e, "string concatenation with client.country");
            return err();   // String + a value of another type
        }
        if (a.k == Sym::COUNTRY_VAR || b.k == Sym::COUNTRY_VAR)
            if (e.op == Expr::OP_ADD) unsupported(e, "string concatenation with client.country");
        if (a.k == Sym::INT_C && b.k == Sym::INT_C) {
            int64_t r = 0;
            switch (e.op) {
                case Expr::OP_ADD: if (__builtin_add_overflow(a.i, b.i, &r)) return err(); break;
                case Expr::OP_SUB: if (__builtin_sub_overflow(a.i, b.i, &r)) return err(); break;
                case Expr::OP_MUL: if (__builtin_mul_overflow(a.i, b.i, &r)) return err(); break;
                case Expr::OP_DIV: if (b.i == 0 || (a.i == INT64_MIN && b.i == -1)) return err(); r = a.i / b.i; break;
                case Expr::OP_MOD: if (b.i == 0 || (a.i == INT64_MIN && b.i == -1)) return err(); r = a.i % b.i; break;
                default: return err();
            }
            return const_int(r);
        }
        if (a.k == Sym::FLOAT_C && b.k == Sym::FLOAT_C) {
            Sym s; s.k = Sym::FLOAT_C;
            switch (e.op) {
                case Expr::OP_ADD: s.f = a.f + b.f; break;
                case Expr::OP_SUB: s.f = a.f - b.f; break;
                case Expr::OP_MUL: s.f = a.f * b.f; break;
                case Expr::OP_DIV: s.f = a.f / b.f; break;
                default: return err();
            }
            return s;
        }
        if (a.k == Sym::STR_C && b.k == Sym::STR_C && e.op == Expr::OP_ADD) return const_str(a.s + b.s);
        if (a.k == Sym::LIST_C && b.k == Sym::LIST_C && e.op == Expr::OP_ADD) {
            Sym s = a;
            s.items.insert(s.items.end(), b.items.begin(), b.items.end());
            return s;
        }
        return err();
    }
};

}  // namespace

int lower_rule_expression(Model& model, const Expr& e, const std::string& rule_name) {
    Lowerer L(model, rule_name);
    Sym r0 = L.lower(e);
    // Rule::match_request: matched iff execute() == Ok(Bool(true))  (pingoo/rules.rs:36-52); a conditional whose taken
    // branch is not a Bool is "no match" like any other non-bool value
    Sym r = L.lift1(r0, [&](const Sym& v) { return v.k == Sym::BOOL ? v : L.boolean(0, 0); });
    if (r.k != Sym::BOOL) return model.pool.constant(false);
    return r.bv.t;
}

}  // namespace pgw
