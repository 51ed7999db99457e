// Model -> HostProgram: DFA groups per field, rule bytecode, candidate indexes.
#include "compile.hpp"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iterator>
#include <map>
#include <set>
#include <sstream>

#include "dfa.hpp"
#include "gate.hpp"
#include "nfa_bits.hpp"

namespace pgw {
namespace {

void pad16(std::vector<uint8_t>& v) {
    while (v.size() % 16) v.push_back(0);
}

// Emit postfix code for formula `n`; returns the stack depth it needs.
int emit_code(const BoolPool& P, int n, std::vector<uint16_t>& code) {
    const BoolNode& b = P.at(n);
    switch (b.kind) {
        case BoolNode::CONST: code.push_back(b.v ? OP_PUSH1 : OP_PUSH0); return 1;
        case BoolNode::ATOM: code.push_back((uint16_t)b.a); return 1;
        case BoolNode::NOT: {
            int d = emit_code(P, b.a, code);
            code.push_back(OP_NOT);
            return d;
        }
        case BoolNode::AND:
        case BoolNode::OR: {
            std::vector<uint16_t> ca, cb;
            int da = emit_code(P, b.a, ca), db = emit_code(P, b.b, cb);
            // deeper operand first keeps the stack shallow
            if (da >= db) { code.insert(code.end(), ca.begin(), ca.end()); code.insert(code.end(), cb.begin(), cb.end()); }
            else { code.insert(code.end(), cb.begin(), cb.end()); code.insert(code.end(), ca.begin(), ca.end()); }
            code.push_back(b.kind == BoolNode::AND ? OP_AND : OP_OR);
            return std::max(std::max(da, db), std::min(da, db) + 1);
        }
    }
    return 1;
}

void collect_atoms(const BoolPool& P, int n, bool neg, std::vector<std::pair<int, bool>>& out) {
    const BoolNode& b = P.at(n);
    switch (b.kind) {
        case BoolNode::CONST: return;
        case BoolNode::ATOM: out.emplace_back(b.a, neg); return;
        case BoolNode::NOT: collect_atoms(P, b.a, !neg, out); return;
        default:
            collect_atoms(P, b.a, neg, out);
            collect_atoms(P, b.b, neg, out);
    }
}

uint8_t terminal_for(const std::vector<uint8_t>& actions, bool captcha_verified) {
    // http_listener.rs:253-261: walk the actions of a matched rule in order
    for (uint8_t a : actions) {
        if (a == ACT_BLOCK) return V_BLOCK;
        if (a == ACT_CAPTCHA && !captcha_verified) return V_CAPTCHA;
    }
    return 0;
}

}  // namespace

std::string HostProgram::summary() const {
    std::ostringstream o;
    o << "rules=" << n_rules << " atoms=" << n_atoms << " units=" << units.size() << " arena_bytes=" << arena.size();
    for (size_t u = 0; u < units.size(); ++u)
        o << " [" << kFieldNames[units[u].field] << (units[u].mode == UM_CANDIDATES ? "/gated" : units[u].abs0 != 0xFFFFFFFFu ? "/early-exit" : "")
          << ": states=" << units[u].n_states << " classes=" << units[u].n_classes << "]";
    for (const BitsetUnitDesc& b : bitset_units)
        o << " [" << kFieldNames[b.field] << "/bitset-nfa: positions=" << b.n_pos << " contexts=" << b.n_tables << " classes=" << b.n_classes << "]";
    for (int f = 0; f < N_FIELDS; ++f)
        if (gate[f].present) {
            o << " gate(" << kFieldNames[f] << ": grams=" << gate[f].n_grams << " bloom=2^" << gate[f].k1 << " table=2^" << gate[f].kt;
            if (gate[f].lits.size() > 1 || gate[f].lits[0].len) o << " literals=" << gate[f].lits.size();
            if (gate[f].slot_words == 4) o << " wide-slots";
            o << ")";
        }
    o << " ns_atoms=" << ns_atoms.size() << " lpm=" << (lpm.present ? 1 : 0);
    return o.str();
}

bool compile_program(Model& M, const CompileOptions& opt, const std::vector<uint8_t>& geo_mmdb, HostProgram* out,
                     std::string& err) {
    HostProgram& H = *out;
    H = HostProgram();
    H.eval_gates = opt.eval_gates;
    H.warnings = M.warnings;
    H.n_rules = (uint32_t)M.rules.size();
    H.n_waf_rules = M.n_waf_rules <= M.rules.size() ? M.n_waf_rules : (uint32_t)M.rules.size();
    if (H.n_rules - H.n_waf_rules >= 0xFFFFu) {
        err = "too many services";
        return false;
    }

    // ---- internal gate atom: path.starts_with("/__pingoo/captcha") (http_listener.rs:200-204)
    if (opt.eval_gates) {
        std::string key = "S|" + std::to_string((int)F_PATH) + "|^|17:/__pingoo/captcha";
        auto it = M.atom_index.find(key);
        if (it == M.atom_index.end()) {
            AtomDesc a;
            a.kind = AtomDesc::STR_PATTERN;
            a.field = F_PATH;
            a.key = key;
            int id = (int)M.atoms.size();
            a.event_base = (int)M.events.size();
            a.nfa_starts.push_back(nfa_literal(M.nfa[F_PATH], "/__pingoo/captcha", true, false, a.event_base));
            M.events.push_back(PatternEvent{EV_FIRE, id});
            M.atom_index[key] = id;
            M.atoms.push_back(a);
            H.gate_bypass_atom = id;
        } else H.gate_bypass_atom = it->second;
    }

    // ---- complement events for atoms that are expected to be true -------------------------------------------
    // `!ua.starts_with("Mozilla/")` (docs/configuration.md:67-70) makes the atom true for most requests, which would
    // fire one accept event per request.  For literals anchored at the start the complement language is regular and
    // cheap ("some byte of the prefix differs, or the field ends early"): scan for THAT, and negate in the formula,
    // so the common case fires nothing.
    {
        std::vector<int> pos(M.atoms.size(), 0), neg(M.atoms.size(), 0);
        for (auto& r : M.rules) {
            std::vector<std::pair<int, bool>> refs;
            collect_atoms(M.pool, r.formula, false, refs);
            for (auto& pr : refs) (pr.second ? neg : pos)[pr.first]++;
        }
        std::vector<int> repl(M.atoms.size(), -1);
        const size_t n0 = M.atoms.size();
        for (size_t a = 0; a < n0; ++a) {
            if (M.atoms[a].kind != AtomDesc::STR_PATTERN || M.atoms[a].lit_kind == 0 || neg[a] <= pos[a]) continue;
            if ((int)a == H.gate_bypass_atom) continue;
            const int f = M.atoms[a].field;
            const std::string lit = M.atoms[a].lit;
            AtomDesc c;
            c.kind = AtomDesc::STR_PATTERN;
            c.field = f;
            c.key = "S|" + std::to_string(f) + "|not" + (M.atoms[a].lit_kind == 2 ? "eq" : "sw") + "|" + lit;
            c.event_base = (int)M.events.size();
            const int id = (int)M.atoms.size();
            Nfa& nfa = M.nfa[f];
            auto add = [&](NfaKind k) { NfaNode n; n.kind = k; nfa.nodes.push_back(n); return (int)nfa.nodes.size() - 1; };
            int m = add(N_MATCH);
            nfa.nodes[m].pattern = c.event_base;
            int next = -1;  // continuation after the whole literal matched: dead for starts_with, "any further byte" for ==
            if (M.atoms[a].lit_kind == 2) {
                ByteSet any;
                any.negate();
                next = add(N_CHAR);
                nfa.nodes[next].set = nfa.add_set(any);
                nfa.nodes[next].out = m;
            }
            for (size_t k = lit.size(); k-- > 0;) {
                ByteSet is, isnot;
                is.set((unsigned char)lit[k]);
                isnot = is;
                isnot.negate();
                int diff = add(N_CHAR);  // a different byte here
                nfa.nodes[diff].set = nfa.add_set(isnot);
                nfa.nodes[diff].out = m;
                int eol = add(N_ASSERT);  // or the field ends here
                nfa.nodes[eol].assert_kind = A_EOL_TEXT;
                nfa.nodes[eol].out = m;
                int alt = add(N_SPLIT);
                nfa.nodes[alt].out = diff;
                nfa.nodes[alt].out1 = eol;
                int node = alt;
                if (next >= 0) {
                    int same = add(N_CHAR);
                    nfa.nodes[same].set = nfa.add_set(is);
                    nfa.nodes[same].out = next;
                    int sp = add(N_SPLIT);
                    nfa.nodes[sp].out = alt;
                    nfa.nodes[sp].out1 = same;
                    node = sp;
                }
                next = node;
            }
            if (next < 0) continue;  // empty starts_with literal cannot get here
            int bol = add(N_ASSERT);
            nfa.nodes[bol].assert_kind = A_BOL_TEXT;
            nfa.nodes[bol].out = next;
            c.nfa_starts.push_back(bol);
            M.events.push_back(PatternEvent{EV_FIRE, id});
            M.atom_index[c.key] = id;
            M.atoms.push_back(c);
            repl[a] = M.pool.mk_not(M.pool.atom(id));
        }
        bool any = false;
        for (int r : repl) any |= r >= 0;
        if (any) {
            repl.resize(M.atoms.size(), -1);
            for (auto& r : M.rules) r.formula = M.pool.substitute(r.formula, repl);
        }
    }

    if (M.atoms.size() > 0x3FFF) {
        err = "too many distinct predicates (" + std::to_string(M.atoms.size()) + " > 16383)";
        return false;
    }
    H.n_atoms = (uint32_t)M.atoms.size();
    H.atom_words = std::max<uint32_t>(1, (H.n_atoms + 31) / 32);
    H.expect.assign(H.atom_words, 0);
    H.care.assign(H.atom_words, 0);

    // ---- polarity statistics, atom -> rules index ---------------------------------
    std::vector<std::vector<uint32_t>> atom_rules(H.n_atoms);
    for (size_t r = 0; r < M.rules.size(); ++r) {
        std::vector<std::pair<int, bool>> refs;
        collect_atoms(M.pool, M.rules[r].formula, false, refs);
        for (auto& pr : refs) {
            if (pr.second) M.atoms[pr.first].neg_refs++;
            else M.atoms[pr.first].pos_refs++;
            if (atom_rules[pr.first].empty() || atom_rules[pr.first].back() != (uint32_t)r) atom_rules[pr.first].push_back((uint32_t)r);
            H.care[pr.first >> 5] |= 1u << (pr.first & 31);
        }
    }
    std::vector<uint8_t> expect_vals(H.n_atoms, 0);
    for (uint32_t a = 0; a < H.n_atoms; ++a)
        if (M.atoms[a].neg_refs > M.atoms[a].pos_refs) {
            expect_vals[a] = 1;
            H.expect[a >> 5] |= 1u << (a & 31);
        }
    H.ar_idx.assign(1, 0);
    for (uint32_t a = 0; a < H.n_atoms; ++a) {
        H.ar_rules.insert(H.ar_rules.end(), atom_rules[a].begin(), atom_rules[a].end());
        H.ar_idx.push_back((uint32_t)H.ar_rules.size());
    }

    // ---- rule bytecode, terminal actions, default verdicts -----------------------
    H.rule_off.assign(1, 0);
    for (size_t r = 0; r < M.rules.size(); ++r) {
        std::vector<uint16_t> code;
        int depth = emit_code(M.pool, M.rules[r].formula, code);
        if (depth > (int)kMaxStackDepth) {
            err = "rule '" + M.rules[r].name + "': expression too deeply nested for the evaluator (" + std::to_string(depth) + ")";
            return false;
        }
        H.code.insert(H.code.end(), code.begin(), code.end());
        H.rule_off.push_back((uint32_t)H.code.size());
        uint8_t t0 = terminal_for(M.rules[r].actions, false), t1 = terminal_for(M.rules[r].actions, true);
        if (M.rules[r].is_service) t0 = t1 = 0;
        H.term.push_back((uint8_t)(t0 | (t1 << 2)));
    }
    // The most common deviation from the expected vector is a single atom (a non-GET method, one matched pattern):
    // its verdict and service are tabulated here so that the epilogue needs no rule evaluation for it either.
    {
        H.v1.assign(2 * (size_t)H.n_atoms, 0);
        H.s1.assign(H.n_atoms, (uint16_t)kNoService);
        std::vector<uint8_t> vals = expect_vals;
        for (uint32_t a = 0; a < H.n_atoms; ++a) {
            if (!((H.care[a >> 5] >> (a & 31)) & 1u)) continue;
            vals[a] ^= 1;
            for (int cv = 0; cv < 2; ++cv) {
                uint32_t v = V_ALLOW | (kNoRule << 2);
                for (size_t r = 0; r < H.n_waf_rules; ++r) {
                    uint8_t t = (H.term[r] >> (2 * cv)) & 3;
                    if (!t || !M.pool.eval(M.rules[r].formula, vals)) continue;
                    v = t | ((uint32_t)r << 2);
                    break;
                }
                H.v1[(size_t)cv * H.n_atoms + a] = v;
            }
            for (size_t r = H.n_waf_rules; r < M.rules.size(); ++r)
                if (M.pool.eval(M.rules[r].formula, vals)) { H.s1[a] = (uint16_t)(r - H.n_waf_rules); break; }
            vals[a] ^= 1;
        }
    }
    // services: first route that is true under the expected atom values, and every route that is (candidates when
    // some atom deviates)
    H.s0 = 0xFFFFu;
    H.dflt_services.clear();
    for (size_t r = H.n_waf_rules; r < M.rules.size(); ++r) {
        if (!M.pool.eval(M.rules[r].formula, expect_vals)) continue;
        H.dflt_services.push_back((uint32_t)r);
        if (H.s0 == 0xFFFFu) H.s0 = (uint32_t)(r - H.n_waf_rules);
        if (M.pool.is_const(M.rules[r].formula)) break;  // a route-less service shadows everything after it
    }
    for (int cv = 0; cv < 2; ++cv) {
        H.v0[cv] = V_ALLOW | (kNoRule << 2);
        bool have_v0 = false;
        for (size_t r = 0; r < M.rules.size(); ++r) {
            uint8_t t = (H.term[r] >> (2 * cv)) & 3;
            if (!t) continue;
            if (!M.pool.eval(M.rules[r].formula, expect_vals)) continue;
            H.dflt_rules[cv].push_back((uint32_t)r);
            if (!have_v0) { H.v0[cv] = t | ((uint32_t)r << 2); have_v0 = true; }
            if (M.pool.is_const(M.rules[r].formula)) break;  // an unconditional rule shadows everything after it
        }
    }

    // a request none of whose atoms is true (the overwhelmingly common case once complements are in place): its
    // verdict and service are constants, the epilogue writes them without touching the atom bitmap
    {
        std::vector<uint8_t> zeros(H.n_atoms, 0);
        for (int cv = 0; cv < 2; ++cv) {
            H.vclean[cv] = V_ALLOW | (kNoRule << 2);
            for (size_t r = 0; r < H.n_waf_rules; ++r) {
                uint8_t t = (H.term[r] >> (2 * cv)) & 3;
                if (!t || !M.pool.eval(M.rules[r].formula, zeros)) continue;
                H.vclean[cv] = t | ((uint32_t)r << 2);
                break;
            }
        }
        H.sclean = kNoService;
        for (size_t r = H.n_waf_rules; r < M.rules.size(); ++r)
            if (M.pool.eval(M.rules[r].formula, zeros)) { H.sclean = (uint32_t)(r - H.n_waf_rules); break; }
        // ... and one with exactly one true atom (one matched pattern, a non-GET method, ...): tabulated as well.
        // Only rules that mention the atom, or that are true on the all-false vector, can be true there.
        H.v1z.assign(2 * (size_t)H.n_atoms, 0);
        H.s1z.assign(H.n_atoms, (uint16_t)kNoService);
        std::vector<uint32_t> zero_true;  // rules true when every atom is false
        for (size_t r = 0; r < M.rules.size(); ++r)
            if (M.pool.eval(M.rules[r].formula, zeros)) zero_true.push_back((uint32_t)r);
        for (uint32_t a = 0; a < H.n_atoms; ++a) {
            zeros[a] = 1;
            std::vector<uint32_t> cands(atom_rules[a].begin(), atom_rules[a].end());
            cands.insert(cands.end(), zero_true.begin(), zero_true.end());
            std::sort(cands.begin(), cands.end());
            cands.erase(std::unique(cands.begin(), cands.end()), cands.end());
            std::vector<uint32_t> true_rules;
            for (uint32_t r : cands)
                if (M.pool.eval(M.rules[r].formula, zeros)) true_rules.push_back(r);
            for (int cv = 0; cv < 2; ++cv) {
                uint32_t v = V_ALLOW | (kNoRule << 2);
                for (uint32_t r : true_rules) {
                    if (r >= H.n_waf_rules) break;
                    uint8_t t = (H.term[r] >> (2 * cv)) & 3;
                    if (!t) continue;
                    v = t | (r << 2);
                    break;
                }
                H.v1z[(size_t)cv * H.n_atoms + a] = v;
            }
            for (uint32_t r : true_rules)
                if (r >= H.n_waf_rules) { H.s1z[a] = (uint16_t)(r - H.n_waf_rules); break; }
            zeros[a] = 0;
        }
    }

    // Two true atoms, the second most common case after none and one: when no rule mentions both and neither is mentioned
    // by a rule that is true on the all-false vector, the rules true under {a, b} are exactly those true under {a} alone or
    // {b} alone, so the verdict is the earlier of v1z[a] and v1z[b].  The device tests "no common rule" conservatively on
    // 64-bit signatures (one bit per rule, hashed); an atom a zero-true rule mentions gets an all-ones signature.
    {
        std::vector<uint8_t> zeros(H.n_atoms, 0);
        H.atom_sig.assign(H.n_atoms, 0);
        for (uint32_t a = 0; a < H.n_atoms; ++a)
            for (uint32_t r : atom_rules[a]) {
                H.atom_sig[a] |= 1ull << ((r * 0x9E3779B97F4A7C15ull) >> 58);
                if (M.pool.eval(M.rules[r].formula, zeros)) H.atom_sig[a] = ~0ull;
            }
        for (uint32_t a = 0; a < H.n_atoms; ++a)
            for (uint32_t r : atom_rules[a])
                if (M.pool.eval(M.rules[r].formula, zeros)) H.atom_sig[a] = ~0ull;
    }

    // ---- scan units: DFA groups per field ------------------------------------------
    // The string atoms of a field fall into three classes:
    //   anchored  every triggering pattern is tied to the start of the field (starts_with, ==, ^...): decided by a prefix,
    //             walked for every request but only until the DFA reaches an absorbing state (UnitDesc::abs0/abs1);
    //   gated     url / user_agent / path patterns the candidate gate covers (gate.hpp): walked for gate candidates only;
    //   full      everything else: walked for every request, whole field.
    bool len_feat_used[N_FIELDS] = {false, false, false, false, false};
    static const int kFieldOrder[N_FIELDS] = {F_URL, F_USER_AGENT, F_PATH, F_HOST, F_METHOD};  // longest first
    enum { UC_FULL = 0, UC_ANCH = 1, UC_GATED = 2, N_UC = 3 };
    // arena: all class maps first, then the tables
    struct Pending { Dfa dfa; int field; uint32_t mode; uint32_t gate_bit = 0; std::vector<int> latch_of_event; };
    std::vector<Pending> pend;
    struct GateInput { bool present = false; std::vector<uint32_t> grams, masks; std::vector<GateLiteral> literals; } gate_in[N_FIELDS];
    for (int fo = 0; fo < N_FIELDS; ++fo) {
        int f = kFieldOrder[fo];
        const bool gate_field = opt.candidate_gate && (f == F_URL || f == F_USER_AGENT || f == F_PATH);
        std::vector<PatternBundle> bundles[N_UC];
        std::vector<uint32_t> bundle_atom[N_UC];
        struct GatedGrams { std::vector<uint32_t> grams; };
        std::vector<GatedGrams> gated_grams;
        // atoms whose pattern is a small finite set of strings: confirmed by the gate's resolve kernel, no automaton (gate.hpp)
        std::vector<GateLiteral> literals;
        std::vector<uint32_t> literal_grams;   // distinct grams the literals put into the field's budget
        // What the gate knows about each atom of the field, computed once: the grams of its triggering parts (FIRE, and SET of a
        // gap-split pattern: TEST needs an earlier SET, CLEAR never fires) and, if its language is a small finite set of strings,
        // those strings with the grams that announce them.
        struct Pre {
            bool used = false, all_anch = true, all_gate = true, lit_ok = false;
            std::vector<uint32_t> grams;       // gate_grams_for_pattern over the triggering parts
            std::vector<LitString> strs;
            std::vector<uint32_t> lit_grams;   // sorted, distinct
        };
        std::vector<Pre> pre(H.n_atoms);
        std::map<uint32_t, uint32_t> gram_owners;   // gram -> atoms of this field that have it in either set
        if (gate_field)
            for (uint32_t a = 0; a < H.n_atoms; ++a) {
                const AtomDesc& ad = M.atoms[a];
                if (ad.kind != AtomDesc::STR_PATTERN || ad.field != f) continue;
                if (ad.pos_refs + ad.neg_refs == 0 && (int)a != H.gate_bypass_atom) continue;
                Pre& pr = pre[a];
                pr.used = true;
                for (size_t k = 0; k < ad.nfa_starts.size(); ++k) {
                    const uint8_t kind = M.events[ad.event_base + k].kind;
                    if (kind != EV_FIRE && kind != EV_SET) continue;
                    if (!pattern_is_start_anchored(M.nfa[f], ad.nfa_starts[k])) pr.all_anch = false;
                    if (pr.all_gate && !gate_grams_for_pattern(M.nfa[f], ad.nfa_starts[k], opt.gate_pattern_cap, &pr.grams)) pr.all_gate = false;
                }
                if (pr.all_anch) continue;   // early-exit class: the gate is not involved
                if (opt.literal_confirm && ad.nfa_starts.size() == 1 && !ad.has_latch && M.events[ad.event_base].kind == EV_FIRE &&
                    (int)a != H.gate_bypass_atom && gate_finite_language(M.nfa[f], ad.nfa_starts[0], &pr.strs)) {
                    pr.lit_ok = true;
                    for (const LitString& ls : pr.strs) {
                        std::vector<std::pair<uint32_t, int>> pg;
                        gate_grams_for_literal(ls, &pg);
                        for (auto& x : pg) pr.lit_grams.push_back(x.first);
                    }
                    std::sort(pr.lit_grams.begin(), pr.lit_grams.end());
                    pr.lit_grams.erase(std::unique(pr.lit_grams.begin(), pr.lit_grams.end()), pr.lit_grams.end());
                }
                std::vector<uint32_t> all = pr.lit_grams;
                if (pr.all_gate) all.insert(all.end(), pr.grams.begin(), pr.grams.end());
                std::sort(all.begin(), all.end());
                all.erase(std::unique(all.begin(), all.end()), all.end());
                for (uint32_t g : all) gram_owners[g]++;
            }
        for (uint32_t a = 0; a < H.n_atoms; ++a)
            if (M.atoms[a].kind == AtomDesc::STR_PATTERN && M.atoms[a].field == f) {
                // atoms no rule refers to any more (replaced by their complement) are not scanned
                if (M.atoms[a].pos_refs + M.atoms[a].neg_refs == 0 && (int)a != H.gate_bypass_atom) continue;
                PatternBundle b;
                b.starts = M.atoms[a].nfa_starts;
                b.has_latch = M.atoms[a].has_latch;
                int cls = UC_FULL;
                if (gate_field) {
                    Pre& pr = pre[a];
                    if (pr.all_anch) cls = UC_ANCH;
                    else {
                        // Literal confirmation pays when a window that carries one of the atom's grams costs ONE comparison and
                        // nothing else (measured: a comparison in the resolve kernel costs about 2/3 of walking a candidate,
                        // profiles/README.md): the atom's grams must be its own -- no other atom of the field, confirmed or walked by
                        // a DFA, may share one.  Families of strings with a common prefix (`sqlmap1`, `sqlmap2`, ...) are what an
                        // automaton is for and stay with the DFA units.
                        bool exclusive = pr.lit_ok;
                        for (size_t k = 0; exclusive && k < pr.lit_grams.size(); ++k) exclusive = gram_owners[pr.lit_grams[k]] == 1;
                        if (exclusive) {
                            std::vector<uint32_t> lg;
                            std::set_union(literal_grams.begin(), literal_grams.end(), pr.lit_grams.begin(), pr.lit_grams.end(), std::back_inserter(lg));
                            if (lg.size() <= opt.gate_field_cap / 2) {   // half of the field's gram budget at most
                                literal_grams.swap(lg);
                                for (const LitString& ls : pr.strs) literals.push_back(GateLiteral{ls, a});
                                continue;   // no DFA for this atom
                            }
                        }
                        if (pr.all_gate) { cls = UC_GATED; gated_grams.push_back(GatedGrams{std::move(pr.grams)}); }
                    }
                }
                bundles[cls].push_back(std::move(b));
                bundle_atom[cls].push_back(a);
            }
        // the field's gram budget (distinct grams: patterns built from one template share most of theirs): patterns are
        // admitted smallest set first, the ones that no longer fit fall back to the full class
        if (!bundles[UC_GATED].empty()) {
            std::vector<size_t> order(gated_grams.size());
            for (size_t i = 0; i < order.size(); ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return gated_grams[x].grams.size() < gated_grams[y].grams.size(); });
            std::vector<uint32_t> all = literal_grams;   // sorted, distinct
            std::vector<char> keep(gated_grams.size(), 0);
            for (size_t i : order) {
                std::vector<uint32_t> g = gated_grams[i].grams, merged;
                std::sort(g.begin(), g.end());
                g.erase(std::unique(g.begin(), g.end()), g.end());
                std::set_union(all.begin(), all.end(), g.begin(), g.end(), std::back_inserter(merged));
                if (merged.size() > opt.gate_field_cap) continue;
                all.swap(merged);
                keep[i] = 1;
            }
            std::vector<PatternBundle> kept_b;
            std::vector<uint32_t> kept_a;
            for (size_t i = 0; i < gated_grams.size(); ++i) {
                if (keep[i]) { kept_b.push_back(std::move(bundles[UC_GATED][i])); kept_a.push_back(bundle_atom[UC_GATED][i]); }
                else { bundles[UC_FULL].push_back(std::move(bundles[UC_GATED][i])); bundle_atom[UC_FULL].push_back(bundle_atom[UC_GATED][i]); }
            }
            std::vector<GatedGrams> kept_g;
            for (size_t i = 0; i < gated_grams.size(); ++i)
                if (keep[i]) kept_g.push_back(std::move(gated_grams[i]));
            gated_grams.swap(kept_g);
            bundles[UC_GATED].swap(kept_b);
            bundle_atom[UC_GATED].swap(kept_a);
        }
        bool any = !literals.empty();
        std::vector<uint32_t> field_grams, field_masks;
        for (int cls = 0; cls < N_UC; ++cls) {
            if (bundles[cls].empty()) continue;
            any = true;
            DfaGroups groups;
            std::vector<int> too_big;
            // (Capping gated units at the shared-memory budget so that each is wholly resident -- 8 url units instead of 2 at 512
            // rules, 16 instead of 4 at 1 024 -- was measured: no gain at 512 rules, 25 % slower scan at 1 024: candidates walk
            // more units and CTAs stage more images; the cold-row path is not what limits the candidate scan.)
            build_dfa_groups(M.nfa[f], bundles[cls], opt.max_dfa_states, opt.max_unit_table_bytes, (int)kMaxLatchesPerUnit, &groups, &too_big);
            // a bundle no DFA of acceptable size exists for is simulated as a bit-parallel NFA over every request (nfa_bits.hpp):
            // where Rust `regex` would leave its lazy DFA for the PikeVM, not a configuration error
            for (int bi : too_big) {
                const AtomDesc& ad = M.atoms[bundle_atom[cls][bi]];
                std::vector<int> pids;
                std::vector<uint32_t> words;
                for (size_t k = 0; k < ad.nfa_starts.size(); ++k) {
                    const PatternEvent& ev = M.events[ad.event_base + k];
                    pids.push_back(ad.event_base + (int)k);
                    words.push_back(((uint32_t)ev.kind << kEvKindShift) | (uint32_t)ev.atom);   // latch 0: one bundle per unit
                }
                BitsetUnitDesc bd;
                std::string berr;
                if (!build_bitset_unit(M.nfa[f], bundles[cls][bi].starts, pids, words, f, &H.bitset_blob, &bd, berr)) {
                    err = "pattern on http_request." + std::string(kFieldNames[f]) + " needs a DFA larger than " + std::to_string(opt.max_dfa_states) +
                          " states and " + berr + ": " + ad.key;
                    return false;
                }
                H.bitset_units.push_back(bd);
            }
            if (cls == UC_GATED) {
                // a gram leads to the units whose patterns it came from (bit = unit index among the field's gated units, mod kGateWidth)
                for (size_t g = 0; g < groups.dfas.size(); ++g)
                    for (int bi : groups.members[g])
                        for (uint32_t x : gated_grams[bi].grams) { field_grams.push_back(x); field_masks.push_back(1u << (g % kGateWidth[f])); }
            }
            for (size_t g = 0; g < groups.dfas.size(); ++g) {
                Pending pd;
                pd.dfa = std::move(groups.dfas[g]);
                pd.field = f;
                pd.mode = cls == UC_GATED ? UM_CANDIDATES : UM_ALL;
                pd.gate_bit = cls == UC_GATED ? (uint32_t)(g % kGateWidth[f]) : 0u;
                // latch numbering is local to the unit
                pd.latch_of_event.assign(M.events.size(), 0);
                int next_latch = 0;
                for (int bi : groups.members[g]) {
                    const AtomDesc& ad = M.atoms[bundle_atom[cls][bi]];
                    if (!ad.has_latch) continue;
                    for (size_t k = 0; k < ad.nfa_starts.size(); ++k) pd.latch_of_event[ad.event_base + k] = next_latch;
                    ++next_latch;
                }
                pend.push_back(std::move(pd));
            }
        }
        if (!field_grams.empty() || !literals.empty()) {
            std::set<uint32_t> lit_atoms;
            for (auto& l : literals) lit_atoms.insert(l.atom);
            H.n_literal_atoms += (uint32_t)lit_atoms.size();
            gate_in[f].present = true;
            gate_in[f].grams.swap(field_grams);
            gate_in[f].masks.swap(field_masks);
            gate_in[f].literals.swap(literals);
        }
        if (any) H.scanned_fields_mask |= 1u << f;
    }
    // the exact tables of all gated fields have one slot layout (the resolve kernel is one launch over the fields): the wide one,
    // with literal candidate lists, only if some field confirms literals
    for (int f = 0; f < N_FIELDS; ++f)
        if (gate_in[f].present)
            gate_build_tables(gate_in[f].grams, gate_in[f].masks, gate_in[f].literals, H.n_literal_atoms != 0, f == F_URL ? kGateMaxLog2 : kGateMaxLog2 - 2,
                              &H.gate[f]);
    // class maps
    std::vector<uint32_t> cls_offs;
    for (auto& p : pend) {
        cls_offs.push_back((uint32_t)H.arena.size());
        H.arena.insert(H.arena.end(), p.dfa.classmap, p.dfa.classmap + 256);
    }
    H.acc_idx.clear();
    H.end_idx.clear();
    auto emit_events = [&](const Pending& pd, const std::vector<int>& pattern_ids, std::vector<uint32_t>& out_words) {
        std::vector<uint32_t> w;
        for (int pid : pattern_ids) {
            const PatternEvent& ev = M.events[pid];
            w.push_back(((uint32_t)ev.kind << kEvKindShift) | ((uint32_t)pd.latch_of_event[pid] << kEvLatchShift) | (uint32_t)ev.atom);
        }
        std::sort(w.begin(), w.end());  // kind is the most significant field: FIRE, TEST, CLEAR, SET
        out_words.insert(out_words.end(), w.begin(), w.end());
    };
    for (size_t u = 0; u < pend.size(); ++u) {
        const Dfa& d = pend[u].dfa;
        UnitDesc ud;
        memset(&ud, 0, sizeof ud);
        ud.field = (uint32_t)pend[u].field;
        ud.n_classes = (uint32_t)d.n_classes;
        ud.n_states = (uint32_t)d.n_states;
        ud.start_state = (uint32_t)d.start;
        ud.acc_lo = (uint32_t)d.acc_lo;
        pad16(H.arena);
        ud.tbl_off = (uint32_t)H.arena.size();
        ud.cls_off = cls_offs[u];
        const uint8_t* tb = (const uint8_t*)d.trans.data();
        H.arena.insert(H.arena.end(), tb, tb + d.trans.size() * 2);
        ud.acc_base = (uint32_t)H.acc_idx.size();
        for (int s = d.acc_lo; s < d.n_states; ++s) {
            H.acc_idx.push_back((uint32_t)H.acc_events.size());
            emit_events(pend[u], d.acc[s], H.acc_events);
        }
        H.acc_idx.push_back((uint32_t)H.acc_events.size());
        ud.end_base = (uint32_t)H.end_idx.size();
        for (int s = 0; s < d.n_states; ++s) {
            H.end_idx.push_back((uint32_t)H.end_events.size());
            emit_events(pend[u], d.endacc[s], H.end_events);
            if (!d.endacc[s].empty()) ud.end_any = 1;
        }
        H.end_idx.push_back((uint32_t)H.end_events.size());
        ud.hot_states = ud.n_states;
        ud.hot_off = ud.tbl_off;
        ud.mode = pend[u].mode;
        ud.gate_bit = pend[u].gate_bit;
        // absorbing states: once reached nothing can change any more, the field is finished as if it ended there
        ud.abs0 = ud.abs1 = 0xFFFFFFFFu;
        for (int st = 0; st < d.n_states; ++st) {
            bool absorbing = true;
            for (int c = 0; c < d.n_classes && absorbing; ++c) absorbing = d.trans[(size_t)st * d.n_classes + c] == (uint16_t)st;
            if (!absorbing || (st >= d.acc_lo && !d.acc[st].empty())) continue;  // a sticky accepting state keeps firing: not handled early
            if (ud.abs0 == 0xFFFFFFFFu) ud.abs0 = (uint32_t)st;
            else if (ud.abs1 == 0xFFFFFFFFu) ud.abs1 = (uint32_t)st;
        }
        ud.start_end = d.endacc[d.start].empty() ? 0u : 1u;
        ud.has_latch = 0;
        for (int lv : pend[u].latch_of_event) if (lv) ud.has_latch = 1;
        for (size_t a = 0; a < M.atoms.size(); ++a)
            if (M.atoms[a].kind == AtomDesc::STR_PATTERN && M.atoms[a].has_latch && M.atoms[a].field == pend[u].field) ud.has_latch = 1;
        H.units.push_back(ud);
    }
    pad16(H.arena);
    if (H.acc_events.empty()) H.acc_events.push_back(0);
    if (H.end_events.empty()) H.end_events.push_back(0);

    // ---- non-scan atoms ----------------------------------------------------------------
    H.iset_off.assign(1, 0);
    for (auto& s : M.int_sets) {
        H.iset_vals.insert(H.iset_vals.end(), s.begin(), s.end());
        H.iset_off.push_back((uint32_t)H.iset_vals.size());
    }
    for (auto& cs : M.country_sets)
        for (uint32_t w = 0; w < kCountryWords; ++w) {
            uint32_t v = 0;
            for (int b = 0; b < 32; ++b) {
                size_t i = (size_t)w * 32 + b;
                if (i < 676 && cs.test(i)) v |= 1u << b;
            }
            H.cset_words.push_back(v);
        }
    std::vector<uint32_t> iexpr_off;
    for (auto& prog : M.int_progs) {
        iexpr_off.push_back((uint32_t)H.iexpr.size());
        H.iexpr.insert(H.iexpr.end(), prog.begin(), prog.end());
    }
    if (H.iexpr.empty()) H.iexpr.push_back(0);
    for (uint32_t a = 0; a < H.n_atoms; ++a) {
        const AtomDesc& d = M.atoms[a];
        if (d.kind == AtomDesc::STR_PATTERN) continue;
        // A predicate no rule refers to (its rule folded to a constant, e.g. behind an operand that always errors) is not
        // evaluated at all -- like the unreferenced string atoms, which are not scanned.  This is more than an economy: the
        // two-atom verdict shortcut (atom_sig, above) relies on every atom that can be TRUE being mentioned by some rule; an
        // unmentioned one has an empty signature, which is "disjoint" from anything, including the all-ones signature of an
        // atom that a rule true on the all-false vector negates.
        if (!((H.care[a >> 5] >> (a & 31)) & 1u)) continue;
        NsAtom n;
        memset(&n, 0, sizeof n);
        n.kind = d.kind;
        n.atom = a;
        n.feat = (uint32_t)std::max(0, d.feat);
        n.op = (uint32_t)d.op;
        n.cval = d.cval;
        n.set_id = (uint32_t)std::max(0, d.set_id);
        if (d.kind == AtomDesc::INT_EXPR) {
            // the program's tokens go to the flat token array once per program; features it reads must be supplied
            n.set_id = iexpr_off[d.set_id];
            const std::vector<int64_t>& prog = M.int_progs[d.set_id];
            for (size_t i = 0; i < prog.size(); ++i) {
                const uint32_t op = (uint32_t)((uint64_t)prog[i] >> 56);
                if (op == IT_CONST64) { ++i; continue; }
                if (op != IT_FEAT) continue;
                const int f = (int)(prog[i] & 0xFF);
                if (f == IF_PORT) H.needs_port = true;
                else if (f == IF_ASN) H.needs_geo_cols = true;
                else len_feat_used[f - IF_LEN0] = true;
            }
        }
        if (d.kind == AtomDesc::FIELD_CMP) {
            n.feat = (uint32_t)d.field;
            n.set_id = (uint32_t)d.feat;
            H.scanned_fields_mask |= (1u << d.field) | (1u << d.feat);   // both fields' bytes are read (by the per-request kernel)
        }
        H.ns_atoms.push_back(n);
        if (d.kind == AtomDesc::INT_CMP || d.kind == AtomDesc::INT_SET) {
            if (d.feat == IF_PORT) H.needs_port = true;
            else if (d.feat == IF_ASN) H.needs_geo_cols = true;
            else len_feat_used[d.feat - IF_LEN0] = true;
        }
        if (d.kind == AtomDesc::IP_SET) H.needs_ip = true;
        if (d.kind == AtomDesc::COUNTRY_SET) H.needs_geo_cols = true;
    }
    // group the predicates by feature and derive the per-feature quick reject (most requests satisfy none of them)
    {
        // integer predicates by feature, then the ip / country sets, then the rare kinds (INT_EXPR, FIELD_CMP) at the very end
        auto group = [](const NsAtom& a) {
            return (a.kind == AtomDesc::INT_CMP || a.kind == AtomDesc::INT_SET) ? (int)a.feat
                   : (a.kind == AtomDesc::INT_EXPR || a.kind == AtomDesc::FIELD_CMP) ? (int)N_INT_FEATS + 1 : (int)N_INT_FEATS;
        };
        std::stable_sort(H.ns_atoms.begin(), H.ns_atoms.end(), [&](const NsAtom& x, const NsAtom& y) { return group(x) < group(y); });
        for (int g = 0; g <= N_INT_FEATS + 1; ++g) H.ns_begin[g] = 0;
        H.n_rare = 0;
        for (auto& a : H.ns_atoms) {
            if (group(a) > N_INT_FEATS) { H.n_rare++; continue; }
            H.ns_begin[group(a) + 1]++;
        }
        if (H.n_rare > 64) {
            err = "more than 64 integer-expression / field-comparison predicates";
            return false;
        }
        for (int g = 0; g <= N_INT_FEATS; ++g) H.ns_begin[g + 1] += H.ns_begin[g];
        for (int f = 0; f < N_INT_FEATS; ++f) {
            int64_t lo = INT64_MIN, hi = INT64_MAX, vmin = INT64_MAX, vmax = INT64_MIN;
            for (uint32_t i = H.ns_begin[f]; i < H.ns_begin[f + 1]; ++i) {
                const NsAtom& a = H.ns_atoms[i];
                if (a.kind == AtomDesc::INT_SET) {
                    for (uint32_t k = H.iset_off[a.set_id]; k < H.iset_off[a.set_id + 1]; ++k) { vmin = std::min(vmin, H.iset_vals[k]); vmax = std::max(vmax, H.iset_vals[k]); }
                    continue;
                }
                switch (a.op) {
                    case CMP_LT: lo = std::max(lo, a.cval); break;
                    case CMP_LE: lo = std::max(lo, a.cval == INT64_MAX ? a.cval : a.cval + 1); if (a.cval == INT64_MAX) hi = INT64_MIN; break;
                    case CMP_GT: hi = std::min(hi, a.cval); break;
                    case CMP_GE: hi = std::min(hi, a.cval == INT64_MIN ? a.cval : a.cval - 1); if (a.cval == INT64_MIN) lo = INT64_MAX; break;
                    case CMP_EQ: vmin = std::min(vmin, a.cval); vmax = std::max(vmax, a.cval); break;
                    default: lo = INT64_MAX; hi = INT64_MIN; break;  // (CMP_NE is lowered to NOT EQ; be safe: no quick reject)
                }
            }
            H.ns_lo[f] = lo; H.ns_hi[f] = hi; H.ns_vmin[f] = vmin; H.ns_vmax[f] = vmax;
        }
    }
    if (M.ip_sets.size() > 32) {
        err = "more than 32 distinct Ip lists referenced by rules";
        return false;
    }

    // ---- offset slots: scanned fields, length features, the user-agent gate --------
    for (int f = 0; f < N_FIELDS; ++f) {
        bool need = (H.scanned_fields_mask >> f) & 1 || len_feat_used[f] || (opt.eval_gates && f == F_USER_AGENT);
        if (need) H.field_slot[f] = (int)H.n_slots++;
    }
    for (auto& u : H.units) u.field_slot = (uint32_t)H.field_slot[u.field];

    // ---- LPM tables (ip lists + geoip) -------------------------------------------------
    bool want_geo = !geo_mmdb.empty();
    if (!M.ip_sets.empty() || want_geo) {
        if (!build_lpm(M.ip_sets, geo_mmdb, &H.lpm, err)) return false;
    }
    return true;
}

void build_unit_images(const HostProgram& H, size_t budget, std::vector<uint8_t>* image, std::vector<UnitDesc>* units) {
    *units = H.units;
    image->clear();
    for (size_t u = 0; u < H.units.size(); ++u) {
        while (image->size() % 256) image->push_back(0);
        UnitDesc& ud = (*units)[u];
        const size_t base = image->size();
        ud.img_off = (uint32_t)base;
        image->insert(image->end(), H.arena.begin() + H.units[u].cls_off, H.arena.begin() + H.units[u].cls_off + 256);
        const size_t C = ud.n_classes, row = C * 2;
        // fixed parts: class map, trap row, acc1 (upper bound: every accepting state hot), end1, alignment slack
        size_t fixed = 256 + row + 2 * (size_t)(ud.n_states - ud.acc_lo) + 2 * (size_t)ud.n_states + 64;
        size_t give = ud.n_states;
        if (fixed + give * row > budget) {
            give = budget > fixed ? (budget - fixed) / row : 0;
            // acc1 / end1 only cover hot states: recompute with the real sizes (they shrink with `give`)
            for (;;) {
                size_t acc_hot = give > ud.acc_lo ? give - ud.acc_lo : 0;
                size_t need = 256 + row + 2 * acc_hot + 2 * give + 64 + give * row;
                if (need <= budget || give == 0) break;
                --give;
            }
            size_t more = give;
            for (;;) {  // grow back while it fits
                size_t g2 = more + 1;
                if (g2 > ud.n_states) break;
                size_t acc_hot = g2 > ud.acc_lo ? g2 - ud.acc_lo : 0;
                if (256 + row + 2 * acc_hot + 2 * g2 + 64 + g2 * row > budget) break;
                more = g2;
            }
            give = more;
        }
        ud.hot_states = (uint32_t)give;
        ud.lim = std::min(ud.hot_states, ud.acc_lo);
        while (image->size() % 16) image->push_back(0);
        ud.hot_off = (uint32_t)(image->size() - base);
        const uint16_t* src = reinterpret_cast<const uint16_t*>(H.arena.data() + H.units[u].tbl_off);
        const uint16_t trap = (uint16_t)ud.hot_states;
        for (size_t s = 0; s < give; ++s)
            for (size_t c = 0; c < C; ++c) {
                uint16_t t = src[s * C + c];
                if (t >= ud.hot_states) t = trap;  // only cold states trap; accepting states stay on the fast path
                image->push_back((uint8_t)(t & 0xFF));
                image->push_back((uint8_t)(t >> 8));
            }
        for (size_t c = 0; c < C; ++c) {  // trap row: absorbing
            image->push_back((uint8_t)(trap & 0xFF));
            image->push_back((uint8_t)(trap >> 8));
        }
        // acc1: one-atom FIRE lists resolved without leaving shared memory
        while (image->size() % 4) image->push_back(0);
        ud.acc1_off = (uint32_t)(image->size() - base);
        for (uint32_t st = ud.acc_lo; st < ud.hot_states; ++st) {
            uint32_t ci = ud.acc_base + st - ud.acc_lo;
            uint32_t a = H.acc_idx[ci], b = H.acc_idx[ci + 1];
            uint16_t v = 0xFFFF;
            if (b - a == 1 && (H.acc_events[a] >> kEvKindShift) == 0) v = (uint16_t)(H.acc_events[a] & kEvAtomMask);
            image->push_back((uint8_t)(v & 0xFF));
            image->push_back((uint8_t)(v >> 8));
        }
        // end1: end-of-field events of hot states
        while (image->size() % 4) image->push_back(0);
        ud.end1_off = (uint32_t)(image->size() - base);
        for (uint32_t st = 0; st < ud.hot_states; ++st) {
            uint32_t ci = ud.end_base + st;
            uint32_t a = H.end_idx[ci], b = H.end_idx[ci + 1];
            uint16_t v = 0xFFFF;
            if (a == b) v = 0xFFFE;
            else if (b - a == 1 && (H.end_events[a] >> kEvKindShift) == 0) v = (uint16_t)(H.end_events[a] & kEvAtomMask);
            image->push_back((uint8_t)(v & 0xFF));
            image->push_back((uint8_t)(v >> 8));
        }
        while (image->size() % 16) image->push_back(0);
        ud.img_bytes = (uint32_t)(image->size() - base);
    }
    while (image->size() % 256) image->push_back(0);
}

}  // namespace pgw
