// Gram extraction for the candidate gate (see gate.hpp).
#include "gate.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>

namespace pgw {
namespace {

constexpr int kWild = 256;     // "any byte" symbol in a padded prefix
constexpr int kPrefix = 5;     // padded prefixes are 5 symbols long: windows m[0..4) and m[1..5)
constexpr size_t kMaxPrefixes = 4096;

inline bool is_folded(int v) { return !(v & 0x20); }  // gate_fold clears bit 5

// Enumerates the padded 5-byte prefixes (forward) or suffixes (backward) of a pattern's matches over folded bytes.
struct Walker {
    const Nfa& nfa;
    const bool backward;
    int start_node = -1;
    std::vector<int> stamp;
    int gen = 0;
    std::vector<int> stack;
    std::vector<std::vector<int>> radj;   // backward: predecessors within the pattern's sub-graph
    std::vector<int> char_nodes;          // backward: byte-consuming nodes of the sub-graph
    std::vector<std::array<int, kPrefix>> out;
    bool overflow = false;

    Walker(const Nfa& n, int start, bool back) : nfa(n), backward(back), start_node(start), stamp(n.nodes.size(), 0) {
        if (!backward) return;
        radj.resize(n.nodes.size());
        std::vector<char> seen(n.nodes.size(), 0);
        std::vector<int> st(1, start);
        while (!st.empty()) {
            int x = st.back();
            st.pop_back();
            if (x < 0 || seen[x]) continue;
            seen[x] = 1;
            const NfaNode& nd = nfa.nodes[x];
            if (nd.kind == N_MATCH) continue;
            if (nd.kind == N_CHAR) char_nodes.push_back(x);
            if (nd.out >= 0) { radj[nd.out].push_back(x); st.push_back(nd.out); }
            if (nd.kind == N_SPLIT && nd.out1 >= 0) { radj[nd.out1].push_back(x); st.push_back(nd.out1); }
        }
    }

    // forward: epsilon closure with every assertion treated as satisfied (a superset of the real language);
    //   `chars` = byte-consuming nodes reached, `edge` = a MATCH node is reachable (the match may end here)
    // backward: reverse closure; `chars` = the closed node set itself, `edge` = the pattern's start node is in it
    void closure(const std::vector<int>& from, std::vector<int>& chars, bool& edge) {
        ++gen;
        chars.clear();
        edge = false;
        stack.assign(from.begin(), from.end());
        while (!stack.empty()) {
            int n = stack.back();
            stack.pop_back();
            if (n < 0 || stamp[n] == gen) continue;
            stamp[n] = gen;
            const NfaNode& nd = nfa.nodes[n];
            if (backward) {
                chars.push_back(n);
                if (n == start_node) edge = true;
                for (int p : radj[n])
                    if (nfa.nodes[p].kind != N_CHAR) stack.push_back(p);  // epsilon predecessors only
                continue;
            }
            switch (nd.kind) {
                case N_CHAR: chars.push_back(n); break;
                case N_MATCH: edge = true; break;
                case N_JUMP:
                case N_ASSERT: stack.push_back(nd.out); break;
                case N_SPLIT:
                    stack.push_back(nd.out);
                    stack.push_back(nd.out1);
                    break;
            }
        }
        std::sort(chars.begin(), chars.end());
    }

    // forward: pads the tail; backward: sequences are built last byte first and emitted in text order, padded in front
    void emit(const std::array<int, kPrefix>& seq, int len) {
        std::array<int, kPrefix> o;
        o.fill(kWild);
        if (!backward) for (int k = 0; k < len; ++k) o[k] = seq[k];
        else for (int k = 0; k < len; ++k) o[kPrefix - 1 - k] = seq[k];
        out.push_back(o);
        if (out.size() > kMaxPrefixes) overflow = true;
    }

    void dfs(int depth, const std::vector<int>& cur, std::array<int, kPrefix>& seq) {
        std::vector<int> next, nset;
        std::vector<char> in_cur;
        if (backward) {
            in_cur.assign(nfa.nodes.size(), 0);
            for (int n : cur) in_cur[n] = 1;
        }
        for (int v = 0; v < 256 && !overflow; ++v) {
            if (!is_folded(v)) continue;
            const int alt = v | 0x20;  // the other byte folding to v
            next.clear();
            if (!backward) {
                for (int c : cur) {
                    const ByteSet& bs = nfa.sets[nfa.nodes[c].set];
                    if (bs.test((unsigned)v) || (alt >= 0 && bs.test((unsigned)alt))) next.push_back(nfa.nodes[c].out);
                }
            } else {
                for (int c : char_nodes) {
                    if (!in_cur[nfa.nodes[c].out]) continue;
                    const ByteSet& bs = nfa.sets[nfa.nodes[c].set];
                    if (bs.test((unsigned)v) || (alt >= 0 && bs.test((unsigned)alt))) next.push_back(c);
                }
            }
            if (next.empty()) continue;
            bool edge = false;
            closure(next, nset, edge);
            seq[depth] = v;
            const int d1 = depth + 1;
            if (d1 == kPrefix) {
                emit(seq, d1);
                continue;
            }
            if (edge) emit(seq, d1);  // the match may end (begin) here: what follows (precedes) is arbitrary
            bool more = !nset.empty();
            if (backward && more) {
                // anything byte-consuming before this point?
                more = false;
                std::vector<char> in(nfa.nodes.size(), 0);
                for (int n : nset) in[n] = 1;
                for (int c : char_nodes)
                    if (in[nfa.nodes[c].out]) { more = true; break; }
            }
            if (more) {
                std::vector<int> copy = nset;  // `nset` is reused by the recursion's siblings
                dfs(d1, copy, seq);
            }
        }
    }
};

// number of concrete grams a 4-symbol sequence expands to (128 folded values per wildcard)
size_t expansion(const std::array<int, 4>& s) {
    size_t n = 1;
    for (int x : s)
        if (x == kWild) n *= 128;
    return n;
}

void expand(const std::array<int, 4>& s, std::vector<uint32_t>* out) {
    std::vector<uint32_t> acc(1, 0u);
    for (int k = 0; k < 4; ++k) {
        std::vector<uint32_t> nx;
        if (s[k] != kWild) {
            for (uint32_t a : acc) nx.push_back(a | ((uint32_t)s[k] << (8 * k)));
        } else {
            for (uint32_t a : acc)
                for (int v = 0; v < 256; ++v)
                    if (is_folded(v)) nx.push_back(a | ((uint32_t)v << (8 * k)));
        }
        acc.swap(nx);
    }
    out->insert(out->end(), acc.begin(), acc.end());
}

}  // namespace

bool pattern_is_start_anchored(const Nfa& nfa, int start) {
    // closure that refuses to cross start-of-text assertions: anchored iff nothing byte-consuming (or a match) is left
    std::vector<int> stack(1, start);
    std::vector<char> seen(nfa.nodes.size(), 0);
    while (!stack.empty()) {
        int n = stack.back();
        stack.pop_back();
        if (n < 0 || seen[n]) continue;
        seen[n] = 1;
        const NfaNode& nd = nfa.nodes[n];
        switch (nd.kind) {
            case N_CHAR:
            case N_MATCH: return false;
            case N_JUMP: stack.push_back(nd.out); break;
            case N_SPLIT:
                stack.push_back(nd.out);
                stack.push_back(nd.out1);
                break;
            case N_ASSERT:
                if (nd.assert_kind != A_BOL_TEXT) stack.push_back(nd.out);
                break;
        }
    }
    return true;
}

bool gate_grams_for_pattern(const Nfa& nfa, int start, size_t cap, std::vector<uint32_t>* out) {
    typedef std::vector<std::array<int, 4>> Seqs;
    auto uniq = [](Seqs& v) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    };
    auto total = [&](const Seqs& v) {
        size_t n = 0;
        for (auto& s : v) {
            n += expansion(s);
            if (n > (cap << 4)) break;
        }
        return n;
    };
    // windows tied to the START of a match (gate.hpp): A = m[0..4), B = (any, m[0..3)) or m[1..5)
    Seqs best_a, best_b;
    size_t best = (size_t)-1;
    {
        Walker W(nfa, start, false);
        std::vector<int> chars;
        bool match = false;
        W.closure(std::vector<int>(1, start), chars, match);
        if (match) return false;         // matches the empty string: nothing to look for
        if (chars.empty()) return true;  // matches nothing: no grams needed
        std::array<int, kPrefix> seq;
        seq.fill(kWild);
        W.dfs(0, chars, seq);
        if (!W.overflow) {
            Seqs A, B1, B2;
            for (auto& s : W.out) {
                A.push_back({s[0], s[1], s[2], s[3]});
                B1.push_back({kWild, s[0], s[1], s[2]});
                B2.push_back({s[1], s[2], s[3], s[4]});
            }
            uniq(A);
            uniq(B1);
            uniq(B2);
            const size_t na = total(A), nb1 = total(B1), nb2 = total(B2);
            best = na + std::min(nb1, nb2);
            best_a = A;
            best_b = nb2 <= nb1 ? B2 : B1;
        }
    }
    // windows tied to the END of a match (patterns that begin with a gap but end in a literal): with t = the last five
    // bytes and q the end position, q even -> window [q-4, q) = t[1..5); q odd -> [q-3, q+1) = (t[2..5), any) or
    // [q-5, q-1) = t[0..4)
    {
        Walker W(nfa, start, true);
        int match_node = -1;
        {
            // the pattern's MATCH node: forward search
            std::vector<char> seen(nfa.nodes.size(), 0);
            std::vector<int> st(1, start);
            while (!st.empty() && match_node < 0) {
                int x = st.back();
                st.pop_back();
                if (x < 0 || seen[x]) continue;
                seen[x] = 1;
                const NfaNode& nd = nfa.nodes[x];
                if (nd.kind == N_MATCH) { match_node = x; break; }
                st.push_back(nd.out);
                if (nd.kind == N_SPLIT) st.push_back(nd.out1);
            }
        }
        if (match_node >= 0) {
            std::vector<int> set0;
            bool at_start = false;
            W.closure(std::vector<int>(1, match_node), set0, at_start);
            std::array<int, kPrefix> seq;
            seq.fill(kWild);
            W.dfs(0, set0, seq);
            if (!W.overflow && !W.out.empty()) {
                Seqs A, B1, B2;
                for (auto& t : W.out) {
                    A.push_back({t[1], t[2], t[3], t[4]});
                    B1.push_back({t[2], t[3], t[4], kWild});
                    B2.push_back({t[0], t[1], t[2], t[3]});
                }
                uniq(A);
                uniq(B1);
                uniq(B2);
                const size_t na = total(A), nb1 = total(B1), nb2 = total(B2);
                if (na + std::min(nb1, nb2) < best) {
                    best = na + std::min(nb1, nb2);
                    best_a = A;
                    best_b = nb2 <= nb1 ? B2 : B1;
                }
            }
        }
    }
    if (best > cap) return false;
    for (auto& s : best_a) expand(s, out);
    for (auto& s : best_b) expand(s, out);
    return true;
}

namespace {

struct LangWalker {
    const Nfa& nfa;
    std::vector<LitString>* out;
    bool fail = false;
    int steps = 0;

    // `ended`: a `$` has been crossed -- nothing may be consumed any more
    void walk(int n, LitString cur, bool ended, int eps_depth) {
        if (fail) return;
        if (++steps > 200000 || eps_depth > 4096) { fail = true; return; }
        if (n < 0) return;
        const NfaNode& nd = nfa.nodes[n];
        switch (nd.kind) {
            case N_MATCH:
                cur.anch_end = ended;
                out->push_back(cur);
                if (out->size() > kLitMaxStrings * 4) fail = true;   // duplicates are removed later; a real blow-up stops here
                return;
            case N_JUMP: walk(nd.out, cur, ended, eps_depth + 1); return;
            case N_SPLIT:
                walk(nd.out, cur, ended, eps_depth + 1);
                walk(nd.out1, cur, ended, eps_depth + 1);
                return;
            case N_ASSERT:
                if (nd.assert_kind == A_BOL_TEXT) {
                    if (!cur.bytes.empty()) return;   // `^` after a consumed byte: this path never matches
                    cur.anch_start = true;
                    walk(nd.out, cur, ended, eps_depth + 1);
                } else if (nd.assert_kind == A_EOL_TEXT) {
                    walk(nd.out, cur, true, eps_depth + 1);
                } else fail = true;   // line anchors, word boundaries: left to the automata
                return;
            case N_CHAR: {
                if (ended) return;    // a byte after `$`: this path never matches
                if (cur.bytes.size() >= kLitMaxLen) { fail = true; return; }
                const ByteSet& bs = nfa.sets[nd.set];
                int members[5], nm = 0;
                for (int v = 0; v < 256 && nm <= 4; ++v)
                    if (bs.test((unsigned)v)) members[nm++] = v;
                if (nm == 0) return;
                if (nm > 4) { fail = true; return; }
                const size_t k = cur.bytes.size();
                auto is_letter = [](int v) { return (v >= 'A' && v <= 'Z') || (v >= 'a' && v <= 'z'); };
                if (nm == 2 && is_letter(members[0]) && members[1] == (members[0] ^ 0x20)) {
                    LitString nx = cur;
                    nx.bytes.push_back((char)(members[0] & 0xDF));
                    nx.ci_mask |= 1ull << k;
                    walk(nd.out, nx, false, 0);
                    return;
                }
                for (int i = 0; i < nm; ++i) {
                    LitString nx = cur;
                    nx.bytes.push_back((char)members[i]);
                    walk(nd.out, nx, false, 0);
                }
                return;
            }
        }
    }
};

}  // namespace

bool gate_finite_language(const Nfa& nfa, int start, std::vector<LitString>* out) {
    out->clear();
    LangWalker W{nfa, out};
    W.walk(start, LitString(), false, 0);
    if (W.fail || out->empty()) return false;
    std::sort(out->begin(), out->end());
    out->erase(std::unique(out->begin(), out->end()), out->end());
    if (out->size() > kLitMaxStrings) return false;
    for (const LitString& s : *out)
        if (s.bytes.size() < kLitMinLen) return false;
    return true;
}

void gate_grams_for_literal(const LitString& s, std::vector<std::pair<uint32_t, int>>* out) {
    const size_t n = s.bytes.size();
    auto f = [&](size_t k) -> int { return k < n ? ((unsigned char)s.bytes[k] & 0xDF) : kWild; };
    auto emit = [&](std::array<int, 4> w, int delta) {
        std::vector<uint32_t> g;
        expand(w, &g);
        for (uint32_t x : g) out->emplace_back(x, delta);
    };
    // a literal anchored at the field end may be followed by the next field's bytes: the padding is a wildcard either way
    emit({f(0), f(1), f(2), f(3)}, 0);                       // the match starts at an even position: window = m[0..4)
    if (n >= 5) emit({f(1), f(2), f(3), f(4)}, -1);          // odd position: window [p + 1, p + 5) = m[1..5)
    else emit({kWild, f(0), f(1), f(2)}, +1);                // odd position: window [p - 1, p + 3) = (any, m[0..3))
}

void gate_build_tables(const std::vector<uint32_t>& grams, const std::vector<uint32_t>& masks, const std::vector<GateLiteral>& literals,
                       bool wide_slots, uint32_t max_log2, GateTables* out) {
    struct Ent { uint32_t mask = 0; std::vector<uint32_t> cand; };
    std::map<uint32_t, Ent> byg;
    for (size_t i = 0; i < grams.size(); ++i) byg[grams[i]].mask |= masks[i];
    GateTables& T = *out;
    T = GateTables();
    T.present = true;
    for (size_t li = 0; li < literals.size(); ++li) {
        const LitString& ls = literals[li].str;
        LitDesc d;
        d.off = (uint32_t)T.lit_bytes.size();
        d.len = (uint16_t)ls.bytes.size();
        d.flags = (uint16_t)((ls.anch_start ? 1 : 0) | (ls.anch_end ? 2 : 0));
        d.atom = literals[li].atom;
        d.pad = 0;
        d.pad2 = 0;
        d.ci_mask = ls.ci_mask;
        T.lits.push_back(d);
        T.lit_bytes.insert(T.lit_bytes.end(), ls.bytes.begin(), ls.bytes.end());
        std::vector<std::pair<uint32_t, int>> lg;
        gate_grams_for_literal(ls, &lg);
        for (auto& pr : lg) {
            const uint32_t c = ((uint32_t)li << 2) | (uint32_t)(pr.second + 1);
            std::vector<uint32_t>& v = byg[pr.first].cand;
            if (std::find(v.begin(), v.end(), c) == v.end()) v.push_back(c);
        }
    }
    while (T.lit_bytes.size() % 16) T.lit_bytes.push_back(0);
    T.n_grams = (uint32_t)byg.size();
    // level 1: two bits per gram in one word; density <= 1/64 -> false-positive rate <= 2.5e-4 per window
    uint32_t k = 12;
    while (k < max_log2 && k < kGateMaxLog2 && ((size_t)1 << k) < byg.size() * 128) ++k;
    T.k1 = k;
    T.b1.assign(((size_t)1 << T.k1) / 32, 0u);
    // level 2: load factor <= 1/2
    T.kt = 4;
    while (((size_t)1 << T.kt) < byg.size() * 2) ++T.kt;
    const uint32_t W = (wide_slots || !literals.empty()) ? 4u : 2u;
    T.slot_words = W;
    T.slots.assign(((size_t)W << T.kt), 0u);
    const uint32_t tm = (1u << T.kt) - 1u;
    for (auto& kv : byg) {
        const uint32_t g = kv.first;
        gate_l1_set(T.b1.data(), T.k1, g);
        uint32_t s = (g * kGateHash2) >> (32 - T.kt);
        while (T.slots[W * s + 1] != 0 || (W == 4 && T.slots[W * s + 3] != 0)) s = (s + 1) & tm;
        T.slots[W * s] = g;
        T.slots[W * s + 1] = kv.second.mask;
        if (W == 4) {
            T.slots[W * s + 2] = (uint32_t)T.lit_cand.size();
            T.slots[W * s + 3] = (uint32_t)kv.second.cand.size();
        }
        T.lit_cand.insert(T.lit_cand.end(), kv.second.cand.begin(), kv.second.cand.end());
    }
    if (T.lit_cand.empty()) T.lit_cand.push_back(0);
    if (T.lits.empty()) { LitDesc d; memset(&d, 0, sizeof d); T.lits.push_back(d); }
    if (T.lit_bytes.empty()) T.lit_bytes.assign(16, 0);
}

}  // namespace pgw
