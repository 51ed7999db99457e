// Subset construction / minimisation / grouping (see dfa.hpp).
#include "dfa.hpp"

#include <algorithm>
#include <cstring>
#include <map>
#include <unordered_map>

namespace pgw {
namespace {

static inline bool is_word_byte(int c) {
    return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}

struct VecHash {
    size_t operator()(const std::vector<int>& v) const {
        uint64_t h = 1469598103934665603ull;
        for (int x : v) {
            h ^= (uint32_t)x;
            h *= 1099511628211ull;
        }
        return (size_t)h;
    }
};

enum { F_AT_START = 1, F_PREV_WORD = 2, F_PREV_NL = 4 };

struct Builder {
    const Nfa& nfa;
    const std::vector<int>& starts;
    std::vector<int> stamp;
    int gen = 0;
    std::vector<int> stack;
    bool use_word = false, use_line = false, use_bol = false;

    Builder(const Nfa& n, const std::vector<int>& s) : nfa(n), starts(s), stamp(n.nodes.size(), 0) {}

    // Expand `kern` under the look-around context; collect CHAR nodes and matched patterns.
    void closure(const std::vector<int>& kern, int flags, int next /* byte or -1 for end */, std::vector<int>& chars,
                 std::vector<int>& matched) {
        ++gen;
        chars.clear();
        matched.clear();
        stack.clear();
        for (int n : kern) stack.push_back(n);
        bool prev_word = flags & F_PREV_WORD;
        bool next_word = next >= 0 && is_word_byte(next);
        while (!stack.empty()) {
            int n = stack.back();
            stack.pop_back();
            if (n < 0 || stamp[n] == gen) continue;
            stamp[n] = gen;
            const NfaNode& nd = nfa.nodes[n];
            switch (nd.kind) {
                case N_CHAR: chars.push_back(n); break;
                case N_MATCH: matched.push_back(nd.pattern); break;
                case N_JUMP: stack.push_back(nd.out); break;
                case N_SPLIT:
                    stack.push_back(nd.out);
                    stack.push_back(nd.out1);
                    break;
                case N_ASSERT: {
                    bool ok = false;
                    switch (nd.assert_kind) {
                        case A_BOL_TEXT: ok = flags & F_AT_START; break;
                        case A_EOL_TEXT: ok = next < 0; break;
                        case A_BOL_LINE: ok = (flags & F_AT_START) || (flags & F_PREV_NL); break;
                        case A_EOL_LINE: ok = next < 0 || next == '\n'; break;
                        case A_WORD_B: ok = prev_word != next_word; break;
                        case A_NOT_WORD_B: ok = prev_word == next_word; break;
                        case A_WORD_START: ok = !prev_word && next_word; break;
                        case A_WORD_END: ok = prev_word && !next_word; break;
                        case A_WORD_START_HALF: ok = !prev_word; break;
                        case A_WORD_END_HALF: ok = !next_word; break;
                    }
                    if (ok) stack.push_back(nd.out);
                    break;
                }
            }
        }
        std::sort(matched.begin(), matched.end());
        matched.erase(std::unique(matched.begin(), matched.end()), matched.end());
    }
};

}  // namespace

bool build_dfa(const Nfa& nfa, const std::vector<int>& starts, int max_raw_states, Dfa* out) {
    Builder B(nfa, starts);

    // ---- reachable nodes, assertion kinds in use, byte classes -------------
    std::vector<char> seen(nfa.nodes.size(), 0);
    std::vector<int> st(starts.begin(), starts.end());
    std::vector<int> used_sets;
    while (!st.empty()) {
        int n = st.back();
        st.pop_back();
        if (n < 0 || seen[n]) continue;
        seen[n] = 1;
        const NfaNode& nd = nfa.nodes[n];
        if (nd.kind == N_CHAR) used_sets.push_back(nd.set);
        if (nd.kind == N_ASSERT) {
            if (assert_looks_at_words(nd.assert_kind)) B.use_word = true;
            if (nd.assert_kind == A_BOL_LINE || nd.assert_kind == A_EOL_LINE) B.use_line = true;
            if (nd.assert_kind == A_BOL_TEXT || nd.assert_kind == A_BOL_LINE) B.use_bol = true;
        }
        if (nd.kind != N_MATCH) st.push_back(nd.out);
        if (nd.kind == N_SPLIT) st.push_back(nd.out1);
    }
    std::sort(used_sets.begin(), used_sets.end());
    used_sets.erase(std::unique(used_sets.begin(), used_sets.end()), used_sets.end());

    // raw byte classes: bytes with identical membership in every used set (+ word / newline distinctions)
    int raw_class_of[256];
    std::vector<int> rep;  // representative byte per raw class
    {
        std::map<std::vector<uint8_t>, int> sig2cls;
        for (int c = 0; c < 256; ++c) {
            std::vector<uint8_t> sig;
            sig.reserve(used_sets.size() + 2);
            for (int s : used_sets) sig.push_back(nfa.sets[s].test(c));
            if (B.use_word) sig.push_back(is_word_byte(c));
            if (B.use_line) sig.push_back(c == '\n');
            auto it = sig2cls.find(sig);
            if (it == sig2cls.end()) {
                it = sig2cls.emplace(sig, (int)rep.size()).first;
                rep.push_back(c);
            }
            raw_class_of[c] = it->second;
        }
    }
    const int C = (int)rep.size();

    // ---- subset construction --------------------------------------------
    struct Raw {
        std::vector<int> kern;
        int flags;
        std::vector<int> acc;
    };
    std::vector<Raw> states;
    std::unordered_map<std::vector<int>, int, VecHash> index;
    std::vector<int> trans;  // states x C
    auto intern = [&](std::vector<int>& kern, int flags, const std::vector<int>& acc) -> int {
        std::sort(kern.begin(), kern.end());
        kern.erase(std::unique(kern.begin(), kern.end()), kern.end());
        std::vector<int> key;
        key.reserve(kern.size() + acc.size() + 2);
        key.push_back(flags);
        key.insert(key.end(), kern.begin(), kern.end());
        key.push_back(-1);
        key.insert(key.end(), acc.begin(), acc.end());
        auto it = index.find(key);
        if (it != index.end()) return it->second;
        int id = (int)states.size();
        index.emplace(std::move(key), id);
        states.push_back(Raw{kern, flags, acc});
        return id;
    };

    {
        std::vector<int> k0(starts.begin(), starts.end());
        int f0 = B.use_bol ? F_AT_START : 0;
        intern(k0, f0, {});
    }
    std::vector<int> chars, matched, next_kern;
    std::vector<std::vector<int>> endacc;
    for (size_t s = 0; s < states.size(); ++s) {
        if ((int)states.size() > max_raw_states) return false;
        trans.resize((s + 1) * C);
        // end-of-haystack acceptance
        {
            Raw cur = states[s];
            B.closure(cur.kern, cur.flags, -1, chars, matched);
            endacc.push_back(matched);
        }
        for (int c = 0; c < C; ++c) {
            const Raw& cur = states[s];
            int byte = rep[c];
            B.closure(cur.kern, cur.flags, byte, chars, matched);
            next_kern.clear();
            for (int n : chars)
                if (nfa.sets[nfa.nodes[n].set].test(byte)) next_kern.push_back(nfa.nodes[n].out);
            for (int r : starts) next_kern.push_back(r);  // unanchored search: every position may start a match
            int nf = 0;
            if (B.use_word && is_word_byte(byte)) nf |= F_PREV_WORD;
            if (B.use_line && byte == '\n') nf |= F_PREV_NL;
            std::vector<int> acc = matched;
            int t = intern(next_kern, nf, acc);
            trans[s * C + c] = t;
        }
    }
    const int N = (int)states.size();
    if (N > max_raw_states) return false;

    // ---- Moore minimisation ---------------------------------------------
    std::vector<int> block(N);
    int n_blocks = 0;
    {
        std::map<std::pair<std::vector<int>, std::vector<int>>, int> init;
        for (int s = 0; s < N; ++s) {
            auto key = std::make_pair(states[s].acc, endacc[s]);
            auto it = init.find(key);
            if (it == init.end()) it = init.emplace(key, n_blocks++).first;
            block[s] = it->second;
        }
    }
    for (;;) {
        std::unordered_map<std::vector<int>, int, VecHash> sigs;
        std::vector<int> nb(N);
        std::vector<int> sig(C + 1);
        int count = 0;
        for (int s = 0; s < N; ++s) {
            sig[0] = block[s];
            for (int c = 0; c < C; ++c) sig[c + 1] = block[trans[s * C + c]];
            auto it = sigs.find(sig);
            if (it == sigs.end()) it = sigs.emplace(sig, count++).first;
            nb[s] = it->second;
        }
        bool stable = count == n_blocks;
        block.swap(nb);
        n_blocks = count;
        if (stable) break;
    }

    // ---- renumber: BFS order from start, accepting (acc non-empty) states last
    std::vector<int> rep_state(n_blocks, -1);
    for (int s = 0; s < N; ++s)
        if (rep_state[block[s]] < 0) rep_state[block[s]] = s;
    std::vector<int> order;  // block ids in BFS order
    {
        std::vector<char> vis(n_blocks, 0);
        std::vector<int> q;
        q.push_back(block[0]);
        vis[block[0]] = 1;
        for (size_t h = 0; h < q.size(); ++h) {
            int b = q[h];
            order.push_back(b);
            int s = rep_state[b];
            for (int c = 0; c < C; ++c) {
                int nb2 = block[trans[s * C + c]];
                if (!vis[nb2]) { vis[nb2] = 1; q.push_back(nb2); }
            }
        }
    }
    std::vector<int> newid(n_blocks, -1);
    int M = 0;
    for (int b : order)
        if (states[rep_state[b]].acc.empty()) newid[b] = M++;
    int acc_lo = M;
    for (int b : order)
        if (!states[rep_state[b]].acc.empty()) newid[b] = M++;
    if (M > 65535) return false;

    // ---- merge identical columns (final byte classes) --------------------
    std::vector<int> col_class(C, -1);
    int C2 = 0;
    {
        std::map<std::vector<int>, int> cols;
        for (int c = 0; c < C; ++c) {
            std::vector<int> col(M);
            for (int b : order) col[newid[b]] = newid[block[trans[rep_state[b] * C + c]]];
            auto it = cols.find(col);
            if (it == cols.end()) it = cols.emplace(std::move(col), C2++).first;
            col_class[c] = it->second;
        }
    }

    Dfa& D = *out;
    D = Dfa();
    D.n_states = M;
    D.n_classes = C2;
    D.start = newid[block[0]];
    D.acc_lo = acc_lo;
    for (int c = 0; c < 256; ++c) D.classmap[c] = (uint8_t)col_class[raw_class_of[c]];
    D.trans.assign((size_t)M * C2, 0);
    D.acc.assign(M, {});
    D.endacc.assign(M, {});
    for (int b : order) {
        int s = rep_state[b];
        int id = newid[b];
        for (int c = 0; c < C; ++c) D.trans[(size_t)id * C2 + col_class[c]] = (uint16_t)newid[block[trans[s * C + c]]];
        D.acc[id] = states[s].acc;
        D.endacc[id] = endacc[s];
    }
    return true;
}

namespace {

struct Grouper {
    const Nfa& nfa;
    const std::vector<PatternBundle>& bundles;
    int max_states;
    size_t max_bytes;
    int max_latches;
    DfaGroups* out;
    std::vector<int>* too_big;

    bool try_build(const std::vector<int>& idx, Dfa* d) {
        std::vector<int> s;
        int latches = 0;
        for (int i : idx) {
            s.insert(s.end(), bundles[i].starts.begin(), bundles[i].starts.end());
            latches += bundles[i].has_latch;
        }
        if (latches > max_latches) return false;
        // allow the raw construction some slack over the post-minimisation cap
        if (!build_dfa(nfa, s, max_states * 3, d)) return false;
        return d->n_states <= max_states && d->table_bytes() <= max_bytes;
    }

    bool split(const std::vector<int>& idx) {
        Dfa d;
        if (try_build(idx, &d)) {
            out->dfas.push_back(std::move(d));
            out->members.push_back(idx);
            return true;
        }
        if (idx.size() == 1) { too_big->push_back(idx[0]); return true; }
        size_t h = idx.size() / 2;
        std::vector<int> a(idx.begin(), idx.begin() + h), b(idx.begin() + h, idx.end());
        return split(a) && split(b);
    }
};

}  // namespace

void build_dfa_groups(const Nfa& nfa, const std::vector<PatternBundle>& bundles, int max_states, size_t max_table_bytes,
                      int max_latches, DfaGroups* out, std::vector<int>* too_big) {
    out->dfas.clear();
    out->members.clear();
    too_big->clear();
    if (bundles.empty()) return;
    Grouper G{nfa, bundles, max_states, max_table_bytes, max_latches, out, too_big};
    std::vector<int> all(bundles.size());
    for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
    G.split(all);
    // greedy merge of neighbouring groups (halving can leave mergeable fragments)
    bool merged = true;
    while (merged && out->dfas.size() > 1) {
        merged = false;
        for (size_t i = 0; i + 1 < out->dfas.size(); ++i) {
            if (out->dfas[i].n_states + out->dfas[i + 1].n_states > max_states) continue;  // cheap reject: union is rarely smaller
            std::vector<int> u = out->members[i];
            u.insert(u.end(), out->members[i + 1].begin(), out->members[i + 1].end());
            Dfa d;
            if (G.try_build(u, &d)) {
                out->dfas[i] = std::move(d);
                out->members[i] = u;
                out->dfas.erase(out->dfas.begin() + i + 1);
                out->members.erase(out->members.begin() + i + 1);
                merged = true;
                break;
            }
        }
    }
}

}  // namespace pgw
