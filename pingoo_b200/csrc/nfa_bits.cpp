// Tables of the bit-parallel NFA unit (see nfa_bits.hpp).
#include "nfa_bits.hpp"

#include <algorithm>
#include <cstring>
#include <map>

namespace pgw {
namespace {

inline bool is_word_byte(int c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_'; }

// kinds of the byte on either side of a boundary; 0 = there is none (start / end of the field)
enum { K_NONE = 0, K_OTHER = 1, K_WORD = 2, K_NL = 3 };

struct Closure {
    const Nfa& nfa;
    const std::vector<int>& pos_of_node;   // CHAR node -> position, -1 otherwise
    const std::map<int, int>& local_pattern;  // MATCH pattern id -> local index
    std::vector<int> stamp, stack;
    int gen = 0;
    Closure(const Nfa& n, const std::vector<int>& pn, const std::map<int, int>& lp) : nfa(n), pos_of_node(pn), local_pattern(lp), stamp(n.nodes.size(), 0) {}

    // epsilon closure of `from` at a boundary whose sides have the kinds (pk, nk): the same rules as dfa.cpp Builder::closure
    void run(const std::vector<int>& from, int pk, int nk, uint32_t* row, uint32_t* acc) {
        ++gen;
        stack.assign(from.begin(), from.end());
        const bool at_start = pk == K_NONE, prev_word = pk == K_WORD, prev_nl = pk == K_NL;
        const bool at_end = nk == K_NONE, next_word = nk == K_WORD, next_nl = nk == K_NL;
        while (!stack.empty()) {
            const int n = stack.back();
            stack.pop_back();
            if (n < 0 || stamp[n] == gen) continue;
            stamp[n] = gen;
            const NfaNode& nd = nfa.nodes[n];
            switch (nd.kind) {
                case N_CHAR: row[pos_of_node[n] >> 5] |= 1u << (pos_of_node[n] & 31); break;
                case N_MATCH: *acc |= 1u << local_pattern.at(nd.pattern); break;
                case N_JUMP: stack.push_back(nd.out); break;
                case N_SPLIT:
                    stack.push_back(nd.out);
                    stack.push_back(nd.out1);
                    break;
                case N_ASSERT: {
                    bool ok = false;
                    switch (nd.assert_kind) {
                        case A_BOL_TEXT: ok = at_start; break;
                        case A_EOL_TEXT: ok = at_end; break;
                        case A_BOL_LINE: ok = at_start || prev_nl; break;
                        case A_EOL_LINE: ok = at_end || next_nl; break;
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
    }
};

}  // namespace

bool build_bitset_unit(const Nfa& nfa, const std::vector<int>& starts, const std::vector<int>& pattern_ids,
                       const std::vector<uint32_t>& event_words, int field, std::vector<uint32_t>* blob, BitsetUnitDesc* desc,
                       std::string& err) {
    if (pattern_ids.size() > kBitsetMaxPatterns || pattern_ids.size() != event_words.size()) {
        err = "more than 32 patterns in one bit-parallel NFA unit";
        return false;
    }
    // local pattern order = event order (FIRE, TEST, CLEAR, SET): the walk applies matched patterns by ascending index
    std::vector<size_t> order(pattern_ids.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return event_words[x] < event_words[y]; });
    std::map<int, int> local_pattern;
    for (size_t j = 0; j < order.size(); ++j) local_pattern[pattern_ids[order[j]]] = (int)j;

    // ---- positions: the CHAR nodes reachable from the start nodes; which assertion kinds occur
    std::vector<int> pos_of_node(nfa.nodes.size(), -1), node_of_pos;
    bool use_word = false, use_line = false;
    {
        std::vector<char> seen(nfa.nodes.size(), 0);
        std::vector<int> st(starts.begin(), starts.end());
        while (!st.empty()) {
            const int n = st.back();
            st.pop_back();
            if (n < 0 || seen[n]) continue;
            seen[n] = 1;
            const NfaNode& nd = nfa.nodes[n];
            if (nd.kind == N_CHAR) { pos_of_node[n] = (int)node_of_pos.size(); node_of_pos.push_back(n); }
            if (nd.kind == N_ASSERT) {
                if (assert_looks_at_words(nd.assert_kind)) use_word = true;
                if (nd.assert_kind == A_BOL_LINE || nd.assert_kind == A_EOL_LINE) use_line = true;
            }
            if (nd.kind == N_MATCH) {
                if (!local_pattern.count(nd.pattern)) { err = "internal: MATCH node of a pattern outside the bundle"; return false; }
                continue;
            }
            st.push_back(nd.out);
            if (nd.kind == N_SPLIT) st.push_back(nd.out1);
        }
    }
    const uint32_t P = (uint32_t)node_of_pos.size();
    if (P > kBitsetMaxPositions) {
        err = "needs " + std::to_string(P) + " NFA positions (more than " + std::to_string(kBitsetMaxPositions) + ")";
        return false;
    }
    const uint32_t W = std::max<uint32_t>(1, (P + 31) / 32);

    // ---- byte kinds and classes
    uint8_t kind[256], cmap[256];
    for (int c = 0; c < 256; ++c) kind[c] = (use_word && is_word_byte(c)) ? K_WORD : (use_line && c == '\n') ? K_NL : K_OTHER;
    std::vector<int> used_sets;
    for (int n : node_of_pos) used_sets.push_back(nfa.nodes[n].set);
    std::sort(used_sets.begin(), used_sets.end());
    used_sets.erase(std::unique(used_sets.begin(), used_sets.end()), used_sets.end());
    std::vector<int> rep;
    {
        std::map<std::vector<uint8_t>, int> sig2cls;
        for (int c = 0; c < 256; ++c) {
            std::vector<uint8_t> sig;
            sig.reserve(used_sets.size() + 1);
            for (int s : used_sets) sig.push_back(nfa.sets[s].test(c));
            sig.push_back(kind[c]);
            auto it = sig2cls.find(sig);
            if (it == sig2cls.end()) {
                it = sig2cls.emplace(sig, (int)rep.size()).first;
                rep.push_back(c);
            }
            cmap[c] = (uint8_t)it->second;
        }
    }
    const uint32_t C = (uint32_t)rep.size();

    // ---- per context: follow / accept rows; identical contexts share a table
    Closure cl(nfa, pos_of_node, local_pattern);
    std::map<std::vector<uint32_t>, int> table_index;
    std::vector<std::vector<uint32_t>> tables;   // each: (P + 1) * W follow words, then (P + 1) accept words
    uint8_t ctx[16];
    for (int pk = 0; pk < 4; ++pk)
        for (int nk = 0; nk < 4; ++nk) {
            std::vector<uint32_t> t((size_t)(P + 1) * W + (P + 1), 0);
            uint32_t* follow = t.data();
            uint32_t* accept = t.data() + (size_t)(P + 1) * W;
            for (uint32_t p = 0; p < P; ++p) {
                std::vector<int> from(1, nfa.nodes[node_of_pos[p]].out);
                cl.run(from, pk, nk, follow + (size_t)p * W, accept + p);
            }
            cl.run(starts, pk, nk, follow + (size_t)P * W, accept + P);
            auto it = table_index.find(t);
            if (it == table_index.end()) {
                it = table_index.emplace(t, (int)tables.size()).first;
                tables.push_back(std::move(t));
            }
            ctx[pk * 4 + nk] = (uint8_t)it->second;
        }
    const uint32_t T = (uint32_t)tables.size();

    // ---- blob
    BitsetUnitDesc& d = *desc;
    memset(&d, 0, sizeof d);
    d.field = (uint32_t)field;
    d.n_pos = P;
    d.words = W;
    d.n_tables = T;
    d.n_classes = C;
    d.n_patterns = (uint32_t)pattern_ids.size();
    while (blob->size() % 4) blob->push_back(0);   // 16-byte aligned pieces
    d.blob_off = (uint32_t)blob->size();
    const size_t base = blob->size();
    blob->resize(base + kBitsetBlobHeaderWords, 0);
    memcpy(reinterpret_cast<uint8_t*>(blob->data() + base), cmap, 256);
    memcpy(reinterpret_cast<uint8_t*>(blob->data() + base) + 256, kind, 256);
    memcpy(reinterpret_cast<uint8_t*>(blob->data() + base) + 512, ctx, 16);
    bool all_fire = true;
    for (size_t j = 0; j < order.size(); ++j) {
        (*blob)[base + 132 + j] = event_words[order[j]];
        all_fire &= (event_words[order[j]] >> kEvKindShift) == 0u;
    }
    d.stop_mask = all_fire ? (d.n_patterns >= 32 ? 0xFFFFFFFFu : (1u << d.n_patterns) - 1u) : 0u;
    d.follow_off = (uint32_t)(blob->size() - base);
    for (auto& t : tables) blob->insert(blob->end(), t.begin(), t.begin() + (size_t)(P + 1) * W);
    d.accept_off = (uint32_t)(blob->size() - base);
    for (auto& t : tables) blob->insert(blob->end(), t.begin() + (size_t)(P + 1) * W, t.end());
    d.bmask_off = (uint32_t)(blob->size() - base);
    for (uint32_t c = 0; c < C; ++c) {
        std::vector<uint32_t> m(W, 0);
        for (uint32_t p = 0; p < P; ++p)
            if (nfa.sets[nfa.nodes[node_of_pos[p]].set].test((unsigned)rep[c])) m[p >> 5] |= 1u << (p & 31);
        blob->insert(blob->end(), m.begin(), m.end());
    }
    while (blob->size() % 4) blob->push_back(0);
    d.blob_words = (uint32_t)(blob->size() - base);
    return true;
}

}  // namespace pgw
