// Multi-pattern NFA -> DFA (subset construction with look-around context),
// Moore minimisation, byte-class compression and pattern grouping.
//
// One DFA answers, for a set of patterns over one request field, "which
// patterns match somewhere in this haystack" (Regex::is_match semantics,
// reference call path pingoo/rules.rs:38 -> bel -> regex).  Events:
//   acc[s]    patterns whose match ended just before the byte that led into s
//   endacc[s] patterns that match when the haystack ends in state s
#pragma once
#include <cstdint>
#include <vector>

#include "regex.hpp"

namespace pgw {

struct Dfa {
    int n_states = 0;
    int n_classes = 0;
    int start = 0;
    int acc_lo = 0;  // states >= acc_lo have a non-empty acc list
    uint8_t classmap[256];
    std::vector<uint16_t> trans;             // [n_states][n_classes] -> state id
    std::vector<std::vector<int>> acc;       // pattern ids
    std::vector<std::vector<int>> endacc;    // pattern ids
    size_t table_bytes() const { return (size_t)n_states * n_classes * 2; }
};

// Build one DFA for the patterns whose NFA start nodes are `starts`.
// Returns false if the (unminimised) subset construction exceeds `max_raw_states`.
bool build_dfa(const Nfa& nfa, const std::vector<int>& starts, int max_raw_states, Dfa* out);

// A bundle = the NFA patterns of one atom (1 normally, 2-3 for a gap-split pattern: they share a latch
// and therefore must land in the same DFA).
struct PatternBundle {
    std::vector<int> starts;
    bool has_latch = false;
};

struct DfaGroups {
    std::vector<Dfa> dfas;
    std::vector<std::vector<int>> members;  // indices into `bundles`
};

// Partition bundles into as few DFAs as possible subject to the caps (and at most `max_latches` latch
// bundles per DFA).  A bundle that exceeds the caps on its own lands in `too_big` (indices into `bundles`) and in no
// group: the caller hands it to the bit-parallel NFA unit (nfa_bits.hpp), as Rust `regex` leaves its DFA for the PikeVM.
void build_dfa_groups(const Nfa& nfa, const std::vector<PatternBundle>& bundles, int max_states, size_t max_table_bytes,
                      int max_latches, DfaGroups* out, std::vector<int>* too_big);

}  // namespace pgw
