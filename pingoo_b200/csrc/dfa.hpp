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

struct DfaGroups {
    std::vector<Dfa> dfas;
    std::vector<std::vector<int>> members;  // indices into `starts`
};

// Partition patterns into as few DFAs as possible subject to the caps.
// Returns false (with *failed_index set) if a single pattern alone exceeds the caps.
bool build_dfa_groups(const Nfa& nfa, const std::vector<int>& starts, int max_states, size_t max_table_bytes,
                      DfaGroups* out, int* failed_index);

}  // namespace pgw
