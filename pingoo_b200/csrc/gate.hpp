// Candidate gate: a position-local prefilter in front of the DFA scan.
//
// Rust `regex` (the engine behind bel's `matches`, reference Cargo.lock:1694-1695) runs literal prefilters
// (memchr / Teddy / Aho-Corasick) in front of its automata; this is the batched equivalent.  For every pattern
// that can be gated we derive, from its NFA, a set of 4-byte *grams* such that EVERY occurrence of the pattern in
// a haystack covers (or touches, for very short patterns) at least one even-aligned 4-byte window of the column
// whose case-folded content is in the set.  The gate kernel tests every even-aligned window of a field column
// against a two-hash bitmap of those grams -- stateless, so the column is read as one flat coalesced stream --
// and only requests with a hit ("candidates") are walked by the unit's DFA afterwards.
//
// Soundness argument (tests/test_gate.py checks it against the oracle):
//   a match m starting at column position p either has p even -> window [p, p+4) = m[0..4) (padded with
//   arbitrary bytes after the match end), or p odd -> window [p-1, p+3) = (arbitrary byte, m[0..3)) or window
//   [p+1, p+5) = m[1..5) (padded).  The gram set of a pattern therefore is
//       A  = { m[0..4) padded }                        and
//       B  = the smaller of  { (any, m[0..3)) padded }  and  { m[1..5) padded }
//   with "padded" / "any" positions expanded over every folded byte value.  Assertions (^ $ \b) are treated as
//   always true, which only enlarges the sets.
// Folding: bytes 0x40-0x5F and 0xC0-0xDF get bit 5 set (upper case -> lower case; a few punctuation marks
// alias), the same operation the kernel applies to the window: g | ((g & 0x40404040) >> 1).
#pragma once
#include <cstdint>
#include <vector>

#include "regex.hpp"

namespace pgw {

constexpr uint32_t kGateHash1 = 0x9E3779B1u, kGateHash2 = 0x85EBCA6Bu;
constexpr uint32_t kGateMaxLog2 = 20;  // largest bitmap: 2^20 bits = 128 KB (two of them must fit shared memory)

inline uint32_t gate_fold(uint32_t g) { return g | ((g & 0x40404040u) >> 1); }

struct GateTables {
    bool present = false;
    uint32_t k1 = 0, k2 = 0;          // log2(bits) of the two bitmaps
    std::vector<uint32_t> b1, b2;     // 2^k / 32 words each
    uint32_t n_grams = 0;
    bool test(uint32_t window_le) const {
        const uint32_t g = gate_fold(window_le);
        const uint32_t h1 = (g * kGateHash1) >> (32 - k1), h2 = (g * kGateHash2) >> (32 - k2);
        return ((b1[h1 >> 5] >> (h1 & 31)) & 1u) && ((b2[h2 >> 5] >> (h2 & 31)) & 1u);
    }
};

// True if every path from `start` to a byte-consuming node crosses a start-of-text assertion (such a pattern is
// decided by a prefix of the field: it goes to an early-exit scan unit, not to the gate).
bool pattern_is_start_anchored(const Nfa& nfa, int start);

// Folded little-endian 4-grams for the pattern starting at NFA node `start`; false if it cannot be gated
// (matches shorter than 3 bytes, or more than `cap` grams).
bool gate_grams_for_pattern(const Nfa& nfa, int start, size_t cap, std::vector<uint32_t>* out);

void gate_build_tables(std::vector<uint32_t> grams, GateTables* out);

}  // namespace pgw
