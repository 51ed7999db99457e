// Candidate gate: a position-local prefilter in front of the DFA scan.
//
// Rust `regex` (the engine behind bel's `matches`, reference Cargo.lock:1694-1695) runs literal prefilters
// (memchr / Teddy / Aho-Corasick) in front of its automata; this is the batched equivalent.  For every pattern
// that can be gated we derive, from its NFA, a set of 4-byte *grams* such that EVERY occurrence of the pattern in
// a haystack covers (or touches, for very short patterns) at least one even-aligned 4-byte window of the column
// whose case-folded content is in the set.  The gate kernel tests every even-aligned window of a field column
// against a bitmap of those grams (then an exact table) -- stateless, so the column is read as one flat coalesced
// stream -- and only requests with a hit ("candidates") are walked, by the DFAs of the units the hit gram belongs to.
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
constexpr uint32_t kGateMaxLog2 = 20;  // largest first-level bitmap: 2^20 bits = 128 KB of shared memory

inline uint32_t gate_fold(uint32_t g) { return g | ((g & 0x40404040u) >> 1); }

// Two levels.  Level 1 (shared memory, probed for every window): a blocked Bloom filter -- the hash selects one
// 32-bit word and two bit positions in it, both must be set (false-positive rate = density^2, a few 1e-4).
// Level 2 (global memory, probed only for level-1 survivors): an exact open-addressing table folded gram -> mask of
// the field's gated scan units whose patterns contain the gram, so a candidate is only walked by those units.
struct GateTables {
    bool present = false;
    uint32_t k1 = 0;                  // log2(bits) of the level-1 bitmap
    std::vector<uint32_t> b1;         // 2^k1 / 32 words
    uint32_t kt = 0;                  // log2(slots) of the level-2 table
    std::vector<uint32_t> slots;      // 2 words per slot: {gram, unit mask}; mask 0 = empty slot
    uint32_t n_grams = 0;
    // unit mask of the window (0: not a candidate window)
    uint32_t probe(uint32_t window_le) const {
        const uint32_t g = gate_fold(window_le);
        const uint32_t h = g * kGateHash1, sh = 32 - k1;
        const uint32_t word = b1[h >> (sh + 5)];
        if (!((word >> ((h >> sh) & 31)) & (word >> ((h >> (sh - 5)) & 31)) & 1u)) return 0;
        const uint32_t tm = (1u << kt) - 1u;
        for (uint32_t s = (g * kGateHash2) >> (32 - kt);; s = (s + 1) & tm) {
            if (slots[2 * s + 1] == 0) return 0;
            if (slots[2 * s] == g) return slots[2 * s + 1];
        }
    }
};

// True if every path from `start` to a byte-consuming node crosses a start-of-text assertion (such a pattern is
// decided by a prefix of the field: it goes to an early-exit scan unit, not to the gate).
bool pattern_is_start_anchored(const Nfa& nfa, int start);

// Folded little-endian 4-grams for the pattern starting at NFA node `start`; false if it cannot be gated
// (matches shorter than 3 bytes, or more than `cap` grams).
bool gate_grams_for_pattern(const Nfa& nfa, int start, size_t cap, std::vector<uint32_t>* out);

// `grams[i]` belongs to the units in `masks[i]` (duplicates are merged by OR)
void gate_build_tables(const std::vector<uint32_t>& grams, const std::vector<uint32_t>& masks, GateTables* out);

}  // namespace pgw
