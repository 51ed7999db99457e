// Bit-parallel NFA unit: the patterns no DFA of acceptable size exists for.
//
// Rust `regex` (Cargo.lock:1694-1695, reached through bel from pingoo/rules.rs:38) never refuses a pattern because its
// DFA would be large: the lazy DFA gives up and the PikeVM runs the NFA directly.  This is the engine's counterpart.
// A bundle (dfa.hpp) that exceeds the scan-unit caps on its own is simulated as a SET of NFA positions held in a bit
// vector, one request per thread (kernel_bitset.cuh):
//
//   positions        the byte-consuming nodes reachable from the bundle's start nodes, p = 0 .. P-1; bit p of the state
//                    means "position p has just consumed a byte" (its successor node is in the kernel of dfa.cpp's
//                    subset construction);  row P stands for the start nodes, which are active at every boundary
//                    (unanchored search);
//   context          what an assertion may look at, at the boundary between two bytes: the kind of the previous byte
//                    (none = start of the field, word, newline, other) x the kind of the next one (none = end of the
//                    field, word, newline, other) -- 16 combinations, mapped to the few DISTINCT tables they produce
//                    (one, for a pattern without assertions);
//   follow[t][k]     the positions reachable from row k through epsilon edges whose assertions hold in context t;
//   accept[t][k]     the bundle's patterns (bit i = pattern i, at most 32) whose MATCH node is reachable likewise;
//   bmask[c]         the positions whose byte set contains the bytes of class c.
//
// One step over byte b in context t:  R = follow[t][P] | OR_{p in S} follow[t][p];  matched = accept[t][P] | OR accept[t][p];
// S' = R & bmask[class(b)].  `matched` is applied through the bundle's events (FIRE / TEST / CLEAR / SET, regex.hpp) in
// that order, as the DFA units do; at the end of the field one more boundary is evaluated with "next = none".
// The semantics are exactly those of dfa.cpp's Builder::closure + build_dfa (same contexts, same kernel), so that a
// pattern gives the same answer whichever unit walks it; tests/test_bitset_nfa.py forces every pattern of the regex
// corpus through this unit and compares with the oracle.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "program.hpp"
#include "regex.hpp"

namespace pgw {

constexpr uint32_t kBitsetMaxPositions = 2048;   // positions of one bundle (64 state words per thread at most)
constexpr uint32_t kBitsetMaxPatterns = 32;

// Word layout of one unit's table blob (uint32 words; the device copy is the same bytes):
//   [0, 64) class map (256 bytes)   [64, 128) byte kind (256 bytes: 1 other, 2 word, 3 newline)
//   [128, 132) context -> table (16 bytes, index = previous kind * 4 + next kind; kind 0 = no byte on that side)
//   [132, 164) event words of the patterns (program.hpp: kind << 30 | latch << 24 | atom), sorted by kind
//   follow_off: follow[T][P + 1][W]   accept_off: accept[T][P + 1]   bmask_off: bmask[n_classes][W]
constexpr uint32_t kBitsetBlobHeaderWords = 164;

// Builds the unit for the bundle whose NFA start nodes are `starts`; `event_words[i]` is the event of the pattern whose
// MATCH node carries id `pattern_ids[i]`.  Appends the tables to `blob` and fills `desc` (blob_off = word offset).
// False with `err` set when the bundle has more than kBitsetMaxPositions positions or kBitsetMaxPatterns patterns.
bool build_bitset_unit(const Nfa& nfa, const std::vector<int>& starts, const std::vector<int>& pattern_ids,
                       const std::vector<uint32_t>& event_words, int field, std::vector<uint32_t>* blob, BitsetUnitDesc* desc,
                       std::string& err);

// The walk itself on the host tables (the simulator's mirror of kernel_bitset.cuh): calls fire(atom) for every atom the
// field [s, e) of `bytes` makes true.
template <class Fire>
inline void bitset_walk_host(const BitsetUnitDesc& d, const uint32_t* tab, const uint8_t* bytes, uint32_t s, uint32_t e, Fire&& fire) {
    const uint8_t* cmap = reinterpret_cast<const uint8_t*>(tab);
    const uint8_t* kind = cmap + 256;
    const uint8_t* ctx = cmap + 512;
    const uint32_t* events = tab + 132;
    const uint32_t W = d.words, P = d.n_pos;
    std::vector<uint32_t> S(W, 0), R(W, 0);
    uint32_t latch = 0, pk = 0, fired = 0;
    for (uint32_t i = s;; ++i) {
        const bool end = i >= e;
        const uint32_t byte = end ? 0u : bytes[i];
        const uint32_t t = ctx[pk * 4u + (end ? 0u : kind[byte])];
        const uint32_t* F = tab + d.follow_off + (size_t)t * (P + 1) * W;
        const uint32_t* A = tab + d.accept_off + (size_t)t * (P + 1);
        for (uint32_t w = 0; w < W; ++w) R[w] = F[(size_t)P * W + w];
        uint32_t acc = A[P];
        for (uint32_t w = 0; w < W; ++w)
            for (uint32_t x = S[w]; x; x &= x - 1) {
                const uint32_t p = w * 32u + (uint32_t)__builtin_ctz(x);
                acc |= A[p];
                for (uint32_t v = 0; v < W; ++v) R[v] |= F[(size_t)p * W + v];
            }
        for (uint32_t x = acc; x; x &= x - 1) {
            const uint32_t j = (uint32_t)__builtin_ctz(x), ev = events[j];
            const uint32_t k = ev >> kEvKindShift, lb = 1u << ((ev >> kEvLatchShift) & 31u);
            if (k == 0u || (k == 1u && (latch & lb))) { fire(ev & kEvAtomMask); fired |= 1u << j; }
            else if (k == 2u) latch &= ~lb;
            else if (k == 3u) latch |= lb;
        }
        if (end) break;
        if (d.stop_mask && (fired & d.stop_mask) == d.stop_mask) break;
        const uint32_t* B = tab + d.bmask_off + (size_t)cmap[byte] * W;
        for (uint32_t w = 0; w < W; ++w) S[w] = R[w] & B[w];
        pk = kind[byte];
    }
}

}  // namespace pgw
