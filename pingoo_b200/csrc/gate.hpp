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
// Folding: bit 5 of every byte is cleared (lower case -> upper case; digits and punctuation alias with control bytes
// that never occur in a request), the same single AND the kernel applies to each word: g & 0xDFDFDFDF.
#pragma once
#include <cstdint>
#include <string>
#include <tuple>
#include <vector>

#include "regex.hpp"

namespace pgw {

// Level-1 hashes (pipe rates measured by tools/microbench_int.cu: IMAD and the ALU ops SHF / LOP3 issue every 2 cycles
// per scheduler on different pipes, IMAD.HI every 4): the Bloom WORD is chosen by the top bits of the low product
// g * K (one IMAD + one SHF; a low-byte change of g moves the top bits of the low product a lot, so the 128 grams of a
// wildcard expansion land in 128 different words), the two BIT positions are the low five bits of the high products
// hi(g * B) and hi(g * C) (IMAD.HI: no shift needed, a rotate takes its amount mod 32).  Level 2 uses g * kGateHash2.
constexpr uint32_t kGateHashK = 0x9E3779B1u, kGateHashB = 0x85EBCA6Bu, kGateHashC = 0xC2B2AE35u, kGateHash2 = 0x85EBCA6Bu;
constexpr uint32_t kGateMaxLog2 = 20;  // largest first-level bitmap: 2^20 bits = 128 KB of shared memory
constexpr uint32_t kGateFoldMask = 0xDFDFDFDFu;

inline uint32_t gate_fold(uint32_t g) { return g & kGateFoldMask; }
inline uint32_t gate_mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
// level 1 of a folded gram on bitmap `b1` of 2^k1 bits (k1 >= 12)
inline bool gate_l1_test(const uint32_t* b1, uint32_t k1, uint32_t g) {
    const uint32_t word = b1[(g * kGateHashK) >> (32 - (k1 - 5))];
    return ((word >> (gate_mulhi(g, kGateHashB) & 31)) & (word >> (gate_mulhi(g, kGateHashC) & 31)) & 1u) != 0;
}
inline void gate_l1_set(uint32_t* b1, uint32_t k1, uint32_t g) {
    b1[(g * kGateHashK) >> (32 - (k1 - 5))] |= (1u << (gate_mulhi(g, kGateHashB) & 31)) | (1u << (gate_mulhi(g, kGateHashC) & 31));
}

// Two levels.  Level 1 (shared memory, probed for every window): a blocked Bloom filter -- the hashes select one
// 32-bit word and two bit positions in it, both must be set (false-positive rate = density^2, a few 1e-4).
// Level 2 (global memory, probed only for level-1 survivors): an exact open-addressing table folded gram -> mask of
// the field's gated scan units whose patterns contain the gram, so a candidate is only walked by those units.
// A pattern whose language is a small finite set of byte strings (alternations of literals, `contains`, `ends_with`,
// optional / case-insensitive letters) needs no automaton: every occurrence covers a gate window whose gram is known,
// so the resolve kernel CONFIRMS it by comparing the string at the hit position and fires the atom itself -- the request
// becomes a scan candidate only for the grams of real regex patterns.  (What Rust `regex` does when a pattern is a
// literal alternation: it never builds an automaton for it either.)
struct LitString {
    std::string bytes;        // case-insensitive positions hold the upper-case letter
    uint64_t ci_mask = 0;     // bit k: byte k matches either case
    bool anch_start = false;  // must start at the first byte of the field
    bool anch_end = false;    // must end at the last byte of the field
    bool operator<(const LitString& o) const {
        return std::tie(bytes, ci_mask, anch_start, anch_end) < std::tie(o.bytes, o.ci_mask, o.anch_start, o.anch_end);
    }
    bool operator==(const LitString& o) const { return bytes == o.bytes && ci_mask == o.ci_mask && anch_start == o.anch_start && anch_end == o.anch_end; }
};
constexpr size_t kLitMaxStrings = 48, kLitMaxLen = 64, kLitMinLen = 3;

// device / table form of one literal (program.hpp style POD)
struct LitDesc {
    uint32_t off;       // into the byte pool
    uint16_t len;
    uint16_t flags;     // 1: anchored at the field start, 2: anchored at the field end
    uint32_t atom;
    uint32_t pad;
    uint64_t ci_mask;
    uint64_t pad2;      // 32 bytes: the resolve kernel reads a descriptor as two 16-byte loads
};
static_assert(sizeof(LitDesc) == 32, "LitDesc is read as two uint4");

struct GateTables {
    bool present = false;
    uint32_t k1 = 0;                  // log2(bits) of the level-1 bitmap
    std::vector<uint32_t> b1;         // 2^k1 / 32 words
    uint32_t kt = 0;                  // log2(slots) of the level-2 table
    uint32_t slot_words = 2;          // 2: {gram, unit mask}, mask 0 = empty slot; 4 (some field of the rule set confirms literals):
                                      // {gram, unit mask, first literal candidate, number of them}, mask 0 and count 0 = empty slot
    std::vector<uint32_t> slots;
    uint32_t n_grams = 0;
    // literal candidates of a gram: (literal index << 2) | (delta + 1), the literal would start at window position + delta
    std::vector<uint32_t> lit_cand;
    std::vector<LitDesc> lits;
    std::vector<uint8_t> lit_bytes;
    // unit mask of the window (0: no scan unit asks for it); `lit` (optional) receives the gram's literal candidates
    uint32_t probe(uint32_t window_le, uint32_t* lit_begin = nullptr, uint32_t* lit_count = nullptr) const {
        if (lit_count) *lit_count = 0;
        const uint32_t g = gate_fold(window_le);
        if (!gate_l1_test(b1.data(), k1, g)) return 0;
        const uint32_t tm = (1u << kt) - 1u;
        const uint32_t W = slot_words;
        for (uint32_t s = (g * kGateHash2) >> (32 - kt);; s = (s + 1) & tm) {
            if (slots[W * s + 1] == 0 && (W == 2 || slots[W * s + 3] == 0)) return 0;
            if (slots[W * s] == g) {
                if (W == 4 && lit_begin) *lit_begin = slots[W * s + 2];
                if (W == 4 && lit_count) *lit_count = slots[W * s + 3];
                return slots[W * s + 1];
            }
        }
    }
    // does literal `d` occur at column position `at` of the field [s, e) of column `col`?
    bool lit_matches(const LitDesc& d, const uint8_t* col, uint32_t s, uint32_t e, int64_t at) const {
        if (at < (int64_t)s || at + d.len > (int64_t)e) return false;
        if ((d.flags & 1) && at != (int64_t)s) return false;
        if ((d.flags & 2) && at + d.len != (int64_t)e) return false;
        for (uint32_t k = d.len; k-- > 0;) {   // last byte first: the bytes behind the announcing gram tell candidates apart
            const uint8_t b = col[at + k], want = lit_bytes[d.off + k];
            if (b != want && !(((d.ci_mask >> k) & 1) && (uint8_t)(b ^ 0x20) == want)) return false;
        }
        return true;
    }
};

// True if every path from `start` to a byte-consuming node crosses a start-of-text assertion (such a pattern is
// decided by a prefix of the field: it goes to an early-exit scan unit, not to the gate).
bool pattern_is_start_anchored(const Nfa& nfa, int start);

// Folded little-endian 4-grams for the pattern starting at NFA node `start`; false if it cannot be gated
// (matches shorter than 3 bytes, or more than `cap` grams).
bool gate_grams_for_pattern(const Nfa& nfa, int start, size_t cap, std::vector<uint32_t>* out);

// The finite language of the pattern starting at NFA node `start`, if it is one the literal path can take: at most
// kLitMaxStrings strings of kLitMinLen..kLitMaxLen bytes, every byte a single value or a letter in both cases, `^` only in
// front and `$` only behind, no other assertion, no loop.
bool gate_finite_language(const Nfa& nfa, int start, std::vector<LitString>* out);

// A literal of a gated field and the grams that announce it: (folded gram, delta) with the literal starting at
// window position + delta (gate.hpp soundness argument: A = m[0..4) at delta 0; B = m[1..5) at delta -1 when the literal
// has five bytes, else (any, m[0..3)) at delta +1).
void gate_grams_for_literal(const LitString& s, std::vector<std::pair<uint32_t, int>>* out);

struct GateLiteral {
    LitString str;
    uint32_t atom;
};

// `grams[i]` belongs to the units in `masks[i]` (duplicates are merged by OR); `literals` are confirmed by the resolve kernel
// `max_log2`: largest level-1 bitmap the field may use (all gated fields' bitmaps are resident in shared memory together)
// `wide_slots`: the 4-word slot layout (required when `literals` is not empty; chosen for every field of a rule set alike)
void gate_build_tables(const std::vector<uint32_t>& grams, const std::vector<uint32_t>& masks, const std::vector<GateLiteral>& literals,
                       bool wide_slots, uint32_t max_log2, GateTables* out);

}  // namespace pgw
