// Regex front-end for the WAF engine: Rust-`regex`-syntax subset -> Thompson NFA.
//
// Replaces (for the rule hot path) what `bel` delegates to the `regex 1.12.2`
// crate (reference Cargo.lock:1694-1695; call site pingoo/rules.rs:38 ->
// bel::Program::execute).  Only *match existence* (`Regex::is_match`,
// unanchored search) is needed, so greediness/captures are parsed and ignored.
//
// Byte semantics: haystacks are ASCII (SEMANTICS.md A9).  Classes are
// intersected with ASCII at compile time; a byte >= 0x80 only matches `.`,
// negated classes and negated escapes (it is treated as one opaque char).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pgw {

struct ByteSet {
    uint64_t w[4] = {0, 0, 0, 0};
    void set(unsigned c) { w[c >> 6] |= 1ull << (c & 63); }
    void set_range(unsigned lo, unsigned hi) { for (unsigned c = lo; c <= hi; ++c) set(c); }
    bool test(unsigned c) const { return (w[c >> 6] >> (c & 63)) & 1; }
    void negate() { for (auto& x : w) x = ~x; }
    void or_with(const ByteSet& o) { for (int i = 0; i < 4; ++i) w[i] |= o.w[i]; }
    bool operator==(const ByteSet& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
    bool empty() const { return !(w[0] | w[1] | w[2] | w[3]); }
};

enum AssertKind : uint8_t {
    A_BOL_TEXT = 0,   // ^ (no m flag), \A
    A_EOL_TEXT = 1,   // $ (no m flag), \z
    A_BOL_LINE = 2,   // ^ with (?m)
    A_EOL_LINE = 3,   // $ with (?m)
    A_WORD_B = 4,     // \b
    A_NOT_WORD_B = 5, // \B
    A_WORD_START = 6,       // \b{start}, \<      : no word byte before, a word byte after
    A_WORD_END = 7,         // \b{end}, \>        : a word byte before, no word byte after
    A_WORD_START_HALF = 8,  // \b{start-half}     : no word byte before
    A_WORD_END_HALF = 9     // \b{end-half}       : no word byte after
};
inline bool assert_looks_at_words(int a) { return a >= A_WORD_B && a <= A_WORD_END_HALF; }

enum NfaKind : uint8_t { N_CHAR, N_SPLIT, N_ASSERT, N_MATCH, N_JUMP };

struct NfaNode {
    NfaKind kind;
    uint8_t assert_kind = 0;
    int out = -1;    // CHAR/ASSERT/JUMP successor, SPLIT first branch
    int out1 = -1;   // SPLIT second branch
    int set = -1;    // CHAR: index into Nfa::sets
    int pattern = -1;  // MATCH: pattern id
};

struct Nfa {
    std::vector<NfaNode> nodes;
    std::vector<ByteSet> sets;
    int add_set(const ByteSet& s);
};

enum RegexStatus {
    RX_OK = 0,
    RX_INVALID = 1,      // Rust `regex` would reject it -> runtime error -> rule is "no match"
    RX_UNSUPPORTED = 2,  // valid Rust syntax this engine does not implement -> loud finalize error
    RX_TOO_BIG = 3       // exceeds the compiled-size cap -> treated like RX_INVALID (regex size_limit)
};

// What a DFA accept event does (see RegexParts).
enum EventKind : uint8_t { EV_FIRE = 0, EV_TEST = 1, EV_CLEAR = 2, EV_SET = 3 };

// A compiled pattern is one NFA pattern (EV_FIRE: the atom is true when it matches), or -- for patterns of
// the shape  X G* S  where G is a byte class whose complement has at most 4 bytes and S is a single byte
// class (e.g. `<script[^>]*>`) -- three cooperating patterns that avoid the 2^k state blow-up such "sticky"
// gap loops cause in a multi-pattern DFA:
//   part 0  X          EV_SET    latch := 1 when X has matched ending here
//   part 1  S          EV_TEST   atom := true if latch is set when a byte of S is seen
//   part 2  not-G      EV_CLEAR  latch := 0 when a byte outside G is seen      (absent if G is every byte)
// Events of one position are applied in the order TEST, CLEAR, SET.
struct RegexParts {
    int n = 0;
    int start[3] = {-1, -1, -1};
    uint8_t kind[3] = {EV_FIRE, EV_FIRE, EV_FIRE};
};

struct RegexInfo {
    bool always_true = false;      // nullable without crossing an assertion: is_match is true on every haystack
    bool uses_word_boundary = false;
    bool uses_multiline = false;
    bool uses_bol = false;
};

// Compile `pattern` into `nfa`; part k ends in a MATCH node carrying pattern id `first_pattern_id + k`.
// On failure `err` holds a message.
RegexStatus regex_compile(const std::string& pattern, int first_pattern_id, Nfa& nfa, RegexParts* parts, RegexInfo* info,
                          std::string& err, bool allow_split = true);

// Literal helpers used for ==, starts_with, ends_with, contains: build the
// equivalent anchored/unanchored literal pattern straight into the NFA.
int nfa_literal(Nfa& nfa, const std::string& lit, bool anchor_start, bool anchor_end, int pattern_id);

// Hard cap on NFA nodes contributed by one pattern (counted repetitions expand).
constexpr int kMaxNfaNodesPerPattern = 200000;

}  // namespace pgw
