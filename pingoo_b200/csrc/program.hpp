// Device program layout: what the host compiler hands to the CUDA kernels.
// Plain-old-data only; included from both host C++ and .cu files.
#pragma once
#include <cstdint>

namespace pgw {

constexpr uint32_t kNoService = 0xFFFFu;             // service index when no service takes the request
constexpr uint32_t kNoRule = 0x3FFFFFFFu;          // verdict rule index when no rule decided
constexpr uint32_t kMaxStackDepth = 32;            // rule bytecode evaluation stack (one 32-bit register)
constexpr uint32_t kCountryWords = 22;             // ceil(676 / 32)

// verdict word = action | rule_index << 2
enum VerdictAction : uint32_t { V_ALLOW = 0, V_BLOCK = 1, V_CAPTCHA = 2, V_BYPASS = 3 };

// request flag bits (pgw_batch.flags)
enum ReqFlags : uint8_t {
    RF_CAPTCHA_VERIFIED = 1,  // valid __pingoo_captcha_verified cookie (http_listener.rs:222-236)
    RF_PRE_BLOCK = 2,         // host-side gate decided "blocked" (e.g. non-ASCII user-agent, http_listener.rs:159-165,196-198)
    RF_PRE_CAPTCHA = 4,       // captcha cookie present but invalid -> serve captcha (http_listener.rs:231-235)
    RF_BYPASS = 8             // request is for the captcha API itself (http_listener.rs:200-204)
};

// DFA accept-event word (uint32): kind << 30 | latch << 24 | atom.  Kinds follow regex.hpp EventKind:
//   0 FIRE  atom := 1                     1 TEST  if latch set: atom := 1
//   2 CLEAR latch := 0                    3 SET   latch := 1
// Lists are sorted by kind, which is also the order events of one position must be applied in.
constexpr uint32_t kEvKindShift = 30, kEvLatchShift = 24, kEvAtomMask = 0x3FFF;
constexpr uint32_t kMaxLatchesPerUnit = 32;

// Candidate unit masks: a gram of a gated field leads to the field's gated scan units whose patterns contain it
// (unit g of the field -> bit g % kGateWidth[field] of the mask).  Index = Field enum.
constexpr uint32_t kGateWidth[5] = {0, 32, 32, 0, 32};   // only url, path and user_agent are gated

// rule bytecode (uint16): 0x0000..0x3FFF push atom, else opcode
enum RuleOp : uint16_t { OP_NOT = 0x4000, OP_AND = 0x4001, OP_OR = 0x4002, OP_PUSH0 = 0x4003, OP_PUSH1 = 0x4004 };

// one scan unit = one DFA over one string field
struct UnitDesc {
    uint32_t field;        // Field enum
    uint32_t n_classes;    // row width (entries)
    uint32_t n_states;
    uint32_t start_state;
    uint32_t acc_lo;       // states >= acc_lo fire accept events
    uint32_t tbl_off;      // byte offset of uint16 trans[n_states][n_classes] in the table arena
    uint32_t cls_off;      // byte offset of the 256-byte class map in the table arena
    uint32_t acc_base;     // acc_idx[acc_base + (s - acc_lo)] .. [+1] -> range in acc_events
    uint32_t end_base;     // end_idx[end_base + s] .. [+1] -> range in end_events
    uint32_t end_any;      // 0 if no state of this DFA has end-of-input accepts (skip finalisation)
    uint32_t field_slot;   // index among the fields that are actually scanned
    uint32_t hot_states;   // states < hot_states have their rows in the shared-memory image; row `hot_states` is the trap row
    uint32_t hot_off;      // byte offset of those rows in the unit's shared-memory image
    uint32_t lim;          // min(hot_states, acc_lo): a walked word whose maximum state is < lim needs no attention at all.
                           // In the image every transition to a cold state (>= hot_states) is replaced by the trap row index.
    uint32_t acc1_off;     // unit-image offset of uint16 acc1[s - acc_lo] for acc_lo <= s < hot_states: the atom of a single-FIRE
                           // event list, or 0xFFFF when the list needs the general path
    uint32_t end1_off;     // unit-image offset of uint16 end1[s] for s < hot_states: end-of-field events of state s:
                           // 0xFFFE none, an atom id for a single FIRE, 0xFFFF general list
    uint32_t mode;         // UnitMode: which requests the unit walks
    uint32_t has_latch;    // the unit has gap-split patterns (latch events must be applied in string order)
    uint32_t abs0, abs1;   // absorbing states (every transition leads back to the state itself): a string that reaches one
                           // is finished early, as if the field ended there; 0xFFFFFFFF when absent
    uint32_t img_off;      // byte offset of this unit's shared-memory image in the image buffer (256-byte aligned)
    uint32_t img_bytes;    // size of that image: class map (offset 0), hot rows + trap row (hot_off), acc1, end1
    uint32_t start_end;    // 1 if the start state has end-of-field events (empty fields are finished by the epilogue)
    uint32_t gate_bit;     // UM_CANDIDATES: this unit's bit in a candidate's unit mask (unit index among the field's gated units % kGateWidth)
};

// one bit-parallel NFA unit (nfa_bits.hpp): a bundle of patterns too large for any DFA unit, walked for every request
struct BitsetUnitDesc {
    uint32_t field;        // Field enum
    uint32_t n_pos;        // P: byte-consuming NFA positions; row P of the tables = the start nodes
    uint32_t words;        // W = ceil(P / 32): state words per request
    uint32_t n_tables;     // T: distinct look-around contexts
    uint32_t n_classes;
    uint32_t n_patterns;   // <= 32
    uint32_t blob_off;     // word offset of the unit's tables in the blob
    uint32_t blob_words;
    uint32_t follow_off, accept_off, bmask_off;   // word offsets inside the unit's tables (layout: nfa_bits.hpp)
    uint32_t stop_mask;    // all patterns, if every event is a plain FIRE: the walk stops once they have all fired; else 0
};

// which requests a scan unit walks
enum UnitMode : uint32_t {
    UM_ALL = 0,        // every request (patterns the gate cannot cover; start-anchored patterns, which finish early)
    UM_CANDIDATES = 1, // only the requests the candidate gate flagged for the unit's field
    UM_PREPASS = 2     // every request, but inside the per-request (epilogue) kernel: small early-exit units whose tables fit its
                       // shared memory next to each other (set at finalize, not by the compiler)
};

// INT_EXPR programs: postfix tokens, one int64 each: opcode in the top byte, operand in the low 56 bits (sign-extended
// constants that do not fit use IT_CONST64 followed by a full word).  Arithmetic is checked i64 (overflow, division by
// zero, INT64_MIN / -1 are errors: the comparison is false and the "errors" atom true), as bel's Int (SEMANTICS.md A4).
enum IntTok : uint32_t { IT_END = 0, IT_CONST = 1, IT_CONST64 = 2, IT_FEAT = 3, IT_ADD = 4, IT_SUB = 5, IT_MUL = 6, IT_DIV = 7, IT_MOD = 8, IT_NEG = 9 };
constexpr uint32_t kIntExprIsError = 6;   // NsAtom::op for "the program errors" (0..5 = CmpOp on the two results)
constexpr uint32_t kIntExprStack = 8;     // deepest operand stack a program may need

// predicates evaluated once per request outside the byte scan
struct NsAtom {
    uint32_t kind;     // AtomDesc::Kind
    uint32_t atom;     // atom bit index
    uint32_t feat;     // IntFeat
    uint32_t op;       // CmpOp
    int64_t cval;
    uint32_t set_id;   // int set / ip set bit / country set; INT_EXPR: offset of the program in the token array;
                       // FIELD_CMP: the second field (feat = the first, op = 0 ==, 1 starts_with, 2 ends_with, 3 contains, 4 <, 5 <=, 6 >, 7 >=)
    uint32_t pad;
};

struct LpmLeaf {
    uint32_t asn;
    uint16_t country;   // two ASCII bytes, first letter in the low byte
    uint16_t pad;
    uint32_t set_mask;  // bit i: address is inside ip set i
};

}  // namespace pgw
