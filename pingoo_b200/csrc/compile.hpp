// Model (atoms + formulas) -> HostProgram (flat tables ready for upload).
#pragma once
#include <string>
#include <vector>

#include "gate.hpp"
#include "model.hpp"
#include "program.hpp"

namespace pgw {

struct CompileOptions {
    int max_dfa_states = 16384;                // per scan unit, after minimisation; a pattern above it alone -> bit-parallel NFA unit
    size_t max_unit_table_bytes = 8u << 20;    // per scan unit transition table (hot rows go to shared memory, the rest stays in L2)
    bool eval_gates = true;                    // evaluate http_listener.rs:196-204 gates inside the engine
    bool candidate_gate = true;                // gram prefilter in front of the DFA scan of url / user_agent / path (gate.hpp)
    bool literal_confirm = true;               // finite-string patterns of gated fields are confirmed by the gate, not walked by a DFA
    size_t gate_pattern_cap = 4096;            // grams one pattern may contribute before it is left to an ungated unit
    size_t gate_field_cap = 12288;             // grams per field (bitmaps of at most 2^19 bits each)
};

struct LpmTables {
    bool present = false;        // any ip set or geoip database loaded
    bool geo_loaded = false;
    std::vector<uint32_t> dir24; // [1<<24]: bit31 set -> low bits index a 256-entry block in tbl8, else leaf id
    std::vector<uint32_t> tbl8;
    std::vector<LpmLeaf> leaves; // leaf 0 = {asn 0, "XX", no sets}
    // IPv6: sorted, disjoint, covering ranges; range i = [start_i, start_{i+1})
    std::vector<uint64_t> v6_hi, v6_lo;
    std::vector<uint32_t> v6_leaf;
    std::vector<uint32_t> v6_top;   // [65537]: index on the first 16 address bits (lpm.cpp)
};

struct HostProgram {
    std::vector<UnitDesc> units;
    std::vector<uint8_t> arena;  // class maps then transition tables (16-byte aligned pieces)
    std::vector<BitsetUnitDesc> bitset_units;   // bundles too large for any DFA unit: bit-parallel NFA, every request (nfa_bits.hpp)
    std::vector<uint32_t> bitset_blob;          // their tables
    std::vector<uint32_t> acc_idx;
    std::vector<uint32_t> acc_events;  // event words (program.hpp), sorted by kind within each list
    std::vector<uint32_t> end_idx;
    std::vector<uint32_t> end_events;
    uint32_t n_atoms = 0, atom_words = 0;
    std::vector<uint32_t> expect;  // expected atom values (perf heuristic only)
    std::vector<uint32_t> care;    // atoms referenced by at least one rule
    std::vector<NsAtom> ns_atoms;            // grouped: integer predicates by feature (IntFeat order), then ip / country sets
    uint32_t ns_begin[N_INT_FEATS + 2] = {0};  // group g = ns_atoms[ns_begin[g], ns_begin[g + 1]); group N_INT_FEATS = the sets
    // quick reject per integer feature: lo <= x <= hi and (x < vmin or x > vmax) => every predicate on the feature is false
    int64_t ns_lo[N_INT_FEATS], ns_hi[N_INT_FEATS], ns_vmin[N_INT_FEATS], ns_vmax[N_INT_FEATS];
    uint32_t n_literal_atoms = 0;            // atoms confirmed by the gate's resolve kernel (finite string sets)
    uint32_t n_rare = 0;                     // INT_EXPR / FIELD_CMP predicates: the last n_rare entries of ns_atoms
    std::vector<int64_t> iexpr;              // INT_EXPR programs, flattened (program.hpp IntTok); NsAtom::set_id = offset
    std::vector<uint64_t> atom_sig;          // per atom: hashed set of the rules that mention it (all ones: never pair-independent)
    std::vector<uint16_t> code;
    std::vector<uint32_t> rule_off;  // n_rules + 1
    std::vector<uint8_t> term;       // per rule: terminal action for cv=0 (bits 0-1) and cv=1 (bits 2-3)
    std::vector<uint32_t> ar_idx, ar_rules;  // atom -> rules CSR
    std::vector<uint32_t> dflt_rules[2];     // rules true under `expect` with a terminal action, per captcha_verified
    uint32_t v0[2] = {0, 0};                 // verdict when every cared atom has its expected value
    std::vector<uint32_t> v1;                // [cv][atom]: verdict when exactly that one cared atom deviates
    std::vector<uint16_t> s1;                // [atom]: service in that case
    // service routes: rules [n_waf_rules, n_rules) of the same program (term = 0: the verdict loop skips them)
    uint32_t n_waf_rules = 0;
    std::vector<uint32_t> dflt_services;     // routes true under `expect` (rule indices)
    uint32_t s0 = 0xFFFFu;                   // service when every cared atom has its expected value
    std::vector<int64_t> iset_vals;
    std::vector<uint32_t> iset_off;
    std::vector<uint32_t> cset_words;
    GateTables gate[N_FIELDS];               // candidate gate of a field (present only if it has UM_CANDIDATES units)
    uint32_t vclean[2] = {0, 0};             // verdict of a request none of whose atoms is true, per captcha_verified
    uint32_t sclean = 0xFFFFu;               // its service
    std::vector<uint32_t> v1z;               // [cv][atom]: verdict when exactly that atom is true, every other one false
    std::vector<uint16_t> s1z;               // [atom]: service in that case
    int field_slot[N_FIELDS] = {-1, -1, -1, -1, -1};  // fields whose offsets the kernel stages
    uint32_t n_slots = 0;
    uint32_t scanned_fields_mask = 0;  // fields whose bytes are read (algorithmic-bytes accounting)
    int32_t gate_bypass_atom = -1;
    bool eval_gates = true;
    bool needs_ip = false, needs_geo_cols = false, needs_port = false;
    uint32_t n_rules = 0;
    LpmTables lpm;
    std::vector<std::string> warnings;
    std::string summary() const;
};

// `geo_mmdb` may be empty (no database: every client is {0,"XX"}, http_listener.rs:156).
bool compile_program(Model& model, const CompileOptions& opt, const std::vector<uint8_t>& geo_mmdb, HostProgram* out,
                     std::string& err);

// Shared-memory images, one per unit (the scan kernel re-stages shared memory for every unit, so each DFA gets the
// whole budget while it is being walked): class map, the rows of the first `hot_states` states (BFS order from the
// start state, so these are the shallow, frequently visited ones) with transitions to deeper states redirected to a
// trap row, and the acc1 / end1 event tables.  Fills units[u].{hot_states, hot_off, lim, acc1_off, end1_off, img_off,
// img_bytes}; `image` is the concatenation (each piece 256-byte aligned, at most `budget_bytes` long).
void build_unit_images(const HostProgram& prog, size_t budget_bytes, std::vector<uint8_t>* image, std::vector<UnitDesc>* units);

// lists (pingoo/lists.rs:62-113)
bool parse_list_csv(const std::string& name, ListType type, const uint8_t* csv, size_t len, ListData* out, std::string& err);
bool parse_ip_network(const std::string& s, IpNet* out, std::string& err);

// LPM construction (lpm.cpp)
bool build_lpm(const std::vector<std::vector<IpNet>>& ip_sets, const std::vector<uint8_t>& geo_mmdb, LpmTables* out,
               std::string& err);

}  // namespace pgw
