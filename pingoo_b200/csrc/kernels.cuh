// Kernel parameter block shared by kernels.cu and the C-ABI host code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gate.hpp"
#include "program.hpp"

namespace pgw {

constexpr uint32_t kMaxConstNs = 16;
constexpr uint32_t kMaxConstUnits = 64;  // scan units per scan-kernel launch (their descriptors ride in the parameter bank);
                                         // rule sets with more units are scanned by several launches
constexpr uint32_t kMaxGateFields = 3;   // gated fields: url, user_agent, path
constexpr uint32_t kMaxPrefixUnits = 8;  // small early-exit units walked inside the per-request (epilogue) kernel

struct KParams {
    // ---- batch (device pointers, SoA) ----
    const uint8_t* col[5];      // field bytes, Field order; 32-byte aligned, readable to round_up(len, 32)
    const uint32_t* off[5];     // n+1 offsets per field
    const uint8_t* ip;          // n x 16, network order; IPv4 in bytes [0,4)
    const uint8_t* is_v6;       // n
    const int32_t* port;        // n
    const int64_t* asn;         // n or null (null: resolve through the loaded GeoIP table, else 0)
    const uint16_t* country;    // n or null
    const uint8_t* flags;       // n or null
    uint32_t* verdict;          // n
    uint32_t n;
    // ---- per-launch scratch ----
    uint32_t* rows;             // n x atom_words atom bitmaps; all zero between batches (the epilogue re-zeroes what was touched)
    uint32_t* info;             // 2 words per request, zero between batches: [0] = largest fired atom + 1 (0: none fired),
                                // [1] = 0x4000 - smallest fired atom; one distinct atom fired <=> [0] - 1 == 0x4000 - [1]
    uint32_t* multi_count;      // requests with several true atoms: list filled by the epilogue, evaluated by waf_multi_kernel
    uint32_t* multi_list;
    uint32_t* counters;         // per scan unit: next unclaimed request / candidate (zeroed before each batch)
    const uint32_t* cand_count[5];  // gate candidates of a field (null: the field has no gate)
    const uint32_t* cand_idx[5];
    const uint32_t* cand_start[5];
    const uint32_t* cand_end[5];
    const uint32_t* cand_mask[5];   // per candidate: the gated units of the field it is a candidate for (UnitDesc::gate_bit)
    // ---- program ----
    uint32_t n_units;           // units of THIS launch: udesc[0 .. n_units)
    uint32_t unit_base;         // index of udesc[0] in the program (claim counter = counters[unit_base + u])
    uint32_t n_units_total;
    const UnitDesc* units;      // all units (global memory copy, read by the epilogue beyond the parameter bank)
    const uint8_t* arena;       // full tables (global memory)
    const uint8_t* images;      // per-unit shared-memory images (UnitDesc::img_off / img_bytes)
    const uint32_t* acc_idx;
    const uint32_t* acc_events;
    const uint32_t* end_idx;
    const uint32_t* end_events;
    uint32_t n_atoms, atom_words;
    const uint32_t* expect;
    const uint32_t* care;
    const NsAtom* ns;           // grouped by integer feature, then the ip / country sets
    uint32_t n_ns;
    uint32_t ns_begin[9];       // group g = ns[ns_begin[g], ns_begin[g + 1]); groups 0..6 = IntFeat, 7 = sets
    uint32_t n_feat_used, feat_used[7];   // the integer features that have predicates (groups with ns_begin[g] < ns_begin[g + 1])
    uint32_t rare_begin, n_rare;  // the INT_EXPR / FIELD_CMP predicates: the last n_rare (<= 64) entries of ns
    int64_t ns_lo[7], ns_hi[7], ns_vmin[7], ns_vmax[7];  // quick reject of a whole integer group (compile.hpp)
    const uint16_t* code;
    const uint32_t* rule_off;
    const uint8_t* term;
    uint32_t n_rules;
    const uint32_t* ar_idx;
    const uint32_t* ar_rules;
    const uint32_t* dflt[2];
    uint32_t n_dflt[2];
    uint32_t v0[2];
    const uint32_t* v1;         // [cv][atom] verdict when exactly one cared atom deviates from `expect`
    const uint16_t* s1;         // [atom] service in that case
    uint32_t vclean[2];         // verdict of a request whose atom bitmap is all zero
    uint32_t sclean;            // its service
    const uint32_t* v1z;        // [cv][atom] verdict when exactly that atom is true and every other one false
    const uint16_t* s1z;        // [atom] service in that case
    const uint64_t* atom_sig;   // [atom] hashed set of the rules mentioning the atom: two true atoms with disjoint signatures combine
                                // their v1z / s1z entries (compile.cpp)
    // service routes (rules [n_waf_rules, n_rules)); `service` null: not requested for this batch
    uint32_t n_waf_rules;
    uint32_t s0;
    const uint32_t* dflt_services;
    uint32_t n_dflt_services;
    uint16_t* service;
    const int64_t* iset_vals;
    const int64_t* iexpr;       // INT_EXPR programs (program.hpp IntTok)
    const uint32_t* iset_off;
    const uint32_t* cset;
    int32_t gate_atom;
    uint32_t eval_gates;
    // units whose start state has end-of-field events: the epilogue finishes their EMPTY fields (they never reach the scan)
    uint32_t n_start_end;
    uint32_t start_end_unit[8];
    // ---- longest-prefix tables ----
    const uint32_t* dir24;
    const uint32_t* tbl8;
    const LpmLeaf* leaves;
    const uint64_t* v6_hi;
    const uint64_t* v6_lo;
    const uint32_t* v6_leaf;
    const uint32_t* v6_top;     // [65537] range index on the first 16 address bits
    uint32_t n_v6;
    uint32_t lpm_present;
    uint32_t geo_loaded;
    uint32_t need_lpm;          // any ip-set atom, or geo columns needed and resolved on device
    // ---- bit-parallel NFA units (nfa_bits.hpp): bundles no DFA unit can hold; one extra kernel, only when present ----
    const BitsetUnitDesc* bitset_units;
    const uint32_t* bitset_blob;
    uint32_t n_bitset;
    uint32_t bitset_smem_words;   // dynamic shared memory of that kernel: the largest unit table that fits the budget
    // ---- small early-exit units walked by the epilogue kernel (UM_PREPASS): descriptors here, tables in its shared memory ----
    uint32_t n_prefix;
    uint32_t prefix_area;                     // bytes of shared memory their images take (multiple of 256)
    uint32_t prefix_img[kMaxPrefixUnits];     // offset of unit k's image inside that area
    UnitDesc pdesc[kMaxPrefixUnits];
    // ---- unit descriptors of this launch in the parameter (constant) bank: the scan kernel reads them with a
    //      warp-uniform index, which keeps the per-unit parameters out of the vector register file ----
    UnitDesc udesc[kMaxConstUnits];
    NsAtom nsd[kMaxConstNs];    // likewise for the first non-scan atoms (read by every request's epilogue)
};

// candidate gate (kernel_gate.cuh): one entry per gated field
struct GateField {
    const uint8_t* col;      // field bytes
    const uint32_t* off;     // n + 1 offsets
    const uint32_t* b1;      // level 1: blocked Bloom filter (2^k1 bits), global memory copy (staged into shared memory)
    const uint32_t* slots;   // level 2: exact table, 2^kt slots of {gram, unit mask} or, with GateParams::wide_slots,
                             // {gram, unit mask, first literal candidate, their number}
    const uint32_t* lit_cand;   // literal candidates of the grams: (literal << 2) | (delta + 1)
    const LitDesc* lits;        // finite-string patterns confirmed by the resolve kernel (gate.hpp)
    const uint8_t* lit_bytes;
    uint32_t k1, kt;
    uint32_t bloom_off;      // byte offset of the field's bitmap in the gate kernel's shared memory (all fields resident)
    uint32_t* bitmap;        // hit bitmap: one bit per 16-byte chunk of the column (index = column position >> 4); the gate
                             // kernel writes every word that covers the batch's bytes
    uint32_t* maybe_count;   // requests whose field overlaps a hit chunk: one counter ...
    uint32_t* maybe_idx;     // ... and the request indices
    uint32_t* cand_count;    // candidate list of the field (confirmed by the exact gram table): one counter ...
    uint32_t* cand_idx;      // ... and request index / field start / field end / unit mask per candidate
    uint32_t* cand_start;
    uint32_t* cand_end;
    uint32_t* cand_mask;
};

struct GateParams {
    GateField f[kMaxGateFields];
    uint32_t n_fields;
    uint32_t n;              // requests
    uint32_t wide_slots;     // the rule set confirms literals in the resolve step (gate.hpp): 4-word slots, waf_gate_resolve_lit_kernel
    // where a confirmed literal's atom goes: the request's bitmap row and info words (as KParams)
    uint32_t* rows;
    uint32_t* info;
    uint32_t atom_words;
};

// host-callable wrappers (kernels.cu)
size_t waf_scan_smem_bytes(uint32_t max_image_bytes);
size_t waf_scan_image_budget(size_t max_smem_optin);  // bytes a unit image may take
int waf_scan_threads();
size_t waf_gate_smem_bytes(const GateParams& g);
size_t waf_bitset_smem_budget();   // bytes of shared memory a bit-parallel NFA unit's tables may take (larger ones stay in global memory)
size_t waf_prefix_budget();  // shared memory the images of all early-exit units walked by the epilogue kernel may take
// One batch: [gate -> maybe -> resolve] -> scan (one launch per kMaxConstUnits units) [-> bitset NFA] -> epilogue -> multi, all on
// `stream`.  `all_units` = the program's unit descriptors (host copy); `small` = the block of claim counters, candidate
// counters and list counters to zero first (`small_words` words).  `ev` (optional): four events recorded before the
// gate kernels, after them, after the scan launches and after the epilogue + multi kernels.
const char* waf_batch_launch(KParams& p, GateParams& g, const UnitDesc* all_units, uint32_t* small, uint32_t small_words, int sm_count,
                             size_t scan_smem, size_t gate_smem, void* stream, cudaEvent_t* ev = nullptr, uint32_t* launches = nullptr);
const char* client_id_launch(const uint8_t* ip, const uint8_t* is_v6, const uint8_t* ua_bytes, const uint32_t* ua_off, const uint8_t* host_bytes,
                             const uint32_t* host_off, uint32_t n, uint8_t* out44, void* stream);
const char* geoip_launch(const KParams& p, const uint8_t* ip, const uint8_t* is_v6, uint32_t n, uint32_t* asn_out,
                         uint16_t* country_out, void* stream);
const char* waf_configure(int device, size_t* max_smem_optin, int* sm_count);

}  // namespace pgw
