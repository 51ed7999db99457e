// Kernel parameter block shared by kernels.cu and the C-ABI host code.
#pragma once
#include <cstdint>

#include "program.hpp"

namespace pgw {

constexpr int kThreads = 1024;      // one persistent CTA per SM, 32 warps
constexpr int kChunk = 16;          // bytes per lane per scan iteration (one 128-bit load)

struct KParams {
    // ---- batch (device pointers, SoA) ----
    const uint8_t* col[5];      // field bytes, Field order; 16-byte aligned, readable to round_up(len,16)
    const uint32_t* off[5];     // n+1 offsets per field
    const uint8_t* ip;          // n x 16, network order; IPv4 in bytes [0,4)
    const uint8_t* is_v6;       // n
    const int32_t* port;        // n
    const int64_t* asn;         // n or null (null: resolve through the loaded GeoIP table, else 0)
    const uint16_t* country;    // n or null
    const uint8_t* flags;       // n or null
    uint32_t* verdict;          // n
    uint32_t n;
    uint32_t tile_log2;
    uint32_t n_tiles;
    // ---- program ----
    const UnitDesc* units;
    uint32_t n_units;
    const uint8_t* arena;
    uint32_t arena_bytes;
    uint32_t cls_bytes;         // leading part of the arena holding the class maps
    const uint32_t* acc_idx;
    const uint16_t* acc_atoms;
    const uint32_t* end_idx;
    const uint16_t* end_atoms;
    uint32_t n_atoms, atom_words;
    const uint32_t* expect;
    const uint32_t* care;
    const NsAtom* ns;
    uint32_t n_ns;
    const uint16_t* code;
    const uint32_t* rule_off;
    const uint8_t* term;
    uint32_t n_rules;
    const uint32_t* ar_idx;
    const uint32_t* ar_rules;
    const uint32_t* dflt[2];
    uint32_t n_dflt[2];
    uint32_t v0[2];
    const int64_t* iset_vals;
    const uint32_t* iset_off;
    const uint32_t* cset;
    int32_t slot[5];
    uint32_t n_slots;
    int32_t gate_atom;
    uint32_t eval_gates;
    // ---- longest-prefix tables ----
    const uint32_t* dir24;
    const uint32_t* tbl8;
    const LpmLeaf* leaves;
    const uint64_t* v6_hi;
    const uint64_t* v6_lo;
    const uint32_t* v6_leaf;
    uint32_t n_v6;
    uint32_t lpm_present;
    uint32_t geo_loaded;
    uint32_t need_lpm;          // any ip-set atom, or geo columns needed and resolved on device
};

struct LaunchPlan {
    bool smem_tables;
    uint32_t tile_log2;
    size_t smem_bytes;
    int grid;
};

// host-callable wrappers (kernels.cu)
size_t waf_smem_bytes(const KParams& p, bool smem_tables, uint32_t tile_log2);
const char* waf_launch(const KParams& p, const LaunchPlan& plan, void* stream);
const char* geoip_launch(const KParams& p, const uint8_t* ip, const uint8_t* is_v6, uint32_t n, uint32_t* asn_out,
                         uint16_t* country_out, void* stream);
const char* waf_configure(int device, size_t* max_smem_optin, int* sm_count);

}  // namespace pgw
