// Kernel parameter block shared by kernels.cu and the C-ABI host code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "program.hpp"

namespace pgw {

#ifndef PGW_THREADS
#define PGW_THREADS 640
#endif
constexpr int kThreads = PGW_THREADS;  // one persistent CTA per SM
#ifndef PGW_CHUNK
#define PGW_CHUNK 16
#endif
constexpr int kChunk = PGW_CHUNK;   // bytes per lane per scan iteration (PGW_CHUNK/16 128-bit loads)
constexpr int kRowsPerLane = 2;     // atom bitmaps per lane: the request being scanned + one awaiting its epilogue
constexpr uint32_t kClaim = 32;     // requests a warp claims from the global counter at a time

constexpr uint32_t kMaxConstNs = 16;
constexpr uint32_t kMaxConstUnits = 64;  // rule sets with more scan units run on the lane path (parameter block stays under 8 KB)

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
    uint32_t* work_counter;     // zeroed before each launch: next unclaimed request index
    // ---- program ----
    const UnitDesc* units;      // with hot_states / hot_off filled for the shared-memory image
    uint32_t n_units;
    const uint8_t* arena;       // full tables (global memory)
    const uint8_t* image;       // shared-memory image: class maps + hot rows
    uint32_t image_bytes;
    const uint32_t* acc_idx;
    const uint32_t* acc_events;
    const uint32_t* end_idx;
    const uint32_t* end_events;
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
    const uint32_t* v1;         // [cv][atom] verdict when exactly one cared atom deviates from `expect`
    const uint16_t* s1;         // [atom] service in that case
    // service routes (rules [n_waf_rules, n_rules)); `service` null: not requested for this batch
    uint32_t n_waf_rules;
    uint32_t s0;
    const uint32_t* dflt_services;
    uint32_t n_dflt_services;
    uint16_t* service;
    const int64_t* iset_vals;
    const uint32_t* iset_off;
    const uint32_t* cset;
    int32_t gate_atom;
    uint32_t eval_gates;
    uint32_t n_slots;           // scanned fields
    uint32_t slot_field[5];     // slot -> Field
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
    // ---- copy of the first unit descriptors in the parameter (constant) bank: the field-scan kernel reads them with a
    //      warp-uniform index, which keeps the per-unit parameters out of the vector register file ----
    UnitDesc udesc[kMaxConstUnits];
    NsAtom nsd[kMaxConstNs];    // likewise for the first non-scan atoms (read by every request's epilogue)
};

struct LaunchPlan {
    size_t smem_bytes;
    int grid;
};

// host-callable wrappers (kernels.cu)
size_t waf_smem_bytes(uint32_t image_bytes, uint32_t n_units, uint32_t atom_words, uint32_t n_slots);
size_t waf_smem_fixed_bytes(uint32_t n_units, uint32_t atom_words, uint32_t n_slots);  // everything except the image
const char* waf_launch(const KParams& p, const LaunchPlan& plan, void* stream);
// stream-scan path: scan kernel + epilogue kernel; `rows` = n * atom_words words of scratch, `task_counter` one word
size_t waf_stream_smem_bytes(uint32_t image_bytes, uint32_t n_units);
const char* waf_stream_launch(const KParams& p, uint32_t* rows, uint32_t* task_counter, int sm_count, size_t smem_bytes, void* stream);
// field-scan path: scan kernel + epilogue kernel; `rows` = n * atom_words words immediately followed by
// kFieldCounters claim counters (`counters` points at them)
constexpr uint32_t kFieldCounters = 64;
size_t waf_field_smem_bytes(uint32_t image_bytes, uint32_t n_units);
int waf_field_threads();  // threads per CTA of the field-scan kernel
const char* waf_field_launch(const KParams& p, uint32_t* rows, uint32_t* counters, int sm_count, size_t smem_bytes, void* stream,
                             cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);  // optional events around the scan kernel
const char* client_id_launch(const uint8_t* ip, const uint8_t* is_v6, const uint8_t* ua_bytes, const uint32_t* ua_off, const uint8_t* host_bytes,
                             const uint32_t* host_off, uint32_t n, uint8_t* out44, void* stream);
const char* geoip_launch(const KParams& p, const uint8_t* ip, const uint8_t* is_v6, uint32_t n, uint32_t* asn_out,
                         uint16_t* country_out, void* stream);
const char* waf_configure(int device, size_t* max_smem_optin, int* sm_count);

}  // namespace pgw
