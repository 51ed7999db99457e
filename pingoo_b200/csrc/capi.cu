// extern "C" boundary (include/pingoo_waf.h): ruleset lifecycle, device upload,
// launch planning and the device / host-pointer evaluate entry points.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pingoo_waf.h"
#include "kernels.cuh"
#include "ruleset.hpp"

using namespace pgw;

namespace {

thread_local std::string g_last_error;

int fail(const std::string& msg, char* err, size_t cap) {
    g_last_error = msg;
    if (err && cap) {
        size_t n = msg.size() < cap - 1 ? msg.size() : cap - 1;
        memcpy(err, msg.data(), n);
        err[n] = 0;
    }
    return 1;
}

struct DevMem {
    std::vector<void*> ptrs;
    template <class T>
    const T* upload(const std::vector<T>& v, size_t pad_to = 16) {
        size_t bytes = v.size() * sizeof(T);
        size_t alloc = ((bytes + pad_to - 1) / pad_to) * pad_to;
        if (alloc == 0) alloc = pad_to;
        void* d = nullptr;
        if (cudaMalloc(&d, alloc) != cudaSuccess) return nullptr;
        ptrs.push_back(d);
        cudaMemset(d, 0, alloc);
        if (bytes && cudaMemcpy(d, v.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        return (const T*)d;
    }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

constexpr uint32_t kProfRing = 256;
constexpr uint32_t kHostSlices = 16;  // host-pointer path: copy/compute pipeline depth

struct Staging {
    void* d = nullptr;
    size_t cap = 0;
    bool ensure(size_t bytes) {
        if (bytes <= cap) return true;
        if (d) cudaFree(d);
        d = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        if (cudaMalloc(&d, want) != cudaSuccess) return false;
        cap = want;
        return true;
    }
    void release() {
        if (d) cudaFree(d);
        d = nullptr;
        cap = 0;
    }
};

// Per-batch device scratch.  rows / dirty obey the invariant "all zero between batches" (the epilogue re-zeroes what
// the scan touched), so a batch costs no memset of the n x atom_words bitmap; `small` (claim counters + candidate
// counters) is zeroed at every launch.  Blocks are reused in stream order: a block released on stream S can be taken
// again on S at once, on another stream only after its event has completed.
struct Scratch {
    uint8_t* base = nullptr;
    size_t cap_requests = 0;
    cudaEvent_t done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool busy = false;
    // carved out of `base`
    uint32_t* rows = nullptr;
    uint32_t* info = nullptr;
    uint32_t* small = nullptr;       // [0, kSmallCounters): per-unit claim counters; [kSmallCounters, +5): candidate counters; [+5]: multi count;
                                     // [+8, +13): counters of the gate's "maybe" lists
    uint32_t* multi = nullptr;       // multi list (request indices)
    uint32_t* cand[5][5] = {};       // per gated field: idx, start, end, unit mask; [4] = the "maybe" list (request indices)
    uint32_t* bitmap[5] = {};        // per gated field: hit bitmap, one bit per 16-byte chunk of the column
};
constexpr uint32_t kSmallCounters = 1024;   // scan units a program may have
constexpr uint32_t kSmallWords = kSmallCounters + 16;
constexpr size_t kHitBitmapBytes = (size_t)32 << 20;  // a column holds less than 4 GiB (32-bit offsets): 2^28 chunks of 16 bytes

struct HostStage {   // device staging of one host-pointer call (pgw_evaluate_batch_host), pooled so that calls may overlap
    Staging cols[5], offs[5], ip, v6, port, asn, country, flags, verdict, service;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t slice_ready[kHostSlices] = {};
    bool busy = false;
    bool init() {
        bool ok = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking) == cudaSuccess;
        for (uint32_t i = 0; i < kHostSlices && ok; ++i) ok = cudaEventCreateWithFlags(&slice_ready[i], cudaEventDisableTiming) == cudaSuccess;
        return ok;
    }
    void release() {
        for (int f = 0; f < 5; ++f) { cols[f].release(); offs[f].release(); }
        ip.release(); v6.release(); port.release(); asn.release(); country.release(); flags.release(); verdict.release(); service.release();
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        for (auto& ev : slice_ready)
            if (ev) cudaEventDestroy(ev);
    }
};

}  // namespace

struct pgw_ruleset {
    RulesetBuilder builder;
    HostProgram prog;
    bool finalized = false;
    int device = -1;
    int sm_count = 0;
    size_t max_smem = 0;
    DevMem mem;
    KParams base;  // program pointers filled in, batch fields zero
    GateParams gate_base;  // bitmaps and shifts filled in
    int gate_field[kMaxGateFields] = {0, 0, 0};
    std::vector<UnitDesc> units;   // with the image fields filled in
    size_t scan_smem = 0, gate_smem = 0;
    uint32_t hot_states_total = 0;
    std::atomic<uint64_t> launches{0};
    // scratch pool (device-pointer and host-pointer paths)
    std::mutex pool_mu;
    std::vector<Scratch*> pool;
    // host-pointer path: staging pool
    std::vector<HostStage*> stages;
    std::atomic<uint64_t> last_h2d{0}, last_d2h{0};
    // measurement hook: event pairs around the gate + scan kernels
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;  // 2 * kProfRing events, created on first enable
    std::atomic<uint64_t> prof_n{0};
};

namespace {

Scratch* scratch_acquire(pgw_ruleset* rs, uint32_t n, cudaStream_t stream, std::string& e) {
    const HostProgram& H = rs->prog;
    std::lock_guard<std::mutex> lk(rs->pool_mu);
    for (Scratch* sc : rs->pool) {
        if (sc->busy || sc->cap_requests < n) continue;
        if (sc->last_stream != stream && cudaEventQuery(sc->done) != cudaSuccess) continue;
        sc->busy = true;
        return sc;
    }
    // a new block, sized with head-room so that slightly larger batches reuse it
    Scratch* sc = new Scratch();
    size_t cap = (size_t)n + n / 8 + 1024;
    cap = (cap + 31) & ~(size_t)31;
    size_t n_gated = 0;
    for (int f = 0; f < 5; ++f) n_gated += H.gate[f].present ? 1 : 0;
    const size_t rows_b = cap * H.atom_words * 4, dirty_b = cap * 8 + 64, small_b = kSmallWords * 4,
                 cand_b = (n_gated * 5 + 1) * cap * 4 + n_gated * kHitBitmapBytes;
    const size_t total = rows_b + dirty_b + small_b + cand_b + 1024;
    if (cudaMalloc((void**)&sc->base, total) != cudaSuccess || cudaEventCreateWithFlags(&sc->done, cudaEventDisableTiming) != cudaSuccess) {
        e = std::string("CUDA: scratch allocation failed (") + std::to_string(total >> 20) + " MiB): " + cudaGetErrorString(cudaGetLastError());
        if (sc->base) cudaFree(sc->base);
        delete sc;
        return nullptr;
    }
    // the zero invariant starts here; ordered before the first kernel that uses the block
    if (cudaMemsetAsync(sc->base, 0, rows_b + dirty_b + small_b, stream) != cudaSuccess) { e = "CUDA: scratch memset failed"; cudaFree(sc->base); delete sc; return nullptr; }
    uint8_t* q = sc->base;
    sc->rows = (uint32_t*)q; q += rows_b;
    sc->info = (uint32_t*)q; q += dirty_b;
    sc->small = (uint32_t*)q; q += small_b;
    sc->multi = (uint32_t*)q; q += cap * 4;
    for (int f = 0; f < 5; ++f)
        if (H.gate[f].present) {
            for (int k = 0; k < 5; ++k) { sc->cand[f][k] = (uint32_t*)q; q += cap * 4; }
            sc->bitmap[f] = (uint32_t*)q; q += kHitBitmapBytes;
        }
    sc->cap_requests = cap;
    sc->busy = true;
    rs->pool.push_back(sc);
    return sc;
}

void scratch_release(pgw_ruleset* rs, Scratch* sc, cudaStream_t stream) {
    cudaEventRecord(sc->done, stream);
    std::lock_guard<std::mutex> lk(rs->pool_mu);
    sc->last_stream = stream;
    sc->busy = false;
}

}  // namespace

extern "C" {

const char* pgw_last_error(void) { return g_last_error.c_str(); }

int pgw_compile_expression(const char* expression, char* err, size_t err_cap) {
    if (!expression) return fail("Expression is not valid: null", err, err_cap);
    std::string e;
    if (!RulesetBuilder::compile_expression(expression, e)) return fail(e, err, err_cap);
    return 0;
}

int pgw_validate_expression(const char* expression, char* err, size_t err_cap) {
    if (!expression) return fail("Expression is not valid: null", err, err_cap);
    std::string e;
    if (!RulesetBuilder::validate_expression(expression, e)) return fail(e, err, err_cap);
    return 0;
}

int pgw_ruleset_create(const pgw_rule_desc* rules, uint32_t n_rules, const pgw_options* options, pgw_ruleset** out,
                       char* err, size_t err_cap) {
    if (!out) return fail("out is null", err, err_cap);
    *out = nullptr;
    if (n_rules && !rules) return fail("rules is null", err, err_cap);
    pgw_ruleset* rs = new pgw_ruleset();
    memset(&rs->base, 0, sizeof rs->base);
    if (options) {
        if (options->max_dfa_states > 0) rs->builder.options.max_dfa_states = options->max_dfa_states;
        if (options->max_unit_table_bytes > 0) rs->builder.options.max_unit_table_bytes = (size_t)options->max_unit_table_bytes;
        rs->builder.options.eval_gates = options->eval_gates != 0;
        rs->builder.options.candidate_gate = (options->disable_candidate_gate & 1) == 0;
        rs->builder.options.literal_confirm = (options->disable_candidate_gate & 2) == 0;
    }
    for (uint32_t i = 0; i < n_rules; ++i) {
        std::string e;
        if (!rs->builder.add_rule(rules[i].name, rules[i].expression, rules[i].actions, rules[i].n_actions, e)) {
            delete rs;
            return fail(e, err, err_cap);
        }
    }
    *out = rs;
    return 0;
}

int pgw_ruleset_load_dir(const char* config_folder, const char* listener, const char* const* geoip_dirs, uint32_t n_geoip_dirs,
                         const pgw_options* options, pgw_ruleset** out, char* err, size_t err_cap) {
    if (!out || !config_folder) return fail("null argument", err, err_cap);
    *out = nullptr;
    pgw_ruleset* rs = new pgw_ruleset();
    memset(&rs->base, 0, sizeof rs->base);
    if (options) {
        if (options->max_dfa_states > 0) rs->builder.options.max_dfa_states = options->max_dfa_states;
        if (options->max_unit_table_bytes > 0) rs->builder.options.max_unit_table_bytes = (size_t)options->max_unit_table_bytes;
        rs->builder.options.eval_gates = options->eval_gates != 0;
        rs->builder.options.candidate_gate = (options->disable_candidate_gate & 1) == 0;
        rs->builder.options.literal_confirm = (options->disable_candidate_gate & 2) == 0;
    }
    std::vector<std::string> dirs;
    for (uint32_t i = 0; i < n_geoip_dirs; ++i) dirs.push_back(geoip_dirs[i]);
    if (!n_geoip_dirs) { dirs.push_back(config_folder); dirs.push_back("/usr/share/pingoo"); }  // config.rs:31-36
    std::string e;
    if (!load_config_dir(config_folder, listener, dirs, &rs->builder, nullptr, e)) {
        delete rs;
        return fail(e, err, err_cap);
    }
    *out = rs;
    return 0;
}

int pgw_lists_add(pgw_ruleset* rs, const char* name, int list_type, const uint8_t* csv, size_t csv_len, char* err,
                  size_t err_cap) {
    if (!rs || !name) return fail("null argument", err, err_cap);
    std::string e;
    if (!rs->builder.add_list(name, list_type, csv, csv_len, e)) return fail(e, err, err_cap);
    return 0;
}

int pgw_services_set(pgw_ruleset* rs, const pgw_service_desc* services, uint32_t n, char* err, size_t err_cap) {
    if (!rs || (n && !services)) return fail("null argument", err, err_cap);
    if (rs->finalized) return fail("ruleset already finalized", err, err_cap);
    if (rs->builder.n_services()) return fail("services are already set", err, err_cap);
    std::string e;
    for (uint32_t i = 0; i < n; ++i)
        if (!rs->builder.add_service(services[i].name, services[i].route, e)) return fail(e, err, err_cap);
    return 0;
}

int pgw_geoip_load(pgw_ruleset* rs, const uint8_t* mmdb, size_t mmdb_len, char* err, size_t err_cap) {
    if (!rs || !mmdb) return fail("null argument", err, err_cap);
    std::string e;
    if (!rs->builder.load_geoip(mmdb, mmdb_len, e)) return fail(e, err, err_cap);
    return 0;
}

int pgw_ruleset_finalize(pgw_ruleset* rs, int device, char* err, size_t err_cap) {
    if (!rs) return fail("null ruleset", err, err_cap);
    if (rs->finalized) return fail("ruleset already finalized", err, err_cap);
    // The device is probed first: without a CUDA device there is nothing this engine can do.
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); this engine has no CPU fallback", err, err_cap);
    if (device < 0 || device >= ndev) return fail("invalid device index", err, err_cap);
    if (const char* m = waf_configure(device, &rs->max_smem, &rs->sm_count)) return fail(std::string("CUDA: ") + m, err, err_cap);

    std::string e;
    if (!rs->builder.finalize(&rs->prog, e)) return fail(e, err, err_cap);
    const HostProgram& H = rs->prog;
    rs->device = device;

    KParams& P = rs->base;
    DevMem& M = rs->mem;
    bool ok = true;
    auto chk = [&](const void* p) { if (!p) ok = false; return p; };
    // per-unit shared-memory images: each unit gets the whole budget while it is being walked
    if (H.units.size() > kSmallCounters) return fail("ruleset has too many scan units", err, err_cap);
    std::vector<uint8_t> image;
    std::vector<UnitDesc>& units = rs->units;
    size_t image_budget = waf_scan_image_budget(rs->max_smem);
    build_unit_images(H, image_budget, &image, &units);
    uint32_t max_img = 0;
    for (auto& u : units) {
        rs->hot_states_total += u.hot_states;
        if (u.img_bytes > max_img) max_img = u.img_bytes;
    }
    rs->scan_smem = waf_scan_smem_bytes(max_img);
    if (rs->scan_smem > rs->max_smem) return fail("ruleset does not fit the shared-memory plan", err, err_cap);
    memset(P.udesc, 0, sizeof P.udesc);
    P.n_units_total = (uint32_t)units.size();
    P.n_units = 0;
    P.unit_base = 0;
    P.arena = (const uint8_t*)chk(M.upload(H.arena));
    P.images = (const uint8_t*)chk(M.upload(image, 256));
    P.n_start_end = 0;
    for (size_t u = 0; u < units.size(); ++u)
        if (units[u].start_end) {
            if (P.n_start_end >= 8) return fail("too many scan units with patterns that match an empty field", err, err_cap);
            P.start_end_unit[P.n_start_end++] = (uint32_t)u;
        }
    // small early-exit units (start-anchored patterns, whole table in shared memory): walked by the epilogue kernel, one
    // thread per request, instead of costing a pass of the scan kernel each; their images sit next to each other in its
    // shared memory
    P.n_prefix = 0;
    {
        size_t used = 0;
        for (size_t u = 0; u < units.size() && P.n_prefix < kMaxPrefixUnits; ++u) {
            UnitDesc& ud = units[u];
            if (ud.mode != UM_ALL || ud.abs0 == 0xFFFFFFFFu || ud.hot_states != ud.n_states) continue;
            const size_t need = ((size_t)ud.img_bytes + 255) & ~(size_t)255;
            if (used + need > waf_prefix_budget()) continue;
            ud.mode = UM_PREPASS;
            P.prefix_img[P.n_prefix] = (uint32_t)used;
            P.pdesc[P.n_prefix++] = ud;
            used += need;
        }
        P.prefix_area = (uint32_t)used;
    }
    // candidate gate: the level-1 bitmaps of all gated fields are resident in the gate kernel's shared memory together
    GateParams& G = rs->gate_base;
    memset(&G, 0, sizeof G);
    static const int kGateOrder[3] = {F_URL, F_USER_AGENT, F_PATH};
    size_t bloom_used = 0;
    for (int fo = 0; fo < 3; ++fo) {
        const int f = kGateOrder[fo];
        if (!H.gate[f].present) continue;
        GateField gf;
        memset(&gf, 0, sizeof gf);
        gf.b1 = (const uint32_t*)chk(M.upload(H.gate[f].b1));
        gf.slots = (const uint32_t*)chk(M.upload(H.gate[f].slots));
        gf.lit_cand = (const uint32_t*)chk(M.upload(H.gate[f].lit_cand));
        gf.lits = (const LitDesc*)chk(M.upload(H.gate[f].lits));
        gf.lit_bytes = (const uint8_t*)chk(M.upload(H.gate[f].lit_bytes));
        if (H.gate[f].slot_words == 4) G.wide_slots = 1u;
        gf.k1 = H.gate[f].k1;
        gf.kt = H.gate[f].kt;
        gf.bloom_off = (uint32_t)bloom_used;
        bloom_used += ((size_t)1 << gf.k1) / 8;
        rs->gate_field[G.n_fields] = f;
        G.f[G.n_fields++] = gf;
    }
    // the scan kernel reads the unit descriptors from the host copy (parameter bank) and the epilogue from P.units: both
    // must see the UM_PREPASS marks
    P.units = (const UnitDesc*)chk(M.upload(units));
    rs->gate_smem = waf_gate_smem_bytes(G);
    if (rs->gate_smem > rs->max_smem) return fail("candidate-gate bitmaps do not fit shared memory", err, err_cap);
    // bit-parallel NFA units: descriptors and tables in global memory; the kernel stages a unit's tables when they fit
    P.n_bitset = (uint32_t)H.bitset_units.size();
    if (P.n_bitset > 65535u) return fail("too many patterns that need the bit-parallel NFA unit", err, err_cap);
    P.bitset_units = (const BitsetUnitDesc*)chk(M.upload(H.bitset_units));
    P.bitset_blob = (const uint32_t*)chk(M.upload(H.bitset_blob));
    P.bitset_smem_words = 0;
    for (const BitsetUnitDesc& b : H.bitset_units)
        if ((size_t)b.blob_words * 4 <= waf_bitset_smem_budget() && b.blob_words > P.bitset_smem_words) P.bitset_smem_words = b.blob_words;
    P.acc_idx = (const uint32_t*)chk(M.upload(H.acc_idx));
    P.acc_events = (const uint32_t*)chk(M.upload(H.acc_events));
    P.end_idx = (const uint32_t*)chk(M.upload(H.end_idx));
    P.end_events = (const uint32_t*)chk(M.upload(H.end_events));
    P.n_atoms = H.n_atoms;
    P.atom_words = H.atom_words;
    P.expect = (const uint32_t*)chk(M.upload(H.expect));
    P.care = (const uint32_t*)chk(M.upload(H.care));
    P.ns = (const NsAtom*)chk(M.upload(H.ns_atoms));
    P.n_ns = (uint32_t)H.ns_atoms.size();
    P.n_rare = H.n_rare;
    P.rare_begin = P.n_ns - H.n_rare;
    for (int g = 0; g < 9; ++g) P.ns_begin[g] = H.ns_begin[g];
    P.n_feat_used = 0;
    for (uint32_t g = 0; g < 7; ++g)
        if (H.ns_begin[g] < H.ns_begin[g + 1]) P.feat_used[P.n_feat_used++] = g;
    for (int f = 0; f < 7; ++f) { P.ns_lo[f] = H.ns_lo[f]; P.ns_hi[f] = H.ns_hi[f]; P.ns_vmin[f] = H.ns_vmin[f]; P.ns_vmax[f] = H.ns_vmax[f]; }
    memset(P.nsd, 0, sizeof P.nsd);
    for (size_t i = 0; i < H.ns_atoms.size() && i < kMaxConstNs; ++i) P.nsd[i] = H.ns_atoms[i];
    P.code = (const uint16_t*)chk(M.upload(H.code));
    P.rule_off = (const uint32_t*)chk(M.upload(H.rule_off));
    P.term = (const uint8_t*)chk(M.upload(H.term));
    P.n_rules = H.n_rules;
    P.v1 = (const uint32_t*)chk(M.upload(H.v1));
    P.s1 = (const uint16_t*)chk(M.upload(H.s1));
    P.n_waf_rules = H.n_waf_rules;
    P.s0 = H.s0;
    P.vclean[0] = H.vclean[0];
    P.vclean[1] = H.vclean[1];
    P.sclean = H.sclean;
    P.v1z = (const uint32_t*)chk(M.upload(H.v1z));
    P.s1z = (const uint16_t*)chk(M.upload(H.s1z));
    P.atom_sig = (const uint64_t*)chk(M.upload(H.atom_sig));
    P.dflt_services = (const uint32_t*)chk(M.upload(H.dflt_services));
    P.n_dflt_services = (uint32_t)H.dflt_services.size();
    P.service = nullptr;
    P.ar_idx = (const uint32_t*)chk(M.upload(H.ar_idx));
    P.ar_rules = (const uint32_t*)chk(M.upload(H.ar_rules));
    for (int cv = 0; cv < 2; ++cv) {
        P.dflt[cv] = (const uint32_t*)chk(M.upload(H.dflt_rules[cv]));
        P.n_dflt[cv] = (uint32_t)H.dflt_rules[cv].size();
        P.v0[cv] = H.v0[cv];
    }
    P.iset_vals = (const int64_t*)chk(M.upload(H.iset_vals));
    P.iset_off = (const uint32_t*)chk(M.upload(H.iset_off));
    P.iexpr = (const int64_t*)chk(M.upload(H.iexpr));
    P.cset = (const uint32_t*)chk(M.upload(H.cset_words));
    P.gate_atom = H.gate_bypass_atom;
    P.eval_gates = H.eval_gates ? 1u : 0u;
    P.lpm_present = H.lpm.present ? 1u : 0u;
    P.geo_loaded = H.lpm.geo_loaded ? 1u : 0u;
    if (H.lpm.present) {
        P.dir24 = (const uint32_t*)chk(M.upload(H.lpm.dir24));
        P.tbl8 = (const uint32_t*)chk(M.upload(H.lpm.tbl8));
        P.leaves = (const LpmLeaf*)chk(M.upload(H.lpm.leaves));
        P.v6_hi = (const uint64_t*)chk(M.upload(H.lpm.v6_hi));
        P.v6_lo = (const uint64_t*)chk(M.upload(H.lpm.v6_lo));
        P.v6_leaf = (const uint32_t*)chk(M.upload(H.lpm.v6_leaf));
        P.v6_top = (const uint32_t*)chk(M.upload(H.lpm.v6_top));
        P.n_v6 = (uint32_t)H.lpm.v6_leaf.size();
    }
    if (!ok) {
        M.release();
        return fail(std::string("CUDA: device allocation/upload failed: ") + cudaGetErrorString(cudaGetLastError()), err, err_cap);
    }
    rs->finalized = true;
    return 0;
}

static int launch_on(pgw_ruleset* rs, const pgw_batch* b, uint32_t* verdict_out, void* stream, std::string& e, uint16_t* service_out = nullptr) {
    const HostProgram& H = rs->prog;
    KParams P = rs->base;
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    for (int f = 0; f < 5; ++f) {
        P.col[f] = cols[f]->bytes;
        P.off[f] = cols[f]->offsets;
        if (H.field_slot[f] >= 0 && !cols[f]->offsets) { e = std::string("batch is missing offsets for http_request.") + kFieldNames[f]; return 1; }
        if (((H.scanned_fields_mask >> f) & 1) && !cols[f]->bytes) { e = std::string("batch is missing bytes for http_request.") + kFieldNames[f]; return 1; }
        if (((H.scanned_fields_mask >> f) & 1) && ((uintptr_t)cols[f]->bytes & 31)) { e = std::string("bytes of http_request.") + kFieldNames[f] + " are not 32-byte aligned"; return 1; }
    }
    P.ip = b->ip;
    P.is_v6 = b->ip_is_v6;
    P.port = b->remote_port;
    // the geo columns count only when BOTH are supplied (the reference resolves asn and country together, before the
    // context is built: http_listener.rs:143-157); otherwise both come from the loaded database, or are {0, "XX"}
    const bool geo_cols = b->asn && b->country;
    P.asn = geo_cols ? b->asn : nullptr;
    P.country = geo_cols ? b->country : nullptr;
    P.flags = b->flags;
    bool geo_on_device = H.lpm.geo_loaded && H.needs_geo_cols && !geo_cols;
    P.need_lpm = (H.needs_ip || geo_on_device) ? 1u : 0u;
    if (P.need_lpm && (!b->ip || !b->ip_is_v6)) { e = "batch is missing client.ip columns"; return 1; }
    if (H.needs_port && !b->remote_port) { e = "batch is missing client.remote_port"; return 1; }
    P.verdict = verdict_out;
    P.service = service_out;
    P.n = b->n;
    if (b->n == 0) return 0;
    cudaStream_t cs = (cudaStream_t)stream;
    Scratch* sc = scratch_acquire(rs, b->n, cs, e);
    if (!sc) return 1;
    P.rows = sc->rows;
    P.info = sc->info;
    P.counters = sc->small;
    P.multi_count = sc->small + kSmallCounters + 5;
    P.multi_list = sc->multi;
    GateParams G = rs->gate_base;
    G.n = b->n;
    G.rows = sc->rows;
    G.info = sc->info;
    G.atom_words = H.atom_words;
    for (uint32_t i = 0; i < G.n_fields; ++i) {
        const int f = rs->gate_field[i];
        G.f[i].col = cols[f]->bytes;
        G.f[i].off = cols[f]->offsets;
        G.f[i].bitmap = sc->bitmap[f];
        G.f[i].maybe_count = sc->small + kSmallCounters + 8 + f;
        G.f[i].maybe_idx = sc->cand[f][4];
        G.f[i].cand_count = sc->small + kSmallCounters + f;
        G.f[i].cand_idx = sc->cand[f][0];
        G.f[i].cand_start = sc->cand[f][1];
        G.f[i].cand_end = sc->cand[f][2];
        G.f[i].cand_mask = sc->cand[f][3];
        P.cand_mask[f] = G.f[i].cand_mask;
        P.cand_count[f] = G.f[i].cand_count;
        P.cand_idx[f] = G.f[i].cand_idx;
        P.cand_start[f] = G.f[i].cand_start;
        P.cand_end[f] = G.f[i].cand_end;
    }
    cudaEvent_t* ev = nullptr;
    if (rs->profiling) {
        const uint64_t k = rs->prof_n.fetch_add(1, std::memory_order_relaxed) % kProfRing;
        ev = &rs->prof_ev[4 * k];
    }
    uint32_t nl = 0;
    const char* m = waf_batch_launch(P, G, rs->units.data(), sc->small, kSmallWords, rs->sm_count, rs->scan_smem, rs->gate_smem, stream, ev, &nl);
    scratch_release(rs, sc, cs);
    rs->launches.fetch_add(nl, std::memory_order_relaxed);
    if (m) { e = std::string("CUDA launch failed: ") + m; return 1; }
    return 0;
}

int pgw_evaluate_batch(const pgw_ruleset* rs, const pgw_batch* batch, uint32_t* verdict_out, void* stream) {
    return pgw_evaluate_batch_routed(rs, batch, verdict_out, nullptr, stream);
}

int pgw_evaluate_batch_routed(const pgw_ruleset* rs, const pgw_batch* batch, uint32_t* verdict_out, uint16_t* service_out, void* stream) {
    if (!rs || !batch) return fail("null argument", nullptr, 0);
    if (!rs->finalized) return fail("ruleset is not finalized", nullptr, 0);
    if (batch->n && !verdict_out) return fail("verdict_out is null", nullptr, 0);
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != rs->device) cudaSetDevice(rs->device);
    std::string e;
    int rc = launch_on(const_cast<pgw_ruleset*>(rs), batch, verdict_out, stream, e, service_out);
    if (cur != rs->device && cur >= 0) cudaSetDevice(cur);
    if (rc) return fail(e, nullptr, 0);
    return 0;
}

int pgw_evaluate_batch_host(pgw_ruleset* rs, const pgw_batch* b, uint32_t* verdict_out) {
    return pgw_evaluate_batch_routed_host(rs, b, verdict_out, nullptr);
}

int pgw_evaluate_batch_routed_host(pgw_ruleset* rs, const pgw_batch* b, uint32_t* verdict_out, uint16_t* service_out) {
    if (!rs || !b) return fail("null argument", nullptr, 0);
    if (!rs->finalized) return fail("ruleset is not finalized", nullptr, 0);
    if (b->n == 0) return 0;
    if (!verdict_out) return fail("verdict_out is null", nullptr, 0);
    cudaSetDevice(rs->device);
    const HostProgram& H = rs->prog;
    const uint32_t n = b->n;
    // staging (device buffers, two streams, slice events) comes from a pool: concurrent callers each get their own
    HostStage* hs = nullptr;
    {
        std::lock_guard<std::mutex> lk(rs->pool_mu);
        for (HostStage* h : rs->stages)
            if (!h->busy) { hs = h; break; }
        if (!hs) {
            hs = new HostStage();
            if (!hs->init()) { hs->release(); delete hs; return fail("CUDA: stream creation failed", nullptr, 0); }
            rs->stages.push_back(hs);
        }
        hs->busy = true;
    }
    struct Unbusy {
        pgw_ruleset* rs; HostStage* hs;
        ~Unbusy() { std::lock_guard<std::mutex> lk(rs->pool_mu); hs->busy = false; }
    } unbusy{rs, hs};
    cudaStream_t s = hs->stream, cs = hs->copy_stream;
    const pgw_strcol* hc[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    const bool geo_on_device = H.lpm.geo_loaded && H.needs_geo_cols && !(b->asn && b->country);
    const bool need_ip = H.needs_ip || geo_on_device;
    const bool geo_cols = H.needs_geo_cols && b->asn && b->country;

    // device staging for the whole batch (grown on demand, reused by later calls)
    pgw_batch d;
    memset(&d, 0, sizeof d);
    pgw_strcol* dc[5] = {&d.host, &d.url, &d.path, &d.method, &d.user_agent};
    bool ok = true;
    for (int f = 0; f < 5 && ok; ++f) {
        if (H.field_slot[f] < 0) continue;
        if (!hc[f]->offsets) return fail(std::string("batch is missing offsets for http_request.") + kFieldNames[f], nullptr, 0);
        ok = hs->offs[f].ensure((size_t)(n + 1) * 4);
        dc[f]->offsets = (const uint32_t*)hs->offs[f].d;
        if (ok && ((H.scanned_fields_mask >> f) & 1)) {
            if (!hc[f]->bytes) return fail(std::string("batch is missing bytes for http_request.") + kFieldNames[f], nullptr, 0);
            ok = hs->cols[f].ensure((size_t)hc[f]->offsets[n] + 64);
            dc[f]->bytes = (const uint8_t*)hs->cols[f].d;
        }
    }
    if (need_ip) {
        if (!b->ip || !b->ip_is_v6) return fail("batch is missing client.ip columns", nullptr, 0);
        ok = ok && hs->ip.ensure((size_t)n * 16) && hs->v6.ensure(n);
        d.ip = (const uint8_t*)hs->ip.d;
        d.ip_is_v6 = (const uint8_t*)hs->v6.d;
    }
    if (H.needs_port) {
        if (!b->remote_port) return fail("batch is missing client.remote_port", nullptr, 0);
        ok = ok && hs->port.ensure((size_t)n * 4);
        d.remote_port = (const int32_t*)hs->port.d;
    }
    if (geo_cols) {
        ok = ok && hs->asn.ensure((size_t)n * 8) && hs->country.ensure((size_t)n * 2);
        d.asn = (const int64_t*)hs->asn.d;
        d.country = (const uint16_t*)hs->country.d;
    }
    if (b->flags) {
        ok = ok && hs->flags.ensure(n);
        d.flags = (const uint8_t*)hs->flags.d;
    }
    ok = ok && hs->verdict.ensure((size_t)n * 4) && (!service_out || hs->service.ensure((size_t)n * 2));
    if (!ok) return fail("CUDA: staging allocation failed", nullptr, 0);

    // The batch is cut into slices of whole requests: slice k is copied on the copy stream while slice k-1 is evaluated
    // on the compute stream, so that only the first copy and the last kernel are exposed.  Offsets stay absolute (the
    // column base does not move), a slice is just a window of the offset arrays.
    const uint32_t slice = 262144;
    uint32_t n_slices = (n + slice / 2) / slice;
    if (n_slices < 1) n_slices = 1;
    if (n_slices > kHostSlices) n_slices = kHostSlices;
    const uint32_t per = ((n + n_slices - 1) / n_slices + 31u) & ~31u;
    uint64_t h2d = 0;
    cudaError_t ce = cudaSuccess;
    auto up = [&](const void* dst, const void* src, size_t off, size_t bytes) {
        if (!bytes || ce != cudaSuccess) return;
        ce = cudaMemcpyAsync((uint8_t*)dst + off, (const uint8_t*)src + off, bytes, cudaMemcpyHostToDevice, cs);
        h2d += bytes;
    };
    std::string e;
    uint32_t k = 0;
    // the small fixed-width columns go first, whole (a dozen tiny copies per slice would cost more in call overhead than
    // they move); only the byte columns -- 90 % of the traffic -- are sliced
    for (int f = 0; f < 5; ++f)
        if (H.field_slot[f] >= 0) up(dc[f]->offsets, hc[f]->offsets, 0, ((size_t)n + 1) * 4);
    if (need_ip) { up(d.ip, b->ip, 0, (size_t)n * 16); up(d.ip_is_v6, b->ip_is_v6, 0, n); }
    if (H.needs_port) up(d.remote_port, b->remote_port, 0, (size_t)n * 4);
    if (geo_cols) { up(d.asn, b->asn, 0, (size_t)n * 8); up(d.country, b->country, 0, (size_t)n * 2); }
    if (b->flags) up(d.flags, b->flags, 0, n);
    // the kernels read whole 16-byte chunks up to round_up(total, 32): the bytes past the column's end are never
    // interpreted as request data, but they are read -- give them a defined value
    for (int f = 0; f < 5; ++f)
        if (H.field_slot[f] >= 0 && ((H.scanned_fields_mask >> f) & 1) && ce == cudaSuccess)
            ce = cudaMemsetAsync((uint8_t*)dc[f]->bytes + hc[f]->offsets[n], 0, 64, cs);
    for (uint32_t a = 0; a < n; a += per, ++k) {
        const uint32_t z = a + per < n ? a + per : n, m = z - a;
        for (int f = 0; f < 5; ++f)
            if (H.field_slot[f] >= 0 && ((H.scanned_fields_mask >> f) & 1))
                up(dc[f]->bytes, hc[f]->bytes, hc[f]->offsets[a], (size_t)hc[f]->offsets[z] - hc[f]->offsets[a]);
        if (ce != cudaSuccess) break;
        if ((ce = cudaEventRecord(hs->slice_ready[k], cs)) != cudaSuccess) break;
        if ((ce = cudaStreamWaitEvent(s, hs->slice_ready[k], 0)) != cudaSuccess) break;
        pgw_batch v = d;
        v.n = m;
        pgw_strcol* vc[5] = {&v.host, &v.url, &v.path, &v.method, &v.user_agent};
        for (int f = 0; f < 5; ++f)
            if (vc[f]->offsets) vc[f]->offsets += a;
        if (v.ip) { v.ip += (size_t)a * 16; v.ip_is_v6 += a; }
        if (v.remote_port) v.remote_port += a;
        if (v.asn) { v.asn += a; v.country += a; }
        if (v.flags) v.flags += a;
        uint32_t* vd = (uint32_t*)hs->verdict.d + a;
        uint16_t* sd = service_out ? (uint16_t*)hs->service.d + a : nullptr;
        if (launch_on(rs, &v, vd, s, e, sd)) { cudaStreamSynchronize(cs); cudaStreamSynchronize(s); return fail(e, nullptr, 0); }
        if ((ce = cudaMemcpyAsync(verdict_out + a, vd, (size_t)m * 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        if (sd && (ce = cudaMemcpyAsync(service_out + a, sd, (size_t)m * 2, cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
    }
    cudaError_t c1 = cudaStreamSynchronize(cs), c2 = cudaStreamSynchronize(s);
    if (ce == cudaSuccess) ce = c1 != cudaSuccess ? c1 : c2;
    if (ce != cudaSuccess) return fail(std::string("CUDA: ") + cudaGetErrorString(ce), nullptr, 0);
    rs->last_h2d.store(h2d);
    rs->last_d2h.store((uint64_t)n * (service_out ? 6 : 4));
    return 0;
}

int pgw_geoip_lookup_batch(const pgw_ruleset* rs, const uint8_t* ip, const uint8_t* ip_is_v6, uint32_t n, uint32_t* asn_out,
                           uint16_t* country_out, void* stream) {
    if (!rs || !rs->finalized) return fail("ruleset is not finalized", nullptr, 0);
    if (n && (!ip || !ip_is_v6 || !asn_out || !country_out)) return fail("null argument", nullptr, 0);
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != rs->device) cudaSetDevice(rs->device);
    KParams P = rs->base;
    const char* m = geoip_launch(P, ip, ip_is_v6, n, asn_out, country_out, stream);
    if (cur != rs->device && cur >= 0) cudaSetDevice(cur);
    if (m) return fail(std::string("CUDA launch failed: ") + m, nullptr, 0);
    if (n) const_cast<pgw_ruleset*>(rs)->launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int pgw_captcha_client_id_batch(const pgw_batch* batch, uint8_t* out44_dev, void* stream) {
    if (!batch) return fail("null argument", nullptr, 0);
    if (batch->n == 0) return 0;
    if (!out44_dev || !batch->ip || !batch->ip_is_v6 || !batch->user_agent.offsets || !batch->host.offsets)
        return fail("client id needs ip, ip_is_v6, user_agent and host columns", nullptr, 0);
    if (const char* m = client_id_launch(batch->ip, batch->ip_is_v6, batch->user_agent.bytes, batch->user_agent.offsets, batch->host.bytes,
                                        batch->host.offsets, batch->n, out44_dev, stream))
        return fail(std::string("CUDA launch failed: ") + m, nullptr, 0);
    return 0;
}

void* pgw_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void pgw_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int pgw_ruleset_set_profiling(pgw_ruleset* rs, int enable) {
    if (!rs || !rs->finalized) return fail("ruleset is not finalized", nullptr, 0);
    cudaSetDevice(rs->device);
    if (enable && rs->prof_ev.empty()) {
        rs->prof_ev.resize(4 * kProfRing, nullptr);
        for (auto& ev : rs->prof_ev)
            if (cudaEventCreate(&ev) != cudaSuccess) return fail("CUDA: event creation failed", nullptr, 0);
    }
    rs->prof_n.store(0);
    rs->profiling = enable != 0;
    return 0;
}

int pgw_ruleset_profile_kernels(pgw_ruleset* rs, double ms_sum[3], uint32_t* batches) {
    if (!rs || !ms_sum || !batches) return fail("null argument", nullptr, 0);
    ms_sum[0] = ms_sum[1] = ms_sum[2] = 0.0;
    *batches = 0;
    if (rs->prof_ev.empty()) return 0;
    cudaSetDevice(rs->device);
    const uint64_t n = rs->prof_n.load();
    const uint32_t have = (uint32_t)(n < kProfRing ? n : kProfRing);
    for (uint32_t k = 0; k < have; ++k)
        for (int j = 0; j < 3; ++j) {
            float ms = 0.f;
            if (cudaEventSynchronize(rs->prof_ev[4 * k + j + 1]) != cudaSuccess ||
                cudaEventElapsedTime(&ms, rs->prof_ev[4 * k + j], rs->prof_ev[4 * k + j + 1]) != cudaSuccess)
                return fail("CUDA: profiling events are not complete", nullptr, 0);
            ms_sum[j] += ms;
        }
    *batches = have;
    rs->prof_n.store(0);
    return 0;
}

int pgw_ruleset_profile(pgw_ruleset* rs, double* scan_ms_sum, uint32_t* launches) {
    if (!scan_ms_sum) return fail("null argument", nullptr, 0);
    double ms[3];
    if (int rc = pgw_ruleset_profile_kernels(rs, ms, launches)) return rc;
    *scan_ms_sum = ms[0] + ms[1];
    return 0;
}

int pgw_ruleset_info(const pgw_ruleset* rs, pgw_info* out) {
    if (!rs || !out) return fail("null argument", nullptr, 0);
    memset(out, 0, sizeof *out);
    const HostProgram& H = rs->prog;
    out->n_rules = (uint32_t)rs->builder.n_rules();
    if (!rs->finalized) return 0;
    out->n_atoms = H.n_atoms;
    out->n_scan_units = (uint32_t)H.units.size();
    out->n_nonscan_atoms = (uint32_t)H.ns_atoms.size();
    out->scanned_fields_mask = H.scanned_fields_mask;
    for (int f = 0; f < 5; ++f)
        if (H.field_slot[f] >= 0) out->offset_fields_mask |= 1u << f;
    out->reads_ip = H.needs_ip || (H.lpm.geo_loaded && H.needs_geo_cols);
    out->reads_port = H.needs_port;
    out->reads_geo_columns = H.needs_geo_cols;
    out->table_arena_bytes = H.arena.size();
    out->smem_bytes = rs->scan_smem;
    for (auto& u : H.units) out->total_dfa_states += u.n_states;
    out->tables_in_smem = rs->hot_states_total == out->total_dfa_states;
    out->hot_dfa_states = rs->hot_states_total;  /* states whose rows live in shared memory */
    out->grid = (uint32_t)rs->sm_count;
    out->threads = (uint32_t)waf_scan_threads();
    for (int f = 0; f < 5; ++f)
        if (H.gate[f].present) { out->gated_fields_mask |= 1u << f; out->gate_grams += H.gate[f].n_grams; }
    out->gate_smem_bytes = rs->gate_smem;
    out->n_bitset_units = (uint32_t)H.bitset_units.size();
    for (auto& b : H.bitset_units) out->bitset_positions += b.n_pos;
    out->lpm_present = H.lpm.present;
    out->geoip_loaded = H.lpm.geo_loaded;
    out->kernel_launches = rs->launches.load();
    out->last_h2d_bytes = rs->last_h2d.load();
    out->last_d2h_bytes = rs->last_d2h.load();
    return 0;
}

size_t pgw_ruleset_describe(const pgw_ruleset* rs, char* buf, size_t cap) {
    if (!rs) return 0;
    std::string s = rs->finalized ? rs->prog.summary() : std::string("(not finalized)");
    for (auto& w : rs->prog.warnings) s += "\nwarning: " + w;
    if (buf && cap) {
        size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

void pgw_ruleset_destroy(pgw_ruleset* rs) {
    if (!rs) return;
    if (rs->device >= 0) cudaSetDevice(rs->device);
    rs->mem.release();
    for (HostStage* h : rs->stages) { h->release(); delete h; }
    for (Scratch* sc : rs->pool) {
        if (sc->done) { cudaEventSynchronize(sc->done); cudaEventDestroy(sc->done); }
        if (sc->base) cudaFree(sc->base);
        delete sc;
    }
    for (auto& ev : rs->prof_ev)
        if (ev) cudaEventDestroy(ev);
    delete rs;
}

}  // extern "C"
