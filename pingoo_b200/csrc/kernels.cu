// sm_100a kernels for the batched WAF verdict path.
//
// Per batch they replace the reference's per-request
//   ctx build (http_listener.rs:239-249) -> rule loop (http_listener.rs:251-264)
//   -> bel::Program::execute (pingoo/rules.rs:36-52) -> regex / list / geoip work -> service routing (:266-272).
//
// File map (one translation unit: the .cuh fragments are included below, device code is not linked across files)
//   kernel_common.cuh   shared-window loads/stores (explicit ld.shared: no generic-address conversion per access),
//                       mbarrier + TMA bulk copy, rule bytecode, longest-prefix lookup,
//                       request_epilogue: per-request predicates outside the byte scan, gates, verdict and service
//   kernel_gate.cuh     waf_gate_kernel -- candidate gate: flat coalesced pass over the url / user_agent / path columns,
//                       4-byte windows against two hashed bitmaps in shared memory, candidate lists per field
//   kernel_field.cuh    waf_field_scan_kernel -- DFA scan: unit-major, lane-owned strings, pooled claims, per-unit
//                       shared-memory images, early exit on absorbing states, bitmaps + dirty bits in global memory;
//                       waf_epilogue_kernel -- verdicts (one thread per request, warp-shared evaluation)
//   kernel_bitset.cuh   waf_bitset_nfa_kernel -- bit-parallel NFA simulation (active-position bit vectors, tables in shared
//                       memory) for the patterns whose DFA would exceed the scan-unit caps
//   kernel_misc.cuh     geoip_lookup_kernel (geoip.rs:73-91), captcha_client_id_kernel (captcha.rs:409-421)
//   kernels.cu          this file: launch wrappers, host-callable, no CUDA types in their signatures beyond the stream
//
// One batch = memset of the claim / candidate counters -> gate -> scan -> epilogue on the caller's stream.
//   DFA tables: each unit's class map and the rows of its shallow ("hot") states are staged into shared memory by TMA
//     bulk copies (cp.async.bulk + mbarrier) when the CTA starts on the unit; transitions into deeper states lead to a
//     trap row and the word is re-walked on the full table in global memory (L1/L2);
//   request bytes: 128-bit ld.global.nc loads issued ahead of their use;
//   events (rare): accepting states carry event lists -- FIRE atom / SET, TEST, CLEAR of a per-scan latch register
//     (gap-split patterns such as `<tag[^>]*>`, see regex.hpp);
//   verdict = first matching rule with a terminal action; a request none of whose atoms fired takes a precomputed
//     verdict without its bitmap row being read, one deviating atom a tabulated one, otherwise only rules that mention
//     a deviating atom run.
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "gate.hpp"
#include "kernels.cuh"

#ifndef PGW_LD_MODE
#define PGW_LD_MODE 0
#endif

namespace pgw {

namespace {
#include "kernel_common.cuh"
#include "kernel_gate.cuh"
#include "kernel_field.cuh"
#include "kernel_bitset.cuh"
#include "kernel_misc.cuh"

}  // namespace

const char* client_id_launch(const uint8_t* ip, const uint8_t* is_v6, const uint8_t* ua_bytes, const uint32_t* ua_off, const uint8_t* host_bytes,
                             const uint32_t* host_off, uint32_t n, uint8_t* out44, void* stream) {
    if (n == 0) return nullptr;
    captcha_client_id_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ip, is_v6, ua_bytes, ua_off, host_bytes, host_off, n, out44);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* geoip_launch(const KParams& p, const uint8_t* ip, const uint8_t* is_v6, uint32_t n, uint32_t* asn_out,
                         uint16_t* country_out, void* stream) {
    if (n == 0) return nullptr;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    geoip_lookup_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, ip, is_v6, n, asn_out, country_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* waf_configure(int device, size_t* max_smem_optin, int* sm_count) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    int v = 0;
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    *max_smem_optin = (size_t)v;
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    *sm_count = v;
    e = cudaFuncSetAttribute(waf_field_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*max_smem_optin);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    e = cudaFuncSetAttribute(waf_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*max_smem_optin);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    e = cudaFuncSetAttribute(waf_epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(waf_prefix_budget() + 512));
    if (e != cudaSuccess) return cudaGetErrorString(e);
    e = cudaFuncSetAttribute(waf_bitset_nfa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)waf_bitset_smem_budget());
    if (e != cudaSuccess) return cudaGetErrorString(e);
    return nullptr;
}

int waf_scan_threads() { return kFsThreads; }

size_t waf_scan_image_budget(size_t max_smem_optin) { return max_smem_optin > kFsFront + 256 ? max_smem_optin - kFsFront - 256 : 0; }

size_t waf_scan_smem_bytes(uint32_t max_image_bytes) { return 256 + kFsFront + r16(max_image_bytes); }

size_t waf_gate_smem_bytes(const GateParams& g) {
    size_t m = 0;
    for (uint32_t i = 0; i < g.n_fields; ++i) {
        const size_t end = g.f[i].bloom_off + (((size_t)1 << g.f[i].k1) / 8);
        if (end > m) m = end;
    }
    return kGateCtrBytes + m + 128;
}

size_t waf_prefix_budget() { return 64u << 10; }

size_t waf_bitset_smem_budget() { return 96u << 10; }

const char* waf_batch_launch(KParams& p, GateParams& g, const UnitDesc* all_units, uint32_t* small, uint32_t small_words, int sm_count,
                             size_t scan_smem, size_t gate_smem, void* stream, cudaEvent_t* ev, uint32_t* launches) {
    if (p.n == 0) return nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t nl = 0;
    cudaError_t e = cudaMemsetAsync(small, 0, (size_t)small_words * 4, s);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    if (ev) cudaEventRecord(ev[0], s);
    if (g.n_fields) {
        // flat stream over the gated columns: one CTA per SM (fewer for small batches: each stages the level-1 bitmaps)
        uint32_t gg = (p.n + 127u) / 128u;
        if (gg > (uint32_t)sm_count) gg = (uint32_t)sm_count;
        waf_gate_kernel<<<(int)gg, kGateThreads, gate_smem, s>>>(g);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
        uint32_t lb = (p.n + kListThreads - 1u) / kListThreads;
        if (lb > (uint32_t)sm_count * 2u) lb = (uint32_t)sm_count * 2u;
        waf_gate_maybe_kernel<<<lb, kListThreads, 0, s>>>(g);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
        if (g.wide_slots) {
            uint32_t rb = (p.n + kResolveThreads - 1u) / kResolveThreads;
            if (rb > (uint32_t)sm_count * 8u) rb = (uint32_t)sm_count * 8u;
            waf_gate_resolve_lit_kernel<<<dim3(rb, g.n_fields), kResolveThreads, 0, s>>>(g);
        } else waf_gate_resolve_kernel<<<dim3(lb, g.n_fields), kListThreads, 0, s>>>(g);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
        nl += 3;
    }
    if (ev) cudaEventRecord(ev[1], s);
    for (uint32_t ub = 0; ub < p.n_units_total; ub += kMaxConstUnits) {
        p.unit_base = ub;
        p.n_units = p.n_units_total - ub < kMaxConstUnits ? p.n_units_total - ub : kMaxConstUnits;
        memcpy(p.udesc, all_units + ub, p.n_units * sizeof(UnitDesc));
        bool any = false;
        for (uint32_t u = 0; u < p.n_units; ++u) any |= p.udesc[u].mode != UM_PREPASS;
        if (!any) continue;   // every unit of this group is walked by the epilogue kernel
        const uint32_t want = (p.n + kFsThreads - 1) / kFsThreads;
        const int grid = (int)(want < (uint32_t)sm_count ? want : (uint32_t)sm_count);
        waf_field_scan_kernel<<<grid, kFsThreads, scan_smem, s>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
        ++nl;
    }
    if (p.n_bitset) {
        BitsetParams bp;
        for (int f = 0; f < 5; ++f) { bp.col[f] = p.col[f]; bp.off[f] = p.off[f]; }
        bp.n = p.n;
        bp.rows = p.rows;
        bp.info = p.info;
        bp.atom_words = p.atom_words;
        bp.units = p.bitset_units;
        bp.blob = p.bitset_blob;
        bp.smem_words = p.bitset_smem_words;
        uint32_t blocks = (p.n + kBitsetThreads - 1) / kBitsetThreads;
        if (blocks > (uint32_t)sm_count * 8u) blocks = (uint32_t)sm_count * 8u;
        waf_bitset_nfa_kernel<<<dim3(blocks, p.n_bitset), kBitsetThreads, (size_t)p.bitset_smem_words * 4, s>>>(bp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
        ++nl;
    }
    if (ev) cudaEventRecord(ev[2], s);
    {
        uint32_t blocks = (p.n + kEpiThreads - 1) / kEpiThreads;
        if (blocks > (uint32_t)sm_count * 4u) blocks = (uint32_t)sm_count * 4u;
        waf_epilogue_kernel<<<(int)blocks, kEpiThreads, p.prefix_area + 256, s>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
        ++nl;
    }
    {
        // at most a few per cent of the requests; sized for the machine, the kernel reads the count from device memory
        const int mb = sm_count * 8;
        waf_multi_kernel<<<mb, 256, 0, s>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
        ++nl;
    }
    if (ev) cudaEventRecord(ev[3], s);
    if (launches) *launches = nl;
    return nullptr;
}

}  // namespace pgw
