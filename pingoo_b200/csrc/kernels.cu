// sm_100a kernels for the batched WAF verdict path.
//
// Per batch they replace the reference's per-request
//   ctx build (http_listener.rs:239-249) -> rule loop (http_listener.rs:251-264)
//   -> bel::Program::execute (pingoo/rules.rs:36-52) -> regex / list / geoip work -> service routing (:266-272).
//
// File map (one translation unit: the .cuh fragments are included below, device code is not linked across files)
//   kernel_common.cuh   shared-window loads/stores (explicit ld.shared: no generic-address conversion per access),
//                       mbarrier + TMA bulk copy, event lists, rule bytecode, longest-prefix lookup,
//                       request_epilogue_t: per-request predicates outside the byte scan, gates, verdict and service
//   kernel_field.cuh    waf_field_scan_kernel -- DEFAULT path: unit-major, lane-owned strings, pooled claims, bitmaps in
//                       global memory; waf_epilogue_kernel -- verdicts (one thread per request, warp-shared evaluation)
//   kernel_lane.cuh     waf_verdict_kernel -- "lane" path (PGW_KERNEL=lane; also the fall-back beyond kMaxConstUnits
//                       scan units): request-major persistent kernel, bitmaps in shared memory, epilogue fused
//   kernel_stream.cuh   waf_stream_scan_kernel -- "stream" path (PGW_KERNEL=stream): coalesced segments, speculated states
//   kernel_misc.cuh     geoip_lookup_kernel (geoip.rs:73-91), captcha_client_id_kernel (captcha.rs:409-421)
//   kernels.cu          this file: launch wrappers, host-callable, no CUDA types in their signatures beyond the stream
//
// Common to all scan paths
//   DFA tables: class maps + the rows of the shallow ("hot") states of every DFA are staged once per CTA into shared
//     memory by TMA bulk copies (cp.async.bulk + mbarrier); transitions into deeper states lead to a trap row and the
//     word is re-walked on the full table in global memory (L1/L2);
//   request bytes: 128-bit ld.global.nc loads issued one iteration ahead of their use;
//   events (rare): accepting states carry event lists -- FIRE atom / SET, TEST, CLEAR of a per-scan latch register
//     (gap-split patterns such as `<tag[^>]*>`, see regex.hpp);
//   verdict = first matching rule with a terminal action; a request whose atom vector equals the expected vector takes a
//     precomputed verdict, one deviating atom a tabulated one, otherwise only rules that mention a deviating atom run.
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>

#include "kernels.cuh"

#ifndef PGW_LD_MODE
#define PGW_LD_MODE 0
#endif
#ifndef PGW_L2_PREFETCH
#define PGW_L2_PREFETCH 0
#endif

namespace pgw {

namespace {
#include "kernel_common.cuh"
#include "kernel_lane.cuh"
#include "kernel_stream.cuh"
#include "kernel_field.cuh"
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

size_t waf_smem_bytes(uint32_t image_bytes, uint32_t n_units, uint32_t atom_words, uint32_t n_slots) { return smem_layout(image_bytes, n_units, atom_words, n_slots).total; }
size_t waf_smem_fixed_bytes(uint32_t n_units, uint32_t atom_words, uint32_t n_slots) { return smem_layout(0, n_units, atom_words, n_slots).total; }

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
    e = cudaFuncSetAttribute(waf_verdict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*max_smem_optin);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    e = cudaFuncSetAttribute(waf_stream_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*max_smem_optin);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    e = cudaFuncSetAttribute(waf_field_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*max_smem_optin);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    return nullptr;
}

int waf_field_threads() { return kFsThreads; }

size_t waf_field_smem_bytes(uint32_t image_bytes, uint32_t n_units) {
    return 256 + r16(image_bytes) + r16(n_units * (uint32_t)sizeof(UnitDesc)) + 64 + (kFsThreads / 32) * 2 * kFsPoolBytes + 4 * kFsSlotStride;
}

const char* waf_field_launch(const KParams& p, uint32_t* rows, uint32_t* counters, int sm_count, size_t smem_bytes, void* stream, cudaEvent_t ev0,
                             cudaEvent_t ev1) {
    if (p.n == 0) return nullptr;
    if (p.n_units > kFieldCounters || p.n_units > kMaxConstUnits) return "too many scan units for the field-scan path";
    cudaStream_t s = (cudaStream_t)stream;
    // bitmaps and the per-unit claim counters are one allocation: one memset
    cudaError_t e = cudaMemsetAsync(rows, 0, ((size_t)p.n * p.atom_words + kFieldCounters) * 4, s);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    if (p.n_units) {
        const uint32_t want = (p.n + kFsThreads - 1) / kFsThreads;
        const int grid = (int)(want < (uint32_t)sm_count ? want : (uint32_t)sm_count);
        if (ev0) cudaEventRecord(ev0, s);
        waf_field_scan_kernel<<<grid, kFsThreads, smem_bytes, s>>>(p, rows, counters);
        if (ev1) cudaEventRecord(ev1, s);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
    }
    int blocks = (int)((p.n + 255) / 256);
    if (blocks > sm_count * 8) blocks = sm_count * 8;
    waf_epilogue_kernel<<<blocks, 256, 0, s>>>(p, rows);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

size_t waf_stream_smem_bytes(uint32_t image_bytes, uint32_t n_units) {
    return r16(image_bytes) + r16(n_units * (uint32_t)sizeof(UnitDesc)) + (kStreamThreads / 32) * (kStreamNB + 4) * 4 + 64;
}

const char* waf_stream_launch(const KParams& p, uint32_t* rows, uint32_t* task_counter, int sm_count, size_t smem_bytes, void* stream) {
    if (p.n == 0) return nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(task_counter, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    e = cudaMemsetAsync(rows, 0, (size_t)p.n * p.atom_words * 4, s);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    const uint32_t n_blocks = (p.n + kStreamNB - 1) / kStreamNB;
    const uint32_t n_tasks = n_blocks * p.n_units;
    if (n_tasks) {
        int ctas_per_sm = smem_bytes * 2 + 2048 <= 227 * 1024 ? 2 : 1;
        waf_stream_scan_kernel<<<sm_count * ctas_per_sm, kStreamThreads, smem_bytes, s>>>(p, rows, task_counter, n_blocks, n_tasks);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cudaGetErrorString(e);
    }
    int blocks = (int)((p.n + 255) / 256);
    if (blocks > sm_count * 8) blocks = sm_count * 8;
    waf_epilogue_kernel<<<blocks, 256, 0, s>>>(p, rows);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* waf_launch(const KParams& p, const LaunchPlan& plan, void* stream) {
    if (p.n == 0) return nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    waf_verdict_kernel<<<plan.grid, kThreads, plan.smem_bytes, s>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace pgw
