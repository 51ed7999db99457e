// sm_100a kernels for the batched WAF verdict path.
//
// Per batch they replace the reference's per-request
//   ctx build (http_listener.rs:239-249) -> rule loop (http_listener.rs:251-264)
//   -> bel::Program::execute (pingoo/rules.rs:36-52) -> regex / list / geoip work -> service routing (:266-272).
//
// File map
//   helpers                    shared-window loads/stores (explicit ld.shared: no generic-address conversion per access),
//                              mbarrier + TMA bulk copy, event lists, rule bytecode, longest-prefix lookup
//   request_epilogue_t         per-request predicates outside the byte scan, gates, verdict and service
//   waf_verdict_kernel         "lane" path (PGW_KERNEL=lane; also the fall-back beyond kMaxConstUnits scan units):
//                              request-major persistent kernel, bitmaps in shared memory, epilogue fused
//   waf_stream_scan_kernel     "stream" path (PGW_KERNEL=stream): coalesced 16-byte segments, speculated start states
//   waf_field_scan_kernel      DEFAULT path: unit-major, lane-owned strings, pooled claims, bitmaps in global memory
//   waf_epilogue_kernel        verdicts for the field / stream paths (one thread per request, warp-shared evaluation)
//   geoip_lookup_kernel        GeoipDB::lookup for a batch (geoip.rs:73-91)
//   captcha_client_id_kernel   generate_captcha_client_id for a batch (captcha.rs:409-421)
//   launch wrappers            host-callable, no CUDA types in their signatures beyond the stream handle
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint4 ld_nc_v4(const uint8_t* p) {
    uint4 r;
#if PGW_LD_MODE == 1
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#elif PGW_LD_MODE == 2
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#endif
    return r;
}

struct SmemLayout {
    uint32_t image, units, rows, ext, misc, total;
};

__host__ __device__ inline uint32_t r16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline SmemLayout smem_layout(uint32_t image_bytes, uint32_t n_units, uint32_t atom_words, uint32_t n_slots) {
    SmemLayout L;
    uint32_t o = 0;
    L.image = o;
    o += r16(image_bytes);
    L.units = o;
    o += r16(n_units * (uint32_t)sizeof(UnitDesc));
    L.rows = o;
    o += r16((uint32_t)kThreads * kRowsPerLane * atom_words * 4u);
    L.ext = o;  // per lane: (start, end) offsets of every scanned field of its next request
    o += r16((uint32_t)kThreads * 2u * n_slots * 4u);
    L.misc = o;
    o += 64;
    L.total = o;
    return L;
}

// Apply the events of CSR row `row` (sorted FIRE, TEST, CLEAR, SET) to the lane's bitmap and latch register.
// Returns true if every event was a plain FIRE (idempotent: the caller may skip an immediate repeat).
__device__ __noinline__ bool run_events(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ events, uint32_t row,
                                        uint32_t* bits, uint32_t stride, uint32_t* latch) {
    uint32_t a = __ldg(idx + row), b = __ldg(idx + row + 1);
    bool pure = true;
    uint32_t l = *latch;
    for (uint32_t i = a; i < b; ++i) {
        const uint32_t e = __ldg(events + i);
        const uint32_t kind = e >> kEvKindShift, lb = 1u << ((e >> kEvLatchShift) & 31u), at = e & kEvAtomMask;
        if (kind == 0u || (kind == 1u && (l & lb))) bits[(at >> 5) * stride] |= 1u << (at & 31);
        else if (kind == 2u) l &= ~lb;
        else if (kind == 3u) l |= lb;
        pure &= kind == 0u;
    }
    *latch = l;
    return pure;
}

__device__ __forceinline__ bool eval_rule(const uint16_t* __restrict__ code, uint32_t a, uint32_t b, const uint32_t* row, uint32_t stride) {
    uint32_t st = 0;
    for (uint32_t i = a; i < b; ++i) {
        uint32_t op = __ldg(code + i);
        if (op < 0x4000u) st = (st << 1) | ((row[(op >> 5) * stride] >> (op & 31)) & 1u);
        else if (op == OP_NOT) st ^= 1u;
        else if (op == OP_AND) st = ((st >> 1) & ~1u) | (st & (st >> 1) & 1u);
        else if (op == OP_OR) st = ((st >> 1) & ~1u) | ((st | (st >> 1)) & 1u);
        else if (op == OP_PUSH0) st <<= 1;
        else st = (st << 1) | 1u;
    }
    return st & 1u;
}

__device__ __forceinline__ uint32_t lpm_lookup(const KParams& p, const uint8_t* ip16, bool v6) {
    if (!v6) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(ip16);
        uint32_t a = __byte_perm(w, 0, 0x0123);  // network order -> host integer
        uint32_t e = __ldg(p.dir24 + (a >> 8));
        if (e & 0x80000000u) e = __ldg(p.tbl8 + ((e & 0x7FFFFFFFu) << 8) + (a & 0xFFu));
        return e;
    }
    const uint32_t* w = reinterpret_cast<const uint32_t*>(ip16);
    uint64_t hi = ((uint64_t)__byte_perm(w[0], 0, 0x0123) << 32) | __byte_perm(w[1], 0, 0x0123);
    uint64_t lo = ((uint64_t)__byte_perm(w[2], 0, 0x0123) << 32) | __byte_perm(w[3], 0, 0x0123);
    // last range whose start <= (hi,lo); range 0 starts at 0
    uint32_t l = 0, r = p.n_v6;
    while (r - l > 1) {
        uint32_t m = (l + r) >> 1;
        uint64_t mh = __ldg(p.v6_hi + m), ml = __ldg(p.v6_lo + m);
        bool le = mh < hi || (mh == hi && ml <= lo);
        if (le) l = m;
        else r = m;
    }
    return __ldg(p.v6_leaf + l);
}

// Per-request predicates outside the byte scan + the verdict (http_listener.rs:196-264).
// `row[w * stride]` is the request's atom bitmap (scan atoms already set).
// WARP: called by all 32 lanes of a converged warp (`valid` false for lanes past the end of the batch, which shadow the
// last request without storing anything): requests of the warp that deviate from the expected atom vector in the same
// way are evaluated once -- verdict and service are functions of the deviation and of `captcha_verified` alone -- and
// the result is shared by shuffle.
template <bool WARP>
__device__ __forceinline__ void request_epilogue_t(const KParams& p, uint32_t r, uint32_t* row, uint32_t stride, bool valid) {
    const uint32_t Aw = p.atom_words;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t flags = p.flags ? p.flags[r] : 0u;
    int64_t asn = 0;
    uint32_t country = (uint32_t)'X' | ((uint32_t)'X' << 8);
    uint32_t set_mask = 0;
    if (p.need_lpm) {
        const uint8_t* ip16 = p.ip + (size_t)r * 16;
        const bool v6 = p.is_v6[r] != 0;
        const LpmLeaf lf = p.leaves[lpm_lookup(p, ip16, v6)];
        set_mask = lf.set_mask;
        if (p.geo_loaded && p.asn == nullptr) {
            // geoip.rs:74-76: loopback / multicast are never looked up
            bool skip;
            if (!v6) skip = ip16[0] == 127 || (ip16[0] >> 4) == 0xE;
            else {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(ip16);
                skip = ip16[0] == 0xFF || (w[0] == 0 && w[1] == 0 && w[2] == 0 && w[3] == 0x01000000u);
            }
            if (!skip) { asn = lf.asn; country = lf.country; }
        }
    }
    if (p.asn) asn = p.asn[r];
    if (p.country) country = p.country[r];

    for (uint32_t i = 0; i < p.n_ns; ++i) {
        const NsAtom a = p.n_ns <= kMaxConstNs ? p.nsd[i] : p.ns[i];
        bool v = false;
        if (a.kind == 1 || a.kind == 2) {  // INT_CMP / INT_SET
            int64_t x;
            if (a.feat == 0) x = p.port ? (int64_t)p.port[r] : 0;
            else if (a.feat == 1) x = asn;
            else {
                const uint32_t* o = p.off[a.feat - 2] + r;
                x = (int64_t)(o[1] - o[0]);
            }
            if (a.kind == 1) {
                switch (a.op) {
                    case 0: v = x == a.cval; break;
                    case 1: v = x != a.cval; break;
                    case 2: v = x < a.cval; break;
                    case 3: v = x <= a.cval; break;
                    case 4: v = x > a.cval; break;
                    default: v = x >= a.cval; break;
                }
            } else {
                uint32_t l = p.iset_off[a.set_id], h = p.iset_off[a.set_id + 1];
                while (l < h) {
                    uint32_t m = (l + h) >> 1;
                    int64_t mv = __ldg(p.iset_vals + m);
                    if (mv == x) { v = true; break; }
                    if (mv < x) l = m + 1;
                    else h = m;
                }
            }
        } else if (a.kind == 3) {  // IP_SET
            v = (set_mask >> a.set_id) & 1u;
        } else {  // COUNTRY_SET
            uint32_t c0 = (country & 0xFFu) - 'A', c1 = ((country >> 8) & 0xFFu) - 'A';
            if (c0 < 26u && c1 < 26u) {
                uint32_t bit = c0 * 26u + c1;
                v = (__ldg(p.cset + a.set_id * kCountryWords + (bit >> 5)) >> (bit & 31)) & 1u;
            }
        }
        if (v && valid) row[(a.atom >> 5) * stride] |= 1u << (a.atom & 31);
    }

    const uint32_t cv = flags & RF_CAPTCHA_VERIFIED;
    uint32_t verdict = V_ALLOW | (kNoRule << 2);
    bool decided = false;
    if (flags & RF_PRE_BLOCK) { verdict = V_BLOCK | (kNoRule << 2); decided = true; }
    if (!decided && p.eval_gates) {
        // http_listener.rs:196-198: empty or over-long user agent is blocked before any rule
        const uint32_t* o = p.off[4] + r;
        uint32_t ual = o[1] - o[0];
        if (ual == 0 || ual >= 256) { verdict = V_BLOCK | (kNoRule << 2); decided = true; }
    }
    if (!decided) {
        bool bypass = flags & RF_BYPASS;
        if (p.eval_gates && p.gate_atom >= 0) bypass |= (row[(p.gate_atom >> 5) * stride] >> (p.gate_atom & 31)) & 1u;
        if (bypass) { verdict = V_BYPASS | (kNoRule << 2); decided = true; }
    }
    if (!decided && (flags & RF_PRE_CAPTCHA)) { verdict = V_CAPTCHA | (kNoRule << 2); decided = true; }

    // deviations from the expected atom vector: none -> v0 / s0, exactly one -> v1[atom] / s1[atom], otherwise the
    // candidate rules (those that mention a deviating atom, plus the ones true by default) are evaluated
    const bool routes = p.service != nullptr && p.n_rules > p.n_waf_rules;
    uint32_t ndev = 0, dev_atom = 0, sig = cv;
    for (uint32_t w = 0; w < Aw; ++w) {
        const uint32_t x = (row[w * stride] ^ __ldg(p.expect + w)) & __ldg(p.care + w);
        if (x) dev_atom = w * 32u + (uint32_t)__ffs(x) - 1u;
        ndev += (uint32_t)__popc(x);
        sig = sig * 0x9E3779B1u + x;
    }
    uint32_t svc = kNoService;
    if (!decided) {
        if (ndev == 0u) { verdict = p.v0[cv]; svc = p.s0; }
        else if (ndev == 1u) { verdict = __ldg(p.v1 + cv * p.n_atoms + dev_atom); svc = routes ? (uint32_t)__ldg(p.s1 + dev_atom) : kNoService; }
    }
    const bool need_eval = !decided && ndev >= 2u;
    if (WARP ? __any_sync(FULL, need_eval) : need_eval) {
        bool do_eval = need_eval, same = false;
        uint32_t leader = 0;
        if (WARP) {
            const uint32_t lane = threadIdx.x & 31u;
            const uint32_t peers = __match_any_sync(FULL, need_eval ? (sig & 0x7FFFFFFFu) : (0x80000000u | lane));
            leader = (uint32_t)__ffs(peers) - 1u;
            same = true;  // equal signature: confirm that the deviation really is the leader's
            for (uint32_t w = 0; w < Aw; ++w) {
                const uint32_t x = (row[w * stride] ^ __ldg(p.expect + w)) & __ldg(p.care + w);
                same &= __shfl_sync(FULL, x, leader) == x;
            }
            same &= __shfl_sync(FULL, cv, leader) == cv;
            do_eval = need_eval && (leader == lane || !same);
        }
        if (do_eval) {
            const uint32_t tshift = 2 * cv;
            uint32_t best = kNoRule, best_svc = kNoRule;
            for (uint32_t w = 0; w < Aw; ++w) {
                uint32_t x = (row[w * stride] ^ __ldg(p.expect + w)) & __ldg(p.care + w);
                while (x) {
                    uint32_t b = __ffs(x) - 1;
                    x &= x - 1;
                    uint32_t atom = w * 32 + b;
                    uint32_t i0 = __ldg(p.ar_idx + atom), i1 = __ldg(p.ar_idx + atom + 1);
                    for (uint32_t i = i0; i < i1; ++i) {
                        uint32_t rule = __ldg(p.ar_rules + i);  // ascending; WAF rules first, then service routes
                        if (rule < p.n_waf_rules) {
                            if (rule >= best || ((__ldg(p.term + rule) >> tshift) & 3u) == 0) continue;
                            if (eval_rule(p.code, __ldg(p.rule_off + rule), __ldg(p.rule_off + rule + 1), row, stride)) best = rule;
                        } else {
                            if (!routes || rule >= best_svc) break;
                            if (eval_rule(p.code, __ldg(p.rule_off + rule), __ldg(p.rule_off + rule + 1), row, stride)) best_svc = rule;
                        }
                    }
                }
            }
            for (uint32_t i = 0; i < p.n_dflt[cv]; ++i) {
                uint32_t rule = __ldg(p.dflt[cv] + i);
                if (rule >= best) break;
                if (eval_rule(p.code, __ldg(p.rule_off + rule), __ldg(p.rule_off + rule + 1), row, stride)) best = rule;
            }
            if (routes)
                for (uint32_t i = 0; i < p.n_dflt_services; ++i) {
                    uint32_t rule = __ldg(p.dflt_services + i);
                    if (rule >= best_svc) break;
                    if (eval_rule(p.code, __ldg(p.rule_off + rule), __ldg(p.rule_off + rule + 1), row, stride)) best_svc = rule;
                }
            verdict = best == kNoRule ? (V_ALLOW | (kNoRule << 2)) : (((__ldg(p.term + best) >> tshift) & 3u) | (best << 2));
            svc = best_svc == kNoRule ? kNoService : best_svc - p.n_waf_rules;
        }
        if (WARP) {
            const uint32_t lv = __shfl_sync(FULL, verdict, leader), ls = __shfl_sync(FULL, svc, leader);
            if (need_eval && same) { verdict = lv; svc = ls; }
        }
    }
    if (!valid) return;
    p.verdict[r] = verdict;
    // http_listener.rs:266-272: only a request the rules let through reaches the services; the first service whose
    // route is absent or true takes it, none => 404 (kNoService)
    if (p.service) p.service[r] = (uint16_t)(((verdict & 3u) == V_ALLOW && routes) ? svc : kNoService);
}

__device__ __noinline__ void request_epilogue(const KParams& p, uint32_t r, uint32_t* row, uint32_t stride) {
    request_epilogue_t<false>(p, r, row, stride, true);
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32_v(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

// Accept events of one walked word (all four states were hot, at least one is accepting): `s01`/`s23` hold the four
// 16-bit states the speculative walk produced, `m4` the bytes that belong to the field.  One-atom FIRE lists are
// resolved from the shared-memory acc1 table; anything else takes the general event list in global memory.
__device__ __noinline__ uint32_t events_word(const KParams& p, const UnitDesc* ud, uint32_t acc1addr, uint32_t s01, uint32_t s23, uint32_t m4,
                                             uint32_t last, uint32_t* latch, uint32_t* row, uint32_t stride) {
    const uint32_t acclo = ud->acc_lo;
    for (uint32_t b = 0; b < 4; ++b) {
        if (!((m4 >> b) & 1u)) continue;
        const uint32_t st = ((b < 2 ? s01 : s23) >> (16 * (b & 1))) & 0xFFFFu;
        if (st >= acclo && st != last) {
            const uint32_t a1 = lds_u16(acc1addr + 2u * (st - acclo));
            if (a1 != 0xFFFFu) {
                row[(a1 >> 5) * stride] |= 1u << (a1 & 31);
                last = st;
            } else {
                const bool pure = run_events(p.acc_idx, p.acc_events, ud->acc_base + st - acclo, row, stride, latch);
                last = pure ? st : 0xFFFFFFFFu;
            }
        }
    }
    return last;
}

// Careful re-walk of one 32-bit word of a field (rare): true transitions from the full table in global memory,
// accept events with latches.  `m4` selects which of the 4 bytes belong to the field.
__device__ __noinline__ void slow_word(const KParams& p, const UnitDesc* ud, uint32_t clsaddr, uint32_t w, uint32_t m4, uint32_t* state,
                                       uint32_t* last, uint32_t* latch, uint32_t* row, uint32_t stride) {
    const uint16_t* tbl = reinterpret_cast<const uint16_t*>(p.arena + ud->tbl_off);
    uint32_t st = *state, la = *last;
    const uint32_t C = ud->n_classes, acclo = ud->acc_lo;
    for (uint32_t b = 0; b < 4; ++b) {
        if (!((m4 >> b) & 1u)) continue;
        const uint32_t byte = (w >> (8 * b)) & 0xFFu;
        st = __ldg(tbl + st * C + lds_u8(clsaddr + byte));
        if (st >= acclo && st != la) {
            const bool pure = run_events(p.acc_idx, p.acc_events, ud->acc_base + st - acclo, row, stride, latch);
            la = pure ? st : 0xFFFFFFFFu;
        }
    }
    *state = st;
    *last = la;
}

__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1) waf_verdict_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const SmemLayout L = smem_layout(p.image_bytes, p.n_units, p.atom_words, p.n_slots);
    uint8_t* s_img = smem + L.image;
    UnitDesc* s_units = reinterpret_cast<UnitDesc*>(smem + L.units);
    uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem + L.rows);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L.misc);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const uint32_t Aw = p.atom_words;
    const uint32_t U = p.n_units;

    // ---- one-time staging: table image via TMA bulk copies, unit descriptors by plain loads ----
    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t stage_bytes = r16(p.image_bytes);
    if (tid == 0 && stage_bytes) {
        mbar_expect_tx(s_bar, stage_bytes);
        for (uint32_t o = 0; o < stage_bytes; o += 32768u) {
            uint32_t n = stage_bytes - o < 32768u ? stage_bytes - o : 32768u;
            bulk_g2s(s_img + o, p.image + o, n, s_bar);
        }
    }
    for (uint32_t i = tid; i < U * (sizeof(UnitDesc) / 4); i += kThreads)
        reinterpret_cast<uint32_t*>(s_units)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.units) + i);
    if (stage_bytes) mbar_wait(s_bar, 0);
    __syncthreads();

    // private bitmap rows: word w of row k of this lane lives at s_rows[(k * Aw + w) * kThreads + tid]
    const uint32_t stride = kThreads;
    uint32_t* my_rows = s_rows + tid;

    if (U == 0) {
        // no string predicate at all: only the per-request epilogue runs
        for (uint32_t r = blockIdx.x * kThreads + tid; r < p.n; r += gridDim.x * kThreads) {
            for (uint32_t w = 0; w < Aw; ++w) my_rows[w * stride] = 0;
            request_epilogue(p, r, my_rows, stride);
        }
        return;
    }

    // 32-bit shared-window addresses (computed once: no per-access generic->shared conversion)
    const uint32_t a_img = smem_u32(s_img);
    const uint32_t a_units = smem_u32(s_units);
    const uint32_t a_rows = smem_u32(my_rows);
    const uint32_t a_ext = smem_u32(smem + L.ext) + tid * 4u;  // word k of this lane: a_ext + k * kThreads * 4
    constexpr uint32_t kExtStride = kThreads * 4u;
    const uint32_t a_ext_w = a_ext - lane * 4u;    // the same for lane 0 of this warp
    const uint32_t a_rows_w = a_rows - lane * 4u;
    // lane k of a warp fetches word k of a claimed request's offsets: (field slot k/2, entry r + k%2)
    const uint32_t* my_off = lane < 2u * p.n_slots ? p.off[p.slot_field[lane >> 1]] + (lane & 1u) : nullptr;

    // ---- per-lane state ----
    bool c_have = false;   // a unit is being scanned (its chunk for this iteration is in `cur`)
    bool n_have = false;   // the next unit is prepared: extents known, first chunk load issued into `nxt`
    bool n_new = false;    //   ... and it is unit 0 of the queued request
    bool q_have = false;   // a request is queued: claimed, bitmap row cleared, field offsets landing in `ext`
    bool q_fresh = false;  //   ... claimed in this very iteration (offsets not yet usable)
    bool own = false;      // a request is in progress (between its first adoption and the end of its last unit)
    bool p_have = false;   // a finished request waits for its epilogue
    uint32_t c_req = 0, c_unit = 0, c_rowi = 0;
    uint32_t c_base = 0, c_start = 0, c_end = 0, c_state = 0, c_C2 = 0, c_lim = 0, c_trap = 0, c_acclo = 0, c_clsaddr = 0, c_hotaddr = 0, c_acc1 = 0, c_end1 = 0;
    uint32_t c_latch = 0, c_last = 0xFFFFFFFFu;
    const uint8_t* c_col = nullptr;
    uint32_t n_unit = 0, n_start = 0, n_end = 0;
    const uint8_t* n_col = nullptr;
    uint32_t q_req = 0, q_rowi = 0;
    uint32_t p_req = 0, p_rowi = 0;
    constexpr int kVec = kChunk / 16;
    constexpr uint32_t kAlign = ~(uint32_t)(kChunk - 1);
    uint4 cur[kVec], nxt[kVec];
#pragma unroll
    for (int v = 0; v < kVec; ++v) cur[v] = nxt[v] = make_uint4(0, 0, 0, 0);
    // warp-uniform pool of claimed requests
    uint32_t pool_next = 0, pool_end = 0;
    bool pool_dry = p.n == 0;

    auto flush = [&]() {
        if (p_have) request_epilogue(p, p_req, my_rows + p_rowi * Aw * stride, stride);
        p_have = false;
    };

    for (;;) {
        // ---- (1) rotate: continue the current unit or adopt the prepared one ----
        if (c_have) {
            c_base += kChunk;
#pragma unroll
            for (int v = 0; v < kVec; ++v) cur[v] = nxt[v];
        } else if (n_have) {
            const uint32_t ua = a_units + n_unit * (uint32_t)sizeof(UnitDesc);
            if (n_new) {
                c_req = q_req;
                c_rowi = q_rowi;
                q_have = false;
                own = true;
            }
            c_unit = n_unit;
            c_start = n_start;
            c_end = n_end;
            c_base = n_start & kAlign;
            c_col = n_col;
            c_C2 = 2u * lds_u32(ua + offsetof(UnitDesc, n_classes));
            c_state = lds_u32(ua + offsetof(UnitDesc, start_state));
            c_trap = lds_u32(ua + offsetof(UnitDesc, hot_states));
            c_lim = lds_u32(ua + offsetof(UnitDesc, lim));
            c_acclo = lds_u32(ua + offsetof(UnitDesc, acc_lo));
            c_clsaddr = a_img + lds_u32(ua + offsetof(UnitDesc, cls_off));
            c_hotaddr = a_img + lds_u32(ua + offsetof(UnitDesc, hot_off));
            c_acc1 = a_img + lds_u32(ua + offsetof(UnitDesc, acc1_off));
            c_end1 = a_img + lds_u32(ua + offsetof(UnitDesc, end1_off));
            c_latch = 0;
            c_last = 0xFFFFFFFFu;
#pragma unroll
            for (int v = 0; v < kVec; ++v) cur[v] = nxt[v];
            c_have = true;
            n_have = false;
        }
        const bool any_have = __any_sync(0xFFFFFFFFu, c_have);
        if (!any_have && pool_dry && pool_next == pool_end && !__any_sync(0xFFFFFFFFu, q_have)) {
            flush();
            break;
        }

        // ---- (2) queue the next request early: while scanning the last unit, or when idle ----
        const bool last_unit = c_unit + 1 >= U;
        const bool want_claim = !q_have && (own ? (c_have && last_unit) : !n_have);
        // rows: current + pending + queued would be three; the pending one goes first
        if (__any_sync(0xFFFFFFFFu, want_claim && own && p_have)) flush();
        const uint32_t need_mask = __ballot_sync(0xFFFFFFFFu, want_claim);
        q_fresh = false;
        if (need_mask) {
            if (pool_next == pool_end && !pool_dry) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(p.work_counter, kClaim);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base >= p.n) pool_dry = true;
                else { pool_next = base; pool_end = min(base + kClaim, p.n); }
            }
            const uint32_t rank = __popc(need_mask & ((1u << lane) - 1u));
            const bool got = want_claim && pool_next + rank < pool_end;
            if (got) {
                q_req = pool_next + rank;
                q_rowi = own ? (c_rowi ^ 1u) : (p_have ? (p_rowi ^ 1u) : 0u);
                q_have = true;
                q_fresh = true;
            }
            // The per-request setup is done by the whole warp for each claiming lane `t` (claims trickle in one or two
            // lanes at a time, so doing it in the claiming lane alone would run at 1/32 efficiency): lane k fetches word k
            // of the field offsets of t's request into t's `ext` slots (cp.async, lands before the next iteration's use)
            // and lanes < Aw clear t's bitmap row.
            uint32_t gm = __ballot_sync(0xFFFFFFFFu, got);
            while (gm) {
                const uint32_t t = __ffs(gm) - 1u;
                gm &= gm - 1u;
                const uint32_t req_t = pool_next + __popc(need_mask & ((1u << t) - 1u));
                const uint32_t rowi_t = __shfl_sync(0xFFFFFFFFu, q_rowi, t);
                if (lane < 2u * p.n_slots) cp_async4(a_ext_w + t * 4u + lane * kExtStride, my_off + req_t);
                for (uint32_t w = lane; w < Aw; w += 32u) sts_u32(a_rows_w + t * 4u + (rowi_t * Aw + w) * stride * 4u, 0u);
            }
            __syncwarp();
            pool_next = min(pool_end, pool_next + (uint32_t)__popc(need_mask));
        }

        // ---- (3) issue the loads each lane consumes in the NEXT iteration ----
        const bool finishing = c_have && (c_end <= c_base + kChunk);
        const bool to_new = q_have && !q_fresh && !n_have && (own ? (finishing && last_unit) : true);
        const bool to_same = finishing && !last_unit;
        if (__any_sync(0xFFFFFFFFu, to_new)) {
            // offsets of queued requests were fetched by other lanes of the warp in an earlier iteration
            cp_async_commit_wait();
            __syncwarp();
        }
        if (to_same || to_new) {
            n_unit = to_new ? 0u : c_unit + 1u;
            n_new = to_new;
            const uint32_t ua = a_units + n_unit * (uint32_t)sizeof(UnitDesc);
            const uint32_t sl = lds_u32(ua + offsetof(UnitDesc, field_slot));
            n_start = lds_u32(a_ext + (2u * sl) * kExtStride);
            n_end = lds_u32(a_ext + (2u * sl + 1u) * kExtStride);
            n_col = p.col[lds_u32(ua + offsetof(UnitDesc, field))];
            const uint8_t* src = n_col + (n_start & kAlign);
#pragma unroll
            for (int v = 0; v < kVec; ++v) nxt[v] = ld_nc_v4(src + 16 * v);
            n_have = true;
        } else if (c_have && !finishing) {
            const uint8_t* src = c_col + c_base + kChunk;
#pragma unroll
            for (int v = 0; v < kVec; ++v) nxt[v] = ld_nc_v4(src + 16 * v);
#if PGW_L2_PREFETCH
            // pull the line a few chunks ahead into L2 so the next loads see L2 rather than HBM latency
            if (c_end > c_base + PGW_L2_PREFETCH) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + PGW_L2_PREFETCH));
#endif
        }

        // ---- (4) walk the bytes of the current chunk that belong to the field ----
        if (any_have) {
            uint32_t mk = 0;  // bit k set: byte k of the chunk belongs to this lane's field
            if (c_have) {
                const uint32_t lo = c_start > c_base ? c_start - c_base : 0u;
                const uint32_t hi = min(c_end - c_base, (uint32_t)kChunk);
                mk = (hi >= 32u ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
            }
#pragma unroll
            for (int wi = 0; wi < kChunk / 4; ++wi) {
                const uint32_t m4 = (mk >> (4 * wi)) & 0xFu;
                if (!__any_sync(0xFFFFFFFFu, m4)) continue;
                const uint4 q = cur[wi / 4];
                const uint32_t w = (wi % 4) == 0 ? q.x : (wi % 4) == 1 ? q.y : (wi % 4) == 2 ? q.z : q.w;
                // speculative walk on the shared-memory rows: transitions to cold states lead to the absorbing trap row
                uint32_t spec = min(c_state, c_trap);
                uint32_t sv[4];
#pragma unroll
                for (int bi = 0; bi < 4; ++bi) {
                    const uint32_t byte = __byte_perm(w, 0, 0x4440 + bi);
                    const uint32_t cls = lds_u8(c_clsaddr + byte);
                    const uint32_t st = lds_u16(c_hotaddr + spec * c_C2 + 2u * cls);
                    spec = (m4 & (1u << bi)) ? st : spec;
                    sv[bi] = spec;
                }
                const uint32_t mx = max(max(max(sv[0], sv[1]), max(sv[2], sv[3])), c_state);
                if (mx >= c_lim) {
                    uint32_t* row = my_rows + c_rowi * Aw * stride;
                    if (mx >= c_trap) {
                        // a cold state is involved: re-walk the word on the full table (copies keep the fast-path state in registers)
                        uint32_t t_state = c_state, t_last = c_last, t_latch = c_latch;
                        slow_word(p, &s_units[c_unit], c_clsaddr, w, m4, &t_state, &t_last, &t_latch, row, stride);
                        c_state = t_state;
                        c_last = t_last;
                        c_latch = t_latch;
                    } else {
                        if (max(max(sv[0], sv[1]), max(sv[2], sv[3])) >= c_acclo) {
                            // accept events straight from the four states in registers; one-atom FIRE lists are resolved from the
                            // shared-memory acc1 table inline, anything else (latches, multi-atom lists) goes out of line
                            bool general = false;
#pragma unroll
                            for (int bi = 0; bi < 4; ++bi) {
                                const uint32_t st = sv[bi];
                                if ((m4 & (1u << bi)) && st >= c_acclo && st != c_last) {
                                    const uint32_t a1 = lds_u16(c_acc1 + 2u * (st - c_acclo));
                                    if (a1 != 0xFFFFu) {
                                        const uint32_t wa = a_rows + (c_rowi * Aw + (a1 >> 5)) * stride * 4u;
                                        sts_u32(wa, lds_u32_v(wa) | (1u << (a1 & 31)));
                                        c_last = st;
                                    } else {
                                        general = true;
                                    }
                                }
                            }
                            if (general) {
                                uint32_t t_latch = c_latch;
                                c_last = events_word(p, &s_units[c_unit], c_acc1, sv[0] | (sv[1] << 16), sv[2] | (sv[3] << 16), m4, 0xFFFFFFFFu, &t_latch, row, stride);
                                c_latch = t_latch;
                            }
                        }
                        c_state = spec;
                    }
                } else {
                    c_state = spec;
                }
            }
            if (finishing) {
                // end-of-field events of the final state: resolved from the shared-memory end1 table when the state is hot
                uint32_t e1 = 0xFFFFu;
                if (c_state < c_trap) e1 = lds_u16(c_end1 + 2u * c_state);
                if (e1 != 0xFFFEu) {
                    if (e1 != 0xFFFFu) {
                        const uint32_t wa = a_rows + (c_rowi * Aw + (e1 >> 5)) * stride * 4u;
                        sts_u32(wa, lds_u32_v(wa) | (1u << (e1 & 31)));
                    } else {
                        const UnitDesc& ud = s_units[c_unit];
                        uint32_t t_latch = c_latch;
                        if (ud.end_any) run_events(p.end_idx, p.end_events, ud.end_base + c_state, my_rows + c_rowi * Aw * stride, stride, &t_latch);
                    }
                }
                c_have = false;
                if (last_unit) {
                    p_have = true;
                    p_req = c_req;
                    p_rowi = c_rowi;
                    own = false;
                }
            }
        }
    }
}

// =====================================================================================================================
// Stream scan (kernel path "v5"): each scan unit's field column is one contiguous byte stream.  A warp owns a block of
// kStreamNB consecutive requests of one unit and walks the block's bytes in windows of 512 B: lane j takes the 16 bytes
// [W+16j, W+16j+16) with one coalesced 128-bit load.  Lane 0 starts from the exact carried state; the other lanes start
// from a speculated state (the unit's idle state warmed up on the 4 preceding bytes, or the DFA start state if their
// segment begins a request).  Validation: lane j's assumed start must equal lane j-1's end state (shuffle + vote);
// mismatching lanes re-walk from the correct state until the chain is consistent, which makes every state exact by
// induction.  Request boundaries inside a segment reset the DFA to its start state.  Side effects (accept events,
// end-of-field events) are applied after validation, to atom bitmaps in global memory (atomicOr, rare).
// The verdict is produced by waf_epilogue_kernel once every unit has been scanned.
// =====================================================================================================================
constexpr int kStreamThreads = 512;
constexpr int kStreamNB = 64;  // requests per task

struct StreamCtx {
    const UnitDesc* ud;     // shared memory
    uint32_t clsaddr;       // shared-window address of the class map
    const uint16_t* gtbl;   // full transition table (global)
    uint32_t C;
    uint32_t D0;
    uint32_t acclo;
    uint32_t end1addr;      // shared-window address of end1 (valid for states < hot)
    uint32_t hot;
    uint32_t Aw;
    uint32_t* rows;         // global atom bitmaps [n][Aw]
    const uint32_t* s_off;  // this task's offsets (shared memory), s_off[i] = off[r0 + i]
    uint32_t r0, nreq;
};

__device__ __forceinline__ void st_fire_list(const KParams& p, const uint32_t* idx, const uint32_t* events, uint32_t ci, uint32_t* row, uint32_t* latch) {
    uint32_t a = __ldg(idx + ci), b = __ldg(idx + ci + 1);
    uint32_t l = *latch;
    for (uint32_t i = a; i < b; ++i) {
        const uint32_t e = __ldg(events + i);
        const uint32_t kind = e >> kEvKindShift, lb = 1u << ((e >> kEvLatchShift) & 31u), at = e & kEvAtomMask;
        if (kind == 0u || (kind == 1u && (l & lb))) atomicOr(row + (at >> 5), 1u << (at & 31));
        else if (kind == 2u) l &= ~lb;
        else if (kind == 3u) l |= lb;
    }
    *latch = l;
}

// end-of-field events of `state` for request `req`
__device__ __forceinline__ void st_apply_end(const KParams& p, const StreamCtx& c, uint32_t state, uint32_t req, uint32_t* latch) {
    uint32_t e1 = 0xFFFFu;
    if (state < c.hot) e1 = lds_u16(c.end1addr + 2u * state);
    if (e1 == 0xFFFEu) return;
    uint32_t* row = c.rows + (size_t)req * c.Aw;
    if (e1 != 0xFFFFu) atomicOr(row + (e1 >> 5), 1u << (e1 & 31));
    else st_fire_list(p, p.end_idx, p.end_events, c.ud->end_base + state, row, latch);
}

// Exact walk of one 16-byte segment on the full table.  `vm`: bytes that belong to the block; `bm`: positions where a
// new request starts.  `req` = request (absolute index) owning the first valid byte.  With `apply`, accept and
// end-of-field events are applied to the global bitmaps.  Returns the end state.
__device__ __noinline__ uint32_t st_careful_segment(const KParams& p, const StreamCtx& c, uint4 data, uint32_t seg, uint32_t vm, uint32_t bm, uint32_t state,
                                                    uint32_t req, bool apply, uint32_t* latch_io) {
    const uint32_t words[4] = {data.x, data.y, data.z, data.w};
    uint32_t latch = *latch_io;
    uint32_t last = 0xFFFFFFFFu;
    for (uint32_t k = 0; k < 16; ++k) {
        if (!((vm >> k) & 1u)) continue;
        if ((bm >> k) & 1u) {
            if (apply) st_apply_end(p, c, state, req, &latch);
            state = c.D0;
            latch = 0;
            last = 0xFFFFFFFFu;
            // the request that owns this byte: the last one starting at or before it (empty requests share offsets)
            const uint32_t pos = seg + k;
            uint32_t i = req - c.r0 + 1;
            while (i + 1 <= c.nreq && c.s_off[i + 1] <= pos) ++i;
            req = c.r0 + i;
        }
        const uint32_t byte = (words[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        state = __ldg(c.gtbl + state * c.C + lds_u8(c.clsaddr + byte));
        if (apply && state >= c.acclo && state != last) {
            const uint32_t l0 = latch;
            st_fire_list(p, p.acc_idx, p.acc_events, c.ud->acc_base + state - c.acclo, c.rows + (size_t)req * c.Aw, &latch);
            last = (c.ud->has_latch || l0 != latch) ? 0xFFFFFFFFu : state;  // only plain FIRE lists are idempotent
        }
    }
    *latch_io = latch;
    return state;
}

struct StreamWalk {
    uint32_t end, mx, nb, pre0, pre1;
};

// Re-walk of one segment on the shared-memory rows from an exact start state (rare path of the stream scan, cheaper than
// st_careful_segment: no global table reads).  Without `apply` it recomputes what P1 computes (end state, max state,
// states before the first two request boundaries); with `apply` it applies accept / end-of-field events.  Falls back to
// the full table when a cold state is met.
__device__ __noinline__ void st_rewalk(const KParams& p, const StreamCtx& c, uint32_t hotaddr, uint32_t C2, uint32_t acc1addr, uint4 data, uint32_t seg,
                                       uint32_t vm, uint32_t bm, uint32_t start, uint32_t req, bool apply, uint32_t* latch_io, StreamWalk* out) {
    const uint32_t words[4] = {data.x, data.y, data.z, data.w};
    const uint32_t trap = c.hot;
    uint32_t latch = *latch_io, last = 0xFFFFFFFFu;
    uint32_t s = start, mx = start >= trap ? start : 0, nb = 0, pre0 = 0, pre1 = 0;
    const uint32_t req_in = req;
    bool cold = start >= trap;
    for (uint32_t k = 0; k < 16 && !cold; ++k) {
        if (!((vm >> k) & 1u)) continue;
        if ((bm >> k) & 1u) {
            if (nb == 0) pre0 = s;
            else if (nb == 1) pre1 = s;
            ++nb;
            if (apply) st_apply_end(p, c, s, req, &latch);
            s = c.D0;
            latch = 0;
            last = 0xFFFFFFFFu;
            const uint32_t pos = seg + k;
            uint32_t i = req - c.r0 + 1;
            while (i + 1 <= c.nreq && c.s_off[i + 1] <= pos) ++i;
            req = c.r0 + i;
        }
        const uint32_t byte = (words[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        s = lds_u16(hotaddr + s * C2 + 2u * lds_u8(c.clsaddr + byte));
        mx = max(mx, s);
        if (s >= trap) { cold = true; break; }
        if (apply && s >= c.acclo && s != last) {
            const uint32_t a1 = lds_u16(acc1addr + 2u * (s - c.acclo));
            uint32_t* row = c.rows + (size_t)req * c.Aw;
            if (a1 != 0xFFFFu) {
                atomicOr(row + (a1 >> 5), 1u << (a1 & 31));
                last = s;
            } else {
                st_fire_list(p, p.acc_idx, p.acc_events, c.ud->acc_base + s - c.acclo, row, &latch);
            }
        }
    }
    if (cold) {
        // a cold state: redo the whole segment exactly on the full table (events are idempotent / replayed from the same latch)
        uint32_t l2 = *latch_io;
        s = st_careful_segment(p, c, data, seg, vm, bm, start, req_in, apply, &l2);
        latch = l2;
        mx = trap;  // forces the apply pass for this lane
        nb = 3;
    }
    *latch_io = latch;
    out->end = s;
    out->mx = mx;
    out->nb = nb;
    out->pre0 = pre0;
    out->pre1 = pre1;
}

__global__ void __launch_bounds__(kStreamThreads, 2) waf_stream_scan_kernel(const __grid_constant__ KParams p, uint32_t* __restrict__ rows,
                                                                            uint32_t* __restrict__ task_counter, uint32_t n_blocks, uint32_t n_tasks) {
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: image | units | per-warp offsets
    const uint32_t img_bytes = r16(p.image_bytes);
    uint8_t* s_img = smem;
    UnitDesc* s_units = reinterpret_cast<UnitDesc*>(smem + img_bytes);
    uint32_t* s_offs_all = reinterpret_cast<uint32_t*>(smem + img_bytes + r16(p.n_units * (uint32_t)sizeof(UnitDesc)));
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + img_bytes + r16(p.n_units * (uint32_t)sizeof(UnitDesc)) + (kStreamThreads / 32) * (kStreamNB + 4) * 4);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && img_bytes) {
        mbar_expect_tx(s_bar, img_bytes);
        for (uint32_t o = 0; o < img_bytes; o += 32768u) {
            uint32_t n = img_bytes - o < 32768u ? img_bytes - o : 32768u;
            bulk_g2s(s_img + o, p.image + o, n, s_bar);
        }
    }
    for (uint32_t i = tid; i < p.n_units * (sizeof(UnitDesc) / 4); i += kStreamThreads)
        reinterpret_cast<uint32_t*>(s_units)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.units) + i);
    if (img_bytes) mbar_wait(s_bar, 0);
    __syncthreads();

    const uint32_t a_img = smem_u32(s_img);
    uint32_t* s_off = s_offs_all + warp * (kStreamNB + 4);
    const uint32_t a_off = smem_u32(s_off);
    const uint32_t FULL = 0xFFFFFFFFu;

    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(task_counter, 1u);
        t = __shfl_sync(FULL, t, 0);
        if (t >= n_tasks) break;
        const uint32_t u = t / n_blocks, b = t - u * n_blocks;
        const UnitDesc& ud = s_units[u];
        StreamCtx c;
        c.ud = &ud;
        c.clsaddr = a_img + ud.cls_off;
        c.gtbl = reinterpret_cast<const uint16_t*>(p.arena + ud.tbl_off);
        c.C = ud.n_classes;
        c.D0 = ud.start_state;
        c.acclo = ud.acc_lo;
        c.end1addr = a_img + ud.end1_off;
        c.hot = ud.hot_states;
        c.Aw = p.atom_words;
        c.rows = rows;
        c.s_off = s_off;
        c.r0 = b * kStreamNB;
        c.nreq = min((uint32_t)kStreamNB, p.n - c.r0);
        const uint32_t C2 = 2u * ud.n_classes, trap = ud.hot_states, lim = ud.lim, idle = ud.idle_state;
        const uint32_t hotaddr = a_img + ud.hot_off, acc1addr = a_img + ud.acc1_off;
        const bool has_latch = ud.has_latch != 0;
        const uint8_t* col = p.col[ud.field];
        const uint32_t* goff = p.off[ud.field] + c.r0;
        __syncwarp();
        for (uint32_t i = lane; i <= c.nreq; i += 32) s_off[i] = __ldg(goff + i);
        __syncwarp();
        const uint32_t B0 = s_off[0], B1 = s_off[c.nreq];
        uint32_t carry = c.D0;
        uint32_t latch = 0;  // warp-uniform
        uint32_t prev_w3 = 0;  // last word of lane 31 of the previous window (warm-up bytes for lane 0 are never needed: lane 0 is exact)

        for (uint32_t Wb = B0 & ~15u; Wb < B1; Wb += 512u) {
            const uint32_t seg = Wb + 16u * lane;
            uint4 d = make_uint4(0, 0, 0, 0);
            if (seg < B1 && seg + 16u > B0) d = ld_nc_v4(col + seg);
            // valid bytes of this segment
            uint32_t vm = 0;
            {
                const uint32_t lo = B0 > seg ? min(B0 - seg, 16u) : 0u;
                const uint32_t hi = B1 > seg ? min(B1 - seg, 16u) : 0u;
                if (hi > lo) vm = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
            }
            // request owning the first valid byte, and the request starts inside the segment
            uint32_t ri = 0, bm = 0;
            bool starts_exact = false;
            if (vm) {
                const uint32_t p0 = seg + (__ffs(vm) - 1);
                // upper_bound over s_off[0..nreq]: first index with s_off[i] > p0
                uint32_t lo = 0, hi = c.nreq + 1;
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (lds_u32_v(a_off + 4u * mid) <= p0) lo = mid + 1;
                    else hi = mid;
                }
                ri = lo - 1;  // last request starting at or before p0: it is non-empty and owns p0
                starts_exact = lds_u32_v(a_off + 4u * ri) == p0;
                for (uint32_t i = ri + 1; i < c.nreq; ++i) {
                    const uint32_t o = lds_u32_v(a_off + 4u * i);
                    if (o >= seg + 16u) break;
                    if (o > p0 && lds_u32_v(a_off + 4u * (i + 1)) > o) bm |= 1u << (o - seg);  // non-empty request starting at o
                }
            }
            const uint32_t req0 = c.r0 + ri;

            // ---- P1: speculative walk on the shared-memory rows (no side effects) ----
            const uint32_t words[4] = {d.x, d.y, d.z, d.w};
            uint32_t assumed;
            {
                const uint32_t up2 = __shfl_up_sync(FULL, d.z, 1);
                const uint32_t up3 = __shfl_up_sync(FULL, d.w, 1);
                if (lane == 0 || starts_exact) assumed = starts_exact ? c.D0 : carry;
                else {
                    // warm up on the 8 bytes before the segment (the previous lane's last two words)
                    uint32_t s = idle;
#pragma unroll
                    for (int bi = 0; bi < 8; ++bi) {
                        const uint32_t byte = __byte_perm(bi < 4 ? up2 : up3, 0, 0x4440 + (bi & 3));
                        s = lds_u16(hotaddr + min(s, trap) * C2 + 2u * lds_u8(c.clsaddr + byte));
                    }
                    assumed = s;
                }
            }
            uint32_t s_end = assumed, mx = 0, nb = 0, s_pre0 = 0, s_pre1 = 0;
            {
                uint32_t s = min(assumed, trap);
                mx = assumed >= trap ? assumed : 0;
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    const uint32_t w = words[wi];
                    const uint32_t v4 = (vm >> (4 * wi)) & 0xFu, b4 = (bm >> (4 * wi)) & 0xFu;
                    if (!__any_sync(FULL, v4)) continue;
                    if (b4 == 0) {
#pragma unroll
                        for (int bi = 0; bi < 4; ++bi) {
                            const uint32_t byte = __byte_perm(w, 0, 0x4440 + bi);
                            const uint32_t st = lds_u16(hotaddr + s * C2 + 2u * lds_u8(c.clsaddr + byte));
                            s = (v4 & (1u << bi)) ? st : s;
                            mx = max(mx, s);
                        }
                    } else {
#pragma unroll
                        for (int bi = 0; bi < 4; ++bi) {
                            if (!(v4 & (1u << bi))) continue;
                            if (b4 & (1u << bi)) {
                                if (nb == 0) s_pre0 = s;
                                else if (nb == 1) s_pre1 = s;
                                ++nb;
                                s = min(c.D0, trap);
                            }
                            const uint32_t byte = __byte_perm(w, 0, 0x4440 + bi);
                            s = lds_u16(hotaddr + s * C2 + 2u * lds_u8(c.clsaddr + byte));
                            mx = max(mx, s);
                        }
                    }
                }
                s_end = s;
            }
            bool need_full = vm && (mx >= lim || nb > 2);  // accept events, a cold state, or more boundaries than recorded
            bool trapped = vm && mx >= trap;

            // ---- P2: make the chain of states exact ----
            const uint32_t p0pos = vm ? seg + (__ffs(vm) - 1) : 0u;
            const bool pre_b = vm && starts_exact && p0pos > B0;  // a request ends exactly where this segment starts
            uint32_t prev_end;
            for (;;) {
                prev_end = __shfl_up_sync(FULL, s_end, 1);
                if (lane == 0) prev_end = carry;
                const uint32_t true_start = starts_exact ? c.D0 : prev_end;
                const bool bad = vm && (trapped || assumed != true_start);
                if (!__any_sync(FULL, bad)) break;
                if (bad) {
                    uint32_t l2 = 0;
                    StreamWalk wk;
                    st_rewalk(p, c, hotaddr, C2, acc1addr, d, seg, vm, bm, true_start, req0, false, &l2, &wk);
                    s_end = wk.end;
                    mx = wk.mx;
                    nb = wk.nb;
                    s_pre0 = wk.pre0;
                    s_pre1 = wk.pre1;
                    assumed = true_start;
                    trapped = false;
                    need_full = mx >= lim || nb > 2;
                }
            }

            // ---- P3: side effects, from validated states ----
            // request whose field ends right before this segment (pre_b): its final state is the previous lane's end state
            uint32_t preq = 0;
            if (pre_b) {
                uint32_t i = ri - 1;
                while (i > 0 && s_off[i] == s_off[i + 1]) --i;
                preq = c.r0 + i;
            }
            bool pre_general = false;
            if (has_latch && vm) {
                // end-of-field lists that involve latches (or are not in the shared-memory table) must be applied in string order
                if (!need_full && nb) {
                    if (s_pre0 >= c.hot || lds_u16(c.end1addr + 2u * s_pre0) == 0xFFFFu) need_full = true;
                    if (nb > 1 && (s_pre1 >= c.hot || lds_u16(c.end1addr + 2u * s_pre1) == 0xFFFFu)) need_full = true;
                }
                if (pre_b && (prev_end >= c.hot || lds_u16(c.end1addr + 2u * prev_end) == 0xFFFFu)) pre_general = true;
            }
            const uint32_t bnd_mask = __ballot_sync(FULL, vm && (nb > 0 || pre_b));
            if (!has_latch) {
                uint32_t l2 = 0;
                if (pre_b) st_apply_end(p, c, prev_end, preq, &l2);
                if (need_full) {
                    StreamWalk wk;
                    st_rewalk(p, c, hotaddr, C2, acc1addr, d, seg, vm, bm, assumed, req0, true, &l2, &wk);
                } else if (nb) {
                    st_apply_end(p, c, s_pre0, req0, &l2);
                    if (nb > 1) {
                        // request owning the byte at the first boundary
                        const uint32_t pos = seg + (__ffs(bm) - 1);
                        uint32_t i = ri + 1;
                        while (i + 1 <= c.nreq && s_off[i + 1] <= pos) ++i;
                        st_apply_end(p, c, s_pre1, c.r0 + i, &l2);
                    }
                }
            } else {
                // latch events are order dependent: lanes that need the general path run one after the other, the latch
                // register travelling with them; a request boundary anywhere resets it.  Lists without latch kinds are
                // applied in parallel first (they commute with everything).
                {
                    uint32_t l2 = 0;
                    if (pre_b && !pre_general) st_apply_end(p, c, prev_end, preq, &l2);
                    if (!need_full && nb) {
                        st_apply_end(p, c, s_pre0, req0, &l2);
                        if (nb > 1) {
                            const uint32_t pos = seg + (__ffs(bm) - 1);
                            uint32_t i = ri + 1;
                            while (i + 1 <= c.nreq && s_off[i + 1] <= pos) ++i;
                            st_apply_end(p, c, s_pre1, c.r0 + i, &l2);
                        }
                    }
                }
                uint32_t m = __ballot_sync(FULL, need_full || pre_general);
                int prev_lane = -1;
                while (m) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    // boundaries in lanes strictly between the previous ordered lane and this one reset the latch
                    const uint32_t below_l = (1u << l) - 1u;
                    const uint32_t upto_prev = prev_lane < 0 ? 0u : ((2u << prev_lane) - 1u);
                    if (bnd_mask & below_l & ~upto_prev) latch = 0;
                    uint32_t lt = latch;
                    if ((int)lane == l) {
                        if (pre_general) st_apply_end(p, c, prev_end, preq, &lt);
                        if (pre_b) lt = 0;
                        if (need_full) {
                            StreamWalk wk;
                            st_rewalk(p, c, hotaddr, C2, acc1addr, d, seg, vm, bm, assumed, req0, true, &lt, &wk);
                        } else if (nb) lt = 0;
                    }
                    latch = __shfl_sync(FULL, lt, l);
                    prev_lane = l;
                }
                {
                    const uint32_t upto_prev = prev_lane < 0 ? 0u : ((2u << prev_lane) - 1u);
                    if (bnd_mask & ~upto_prev) latch = 0;
                }
            }

            // carry for the next window: end state of the last valid lane
            const uint32_t vlanes = __ballot_sync(FULL, vm != 0);
            const int last_lane = 31 - __clz(vlanes);
            carry = __shfl_sync(FULL, s_end, last_lane);
            prev_w3 = __shfl_sync(FULL, d.w, 31);
        }
        // end of the block: the last non-empty request ends at B1
        if (lane == 0 && B1 > B0) {
            uint32_t i = c.nreq - 1;
            while (i > 0 && s_off[i] == B1) --i;  // trailing empty requests
            uint32_t lt = latch;
            st_apply_end(p, c, carry, c.r0 + i, &lt);
        }
    }
}

// =====================================================================================================================
// Field scan (kernel path "field"): unit-major, lane-owned strings.  A warp works on ONE scan unit at a time, so all the
// per-unit parameters are warp-uniform; each lane owns one request's field of that unit and walks it 16 bytes per
// iteration exactly like the lane path (speculative 4-byte word walk on the shared-memory rows).  A lane that finishes
// takes the next request of the unit from a warp pool of 32 claimed requests whose field offsets were fetched coalesced
// one pool ahead, so the per-request setup is a shuffle.  Atom bits go to bitmaps in global memory (red.or, rare) and
// waf_epilogue_kernel turns them into verdicts.  Warps move to the next unit on their own when a unit runs dry.
// =====================================================================================================================
#ifndef PGW_FS_THREADS
#define PGW_FS_THREADS 1024
#endif
constexpr int kFsThreads = PGW_FS_THREADS;
constexpr uint32_t kFsSlotStride = kFsThreads * 4u;  // per-lane slots: request index, latch register, last fired state, next request
#ifndef PGW_FS_TICKET
#define PGW_FS_TICKET 64
#endif
constexpr uint32_t kFsTicket = PGW_FS_TICKET;  // requests per atomic claim (two 32-request pools: 128 and 256 measured worse, tail imbalance)
constexpr uint32_t kFsPoolBytes = 144;  // 33 offsets of a claimed pool (+pad), two buffers per warp

__device__ __forceinline__ void red_or(uint32_t* addr, uint32_t v) { asm volatile("red.global.or.b32 [%0], %1;" ::"l"(addr), "r"(v) : "memory"); }

// events of CSR row `ci` applied to a bitmap in global memory; true if all of them were plain FIREs
__device__ __forceinline__ bool fs_fire_list(const uint32_t* idx, const uint32_t* events, uint32_t ci, uint32_t* row, uint32_t* latch) {
    uint32_t a = __ldg(idx + ci), b = __ldg(idx + ci + 1);
    uint32_t l = *latch;
    bool pure = true;
    for (uint32_t i = a; i < b; ++i) {
        const uint32_t e = __ldg(events + i);
        const uint32_t kind = e >> kEvKindShift, lb = 1u << ((e >> kEvLatchShift) & 31u), at = e & kEvAtomMask;
        if (kind == 0u || (kind == 1u && (l & lb))) red_or(row + (at >> 5), 1u << (at & 31));
        else if (kind == 2u) l &= ~lb;
        else if (kind == 3u) l |= lb;
        pure &= kind == 0u;
    }
    *latch = l;
    return pure;
}

// accept events of four hot states of one word (at least one accepting); returns the new `last`
__device__ __noinline__ uint32_t fs_events_word(const KParams& p, const UnitDesc* ud, uint32_t acc1addr, uint32_t s01, uint32_t s23, uint32_t m4,
                                                uint32_t last, uint32_t* latch, uint32_t* row) {
    const uint32_t acclo = ud->acc_lo;
    // positions whose state is accepting
    uint32_t am = m4;
    if ((s01 & 0xFFFFu) < acclo) am &= ~1u;
    if ((s01 >> 16) < acclo) am &= ~2u;
    if ((s23 & 0xFFFFu) < acclo) am &= ~4u;
    if ((s23 >> 16) < acclo) am &= ~8u;
#pragma unroll 1
    while (am) {
        const int bi = __ffs(am) - 1;
        am &= am - 1u;
        const uint32_t st = ((bi < 2 ? s01 : s23) >> (16 * (bi & 1))) & 0xFFFFu;
        if (st == last) continue;
        const uint32_t a1 = lds_u16(acc1addr + 2u * (st - acclo));
        if (a1 != 0xFFFFu) {
            red_or(row + (a1 >> 5), 1u << (a1 & 31));
            last = st;
        } else {
            last = fs_fire_list(p.acc_idx, p.acc_events, ud->acc_base + st - acclo, row, latch) ? st : 0xFFFFFFFFu;
        }
    }
    return last;
}

// one word walked on the full table in global memory (a cold state is involved)
__device__ __noinline__ void fs_slow_word(const KParams& p, const UnitDesc* ud, uint32_t clsaddr, uint32_t w, uint32_t m4, uint32_t* state,
                                          uint32_t* last, uint32_t* latch, uint32_t* row) {
    const uint16_t* tbl = reinterpret_cast<const uint16_t*>(p.arena + ud->tbl_off);
    const uint32_t C = ud->n_classes, acclo = ud->acc_lo;
    uint32_t st = *state, la = *last;
#pragma unroll 1
    for (int bi = 0; bi < 4; ++bi) {
        if (!((m4 >> bi) & 1u)) continue;
        const uint32_t byte = (w >> (8 * bi)) & 0xFFu;
        st = __ldg(tbl + st * C + lds_u8(clsaddr + byte));
        if (st >= acclo && st != la) la = fs_fire_list(p.acc_idx, p.acc_events, ud->acc_base + st - acclo, row, latch) ? st : 0xFFFFFFFFu;
    }
    *state = st;
    *last = la;
}

__global__ void __launch_bounds__(kFsThreads, 1) waf_field_scan_kernel(const __grid_constant__ KParams p, uint32_t* __restrict__ rows,
                                                                       uint32_t* __restrict__ counters) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t img_bytes = r16(p.image_bytes);
    uint8_t* s_img = smem + ((0u - smem_u32(smem)) & 255u);  // class maps (image offsets u * 256) on 256-byte boundaries
    UnitDesc* s_units = reinterpret_cast<UnitDesc*>(s_img + img_bytes);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_img + img_bytes + r16(p.n_units * (uint32_t)sizeof(UnitDesc)));

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && img_bytes) {
        mbar_expect_tx(s_bar, img_bytes);
        for (uint32_t o = 0; o < img_bytes; o += 32768u) {
            uint32_t n = img_bytes - o < 32768u ? img_bytes - o : 32768u;
            bulk_g2s(s_img + o, p.image + o, n, s_bar);
        }
    }
    for (uint32_t i = tid; i < p.n_units * (sizeof(UnitDesc) / 4); i += kFsThreads)
        reinterpret_cast<uint32_t*>(s_units)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.units) + i);
    if (img_bytes) mbar_wait(s_bar, 0);
    __syncthreads();

    const uint32_t a_img = smem_u32(s_img);
    const uint32_t a_pool = smem_u32(s_bar) + 64u + (tid >> 5) * 2u * kFsPoolBytes;
    const uint32_t a_slot = smem_u32(s_bar) + 64u + (kFsThreads / 32) * 2u * kFsPoolBytes + tid * 4u;  // word k at a_slot + k * kFsSlotStride
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t Aw = p.atom_words, N = p.n;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (uint32_t u = 0; u < p.n_units; ++u) {
        const UnitDesc* ud = &s_units[u];  // for the out-of-line event paths
        const UnitDesc& cu = p.udesc[u];   // constant bank, uniform index
        const uint32_t C2 = 2u * cu.n_classes, D0 = cu.start_state, trap = cu.hot_states, lim = cu.lim, acclo = cu.acc_lo;
        const uint32_t clsaddr = a_img + cu.cls_off, hotaddr = a_img + cu.hot_off, acc1addr = a_img + cu.acc1_off, end1addr = a_img + cu.end1_off;
        const uint8_t* col = p.col[cu.field];
        const uint32_t* off = p.off[cu.field];
        uint32_t* ctr = counters + u;

        // warp pools of 32 claimed requests, double buffered in shared memory: buffer `pb` is being handed out, the other
        // one holds the next claim whose 33 field offsets are landing through cp.async (no registers, no stall)
        uint32_t pool_next = 0, pool_end = 0, pb = 0, ah_base = 0;
        // claims are pipelined three deep so that no global latency is ever waited for: `ticket` (atomicAdd issued, result
        // not looked at yet) -> `ahead` (offsets landing in the spare buffer) -> the pool being handed out
        uint32_t ticket = 0;
        bool tk_valid = false, ah_valid = false;
        auto issue_ticket = [&]() {
            if (lane == 0) ticket = atomicAdd(ctr, kFsTicket);  // a ticket covers kFsTicket / 32 consecutive pools
            tk_valid = true;
        };
        uint32_t more = 0;  // end of the current ticket's range (pools still to take from it start at ah_base + 32)
        auto claim_ahead = [&]() {
            ah_valid = false;
            uint32_t b = ah_base + 32u;
            if (b >= more) {
                if (!tk_valid) return;
                b = __shfl_sync(FULL, ticket, 0);
                if (b >= N) { tk_valid = false; return; }
                more = min(b + kFsTicket, N);
                issue_ticket();
            }
            ah_base = b;
            const uint32_t dst = a_pool + (pb ^ 1u) * kFsPoolBytes;
            cp_async4(dst + lane * 4u, off + min(b + lane, N));
            if (lane == 0) cp_async4(dst + 128u, off + min(b + 32u, N));
            asm volatile("cp.async.commit_group;" ::: "memory");
            ah_valid = true;
        };
        __syncwarp();
        if (N) issue_ticket();
        claim_ahead();

        bool have = false, pend = false;
        // hot per-lane state lives in registers; what only the (rare) event paths need -- request index, latch register,
        // last fired state -- lives in this lane's shared-memory slots
        uint32_t base = 0, skip = 0, end = 0, state = 0;
        uint4 cur = make_uint4(0, 0, 0, 0), nxt = make_uint4(0, 0, 0, 0);

        for (;;) {
            // ---- rotate: next chunk of the current string, or the first chunk of the string claimed last iteration ----
            skip = 0;
            if (pend) {
                // a claim leaves `base` one chunk before the string's first one, its low four bits carry the offset of the
                // first byte in that chunk; the request index waits in the lane's fourth slot word
                skip = base & 15u;
                base &= ~15u;
                state = D0;
                sts_u32(a_slot, lds_u32_v(a_slot + 3u * kFsSlotStride));
                sts_u32(a_slot + kFsSlotStride, 0u);
                sts_u32(a_slot + 2u * kFsSlotStride, 0xFFFFFFFFu);
                have = true;
                pend = false;
                base += 16u;
                cur = nxt;
            } else if (have) {
                base += 16u;
                cur = nxt;
            }
            // everything this iteration's walk needs from (base, end, skip) is derived here, so that a lane on its last
            // chunk can overwrite them with its next string right away
            const bool finishing = have && end <= base + 16u;
            uint32_t mk = 0;
            if (have) {
                const uint32_t hi = min(end - base, 16u);
                mk = ((1u << hi) - 1u) & ~((1u << skip) - 1u);
            }
            // ---- lanes that run out of bytes in this iteration (or are idle) take the next request of the pool ----
            const bool want = !have || finishing;
            bool do_ld = have && !finishing;
            uint32_t ld_off = base + 16u;
            const uint32_t need = __ballot_sync(FULL, want);
            if (need) {
                if (pool_next == pool_end && ah_valid) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                    pb ^= 1u;
                    pool_next = ah_base;
                    pool_end = min(ah_base + 32u, N);
                    claim_ahead();
                }
                const uint32_t idx = pool_next + __popc(need & lt_mask);
                if (want && idx < pool_end) {
                    const uint32_t sa = a_pool + pb * kFsPoolBytes + (idx & 31u) * 4u;
                    const uint32_t s0 = lds_u32_v(sa), e0 = lds_u32_v(sa + 4u);
                    if (e0 > s0) {  // empty fields are left to the epilogue kernel
                        sts_u32(a_slot + 3u * kFsSlotStride, idx);
                        pend = true;
                        end = e0;
                        ld_off = s0 & ~15u;
                        base = (ld_off - 16u) | (s0 & 15u);
                        do_ld = true;
                    }
                }
                pool_next = min(pool_end, pool_next + (uint32_t)__popc(need));
            }
            // one load per lane and iteration, consumed in the next one: the next chunk of the current string or the first
            // chunk of the string just claimed (a single instruction: two loads into the same registers would serialise)
            if (do_ld) nxt = ld_nc_v4(col + ld_off);
            if (!__any_sync(FULL, have)) {
                if (!__any_sync(FULL, pend) && pool_next == pool_end && !ah_valid) break;
                continue;
            }

            // ---- walk the bytes of this chunk that belong to the field (no per-word vote: some lane almost always has
            //      bytes in every word, the vote cost more than the words it skipped) ----
#pragma unroll
            for (int wi = 0; wi < 4; ++wi) {
                const uint32_t m4 = (mk >> (4 * wi)) & 0xFu;
                const uint32_t w = wi == 0 ? cur.x : wi == 1 ? cur.y : wi == 2 ? cur.z : cur.w;
                uint32_t spec = min(state, trap);
                uint32_t sv[4];
#pragma unroll
                for (int bi = 0; bi < 4; ++bi) {
                    // class maps sit on 256-byte boundaries of the shared window: one PRMT extracts the byte AND adds the base
                    const uint32_t cls = lds_u8(__byte_perm(w, clsaddr, 0x7650 + bi));
                    // column address first (independent of the state): the state chain is IMAD -> LDS -> SEL only
                    uint32_t colad = hotaddr + 2u * cls, ad;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ad) : "r"(spec), "r"(C2), "r"(colad));
                    const uint32_t st = lds_u16(ad);
                    spec = (m4 & (1u << bi)) ? st : spec;
                    sv[bi] = spec;
                }
                // a cold start state walks the trap row, so the four new states alone tell whether the word needs attention
                const uint32_t mx4 = max(max(sv[0], sv[1]), max(sv[2], sv[3]));
                if (mx4 >= lim) {
                    if (mx4 >= trap || mx4 >= acclo) {
                        uint32_t* row = rows + (size_t)lds_u32_v(a_slot) * Aw;
                        uint32_t t_latch = lds_u32_v(a_slot + kFsSlotStride), t_last = lds_u32_v(a_slot + 2u * kFsSlotStride);
                        if (mx4 >= trap) {
                            uint32_t t_state = state;
                            fs_slow_word(p, ud, clsaddr, w, m4, &t_state, &t_last, &t_latch, row);
                            spec = t_state;
                        } else {
                            // a string sitting in a sticky accepting state whose events were already applied: nothing to do
                            const uint32_t s01 = sv[0] | (sv[1] << 16), s23 = sv[2] | (sv[3] << 16), ll = t_last | (t_last << 16);
                            if (t_last > 0xFFFFu || s01 != ll || s23 != ll) {
                                // the common event -- one accepting position whose list is a single FIRE -- is applied inline
                                uint32_t am = m4;
                                if (sv[0] < acclo) am &= ~1u;
                                if (sv[1] < acclo) am &= ~2u;
                                if (sv[2] < acclo) am &= ~4u;
                                if (sv[3] < acclo) am &= ~8u;
                                const uint32_t pos = __ffs(am) - 1u;
                                const uint32_t st1 = ((pos < 2u ? s01 : s23) >> (16u * (pos & 1u))) & 0xFFFFu;
                                uint32_t a1 = 0xFFFFu;
                                if ((am & (am - 1u)) == 0u) a1 = lds_u16(acc1addr + 2u * (st1 - acclo));
                                if (a1 != 0xFFFFu) {
                                    if (st1 != t_last) red_or(row + (a1 >> 5), 1u << (a1 & 31));
                                    t_last = st1;
                                } else {
                                    t_last = fs_events_word(p, ud, acc1addr, s01, s23, m4, t_last, &t_latch, row);
                                }
                            }
                        }
                        sts_u32(a_slot + kFsSlotStride, t_latch);
                        sts_u32(a_slot + 2u * kFsSlotStride, t_last);
                    }
                }
                state = spec;
            }
            if (finishing) {
                uint32_t e1 = 0xFFFFu;
                if (state < trap) e1 = lds_u16(end1addr + 2u * state);
                if (e1 != 0xFFFEu) {
                    uint32_t* row = rows + (size_t)lds_u32_v(a_slot) * Aw;
                    if (e1 != 0xFFFFu) red_or(row + (e1 >> 5), 1u << (e1 & 31));
                    else if (cu.end_any) {
                        uint32_t t_latch = lds_u32_v(a_slot + kFsSlotStride);
                        fs_fire_list(p.end_idx, p.end_events, cu.end_base + state, row, &t_latch);
                    }
                }
                have = false;
                state = 0;  // an idle lane must not look like it sits in a cold or accepting state
            }
        }
    }
}

// Verdicts once every unit has been scanned: one thread per request.  Empty fields never reach the stream scan;
// their end-of-field events (the DFA's start state at end of input) are applied here.
__global__ void __launch_bounds__(256) waf_epilogue_kernel(const __grid_constant__ KParams p, uint32_t* __restrict__ rows) {
    // warp-uniform trip count: every lane of a warp goes through request_epilogue_t<true> together
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < p.n; base += gridDim.x * blockDim.x) {
        const uint32_t rr = base + (threadIdx.x & 31u);
        const bool valid = rr < p.n;
        const uint32_t r = valid ? rr : p.n - 1u;
        uint32_t* row = rows + (size_t)r * p.atom_words;
        if (valid)
            for (uint32_t u = 0; u < p.n_units; ++u) {
                const UnitDesc& ud = p.udesc[u];  // parameter bank (both callers keep n_units <= kMaxConstUnits)
                if (!ud.end_any) continue;
                const uint32_t* o = p.off[ud.field] + r;
                if (o[0] != o[1]) continue;
                uint32_t a = __ldg(p.end_idx + ud.end_base + ud.start_state), b = __ldg(p.end_idx + ud.end_base + ud.start_state + 1);
                for (uint32_t i = a; i < b; ++i) {
                    const uint32_t e = __ldg(p.end_events + i);
                    if ((e >> kEvKindShift) == 0u) row[(e & kEvAtomMask) >> 5] |= 1u << (e & 31);  // latch kinds cannot fire on an empty field
                }
            }
        __syncwarp();
        request_epilogue_t<true>(p, r, row, 1, valid);
    }
}

// GeoipDB::lookup for a batch of addresses (pingoo/geoip.rs:73-91)
__global__ void geoip_lookup_kernel(const __grid_constant__ KParams p, const uint8_t* __restrict__ ip,
                                    const uint8_t* __restrict__ is_v6, uint32_t n, uint32_t* __restrict__ asn_out,
                                    uint16_t* __restrict__ country_out) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint8_t* ip16 = ip + (size_t)r * 16;
        const bool v6 = is_v6[r] != 0;
        uint32_t asn = 0, country = (uint32_t)'X' | ((uint32_t)'X' << 8);
        bool skip;
        if (!v6) skip = ip16[0] == 127 || (ip16[0] >> 4) == 0xE;
        else {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(ip16);
            skip = ip16[0] == 0xFF || (w[0] == 0 && w[1] == 0 && w[2] == 0 && w[3] == 0x01000000u);
        }
        if (!skip && p.geo_loaded) {
            const LpmLeaf lf = p.leaves[lpm_lookup(p, ip16, v6)];
            asn = lf.asn;
            country = lf.country;
        }
        asn_out[r] = asn;
        country_out[r] = (uint16_t)country;
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// captcha client id: generate_captcha_client_id (pingoo/captcha.rs:409-421) for a batch --
// base64url-no-pad( SHA-256( ip octets (4 or 16) || user_agent || host ) ), 43 characters per request.
// One thread per request; the message is at most 16 + 256 + 256 bytes, i.e. nine 64-byte blocks.
// ---------------------------------------------------------------------------------------------------------------------
__constant__ uint32_t kSha256K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

__global__ void __launch_bounds__(128) captcha_client_id_kernel(const uint8_t* __restrict__ ip, const uint8_t* __restrict__ is_v6,
                                                                const uint8_t* __restrict__ ua_bytes, const uint32_t* __restrict__ ua_off,
                                                                const uint8_t* __restrict__ host_bytes, const uint32_t* __restrict__ host_off, uint32_t n,
                                                                uint8_t* __restrict__ out44) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t* ipp = ip + (size_t)r * 16;
    const uint32_t ipl = is_v6[r] ? 16u : 4u;
    const uint8_t* uap = ua_bytes + ua_off[r];
    const uint32_t ual = ua_off[r + 1] - ua_off[r];
    const uint8_t* hop = host_bytes + host_off[r];
    const uint32_t hol = host_off[r + 1] - host_off[r];
    const uint32_t total = ipl + ual + hol;
    const uint32_t n_blocks = (total + 9u + 63u) / 64u;
    auto msg_byte = [&](uint32_t i) -> uint32_t {
        if (i < ipl) return ipp[i];
        if (i < ipl + ual) return uap[i - ipl];
        if (i < total) return hop[i - ipl - ual];
        return i == total ? 0x80u : 0u;
    };
    uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (uint32_t b = 0; b < n_blocks; ++b) {
        uint32_t w[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const uint32_t p = b * 64u + 4u * t;
            w[t] = (msg_byte(p) << 24) | (msg_byte(p + 1) << 16) | (msg_byte(p + 2) << 8) | msg_byte(p + 3);
        }
        if (b == n_blocks - 1) {  // message length in bits, big endian, in the last eight bytes
            w[14] = 0;
            w[15] = total * 8u;
        }
        uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
        for (int t = 0; t < 64; ++t) {
            if (t >= 16) {
                const uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
                const uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
                const uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
                w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
            }
            const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t t1 = hh + S1 + ch + kSha256K[t] + w[t & 15];
            const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
            const uint32_t mj = (a & bb) ^ (a & c) ^ (bb & c);
            const uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
        }
        h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    // base64url without padding: 32 bytes -> 43 characters (+ a terminating 0 in the 44th byte)
    auto digest_byte = [&](uint32_t i) -> uint32_t { return i < 32u ? (h[i >> 2] >> (24u - 8u * (i & 3u))) & 0xFFu : 0u; };
    auto b64 = [](uint32_t v) -> uint8_t {
        return (uint8_t)(v < 26u ? 'A' + v : v < 52u ? 'a' + (v - 26u) : v < 62u ? '0' + (v - 52u) : v == 62u ? '-' : '_');
    };
    uint8_t* o = out44 + (size_t)r * 44;
    for (uint32_t i = 0, j = 0; i < 33u; i += 3u, j += 4u) {
        const uint32_t v = (digest_byte(i) << 16) | (digest_byte(i + 1) << 8) | digest_byte(i + 2);
        o[j] = b64(v >> 18);
        o[j + 1] = b64((v >> 12) & 63u);
        if (j + 2 < 43u) o[j + 2] = b64((v >> 6) & 63u);
        if (j + 3 < 43u) o[j + 3] = b64(v & 63u);
    }
    o[43] = 0;
}

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
