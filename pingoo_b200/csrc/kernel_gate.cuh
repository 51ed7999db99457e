// Part of kernels.cu (included inside namespace pgw { namespace { ... } }).
//
// Candidate gate (gate.hpp): the literal prefilter in front of the DFA scan -- what Rust `regex` does with memchr /
// Teddy in front of its automata (reference call path pingoo/rules.rs:38 -> bel -> regex 1.12.2), restated for a
// batch.  The gate is position-local (no automaton state, no request boundaries), so it is three kernels:
//
//   waf_gate_kernel      streams every gated field column as ONE flat byte range: a warp takes 512 consecutive bytes per
//                        iteration, every lane one coalesced 16-byte load.  Each lane tests the eight even-aligned
//                        4-byte windows that start in its 16 bytes against the field's level-1 blocked Bloom filter in
//                        shared memory (all gated fields' filters are resident together).  The warp's 32 verdicts are
//                        one word of the field's HIT BITMAP (one bit per 16-byte chunk of the column): one ballot, one
//                        store -- nothing else happens in the streaming loop, whatever the input looks like.
//   waf_gate_maybe_kernel    one thread per request: the bits of the chunks its field overlaps (one or two words, read
//                        coalesced); requests with a bit set (a few per cent) are listed per field.
//   waf_gate_resolve_kernel  one thread per listed request: its hit chunks' windows again, now against the exact gram
//                        table (gram -> mask of scan units); requests with a confirmed gram become the field's
//                        candidates (request, field start, field end, unit mask) for the scan kernel.
// Both lists are built with one block-wide scan and one atomicAdd per 1024 requests, in request order.
constexpr int kGateThreads = 1024;
constexpr uint32_t kGateCtrBytes = 64;   // front of the shared window (alignment pad)

__device__ __forceinline__ uint32_t shf_wrap_r(uint32_t x, uint32_t n) { return __funnelshift_r(x, x, n); }  // rotate: amount taken mod 32

// level 1: both bits of the window's Bloom word set?  (bit 0 of the result).  FMA pipe: one IMAD + two IMAD.HI; ALU
// pipe: one SHF (word index; the LDS scales it by 4), two SHF.W, and the caller's LOP3 that ORs the eight windows.
__device__ __forceinline__ uint32_t gate_l1(const uint32_t* __restrict__ sb, uint32_t wshift, uint32_t g) {
    const uint32_t word = sb[(g * kGateHashK) >> wshift];
    return shf_wrap_r(word, __umulhi(g, kGateHashB)) & shf_wrap_r(word, __umulhi(g, kGateHashC));
}

// level 2: exact table in global memory, narrow slots {gram, unit mask}: the unit mask of the gram, or 0
__device__ __forceinline__ uint32_t gate_l2(const uint2* __restrict__ slots, uint32_t kt, uint32_t g) {
    const uint32_t tm = (1u << kt) - 1u;
    uint32_t s = (g * kGateHash2) >> (32u - kt);
    for (;;) {
        const uint2 e = __ldg(slots + s);
        if (e.y == 0u) return 0u;
        if (e.x == g) return e.y;
        s = (s + 1u) & tm;
    }
}
// ... wide slots {gram, unit mask, first literal candidate, their number} (rule sets that confirm literals): the slot, or zeros
__device__ __forceinline__ uint4 gate_l2(const uint4* __restrict__ slots, uint32_t kt, uint32_t g) {
    const uint32_t tm = (1u << kt) - 1u;
    uint32_t s = (g * kGateHash2) >> (32u - kt);
    for (;;) {
        const uint4 e = __ldg(slots + s);
        if (e.y == 0u && e.w == 0u) return make_uint4(0, 0, 0, 0);
        if (e.x == g) return e;
        s = (s + 1u) & tm;
    }
}

// does literal `d` occur at column position `at` of the field [s, e)?  (GateTables::lit_matches on the device)
// Last byte first: the announcing gram already matched around the literal's start, the bytes behind it tell candidates apart.
__device__ __forceinline__ bool lit_matches(const GateField& F, uint32_t off, uint32_t len, uint32_t flags, uint64_t ci_mask, uint32_t s, uint32_t e,
                                            int64_t at) {
    if (at < (int64_t)s || at + len > (int64_t)e) return false;
    if ((flags & 1u) && at != (int64_t)s) return false;
    if ((flags & 2u) && at + len != (int64_t)e) return false;
    const uint8_t* col = F.col + at;
    const uint8_t* lit = F.lit_bytes + off;
    for (uint32_t k = len; k-- > 0u;) {
        const uint32_t b = __ldg(col + k), want = __ldg(lit + k);
        if (b != want && !(((ci_mask >> k) & 1ull) && (b ^ 0x20u) == want)) return false;
    }
    return true;
}

__device__ __forceinline__ uint32_t ld_nc_u32(const uint8_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(kGateThreads, 1) waf_gate_kernel(const __grid_constant__ GateParams gp) {
    extern __shared__ __align__(128) uint8_t gsm[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t warps_total = gridDim.x * (kGateThreads / 32), warp_global = blockIdx.x * (kGateThreads / 32) + (tid >> 5);
    uint8_t* const blooms = gsm + kGateCtrBytes;

    // every gated field's level-1 bitmap is staged once
    for (uint32_t fi = 0; fi < gp.n_fields; ++fi) {
        const GateField& F = gp.f[fi];
        uint4* d1 = reinterpret_cast<uint4*>(blooms + F.bloom_off);
        const uint4* s1 = reinterpret_cast<const uint4*>(F.b1);
        const uint32_t q = (1u << (F.k1 - 3u)) / 16u;
        for (uint32_t i = tid; i < q; i += kGateThreads) d1[i] = __ldg(s1 + i);
    }
    __syncthreads();

    for (uint32_t fi = 0; fi < gp.n_fields; ++fi) {
        const GateField& F = gp.f[fi];
        const uint32_t* const t1 = reinterpret_cast<const uint32_t*>(blooms + F.bloom_off);
        const uint32_t wmask = 32u - (F.k1 - 5u);   // the Bloom word = the top k1 - 5 bits of the low product
        const uint8_t* col = F.col;
        // the batch's bytes of this column: [off[0], off[n]) (a batch may be a window of a longer column); blocks of 512
        // bytes (one word of the hit bitmap) from the one holding off[0]; the column is readable up to
        // round_up(off[n], 32) (pgw_strcol contract)
        const uint32_t first = __ldg(F.off) & ~511u, total = __ldg(F.off + gp.n);
        const uint32_t limit = (total + 15u) & ~15u;

        // a warp walks blocks of 512 bytes, warps_total blocks apart; loads run two blocks ahead
        uint32_t pos = first + warp_global * 512u + lane * 16u;
        const uint32_t stride = warps_total * 512u;
        auto load16 = [&](uint32_t p) { return p < limit ? ld_nc_v4(col + p) : make_uint4(0, 0, 0, 0); };
        // the word after lane 31's chunk belongs to another warp's block: lane 31 fetches it itself
        auto load_la = [&](uint32_t p) { return (lane == 31u && p + 16u < limit) ? ld_nc_u32(col + p + 16u) : 0u; };
        if (pos - lane * 16u >= limit) continue;   // warp-uniform: nothing for this warp in this column
        uint4 cur = load16(pos), nx1 = make_uint4(0, 0, 0, 0), nx2;
        uint32_t la_cur = load_la(pos), la_nx1 = 0u, la_nx2;
        // positions past 2^32 cannot occur: a column holds less than 4 GiB (pgw_strcol offsets are 32-bit)
        const bool more1 = (uint64_t)pos - lane * 16u + stride < limit;
        if (more1) { nx1 = load16(pos + stride); la_nx1 = load_la(pos + stride); }
        for (;;) {
            const uint64_t wb2 = (uint64_t)pos - lane * 16u + 2ull * stride;   // warp-uniform
            const bool have2 = wb2 < limit;
            nx2 = make_uint4(0, 0, 0, 0);
            la_nx2 = 0u;
            if (have2) { nx2 = load16(pos + 2u * stride); la_nx2 = load_la(pos + 2u * stride); }
            uint32_t la = __shfl_down_sync(FULL, cur.x, 1);
            if (lane == 31u) la = la_cur;
            // fold once per word, then the eight windows at byte offsets 0, 2, .., 14
            const uint32_t f0 = cur.x & kGateFoldMask, f1 = cur.y & kGateFoldMask, f2 = cur.z & kGateFoldMask, f3 = cur.w & kGateFoldMask,
                           f4 = la & kGateFoldMask;
            uint32_t acc = gate_l1(t1, wmask, f0);
            acc |= gate_l1(t1, wmask, __funnelshift_r(f0, f1, 16));
            acc |= gate_l1(t1, wmask, f1);
            acc |= gate_l1(t1, wmask, __funnelshift_r(f1, f2, 16));
            acc |= gate_l1(t1, wmask, f2);
            acc |= gate_l1(t1, wmask, __funnelshift_r(f2, f3, 16));
            acc |= gate_l1(t1, wmask, f3);
            acc |= gate_l1(t1, wmask, __funnelshift_r(f3, f4, 16));
            // one word of the hit bitmap per warp and iteration: bit = lane = chunk (pos >> 4) & 31
            const uint32_t hits = __ballot_sync(FULL, (acc & 1u) && pos < limit);
            if (lane == 0) F.bitmap[pos >> 9] = hits;
            if ((uint64_t)pos - lane * 16u + stride >= limit) break;   // warp-uniform
            pos += stride;
            cur = nx1;
            la_cur = la_nx1;
            nx1 = nx2;
            la_nx1 = la_nx2;
        }
    }
}

// Block-wide append of the threads with `has` to a list: one scan, one atomicAdd; returns this thread's slot.
// Every thread of the block must call it (barriers inside).
__device__ __forceinline__ uint32_t block_append_slot(bool has, uint32_t* counter, uint32_t* s_warp, uint32_t* s_base) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, FULL = 0xFFFFFFFFu;
    const uint32_t bal = __ballot_sync(FULL, has);
    if (lane == 0) s_warp[warp] = (uint32_t)__popc(bal);
    __syncthreads();
    if (warp == 0) {
        const uint32_t c = lane < (blockDim.x >> 5) ? s_warp[lane] : 0u;
        uint32_t x = c;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, o);
            if ((int)lane >= o) x += y;
        }
        s_warp[lane] = x - c;   // exclusive prefix of the warp counts
        if (lane == 31u) *s_base = x ? atomicAdd(counter, x) : 0u;
    }
    __syncthreads();
    const uint32_t k = *s_base + s_warp[warp] + (uint32_t)__popc(bal & ((1u << lane) - 1u));
    __syncthreads();   // s_warp / s_base may be reused by the caller's next list
    return k;
}

// chunks (16 bytes, index = column position >> 4) holding an even-aligned window [j, j + 4) that overlaps the field [s, e)
__device__ __forceinline__ void field_chunks(uint32_t s, uint32_t e, uint32_t* c_lo, uint32_t* c_hi) {
    uint32_t j0 = s >= 3u ? s - 3u : 0u;   // first window with j + 4 > s ...
    j0 += j0 & 1u;                         // ... at an even position
    *c_lo = j0 >> 4;
    *c_hi = ((e - 1u) & ~1u) >> 4;         // last even j < e
}

constexpr uint32_t kListThreads = 1024;
constexpr uint32_t kResolveThreads = 256;   // the resolve kernel has no block-wide step: small CTAs, many of them

// Requests whose field overlaps a chunk with a level-1 hit.  One thread per request; the (up to three) fields' lists are
// appended with ONE block-wide scan: the per-field counts ride in 10-bit lanes of one word.
__global__ void __launch_bounds__(kListThreads) waf_gate_maybe_kernel(const __grid_constant__ GateParams gp) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base[kMaxGateFields];
    const uint32_t n = gp.n, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, FULL = 0xFFFFFFFFu;
    static_assert(kMaxGateFields <= 3 && kListThreads <= 1024, "three 10-bit (+carry) counters per word");
    for (uint32_t blk = blockIdx.x * kListThreads; blk < n; blk += gridDim.x * kListThreads) {
        const uint32_t r = blk + threadIdx.x;
        uint32_t has = 0;   // bit fi: the request goes to field fi's list
        if (r < n)
            for (uint32_t fi = 0; fi < gp.n_fields; ++fi) {
                const GateField& F = gp.f[fi];
                const uint32_t s = __ldg(F.off + r), e = __ldg(F.off + r + 1u);
                if (e <= s) continue;
                uint32_t c_lo, c_hi;
                field_chunks(s, e, &c_lo, &c_hi);
                bool any = false;
                for (uint32_t wi = c_lo >> 5; wi <= (c_hi >> 5) && !any; ++wi) {
                    uint32_t bits = F.bitmap[wi];
                    if (wi == (c_lo >> 5)) bits &= 0xFFFFFFFFu << (c_lo & 31u);
                    if (wi == (c_hi >> 5)) bits &= 0xFFFFFFFFu >> (31u - (c_hi & 31u));
                    any = bits != 0u;
                }
                if (any) has |= 1u << fi;
            }
        // per-field ballots; counts packed 11 bits apart (a block appends at most 1024 per field)
        const uint32_t b0 = __ballot_sync(FULL, has & 1u), b1 = __ballot_sync(FULL, has & 2u), b2 = __ballot_sync(FULL, has & 4u);
        if (lane == 0) s_warp[warp] = (uint32_t)__popc(b0) | ((uint32_t)__popc(b1) << 11) | ((uint32_t)__popc(b2) << 22);
        __syncthreads();
        if (warp == 0) {
            const uint32_t c = s_warp[lane];
            uint32_t x = c;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, x, o);
                if ((int)lane >= o) x += y;
            }
            s_warp[lane] = x - c;   // exclusive prefixes of the warp counts, still packed
            if (lane == 31u)
                for (uint32_t fi = 0; fi < gp.n_fields; ++fi) {
                    // exclusive prefix + own count: the top lane has 10 bits, a full block (1024) would not fit as one number
                    const uint32_t tot = (((x - c) >> (11u * fi)) & 0x7FFu) + ((c >> (11u * fi)) & 0x7FFu);
                    s_base[fi] = tot ? atomicAdd(gp.f[fi].maybe_count, tot) : 0u;
                }
        }
        __syncthreads();
        const uint32_t pre = s_warp[warp], lt = (1u << lane) - 1u;
        if (has & 1u) gp.f[0].maybe_idx[s_base[0] + (pre & 0x7FFu) + (uint32_t)__popc(b0 & lt)] = r;
        if (has & 2u) gp.f[1].maybe_idx[s_base[1] + ((pre >> 11) & 0x7FFu) + (uint32_t)__popc(b1 & lt)] = r;
        if (has & 4u) gp.f[2].maybe_idx[s_base[2] + ((pre >> 22) & 0x7FFu) + (uint32_t)__popc(b2 & lt)] = r;
        __syncthreads();   // s_warp / s_base are reused by the next block of requests
    }
}

// The listed requests' hit chunks against the exact gram table (narrow slots).  grid.y = gated fields; one thread per listed request.
__global__ void __launch_bounds__(kListThreads) waf_gate_resolve_kernel(const __grid_constant__ GateParams gp) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    const GateField& F = gp.f[blockIdx.y];
    const uint32_t count = *F.maybe_count;
    const uint32_t total = __ldg(F.off + gp.n);
    const uint32_t limit = (total + 15u) & ~15u;
    for (uint32_t blk = blockIdx.x * kListThreads; blk < count; blk += gridDim.x * kListThreads) {
        const uint32_t k = blk + threadIdx.x;
        uint32_t r = 0, s = 0, e = 0, mask = 0;
        if (k < count) {
            r = F.maybe_idx[k];
            s = __ldg(F.off + r);
            e = __ldg(F.off + r + 1u);
            uint32_t c_lo, c_hi;
            field_chunks(s, e, &c_lo, &c_hi);
            for (uint32_t wi = c_lo >> 5; wi <= (c_hi >> 5); ++wi) {
                uint32_t bits = F.bitmap[wi];
                if (wi == (c_lo >> 5)) bits &= 0xFFFFFFFFu << (c_lo & 31u);
                if (wi == (c_hi >> 5)) bits &= 0xFFFFFFFFu >> (31u - (c_hi & 31u));
                while (bits) {
                    const uint32_t pos = (wi * 32u + (uint32_t)__ffs(bits) - 1u) << 4;
                    bits &= bits - 1u;
                    const uint4 c = ld_nc_v4(F.col + pos);
                    const uint32_t la = pos + 16u < limit ? ld_nc_u32(F.col + pos + 16u) : 0u;
                    const uint32_t f0 = c.x & kGateFoldMask, f1 = c.y & kGateFoldMask, f2 = c.z & kGateFoldMask, f3 = c.w & kGateFoldMask,
                                   f4 = la & kGateFoldMask;
                    uint32_t g[8];
                    g[0] = f0; g[1] = __funnelshift_r(f0, f1, 16); g[2] = f1; g[3] = __funnelshift_r(f1, f2, 16);
                    g[4] = f2; g[5] = __funnelshift_r(f2, f3, 16); g[6] = f3; g[7] = __funnelshift_r(f3, f4, 16);
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        const uint32_t j = pos + 2u * (uint32_t)w;
                        // the window must overlap the field (and lie inside the batch's bytes: j < e <= total)
                        if (j + 4u > s && j < e) mask |= gate_l2(reinterpret_cast<const uint2*>(F.slots), F.kt, g[w]);
                    }
                }
            }
        }
        const bool has = mask != 0u;
        const uint32_t q = block_append_slot(has, F.cand_count, s_warp, &s_base);
        if (has) {
            F.cand_idx[q] = r;
            F.cand_start[q] = s;
            F.cand_end[q] = e;
            F.cand_mask[q] = mask;
        }
    }
}

// The same for rule sets that confirm literals (wide slots, gate.hpp): besides collecting unit masks, a gram's literal candidates
// are compared in place and their atoms fired here.  Every WARP works on its own -- no block-wide barrier, so a lane with many
// windows or candidates holds up 31 neighbours at most (with the block-wide append of the kernel above one such lane in 1 024
// stalled the whole CTA: measured 5.9 instead of 2.1 ms per 10 M requests for the gate group); candidates are appended warp by
// warp (one atomicAdd per warp that has any): the list is in request order within 32 entries, which is all the scan's pools need.
__global__ void __launch_bounds__(kResolveThreads) waf_gate_resolve_lit_kernel(const __grid_constant__ GateParams gp) {
    const GateField& F = gp.f[blockIdx.y];
    const uint32_t count = *F.maybe_count;
    const uint32_t total = __ldg(F.off + gp.n);
    const uint32_t limit = (total + 15u) & ~15u;
    const uint32_t lane = threadIdx.x & 31u, FULL = 0xFFFFFFFFu;
    const uint32_t gwarp = (blockIdx.x * kResolveThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kResolveThreads) >> 5;
    const uint4* lits4 = reinterpret_cast<const uint4*>(F.lits);
    for (uint32_t base = gwarp * 32u; base < count; base += nwarps * 32u) {
        const uint32_t k = base + lane;
        uint32_t r = 0, s = 0, e = 0, mask = 0;
        if (k < count) {
            r = F.maybe_idx[k];
            s = __ldg(F.off + r);
            e = __ldg(F.off + r + 1u);
            uint32_t c_lo, c_hi;
            field_chunks(s, e, &c_lo, &c_hi);
            for (uint32_t wi = c_lo >> 5; wi <= (c_hi >> 5); ++wi) {
                uint32_t bits = F.bitmap[wi];
                if (wi == (c_lo >> 5)) bits &= 0xFFFFFFFFu << (c_lo & 31u);
                if (wi == (c_hi >> 5)) bits &= 0xFFFFFFFFu >> (31u - (c_hi & 31u));
                while (bits) {
                    const uint32_t pos = (wi * 32u + (uint32_t)__ffs(bits) - 1u) << 4;
                    bits &= bits - 1u;
                    const uint4 c = ld_nc_v4(F.col + pos);
                    const uint32_t la = pos + 16u < limit ? ld_nc_u32(F.col + pos + 16u) : 0u;
                    const uint32_t f0 = c.x & kGateFoldMask, f1 = c.y & kGateFoldMask, f2 = c.z & kGateFoldMask, f3 = c.w & kGateFoldMask,
                                   f4 = la & kGateFoldMask;
                    uint32_t g[8];
                    g[0] = f0; g[1] = __funnelshift_r(f0, f1, 16); g[2] = f1; g[3] = __funnelshift_r(f1, f2, 16);
                    g[4] = f2; g[5] = __funnelshift_r(f2, f3, 16); g[6] = f3; g[7] = __funnelshift_r(f3, f4, 16);
                    // the eight probes are independent loads: issue them together, then look at what came back
                    uint4 en[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        const uint32_t j = pos + 2u * (uint32_t)w;
                        // the window must overlap the field (and lie inside the batch's bytes: j < e <= total)
                        en[w] = (j + 4u > s && j < e) ? gate_l2(reinterpret_cast<const uint4*>(F.slots), F.kt, g[w]) : make_uint4(0, 0, 0, 0);
                        mask |= en[w].y;
                    }
                    // finite-string patterns announced by a gram (one, for an atom whose grams are its own: compile.cpp): compared in place, their atoms fired here
#pragma unroll 1
                    for (int w = 0; w < 8; ++w) {
                        const uint32_t nc = en[w].w;
                        if (nc == 0u) continue;
                        const uint32_t j = pos + 2u * (uint32_t)w;
                        for (uint32_t ci = 0; ci < nc; ++ci) {
                            const uint32_t cd = __ldg(F.lit_cand + en[w].z + ci);
                            const uint4 d0 = __ldg(lits4 + 2u * (cd >> 2)), d1 = __ldg(lits4 + 2u * (cd >> 2) + 1u);   // LitDesc: off, len | flags << 16, atom, pad; ci_mask
                            const uint64_t cim = (uint64_t)d1.x | ((uint64_t)d1.y << 32);
                            if (lit_matches(F, d0.x, d0.y & 0xFFFFu, d0.y >> 16, cim, s, e, (int64_t)j + (int64_t)(cd & 3u) - 1))
                                fire_atom(Sink{gp.rows + (size_t)r * gp.atom_words, gp.info + 2u * (size_t)r}, d0.z);
                        }
                    }
                }
            }
        }
        const bool has = mask != 0u;
        const uint32_t bal = __ballot_sync(FULL, has);
        if (bal) {
            uint32_t q0 = 0;
            if (lane == 0) q0 = atomicAdd(F.cand_count, (uint32_t)__popc(bal));
            q0 = __shfl_sync(FULL, q0, 0);
            if (has) {
                const uint32_t q = q0 + (uint32_t)__popc(bal & ((1u << lane) - 1u));
                F.cand_idx[q] = r;
                F.cand_start[q] = s;
                F.cand_end[q] = e;
                F.cand_mask[q] = mask;
            }
        }
    }
}
