// Part of kernels.cu (included inside namespace pgw { namespace { ... } }).
//
// Candidate gate (gate.hpp): the literal prefilter in front of the DFA scan -- what Rust `regex` does with memchr /
// Teddy in front of its automata (reference call path pingoo/rules.rs:38 -> bel -> regex 1.12.2), restated for a
// batch.  The gate is position-local (no automaton state, no request boundaries), so it is three kernels:
//
//   waf_gate_kernel      streams every gated field column as ONE flat byte range: a warp takes 512 consecutive bytes per
//                        iteration, every lane one coalesced 16-byte load.  Each lane tests the eight even-aligned
//                        4-byte windows that start in its 16 bytes against the field's level-1 blocked Bloom filter in
//                        shared memory (all gated fields' filters are resident together).  A lane that saw a level-1
//                        hit (about 1 % of the chunks on benign traffic) appends the chunk's index to its CTA's segment
//                        of the hit queue -- nothing else happens in the streaming loop.
//   waf_gate_resolve_kernel  one thread per queued chunk: the eight windows again, now against the exact gram table
//                        (gram -> mask of scan units), the requests a hit window overlaps (binary search over the
//                        field's offsets), their unit masks (atomicOr into one word per request) and, for the first
//                        marker of a request, its entry in the field's candidate list.
//   waf_gate_finalize_kernel copies each candidate's final unit mask next to its list entry for the scan kernel.
//
// A hit queue segment that overflows (adversarial input: every chunk hits) sets the field's overflow flag; the resolve
// kernel then lists every request as a candidate of every gated unit of the field -- the ungated behaviour, still exact.
constexpr int kGateThreads = 1024;
constexpr uint32_t kGateCtrBytes = 64;   // front of the shared window: one hit counter per gated field

__device__ __forceinline__ uint32_t shf_wrap_r(uint32_t x, uint32_t n) { return __funnelshift_r(x, x, n); }  // rotate: amount taken mod 32

// level 1: both bits of the window's Bloom word set?  (bit 0 of the result).  FMA pipe: one IMAD + two IMAD.HI; ALU
// pipe: one SHF (word index; the LDS scales it by 4), two SHF.W, and the caller's LOP3 that ORs the eight windows.
__device__ __forceinline__ uint32_t gate_l1(const uint32_t* __restrict__ sb, uint32_t wshift, uint32_t g) {
    const uint32_t word = sb[(g * kGateHashK) >> wshift];
    return shf_wrap_r(word, __umulhi(g, kGateHashB)) & shf_wrap_r(word, __umulhi(g, kGateHashC));
}

// level 2: exact table in global memory; unit mask of the gram or 0
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
    uint32_t* const s_cnt = reinterpret_cast<uint32_t*>(gsm);
    uint8_t* const blooms = gsm + kGateCtrBytes;

    // every gated field's level-1 bitmap is staged once
    if (tid < kMaxGateFields) s_cnt[tid] = 0u;
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
        // the batch's bytes of this column: [off[0], off[n]) (a batch may be a window of a longer column); chunks of 16
        // bytes from the one holding off[0]; the column is readable up to round_up(off[n], 32) (pgw_strcol contract)
        const uint32_t first = __ldg(F.off) & ~15u, total = __ldg(F.off + gp.n);
        const uint32_t limit = (total + 15u) & ~15u;
        uint32_t* const hq = F.hq + (size_t)blockIdx.x * F.hq_cap;
        const uint32_t a_cnt = smem_u32(s_cnt + fi);

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
            if ((acc & 1u) && pos < limit) {
                uint32_t slot;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(a_cnt) : "memory");
                if (slot < F.hq_cap) hq[slot] = pos >> 4;
                else *F.overflow = 1u;
            }
            if ((uint64_t)pos - lane * 16u + stride >= limit) break;   // warp-uniform
            pos += stride;
            cur = nx1;
            la_cur = la_nx1;
            nx1 = nx2;
            la_nx1 = la_nx2;
        }
    }
    __syncthreads();
    if (tid < gp.n_fields) {
        const GateField& F = gp.f[tid];
        F.hq_count[blockIdx.x] = min(s_cnt[tid], F.hq_cap);
    }
}

// One thread per queued chunk.  grid = (segments * kResolveParts, gated fields); block b of a field works on segment
// b % segments, interleaved with the other kResolveParts - 1 blocks of that segment.
constexpr uint32_t kResolveParts = 4, kResolveThreads = 256;

__global__ void __launch_bounds__(kResolveThreads) waf_gate_resolve_kernel(const __grid_constant__ GateParams gp) {
    const GateField& F = gp.f[blockIdx.y];
    const uint32_t n = gp.n, lane = threadIdx.x & 31u;
    const uint32_t fshift = F.mask_shift, fmask = F.mask_bits;   // this field's bits in a request's candidate word
    const uint32_t* __restrict__ off = F.off;

    // appends request r (field bytes [s, e)) to the field's candidate list; called under divergence: the lanes that are
    // here together share one atomicAdd
    auto append = [&](uint32_t r, uint32_t s, uint32_t e) {
        const uint32_t act = __activemask();
        const uint32_t leader = __ffs(act) - 1u;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(F.cand_count, (uint32_t)__popc(act));
        base = __shfl_sync(act, base, leader);
        const uint32_t k = base + (uint32_t)__popc(act & ((1u << lane) - 1u));
        F.cand_idx[k] = r;
        F.cand_start[k] = s;
        F.cand_end[k] = e;
    };

    if (*reinterpret_cast<volatile uint32_t*>(F.overflow)) {
        // the hit queue overflowed: every request is a candidate of every gated unit of the field
        const uint32_t stride = gridDim.x * kResolveThreads;
        for (uint32_t r = blockIdx.x * kResolveThreads + threadIdx.x; r < n; r += stride) {
            atomicOr(gp.reqmask + r, fmask << fshift);
            F.cand_idx[r] = r;
            F.cand_start[r] = __ldg(off + r);
            F.cand_end[r] = __ldg(off + r + 1u);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) *F.cand_count = n;
        return;
    }

    const uint32_t seg = blockIdx.x % gp.n_seg, part = blockIdx.x / gp.n_seg, parts = gridDim.x / gp.n_seg;
    const uint32_t count = F.hq_count[seg];
    const uint32_t* hq = F.hq + (size_t)seg * F.hq_cap;
    const uint32_t first_byte = __ldg(off), total = __ldg(off + n);
    const uint32_t limit = (total + 15u) & ~15u;
    for (uint32_t i = part * kResolveThreads + threadIdx.x; i < count; i += parts * kResolveThreads) {
        const uint32_t pos = hq[i] << 4;
        const uint4 c = ld_nc_v4(F.col + pos);
        const uint32_t la = pos + 16u < limit ? ld_nc_u32(F.col + pos + 16u) : 0u;
        const uint32_t f0 = c.x & kGateFoldMask, f1 = c.y & kGateFoldMask, f2 = c.z & kGateFoldMask, f3 = c.w & kGateFoldMask, f4 = la & kGateFoldMask;
        uint32_t g[8];
        g[0] = f0; g[1] = __funnelshift_r(f0, f1, 16); g[2] = f1; g[3] = __funnelshift_r(f1, f2, 16);
        g[4] = f2; g[5] = __funnelshift_r(f2, f3, 16); g[6] = f3; g[7] = __funnelshift_r(f3, f4, 16);
        uint32_t m[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) m[w] = gate_l2(reinterpret_cast<const uint2*>(F.slots), F.kt, g[w]);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (!m[w]) continue;
            const uint32_t j = pos + 2u * (uint32_t)w;
            if (j >= total || j + 4u <= first_byte) continue;   // the window lies outside the batch's bytes
            // largest r with off[r] <= j (0 when the window starts before the batch's first byte)
            uint32_t lo = 0, hi = n;   // invariant: off[lo] <= j or lo == 0; off[hi] > j
            while (hi - lo > 1u) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(off + mid) <= j) lo = mid;
                else hi = mid;
            }
            for (uint32_t r = lo; r < n; ++r) {
                const uint32_t s = __ldg(off + r);
                if (s >= j + 4u) break;
                const uint32_t e = __ldg(off + r + 1u);
                if (e <= s || e <= j) continue;   // empty field, or the window starts at or after the field's end
                const uint32_t old = atomicOr(gp.reqmask + r, (m[w] & fmask) << fshift);
                if (((old >> fshift) & fmask) == 0u) append(r, s, e);
            }
        }
    }
}

// The candidates' final unit masks, next to their list entries.  grid.y = gated fields.
__global__ void __launch_bounds__(256) waf_gate_finalize_kernel(const __grid_constant__ GateParams gp) {
    const GateField& F = gp.f[blockIdx.y];
    const uint32_t count = *F.cand_count;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x)
        F.cand_mask[k] = (gp.reqmask[F.cand_idx[k]] >> F.mask_shift) & F.mask_bits;
}
