// Part of kernels.cu (included inside namespace pgw { namespace { ... } }).
//
// Candidate gate (gate.hpp): the literal prefilter in front of the DFA scan -- what Rust `regex` does with memchr /
// Teddy in front of its automata (reference call path pingoo/rules.rs:38 -> bel -> regex 1.12.2), restated for a
// batch.  The gate is position-local (no automaton state), so a field column is read as ONE flat stream: a warp
// takes a tile of 32 consecutive requests -- a contiguous byte range of the column -- and walks it 512 bytes per
// iteration, every lane one coalesced 16-byte load.  Each lane tests the eight even-aligned 4-byte windows that start
// in its 16 bytes against the level-1 blocked Bloom filter in shared memory (fold case, multiplicative hash, one
// word, two bits).  A lane that saw a level-1 hit (a few 1e-4 of the windows) re-tests its windows one by one, looks
// the survivors up in the exact gram table in global memory (gram -> mask of scan units) and ORs the mask into every
// request of the tile the window overlaps (binary search over the tile's 33 offsets in shared memory).  A tile's
// candidates are compacted with one ballot and appended to the field's candidate list (request, start, end, unit
// mask) with one atomicAdd.
constexpr int kGateThreads = 1024;
constexpr uint32_t kGateWarpSmem = 68 * 4;  // per warp: 33 offsets of the tile, 32 unit masks, pad (multiple of 16 bytes)

__device__ __forceinline__ uint32_t gate_fold_dev(uint32_t g) { return g | ((g & 0x40404040u) >> 1); }

__device__ __forceinline__ uint32_t shf_wrap_r(uint32_t x, uint32_t n) { return __funnelshift_r(x, x, n); }  // rotate: amount taken mod 32

// level 1: both bits of the window's Bloom word set?  (bit 0 of the result)
__device__ __forceinline__ uint32_t gate_l1(uint32_t tbl, uint32_t g, uint32_t sh) {
    const uint32_t h = g * kGateHash1;
    const uint32_t word = lds_u32(tbl + ((h >> (sh + 3u)) & ~3u));
    return shf_wrap_r(word, h >> sh) & shf_wrap_r(word, h >> (sh - 5u));
}

// level 2: exact table in global memory; unit mask of the gram or 0
__device__ __noinline__ uint32_t gate_l2(const uint2* __restrict__ slots, uint32_t kt, uint32_t g) {
    const uint32_t tm = (1u << kt) - 1u;
    uint32_t s = (g * kGateHash2) >> (32u - kt);
    for (;;) {
        const uint2 e = __ldg(slots + s);
        if (e.y == 0u) return 0u;
        if (e.x == g) return e.y;
        s = (s + 1u) & tm;
    }
}

// OR `m` into the mask of every request of the tile whose field overlaps the window [j, j+4)
__device__ __noinline__ void gate_mark(uint32_t a_offs, uint32_t a_mask, uint32_t j, uint32_t m) {
    // smallest r in [0, 31] with end_r = offs[r + 1] > j  (offs[32] = B > j holds for every window of the tile)
    uint32_t lo = 0, hi = 31;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (lds_u32_v(a_offs + 4u * (mid + 1u)) > j) hi = mid;
        else lo = mid + 1u;
    }
    for (uint32_t r = lo; r < 32u; ++r) {
        const uint32_t s = lds_u32_v(a_offs + 4u * r), e = lds_u32_v(a_offs + 4u * (r + 1u));
        if (s >= j + 4u) break;
        if (e > s && e > j) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a_mask + 4u * r), "r"(m) : "memory");
    }
}

// One request's field walked on a small early-exit DFA whose whole table is in shared memory at `img` (class map,
// rows, acc1, end1: the unit image of compile.hpp): start-anchored patterns are decided within the first few bytes.
__device__ __noinline__ void prefix_walk(const KParams& p, const UnitDesc& ud, uint32_t img, const uint8_t* __restrict__ col, uint32_t s, uint32_t e,
                                         uint32_t ridx) {
    const uint32_t C2 = 2u * ud.n_classes, acclo = ud.acc_lo, abs0 = ud.abs0, abs1 = ud.abs1;
    const uint32_t hot = img + ud.hot_off;
    uint32_t st = ud.start_state, latch = 0u;
    const Sink sink = sink_of(p, ridx);
    // start-anchored patterns die (or are decided) within a few bytes: a plain byte loop that stops at an absorbing state
#pragma unroll 1
    for (uint32_t pos = s; pos < e; ++pos) {
        const uint32_t cls = lds_u8(img + (uint32_t)__ldg(col + pos));
        st = lds_u16(hot + st * C2 + 2u * cls);
        if (st >= acclo) {
            const uint32_t a1 = lds_u16(img + ud.acc1_off + 2u * (st - acclo));
            if (a1 != 0xFFFFu) fire_atom(sink, a1);
            else fs_fire_list(p.acc_idx, p.acc_events, ud.acc_base + st - acclo, sink, &latch);
        }
        if (st == abs0 || st == abs1) break;  // absorbing: nothing can change any more
    }
    const uint32_t e1 = lds_u16(img + ud.end1_off + 2u * st);
    if (e1 != 0xFFFEu) {
        if (e1 != 0xFFFFu) fire_atom(sink, e1);
        else if (ud.end_any) fs_fire_list(p.end_idx, p.end_events, ud.end_base + st, sink, &latch);
    }
}

__global__ void __launch_bounds__(kGateThreads, 1) waf_gate_kernel(const __grid_constant__ GateParams gp, const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) uint8_t gsm[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t warps_total = gridDim.x * (kGateThreads / 32), warp_global = blockIdx.x * (kGateThreads / 32) + (tid >> 5);
    const uint32_t n = gp.n;
    const uint32_t n_tiles = (n + 31u) / 32u;
    const uint32_t a_offs = smem_u32(gsm) + (tid >> 5) * kGateWarpSmem, a_mask = a_offs + 33u * 4u;
    uint8_t* const images = gsm + (kGateThreads / 32) * kGateWarpSmem;  // prefix-unit images (256-byte aligned pieces), then the Bloom bitmap
    uint8_t* const bloom = images + gp.image_area;

    for (uint32_t fi = 0; fi < gp.n_fields; ++fi) {
        const GateField& F = gp.f[fi];
        const bool gated = F.b1 != nullptr;
        const uint32_t words1 = gated ? 1u << (F.k1 - 5u) : 0u;
        __syncthreads();  // everybody is done with the previous field's tables
        {
            uint4* d1 = reinterpret_cast<uint4*>(bloom);
            const uint4* s1 = reinterpret_cast<const uint4*>(F.b1);
            for (uint32_t i = tid; i < words1 / 4u; i += kGateThreads) d1[i] = __ldg(s1 + i);
            for (uint32_t k = 0; k < F.n_prefix; ++k) {
                uint4* di = reinterpret_cast<uint4*>(images + F.prefix_img[k]);
                const uint4* si = reinterpret_cast<const uint4*>(p.images + F.prefix[k].img_off);
                for (uint32_t i = tid; i < F.prefix[k].img_bytes / 16u; i += kGateThreads) di[i] = __ldg(si + i);
            }
        }
        __syncthreads();
        const uint32_t t1 = smem_u32(bloom);
        const uint32_t sh = 32u - F.k1;
        const uint32_t total = __ldg(F.off + n);
        const uint32_t limit = (total + 15u) & ~15u;  // the column is readable up to here (pgw_strcol contract: round_up(.., 32))
        const uint8_t* col = F.col;

        for (uint32_t tile = warp_global; tile < n_tiles; tile += warps_total) {
            const uint32_t r = tile * 32u + lane;
            const uint32_t s_l = __ldg(F.off + min(r, n)), e_l = __ldg(F.off + min(r + 1u, n));
            const uint32_t A = __shfl_sync(FULL, s_l, 0), B = __shfl_sync(FULL, e_l, 31);
            if (A == B) continue;  // 32 empty fields
            if (gated) {
            __syncwarp();
            sts_u32(a_offs + 4u * lane, s_l);
            sts_u32(a_mask + 4u * lane, 0u);
            if (lane == 31u) sts_u32(a_offs + 4u * 32u, e_l);
            __syncwarp();
            // windows [j, j+4) with j even, j + 4 > A, j < B: chunks from the one holding A - 3 on
            uint32_t pos = ((A >= 3u ? A - 3u : 0u) & ~15u) + lane * 16u;
            uint4 cur = make_uint4(0, 0, 0, 0), nx1 = make_uint4(0, 0, 0, 0), nx2 = make_uint4(0, 0, 0, 0);
            if (pos < limit) cur = ld_nc_v4(col + pos);
            if (pos + 512u < limit && pos + 512u - lane * 16u < B) nx1 = ld_nc_v4(col + pos + 512u);
            for (;; pos += 512u) {
                const uint32_t wbase = pos - lane * 16u;  // position of lane 0's chunk: warp-uniform
                if (wbase >= B) break;
                // two iterations ahead (zeros past the readable end of the column or past the tile)
                nx2 = make_uint4(0, 0, 0, 0);
                if (pos + 1024u < limit && wbase + 1024u < B) nx2 = ld_nc_v4(col + pos + 1024u);
                // the word after this lane's 16 bytes: the next lane's first word, for lane 31 the first word of the next iteration
                uint32_t la = __shfl_down_sync(FULL, cur.x, 1);
                const uint32_t la31 = __shfl_sync(FULL, nx1.x, 0);
                if (lane == 31u) la = la31;
                // fold once per word, then the eight windows at byte offsets 0, 2, .., 14
                const uint32_t f0 = gate_fold_dev(cur.x), f1 = gate_fold_dev(cur.y), f2 = gate_fold_dev(cur.z), f3 = gate_fold_dev(cur.w),
                               f4 = gate_fold_dev(la);
                uint32_t g[8];
                g[0] = f0;
                g[1] = __funnelshift_r(f0, f1, 16);
                g[2] = f1;
                g[3] = __funnelshift_r(f1, f2, 16);
                g[4] = f2;
                g[5] = __funnelshift_r(f2, f3, 16);
                g[6] = f3;
                g[7] = __funnelshift_r(f3, f4, 16);
                uint32_t acc = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc |= gate_l1(t1, g[i], sh);
                if (acc & 1u) {
                    // rare (a few 1e-4 of the windows): which windows, exact table, mark the requests they overlap
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (!(gate_l1(t1, g[i], sh) & 1u)) continue;
                        const uint32_t j = pos + 2u * i;
                        if (!(j < B && j + 4u > A)) continue;  // windows outside the tile's byte range belong to the neighbouring tiles
                        const uint32_t m = gate_l2(reinterpret_cast<const uint2*>(F.slots), F.kt, g[i]);
                        if (m) gate_mark(a_offs, a_mask, j, m);
                    }
                }
                cur = nx1;
                nx1 = nx2;
            }
            __syncwarp();
            const uint32_t mine = r < n ? lds_u32_v(a_mask + 4u * lane) : 0u;
            const uint32_t cm = __ballot_sync(FULL, mine != 0u);
            if (cm) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(F.cand_count, (uint32_t)__popc(cm));
                base = __shfl_sync(FULL, base, 0);
                if (mine) {
                    const uint32_t k = base + (uint32_t)__popc(cm & ((1u << lane) - 1u));
                    F.cand_idx[k] = r;
                    F.cand_start[k] = s_l;
                    F.cand_end[k] = e_l;
                    F.cand_mask[k] = mine;
                }
            }
            }
            // the field's small early-exit units: one lane per request, the bytes are in the cache
            if (r < n && e_l > s_l)
                for (uint32_t k = 0; k < F.n_prefix; ++k) prefix_walk(p, F.prefix[k], smem_u32(images) + F.prefix_img[k], col, s_l, e_l, r);
        }
    }
}
