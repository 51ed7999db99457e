// Part of kernels.cu (included inside namespace pgw { namespace { ... } }).
//
// Candidate gate (gate.hpp): the literal prefilter in front of the DFA scan -- what Rust `regex` does with memchr /
// Teddy in front of its automata (reference call path pingoo/rules.rs:38 -> bel -> regex 1.12.2), restated for a
// batch.  The gate is position-local (no automaton state), so a field column is read as ONE flat stream: a warp
// takes a tile of 32 consecutive requests -- a contiguous byte range of the column -- and walks it 512 bytes per
// iteration, every lane one coalesced 16-byte load.  Each lane tests the eight even-aligned 4-byte windows that start
// in its 16 bytes: fold case, multiplicative hash, one bit of the first bitmap in shared memory.  Windows that pass
// (about 1 %) are tested against the second bitmap (independent hash); windows that pass both mark every request of
// the tile they overlap.  A tile's 32 requests are exactly one word of the candidate bitmap, so a tile's candidates
// are compacted with one ballot and appended to the field's candidate list (request, start, end) with one atomicAdd.
constexpr int kGateThreads = 1024;

__device__ __forceinline__ uint32_t gate_fold_dev(uint32_t g) { return g | ((g & 0x40404040u) >> 1); }

// 1 if the window's bit is set in the bitmap at shared address `tbl` (2^k bits, k = 32 - sh)
__device__ __forceinline__ uint32_t gate_probe(uint32_t tbl, uint32_t g, uint32_t mult, uint32_t sh) {
    const uint32_t h = g * mult;
    const uint32_t word = lds_u32(tbl + ((h >> (sh + 3u)) & ~3u));
    return __funnelshift_r(word, word, h >> sh) & 1u;  // wrap mode: the shift uses the low five bits of (h >> sh)
}

__global__ void __launch_bounds__(kGateThreads, 1) waf_gate_kernel(const __grid_constant__ GateParams gp) {
    extern __shared__ __align__(128) uint8_t gsm[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t warps_total = gridDim.x * (kGateThreads / 32), warp_global = blockIdx.x * (kGateThreads / 32) + (tid >> 5);
    const uint32_t n = gp.n;
    const uint32_t n_tiles = (n + 31u) / 32u;

    for (uint32_t fi = 0; fi < gp.n_fields; ++fi) {
        const GateField& F = gp.f[fi];
        const uint32_t words1 = 1u << (F.k1 - 5u), words2 = 1u << (F.k2 - 5u);
        __syncthreads();  // everybody is done with the previous field's bitmaps
        {
            uint4* d1 = reinterpret_cast<uint4*>(gsm);
            const uint4* s1 = reinterpret_cast<const uint4*>(F.b1);
            for (uint32_t i = tid; i < words1 / 4u; i += kGateThreads) d1[i] = __ldg(s1 + i);
            uint4* d2 = reinterpret_cast<uint4*>(gsm + words1 * 4u);
            const uint4* s2 = reinterpret_cast<const uint4*>(F.b2);
            for (uint32_t i = tid; i < words2 / 4u; i += kGateThreads) d2[i] = __ldg(s2 + i);
        }
        __syncthreads();
        const uint32_t t1 = smem_u32(gsm), t2 = t1 + words1 * 4u;
        const uint32_t sh1 = 32u - F.k1, sh2 = 32u - F.k2;
        const uint32_t total = __ldg(F.off + n);
        const uint32_t limit = (total + 15u) & ~15u;  // the column is readable up to here (pgw_strcol contract: round_up(.., 32))
        const uint8_t* col = F.col;

        for (uint32_t tile = warp_global; tile < n_tiles; tile += warps_total) {
            const uint32_t r = tile * 32u + lane;
            const uint32_t s_l = __ldg(F.off + min(r, n)), e_l = __ldg(F.off + min(r + 1u, n));
            const uint32_t A = __shfl_sync(FULL, s_l, 0), B = __shfl_sync(FULL, e_l, 31);
            if (A == B) continue;  // 32 empty fields
            // windows [j, j+4) with j even, j + 4 > A, j < B: chunks from the one holding A - 3 on
            uint32_t pos = ((A >= 3u ? A - 3u : 0u) & ~15u) + lane * 16u;
            bool mine = false;
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
                // eight windows at byte offsets 0, 2, .., 14
                uint32_t g[8];
                g[0] = cur.x;
                g[1] = __funnelshift_r(cur.x, cur.y, 16);
                g[2] = cur.y;
                g[3] = __funnelshift_r(cur.y, cur.z, 16);
                g[4] = cur.z;
                g[5] = __funnelshift_r(cur.z, cur.w, 16);
                g[6] = cur.w;
                g[7] = __funnelshift_r(cur.w, la, 16);
                uint32_t hits = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    g[i] = gate_fold_dev(g[i]);
                    hits |= gate_probe(t1, g[i], kGateHash1, sh1) << i;
                }
                if (hits) {
                    // second bitmap, only for windows that passed the first
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (hits & (1u << i))
                            if (!gate_probe(t2, g[i], kGateHash2, sh2)) hits &= ~(1u << i);
                    // windows outside the tile's byte range belong to the neighbouring tiles
                    if (hits) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t j = pos + 2u * i;
                            if (!(j < B && j + 4u > A)) hits &= ~(1u << i);
                        }
                    }
                }
                uint32_t hm = __ballot_sync(FULL, hits != 0u);
                while (hm) {  // rare: mark the requests each hit window overlaps
                    const int src = __ffs(hm) - 1;
                    hm &= hm - 1u;
                    uint32_t f = __shfl_sync(FULL, hits, src);
                    const uint32_t p0 = __shfl_sync(FULL, pos, src);
                    while (f) {
                        const uint32_t j = p0 + 2u * (uint32_t)(__ffs(f) - 1);
                        f &= f - 1u;
                        mine |= (s_l < j + 4u) && (e_l > j) && (e_l > s_l);
                    }
                }
                cur = nx1;
                nx1 = nx2;
            }
            mine = mine && r < n;
            const uint32_t cm = __ballot_sync(FULL, mine);
            if (cm) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(F.cand_count, (uint32_t)__popc(cm));
                base = __shfl_sync(FULL, base, 0);
                if (mine) {
                    const uint32_t k = base + (uint32_t)__popc(cm & ((1u << lane) - 1u));
                    F.cand_idx[k] = r;
                    F.cand_start[k] = s_l;
                    F.cand_end[k] = e_l;
                }
            }
        }
    }
}
