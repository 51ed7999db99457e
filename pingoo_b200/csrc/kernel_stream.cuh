// Part of kernels.cu (included inside namespace pgw { namespace { ... } }, one translation unit: device functions are
// not linked across files).  The "stream" path (PGW_KERNEL=stream).

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

