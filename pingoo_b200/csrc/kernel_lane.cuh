// Part of kernels.cu (included inside namespace pgw { namespace { ... } }, one translation unit: device functions are
// not linked across files).  The request-major "lane" path (PGW_KERNEL=lane; fall-back beyond kMaxConstUnits scan units).

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

