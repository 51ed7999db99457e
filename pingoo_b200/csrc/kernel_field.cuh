// Part of kernels.cu (included inside namespace pgw { namespace { ... } }, one translation unit: device functions are
// not linked across files).  The DFA scan and the epilogue kernel.

// =====================================================================================================================
// Field scan: unit-major, lane-owned strings.  All warps of a CTA work on ONE scan unit at a time: its shared-memory
// image (class map, hot rows, event tables -- up to the whole shared-memory budget) is staged by TMA bulk copies at the
// start of the unit, so every per-unit parameter is warp-uniform.  Each lane owns one request's field of that unit and
// walks it 16 bytes per iteration (speculative 4-byte word walk on the shared-memory rows).  A unit walks either every
// request (UM_ALL) or the candidates the gate kernel listed for its field (UM_CANDIDATES).  A lane that finishes --
// end of the field, or an absorbing DFA state (UnitDesc::abs0/abs1: nothing can change any more) -- takes the next
// request from a warp pool of 32 claimed entries whose (start, end, request) triples were fetched one pool ahead.
// Atom bits go to bitmaps in global memory (red.or, rare) together with the request's info words;
// waf_epilogue_kernel turns them into verdicts.
// =====================================================================================================================
#ifndef PGW_FS_THREADS
#define PGW_FS_THREADS 1024
#endif
constexpr int kFsThreads = PGW_FS_THREADS;
constexpr uint32_t kFsSlotStride = kFsThreads * 4u;  // per-lane slots: request index, latch register, last fired state, next request
#ifndef PGW_FS_TICKET
#define PGW_FS_TICKET 64
#endif
#ifndef PGW_FS_TICKET_CAND
#define PGW_FS_TICKET_CAND 32u
#endif
constexpr uint32_t kFsTicket = PGW_FS_TICKET;  // entries per atomic claim (two 32-entry pools: 128 and 256 measured worse, tail imbalance)
constexpr uint32_t kFsPoolBytes = 512;  // a claimed pool: 32 field starts, 32 field ends, 32 request indices, 32 unit masks; two buffers per warp
// front of the shared window (mbarrier + skip flag, claim pools, per-lane slots), rounded so that the unit image behind it starts on a 256-byte boundary
constexpr uint32_t kFsFront = (64u + (kFsThreads / 32) * 2u * kFsPoolBytes + 4u * kFsSlotStride + 255u) & ~255u;

// accept events of four hot states of one word (at least one accepting); returns the new `last`
__device__ __noinline__ uint32_t fs_events_word(const KParams& p, uint32_t acclo, uint32_t acc_base, uint32_t acc1addr, uint32_t s01, uint32_t s23,
                                                uint32_t m4, uint32_t last, uint32_t* latch, const Sink& row) {
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
            fire_atom(row, a1);
            last = st;
        } else {
            last = fs_fire_list(p.acc_idx, p.acc_events, acc_base + st - acclo, row, latch) ? st : 0xFFFFFFFFu;
        }
    }
    return last;
}

// one word walked on the full table in global memory (a cold state is involved)
__device__ __noinline__ void fs_slow_word(const KParams& p, uint32_t tbl_off, uint32_t C, uint32_t acclo, uint32_t acc_base, uint32_t clsaddr,
                                          uint32_t w, uint32_t m4, uint32_t* state, uint32_t* last, uint32_t* latch, const Sink& row) {
    const uint16_t* tbl = reinterpret_cast<const uint16_t*>(p.arena + tbl_off);
    uint32_t st = *state, la = *last;
#pragma unroll 1
    for (int bi = 0; bi < 4; ++bi) {
        if (!((m4 >> bi) & 1u)) continue;
        const uint32_t byte = (w >> (8 * bi)) & 0xFFu;
        st = __ldg(tbl + st * C + lds_u8(clsaddr + byte));
        if (st >= acclo && st != la) la = fs_fire_list(p.acc_idx, p.acc_events, acc_base + st - acclo, row, latch) ? st : 0xFFFFFFFFu;
    }
    *state = st;
    *last = la;
}

__global__ void __launch_bounds__(kFsThreads, 1) waf_field_scan_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_img = smem + ((0u - smem_u32(smem)) & 255u);  // the class map (image offset 0) sits on a 256-byte boundary
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // fixed part of the shared window sits in FRONT of the image: mbarrier, pools, slots
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_img);
    const uint32_t a_pool = smem_u32(s_img) + 64u + (tid >> 5) * 2u * kFsPoolBytes;
    const uint32_t a_slot = smem_u32(s_img) + 64u + (kFsThreads / 32) * 2u * kFsPoolBytes + tid * 4u;  // word k at a_slot + k * kFsSlotStride
    uint8_t* const s_image = s_img + kFsFront;
    const uint32_t a_img = smem_u32(s_image);

    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // Plan: every CTA computes the same assignment of CTAs to units from the units' work (entries x mean field length,
    // known on the device only: candidate counts): each unit with work gets one CTA plus a share of the rest in
    // proportion to its work, CTA b's HOME unit follows from the cumulative shares.  A CTA stages its home unit's image
    // once and works there until the unit's entries are all claimed; afterwards it goes round the other units and
    // joins one only if a worthwhile amount of work is left there (or nobody is at home).
    constexpr uint32_t kJoinMin = 4096;
    volatile uint32_t* s_skip = reinterpret_cast<volatile uint32_t*>(s_img + 16);
    volatile uint32_t* s_plan = reinterpret_cast<volatile uint32_t*>(s_img + 20);   // [0] home unit, [1..2] units that have a home CTA
    if (tid == 0) {
        uint64_t work[kMaxConstUnits];
        uint64_t total = 0;
        uint32_t k = 0;
        for (uint32_t u = 0; u < p.n_units; ++u) {
            const UnitDesc& d = p.udesc[u];
            work[u] = 0;
            if (d.mode == UM_PREPASS) continue;
            const uint32_t N = d.mode == UM_CANDIDATES ? __ldg(p.cand_count[d.field]) : p.n;
            if (!N) continue;
            const uint64_t bytes = (uint64_t)(__ldg(p.off[d.field] + p.n) - __ldg(p.off[d.field]));
            work[u] = (uint64_t)N * bytes / p.n + N;
            total += work[u];
            ++k;
        }
        uint32_t home = 0, mask_lo = 0, mask_hi = 0;
        if (k) {
            const uint32_t G = gridDim.x, spare = G > k ? G - k : 0u;
            // shares: one CTA each (while CTAs last), the spare ones in proportion; what rounding leaves goes to the largest unit
            uint32_t given = 0, big = 0;
            uint32_t share[kMaxConstUnits];
            uint32_t seen = 0;
            for (uint32_t u = 0; u < p.n_units; ++u) {
                share[u] = 0;
                if (!work[u]) continue;
                if (seen < G) share[u] = 1u + (uint32_t)((uint64_t)spare * work[u] / total);
                ++seen;
                given += share[u];
                if (work[u] > work[big] || !work[big]) big = u;
            }
            if (given < G) share[big] += G - given;
            uint32_t cum = 0;
            bool placed = false;
            for (uint32_t u = 0; u < p.n_units; ++u) {
                if (!share[u]) continue;
                if (u < 32u) mask_lo |= 1u << u; else mask_hi |= 1u << (u - 32u);
                if (!placed && blockIdx.x < cum + share[u]) { home = u; placed = true; }
                cum += share[u];
            }
            if (!placed) home = blockIdx.x % p.n_units;
        }
        s_plan[0] = home;
        s_plan[1] = mask_lo;
        s_plan[2] = mask_hi;
    }
    __syncthreads();
    const uint32_t u_home = s_plan[0], home_lo = s_plan[1], home_hi = s_plan[2];
    uint32_t staged = 0;
    for (uint32_t uk = 0; uk < p.n_units; ++uk) {
        const uint32_t u = (uk + u_home) % p.n_units;
        const UnitDesc& cu = p.udesc[u];   // constant bank, uniform index
        if (cu.mode == UM_PREPASS) continue;  // walked by the per-request kernel
        const bool cand = cu.mode == UM_CANDIDATES;
        const uint32_t N = cand ? __ldg(p.cand_count[cu.field]) : p.n;
        uint32_t* ctr = p.counters + p.unit_base + u;
        // ---- stage this unit's image: everybody has left the previous unit's tables ----
        __syncthreads();
        if (tid == 0) {
            const uint32_t taken = *reinterpret_cast<volatile uint32_t*>(ctr);
            const uint32_t left = taken >= N ? 0u : N - taken;
            const bool has_home = ((u < 32u ? home_lo >> u : home_hi >> (u - 32u)) & 1u) != 0u;
            const uint32_t skip_unit = (left == 0u || (uk != 0u && has_home && left < kJoinMin)) ? 1u : 0u;
            *s_skip = skip_unit;
            if (!skip_unit) {
                const uint32_t bytes = cu.img_bytes;
                mbar_expect_tx(s_bar, bytes);
                for (uint32_t o = 0; o < bytes; o += 32768u) {
                    const uint32_t nb = bytes - o < 32768u ? bytes - o : 32768u;
                    bulk_g2s(s_image + o, p.images + cu.img_off + o, nb, s_bar);
                }
            }
        }
        __syncthreads();
        if (*s_skip) continue;
        mbar_wait(s_bar, staged & 1u);
        ++staged;
#ifdef PGW_EXP_SCANLOG
        unsigned long long t_unit0 = 0;
        uint32_t n_strings = 0, n_cold = 0, n_iter = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_unit0));
#endif

        const uint32_t C2 = 2u * cu.n_classes, D0 = cu.start_state, trap = cu.hot_states, lim = cu.lim, acclo = cu.acc_lo;
        const uint32_t abs0 = cu.abs0, abs1 = cu.abs1;
        const uint32_t clsaddr = a_img, hotaddr = a_img + cu.hot_off, acc1addr = a_img + cu.acc1_off, end1addr = a_img + cu.end1_off;
        const uint8_t* col = p.col[cu.field];
        const uint32_t* off = p.off[cu.field];
        const uint32_t* c_idx = p.cand_idx[cu.field];
        const uint32_t* c_start = p.cand_start[cu.field];
        const uint32_t* c_end = p.cand_end[cu.field];
        const uint32_t* c_mask = p.cand_mask[cu.field];
        const uint32_t gate_bit = cu.gate_bit;

        // warp pools of 32 claimed entries, double buffered in shared memory: buffer `pb` is being handed out, the other
        // one holds the next claim whose triples are landing through cp.async (no registers, no stall)
        uint32_t pool_next = 0, pool_end = 0, pb = 0, ah_base = 0;
        // claims are pipelined three deep so that no global latency is ever waited for: `ticket` (atomicAdd issued, result
        // not looked at yet) -> `ahead` (triples landing in the spare buffer) -> the pool being handed out
        uint32_t ticket = 0;
        // candidate lists are short (a handful of strings per lane): one pool per claim keeps the tail of a unit short
        const uint32_t ticket_sz = cand ? PGW_FS_TICKET_CAND : kFsTicket;
        bool tk_valid = false, ah_valid = false;
        auto issue_ticket = [&]() {
            if (lane == 0) ticket = atomicAdd(ctr, ticket_sz);  // a ticket covers ticket_sz / 32 consecutive pools
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
                more = min(b + ticket_sz, N);
                issue_ticket();
            }
            ah_base = b;
            const uint32_t dst = a_pool + (pb ^ 1u) * kFsPoolBytes + lane * 4u;
            const uint32_t e = min(b + lane, N - 1u);  // entries past N are never handed out
            if (cand) {
                cp_async4(dst, c_start + e);
                cp_async4(dst + 128u, c_end + e);
                cp_async4(dst + 256u, c_idx + e);
                cp_async4(dst + 384u, c_mask + e);
            } else {
                cp_async4(dst, off + e);
                cp_async4(dst + 128u, off + e + 1u);
                sts_u32(dst + 256u, e);
                sts_u32(dst + 384u, 0xFFFFFFFFu);
            }
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
            // ---- lanes that run out of bytes in this iteration (or are idle) take the next entry of the pool ----
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
                    if (cand) {
                        // candidates are scattered over the column: pull this pool's strings towards L2 before lanes adopt them
                        const uint32_t sa = a_pool + pb * kFsPoolBytes + lane * 4u;
                        const uint32_t s0 = lds_u32_v(sa), e0 = lds_u32_v(sa + 128u);
                        if (ah_base + lane < N)
                            for (uint32_t a = s0 & ~127u, k = 0; a < e0 && k < 4u; a += 128u, ++k)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(col + a));
                    }
                    claim_ahead();
                }
                const uint32_t idx = pool_next + __popc(need & lt_mask);
                if (want && idx < pool_end) {
                    const uint32_t sa = a_pool + pb * kFsPoolBytes + (idx & 31u) * 4u;
                    const uint32_t s0 = lds_u32_v(sa), e0 = lds_u32_v(sa + 128u);
                    // empty fields are left to the epilogue kernel; a candidate of other units of the field is not ours
                    if (e0 > s0 && ((lds_u32_v(sa + 384u) >> gate_bit) & 1u)) {
                        sts_u32(a_slot + 3u * kFsSlotStride, lds_u32_v(sa + 256u));
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
#ifdef PGW_EXP_SCANLOG
            ++n_iter;
#endif

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
                    // the class map sits on a 256-byte boundary of the shared window: one PRMT extracts the byte AND adds the base
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
                        const Sink row = sink_of(p, lds_u32_v(a_slot));
                        uint32_t t_latch = lds_u32_v(a_slot + kFsSlotStride), t_last = lds_u32_v(a_slot + 2u * kFsSlotStride);
                        if (mx4 >= trap) {
                            uint32_t t_state = state;
                            fs_slow_word(p, cu.tbl_off, cu.n_classes, acclo, cu.acc_base, clsaddr, w, m4, &t_state, &t_last, &t_latch, row);
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
                                    if (st1 != t_last) fire_atom(row, a1);
                                    t_last = st1;
                                } else {
                                    t_last = fs_events_word(p, acclo, cu.acc_base, acc1addr, s01, s23, m4, t_last, &t_latch, row);
                                }
                            }
                        }
                        sts_u32(a_slot + kFsSlotStride, t_latch);
                        sts_u32(a_slot + 2u * kFsSlotStride, t_last);
                    }
                }
                state = spec;
            }
            // the field ends in this chunk, or the DFA reached an absorbing state: finish as at the end of the field
            if (finishing || (have && (state == abs0 || state == abs1))) {
                uint32_t e1 = 0xFFFFu;
                if (state < trap) e1 = lds_u16(end1addr + 2u * state);
                if (e1 != 0xFFFEu) {
                    const Sink row = sink_of(p, lds_u32_v(a_slot));
                    if (e1 != 0xFFFFu) {
                        fire_atom(row, e1);
                    } else if (cu.end_any) {
                        uint32_t t_latch = lds_u32_v(a_slot + kFsSlotStride);
                        fs_fire_list(p.end_idx, p.end_events, cu.end_base + state, row, &t_latch);
                    }
                }
                have = false;
                state = 0;  // an idle lane must not look like it sits in a cold or accepting state
#ifdef PGW_EXP_SCANLOG
                ++n_strings;
#endif
            }
        }
#ifdef PGW_EXP_SCANLOG
        {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            // per warp: lane 0 reports; strings summed over the warp
            for (int o = 16; o; o >>= 1) n_strings += __shfl_xor_sync(FULL, n_strings, o);
            if (lane == 0 && (tid >> 5) % 8 == 0)
                printf("SCANLOG cta %u warp %u unit %u uk %u N %u us %llu..%llu strings %u iters %u\n", blockIdx.x, tid >> 5, u, uk, N, t_unit0 / 1000ull % 100000000ull,
                       t1 / 1000ull % 100000000ull, n_strings, n_iter);
        }
#endif
    }
}

// Verdicts once every unit has been scanned: one thread per request (request_epilogue).  The tables of the small
// early-exit units it walks are staged into the CTA's shared memory first.
constexpr int kEpiThreads = 512;
// CTAs per SM the register allocation is held to: 4 x 512 = every thread slot of the SM (32 registers, a few spills).  The
// kernel is bound by the latency of dependent loads; measured 0.465 ms per 4 M requests against 0.628 ms at 3 CTAs (40 registers).
#ifndef PGW_EPI_CTAS
#define PGW_EPI_CTAS 4
#endif
__global__ void __launch_bounds__(kEpiThreads, PGW_EPI_CTAS) waf_epilogue_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(256) uint8_t esm[];
    uint8_t* img = esm + ((0u - smem_u32(esm)) & 255u);   // class maps sit on 256-byte boundaries
    for (uint32_t k = 0; k < p.n_prefix; ++k) {
        uint4* d = reinterpret_cast<uint4*>(img + p.prefix_img[k]);
        const uint4* s = reinterpret_cast<const uint4*>(p.images + p.pdesc[k].img_off);
        for (uint32_t i = threadIdx.x; i < p.pdesc[k].img_bytes / 16u; i += kEpiThreads) d[i] = __ldg(s + i);
    }
    __syncthreads();
    const uint32_t a_img = smem_u32(img);
    // warp-uniform trip count: every lane of a warp goes through request_epilogue together
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < p.n; base += gridDim.x * blockDim.x) {
        const uint32_t rr = base + (threadIdx.x & 31u);
        const bool valid = rr < p.n;
        request_epilogue(p, valid ? rr : p.n - 1u, valid, a_img);
    }
}

// Requests with several true atoms (the multi list of the epilogue): one thread per request.
__global__ void __launch_bounds__(256) waf_multi_kernel(const __grid_constant__ KParams p) {
    const uint32_t count = *p.multi_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) request_multi_thread(p, p.multi_list[i]);
}
