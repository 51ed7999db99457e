// Part of kernels.cu (included inside namespace pgw { namespace { ... } }, one translation unit: device functions are
// not linked across files).  Helpers shared by every kernel path: shared-window accessors, mbarrier / TMA bulk copy, event lists,
// rule bytecode, longest-prefix lookup and the per-request epilogue.

__host__ __device__ inline uint32_t r16(uint32_t x) { return (x + 15u) & ~15u; }


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

__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---- what a scan reports ----
__device__ __forceinline__ void red_or(uint32_t* addr, uint32_t v) { asm volatile("red.global.or.b32 [%0], %1;" ::"l"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void red_max(uint32_t* addr, uint32_t v) { asm volatile("red.global.max.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory"); }

// Where a scan writes what it found about one request: the request's atom bitmap row and its two info words
// (KParams::info: largest fired atom + 1, 0x4000 - smallest fired atom).  All three are fire-and-forget reductions.
struct Sink {
    uint32_t* row;
    uint32_t* inf;
};
__device__ __forceinline__ Sink sink_of(const KParams& p, uint32_t ridx) { return Sink{p.rows + (size_t)ridx * p.atom_words, p.info + 2u * (size_t)ridx}; }
__device__ __forceinline__ void fire_atom(const Sink& k, uint32_t at) {
    red_or(k.row + (at >> 5), 1u << (at & 31));
    red_max(k.inf, at + 1u);
    red_max(k.inf + 1, 0x4000u - at);
}

// events of CSR row `ci` applied through `fire(atom)`; true if all of them were plain FIREs
template <class Fire>
__device__ __forceinline__ bool fs_apply_list(const uint32_t* idx, const uint32_t* events, uint32_t ci, Fire&& fire, uint32_t* latch) {
    uint32_t a = __ldg(idx + ci), b = __ldg(idx + ci + 1);
    uint32_t l = *latch;
    bool pure = true;
    for (uint32_t i = a; i < b; ++i) {
        const uint32_t e = __ldg(events + i);
        const uint32_t kind = e >> kEvKindShift, lb = 1u << ((e >> kEvLatchShift) & 31u), at = e & kEvAtomMask;
        if (kind == 0u || (kind == 1u && (l & lb))) fire(at);
        else if (kind == 2u) l &= ~lb;
        else if (kind == 3u) l |= lb;
        pure &= kind == 0u;
    }
    *latch = l;
    return pure;
}
// ... to a request's sink in global memory (the scan kernel)
__device__ __forceinline__ bool fs_fire_list(const uint32_t* idx, const uint32_t* events, uint32_t ci, const Sink& row, uint32_t* latch) {
    return fs_apply_list(idx, events, ci, [&](uint32_t at) { fire_atom(row, at); }, latch);
}

__device__ __forceinline__ bool eval_rule(const uint16_t* __restrict__ code, uint32_t a, uint32_t b, const uint32_t* row) {
    uint32_t st = 0;
    for (uint32_t i = a; i < b; ++i) {
        uint32_t op = __ldg(code + i);
        if (op < 0x4000u) st = (st << 1) | ((row[op >> 5] >> (op & 31)) & 1u);
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
    // last range whose start <= (hi,lo), searched inside the address's 16-bit bucket (lpm.cpp: v6_top)
    const uint32_t t = (uint32_t)(hi >> 48);
    uint32_t l = __ldg(p.v6_top + t), r = __ldg(p.v6_top + t + 1u) + 1u;
    while (r - l > 1) {
        uint32_t m = (l + r) >> 1;
        uint64_t mh = __ldg(p.v6_hi + m), ml = __ldg(p.v6_lo + m);
        bool le = mh < hi || (mh == hi && ml <= lo);
        if (le) l = m;
        else r = m;
    }
    return __ldg(p.v6_leaf + l);
}

// INT_EXPR (program.hpp IntTok): both operand expressions of a comparison, evaluated with checked i64 arithmetic.
// Returns false when the evaluation errors (overflow, division by zero, INT64_MIN / -1).
__device__ __noinline__ bool int_expr_eval(const KParams& p, uint32_t off, uint32_t r, int64_t asn, int64_t* lhs, int64_t* rhs) {
    int64_t st[kIntExprStack];
    int sp = 0;
    for (const int64_t* t = p.iexpr + off;; ++t) {
        const int64_t w = __ldg(t);
        const uint32_t op = (uint32_t)((uint64_t)w >> 56);
        if (op == IT_END) break;
        if (op == IT_CONST) { st[sp++] = (w << 8) >> 8; continue; }   // sign-extend the 56-bit operand
        if (op == IT_CONST64) { st[sp++] = __ldg(++t); continue; }
        if (op == IT_FEAT) {
            const uint32_t f = (uint32_t)(w & 0xFF);
            int64_t x;
            if (f == 0u) x = p.port ? (int64_t)p.port[r] : 0;
            else if (f == 1u) x = asn;
            else { const uint32_t* o = p.off[f - 2u] + r; x = (int64_t)(o[1] - o[0]); }
            st[sp++] = x;
            continue;
        }
        if (op == IT_NEG) {
            if (st[sp - 1] == INT64_MIN) return false;
            st[sp - 1] = -st[sp - 1];
            continue;
        }
        const int64_t b = st[--sp], a = st[sp - 1];
        int64_t v;
        if (op == IT_ADD) {
            v = (int64_t)((uint64_t)a + (uint64_t)b);
            if (((a ^ v) & (b ^ v)) < 0) return false;   // both operands differ in sign from the result
        } else if (op == IT_SUB) {
            v = (int64_t)((uint64_t)a - (uint64_t)b);
            if (((a ^ b) & (a ^ v)) < 0) return false;
        }
        else if (op == IT_MUL) {
            // checked multiply without a 128-bit product
            v = (int64_t)((uint64_t)a * (uint64_t)b);
            if (a != 0 && ((a == -1 && b == INT64_MIN) || (b == -1 && a == INT64_MIN) || v / a != b)) return false;
        } else {
            if (b == 0 || (a == INT64_MIN && b == -1)) return false;
            v = op == IT_DIV ? a / b : a % b;
        }
        st[sp - 1] = v;
    }
    *lhs = st[0];
    *rhs = st[1];
    return true;
}

// FIELD_CMP: one http_request field against another (0 ==, 1 starts_with, 2 ends_with, 3 contains, 4 <, 5 <=, 6 >, 7 >=)
__device__ __noinline__ bool field_cmp_eval(const KParams& p, uint32_t f1, uint32_t f2, uint32_t op, uint32_t r) {
    const uint32_t s1 = p.off[f1][r], n1 = p.off[f1][r + 1] - s1, s2 = p.off[f2][r], n2 = p.off[f2][r + 1] - s2;
    const uint8_t* a = p.col[f1] + s1;
    const uint8_t* b = p.col[f2] + s2;
    if (op >= 4u) {
        // byte-wise lexicographic order (Rust str::cmp): the first differing byte decides, else the shorter string is smaller
        const uint32_t m = n1 < n2 ? n1 : n2;
        int c = 0;
        for (uint32_t i = 0; i < m && c == 0; ++i) c = (int)__ldg(a + i) - (int)__ldg(b + i);
        if (c == 0) c = n1 < n2 ? -1 : n1 > n2 ? 1 : 0;
        return op == 4u ? c < 0 : op == 5u ? c <= 0 : op == 6u ? c > 0 : c >= 0;
    }
    if (op == 0u && n1 != n2) return false;
    if (n2 > n1) return false;
    auto same = [&](uint32_t at) {
        for (uint32_t i = 0; i < n2; ++i)
            if (__ldg(a + at + i) != __ldg(b + i)) return false;
        return true;
    };
    if (op == 0u || op == 1u) return same(0);
    if (op == 2u) return same(n1 - n2);
    for (uint32_t at = 0; at + n2 <= n1; ++at)
        if (same(at)) return true;
    return false;
}

// The INT_EXPR / FIELD_CMP predicates of a program (ns[rare_begin .. rare_begin + n_rare), at most 64): bit k = predicate k holds
__device__ __noinline__ uint64_t rare_atoms(const KParams& p, uint32_t r, int64_t asn) {
    uint64_t m = 0;
    for (uint32_t k = 0; k < p.n_rare; ++k) {
        const NsAtom a = p.ns[p.rare_begin + k];
        bool v = false;
        if (a.kind == 5) {  // (lhs <op> rhs), or "the evaluation errors"
            int64_t x = 0, y = 0;
            const bool ok = int_expr_eval(p, a.set_id, r, asn, &x, &y);
            if (a.op == kIntExprIsError) v = !ok;
            else if (ok) {
                switch (a.op) {
                    case 0: v = x == y; break;
                    case 1: v = x != y; break;
                    case 2: v = x < y; break;
                    case 3: v = x <= y; break;
                    case 4: v = x > y; break;
                    default: v = x >= y; break;
                }
            }
        } else {
            v = field_cmp_eval(p, a.feat, a.set_id, a.op, r);
        }
        if (v) m |= 1ull << k;
    }
    return m;
}

// One request's field walked on a small early-exit DFA whose whole table is in shared memory at `img` (class map,
// rows, acc1, end1: the unit image of compile.hpp): start-anchored patterns (starts_with, ==, ^...) are decided within
// the first few bytes, the walk stops at an absorbing state.  Fired atoms are reported through `fire(atom)`.
// (A variant that fetched the first eight bytes with three word loads and walked them in eight predicated steps was
// measured: 4 % slower -- it issues all eight steps although a walk ends after about five.)
template <class Fire>
__device__ __forceinline__ void prefix_walk(const KParams& p, const UnitDesc& ud, uint32_t img, const uint8_t* __restrict__ col, uint32_t s, uint32_t e,
                                            Fire&& fire) {
    const uint32_t C2 = 2u * ud.n_classes, acclo = ud.acc_lo, abs0 = ud.abs0, abs1 = ud.abs1;
    const uint32_t hot = img + ud.hot_off;
    uint32_t st = ud.start_state, latch = 0u;
#pragma unroll 1
    for (uint32_t pos = s; pos < e; ++pos) {
        const uint32_t cls = lds_u8(img + (uint32_t)__ldg(col + pos));
        st = lds_u16(hot + st * C2 + 2u * cls);
        if (st >= acclo) {
            const uint32_t a1 = lds_u16(img + ud.acc1_off + 2u * (st - acclo));
            if (a1 != 0xFFFFu) fire(a1);
            else fs_apply_list(p.acc_idx, p.acc_events, ud.acc_base + st - acclo, fire, &latch);
        }
        if (st == abs0 || st == abs1) break;  // absorbing: nothing can change any more
    }
    const uint32_t e1 = lds_u16(img + ud.end1_off + 2u * st);
    if (e1 != 0xFFFEu) {
        if (e1 != 0xFFFFu) fire(e1);
        else if (ud.end_any) fs_apply_list(p.end_idx, p.end_events, ud.end_base + st, fire, &latch);
    }
}

// What the per-request kernel finds true outside the scan kernel: up to four atoms packed 16 bits each (`n` counts all of
// them, also those beyond four), and the running largest / smallest true atom in the info-word encoding.
struct Extras {
    uint64_t xl;
    uint32_t n, amax, binv;
};

// The atoms that become true in the per-request kernel: client address -> {asn, country, ip-set mask} (DIR-24-8 / IPv6
// index), the small early-exit units (KParams::pdesc, tables in shared memory at `a_img`), end-of-field events of empty
// fields, the integer / list / country predicates, the rare expression kinds.  With `row` == nullptr the atoms are
// collected into `ex` (its amax / binv come in initialised from the scan's info words); with a row they are ORed into it
// (second pass of the few requests with more than four such atoms: collect_extras_row, out of line).
__device__ __forceinline__ void collect_extras(const KParams& p, uint32_t r, uint32_t a_img, Extras* ex, uint32_t* row) {
    uint64_t xl = 0;
    uint32_t nx = 0, xlast = 0xFFFFFFFFu, amax = ex->amax, binv = ex->binv;
    auto fn = [&](uint32_t a) {
        if (row) { row[a >> 5] |= 1u << (a & 31); return; }
        if (a == xlast) return;   // a sticky accepting state fires at every byte
        xlast = a;
        amax = max(amax, a + 1u);
        binv = max(binv, 0x4000u - a);
        if (nx < 4u) xl |= (uint64_t)a << (16u * nx);
        ++nx;
    };
    int64_t asn = 0;
    uint32_t country = (uint32_t)'X' | ((uint32_t)'X' << 8);
    uint32_t set_mask = 0;
    // client address -> leaf, in two steps around the early-exit walks: the DIR-24-8 word (or, for IPv6, nothing yet) is
    // requested first, so that its DRAM latency passes while the walks run; the leaf is fetched after them
    const bool lpm = p.need_lpm;
    const uint8_t* ip16 = p.ip + (size_t)r * 16;
    bool v6 = false;
    uint32_t dir_e = 0, ip_host = 0;
    if (lpm) {
        v6 = p.is_v6[r] != 0;
        if (!v6) {
            ip_host = __byte_perm(*reinterpret_cast<const uint32_t*>(ip16), 0, 0x0123);  // network order -> host integer
            dir_e = __ldg(p.dir24 + (ip_host >> 8));
        }
    }

    // small early-exit units: one walk over the first bytes of the field
    for (uint32_t k = 0; k < p.n_prefix; ++k) {
        const UnitDesc& ud = p.pdesc[k];
        const uint32_t* o = p.off[ud.field] + r;
        const uint32_t s0 = o[0], e0 = o[1];
        if (e0 > s0) prefix_walk(p, ud, a_img + p.prefix_img[k], p.col[ud.field], s0, e0, fn);
    }
    // end-of-field events of EMPTY fields (they reach neither the scan nor a prefix walk)
    for (uint32_t k = 0; k < p.n_start_end; ++k) {
        const UnitDesc& ud = p.units[p.start_end_unit[k]];
        const uint32_t* o = p.off[ud.field] + r;
        if (o[0] != o[1]) continue;
        const uint32_t a = __ldg(p.end_idx + ud.end_base + ud.start_state), b = __ldg(p.end_idx + ud.end_base + ud.start_state + 1);
        for (uint32_t i = a; i < b; ++i) {
            const uint32_t e = __ldg(p.end_events + i);
            if ((e >> kEvKindShift) == 0u) fn(e & kEvAtomMask);  // latch kinds cannot fire on an empty field
        }
    }
    if (lpm) {
        uint32_t leaf;
        if (!v6) leaf = (dir_e & 0x80000000u) ? __ldg(p.tbl8 + ((dir_e & 0x7FFFFFFFu) << 8) + (ip_host & 0xFFu)) : dir_e;
        else leaf = lpm_lookup(p, ip16, true);
        const LpmLeaf lf = p.leaves[leaf];
        set_mask = lf.set_mask;
        if (p.geo_loaded) {
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
    // integer predicates, one feature at a time: the feature's quick reject (compile.hpp) settles almost every request
#pragma unroll 1
    for (uint32_t j = 0; j < p.n_feat_used; ++j) {
        const uint32_t fe = p.feat_used[j];
        const uint32_t b0 = p.ns_begin[fe], b1 = p.ns_begin[fe + 1u];
        int64_t x;
        if (fe == 0u) x = p.port ? (int64_t)p.port[r] : 0;
        else if (fe == 1u) x = asn;
        else {
            const uint32_t* o = p.off[fe - 2u] + r;
            x = (int64_t)(o[1] - o[0]);
        }
        if (x >= p.ns_lo[fe] && x <= p.ns_hi[fe] && (x < p.ns_vmin[fe] || x > p.ns_vmax[fe])) continue;
        for (uint32_t i = b0; i < b1; ++i) {
            const NsAtom a = p.n_ns <= kMaxConstNs ? p.nsd[i] : p.ns[i];
            bool v = false;
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
            if (v) fn(a.atom);
        }
    }
    for (uint32_t i = p.ns_begin[7]; i < p.rare_begin; ++i) {
        const NsAtom a = p.n_ns <= kMaxConstNs ? p.nsd[i] : p.ns[i];
        bool v = false;
        if (a.kind == 3) {  // IP_SET
            v = (set_mask >> a.set_id) & 1u;
        } else {  // COUNTRY_SET
            uint32_t c0 = (country & 0xFFu) - 'A', c1 = ((country >> 8) & 0xFFu) - 'A';
            if (c0 < 26u && c1 < 26u) {
                uint32_t bit = c0 * 26u + c1;
                v = (__ldg(p.cset + a.set_id * kCountryWords + (bit >> 5)) >> (bit & 31)) & 1u;
            }
        }
        if (v) fn(a.atom);
    }
    // integer expressions and field-against-field predicates (rare in rule sets): one out-of-line call
    if (p.n_rare) {
        uint64_t m = rare_atoms(p, r, asn);
        while (m) {
            const uint32_t k = (uint32_t)__ffsll((long long)m) - 1u;
            m &= m - 1ull;
            fn((p.n_ns <= kMaxConstNs ? p.nsd[p.rare_begin + k] : p.ns[p.rare_begin + k]).atom);
        }
    }
    ex->xl = xl;
    ex->n = nx;
    ex->amax = amax;
    ex->binv = binv;
}

__device__ __noinline__ void collect_extras_row(const KParams& p, uint32_t r, uint32_t a_img, uint32_t* row) {
    Extras ex;
    ex.amax = ex.binv = 0u;
    collect_extras(p, r, a_img, &ex, row);
}

// Per-request work outside the gate and the scan + the verdict (http_listener.rs:196-264): collect_extras, the gates,
// the verdict and the service.
// The scan left, per request, two info words (KParams::info) and the bits of the fired atoms in the request's bitmap
// row.  Called by all 32 lanes of a converged warp (`valid` false for lanes past the end of the batch, which shadow the
// last request without storing anything).
//   no atom true            -> the precomputed verdict `vclean` (the row is never read)
//   one distinct atom true  -> the tabulated verdict `v1z[atom]`
//   two, disjoint rule sets -> the earlier of their `v1z` entries
//   otherwise               -> the row is completed and the request is appended to the multi list (waf_multi_kernel)
// Whatever the scan or this function wrote to the row / info words is written back to zero (the scratch invariant).
__device__ __forceinline__ void request_epilogue(const KParams& p, uint32_t r, bool valid, uint32_t a_img) {
    const uint32_t Aw = p.atom_words;
    const uint32_t FULL = 0xFFFFFFFFu;
    // everything that does not depend on another load is requested first: the flags, the scan's info words, and the
    // extents of the fields the early-exit units walk, whose first bytes are prefetched -- the walks would otherwise each
    // wait for DRAM in turn (offset -> first byte -> ...)
    const uint32_t flags = p.flags ? p.flags[r] : 0u;
    const uint2 inf = *reinterpret_cast<const uint2*>(p.info + 2u * (size_t)r);
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k)
        if (k < p.n_prefix) {
            const uint32_t* o = p.off[p.pdesc[k].field] + r;
            const uint32_t s0 = o[0];
            if (o[1] > s0) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.col[p.pdesc[k].field] + s0));
        }
    uint32_t* const row = p.rows + (size_t)r * Aw;
    Extras ex;
    ex.amax = inf.x;   // largest true atom + 1 (0: none), 0x4000 - smallest true atom
    ex.binv = inf.y;
    collect_extras(p, r, a_img, &ex, nullptr);
    const uint32_t amax = ex.amax, binv = ex.binv, nx = ex.n;
    const uint64_t xl = ex.xl;
    const bool any_atom = amax != 0u;
    const bool single = any_atom && (amax - 1u == 0x4000u - binv);
    const bool multi = any_atom && !single;
    // exactly two distinct true atoms (the smallest and the largest) whose rule sets are disjoint: their single-atom
    // table entries combine (compile.cpp: atom_sig), no rule is evaluated
    bool pair = false;
    const uint32_t a_lo = 0x4000u - binv, a_hi = amax - 1u;
    if (multi && nx <= 4u) {
        pair = true;
        for (uint32_t k = 0; k < nx; ++k) {
            const uint32_t a = (uint32_t)(xl >> (16u * k)) & 0xFFFFu;
            pair &= a == a_lo || a == a_hi;
        }
        if (pair && inf.x != 0u)   // atoms the scan fired: their bits are in the row
            for (uint32_t w = 0; w < Aw; ++w) {
                uint32_t v = row[w];
                if (w == (a_lo >> 5)) v &= ~(1u << (a_lo & 31));
                if (w == (a_hi >> 5)) v &= ~(1u << (a_hi & 31));
                if (v) { pair = false; break; }
            }
        if (pair) pair = (__ldg(p.atom_sig + a_lo) & __ldg(p.atom_sig + a_hi)) == 0ull;
    }
    if (multi && !pair && valid && nx) {  // complete the row (the scan's bits are in it already)
        if (nx <= 4u) {
            for (uint32_t k = 0; k < nx; ++k) {
                const uint32_t a = (uint32_t)(xl >> (16u * k)) & 0xFFFFu;
                row[a >> 5] |= 1u << (a & 31);
            }
        } else {
            collect_extras_row(p, r, a_img, row);
        }
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
        if (p.eval_gates && p.gate_atom >= 0) {
            if (single) bypass |= a_hi == (uint32_t)p.gate_atom;
            else if (pair) bypass |= a_lo == (uint32_t)p.gate_atom || a_hi == (uint32_t)p.gate_atom;
            else if (multi) bypass |= (row[p.gate_atom >> 5] >> (p.gate_atom & 31)) & 1u;
        }
        if (bypass) { verdict = V_BYPASS | (kNoRule << 2); decided = true; }
    }
    if (!decided && (flags & RF_PRE_CAPTCHA)) { verdict = V_CAPTCHA | (kNoRule << 2); decided = true; }

    const bool routes = p.service != nullptr && p.n_rules > p.n_waf_rules;
    uint32_t svc = kNoService;
    if (!decided && !any_atom) { verdict = p.vclean[cv]; svc = p.sclean; decided = true; }
    if (!decided && single) {
        verdict = __ldg(p.v1z + cv * p.n_atoms + a_hi);
        svc = routes ? (uint32_t)__ldg(p.s1z + a_hi) : kNoService;
        decided = true;
    }
    if (!decided && pair) {
        // first match over the union of the two rule sets: the entry with the smaller rule index (kNoRule is the largest)
        const uint32_t va = __ldg(p.v1z + cv * p.n_atoms + a_lo), vb = __ldg(p.v1z + cv * p.n_atoms + a_hi);
        verdict = (va >> 2) <= (vb >> 2) ? va : vb;
        svc = routes ? min((uint32_t)__ldg(p.s1z + a_lo), (uint32_t)__ldg(p.s1z + a_hi)) : kNoService;
        decided = true;
    }
    // several atoms and no gate decided: the request goes to the multi list (one ballot + one atomicAdd per warp)
    const bool to_list = valid && !decided;
    const uint32_t lm = __ballot_sync(FULL, to_list);
    if (lm) {
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t base = 0;
        if (lane == (uint32_t)__ffs(lm) - 1u) base = atomicAdd(p.multi_count, (uint32_t)__popc(lm));
        base = __shfl_sync(FULL, base, __ffs(lm) - 1);
        if (to_list) p.multi_list[base + (uint32_t)__popc(lm & ((1u << lane) - 1u))] = r;
    }
    if (!valid) return;
    // scratch goes back all-zero: the info words, the scan's bit if a single atom fired, a completed row that is not
    // handed to the multi kernel (which zeroes the rows it evaluates)
    if (inf.x != 0u) {
        *reinterpret_cast<uint2*>(p.info + 2u * (size_t)r) = make_uint2(0u, 0u);
        if (inf.x - 1u == 0x4000u - inf.y && !to_list) row[(inf.x - 1u) >> 5] = 0u;
    }
    if (multi && !to_list)
        for (uint32_t w = 0; w < Aw; ++w) row[w] = 0u;
    if (to_list) return;
    p.verdict[r] = verdict;
    // http_listener.rs:266-272: only a request the rules let through reaches the services; the first service whose
    // route is absent or true takes it, none => 404 (kNoService)
    if (p.service) p.service[r] = (uint16_t)(((verdict & 3u) == V_ALLOW && routes) ? svc : kNoService);
}

// A request with several true atoms, evaluated by one THREAD of waf_multi_kernel (the work is a chain of dependent
// look-ups -- deviating atom -> its rules -> their bytecode -- so what counts is how many requests are in flight, not
// how many lanes share one).  Deviations from the expected atom vector: none -> v0 / s0, exactly one -> v1[atom] /
// s1[atom], otherwise the rules that mention a deviating atom plus the ones true by default, first terminal one wins
// (http_listener.rs:251-264).  The bitmap row is read in place and goes back all-zero.
__device__ __forceinline__ void request_multi_thread(const KParams& p, uint32_t r) {
    const uint32_t Aw = p.atom_words;
    uint32_t* const row = p.rows + (size_t)r * Aw;
    const uint32_t flags = p.flags ? p.flags[r] : 0u;
    const uint32_t cv = flags & RF_CAPTCHA_VERIFIED, tshift = 2u * cv;
    const bool routes = p.service != nullptr && p.n_rules > p.n_waf_rules;
    uint32_t ndev = 0, dev_atom = 0;
    for (uint32_t w = 0; w < Aw; ++w) {
        const uint32_t x = (row[w] ^ __ldg(p.expect + w)) & __ldg(p.care + w);
        if (x) dev_atom = w * 32u + (uint32_t)__ffs(x) - 1u;
        ndev += (uint32_t)__popc(x);
    }
    uint32_t verdict, svc;
    if (ndev == 0u) { verdict = p.v0[cv]; svc = p.s0; }
    else if (ndev == 1u) { verdict = __ldg(p.v1 + cv * p.n_atoms + dev_atom); svc = routes ? (uint32_t)__ldg(p.s1 + dev_atom) : kNoService; }
    else {
        uint32_t best = kNoRule, best_svc = kNoRule;
        for (uint32_t w = 0; w < Aw; ++w) {
            uint32_t x = (row[w] ^ __ldg(p.expect + w)) & __ldg(p.care + w);
            while (x) {
                const uint32_t atom = w * 32u + (uint32_t)__ffs(x) - 1u;
                x &= x - 1u;
                const uint32_t i0 = __ldg(p.ar_idx + atom), i1 = __ldg(p.ar_idx + atom + 1);
                for (uint32_t i = i0; i < i1; ++i) {
                    const uint32_t rule = __ldg(p.ar_rules + i);  // ascending: WAF rules first, then service routes
                    if (rule < p.n_waf_rules) {
                        if (rule >= best || ((__ldg(p.term + rule) >> tshift) & 3u) == 0) continue;
                        if (eval_rule(p.code, __ldg(p.rule_off + rule), __ldg(p.rule_off + rule + 1), row)) best = rule;
                    } else if (routes && rule < best_svc) {
                        if (eval_rule(p.code, __ldg(p.rule_off + rule), __ldg(p.rule_off + rule + 1), row)) best_svc = rule;
                    }
                }
            }
        }
        // rules true by default (ascending): the first one that still holds
        for (uint32_t i = 0; i < p.n_dflt[cv]; ++i) {
            const uint32_t rule = __ldg(p.dflt[cv] + i);
            if (rule >= best) break;
            if (eval_rule(p.code, __ldg(p.rule_off + rule), __ldg(p.rule_off + rule + 1), row)) { best = rule; break; }
        }
        if (routes)
            for (uint32_t i = 0; i < p.n_dflt_services; ++i) {
                const uint32_t rule = __ldg(p.dflt_services + i);
                if (rule >= best_svc) break;
                if (eval_rule(p.code, __ldg(p.rule_off + rule), __ldg(p.rule_off + rule + 1), row)) { best_svc = rule; break; }
            }
        verdict = best == kNoRule ? (V_ALLOW | (kNoRule << 2)) : (((__ldg(p.term + best) >> tshift) & 3u) | (best << 2));
        svc = best_svc == kNoRule ? kNoService : best_svc - p.n_waf_rules;
    }
    for (uint32_t w = 0; w < Aw; ++w) row[w] = 0u;  // scratch goes back all-zero
    p.verdict[r] = verdict;
    if (p.service) p.service[r] = (uint16_t)(((verdict & 3u) == V_ALLOW && routes) ? svc : kNoService);
}

