// TEST-ONLY: CPU walk over the compiled device program (HostProgram).
//
// This is not a product path and not the oracle.  It executes, on the CPU and
// without CUDA, exactly the tables the host compiler would upload (DFAs, accept
// lists, rule bytecode, candidate indexes, LPM tables), mirroring the kernel's
// logic step by step.  `-m "not gpu"` tests use it to check the *compiler*
// against the oracle in this GPU-less container; the GPU tests then only have
// to establish kernel == tables.  The product library never links this file.
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pingoo_waf.h"
#include "../../pingoo_b200/csrc/nfa_bits.hpp"
#include "../../pingoo_b200/csrc/ruleset.hpp"
#include "../../pingoo_b200/csrc/yaml.hpp"

using namespace pgw;

namespace {

struct Sim {
    RulesetBuilder builder;
    HostProgram H;
    bool finalized = false;
    std::string last_error;
    uint64_t* stats = nullptr;  // optional: per field {requests gated, candidates}
    uint64_t* atom_hist = nullptr;
    uint64_t* lit_stats = nullptr;  // optional: {windows with a gram that has literal candidates, candidates compared, confirmed, largest candidate list, out of bounds, anchor misses}
};

int fail(Sim* s, const std::string& m, char* err, size_t cap) {
    if (s) s->last_error = m;
    if (err && cap) {
        size_t n = m.size() < cap - 1 ? m.size() : cap - 1;
        memcpy(err, m.data(), n);
        err[n] = 0;
    }
    return 1;
}

bool eval_rule(const HostProgram& H, uint32_t rule, const std::vector<uint32_t>& row) {
    uint32_t st = 0;
    for (uint32_t i = H.rule_off[rule]; i < H.rule_off[rule + 1]; ++i) {
        uint32_t op = H.code[i];
        if (op < 0x4000u) st = (st << 1) | ((row[op >> 5] >> (op & 31)) & 1u);
        else if (op == OP_NOT) st ^= 1u;
        else if (op == OP_AND) st = ((st >> 1) & ~1u) | (st & (st >> 1) & 1u);
        else if (op == OP_OR) st = ((st >> 1) & ~1u) | ((st | (st >> 1)) & 1u);
        else if (op == OP_PUSH0) st <<= 1;
        else st = (st << 1) | 1u;
    }
    return st & 1u;
}

uint32_t lpm_lookup(const LpmTables& T, const uint8_t* ip16, bool v6) {
    if (!v6) {
        uint32_t a = (uint32_t)ip16[0] << 24 | (uint32_t)ip16[1] << 16 | (uint32_t)ip16[2] << 8 | ip16[3];
        uint32_t e = T.dir24[a >> 8];
        if (e & 0x80000000u) e = T.tbl8[((size_t)(e & 0x7FFFFFFFu) << 8) + (a & 0xFF)];
        return e;
    }
    uint64_t hi = 0, lo = 0;
    for (int k = 0; k < 8; ++k) hi = (hi << 8) | ip16[k];
    for (int k = 8; k < 16; ++k) lo = (lo << 8) | ip16[k];
    uint32_t l = 0, r = (uint32_t)T.v6_leaf.size();
    while (r - l > 1) {
        uint32_t m = (l + r) >> 1;
        bool le = T.v6_hi[m] < hi || (T.v6_hi[m] == hi && T.v6_lo[m] <= lo);
        if (le) l = m;
        else r = m;
    }
    return T.v6_leaf[l];
}

bool geo_skip(const uint8_t* ip16, bool v6) {
    if (!v6) return ip16[0] == 127 || (ip16[0] >> 4) == 0xE;
    if (ip16[0] == 0xFF) return true;
    for (int k = 0; k < 15; ++k) if (ip16[k]) return false;
    return ip16[15] == 1;
}

}  // namespace

extern "C" {

void* pgwsim_create(const pgw_rule_desc* rules, uint32_t n, const pgw_options* opt, char* err, size_t cap) {
    Sim* s = new Sim();
    if (opt) {
        if (opt->max_dfa_states > 0) s->builder.options.max_dfa_states = opt->max_dfa_states;
        if (opt->max_unit_table_bytes > 0) s->builder.options.max_unit_table_bytes = (size_t)opt->max_unit_table_bytes;
        s->builder.options.eval_gates = opt->eval_gates != 0;
        s->builder.options.candidate_gate = (opt->disable_candidate_gate & 1) == 0;
        s->builder.options.literal_confirm = (opt->disable_candidate_gate & 2) == 0;
    }
    for (uint32_t i = 0; i < n; ++i) {
        std::string e;
        if (!s->builder.add_rule(rules[i].name, rules[i].expression, rules[i].actions, rules[i].n_actions, e)) {
            fail(nullptr, e, err, cap);
            delete s;
            return nullptr;
        }
    }
    return s;
}

int pgwsim_lists_add(void* h, const char* name, int type, const uint8_t* csv, size_t len, char* err, size_t cap) {
    Sim* s = (Sim*)h;
    std::string e;
    if (!s->builder.add_list(name, type, csv, len, e)) return fail(s, e, err, cap);
    return 0;
}

int pgwsim_geoip_load(void* h, const uint8_t* mmdb, size_t len, char* err, size_t cap) {
    Sim* s = (Sim*)h;
    std::string e;
    if (!s->builder.load_geoip(mmdb, len, e)) return fail(s, e, err, cap);
    return 0;
}

int pgwsim_finalize(void* h, char* err, size_t cap) {
    Sim* s = (Sim*)h;
    std::string e;
    if (!s->builder.finalize(&s->H, e)) return fail(s, e, err, cap);
    s->finalized = true;
    return 0;
}

size_t pgwsim_describe(void* h, char* buf, size_t cap) {
    Sim* s = (Sim*)h;
    std::string d = s->H.summary();
    for (auto& w : s->H.warnings) d += "\nwarning: " + w;
    if (buf && cap) {
        size_t n = d.size() < cap - 1 ? d.size() : cap - 1;
        memcpy(buf, d.data(), n);
        buf[n] = 0;
    }
    return d.size();
}

// geo lookup exactly as the kernel resolves it
int pgwsim_geoip_lookup(void* h, const uint8_t* ip, const uint8_t* is_v6, uint32_t n, uint32_t* asn, uint16_t* country) {
    Sim* s = (Sim*)h;
    const HostProgram& H = s->H;
    for (uint32_t r = 0; r < n; ++r) {
        asn[r] = 0;
        country[r] = (uint16_t)('X' | ('X' << 8));
        const uint8_t* ip16 = ip + (size_t)r * 16;
        if (H.lpm.geo_loaded && !geo_skip(ip16, is_v6[r] != 0)) {
            const LpmLeaf& lf = H.lpm.leaves[lpm_lookup(H.lpm, ip16, is_v6[r] != 0)];
            asn[r] = lf.asn;
            country[r] = lf.country;
        }
    }
    return 0;
}

int pgwsim_services_set(void* h, const pgw_service_desc* sv, uint32_t n, char* err, size_t cap) {
    Sim* s = (Sim*)h;
    std::string e;
    for (uint32_t i = 0; i < n; ++i)
        if (!s->builder.add_service(sv[i].name, sv[i].route, e)) { snprintf(err, cap, "%s", e.c_str()); return 1; }
    return 0;
}

int pgwsim_evaluate_routed(void* h, const pgw_batch* b, uint32_t* out, uint16_t* svc_out);
int pgwsim_evaluate(void* h, const pgw_batch* b, uint32_t* out) { return pgwsim_evaluate_routed(h, b, out, nullptr); }
static int evaluate_range(Sim* s, const pgw_batch* b, uint32_t* out, uint16_t* svc_out, uint32_t r_lo, uint32_t r_hi);

int pgwsim_evaluate_routed(void* h, const pgw_batch* b, uint32_t* out, uint16_t* svc_out) {
    Sim* s = (Sim*)h;
    if (!s->finalized) return 1;
    return evaluate_range(s, b, out, svc_out, 0, b->n);
}

// the same walk on `threads` host threads (contiguous ranges): the table-driven CPU leg of bench.py
int pgwsim_evaluate_mt(void* h, const pgw_batch* b, uint32_t* out, int threads) {
    Sim* s = (Sim*)h;
    if (!s->finalized || s->stats || s->atom_hist) return 1;
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    std::vector<int> rc((size_t)threads, 0);
    for (int t = 0; t < threads; ++t) {
        const uint32_t lo = (uint32_t)((uint64_t)b->n * t / threads), hi = (uint32_t)((uint64_t)b->n * (t + 1) / threads);
        th.emplace_back([=, &rc]() { rc[t] = evaluate_range(s, b, out, nullptr, lo, hi); });
    }
    for (auto& x : th) x.join();
    for (int v : rc) if (v) return v;
    return 0;
}

static int evaluate_range(Sim* s, const pgw_batch* b, uint32_t* out, uint16_t* svc_out, uint32_t r_lo, uint32_t r_hi) {
    const HostProgram& H = s->H;
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    const uint32_t Aw = H.atom_words;
    std::vector<uint32_t> row(Aw);
    for (uint32_t r = r_lo; r < r_hi; ++r) {
        std::fill(row.begin(), row.end(), 0);
        // candidate gate (gate.hpp): every even-aligned 4-byte window of the column that overlaps the field
        uint32_t cand[N_FIELDS] = {0, 0, 0, 0, 0};  // per field: mask of the gated units the request is a candidate for
        for (int f = 0; f < N_FIELDS; ++f) {
            const GateTables& G = H.gate[f];
            if (!G.present) continue;
            const uint8_t* bytes = cols[f]->bytes;
            const uint32_t a = cols[f]->offsets[r], e = cols[f]->offsets[r + 1], total = cols[f]->offsets[b->n];
            uint32_t j = a >= 3 ? a - 3 : 0;
            j += j & 1u;
            for (; j < e; j += 2) {
                uint32_t w = 0;
                for (uint32_t k = 0; k < 4; ++k)
                    if (j + k < total) w |= (uint32_t)bytes[j + k] << (8 * k);  // past the column: zeros
                uint32_t lb = 0, lc = 0;
                cand[f] |= G.probe(w, &lb, &lc);
                // finite-string patterns announced by the window's gram: compared in place (the resolve kernel's job)
                if (s->lit_stats && lc) { s->lit_stats[0]++; s->lit_stats[1] += lc; if (lc > s->lit_stats[3]) s->lit_stats[3] = lc; }
                for (uint32_t c = 0; c < lc; ++c) {
                    const uint32_t cd = G.lit_cand[lb + c];
                    const LitDesc& d = G.lits[cd >> 2];
                    if (s->lit_stats) {
                        const int64_t at = (int64_t)j + (int64_t)(cd & 3u) - 1;
                        if (at < (int64_t)a || at + d.len > (int64_t)e) s->lit_stats[4]++;
                        else if (((d.flags & 1) && at != (int64_t)a) || ((d.flags & 2) && at + d.len != (int64_t)e)) s->lit_stats[5]++;
                    }
                    if (G.lit_matches(d, bytes, a, e, (int64_t)j + (int64_t)(cd & 3u) - 1)) { row[d.atom >> 5] |= 1u << (d.atom & 31); if (s->lit_stats) s->lit_stats[2]++; }
                }
            }
            if (s->stats) { s->stats[2 * f] += 1; s->stats[2 * f + 1] += cand[f] ? 1 : 0; }
        }
        // scan units
        for (const UnitDesc& u : H.units) {
            if (u.mode == UM_CANDIDATES && !((cand[u.field] >> u.gate_bit) & 1u)) continue;
            const uint8_t* bytes = cols[u.field]->bytes;
            uint32_t a = cols[u.field]->offsets[r], e = cols[u.field]->offsets[r + 1];
            const uint8_t* cls = H.arena.data() + u.cls_off;
            const uint16_t* tbl = (const uint16_t*)(H.arena.data() + u.tbl_off);
            uint32_t st = u.start_state;
            uint32_t latch = 0;
            auto run_events = [&](const std::vector<uint32_t>& idx, const std::vector<uint32_t>& ev, uint32_t ci) {
                for (uint32_t k = idx[ci]; k < idx[ci + 1]; ++k) {
                    uint32_t w = ev[k], kind = w >> kEvKindShift, lb = 1u << ((w >> kEvLatchShift) & 31u), at = w & kEvAtomMask;
                    if (kind == 0 || (kind == 1 && (latch & lb))) row[at >> 5] |= 1u << (at & 31);
                    else if (kind == 2) latch &= ~lb;
                    else if (kind == 3) latch |= lb;
                }
            };
            for (uint32_t i = a; i < e; ++i) {
                st = tbl[st * u.n_classes + cls[bytes[i]]];
                if (st >= u.acc_lo) run_events(H.acc_idx, H.acc_events, u.acc_base + st - u.acc_lo);
                if (st == u.abs0 || st == u.abs1) break;  // absorbing: the kernel finishes the field here
            }
            if (u.end_any) run_events(H.end_idx, H.end_events, u.end_base + st);
        }
        // bit-parallel NFA units (kernel_bitset.cuh): every request, whole field
        for (const BitsetUnitDesc& bu : H.bitset_units)
            bitset_walk_host(bu, H.bitset_blob.data() + bu.blob_off, cols[bu.field]->bytes, cols[bu.field]->offsets[r], cols[bu.field]->offsets[r + 1],
                             [&](uint32_t at) { row[at >> 5] |= 1u << (at & 31); });
        // per-request predicates
        uint32_t flags = b->flags ? b->flags[r] : 0;
        int64_t asn = 0;
        uint32_t country = 'X' | ('X' << 8);
        uint32_t set_mask = 0;
        bool geo_on_device = H.lpm.geo_loaded && H.needs_geo_cols && !(b->asn && b->country);
        if (H.needs_ip || geo_on_device) {
            const uint8_t* ip16 = b->ip + (size_t)r * 16;
            bool v6 = b->ip_is_v6[r] != 0;
            const LpmLeaf& lf = H.lpm.leaves[lpm_lookup(H.lpm, ip16, v6)];
            set_mask = lf.set_mask;
            if (H.lpm.geo_loaded && !(b->asn && b->country) && !geo_skip(ip16, v6)) { asn = lf.asn; country = lf.country; }
        }
        // the geo columns count only when BOTH are supplied (capi.cu launch_on nulls them otherwise; the oracle likewise)
        if (b->asn && b->country) { asn = b->asn[r]; country = b->country[r]; }
        for (const NsAtom& a : H.ns_atoms) {
            bool v = false;
            if (a.kind == AtomDesc::INT_CMP || a.kind == AtomDesc::INT_SET) {
                int64_t x;
                if (a.feat == IF_PORT) x = b->remote_port ? b->remote_port[r] : 0;
                else if (a.feat == IF_ASN) x = asn;
                else { int f = a.feat - IF_LEN0; x = (int64_t)(cols[f]->offsets[r + 1] - cols[f]->offsets[r]); }
                if (a.kind == AtomDesc::INT_CMP) {
                    switch (a.op) {
                        case CMP_EQ: v = x == a.cval; break;
                        case CMP_NE: v = x != a.cval; break;
                        case CMP_LT: v = x < a.cval; break;
                        case CMP_LE: v = x <= a.cval; break;
                        case CMP_GT: v = x > a.cval; break;
                        default: v = x >= a.cval; break;
                    }
                } else {
                    uint32_t l = H.iset_off[a.set_id], hgh = H.iset_off[a.set_id + 1];
                    while (l < hgh) {
                        uint32_t m = (l + hgh) >> 1;
                        if (H.iset_vals[m] == x) { v = true; break; }
                        if (H.iset_vals[m] < x) l = m + 1;
                        else hgh = m;
                    }
                }
            } else if (a.kind == AtomDesc::IP_SET) {
                v = (set_mask >> a.set_id) & 1u;
            } else if (a.kind == AtomDesc::INT_EXPR) {
                // program.hpp IntTok: both operands with checked i64 arithmetic
                int64_t st[kIntExprStack];
                int sp = 0;
                bool ok = true;
                for (size_t t = a.set_id; ok; ++t) {
                    const int64_t w = H.iexpr[t];
                    const uint32_t op = (uint32_t)((uint64_t)w >> 56);
                    if (op == IT_END) break;
                    if (op == IT_CONST) { st[sp++] = (int64_t)((uint64_t)w << 8) >> 8; continue; }
                    if (op == IT_CONST64) { st[sp++] = H.iexpr[++t]; continue; }
                    if (op == IT_FEAT) {
                        const int f = (int)(w & 0xFF);
                        int64_t x;
                        if (f == IF_PORT) x = b->remote_port ? b->remote_port[r] : 0;
                        else if (f == IF_ASN) x = asn;
                        else { int ff = f - IF_LEN0; x = (int64_t)(cols[ff]->offsets[r + 1] - cols[ff]->offsets[r]); }
                        st[sp++] = x;
                        continue;
                    }
                    if (op == IT_NEG) { if (st[sp - 1] == INT64_MIN) ok = false; else st[sp - 1] = -st[sp - 1]; continue; }
                    const int64_t y = st[--sp], x = st[sp - 1];
                    int64_t z = 0;
                    if (op == IT_ADD) ok = !__builtin_add_overflow(x, y, &z);
                    else if (op == IT_SUB) ok = !__builtin_sub_overflow(x, y, &z);
                    else if (op == IT_MUL) ok = !__builtin_mul_overflow(x, y, &z);
                    else if (y == 0 || (x == INT64_MIN && y == -1)) ok = false;
                    else z = op == IT_DIV ? x / y : x % y;
                    st[sp - 1] = z;
                }
                if (a.op == kIntExprIsError) v = !ok;
                else if (ok) {
                    const int64_t x = st[0], y = st[1];
                    switch (a.op) {
                        case CMP_EQ: v = x == y; break;
                        case CMP_NE: v = x != y; break;
                        case CMP_LT: v = x < y; break;
                        case CMP_LE: v = x <= y; break;
                        case CMP_GT: v = x > y; break;
                        default: v = x >= y; break;
                    }
                }
            } else if (a.kind == AtomDesc::FIELD_CMP) {
                const uint32_t f1 = a.feat, f2 = a.set_id;
                const uint32_t s1 = cols[f1]->offsets[r], n1 = cols[f1]->offsets[r + 1] - s1, s2 = cols[f2]->offsets[r], n2 = cols[f2]->offsets[r + 1] - s2;
                const uint8_t* x = cols[f1]->bytes + s1;
                const uint8_t* y = cols[f2]->bytes + s2;
                if (a.op >= 4) {
                    const uint32_t m = n1 < n2 ? n1 : n2;
                    int c = m ? memcmp(x, y, m) : 0;
                    if (c == 0) c = n1 < n2 ? -1 : n1 > n2 ? 1 : 0;
                    v = a.op == 4 ? c < 0 : a.op == 5 ? c <= 0 : a.op == 6 ? c > 0 : c >= 0;
                } else if (a.op == 0) v = n1 == n2 && memcmp(x, y, n1) == 0;
                else if (n2 > n1) v = false;
                else if (a.op == 1) v = memcmp(x, y, n2) == 0;
                else if (a.op == 2) v = memcmp(x + n1 - n2, y, n2) == 0;
                else {
                    for (uint32_t at = 0; at + n2 <= n1 && !v; ++at) v = memcmp(x + at, y, n2) == 0;
                }
            } else {
                uint32_t c0 = (country & 0xFF) - 'A', c1 = ((country >> 8) & 0xFF) - 'A';
                if (c0 < 26 && c1 < 26) {
                    uint32_t bit = c0 * 26 + c1;
                    v = (H.cset_words[a.set_id * kCountryWords + (bit >> 5)] >> (bit & 31)) & 1u;
                }
            }
            if (v) row[a.atom >> 5] |= 1u << (a.atom & 31);
        }
        uint32_t cv = flags & RF_CAPTCHA_VERIFIED;
        uint32_t verdict = 0;
        uint32_t single_dev = 0xFFFFFFFFu;
        bool decided = false;
        if (flags & RF_PRE_BLOCK) { verdict = V_BLOCK | (kNoRule << 2); decided = true; }
        if (!decided && H.eval_gates) {
            uint32_t ual = b->user_agent.offsets[r + 1] - b->user_agent.offsets[r];
            if (ual == 0 || ual >= 256) { verdict = V_BLOCK | (kNoRule << 2); decided = true; }
        }
        if (!decided) {
            bool bypass = flags & RF_BYPASS;
            if (H.eval_gates && H.gate_bypass_atom >= 0) bypass |= (row[H.gate_bypass_atom >> 5] >> (H.gate_bypass_atom & 31)) & 1u;
            if (bypass) { verdict = V_BYPASS | (kNoRule << 2); decided = true; }
        }
        if (!decided && (flags & RF_PRE_CAPTCHA)) { verdict = V_CAPTCHA | (kNoRule << 2); decided = true; }
        bool dirty = false;
        for (uint32_t w = 0; w < Aw; ++w) dirty |= row[w] != 0;
        if (!decided && !dirty) { verdict = H.vclean[cv]; decided = true; }  // the epilogue's clean path
        uint32_t n_true = 0, the_atom = 0;
        for (uint32_t w = 0; w < Aw; ++w) {
            if (row[w]) the_atom = w * 32 + (uint32_t)__builtin_ctz(row[w]);
            n_true += (uint32_t)__builtin_popcount(row[w]);
        }
        const bool single_true = n_true == 1;
        if (s->atom_hist) s->atom_hist[n_true < 3 ? n_true : 3]++;
        if (!decided && single_true) { verdict = H.v1z[(size_t)cv * H.n_atoms + the_atom]; decided = true; }  // the single-atom table
        // exactly two true atoms with disjoint rule signatures: the epilogue's pair path (compile.cpp: atom_sig)
        bool pair_true = false;
        uint32_t pa = 0, pb = 0;
        if (n_true == 2) {
            bool first = true;
            for (uint32_t w = 0; w < Aw; ++w)
                for (uint32_t x = row[w]; x; x &= x - 1) {
                    const uint32_t at = w * 32 + (uint32_t)__builtin_ctz(x);
                    if (first) { pa = at; first = false; } else pb = at;
                }
            pair_true = (H.atom_sig[pa] & H.atom_sig[pb]) == 0;
        }
        if (!decided && pair_true) {
            const uint32_t va = H.v1z[(size_t)cv * H.n_atoms + pa], vb = H.v1z[(size_t)cv * H.n_atoms + pb];
            verdict = (va >> 2) <= (vb >> 2) ? va : vb;
            decided = true;
        }
        if (!decided) {
            uint32_t diff = 0, ndev = 0, dev_atom = 0;
            for (uint32_t w = 0; w < Aw; ++w) {
                uint32_t x = (row[w] ^ H.expect[w]) & H.care[w];
                if (x) dev_atom = w * 32 + (uint32_t)__builtin_ctz(x);
                ndev += (uint32_t)__builtin_popcount(x);
                diff |= x;
            }
            if (ndev == 1) single_dev = dev_atom;
            if (!diff) verdict = H.v0[cv];
            else if (ndev == 1) verdict = H.v1[(size_t)cv * H.n_atoms + dev_atom];
            else {
                uint32_t best = kNoRule;
                uint32_t tshift = 2 * cv;
                for (uint32_t w = 0; w < Aw; ++w) {
                    uint32_t x = (row[w] ^ H.expect[w]) & H.care[w];
                    while (x) {
                        uint32_t bit = __builtin_ctz(x);
                        x &= x - 1;
                        uint32_t atom = w * 32 + bit;
                        for (uint32_t i = H.ar_idx[atom]; i < H.ar_idx[atom + 1]; ++i) {
                            uint32_t rule = H.ar_rules[i];
                            if (rule >= best) break;
                            if (((H.term[rule] >> tshift) & 3u) == 0) continue;
                            if (eval_rule(H, rule, row)) best = rule;
                        }
                    }
                }
                for (uint32_t rule : H.dflt_rules[cv]) {
                    if (rule >= best) break;
                    if (eval_rule(H, rule, row)) best = rule;
                }
                verdict = best == kNoRule ? (V_ALLOW | (kNoRule << 2)) : (((H.term[best] >> tshift) & 3u) | (best << 2));
            }
        }
        out[r] = verdict;
        if (svc_out) {
            uint32_t svc = kNoService;
            if ((verdict & 3u) == V_ALLOW && H.n_rules > H.n_waf_rules && !dirty) svc = H.sclean;
            else if ((verdict & 3u) == V_ALLOW && H.n_rules > H.n_waf_rules && single_true) svc = H.s1z[the_atom];
            else if ((verdict & 3u) == V_ALLOW && H.n_rules > H.n_waf_rules && pair_true) svc = std::min<uint32_t>(H.s1z[pa], H.s1z[pb]);
            else if ((verdict & 3u) == V_ALLOW && H.n_rules > H.n_waf_rules) {
                uint32_t diff = 0;
                for (uint32_t w = 0; w < Aw; ++w) diff |= (row[w] ^ H.expect[w]) & H.care[w];
                if (single_dev != 0xFFFFFFFFu) svc = H.s1[single_dev];
                else if (!diff) svc = H.s0;
                else {
                    uint32_t best = kNoRule;
                    for (uint32_t w = 0; w < Aw; ++w) {
                        uint32_t x = (row[w] ^ H.expect[w]) & H.care[w];
                        while (x) {
                            uint32_t bit = __builtin_ctz(x);
                            x &= x - 1;
                            uint32_t atom = w * 32 + bit;
                            for (uint32_t i = H.ar_idx[atom]; i < H.ar_idx[atom + 1]; ++i) {
                                uint32_t rule = H.ar_rules[i];
                                if (rule >= best) break;
                                if (rule < H.n_waf_rules) continue;
                                if (eval_rule(H, rule, row)) best = rule;
                            }
                        }
                    }
                    for (uint32_t rule : H.dflt_services) {
                        if (rule >= best) break;
                        if (eval_rule(H, rule, row)) best = rule;
                    }
                    if (best != kNoRule) svc = best - H.n_waf_rules;
                }
            }
            svc_out[r] = (uint16_t)svc;
        }
    }
    return 0;
}

// debug: per-state visit counts of unit `unit` over a batch (counts must hold n_states entries); returns n_states
uint32_t pgwsim_state_histogram(void* h, const pgw_batch* b, uint32_t unit, uint64_t* counts, uint32_t* acc_lo) {
    Sim* s = (Sim*)h;
    const HostProgram& H = s->H;
    if (unit >= H.units.size()) return 0;
    const UnitDesc& u = H.units[unit];
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    if (acc_lo) *acc_lo = u.acc_lo;
    if (!counts) return u.n_states;
    const uint8_t* cls = H.arena.data() + u.cls_off;
    const uint16_t* tbl = (const uint16_t*)(H.arena.data() + u.tbl_off);
    for (uint32_t r = 0; r < b->n; ++r) {
        uint32_t st = u.start_state;
        for (uint32_t i = cols[u.field]->offsets[r]; i < cols[u.field]->offsets[r + 1]; ++i) {
            st = tbl[st * u.n_classes + cls[cols[u.field]->bytes[i]]];
            counts[st]++;
        }
    }
    return u.n_states;
}

// debug: per field {requests seen by the gate, candidates} accumulated over later evaluate calls (10 counters)
void pgwsim_set_gate_stats(void* h, uint64_t* stats10) { ((Sim*)h)->stats = stats10; }

void pgwsim_destroy(void* h) { delete (Sim*)h; }

// the configuration-directory loader (config_dir.cpp) on the simulator's builder: the CPU half of pgw_ruleset_load_dir
void* pgwsim_load_dir(const char* folder, const char* listener, const char* geoip_dir, char* err, size_t cap) {
    Sim* s = new Sim();
    std::vector<std::string> dirs;
    if (geoip_dir) dirs.push_back(geoip_dir);
    else { dirs.push_back(folder); dirs.push_back("/usr/share/pingoo"); }
    std::string e;
    if (!load_config_dir(folder, listener, dirs, &s->builder, nullptr, e)) {
        fail(nullptr, e, err, cap);
        delete s;
        return nullptr;
    }
    return s;
}

}  // extern "C"

// Analysis tool (no product counterpart): shared-memory wavefronts of the scan's row look-ups for one unit, emulating
// one warp of 32 lanes that each walk one request's field and take the next request when done (the field path's
// schedule).  For every lock-step byte the active lanes read 2 bytes at row(state) * row_stride + 2 * column(class); a
// wavefront serves all lanes whose addresses fall in distinct banks or in the same 4-byte word.  `perm` (class -> column)
// and `state_rows` (state -> row) may be null (identity).
// out[0] = lock-steps with at least one active lane, out[1] = wavefronts, out[2] = active lane-steps,
// out[3] = distinct states summed over steps.
extern "C" int pgwsim_bank_stats(void* h, const pgw_batch* b, uint32_t unit, uint32_t row_stride, const uint16_t* perm, const uint32_t* state_rows,
                                 uint64_t* out) {
    Sim* s = (Sim*)h;
    const HostProgram& H = s->H;
    if (unit >= H.units.size()) return 1;
    const UnitDesc& u = H.units[unit];
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    const uint8_t* cls = H.arena.data() + u.cls_off;
    const uint16_t* tbl = (const uint16_t*)(H.arena.data() + u.tbl_off);
    const pgw_strcol* col = cols[u.field];
    uint64_t steps = 0, waves = 0, lane_steps = 0, distinct = 0;
    uint32_t next = 0;
    struct Lane { uint32_t pos, end, st; bool on; } L[32];
    for (auto& l : L) l.on = false;
    for (;;) {
        bool any = false;
        for (auto& l : L) {
            while (!l.on && next < b->n) {
                l.pos = col->offsets[next];
                l.end = col->offsets[next + 1];
                l.st = u.start_state;
                l.on = l.end > l.pos;
                ++next;
            }
            any |= l.on;
        }
        if (!any) break;
        uint32_t words[32], nw = 0, sts[32], ns = 0;
        for (auto& l : L) {
            if (!l.on) continue;
            const uint32_t c = cls[col->bytes[l.pos]];
            const uint32_t row = state_rows ? state_rows[l.st] : l.st;
            const uint32_t addr = row * row_stride + 2u * (perm ? perm[c] : c);
            const uint32_t w = addr >> 2;
            bool seen = false;
            for (uint32_t k = 0; k < nw; ++k) seen |= words[k] == w;
            if (!seen) words[nw++] = w;
            bool ss = false;
            for (uint32_t k = 0; k < ns; ++k) ss |= sts[k] == l.st;
            if (!ss) sts[ns++] = l.st;
            l.st = tbl[l.st * u.n_classes + c];
            if (++l.pos >= l.end) l.on = false;
            ++lane_steps;
        }
        uint32_t load[32] = {0}, mx = 0;
        for (uint32_t k = 0; k < nw; ++k) {
            const uint32_t v = ++load[words[k] & 31];
            if (v > mx) mx = v;
        }
        waves += mx;
        distinct += ns;
        ++steps;
    }
    out[0] = steps;
    out[1] = waves;
    out[2] = lane_steps;
    out[3] = distinct;
    return 0;
}

// debug: level-1 pass rate of the gate of field `f` over every even-aligned window of the column
// out[0] = windows, out[1] = level-1 passes, out[2] = exact-table hits
extern "C" void pgwsim_gate_window_stats(void* h, const pgw_batch* b, int f, uint64_t* out) {
    Sim* s = (Sim*)h;
    const GateTables& G = s->H.gate[f];
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    out[0] = out[1] = out[2] = 0;
    if (!G.present) return;
    const uint32_t total = cols[f]->offsets[b->n];
    for (uint32_t j = 0; j + 4 <= total; j += 2) {
        uint32_t w;
        memcpy(&w, cols[f]->bytes + j, 4);
        const uint32_t g = gate_fold(w);
        out[0]++;
        if (gate_l1_test(G.b1.data(), G.k1, g)) {
            out[1]++;
            if (G.probe(w)) out[2]++;
        }
    }
}

// debug: literal-confirmation work {windows whose gram has literal candidates, candidates compared, confirmed, longest list}
extern "C" void pgwsim_set_lit_stats(void* h, uint64_t* out4) { ((Sim*)h)->lit_stats = out4; }

// debug: histogram of the number of true atoms per request (out[0..3] = 0, 1, 2, >=3) -- sizes the epilogue's paths
extern "C" void pgwsim_set_atom_hist(void* h, uint64_t* out4) { ((Sim*)h)->atom_hist = out4; }

// the YAML subset reader, canonical dump (compared with PyYAML by tests/test_config_formats.py); returns the length needed
extern "C" size_t pgwsim_yaml_dump(const char* text, size_t len, char* buf, size_t cap, int* ok) {
    YNode root;
    std::string err;
    std::string out;
    if (yaml_parse(std::string(text, len), &root, err)) { *ok = 1; out = yaml_dump(root); }
    else { *ok = 0; out = err; }
    if (buf && cap) {
        size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size();
}
