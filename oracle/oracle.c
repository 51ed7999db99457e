/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Request loop, list loading and GeoIP lookup, restated from the reference:
 *   verdict loop      pingoo/listeners/http_listener.rs:196-264
 *   Rule semantics    pingoo/rules.rs:36-52
 *   list CSV loading  pingoo/lists.rs:62-113
 *   GeoIP             pingoo/geoip.rs:73-91,111-142 + pingoo/serde_utils.rs:1-9
 *   MMDB walk/decode  MaxMind DB file format 2.0 (maxminddb 0.24 Reader::lookup)
 */
#include "oracle.h"

#include <ctype.h>
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bel.h"
#include "rx.h"

typedef struct {
    char* name;
    bel_expr* expr; /* NULL: no expression => always matches */
    uint8_t* actions;
    uint32_t n_actions;
} orc_rule;

typedef struct {
    uint8_t* buf;
    size_t len;
    uint32_t node_count, record_size, ip_version;
    size_t tree_bytes, data_start, data_len;
    uint32_t v4_start; /* node reached after 96 zero bits in an IPv6 tree */
    int v4_start_bits; /* how many of the 96 bits were actually consumed before a non-node record */
} orc_mmdb;

struct orc_ruleset {
    orc_rule* rules;
    uint32_t n_rules;
    bel_list* lists;
    size_t n_lists;
    bel_lists lists_view;
    int eval_gates;
    orc_mmdb* geo;
    /* services (http_listener.rs:266-270): tried in order; route == NULL matches every request */
    orc_rule* services;
    uint32_t n_services;
};

static void set_err(char* err, size_t cap, const char* msg) {
    if (err && cap) snprintf(err, cap, "%s", msg);
}

int orc_compile_expression(const char* expr, char* err, size_t cap) {
    bel_expr* e = bel_compile(expr, err, cap);
    if (!e) return 1;
    bel_free(e);
    return 0;
}

int orc_validate_expression(const char* expr, char* err, size_t cap) {
    if (!expr[0]) { set_err(err, cap, "Expression is not valid: expression is empty"); return 1; }
    bel_expr* e = bel_compile(expr, err, cap);
    if (!e) return 1;
    int uses_in = bel_uses_in(e);
    bel_free(e);
    if (uses_in) { set_err(err, cap, "Expression is not valid: unknown operator: in"); return 1; }
    return 0;
}

orc_ruleset* orc_create(const pgw_rule_desc* rules, uint32_t n, int eval_gates, char* err, size_t cap) {
    orc_ruleset* rs = (orc_ruleset*)calloc(1, sizeof *rs);
    rs->rules = (orc_rule*)calloc(n ? n : 1, sizeof(orc_rule));
    rs->n_rules = n;
    rs->eval_gates = eval_gates;
    for (uint32_t i = 0; i < n; ++i) {
        orc_rule* r = &rs->rules[i];
        r->name = strdup(rules[i].name ? rules[i].name : "");
        r->n_actions = rules[i].n_actions;
        r->actions = (uint8_t*)malloc(r->n_actions ? r->n_actions : 1);
        memcpy(r->actions, rules[i].actions, r->n_actions);
        if (rules[i].expression) {
            char msg[200];
            r->expr = bel_compile(rules[i].expression, msg, sizeof msg);
            if (!r->expr) {
                char full[260];
                snprintf(full, sizeof full, "error parsing rules: %s", msg); /* config.rs:268 */
                set_err(err, cap, full);
                orc_destroy(rs);
                return NULL;
            }
        }
    }
    return rs;
}

/* config_file.rs:257-265: a service's route is compiled with rules::compile_expression */
int orc_services_set(orc_ruleset* rs, const pgw_service_desc* sv, uint32_t n, char* err, size_t cap) {
    rs->services = (orc_rule*)calloc(n ? n : 1, sizeof(orc_rule));
    rs->n_services = n;
    for (uint32_t i = 0; i < n; ++i) {
        orc_rule* r = &rs->services[i];
        r->name = strdup(sv[i].name ? sv[i].name : "");
        if (sv[i].route) {
            char msg[200];
            r->expr = bel_compile(sv[i].route, msg, sizeof msg);
            if (!r->expr) {
                char full[300];
                snprintf(full, sizeof full, "error parsing route for service %s: %s", r->name, msg);
                set_err(err, cap, full);
                return 1;
            }
        }
    }
    return 0;
}

/* ---- lists (pingoo/lists.rs:62-113) ------------------------------------------------------ */
typedef struct {
    char** f;
    size_t* fl;
    int n, cap;
} record;

static void rec_push(record* r, const char* s, size_t n) {
    if (r->n == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 4;
        r->f = (char**)realloc(r->f, sizeof(char*) * (size_t)r->cap);
        r->fl = (size_t*)realloc(r->fl, sizeof(size_t) * (size_t)r->cap);
    }
    r->f[r->n] = (char*)malloc(n + 1);
    memcpy(r->f[r->n], s, n);
    r->f[r->n][n] = 0;
    r->fl[r->n] = n;
    r->n++;
}
static void rec_clear(record* r) {
    for (int i = 0; i < r->n; ++i) free(r->f[i]);
    r->n = 0;
}

/* csv::ReaderBuilder{has_headers:false, flexible:true}: RFC-4180 quoting, \n | \r\n | \r terminators,
 * blank lines skipped.  Returns 0 at end of input. */
static int csv_next(const uint8_t** pp, const uint8_t* end, record* rec) {
    const uint8_t* p = *pp;
    rec_clear(rec);
    while (p < end && (*p == '\n' || *p == '\r')) p++;
    if (p >= end) { *pp = p; return 0; }
    char* field = NULL;
    size_t fn = 0, fcap = 0;
    int quoted = 0, started = 0;
#define FPUSH(ch) do { if (fn + 1 >= fcap) { fcap = fcap ? fcap * 2 : 32; field = (char*)realloc(field, fcap); } field[fn++] = (char)(ch); } while (0)
    for (;;) {
        if (p >= end) { rec_push(rec, field ? field : "", fn); break; }
        uint8_t ch = *p;
        if (quoted) {
            if (ch == '"') {
                if (p + 1 < end && p[1] == '"') { FPUSH('"'); p += 2; continue; }
                quoted = 0;
                p++;
                continue;
            }
            FPUSH(ch);
            p++;
            continue;
        }
        if (ch == '"' && fn == 0 && !started) { quoted = 1; started = 1; p++; continue; }
        if (ch == ',') { rec_push(rec, field ? field : "", fn); fn = 0; started = 0; p++; continue; }
        if (ch == '\n' || ch == '\r') {
            rec_push(rec, field ? field : "", fn);
            if (ch == '\r' && p + 1 < end && p[1] == '\n') p++;
            p++;
            break;
        }
        FPUSH(ch);
        started = 1;
        p++;
    }
#undef FPUSH
    free(field);
    *pp = p;
    return 1;
}

int orc_lists_add(orc_ruleset* rs, const char* name, int type, const uint8_t* csv, size_t len, char* err, size_t cap) {
    bel_list L;
    memset(&L, 0, sizeof L);
    L.name = strdup(name);
    L.type = type;
    size_t capn = 0;
    record rec = {0};
    const uint8_t* p = csv;
    size_t line = 0;
    char msg[256];
    while (csv_next(&p, csv + len, &rec)) {
        line++;
        if (rec.n > 2 || rec.n < 1) {
            snprintf(msg, sizeof msg, "error parsing list %s at line %zu: invalid number of columns. Min: 1, Max: 2", name, line);
            set_err(err, cap, msg);
            return 1;
        }
        /* record[0].trim() */
        char* v = rec.f[0];
        size_t a = 0, b = rec.fl[0];
        while (a < b && (v[a] == ' ' || (v[a] >= 9 && v[a] <= 13))) a++;
        while (b > a && (v[b - 1] == ' ' || (v[b - 1] >= 9 && v[b - 1] <= 13))) b--;
        v[b] = 0;
        v += a;
        size_t vl = b - a;
        if (L.n == capn) {
            capn = capn ? capn * 2 : 64;
            L.strs = (char**)realloc(L.strs, sizeof(char*) * capn);
            L.str_lens = (size_t*)realloc(L.str_lens, sizeof(size_t) * capn);
            L.ints = (int64_t*)realloc(L.ints, sizeof(int64_t) * capn);
            L.nets = (bel_ipnet*)realloc(L.nets, sizeof(bel_ipnet) * capn);
        }
        L.strs[L.n] = NULL;
        if (type == BEL_LIST_STRING) {
            L.strs[L.n] = strdup(v);
            L.str_lens[L.n] = vl;
        } else if (type == BEL_LIST_INT) {
            size_t k = 0;
            int ok = vl > 0;
            if (ok && (v[0] == '+' || v[0] == '-')) k = 1;
            if (k >= vl) ok = 0;
            for (size_t j = k; ok && j < vl; ++j) if (!isdigit((unsigned char)v[j])) ok = 0;
            errno = 0;
            long long x = ok ? strtoll(v, NULL, 10) : 0;
            if (!ok || errno == ERANGE) {
                snprintf(msg, sizeof msg, "error parsing list %s at line %zu: error parsing int: invalid digit found in string", name, line);
                set_err(err, cap, msg);
                return 1;
            }
            L.ints[L.n] = (int64_t)x;
        } else {
            if (!bel_parse_ipnet(v, &L.nets[L.n])) {
                snprintf(msg, sizeof msg, "error parsing list %s at line %zu: error parsing IP network: invalid address: %s", name, line, v);
                set_err(err, cap, msg);
                return 1;
            }
        }
        L.n++;
    }
    rec_clear(&rec);
    free(rec.f);
    free(rec.fl);
    /* replace an existing list of the same name (HashMap insert) */
    for (size_t i = 0; i < rs->n_lists; ++i)
        if (!strcmp(rs->lists[i].name, name)) { rs->lists[i] = L; goto done; }
    rs->lists = (bel_list*)realloc(rs->lists, sizeof(bel_list) * (rs->n_lists + 1));
    rs->lists[rs->n_lists++] = L;
done:
    rs->lists_view.lists = rs->lists;
    rs->lists_view.n_lists = rs->n_lists;
    return 0;
}

/* ---- MaxMind DB -------------------------------------------------------------------------- */
typedef struct {
    int type;
    size_t size, next, ptr;
    int bad;
} mm_hdr;

static mm_hdr mm_header(const uint8_t* d, size_t n, size_t off) {
    mm_hdr h;
    memset(&h, 0, sizeof h);
    if (off >= n) { h.bad = 1; return h; }
    uint8_t ctrl = d[off++];
    int type = ctrl >> 5;
    if (type == 1) {
        int ss = (ctrl >> 3) & 3;
        if (off + (size_t)ss + 1 > n) { h.bad = 1; return h; }
        size_t v;
        if (ss == 0) v = ((size_t)(ctrl & 7) << 8) | d[off];
        else if (ss == 1) v = (((size_t)(ctrl & 7) << 16) | ((size_t)d[off] << 8) | d[off + 1]) + 2048;
        else if (ss == 2) v = (((size_t)(ctrl & 7) << 24) | ((size_t)d[off] << 16) | ((size_t)d[off + 1] << 8) | d[off + 2]) + 526336;
        else v = ((size_t)d[off] << 24) | ((size_t)d[off + 1] << 16) | ((size_t)d[off + 2] << 8) | d[off + 3];
        h.type = 1;
        h.ptr = v;
        h.next = off + (size_t)ss + 1;
        return h;
    }
    if (type == 0) {
        if (off >= n) { h.bad = 1; return h; }
        type = 7 + d[off++];
    }
    size_t size = ctrl & 0x1f;
    if (size == 29) { if (off + 1 > n) { h.bad = 1; return h; } size = 29 + d[off]; off += 1; }
    else if (size == 30) { if (off + 2 > n) { h.bad = 1; return h; } size = 285 + (((size_t)d[off] << 8) | d[off + 1]); off += 2; }
    else if (size == 31) { if (off + 3 > n) { h.bad = 1; return h; } size = 65821 + (((size_t)d[off] << 16) | ((size_t)d[off + 1] << 8) | d[off + 2]); off += 3; }
    h.type = type;
    h.size = size;
    h.next = off;
    return h;
}

/* offset just past the value at `off` (pointers are one value; they are not followed); (size_t)-1 on error */
static size_t mm_skip(const uint8_t* d, size_t n, size_t off, int depth) {
    mm_hdr h = mm_header(d, n, off);
    if (h.bad || depth > 32) return (size_t)-1;
    if (h.type == 1 || h.type == 14) return h.next;
    if (h.type == 7 || h.type == 11) {
        size_t p = h.next;
        size_t items = h.type == 7 ? h.size * 2 : h.size;
        for (size_t k = 0; k < items; ++k) {
            p = mm_skip(d, n, p, depth + 1);
            if (p == (size_t)-1) return p;
        }
        return p;
    }
    if (h.next + h.size > n) return (size_t)-1;
    return h.next + h.size;
}

/* read a UTF-8 string value (following pointers); *after = offset past the value as laid out at `off` */
static int mm_string(const uint8_t* d, size_t n, size_t off, const uint8_t** s, size_t* sl, size_t* after) {
    mm_hdr h = mm_header(d, n, off);
    if (h.bad) return 0;
    int was_ptr = h.type == 1;
    if (was_ptr) *after = h.next;
    int hops = 0;
    while (h.type == 1) {
        if (++hops > 8) return 0;
        h = mm_header(d, n, h.ptr);
        if (h.bad) return 0;
    }
    if (h.type != 2 || h.next + h.size > n) return 0;
    *s = d + h.next;
    *sl = h.size;
    if (!was_ptr) *after = h.next + h.size;
    return 1;
}

static int mm_open(orc_mmdb* m, char* err, size_t cap) {
    static const uint8_t marker[14] = {0xab, 0xcd, 0xef, 'M', 'a', 'x', 'M', 'i', 'n', 'd', '.', 'c', 'o', 'm'};
    if (m->len < 14) { set_err(err, cap, "mmdb file is not valid: metadata marker not found"); return 0; }
    size_t pos = (size_t)-1;
    for (size_t i = m->len - 14 + 1; i-- > 0;)
        if (!memcmp(m->buf + i, marker, 14)) { pos = i; break; }
    if (pos == (size_t)-1) { set_err(err, cap, "mmdb file is not valid: metadata marker not found"); return 0; }
    const uint8_t* d = m->buf + pos + 14;
    size_t n = m->len - pos - 14;
    mm_hdr h = mm_header(d, n, 0);
    if (h.bad || h.type != 7) { set_err(err, cap, "mmdb file is not valid: metadata is not a map"); return 0; }
    size_t p = h.next;
    int got = 0;
    for (size_t k = 0; k < h.size; ++k) {
        const uint8_t* ks;
        size_t kl, after;
        if (!mm_string(d, n, p, &ks, &kl, &after)) { set_err(err, cap, "mmdb file is not valid: bad metadata key"); return 0; }
        p = after;
        mm_hdr v = mm_header(d, n, p);
        if (v.bad) { set_err(err, cap, "mmdb file is not valid: bad metadata value"); return 0; }
        uint64_t x = 0;
        int is_uint = (v.type == 5 || v.type == 6 || v.type == 9) && v.size <= 8 && v.next + v.size <= n;
        if (is_uint) for (size_t j = 0; j < v.size; ++j) x = (x << 8) | d[v.next + j];
        if (kl == 10 && !memcmp(ks, "node_count", 10) && is_uint) { m->node_count = (uint32_t)x; got |= 1; }
        else if (kl == 11 && !memcmp(ks, "record_size", 11) && is_uint) { m->record_size = (uint32_t)x; got |= 2; }
        else if (kl == 10 && !memcmp(ks, "ip_version", 10) && is_uint) { m->ip_version = (uint32_t)x; got |= 4; }
        p = mm_skip(d, n, p, 0);
        if (p == (size_t)-1) { set_err(err, cap, "mmdb file is not valid: bad metadata value"); return 0; }
    }
    if (got != 7) { set_err(err, cap, "mmdb file is not valid: missing metadata fields"); return 0; }
    if (m->record_size != 24 && m->record_size != 28 && m->record_size != 32) { set_err(err, cap, "mmdb file is not valid: unsupported record size"); return 0; }
    if (m->ip_version != 4 && m->ip_version != 6) { set_err(err, cap, "mmdb file is not valid: bad ip_version"); return 0; }
    m->tree_bytes = (size_t)m->node_count * m->record_size / 4;
    if (m->tree_bytes + 16 > pos) { set_err(err, cap, "mmdb file is not valid: search tree larger than file"); return 0; }
    m->data_start = m->tree_bytes + 16;
    m->data_len = pos - m->data_start;
    return 1;
}

static uint32_t mm_record(const orc_mmdb* m, uint32_t node, int bit) {
    const uint8_t* b = m->buf + (size_t)node * m->record_size / 4;
    if (m->record_size == 24) { b += bit * 3; return (uint32_t)b[0] << 16 | (uint32_t)b[1] << 8 | b[2]; }
    if (m->record_size == 28) {
        if (!bit) return ((uint32_t)(b[3] & 0xF0) << 20) | (uint32_t)b[0] << 16 | (uint32_t)b[1] << 8 | b[2];
        return ((uint32_t)(b[3] & 0x0F) << 24) | (uint32_t)b[4] << 16 | (uint32_t)b[5] << 8 | b[6];
    }
    b += bit * 4;
    return (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
}

int orc_geoip_load(orc_ruleset* rs, const uint8_t* mmdb, size_t len, char* err, size_t cap) {
    orc_mmdb* m = (orc_mmdb*)calloc(1, sizeof *m);
    m->buf = (uint8_t*)malloc(len ? len : 1);
    memcpy(m->buf, mmdb, len);
    m->len = len;
    if (!mm_open(m, err, cap)) { free(m->buf); free(m); return 1; }
    /* ipv4_start: walk 96 zero bits of an IPv6 tree */
    m->v4_start = 0;
    if (m->ip_version == 6) {
        uint32_t node = 0;
        for (int i = 0; i < 96 && node < m->node_count; ++i) node = mm_record(m, node, 0);
        m->v4_start = node;
    }
    if (rs->geo) { free(rs->geo->buf); free(rs->geo); }
    rs->geo = m;
    return 0;
}

/* serde_utils::asn: trim every leading "AS", parse::<u32>() or 0 */
static uint32_t parse_asn(const uint8_t* s, size_t n) {
    while (n >= 2 && s[0] == 'A' && s[1] == 'S') { s += 2; n -= 2; }
    size_t k = 0;
    if (n && s[0] == '+') k = 1;
    if (k >= n) return 0;
    uint64_t v = 0;
    for (; k < n; ++k) {
        if (s[k] < '0' || s[k] > '9') return 0;
        v = v * 10 + (uint64_t)(s[k] - '0');
        if (v > 0xFFFFFFFFull) return 0;
    }
    return (uint32_t)v;
}

/* returns 1 and fills the record, or 0 for "any error / not found" (caller substitutes the default) */
static int mm_lookup(const orc_mmdb* m, const uint8_t* ip, int is_v6, uint32_t* asn, char country[2]) {
    if (m->node_count == 0) return 0;
    uint32_t node;
    int nbits;
    if (is_v6) {
        if (m->ip_version == 4) return 0; /* IPv6 address in an IPv4-only database */
        node = 0;
        nbits = 128;
    } else {
        node = m->ip_version == 6 ? m->v4_start : 0;
        nbits = 32;
    }
    for (int i = 0; i < nbits && node < m->node_count; ++i) {
        int bit = (ip[i >> 3] >> (7 - (i & 7))) & 1;
        node = mm_record(m, node, bit);
    }
    if (node == m->node_count) return 0; /* empty record: AddressNotFound */
    if (node < m->node_count) return 0;  /* ran out of bits on an internal node: invalid tree */
    size_t off = (size_t)node - m->node_count - 16;
    const uint8_t* d = m->buf + m->data_start;
    size_t n = m->data_len;
    if (off >= n) return 0;
    mm_hdr h = mm_header(d, n, off);
    int hops = 0;
    while (!h.bad && h.type == 1) {
        if (++hops > 8) return 0;
        h = mm_header(d, n, h.ptr);
    }
    if (h.bad || h.type != 7) return 0;
    size_t p = h.next;
    int have_asn = 0, have_cc = 0;
    uint32_t a = 0;
    char cc[2] = {'X', 'X'};
    for (size_t k = 0; k < h.size; ++k) {
        const uint8_t* ks;
        size_t kl, after;
        if (!mm_string(d, n, p, &ks, &kl, &after)) return 0;
        p = after;
        int is_asn = kl == 3 && !memcmp(ks, "asn", 3), is_cc = kl == 7 && !memcmp(ks, "country", 7);
        if (is_asn || is_cc) {
            const uint8_t* vs;
            size_t vl;
            if (!mm_string(d, n, p, &vs, &vl, &after)) return 0; /* not a string: serde type error */
            p = after;
            if (is_asn) {
                if (have_asn) return 0;
                have_asn = 1;
                a = parse_asn(vs, vl);
            } else {
                if (have_cc) return 0;
                have_cc = 1;
                if (vl != 2 || vs[0] < 'A' || vs[0] > 'Z' || vs[1] < 'A' || vs[1] > 'Z') return 0; /* geoip.rs:128-142 */
                cc[0] = (char)vs[0];
                cc[1] = (char)vs[1];
            }
        } else {
            p = mm_skip(d, n, p, 0);
            if (p == (size_t)-1) return 0;
        }
    }
    if (!have_asn || !have_cc) return 0;
    *asn = a;
    country[0] = cc[0];
    country[1] = cc[1];
    return 1;
}

void orc_geoip_lookup(const orc_ruleset* rs, const uint8_t ip[16], int is_v6, uint32_t* asn, uint16_t* country) {
    uint32_t a = 0;
    char cc[2] = {'X', 'X'};
    int skip; /* geoip.rs:74-76 */
    if (!is_v6) skip = ip[0] == 127 || (ip[0] & 0xF0) == 0xE0;
    else {
        int loop = ip[15] == 1;
        for (int k = 0; k < 15; ++k) if (ip[k]) loop = 0;
        skip = loop || ip[0] == 0xFF;
    }
    if (rs->geo && !skip) {
        uint32_t ta;
        char tc[2];
        if (mm_lookup(rs->geo, ip, is_v6, &ta, tc)) { a = ta; cc[0] = tc[0]; cc[1] = tc[1]; }
    }
    *asn = a;
    *country = (uint16_t)((uint8_t)cc[0] | ((uint8_t)cc[1] << 8));
}

/* ---- request loop -------------------------------------------------------------------------- */
static uint32_t verdict_for(const orc_ruleset* rs, const pgw_batch* b, uint32_t r, uint16_t* svc) {
    if (svc) *svc = (uint16_t)PGW_NO_SERVICE;
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    bel_ctx c;
    memset(&c, 0, sizeof c);
    for (int f = 0; f < 5; ++f) {
        if (cols[f]->offsets) {
            uint32_t a = cols[f]->offsets[r], e = cols[f]->offsets[r + 1];
            c.str[f] = cols[f]->bytes ? cols[f]->bytes + a : (const uint8_t*)"";
            c.len[f] = cols[f]->bytes ? e - a : 0;
        } else {
            c.str[f] = (const uint8_t*)"";
            c.len[f] = 0;
        }
    }
    uint32_t flags = b->flags ? b->flags[r] : 0;
    if (b->ip) memcpy(c.ip, b->ip + (size_t)r * 16, 16);
    c.ip_is_v6 = b->ip_is_v6 ? b->ip_is_v6[r] != 0 : 0;
    c.remote_port = b->remote_port ? b->remote_port[r] : 0;
    c.asn = 0;
    c.country[0] = 'X';
    c.country[1] = 'X';
    if (b->asn && b->country) {
        c.asn = b->asn[r];
        c.country[0] = (char)(b->country[r] & 0xFF);
        c.country[1] = (char)(b->country[r] >> 8);
    } else if (rs->geo && b->ip) {
        uint32_t a;
        uint16_t cc;
        orc_geoip_lookup(rs, c.ip, c.ip_is_v6, &a, &cc);
        c.asn = a;
        c.country[0] = (char)(cc & 0xFF);
        c.country[1] = (char)(cc >> 8);
    }
    c.lists = &rs->lists_view;

    /* gates, in the reference's order (http_listener.rs:196-236) */
    if (flags & PGW_FLAG_PRE_BLOCK) return PGW_BLOCK | (PGW_NO_RULE << 2);
    if (rs->eval_gates) {
        size_t ual = cols[4]->offsets ? cols[4]->offsets[r + 1] - cols[4]->offsets[r] : 0;
        if (ual == 0 || ual >= 256) return PGW_BLOCK | (PGW_NO_RULE << 2);
    }
    int bypass = (flags & PGW_FLAG_BYPASS) != 0;
    if (rs->eval_gates && c.len[2] >= 17 && !memcmp(c.str[2], "/__pingoo/captcha", 17)) bypass = 1;
    if (bypass) return PGW_BYPASS_CAPTCHA_API | (PGW_NO_RULE << 2);
    if (flags & PGW_FLAG_PRE_CAPTCHA) return PGW_CAPTCHA | (PGW_NO_RULE << 2);
    int captcha_verified = (flags & PGW_FLAG_CAPTCHA_VERIFIED) != 0;

    /* http_listener.rs:251-264 */
    for (uint32_t i = 0; i < rs->n_rules; ++i) {
        const orc_rule* rule = &rs->rules[i];
        int matched = rule->expr ? bel_matches(rule->expr, &c) : 1;
        if (!matched) continue;
        for (uint32_t k = 0; k < rule->n_actions; ++k) {
            if (rule->actions[k] == PGW_ACTION_BLOCK) return PGW_BLOCK | (i << 2);
            if (rule->actions[k] == PGW_ACTION_CAPTCHA && !captcha_verified) return PGW_CAPTCHA | (i << 2);
        }
    }
    /* http_listener.rs:266-272 + HttpService::match_request (http_proxy_service.rs:84-95): first service whose route
     * is absent or evaluates to true; none => 404 (PGW_NO_SERVICE) */
    if (svc)
        for (uint32_t i = 0; i < rs->n_services; ++i) {
            const orc_rule* sv = &rs->services[i];
            if (!sv->expr || bel_matches(sv->expr, &c)) { *svc = (uint16_t)i; break; }
        }
    return PGW_ALLOW | (PGW_NO_RULE << 2);
}

typedef struct {
    const orc_ruleset* rs;
    const pgw_batch* b;
    uint32_t* out;
    uint16_t* svc;
    uint32_t lo, hi;
} job;

static void* worker(void* arg) {
    job* j = (job*)arg;
    for (uint32_t r = j->lo; r < j->hi; ++r) j->out[r] = verdict_for(j->rs, j->b, r, j->svc ? j->svc + r : NULL);
    return NULL;
}

int orc_evaluate(const orc_ruleset* rs, const pgw_batch* batch, uint32_t* out, int n_threads) {
    return orc_evaluate_routed(rs, batch, out, NULL, n_threads);
}

int orc_evaluate_routed(const orc_ruleset* rs, const pgw_batch* batch, uint32_t* out, uint16_t* svc, int n_threads) {
    uint32_t n = batch->n;
    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > n) n_threads = n ? (int)n : 1;
    if (n_threads == 1) {
        job j = {rs, batch, out, svc, 0, n};
        worker(&j);
        return 0;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job* jobs = (job*)malloc(sizeof(job) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].rs = rs;
        jobs[t].b = batch;
        jobs[t].out = out;
        jobs[t].svc = svc;
        jobs[t].lo = (uint32_t)((uint64_t)n * (uint64_t)t / (uint64_t)n_threads);
        jobs[t].hi = (uint32_t)((uint64_t)n * (uint64_t)(t + 1) / (uint64_t)n_threads);
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}

void orc_destroy(orc_ruleset* rs) {
    if (!rs) return;
    for (uint32_t i = 0; i < rs->n_rules; ++i) {
        free(rs->rules[i].name);
        free(rs->rules[i].actions);
        bel_free(rs->rules[i].expr);
    }
    free(rs->rules);
    for (uint32_t i = 0; i < rs->n_services; ++i) {
        free(rs->services[i].name);
        bel_free(rs->services[i].expr);
    }
    free(rs->services);
    for (size_t i = 0; i < rs->n_lists; ++i) {
        bel_list* L = &rs->lists[i];
        if (L->strs) for (size_t k = 0; k < L->n; ++k) free(L->strs[k]);
        free(L->strs); free(L->str_lens); free(L->ints); free(L->nets); free(L->name);
    }
    free(rs->lists);
    if (rs->geo) { free(rs->geo->buf); free(rs->geo); }
    free(rs);
}

/* ---- hooks for differential tests -------------------------------------------------------------- */
int orc_regex_is_match(const char* pattern, size_t plen, const uint8_t* hay, size_t n) {
    int st;
    char msg[128];
    rx_prog* p = rx_compile(pattern, plen, &st, msg);
    if (!p) return -st;
    int r = rx_is_match(p, hay, n);
    rx_free(p);
    return r;
}

int orc_ipnet_contains(const char* net, const uint8_t ip[16], int is_v6) {
    bel_ipnet n;
    if (!bel_parse_ipnet(net, &n)) return -1;
    return bel_ipnet_contains(&n, ip, is_v6);
}

int orc_eval_kind(const char* expr, const pgw_batch* b) {
    char msg[200];
    bel_expr* e = bel_compile(expr, msg, sizeof msg);
    if (!e) return -1;
    /* build the context exactly as verdict_for does, without lists/geo */
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    bel_ctx c;
    memset(&c, 0, sizeof c);
    for (int f = 0; f < 5; ++f) {
        c.str[f] = cols[f]->bytes ? cols[f]->bytes + cols[f]->offsets[0] : (const uint8_t*)"";
        c.len[f] = cols[f]->bytes ? cols[f]->offsets[1] - cols[f]->offsets[0] : 0;
    }
    if (b->ip) memcpy(c.ip, b->ip, 16);
    c.ip_is_v6 = b->ip_is_v6 ? b->ip_is_v6[0] : 0;
    c.remote_port = b->remote_port ? b->remote_port[0] : 0;
    c.asn = b->asn ? b->asn[0] : 0;
    c.country[0] = b->country ? (char)(b->country[0] & 0xFF) : 'X';
    c.country[1] = b->country ? (char)(b->country[0] >> 8) : 'X';
    bel_lists none = {NULL, 0};
    c.lists = &none;
    int r = bel_eval_kind(e, &c);
    bel_free(e);
    return r;
}
