// Flatten IP lists (pingoo/lists.rs:102, bel List<IpNetwork>.contains(ip)) and
// the GeoIP database (pingoo/geoip.rs:73-91 -> maxminddb 0.24 Reader::lookup)
// into one longest-prefix structure:
//   IPv4: DIR-24-8 (dir24[ip>>8], optional 256-entry tbl8 block)
//   IPv6: sorted disjoint ranges, binary searched
// Every entry resolves to a leaf {asn, country, bitmask of ip sets containing the address}.
#include <algorithm>
#include <cstring>
#include <map>
#include <unordered_map>

#include "compile.hpp"

namespace pgw {
namespace {

typedef unsigned __int128 u128;

struct Mmdb {
    const uint8_t* buf = nullptr;
    size_t len = 0;
    uint32_t node_count = 0;
    uint32_t record_size = 0;
    uint32_t ip_version = 0;
    size_t tree_bytes = 0;
    size_t data_start = 0;  // offset of the data section
    size_t data_len = 0;
};

// ---- MaxMind DB data-section decoding (format spec v2.0) -------------------------
struct Dec {
    const uint8_t* d;  // data section
    size_t n;
    bool fail = false;

    struct Hdr { int type; size_t size; size_t next; size_t ptr; };

    Hdr header(size_t off) {
        Hdr h{0, 0, 0, 0};
        if (off >= n) { fail = true; return h; }
        uint8_t ctrl = d[off++];
        int type = ctrl >> 5;
        if (type == 1) {  // pointer
            int ss = (ctrl >> 3) & 3;
            if (off + ss + 1 > n) { fail = true; return h; }
            size_t v = 0;
            switch (ss) {
                case 0: v = ((size_t)(ctrl & 7) << 8) | d[off]; break;
                case 1: v = (((size_t)(ctrl & 7) << 16) | ((size_t)d[off] << 8) | d[off + 1]) + 2048; break;
                case 2: v = (((size_t)(ctrl & 7) << 24) | ((size_t)d[off] << 16) | ((size_t)d[off + 1] << 8) | d[off + 2]) + 526336; break;
                default: v = ((size_t)d[off] << 24) | ((size_t)d[off + 1] << 16) | ((size_t)d[off + 2] << 8) | d[off + 3]; break;
            }
            h.type = 1;
            h.ptr = v;
            h.next = off + ss + 1;
            return h;
        }
        if (type == 0) {
            if (off >= n) { fail = true; return h; }
            type = 7 + d[off++];
        }
        size_t size = ctrl & 0x1f;
        if (size == 29) { if (off + 1 > n) { fail = true; return h; } size = 29 + d[off]; off += 1; }
        else if (size == 30) { if (off + 2 > n) { fail = true; return h; } size = 285 + (((size_t)d[off] << 8) | d[off + 1]); off += 2; }
        else if (size == 31) { if (off + 3 > n) { fail = true; return h; } size = 65821 + (((size_t)d[off] << 16) | ((size_t)d[off + 1] << 8) | d[off + 2]); off += 3; }
        h.type = type;
        h.size = size;
        h.next = off;
        return h;
    }

    // resolve pointers; returns header of the pointed-to value and its offset
    Hdr resolve(size_t off, size_t* after) {
        Hdr h = header(off);
        if (fail) return h;
        if (after) *after = h.next;
        int hops = 0;
        while (h.type == 1) {
            if (++hops > 8) { fail = true; return h; }
            h = header(h.ptr);
            if (fail) return h;
        }
        return h;
    }

    // skip a value starting at `off`, returning the offset just past it (pointers are not followed)
    size_t skip(size_t off, int depth = 0) {
        Hdr h = header(off);
        if (fail || depth > 32) { fail = true; return off; }
        switch (h.type) {
            case 1: return h.next;
            case 7: {  // map
                size_t p = h.next;
                for (size_t k = 0; k < h.size && !fail; ++k) { p = skip(p, depth + 1); p = skip(p, depth + 1); }
                return p;
            }
            case 11: {  // array
                size_t p = h.next;
                for (size_t k = 0; k < h.size && !fail; ++k) p = skip(p, depth + 1);
                return p;
            }
            case 14: return h.next;  // boolean: size is the value
            default:
                if (h.next + h.size > n) { fail = true; return off; }
                return h.next + h.size;
        }
    }

    bool read_string(size_t off, std::string* out, size_t* after) {
        Hdr h = resolve(off, after);
        if (fail || h.type != 2 || h.next + h.size > n) return false;
        out->assign((const char*)d + h.next, h.size);
        if (after && header(off).type != 1) *after = h.next + h.size;
        return true;
    }
};

// serde_utils::asn (pingoo/serde_utils.rs:1-9): strip every leading "AS", parse u32, else 0
uint32_t parse_asn(const std::string& s) {
    size_t p = 0;
    while (s.compare(p, 2, "AS") == 0) p += 2;
    std::string t = s.substr(p);
    size_t k = 0;
    if (!t.empty() && t[0] == '+') k = 1;
    if (k >= t.size()) return 0;
    uint64_t v = 0;
    for (; k < t.size(); ++k) {
        if (!isdigit((unsigned char)t[k])) return 0;
        v = v * 10 + (t[k] - '0');
        if (v > 0xFFFFFFFFull) return 0;
    }
    return (uint32_t)v;
}

// Decode the record the reference deserialises into GeoipRecord{asn, country} (geoip.rs:17-23).
// Any decoding error makes the whole lookup fall back to the default record (http_listener.rs:143-157).
bool decode_record(Dec& dec, size_t off, GeoRecord* out) {
    dec.fail = false;
    Dec::Hdr h = dec.resolve(off, nullptr);
    if (dec.fail || h.type != 7) return false;
    size_t p = h.next;
    bool have_asn = false, have_country = false;
    GeoRecord r;
    for (size_t k = 0; k < h.size; ++k) {
        std::string key;
        size_t after = 0;
        if (!dec.read_string(p, &key, &after)) return false;
        p = after;
        if (key == "asn" || key == "country") {
            std::string val;
            size_t after_v = 0;
            if (!dec.read_string(p, &val, &after_v)) return false;  // wrong type -> serde error
            p = after_v;
            if (key == "asn") {
                if (have_asn) return false;  // duplicate field
                have_asn = true;
                r.asn = parse_asn(val);
            } else {
                if (have_country) return false;
                have_country = true;
                if (val.size() != 2 || val[0] < 'A' || val[0] > 'Z' || val[1] < 'A' || val[1] > 'Z') return false;  // geoip.rs:128-142
                r.country[0] = val[0];
                r.country[1] = val[1];
            }
        } else {
            p = dec.skip(p);
            if (dec.fail) return false;
        }
    }
    if (!have_asn || !have_country) return false;  // serde: missing field
    *out = r;
    return true;
}

bool open_mmdb(const std::vector<uint8_t>& buf, Mmdb* m, std::string& err) {
    static const uint8_t marker[] = {0xab, 0xcd, 0xef, 'M', 'a', 'x', 'M', 'i', 'n', 'd', '.', 'c', 'o', 'm'};
    const size_t ml = sizeof marker;
    if (buf.size() < ml) { err = "mmdb file is not valid: metadata marker not found"; return false; }
    size_t pos = std::string::npos;
    for (size_t i = buf.size() - ml + 1; i-- > 0;) {
        if (memcmp(buf.data() + i, marker, ml) == 0) { pos = i; break; }
    }
    if (pos == std::string::npos) { err = "mmdb file is not valid: metadata marker not found"; return false; }
    Dec md{buf.data() + pos + ml, buf.size() - pos - ml};
    Dec::Hdr h = md.header(0);
    if (md.fail || h.type != 7) { err = "mmdb file is not valid: metadata is not a map"; return false; }
    size_t p = h.next;
    bool got_nc = false, got_rs = false, got_iv = false;
    for (size_t k = 0; k < h.size; ++k) {
        std::string key;
        size_t after = 0;
        if (!md.read_string(p, &key, &after)) { err = "mmdb file is not valid: bad metadata key"; return false; }
        p = after;
        Dec::Hdr v = md.header(p);
        if (md.fail) { err = "mmdb file is not valid: bad metadata value"; return false; }
        auto read_uint = [&](uint64_t* out) {
            if ((v.type != 5 && v.type != 6 && v.type != 9) || v.size > 8 || v.next + v.size > md.n) return false;
            uint64_t x = 0;
            for (size_t j = 0; j < v.size; ++j) x = (x << 8) | md.d[v.next + j];
            *out = x;
            return true;
        };
        uint64_t x = 0;
        if (key == "node_count") { if (!read_uint(&x)) { err = "mmdb file is not valid: node_count"; return false; } m->node_count = (uint32_t)x; got_nc = true; }
        else if (key == "record_size") { if (!read_uint(&x)) { err = "mmdb file is not valid: record_size"; return false; } m->record_size = (uint32_t)x; got_rs = true; }
        else if (key == "ip_version") { if (!read_uint(&x)) { err = "mmdb file is not valid: ip_version"; return false; } m->ip_version = (uint32_t)x; got_iv = true; }
        p = md.skip(p);
        if (md.fail) { err = "mmdb file is not valid: bad metadata value"; return false; }
    }
    if (!got_nc || !got_rs || !got_iv) { err = "mmdb file is not valid: missing metadata fields"; return false; }
    if (m->record_size != 24 && m->record_size != 28 && m->record_size != 32) { err = "mmdb file is not valid: unsupported record size"; return false; }
    if (m->ip_version != 4 && m->ip_version != 6) { err = "mmdb file is not valid: bad ip_version"; return false; }
    m->buf = buf.data();
    m->len = buf.size();
    m->tree_bytes = (size_t)m->node_count * m->record_size / 4;
    if (m->tree_bytes + 16 > pos) { err = "mmdb file is not valid: search tree larger than file"; return false; }
    m->data_start = m->tree_bytes + 16;
    m->data_len = pos - m->data_start;
    return true;
}

inline uint32_t read_record(const Mmdb& m, uint32_t node, int bit) {
    const uint8_t* b = m.buf + (size_t)node * m.record_size / 4;
    switch (m.record_size) {
        case 24: b += bit * 3; return (uint32_t)b[0] << 16 | (uint32_t)b[1] << 8 | b[2];
        case 28:
            if (bit == 0) return ((uint32_t)(b[3] & 0xF0) << 20) | (uint32_t)b[0] << 16 | (uint32_t)b[1] << 8 | b[2];
            return ((uint32_t)(b[3] & 0x0F) << 24) | (uint32_t)b[4] << 16 | (uint32_t)b[5] << 8 | b[6];
        default: b += bit * 4; return (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
    }
}

struct GeoLeaves {
    // terminal records in address order: start address, depth, geo record index (0 = default)
    std::vector<u128> start;
    std::vector<uint8_t> depth;
    std::vector<uint32_t> geo;
};

struct GeoBuild {
    const Mmdb& m;
    Dec dec;
    std::vector<GeoRecord> records;                 // [0] = default
    std::unordered_map<uint32_t, uint32_t> by_ptr;  // record value -> geo index
    std::map<std::pair<uint32_t, uint16_t>, uint32_t> by_val;

    explicit GeoBuild(const Mmdb& mm) : m(mm), dec{mm.buf + mm.data_start, mm.data_len} { records.push_back(GeoRecord()); }

    uint32_t geo_of(uint32_t rec) {
        if (rec == m.node_count) return 0;  // empty: AddressNotFound -> default
        auto it = by_ptr.find(rec);
        if (it != by_ptr.end()) return it->second;
        uint32_t idx = 0;
        size_t off = (size_t)rec - m.node_count - 16;
        GeoRecord r;
        if (rec > m.node_count && off < m.data_len && decode_record(dec, off, &r)) {
            auto key = std::make_pair(r.asn, (uint16_t)((uint8_t)r.country[0] | (uint8_t)r.country[1] << 8));
            auto jt = by_val.find(key);
            if (jt != by_val.end()) idx = jt->second;
            else {
                idx = (uint32_t)records.size();
                records.push_back(r);
                by_val[key] = idx;
            }
        }
        by_ptr[rec] = idx;
        return idx;
    }

    // enumerate terminals below `node` (which sits at `depth` bits with prefix `base` in a `width`-bit space)
    void walk(uint32_t root_rec, int width, GeoLeaves* out) {
        struct Item { uint32_t rec; int depth; u128 base; };
        std::vector<Item> st;
        st.push_back(Item{root_rec, 0, 0});
        while (!st.empty()) {
            Item it = st.back();
            st.pop_back();
            if (it.rec >= m.node_count || it.depth == width) {
                // terminal (a node reached at full depth is malformed: treat as not found)
                out->start.push_back(it.base);
                out->depth.push_back((uint8_t)it.depth);
                out->geo.push_back(it.rec >= m.node_count ? geo_of(it.rec) : 0);
                continue;
            }
            uint32_t l = read_record(m, it.rec, 0), r = read_record(m, it.rec, 1);
            u128 bitv = (u128)1 << (width - 1 - it.depth);
            st.push_back(Item{r, it.depth + 1, it.base | bitv});  // right pushed first: left pops first (address order)
            st.push_back(Item{l, it.depth + 1, it.base});
        }
    }
};

}  // namespace

bool build_lpm(const std::vector<std::vector<IpNet>>& ip_sets, const std::vector<uint8_t>& geo_mmdb, LpmTables* out,
               std::string& err) {
    LpmTables& T = *out;
    T = LpmTables();
    T.present = true;

    GeoLeaves v4geo, v6geo;
    std::vector<GeoRecord> georecs(1);
    if (!geo_mmdb.empty()) {
        Mmdb m;
        if (!open_mmdb(geo_mmdb, &m, err)) return false;
        GeoBuild gb(m);
        // IPv4 addresses in an IPv6 tree live under ::/96 (maxminddb: ipv4_start)
        uint32_t v4root = 0;
        bool v4root_is_node = m.node_count > 0;
        if (m.node_count == 0) { v4root = 0; v4root_is_node = false; }
        uint32_t rec = 0;  // node 0
        if (m.ip_version == 6) {
            rec = 0;
            for (int i = 0; i < 96 && rec < m.node_count; ++i) rec = read_record(m, rec, 0);
            v4root = rec;
            v4root_is_node = rec < m.node_count;
        }
        if (m.node_count == 0) {
            v4geo.start.push_back(0); v4geo.depth.push_back(0); v4geo.geo.push_back(0);
        } else if (m.ip_version == 4 || v4root_is_node) {
            gb.walk(m.ip_version == 4 ? 0 : v4root, 32, &v4geo);
        } else {
            // the walk ended before 96 bits: one record (or none) covers all of IPv4
            v4geo.start.push_back(0); v4geo.depth.push_back(0); v4geo.geo.push_back(gb.geo_of(v4root));
        }
        if (m.ip_version == 6 && m.node_count > 0) gb.walk(0, 128, &v6geo);
        else { v6geo.start.push_back(0); v6geo.depth.push_back(0); v6geo.geo.push_back(0); }  // IPv6 in a v4-only db: not found
        georecs = gb.records;
        T.geo_loaded = true;
    } else {
        v4geo.start.push_back(0); v4geo.depth.push_back(0); v4geo.geo.push_back(0);
        v6geo.start.push_back(0); v6geo.depth.push_back(0); v6geo.geo.push_back(0);
    }

    // ---- leaves interning -----------------------------------------------------------------
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> leaf_of;  // (geo idx, mask) -> leaf id
    auto leaf = [&](uint32_t geo, uint32_t mask) -> uint32_t {
        auto key = std::make_pair(geo, mask);
        auto it = leaf_of.find(key);
        if (it != leaf_of.end()) return it->second;
        LpmLeaf l;
        l.asn = georecs[geo].asn;
        l.country = (uint16_t)((uint8_t)georecs[geo].country[0] | ((uint8_t)georecs[geo].country[1] << 8));
        l.pad = 0;
        l.set_mask = mask;
        T.leaves.push_back(l);
        leaf_of[key] = (uint32_t)T.leaves.size() - 1;
        return (uint32_t)T.leaves.size() - 1;
    };
    leaf(0, 0);  // leaf 0 = default record, no sets

    // ---- IPv4: paint (geo, mask) over the /24 grid, then refine with tbl8 blocks ----------
    const size_t N24 = 1u << 24;
    std::vector<uint32_t> g24(N24, 0), m24(N24, 0);
    struct Long { uint32_t addr; int prefix; uint32_t geo; uint32_t mask; bool is_geo; };
    std::vector<Long> longs;
    for (size_t i = 0; i < v4geo.start.size(); ++i) {
        uint32_t a = (uint32_t)v4geo.start[i];
        int d = v4geo.depth[i];
        if (v4geo.geo[i] == 0) continue;
        if (d <= 24) {
            size_t cnt = (size_t)1 << (24 - d);
            std::fill(g24.begin() + (a >> 8), g24.begin() + (a >> 8) + cnt, v4geo.geo[i]);
        } else longs.push_back(Long{a, d, v4geo.geo[i], 0, true});
    }
    for (size_t s = 0; s < ip_sets.size(); ++s)
        for (const IpNet& n : ip_sets[s]) {
            if (n.v6) continue;
            uint32_t a = (uint32_t)n.addr[0] << 24 | (uint32_t)n.addr[1] << 16 | (uint32_t)n.addr[2] << 8 | n.addr[3];
            uint32_t mask = n.prefix == 0 ? 0 : ~0u << (32 - n.prefix);
            a &= mask;  // ipnetwork contains(): compares masked addresses
            if (n.prefix <= 24) {
                size_t cnt = (size_t)1 << (24 - n.prefix);
                for (size_t k = 0; k < cnt; ++k) m24[(a >> 8) + k] |= 1u << s;
            } else longs.push_back(Long{a, n.prefix, 0, 1u << s, false});
        }
    std::unordered_map<uint32_t, uint32_t> block_of;  // /24 index -> tbl8 block
    std::vector<uint32_t> g8, m8;
    for (const Long& L : longs) {
        uint32_t top = L.addr >> 8;
        auto it = block_of.find(top);
        uint32_t b;
        if (it == block_of.end()) {
            b = (uint32_t)(g8.size() / 256);
            block_of[top] = b;
            g8.insert(g8.end(), 256, g24[top]);
            m8.insert(m8.end(), 256, m24[top]);
        } else b = it->second;
        uint32_t lo = L.addr & 0xFF, cnt = 1u << (32 - L.prefix);
        for (uint32_t k = 0; k < cnt; ++k) {
            if (L.is_geo) g8[b * 256 + lo + k] = L.geo;
            else m8[b * 256 + lo + k] |= L.mask;
        }
    }
    T.dir24.resize(N24);
    for (size_t i = 0; i < N24; ++i) T.dir24[i] = leaf(g24[i], m24[i]);
    T.tbl8.resize(g8.size());
    for (size_t i = 0; i < g8.size(); ++i) T.tbl8[i] = leaf(g8[i], m8[i]);
    for (auto& kv : block_of) T.dir24[kv.first] = 0x80000000u | kv.second;
    if (T.tbl8.empty()) T.tbl8.push_back(0);

    // ---- IPv6: elementary ranges from geo terminals and set prefixes ------------------------
    struct Ev { u128 at; int set; int delta; };
    std::vector<Ev> evs;
    std::vector<u128> cuts;
    for (size_t i = 0; i < v6geo.start.size(); ++i) cuts.push_back(v6geo.start[i]);
    for (size_t s = 0; s < ip_sets.size(); ++s)
        for (const IpNet& n : ip_sets[s]) {
            if (!n.v6) continue;
            u128 a = 0;
            for (int k = 0; k < 16; ++k) a = (a << 8) | n.addr[k];
            u128 mask = n.prefix == 0 ? 0 : ~(u128)0 << (128 - n.prefix);
            a &= mask;
            u128 last = a | ~mask;
            evs.push_back(Ev{a, (int)s, +1});
            cuts.push_back(a);
            if (last != ~(u128)0) {
                evs.push_back(Ev{last + 1, (int)s, -1});
                cuts.push_back(last + 1);
            }
        }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    std::sort(evs.begin(), evs.end(), [](const Ev& a, const Ev& b) { return a.at < b.at; });
    std::vector<int> cnt(ip_sets.size(), 0);
    size_t ei = 0, gi = 0;
    uint32_t prev_leaf = 0xFFFFFFFFu;
    for (u128 c : cuts) {
        while (ei < evs.size() && evs[ei].at <= c) { cnt[evs[ei].set] += evs[ei].delta; ++ei; }
        while (gi + 1 < v6geo.start.size() && v6geo.start[gi + 1] <= c) ++gi;
        uint32_t mask = 0;
        for (size_t s = 0; s < cnt.size(); ++s) if (cnt[s] > 0) mask |= 1u << s;
        uint32_t lf = leaf(v6geo.geo[gi], mask);
        if (lf == prev_leaf) continue;  // merge equal neighbours
        prev_leaf = lf;
        T.v6_hi.push_back((uint64_t)(c >> 64));
        T.v6_lo.push_back((uint64_t)c);
        T.v6_leaf.push_back(lf);
    }
    if (T.v6_leaf.empty()) { T.v6_hi.push_back(0); T.v6_lo.push_back(0); T.v6_leaf.push_back(0); }
    // top-level index on the first 16 address bits: the range holding the first address of bucket t.  A lookup then
    // searches ranges [v6_top[t], v6_top[t + 1]] only -- a handful instead of log2(all ranges) dependent probes.
    T.v6_top.assign(65537, 0);
    {
        size_t i = 0;
        const size_t n = T.v6_leaf.size();
        for (uint32_t t = 0; t < 65536; ++t) {
            const uint64_t key = (uint64_t)t << 48;
            while (i + 1 < n && (T.v6_hi[i + 1] < key || (T.v6_hi[i + 1] == key && T.v6_lo[i + 1] == 0))) ++i;
            T.v6_top[t] = (uint32_t)i;
        }
        T.v6_top[65536] = (uint32_t)(n - 1);
    }
    return true;
}

}  // namespace pgw
