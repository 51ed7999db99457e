// List loading: CSV -> typed vectors, with the reference's validation and
// messages (pingoo/lists.rs:62-113), and IpNetwork parsing (ipnetwork 0.21
// FromStr: bare address = host prefix, "addr/len", and IPv4 "addr/netmask").
#include <cerrno>
#include <cstdlib>
#include <cstring>

#include "compile.hpp"

namespace pgw {
namespace {

bool parse_ipv4(const std::string& s, uint8_t out[4]) {
    // std::net::Ipv4Addr::from_str: exactly four decimal octets, no leading zeros, each <= 255
    size_t p = 0;
    for (int k = 0; k < 4; ++k) {
        if (p >= s.size() || !isdigit((unsigned char)s[p])) return false;
        size_t st = p;
        unsigned v = 0;
        while (p < s.size() && isdigit((unsigned char)s[p])) {
            v = v * 10 + (s[p] - '0');
            if (v > 255 || p - st >= 3) return false;
            ++p;
        }
        if (p - st > 1 && s[st] == '0') return false;
        out[k] = (uint8_t)v;
        if (k < 3) {
            if (p >= s.size() || s[p] != '.') return false;
            ++p;
        }
    }
    return p == s.size();
}

bool parse_ipv6(const std::string& s, uint8_t out[16]) {
    // std::net::Ipv6Addr::from_str: up to 8 groups of 1-4 hex digits, one "::", optional trailing IPv4
    uint16_t head[8], tail[8];
    int nh = 0, nt = 0;
    bool seen_gap = false;
    size_t p = 0;
    if (s.size() >= 2 && s[0] == ':' && s[1] == ':') { seen_gap = true; p = 2; }
    else if (!s.empty() && s[0] == ':') return false;
    bool expect_group = p < s.size();
    while (p < s.size()) {
        // try embedded IPv4 at the tail
        size_t q = p;
        while (q < s.size() && s[q] != ':') ++q;
        std::string tok = s.substr(p, q - p);
        if (tok.find('.') != std::string::npos) {
            if (q != s.size()) return false;
            uint8_t v4[4];
            if (!parse_ipv4(tok, v4)) return false;
            uint16_t a = (uint16_t)(v4[0] << 8 | v4[1]), b = (uint16_t)(v4[2] << 8 | v4[3]);
            if (seen_gap) { if (nt > 6) return false; tail[nt++] = a; tail[nt++] = b; }
            else { if (nh > 6) return false; head[nh++] = a; head[nh++] = b; }
            p = q;
            expect_group = false;
            break;
        }
        if (tok.empty() || tok.size() > 4) return false;
        unsigned v = 0;
        for (char c : tok) {
            int h = isdigit((unsigned char)c) ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1;
            if (h < 0) return false;
            v = v * 16 + h;
        }
        if (seen_gap) { if (nt >= 8) return false; tail[nt++] = (uint16_t)v; }
        else { if (nh >= 8) return false; head[nh++] = (uint16_t)v; }
        p = q;
        expect_group = false;
        if (p < s.size()) {
            // s[p] == ':'
            if (p + 1 < s.size() && s[p + 1] == ':') {
                if (seen_gap) return false;
                seen_gap = true;
                p += 2;
                expect_group = false;
            } else {
                ++p;
                expect_group = true;
            }
        }
    }
    if (expect_group) return false;
    int total = nh + nt;
    if (seen_gap) { if (total > 7) return false; }
    else if (total != 8) return false;
    uint16_t g[8] = {0};
    for (int k = 0; k < nh; ++k) g[k] = head[k];
    for (int k = 0; k < nt; ++k) g[8 - nt + k] = tail[k];
    for (int k = 0; k < 8; ++k) { out[2 * k] = (uint8_t)(g[k] >> 8); out[2 * k + 1] = (uint8_t)g[k]; }
    return true;
}

}  // namespace

// `IpNetwork::from_str` (ipnetwork 0.21, reached by lists.rs:100 `item_value.parse()`) tries the IPv4 and then the IPv6 network
// parser and reports ANY failure -- address, prefix or netmask -- as InvalidAddr(<the whole string>): "invalid address: <s>".
bool parse_ip_network(const std::string& s, IpNet* out, std::string& err) {
    IpNet n;
    size_t slash = s.find('/');
    std::string addr = slash == std::string::npos ? s : s.substr(0, slash);
    std::string pfx = slash == std::string::npos ? "" : s.substr(slash + 1);
    if (parse_ipv4(addr, n.addr)) {
        n.v6 = false;
        n.prefix = 32;
    } else if (parse_ipv6(addr, n.addr)) {
        n.v6 = true;
        n.prefix = 128;
    } else {
        err = "invalid address: " + s;
        return false;
    }
    if (slash != std::string::npos) {
        // the prefix goes through `str::parse::<u8>()` (ipnetwork parse_prefix): an optional '+', digits (leading zeros allowed),
        // a value that fits the type; then <= 32 / 128
        const size_t d0 = (!pfx.empty() && pfx[0] == '+') ? 1 : 0;
        bool numeric = pfx.size() > d0;
        for (size_t k = d0; k < pfx.size(); ++k) if (!isdigit((unsigned char)pfx[k])) numeric = false;
        if (numeric) {
            unsigned v = 0;
            for (size_t k = d0; k < pfx.size() && v <= 255; ++k) v = v * 10 + (unsigned)(pfx[k] - '0');
            if (v > (n.v6 ? 128u : 32u)) { err = "invalid address: " + s; return false; }
            n.prefix = (int)v;
        } else if (!n.v6) {
            // dotted netmask, must be contiguous ones
            uint8_t m[4];
            if (!parse_ipv4(pfx, m)) { err = "invalid address: " + s; return false; }
            uint32_t mask = (uint32_t)m[0] << 24 | (uint32_t)m[1] << 16 | (uint32_t)m[2] << 8 | m[3];
            int len = 0;
            while (len < 32 && (mask & (0x80000000u >> len))) ++len;
            if (len < 32 && (mask << len) != 0) { err = "invalid address: " + s; return false; }
            n.prefix = len;
        } else {
            err = "invalid address: " + s;
            return false;
        }
    }
    *out = n;
    return true;
}

namespace {

// Minimal RFC-4180 reader matching csv::ReaderBuilder{has_headers:false, flexible:true}:
// ',' delimiter, '"' quoting with "" escapes, records end at \n, \r\n or \r; empty lines are skipped.
struct CsvReader {
    const uint8_t* p;
    const uint8_t* end;
    // returns false at end of input
    bool next(std::vector<std::string>& rec) {
        rec.clear();
        for (;;) {
            // skip empty lines
            while (p < end && (*p == '\n' || *p == '\r')) ++p;
            if (p >= end) return false;
            break;
        }
        std::string field;
        bool in_quotes = false, any = false;
        for (;;) {
            if (p >= end) {
                rec.push_back(field);
                return true;
            }
            uint8_t c = *p;
            if (in_quotes) {
                if (c == '"') {
                    if (p + 1 < end && p[1] == '"') { field.push_back('"'); p += 2; continue; }
                    in_quotes = false;
                    ++p;
                    continue;
                }
                field.push_back((char)c);
                ++p;
                continue;
            }
            if (c == '"' && field.empty() && !any) { in_quotes = true; any = true; ++p; continue; }
            if (c == ',') { rec.push_back(field); field.clear(); any = false; ++p; continue; }
            if (c == '\n' || c == '\r') {
                rec.push_back(field);
                if (c == '\r' && p + 1 < end && p[1] == '\n') ++p;
                ++p;
                return true;
            }
            field.push_back((char)c);
            any = true;
            ++p;
        }
    }
};

std::string trim(const std::string& s) {
    // str::trim: Unicode White_Space; ASCII subset is enough for list files
    size_t a = 0, b = s.size();
    auto ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
    while (a < b && ws(s[a])) ++a;
    while (b > a && ws(s[b - 1])) --b;
    return s.substr(a, b - a);
}

}  // namespace

bool parse_list_csv(const std::string& name, ListType type, const uint8_t* csv, size_t len, ListData* out, std::string& err) {
    ListData L;
    L.type = type;
    CsvReader rd{csv, csv + len};
    std::vector<std::string> rec;
    size_t line = 0;
    while (rd.next(rec)) {
        ++line;
        if (rec.size() > 2 || rec.size() < 1) {
            err = "error parsing list " + name + " at line " + std::to_string(line) + ": invalid number of columns. Min: 1, Max: 2";
            return false;
        }
        std::string v = trim(rec[0]);
        switch (type) {
            case LT_STRING: L.strs.push_back(v); break;
            case LT_INT: {
                // i64::from_str: optional sign, decimal digits only
                bool ok = !v.empty();
                size_t k = 0;
                if (ok && (v[0] == '+' || v[0] == '-')) k = 1;
                if (k >= v.size()) ok = false;
                for (size_t j = k; ok && j < v.size(); ++j) if (!isdigit((unsigned char)v[j])) ok = false;
                errno = 0;
                long long x = ok ? strtoll(v.c_str(), nullptr, 10) : 0;
                if (!ok || errno == ERANGE) {
                    err = "error parsing list " + name + " at line " + std::to_string(line) + ": error parsing int: invalid digit found in string";
                    return false;
                }
                L.ints.push_back((int64_t)x);
                break;
            }
            case LT_IP: {
                IpNet n;
                std::string e2;
                if (!parse_ip_network(v, &n, e2)) {
                    err = "error parsing list " + name + " at line " + std::to_string(line) + ": error parsing IP network: " + e2;
                    return false;
                }
                L.nets.push_back(n);
                break;
            }
        }
    }
    *out = std::move(L);
    return true;
}

}  // namespace pgw
