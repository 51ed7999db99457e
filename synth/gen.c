/* Synthetic request-stream generator (bench/test input only; not product, not oracle).
 * Follows SURVEY.md section 8(d): xoshiro256** seeded through SplitMix64 from
 * seed = 0x50494E474F4F0000 + config_id; every request is generated from
 * (seed, index) alone so shards regenerate independently. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s[4]; } rng_t;

static uint64_t splitmix(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, uint64_t seed) {
    uint64_t x = seed;
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix(&x);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t* r) {
    uint64_t* s = r->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
static double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

typedef struct { char* s; uint32_t len; } str_t;

typedef struct {
    char* data;
    size_t len, cap;
} buf_t;
static void buf_put(buf_t* b, const char* s, size_t n) {
    if (b->len + n + 1 > b->cap) {
        while (b->len + n + 1 > b->cap) b->cap = b->cap ? b->cap * 2 : 4096;
        b->data = (char*)realloc(b->data, b->cap);
    }
    memcpy(b->data + b->len, s, n);
    b->len += n;
}

typedef struct {
    uint64_t seed;
    /* vocabularies */
    str_t hosts[1024];
    str_t* paths; /* 65536 */
    double* host_cdf;
    double* path_cdf;
    str_t keys[256];
    /* knobs */
    double attack_rate, captcha_verified_rate, v6_rate, blocklist_rate, special_ip_rate;
    int get_only;
    uint32_t long_url_bytes; /* >0: adversarial config, URLs padded with trigger runs to this size */
    /* attack payloads */
    str_t* payloads;
    uint8_t* payload_field; /* 0 url (query), 1 path, 2 user agent */
    uint32_t n_payloads;
    /* addresses that must hit the blocklist */
    uint8_t* bl_ips; /* n x 17: 16 bytes + is_v6 */
    uint32_t n_bl;
    /* output buffers */
    buf_t col[5];
    uint32_t* off[5];
    uint8_t *ip, *is_v6, *flags;
    int32_t* port;
    uint32_t out_cap;
} gen_t;

static const char* SYL[] = {"ka", "lo", "mi", "ta", "re", "no", "vi", "su", "pe", "da", "ro", "li", "ne", "mo", "sa", "tu",
                            "be", "ci", "fo", "ga", "he", "ju", "ky", "wa", "xe", "zo", "an", "er", "in", "on", "st", "ar"};
static const char* TLD[] = {".com", ".net", ".org", ".io", ".dev", ".app", ".co", ".example"};
static const char* EXT[] = {"", "", "", ".html", ".php", ".js", ".css", ".png", ".json", ".svg"};
static const char* UA_T[] = {
    "Mozilla/5.0 (Windows NT 10.0; Win64; x64) AppleWebKit/537.36 (KHTML, like Gecko) Chrome/%u.0.%u.%u Safari/537.36",
    "Mozilla/5.0 (Macintosh; Intel Mac OS X 10_15_7) AppleWebKit/605.1.15 (KHTML, like Gecko) Version/%u.%u Safari/605.1.%u",
    "Mozilla/5.0 (X11; Linux x86_64; rv:%u.0) Gecko/20100101 Firefox/%u.%u",
    "Mozilla/5.0 (iPhone; CPU iPhone OS %u_%u like Mac OS X) AppleWebKit/605.1.15 (KHTML, like Gecko) Version/%u.0 Mobile/15E148 Safari/604.1",
    "Mozilla/5.0 (Linux; Android %u; Pixel %u) AppleWebKit/537.36 (KHTML, like Gecko) Chrome/%u.0.0.0 Mobile Safari/537.36",
    "Mozilla/5.0 (Windows NT 10.0; Win64; x64) AppleWebKit/537.36 (KHTML, like Gecko) Chrome/%u.0.0.0 Safari/537.36 Edg/%u.0.%u.0",
    "Mozilla/5.0 (compatible; Googlebot/2.%u; +http://www.google.com/bot.html) rev/%u.%u",
    "Mozilla/5.0 (compatible; bingbot/2.%u; +http://www.bing.com/bingbot.htm) b/%u.%u",
    "curl/%u.%u.%u",
    "python-requests/2.%u.%u k/%u",
    "Go-http-client/%u.%u (x%u)",
    "okhttp/%u.%u.%u",
    "Wget/1.%u.%u (linux-gnu) r%u",
    "PostmanRuntime/7.%u.%u p%u",
    "Apache-HttpClient/4.5.%u (Java/%u.0.%u)",
    "Mozilla/5.0 (Windows NT 6.1; WOW64; Trident/7.0; rv:%u.0) like Gecko t/%u.%u",
};
#define N_UA_T (sizeof UA_T / sizeof UA_T[0])

static str_t mkstr(const char* s, size_t n) {
    str_t r;
    r.s = (char*)malloc(n + 1);
    memcpy(r.s, s, n);
    r.s[n] = 0;
    r.len = (uint32_t)n;
    return r;
}

static size_t word(rng_t* r, char* out, int min_syl, int max_syl) {
    int k = min_syl + (int)rng_below(r, (uint32_t)(max_syl - min_syl + 1));
    size_t n = 0;
    for (int i = 0; i < k; ++i) {
        const char* s = SYL[rng_below(r, 32)];
        out[n++] = s[0];
        out[n++] = s[1];
    }
    return n;
}

static double* zipf_cdf(uint32_t n, double s) {
    double* c = (double*)malloc(sizeof(double) * n);
    double acc = 0;
    for (uint32_t i = 0; i < n; ++i) { acc += 1.0 / pow((double)(i + 1), s); c[i] = acc; }
    for (uint32_t i = 0; i < n; ++i) c[i] /= acc;
    return c;
}
static uint32_t cdf_pick(const double* c, uint32_t n, double u) {
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi) {
        uint32_t m = (lo + hi) / 2;
        if (c[m] < u) lo = m + 1;
        else hi = m;
    }
    return lo;
}

gen_t* synth_create(uint64_t seed) {
    gen_t* g = (gen_t*)calloc(1, sizeof *g);
    g->seed = seed;
    g->attack_rate = 0.03;
    g->captcha_verified_rate = 0.05;
    g->v6_rate = 0.10;
    g->blocklist_rate = 0.01;
    g->special_ip_rate = 0.001;
    rng_t r;
    rng_seed(&r, seed ^ 0xA5A5A5A5DEADBEEFull);
    char tmp[512];
    for (int i = 0; i < 1024; ++i) {
        size_t n = 0;
        if (rng_below(&r, 2) == 0) { n += word(&r, tmp + n, 1, 3); tmp[n++] = '.'; }
        n += word(&r, tmp + n, 3, 9);
        const char* t = TLD[rng_below(&r, 8)];
        memcpy(tmp + n, t, strlen(t));
        n += strlen(t);
        g->hosts[i] = mkstr(tmp, n);
    }
    g->paths = (str_t*)malloc(sizeof(str_t) * 65536);
    for (int i = 0; i < 65536; ++i) {
        size_t n = 0;
        int segs = 2 + (int)rng_below(&r, 8);
        for (int s = 0; s < segs && n < 140; ++s) {
            tmp[n++] = '/';
            if (rng_below(&r, 5) == 0) n += (size_t)sprintf(tmp + n, "%u", rng_below(&r, 100000));
            else n += word(&r, tmp + n, 1, 8);
        }
        const char* e = EXT[rng_below(&r, 10)];
        memcpy(tmp + n, e, strlen(e));
        n += strlen(e);
        while (n < 8) tmp[n++] = (char)('a' + rng_below(&r, 26));
        g->paths[i] = mkstr(tmp, n);
    }
    for (int i = 0; i < 256; ++i) {
        size_t n = word(&r, tmp, 1, 4);
        g->keys[i] = mkstr(tmp, n);
    }
    g->host_cdf = zipf_cdf(1024, 1.0);
    g->path_cdf = zipf_cdf(65536, 1.1);
    return g;
}

void synth_set_rates(gen_t* g, double attack, double captcha_verified, double v6, double blocklist, double special_ip, int get_only,
                     uint32_t long_url_bytes) {
    g->attack_rate = attack;
    g->captcha_verified_rate = captcha_verified;
    g->v6_rate = v6;
    g->blocklist_rate = blocklist;
    g->special_ip_rate = special_ip;
    g->get_only = get_only;
    g->long_url_bytes = long_url_bytes;
}

void synth_set_payloads(gen_t* g, const char** payloads, const uint32_t* lens, const uint8_t* fields, uint32_t n) {
    g->payloads = (str_t*)malloc(sizeof(str_t) * (n ? n : 1));
    g->payload_field = (uint8_t*)malloc(n ? n : 1);
    for (uint32_t i = 0; i < n; ++i) {
        g->payloads[i] = mkstr(payloads[i], lens[i]);
        g->payload_field[i] = fields[i];
    }
    g->n_payloads = n;
}

void synth_set_blocklist_ips(gen_t* g, const uint8_t* ips17, uint32_t n) {
    g->bl_ips = (uint8_t*)malloc((size_t)n * 17 + 1);
    memcpy(g->bl_ips, ips17, (size_t)n * 17);
    g->n_bl = n;
}

static const char URLSAFE[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-_.~";

static void gen_one(gen_t* g, uint64_t index, buf_t* host, buf_t* url, buf_t* path, buf_t* method, buf_t* ua, uint8_t* ip16, uint8_t* v6,
                    int32_t* port, uint8_t* flags) {
    rng_t r;
    rng_seed(&r, g->seed + 0x9E3779B97F4A7C15ull * (index + 1));
    char tmp[1024];
    /* method */
    const char* m = "GET";
    if (!g->get_only) {
        uint32_t x = rng_below(&r, 100);
        if (x >= 80 && x < 95) m = "POST";
        else if (x >= 95) { static const char* O[] = {"HEAD", "PUT", "DELETE", "OPTIONS"}; m = O[rng_below(&r, 4)]; }
    }
    buf_put(method, m, strlen(m));
    /* host */
    const str_t* h = &g->hosts[cdf_pick(g->host_cdf, 1024, rng_unit(&r))];
    buf_put(host, h->s, h->len);
    /* attack decision */
    int attack = g->n_payloads && rng_unit(&r) < g->attack_rate;
    const str_t* pay = NULL;
    int pay_field = 0;
    if (attack) {
        uint32_t pi = rng_below(&r, g->n_payloads);
        pay = &g->payloads[pi];
        pay_field = g->payload_field[pi];
    }
    /* path */
    const str_t* p = &g->paths[cdf_pick(g->path_cdf, 65536, rng_unit(&r))];
    size_t url_start = url->len;
    buf_put(path, p->s, p->len);
    buf_put(url, p->s, p->len);
    if (attack && pay_field == 1) {
        buf_put(path, "/", 1);
        buf_put(path, pay->s, pay->len);
        buf_put(url, "/", 1);
        buf_put(url, pay->s, pay->len);
    }
    /* query: 0-8 params, value lengths log-normal; tuned for a ~300-byte mean URL */
    uint32_t np = rng_below(&r, 9);
    int first = 1;
    for (uint32_t k = 0; k < np; ++k) {
        const str_t* key = &g->keys[rng_below(&r, 256)];
        buf_put(url, first ? "?" : "&", 1);
        first = 0;
        buf_put(url, key->s, key->len);
        buf_put(url, "=", 1);
        /* log-normal(mu=3.72, sigma=0.7): mean ~ 52 */
        double u1 = rng_unit(&r), u2 = rng_unit(&r);
        double z = sqrt(-2.0 * log(u1 + 1e-300)) * cos(6.283185307179586 * u2);
        int vl = (int)exp(3.72 + 0.7 * z);
        if (vl < 1) vl = 1;
        if (vl > 600) vl = 600;
        for (int i = 0; i < vl; ++i) {
            uint32_t x = rng_below(&r, 80);
            if (x < 66) tmp[0] = URLSAFE[x], buf_put(url, tmp, 1);
            else if (x < 72) buf_put(url, "%20", 3), i += 2;
            else if (x < 76) buf_put(url, "+", 1);
            else { sprintf(tmp, "%%%02X", 0x21 + rng_below(&r, 0x5E)); buf_put(url, tmp, 3); i += 2; }
        }
    }
    if (attack && pay_field == 0) {
        buf_put(url, first ? "?" : "&", 1);
        first = 0;
        buf_put(url, "q=", 2);
        buf_put(url, pay->s, pay->len);
    }
    if (g->long_url_bytes) {
        /* adversarial: pad with long runs of trigger characters */
        static const char* RUNS[] = {"a", "x", "aa", "ab ", "select", "xy", "a1 ", "aaaaaaaaab"};
        const char* run = RUNS[rng_below(&r, 8)];
        size_t rl = strlen(run);
        buf_put(url, first ? "?" : "&", 1);
        buf_put(url, "z=", 2);
        while (url->len - url_start + rl <= g->long_url_bytes) buf_put(url, run, rl);
        while (url->len - url_start < g->long_url_bytes) buf_put(url, "a", 1);
    }
    /* user agent */
    if (attack && pay_field == 2) {
        buf_put(ua, pay->s, pay->len);
    } else {
        /* 64 "templates": 16 families x 4 version bands */
        uint32_t fam = rng_below(&r, 100);
        uint32_t t = fam < 70 ? rng_below(&r, 6) : 6 + rng_below(&r, (uint32_t)N_UA_T - 6);
        uint32_t band = rng_below(&r, 4);
        int n = snprintf(tmp, sizeof tmp, UA_T[t], 90 + band * 10 + rng_below(&r, 10), rng_below(&r, 7000), rng_below(&r, 200));
        if (n > 0 && n < 230) n += snprintf(tmp + n, sizeof tmp - (size_t)n, " Build/%06u.%05u", rng_below(&r, 1000000), rng_below(&r, 100000));
        if (n > 250) n = 250;
        buf_put(ua, tmp, (size_t)n);
    }
    /* client */
    memset(ip16, 0, 16);
    *v6 = 0;
    double u = rng_unit(&r);
    if (g->n_bl && u < g->blocklist_rate) {
        const uint8_t* e = g->bl_ips + (size_t)rng_below(&r, g->n_bl) * 17;
        memcpy(ip16, e, 16);
        *v6 = e[16];
    } else if (u < g->blocklist_rate + g->special_ip_rate) {
        uint32_t k = rng_below(&r, 4);
        if (k == 0) { ip16[0] = 127; ip16[3] = 1; }
        else if (k == 1) { ip16[0] = 224; ip16[3] = 251; }
        else if (k == 2) { ip16[15] = 1; *v6 = 1; }
        else { ip16[0] = 0xFF; ip16[1] = 0x02; ip16[15] = 1; *v6 = 1; }
    } else if (rng_unit(&r) < g->v6_rate) {
        *v6 = 1;
        ip16[0] = 0x20 + (uint8_t)rng_below(&r, 2) * 6; /* 2000::/8 or 2600::/8 */
        ip16[1] = (uint8_t)rng_below(&r, 256);
        for (int k = 2; k < 8; ++k) ip16[k] = (uint8_t)rng_below(&r, 16);
        for (int k = 8; k < 16; ++k) ip16[k] = (uint8_t)rng_below(&r, 256);
    } else {
        ip16[0] = (uint8_t)(1 + rng_below(&r, 223));
        if (ip16[0] == 127) ip16[0] = 128;
        ip16[1] = (uint8_t)rng_below(&r, 256);
        ip16[2] = (uint8_t)rng_below(&r, 256);
        ip16[3] = (uint8_t)rng_below(&r, 256);
    }
    *port = 1024 + (int32_t)rng_below(&r, 65536 - 1024);
    *flags = rng_unit(&r) < g->captcha_verified_rate ? 1 : 0;
}

/* Generate requests [first, first+n) into internal buffers; pointers stay valid until the next call. */
int synth_generate(gen_t* g, uint64_t first, uint32_t n) {
    if (n > g->out_cap) {
        for (int f = 0; f < 5; ++f) g->off[f] = (uint32_t*)realloc(g->off[f], sizeof(uint32_t) * ((size_t)n + 1));
        g->ip = (uint8_t*)realloc(g->ip, (size_t)n * 16);
        g->is_v6 = (uint8_t*)realloc(g->is_v6, n);
        g->flags = (uint8_t*)realloc(g->flags, n);
        g->port = (int32_t*)realloc(g->port, sizeof(int32_t) * (size_t)n);
        g->out_cap = n;
    }
    for (int f = 0; f < 5; ++f) g->col[f].len = 0;
    for (uint32_t i = 0; i < n; ++i) {
        for (int f = 0; f < 5; ++f) g->off[f][i] = (uint32_t)g->col[f].len;
        gen_one(g, first + i, &g->col[0], &g->col[1], &g->col[2], &g->col[3], &g->col[4], g->ip + (size_t)i * 16, &g->is_v6[i], &g->port[i], &g->flags[i]);
        for (int f = 0; f < 5; ++f)
            if (g->col[f].len > 0xFFFFFFF0ull) return 1; /* column exceeds 4 GiB: caller must split the batch */
    }
    for (int f = 0; f < 5; ++f) {
        g->off[f][n] = (uint32_t)g->col[f].len;
        static const char zeros[64] = {0};
        buf_put(&g->col[f], zeros, 64); /* readable padding */
        g->col[f].len -= 64;
    }
    return 0;
}

const uint8_t* synth_col_bytes(gen_t* g, int f) { return (const uint8_t*)g->col[f].data; }
uint64_t synth_col_len(gen_t* g, int f) { return g->col[f].len; }
const uint32_t* synth_col_offsets(gen_t* g, int f) { return g->off[f]; }
const uint8_t* synth_ip(gen_t* g) { return g->ip; }
const uint8_t* synth_is_v6(gen_t* g) { return g->is_v6; }
const uint8_t* synth_flags(gen_t* g) { return g->flags; }
const int32_t* synth_port(gen_t* g) { return g->port; }

void synth_destroy(gen_t* g) {
    if (!g) return;
    for (int i = 0; i < 1024; ++i) free(g->hosts[i].s);
    for (int i = 0; i < 65536; ++i) free(g->paths[i].s);
    for (int i = 0; i < 256; ++i) free(g->keys[i].s);
    free(g->paths); free(g->host_cdf); free(g->path_cdf);
    for (uint32_t i = 0; i < g->n_payloads; ++i) free(g->payloads[i].s);
    free(g->payloads); free(g->payload_field); free(g->bl_ips);
    for (int f = 0; f < 5; ++f) { free(g->col[f].data); free(g->off[f]); }
    free(g->ip); free(g->is_v6); free(g->flags); free(g->port);
    free(g);
}

/* ---- MaxMind-DB search tree builder (format 2.0, record size 28, ip_version 6) -------------------------------
 * Input: n networks, each 17 bytes = 16 address bytes (IPv4 networks already mapped under ::/96) + prefix length
 * (1..128, already +96 for IPv4), and for each the byte offset of its record inside the data section.  Networks must
 * be prefix-free.  Output: the tree section only (node_count * 7 bytes); the caller appends the 16-byte separator,
 * the data section, the metadata marker and the metadata.  Returns node_count, or 0 on allocation failure. */
typedef struct { uint32_t kid[2]; } mnode_t; /* 0 = empty, 1..: node index + 1 when < 0x80000000, else 0x80000000 | leaf */

uint32_t synth_mmdb_tree(const uint8_t* nets17, const uint32_t* data_off, uint32_t n, uint8_t** out, size_t* out_len) {
    size_t cap = (size_t)n * 24 + 1024, cnt = 1;
    mnode_t* nodes = (mnode_t*)calloc(cap, sizeof(mnode_t));
    if (!nodes) return 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t* a = nets17 + (size_t)i * 17;
        int plen = a[16];
        size_t cur = 0;
        for (int d = 0; d < plen; ++d) {
            int bit = (a[d >> 3] >> (7 - (d & 7))) & 1;
            if (d == plen - 1) {
                nodes[cur].kid[bit] = 0x80000000u | i;
            } else {
                uint32_t k = nodes[cur].kid[bit];
                if (k == 0 || (k & 0x80000000u)) {
                    if (cnt == cap) {
                        cap *= 2;
                        mnode_t* nn = (mnode_t*)realloc(nodes, cap * sizeof(mnode_t));
                        if (!nn) { free(nodes); return 0; }
                        memset(nn + cnt, 0, (cap - cnt) * sizeof(mnode_t));
                        nodes = nn;
                    }
                    nodes[cur].kid[bit] = (uint32_t)cnt + 1;
                    k = (uint32_t)cnt + 1;
                    ++cnt;
                }
                cur = k - 1;
            }
        }
    }
    uint32_t node_count = (uint32_t)cnt;
    uint8_t* t = (uint8_t*)malloc((size_t)node_count * 7);
    if (!t) { free(nodes); return 0; }
    for (uint32_t i = 0; i < node_count; ++i) {
        uint32_t rec[2];
        for (int b = 0; b < 2; ++b) {
            uint32_t k = nodes[i].kid[b];
            if (k == 0) rec[b] = node_count;                                   /* no data */
            else if (k & 0x80000000u) rec[b] = node_count + 16 + data_off[k & 0x7FFFFFFFu];
            else rec[b] = k - 1;
        }
        uint8_t* o = t + (size_t)i * 7;
        o[0] = (uint8_t)(rec[0] >> 16); o[1] = (uint8_t)(rec[0] >> 8); o[2] = (uint8_t)rec[0];
        o[3] = (uint8_t)((((rec[0] >> 24) & 0xF) << 4) | ((rec[1] >> 24) & 0xF));
        o[4] = (uint8_t)(rec[1] >> 16); o[5] = (uint8_t)(rec[1] >> 8); o[6] = (uint8_t)rec[1];
    }
    free(nodes);
    *out = t;
    *out_len = (size_t)node_count * 7;
    return node_count;
}
void synth_free(void* p) { free(p); }
