/* ORACLE (test infrastructure): parser + tree-walking evaluator for the rule
 * language.  See bel.h.  Deliberately naive: values are materialised, lists are
 * scanned linearly, `&&`/`||` short-circuit left to right, any error aborts the
 * evaluation -- the behaviour `Rule::match_request` observes (pingoo/rules.rs:36-52). */
#include "bel.h"

#include <ctype.h>
#include <errno.h>
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rx.h"

/* ---- AST ---------------------------------------------------------------------------- */
enum {
    K_NULL, K_BOOL, K_INT, K_UINT, K_FLOAT, K_STR, K_BYTES, K_IDENT, K_MEMBER, K_INDEX, K_CALL, K_METHOD,
    K_NOT, K_NEG, K_OR, K_AND, K_EQ, K_NE, K_LT, K_LE, K_GT, K_GE, K_IN, K_ADD, K_SUB, K_MUL, K_DIV, K_MOD,
    K_COND, K_LIST, K_MAP
};

struct bel_expr {
    int kind;
    char* s; /* identifier / member / function name, or string payload */
    size_t slen;
    int64_t i;
    double f;
    struct bel_expr** kid;
    int nkid;
    rx_prog* rx; /* precompiled pattern for matches("literal") */
    int rx_status;
};

/* ---- lexer -------------------------------------------------------------------------- */
enum {
    T_END, T_IDENT, T_INT, T_UINT, T_FLOAT, T_STR, T_BYTES, T_LP, T_RP, T_LB, T_RB, T_LC, T_RC, T_DOT, T_COMMA, T_COLON,
    T_Q, T_NOT, T_MINUS, T_PLUS, T_STAR, T_SLASH, T_PCT, T_OR, T_AND, T_EQ, T_NE, T_LT, T_LE, T_GT, T_GE
};

typedef struct {
    int t;
    char* s;
    size_t slen;
    int64_t i;
    int is_min; /* literal 9223372036854775808, valid only after unary minus */
    double f;
    size_t pos;
} token;

typedef struct {
    const char* src;
    size_t n, p;
    token cur;
    jmp_buf jb;
    char* err;
    size_t cap;
    int depth;
} P;

__attribute__((noreturn)) static void perr(P* p, const char* msg, size_t at) {
    snprintf(p->err, p->cap, "%s at offset %zu", msg, at);
    longjmp(p->jb, 1);
}

static void buf_push(char** b, size_t* n, size_t* cap, char c) {
    if (*n + 1 >= *cap) {
        *cap = *cap ? *cap * 2 : 32;
        *b = (char*)realloc(*b, *cap);
    }
    (*b)[(*n)++] = c;
    (*b)[*n] = 0;
}

static void push_utf8(char** b, size_t* n, size_t* cap, uint32_t cp) {
    if (cp < 0x80) buf_push(b, n, cap, (char)cp);
    else if (cp < 0x800) { buf_push(b, n, cap, (char)(0xC0 | (cp >> 6))); buf_push(b, n, cap, (char)(0x80 | (cp & 63))); }
    else if (cp < 0x10000) { buf_push(b, n, cap, (char)(0xE0 | (cp >> 12))); buf_push(b, n, cap, (char)(0x80 | ((cp >> 6) & 63))); buf_push(b, n, cap, (char)(0x80 | (cp & 63))); }
    else { buf_push(b, n, cap, (char)(0xF0 | (cp >> 18))); buf_push(b, n, cap, (char)(0x80 | ((cp >> 12) & 63))); buf_push(b, n, cap, (char)(0x80 | ((cp >> 6) & 63))); buf_push(b, n, cap, (char)(0x80 | (cp & 63))); }
}

static void lex_string(P* p, int raw, int bytes, token* k) {
    size_t st = p->p;
    char q = p->src[p->p];
    int triple = p->p + 2 < p->n && p->src[p->p + 1] == q && p->src[p->p + 2] == q;
    p->p += triple ? 3 : 1;
    char* out = NULL;
    size_t on = 0, ocap = 0;
    buf_push(&out, &on, &ocap, 0);
    on = 0;
    for (;;) {
        if (p->p >= p->n) { free(out); perr(p, "unterminated string literal", st); }
        char c = p->src[p->p];
        if (c == q) {
            if (!triple) { p->p++; break; }
            if (p->p + 2 < p->n && p->src[p->p + 1] == q && p->src[p->p + 2] == q) { p->p += 3; break; }
            buf_push(&out, &on, &ocap, c);
            p->p++;
            continue;
        }
        if (!triple && (c == '\n' || c == '\r')) { free(out); perr(p, "newline in string literal", p->p); }
        if (c != '\\' || raw) { buf_push(&out, &on, &ocap, c); p->p++; continue; }
        p->p++;
        if (p->p >= p->n) { free(out); perr(p, "unterminated escape", p->p); }
        char e = p->src[p->p++];
        switch (e) {
            case 'a': buf_push(&out, &on, &ocap, '\a'); break;
            case 'b': buf_push(&out, &on, &ocap, '\b'); break;
            case 'f': buf_push(&out, &on, &ocap, '\f'); break;
            case 'n': buf_push(&out, &on, &ocap, '\n'); break;
            case 'r': buf_push(&out, &on, &ocap, '\r'); break;
            case 't': buf_push(&out, &on, &ocap, '\t'); break;
            case 'v': buf_push(&out, &on, &ocap, '\v'); break;
            case '\\': case '?': case '"': case '\'': case '`': buf_push(&out, &on, &ocap, e); break;
            case 'x': case 'X': case 'u': case 'U': {
                int nd = (e == 'x' || e == 'X') ? 2 : (e == 'u' ? 4 : 8);
                if (bytes && nd != 2) { free(out); perr(p, "\\u escape in bytes literal", p->p); }
                uint32_t v = 0;
                for (int d = 0; d < nd; ++d) {
                    if (p->p >= p->n || !isxdigit((unsigned char)p->src[p->p])) { free(out); perr(p, "invalid hex escape", p->p); }
                    char h = p->src[p->p++];
                    v = v * 16 + (uint32_t)(isdigit((unsigned char)h) ? h - '0' : tolower(h) - 'a' + 10);
                }
                if (bytes) buf_push(&out, &on, &ocap, (char)v);
                else {
                    if (v > 0x10FFFF || (v >= 0xD800 && v <= 0xDFFF)) { free(out); perr(p, "invalid code point", p->p); }
                    push_utf8(&out, &on, &ocap, v);
                }
                break;
            }
            default:
                if (e >= '0' && e <= '3') {
                    uint32_t v = (uint32_t)(e - '0');
                    for (int d = 0; d < 2; ++d) {
                        if (p->p >= p->n || p->src[p->p] < '0' || p->src[p->p] > '7') { free(out); perr(p, "invalid octal escape", p->p); }
                        v = v * 8 + (uint32_t)(p->src[p->p++] - '0');
                    }
                    if (bytes) buf_push(&out, &on, &ocap, (char)v);
                    else push_utf8(&out, &on, &ocap, v);
                    break;
                }
                free(out);
                perr(p, "invalid escape sequence", p->p - 2);
        }
    }
    k->t = bytes ? T_BYTES : T_STR;
    k->s = out;
    k->slen = on;
}

static void next_token(P* p) {
    token* k = &p->cur;
    memset(k, 0, sizeof *k);
    for (;;) {
        while (p->p < p->n && isspace((unsigned char)p->src[p->p])) p->p++;
        if (p->p + 1 < p->n && p->src[p->p] == '/' && p->src[p->p + 1] == '/') {
            while (p->p < p->n && p->src[p->p] != '\n') p->p++;
            continue;
        }
        break;
    }
    k->pos = p->p;
    if (p->p >= p->n) { k->t = T_END; return; }
    unsigned char c = (unsigned char)p->src[p->p];
    if (isalpha(c) || c == '_') {
        size_t q = p->p;
        int raw = 0, bytes = 0;
        while (q < p->n && q - p->p < 2 && strchr("rRbB", p->src[q])) {
            if (p->src[q] == 'r' || p->src[q] == 'R') { if (raw) break; raw = 1; }
            else { if (bytes) break; bytes = 1; }
            q++;
        }
        if (q > p->p && q < p->n && (p->src[q] == '"' || p->src[q] == '\'')) {
            p->p = q;
            lex_string(p, raw, bytes, k);
            return;
        }
        size_t st = p->p;
        while (p->p < p->n && (isalnum((unsigned char)p->src[p->p]) || p->src[p->p] == '_')) p->p++;
        k->t = T_IDENT;
        k->slen = p->p - st;
        k->s = (char*)malloc(k->slen + 1);
        memcpy(k->s, p->src + st, k->slen);
        k->s[k->slen] = 0;
        return;
    }
    if (isdigit(c) || (c == '.' && p->p + 1 < p->n && isdigit((unsigned char)p->src[p->p + 1]))) {
        size_t st = p->p;
        unsigned long long v = 0;
        int is_float = 0;
        if (c == '0' && p->p + 1 < p->n && (p->src[p->p + 1] == 'x' || p->src[p->p + 1] == 'X')) {
            p->p += 2;
            size_t hs = p->p;
            while (p->p < p->n && isxdigit((unsigned char)p->src[p->p])) p->p++;
            if (p->p == hs) perr(p, "invalid hex literal", st);
            char tmp[64];
            size_t l = p->p - hs;
            if (l >= sizeof tmp) perr(p, "integer literal out of range", st);
            memcpy(tmp, p->src + hs, l);
            tmp[l] = 0;
            errno = 0;
            v = strtoull(tmp, NULL, 16);
            if (errno) perr(p, "integer literal out of range", st);
        } else {
            while (p->p < p->n && isdigit((unsigned char)p->src[p->p])) p->p++;
            if (p->p < p->n && p->src[p->p] == '.' && p->p + 1 < p->n && isdigit((unsigned char)p->src[p->p + 1])) {
                is_float = 1;
                p->p++;
                while (p->p < p->n && isdigit((unsigned char)p->src[p->p])) p->p++;
            }
            if (p->p < p->n && (p->src[p->p] == 'e' || p->src[p->p] == 'E')) {
                size_t save = p->p;
                p->p++;
                if (p->p < p->n && (p->src[p->p] == '+' || p->src[p->p] == '-')) p->p++;
                if (p->p < p->n && isdigit((unsigned char)p->src[p->p])) {
                    is_float = 1;
                    while (p->p < p->n && isdigit((unsigned char)p->src[p->p])) p->p++;
                } else p->p = save;
            }
            char tmp[128];
            size_t l = p->p - st;
            if (l >= sizeof tmp) perr(p, "numeric literal too long", st);
            memcpy(tmp, p->src + st, l);
            tmp[l] = 0;
            if (is_float) {
                k->t = T_FLOAT;
                k->f = strtod(tmp, NULL);
                return;
            }
            errno = 0;
            v = strtoull(tmp, NULL, 10);
            if (errno) perr(p, "integer literal out of range", st);
        }
        if (p->p < p->n && (p->src[p->p] == 'u' || p->src[p->p] == 'U')) {
            p->p++;
            k->t = T_UINT;
            k->i = (int64_t)v;
            return;
        }
        if (v > 9223372036854775808ull) perr(p, "integer literal out of range", st);
        k->t = T_INT;
        k->i = (int64_t)v;
        k->is_min = v == 9223372036854775808ull;
        return;
    }
    if (c == '"' || c == '\'') { lex_string(p, 0, 0, k); return; }
    p->p++;
    int two = p->p < p->n ? p->src[p->p] : 0;
    switch (c) {
        case '(': k->t = T_LP; return;
        case ')': k->t = T_RP; return;
        case '[': k->t = T_LB; return;
        case ']': k->t = T_RB; return;
        case '{': k->t = T_LC; return;
        case '}': k->t = T_RC; return;
        case '.': k->t = T_DOT; return;
        case ',': k->t = T_COMMA; return;
        case ':': k->t = T_COLON; return;
        case '?': k->t = T_Q; return;
        case '+': k->t = T_PLUS; return;
        case '-': k->t = T_MINUS; return;
        case '*': k->t = T_STAR; return;
        case '/': k->t = T_SLASH; return;
        case '%': k->t = T_PCT; return;
        case '!': if (two == '=') { p->p++; k->t = T_NE; } else k->t = T_NOT; return;
        case '=': if (two == '=') { p->p++; k->t = T_EQ; return; } break;
        case '<': if (two == '=') { p->p++; k->t = T_LE; } else k->t = T_LT; return;
        case '>': if (two == '=') { p->p++; k->t = T_GE; } else k->t = T_GT; return;
        case '|': if (two == '|') { p->p++; k->t = T_OR; return; } break;
        case '&': if (two == '&') { p->p++; k->t = T_AND; return; } break;
        default: break;
    }
    perr(p, "unexpected character", k->pos);
}

/* Parsed nodes are leaked on a syntax error inside setjmp; acceptable for test infrastructure. */
static bel_expr* mk(int kind) {
    bel_expr* e = (bel_expr*)calloc(1, sizeof *e);
    e->kind = kind;
    return e;
}
static void add_kid(bel_expr* e, bel_expr* k) {
    e->kid = (bel_expr**)realloc(e->kid, sizeof(bel_expr*) * (size_t)(e->nkid + 1));
    e->kid[e->nkid++] = k;
}
static bel_expr* bin(int kind, bel_expr* a, bel_expr* b) {
    bel_expr* e = mk(kind);
    add_kid(e, a);
    add_kid(e, b);
    return e;
}
static void advance(P* p) {
    /* ownership of cur.s moves to the AST when used; otherwise free */
    next_token(p);
}
static int accept(P* p, int t) {
    if (p->cur.t == t) { advance(p); return 1; }
    return 0;
}
static void expect(P* p, int t, const char* what) {
    if (!accept(p, t)) {
        char m[64];
        snprintf(m, sizeof m, "expected %s", what);
        perr(p, m, p->cur.pos);
    }
}

static bel_expr* parse_expr(P* p);
static bel_expr* parse_unary(P* p);

static void parse_args(P* p, bel_expr* call) {
    if (accept(p, T_RP)) return;
    for (;;) {
        add_kid(call, parse_expr(p));
        if (accept(p, T_COMMA)) {
            if (accept(p, T_RP)) return;
            continue;
        }
        expect(p, T_RP, "')'");
        return;
    }
}

static bel_expr* parse_primary(P* p) {
    token k = p->cur;
    switch (k.t) {
        case T_INT: {
            if (k.is_min) perr(p, "integer literal out of range", k.pos);
            advance(p);
            bel_expr* e = mk(K_INT);
            e->i = k.i;
            return e;
        }
        case T_UINT: { advance(p); bel_expr* e = mk(K_UINT); e->i = k.i; return e; }
        case T_FLOAT: { advance(p); bel_expr* e = mk(K_FLOAT); e->f = k.f; return e; }
        case T_STR: case T_BYTES: {
            advance(p);
            bel_expr* e = mk(k.t == T_STR ? K_STR : K_BYTES);
            e->s = k.s;
            e->slen = k.slen;
            return e;
        }
        case T_IDENT: {
            advance(p);
            if (!strcmp(k.s, "true") || !strcmp(k.s, "false")) {
                bel_expr* e = mk(K_BOOL);
                e->i = k.s[0] == 't';
                free(k.s);
                return e;
            }
            if (!strcmp(k.s, "null")) { free(k.s); return mk(K_NULL); }
            if (!strcmp(k.s, "in")) perr(p, "unexpected 'in'", k.pos);
            if (p->cur.t == T_LP) {
                advance(p);
                bel_expr* c = mk(K_CALL);
                c->s = k.s;
                parse_args(p, c);
                return c;
            }
            bel_expr* e = mk(K_IDENT);
            e->s = k.s;
            return e;
        }
        case T_LP: {
            advance(p);
            bel_expr* e = parse_expr(p);
            expect(p, T_RP, "')'");
            return e;
        }
        case T_LB: {
            advance(p);
            bel_expr* l = mk(K_LIST);
            if (accept(p, T_RB)) return l;
            for (;;) {
                add_kid(l, parse_expr(p));
                if (accept(p, T_COMMA)) {
                    if (accept(p, T_RB)) return l;
                    continue;
                }
                expect(p, T_RB, "']'");
                return l;
            }
        }
        case T_LC: {
            advance(p);
            bel_expr* m = mk(K_MAP);
            if (accept(p, T_RC)) return m;
            for (;;) {
                add_kid(m, parse_expr(p));
                expect(p, T_COLON, "':'");
                add_kid(m, parse_expr(p));
                if (accept(p, T_COMMA)) {
                    if (accept(p, T_RC)) return m;
                    continue;
                }
                expect(p, T_RC, "'}'");
                return m;
            }
        }
        case T_END: perr(p, "unexpected end of expression", k.pos);
        default: perr(p, "unexpected token", k.pos);
    }
    return NULL;
}

static bel_expr* parse_postfix(P* p, bel_expr* e) {
    for (;;) {
        if (p->cur.t == T_DOT) {
            advance(p);
            if (p->cur.t != T_IDENT) perr(p, "expected identifier after '.'", p->cur.pos);
            char* name = p->cur.s;
            advance(p);
            if (p->cur.t == T_LP) {
                advance(p);
                bel_expr* m = mk(K_METHOD);
                m->s = name;
                add_kid(m, e);
                parse_args(p, m);
                e = m;
            } else {
                bel_expr* m = mk(K_MEMBER);
                m->s = name;
                add_kid(m, e);
                e = m;
            }
            continue;
        }
        if (p->cur.t == T_LB) {
            advance(p);
            bel_expr* ix = mk(K_INDEX);
            add_kid(ix, e);
            add_kid(ix, parse_expr(p));
            expect(p, T_RB, "']'");
            e = ix;
            continue;
        }
        return e;
    }
}

static bel_expr* parse_unary(P* p) {
    if (++p->depth > 200) perr(p, "expression nesting too deep", p->cur.pos);
    bel_expr* r;
    if (p->cur.t == T_NOT) {
        advance(p);
        r = mk(K_NOT);
        add_kid(r, parse_unary(p));
    } else if (p->cur.t == T_MINUS) {
        advance(p);
        if (p->cur.t == T_INT && p->cur.is_min) {
            advance(p);
            bel_expr* e = mk(K_INT);
            e->i = INT64_MIN;
            r = parse_postfix(p, e);
        } else {
            r = mk(K_NEG);
            add_kid(r, parse_unary(p));
        }
    } else {
        r = parse_postfix(p, parse_primary(p));
    }
    p->depth--;
    return r;
}

static bel_expr* parse_mul(P* p) {
    bel_expr* l = parse_unary(p);
    for (;;) {
        int k = p->cur.t == T_STAR ? K_MUL : p->cur.t == T_SLASH ? K_DIV : p->cur.t == T_PCT ? K_MOD : -1;
        if (k < 0) return l;
        advance(p);
        l = bin(k, l, parse_unary(p));
    }
}
static bel_expr* parse_add(P* p) {
    bel_expr* l = parse_mul(p);
    for (;;) {
        int k = p->cur.t == T_PLUS ? K_ADD : p->cur.t == T_MINUS ? K_SUB : -1;
        if (k < 0) return l;
        advance(p);
        l = bin(k, l, parse_mul(p));
    }
}
static bel_expr* parse_rel(P* p) {
    bel_expr* l = parse_add(p);
    for (;;) {
        int k = -1;
        switch (p->cur.t) {
            case T_EQ: k = K_EQ; break;
            case T_NE: k = K_NE; break;
            case T_LT: k = K_LT; break;
            case T_LE: k = K_LE; break;
            case T_GT: k = K_GT; break;
            case T_GE: k = K_GE; break;
            case T_IDENT: if (!strcmp(p->cur.s, "in")) k = K_IN; break;
            default: break;
        }
        if (k < 0) return l;
        if (k == K_IN) free(p->cur.s);
        advance(p);
        l = bin(k, l, parse_add(p));
    }
}
static bel_expr* parse_and(P* p) {
    bel_expr* l = parse_rel(p);
    while (p->cur.t == T_AND) {
        advance(p);
        l = bin(K_AND, l, parse_rel(p));
    }
    return l;
}
static bel_expr* parse_or(P* p) {
    bel_expr* l = parse_and(p);
    while (p->cur.t == T_OR) {
        advance(p);
        l = bin(K_OR, l, parse_and(p));
    }
    return l;
}
static bel_expr* parse_expr(P* p) {
    if (++p->depth > 200) perr(p, "expression nesting too deep", p->cur.pos);
    bel_expr* c = parse_or(p);
    if (p->cur.t == T_Q) {
        advance(p);
        bel_expr* a = parse_or(p);
        expect(p, T_COLON, "':'");
        bel_expr* b = parse_expr(p);
        bel_expr* t = mk(K_COND);
        add_kid(t, c);
        add_kid(t, a);
        add_kid(t, b);
        c = t;
    }
    p->depth--;
    return c;
}

static void precompile(bel_expr* e) {
    for (int k = 0; k < e->nkid; ++k) precompile(e->kid[k]);
    if (e->kind == K_METHOD && !strcmp(e->s, "matches") && e->nkid == 2 && e->kid[1]->kind == K_STR) {
        char msg[128];
        e->rx = rx_compile(e->kid[1]->s, e->kid[1]->slen, &e->rx_status, msg);
    }
}

bel_expr* bel_compile(const char* src, char* err, size_t cap) {
    P p;
    memset(&p, 0, sizeof p);
    p.src = src;
    p.n = strlen(src);
    char local[160];
    p.err = local;
    p.cap = sizeof local;
    if (setjmp(p.jb)) {
        if (err && cap) snprintf(err, cap, "Expression is not valid: %s", local);
        return NULL;
    }
    next_token(&p);
    bel_expr* e = parse_expr(&p);
    if (p.cur.t != T_END) perr(&p, "unexpected token", p.cur.pos);
    precompile(e);
    return e;
}

void bel_free(bel_expr* e) {
    if (!e) return;
    for (int k = 0; k < e->nkid; ++k) bel_free(e->kid[k]);
    free(e->kid);
    free(e->s);
    rx_free(e->rx);
    free(e);
}

int bel_uses_in(const bel_expr* e) {
    if (e->kind == K_IN) return 1;
    for (int k = 0; k < e->nkid; ++k)
        if (bel_uses_in(e->kid[k])) return 1;
    return 0;
}

/* ---- IpNetwork (ipnetwork 0.21 FromStr / contains) ------------------------------------- */
static int parse_v4(const char* s, size_t n, uint8_t out[4]) {
    size_t p = 0;
    for (int k = 0; k < 4; ++k) {
        if (p >= n || !isdigit((unsigned char)s[p])) return 0;
        size_t st = p;
        unsigned v = 0;
        while (p < n && isdigit((unsigned char)s[p])) {
            v = v * 10 + (unsigned)(s[p] - '0');
            if (v > 255 || p - st >= 3) return 0;
            p++;
        }
        if (p - st > 1 && s[st] == '0') return 0; /* Rust rejects leading zeros */
        out[k] = (uint8_t)v;
        if (k < 3) {
            if (p >= n || s[p] != '.') return 0;
            p++;
        }
    }
    return p == n;
}

static int parse_v6(const char* s, size_t n, uint8_t out[16]) {
    /* RFC 4291 text form: groups split on ':', at most one "::", optional dotted quad at the end */
    uint16_t g[8];
    int ng = 0, gap_at = -1;
    size_t p = 0;
    if (n == 0) return 0;
    if (n >= 2 && s[0] == ':' && s[1] == ':') { gap_at = 0; p = 2; }
    else if (s[0] == ':') return 0;
    while (p < n) {
        size_t q = p;
        while (q < n && s[q] != ':') q++;
        size_t l = q - p;
        if (memchr(s + p, '.', l)) {
            uint8_t v4[4];
            if (q != n || !parse_v4(s + p, l, v4) || ng > 6) return 0;
            g[ng++] = (uint16_t)(v4[0] << 8 | v4[1]);
            g[ng++] = (uint16_t)(v4[2] << 8 | v4[3]);
            p = q;
            break;
        }
        if (l == 0 || l > 4 || ng >= 8) return 0;
        unsigned v = 0;
        for (size_t i = 0; i < l; ++i) {
            if (!isxdigit((unsigned char)s[p + i])) return 0;
            char h = s[p + i];
            v = v * 16 + (unsigned)(isdigit((unsigned char)h) ? h - '0' : tolower(h) - 'a' + 10);
        }
        g[ng++] = (uint16_t)v;
        p = q;
        if (p < n) {
            if (p + 1 < n && s[p + 1] == ':') {
                if (gap_at >= 0) return 0;
                gap_at = ng;
                p += 2;
            } else {
                p++;
                if (p >= n) return 0; /* trailing single ':' */
            }
        }
    }
    uint16_t full[8] = {0};
    if (gap_at >= 0) {
        if (ng > 7) return 0;
        for (int k = 0; k < gap_at; ++k) full[k] = g[k];
        for (int k = gap_at; k < ng; ++k) full[8 - (ng - k)] = g[k];
    } else {
        if (ng != 8) return 0;
        memcpy(full, g, sizeof full);
    }
    for (int k = 0; k < 8; ++k) { out[2 * k] = (uint8_t)(full[k] >> 8); out[2 * k + 1] = (uint8_t)full[k]; }
    return 1;
}

int bel_parse_ipnet(const char* s, bel_ipnet* out) {
    memset(out, 0, sizeof *out);
    const char* slash = strchr(s, '/');
    size_t alen = slash ? (size_t)(slash - s) : strlen(s);
    if (parse_v4(s, alen, out->addr)) { out->v6 = 0; out->prefix = 32; }
    else if (parse_v6(s, alen, out->addr)) { out->v6 = 1; out->prefix = 128; }
    else return 0;
    if (slash) {
        const char* pf = slash + 1;
        size_t pl = strlen(pf);
        /* parse_prefix = `s.parse::<u8>()` then `<= max`: Rust's integer FromStr takes an optional '+', any number of leading
         * zeros, nothing else, and fails on overflow of the type (255) */
        size_t d0 = (pl > 0 && pf[0] == '+') ? 1 : 0;
        int digits = pl > d0;
        unsigned v = 0;
        for (size_t i = d0; i < pl && digits; ++i) {
            if (!isdigit((unsigned char)pf[i])) digits = 0;
            else { v = v * 10 + (unsigned)(pf[i] - '0'); if (v > 255) digits = 0; }
        }
        int numeric = pl > d0;
        for (size_t i = d0; i < pl; ++i) if (!isdigit((unsigned char)pf[i])) numeric = 0;
        if (numeric) {
            if (!digits || v > (unsigned)(out->v6 ? 128 : 32)) return 0;
            out->prefix = (int)v;
        } else if (!out->v6) {
            uint8_t m[4];
            if (!parse_v4(pf, pl, m)) return 0;
            uint32_t mask = (uint32_t)m[0] << 24 | (uint32_t)m[1] << 16 | (uint32_t)m[2] << 8 | m[3];
            int len = 0;
            while (len < 32 && (mask & (0x80000000u >> len))) len++;
            if (len < 32 && (uint32_t)(mask << len) != 0) return 0;
            out->prefix = len;
        } else return 0;
    }
    return 1;
}

int bel_ipnet_contains(const bel_ipnet* net, const uint8_t* ip, int is_v6) {
    /* Ipv4Network::contains / Ipv6Network::contains: masked compare; families never mix (A7) */
    if ((net->v6 != 0) != (is_v6 != 0)) return 0;
    int bits = net->prefix;
    int nbytes = is_v6 ? 16 : 4;
    for (int k = 0; k < nbytes && bits > 0; ++k, bits -= 8) {
        uint8_t mask = bits >= 8 ? 0xFF : (uint8_t)(0xFF << (8 - bits));
        if ((ip[k] & mask) != (net->addr[k] & mask)) return 0;
    }
    return 1;
}

/* ---- values & evaluation ------------------------------------------------------------------ */
typedef enum { V_ERR, V_NULL, V_BOOL, V_INT, V_UINT, V_FLOAT, V_STR, V_BYTES, V_IP, V_IPNET, V_LIST, V_MAP } vtype;
enum { M_HTTP = 1, M_CLIENT, M_LISTS, M_LITERAL };

typedef struct val {
    vtype t;
    int64_t i;
    double f;
    const uint8_t* s;
    size_t n;
    uint8_t ip[16];
    int v6;
    const bel_ipnet* net;
    struct val* items; /* literal list items / literal map key,value pairs */
    size_t n_items;
    const bel_list* lref;
    int mapkind;
} val;

typedef struct blk {
    struct blk* next;
} blk;
typedef struct {
    blk* head;
} arena;
static void* aalloc(arena* A, size_t n) {
    blk* b = (blk*)malloc(sizeof(blk) + n);
    b->next = A->head;
    A->head = b;
    return b + 1;
}
static void afree(arena* A) {
    for (blk* b = A->head; b;) { blk* nx = b->next; free(b); b = nx; }
    A->head = NULL;
}

static val v_err(void) { val v; memset(&v, 0, sizeof v); v.t = V_ERR; return v; }
static val v_bool(int b) { val v; memset(&v, 0, sizeof v); v.t = V_BOOL; v.i = b != 0; return v; }
static val v_int(int64_t i) { val v; memset(&v, 0, sizeof v); v.t = V_INT; v.i = i; return v; }
static val v_str(const uint8_t* s, size_t n) { val v; memset(&v, 0, sizeof v); v.t = V_STR; v.s = s; v.n = n; return v; }

static const char* const FIELD_NAMES[5] = {"host", "url", "path", "method", "user_agent"};

static size_t list_len(const val* l) { return l->lref ? l->lref->n : l->n_items; }

static val list_get(const val* l, size_t i) {
    if (!l->lref) return l->items[i];
    const bel_list* L = l->lref;
    val v;
    memset(&v, 0, sizeof v);
    switch (L->type) {
        case BEL_LIST_STRING: return v_str((const uint8_t*)L->strs[i], L->str_lens[i]);
        case BEL_LIST_INT: return v_int(L->ints[i]);
        default: v.t = V_IPNET; v.net = &L->nets[i]; return v;
    }
}

/* 1 equal, 0 different, -1 values of different types */
static int val_equal(const val* a, const val* b) {
    if (a->t == V_IPNET && b->t == V_IP) return bel_ipnet_contains(a->net, b->ip, b->v6);
    if (a->t == V_IP && b->t == V_IPNET) return bel_ipnet_contains(b->net, a->ip, a->v6);
    if (a->t != b->t) return -1;
    switch (a->t) {
        case V_NULL: return 1;
        case V_BOOL: case V_INT: case V_UINT: return a->i == b->i;
        case V_FLOAT: return a->f == b->f;
        case V_STR: case V_BYTES: return a->n == b->n && memcmp(a->s, b->s, a->n) == 0;
        case V_IP: return a->v6 == b->v6 && memcmp(a->ip, b->ip, a->v6 ? 16 : 4) == 0;
        case V_LIST: {
            if (list_len(a) != list_len(b)) return 0;
            for (size_t i = 0; i < list_len(a); ++i) {
                val x = list_get(a, i), y = list_get(b, i);
                int r = val_equal(&x, &y);
                if (r != 1) return r < 0 ? -1 : 0;
            }
            return 1;
        }
        default: return -1;
    }
}

static val member_of(const val* base, const char* name, size_t nlen, const bel_ctx* c) {
    val v;
    memset(&v, 0, sizeof v);
    if (base->t != V_MAP) return v_err();
    if (base->mapkind == M_HTTP) {
        for (int f = 0; f < 5; ++f)
            if (strlen(FIELD_NAMES[f]) == nlen && !memcmp(FIELD_NAMES[f], name, nlen)) return v_str(c->str[f], c->len[f]);
        return v_err();
    }
    if (base->mapkind == M_CLIENT) {
        if (nlen == 2 && !memcmp(name, "ip", 2)) { v.t = V_IP; memcpy(v.ip, c->ip, 16); v.v6 = c->ip_is_v6; return v; }
        if (nlen == 11 && !memcmp(name, "remote_port", 11)) return v_int(c->remote_port);
        if (nlen == 3 && !memcmp(name, "asn", 3)) return v_int(c->asn);
        if (nlen == 7 && !memcmp(name, "country", 7)) return v_str((const uint8_t*)c->country, 2);
        return v_err();
    }
    if (base->mapkind == M_LISTS) {
        if (!c->lists) return v_err();
        for (size_t i = 0; i < c->lists->n_lists; ++i) {
            const bel_list* L = &c->lists->lists[i];
            if (strlen(L->name) == nlen && !memcmp(L->name, name, nlen)) { v.t = V_LIST; v.lref = L; return v; }
        }
        return v_err(); /* A6: missing key is a runtime error */
    }
    for (size_t i = 0; i + 1 < base->n_items; i += 2) {
        const val* k = &base->items[i];
        if (k->t == V_STR && k->n == nlen && !memcmp(k->s, name, nlen)) return base->items[i + 1];
    }
    return v_err();
}

static size_t utf8_count(const uint8_t* s, size_t n) {
    size_t c = 0;
    for (size_t i = 0; i < n; ++i) if ((s[i] & 0xC0) != 0x80) c++;
    return c;
}

static int bytes_find(const uint8_t* h, size_t hn, const uint8_t* nd, size_t nn) {
    if (nn == 0) return 1;
    if (nn > hn) return 0;
    for (size_t i = 0; i + nn <= hn; ++i)
        if (memcmp(h + i, nd, nn) == 0) return 1;
    return 0;
}

static val eval(const bel_expr* e, const bel_ctx* c, arena* A);

static val eval_method(const bel_expr* e, const bel_ctx* c, arena* A) {
    val recv = eval(e->kid[0], c, A);
    if (recv.t == V_ERR) return recv;
    int nargs = e->nkid - 1;
    val arg;
    memset(&arg, 0, sizeof arg);
    if (nargs >= 1) {
        arg = eval(e->kid[1], c, A);
        if (arg.t == V_ERR) return arg;
        for (int k = 2; k < e->nkid; ++k) {
            val x = eval(e->kid[k], c, A);
            if (x.t == V_ERR) return x;
        }
    }
    const char* fn = e->s;
    if (!strcmp(fn, "length")) {
        if (nargs != 0) return v_err();
        switch (recv.t) {
            case V_STR: return v_int((int64_t)utf8_count(recv.s, recv.n));
            case V_LIST: return v_int((int64_t)list_len(&recv));
            case V_MAP:
                if (recv.mapkind == M_HTTP) return v_int(5);
                if (recv.mapkind == M_CLIENT) return v_int(4);
                if (recv.mapkind == M_LISTS) return v_int(c->lists ? (int64_t)c->lists->n_lists : 0);
                return v_int((int64_t)recv.n_items / 2);
            default: return v_err();
        }
    }
    if (!strcmp(fn, "contains")) {
        if (nargs != 1) return v_err();
        if (recv.t == V_STR) {
            if (arg.t != V_STR) return v_err();
            return v_bool(bytes_find(recv.s, recv.n, arg.s, arg.n));
        }
        if (recv.t == V_LIST) {
            /* A2: any element equal; Ip-network elements match by containment; other types never equal */
            size_t n = list_len(&recv);
            for (size_t i = 0; i < n; ++i) {
                val it = list_get(&recv, i);
                if (val_equal(&it, &arg) == 1) return v_bool(1);
            }
            return v_bool(0);
        }
        if (recv.t == V_MAP) {
            if (arg.t != V_STR) return v_bool(0);
            val m = member_of(&recv, (const char*)arg.s, arg.n, c);
            return v_bool(m.t != V_ERR);
        }
        return v_err();
    }
    if (!strcmp(fn, "starts_with") || !strcmp(fn, "ends_with")) {
        if (nargs != 1 || recv.t != V_STR || arg.t != V_STR) return v_err();
        if (arg.n > recv.n) return v_bool(0);
        if (fn[0] == 's') return v_bool(memcmp(recv.s, arg.s, arg.n) == 0);
        return v_bool(memcmp(recv.s + recv.n - arg.n, arg.s, arg.n) == 0);
    }
    if (!strcmp(fn, "matches")) {
        if (nargs != 1 || recv.t != V_STR || arg.t != V_STR) return v_err();
        if (e->kid[1]->kind == K_STR) {
            if (!e->rx) return v_err(); /* pattern does not compile: runtime error (A1) */
            return v_bool(rx_is_match(e->rx, recv.s, recv.n));
        }
        int st;
        char msg[128];
        rx_prog* pr = rx_compile((const char*)arg.s, arg.n, &st, msg);
        if (!pr) return v_err();
        int r = rx_is_match(pr, recv.s, recv.n);
        rx_free(pr);
        return v_bool(r);
    }
    return v_err(); /* unknown method */
}

static val eval(const bel_expr* e, const bel_ctx* c, arena* A) {
    val v;
    memset(&v, 0, sizeof v);
    switch (e->kind) {
        case K_NULL: v.t = V_NULL; return v;
        case K_BOOL: return v_bool((int)e->i);
        case K_INT: return v_int(e->i);
        case K_UINT: v.t = V_UINT; v.i = e->i; return v;
        case K_FLOAT: v.t = V_FLOAT; v.f = e->f; return v;
        case K_STR: return v_str((const uint8_t*)e->s, e->slen);
        case K_BYTES: v.t = V_BYTES; v.s = (const uint8_t*)e->s; v.n = e->slen; return v;
        case K_IDENT:
            v.t = V_MAP;
            if (!strcmp(e->s, "http_request")) { v.mapkind = M_HTTP; return v; }
            if (!strcmp(e->s, "client")) { v.mapkind = M_CLIENT; return v; }
            if (!strcmp(e->s, "lists")) { v.mapkind = M_LISTS; return v; }
            return v_err(); /* undeclared reference */
        case K_MEMBER: {
            val b = eval(e->kid[0], c, A);
            if (b.t == V_ERR) return b;
            return member_of(&b, e->s, strlen(e->s), c);
        }
        case K_INDEX: {
            val b = eval(e->kid[0], c, A);
            if (b.t == V_ERR) return b;
            val ix = eval(e->kid[1], c, A);
            if (ix.t == V_ERR) return ix;
            if (b.t == V_MAP) {
                if (b.mapkind == M_LITERAL) {
                    for (size_t i = 0; i + 1 < b.n_items; i += 2)
                        if (val_equal(&b.items[i], &ix) == 1) return b.items[i + 1];
                    return v_err();
                }
                if (ix.t != V_STR) return v_err();
                return member_of(&b, (const char*)ix.s, ix.n, c);
            }
            if (b.t == V_LIST) {
                if (ix.t != V_INT || ix.i < 0 || (uint64_t)ix.i >= list_len(&b)) return v_err();
                val it = list_get(&b, (size_t)ix.i);
                if (it.t == V_IPNET) return v_err();
                return it;
            }
            return v_err();
        }
        case K_CALL: {
            for (int k = 0; k < e->nkid; ++k) {
                val x = eval(e->kid[k], c, A);
                if (x.t == V_ERR) return x;
            }
            return v_err(); /* no global functions in the documented language */
        }
        case K_METHOD: return eval_method(e, c, A);
        case K_NOT: {
            val x = eval(e->kid[0], c, A);
            if (x.t != V_BOOL) return v_err();
            return v_bool(!x.i);
        }
        case K_NEG: {
            val x = eval(e->kid[0], c, A);
            if (x.t == V_INT) {
                if (x.i == INT64_MIN) return v_err();
                return v_int(-x.i);
            }
            if (x.t == V_FLOAT) { x.f = -x.f; return x; }
            return v_err();
        }
        case K_AND: {
            val a = eval(e->kid[0], c, A);
            if (a.t != V_BOOL) return v_err();
            if (!a.i) return v_bool(0);
            val b = eval(e->kid[1], c, A);
            if (b.t != V_BOOL) return v_err();
            return b;
        }
        case K_OR: {
            val a = eval(e->kid[0], c, A);
            if (a.t != V_BOOL) return v_err();
            if (a.i) return v_bool(1);
            val b = eval(e->kid[1], c, A);
            if (b.t != V_BOOL) return v_err();
            return b;
        }
        case K_COND: {
            val k = eval(e->kid[0], c, A);
            if (k.t != V_BOOL) return v_err();
            return eval(e->kid[k.i ? 1 : 2], c, A);
        }
        case K_IN: {
            val a = eval(e->kid[0], c, A);
            if (a.t == V_ERR) return a;
            val b = eval(e->kid[1], c, A);
            if (b.t == V_ERR) return b;
            return v_err(); /* operator "@in" is not provided (rules/rules.rs:67-71) */
        }
        case K_EQ: case K_NE: case K_LT: case K_LE: case K_GT: case K_GE: {
            val a = eval(e->kid[0], c, A);
            if (a.t == V_ERR) return a;
            val b = eval(e->kid[1], c, A);
            if (b.t == V_ERR) return b;
            if (e->kind == K_EQ || e->kind == K_NE) {
                if (a.t == V_IPNET || b.t == V_IPNET || a.t == V_MAP || b.t == V_MAP) return v_err();
                int r = val_equal(&a, &b);
                if (r < 0) return v_err(); /* A4: cross-type comparison */
                return v_bool(e->kind == K_EQ ? r == 1 : r == 0);
            }
            if (a.t != b.t) return v_err();
            int cmp;
            if (a.t == V_INT) cmp = a.i < b.i ? -1 : a.i > b.i;
            else if (a.t == V_UINT) cmp = (uint64_t)a.i < (uint64_t)b.i ? -1 : (uint64_t)a.i > (uint64_t)b.i;
            else if (a.t == V_FLOAT) {
                if (a.f != a.f || b.f != b.f) return v_bool(0);
                cmp = a.f < b.f ? -1 : a.f > b.f;
            } else if (a.t == V_STR || a.t == V_BYTES) {
                size_t m = a.n < b.n ? a.n : b.n;
                cmp = memcmp(a.s, b.s, m);
                if (cmp == 0) cmp = a.n < b.n ? -1 : a.n > b.n;
            } else return v_err();
            switch (e->kind) {
                case K_LT: return v_bool(cmp < 0);
                case K_LE: return v_bool(cmp <= 0);
                case K_GT: return v_bool(cmp > 0);
                default: return v_bool(cmp >= 0);
            }
        }
        case K_ADD: case K_SUB: case K_MUL: case K_DIV: case K_MOD: {
            val a = eval(e->kid[0], c, A);
            if (a.t == V_ERR) return a;
            val b = eval(e->kid[1], c, A);
            if (b.t == V_ERR) return b;
            if (a.t == V_INT && b.t == V_INT) {
                int64_t r = 0;
                switch (e->kind) {
                    case K_ADD: if (__builtin_add_overflow(a.i, b.i, &r)) return v_err(); break;
                    case K_SUB: if (__builtin_sub_overflow(a.i, b.i, &r)) return v_err(); break;
                    case K_MUL: if (__builtin_mul_overflow(a.i, b.i, &r)) return v_err(); break;
                    case K_DIV: if (b.i == 0 || (a.i == INT64_MIN && b.i == -1)) return v_err(); r = a.i / b.i; break;
                    default: if (b.i == 0 || (a.i == INT64_MIN && b.i == -1)) return v_err(); r = a.i % b.i; break;
                }
                return v_int(r);
            }
            if (a.t == V_FLOAT && b.t == V_FLOAT) {
                switch (e->kind) {
                    case K_ADD: a.f += b.f; return a;
                    case K_SUB: a.f -= b.f; return a;
                    case K_MUL: a.f *= b.f; return a;
                    case K_DIV: a.f /= b.f; return a;
                    default: return v_err();
                }
            }
            if (a.t == V_STR && b.t == V_STR && e->kind == K_ADD) {
                uint8_t* s = (uint8_t*)aalloc(A, a.n + b.n + 1);
                memcpy(s, a.s, a.n);
                memcpy(s + a.n, b.s, b.n);
                return v_str(s, a.n + b.n);
            }
            if (a.t == V_LIST && b.t == V_LIST && e->kind == K_ADD && !a.lref && !b.lref) {
                val r = a;
                r.items = (val*)aalloc(A, sizeof(val) * (a.n_items + b.n_items + 1));
                memcpy(r.items, a.items, sizeof(val) * a.n_items);
                memcpy(r.items + a.n_items, b.items, sizeof(val) * b.n_items);
                r.n_items = a.n_items + b.n_items;
                return r;
            }
            return v_err();
        }
        case K_LIST: case K_MAP: {
            v.t = e->kind == K_LIST ? V_LIST : V_MAP;
            v.mapkind = M_LITERAL;
            v.n_items = (size_t)e->nkid;
            v.items = (val*)aalloc(A, sizeof(val) * ((size_t)e->nkid + 1));
            for (int k = 0; k < e->nkid; ++k) {
                v.items[k] = eval(e->kid[k], c, A);
                if (v.items[k].t == V_ERR) return v_err();
            }
            return v;
        }
    }
    return v_err();
}

int bel_eval_kind(const bel_expr* e, const bel_ctx* ctx) {
    arena A = {NULL};
    val v = eval(e, ctx, &A);
    int r = v.t == V_ERR ? 2 : v.t != V_BOOL ? 3 : (v.i ? 1 : 0);
    afree(&A);
    return r;
}

int bel_matches(const bel_expr* e, const bel_ctx* ctx) { return bel_eval_kind(e, ctx) == 1; }
