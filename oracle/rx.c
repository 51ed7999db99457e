/* ORACLE (test infrastructure): Rust-regex-syntax subset -> Pike VM.  See rx.h.
 * Independent of the product compiler (pingoo_b200/csrc/regex.cpp + dfa.cpp):
 * this one never builds a DFA; it simulates the NFA position set byte by byte. */
#include "rx.h"

#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- AST ------------------------------------------------------------------------ */
enum { N_EMPTY, N_SET, N_ASSERT, N_CAT, N_ALT, N_REP };
/* AS_WS .. AS_WEH: \b{start} = \<, \b{end} = \>, \b{start-half}, \b{end-half} (regex 1.10+, docs "Empty matches / word boundaries") */
enum { AS_BOT, AS_EOT, AS_BOL, AS_EOL, AS_WB, AS_NWB, AS_WS, AS_WE, AS_WSH, AS_WEH };

typedef struct node {
    int kind;
    uint8_t set[32];
    int as;
    struct node* a;
    struct node* b;
    int min, max; /* max < 0: unbounded */
    struct node* next_alloc;
} node;

typedef struct {
    int i, m, s, x;
} flags_t;

typedef struct {
    const unsigned char* p;
    size_t n, pos;
    node* allocs;
    jmp_buf jb;
    int status;
    char* err;
    int depth;
} parser;

__attribute__((noreturn)) static void pfail(parser* P, int status, const char* msg) {
    P->status = status;
    snprintf(P->err, 128, "%s at offset %zu", msg, P->pos);
    longjmp(P->jb, 1);
}

static node* mk(parser* P, int kind) {
    node* n = (node*)calloc(1, sizeof(node));
    if (!n) pfail(P, RX_TOO_BIG, "out of memory");
    n->kind = kind;
    n->next_alloc = P->allocs;
    P->allocs = n;
    return n;
}

static void set_add(uint8_t* s, unsigned c) { s[c >> 3] |= (uint8_t)(1u << (c & 7)); }
static int set_has(const uint8_t* s, unsigned c) { return (s[c >> 3] >> (c & 7)) & 1; }
static void set_range(uint8_t* s, unsigned lo, unsigned hi) {
    for (unsigned c = lo; c <= hi; ++c) set_add(s, c);
}
static void set_not(uint8_t* s) {
    for (int i = 0; i < 32; ++i) s[i] = (uint8_t)~s[i];
}
static void set_fold(uint8_t* s) {
    for (unsigned c = 'a'; c <= 'z'; ++c)
        if (set_has(s, c) || set_has(s, c - 32)) {
            set_add(s, c);
            set_add(s, c - 32);
        }
}
static int is_word(int c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }

static void perl_set(uint8_t* s, int k) {
    memset(s, 0, 32);
    if (k == 'd') set_range(s, '0', '9');
    else if (k == 's') {
        set_add(s, ' ');
        set_range(s, 9, 13); /* \t \n \v \f \r */
    } else {
        for (unsigned c = 0; c < 128; ++c)
            if (is_word((int)c)) set_add(s, c);
    }
}

/* \p{..} on ASCII haystacks (oracle.h: request strings are ASCII; a byte >= 0x80 is one opaque character).  Written from
 * the code charts, member by member: the ASCII characters of each general category, then the names that select them.
 * Returns 1 if the (normalised: lower case, no ' ', '_', '-') name is known and fills `s` with its ASCII members. */
static void set_chars(uint8_t* s, const char* chars) {
    for (; *chars; ++chars) set_add(s, (unsigned char)*chars);
}
static int uni_gc(const char* v, uint8_t* s) {
    static const char* PO = "!\"#%&'*,./:;?@\\";
#define V(x) (strcmp(v, x) == 0)
    if (V("l") || V("letter") || V("lc") || V("casedletter")) { set_range(s, 'A', 'Z'); set_range(s, 'a', 'z'); }
    else if (V("lu") || V("uppercaseletter")) set_range(s, 'A', 'Z');
    else if (V("ll") || V("lowercaseletter")) set_range(s, 'a', 'z');
    else if (V("n") || V("number") || V("nd") || V("decimalnumber") || V("digit")) set_range(s, '0', '9');
    else if (V("p") || V("punctuation") || V("punct")) { set_chars(s, PO); set_chars(s, "_-([{)]}"); }
    else if (V("pc") || V("connectorpunctuation")) set_chars(s, "_");
    else if (V("pd") || V("dashpunctuation")) set_chars(s, "-");
    else if (V("ps") || V("openpunctuation")) set_chars(s, "([{");
    else if (V("pe") || V("closepunctuation")) set_chars(s, ")]}");
    else if (V("po") || V("otherpunctuation")) set_chars(s, PO);
    else if (V("s") || V("symbol")) set_chars(s, "+<=>|~$^`");
    else if (V("sm") || V("mathsymbol")) set_chars(s, "+<=>|~");
    else if (V("sc") || V("currencysymbol")) set_chars(s, "$");
    else if (V("sk") || V("modifiersymbol")) set_chars(s, "^`");
    else if (V("z") || V("separator") || V("zs") || V("spaceseparator")) set_chars(s, " ");
    else if (V("c") || V("other") || V("cc") || V("control") || V("cntrl")) { set_range(s, 0, 31); set_add(s, 127); }
    else if (V("lt") || V("titlecaseletter") || V("lm") || V("modifierletter") || V("lo") || V("otherletter") || V("m") || V("mark") ||
             V("combiningmark") || V("mn") || V("nonspacingmark") || V("mc") || V("spacingmark") || V("me") || V("enclosingmark") || V("nl") ||
             V("letternumber") || V("no") || V("othernumber") || V("pi") || V("initialpunctuation") || V("pf") || V("finalpunctuation") ||
             V("so") || V("othersymbol") || V("zl") || V("lineseparator") || V("zp") || V("paragraphseparator") || V("cf") || V("format") ||
             V("cs") || V("surrogate") || V("co") || V("privateuse") || V("cn") || V("unassigned")) { /* no ASCII member */ }
    else return 0;
    return 1;
}
static int uni_script(const char* v, uint8_t* s) {
    /* every Unicode script except Latin and Common: no ASCII member (checked against the `regex` module by tests/test_oracle.py) */
    static const char* none[] = {
        "adlam", "ahom", "anatolianhieroglyphs", "arabic", "armenian", "avestan", "balinese", "bamum", "bassavah", "batak", "bengali", "bhaiksuki",
        "bopomofo", "brahmi", "braille", "buginese", "buhid", "canadianaboriginal", "carian", "caucasianalbanian", "chakma", "cham", "cherokee", "chorasmian",
        "coptic", "cuneiform", "cypriot", "cyprominoan", "cyrillic", "deseret", "devanagari", "divesakuru", "dogra", "duployan", "egyptianhieroglyphs", "elbasan",
        "elymaic", "ethiopic", "georgian", "glagolitic", "gothic", "grantha", "greek", "gujarati", "gunjalagondi", "gurmukhi", "han", "hangul",
        "hanifirohingya", "hanunoo", "hatran", "hebrew", "hiragana", "imperialaramaic", "inherited", "inscriptionalpahlavi", "inscriptionalparthian", "javanese", "kaithi", "kannada",
        "katakana", "kawi", "kayahli", "kharoshthi", "khitansmallscript", "khmer", "khojki", "khudawadi", "lao", "lepcha", "limbu", "lineara",
        "linearb", "lisu", "lycian", "lydian", "mahajani", "makasar", "malayalam", "mandaic", "manichaean", "marchen", "masaramgondi", "medefaidrin",
        "meeteimayek", "mendekikakui", "meroiticcursive", "meroitichieroglyphs", "miao", "modi", "mongolian", "mro", "multani", "myanmar", "nabataean", "nagmundari",
        "nandinagari", "newa", "newtailue", "nko", "nushu", "nyiakengpuachuehmong", "ogham", "olchiki", "oldhungarian", "olditalic", "oldnortharabian", "oldpermic",
        "oldpersian", "oldsogdian", "oldsoutharabian", "oldturkic", "olduyghur", "oriya", "osage", "osmanya", "pahawhhmong", "palmyrene", "paucinhau", "phagspa",
        "phoenician", "psalterpahlavi", "rejang", "runic", "samaritan", "saurashtra", "sharada", "shavian", "siddham", "signwriting", "sinhala", "sogdian",
        "sorasompeng", "soyombo", "sundanese", "sylotinagri", "syriac", "tagalog", "tagbanwa", "taile", "taitham", "taiviet", "takri", "tamil",
        "tangsa", "tangut", "telugu", "thaana", "thai", "tibetan", "tifinagh", "tirhuta", "toto", "ugaritic", "vai", "vithkuqi",
        "wancho", "warangciti", "yezidi", "yi", "zanabazarsquare", "grek", "cyrl", "hani", "arab", "hebr", "hira", "kana",
        "deva", "hang", "armn", "geor", "ethi", "beng", "taml", "telu", "gujr", "guru", "knda", "mlym",
        "sinh", "khmr", "laoo", "tibt", "mymr", "mong", "syrc", "thaa", "copt", "cher", "bopo", "brai",
        "zinh", "qaai",
        NULL};
    if (V("latin") || V("latn")) { set_range(s, 'A', 'Z'); set_range(s, 'a', 'z'); return 1; }
    if (V("common") || V("zyyy")) {
        for (unsigned c = 0; c < 128; ++c)
            if (!((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) set_add(s, c);
        return 1;
    }
    for (int i = 0; none[i]; ++i)
        if (strcmp(v, none[i]) == 0) return 1;
    return 0;
}
/* 0 ok, 1 unknown name, 2 malformed */
static int uni_class(const char* raw, size_t len, uint8_t* s, int* negated) {
    char key[64], val[64];
    size_t nk = 0, nv = 0;
    int neg = 0, have_key = 0;
    memset(s, 0, 32);
    size_t i = 0;
    if (i < len && raw[i] == '^') { neg = 1; ++i; }
    for (; i < len; ++i) {
        char ch = raw[i];
        if (!have_key && (ch == '=' || ch == ':')) {
            if (ch == '=' && nv > 0 && val[nv - 1] == '!') { neg = !neg; --nv; }
            memcpy(key, val, nv);
            nk = nv;
            nv = 0;
            have_key = 1;
            continue;
        }
        if (ch == ' ' || ch == '_' || ch == '-') continue;
        if (ch >= 'A' && ch <= 'Z') ch = (char)(ch + 32);
        if (nv + 1 >= sizeof val) return 1;
        val[nv++] = ch;
    }
    key[nk] = 0;
    val[nv] = 0;
    if (nv == 0) return 2;
    const char* v = val;
    int ok;
    if (have_key) {
        if (!strcmp(key, "gc") || !strcmp(key, "generalcategory")) ok = uni_gc(val, s);
        else if (!strcmp(key, "sc") || !strcmp(key, "script") || !strcmp(key, "scx") || !strcmp(key, "scriptextensions")) ok = uni_script(val, s);
        else return 1;
    } else if (V("any")) { memset(s, 0xFF, 32); ok = 1; }
    else if (V("ascii")) { set_range(s, 0, 127); ok = 1; }
    else if (V("assigned")) { memset(s, 0xFF, 32); ok = 1; }
    else if (V("alphabetic") || V("alpha") || V("cased")) { set_range(s, 'A', 'Z'); set_range(s, 'a', 'z'); ok = 1; }
    else if (V("uppercase") || V("upper")) { set_range(s, 'A', 'Z'); ok = 1; }
    else if (V("lowercase") || V("lower")) { set_range(s, 'a', 'z'); ok = 1; }
    else if (V("whitespace") || V("wspace") || V("space")) { set_add(s, ' '); set_range(s, 9, 13); ok = 1; }
    else if (V("hexdigit") || V("hex") || V("asciihexdigit") || V("ahex")) { set_range(s, '0', '9'); set_range(s, 'A', 'F'); set_range(s, 'a', 'f'); ok = 1; }
    else ok = uni_gc(val, s) || uni_script(val, s);
#undef V
    if (!ok) return 1;
    *negated = neg; /* applied by the caller AFTER case folding (regex-syntax: fold, then negate) */
    return 0;
}

static int at_end(parser* P) { return P->pos >= P->n; }
static int peekc(parser* P) { return at_end(P) ? -1 : P->p[P->pos]; }
static int peek2(parser* P) { return P->pos + 1 < P->n ? P->p[P->pos + 1] : -1; }

static node* parse_alt(parser* P, flags_t* f);

static node* lit_byte(parser* P, unsigned c, const flags_t* f) {
    node* n = mk(P, N_SET);
    set_add(n->set, c);
    if (f->i) set_fold(n->set);
    return n;
}

static node* cat2(parser* P, node* a, node* b) {
    if (!a) return b;
    node* n = mk(P, N_CAT);
    n->a = a;
    n->b = b;
    return n;
}

/* a code point as a literal; non-ASCII becomes its UTF-8 byte sequence (case-sensitive) */
static node* lit_cp(parser* P, uint32_t cp, const flags_t* f) {
    if (cp < 0x80) return lit_byte(P, cp, f);
    if (cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) pfail(P, RX_INVALID, "invalid code point");
    unsigned char b[4];
    int n;
    if (cp < 0x800) { b[0] = (unsigned char)(0xC0 | (cp >> 6)); b[1] = (unsigned char)(0x80 | (cp & 63)); n = 2; }
    else if (cp < 0x10000) { b[0] = (unsigned char)(0xE0 | (cp >> 12)); b[1] = (unsigned char)(0x80 | ((cp >> 6) & 63)); b[2] = (unsigned char)(0x80 | (cp & 63)); n = 3; }
    else { b[0] = (unsigned char)(0xF0 | (cp >> 18)); b[1] = (unsigned char)(0x80 | ((cp >> 12) & 63)); b[2] = (unsigned char)(0x80 | ((cp >> 6) & 63)); b[3] = (unsigned char)(0x80 | (cp & 63)); n = 4; }
    flags_t nf = *f;
    nf.i = 0;
    node* acc = NULL;
    for (int k = 0; k < n; ++k) acc = cat2(P, acc, lit_byte(P, b[k], &nf));
    return acc;
}

static int hexv(int c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

static uint32_t parse_hex(parser* P, int kind) {
    uint32_t v = 0;
    if (peekc(P) == '{') {
        P->pos++;
        int nd = 0;
        while (!at_end(P) && peekc(P) != '}') {
            int h = hexv(peekc(P));
            if (h < 0) pfail(P, RX_INVALID, "invalid hexadecimal digit");
            v = v * 16 + (uint32_t)h;
            if (++nd > 8) pfail(P, RX_INVALID, "hexadecimal literal too long");
            P->pos++;
        }
        if (at_end(P)) pfail(P, RX_INVALID, "unclosed hexadecimal literal");
        P->pos++;
        if (!nd) pfail(P, RX_INVALID, "empty hexadecimal literal");
        return v;
    }
    int want = kind == 'x' ? 2 : kind == 'u' ? 4 : 8;
    for (int k = 0; k < want; ++k) {
        int h = at_end(P) ? -1 : hexv(peekc(P));
        if (h < 0) pfail(P, RX_INVALID, "invalid hexadecimal literal");
        v = v * 16 + (uint32_t)h;
        P->pos++;
    }
    return v;
}

/* escape result */
enum { E_LIT, E_SET, E_ASSERT };
typedef struct {
    int kind;
    uint32_t cp;
    uint8_t set[32];
    int as;
} esc_t;

static esc_t parse_escape(parser* P, int in_class, const flags_t* f) {
    esc_t e;
    memset(&e, 0, sizeof e);
    if (at_end(P)) pfail(P, RX_INVALID, "incomplete escape sequence");
    int c = P->p[P->pos++];
    switch (c) {
        case 'd': case 's': case 'w': e.kind = E_SET; perl_set(e.set, c); return e;
        case 'D': case 'S': case 'W': e.kind = E_SET; perl_set(e.set, c + 32); set_not(e.set); return e;
        case 'n': e.cp = '\n'; return e;
        case 't': e.cp = '\t'; return e;
        case 'r': e.cp = '\r'; return e;
        case 'a': e.cp = 7; return e;
        case 'f': e.cp = 12; return e;
        case 'v': e.cp = 11; return e;
        case 'x': case 'u': case 'U': e.cp = parse_hex(P, c); return e;
        case 'p': case 'P': {
            /* \pL, \p{Letter}, \p{^L}, \p{gc=Lu}, \p{sc:Latin}; \P negates */
            const char* nm;
            size_t len;
            if (at_end(P)) pfail(P, RX_INVALID, "incomplete Unicode class");
            if (P->p[P->pos] == '{') {
                size_t q = P->pos + 1;
                while (q < P->n && P->p[q] != '}') ++q;
                if (q >= P->n) pfail(P, RX_INVALID, "unclosed Unicode class");
                nm = (const char*)P->p + P->pos + 1;
                len = q - P->pos - 1;
                P->pos = q + 1;
            } else {
                nm = (const char*)P->p + P->pos;
                len = 1;
                P->pos++;
            }
            int neg = 0;
            int rc = uni_class(nm, len, e.set, &neg);
            if (rc == 2) pfail(P, RX_INVALID, "malformed Unicode class");
            if (rc == 1) pfail(P, RX_UNSUPPORTED, "Unicode property outside the known set");
            if (f->i) set_fold(e.set);
            if (neg != (c == 'P')) set_not(e.set);
            e.kind = E_SET;
            return e;
        }
        case 'A': case 'z': case 'b': case 'B':
            if (in_class) pfail(P, RX_INVALID, "unrecognized escape sequence in class");
            e.kind = E_ASSERT;
            e.as = c == 'A' ? AS_BOT : c == 'z' ? AS_EOT : c == 'b' ? AS_WB : AS_NWB;
            if (c == 'b' && peekc(P) == '{') {
                /* a name made of letters and '-' up to the closing brace; a digit after the brace would be a counted
                 * repetition of \b itself, which neither this oracle nor the engine implements */
                size_t q = P->pos + 1, q0 = q;
                if (q >= P->n || !((P->p[q] | 32) >= 'a' && (P->p[q] | 32) <= 'z') ) {
                    if (q < P->n && P->p[q] == '-') { /* falls through to the name scan below */ }
                    else pfail(P, RX_UNSUPPORTED, "a counted repetition of \\b is not supported");
                }
                while (q < P->n && ((((P->p[q] | 32) >= 'a') && ((P->p[q] | 32) <= 'z')) || P->p[q] == '-')) ++q;
                if (q >= P->n || P->p[q] != '}') pfail(P, RX_INVALID, "unclosed or malformed special word boundary");
                size_t len = q - q0;
                const char* nm = (const char*)P->p + q0;
                if (len == 5 && !memcmp(nm, "start", 5)) e.as = AS_WS;
                else if (len == 3 && !memcmp(nm, "end", 3)) e.as = AS_WE;
                else if (len == 10 && !memcmp(nm, "start-half", 10)) e.as = AS_WSH;
                else if (len == 8 && !memcmp(nm, "end-half", 8)) e.as = AS_WEH;
                else pfail(P, RX_INVALID, "unrecognized special word boundary assertion");
                P->pos = q + 1;
            }
            return e;
        case '<': case '>':
            if (in_class) pfail(P, RX_INVALID, "unrecognized escape sequence in class");
            e.kind = E_ASSERT;
            e.as = c == '<' ? AS_WS : AS_WE;
            return e;
        default: break;
    }
    if (c >= '0' && c <= '9') pfail(P, RX_INVALID, "backreferences are not supported");
    /* regex-syntax `is_escapeable_character`: any ASCII character but letters and digits (`<` `>` are assertions, above) */
    if (c < 0x80 && !((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'))) {
        e.cp = (uint32_t)c;
        return e;
    }
    pfail(P, RX_INVALID, "unrecognized escape sequence");
    return e;
}

static int posix_named(const char* name, size_t len, uint8_t* s) {
    memset(s, 0, 32);
#define IS(x) (len == strlen(x) && memcmp(name, x, len) == 0)
    if (IS("alnum")) { set_range(s, '0', '9'); set_range(s, 'A', 'Z'); set_range(s, 'a', 'z'); }
    else if (IS("alpha")) { set_range(s, 'A', 'Z'); set_range(s, 'a', 'z'); }
    else if (IS("ascii")) set_range(s, 0, 127);
    else if (IS("blank")) { set_add(s, ' '); set_add(s, '\t'); }
    else if (IS("cntrl")) { set_range(s, 0, 31); set_add(s, 127); }
    else if (IS("digit")) set_range(s, '0', '9');
    else if (IS("graph")) set_range(s, '!', '~');
    else if (IS("lower")) set_range(s, 'a', 'z');
    else if (IS("print")) set_range(s, ' ', '~');
    else if (IS("punct")) { set_range(s, '!', '/'); set_range(s, ':', '@'); set_range(s, '[', '`'); set_range(s, '{', '~'); }
    else if (IS("space")) { set_add(s, ' '); set_range(s, 9, 13); }
    else if (IS("upper")) set_range(s, 'A', 'Z');
    else if (IS("word")) { set_range(s, '0', '9'); set_range(s, 'A', 'Z'); set_range(s, 'a', 'z'); set_add(s, '_'); }
    else if (IS("xdigit")) { set_range(s, '0', '9'); set_range(s, 'A', 'F'); set_range(s, 'a', 'f'); }
    else return 0;
#undef IS
    return 1;
}

static void class_ws(parser* P, const flags_t* f) {
    if (!f->x) return;
    while (!at_end(P)) {
        int c = peekc(P);
        if (c == ' ' || (c >= 9 && c <= 13)) P->pos++;
        else break;
    }
}

static void parse_class(parser* P, const flags_t* f, uint8_t* out);

/* one item: returns 1 if it produced a set (not usable as a range endpoint) */
static int class_item(parser* P, const flags_t* f, uint32_t* cp, uint8_t* set) {
    int c = peekc(P);
    if (c == '[') {
        if (peek2(P) == ':') {
            size_t save = P->pos;
            P->pos += 2;
            int neg = 0;
            if (peekc(P) == '^') { neg = 1; P->pos++; }
            size_t st = P->pos;
            while (!at_end(P) && peekc(P) != ':' && peekc(P) != ']') P->pos++;
            if (peekc(P) == ':' && peek2(P) == ']' && posix_named((const char*)P->p + st, P->pos - st, set)) {
                P->pos += 2;
                if (neg) set_not(set);
                return 1;
            }
            P->pos = save;
        }
        P->pos++;
        parse_class(P, f, set);
        return 1;
    }
    if (c == '\\') {
        P->pos++;
        esc_t e = parse_escape(P, 1, f);
        if (e.kind == E_SET) { memcpy(set, e.set, 32); return 1; }
        *cp = e.cp;
        return 0;
    }
    if (c < 0x80) { P->pos++; *cp = (uint32_t)c; return 0; }
    int n = c >= 0xF0 ? 4 : c >= 0xE0 ? 3 : c >= 0xC0 ? 2 : 0;
    if (!n || P->pos + (size_t)n > P->n) pfail(P, RX_INVALID, "invalid UTF-8 in pattern");
    uint32_t v = (uint32_t)c & (0xFFu >> (n + 1));
    for (int k = 1; k < n; ++k) {
        int cc = P->p[P->pos + (size_t)k];
        if ((cc & 0xC0) != 0x80) pfail(P, RX_INVALID, "invalid UTF-8 in pattern");
        v = (v << 6) | (uint32_t)(cc & 63);
    }
    P->pos += (size_t)n;
    *cp = v;
    return 0;
}

static int is_setop(parser* P) {
    int c = peekc(P), d = peek2(P);
    return (c == '&' && d == '&') || (c == '-' && d == '-') || (c == '~' && d == '~');
}

static void class_union(parser* P, const flags_t* f, uint8_t* acc, int first) {
    memset(acc, 0, 32);
    for (;;) {
        class_ws(P, f);
        if (at_end(P)) pfail(P, RX_INVALID, "unclosed character class");
        int c = peekc(P);
        if (c == ']' && !first) return;
        if (is_setop(P)) return;
        uint32_t lo = 0;
        uint8_t sub[32];
        int isset;
        if (c == ']' && first) { P->pos++; lo = ']'; isset = 0; }
        else if (c == '-') { P->pos++; lo = '-'; isset = 0; }
        else isset = class_item(P, f, &lo, sub);
        first = 0;
        if (isset) {
            for (int i = 0; i < 32; ++i) acc[i] |= sub[i];
            continue;
        }
        class_ws(P, f);
        if (peekc(P) == '-' && peek2(P) != ']' && peek2(P) != -1 && peek2(P) != '-') {
            P->pos++;
            class_ws(P, f);
            if (at_end(P)) pfail(P, RX_INVALID, "unclosed character class");
            uint32_t hi = 0;
            uint8_t sub2[32];
            if (peekc(P) == '[' || class_item(P, f, &hi, sub2)) pfail(P, RX_INVALID, "invalid character class range");
            if (hi < lo) pfail(P, RX_INVALID, "invalid character class range");
            for (uint32_t v = lo; v <= hi && v < 0x80; ++v) set_add(acc, v);
            continue;
        }
        if (lo < 0x80) set_add(acc, lo);
    }
}

/* P->pos is just past '[' */
static void parse_class(parser* P, const flags_t* f, uint8_t* out) {
    int neg = 0;
    if (peekc(P) == '^') { neg = 1; P->pos++; }
    uint8_t acc[32], rhs[32];
    class_union(P, f, acc, 1);
    for (;;) {
        if (at_end(P)) pfail(P, RX_INVALID, "unclosed character class");
        int c = peekc(P);
        if (c == ']') { P->pos++; break; }
        P->pos += 2;
        class_union(P, f, rhs, 0);
        for (int i = 0; i < 32; ++i) {
            if (c == '&') acc[i] &= rhs[i];
            else if (c == '-') acc[i] &= (uint8_t)~rhs[i];
            else acc[i] ^= rhs[i];
        }
    }
    for (int i = 16; i < 32; ++i) acc[i] = 0; /* positive members are ASCII only */
    if (f->i) set_fold(acc);
    if (neg) set_not(acc);
    memcpy(out, acc, 32);
}

static void skip_ws(parser* P, const flags_t* f) {
    if (!f->x) return;
    while (!at_end(P)) {
        int c = peekc(P);
        if (c == ' ' || (c >= 9 && c <= 13)) { P->pos++; continue; }
        if (c == '#') {
            while (!at_end(P) && peekc(P) != '\n') P->pos++;
            continue;
        }
        break;
    }
}

static int parse_dec(parser* P, int* out) {
    size_t st = P->pos;
    long v = 0;
    while (!at_end(P) && peekc(P) >= '0' && peekc(P) <= '9') {
        v = v * 10 + (peekc(P) - '0');
        if (v > 100000000) pfail(P, RX_INVALID, "repetition count too large");
        P->pos++;
    }
    if (P->pos == st) return 0;
    *out = (int)v;
    return 1;
}

static void skip_sp(parser* P) {
    while (peekc(P) == ' ') P->pos++;
}

/* returns NULL for a bare flag group "(?i)" */
static node* parse_atom(parser* P, flags_t* f, int* produced) {
    *produced = 1;
    int c = peekc(P);
    if (c == '(') {
        P->pos++;
        flags_t inner = *f;
        if (peekc(P) == '?') {
            P->pos++;
            if (at_end(P)) pfail(P, RX_INVALID, "unclosed group");
            int d = peekc(P);
            if (d == 'P' || d == '<') {
                if (d == 'P') {
                    P->pos++;
                    if (peekc(P) != '<') pfail(P, RX_INVALID, "invalid named group");
                } else if (peek2(P) == '=' || peek2(P) == '!') pfail(P, RX_INVALID, "look-around is not supported");
                P->pos++;
                size_t st = P->pos;
                while (!at_end(P) && peekc(P) != '>') P->pos++;
                if (at_end(P) || P->pos == st) pfail(P, RX_INVALID, "invalid capture group name");
                P->pos++;
            } else if (d == '=' || d == '!') {
                pfail(P, RX_INVALID, "look-around is not supported");
            } else if (d == ':') {
                P->pos++; /* plain non-capturing group */
            } else {
                int neg = 0, any = 0, scoped = -1;
                while (scoped < 0) {
                    if (at_end(P)) pfail(P, RX_INVALID, "unclosed group");
                    int fc = P->p[P->pos++];
                    if (fc == ')' || fc == ':') {
                        if (!any) pfail(P, RX_INVALID, "missing flags");
                        scoped = fc == ':';
                        break;
                    }
                    if (fc == '-') {
                        if (neg) pfail(P, RX_INVALID, "repeated flag negation");
                        neg = 1;
                        any = 0;
                        continue;
                    }
                    int v = !neg;
                    switch (fc) {
                        case 'i': inner.i = v; break;
                        case 'm': inner.m = v; break;
                        case 's': inner.s = v; break;
                        case 'x': inner.x = v; break;
                        case 'U': case 'u': break; /* greediness / unicode toggles do not change ASCII match existence */
                        case 'R': pfail(P, RX_UNSUPPORTED, "CRLF mode is not supported");
                        default: pfail(P, RX_INVALID, "unrecognized flag");
                    }
                    any = 1;
                }
                if (!scoped) {
                    *f = inner;
                    *produced = 0;
                    return NULL;
                }
            }
        }
        if (++P->depth > 200) pfail(P, RX_INVALID, "nesting too deep");
        node* r = parse_alt(P, &inner);
        P->depth--;
        if (peekc(P) != ')') pfail(P, RX_INVALID, "unclosed group");
        P->pos++;
        return r;
    }
    if (c == '[') {
        P->pos++;
        node* n = mk(P, N_SET);
        parse_class(P, f, n->set);
        return n;
    }
    if (c == '.') {
        P->pos++;
        node* n = mk(P, N_SET);
        memset(n->set, 0xFF, 32);
        if (!f->s) n->set['\n' >> 3] &= (uint8_t)~(1u << ('\n' & 7));
        return n;
    }
    if (c == '^' || c == '$') {
        P->pos++;
        node* n = mk(P, N_ASSERT);
        n->as = c == '^' ? (f->m ? AS_BOL : AS_BOT) : (f->m ? AS_EOL : AS_EOT);
        return n;
    }
    if (c == '\\') {
        P->pos++;
        esc_t e = parse_escape(P, 0, f);
        if (e.kind == E_ASSERT) {
            node* n = mk(P, N_ASSERT);
            n->as = e.as;
            return n;
        }
        if (e.kind == E_SET) {
            node* n = mk(P, N_SET);
            memcpy(n->set, e.set, 32);
            return n;
        }
        return lit_cp(P, e.cp, f);
    }
    if (c < 0x80) {
        P->pos++;
        return lit_byte(P, (unsigned)c, f);
    }
    uint32_t cp = 0;
    uint8_t dummy[32];
    class_item(P, f, &cp, dummy);
    return lit_cp(P, cp, f);
}

static node* parse_cat(parser* P, flags_t* f) {
    node* acc = NULL;
    node* last = NULL; /* last item, target of a following repetition operator */
    node** last_slot = NULL;
    for (;;) {
        skip_ws(P, f);
        if (at_end(P)) break;
        int c = peekc(P);
        if (c == '|') break;
        if (c == ')') {
            if (P->depth == 0) pfail(P, RX_INVALID, "unopened group");
            break;
        }
        if (c == '*' || c == '+' || c == '?' || c == '{') {
            if (!last) pfail(P, RX_INVALID, "repetition operator missing expression");
            int mn = 0, mx = -1;
            P->pos++;
            if (c == '+') mn = 1;
            else if (c == '?') mx = 1;
            else if (c == '{') {
                skip_sp(P);
                if (!parse_dec(P, &mn)) pfail(P, RX_INVALID, "repetition quantifier expects a valid decimal");
                skip_sp(P);
                if (at_end(P)) pfail(P, RX_INVALID, "unclosed counted repetition");
                if (peekc(P) == ',') {
                    P->pos++;
                    skip_sp(P);
                    if (at_end(P)) pfail(P, RX_INVALID, "unclosed counted repetition");
                    if (peekc(P) != '}' && !parse_dec(P, &mx)) pfail(P, RX_INVALID, "repetition quantifier expects a valid decimal");
                    skip_sp(P);
                } else mx = mn;
                if (peekc(P) != '}') pfail(P, RX_INVALID, "unclosed counted repetition");
                P->pos++;
                if (mx >= 0 && mx < mn) pfail(P, RX_INVALID, "invalid repetition count range");
            }
            if (peekc(P) == '?') P->pos++; /* lazy marker */
            node* r = mk(P, N_REP);
            r->a = last;
            r->min = mn;
            r->max = mx;
            *last_slot = r;
            last = r;
            continue;
        }
        int produced;
        node* a = parse_atom(P, f, &produced);
        if (!produced) continue;
        if (!a) a = mk(P, N_EMPTY);
        if (!acc) {
            acc = a;
            last = a;
            last_slot = &acc;
        } else {
            node* cn = mk(P, N_CAT);
            cn->a = acc;
            cn->b = a;
            acc = cn;
            last = a;
            last_slot = &cn->b;
        }
    }
    return acc ? acc : mk(P, N_EMPTY);
}

static node* parse_alt(parser* P, flags_t* f) {
    node* left = parse_cat(P, f);
    while (peekc(P) == '|') {
        P->pos++;
        node* right = parse_cat(P, f);
        node* a = mk(P, N_ALT);
        a->a = left;
        a->b = right;
        left = a;
    }
    return left;
}

/* ---- Pike VM ------------------------------------------------------------------------ */
enum { I_SET, I_SPLIT, I_JMP, I_ASSERT, I_MATCH };
typedef struct {
    int op;
    int x, y;
    uint8_t set[32];
} inst;

struct rx_prog {
    inst* code;
    int n, cap;
    int start;
    int any_bot_only; /* unused optimisation hook */
};

#define RX_MAX_INST 200000

typedef struct {
    rx_prog* pr;
    parser* P;
} emitter;

static int emit(emitter* E, int op) {
    rx_prog* pr = E->pr;
    if (pr->n >= RX_MAX_INST) pfail(E->P, RX_TOO_BIG, "compiled regex exceeds size limit");
    if (pr->n == pr->cap) {
        pr->cap = pr->cap ? pr->cap * 2 : 64;
        pr->code = (inst*)realloc(pr->code, (size_t)pr->cap * sizeof(inst));
        if (!pr->code) pfail(E->P, RX_TOO_BIG, "out of memory");
    }
    memset(&pr->code[pr->n], 0, sizeof(inst));
    pr->code[pr->n].op = op;
    return pr->n++;
}

/* compile n so that it continues at `next`; returns entry pc */
static int comp(emitter* E, const node* n, int next) {
    switch (n->kind) {
        case N_EMPTY: return next;
        case N_SET: {
            int pc = emit(E, I_SET);
            memcpy(E->pr->code[pc].set, n->set, 32);
            E->pr->code[pc].x = next;
            return pc;
        }
        case N_ASSERT: {
            int pc = emit(E, I_ASSERT);
            E->pr->code[pc].y = n->as;
            E->pr->code[pc].x = next;
            return pc;
        }
        case N_CAT: {
            int b = comp(E, n->b, next);
            return comp(E, n->a, b);
        }
        case N_ALT: {
            int a = comp(E, n->a, next);
            int b = comp(E, n->b, next);
            int pc = emit(E, I_SPLIT);
            E->pr->code[pc].x = a;
            E->pr->code[pc].y = b;
            return pc;
        }
        case N_REP: {
            int tail = next;
            if (n->max < 0) {
                int sp = emit(E, I_SPLIT);
                int body = comp(E, n->a, sp);
                E->pr->code[sp].x = body;
                E->pr->code[sp].y = next;
                tail = sp;
            } else {
                for (int k = 0; k < n->max - n->min; ++k) {
                    int sp = emit(E, I_SPLIT);
                    int body = comp(E, n->a, tail);
                    E->pr->code[sp].x = body;
                    E->pr->code[sp].y = next;
                    tail = sp;
                }
            }
            int cur = tail;
            for (int k = 0; k < n->min; ++k) cur = comp(E, n->a, cur);
            return cur;
        }
    }
    return next;
}

rx_prog* rx_compile(const char* pat, size_t len, int* status, char* err) {
    parser P;
    memset(&P, 0, sizeof P);
    P.p = (const unsigned char*)pat;
    P.n = len;
    P.err = err;
    err[0] = 0;
    rx_prog* volatile pr = (rx_prog*)calloc(1, sizeof(rx_prog));
    emitter E = {pr, &P};
    if (setjmp(P.jb)) {
        *status = P.status;
        for (node* n = P.allocs; n;) { node* nx = n->next_alloc; free(n); n = nx; }
        free(pr->code);
        free(pr);
        return NULL;
    }
    flags_t f = {0, 0, 0, 0};
    node* root = parse_alt(&P, &f);
    if (!at_end(&P)) pfail(&P, RX_INVALID, "unopened group");
    int m = emit(&E, I_MATCH);
    pr->start = comp(&E, root, m);
    for (node* n = P.allocs; n;) { node* nx = n->next_alloc; free(n); n = nx; }
    *status = RX_OK;
    return pr;
}

void rx_free(rx_prog* p) {
    if (!p) return;
    free(p->code);
    free(p);
}

typedef struct {
    int* dense;
    int* sparse;
    int n;
} sset;

static int ss_has(const sset* s, int v) { unsigned i = (unsigned)s->sparse[v]; return i < (unsigned)s->n && s->dense[i] == v; }
static void ss_add(sset* s, int v) { s->sparse[v] = s->n; s->dense[s->n++] = v; }

/* follow epsilon edges from pc at position `pos`; returns 1 if MATCH is reachable */
static int addthread(const rx_prog* p, sset* list, int* stack, int pc0, const uint8_t* s, size_t n, size_t pos) {
    int sp = 0, matched = 0;
    stack[sp++] = pc0;
    while (sp) {
        int pc = stack[--sp];
        if (ss_has(list, pc)) continue;
        ss_add(list, pc);
        const inst* in = &p->code[pc];
        switch (in->op) {
            case I_JMP: stack[sp++] = in->x; break;
            case I_SPLIT: stack[sp++] = in->y; stack[sp++] = in->x; break;
            case I_ASSERT: {
                int prev_w = pos > 0 && is_word(s[pos - 1]);
                int next_w = pos < n && is_word(s[pos]);
                int ok = 0;
                switch (in->y) {
                    case AS_BOT: ok = pos == 0; break;
                    case AS_EOT: ok = pos == n; break;
                    case AS_BOL: ok = pos == 0 || s[pos - 1] == '\n'; break;
                    case AS_EOL: ok = pos == n || s[pos] == '\n'; break;
                    case AS_WB: ok = prev_w != next_w; break;
                    case AS_NWB: ok = prev_w == next_w; break;
                    case AS_WS: ok = !prev_w && next_w; break;
                    case AS_WE: ok = prev_w && !next_w; break;
                    case AS_WSH: ok = !prev_w; break;
                    case AS_WEH: ok = !next_w; break;
                }
                if (ok) stack[sp++] = in->x;
                break;
            }
            case I_MATCH: matched = 1; break;
            default: break;
        }
    }
    return matched;
}

int rx_is_match(const rx_prog* p, const uint8_t* s, size_t n) {
    int N = p->n;
    /* per-thread scratch: the baseline should time matching, not malloc */
    static __thread int* scratch = NULL;
    static __thread size_t scratch_cap = 0;
    size_t need = (size_t)N * 7 + 8;
    if (need > scratch_cap) {
        free(scratch);
        scratch = (int*)malloc(need * sizeof(int));
        scratch_cap = need;
    }
    int* mem = scratch;
    sset a = {mem, mem + N, 0}, b = {mem + 2 * N, mem + 3 * N, 0};
    int* stack = mem + 4 * N; /* every visited pc pushes at most two successors: 2N+1 entries suffice */
    sset *cur = &a, *nxt = &b;
    int found = 0;
    for (size_t pos = 0;; ++pos) {
        /* unanchored search: a new thread starts at every position */
        if (addthread(p, cur, stack, p->start, s, n, pos)) { found = 1; break; }
        if (pos == n) break;
        nxt->n = 0;
        unsigned c = s[pos];
        for (int i = 0; i < cur->n; ++i) {
            const inst* in = &p->code[cur->dense[i]];
            if (in->op == I_SET && set_has(in->set, c)) {
                if (addthread(p, nxt, stack, in->x, s, n, pos + 1)) { found = 1; break; }
            }
        }
        if (found) break;
        sset* t = cur; cur = nxt; nxt = t;
    }
    return found;
}
