// Rust-`regex` syntax subset -> AST -> Thompson NFA (see regex.hpp).
#include "regex.hpp"

#include <cctype>
#include <cstring>
#include <functional>

namespace pgw {

int Nfa::add_set(const ByteSet& s) {
    for (size_t i = 0; i < sets.size(); ++i)
        if (sets[i] == s) return (int)i;
    sets.push_back(s);
    return (int)sets.size() - 1;
}

namespace {

struct Flags {
    bool i = false, m = false, s = false, U = false, u = true, x = false;
};

struct Ast {
    enum Kind : uint8_t { EMPTY, SET, ASSERT, CONCAT, ALT, REPEAT } kind = EMPTY;
    ByteSet set;
    uint8_t ak = 0;
    std::vector<int> kids;
    int min = 0, max = -1;  // REPEAT; max -1 = unbounded
};

struct ParseFail {
    RegexStatus st;
    std::string msg;
};

static bool is_word_byte(unsigned c) {
    return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}

// ---- Unicode properties (\p{..}) on ASCII haystacks (SEMANTICS.md A9) ------------------------------------------------------
// Rust `regex` is Unicode-aware by default; with request strings that are ASCII, a property class is its ASCII members
// (bytes >= 0x80 stay one opaque character that only negated classes match).  General category of every ASCII code point
// (UnicodeData.txt, stable since Unicode 1.1), then the property names regex-syntax accepts (loose matching: case, spaces,
// '-' and '_' are ignored) mapped to sets of categories; scripts: Latin = the letters, Common = every other ASCII code point,
// any other known script has no ASCII member.
enum Gc : uint32_t { GC_Cc = 1u << 0, GC_Zs = 1u << 1, GC_Po = 1u << 2, GC_Sc = 1u << 3, GC_Ps = 1u << 4, GC_Pe = 1u << 5, GC_Sm = 1u << 6,
                     GC_Pd = 1u << 7, GC_Nd = 1u << 8, GC_Lu = 1u << 9, GC_Ll = 1u << 10, GC_Sk = 1u << 11, GC_Pc = 1u << 12 };
static uint32_t ascii_gc(unsigned c) {
    if (c < 0x20 || c == 0x7F) return GC_Cc;
    if (c == ' ') return GC_Zs;
    if (c >= '0' && c <= '9') return GC_Nd;
    if (c >= 'A' && c <= 'Z') return GC_Lu;
    if (c >= 'a' && c <= 'z') return GC_Ll;
    switch (c) {
        case '$': return GC_Sc;
        case '(': case '[': case '{': return GC_Ps;
        case ')': case ']': case '}': return GC_Pe;
        case '+': case '<': case '=': case '>': case '|': case '~': return GC_Sm;
        case '-': return GC_Pd;
        case '^': case '`': return GC_Sk;
        case '_': return GC_Pc;
        default: return GC_Po;   // ! " # % & ' * , . / : ; ? @ backslash
    }
}
struct PropName { const char* name; uint32_t gcs; };
static const PropName kGcNames[] = {
    {"l", GC_Lu | GC_Ll}, {"letter", GC_Lu | GC_Ll}, {"lu", GC_Lu}, {"uppercaseletter", GC_Lu}, {"ll", GC_Ll}, {"lowercaseletter", GC_Ll},
    {"lc", GC_Lu | GC_Ll}, {"casedletter", GC_Lu | GC_Ll}, {"lt", 0}, {"titlecaseletter", 0}, {"lm", 0}, {"modifierletter", 0}, {"lo", 0}, {"otherletter", 0},
    {"m", 0}, {"mark", 0}, {"combiningmark", 0}, {"mn", 0}, {"nonspacingmark", 0}, {"mc", 0}, {"spacingmark", 0}, {"me", 0}, {"enclosingmark", 0},
    {"n", GC_Nd}, {"number", GC_Nd}, {"nd", GC_Nd}, {"decimalnumber", GC_Nd}, {"digit", GC_Nd}, {"nl", 0}, {"letternumber", 0}, {"no", 0}, {"othernumber", 0},
    {"p", GC_Pc | GC_Pd | GC_Ps | GC_Pe | GC_Po}, {"punctuation", GC_Pc | GC_Pd | GC_Ps | GC_Pe | GC_Po}, {"punct", GC_Pc | GC_Pd | GC_Ps | GC_Pe | GC_Po},
    {"pc", GC_Pc}, {"connectorpunctuation", GC_Pc}, {"pd", GC_Pd}, {"dashpunctuation", GC_Pd}, {"ps", GC_Ps}, {"openpunctuation", GC_Ps},
    {"pe", GC_Pe}, {"closepunctuation", GC_Pe}, {"pi", 0}, {"initialpunctuation", 0}, {"pf", 0}, {"finalpunctuation", 0}, {"po", GC_Po}, {"otherpunctuation", GC_Po},
    {"s", GC_Sm | GC_Sc | GC_Sk}, {"symbol", GC_Sm | GC_Sc | GC_Sk}, {"sm", GC_Sm}, {"mathsymbol", GC_Sm}, {"sc", GC_Sc}, {"currencysymbol", GC_Sc},
    {"sk", GC_Sk}, {"modifiersymbol", GC_Sk}, {"so", 0}, {"othersymbol", 0},
    {"z", GC_Zs}, {"separator", GC_Zs}, {"zs", GC_Zs}, {"spaceseparator", GC_Zs}, {"zl", 0}, {"lineseparator", 0}, {"zp", 0}, {"paragraphseparator", 0},
    {"c", GC_Cc}, {"other", GC_Cc}, {"cc", GC_Cc}, {"control", GC_Cc}, {"cntrl", GC_Cc}, {"cf", 0}, {"format", 0}, {"cs", 0}, {"surrogate", 0},
    {"co", 0}, {"privateuse", 0}, {"cn", 0}, {"unassigned", 0},
};
static const char* const kOtherScripts[] = {   // every Unicode script but Latin and Common (names normalised; some ISO 15924 codes): none has an ASCII member
    "adlam", "ahom", "anatolianhieroglyphs", "arabic", "armenian", "avestan", "balinese", "bamum", "bassavah", "batak", "bengali", "bhaiksuki", "bopomofo", "brahmi",
    "braille", "buginese", "buhid", "canadianaboriginal", "carian", "caucasianalbanian", "chakma", "cham", "cherokee", "chorasmian", "coptic", "cuneiform", "cypriot", "cyprominoan",
    "cyrillic", "deseret", "devanagari", "divesakuru", "dogra", "duployan", "egyptianhieroglyphs", "elbasan", "elymaic", "ethiopic", "georgian", "glagolitic", "gothic", "grantha",
    "greek", "gujarati", "gunjalagondi", "gurmukhi", "han", "hangul", "hanifirohingya", "hanunoo", "hatran", "hebrew", "hiragana", "imperialaramaic", "inherited", "inscriptionalpahlavi",
    "inscriptionalparthian", "javanese", "kaithi", "kannada", "katakana", "kawi", "kayahli", "kharoshthi", "khitansmallscript", "khmer", "khojki", "khudawadi", "lao", "lepcha",
    "limbu", "lineara", "linearb", "lisu", "lycian", "lydian", "mahajani", "makasar", "malayalam", "mandaic", "manichaean", "marchen", "masaramgondi", "medefaidrin",
    "meeteimayek", "mendekikakui", "meroiticcursive", "meroitichieroglyphs", "miao", "modi", "mongolian", "mro", "multani", "myanmar", "nabataean", "nagmundari", "nandinagari", "newa",
    "newtailue", "nko", "nushu", "nyiakengpuachuehmong", "ogham", "olchiki", "oldhungarian", "olditalic", "oldnortharabian", "oldpermic", "oldpersian", "oldsogdian", "oldsoutharabian", "oldturkic",
    "olduyghur", "oriya", "osage", "osmanya", "pahawhhmong", "palmyrene", "paucinhau", "phagspa", "phoenician", "psalterpahlavi", "rejang", "runic", "samaritan", "saurashtra",
    "sharada", "shavian", "siddham", "signwriting", "sinhala", "sogdian", "sorasompeng", "soyombo", "sundanese", "sylotinagri", "syriac", "tagalog", "tagbanwa", "taile",
    "taitham", "taiviet", "takri", "tamil", "tangsa", "tangut", "telugu", "thaana", "thai", "tibetan", "tifinagh", "tirhuta", "toto", "ugaritic",
    "vai", "vithkuqi", "wancho", "warangciti", "yezidi", "yi", "zanabazarsquare", "grek", "cyrl", "hani", "arab", "hebr", "hira", "kana",
    "deva", "hang", "armn", "geor", "ethi", "beng", "taml", "telu", "gujr", "guru", "knda", "mlym", "sinh", "khmr",
    "laoo", "tibt", "mymr", "mong", "syrc", "thaa", "copt", "cher", "bopo", "brai", "zinh", "qaai",
};
// 0: ok, 1: not a name this engine knows (loud), 2: malformed
static int unicode_property_ascii(const std::string& raw, ByteSet* out, bool* negated) {
    std::string key, val;
    bool have_key = false, neg = false;
    auto norm = [](const std::string& t) {
        std::string r;
        for (char ch : t)
            if (ch != ' ' && ch != '_' && ch != '-') r.push_back((char)std::tolower((unsigned char)ch));
        return r;
    };
    std::string body = raw;
    if (!body.empty() && body[0] == '^') { neg = true; body.erase(0, 1); }
    size_t eq = body.find_first_of("=:");
    if (eq != std::string::npos) {
        have_key = true;
        size_t kend = eq;
        if (body[eq] == '=' && eq > 0 && body[eq - 1] == '!') { neg = !neg; kend = eq - 1; }
        key = norm(body.substr(0, kend));
        val = norm(body.substr(eq + 1));
    } else val = norm(body);
    if (val.empty()) return 2;
    ByteSet s;
    auto by_gc = [&](uint32_t gcs) { for (unsigned c = 0; c < 128; ++c) if (ascii_gc(c) & gcs) s.set(c); };
    auto script = [&](const std::string& v) -> bool {
        if (v == "latin" || v == "latn") { by_gc(GC_Lu | GC_Ll); return true; }
        if (v == "common" || v == "zyyy") { by_gc(~(uint32_t)(GC_Lu | GC_Ll)); return true; }
        for (const char* o : kOtherScripts) if (v == o) return true;
        return false;
    };
    auto gc = [&](const std::string& v) -> bool {
        for (const PropName& pn : kGcNames) if (v == pn.name) { by_gc(pn.gcs); return true; }
        return false;
    };
    bool ok = false;
    if (have_key) {
        if (key == "gc" || key == "generalcategory") ok = gc(val);
        else if (key == "sc" || key == "script" || key == "scx" || key == "scriptextensions") ok = script(val);
        else return 1;
    } else if (val == "any") { s.negate(); ok = true; }   // every code point, the opaque one included
    else if (val == "ascii" || val == "assigned") { s.set_range(0, 127); ok = true; if (val == "assigned") { for (unsigned c = 128; c < 256; ++c) s.set(c); } }
    else if (val == "alphabetic" || val == "alpha" || val == "cased") { by_gc(GC_Lu | GC_Ll); ok = true; }
    else if (val == "uppercase" || val == "upper") { by_gc(GC_Lu); ok = true; }
    else if (val == "lowercase" || val == "lower") { by_gc(GC_Ll); ok = true; }
    else if (val == "whitespace" || val == "wspace" || val == "space") { s.set(' '); s.set_range(9, 13); ok = true; }
    else if (val == "hexdigit" || val == "hex" || val == "asciihexdigit" || val == "ahex") { s.set_range('0', '9'); s.set_range('A', 'F'); s.set_range('a', 'f'); ok = true; }
    else ok = gc(val) || script(val);
    if (!ok) return 1;
    *out = s;
    *negated = neg;   // the caller folds case first and negates afterwards, as regex-syntax does (hir/translate.rs unicode_fold_and_negate)
    return 0;
}

static ByteSet perl_class(char k) {
    ByteSet s;
    switch (k) {
        case 'd': s.set_range('0', '9'); break;
        case 's':
            s.set('\t'); s.set('\n'); s.set(0x0B); s.set(0x0C); s.set('\r'); s.set(' ');
            break;
        case 'w':
            for (unsigned c = 0; c < 128; ++c)
                if (is_word_byte(c)) s.set(c);
            break;
    }
    return s;
}

static void fold_case(ByteSet& s) {
    for (unsigned c = 'a'; c <= 'z'; ++c) {
        unsigned C = c - 32;
        if (s.test(c) || s.test(C)) { s.set(c); s.set(C); }
    }
}

class Parser {
  public:
    Parser(const std::string& p, std::vector<Ast>& pool, RegexInfo* info) : p_(p), pool_(pool), info_(info) {}

    int parse() {
        Flags f;
        int r = parse_alt(f, 0);
        if (pos_ < p_.size()) {
            // only an unmatched ')' can stop parse_alt at depth 0
            fail(RX_INVALID, "unopened group");
        }
        return r;
    }

  private:
    const std::string& p_;
    std::vector<Ast>& pool_;
    RegexInfo* info_;
    size_t pos_ = 0;

    [[noreturn]] void fail(RegexStatus st, const std::string& m) { throw ParseFail{st, m + " at offset " + std::to_string(pos_)}; }
    bool eof() const { return pos_ >= p_.size(); }
    unsigned char peek() const { return (unsigned char)p_[pos_]; }
    unsigned char peek_at(size_t k) const { return pos_ + k < p_.size() ? (unsigned char)p_[pos_ + k] : 0; }

    int mk(Ast::Kind k) {
        pool_.emplace_back();
        pool_.back().kind = k;
        return (int)pool_.size() - 1;
    }
    int mk_set(ByteSet s, const Flags& f) {
        if (f.i) fold_case(s);
        int n = mk(Ast::SET);
        pool_[n].set = s;
        return n;
    }
    int mk_byte(unsigned c, const Flags& f) {
        ByteSet s;
        s.set(c);
        return mk_set(s, f);
    }
    int mk_assert(AssertKind a) {
        int n = mk(Ast::ASSERT);
        pool_[n].ak = a;
        if (assert_looks_at_words(a)) info_->uses_word_boundary = true;
        if (a == A_BOL_LINE || a == A_EOL_LINE) info_->uses_multiline = true;
        if (a == A_BOL_TEXT || a == A_BOL_LINE) info_->uses_bol = true;
        return n;
    }
    // A code point as a literal: ASCII -> one byte set; otherwise its UTF-8 bytes in sequence.
    int mk_codepoint(uint32_t cp, const Flags& f) {
        if (cp < 0x80) return mk_byte(cp, f);
        if (cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) fail(RX_INVALID, "invalid code point");
        unsigned char b[4];
        int n;
        if (cp < 0x800) { b[0] = 0xC0 | (cp >> 6); b[1] = 0x80 | (cp & 63); n = 2; }
        else if (cp < 0x10000) { b[0] = 0xE0 | (cp >> 12); b[1] = 0x80 | ((cp >> 6) & 63); b[2] = 0x80 | (cp & 63); n = 3; }
        else { b[0] = 0xF0 | (cp >> 18); b[1] = 0x80 | ((cp >> 12) & 63); b[2] = 0x80 | ((cp >> 6) & 63); b[3] = 0x80 | (cp & 63); n = 4; }
        int c = mk(Ast::CONCAT);
        Flags nf = f;
        nf.i = false;  // non-ASCII case folding is out of scope (ASCII haystacks)
        for (int k = 0; k < n; ++k) {
            int leaf = mk_byte(b[k], nf);
            pool_[c].kids.push_back(leaf);
        }
        return c;
    }

    void skip_ws(const Flags& f) {
        if (!f.x) return;
        while (!eof()) {
            unsigned char c = peek();
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0x0B || c == 0x0C) { ++pos_; continue; }
            if (c == '#') {
                while (!eof() && peek() != '\n') ++pos_;
                continue;
            }
            break;
        }
    }

    // alternation := concat ('|' concat)*   (stops at ')' or end)
    int parse_alt(Flags& f, int depth) {
        std::vector<int> branches;
        branches.push_back(parse_concat(f, depth));
        while (!eof() && peek() == '|') {
            ++pos_;
            branches.push_back(parse_concat(f, depth));
        }
        if (branches.size() == 1) return branches[0];
        int a = mk(Ast::ALT);
        pool_[a].kids = branches;
        return a;
    }

    int parse_concat(Flags& f, int depth) {
        std::vector<int> items;
        for (;;) {
            skip_ws(f);
            if (eof()) break;
            unsigned char c = peek();
            if (c == '|') break;
            if (c == ')') {
                if (depth == 0) fail(RX_INVALID, "unopened group");
                break;
            }
            if (c == '*' || c == '+' || c == '?' || c == '{') {
                if (items.empty()) fail(RX_INVALID, "repetition operator missing expression");
                int child = items.back();
                int rep = parse_repeat_op(child, f);
                items.back() = rep;
                continue;
            }
            int atom = parse_atom(f, depth);
            if (atom >= 0) items.push_back(atom);
        }
        if (items.empty()) return mk(Ast::EMPTY);
        if (items.size() == 1) return items[0];
        int cnode = mk(Ast::CONCAT);
        pool_[cnode].kids = items;
        return cnode;
    }

    bool parse_decimal(int* out) {
        size_t st = pos_;
        long v = 0;
        while (!eof() && peek() >= '0' && peek() <= '9') {
            v = v * 10 + (peek() - '0');
            if (v > 100000000) fail(RX_INVALID, "repetition count too large");
            ++pos_;
        }
        if (pos_ == st) return false;
        *out = (int)v;
        return true;
    }
    void skip_sp() {
        while (!eof() && peek() == ' ') ++pos_;
    }

    int parse_repeat_op(int child, const Flags& f) {
        int mn = 0, mx = -1;
        unsigned char c = peek();
        ++pos_;
        if (c == '*') { mn = 0; mx = -1; }
        else if (c == '+') { mn = 1; mx = -1; }
        else if (c == '?') { mn = 0; mx = 1; }
        else {  // '{'
            skip_sp();
            if (!parse_decimal(&mn)) fail(RX_INVALID, "repetition quantifier expects a valid decimal");
            skip_sp();
            if (eof()) fail(RX_INVALID, "unclosed counted repetition");
            if (peek() == ',') {
                ++pos_;
                skip_sp();
                if (eof()) fail(RX_INVALID, "unclosed counted repetition");
                if (peek() == '}') mx = -1;
                else if (!parse_decimal(&mx)) fail(RX_INVALID, "repetition quantifier expects a valid decimal");
                skip_sp();
            } else {
                mx = mn;
            }
            if (eof() || peek() != '}') fail(RX_INVALID, "unclosed counted repetition");
            ++pos_;
            if (mx >= 0 && mx < mn) fail(RX_INVALID, "invalid repetition count range");
        }
        // lazy suffix: irrelevant for match existence
        if (!eof() && peek() == '?') ++pos_;
        (void)f;
        int r = mk(Ast::REPEAT);
        pool_[r].kids.push_back(child);
        pool_[r].min = mn;
        pool_[r].max = mx;
        return r;
    }

    static int hexval(unsigned char c) {
        if (c >= '0' && c <= '9') return c - '0';
        if (c >= 'a' && c <= 'f') return c - 'a' + 10;
        if (c >= 'A' && c <= 'F') return c - 'A' + 10;
        return -1;
    }

    // after "\x" / "\u" / "\U": returns code point
    uint32_t parse_hex(unsigned char kind) {
        if (!eof() && peek() == '{') {
            ++pos_;
            uint32_t v = 0;
            int nd = 0;
            while (!eof() && peek() != '}') {
                int h = hexval(peek());
                if (h < 0) fail(RX_INVALID, "invalid hexadecimal digit");
                v = v * 16 + h;
                if (++nd > 8) fail(RX_INVALID, "hexadecimal literal too long");
                ++pos_;
            }
            if (eof()) fail(RX_INVALID, "unclosed hexadecimal literal");
            ++pos_;
            if (nd == 0) fail(RX_INVALID, "empty hexadecimal literal");
            return v;
        }
        int want = kind == 'x' ? 2 : kind == 'u' ? 4 : 8;
        uint32_t v = 0;
        for (int k = 0; k < want; ++k) {
            if (eof()) fail(RX_INVALID, "incomplete hexadecimal literal");
            int h = hexval(peek());
            if (h < 0) fail(RX_INVALID, "invalid hexadecimal digit");
            v = v * 16 + h;
            ++pos_;
        }
        return v;
    }

    // Escape outcomes
    struct Esc {
        enum { LITERAL, CLASS, ASSERTION } kind;
        uint32_t cp = 0;
        ByteSet set;
        AssertKind ak = A_WORD_B;
    };

    // pos_ is just past the backslash
    Esc parse_escape(bool in_class, const Flags& f) {
        if (eof()) fail(RX_INVALID, "incomplete escape sequence");
        unsigned char c = peek();
        ++pos_;
        Esc e;
        e.kind = Esc::LITERAL;
        switch (c) {
            case 'd': case 's': case 'w':
                e.kind = Esc::CLASS; e.set = perl_class((char)c); return e;
            case 'D': case 'S': case 'W':
                e.kind = Esc::CLASS; e.set = perl_class((char)(c + 32)); e.set.negate(); return e;
            case 'n': e.cp = '\n'; return e;
            case 't': e.cp = '\t'; return e;
            case 'r': e.cp = '\r'; return e;
            case 'a': e.cp = 0x07; return e;
            case 'f': e.cp = 0x0C; return e;
            case 'v': e.cp = 0x0B; return e;
            case 'x': case 'u': case 'U': e.cp = parse_hex(c); return e;
            case 'p': case 'P': {
                // \pL, \p{Letter}, \p{^L}, \p{gc=Lu}, \p{sc:Latin}, \P{..} (regex-syntax ast/parse.rs parse_unicode_class)
                if (eof()) fail(RX_INVALID, "incomplete Unicode class");
                std::string name;
                if (peek() == '{') {
                    const size_t close = p_.find('}', pos_);
                    if (close == std::string::npos) fail(RX_INVALID, "unclosed Unicode class");
                    name = p_.substr(pos_ + 1, close - pos_ - 1);
                    pos_ = close + 1;
                } else {
                    name = std::string(1, (char)peek());
                    ++pos_;
                }
                bool neg = false;
                const int rc = unicode_property_ascii(name, &e.set, &neg);
                if (rc == 2) fail(RX_INVALID, "malformed Unicode class");
                if (rc == 1) fail(RX_UNSUPPORTED, "Unicode property \\p{" + name + "} is not one this engine knows");
                if (f.i) fold_case(e.set);            // `(?i)\p{Lu}` matches `a`; folding comes BEFORE negation
                if (neg != (c == 'P')) e.set.negate();
                e.kind = Esc::CLASS;
                return e;
            }
            case 'A':
                if (in_class) fail(RX_INVALID, "unrecognized escape sequence in class");
                e.kind = Esc::ASSERTION; e.ak = A_BOL_TEXT; return e;
            case 'z':
                if (in_class) fail(RX_INVALID, "unrecognized escape sequence in class");
                e.kind = Esc::ASSERTION; e.ak = A_EOL_TEXT; return e;
            case 'b':
                if (in_class) fail(RX_INVALID, "unrecognized escape sequence in class");
                e.kind = Esc::ASSERTION; e.ak = A_WORD_B;
                if (!eof() && peek() == '{') {
                    // \b{start}, \b{end}, \b{start-half}, \b{end-half} (regex >= 1.10, regex-syntax ast/parse.rs
                    // maybe_parse_special_word_boundary): a name of letters and '-' between the braces; anything that does not
                    // start like a name is a counted repetition of a plain \b, which this engine leaves alone (loudly)
                    size_t q = pos_ + 1;
                    auto namech = [&](size_t i) { return i < p_.size() && (std::isalpha((unsigned char)p_[i]) || p_[i] == '-'); };
                    if (!namech(q)) fail(RX_UNSUPPORTED, "a counted repetition of \\b is not supported");
                    while (namech(q)) ++q;
                    if (q >= p_.size() || p_[q] != '}') fail(RX_INVALID, "special word boundary assertion is either unclosed or contains an invalid character");
                    const std::string name = p_.substr(pos_ + 1, q - pos_ - 1);
                    if (name == "start") e.ak = A_WORD_START;
                    else if (name == "end") e.ak = A_WORD_END;
                    else if (name == "start-half") e.ak = A_WORD_START_HALF;
                    else if (name == "end-half") e.ak = A_WORD_END_HALF;
                    else fail(RX_INVALID, "unrecognized special word boundary assertion");
                    pos_ = q + 1;
                }
                return e;
            case 'B':
                if (in_class) fail(RX_INVALID, "unrecognized escape sequence in class");
                e.kind = Esc::ASSERTION; e.ak = A_NOT_WORD_B; return e;
            case '<': case '>':
                if (in_class) fail(RX_INVALID, "unrecognized escape sequence in class");   // an assertion, as \b is (regex >= 1.10)
                e.kind = Esc::ASSERTION; e.ak = c == '<' ? A_WORD_START : A_WORD_END; return e;
            default: break;
        }
        if (c >= '0' && c <= '9') fail(RX_INVALID, "backreferences are not supported");
        // regex-syntax is_escapeable_character: every ASCII character that is not a letter or a digit (and not `<` `>`, taken
        // above) may be escaped and stands for itself -- punctuation, the blank, control characters
        if (c < 0x80 && !((c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) {
            e.cp = c;
            return e;
        }
        fail(RX_INVALID, "unrecognized escape sequence");
    }

    static bool posix_class(const std::string& name, ByteSet* out) {
        ByteSet s;
        if (name == "alnum") { s.set_range('0', '9'); s.set_range('A', 'Z'); s.set_range('a', 'z'); }
        else if (name == "alpha") { s.set_range('A', 'Z'); s.set_range('a', 'z'); }
        else if (name == "ascii") { s.set_range(0, 127); }
        else if (name == "blank") { s.set(' '); s.set('\t'); }
        else if (name == "cntrl") { s.set_range(0, 31); s.set(127); }
        else if (name == "digit") { s.set_range('0', '9'); }
        else if (name == "graph") { s.set_range('!', '~'); }
        else if (name == "lower") { s.set_range('a', 'z'); }
        else if (name == "print") { s.set_range(' ', '~'); }
        else if (name == "punct") { s.set_range('!', '/'); s.set_range(':', '@'); s.set_range('[', '`'); s.set_range('{', '~'); }
        else if (name == "space") { s.set('\t'); s.set('\n'); s.set(0x0B); s.set(0x0C); s.set('\r'); s.set(' '); }
        else if (name == "upper") { s.set_range('A', 'Z'); }
        else if (name == "word") { s.set_range('0', '9'); s.set_range('A', 'Z'); s.set_range('a', 'z'); s.set('_'); }
        else if (name == "xdigit") { s.set_range('0', '9'); s.set_range('A', 'F'); s.set_range('a', 'f'); }
        else return false;
        *out = s;
        return true;
    }

    void skip_class_ws(const Flags& f) {
        if (!f.x) return;
        while (!eof()) {
            unsigned char c = peek();
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0x0B || c == 0x0C) ++pos_;
            else break;
        }
    }

    // one class "atom" that can be a range endpoint: returns code point; or a whole set
    // returns true if *set_out filled (not usable as endpoint)
    bool parse_class_item(const Flags& f, uint32_t* cp, ByteSet* set_out) {
        unsigned char c = peek();
        if (c == '[') {
            // POSIX class or nested class
            if (peek_at(1) == ':') {
                size_t save = pos_;
                pos_ += 2;
                bool neg = false;
                if (!eof() && peek() == '^') { neg = true; ++pos_; }
                size_t st = pos_;
                while (!eof() && peek() != ':' && peek() != ']') ++pos_;
                if (!eof() && peek() == ':' && peek_at(1) == ']') {
                    std::string name = p_.substr(st, pos_ - st);
                    ByteSet s;
                    if (posix_class(name, &s)) {
                        pos_ += 2;
                        if (neg) s.negate();
                        *set_out = s;
                        return true;
                    }
                }
                pos_ = save;  // not a POSIX class: treat as nested class
            }
            ++pos_;
            *set_out = parse_class_body(f);
            return true;
        }
        if (c == '\\') {
            ++pos_;
            Esc e = parse_escape(true, f);
            if (e.kind == Esc::CLASS) { *set_out = e.set; return true; }
            *cp = e.cp;
            return false;
        }
        // plain literal: may be a multi-byte UTF-8 char in the pattern
        if (c < 0x80) { ++pos_; *cp = c; return false; }
        int n = (c >= 0xF0) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC0) ? 2 : 0;
        if (n == 0 || pos_ + n > p_.size()) fail(RX_INVALID, "invalid UTF-8 in pattern");
        uint32_t v = c & (0xFF >> (n + 1));
        for (int k = 1; k < n; ++k) {
            unsigned char cc = (unsigned char)p_[pos_ + k];
            if ((cc & 0xC0) != 0x80) fail(RX_INVALID, "invalid UTF-8 in pattern");
            v = (v << 6) | (cc & 63);
        }
        pos_ += n;
        *cp = v;
        return false;
    }

    // union of items until a set operator or ']' ; pos_ after '[' (and after '^' handled by caller)
    ByteSet parse_class_union(const Flags& f, bool first_item_allowed_bracket) {
        ByteSet acc;
        bool first = first_item_allowed_bracket;
        for (;;) {
            skip_class_ws(f);
            if (eof()) fail(RX_INVALID, "unclosed character class");
            unsigned char c = peek();
            if (c == ']' && !first) break;
            if ((c == '&' && peek_at(1) == '&') || (c == '-' && peek_at(1) == '-') || (c == '~' && peek_at(1) == '~')) break;
            uint32_t lo = 0;
            ByteSet sub;
            bool is_set;
            if (c == ']' && first) { ++pos_; lo = ']'; is_set = false; }
            else if (c == '-' ) { ++pos_; lo = '-'; is_set = false; }
            else is_set = parse_class_item(f, &lo, &sub);
            first = false;
            if (is_set) { acc.or_with(sub); continue; }
            // range?
            skip_class_ws(f);
            if (!eof() && peek() == '-' && peek_at(1) != ']' && peek_at(1) != 0 && !(peek_at(1) == '-')) {
                size_t save = pos_;
                ++pos_;
                skip_class_ws(f);
                if (eof()) fail(RX_INVALID, "unclosed character class");
                uint32_t hi = 0;
                ByteSet sub2;
                bool hs;
                if (peek() == '[') { hs = true; }
                else hs = parse_class_item(f, &hi, &sub2);
                if (hs) {
                    // "a-[" or "a-\d": Rust rejects a class as a range endpoint
                    (void)save;
                    fail(RX_INVALID, "invalid character class range");
                }
                if (hi < lo) fail(RX_INVALID, "invalid character class range");
                for (uint32_t v = lo; v <= hi && v < 0x80; ++v) acc.set(v);
                continue;
            }
            if (lo < 0x80) acc.set(lo);
            // non-ASCII single members never match an ASCII haystack: dropped
        }
        return acc;
    }

    // pos_ just after '['; consumes through the closing ']'
    ByteSet parse_class_body(const Flags& f) {
        bool neg = false;
        if (!eof() && peek() == '^') { neg = true; ++pos_; }
        ByteSet acc = parse_class_union(f, true);
        for (;;) {
            if (eof()) fail(RX_INVALID, "unclosed character class");
            unsigned char c = peek();
            if (c == ']') { ++pos_; break; }
            unsigned char op = c;  // && -- ~~
            pos_ += 2;
            ByteSet rhs = parse_class_union(f, false);
            if (op == '&') { for (int k = 0; k < 4; ++k) acc.w[k] &= rhs.w[k]; }
            else if (op == '-') { for (int k = 0; k < 4; ++k) acc.w[k] &= ~rhs.w[k]; }
            else { for (int k = 0; k < 4; ++k) acc.w[k] ^= rhs.w[k]; }
        }
        // positive sets are ASCII-only by construction (bytes >= 0x80 appear only through negation)
        ByteSet ascii;
        ascii.set_range(0, 127);
        for (int k = 0; k < 4; ++k) acc.w[k] &= ascii.w[k];
        if (f.i) fold_case(acc);
        if (neg) acc.negate();
        return acc;
    }

    void parse_flags(Flags& f, bool* scoped) {
        // pos_ after "(?" ; parses flags up to ')' or ':'
        bool negate = false, any = false;
        for (;;) {
            if (eof()) fail(RX_INVALID, "unclosed group");
            unsigned char c = peek();
            if (c == ')' || c == ':') {
                if (!any) fail(RX_INVALID, "missing flags");
                *scoped = (c == ':');
                ++pos_;
                return;
            }
            ++pos_;
            if (c == '-') {
                if (negate) fail(RX_INVALID, "repeated flag negation");
                negate = true;
                any = false;
                continue;
            }
            bool v = !negate;
            switch (c) {
                case 'i': f.i = v; break;
                case 'm': f.m = v; break;
                case 's': f.s = v; break;
                case 'U': f.U = v; break;
                case 'u': f.u = v; break;
                case 'x': f.x = v; break;
                case 'R': fail(RX_UNSUPPORTED, "CRLF mode (?R) is not supported");
                default: fail(RX_INVALID, "unrecognized flag");
            }
            any = true;
        }
    }

    // returns AST index, or -1 if the atom produced nothing (a bare flags group)
    int parse_atom(Flags& f, int depth) {
        unsigned char c = peek();
        switch (c) {
            case '(': {
                ++pos_;
                Flags inner = f;
                if (!eof() && peek() == '?') {
                    ++pos_;
                    if (eof()) fail(RX_INVALID, "unclosed group");
                    unsigned char d = peek();
                    if (d == 'P' || d == '<') {
                        // named group (?P<name>..) / (?<name>..); look-behind (?<= (?<! is rejected
                        if (d == 'P') {
                            ++pos_;
                            if (eof() || peek() != '<') fail(RX_INVALID, "invalid named group");
                        } else if (peek_at(1) == '=' || peek_at(1) == '!') {
                            fail(RX_INVALID, "look-around is not supported");
                        }
                        ++pos_;
                        size_t st = pos_;
                        while (!eof() && peek() != '>') ++pos_;
                        if (eof() || pos_ == st) fail(RX_INVALID, "invalid capture group name");
                        ++pos_;
                    } else if (d == '=' || d == '!') {
                        fail(RX_INVALID, "look-around is not supported");
                    } else if (d == ':') {
                        ++pos_;  // (?:...) non-capturing group
                    } else {
                        bool scoped = false;
                        parse_flags(inner, &scoped);
                        if (!scoped) {
                            f = inner;  // (?flags) applies to the rest of the enclosing group
                            return -1;
                        }
                    }
                }
                if (depth > 200) fail(RX_INVALID, "nesting too deep");
                int r = parse_alt(inner, depth + 1);
                if (eof() || peek() != ')') fail(RX_INVALID, "unclosed group");
                ++pos_;
                return r;
            }
            case '[': {
                ++pos_;
                ByteSet s = parse_class_body(f);
                int n = mk(Ast::SET);
                pool_[n].set = s;  // folding/negation already applied
                return n;
            }
            case '.': {
                ++pos_;
                ByteSet s;
                s.negate();  // everything
                if (!f.s) { ByteSet nl; nl.set('\n'); for (int k = 0; k < 4; ++k) s.w[k] &= ~nl.w[k]; }
                int n = mk(Ast::SET);
                pool_[n].set = s;
                return n;
            }
            case '^': ++pos_; return mk_assert(f.m ? A_BOL_LINE : A_BOL_TEXT);
            case '$': ++pos_; return mk_assert(f.m ? A_EOL_LINE : A_EOL_TEXT);
            case '\\': {
                ++pos_;
                Esc e = parse_escape(false, f);
                if (e.kind == Esc::ASSERTION) return mk_assert(e.ak);
                if (e.kind == Esc::CLASS) {
                    // Perl classes are not affected by (?i) (already case-closed; \p classes were folded by parse_escape); keep
                    // bytes >= 0x80 of negations
                    int n = mk(Ast::SET);
                    pool_[n].set = e.set;
                    return n;
                }
                return mk_codepoint(e.cp, f);
            }
            default: break;
        }
        if (c < 0x80) { ++pos_; return mk_byte(c, f); }
        // literal non-ASCII char in the pattern
        uint32_t cp = 0;
        ByteSet dummy;
        parse_class_item(f, &cp, &dummy);
        return mk_codepoint(cp, f);
    }
};

static bool nullable_no_assert(const std::vector<Ast>& pool, int n) {
    const Ast& a = pool[n];
    switch (a.kind) {
        case Ast::EMPTY: return true;
        case Ast::SET: return false;
        case Ast::ASSERT: return false;
        case Ast::CONCAT:
            for (int k : a.kids) if (!nullable_no_assert(pool, k)) return false;
            return true;
        case Ast::ALT:
            for (int k : a.kids) if (nullable_no_assert(pool, k)) return true;
            return false;
        case Ast::REPEAT: return a.min == 0 || nullable_no_assert(pool, a.kids[0]);
    }
    return false;
}

struct Emitter {
    Nfa& nfa;
    const std::vector<Ast>& pool;
    size_t base;
    int node(NfaKind k) {
        if (nfa.nodes.size() - base > (size_t)kMaxNfaNodesPerPattern) throw ParseFail{RX_TOO_BIG, "compiled regex exceeds size limit"};
        NfaNode n;
        n.kind = k;
        nfa.nodes.push_back(n);
        return (int)nfa.nodes.size() - 1;
    }
    // emit `a` so that it continues to `next`; returns entry node
    int emit(int a, int next) {
        const Ast& A = pool[a];
        switch (A.kind) {
            case Ast::EMPTY: return next;
            case Ast::SET: {
                int n = node(N_CHAR);
                nfa.nodes[n].set = nfa.add_set(A.set);
                nfa.nodes[n].out = next;
                return n;
            }
            case Ast::ASSERT: {
                int n = node(N_ASSERT);
                nfa.nodes[n].assert_kind = A.ak;
                nfa.nodes[n].out = next;
                return n;
            }
            case Ast::CONCAT: {
                int cur = next;
                for (size_t k = A.kids.size(); k-- > 0;) cur = emit(A.kids[k], cur);
                return cur;
            }
            case Ast::ALT: {
                // chain of splits
                int entry = -1;
                int prev_split = -1;
                for (size_t k = 0; k < A.kids.size(); ++k) {
                    int br = emit(A.kids[k], next);
                    if (k + 1 == A.kids.size()) {
                        if (prev_split >= 0) nfa.nodes[prev_split].out1 = br;
                        else entry = br;
                    } else {
                        int s = node(N_SPLIT);
                        nfa.nodes[s].out = br;
                        if (prev_split >= 0) nfa.nodes[prev_split].out1 = s;
                        else entry = s;
                        prev_split = s;
                    }
                }
                return entry;
            }
            case Ast::REPEAT: {
                int child = A.kids[0];
                int tail = next;
                if (A.max < 0) {
                    // x* loop
                    int s = node(N_SPLIT);
                    int body = emit(child, s);
                    nfa.nodes[s].out = body;
                    nfa.nodes[s].out1 = next;
                    tail = s;
                } else {
                    for (int k = 0; k < A.max - A.min; ++k) {
                        int s = node(N_SPLIT);
                        int body = emit(child, tail);
                        nfa.nodes[s].out = body;
                        nfa.nodes[s].out1 = next;
                        tail = s;
                    }
                }
                int cur = tail;
                for (int k = 0; k < A.min; ++k) cur = emit(child, cur);
                return cur;
            }
        }
        return next;
    }
};

}  // namespace

static int popcount_set(const ByteSet& s) {
    return __builtin_popcountll(s.w[0]) + __builtin_popcountll(s.w[1]) + __builtin_popcountll(s.w[2]) + __builtin_popcountll(s.w[3]);
}

RegexStatus regex_compile(const std::string& pattern, int first_pattern_id, Nfa& nfa, RegexParts* parts, RegexInfo* info,
                          std::string& err, bool allow_split) {
    std::vector<Ast> pool;
    RegexInfo local;
    size_t nodes0 = nfa.nodes.size(), sets0 = nfa.sets.size();
    try {
        Parser p(pattern, pool, &local);
        int root = p.parse();
        local.always_true = nullable_no_assert(pool, root);
        Emitter em{nfa, pool, nodes0};
        *parts = RegexParts();

        // ---- gap split:  X G* S  ------------------------------------------------------------
        bool split = false;
        if (allow_split && !local.always_true && pool[root].kind == Ast::CONCAT && pool[root].kids.size() >= 3) {
            const std::vector<int>& it = pool[root].kids;
            const Ast& last = pool[it[it.size() - 1]];
            const Ast& rep = pool[it[it.size() - 2]];
            if (last.kind == Ast::SET && !last.set.empty() && rep.kind == Ast::REPEAT && rep.max < 0 && pool[rep.kids[0]].kind == Ast::SET &&
                rep.min <= 4) {
                ByteSet G = pool[rep.kids[0]].set;
                ByteSet notG = G;
                notG.negate();
                if (popcount_set(notG) <= 4) {
                    // prefix = items[0 .. n-2) followed by `min` mandatory copies of G
                    int pre = (int)pool.size();
                    pool.emplace_back();
                    pool[pre].kind = Ast::CONCAT;
                    for (size_t k = 0; k + 2 < pool[root].kids.size(); ++k) pool[pre].kids.push_back(pool[root].kids[k]);
                    int gchild = pool[pool[root].kids[pool[root].kids.size() - 2]].kids[0];
                    int gmin = pool[pool[root].kids[pool[root].kids.size() - 2]].min;
                    for (int k = 0; k < gmin; ++k) pool[pre].kids.push_back(gchild);
                    if (!nullable_no_assert(pool, pre)) {
                        int m0 = em.node(N_MATCH);
                        nfa.nodes[m0].pattern = first_pattern_id;
                        parts->start[0] = em.emit(pre, m0);
                        parts->kind[0] = EV_SET;
                        int m1 = em.node(N_MATCH);
                        nfa.nodes[m1].pattern = first_pattern_id + 1;
                        int c1 = em.node(N_CHAR);
                        nfa.nodes[c1].set = nfa.add_set(pool[pool[root].kids.back()].set);
                        nfa.nodes[c1].out = m1;
                        parts->start[1] = c1;
                        parts->kind[1] = EV_TEST;
                        parts->n = 2;
                        if (!notG.empty()) {
                            int m2 = em.node(N_MATCH);
                            nfa.nodes[m2].pattern = first_pattern_id + 2;
                            int c2 = em.node(N_CHAR);
                            nfa.nodes[c2].set = nfa.add_set(notG);
                            nfa.nodes[c2].out = m2;
                            parts->start[2] = c2;
                            parts->kind[2] = EV_CLEAR;
                            parts->n = 3;
                        }
                        split = true;
                    }
                }
            }
        }
        if (!split) {
            int m = em.node(N_MATCH);
            nfa.nodes[m].pattern = first_pattern_id;
            parts->start[0] = em.emit(root, m);
            parts->kind[0] = EV_FIRE;
            parts->n = 1;
        }
    } catch (const ParseFail& f) {
        nfa.nodes.resize(nodes0);
        nfa.sets.resize(sets0);
        err = f.msg;
        return f.st;
    }
    if (info) *info = local;
    return RX_OK;
}

int nfa_literal(Nfa& nfa, const std::string& lit, bool anchor_start, bool anchor_end, int pattern_id) {
    auto add = [&](NfaKind k) {
        NfaNode n;
        n.kind = k;
        nfa.nodes.push_back(n);
        return (int)nfa.nodes.size() - 1;
    };
    int m = add(N_MATCH);
    nfa.nodes[m].pattern = pattern_id;
    int cur = m;
    if (anchor_end) {
        int a = add(N_ASSERT);
        nfa.nodes[a].assert_kind = A_EOL_TEXT;
        nfa.nodes[a].out = cur;
        cur = a;
    }
    for (size_t k = lit.size(); k-- > 0;) {
        ByteSet s;
        s.set((unsigned char)lit[k]);
        int n = add(N_CHAR);
        nfa.nodes[n].set = nfa.add_set(s);
        nfa.nodes[n].out = cur;
        cur = n;
    }
    if (anchor_start) {
        int a = add(N_ASSERT);
        nfa.nodes[a].assert_kind = A_BOL_TEXT;
        nfa.nodes[a].out = cur;
        cur = a;
    }
    return cur;
}

}  // namespace pgw
