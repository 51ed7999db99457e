// Recursive-descent parser for the rule expression language (see expr.hpp).
#include "expr.hpp"

#include <cerrno>
#include <cstdlib>
#include <cstring>

namespace pgw {
namespace {

struct SyntaxError {
    std::string msg;
};

enum Tok : uint8_t {
    T_END, T_IDENT, T_INT, T_UINT, T_FLOAT, T_STR, T_BYTES,
    T_LPAREN, T_RPAREN, T_LBRACK, T_RBRACK, T_LBRACE, T_RBRACE,
    T_DOT, T_COMMA, T_COLON, T_QUESTION,
    T_NOT, T_MINUS, T_PLUS, T_STAR, T_SLASH, T_PERCENT,
    T_OR, T_AND, T_EQ, T_NE, T_LT, T_LE, T_GT, T_GE
};

struct Token {
    Tok t = T_END;
    std::string text;
    int64_t ival = 0;
    double fval = 0;
    size_t pos = 0;
};

class Lexer {
  public:
    explicit Lexer(const std::string& s) : s_(s) {}
    Token next() {
        skip();
        Token k;
        k.pos = p_;
        if (p_ >= s_.size()) return k;
        unsigned char c = s_[p_];
        if (isalpha(c) || c == '_') {
            // string prefixes r"..", b"..", rb"..", br".."
            size_t q = p_;
            bool raw = false, bytes = false;
            while (q < s_.size() && q - p_ < 2 && strchr("rRbB", s_[q])) {
                if (s_[q] == 'r' || s_[q] == 'R') { if (raw) break; raw = true; }
                else { if (bytes) break; bytes = true; }
                ++q;
            }
            if (q > p_ && q < s_.size() && (s_[q] == '"' || s_[q] == '\'')) {
                p_ = q;
                k.text = lex_string(raw, bytes);
                k.t = bytes ? T_BYTES : T_STR;
                return k;
            }
            size_t st = p_;
            while (p_ < s_.size() && (isalnum((unsigned char)s_[p_]) || s_[p_] == '_')) ++p_;
            k.t = T_IDENT;
            k.text = s_.substr(st, p_ - st);
            return k;
        }
        if (isdigit(c) || (c == '.' && p_ + 1 < s_.size() && isdigit((unsigned char)s_[p_ + 1]))) return lex_number();
        if (c == '"' || c == '\'') {
            k.text = lex_string(false, false);
            k.t = T_STR;
            return k;
        }
        ++p_;
        auto two = [&](char n) { if (p_ < s_.size() && s_[p_] == n) { ++p_; return true; } return false; };
        switch (c) {
            case '(': k.t = T_LPAREN; return k;
            case ')': k.t = T_RPAREN; return k;
            case '[': k.t = T_LBRACK; return k;
            case ']': k.t = T_RBRACK; return k;
            case '{': k.t = T_LBRACE; return k;
            case '}': k.t = T_RBRACE; return k;
            case '.': k.t = T_DOT; return k;
            case ',': k.t = T_COMMA; return k;
            case ':': k.t = T_COLON; return k;
            case '?': k.t = T_QUESTION; return k;
            case '+': k.t = T_PLUS; return k;
            case '-': k.t = T_MINUS; return k;
            case '*': k.t = T_STAR; return k;
            case '/': k.t = T_SLASH; return k;
            case '%': k.t = T_PERCENT; return k;
            case '!': k.t = two('=') ? T_NE : T_NOT; return k;
            case '=': if (two('=')) { k.t = T_EQ; return k; } break;
            case '<': k.t = two('=') ? T_LE : T_LT; return k;
            case '>': k.t = two('=') ? T_GE : T_GT; return k;
            case '|': if (two('|')) { k.t = T_OR; return k; } break;
            case '&': if (two('&')) { k.t = T_AND; return k; } break;
            default: break;
        }
        throw SyntaxError{"unexpected character '" + std::string(1, (char)c) + "' at offset " + std::to_string(k.pos)};
    }

  private:
    const std::string& s_;
    size_t p_ = 0;

    void skip() {
        for (;;) {
            while (p_ < s_.size() && isspace((unsigned char)s_[p_])) ++p_;
            if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '/') {
                while (p_ < s_.size() && s_[p_] != '\n') ++p_;
                continue;
            }
            break;
        }
    }

    Token lex_number() {
        Token k;
        k.pos = p_;
        size_t st = p_;
        if (s_[p_] == '0' && p_ + 1 < s_.size() && (s_[p_ + 1] == 'x' || s_[p_ + 1] == 'X')) {
            p_ += 2;
            size_t hs = p_;
            while (p_ < s_.size() && isxdigit((unsigned char)s_[p_])) ++p_;
            if (p_ == hs) throw SyntaxError{"invalid hex literal at offset " + std::to_string(st)};
            errno = 0;
            unsigned long long v = strtoull(s_.substr(hs, p_ - hs).c_str(), nullptr, 16);
            if (errno) throw SyntaxError{"integer literal out of range at offset " + std::to_string(st)};
            return finish_int(k, v, st);
        }
        while (p_ < s_.size() && isdigit((unsigned char)s_[p_])) ++p_;
        bool is_float = false;
        if (p_ < s_.size() && s_[p_] == '.' && p_ + 1 < s_.size() && isdigit((unsigned char)s_[p_ + 1])) {
            is_float = true;
            ++p_;
            while (p_ < s_.size() && isdigit((unsigned char)s_[p_])) ++p_;
        }
        if (p_ < s_.size() && (s_[p_] == 'e' || s_[p_] == 'E')) {
            size_t save = p_;
            ++p_;
            if (p_ < s_.size() && (s_[p_] == '+' || s_[p_] == '-')) ++p_;
            if (p_ < s_.size() && isdigit((unsigned char)s_[p_])) {
                is_float = true;
                while (p_ < s_.size() && isdigit((unsigned char)s_[p_])) ++p_;
            } else {
                p_ = save;
            }
        }
        if (is_float) {
            k.t = T_FLOAT;
            k.fval = strtod(s_.substr(st, p_ - st).c_str(), nullptr);
            return k;
        }
        errno = 0;
        unsigned long long v = strtoull(s_.substr(st, p_ - st).c_str(), nullptr, 10);
        if (errno) throw SyntaxError{"integer literal out of range at offset " + std::to_string(st)};
        return finish_int(k, v, st);
    }

    Token finish_int(Token k, unsigned long long v, size_t st) {
        if (p_ < s_.size() && (s_[p_] == 'u' || s_[p_] == 'U')) {
            ++p_;
            k.t = T_UINT;
            k.ival = (int64_t)v;
            return k;
        }
        // 9223372036854775808 is only valid under unary minus; the parser handles that case
        if (v > 9223372036854775808ull) throw SyntaxError{"integer literal out of range at offset " + std::to_string(st)};
        k.t = T_INT;
        k.ival = (int64_t)v;  // 2^63 wraps to INT64_MIN; parser rejects it unless negated
        k.text = v == 9223372036854775808ull ? "min" : "";
        return k;
    }

    static void put_utf8(std::string& out, uint32_t cp) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 63))); }
        else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 63))); out.push_back((char)(0x80 | (cp & 63))); }
        else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 63))); out.push_back((char)(0x80 | ((cp >> 6) & 63))); out.push_back((char)(0x80 | (cp & 63))); }
    }

    std::string lex_string(bool raw, bool bytes) {
        size_t st = p_;
        char q = s_[p_];
        bool triple = p_ + 2 < s_.size() && s_[p_ + 1] == q && s_[p_ + 2] == q;
        p_ += triple ? 3 : 1;
        std::string out;
        for (;;) {
            if (p_ >= s_.size()) throw SyntaxError{"unterminated string literal at offset " + std::to_string(st)};
            char c = s_[p_];
            if (c == q) {
                if (!triple) { ++p_; break; }
                if (p_ + 2 < s_.size() + 0 && s_.compare(p_, 3, std::string(3, q)) == 0) { p_ += 3; break; }
                out.push_back(c);
                ++p_;
                continue;
            }
            if (!triple && (c == '\n' || c == '\r')) throw SyntaxError{"newline in string literal at offset " + std::to_string(p_)};
            if (c != '\\' || raw) { out.push_back(c); ++p_; continue; }
            ++p_;
            if (p_ >= s_.size()) throw SyntaxError{"unterminated escape at offset " + std::to_string(p_)};
            char e = s_[p_++];
            switch (e) {
                case 'a': out.push_back('\a'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'v': out.push_back('\v'); break;
                case '\\': case '?': case '"': case '\'': case '`': out.push_back(e); break;
                case 'x': case 'X': case 'u': case 'U': {
                    int n = (e == 'x' || e == 'X') ? 2 : (e == 'u' ? 4 : 8);
                    if (bytes && n != 2) throw SyntaxError{"\\u escapes are not allowed in bytes literals"};
                    uint32_t v = 0;
                    for (int k = 0; k < n; ++k) {
                        if (p_ >= s_.size() || !isxdigit((unsigned char)s_[p_])) throw SyntaxError{"invalid hex escape at offset " + std::to_string(p_)};
                        char h = s_[p_++];
                        v = v * 16 + (isdigit((unsigned char)h) ? h - '0' : (tolower(h) - 'a' + 10));
                    }
                    if (bytes) out.push_back((char)v);
                    else {
                        if (v > 0x10FFFF || (v >= 0xD800 && v <= 0xDFFF)) throw SyntaxError{"invalid code point in string literal"};
                        put_utf8(out, v);
                    }
                    break;
                }
                default:
                    if (e >= '0' && e <= '3') {
                        uint32_t v = e - '0';
                        for (int k = 0; k < 2; ++k) {
                            if (p_ >= s_.size() || s_[p_] < '0' || s_[p_] > '7') throw SyntaxError{"invalid octal escape at offset " + std::to_string(p_)};
                            v = v * 8 + (s_[p_++] - '0');
                        }
                        if (bytes) out.push_back((char)v);
                        else put_utf8(out, v);
                        break;
                    }
                    throw SyntaxError{"invalid escape sequence '\\" + std::string(1, e) + "' at offset " + std::to_string(p_ - 2)};
            }
        }
        return out;
    }
};

class Parser {
  public:
    explicit Parser(const std::string& s) : lex_(s) { advance(); }

    ExprP parse_all() {
        ExprP e = parse_expr(0);
        if (cur_.t != T_END) throw SyntaxError{"unexpected token at offset " + std::to_string(cur_.pos)};
        return e;
    }

  private:
    Lexer lex_;
    Token cur_;

    void advance() { cur_ = lex_.next(); }
    bool accept(Tok t) {
        if (cur_.t == t) { advance(); return true; }
        return false;
    }
    void expect(Tok t, const char* what) {
        if (!accept(t)) throw SyntaxError{std::string("expected ") + what + " at offset " + std::to_string(cur_.pos)};
    }
    static ExprP mk(Expr::Kind k, size_t pos) {
        ExprP e(new Expr());
        e->kind = k;
        e->pos = pos;
        return e;
    }
    static ExprP bin(Expr::Op op, ExprP a, ExprP b, size_t pos) {
        ExprP e = mk(Expr::BINARY, pos);
        e->op = op;
        e->kids.push_back(std::move(a));
        e->kids.push_back(std::move(b));
        return e;
    }

    ExprP parse_expr(int depth) {
        if (depth > 200) throw SyntaxError{"expression nesting too deep"};
        ExprP c = parse_or(depth);
        if (cur_.t == T_QUESTION) {
            size_t pos = cur_.pos;
            advance();
            ExprP a = parse_or(depth + 1);
            expect(T_COLON, "':'");
            ExprP b = parse_expr(depth + 1);
            ExprP t = mk(Expr::TERNARY, pos);
            t->kids.push_back(std::move(c));
            t->kids.push_back(std::move(a));
            t->kids.push_back(std::move(b));
            return t;
        }
        return c;
    }
    ExprP parse_or(int depth) {
        ExprP l = parse_and(depth);
        while (cur_.t == T_OR) {
            size_t pos = cur_.pos;
            advance();
            l = bin(Expr::OP_OR, std::move(l), parse_and(depth), pos);
        }
        return l;
    }
    ExprP parse_and(int depth) {
        ExprP l = parse_rel(depth);
        while (cur_.t == T_AND) {
            size_t pos = cur_.pos;
            advance();
            l = bin(Expr::OP_AND, std::move(l), parse_rel(depth), pos);
        }
        return l;
    }
    ExprP parse_rel(int depth) {
        ExprP l = parse_add(depth);
        for (;;) {
            Expr::Op op = Expr::OP_NONE;
            switch (cur_.t) {
                case T_EQ: op = Expr::OP_EQ; break;
                case T_NE: op = Expr::OP_NE; break;
                case T_LT: op = Expr::OP_LT; break;
                case T_LE: op = Expr::OP_LE; break;
                case T_GT: op = Expr::OP_GT; break;
                case T_GE: op = Expr::OP_GE; break;
                case T_IDENT: if (cur_.text == "in") op = Expr::OP_IN; break;
                default: break;
            }
            if (op == Expr::OP_NONE) return l;
            size_t pos = cur_.pos;
            advance();
            l = bin(op, std::move(l), parse_add(depth), pos);
        }
    }
    ExprP parse_add(int depth) {
        ExprP l = parse_mul(depth);
        while (cur_.t == T_PLUS || cur_.t == T_MINUS) {
            Expr::Op op = cur_.t == T_PLUS ? Expr::OP_ADD : Expr::OP_SUB;
            size_t pos = cur_.pos;
            advance();
            l = bin(op, std::move(l), parse_mul(depth), pos);
        }
        return l;
    }
    ExprP parse_mul(int depth) {
        ExprP l = parse_unary(depth);
        while (cur_.t == T_STAR || cur_.t == T_SLASH || cur_.t == T_PERCENT) {
            Expr::Op op = cur_.t == T_STAR ? Expr::OP_MUL : cur_.t == T_SLASH ? Expr::OP_DIV : Expr::OP_MOD;
            size_t pos = cur_.pos;
            advance();
            l = bin(op, std::move(l), parse_unary(depth), pos);
        }
        return l;
    }
    ExprP parse_unary(int depth) {
        if (depth > 200) throw SyntaxError{"expression nesting too deep"};
        if (cur_.t == T_NOT) {
            size_t pos = cur_.pos;
            advance();
            ExprP e = mk(Expr::UNARY, pos);
            e->op = Expr::OP_NOT;
            e->kids.push_back(parse_unary(depth + 1));
            return e;
        }
        if (cur_.t == T_MINUS) {
            size_t pos = cur_.pos;
            advance();
            if (cur_.t == T_INT && cur_.text == "min") {
                // -9223372036854775808
                ExprP e = mk(Expr::LIT_INT, pos);
                e->ival = INT64_MIN;
                advance();
                return parse_postfix(std::move(e), depth);
            }
            ExprP e = mk(Expr::UNARY, pos);
            e->op = Expr::OP_NEG;
            e->kids.push_back(parse_unary(depth + 1));
            return e;
        }
        return parse_postfix(parse_primary(depth), depth);
    }

    void parse_args(std::vector<ExprP>& out, int depth) {
        // after '('
        if (accept(T_RPAREN)) return;
        for (;;) {
            out.push_back(parse_expr(depth + 1));
            if (accept(T_COMMA)) {
                if (cur_.t == T_RPAREN) { advance(); return; }  // trailing comma
                continue;
            }
            expect(T_RPAREN, "')'");
            return;
        }
    }

    ExprP parse_postfix(ExprP e, int depth) {
        for (;;) {
            if (cur_.t == T_DOT) {
                size_t pos = cur_.pos;
                advance();
                if (cur_.t != T_IDENT) throw SyntaxError{"expected identifier after '.' at offset " + std::to_string(cur_.pos)};
                std::string name = cur_.text;
                advance();
                if (cur_.t == T_LPAREN) {
                    advance();
                    ExprP m = mk(Expr::METHOD, pos);
                    m->name = name;
                    m->kids.push_back(std::move(e));
                    parse_args(m->kids, depth);
                    e = std::move(m);
                } else {
                    ExprP m = mk(Expr::MEMBER, pos);
                    m->name = name;
                    m->kids.push_back(std::move(e));
                    e = std::move(m);
                }
                continue;
            }
            if (cur_.t == T_LBRACK) {
                size_t pos = cur_.pos;
                advance();
                ExprP ix = mk(Expr::INDEX, pos);
                ix->kids.push_back(std::move(e));
                ix->kids.push_back(parse_expr(depth + 1));
                expect(T_RBRACK, "']'");
                e = std::move(ix);
                continue;
            }
            return e;
        }
    }

    ExprP parse_primary(int depth) {
        Token k = cur_;
        switch (k.t) {
            case T_INT: {
                if (k.text == "min") throw SyntaxError{"integer literal out of range at offset " + std::to_string(k.pos)};
                advance();
                ExprP e = mk(Expr::LIT_INT, k.pos);
                e->ival = k.ival;
                return e;
            }
            case T_UINT: {
                advance();
                ExprP e = mk(Expr::LIT_UINT, k.pos);
                e->ival = k.ival;
                return e;
            }
            case T_FLOAT: {
                advance();
                ExprP e = mk(Expr::LIT_FLOAT, k.pos);
                e->fval = k.fval;
                return e;
            }
            case T_STR: case T_BYTES: {
                advance();
                ExprP e = mk(k.t == T_STR ? Expr::LIT_STR : Expr::LIT_BYTES, k.pos);
                e->name = k.text;
                return e;
            }
            case T_IDENT: {
                advance();
                if (k.text == "true" || k.text == "false") {
                    ExprP e = mk(Expr::LIT_BOOL, k.pos);
                    e->bval = k.text == "true";
                    return e;
                }
                if (k.text == "null") return mk(Expr::LIT_NULL, k.pos);
                if (k.text == "in") throw SyntaxError{"unexpected 'in' at offset " + std::to_string(k.pos)};
                if (cur_.t == T_LPAREN) {
                    advance();
                    ExprP c = mk(Expr::CALL, k.pos);
                    c->name = k.text;
                    parse_args(c->kids, depth);
                    return c;
                }
                ExprP e = mk(Expr::IDENT, k.pos);
                e->name = k.text;
                return e;
            }
            case T_LPAREN: {
                advance();
                ExprP e = parse_expr(depth + 1);
                expect(T_RPAREN, "')'");
                return e;
            }
            case T_LBRACK: {
                advance();
                ExprP l = mk(Expr::LIST, k.pos);
                if (accept(T_RBRACK)) return l;
                for (;;) {
                    l->kids.push_back(parse_expr(depth + 1));
                    if (accept(T_COMMA)) {
                        if (accept(T_RBRACK)) return l;
                        continue;
                    }
                    expect(T_RBRACK, "']'");
                    return l;
                }
            }
            case T_LBRACE: {
                advance();
                ExprP m = mk(Expr::MAP, k.pos);
                if (accept(T_RBRACE)) return m;
                for (;;) {
                    m->kids.push_back(parse_expr(depth + 1));
                    expect(T_COLON, "':'");
                    m->kids.push_back(parse_expr(depth + 1));
                    if (accept(T_COMMA)) {
                        if (accept(T_RBRACE)) return m;
                        continue;
                    }
                    expect(T_RBRACE, "'}'");
                    return m;
                }
            }
            case T_END: throw SyntaxError{"unexpected end of expression"};
            default: throw SyntaxError{"unexpected token at offset " + std::to_string(k.pos)};
        }
    }
};

}  // namespace

ExprP parse_expression(const std::string& src, std::string& err) {
    try {
        Parser p(src);
        return p.parse_all();
    } catch (const SyntaxError& e) {
        err = e.msg;
        return nullptr;
    }
}

void collect_functions(const Expr& e, std::vector<std::string>& out) {
    if (e.kind == Expr::CALL || e.kind == Expr::METHOD) out.push_back(e.name);
    if (e.kind == Expr::BINARY && e.op == Expr::OP_IN) out.push_back("@in");
    for (const auto& k : e.kids) collect_functions(*k, out);
}

}  // namespace pgw
