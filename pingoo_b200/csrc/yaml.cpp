// YAML subset reader (see yaml.hpp).
#include "yaml.hpp"

#include <cstdio>

namespace pgw {
namespace {

struct Line {
    int indent = 0;
    std::string text;  // content after the indentation, comment stripped, right-trimmed
    std::string raw;   // the whole line as written (block scalars keep comments and trailing blanks out of `text`'s way)
    int no = 0;
    bool blank = false;
};

struct Fail {
    std::string msg;
};

[[noreturn]] void fail(int line, const std::string& m) { throw Fail{"line " + std::to_string(line) + ": " + m}; }

std::string rtrim(std::string s) {
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back();
    return s;
}
std::string ltrim(const std::string& s) {
    size_t i = 0;
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i;
    return s.substr(i);
}

// strip a trailing comment: '#' at the start or after white space, outside quotes
std::string strip_comment(const std::string& s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i];
        if (q) {
            if (q == '"' && c == '\\') { ++i; continue; }
            if (c == q) {
                if (q == '\'' && i + 1 < s.size() && s[i + 1] == '\'') { ++i; continue; }
                q = 0;
            }
        } else if ((c == '"' || c == '\'') && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t' || s[i - 1] == '[' || s[i - 1] == '{' || s[i - 1] == ',' || s[i - 1] == ':' )) {
            q = c;
        } else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) {
            return s.substr(0, i);
        }
    }
    return s;
}

void append_utf8(std::string& o, unsigned cp) {
    if (cp < 0x80) o += (char)cp;
    else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 63)); o += (char)(0x80 | (cp & 63)); }
    else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 63)); o += (char)(0x80 | ((cp >> 6) & 63)); o += (char)(0x80 | (cp & 63)); }
}

class Parser {
  public:
    explicit Parser(const std::string& text) {
        size_t pos = 0;
        int no = 0;
        while (pos <= text.size()) {
            size_t nl = text.find('\n', pos);
            std::string raw = text.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
            pos = nl == std::string::npos ? text.size() + 1 : nl + 1;
            ++no;
            if (!raw.empty() && raw.back() == '\r') raw.pop_back();
            Line l;
            l.no = no;
            l.raw = raw;
            size_t i = 0;
            while (i < raw.size() && raw[i] == ' ') ++i;
            if (i < raw.size() && raw[i] == '\t') {
                // tabs are not indentation in YAML; tolerate a tab-only blank line
                if (rtrim(raw).empty()) { l.blank = true; lines_.push_back(l); continue; }
            }
            l.indent = (int)i;
            l.text = rtrim(strip_comment(raw.substr(i)));
            l.blank = l.text.empty();
            lines_.push_back(l);
        }
    }

    YNode parse() {
        size_t i = next(0);
        // directives / document start
        while (i < lines_.size() && (lines_[i].text == "---" || (lines_[i].indent == 0 && !lines_[i].text.empty() && lines_[i].text[0] == '%'))) i = next(i + 1);
        if (i < lines_.size() && lines_[i].text.compare(0, 4, "--- ") == 0) {
            lines_[i].text = ltrim(lines_[i].text.substr(4));
            lines_[i].indent = 0;
        }
        YNode root;
        if (i >= lines_.size()) return root;
        root = block(i, lines_[i].indent);
        i = next(i);
        if (i < lines_.size()) {
            if (lines_[i].text == "..." || lines_[i].text == "---") {
                size_t j = next(i + 1);
                if (j < lines_.size()) fail(lines_[j].no, "multiple documents are not supported");
            } else fail(lines_[i].no, "unexpected content (bad indentation?)");
        }
        return root;
    }

  private:
    std::vector<Line> lines_;

    size_t next(size_t i) const {
        while (i < lines_.size() && lines_[i].blank) ++i;
        return i;
    }

    static bool is_seq_item(const std::string& t) { return !t.empty() && t[0] == '-' && (t.size() == 1 || t[1] == ' '); }

    // position of the ':' that ends a mapping key on this line, or npos
    static size_t key_colon(const std::string& t) {
        if (t.empty()) return std::string::npos;
        size_t i = 0;
        if (t[0] == '"' || t[0] == '\'') {
            char q = t[0];
            for (i = 1; i < t.size(); ++i) {
                if (q == '"' && t[i] == '\\') { ++i; continue; }
                if (t[i] == q) {
                    if (q == '\'' && i + 1 < t.size() && t[i + 1] == '\'') { ++i; continue; }
                    break;
                }
            }
            if (i >= t.size()) return std::string::npos;
            ++i;
            while (i < t.size() && t[i] == ' ') ++i;
            return (i < t.size() && t[i] == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) ? i : std::string::npos;
        }
        if (t[0] == '[' || t[0] == '{' || t[0] == '|' || t[0] == '>' || t[0] == '&' || t[0] == '*' || t[0] == '!') return std::string::npos;
        for (i = 0; i < t.size(); ++i)
            if (t[i] == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
        return std::string::npos;
    }

    std::string unquote_key(const std::string& k, int line) {
        std::string t = rtrim(k);
        if (!t.empty() && (t[0] == '"' || t[0] == '\'')) {
            size_t p = 0;
            YNode n = quoted(t, p, line);
            return n.s;
        }
        if (!t.empty() && (t[0] == '?' )) fail(line, "complex mapping keys are not supported");
        return t;
    }

    YNode quoted(const std::string& t, size_t& p, int line) {
        YNode n;
        n.kind = YNode::SCALAR;
        n.quoted = true;
        n.line = line;
        char q = t[p++];
        // A line break inside a quoted scalar (multi-line scalars reach this function with their '\n' in place) is folded as YAML
        // folds flow scalars: white space around the break is dropped, one break becomes a space, k > 1 breaks become k - 1 newlines.
        auto fold_break = [&]() {
            while (!n.s.empty() && (n.s.back() == ' ' || n.s.back() == '\t')) n.s.pop_back();
            int breaks = 1;
            for (;;) {
                while (p < t.size() && (t[p] == ' ' || t[p] == '\t')) ++p;
                if (p < t.size() && t[p] == '\n') { ++breaks; ++p; continue; }
                break;
            }
            if (breaks == 1) n.s += ' ';
            else n.s.append((size_t)breaks - 1, '\n');
        };
        for (;;) {
            if (p >= t.size()) fail(line, "unterminated quoted scalar");
            char c = t[p++];
            if (c == '\n') { fold_break(); continue; }
            if (q == '\'') {
                if (c == '\'') {
                    if (p < t.size() && t[p] == '\'') { n.s += '\''; ++p; continue; }
                    return n;
                }
                n.s += c;
            } else {
                if (c == '"') return n;
                if (c != '\\') { n.s += c; continue; }
                if (p >= t.size()) fail(line, "unterminated escape");
                char e = t[p++];
                if (e == '\n') {   // escaped line break: the break disappears, the white space in front of it stays
                    while (p < t.size() && (t[p] == ' ' || t[p] == '\t')) ++p;
                    continue;
                }
                switch (e) {
                    case 'n': n.s += '\n'; break;
                    case 't': n.s += '\t'; break;
                    case 'r': n.s += '\r'; break;
                    case '0': n.s += '\0'; break;
                    case 'a': n.s += '\a'; break;
                    case 'b': n.s += '\b'; break;
                    case 'e': n.s += '\x1b'; break;
                    case 'f': n.s += '\f'; break;
                    case 'v': n.s += '\v'; break;
                    case ' ': n.s += ' '; break;
                    case '/': n.s += '/'; break;
                    case '"': n.s += '"'; break;
                    case '\\': n.s += '\\'; break;
                    case 'x': case 'u': case 'U': {
                        int digits = e == 'x' ? 2 : e == 'u' ? 4 : 8;
                        unsigned cp = 0;
                        for (int k = 0; k < digits; ++k) {
                            if (p >= t.size()) fail(line, "bad escape");
                            char h = t[p++];
                            cp <<= 4;
                            if (h >= '0' && h <= '9') cp |= h - '0';
                            else if (h >= 'a' && h <= 'f') cp |= h - 'a' + 10;
                            else if (h >= 'A' && h <= 'F') cp |= h - 'A' + 10;
                            else fail(line, "bad escape");
                        }
                        append_utf8(n.s, cp);
                        break;
                    }
                    default: fail(line, std::string("unknown escape \\") + e);
                }
            }
        }
    }

    // a flow value starting at t[p]; `in_flow`: stop plain scalars at , ] }
    YNode flow(const std::string& t, size_t& p, int line, bool in_flow) {
        while (p < t.size() && t[p] == ' ') ++p;
        YNode n;
        n.line = line;
        if (p >= t.size()) return n;
        char c = t[p];
        if (c == '[') {
            n.kind = YNode::SEQ;
            ++p;
            for (;;) {
                while (p < t.size() && t[p] == ' ') ++p;
                if (p >= t.size()) fail(line, "unterminated flow sequence");
                if (t[p] == ']') { ++p; return n; }
                n.seq.push_back(flow(t, p, line, true));
                while (p < t.size() && t[p] == ' ') ++p;
                if (p < t.size() && t[p] == ',') { ++p; continue; }
                if (p < t.size() && t[p] == ']') { ++p; return n; }
                fail(line, "expected ',' or ']' in flow sequence");
            }
        }
        if (c == '{') {
            n.kind = YNode::MAP;
            ++p;
            for (;;) {
                while (p < t.size() && t[p] == ' ') ++p;
                if (p >= t.size()) fail(line, "unterminated flow mapping");
                if (t[p] == '}') { ++p; return n; }
                YNode k = flow_key(t, p, line);
                while (p < t.size() && t[p] == ' ') ++p;
                YNode v;
                if (p < t.size() && t[p] == ':') { ++p; v = flow(t, p, line, true); }
                n.map.emplace_back(k.s, v);
                while (p < t.size() && t[p] == ' ') ++p;
                if (p < t.size() && t[p] == ',') { ++p; continue; }
                if (p < t.size() && t[p] == '}') { ++p; return n; }
                fail(line, "expected ',' or '}' in flow mapping");
            }
        }
        if (c == '"' || c == '\'') return quoted(t, p, line);
        if (c == '&' || c == '*' || c == '!') fail(line, "anchors, aliases and tags are not supported");
        n.kind = YNode::SCALAR;
        size_t b = p;
        while (p < t.size()) {
            if (in_flow && (t[p] == ',' || t[p] == ']' || t[p] == '}')) break;
            ++p;
        }
        n.s = rtrim(t.substr(b, p - b));
        return n;
    }

    YNode flow_key(const std::string& t, size_t& p, int line) {
        while (p < t.size() && t[p] == ' ') ++p;
        if (p < t.size() && (t[p] == '"' || t[p] == '\'')) return quoted(t, p, line);
        YNode n;
        n.kind = YNode::SCALAR;
        size_t b = p;
        while (p < t.size() && t[p] != ':' && t[p] != ',' && t[p] != '}') ++p;
        n.s = rtrim(t.substr(b, p - b));
        return n;
    }

    // bracket depth of `t` and whether it ends inside a quoted scalar
    static void flow_state(const std::string& t, int* depth_out, bool* in_quote) {
        int depth = 0;
        char q = 0;
        for (size_t i = 0; i < t.size(); ++i) {
            char c = t[i];
            if (q) {
                if (q == '"' && c == '\\') { ++i; continue; }
                if (c == q) q = 0;
            } else if (c == '"' || c == '\'') {
                // a quote opens a scalar only where a scalar can begin; inside a plain one (it's) it is an ordinary character
                size_t b = i;
                while (b > 0 && (t[b - 1] == ' ' || t[b - 1] == '\t' || t[b - 1] == '\n')) --b;
                if (b == 0 || t[b - 1] == '[' || t[b - 1] == '{' || t[b - 1] == ',' || t[b - 1] == ':') q = c;
            } else if (c == '[' || c == '{') ++depth;
            else if (c == ']' || c == '}') --depth;
        }
        *depth_out = depth;
        *in_quote = q != 0;
    }
    static bool balanced(const std::string& t) {
        int d;
        bool q;
        flow_state(t, &d, &q);
        return d <= 0 && !q;
    }
    static bool open_quote(const std::string& t) {
        int d;
        bool q;
        flow_state(t, &d, &q);
        return q;
    }

    // literal / folded block scalar whose header (`|`, `>`, with indicators) is `hdr`; content = following lines more
    // indented than `parent_indent`
    YNode block_scalar(const std::string& hdr, size_t& i, int parent_indent, int line) {
        YNode n;
        n.kind = YNode::SCALAR;
        n.quoted = true;
        n.line = line;
        const bool folded = hdr[0] == '>';
        char chomp = 0;
        int explicit_indent = 0;
        for (size_t k = 1; k < hdr.size(); ++k) {
            if (hdr[k] == '-' || hdr[k] == '+') chomp = hdr[k];
            else if (hdr[k] >= '1' && hdr[k] <= '9') explicit_indent = hdr[k] - '0';
            else fail(line, "bad block scalar header");
        }
        std::vector<std::string> content;
        int bi = explicit_indent ? parent_indent + explicit_indent : -1;
        size_t j = i;
        for (; j < lines_.size(); ++j) {
            const std::string& raw = lines_[j].raw;
            size_t ind = 0;
            while (ind < raw.size() && raw[ind] == ' ') ++ind;
            const bool empty = rtrim(raw).empty();
            if (empty) { content.push_back(""); continue; }
            if (bi < 0) {
                if ((int)ind <= parent_indent) break;
                bi = (int)ind;
            }
            if ((int)ind < bi) break;
            content.push_back(raw.substr((size_t)bi));
        }
        i = j;
        // trailing blank lines belong to the chomping decision
        size_t trailing = 0;
        while (!content.empty() && content.back().empty()) { content.pop_back(); ++trailing; }
        std::string out;
        for (size_t k = 0; k < content.size(); ++k) {
            if (!folded) {
                out += content[k];
                if (k + 1 < content.size()) out += '\n';
            } else {
                out += content[k];
                if (k + 1 < content.size()) {
                    // folding: a single line break between two non-empty, non-indented lines becomes a space
                    const bool more = content[k + 1].empty() || content[k + 1][0] == ' ' || content[k].empty() || content[k][0] == ' ';
                    if (content[k + 1].empty()) {
                        size_t e = k + 1;
                        while (e < content.size() && content[e].empty()) { out += '\n'; ++e; }
                        k = e - 1;
                    } else out += more ? '\n' : ' ';
                }
            }
        }
        if (!content.empty()) {
            if (chomp != '-') out += '\n';
            if (chomp == '+') out.append(trailing, '\n');
        } else if (chomp == '+') out.append(trailing, '\n');
        n.s = out;
        return n;
    }

    // an inline value (after `key: ` or `- `) starting on line i; may pull in continuation lines
    YNode inline_value(std::string v, size_t& i, int parent_indent, int line) {
        if (v[0] == '|' || v[0] == '>') return block_scalar(v, i, parent_indent, line);
        if (v[0] == '[' || v[0] == '{') {
            while (!balanced(v)) {
                if (open_quote(v)) {
                    // the collection breaks off inside a quoted scalar: the next line (blank ones included) continues the scalar
                    if (i >= lines_.size()) fail(line, "unterminated quoted scalar");
                    v += "\n" + lines_[i].raw;
                    ++i;
                    continue;
                }
                size_t j = next(i);
                if (j >= lines_.size()) fail(line, "unterminated flow collection");
                v += " " + lines_[j].text;
                i = j + 1;
            }
            size_t p = 0;
            YNode n = flow(v, p, line, false);
            while (p < v.size() && v[p] == ' ') ++p;
            if (p < v.size()) fail(line, "unexpected characters after flow collection");
            return n;
        }
        if (v[0] == '"' || v[0] == '\'') {
            // a quoted scalar may continue on the following lines (folded)
            for (;;) {
                bool closed = false;
                char q = v[0];
                for (size_t k = 1; k < v.size(); ++k) {
                    if (q == '"' && v[k] == '\\') { ++k; continue; }
                    if (v[k] == q) {
                        if (q == '\'' && k + 1 < v.size() && v[k + 1] == '\'') { ++k; continue; }
                        closed = true;
                        break;
                    }
                }
                if (closed) break;
                if (i >= lines_.size()) fail(line, "unterminated quoted scalar");
                v += "\n" + lines_[i].raw;   // folded by quoted()
                ++i;
            }
            size_t p = 0;
            YNode n = quoted(v, p, line);
            while (p < v.size() && v[p] == ' ') ++p;
            if (p < v.size()) fail(line, "unexpected characters after quoted scalar");
            return n;
        }
        if (v[0] == '&' || v[0] == '*' || v[0] == '!') fail(line, "anchors, aliases and tags are not supported");
        // plain scalar: following lines that are more indented continue it (folded with one space)
        if (v.find(": ") != std::string::npos || v.back() == ':') fail(line, "mapping values are not allowed here");
        YNode n;
        n.kind = YNode::SCALAR;
        n.line = line;
        n.s = v;
        for (;;) {
            size_t j = next(i);
            if (j >= lines_.size() || lines_[j].indent <= parent_indent) break;
            if (lines_[j].text.find(": ") != std::string::npos || lines_[j].text.back() == ':') fail(lines_[j].no, "mapping values are not allowed here");
            n.s += " " + lines_[j].text;
            i = j + 1;
        }
        return n;
    }

    YNode block(size_t& i, int indent) {
        i = next(i);
        const Line& l = lines_[i];
        if (is_seq_item(l.text)) return sequence(i, indent);
        if (key_colon(l.text) != std::string::npos) return mapping(i, indent);
        // a bare scalar / flow collection as the whole block
        std::string v = l.text;
        int line = l.no;
        ++i;
        return inline_value(v, i, indent - 1, line);
    }

    YNode mapping(size_t& i, int indent) {
        YNode n;
        n.kind = YNode::MAP;
        n.line = lines_[i].no;
        for (;;) {
            i = next(i);
            if (i >= lines_.size() || lines_[i].indent < indent) break;
            if (lines_[i].indent > indent) fail(lines_[i].no, "bad indentation of a mapping entry");
            const std::string t = lines_[i].text;
            const int line = lines_[i].no;
            if (is_seq_item(t)) break;  // `key:` followed by a sequence at the same indentation ends up here for the parent
            size_t c = key_colon(t);
            if (c == std::string::npos) fail(line, "expected `key: value`");
            std::string key = unquote_key(t.substr(0, c), line);
            for (auto& kv : n.map)
                if (kv.first == key) fail(line, "duplicate key: " + key);
            std::string rest = ltrim(t.substr(c + 1));
            ++i;
            YNode v;
            v.line = line;
            if (rest.empty()) {
                size_t j = next(i);
                if (j < lines_.size() && lines_[j].indent > indent) {
                    i = j;
                    v = block(i, lines_[j].indent);
                } else if (j < lines_.size() && lines_[j].indent == indent && is_seq_item(lines_[j].text)) {
                    i = j;
                    v = sequence(i, indent);
                }
            } else v = inline_value(rest, i, indent, line);
            n.map.emplace_back(key, v);
        }
        return n;
    }

    YNode sequence(size_t& i, int indent) {
        YNode n;
        n.kind = YNode::SEQ;
        n.line = lines_[i].no;
        for (;;) {
            i = next(i);
            if (i >= lines_.size() || lines_[i].indent != indent || !is_seq_item(lines_[i].text)) {
                if (i < lines_.size() && lines_[i].indent > indent) fail(lines_[i].no, "bad indentation of a sequence entry");
                break;
            }
            const std::string t = lines_[i].text;
            const int line = lines_[i].no;
            std::string rest = t.size() > 1 ? t.substr(2) : std::string();
            size_t lead = 0;
            while (lead < rest.size() && rest[lead] == ' ') ++lead;
            rest = rest.substr(lead);
            if (rest.empty()) {
                ++i;
                size_t j = next(i);
                YNode v;
                if (j < lines_.size() && lines_[j].indent > indent) {
                    i = j;
                    v = block(i, lines_[j].indent);
                }
                n.seq.push_back(v);
                continue;
            }
            const int col = indent + 2 + (int)lead;
            if (is_seq_item(rest) || key_colon(rest) != std::string::npos) {
                // a nested block starts on the dash line: re-read the line as if it began at that column
                lines_[i].indent = col;
                lines_[i].text = rest;
                n.seq.push_back(block(i, col));
            } else {
                ++i;
                n.seq.push_back(inline_value(rest, i, indent, line));
            }
        }
        return n;
    }
};

void dump(const YNode& n, std::string& o) {
    auto str = [&](const std::string& s) {
        o += '"';
        for (unsigned char c : s) {
            if (c == '"' || c == '\\') { o += '\\'; o += (char)c; }
            else if (c == '\n') o += "\\n";
            else if (c == '\t') o += "\\t";
            else if (c == '\r') o += "\\r";
            else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
            else o += (char)c;
        }
        o += '"';
    };
    switch (n.kind) {
        case YNode::NUL: o += "null"; break;
        case YNode::SCALAR:
            if (n.is_null()) o += "null";
            else str(n.s);
            break;
        case YNode::SEQ:
            o += '[';
            for (size_t i = 0; i < n.seq.size(); ++i) {
                if (i) o += ',';
                dump(n.seq[i], o);
            }
            o += ']';
            break;
        case YNode::MAP:
            o += '{';
            for (size_t i = 0; i < n.map.size(); ++i) {
                if (i) o += ',';
                str(n.map[i].first);
                o += ':';
                dump(n.map[i].second, o);
            }
            o += '}';
            break;
    }
}

}  // namespace

bool yaml_parse(const std::string& text, YNode* root, std::string& err) {
    try {
        Parser p(text);
        *root = p.parse();
        return true;
    } catch (const Fail& f) {
        err = f.msg;
        return false;
    }
}

std::string yaml_dump(const YNode& n) {
    std::string o;
    dump(n, o);
    return o;
}

}  // namespace pgw
