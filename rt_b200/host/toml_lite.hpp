// toml_lite.hpp -- a small TOML reader for scene files (the reference uses toml++, which is not available here).
// Covers what scene files use and a bit more: comments, bare / quoted / dotted keys, [tables], [[arrays of tables]],
// inline tables, (multi-line) arrays with trailing commas, basic and literal strings, integers (dec/hex/oct/bin, '_'),
// floats (exponent, inf, nan), booleans, multi-line strings, every escape.  Dates and times are rejected with a message.
#pragma once
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace toml_lite {

struct node;
using node_ptr = std::shared_ptr<node>;

struct node {
    enum kind_t { integer, floating, boolean, string, array, table } kind = table;
    int64_t i = 0;
    double f = 0.0;
    bool b = false;
    std::string s;
    std::vector<node_ptr> items;                        // array
    std::vector<std::pair<std::string, node_ptr>> kv;   // table, insertion order
    bool inline_or_defined = false;                      // inline table / explicitly defined (no re-opening)
    int line = 0;

    const node* get(const std::string& key) const
    {
        for (auto& p : kv)
            if (p.first == key) return p.second.get();
        return nullptr;
    }
    const char* type_name() const
    {
        switch (kind)
        {
            case integer: return "integer";
            case floating: return "floating-point";
            case boolean: return "boolean";
            case string: return "string";
            case array: return "array";
            default: return "table";
        }
    }
};

struct parse_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class parser {
    const std::string& src_;
    size_t pos_ = 0;
    int line_ = 1;

    [[noreturn]] void fail(const std::string& msg) const { throw parse_error("TOML parse error at line " + std::to_string(line_) + ": " + msg); }
    bool eof() const { return pos_ >= src_.size(); }
    char peek() const { return eof() ? '\0' : src_[pos_]; }
    char next()
    {
        const char c = peek();
        if (c == '\n') line_++;
        pos_++;
        return c;
    }
    void skip_ws() { while (peek() == ' ' || peek() == '\t') pos_++; }
    void skip_comment() { if (peek() == '#') while (!eof() && peek() != '\n') pos_++; }
    void skip_ws_nl()
    {
        for (;;)
        {
            skip_ws();
            skip_comment();
            if (peek() == '\n' || peek() == '\r') next();
            else break;
        }
    }
    void expect_eol()
    {
        skip_ws();
        skip_comment();
        if (peek() == '\r') next();
        if (!eof() && peek() != '\n') fail(std::string("unexpected '") + peek() + "' after value");
        if (!eof()) next();
    }

    void append_utf8(std::string& out, uint32_t cp)
    {
        if (cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) fail("escape is not a Unicode scalar value");
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    void parse_escape(std::string& out) // after the backslash
    {
        const char c = next();
        switch (c)
        {
            case 'b': out += '\b'; break;
            case 't': out += '\t'; break;
            case 'n': out += '\n'; break;
            case 'f': out += '\f'; break;
            case 'r': out += '\r'; break;
            case '\\': out += '\\'; break;
            case '"': out += '"'; break;
            case 'u':
            case 'U':
            {
                uint32_t cp = 0;
                for (int k = 0, n = c == 'u' ? 4 : 8; k < n; k++)
                {
                    const char h = next();
                    if (!std::isxdigit(static_cast<unsigned char>(h))) fail("malformed unicode escape");
                    cp = cp * 16 + (uint32_t)(std::isdigit(static_cast<unsigned char>(h)) ? h - '0' : (std::tolower(h) - 'a' + 10));
                }
                append_utf8(out, cp);
                break;
            }
            default: fail("unsupported escape sequence");
        }
    }
    static bool control(char c) { return (static_cast<unsigned char>(c) < 0x20 && c != '\t') || c == 0x7F; }
    // the body of a multi-line string, after the opening delimiter: a newline right after it is dropped, up to two quotes may
    // precede the closing delimiter, and (basic strings) a backslash at the end of a line swallows the whitespace that follows
    std::string parse_multiline(char q)
    {
        const bool basic = q == '"';
        if (peek() == '\r' && pos_ + 1 < src_.size() && src_[pos_ + 1] == '\n') next();
        if (peek() == '\n') next();
        std::string out;
        for (;;)
        {
            if (eof()) fail("unterminated multi-line string");
            const char c = next();
            if (c == q)
            {
                size_t run = 1;
                while (peek() == q && run < 5) { next(); run++; }
                if (run >= 3)
                {
                    out.append(run - 3, q); // """"" ends a string that ends in two quotes
                    return out;
                }
                out.append(run, q);
            }
            else if (basic && c == '\\')
            {
                size_t p = pos_;
                while (p < src_.size() && (src_[p] == ' ' || src_[p] == '\t')) p++;
                if (p < src_.size() && (src_[p] == '\n' || (src_[p] == '\r' && p + 1 < src_.size() && src_[p + 1] == '\n')))
                    while (!eof() && (peek() == ' ' || peek() == '\t' || peek() == '\n' || peek() == '\r')) next(); // line-ending backslash
                else
                    parse_escape(out);
            }
            else if (c == '\r')
            {
                if (peek() != '\n') fail("bare carriage return in a string");
            }
            else if (c != '\n' && control(c))
                fail("control character in a string");
            else
                out += c;
        }
    }
    std::string parse_basic_string()
    {
        next(); // "
        if (src_.compare(pos_, 2, "\"\"") == 0)
        {
            next(); next();
            return parse_multiline('"');
        }
        std::string out;
        for (;;)
        {
            if (eof() || peek() == '\n') fail("unterminated string");
            const char c = next();
            if (c == '"') break;
            if (c == '\\') parse_escape(out);
            else if (control(c)) fail("control character in a string");
            else out += c;
        }
        return out;
    }
    std::string parse_literal_string()
    {
        next(); // '
        if (src_.compare(pos_, 2, "''") == 0)
        {
            next(); next();
            return parse_multiline('\'');
        }
        std::string out;
        for (;;)
        {
            if (eof() || peek() == '\n') fail("unterminated string");
            const char c = next();
            if (c == '\'') break;
            if (control(c)) fail("control character in a string");
            out += c;
        }
        return out;
    }
    std::string parse_key_part()
    {
        skip_ws();
        if (src_.compare(pos_, 3, "\"\"\"") == 0 || src_.compare(pos_, 3, "'''") == 0) fail("a key cannot be a multi-line string");
        if (peek() == '"') return parse_basic_string();
        if (peek() == '\'') return parse_literal_string();
        std::string k;
        while (std::isalnum(static_cast<unsigned char>(peek())) || peek() == '_' || peek() == '-') k += next();
        if (k.empty()) fail("expected a key");
        return k;
    }
    std::vector<std::string> parse_key()
    {
        std::vector<std::string> parts{ parse_key_part() };
        skip_ws();
        while (peek() == '.')
        {
            next();
            parts.push_back(parse_key_part());
            skip_ws();
        }
        return parts;
    }

    node_ptr parse_number_or_bool()
    {
        std::string tok;
        while (!eof() && (std::isalnum(static_cast<unsigned char>(peek())) || peek() == '+' || peek() == '-' || peek() == '.' || peek() == '_' || peek() == ':'))
            tok += next();
        auto n = std::make_shared<node>();
        n->line = line_;
        if (tok == "true" || tok == "false")
        {
            n->kind = node::boolean;
            n->b = tok == "true";
            return n;
        }
        // TOML's number grammar, checked on the token itself (strtod / strtoll alone would accept "1.", ".5", "007", "1__0", hex floats)
        auto digits = [&](size_t& p, auto is_digit) { // digit ( '_'? digit )*
            if (p >= tok.size() || !is_digit(tok[p])) return false;
            for (p++; p < tok.size();)
            {
                if (is_digit(tok[p])) p++;
                else if (tok[p] == '_' && p + 1 < tok.size() && is_digit(tok[p + 1])) p += 2;
                else break;
            }
            return true;
        };
        auto dec = [](char c) { return c >= '0' && c <= '9'; };
        std::string t;
        for (char c : tok)
            if (c != '_') t += c;
        const bool sign = !tok.empty() && (tok[0] == '+' || tok[0] == '-');
        const std::string body = sign ? t.substr(1) : t;
        if (body == "inf" || body == "nan")
        {
            if (t.size() != tok.size()) fail("malformed number '" + tok + "'");
            n->kind = node::floating;
            n->f = body == "inf" ? INFINITY : NAN;
            if (t[0] == '-') n->f = -n->f;
            return n;
        }
        if (tok.empty()) fail("expected a value");
        const bool date_like = tok.size() > 4 && dec(tok[0]) && dec(tok[1]) && dec(tok[2]) && dec(tok[3]) && tok[4] == '-';
        const bool time_like = tok.size() > 2 && dec(tok[0]) && dec(tok[1]) && tok[2] == ':';
        if (date_like || time_like) fail("dates and times are not supported");
        char* end = nullptr;
        if (!sign && tok.size() > 2 && tok[0] == '0' && (tok[1] == 'x' || tok[1] == 'o' || tok[1] == 'b'))
        {
            const int base = tok[1] == 'x' ? 16 : (tok[1] == 'o' ? 8 : 2);
            size_t p = 2;
            const bool ok = base == 16 ? digits(p, [](char c) { return std::isxdigit(static_cast<unsigned char>(c)) != 0; })
                          : base == 8 ? digits(p, [](char c) { return c >= '0' && c <= '7'; })
                                      : digits(p, [](char c) { return c == '0' || c == '1'; });
            if (!ok || p != tok.size()) fail("malformed integer '" + tok + "'");
            errno = 0;
            const unsigned long long u = std::strtoull(body.c_str() + 2, &end, base);
            if (*end || errno == ERANGE || u > (unsigned long long)INT64_MAX) fail("integer '" + tok + "' is out of range");
            n->kind = node::integer;
            n->i = (int64_t)u;
            return n;
        }
        // [+-] ( '0' | nonzero digits ) [ '.' digits ] [ (e|E) [+-] digits ]
        size_t p = sign ? 1 : 0;
        const size_t int_begin = p;
        if (!digits(p, dec)) fail("malformed number '" + tok + "'");
        if (tok[int_begin] == '0' && p - int_begin > 1) fail("malformed number '" + tok + "' (leading zero)");
        bool is_float = false;
        if (p < tok.size() && tok[p] == '.')
        {
            is_float = true;
            if (!digits(++p, dec)) fail("malformed number '" + tok + "'");
        }
        if (p < tok.size() && (tok[p] == 'e' || tok[p] == 'E'))
        {
            is_float = true;
            p++;
            if (p < tok.size() && (tok[p] == '+' || tok[p] == '-')) p++;
            if (!digits(p, dec)) fail("malformed number '" + tok + "'");
        }
        if (p != tok.size()) fail("malformed number '" + tok + "'");
        if (is_float)
        {
            n->kind = node::floating;
            n->f = std::strtod(t.c_str(), &end);
            if (*end) fail("malformed number '" + tok + "'");
            return n;
        }
        errno = 0;
        n->kind = node::integer;
        n->i = std::strtoll(t.c_str(), &end, 10);
        if (*end) fail("malformed number '" + tok + "'");
        if (errno == ERANGE) fail("integer '" + tok + "' is out of range");
        return n;
    }

    // arrays and inline tables nest by recursion: bounded like toml++'s parser (TOML_MAX_NESTED_VALUES = 256), so that a document of
    // 200 000 opening brackets is a parse error (main.cpp:370-379: "error: ...", exit 1), not a stack overflow
    static constexpr int max_nesting = 256;
    int depth_ = 0;
    struct depth_guard {
        int& d;
        explicit depth_guard(int& depth) : d(depth) { ++d; }
        ~depth_guard() { --d; }
    };

    node_ptr parse_value()
    {
        const depth_guard nested{ depth_ };
        if (depth_ > max_nesting) fail("exceeded maximum nested value depth of " + std::to_string(max_nesting));
        skip_ws();
        auto n = std::make_shared<node>();
        n->line = line_;
        const char c = peek();
        if (c == '"' || c == '\'')
        {
            n->kind = node::string;
            n->s = c == '"' ? parse_basic_string() : parse_literal_string();
            return n;
        }
        if (c == '[')
        {
            next();
            n->kind = node::array;
            for (;;)
            {
                skip_ws_nl();
                if (peek() == ']') { next(); break; }
                n->items.push_back(parse_value());
                skip_ws_nl();
                if (peek() == ',') { next(); continue; }
                if (peek() == ']') { next(); break; }
                fail("expected ',' or ']' in array");
            }
            return n;
        }
        if (c == '{')
        {
            next();
            n->kind = node::table;
            n->inline_or_defined = true;
            skip_ws();
            if (peek() == '}') { next(); return n; }
            for (;;)
            {
                skip_ws();
                const auto key = parse_key();
                skip_ws();
                if (next() != '=') fail("expected '=' in inline table");
                insert(*n, key, parse_value());
                skip_ws();
                if (peek() == ',') { next(); continue; }
                if (peek() == '}') { next(); break; }
                fail("expected ',' or '}' in inline table");
            }
            return n;
        }
        return parse_number_or_bool();
    }

    node& descend(node& root, const std::vector<std::string>& path, size_t count)
    {
        node* cur = &root;
        for (size_t k = 0; k < count; k++)
        {
            node* child = nullptr;
            for (auto& p : cur->kv)
                if (p.first == path[k]) child = p.second.get();
            if (!child)
            {
                auto t = std::make_shared<node>();
                t->kind = node::table;
                t->line = line_;
                cur->kv.emplace_back(path[k], t);
                child = t.get();
            }
            if (child->kind == node::array && !child->items.empty() && child->items.back()->kind == node::table)
                child = child->items.back().get(); // [[a]] then [a.b]
            if (child->kind != node::table) fail("key '" + path[k] + "' is not a table");
            cur = child;
        }
        return *cur;
    }
    void insert(node& tbl, const std::vector<std::string>& key, node_ptr value)
    {
        node& parent = descend(tbl, key, key.size() - 1);
        if (parent.get(key.back())) fail("duplicate key '" + key.back() + "'");
        parent.kv.emplace_back(key.back(), std::move(value));
    }

  public:
    explicit parser(const std::string& src) : src_(src) {}

    // a TOML document is UTF-8: reject overlong forms, surrogates, values past U+10FFFF and stray continuation bytes up front
    void check_utf8() const
    {
        int line = 1;
        for (size_t i = 0; i < src_.size();)
        {
            const unsigned char c = (unsigned char)src_[i];
            if (c == '\n') line++;
            size_t len = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 0;
            bool ok = len != 0 && i + len <= src_.size();
            uint32_t cp = len == 1 ? c : len == 2 ? c & 0x1Fu : len == 3 ? c & 0x0Fu : c & 0x07u;
            for (size_t k = 1; ok && k < len; k++)
            {
                const unsigned char d = (unsigned char)src_[i + k];
                ok = (d & 0xC0) == 0x80;
                cp = (cp << 6) | (d & 0x3Fu);
            }
            static const uint32_t smallest[5] = { 0, 0, 0x80, 0x800, 0x10000 };
            if (ok && len > 1) ok = cp >= smallest[len] && cp <= 0x10FFFF && !(cp >= 0xD800 && cp <= 0xDFFF);
            if (!ok) throw parse_error("TOML parse error at line " + std::to_string(line) + ": the document is not valid UTF-8");
            i += len;
        }
    }

    node_ptr parse()
    {
        check_utf8();
        auto root = std::make_shared<node>();
        root->kind = node::table;
        node* current = root.get();
        for (;;)
        {
            skip_ws_nl();
            if (eof()) break;
            if (peek() == '[')
            {
                next();
                const bool is_array = peek() == '[';
                if (is_array) next();
                const auto key = parse_key();
                skip_ws();
                if (next() != ']') fail("expected ']'");
                if (is_array && next() != ']') fail("expected ']]'");
                expect_eol();
                node& parent = descend(*root, key, key.size() - 1);
                node* existing = nullptr;
                for (auto& p : parent.kv)
                    if (p.first == key.back()) existing = p.second.get();
                if (is_array)
                {
                    if (!existing)
                    {
                        auto arr = std::make_shared<node>();
                        arr->kind = node::array;
                        arr->line = line_;
                        parent.kv.emplace_back(key.back(), arr);
                        existing = arr.get();
                    }
                    if (existing->kind != node::array) fail("'" + key.back() + "' is not an array of tables");
                    auto t = std::make_shared<node>();
                    t->kind = node::table;
                    t->line = line_;
                    existing->items.push_back(t);
                    current = t.get();
                }
                else
                {
                    if (existing && (existing->kind != node::table || existing->inline_or_defined)) fail("table '" + key.back() + "' defined twice");
                    if (!existing)
                    {
                        auto t = std::make_shared<node>();
                        t->kind = node::table;
                        t->line = line_;
                        parent.kv.emplace_back(key.back(), t);
                        existing = t.get();
                    }
                    existing->inline_or_defined = true;
                    current = existing;
                }
                continue;
            }
            const auto key = parse_key();
            skip_ws();
            if (next() != '=') fail("expected '=' after key '" + key.back() + "'");
            insert(*current, key, parse_value());
            expect_eol();
        }
        return root;
    }
};

inline node_ptr parse(const std::string& text) { return parser(text).parse(); }

} // namespace toml_lite
