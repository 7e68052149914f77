// toml_lite.hpp -- a small TOML reader for scene files (the reference uses toml++, which is not available here).
// Covers what scene files use and a bit more: comments, bare / quoted / dotted keys, [tables], [[arrays of tables]],
// inline tables, (multi-line) arrays with trailing commas, basic and literal strings, integers (dec/hex/oct/bin, '_'),
// floats (exponent, inf, nan), booleans.  Dates and multi-line strings are rejected with a message.
#pragma once
#include <cctype>
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

    std::string parse_basic_string()
    {
        next(); // "
        if (src_.compare(pos_, 2, "\"\"") == 0) fail("multi-line strings are not supported");
        std::string out;
        for (;;)
        {
            if (eof() || peek() == '\n') fail("unterminated string");
            char c = next();
            if (c == '"') break;
            if (c == '\\')
            {
                c = next();
                switch (c)
                {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case '\\': out += '\\'; break;
                    case '"': out += '"'; break;
                    default: fail("unsupported escape sequence");
                }
            }
            else
                out += c;
        }
        return out;
    }
    std::string parse_literal_string()
    {
        next(); // '
        if (src_.compare(pos_, 2, "''") == 0) fail("multi-line strings are not supported");
        std::string out;
        for (;;)
        {
            if (eof() || peek() == '\n') fail("unterminated string");
            const char c = next();
            if (c == '\'') break;
            out += c;
        }
        return out;
    }
    std::string parse_key_part()
    {
        skip_ws();
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
        std::string t;
        for (char c : tok)
            if (c != '_') t += c;
        const std::string body = (!t.empty() && (t[0] == '+' || t[0] == '-')) ? t.substr(1) : t;
        if (body == "inf" || body == "nan")
        {
            n->kind = node::floating;
            n->f = body == "inf" ? INFINITY : NAN;
            if (t[0] == '-') n->f = -n->f;
            return n;
        }
        if (t.find(':') != std::string::npos || (t.size() > 4 && t[4] == '-' && std::isdigit(static_cast<unsigned char>(t[0])))) fail("dates and times are not supported");
        if (t.empty()) fail("expected a value");
        char* end = nullptr;
        if (body.size() > 2 && body[0] == '0' && (body[1] == 'x' || body[1] == 'o' || body[1] == 'b'))
        {
            const int base = body[1] == 'x' ? 16 : (body[1] == 'o' ? 8 : 2);
            n->kind = node::integer;
            n->i = std::strtoll(body.c_str() + 2, &end, base);
            if (*end) fail("malformed integer '" + tok + "'");
            return n;
        }
        if (t.find_first_of(".eE") != std::string::npos)
        {
            n->kind = node::floating;
            n->f = std::strtod(t.c_str(), &end);
            if (*end) fail("malformed number '" + tok + "'");
            return n;
        }
        n->kind = node::integer;
        n->i = std::strtoll(t.c_str(), &end, 10);
        if (*end) fail("malformed number '" + tok + "'");
        return n;
    }

    node_ptr parse_value()
    {
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

    node_ptr parse()
    {
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
