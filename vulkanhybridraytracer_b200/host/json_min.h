// json_min.h -- small JSON DOM for the glTF reader (the reference parses glTF through tinygltf + nlohmann::json,
// both un-vendored upstream, .gitignore:7).  RFC 8259 values; numbers are kept as double like tinygltf does.
#pragma once
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace jsonmin {

struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type{Null};
    bool boolean{false};
    double number{0.0};
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;   // insertion order kept

    bool is_object() const { return type == Object; }
    bool is_array() const { return type == Array; }
    bool is_number() const { return type == Number; }
    bool is_string() const { return type == String; }
    size_t size() const { return type == Array ? arr.size() : type == Object ? obj.size() : 0; }
    const Value* find(const char* key) const
    {
        if (type != Object) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const Value& at(size_t i) const
    {
        if (type != Array || i >= arr.size()) throw std::runtime_error("json: array index out of range");
        return arr[i];
    }
    // typed getters with the defaults tinygltf applies when a property is absent
    int get_int(const char* key, int dflt) const { const Value* v = find(key); return v && v->is_number() ? (int)v->number : dflt; }
    double get_num(const char* key, double dflt) const { const Value* v = find(key); return v && v->is_number() ? v->number : dflt; }
    std::string get_str(const char* key, const char* dflt = "") const { const Value* v = find(key); return v && v->is_string() ? v->str : std::string(dflt); }
    std::vector<double> get_numbers(const char* key) const
    {
        std::vector<double> out;
        const Value* v = find(key);
        if (v && v->is_array()) for (const Value& e : v->arr) if (e.is_number()) out.push_back(e.number);
        return out;
    }
};

class Parser {
public:
    Parser(const char* begin, const char* end) : p_(begin), end_(end) {}
    Value parse()
    {
        if (end_ - p_ >= 3 && (unsigned char)p_[0] == 0xEF && (unsigned char)p_[1] == 0xBB && (unsigned char)p_[2] == 0xBF) p_ += 3;   // BOM
        Value v = value(0);
        ws();
        if (p_ != end_) fail("trailing characters");
        return v;
    }

private:
    const char* p_;
    const char* end_;

    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("json: ") + what); }
    void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }
    bool eat(char c) { ws(); if (p_ < end_ && *p_ == c) { ++p_; return true; } return false; }
    void expect(const char* lit) { size_t n = std::strlen(lit); if ((size_t)(end_ - p_) < n || std::memcmp(p_, lit, n) != 0) fail("bad literal"); p_ += n; }

    static void utf8(std::string& s, unsigned cp)
    {
        if (cp < 0x80) s += (char)cp;
        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
        else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4()
    {
        if (end_ - p_ < 4) fail("bad \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string string()
    {
        std::string s;
        if (p_ >= end_ || *p_ != '"') fail("expected string");
        ++p_;
        while (true) {
            if (p_ >= end_) fail("unterminated string");
            char c = *p_++;
            if (c == '"') break;
            if (c != '\\') { s += c; continue; }
            if (p_ >= end_) fail("unterminated escape");
            char e = *p_++;
            switch (e) {
            case '"': s += '"'; break;
            case '\\': s += '\\'; break;
            case '/': s += '/'; break;
            case 'b': s += '\b'; break;
            case 'f': s += '\f'; break;
            case 'n': s += '\n'; break;
            case 'r': s += '\r'; break;
            case 't': s += '\t'; break;
            case 'u': {
                unsigned cp = hex4();
                if (cp >= 0xD800 && cp < 0xDC00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                    p_ += 2;
                    unsigned lo = hex4();
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                utf8(s, cp);
                break;
            }
            default: fail("bad escape");
            }
        }
        return s;
    }
    Value value(int depth)
    {
        if (depth > 256) fail("nesting too deep");
        ws();
        if (p_ >= end_) fail("unexpected end");
        Value v;
        char c = *p_;
        if (c == '{') {
            ++p_;
            v.type = Value::Object;
            if (eat('}')) return v;
            while (true) {
                ws();
                std::string key = string();
                if (!eat(':')) fail("expected ':'");
                v.obj.emplace_back(std::move(key), value(depth + 1));
                if (eat(',')) continue;
                if (eat('}')) break;
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            ++p_;
            v.type = Value::Array;
            if (eat(']')) return v;
            while (true) {
                v.arr.push_back(value(depth + 1));
                if (eat(',')) continue;
                if (eat(']')) break;
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.type = Value::String;
            v.str = string();
        } else if (c == 't') { expect("true"); v.type = Value::Bool; v.boolean = true; }
        else if (c == 'f') { expect("false"); v.type = Value::Bool; v.boolean = false; }
        else if (c == 'n') { expect("null"); v.type = Value::Null; }
        else {
            const char* q = p_;
            if (q < end_ && *q == '-') ++q;
            if (q >= end_ || *q < '0' || *q > '9') fail("unexpected character");
            while (q < end_ && ((*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '+' || *q == '-')) ++q;
            std::string num(p_, q);
            char* stop = nullptr;
            v.type = Value::Number;
            v.number = std::strtod(num.c_str(), &stop);
            if (stop == num.c_str() || *stop != '\0') fail("bad number");
            p_ = q;
        }
        return v;
    }
};

inline Value parse(const std::string& text) { return Parser(text.data(), text.data() + text.size()).parse(); }

} // namespace jsonmin
