// Minimal JSON DOM reader for the scene / mesh files (replaces the reference's use of the
// vendored nlohmann/json, Raytracer.cpp:3, 602, 646).  Numbers keep nlohmann's distinction:
// integer syntax -> int64/uint64, anything with '.', 'e' or 'E' -> double via strtod; the
// caller converts to float with a static_cast, so every value ends up as the same float the
// reference loader stores (double->float or int64->float, one rounding).
#pragma once
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace jsonmin {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
    enum Kind { Null, Bool, Int, UInt, Float, String, Array, Object } kind = Null;
    bool b = false;
    int64_t i = 0;
    uint64_t u = 0;
    double d = 0.0;
    std::string s;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;   // insertion order; duplicate keys: last wins on lookup

    bool is_array() const { return kind == Array; }
    bool is_object() const { return kind == Object; }
    bool is_number() const { return kind == Int || kind == UInt || kind == Float; }
    bool contains(const std::string& key) const { return kind == Object && find(key) != nullptr; }
    const Value* find(const std::string& key) const {
        const Value* r = nullptr;
        for (auto& kv : obj) if (kv.first == key) r = kv.second.get();
        return r;
    }
    const Value& at(const std::string& key) const {
        if (kind != Object) throw std::runtime_error("cannot use key '" + key + "' on a non-object");
        const Value* v = find(key);
        if (!v) throw std::runtime_error("key '" + key + "' not found");
        return *v;
    }
    const Value& at(size_t idx) const {
        if (kind != Array) throw std::runtime_error("cannot index a non-array");
        if (idx >= arr.size()) throw std::runtime_error("array index " + std::to_string(idx) + " is out of range");
        return *arr[idx];
    }
    size_t size() const { return kind == Array ? arr.size() : (kind == Object ? obj.size() : 0); }
    float as_float() const {
        switch (kind) {
        case Int: return static_cast<float>(i);
        case UInt: return static_cast<float>(u);
        case Float: return static_cast<float>(d);
        case Bool: return b ? 1.0f : 0.0f;      // nlohmann converts booleans too
        default: throw std::runtime_error("type must be number");
        }
    }
    int as_int() const {
        switch (kind) {
        case Int: return static_cast<int>(i);
        case UInt: return static_cast<int>(u);
        case Float: return static_cast<int>(d);
        case Bool: return b ? 1 : 0;
        default: throw std::runtime_error("type must be number");
        }
    }
    const std::string& as_string() const {
        if (kind != String) throw std::runtime_error("type must be string");
        return s;
    }
};

class Parser {
public:
    explicit Parser(const std::string& text) : p_(text.c_str()), end_(text.c_str() + text.size()) {}
    ValuePtr parse() {
        // tolerate a UTF-8 byte order mark like nlohmann does
        if (end_ - p_ >= 3 && (unsigned char)p_[0] == 0xEF && (unsigned char)p_[1] == 0xBB && (unsigned char)p_[2] == 0xBF) p_ += 3;
        ValuePtr v = value();
        ws();
        if (p_ != end_) fail("unexpected trailing characters");
        return v;
    }

private:
    const char* p_;
    const char* end_;
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("parse error: ") + what); }
    void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++; }
    ValuePtr value() {
        ws();
        if (p_ >= end_) fail("unexpected end of input");
        switch (*p_) {
        case '{': return object();
        case '[': return array();
        case '"': { auto v = std::make_shared<Value>(); v->kind = Value::String; v->s = string(); return v; }
        case 't': literal("true"); { auto v = std::make_shared<Value>(); v->kind = Value::Bool; v->b = true; return v; }
        case 'f': literal("false"); { auto v = std::make_shared<Value>(); v->kind = Value::Bool; v->b = false; return v; }
        case 'n': literal("null"); return std::make_shared<Value>();
        default: return number();
        }
    }
    void literal(const char* lit) {
        size_t n = strlen(lit);
        if ((size_t)(end_ - p_) < n || strncmp(p_, lit, n) != 0) fail("invalid literal");
        p_ += n;
    }
    std::string string() {
        std::string out;
        p_++;   // opening quote
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\') {
                p_++;
                if (p_ >= end_) fail("bad escape");
                switch (*p_) {
                case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
                case 'r': out += '\r'; break; case 't': out += '\t'; break;
                case 'u': {
                    if (end_ - p_ < 5) fail("bad \\u escape");
                    unsigned cp = (unsigned)strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
                    p_ += 4;
                    if (cp < 0x80) out += (char)cp;
                    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                    else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: fail("bad escape");
                }
                p_++;
            } else out += *p_++;
        }
        if (p_ >= end_) fail("unterminated string");
        p_++;
        return out;
    }
    ValuePtr number() {
        const char* start = p_;
        bool is_float = false;
        if (p_ < end_ && *p_ == '-') p_++;
        if (p_ >= end_ || !(*p_ >= '0' && *p_ <= '9')) fail("invalid number");
        while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++;
        if (p_ < end_ && *p_ == '.') { is_float = true; p_++; while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++; }
        if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
            is_float = true; p_++;
            if (p_ < end_ && (*p_ == '+' || *p_ == '-')) p_++;
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++;
        }
        std::string tok(start, p_);
        auto v = std::make_shared<Value>();
        if (!is_float) {
            errno = 0;
            char* e = nullptr;
            if (tok[0] == '-') {
                long long x = strtoll(tok.c_str(), &e, 10);
                if (errno == 0) { v->kind = Value::Int; v->i = x; return v; }
            } else {
                unsigned long long x = strtoull(tok.c_str(), &e, 10);
                if (errno == 0) { v->kind = Value::UInt; v->u = x; return v; }
            }
        }
        v->kind = Value::Float;
        v->d = strtod(tok.c_str(), nullptr);
        return v;
    }
    ValuePtr array() {
        auto v = std::make_shared<Value>(); v->kind = Value::Array;
        p_++; ws();
        if (p_ < end_ && *p_ == ']') { p_++; return v; }
        for (;;) {
            v->arr.push_back(value());
            ws();
            if (p_ >= end_) fail("unterminated array");
            if (*p_ == ',') { p_++; continue; }
            if (*p_ == ']') { p_++; return v; }
            fail("expected ',' or ']'");
        }
    }
    ValuePtr object() {
        auto v = std::make_shared<Value>(); v->kind = Value::Object;
        p_++; ws();
        if (p_ < end_ && *p_ == '}') { p_++; return v; }
        for (;;) {
            ws();
            if (p_ >= end_ || *p_ != '"') fail("expected string key");
            std::string key = string();
            ws();
            if (p_ >= end_ || *p_ != ':') fail("expected ':'");
            p_++;
            ValuePtr val = value();
            v->obj.emplace_back(std::move(key), std::move(val));
            ws();
            if (p_ >= end_) fail("unterminated object");
            if (*p_ == ',') { p_++; continue; }
            if (*p_ == '}') { p_++; return v; }
            fail("expected ',' or '}'");
        }
    }
};

inline ValuePtr parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace jsonmin
