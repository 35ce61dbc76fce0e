// mon_json.h — minimal JSON reader (objects, arrays, strings, numbers, bools, null; // and /* */
// comments are skipped, like the reference's json::parse(file, nullptr, true, /*ignore_comments=*/true),
// MON/Core/src/nerf_model.cu:1281).  Only what the network configuration (base.json) needs.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace monjson {

struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> obj;

    const Value* find(const std::string& key) const {
        if (type != Object) return nullptr;
        auto it = obj.find(key);
        return it == obj.end() ? nullptr : &it->second;
    }
    double number_or(const std::string& key, double dflt) const {
        const Value* v = find(key);
        return (v && v->type == Number) ? v->num : dflt;
    }
    std::string string_or(const std::string& key, const std::string& dflt) const {
        const Value* v = find(key);
        return (v && v->type == String) ? v->str : dflt;
    }
};

class Parser {
public:
    explicit Parser(const std::string& text) : s_(text) {}
    bool parse(Value& out, std::string& err) {
        try {
            skip();
            out = value();
            skip();
            if (p_ != s_.size()) throw std::string("trailing characters");
            return true;
        } catch (const std::string& e) {
            err = e + " at byte " + std::to_string(p_);
            return false;
        }
    }

private:
    const std::string& s_;
    size_t p_ = 0;

    void skip() {
        for (;;) {
            while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_;
            if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '/') {
                while (p_ < s_.size() && s_[p_] != '\n') ++p_;
            } else if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '*') {
                p_ += 2;
                while (p_ + 1 < s_.size() && !(s_[p_] == '*' && s_[p_ + 1] == '/')) ++p_;
                if (p_ + 1 >= s_.size()) throw std::string("unterminated comment");
                p_ += 2;
            } else {
                return;
            }
        }
    }
    char peek() const { return p_ < s_.size() ? s_[p_] : '\0'; }
    void expect(char c) {
        if (peek() != c) throw std::string("expected '") + c + "'";
        ++p_;
    }
    Value value() {
        skip();
        const char c = peek();
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') { Value v; v.type = Value::String; v.str = string(); return v; }
        if (s_.compare(p_, 4, "true") == 0) { p_ += 4; Value v; v.type = Value::Bool; v.b = true; return v; }
        if (s_.compare(p_, 5, "false") == 0) { p_ += 5; Value v; v.type = Value::Bool; v.b = false; return v; }
        if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return Value(); }
        return number();
    }
    Value number() {
        const char* start = s_.c_str() + p_;
        char* end = nullptr;
        const double d = std::strtod(start, &end);
        if (end == start) throw std::string("invalid value");
        p_ += (size_t)(end - start);
        Value v; v.type = Value::Number; v.num = d;
        return v;
    }
    std::string string() {
        expect('"');
        std::string out;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if (c == '\\') {
                if (p_ >= s_.size()) throw std::string("bad escape");
                const char e = s_[p_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': p_ += 4; out += '?'; break;  // non-ASCII never appears in the keys we read
                    default: out += e;
                }
            } else {
                out += c;
            }
        }
        expect('"');
        return out;
    }
    Value array() {
        Value v; v.type = Value::Array;
        expect('[');
        skip();
        if (peek() == ']') { ++p_; return v; }
        for (;;) {
            v.arr.push_back(value());
            skip();
            if (peek() == ',') { ++p_; continue; }
            expect(']');
            return v;
        }
    }
    Value object() {
        Value v; v.type = Value::Object;
        expect('{');
        skip();
        if (peek() == '}') { ++p_; return v; }
        for (;;) {
            skip();
            std::string k = string();
            skip();
            expect(':');
            v.obj[k] = value();
            skip();
            if (peek() == ',') { ++p_; continue; }
            expect('}');
            return v;
        }
    }
};

}  // namespace monjson
