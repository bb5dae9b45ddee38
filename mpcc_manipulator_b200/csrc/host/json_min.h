// Minimal JSON reader for the reference's Params/*.json schema (flat objects of numbers / bools /
// strings and arrays of numbers).  The reference uses nlohmann/json, which this repo does not ship.
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mpcc {
namespace json {

struct Value {
    enum Kind { Null, Number, Bool, String, Array, Object } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> obj;

    bool has(const std::string& k) const { return kind == Object && obj.count(k); }
    const Value& at(const std::string& k) const {
        auto it = obj.find(k);
        if (kind != Object || it == obj.end()) throw std::runtime_error("json: missing key '" + k + "'");
        return it->second;
    }
    double number(const std::string& k) const {
        const Value& v = at(k);
        if (v.kind == Number) return v.num;
        if (v.kind == Bool) return v.b ? 1.0 : 0.0;
        throw std::runtime_error("json: key '" + k + "' is not a number");
    }
    std::vector<double> numbers(const std::string& k) const {
        const Value& v = at(k);
        if (v.kind != Array) throw std::runtime_error("json: key '" + k + "' is not an array");
        std::vector<double> out;
        out.reserve(v.arr.size());
        for (const Value& e : v.arr) {
            if (e.kind != Number) throw std::runtime_error("json: array '" + k + "' holds a non-number");
            out.push_back(e.num);
        }
        return out;
    }
    std::string string(const std::string& k) const {
        const Value& v = at(k);
        if (v.kind != String) throw std::runtime_error("json: key '" + k + "' is not a string");
        return v.str;
    }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s) {}
    Value parse() {
        Value v = value();
        ws();
        if (i_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t i_ = 0;
    [[noreturn]] void fail(const std::string& m) const { throw std::runtime_error("json: " + m + " at offset " + std::to_string(i_)); }
    void ws() { while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) i_++; }
    bool eat(char c) { ws(); if (i_ < s_.size() && s_[i_] == c) { i_++; return true; } return false; }
    Value value() {
        ws();
        if (i_ >= s_.size()) fail("unexpected end");
        char c = s_[i_];
        Value v;
        if (c == '{') {
            i_++;
            v.kind = Value::Object;
            if (eat('}')) return v;
            do {
                ws();
                Value k = string_();
                if (!eat(':')) fail("expected ':'");
                v.obj[k.str] = value();
            } while (eat(','));
            if (!eat('}')) fail("expected '}'");
        } else if (c == '[') {
            i_++;
            v.kind = Value::Array;
            if (eat(']')) return v;
            do { v.arr.push_back(value()); } while (eat(','));
            if (!eat(']')) fail("expected ']'");
        } else if (c == '"') {
            v = string_();
        } else if (s_.compare(i_, 4, "true") == 0) { v.kind = Value::Bool; v.b = true; i_ += 4; }
        else if (s_.compare(i_, 5, "false") == 0) { v.kind = Value::Bool; v.b = false; i_ += 5; }
        else if (s_.compare(i_, 4, "null") == 0) { v.kind = Value::Null; i_ += 4; }
        else {
            const char* b = s_.c_str() + i_;
            char* e = nullptr;
            v.num = std::strtod(b, &e);
            if (e == b) fail("bad token");
            v.kind = Value::Number;
            i_ += (size_t)(e - b);
        }
        return v;
    }
    Value string_() {
        if (i_ >= s_.size() || s_[i_] != '"') fail("expected string");
        i_++;
        Value v;
        v.kind = Value::String;
        while (i_ < s_.size() && s_[i_] != '"') {
            if (s_[i_] == '\\' && i_ + 1 < s_.size()) {
                char n = s_[i_ + 1];
                v.str += (n == 'n') ? '\n' : (n == 't') ? '\t' : n;
                i_ += 2;
            } else v.str += s_[i_++];
        }
        if (i_ >= s_.size()) fail("unterminated string");
        i_++;
        return v;
    }
};

inline Value parse_file(const std::string& path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("json: cannot open '" + path + "'");
    std::stringstream ss;
    ss << f.rdbuf();
    std::string s = ss.str();
    try {
        return Parser(s).parse();
    } catch (const std::exception& e) {
        throw std::runtime_error(std::string(e.what()) + " in '" + path + "'");
    }
}

}  // namespace json
}  // namespace mpcc
