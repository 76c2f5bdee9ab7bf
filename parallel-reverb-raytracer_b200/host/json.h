// Minimal JSON document model + recursive-descent parser for the rayverb config and material
// files (the reference uses rapidjson, which is not available in this image).  Header-only.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rvbjson {

class Value {
public:
    enum Type { Null, Bool, Number, String, Array, Object };
    Type type = Null;
    bool boolean = false;
    double number = 0;
    bool is_integer = false;     // written without fraction / exponent and fits an int
    std::string string;
    std::vector<Value> array;
    std::vector<std::pair<std::string, Value>> members;   // insertion order, like rapidjson

    bool IsNull() const { return type == Null; }
    bool IsBool() const { return type == Bool; }
    bool IsNumber() const { return type == Number; }
    bool IsInt() const { return type == Number && is_integer; }
    bool IsString() const { return type == String; }
    bool IsArray() const { return type == Array; }
    bool IsObject() const { return type == Object; }
    bool GetBool() const { return boolean; }
    double GetDouble() const { return number; }
    int GetInt() const { return (int)number; }
    const std::string &GetString() const { return string; }
    size_t Size() const { return array.size(); }
    const Value &operator[](size_t i) const { return array[i]; }
    bool HasMember(const std::string &k) const { return find(k) != nullptr; }
    const Value &operator[](const std::string &k) const
    {
        const Value *v = find(k);
        if (!v) throw std::runtime_error("key " + k + " not found in config object");
        return *v;
    }
    const Value &operator[](const char *k) const { return (*this)[std::string(k)]; }
    const Value *find(const std::string &k) const
    {
        if (type != Object) return nullptr;
        for (const auto &m : members)
            if (m.first == k) return &m.second;
        return nullptr;
    }
};

// A parsed file: like rapidjson::Document it records a parse error instead of throwing.
class Document : public Value {
public:
    bool HasParseError() const { return !error.empty(); }
    const std::string &GetParseError() const { return error; }
    void Parse(const std::string &text);
    std::string error;
};

namespace detail {
struct Parser {
    const std::string &s;
    size_t i = 0;
    explicit Parser(const std::string &t) : s(t) {}
    [[noreturn]] void fail(const std::string &m) const { throw std::runtime_error(m + " at offset " + std::to_string(i)); }
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    Value value()
    {
        ws();
        if (i >= s.size()) fail("The document is empty");
        char c = s[i];
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') { Value v; v.type = Value::String; v.string = str(); return v; }
        if (c == 't' || c == 'f' || c == 'n') return literal();
        return number();
    }
    Value literal()
    {
        Value v;
        if (s.compare(i, 4, "true") == 0) { v.type = Value::Bool; v.boolean = true; i += 4; }
        else if (s.compare(i, 5, "false") == 0) { v.type = Value::Bool; v.boolean = false; i += 5; }
        else if (s.compare(i, 4, "null") == 0) { v.type = Value::Null; i += 4; }
        else fail("Invalid value");
        return v;
    }
    Value number()
    {
        size_t b = i;
        if (i < s.size() && (s[i] == '-' || s[i] == '+')) ++i;
        bool digits = false, integral = true;
        while (i < s.size() && std::isdigit((unsigned char)s[i])) { ++i; digits = true; }
        if (i < s.size() && s[i] == '.') { integral = false; ++i; while (i < s.size() && std::isdigit((unsigned char)s[i])) { ++i; digits = true; } }
        if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
            integral = false; ++i;
            if (i < s.size() && (s[i] == '-' || s[i] == '+')) ++i;
            while (i < s.size() && std::isdigit((unsigned char)s[i])) ++i;
        }
        if (!digits) fail("Invalid value");
        Value v;
        v.type = Value::Number;
        v.number = std::strtod(s.substr(b, i - b).c_str(), nullptr);
        v.is_integer = integral && v.number >= -2147483648.0 && v.number <= 2147483647.0;
        return v;
    }
    std::string str()
    {
        ++i;   // opening quote
        std::string out;
        while (i < s.size() && s[i] != '"') {
            char c = s[i++];
            if (c == '\\') {
                if (i >= s.size()) fail("Invalid escape character in string");
                char e = s[i++];
                switch (e) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': {
                    if (i + 4 > s.size()) fail("Incorrect hex digit after \\u escape in string");
                    unsigned cp = (unsigned)std::strtoul(s.substr(i, 4).c_str(), nullptr, 16);
                    i += 4;
                    if (cp < 0x80) out += (char)cp;
                    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                    else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: out += e;
                }
            } else out += c;
        }
        if (i >= s.size()) fail("Missing a closing quotation mark in string");
        ++i;
        return out;
    }
    Value array()
    {
        Value v; v.type = Value::Array;
        ++i;
        if (eat(']')) return v;
        for (;;) {
            v.array.push_back(value());
            if (eat(',')) continue;
            if (eat(']')) return v;
            fail("Missing a comma or ']' after an array element");
        }
    }
    Value object()
    {
        Value v; v.type = Value::Object;
        ++i;
        if (eat('}')) return v;
        for (;;) {
            ws();
            if (i >= s.size() || s[i] != '"') fail("Missing a name for object member");
            std::string k = str();
            if (!eat(':')) fail("Missing a colon after a name of object member");
            v.members.emplace_back(std::move(k), value());
            if (eat(',')) continue;
            if (eat('}')) return v;
            fail("Missing a comma or '}' after an object member");
        }
    }
};
}  // namespace detail

inline void Document::Parse(const std::string &text)
{
    error.clear();
    try {
        detail::Parser p(text);
        Value v = p.value();
        p.ws();
        if (p.i != text.size()) p.fail("The document root must not be followed by other values");
        static_cast<Value &>(*this) = std::move(v);
    } catch (const std::runtime_error &e) {
        static_cast<Value &>(*this) = Value();
        error = e.what();
    }
}

}  // namespace rvbjson
