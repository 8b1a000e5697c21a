// json.hpp -- minimal JSON reader for the per-rank partition files written by the reference pre-processor
// (01-Pre_Process/Core/Outputs.py:29-51, json.dumps indent=4; schema: SURVEY.md App. D).  Objects keep their
// keys in a std::map so that iteration is in ascending key order; numeric-tag maps are re-sorted numerically by
// the driver exactly like the reference's std::map<unsigned int, ...> containers.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace svlhost {

struct JValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    double num = 0.0;
    bool b = false;
    std::string str;
    std::vector<JValue> arr;
    std::map<std::string, JValue> obj;

    bool has(const std::string &k) const { return kind == Object && obj.count(k); }
    const JValue &operator[](const std::string &k) const {
        static const JValue null_value;
        if (kind != Object) return null_value;
        auto it = obj.find(k);
        return it == obj.end() ? null_value : it->second;
    }
    const JValue &operator[](size_t i) const { return arr.at(i); }
    size_t size() const { return kind == Array ? arr.size() : kind == Object ? obj.size() : 0; }
    double as_double(double dflt = 0.0) const { return kind == Number ? num : (kind == Bool ? (b ? 1.0 : 0.0) : dflt); }
    int as_int(int dflt = 0) const { return kind == Number ? (int)num : dflt; }
    std::string as_string(const std::string &dflt = "") const { return kind == String ? str : dflt; }
};

class JParser {
  public:
    explicit JParser(const std::string &text) : s(text) {}
    JValue parse() {
        JValue v = value();
        ws();
        if (p != s.size()) fail("trailing characters");
        return v;
    }

  private:
    const std::string &s;
    size_t p = 0;
    [[noreturn]] void fail(const char *msg) { throw std::runtime_error(std::string("JSON: ") + msg + " at offset " + std::to_string(p)); }
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++; }
    JValue value() {
        ws();
        if (p >= s.size()) fail("unexpected end");
        const char c = s[p];
        JValue v;
        if (c == '{') {
            v.kind = JValue::Object;
            p++; ws();
            if (s[p] == '}') { p++; return v; }
            for (;;) {
                ws();
                std::string k = string();
                ws();
                if (s[p++] != ':') fail("':' expected");
                v.obj.emplace(std::move(k), value());
                ws();
                if (s[p] == ',') { p++; continue; }
                if (s[p] == '}') { p++; break; }
                fail("',' or '}' expected");
            }
        } else if (c == '[') {
            v.kind = JValue::Array;
            p++; ws();
            if (s[p] == ']') { p++; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (s[p] == ',') { p++; continue; }
                if (s[p] == ']') { p++; break; }
                fail("',' or ']' expected");
            }
        } else if (c == '"') {
            v.kind = JValue::String;
            v.str = string();
        } else if (s.compare(p, 4, "true") == 0) { v.kind = JValue::Bool; v.b = true; p += 4; }
        else if (s.compare(p, 5, "false") == 0) { v.kind = JValue::Bool; v.b = false; p += 5; }
        else if (s.compare(p, 4, "null") == 0) { p += 4; }
        else {
            char *end = nullptr;
            v.num = std::strtod(s.c_str() + p, &end);
            if (end == s.c_str() + p) fail("value expected");
            v.kind = JValue::Number;
            p = (size_t)(end - s.c_str());
        }
        return v;
    }
    std::string string() {
        if (s[p] != '"') fail("string expected");
        p++;
        std::string out;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\') {
                p++;
                switch (s[p]) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': out += '?'; p += 4; break;
                default: out += s[p];
                }
                p++;
            } else out += s[p++];
        }
        if (p >= s.size()) fail("unterminated string");
        p++;
        return out;
    }
};

}  // namespace svlhost
