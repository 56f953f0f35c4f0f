// Jet::TensorNetworkSerializer / TensorNetworkFile — drop-in for
// /root/reference/include/jet/TensorNetworkIO.hpp.  Same file format:
//   {"path": [[i, j], ...], "tensors": [[tags, indices, shape, [[re, im], ...]], ...]}
// The reference parses with the vendored nlohmann/json; this header carries its own small JSON
// reader/writer (objects are written with sorted keys and numbers in shortest round-trip form, so
// load -> dump(-1) reproduces the reference's strings).
#pragma once

#include <charconv>
#include <complex>
#include <cstdint>
#include <exception>
#include <map>
#include <optional>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "Abort.hpp"
#include "PathInfo.hpp"
#include "Tensor.hpp"
#include "TensorNetwork.hpp"

namespace Jet {

/// Thrown for malformed JSON text (the role nlohmann's json::exception plays in the reference).
class JsonException : public std::exception {
  public:
    explicit JsonException(std::string msg) : msg_(std::move(msg)) {}
    const char *what() const noexcept override { return msg_.c_str(); }

  private:
    std::string msg_;
};

/// Thrown for valid JSON that is not a valid tensor network file.
class TensorFileException : public Exception {
  public:
    explicit TensorFileException(const std::string &what_arg)
        : Exception("Error parsing tensor network file: " + what_arg)
    {
    }
    explicit TensorFileException(const char *what_arg) : TensorFileException(std::string(what_arg)) {}
};

namespace JsonLite {

struct Value {
    enum class Kind { Null, Bool, Int, Float, String, Array, Object } kind = Kind::Null;
    bool b = false;
    int64_t i = 0;
    double d = 0;
    std::string s;
    std::vector<Value> a;
    std::map<std::string, Value> o;

    bool is_number() const { return kind == Kind::Int || kind == Kind::Float; }
    double number() const { return kind == Kind::Int ? static_cast<double>(i) : d; }
};

class Parser {
  public:
    explicit Parser(const std::string &text) : t_(text) {}
    Value Parse()
    {
        Skip_();
        if (p_ >= t_.size())
            throw JsonException("parse error: unexpected end of input");
        Value v = ParseValue_();
        Skip_();
        if (p_ != t_.size())
            throw JsonException("parse error: trailing characters at offset " + std::to_string(p_));
        return v;
    }

  private:
    const std::string &t_;
    size_t p_ = 0;

    void Skip_()
    {
        while (p_ < t_.size() && (t_[p_] == ' ' || t_[p_] == '\n' || t_[p_] == '\t' || t_[p_] == '\r'))
            p_++;
    }
    [[noreturn]] void Fail_(const std::string &what) const
    {
        throw JsonException("parse error at offset " + std::to_string(p_) + ": " + what);
    }
    Value ParseValue_()
    {
        Skip_();
        if (p_ >= t_.size())
            Fail_("unexpected end of input");
        const char c = t_[p_];
        if (c == '{')
            return ParseObject_();
        if (c == '[')
            return ParseArray_();
        if (c == '"') {
            Value v;
            v.kind = Value::Kind::String;
            v.s = ParseString_();
            return v;
        }
        if (t_.compare(p_, 4, "true") == 0) {
            p_ += 4;
            Value v;
            v.kind = Value::Kind::Bool;
            v.b = true;
            return v;
        }
        if (t_.compare(p_, 5, "false") == 0) {
            p_ += 5;
            Value v;
            v.kind = Value::Kind::Bool;
            return v;
        }
        if (t_.compare(p_, 4, "null") == 0) {
            p_ += 4;
            return Value();
        }
        return ParseNumber_();
    }
    Value ParseNumber_()
    {
        const size_t start = p_;
        bool is_float = false;
        if (p_ < t_.size() && t_[p_] == '-')
            p_++;
        while (p_ < t_.size()) {
            const char c = t_[p_];
            if (c >= '0' && c <= '9')
                p_++;
            else if (c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-') {
                is_float = true;
                p_++;
            }
            else
                break;
        }
        if (p_ == start)
            Fail_("invalid literal");
        Value v;
        const char *b = t_.data() + start, *e = t_.data() + p_;
        if (!is_float) {
            v.kind = Value::Kind::Int;
            const auto r = std::from_chars(b, e, v.i);
            if (r.ec == std::errc() && r.ptr == e)
                return v;
        }
        v.kind = Value::Kind::Float;
        const auto r = std::from_chars(b, e, v.d);
        if (r.ec != std::errc() || r.ptr != e)
            Fail_("invalid number");
        return v;
    }
    std::string ParseString_()
    {
        std::string out;
        p_++; // opening quote
        while (true) {
            if (p_ >= t_.size())
                Fail_("unterminated string");
            const char c = t_[p_++];
            if (c == '"')
                break;
            if (c != '\\') {
                out += c;
                continue;
            }
            if (p_ >= t_.size())
                Fail_("unterminated escape");
            const char e = t_[p_++];
            switch (e) {
            case '"': out += '"'; break;
            case '\\': out += '\\'; break;
            case '/': out += '/'; break;
            case 'b': out += '\b'; break;
            case 'f': out += '\f'; break;
            case 'n': out += '\n'; break;
            case 'r': out += '\r'; break;
            case 't': out += '\t'; break;
            case 'u': {
                if (p_ + 4 > t_.size())
                    Fail_("bad unicode escape");
                unsigned cp = std::stoul(t_.substr(p_, 4), nullptr, 16);
                p_ += 4;
                if (cp < 0x80)
                    out += static_cast<char>(cp);
                else if (cp < 0x800) {
                    out += static_cast<char>(0xC0 | (cp >> 6));
                    out += static_cast<char>(0x80 | (cp & 0x3F));
                }
                else {
                    out += static_cast<char>(0xE0 | (cp >> 12));
                    out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
                    out += static_cast<char>(0x80 | (cp & 0x3F));
                }
                break;
            }
            default: Fail_("bad escape");
            }
        }
        return out;
    }
    Value ParseArray_()
    {
        Value v;
        v.kind = Value::Kind::Array;
        p_++;
        Skip_();
        if (p_ < t_.size() && t_[p_] == ']') {
            p_++;
            return v;
        }
        while (true) {
            v.a.push_back(ParseValue_());
            Skip_();
            if (p_ >= t_.size())
                Fail_("unterminated array");
            if (t_[p_] == ',') {
                p_++;
                continue;
            }
            if (t_[p_] == ']') {
                p_++;
                return v;
            }
            Fail_("expected ',' or ']'");
        }
    }
    Value ParseObject_()
    {
        Value v;
        v.kind = Value::Kind::Object;
        p_++;
        Skip_();
        if (p_ < t_.size() && t_[p_] == '}') {
            p_++;
            return v;
        }
        while (true) {
            Skip_();
            if (p_ >= t_.size() || t_[p_] != '"')
                Fail_("expected string key");
            std::string key = ParseString_();
            Skip_();
            if (p_ >= t_.size() || t_[p_] != ':')
                Fail_("expected ':'");
            p_++;
            v.o[key] = ParseValue_();
            Skip_();
            if (p_ >= t_.size())
                Fail_("unterminated object");
            if (t_[p_] == ',') {
                p_++;
                continue;
            }
            if (t_[p_] == '}') {
                p_++;
                return v;
            }
            Fail_("expected ',' or '}'");
        }
    }
};

inline void DumpString(const std::string &s, std::string &out)
{
    out += '"';
    for (const char c : s) {
        switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\t': out += "\\t"; break;
        case '\r': out += "\\r"; break;
        default: out += c;
        }
    }
    out += '"';
}

inline void DumpDouble(double d, std::string &out)
{
    char buf[40];
    const auto r = std::to_chars(buf, buf + sizeof(buf), d);
    std::string s(buf, r.ptr);
    if (s.find_first_of(".eEn") == std::string::npos)
        s += ".0";
    out += s;
}

inline void Dump(const Value &v, int indent, int depth, std::string &out)
{
    const bool pretty = indent >= 0;
    auto newline = [&](int dpt) {
        if (pretty) {
            out += '\n';
            out.append(static_cast<size_t>(indent * dpt), ' ');
        }
    };
    switch (v.kind) {
    case Value::Kind::Null: out += "null"; break;
    case Value::Kind::Bool: out += v.b ? "true" : "false"; break;
    case Value::Kind::Int: out += std::to_string(v.i); break;
    case Value::Kind::Float: DumpDouble(v.d, out); break;
    case Value::Kind::String: DumpString(v.s, out); break;
    case Value::Kind::Array:
        if (v.a.empty()) {
            out += "[]";
            break;
        }
        out += '[';
        for (size_t i = 0; i < v.a.size(); i++) {
            if (i)
                out += ',';
            newline(depth + 1);
            Dump(v.a[i], indent, depth + 1, out);
        }
        newline(depth);
        out += ']';
        break;
    case Value::Kind::Object:
        if (v.o.empty()) {
            out += "{}";
            break;
        }
        out += '{';
        {
            bool first = true;
            for (const auto &[k, e] : v.o) {
                if (!first)
                    out += ',';
                first = false;
                newline(depth + 1);
                DumpString(k, out);
                out += pretty ? ": " : ":";
                Dump(e, indent, depth + 1, out);
            }
        }
        newline(depth);
        out += '}';
        break;
    }
}

} // namespace JsonLite

/// A tensor network and (optionally) a contraction path, as stored in a file.
template <class TensorType> struct TensorNetworkFile {
    std::optional<PathInfo> path;
    TensorNetwork<TensorType> tensors;
};

template <class TensorType> class TensorNetworkSerializer {
  public:
    TensorNetworkSerializer(int indent = -1) : indent_(indent) {}

    /// Network + path -> JSON text.
    std::string operator()(const TensorNetwork<TensorType> &tn, const PathInfo &path)
    {
        JsonLite::Value root;
        root.kind = JsonLite::Value::Kind::Object;
        JsonLite::Value p;
        p.kind = JsonLite::Value::Kind::Array;
        for (const auto &[a, b] : path.GetPath()) {
            JsonLite::Value pair;
            pair.kind = JsonLite::Value::Kind::Array;
            pair.a.push_back(Int_(a));
            pair.a.push_back(Int_(b));
            p.a.push_back(std::move(pair));
        }
        root.o["path"] = std::move(p);
        return Dump_(tn, std::move(root));
    }

    /// Network -> JSON text.
    std::string operator()(const TensorNetwork<TensorType> &tn)
    {
        JsonLite::Value root;
        root.kind = JsonLite::Value::Kind::Object;
        return Dump_(tn, std::move(root));
    }

    /// JSON text -> network (+ path).  `col_major` reverses indices and shapes on load (the
    /// reference's switch for column-major back ends).
    TensorNetworkFile<TensorType> operator()(std::string js_str, bool col_major = false)
    {
        using S = typename TensorType::scalar_type_t;
        const JsonLite::Value root = JsonLite::Parser(js_str).Parse();
        if (root.kind != JsonLite::Value::Kind::Object)
            throw TensorFileException("root element must be an object.");
        const auto tensors = root.o.find("tensors");
        if (tensors == root.o.end())
            throw TensorFileException("root object must contain 'tensors' key");
        if (tensors->second.kind != JsonLite::Value::Kind::Array)
            throw TensorFileException("'tensors' must be an array.");

        TensorNetworkFile<TensorType> file;
        size_t t = 0;
        for (const auto &entry : tensors->second.a) {
            if (entry.kind != JsonLite::Value::Kind::Array || entry.a.size() != 4)
                throw TensorFileException("tensor " + std::to_string(t) +
                                          " must be an array of tags, indices, shape and data.");
            std::vector<std::string> tags = Strings_(entry.a[0], t);
            std::vector<std::string> indices = Strings_(entry.a[1], t);
            std::vector<size_t> shape;
            for (const auto &v : entry.a[2].a) {
                if (!v.is_number())
                    throw TensorFileException("tensor " + std::to_string(t) + " has an invalid shape.");
                shape.push_back(static_cast<size_t>(v.number()));
            }
            std::vector<S> data(entry.a[3].a.size());
            for (size_t i = 0; i < data.size(); i++) {
                const auto &z = entry.a[3].a[i];
                if (z.kind != JsonLite::Value::Kind::Array || z.a.size() < 2 || !z.a[0].is_number() ||
                    !z.a[1].is_number()) {
                    std::string txt;
                    JsonLite::Dump(z, -1, 0, txt);
                    throw TensorFileException("Invalid element at index " + std::to_string(i) +
                                              " of tensor " + std::to_string(t) + ": Could not parse " +
                                              txt + " as complex.");
                }
                using R = typename S::value_type;
                data[i] = S{static_cast<R>(z.a[0].number()), static_cast<R>(z.a[1].number())};
            }
            if (col_major) {
                std::reverse(indices.begin(), indices.end());
                std::reverse(shape.begin(), shape.end());
            }
            file.tensors.AddTensor(TensorType(indices, shape, data), tags);
            t++;
        }
        const auto path = root.o.find("path");
        if (path != root.o.end()) {
            PathInfo::Path p;
            for (const auto &pair : path->second.a) {
                if (pair.kind != JsonLite::Value::Kind::Array || pair.a.size() != 2 ||
                    !pair.a[0].is_number() || !pair.a[1].is_number())
                    throw TensorFileException("path entries must be pairs of node ids.");
                p.emplace_back(static_cast<size_t>(pair.a[0].number()),
                               static_cast<size_t>(pair.a[1].number()));
            }
            file.path = PathInfo(file.tensors, p);
        }
        return file;
    }

  private:
    int indent_;

    static JsonLite::Value Int_(size_t v)
    {
        JsonLite::Value x;
        x.kind = JsonLite::Value::Kind::Int;
        x.i = static_cast<int64_t>(v);
        return x;
    }
    static JsonLite::Value Str_(const std::string &s)
    {
        JsonLite::Value x;
        x.kind = JsonLite::Value::Kind::String;
        x.s = s;
        return x;
    }
    static std::vector<std::string> Strings_(const JsonLite::Value &v, size_t t)
    {
        std::vector<std::string> out;
        if (v.kind != JsonLite::Value::Kind::Array)
            throw TensorFileException("tensor " + std::to_string(t) + " has a malformed string list.");
        for (const auto &e : v.a) {
            if (e.kind != JsonLite::Value::Kind::String)
                throw TensorFileException("tensor " + std::to_string(t) + " has a malformed string list.");
            out.push_back(e.s);
        }
        return out;
    }

    std::string Dump_(const TensorNetwork<TensorType> &tn, JsonLite::Value root)
    {
        using K = JsonLite::Value::Kind;
        JsonLite::Value list;
        list.kind = K::Array;
        for (const auto &node : tn.GetNodes()) {
            JsonLite::Value entry, tags, indices, shape, data;
            entry.kind = tags.kind = indices.kind = shape.kind = data.kind = K::Array;
            for (const auto &s : node.tags)
                tags.a.push_back(Str_(s));
            for (const auto &s : node.tensor.GetIndices())
                indices.a.push_back(Str_(s));
            for (const auto s : node.tensor.GetShape())
                shape.a.push_back(Int_(s));
            for (const auto &z : node.tensor.GetData()) {
                JsonLite::Value c, re, im;
                c.kind = K::Array;
                re.kind = im.kind = K::Float;
                re.d = static_cast<double>(z.real());
                im.d = static_cast<double>(z.imag());
                c.a.push_back(re);
                c.a.push_back(im);
                data.a.push_back(std::move(c));
            }
            entry.a.push_back(std::move(tags));
            entry.a.push_back(std::move(indices));
            entry.a.push_back(std::move(shape));
            entry.a.push_back(std::move(data));
            list.a.push_back(std::move(entry));
        }
        root.o["tensors"] = std::move(list);
        std::string out;
        JsonLite::Dump(root, indent_, 0, out);
        return out;
    }
};

} // namespace Jet
