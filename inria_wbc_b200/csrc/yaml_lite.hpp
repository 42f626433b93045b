// yaml_lite.hpp -- the subset of YAML the reference's etc/ files use, with a yaml-cpp-shaped API.
//
// The reference reads every config through yaml-cpp (`YAML::LoadFile`, `node["k"].as<T>()`,
// iteration in insertion order: /root/reference/src/controllers/pos_tracker.cpp:161-189).  yaml-cpp
// is not available here, so this header implements what those files need: block maps, block
// sequences, (multi-line, nested) flow maps / sequences, plain / quoted scalars and comments.
// Maps keep insertion order, because task order is the Hessian summation order.
#pragma once
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace iwbc {
namespace yaml {

    struct Error : std::runtime_error {
        explicit Error(const std::string& m) : std::runtime_error(m) {}
    };

    class Node {
    public:
        enum Kind { Null, Scalar, Seq, Map };
        Kind kind = Null;
        std::string scalar;
        bool quoted = false;
        std::vector<Node> seq;
        std::vector<std::pair<std::string, Node>> map;

        bool defined() const { return kind != Null; }
        explicit operator bool() const { return defined(); }
        bool is_map() const { return kind == Map; }
        bool is_seq() const { return kind == Seq; }
        bool is_scalar() const { return kind == Scalar; }
        size_t size() const { return kind == Seq ? seq.size() : kind == Map ? map.size() : 0; }

        // yaml-cpp semantics: a missing key gives an undefined node (tested with `if (c["frames"])`)
        const Node& operator[](const std::string& key) const
        {
            static const Node undefined;
            if (kind != Map)
                return undefined;
            for (auto& kv : map)
                if (kv.first == key)
                    return kv.second;
            return undefined;
        }
        Node& at(const std::string& key)
        {
            if (kind == Null)
                kind = Map;
            if (kind != Map)
                throw Error("yaml: not a map while setting '" + key + "'");
            for (auto& kv : map)
                if (kv.first == key)
                    return kv.second;
            map.emplace_back(key, Node());
            return map.back().second;
        }
        const Node& operator[](size_t i) const
        {
            if (kind != Seq || i >= seq.size())
                throw Error("yaml: bad sequence index");
            return seq[i];
        }
        void set(const std::string& key, const std::string& value)
        {
            Node& n = at(key);
            n = Node();
            n.kind = Scalar;
            n.scalar = value;
        }

        template <typename T>
        T as() const;

        std::string dump(int indent = 0) const
        {
            std::ostringstream o;
            std::string pad(indent, ' ');
            if (kind == Scalar)
                o << (quoted || scalar.empty() ? "\"" + scalar + "\"" : scalar);
            else if (kind == Seq) {
                o << "[";
                for (size_t i = 0; i < seq.size(); ++i)
                    o << (i ? ", " : "") << seq[i].dump(0);
                o << "]";
            }
            else if (kind == Map) {
                for (auto& kv : map) {
                    o << pad << kv.first << ":";
                    if (kv.second.kind == Map)
                        o << "\n" << kv.second.dump(indent + 2);
                    else
                        o << " " << kv.second.dump(0) << "\n";
                }
            }
            return o.str();
        }
    };

    namespace detail {
        inline std::string trim(const std::string& s)
        {
            size_t a = s.find_first_not_of(" \t\r\n");
            if (a == std::string::npos)
                return "";
            size_t b = s.find_last_not_of(" \t\r\n");
            return s.substr(a, b - a + 1);
        }

        // remove a trailing comment (a '#' outside quotes, at start or preceded by whitespace)
        inline std::string strip_comment(const std::string& s)
        {
            char q = 0;
            for (size_t i = 0; i < s.size(); ++i) {
                char c = s[i];
                if (q) {
                    if (c == q)
                        q = 0;
                }
                else if (c == '"' || c == '\'')
                    q = c;
                else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t'))
                    return s.substr(0, i);
            }
            return s;
        }

        struct Line {
            int indent;
            std::string text; // comment-stripped, trimmed
            int no;
        };

        inline Node make_scalar(const std::string& raw)
        {
            Node n;
            std::string t = trim(raw);
            if (t.empty() || t == "~" || t == "null")
                return n; // Null
            n.kind = Node::Scalar;
            if (t.size() >= 2 && ((t.front() == '"' && t.back() == '"') || (t.front() == '\'' && t.back() == '\''))) {
                n.scalar = t.substr(1, t.size() - 2);
                n.quoted = true;
            }
            else
                n.scalar = t;
            return n;
        }

        // flow collection parser over a character buffer
        struct Flow {
            const std::string& s;
            size_t p = 0;
            explicit Flow(const std::string& str) : s(str) {}
            void ws()
            {
                while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r'))
                    ++p;
            }
            std::string token(const char* stops)
            {
                ws();
                std::string out;
                if (p < s.size() && (s[p] == '"' || s[p] == '\'')) {
                    char q = s[p];
                    size_t e = s.find(q, p + 1);
                    if (e == std::string::npos)
                        throw Error("yaml: unterminated quote in flow collection");
                    out = s.substr(p, e - p + 1);
                    p = e + 1;
                    return out;
                }
                while (p < s.size() && !std::strchr(stops, s[p]))
                    out += s[p++];
                return trim(out);
            }
            Node value()
            {
                ws();
                if (p >= s.size())
                    return Node();
                if (s[p] == '[') {
                    ++p;
                    Node n;
                    n.kind = Node::Seq;
                    for (;;) {
                        ws();
                        if (p >= s.size())
                            throw Error("yaml: unterminated '['");
                        if (s[p] == ']') {
                            ++p;
                            break;
                        }
                        n.seq.push_back(value());
                        ws();
                        if (p < s.size() && s[p] == ',')
                            ++p;
                    }
                    return n;
                }
                if (s[p] == '{') {
                    ++p;
                    Node n;
                    n.kind = Node::Map;
                    for (;;) {
                        ws();
                        if (p >= s.size())
                            throw Error("yaml: unterminated '{'");
                        if (s[p] == '}') {
                            ++p;
                            break;
                        }
                        std::string key = token(":,}");
                        Node kn = make_scalar(key);
                        ws();
                        if (p < s.size() && s[p] == ':') {
                            ++p;
                            n.map.emplace_back(kn.scalar, value());
                        }
                        else
                            n.map.emplace_back(kn.scalar, Node());
                        ws();
                        if (p < s.size() && s[p] == ',')
                            ++p;
                    }
                    return n;
                }
                return make_scalar(token(",]}"));
            }
        };

        inline int flow_balance(const std::string& t)
        {
            int b = 0;
            char q = 0;
            for (char c : t) {
                if (q) {
                    if (c == q)
                        q = 0;
                }
                else if (c == '"' || c == '\'')
                    q = c;
                else if (c == '[' || c == '{')
                    ++b;
                else if (c == ']' || c == '}')
                    --b;
            }
            return b;
        }

        struct Parser {
            std::vector<Line> lines;
            size_t i = 0;

            explicit Parser(const std::string& text)
            {
                std::istringstream in(text);
                std::string l;
                int no = 0;
                while (std::getline(in, l)) {
                    ++no;
                    std::string t = strip_comment(l);
                    size_t a = t.find_first_not_of(" \t");
                    if (a == std::string::npos)
                        continue;
                    std::string body = trim(t);
                    if (body == "---" || body == "...")
                        continue;
                    lines.push_back({(int)a, body, no});
                }
                // join multi-line flow collections into one logical line
                std::vector<Line> joined;
                for (size_t k = 0; k < lines.size(); ++k) {
                    Line cur = lines[k];
                    int bal = flow_balance(cur.text);
                    while (bal > 0 && k + 1 < lines.size()) {
                        ++k;
                        cur.text += " " + lines[k].text;
                        bal = flow_balance(cur.text);
                    }
                    joined.push_back(cur);
                }
                lines.swap(joined);
            }

            // find "key:" split at top level of a block line; returns npos if the line is not "key: value"
            static size_t key_colon(const std::string& t)
            {
                char q = 0;
                for (size_t k = 0; k < t.size(); ++k) {
                    char c = t[k];
                    if (q) {
                        if (c == q)
                            q = 0;
                        continue;
                    }
                    if (c == '"' || c == '\'')
                        q = c;
                    else if (c == '[' || c == '{')
                        return std::string::npos;
                    else if (c == ':' && (k + 1 == t.size() || t[k + 1] == ' ' || t[k + 1] == '\t'))
                        return k;
                }
                return std::string::npos;
            }

            Node inline_value(const std::string& v)
            {
                std::string t = trim(v);
                if (!t.empty() && (t[0] == '[' || t[0] == '{')) {
                    Flow f(t);
                    return f.value();
                }
                return make_scalar(t);
            }

            Node block(int indent)
            {
                if (i >= lines.size() || lines[i].indent < indent)
                    return Node();
                int ind = lines[i].indent;
                const std::string& first = lines[i].text;
                if (first.size() >= 1 && first[0] == '-' && (first.size() == 1 || first[1] == ' ')) {
                    Node n;
                    n.kind = Node::Seq;
                    while (i < lines.size() && lines[i].indent == ind && lines[i].text[0] == '-'
                        && (lines[i].text.size() == 1 || lines[i].text[1] == ' ')) {
                        std::string rest = trim(lines[i].text.substr(1));
                        if (rest.empty()) {
                            ++i;
                            n.seq.push_back(block(ind + 1));
                        }
                        else if (key_colon(rest) != std::string::npos) {
                            // "- key: value" starts an inline map: re-interpret as a map indented past the dash
                            lines[i].text = rest;
                            lines[i].indent = ind + 2;
                            n.seq.push_back(block(ind + 2));
                        }
                        else {
                            ++i;
                            n.seq.push_back(inline_value(rest));
                        }
                    }
                    return n;
                }
                size_t kc = key_colon(first);
                if (kc == std::string::npos) {
                    Node n = inline_value(first);
                    ++i;
                    return n;
                }
                Node n;
                n.kind = Node::Map;
                while (i < lines.size() && lines[i].indent == ind) {
                    const std::string t = lines[i].text;
                    size_t c = key_colon(t);
                    if (c == std::string::npos)
                        throw Error("yaml: expected 'key: value' at line " + std::to_string(lines[i].no) + ": " + t);
                    Node kn = make_scalar(t.substr(0, c));
                    std::string rest = trim(t.substr(c + 1));
                    ++i;
                    if (!rest.empty())
                        n.map.emplace_back(kn.scalar, inline_value(rest));
                    else if (i < lines.size() && (lines[i].indent > ind || (lines[i].indent == ind && lines[i].text[0] == '-' && (lines[i].text.size() == 1 || lines[i].text[1] == ' '))))
                        n.map.emplace_back(kn.scalar, block(lines[i].indent));
                    else
                        n.map.emplace_back(kn.scalar, Node());
                }
                if (i < lines.size() && lines[i].indent > ind)
                    throw Error("yaml: bad indentation at line " + std::to_string(lines[i].no));
                return n;
            }
        };
    } // namespace detail

    inline Node Load(const std::string& text)
    {
        detail::Parser p(text);
        if (p.lines.empty())
            return Node();
        Node n = p.block(0);
        if (p.i < p.lines.size())
            throw Error("yaml: trailing content at line " + std::to_string(p.lines[p.i].no));
        return n;
    }

    inline Node LoadFile(const std::string& path)
    {
        std::ifstream f(path);
        if (!f)
            throw Error("yaml: cannot open '" + path + "'");
        std::stringstream ss;
        ss << f.rdbuf();
        return Load(ss.str());
    }

    template <>
    inline std::string Node::as<std::string>() const
    {
        if (kind == Null)
            return "";
        if (kind != Scalar)
            throw Error("yaml: not a scalar");
        return scalar;
    }
    template <>
    inline double Node::as<double>() const
    {
        if (kind != Scalar)
            throw Error("yaml: not a scalar (expected a number)");
        char* e = nullptr;
        double v = std::strtod(scalar.c_str(), &e);
        if (e == scalar.c_str() || *e != 0)
            throw Error("yaml: bad number '" + scalar + "'");
        return v;
    }
    template <>
    inline float Node::as<float>() const { return (float)as<double>(); }
    template <>
    inline int Node::as<int>() const { return (int)as<double>(); }
    template <>
    inline bool Node::as<bool>() const
    {
        if (kind != Scalar)
            throw Error("yaml: not a scalar (expected a bool)");
        if (scalar == "true" || scalar == "True" || scalar == "yes" || scalar == "on")
            return true;
        if (scalar == "false" || scalar == "False" || scalar == "no" || scalar == "off")
            return false;
        throw Error("yaml: bad bool '" + scalar + "'");
    }
    template <>
    inline std::vector<double> Node::as<std::vector<double>>() const
    {
        if (kind != Seq)
            throw Error("yaml: not a sequence");
        std::vector<double> v;
        for (auto& e : seq)
            v.push_back(e.as<double>());
        return v;
    }
    template <>
    inline std::vector<std::string> Node::as<std::vector<std::string>>() const
    {
        if (kind == Null)
            return {};
        if (kind != Seq)
            throw Error("yaml: not a sequence");
        std::vector<std::string> v;
        for (auto& e : seq)
            v.push_back(e.as<std::string>());
        return v;
    }

} // namespace yaml
} // namespace iwbc
