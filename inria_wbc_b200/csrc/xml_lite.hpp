// xml_lite.hpp -- minimal XML element tree (elements + attributes; text, comments, PIs skipped).
// Enough for URDF (<robot><link><inertial>..</link><joint>..</joint></robot>) and the SRDF
// <group_state><joint name= value=/> lists of etc/*/configurations.srdf.  Replaces urdfdom /
// pinocchio::urdf / pinocchio::srdf, which are not available here (SURVEY.md finding 2).
#pragma once
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace iwbc {
namespace xml {
    struct Element {
        std::string name;
        std::vector<std::pair<std::string, std::string>> attrs;
        std::vector<Element> children;

        const std::string* attr(const std::string& k) const
        {
            for (auto& a : attrs)
                if (a.first == k)
                    return &a.second;
            return nullptr;
        }
        std::string attr_or(const std::string& k, const std::string& dflt) const
        {
            auto* a = attr(k);
            return a ? *a : dflt;
        }
        const Element* child(const std::string& n) const
        {
            for (auto& c : children)
                if (c.name == n)
                    return &c;
            return nullptr;
        }
    };

    class Parser {
    public:
        explicit Parser(const std::string& s) : s_(s) {}
        Element root()
        {
            skip_misc();
            Element e;
            if (!element(e))
                throw std::runtime_error("xml: no root element");
            return e;
        }

    private:
        const std::string& s_;
        size_t p_ = 0;
        void ws()
        {
            while (p_ < s_.size() && isspace((unsigned char)s_[p_]))
                ++p_;
        }
        bool starts(const char* t) const { return s_.compare(p_, strlen(t), t) == 0; }
        void skip_misc()
        {
            for (;;) {
                // skip text up to next '<'
                while (p_ < s_.size() && s_[p_] != '<')
                    ++p_;
                if (p_ >= s_.size())
                    return;
                if (starts("<!--")) {
                    size_t e = s_.find("-->", p_);
                    if (e == std::string::npos)
                        throw std::runtime_error("xml: unterminated comment");
                    p_ = e + 3;
                }
                else if (starts("<?")) {
                    size_t e = s_.find("?>", p_);
                    if (e == std::string::npos)
                        throw std::runtime_error("xml: unterminated processing instruction");
                    p_ = e + 2;
                }
                else if (starts("<!")) {
                    size_t e = s_.find(">", p_);
                    p_ = e + 1;
                }
                else
                    return;
            }
        }
        std::string ident()
        {
            size_t a = p_;
            while (p_ < s_.size() && (isalnum((unsigned char)s_[p_]) || s_[p_] == '_' || s_[p_] == '-' || s_[p_] == ':' || s_[p_] == '.'))
                ++p_;
            return s_.substr(a, p_ - a);
        }
        bool element(Element& e)
        {
            if (p_ >= s_.size() || s_[p_] != '<' || starts("</"))
                return false;
            ++p_;
            e.name = ident();
            for (;;) {
                ws();
                if (p_ >= s_.size())
                    throw std::runtime_error("xml: unterminated tag <" + e.name);
                if (starts("/>")) {
                    p_ += 2;
                    return true;
                }
                if (s_[p_] == '>') {
                    ++p_;
                    break;
                }
                std::string k = ident();
                ws();
                if (k.empty() || p_ >= s_.size() || s_[p_] != '=')
                    throw std::runtime_error("xml: bad attribute in <" + e.name + ">");
                ++p_;
                ws();
                char q = s_[p_];
                if (q != '"' && q != '\'')
                    throw std::runtime_error("xml: unquoted attribute in <" + e.name + ">");
                size_t end = s_.find(q, p_ + 1);
                if (end == std::string::npos)
                    throw std::runtime_error("xml: unterminated attribute in <" + e.name + ">");
                e.attrs.emplace_back(k, s_.substr(p_ + 1, end - p_ - 1));
                p_ = end + 1;
            }
            for (;;) {
                skip_misc();
                if (p_ >= s_.size())
                    throw std::runtime_error("xml: missing </" + e.name + ">");
                if (starts("</")) {
                    p_ += 2;
                    std::string n = ident();
                    if (n != e.name)
                        throw std::runtime_error("xml: mismatched </" + n + "> for <" + e.name + ">");
                    ws();
                    if (p_ < s_.size() && s_[p_] == '>')
                        ++p_;
                    return true;
                }
                Element c;
                if (element(c))
                    e.children.push_back(std::move(c));
            }
        }
    };

    inline Element LoadFile(const std::string& path)
    {
        std::ifstream f(path);
        if (!f)
            throw std::runtime_error("xml: cannot open '" + path + "'");
        std::stringstream ss;
        ss << f.rdbuf();
        std::string text = ss.str();
        Parser p(text);
        return p.root();
    }
} // namespace xml
} // namespace iwbc
