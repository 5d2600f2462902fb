/*
 * yaml-cpp/yaml.h -- stand-in for the sliver of yaml-cpp that SARAS's lib/io/parser.cc uses:
 * YAML::Load(std::istream&) / YAML::LoadFile, Node::operator[](key) on nested block mappings and
 * Node::as<T>() for int / float / double / bool / std::string, including '>' folded block scalars
 * (the "Probes" entry) and '#' comments.  It understands exactly the shape of input/parameters.yaml:
 * indentation-nested maps of "quoted or plain key": scalar.  With a real yaml-cpp installation this
 * directory is left off the include path.
 */
#ifndef SARAS_COMPAT_YAML_H
#define SARAS_COMPAT_YAML_H

#include <fstream>
#include <istream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace YAML {

class Exception : public std::runtime_error {
   public:
    explicit Exception(const std::string &m) : std::runtime_error(m) {}
};

class Node {
    struct Data {
        bool isMap = false;
        std::string scalar;
        std::map<std::string, std::shared_ptr<Data>> kids;
    };
    std::shared_ptr<Data> d_;
    std::string path_;

    static std::string trim(const std::string &s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return (a == std::string::npos) ? std::string() : s.substr(a, b - a + 1);
    }
    static std::string unquote(const std::string &s) {
        if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
        return s;
    }
    static std::string stripComment(const std::string &s) {
        bool inq = false;
        char q = 0;
        for (size_t i = 0; i < s.size(); i++) {
            if (inq) { if (s[i] == q) inq = false; }
            else if (s[i] == '"' || s[i] == '\'') { inq = true; q = s[i]; }
            else if (s[i] == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
        }
        return s;
    }
    static size_t keyColon(const std::string &s) {
        bool inq = false;
        char q = 0;
        for (size_t i = 0; i < s.size(); i++) {
            if (inq) { if (s[i] == q) inq = false; }
            else if (s[i] == '"' || s[i] == '\'') { inq = true; q = s[i]; }
            else if (s[i] == ':' && (i + 1 == s.size() || s[i + 1] == ' ' || s[i + 1] == '\t')) return i;
        }
        return std::string::npos;
    }

   public:
    Node() : d_(std::make_shared<Data>()) {}

    static Node parse(std::istream &in) {
        Node root;
        root.d_->isMap = true;
        struct Frame { int indent; std::shared_ptr<Data> node; };
        std::vector<Frame> stack;
        stack.push_back({-1, root.d_});
        std::string line;
        std::shared_ptr<Data> folded;
        int foldedIndent = -1;
        while (std::getline(in, line)) {
            size_t ind = line.find_first_not_of(' ');
            if (folded) {
                if (ind == std::string::npos) continue;
                if ((int)ind > foldedIndent) {
                    std::string piece = trim(line);
                    if (!folded->scalar.empty()) folded->scalar += " ";
                    folded->scalar += piece;
                    continue;
                }
                folded.reset();
            }
            std::string body = stripComment(line);
            if (trim(body).empty()) continue;
            ind = body.find_first_not_of(' ');
            std::string t = trim(body);
            size_t c = keyColon(t);
            if (c == std::string::npos) continue;       /* not a "key: value" line: ignored */
            std::string key = unquote(trim(t.substr(0, c)));
            std::string val = trim(t.substr(c + 1));
            while (stack.size() > 1 && (int)ind <= stack.back().indent) stack.pop_back();
            auto node = std::make_shared<Data>();
            stack.back().node->isMap = true;
            stack.back().node->kids[key] = node;
            if (val.empty()) { node->isMap = true; stack.push_back({(int)ind, node}); }
            else if (val == ">" || val == "|" || val == ">-" || val == "|-") { folded = node; foldedIndent = (int)ind; }
            else node->scalar = unquote(val);
        }
        return root;
    }

    bool IsDefined() const { return (bool)d_; }
    Node operator[](const std::string &key) const {
        Node n;
        n.path_ = path_ + "/" + key;
        auto it = d_ ? d_->kids.find(key) : std::map<std::string, std::shared_ptr<Data>>::const_iterator();
        if (!d_ || it == d_->kids.end()) n.d_.reset(); else n.d_ = it->second;
        return n;
    }
    Node operator[](const char *key) const { return (*this)[std::string(key)]; }

    template <class T>
    T as() const {
        if (!d_) throw Exception("yaml: missing key " + path_);
        return convert<T>(d_->scalar);
    }

   private:
    template <class T>
    T convert(const std::string &s) const {
        std::istringstream is(s);
        T v;
        is >> v;
        if (is.fail()) throw Exception("yaml: bad value '" + s + "' for " + path_);
        return v;
    }
};

template <>
inline std::string Node::convert<std::string>(const std::string &s) const { return s; }
template <>
inline bool Node::convert<bool>(const std::string &s) const {
    if (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on" || s == "1") return true;
    if (s == "false" || s == "False" || s == "FALSE" || s == "no" || s == "off" || s == "0") return false;
    throw Exception("yaml: bad bool '" + s + "' for " + path_);
}

inline Node Load(std::istream &in) { return Node::parse(in); }
inline Node Load(const std::string &text) { std::istringstream is(text); return Node::parse(is); }
inline Node LoadFile(const std::string &path) {
    std::ifstream f(path.c_str());
    if (!f) throw Exception("yaml: cannot open " + path);
    return Node::parse(f);
}

}  // namespace YAML

#endif
