// A YAML reader for Pingoo configuration files (pingoo.yml, rules/*.yml): the subset serde_yaml is handed by the
// documented configuration format (docs/configuration.md, docs/rules.md) -- block mappings and sequences, flow
// sequences / mappings, plain, single- and double-quoted scalars, literal (|) and folded (>) block scalars with
// chomping indicators, comments, `---`.  Anchors, aliases, tags, multi-document streams and complex keys are
// rejected with an error (the reference's own examples use none of them).
// Mapping order is preserved: Pingoo's rule and service order is the YAML order (IndexMap, config_file.rs:49-101).
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace pgw {

struct YNode {
    enum Kind { NUL, SCALAR, MAP, SEQ } kind = NUL;
    std::string s;       // SCALAR
    bool quoted = false; // SCALAR: came from a quoted or block scalar (never null / a number)
    std::vector<std::pair<std::string, YNode>> map;
    std::vector<YNode> seq;
    int line = 0;
    const YNode* get(const std::string& key) const {
        for (auto& kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_null() const { return kind == NUL || (kind == SCALAR && !quoted && (s.empty() || s == "~" || s == "null" || s == "Null" || s == "NULL")); }
};

bool yaml_parse(const std::string& text, YNode* root, std::string& err);
// canonical dump (tests compare it with PyYAML's reading of the same text)
std::string yaml_dump(const YNode& n);

}  // namespace pgw
