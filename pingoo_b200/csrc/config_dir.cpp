// Configuration directory loader: the rule-path inputs of a Pingoo configuration directory, read the way
// `config::load_and_validate` and `Server::run` read them, into an un-finalized pgw_ruleset.
//
//   <dir>/pingoo.yml   rules: {name: {expression?, actions: [{action: block|captcha}]}}          config_file.rs:97-101
//                      services: {name: {route?, http_proxy? | static? | tcp_proxy?}}            config_file.rs:49-66, 180-265
//                      listeners: {name: {address, services?}}                                   config.rs:217-244
//                      lists: {name: {type: String|Int|Ip, file}}                                config.rs:157-161
//   <dir>/rules/*.yml  more rules, appended after the file's own, directory order of the OS      config.rs:206-213, 378-422
//   geoip.mmdb[.zst]   first existing of <dir>/geoip.mmdb, <dir>/geoip.mmdb.zst, then the same two under /usr/share/pingoo
//                      (config.rs:31-36, geoip.rs:44-58, 94-109); `.zst` is decoded with the system libzstd (dlopen)
// An HTTP listener routes over the services that have `http_proxy` or `static` (config.rs:217-221, server.rs:62-74), in
// configuration order, or over its own `services:` list in that list's order (server.rs:104-110); tcp_proxy services
// never carry a route (config_file.rs:240-245) and are not offered to HTTP requests.
#include <dirent.h>
#include <dlfcn.h>

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "ruleset.hpp"
#include "yaml.hpp"

namespace pgw {

namespace {

std::string os_error(int e) { return std::string(strerror(e)) + " (os error " + std::to_string(e) + ")"; }

bool read_file(const std::string& path, std::string* out, int* err_no) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { *err_no = errno; return false; }
    out->clear();
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->append(buf, n);
    bool bad = ferror(f);
    *err_no = bad ? errno : 0;
    fclose(f);
    return !bad;
}

bool file_exists(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

// zstd::decode_all through the system library (no headers needed: the streaming API takes plain structs)
bool zstd_decode_all(const std::string& in, std::string* out, std::string& err) {
    void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { err = "libzstd is not available"; return false; }
    struct Buf { const void* p; size_t size; size_t pos; };
    struct OBuf { void* p; size_t size; size_t pos; };
    auto create = (void* (*)())dlsym(h, "ZSTD_createDStream");
    auto destroy = (size_t (*)(void*))dlsym(h, "ZSTD_freeDStream");
    auto step = (size_t (*)(void*, OBuf*, Buf*))dlsym(h, "ZSTD_decompressStream");
    auto is_err = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
    auto err_name = (const char* (*)(size_t))dlsym(h, "ZSTD_getErrorName");
    if (!create || !destroy || !step || !is_err || !err_name) { err = "libzstd lacks the streaming API"; dlclose(h); return false; }
    void* ds = create();
    Buf ib{in.data(), in.size(), 0};
    std::vector<char> chunk(1 << 20);
    out->clear();
    bool ok = true;
    size_t ret = 1;
    while (ib.pos < ib.size || ret != 0) {
        OBuf ob{chunk.data(), chunk.size(), 0};
        ret = step(ds, &ob, &ib);
        if (is_err(ret)) { err = err_name(ret); ok = false; break; }
        out->append(chunk.data(), ob.pos);
        if (ib.pos >= ib.size && ob.pos < ob.size) {
            if (ret != 0) { err = "incomplete frame"; ok = false; }
            break;
        }
    }
    destroy(ds);
    dlclose(h);
    return ok;
}

struct RuleCfg {
    std::string name;
    bool has_expression = false;
    std::string expression;
    std::vector<uint8_t> actions;
};

// IndexMap<String, RuleConfigFile>
bool rules_from_mapping(const YNode* m, const std::string& where, std::vector<RuleCfg>* out, std::string& err) {
    if (!m || m->is_null()) return true;
    if (m->kind != YNode::MAP) { err = "error parsing " + where + ": invalid type: expected a map of rules"; return false; }
    for (auto& kv : m->map) {
        const YNode& c = kv.second;
        RuleCfg r;
        r.name = kv.first;
        const YNode* acts = c.kind == YNode::MAP ? c.get("actions") : nullptr;
        if (!acts) { err = "error parsing " + where + ": " + r.name + ": missing field `actions`"; return false; }
        if (!acts->is_null() && acts->kind != YNode::SEQ) { err = "error parsing " + where + ": " + r.name + ": actions: invalid type: expected a sequence"; return false; }
        for (auto& a : acts->seq) {
            const YNode* tag = a.kind == YNode::MAP ? a.get("action") : nullptr;
            if (!tag || tag->kind != YNode::SCALAR) { err = "error parsing " + where + ": " + r.name + ": missing field `action`"; return false; }
            if (tag->s == "block") r.actions.push_back(ACT_BLOCK);
            else if (tag->s == "captcha") r.actions.push_back(ACT_CAPTCHA);
            else { err = "error parsing " + where + ": " + r.name + ": unknown variant `" + tag->s + "`, expected `block` or `captcha`"; return false; }
        }
        const YNode* e = c.get("expression");
        if (e && !e->is_null()) {
            if (e->kind != YNode::SCALAR) { err = "error parsing " + where + ": " + r.name + ": expression: invalid type: expected a string"; return false; }
            r.has_expression = true;
            r.expression = e->s;
        }
        out->push_back(std::move(r));
    }
    return true;
}

}  // namespace

bool load_config_dir(const std::string& folder, const char* listener, const std::vector<std::string>& geoip_dirs, RulesetBuilder* B,
                     std::string* geoip_path, std::string& err) {
    const std::string cfg_path = folder + "/pingoo.yml";
    std::string raw;
    int en = 0;
    if (!read_file(cfg_path, &raw, &en)) { err = "error reading config file (" + cfg_path + "): " + os_error(en); return false; }
    YNode doc;
    std::string perr;
    if (!yaml_parse(raw, &doc, perr)) { err = "error parsing config file (" + cfg_path + "): " + perr; return false; }
    if (!doc.is_null() && doc.kind != YNode::MAP) { err = "error parsing config file (" + cfg_path + "): invalid type: expected a map"; return false; }

    std::vector<RuleCfg> rules;
    if (!rules_from_mapping(doc.get("rules"), "config file (" + cfg_path + ")", &rules, err)) return false;

    // rules folder: *.yml only, in the order the OS lists them (config.rs:378-422)
    std::vector<RuleCfg> folder_rules;
    const std::string rules_dir = folder + "/rules";
    if (DIR* d = opendir(rules_dir.c_str())) {
        while (dirent* ent = readdir(d)) {
            const std::string name = ent->d_name;
            const size_t dot = name.rfind('.');
            if (name == "." || name == ".." || dot == std::string::npos || dot == 0 || name.substr(dot) != ".yml") continue;
            const std::string path = rules_dir + "/" + name;
            std::string content;
            if (!read_file(path, &content, &en)) { closedir(d); err = "error reading rules file \"" + path + "\": " + os_error(en); return false; }
            YNode m;
            if (!yaml_parse(content, &m, perr)) { closedir(d); err = "error parsing rules file \"" + path + "\": " + perr; return false; }
            std::vector<RuleCfg> fresh;
            if (!rules_from_mapping(&m, "rules file \"" + path + "\"", &fresh, err)) { closedir(d); return false; }
            for (auto& r : fresh) {
                for (auto& o : folder_rules)
                    if (o.name == r.name) { closedir(d); err = "duplicate rule name: " + r.name; return false; }
            }
            for (auto& r : fresh) folder_rules.push_back(std::move(r));
        }
        closedir(d);
    } else if (errno != ENOENT) {
        err = "error reading rules folder \"" + rules_dir + "\": " + os_error(errno);
        return false;
    }
    for (auto& r : folder_rules)
        for (auto& o : rules)
            if (o.name == r.name) { err = "duplicate rule name: " + r.name; return false; }
    for (auto& r : folder_rules) rules.push_back(std::move(r));

    // services (config_file.rs:180-265): exactly one of http_proxy / static / tcp_proxy; a tcp proxy has no route
    struct Svc { std::string name; bool has_route = false; std::string route; bool http = false; };
    std::vector<Svc> services;
    if (const YNode* sv = doc.get("services")) {
        if (!sv->is_null() && sv->kind != YNode::MAP) { err = "error parsing config file (" + cfg_path + "): services: invalid type: expected a map"; return false; }
        for (auto& kv : sv->map) {
            Svc s;
            s.name = kv.first;
            const YNode& c = kv.second;
            auto present = [&](const char* k) { const YNode* x = c.kind == YNode::MAP ? c.get(k) : nullptr; return x && !(x->kind == YNode::SCALAR && x->is_null()) && x->kind != YNode::NUL; };
            const bool hp = present("http_proxy"), st = present("static"), tp = present("tcp_proxy");
            if ((int)hp + (int)st + (int)tp != 1) {
                err = "invalid service definition for " + s.name + ": services must have exactly 1 http_proxy, tcp_proxy or static field";
                return false;
            }
            const YNode* rt = c.get("route");
            if (rt && !rt->is_null()) {
                if (rt->kind != YNode::SCALAR) { err = "error parsing config file (" + cfg_path + "): services." + s.name + ".route: invalid type: expected a string"; return false; }
                s.has_route = true;
                s.route = rt->s;
            }
            if (tp && s.has_route) { err = "Invalid service definition for " + s.name + ": TCP proxy can't have a route"; return false; }
            // parse_service compiles the route right here (config_file.rs:257-265), service by service, before the listeners are
            // validated and before any rule is compiled (config.rs:217-269): with several mistakes in a file, this is the one reported
            if (s.has_route) {
                std::string cerr;
                if (!RulesetBuilder::compile_expression(s.route, cerr)) { err = "error parsing route for service " + s.name + ": " + cerr; return false; }
            }
            s.http = hp || st;
            services.push_back(std::move(s));
        }
    }
    // which services an HTTP listener offers a request to, and in which order
    std::vector<const Svc*> offered;
    bool explicit_list = false;
    if (listener) {
        const YNode* ls = doc.get("listeners");
        const YNode* l = ls && ls->kind == YNode::MAP ? ls->get(listener) : nullptr;
        if (!l) { err = std::string("config: listeners: ") + listener + ": no such listener"; return false; }
        const YNode* lsv = l->kind == YNode::MAP ? l->get("services") : nullptr;
        if (lsv && !lsv->is_null()) {
            if (lsv->kind != YNode::SEQ) { err = std::string("error parsing config file (") + cfg_path + "): listeners." + listener + ".services: invalid type: expected a sequence"; return false; }
            explicit_list = true;
            std::set<std::string> seen;
            for (auto& item : lsv->seq) {
                const Svc* found = nullptr;
                for (auto& s : services)
                    if (s.name == item.s) found = &s;
                if (!found) { err = std::string("config: listeners: ") + listener + ": service " + item.s + " doesn't exist"; return false; }
                if (!seen.insert(item.s).second) { err = std::string("config: listeners: ") + listener + ": duplicate services are not allowed (" + item.s + ")"; return false; }
                if (!found->http) { err = std::string("config: listeners: ") + listener + ": service " + item.s + " is not an HTTP service"; return false; }
                offered.push_back(found);
            }
        }
    }
    if (!explicit_list)
        for (auto& s : services)
            if (s.http) offered.push_back(&s);

    // hand everything to the builder; compile errors are fatal with the reference's texts (config.rs:255-269, config_file.rs:257-265)
    for (auto& r : rules)
        if (!B->add_rule(r.name.c_str(), r.has_expression ? r.expression.c_str() : nullptr, r.actions.data(), (uint32_t)r.actions.size(), err)) return false;
    for (const Svc* s : offered)
        if (!B->add_service(s->name.c_str(), s->has_route ? s->route.c_str() : nullptr, err)) return false;

    if (const YNode* ls = doc.get("lists")) {
        if (!ls->is_null() && ls->kind != YNode::MAP) { err = "error parsing config file (" + cfg_path + "): lists: invalid type: expected a map"; return false; }
        for (auto& kv : ls->map) {
            const YNode& c = kv.second;
            const YNode* ty = c.kind == YNode::MAP ? c.get("type") : nullptr;
            const YNode* fl = c.kind == YNode::MAP ? c.get("file") : nullptr;
            if (!ty || ty->kind != YNode::SCALAR) { err = "error parsing config file (" + cfg_path + "): lists." + kv.first + ": missing field `type`"; return false; }
            if (!fl || fl->kind != YNode::SCALAR) { err = "error parsing config file (" + cfg_path + "): lists." + kv.first + ": missing field `file`"; return false; }
            int type;
            if (ty->s == "String") type = LT_STRING;
            else if (ty->s == "Int") type = LT_INT;
            else if (ty->s == "Ip") type = LT_IP;
            else { err = "error parsing config file (" + cfg_path + "): lists." + kv.first + ": unknown variant `" + ty->s + "`, expected one of `String`, `Int`, `Ip`"; return false; }
            std::string csv;
            // lists.rs:62-66 and :82-109 name the list by its PATH in every message
            if (!read_file(fl->s, &csv, &en)) { err = "error reading list " + fl->s + ": " + os_error(en); return false; }
            if (!B->add_list(kv.first.c_str(), type, (const uint8_t*)csv.data(), csv.size(), err)) {
                const std::string by_name = "error parsing list " + kv.first + " at line";
                if (err.compare(0, by_name.size(), by_name) == 0) err = "error parsing list " + fl->s + " at line" + err.substr(by_name.size());
                return false;
            }
        }
    }

    for (auto& dir : geoip_dirs)
        for (const char* nm : {"geoip.mmdb", "geoip.mmdb.zst"}) {
            const std::string p = dir + "/" + nm;
            if (!file_exists(p)) continue;
            std::string data;
            if (!read_file(p, &data, &en)) { err = "error reading geoip database (" + p + "): " + os_error(en); return false; }
            if (p.size() >= 4 && p.compare(p.size() - 4, 4, ".zst") == 0) {
                std::string plain, zerr;
                if (!zstd_decode_all(data, &plain, zerr)) { err = "error decompressing geoip database (" + p + "): " + zerr; return false; }
                data.swap(plain);
            }
            if (!B->load_geoip((const uint8_t*)data.data(), data.size(), err)) return false;
            if (geoip_path) *geoip_path = p;
            return true;
        }
    return true;
}

}  // namespace pgw
