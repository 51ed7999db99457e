/* ORACLE (test infrastructure, never shipped): tree-walking evaluator restating
 * `bel 0.11.0` (pingooio/stdx-rs @ 70c3c14, reference Cargo.lock:142-154) for the
 * documented subset of the rule language (reference docs/rules.md:35-76).
 * `bel` is not vendored under /root/reference: PARITY UNPINNED -- semantics follow
 * SEMANTICS.md assumptions A1-A9 and are pinned only against the reference's doc
 * examples plus independent cross-checks (Python re / ipaddress). */
#ifndef ORACLE_BEL_H
#define ORACLE_BEL_H
#include <stddef.h>
#include <stdint.h>

typedef struct bel_expr bel_expr;

typedef struct {
    int v6;
    uint8_t addr[16]; /* IPv4 in addr[0..4) */
    int prefix;
} bel_ipnet;

enum { BEL_LIST_STRING = 0, BEL_LIST_INT = 1, BEL_LIST_IP = 2 };

typedef struct {
    char* name;
    int type;
    size_t n;
    char** strs;
    size_t* str_lens;
    int64_t* ints;
    bel_ipnet* nets;
} bel_list;

typedef struct {
    const bel_list* lists;
    size_t n_lists;
} bel_lists;

/* request context: the variables of pingoo/rules.rs:16-34 */
typedef struct {
    const uint8_t* str[5]; /* host, url, path, method, user_agent */
    size_t len[5];
    uint8_t ip[16];
    int ip_is_v6;
    int64_t remote_port;
    int64_t asn;
    char country[2];
    const bel_lists* lists;
} bel_ctx;

/* rules::compile_expression: NULL + message on a syntax error */
bel_expr* bel_compile(const char* src, char* err, size_t cap);
void bel_free(bel_expr* e);
/* 1 if the program references the `in` operator ("@in", rules/rules.rs:67-71) */
int bel_uses_in(const bel_expr* e);

/* bel::Program::execute + `== true.into()` (pingoo/rules.rs:36-52):
 * returns 1 iff the expression evaluates, without error, to Bool(true). */
int bel_matches(const bel_expr* e, const bel_ctx* ctx);
/* raw outcome for tests: 0 false, 1 true, 2 error, 3 non-bool value */
int bel_eval_kind(const bel_expr* e, const bel_ctx* ctx);

int bel_parse_ipnet(const char* s, bel_ipnet* out);
int bel_ipnet_contains(const bel_ipnet* net, const uint8_t* ip16, int is_v6);

#endif
