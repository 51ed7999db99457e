/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called
 * from the product (pingoo_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, as the checker or
 * the timed CPU baseline.
 *
 * CPU restatement of Pingoo's per-request WAF path:
 *   rules/rules.rs:22-77, pingoo/rules.rs:9-52, pingoo/lists.rs:11-125,
 *   pingoo/geoip.rs:12-174, pingoo/serde_utils.rs:1-9,
 *   pingoo/listeners/http_listener.rs:139-272, pingoo/services/http_proxy_service.rs:84-95.
 * The arithmetic lives in crates that are NOT vendored in the reference
 * (bel 0.11.0, regex 1.12.2, ipnetwork 0.21.1, maxminddb 0.24.0 -- Cargo.lock)
 * and the reference has no tests or golden vectors: PARITY UNPINNED.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#include "../include/pingoo_waf.h" /* shares the batch / rule descriptor structs and verdict encoding */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ruleset orc_ruleset;

int orc_compile_expression(const char* expr, char* err, size_t cap);   /* rules::compile_expression */
int orc_validate_expression(const char* expr, char* err, size_t cap);  /* rules::validate_expression */

orc_ruleset* orc_create(const pgw_rule_desc* rules, uint32_t n, int eval_gates, char* err, size_t cap);
int orc_lists_add(orc_ruleset* rs, const char* name, int type, const uint8_t* csv, size_t len, char* err, size_t cap);
int orc_geoip_load(orc_ruleset* rs, const uint8_t* mmdb, size_t len, char* err, size_t cap);
/* Evaluate a host batch request by request with `n_threads` workers (static partition). */
int orc_evaluate(const orc_ruleset* rs, const pgw_batch* batch, uint32_t* verdict_out, int n_threads);
/* services (http_listener.rs:266-272, http_proxy_service.rs:84-95): set before evaluating; the routed form also writes the
 * index of the first matching service for allowed requests (PGW_NO_SERVICE otherwise) */
int orc_services_set(orc_ruleset* rs, const pgw_service_desc* services, uint32_t n, char* err, size_t cap);
int orc_evaluate_routed(const orc_ruleset* rs, const pgw_batch* batch, uint32_t* verdict_out, uint16_t* service_out, int n_threads);
/* GeoipDB::lookup + the caller's fallback (http_listener.rs:143-157): always fills a record. */
void orc_geoip_lookup(const orc_ruleset* rs, const uint8_t ip[16], int is_v6, uint32_t* asn, uint16_t* country);
void orc_destroy(orc_ruleset* rs);

/* single-purpose hooks for differential tests */
int orc_regex_is_match(const char* pattern, size_t plen, const uint8_t* hay, size_t n); /* 1/0, -1 invalid, -2 unsupported, -3 too big */
int orc_ipnet_contains(const char* net, const uint8_t ip[16], int is_v6);               /* 1/0, -1 parse error */
int orc_eval_kind(const char* expr, const pgw_batch* one_request_batch);                /* 0 false 1 true 2 error 3 non-bool, -1 syntax */

#ifdef __cplusplus
}
#endif
#endif
