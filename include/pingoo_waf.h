/* pingoo_waf.h -- C ABI of the B200 batched WAF verdict engine.
 *
 * Drop-in boundary for Pingoo's per-request rules / lists / GeoIP hot path.
 * The reference has no FFI; each entry point below names the Rust-internal
 * call it replaces (paths relative to the pingooio/pingoo tree @ eecc74d).
 * INTEGRATION.md shows the Rust `extern "C"` block a maintainer would add.
 *
 * All functions return 0 on success and non-zero on failure; when `err` is
 * non-NULL the failure text is written there (NUL-terminated, truncated to
 * `err_cap`), and it is also available from pgw_last_error() (thread-local).
 * There is no CPU fallback: every evaluate call runs CUDA kernels on an
 * sm_100a device or fails.
 */
#ifndef PINGOO_WAF_H
#define PINGOO_WAF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgw_ruleset pgw_ruleset;

/* rules::Action (rules/rules.rs:30-35) */
enum { PGW_ACTION_BLOCK = 1, PGW_ACTION_CAPTCHA = 2 };

/* ListType (pingoo/lists.rs:17-22) */
enum { PGW_LIST_STRING = 0, PGW_LIST_INT = 1, PGW_LIST_IP = 2 };

/* verdict word = action | rule_index << 2 ; rule_index == PGW_NO_RULE when no rule decided */
enum { PGW_ALLOW = 0, PGW_BLOCK = 1, PGW_CAPTCHA = 2, PGW_BYPASS_CAPTCHA_API = 3 };
#define PGW_NO_RULE 0x3FFFFFFFu
#define PGW_VERDICT_ACTION(v) ((v) & 3u)
#define PGW_VERDICT_RULE(v) ((v) >> 2)

/* per-request flag bits (pgw_batch.flags) */
enum {
    PGW_FLAG_CAPTCHA_VERIFIED = 1, /* valid __pingoo_captcha_verified cookie  (listeners/http_listener.rs:222-236) */
    PGW_FLAG_PRE_BLOCK = 2,        /* host gate already blocked the request   (http_listener.rs:159-165,196-198)   */
    PGW_FLAG_PRE_CAPTCHA = 4,      /* cookie present but invalid              (http_listener.rs:231-235)           */
    PGW_FLAG_BYPASS = 8            /* request addressed to /__pingoo/captcha  (http_listener.rs:200-204)           */
};

/* pingoo::rules::Rule{name, expression: Option<..>, actions}  (pingoo/rules.rs:9-14) */
typedef struct pgw_rule_desc {
    const char* name;
    const char* expression;  /* NULL => the rule matches every request (pingoo/rules.rs:49-51) */
    const uint8_t* actions;  /* PGW_ACTION_* in configuration order */
    uint32_t n_actions;
} pgw_rule_desc;

typedef struct pgw_options {
    int32_t max_dfa_states;        /* per scan unit; 0 = default (16384); a pattern that needs more on its own is not refused:
                                      it runs as a bit-parallel NFA (up to 2048 positions) over every request */
    uint64_t max_unit_table_bytes; /* per scan unit; 0 = default (8 MiB) */
    int32_t eval_gates;            /* 1 (default): evaluate the user-agent and captcha-path gates of
                                      http_listener.rs:196-204 inside the engine; 0: rules only */
    int32_t disable_candidate_gate; /* 0 (default): url / user_agent / path patterns are prefiltered by the gram gate and
                                      only candidate requests are walked by their DFAs; 1: every request is walked
                                      (same verdicts; kept for measurements and tests); 2: the gate stays, but finite-string
                                      patterns are walked by the DFAs like every other pattern instead of being confirmed by
                                      the gate's resolve kernel (same verdicts; measurements and tests) */
} pgw_options;

/* One string column: concatenated bytes + n+1 offsets.  `bytes` must be 32-byte aligned and readable up to
 * round_up(offsets[n], 32) (the kernels load aligned 16-byte chunks; what lies behind offsets[n] is read but never
 * interpreted as request data).  A column holds less than 4 GiB - 4 KiB (32-bit offsets); callers split larger batches.
 * A batch may be a window of a longer column: offsets are absolute positions in `bytes`, offsets[0] need not be 0. */
typedef struct pgw_strcol {
    const uint8_t* bytes;
    const uint32_t* offsets;
} pgw_strcol;

/* Columnar request batch = RequestData + ClientData (pingoo/rules.rs:16-34) for n requests. */
typedef struct pgw_batch {
    uint32_t n;
    pgw_strcol host, url, path, method, user_agent;
    const uint8_t* ip;          /* n x 16 bytes, network order; IPv4 in bytes [0,4), rest zero */
    const uint8_t* ip_is_v6;    /* n */
    const int32_t* remote_port; /* n */
    const int64_t* asn;         /* n, or NULL: resolved from the loaded GeoIP db (else 0)  */
    const uint16_t* country;    /* n, or NULL; two ASCII letters, first letter in the low byte */
    const uint8_t* flags;       /* n, or NULL (= all zero) */
} pgw_batch;

typedef struct pgw_info {
    uint32_t n_rules, n_atoms, n_scan_units, n_nonscan_atoms;
    uint32_t scanned_fields_mask; /* bit f set: bytes of field f are read (host,url,path,method,user_agent = 0..4) */
    uint32_t offset_fields_mask;  /* bit f set: offsets of field f are read */
    uint32_t reads_ip, reads_port, reads_geo_columns;
    uint64_t table_arena_bytes, smem_bytes;
    uint32_t tables_in_smem;      /* every DFA row is resident in shared memory */
    uint32_t hot_dfa_states;      /* DFA states whose rows are in the shared-memory image (the rest is read from L2) */
    uint32_t grid, threads;       /* launch shape of the scan kernel of the active path */
    uint32_t total_dfa_states, lpm_present, geoip_loaded;
    uint64_t kernel_launches;     /* launches issued through this ruleset so far */
    uint64_t last_h2d_bytes, last_d2h_bytes; /* bytes moved by the last pgw_evaluate_batch_host call */
    uint32_t gated_fields_mask;   /* fields in front of whose DFAs the candidate gate runs */
    uint32_t gate_grams;          /* 4-byte grams in the gate bitmaps, all fields */
    uint64_t gate_smem_bytes;     /* shared memory of the gate kernel (the level-1 bitmaps of all gated fields) */
    uint32_t n_bitset_units;      /* patterns simulated as bit-parallel NFAs because no DFA unit can hold them */
    uint32_t bitset_positions;    /* NFA positions of those units, all together */
} pgw_info;

/* rules::compile_expression(&str) -> Result<CompiledExpression, Error>   (rules/rules.rs:45-53) */
int pgw_compile_expression(const char* expression, char* err, size_t err_cap);
/* rules::validate_expression(&str) -> Result<(), Error>                  (rules/rules.rs:55-77) */
int pgw_validate_expression(const char* expression, char* err, size_t err_cap);

/* Config rule compilation: the loop of pingoo/config/config.rs:255-269.  Rule order = array order. */
int pgw_ruleset_create(const pgw_rule_desc* rules, uint32_t n_rules, const pgw_options* options, pgw_ruleset** out,
                       char* err, size_t err_cap);
/* config::load_and_validate + the loading half of Server::run for the rule path (pingoo/config/config.rs:194-323, 378-422;
 * config_file.rs:97-101, 180-265; server.rs:40-47; geoip.rs:44-58, 94-109), on a configuration DIRECTORY:
 * <dir>/pingoo.yml (rules, services, listeners, lists), <dir>/rules/ (*.yml), list CSV files, geoip.mmdb[.zst] (first of
 * `geoip_dirs`; n_geoip_dirs == 0: <dir> then /usr/share/pingoo; `.zst` needs the system libzstd).  Rule order = file
 * order, then the rules folder; duplicate rule names, expressions / routes that do not compile, unknown actions and
 * malformed services are errors with the reference's messages.  The services installed (pgw_services_set) are the ones
 * an HTTP listener offers a request to: `listener` NULL = every service with http_proxy or static, in configuration
 * order (config.rs:217-221); else the `services:` list of that listener, in its own order (server.rs:104-110).
 * The ruleset comes back un-finalized: pgw_ruleset_finalize uploads it. */
int pgw_ruleset_load_dir(const char* config_folder, const char* listener, const char* const* geoip_dirs, uint32_t n_geoip_dirs,
                         const pgw_options* options, pgw_ruleset** out, char* err, size_t err_cap);
/* pingoo::lists::load_list on an in-memory CSV                          (pingoo/lists.rs:62-113) */
int pgw_lists_add(pgw_ruleset* rs, const char* name, int list_type, const uint8_t* csv, size_t csv_len, char* err,
                  size_t err_cap);
/* GeoipDB::load on an already-decompressed .mmdb image                  (pingoo/geoip.rs:44-71) */
int pgw_geoip_load(pgw_ruleset* rs, const uint8_t* mmdb, size_t mmdb_len, char* err, size_t err_cap);
/* Lower rules against the loaded lists, build device tables, upload to `device`. */
int pgw_ruleset_finalize(pgw_ruleset* rs, int device, char* err, size_t err_cap);

/* The per-request body of http_listener.rs:239-264 for a whole batch: context build, rule loop,
 * action interpretation.  All pointers in `batch` and `verdict_out` are DEVICE pointers; the kernel is
 * enqueued on `stream` (a cudaStream_t, NULL = default stream) and the call does not synchronise.
 * Thread-safe on a finalized ruleset. */
int pgw_evaluate_batch(const pgw_ruleset* rs, const pgw_batch* batch, uint32_t* verdict_out, void* stream);
/* Same with HOST pointers: copies the columns to the device, evaluates, copies verdicts back, synchronises.
 * Thread-safe on a finalized ruleset: every call takes its own staging buffers and streams from a pool
 * (the reference evaluates concurrently on shared Arcs, http_listener.rs:91-103,134-138). */
int pgw_evaluate_batch_host(pgw_ruleset* rs, const pgw_batch* batch, uint32_t* verdict_out);

/* GeoipDB::lookup for a batch of addresses (pingoo/geoip.rs:73-91); device pointers, not-found => {0,"XX"}. */
int pgw_geoip_lookup_batch(const pgw_ruleset* rs, const uint8_t* ip, const uint8_t* ip_is_v6, uint32_t n,
                           uint32_t* asn_out, uint16_t* country_out, void* stream);

/* Pinned (page-locked) host memory for batch columns, so the host-pointer path runs at PCIe speed. */
void* pgw_host_alloc(size_t bytes);
void pgw_host_free(void* p);

/* ---- service routes, evaluated in the same pass -----------------------------------------------
 * http_listener.rs:266-272 + HttpService::match_request (services/mod.rs:33-37, http_proxy_service.rs:84-95,
 * http_static_site_service.rs:70-81): a request the rules let through is offered to the services in configuration
 * order; a service takes it if it has no `route` or its route evaluates to Bool(true) (an error or a non-bool value
 * is "no match"); if none does the listener answers 404.  A route is compiled like a rule expression
 * (config_file.rs:257-265: a route that does not compile is a fatal configuration error).
 * pgw_services_set is called between pgw_ruleset_create and pgw_ruleset_finalize.  The routed entry points
 * additionally write, per request, the index of the service that takes it, or PGW_NO_SERVICE when none matches
 * or when the verdict's action is not PGW_ALLOW (the reference never consults the services for those). */
typedef struct pgw_service_desc {
    const char* name;
    const char* route; /* NULL: the service matches every request */
} pgw_service_desc;
#define PGW_NO_SERVICE 0xFFFFu
int pgw_services_set(pgw_ruleset* rs, const pgw_service_desc* services, uint32_t n, char* err, size_t err_cap);
int pgw_evaluate_batch_routed(const pgw_ruleset* rs, const pgw_batch* batch, uint32_t* verdict_dev, uint16_t* service_dev, void* stream);
int pgw_evaluate_batch_routed_host(pgw_ruleset* rs, const pgw_batch* batch, uint32_t* verdict_host, uint16_t* service_host);

/* ---- request packer + micro-batching queue ------------------------------------------------------
 * The reference evaluates each request inline on the worker that serves the connection (http_listener.rs:139-272);
 * nothing like this queue exists there.  Worker threads hand single requests -- the strings the listener has in hand --
 * to pgw_queue_evaluate, which blocks until the verdict is known.  The queue applies the listener's shaping (host:
 * header to_str + trim + at most 256 bytes, http_listener.rs:284-296; path: trailing '/' trimmed, http_utils.rs:114-116;
 * user agent: to_str + trim + at most 256 bytes, http_listener.rs:159-165), packs the requests into the columnar
 * pgw_batch in pinned memory and evaluates a batch (pgw_evaluate_batch_routed_host) when `max_batch` requests are
 * waiting or the oldest has waited `max_delay_us`.  Two batches are in flight at most: one being filled, one being
 * evaluated.  While a queue exists it must be the only caller of the host-pointer entry points of its ruleset.
 * asn / country columns are not part of a queued request: GeoIP is resolved on the device (pgw_geoip_load). */
typedef struct pgw_request {
    const char* host; size_t host_len;             /* uri.host() or the Host header, raw */
    const char* url; size_t url_len;               /* uri.to_string() */
    const char* path; size_t path_len;             /* uri.path() */
    const char* method; size_t method_len;
    const char* user_agent; size_t user_agent_len; /* raw header value */
    uint8_t ip[16];                                /* network order, IPv4 in bytes 0-3 */
    uint8_t ip_is_v6;
    uint8_t flags;                                 /* PGW_FLAG_* */
    int32_t remote_port;
} pgw_request;
typedef struct pgw_queue_stats {
    uint64_t batches, requests, full_flushes, deadline_flushes;
    uint32_t largest_batch, reserved;
} pgw_queue_stats;
typedef struct pgw_queue pgw_queue;
int pgw_queue_create(pgw_ruleset* rs, uint32_t max_batch, uint32_t max_delay_us, pgw_queue** out, char* err, size_t err_cap);
/* Thread-safe, blocking.  0 on success; `service` may be NULL. */
int pgw_queue_evaluate(pgw_queue* q, const pgw_request* req, uint32_t* verdict, uint16_t* service);
/* Thread-safe, returns as soon as the request is packed (its strings are copied); `done` is called exactly once from the
 * queue's dispatcher thread when the batch has been evaluated -- the shape an async runtime (tokio) wants: wake a task. */
typedef void (*pgw_done_fn)(void* user, uint32_t verdict, uint16_t service, int rc);
int pgw_queue_submit(pgw_queue* q, const pgw_request* req, pgw_done_fn done, void* user);
int pgw_queue_get_stats(pgw_queue* q, pgw_queue_stats* out);
/* Message of the most recent batch that failed (rc 4 of pgw_queue_evaluate / the callback): the batch is evaluated on the
 * queue's dispatcher thread, whose pgw_last_error() the callers cannot see.  Returns the message length. */
size_t pgw_queue_last_error(pgw_queue* q, char* buf, size_t cap);
/* Flushes what is pending; no thread may be inside pgw_queue_evaluate any more. */
void pgw_queue_destroy(pgw_queue* q);
/* The shaping alone (no device needed): pointers into the request's own strings, Field order host,url,path,method,user_agent. */
int pgw_shape_request(const pgw_request* req, const char* out_ptr[5], size_t out_len[5]);

/* ---- captcha client id for a batch -------------------------------------------------------------
 * generate_captcha_client_id (pingoo/captcha.rs:409-421), called per request at http_listener.rs:167:
 * base64url-no-pad(SHA-256(ip octets (4 or 16) || user_agent || host)) = 43 characters.  `batch` holds DEVICE pointers
 * on the current device (ip, ip_is_v6, user_agent, host are read); out44_dev receives n x 44 bytes (43 characters and
 * a terminating 0).  No ruleset is involved. */
int pgw_captcha_client_id_batch(const pgw_batch* batch, uint8_t* out44_dev, void* stream);

int pgw_ruleset_info(const pgw_ruleset* rs, pgw_info* out);
/* Measurement hook (no reference counterpart): while enabled, every batch is bracketed by CUDA events around its
 * kernels, on the stream they are launched on (a ring of 256 batches).  pgw_ruleset_profile synchronises those events
 * and returns the summed pre-pass + scan time and the number of batches it covers (at most the 256 most recent), then
 * clears the ring. */
int pgw_ruleset_set_profiling(pgw_ruleset* rs, int enable);
int pgw_ruleset_profile(pgw_ruleset* rs, double* scan_ms_sum, uint32_t* launches);
/* The same ring, per kernel group: ms_sum[0] = pre-pass kernel (candidate gate + small early-exit units), [1] = DFA scan
 * launches, [2] = epilogue + multi-atom kernels; `batches` = how many batches the sums cover. */
int pgw_ruleset_profile_kernels(pgw_ruleset* rs, double ms_sum[3], uint32_t* batches);
/* Human-readable compile summary / warnings (e.g. a regex that does not compile => that rule never matches). */
size_t pgw_ruleset_describe(const pgw_ruleset* rs, char* buf, size_t cap);
void pgw_ruleset_destroy(pgw_ruleset* rs);
const char* pgw_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PINGOO_WAF_H */
