// Request packer + micro-batching queue.
//
// The reference evaluates every request inline on the tokio worker that serves the connection
// (pingoo/listeners/http_listener.rs:139-272); a batched engine needs the opposite shape: worker threads hand single
// requests to a queue, the queue packs them into the columnar pgw_batch and evaluates a batch when it is full or when
// its oldest request has waited `max_delay_us`.  This file is that host-side piece, above the C ABI it only uses
// pgw_host_alloc / pgw_evaluate_batch_routed_host.
//
// Shaping restated from the listener (the rules must see exactly what the reference's rules see):
//   host        uri.host() or the Host header, `to_str()` (visible ASCII or tab, else ""), trimmed, more than 256 bytes
//               -> ""                                                         http_listener.rs:284-296
//   path        uri.path() with trailing '/' trimmed ("/" -> "")              services/http_utils.rs:114-116
//   user_agent  header `to_str()` (visible ASCII or tab, else ""), trimmed, more than 256 bytes -> ""
//                                                                              http_listener.rs:159-165
//   url, method verbatim                                                      http_listener.rs:239-249
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include "../../include/pingoo_waf.h"

namespace {

using Clock = std::chrono::steady_clock;

bool header_to_str_ok(const char* s, size_t n) {  // http::HeaderValue::to_str: every byte visible ASCII (32..126) or tab
    for (size_t i = 0; i < n; ++i) {
        const unsigned char c = (unsigned char)s[i];
        if (!((c >= 32 && c < 127) || c == '\t')) return false;
    }
    return true;
}

bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }  // str::trim on ASCII input

void trim(const char*& s, size_t& n) {
    while (n && is_space((unsigned char)s[0])) { ++s; --n; }
    while (n && is_space((unsigned char)s[n - 1])) --n;
}

struct Shaped {
    const char* p[5];
    size_t n[5];
};

void shape(const pgw_request* r, Shaped* out) {
    const char* host = r->host ? r->host : "";
    size_t hn = r->host ? r->host_len : 0;
    if (!header_to_str_ok(host, hn)) hn = 0;
    trim(host, hn);
    if (hn > 256) hn = 0;
    const char* path = r->path ? r->path : "";
    size_t pn = r->path ? r->path_len : 0;
    while (pn && path[pn - 1] == '/') --pn;
    const char* ua = r->user_agent ? r->user_agent : "";
    size_t un = r->user_agent ? r->user_agent_len : 0;
    if (!header_to_str_ok(ua, un)) un = 0;
    trim(ua, un);
    if (un > 256) un = 0;
    out->p[0] = host; out->n[0] = hn;
    out->p[1] = r->url ? r->url : ""; out->n[1] = r->url ? r->url_len : 0;
    out->p[2] = path; out->n[2] = pn;
    out->p[3] = r->method ? r->method : ""; out->n[3] = r->method ? r->method_len : 0;
    out->p[4] = ua; out->n[4] = un;
}

struct Side {
    uint8_t* bytes[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[5] = {0, 0, 0, 0, 0}, used[5] = {0, 0, 0, 0, 0};
    uint32_t* offs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    uint8_t *ip = nullptr, *v6 = nullptr, *flags = nullptr;
    int32_t* port = nullptr;
    uint32_t* verdict = nullptr;
    uint16_t* service = nullptr;
    pgw_done_fn* fn = nullptr;   // per request: completion callback (null: a thread blocks in pgw_queue_evaluate)
    void** user = nullptr;
    uint32_t n = 0;
    uint32_t readers_left = 0;
    bool done = false, free_ = true;
    int rc = 0;
    Clock::time_point first;
};

bool grow(Side& s, int f, size_t need) {
    if (need <= s.cap[f]) return true;
    size_t cap = s.cap[f] ? s.cap[f] : 4096;
    while (cap < need) cap *= 2;
    uint8_t* nb = (uint8_t*)pgw_host_alloc(cap);
    if (!nb) return false;
    if (s.used[f]) memcpy(nb, s.bytes[f], s.used[f]);
    if (s.bytes[f]) pgw_host_free(s.bytes[f]);
    s.bytes[f] = nb;
    s.cap[f] = cap;
    return true;
}

}  // namespace

struct pgw_queue {
    pgw_ruleset* rs = nullptr;
    uint32_t max_batch = 0;
    std::chrono::microseconds max_delay{0};
    Side side[2];
    int fill = 0;  // index of the side being filled
    std::mutex mu;
    std::condition_variable cv_work, cv_space, cv_done, cv_free;
    std::thread worker;
    bool stop = false;
    pgw_queue_stats st{};
};

static void dispatcher(pgw_queue* q) {
    std::unique_lock<std::mutex> lk(q->mu);
    for (;;) {
        q->cv_work.wait(lk, [&] { return q->stop || q->side[q->fill].n > 0; });
        if (q->stop && q->side[q->fill].n == 0) return;
        Side* s = &q->side[q->fill];
        // full, or the oldest request has waited long enough
        const bool full = q->cv_work.wait_until(lk, s->first + q->max_delay, [&] { return q->stop || s->n == q->max_batch; }) && s->n == q->max_batch;
        if (s->n == 0) continue;
        Side* other = &q->side[q->fill ^ 1];
        q->cv_free.wait(lk, [&] { return other->free_; });
        // the other side becomes the one being filled
        other->free_ = false;
        other->n = 0;
        other->done = false;
        for (int f = 0; f < 5; ++f) other->used[f] = 0;
        q->fill ^= 1;
        q->cv_space.notify_all();
        const uint32_t n = s->n;
        uint32_t blocking = 0;
        for (uint32_t i = 0; i < n; ++i) blocking += s->fn[i] == nullptr;
        s->readers_left = blocking;
        if (full) q->st.full_flushes++;
        else q->st.deadline_flushes++;
        q->st.batches++;
        q->st.requests += n;
        if (n > q->st.largest_batch) q->st.largest_batch = n;
        lk.unlock();
        pgw_batch b;
        memset(&b, 0, sizeof b);
        b.n = n;
        pgw_strcol* cols[5] = {&b.host, &b.url, &b.path, &b.method, &b.user_agent};
        for (int f = 0; f < 5; ++f) {
            s->offs[f][n] = (uint32_t)s->used[f];
            cols[f]->bytes = s->bytes[f];
            cols[f]->offsets = s->offs[f];
        }
        b.ip = s->ip;
        b.ip_is_v6 = s->v6;
        b.remote_port = s->port;
        b.flags = s->flags;
        const int rc = pgw_evaluate_batch_routed_host(q->rs, &b, s->verdict, s->service);
        for (uint32_t i = 0; i < n; ++i)
            if (s->fn[i]) s->fn[i](s->user[i], s->verdict[i], s->service[i], rc ? 4 : 0);
        lk.lock();
        s->rc = rc;
        s->done = true;
        if (s->readers_left == 0) {  // nobody blocks on this batch: the side is reusable at once
            s->free_ = true;
            q->cv_free.notify_all();
        }
        q->cv_done.notify_all();
    }
}

extern "C" {

int pgw_shape_request(const pgw_request* req, const char* out_ptr[5], size_t out_len[5]) {
    if (!req || !out_ptr || !out_len) return 1;
    Shaped sh;
    shape(req, &sh);
    for (int f = 0; f < 5; ++f) { out_ptr[f] = sh.p[f]; out_len[f] = sh.n[f]; }
    return 0;
}

int pgw_queue_create(pgw_ruleset* rs, uint32_t max_batch, uint32_t max_delay_us, pgw_queue** out, char* err, size_t err_cap) {
    auto fail = [&](const char* m) { if (err && err_cap) snprintf(err, err_cap, "%s", m); return 1; };
    if (!rs || !out) return fail("null argument");
    if (max_batch == 0 || max_batch > (1u << 22)) return fail("max_batch must be in 1..4194304");
    pgw_queue* q = new pgw_queue();
    q->rs = rs;
    q->max_batch = max_batch;
    q->max_delay = std::chrono::microseconds(max_delay_us);
    static const size_t kGuess[5] = {32, 384, 64, 8, 128};  // initial bytes per request and column (grown on demand)
    bool ok = true;
    for (Side& s : q->side) {
        for (int f = 0; f < 5 && ok; ++f) {
            ok = grow(s, f, (size_t)max_batch * kGuess[f]);
            s.offs[f] = (uint32_t*)pgw_host_alloc(((size_t)max_batch + 1) * 4);
            ok = ok && s.offs[f];
            if (ok) s.offs[f][0] = 0;
        }
        s.ip = (uint8_t*)pgw_host_alloc((size_t)max_batch * 16);
        s.v6 = (uint8_t*)pgw_host_alloc(max_batch);
        s.flags = (uint8_t*)pgw_host_alloc(max_batch);
        s.port = (int32_t*)pgw_host_alloc((size_t)max_batch * 4);
        s.verdict = (uint32_t*)pgw_host_alloc((size_t)max_batch * 4);
        s.service = (uint16_t*)pgw_host_alloc((size_t)max_batch * 2);
        s.fn = new pgw_done_fn[max_batch]();
        s.user = new void*[max_batch]();
        ok = ok && s.ip && s.v6 && s.flags && s.port && s.verdict && s.service;
    }
    if (!ok) {
        pgw_queue_destroy(q);
        return fail("pinned host allocation failed");
    }
    q->side[0].free_ = false;  // side 0 starts as the one being filled
    q->worker = std::thread(dispatcher, q);
    *out = q;
    return 0;
}

// appends one shaped request to the side being filled; returns the side and the request's index in it (lock held)
static int enqueue(pgw_queue* q, std::unique_lock<std::mutex>& lk, const pgw_request* req, pgw_done_fn fn, void* user, Side** side, uint32_t* index) {
    Shaped sh;
    shape(req, &sh);
    q->cv_space.wait(lk, [&] { return q->stop || q->side[q->fill].n < q->max_batch; });
    if (q->stop) return 2;
    Side* s = &q->side[q->fill];
    const uint32_t i = s->n;
    // every column is grown before any of them is touched: a failed allocation leaves the side exactly as it was
    for (int f = 0; f < 5; ++f)
        if (!grow(*s, f, s->used[f] + sh.n[f])) return 3;
    for (int f = 0; f < 5; ++f) {
        s->offs[f][i] = (uint32_t)s->used[f];
        if (sh.n[f]) memcpy(s->bytes[f] + s->used[f], sh.p[f], sh.n[f]);
        s->used[f] += sh.n[f];
    }
    memcpy(s->ip + (size_t)i * 16, req->ip, 16);
    s->v6[i] = req->ip_is_v6;
    s->port[i] = req->remote_port;
    s->flags[i] = req->flags;
    s->fn[i] = fn;
    s->user[i] = user;
    if (i == 0) s->first = Clock::now();
    s->n = i + 1;
    if (i == 0 || s->n == q->max_batch) q->cv_work.notify_one();
    *side = s;
    *index = i;
    return 0;
}

int pgw_queue_evaluate(pgw_queue* q, const pgw_request* req, uint32_t* verdict, uint16_t* service) {
    if (!q || !req || !verdict) return 1;
    std::unique_lock<std::mutex> lk(q->mu);
    Side* s = nullptr;
    uint32_t i = 0;
    if (int rc = enqueue(q, lk, req, nullptr, nullptr, &s, &i)) return rc;
    // the side cannot be reused before every blocked request of the batch has read its result
    q->cv_done.wait(lk, [&] { return s->done; });
    const int rc = s->rc;
    *verdict = s->verdict[i];
    if (service) *service = s->service[i];
    if (--s->readers_left == 0) {
        s->free_ = true;
        q->cv_free.notify_all();
    }
    return rc ? 4 : 0;
}

int pgw_queue_submit(pgw_queue* q, const pgw_request* req, pgw_done_fn done, void* user) {
    if (!q || !req || !done) return 1;
    std::unique_lock<std::mutex> lk(q->mu);
    Side* s = nullptr;
    uint32_t i = 0;
    return enqueue(q, lk, req, done, user, &s, &i);
}

int pgw_queue_get_stats(pgw_queue* q, pgw_queue_stats* out) {
    if (!q || !out) return 1;
    std::lock_guard<std::mutex> lk(q->mu);
    *out = q->st;
    return 0;
}

void pgw_queue_destroy(pgw_queue* q) {
    if (!q) return;
    {
        std::lock_guard<std::mutex> lk(q->mu);
        q->stop = true;
    }
    q->cv_work.notify_all();
    q->cv_space.notify_all();
    if (q->worker.joinable()) q->worker.join();
    for (Side& s : q->side) {
        for (int f = 0; f < 5; ++f) {
            if (s.bytes[f]) pgw_host_free(s.bytes[f]);
            if (s.offs[f]) pgw_host_free(s.offs[f]);
        }
        pgw_host_free(s.ip);
        pgw_host_free(s.v6);
        pgw_host_free(s.flags);
        pgw_host_free(s.port);
        pgw_host_free(s.verdict);
        pgw_host_free(s.service);
        delete[] s.fn;
        delete[] s.user;
    }
    delete q;
}

}  // extern "C"
