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
#include <utility>
#include <vector>
#include <algorithm>

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
    std::atomic<uint32_t> n{0};            // requests whose slots are reserved (written under the queue's spin lock)
    std::atomic<uint32_t> filled{0};       // requests whose bytes are in place (the copies run outside any lock)
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

// A producer thread's private mini-batch (pgw_queue_submit): requests are packed here with no shared cache line touched,
// and moved into the side being filled kLocalMax at a time -- or when the dispatcher collects what is pending at a
// deadline.  (Packing straight into the shared columns costs a dozen cache-line transfers between cores per request:
// adjacent requests of different threads share the lines of every offset / fixed-width column.)
constexpr uint32_t kLocalMax = 128;
struct Local {
    std::atomic<bool> lock{false};   // owner thread vs dispatcher
    uint32_t n = 0;
    std::vector<uint8_t> bytes[5];
    uint32_t offs[5][kLocalMax + 1];
    uint8_t ip[kLocalMax][16];
    uint8_t v6[kLocalMax], flags[kLocalMax];
    int32_t port[kLocalMax];
    pgw_done_fn fn[kLocalMax];
    void* user[kLocalMax];
    Clock::time_point first;
    void acquire() { while (lock.exchange(true, std::memory_order_acquire)) { while (lock.load(std::memory_order_relaxed)) {} } }
    void release() { lock.store(false, std::memory_order_release); }
};

std::atomic<uint64_t> g_queue_ids{1};

}  // namespace

struct pgw_queue {
    uint64_t id = 0;                       // distinguishes queues in the producers' thread-local caches
    std::vector<Local*> locals;            // every producer's mini-batch (registered under `mu`, owned by the queue)
    std::atomic<uint32_t> local_pending{0};  // mini-batches holding requests: the dispatcher must not sleep without a deadline
    pgw_ruleset* rs = nullptr;
    uint32_t max_batch = 0;
    std::chrono::microseconds max_delay{0};
    Side side[2];
    std::atomic<int> fill{0};  // index of the side being filled (changed by the dispatcher under `spin`)
    // `spin` guards the reservation state of the side being filled (n, used[], offsets, first): a producer holds it for a
    // few dozen nanoseconds per request -- a sleeping mutex hands over at about a million acquisitions per second under
    // contention, which capped the queue at ~1 M requests/s.  `mu` + the condition variables are only for sleeping.
    std::atomic<bool> spin{false};
    void spin_lock() { while (spin.exchange(true, std::memory_order_acquire)) { while (spin.load(std::memory_order_relaxed)) {} } }
    void spin_unlock() { spin.store(false, std::memory_order_release); }
    std::mutex mu;
    std::condition_variable cv_work, cv_space, cv_done, cv_free;
    std::thread worker;
    std::atomic<bool> stop{false};
    pgw_queue_stats st{};
    std::string last_error;   // message of the most recent failed batch (pgw_last_error is thread-local to the dispatcher)
};

extern "C" {
static int flush_local(pgw_queue* q, Local& L, bool from_dispatcher);
}

static void dispatcher(pgw_queue* q) {
    std::unique_lock<std::mutex> lk(q->mu);
    for (;;) {
        q->cv_work.wait(lk, [&] { return q->stop || q->side[q->fill].n.load() > 0 || q->local_pending.load() > 0; });
        if (q->stop && q->side[q->fill].n.load() == 0 && q->local_pending.load() == 0) return;
        const int cur = q->fill.load();
        Side* s = &q->side[cur];
        // the deadline counts from the oldest pending request, in the side or in a producer's mini-batch
        Clock::time_point first = Clock::now();
        q->spin_lock();
        if (s->n.load() > 0 && s->first < first) first = s->first;
        q->spin_unlock();
        std::vector<Local*> locals = q->locals;   // registered under `mu`, which is held here
        for (Local* L : locals) {
            // never wait for a mini-batch here: its owner may be waiting for THIS thread to switch sides
            if (L->lock.exchange(true, std::memory_order_acquire)) continue;
            if (L->n && L->first < first) first = L->first;
            L->release();
        }
        const bool full = q->cv_work.wait_until(lk, first + q->max_delay, [&] { return q->stop || s->n.load() == q->max_batch; }) && s->n.load() == q->max_batch;
        // collect the mini-batches; a producer that is busy with its own (try-lock fails) flushes it itself
        lk.unlock();
        for (Local* L : locals) {
            if (L->lock.exchange(true, std::memory_order_acquire)) continue;
            if (L->n) flush_local(q, *L, true);
            L->release();
        }
        lk.lock();
        if (s->n.load() == 0) continue;
        Side* other = &q->side[cur ^ 1];
        q->cv_free.wait(lk, [&] { return other->free_; });
        // the other side becomes the one being filled
        other->free_ = false;
        other->done = false;
        q->spin_lock();
        other->n.store(0);
        other->filled.store(0, std::memory_order_relaxed);
        for (int f = 0; f < 5; ++f) other->used[f] = 0;
        q->fill.store(cur ^ 1);
        const uint32_t n = s->n.load();   // final: producers reserve only in the side `fill` names
        q->spin_unlock();
        q->cv_space.notify_all();
        if (full) q->st.full_flushes++;
        else q->st.deadline_flushes++;
        q->st.batches++;
        q->st.requests += n;
        if (n > q->st.largest_batch) q->st.largest_batch = n;
        lk.unlock();
        // producers copy their bytes outside the mutex: wait for the last ones (a copy is a few hundred nanoseconds)
        while (s->filled.load(std::memory_order_acquire) != n) std::this_thread::yield();
        uint32_t blocking = 0;   // requests a thread waits for in pgw_queue_evaluate (the others have completion callbacks)
        for (uint32_t i = 0; i < n; ++i) blocking += s->fn[i] == nullptr;
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
        if (rc) q->last_error = pgw_last_error();
        s->rc = rc;
        s->readers_left = blocking;
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
    q->id = g_queue_ids.fetch_add(1);
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

// Appends one request to the side being filled; returns the side and the request's index in it.  The request is shaped
// first, its slot and byte ranges are RESERVED under the spin lock (a few dozen nanoseconds), and its bytes are copied
// with no lock held -- producers pack in parallel; the dispatcher waits for `filled` before it reads a side.
static int enqueue(pgw_queue* q, const pgw_request* req, pgw_done_fn fn, void* user, Side** side, uint32_t* index) {
    Shaped sh;
    shape(req, &sh);
    Side* s;
    uint32_t i;
    uint8_t* dst[5];
    bool wake = false;
    for (;;) {
        q->spin_lock();
        s = &q->side[q->fill.load(std::memory_order_relaxed)];
        i = s->n.load(std::memory_order_relaxed);
        if (i < q->max_batch) break;
        q->spin_unlock();
        // the side is full: sleep until the dispatcher has switched sides
        std::unique_lock<std::mutex> lk(q->mu);
        q->cv_space.wait(lk, [&] { return q->stop || q->side[q->fill.load()].n.load() < q->max_batch; });
        if (q->stop) return 2;
    }
    bool need_grow = false;
    for (int f = 0; f < 5; ++f) need_grow |= s->used[f] + sh.n[f] > s->cap[f];
    if (need_grow) {
        // a column moves: every earlier producer of this side must have finished copying into the old buffer
        while (s->filled.load(std::memory_order_acquire) != i) std::this_thread::yield();
        // every column is grown before any of them is touched: a failed allocation leaves the side exactly as it was
        for (int f = 0; f < 5; ++f)
            if (!grow(*s, f, s->used[f] + sh.n[f])) { q->spin_unlock(); return 3; }
    }
    for (int f = 0; f < 5; ++f) {
        s->offs[f][i] = (uint32_t)s->used[f];
        dst[f] = s->bytes[f] + s->used[f];
        s->used[f] += sh.n[f];
    }
    if (i == 0) s->first = Clock::now();
    s->n.store(i + 1, std::memory_order_release);
    wake = i == 0 || i + 1 == q->max_batch;
    q->spin_unlock();
    for (int f = 0; f < 5; ++f)
        if (sh.n[f]) memcpy(dst[f], sh.p[f], sh.n[f]);
    memcpy(s->ip + (size_t)i * 16, req->ip, 16);
    s->v6[i] = req->ip_is_v6;
    s->port[i] = req->remote_port;
    s->flags[i] = req->flags;
    s->fn[i] = fn;
    s->user[i] = user;
    s->filled.fetch_add(1, std::memory_order_release);
    if (wake) {
        // first request of a side, or the side is full: the dispatcher may be asleep (the empty critical section orders
        // the store of `n` before its predicate check)
        { std::lock_guard<std::mutex> g(q->mu); }
        q->cv_work.notify_one();
    }
    *side = s;
    *index = i;
    return 0;
}

// Moves the requests of mini-batch `L` (locked by the caller) into the side being filled: slots and byte ranges are
// reserved under the spin lock once per move, the bytes are copied outside it.  The dispatcher (`from_dispatcher`) never
// waits for space: what does not fit stays in the mini-batch for the next batch.
static int flush_local(pgw_queue* q, Local& L, bool from_dispatcher) {
    uint32_t k0 = 0;
    int rc = 0;
    while (k0 < L.n) {
        q->spin_lock();
        Side* s = &q->side[q->fill.load(std::memory_order_relaxed)];
        const uint32_t i = s->n.load(std::memory_order_relaxed);
        if (i >= q->max_batch) {
            q->spin_unlock();
            if (from_dispatcher) break;
            std::unique_lock<std::mutex> lk(q->mu);
            q->cv_space.wait(lk, [&] { return q->stop || q->side[q->fill.load()].n.load() < q->max_batch; });
            if (q->stop) { rc = 2; break; }
            continue;
        }
        const uint32_t m = std::min(L.n - k0, q->max_batch - i);
        size_t need[5];
        bool need_grow = false;
        for (int f = 0; f < 5; ++f) {
            need[f] = L.offs[f][k0 + m] - L.offs[f][k0];
            need_grow |= s->used[f] + need[f] > s->cap[f];
        }
        if (need_grow) {
            while (s->filled.load(std::memory_order_acquire) != i) std::this_thread::yield();
            bool ok = true;
            for (int f = 0; f < 5 && ok; ++f) ok = grow(*s, f, s->used[f] + need[f]);
            if (!ok) { q->spin_unlock(); rc = 3; break; }
        }
        uint8_t* dst[5];
        for (int f = 0; f < 5; ++f) {
            const uint32_t base = (uint32_t)s->used[f], l0 = L.offs[f][k0];
            for (uint32_t k = 0; k < m; ++k) s->offs[f][i + k] = base + (L.offs[f][k0 + k] - l0);
            dst[f] = s->bytes[f] + s->used[f];
            s->used[f] += need[f];
        }
        if (i == 0 || L.first < s->first) s->first = L.first;
        s->n.store(i + m, std::memory_order_release);
        const bool wake = i == 0 || i + m == q->max_batch;
        q->spin_unlock();
        for (int f = 0; f < 5; ++f)
            if (need[f]) memcpy(dst[f], L.bytes[f].data() + L.offs[f][k0], need[f]);
        memcpy(s->ip + (size_t)i * 16, L.ip[k0], (size_t)m * 16);
        memcpy(s->v6 + i, L.v6 + k0, m);
        memcpy(s->flags + i, L.flags + k0, m);
        memcpy(s->port + i, L.port + k0, (size_t)m * 4);
        memcpy(s->fn + i, L.fn + k0, (size_t)m * sizeof(pgw_done_fn));
        memcpy(s->user + i, L.user + k0, (size_t)m * sizeof(void*));
        s->filled.fetch_add(m, std::memory_order_release);
        if (wake) {
            { std::lock_guard<std::mutex> g(q->mu); }
            q->cv_work.notify_one();
        }
        k0 += m;
    }
    if (rc) {
        // the queue is stopping or out of memory: the requests left in the mini-batch are failed through their callbacks
        for (uint32_t k = k0; k < L.n; ++k) L.fn[k](L.user[k], 0, 0xFFFF, rc);
        k0 = L.n;
    }
    if (k0 >= L.n) {
        L.n = 0;
        for (int f = 0; f < 5; ++f) L.bytes[f].clear();
        q->local_pending.fetch_sub(1, std::memory_order_release);
    } else if (k0) {
        // the dispatcher moved a prefix: slide the rest to the front
        const uint32_t rest = L.n - k0;
        for (int f = 0; f < 5; ++f) {
            const uint32_t l0 = L.offs[f][k0];
            L.bytes[f].erase(L.bytes[f].begin(), L.bytes[f].begin() + l0);
            for (uint32_t k = 0; k <= rest; ++k) L.offs[f][k] = L.offs[f][k0 + k] - l0;
        }
        memmove(L.ip[0], L.ip[k0], (size_t)rest * 16);
        memmove(L.v6, L.v6 + k0, rest);
        memmove(L.flags, L.flags + k0, rest);
        memmove(L.port, L.port + k0, (size_t)rest * 4);
        memmove(L.fn, L.fn + k0, (size_t)rest * sizeof(pgw_done_fn));
        memmove(L.user, L.user + k0, (size_t)rest * sizeof(void*));
        L.n = rest;
    }
    return rc;
}

// this thread's mini-batch for queue `q`
static Local* local_of(pgw_queue* q) {
    thread_local std::vector<std::pair<uint64_t, Local*>> cache;
    for (auto& e : cache)
        if (e.first == q->id) return e.second;
    Local* L = new Local();
    for (int f = 0; f < 5; ++f) L->offs[f][0] = 0;
    {
        std::lock_guard<std::mutex> g(q->mu);
        q->locals.push_back(L);
    }
    cache.emplace_back(q->id, L);
    return L;
}

int pgw_queue_evaluate(pgw_queue* q, const pgw_request* req, uint32_t* verdict, uint16_t* service) {
    if (!q || !req || !verdict) return 1;
    Side* s = nullptr;
    uint32_t i = 0;
    if (int rc = enqueue(q, req, nullptr, nullptr, &s, &i)) return rc;
    std::unique_lock<std::mutex> lk(q->mu);
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
    if (q->stop.load(std::memory_order_acquire)) return 2;
    Shaped sh;
    shape(req, &sh);
    Local* L = local_of(q);
    L->acquire();
    const uint32_t k = L->n;
    for (int f = 0; f < 5; ++f) {
        L->bytes[f].insert(L->bytes[f].end(), (const uint8_t*)sh.p[f], (const uint8_t*)sh.p[f] + sh.n[f]);
        L->offs[f][k + 1] = (uint32_t)L->bytes[f].size();
    }
    memcpy(L->ip[k], req->ip, 16);
    L->v6[k] = req->ip_is_v6;
    L->flags[k] = req->flags;
    L->port[k] = req->remote_port;
    L->fn[k] = done;
    L->user[k] = user;
    L->n = k + 1;
    int rc = 0;
    bool wake = false;
    if (k == 0) {
        L->first = Clock::now();
        wake = q->local_pending.fetch_add(1, std::memory_order_acq_rel) == 0;   // the dispatcher may be asleep without a deadline
    }
    if (L->n == kLocalMax) rc = flush_local(q, *L, false);
    L->release();
    if (wake) {
        { std::lock_guard<std::mutex> g(q->mu); }
        q->cv_work.notify_one();
    }
    return rc;
}

int pgw_queue_get_stats(pgw_queue* q, pgw_queue_stats* out) {
    if (!q || !out) return 1;
    std::lock_guard<std::mutex> lk(q->mu);
    *out = q->st;
    return 0;
}

size_t pgw_queue_last_error(pgw_queue* q, char* buf, size_t cap) {
    if (!q) return 0;
    std::lock_guard<std::mutex> lk(q->mu);
    if (buf && cap) {
        const size_t n = q->last_error.size() < cap - 1 ? q->last_error.size() : cap - 1;
        memcpy(buf, q->last_error.data(), n);
        buf[n] = 0;
    }
    return q->last_error.size();
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
    for (Local* L : q->locals) delete L;
    delete q;
}

}  // extern "C"
