// CPU-only stress of the micro-batching queue (pingoo_b200/csrc/queue.cpp) under ThreadSanitizer: the device entry points
// are stubbed (verdict = a checksum of the packed request, computed from the batch the queue assembled), so that lost,
// duplicated or mixed-up requests, early buffer reuse and data races in the queue itself are caught without a GPU.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pingoo_waf.h"

static uint32_t fnv(const uint8_t* p, size_t n, uint32_t h) {
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 16777619u;
    return h;
}

static std::atomic<int> g_eval_calls{0};

extern "C" {
void* pgw_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void pgw_host_free(void* p) { free(p); }
const char* pgw_last_error(void) { return "stub"; }
// stub of the batch evaluation: checks the batch is well formed and derives verdict/service from the packed bytes
int pgw_evaluate_batch_routed_host(pgw_ruleset*, const pgw_batch* b, uint32_t* verdict, uint16_t* service) {
    g_eval_calls++;
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    for (uint32_t r = 0; r < b->n; ++r) {
        uint32_t h = 2166136261u;
        for (int f = 0; f < 5; ++f) {
            if (cols[f]->offsets[r + 1] < cols[f]->offsets[r]) return 1;
            h = fnv(cols[f]->bytes + cols[f]->offsets[r], cols[f]->offsets[r + 1] - cols[f]->offsets[r], h);
        }
        h = fnv(b->ip + (size_t)r * 16, 16, h);
        h = fnv((const uint8_t*)&b->remote_port[r], 4, h) ^ b->flags[r] ^ ((uint32_t)b->ip_is_v6[r] << 8);
        verdict[r] = h;
        service[r] = (uint16_t)(h >> 16);
    }
    std::this_thread::sleep_for(std::chrono::microseconds(50 + (b->n % 7) * 20));  // an "evaluation" takes a while
    return 0;
}
}

struct Req {
    std::string host, url, path, method, ua;
    pgw_request r;
    uint32_t want;
};

static Req make(uint32_t i) {
    Req q;
    q.host = "h" + std::to_string(i % 97) + ".example";
    q.url = "/p/" + std::to_string(i) + std::string(i % 300, 'x');
    q.path = "/p/" + std::to_string(i);
    q.method = (i % 5) ? "GET" : "POST";
    q.ua = "agent/" + std::to_string(i % 13);
    memset(&q.r, 0, sizeof q.r);
    q.r.host = q.host.data(); q.r.host_len = q.host.size();
    q.r.url = q.url.data(); q.r.url_len = q.url.size();
    q.r.path = q.path.data(); q.r.path_len = q.path.size();
    q.r.method = q.method.data(); q.r.method_len = q.method.size();
    q.r.user_agent = q.ua.data(); q.r.user_agent_len = q.ua.size();
    q.r.ip[0] = 10; q.r.ip[3] = (uint8_t)i;
    q.r.remote_port = (int32_t)(1000 + i % 5000);
    q.r.flags = (uint8_t)(i % 3);
    uint32_t h = 2166136261u;
    h = fnv((const uint8_t*)q.host.data(), q.host.size(), h);
    h = fnv((const uint8_t*)q.url.data(), q.url.size(), h);
    h = fnv((const uint8_t*)q.path.data(), q.path.size(), h);
    h = fnv((const uint8_t*)q.method.data(), q.method.size(), h);
    h = fnv((const uint8_t*)q.ua.data(), q.ua.size(), h);
    h = fnv(q.r.ip, 16, h);
    h = fnv((const uint8_t*)&q.r.remote_port, 4, h) ^ q.r.flags;
    q.want = h;
    return q;
}

static std::atomic<int> g_bad{0}, g_done{0};
static void on_done(void* user, uint32_t verdict, uint16_t service, int rc) {
    const Req* q = (const Req*)user;
    if (rc || verdict != q->want || service != (uint16_t)(q->want >> 16)) g_bad++;
    g_done++;
}

int main(int argc, char** argv) {
    const uint32_t max_batch = argc > 1 ? (uint32_t)atoi(argv[1]) : 64;
    const uint32_t delay_us = argc > 2 ? (uint32_t)atoi(argv[2]) : 300;
    pgw_queue* q = nullptr;
    char err[128];
    if (pgw_queue_create((pgw_ruleset*)0x1, max_batch, delay_us, &q, err, sizeof err)) { printf("create failed: %s\n", err); return 2; }
    const uint32_t T = 8, PER = 3000;
    std::vector<std::vector<Req>> reqs(T);
    for (uint32_t t = 0; t < T; ++t) {
        reqs[t].reserve(PER);  // the pgw_request points into the strings: no reallocation afterwards
        for (uint32_t k = 0; k < PER; ++k) reqs[t].push_back(make(t * PER + k));
        for (auto& x : reqs[t]) {  // re-point after the moves of push_back
            x.r.host = x.host.data(); x.r.url = x.url.data(); x.r.path = x.path.data(); x.r.method = x.method.data(); x.r.user_agent = x.ua.data();
        }
    }
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            for (uint32_t k = 0; k < PER; ++k) {
                Req& x = reqs[t][k];
                if ((t + k) % 3 == 0) {  // blocking and callback submissions mixed in the same batches
                    uint32_t v = 0;
                    uint16_t s = 0;
                    int rc = pgw_queue_evaluate(q, &x.r, &v, &s);
                    if (rc || v != x.want || s != (uint16_t)(x.want >> 16)) g_bad++;
                    g_done++;
                } else if (pgw_queue_submit(q, &x.r, on_done, &x)) {
                    g_bad++;
                    g_done++;
                }
            }
        });
    for (auto& x : th) x.join();
    while (g_done.load() < (int)(T * PER)) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    pgw_queue_stats st;
    pgw_queue_get_stats(q, &st);
    pgw_queue_destroy(q);
    printf("requests %llu batches %llu full %llu deadline %llu largest %u eval_calls %d bad %d\n", (unsigned long long)st.requests,
           (unsigned long long)st.batches, (unsigned long long)st.full_flushes, (unsigned long long)st.deadline_flushes, st.largest_batch,
           g_eval_calls.load(), g_bad.load());
    return (g_bad.load() == 0 && st.requests == T * PER && st.largest_batch <= max_batch) ? 0 : 1;
}
