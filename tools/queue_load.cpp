// Load generator for the micro-batching queue (measurement tool, not part of the product library):
// `n_threads` C++ threads replay the requests of a host pgw_batch through pgw_queue_evaluate as fast as they can.
// Built by tools/queue_bench.py into tools/libqueue_load.so.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/pingoo_waf.h"

extern "C" int queue_load_run(pgw_queue* q, const pgw_batch* b, uint32_t n_threads, uint32_t per_thread, uint32_t* verdict_out,
                              double* seconds, double* lat_us_p50, double* lat_us_p99, double* lat_us_max) {
    using Clock = std::chrono::steady_clock;
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    std::vector<std::vector<float>> lats(n_threads);
    std::atomic<int> failures{0};
    auto t0 = Clock::now();
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            lats[t].reserve(per_thread);
            for (uint32_t k = 0; k < per_thread; ++k) {
                const uint32_t i = (uint32_t)(((uint64_t)t * per_thread + k) % b->n);
                pgw_request r;
                memset(&r, 0, sizeof r);
                const char** ps[5] = {&r.host, &r.url, &r.path, &r.method, &r.user_agent};
                size_t* ls[5] = {&r.host_len, &r.url_len, &r.path_len, &r.method_len, &r.user_agent_len};
                for (int f = 0; f < 5; ++f) {
                    *ps[f] = (const char*)cols[f]->bytes + cols[f]->offsets[i];
                    *ls[f] = cols[f]->offsets[i + 1] - cols[f]->offsets[i];
                }
                memcpy(r.ip, b->ip + (size_t)i * 16, 16);
                r.ip_is_v6 = b->ip_is_v6[i];
                r.remote_port = b->remote_port ? b->remote_port[i] : 0;
                r.flags = b->flags ? b->flags[i] : 0;
                uint32_t v = 0;
                auto a = Clock::now();
                if (pgw_queue_evaluate(q, &r, &v, nullptr)) failures++;
                lats[t].push_back(std::chrono::duration<float, std::micro>(Clock::now() - a).count());
                if (verdict_out && (uint64_t)t * per_thread + k < b->n) verdict_out[i] = v;
            }
        });
    for (auto& x : th) x.join();
    *seconds = std::chrono::duration<double>(Clock::now() - t0).count();
    std::vector<float> all;
    for (auto& l : lats) all.insert(all.end(), l.begin(), l.end());
    std::sort(all.begin(), all.end());
    if (!all.empty()) {
        *lat_us_p50 = all[all.size() / 2];
        *lat_us_p99 = all[(size_t)(all.size() * 0.99)];
        *lat_us_max = all.back();
    }
    return failures.load();
}


// Asynchronous variant: every thread keeps up to `window` requests in flight through pgw_queue_submit (the shape of an
// async server with many connections per worker); completions arrive on the queue's dispatcher thread.
namespace {
struct Ctx {
    std::chrono::steady_clock::time_point t0;
    std::atomic<int>* inflight;
    float* lat;
    uint32_t* vout;
    std::atomic<int>* failures;
};
void on_done(void* user, uint32_t verdict, uint16_t, int rc) {
    Ctx* c = (Ctx*)user;
    *c->lat = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - c->t0).count();
    if (c->vout) *c->vout = verdict;
    if (rc) c->failures->fetch_add(1);
    c->inflight->fetch_sub(1, std::memory_order_release);
}
}  // namespace

extern "C" int queue_load_run_async(pgw_queue* q, const pgw_batch* b, uint32_t n_threads, uint32_t per_thread, uint32_t window, uint32_t* verdict_out,
                                    double* seconds, double* lat_us_p50, double* lat_us_p99, double* lat_us_max) {
    using Clock = std::chrono::steady_clock;
    const pgw_strcol* cols[5] = {&b->host, &b->url, &b->path, &b->method, &b->user_agent};
    std::vector<std::vector<float>> lats(n_threads, std::vector<float>(per_thread, 0.f));
    std::vector<std::vector<Ctx>> ctxs(n_threads, std::vector<Ctx>(per_thread));
    std::vector<std::atomic<int>> inflight(n_threads);
    std::atomic<int> failures{0};
    for (auto& x : inflight) x.store(0);
    auto t0 = Clock::now();
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            for (uint32_t k = 0; k < per_thread; ++k) {
                while (inflight[t].load(std::memory_order_acquire) >= (int)window) std::this_thread::yield();
                const uint64_t g = (uint64_t)t * per_thread + k;
                const uint32_t i = (uint32_t)(g % b->n);
                pgw_request r;
                memset(&r, 0, sizeof r);
                const char** ps[5] = {&r.host, &r.url, &r.path, &r.method, &r.user_agent};
                size_t* ls[5] = {&r.host_len, &r.url_len, &r.path_len, &r.method_len, &r.user_agent_len};
                for (int f = 0; f < 5; ++f) {
                    *ps[f] = (const char*)cols[f]->bytes + cols[f]->offsets[i];
                    *ls[f] = cols[f]->offsets[i + 1] - cols[f]->offsets[i];
                }
                memcpy(r.ip, b->ip + (size_t)i * 16, 16);
                r.ip_is_v6 = b->ip_is_v6[i];
                r.remote_port = b->remote_port ? b->remote_port[i] : 0;
                r.flags = b->flags ? b->flags[i] : 0;
                Ctx& c = ctxs[t][k];
                c.t0 = Clock::now();
                c.inflight = &inflight[t];
                c.lat = &lats[t][k];
                c.vout = (verdict_out && g < b->n) ? verdict_out + i : nullptr;
                c.failures = &failures;
                inflight[t].fetch_add(1);
                if (pgw_queue_submit(q, &r, on_done, &c)) { failures++; inflight[t].fetch_sub(1); }
            }
            while (inflight[t].load(std::memory_order_acquire) > 0) std::this_thread::yield();
        });
    for (auto& x : th) x.join();
    *seconds = std::chrono::duration<double>(Clock::now() - t0).count();
    std::vector<float> all;
    for (auto& l : lats) all.insert(all.end(), l.begin(), l.end());
    std::sort(all.begin(), all.end());
    if (!all.empty()) {
        *lat_us_p50 = all[all.size() / 2];
        *lat_us_p99 = all[(size_t)(all.size() * 0.99)];
        *lat_us_max = all.back();
    }
    return failures.load();
}
