// Host-only rate of the micro-batching queue (pingoo_b200/csrc/queue.cpp) with the device entry point stubbed (300 us per batch):
// separates the packing / hand-off cost from the CUDA path.  tools only.
//   g++ -O2 -std=c++17 -pthread -o queue_microbench queue_microbench.cpp ../pingoo_b200/csrc/queue.cpp && ./queue_microbench <threads> <requests per thread>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <string>
#include "../include/pingoo_waf.h"
extern "C" {
void* pgw_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void pgw_host_free(void* p) { free(p); }
static std::atomic<long> g_batches{0}, g_reqs{0};
int pgw_evaluate_batch_routed_host(pgw_ruleset*, const pgw_batch* b, uint32_t* verdict, uint16_t* service) {
    g_batches++; g_reqs += b->n;
    std::this_thread::sleep_for(std::chrono::microseconds(300));
    for (uint32_t r = 0; r < b->n; ++r) { verdict[r] = r; service[r] = 0; }
    return 0;
}
const char* pgw_last_error(void) { return ""; }
}
static std::atomic<long> done_cnt{0};
static void on_done(void* u, uint32_t, uint16_t, int) { ((std::atomic<int>*)u)->fetch_sub(1, std::memory_order_release); done_cnt++; }
int main(int argc, char** argv) {
    int threads = argc > 1 ? atoi(argv[1]) : 8; int per = argc > 2 ? atoi(argv[2]) : 400000; int window = 8192;
    pgw_queue* q = nullptr; char err[256];
    if (pgw_queue_create((pgw_ruleset*)1, 16384, 500, &q, err, sizeof err)) { puts(err); return 1; }
    std::string url(297, 'a'), ua(113, 'M'), path(59, 'p'), host(18, 'h');
    std::vector<std::atomic<int>> infl(threads); for (auto& x : infl) x = 0;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([&, t] {
        for (int k = 0; k < per; ++k) {
            while (infl[t].load(std::memory_order_acquire) >= window) std::this_thread::yield();
            pgw_request r; memset(&r, 0, sizeof r);
            r.host = host.data(); r.host_len = host.size(); r.url = url.data(); r.url_len = url.size(); r.path = path.data(); r.path_len = path.size();
            r.method = "GET"; r.method_len = 3; r.user_agent = ua.data(); r.user_agent_len = ua.size();
            infl[t].fetch_add(1);
            if (pgw_queue_submit(q, &r, on_done, &infl[t])) { puts("fail"); }
        }
        while (infl[t].load() > 0) std::this_thread::yield();
    });
    for (auto& x : th) x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %d: %.2f M req/s, batches %ld avg %.0f\n", threads, threads * (double)per / s / 1e6, g_batches.load(), (double)g_reqs / g_batches);
    pgw_queue_destroy(q);
}
