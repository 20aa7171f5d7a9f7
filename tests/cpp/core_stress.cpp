// Threaded stress of the CPU shard engine's C ABI, meant to be built with sanitizers:
//   g++ -std=c++17 -O1 -g -fsanitize=thread            core_stress.cpp exb_core.cpp -lpthread   (data races)
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined core_stress.cpp exb_core.cpp -lpthread   (memory / UB)
// Mirrors the reference's randomized multi-thread pull/push test (openembedding/entry/c_api_test.h:49-98:
// raw threads and pool threads pulling with duplicate keys while updates run) -- the reference has no
// sanitizer build at all. Uses the `test` optimizer, so the final weights have a closed form.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <thread>
#include <vector>

extern "C" {
void* exb_var_create(int dtype, int dim, uint64_t vocab, int shard_id, int shard_num, int use_hash);
void exb_var_destroy(void* v);
void exb_var_set_initializer(void* v, int kind, double p0, double p1, double p2, uint64_t seed);
void exb_var_set_optimizer(void* v, int kind, const double* p, int np);
void exb_var_pull(void* v, const uint64_t* keys, uint64_t n, void* out);
void exb_var_push(void* v, const uint64_t* keys, uint64_t n, const void* grads, const uint64_t* counts);
void exb_var_update(void* v);
uint64_t exb_var_num_items(void* v);
}

int main() {
    const int dim = 8, threads = 4, rounds = 50;
    const uint64_t vocab = 1000;
    int fails = 0;
    for (int use_hash = 0; use_hash < 2; ++use_hash) {
        void* v = exb_var_create(0x104, dim, use_hash ? (1ull << 63) : vocab, 0, 1, use_hash);
        exb_var_set_initializer(v, /*constant*/ 0, 0.0, 0.0, 0.0, 1);
        double p[8] = {1.0, 0.0, 0, 0, 0, 0, 0, 0};      // sgd-like: lr 1, no momentum
        exb_var_set_optimizer(v, /*OPT_SGD*/ 7, p, 3);
        std::atomic<int> go{0};
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([&, t]() {
                while (!go.load()) {}
                std::vector<uint64_t> keys(64);
                std::vector<float> out(64 * dim), g(64 * dim, 1.0f);
                std::vector<uint64_t> cnt(64, 1);
                unsigned s = 1234u + t;
                for (int r = 0; r < rounds; ++r) {
                    for (auto& k : keys) { s = s * 1664525u + 1013904223u; k = (s >> 8) % 200; }   // duplicates on purpose
                    exb_var_pull(v, keys.data(), keys.size(), out.data());       // readers ...
                    exb_var_push(v, keys.data(), keys.size(), g.data(), cnt.data());   // ... concurrent with writers
                    if (t == 0 && r % 5 == 4) exb_var_update(v);                  // ... and with the update
                }
            });
        go.store(1);
        for (auto& x : th) x.join();
        exb_var_update(v);
        // every push of key k subtracted lr * 1 per element: weights are non-positive integers
        std::vector<uint64_t> all(200);
        for (uint64_t i = 0; i < 200; ++i) all[i] = i;
        std::vector<float> w(200 * dim);
        exb_var_pull(v, all.data(), all.size(), w.data());
        double total = 0;
        for (float x : w) { if (x > 0 || x != (float)(long long)x) ++fails; total += x; }
        const double expect = -(double)threads * rounds * 64 * dim;
        if (total != expect) { fprintf(stderr, "sum %.1f != %.1f (use_hash=%d)\n", total, expect, use_hash); ++fails; }
        exb_var_destroy(v);
    }
    printf(fails ? "CORE_STRESS_FAILED %d\n" : "CORE_STRESS_OK\n", fails);
    return fails ? 1 : 0;
}
