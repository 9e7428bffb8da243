// include/dvfe/frontend_io.hpp's FeatureQueue against the reference's own FeatureQueue (basic/feature_queue.h:19-71, compiled into
// oracle/_ref/libdvref.so and reached through its C glue): the same operation script runs on both, every observable result must be
// equal.  usage: test_queue_vs_reference <path to libdvref.so>
// The class under test is renamed in this translation unit so that its inline members can never be confused with the
// reference's same-named ones inside the loaded library.
#include <dlfcn.h>

#include <cstdio>
#include <thread>

#define FeatureQueue DvfeFeatureQueue
#include "dvfe/frontend_io.hpp"
#undef FeatureQueue

using namespace dynamic_vins;

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #cond); return 1; } \
    } while (0)

template <class F> static F sym(void* h, const char* name) { return reinterpret_cast<F>(dlsym(h, name)); }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    auto q_new = sym<void* (*)()>(h, "dvref_queue_new");
    auto q_free = sym<void (*)(void*)>(h, "dvref_queue_free");
    auto q_push = sym<void (*)(void*, unsigned, double)>(h, "dvref_queue_push");
    auto q_request = sym<int (*)(void*, unsigned*, double*)>(h, "dvref_queue_request");
    auto q_size = sym<int (*)(void*)>(h, "dvref_queue_size");
    auto q_empty = sym<int (*)(void*)>(h, "dvref_queue_empty");
    auto q_clear = sym<void (*)(void*)>(h, "dvref_queue_clear");
    auto q_front = sym<int (*)(void*, double*)>(h, "dvref_queue_front_time");
    CHECK(q_new && q_free && q_push && q_request && q_size && q_empty && q_clear && q_front);
    void* ref = q_new();
    DvfeFeatureQueue mine;

    auto push = [&](unsigned seq) {
        FrontendFeature f;
        f.seq_id = seq; f.time = 10.0 + 0.05 * seq;
        mine.push_back(f);
        q_push(ref, seq, f.time);
    };
    auto same_state = [&]() {
        double t = 0;
        const int has = q_front(ref, &t);
        auto ft = mine.front_time();
        return mine.size() == q_size(ref) && (mine.empty() ? 1 : 0) == q_empty(ref) && has == (ft.has_value() ? 1 : 0) && (!has || *ft == t);
    };
    auto request_both = [&]() {          // returns -2 on a mismatch, -1 when both time out, else the sequence number both delivered
        unsigned seq = 0; double t = 0;
        const int got = q_request(ref, &seq, &t);
        auto f = mine.request();
        if ((got != 0) != f.has_value()) return -2;
        if (!got) return -1;
        return (f->seq_id == seq && f->time == t) ? (int)seq : -2;
    };

    CHECK(same_state());
    CHECK(request_both() == -1);                                   // both wait 30 ms on an empty queue and give up
    for (unsigned k = 0; k < 7; k++) push(k);
    CHECK(same_state());
    for (int k = 0; k < 3; k++) CHECK(request_both() == k);        // FIFO
    CHECK(same_state());
    for (unsigned k = 7; k < 140; k++) { push(k); CHECK(same_state()); }      // the bound: frames beyond kImageQueueSize are dropped
    CHECK(mine.size() == kImageQueueSize);
    int last = 2;
    for (;;) {
        const int r = request_both();
        CHECK(r != -2);
        if (r == -1) break;
        CHECK(r > last);
        last = r;
        CHECK(same_state());
    }
    CHECK(last == 102);                                            // 3..102 were queued (100 slots), 103..139 dropped
    push(500); push(501);
    mine.clear(); q_clear(ref);
    CHECK(same_state() && request_both() == -1);
    q_free(ref);
    std::printf("queue equals the reference's FeatureQueue\n");
    return 0;
}
