// Same name and interface as the reference's include/threadsafequeue.h:9-76 (ThreadSafeQueue<T>: push / wait_and_pop /
// prepare), so that /root/reference/src/project.cpp:28 and code written against the reference compile unchanged.
//
// Contract kept (it is what the three thread bodies rely on):
//   * one producer, two consumers (indicator 0 = audio thread, 1 = RDS thread), ONE payload in flight;
//   * push(v) blocks until both consumers have called prepare() for the previous payload (both count as prepared
//     before the first push), then deletes the previous payload and publishes v;
//   * wait_and_pop(v, who) blocks until a payload is published that consumer `who` has not taken yet;
//   * prepare(who): consumer `who` no longer needs the payload (the producer may replace and delete it).
// Own implementation: a generation counter instead of the reference's flag set.
#pragma once

#include <condition_variable>
#include <mutex>

template <typename T>
class ThreadSafeQueue {
public:
    ThreadSafeQueue() = default;
    ThreadSafeQueue(const ThreadSafeQueue&) = delete;
    ThreadSafeQueue& operator=(const ThreadSafeQueue&) = delete;

    void push(const T value) {
        std::unique_lock<std::mutex> lk(m_);
        producer_.wait(lk, [&] { return released_[0] == generation_ && released_[1] == generation_; });
        if (payload_) delete payload_;
        payload_ = value;
        ++generation_;
        consumers_.notify_all();
    }

    void wait_and_pop(T& value, int indicator) {
        std::unique_lock<std::mutex> lk(m_);
        const int who = indicator ? 1 : 0;
        consumers_.wait(lk, [&] { return taken_[who] != generation_; });
        value = payload_;
        taken_[who] = generation_;
    }

    void prepare(int indicator) {
        std::lock_guard<std::mutex> lk(m_);
        if (indicator == 0 || indicator == 1) released_[indicator] = taken_[indicator];
        producer_.notify_all();
    }

private:
    std::mutex m_;
    std::condition_variable consumers_, producer_;
    T payload_ = nullptr;
    unsigned long long generation_ = 0;          // number of payloads published so far
    unsigned long long taken_[2] = {0, 0};       // generation each consumer has popped
    unsigned long long released_[2] = {0, 0};    // generation each consumer has released with prepare()
};
