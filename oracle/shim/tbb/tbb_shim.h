// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Stand-in for the oneTBB subset the reference's Registration.cpp uses (includes at
// cpp/kinematic_icp/registration/Registration.cpp:25-31), so that file can be compiled in place for oracle/_ref
// (oneTBB is not installed offline).  A PERSISTENT pool of std::thread workers (created once, parked on a condition
// variable between jobs — like TBB's arena, no thread creation per parallel_for) over a static contiguous partition;
// the worker count is the process-wide tbb::global_control value, like the reference's function-local static cap
// (:147-148).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <iterator>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

namespace tbb {

namespace detail_shim {
inline std::atomic<size_t> &max_parallelism() {
    static std::atomic<size_t> v{0};  // 0 = hardware concurrency
    return v;
}
inline size_t workers() {
    size_t v = max_parallelism().load();
    if (v == 0) v = std::max(1u, std::thread::hardware_concurrency());
    return v;
}

// Fork-join pool: run(T, f) executes f(0) .. f(T-1), f(0) on the calling thread, the rest on parked workers.
class Pool {
public:
    static Pool &instance() {
        static Pool p;
        return p;
    }
    void run(size_t T, const std::function<void(size_t)> &f) {
        if (T <= 1) {
            f(0);
            return;
        }
        std::unique_lock<std::mutex> outer(run_mutex_);  // one job at a time (callers are single-threaded anyway)
        {
            std::unique_lock<std::mutex> g(m_);
            while (threads_.size() + 1 < T) {
                const size_t idx = threads_.size();
                threads_.emplace_back([this, idx] { worker(idx); });
            }
            job_ = &f, job_T_ = T, remaining_ = T - 1, ++generation_;
        }
        cv_job_.notify_all();
        f(0);
        std::unique_lock<std::mutex> g(m_);
        cv_done_.wait(g, [this] { return remaining_ == 0; });
        job_ = nullptr;
    }
    ~Pool() {
        {
            std::unique_lock<std::mutex> g(m_);
            stop_ = true;
        }
        cv_job_.notify_all();
        for (auto &t : threads_) t.join();
    }

private:
    void worker(size_t idx) {
        size_t seen = 0;
        for (;;) {
            const std::function<void(size_t)> *f = nullptr;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_job_.wait(g, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                if (idx + 1 < job_T_) f = job_;
            }
            if (f) {
                (*f)(idx + 1);
                std::unique_lock<std::mutex> g(m_);
                if (--remaining_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::mutex m_, run_mutex_;
    std::condition_variable cv_job_, cv_done_;
    std::vector<std::thread> threads_;
    const std::function<void(size_t)> *job_ = nullptr;
    size_t job_T_ = 0, remaining_ = 0, generation_ = 0;
    bool stop_ = false;
};
}  // namespace detail_shim

class global_control {
public:
    enum parameter { max_allowed_parallelism, thread_stack_size };
    global_control(parameter p, size_t value) {
        if (p == max_allowed_parallelism) detail_shim::max_parallelism().store(value);
    }
};

namespace this_task_arena {
inline int max_concurrency() { return static_cast<int>(std::max(1u, std::thread::hardware_concurrency())); }
}  // namespace this_task_arena

template <typename It>
class blocked_range {
public:
    using const_iterator = It;
    blocked_range(It b, It e) : b_(b), e_(e) {}
    It begin() const { return b_; }
    It end() const { return e_; }
    size_t size() const { return static_cast<size_t>(e_ - b_); }
    bool empty() const { return !(b_ < e_); }

private:
    It b_, e_;
};

template <typename Range, typename Body>
void parallel_for(const Range &range, const Body &body) {
    const size_t n = range.size();
    const size_t T = std::min(detail_shim::workers(), std::max<size_t>(n, 1));
    if (T <= 1 || n == 0) {
        body(range);
        return;
    }
    detail_shim::Pool::instance().run(T, [&](size_t t) { body(Range(range.begin() + n * t / T, range.begin() + n * (t + 1) / T)); });
}

template <typename Range, typename Value, typename Func, typename Reduction>
Value parallel_reduce(const Range &range, const Value &identity, const Func &func, const Reduction &reduction) {
    const size_t n = range.size();
    const size_t T = std::min(detail_shim::workers(), std::max<size_t>(n, 1));
    if (T <= 1 || n == 0) return func(range, identity);
    std::vector<Value> parts(T, identity);
    detail_shim::Pool::instance().run(
        T, [&](size_t t) { parts[t] = func(Range(range.begin() + n * t / T, range.begin() + n * (t + 1) / T), identity); });
    Value acc = parts[0];
    for (size_t t = 1; t < T; ++t) acc = reduction(acc, parts[t]);
    return acc;
}

// Append-only vector with a lock-free emplace_back into storage reserved up front (the reference reserves
// points.size() before the parallel fill, Registration.cpp:67-68); falls back to a mutex when it has to grow.
template <typename T>
class concurrent_vector {
public:
    using value_type = T;
    using const_iterator = const T *;
    using iterator = T *;
    concurrent_vector() = default;
    concurrent_vector(const concurrent_vector &o) { assign(o); }
    concurrent_vector(concurrent_vector &&o) noexcept { steal(o); }
    concurrent_vector &operator=(const concurrent_vector &o) {
        if (this != &o) {
            destroy();
            assign(o);
        }
        return *this;
    }
    concurrent_vector &operator=(concurrent_vector &&o) noexcept {
        if (this != &o) {
            destroy();
            steal(o);
        }
        return *this;
    }
    ~concurrent_vector() { destroy(); }

    void reserve(size_t n) {
        std::lock_guard<std::mutex> g(m_);
        grow_locked(n);
    }
    template <typename... Args>
    iterator emplace_back(Args &&...args) {
        size_t idx = size_.fetch_add(1, std::memory_order_relaxed);
        if (idx >= cap_.load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> g(m_);
            if (idx >= cap_.load()) grow_locked(std::max<size_t>(2 * cap_.load(), idx + 1));
        }
        return new (data_ + idx) T(std::forward<Args>(args)...);
    }
    size_t size() const { return size_.load(); }
    bool empty() const { return size() == 0; }
    const_iterator cbegin() const { return data_; }
    const_iterator cend() const { return data_ + size(); }
    const_iterator begin() const { return data_; }
    const_iterator end() const { return data_ + size(); }
    const T &operator[](size_t i) const { return data_[i]; }

private:
    // NOTE: growth while other threads emplace is not safe in this shim; the reference always reserves first.
    void grow_locked(size_t n) {
        if (n <= cap_.load()) return;
        T *nd = static_cast<T *>(::operator new(n * sizeof(T)));
        const size_t s = std::min(size_.load(), cap_.load());
        for (size_t i = 0; i < s; ++i) {
            new (nd + i) T(std::move(data_[i]));
            data_[i].~T();
        }
        ::operator delete(data_);
        data_ = nd;
        cap_.store(n, std::memory_order_release);
    }
    void destroy() {
        const size_t s = std::min(size_.load(), cap_.load());
        for (size_t i = 0; i < s; ++i) data_[i].~T();
        ::operator delete(data_);
        data_ = nullptr;
        size_.store(0);
        cap_.store(0);
    }
    void assign(const concurrent_vector &o) {
        const size_t s = o.size();
        data_ = s ? static_cast<T *>(::operator new(s * sizeof(T))) : nullptr;
        for (size_t i = 0; i < s; ++i) new (data_ + i) T(o.data_[i]);
        size_.store(s);
        cap_.store(s);
    }
    void steal(concurrent_vector &o) {
        data_ = o.data_;
        size_.store(o.size_.load());
        cap_.store(o.cap_.load());
        o.data_ = nullptr;
        o.size_.store(0);
        o.cap_.store(0);
    }
    T *data_ = nullptr;
    std::atomic<size_t> size_{0}, cap_{0};
    std::mutex m_;
};

}  // namespace tbb
