// Parallel staging of a pageable host buffer into page-locked memory (host-only C++, no CUDA in this header).
//
// The reference's callers hold their frames in ordinary memory — std::vector<Eigen::Vector3d>, the byte vector of a
// sensor_msgs::PointCloud2 (ros/src/kinematic_icp_ros/utils/RosUtils.cpp:30-39) — and a cudaMemcpyAsync from such memory is
// staged by the driver on the calling thread.  Here a few helper threads copy the buffer into the context's own page-locked
// staging area granule by granule, in address order, while the calling thread hands every finished prefix to the copy engine:
// staging and DMA overlap, and the staging itself runs at several cores' memory bandwidth.
//
// A job's granules are claimed through one atomic counter, so the calling thread can always finish the job alone (it copies
// granules itself while it waits): with zero helpers, or helpers that wake up late, the result is the same, only slower.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace kicp {

class Stager {
public:
    struct Job {
        const unsigned char *src = nullptr;
        unsigned char *dst = nullptr;
        size_t bytes = 0, gran = 0, ngran = 0;
        std::atomic<size_t> next{0};                   // next granule to claim
        std::unique_ptr<std::atomic<uint8_t>[]> done;  // per granule: 1 once its bytes are in dst
        size_t cursor = 0;                             // calling thread only: granules [0, cursor) are known to be done
    };

    explicit Stager(int helpers) {
        for (int k = 0; k < helpers; ++k) threads_.emplace_back([this]() { helper_loop(); });
    }
    ~Stager() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    Stager(const Stager &) = delete;
    Stager &operator=(const Stager &) = delete;
    int helpers() const { return (int)threads_.size(); }

    // Begin staging src[0, bytes) into dst[0, bytes) in granules of `gran` bytes.  Both ranges must stay valid until
    // wait_prefix(job, bytes) has returned.
    std::shared_ptr<Job> start(const void *src, void *dst, size_t bytes, size_t gran) {
        auto j = std::make_shared<Job>();
        j->src = static_cast<const unsigned char *>(src), j->dst = static_cast<unsigned char *>(dst);
        j->bytes = bytes, j->gran = gran ? gran : 1, j->ngran = (bytes + j->gran - 1) / j->gran;
        j->done.reset(new std::atomic<uint8_t>[j->ngran ? j->ngran : 1]);
        for (size_t g = 0; g < j->ngran; ++g) j->done[g].store(0, std::memory_order_relaxed);
        if (!threads_.empty() && j->ngran > 1) {
            {
                std::lock_guard<std::mutex> lk(m_);
                cur_ = j, seq_++;
            }
            cv_.notify_all();
        }
        return j;
    }

    // Returns once dst[0, upto) holds the source bytes.  The calling thread copies granules itself while it waits.
    void wait_prefix(Job &j, size_t upto) {
        if (upto > j.bytes) upto = j.bytes;
        const size_t gend = (upto + j.gran - 1) / j.gran;
        while (j.cursor < gend) {
            if (j.done[j.cursor].load(std::memory_order_acquire)) {
                j.cursor++;
            } else if (!copy_one(j)) {
                relax();  // every granule is claimed: the one we wait for is being copied right now
            }
        }
    }

private:
    static void relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    static bool copy_one(Job &j) {
        const size_t g = j.next.fetch_add(1, std::memory_order_relaxed);
        if (g >= j.ngran) return false;
        const size_t off = g * j.gran, len = (off + j.gran <= j.bytes) ? j.gran : j.bytes - off;
        std::memcpy(j.dst + off, j.src + off, len);
        j.done[g].store(1, std::memory_order_release);
        return true;
    }
    void helper_loop() {
        uint64_t seen = 0;
        for (;;) {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&]() { return stop_ || seq_ != seen; });
                if (stop_) return;
                j = cur_, seen = seq_;
            }
            // (a helper that wakes up after its job has finished finds every granule claimed and touches nothing)
            while (copy_one(*j)) {
            }
        }
    }

    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    std::shared_ptr<Job> cur_;
    uint64_t seq_ = 0;
    bool stop_ = false;
};

}  // namespace kicp
