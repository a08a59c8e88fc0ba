// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the one Ygor symbol that the reference's
// SYCL_Fallback.h needs (`work_queue<T>`: ctor(n), submit_task(T), joining dtor; used at
// /root/reference/src/SYCL_Fallback.h:220,251,354).  Ygor itself (github.com/hdclark/Ygor, branch master,
// unpinned by cmake/UpstreamAuthorDeps.cmake:71-73) is not vendored in the reference and is not installed here.
// This is our own code, written from the usage sites; it holds no arithmetic.
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

template <class T>
class work_queue {
  public:
    explicit work_queue(unsigned int n_workers = std::thread::hardware_concurrency()) {
        if (n_workers == 0U) n_workers = 2U;
        for (unsigned int i = 0; i < n_workers; ++i) {
            workers_.emplace_back([this]() { this->run(); });
        }
    }
    work_queue(const work_queue &) = delete;
    work_queue &operator=(const work_queue &) = delete;

    void submit_task(T task) {
        {
            std::lock_guard<std::mutex> lk(m_);
            tasks_.push_back(std::move(task));
        }
        cv_.notify_one();
    }

    ~work_queue() {
        {
            std::lock_guard<std::mutex> lk(m_);
            closing_ = true;
        }
        cv_.notify_all();
        for (auto &w : workers_) {
            if (w.joinable()) w.join();
        }
    }

  private:
    void run() {
        for (;;) {
            T task;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this]() { return closing_ || !tasks_.empty(); });
                if (tasks_.empty()) return;  // closing and drained
                task = std::move(tasks_.front());
                tasks_.pop_front();
            }
            task();
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> tasks_;
    std::vector<std::thread> workers_;
    bool closing_ = false;
};
