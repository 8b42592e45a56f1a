// ORACLE BUILD SHIM: std::thread implementation of the reference's
// TaskSchedulerManager (header: BasicRenderer/include/Managers/Singletons/TaskSchedulerManager.h).
// The reference implements ParallelForImpl on a oneTBB arena sized to
// hardware_concurrency (TaskSchedulerManager.cpp:67-75, 328-390); oneTBB is not
// in this image, so this shim shares work over an atomic counter instead.
// Thread count can be overridden with CLODREF_THREADS.
#include "Managers/Singletons/TaskSchedulerManager.h"
#include <cstdlib>

namespace br {

struct TaskSchedulerManager::RuntimeState {};

TaskSchedulerManager& TaskSchedulerManager::GetInstance() {
    static TaskSchedulerManager instance;
    return instance;
}

void TaskSchedulerManager::Initialize(uint32_t, uint32_t) {
    uint32_t n = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("CLODREF_THREADS")) n = (uint32_t)std::atoi(e);
    m_workerThreadCount = n ? n : 1u;
    m_initialized = true;
}

void TaskSchedulerManager::Cleanup() {
    m_initialized = false;
    m_workerThreadCount = 0;
}

void TaskSchedulerManager::IoWorkerLoop() {}
void TaskSchedulerManager::BackgroundWorkerLoop() {}
void TaskSchedulerManager::RunIoTask(std::function<void()>&& t) { t(); }
void TaskSchedulerManager::RunIoTask(std::string_view, std::function<void()>&& t) { t(); }
void TaskSchedulerManager::QueueIoTask(std::function<void()>&& t) { t(); }
void TaskSchedulerManager::QueueIoTask(std::string_view, std::function<void()>&& t) { t(); }
void TaskSchedulerManager::RunBackgroundTask(std::function<void()>&& t) { t(); }
void TaskSchedulerManager::RunBackgroundTask(std::string_view, std::function<void()>&& t) { t(); }

void TaskSchedulerManager::ParallelForImpl(std::string_view, size_t itemCount, std::function<void(size_t)>&& func) {
    uint32_t nthreads = m_initialized ? m_workerThreadCount : 1u;
    if (nthreads <= 1 || itemCount <= 1) {
        for (size_t i = 0; i < itemCount; ++i) func(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex errMutex;
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= itemCount) break;
            try { func(i); } catch (...) { std::lock_guard<std::mutex> l(errMutex); if (!err) err = std::current_exception(); }
        }
    };
    std::vector<std::thread> threads;
    size_t n = std::min<size_t>(nthreads, itemCount);
    for (size_t t = 1; t < n; ++t) threads.emplace_back(worker);
    worker();
    for (auto& t : threads) t.join();
    if (err) std::rethrow_exception(err);
}

}
