// c2d_copy_pool.hpp — helper threads for the copy-out of Registry::getCollidingPairs().
//
// The reference returns its collisions in a fresh std::vector (Registry.hpp:464-549); at 1 M shapes that is ~35 MB of
// first-touch memory per frame, which one host thread fills at ~8 GB/s — four times the whole device pipeline.  The
// pool lets the library (a) touch the pages of the caller's reserved storage while the GPU is still working and
// (b) move the records out of the pinned staging buffer chunk by chunk, as the device-to-host copies land.
#pragma once

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace c2d_host
{

class CopyPool
{
  public:
    explicit CopyPool(unsigned helpers)
    {
        for (unsigned i = 0; i < helpers; ++i)
            mWorkers.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lock(mMutex);
            mStop = true;
        }
        mWork.notify_all();
        for (std::thread& t : mWorkers)
            t.join();
    }
    CopyPool(const CopyPool&) = delete;
    CopyPool& operator=(const CopyPool&) = delete;

    unsigned helpers() const
    {
        return static_cast<unsigned>(mWorkers.size());
    }

    // Starts fn(helperIndex) on every helper thread and returns; one job at a time (waits for the previous one).
    void launch(std::function<void(unsigned)> fn)
    {
        std::unique_lock<std::mutex> lock(mMutex);
        mDone.wait(lock, [this] { return mPending == 0; });
        if (mWorkers.empty())
            return;
        mJob = std::move(fn);
        mPending = static_cast<unsigned>(mWorkers.size());
        ++mGeneration;
        lock.unlock();
        mWork.notify_all();
    }

    void wait()
    {
        std::unique_lock<std::mutex> lock(mMutex);
        mDone.wait(lock, [this] { return mPending == 0; });
    }

  private:
    void loop(unsigned index)
    {
        uint64_t seen = 0;
        while (true)
        {
            std::function<void(unsigned)> job;
            {
                std::unique_lock<std::mutex> lock(mMutex);
                mWork.wait(lock, [&] { return mStop || mGeneration != seen; });
                if (mStop)
                    return;
                seen = mGeneration;
                job = mJob;
            }
            job(index);
            {
                std::lock_guard<std::mutex> lock(mMutex);
                if (--mPending == 0)
                    mDone.notify_all();
            }
        }
    }

    std::vector<std::thread> mWorkers;
    std::mutex mMutex;
    std::condition_variable mWork, mDone;
    std::function<void(unsigned)> mJob;
    uint64_t mGeneration = 0;
    unsigned mPending = 0;
    bool mStop = false;
};

} // namespace c2d_host
