// hostcopy.cu -- host <-> device staging engine of the host tier.
//
// The reference API hands over and returns ordinary (pageable) NumPy arrays; for
// cfg2 the horizon array alone is 2 GB in each direction (D2H after the search,
// H2D again for sky_view_factor).  cudaMemcpy on pageable memory moves that with
// one driver thread (and takes the first-touch page faults of a fresh ndarray on
// the same thread): 4-5 GB/s, i.e. as long as the traversal itself.  This engine
//   * pre-faults fresh output pages with MADV_POPULATE_WRITE from several threads
//     while the kernel runs,
//   * moves data through a small ring of pinned buffers: DMA of chunk i overlaps the
//     multi-threaded memcpy between the pinned buffer and the caller's array of
//     chunk i-1 (D2H) / i+1 (H2D).
// One process-wide instance (worker threads + pinned ring), serialised by a mutex.
#include "hzb_common.cuh"
#include <sys/mman.h>
#include <unistd.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif

namespace hzb {
namespace {

class Workers {   // persistent pool: run(f, n) executes f(0..n-1) on the pool and the caller
 public:
    Workers() {
        unsigned hc = std::thread::hardware_concurrency();
        int n = (int)(hc ? hc : 8) / 2;
        const char* lw = getenv("LOCAL_WORLD_SIZE");     // one process per GPU (torchrun): share the host cores
        if (lw && atoi(lw) > 1) n /= atoi(lw);
        n = std::max(2, std::min(n, 12));
        const char* e = getenv("HZB_HOST_THREADS");
        if (e && atoi(e) > 0) n = std::min(atoi(e), 64);
        for (int i = 0; i + 1 < n; ++i) th_.emplace_back([this] { loop(); });
        nthreads_ = n;
    }
    ~Workers() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return nthreads_; }
    void run(const std::function<void(int)>& f, int parts) {
        if (parts <= 0) return;
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = &f; parts_ = parts; next_.store(0); left_ = parts; ++gen_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return left_ == 0 && active_ == 0; });   // no worker is still inside work()
        job_ = nullptr;
    }

 private:
    void work() {
        while (true) {
            const int i = next_.fetch_add(1);
            if (i >= parts_) break;
            (*job_)(i);
            std::lock_guard<std::mutex> l(m_);
            if (--left_ == 0) done_.notify_all();
        }
    }
    void loop() {
        unsigned long long seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || (gen_ != seen && job_ != nullptr); });
                if (stop_) return;
                seen = gen_;
                ++active_;
            }
            work();
            {
                std::lock_guard<std::mutex> l(m_);
                if (--active_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr;
    std::atomic<int> next_{0};
    int parts_ = 0, left_ = 0, nthreads_ = 1, active_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

constexpr int RING = 3;
constexpr size_t CHUNK = (size_t)32 << 20;

struct Engine {
    std::mutex mu;
    Workers pool;
    void* pin[RING] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[RING] = {nullptr, nullptr, nullptr};
    bool ready = false;
    int dev = -1;          // device the events belong to
    int init() {
        int cur = 0;
        HZB_CUDA(cudaGetDevice(&cur));
        if (!ready) {
            for (int i = 0; i < RING; ++i) HZB_CUDA(cudaHostAlloc(&pin[i], CHUNK, cudaHostAllocPortable));
            ready = true;
        }
        if (cur != dev) {   // events are per device; the pinned ring is portable
            for (int i = 0; i < RING; ++i) {
                if (ev[i]) cudaEventDestroy(ev[i]);
                HZB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
            }
            dev = cur;
        }
        return 0;
    }
    void pcopy(void* dst, const void* src, size_t bytes) {
        const int parts = (int)std::min<size_t>((size_t)pool.size(), std::max<size_t>(1, bytes >> 20));
        const size_t per = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
        pool.run([&](int i) {
            const size_t o = (size_t)i * per;
            if (o < bytes) memcpy((char*)dst + o, (const char*)src + o, std::min(per, bytes - o));
        }, parts);
    }
};

Engine& engine() { static Engine* e = new Engine(); return *e; }   // leaked on purpose: no CUDA calls at exit

}  // namespace

void host_prefault(void* p, size_t bytes) {
    if (!p || bytes < ((size_t)4 << 20) || getenv("HZB_NO_PREFAULT")) return;
    Engine& e = engine();
    std::lock_guard<std::mutex> l(e.mu);
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const uintptr_t a0 = (uintptr_t)p & ~(uintptr_t)(page - 1), a1 = ((uintptr_t)p + bytes + page - 1) & ~(uintptr_t)(page - 1);
    const size_t slice = (size_t)16 << 20;
    const int parts = (int)((a1 - a0 + slice - 1) / slice);
    e.pool.run([&](int i) {
        const uintptr_t b = a0 + (uintptr_t)i * slice, t = std::min<uintptr_t>(b + slice, a1);
        madvise((void*)b, t - b, MADV_POPULATE_WRITE);   // content untouched; an old kernel just returns EINVAL
    }, parts);
}

int staged_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    Engine& e = engine();
    std::lock_guard<std::mutex> l(e.mu);
    HZB_TRY(e.init());
    const size_t n = (bytes + CHUNK - 1) / CHUNK;
    for (size_t i = 0; i <= n; ++i) {
        if (i < n) {   // DMA of chunk i (its slot was drained when chunk i-RING+... was copied out: host copies are in order)
            const size_t o = i * CHUNK, len = std::min(CHUNK, bytes - o);
            HZB_CUDA(cudaMemcpyAsync(e.pin[i % RING], (const char*)src_dev + o, len, cudaMemcpyDeviceToHost, st));
            HZB_CUDA(cudaEventRecord(e.ev[i % RING], st));
        }
        if (i >= 1) {  // meanwhile chunk i-1 goes from the pinned buffer to the caller's array
            const size_t o = (i - 1) * CHUNK, len = std::min(CHUNK, bytes - o);
            HZB_CUDA(cudaEventSynchronize(e.ev[(i - 1) % RING]));
            e.pcopy((char*)dst_host + o, e.pin[(i - 1) % RING], len);
        }
    }
    return 0;
}

int staged_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    Engine& e = engine();
    std::lock_guard<std::mutex> l(e.mu);
    HZB_TRY(e.init());
    const size_t n = (bytes + CHUNK - 1) / CHUNK;
    for (size_t i = 0; i < n; ++i) {
        const size_t o = i * CHUNK, len = std::min(CHUNK, bytes - o);
        if (i >= RING) HZB_CUDA(cudaEventSynchronize(e.ev[i % RING]));   // its previous DMA has left the buffer
        e.pcopy(e.pin[i % RING], (const char*)src_host + o, len);
        HZB_CUDA(cudaMemcpyAsync((char*)dst_dev + o, e.pin[i % RING], len, cudaMemcpyHostToDevice, st));
        HZB_CUDA(cudaEventRecord(e.ev[i % RING], st));
    }
    // the pinned ring is shared: the DMAs must have left it before the lock is released
    for (int k = 0; k < RING; ++k) if ((size_t)k < n) HZB_CUDA(cudaEventSynchronize(e.ev[k]));
    return 0;
}

// ---- pooled pinned host blocks for large OUTPUT arrays --------------------------------
// The wrappers allocate the arrays they return; backing the big ones (the 2 GB horizon
// array) with page-locked memory lets the finished row blocks leave by plain DMA while the
// kernel runs and lets sky_view_factor read them back by DMA.  Page-locking 2 GB costs a
// few 100 ms, so ONE freed block (of at most 4 GB) is kept and handed out again.
namespace {
struct PinPool {
    std::mutex mu;
    struct Blk { void* p; size_t cap; };
    std::vector<Blk> free_blocks, live;
};
PinPool& pin_pool() { static PinPool* p = new PinPool(); return *p; }
}  // namespace

void* host_block_alloc(size_t bytes) {
    if (bytes == 0) return nullptr;
    PinPool& P = pin_pool();
    std::lock_guard<std::mutex> l(P.mu);
    int best = -1;
    for (int i = 0; i < (int)P.free_blocks.size(); ++i)
        if (P.free_blocks[i].cap >= bytes && P.free_blocks[i].cap <= bytes + bytes / 4 &&
            (best < 0 || P.free_blocks[i].cap < P.free_blocks[best].cap)) best = i;
    PinPool::Blk b{nullptr, 0};
    if (best >= 0) { b = P.free_blocks[best]; P.free_blocks.erase(P.free_blocks.begin() + best); }
    else {
        if (cudaHostAlloc(&b.p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        b.cap = bytes;
    }
    P.live.push_back(b);
    return b.p;
}

void host_block_free(void* p) {
    if (!p) return;
    PinPool& P = pin_pool();
    std::lock_guard<std::mutex> l(P.mu);
    for (size_t i = 0; i < P.live.size(); ++i)
        if (P.live[i].p == p) {
            P.free_blocks.push_back(P.live[i]);
            P.live.erase(P.live.begin() + i);
            // bound the idle page-locked memory: one block of at most 4 GB stays (hzb_trim() releases it too)
            while (P.free_blocks.size() > 1 || (!P.free_blocks.empty() && P.free_blocks[0].cap > ((size_t)4 << 30))) {
                if (cudaFreeHost(P.free_blocks[0].p) != cudaSuccess) cudaGetLastError();
                P.free_blocks.erase(P.free_blocks.begin());
            }
            return;
        }
}

void host_block_trim() {
    PinPool& P = pin_pool();
    std::lock_guard<std::mutex> l(P.mu);
    for (PinPool::Blk& b : P.free_blocks) if (cudaFreeHost(b.p) != cudaSuccess) cudaGetLastError();
    P.free_blocks.clear();
}

bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace hzb
