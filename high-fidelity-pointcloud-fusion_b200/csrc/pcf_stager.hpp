// pcf_stager.hpp -- host staging pool of libpcfusion.so (plain C++17, no CUDA types on this side).
//
// Replaces the reference's first pipeline stage: the `addPoints()` thread that pops a PointCloud2 message from `clouds_`,
// decodes it and copies the points inside the camera-frame depth clip into `clouds_processed_` (node.cpp:218-263,
// decode node.cpp:182-216, clip node.cpp:248-255).  Here a pool of threads does that walk once per message, in
// parallel ACROSS frames: every point is read, points outside the clip are dropped (the float thresholds are the
// ones the kernel uses, exactly equivalent to the reference's double compares), and the survivors are packed, in
// point order, as 12-byte xyz into a pinned slot.  Slots are handed to the GPU strictly in submission order, so the
// arrival order every per-voxel buffer depends on (OG.hpp:211,230,239) is unchanged; the kernel still applies the
// FP64 transform, the box test and (again, harmlessly) the clip.  What crosses PCIe is the clipped cloud
// (C2: 5.5 instead of 16 bytes per input point), which is what bounds the end-to-end rate.
//
// Raw lanes.  Packing is bound by the host's memory bandwidth and cores, the unstaged upload by PCIe: while the packers
// are saturated the link is half idle.  A few extra "raw lane" threads take clouds from the same queue and upload them
// as they are (pinned float4 / packed-xyz arrays only; anything else they pack like everybody else).  A raw lane waits for
// its own copy to finish before it takes the next cloud, so it only ever uses link time the staged copies leave free, and
// the same in-order hand-over keeps the arrival order.  Results are identical either way: the kernel applies the clip.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

namespace pcf {

struct StageJob {
    const uint8_t* data = nullptr;   // first byte of the cloud (message payload); must stay valid until the job is staged
    uint32_t rows = 1, cols = 0;     // organized cloud: rows x cols points (flat cloud: 1 x n)
    uint64_t row_step = 0;           // bytes between rows
    uint32_t point_step = 16;        // bytes between points of a row
    uint32_t x_offset = 0;           // byte offset of the float32 x field; y and z follow it
    double pose[16];
    uint32_t frame_idx = 0;
};

// Order-preserving clip-and-pack of one cloud (pcf_pack.cpp: AVX-512 / AVX2 / scalar, selected at run time).  `out` must
// hold 3 * (rows * cols) + 16 floats.  Returns the number of packed points, padded with NaN points (which fail the kernel's
// clip) to a multiple of 4 so that every 256-point chunk of the packed cloud is a 16-byte multiple for the bulk-copy kernel.
uint32_t clip_pack(const StageJob& j, float clip_lo, float clip_hi, float* out);
int clip_pack_isa();     // 0 scalar, 1 AVX2, 2 AVX-512
uint32_t clip_pack_with(int isa, const StageJob& j, float clip_lo, float clip_hi, float* out);   // tests: force an implementation

class Stager {
   public:
    struct Hooks {
        std::function<void*(size_t)> alloc_pinned;
        std::function<void(void*)> free_pinned;
        std::function<void(int)> thread_init;                       // once per worker thread (cudaSetDevice)
        std::function<void(int)> slot_wait;                         // block until the last upload out of slot s has completed
        // hand a staged cloud to the GPU (H2D copy + integration launch); called in submission order, one at a time
        std::function<int(int slot, const float* xyz, uint32_t n_staged, uint32_t n_offered, const double* pose, uint32_t frame_idx)> push;
        // raw lanes: can this cloud be uploaded as it is (pinned, plain [n, stride] array)?  upload it from the caller's
        // memory (in submission order, like push); block until lane `lane`'s last upload has left the caller's buffer
        std::function<bool(const StageJob&)> raw_ok;
        std::function<int(int lane, const StageJob&)> push_raw;
        std::function<void(int lane)> raw_wait;
    };

    Stager(int threads, int raw_lanes, float clip_lo, float clip_hi, Hooks hooks)
        : clip_lo_(clip_lo), clip_hi_(clip_hi), hooks_(std::move(hooks)) {
        n_threads_ = threads < 1 ? 1 : threads;
        n_raw_ = raw_lanes < 0 ? 0 : raw_lanes;
        slots_.resize((size_t)(n_threads_ + n_raw_) * 2);
        max_queue_ = (size_t)n_threads_ * 4;
        for (int t = 0; t < n_threads_; t++) pool_.emplace_back(&Stager::worker, this, -1);
        for (int t = 0; t < n_raw_; t++) pool_.emplace_back(&Stager::worker, this, t);
    }
    ~Stager() {
        drain();
        {
            std::lock_guard<std::mutex> lk(m_);
            quit_ = true;
        }
        cv_work_.notify_all();
        cv_raw_.notify_all();
        cv_order_.notify_all();
        for (std::thread& t : pool_) t.join();
        for (Slot& s : slots_) if (s.p) hooks_.free_pinned(s.p);
    }
    int n_slots() const { return (int)slots_.size(); }
    int n_threads() const { return n_threads_; }
    int n_raw_lanes() const { return n_raw_; }
    uint64_t raw_pushed() {
        std::lock_guard<std::mutex> lk(m_);
        return raw_pushed_;
    }

    void submit(const StageJob& j) {                                 // node.cpp:345-347 (clouds_.push_back)
        std::unique_lock<std::mutex> lk(m_);
        cv_space_.wait(lk, [&] { return queue_.size() < max_queue_; });
        queue_.push_back(j);
        const bool backlog = n_raw_ > 0 && queue_.size() >= (size_t)n_threads_;
        lk.unlock();
        cv_work_.notify_one();
        if (backlog) cv_raw_.notify_one();       // clouds are piling up behind the packers: a raw lane may help
    }
    // node.cpp:356: drop what no worker has taken yet.  Returns the number of dropped clouds.
    size_t drop_queued() {
        std::lock_guard<std::mutex> lk(m_);
        size_t n = queue_.size();
        queue_.clear();
        cv_space_.notify_all();
        cv_order_.notify_all();
        return n;
    }
    // wait until every submitted cloud has been handed to the GPU; returns the first push error (0 = none) and clears it
    int drain() {
        std::unique_lock<std::mutex> lk(m_);
        cv_order_.wait(lk, [&] { return queue_.empty() && next_push_ == next_seq_; });
        int rc = err_;
        err_ = 0;
        return rc;
    }
    // number of clouds handed to the GPU so far (their source buffers are no longer read)
    uint64_t staged() {
        std::lock_guard<std::mutex> lk(m_);
        return next_push_;
    }
    // block until `n` clouds have been handed over -- or nothing is pending any more (the n-th may have been dropped)
    void wait_staged(uint64_t n) {
        std::unique_lock<std::mutex> lk(m_);
        cv_order_.wait(lk, [&] { return next_push_ >= n || (queue_.empty() && next_push_ == next_seq_); });
    }

   private:
    struct Slot { float* p = nullptr; size_t cap = 0; };

    // raw_lane < 0: packer.  raw_lane >= 0: raw lane (uploads eligible clouds unstaged, packs the others)
    void worker(int raw_lane) {
        if (hooks_.thread_init) hooks_.thread_init(0);
        for (;;) {
            StageJob j;
            uint64_t seq;
            if (raw_lane >= 0) hooks_.raw_wait(raw_lane);            // the link is free of this lane's previous cloud
            {
                std::unique_lock<std::mutex> lk(m_);
                // a raw lane only helps out when clouds are piling up behind the packers
                if (raw_lane < 0) cv_work_.wait(lk, [&] { return quit_ || !queue_.empty(); });
                else cv_raw_.wait(lk, [&] { return quit_ || queue_.size() >= (size_t)n_threads_; });
                if (quit_ && queue_.empty()) return;
                if (queue_.empty()) continue;
                j = queue_.front();
                queue_.pop_front();
                seq = next_seq_++;                                   // taken in FIFO order under the lock: seq order == submission order
                cv_space_.notify_one();
                if (raw_lane >= 0 && hooks_.raw_ok(j)) {
                    cv_order_.wait(lk, [&] { return next_push_ == seq; });
                    int rc = hooks_.push_raw(raw_lane, j);
                    if (rc < 0 && err_ == 0) err_ = rc;
                    next_push_++;
                    raw_pushed_++;
                    lk.unlock();
                    cv_order_.notify_all();
                    continue;
                }
                // the slot was last used by cloud seq - n_slots: that one must have been handed over before it is refilled
                const uint64_t ns = slots_.size();
                cv_order_.wait(lk, [&] { return seq < ns || next_push_ > seq - ns; });
            }
            const int s = (int)(seq % slots_.size());
            hooks_.slot_wait(s);
            Slot& sl = slots_[s];
            const size_t need = 3 * (size_t)j.rows * j.cols + 16;
            if (need > sl.cap) {
                if (sl.p) hooks_.free_pinned(sl.p);
                sl.cap = need + need / 8;
                sl.p = static_cast<float*>(hooks_.alloc_pinned(sl.cap * sizeof(float)));
            }
            uint32_t n = 0;
            if (sl.p) n = clip_pack(j, clip_lo_, clip_hi_, sl.p);
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_order_.wait(lk, [&] { return next_push_ == seq; });
                int rc = sl.p ? hooks_.push(s, sl.p, n, j.rows * j.cols, j.pose, j.frame_idx) : -2;
                if (rc < 0 && err_ == 0) err_ = rc;
                next_push_++;
            }
            cv_order_.notify_all();
        }
    }

    float clip_lo_, clip_hi_;
    Hooks hooks_;
    int n_threads_ = 1, n_raw_ = 0;
    uint64_t raw_pushed_ = 0;
    std::vector<Slot> slots_;
    std::vector<std::thread> pool_;
    std::deque<StageJob> queue_;
    size_t max_queue_ = 8;
    std::mutex m_;
    std::condition_variable cv_work_, cv_raw_, cv_space_, cv_order_;
    uint64_t next_seq_ = 0, next_push_ = 0;
    int err_ = 0;
    bool quit_ = false;
};

}  // namespace pcf
