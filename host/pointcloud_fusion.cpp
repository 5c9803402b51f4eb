// host/pointcloud_fusion.cpp -- see pointcloud_fusion.hpp.  Plain C++17 + the C ABI; no CUDA headers on this side.
#include "pointcloud_fusion.hpp"

#include <chrono>
#include <cstring>
#include <iostream>

namespace pcfusion {

PointcloudFusion::PointcloudFusion(const Params& p) : prm_(p) {
    pcf_config cfg;
    pcf_default_config(&cfg);
    std::memcpy(cfg.box, p.box, sizeof cfg.box);
    cfg.res[0] = cfg.res[1] = cfg.res[2] = p.res;                 // node.cpp:161
    cfg.clip_zmin = p.clip_zmin;
    cfg.clip_zmax = p.clip_zmax;
    cfg.device = p.device;
    cfg.log_capacity_hint = p.log_capacity_hint;
    if (pcf_create(&cfg, &ctx_) != PCF_OK) {                      // node.cpp:161-164
        err_ = pcf_last_error(nullptr);
        ctx_ = nullptr;
        return;
    }
    pcf_start(ctx_);       // the library-side gate stays open; the service gate is start_ below, like node.cpp:329
    slots_.resize(p.staging_slots ? p.staging_slots : 1);
    thread_ = std::thread(&PointcloudFusion::worker, this);      // node.cpp:166-168
}

PointcloudFusion::~PointcloudFusion() {
    {
        std::lock_guard<std::mutex> lk(mtx_);
        quit_ = true;
    }
    cv_work_.notify_all();
    if (thread_.joinable()) thread_.join();
    if (ctx_) pcf_sync(ctx_);
    for (Slot& s : slots_) pcf_host_free(s.p);
    pcf_destroy(ctx_);
}

float* PointcloudFusion::acquire(size_t floats) {
    if (!ctx_) return nullptr;
    std::unique_lock<std::mutex> lk(mtx_);
    int idx = -1;
    cv_free_.wait(lk, [&] {
        for (size_t i = 0; i < slots_.size(); i++)
            if (!slots_[i].busy) { idx = (int)i; return true; }
        return false;
    });
    Slot& s = slots_[idx];
    s.busy = true;
    uint64_t ticket = s.ticket;
    lk.unlock();
    pcf_wait_upload(ctx_, ticket);        // the H2D copy that last read this slot must have finished
    if (floats > s.cap) {
        pcf_host_free(s.p);
        s.cap = floats + floats / 8;
        s.p = static_cast<float*>(pcf_host_alloc(s.cap * sizeof(float)));
        if (!s.p) { s.cap = 0; err_ = "pinned allocation failed"; std::lock_guard<std::mutex> g(mtx_); s.busy = false; return nullptr; }
    }
    return s.p;
}

bool PointcloudFusion::submit(float* slot, uint32_t n, uint32_t stride, const double pose[16]) {
    std::unique_lock<std::mutex> lk(mtx_);
    int idx = -1;
    for (size_t i = 0; i < slots_.size(); i++) if (slots_[i].p == slot && slots_[i].busy) idx = (int)i;
    if (idx < 0) return false;
    cnt_.received++;
    if (!start_) {                       // node.cpp:329-331: messages are ignored unless started
        cnt_.dropped++;
        slots_[idx].busy = false;
        lk.unlock();
        cv_free_.notify_one();
        return false;
    }
    Item it;
    it.slot = idx; it.n = n; it.stride = stride;
    std::memcpy(it.pose, pose, sizeof it.pose);
    clouds_.push_back(it);               // node.cpp:345-347
    lk.unlock();
    cv_work_.notify_one();
    return true;
}

bool PointcloudFusion::onReceivedPointCloud(const float* xyz, uint32_t n, uint32_t stride, const double pose[16]) {
    float* s = acquire((size_t)n * stride);
    if (!s) return false;
    std::memcpy(s, xyz, (size_t)n * stride * sizeof(float));
    return submit(s, n, stride, pose);
}

void PointcloudFusion::worker() {        // node.cpp:218-299 (addPoints + updateStates threads) and 301-325 (cleanGrid)
    for (;;) {
        Item it;
        {
            std::unique_lock<std::mutex> lk(mtx_);
            busy_ = false;
            cv_idle_.notify_all();
            cv_work_.wait(lk, [&] { return quit_ || !clouds_.empty(); });
            if (quit_ && clouds_.empty()) return;
            it = clouds_.front();
            clouds_.pop_front();
            busy_ = true;
        }
        int rc = pcf_push_frame(ctx_, slots_[it.slot].p, it.n, it.stride, it.pose, next_frame_++);
        uint64_t ticket = 0;
        pcf_upload_ticket(ctx_, &ticket);
        bool do_update = false;
        {
            std::lock_guard<std::mutex> lk(mtx_);
            slots_[it.slot].ticket = ticket;
            slots_[it.slot].busy = false;
            if (rc == PCF_OK) {
                cnt_.integrated++;
                do_update = prm_.update_every > 0 && cnt_.integrated % (uint64_t)prm_.update_every == 0;
            } else if (rc < 0) {
                err_ = pcf_last_error(ctx_);
            }
        }
        cv_free_.notify_one();
        if (do_update && pcf_update(ctx_) == PCF_OK) {           // node.cpp:305-321
            std::lock_guard<std::mutex> lk(mtx_);
            cnt_.updates++;
        }
    }
}

void PointcloudFusion::drain() {
    std::unique_lock<std::mutex> lk(mtx_);
    cv_idle_.wait(lk, [&] { return clouds_.empty() && !busy_; });
    lk.unlock();
    if (ctx_) pcf_sync(ctx_);
}

bool PointcloudFusion::reset() {         // node.cpp:351-359
    std::cout << "RESET" << std::endl;
    std::lock_guard<std::mutex> lk(mtx_);
    for (const Item& it : clouds_) { slots_[it.slot].busy = false; cnt_.discarded_by_reset++; }
    clouds_.clear();
    cv_free_.notify_all();
    return true;
}
bool PointcloudFusion::start() {         // node.cpp:361-367
    std::cout << "START" << std::endl;
    std::lock_guard<std::mutex> lk(mtx_);
    start_ = true;
    return true;
}
bool PointcloudFusion::stop() {          // node.cpp:369-375
    std::cout << "STOP" << std::endl;
    std::lock_guard<std::mutex> lk(mtx_);
    start_ = false;
    return true;
}

bool PointcloudFusion::getFusedCloud() { // node.cpp:377-440
    if (!ctx_) return false;
    std::cout << "Downloading cloud." << std::endl;
    drain();                             // node.cpp:380-394: wait until both queues are empty
    auto t0 = std::chrono::steady_clock::now();
    if (pcf_update(ctx_) != PCF_OK) { err_ = pcf_last_error(ctx_); return false; }       // D4: final updateThicknessVectors
    {
        std::lock_guard<std::mutex> lk(mtx_);
        cnt_.updates++;
    }
    const std::string cloud = prm_.directory_name + "/test_cloud.pcd", meta = prm_.directory_name + "/meta.csv";
    int rc = pcf_process(ctx_, cloud.c_str(), meta.c_str());     // downloadData + clearVoxels, node.cpp:395-398,438
    process_ms_ = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc != PCF_OK) { err_ = pcf_last_error(ctx_); return false; }
    next_frame_ = 0;
    std::cout << "Fused cloud saved to " << cloud << std::endl;
    return true;                         // D8: the reference's handler never sets res.success
}

PointcloudFusion::Counters PointcloudFusion::counters() {
    std::lock_guard<std::mutex> lk(mtx_);
    return cnt_;
}

}  // namespace pcfusion
