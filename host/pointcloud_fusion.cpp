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
    cfg.stage_threads = p.stage_threads;
    if (pcf_create(&cfg, &ctx_) != PCF_OK) {                      // node.cpp:161-164
        err_ = pcf_last_error(nullptr);
        ctx_ = nullptr;
        return;
    }
    // a scan of known size: pre-size the process() scratch so that the first getFusedCloud() does not pay for device allocations
    if (p.log_capacity_hint) pcf_reserve_process(ctx_, p.log_capacity_hint, p.log_capacity_hint / 8 + 1024);
    slots_.resize(p.staging_slots ? p.staging_slots : 1);        // start_ is false until start(), like node.cpp:135
}

PointcloudFusion::~PointcloudFusion() {
    if (ctx_) pcf_sync(ctx_);
    for (Slot& s : slots_) pcf_host_free(s.p);
    pcf_destroy(ctx_);
}

float* PointcloudFusion::acquire(size_t floats) {
    if (!ctx_) return nullptr;
    Slot& s = slots_[next_slot_];                                 // round robin: the oldest buffer is the first to come free
    if (s.busy) {
        pcf_wait_staged(ctx_, s.staged_at);                       // its cloud has been packed into a pinned slot (or dropped)
        s.busy = false;
    }
    if (floats > s.cap) {
        pcf_host_free(s.p);
        s.cap = floats + floats / 8;
        s.p = static_cast<float*>(pcf_host_alloc(s.cap * sizeof(float)));
        if (!s.p) { s.cap = 0; err_ = "pinned allocation failed"; return nullptr; }
    }
    return s.p;
}

bool PointcloudFusion::push(const float* xyz, uint32_t n, uint32_t stride, const double pose[16]) {
    cnt_.received++;
    int rc = pcf_submit_frame(ctx_, xyz, n, stride, pose, next_frame_);     // node.cpp:329-347 (gate + clouds_.push_back)
    if (rc == PCF_DROPPED) { cnt_.dropped++; return false; }               // node.cpp:329-331: ignored unless started
    if (rc < 0) { err_ = pcf_last_error(ctx_); return false; }
    next_frame_++;
    submitted_++;
    if (prm_.update_every > 0 && ++since_update_ == (uint64_t)prm_.update_every) {      // node.cpp:305-321
        since_update_ = 0;
        if (pcf_update(ctx_) == PCF_OK) cnt_.updates++;                     // drains the staging pool first: frame order is kept
        else err_ = pcf_last_error(ctx_);
    }
    return true;
}

bool PointcloudFusion::submit(float* slot, uint32_t n, uint32_t stride, const double pose[16]) {
    Slot& s = slots_[next_slot_];
    if (!ctx_ || slot != s.p) return false;
    if (!push(slot, n, stride, pose)) return false;              // dropped: the buffer is free again at once
    s.busy = true;
    s.staged_at = base_ + submitted_;
    next_slot_ = (next_slot_ + 1) % slots_.size();
    return true;
}

bool PointcloudFusion::onReceivedPointCloud(const float* xyz, uint32_t n, uint32_t stride, const double pose[16]) {
    return ctx_ && push(xyz, n, stride, pose);
}

void PointcloudFusion::drain() {
    if (ctx_ && pcf_sync(ctx_) < 0) err_ = pcf_last_error(ctx_);
}

bool PointcloudFusion::reset() {         // node.cpp:351-359
    std::cout << "RESET" << std::endl;
    if (!ctx_) return false;
    pcf_reset(ctx_);                     // start_ = false; clouds_.clear()
    pcf_drain(ctx_);                     // what the staging threads had already taken still integrates (clouds_processed_)
    for (Slot& s : slots_) s.busy = false;
    pcf_staged_count(ctx_, &base_);
    submitted_ = 0;
    return true;
}
bool PointcloudFusion::start() {         // node.cpp:361-367
    std::cout << "START" << std::endl;
    return ctx_ && pcf_start(ctx_) == PCF_OK;
}
bool PointcloudFusion::stop() {          // node.cpp:369-375
    std::cout << "STOP" << std::endl;
    return ctx_ && pcf_stop(ctx_) == PCF_OK;
}

bool PointcloudFusion::getFusedCloud() { // node.cpp:377-440
    if (!ctx_) return false;
    std::cout << "Downloading cloud." << std::endl;
    drain();                             // node.cpp:380-394: wait until both queues are empty
    auto t0 = std::chrono::steady_clock::now();
    if (pcf_update(ctx_) != PCF_OK) { err_ = pcf_last_error(ctx_); return false; }       // D4: final updateThicknessVectors
    cnt_.updates++;
    const std::string cloud = prm_.directory_name + "/test_cloud.pcd", meta = prm_.directory_name + "/meta.csv";
    int rc = pcf_process(ctx_, cloud.c_str(), meta.c_str());     // downloadData + clearVoxels, node.cpp:395-398,438
    process_ms_ = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc != PCF_OK) { err_ = pcf_last_error(ctx_); return false; }
    next_frame_ = 0;
    since_update_ = 0;
    std::cout << "Fused cloud saved to " << cloud << std::endl;
    return true;                         // D8: the reference's handler never sets res.success
}

PointcloudFusion::Counters PointcloudFusion::counters() {
    if (ctx_) {
        pcf_stats st;
        pcf_get_stats(ctx_, &st);
        cnt_.integrated = st.frames_pushed;
        cnt_.discarded_by_reset = st.staged_dropped;
    }
    return cnt_;
}

}  // namespace pcfusion
