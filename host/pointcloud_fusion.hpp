// host/pointcloud_fusion.hpp -- offline replay counterpart of the reference's `class PointcloudFusion`
// (pointcloud_fusion/pointcloud_fusion/src/pointcloud_fusion_and_filter.cpp:99-440).
//
// Same service semantics, ROS removed:
//   onReceivedPointCloud()  node.cpp:327-349  enqueue a (cloud, pose) pair while `start_` is set, drop it otherwise
//   start() / stop()        node.cpp:361-375  flip the gate; frames already queued keep integrating
//   reset()                 node.cpp:351-359  drop the not-yet-integrated input, keep the grid
//   getFusedCloud()         node.cpp:377-440  drain, write <dir>/test_cloud.pcd + <dir>/meta.csv, clear the grid
// What the three worker threads + two mutex-protected deques of the reference (node.cpp:130-143,218-325) did on the
// CPU is now: one pool of PINNED staging slots (filled by the producer), one worker thread that hands every queued
// slot to pcf_push_frame (async H2D + the integration kernel on the context's CUDA streams), and an explicit
// update schedule instead of the 5 s cleanGrid timer (D4): `update_every` frames, plus once at process().
#pragma once
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../include/pcfusion.h"

namespace pcfusion {

class PointcloudFusion {
   public:
    struct Params {
        double box[6] = {-0.8, 1.8, -1.5, 1.5, 0.0, 1.0};   // launch:8 (xmin xmax ymin ymax zmin zmax)
        float res = 0.005f;                                 // node.cpp:91
        double clip_zmin = 0.28, clip_zmax = 0.6;           // node.cpp:92-93
        std::string directory_name = ".";                   // launch:6
        int device = 0;
        int update_every = 0;        // run updateThicknessVectors after every k integrated frames (0: only at process)
        uint32_t staging_slots = 8;  // pinned slots = frames that can be in flight (the reference queues up to 100 messages, node.cpp:152)
        uint64_t log_capacity_hint = 0;
    };
    struct Counters {
        uint64_t received = 0, dropped = 0, integrated = 0, discarded_by_reset = 0, updates = 0;
    };

    explicit PointcloudFusion(const Params& p);
    ~PointcloudFusion();
    bool ok() const { return ctx_ != nullptr; }
    const std::string& last_error() const { return err_; }

    // Zero-copy producer interface: get a pinned slot able to hold `floats` floats (blocks while all slots are in
    // flight), fill it, submit it.  submit() returns false when the frame was dropped because fusion is stopped.
    float* acquire(size_t floats);
    bool submit(float* slot, uint32_t n_points, uint32_t stride_floats, const double pose[16]);
    // node.cpp:327-349: copying variant for callers that own their cloud memory
    bool onReceivedPointCloud(const float* xyz, uint32_t n_points, uint32_t stride_floats, const double pose[16]);

    bool reset();          // node.cpp:351-359
    bool start();          // node.cpp:361-367
    bool stop();           // node.cpp:369-375
    bool getFusedCloud();  // node.cpp:377-440 ("process"); D8: returns the success flag the reference forgets to set
    void drain();          // wait until every queued frame has been handed to the GPU and integrated

    Counters counters();
    pcf_ctx* handle() const { return ctx_; }
    float last_process_ms() const { return process_ms_; }

   private:
    struct Slot { float* p = nullptr; size_t cap = 0; uint64_t ticket = 0; bool busy = false; };
    struct Item { int slot; uint32_t n, stride; double pose[16]; };
    void worker();

    Params prm_;
    pcf_ctx* ctx_ = nullptr;
    std::string err_;
    std::vector<Slot> slots_;
    std::deque<Item> clouds_;             // node.cpp:137
    std::mutex mtx_;                      // node.cpp:140
    std::condition_variable cv_work_, cv_free_, cv_idle_;
    std::thread thread_;
    bool start_ = false, quit_ = false, busy_ = false;
    uint32_t next_frame_ = 0;
    Counters cnt_;
    float process_ms_ = 0.f;
};

}  // namespace pcfusion
