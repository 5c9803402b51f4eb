// host/pointcloud_fusion.hpp -- offline replay counterpart of the reference's `class PointcloudFusion`
// (pointcloud_fusion/pointcloud_fusion/src/pointcloud_fusion_and_filter.cpp:99-440).
//
// Same service semantics, ROS removed:
//   onReceivedPointCloud()  node.cpp:327-349  enqueue a (cloud, pose) pair while `start_` is set, drop it otherwise
//   start() / stop()        node.cpp:361-375  flip the gate; frames already queued keep integrating
//   reset()                 node.cpp:351-359  start_ = false, drop the not-yet-staged input, keep the grid
//   getFusedCloud()         node.cpp:377-440  drain, write <dir>/test_cloud.pcd + <dir>/meta.csv, clear the grid
// The reference's pipeline -- clouds_ deque -> addPoints thread (decode + depth clip) -> clouds_processed_ deque ->
// updateStates thread (transform + grid insert), node.cpp:130-143,218-299 -- maps onto the library's staging pool
// (pcf_submit_frame: queue -> clip-and-pack threads -> pinned slots -> in-order H2D + integration kernel), so this
// class owns no thread of its own: every call comes from the caller's thread and one context is never driven from two
// threads at once.  The 5 s cleanGrid timer (node.cpp:301-325) becomes an explicit schedule (D4): `update_every`
// frames, plus once at process().
#pragma once
#include <cstdint>
#include <string>
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
        uint32_t staging_slots = 8;  // caller-side cloud buffers that can be in flight (the reference queues up to 100 messages, node.cpp:152)
        uint64_t log_capacity_hint = 0;
        int stage_threads = 0;       // host staging threads inside the library (0: auto)
    };
    struct Counters {
        uint64_t received = 0, dropped = 0, integrated = 0, discarded_by_reset = 0, updates = 0;
    };

    explicit PointcloudFusion(const Params& p);
    ~PointcloudFusion();
    bool ok() const { return ctx_ != nullptr; }
    const std::string& last_error() const { return err_; }

    // Producer interface without an extra copy: get a cloud buffer able to hold `floats` floats (blocks while all of
    // them are still being staged), fill it, submit it.  submit() returns false when the frame was dropped because
    // fusion is stopped.
    float* acquire(size_t floats);
    bool submit(float* slot, uint32_t n_points, uint32_t stride_floats, const double pose[16]);
    // node.cpp:327-349: variant for callers that own their cloud memory.  The cloud is read by the staging threads
    // after this call returns: it must stay valid until drain() (a ROS bridge holds the message pointer that long).
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
    struct Slot { float* p = nullptr; size_t cap = 0; uint64_t staged_at = 0; bool busy = false; };
    bool push(const float* xyz, uint32_t n, uint32_t stride, const double pose[16]);

    Params prm_;
    pcf_ctx* ctx_ = nullptr;
    std::string err_;
    std::vector<Slot> slots_;
    size_t next_slot_ = 0;
    uint64_t submitted_ = 0;      // clouds accepted by pcf_submit_frame since the last reset (the pool hands them over in this order)
    uint64_t base_ = 0;           // pcf_staged_count at the last reset
    uint32_t next_frame_ = 0;
    uint64_t since_update_ = 0;
    Counters cnt_;
    float process_ms_ = 0.f;
};

}  // namespace pcfusion
