// b200_host.h — what every drop-in main (main_hw*_b200.cpp, main_ppm_b200.cpp) shares on the host side.
//
// 1. B200Devices: one b200rt_scene per GPU of the box (B200RT_DEVICES=n caps it), the replacement for the reference's
//    `thread_count = std::thread::hardware_concurrency()` (HW7/src/main.cpp:31-34): the frame is fanned out over GPUs
//    by b200rt_render_multi the way main fans it out over threads.
// 2. AsyncWriter: the reference encodes camera i's PNG / EXR on the main thread before it starts camera i+1
//    (HW7/src/main.cpp:105-198, HW3/src/main.cpp:44-62).  With the render down to milliseconds the encode IS the frame
//    time, so it moves to a worker thread behind a bounded queue: camera i's encode overlaps camera i+1's render, at
//    most `depth` finished frames wait in memory, and an encoder error still ends the program with a non-zero code.
//    The bytes written are the same: the job is the reference's own lodepng / tinyexr call on the same buffer.
#pragma once
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <iostream>
#include <mutex>
#include <thread>
#include <vector>

#include "b200rt.h"

struct B200Devices {
  std::vector<b200rt_scene*> scenes;
  // Returns 0, or 2 after printing why (no device, out of memory ...): there is no CPU fallback.
  int create(const B200SceneDesc* desc) {
    int n = b200rt_device_count();
    if (const char* e = std::getenv("B200RT_DEVICES")) {
      const int cap = std::atoi(e);
      if (cap >= 1 && cap < n) n = cap;
    }
    if (n < 1) n = 1;   // b200rt_scene_create then reports "no CUDA device"
    for (int dev = 0; dev < n; dev++) {
      b200rt_scene* s = nullptr;
      if (b200rt_scene_create(desc, dev, &s) != B200RT_OK) {
        std::cerr << "b200rt: " << b200rt_last_error() << std::endl;
        destroy();
        return 2;
      }
      scenes.push_back(s);
    }
    return 0;
  }
  void destroy() {
    for (b200rt_scene* s : scenes) b200rt_scene_destroy(s);
    scenes.clear();
  }
  int size() const { return (int)scenes.size(); }
  // The reference's fan-out over Scene::render_image, over GPUs.
  int render(int camera_index, const B200RenderOpts* opts, float* rgbw, B200RenderStats* stats) {
    if (scenes.size() == 1) return b200rt_render(scenes[0], camera_index, opts, rgbw, stats);
    return b200rt_render_multi(scenes.data(), (int)scenes.size(), camera_index, opts, rgbw, stats);
  }
};

class AsyncWriter {
 public:
  explicit AsyncWriter(size_t depth = 2) : depth_(depth), worker_([this] { run(); }) {}
  ~AsyncWriter() { finish(); }
  // Blocks while `depth` jobs are waiting.  A job returns 0 or an error code.
  void submit(std::function<int()> job) {
    std::unique_lock<std::mutex> lk(m_);
    not_full_.wait(lk, [this] { return q_.size() < depth_; });
    q_.push_back(std::move(job));
    not_empty_.notify_one();
  }
  // Waits for every submitted job; returns the first non-zero code any of them returned.
  int finish() {
    {
      std::unique_lock<std::mutex> lk(m_);
      if (done_) return rc_;
      done_ = true;
      not_empty_.notify_one();
    }
    worker_.join();
    return rc_;
  }

 private:
  void run() {
    for (;;) {
      std::function<int()> job;
      {
        std::unique_lock<std::mutex> lk(m_);
        not_empty_.wait(lk, [this] { return !q_.empty() || done_; });
        if (q_.empty()) return;
        job = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
      }
      const int rc = job();
      if (rc && !rc_) rc_ = rc;
    }
  }
  size_t depth_;
  std::mutex m_;
  std::condition_variable not_full_, not_empty_;
  std::deque<std::function<int()>> q_;
  bool done_ = false;
  int rc_ = 0;
  std::thread worker_;
};
