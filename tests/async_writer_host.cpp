// Host check of integration/b200_host.h: AsyncWriter (the bounded encode queue behind the drop-in mains).
//   g++ -std=c++14 -O1 -pthread -I include -I integration tests/async_writer_host.cpp -o <exe>
// Prints one line per check; exit code 0 = all held.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "b200_host.h"

using clk = std::chrono::steady_clock;
static double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

int main() {
  int bad = 0;
  {  // (1) jobs run in submission order, on another thread, and the producer is not blocked while the queue has room
    AsyncWriter w(2);
    std::vector<int> order;
    std::atomic<int> running{0};
    const std::thread::id me = std::this_thread::get_id();
    bool other_thread = true;
    const auto t0 = clk::now();
    for (int i = 0; i < 2; i++)
      w.submit([&, i]() -> int {
        running++;
        if (std::this_thread::get_id() == me) other_thread = false;
        std::this_thread::sleep_for(std::chrono::milliseconds(60));
        order.push_back(i);
        return 0;
      });
    const double t_submit = ms_since(t0);   // two 60 ms jobs accepted without waiting for either
    const int rc = w.finish();
    const double t_all = ms_since(t0);
    const bool ok = rc == 0 && order.size() == 2 && order[0] == 0 && order[1] == 1 && other_thread && t_submit < 30.0 && t_all >= 110.0;
    std::printf("overlap: submit %.1f ms, all %.1f ms, order %zu -> %s\n", t_submit, t_all, order.size(), ok ? "ok" : "FAILED");
    bad += !ok;
  }
  {  // (2) bounded: with depth 1 the third submit waits for the first job to finish (at most one frame waits in memory)
    AsyncWriter w(1);
    const auto t0 = clk::now();
    for (int i = 0; i < 3; i++) w.submit([]() -> int { std::this_thread::sleep_for(std::chrono::milliseconds(50)); return 0; });
    const double t_submit = ms_since(t0);
    w.finish();
    const bool ok = t_submit >= 45.0 && t_submit < 140.0;
    std::printf("bounded: three submits took %.1f ms -> %s\n", t_submit, ok ? "ok" : "FAILED");
    bad += !ok;
  }
  {  // (3) the first error code survives later successes; finish() twice is harmless
    AsyncWriter w(2);
    w.submit([]() -> int { return 0; });
    w.submit([]() -> int { return 4; });
    w.submit([]() -> int { return 7; });
    w.submit([]() -> int { return 0; });
    const int rc = w.finish(), again = w.finish();
    const bool ok = rc == 4 && again == 4;
    std::printf("errors: finish() = %d, %d -> %s\n", rc, again, ok ? "ok" : "FAILED");
    bad += !ok;
  }
  return bad;
}
