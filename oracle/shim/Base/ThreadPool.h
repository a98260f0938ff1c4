// ORACLE (test infrastructure, not product code).
// Include-path shim for the reference's "Base/ThreadPool.h" (the reference header is
// src/Base/ThreadPool.h:19-121; RendererSoft.h:11 includes it by that relative name, so putting
// this directory first on the include path swaps it without touching any reference file).
//
// Why: the reference rasteriser pushes one task per (triangle x 32x32 block) and lets
// hardware_concurrency() workers race on depth/colour read-modify-writes of the same pixels
// (RendererSoft.cpp:731-768).  For an order-deterministic parity image every
// task must run in submission (= primitive) order.  This pool therefore executes each task inline
// on the submitting thread; worker id is always 0 and getThreadCnt() is 1.
#pragma once

#include <atomic>
#include <cstddef>
#include <functional>

namespace SoftGL {

class ThreadPool {
 public:
  explicit ThreadPool(const size_t threadCnt = 1) { (void) threadCnt; }

  inline size_t getThreadCnt() const { return 1; }

  template<typename F>
  void pushTask(const F &task) {
    task(0);
  }

  void waitTasksFinish() const {}

  std::atomic<bool> paused{false};
};

}  // namespace SoftGL
