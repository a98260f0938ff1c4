// CLI of the headless offscreen harness.
//   player <trace.sglt> [--out file] [--data-dir dir] [--frames N] [--warmup W]
// Replays the trace once; with --frames N the FRAME section is re-executed N more times and each
// repetition is wall-clock timed (steady_clock around submit + waitIdle, SURVEY.md section 8d
// "CPU baseline timing").  Prints one JSON line with per-frame milliseconds.
#include "trace_player.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s trace.sglt [--out f] [--data-dir d] [--frames N] [--warmup W]\n", argv[0]);
    return 2;
  }
  std::string trace = argv[1], out, dataDir = ".";
  int frames = 0, warmup = 0;
  for (int i = 2; i < argc; i++) {
    if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
    else if (!strcmp(argv[i], "--data-dir") && i + 1 < argc) dataDir = argv[++i];
    else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
  }
  TracePlayer player;
  if (!player.load(trace)) return 1;
  player.setDataDir(dataDir);
  player.setOutput(out);
  if (!player.runSetup()) return 1;
  if (!player.runFrame(true)) return 1;
  for (int i = 0; i < warmup; i++) player.runFrame(true);
  std::vector<double> ms;
  for (int i = 0; i < frames; i++) {
    auto t0 = std::chrono::steady_clock::now();
    player.runFrame(true);
    auto t1 = std::chrono::steady_clock::now();
    ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
  }
  if (!player.runTail()) return 1;
  if (!ms.empty()) {
    std::vector<double> s = ms;
    std::sort(s.begin(), s.end());
    double sum = 0;
    for (double v : ms) sum += v;
    printf("{\"backend\": \"%s\", \"frames\": %d, \"ms_median\": %.4f, \"ms_mean\": %.4f, \"ms_min\": %.4f, "
           "\"threads\": %u}\n",
           PlayerBackend::name(), frames, s[s.size() / 2], sum / ms.size(), s[0],
           std::thread::hardware_concurrency());
  }
  return 0;
}
