// C entry points over TracePlayer so that Python (bench.py, tests) can drive the RendererCUDA harness in-process
// and time it with the library's CUDA events.  Built into lib/libsglhost.so (links libsglcuda.so).
#include "trace_player.h"

#include <cstring>

extern "C" {

void *sglp_create(void) { return new TracePlayer(); }
void sglp_destroy(void *p) { delete static_cast<TracePlayer *>(p); }
int sglp_load(void *p, const char *trace_path, const char *data_dir) {
  auto *tp = static_cast<TracePlayer *>(p);
  if (!tp->load(trace_path)) return -1;
  tp->setDataDir(data_dir ? data_dir : ".");
  return 0;
}
int sglp_setup(void *p) { return static_cast<TracePlayer *>(p)->runSetup() ? 0 : -1; }
// one steady-state frame: records and submits every pass of the FRAME section; sync != 0 waits for the GPU
int sglp_frame(void *p, int sync) { return static_cast<TracePlayer *>(p)->runFrame(sync != 0) ? 0 : -1; }
int sglp_tail(void *p, const char *out_path) {
  auto *tp = static_cast<TracePlayer *>(p);
  tp->setOutput(out_path ? out_path : "");
  return tp->runTail() ? 0 : -1;
}
// read back the attachment recorded under `tag` (resolved colour for multisample targets) into host memory
int sglp_readback(void *p, const char *tag, void *dst, size_t cap, int *w, int *h, int *format, int *samples) {
  PlayerBackend::Blob b;
  if (!static_cast<TracePlayer *>(p)->readbackTagged(tag, b)) return -1;
  if (w) *w = b.width;
  if (h) *h = b.height;
  if (format) *format = b.format;
  if (samples) *samples = b.samples;
  if (dst) {
    if (b.data.size() > cap) return -2;
    memcpy(dst, b.data.data(), b.data.size());
  }
  return (int) (b.data.size() >> 0 > 0x7fffffff ? 0x7fffffff : b.data.size());
}
// C-ABI texture handle of the attachment recorded under `tag` (for zero-copy read-back into pinned memory)
int sglp_texture_handle(void *p, const char *tag) {
  return static_cast<TracePlayer *>(p)->textureHandleTagged(tag);
}
const char *sglp_backend_name(void) { return PlayerBackend::name(); }

}  // extern "C"
