// api.cu -- the C-ABI of libpigo_b200.so (include/pigo_b200.h): cascade parsing, scan planning on the
// host (float64 ladder arithmetic exactly as core/pigo.go:226-231,:255), scratch management and the
// kernel sequence of each entry point.  No CPU compute path exists: without a device the calls fail.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "host.h"

namespace pigo {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
static std::atomic<int> g_device{-1};
static int g_num_sms = 148;
Options g_opt;

// ---- per-kernel event timing --------------------------------------------------------------------------------
struct TimingSlot {
  std::vector<cudaEvent_t> ev;  // begin/end pairs
  size_t used = 0;
};
static TimingSlot g_tslot[T_NSLOTS];
static std::mutex g_tmu;
static const char* kSlotNames[T_NSLOTS] = {"tiled", "gather", "deep", "finalize", "cluster", "puploc", "gray"};

void timing_reset() {
  std::lock_guard<std::mutex> g(g_tmu);
  for (auto& s : g_tslot) s.used = 0;
}
static void timing_mark(int slot, cudaStream_t st) {
  if (!g_opt.timing.load()) return;
  std::lock_guard<std::mutex> g(g_tmu);
  TimingSlot& s = g_tslot[slot];
  if (s.used >= 1u << 16) return;
  if (s.used == s.ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
    s.ev.push_back(e);
  }
  cudaEventRecord(s.ev[s.used++], st);
}
void timing_begin(int slot, cudaStream_t st) { timing_mark(slot, st); }
void timing_end(int slot, cudaStream_t st) { timing_mark(slot, st); }
long long timing_query(const std::string& key) {
  std::lock_guard<std::mutex> g(g_tmu);
  for (int i = 0; i < T_NSLOTS; ++i) {
    const std::string base = std::string("t_") + kSlotNames[i];
    TimingSlot& s = g_tslot[i];
    if (key == base + "_n") return (long long)(s.used / 2);
    if (key == base + "_ns") {
      double total = 0;
      for (size_t k = 0; k + 1 < s.used; k += 2) {
        float ms = 0;
        if (cudaEventSynchronize(s.ev[k + 1]) != cudaSuccess || cudaEventElapsedTime(&ms, s.ev[k], s.ev[k + 1]) != cudaSuccess) {
          cudaGetLastError();
          return -2;
        }
        total += (double)ms * 1e6;
      }
      return (long long)total;
    }
  }
  return -1;
}

// A workspace can be handed to another caller (another thread / stream) while the asynchronous work of its previous
// user is still running on that user's stream: order every new user behind the last recorded use.
static int ws_enter(Workspace* w, cudaStream_t st) {
  if (!w->busy && cudaEventCreateWithFlags(&w->busy, cudaEventDisableTiming) != cudaSuccess) return set_err(PIGO_E_CUDA, "event creation failed");
  if (w->busy_valid && cudaStreamWaitEvent(st, w->busy, 0) != cudaSuccess) return set_err(PIGO_E_CUDA, "cudaStreamWaitEvent failed");
  w->active_stream = st; w->active_stream_set = true;   // WsGuard records `busy` on it when the call ends, however it ends
  return PIGO_OK;
}
static int ws_leave_async(Workspace* w, cudaStream_t st) {
  if (cudaEventRecord(w->busy, st) != cudaSuccess) return set_err(PIGO_E_CUDA, "cudaEventRecord failed");
  w->busy_valid = true;
  w->active_stream_set = false;
  return PIGO_OK;
}

int set_err(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return set_err(PIGO_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

static int ensure_device() {
  int dev = g_device.load();
  if (dev < 0) {
    int rc = pigo_init(0);
    if (rc != PIGO_OK) return rc;
    dev = g_device.load();
  }
  CUDA_TRY(cudaSetDevice(dev));
  return PIGO_OK;
}

// ---- scale ladder / grid: core/pigo.go:226-231,:255 ---------------------------------------------------
static int build_plan(int rows, int cols, int min_size, int max_size, double shift, double scale_factor,
                      std::vector<ScaleEntry>& plan, uint64_t& total) {
  plan.clear();
  total = 0;
  long long scale = min_size;
  int guard = 0;
  while (scale <= max_size) {
    if (++guard > (1 << 16)) return set_err(PIGO_E_INVALID, "scale ladder longer than 65536 entries");
    if (scale > 0) {
      const long long step = (long long)std::fmax(shift * (double)scale, 1.0);  // :227
      const long long off = scale / 2 + 1;                                      // :228
      long long nr = 0, nc = 0;
      if (rows - off >= off) nr = (rows - off - off) / step + 1;                // :230
      if (cols - off >= off) nc = (cols - off - off) / step + 1;                // :231
      if (nr > 0 && nc > 0) {
        ScaleEntry e;
        e.s = (int)scale; e.step = (int)step; e.off = (int)off; e.nrows = (int)nr; e.ncols = (int)nc;
        e.wbase = (uint32_t)total; e.nwin = (uint32_t)(nr * nc); e.pad = 0;
        total += (uint64_t)(nr * nc);
        if (total > 0x7fffffffull) return set_err(PIGO_E_INVALID, "more than 2^31 windows per frame");
        plan.push_back(e);
      }
    } else {
      // scale <= 0: offset <= 1 and step >= 1; the reference would index out of bounds. Reject.
      return set_err(PIGO_E_INVALID, "MinSize must be positive");
    }
    const double next = (double)scale + std::fmax(2.0, ((double)scale * scale_factor) - (double)scale);  // :255
    if (!(next < 9.0e15)) break;
    scale = (long long)next;
  }
  return PIGO_OK;
}

}  // namespace pigo

using namespace pigo;

// =========================================================================================================
extern "C" {

const char* pigo_last_error(void) { return g_err.c_str(); }
int pigo_version(void) { return PIGO_B200_VERSION; }
int64_t pigo_launch_count(void) { return g_launches.load(); }

int pigo_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_err(PIGO_E_NODEVICE, "no CUDA device visible (%s); libpigo_b200 has no CPU fallback",
                   e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return set_err(PIGO_E_INVALID, "device %d out of range (0..%d)", device, n - 1);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_err(PIGO_E_NODEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));
  g_num_sms = prop.multiProcessorCount;
  g_device.store(device);
  return PIGO_OK;
}

int pigo_shutdown(void) { return PIGO_OK; }

int pigo_alloc_pinned(void** ptr, size_t bytes) {
  if (!ptr) return set_err(PIGO_E_INVALID, "null ptr");
  int rc = ensure_device();
  if (rc) return rc;
  CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return PIGO_OK;
}
int pigo_free_pinned(void* ptr) {
  if (ptr) CUDA_TRY(cudaFreeHost(ptr));
  return PIGO_OK;
}

int pigo_device_alloc(void** ptr, size_t bytes) {
  if (!ptr) return set_err(PIGO_E_INVALID, "null ptr");
  int rc = ensure_device();
  if (rc) return rc;
  if (cudaMalloc(ptr, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return set_err(PIGO_E_NOMEM, "cudaMalloc(%zu) failed", bytes); }
  return PIGO_OK;
}
int pigo_device_free(void* ptr) {
  if (ptr) CUDA_TRY(cudaFree(ptr));
  return PIGO_OK;
}
int pigo_device_upload(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return PIGO_OK;
  if (!dst || !src) return set_err(PIGO_E_INVALID, "null argument");
  int rc = ensure_device();
  if (rc) return rc;
  CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return PIGO_OK;
}

int pigo_set_option(const char* name, int64_t value) {
  if (!name) return set_err(PIGO_E_INVALID, "null option name");
  return g_opt.set(name, value) ? PIGO_OK : set_err(PIGO_E_INVALID, "unknown option '%s'", name);
}
int64_t pigo_get_option(const char* name) { return name ? g_opt.get(name) : -1; }

// ---- (*Pigo).Unpack, core/pigo.go:51-110 ----------------------------------------------------------------
static uint32_t rd_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

int pigo_cascade_create(const uint8_t* packet, size_t len, pigo_cascade** out) {
  if (!packet || !out) return set_err(PIGO_E_INVALID, "null argument");
  if (len < 16) return set_err(PIGO_E_INVALID, "cascade packet shorter than its 16-byte header");
  const uint32_t depth = rd_u32(packet + 8);    // :64 (bytes 0..7 are skipped, :61)
  const uint32_t ntrees = rd_u32(packet + 12);  // :68
  if (depth < 1 || depth > 12) return set_err(PIGO_E_INVALID, "unsupported tree depth %u", depth);
  const size_t leaves = (size_t)1 << depth;
  const size_t per_tree = (4 * leaves - 4) + 4 * leaves + 4;
  if (ntrees > (1u << 20) || len < 16 + (size_t)ntrees * per_tree)
    return set_err(PIGO_E_INVALID, "cascade packet truncated: %zu bytes, need %zu", len, 16 + (size_t)ntrees * per_tree);
  int rc = ensure_device();
  if (rc) return rc;
  std::vector<int8_t> codes((size_t)ntrees * 4 * leaves + 16, 0);
  std::vector<float> preds((size_t)ntrees * leaves + 4, 0.f), thr(ntrees + 4, 0.f);
  size_t pos = 16;
  for (uint32_t t = 0; t < ntrees; ++t) {
    memcpy(codes.data() + (size_t)t * 4 * leaves + 4, packet + pos, 4 * leaves - 4);  // 4 zero bytes first, :79-86
    pos += 4 * leaves - 4;
    memcpy(preds.data() + (size_t)t * leaves, packet + pos, 4 * leaves);              // :89-95 (LE f32 bit copy)
    pos += 4 * leaves;
    memcpy(thr.data() + t, packet + pos, 4);                                          // :96-100
    pos += 4;
  }
  pigo_cascade* c = new pigo_cascade();
  c->depth = depth; c->ntrees = ntrees; c->leaves = (uint32_t)leaves;
  cudaGetDevice(&c->device);
  if ((rc = c->codes.reserve(codes.size())) || (rc = c->preds.reserve(preds.size() * 4)) || (rc = c->thresh.reserve(thr.size() * 4))) {
    delete c; return rc;
  }
  cudaMemcpy(c->codes.p, codes.data(), codes.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(c->preds.p, preds.data(), preds.size() * 4, cudaMemcpyHostToDevice);
  cudaError_t e = cudaMemcpy(c->thresh.p, thr.data(), thr.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete c; return set_err(PIGO_E_CUDA, "table upload failed: %s", cudaGetErrorString(e)); }
  c->tab.codes = (const int8_t*)c->codes.p; c->tab.preds = (const float*)c->preds.p; c->tab.thresh = (const float*)c->thresh.p;
  c->tab.depth = (int)depth; c->tab.ntrees = (int)ntrees; c->tab.leaves = (int)leaves;
  rc = build_tiled_tables(c->tab, codes, preds, thr, c->tiled_tab);
  if (rc) { delete c; return rc; }
  *out = c;
  return PIGO_OK;
}

void pigo_cascade_destroy(pigo_cascade* c) {
  if (!c) return;
  c->codes.release(); c->preds.release(); c->thresh.release(); c->tiled_tab.release();
  delete c;
}

int pigo_cascade_info(const pigo_cascade* c, uint32_t* tree_depth, uint32_t* tree_num) {
  if (!c) return set_err(PIGO_E_INVALID, "null cascade");
  if (tree_depth) *tree_depth = c->depth;
  if (tree_num) *tree_num = c->ntrees;
  return PIGO_OK;
}

int pigo_scale_ladder(int min_size, int max_size, double scale_factor, int* scales, int cap, int* n_out) {
  long long scale = min_size;
  int n = 0;
  while (scale <= max_size) {
    if (scales && n < cap) scales[n] = (int)scale;
    if (++n > (1 << 16)) return set_err(PIGO_E_INVALID, "scale ladder longer than 65536 entries");
    scale = (long long)((double)scale + std::fmax(2.0, ((double)scale * scale_factor) - (double)scale));
  }
  if (n_out) *n_out = n;
  return (scales && n > cap) ? set_err(PIGO_E_CAP, "ladder has %d entries", n) : PIGO_OK;
}

int64_t pigo_count_windows(int rows, int cols, int min_size, int max_size, double shift_factor, double scale_factor) {
  std::vector<ScaleEntry> plan;
  uint64_t total = 0;
  if (build_plan(rows, cols, min_size, max_size, shift_factor, scale_factor, plan, total) != PIGO_OK) return -1;
  return (int64_t)total;
}

int pigo_describe_plan(int rows, int cols, int min_size, int max_size, double shift_factor, double scale_factor, char* json, size_t cap) {
  if (!json) return set_err(PIGO_E_INVALID, "null buffer");
  std::vector<ScaleEntry> plan;
  uint64_t total = 0;
  int rc = build_plan(rows, cols, min_size, max_size, shift_factor, scale_factor, plan, total);
  if (rc) return rc;
  return describe_plan(plan, total, 468, json, cap);   // geometry of the stock 468-tree facefinder cascade
}

// ---- RunCascade ------------------------------------------------------------------------------------------
int pigo_run_cascade_batch(const pigo_cascade* cc, const uint8_t* frames, int nframes, size_t frame_stride, int rows, int cols,
                           int dim, int min_size, int max_size, double shift_factor, double scale_factor, double angle,
                           pigo_det* out, int cap_per_frame, int* n_out, unsigned flags, void* stream_) {
  pigo_cascade* c = const_cast<pigo_cascade*>(cc);
  if (!c || !n_out || (!out && cap_per_frame > 0)) return set_err(PIGO_E_INVALID, "null argument");
  if (nframes < 0 || nframes > 65535) return set_err(PIGO_E_INVALID, "nframes must be 0..65535");
  if (rows < 0 || cols < 0 || dim < cols) return set_err(PIGO_E_INVALID, "bad geometry rows=%d cols=%d dim=%d", rows, cols, dim);
  if (cap_per_frame < 0) return set_err(PIGO_E_INVALID, "negative capacity");
  if ((uint64_t)rows * (uint64_t)dim > 0x7fffffffull) return set_err(PIGO_E_INVALID, "frames larger than 2^31 bytes are not supported");
  if (nframes > 0 && !frames) return set_err(PIGO_E_INVALID, "null frames");
  if (nframes > 1 && frame_stride < (size_t)rows * dim - (size_t)(dim - cols))
    return set_err(PIGO_E_INVALID, "frame_stride smaller than a frame");
  int rc = ensure_device();
  if (rc) return rc;
  const bool frames_dev = flags & PIGO_FRAMES_DEVICE, out_dev = flags & PIGO_OUT_DEVICE;
  if (nframes == 0) return PIGO_OK;

  WsGuard g(c->pool);
  Workspace* w = g.w;
  if (!w) return set_err(PIGO_E_CUDA, "could not create a CUDA stream");
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : w->stream;
  if ((rc = ws_enter(w, st))) return rc;

  // plan (cached per workspace)
  if (w->p_rows != rows || w->p_cols != cols || w->p_min != min_size || w->p_max != max_size || w->p_shift != shift_factor ||
      w->p_scale != scale_factor) {
    rc = build_plan(rows, cols, min_size, max_size, shift_factor, scale_factor, w->plan_host, w->wins);
    if (rc) { w->p_rows = -1; return rc; }
    if ((rc = w->plan.reserve((w->plan_host.size() + 1) * sizeof(ScaleEntry)))) return rc;
    if (!w->plan_host.empty())
      CUDA_TRY(cudaMemcpyAsync(w->plan.p, w->plan_host.data(), w->plan_host.size() * sizeof(ScaleEntry), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));  // plan_host may be rebuilt by the next call before the copy ran
    w->pad_first_untiled = -1;
    w->p_rows = rows; w->p_cols = cols; w->p_min = min_size; w->p_max = max_size; w->p_shift = shift_factor; w->p_scale = scale_factor;
  }
  const int nscales = (int)w->plan_host.size();
  const int cap = cap_per_frame > 0 ? cap_per_frame : 1;

  // device buffers
  const uint8_t* d_frames = frames;
  const size_t frame_bytes = (size_t)rows * dim;
  size_t d_stride = frame_stride;
  if (!frames_dev) {
    d_stride = (frame_bytes + 255) & ~(size_t)255;
    if ((rc = w->frames.reserve(d_stride * nframes + 256))) return rc;
    d_frames = (const uint8_t*)w->frames.p;
  }
  pigo_det* d_out = out;
  int32_t* d_nout = n_out;
  if (!out_dev) {
    if ((rc = w->out.reserve((size_t)nframes * cap * sizeof(pigo_det)))) return rc;
    if ((rc = w->nout.reserve((size_t)nframes * sizeof(int32_t)))) return rc;
    d_out = (pigo_det*)w->out.p;
    d_nout = (int32_t*)w->nout.p;
  }
  if ((rc = w->raw.reserve((size_t)nframes * cap * sizeof(RawDet)))) return rc;
  bool zero_out_staging = !out_dev;   // slots past a frame's count travel back to the host too: keep them defined (zero)

  // Sub-batch pipeline: the batch is cut into groups of `sub_batch` frames that alternate between `lanes` internal
  // streams.  (1) The deferred queues (Q1/Q2) of a group are consumed while its frames are still L2-resident;
  // (2) the tail kernels of one group overlap the bulk kernels of the next; (3) with host frames, the H2D copy of
  // group k+1 overlaps the scan of group k.
  // Group size: measured on 256 x 1080p -- resident frames 10.3 ms/step with 128-frame groups vs 10.9 with 64 (fewer
  // kernel tails); host frames 13.4 ms with 64 vs 15.3 with 128 (the first group's copy is not overlapped).
  int sub = (int)g_opt.sub_batch.load();
  if (sub <= 0) sub = frames_dev ? 128 : 64;
  if (sub > nframes) sub = nframes;
  const int nsub = (nframes + sub - 1) / sub;
  int lanes = (int)std::min<long long>(std::max<long long>(1, g_opt.lanes.load()), kMaxLanes);
  if (nsub == 1) lanes = 1;
  if ((rc = w->ensure_lanes(lanes))) return rc;

  // counters: [0..nframes) raw counts | 8 x u64 work counters per sub-batch
  const size_t work_off = ((size_t)nframes * 4 + 15) & ~(size_t)15;
  const size_t cnt_bytes = work_off + (size_t)nsub * 64 + 64;
  if ((rc = w->counters.reserve(cnt_bytes))) return rc;
  CUDA_TRY(cudaMemsetAsync(w->counters.p, 0, cnt_bytes, st));
  int32_t* d_rawcount = (int32_t*)w->counters.p;
  unsigned long long* d_work = (unsigned long long*)((char*)w->counters.p + work_off);

  if (lanes > 1 || !frames_dev) {
    CUDA_TRY(cudaEventRecord(w->ev_fork, st));   // everything queued on `st` so far (previous results, memset) comes first
    for (int l = 0; l < lanes && lanes > 1; ++l) CUDA_TRY(cudaStreamWaitEvent(w->lane_stream[l], w->ev_fork, 0));
    if (!frames_dev) {
      if (!w->copy_stream && cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
        return set_err(PIGO_E_CUDA, "stream creation failed");
      CUDA_TRY(cudaStreamWaitEvent(w->copy_stream, w->ev_fork, 0));
    }
  }
  int rot_slot = -1;
  if (angle > 0.0) {                      // core/pigo.go:232-236
    const double a = angle > 1.0 ? 1.0 : angle;
    rot_slot = (int)(32.0 * a);           // :159
  }
  for (int k = 0; k < nsub; ++k) {
    const int f0 = k * sub, nf = std::min(sub, nframes - f0);
    const int lane = k % lanes;
    cudaStream_t ls = lanes > 1 ? w->lane_stream[lane] : st;
    if (!frames_dev) {
      // the copy of group k runs on the copy stream and overlaps the scan of group k-1 (pinned source memory)
      uint8_t* dst = (uint8_t*)w->frames.p + (size_t)f0 * d_stride;
      const uint8_t* src = frames + (size_t)f0 * frame_stride;
      cudaStream_t cs = w->copy_stream;
      if (frame_stride == d_stride || nf == 1) {
        CUDA_TRY(cudaMemcpyAsync(dst, src, nf == 1 ? frame_bytes : d_stride * (nf - 1) + frame_bytes, cudaMemcpyHostToDevice, cs));
      } else {
        CUDA_TRY(cudaMemcpy2DAsync(dst, d_stride, src, frame_stride, frame_bytes, nf, cudaMemcpyHostToDevice, cs));
      }
      cudaEvent_t ev = w->copy_event(k);
      if (!ev) return set_err(PIGO_E_CUDA, "event creation failed");
      CUDA_TRY(cudaEventRecord(ev, cs));
      CUDA_TRY(cudaStreamWaitEvent(ls, ev, 0));
    }
    if (nscales > 0 && c->ntrees > 0) {
      ScanArgs A{};
      A.tab = c->tab;
      A.frames = d_frames + (size_t)f0 * d_stride; A.frame_stride = d_stride; A.nframes = nf; A.rows = rows; A.cols = cols; A.dim = dim;
      A.plan = (const ScaleEntry*)w->plan.p; A.nscales = nscales; A.wins_per_frame = (uint32_t)w->wins;
      A.rot_slot = rot_slot;
      A.raw = (RawDet*)w->raw.p + (size_t)f0 * cap; A.raw_count = d_rawcount + f0; A.cap = cap;
      rc = run_scan(c, w, lane, A, d_work + 8 * (size_t)k, ls, g_num_sms);
      if (rc) return rc;
    }
  }
  if (lanes > 1) {
    for (int l = 0; l < lanes; ++l) {
      CUDA_TRY(cudaEventRecord(w->ev_join[l], w->lane_stream[l]));
      CUDA_TRY(cudaStreamWaitEvent(st, w->ev_join[l], 0));
    }
  }
  if (zero_out_staging) CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)nframes * cap * sizeof(pigo_det), st));
  timing_begin(T_FINALIZE, st);
  launch_finalize((const RawDet*)w->raw.p, d_rawcount, cap, (const ScaleEntry*)w->plan.p, nscales, d_out, d_nout, nframes, st);
  timing_end(T_FINALIZE, st);
  g_launches++;
  CUDA_TRY(cudaGetLastError());

  if (out_dev) return ws_leave_async(w, st);
  CUDA_TRY(cudaMemcpyAsync(n_out, d_nout, (size_t)nframes * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  int need_more = 0;
  size_t max_n = 0;
  for (int f = 0; f < nframes; ++f) {
    if (n_out[f] > cap_per_frame) need_more = 1;
    size_t k = (size_t)std::min(n_out[f], cap_per_frame);
    if (k > max_n) max_n = k;
  }
  if (cap_per_frame > 0 && max_n > 0) {
    // copy only the used prefix of every frame's slice
    CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)cap_per_frame * sizeof(pigo_det), d_out, (size_t)cap * sizeof(pigo_det),
                               max_n * sizeof(pigo_det), nframes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  if (need_more) return set_err(PIGO_E_CAP, "output capacity %d per frame too small", cap_per_frame);
  return PIGO_OK;
}

int pigo_run_cascade(const pigo_cascade* c, const uint8_t* pixels, int rows, int cols, int dim, int min_size, int max_size,
                     double shift_factor, double scale_factor, double angle, pigo_det* out, int cap, int* n_out) {
  return pigo_run_cascade_batch(c, pixels, 1, (size_t)rows * (size_t)dim, rows, cols, dim, min_size, max_size, shift_factor,
                                scale_factor, angle, out, cap, n_out, PIGO_MEM_HOST, nullptr);
}

// ---- ClusterDetections ---------------------------------------------------------------------------------
static WorkspacePool g_cluster_pool;

int pigo_cluster_batch(pigo_det* dets, const int* n, int nframes, int cap_per_frame, double iou_threshold, pigo_det* out,
                       int out_cap_per_frame, int* n_out, unsigned flags, void* stream_) {
  if (!n || !n_out || nframes < 0) return set_err(PIGO_E_INVALID, "null argument");
  if (cap_per_frame < 0 || out_cap_per_frame < 0) return set_err(PIGO_E_INVALID, "negative capacity");
  int rc = ensure_device();
  if (rc) return rc;
  if (nframes == 0) return PIGO_OK;
  const bool dev = (flags & PIGO_OUT_DEVICE) != 0;  // dets, n, out, n_out all on the device
  WsGuard g(g_cluster_pool);
  Workspace* w = g.w;
  if (!w) return set_err(PIGO_E_CUDA, "could not create a CUDA stream");
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : w->stream;
  if ((rc = ws_enter(w, st))) return rc;
  const int cap = std::max(cap_per_frame, 1), ocap = std::max(out_cap_per_frame, 1);
  const size_t nd = (size_t)nframes * cap;
  if ((rc = w->scratch_a.reserve(nd * sizeof(pigo_det)))) return rc;
  if ((rc = w->scratch_b.reserve(nd))) return rc;
  if ((rc = w->scratch_c.reserve(nd * sizeof(int32_t)))) return rc;
  pigo_det *d_dets = dets, *d_out = out;
  const int32_t* d_n = n;
  int32_t* d_nout = n_out;
  if (!dev) {
    if ((rc = w->raw.reserve(nd * sizeof(pigo_det)))) return rc;
    if ((rc = w->out.reserve((size_t)nframes * ocap * sizeof(pigo_det)))) return rc;
    if ((rc = w->nout.reserve((size_t)nframes * 8))) return rc;
    d_dets = (pigo_det*)w->raw.p; d_out = (pigo_det*)w->out.p;
    d_nout = (int32_t*)w->nout.p;
    int32_t* dn = d_nout + nframes;
    d_n = dn;
    if (cap_per_frame > 0 && dets) CUDA_TRY(cudaMemcpyAsync(d_dets, dets, nd * sizeof(pigo_det), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dn, n, (size_t)nframes * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)nframes * ocap * sizeof(pigo_det), st));   // defined padding in the host copy
  }
  timing_begin(T_CLUSTER, st);
  launch_cluster(d_dets, d_n, cap, iou_threshold, (pigo_det*)w->scratch_a.p, (uint8_t*)w->scratch_b.p, (int32_t*)w->scratch_c.p,
                 d_out, ocap, d_nout, nframes, st);
  timing_end(T_CLUSTER, st);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  if (dev) return ws_leave_async(w, st);
  if (cap_per_frame > 0 && dets) CUDA_TRY(cudaMemcpyAsync(dets, d_dets, nd * sizeof(pigo_det), cudaMemcpyDeviceToHost, st));  // in-place sort
  CUDA_TRY(cudaMemcpyAsync(n_out, d_nout, (size_t)nframes * 4, cudaMemcpyDeviceToHost, st));
  if (out_cap_per_frame > 0 && out)
    CUDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)nframes * ocap * sizeof(pigo_det), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int f = 0; f < nframes; ++f)
    if (n_out[f] > out_cap_per_frame) return set_err(PIGO_E_CAP, "cluster capacity %d too small", out_cap_per_frame);
  return PIGO_OK;
}

int pigo_cluster(pigo_det* dets, int n, double iou_threshold, pigo_det* out, int cap, int* n_out) {
  if (n < 0) return set_err(PIGO_E_INVALID, "negative n");
  if (n == 0) { if (n_out) *n_out = 0; return ensure_device(); }
  return pigo_cluster_batch(dets, &n, 1, n, iou_threshold, out, cap, n_out, PIGO_MEM_HOST, nullptr);
}

// ---- UnpackCascade / RunDetector / GetLandmarkPoint --------------------------------------------------------
int pigo_puploc_create(const uint8_t* packet, size_t len, pigo_puploc** out) {
  if (!packet || !out) return set_err(PIGO_E_INVALID, "null argument");
  if (len < 16) return set_err(PIGO_E_INVALID, "puploc packet shorter than its 16-byte header");
  const uint32_t stages = rd_u32(packet + 0);   // core/puploc.go:51
  float scales; memcpy(&scales, packet + 4, 4); // :55-57
  const uint32_t trees = rd_u32(packet + 8);    // :61
  const uint32_t depth = rd_u32(packet + 12);   // :65
  if (depth < 1 || depth > 16 || stages > 4096 || trees > 65536) return set_err(PIGO_E_INVALID, "implausible puploc header");
  const size_t leaves = (size_t)1 << depth, ncode = 4 * leaves - 4, npred = 2 * leaves;
  const size_t nt = (size_t)stages * trees;
  if (len < 16 + nt * (ncode + npred * 4)) return set_err(PIGO_E_INVALID, "puploc packet truncated");
  int rc = ensure_device();
  if (rc) return rc;
  std::vector<int8_t> codes(nt * ncode + 16);
  std::vector<float> preds(nt * npred + 4);
  size_t pos = 16;
  for (size_t t = 0; t < nt; ++t) {
    memcpy(codes.data() + t * ncode, packet + pos, ncode); pos += ncode;          // :75-80
    memcpy(preds.data() + t * npred, packet + pos, npred * 4); pos += npred * 4;  // :83-91
  }
  pigo_puploc* p = new pigo_puploc();
  if ((rc = p->codes.reserve(codes.size())) || (rc = p->preds.reserve(preds.size() * 4))) { delete p; return rc; }
  cudaMemcpy(p->codes.p, codes.data(), codes.size(), cudaMemcpyHostToDevice);
  cudaError_t e = cudaMemcpy(p->preds.p, preds.data(), preds.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete p; return set_err(PIGO_E_CUDA, "table upload failed: %s", cudaGetErrorString(e)); }
  p->tab.codes = (const int8_t*)p->codes.p; p->tab.preds = (const float*)p->preds.p;
  p->tab.stages = (int)stages; p->tab.trees = (int)trees; p->tab.depth = (int)depth; p->tab.leaves = (int)leaves;
  p->tab.scales = scales;
  *out = p;
  return PIGO_OK;
}

void pigo_puploc_destroy(pigo_puploc* p) {
  if (!p) return;
  p->codes.release(); p->preds.release();
  delete p;
}

int pigo_puploc_info(const pigo_puploc* p, uint32_t* stages, float* scale_mul, uint32_t* trees, uint32_t* depth) {
  if (!p) return set_err(PIGO_E_INVALID, "null cascade");
  if (stages) *stages = p->tab.stages;
  if (scale_mul) *scale_mul = p->tab.scales;
  if (trees) *trees = p->tab.trees;
  if (depth) *depth = p->tab.depth;
  return PIGO_OK;
}

int pigo_puploc_run(const pigo_puploc* pc, const pigo_point* seeds, int nseeds, const float* randoms, uint64_t rng_seed,
                    const uint8_t* pixels, int rows, int cols, int dim, double angle, const uint8_t* flipv, pigo_point* out,
                    unsigned flags, void* stream_) {
  return pigo_puploc_run_frames(pc, seeds, nseeds, nullptr, randoms, rng_seed, pixels, 1, (size_t)std::max(rows, 0) * (size_t)std::max(dim, 0),
                                rows, cols, dim, angle, flipv, out, flags, stream_);
}

int pigo_puploc_run_frames(const pigo_puploc* pc, const pigo_point* seeds, int nseeds, const int32_t* seed_frame, const float* randoms,
                           uint64_t rng_seed, const uint8_t* pixels, int nframes, size_t frame_stride, int rows, int cols, int dim,
                           double angle, const uint8_t* flipv, pigo_point* out, unsigned flags, void* stream_) {
  pigo_puploc* p = const_cast<pigo_puploc*>(pc);
  if (!p || !seeds || !out || !pixels) return set_err(PIGO_E_INVALID, "null argument");
  if (nseeds < 0 || rows <= 0 || cols <= 0 || dim < cols || nframes < 1) return set_err(PIGO_E_INVALID, "bad geometry");
  if (nframes > 1 && !seed_frame) return set_err(PIGO_E_INVALID, "seed_frame is required when nframes > 1");
  int rc = ensure_device();
  if (rc) return rc;
  if (nseeds == 0) return PIGO_OK;
  const bool frames_dev = flags & PIGO_FRAMES_DEVICE, out_dev = flags & PIGO_OUT_DEVICE;
  if (!out_dev)
    for (int i = 0; i < nseeds; ++i) {
      if (seeds[i].perturbs < 0 || seeds[i].perturbs > 63)
        return set_err(PIGO_E_INVALID, "seed %d: Perturbs=%d outside 0..63 (the reference panics, core/puploc.go:261)", i, seeds[i].perturbs);
      if (seed_frame && (seed_frame[i] < 0 || seed_frame[i] >= nframes)) return set_err(PIGO_E_INVALID, "seed %d: frame index out of range", i);
    }
  WsGuard g(p->pool);
  Workspace* w = g.w;
  if (!w) return set_err(PIGO_E_CUDA, "could not create a CUDA stream");
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : w->stream;
  if ((rc = ws_enter(w, st))) return rc;
  const uint8_t* d_pix = pixels;
  if (!frames_dev) {
    const size_t bytes = frame_stride * (size_t)(nframes - 1) + (size_t)rows * dim;
    if ((rc = w->frames.reserve(bytes))) return rc;
    CUDA_TRY(cudaMemcpyAsync(w->frames.p, pixels, bytes, cudaMemcpyHostToDevice, st));
    d_pix = (const uint8_t*)w->frames.p;
  }
  const int32_t* d_sf = seed_frame;
  const pigo_point* d_seeds = seeds;
  pigo_point* d_out = out;
  const float* d_rnd = randoms;
  const uint8_t* d_flip = flipv;
  if (!out_dev) {
    const size_t sb = (size_t)nseeds * sizeof(pigo_point), rb = (size_t)nseeds * 63 * 3 * sizeof(float);
    if ((rc = w->scratch_a.reserve(2 * sb))) return rc;
    d_seeds = (const pigo_point*)w->scratch_a.p;
    d_out = (pigo_point*)((char*)w->scratch_a.p + sb);
    CUDA_TRY(cudaMemcpyAsync(w->scratch_a.p, seeds, sb, cudaMemcpyHostToDevice, st));
    if (randoms) {
      if ((rc = w->scratch_b.reserve(rb))) return rc;
      CUDA_TRY(cudaMemcpyAsync(w->scratch_b.p, randoms, rb, cudaMemcpyHostToDevice, st));
      d_rnd = (const float*)w->scratch_b.p;
    }
    if (flipv) {
      if ((rc = w->scratch_c.reserve(nseeds))) return rc;
      CUDA_TRY(cudaMemcpyAsync(w->scratch_c.p, flipv, nseeds, cudaMemcpyHostToDevice, st));
      d_flip = (const uint8_t*)w->scratch_c.p;
    }
    if (seed_frame) {
      if ((rc = w->nout.reserve((size_t)nseeds * 4))) return rc;
      CUDA_TRY(cudaMemcpyAsync(w->nout.p, seed_frame, (size_t)nseeds * 4, cudaMemcpyHostToDevice, st));
      d_sf = (const int32_t*)w->nout.p;
    }
  }
  int rot_slot = -1;
  if (angle > 0.0) rot_slot = (int)(32.0 * (angle > 1.0 ? 1.0 : angle));  // core/puploc.go:252-256, :166
  timing_begin(T_PUPLOC, st);
  launch_puploc(p->tab, d_seeds, nseeds, d_rnd, rng_seed, d_pix, d_sf, frame_stride, rows, cols, dim, rot_slot, d_flip, d_out, st);
  timing_end(T_PUPLOC, st);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  if (out_dev) return ws_leave_async(w, st);
  CUDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)nseeds * sizeof(pigo_point), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return PIGO_OK;
}

int pigo_get_landmark_point(const pigo_puploc* p, const pigo_point* left_eye, const pigo_point* right_eye, const uint8_t* pixels,
                            int rows, int cols, int dim, int perturb, int flipv, const float* randoms, uint64_t rng_seed,
                            pigo_point* out) {
  if (!left_eye || !right_eye) return set_err(PIGO_E_INVALID, "null argument");
  // core/flploc.go:37-50 (float64 on the host)
  const long long dx = (long long)(left_eye->row - right_eye->row) * (left_eye->row - right_eye->row);
  const long long dy = (long long)(left_eye->col - right_eye->col) * (left_eye->col - right_eye->col);
  const double dist = std::sqrt((double)(dx + dy));
  const double row = (double)(left_eye->row + right_eye->row) / 2.0 + 0.25 * dist;
  const double col = (double)(left_eye->col + right_eye->col) / 2.0 + 0.15 * dist;
  const double scale = 3.0 * dist;
  pigo_point seed;
  seed.row = (int)row; seed.col = (int)col; seed.scale = (float)scale; seed.perturbs = perturb;
  const uint8_t fl = flipv ? 1 : 0;
  return pigo_puploc_run(p, &seed, 1, randoms, rng_seed, pixels, rows, cols, dim, 0.0, &fl, out, PIGO_MEM_HOST, nullptr);
}

// ---- RgbToGrayscale, core/grayscale.go:8-23 ----------------------------------------------------------------
static WorkspacePool g_gray_pool;

int pigo_rgba_to_gray(const uint8_t* rgba, size_t npixels, uint8_t* gray, unsigned flags, void* stream_) {
  if (npixels == 0) return ensure_device();
  if (!rgba || !gray) return set_err(PIGO_E_INVALID, "null argument");
  int rc = ensure_device();
  if (rc) return rc;
  WsGuard g(g_gray_pool);
  Workspace* w = g.w;
  if (!w) return set_err(PIGO_E_CUDA, "could not create a CUDA stream");
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : w->stream;
  if ((rc = ws_enter(w, st))) return rc;   // a device-output call returns while the kernel still reads the staging buffer
  const bool in_dev = flags & PIGO_FRAMES_DEVICE, out_dev = flags & PIGO_OUT_DEVICE;
  const uint8_t* d_in = rgba;
  uint8_t* d_out = gray;
  if (!in_dev) {
    if ((rc = w->frames.reserve(npixels * 4))) return rc;
    CUDA_TRY(cudaMemcpyAsync(w->frames.p, rgba, npixels * 4, cudaMemcpyHostToDevice, st));
    d_in = (const uint8_t*)w->frames.p;
  }
  if (!out_dev) {
    if ((rc = w->out.reserve(npixels))) return rc;
    d_out = (uint8_t*)w->out.p;
  }
  const size_t want = (npixels / 16 + 255) / 256 + 1;
  const int grid = (int)std::min<size_t>(want, (size_t)g_num_sms * 16);
  timing_begin(T_GRAY, st);
  launch_gray(d_in, npixels, d_out, grid, st);
  timing_end(T_GRAY, st);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  if (out_dev) return PIGO_OK;
  CUDA_TRY(cudaMemcpyAsync(gray, d_out, npixels, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return PIGO_OK;
}

}  // extern "C"
