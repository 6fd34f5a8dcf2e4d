// scan_driver.cu -- host-side planning and sequencing of the scan kernels for one RunCascade batch:
//   tiled kernel  : scales grouped into bands whose pixel tiles fit a per-warp shared-memory buffer
//   gather kernel : the remaining (large) scales, every scale of the rotated path, non-depth-6 cascades
//   resume kernel : finishes the long-lived windows the tiled kernel parked in the deep queue
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "host.h"

namespace pigo {

static constexpr size_t kSmemPerCta = 232448;  // 227 KB opt-in maximum on sm_100

// Builds the shared-memory-layout cascade table: one 516-byte record per tree (depth 6 only).
int build_tiled_tables(const FaceTables& tab, const std::vector<int8_t>& codes, const std::vector<float>& preds,
                       const std::vector<float>& thr, DevBuf& out) {
  if (tab.depth != 6) return PIGO_OK;  // other depths use the gather kernel only
  const size_t n = (size_t)tab.ntrees;
  std::vector<uint8_t> rec(n * 516 + 64, 0);
  for (size_t t = 0; t < n; ++t) {
    memcpy(rec.data() + t * 516, codes.data() + t * 256, 256);
    memcpy(rec.data() + t * 516 + 256, preds.data() + t * 64, 256);
    memcpy(rec.data() + t * 516 + 512, thr.data() + t, 4);
  }
  int rc = out.reserve(rec.size());
  if (rc) return rc;
  if (cudaMemcpy(out.p, rec.data(), rec.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    return set_err(PIGO_E_CUDA, "tiled table upload failed");
  }
  return PIGO_OK;
}

struct TilePlan {
  int nbands = 0;
  TileBand band[kMaxBands];
  int first_untiled = 0;  // ladder index of the first scale left to the gather kernel
  uint32_t tiles_per_frame = 0;
};

static TilePlan plan_bands(const std::vector<ScaleEntry>& plan, uint32_t tile_bytes, int max_scale, int ratio_pct) {
  TilePlan tp;
  const int n = (int)plan.size();
  int a = 0;
  while (a < n && tp.nbands < kMaxBands) {
    const int s0 = plan[a].s;
    if (s0 > max_scale) break;
    // longest band [a, b) within the ratio whose tile still has a useful core
    int best_b = -1, best_core = 0;
    for (int b = a + 1; b <= n && b - a <= 32; ++b) {
      const int smax = plan[b - 1].s;
      if (smax > max_scale || (long long)smax * 100 > (long long)s0 * ratio_pct) break;
      const int halo_lo = (smax + 1) / 2, halo_hi = (127 * smax) >> 8;
      int core = 0;
      for (int c = 16; c <= 512; c += 16) {
        const int rows_t = halo_lo + c + halo_hi;
        const int pitch = (rows_t + 15) & ~15;
        if ((size_t)rows_t * pitch <= tile_bytes) core = c; else break;
      }
      const int step_max = plan[b - 1].step;
      if (core >= 32 && core >= 3 * step_max) { best_b = b; best_core = core; }
    }
    if (best_b < 0) break;
    TileBand B{};
    const int smax = plan[best_b - 1].s;
    B.scale_lo = a; B.nscales = best_b - a;
    B.halo_lo = (smax + 1) / 2;
    const int halo_hi = (127 * smax) >> 8;
    B.core = best_core;
    B.rows_t = B.halo_lo + B.core + halo_hi;
    B.pitch = (B.rows_t + 15) & ~15;
    int org = B.halo_lo % 16;
    const int off_min = plan[a].off;
    if (org > off_min) org -= 16;
    B.org_x = org;
    int max_c = 0, max_r = 0;
    for (int i = a; i < best_b; ++i) {
      max_c = std::max(max_c, plan[i].off + (plan[i].ncols - 1) * plan[i].step);
      max_r = std::max(max_r, plan[i].off + (plan[i].nrows - 1) * plan[i].step);
    }
    B.tiles_x = (max_c - B.org_x) / B.core + 1;
    const int tiles_y = max_r / B.core + 1;
    B.ntiles = B.tiles_x * tiles_y;
    tp.band[tp.nbands++] = B;
    tp.tiles_per_frame += (uint32_t)B.ntiles;
    a = best_b;
  }
  tp.first_untiled = a;
  return tp;
}

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(PIGO_E_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return PIGO_OK;
}

int run_scan(pigo_cascade* c, Workspace* w, ScanArgs& A, unsigned long long* d_work, cudaStream_t st, int num_sms) {
  int rc;
  const bool rot = A.rot_slot >= 0;
  const long long mode = g_opt.scan_mode.load();
  A.deep = nullptr; A.deep_count = (unsigned int*)(d_work + 3); A.deep_cap = 0; A.deep_tree = 0x7fffffff;
  A.chunk = (uint32_t)std::max<long long>(32, g_opt.chunk.load());

  // ---- tiled kernel over the small/medium scales
  int first_gather_scale = 0;
  const bool can_tile = !rot && A.tab.depth == 6 && c->tiled_tab.p != nullptr && mode != 1;
  if (can_tile) {
    const int W = (int)std::min<long long>(std::max<long long>(1, g_opt.tile_warps.load()), kTiledMaxThreads / 32);
    int ks = (int)std::min<long long>(std::max<long long>(4, g_opt.tile_ks.load()), A.tab.ntrees);
    ks &= ~3;
    if (ks < 4) ks = A.tab.ntrees;  // tiny cascades: everything resident (record size keeps 16-byte multiples only for ks%4==0)
    const int ni = (int)std::min<long long>(std::max<long long>(1, g_opt.tile_ni.load()), 4);
    const size_t casc_bytes = (size_t)ks * 516;
    const size_t tiles0 = (16 + casc_bytes + 127) & ~(size_t)127;
    if (tiles0 + 4096 * (size_t)W < kSmemPerCta && casc_bytes % 16 == 0) {
      const uint32_t tile_bytes = (uint32_t)(((kSmemPerCta - tiles0) / W) & ~(size_t)127);
      int max_scale = (int)g_opt.tile_max_scale.load();
      if (max_scale <= 0) max_scale = 1 << 30;
      const TilePlan tp = plan_bands(w->plan_host, tile_bytes, max_scale, (int)std::max<long long>(100, g_opt.tile_band_ratio.load()));
      if (tp.nbands > 0) {
        const uint64_t total_windows = (uint64_t)A.wins_per_frame * A.nframes;
        uint64_t deep_cap = std::min<uint64_t>(total_windows / 12 + 65536, 1ull << 26);
        if ((rc = w->deep.reserve(deep_cap * sizeof(DeepItem)))) return rc;
        TiledArgs T{};
        T.scan = A;
        T.scan.deep = (DeepItem*)w->deep.p; T.scan.deep_cap = (uint32_t)deep_cap;
        T.scan.chunk_counter = d_work + 1;
        T.tab_tiled = (const uint8_t*)c->tiled_tab.p;
        T.ks = ks; T.tile_bytes = tile_bytes; T.nbands = tp.nbands;
        T.tail_min = (int)g_opt.tile_tail_min.load();
        for (int b = 0; b < tp.nbands; ++b) T.band[b] = tp.band[b];
        T.tiles_per_frame = tp.tiles_per_frame;
        T.total_tiles = (unsigned long long)tp.tiles_per_frame * A.nframes;
        const bool aligned = (A.dim % 16 == 0) && (A.frame_stride % 16 == 0) && (((uintptr_t)A.frames) % 16 == 0);
        const long long grid = std::max<long long>(1, std::min<long long>(num_sms, (long long)((T.total_tiles + W - 1) / W)));
        const size_t smem = tiles0 + (size_t)tile_bytes * W;
        timing_begin(T_TILED, st);
        launch_scan_tiled(T, (int)grid, W * 32, smem, ni, aligned, st);
        timing_end(T_TILED, st);
        g_launches++;
        if ((rc = check_launch("tiled scan"))) return rc;
        first_gather_scale = tp.first_untiled;
        A.deep = T.scan.deep; A.deep_cap = T.scan.deep_cap;
      }
    }
  }

  // ---- gather kernel over whatever is left
  if (first_gather_scale < A.nscales) {
    ScanArgs G = A;
    G.scale_lo = first_gather_scale; G.scale_hi = A.nscales;
    const uint32_t w_lo = w->plan_host[first_gather_scale].wbase;
    const uint32_t span = A.wins_per_frame - w_lo;
    G.chunks_per_frame = (span + G.chunk - 1) / G.chunk;
    G.chunk_counter = d_work;
    int per_sm = (int)g_opt.gather_ctas_per_sm.load();
    if (per_sm <= 0) per_sm = gather_max_ctas_per_sm(A.tab.depth, rot);
    const unsigned long long total_chunks = (unsigned long long)G.chunks_per_frame * A.nframes;
    long long grid = std::max(1ll, std::min<long long>((long long)num_sms * per_sm, (long long)((total_chunks + 7) / 8)));
    timing_begin(T_GATHER, st);
    launch_scan_gather(G, (int)grid, st);
    timing_end(T_GATHER, st);
    g_launches++;
    if ((rc = check_launch("gather scan"))) return rc;
  }

  // ---- resume kernel for the deep queue
  if (A.deep != nullptr) {
    ScanArgs R = A;
    R.scale_lo = 0; R.scale_hi = A.nscales;
    R.chunk_counter = d_work + 2;
    R.chunks_per_frame = 0;
    const int per_sm = gather_max_ctas_per_sm(6, false);
    timing_begin(T_DEEP, st);
    launch_scan_resume(R, num_sms * std::min(per_sm, 4), st);
    timing_end(T_DEEP, st);
    g_launches++;
    if ((rc = check_launch("resume scan"))) return rc;
  }
  return PIGO_OK;
}

}  // namespace pigo
