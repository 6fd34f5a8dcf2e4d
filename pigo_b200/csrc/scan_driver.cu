// scan_driver.cu -- host-side sequencing of the scan kernels for one RunCascade batch.
#include "common.cuh"
#include "host.h"

namespace pigo {

int build_tiled_tables(const FaceTables&, const std::vector<int8_t>&, const std::vector<float>&, const std::vector<float>&,
                       DevBuf&) {
  return PIGO_OK;  // filled in by the tiled kernel
}

int run_scan(pigo_cascade* c, Workspace* w, ScanArgs& A, unsigned long long* d_work, cudaStream_t st, int num_sms) {
  (void)c; (void)w;
  A.scale_lo = 0;
  A.scale_hi = A.nscales;
  A.deep = nullptr; A.deep_count = nullptr; A.deep_cap = 0; A.deep_tree = 0x7fffffff;
  A.chunk = (uint32_t)std::max<long long>(32, g_opt.chunk.load());
  A.chunks_per_frame = (A.wins_per_frame + A.chunk - 1) / A.chunk;
  A.chunk_counter = d_work;
  int per_sm = (int)g_opt.gather_ctas_per_sm.load();
  if (per_sm <= 0) per_sm = gather_max_ctas_per_sm(A.tab.depth, A.rot_slot >= 0);
  const unsigned long long total_chunks = (unsigned long long)A.chunks_per_frame * A.nframes;
  long long grid = (long long)num_sms * per_sm;
  const long long warps_needed = (long long)total_chunks;       // one chunk keeps a warp busy
  grid = std::max(1ll, std::min(grid, (warps_needed + 7) / 8));
  timing_begin(T_GATHER, st);
  launch_scan_gather(A, (int)grid, st);
  timing_end(T_GATHER, st);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(PIGO_E_CUDA, "scan launch failed: %s", cudaGetErrorString(e));
  return PIGO_OK;
}

}  // namespace pigo
