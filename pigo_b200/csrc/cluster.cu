// cluster.cu -- ClusterDetections (core/pigo.go:262-308) on the device, one CTA per frame.
//
//  1. in-place sort of the frame's detections by Q ascending (:264-266).  Go's sort.Slice is unstable and
//     leaves tie order unspecified; we sort stably (rank = #{q_j < q_i} + #{q_j == q_i, j < i}).
//  2. greedy seed selection (:282-301): detection i is a seed iff no earlier seed overlapped it with
//     IoU > threshold.  Sequential over seeds, parallel over j; one barrier per SEED, not per detection.
//  3. per-seed accumulation, one thread per seed, j ascending so that the float32 q-sum (:298) has the
//     reference's addition order; integer sums and the truncating division (:303) as in Go.
// IoU is float64 with the reference's operation order (:268-278); the library is built with -fmad=false.
#include "common.cuh"
#include "host.h"

namespace pigo {

__device__ __forceinline__ double calc_iou(const pigo_det& a, const pigo_det& b) {
  const double r1 = a.row, c1 = a.col, s1 = a.scale;
  const double r2 = b.row, c2 = b.col, s2 = b.scale;
  const double over_row = fmax(0.0, fmin(r1 + s1 / 2, r2 + s2 / 2) - fmax(r1 - s1 / 2, r2 - s2 / 2));
  const double over_col = fmax(0.0, fmin(c1 + s1 / 2, c2 + s2 / 2) - fmax(c1 - s1 / 2, c2 - s2 / 2));
  return __ddiv_rn(__dmul_rn(over_row, over_col),
                   __dsub_rn(__dadd_rn(__dmul_rn(s1, s1), __dmul_rn(s2, s2)), __dmul_rn(over_row, over_col)));
}

// scratch per frame: tmp[cap] (pigo_det), flags[cap] (u8: assigned), seeds[cap] (i32)
__global__ void __launch_bounds__(256) cluster_kernel(pigo_det* __restrict__ dets, const int32_t* __restrict__ n_in, int cap,
                                                      double thr, pigo_det* __restrict__ tmp, uint8_t* flags,
                                                      int32_t* __restrict__ seeds, pigo_det* __restrict__ out, int out_cap,
                                                      int32_t* __restrict__ n_out) {
  const int frame = blockIdx.x;
  const int n = max(0, min(n_in[frame], cap));
  pigo_det* d = dets + (size_t)frame * cap;
  pigo_det* tp = tmp + (size_t)frame * cap;
  uint8_t* asg = flags + (size_t)frame * cap;
  int32_t* sd = seeds + (size_t)frame * cap;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int s_nseeds;

  // ---- 1. stable rank sort by q ascending
  for (int i = tid; i < n; i += nt) {
    const float qi = d[i].q;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float qj = d[j].q;
      rank += (qj < qi || (qj == qi && j < i)) ? 1 : 0;
    }
    tp[rank] = d[i];
    asg[i] = 0;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) d[i] = tp[i];
  if (tid == 0) { s_nseeds = 0; }
  __syncthreads();

  // ---- 2. seeds
  for (int i = 0; i < n; ++i) {
    // Uniform branch: asg[i] is written only inside the marking loops of EARLIER seed iterations, each of which ends in a
    // barrier.  The seed's own flag is deliberately not set below (j != i): the reference sets assignments[i] there
    // (IoU(i,i) = 1 > thr, core/pigo.go:290-296) but never reads it again, and setting it here would let a lagging warp
    // of this same iteration read 1 and skip the barrier the others are waiting at.
    if (*(volatile uint8_t*)(asg + i)) continue;
    const pigo_det di = d[i];
    for (int j = tid; j < n; j += nt)
      if (j != i && calc_iou(di, d[j]) > thr) asg[j] = 1;
    if (tid == 0) sd[s_nseeds++] = i;
    __syncthreads();
  }
  __syncthreads();
  const int nseeds = s_nseeds;

  // ---- 3. accumulate each seed's cluster (j ascending); tp[] is reused for the per-seed results
  for (int k = tid; k < nseeds; k += nt) {
    const pigo_det di = d[sd[k]];
    long long r = 0, c = 0, s = 0, cnt = 0;
    float q = 0.f;
    for (int j = 0; j < n; ++j) {
      const pigo_det dj = d[j];
      if (calc_iou(di, dj) > thr) {
        r += dj.row; c += dj.col; s += dj.scale;
        q = __fadd_rn(q, dj.q);
        ++cnt;
      }
    }
    pigo_det o;
    if (cnt > 0) { o.row = (int)(r / cnt); o.col = (int)(c / cnt); o.scale = (int)(s / cnt); o.q = q; }
    else { o.row = o.col = 0; o.scale = -1; o.q = 0.f; }  // scale -1 marks "n == 0" (:302), dropped below
    tp[k] = o;
  }
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    pigo_det* of = out + (size_t)frame * out_cap;
    for (int k = 0; k < nseeds; ++k) {
      if (tp[k].scale == -1 && tp[k].row == 0 && tp[k].col == 0) {
        // distinguish the marker from a genuine cluster: genuine clusters have cnt > 0 and scale >= 0 whenever
        // input scales are >= 0 (always true for RunCascade output)
        continue;
      }
      if (m < out_cap) of[m] = tp[k];
      ++m;
    }
    n_out[frame] = m;
  }
}

void launch_cluster(pigo_det* dets, const int32_t* n_in, int cap, double thr, pigo_det* tmp, uint8_t* flags, int32_t* seeds,
                    pigo_det* out, int out_cap, int32_t* n_out, int nframes, cudaStream_t st) {
  cluster_kernel<<<nframes, 256, 0, st>>>(dets, n_in, cap, thr, tmp, flags, seeds, out, out_cap, n_out);
}

}  // namespace pigo
