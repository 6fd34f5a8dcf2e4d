// deep.cu -- finishes the windows that survived the KS shared-memory-resident trees (queue Q2), ONE WARP PER WINDOW,
// ONE TREE PER LANE: lanes walk 32 consecutive trees of the same window in parallel (tree walks are independent of
// each other -- only the early-exit test couples them), then the warp replays the reference's float32 accumulation
// and threshold tests in tree order (core/pigo.go:137-141) with shuffles.  A full survivor of the 468-tree cascade
// costs ~13 steps instead of a ~400-tree serial chain of dependent L2 round trips; trees evaluated past the
// rejecting one are wasted work, acceptable because windows that reach tree KS usually live long.
// Bit-exactness: each lane produces the same leaf the serial walk would, and the sum is formed in the same order.
#include "common.cuh"
#include "host.h"

namespace pigo {

// GROUP = lanes (= trees per step) per window: 32 -> one window per warp; 16 / 8 -> two / four windows per warp, each
// an independent lane group (all sync ops use the half's mask), which halves the speculation past the rejecting
// tree and doubles the windows in flight per warp.
template <int GROUP>
__global__ void __launch_bounds__(256) deep_kernel(const ScanArgs A, unsigned long long* counter) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (GROUP - 1);
  const unsigned gmask = GROUP == 32 ? 0xffffffffu : (GROUP == 16 ? (0xffffu << (lane & 16)) : (0xffu << (lane & 24)));
  const int leader = lane & ~(GROUP - 1);
  const FaceTables T = A.tab;
  const uint32_t qn = min(*A.long_count, A.long_cap);
  for (;;) {
    unsigned long long g = 0;
    if (sub == 0) g = atomicAdd(counter, 1ull);
    g = __shfl_sync(gmask, g, leader);
    if (g >= qn) break;
    const DeepItem it = A.longq[g];
    const int si = find_scale(A.plan, A.nscales, it.wid);
    const ScaleEntry e = A.plan[si];
    const uint32_t local = it.wid - e.wbase;
    const uint32_t ri = local / (uint32_t)e.ncols, ci = local - ri * (uint32_t)e.ncols;
    const int s = e.s;
    const uint8_t* pc = A.frames + (size_t)it.frame * A.frame_stride + (size_t)(e.off + (int)ri * e.step) * A.dim + (e.off + (int)ci * e.step);
    int t0 = it.tree;
    float acc = it.acc;
    bool rejected = false;
    float thr_prev = 0.f;
    while (t0 < T.ntrees && !rejected) {
      const int t = min(t0 + sub, T.ntrees - 1);      // lanes past the last tree redo it harmlessly
      // child-pair prefetch: both children of node idx (codes: bytes 8*idx.., leaves: floats 2*idx-64..) are fetched with
      // one 64-bit load that is in flight together with the two pixel gathers, so a level costs ONE dependent L2 round
      // trip instead of two
      const int2* tc2 = reinterpret_cast<const int2*>(T.codes + (size_t)t * 256);
      const int2* tp2 = reinterpret_cast<const int2*>(T.preds + (size_t)t * 64);
      const float thr = __ldg(T.thresh + t);
      int idx = 1;
      int cw = __ldg(reinterpret_cast<const int*>(tc2) + 1);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int2 kids = j < 5 ? __ldg(tc2 + idx) : __ldg(tp2 + (idx - 32));
        const int o1 = (((int)(int8_t)(cw) * s) >> 8) * A.dim + (((int)(int8_t)(cw >> 8) * s) >> 8);
        const int o2 = (((int)(int8_t)(cw >> 16) * s) >> 8) * A.dim + (((cw >> 24) * s) >> 8);
        const unsigned p1 = __ldg(pc + o1), p2 = __ldg(pc + o2);
        const bool right = p1 <= p2;                  // core/pigo.go:129-135
        cw = right ? kids.y : kids.x;
        idx = 2 * idx + (right ? 1 : 0);
      }
      const float pred = __int_as_float(cw);
      const int nvalid = min(GROUP, T.ntrees - t0);
      for (int j = 0; j < nvalid; ++j) {              // the reference's sequential accumulation, :137-141
        acc += __shfl_sync(gmask, pred, leader + j);
        thr_prev = __shfl_sync(gmask, thr, leader + j);
        if (acc <= thr_prev) { rejected = true; break; }
      }
      t0 += GROUP;
    }
    if (!rejected && sub == 0) {
      const float q = acc - thr_prev;                 // :144 (thr_prev == threshold of the last tree)
      if (q > 0.0f) {                                 // :246
        const int pos = atomicAdd(A.raw_count + it.frame, 1);
        if (pos < A.cap) A.raw[(size_t)it.frame * A.cap + pos] = RawDet{it.wid, q};
      }
    }
  }
}

void launch_deep(const ScanArgs& A, unsigned long long* counter, int grid, int group, cudaStream_t st) {
  if (group == 8) deep_kernel<8><<<grid, 256, 0, st>>>(A, counter);
  else if (group == 16) deep_kernel<16><<<grid, 256, 0, st>>>(A, counter);
  else deep_kernel<32><<<grid, 256, 0, st>>>(A, counter);
}

}  // namespace pigo
