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

// GROUP = lanes (= trees per step) per window: 32 -> one window per warp; 16 -> two windows per warp, each half-warp
// an independent 16-lane group (all sync ops use the half's mask), which halves the speculation past the rejecting
// tree and doubles the windows in flight per warp.
template <int GROUP>
__global__ void __launch_bounds__(256) deep_kernel(const ScanArgs A, unsigned long long* counter) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (GROUP - 1);
  const unsigned gmask = GROUP == 32 ? 0xffffffffu : (0xffffu << (lane & 16));
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
      const int* tc = reinterpret_cast<const int*>(T.codes + (size_t)t * 256);
      int idx = 1;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int cw = __ldg(tc + idx);
        const int o1 = (((int)(int8_t)(cw) * s) >> 8) * A.dim + (((int)(int8_t)(cw >> 8) * s) >> 8);
        const int o2 = (((int)(int8_t)(cw >> 16) * s) >> 8) * A.dim + (((cw >> 24) * s) >> 8);
        const unsigned p1 = __ldg(pc + o1), p2 = __ldg(pc + o2);
        idx = 2 * idx + (p1 <= p2 ? 1 : 0);           // core/pigo.go:129-135
      }
      const float pred = __ldg(T.preds + (size_t)t * 64 + idx - 64);
      const float thr = __ldg(T.thresh + t);
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
  if (group == 16) deep_kernel<16><<<grid, 256, 0, st>>>(A, counter);
  else deep_kernel<32><<<grid, 256, 0, st>>>(A, counter);
}

// ---- v2: the cascade tail [kd, ntrees) resident in shared memory (one persistent CTA per SM) ----------------------
// Every item in Q2 sits at tree >= kd, so node codes, leaves and thresholds never leave the SM; only the 12 pixel
// gathers per tree go to L1/L2, and the 32 lanes of a warp sample the same s x s window, so they hit the same lines.
__device__ __forceinline__ void mbar_init_d(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__global__ void __launch_bounds__(1024, 1) deep_smem_kernel(const ScanArgs A, unsigned long long* counter, const uint8_t* tab_tiled,
                                                           int kd, uint32_t copy_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const FaceTables T = A.tab;
  const uint32_t qn = min(*A.long_count, A.long_cap);
  if (qn == 0) return;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bar = smem_base, casc = 128;
  if (threadIdx.x == 0) mbar_init_d(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(copy_bytes) : "memory");
    const uint8_t* src = tab_tiled + (size_t)kd * 516;   // kd*516 is a multiple of 16 when kd % 4 == 0 (host guarantees)
    for (uint32_t off = 0; off < copy_bytes; off += 32768) {
      const uint32_t n = min(32768u, copy_bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_base + casc + off),
                   "l"(src + off), "r"(n), "r"(bar)
                   : "memory");
    }
  }
  __syncthreads();
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITD_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra DONED_%=;\n\t"
      "bra WAITD_%=;\n\t"
      "DONED_%=:\n\t}" ::"r"(bar)
      : "memory");

  for (;;) {
    unsigned long long g = 0;
    if (lane == 0) g = atomicAdd(counter, 1ull);
    g = __shfl_sync(FULL, g, 0);
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
      const int t = min(t0 + lane, T.ntrees - 1);
      const uint32_t tb = casc + (uint32_t)(t - kd) * 516u;
      int idx = 1;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int cw = *reinterpret_cast<const int*>(smem + tb + 4 * idx);
        const int o1 = (((int)(int8_t)(cw) * s) >> 8) * A.dim + (((int)(int8_t)(cw >> 8) * s) >> 8);
        const int o2 = (((int)(int8_t)(cw >> 16) * s) >> 8) * A.dim + (((cw >> 24) * s) >> 8);
        const unsigned p1 = __ldg(pc + o1), p2 = __ldg(pc + o2);
        idx = 2 * idx + (p1 <= p2 ? 1 : 0);
      }
      const float pred = *reinterpret_cast<const float*>(smem + tb + 4 * idx);
      const float thr = *reinterpret_cast<const float*>(smem + tb + 512);
      const int nvalid = min(32, T.ntrees - t0);
      for (int j = 0; j < nvalid; ++j) {
        acc += __shfl_sync(FULL, pred, j);
        thr_prev = __shfl_sync(FULL, thr, j);
        if (acc <= thr_prev) { rejected = true; break; }
      }
      t0 += 32;
    }
    if (!rejected && lane == 0) {
      const float q = acc - thr_prev;
      if (q > 0.0f) {
        const int pos = atomicAdd(A.raw_count + it.frame, 1);
        if (pos < A.cap) A.raw[(size_t)it.frame * A.cap + pos] = RawDet{it.wid, q};
      }
    }
  }
}

void launch_deep_smem(const ScanArgs& A, unsigned long long* counter, const uint8_t* tab_tiled, int kd, int grid, cudaStream_t st) {
  const uint32_t copy_bytes = (uint32_t)((((size_t)(A.tab.ntrees - kd) * 516) + 15) & ~(size_t)15);
  const size_t smem = 128 + copy_bytes;
  cudaFuncSetAttribute(deep_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  deep_smem_kernel<<<grid, 1024, smem, st>>>(A, counter, tab_tiled, kd, copy_bytes);
}

}  // namespace pigo
