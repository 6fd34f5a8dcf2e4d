// scan_gather.cu -- "gather" scan kernel: the universal (any scale, rotated or not) implementation of
// the RunCascade grid (core/pigo.go:212-258) with classifyRegion / classifyRotatedRegion
// (core/pigo.go:113-191) evaluated by global-memory gathers.
//
// Execution model: persistent warps, one window per lane, LANE REFILL: every loop iteration each live
// lane walks ONE tree of its window; lanes whose window was rejected (out <= threshold, :139) or
// finished are re-armed in place with the next unprocessed window index (ballot + popc prefix), so
// warps stay full although ~60% of windows die at the first tree.  Windows are handed out in chunks
// of `chunk` consecutive in-frame indices through one global atomic per chunk.
//
// This is the universal path: every scale of the rotated scan (angle > 0), cascades whose tree depth is not 6, and
// scan_mode=1.  The unrotated depth-6 scan uses the fused kernel of scan_tiled.cu (+ gather-v2 and deep kernels).
#include <algorithm>

#include "common.cuh"
#include "host.h"

namespace pigo {

// Both walks use the child-pair prefetch: the two children of node idx are adjacent (codes at bytes 8*idx.., leaves at
// floats 2*idx-L..), so one 64-bit load fetches both while the pixel gathers of the current level are in flight and a
// select picks the child afterwards -- one dependent L2 round trip per level instead of two.  They return the LEAF VALUE.
template <int DEPTH>  // DEPTH = 0: runtime depth
__device__ __forceinline__ float walk_tree(const int8_t* __restrict__ tc, const float* __restrict__ tp, const uint8_t* __restrict__ pc,
                                           int s, int dim, int depth, int leaves) {
  const int D = DEPTH ? DEPTH : depth;
  const int2* tc2 = reinterpret_cast<const int2*>(tc);
  const int2* tp2 = reinterpret_cast<const int2*>(tp);
  int idx = 1;
  int cw = __ldg(reinterpret_cast<const int*>(tc) + 1);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const int2 kids = (j < D - 1) ? __ldg(tc2 + idx) : __ldg(tp2 + (idx - (leaves >> 1)));
    // ((r*256 + code*s) >> 8) == r + ((code*s) >> 8): r*256 is a multiple of 256 (core/pigo.go:126-127)
    const int o1 = (((int)(int8_t)(cw) * s) >> 8) * dim + (((int)(int8_t)(cw >> 8) * s) >> 8);
    const int o2 = (((int)(int8_t)(cw >> 16) * s) >> 8) * dim + (((cw >> 24) * s) >> 8);
    const unsigned p1 = __ldg(pc + o1), p2 = __ldg(pc + o2);
    const bool right = p1 <= p2;           // core/pigo.go:129-135
    cw = right ? kids.y : kids.x;
    idx = 2 * idx + (right ? 1 : 0);
  }
  return __int_as_float(cw);
}

// classifyRotatedRegion node walk (core/pigo.go:164-180).  NB: both coordinates are clamped with
// nrows-1 (:167-171) -- reproduced on purpose.  I = long long mirrors Go's 64-bit int; I = int is chosen by the host
// when 65536*max(rows, cols) + 2*256*128*max_scale provably fits 31 bits (every realistic frame), which halves the
// instruction count of the coordinate arithmetic.
template <int DEPTH, typename I>
__device__ __forceinline__ float walk_tree_rot(const int8_t* __restrict__ tc, const float* __restrict__ tp, const uint8_t* __restrict__ frame,
                                               int r, int c, I qsin, I qcos, int nrows, int dim, int depth, int leaves) {
  const int D = DEPTH ? DEPTH : depth;
  const I lim = nrows - 1;
  const I r16 = (I)65536 * r, c16 = (I)65536 * c;
  const int2* tc2 = reinterpret_cast<const int2*>(tc);
  const int2* tp2 = reinterpret_cast<const int2*>(tp);
  int idx = 1;
  int cw = __ldg(reinterpret_cast<const int*>(tc) + 1);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const int2 kids = (j < D - 1) ? __ldg(tc2 + idx) : __ldg(tp2 + (idx - (leaves >> 1)));
    const I k0 = (int8_t)(cw), k1 = (int8_t)(cw >> 8), k2 = (int8_t)(cw >> 16), k3 = (cw >> 24);
    I r1 = min(lim, max((I)0, r16 + qcos * k0 - qsin * k1) >> 16);
    I c1 = min(lim, max((I)0, c16 + qsin * k0 + qcos * k1) >> 16);
    I r2 = min(lim, max((I)0, r16 + qcos * k2 - qsin * k3) >> 16);
    I c2 = min(lim, max((I)0, c16 + qsin * k2 + qcos * k3) >> 16);
    r1 = r1 < 0 ? -r1 : r1; c1 = c1 < 0 ? -c1 : c1; r2 = r2 < 0 ? -r2 : r2; c2 = c2 < 0 ? -c2 : c2;  // abs(), :167
    const unsigned p1 = __ldg(frame + (size_t)r1 * dim + (size_t)c1), p2 = __ldg(frame + (size_t)r2 * dim + (size_t)c2);
    const bool right = p1 <= p2;
    cw = right ? kids.y : kids.x;
    idx = 2 * idx + (right ? 1 : 0);
  }
  return __int_as_float(cw);
}

// ROT: 0 = unrotated, 1 = rotated with 64-bit coordinates, 2 = rotated with 32-bit coordinates (host-proven safe)
template <int DEPTH, int ROT>
__global__ void __launch_bounds__(256) scan_gather_kernel(const ScanArgs A) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const FaceTables T = A.tab;
  const int L = T.leaves;
  const int code_stride = 4 * L;

  // window range of the scales this launch covers
  const uint32_t w_lo = A.plan[A.scale_lo].wbase;
  const uint32_t w_hi = (A.scale_hi < A.nscales) ? A.plan[A.scale_hi].wbase : A.wins_per_frame;

  const unsigned long long total_chunks = (unsigned long long)A.chunks_per_frame * A.nframes;

  // per-lane item
  bool alive = false;
  const uint8_t* pc = nullptr;  // unrotated: centre pixel; rotated: frame base
  int s = 0, t = 0, frame = 0, r = 0, c = 0;
  uint32_t wid = 0;
  float acc = 0.f;
  // per-warp chunk cursor (uniform)
  uint32_t cur = 0, end = 0;
  int cframe = 0;
  bool more = true;

  for (;;) {
    unsigned need = __ballot_sync(FULL, !alive);
    while (need && more) {
      if (cur == end) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(A.chunk_counter, 1ull);
        g = __shfl_sync(FULL, g, 0);
        if (g >= total_chunks) { more = false; break; }
        cframe = (int)(g / A.chunks_per_frame);
        const uint32_t k = (uint32_t)(g % A.chunks_per_frame);
        cur = w_lo + k * A.chunk;
        end = min(cur + A.chunk, w_hi);
        continue;
      }
      const uint32_t avail = end - cur;
      const uint32_t rank = __popc(need & lanemask_lt());
      if (!alive && rank < avail) {
        wid = cur + rank;
        const int si = find_scale(A.plan, A.nscales, wid);
        const ScaleEntry e = A.plan[si];
        const uint32_t local = wid - e.wbase;
        const uint32_t ri = local / (uint32_t)e.ncols;
        const uint32_t ci = local - ri * (uint32_t)e.ncols;
        r = e.off + (int)ri * e.step;
        c = e.off + (int)ci * e.step;
        s = e.s;
        frame = cframe;
        const uint8_t* fb = A.frames + (size_t)cframe * A.frame_stride;
        pc = ROT ? fb : fb + (size_t)r * A.dim + c;
        t = 0;
        acc = 0.f;
        alive = true;
      }
      cur += min((uint32_t)__popc(need), avail);
      need = __ballot_sync(FULL, !alive);
    }
    if (!__any_sync(FULL, alive)) break;
    if (alive) {
      const int8_t* tc = T.codes + (size_t)t * code_stride;
      const float* tp = T.preds + (size_t)t * L;
      float pred;
      if (ROT == 1) {
        const long long qsin = (long long)s * c_qsin[A.rot_slot];  // core/pigo.go:159
        const long long qcos = (long long)s * c_qcos[A.rot_slot];  // :160
        pred = walk_tree_rot<DEPTH, long long>(tc, tp, pc, r, c, qsin, qcos, A.rows, A.dim, T.depth, L);
      } else if (ROT == 2) {
        pred = walk_tree_rot<DEPTH, int>(tc, tp, pc, r, c, s * c_qsin[A.rot_slot], s * c_qcos[A.rot_slot], A.rows, A.dim, T.depth, L);
      } else {
        pred = walk_tree<DEPTH>(tc, tp, pc, s, A.dim, T.depth, L);
      }
      acc += pred;                                       // core/pigo.go:137 (float32, tree order)
      const float thr = __ldg(T.thresh + t);
      if (acc <= thr) {                                  // :139-141
        alive = false;
      } else if (++t == T.ntrees) {
        const float q = acc - thr;                       // :144 ; q > 0 always holds here, kept for :246
        if (q > 0.0f) {
          const int pos = atomicAdd(A.raw_count + frame, 1);
          if (pos < A.cap) A.raw[(size_t)frame * A.cap + pos] = RawDet{wid, q};
        }
        alive = false;
      }
    }
  }
}

static bool rot32_safe(const ScanArgs& A, int max_scale) {
  // |65536*coord| + |qcos*k| + |qsin*k| <= 65536*max(rows, cols) + 2 * (256*max_scale) * 128  must stay below 2^31
  const long long bound = 65536ll * std::max(A.rows, A.cols) + 2ll * 256 * 128 * (long long)max_scale;
  return bound < 0x7fffffffll;
}

void launch_scan_gather(const ScanArgs& A, int grid, int max_scale, cudaStream_t st) {
  const int rot = A.rot_slot < 0 ? 0 : (rot32_safe(A, max_scale) ? 2 : 1);
  if (A.tab.depth == 6) {
    if (rot == 2) scan_gather_kernel<6, 2><<<grid, 256, 0, st>>>(A);
    else if (rot == 1) scan_gather_kernel<6, 1><<<grid, 256, 0, st>>>(A);
    else scan_gather_kernel<6, 0><<<grid, 256, 0, st>>>(A);
  } else {
    if (rot == 2) scan_gather_kernel<0, 2><<<grid, 256, 0, st>>>(A);
    else if (rot == 1) scan_gather_kernel<0, 1><<<grid, 256, 0, st>>>(A);
    else scan_gather_kernel<0, 0><<<grid, 256, 0, st>>>(A);
  }
}

}  // namespace pigo

namespace pigo {
int gather_max_ctas_per_sm(int depth, bool rot) {
  int n = 0;
  const void* f;
  if (depth == 6) f = rot ? (const void*)scan_gather_kernel<6, 1> : (const void*)scan_gather_kernel<6, 0>;
  else f = rot ? (const void*)scan_gather_kernel<0, 1> : (const void*)scan_gather_kernel<0, 0>;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, f, 256, 0) != cudaSuccess || n < 1) { cudaGetLastError(); n = 4; }
  return n;
}
}  // namespace pigo
