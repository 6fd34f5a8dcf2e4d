// scan_tiled.cu -- the hot kernel: RunCascade's (scale,row,col) grid (core/pigo.go:226-249) with classifyRegion
// (core/pigo.go:113-147) for the small and medium scales, ONE WARP PER IMAGE TILE.
//
//  * Persistent CTA (one per SM), W independent warps.  The first KS trees of the cascade (codes, leaves,
//    threshold: one 516-byte record per tree, 129 words so consecutive trees are skewed by one bank) are staged
//    ONCE per CTA into shared memory by the TMA bulk engine (cp.async.bulk + mbarrier).
//  * Each warp owns a private pixel-tile buffer in shared memory.  A tile is an image region (core + halo) that
//    serves EVERY scale of a band (e.g. 20..39 px) at once, so the frame is read from L2/HBM once per band rather
//    than once per scale; it is filled with 16-byte cp.async vector loads (coalesced 128-bit rows).
//  * Windows whose centre lies in the tile's core are evaluated one-per-lane with LANE REFILL: every iteration
//    each live lane walks ONE tree (6 levels: 1 LDS.32 for the node's 4 codes, 2 LDS.U8 pixel gathers); lanes
//    whose window was rejected are re-armed with the next window of the tile via ballot+popc, so warps stay full
//    although ~60% of the windows die at tree 0.  NI independent item slots per lane give the ILP that hides the
//    shared-memory latency with only W warps per SM.
//  * Long-lived windows must not pin a tile: a window that reaches tree KS, and whatever is still alive when
//    fewer than `tail_min` items remain after the tile's windows ran out, is appended (wid, frame, tree, score)
//    to the global "deep" queue and finished by the resume kernel.  If the queue is full the lane simply keeps
//    going (cascade rows beyond KS are then read from global memory), so the queue is an optimisation only.
//
// Scores are float32 sums in tree order and the result is bit-identical to the reference.
#include "common.cuh"
#include "host.h"

namespace pigo {

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA bulk copy of the cascade prefix -------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

constexpr int kTreeRec = 516;  // bytes per tree record in the tiled table: 256 codes + 256 leaves + 4 threshold (depth 6)

// Walks tree record `tb` (shared-memory byte address) for the window whose centre pixel is at shared address pb.
__device__ __forceinline__ int walk_smem(uint32_t tb, uint32_t pb, int s, int pitch) {
  int idx = 1;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int cw = (int)lds_u32(tb + 4 * idx);
    // ((r*256 + code*s) >> 8) == r + ((code*s) >> 8)  (core/pigo.go:126-127)
    const int o1 = (((int)(int8_t)(cw) * s) >> 8) * pitch + (((int)(int8_t)(cw >> 8) * s) >> 8);
    const int o2 = (((int)(int8_t)(cw >> 16) * s) >> 8) * pitch + (((cw >> 24) * s) >> 8);
    const uint32_t p1 = lds_u8(pb + o1), p2 = lds_u8(pb + o2);
    idx = 2 * idx + (p1 <= p2 ? 1 : 0);  // core/pigo.go:129-135
  }
  return idx;
}
// Same walk with the node codes read from the reference-layout table in global memory (trees >= KS, queue full).
__device__ __forceinline__ int walk_smem_pixels_global_codes(const int8_t* __restrict__ tc, uint32_t pb, int s, int pitch) {
  int idx = 1;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int cw = __ldg(reinterpret_cast<const int*>(tc) + idx);
    const int o1 = (((int)(int8_t)(cw) * s) >> 8) * pitch + (((int)(int8_t)(cw >> 8) * s) >> 8);
    const int o2 = (((int)(int8_t)(cw >> 16) * s) >> 8) * pitch + (((cw >> 24) * s) >> 8);
    const uint32_t p1 = lds_u8(pb + o1), p2 = lds_u8(pb + o2);
    idx = 2 * idx + (p1 <= p2 ? 1 : 0);
  }
  return idx;
}

__device__ __forceinline__ int ceil_div_pos(int num, int den) { return num <= 0 ? 0 : (num + den - 1) / den; }

template <int NI, bool ALIGNED>
__global__ void __launch_bounds__(kTiledMaxThreads, 1) scan_tiled_kernel(const TiledArgs A) {
  extern __shared__ __align__(128) uint8_t smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bar = smem_base;                 // 8-byte mbarrier at offset 0
  const uint32_t casc = smem_base + 16;           // cascade prefix records
  const uint32_t casc_bytes = (uint32_t)A.ks * kTreeRec;
  const uint32_t tiles0 = (16 + casc_bytes + 127) & ~127u;
  const uint32_t my_tile = smem_base + tiles0 + (uint32_t)warp * A.tile_bytes;

  // ---- stage the cascade prefix with the TMA bulk engine (one elected thread issues, all threads wait)
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, casc_bytes);
    for (uint32_t off = 0; off < casc_bytes; off += 32768) {
      const uint32_t n = min(32768u, casc_bytes - off);
      tma_bulk_g2s(casc + off, A.tab_tiled + off, n, bar);
    }
  }
  __syncthreads();
  mbar_wait(bar, 0);

  const ScanArgs& S = A.scan;

  for (;;) {
    // ---- next tile for this warp
    unsigned long long tg = 0;
    if (lane == 0) tg = atomicAdd(S.chunk_counter, 1ull);
    tg = __shfl_sync(FULL, tg, 0);
    if (tg >= A.total_tiles) break;
    const int frame = (int)(tg / A.tiles_per_frame);
    int tf = (int)(tg % A.tiles_per_frame);
    int b = 0;
    while (b + 1 < A.nbands && tf >= A.band[b].ntiles) { tf -= A.band[b].ntiles; ++b; }
    const TileBand B = A.band[b];
    const int ty = tf / B.tiles_x, tx = tf - ty * B.tiles_x;
    const int cx0 = B.org_x + tx * B.core, cy0 = ty * B.core;          // core origin (window centres)
    const int gx0 = cx0 - B.halo_lo, gy0 = cy0 - B.halo_lo;            // tile origin in the frame (gx0 % 16 == 0)
    const int pitch = B.pitch;
    const uint8_t* fb = S.frames + (size_t)frame * S.frame_stride;

    // ---- per-scale window sub-grids of this tile: lane l describes band scale l
    int sc_i0 = 0, sc_j0 = 0, sc_nj = 0, sc_n = 0;
    ScaleEntry e{};
    if (lane < B.nscales) {
      e = S.plan[B.scale_lo + lane];
      const int i0 = ceil_div_pos(cy0 - e.off, e.step), i1 = min(e.nrows, ceil_div_pos(cy0 + B.core - e.off, e.step));
      const int j0 = ceil_div_pos(cx0 - e.off, e.step), j1 = min(e.ncols, ceil_div_pos(cx0 + B.core - e.off, e.step));
      sc_i0 = i0; sc_j0 = j0;
      sc_nj = max(0, j1 - j0);
      sc_n = max(0, i1 - i0) * sc_nj;
    }
    if (!__any_sync(FULL, sc_n > 0)) continue;  // no window centre falls into this core (frame border)

    // ---- fill the tile: rows [gy0, gy0+rows_t) x bytes [gx0, gx0+pitch) clipped to the frame
    __syncwarp();
    if (ALIGNED) {
      const int cpr = pitch >> 4;
      const int nchunks = B.rows_t * cpr;
      for (int q = lane; q < nchunks; q += 32) {
        const int row = q / cpr, cxk = q - row * cpr;
        const int y = gy0 + row, x = gx0 + 16 * cxk;
        if (y >= 0 && y < S.rows && x >= 0 && x < S.dim) cp_async16(my_tile + row * pitch + 16 * cxk, fb + (size_t)y * S.dim + x);
      }
      cp_async_wait_all();
    } else {
      const int nbytes = B.rows_t * pitch;
      for (int q = lane; q < nbytes; q += 32) {
        const int row = q / pitch, xx = q - row * pitch;
        const int y = gy0 + row, x = gx0 + xx;
        if (y >= 0 && y < S.rows && x >= 0 && x < S.dim) {
          const uint32_t v = __ldg(fb + (size_t)y * S.dim + x);
          asm volatile("st.shared.u8 [%0], %1;" ::"r"(my_tile + q), "r"(v) : "memory");
        }
      }
    }
    __syncwarp();

    // uniform cursor over the tile's window list (scale-major)
    int cur_si = -1, cur_k = 0, cur_n = 0;
    int u_s = 0, u_step = 0, u_off = 0, u_i0 = 0, u_j0 = 0, u_nj = 1, u_ncols = 0;
    uint32_t u_wbase = 0, u_magic = 0;
    bool exhausted = false;

    bool alive[NI];
    uint32_t pb[NI], tb[NI], wid[NI];
    int sv[NI], tv[NI];
    float acc[NI];
#pragma unroll
    for (int u = 0; u < NI; ++u) { alive[u] = false; pb[u] = my_tile; tb[u] = casc; wid[u] = 0; sv[u] = 0; tv[u] = 0; acc[u] = 0.f; }

    for (;;) {
      // ---- refill dead slots from the tile's window list
      bool any_alive = false;
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        unsigned need = __ballot_sync(FULL, !alive[u]);
        while (need && !exhausted) {
          if (cur_k == cur_n) {
            // advance to the next scale of the band that has windows in this tile
            int nsi = cur_si + 1;
            int n = 0;
            while (nsi < B.nscales && (n = __shfl_sync(FULL, sc_n, nsi)) == 0) ++nsi;
            if (nsi >= B.nscales) { exhausted = true; break; }
            cur_si = nsi; cur_k = 0; cur_n = n;
            u_s = __shfl_sync(FULL, e.s, nsi); u_step = __shfl_sync(FULL, e.step, nsi); u_off = __shfl_sync(FULL, e.off, nsi);
            u_i0 = __shfl_sync(FULL, sc_i0, nsi); u_j0 = __shfl_sync(FULL, sc_j0, nsi); u_nj = __shfl_sync(FULL, sc_nj, nsi);
            u_ncols = __shfl_sync(FULL, e.ncols, nsi); u_wbase = __shfl_sync(FULL, e.wbase, nsi);
            u_magic = (uint32_t)((0x100000000ull + (unsigned)u_nj - 1) / (unsigned)u_nj);  // ceil(2^32 / nj)
            continue;
          }
          const int avail = cur_n - cur_k;
          const int rank = __popc(need & lanemask_lt());
          if (!alive[u] && rank < avail) {
            const uint32_t k = (uint32_t)(cur_k + rank);
            const uint32_t i = u_nj == 1 ? k : __umulhi(k, u_magic);   // k / nj, exact for k*nj < 2^32
            const uint32_t j = k - i * (uint32_t)u_nj;
            const int gi = u_i0 + (int)i, gj = u_j0 + (int)j;
            const int r = u_off + gi * u_step, c = u_off + gj * u_step;
            pb[u] = my_tile + (uint32_t)((r - gy0) * pitch + (c - gx0));
            wid[u] = u_wbase + (uint32_t)gi * (uint32_t)u_ncols + (uint32_t)gj;
            sv[u] = u_s; tv[u] = 0; tb[u] = casc; acc[u] = 0.f;
            alive[u] = true;
          }
          cur_k += min(__popc(need), avail);
          need = __ballot_sync(FULL, !alive[u]);
        }
        // ---- tail policy: once the tile is drained, a thin slot group is handed to the deep queue
        const unsigned live = ~need;
        if (exhausted && live != 0u && __popc(live) < A.tail_min) {
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(S.deep_count, (unsigned)__popc(live));
          base = __shfl_sync(FULL, base, 0);
          const unsigned pos = base + __popc(live & lanemask_lt());
          if (alive[u] && pos < S.deep_cap) {
            S.deep[pos] = DeepItem{wid[u], frame, tv[u], acc[u]};
            alive[u] = false;
          }
        }
        any_alive |= __any_sync(FULL, alive[u]);
      }
      if (!any_alive) break;

      // ---- one tree per live item (dead slots walk a harmless dummy: tree 0 at the tile origin with s = 0)
      int idx[NI];
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        if (!alive[u]) { tv[u] = 0; tb[u] = casc; sv[u] = 0; }
        if (tv[u] < A.ks) {
          idx[u] = walk_smem(tb[u], pb[u], sv[u], pitch);
        } else {
          idx[u] = walk_smem_pixels_global_codes(S.tab.codes + (size_t)tv[u] * 256, pb[u], sv[u], pitch);
        }
      }
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        float pred, thr;
        if (tv[u] < A.ks) {
          pred = lds_f32(tb[u] + 256 + 4 * (idx[u] - 64));
          thr = lds_f32(tb[u] + 512);
        } else {
          pred = __ldg(S.tab.preds + (size_t)tv[u] * 64 + idx[u] - 64);
          thr = __ldg(S.tab.thresh + tv[u]);
        }
        if (alive[u]) {
          acc[u] += pred;                                   // core/pigo.go:137 (float32, tree order)
          if (acc[u] <= thr) {                              // :139-141
            alive[u] = false;
          } else {
            ++tv[u];
            tb[u] += kTreeRec;
            if (tv[u] == S.tab.ntrees) {
              const float q = acc[u] - thr;                 // :144
              if (q > 0.0f) {                               // :246
                const int pos = atomicAdd(S.raw_count + frame, 1);
                if (pos < S.cap) S.raw[(size_t)frame * S.cap + pos] = RawDet{wid[u], q};
              }
              alive[u] = false;
            } else if (tv[u] == A.ks) {
              const unsigned pos = atomicAdd(S.deep_count, 1u);
              if (pos < S.deep_cap) {
                S.deep[pos] = DeepItem{wid[u], frame, tv[u], acc[u]};
                alive[u] = false;
              }  // else: queue full -> keep walking with the global cascade rows
            }
          }
        }
      }
    }
  }
}

template <int NI>
static void launch_tiled_ni(const TiledArgs& A, int grid, int threads, size_t smem, bool aligned, cudaStream_t st) {
  if (aligned) {
    cudaFuncSetAttribute(scan_tiled_kernel<NI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_tiled_kernel<NI, true><<<grid, threads, smem, st>>>(A);
  } else {
    cudaFuncSetAttribute(scan_tiled_kernel<NI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_tiled_kernel<NI, false><<<grid, threads, smem, st>>>(A);
  }
}

void launch_scan_tiled(const TiledArgs& A, int grid, int threads, size_t smem, int ni, bool aligned, cudaStream_t st) {
  switch (ni) {
    case 1: launch_tiled_ni<1>(A, grid, threads, smem, aligned, st); break;
    case 2: launch_tiled_ni<2>(A, grid, threads, smem, aligned, st); break;
    case 3: launch_tiled_ni<3>(A, grid, threads, smem, aligned, st); break;
    default: launch_tiled_ni<4>(A, grid, threads, smem, aligned, st); break;
  }
}

}  // namespace pigo
