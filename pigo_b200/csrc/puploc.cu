// puploc.cu -- PuplocCascade.RunDetector (core/puploc.go:239-277) for a batch of seeds on one image:
// one CTA per seed, one warp per perturbation, one lane per tree of a stage, then the three independent 63-slot sorts
// and the "median" pick of the reference.  classifyRegion / classifyRotatedRegion (core/puploc.go:106-217) walk
// stages x trees x depth pixel-pair tests with global-memory gathers; tables stay L2/L1-resident.
//
// float32 arithmetic uses explicit round-to-nearest intrinsics (and the library is built with
// -fmad=false): Go on amd64 never fuses r += dr*s (core/puploc.go:149-151).
#include "common.cuh"
#include "host.h"

namespace pigo {

__device__ __constant__ float c_qcosf[33] = {256, 251, 236, 212, 181, 142, 97, 49, 0, -49, -97, -142, -181, -212, -236, -251, -256,
                                             -251, -236, -212, -181, -142, -97, -49, 0, 49, 97, 142, 181, 212, 236, 251, 256};
__device__ __constant__ float c_qsinf[33] = {0, 49, 97, 142, 181, 212, 236, 251, 256, 251, 236, 212, 181, 142, 97, 49, 0,
                                             -49, -97, -142, -181, -212, -236, -251, -256, -251, -236, -212, -181, -142, -97, -49, 0};

// int8 negation wraps in Go: -int8(-128) == -128 (core/puploc.go:124-125, :181-182)
__device__ __forceinline__ int neg_i8(int v) { return (int)(int8_t)(-v); }

__device__ __forceinline__ uint32_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 40);  // 24 random bits
}

// One CTA (32 warps) per seed; one WARP per perturbation (two rounds cover the 63 slots) and one LANE per tree of the
// current stage: the trees of a stage are independent of each other (they all start from the stage's (r, c, s)), only
// the float32 sums dr, dc are order-sensitive -- lanes walk the trees in parallel, then the sums are formed in tree order
// with shuffles, exactly like the reference's sequential loop (core/puploc.go:112-151).  That turns the reference's
// stages*trees*depth = 1000-deep dependent gather chain into stages*depth = 50; the two children of a node are fetched
// together with the node's pixel pair (child-pair prefetch).
__global__ void __launch_bounds__(1024, 1) puploc_kernel(PuplocTables T, const pigo_point* __restrict__ seeds, int nseeds,
                                                         const float* __restrict__ randoms, uint64_t rng_seed,
                                                         const uint8_t* __restrict__ frames, const int32_t* __restrict__ seed_frame,
                                                         size_t frame_stride, int nrows, int ncols, int dim,
                                                         int rot_slot, const uint8_t* __restrict__ flipv_arr,
                                                         pigo_point* __restrict__ out) {
  const unsigned FULL = 0xffffffffu;
  const int sidx = blockIdx.x;
  const uint8_t* __restrict__ pixels = frames + (seed_frame ? (size_t)seed_frame[sidx] * frame_stride : 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float rows_[64], cols_[64], scal_[64];
  const pigo_point seed = seeds[sidx];
  const int P = seed.perturbs;
  const bool flip = flipv_arr ? (flipv_arr[sidx] != 0) : false;
  const int L = T.leaves;
  const int tree_codes = 4 * L - 4;

  for (int i = warp; i < 63; i += 32) {       // warp-uniform: perturbation i (pool slot i)
    float r = 0.f, c = 0.f, s = 0.f;
    if (i < P) {
      float u0, u1, u2;
      if (randoms) {
        const float* rr = randoms + ((size_t)sidx * 63 + i) * 3;
        u0 = rr[0]; u1 = rr[1]; u2 = rr[2];
      } else {
        const uint64_t key = rng_seed * 0xD1342543DE82EF95ull + (uint64_t)sidx * 64 + i;
        u0 = mix64(key * 3 + 0) * (1.0f / 16777216.0f);
        u1 = mix64(key * 3 + 1) * (1.0f / 16777216.0f);
        u2 = mix64(key * 3 + 2) * (1.0f / 16777216.0f);
      }
      const float t1 = __fmul_rn(seed.scale, 0.15f);
      r = __fadd_rn((float)seed.row, __fmul_rn(t1, __fsub_rn(0.5f, u0)));   // core/puploc.go:248
      c = __fadd_rn((float)seed.col, __fmul_rn(t1, __fsub_rn(0.5f, u1)));   // :249
      s = __fmul_rn(seed.scale, __fadd_rn(0.925f, __fmul_rn(0.15f, u2)));   // :250
      int iqs = 0, iqc = 0;
      if (rot_slot >= 0) {
        iqs = (int)__fmul_rn(s, c_qsinf[rot_slot]);  // int(qsin), :166, :188 (from the INITIAL s)
        iqc = (int)__fmul_rn(s, c_qcosf[rot_slot]);
      }
      for (int st = 0; st < T.stages; ++st) {
        float dr = 0.f, dc = 0.f;
        const long long ir = (long long)r, ic = (long long)c;  // int(r): truncation toward zero
        const long long rs = llround((double)s);               // int(math.Round(float64(s)))
        for (int j0 = 0; j0 < T.trees; j0 += 32) {             // 32 trees of the stage at a time, one per lane
          const int j = min(j0 + lane, T.trees - 1);           // surplus lanes redo the last tree (ignored below)
          const size_t tg = (size_t)st * T.trees + j;
          const int* tc = reinterpret_cast<const int*>(T.codes + tg * tree_codes);   // 4-byte aligned: tree_codes % 4 == 0
          const float2* tp = reinterpret_cast<const float2*>(T.preds + tg * 2 * L);
          int idx = 0;
          int cw = __ldg(tc);
          for (int k = 0; k < T.depth; ++k) {
            // children of node idx are nodes 2idx+1, 2idx+2: fetch both while this node's pixels are in flight
            int kl = 0, kr = 0;
            if (k + 1 < T.depth) { kl = __ldg(tc + 2 * idx + 1); kr = __ldg(tc + 2 * idx + 2); }
            const int k0 = (int8_t)(cw), k2 = (int8_t)(cw >> 16);
            const int k1 = flip ? neg_i8((int8_t)(cw >> 8)) : (int)(int8_t)(cw >> 8);
            const int k3 = flip ? neg_i8(cw >> 24) : (cw >> 24);
            long long r1, c1, r2, c2;
            int bit;
            if (rot_slot < 0) {
              r1 = min((long long)nrows - 1, max(0ll, (256 * ir + k0 * rs) >> 8));   // :118
              r2 = min((long long)nrows - 1, max(0ll, (256 * ir + k2 * rs) >> 8));   // :119
              c1 = min((long long)ncols - 1, max(0ll, (256 * ic + k1 * rs) >> 8));   // :124/:127
              c2 = min((long long)ncols - 1, max(0ll, (256 * ic + k3 * rs) >> 8));
              bit = __ldg(pixels + r1 * dim + c1) > __ldg(pixels + r2 * dim + c2) ? 1 : 0;   // :130-136
            } else {
              r1 = min((long long)nrows - 1, max(0ll, 65536 * ir + (long long)iqc * k0 - (long long)iqs * k1) >> 16);  // :188
              c1 = min((long long)ncols - 1, max(0ll, 65536 * ic + (long long)iqs * k0 + (long long)iqc * k1) >> 16);  // :189
              r2 = min((long long)nrows - 1, max(0ll, 65536 * ir + (long long)iqc * k2 - (long long)iqs * k3) >> 16);  // :190
              c2 = min((long long)ncols - 1, max(0ll, 65536 * ic + (long long)iqs * k2 + (long long)iqc * k3) >> 16);  // :191
              bit = __ldg(pixels + r1 * dim + c1) <= __ldg(pixels + r2 * dim + c2) ? 1 : 0;  // :193-199
            }
            cw = bit ? kr : kl;
            idx = 2 * idx + 1 + bit;
          }
          const float2 leaf = __ldg(tp + (idx - (L - 1)));
          const int nv = min(32, T.trees - j0);
          for (int q = 0; q < nv; ++q) {                       // the reference's tree-ordered float32 sums, :140-145
            const float pr = __shfl_sync(FULL, leaf.x, q), pc = __shfl_sync(FULL, leaf.y, q);
            dr = __fadd_rn(dr, pr);
            dc = __fadd_rn(dc, flip ? -pc : pc);
          }
        }
        r = __fadd_rn(r, __fmul_rn(dr, s));   // :149
        c = __fadd_rn(c, __fmul_rn(dc, s));   // :150
        s = __fmul_rn(s, T.scales);           // :151
      }
    }
    // pool slots >= Perturbs stay 0 (fresh pool object, :228-236); all 63 slots are sorted (:267-269)
    if (lane == 0) { rows_[i] = r; cols_[i] = c; scal_[i] = s; }
  }
  __syncthreads();
  const int i = threadIdx.x;
  const int mid = (int)llround((double)P / 2);  // int(math.Round(float64(Perturbs)/2)), :273
  if (i < 63) {
    // rank sort of 63 values (ties: index order); the thread whose rank == mid publishes it
    const float vr = rows_[i], vc = cols_[i], vs = scal_[i];
    int kr = 0, kc = 0, ks = 0;
    for (int j = 0; j < 63; ++j) {
      kr += (rows_[j] < vr || (rows_[j] == vr && j < i)) ? 1 : 0;
      kc += (cols_[j] < vc || (cols_[j] == vc && j < i)) ? 1 : 0;
      ks += (scal_[j] < vs || (scal_[j] == vs && j < i)) ? 1 : 0;
    }
    if (kr == mid) out[sidx].row = (int)vr;      // int() truncation, :273
    if (kc == mid) out[sidx].col = (int)vc;
    if (ks == mid) out[sidx].scale = vs;
    if (i == 0) out[sidx].perturbs = 0;          // the returned Puploc leaves Perturbs unset (:272-276)
  }
}

void launch_puploc(const PuplocTables& T, const pigo_point* seeds, int nseeds, const float* randoms, uint64_t rng_seed,
                   const uint8_t* frames, const int32_t* seed_frame, size_t frame_stride, int rows, int cols, int dim, int rot_slot,
                   const uint8_t* flipv, pigo_point* out, cudaStream_t st) {
  puploc_kernel<<<nseeds, 1024, 0, st>>>(T, seeds, nseeds, randoms, rng_seed, frames, seed_frame, frame_stride, rows, cols, dim, rot_slot,
                                         flipv, out);
}

}  // namespace pigo
