// Halo-staged tcgen05 convolution for the full-resolution 16/32-channel 3x3(x3) layers (sm_100a, bf16).
//
// conv_tc.cu loads one shifted activation box per tap through TMA: 27 L2->smem transfers of the same
// voxels per output tile, which makes the 16/32-channel layers (80 % of the network's bytes) L2-bound.
// This kernel stages each input d-slice ONCE in shared memory and lets the 27 taps read shifted VIEWS of it:
//   * a CTA owns an (8 wide x 16 high) output column and marches along d; a ring of input slices
//     (18 x 10 voxels each, one-voxel halo, zero filled = conv padding) lives in smem, so each input voxel
//     is fetched ~1.4x instead of 27x;
//   * slices are stored channel-planar, [C/8 planes][18][10][8 ch], i.e. every voxel contributes one 16-byte
//     K-chunk per plane and 8 consecutive voxels along w form one UMMA core matrix (8 rows x 16 B).  In the
//     no-swizzle K-major canonical layout the 16 core-matrix rows of the M = 128 tile are then a constant
//     SBO = 160 B (one halo row) apart and the two K-chunks LBO = one plane apart, for EVERY tap: the tap only
//     moves the descriptor start address by (kh*10 + kw)*16 B inside the slice selected by kd;
//   * the packed weights of all taps stay resident in smem for the CTA's lifetime;
//   * loader warps move global -> registers -> smem with 128-bit accesses (this is where norm-on-load will be
//     fused); MMA issue, TMEM double buffering and the epilogue (bias, GroupNorm statistics, residual
//     addend, bf16 NDHWC stores) follow conv_tc.cu.
// Warp roles (416 threads): warp 0 MMA issuer + TMEM allocator, warps 1-4 epilogue, warps 5-12 loaders
// (software-pipelined: the loads of slice i+1 are in flight while slice i is stored).
#include <stdlib.h>

#include "tc_common.cuh"

namespace b200seg {

constexpr int HT_W = 8, HT_H = 16;            // output tile (w, h); M = 128 rows = (hh, ww)
constexpr int HP_W = HT_W + 2, HP_H = HT_H + 2;
constexpr int kHaloMaxSlices = 16;
constexpr int kLoaderWarps = 8;
constexpr int kHaloThreads = 32 * (1 + 4 + kLoaderWarps);
constexpr int kMaxPieces = (HP_H * HP_W * 4 + 32 * kLoaderWarps - 1) / (32 * kLoaderWarps);   // Cin <= 32

struct HaloArgs {
  const bf16* x;
  const bf16* w;          // packed [tap][Cin/8][Cout][8]
  bf16* y;
  const bf16* addend;
  const float* bias;
  double* stats;
  long long xld, yld, ald;
  int N, D, H, W;
  int Cin, Cout;
  int kd;                 // 3 (3-D) or 1 (2-D)
  int tw, th;             // tiles along w, h
  int dchunk, ndchunks;   // output slices per work item, items along d
  int nitems;             // N * th * tw * ndchunks
  int nslices;            // ring size
  int tmem_cols;
  int swap_lbo_sbo;       // debug: swap the roles of the two descriptor strides
  int cp_async;           // 1: loaders use cp.async (zfill) + mbarrier completion; 0: register-staged copies
  // GroupNorm-backward sums fused into the data-gradient epilogue (b200seg_conv_bwdstats): the output of this conv is
  // g = dL/d(act) of the layer whose raw output is yfwd; with m = [yfwd*A + B > 0] (A, B = that layer's GroupNorm /
  // dropout coefficients, derived here from its statistics exactly as gn_cta_coefs does) the epilogue accumulates
  // sums[n][c][0] += g*m, sums[n][c][1] += g*m*yfwd  (stride 3: [2] = sum yfwd is taken from the forward statistics).
  const bf16* yfwd;       // nullptr: forward statistics mode (stats = [N][Cout][2])
  long long yfld;
  const double* gstats;   // [N][Cout][2] forward statistics of the layer being differentiated
  const float* ggamma;
  const float* gbeta;
  const float* gscale;    // [N][Cout] dropout scale or nullptr
  int ggroups;
  double gm;              // elements per group
  float geps;
  int exp;                // debug experiments (B200SEG_HALO_EXP bitmask): 1 no MMAs, 2 no epilogue work, 4 no loads
  long long* dbg;         // debug timeline (B200SEG_HALO_DBG=1): [role 0..3][64 events] clock64 stamps of CTA 0
};

__device__ __forceinline__ void halo_stamp(const HaloArgs& p, int role, int idx) {
  if (p.dbg != nullptr && blockIdx.x == 0 && idx < 64) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.dbg[role * 64 + idx] = t;
  }
}

__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

template <int CIN, int COUT, int KD>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const HaloArgs p) {
  PDL_ENTER();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  constexpr int CP = CIN / 8;                                // 8-channel planes
  constexpr uint32_t PLANE = HP_H * HP_W * 16u;              // bytes of one plane of one slice
  constexpr uint32_t SLICE = (uint32_t)CP * PLANE;
  constexpr int taps = KD * 9;
  constexpr uint32_t W_BYTES = (uint32_t)taps * CIN * COUT * 2u;
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + ((W_BYTES + 127u) & ~127u);
  uint8_t* tail = s_ring + (size_t)p.nslices * SLICE;
  tail = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tail) + 15) & ~(uintptr_t)15);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(tail);
  uint64_t* sempty = sfull + kHaloMaxSlices;
  uint64_t* tfull = sempty + kHaloMaxSlices;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_stat = reinterpret_cast<float*>(tmem_slot + 4);   // [4 epilogue warps][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int items_per_cta = (p.nitems + gridDim.x - 1) / gridDim.x;
  const int item_begin = blockIdx.x * items_per_cta;
  const int item_end = min(p.nitems, item_begin + items_per_cta);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslices; ++s) {
      // register loader: one arrive per loader warp; cp.async loader: one (deferred) arrive per loader thread
      mbar_init(&sfull[s], p.cp_async ? 32 * kLoaderWarps : kLoaderWarps);
      mbar_init(&sempty[s], 1);     // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // resident weights: straight 16-byte copy of the pre-packed smem image
  for (uint32_t i = threadIdx.x * 16u; i < W_BYTES; i += blockDim.x * 16u)
    *reinterpret_cast<uint4*>(s_w + i) = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.w) + i);
  for (int i = threadIdx.x; i < 8 * p.Cout; i += blockDim.x) s_stat[i] = 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int pd = KD / 2;

  auto decode = [&](int item, int& n, int& h0, int& w0, int& d0, int& nd) {
    int t = item;
    const int dc = t % p.ndchunks; t /= p.ndchunks;
    const int iw = t % p.tw; t /= p.tw;
    const int ih = t % p.th;
    n = t / p.th;
    w0 = iw * HT_W;
    h0 = ih * HT_H;
    d0 = dc * p.dchunk;
    nd = min(p.dchunk, p.D - d0);
  };

  if (warp == 0) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_lbo = p.swap_lbo_sbo ? HP_W * 16u : PLANE;
    const uint32_t a_sbo = p.swap_lbo_sbo ? PLANE : HP_W * 16u;
    const uint32_t b_lbo = p.swap_lbo_sbo ? 128u : (uint32_t)COUT * 16u;
    const uint32_t b_sbo = p.swap_lbo_sbo ? (uint32_t)COUT * 16u : 128u;
    const uint64_t a_hi = make_nosw_desc(0, a_lbo, a_sbo);
    const uint64_t b_hi = make_nosw_desc(0, b_lbo, b_sbo);
    const uint32_t ring_u32 = smem_u32(s_ring);
    const uint64_t b_base = b_hi | (uint64_t)(smem_u32(s_w) >> 4);
    constexpr int kchunks = CIN / 16;
    uint32_t gs = 0;   // slices consumed so far (ring position of the item's first slice)
    uint32_t go = 0;   // output slice-tiles produced so far
    for (int item = item_begin; item < item_end; ++item) {
      int n, h0, w0, d0, nd;
      decode(item, n, h0, w0, d0, nd);
      const int nsl = nd + KD - 1;
      for (int o = 0; o < nd; ++o, ++go) {
        // input slices o .. o+KD-1 of this item must have landed
        const int first_wait = (o == 0) ? 0 : KD - 1;
        for (int k = first_wait; k < KD; ++k) {
          const uint32_t sl = gs + o + k;
          mbar_wait(&sfull[sl % p.nslices], (sl / p.nslices) & 1u);
        }
        const uint32_t as = go & 1u;
        mbar_wait(&tempty[as], ((go >> 1) & 1u) ^ 1u);
        // slices written by cp.async (generic proxy) must be ordered before the tensor core's async-proxy reads
        fence_proxy_async();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tacc = tmem_base + as * (uint32_t)COUT;
          // straight-line issue: every descriptor is (per-slice base) + compile-time offset, in 16-byte units
#pragma unroll
          for (int kd_ = 0; kd_ < KD; ++kd_) {
            const uint32_t sl = gs + o + kd_;
            const uint64_t a_base = a_hi | (uint64_t)((ring_u32 + (sl % p.nslices) * SLICE) >> 4);
#pragma unroll
            for (int kh_ = 0; kh_ < 3; ++kh_)
#pragma unroll
              for (int kw_ = 0; kw_ < 3; ++kw_)
#pragma unroll
                for (int kc = 0; kc < kchunks; ++kc) {
                  constexpr uint32_t dummy = 0;
                  (void)dummy;
                  const uint32_t a_off = ((uint32_t)(2 * kc) * PLANE + (uint32_t)(kh_ * HP_W + kw_) * 16u) >> 4;
                  const uint32_t b_off = ((uint32_t)((((kd_ * 3 + kh_) * 3 + kw_) * kchunks + kc) * 2 * COUT) * 16u) >> 4;
                  umma_bf16(tacc, a_base + a_off, b_base + b_off, idesc, (kd_ | kh_ | kw_ | kc) != 0 ? 1u : 0u);
                }
          }
          umma_commit(&tfull[as]);
          // input slice o is no longer needed; at the end of the item neither are the trailing KD-1
          umma_commit(&sempty[(gs + o) % p.nslices]);
          if (o == nd - 1)
            for (int k = 1; k < KD; ++k) umma_commit(&sempty[(gs + o + k) % p.nslices]);
        }
        __syncwarp();
      }
      gs += (uint32_t)nsl;
    }
  } else if (warp >= 5) {
    // ===================================================== loaders (8 warps, software pipelined)
    constexpr int LT = 32 * kLoaderWarps;
    const int lt = threadIdx.x - 160;
    const int pieces = HP_H * HP_W * CP;
    // this thread's pieces: q = lt + j*LT  ->  (voxel v, plane) -> smem offset and (hh, ww)
    int poff[kMaxPieces], phh[kMaxPieces], pww[kMaxPieces];
    bool pval[kMaxPieces];
#pragma unroll
    for (int j = 0; j < kMaxPieces; ++j) {
      const int q = lt + j * LT;
      pval[j] = q < pieces;
      const int qq = pval[j] ? q : 0;
      const int v = qq / CP, plane = qq - v * CP;
      phh[j] = v / HP_W;
      pww[j] = v - phh[j] * HP_W;
      poff[j] = plane * (int)PLANE + v * 16;
      // global element offset of the plane is added per slice below
      pww[j] |= plane << 16;
    }
    // iterator over (item, slice)
    int it_item = item_begin, it_i = 0, it_nsl = 0, n = 0, h0 = 0, w0 = 0, d0 = 0, nd = 0;
    if (it_item < item_end) {
      decode(it_item, n, h0, w0, d0, nd);
      it_nsl = nd + KD - 1;
    }
    uint4 regs[kMaxPieces];
    auto issue_loads = [&]() {
      const int d = d0 - pd + it_i;
      const bool dok = (unsigned)d < (unsigned)p.D;
      const bf16* src = p.x + (((long long)n * p.D + (dok ? d : 0)) * p.H) * p.W * p.xld;
#pragma unroll
      for (int j = 0; j < kMaxPieces; ++j) {
        const int plane = pww[j] >> 16, ww = pww[j] & 0xffff;
        const int h = h0 - 1 + phh[j], w = w0 - 1 + ww;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (pval[j] && dok && (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W)
          val = *reinterpret_cast<const uint4*>(src + ((long long)h * p.W + w) * p.xld + plane * 8);
        regs[j] = val;
      }
    };
    uint32_t sl = 0;
    if (p.cp_async) {
      // fully asynchronous: as many slices in flight as the ring has free slots
      while (it_item < item_end) {
        const uint32_t slot = sl % p.nslices;
        mbar_wait(&sempty[slot], ((sl / p.nslices) & 1u) ^ 1u);
        const uint32_t dst = smem_u32(s_ring + (size_t)slot * SLICE);
        const int d = d0 - pd + it_i;
        const bool dok = (unsigned)d < (unsigned)p.D;
        const bf16* src = p.x + (((long long)n * p.D + (dok ? d : 0)) * p.H) * p.W * p.xld;
#pragma unroll
        for (int j = 0; j < kMaxPieces; ++j) {
          if (!pval[j]) continue;
          const int plane = pww[j] >> 16, ww = pww[j] & 0xffff;
          const int h = h0 - 1 + phh[j], w = w0 - 1 + ww;
          const bool ok = dok && (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
          const bf16* g = ok ? src + ((long long)h * p.W + w) * p.xld + plane * 8 : p.x;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)poff[j]), "l"(g),
                       "r"(ok ? 16 : 0)
                       : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sfull[slot])) : "memory");
        ++it_i;
        if (it_i == it_nsl) {
          ++it_item;
          it_i = 0;
          if (it_item < item_end) {
            decode(it_item, n, h0, w0, d0, nd);
            it_nsl = nd + KD - 1;
          }
        }
        ++sl;
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
    if (it_item < item_end) issue_loads();
    while (it_item < item_end) {
      const uint32_t slot = sl % p.nslices;
      mbar_wait(&sempty[slot], ((sl / p.nslices) & 1u) ^ 1u);
      uint8_t* dst = s_ring + (size_t)slot * SLICE;
#pragma unroll
      for (int j = 0; j < kMaxPieces; ++j)
        if (pval[j]) *reinterpret_cast<uint4*>(dst + poff[j]) = regs[j];
      // advance and put the next slice's loads in flight before signalling this one
      ++it_i;
      if (it_i == it_nsl) {
        ++it_item;
        it_i = 0;
        if (it_item < item_end) {
          decode(it_item, n, h0, w0, d0, nd);
          it_nsl = nd + KD - 1;
        }
      }
      fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[slot]);
      if (it_item < item_end) issue_loads();
      ++sl;
    }
    }
  } else {
    // ===================================================== epilogue warps 1..4
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rw = row % HT_W, rh = row / HT_W;
    const int etid = (warp - 1) * 32 + lane;
    uint32_t go = 0;
    int cur_n = -1;
    auto flush_stats = [&](int n) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (p.stats != nullptr && n >= 0) {
        for (int i = etid; i < 2 * p.Cout; i += 128) {
          const int which = i / p.Cout, c = i - which * p.Cout;
          double t = 0.0;
#pragma unroll
          for (int wq = 0; wq < 4; ++wq) {
            t += (double)s_stat[wq * 2 * p.Cout + i];
            s_stat[wq * 2 * p.Cout + i] = 0.f;
          }
          atomicAdd(p.stats + ((long long)n * p.Cout + c) * 2 + which, t);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    for (int item = item_begin; item < item_end; ++item) {
      int n, h0, w0, d0, nd;
      decode(item, n, h0, w0, d0, nd);
      if (n != cur_n) {
        if (cur_n >= 0 && p.stats != nullptr) flush_stats(cur_n);
        cur_n = n;
      }
      const int oh = h0 + rh, ow = w0 + rw;
      const bool valid = oh < p.H && ow < p.W;
      for (int o = 0; o < nd; ++o, ++go) {
        const long long vox = (((long long)n * p.D + (d0 + o)) * p.H + oh) * p.W + ow;
        const uint32_t as = go & 1u;
        mbar_wait(&tfull[as], (go >> 1) & 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * (uint32_t)p.Cout + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.Cout; c0 += 16) {
          float v[16];
          tmem_ld16(tacc + (uint32_t)c0, v);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias + c0 + j);
          }
          if (p.stats != nullptr) {
            float s[16], qq[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              s[j] = valid ? v[j] : 0.f;
              qq[j] = s[j] * s[j];
            }
#pragma unroll
            for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int j = 0; j < half; ++j) {
                const float keep_s = up ? s[j + half] : s[j];
                const float send_s = up ? s[j] : s[j + half];
                const float keep_q = up ? qq[j + half] : qq[j];
                const float send_q = up ? qq[j] : qq[j + half];
                s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
                qq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
              }
            }
            s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
            qq[0] += __shfl_xor_sync(0xffffffffu, qq[0], 1);
            if ((lane & 1) == 0) {
              const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
              // this lane is the only writer of its column in this warp's private row: no atomics, fixed order
              float* sw_ = s_stat + q * 2 * p.Cout;
              sw_[c0 + col] += s[0];
              sw_[p.Cout + c0 + col] += qq[0];
            }
          }
          if (valid) {
            if (p.addend != nullptr) {
              float r[16];
              load8(p.addend + vox * p.ald + c0, r);
              load8(p.addend + vox * p.ald + c0 + 8, r + 8);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += r[j];
            }
            store8(p.y + vox * p.yld + c0, v);
            store8(p.y + vox * p.yld + c0 + 8, v + 8);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
      }
    }
    if (p.stats != nullptr && cur_n >= 0) flush_stats(cur_n);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}


// ------------------------------------------------------------------------------------------------
// Input-slice-major form of the 3-D kernel (KD = 3): every staged input slice is multiplied ONCE per (kh, kw, K chunk)
// against the weights of all three kd taps side by side, B = [W(kd=2) | W(kd=1) | W(kd=0)] (N = 3*Cout columns), and
// accumulates into the three output slices (i-2, i-1, i) it contributes to, which are neighbouring column ranges
// of a TMEM ring of accumulators.  The M128 x N x K16 instruction is bound by the fetch of its A operand (128 voxel
// rows re-read from shared memory per K step whatever N is: measured ~27 + 0.6*N cycles), so 9 instructions of
// N = 3*Cout per input slice instead of 27 of N = Cout per output slice do the same MACs with a third of the A reads
// (16 channels: 9 x ~57 instead of 27 x ~37 cycles per slice-tile; 32 channels: 18 x ~87 instead of 54 x ~47).
//   * accumulators: ring of kAccRing slots of Cout columns; slot of output o = (running output count) % ring; an
//     output is complete when the input slice two further on has been issued (commit -> tfull[slot]);
//   * all MMAs accumulate (the instruction-wide accumulate flag cannot distinguish the fresh third of its columns):
//     the epilogue reads a slot into registers, hands it back ZEROED (tcgen05.st) at once and only then does its
//     arithmetic; the whole ring is zeroed once at kernel start.  (Clearing a slot with an extra non-accumulating
//     zero x zero MMA instead was measured: same time at 16 channels, 13 % slower at 32.)
//   * at the ends of an item's d-range the column range shrinks to the outputs that exist (N = Cout or 2*Cout,
//     weight rows offset accordingly), and a range that would wrap around the ring is issued in two pieces;
//   * the weight image in smem is the packed image permuted at load time to [(kh,kw)][Cin/8][kd reversed][Cout][8].
// The per-slice chain of one CTA (barrier waits, MMA issue by a single thread, commits, TMEM read-and-zero) is latency
// bound (tools/halo_timeline.py: ~460 ns of synchronisation + ~220 ns of MMA issue per slice, tensor pipe 15 % busy),
// so the kernel is built to run TWO CTAs per SM: 288 threads (1 MMA issuer, 4 epilogue, 4 loader warps), <= 112 KB of
// shared memory (4-8 slices), 128 / 256 TMEM columns, and the host splits d so that there are ~2 items per SM.
// ------------------------------------------------------------------------------------------------
constexpr int kAccRing = 8;
constexpr int kH3Loaders = 4;          // loader warps of the input-slice-major kernel (warps 5..8)
constexpr int kH3Threads = 32 * (1 + 4 + kH3Loaders);   // 288 threads, TWO CTAs per SM
constexpr int kHaloDepth = 6;          // cp.async groups (slices) in flight per loader thread; must be < ring slices

template <int CIN, int COUT, int NCTA, bool BWD>
__global__ void __launch_bounds__(kH3Threads, NCTA) conv_halo3_kernel(const HaloArgs p) {
  PDL_TRIGGER();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  constexpr int KD = 3;
  constexpr int CP = CIN / 8;                                // 8-channel planes
  constexpr uint32_t PLANE = HP_H * HP_W * 16u;              // bytes of one plane of one slice
  constexpr uint32_t SLICE = (uint32_t)CP * PLANE;
  constexpr int taps = 27;
  constexpr uint32_t W_BYTES = (uint32_t)taps * CIN * COUT * 2u;
  constexpr int R = kAccRing;
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + ((W_BYTES + 127u) & ~127u);
  uint8_t* tail = s_ring + (size_t)p.nslices * SLICE;
  tail = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tail) + 15) & ~(uintptr_t)15);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(tail);
  uint64_t* sempty = sfull + kHaloMaxSlices;
  uint64_t* tfull = sempty + kHaloMaxSlices;
  uint64_t* tempty = tfull + R;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + R);
  float* s_stat = reinterpret_cast<float*>(tmem_slot + 4);   // [4 epilogue warps][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int items_per_cta = (p.nitems + gridDim.x - 1) / gridDim.x;
  const int item_begin = blockIdx.x * items_per_cta;
  const int item_end = min(p.nitems, item_begin + items_per_cta);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslices; ++s) {
      mbar_init(&sfull[s], 1);                   // ONE arrive: by the loader warp that staged the slice
      mbar_init(&sempty[s], 1);                  // tcgen05.commit
    }
    for (int a = 0; a < R; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  for (int i = threadIdx.x; i < 8 * p.Cout; i += blockDim.x) s_stat[i] = 0.f;
  float* s_bias = s_stat + 8 * COUT;
  float* s_ab = s_bias + COUT;                                       // [2][Cout]: A, B of the sample being processed
  double* s_gd = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_ab + 2 * COUT) + 7) & ~(uintptr_t)7);   // [2][Cout] + [8][2]
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) halo_stamp(p, 0, 63);
  if (warp >= 1 && warp <= 4) {
    // the accumulator ring starts at zero (every MMA accumulates)
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = 0; c < R * COUT; c += 16) tmem_st16_zero(trow + (uint32_t)c);
    tc_fence_before();
  }
  // everything above is on-chip set-up and runs under the tail of the previous kernel; global memory from here on
  PDL_WAIT();
  // resident weights: 16-byte granules of the packed image [tap = (kd,kh,kw)][plane][co] -> [(kh,kw)][plane][2-kd][co]
  for (uint32_t g = threadIdx.x; g < W_BYTES / 16u; g += blockDim.x) {
    const uint32_t co = g % COUT;
    const uint32_t t1 = g / COUT;
    const uint32_t plane = t1 % CP;
    const uint32_t tap = t1 / CP;
    const uint32_t kd = tap / 9u, t9 = tap - kd * 9u;
    const uint32_t dst = ((t9 * CP + plane) * 3u + (2u - kd)) * COUT + co;
    *reinterpret_cast<uint4*>(s_w + dst * 16u) = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.w) + g * 16u);
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_bias[i] = p.bias != nullptr ? p.bias[i] : 0.f;
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  constexpr int pd = 1;

  auto decode = [&](int item, int& n, int& h0, int& w0, int& d0, int& nd) {
    int t = item;
    const int dc = t % p.ndchunks; t /= p.ndchunks;
    const int iw = t % p.tw; t /= p.tw;
    const int ih = t % p.th;
    n = t / p.th;
    w0 = iw * HT_W;
    h0 = ih * HT_H;
    d0 = dc * p.dchunk;
    nd = min(p.dchunk, p.D - d0);
  };

  if (warp == 0) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);   // | (N >> 3) << 17
    const uint64_t a_hi = make_nosw_desc(0, PLANE, HP_W * 16u);
    const uint64_t b_hi = make_nosw_desc(0, 3u * COUT * 16u, 128u);
    const uint32_t ring_u32 = smem_u32(s_ring);
    const uint32_t w_u32 = smem_u32(s_w);
    constexpr int kchunks = CIN / 16;
    uint32_t gs = 0;   // input slices consumed so far (ring position of the item's first slice)
    uint32_t go = 0;   // output slices started so far (accumulator slot of the item's first output)
    for (int item = item_begin; item < item_end; ++item) {
      int n, h0, w0, d0, nd;
      decode(item, n, h0, w0, d0, nd);
      const int nsl = nd + KD - 1;
      for (int i = 0; i < nsl; ++i) {
        const uint32_t sl = gs + (uint32_t)i;
        mbar_wait(&sfull[sl % p.nslices], (sl / p.nslices) & 1u);
        if (lane == 0) halo_stamp(p, 0, (int)sl);
        if (i < nd) {
          // output i is touched for the first time: its slot must have been drained (and zeroed) by the epilogue
          const uint32_t gn_ = go + (uint32_t)i;
          mbar_wait(&tempty[gn_ % R], ((gn_ / R) & 1u) ^ 1u);
        }
        // (the loader warp that staged the slice has waited for its cp.async group and executed
        //  fence.proxy.async BEFORE the arrive: the data is already ordered for the tensor core's async-proxy reads)
        if (!(p.exp & 8)) tc_fence_after();
        if (elect_one()) {
          const int lo = i - 2 > 0 ? i - 2 : 0;
          const int hi = i < nd - 1 ? i : nd - 1;
          const uint64_t a_base = a_hi | (uint64_t)((ring_u32 + (sl % p.nslices) * SLICE) >> 4);
          // pieces of [lo, hi] that are contiguous in the accumulator ring
          int o0 = lo;
          while (o0 <= hi) {
            const uint32_t slot0 = (go + (uint32_t)o0) % R;
            int cnt = hi - o0 + 1;
            if ((int)slot0 + cnt > R) cnt = R - (int)slot0;
            const uint32_t ncols = (uint32_t)cnt * COUT;
            const uint32_t idesc = idesc0 | ((ncols >> 3) << 17);
            const uint32_t tacc = tmem_base + slot0 * (uint32_t)COUT;
            const uint32_t row0 = (uint32_t)(2 - i + o0) * COUT;            // weight rows: kd = i - o  ->  (2 - kd) * Cout
            const uint64_t b_base = b_hi | (uint64_t)((w_u32 + row0 * 16u) >> 4);
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9)
#pragma unroll
              for (int kc = 0; kc < kchunks; ++kc) {
                const uint32_t a_off = ((uint32_t)(2 * kc) * PLANE + (uint32_t)((t9 / 3) * HP_W + (t9 % 3)) * 16u) >> 4;
                const uint32_t b_off = ((uint32_t)(t9 * CP + 2 * kc) * 3u * COUT * 16u) >> 4;
                if (!(p.exp & 1)) umma_bf16(tacc, a_base + a_off, b_base + b_off, idesc, 1u);
              }
            o0 += cnt;
          }
          if (i >= 2) umma_commit(&tfull[(go + (uint32_t)(i - 2)) % R]);   // output i-2 has all three contributions
          umma_commit(&sempty[sl % p.nslices]);                            // this input slice is consumed
          halo_stamp(p, 1, (int)sl);
        }
        __syncwarp();
      }
      gs += (uint32_t)nsl;
      go += (uint32_t)nd;
    }
  } else if (warp >= 5) {
    // ===================================================== loaders: 4 warps, warp w stages the input slices
    // sl = w (mod 4) on its own (all pieces of the slice, 16-byte cp.async with zero fill = conv padding), waits
    // for them to land and signals the slice with ONE mbarrier arrive.  The warps run independently: four slices
    // are in flight per CTA, eight per SM (two CTAs are resident).
    const int lw = warp - 5;
    constexpr int pieces = HP_H * HP_W * CP;
    int it_item = item_begin, it_i = 0, it_nsl = 0, n = 0, h0 = 0, w0 = 0, d0 = 0, nd = 0;
    auto load_item = [&]() {
      if (it_item < item_end) {
        decode(it_item, n, h0, w0, d0, nd);
        it_nsl = nd + KD - 1;
      }
    };
    auto advance = [&](int steps) {
      it_i += steps;
      while (it_item < item_end && it_i >= it_nsl) {
        it_i -= it_nsl;
        ++it_item;
        load_item();
      }
    };
    load_item();
    advance(lw);
    uint32_t sl = (uint32_t)lw, prev_sl = 0u;
    bool have_prev = false;
    auto signal = [&](uint32_t which) {
      fence_proxy_async();            // landed cp.async data -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[which % p.nslices]);
    };
    while (it_item < item_end) {
      const uint32_t slot = sl % p.nslices;
      mbar_wait(&sempty[slot], ((sl / p.nslices) & 1u) ^ 1u);
      const uint32_t dst = smem_u32(s_ring + (size_t)slot * SLICE);
      const int d = d0 - pd + it_i;
      const bool dok = (unsigned)d < (unsigned)p.D;
      const bf16* src = p.x + (((long long)n * p.D + (dok ? d : 0)) * p.H) * p.W * p.xld;
      if (!(p.exp & 4)) {
#pragma unroll 4
        for (int q = lane; q < pieces; q += 32) {
          const int v = q / CP, plane = q - v * CP;
          const int hh = v / HP_W, ww = v - hh * HP_W;
          const int h = h0 - 1 + hh, w = w0 - 1 + ww;
          const bool ok = dok && (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
          const bf16* g = ok ? src + ((long long)h * p.W + w) * p.xld + plane * 8 : p.x;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(plane * (int)PLANE + v * 16)),
                       "l"(g), "r"(ok ? 16 : 0)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (lane == 0) halo_stamp(p, 2, (int)sl);
      if (NCTA == 2) {
        // two CTAs per SM, 4-8 slice ring: one slice in flight per warp (eight per SM)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        signal(sl);
      } else {
        // one CTA per SM, deep ring: two cp.async groups in flight per warp; signal the older one
        if (have_prev) {
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          signal(prev_sl);
        }
        prev_sl = sl;
        have_prev = true;
      }
      advance(kH3Loaders);
      sl += kH3Loaders;
    }
    if (NCTA != 2 && have_prev) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      signal(prev_sl);
    }
  } else {
    // ===================================================== epilogue warps 1..4
    // GroupNorm statistics: per-thread (= per output row) partial sums in registers over ALL slices of an item, one
    // butterfly fold over the warp per item (not per slice); bias from shared memory; the zero hand-back of the
    // accumulator columns is issued right after the read and only waited for before the slot is released.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rw = row % HT_W, rh = row / HT_W;
    const int etid = (warp - 1) * 32 + lane;
    uint32_t go = 0;
    int cur_n = -1;
    constexpr bool bwd = BWD;                 // backward-sums epilogue (b200seg_conv_bwdstats): its own instantiation,
    constexpr int sstride = BWD ? 3 : 2;      // so the plain kernel does not carry its registers
    auto flush_stats = [&](int n) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (p.stats != nullptr && n >= 0) {
        for (int i = etid; i < 2 * p.Cout; i += 128) {
          const int which = i / p.Cout, c = i - which * p.Cout;
          double t = 0.0;
#pragma unroll
          for (int wq = 0; wq < 4; ++wq) {
            t += (double)s_stat[wq * 2 * p.Cout + i];
            s_stat[wq * 2 * p.Cout + i] = 0.f;
          }
          atomicAdd(p.stats + ((long long)n * p.Cout + c) * sstride + which, t);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    // A, B of sample n for the ReLU mask of the backward sums: the same expressions, in the same precision, as
    // gn_cta_coefs (elementwise.cu), so the mask here equals the one gn_bwd_apply derives later
    auto load_coefs = [&](int n) {
      const int cpg = COUT / p.ggroups;
      if (etid < COUT) {
        const double* q_ = p.gstats + ((long long)n * COUT + etid) * 2;
        s_gd[etid] = q_[0];
        s_gd[COUT + etid] = q_[1];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (etid < p.ggroups) {
        double sm = 0.0, sq = 0.0;
        for (int k_ = 0; k_ < cpg; ++k_) {
          sm += s_gd[etid * cpg + k_];
          sq += s_gd[COUT + etid * cpg + k_];
        }
        const double mean = sm / p.gm;
        double var = sq / p.gm - mean * mean;
        if (var < 0.0) var = 0.0;
        s_gd[2 * COUT + etid * 2 + 0] = mean;
        s_gd[2 * COUT + etid * 2 + 1] = rsqrt(var + (double)p.geps);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (etid < COUT) {
        const int g_ = etid / cpg;
        const double mean = s_gd[2 * COUT + g_ * 2 + 0], rstd = s_gd[2 * COUT + g_ * 2 + 1];
        const double sc = p.gscale ? (double)p.gscale[(long long)n * COUT + etid] : 1.0;
        const double ga = (double)p.ggamma[etid], be = (double)p.gbeta[etid];
        s_ab[etid] = (float)(rstd * ga * sc);
        s_ab[COUT + etid] = (float)((be - mean * rstd * ga) * sc);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    const bool want_stats = p.stats != nullptr;
    for (int item = item_begin; item < item_end; ++item) {
      int n, h0, w0, d0, nd;
      decode(item, n, h0, w0, d0, nd);
      if (n != cur_n) {
        if (cur_n >= 0 && want_stats) flush_stats(cur_n);
        cur_n = n;
        if (bwd) load_coefs(n);
      }
      const int oh = h0 + rh, ow = w0 + rw;
      const bool valid = oh < p.H && ow < p.W;
      float rs[COUT], rq[COUT];
#pragma unroll
      for (int j = 0; j < COUT; ++j) rs[j] = rq[j] = 0.f;
      for (int o = 0; o < nd; ++o, ++go) {
        const long long vox = (((long long)n * p.D + (d0 + o)) * p.H + oh) * p.W + ow;
        // backward-sums form: the epilogue's global operands (residual addend, the producer's raw output) are requested
        // BEFORE waiting for the accumulator
        uint4 adv[BWD ? COUT / 8 : 1], yfv[BWD ? COUT / 8 : 1];
        if constexpr (BWD) {
#pragma unroll
          for (int k_ = 0; k_ < COUT / 8; ++k_) {
            adv[k_] = make_uint4(0u, 0u, 0u, 0u);
            yfv[k_] = make_uint4(0u, 0u, 0u, 0u);
          }
          if (valid) {
            if (p.addend != nullptr) {
#pragma unroll
              for (int k_ = 0; k_ < COUT / 8; ++k_)
                adv[k_] = *reinterpret_cast<const uint4*>(p.addend + vox * p.ald + 8 * k_);
            }
#pragma unroll
            for (int k_ = 0; k_ < COUT / 8; ++k_)
              yfv[k_] = *reinterpret_cast<const uint4*>(p.yfwd + vox * p.yfld + 8 * k_);
          }
        }
        const uint32_t as = go % R;
        mbar_wait(&tfull[as], (go / R) & 1u);
        tc_fence_after();
        if (etid == 0) halo_stamp(p, 3, (int)go);
        const uint32_t tacc = tmem_base + as * (uint32_t)COUT + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16) {
          float v[16];
          tmem_ld16(tacc + (uint32_t)c0, v);
          {   // hand the columns back zeroed (all MMAs accumulate); completion is awaited before the release below
            const uint32_t z = 0u;
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                ::"r"(tacc + (uint32_t)c0), "r"(z)
                : "memory");
          }
          if (c0 + 16 >= COUT) {
            // every column of the slot is read and its zeroing issued: release it before the arithmetic
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
          }
          if (p.exp & 2) continue;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j);
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
          if (want_stats && valid && !bwd) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              rs[c0 + j] += v[j];
              rq[c0 + j] = fmaf(v[j], v[j], rq[c0 + j]);
            }
          }
          if (valid) {
            if (p.addend != nullptr) {
              if constexpr (BWD) {
#pragma unroll
                for (int h_ = 0; h_ < 2; ++h_) {
                  const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&adv[c0 / 8 + h_]);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(a2[j]);
                    v[8 * h_ + 2 * j] += f.x;
                    v[8 * h_ + 2 * j + 1] += f.y;
                  }
                }
              } else {
                float r[16];
                load8(p.addend + vox * p.ald + c0, r);
                load8(p.addend + vox * p.ald + c0 + 8, r + 8);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += r[j];
              }
            }
            if constexpr (!BWD) {
              store8(p.y + vox * p.yld + c0, v);
              store8(p.y + vox * p.yld + c0 + 8, v + 8);
            } else {
              // round g to bf16 once: the packed values are what is stored AND what the sums see (a separate reduce
              // pass would read the stored tensor); coefficients come as 128-bit shared-memory loads
#pragma unroll
              for (int h_ = 0; h_ < 2; ++h_) {
                uint4 pk;
                __nv_bfloat162* g2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int j = 0; j < 4; ++j) g2[j] = __floats2bfloat162_rn(v[8 * h_ + 2 * j], v[8 * h_ + 2 * j + 1]);
                *reinterpret_cast<uint4*>(p.y + vox * p.yld + c0 + 8 * h_) = pk;
                const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&yfv[c0 / 8 + h_]);
                const int cb = c0 + 8 * h_;
                const float4 a0 = *reinterpret_cast<const float4*>(s_ab + cb), a1 = *reinterpret_cast<const float4*>(s_ab + cb + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(s_ab + COUT + cb);
                const float4 b1 = *reinterpret_cast<const float4*>(s_ab + COUT + cb + 4);
                const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 yy = __bfloat1622float2(y2[j]);
                  const float2 gg = __bfloat1622float2(g2[j]);
                  const float d0_ = fmaf(yy.x, aa[2 * j], bb[2 * j]) > 0.f ? gg.x : 0.f;
                  const float d1_ = fmaf(yy.y, aa[2 * j + 1], bb[2 * j + 1]) > 0.f ? gg.y : 0.f;
                  rs[cb + 2 * j] += d0_;
                  rs[cb + 2 * j + 1] += d1_;
                  rq[cb + 2 * j] = fmaf(d0_, yy.x, rq[cb + 2 * j]);
                  rq[cb + 2 * j + 1] = fmaf(d1_, yy.y, rq[cb + 2 * j + 1]);
                }
              }
            }
          }
        }
      }
      if (want_stats) {
        // fold the 32 rows of this warp: after the butterfly lane pair (2k, 2k+1) holds column col(k) of the chunk
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16) {
          float s[16], qq[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s[j] = rs[c0 + j];
            qq[j] = rq[c0 + j];
          }
#pragma unroll
          for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < half; ++j) {
              const float keep_s = up ? s[j + half] : s[j];
              const float send_s = up ? s[j] : s[j + half];
              const float keep_q = up ? qq[j + half] : qq[j];
              const float send_q = up ? qq[j] : qq[j + half];
              s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
              qq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
            }
          }
          s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
          qq[0] += __shfl_xor_sync(0xffffffffu, qq[0], 1);
          if ((lane & 1) == 0) {
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            float* sw_ = s_stat + (warp - 1) * 2 * p.Cout;   // this warp's private row: no atomics, fixed order
            sw_[c0 + col] += s[0];
            sw_[p.Cout + c0 + col] += qq[0];
          }
        }
      }
    }
    if (want_stats && cur_n >= 0) flush_stats(cur_n);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) halo_stamp(p, 1, 63);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
static bool al16h(const void* p) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0; }

int conv_halo_channels_ok(int kind, int cin, int cout) {
  if (kind != B200SEG_K3) return 0;
  if (cin != 16 && cin != 32) return 0;
  if (cout != 16 && cout != 32) return 0;
  return 1;
}

// the packed-weight image: [tap][Cin/8][Cout][8]  (== T taps, K = Cin/8 chunks, N2 = Cout, N1 = 8)
int conv_halo_supported(int kind, int dims, const b200seg_tensor* x, int w_dtype, const b200seg_tensor* y,
                        const b200seg_tensor* addend) {
  (void)dims;
  if (w_dtype != B200SEG_BF16_HALO) return 0;
  if (!conv_halo_channels_ok(kind, x->c, y->c)) return 0;
  if (x->dtype != B200SEG_BF16 || y->dtype != B200SEG_BF16) return 0;
  if (addend && addend->dtype != B200SEG_BF16) return 0;
  if ((x->ld % 8) || (y->ld % 8) || !al16h(x->ptr) || !al16h(y->ptr)) return 0;
  if (addend && ((addend->ld % 8) || !al16h(addend->ptr))) return 0;
  if (x->d != y->d || x->h != y->h || x->w != y->w) return 0;
  return 1;
}

static int g_halo_init[64] = {0};

// 1 if conv_halo can fuse the GroupNorm-backward sums of the layer behind `y` into its epilogue (3-D layers only)
int conv_halo_bwdstats_ok(int dims) {
  static const int halo3 = [] {
    const char* e = getenv("B200SEG_HALO3");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return halo3 && dims == 3;
}

int conv_halo(int kind, int dims, const b200seg_tensor* x, const void* wpk, const float* bias, const b200seg_tensor* y,
              double* stats, const b200seg_tensor* addend, int device, cudaStream_t st, const b200seg_tensor* yfwd,
              const b200seg_gn* gn, double* sums) {
  (void)kind;
  const int maxsm = tc_max_smem(device);
  if (device >= 0 && device < 64 && !g_halo_init[device]) {
#define HALO_ATTR(CI, CO, K) \
  B200_CUDA(cudaFuncSetAttribute(conv_halo_kernel<CI, CO, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm))
    HALO_ATTR(16, 16, 3); HALO_ATTR(16, 32, 3); HALO_ATTR(32, 16, 3); HALO_ATTR(32, 32, 3);
    HALO_ATTR(16, 16, 1); HALO_ATTR(16, 32, 1); HALO_ATTR(32, 16, 1); HALO_ATTR(32, 32, 1);
#undef HALO_ATTR
#define HALO3_ATTR(CI, CO, NC) \
  B200_CUDA(cudaFuncSetAttribute(conv_halo3_kernel<CI, CO, NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm)); \
  B200_CUDA(cudaFuncSetAttribute(conv_halo3_kernel<CI, CO, NC, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
  B200_CUDA(cudaFuncSetAttribute(conv_halo3_kernel<CI, CO, NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm)); \
  B200_CUDA(cudaFuncSetAttribute(conv_halo3_kernel<CI, CO, NC, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100))
    HALO3_ATTR(16, 16, 2); HALO3_ATTR(16, 32, 1); HALO3_ATTR(32, 16, 2); HALO3_ATTR(32, 32, 1);
#undef HALO3_ATTR
    g_halo_init[device] = 1;
  }
  HaloArgs p;
  p.x = static_cast<const bf16*>(x->ptr);
  p.w = static_cast<const bf16*>(wpk);
  p.y = static_cast<bf16*>(y->ptr);
  p.addend = addend ? static_cast<const bf16*>(addend->ptr) : nullptr;
  p.bias = bias;
  p.stats = stats;
  p.xld = x->ld; p.yld = y->ld; p.ald = addend ? addend->ld : 0;
  p.N = x->n; p.D = x->d; p.H = x->h; p.W = x->w;
  p.Cin = x->c; p.Cout = y->c;
  p.kd = dims == 3 ? 3 : 1;
  p.yfwd = nullptr; p.yfld = 0; p.gstats = nullptr; p.ggamma = nullptr; p.gbeta = nullptr; p.gscale = nullptr;
  p.ggroups = 1; p.gm = 1.0; p.geps = 0.f;
  if (yfwd != nullptr) {
    B200_CHECK_ARG(conv_halo_bwdstats_ok(dims) && gn != nullptr && sums != nullptr && stats == nullptr,
                   "conv_halo: backward statistics need the 3-D input-slice-major kernel");
    B200_CHECK_ARG(same_geom(yfwd, y) && yfwd->dtype == B200SEG_BF16 && (yfwd->ld % 8) == 0 && al16h(yfwd->ptr) &&
                       gn->groups > 0 && (y->c % gn->groups) == 0 && gn->groups <= 8,
                   "conv_halo: bad forward tensor / GroupNorm reference for the backward statistics");
    p.yfwd = static_cast<const bf16*>(yfwd->ptr);
    p.yfld = yfwd->ld;
    p.gstats = gn->stats; p.ggamma = gn->gamma; p.gbeta = gn->beta; p.gscale = gn->scale;
    p.ggroups = gn->groups;
    p.gm = (double)(y->c / gn->groups) * (double)gn->vox;
    p.geps = gn->eps;
    p.stats = sums;
  }
  p.tw = (p.W + HT_W - 1) / HT_W;
  p.th = (p.H + HT_H - 1) / HT_H;
  const int cols = p.N * p.th * p.tw;
  const int sms = num_sms(device);
  static const int halo3 = [] {
    const char* e = getenv("B200SEG_HALO3");
    return (e && e[0] == '0') ? 0 : 1;            // input-slice-major kernel for the 3-D layers (default on)
  }();
  const bool use3 = halo3 && p.kd == 3;
  // split d so that the grid covers the chip (each extra chunk re-loads kd-1 halo slices).  The input-slice-major
  // kernel runs TWO CTAs per SM (and wants two items per SM) for 16 output channels; with 32 output channels the
  // epilogue's register statistics do not fit the 96-register budget of two CTAs (and 32 -> 32 has 55 KB of weights):
  // one CTA per SM with a deep ring (measured at 32 -> 32 @48^3: 30.7 us at one CTA, 34.8 us at two).
  const uint32_t slice = (uint32_t)(p.Cin / 8) * HP_H * HP_W * 16u;
  const uint32_t wbytes = ((uint32_t)(p.kd * 9) * p.Cin * p.Cout * 2u + 127u) & ~127u;
  const uint32_t tail = (2 * kHaloMaxSlices + (use3 ? 2 * kAccRing : 4)) * 8 + 16 + (use3 ? 11 : 8) * p.Cout * 4 + 64 + (use3 ? (2 * p.Cout + 16) * 8 + 8 : 0);
  const bool two = use3 && p.Cout == 16;       // 32 output channels: the epilogue's per-thread statistics need > 96 registers
  const int slots = two ? 2 * sms : sms;
  int ndch = 1;
  if (cols < slots) {
    ndch = slots / cols;               // items <= slots: one item per CTA, no second wave
    if (ndch > p.D) ndch = p.D;
    if (ndch < 1) ndch = 1;
  }
  p.dchunk = (p.D + ndch - 1) / ndch;
  p.ndchunks = (p.D + p.dchunk - 1) / p.dchunk;
  p.nitems = cols * p.ndchunks;
  const int budget = two ? 112 * 1024 : maxsm;
  int ns = (int)((budget - 256 - (int)wbytes - (int)tail) / (int)slice);
  if (ns > (two ? 8 : kHaloMaxSlices)) ns = two ? 8 : kHaloMaxSlices;
  B200_CHECK_ARG(ns >= p.kd + 1, "conv_halo: slices do not fit in shared memory");
  p.nslices = ns;
  int tc = 32;
  while (tc < (use3 ? kAccRing : 2) * p.Cout) tc *= 2;
  p.tmem_cols = tc;
  static const int swap = [] {
    const char* e = getenv("B200SEG_HALO_SWAP");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  p.swap_lbo_sbo = swap;
  static const int cpa = [] {
    const char* e = getenv("B200SEG_HALO_LOADER");
    return (e && e[0] == 'r') ? 0 : 1;          // "regs" selects the register-staged loader
  }();
  p.cp_async = cpa;
  static const int exp_bits = [] {
    const char* e = getenv("B200SEG_HALO_EXP");
    return e ? atoi(e) : 0;
  }();
  p.exp = exp_bits;
  p.dbg = nullptr;
  static const int dbg_on = [] {
    const char* e = getenv("B200SEG_HALO_DBG");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  static long long* dbg_buf = nullptr;
  if (dbg_on) {               // development aid only: allocates and synchronises
    if (dbg_buf == nullptr) cudaMalloc(&dbg_buf, 4 * 64 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 4 * 64 * sizeof(long long), st);
    p.dbg = dbg_buf;
  }
  const size_t smem_bytes = 128 + wbytes + (size_t)ns * slice + tail;
  int grid = slots < p.nitems ? slots : p.nitems;
#define HALO_LAUNCH(CI, CO, K) launch_k(conv_halo_kernel<CI, CO, K>, grid, kHaloThreads, smem_bytes, st, p)
#define HALO3_LAUNCH(CI, CO, NC)                                                                         \
  do {                                                                                                   \
    if (p.yfwd != nullptr) launch_k(conv_halo3_kernel<CI, CO, NC, true>, grid, kH3Threads, smem_bytes, st, p);  \
    else launch_k(conv_halo3_kernel<CI, CO, NC, false>, grid, kH3Threads, smem_bytes, st, p);            \
  } while (0)
  if (use3) {
    if (p.Cin == 16 && p.Cout == 16) HALO3_LAUNCH(16, 16, 2);
    else if (p.Cin == 16) HALO3_LAUNCH(16, 32, 1);
    else if (p.Cout == 16) HALO3_LAUNCH(32, 16, 2);
    else HALO3_LAUNCH(32, 32, 1);
  } else if (p.kd == 3) {
    if (p.Cin == 16 && p.Cout == 16) HALO_LAUNCH(16, 16, 3);
    else if (p.Cin == 16) HALO_LAUNCH(16, 32, 3);
    else if (p.Cout == 16) HALO_LAUNCH(32, 16, 3);
    else HALO_LAUNCH(32, 32, 3);
  } else {
    if (p.Cin == 16 && p.Cout == 16) HALO_LAUNCH(16, 16, 1);
    else if (p.Cin == 16) HALO_LAUNCH(16, 32, 1);
    else if (p.Cout == 16) HALO_LAUNCH(32, 16, 1);
    else HALO_LAUNCH(32, 32, 1);
  }
#undef HALO_LAUNCH
#undef HALO3_LAUNCH
  if (dbg_on && use3) {
    long long h[4 * 64];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[63];
    fprintf(stderr, "[halo3 dbg] Cin=%d Cout=%d D=%d H=%d W=%d items=%d dchunk=%d ns=%d  end=%lld ns\n", p.Cin, p.Cout, p.D, p.H,
            p.W, p.nitems, p.dchunk, p.nslices, h[64 + 63] - t0);
    for (int i = 0; i < 40; ++i)
      fprintf(stderr, "  sl %2d: load_issued %7lld  mma_sees_full %7lld  mma_issued %7lld | out %2d epi_sees_full %7lld\n", i,
              h[128 + i] ? h[128 + i] - t0 : -1, h[i] ? h[i] - t0 : -1, h[64 + i] ? h[64 + i] - t0 : -1, i,
              h[192 + i] ? h[192 + i] - t0 : -1);
  }
  B200_LAUNCH_CHECK();
  return B200SEG_OK;
}

}  // namespace b200seg
