// tcgen05 + TMA weight-gradient kernel for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
//   dwp[t][ka][kb] += sum_{n,o} a[n, o + t - p, ka] * b[n, o, kb]        (stride-1 kinds: 3x3x3 / 3x3 / 1x1)
//
// GEMM view: M = flattened (tap, ka) rows, N = kb, K = output voxels.  Both operands sit in shared
// memory exactly as TMA delivers NDHWC boxes -- [voxel rows][channels] -- which is the MN-major
// canonical UMMA layout (rows = K index, 32/64/128-byte swizzled rows of MN elements), so no
// transposition is ever materialised:
//   A stage : 128/CA sub-tiles, each the activation box of one (tap, 16/32/64-channel block) shifted
//             by the tap offset (TMA zero fill = conv padding); consecutive sub-tiles are consecutive
//             MN atoms (descriptor LBO = sub-tile bytes), i.e. several taps are STACKED along M so
//             that 16- and 32-channel layers still fill the M = 128 instruction.
//   B stage : the dy box of the same 128 voxels, kb channels in 64-channel sub-tiles.
//   MMA     : 8 x tcgen05.mma (M128 x N=kb x K16 voxels) per stage, a_major = b_major = MN.
// A CTA owns a contiguous chunk of voxel tiles and up to 512/kb accumulator blocks (128 rows each)
// resident in TMEM for its whole lifetime; the epilogue adds them into dwp with vectorised fp32 atomics
// (split-K over the grid).  Warp roles as in conv_tc.cu.
#include <stdlib.h>

#include "tc_common.cuh"

namespace b200seg {

struct WgArgs {
  float* dwp;
  int N, D, H, W;
  int Ka, Kb, Rtot;          // rows per tap, columns, taps*Ka
  int kd, kh, kw, pd, ph, pw, sd, sh, sw;
  int bw, bh, bd, tw, th, td, ntiles;
  int CA, CB;                // channels per A / B sub-tile (<= 64)
  int mblocks_total, mb_per_cta;
  int nstages, tmem_cols;
};

constexpr int kWgMaxStages = 6;

__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB, const WgArgs p) {
  PDL_TRIGGER();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t A_SUB = 128u * (uint32_t)p.CA * 2u;          // bytes of one A sub-tile
  const uint32_t A_BYTES = 128u * 128u * 2u;                  // 128/CA sub-tiles
  const uint32_t B_SUB = 128u * (uint32_t)p.CB * 2u;
  const uint32_t nbsub = (uint32_t)(p.Kb / p.CB);
  const uint32_t B_BYTES = B_SUB * nbsub;
  const uint32_t STAGE = A_BYTES + B_BYTES;
  uint8_t* tail = smem + (size_t)p.nstages * STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* tfull = empty + kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);
  int4* s_sub = reinterpret_cast<int4*>(tmem_slot + 4);       // [mb_per_cta][128/CA] : channel, dw, dh, dd of a sub-tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_per_cta = (p.ntiles + gridDim.x - 1) / gridDim.x;
  const int tile_begin = blockIdx.x * tiles_per_cta;
  const int tile_end = min(p.ntiles, tile_begin + tiles_per_cta);
  const int mb0 = blockIdx.y * p.mb_per_cta;
  const int nmb = min(p.mb_per_cta, p.mblocks_total - mb0);
  const int nasub = 128 / p.CA;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // tap decode of every A sub-tile this CTA will ever load (keeps integer division out of the producer loop)
  for (int i = threadIdx.x; i < nmb * nasub; i += blockDim.x) {
    const int mb = i / nasub, j = i - mb * nasub;
    const int R = (mb0 + mb) * 128 + j * p.CA;
    int4 e = make_int4(-1, 0, 0, 0);
    if (R < p.Rtot) {
      const int tap = R / p.Ka;
      e.x = R - tap * p.Ka;
      e.y = tap % p.kw - p.pw;
      e.z = (tap / p.kw) % p.kh - p.ph;
      e.w = tap / (p.kw * p.kh) - p.pd;
    }
    s_sub[i] = e;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PDL_WAIT();               // everything above is on-chip set-up: it runs under the tail of the previous kernel
  const bool has_work = tile_begin < tile_end && nmb > 0;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (has_work && elect_one()) {
      uint32_t it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int t = tile;
        const int iw = t % p.tw; t /= p.tw;
        const int ih = t % p.th; t /= p.th;
        const int id = t % p.td;
        const int n = t / p.td;
        const int w0 = iw * p.bw, h0 = ih * p.bh, d0 = id * p.bd;
        for (int mb = 0; mb < nmb; ++mb, ++it) {
          const uint32_t s = it % p.nstages;
          const uint32_t ph = (it / p.nstages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sa = smem + (size_t)s * STAGE;
          // the last M block may be partial: skip (and do not expect) the sub-tiles that do not exist
          int valid = 0;
          for (int j = 0; j < nasub; ++j) valid += s_sub[mb * nasub + j].x >= 0 ? 1 : 0;
          mbar_expect_tx(&full[s], (uint32_t)valid * A_SUB + B_BYTES);
          for (int j = 0; j < valid; ++j) {
            const int4 e = s_sub[mb * nasub + j];
            tma_load_5d(&tmA, sa + (size_t)j * A_SUB, &full[s], e.x, w0 * p.sw + e.y, h0 * p.sh + e.z, d0 * p.sd + e.w, n);
          }
          for (uint32_t j = 0; j < nbsub; ++j)
            tma_load_5d(&tmB, sa + A_BYTES + (size_t)j * B_SUB, &full[s], (int)j * p.CB, w0, h0, d0, n);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (has_work) {
      // c_format f32 | a,b bf16 | a_major = b_major = MN | N | M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.Kb >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t swa = (uint32_t)p.CA * 2u, swb = (uint32_t)p.CB * 2u;
      uint32_t it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        for (int mb = 0; mb < nmb; ++mb, ++it) {
          const uint32_t s = it % p.nstages;
          const uint32_t ph = (it / p.nstages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + (size_t)s * STAGE);
            const uint64_t adesc = make_mnmajor_desc(sa, swa, A_SUB);
            const uint64_t bdesc = make_mnmajor_desc(sa + A_BYTES, swb, B_SUB);
            const uint32_t tacc = tmem_base + (uint32_t)(mb * p.Kb);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              // 16 voxel rows per K step: advance the start address by 16 rows of the operand's row pitch
              const uint64_t ka = (uint64_t)((16u * swa * (uint32_t)k) >> 4);
              const uint64_t kb = (uint64_t)((16u * swb * (uint32_t)k) >> 4);
              umma_bf16(tacc, adesc + ka, bdesc + kb, idesc, (tile != tile_begin || k != 0) ? 1u : 0u);
            }
            umma_commit(&empty[s]);
            if (tile == tile_end - 1 && mb == nmb - 1) umma_commit(&tfull[0]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================================================== epilogue warps (2..5)
    if (has_work) {
      const int q = warp & 3;
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      for (int mb = 0; mb < nmb; ++mb) {
        const int R = (mb0 + mb) * 128 + q * 32 + lane;
        const bool valid = R < p.Rtot;
        const uint32_t tacc = tmem_base + (uint32_t)(mb * p.Kb) + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.Kb; c0 += 16) {
          float v[16];
          tmem_ld16(tacc + (uint32_t)c0, v);
          if (valid) {
            float* dst = p.dwp + (size_t)R * p.Kb + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
static bool al16w(const void* p) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0; }

static int sub_channels(int c) {
  if (c % 64 == 0) return 64;
  if (c == 32) return 32;
  if (c == 16) return 16;
  return 0;
}

int wgrad_tc_init(int device, int maxsm) {
  (void)device;
  B200_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
  return B200SEG_OK;
}

int wgrad_tc_supported(int kind, int dims, const b200seg_tensor* a, const b200seg_tensor* b) {
  ConvGeom g;
  if (conv_geometry(kind, dims, &g) != 0 || g.up) return 0;
  if (a->dtype != B200SEG_BF16 || b->dtype != B200SEG_BF16) return 0;
  if (sub_channels(a->c) == 0 || sub_channels(b->c) == 0) return 0;
  if (a->c > 256 || b->c > 256) return 0;
  if ((a->ld % 8) || (b->ld % 8) || !al16w(a->ptr) || !al16w(b->ptr)) return 0;
  if (a->n != b->n || a->d != b->d * g.sd || a->h != b->h * g.sh || a->w != b->w * g.sw) return 0;
  return 1;
}

static int encode_box_map(CUtensorMap* tm, const b200seg_tensor* t, int cbox, int bw, int bh, int bd, int sw_ = 1,
                          int sh_ = 1, int sd_ = 1) {
  EncodeTiledFn enc = tc_encode_fn();
  if (!enc) return B200SEG_ECUDA;
  const CUtensorMapSwizzle sw = cbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                           : (cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  cuuint64_t dims5[5] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->d, (cuuint64_t)t->n};
  cuuint64_t strides[4] = {(cuuint64_t)t->ld * 2, (cuuint64_t)t->ld * 2 * t->w, (cuuint64_t)t->ld * 2 * t->w * t->h,
                           (cuuint64_t)t->ld * 2 * t->w * t->h * t->d};
  cuuint32_t box[5] = {(cuuint32_t)cbox, (cuuint32_t)(bw * sw_), (cuuint32_t)(bh * sh_), (cuuint32_t)(bd * sd_), 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)sw_, (cuuint32_t)sh_, (cuuint32_t)sd_, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, t->ptr, dims5, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "wgrad_tc: cuTensorMapEncodeTiled failed with %d", (int)r);
  return B200SEG_OK;
}

int wgrad_tc(int kind, int dims, const b200seg_tensor* a, const b200seg_tensor* b, float* dwp, int device,
             cudaStream_t st) {
  ConvGeom g;
  conv_geometry(kind, dims, &g);
  WgArgs p;
  p.dwp = dwp;
  p.N = b->n; p.D = b->d; p.H = b->h; p.W = b->w;
  p.Ka = a->c; p.Kb = b->c;
  p.kd = g.kd; p.kh = g.kh; p.kw = g.kw; p.pd = g.pd; p.ph = g.ph; p.pw = g.pw;
  p.sd = g.sd; p.sh = g.sh; p.sw = g.sw;
  p.Rtot = g.kd * g.kh * g.kw * p.Ka;
  tc_pick_box(p.W, p.H, p.D, &p.bw, &p.bh, &p.bd);
  p.tw = (p.W + p.bw - 1) / p.bw;
  p.th = (p.H + p.bh - 1) / p.bh;
  p.td = (p.D + p.bd - 1) / p.bd;
  p.ntiles = p.N * p.td * p.th * p.tw;
  p.CA = sub_channels(p.Ka);
  p.CB = sub_channels(p.Kb);
  p.mblocks_total = (p.Rtot + 127) / 128;
  // Decomposition: grid.y = groups of accumulator blocks (128 rows of (tap, ka) each), grid.x = split-K over the
  // voxel tiles.  Every K split adds its partial dW with fp32 atomics, i.e. grid.x * |dW| bytes of L2 atomic traffic:
  // at the small pyramid levels (few voxel tiles, large dW) that flush, not the K loop, is the run time.  The B tile
  // is re-loaded per accumulator block anyway, so fewer blocks per CTA (more groups, fewer K splits) costs nothing
  // in the loop.  Pick the block count per CTA that minimises  stages * t_stage + flush / rate.
  const int sms_ = num_sms(device);
  const int max_mb = (512 / p.Kb) < p.mblocks_total ? (512 / p.Kb) : p.mblocks_total;
  {
    const double stage_bytes = 128.0 * 128.0 * 2.0 + 128.0 * p.Kb * 2.0;
    const double stage_rows = 128.0 * (128 / p.CA) + 128.0 * (p.Kb / p.CB);             // TMA rows per stage
    double t_stage = stage_bytes / 88e3;                                                // us: ~88 GB/s ingest per SM
    if (t_stage < stage_rows * 1.06e-3) t_stage = stage_rows * 1.06e-3;                 // ~1 row/ns for short rows
    const double dw_bytes = (double)p.Rtot * p.Kb * 4.0;
    const double last_w = (double)(p.Rtot - 128 * (p.mblocks_total - 1)) / 128.0;       // the last block may be partial
    auto estimate = [&](int mb) {
      const int gy_ = (p.mblocks_total + mb - 1) / mb;
      int gx_ = sms_ / gy_;
      if (gx_ < 1) gx_ = 1;
      if (gx_ > p.ntiles) gx_ = p.ntiles;
      const int mbc = (p.mblocks_total + gy_ - 1) / gy_;
      // heaviest row group (the last block of the last group may be partial)
      const double heavy = gy_ == 1 ? (double)(mbc - 1) + last_w : (double)mbc;
      const double stages = (double)((p.ntiles + gx_ - 1) / gx_) * heavy;
      const double waves = gy_ > sms_ ? (double)((gy_ + sms_ - 1) / sms_) : 1.0;
      return waves * stages * t_stage + gx_ * dw_bytes / 1.5e6;                          // us: ~1.5 TB/s of atomics
    };
    // default: as many accumulator blocks per CTA as TMEM holds (fewest CTAs touching a tile); switch only for a
    // clear predicted gain (the model is calibrated on the VNet3d layer shapes, tools/microbench_ops.py)
    int best_mb = max_mb;
    double best = estimate(max_mb) * 0.85;
    for (int mb = 1; mb < max_mb; ++mb) {
      const double t = estimate(mb);
      if (t < best - 1e-9) {
        best = t;
        best_mb = mb;
      }
    }
    static const int forced = [] {
      const char* e = getenv("B200SEG_WGRAD_MB");
      return e ? atoi(e) : 0;
    }();
    p.mb_per_cta = forced > 0 ? (forced < max_mb ? forced : max_mb) : best_mb;
  }
  const int gy = (p.mblocks_total + p.mb_per_cta - 1) / p.mb_per_cta;
  // balance the accumulator blocks over the row groups
  p.mb_per_cta = (p.mblocks_total + gy - 1) / gy;
  int cols = 32;
  while (cols < p.mb_per_cta * p.Kb) cols *= 2;
  p.tmem_cols = cols;
  const uint32_t stage = 128u * 128u * 2u + 128u * (uint32_t)p.Kb * 2u;
  const uint32_t tail = (2 * kWgMaxStages + 2) * 8 + 16 + 16 + (uint32_t)(p.mb_per_cta * (128 / p.CA)) * 16;
  const int maxsm = tc_max_smem(device);
  int nst = (int)((maxsm - 1024 - (int)tail - 256) / (int)stage);
  if (nst > kWgMaxStages) nst = kWgMaxStages;
  B200_CHECK_ARG(nst >= 2, "wgrad_tc: stage does not fit in shared memory (Kb=%d)", p.Kb);
  p.nstages = nst;
  const size_t smem_bytes = 1024 + (size_t)nst * stage + tail + 128;
  CUtensorMap tmA, tmB;
  int rc = encode_box_map(&tmA, a, p.CA, p.bw, p.bh, p.bd, g.sw, g.sh, g.sd);
  if (rc != B200SEG_OK) return rc;
  rc = encode_box_map(&tmB, b, p.CB, p.bw, p.bh, p.bd);
  if (rc != B200SEG_OK) return rc;
  int gx = num_sms(device) / gy;
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  dim3 grid(gx, gy);
  launch_k(wgrad_tc_kernel, grid, 192, smem_bytes, st, tmA, tmB, p);
  B200_LAUNCH_CHECK();
  return B200SEG_OK;
}

}  // namespace b200seg
