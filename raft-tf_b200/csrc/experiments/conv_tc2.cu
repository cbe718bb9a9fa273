// 2-CTA (cta_group::2) variant of the persistent implicit-GEMM conv of conv_tc.cu -- EXPERIMENTAL, opt-in with
// RAFT_B200_CTA2=1.
//
// Motivation (profiles/r01_notes.md): the MMA loop is bound by the ~38 B/clk an SM can ingest from L2, and for the
// wide GRU convs half of those bytes are the weight tile, which every CTA copies into its own shared memory.  With
// cta_group::2 a PAIR of CTAs (two pixel tiles, same cout tile) issues ONE tcgen05.mma with M = 256: each CTA holds
// its own 128 activation rows and only HALF of the weight rows (N/2 of B_hi and N/2 of B_lo); the tensor cores of
// both SMs read both halves.  Per k-iteration a CTA then ingests 32 KB + N*128 B instead of 32 KB + N*256 B.
//   * both CTAs issue their TMA loads with .cta_group::2 so that the bytes are credited to the LEADER's full barrier;
//     the follower additionally arrives on it remotely (2 arrivals + all bytes complete a phase);
//   * only the leader's MMA thread issues MMAs (3 per 16-wide k-slice: hi*hi -> D0, hi*lo -> D1, lo*hi -> D1) and
//     commits with multicast to both CTAs' empty / tmem_full barriers;
//   * each CTA's 16 epilogue warps drain their own TMEM half; the follower's warps arrive on the leader's tmem_empty
//     barrier remotely.
#include <stdlib.h>
#include <string.h>

#include "../tc_common.cuh"

namespace rb {
using namespace tc;

constexpr int kT2Threads = 576;
constexpr int kA2Bytes = 128 * 128;  // one 128-row x 64-channel fp16 tile

template <int BLOCK_N>
struct Tc2Cfg {
  static constexpr int kBHalfBytes = (BLOCK_N / 2) * 128;                 // N/2 rows of B_hi (or B_lo)
  static constexpr int kStageBytes = 2 * kA2Bytes + 2 * kBHalfBytes;      // per CTA
  static constexpr int kStages = (200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes;
  static constexpr int kAccCols = 2 * BLOCK_N;
  static constexpr int kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128) ? 128
                                   : (2 * kAccCols <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static constexpr int kColsPerWarp = BLOCK_N >= 96 ? 32 : 16;
  static constexpr int kGroups = BLOCK_N / kColsPerWarp;
};

struct Tile2Geom {
  int bw_log2, bh_log2, tiles_x, tiles_y, n_tiles, m_tiles, total_pairs;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kT2Threads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const ConvParams p, const Tile2Geom g) {
  using Cfg = Tc2Cfg<BLOCK_N>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);  // used in the leader
  uint64_t* empty_bar = full_bar + STAGES;                                             // each CTA's own
  uint64_t* tmem_full_bar = empty_bar + STAGES;                                        // [2], each CTA's own
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                                        // [2], used in the leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int chunks = conv_chunks(p), taps = p.kh * p.kw, kiters = taps * chunks;
  const int tiles_per_img = g.tiles_x * g.tiles_y;
  const int first = (int)(blockIdx.x >> 1), stride = (int)(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 2 * Cfg::kGroups * 4); }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before the paired TMEM allocation / any remote signal
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);  // uniform register (see conv_tc.cu)

  if (warp == 0) {
    if (elect_one()) {
      const int ph = (p.kh - 1) / 2, pw = (p.kw - 1) / 2;
      int s = 0;
      uint32_t phase = 0;
      for (int tile = first; tile < g.total_pairs; tile += stride) {
        const int mq = tile / g.n_tiles, nt = tile - mq * g.n_tiles;
        const int mt = 2 * mq + rank;
        const int b = mt / tiles_per_img, trem = mt - b * tiles_per_img;
        const int ty = trem / g.tiles_x, tx = trem - ty * g.tiles_x;
        const int y0 = ty << g.bh_log2, x0 = tx << g.bw_log2;
        const int nrow0 = nt * BLOCK_N + rank * (BLOCK_N / 2);  // this CTA's half of the weight rows
        const int wb = p.w_per_batch ? min(b, p.B - 1) : 0;
        int cki = 0, kx = 0, ky = 0, ck = conv_chunk(p, 0);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&empty_bar[s], phase ^ 1);
          uint8_t* st = smem + s * Cfg::kStageBytes;
          const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);  // bytes of BOTH CTAs
          const int t = ky * p.kw + kx;
          const int c0 = p.in_choff + ck * 64;
          tma_load_4d_2sm(&tmA_hi, lead_full, st, c0, x0 + kx - pw, y0 + ky - ph, b);
          tma_load_4d_2sm(&tmA_lo, lead_full, st + kA2Bytes, c0, x0 + kx - pw, y0 + ky - ph, b);
          const int kcol = t * p.cin_pad + ck * 64;
          tma_load_3d_2sm(&tmB_hi, lead_full, st + 2 * kA2Bytes, kcol, nrow0, wb);
          tma_load_3d_2sm(&tmB_lo, lead_full, st + 2 * kA2Bytes + Cfg::kBHalfBytes, kcol, nrow0, wb);
          if (!leader) mbar_arrive_remote(lead_full);
          if (++ky == p.kh) {
            ky = 0;
            if (++kx == p.kw) { kx = 0; ck = conv_chunk(p, ++cki); }
          }
          if (++s == STAGES) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16_m256(BLOCK_N);
      int s = 0, li = 0;
      uint32_t phase = 0;
      for (int tile = first; tile < g.total_pairs; tile += stride, ++li) {
        const int ab = li & 1;
        mbar_wait(&tmem_empty_bar[ab], ((li >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + ab * Cfg::kAccCols;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full_bar[s], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kA2Bytes);
          const uint64_t b_hi = umma_desc_sw128(st + 2 * kA2Bytes), b_lo = umma_desc_sw128(st + 2 * kA2Bytes + Cfg::kBHalfBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t koff = (uint64_t)(k * 2);
            umma_f16_2sm(acc, a_hi + koff, b_hi + koff, idesc, (it | k) != 0);
            umma_f16_2sm(acc + BLOCK_N, a_hi + koff, b_lo + koff, idesc, (it | k) != 0);
            umma_f16_2sm(acc + BLOCK_N, a_lo + koff, b_hi + koff, idesc, 1u);
          }
          umma_commit_2sm_mc(&empty_bar[s], (uint16_t)3);
          if (++s == STAGES) { s = 0; phase ^= 1; }
        }
        umma_commit_2sm_mc(&tmem_full_bar[ab], (uint16_t)3);
      }
    }
  } else {
    const int q = warp & 3, grp = (warp - 2) >> 2, r = q * 32 + lane;
    const bool wide = epilogue_wide_ok(p);
    if (grp < Cfg::kGroups) {
      int li = 0;
      for (int tile = first; tile < g.total_pairs; tile += stride, ++li) {
        const int mq = tile / g.n_tiles, nt = tile - mq * g.n_tiles;
        const int mt = 2 * mq + rank;
        const int b = mt / tiles_per_img, trem = mt - b * tiles_per_img;
        const int ty = trem / g.tiles_x, tx = trem - ty * g.tiles_x;
        const int y0 = ty << g.bh_log2, x0 = tx << g.bw_log2, n0 = nt * BLOCK_N;
        const int py = y0 + (r >> g.bw_log2), px = x0 + (r & ((1 << g.bw_log2) - 1));
        const bool valid = (py < p.h) && (px < p.w) && (mt < g.m_tiles);
        const int pix = (b * p.h + py) * p.w + px;
        const int ab = li & 1;
        mbar_wait(&tmem_full_bar[ab], (li >> 1) & 1);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + ab * Cfg::kAccCols;
#pragma unroll 1
        for (int cc = 0; cc < Cfg::kColsPerWarp; cc += 16) {
          const int c = grp * Cfg::kColsPerWarp + cc;
          if (n0 + c >= p.cout) break;
          uint32_t d0[16], d1[16];
          tmem_ld16(trow + c, d0);
          tmem_ld16(trow + BLOCK_N + c, d1);
          tmem_ld_wait(d0, d1);
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(d0[i]) + __uint_as_float(d1[i]) * kLoInv;
          if (valid) {
            if (wide) {
              epilogue_wide16(p, pix, n0 + c, v);
            } else {
              epilogue_store<8>(p, pix, n0 + c, v);
              epilogue_store<8>(p, pix, n0 + c + 8, v + 8);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tmem_empty_bar[ab]), 0));  // the leader's barrier
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N>
static int launch_tc2_cfg(const ConvParams& p, Tile2Geom g, const CUtensorMap* maps, cudaStream_t s) {
  using Cfg = Tc2Cfg<BLOCK_N>;
  static PerDeviceOnce attr_set;
  int dev = 0, rc_dev;
  if ((rc_dev = current_device(&dev))) return rc_dev;
  if (!attr_set.test(dev)) {
    RB_CHECK_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.set(dev);
  }
  const int num_sms = device_sm_count(dev);
  g.n_tiles = (p.cout + BLOCK_N - 1) / BLOCK_N;
  g.m_tiles = p.B * g.tiles_x * g.tiles_y;
  g.total_pairs = ((g.m_tiles + 1) / 2) * g.n_tiles;
  const int units = num_sms / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * (g.total_pairs < units ? g.total_pairs : units));
  cfg.blockDim = dim3(kT2Threads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<BLOCK_N>, maps[0], maps[1], maps[2], maps[3], p, g));
  RB_CHECK_LAUNCH("conv_tc2_kernel");
  return RB_OK;
}

// returns handled = false when the shape is not eligible (caller falls back to conv_tc.cu)
int launch_conv_tc2(const ConvParams& p, cudaStream_t s, int bn, int bw_log2, int bh_log2, int tiles_x, int tiles_y, bool* handled) {
  *handled = false;
  static const bool enabled = getenv("RAFT_B200_CTA2") != nullptr;
  if (!enabled || bn < 32 || p.B * tiles_x * tiles_y < 2) return RB_OK;
  Tile2Geom g;
  memset(&g, 0, sizeof(g));
  g.bw_log2 = bw_log2; g.bh_log2 = bh_log2; g.tiles_x = tiles_x; g.tiles_y = tiles_y;
  CUtensorMap maps[4];
  {
    uint64_t dims[4] = {(uint64_t)p.in_stride, (uint64_t)p.w, (uint64_t)p.h, (uint64_t)p.B};
    uint64_t str[3] = {(uint64_t)p.in_stride * 2, (uint64_t)p.in_stride * 2 * p.w, (uint64_t)p.in_stride * 2 * p.w * p.h};
    uint32_t box[4] = {64, 1u << bw_log2, 1u << bh_log2, 1};
    int rc;
    if ((rc = cached_tmap(&maps[0], p.in_hi, 4, dims, str, box))) return rc;
    if ((rc = cached_tmap(&maps[1], p.in_lo, 4, dims, str, box))) return rc;
  }
  {
    const uint64_t ktot = (uint64_t)p.kh * p.kw * p.cin_pad;
    uint64_t dims[3] = {ktot, (uint64_t)p.cout_pad, (uint64_t)(p.w_per_batch ? p.B : 1)};
    uint64_t str[2] = {ktot * 2, ktot * 2 * p.cout_pad};
    uint32_t box[3] = {64, (uint32_t)(bn / 2), 1};
    int rc;
    if ((rc = cached_tmap(&maps[2], p.w_hi, 3, dims, str, box))) return rc;
    if ((rc = cached_tmap(&maps[3], p.w_lo, 3, dims, str, box))) return rc;
  }
  *handled = true;
  switch (bn) {
    case 32: return launch_tc2_cfg<32>(p, g, maps, s);
    case 64: return launch_tc2_cfg<64>(p, g, maps, s);
    case 96: return launch_tc2_cfg<96>(p, g, maps, s);
    default: return launch_tc2_cfg<128>(p, g, maps, s);
  }
}

}  // namespace rb
