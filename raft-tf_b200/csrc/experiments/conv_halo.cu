// tcgen05 implicit-GEMM convolution with HALO tiles for the multi-tap (3x3, 1x5, 5x1) convs.
//
// r01 measurement (profiles/r01_notes.md): the MMA loops of conv_tc.cu run at ~45 % tensor-pipe activity
// because every CTA pulls ~38 B/clk from L2 -- the sustained L2->SM rate -- while one 128x128x64
// k-iteration of the split GEMM needs 85 B/clk.  Most of those bytes are redundant: a 3x3 conv fetches
// every input pixel 9 times (once per tap), a 1x5 conv 5 times.
//
// Here the 128 output pixels of a CTA are a P x Q patch (P pixels along the "fast" image axis F, Q along
// the "slow" axis S, P a multiple of 8) and the filter taps are split the same way (kF x kS).  Per
// (64-channel chunk, F-tap) ONE TMA box of P x (Q + kS - 1) pixels is loaded -- rows ordered F-fastest --
// and the kS taps along S are views of that tile shifted by s*P rows = s*P*128 bytes, a multiple of the
// 1024-byte swizzle atom, so the UMMA descriptor simply starts s*P rows later.  A traffic drops by
// kS*Q/(Q+kS-1) (4x for 1x5 with Q=16), the B (weight) tiles stream through their own ring.
//   3x3, 5x1 : F = x, S = y, tensor map dims (C, w, h, B)
//   1x5      : F = y, S = x, tensor map dims (C, h, w, B)  (same memory, permuted strides)
// Two mbarrier rings (A halo tiles: 2 stages; B tiles: as many as fit), one producer thread, one MMA
// thread, 16 epilogue warps; arithmetic identical to conv_tc.cu.
#include <stdlib.h>
#include <string.h>

#include "../tc_common.cuh"

namespace rb {
using namespace tc;

constexpr int kHaloThreads = 576;
constexpr int kMaxHaloRows = 192;                       // rows (pixels) per staged A tile
constexpr int kAStageBytes = 2 * kMaxHaloRows * 128;    // hi + lo
constexpr int kAStages = 2;

template <int BLOCK_N>
struct HaloCfg {
  static constexpr int kBStageBytes = 2 * BLOCK_N * 128;
  static constexpr int kBudget = 226 * 1024 - 1280 - kAStages * kAStageBytes;
  static constexpr int kBStages = kBudget / kBStageBytes > 8 ? 8 : kBudget / kBStageBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : 256;
  static constexpr int kSmemBytes = kAStages * kAStageBytes + kBStages * kBStageBytes + 1024 + 256;
};

struct HaloGeom {
  int p_log2;        // P = pixels along F per tile
  int q;             // Q = 128 / P pixels along S
  int s_is_x;        // 1: S axis = x (1x5 convs), F = y; 0: S = y, F = x
  int kF, kS;        // taps along F and S
  int tiles_f, tiles_s;
  int halo_rows;     // P * (Q + kS - 1)
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const ConvParams p, const HaloGeom g) {
  using Cfg = HaloCfg<BLOCK_N>;
  constexpr int NB = Cfg::kBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + kAStages * kAStageBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smemB + NB * Cfg::kBStageBytes);
  uint64_t* a_empty = a_full + kAStages;
  uint64_t* b_full = a_empty + kAStages;
  uint64_t* b_empty = b_full + NB;
  uint64_t* tmem_full_bar = b_empty + NB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = p.dbg ? p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = gtime_ns();
  const int P = 1 << g.p_log2;
  const int tiles_per_img = g.tiles_f * g.tiles_s;
  const int b = blockIdx.x / tiles_per_img;
  const int trem = blockIdx.x - b * tiles_per_img;
  const int ts = trem / g.tiles_f, tf = trem - ts * g.tiles_f;
  const int f0 = tf << g.p_log2, s0 = ts * g.q;  // tile origin along F and S
  const int n0 = blockIdx.y * BLOCK_N;
  const int chunks = conv_chunks(p);
  const int groups = chunks * g.kF;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
    for (int i = 0; i < kAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);  // uniform register (see conv_tc.cu)
  if (dbg && threadIdx.x == 0) dbg[1] = gtime_ns();
  if (p.pdl_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (dbg && threadIdx.x == 0) dbg[2] = gtime_ns();

  if (warp == 0) {
    if (elect_one()) {
      // no integer division inside the loops: (chunk, F-tap) groups and ring slots advance incrementally
      const int padF = (g.kF - 1) / 2, padS = (g.kS - 1) / 2;
      const uint32_t a_bytes = 2u * (uint32_t)g.halo_rows * 128u;
      int ast = 0;
      uint32_t aph = 0;
      auto issue_a = [&](int cki, int f) {
        mbar_wait(&a_empty[ast], aph ^ 1);
        mbar_arrive_expect_tx(&a_full[ast], a_bytes);
        const int ck = conv_chunk(p, cki);
        uint8_t* dst = smemA + ast * kAStageBytes;
        const int c0 = p.in_choff + ck * 64;
        // box origin: F coordinate shifted by the F-tap, S coordinate by -padS (halo covers all S-taps)
        tma_load_4d(&tmA_hi, &a_full[ast], dst, c0, f0 + f - padF, s0 - padS, b);
        tma_load_4d(&tmA_lo, &a_full[ast], dst + g.halo_rows * 128, c0, f0 + f - padF, s0 - padS, b);
        if (++ast == kAStages) { ast = 0; aph ^= 1; }
      };
      const int wb = p.w_per_batch ? b : 0;
      int bst = 0;
      uint32_t bph = 0;
      issue_a(0, 0);
      for (int cki = 0; cki < chunks; ++cki) {
        const int ck = conv_chunk(p, cki);
        for (int f = 0; f < g.kF; ++f) {
          for (int s = 0; s < g.kS; ++s) {
            mbar_wait(&b_empty[bst], bph ^ 1);
            mbar_arrive_expect_tx(&b_full[bst], Cfg::kBStageBytes);
            // filter tap (ky,kx): the S-tap walks x for 1x5 convs, y otherwise
            const int ky = g.s_is_x ? f : s, kx = g.s_is_x ? s : f;
            const int kcol = (ky * p.kw + kx) * p.cin_pad + ck * 64;
            uint8_t* dst = smemB + bst * Cfg::kBStageBytes;
            tma_load_3d(&tmB_hi, &b_full[bst], dst, kcol, n0, wb);
            tma_load_3d(&tmB_lo, &b_full[bst], dst + BLOCK_N * 128, kcol, n0, wb);
            if (++bst == NB) { bst = 0; bph ^= 1; }
          }
          if (f + 1 < g.kF) issue_a(cki, f + 1);
          else if (cki + 1 < chunks) issue_a(cki + 1, 0);
        }
      }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // every warp issues the PDL trigger once its part is done
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_2n = umma_idesc_f16(2 * BLOCK_N);
      constexpr uint32_t idesc_n = umma_idesc_f16(BLOCK_N);
      int ast = 0, bst = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int grp = 0; grp < groups; ++grp) {
        mbar_wait(&a_full[ast], aph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smemA + ast * kAStageBytes);
        for (int s = 0; s < g.kS; ++s) {
          mbar_wait(&b_full[bst], bph);
          tc_fence_after();
          if (dbg && first) dbg[3] = gtime_ns();
          const uint32_t view = a_base + (uint32_t)(s * P) * 128u;  // shifted by s taps along S: multiple of 1024 B
          const uint64_t a_hi = umma_desc_sw128(view);
          const uint64_t a_lo = umma_desc_sw128(view + (uint32_t)g.halo_rows * 128u);
          const uint64_t b_all = umma_desc_sw128(smem_u32(smemB + bst * Cfg::kBStageBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t koff = (uint64_t)(k * 2);
            umma_f16(tmem_base, a_hi + koff, b_all + koff, idesc_2n, (!first || k != 0) ? 1u : 0u);
            umma_f16(tmem_base + BLOCK_N, a_lo + koff, b_all + koff, idesc_n, 1u);
          }
          first = false;
          umma_commit(&b_empty[bst]);
          if (++bst == NB) { bst = 0; bph ^= 1; }
        }
        umma_commit(&a_empty[ast]);
        if (++ast == kAStages) { ast = 0; aph ^= 1; }
      }
      umma_commit(tmem_full_bar);
      if (dbg) dbg[4] = gtime_ns();
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  } else {
    const int q = warp & 3;
    constexpr int kColsPerWarp = BLOCK_N >= 128 ? 32 : (BLOCK_N == 96 ? 32 : 16);
    constexpr int kGroups = BLOCK_N / kColsPerWarp;
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int fi = f0 + (r & (P - 1)), si = s0 + (r >> g.p_log2);
    const int px = g.s_is_x ? si : fi, py = g.s_is_x ? fi : si;
    const bool valid = (py < p.h) && (px < p.w) && (si < s0 + g.q);
    const int pix = (b * p.h + py) * p.w + px;
    if (grp >= kGroups) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (grp < kGroups) {
      mbar_wait_warp(tmem_full_bar, 0);
      tc_fence_after();
      const bool wide = epilogue_wide_ok(p);
      // PDL trigger: only now (MMA loop done, epilogue starting) may the next kernel's CTAs be scheduled -- triggering
      // at kernel start let them take the SMs this kernel's own late CTAs were waiting for (phase_times.py).
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      if (dbg && warp == 2 && lane == 0) dbg[5] = gtime_ns();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < kColsPerWarp; cc += 16) {
        const int c = grp * kColsPerWarp + cc;
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
    }
  }
  if (dbg && warp == 2 && lane == 0) dbg[6] = gtime_ns();
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[7] = gtime_ns();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host --------------------------------------------------------------------------------------------
static bool halo_geom(const ConvParams& p, HaloGeom* g) {
  const int s_is_x = (p.kh == 1 && p.kw > 1) ? 1 : 0;
  const int kS = s_is_x ? p.kw : p.kh, kF = s_is_x ? p.kh : p.kw;
  const int lenF = s_is_x ? p.h : p.w, lenS = s_is_x ? p.w : p.h;
  long best = -1;
  for (int pl = 3; pl <= 5; ++pl) {  // P = 8, 16, 32
    const int P = 1 << pl, Q = 128 >> pl;
    const int halo = P * (Q + kS - 1);
    if (halo > kMaxHaloRows) continue;
    const long tiles = (long)((lenF + P - 1) / P) * ((lenS + Q - 1) / Q);
    const long cost = tiles * 4096 + halo;  // fewest tiles first, then the smallest halo
    if (best < 0 || cost < best) {
      best = cost;
      g->p_log2 = pl; g->q = Q; g->s_is_x = s_is_x; g->kF = kF; g->kS = kS;
      g->tiles_f = (lenF + P - 1) / P; g->tiles_s = (lenS + Q - 1) / Q; g->halo_rows = halo;
    }
  }
  return best >= 0;
}

static int halo_block_n(int cout, long m_tiles) {
  if (cout <= 16) return 16;
  const int cand[4] = {128, 96, 64, 32};
  int bestn = 128;
  long best = -1;
  for (int i = 0; i < 4; ++i) {
    const int n = cand[i];
    const long tiles = m_tiles * ((cout + n - 1) / n);
    const long waves = (tiles + 147) / 148;
    // per S-tap k-iteration: tensor 4*1.5n cycles; L2->SM ~38 B/clk over the B tile (256n B) + amortised A
    long math = 6 * n, mem = (256L * n + 8192) / 38;
    const long cost = waves * (math > mem ? math : mem);
    if (best < 0 || cost < best) { best = cost; bestn = n; }
  }
  return bestn;
}

template <int BLOCK_N>
static int launch_halo_cfg(const ConvParams& p, const HaloGeom& g, const CUtensorMap* maps, cudaStream_t s) {
  using Cfg = HaloCfg<BLOCK_N>;
  static PerDeviceOnce attr_set;
  int dev = 0, rc_dev;
  if ((rc_dev = current_device(&dev))) return rc_dev;
  if (!attr_set.test(dev)) {
    RB_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.set(dev);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p.B * g.tiles_f * g.tiles_s, (p.cout + BLOCK_N - 1) / BLOCK_N);
  cfg.blockDim = dim3(kHaloThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // Programmatic dependent launch is OFF by default: measured on the update block (profiles/r01_notes.md) it hides
  // the ~3.5 us launch gap but the dependents' CTAs then wait just as long for the grid-completion signal (early
  // trigger: 222 us, trigger after the MMA loop: 220 us, no PDL: 217 us per step).  RAFT_B200_PDL=1 enables it.
  static const int pdl = getenv("RAFT_B200_PDL") ? 1 : 0;
  cfg.numAttrs = pdl;
  RB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<BLOCK_N>, maps[0], maps[1], maps[2], maps[3], p, g));
  RB_CHECK_LAUNCH("conv_halo_kernel");
  return RB_OK;
}

int launch_conv_halo(const ConvParams& p, cudaStream_t s, bool* handled) {
  *handled = false;
  // Opt-in (RAFT_B200_HALO=1).  Same-box A/B on the update block at batch 1 (profiles/r01_notes.md): 236.6 us with the
  // halo kernel for every multi-tap conv, 236.1 us with the persistent per-tap kernel -- the activation re-fetches it
  // removes are not what bounds the MMA loop (the weight tiles are), and at several waves it is 13 % slower.
  static const bool enabled = getenv("RAFT_B200_HALO") != nullptr && getenv("RAFT_B200_NO_HALO") == nullptr;
  if (!enabled) return RB_OK;
  if (p.cin_pad % 64 || p.in_stride % 8 || p.in_choff % 8) return RB_OK;
  HaloGeom g;
  if (!halo_geom(p, &g)) return RB_OK;
  const long m_tiles = (long)p.B * g.tiles_f * g.tiles_s;
  const int bn = halo_block_n(p.cout, m_tiles);
  // Same-box measurements (profiles/r01_notes.md): with a single wave of CTAs (batch 1) the halo kernel is ~2 % faster
  // than the per-tap kernel; with several waves (batch 8: 6 waves) its two-stage A ring stalls at every tap-row boundary
  // and it is 13 % slower (1317 vs 1142 us per update step) -- use it only for single-wave launches.
  if (m_tiles * ((p.cout + bn - 1) / bn) > 148) return RB_OK;
  static const int min_n = getenv("RAFT_B200_HALO_MIN_N") ? atoi(getenv("RAFT_B200_HALO_MIN_N")) : 0;  // tuning knob
  if (bn < min_n) return RB_OK;
  CUtensorMap maps[4];
  {
    const uint64_t C = (uint64_t)p.in_stride;
    const uint64_t sx = C * 2, sy = C * 2 * p.w, sb = C * 2 * p.w * p.h;  // byte strides of x, y, batch
    uint64_t dims[4], str[3];
    uint32_t box[4] = {64, 1u << g.p_log2, (uint32_t)(g.q + g.kS - 1), 1};
    dims[0] = C;
    if (g.s_is_x) { dims[1] = p.h; dims[2] = p.w; str[0] = sy; str[1] = sx; }
    else          { dims[1] = p.w; dims[2] = p.h; str[0] = sx; str[1] = sy; }
    dims[3] = p.B; str[2] = sb;
    int rc = cached_tmap(&maps[0], p.in_hi, 4, dims, str, box);
    if (rc == RB_OK) rc = cached_tmap(&maps[1], p.in_lo, 4, dims, str, box);
    if (rc != RB_OK) {
      if (g.s_is_x) return RB_OK;  // driver rejected the permuted strides: fall back to the per-tap kernel
      return rc;
    }
  }
  {
    const uint64_t ktot = (uint64_t)p.kh * p.kw * p.cin_pad;
    uint64_t dims[3] = {ktot, (uint64_t)p.cout_pad, (uint64_t)(p.w_per_batch ? p.B : 1)};
    uint64_t str[2] = {ktot * 2, ktot * 2 * p.cout_pad};
    uint32_t box[3] = {64, (uint32_t)bn, 1};
    int rc;
    if ((rc = cached_tmap(&maps[2], p.w_hi, 3, dims, str, box))) return rc;
    if ((rc = cached_tmap(&maps[3], p.w_lo, 3, dims, str, box))) return rc;
  }
  *handled = true;
  switch (bn) {
    case 16: return launch_halo_cfg<16>(p, g, maps, s);
    case 32: return launch_halo_cfg<32>(p, g, maps, s);
    case 64: return launch_halo_cfg<64>(p, g, maps, s);
    case 96: return launch_halo_cfg<96>(p, g, maps, s);
    default: return launch_halo_cfg<128>(p, g, maps, s);
  }
}

}  // namespace rb
