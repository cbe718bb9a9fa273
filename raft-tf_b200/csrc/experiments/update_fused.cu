// Fused update step: ALL tensor-core convs of one update-block application in ONE persistent kernel launch.
//
// At batch 1 a conv of the update block is a single wave of 55-112 tiles that takes ~20 us, of which ~3.5 us are the
// kernel-to-kernel gap of the CUDA graph and ~2 us the prologue / pipeline fill (profiles/r01_notes.md) -- ten
// dependent convs per iteration make that fixed cost ~25 % of the step.  Here one kernel (<= 1 CTA per SM, cooperative
// launch so that all CTAs are co-resident) walks a JOB LIST: job = one conv with its own tensor maps, tile width and
// fused epilogue.  The warp roles and the mbarrier rings are those of conv_tc.cu (TMA producer / MMA issuer / 16
// epilogue warps, double-buffered TMEM accumulators) and simply continue across jobs; a job that depends on earlier
// jobs is separated from them by a grid barrier:
//   * every epilogue warp, after its last tile of job j: fence (generic + async proxy), then counter[j] += 1;
//   * the TMA producer of every CTA, before the first activation load of a dependent job: spins until
//     counter[j-1] == 16 * gridDim.x (acquire), then a proxy fence -- the WEIGHT tiles of the first pipeline stages
//     are requested before the wait, so they travel during the barrier;
//   * the epilogue warps of the dependent job execute a gpu-scope fence after their first accumulator wait
//     (their z / h / addend reads were written by other CTAs inside the same launch; L1 is not coherent).
// BLOCK_N is a run-time property of the job (instruction descriptor, stage layout and epilogue column groups are
// computed per job); the ring has 3 stages of 64 KB (the stage sweep is flat beyond 3).
#include <stdlib.h>
#include <string.h>

#include "../tc_common.cuh"

namespace rb {
using namespace tc;

constexpr int kFThreads = 576;
constexpr int kFStages = 3;
constexpr int kFStageStride = 64 * 1024;  // A_hi 16 KB | A_lo 16 KB | [B_hi ; B_lo] up to 32 KB
constexpr int kFATile = 128 * 128;
constexpr int kFSmemBytes = kFStages * kFStageStride + 1024 + 256;

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kFThreads, 1) fused_conv_kernel(const __grid_constant__ FusedJobs J) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kFStages * kFStageStride);
  uint64_t* empty_bar = full_bar + kFStages;
  uint64_t* tmem_full_bar = empty_bar + kFStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x;
  const int NS = J.stages;  // <= kFStages (tuning knob RAFT_B200_FUSED_STAGES)

  if (warp == 0 && lane == 0) {
    for (int j = 0; j < J.n; ++j)
      for (int k = 0; k < 4; ++k) prefetch_tmap(&J.job[j].m[k]);
    for (int s = 0; s < kFStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 16); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // warp-wide OR of identical values: lands in a UNIFORM register, so that ptxas does not wrap every tcgen05.mma of the
  // single issuing lane in an elect / R2UR.BROADCAST "waterfall" loop (that was ~50 cycles per MMA, 8 MMAs per k-iteration)
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);

  if (warp == 0) {
    if (elect_one()) {
      int rs = 0;          // ring slot and its phase (continue across tiles and jobs; no integer division per k-iteration)
      uint32_t rph = 0;
      for (int j = 0; j < J.n; ++j) {
        const FusedJob& jb = J.job[j];
        const ConvParams& p = jb.p;
        const TileGeom& g = jb.g;
        const int bn = jb.block_n, btile = bn * 128;
        const int wi = J.whatif;
        const uint32_t stage_tx = ((wi & 8) ? 0 : kFATile) + ((wi & 2) ? 0 : kFATile) + ((wi & 4) ? 0 : 2 * btile);
        const int chunks = conv_chunks(p), taps = p.kh * p.kw, kiters = taps * chunks;
        const int ph = (p.kh - 1) / 2, pw = (p.kw - 1) / 2;
        const int tiles_per_img = g.tiles_x * g.tiles_y;
        const int vcta = (int)((blockIdx.x + G - jb.cta_offset % G) % G);
        bool must_wait = jb.wait_prev && j > 0;
        long long* dbg = J.dbg ? J.dbg + ((size_t)j * 4096 + blockIdx.x) * 8 : nullptr;
        if (dbg && vcta < g.total_tiles) dbg[0] = gtime_ns();
        for (int tile = vcta; tile < g.total_tiles; tile += G) {
          const int mt = tile / g.n_tiles, nt = tile - mt * g.n_tiles;
          const int b = mt / tiles_per_img, trem = mt - b * tiles_per_img;
          const int ty = trem / g.tiles_x, tx = trem - ty * g.tiles_x;
          const int y0 = ty << g.bh_log2, x0 = tx << g.bw_log2, n0 = nt * bn;
          const int wb = p.w_per_batch ? b : 0;
          // k-iteration -> (channel chunk, kx, ky): chunk outermost, then kx, then ky (the order of conv_tc.cu)
          struct KIter {
            int cki, kx, ky, ck;
          };
          auto k_next = [&](KIter& k) {
            if (++k.ky == p.kh) {
              k.ky = 0;
              if (++k.kx == p.kw) { k.kx = 0; k.ck = conv_chunk(p, ++k.cki); }
            }
          };
          auto load_a = [&](const KIter& k, int s) {
            uint8_t* st = smem + s * kFStageStride;
            const int c0 = p.in_choff + k.ck * 64;
            if (!(wi & 8)) tma_load_4d(&jb.m[0], &full_bar[s], st, c0, x0 + k.kx - pw, y0 + k.ky - ph, b);
            if (!(wi & 2)) tma_load_4d(&jb.m[1], &full_bar[s], st + kFATile, c0, x0 + k.kx - pw, y0 + k.ky - ph, b);
          };
          auto load_b = [&](const KIter& k, int s) {
            uint8_t* st = smem + s * kFStageStride;
            const int kcol = (k.ky * p.kw + k.kx) * p.cin_pad + k.ck * 64;
            if (!(wi & 4)) {
              tma_load_3d(&jb.m[2], &full_bar[s], st + 2 * kFATile, kcol, n0, wb);
              tma_load_3d(&jb.m[3], &full_bar[s], st + 2 * kFATile + btile, kcol, n0, wb);
            }
          };
          KIter k = {0, 0, 0, conv_chunk(p, 0)};
          int it0 = 0;
          if (must_wait) {
            // weights first (they do not depend on earlier jobs), for as many stages as the ring has ...
            const int pre = kiters < NS ? kiters : NS;
            KIter kb = k;
            int s = rs;
            uint32_t sph = rph;
            for (int it = 0; it < pre; ++it) {
              mbar_wait(&empty_bar[s], sph ^ 1);
              mbar_arrive_expect_tx(&full_bar[s], stage_tx);
              load_b(kb, s);
              k_next(kb);
              if (++s == NS) { s = 0; sph ^= 1; }
            }
            // ... then the grid barrier: every epilogue warp of every CTA has finished the previous jobs
            const unsigned target = 16u * (unsigned)G;
            unsigned spins = 0;
            long long t0 = 0;
            while (ld_acquire_gpu(J.counters + (j - 1)) < target) {
              if ((++spins & 0x3FF) == 0) {
                long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000LL) { printf("raft_b200: fused grid barrier %d timed out (cta %d)\n", j, blockIdx.x); __trap(); }
              }
            }
            fence_proxy_async();
            if (dbg) dbg[1] = gtime_ns();
            for (int it = 0; it < pre; ++it) {
              load_a(k, rs);
              k_next(k);
              if (++rs == NS) { rs = 0; rph ^= 1; }
            }
            it0 = pre;
            must_wait = false;
          }
          for (int it = it0; it < kiters; ++it) {
            mbar_wait(&empty_bar[rs], rph ^ 1);
            mbar_arrive_expect_tx(&full_bar[rs], stage_tx);
            load_a(k, rs);
            load_b(k, rs);
            k_next(k);
            if (++rs == NS) { rs = 0; rph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int rs = 0, li = 0;
      uint32_t rph = 0;
      for (int j = 0; j < J.n; ++j) {
        const FusedJob& jb = J.job[j];
        const ConvParams& p = jb.p;
        const int bn = jb.block_n;
        const uint32_t idesc_2n = umma_idesc_f16(2 * bn), idesc_n = umma_idesc_f16(bn);
        const int kiters = p.kh * p.kw * conv_chunks(p);
        const int vcta = (int)((blockIdx.x + G - jb.cta_offset % G) % G);
        long long* dbg = J.dbg ? J.dbg + ((size_t)j * 4096 + blockIdx.x) * 8 : nullptr;
        bool first_tile = true;
        for (int tile = vcta; tile < jb.g.total_tiles; tile += G, ++li) {
          const int ab = li & 1;
          mbar_wait(&tmem_empty_bar[ab], ((li >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc = tmem_base + ab * 256;
          for (int it = 0; it < kiters; ++it) {
            const int s = rs;
            mbar_wait(&full_bar[s], rph);
            tc_fence_after();
            if (dbg && first_tile && it == 0) dbg[2] = gtime_ns();
            const uint32_t st = smem_u32(smem + s * kFStageStride);
            const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kFATile);
            const uint64_t b_all = umma_desc_sw128(st + 2 * kFATile);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t koff = (uint64_t)(k * 2);
              if (!(J.whatif & 16)) umma_f16(acc, a_hi + koff, b_all + koff, idesc_2n, (it | k) != 0);
              if (!(J.whatif & 17)) umma_f16(acc + bn, a_lo + koff, b_all + koff, idesc_n, 1u);
            }
            umma_commit(&empty_bar[s]);
            if (++rs == NS) { rs = 0; rph ^= 1; }
          }
          umma_commit(&tmem_full_bar[ab]);
          if (dbg && first_tile) dbg[3] = gtime_ns();
          first_tile = false;
        }
      }
    }
  } else {
    const int q = warp & 3, grp = (warp - 2) >> 2, r = q * 32 + lane;
    int li = 0;
    for (int j = 0; j < J.n; ++j) {
      const FusedJob& jb = J.job[j];
      const ConvParams& p = jb.p;
      const TileGeom& g = jb.g;
      const int bn = jb.block_n;
      const int cpw = bn >= 96 ? 32 : 16, groups = bn / cpw;
      const bool wide = epilogue_wide_ok(p);
      const int tiles_per_img = g.tiles_x * g.tiles_y;
      const int vcta = (int)((blockIdx.x + G - jb.cta_offset % G) % G);
      bool first = true;
      long long* dbg = (J.dbg && warp == 2 && lane == 0 && vcta < g.total_tiles) ? J.dbg + ((size_t)j * 4096 + blockIdx.x) * 8 : nullptr;
      for (int tile = vcta; tile < g.total_tiles; tile += G, ++li) {
        const int mt = tile / g.n_tiles, nt = tile - mt * g.n_tiles;
        const int b = mt / tiles_per_img, trem = mt - b * tiles_per_img;
        const int ty = trem / g.tiles_x, tx = trem - ty * g.tiles_x;
        const int y0 = ty << g.bh_log2, x0 = tx << g.bw_log2, n0 = nt * bn;
        const int py = y0 + (r >> g.bw_log2), px = x0 + (r & ((1 << g.bw_log2) - 1));
        const bool valid = (py < p.h) && (px < p.w);
        const int pix = (b * p.h + py) * p.w + px;
        const int ab = li & 1;
        mbar_wait_warp(&tmem_full_bar[ab], (li >> 1) & 1);
        tc_fence_after();
        if (dbg && first) dbg[4] = gtime_ns();
        if (first) {
          if (jb.wait_prev && j > 0) __threadfence();  // acquire side of the grid barrier for this warp's generic loads
          first = false;
        }
        if (grp < groups) {
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256;
#pragma unroll 1
          for (int cc = 0; cc < cpw; cc += 16) {
            const int c = grp * cpw + cc;
            if (n0 + c >= p.cout) break;
            uint32_t d0[16], d1[16];
            tmem_ld16(trow + c, d0);
            tmem_ld16(trow + bn + c, d1);
            tmem_ld_wait(d0, d1);
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(d0[i]) + __uint_as_float(d1[i]) * kLoInv;
            if (valid && !(J.whatif & 32)) {
              if (wide) {
                epilogue_wide16<true>(p, pix, n0 + c, v);
              } else {
                epilogue_store<8>(p, pix, n0 + c, v);
                epilogue_store<8>(p, pix, n0 + c + 8, v + 8);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[ab])) : "memory");
      }
      if (dbg) dbg[5] = gtime_ns();
      // this warp is done with job j: publish its global writes (generic proxy -> everybody, incl. other CTAs' TMA reads)
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
      if (dbg) dbg[6] = gtime_ns();
      if (lane == 0) atomicAdd(J.counters + j, 1u);
      if (dbg) dbg[7] = gtime_ns();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_fused_jobs(const FusedJobs& jobs, cudaStream_t s) {
  static PerDeviceOnce attr_set;
  int dev = 0, rc_dev;
  if ((rc_dev = current_device(&dev))) return rc_dev;
  if (!attr_set.test(dev)) {
    RB_CHECK_CUDA(cudaFuncSetAttribute(fused_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmemBytes));
    attr_set.set(dev);
  }
  const int num_sms = device_sm_count(dev);
  RB_REQUIRE(jobs.n > 0 && jobs.n <= kMaxFusedJobs && jobs.counters, RB_ERR_BAD_ARG, "fused update: bad job list");
  int max_tiles = 0;
  for (int j = 0; j < jobs.n; ++j) {
    RB_REQUIRE(jobs.job[j].block_n <= 128, RB_ERR_UNSUPPORTED, "fused update: tile width %d", jobs.job[j].block_n);
    if (jobs.job[j].g.total_tiles > max_tiles) max_tiles = jobs.job[j].g.total_tiles;
  }
  RB_CHECK_CUDA(cudaMemsetAsync(jobs.counters, 0, kMaxFusedJobs * sizeof(unsigned), s));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(max_tiles < num_sms ? max_tiles : num_sms);
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = kFSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the in-kernel grid barriers cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fused_conv_kernel, jobs));
  RB_CHECK_LAUNCH("fused_conv_kernel");
  return RB_OK;
}

}  // namespace rb
