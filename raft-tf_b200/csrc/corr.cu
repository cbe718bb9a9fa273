// A1 (correlation volume + pyramid), A2/A3 (pyramid lookup), A4 (coords grid).
// Reference: networks/model_utils.py:199-249, networks/utils.py:4-103.
#include "common.cuh"

namespace rb {

// ---------------------------------------------------------------------------------------------
// A4  coords_grid  (utils.py:4-11)
// ---------------------------------------------------------------------------------------------
__global__ void coords_grid_kernel(float2* __restrict__ coords, int B, int h, int w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = B * h * w;
  if (i >= n) return;
  int x = i % w, y = (i / w) % h;
  coords[i] = make_float2((float)x, (float)y);
}

// ---------------------------------------------------------------------------------------------
// A1  SIMT fp32 correlation GEMM: vol[b, n, m] = <f1[b,n,:], f2[b,m,:]> / sqrt(C)
//     (model_utils.py:206-215).  64x64 tile, 16-deep k slices, 4x4 micro-tile per thread.
//     Bring-up / cross-check back end (RB_MATH_SIMT); the tcgen05 build lives in gemm_tc.cu.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) corr_gemm_simt_kernel(const float* __restrict__ f1,
                                                             const float* __restrict__ f2,
                                                             float* __restrict__ vol, int N, int C,
                                                             float inv_sqrt_c_is_unused) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int b = blockIdx.z;
  const float* A = f1 + (size_t)b * N * C;
  const float* Bm = f2 + (size_t)b * N * C;
  float* out = vol + (size_t)b * N * N;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < C; k0 += 16) {
    {
      int r = tid / 4, kq = (tid % 4) * 4;
      float4 va = make_float4(0, 0, 0, 0), vb = va;
      if (m0 + r < N) va = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * C + k0 + kq);
      if (n0 + r < N) vb = *reinterpret_cast<const float4*>(Bm + (size_t)(n0 + r) * C + k0 + kq);
      As[kq + 0][r] = va.x; As[kq + 1][r] = va.y; As[kq + 2][r] = va.z; As[kq + 3][r] = va.w;
      Bs[kq + 0][r] = vb.x; Bs[kq + 1][r] = vb.y; Bs[kq + 2][r] = vb.z; Bs[kq + 3][r] = vb.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; bb[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float sq = sqrtf((float)C);
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= N) continue;
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) out[(size_t)m * N + n] = acc[i][j] / sq;  // divide AFTER the matmul (:213)
    }
  }
}

// 2x2 average pooling, stride 2, VALID (floor) -- tensorpack AvgPooling (model_utils.py:217-219).
__global__ void avgpool2_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t rows,
                                int hs, int ws, int hd, int wd) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = rows * hd * wd;
  if (i >= total) return;
  int x = i % wd;
  int y = (i / wd) % hd;
  size_t r = i / ((size_t)wd * hd);
  const float* s = src + r * hs * ws + (size_t)(2 * y) * ws + 2 * x;
  dst[i] = (s[0] + s[1] + s[ws] + s[ws + 1]) * 0.25f;
}

// ---------------------------------------------------------------------------------------------
// A2/A3  pyramid lookup (model_utils.py:224-249 + utils.py:39-103).
//
// A block handles PB = 8 query pixels x 4 levels = 32 "units"; thread = (unit, window row j).
//   phase 1  the block stages, per unit, the (2r+3) rows x 16 columns of the level that cover every
//            tap footprint ((2r+1)^2 taps share one fractional offset, +1 row/col of slack for the
//            fp32 rounding of cx+dx): 16-byte loads from a 4-float-aligned column, row index clamped.
//   phase 2  thread j of a unit computes the x-side of window column i=j (clamped x0/x1 relative to
//            the staged columns, qx = x1c - x and 1 - qx) into a small shared table, and keeps the
//            y-side of its own row j in registers: the separable index math is done once per row /
//            column instead of once per tap.
//   phase 3  each thread walks the (2r+1) taps of its row with the reference's exact arithmetic
//            (trunc toward zero, clamp, weights from the CLAMPED x1/y1, add_n order; __fmul_rn /
//            __fadd_rn forbid FMA contraction => bit-identical to the fp32 CPU oracle) and stores
//            channel lvl*K + i*(2r+1) + j: the threads of a unit write consecutive addresses.
// Output: fp32 [pix][4K] (rb_corr_lookup) or hi/lo fp16 planes with a padded channel stride (the
// layout convc1 consumes).  r01 profile of the previous one-warp-per-unit kernel: 391 warp
// instructions per unit, issue-bound (70 % issue slots, 22 % DRAM); this formulation needs ~110.
// ---------------------------------------------------------------------------------------------
struct PyramidView {
  const float* base[RB_NUM_LEVELS];
  int hl[RB_NUM_LEVELS], wl[RB_NUM_LEVELS];
  int vec_ok[RB_NUM_LEVELS];  // level rows are 16-byte aligned (W % 4 == 0 and aligned base)
};

constexpr int kLookupPB = 8;    // pixels per block
constexpr int kPatchCols = 16;  // staged columns per row
constexpr int kPatchPitch = 20; // floats per staged row (16 + 4 padding: conflict-light column reads)

template <int R, bool SPLIT>
__global__ void __launch_bounds__(kLookupPB * 4 * (2 * R + 1))
corr_lookup_kernel(const __grid_constant__ PyramidView pv, const float2* __restrict__ coords,
                   float* __restrict__ out_f32, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                   int out_stride, int npix) {
  constexpr int D = 2 * R + 1, K = D * D, P = D + 2, UNITS = kLookupPB * 4, NT = UNITS * D;
  __shared__ __align__(16) float patch[UNITS][P][kPatchPitch];
  __shared__ __align__(16) float4 xtab[UNITS][D];
  __shared__ int ubase[UNITS][2];  // bx4, by per unit
  const int tid = threadIdx.x;
  const int pix0 = blockIdx.x * kLookupPB;

  // ---- phase 0: per-unit origin -------------------------------------------------------------
  if (tid < UNITS) {
    const int u = tid, pl = u >> 2, lvl = u & 3;
    const int pix = min(pix0 + pl, npix - 1);
    const float2 c = __ldg(coords + pix);
    const float inv = 1.0f / (float)(1 << lvl);  // centroid / 2**i (model_utils.py:239), exact
    const int H = pv.hl[lvl], W = pv.wl[lvl];
    const float xf0 = __fadd_rn(c.x * inv, (float)(-R)), yf0 = __fadd_rn(c.y * inv, (float)(-R));
    const int bx = min(max((int)xf0, 0), W - 1);
    ubase[u][0] = bx & ~3;
    ubase[u][1] = min(max((int)yf0, 0), H - 1);
  }
  __syncthreads();
  // ---- phase 1: stage patches (UNITS * P rows * 4 chunks of 4 floats) -------------------------
  for (int e = tid; e < UNITS * P * 4; e += NT) {
    const int u = e / (P * 4), rem = e - u * (P * 4), py = rem >> 2, ch = rem & 3;
    const int pl = u >> 2, lvl = u & 3;
    const int pix = min(pix0 + pl, npix - 1);
    const int H = pv.hl[lvl], W = pv.wl[lvl];
    const int yy = min(ubase[u][1] + py, H - 1);
    const int col = ubase[u][0] + ch * 4;
    const float* row = pv.base[lvl] + ((size_t)pix * H + yy) * W;
    float4 v;
    if (pv.vec_ok[lvl]) {
      v = __ldg(reinterpret_cast<const float4*>(row + col));  // may run past the row end: never indexed
    } else {
      v.x = __ldg(row + min(col + 0, W - 1)); v.y = __ldg(row + min(col + 1, W - 1));
      v.z = __ldg(row + min(col + 2, W - 1)); v.w = __ldg(row + min(col + 3, W - 1));
    }
    *reinterpret_cast<float4*>(&patch[u][py][ch * 4]) = v;
  }
  // ---- phase 2: separable index math ------------------------------------------------------------
  const int u = tid / D, j = tid - u * D;
  const int pl = u >> 2, lvl = u & 3;
  const int pix = pix0 + pl;
  const int H = pv.hl[lvl], W = pv.wl[lvl];
  const float2 c = __ldg(coords + min(pix, npix - 1));
  const float inv = 1.0f / (float)(1 << lvl);
  const float cx = c.x * inv, cy = c.y * inv;
  {
    // x side of window column i = j (i walks x: model_utils.py:235-237)
    const float x = __fadd_rn(cx, (float)(j - R));
    int x0 = (int)x;  // tf.cast truncates toward zero (utils.py:54-57)
    int x1 = x0 + 1;
    x0 = min(max(x0, 0), W - 1);
    x1 = min(max(x1, 0), W - 1);
    const float qx = __fsub_rn((float)x1, x);  // utils.py:84 (clamped x1)
    const int bx4 = ubase[u][0];
    const int ax0 = min(max(x0 - bx4, 0), kPatchCols - 1), ax1 = min(max(x1 - bx4, 0), kPatchCols - 1);
    xtab[u][j] = make_float4(qx, __fsub_rn(1.0f, qx), __int_as_float(ax0), __int_as_float(ax1));
  }
  const float y = __fadd_rn(cy, (float)(j - R));
  int y0 = (int)y;
  int y1 = y0 + 1;
  y0 = min(max(y0, 0), H - 1);
  y1 = min(max(y1, 0), H - 1);
  const float qy = __fsub_rn((float)y1, y), pyw = __fsub_rn(1.0f, qy);  // utils.py:85
  const int by = ubase[u][1];
  const float* row0 = &patch[u][min(max(y0 - by, 0), P - 1)][0];
  const float* row1 = &patch[u][min(max(y1 - by, 0), P - 1)][0];
  __syncthreads();
  if (pix >= npix) return;
  // ---- phase 3: taps of window row j ----------------------------------------------------------------
  const size_t obase = (size_t)pix * out_stride + lvl * K + j;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const float4 xt = xtab[u][i];
    const int ax0 = __float_as_int(xt.z), ax1 = __float_as_int(xt.w);
    const float wa = __fmul_rn(xt.x, qy), wb = __fmul_rn(xt.x, pyw);  // utils.py:86-89
    const float wc = __fmul_rn(xt.y, qy), wd = __fmul_rn(xt.y, pyw);
    const float Ia = row0[ax0], Ib = row1[ax0], Ic = row0[ax1], Id = row1[ax1];
    const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wa, Ia), __fmul_rn(wb, Ib)), __fmul_rn(wc, Ic)),
                              __fmul_rn(wd, Id));  // tf.add_n order (utils.py:98)
    if constexpr (SPLIT) {
      __half hi, lo;
      split_f32(v, hi, lo);
      out_hi[obase + i * D] = hi;
      out_lo[obase + i * D] = lo;
    } else {
      out_f32[obase + i * D] = v;
    }
  }
}

// General form of bilinear_sampler / tf_grid_sample (utils.py:39-103) for single-channel images:
// img [n,H,W,1], coords [n,S,2] -> out [n,S].  One thread per sample; same arithmetic as above.
__global__ void bilinear_sample_kernel(const float* __restrict__ img, const float2* __restrict__ coords,
                                       float* __restrict__ out, int n, int H, int W, int S) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * S) return;
  const float* im = img + (i / S) * (size_t)H * W;
  const float2 c = coords[i];
  int x0 = (int)c.x, y0 = (int)c.y;
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1);
  y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
  const float qx = __fsub_rn((float)x1, c.x), qy = __fsub_rn((float)y1, c.y);
  const float pxw = __fsub_rn(1.0f, qx), pyw = __fsub_rn(1.0f, qy);
  const float Ia = im[(size_t)y0 * W + x0], Ib = im[(size_t)y1 * W + x0];
  const float Ic = im[(size_t)y0 * W + x1], Id = im[(size_t)y1 * W + x1];
  out[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(qx, qy), Ia), __fmul_rn(__fmul_rn(qx, pyw), Ib)),
                               __fmul_rn(__fmul_rn(pxw, qy), Ic)),
                     __fmul_rn(__fmul_rn(pxw, pyw), Id));
}

int pyramid_view(const float* pyramid, int B, int h, int w, PyramidView* pv) {
  size_t off = 0;
  size_t rows = (size_t)B * h * w;
  for (int l = 0; l < RB_NUM_LEVELS; ++l) {
    int hl = level_dim(h, l), wl = level_dim(w, l);
    if (hl < 1 || wl < 1) {
      set_error("pyramid level %d is empty for a %dx%d grid", l, h, w);
      return RB_ERR_BAD_SHAPE;
    }
    pv->base[l] = pyramid + off;
    pv->hl[l] = hl;
    pv->wl[l] = wl;
    pv->vec_ok[l] = (wl % 4 == 0) && (reinterpret_cast<uintptr_t>(pyramid + off) % 16 == 0);
    off += rows * hl * wl;
  }
  return RB_OK;
}

int launch_lookup(const float* pyramid, const float* coords, float* out_f32, __half* out_hi,
                  __half* out_lo, int out_stride, int B, int h, int w, int radius, cudaStream_t s) {
  PyramidView pv;
  int rc = pyramid_view(pyramid, B, h, w, &pv);
  if (rc) return rc;
  int npix = B * h * w;
  dim3 grid((npix + kLookupPB - 1) / kLookupPB);
  const float2* c2 = reinterpret_cast<const float2*>(coords);
  bool split = out_hi != nullptr;
  if (radius == 4) {
    dim3 block(kLookupPB * 4 * 9);
    if (split) corr_lookup_kernel<4, true><<<grid, block, 0, s>>>(pv, c2, nullptr, out_hi, out_lo, out_stride, npix);
    else corr_lookup_kernel<4, false><<<grid, block, 0, s>>>(pv, c2, out_f32, nullptr, nullptr, out_stride, npix);
  } else if (radius == 3) {
    dim3 block(kLookupPB * 4 * 7);
    if (split) corr_lookup_kernel<3, true><<<grid, block, 0, s>>>(pv, c2, nullptr, out_hi, out_lo, out_stride, npix);
    else corr_lookup_kernel<3, false><<<grid, block, 0, s>>>(pv, c2, out_f32, nullptr, nullptr, out_stride, npix);
  } else {
    set_error("radius %d unsupported (3 = raft-small, 4 = raft-things)", radius);
    return RB_ERR_UNSUPPORTED;
  }
  RB_CHECK_LAUNCH("corr_lookup_kernel");
  return RB_OK;
}

int corr_build_tc(const float* fmap1, const float* fmap2, float* pyramid, int B, int h, int w, int C,
                  void* ws, size_t ws_bytes, cudaStream_t s);
size_t corr_tc_workspace_bytes(int B, int h, int w, int C);

}  // namespace rb

using namespace rb;

extern "C" int rb_coords_grid(float* coords, int B, int h, int w, void* stream) {
  RB_REQUIRE(coords && B > 0 && h > 0 && w > 0, RB_ERR_BAD_ARG, "rb_coords_grid: bad argument");
  int n = B * h * w;
  coords_grid_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(coords), B, h, w);
  RB_CHECK_LAUNCH("coords_grid_kernel");
  return RB_OK;
}

extern "C" int rb_corr_level_offset(int B, int h, int w, int level, size_t* offset_floats, int* hl, int* wl) {
  RB_REQUIRE(B > 0 && h > 0 && w > 0 && level >= 0 && level <= RB_NUM_LEVELS, RB_ERR_BAD_ARG,
             "rb_corr_level_offset: bad argument");
  size_t off = 0, rows = (size_t)B * h * w;
  for (int l = 0; l < level; ++l) off += rows * level_dim(h, l) * level_dim(w, l);
  if (offset_floats) *offset_floats = off;
  if (hl) *hl = level < RB_NUM_LEVELS ? level_dim(h, level) : 0;
  if (wl) *wl = level < RB_NUM_LEVELS ? level_dim(w, level) : 0;
  return RB_OK;
}

extern "C" int rb_corr_pyramid_bytes(int B, int h, int w, size_t* bytes) {
  RB_REQUIRE(bytes, RB_ERR_BAD_ARG, "rb_corr_pyramid_bytes: null output");
  RB_REQUIRE(B > 0 && (h >> 3) >= 1 && (w >> 3) >= 1, RB_ERR_BAD_SHAPE,
             "rb_corr_pyramid_bytes: grid %dx%d too small for 4 levels", h, w);
  size_t off;
  rb_corr_level_offset(B, h, w, RB_NUM_LEVELS, &off, nullptr, nullptr);
  *bytes = off * sizeof(float) + 256;  // tail padding: the lookup stages 16-column row segments with vector loads
  return RB_OK;
}

extern "C" int rb_corr_workspace_bytes(int B, int h, int w, int C, size_t* bytes) {
  RB_REQUIRE(bytes && B > 0 && h > 0 && w > 0 && C > 0, RB_ERR_BAD_ARG, "rb_corr_workspace_bytes: bad argument");
  *bytes = corr_tc_workspace_bytes(B, h, w, C);
  return RB_OK;
}

extern "C" int rb_corr_build(const float* fmap1, const float* fmap2, float* pyramid, int B, int h, int w,
                             int C, void* workspace, size_t workspace_bytes, void* stream) {
  RB_REQUIRE(fmap1 && fmap2 && pyramid, RB_ERR_BAD_ARG, "rb_corr_build: null pointer");
  RB_REQUIRE(B > 0 && (h >> 3) >= 1 && (w >> 3) >= 1, RB_ERR_BAD_SHAPE, "rb_corr_build: grid %dx%d too small", h, w);
  RB_REQUIRE(C > 0 && C % 16 == 0, RB_ERR_BAD_SHAPE, "rb_corr_build: C=%d must be a multiple of 16", C);
  cudaStream_t s = (cudaStream_t)stream;
  if (math_mode() == RB_MATH_TC) return corr_build_tc(fmap1, fmap2, pyramid, B, h, w, C, workspace, workspace_bytes, s);
  const int N = h * w;
  dim3 grid((N + 63) / 64, (N + 63) / 64, B);
  corr_gemm_simt_kernel<<<grid, 256, 0, s>>>(fmap1, fmap2, pyramid, N, C, 0.f);
  RB_CHECK_LAUNCH("corr_gemm_simt_kernel");
  size_t rows = (size_t)B * N;
  for (int l = 0; l + 1 < RB_NUM_LEVELS; ++l) {
    size_t so, dofs;
    int hs, ws, hd, wd;
    rb_corr_level_offset(B, h, w, l, &so, &hs, &ws);
    rb_corr_level_offset(B, h, w, l + 1, &dofs, &hd, &wd);
    size_t total = rows * hd * wd;
    avgpool2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(pyramid + so, pyramid + dofs, rows, hs, ws, hd, wd);
    RB_CHECK_LAUNCH("avgpool2_kernel");
  }
  return RB_OK;
}

extern "C" int rb_bilinear_sample(const float* img, const float* coords, float* out, int n, int H, int W, int S,
                                  void* stream) {
  RB_REQUIRE(img && coords && out, RB_ERR_BAD_ARG, "rb_bilinear_sample: null pointer");
  RB_REQUIRE(n > 0 && H > 0 && W > 0 && S > 0, RB_ERR_BAD_SHAPE, "rb_bilinear_sample: bad shape");
  size_t total = (size_t)n * S;
  bilinear_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      img, reinterpret_cast<const float2*>(coords), out, n, H, W, S);
  RB_CHECK_LAUNCH("bilinear_sample_kernel");
  return RB_OK;
}

extern "C" int rb_corr_lookup(const float* pyramid, const float* coords, float* out, int B, int h, int w,
                              int radius, void* stream) {
  RB_REQUIRE(pyramid && coords && out, RB_ERR_BAD_ARG, "rb_corr_lookup: null pointer");
  RB_REQUIRE(B > 0 && h > 0 && w > 0, RB_ERR_BAD_SHAPE, "rb_corr_lookup: bad shape");
  int K = (2 * radius + 1) * (2 * radius + 1);
  return launch_lookup(pyramid, coords, out, nullptr, nullptr, 4 * K, B, h, w, radius, (cudaStream_t)stream);
}
