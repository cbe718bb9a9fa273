// A1 on tensor cores: the all-pairs correlation volume and its pyramid as four GEMMs.
// Reference: networks/model_utils.py:199-221 (GetCorrPyramid).
//
// level 0:  vol0[n, m] = <fmap1[n,:], fmap2[m,:]> / sqrt(C)
// level l:  the reference average-pools the h2 x w2 axes of the volume (2x2, stride 2, VALID).
//           Pooling is linear, so  pool(vol)[n, m'] = <fmap1[n,:], pool(fmap2)[m',:]> / sqrt(C):
//           instead of re-reading the 4*N^2-byte volume three times we pool the (tiny) fmap2 and
//           run three more GEMMs (+33% MMA work, no extra HBM reads).  Identical in exact
//           arithmetic; differs from pooling the stored fp32 volume by fp32 rounding only
//           (covered by tests/test_corr_parity.py).
// The GEMMs reuse the implicit-GEMM kernel of conv_tc.cu with a 1x1 "filter" whose weight operand
// is the (pooled) fmap2 of the same batch element.
#include "common.cuh"

namespace rb {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct CorrWs {
  __half* f1_hi;
  __half* f1_lo;
  __half* f2_hi[RB_NUM_LEVELS];
  __half* f2_lo[RB_NUM_LEVELS];
  float* pooled[RB_NUM_LEVELS];  // fp32 pooled fmap2, levels 1..3 (level 0 aliases the input)
  size_t total;
};

static CorrWs corr_ws_layout(int B, int h, int w, int C, void* base) {
  CorrWs W;
  char* b = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = b + off; off += align_up(bytes, 1024); return p; };
  size_t n0 = (size_t)B * h * w * C;
  W.f1_hi = reinterpret_cast<__half*>(take(n0 * 2));
  W.f1_lo = reinterpret_cast<__half*>(take(n0 * 2));
  for (int l = 0; l < RB_NUM_LEVELS; ++l) {
    // rows padded to a multiple of 16 so the packed "weight" operand has a legal cout_pad
    size_t rows = (size_t)((level_dim(h, l) * level_dim(w, l) + 15) / 16 * 16);
    size_t nl = (size_t)B * rows * C;
    W.f2_hi[l] = reinterpret_cast<__half*>(take(nl * 2));
    W.f2_lo[l] = reinterpret_cast<__half*>(take(nl * 2));
    W.pooled[l] = l == 0 ? nullptr : reinterpret_cast<float*>(take((size_t)B * level_dim(h, l) * level_dim(w, l) * C * 4));
  }
  W.total = off;
  return W;
}

size_t corr_tc_workspace_bytes(int B, int h, int w, int C) { return corr_ws_layout(B, h, w, C, nullptr).total; }

// fp32 [B, rows, C] -> split planes [B, rows_pad, C] (rows beyond `rows` are left zero by the caller's memset... they
// are never read into valid outputs: the epilogue masks columns >= cout)
__global__ void split_rows_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo,
                                  int B, int rows, int rows_pad, int C) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per = (size_t)rows_pad * C;
  if (i >= (size_t)B * per) return;
  int b = i / per;
  size_t rem = i - (size_t)b * per;
  int r = rem / C, c = rem % C;
  float v = r < rows ? src[((size_t)b * rows + r) * C + c] : 0.f;
  __half hh, ll;
  split_f32(v, hh, ll);
  hi[i] = hh;
  lo[i] = ll;
}

// 2x2 VALID average pool of an NHWC fp32 feature map
__global__ void pool_fmap_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int hs, int ws, int hd,
                                 int wd, int C) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * hd * wd * C) return;
  int c = i % C;
  size_t t = i / C;
  int x = t % wd; t /= wd;
  int y = t % hd;
  int b = t / hd;
  const float* s = src + (((size_t)b * hs + 2 * y) * ws + 2 * x) * C + c;
  dst[i] = (s[0] + s[C] + s[(size_t)ws * C] + s[(size_t)ws * C + C]) * 0.25f;
}

int corr_build_tc(const float* fmap1, const float* fmap2, float* pyramid, int B, int h, int w, int C, void* ws,
                  size_t ws_bytes, cudaStream_t s) {
  RB_REQUIRE(C % 64 == 0, RB_ERR_BAD_SHAPE, "rb_corr_build (tensor-core): C=%d must be a multiple of 64", C);
  size_t need = corr_tc_workspace_bytes(B, h, w, C);
  RB_REQUIRE(ws && ws_bytes >= need, RB_ERR_WORKSPACE, "rb_corr_build: workspace has %zu bytes, need %zu", ws_bytes, need);
  CorrWs W = corr_ws_layout(B, h, w, C, ws);
  const int N = h * w;
  {
    size_t n = (size_t)B * N * C;
    split_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(fmap1, W.f1_hi, W.f1_lo, B, N, N, C);
    RB_CHECK_LAUNCH("split_rows_kernel");
  }
  const float* prev = fmap2;
  size_t lvl_off = 0;
  for (int l = 0; l < RB_NUM_LEVELS; ++l) {
    const int hl = level_dim(h, l), wl = level_dim(w, l);
    const int rows = hl * wl, rows_pad = (rows + 15) / 16 * 16;
    const float* cur = prev;
    if (l > 0) {
      size_t n = (size_t)B * rows * C;
      pool_fmap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(prev, W.pooled[l], B, level_dim(h, l - 1),
                                                                   level_dim(w, l - 1), hl, wl, C);
      RB_CHECK_LAUNCH("pool_fmap_kernel");
      cur = W.pooled[l];
    }
    {
      size_t n = (size_t)B * rows_pad * C;
      split_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cur, W.f2_hi[l], W.f2_lo[l], B, rows, rows_pad, C);
      RB_CHECK_LAUNCH("split_rows_kernel");
    }
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.in_hi = W.f1_hi; p.in_lo = W.f1_lo; p.in_stride = C; p.in_choff = 0; p.cin_pad = C;
    p.w_hi = W.f2_hi[l]; p.w_lo = W.f2_lo[l]; p.bias = nullptr;
    p.cout = rows; p.cout_pad = rows_pad; p.kh = 1; p.kw = 1; p.w_per_batch = 1;
    p.B = B; p.h = h; p.w = w;
    p.epi = EPI_F32; p.scale = 1.f; p.div = sqrtf((float)C);  // divide after the matmul (:213)
    p.f0 = pyramid + lvl_off;
    int rc = launch_conv_tc(p, s);
    if (rc) return rc;
    lvl_off += (size_t)B * N * rows;
    prev = cur;
  }
  return RB_OK;
}

}  // namespace rb
