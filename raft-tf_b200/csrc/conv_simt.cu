// CUDA-core (fp32 FMA) implicit-GEMM convolution over split fp16 operands.
// Bring-up and cross-check back end (RB_MATH_SIMT): same buffers, packed weights and fused
// epilogues as the tcgen05 kernel in conv_tc.cu, so either can run any conv of the update block.
// Implements tensorpack Conv2D(stride 1, 'same') = zero padding (k-1)/2  (SURVEY A14), and the strided / windowed input
// views of the encoders (ConvParams, common.cuh).
#include "common.cuh"

namespace rb {

// tile: 64 pixels x 64 output channels, 16 input channels per step; 256 threads, 4x4 each.
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int npix = p.B * p.h * p.w;
  const int pix0 = blockIdx.x * 64, co0 = blockIdx.y * 64;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int lr = tid / 4, lk = (tid % 4) * 4;  // loader: row (pixel / cout) and 4-channel group
  const int ph = conv_pad_y(p), pw = conv_pad_x(p), csx = conv_sx(p), csy = conv_sy(p);
  const int iw = conv_in_w(p), ih = conv_in_h(p), rowpitch = conv_rowpitch(p);  // input view (common.cuh)
  const int taps = p.kh * p.kw;
  // loader pixel coordinates
  const int lpix = pix0 + lr;
  int lb = 0, ly = 0, lx = 0;
  if (lpix < npix) { lx = lpix % p.w; ly = (lpix / p.w) % p.h; lb = lpix / (p.w * p.h); }
  float acc[4][4] = {};
  for (int t = 0; t < taps; ++t) {
    const int sy = ly * csy + t / p.kw - ph, sx = lx * csx + t % p.kw - pw;
    const bool in_img = (lpix < npix) && sy >= 0 && sy < ih && sx >= 0 && sx < iw;
    const size_t src = ((size_t)lb * ih + sy) * rowpitch + (size_t)sx * p.in_stride + p.in_choff;
    const size_t wrow = ((size_t)(co0 + lr) * taps + t) * p.cin_pad;
    const bool w_ok = (co0 + lr) < p.cout_pad;
    for (int kk = 0; kk < conv_chunks(p) * 64; kk += 16) {
      const int k0 = conv_chunk(p, kk >> 6) * 64 + (kk & 63);
      float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
      if (in_img) {
        uint2 h = *reinterpret_cast<const uint2*>(p.in_hi + src + k0 + lk);
        uint2 l = *reinterpret_cast<const uint2*>(p.in_lo + src + k0 + lk);
        const __half* hh = reinterpret_cast<const __half*>(&h);
        const __half* ll = reinterpret_cast<const __half*>(&l);
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = join_f32(hh[i], ll[i]);
      }
      if (w_ok) {
        uint2 h = *reinterpret_cast<const uint2*>(p.w_hi + wrow + k0 + lk);
        uint2 l = *reinterpret_cast<const uint2*>(p.w_lo + wrow + k0 + lk);
        const __half* hh = reinterpret_cast<const __half*>(&h);
        const __half* ll = reinterpret_cast<const __half*>(&l);
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = join_f32(hh[i], ll[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { As[lk + i][lr] = a[i]; Bs[lk + i][lr] = b[i]; }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { av[i] = As[k][ty * 4 + i]; bv[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int pix = pix0 + ty * 4 + i;
    if (pix < npix) epilogue_store<4>(p, pix, co0 + tx * 4, acc[i]);
  }
}

int launch_conv_simt(const ConvParams& p, cudaStream_t s) {
  const int npix = p.B * p.h * p.w;
  dim3 grid((npix + 63) / 64, (p.cout + 63) / 64);
  conv_simt_kernel<<<grid, 256, 0, s>>>(p);
  RB_CHECK_LAUNCH("conv_simt_kernel");
  return RB_OK;
}

}  // namespace rb
