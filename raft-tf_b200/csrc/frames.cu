// F3 (input edge): frame preprocessing on the GPU.
// Reference: dataflow/test_dataflow.py:56,61 (BGR decode), :96-97 (np.float32(x) / 255.0); the 2x-1 of
// RAFT.input_preprocess (networks/RAFT.py:53-59) is applied by the encoder stem (encoder.cu).  H/W that are not multiples of
// 8 are replicate-padded here (SURVEY 8(d) shape policy: upstream InputPadder, pad split between both sides) -- one pass,
// no torch ops on the timed path.
#include "common.cuh"

namespace rb {

// dst [B,Hp,Wp,3] fp32 in [0,1]; src [B,H,W,3] fp32 in [0,1] (U8 = false) or uint8 in [0,255] (U8 = true).
// One thread per destination pixel; the source pixel is the clamped (replicate) coordinate.
template <bool U8>
__global__ void frames_prepare_kernel(const void* __restrict__ src, float* __restrict__ dst, int B, int H, int W, int Hp,
                                      int Wp, int pad_top, int pad_left) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * Hp * Wp) return;
  const int x = (int)(i % Wp), y = (int)((i / Wp) % Hp), b = (int)(i / ((size_t)Wp * Hp));
  const int sx = min(max(x - pad_left, 0), W - 1), sy = min(max(y - pad_top, 0), H - 1);
  const size_t so = (((size_t)b * H + sy) * W + sx) * 3;
  float v0, v1, v2;
  if (U8) {
    const uint8_t* s = reinterpret_cast<const uint8_t*>(src) + so;
    v0 = __fdiv_rn((float)s[0], 255.0f);  // np.float32(x) / 255.0: correctly rounded fp32 division
    v1 = __fdiv_rn((float)s[1], 255.0f);
    v2 = __fdiv_rn((float)s[2], 255.0f);
  } else {
    const float* s = reinterpret_cast<const float*>(src) + so;
    v0 = s[0]; v1 = s[1]; v2 = s[2];
  }
  float* d = dst + i * 3;
  d[0] = v0; d[1] = v1; d[2] = v2;
}

}  // namespace rb

using namespace rb;

extern "C" int rb_frames_prepare(const void* src, int src_is_u8, float* dst, int B, int H, int W, int pad_top,
                                 int pad_bottom, int pad_left, int pad_right, void* stream) {
  RB_REQUIRE(src && dst, RB_ERR_BAD_ARG, "rb_frames_prepare: null pointer");
  RB_REQUIRE(B > 0 && H > 0 && W > 0 && pad_top >= 0 && pad_bottom >= 0 && pad_left >= 0 && pad_right >= 0, RB_ERR_BAD_SHAPE,
             "rb_frames_prepare: bad shape B=%d H=%d W=%d pad=(%d,%d,%d,%d)", B, H, W, pad_top, pad_bottom, pad_left, pad_right);
  const int Hp = H + pad_top + pad_bottom, Wp = W + pad_left + pad_right;
  const size_t n = (size_t)B * Hp * Wp;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (src_is_u8)
    frames_prepare_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, B, H, W, Hp, Wp, pad_top, pad_left);
  else
    frames_prepare_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, B, H, W, Hp, Wp, pad_top, pad_left);
  RB_CHECK_LAUNCH("frames_prepare_kernel");
  return RB_OK;
}
