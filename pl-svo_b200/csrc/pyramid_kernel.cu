// pyramid_kernel.cu — frame_utils::createImgPyramid (src/frame.cpp:171-180) = repeated vk::halfSample
// (rpg_vikit vision.cpp, scalar path: truncating mean of each 2x2 block) for a batch of frames.
// SURVEY.md §8f rank 2 ("next"): the producer of the alignment path's image input.
//
// One CTA owns a 64x64 tile of level 0 and produces the corresponding tiles of every level from
// shared memory, so level 0 is read from HBM exactly once and each level is written once
// (traffic = 4/3 x the level-0 bytes; HBM-bound byte work, no tensor cores).  Tiles are aligned to
// 64 pixels, so the truncating 2x2 means are identical to the level-by-level computation.
#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"

namespace plsvo {
namespace {

constexpr int kTile = 64;

__global__ void __launch_bounds__(256) pyramid_kernel(const PyramidArgs a) {
  __shared__ __align__(16) uint8_t t0[kTile * kTile];       // level 0 tile
  __shared__ uint8_t t1[(kTile / 2) * (kTile / 2)];         // level 1 tile, then reused downwards
  __shared__ uint8_t t2[(kTile / 4) * (kTile / 4)];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const uint8_t* src = a.level[0] + (size_t)b * a.stride[0];
  const int tid = threadIdx.x;
  // load: 64 rows x 64 bytes, 16 bytes per thread (rows are 16B-pitched on the device)
  {
    const int row = tid >> 2, cx = (tid & 3) * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y0 + row < a.height && x0 + cx < (int)a.pitch[0])
      v = *reinterpret_cast<const uint4*>(src + (size_t)(y0 + row) * a.pitch[0] + x0 + cx);
    *reinterpret_cast<uint4*>(t0 + row * kTile + cx) = v;
  }
  __syncthreads();
  const uint8_t* in = t0;
  int in_dim = kTile;
  for (int l = 1; l < a.n_levels; ++l) {
    const int out_dim = in_dim >> 1;
    if (out_dim == 0) break;
    uint8_t* out = (l & 1) ? t1 : t2;
    const int W = a.width >> l, H = a.height >> l;
    const int ox0 = x0 >> l, oy0 = y0 >> l;
    uint8_t* dst = a.level[l] + (size_t)b * a.stride[l];
    for (int i = tid; i < out_dim * out_dim; i += 256) {
      const int oy = i / out_dim, ox = i - oy * out_dim;
      const uint8_t* p = in + (2 * oy) * in_dim + 2 * ox;
      const int s = (int)p[0] + (int)p[1] + (int)p[in_dim] + (int)p[in_dim + 1];
      const uint8_t v = (uint8_t)(s / 4);  // truncating, as the scalar vk::halfSample
      out[i] = v;
      if (ox0 + ox < W && oy0 + oy < H) dst[(size_t)(oy0 + oy) * a.pitch[l] + ox0 + ox] = v;
    }
    __syncthreads();
    in = out;
    in_dim = out_dim;
  }
}

}  // namespace

cudaError_t pyramid_kernel_launch(const PyramidArgs& a, cudaStream_t s) {
  dim3 grid((a.width + kTile - 1) / kTile, (a.height + kTile - 1) / kTile, a.B);
  pyramid_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace plsvo
