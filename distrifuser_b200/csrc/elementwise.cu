// distrifuser_b200 -- fused GEGLU gate for the transformer feed-forward (diffusers FeedForward.net[0], untouched by the
// reference's wrappers but 13 % of a 1024^2 SDXL step as two eager torch kernels: gelu(gate) then hidden * gelu).
// One pass: reads the [rows, 2*cols] projection once, writes [rows, cols].   Bound: HBM (3 * rows * cols * 2 B).
#include "common.cuh"

using namespace df;

namespace {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(256) geglu_kernel(const __half* __restrict__ in, __half* __restrict__ out, int64_t rows,
                                                    int vec_per_row, int64_t in_pitch, int64_t out_pitch, int cols) {
  const int64_t total = rows * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int q = (int)(i - r * vec_per_row);
    const __half* src = in + r * in_pitch + (int64_t)q * 8;
    int4 hv = ld_nc_v4(src), gv = ld_nc_v4(src + cols);
    const __half2* h2 = reinterpret_cast<const __half2*>(&hv);
    const __half2* g2 = reinterpret_cast<const __half2*>(&gv);
    int4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 h = __half22float2(h2[j]), g = __half22float2(g2[j]);
      o2[j] = __floats2half2_rn(h.x * gelu_erf(g.x), h.y * gelu_erf(g.y));
    }
    st_v4(out + r * out_pitch + (int64_t)q * 8, o);
  }
}

}  // namespace

extern "C" int df_geglu(const void* in, void* out, int64_t rows, int cols, int64_t in_pitch, int64_t out_pitch, void* stream) {
  DF_REQUIRE(cols % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0,
             "df_geglu: 16-byte alignment required (cols=%d)", cols);
  if (rows == 0) return 0;
  const int64_t total = rows * (cols / 8);
  int64_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  geglu_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>((const __half*)in, (__half*)out, rows, cols / 8, in_pitch, out_pitch, cols);
  DF_CHECK_LAUNCH();
  return 0;
}
