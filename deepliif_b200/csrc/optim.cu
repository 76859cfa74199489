// Fused Adam over a flat parameter bucket (reference: torch.optim.Adam(lr, betas=(beta1, 0.999)) on all generator /
// all discriminator parameters, DeepLIIF_model.py:133-146).  One launch per bucket: p, g, m, v are contiguous fp32
// arrays (the gradient bucket is the same memory NCCL all-reduces).  PyTorch's update, single-tensor form:
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "internal.h"

namespace dlb {
namespace {
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2_sqrt, float grad_scale) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}
// Same update with the step-dependent scalars read from device memory: hyper = {lr, bias_correction1,
// sqrt(bias_correction2), grad_scale}.  A captured CUDA graph replays this launch unchanged while the host refreshes
// the four floats between replays (lr schedule, step count).
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, long long n, const float* __restrict__ hyper,
                                                       float b1, float b2, float eps) {
  const float lr = hyper[0], bc1 = hyper[1], bc2_sqrt = hyper[2], grad_scale = hyper[3];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}
}  // namespace
}  // namespace dlb

using namespace dlb;

extern "C" int dlb_adam_hyper(float lr, float beta1, float beta2, int step, float grad_scale, float* hyper4_host) {
  if (step < 1) return set_error("dlb_adam_hyper: step starts at 1");
  hyper4_host[0] = lr;
  hyper4_host[1] = 1.f - powf(beta1, static_cast<float>(step));
  hyper4_host[2] = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  hyper4_host[3] = grad_scale;
  return 0;
}

extern "C" int dlb_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper4_dev,
                                 float beta1, float beta2, float eps, dlb_stream_t stream) {
  long long grid = (n + 255) / 256; if (grid > 148 * 16) grid = 148 * 16; if (grid < 1) grid = 1;
  adam_dev_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(p, g, m, v, n, hyper4_dev, beta1, beta2, eps);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("adam_dev_kernel launch");
  return 0;
}

extern "C" int dlb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                             float eps, int step, float grad_scale, dlb_stream_t stream) {
  if (step < 1) return set_error("dlb_adam_step: step starts at 1");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  long long grid = (n + 255) / 256; if (grid > 148 * 16) grid = 148 * 16; if (grid < 1) grid = 1;
  adam_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("adam_kernel launch");
  return 0;
}
