// Multi-tensor Adam: the optimizer half of the train step the hot path serves
// (reference train_context_app_v2.py:113-127,174,189: torch.optim.Adam, betas (0, 0.999), one
// parameter group per tensor -> hundreds of tiny launches per step in the reference).
// One launch updates every parameter of a network.  HBM-bound: 4 reads + 3 writes of 4 bytes
// per element.  The arithmetic follows torch's (non-capturable, non-amsgrad) Adam step by step:
//   m = lerp(m, g, 1 - beta1);  v = v * beta2 + (1 - beta2) * g * g
//   p += -(lr / bias_correction1) * (m / (sqrt(v) / sqrt(bias_correction2) + eps))
#include "common.cuh"
#include "kernels.h"

namespace l2i {

__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
  // at::native::lerp: the form that is exact at both ends
  const float d = b - a;
  return (fabsf(w) < 0.5f) ? a + w * d : b - d * (1.0f - w);
}

struct AdamConsts { float beta2, omb1, omb2, eps; };   // omb = 1 - beta, rounded from double like torch's scalars

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_over_bc1, float bc2_sqrt,
                                         const AdamConsts& c) {
  m = torch_lerp(m, g, c.omb1);
  v = __fmul_rn(v, c.beta2);
  v = __fadd_rn(v, __fmul_rn(__fmul_rn(c.omb2, g), g));            // addcmul_: v + value * g * g
  const float eps = c.eps;
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
  p = __fadd_rn(p, __fmul_rn(-lr_over_bc1, __fdiv_rn(m, denom)));  // addcdiv_: p + value * (m / denom)
}

__global__ void __launch_bounds__(256)
adam_kernel(const AdamTensor* __restrict__ tensors, const int2* __restrict__ chunks, int chunk_elems, AdamConsts c,
            const long long* __restrict__ step_dev, double beta1, double beta2) {
  const int2 ck = chunks[blockIdx.x];
  AdamTensor t = tensors[ck.x];
  if (step_dev) {
    // graph-capturable form: the step count lives on the device (advanced by the caller before this launch), the
    // table's step_size field holds the plain learning rate and the bias corrections are formed here, in double like
    // torch's host code
    const double st = static_cast<double>(*step_dev);
    t.step_size = static_cast<float>(static_cast<double>(t.step_size) / (1.0 - pow(beta1, st)));
    t.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(beta2, st)));
  }
  const long long begin = 1LL * ck.y * chunk_elems;
  long long end = begin + chunk_elems;
  if (end > t.n) end = t.n;
  if (t.n <= 0) return;                              // parameter without a gradient this step: skipped, as torch does
  const float lr1 = t.step_size, b2 = t.bc2_sqrt;   // lr / (1 - beta1^step) and sqrt(1 - beta2^step) of THIS tensor's step
  float* __restrict__ p = t.p + begin;
  const float* __restrict__ g = t.g + begin;
  float* __restrict__ m = t.m + begin;
  float* __restrict__ v = t.v + begin;
  const int n = static_cast<int>(end - begin);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, lr1, b2, c);
      adam_one(pp.y, gg.y, mm.y, vv.y, lr1, b2, c);
      adam_one(pp.z, gg.z, mm.z, vv.z, lr1, b2, c);
      adam_one(pp.w, gg.w, mm.w, vv.w, lr1, b2, c);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) adam_one(p[i], g[i], m[i], v[i], lr1, b2, c);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) adam_one(p[i], g[i], m[i], v[i], lr1, b2, c);
  }
}

int adam_step(const void* tensors, const int* chunks, int n_chunks, int chunk_elems, double beta1, double beta2, double eps,
              const long long* step_dev, cudaStream_t stream) {
  if (!tensors || !chunks || n_chunks <= 0 || chunk_elems <= 0 || (chunk_elems & 3)) {
    set_error("adam_step: bad arguments (n_chunks=%d chunk_elems=%d)", n_chunks, chunk_elems);
    return L2I_ERR_BAD_ARG;
  }
  adam_kernel<<<n_chunks, 256, 0, stream>>>(reinterpret_cast<const AdamTensor*>(tensors),
                                            reinterpret_cast<const int2*>(chunks), chunk_elems,
                                            AdamConsts{static_cast<float>(beta2), static_cast<float>(1.0 - beta1),
                                                       static_cast<float>(1.0 - beta2), static_cast<float>(eps)},
                                            step_dev, beta1, beta2);
  return check_launch("adam_kernel");
}

}  // namespace l2i
