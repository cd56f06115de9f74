// Multi-tensor fused Adam / AdamW: every tensor of a param group updated by ONE launch.
//
// Replaces torch.optim.AdamW over the factor + dense groups (scripts/simple_train.py:502-511), which in eager
// mode issues ~8 elementwise kernels per group chunk and, for bf16 parameters, rounds to bf16 after every one
// of them.  Here each element is read once (p, g, m, v), updated in fp32 registers and written once:
// 4 reads + 3 writes of the parameter dtype per element -> 14 B/element in bf16, the HBM floor for Adam.
// State stays per-parameter and in the parameter dtype, exactly where torch keeps it, so
// scripts/utils/training_utils.py:257-277 (reset_optimizer, which REBINDS exp_avg / exp_avg_sq) keeps working:
// the host side re-resolves the pointers whenever they change.
#include "common.cuh"

namespace sowb {

struct AdamChunk {   // 40 bytes; the host passes an int64[n_chunks][5] table with exactly this layout
  void* p;
  const void* g;
  void* m;
  void* v;
  int64_t n;
};

struct AdamHyper {
  float lr, beta1, omb1, beta2, omb2, eps, decay_mul, weight_decay, step_size, inv_sqrt_bc2;
  int decoupled;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h) {
  if (h.decoupled) p *= h.decay_mul;
  else g = fmaf(h.weight_decay, p, g);
  m = fmaf(h.beta1, m, h.omb1 * g);
  v = fmaf(h.beta2, v, h.omb2 * g * g);
  const float denom = sqrtf(v) * h.inv_sqrt_bc2 + h.eps;
  p -= h.step_size * (m / denom);
}

// CUDA-graph friendly step counter: the bias corrections are derived ON THE DEVICE from a counter that the launch
// sequence itself advances, so a captured optimizer step replays with the right corrections (host-computed ones would
// be frozen into the graph).
__global__ void adam_step_inc_kernel(float* step) { *step += 1.0f; }

__device__ __forceinline__ AdamHyper with_device_step(AdamHyper h, const float* __restrict__ step_dev) {
  if (step_dev != nullptr) {
    const double t = static_cast<double>(*step_dev);
    const double bc1 = 1.0 - pow(static_cast<double>(h.beta1), t);
    const double bc2 = 1.0 - pow(static_cast<double>(h.beta2), t);
    h.step_size = static_cast<float>(static_cast<double>(h.lr) / bc1);
    h.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  }
  return h;
}

__global__ void __launch_bounds__(256)
adam_multi_bf16_kernel(const AdamChunk* __restrict__ chunks, const AdamHyper h0, const float* __restrict__ step_dev) {
  const AdamHyper h = with_device_step(h0, step_dev);
  const AdamChunk c = chunks[blockIdx.x];
  __nv_bfloat16* p = static_cast<__nv_bfloat16*>(c.p);
  const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(c.g);
  __nv_bfloat16* m = static_cast<__nv_bfloat16*>(c.m);
  __nv_bfloat16* v = static_cast<__nv_bfloat16*>(c.v);
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int64_t done = 0;
  if (aligned) {
    const int64_t nvec = c.n / 8;
    for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
      uint4 pp = reinterpret_cast<uint4*>(p)[i];
      const uint4 gg = reinterpret_cast<const uint4*>(g)[i];
      uint4 mm = reinterpret_cast<uint4*>(m)[i];
      uint4 vv = reinterpret_cast<uint4*>(v)[i];
      __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pp);
      const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gg);
      __nv_bfloat162* m2 = reinterpret_cast<__nv_bfloat162*>(&mm);
      __nv_bfloat162* v2 = reinterpret_cast<__nv_bfloat162*>(&vv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 pf = __bfloat1622float2(p2[j]), gf = __bfloat1622float2(g2[j]);
        float2 mf = __bfloat1622float2(m2[j]), vf = __bfloat1622float2(v2[j]);
        adam_update(pf.x, gf.x, mf.x, vf.x, h);
        adam_update(pf.y, gf.y, mf.y, vf.y, h);
        p2[j] = __floats2bfloat162_rn(pf.x, pf.y);
        m2[j] = __floats2bfloat162_rn(mf.x, mf.y);
        v2[j] = __floats2bfloat162_rn(vf.x, vf.y);
      }
      reinterpret_cast<uint4*>(p)[i] = pp;
      reinterpret_cast<uint4*>(m)[i] = mm;
      reinterpret_cast<uint4*>(v)[i] = vv;
    }
    done = nvec * 8;
  }
  for (int64_t i = done + threadIdx.x; i < c.n; i += blockDim.x) {
    float pf = __bfloat162float(p[i]), mf = __bfloat162float(m[i]), vf = __bfloat162float(v[i]);
    adam_update(pf, __bfloat162float(g[i]), mf, vf, h);
    p[i] = __float2bfloat16(pf);
    m[i] = __float2bfloat16(mf);
    v[i] = __float2bfloat16(vf);
  }
}

__global__ void __launch_bounds__(256)
adam_multi_f32_kernel(const AdamChunk* __restrict__ chunks, const AdamHyper h0, const float* __restrict__ step_dev) {
  const AdamHyper h = with_device_step(h0, step_dev);
  const AdamChunk c = chunks[blockIdx.x];
  float* p = static_cast<float*>(c.p);
  const float* g = static_cast<const float*>(c.g);
  float* m = static_cast<float*>(c.m);
  float* v = static_cast<float*>(c.v);
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int64_t done = 0;
  if (aligned) {
    const int64_t nvec = c.n / 4;
    for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adam_update(pp.x, gg.x, mm.x, vv.x, h);
      adam_update(pp.y, gg.y, mm.y, vv.y, h);
      adam_update(pp.z, gg.z, mm.z, vv.z, h);
      adam_update(pp.w, gg.w, mm.w, vv.w, h);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    done = nvec * 4;
  }
  for (int64_t i = done + threadIdx.x; i < c.n; i += blockDim.x) adam_update(p[i], g[i], m[i], v[i], h);
}

}  // namespace sowb

using namespace sowb;

extern "C" {

int sow_adam_chunk_elems(void) { return 32768; }

int sow_adam_multi_ex(const void* chunks_dev, int n_chunks, int64_t total_elems, double lr, double beta1, double beta2,
                      double eps, double weight_decay, double bias_correction1, double bias_correction2, int decoupled,
                      int dtype, void* stream_);

int sow_adam_multi(const void* chunks_dev, int n_chunks, double lr, double beta1, double beta2, double eps,
                   double weight_decay, double bias_correction1, double bias_correction2, int decoupled, int dtype,
                   void* stream_) {
  return sow_adam_multi_ex(chunks_dev, n_chunks, 0, lr, beta1, beta2, eps, weight_decay, bias_correction1,
                           bias_correction2, decoupled, dtype, stream_);
}

static int adam_launch(const void* chunks_dev, int n_chunks, int64_t total_elems, double lr, double beta1, double beta2,
                       double eps, double weight_decay, double bias_correction1, double bias_correction2, float* step_dev,
                       int decoupled, int dtype, void* stream_);

int sow_adam_multi_ex(const void* chunks_dev, int n_chunks, int64_t total_elems, double lr, double beta1, double beta2, double eps,
                      double weight_decay, double bias_correction1, double bias_correction2, int decoupled,
                      int dtype, void* stream_) {
  SOWB_REQUIRE(bias_correction1 > 0.0 && bias_correction2 > 0.0, "sow_adam_multi: bias corrections must be positive");
  return adam_launch(chunks_dev, n_chunks, total_elems, lr, beta1, beta2, eps, weight_decay, bias_correction1,
                     bias_correction2, nullptr, decoupled, dtype, stream_);
}

int sow_adam_multi_dev(const void* chunks_dev, int n_chunks, int64_t total_elems, double lr, double beta1, double beta2,
                       double eps, double weight_decay, float* step_dev, int decoupled, int dtype, void* stream_) {
  SOWB_REQUIRE(step_dev != nullptr, "sow_adam_multi_dev: null step counter");
  return adam_launch(chunks_dev, n_chunks, total_elems, lr, beta1, beta2, eps, weight_decay, 1.0, 1.0, step_dev, decoupled,
                     dtype, stream_);
}

static int adam_launch(const void* chunks_dev, int n_chunks, int64_t total_elems, double lr, double beta1, double beta2,
                       double eps, double weight_decay, double bias_correction1, double bias_correction2, float* step_dev,
                       int decoupled, int dtype, void* stream_) {
  if (n_chunks <= 0) return SOWB_OK;
  SOWB_REQUIRE(chunks_dev != nullptr, "sow_adam_multi: null chunk table");
  if (int rc0 = ensure_context_for(chunks_dev)) return rc0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (step_dev != nullptr) {
    adam_step_inc_kernel<<<1, 1, 0, stream>>>(step_dev);
    SOWB_CHECK_CUDA(cudaGetLastError());
  }
  AdamHyper h;
  h.lr = float(lr);
  h.beta1 = float(beta1);
  h.omb1 = float(1.0 - beta1);
  h.beta2 = float(beta2);
  h.omb2 = float(1.0 - beta2);
  h.eps = float(eps);
  h.decay_mul = float(1.0 - lr * weight_decay);
  h.weight_decay = float(weight_decay);
  h.step_size = float(lr / bias_correction1);
  h.inv_sqrt_bc2 = float(1.0 / sqrt(bias_correction2));
  h.decoupled = decoupled;
  // algorithmic bytes: p, m, v read+written, g read = 7 accesses of the element size
  ProfileScope prof(stream, PROF_ADAM, 7.0 * double(total_elems) * (dtype == SOWB_BF16 ? 2 : 4));
  if (dtype == SOWB_BF16)
    adam_multi_bf16_kernel<<<n_chunks, 256, 0, stream>>>(static_cast<const AdamChunk*>(chunks_dev), h, step_dev);
  else if (dtype == SOWB_F32)
    adam_multi_f32_kernel<<<n_chunks, 256, 0, stream>>>(static_cast<const AdamChunk*>(chunks_dev), h, step_dev);
  else
    return set_error(SOWB_EINVAL, "sow_adam_multi: unknown dtype %d", dtype);
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

}  // extern "C"
