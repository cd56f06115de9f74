// Tensor-train path: truncated-QR decomposition pieces, reconstruction and the fused TT-Adam update.
//
// The reference's "TT-SVD" is a left-to-right sweep of COMPLETE Householder QR + truncation
// (tn_gradient/tt.py:127-136).  Keeping Q[:, :r] and R[:r, :] of a complete QR of L is exactly the orthogonal
// projection of L onto span(L[:, :r]):  Q_r = thin-QR(L[:, :r]),  R_r = Q_r^T L   (verified
// to 1.9e-07), so the m x m orthogonal factor is never formed:
//     sow_thin_qr  : Q_r  (CGS2 -- classical Gram-Schmidt with re-orthogonalisation, fp32)
//     tt_project   : R_r = Q_r^T L   (fp32 FMA accumulate; TF32/bf16 would break the 1e-5 reconstruction bound)
// Reconstruction results are gauge-invariant, so parity is defined on reconstruct()/to_matrix(), never on cores.
#include "common.cuh"

#include <algorithm>

namespace sowb {

// ------------------------------------------------------------------------------------------------
// thin QR (batched; one CTA per matrix)
// ------------------------------------------------------------------------------------------------
constexpr int kQrThreads = 512;
constexpr int kQrWarps = kQrThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CGS2 on a COLUMN-MAJOR work copy Wt[k][i] (r x m): every sweep is thread-per-row, so with this layout the 32
// lanes of a warp touch 32 consecutive floats of one column (one 128-byte wavefront) for every k.  (On the
// row-major layout the same sweeps cost one wavefront per element: measured 20 ms instead of < 1 ms for 4096x64.)
// Prologue / epilogue transpose X[:, :r] -> Wt and Wt -> Q through per-warp 32x33 smem tiles.
__device__ __forceinline__ void qr_transpose_in(const float* __restrict__ X, int ldx, float* __restrict__ Wt, int m,
                                                int r, float (*tile)[33], int warp, int lane) {
  const int tiles_i = (m + 31) / 32, tiles_k = (r + 31) / 32;
  for (int t = warp; t < tiles_i * tiles_k; t += kQrWarps) {
    const int i0 = (t / tiles_k) * 32, k0 = (t % tiles_k) * 32;
    for (int ii = 0; ii < 32; ++ii) {
      const int i = i0 + ii, k = k0 + lane;
      tile[ii][lane] = (i < m && k < r) ? X[static_cast<int64_t>(i) * ldx + k] : 0.f;
    }
    __syncwarp();
    for (int kk = 0; kk < 32; ++kk) {
      const int k = k0 + kk, i = i0 + lane;
      if (k < r && i < m) Wt[static_cast<int64_t>(k) * m + i] = tile[lane][kk];
    }
    __syncwarp();
  }
}
__device__ __forceinline__ void qr_transpose_out(const float* __restrict__ Wt, float* __restrict__ Q, int m, int r,
                                                 float (*tile)[33], int warp, int lane) {
  const int tiles_i = (m + 31) / 32, tiles_k = (r + 31) / 32;
  for (int t = warp; t < tiles_i * tiles_k; t += kQrWarps) {
    const int i0 = (t / tiles_k) * 32, k0 = (t % tiles_k) * 32;
    for (int kk = 0; kk < 32; ++kk) {
      const int k = k0 + kk, i = i0 + lane;
      tile[kk][lane] = (k < r && i < m) ? Wt[static_cast<int64_t>(k) * m + i] : 0.f;
    }
    __syncwarp();
    for (int ii = 0; ii < 32; ++ii) {
      const int i = i0 + ii, k = k0 + lane;
      if (i < m && k < r) Q[static_cast<int64_t>(i) * r + k] = tile[lane][ii];
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kQrThreads, 1)
thin_qr_kernel(const float* __restrict__ X, int64_t x_bs, int ldx, float* __restrict__ Q, int64_t q_bs,
               float* __restrict__ work, int m, int r) {
  extern __shared__ float qr_smem[];
  float* c = qr_smem;                 // [r]      projection coefficients
  float* red = qr_smem + r;           // [kQrWarps * 64] cross-warp partials
  float (*tiles)[33] = reinterpret_cast<float (*)[33]>(red + kQrWarps * 64);   // [kQrWarps][32][33]
  __shared__ float s_norm;
  const float* Xb = X + blockIdx.x * x_bs;
  float* Qb = Q + blockIdx.x * q_bs;
  float* Wt = work + static_cast<int64_t>(blockIdx.x) * m * r;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  qr_transpose_in(Xb, ldx, Wt, m, r, tiles + warp * 32, warp, lane);
  __syncthreads();

  for (int j = 0; j < r; ++j) {
    float* colj = Wt + static_cast<int64_t>(j) * m;
    for (int pass = 0; pass < 2 && j > 0; ++pass) {
      // c[k] = sum_i Wt[k][i] * Wt[j][i], k < j, in chunks of 32 coefficients held in registers
      for (int k0 = 0; k0 < j; k0 += 32) {
        const int kn = min(32, j - k0);
        float part[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) part[k] = 0.f;
        for (int i = tid; i < m; i += kQrThreads) {
          const float v = colj[i];
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k < kn) part[k] = fmaf(Wt[static_cast<int64_t>(k0 + k) * m + i], v, part[k]);
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (k < kn) {
            const float s = warp_sum(part[k]);
            if (lane == 0) red[warp * 64 + k] = s;
          }
        }
        __syncthreads();
        if (tid < kn) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < kQrWarps; ++w) s += red[w * 64 + tid];
          c[k0 + tid] = s;
        }
        __syncthreads();
      }
      // Wt[j][i] -= sum_k Wt[k][i] c[k]
      for (int i = tid; i < m; i += kQrThreads) {
        float acc = 0.f;
        for (int k = 0; k < j; ++k) acc = fmaf(Wt[static_cast<int64_t>(k) * m + i], c[k], acc);
        colj[i] -= acc;
      }
      __syncthreads();
    }
    // normalise
    float ss = 0.f;
    for (int i = tid; i < m; i += kQrThreads) {
      const float v = colj[i];
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < kQrWarps; ++w) s += red[w];
      s_norm = s;
    }
    __syncthreads();
    const float nrm2 = s_norm;
    const float inv = (nrm2 > 1e-37f) ? rsqrtf(nrm2) : 0.f;   // rank-deficient column -> zero vector
    for (int i = tid; i < m; i += kQrThreads) colj[i] *= inv;
    __syncthreads();
  }
  qr_transpose_out(Wt, Qb, m, r, tiles + warp * 32, warp, lane);
}

// ------------------------------------------------------------------------------------------------
// projection R[r, n] = Q[m, r]^T L[m, n]    (fp32, register-tiled; split over m with fp32 red.add)
// ------------------------------------------------------------------------------------------------
constexpr int kPjTN = 128;   // columns of L per CTA
constexpr int kPjTM = 32;    // rows of L per smem stage
constexpr int kPjThreads = 256;
constexpr int kPjRT = 64;    // rank tile

// thread layout: 16 (r) x 16 (n); each thread owns 4 (r) x 8 (n) outputs of a 64 x 128 tile
__global__ void __launch_bounds__(kPjThreads)
tt_project_kernel(const float* __restrict__ L, int64_t l_bs, const float* __restrict__ Q, int64_t q_bs,
                  float* __restrict__ R, int64_t r_bs, int m, int n, int r, int m_per_split) {
  __shared__ float sL[kPjTM][kPjTN];
  __shared__ float sQ[kPjTM][kPjRT + 1];
  const int b = blockIdx.z;
  const float* Lb = L + b * l_bs;
  const float* Qb = Q + b * q_bs;
  float* Rb = R + b * r_bs;
  const int n0 = blockIdx.x * kPjTN;
  const int m_begin = blockIdx.y * m_per_split;
  const int m_end = min(m, m_begin + m_per_split);
  const int tid = threadIdx.x;
  const int tr = tid / 16, tn = tid % 16;
  for (int r0 = 0; r0 < r; r0 += kPjRT) {
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;
    for (int mm0 = m_begin; mm0 < m_end; mm0 += kPjTM) {
      // stage L[mm0:mm0+32, n0:n0+128] and Q[mm0:mm0+32, r0:r0+64]
      for (int idx = tid; idx < kPjTM * kPjTN; idx += kPjThreads) {
        const int i = idx / kPjTN, c = idx % kPjTN;
        const int gi = mm0 + i, gc = n0 + c;
        sL[i][c] = (gi < m_end && gc < n) ? Lb[static_cast<int64_t>(gi) * n + gc] : 0.f;
      }
      for (int idx = tid; idx < kPjTM * kPjRT; idx += kPjThreads) {
        const int i = idx / kPjRT, k = idx % kPjRT;
        const int gi = mm0 + i, gk = r0 + k;
        sQ[i][k] = (gi < m_end && gk < r) ? Qb[static_cast<int64_t>(gi) * r + gk] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int i = 0; i < kPjTM; ++i) {
        float qv[4], lv[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) qv[a] = sQ[i][tr * 4 + a];
#pragma unroll
        for (int c = 0; c < 8; ++c) lv[c] = sL[i][tn + 16 * c];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(qv[a], lv[c], acc[a][c]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int gr = r0 + tr * 4 + a;
      if (gr >= r) continue;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int gc = n0 + tn + 16 * c;
        if (gc < n) {
          if (gridDim.y == 1) Rb[static_cast<int64_t>(gr) * n + gc] = acc[a][c];
          else atomicAdd(&Rb[static_cast<int64_t>(gr) * n + gc], acc[a][c]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pad + interleave / de-interleave for arbitrary order (tt.py:58,65,33 ; tt.py:242-247, utils.py:86-87)
//   interleaved index: digits (i1,o1,i2,o2,...,id,od), i_k base mm, o_k base nn
//   matrix index     : row = (i1..id) base mm, col = (o1..od) base nn
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_as_f32(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_f32<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_f32<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}
template <typename T>
__device__ __forceinline__ void store_from_f32(T* p, int64_t i, float v);
template <>
__device__ __forceinline__ void store_from_f32<float>(float* p, int64_t i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void store_from_f32<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) {
  p[i] = __float2bfloat16(v);
}

__device__ __forceinline__ void decode_interleaved(int64_t idx, int mm, int nn, int order, int64_t& row, int64_t& col) {
  // idx = ((((i1*nn + o1)*mm + i2)*nn + o2) ... ); peel digits from the least significant end
  int64_t rmul = 1, cmul = 1;
  row = 0;
  col = 0;
  for (int k = 0; k < order; ++k) {
    const int o = static_cast<int>(idx % nn);
    idx /= nn;
    const int i = static_cast<int>(idx % mm);
    idx /= mm;
    row += i * rmul;
    col += o * cmul;
    rmul *= mm;
    cmul *= nn;
  }
}

template <typename T>
__global__ void tt_interleave_kernel(const T* __restrict__ src, int M, int N, int mm, int nn, int order,
                                     float* __restrict__ out, int64_t total) {
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t row, col;
    decode_interleaved(idx, mm, nn, order, row, col);
    out[idx] = (row < M && col < N) ? load_as_f32<T>(src, row * N + col) : 0.f;
  }
}

template <typename T>
__global__ void tt_deinterleave_kernel(const float* __restrict__ src, int M, int N, int mm, int nn, int order,
                                       T* __restrict__ out, int64_t total) {
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t row, col;
    decode_interleaved(idx, mm, nn, order, row, col);
    if (row < M && col < N) store_from_f32<T>(out, row * N + col, src[idx]);
  }
}

// ------------------------------------------------------------------------------------------------
// small-K fp32 matmul  C[m, n] = A[m, r] . B[r, n]   (TT reconstruction chain, tt.py:213-237)
// 64 x 64 output tile per CTA, 4 x 4 per thread, whole K (= r <= 64 per pass) in smem.
// ------------------------------------------------------------------------------------------------
constexpr int kRkTile = 64;
__global__ void __launch_bounds__(256)
tt_matmul_rk_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int m, int n, int r) {
  __shared__ float sA[kRkTile][kRkTile + 1];  // [row][k]
  __shared__ float sB[kRkTile][kRkTile];      // [k][col]
  const int m0 = blockIdx.y * kRkTile, n0 = blockIdx.x * kRkTile;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  for (int k0 = 0; k0 < r; k0 += kRkTile) {
    for (int idx = tid; idx < kRkTile * kRkTile; idx += 256) {
      const int i = idx / kRkTile, k = idx % kRkTile;
      sA[i][k] = (m0 + i < m && k0 + k < r) ? A[static_cast<int64_t>(m0 + i) * r + k0 + k] : 0.f;
      const int kk = idx / kRkTile, c = idx % kRkTile;
      sB[kk][c] = (k0 + kk < r && n0 + c < n) ? B[static_cast<int64_t>(k0 + kk) * n + n0 + c] : 0.f;
    }
    __syncthreads();
    const int kn = min(kRkTile, r - k0);
    for (int k = 0; k < kn; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = sA[ty * 4 + a][k];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = sB[k][tx + 16 * c];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = m0 + ty * 4 + a;
    if (gi >= m) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gc = n0 + tx + 16 * c;
      if (gc < n) C[static_cast<int64_t>(gi) * n + gc] = acc[a][c];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Order-2 reconstruction fused with the TT-Adam update (see include/sow_b200.h: tt_adam_fused2)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
tt_adam_fused2_kernel(T* __restrict__ p, const T* __restrict__ g, const float* __restrict__ G1m,
                      const float* __restrict__ G2m, const float* __restrict__ G1v, const float* __restrict__ G2v,
                      int r, float* __restrict__ m_out, float* __restrict__ v_out, int M, int N, int mm, int nn,
                      float beta1, float omb1, float beta2, float omb2, float eps, float step_size, float lr_wd, int first_step) {
  extern __shared__ float fs[];
  const int P = mm * nn;  // interleaved matrix is P x P
  // smem: G1m tile [64][r+1], G1v tile [64][r+1], G2m tile [r][64], G2v tile [r][64]
  float* s1m = fs;
  float* s1v = s1m + 64 * (r + 1);
  float* s2m = s1v + 64 * (r + 1);
  float* s2v = s2m + r * 64;
  const int a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  float am[4][4], av[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) am[a][c] = av[a][c] = 0.f;
  if (!first_step) {
    for (int idx = tid; idx < 64 * r; idx += 256) {
      const int i = idx / r, k = idx % r;
      const bool ok = a0 + i < P;
      s1m[i * (r + 1) + k] = ok ? G1m[static_cast<int64_t>(a0 + i) * r + k] : 0.f;
      s1v[i * (r + 1) + k] = ok ? G1v[static_cast<int64_t>(a0 + i) * r + k] : 0.f;
      const int kk = idx / 64, c = idx % 64;
      const bool ok2 = b0 + c < P;
      s2m[kk * 64 + c] = ok2 ? G2m[static_cast<int64_t>(kk) * P + b0 + c] : 0.f;
      s2v[kk * 64 + c] = ok2 ? G2v[static_cast<int64_t>(kk) * P + b0 + c] : 0.f;
    }
    __syncthreads();
    for (int k = 0; k < r; ++k) {
      float x1m[4], x1v[4], x2m[4], x2v[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        x1m[a] = s1m[(ty * 4 + a) * (r + 1) + k];
        x1v[a] = s1v[(ty * 4 + a) * (r + 1) + k];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        x2m[c] = s2m[k * 64 + tx + 16 * c];
        x2v[c] = s2v[k * 64 + tx + 16 * c];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          am[a][c] = fmaf(x1m[a], x2m[c], am[a][c]);
          av[a][c] = fmaf(x1v[a], x2v[c], av[a][c]);
        }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int ga = a0 + ty * 4 + a;
    if (ga >= P) continue;
    const int i1 = ga / nn, o1 = ga % nn;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gb = b0 + tx + 16 * c;
      if (gb >= P) continue;
      const int i2 = gb / nn, o2 = gb % nn;
      const int64_t row = static_cast<int64_t>(i1) * mm + i2, col = static_cast<int64_t>(o1) * nn + o2;
      float mo = 0.f, vo = 0.f;
      if (row < M && col < N) {
        const int64_t e = row * N + col;
        const float gv = load_as_f32<T>(g, e);
        float pv = load_as_f32<T>(p, e);
        const float mp = am[a][c];
        const float vp = fmaxf(av[a][c], 0.f);                     // ttadam.py:84
        mo = beta1 * mp + omb1 * gv;                      // ttadam.py:92
        vo = beta2 * vp + omb2 * gv * gv;                 // ttadam.py:93
        pv -= step_size * (mo / (sqrtf(vo) + eps));                // ttadam.py:94,103,108
        if (lr_wd > 0.f) pv -= lr_wd * pv;                         // ttadam.py:110-111
        store_from_f32<T>(p, e, pv);
      }
      m_out[static_cast<int64_t>(ga) * P + gb] = mo;
      v_out[static_cast<int64_t>(ga) * P + gb] = vo;
    }
  }
}

// Dense variant for order > 2: m, v are fp32 (M, N) work matrices (already reconstructed + de-interleaved).
template <typename T>
__global__ void tt_adam_dense_kernel(T* __restrict__ p, const T* __restrict__ g, float* __restrict__ m,
                                     float* __restrict__ v, int64_t n, float beta1, float omb1, float beta2, float omb2,
                                     float eps, float step_size, float lr_wd) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gv = load_as_f32<T>(g, i);
    float pv = load_as_f32<T>(p, i);
    const float mo = beta1 * m[i] + omb1 * gv;
    const float vo = beta2 * fmaxf(v[i], 0.f) + omb2 * gv * gv;
    pv -= step_size * (mo / (sqrtf(vo) + eps));
    if (lr_wd > 0.f) pv -= lr_wd * pv;
    store_from_f32<T>(p, i, pv);
    m[i] = mo;
    v[i] = vo;
  }
}

static inline int grid_for(int64_t n, int threads) {
  const int64_t b = (n + threads - 1) / threads;
  return static_cast<int>(std::min<int64_t>(b, int64_t(num_sms()) * 16));
}

}  // namespace sowb

using namespace sowb;

extern "C" {

int sow_thin_qr(const float* X, int64_t x_batch_stride, int ldx, float* Q, int64_t q_batch_stride, int m, int r,
                int batch, void* ws, size_t ws_bytes, void* stream_) {
  SOWB_REQUIRE(X && Q && ws, "sow_thin_qr: null pointer argument");
  SOWB_REQUIRE(m > 0 && r > 0 && batch > 0 && ldx >= r, "sow_thin_qr: bad dimensions (m=%d r=%d ldx=%d batch=%d)", m, r, ldx, batch);
  SOWB_REQUIRE(r <= m, "sow_thin_qr: rank %d exceeds the row count %d (the reference fails here too: tt.py:135)", r, m);
  SOWB_REQUIRE(r <= 4096, "sow_thin_qr: rank %d too large", r);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t need = size_t(batch) * m * r * sizeof(float);
  if (ws_bytes < need) return set_error(SOWB_EWORKSPACE, "sow_thin_qr: workspace %zu B < required %zu B", ws_bytes, need);
  const size_t smem = (size_t(r) + kQrWarps * 64 + size_t(kQrWarps) * 32 * 33) * sizeof(float);
  SOWB_CHECK_CUDA(cudaFuncSetAttribute(thin_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  thin_qr_kernel<<<batch, kQrThreads, smem, stream>>>(X, x_batch_stride, ldx, Q, q_batch_stride, static_cast<float*>(ws), m, r);
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

int tt_project(const float* L, int64_t l_batch_stride, const float* Q, int64_t q_batch_stride, float* R,
               int64_t r_batch_stride, int m, int n, int r, int batch, void* stream_) {
  SOWB_REQUIRE(L && Q && R, "tt_project: null pointer argument");
  SOWB_REQUIRE(m > 0 && n > 0 && r > 0 && batch > 0, "tt_project: bad dimensions");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n_tiles = ceil_div(n, kPjTN);
  int splits = std::max(1, (2 * num_sms()) / std::max(1, n_tiles * batch));
  splits = std::min(splits, ceil_div(m, 4 * kPjTM));
  splits = std::max(1, std::min(splits, 65535));
  int m_per = round_up(ceil_div(m, splits), kPjTM);
  splits = ceil_div(m, m_per);
  if (splits > 1) {
    if (r_batch_stride == int64_t(r) * n || batch == 1) {
      const size_t bytes = (batch == 1) ? size_t(r) * n * 4 : size_t(batch) * r * n * 4;
      SOWB_CHECK_CUDA(cudaMemsetAsync(R, 0, bytes, stream));
    } else {
      for (int b = 0; b < batch; ++b) SOWB_CHECK_CUDA(cudaMemsetAsync(R + b * r_batch_stride, 0, size_t(r) * n * 4, stream));
    }
  }
  dim3 grid(n_tiles, splits, batch);
  tt_project_kernel<<<grid, kPjThreads, 0, stream>>>(L, l_batch_stride, Q, q_batch_stride, R, r_batch_stride, m, n, r, m_per);
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

int tt_interleave(const void* src, int M, int N, int mm, int nn, int order, float* out, int dtype, void* stream_) {
  SOWB_REQUIRE(src && out, "tt_interleave: null pointer argument");
  SOWB_REQUIRE(order >= 1 && order <= 8 && mm > 0 && nn > 0, "tt_interleave: bad order/shape");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int64_t total = 1;
  for (int k = 0; k < order; ++k) total *= int64_t(mm) * nn;
  if (dtype == SOWB_BF16)
    tt_interleave_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), M, N, mm, nn, order, out, total);
  else if (dtype == SOWB_F32)
    tt_interleave_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const float*>(src), M, N, mm, nn, order, out, total);
  else
    return set_error(SOWB_EINVAL, "tt_interleave: unknown dtype %d", dtype);
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

int tt_deinterleave(const float* src, int M, int N, int mm, int nn, int order, void* out, int dtype, void* stream_) {
  SOWB_REQUIRE(src && out, "tt_deinterleave: null pointer argument");
  SOWB_REQUIRE(order >= 1 && order <= 8 && mm > 0 && nn > 0, "tt_deinterleave: bad order/shape");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int64_t total = 1;
  for (int k = 0; k < order; ++k) total *= int64_t(mm) * nn;
  if (dtype == SOWB_BF16)
    tt_deinterleave_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, stream>>>(src, M, N, mm, nn, order, static_cast<__nv_bfloat16*>(out), total);
  else if (dtype == SOWB_F32)
    tt_deinterleave_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>(src, M, N, mm, nn, order, static_cast<float*>(out), total);
  else
    return set_error(SOWB_EINVAL, "tt_deinterleave: unknown dtype %d", dtype);
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

int tt_matmul_rk(const float* A, const float* B, float* C, int m, int n, int r, void* stream_) {
  SOWB_REQUIRE(A && B && C, "tt_matmul_rk: null pointer argument");
  SOWB_REQUIRE(m > 0 && n > 0 && r > 0, "tt_matmul_rk: bad dimensions");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dim3 grid(ceil_div(n, kRkTile), ceil_div(m, kRkTile));
  SOWB_REQUIRE(grid.y <= 65535, "tt_matmul_rk: m too large");
  tt_matmul_rk_kernel<<<grid, 256, 0, stream>>>(A, B, C, m, n, r);
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

int tt_adam_fused2(void* p, const void* g, const float* G1m, const float* G2m, const float* G1v,
                   const float* G2v, int r, float* m_out, float* v_out, int M, int N, int mm, int nn,
                   double beta1_d, double beta2_d, double eps_d, double step_size_d, double lr_wd_d, int first_step,
                   int dtype, void* stream_) {
  const float beta1 = float(beta1_d), beta2 = float(beta2_d), omb1 = float(1.0 - beta1_d), omb2 = float(1.0 - beta2_d);
  const float eps = float(eps_d), step_size = float(step_size_d), lr_wd = float(lr_wd_d);
  SOWB_REQUIRE(p && g && m_out && v_out, "tt_adam_fused2: null pointer argument");
  SOWB_REQUIRE(first_step || (G1m && G2m && G1v && G2v), "tt_adam_fused2: null core pointer");
  SOWB_REQUIRE(r > 0 && r <= 128, "tt_adam_fused2: rank %d unsupported (1..128)", r);
  SOWB_REQUIRE(int64_t(mm) * mm >= M && int64_t(nn) * nn >= N, "tt_adam_fused2: mm/nn too small for (M,N)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int P = mm * nn;
  dim3 grid(ceil_div(P, 64), ceil_div(P, 64));
  const size_t smem = (size_t(2) * 64 * (r + 1) + size_t(2) * r * 64) * sizeof(float);
  if (dtype == SOWB_BF16) {
    auto k = tt_adam_fused2_kernel<__nv_bfloat16>;
    SOWB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    k<<<grid, 256, smem, stream>>>(static_cast<__nv_bfloat16*>(p), static_cast<const __nv_bfloat16*>(g), G1m, G2m, G1v,
                                   G2v, r, m_out, v_out, M, N, mm, nn, beta1, omb1, beta2, omb2, eps, step_size, lr_wd, first_step);
  } else if (dtype == SOWB_F32) {
    auto k = tt_adam_fused2_kernel<float>;
    SOWB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    k<<<grid, 256, smem, stream>>>(static_cast<float*>(p), static_cast<const float*>(g), G1m, G2m, G1v, G2v, r, m_out,
                                   v_out, M, N, mm, nn, beta1, omb1, beta2, omb2, eps, step_size, lr_wd, first_step);
  } else {
    return set_error(SOWB_EINVAL, "tt_adam_fused2: unknown dtype %d", dtype);
  }
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

int tt_adam_dense(void* p, const void* g, float* m, float* v, int64_t numel, double beta1_d, double beta2_d,
                  double eps_d, double step_size_d, double lr_wd_d, int dtype, void* stream_) {
  const float beta1 = float(beta1_d), beta2 = float(beta2_d), omb1 = float(1.0 - beta1_d), omb2 = float(1.0 - beta2_d);
  const float eps = float(eps_d), step_size = float(step_size_d), lr_wd = float(lr_wd_d);
  SOWB_REQUIRE(p && g && m && v && numel > 0, "tt_adam_dense: bad argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (dtype == SOWB_BF16)
    tt_adam_dense_kernel<__nv_bfloat16><<<grid_for(numel, 256), 256, 0, stream>>>(
        static_cast<__nv_bfloat16*>(p), static_cast<const __nv_bfloat16*>(g), m, v, numel, beta1, omb1, beta2, omb2, eps, step_size, lr_wd);
  else if (dtype == SOWB_F32)
    tt_adam_dense_kernel<float><<<grid_for(numel, 256), 256, 0, stream>>>(static_cast<float*>(p), static_cast<const float*>(g),
                                                                         m, v, numel, beta1, omb1, beta2, omb2, eps, step_size, lr_wd);
  else
    return set_error(SOWB_EINVAL, "tt_adam_dense: unknown dtype %d", dtype);
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

}  // extern "C"
