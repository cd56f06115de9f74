// C-ABI entry points for the SoW linear forward / backward (include/sow_b200.h) on top of the tcgen05 GEMM.
#include "common.cuh"
#include "gemm.cuh"

#include <algorithm>

namespace sowb {

// ------------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------------
// A (in, r) -> A_pad (in, r_pad), zero padded: gives the factor a TMA-legal 128-byte row pitch.
__global__ void pack_factor_kernel(const __nv_bfloat16* __restrict__ A, __nv_bfloat16* __restrict__ A_pad, int in,
                                   int r, int r_pad) {
  const int64_t n = static_cast<int64_t>(in) * r_pad;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / r_pad), col = static_cast<int>(i % r_pad);
    A_pad[i] = (col < r) ? A[static_cast<int64_t>(row) * r + col] : __float2bfloat16(0.f);
  }
}

// fp32 split-K PARTIALS -> bf16 gradients in the reference layouts, summed over the splits in a fixed order (so the
// factor gradients are bit-reproducible run to run):
//   dA[i, j] = sum_s partA[s][i, j]       (partA is [splitsA][in , r_pad])
//   dB[j, o] = sum_s partB[s][o, j]       (partB is [splitsB][out, r_pad], i.e. dB^T)
__global__ void finalize_factor_grads_kernel(const float* __restrict__ partA, const float* __restrict__ partB,
                                             __nv_bfloat16* __restrict__ dA, __nv_bfloat16* __restrict__ dB, int in,
                                             int out, int r, int r_pad, int splitsA, int splitsB) {
  __shared__ float tile[32][33];
  const int nA_blocks = ceil_div(in * r, 1024);
  if (static_cast<int>(blockIdx.x) < nA_blocks) {
    const int i = blockIdx.x * 1024 + threadIdx.y * 32 + threadIdx.x;
    if (i < in * r) {
      const int row = i / r, col = i % r;
      const int64_t off = static_cast<int64_t>(row) * r_pad + col, stride = static_cast<int64_t>(in) * r_pad;
      float s = 0.f;
      for (int k = 0; k < splitsA; ++k) s += partA[k * stride + off];
      dA[i] = __float2bfloat16(s);
    }
    return;
  }
  // transpose tiles of partB: block handles 32 (o) x 32 (j)
  const int b = blockIdx.x - nA_blocks;
  const int jt = ceil_div(r, 32);
  const int o0 = (b / jt) * 32, j0 = (b % jt) * 32;
  {
    const int o = o0 + threadIdx.y, j = j0 + threadIdx.x;
    float s = 0.f;
    if (o < out && j < r) {
      const int64_t off = static_cast<int64_t>(o) * r_pad + j, stride = static_cast<int64_t>(out) * r_pad;
      for (int k = 0; k < splitsB; ++k) s += partB[k * stride + off];
    }
    tile[threadIdx.y][threadIdx.x] = s;
  }
  __syncthreads();
  {
    const int j = j0 + threadIdx.y, o = o0 + threadIdx.x;
    if (o < out && j < r) dB[static_cast<int64_t>(j) * out + o] = __float2bfloat16(tile[threadIdx.x][threadIdx.y]);
  }
}

// dbias[o] = sum_t dY[t, o]: block = 64 columns x 4 row-lanes, grid.y splits T; every block writes its partial row,
// the conversion kernel sums the partial rows in a fixed order (bit-reproducible).
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ part, int64_t T, int out) {
  __shared__ float red[4][64];
  const int col = blockIdx.x * 64 + threadIdx.x;
  const int64_t rows_per = (T + gridDim.y - 1) / gridDim.y;
  const int64_t t0 = blockIdx.y * rows_per, t1 = min(T, t0 + rows_per);
  float s = 0.f;
  if (col < out)
    for (int64_t t = t0 + threadIdx.y; t < t1; t += blockDim.y) s += __bfloat162float(dy[t * out + col]);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < out)
    part[static_cast<int64_t>(blockIdx.y) * out + col] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}
__global__ void sum_partials_to_bf16_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ dst, int n, int parts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float s = 0.f;
    for (int k = 0; k < parts; ++k) s += part[static_cast<int64_t>(k) * n + i];
    dst[i] = __float2bfloat16(s);
  }
}

// ------------------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------------------
struct Operand {
  const void* ptr;
  uint64_t rows, cols;  // as stored in global memory, row-major, `cols` contiguous
  uint64_t pitch_elems;
};

// Operand A of D = A.B: logical [M, K].  K-major  <=> stored [M rows, K cols];  MN-major <=> stored [K rows, M cols].
// Operand B of D = A.B: logical [K, N].  K-major  <=> stored [N rows, K cols];  MN-major <=> stored [K rows, N cols].
template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const Operand& A, const Operand& B, const Operand* A2, const Operand* B2, void* C_bf16,
                       float* C_f32, int ldc, int M, int N, int K, int K2, float alpha, const void* bias,
                       bool split_k, cudaStream_t stream, int prof_class = PROF_GEMM_SKINNY, int* splits_out = nullptr) {
  using S = GemmSmem<BN>;
  CUtensorMap tmA, tmB, tmA2, tmB2, tmC;
  int rc;
  rc = make_tensor_map_2d(&tmA, A.ptr, A.cols, A.rows, A.pitch_elems * 2, 64, A_MN ? 64 : kBM, 2);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmB, B.ptr, B.cols, B.rows, B.pitch_elems * 2, 64, B_MN ? 64 : BN, 2);
  if (rc) return rc;
  tmA2 = tmA;
  tmB2 = tmB;
  if (K2 > 0) {
    rc = make_tensor_map_2d(&tmA2, A2->ptr, A2->cols, A2->rows, A2->pitch_elems * 2, 64, A_MN ? 64 : kBM, 2);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmB2, B2->ptr, B2->cols, B2->rows, B2->pitch_elems * 2, 64, B_MN ? 64 : BN, 2);
    if (rc) return rc;
  }
  tmC = tmA;
  if (EPI == EPI_BF16_TMA) {
    rc = make_tensor_map_2d(&tmC, C_bf16, N, M, static_cast<uint64_t>(ldc) * 2, kStoreBoxCols, kBM, 2);
    if (rc) return rc;
  }
  GemmParams p;
  p.M = M;
  p.N = N;
  p.kb_main = ceil_div(K, kBK);
  p.kb_tail = ceil_div(K2, kBK);
  p.m_tiles = ceil_div(M, kBM);
  p.n_tiles = ceil_div(N, BN);
  const int kb_total = p.kb_main + p.kb_tail;
  const int sms = num_sms();
  int splits = 1;
  if (split_k) {
    splits = sms / (p.m_tiles * p.n_tiles);
    if (splits < 1) splits = 1;
    if (splits > kb_total) splits = kb_total;
  }
  p.kb_per_split = ceil_div(kb_total, splits);
  p.splits = ceil_div(kb_total, p.kb_per_split);
  p.alpha = alpha;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.out_f32 = C_f32;
  p.split_stride = static_cast<int64_t>(M) * ldc;
  p.ldc = ldc;
  if (splits_out) *splits_out = p.splits;
  const int total = p.m_tiles * p.n_tiles * p.splits;
  if (total <= 0 || kb_total <= 0) return SOWB_OK;
  auto kern = sow_gemm_kernel<BN, A_MN, B_MN, EPI>;
  SOWB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  const int grid = total < sms ? total : sms;
  // algorithmic flops: un-padded contraction lengths
  ProfileScope prof(stream, prof_class, 2.0 * double(M) * double(N) * (double(K) + double(K2)));
  kern<<<grid, kGemmThreads, S::kTotal, stream>>>(tmA, tmB, tmA2, tmB2, tmC, p);
  SOWB_CHECK_CUDA(cudaGetLastError());
  return SOWB_OK;
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// number of split-K partials launch_gemm<64, ..., EPI_F32_PARTIAL> produces for an [M x 64] output contracted over K
static int splitk_count(int M, int K) {
  const int tiles = ceil_div(M, kBM), kb_total = ceil_div(K, kBK);
  int splits = num_sms() / std::max(1, tiles);
  splits = std::max(1, std::min(splits, kb_total));
  const int per = ceil_div(kb_total, splits);
  return ceil_div(kb_total, per);
}
constexpr int kBiasParts = 64;

static int check_common(const char* fn, int64_t T, int in, int out, int r, int dtype, const void* any_dev_ptr = nullptr) {
  if (any_dev_ptr != nullptr) {
    const int rc0 = ensure_context_for(any_dev_ptr);   // backward runs on autograd's worker thread
    if (rc0) return rc0;
  }
  if (dtype != SOWB_BF16)
    return set_error(SOWB_EINVAL, "%s: only SOWB_BF16 is implemented on device (fp32 modules use the host-side bf16 compute policy)", fn);
  if (T <= 0 || in <= 0 || out <= 0 || r <= 0) return set_error(SOWB_EINVAL, "%s: non-positive dimension", fn);
  if (in % 8 != 0 || out % 8 != 0)
    return set_error(SOWB_EINVAL, "%s: in=%d and out=%d must be multiples of 8 (16-byte TMA row pitch)", fn, in, out);
  if (T >= (int64_t(1) << 31)) return set_error(SOWB_EINVAL, "%s: T too large", fn);
  return require_sm100();
}

}  // namespace sowb

using namespace sowb;

extern "C" {

int sow_rank_pad(int r) { return round_up(r, kBK); }

size_t sow_workspace_bytes(int op, int64_t T, int in, int out, int r) {
  const size_t r_pad = round_up(r, kBK);
  switch (op) {
    case SOWB_OP_LINEAR_FWD:
      return align256(size_t(in) * r_pad * 2);
    case SOWB_OP_LINEAR_BWD: {
      // bwd_factors: fp32 split-K partials of dA and dB^T, partial rows of dbias; bwd_dx: padded A
      const int sa = splitk_count(in, static_cast<int>(T)), sb = splitk_count(out, static_cast<int>(T));
      return align256(size_t(sa) * in * r_pad * 4) + align256(size_t(sb) * out * r_pad * 4) +
             align256(size_t(kBiasParts) * out * 4) + align256(size_t(in) * r_pad * 2);
    }
    default:
      return 0;
  }
}

int sow_linear_fwd(const void* x, const void* W, const void* A, const void* B, const void* bias, void* y,
                   void* t_out, int64_t T, int in, int out, int r, float scale, int dtype, void* ws,
                   size_t ws_bytes, void* stream_) {
  int rc = check_common("sow_linear_fwd", T, in, out, r, dtype, x);
  if (rc) return rc;
  SOWB_REQUIRE(x && A && B && y && t_out && ws, "sow_linear_fwd: null pointer argument");
  if (ws_bytes < sow_workspace_bytes(SOWB_OP_LINEAR_FWD, T, in, out, r))
    return set_error(SOWB_EWORKSPACE, "sow_linear_fwd: workspace %zu B < required %zu B", ws_bytes,
                     sow_workspace_bytes(SOWB_OP_LINEAR_FWD, T, in, out, r));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int r_pad = round_up(r, kBK);
  __nv_bfloat16* A_pad = static_cast<__nv_bfloat16*>(ws);

  {
    const int64_t n = int64_t(in) * r_pad;
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 4096));
    pack_factor_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(A), A_pad, in, r, r_pad);
    SOWB_CHECK_CUDA(cudaGetLastError());
  }
  // t = scale * x . A_pad            [T, r_pad]
  Operand opX{x, uint64_t(T), uint64_t(in), uint64_t(in)};
  Operand opApadMN{A_pad, uint64_t(in), uint64_t(r_pad), uint64_t(r_pad)};  // [K=in rows, N=r_pad cols]
  rc = launch_gemm<64, false, true, EPI_BF16_TMA>(opX, opApadMN, nullptr, nullptr, t_out, nullptr, r_pad,
                                                  static_cast<int>(T), r_pad, in, 0, scale, nullptr, false, stream);
  if (rc) return rc;
  // y = x . W + t . B (+ bias)
  Operand opT{t_out, uint64_t(T), uint64_t(r_pad), uint64_t(r_pad)};
  Operand opBMN{B, uint64_t(r), uint64_t(out), uint64_t(out)};  // [K=r rows (OOB rows read as 0), N=out cols]
  if (W != nullptr) {
    Operand opWMN{W, uint64_t(in), uint64_t(out), uint64_t(out)};  // [K=in rows, N=out cols]
    rc = launch_gemm<256, false, true, EPI_BF16_TMA>(opX, opWMN, &opT, &opBMN, y, nullptr, out, static_cast<int>(T),
                                                     out, in, r_pad, 1.0f, bias, false, stream, PROF_GEMM_FWD);
  } else {
    rc = launch_gemm<256, false, true, EPI_BF16_TMA>(opT, opBMN, nullptr, nullptr, y, nullptr, out,
                                                     static_cast<int>(T), out, r_pad, 0, 1.0f, bias, false, stream,
                                                     PROF_GEMM_FWD);
  }
  return rc;
}

int sow_linear_bwd_factors(const void* dy, const void* x, const void* t, const void* B, void* dt, void* dA,
                           void* dB, void* dbias, int64_t T, int in, int out, int r, float scale, int dtype,
                           void* ws, size_t ws_bytes, void* stream_) {
  int rc = check_common("sow_linear_bwd_factors", T, in, out, r, dtype, dy);
  if (rc) return rc;
  SOWB_REQUIRE(dy && x && t && B && dt && dA && dB && ws, "sow_linear_bwd_factors: null pointer argument");
  if (ws_bytes < sow_workspace_bytes(SOWB_OP_LINEAR_BWD, T, in, out, r))
    return set_error(SOWB_EWORKSPACE, "sow_linear_bwd_factors: workspace %zu B < required %zu B", ws_bytes,
                     sow_workspace_bytes(SOWB_OP_LINEAR_BWD, T, in, out, r));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int r_pad = round_up(r, kBK);
  uint8_t* wsp = static_cast<uint8_t*>(ws);
  const int sa_max = splitk_count(in, static_cast<int>(T)), sb_max = splitk_count(out, static_cast<int>(T));
  float* partA = reinterpret_cast<float*>(wsp);
  float* partB = reinterpret_cast<float*>(wsp + align256(size_t(sa_max) * in * r_pad * 4));
  float* partBias = reinterpret_cast<float*>(wsp + align256(size_t(sa_max) * in * r_pad * 4) + align256(size_t(sb_max) * out * r_pad * 4));
  int splitsA = 1, splitsB = 1;

  // dt = scale * dY . B^T     [T, r_pad]; B is read K-major: [N=r rows (OOB rows -> 0), K=out cols]
  Operand opDY{dy, uint64_t(T), uint64_t(out), uint64_t(out)};
  Operand opBK{B, uint64_t(r), uint64_t(out), uint64_t(out)};
  rc = launch_gemm<64, false, false, EPI_BF16_TMA>(opDY, opBK, nullptr, nullptr, dt, nullptr, r_pad,
                                                   static_cast<int>(T), r_pad, out, 0, scale, nullptr, false, stream);
  if (rc) return rc;
  // dB^T [out, r_pad] = dY^T . t   (both operands MN-major; K = T, split-K into fp32 partials summed by finalize)
  Operand opT{t, uint64_t(T), uint64_t(r_pad), uint64_t(r_pad)};
  rc = launch_gemm<64, true, true, EPI_F32_PARTIAL>(opDY, opT, nullptr, nullptr, nullptr, partB, r_pad, out, r_pad,
                                                   static_cast<int>(T), 0, 1.0f, nullptr, true, stream, PROF_GEMM_SPLITK,
                                                   &splitsB);
  if (rc) return rc;
  // dA [in, r_pad] = x^T . dt
  Operand opX{x, uint64_t(T), uint64_t(in), uint64_t(in)};
  Operand opDT{dt, uint64_t(T), uint64_t(r_pad), uint64_t(r_pad)};
  rc = launch_gemm<64, true, true, EPI_F32_PARTIAL>(opX, opDT, nullptr, nullptr, nullptr, partA, r_pad, in, r_pad,
                                                   static_cast<int>(T), 0, 1.0f, nullptr, true, stream, PROF_GEMM_SPLITK,
                                                   &splitsA);
  if (rc) return rc;
  {
    const int nA_blocks = ceil_div(in * r, 1024);
    const int nB_blocks = ceil_div(out, 32) * ceil_div(r, 32);
    if (splitsA > sa_max || splitsB > sb_max)
      return set_error(SOWB_EWORKSPACE, "sow_linear_bwd_factors: split-K count exceeds the workspace plan");
    finalize_factor_grads_kernel<<<nA_blocks + nB_blocks, dim3(32, 32), 0, stream>>>(
        partA, partB, static_cast<__nv_bfloat16*>(dA), static_cast<__nv_bfloat16*>(dB), in, out, r, r_pad, splitsA, splitsB);
    SOWB_CHECK_CUDA(cudaGetLastError());
  }
  if (dbias != nullptr) {
    dim3 grid(ceil_div(out, 64), static_cast<unsigned>(std::min<int64_t>(kBiasParts, (T + 255) / 256)));
    colsum_kernel<<<grid, dim3(64, 4), 0, stream>>>(static_cast<const __nv_bfloat16*>(dy), partBias, T, out);
    SOWB_CHECK_CUDA(cudaGetLastError());
    sum_partials_to_bf16_kernel<<<ceil_div(out, 256), 256, 0, stream>>>(partBias, static_cast<__nv_bfloat16*>(dbias), out,
                                                                       static_cast<int>(grid.y));
    SOWB_CHECK_CUDA(cudaGetLastError());
  }
  return SOWB_OK;
}

int sow_linear_bwd_dx(const void* dy, const void* dt, const void* W, const void* A, void* dx, int64_t T, int in,
                      int out, int r, int dtype, void* ws, size_t ws_bytes, void* stream_) {
  int rc = check_common("sow_linear_bwd_dx", T, in, out, r, dtype, dy);
  if (rc) return rc;
  SOWB_REQUIRE(dy && dt && A && dx && ws, "sow_linear_bwd_dx: null pointer argument");
  const int r_pad = round_up(r, kBK);
  if (ws_bytes < align256(size_t(in) * r_pad * 2))
    return set_error(SOWB_EWORKSPACE, "sow_linear_bwd_dx: workspace %zu B < required %zu B", ws_bytes,
                     align256(size_t(in) * r_pad * 2));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  __nv_bfloat16* A_pad = static_cast<__nv_bfloat16*>(ws);
  {
    const int64_t n = int64_t(in) * r_pad;
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 4096));
    pack_factor_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(A), A_pad, in, r, r_pad);
    SOWB_CHECK_CUDA(cudaGetLastError());
  }
  // dX = dY . W^T + dt . A_pad^T ;  W (in,out) is K-major for this product: [N=in rows, K=out cols]
  Operand opDT{dt, uint64_t(T), uint64_t(r_pad), uint64_t(r_pad)};
  Operand opApadK{A_pad, uint64_t(in), uint64_t(r_pad), uint64_t(r_pad)};  // [N=in rows, K=r_pad cols]
  if (W != nullptr) {
    Operand opDY{dy, uint64_t(T), uint64_t(out), uint64_t(out)};
    Operand opWK{W, uint64_t(in), uint64_t(out), uint64_t(out)};
    rc = launch_gemm<256, false, false, EPI_BF16_TMA>(opDY, opWK, &opDT, &opApadK, dx, nullptr, in,
                                                      static_cast<int>(T), in, out, r_pad, 1.0f, nullptr, false, stream,
                                                      PROF_GEMM_DX);
  } else {
    rc = launch_gemm<256, false, false, EPI_BF16_TMA>(opDT, opApadK, nullptr, nullptr, dx, nullptr, in,
                                                      static_cast<int>(T), in, r_pad, 0, 1.0f, nullptr, false, stream,
                                                      PROF_GEMM_DX);
  }
  return rc;
}

}  // extern "C"
