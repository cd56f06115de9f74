// Grouped SoW merge:  W <- W_prev + scale * A . B   for every SoW linear of a model in ONE launch.
//
// Replaces the dense branch of SoWLinear.accumulate (tn_gradient/layer/sow.py:131-134,140,151-153), which in the
// reference materialises >= 4 in x out temporaries per layer (stack, sum, scale, add) through separate kernels.
//
// The op is HBM-bound (2 bytes read + 2 bytes written per element) but at r = 50 it needs 25 flop/byte, above what
// the CUDA cores sustain at full bandwidth, so the rank-r product runs on tcgen05 (one 128x128x64 UMMA block per
// tile) and everything else is a streaming read-modify-write pipeline:
//
//   warps 4..7 (producer group): thread 0 issues TMA loads of the W_prev tile (2 x [128 rows x 64 cols], swizzle
//               128B) and of the B tile ([64 k-rows x 128 cols], rows >= r zero-filled by TMA bounds); all 128
//               threads gather the A tile (pitch r*2 bytes is not TMA-legal) into a zero-padded swizzled smem tile
//   warp  8   : TMEM allocation; lane 0 issues the UMMA and commits to the accumulator barrier
//   warps 0..3 (epilogue group): tcgen05.ld accumulator row, add in place onto the W tile in smem (conflict-free
//               16-byte swizzled accesses), TMA-store the tile back (coalesced, clipped at the matrix edge)
//
// 3 smem slots (W 32 KB + B 16 KB + A 16 KB) keep >= 2 tiles of loads in flight per SM; TMEM double-buffers the
// accumulator.  Tiles of all layers are flattened into one index space walked persistently by <= #SM CTAs.
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <string.h>
#include <vector>

namespace sowb {

constexpr int kMgBM = 128, kMgBN = 128, kMgSlots = 3;
constexpr int kMgWBytes = kMgBM * kMgBN * 2;        // 32 KB
constexpr int kMgBBytes = 64 * kMgBN * 2;           // 16 KB
constexpr int kMgABytes = kMgBM * 64 * 2;           // 16 KB
constexpr int kMgSlotBytes = kMgWBytes + kMgBBytes + kMgABytes;
constexpr int kMgMaxEntries = 2048;   // per launch (host splits longer tables)
constexpr int kMgSmemTotal = 1024 + kMgSlots * kMgSlotBytes + 256 + kMgMaxEntries * 4;
constexpr int kMgThreads = 288;
constexpr uint32_t kMgTmemCols = 256;

struct alignas(128) MergeDevEntry {
  CUtensorMap tmWin;   // W_prev loads  (box 64 cols x 128 rows)
  CUtensorMap tmWout;  // W stores
  CUtensorMap tmB;     // B[r, out] as MN-major operand (box 64 cols x 64 rows)
  const __nv_bfloat16* A;
  int lda;             // row pitch of A in elements (= full rank)
  int in, out, r;      // r = rank chunk handled by this entry (<= 64)
  float scale;
  int m_tiles, n_tiles, tile_begin;
  int has_prev;
};

// Tiles are visited in increasing order by every role, so the entry index only ever moves forward: walk it.
// `ends` (smem) holds tile_begin + #tiles of every entry, staged once per CTA.
__device__ __forceinline__ int advance_entry(const int* __restrict__ ends, int ei, int tile) {
  while (tile >= ends[ei]) ++ei;
  return ei;
}

__global__ void __launch_bounds__(kMgThreads, 1)
sow_merge_kernel(const MergeDevEntry* __restrict__ tab, int n_entries, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMgSlots * kMgSlotBytes);
  uint64_t* full_bar = bars;                    // [slots] 128 A-writers + 1 expect_tx
  uint64_t* empty_bar = bars + kMgSlots;        // [slots] 1
  uint64_t* tfull_bar = bars + 2 * kMgSlots;    // [2]
  uint64_t* tempty_bar = bars + 2 * kMgSlots + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMgSlots + 4);
  int* ends = reinterpret_cast<int*>(smem + kMgSlots * kMgSlotBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < n_entries; i += kMgThreads) ends[i] = tab[i].tile_begin + tab[i].m_tiles * tab[i].n_tiles;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMgSlots; ++i) {
      mbar_init(&full_bar[i], 129);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, kMgTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ===================== producer group =====================
    const int pt = threadIdx.x - 128;  // 0..127
    const MergeDevEntry* last_e = nullptr;
    int slot = 0, ei = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ei = advance_entry(ends, ei, tile);
      const MergeDevEntry* e = &tab[ei];
      const int local = tile - e->tile_begin;
      const int m0 = (local / e->n_tiles) * kMgBM;
      const int n0 = (local % e->n_tiles) * kMgBN;
      uint8_t* sW = smem + slot * kMgSlotBytes;
      uint8_t* sB = sW + kMgWBytes;
      uint8_t* sA = sB + kMgBBytes;
      // issue the A loads BEFORE waiting for the slot: their latency overlaps the wait
      const __nv_bfloat16* A = e->A;
      const int r = e->r, lda = e->lda, rows = min(kMgBM, e->in - m0);
      const bool vec = (lda == r) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((r & 7) == 0 || (m0 * r) % 8 == 0);
      uint4 av[8];
      const int nvec = (rows * r) >> 3;                 // whole 8-element groups of the contiguous [rows x r] block
      if (vec) {
        const uint4* src = reinterpret_cast<const uint4*>(A + static_cast<int64_t>(m0) * r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int v = pt + 128 * j;
          if (v < nvec) av[j] = __ldg(src + v);
        }
      }
      mbar_wait(&empty_bar[slot], phase ^ 1);
      if (pt == 0) {
        if (e != last_e) {
          tma_acquire_desc(&e->tmWin);
          tma_acquire_desc(&e->tmB);
          last_e = e;
        }
        mbar_expect_tx(&full_bar[slot], kMgBBytes + (e->has_prev ? kMgWBytes : 0));
        if (e->has_prev) {
          tma_load_2d(sW, &e->tmWin, &full_bar[slot], n0, m0);
          tma_load_2d(sW + 16384, &e->tmWin, &full_bar[slot], n0 + 64, m0);
        }
        tma_load_2d(sB, &e->tmB, &full_bar[slot], n0, 0);
        tma_load_2d(sB + 8192, &e->tmB, &full_bar[slot], n0 + 64, 0);
      }
      // A[m0 : m0+128, 0 : r] -> K-major swizzled tile, zero padded to 64 columns
      auto put = [&](int row, int col, __nv_bfloat16 v) {
        *reinterpret_cast<__nv_bfloat16*>(sA + row * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2) = v;
      };
      if (vec) {
        // 1) zero the padding columns r..63 (and rows beyond the matrix edge), 16 B at a time where possible
        for (int idx = pt; idx < kMgBM * 8; idx += 128) {
          const int row = idx >> 3, chunk = idx & 7;
          if (row >= rows || chunk * 8 >= r)
            *reinterpret_cast<uint4*>(sA + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(0, 0, 0, 0);
          else if (chunk * 8 + 8 > r)
            for (int c = r; c < chunk * 8 + 8; ++c) put(row, c, __float2bfloat16(0.f));
        }
        // 2) scatter the prefetched 16-byte groups (a group may straddle two rows when r % 8 != 0)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int v = pt + 128 * j;
          if (v < nvec) {
            const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&av[j]);
            int e0 = v * 8;
            int row = e0 / r, col = e0 - row * r;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              put(row, col, h[q]);
              if (++col == r) {
                col = 0;
                ++row;
              }
            }
          }
        }
        for (int e1 = nvec * 8 + pt; e1 < rows * r; e1 += 128) {   // tail elements of the block
          const int row = e1 / r, col = e1 - row * r;
          put(row, col, A[static_cast<int64_t>(m0) * r + e1]);
        }
      } else {
        const int col = pt & 63;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) {
          const int row = (pt >> 6) + 2 * j;
          __nv_bfloat16 v = __float2bfloat16(0.f);
          if (col < r && row < rows) v = A[static_cast<int64_t>(m0 + row) * lda + col];
          put(row, col, v);
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[slot]);
      if (++slot == kMgSlots) {
        slot = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 8 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(1, kMgBM, kMgBN, 0, 1);
    int slot = 0, acc = 0, ei = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ei = advance_entry(ends, ei, tile);
      const MergeDevEntry* e = &tab[ei];
      const int ksteps = (e->r + 15) >> 4;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[slot], phase);
      tc_fence_after();
      const uint32_t sW = smem_u32(smem + slot * kMgSlotBytes);
      const uint32_t sB = sW + kMgWBytes, sA = sB + kMgBBytes;
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t ad = make_smem_desc(sA + k * 32, 16, 1024);
        const uint64_t bd = make_smem_desc(sB + k * 2048, 8192, 1024);
        umma_bf16(tmem_base + acc * kMgBN, ad, bd, idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(&tfull_bar[acc]);
      if (++slot == kMgSlots) {
        slot = 0;
        phase ^= 1;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue group =====================
    const int et = threadIdx.x;
    const int row = warp * 32 + lane;
    int slot = 0, acc = 0, ei = 0;
    uint32_t phase = 0, acc_phase = 0;
    int prev_slot = -1;
    const MergeDevEntry* last_e = nullptr;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ei = advance_entry(ends, ei, tile);
      const MergeDevEntry* e = &tab[ei];
      const int local = tile - e->tile_begin;
      const int m0 = (local / e->n_tiles) * kMgBM;
      const int n0 = (local % e->n_tiles) * kMgBN;
      const float scale = e->scale;
      const bool has_prev = e->has_prev != 0;
      uint8_t* sW = smem + slot * kMgSlotBytes;
      mbar_wait(&full_bar[slot], phase);  // W tile landed (acquire on the TMA barrier)
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * kMgBN;
#pragma unroll
      for (int q = 0; q < kMgBN / 32; ++q) {
        uint32_t v[32];
        tmem_ld32(taddr + q * 32, v);
        tmem_ld_wait();
        uint8_t* box = sW + (q >> 1) * 16384 + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = (q & 1) * 4 + c;
          uint4* p = reinterpret_cast<uint4*>(box + ((chunk ^ (row & 7)) << 4));
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = scale * __uint_as_float(v[c * 8 + j]);
          if (has_prev) {
            const uint4 old = *p;
            const uint32_t w[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
              f[2 * j] += __low2float(b);
              f[2 * j + 1] += __high2float(b);
            }
          }
          uint4 pk;
          pk.x = pack_bf16x2(f[0], f[1]);
          pk.y = pack_bf16x2(f[2], f[3]);
          pk.z = pack_bf16x2(f[4], f[5]);
          pk.w = pack_bf16x2(f[6], f[7]);
          *p = pk;
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      fence_proxy_async_smem();
      named_barrier_sync(1, 128);
      if (et == 0) {
        if (e != last_e) {
          tma_acquire_desc(&e->tmWout);
          last_e = e;
        }
        tma_store_2d(&e->tmWout, sW, n0, m0);
        if (n0 + 64 < e->out) tma_store_2d(&e->tmWout, sW + 16384, n0 + 64, m0);
        tma_store_commit();
        if (prev_slot >= 0) {
          tma_store_wait_read<1>();  // the previous tile's store has finished reading its slot
          mbar_arrive(&empty_bar[prev_slot]);
        }
      }
      prev_slot = slot;
      if (++slot == kMgSlots) {
        slot = 0;
        phase ^= 1;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (et == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(tmem_base, kMgTmemCols);
}

}  // namespace sowb

using namespace sowb;

extern "C" {

size_t sow_merge_table_stride(void) {
  // one device entry per 64-wide rank chunk; callers size the table for ceil(r/64) chunks per layer
  return sizeof(MergeDevEntry);
}

int sow_merge_grouped(const sowb_merge_entry* entries, int n, int dtype, void* table_dev, size_t table_bytes,
                      void* stream_) {
  if (n <= 0) return SOWB_OK;
  if (dtype != SOWB_BF16) return set_error(SOWB_EINVAL, "sow_merge_grouped: only SOWB_BF16 is implemented");
  SOWB_REQUIRE(entries != nullptr && table_dev != nullptr, "sow_merge_grouped: null pointer argument");
  SOWB_REQUIRE((reinterpret_cast<uintptr_t>(table_dev) & 127) == 0, "sow_merge_grouped: table_dev must be 128-byte aligned");
  int rc = require_sm100();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int max_chunks = 1;
  for (int i = 0; i < n; ++i) {
    const sowb_merge_entry& e = entries[i];
    SOWB_REQUIRE(e.W && e.A && e.B, "sow_merge_grouped: entry %d has a null W/A/B pointer", i);
    SOWB_REQUIRE(e.in > 0 && e.out > 0 && e.r > 0, "sow_merge_grouped: entry %d has a non-positive dimension", i);
    SOWB_REQUIRE(e.out % 8 == 0, "sow_merge_grouped: entry %d: out=%d must be a multiple of 8", i, e.out);
    max_chunks = std::max(max_chunks, ceil_div(e.r, 64));
  }
  if (table_bytes < size_t(n) * sizeof(MergeDevEntry))
    return set_error(SOWB_EWORKSPACE, "sow_merge_grouped: table %zu B < required %zu B", table_bytes,
                     size_t(n) * sizeof(MergeDevEntry));
  SOWB_CHECK_CUDA(cudaFuncSetAttribute(sow_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMgSmemTotal));
  // ranks above 64 are applied as successive 64-wide chunks (W_prev = W after the first chunk)
  for (int chunk = 0; chunk < max_chunks; ++chunk) {
    std::vector<MergeDevEntry> host;
    host.reserve(n);
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
      const sowb_merge_entry& e = entries[i];
      const int r0 = chunk * 64;
      if (r0 >= e.r) continue;
      MergeDevEntry d;
      memset(&d, 0, sizeof(d));
      const void* prev = (chunk == 0) ? e.W_prev : e.W;
      d.has_prev = prev != nullptr;
      rc = make_tensor_map_2d(&d.tmWout, e.W, e.out, e.in, uint64_t(e.out) * 2, 64, kMgBM, 2);
      if (rc) return rc;
      d.tmWin = d.tmWout;
      if (prev != nullptr && prev != e.W) {
        rc = make_tensor_map_2d(&d.tmWin, prev, e.out, e.in, uint64_t(e.out) * 2, 64, kMgBM, 2);
        if (rc) return rc;
      }
      const int rc_rows = std::min(64, e.r - r0);
      const __nv_bfloat16* Bp = static_cast<const __nv_bfloat16*>(e.B) + size_t(r0) * e.out;
      rc = make_tensor_map_2d(&d.tmB, Bp, e.out, rc_rows, uint64_t(e.out) * 2, 64, 64, 2);
      if (rc) return rc;
      d.A = static_cast<const __nv_bfloat16*>(e.A) + r0;
      d.lda = e.r;
      d.in = e.in;
      d.out = e.out;
      d.r = rc_rows;
      d.scale = e.scale;
      d.m_tiles = ceil_div(e.in, kMgBM);
      d.n_tiles = ceil_div(e.out, kMgBN);
      d.tile_begin = tiles;
      tiles += d.m_tiles * d.n_tiles;
      host.push_back(d);
    }
    if (host.empty()) break;
    if (host.size() > size_t(kMgMaxEntries))
      return set_error(SOWB_EINVAL, "sow_merge_grouped: at most %d entries per call", kMgMaxEntries);
    SOWB_CHECK_CUDA(cudaMemcpyAsync(table_dev, host.data(), host.size() * sizeof(MergeDevEntry),
                                    cudaMemcpyHostToDevice, stream));
    const int sms = num_sms();
    const int grid = tiles < sms ? tiles : sms;
    double bytes = 0;   // algorithmic: W read (if any) + W write + A + B
    for (const auto& d : host) bytes += 2.0 * d.in * d.out * (1 + d.has_prev) + 2.0 * d.r * (double(d.in) + d.out);
    ProfileScope prof(stream, PROF_MERGE, bytes);
    sow_merge_kernel<<<grid, kMgThreads, kMgSmemTotal, stream>>>(static_cast<const MergeDevEntry*>(table_dev),
                                                                 static_cast<int>(host.size()), tiles);
    SOWB_CHECK_CUDA(cudaGetLastError());
  }
  return SOWB_OK;
}

}  // extern "C"
