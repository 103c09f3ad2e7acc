// LSD radix sort of (Morton key, input index) pairs for inputs of millions of keys (voxel grids and point clouds of the
// cfg2 / cfg5 size class; the reference sorts with thrust::sort_by_key, svo.cu:602,660).  ONE cooperative launch, like
// k_sort (osl_integrate.cu), which stays the sort of mid-sized inputs; what differs at this size:
//   * 8- or 9-bit digits, whichever needs fewer passes for the key width (36-bit keys of a depth-12 tree: 4 passes);
//   * keys only for voxel grids -- their colour index is the SORTED POSITION of the key (quirk Q11), no payload moves;
//   * every CTA owns a contiguous range of tiles (4096 keys, or 2048 pairs); the count phase keeps 8 loads per thread in flight, the
//     scatter phase prefetches the next tile into shared memory (one TMA bulk copy, cp.async.bulk + mbarrier, issued by
//     one thread; cp.async for the ragged last tile) while the current one is ranked, staged in digit order and written
//     out as runs of equal digits;
//   * the cross-CTA prefix is a column scan of the [CTA][digit] count matrix done once (one CTA per digit, one L2 round
//     trip) instead of every CTA summing every other CTA's counts.
// A chained-scan ("one-sweep") variant with per-tile look-back was built first and measured on B200: with 14 k tiles per
// pass and ~300 resident CTAs a tile starts every 0.1 us while one look-back step costs an L2 round trip (0.7 us), the
// chain cannot keep up and the sort took 3.2 - 13 ms (run to run) against 4.2 ms for k_sort.  DESIGN.md section 3.
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "osl_internal.cuh"

namespace cg = cooperative_groups;

#define FULL 0xFFFFFFFFu

namespace {

constexpr int SB_THREADS = 256;
// keys per thread and tile: 16 when only keys move (4096-key tiles), 12 with a payload (3072 pairs); 2 CTAs per SM
template <bool PAY> struct SbCfg {
  static constexpr int ITEMS = PAY ? 12 : 16;
  static constexpr int TILE = SB_THREADS * ITEMS;
};
constexpr int SB_WARPS = SB_THREADS / 32;
constexpr int SB_RADIX = 512;   // most digit values (9-bit digits)
constexpr int SB_MAXG = 768;    // most CTAs (3 rows of the column scan per thread)

struct SortBigArgs {
  u64* kA; u32* pA; u64* kB; u32* pB;
  const int* n_ptr;     // number of keys (device side: the emit stage counted them)
  const int* run_flag;  // NULL, or: sort only when *run_flag != 0 (voxel grids that arrived in Morton order are not sorted)
  u32* hist;            // [grid][SB_RADIX] per-CTA digit counts -> exclusive prefix over the CTAs; then [SB_RADIX] totals
  int passes, bits;
};

template <bool PAY>
constexpr int sb_smem() {
  return SbCfg<PAY>::TILE * 8 * 2 + (PAY ? SbCfg<PAY>::TILE * 4 * 2 : 0) + SB_WARPS * SB_RADIX * 4 + 3 * SB_RADIX * 4 + 64;
}

// exclusive scan of one value per thread over the block (8 warps); total = sum over the block
__device__ __forceinline__ u32 sb_scan(u32 v, u32* s_wsum, int lane, int warp, u32& total) {
  u32 incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // (s_wsum may still be read from the previous scan)
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  u32 woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SB_WARPS; w++) {
    const u32 t = s_wsum[w];
    if (w < warp) woff += t;
    tot += t;
  }
  total = tot;
  return woff + incl - v;
}

__device__ __forceinline__ void sb_cp16(void* smem, const void* gmem, int bytes) {  // bytes in 0..16, the rest is zero-filled
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}

// TMA bulk copy (cp.async.bulk, 1-D) of a whole tile into shared memory: ONE thread issues it, the bytes arrive on an
// mbarrier the CTA waits on -- instead of 4 (+2) cp.async per thread with their address arithmetic.
__device__ __forceinline__ unsigned sb_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_bar_init(u64* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sb_saddr(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
}
__device__ __forceinline__ void sb_bar_expect(u64* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(sb_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_bulk(void* dst, const void* src, unsigned bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(sb_saddr(dst)),
               "l"(src), "r"(bytes), "r"(sb_saddr(bar))
               : "memory");
}
__device__ __forceinline__ void sb_bar_wait(u64* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(sb_saddr(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

template <bool PAY>
__device__ __forceinline__ void sb_prefetch(u64* s_in_k, u32* s_in_p, const u64* kin, const u32* pin, int tile, int n, int tid,
                                            bool bulk, u64* bar) {
  constexpr int SB_TILE = SbCfg<PAY>::TILE;
  const long long g0 = (long long)tile * SB_TILE;
  if (bulk) {  // (a full tile at 16-byte aligned addresses; uniform over the CTA)
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the CTA's reads of the previous tile come first
      sb_bar_expect(bar, SB_TILE * 8 + (PAY ? SB_TILE * 4 : 0));
      sb_bulk(s_in_k, kin + g0, SB_TILE * 8, bar);
      if (PAY) sb_bulk(s_in_p, pin + g0, SB_TILE * 4, bar);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < SB_TILE * 8 / 16 / SB_THREADS; i++) {  // 4 chunks of two keys
    const int c = i * SB_THREADS + tid;
    const long long left = (long long)n - (g0 + 2 * c);
    const int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
    sb_cp16(s_in_k + 2 * c, kin + g0 + 2 * c, bytes);
  }
  if (PAY) {
#pragma unroll
    for (int i = 0; i < SB_TILE * 4 / 16 / SB_THREADS; i++) {  // 2 chunks of four payloads
      const int c = i * SB_THREADS + tid;
      const long long left = (long long)n - (g0 + 4 * c);
      const int bytes = left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0);
      sb_cp16(s_in_p + 4 * c, pin + g0 + 4 * c, bytes);
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
}

template <bool PAY>
__global__ void __launch_bounds__(SB_THREADS, 2) k_sort_big(SortBigArgs A) {
  constexpr int SB_ITEMS = SbCfg<PAY>::ITEMS, SB_TILE = SbCfg<PAY>::TILE;
  cg::grid_group grid = cg::this_grid();
  if (A.run_flag && *A.run_flag == 0) return;  // (uniform over the grid)
  extern __shared__ __align__(16) unsigned char s_raw[];
  u64* s_in_k = reinterpret_cast<u64*>(s_raw);
  u64* s_skey = s_in_k + SB_TILE;
  u32* s_in_p = reinterpret_cast<u32*>(s_skey + SB_TILE);
  u32* s_sval = s_in_p + (PAY ? SB_TILE : 0);
  u32 (*s_whist)[SB_RADIX] = reinterpret_cast<u32 (*)[SB_RADIX]>(s_sval + (PAY ? SB_TILE : 0));
  u32* s_hist = &s_whist[0][0];  // (count phase only: aliases the scatter phase's per-warp counters)
  u32* s_run = &s_whist[0][0] + SB_WARPS * SB_RADIX;
  u32* s_base = s_run + SB_RADIX;
  u32* s_goff = s_base + SB_RADIX;
  u32* s_wsum = s_goff + SB_RADIX;
  u64* s_bar = reinterpret_cast<u64*>(s_wsum + 8);  // mbarrier of the tile copies
  unsigned bar_phase = 0;
  if (threadIdx.x == 0) sb_bar_init(s_bar);
  __syncthreads();

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bid = blockIdx.x;
  const int n = *A.n_ptr;
  const int tiles = (n + SB_TILE - 1) / SB_TILE;
  const int G = gridDim.x;
  const int per = (tiles + G - 1) / G;
  const int t0 = min(tiles, bid * per), t1 = min(tiles, t0 + per);
  const int Gused = per > 0 ? (tiles + per - 1) / per : 0;  // CTAs that own tiles
  const u32 lt = lanemask_lt();
  const int bits = A.bits, RB = 1 << bits;
  const u32 mask = (u32)RB - 1u;
  u32* tot = A.hist + (size_t)G * SB_RADIX;

#ifdef SB_PROFILE
  long long c_count = 0, c_sync1 = 0, c_col = 0, c_scat = 0, c_sync3 = 0, c0 = 0, c1 = 0, d_wait = 0, d_rank = 0, d_digit = 0, d_stage = 0, d_write = 0, d0 = 0, d1 = 0;
#define SBD(x) do { d1 = clock64(); x += d1 - d0; d0 = d1; } while (0)
#define SBP(x) do { c1 = clock64(); x += c1 - c0; c0 = c1; } while (0)
  c0 = clock64();
#else
#define SBP(x)
#define SBD(x)
#endif
  for (int pass = 0; pass < A.passes; pass++) {
    const int shift = bits * pass;
    const u64* kin = (pass & 1) ? A.kB : A.kA;
    const u32* pin = (pass & 1) ? A.pB : A.pA;
    u64* kout = (pass & 1) ? A.kA : A.kB;
    u32* pout = (pass & 1) ? A.pA : A.pB;

    // ---- count: this CTA's digit histogram over its tiles (8 independent loads per thread, warp-aggregated adds)
    for (int d = tid; d < SB_RADIX; d += SB_THREADS) s_hist[d] = 0;
    __syncthreads();
    for (int tile = t0; tile < t1; tile++) {
      const long long g0 = (long long)tile * SB_TILE;
      u64 k[SB_ITEMS];
#pragma unroll
      for (int i = 0; i < SB_ITEMS; i++) {
        const long long j = g0 + i * SB_THREADS + tid;
        k[i] = j < n ? kin[j] : ~0ull;
      }
#pragma unroll
      for (int i = 0; i < SB_ITEMS; i++) {
        const bool ok = (g0 + i * SB_THREADS + tid) < n;
        // (shuffled keys: 32 different digits, conflict-free adds; the high digits of the keys of one surface: the
        // whole warp agrees and adds once)
        const u32 digit = ok ? ((u32)(k[i] >> shift) & mask) : (u32)RB;
        int same;
        __match_all_sync(FULL, digit, &same);
        if (same) {
          if (lane == 0 && ok) atomicAdd(&s_hist[digit], 32u);
        } else if (ok) {
          atomicAdd(&s_hist[digit], 1u);
        }
      }
    }
    __syncthreads();
    for (int d = tid; d < RB; d += SB_THREADS) __stcg(&A.hist[(size_t)bid * SB_RADIX + d], s_hist[d]);
    SBP(c_count);
    grid.sync();
    SBP(c_sync1);

    // ---- column scan: digit `col` over the CTAs (one CTA per digit, all rows in one round trip); in place
    for (int col = bid; col < RB; col += G) {
      u32 v[3], e[3], t3[3];
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int row = q * SB_THREADS + tid;
        v[q] = row < Gused ? __ldcg(&A.hist[(size_t)row * SB_RADIX + col]) : 0u;
      }
#pragma unroll
      for (int q = 0; q < 3; q++) e[q] = sb_scan(v[q], s_wsum, lane, warp, t3[q]);
      u32 before = 0;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int row = q * SB_THREADS + tid;
        if (row < Gused) __stcg(&A.hist[(size_t)row * SB_RADIX + col], before + e[q]);
        before += t3[q];
      }
      if (tid == 0) __stcg(&tot[col], before);
    }
    grid.sync();
    SBP(c_col);

    // ---- bases: first output slot of digit d for this CTA = (all smaller digits) + (same digit in lower CTAs)
    {
      u32 tv[2], ex[2], tt[2];
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int d = q * SB_THREADS + tid;
        tv[q] = d < RB ? __ldcg(&tot[d]) : 0u;
      }
#pragma unroll
      for (int q = 0; q < 2; q++) ex[q] = sb_scan(tv[q], s_wsum, lane, warp, tt[q]);
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int d = q * SB_THREADS + tid;
        if (d < RB) s_run[d] = (q ? tt[0] : 0u) + ex[q] + (bid < Gused ? __ldcg(&A.hist[(size_t)bid * SB_RADIX + d]) : 0u);
      }
    }
    __syncthreads();

    // ---- scatter, tile by tile; the next tile is on its way into shared memory while this one is processed
    const bool aligned = ((reinterpret_cast<uintptr_t>(kin) | reinterpret_cast<uintptr_t>(pin)) & 15) == 0;
    auto full_tile = [&](int tl) { return aligned && (long long)(tl + 1) * SB_TILE <= n; };
    if (t0 < t1) sb_prefetch<PAY>(s_in_k, s_in_p, kin, pin, t0, n, tid, full_tile(t0), s_bar);
    for (int tile = t0; tile < t1; tile++) {
      const int base = tile * SB_TILE + warp * (32 * SB_ITEMS);
#ifdef SB_PROFILE
      d0 = clock64();
#endif
      if (full_tile(tile)) {
        sb_bar_wait(s_bar, bar_phase);
        bar_phase ^= 1u;
      } else {
        asm volatile("cp.async.wait_group 0;\n" ::);
      }
      __syncthreads();
      SBD(d_wait);
      u64 key[SB_ITEMS];
      u32 val[SB_ITEMS], rank[SB_ITEMS];
#pragma unroll
      for (int i = 0; i < SB_ITEMS; i++) {  // warp-striped: (warp, item, lane) order == input order (stability)
        key[i] = s_in_k[warp * (32 * SB_ITEMS) + i * 32 + lane];
        if (PAY) val[i] = s_in_p[warp * (32 * SB_ITEMS) + i * 32 + lane];
      }
      for (int d = lane; d < RB; d += 32) s_whist[warp][d] = 0;
      __syncthreads();
      if (tile + 1 < t1) sb_prefetch<PAY>(s_in_k, s_in_p, kin, pin, tile + 1, n, tid, full_tile(tile + 1), s_bar);
#pragma unroll
      for (int i = 0; i < SB_ITEMS; i++) {
        const bool ok = (base + i * 32 + lane) < n;
        // Rank among the items of this warp with the same digit, in lane order.  Optimistic: every lane adds 1 to the
        // warp's counter of its digit; a lane whose counter moved by exactly 1 was alone in this row (94 % of the
        // lanes for shuffled 9-bit digits) and its rank is the old count.  Only the lanes that met a peer sort it out
        // among themselves with match.any -- over 2-4 lanes, not 32 distinct values (match.any over the whole warp takes
        // time proportional to the number of distinct values; one ballot per digit bit costs ~85 instructions per item).
        const u32 digit = (u32)(key[i] >> shift) & mask;
        u32 before = 0, after = 0;
        if (ok) before = s_whist[warp][digit];
        __syncwarp();
        if (ok) atomicAdd(&s_whist[warp][digit], 1u);
        __syncwarp();
        if (ok) after = s_whist[warp][digit];
        rank[i] = before;
        const bool met = ok && after != before + 1u;
        const u32 crowd = __ballot_sync(FULL, met);
        if (met) {
          const u32 peers = __match_any_sync(crowd, digit);
          rank[i] = before + __popc(peers & lt);
        }
        __syncwarp();
      }
      __syncthreads();
      SBD(d_rank);
      {  // digit d: exclusive scan over the warps, advance the running base; tile-local order of the digits
        u32 sum[2], lex[2], tt[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int d = q * SB_THREADS + tid;
          sum[q] = 0;
          if (d < RB) {
#pragma unroll
            for (int w = 0; w < SB_WARPS; w++) {
              const u32 v = s_whist[w][d];
              s_whist[w][d] = sum[q];
              sum[q] += v;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 2; q++) lex[q] = sb_scan(sum[q], s_wsum, lane, warp, tt[q]);
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int d = q * SB_THREADS + tid;
          if (d < RB) {
            const u32 l = (q ? tt[0] : 0u) + lex[q];
            const u32 b = s_run[d];
            s_run[d] = b + sum[q];
            s_base[d] = l;      // first tile-local slot of the digit
            s_goff[d] = b - l;  // global position = s_goff[digit] + tile-local slot
          }
        }
      }
      __syncthreads();
      SBD(d_digit);
      // stage the tile in digit order, then write it out linearly: runs of equal digits go to consecutive addresses
#pragma unroll
      for (int i = 0; i < SB_ITEMS; i++) {
        if ((base + i * 32 + lane) < n) {
          const u32 digit = (u32)(key[i] >> shift) & mask;
          const u32 lp = s_base[digit] + s_whist[warp][digit] + rank[i];
          s_skey[lp] = key[i];
          if (PAY) s_sval[lp] = val[i];
        }
      }
      __syncthreads();
      SBD(d_stage);
      const int cnt = min(SB_TILE, n - tile * SB_TILE);
      for (int l = tid; l < cnt; l += SB_THREADS) {
        const u64 k = s_skey[l];
        const u32 pos = s_goff[(u32)(k >> shift) & mask] + (u32)l;
        kout[pos] = k;
        if (PAY) pout[pos] = s_sval[l];
      }
      SBD(d_write);
    }
    SBP(c_scat);
    grid.sync();
    SBP(c_sync3);
  }
#ifdef SB_PROFILE
  if (tid == 0 && (bid == 0 || bid == G / 2))
    printf("k_sort_big cta %d: count %.0f us, sync %.0f, colscan+sync %.0f, scatter %.0f, end sync %.0f (tiles %d)\n", bid,
           c_count / 1965.0, c_sync1 / 1965.0, c_col / 1965.0, c_scat / 1965.0, c_sync3 / 1965.0, t1 - t0);
  if (tid == 0 && (bid == 0 || bid == G / 2))
    printf("   scatter: wait %.0f us, rank %.0f, digit %.0f, stage %.0f, write %.0f\n", d_wait / 1965.0, d_rank / 1965.0,
           d_digit / 1965.0, d_stage / 1965.0, d_write / 1965.0);
#endif
}

}  // namespace

void osl_sort_big_free(OslSortWs* ws) {
  if (ws->hist) cudaFree(ws->hist);
  ws->hist = nullptr; ws->grid = 0;
}

osl_status osl_sort_big_reserve(OslSortWs* ws) {
  if (ws->hist) return OSL_OK;
  OSL_CUDA(cudaMalloc(&ws->hist, ((size_t)SB_MAXG + 1) * SB_RADIX * sizeof(u32)));
  ws->grid = SB_MAXG;
  return OSL_OK;
}

// digit width / number of passes for a key of key_bits bits
int osl_sort_big_passes(int key_bits, int* bits_out) {
  if (key_bits < 1) key_bits = 1;
  const int passes = (key_bits + 8) / 9;
  const int bits = (key_bits + passes - 1) / passes;
  if (bits_out) *bits_out = bits;
  return passes;
}

// Sorts the pairs in (kA, pA) by the low key_bits key bits; the result is in (kB, pB) when the number of passes
// (osl_sort_big_passes) is odd, else in (kA, pA).  with_pay = false moves keys only.  grid_cap bounds the cooperative grid
// (other cooperative kernels of the frame pipeline may be resident).
osl_status osl_sort_big(OslSortWs* ws, u64* kA, u32* pA, u64* kB, u32* pB, const int* d_n, const int* d_run_flag,
                        long long n_upper, int key_bits, bool with_pay, int grid_cap, cudaStream_t st) {
  if (n_upper <= 0) return OSL_OK;
  // (function attributes and occupancy are per device; one process may drive several)
  static int s_occ_pay[64], s_occ_keys[64], s_sms[64];
  static bool s_have[64] = {false};
  int dev = 0;
  OSL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return OSL_ERR_INVALID;
  if (!s_have[dev]) {
    OSL_CUDA(cudaDeviceGetAttribute(&s_sms[dev], cudaDevAttrMultiProcessorCount, dev));
    OSL_CUDA(cudaFuncSetAttribute((const void*)k_sort_big<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb_smem<true>()));
    OSL_CUDA(cudaFuncSetAttribute((const void*)k_sort_big<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb_smem<false>()));
    OSL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_occ_pay[dev], (const void*)k_sort_big<true>, SB_THREADS, sb_smem<true>()));
    OSL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_occ_keys[dev], (const void*)k_sort_big<false>, SB_THREADS, sb_smem<false>()));
    s_have[dev] = true;
  }
  const int occ_pay = s_occ_pay[dev], occ_keys = s_occ_keys[dev], sms = s_sms[dev];
  const int occ = with_pay ? occ_pay : occ_keys;
  if (occ < 1) return OSL_ERR_CUDA;
  long long grid = (long long)occ * sms;
  if (grid > SB_MAXG) grid = SB_MAXG;
  if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
  const int tile_keys = with_pay ? SbCfg<true>::TILE : SbCfg<false>::TILE;
  const long long tiles = (n_upper + tile_keys - 1) / tile_keys;
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  osl_status rr = osl_sort_big_reserve(ws);
  if (rr) return rr;
  SortBigArgs A;
  A.kA = kA; A.pA = pA; A.kB = kB; A.pB = pB; A.n_ptr = d_n; A.run_flag = d_run_flag; A.hist = ws->hist;
  A.passes = osl_sort_big_passes(key_bits, &A.bits);
  void* args[] = {&A};
  if (with_pay)
    OSL_CUDA(cudaLaunchCooperativeKernel((void*)k_sort_big<true>, dim3((unsigned)grid), dim3(SB_THREADS), args, sb_smem<true>(), st));
  else
    OSL_CUDA(cudaLaunchCooperativeKernel((void*)k_sort_big<false>, dim3((unsigned)grid), dim3(SB_THREADS), args, sb_smem<false>(), st));
  OSL_LAUNCHED(1);
  return OSL_OK;
}
