// Stable LSD radix sort of (key, value) pairs on the device -- the engine's own sorter.
//
// It stands behind every reordering of the sparse data: CSR -> CSC (SMatrix::transpose, reference
// src/util/Smatrix.h:155-185: equal columns must keep their rows ascending, i.e. the sort must be STABLE), the per-batch
// CSC of the minibatch trainers, the row order of the dense ALS layout and the |score| order of AUC
// (src/core/Evaluation.h:56-78).
//
// Structure (8-bit digits, one sweep per digit):
//   radix_hist_kernel   one read of the keys -> the digit histograms of ALL passes (warp-aggregated shared-memory counts)
//   radix_scan_kernel   exclusive scan of each pass's 256 bins -> global bin bases
//   radix_pass_kernel   per pass, ONE kernel reads a tile of keys/values once and writes it once:
//                         * ranks inside the tile are stable: every warp owns a contiguous slice, items are taken in order and
//                           lanes of one item are ordered by `match_any` masks (rank = earlier items' count + lower lanes),
//                         * the tile's position among the earlier tiles comes from a chained scan with decoupled look-back
//                           (tile ids are handed out by an atomic ticket, so a tile only ever waits for tiles that already run),
//                         * keys/values are regrouped by digit in shared memory so the global writes are runs, not single words.
// HBM traffic per pass = read + write of the pairs (16 B per 32-bit pair); nothing is re-read.
#include "common.cuh"

namespace fmwr {

namespace {

constexpr int RS_BITS = 8;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 8;
#ifndef RS_NAP
#define RS_NAP 20
#endif

template <class K> struct RsTile { enum { ITEMS = sizeof(K) == 4 ? 16 : 12, SIZE = RS_THREADS * ITEMS }; };

// ---- histograms of every pass in one read of the keys ----------------------------------------------------------------
template <class K>
__global__ void __launch_bounds__(256) radix_hist_kernel(const K* __restrict__ keys, int64_t n, int passes, int bits, uint32_t* __restrict__ hist /*[passes][256]*/)
{
  __shared__ uint32_t sh[RS_MAX_PASSES * RS_BINS];
  for (int i = threadIdx.x; i < passes * RS_BINS; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // Every lane takes HU consecutive keys per trip (one 16-byte load for 32-bit keys).  A digit that is the same across the warp's
  // lanes (the high digits of sorted-ish keys: the batch number of consecutive rows) is counted by one lane; anything else goes
  // to the shared-memory bins lane by lane -- low digits are near-random, 32 lanes into 256 bins rarely collide.
  constexpr int HU = 16 / (int)sizeof(K);
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp_global * 32 * HU; base < n; base += n_warps * 32 * HU) {
    const int64_t i0 = base + (int64_t)lane * HU;
    K key[HU];
    bool ok[HU];
    if (i0 + HU <= n && (reinterpret_cast<uintptr_t>(keys + i0) & 15) == 0) {
      const uint4 q4 = *reinterpret_cast<const uint4*>(keys + i0);
      if (sizeof(K) == 4) { key[0] = (K)q4.x; key[1 % HU] = (K)q4.y; key[2 % HU] = (K)q4.z; key[3 % HU] = (K)q4.w; }
      else { key[0] = (K)(((uint64_t)q4.y << 32) | q4.x); key[1 % HU] = (K)(((uint64_t)q4.w << 32) | q4.z); }
#pragma unroll
      for (int u = 0; u < HU; ++u) ok[u] = true;
    } else {
#pragma unroll
      for (int u = 0; u < HU; ++u) { ok[u] = i0 + u < n; key[u] = ok[u] ? keys[i0 + u] : K(0); }
    }
    for (int ps = 0; ps < passes; ++ps) {
      const int shift = ps * RS_BITS;
      const int nb = min(RS_BITS, bits - shift);
      const K dm = (K)((1u << nb) - 1u);
      uint32_t d[HU];
#pragma unroll
      for (int u = 0; u < HU; ++u) d[u] = ok[u] ? (uint32_t)((key[u] >> shift) & dm) : (uint32_t)RS_BINS;
      bool same = true;
#pragma unroll
      for (int u = 1; u < HU; ++u) same &= d[u] == d[0];
      const uint32_t d00 = __shfl_sync(0xffffffffu, d[0], 0);
      if (__all_sync(0xffffffffu, same && d[0] == d00)) {
        if (lane == 0 && d00 < (uint32_t)RS_BINS) atomicAdd(&sh[ps * RS_BINS + d00], 32u * HU);
      } else {
#pragma unroll
        for (int u = 0; u < HU; ++u) if (ok[u]) atomicAdd(&sh[ps * RS_BINS + d[u]], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * RS_BINS; i += blockDim.x) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// exclusive scan of each pass's bins, in place (one warp per pass; counts < 2^32 because n < 2^32)
__global__ void radix_scan_kernel(uint32_t* __restrict__ hist, int passes)
{
  const int ps = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (ps >= passes) return;
  uint32_t* h = hist + ps * RS_BINS;
  uint32_t run = 0;
  for (int c = 0; c < RS_BINS / 32; ++c) {
    const uint32_t x = h[c * 32 + lane];
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    h[c * 32 + lane] = run + inc - x;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
}

// look-back word: the two top bits say what the low bits hold (0 nothing yet, 1 this tile's count, 2 inclusive count up to this tile)
template <class W> struct Lb;
template <> struct Lb<uint32_t> {
  static constexpr uint32_t LOCAL = 1u << 30, INCL = 2u << 30, MASK = (1u << 30) - 1u;
};
template <> struct Lb<uint64_t> {
  static constexpr uint64_t LOCAL = 1ull << 62, INCL = 2ull << 62, MASK = (1ull << 62) - 1ull;
};

template <class K, class W, bool IOTA>
__global__ void __launch_bounds__(RS_THREADS, 4)
radix_pass_kernel(const K* __restrict__ kin, K* __restrict__ kout, const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout, int64_t n,
                  int shift, int nbits, const uint32_t* __restrict__ gbase /*[256] exclusive bin bases of this pass*/,
                  W* __restrict__ state /*[tiles][256], zeroed*/, uint32_t* __restrict__ ticket)
{
  constexpr int ITEMS = RsTile<K>::ITEMS;
  constexpr int TILE = RsTile<K>::SIZE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K* skeys = reinterpret_cast<K*>(smem_raw);                                    // [TILE]
  uint32_t* svals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(K) * TILE);   // [TILE]
  uint32_t* whist = svals + TILE;                                               // [RS_WARPS][RS_BINS + 1]   (+1: the bin of out-of-range lanes)
  uint32_t* bin_excl = whist + RS_WARPS * (RS_BINS + 1);                        // [RS_BINS] first local position of a bin
  uint32_t* bin_goff = bin_excl + RS_BINS;                                      // [RS_BINS] global position of the bin's first element - local position
  __shared__ uint32_t tile_s;
  __shared__ uint32_t warp_tot[RS_BINS / 32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) tile_s = atomicAdd(ticket, 1u);
  for (int i = tid; i < RS_WARPS * (RS_BINS + 1); i += RS_THREADS) whist[i] = 0u;
  __syncthreads();
  const uint32_t tile = tile_s;
  const int64_t tile_base = (int64_t)tile * TILE;
  const int tile_n = (n - tile_base) < (int64_t)TILE ? (int)(n - tile_base) : TILE;
  const K dmask = (K)((1u << nbits) - 1u);

  // ---- load (warp w owns elements [w * 32 * ITEMS, (w+1) * 32 * ITEMS) of the tile, item-major inside) and rank
  K key[ITEMS];
  uint32_t val[ITEMS];
  uint32_t rank[ITEMS];
  const int wbase = wid * 32 * ITEMS;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int li = wbase + i * 32 + lane;
    key[i] = K(0); val[i] = 0u;
    if (li < tile_n) {
      key[i] = kin[tile_base + li];
      val[i] = IOTA ? (uint32_t)(tile_base + li) : vin[tile_base + li];
    }
  }
  uint32_t* myh = whist + wid * (RS_BINS + 1);
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int li = wbase + i * 32 + lane;
    const uint32_t d = li < tile_n ? (uint32_t)((key[i] >> shift) & dmask) : (uint32_t)RS_BINS;
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(m) - 1;
    uint32_t pre = 0u;
    if (lane == leader) { pre = myh[d]; myh[d] = pre + (uint32_t)__popc(m); }
    pre = __shfl_sync(0xffffffffu, pre, leader);
    rank[i] = pre + (uint32_t)__popc(m & lt);
    __syncwarp();
  }
  __syncthreads();

  // ---- per bin: prefix over the warps (whist becomes the offset of the warp inside the bin), tile count, local exclusive scan
  uint32_t cnt = 0u;
  if (tid < RS_BINS) {
#pragma unroll 4
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = whist[w * (RS_BINS + 1) + tid];
      whist[w * (RS_BINS + 1) + tid] = cnt;
      cnt += c;
    }
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[wid] = inc;
    bin_excl[tid] = inc - cnt;                      // within its warp of bins for now
  }
  __syncthreads();
  if (tid < RS_BINS) {
    uint32_t off = 0u;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    const uint32_t ex = bin_excl[tid] + off;
    bin_excl[tid] = ex;
    // ---- chained scan with look-back: how many elements of this bin sit in earlier tiles
    W* mine = state + (size_t)tile * RS_BINS + tid;
    volatile W* vstate = state;
    uint32_t before = 0u;
    if (tile == 0) {
      *reinterpret_cast<volatile W*>(mine) = (W)cnt | Lb<W>::INCL;
    } else {
      *reinterpret_cast<volatile W*>(mine) = (W)cnt | Lb<W>::LOCAL;
      int64_t look = (int64_t)tile - 1;
      W acc = 0;
      for (;;) {
        W s = vstate[(size_t)look * RS_BINS + tid];
        while ((s & ~Lb<W>::MASK) == 0) { if (RS_NAP) __nanosleep(RS_NAP); s = vstate[(size_t)look * RS_BINS + tid]; }
        acc += s & Lb<W>::MASK;
        if (s & Lb<W>::INCL) break;
        --look;
      }
      *reinterpret_cast<volatile W*>(mine) = (acc + (W)cnt) | Lb<W>::INCL;
      before = (uint32_t)acc;
    }
    bin_goff[tid] = gbase[tid] + before - ex;
  }
  __syncthreads();

  // ---- regroup by digit in shared memory
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int li = wbase + i * 32 + lane;
    if (li < tile_n) {
      const uint32_t d = (uint32_t)((key[i] >> shift) & dmask);
      const uint32_t pos = bin_excl[d] + myh[d] + rank[i];
      skeys[pos] = key[i];
      svals[pos] = val[i];
    }
  }
  __syncthreads();
  // ---- write runs
  for (int j = tid; j < tile_n; j += RS_THREADS) {
    const K k2 = skeys[j];
    const uint32_t d = (uint32_t)((k2 >> shift) & dmask);
    const uint32_t pos = bin_goff[d] + (uint32_t)j;
    kout[pos] = k2;
    vout[pos] = svals[j];
  }
}

template <class K>
size_t pass_smem_bytes()
{
  return (sizeof(K) + 4) * (size_t)RsTile<K>::SIZE + 4 * ((size_t)RS_WARPS * (RS_BINS + 1) + 2 * RS_BINS);
}

template <class K, class W>
void launch_pass(fmwr_ctx* ctx, const K* kin, K* kout, const uint32_t* vin, uint32_t* vout, int64_t n, int shift, int nbits, const uint32_t* gbase,
                 W* state, uint32_t* ticket, int tiles)
{
  const size_t smem = pass_smem_bytes<K>();
  if (vin == nullptr) {
    FMWR_CUDA(cudaFuncSetAttribute(radix_pass_kernel<K, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FMWR_LAUNCH(ctx, (radix_pass_kernel<K, W, true>), tiles, RS_THREADS, smem, kin, kout, vin, vout, n, shift, nbits, gbase, state, ticket);
  } else {
    FMWR_CUDA(cudaFuncSetAttribute(radix_pass_kernel<K, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FMWR_LAUNCH(ctx, (radix_pass_kernel<K, W, false>), tiles, RS_THREADS, smem, kin, kout, vin, vout, n, shift, nbits, gbase, state, ticket);
  }
}

// key_out / val_out receive the result; key_in is only read; val_in == nullptr stands for the identity 0, 1, 2, ...
template <class K>
void radix_sort_pairs(fmwr_ctx* ctx, const K* key_in, K* key_out, const uint32_t* val_in, uint32_t* val_out, int64_t n, int bits)
{
  if (n <= 0) return;
  FMWR_REQUIRE(n < (int64_t)0xffffffffll, FMWR_ERR_UNSUPPORTED, "radix sort: more than 2^32 - 1 pairs");
  if (bits < 1) bits = 1;
  FMWR_REQUIRE(bits <= (int)(8 * sizeof(K)), FMWR_ERR_ARG, "radix sort: more key bits than the key type holds");
  const int passes = (bits + RS_BITS - 1) / RS_BITS;
  constexpr int TILE = RsTile<K>::SIZE;
  const int64_t tiles = ceil_div64(n, TILE);
  const bool wide = n >= (1ll << 30);              // the look-back words then need more than 30 value bits
  DBuf<uint32_t> hist;                              // [passes][256] + [passes] tickets
  hist.alloc((size_t)passes * RS_BINS + RS_MAX_PASSES);
  hist.zero(ctx->stream);
  uint32_t* tickets = hist.p + (size_t)passes * RS_BINS;
  DBuf<char> state;
  const size_t state_bytes = (size_t)tiles * RS_BINS * (wide ? 8 : 4);
  state.alloc(state_bytes);
  DBuf<K> kalt;
  DBuf<uint32_t> valt;
  if (passes > 1) { kalt.alloc(n); valt.alloc(n); }
  {
    const int grid = (int)std::min<int64_t>(ceil_div64(n, 256 * 8), (int64_t)ctx->sm_count * 8);
    FMWR_LAUNCH(ctx, radix_hist_kernel<K>, grid, 256, 0, key_in, n, passes, bits, hist.p);
    FMWR_LAUNCH(ctx, radix_scan_kernel, 1, 32 * RS_MAX_PASSES, 0, hist.p, passes);
  }
  const K* kin = key_in;
  const uint32_t* vin = val_in;
  for (int ps = 0; ps < passes; ++ps) {
    // the last pass lands in the caller's output; the ones before alternate so that it does
    const bool to_out = ((passes - 1 - ps) & 1) == 0;
    K* ko = to_out ? key_out : kalt.p;
    uint32_t* vo = to_out ? val_out : valt.p;
    FMWR_CUDA(cudaMemsetAsync(state.p, 0, state_bytes, ctx->stream));
    const int shift = ps * RS_BITS;
    const int nb = std::min(RS_BITS, bits - shift);
    if (wide) launch_pass<K, uint64_t>(ctx, kin, ko, vin, vo, n, shift, nb, hist.p + (size_t)ps * RS_BINS, (uint64_t*)state.p, tickets + ps, (int)tiles);
    else launch_pass<K, uint32_t>(ctx, kin, ko, vin, vo, n, shift, nb, hist.p + (size_t)ps * RS_BINS, (uint32_t*)state.p, tickets + ps, (int)tiles);
    kin = ko; vin = vo;
  }
  FMWR_CUDA(cudaStreamSynchronize(ctx->stream));     // the scratch buffers above are released on return
}

}  // namespace

void sort_pairs_u32(fmwr_ctx* ctx, const uint32_t* key_in, uint32_t* key_out, const uint32_t* val_in, uint32_t* val_out, int64_t n, int bits)
{
  radix_sort_pairs<uint32_t>(ctx, key_in, key_out, val_in, val_out, n, bits);
}

void sort_pairs_u64(fmwr_ctx* ctx, const uint64_t* key_in, uint64_t* key_out, const uint32_t* val_in, uint32_t* val_out, int64_t n, int bits)
{
  radix_sort_pairs<uint64_t>(ctx, key_in, key_out, val_in, val_out, n, bits);
}

}  // namespace fmwr
