// Bandwidth-bound stages of the segmentation path: tile extraction (K1), glue +
// threshold (K6), connected-component labelling with OpenCV-identical numbering
// (K7), island stats (K8), group canvases (K9), plus the host-side planners.
//
// Data layout in HBM
//   lines_rgb : packed (128, W, 3) u8 images at sd_line.img_off
//   planes    : packed (128, pitch) planes at sd_line.px_off, pitch = round_up(W,16);
//               mask planes are u8, label planes int32 (same element offsets).
//               Every row therefore starts 16-byte aligned and a thread moves
//               16 pixels with one 128-bit access.
//   CCL block space: 2x2 pixel blocks, 64 rows x bw (= pitch/2) columns per line at
//               sd_line.blk_off, key = br*bw + bc (order-isomorphic to OpenCV's
//               (r/2)*ceil(W/2)+c/2 first-block raster key).
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <limits.h>

namespace sd {

// ---------------------------------------------------------------------------
// line lookup: which line owns packed offset `off`?  lines are sorted by offset.
// ---------------------------------------------------------------------------
template <bool kBlk>
__device__ __forceinline__ int64_t line_off(const sd_line* L, int i) { return kBlk ? L[i].blk_off : L[i].px_off; }

template <bool kBlk>
__device__ __forceinline__ int find_line(const sd_line* __restrict__ L, int n, int64_t off) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (line_off<kBlk>(L, mid) <= off) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// CTA-cooperative variant: thread 0 searches for the CTA's first unit, every
// thread then walks forward (a CTA spans at most a few lines).
template <bool kBlk>
__device__ __forceinline__ int find_line_cta(const sd_line* __restrict__ L, int n, int64_t cta_first, int64_t mine) {
  __shared__ int s_line;
  if (threadIdx.x == 0) s_line = find_line<kBlk>(L, n, cta_first);
  __syncthreads();
  int l = s_line;
  while (l + 1 < n && line_off<kBlk>(L, l + 1) <= mine) ++l;
  return l;
}

__device__ __forceinline__ int tile_width(const sd_line& ln, int i) {
  // helper/split.py:31-34: tile i spans [i*wu, min((i+1)*wu + overlap, W))
  if (ln.n_tiles == 1) return ln.width < ln.tile_w ? ln.width : ln.tile_w;
  int end = (i + 1) * ln.wu + ln.overlap;
  if (end > ln.width) end = ln.width;
  int w = end - i * ln.wu;
  return w < ln.tile_w ? w : ln.tile_w;   // pad_image's crop branch (split.py:52-53)
}

// ---------------------------------------------------------------------------
// K1: tile extraction
// ---------------------------------------------------------------------------
__device__ __forceinline__ int find_tile_line(const sd_line* __restrict__ L, int n, int tile) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (L[mid].first_tile <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// one CTA per (tile, row); thread x handles output columns x, x+blockDim, ...
__global__ void __launch_bounds__(128) tile_extract_f16_kernel(
    const uint8_t* __restrict__ rgb, const sd_line* __restrict__ L, int n_lines, int tile_w,
    uint4* __restrict__ out) {
  const int tile = blockIdx.x >> 7, row = blockIdx.x & 127;
  __shared__ int s_l;
  if (threadIdx.x == 0) s_l = find_tile_line(L, n_lines, tile);
  __syncthreads();
  const sd_line ln = L[s_l];
  const int ti = tile - ln.first_tile;
  const int wd = tile_width(ln, ti);
  const int x0 = (ln.n_tiles == 1) ? 0 : ti * ln.wu;
  const uint8_t* src = rgb + ln.img_off + ((int64_t)row * ln.width + x0) * 3;
  uint4* dst = out + ((int64_t)tile * SD_TILE_H + row) * tile_w;
  for (int x = threadIdx.x; x < tile_w; x += blockDim.x) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (x < wd) {
      // (x / 255.).astype(float32) then to fp16 (evaluate_binarize.py:99)
      // x/255 is never within an f32 ulp of an fp16 rounding boundary, so the
      // f32 division rounds to the same half as numpy's float64 path (tested
      // for all 256 values).
      float r = (float)src[3 * x + 0] / 255.f;
      float g = (float)src[3 * x + 1] / 255.f;
      float b = (float)src[3 * x + 2] / 255.f;
      __half2 rg = __floats2half2_rn(r, g);
      __half2 b0 = __floats2half2_rn(b, 0.f);
      v.x = *reinterpret_cast<uint32_t*>(&rg);
      v.y = *reinterpret_cast<uint32_t*>(&b0);
    }
    dst[x] = v;
  }
}

// one CTA per (tile, channel, row); thread handles 4 output columns.
__global__ void __launch_bounds__(96) tile_extract_u8_kernel(
    const uint8_t* __restrict__ rgb, const sd_line* __restrict__ L, int n_lines, int tile_w,
    uint8_t* __restrict__ out) {
  const int row = blockIdx.x & 127, c = (blockIdx.x >> 7) % 3, tile = (blockIdx.x >> 7) / 3;
  __shared__ int s_l;
  if (threadIdx.x == 0) s_l = find_tile_line(L, n_lines, tile);
  __syncthreads();
  const sd_line ln = L[s_l];
  const int ti = tile - ln.first_tile;
  const int wd = tile_width(ln, ti);
  const int x0 = (ln.n_tiles == 1) ? 0 : ti * ln.wu;
  const uint8_t* src = rgb + ln.img_off + ((int64_t)row * ln.width + x0) * 3 + c;
  uint8_t* dst = out + (((int64_t)tile * 3 + c) * SD_TILE_H + row) * tile_w;
  for (int x = threadIdx.x * 4; x < tile_w; x += blockDim.x * 4) {
    uchar4 v;
    v.x = (x + 0 < wd) ? src[3 * (x + 0)] : 0;
    v.y = (x + 1 < wd) ? src[3 * (x + 1)] : 0;
    v.z = (x + 2 < wd) ? src[3 * (x + 2)] : 0;
    v.w = (x + 3 < wd) ? src[3 * (x + 3)] : 0;
    *reinterpret_cast<uchar4*>(dst + x) = v;   // tile_w % 4 == 0
  }
}

// ---------------------------------------------------------------------------
// K6: glue (+ threshold).  Gather form of reconstruct_images:
// column x of a line is covered by tile i = min(x / wu, n-1) and, inside the
// overlap, by tile i-1; the output is the max (OR) of the covering tiles.
// A thread produces 16 aligned output bytes of one row with one 128-bit store.
// ---------------------------------------------------------------------------
// 16 consecutive source bytes starting at (possibly unaligned, possibly
// negative) element offset `e` of a tile buffer with `n_elems` elements.
__device__ __forceinline__ void load16_u8(const uint8_t* __restrict__ base, int64_t e, int64_t n_elems,
                                          uint32_t out[4]) {
  const int64_t w0 = e >> 2;                      // floor (e may be negative)
  const int sh = (int)(e & 3) * 8;
  const int64_t wmax = (n_elems >> 2) - 1;
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(base);
  uint32_t w[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    int64_t wi = w0 + k;
    wi = wi < 0 ? 0 : (wi > wmax ? wmax : wi);    // clamped words are masked by the caller
    w[k] = __ldg(wp + wi);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = __funnelshift_r(w[k], w[k + 1], sh);
}

// 16 consecutive fp16 probabilities -> 16 bytes of 0xFF/0x00 by `> thr`.
__device__ __forceinline__ void load16_f16_thr(const __half* __restrict__ base, int64_t e, int64_t n_elems,
                                               float thr, uint32_t out[4]) {
  const int64_t w0 = e >> 1;
  const int sh = (int)(e & 1) * 16;
  const int64_t wmax = (n_elems >> 1) - 1;
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(base);
  uint32_t w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    int64_t wi = w0 + k;
    wi = wi < 0 ? 0 : (wi > wmax ? wmax : wi);
    w[k] = __ldg(wp + wi);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t pr = __funnelshift_r(w[2 * k + j], w[2 * k + j + 1], sh);
      __half2 h2 = *reinterpret_cast<__half2*>(&pr);
      float2 f = __half22float2(h2);
      acc |= (f.x > thr ? 0xFFu : 0u) << (16 * j);
      acc |= (f.y > thr ? 0xFFu : 0u) << (16 * j + 8);
    }
    out[k] = acc;
  }
}

// byte mask 0xFF for k in [lo, hi) relative to a 16-byte run, word `wi`.
__device__ __forceinline__ uint32_t range_mask(int wi, int lo, int hi) {
  uint32_t m = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    int k = wi * 4 + b;
    if (k >= lo && k < hi) m |= 0xFFu << (8 * b);
  }
  return m;
}

template <bool kProb>
__global__ void __launch_bounds__(256) glue_kernel(
    const void* __restrict__ tiles, const sd_line* __restrict__ L, int n_lines, int64_t n_units,
    int64_t tile_elems_total, float thr, uint32_t on_rep, uint4* __restrict__ out) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // 16-px unit in packed space
  const int64_t cta_first = (int64_t)blockIdx.x * blockDim.x * 16;
  const int64_t mine = (unit < n_units ? unit : n_units - 1) * 16;
  const int l = find_line_cta<false>(L, n_lines, cta_first, mine);
  if (unit >= n_units) return;
  const sd_line ln = L[l];
  const int64_t rel = unit * 16 - ln.px_off;
  const int row = (int)(rel / ln.pitch);
  const int x0 = (int)(rel - (int64_t)row * ln.pitch);
  uint32_t acc[4] = {0u, 0u, 0u, 0u};
  if (x0 < ln.width) {
    const int xl = min(x0 + 15, ln.width - 1);
    const int iA = (ln.n_tiles == 1) ? 0 : min(xl / ln.wu, ln.n_tiles - 1);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = iA - t;
      if (i < 0) break;
      const int start = (ln.n_tiles == 1) ? 0 : i * ln.wu;
      const int wd = tile_width(ln, i);
      // valid run positions k: 0 <= x0 + k - start < wd  and x0 + k < W
      const int lo = max(0, start - x0), hi = min(min(16, start + wd - x0), ln.width - x0);
      if (lo >= hi) continue;
      const int64_t e = ((int64_t)(ln.first_tile + i) * SD_TILE_H + row) * ln.tile_w + (x0 - start);
      uint32_t v[4];
      if (kProb) load16_f16_thr(reinterpret_cast<const __half*>(tiles), e, tile_elems_total, thr, v);
      else load16_u8(reinterpret_cast<const uint8_t*>(tiles), e, tile_elems_total, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t m = v[k] & range_mask(k, lo, hi);
        acc[k] = kProb ? (acc[k] | m) : __vmaxu4(acc[k], m);
      }
    }
    if (kProb) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] &= on_rep;
    }
  }
  out[unit] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
}

// ---------------------------------------------------------------------------
// K7: connected-component labelling (block-based union-find, min-root).
// A thread owns 8 horizontally adjacent 2x2 blocks (16 px x 2 rows).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nz16(uint4 v) {   // bit k = byte k non-zero
  uint32_t w[4] = {v.x, v.y, v.z, v.w}, r = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t t = w[k];
    t = (t | ((t & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;   // msb of each non-zero byte
    t = (t >> 7) * 0x00204081u;                                  // gather to bits 21..24
    r |= ((t >> 21) & 0xFu) << (4 * k);
  }
  return r;
}

struct RowBits {  // bit (k+1) = pixel x0 + k, bit 0 = pixel x0-1, bit 17 = pixel x0+16
  uint32_t up, top, bot;
};

__device__ __forceinline__ uint32_t row_bits18(const uint8_t* __restrict__ rowp, int x0, int pitch) {
  uint32_t b = nz16(__ldg(reinterpret_cast<const uint4*>(rowp + x0))) << 1;
  if (x0 > 0) b |= (rowp[x0 - 1] != 0) ? 1u : 0u;
  if (x0 + 16 < pitch) b |= (rowp[x0 + 16] != 0) ? (1u << 17) : 0u;
  return b;
}

__device__ __forceinline__ int uf_find(const int* __restrict__ parent, int a) {
  // .cg loads: other SMs re-parent nodes concurrently; a stale value would still
  // be an ancestor (correct), but fresh ones shorten the walk.
  int p = __ldcg(parent + a);
  while (p != a) { a = p; p = __ldcg(parent + a); }
  return a;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  // min-root union (the root of a component is its smallest block key)
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }       // a > b: hang a under b
    int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;                                      // somebody re-parented a meanwhile; retry
  }
}

struct BlkCtx {
  int l, br, bc0, key0;     // line, block row, first block column, key of first block
  const sd_line* ln;
};

// maps packed block-space unit (8 blocks) -> line / block row / column.
__device__ __forceinline__ bool blk_unit(const sd_line* __restrict__ L, int n_lines, int64_t n_units,
                                         sd_line& ln, int& br, int& bc0) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cta_first = (int64_t)blockIdx.x * blockDim.x * 8;
  const int64_t mine = (unit < n_units ? unit : n_units - 1) * 8;
  const int l = find_line_cta<true>(L, n_lines, cta_first, mine);
  if (unit >= n_units) return false;
  ln = L[l];
  const int64_t rel = unit * 8 - ln.blk_off;
  br = (int)(rel / ln.bw);
  bc0 = (int)(rel - (int64_t)br * ln.bw);
  return br < SD_TILE_H / 2;       // beyond: chunk padding
}

__global__ void __launch_bounds__(256) ccl_init_kernel(
    const uint8_t* __restrict__ mask, const sd_line* __restrict__ L, int n_lines, int64_t n_units,
    int* __restrict__ parent) {
  sd_line ln; int br, bc0;
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = blk_unit(L, n_lines, n_units, ln, br, bc0);
  if (unit >= n_units) return;
  int4 p0 = make_int4(-1, -1, -1, -1), p1 = p0;
  if (ok) {
    const uint8_t* m = mask + ln.px_off + (int64_t)(2 * br) * ln.pitch + 2 * bc0;
    uint32_t t = nz16(__ldg(reinterpret_cast<const uint4*>(m)));
    uint32_t b = nz16(__ldg(reinterpret_cast<const uint4*>(m + ln.pitch)));
    uint32_t a = t | b;
    const int key0 = br * ln.bw + bc0;
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ((a >> (2 * k)) & 3u) ? key0 + k : -1;
    p0 = make_int4(v[0], v[1], v[2], v[3]);
    p1 = make_int4(v[4], v[5], v[6], v[7]);
  }
  int4* dst = reinterpret_cast<int4*>(parent + unit * 8);
  dst[0] = p0; dst[1] = p1;
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(
    const uint8_t* __restrict__ mask, const sd_line* __restrict__ L, int n_lines, int64_t n_units,
    int* __restrict__ parent_all) {
  sd_line ln; int br, bc0;
  if (!blk_unit(L, n_lines, n_units, ln, br, bc0)) return;
  const uint8_t* m = mask + ln.px_off + (int64_t)(2 * br) * ln.pitch;
  const int x0 = 2 * bc0;
  const uint32_t top = row_bits18(m, x0, ln.pitch);
  const uint32_t bot = row_bits18(m + ln.pitch, x0, ln.pitch);
  if (((top | bot) & 0x1FFFEu) == 0) return;
  const uint32_t up = (br > 0) ? row_bits18(m - ln.pitch, x0, ln.pitch) : 0u;
  int* parent = parent_all + ln.blk_off;
  const int key0 = br * ln.bw + bc0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t p00 = (top >> (1 + 2 * k)) & 1u, p01 = (top >> (2 + 2 * k)) & 1u;
    const uint32_t p10 = (bot >> (1 + 2 * k)) & 1u, p11 = (bot >> (2 + 2 * k)) & 1u;
    if (!(p00 | p01 | p10 | p11)) continue;
    const int key = key0 + k;
    const uint32_t l01 = (top >> (2 * k)) & 1u, l11 = (bot >> (2 * k)) & 1u;
    const uint32_t ul = (up >> (2 * k)) & 1u, u10 = (up >> (1 + 2 * k)) & 1u;
    const uint32_t u11 = (up >> (2 + 2 * k)) & 1u, ur = (up >> (3 + 2 * k)) & 1u;
    if ((p00 | p10) & (l01 | l11)) uf_union(parent, key, key - 1);
    const bool cu = (p00 | p01) & (u10 | u11);
    if (cu) uf_union(parent, key, key - ln.bw);
    // up-left / up-right diagonals; redundant when the block above already
    // links us and itself touches that diagonal pixel's block.
    if ((p00 & ul) && !(cu && u10)) uf_union(parent, key, key - ln.bw - 1);
    if ((p01 & ur) && !(cu && u11)) uf_union(parent, key, key - ln.bw + 1);
  }
}

// CTA == one 2048-block chunk: flatten parents, count roots of the chunk.
__global__ void __launch_bounds__(256) ccl_flatten_count_kernel(
    const sd_line* __restrict__ L, int n_lines, int* __restrict__ parent_all, int* __restrict__ chunk_count) {
  const int64_t base = (int64_t)blockIdx.x * SD_CCL_CHUNK + threadIdx.x * 8;
  __shared__ int s_l;
  if (threadIdx.x == 0) s_l = find_line<true>(L, n_lines, (int64_t)blockIdx.x * SD_CCL_CHUNK);
  __syncthreads();
  const int64_t blk_off = L[s_l].blk_off;
  int* parent = parent_all + blk_off;
  const int k0 = (int)(base - blk_off);
  int4* pp = reinterpret_cast<int4*>(parent + k0);
  int4 a = pp[0], b = pp[1];
  int v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  int roots = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (v[k] >= 0) {
      int r = v[k];
      if (r != k0 + k) { r = uf_find(parent, r); }
      v[k] = r;
      roots += (r == k0 + k);
    }
  }
  // NOTE: other threads may still be walking through these entries; writing a
  // shorter path (still an ancestor) keeps every walk correct.
  pp[0] = make_int4(v[0], v[1], v[2], v[3]);
  pp[1] = make_int4(v[4], v[5], v[6], v[7]);
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) roots += __shfl_xor_sync(0xffffffffu, roots, o);
  if ((threadIdx.x & 31) == 0 && roots) atomicAdd(&s_cnt, roots);
  __syncthreads();
  if (threadIdx.x == 0) chunk_count[blockIdx.x] = s_cnt;
}

// one CTA per line: exclusive scan of the line's chunk counts.
__global__ void __launch_bounds__(256) ccl_scan_kernel(
    const sd_line* __restrict__ L, int n_lines, int64_t blk_total, const int* __restrict__ chunk_count,
    int* __restrict__ chunk_base, int* __restrict__ num_out) {
  const int l = blockIdx.x;
  const int64_t c0 = L[l].blk_off / SD_CCL_CHUNK;
  const int64_t c1 = ((l + 1 < n_lines) ? L[l + 1].blk_off : blk_total) / SD_CCL_CHUNK;
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t c = c0; c < c1; c += 256) {
    const int64_t i = c + threadIdx.x;
    const int v = (i < c1) ? chunk_count[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += s_warp[w];
    const int carry = s_carry;
    if (i < c1) chunk_base[i] = carry + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == 255) s_carry = carry + woff + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) num_out[l] = s_carry + 1;     // cv2 counts the background label
}

// CTA == chunk: rank roots inside the chunk, write final label at the root's slot.
__global__ void __launch_bounds__(256) ccl_rank_kernel(
    const int* __restrict__ parent_all, const int* __restrict__ chunk_base, const sd_line* __restrict__ L,
    int n_lines, int* __restrict__ rlabel_all) {
  const int64_t base = (int64_t)blockIdx.x * SD_CCL_CHUNK + threadIdx.x * 8;
  __shared__ int s_l;
  __shared__ int s_warp[8];
  if (threadIdx.x == 0) s_l = find_line<true>(L, n_lines, (int64_t)blockIdx.x * SD_CCL_CHUNK);
  __syncthreads();
  const int64_t blk_off = L[s_l].blk_off;
  const int k0 = (int)(base - blk_off);
  const int4* pp = reinterpret_cast<const int4*>(parent_all + base);
  int4 a = pp[0], b = pp[1];
  int v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) cnt += (v[k] == k0 + k);
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += s_warp[w];
  int next = chunk_base[blockIdx.x] + woff + inc - cnt + 1;   // labels start at 1
  int out[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) out[k] = (v[k] == k0 + k) ? next++ : 0;
  int4* rp = reinterpret_cast<int4*>(rlabel_all + base);
  rp[0] = make_int4(out[0], out[1], out[2], out[3]);
  rp[1] = make_int4(out[4], out[5], out[6], out[7]);
}

__global__ void __launch_bounds__(256) ccl_write_kernel(
    const uint8_t* __restrict__ mask, const sd_line* __restrict__ L, int n_lines, int64_t n_units,
    const int* __restrict__ parent_all, const int* __restrict__ rlabel_all, int* __restrict__ labels) {
  sd_line ln; int br, bc0;
  if (!blk_unit(L, n_lines, n_units, ln, br, bc0)) return;
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t poff = ln.px_off + (int64_t)(2 * br) * ln.pitch + 2 * bc0;
  const uint32_t t = nz16(__ldg(reinterpret_cast<const uint4*>(mask + poff)));
  const uint32_t b = nz16(__ldg(reinterpret_cast<const uint4*>(mask + poff + ln.pitch)));
  const int4* pp = reinterpret_cast<const int4*>(parent_all + unit * 8);
  int lab[8];
  if (t | b) {
    int4 pa = pp[0], pb = pp[1];
    int v[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
    const int* rl = rlabel_all + ln.blk_off;
#pragma unroll
    for (int k = 0; k < 8; ++k) lab[k] = (v[k] >= 0) ? __ldg(rl + v[k]) : 0;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) lab[k] = 0;
  }
  int4* o0 = reinterpret_cast<int4*>(labels + poff);
  int4* o1 = reinterpret_cast<int4*>(labels + poff + ln.pitch);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int4 r0, r1;
    r0.x = ((t >> (4 * q + 0)) & 1u) ? lab[2 * q] : 0;
    r0.y = ((t >> (4 * q + 1)) & 1u) ? lab[2 * q] : 0;
    r0.z = ((t >> (4 * q + 2)) & 1u) ? lab[2 * q + 1] : 0;
    r0.w = ((t >> (4 * q + 3)) & 1u) ? lab[2 * q + 1] : 0;
    r1.x = ((b >> (4 * q + 0)) & 1u) ? lab[2 * q] : 0;
    r1.y = ((b >> (4 * q + 1)) & 1u) ? lab[2 * q] : 0;
    r1.z = ((b >> (4 * q + 2)) & 1u) ? lab[2 * q + 1] : 0;
    r1.w = ((b >> (4 * q + 3)) & 1u) ? lab[2 * q + 1] : 0;
    o0[q] = r0; o1[q] = r1;
  }
}

// ---------------------------------------------------------------------------
// K8: island stats.  A thread scans 16 pixels of one row, merges equal-label
// runs locally, then issues one set of atomics per run.
// rows of d_stats during accumulation: (minx, miny, maxx, maxy, area).
// ---------------------------------------------------------------------------
__global__ void stats_init_kernel(int32_t* __restrict__ st, int64_t n_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  st[i * 5 + 0] = INT_MAX; st[i * 5 + 1] = INT_MAX; st[i * 5 + 2] = -1; st[i * 5 + 3] = -1; st[i * 5 + 4] = 0;
}

__global__ void __launch_bounds__(256) stats_accum_kernel(
    const int32_t* __restrict__ labels, const sd_line* __restrict__ L, int n_lines, int64_t n_units,
    const int64_t* __restrict__ stat_off, int32_t* __restrict__ st) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cta_first = (int64_t)blockIdx.x * blockDim.x * 16;
  const int64_t mine = (unit < n_units ? unit : n_units - 1) * 16;
  const int l = find_line_cta<false>(L, n_lines, cta_first, mine);
  if (unit >= n_units) return;
  const sd_line ln = L[l];
  const int64_t rel = unit * 16 - ln.px_off;
  const int row = (int)(rel / ln.pitch);
  const int x0 = (int)(rel - (int64_t)row * ln.pitch);
  const int4* p = reinterpret_cast<const int4*>(labels + unit * 16);
  int v[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) { int4 t = __ldg(p + q); v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
  int32_t* base = st + stat_off[l] * 5;
  int cur = 0, xs = 0, xe = 0;
#pragma unroll
  for (int k = 0; k <= 16; ++k) {
    const int lab = (k < 16) ? v[k] : 0;
    if (lab != cur) {
      if (cur > 0) {
        int32_t* r = base + (int64_t)(cur - 1) * 5;
        atomicMin(r + 0, x0 + xs); atomicMin(r + 1, row);
        atomicMax(r + 2, x0 + xe); atomicMax(r + 3, row);
        atomicAdd(r + 4, xe - xs + 1);
      }
      cur = lab; xs = k;
    }
    xe = k;
  }
}

__global__ void stats_finish_kernel(int32_t* __restrict__ st, int64_t n_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int minx = st[i * 5 + 0], miny = st[i * 5 + 1], maxx = st[i * 5 + 2], maxy = st[i * 5 + 3];
  st[i * 5 + 2] = maxx - minx + 1;
  st[i * 5 + 3] = maxy - miny + 1;
}

// ---------------------------------------------------------------------------
// K9: group canvases.  One CTA per group.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_canvas_kernel(
    const int32_t* __restrict__ labels, const sd_line* __restrict__ L, const int64_t* __restrict__ groups,
    const int32_t* __restrict__ group_of, const int64_t* __restrict__ stat_off, uint8_t* __restrict__ canvas) {
  const int g = blockIdx.x;
  const int64_t* G = groups + (int64_t)g * 6;
  const int l = (int)G[0], left = (int)G[1], top = (int)G[2], right = (int)G[3], bottom = (int)G[4];
  const sd_line ln = L[l];
  const int w = right - left, h = bottom - top;
  const int32_t* lab = labels + ln.px_off;
  const int32_t* gof = group_of + stat_off[l];
  uint8_t* out = canvas + G[5];
  const int n = w * h;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    const int v = __ldg(lab + (int64_t)(top + y) * ln.pitch + left + x);
    out[i] = (v > 0 && __ldg(gof + v - 1) == g) ? 1 : 0;
  }
}

}  // namespace sd

// ===========================================================================
// C ABI
// ===========================================================================
using namespace sd;

extern "C" int sd_plan_lines(const int32_t* h_widths, int n_lines, int tile_w, int overlap,
                             sd_line* out, sd_plan* plan) {
  SD_REQUIRE(h_widths && out && plan && n_lines >= 0, "sd_plan_lines: null argument");
  SD_REQUIRE(tile_w > overlap && overlap >= 0 && tile_w % 4 == 0, "sd_plan_lines: bad tile_w/overlap");
  int64_t img = 0, px = 0, blk = 0;
  int tiles = 0;
  for (int i = 0; i < n_lines; ++i) {
    const int W = h_widths[i];
    SD_REQUIRE(W > 0, "sd_plan_lines: line %d has width %d", i, W);
    sd_line& ln = out[i];
    ln.width = W;
    if (W < tile_w) { ln.n_tiles = 1; ln.wu = W; }               // helper/split.py:19-21
    else { ln.n_tiles = W / (tile_w - overlap) + 1; ln.wu = W / ln.n_tiles; }   // :25-26
    ln.first_tile = tiles;
    ln.tile_w = tile_w; ln.overlap = overlap;
    ln.pitch = (W + 15) / 16 * 16;
    ln.bw = ln.pitch / 2;
    ln.img_off = img; ln.px_off = px; ln.blk_off = blk;
    tiles += ln.n_tiles;
    img += ((int64_t)SD_TILE_H * W * 3 + 15) / 16 * 16;
    px += (int64_t)SD_TILE_H * ln.pitch;
    blk += ((int64_t)(SD_TILE_H / 2) * ln.bw + SD_CCL_CHUNK - 1) / SD_CCL_CHUNK * SD_CCL_CHUNK;
  }
  plan->img_bytes = img; plan->px_total = px; plan->blk_total = blk;
  plan->n_tiles = tiles; plan->n_lines = n_lines;
  return SD_OK;
}

namespace sd {
// group_intervals + group_connections + add_to_group (helper/partition.py:248-358).
// iv: n pairs (a,b) sorted by a.  Appends member indices to `members`, group ends to `ends`.
static void group_intervals_core(const int64_t* iv, int n, int64_t width, std::vector<int32_t>& members,
                                 std::vector<int32_t>& ends) {
  // :255-281: every interval wider than `width` links to all intervals it contains
  // (scan from the left, stop at the first a_i > b_o).
  std::vector<std::vector<int>> adj;
  std::vector<char> contained(n, 0);
  bool any_long = false;
  for (int o = 0; o < n; ++o) {
    const int64_t ao = iv[2 * o], bo = iv[2 * o + 1];
    if (bo - ao <= width) continue;
    if (!any_long) { adj.resize(n); any_long = true; }
    for (int k = 0; k < n; ++k) {
      if (k == o) continue;
      const int64_t ai = iv[2 * k], bi = iv[2 * k + 1];
      if (ai > bo) break;
      if (ao <= ai && bo >= bi) {
        adj[o].push_back(k); adj[k].push_back(o);
        contained[o] = contained[k] = 1;
      }
    }
  }
  // group_connections / add_to_group (:321-358): pre-order DFS in adjacency order from each
  // not-yet-grouped node, in index order; the start node is appended when first reached back
  // from a neighbour.
  if (any_long) {
    std::vector<char> done(n, 0), seen(n, 0);
    std::vector<std::pair<int, size_t>> stack;
    for (int f = 0; f < n; ++f) {
      if (adj[f].empty() || done[f]) continue;
      const size_t g0 = members.size();
      stack.clear();
      stack.emplace_back(f, 0);
      while (!stack.empty()) {
        auto& top = stack.back();
        if (top.second >= adj[top.first].size()) { stack.pop_back(); continue; }
        const int nxt = adj[top.first][top.second++];
        if (!seen[nxt]) {
          seen[nxt] = 1;
          members.push_back(nxt);
          stack.emplace_back(nxt, 0);
        }
      }
      for (size_t i = g0; i < members.size(); ++i) { done[members[i]] = 1; seen[members[i]] = 0; }
      done[f] = 1;
      if (members.size() > g0) ends.push_back((int32_t)members.size());
    }
  }
  // greedy packing of the remaining intervals (:287-310)
  int64_t w = 0, left = 0;
  size_t cur0 = members.size();
  for (int i = 0; i < n; ++i) {
    if (contained[i]) continue;
    const int64_t a = iv[2 * i], b = iv[2 * i + 1];
    const int64_t new_w = std::max(b - left, w);
    if (new_w > width) {
      if (members.size() > cur0) ends.push_back((int32_t)members.size());
      cur0 = members.size();
      members.push_back(i);
      w = b - a; left = a;
    } else {
      members.push_back(i);
      w = new_w;
    }
  }
  if (members.size() > cur0) ends.push_back((int32_t)members.size());
}
}  // namespace sd

extern "C" int sd_group_intervals(const int64_t* iv, int n, int64_t width, int32_t* members, int32_t* group_start) {
  SD_REQUIRE(n >= 0 && (n == 0 || (iv && members)) && group_start, "sd_group_intervals: null argument");
  std::vector<int32_t> mem, ends;
  group_intervals_core(iv, n, width, mem, ends);
  for (size_t i = 0; i < mem.size(); ++i) members[i] = mem[i];
  group_start[0] = 0;
  for (size_t g = 0; g < ends.size(); ++g) group_start[g + 1] = ends[g];
  return (int)ends.size();
}

extern "C" int64_t sd_group_lines(const int32_t* stats, const int64_t* stat_off, const int32_t* widths, int n_lines,
                                  const int64_t* order, int margin, int img_h, int64_t target_w, int64_t* groups,
                                  int32_t* group_of, int64_t* line_group_start, int64_t* canvas_bytes) {
  SD_REQUIRE(stat_off && widths && line_group_start && canvas_bytes && n_lines >= 0, "sd_group_lines: null argument");
  int64_t ng = 0, off = 0;
  std::vector<int64_t> iv, xs, ys, xf, yf;
  std::vector<int32_t> mem, ends;
  line_group_start[0] = 0;
  for (int l = 0; l < n_lines; ++l) {
    const int64_t r0 = stat_off[l];
    const int n = (int)(stat_off[l + 1] - r0);
    if (n > 0) {
      SD_REQUIRE(stats && order && groups && group_of, "sd_group_lines: null argument");
      const int W = widths[l];
      xs.resize(n); ys.resize(n); xf.resize(n); yf.resize(n); iv.resize(2 * (size_t)n);
      for (int k = 0; k < n; ++k) {
        // helper/partition.py:19-24 (2 px before, 3 px after, clipped to the image)
        const int32_t* st = stats + (r0 + k) * 5;
        xs[k] = std::max<int64_t>(st[0] - margin, 0);
        ys[k] = std::max<int64_t>(st[1] - margin, 0);
        xf[k] = std::min<int64_t>((int64_t)st[0] + st[2] + margin + 1, W);
        yf[k] = std::min<int64_t>((int64_t)st[1] + st[3] + margin + 1, img_h);
      }
      for (int k = 0; k < n; ++k) {
        const int64_t o = order[r0 + k];
        SD_REQUIRE(o >= 0 && o < n, "sd_group_lines: order[%lld] out of range", (long long)(r0 + k));
        iv[2 * k] = xs[o]; iv[2 * k + 1] = xf[o];          // intervals (xs, xs + crop_w), partition.py:42-46
      }
      mem.clear(); ends.clear();
      group_intervals_core(iv.data(), n, target_w, mem, ends);
      size_t m0 = 0;
      for (size_t g = 0; g < ends.size(); ++g) {
        int64_t left = INT64_MAX, top = INT64_MAX, right = INT64_MIN, bottom = INT64_MIN;
        for (size_t i = m0; i < (size_t)ends[g]; ++i) {
          const int64_t k = order[r0 + mem[i]];             // label - 1 of this member
          left = std::min(left, xs[k]); top = std::min(top, ys[k]);
          right = std::max(right, xf[k]); bottom = std::max(bottom, yf[k]);
          group_of[r0 + k] = (int32_t)ng;
        }
        int64_t* G = groups + ng * 6;
        G[0] = l; G[1] = left; G[2] = top; G[3] = right; G[4] = bottom; G[5] = off;
        off += (right - left) * (bottom - top);
        ++ng;
        m0 = ends[g];
      }
    }
    line_group_start[l + 1] = ng;
  }
  *canvas_bytes = off;
  return ng;
}

extern "C" int sd_tile_extract_f16(const uint8_t* d_rgb, const sd_line* d_lines, int n_lines, int n_tiles,
                                   void* d_out, void* stream) {
  SD_REQUIRE(d_rgb && d_lines && d_out && n_lines > 0 && n_tiles > 0, "sd_tile_extract_f16: bad argument");
  tile_extract_f16_kernel<<<n_tiles * SD_TILE_H, 128, 0, (cudaStream_t)stream>>>(
      d_rgb, d_lines, n_lines, SD_TILE_W, reinterpret_cast<uint4*>(d_out));
  SD_LAUNCH_CHECK("tile_extract_f16_kernel");
  return SD_OK;
}

extern "C" int sd_tile_extract_u8(const uint8_t* d_rgb, const sd_line* d_lines, int n_lines, int n_tiles,
                                  uint8_t* d_out, void* stream) {
  SD_REQUIRE(d_rgb && d_lines && d_out && n_lines > 0 && n_tiles > 0, "sd_tile_extract_u8: bad argument");
  tile_extract_u8_kernel<<<n_tiles * 3 * SD_TILE_H, 96, 0, (cudaStream_t)stream>>>(
      d_rgb, d_lines, n_lines, SD_TILE_W, d_out);
  SD_LAUNCH_CHECK("tile_extract_u8_kernel");
  return SD_OK;
}

extern "C" int sd_glue_u8(const uint8_t* d_tiles, int n_tiles, const sd_line* d_lines, int n_lines,
                          int64_t px_total, uint8_t* d_out, void* stream) {
  SD_REQUIRE(d_tiles && d_lines && d_out && n_lines > 0 && n_tiles > 0 && px_total > 0 && px_total % 16 == 0,
             "sd_glue_u8: bad argument");
  const int64_t elems = (int64_t)n_tiles * SD_TILE_H * SD_TILE_W;   // clamps edge loads
  const int64_t units = px_total / 16;
  glue_kernel<false><<<ceil_div(units, 256), 256, 0, (cudaStream_t)stream>>>(
      d_tiles, d_lines, n_lines, units, elems, 0.f, 0u, reinterpret_cast<uint4*>(d_out));
  SD_LAUNCH_CHECK("glue_kernel<u8>");
  return SD_OK;
}

extern "C" int sd_glue_threshold_f16(const void* d_prob, int n_tiles, const sd_line* d_lines, int n_lines,
                                     int64_t px_total, float bin_thr, int on_value, uint8_t* d_out, void* stream) {
  SD_REQUIRE(d_prob && d_lines && d_out && n_lines > 0 && n_tiles > 0 && px_total > 0 && px_total % 16 == 0,
             "sd_glue_threshold_f16: bad argument");
  SD_REQUIRE(on_value > 0 && on_value <= 255, "sd_glue_threshold_f16: on_value %d", on_value);
  const int64_t elems = (int64_t)n_tiles * SD_TILE_H * SD_TILE_W;
  const int64_t units = px_total / 16;
  const uint32_t rep = (uint32_t)on_value * 0x01010101u;
  glue_kernel<true><<<ceil_div(units, 256), 256, 0, (cudaStream_t)stream>>>(
      d_prob, d_lines, n_lines, units, elems, bin_thr, rep, reinterpret_cast<uint4*>(d_out));
  SD_LAUNCH_CHECK("glue_kernel<f16>");
  return SD_OK;
}

extern "C" size_t sd_ccl_workspace_bytes(int64_t blk_total, int n_lines) {
  (void)n_lines;
  const int64_t chunks = blk_total / SD_CCL_CHUNK;
  return (size_t)(blk_total * 4 * 2 + chunks * 4 * 2 + 256);
}

extern "C" int sd_ccl_label(const uint8_t* d_mask, const sd_line* d_lines, int n_lines, int64_t px_total,
                            int64_t blk_total, int32_t* d_labels, int32_t* d_num, void* d_work, void* stream) {
  SD_REQUIRE(d_mask && d_lines && d_labels && d_num && d_work && n_lines > 0, "sd_ccl_label: null argument");
  SD_REQUIRE(blk_total > 0 && blk_total % SD_CCL_CHUNK == 0 && px_total > 0, "sd_ccl_label: bad totals");
  cudaStream_t s = (cudaStream_t)stream;
  int* parent = reinterpret_cast<int*>(d_work);
  int* rlabel = parent + blk_total;
  const int64_t chunks = blk_total / SD_CCL_CHUNK;
  int* chunk_count = rlabel + blk_total;
  int* chunk_base = chunk_count + chunks;
  const int64_t units = blk_total / 8;
  const int grid = ceil_div(units, 256);
  ccl_init_kernel<<<grid, 256, 0, s>>>(d_mask, d_lines, n_lines, units, parent);
  SD_LAUNCH_CHECK("ccl_init_kernel");
  ccl_merge_kernel<<<grid, 256, 0, s>>>(d_mask, d_lines, n_lines, units, parent);
  SD_LAUNCH_CHECK("ccl_merge_kernel");
  ccl_flatten_count_kernel<<<(int)chunks, 256, 0, s>>>(d_lines, n_lines, parent, chunk_count);
  SD_LAUNCH_CHECK("ccl_flatten_count_kernel");
  ccl_scan_kernel<<<n_lines, 256, 0, s>>>(d_lines, n_lines, blk_total, chunk_count, chunk_base, d_num);
  SD_LAUNCH_CHECK("ccl_scan_kernel");
  ccl_rank_kernel<<<(int)chunks, 256, 0, s>>>(parent, chunk_base, d_lines, n_lines, rlabel);
  SD_LAUNCH_CHECK("ccl_rank_kernel");
  ccl_write_kernel<<<grid, 256, 0, s>>>(d_mask, d_lines, n_lines, units, parent, rlabel, d_labels);
  SD_LAUNCH_CHECK("ccl_write_kernel");
  return SD_OK;
}

extern "C" int sd_island_stats(const int32_t* d_labels, const sd_line* d_lines, int n_lines, int64_t px_total,
                               const int64_t* d_stat_off, int64_t n_rows, int32_t* d_stats, void* stream) {
  SD_REQUIRE(d_labels && d_lines && d_stat_off && n_lines > 0 && px_total % 16 == 0, "sd_island_stats: bad argument");
  if (n_rows == 0) return SD_OK;
  SD_REQUIRE(d_stats, "sd_island_stats: null stats");
  cudaStream_t s = (cudaStream_t)stream;
  stats_init_kernel<<<ceil_div(n_rows, 256), 256, 0, s>>>(d_stats, n_rows);
  SD_LAUNCH_CHECK("stats_init_kernel");
  const int64_t units = px_total / 16;
  stats_accum_kernel<<<ceil_div(units, 256), 256, 0, s>>>(d_labels, d_lines, n_lines, units, d_stat_off, d_stats);
  SD_LAUNCH_CHECK("stats_accum_kernel");
  stats_finish_kernel<<<ceil_div(n_rows, 256), 256, 0, s>>>(d_stats, n_rows);
  SD_LAUNCH_CHECK("stats_finish_kernel");
  return SD_OK;
}

extern "C" int sd_group_canvas(const int32_t* d_labels, const sd_line* d_lines, const int64_t* d_groups,
                               int n_groups, const int32_t* d_group_of, const int64_t* d_stat_off,
                               uint8_t* d_canvas, void* stream) {
  if (n_groups == 0) return SD_OK;
  SD_REQUIRE(d_labels && d_lines && d_groups && d_group_of && d_stat_off && d_canvas && n_groups > 0,
             "sd_group_canvas: bad argument");
  group_canvas_kernel<<<n_groups, 256, 0, (cudaStream_t)stream>>>(d_labels, d_lines, d_groups, d_group_of,
                                                                 d_stat_off, d_canvas);
  SD_LAUNCH_CHECK("group_canvas_kernel");
  return SD_OK;
}
