// Bandwidth-bound stages of the segmentation path: tile extraction (K1), glue +
// threshold (K6, legacy API: the product path glues in the UNet head), connected-component labelling with
// OpenCV-identical numbering + island stats (K7 / K8, ccl_warp.cuh), group canvases (K9), 224x224 crops (K10),
// general-height resize (K11), the stroke front end's encode post-process (K12), plus the host-side planners.
//
// Data layout in HBM
//   lines_rgb : packed (128, W, 3) u8 images at sd_line.img_off
//   planes    : packed (128, pitch) planes at sd_line.px_off, pitch = round_up(W,128);
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
#include <cstdlib>

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

// one CTA per (tile, group of kExtRows rows); a thread converts 3 pixels of every row of the group, all source
// bytes of the group are requested before the first conversion (the kernel is a pure stream: 3 B in, 16 B out).
constexpr int kExtRows = 8;
__global__ void __launch_bounds__(128) tile_extract_f16_kernel(
    const uint8_t* __restrict__ rgb, const sd_line* __restrict__ L, int n_lines, int tile_w,
    uint4* __restrict__ out) {
  constexpr int kGroups = SD_TILE_H / kExtRows;
  const int tile = blockIdx.x / kGroups, row0 = (blockIdx.x % kGroups) * kExtRows;
  __shared__ int s_l;
  if (threadIdx.x == 0) s_l = find_tile_line(L, n_lines, tile);
  __syncthreads();
  const sd_line ln = L[s_l];
  const int ti = tile - ln.first_tile;
  const int wd = tile_width(ln, ti);
  const int x0 = (ln.n_tiles == 1) ? 0 : ti * ln.wu;
  for (int x = threadIdx.x; x < tile_w; x += blockDim.x) {
    uint8_t px[kExtRows][3];
    const bool in = x < wd;
#pragma unroll
    for (int r = 0; r < kExtRows; ++r) {
      const uint8_t* src = rgb + ln.img_off + ((int64_t)(row0 + r) * ln.width + x0 + x) * 3;
      px[r][0] = in ? __ldg(src + 0) : 0; px[r][1] = in ? __ldg(src + 1) : 0; px[r][2] = in ? __ldg(src + 2) : 0;
    }
#pragma unroll
    for (int r = 0; r < kExtRows; ++r) {
      // (x / 255.).astype(float32) then to fp16 (evaluate_binarize.py:99)
      // x/255 is never within an f32 ulp of an fp16 rounding boundary, so the
      // f32 division rounds to the same half as numpy's float64 path (tested
      // for all 256 values).
      act2_t rg = floats2act2((float)px[r][0] / 255.f, (float)px[r][1] / 255.f);
      act2_t b0 = floats2act2((float)px[r][2] / 255.f, 0.f);
      uint4 v = make_uint4(*reinterpret_cast<uint32_t*>(&rg), *reinterpret_cast<uint32_t*>(&b0), 0u, 0u);
      out[((int64_t)tile * SD_TILE_H + row0 + r) * tile_w + x] = v;
    }
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
// 16 consecutive source elements of one tile row, starting at element offset `off` of the row (may be negative
// or run past the row: those positions are masked by the caller, so word indices are simply clamped to the row).
__device__ __forceinline__ void load16_u8(const uint32_t* __restrict__ row_words, int off, int row_words_n, uint32_t out[4]) {
  const int w0 = off >> 2;                        // floor
  const int sh = (off & 3) * 8;
  uint32_t w[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) w[k] = __ldg(row_words + min(max(w0 + k, 0), row_words_n - 1));
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = __funnelshift_r(w[k], w[k + 1], sh);
}

// 16 consecutive fp16 probabilities -> 16 bytes of 0xFF/0x00 by `> thr`.
__device__ __forceinline__ void load16_f16_thr(const uint32_t* __restrict__ row_words, int off, int row_words_n, float thr,
                                               uint32_t out[4]) {
  const int w0 = off >> 1;
  const int sh = (off & 1) * 16;
  uint32_t w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = __ldg(row_words + min(max(w0 + k, 0), row_words_n - 1));
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

// one CTA per (line, group of kGlueRows rows): a thread owns one 16-px unit column, resolves the covering
// tiles once and then issues the loads of all its rows back to back (the kernel is latency bound otherwise).
constexpr int kGlueRows = 8;
template <bool kProb>
__global__ void __launch_bounds__(64) glue_kernel(
    const void* __restrict__ tiles, const sd_line* __restrict__ L, int64_t tile_elems_total, float thr,
    uint32_t on_rep, uint4* __restrict__ out) {
  constexpr int kGroups = SD_TILE_H / kGlueRows;
  const int l = blockIdx.x / kGroups, r0 = (blockIdx.x % kGroups) * kGlueRows;
  const sd_line ln = L[l];
  const int units = ln.pitch >> 4;
  for (int u = threadIdx.x; u < units; u += blockDim.x) {
    const int x0 = u * 16;
    uint32_t acc[kGlueRows][4];
#pragma unroll
    for (int r = 0; r < kGlueRows; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0u; }
    if (x0 < ln.width) {
      const int xl = min(x0 + 15, ln.width - 1);
      const int iA = (ln.n_tiles == 1) ? 0 : min(xl / ln.wu, ln.n_tiles - 1);
      // every tile that covers the run: two at the default 384 / 64 geometry (wu >= 192 > overlap), more when
      // overlap > wu (helper/split.py accepts any 0 <= overlap < tile_w)
      for (int t = 0; t <= iA; ++t) {
        const int i = iA - t;
        const int start = (ln.n_tiles == 1) ? 0 : i * ln.wu;
        const int wd = tile_width(ln, i);
        if (start + wd <= x0) break;                  // tile ends left of the run; so do all earlier tiles
        // valid run positions k: 0 <= x0 + k - start < wd  and x0 + k < W
        const int lo = max(0, start - x0), hi = min(min(16, start + wd - x0), ln.width - x0);
        if (lo >= hi) continue;
        const uint32_t m0 = range_mask(0, lo, hi), m1 = range_mask(1, lo, hi), m2 = range_mask(2, lo, hi), m3 = range_mask(3, lo, hi);
        // tile rows are tile_w elements = a whole number of 32-bit words, 4-byte aligned
        const int wpr = kProb ? ln.tile_w >> 1 : ln.tile_w >> 2;
        const uint32_t* row0 = reinterpret_cast<const uint32_t*>(tiles) + ((int64_t)(ln.first_tile + i) * SD_TILE_H + r0) * wpr;
        const int off = x0 - start;
        uint32_t v[kGlueRows][4];
#pragma unroll
        for (int r = 0; r < kGlueRows; ++r) {
          if (kProb) load16_f16_thr(row0 + r * wpr, off, wpr, thr, v[r]);
          else load16_u8(row0 + r * wpr, off, wpr, v[r]);
        }
        if (t == 0) {                                 // first covering tile: max(0, v) == v
#pragma unroll
          for (int r = 0; r < kGlueRows; ++r) { acc[r][0] = v[r][0] & m0; acc[r][1] = v[r][1] & m1; acc[r][2] = v[r][2] & m2; acc[r][3] = v[r][3] & m3; }
        } else {
#pragma unroll
          for (int r = 0; r < kGlueRows; ++r) {
            acc[r][0] = kProb ? (acc[r][0] | (v[r][0] & m0)) : __vmaxu4(acc[r][0], v[r][0] & m0);
            acc[r][1] = kProb ? (acc[r][1] | (v[r][1] & m1)) : __vmaxu4(acc[r][1], v[r][1] & m1);
            acc[r][2] = kProb ? (acc[r][2] | (v[r][2] & m2)) : __vmaxu4(acc[r][2], v[r][2] & m2);
            acc[r][3] = kProb ? (acc[r][3] | (v[r][3] & m3)) : __vmaxu4(acc[r][3], v[r][3] & m3);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kGlueRows; ++r) {
      if (kProb) { acc[r][0] &= on_rep; acc[r][1] &= on_rep; acc[r][2] &= on_rep; acc[r][3] &= on_rep; }
      out[((ln.px_off + (int64_t)(r0 + r) * ln.pitch) >> 4) + u] = make_uint4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
  }
}

// ---------------------------------------------------------------------------
// K7 / K8 fused: connected-component labelling with OpenCV's numbering and the cv2 stats rows live
// in ccl_warp.cuh (warp-per-strip run-level union-find; the strip-per-CTA kernels of round 1 are gone).
// A line (128 x pitch, pitch % 128 == 0) is cut into strips of 128 x 128 px = 64 x 64 blocks of 2 x 2 px; block
// indices are global: line.blk_off + br * bw + bc.
// ---------------------------------------------------------------------------
constexpr int kStripBlocks = 4096;
}  // namespace sd
#include "ccl_warp.cuh"
namespace sd {

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

// ---------------------------------------------------------------------------
// K10: 224x224 stroke-estimator crops (evaluate_strokes.py:202-222): per group
//   canvas {0,1} -> cv2.normalize MINMAX (0/255; a constant canvas becomes 0) -> cv2.resize INTER_LINEAR to
//   (rs_w, rs_h) -> zero pad to size x size (odd remainder goes right / bottom) = `image`;
//   `image_input` = per channel ((MINMAX(image) / 255 - mean) / std) as f32, through a 3 x 256 table the host
//   fills in float64 exactly like the reference expression.
// cv2's 8-bit bilinear is fixed point and is reproduced bit for bit (oracle/segmentation_ref.py
// resize_linear_u8, checked against cv2 itself): 11-bit coefficients rounded from float, horizontal pass in
// int32, vertical pass ((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2; x offsets clamp with the
// coefficient zeroed, y offsets clamp the row only; an exact 2x decimation takes cv2's area path.
// One CTA per group; the padded crop is built in shared memory so its min / max are known before it is written.
// ---------------------------------------------------------------------------
constexpr int kCropMax = 256;   // largest supported crop edge

__device__ __forceinline__ void crop_coeff(int d, int dn, int sn, bool clamp, int& ofs, int& a0, int& a1) {
  const double scale = 1.0 / ((double)dn / (double)sn);
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (clamp) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  }
  ofs = s;
  a0 = __float2int_rn((1.f - f) * 2048.f);
  a1 = __float2int_rn(f * 2048.f);
}

__global__ void __launch_bounds__(256) group_crop_kernel(
    const uint8_t* __restrict__ canvas, const int64_t* __restrict__ groups, const int32_t* __restrict__ rs_dims,
    int size, uint8_t* __restrict__ out_u8, float* __restrict__ out_f32, const float* __restrict__ lut) {
  extern __shared__ uint8_t s_img[];                 // size * size
  __shared__ int s_xo[kCropMax], s_yo[kCropMax];
  __shared__ short s_xa[kCropMax][2], s_yb[kCropMax][2];
  __shared__ int s_red[2][8];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int64_t* G = groups + (int64_t)g * 6;
  const int w = (int)(G[3] - G[1]), h = (int)(G[4] - G[2]);
  const uint8_t* cv = canvas + G[5];
  const int rs_w = rs_dims[2 * g], rs_h = rs_dims[2 * g + 1];
  const int pad_l = (size - rs_w) / 2, pad_t = (size - rs_h) / 2;
  // cv2.normalize(canvas, 0, 255, MINMAX): all-equal input -> zeros
  int has_zero = 0;
  for (int i = tid; i < w * h; i += blockDim.x) has_zero |= (cv[i] == 0);
  has_zero = __syncthreads_or(has_zero);
  const int on = has_zero ? 255 : 0;
  const bool area2 = (w == 2 * rs_w) && (h == 2 * rs_h);
  for (int d = tid; d < rs_w; d += blockDim.x) { int o, a0, a1; crop_coeff(d, rs_w, w, true, o, a0, a1); s_xo[d] = o; s_xa[d][0] = (short)a0; s_xa[d][1] = (short)a1; }
  for (int d = tid; d < rs_h; d += blockDim.x) { int o, a0, a1; crop_coeff(d, rs_h, h, false, o, a0, a1); s_yo[d] = o; s_yb[d][0] = (short)a0; s_yb[d][1] = (short)a1; }
  __syncthreads();
  int vmax = 0, vmin = 255;
  for (int p = tid; p < size * size; p += blockDim.x) {
    const int oy = p / size, ox = p - oy * size;
    const int dx = ox - pad_l, dy = oy - pad_t;
    int v = 0;
    if (dx >= 0 && dx < rs_w && dy >= 0 && dy < rs_h) {
      if (area2) {
        const uint8_t* r0 = cv + (int64_t)(2 * dy) * w + 2 * dx;
        const int sum = (r0[0] != 0) + (r0[1] != 0) + (r0[w] != 0) + (r0[w + 1] != 0);
        v = (sum * on + 2) >> 2;
      } else {
        const int x0 = s_xo[dx], x1 = min(x0 + 1, w - 1);
        const int y0 = min(max(s_yo[dy], 0), h - 1), y1 = min(max(s_yo[dy] + 1, 0), h - 1);
        const int a0 = s_xa[dx][0], a1 = s_xa[dx][1], b0 = s_yb[dy][0], b1 = s_yb[dy][1];
        const uint8_t* r0 = cv + (int64_t)y0 * w;
        const uint8_t* r1 = cv + (int64_t)y1 * w;
        const int H0 = ((r0[x0] != 0) * a0 + (r0[x1] != 0) * a1) * on;
        const int H1 = ((r1[x0] != 0) * a0 + (r1[x1] != 0) * a1) * on;
        v = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
      }
    }
    s_img[p] = (uint8_t)v;
    vmax = max(vmax, v); vmin = min(vmin, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
  }
  if ((tid & 31) == 0) { s_red[0][tid >> 5] = vmax; s_red[1][tid >> 5] = vmin; }
  __syncthreads();
  vmax = s_red[0][0]; vmin = s_red[1][0];
#pragma unroll
  for (int k = 1; k < 8; ++k) { vmax = max(vmax, s_red[0][k]); vmin = min(vmin, s_red[1][k]); }
  // image: 4 bytes per thread (size * size % 4 == 0)
  uint32_t* o8 = reinterpret_cast<uint32_t*>(out_u8 + (int64_t)g * size * size);
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s_img);
  for (int i = tid; i < size * size / 4; i += blockDim.x) o8[i] = s32[i];
  if (out_f32) {
    // second MINMAX (evaluate_strokes.py:58-69): cv2 converts with float scale / shift and one fused rounding
    const bool flat = vmax == vmin;
    const double scale_d = flat ? 0.0 : 255.0 / (double)(vmax - vmin);
    const float scale = (float)scale_d, shift = (float)(0.0 - (double)vmin * scale_d);
    float* of = out_f32 + (int64_t)g * 3 * size * size;
    for (int p = tid; p < size * size; p += blockDim.x) {
      int n = 0;
      if (!flat) n = min(max(__float2int_rn(fmaf((float)s_img[p], scale, shift)), 0), 255);
#pragma unroll
      for (int c = 0; c < 3; ++c) of[(int64_t)c * size * size + p] = __ldg(lut + c * 256 + n);
    }
  }
}

// ---------------------------------------------------------------------------
// K11: general-height line resize (common.py:85-93 / helper/split.py:127-135): cv2.resize(img, (int(w * (128 / h)),
//   128)) with the default INTER_LINEAR on (h, w, 3) u8, written straight into the packed line buffer that
//   tile_extract reads.  Same bit-exact restatement of cv2's 8-bit fixed-point bilinear as K10 (oracle:
//   resize_linear_u8 per channel, pinned against cv2 on 3-channel images in tests/test_oracle.py).
// One CTA = 128 output columns x 32 rows of one line; the 128 + 32 coefficient pairs are computed once in shared
// memory (they need a double division each), then every thread walks one column through 16 rows, four rows of
// loads in flight at a time (the first version, 64 dependent rows per thread, was latency-bound).  The two
// horizontally adjacent source pixels (6 bytes) come in as aligned 32-bit words realigned with funnel shifts
// (a third of the load instructions of byte gathers).  The kernel is bound by instruction issue (~100 integer
// instructions per output pixel), not by HBM or latency: staging the source rows in shared memory (+25 %) and
// packing a warp's 96 output bytes into aligned words through shared memory (+40 %) were both measured slower.
// ---------------------------------------------------------------------------
constexpr int kRsCols = 128;    // output columns per CTA
constexpr int kRsBand = 32;     // output rows per CTA (two threads per column, 16 rows each)
constexpr int kRsBatch = 4;     // rows whose loads are in flight together
static_assert(SD_TILE_H % kRsBand == 0 && kRsBand % (2 * kRsBatch) == 0, "resize band");

// bytes p[0..5] of a 4-byte aligned buffer region as (lo = p[0..3], hi = p[4..7]); reads up to 8 bytes past p + 3
__device__ __forceinline__ void load6(const uint8_t* p, uint32_t& lo, uint32_t& hi) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const int o = (int)(a & 3);
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = (o == 3) ? __ldg(w + 2) : 0u;
  lo = __funnelshift_r(w0, w1, 8 * o);
  hi = __funnelshift_r(w1, w2, 8 * o);
}

__device__ __forceinline__ int resize_px(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1, bool area2) {
  if (area2) return (p00 + p01 + p10 + p11 + 2) >> 2;             // cv2: exact 2x decimation -> INTER_AREA
  const int H0 = p00 * a0 + p01 * a1, H1 = p10 * a0 + p11 * a1;
  const int v = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2;
  return min(max(v, 0), 255);
}

__global__ void __launch_bounds__(256) resize_lines_kernel(const uint8_t* __restrict__ src,
                                                           const sd_resize_job* __restrict__ jobs,
                                                           uint8_t* __restrict__ dst) {
  const sd_resize_job J = jobs[blockIdx.y];
  const int xc = blockIdx.x * kRsCols;
  if (xc >= J.dst_w) return;
  __shared__ int s_xo[kRsCols], s_yo[kRsBand];
  __shared__ short s_xa[kRsCols][2], s_yb[kRsBand][2];
  const int tid = threadIdx.x, col = tid & (kRsCols - 1), x = xc + col;
  const int yb = blockIdx.z * kRsBand;                  // this CTA's band of output rows
  const bool area2 = J.src_w == 2 * J.dst_w && J.src_h == 2 * SD_TILE_H;
  {
    int o = 0, a0 = 0, a1 = 0;
    if (tid < kRsCols) {
      if (x < J.dst_w) {
        if (area2) o = 2 * x;
        else crop_coeff(x, J.dst_w, J.src_w, true, o, a0, a1);
      }
      s_xo[col] = o; s_xa[col][0] = (short)a0; s_xa[col][1] = (short)a1;
    } else if (col < kRsBand) {
      if (area2) o = 2 * (yb + col);
      else crop_coeff(yb + col, SD_TILE_H, J.src_h, false, o, a0, a1);
      s_yo[col] = o; s_yb[col][0] = (short)a0; s_yb[col][1] = (short)a1;
    }
  }
  __syncthreads();
  if (x >= J.dst_w) return;
  const uint8_t* S = src + J.src_off;
  uint8_t* D = dst + J.dst_off;
  const int64_t srow = (int64_t)J.src_w * 3;
  const int x0 = s_xo[col] * 3;
  const int a0 = s_xa[col][0], a1 = s_xa[col][1];
  // x1 = min(x0 + 1, src_w - 1): the clamp only happens where a1 == 0 (crop_coeff), so the pixel after x0 may stand
  // in for it as long as its bytes are readable: the layout contract pads every image (include/sd_b200.h)
  const bool words = (reinterpret_cast<uintptr_t>(S) & 3) == 0;
  if (!words) {                       // misaligned source: byte gathers
    const int dx = (s_xo[col] + 1 < J.src_w) ? 3 : 0;
    for (int r = tid >> 7; r < kRsBand; r += 2) {
      const int y0 = min(max(s_yo[r], 0), J.src_h - 1), y1 = min(max(s_yo[r] + 1, 0), J.src_h - 1);
      const uint8_t* r0 = S + (int64_t)y0 * srow + x0;
      const uint8_t* r1 = S + (int64_t)y1 * srow + x0;
      uint8_t* o = D + ((int64_t)(yb + r) * J.dst_w + x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = (uint8_t)resize_px(r0[c], r0[dx + c], r1[c], r1[dx + c], a0, a1, s_yb[r][0], s_yb[r][1], area2);
    }
    return;
  }
  // 16 rows per thread in batches of kRsBatch: all loads of a batch are issued before any arithmetic
#pragma unroll 1
  for (int rb = tid >> 7; rb < kRsBand; rb += 2 * kRsBatch) {
    uint32_t l0[kRsBatch], h0[kRsBatch], l1[kRsBatch], h1[kRsBatch];
#pragma unroll
    for (int k = 0; k < kRsBatch; ++k) {
      const int r = rb + 2 * k;
      const int y0 = min(max(s_yo[r], 0), J.src_h - 1), y1 = min(max(s_yo[r] + 1, 0), J.src_h - 1);
      load6(S + (int64_t)y0 * srow + x0, l0[k], h0[k]);
      load6(S + (int64_t)y1 * srow + x0, l1[k], h1[k]);
    }
#pragma unroll
    for (int k = 0; k < kRsBatch; ++k) {
      const int r = rb + 2 * k;
      const int b0 = s_yb[r][0], b1 = s_yb[r][1];
      uint8_t* o = D + ((int64_t)(yb + r) * J.dst_w + x) * 3;
      o[0] = (uint8_t)resize_px(l0[k] & 255, l0[k] >> 24, l1[k] & 255, l1[k] >> 24, a0, a1, b0, b1, area2);
      o[1] = (uint8_t)resize_px((l0[k] >> 8) & 255, h0[k] & 255, (l1[k] >> 8) & 255, h1[k] & 255, a0, a1, b0, b1, area2);
      o[2] = (uint8_t)resize_px((l0[k] >> 16) & 255, (h0[k] >> 8) & 255, (l1[k] >> 16) & 255, (h1[k] >> 8) & 255, a0, a1, b0, b1, area2);
    }
  }
}
// ---------------------------------------------------------------------------
// K12: stroke-estimator front end, evaluate_strokes.py:72-91 (_encode_postprocess): the encoder output (B, C, h, w)
//   f32 becomes (B, 2h * 2w, C) f32: every value repeated on a 2 x 2 grid (the reference's stand-in for the encoder's
//   AdaptiveAvgPool2d), channels last, positions flattened.  Pure data movement: one CTA per (image, 32-channel slab)
//   reads the slab along its contiguous h*w axis, transposes it through shared memory and writes 128-byte channel
//   runs for each of the 4 h w output positions.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) encode_postprocess_kernel(const float* __restrict__ enc, int C, int h, int w,
                                                                 float* __restrict__ out) {
  extern __shared__ float s_slab[];                 // [32][h*w + 1]
  const int hw = h * w, b = blockIdx.y, c0 = blockIdx.x * 32;
  const int nc = min(32, C - c0);
  const float* src = enc + ((int64_t)b * C + c0) * hw;
  for (int i = threadIdx.x; i < nc * hw; i += blockDim.x) s_slab[(i / hw) * (hw + 1) + i % hw] = __ldg(src + i);
  __syncthreads();
  const int W2 = 2 * w, P = 4 * hw;
  float* dst = out + (int64_t)b * P * C + c0;
  for (int i = threadIdx.x; i < P * 32; i += blockDim.x) {
    const int c = i & 31, pos = i >> 5;
    if (c >= nc) continue;
    const int oy = pos / W2, ox = pos - oy * W2;
    dst[(int64_t)pos * C + c] = s_slab[c * (hw + 1) + (oy >> 1) * w + (ox >> 1)];
  }
}
}  // namespace sd

// ===========================================================================
// C ABI
// ===========================================================================
using namespace sd;

extern "C" int sd_encode_postprocess(const float* d_enc, int B, int C, int h, int w, float* d_out, void* stream) {
  if (B == 0) return SD_OK;
  SD_REQUIRE(d_enc && d_out && B > 0 && C > 0 && h > 0 && w > 0, "sd_encode_postprocess: bad argument");
  SD_REQUIRE(B <= 65535, "sd_encode_postprocess: %d images in one call (max 65535)", B);
  const int smem = 32 * (h * w + 1) * (int)sizeof(float);
  SD_REQUIRE(smem <= 48 * 1024, "sd_encode_postprocess: feature map %dx%d too large", h, w);
  encode_postprocess_kernel<<<dim3((C + 31) / 32, B), 256, smem, (cudaStream_t)stream>>>(d_enc, C, h, w, d_out);
  SD_LAUNCH_CHECK("encode_postprocess_kernel");
  return SD_OK;
}

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
    // reconstruct_images pastes tile k at sum_{j<k} (width_j - overlap) (helper/split.py:119), which is k * wu only
    // while no tile before the last one is clipped at the image edge, and the gather-form glue / fused head assume
    // at most two tiles per column.  True for the default 384 / 64 geometry (wu >= 192); anything else is refused
    // loudly instead of pasted differently from the reference.
    SD_REQUIRE(ln.n_tiles == 1 || (overlap <= ln.wu && (int64_t)(ln.n_tiles - 1) * ln.wu + overlap <= W),
               "sd_plan_lines: line %d (width %d): overlap %d is too large for tile stride %d (clipped inner tiles are not supported)",
               i, W, overlap, ln.wu);
    ln.first_tile = tiles;
    ln.tile_w = tile_w; ln.overlap = overlap;
    ln.pitch = (W + 127) / 128 * 128;                 // whole 128-px CCL strips; rows start 128-B aligned
    ln.bw = ln.pitch / 2;
    ln.img_off = img; ln.px_off = px; ln.blk_off = blk;
    tiles += ln.n_tiles;
    img += ((int64_t)SD_TILE_H * W * 3 + 15) / 16 * 16;
    px += (int64_t)SD_TILE_H * ln.pitch;
    blk += (int64_t)(SD_TILE_H / 2) * ln.bw;          // a multiple of SD_CCL_CHUNK (64 x 64 blocks per strip)
  }
  plan->img_bytes = img; plan->px_total = px; plan->blk_total = blk;
  plan->n_tiles = tiles; plan->n_lines = n_lines;
  return SD_OK;
}

extern "C" int sd_tile_dst_table(const sd_line* h_lines, int n_lines, uint8_t* d_planes_base, sd_tile_dst* h_out) {
  SD_REQUIRE(n_lines >= 0 && (n_lines == 0 || (h_lines && h_out)), "sd_tile_dst_table: null argument");
  for (int l = 0; l < n_lines; ++l) {
    const sd_line& ln = h_lines[l];
    for (int i = 0; i < ln.n_tiles; ++i) {
      const int start = ln.n_tiles == 1 ? 0 : i * ln.wu;            // helper/split.py:119: s += width_k - overlap == k * wu
      int end = ln.n_tiles == 1 ? ln.width : std::min((i + 1) * ln.wu + ln.overlap, ln.width);
      sd_tile_dst& t = h_out[ln.first_tile + i];
      t.d_dst = d_planes_base + ln.px_off + start;
      t.pitch = ln.pitch;
      t.width = std::min(end - start, ln.tile_w);
    }
  }
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
  tile_extract_f16_kernel<<<n_tiles * (SD_TILE_H / kExtRows), 128, 0, (cudaStream_t)stream>>>(
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
  glue_kernel<false><<<n_lines * (SD_TILE_H / kGlueRows), 64, 0, (cudaStream_t)stream>>>(
      d_tiles, d_lines, elems, 0.f, 0u, reinterpret_cast<uint4*>(d_out));
  SD_LAUNCH_CHECK("glue_kernel<u8>");
  return SD_OK;
}

extern "C" int sd_glue_threshold_f16(const void* d_prob, int n_tiles, const sd_line* d_lines, int n_lines,
                                     int64_t px_total, float bin_thr, int on_value, uint8_t* d_out, void* stream) {
  SD_REQUIRE(d_prob && d_lines && d_out && n_lines > 0 && n_tiles > 0 && px_total > 0 && px_total % 16 == 0,
             "sd_glue_threshold_f16: bad argument");
  SD_REQUIRE(on_value > 0 && on_value <= 255, "sd_glue_threshold_f16: on_value %d", on_value);
  const int64_t elems = (int64_t)n_tiles * SD_TILE_H * SD_TILE_W;
  const uint32_t rep = (uint32_t)on_value * 0x01010101u;
  glue_kernel<true><<<n_lines * (SD_TILE_H / kGlueRows), 64, 0, (cudaStream_t)stream>>>(
      d_prob, d_lines, elems, bin_thr, rep, reinterpret_cast<uint4*>(d_out));
  SD_LAUNCH_CHECK("glue_kernel<f16>");
  return SD_OK;
}

namespace sd {
static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
// CCL workspace (ccl_warp.cuh)
static size_t ccl_warp_carve(void* base, int64_t blk_total, CclWarpWork* w) {
  const size_t strips = (size_t)(blk_total / kStripBlocks);
  size_t off = 0;
  char* p = reinterpret_cast<char*>(base);
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += up256(bytes); return r; };
  int* parent = reinterpret_cast<int*>(take((size_t)blk_total * 4));
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(take((size_t)blk_total / 32 * 4));
  int* prefix = reinterpret_cast<int*>(take((size_t)blk_total / 32 * 4));
  uint4* pix = reinterpret_cast<uint4*>(take(strips * 128 * 16));
  uint2* rs = reinterpret_cast<uint2*>(take(strips * 64 * 8));
  uint16_t* roots = reinterpret_cast<uint16_t*>(take(strips * kStripBlocks * 2));
  uint16_t* parent_fb = reinterpret_cast<uint16_t*>(take(strips * kStripBlocks * 2));
  int* bnd_root = reinterpret_cast<int*>(take(strips * 128 * 4));
  uint32_t* bnd_bits = reinterpret_cast<uint32_t*>(take(strips * 8 * 4));
  unsigned int* ticket = reinterpret_cast<unsigned int*>(take(256));
  if (w) { w->blk_total = blk_total; w->n_strips = (int)strips; w->parent = parent; w->bitmap = bitmap; w->prefix = prefix; w->pix = pix; w->rs = rs; w->roots = roots; w->parent_fb = parent_fb;
           w->bnd_root = bnd_root; w->bnd_bits = bnd_bits; w->ticket = ticket; }
  return off;
}

// launch with programmatic stream serialization (the kernel calls pdl_wait() before it touches its predecessor's output)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// one instantiation of the label kernel: launch geometry for this device, then the launch itself
template <int MINB, int MAXRUNS>
static cudaError_t ccl_launch_label(const uint8_t* d_mask, const sd_line* d_lines, int n_lines, int strips, const CclWarpWork& w,
                                    cudaStream_t s) {
  const int smem_l = kCw * (int)sizeof(CwLabelSmem<MAXRUNS>);
  static PerDeviceOnce attr_once;
  static int per_sm_l = 0, sms = 148;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(ccl_warp_label_kernel<MINB, MAXRUNS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_l);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_l, ccl_warp_label_kernel<MINB, MAXRUNS>, 32 * kCw, smem_l);
    if (per_sm_l < 1) per_sm_l = 1;
  }
  const int ctas = (strips + kCw - 1) / kCw;
  ccl_warp_label_kernel<MINB, MAXRUNS><<<std::min(ctas, sms * per_sm_l), 32 * kCw, smem_l, s>>>(d_mask, d_lines, n_lines, strips, w);
  return cudaGetLastError();
}

// label (+ stats) with the warp-per-strip kernels
static int ccl_warp_run(const uint8_t* d_mask, const sd_line* d_lines, int n_lines, int64_t blk_total, int32_t* d_labels,
                        int32_t* d_num, int64_t* d_stat_off, int32_t* d_stats, int64_t cap_rows, void* d_work, cudaStream_t s) {
  CclWarpWork w;
  ccl_warp_carve(reinterpret_cast<void*>(((uintptr_t)d_work + 255) / 256 * 256), blk_total, &w);
  const int strips = (int)(blk_total / kStripBlocks);
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  if (d_stats && cap_rows > 0) SD_CUDA_CHECK(cudaMemsetAsync(d_stats, 0x7f, (size_t)cap_rows * 20, s));
  // SD_CCL_STAGE=k (debug / profiling): stop after the k-th kernel of the chain, so that CUDA-event times of the call
  // for k = 1..4 give every kernel's steady-state share (tools/ccl_bench.py --stages)
  const char* stage_env = getenv("SD_CCL_STAGE");
  const int stage = stage_env ? atoi(stage_env) : 99;
  // SD_CCL_OCC = resident label CTAs per SM the kernel is compiled for (8 / 10 / 12, see ccl_warp_label_kernel);
  // SD_CCL_MERGE=1 folds the seam merges into the line kernel (A/B switches; the defaults are the measured optimum)
  const char* occ_env = getenv("SD_CCL_OCC");
  const int occ = occ_env ? atoi(occ_env) : SD_CCL_OCC_DEFAULT;
  const char* merge_env = getenv("SD_CCL_MERGE");
  const bool merge_in_line = merge_env ? atoi(merge_env) != 0 : false;   // measured slower (profiles/r02_ccl_occ_merge_ab.json)
  cudaError_t le;
  if (occ >= 12) le = ccl_launch_label<12, 1536>(d_mask, d_lines, n_lines, strips, w, s);
  else if (occ >= 10) le = ccl_launch_label<10, 2048>(d_mask, d_lines, n_lines, strips, w, s);
  else le = ccl_launch_label<8, 2048>(d_mask, d_lines, n_lines, strips, w, s);
  if (le != cudaSuccess) { set_error("ccl_warp_label_kernel: %s", cudaGetErrorString(le)); return SD_ECUDA; }
  count_launch();
  if (stage < 2) return SD_OK;
  if (!merge_in_line) {
    launch_pdl(ccl_seam_merge_kernel, dim3((unsigned)ceil_div((int64_t)strips * 64, 256)), dim3(256), 0, s, w, strips);
    SD_LAUNCH_CHECK("ccl_seam_merge_kernel");
  }
  if (stage < 3) return SD_OK;
  if (merge_in_line) launch_pdl(ccl_line_kernel<true>, dim3(n_lines), dim3(1024), 0, s, d_lines, n_lines, w, d_num, d_stat_off);
  else launch_pdl(ccl_line_kernel<false>, dim3(n_lines), dim3(1024), 0, s, d_lines, n_lines, w, d_num, d_stat_off);
  SD_LAUNCH_CHECK("ccl_line_kernel");
  if (stage < 4) return SD_OK;
  launch_pdl(ccl_strip_write2_kernel, dim3(strips), dim3(256), 0, s, d_lines, n_lines, w, d_labels, (const int64_t*)d_stat_off, d_stats, cap_rows);
  SD_LAUNCH_CHECK("ccl_strip_write2_kernel");
  if (d_stats && cap_rows > 0) {
    launch_pdl(ccl_stats_finish_kernel, dim3(sms * 2), dim3(256), 0, s, d_stats, (const int64_t*)d_stat_off, n_lines, cap_rows);
    SD_LAUNCH_CHECK("ccl_stats_finish_kernel");
  }
  return SD_OK;
}
}  // namespace sd

extern "C" size_t sd_ccl_workspace_bytes(int64_t blk_total, int n_lines) {
  (void)n_lines;
  return ccl_warp_carve(nullptr, blk_total, nullptr) + 256;
}

extern "C" int sd_ccl_label(const uint8_t* d_mask, const sd_line* d_lines, int n_lines, int64_t px_total,
                            int64_t blk_total, int32_t* d_labels, int32_t* d_num, void* d_work, void* stream) {
  SD_REQUIRE(d_mask && d_lines && d_labels && d_num && d_work && n_lines > 0, "sd_ccl_label: null argument");
  SD_REQUIRE(blk_total > 0 && blk_total % SD_CCL_CHUNK == 0 && px_total == blk_total * 4 && blk_total < INT_MAX,
             "sd_ccl_label: bad totals");
  SD_REQUIRE(((uintptr_t)d_mask & 15) == 0 && ((uintptr_t)d_labels & 15) == 0, "sd_ccl_label: planes must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  return ccl_warp_run(d_mask, d_lines, n_lines, blk_total, d_labels, d_num, nullptr, nullptr, 0, d_work, s);
}

extern "C" int sd_ccl_label_stats(const uint8_t* d_mask, const sd_line* d_lines, int n_lines, int64_t px_total,
                                  int64_t blk_total, int32_t* d_labels, int32_t* d_num, int64_t* d_stat_off,
                                  int32_t* d_stats, int64_t cap_rows, void* d_work, void* stream) {
  SD_REQUIRE(d_mask && d_lines && d_labels && d_num && d_stat_off && d_stats && d_work && n_lines > 0, "sd_ccl_label_stats: null argument");
  SD_REQUIRE(blk_total > 0 && blk_total % SD_CCL_CHUNK == 0 && px_total == blk_total * 4 && blk_total < INT_MAX,
             "sd_ccl_label_stats: bad totals");
  SD_REQUIRE(cap_rows > 0, "sd_ccl_label_stats: cap_rows %lld", (long long)cap_rows);
  SD_REQUIRE(((uintptr_t)d_mask & 15) == 0 && ((uintptr_t)d_labels & 15) == 0, "sd_ccl_label_stats: planes must be 16-byte aligned");
  return ccl_warp_run(d_mask, d_lines, n_lines, blk_total, d_labels, d_num, d_stat_off, d_stats, cap_rows, d_work, (cudaStream_t)stream);
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

extern "C" int sd_group_crops(const uint8_t* d_canvas, const int64_t* d_groups, const int32_t* d_rs_dims, int n_groups,
                              int size, uint8_t* d_image_u8, float* d_input_f32, const float* d_lut, void* stream) {
  if (n_groups == 0) return SD_OK;
  SD_REQUIRE(d_canvas && d_groups && d_rs_dims && d_image_u8 && n_groups > 0, "sd_group_crops: bad argument");
  SD_REQUIRE(size > 0 && size <= kCropMax && size % 2 == 0, "sd_group_crops: size %d (even, <= %d)", size, kCropMax);
  SD_REQUIRE(!d_input_f32 || d_lut, "sd_group_crops: the f32 output needs the 3x256 table");
  const int smem = size * size;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) SD_CUDA_CHECK(cudaFuncSetAttribute(group_crop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCropMax * kCropMax));
  group_crop_kernel<<<n_groups, 256, smem, (cudaStream_t)stream>>>(d_canvas, d_groups, d_rs_dims, size, d_image_u8,
                                                                   d_input_f32, d_lut);
  SD_LAUNCH_CHECK("group_crop_kernel");
  return SD_OK;
}

extern "C" int sd_resize_lines(const uint8_t* d_src, const sd_resize_job* d_jobs, int n_jobs, int max_dst_w,
                               uint8_t* d_rgb, void* stream) {
  if (n_jobs == 0) return SD_OK;
  SD_REQUIRE(d_src && d_jobs && d_rgb && n_jobs > 0 && max_dst_w > 0, "sd_resize_lines: bad argument");
  SD_REQUIRE(n_jobs <= 65535, "sd_resize_lines: %d jobs in one call (max 65535)", n_jobs);
  const dim3 grid((max_dst_w + kRsCols - 1) / kRsCols, n_jobs, SD_TILE_H / kRsBand);
  resize_lines_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, d_jobs, d_rgb);
  SD_LAUNCH_CHECK("resize_lines_kernel");
  return SD_OK;
}
