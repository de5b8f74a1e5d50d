// Attention-UNet binarizer engine: weight packing, TMA descriptors, layer schedule,
// the CUDA-core kernels that are not GEMM-shaped (first 3->64 conv, 2x2 max-pool)
// and a slow all-CUDA-core debug implementation (impl=1) used only to cross-check
// the tcgen05 path during bring-up.
//
// Replaces onnxruntime's execution of the exported graph
// (/root/reference/derenderer/evaluate_binarize.py:48-53, :99-100).  Topology:
// SURVEY.md Appendix B.  All activations are NHWC fp16, accumulation fp32.
#include "conv_umma.cuh"
#include <vector>
#include <string>
#include <functional>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace sd {

// ---------------------------------------------------------------------------
// first conv: 3(+5 zero) -> 64 channels, 3x3, pad 1, bias + ReLU.  K = 27 is too
// thin for the tensor pipe (0.17 % of the network's FLOPs); one thread per pixel,
// weights broadcast from shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_first_kernel(
    const uint2* __restrict__ in /* NHWC8: 16 B/px, first 8 B = r,g,b,0 */, const float* __restrict__ w /* [27][64] */,
    const float* __restrict__ bias, act_t* __restrict__ out, int B, int H, int W) {
  __shared__ float4 sw[27 * 16];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(w)[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * H * W;
  if (pix >= total) return;
  const int x = (int)(pix % W); const int64_t r = pix / W;
  const int y = (int)(r % H);
  float acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = sb[j];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      float c[3] = {0.f, 0.f, 0.f};
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const uint2 v = __ldg(in + (pix + (int64_t)(ky - 1) * W + (kx - 1)) * 2);
        const act2_t rg = *reinterpret_cast<const act2_t*>(&v.x);
        const act2_t b0 = *reinterpret_cast<const act2_t*>(&v.y);
        c[0] = __low2float(rg); c[1] = __high2float(rg); c[2] = __low2float(b0);
      }
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float4* wr = sw + ((ky * 3 + kx) * 3 + ci) * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 ww = wr[j];
          acc[4 * j + 0] = fmaf(c[ci], ww.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(c[ci], ww.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(c[ci], ww.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(c[ci], ww.w, acc[4 * j + 3]);
        }
      }
    }
  }
  uint4* o = reinterpret_cast<uint4*>(out + pix * 64);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 t;
    uint32_t* tw = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      act2_t h = floats2act2(fmaxf(acc[8 * j + 2 * k], 0.f), fmaxf(acc[8 * j + 2 * k + 1], 0.f));
      tw[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    o[j] = t;
  }
}

// 2x2 max-pool, NHWC fp16; a thread handles 8 channels of one output pixel.
__global__ void __launch_bounds__(256) maxpool_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                      int B, int Ho, int Wo, int C8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo * C8;
  if (i >= total) return;
  const int c = (int)(i % C8); int64_t r = i / C8;
  const int x = (int)(r % Wo); r /= Wo;
  const int y = (int)(r % Ho); const int64_t n = r / Ho;
  const int Wi = 2 * Wo;
  const uint4* p = in + (((n * 2 * Ho + 2 * y) * Wi + 2 * x) * C8 + c);
  uint4 a = __ldg(p), b = __ldg(p + C8), d = __ldg(p + (int64_t)Wi * C8), e = __ldg(p + (int64_t)Wi * C8 + C8);
  act2_t* ha = reinterpret_cast<act2_t*>(&a); const act2_t* hb = reinterpret_cast<const act2_t*>(&b);
  const act2_t* hd = reinterpret_cast<const act2_t*>(&d); const act2_t* he = reinterpret_cast<const act2_t*>(&e);
#pragma unroll
  for (int k = 0; k < 4; ++k) ha[k] = __hmax2(__hmax2(ha[k], hb[k]), __hmax2(hd[k], he[k]));
  out[i] = a;
}

// ---------------------------------------------------------------------------
// debug implementation (impl=1): plain CUDA-core direct conv, two sources, optional
// nearest-x2 gather.  64 px x 64 cout per CTA, 4x4 per thread.
// ---------------------------------------------------------------------------
struct SimtConv {
  const act_t* in0; const act_t* in1; int c0, c1;
  const act_t* w;           // [taps][c0+c1][cout]
  const float* bias;
  act_t* out; float* out32;
  int B, H, W, cout, ks, up, relu;
};

__global__ void __launch_bounds__(256) conv_simt_kernel(SimtConv a) {
  __shared__ float As[8][64 + 1];
  __shared__ float Ws[8][64];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int cin = a.c0 + a.c1;
  const int IH = a.up ? a.H / 2 : a.H, IW = a.up ? a.W / 2 : a.W;
  float acc[4][4] = {};
  // loader roles
  const int lp = tid >> 2, lc = (tid & 3) * 2;           // A: pixel lp, channels lc, lc+1
  const int wc = tid >> 5, wn = (tid & 31) * 2;          // W: row wc, couts wn, wn+1
  const int64_t pm = m0 + lp;
  const int px = (int)(pm % a.W); const int64_t pr = pm / a.W;
  const int py = (int)(pr % a.H); const int64_t pn = pr / a.H;
  const int taps = a.ks * a.ks, half_k = a.ks / 2;
  for (int t = 0; t < taps; ++t) {
    int yy = py + t / a.ks - half_k, xx = px + t % a.ks - half_k;
    const bool inb = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
    if (a.up) { yy >>= 1; xx >>= 1; }
    const int64_t ip = (pn * IH + yy) * IW + xx;
    for (int cb = 0; cb < cin; cb += 8) {
      float2 av = make_float2(0.f, 0.f);
      if (inb) {
        const int c = cb + lc;
        const act_t* src = (c < a.c0) ? a.in0 + ip * a.c0 + c : a.in1 + ip * a.c1 + (c - a.c0);
        av = act22float2(*reinterpret_cast<const act2_t*>(src));
      }
      As[lc][lp] = av.x; As[lc + 1][lp] = av.y;
      float2 wv = make_float2(0.f, 0.f);
      if (n0 + wn < a.cout) wv = act22float2(*reinterpret_cast<const act2_t*>(a.w + ((int64_t)t * cin + cb + wc) * a.cout + n0 + wn));
      Ws[wc][wn] = wv.x; Ws[wc][wn + 1] = wv.y;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float av4[4], wv4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { av4[i] = As[k][ty * 4 + i]; wv4[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av4[i], wv4[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.cout) continue;
      float v = acc[i][j] + a.bias[n];
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.out32) a.out32[m * a.cout + n] = v;
      else a.out[m * a.cout + n] = f2act(v);
    }
  }
}

// debug / tap reader: a1 = x1 * psi with the arithmetic of the gate epilogue (fp32 product, one rounding)
__global__ void scale_by_psi_kernel(const uint4* __restrict__ x, const float* __restrict__ psi, uint4* __restrict__ out, int64_t px,
                                    int chunks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= px * chunks) return;
  const float sc = psi[i / chunks];
  uint4 v = x[i];
  act2_t* h = reinterpret_cast<act2_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 f = act22float2(h[k]);
    h[k] = floats2act2(f.x * sc, f.y * sc);
  }
  out[i] = v;
}

__global__ void gate_apply_kernel(const float* __restrict__ q, const float* __restrict__ psi_w, float psi_b,
                                  const act_t* __restrict__ x, act_t* __restrict__ out, int64_t M, int fint, int fl) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float dot = 0.f;
  for (int j = 0; j < fint; ++j) dot = fmaf(q[m * fint + j], psi_w[j], dot);
  const float s = 1.f / (1.f + expf(-(dot + psi_b)));
  for (int c = 0; c < fl; ++c) out[m * fl + c] = f2act(act2f(x[m * fl + c]) * s);
}

__global__ void head_kernel(const act_t* __restrict__ d2, const float* __restrict__ w, float b, float thr,
                            float* __restrict__ p32, __half* __restrict__ p16, uint8_t* __restrict__ mask, int64_t M, int c) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float dot = 0.f;
  for (int j = 0; j < c; ++j) dot = fmaf(act2f(d2[m * c + j]), w[j], dot);
  const float pr = 1.f / (1.f + expf(-(dot + b)));
  if (p32) p32[m] = pr;
  if (p16) p16[m] = __float2half_rn(pr);
  if (mask) mask[m] = pr > thr ? 255 : 0;
}

// ---------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Act {
  act_t* p = nullptr;
  int C = 0, H = 0, W = 0;
};

struct Level { int H, W, box_w, box_h, box_n; };

struct Op {
  std::string name;
  std::function<int(int /*B*/, cudaStream_t)> run;
  double flops_per_tile = 0;
};

}  // namespace sd

using namespace sd;

struct sd_engine {
  int device = 0, max_tiles = 0, cap_tiles = 0 /* max_tiles rounded up to the largest box_n */, H = 0, W = 0, impl = -1;
  bool finalized = false;
  std::vector<float> hw[SD_NUM_SLOTS], hb[SD_NUM_SLOTS];
  int cout[SD_NUM_SLOTS] = {}, cin[SD_NUM_SLOTS] = {}, ks[SD_NUM_SLOTS] = {};
  // device weights
  act_t* w_umma[SD_NUM_SLOTS] = {};   // [phases*cout][K]
  act_t* w_simt[SD_NUM_SLOTS] = {};   // [taps][cin][cout]
  float* w_f32[SD_NUM_SLOTS] = {};     // first conv [27][64]; psi / head vectors
  float* bias[SD_NUM_SLOTS] = {};
  float psi_b[4] = {}, head_b = 0.f;
  Level lv[5];
  Act act[SD_NUM_TAPS];                // named taps
  Act c1a, p1, c2a, p2, c3a, p3, c4a, p4, c5a, u5a, u4a, u3a, u2a;
  float* qbuf = nullptr;               // debug impl: gate pre-activation, fp32
  int* err_flag = nullptr;             // device alias of err_flag_host
  int* err_flag_host = nullptr;        // pinned, mapped: still readable after a kernel trapped (barrier wait codes)
  std::vector<void*> allocs;
  std::vector<Op> ops;
  // per-call outputs (captured by the head op)
  const void* in_tiles = nullptr;
  float thr = 0.5f; float* o32 = nullptr; __half* o16 = nullptr; uint8_t* omask = nullptr;
  const sd_tile_dst* odst = nullptr;   // head writes straight into the packed line planes (sd_unet_forward_lines)
  // timing
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  std::vector<float> last_ms;
  PFN_tmapEncodeTiled encode = nullptr;
  int num_sms = 148;
  int cta2 = 1;                        // Cout % 256 == 0 layers on SM pairs (cluster of 2, tcgen05 cta_group::2); SD_CTA2=0: one CTA per tile
  int cta2_n128 = 0;                   // SD_CTA2_N128=1: Cout = 128 layers on SM pairs too (N = 128, one M tile per CTA).  Measured neutral
                                       // (Up3 / Up_conv3.x -2..3 %, Conv2.x +1..3 %, pass unchanged), so the single-CTA MT = 2 kernel stays the default
  int psi_fused = 1;                   // level-1 gate (Att2) writes only psi; Up_conv2.0's band kernel scales the skip rows it stages (SD_PSI_FUSED=0: the gate writes x * psi)
  float* psi1 = nullptr;               // [max_tiles][H][W] fp32: sigma(psi) of the level-1 gate
  bool psi_live = false;               // this engine runs the psi-only gate: tap a1 is materialised on demand
  int gate_tma = 1;                    // gate epilogue moves the skip tensor with TMA (load, scale in smem, store); SD_GATETMA=0: per-thread row walk
  int up4 = 1;                         // Up2: four sub-pixel phases per work item (SD_UP4=0: generic kernel, phase by phase)
  int fuse_pool = 1;                   // MaxPool2x2 fused into the preceding conv's epilogue (SD_FUSEPOOL=0: separate kernel)
  int conv1_tc = 1;                    // Conv1.0 on the tensor pipe (SD_CONV1TC=0: CUDA-core kernel)
  int mt2_max_bn = 128;                // SD_MT2=128: 2 x (128 x BN) tiles per work item for BN <= 128 (0 = off)
  int bn_max = 256;                    // widest N tile of the generic conv kernel (SD_BNMAX=256 to try 128x256 tiles)
  int band = 1;                        // level-1 64-channel 3x3 layers on conv_band_kernel (SD_BAND=0: generic kernel)
};

namespace sd {

// the C ABI never leaves the caller's current device changed (a process may drive several GPUs)
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int dev_alloc(sd_engine* e, void** p, size_t bytes) {
  SD_CUDA_CHECK(cudaMalloc(p, bytes));
  e->allocs.push_back(*p);
  return SD_OK;
}

static int alloc_act(sd_engine* e, Act& a, int lvl, int C) {
  a.C = C; a.H = e->lv[lvl].H; a.W = e->lv[lvl].W;
  return dev_alloc(e, (void**)&a.p, (size_t)e->cap_tiles * a.H * a.W * C * sizeof(act_t));
}

static int make_tmap_act(sd_engine* e, CUtensorMap* tm, const Act& a, const Level& box) {  // box.box_* only
  cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)e->cap_tiles};
  cuuint64_t strides[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.W * a.C * 2, (cuuint64_t)a.H * a.W * a.C * 2};
  cuuint32_t boxd[4] = {64, (cuuint32_t)box.box_w, (cuuint32_t)box.box_h, (cuuint32_t)box.box_n};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = e->encode(tm, SD_TMAP_DTYPE, 4, a.p, dims, strides, boxd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(act C=%d W=%d H=%d) failed: %d", a.C, a.W, a.H, (int)r); return SD_ECUDA; }
  return SD_OK;
}

static int make_tmap_w(sd_engine* e, CUtensorMap* tm, const act_t* w, int rows, int K, int bn) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t boxd[2] = {64, (cuuint32_t)bn};
  cuuint32_t es[2] = {1, 1};
  CUresult r = e->encode(tm, SD_TMAP_DTYPE, 2, (void*)w, dims, strides, boxd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights rows=%d K=%d) failed: %d", rows, K, (int)r); return SD_ECUDA; }
  return SD_OK;
}

// output map of a store-epilogue conv.  For the sub-pixel up-conv, phase (py, px) writes pixel
// (2y+py, 2x+px) of a 2H x 2W tensor: a strided view indexed by the LOW-res (x, y).
static int make_tmap_out(sd_engine* e, CUtensorMap* tm, const Act& o, const Level& box, bool up, int ph) {
  const uint64_t C = o.C;
  cuuint64_t dims[4], strides[3];
  char* base = reinterpret_cast<char*>(o.p);
  if (!up) {
    dims[0] = C; dims[1] = o.W; dims[2] = o.H; dims[3] = e->cap_tiles;
    strides[0] = C * 2; strides[1] = (uint64_t)o.W * C * 2; strides[2] = (uint64_t)o.H * o.W * C * 2;
  } else {
    dims[0] = C; dims[1] = o.W / 2; dims[2] = o.H / 2; dims[3] = e->cap_tiles;
    strides[0] = 2 * C * 2; strides[1] = 2 * (uint64_t)o.W * C * 2; strides[2] = (uint64_t)o.H * o.W * C * 2;
    base += ((uint64_t)(ph >> 1) * o.W + (ph & 1)) * C * 2;
  }
  cuuint32_t boxd[4] = {64, (cuuint32_t)box.box_w, (cuuint32_t)box.box_h, (cuuint32_t)box.box_n};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = e->encode(tm, SD_TMAP_DTYPE, 4, base, dims, strides, boxd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out C=%d W=%d H=%d up=%d) failed: %d", o.C, o.W, o.H, (int)up, (int)r); return SD_ECUDA; }
  return SD_OK;
}

// pooled copy of a store-epilogue output: (C, W/2, H/2, N) with half-size boxes
static int make_tmap_pool(sd_engine* e, CUtensorMap* tm, const Act& o, int bw, int bh, int bn) {
  cuuint64_t dims[4] = {(cuuint64_t)o.C, (cuuint64_t)o.W, (cuuint64_t)o.H, (cuuint64_t)e->cap_tiles};
  cuuint64_t strides[3] = {(cuuint64_t)o.C * 2, (cuuint64_t)o.W * o.C * 2, (cuuint64_t)o.H * o.W * o.C * 2};
  cuuint32_t boxd[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = e->encode(tm, SD_TMAP_DTYPE, 4, o.p, dims, strides, boxd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(pool C=%d W=%d H=%d) failed: %d", o.C, o.W, o.H, (int)r); return SD_ECUDA; }
  return SD_OK;
}

template <int BN, int EPI, int MT = 1>
static int launch_conv(const ConvParams& p, int grid, cudaStream_t s) {
  using Cfg = ConvCfg<BN, EPI, MT>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) SD_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<BN, EPI, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  conv_umma_kernel<BN, EPI, MT><<<grid, Cfg::kThreads, Cfg::kSmemBytes, s>>>(p);
  SD_LAUNCH_CHECK("conv_umma_kernel");
  return SD_OK;
}

template <int CB, int EPI, bool PSI = false>
static int launch_band(const ConvParams& p, int grid, cudaStream_t s) {
  using Cfg = BandCfg<CB, EPI, PSI>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) SD_CUDA_CHECK(cudaFuncSetAttribute(conv_band_kernel<CB, EPI, PSI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  conv_band_kernel<CB, EPI, PSI><<<grid, Cfg::kThreads, Cfg::kSmemBytes, s>>>(p);
  SD_LAUNCH_CHECK("conv_band_kernel");
  return SD_OK;
}

static int dispatch_band(const ConvParams& p, int cb, int epi, int grid, cudaStream_t s) {
  if (cb == 1 && epi == EPI_STORE) return launch_band<1, EPI_STORE>(p, grid, s);
  if (cb == 2 && epi == EPI_STORE) return launch_band<2, EPI_STORE>(p, grid, s);
  if (cb == 1 && epi == EPI_HEAD) return launch_band<1, EPI_HEAD>(p, grid, s);
  set_error("dispatch_band: no kernel for CB=%d epilogue=%d", cb, epi);
  return SD_EINVAL;
}

// 2-SM variant (cluster of 2, cta_group::2) of the store-epilogue conv: N = 256, or N = 128 with one M tile per CTA
template <int BN>
static int launch_conv2(const ConvParams& p, int n_sms, cudaStream_t s) {
  using Cfg = Conv2CfgT<BN>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) SD_CUDA_CHECK(cudaFuncSetAttribute(conv_umma2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int n_work = ((p.m_tiles + 1) / 2) * p.n_tiles * p.n_phases;
  int pairs = n_sms / 2;
  if (n_work < pairs) pairs = n_work;
  conv_umma2_kernel<BN><<<2 * pairs, kConvThreads, Cfg::kSmemBytes, s>>>(p);
  SD_LAUNCH_CHECK("conv_umma2_kernel");
  return SD_OK;
}

static int dispatch_conv(const ConvParams& p, int bn, int mt, int epi, int grid, cudaStream_t s) {
  if (epi == EPI_STORE) {
    if (bn == 64 && mt == 2) return launch_conv<64, EPI_STORE, 2>(p, grid, s);
    if (bn == 128 && mt == 2) return launch_conv<128, EPI_STORE, 2>(p, grid, s);
    if (bn == 64) return launch_conv<64, EPI_STORE>(p, grid, s);
    if (bn == 128) return launch_conv<128, EPI_STORE>(p, grid, s);
    if (bn == 256) return launch_conv<256, EPI_STORE>(p, grid, s);
  } else if (epi == EPI_GATE) {
    if (bn == 32) return launch_conv<32, EPI_GATE>(p, grid, s);
    if (bn == 64) return launch_conv<64, EPI_GATE>(p, grid, s);
    if (bn == 128) return launch_conv<128, EPI_GATE>(p, grid, s);
    if (bn == 256) return launch_conv<256, EPI_GATE>(p, grid, s);
  } else if (epi == EPI_HEAD) {
    if (bn == 64) return launch_conv<64, EPI_HEAD>(p, grid, s);
  }
  set_error("dispatch_conv: no kernel for BN=%d epilogue=%d", bn, epi);
  return SD_EINVAL;
}

// which 3x3 taps collapse onto low-res offset index t (0/1) for output parity p (0/1)
static void phase_taps(int par, int t, int& k_lo, int& k_hi, int& d) {
  if (par == 0) { if (t == 0) { k_lo = 0; k_hi = 0; d = -1; } else { k_lo = 1; k_hi = 2; d = 0; } }
  else          { if (t == 0) { k_lo = 0; k_hi = 1; d = 0; }  else { k_lo = 2; k_hi = 2; d = 1; } }
}

// fp32 OIHW -> fp16 [phase*cout + co][tap*cin + ci]
static void pack_umma(const float* w, int cout, int cin, int ks, bool up, std::vector<act_t>& out) {
  if (!up) {
    const int taps = ks * ks, K = taps * cin;
    out.assign((size_t)cout * K, f2act(0.f));
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < taps; ++t)
          out[(size_t)co * K + t * cin + ci] = f2act(w[((size_t)co * cin + ci) * taps + t]);
  } else {
    const int K = 4 * cin;
    out.assign((size_t)4 * cout * K, f2act(0.f));
    for (int ph = 0; ph < 4; ++ph)
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) {
          int y0, y1, x0, x1, d;
          phase_taps(ph >> 1, ty, y0, y1, d);
          phase_taps(ph & 1, tx, x0, x1, d);
          for (int co = 0; co < cout; ++co)
            for (int ci = 0; ci < cin; ++ci) {
              float s = 0.f;
              for (int ky = y0; ky <= y1; ++ky)
                for (int kx = x0; kx <= x1; ++kx) s += w[((size_t)co * cin + ci) * 9 + ky * 3 + kx];
              out[((size_t)ph * cout + co) * K + (ty * 2 + tx) * cin + ci] = f2act(s);
            }
        }
  }
}

static int upload(sd_engine* e, void** dptr, const void* h, size_t bytes) {
  int r = dev_alloc(e, dptr, bytes);
  if (r) return r;
  SD_CUDA_CHECK(cudaMemcpy(*dptr, h, bytes, cudaMemcpyHostToDevice));
  return SD_OK;
}

// ---- op builders ------------------------------------------------------------
struct ConvSpec {
  const char* name;
  int slot;
  const Act* in0; const Act* in1;   // in1 = second concat source or null
  const Act* out;
  bool up;
  int epi;                          // EPI_*
  int att = -1;                     // gate: index 0..3 (psi slot / bias), x source = in1
  const Act* pool_out = nullptr;    // store epilogue: also write MaxPool2x2(out) here (fused pool)
  float* psi_out = nullptr;         // gate: write only sigma(psi) here (the consumer scales the skip tensor)
  const float* psi_in = nullptr;    // band conv with two sources: scale the staged rows of in0 by this plane
};

static int add_umma_conv(sd_engine* e, const ConvSpec& cs) {
  const int slot = cs.slot;
  const int cin_total = cs.in0->C + (cs.in1 ? cs.in1->C : 0);
  const int co = e->cout[slot];
  ConvParams p;
  memset(&p, 0, sizeof(p));
  // level of the A source
  int lvl = -1;
  for (int i = 0; i < 5; ++i) if (e->lv[i].H == cs.in0->H && e->lv[i].W == cs.in0->W) lvl = i;
  SD_REQUIRE(lvl >= 0, "add_umma_conv(%s): unknown level", cs.name);
  const Level& L = e->lv[lvl];
  int r;
  // level-1 3x3 layers with 64 output channels: dy-stacked band kernel
  const bool band_ok = e->band && lvl == 0 && L.W % 128 == 0 && co == 64 && !cs.up && e->ks[slot] == 3 &&
                       cs.in0->C == 64 && (!cs.in1 || cs.in1->C == 64) && (cs.epi == EPI_STORE || cs.epi == EPI_HEAD) &&
                       !(cs.in1 && cs.epi == EPI_HEAD);
  if (band_ok) {
    const int cb = cs.in1 ? 2 : 1;
    Level halo = L; halo.box_w = 130; halo.box_h = 1; halo.box_n = 1;
    if ((r = make_tmap_act(e, &p.tmA0, *cs.in0, halo))) return r;
    if (cs.in1 && (r = make_tmap_act(e, &p.tmA1, *cs.in1, halo))) return r;
    if ((r = make_tmap_w(e, &p.tmB, e->w_umma[slot], co, 9 * cin_total, 64))) return r;
    p.H = L.H; p.W = L.W; p.cout = co; p.relu = 1; p.n_phases = 1;
    p.bias = e->bias[slot]; p.err_flag = e->err_flag;
    for (int i = 0; i < 64; ++i) { p.bias_c[i] = e->hb[slot][i]; p.vec_c[i] = cs.epi == EPI_HEAD ? e->hw[SD_HEAD][i] : 0.f; }
    if (cs.epi == EPI_STORE) {
      p.out = cs.out->p; p.out_c = cs.out->C;
      if ((r = make_tmap_out(e, &p.tmOut[0], *cs.out, L, false, 0))) return r;
      if (cs.pool_out) {
        SD_REQUIRE(!cs.in1, "add_umma_conv(%s): fused pool needs a single source", cs.name);
        if ((r = make_tmap_pool(e, &p.tmPool, *cs.pool_out, 64, 1, 1))) return r;
        p.pool = 1;
      }
    } else {
      p.head_w = e->w_f32[SD_HEAD];
    }
    const bool psi = cs.psi_in != nullptr;
    if (psi) {
      SD_REQUIRE(cb == 2 && cs.epi == EPI_STORE, "add_umma_conv(%s): psi scaling needs the two-source store form", cs.name);
      p.psi_in = cs.psi_in;
    }
    Op op;
    op.name = std::string(cs.name) + "[band" + (psi ? ",psi" : "") + "]" + (p.pool ? "+pool" : "");
    const int epi = cs.epi, nsm = e->num_sms, segs = L.W / 128, H = L.H;
    op.flops_per_tile = 2.0 * L.H * L.W * co * 9 * cin_total;
    op.run = [e, p, cb, epi, nsm, segs, H, psi](int B, cudaStream_t s) mutable -> int {
      p.B = B;
      if (epi == EPI_HEAD) {
        p.head_b = e->head_b; p.thr = e->thr;
        p.prob_f32 = e->o32; p.prob_f16 = e->o16; p.mask_u8 = e->omask; p.tile_dst = e->odst;
      }
      const int n_work = B * ((H + kBandRows - 1) / kBandRows) * segs;
      if (psi) return launch_band<2, EPI_STORE, true>(p, n_work < nsm ? n_work : nsm, s);
      return dispatch_band(p, cb, epi, n_work < nsm ? n_work : nsm, s);
    };
    e->ops.push_back(op);
    return SD_OK;
  }
  if ((r = make_tmap_act(e, &p.tmA0, *cs.in0, L))) return r;
  if (cs.in1 && (r = make_tmap_act(e, &p.tmA1, *cs.in1, L))) return r;
  const int taps = cs.up ? 4 : e->ks[slot] * e->ks[slot];
  const int K = taps * cin_total;
  int bn = co >= 128 ? 128 : co;
  if (e->bn_max >= 256 && co % 256 == 0 && cs.epi == EPI_STORE) bn = 256;
  const int mt = (cs.epi == EPI_STORE && bn <= e->mt2_max_bn) ? 2 : 1;   // two M tiles share each B k-block
  if (cs.epi == EPI_GATE) bn = co;
  if ((r = make_tmap_w(e, &p.tmB, e->w_umma[slot], (cs.up ? 4 : 1) * co, K, bn))) return r;
  p.H = L.H; p.W = L.W;
  p.box_w = L.box_w; p.box_h = L.box_h; p.box_n = L.box_n;
  p.tiles_x = L.W / L.box_w; p.tiles_y = L.H / L.box_h;
  p.n_tiles = co / bn;
  p.n_phases = cs.up ? 4 : 1;
  p.n_taps = taps;
  p.c0_blocks = cs.in0->C / 64; p.c1_blocks = cs.in1 ? cs.in1->C / 64 : 0;
  p.cout = co; p.up = cs.up ? 1 : 0; p.relu = 1;
  if (cs.up) {
    for (int ph = 0; ph < 4; ++ph)
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) {
          int a, b, dy, dx;
          phase_taps(ph >> 1, ty, a, b, dy);
          phase_taps(ph & 1, tx, a, b, dx);
          p.dy[ph][ty * 2 + tx] = (int8_t)dy; p.dx[ph][ty * 2 + tx] = (int8_t)dx;
        }
  } else if (e->ks[slot] == 3) {
    for (int t = 0; t < 9; ++t) { p.dy[0][t] = (int8_t)(t / 3 - 1); p.dx[0][t] = (int8_t)(t % 3 - 1); }
  }
  p.bias = e->bias[slot];
  p.err_flag = e->err_flag;
  if (cs.epi == EPI_STORE) {
    p.out = cs.out->p; p.out_c = cs.out->C;
    for (int ph = 0; ph < p.n_phases; ++ph)
      if ((r = make_tmap_out(e, &p.tmOut[ph], *cs.out, L, cs.up, ph))) return r;
    if (cs.pool_out) {
      SD_REQUIRE(!cs.up && L.box_w % 2 == 0 && L.box_h % 2 == 0, "add_umma_conv(%s): fused pool needs even box dims", cs.name);
      if ((r = make_tmap_pool(e, &p.tmPool, *cs.pool_out, L.box_w / 2, L.box_h / 2, L.box_n))) return r;
      p.pool = 1;
    }
  }
  if (cs.epi == EPI_GATE) {
    p.out = cs.out->p; p.out_c = cs.out->C;
    p.psi_w = e->w_f32[SD_ATT5_PSI + 6 * cs.att]; p.psi_b = e->psi_b[cs.att];
    p.gate_x = cs.in1->p; p.gate_c = cs.in1->C;
    p.gate_tma = e->gate_tma;
    p.psi_out = cs.psi_out;
    if ((r = make_tmap_out(e, &p.tmOut[0], *cs.out, L, false, 0))) return r;
  }
  if (cs.epi == EPI_HEAD) { p.head_w = e->w_f32[SD_HEAD]; }
  if (cs.up && co == 64 && !cs.in1 && cs.epi == EPI_STORE && e->up4) {
    // all four sub-pixel phases per work item (conv_up4_kernel).  Weights re-packed as [tap][cb][phase in slot
    // order][co] rows of 64 halves, so a stage's B operand is one box of 64 / 128 / 256 rows (three maps).
    {
      const int cbs = cin_total / 64;
      std::vector<act_t> wp((size_t)16 * cbs * 64 * 64);
      std::vector<act_t> src((size_t)4 * co * K);
      SD_CUDA_CHECK(cudaMemcpy(src.data(), e->w_umma[slot], src.size() * 2, cudaMemcpyDeviceToHost));
      for (int t = 0; t < 9; ++t) {
        const Up4Tap tp = up4_tap(t);
        for (int cb = 0; cb < cbs; ++cb)
          for (int i = 0; i < tp.count; ++i) {
            const int ph = up4_slot_phase((tp.first + i) & 3);
            const int ty = tp.dy + 1 - (ph >> 1), tx = tp.dx + 1 - (ph & 1);
            const size_t row0 = ((size_t)tp.prefix * cbs + (size_t)cb * tp.count + i) * 64;
            for (int c2 = 0; c2 < 64; ++c2)
              for (int k = 0; k < 64; ++k)
                wp[(row0 + c2) * 64 + k] = src[((size_t)ph * co + c2) * K + (size_t)(ty * 2 + tx) * cin_total + cb * 64 + k];
          }
      }
      act_t* d_wp = nullptr;
      if ((r = upload(e, (void**)&d_wp, wp.data(), wp.size() * 2))) return r;
      const int rows = 16 * cbs * 64;
      if ((r = make_tmap_w(e, &p.tmB, d_wp, rows, 64, 64))) return r;
      if ((r = make_tmap_w(e, &p.tmA1, d_wp, rows, 64, 128))) return r;
      if ((r = make_tmap_w(e, &p.tmPool, d_wp, rows, 64, 256))) return r;
    }
    Op op4;
    op4.name = std::string(cs.name) + "[4ph]";
    op4.flops_per_tile = 2.0 * 4.0 * L.H * L.W * co * K;
    const int box_n4 = L.box_n, per_img4 = p.tiles_x * p.tiles_y, nsm4 = e->num_sms;
    op4.run = [p, box_n4, per_img4, nsm4](int B, cudaStream_t s) mutable -> int {
      p.B = B;
      p.m_tiles = per_img4 * ((B + box_n4 - 1) / box_n4);
      static PerDeviceOnce attr_once;
      if (attr_once.first()) SD_CUDA_CHECK(cudaFuncSetAttribute(conv_up4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Up4Cfg::kSmemBytes));
      conv_up4_kernel<<<p.m_tiles < nsm4 ? p.m_tiles : nsm4, Up4Cfg::kThreads, Up4Cfg::kSmemBytes, s>>>(p);
      SD_LAUNCH_CHECK("conv_up4_kernel");
      return SD_OK;
    };
    e->ops.push_back(op4);
    return SD_OK;
  }
  Op op;
  op.name = std::string(cs.name) + (p.pool ? "+pool" : "");
  const int epi = cs.epi;
  const int box_n = L.box_n, per_img = p.tiles_x * p.tiles_y;
  const int nsm = e->num_sms;
  op.flops_per_tile = 2.0 * (cs.up ? 4.0 : 1.0) * L.H * L.W * co * K;
  const bool cta2 = epi == EPI_STORE && ((e->cta2 && bn == 256) || (e->cta2_n128 && bn == 128 && mt == 2));
  if (cta2) {
    op.name += "[2sm]";
    // each CTA of a pair fetches its own half (128 / 64 rows) of the B k-block
    if ((r = make_tmap_w(e, &p.tmB, e->w_umma[slot], (cs.up ? 4 : 1) * co, K, bn / 2))) return r;
  }
  op.run = [e, p, bn, mt, epi, box_n, per_img, nsm, cta2](int B, cudaStream_t s) mutable -> int {
    p.B = B;
    p.m_tiles = per_img * ((B + box_n - 1) / box_n);
    if (epi == EPI_HEAD) {
      p.head_b = e->head_b; p.thr = e->thr;
      p.prob_f32 = e->o32; p.prob_f16 = e->o16; p.mask_u8 = e->omask; p.tile_dst = e->odst;
    }
    if (cta2) return bn == 256 ? launch_conv2<256>(p, nsm, s) : launch_conv2<128>(p, nsm, s);
    const int n_work = ((p.m_tiles + mt - 1) / mt) * p.n_tiles * p.n_phases;
    const int grid = n_work < nsm ? n_work : nsm;
    return dispatch_conv(p, bn, mt, epi, grid, s);
  };
  e->ops.push_back(op);
  return SD_OK;
}

static void add_simt_conv(sd_engine* e, const char* name, int slot, const Act* in0, const Act* in1, const Act* out,
                          bool up, float* out32) {
  Op op;
  op.name = name;
  op.run = [e, slot, in0, in1, out, up, out32](int B, cudaStream_t s) -> int {
    SimtConv a;
    a.in0 = in0->p; a.c0 = in0->C; a.in1 = in1 ? in1->p : nullptr; a.c1 = in1 ? in1->C : 0;
    a.w = e->w_simt[slot]; a.bias = e->bias[slot];
    a.out = out ? out->p : nullptr; a.out32 = out32;
    a.B = B; a.H = up ? in0->H * 2 : in0->H; a.W = up ? in0->W * 2 : in0->W;
    a.cout = e->cout[slot]; a.ks = e->ks[slot]; a.up = up; a.relu = 1;
    const int64_t M = (int64_t)B * a.H * a.W;
    dim3 grid((unsigned)(M / 64), (unsigned)((a.cout + 63) / 64));
    conv_simt_kernel<<<grid, 256, 0, s>>>(a);
    SD_LAUNCH_CHECK("conv_simt_kernel");
    return SD_OK;
  };
  e->ops.push_back(op);
}

static void add_pool(sd_engine* e, const char* name, const Act* in, const Act* out) {
  Op op;
  op.name = name;
  op.run = [in, out](int B, cudaStream_t s) -> int {
    const int C8 = in->C / 8;
    const int64_t total = (int64_t)B * out->H * out->W * C8;
    maxpool_kernel<<<ceil_div(total, 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(in->p), reinterpret_cast<uint4*>(out->p),
                                                        B, out->H, out->W, C8);
    SD_LAUNCH_CHECK("maxpool_kernel");
    return SD_OK;
  };
  e->ops.push_back(op);
}

}  // namespace sd

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" int sd_engine_create(int device, int max_tiles, int tile_h, int tile_w, sd_engine** out) {
  SD_REQUIRE(out, "sd_engine_create: null out");
  // 2048 tiles = 150 GB of activations, more than fits beside the weights; it also keeps every work-item count below 2^20,
  // the domain of the kernels' multiply-shift dividers (common.cuh FastDiv)
  SD_REQUIRE(max_tiles > 0 && max_tiles <= 2048, "sd_engine_create: max_tiles %d (1..2048)", max_tiles);
  SD_REQUIRE(tile_h == SD_TILE_H && tile_w == SD_TILE_W, "sd_engine_create: only %dx%d tiles are supported (got %dx%d)",
             SD_TILE_H, SD_TILE_W, tile_h, tile_w);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("sd_engine_create: no CUDA device (this library has no CPU fallback)");
    return SD_ECUDA;
  }
  SD_REQUIRE(device >= 0 && device < ndev, "sd_engine_create: device %d of %d", device, ndev);
  cudaDeviceProp prop;
  SD_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("sd_engine_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return SD_ECUDA;
  }
  sd_engine* e = new sd_engine();
  e->device = device; e->max_tiles = max_tiles; e->cap_tiles = (max_tiles + 1) / 2 * 2; e->H = tile_h; e->W = tile_w;
  e->num_sms = prop.multiProcessorCount;
  *out = e;
  return SD_OK;
}

extern "C" void sd_engine_destroy(sd_engine* e) {
  if (!e) return;
  DeviceGuard guard(e->device);
  for (void* p : e->allocs) cudaFree(p);
  if (e->err_flag_host) cudaFreeHost(e->err_flag_host);
  for (auto ev : e->ev) cudaEventDestroy(ev);
  delete e;
}

extern "C" int sd_engine_set_conv(sd_engine* e, int slot, const float* w, const float* b, int cout, int cin, int k) {
  SD_REQUIRE(e && w && b, "sd_engine_set_conv: null argument");
  SD_REQUIRE(slot >= 0 && slot < SD_NUM_SLOTS, "sd_engine_set_conv: slot %d", slot);
  SD_REQUIRE(!e->finalized, "sd_engine_set_conv: engine already finalized");
  SD_REQUIRE((k == 1 || k == 3) && cout > 0 && cin > 0, "sd_engine_set_conv: bad shape");
  e->hw[slot].assign(w, w + (size_t)cout * cin * k * k);
  e->hb[slot].assign(b, b + cout);
  e->cout[slot] = cout; e->cin[slot] = cin; e->ks[slot] = k;
  return SD_OK;
}

extern "C" int sd_engine_set_head_bias(sd_engine* e, float bias) {
  SD_REQUIRE(e, "sd_engine_set_head_bias: null engine");
  e->head_b = bias;
  return SD_OK;
}

extern "C" int sd_engine_finalize(sd_engine* e, int impl) {
  SD_REQUIRE(e && !e->finalized, "sd_engine_finalize: bad state");
  SD_REQUIRE(impl == 0 || impl == 1, "sd_engine_finalize: impl %d", impl);
  DeviceGuard guard(e->device);
  // expected shapes
  static const int exp_shape[SD_NUM_SLOTS][3] = {
      {64, 3, 3}, {64, 64, 3}, {128, 64, 3}, {128, 128, 3}, {256, 128, 3}, {256, 256, 3}, {512, 256, 3}, {512, 512, 3},
      {1024, 512, 3}, {1024, 1024, 3},
      {512, 1024, 3}, {256, 512, 1}, {256, 512, 1}, {1, 256, 1}, {512, 1024, 3}, {512, 512, 3},
      {256, 512, 3}, {128, 256, 1}, {128, 256, 1}, {1, 128, 1}, {256, 512, 3}, {256, 256, 3},
      {128, 256, 3}, {64, 128, 1}, {64, 128, 1}, {1, 64, 1}, {128, 256, 3}, {128, 128, 3},
      {64, 128, 3}, {32, 64, 1}, {32, 64, 1}, {1, 32, 1}, {64, 128, 3}, {64, 64, 3},
      {1, 64, 1}};
  for (int s = 0; s < SD_NUM_SLOTS; ++s) {
    SD_REQUIRE(!e->hw[s].empty(), "sd_engine_finalize: slot %d has no weights", s);
    SD_REQUIRE(e->cout[s] == exp_shape[s][0] && e->cin[s] == exp_shape[s][1] && e->ks[s] == exp_shape[s][2],
               "sd_engine_finalize: slot %d has shape (%d,%d,%d), expected (%d,%d,%d)", s, e->cout[s], e->cin[s], e->ks[s],
               exp_shape[s][0], exp_shape[s][1], exp_shape[s][2]);
  }
  e->impl = impl;
  if (const char* rm = getenv("SD_BAND")) e->band = atoi(rm);
  if (const char* bm = getenv("SD_BNMAX")) e->bn_max = atoi(bm);
  if (const char* m2 = getenv("SD_MT2")) e->mt2_max_bn = atoi(m2);
  if (const char* c1 = getenv("SD_CONV1TC")) e->conv1_tc = atoi(c1);
  if (const char* fp = getenv("SD_FUSEPOOL")) e->fuse_pool = atoi(fp);
  if (const char* u4 = getenv("SD_UP4")) e->up4 = atoi(u4);
  if (const char* gt = getenv("SD_GATETMA")) e->gate_tma = atoi(gt);
  if (const char* pf = getenv("SD_PSI_FUSED")) e->psi_fused = atoi(pf);
  if (const char* c2 = getenv("SD_CTA2")) e->cta2 = atoi(c2);
  if (const char* c2 = getenv("SD_CTA2_N128")) e->cta2_n128 = atoi(c2);
  if (!e->band) e->fuse_pool = 0;                   // the level-1 pool is fused in the band kernel only
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  SD_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
  e->encode = (PFN_tmapEncodeTiled)fn;

  const int H = e->H, W = e->W;
  e->lv[0] = {H, W, 128, 1, 1};
  e->lv[1] = {H / 2, W / 2, 64, 2, 1};
  e->lv[2] = {H / 4, W / 4, 32, 4, 1};
  e->lv[3] = {H / 8, W / 8, 16, 8, 1};
  e->lv[4] = {H / 16, W / 16, 8, 8, 2};
  int r;
  SD_CUDA_CHECK(cudaHostAlloc((void**)&e->err_flag_host, 16 * sizeof(int), cudaHostAllocMapped));
  for (int i = 0; i < 16; ++i) e->err_flag_host[i] = 0;
  SD_CUDA_CHECK(cudaHostGetDevicePointer((void**)&e->err_flag, e->err_flag_host, 0));

  // ---- weights ----
  for (int s = 0; s < SD_NUM_SLOTS; ++s) {
    if ((r = upload(e, (void**)&e->bias[s], e->hb[s].data(), e->hb[s].size() * 4))) return r;
    const bool is_psi = (s == SD_ATT5_PSI || s == SD_ATT4_PSI || s == SD_ATT3_PSI || s == SD_ATT2_PSI);
    const bool is_up = (s == SD_UP5 || s == SD_UP4 || s == SD_UP3 || s == SD_UP2);
    if (is_psi || s == SD_HEAD) {
      if ((r = upload(e, (void**)&e->w_f32[s], e->hw[s].data(), e->hw[s].size() * 4))) return r;
      if (is_psi) e->psi_b[(s - SD_ATT5_PSI) / 6] = e->hb[s][0];
      else e->head_b = e->hb[s][0];
      continue;
    }
    if (s == SD_CONV1_0) {
      std::vector<float> t(27 * 64);
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < 3; ++ci)
          for (int k = 0; k < 9; ++k) t[(k * 3 + ci) * 64 + co] = e->hw[s][((size_t)co * 3 + ci) * 9 + k];
      if ((r = upload(e, (void**)&e->w_f32[s], t.data(), t.size() * 4))) return r;
      // tensor-core form: [64 x K=80] fp16 in the canonical no-swizzle K-major layout, k = tap * 8 + ci
      std::vector<act_t> c1((size_t)kC1BBytes / 2, f2act(0.f));
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < 3; ++ci)
          for (int k = 0; k < 9; ++k)
            c1[(size_t)(co >> 3) * (kC1GroupBytes / 2) + k * 64 + (co & 7) * 8 + ci] = f2act(e->hw[s][((size_t)co * 3 + ci) * 9 + k]);
      if ((r = upload(e, (void**)&e->w_umma[s], c1.data(), c1.size() * 2))) return r;
    }
    const int co = e->cout[s], ci = e->cin[s], taps = e->ks[s] * e->ks[s];
    if (impl == 1) {
      const int cip = (s == SD_CONV1_0) ? 8 : ci;
      std::vector<act_t> t((size_t)taps * cip * co, f2act(0.f));
      for (int o = 0; o < co; ++o)
        for (int i = 0; i < ci; ++i)
          for (int k = 0; k < taps; ++k) t[((size_t)k * cip + i) * co + o] = f2act(e->hw[s][((size_t)o * ci + i) * taps + k]);
      if ((r = upload(e, (void**)&e->w_simt[s], t.data(), t.size() * 2))) return r;
    } else if (s != SD_CONV1_0) {
      const bool is_gx = (s == SD_ATT5_X || s == SD_ATT4_X || s == SD_ATT3_X || s == SD_ATT2_X);
      const bool is_gg = (s == SD_ATT5_G || s == SD_ATT4_G || s == SD_ATT3_G || s == SD_ATT2_G);
      if (is_gx) continue;   // packed together with the G slot below
      std::vector<act_t> t;
      if (is_gg) {
        // gate GEMM: K = [g channels | x channels], weights [W_g | W_x], bias b_g + b_x
        const int fg = ci, fl = e->cin[s + 1];
        t.assign((size_t)co * (fg + fl), f2act(0.f));
        for (int o = 0; o < co; ++o) {
          for (int i = 0; i < fg; ++i) t[(size_t)o * (fg + fl) + i] = f2act(e->hw[s][(size_t)o * fg + i]);
          for (int i = 0; i < fl; ++i) t[(size_t)o * (fg + fl) + fg + i] = f2act(e->hw[s + 1][(size_t)o * fl + i]);
        }
      } else {
        pack_umma(e->hw[s].data(), co, ci, e->ks[s], is_up, t);
      }
      if ((r = upload(e, (void**)&e->w_umma[s], t.data(), t.size() * 2))) return r;
    }
  }
  // gate bias = b_g + b_x (both folded), replacing the G slot's device bias
  for (int a = 0; a < 4; ++a) {
    const int sg = SD_ATT5_G + 6 * a, sx = sg + 1;
    std::vector<float> bsum(e->cout[sg]);
    for (int i = 0; i < e->cout[sg]; ++i) bsum[i] = e->hb[sg][i] + e->hb[sx][i];
    SD_CUDA_CHECK(cudaMemcpy(e->bias[sg], bsum.data(), bsum.size() * 4, cudaMemcpyHostToDevice));
  }

  // ---- activations ----
  Act* A = e->act;
  if ((r = alloc_act(e, e->c1a, 0, 64)) || (r = alloc_act(e, A[SD_TAP_X1], 0, 64)) || (r = alloc_act(e, e->p1, 1, 64)) ||
      (r = alloc_act(e, e->c2a, 1, 128)) || (r = alloc_act(e, A[SD_TAP_X2], 1, 128)) || (r = alloc_act(e, e->p2, 2, 128)) ||
      (r = alloc_act(e, e->c3a, 2, 256)) || (r = alloc_act(e, A[SD_TAP_X3], 2, 256)) || (r = alloc_act(e, e->p3, 3, 256)) ||
      (r = alloc_act(e, e->c4a, 3, 512)) || (r = alloc_act(e, A[SD_TAP_X4], 3, 512)) || (r = alloc_act(e, e->p4, 4, 512)) ||
      (r = alloc_act(e, e->c5a, 4, 1024)) || (r = alloc_act(e, A[SD_TAP_X5], 4, 1024)) ||
      (r = alloc_act(e, A[SD_TAP_D5U], 3, 512)) || (r = alloc_act(e, A[SD_TAP_A4], 3, 512)) || (r = alloc_act(e, e->u5a, 3, 512)) ||
      (r = alloc_act(e, A[SD_TAP_D5], 3, 512)) ||
      (r = alloc_act(e, A[SD_TAP_D4U], 2, 256)) || (r = alloc_act(e, A[SD_TAP_A3], 2, 256)) || (r = alloc_act(e, e->u4a, 2, 256)) ||
      (r = alloc_act(e, A[SD_TAP_D4], 2, 256)) ||
      (r = alloc_act(e, A[SD_TAP_D3U], 1, 128)) || (r = alloc_act(e, A[SD_TAP_A2], 1, 128)) || (r = alloc_act(e, e->u3a, 1, 128)) ||
      (r = alloc_act(e, A[SD_TAP_D3], 1, 128)) ||
      (r = alloc_act(e, A[SD_TAP_D2U], 0, 64)) || (r = alloc_act(e, A[SD_TAP_A1], 0, 64)) || (r = alloc_act(e, e->u2a, 0, 64)))
    return r;
  if (impl == 1) {
    if ((r = alloc_act(e, A[SD_TAP_D2], 0, 64))) return r;
    if ((r = dev_alloc(e, (void**)&e->qbuf, (size_t)e->max_tiles * H * W * 32 * sizeof(float)))) return r;
  }
  if (impl == 0 && e->psi_fused && (r = dev_alloc(e, (void**)&e->psi1, (size_t)e->cap_tiles * H * W * sizeof(float)))) return r;

  // ---- schedule ----
  Op first;
  first.name = "Conv1.0(simt)";
  first.flops_per_tile = 2.0 * H * W * 64 * 27;
  if (impl == 0 && e->conv1_tc && W % 128 == 0) {
    first.name = "Conv1.0";
    ConvFirstParams cp;
    memset(&cp, 0, sizeof(cp));
    if ((r = make_tmap_out(e, &cp.tmOut, e->c1a, e->lv[0], false, 0))) return r;
    cp.w = reinterpret_cast<const uint4*>(e->w_umma[SD_CONV1_0]); cp.bias = e->bias[SD_CONV1_0];
    cp.H = H; cp.W = W; cp.err_flag = e->err_flag;
    SD_CUDA_CHECK(cudaFuncSetAttribute(conv_first_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1SmemBytes));
    first.run = [e, cp](int B, cudaStream_t s) mutable -> int {
      cp.B = B; cp.in = reinterpret_cast<const uint4*>(e->in_tiles);
      const int n_work = B * e->H * (e->W / 128);
      const int grid = n_work < kC1CtasPerSm * e->num_sms ? n_work : kC1CtasPerSm * e->num_sms;
      conv_first_umma_kernel<<<grid, kC1Threads, kC1SmemBytes, s>>>(cp);
      SD_LAUNCH_CHECK("conv_first_umma_kernel");
      return SD_OK;
    };
  } else if (impl == 0) {
    first.run = [e](int B, cudaStream_t s) -> int {
      const int64_t total = (int64_t)B * e->H * e->W;
      conv_first_kernel<<<ceil_div(total, 128), 128, 0, s>>>(reinterpret_cast<const uint2*>(e->in_tiles), e->w_f32[SD_CONV1_0],
                                                             e->bias[SD_CONV1_0], e->c1a.p, B, e->H, e->W);
      SD_LAUNCH_CHECK("conv_first_kernel");
      return SD_OK;
    };
  } else {
    first.run = [e](int B, cudaStream_t s) -> int {
      SimtConv a;
      a.in0 = reinterpret_cast<const act_t*>(e->in_tiles); a.c0 = 8; a.in1 = nullptr; a.c1 = 0;
      a.w = e->w_simt[SD_CONV1_0]; a.bias = e->bias[SD_CONV1_0]; a.out = e->c1a.p; a.out32 = nullptr;
      a.B = B; a.H = e->H; a.W = e->W; a.cout = 64; a.ks = 3; a.up = 0; a.relu = 1;
      dim3 grid((unsigned)((int64_t)B * e->H * e->W / 64), 1);
      conv_simt_kernel<<<grid, 256, 0, s>>>(a);
      SD_LAUNCH_CHECK("conv_simt_kernel");
      return SD_OK;
    };
  }
  e->ops.push_back(first);

  auto conv = [&](const char* name, int slot, const Act* i0, const Act* i1, const Act* o, bool up) -> int {
    if (impl == 0) { ConvSpec cs{name, slot, i0, i1, o, up, EPI_STORE, -1}; return add_umma_conv(e, cs); }
    add_simt_conv(e, name, slot, i0, i1, o, up, nullptr);
    return SD_OK;
  };
  // conv followed by MaxPool2x2: fused into the conv's store epilogue on the tcgen05 path
  auto conv_pool = [&](const char* name, const char* pool_name, int slot, const Act* i0, const Act* o, const Act* po) -> int {
    if (impl == 0 && e->fuse_pool) {
      ConvSpec cs{name, slot, i0, nullptr, o, false, EPI_STORE, -1};
      cs.pool_out = po;
      return add_umma_conv(e, cs);
    }
    int rr = conv(name, slot, i0, nullptr, o, false);
    if (rr) return rr;
    add_pool(e, pool_name, o, po);
    return SD_OK;
  };
  auto gate = [&](const char* name, int att, const Act* g, const Act* x, const Act* o) -> int {
    const int sg = SD_ATT5_G + 6 * att;
    if (impl == 0) { ConvSpec cs{name, sg, g, x, o, false, EPI_GATE, att}; return add_umma_conv(e, cs); }
    // debug: q = relu(W_g g + W_x x + b) in fp32, then psi / scale
    Op op;
    op.name = name;
    op.run = [e, sg, att, g, x, o](int B, cudaStream_t s) -> int {
      SimtConv a;
      a.in0 = g->p; a.c0 = g->C; a.in1 = x->p; a.c1 = x->C;
      a.w = e->w_simt[sg]; a.bias = e->bias[sg]; a.out = nullptr; a.out32 = e->qbuf;
      a.B = B; a.H = g->H; a.W = g->W; a.cout = e->cout[sg]; a.ks = 1; a.up = 0; a.relu = 1;
      const int64_t M = (int64_t)B * a.H * a.W;
      dim3 grid((unsigned)(M / 64), (unsigned)((a.cout + 63) / 64));
      conv_simt_kernel<<<grid, 256, 0, s>>>(a);
      SD_LAUNCH_CHECK("conv_simt_kernel(gate)");
      gate_apply_kernel<<<ceil_div(M, 128), 128, 0, s>>>(e->qbuf, e->w_f32[sg + 2], e->psi_b[att], x->p, o->p, M, a.cout, x->C);
      SD_LAUNCH_CHECK("gate_apply_kernel");
      return SD_OK;
    };
    e->ops.push_back(op);
    return SD_OK;
  };
  if (impl == 1) {
    // debug gate weights: [1 tap][fg + fl][fint]
    for (int a = 0; a < 4; ++a) {
      const int sg = SD_ATT5_G + 6 * a, sx = sg + 1;
      const int fint = e->cout[sg], fg = e->cin[sg], fl = e->cin[sx];
      std::vector<act_t> t((size_t)(fg + fl) * fint);
      for (int o = 0; o < fint; ++o) {
        for (int i = 0; i < fg; ++i) t[(size_t)i * fint + o] = f2act(e->hw[sg][(size_t)o * fg + i]);
        for (int i = 0; i < fl; ++i) t[(size_t)(fg + i) * fint + o] = f2act(e->hw[sx][(size_t)o * fl + i]);
      }
      e->w_simt[sg] = nullptr;
      if ((r = upload(e, (void**)&e->w_simt[sg], t.data(), t.size() * 2))) return r;
    }
  }

  if ((r = conv_pool("Conv1.3", "pool1", SD_CONV1_1, &e->c1a, &A[SD_TAP_X1], &e->p1))) return r;
  if ((r = conv("Conv2.0", SD_CONV2_0, &e->p1, nullptr, &e->c2a, false))) return r;
  if ((r = conv_pool("Conv2.3", "pool2", SD_CONV2_1, &e->c2a, &A[SD_TAP_X2], &e->p2))) return r;
  if ((r = conv("Conv3.0", SD_CONV3_0, &e->p2, nullptr, &e->c3a, false))) return r;
  if ((r = conv_pool("Conv3.3", "pool3", SD_CONV3_1, &e->c3a, &A[SD_TAP_X3], &e->p3))) return r;
  if ((r = conv("Conv4.0", SD_CONV4_0, &e->p3, nullptr, &e->c4a, false))) return r;
  if ((r = conv_pool("Conv4.3", "pool4", SD_CONV4_1, &e->c4a, &A[SD_TAP_X4], &e->p4))) return r;
  if ((r = conv("Conv5.0", SD_CONV5_0, &e->p4, nullptr, &e->c5a, false))) return r;
  if ((r = conv("Conv5.3", SD_CONV5_1, &e->c5a, nullptr, &A[SD_TAP_X5], false))) return r;

  if ((r = conv("Up5", SD_UP5, &A[SD_TAP_X5], nullptr, &A[SD_TAP_D5U], true))) return r;
  if ((r = gate("Att5", 0, &A[SD_TAP_D5U], &A[SD_TAP_X4], &A[SD_TAP_A4]))) return r;
  if ((r = conv("Up_conv5.0", SD_UPCONV5_0, &A[SD_TAP_A4], &A[SD_TAP_D5U], &e->u5a, false))) return r;
  if ((r = conv("Up_conv5.3", SD_UPCONV5_1, &e->u5a, nullptr, &A[SD_TAP_D5], false))) return r;

  if ((r = conv("Up4", SD_UP4, &A[SD_TAP_D5], nullptr, &A[SD_TAP_D4U], true))) return r;
  if ((r = gate("Att4", 1, &A[SD_TAP_D4U], &A[SD_TAP_X3], &A[SD_TAP_A3]))) return r;
  if ((r = conv("Up_conv4.0", SD_UPCONV4_0, &A[SD_TAP_A3], &A[SD_TAP_D4U], &e->u4a, false))) return r;
  if ((r = conv("Up_conv4.3", SD_UPCONV4_1, &e->u4a, nullptr, &A[SD_TAP_D4], false))) return r;

  if ((r = conv("Up3", SD_UP3, &A[SD_TAP_D4], nullptr, &A[SD_TAP_D3U], true))) return r;
  if ((r = gate("Att3", 2, &A[SD_TAP_D3U], &A[SD_TAP_X2], &A[SD_TAP_A2]))) return r;
  if ((r = conv("Up_conv3.0", SD_UPCONV3_0, &A[SD_TAP_A2], &A[SD_TAP_D3U], &e->u3a, false))) return r;
  if ((r = conv("Up_conv3.3", SD_UPCONV3_1, &e->u3a, nullptr, &A[SD_TAP_D3], false))) return r;

  if ((r = conv("Up2", SD_UP2, &A[SD_TAP_D3], nullptr, &A[SD_TAP_D2U], true))) return r;
  // level-1 gate: psi-only form when its consumer is the band kernel (which scales the skip rows it stages); x1 * psi
  // then never exists in HBM (sd_unet_read_tap materialises it on demand)
  e->psi_live = impl == 0 && e->psi_fused && e->band && e->psi1;
  if (e->psi_live) {
    ConvSpec cg{"Att2", SD_ATT5_G + 6 * 3, &A[SD_TAP_D2U], &A[SD_TAP_X1], &A[SD_TAP_A1], false, EPI_GATE, 3};
    cg.psi_out = e->psi1;
    if ((r = add_umma_conv(e, cg))) return r;
    e->ops.back().name += "[psi]";
    ConvSpec cu{"Up_conv2.0", SD_UPCONV2_0, &A[SD_TAP_X1], &A[SD_TAP_D2U], &e->u2a, false, EPI_STORE, -1};
    cu.psi_in = e->psi1;
    if ((r = add_umma_conv(e, cu))) return r;
  } else {
    if ((r = gate("Att2", 3, &A[SD_TAP_D2U], &A[SD_TAP_X1], &A[SD_TAP_A1]))) return r;
    if ((r = conv("Up_conv2.0", SD_UPCONV2_0, &A[SD_TAP_A1], &A[SD_TAP_D2U], &e->u2a, false))) return r;
  }
  if (impl == 0) {
    ConvSpec cs{"Up_conv2.3+head", SD_UPCONV2_1, &e->u2a, nullptr, nullptr, false, EPI_HEAD, -1};
    if ((r = add_umma_conv(e, cs))) return r;
  } else {
    if ((r = conv("Up_conv2.3", SD_UPCONV2_1, &e->u2a, nullptr, &A[SD_TAP_D2], false))) return r;
    Op op;
    op.name = "head";
    op.run = [e](int B, cudaStream_t s) -> int {
      const int64_t M = (int64_t)B * e->H * e->W;
      head_kernel<<<ceil_div(M, 256), 256, 0, s>>>(e->act[SD_TAP_D2].p, e->w_f32[SD_HEAD], e->head_b, e->thr, e->o32, e->o16,
                                                   e->omask, M, 64);
      SD_LAUNCH_CHECK("head_kernel");
      return SD_OK;
    };
    e->ops.push_back(op);
  }
  e->ev.resize(e->ops.size() + 1);
  for (auto& ev : e->ev) SD_CUDA_CHECK(cudaEventCreate(&ev));
  e->last_ms.assign(e->ops.size(), 0.f);
  e->finalized = true;
  return SD_OK;
}

static int unet_forward(sd_engine* e, const void* d_tiles, int n_tiles, float bin_thr, float* d_prob_f32, void* d_prob_f16,
                        uint8_t* d_mask_u8, const sd_tile_dst* d_dst, void* stream) {
  SD_REQUIRE(e && e->finalized, "sd_unet_forward: engine not finalized");
  SD_REQUIRE(n_tiles >= 0 && n_tiles <= e->max_tiles, "sd_unet_forward: n_tiles %d exceeds max_tiles %d", n_tiles, e->max_tiles);
  if (n_tiles == 0) return SD_OK;   // evaluate_binarize.py:93-100 feeds an empty last minibatch when B % 8 == 0
  SD_REQUIRE(d_tiles, "sd_unet_forward: null input");
  SD_REQUIRE(!d_dst || e->impl == 0, "sd_unet_forward_lines: the debug implementation has no fused glue");
  cudaStream_t s = (cudaStream_t)stream;
  e->in_tiles = d_tiles; e->thr = bin_thr;
  e->o32 = d_prob_f32; e->o16 = reinterpret_cast<__half*>(d_prob_f16); e->omask = d_mask_u8; e->odst = d_dst;
  for (size_t i = 0; i < e->ops.size(); ++i) {
    if (e->timing) SD_CUDA_CHECK(cudaEventRecord(e->ev[i], s));
    int r = e->ops[i].run(n_tiles, s);
    if (r) {
      // a trapped kernel (bounded mbarrier wait) surfaces as a sticky CUDA error at the next call: name the wait
      if (r == SD_ECUDA && e->err_flag_host && *e->err_flag_host) {
        std::string msg = sd_last_error();
        set_error("%s [tcgen05 barrier timeout, wait code %d]", msg.c_str(), *e->err_flag_host);
      }
      return r;
    }
#ifdef SD_CONV_STATS
    if (e->timing) {   // debug build: per-op mbarrier wait cycles, printed as JSON lines on stderr
      unsigned long long wc[8];
      SD_CUDA_CHECK(cudaStreamSynchronize(s));
      SD_CUDA_CHECK(cudaMemcpyFromSymbol(wc, sd::g_wait_cycles, sizeof(wc)));
      unsigned long long z[8] = {};
      SD_CUDA_CHECK(cudaMemcpyToSymbol(sd::g_wait_cycles, z, sizeof(z)));
      fprintf(stderr, "{\"op\": \"%s\", \"wait\": [%llu, %llu, %llu, %llu, %llu, %llu]}\n", e->ops[i].name.c_str(), wc[0], wc[1],
              wc[2], wc[3], wc[4], wc[5]);
    }
#endif
  }
  if (e->timing) {
    SD_CUDA_CHECK(cudaEventRecord(e->ev[e->ops.size()], s));
    SD_CUDA_CHECK(cudaStreamSynchronize(s));
    for (size_t i = 0; i < e->ops.size(); ++i) SD_CUDA_CHECK(cudaEventElapsedTime(&e->last_ms[i], e->ev[i], e->ev[i + 1]));
    int flag = 0;
    flag = *e->err_flag_host;
    SD_REQUIRE(flag == 0, "sd_unet_forward: kernel barrier timeout (code %d)", flag);
  }
  return SD_OK;
}

extern "C" int sd_unet_forward(sd_engine* e, const void* d_tiles, int n_tiles, float bin_thr, float* d_prob_f32,
                               void* d_prob_f16, uint8_t* d_mask_u8, void* stream) {
  return unet_forward(e, d_tiles, n_tiles, bin_thr, d_prob_f32, d_prob_f16, d_mask_u8, nullptr, stream);
}

extern "C" int sd_unet_forward_lines(sd_engine* e, const void* d_tiles, int n_tiles, float bin_thr, const sd_tile_dst* d_dst,
                                     void* stream) {
  SD_REQUIRE(d_dst || n_tiles == 0, "sd_unet_forward_lines: null destination table");
  return unet_forward(e, d_tiles, n_tiles, bin_thr, nullptr, nullptr, nullptr, d_dst, stream);
}

extern "C" int sd_unet_read_tap(sd_engine* e, int tap, int n_tiles, void* d_out, size_t out_bytes, int* c, int* h, int* w,
                                void* stream) {
  SD_REQUIRE(e && e->finalized, "sd_unet_read_tap: engine not finalized");
  SD_REQUIRE(tap >= 0 && tap < SD_NUM_TAPS, "sd_unet_read_tap: tap %d", tap);
  const Act& a = e->act[tap];
  SD_REQUIRE(a.p, "sd_unet_read_tap: tap %d is not materialised by this implementation", tap);
  if (c) *c = a.C;
  if (h) *h = a.H;
  if (w) *w = a.W;
  if (!d_out) return SD_OK;
  const size_t need = (size_t)n_tiles * a.H * a.W * a.C * sizeof(act_t);
  SD_REQUIRE(out_bytes >= need, "sd_unet_read_tap: buffer too small (%zu < %zu)", out_bytes, need);
  if (tap == SD_TAP_A1 && e->psi_live) {      // psi-only gate: x1 * psi exists only inside the consumer; rebuild it for the reader
    const int64_t px = (int64_t)n_tiles * a.H * a.W;
    scale_by_psi_kernel<<<ceil_div(px * (a.C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(e->act[SD_TAP_X1].p), e->psi1, reinterpret_cast<uint4*>(a.p), px, a.C / 8);
    SD_LAUNCH_CHECK("scale_by_psi_kernel");
  }
  SD_CUDA_CHECK(cudaMemcpyAsync(d_out, a.p, need, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return SD_OK;
}

extern "C" int sd_engine_enable_timing(sd_engine* e, int on) {
  SD_REQUIRE(e, "sd_engine_enable_timing: null engine");
  e->timing = on != 0;
  return SD_OK;
}

extern "C" int sd_engine_layer_times(sd_engine* e, float* h_ms, int cap, int* n_out) {
  SD_REQUIRE(e && e->finalized && n_out, "sd_engine_layer_times: bad argument");
  *n_out = (int)e->ops.size();
  for (int i = 0; i < cap && i < (int)e->ops.size(); ++i) h_ms[i] = e->last_ms[i];
  return SD_OK;
}

extern "C" const char* sd_engine_layer_name(sd_engine* e, int i) {
  if (!e || i < 0 || i >= (int)e->ops.size()) return "";
  return e->ops[i].name.c_str();
}

// debug: cycles spent in mbarrier waits per wait code since the last reset (zeros unless built with -DSD_CONV_STATS)
extern "C" int sd_debug_wait_cycles(unsigned long long* h_out8, int reset) {
#ifdef SD_CONV_STATS
  SD_CUDA_CHECK(cudaDeviceSynchronize());
  if (h_out8) SD_CUDA_CHECK(cudaMemcpyFromSymbol(h_out8, sd::g_wait_cycles, 8 * sizeof(unsigned long long)));
  if (reset) { unsigned long long z[8] = {}; SD_CUDA_CHECK(cudaMemcpyToSymbol(sd::g_wait_cycles, z, sizeof(z))); }
#else
  if (h_out8) for (int i = 0; i < 8; ++i) h_out8[i] = 0;
  (void)reset;
#endif
  return SD_OK;
}

// debug: code of the mbarrier wait that timed out (0 = none).  The flag lives in pinned host memory, so it can
// be read after the trapped kernel has poisoned the CUDA context.
extern "C" int sd_engine_wait_error(sd_engine* e) { return (e && e->err_flag_host) ? *e->err_flag_host : -1; }
extern "C" int sd_engine_debug_word(sd_engine* e, int i) { return (e && e->err_flag_host && i >= 0 && i < 16) ? e->err_flag_host[i] : -1; }
