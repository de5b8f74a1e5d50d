// Implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
//   D[128 px, BN] (fp32, TMEM) += A[128 px, 64 ch of one tap] (fp16, smem) x B[BN, 64]^T (fp16, smem)
//
// * activations are NHWC fp16; one A k-block is ONE 4-D TMA box {64 ch, bw, bh, bn}
//   (bw*bh*bn == 128 pixels) fetched at the tap's (dx, dy) offset — the TMA unit
//   zero-fills out-of-image coordinates, which is exactly the conv's padding=1
//   per tile (tiles are independent images, no cross-tile halo).
// * the box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle,
//   i.e. the canonical K-major SWIZZLE_128B UMMA operand; weights [Cout][K] land the
//   same way, so both operands are described by plain UMMA smem descriptors.
// * a second source tensor continues the K loop (channel concat without a cat).
// * nearest-x2 upsample + conv3x3 is run as 4 sub-pixel phases, each a 2x2-tap conv
//   on the LOW-res input with pre-summed weights (2.25x fewer MACs, no upsampled
//   tensor ever exists).
// * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
//   warps 2..5 = epilogue (TMEM -> registers -> fused op -> global).  TMEM holds two
//   accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
// * epilogues: bias+ReLU store | attention gate (ReLU, dot with psi, sigmoid, scale
//   the skip tensor) | head (ReLU, 1x1 conv to one channel, sigmoid, threshold).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace sd {

enum { EPI_STORE = 0, EPI_GATE = 1, EPI_HEAD = 2 };

struct ConvParams {
  CUtensorMap tmA0, tmA1, tmB;
  CUtensorMap tmOut[4];        // store epilogue: one output map per sub-pixel phase
  CUtensorMap tmPool;          // store epilogue with pool != 0: the 2x2 max-pooled copy of the output
  int pool;
  int gate_tma;                // gate epilogue: skip tensor through shared memory (TMA load, scale in place, TMA store)
  float bias_c[64], vec_c[64]; // band kernel: bias and head vector in the constant bank (no smem reads in the epilogue)
  // geometry of the (low-res for up-convs) input grid the M tiles walk over
  int H, W, B;                 // image dims of the A source, live batch
  int box_w, box_h, box_n;     // pixels per M tile = box_w*box_h*box_n = 128
  int tiles_x, tiles_y;        // W/box_w, H/box_h
  int m_tiles, n_tiles;        // m_tiles = tiles_x*tiles_y*ceil(B/box_n)
  int n_phases;                // 1, or 4 for the sub-pixel up-conv
  int n_taps;                  // taps per phase (9, 4 or 1)
  int c0_blocks, c1_blocks;    // 64-channel k-blocks per tap from source 0 / 1
  int cout;                    // output channels (rows per phase in the weight matrix)
  int up;                      // 1: output pixel = (2y+py, 2x+px) on a 2H x 2W grid
  int relu;
  int8_t dy[4][9], dx[4][9];
  const float* bias;           // [cout]
  act_t* out;                 // NHWC fp16, channels = out_c
  int out_c;
  // gate epilogue
  const float* psi_w;          // [cout] fp32
  float psi_b;
  const act_t* gate_x;        // skip tensor to scale, NHWC with gate_c channels
  int gate_c;
  float* psi_out;              // gate epilogue, psi-only form: the gate writes sigma(psi) (one fp32 per pixel) and nothing else;
                               // the consumer (band kernel, PSI = true) scales the skip rows it stages with it
  const float* psi_in;         // band kernel, PSI = true: the plane the gate wrote
  // head epilogue
  const float* head_w;         // [cout]
  float head_b, thr;
  float* prob_f32; __half* prob_f16; uint8_t* mask_u8;
  const sd_tile_dst* tile_dst; // head: per tile, where its columns live in the packed line planes (glue fused into the head)
  int* err_flag;               // set when a barrier wait times out
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
#ifdef SD_CONV_STATS
// debug build only (SD_EXTRA_NVCC_FLAGS=-DSD_CONV_STATS): cycles spent in mbar_wait per wait code, summed over all
// calling threads; [0] counts kernel cycles of thread 0 of every CTA.  Read with sd_debug_wait_cycles().
__device__ unsigned long long g_wait_cycles[8];
#endif
// Bounded wait: a descriptor / protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
#ifdef SD_CONV_STATS
  const long long t0 = clock64();
#endif
  // the bound depends on the role so that, when a pipeline deadlocks, the wait closest to the cause reports:
  // MMA<-full (3) first, then producer<-empty (1), MMA<-epilogue (2), epilogue<-MMA (4)
  const int role = code % 10;
  const uint32_t bound = role == 3 ? 20000000u : (role == 1 ? 30000000u : (role == 2 ? 40000000u : 50000000u));
#pragma unroll 1
  for (uint32_t it = 0; it < bound; ++it) {
    if (mbar_try_wait(bar, parity)) {
#ifdef SD_CONV_STATS
      atomicAdd(&g_wait_cycles[code & 7], (unsigned long long)(clock64() - t0));
#endif
      return;
    }
  }
  if (err_flag) atomicExch(err_flag, code);
  __threadfence_system();
  asm volatile("trap;");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// One lane of a CONVERGED warp, chosen by the hardware.  The producer and MMA-issuer warps must pick their single issuing
// thread this way: under `if (lane == 0)` the compiler cannot prove that one thread is active and wraps EVERY tcgen05.mma /
// TMA instruction (uniform-datapath instructions) in an ELECT / BRA.U.ANY loop over the active lanes — ~50 cycles of issue
// per MMA, which made the issuing thread the bottleneck of every kernel whose MMAs are shorter than that (N <= 128:
// 64 tensor cycles; measured with tools/micro/umma_rate.cu, profiles/r02_umma_rate.txt).  After elect.sync it emits the
// bare instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B operand tile whose rows are 128 B apart and 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4)   // start address  [0,14)
         | (1ull << 16)                            // leading byte offset (unused for swizzled K-major)
         | (64ull << 32)                           // stride byte offset = 1024 B >> 4
         | (1ull << 46)                            // descriptor version (sm_100)
         | (2ull << 61);                           // SWIZZLE_128B
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barrier for the 4 epilogue warps only (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
template <int N> __device__ __forceinline__ void epi_bar_n() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  act2_t r = __hmax2(*reinterpret_cast<act2_t*>(&a), *reinterpret_cast<act2_t*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 hmax2_v4(uint4 a, uint4 b) {
  return make_uint4(hmax2_u32(a.x, b.x), hmax2_u32(a.y, b.y), hmax2_u32(a.z, b.z), hmax2_u32(a.w, b.w));
}
// 16-B chunk `ch` of row `r` of a [rows x 128 B] tile stored with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128(uint32_t base, int r, int ch) {
  return base + (uint32_t)r * 128u + ((uint32_t)(ch ^ (r & 7)) << 4);
}

// ---- store epilogue arithmetic -----------------------------------------------------------------------------------
// The four epilogue warps sit one per scheduler, so nothing hides their latencies: on the short-K layers (Conv2.x, Up2,
// Up3, the level-1 bands) the epilogue, not the tensor pipe, was the critical path (ncu source view: 23 % of the stall
// samples of Conv2.0 waited on scalar shared-memory bias loads, 14 % of its instructions were the ReLU).  So: bias as
// float4 broadcast loads (8 instead of 32 per 32 columns) or constant-bank operands, ReLU folded into the convert
// (cvt.rn.relu), one 16-byte shared store per 8 channels.
// 32 accumulator columns of this thread's pixel row -> chunks 4*cc .. 4*cc+3 of its 128-byte row in the staged tile
template <bool RELU>
__device__ __forceinline__ void epi_pack32(const float (&v)[32], const float* __restrict__ bias32, uint32_t row_base, int row, int cc) {
  const float4* b4 = reinterpret_cast<const float4*>(bias32);        // shared memory, 16-byte aligned
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 b0 = b4[2 * j4], b1 = b4[2 * j4 + 1];
    const int c0 = j4 * 8;
    uint4 pk;
    if (RELU) {
      pk.x = floats2act2_relu_u32(v[c0] + b0.x, v[c0 + 1] + b0.y); pk.y = floats2act2_relu_u32(v[c0 + 2] + b0.z, v[c0 + 3] + b0.w);
      pk.z = floats2act2_relu_u32(v[c0 + 4] + b1.x, v[c0 + 5] + b1.y); pk.w = floats2act2_relu_u32(v[c0 + 6] + b1.z, v[c0 + 7] + b1.w);
    } else {
      pk.x = floats2act2_u32(v[c0] + b0.x, v[c0 + 1] + b0.y); pk.y = floats2act2_u32(v[c0 + 2] + b0.z, v[c0 + 3] + b0.w);
      pk.z = floats2act2_u32(v[c0 + 4] + b1.x, v[c0 + 5] + b1.y); pk.w = floats2act2_u32(v[c0 + 6] + b1.z, v[c0 + 7] + b1.w);
    }
    st_shared_v4(row_base + ((((uint32_t)(cc * 4 + j4)) ^ (uint32_t)(row & 7)) << 4), pk);
  }
}
// the same with the bias in the constant bank (kernel parameters; `bias32` must be indexed with compile-time offsets
// after unrolling so that every add takes its bias as an immediate constant operand)
__device__ __forceinline__ void epi_pack32_const_relu(const float (&v)[32], const float* bias32, uint32_t row_base, int row, int cc) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const int c0 = j4 * 8;
    uint4 pk;
    pk.x = floats2act2_relu_u32(v[c0] + bias32[c0], v[c0 + 1] + bias32[c0 + 1]); pk.y = floats2act2_relu_u32(v[c0 + 2] + bias32[c0 + 2], v[c0 + 3] + bias32[c0 + 3]);
    pk.z = floats2act2_relu_u32(v[c0 + 4] + bias32[c0 + 4], v[c0 + 5] + bias32[c0 + 5]); pk.w = floats2act2_relu_u32(v[c0 + 6] + bias32[c0 + 6], v[c0 + 7] + bias32[c0 + 7]);
    st_shared_v4(row_base + ((((uint32_t)(cc * 4 + j4)) ^ (uint32_t)(row & 7)) << 4), pk);
  }
}

constexpr int kMiscBytes = 4096;       // barriers | tmem slot | bias | psi/head vector | gate scale | pixel index
constexpr int kMaxSmem = 232448;       // 227 KB opt-in limit per CTA

// MT = M tiles (128 pixels each) that share one B k-block in shared memory.  A 128 x 128 tile reads
// 32 KB of operands per 256 MMA cycles = the full 128 B/clk of shared-memory bandwidth and stalls the
// tensor pipe; 128 x 256 (BN = 256) or 2 x (128 x 128) (MT = 2) needs 96 B/clk.
#ifndef SD_EPI_WARPS
#define SD_EPI_WARPS 8
#endif
template <int BN, int EPI, int MT = 1> struct ConvCfg {
  static constexpr int kABytes = 128 * 128;             // 128 px x 64 halves, per M tile
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = MT * kABytes + kBBytes;
  // store: swizzled staging of ONE 64-channel half for the TMA store + its 2x2-pooled copy; gate: two 64-channel
  // blocks of the skip tensor (TMA in, scaled in place, TMA out)
  static constexpr int kOutBytes = (EPI == EPI_STORE) ? 128 * 64 * 2 + 4096 : (EPI == EPI_GATE ? 2 * 16384 : 0);
  static constexpr int kFit = (kMaxSmem - 1024 - kMiscBytes - kOutBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kTmemCols = (2 * MT * BN < 32) ? 32 : 2 * MT * BN;   // 64,128,256,512: powers of two
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 /*align slack*/ + kMiscBytes;
  static constexpr uint32_t kIdesc = kIdescBase | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
  // store epilogue: EIGHT warps, two per TMEM lane quarter (each takes 32 of the 64 staged channels of its pixel rows), so
  // every scheduler holds two epilogue warps that hide each other's TMEM-load / shared-memory / barrier latencies; with
  // four (one per scheduler) the epilogue was the critical path of the short-K layers.  Gate / head: four
  // (thread = pixel row reduces over all columns).
  static constexpr int kEpiWarps = (EPI == EPI_STORE) ? SD_EPI_WARPS : 4;
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  static constexpr int kThreads = 64 + kEpiThreads;
  static_assert(kEpiWarps == 4 || kEpiWarps == 8, "4 or 8 epilogue warps");
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(kTmemCols <= 512, "TMEM has 512 columns");
  static_assert(EPI != EPI_STORE || BN % 64 == 0, "store epilogue writes 64-channel boxes");
  static_assert(MT == 1 || EPI == EPI_STORE, "multi-M tiles are implemented for the store epilogue");
};

constexpr int kConvThreads = 192;

template <int BN, int EPI, int MT>
__global__ void __launch_bounds__((ConvCfg<BN, EPI, MT>::kThreads), 1) conv_umma_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, EPI, MT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t out_base = smem_base + Cfg::kStages * Cfg::kStageBytes;            // 1024-aligned
  uint8_t* misc = smem_al + Cfg::kStages * Cfg::kStageBytes + Cfg::kOutBytes;
  const uint32_t bar_base = out_base + Cfg::kOutBytes;
  // misc layout: [0,256) barriers (8 B each): full[S], empty[S], tfull[2], tempty[2]; [248] TMEM slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  auto xbar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 4 + s); };      // gate: skip-tensor block landed
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(misc + 248);
  float* s_bias = reinterpret_cast<float*>(misc + 256);        // up to 256 floats
  float* s_vec = reinterpret_cast<float*>(misc + 1280);        // psi / head weights, up to 256 floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    if (p.c1_blocks) tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmB);
    if (EPI == EPI_STORE) for (int i = 0; i < p.n_phases; ++i) tma_prefetch_desc(&p.tmOut[i]);
    if (EPI == EPI_STORE && p.pool) tma_prefetch_desc(&p.tmPool);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), Cfg::kEpiThreads); mbar_init(xbar(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (EPI != EPI_STORE && warp >= 2) {
    // whole-N vectors used by the gate / head epilogues (n_tiles == 1 there)
    const int t = threadIdx.x - 64;
    for (int i = t; i < BN; i += 128) {
      s_bias[i] = p.bias[i];
      s_vec[i] = (EPI == EPI_GATE) ? p.psi_w[i] : p.head_w[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_groups = (p.m_tiles + MT - 1) / MT;
  const int n_work = m_groups * p.n_tiles * p.n_phases;
  const int kb_per_tap = p.c0_blocks + p.c1_blocks;
  const int num_kb = p.n_taps * kb_per_tap;
  // work-item and tile indices are decomposed with multiply-shift dividers (the epilogue warps are the critical path
  // of the short-K layers; four integer divides per M tile were 8 % of their stall samples)
  const FastDiv fd_tx((uint32_t)p.tiles_x), fd_ty((uint32_t)p.tiles_y), fd_nt((uint32_t)p.n_tiles), fd_mg((uint32_t)m_groups);
  // M tile index -> pixel origin; a phantom tile (odd tail of an MT group) lands beyond the batch, where the
  // TMA unit zero-fills loads and clips stores
  auto origin = [&](int mt, int& x0, int& y0, int& n0) {
    uint32_t rest, tx, tn, ty;
    fd_tx.divmod((uint32_t)mt, rest, tx); fd_ty.divmod(rest, tn, ty);
    x0 = (int)tx * p.box_w; y0 = (int)ty * p.box_h; n0 = (int)tn * p.box_n;
  };

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        uint32_t nt_u, rest_u, mg_u, ph_u;
        fd_nt.divmod((uint32_t)w, rest_u, nt_u); fd_mg.divmod(rest_u, ph_u, mg_u);
        const int nt = (int)nt_u, mg = (int)mg_u, ph = (int)ph_u;
        int x0[MT], y0[MT], n0[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) origin(mg * MT + m, x0[m], y0[m], n0[m]);
        const int brow = ph * p.cout + nt * BN;
        for (int tap = 0; tap < p.n_taps; ++tap) {
          const int dx = p.dx[ph][tap], dy = p.dy[ph][tap];
          for (int cb = 0; cb < kb_per_tap; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1);
            mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
            const CUtensorMap* tm = cb < p.c0_blocks ? &p.tmA0 : &p.tmA1;
            const int c = (cb < p.c0_blocks ? cb : cb - p.c0_blocks) * 64;
#pragma unroll
            for (int m = 0; m < MT; ++m)
              tma_load_4d(a_dst + m * Cfg::kABytes, tm, full_bar(stage), c, x0[m] + dx, y0[m] + dy, n0[m]);
            tma_load_2d(a_dst + MT * Cfg::kABytes, &p.tmB, full_bar(stage), (tap * kb_per_tap + cb) * 64, brow);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u, p.err_flag, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * MT * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          const uint64_t bdesc = umma_desc_sw128(a_addr + MT * Cfg::kABytes);
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const uint64_t adesc = umma_desc_sw128(a_addr + m * Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // UMMA_K = 16 halves = 32 B -> +2 in the (addr >> 4) field
              umma_f16(d_tmem + (uint32_t)(m * BN), adesc + 2u * k, bdesc + 2u * k, Cfg::kIdesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(empty_bar(stage));          // frees the smem slot when these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));               // accumulators complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue (warps 2..5, store: 2..9) =======================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                // accumulator row == pixel within the M tile
    const int et = threadIdx.x - 64;              // 0 .. kEpiThreads-1 within the epilogue group
    const int eh = (warp - 2) >> 2;               // store epilogue with 8 warps: which 32 of the 64 staged channels
    const int lw = row % p.box_w; int rr = row / p.box_w;
    const int lh = rr % p.box_h; const int ln = rr / p.box_h;
    int as = 0; uint32_t aphase = 0;
    int cur_nt = -1;
    uint32_t xphase[2] = {0u, 0u};                // gate: parity of the two skip-tensor buffers
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      uint32_t nt_u, rest_u, mg_u, ph_u;
      fd_nt.divmod((uint32_t)w, rest_u, nt_u); fd_mg.divmod(rest_u, ph_u, mg_u);
      const int nt = (int)nt_u, mg = (int)mg_u, ph = (int)ph_u;

      if constexpr (EPI == EPI_STORE) {
        if (nt != cur_nt) {                       // bias slice of this N tile -> smem (uniform branch)
          epi_bar_n<Cfg::kEpiThreads>();                              // everybody is done reading the previous slice
          for (int i = et; i < BN; i += Cfg::kEpiThreads) s_bias[i] = __ldg(p.bias + nt * BN + i);
          cur_nt = nt;
          epi_bar_n<Cfg::kEpiThreads>();
        }
        mbar_wait(tfull_bar(as), aphase, p.err_flag, 4);
        tc_fence_after();
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * MT + m) * BN);
          int x0, y0, n0;
          origin(mg * MT + m, x0, y0, n0);
#pragma unroll 1
          for (int hb = 0; hb < BN / 64; ++hb) {
            // the previous TMA store must have finished READING the staging buffer
            if (et == 0) tma_store_wait_read();
            epi_bar_n<Cfg::kEpiThreads>();
#pragma unroll
            for (int cc = (Cfg::kEpiWarps == 8 ? eh : 0); cc < (Cfg::kEpiWarps == 8 ? eh + 1 : 2); ++cc) {
              const int c = hb * 2 + cc;
              float v[32];
              tmem_ld32(taddr + c * 32, v);
              // 128 rows x 128 B, 16-B chunk j of row r lives at chunk (j ^ (r & 7))
              const uint32_t row_base = out_base + (uint32_t)row * 128u;
              if (p.relu) epi_pack32<true>(v, s_bias + c * 32, row_base, row, cc);
              else epi_pack32<false>(v, s_bias + c * 32, row_base, row, cc);
            }
            if (m == MT - 1 && hb == BN / 64 - 1) {
              tc_fence_before();
              mbar_arrive(tempty_bar(as));        // TMEM stage free: the next work item's MMAs may start
            }
            fence_async_smem();                   // generic-proxy smem writes -> visible to the TMA unit
            epi_bar_n<Cfg::kEpiThreads>();
            if (et == 0) {
              tma_store_4d(&p.tmOut[ph], out_base, nt * BN + hb * 64, x0, y0, n0);
              tma_store_commit();
            }
            if (p.pool) {
              // fused MaxPool2x2: the staged tile holds whole 2x2 windows (box dims are even); 32 pooled pixels x 8 chunks
              const int pw = p.box_w >> 1, phh = p.box_h >> 1;
              const int pw_sh = 31 - __clz(pw), ph_sh = 31 - __clz(phh);
              const uint32_t pool_base = out_base + 16384u;
#pragma unroll
              for (int task = et; task < 256; task += Cfg::kEpiThreads) {
                const int pp = task >> 3, ch = task & 7;
                const int px = pp & (pw - 1); const int r2 = pp >> pw_sh;       // box dims are powers of two
                const int py = r2 & (phh - 1), pn = r2 >> ph_sh;
                const int m00 = (pn * p.box_h + 2 * py) * p.box_w + 2 * px, m10 = m00 + p.box_w;
                const uint4 v = hmax2_v4(hmax2_v4(ld_shared_v4(sw128(out_base, m00, ch)), ld_shared_v4(sw128(out_base, m00 + 1, ch))),
                                         hmax2_v4(ld_shared_v4(sw128(out_base, m10, ch)), ld_shared_v4(sw128(out_base, m10 + 1, ch))));
                st_shared_v4(sw128(pool_base, pp, ch), v);
              }
              fence_async_smem();
              epi_bar_n<Cfg::kEpiThreads>();
              if (et == 0) {
                tma_store_4d(&p.tmPool, pool_base, nt * BN + hb * 64, x0 >> 1, y0 >> 1, n0);
                tma_store_commit();
              }
            }
          }
        }
      } else {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
        uint32_t r2_u, tx_u, tn_u, ty_u;
        fd_tx.divmod((uint32_t)mg, r2_u, tx_u); fd_ty.divmod(r2_u, tn_u, ty_u);
        const int tx = (int)tx_u, ty = (int)ty_u, tn = (int)tn_u;
        const int n = tn * p.box_n + ln;
        int y = ty * p.box_h + lh, x = tx * p.box_w + lw;
        const int64_t pix = ((int64_t)n * p.H + y) * p.W + x;     // gate / head never upsample
        const bool live = n < p.B;
        // gate: the skip-tensor row does not depend on the GEMM -> issue its first loads now so
        // their latency hides behind the wait for the accumulator
        uint4 xpre[8];
        const bool gate_tma = EPI == EPI_GATE && p.gate_tma;
        const bool psi_only = EPI == EPI_GATE && p.psi_out != nullptr;
        const int gx0 = tx * p.box_w, gy0 = ty * p.box_h, gn0 = tn * p.box_n;
        if constexpr (EPI == EPI_GATE) {
          if (psi_only) {
            // nothing to prefetch: the skip tensor is scaled by its consumer
          } else if (gate_tma) {
            // first 64-channel block of the skip tensor -> buffer 0, in flight while the accumulator is awaited
            if (et == 0) {
              tma_store_wait_read();                    // the previous tile's stores are done reading the buffers
              mbar_expect_tx(xbar(0), 16384);
              tma_load_4d(out_base, &p.tmA1, xbar(0), 0, gx0, gy0, gn0);
            }
          } else {
            const uint4* xi = reinterpret_cast<const uint4*>(p.gate_x + (live ? pix : 0) * p.gate_c);
#pragma unroll
            for (int c = 0; c < 8; ++c) xpre[c] = __ldg(xi + c);
          }
        }
        mbar_wait(tfull_bar(as), aphase, p.err_flag, 4);
        tc_fence_after();
        float dot = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = v[j] + s_bias[c * 32 + j];
            if (EPI == EPI_GATE || p.relu) a = fmaxf(a, 0.f);
            dot = fmaf(a, s_vec[c * 32 + j], dot);
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(as));
        if constexpr (EPI == EPI_GATE) {
          // thread == pixel row: stream this pixel's skip-tensor row (gate_c halves, contiguous), scale, store.
          // (A 128-thread "coalesced" sweep over the tile measured slower: each row is a whole number of
          // 128-B lines, so the per-row walk already moves full lines.)
          const float sc = 1.f / (1.f + expf(-(dot + p.psi_b)));
          if (psi_only) {
            if (live) p.psi_out[pix] = sc;
          } else if (gate_tma) {
            const int nhb = p.gate_c >> 6;
#pragma unroll 1
            for (int hb = 0; hb < nhb; ++hb) {
              const int b = hb & 1;
              const uint32_t buf = out_base + (uint32_t)b * 16384u;
              if (hb + 1 < nhb && et == 0) {            // next block into the other buffer once its last store has been read
                tma_store_wait_read();
                mbar_expect_tx(xbar(b ^ 1), 16384);
                tma_load_4d(out_base + (uint32_t)(b ^ 1) * 16384u, &p.tmA1, xbar(b ^ 1), (hb + 1) * 64, gx0, gy0, gn0);
              }
              mbar_wait(xbar(b), xphase[b], p.err_flag, 6);
              xphase[b] ^= 1u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t addr = sw128(buf, row, j);
                uint4 t = ld_shared_v4(addr);
                act2_t* h = reinterpret_cast<act2_t*>(&t);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float2 f = act22float2(h[k]);
                  h[k] = floats2act2(f.x * sc, f.y * sc);
                }
                st_shared_v4(addr, t);
              }
              fence_async_smem();
              epi_bar();
              if (et == 0) { tma_store_4d(&p.tmOut[0], buf, hb * 64, gx0, gy0, gn0); tma_store_commit(); }
            }
          } else if (live) {
            const uint4* xi = reinterpret_cast<const uint4*>(p.gate_x + pix * p.gate_c);
            uint4* xo = reinterpret_cast<uint4*>(p.out + pix * p.out_c);
            const int nchunk = p.gate_c >> 3;            // 16-B chunks per row: 8, 16, 32 or 64
            for (int c0 = 0; c0 < nchunk; c0 += 8) {
              uint4 nxt[8];
              const bool more = c0 + 8 < nchunk;
              if (more) {
#pragma unroll
                for (int c = 0; c < 8; ++c) nxt[c] = __ldg(xi + c0 + 8 + c);   // next batch in flight while this one is scaled
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                uint4 t = xpre[c];
                act2_t* h = reinterpret_cast<act2_t*>(&t);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 f = act22float2(h[j]);
                  h[j] = floats2act2(f.x * sc, f.y * sc);
                }
                xo[c0 + c] = t;
              }
              if (more) {
#pragma unroll
                for (int c = 0; c < 8; ++c) xpre[c] = nxt[c];
              }
            }
          }
        } else {
          const float pr = 1.f / (1.f + expf(-(dot + p.head_b)));
          if (live) {
            if (p.prob_f32) p.prob_f32[pix] = pr;
            if (p.prob_f16) p.prob_f16[pix] = __float2half_rn(pr);
            if (p.mask_u8) p.mask_u8[pix] = pr > p.thr ? 255 : 0;
            if (p.tile_dst && pr > p.thr) {            // glue: OR into the pre-zeroed line plane
              const sd_tile_dst td = p.tile_dst[n];
              if (x < td.width) td.d_dst[(int64_t)y * td.pitch + x] = 255;
            }
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (EPI != EPI_HEAD && et == 0) tma_store_wait_all();    // all output bytes written before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// conv_umma2_kernel: the store-epilogue conv for Cout % 256 == 0 on a PAIR of SMs (thread-block cluster of 2,
// tcgen05.mma.cta_group::2): one 256 x 256 tile per pair.  Each CTA stages its own 128-pixel A tile and HALF of
// the 256-row B k-block (the tensor cores of both SMs read both halves), so a pipeline stage is 32 KB instead of
// 48 KB and six stages fit instead of four: the same shared memory now covers 50 % more TMA latency, which is
// what starves the single-CTA kernel (its MMA warp waits on `full` barriers 25-30 % of the time).
//   * both CTAs' TMA loads signal the LEADER's (even CTA's) full barrier (.cta_group::2 form, peer bit cleared);
//   * only the leader issues MMAs; tcgen05.commit multicasts the slot-free / accumulator-full signals to both;
//   * each CTA's epilogue drains its own 128 accumulator lanes and arrives on the leader's TMEM-empty barrier.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the even CTA of the pair

// BN = 256 (Cout % 256 == 0 layers) or 128 (the level-2 layers, ONE M tile per CTA: per MMA a CTA reads 4 KB of A and its
// 2-KB half of B instead of the 4 + 4 KB of the single-CTA MT = 2 form, and a k-block fills 24 KB instead of 48 KB:
// those layers are bound by shared-memory bandwidth).
template <int BN> struct Conv2CfgT {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = (BN / 2) * 128;        // this CTA's half of the BN-row B k-block
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = 128 * 64 * 2 + 4096;
  static constexpr int kFit = (kMaxSmem - 1024 - kMiscBytes - kOutBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;   // 6 (BN = 256) / 8 (BN = 128)
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 + kMiscBytes;
  static constexpr uint32_t kIdesc = kIdescBase | ((uint32_t)(BN >> 3) << 17) | ((256u >> 4) << 24);   // M = 256, N = BN
  static_assert(BN == 256 || BN == 128, "2-SM conv: N = 256 or 128");
};
using Conv2Cfg = Conv2CfgT<256>;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {      // arrives on `bar` in both CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the even CTA's barrier from either CTA
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1) conv_umma2_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = Conv2CfgT<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t out_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  uint8_t* misc = smem_al + Cfg::kStages * Cfg::kStageBytes + Cfg::kOutBytes;
  const uint32_t bar_base = out_base + Cfg::kOutBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(misc + 248);
  float* s_bias = reinterpret_cast<float*>(misc + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();              // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    if (p.c1_blocks) tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < p.n_phases; ++i) tma_prefetch_desc(&p.tmOut[i]);
    if (p.pool) tma_prefetch_desc(&p.tmPool);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

#ifdef SD_CTA2_DEBUG
  if (threadIdx.x == 0 && pair == 0 && p.err_flag) {
    p.err_flag[2 + 4 * rank] = (int)smem_u32(smem_raw);
    p.err_flag[3 + 4 * rank] = (int)full_bar(0);
    p.err_flag[4 + 4 * rank] = (int)(full_bar(0) & kPeerBitMask);
    uint32_t mapped;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(mapped) : "r"(full_bar(0)), "r"(0u));
    p.err_flag[5 + 4 * rank] = (int)mapped;
  }
#endif
  const int m_groups = (p.m_tiles + 1) / 2;             // an M group = the two tiles of a pair
  const int n_work = m_groups * p.n_tiles * p.n_phases;
  const int kb_per_tap = p.c0_blocks + p.c1_blocks;
  const int num_kb = p.n_taps * kb_per_tap;
  const FastDiv fd_tx((uint32_t)p.tiles_x), fd_ty((uint32_t)p.tiles_y), fd_nt((uint32_t)p.n_tiles), fd_mg((uint32_t)m_groups);
  auto origin = [&](int mt, int& x0, int& y0, int& n0) {
    uint32_t rest, tx, tn, ty;
    fd_tx.divmod((uint32_t)mt, rest, tx); fd_ty.divmod(rest, tn, ty);
    x0 = (int)tx * p.box_w; y0 = (int)ty * p.box_h; n0 = (int)tn * p.box_n;
  };

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int w = pair; w < n_work; w += n_pairs) {
        uint32_t nt_u, rest_u, mg_u, ph_u;
        fd_nt.divmod((uint32_t)w, rest_u, nt_u); fd_mg.divmod(rest_u, ph_u, mg_u);
        const int nt = (int)nt_u, mg = (int)mg_u, ph = (int)ph_u;
        int x0, y0, n0;
        origin(mg * 2 + (int)rank, x0, y0, n0);
        const int brow = ph * p.cout + nt * BN + (int)rank * (BN / 2);
        for (int tap = 0; tap < p.n_taps; ++tap) {
          const int dx = p.dx[ph][tap], dy = p.dy[ph][tap];
          for (int cb = 0; cb < kb_per_tap; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1 + 10 * (int)rank);
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);    // both CTAs' bytes land on the leader's barrier
            const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
            const CUtensorMap* tm = cb < p.c0_blocks ? &p.tmA0 : &p.tmA1;
            const int c = (cb < p.c0_blocks ? cb : cb - p.c0_blocks) * 64;
            tma_load_4d_2sm(a_dst, tm, full_bar(stage), c, x0 + dx, y0 + dy, n0);
            tma_load_2d_2sm(a_dst + Cfg::kABytes, &p.tmB, full_bar(stage), (tap * kb_per_tap + cb) * 64, brow);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only) =======================
    if (rank == 0 && elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int w = pair; w < n_work; w += n_pairs) {
        mbar_wait(tempty_bar(as), aphase ^ 1u, p.err_flag, 2 + 10 * (int)rank);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 3 + 10 * (int)rank);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = umma_desc_sw128(a_addr);
          const uint64_t bdesc = umma_desc_sw128(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, Cfg::kIdesc, (uint32_t)((kb | k) != 0));
          umma_commit_2sm(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue (warps 2..5, both CTAs: own 128 accumulator lanes) =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    int as = 0; uint32_t aphase = 0;
    int cur_nt = -1;
    for (int w = pair; w < n_work; w += n_pairs) {
      uint32_t nt_u, rest_u, mg_u, ph_u;
      fd_nt.divmod((uint32_t)w, rest_u, nt_u); fd_mg.divmod(rest_u, ph_u, mg_u);
      const int nt = (int)nt_u, mg = (int)mg_u, ph = (int)ph_u;
      if (nt != cur_nt) {
        epi_bar();
        for (int i = et; i < BN; i += 128) s_bias[i] = __ldg(p.bias + nt * BN + i);
        cur_nt = nt;
        epi_bar();
      }
      mbar_wait(tfull_bar(as), aphase, p.err_flag, 4 + 10 * (int)rank);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      int x0, y0, n0;
      origin(mg * 2 + (int)rank, x0, y0, n0);
#pragma unroll 1
      for (int hb = 0; hb < BN / 64; ++hb) {
        if (et == 0) tma_store_wait_read();
        epi_bar();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = hb * 2 + cc;
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          const uint32_t row_base = out_base + (uint32_t)row * 128u;
          if (p.relu) epi_pack32<true>(v, s_bias + c * 32, row_base, row, cc);
          else epi_pack32<false>(v, s_bias + c * 32, row_base, row, cc);
        }
        if (hb == BN / 64 - 1) {
          tc_fence_before();
          mbar_arrive_leader(tempty_bar(as));       // 128 + 128 arrivals free the accumulator stage of the pair
        }
        fence_async_smem();
        epi_bar();
        if (et == 0) {
          tma_store_4d(&p.tmOut[ph], out_base, nt * BN + hb * 64, x0, y0, n0);
          tma_store_commit();
        }
        if (p.pool) {
          const int pw = p.box_w >> 1, phh = p.box_h >> 1;
          const int pw_sh = 31 - __clz(pw), ph_sh = 31 - __clz(phh);
          const uint32_t pool_base = out_base + 16384u;
#pragma unroll
          for (int task = et; task < 256; task += 128) {
            const int pp = task >> 3, ch = task & 7;
            const int px = pp & (pw - 1); const int r2 = pp >> pw_sh;           // box dims are powers of two
            const int py = r2 & (phh - 1), pn = r2 >> ph_sh;
            const int m00 = (pn * p.box_h + 2 * py) * p.box_w + 2 * px, m10 = m00 + p.box_w;
            const uint4 v = hmax2_v4(hmax2_v4(ld_shared_v4(sw128(out_base, m00, ch)), ld_shared_v4(sw128(out_base, m00 + 1, ch))),
                                     hmax2_v4(ld_shared_v4(sw128(out_base, m10, ch)), ld_shared_v4(sw128(out_base, m10 + 1, ch))));
            st_shared_v4(sw128(pool_base, pp, ch), v);
          }
          fence_async_smem();
          epi_bar();
          if (et == 0) {
            tma_store_4d(&p.tmPool, pool_base, nt * BN + hb * 64, x0 >> 1, y0 >> 1, n0);
            tma_store_commit();
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (et == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // nobody exits while the peer may still signal its barriers / read its smem
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Level-1 halo rows: one A stage of the band kernel is a row segment of 130 pixels x 64 channels (the 128 output
// pixels plus one halo pixel on each side), fetched by one 4-D TMA box; the three dx taps read it through UMMA
// descriptors whose start address is shifted by dx rows (128 B) inside the swizzled buffer.  Measured on B200: the
// 128-B swizzle is applied on absolute shared-memory address bits, so a start that is not 1024-B aligned needs no
// correction (setting base_offset = (addr >> 7) & 7 gives wrong results).
// ---------------------------------------------------------------------------------------------
constexpr int kRowStageBytes = 17 * 1024;     // 130 rows x 128 B = 16640, padded to keep 1024-B alignment
constexpr int kRowBoxBytes = 130 * 128;

// ---------------------------------------------------------------------------------------------
// conv_band_kernel<CB, EPI>: the level-1 layers with 3x3 taps, Cin = CB*64 -> Cout = 64, W % 128 == 0, dy-stacked.
// An N = 64 MMA reads 4 KB of A + 2 KB of B from shared memory for 32 tensor cycles (192 B/clk against the
// 128 B/clk the shared memory delivers) and each input row is fetched three times.  Here ONE input halo row
// (130 px x 64 ch, fetched once) feeds the three output rows it contributes to in a single N = 192 MMA per
// (dx, k-step): the weights of ky = 2, 1, 0 are stacked along N and the accumulators of consecutive output rows
// sit in consecutive 64-column TMEM slots, a ring of 8 slots = all 512 columns.  10 KB per 96 cycles = 107 B/clk.
//   * every slot only ever receives the three valid contributions of its output row, so all MMAs accumulate;
//     the epilogue zeroes a slot (tcgen05.st) after draining it;
//   * a window that wraps around the ring (2 of 8 positions) is issued as two MMAs;
//   * work item = (image, 128-px segment, band of kBandRows output rows); band edges use N = 64 / 128 windows,
//     so no MAC is wasted, only the two halo input rows are fetched twice.
// ---------------------------------------------------------------------------------------------
constexpr int kBandRows = 32;

template <int CB, int EPI, bool PSI = false> struct BandCfg {
  static constexpr int kWBytes = 9 * CB * 8192;        // [dx][cb][ky = 2, 1, 0][64 cout] rows of 128 B
  static constexpr bool kCanPool = (EPI == EPI_STORE) && CB == 1;      // fused MaxPool2x2 (Conv1.3): two row buffers + pooled row
  static constexpr int kOutBytes = (EPI == EPI_STORE ? 16384 : 0) + (kCanPool ? 16384 + 8192 : 0);
  static constexpr int kFit = (kMaxSmem - 1024 - kMiscBytes - kWBytes - kOutBytes) / kRowStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kSmemBytes = kStages * kRowStageBytes + kWBytes + kOutBytes + 1024 + kMiscBytes;
  static constexpr int kEpiWarps = (EPI == EPI_STORE) ? SD_EPI_WARPS : 4;   // as in ConvCfg; the head reduces over all 64 channels per thread
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  // PSI: four more warps scale the staged rows of source 0 (the skip tensor) by the attention gate's psi plane before the
  // MMAs read them: the gate then writes 4 bytes per pixel instead of the scaled copy of the skip tensor (Att2: 1.6 GB per
  // 256 tiles), and x * psi never exists in HBM.  Same arithmetic as the gate's own scaling (fp32 product, one rounding).
  static constexpr int kScaleThreads = PSI ? 128 : 0;
  static constexpr int kThreads = 64 + kEpiThreads + kScaleThreads;
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(!PSI || (CB == 2 && EPI == EPI_STORE), "psi scaling is built for the concat consumer (Up_conv2.0)");
};

__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int CB, int EPI, bool PSI>
__global__ void __launch_bounds__((BandCfg<CB, EPI, PSI>::kThreads), 1) conv_band_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = BandCfg<CB, EPI, PSI>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_base = smem_base + Cfg::kStages * kRowStageBytes;                 // resident weights
  const uint32_t out_base = w_base + Cfg::kWBytes;
  uint8_t* misc = smem_al + Cfg::kStages * kRowStageBytes + Cfg::kWBytes + Cfg::kOutBytes;
  const uint32_t bar_base = out_base + Cfg::kOutBytes;
  // misc: [0,384) barriers: full[S], empty[S], tfull[8], tempty[8], weights; [384] TMEM slot; [512] bias; [768] head vector
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (16 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (24 + s); };
  const uint32_t w_bar = bar_base + 8u * 32;
  auto scaled_bar = [&](int s) { return bar_base + 8u * (33 + s); };     // PSI: stage s has been scaled (or needs no scaling)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(misc + 384);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    if (CB > 1) tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmB);
    if (EPI == EPI_STORE) tma_prefetch_desc(&p.tmOut[0]);
    if (Cfg::kCanPool && p.pool) tma_prefetch_desc(&p.tmPool);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), Cfg::kEpiThreads); }
    if (PSI) for (int s = 0; s < Cfg::kStages; ++s) mbar_init(scaled_bar(s), 128);
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 2 && warp < 6) {                         // all accumulator slots start at zero
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 16; ++c) tmem_st32_zero(t0 + c * 32);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int segs = p.W / 128;
  const int bands = (p.H + kBandRows - 1) / kBandRows;
  const int n_work = p.B * bands * segs;
  // per CTA, the output rows of its work items are numbered consecutively: g -> slot g & 7, use (g >> 3)

  if (warp == 0) {
    if (elect_one()) {
      // one-time: weights -> smem, block (dx, cb, kyr) <- tap (ky = 2 - kyr, kx = dx), 64 rows x 128 B each
      mbar_expect_tx(w_bar, Cfg::kWBytes);
      for (int dx = 0; dx < 3; ++dx)
        for (int cb = 0; cb < CB; ++cb)
          for (int kyr = 0; kyr < 3; ++kyr)
            tma_load_2d(w_base + ((dx * CB + cb) * 3 + kyr) * 8192, &p.tmB, w_bar, (((2 - kyr) * 3 + dx) * CB + cb) * 64, 0);
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int sx = w % segs; int rest = w / segs;
        const int b = rest % bands; const int n = rest / bands;
        const int y0 = b * kBandRows, y1 = min(y0 + kBandRows, p.H);
        for (int i = max(y0 - 1, 0); i <= min(y1, p.H - 1); ++i) {
          for (int cb = 0; cb < CB; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1);
            mbar_expect_tx(full_bar(stage), kRowBoxBytes);
            tma_load_4d(smem_base + stage * kRowStageBytes, cb == 0 ? &p.tmA0 : &p.tmA1, full_bar(stage), 0, sx * 128 - 1, i, n);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(w_bar, 0, p.err_flag, 5);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      int gbase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int b = (w / segs) % bands;
        const int y0 = b * kBandRows, y1 = min(y0 + kBandRows, p.H);
        for (int i = max(y0 - 1, 0); i <= min(y1, p.H - 1); ++i) {
          const int r_lo = max(i - 1, y0), r_hi = min(i + 1, y1 - 1);
          const int nr = r_hi - r_lo + 1;               // output rows fed by this input row (1..3)
          const int kyr_lo = r_lo - i + 1;              // first weight block of the window
          const int g_lo = gbase + (r_lo - y0);
          const int s_lo = g_lo & 7;
          // rows that get their first contribution from this input row (row i+1; at the image top also row 0):
          // their slots must have been drained and zeroed by the epilogue
          if (i == 0 && y0 == 0) {
            mbar_wait(tempty_bar(gbase & 7), (uint32_t)(((gbase >> 3) & 1) ^ 1), p.err_flag, 2);
            tc_fence_after();
          }
          if (r_hi == i + 1) {
            const int g_hi = gbase + (r_hi - y0);
            mbar_wait(tempty_bar(g_hi & 7), (uint32_t)(((g_hi >> 3) & 1) ^ 1), p.err_flag, 2);
            tc_fence_after();
          }
          const int n1 = min(nr, 8 - s_lo), n2 = nr - n1;   // window split at the end of the ring
          const uint32_t idesc1 = kIdescBase | ((uint32_t)(n1 * 8) << 17) | ((128u >> 4) << 24);
          const uint32_t idesc2 = kIdescBase | ((uint32_t)(n2 * 8) << 17) | ((128u >> 4) << 24);
          const uint32_t d1 = tmem_base + (uint32_t)(s_lo * 64);
          for (int cb = 0; cb < CB; ++cb) {
            mbar_wait(PSI ? scaled_bar(stage) : full_bar(stage), phase, p.err_flag, 3);
            tc_fence_after();
            const uint32_t a_addr = smem_base + stage * kRowStageBytes;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const uint64_t adesc = umma_desc_sw128(a_addr + dx * 128);
              const uint32_t wb = w_base + ((dx * CB + cb) * 3 + kyr_lo) * 8192;
              const uint64_t bdesc1 = umma_desc_sw128(wb);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d1, adesc + 2u * k, bdesc1 + 2u * k, idesc1, 1u);
              if (n2) {
                const uint64_t bdesc2 = umma_desc_sw128(wb + n1 * 8192);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + 2u * k, bdesc2 + 2u * k, idesc2, 1u);
              }
            }
            umma_commit(empty_bar(stage));
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
          // rows whose last contribution this was
          if (i - 1 >= y0) umma_commit(tfull_bar((gbase + (i - 1 - y0)) & 7));
          if (i == p.H - 1 && i < y1) umma_commit(tfull_bar((gbase + (i - y0)) & 7));
        }
        gbase += y1 - y0;
      }
    }
    __syncwarp();
  } else if (PSI && warp >= 2 + Cfg::kEpiWarps) {
    // ======================= psi scalers (4 warps): thread = staged pixel row =======================
    // Walks the producer's stage sequence.  Source-0 stages (the skip tensor): wait for the TMA box, multiply the 64
    // channels of pixel row r by psi of that pixel (fp32 product, one rounding: what the gate's own scaling did), make the
    // generic-proxy writes visible to the tensor pipe, arrive.  Source-1 stages: arrive as soon as the box has landed.
    const int t = threadIdx.x - 64 - Cfg::kEpiThreads;  // 0..127
    int stage = 0; uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int sx = w % segs; int rest = w / segs;
      const int b = rest % bands; const int n = rest / bands;
      const int y0 = b * kBandRows, y1 = min(y0 + kBandRows, p.H);
      for (int i = max(y0 - 1, 0); i <= min(y1, p.H - 1); ++i) {
        // psi of this thread's halo pixel(s): columns sx*128 - 1 + r, r = t (and t + 128 for the two last halo rows); columns
        // outside the image hold zero-filled data, any finite factor does
        const float* prow = p.psi_in + ((int64_t)n * p.H + i) * p.W;
        const float sc0 = __ldg(prow + min(max(sx * 128 - 1 + t, 0), p.W - 1));
        const float sc1 = t < 2 ? __ldg(prow + min(sx * 128 + 127 + t, p.W - 1)) : 0.f;
        for (int cb = 0; cb < CB; ++cb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 7);
          if (cb == 0) {
            const uint32_t a_addr = smem_base + stage * kRowStageBytes;
#pragma unroll 1
            for (int rr = 0; rr < (t < 2 ? 2 : 1); ++rr) {
              const int r = t + rr * 128;
              const float sc = rr ? sc1 : sc0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t addr = sw128(a_addr, r, j);
                uint4 v = ld_shared_v4(addr);
                act2_t* h = reinterpret_cast<act2_t*>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float2 f = act22float2(h[k]);
                  h[k] = floats2act2(f.x * sc, f.y * sc);
                }
                st_shared_v4(addr, v);
              }
            }
            fence_async_smem();                         // generic-proxy writes -> visible to the tensor pipe
          }
          mbar_arrive(scaled_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int eh = (warp - 2) >> 2;                     // store epilogue with 8 warps: which 32 of the row's 64 channels
    int g = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int sx = w % segs; int rest = w / segs;
      const int b = rest % bands; const int n = rest / bands;
      const int y0 = b * kBandRows, y1 = min(y0 + kBandRows, p.H);
      uint8_t* line_dst = nullptr; int line_pitch = 0;
      if (EPI == EPI_HEAD && p.tile_dst) {              // this tile's columns in the packed line planes
        const sd_tile_dst td = p.tile_dst[n];
        if (sx * 128 + row < td.width) { line_dst = td.d_dst + sx * 128 + row; line_pitch = td.pitch; }
      }
      for (int y = y0; y < y1; ++y, ++g) {
        const int slot = g & 7;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 64);
        mbar_wait(tfull_bar(slot), (uint32_t)((g >> 3) & 1), p.err_flag, 4);
        tc_fence_after();
        if constexpr (EPI == EPI_STORE) {
          const bool pooling = Cfg::kCanPool && p.pool;
          const uint32_t row_buf = out_base + ((pooling && (y & 1)) ? 16384u : 0u);     // rows alternate buffers when pooling
          if (et == 0) tma_store_wait_read();
          epi_bar_n<Cfg::kEpiThreads>();
#pragma unroll
          for (int c = (Cfg::kEpiWarps == 8 ? eh : 0); c < (Cfg::kEpiWarps == 8 ? eh + 1 : 2); ++c) {
            float v[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_st32_zero(taddr + c * 32);
            const uint32_t rbase = row_buf + (uint32_t)row * 128u;
            if (c == 0) epi_pack32_const_relu(v, p.bias_c, rbase, row, 0);           // two copies: the bias offsets stay compile-time
            else epi_pack32_const_relu(v, p.bias_c + 32, rbase, row, 1);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(tempty_bar(slot));
          fence_async_smem();
          epi_bar_n<Cfg::kEpiThreads>();
          if (et == 0) { tma_store_4d(&p.tmOut[0], row_buf, 0, sx * 128, y, n); tma_store_commit(); }
          if (pooling && (y & 1)) {
            // fused MaxPool2x2 over rows y-1 (buffer 0) and y (buffer 1): 64 pooled pixels x 8 chunks
            const uint32_t pool_base = out_base + 32768u;
#pragma unroll
            for (int task = et; task < 512; task += Cfg::kEpiThreads) {
              const int pp = task >> 3, ch = task & 7;
              const uint4 v = hmax2_v4(hmax2_v4(ld_shared_v4(sw128(out_base, 2 * pp, ch)), ld_shared_v4(sw128(out_base, 2 * pp + 1, ch))),
                                       hmax2_v4(ld_shared_v4(sw128(out_base + 16384u, 2 * pp, ch)),
                                                ld_shared_v4(sw128(out_base + 16384u, 2 * pp + 1, ch))));
              st_shared_v4(sw128(pool_base, pp, ch), v);
            }
            fence_async_smem();
            epi_bar_n<Cfg::kEpiThreads>();
            if (et == 0) { tma_store_4d(&p.tmPool, pool_base, 0, sx * 64, y >> 1, n); tma_store_commit(); }
          }
        } else {
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float v[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_st32_zero(taddr + c * 32);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = fmaxf(v[j] + p.bias_c[c * 32 + j], 0.f);   // stays fp32 into the 1x1 conv (the oracle is fp32)
              dot = fmaf(a, p.vec_c[c * 32 + j], dot);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(tempty_bar(slot));
          const float pr = 1.f / (1.f + expf(-(dot + p.head_b)));
          const int64_t pix = ((int64_t)n * p.H + y) * p.W + sx * 128 + row;
          if (p.prob_f32) p.prob_f32[pix] = pr;
          if (p.prob_f16) p.prob_f16[pix] = __float2half_rn(pr);
          if (p.mask_u8) p.mask_u8[pix] = pr > p.thr ? 255 : 0;
          // glue fused into the head (helper/split.py:109-119): planes are pre-zeroed, every covering tile ORs its
          // foreground pixels in; equal bytes from two tiles of an overlap are a benign race
          if (line_dst && pr > p.thr) line_dst[(int64_t)y * line_pitch] = 255;
        }
      }
    }
    if (EPI == EPI_STORE && et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// conv_up4_kernel: nearest-x2 upsample + conv3x3 with Cout = 64 (Up2), all four sub-pixel phases in one work item.
// Run phase by phase (generic kernel) every MMA has N = 64 and costs the 71-cycle operand-read floor: 45 % of the
// tensor peak at best.  Here a work item is one M tile of 128 LOW-res pixels; its four phase accumulators sit side
// by side in TMEM (4 x 64 columns, order (py,px) = (0,1) (0,0) (1,0) (1,1)) and each of the 9 low-res taps is
// multiplied ONCE against the pre-summed weights of every phase that uses it: N = 256 for the centre tap, 128 for
// the edge taps (one of them wraps the slot ring and is issued as 2 x 64), 64 for the corners -> 80 instead of
// 128 MMAs per tile and 18 instead of 32 A boxes.  Store epilogue as in conv_umma_kernel (one strided output map
// per phase).
// ---------------------------------------------------------------------------------------------
struct Up4Cfg {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = 4 * 8192;              // up to four 64-row weight blocks
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = 16384;
  static constexpr int kStages = (kMaxSmem - 1024 - kMiscBytes - kOutBytes) / kStageBytes;   // 4
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 + kMiscBytes;
  static constexpr int kEpiWarps = SD_EPI_WARPS;        // as in ConvCfg: two epilogue warps per TMEM lane quarter
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  static constexpr int kThreads = 64 + kEpiThreads;
  static_assert(kStages >= 3, "pipeline too shallow");
};

// tap order: the centre tap first (its N = 256 MMA initialises all four accumulators).  Per tap: low-res offset,
// first TMEM slot and number of phases that use it (slots in ring order: 0 = (py,px) (0,1), 1 = (0,0), 2 = (1,0),
// 3 = (1,1); the (0,+1) tap covers slots 3 and 0 and wraps).  The weights are packed on the host in exactly this
// order ([tap][cb][phase in slot order][co] rows of 64 halves), so one TMA fetches a stage's whole B operand.
struct Up4Tap { int dy, dx, first, count, prefix; };
__host__ __device__ constexpr Up4Tap up4_tap(int t) {
  constexpr Up4Tap tab[9] = {{0, 0, 0, 4, 0},  {-1, -1, 1, 1, 4}, {-1, 0, 0, 2, 5}, {-1, 1, 0, 1, 7}, {0, -1, 1, 2, 8},
                             {0, 1, 3, 2, 10}, {1, -1, 2, 1, 12}, {1, 0, 2, 2, 13}, {1, 1, 3, 1, 15}};
  return tab[t];
}
__host__ __device__ constexpr int up4_slot_phase(int slot) { return slot == 0 ? 1 : (slot == 1 ? 0 : (slot == 2 ? 2 : 3)); }   // ph = py * 2 + px

__global__ void __launch_bounds__(Up4Cfg::kThreads, 1) conv_up4_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = Up4Cfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t out_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  uint8_t* misc = smem_al + Cfg::kStages * Cfg::kStageBytes + Cfg::kOutBytes;
  const uint32_t bar_base = out_base + Cfg::kOutBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(misc + 248);
  float* s_bias = reinterpret_cast<float*>(misc + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB); tma_prefetch_desc(&p.tmA1); tma_prefetch_desc(&p.tmPool);   // weight maps with 64 / 128 / 256-row boxes
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmOut[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), Cfg::kEpiThreads); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    if (t < 64) s_bias[t] = p.bias[t];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_work = p.m_tiles;
  const int cbs = p.c0_blocks;                           // 64-channel blocks of the (single) source
  const FastDiv fd_tx((uint32_t)p.tiles_x), fd_ty((uint32_t)p.tiles_y);
  auto origin = [&](int mt, int& x0, int& y0, int& n0) {
    uint32_t rest, tx, tn, ty;
    fd_tx.divmod((uint32_t)mt, rest, tx); fd_ty.divmod(rest, tn, ty);
    x0 = (int)tx * p.box_w; y0 = (int)ty * p.box_h; n0 = (int)tn * p.box_n;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        int x0, y0, n0;
        origin(w, x0, y0, n0);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const Up4Tap tp = up4_tap(t);
          // B operand of the stage: count x 64 rows, one TMA through the map whose box has that many rows
          const CUtensorMap* tmb = tp.count == 4 ? &p.tmPool : (tp.count == 2 ? &p.tmA1 : &p.tmB);
          for (int cb = 0; cb < cbs; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1);
            mbar_expect_tx(full_bar(stage), Cfg::kABytes + tp.count * 8192);
            const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
            tma_load_4d(a_dst, &p.tmA0, full_bar(stage), cb * 64, x0 + tp.dx, y0 + tp.dy, n0);
            tma_load_2d(a_dst + Cfg::kABytes, tmb, full_bar(stage), 0, (tp.prefix * cbs + cb * tp.count) * 64);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u, p.err_flag, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const Up4Tap tp = up4_tap(t);
          const int first = tp.first, count = tp.count;
          const int n1 = count < 4 - first ? count : 4 - first, n2 = count - n1;       // split where the slot ring wraps
          const uint32_t idesc1 = kIdescBase | ((uint32_t)(n1 * 8) << 17) | ((128u >> 4) << 24);
          const uint32_t idesc2 = kIdescBase | ((uint32_t)(n2 * 8) << 17) | ((128u >> 4) << 24);
          for (int cb = 0; cb < cbs; ++cb) {
            mbar_wait(full_bar(stage), phase, p.err_flag, 3);
            tc_fence_after();
            const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
            const uint64_t adesc = umma_desc_sw128(a_addr);
            const uint64_t bdesc1 = umma_desc_sw128(a_addr + Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem + (uint32_t)(first * 64), adesc + 2u * k, bdesc1 + 2u * k, idesc1, (uint32_t)((t | cb | k) != 0));
            if (n2) {
              const uint64_t bdesc2 = umma_desc_sw128(a_addr + Cfg::kABytes + n1 * 8192);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2u * k, bdesc2 + 2u * k, idesc2, 1u);
            }
            umma_commit(empty_bar(stage));
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int eh = (warp - 2) >> 2;               // 8 epilogue warps: which 32 of the 64 channels of a phase (ConvCfg::kEpiWarps)
    int as = 0; uint32_t aphase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int x0, y0, n0;
      origin(w, x0, y0, n0);
      mbar_wait(tfull_bar(as), aphase, p.err_flag, 4);
      tc_fence_after();
#pragma unroll 1
      for (int slot = 0; slot < 4; ++slot) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256 + slot * 64);
        if (et == 0) tma_store_wait_read();
        epi_bar_n<Cfg::kEpiThreads>();
#pragma unroll
        for (int c = (Cfg::kEpiWarps == 8 ? eh : 0); c < (Cfg::kEpiWarps == 8 ? eh + 1 : 2); ++c) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          const uint32_t rbase = out_base + (uint32_t)row * 128u;
          epi_pack32<true>(v, s_bias + c * 32, rbase, row, c);
        }
        if (slot == 3) {
          tc_fence_before();
          mbar_arrive(tempty_bar(as));
        }
        fence_async_smem();
        epi_bar_n<Cfg::kEpiThreads>();
        if (et == 0) { tma_store_4d(&p.tmOut[up4_slot_phase(slot)], out_base, 0, x0, y0, n0); tma_store_commit(); }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// conv_first_umma_kernel: Conv1.0, 3(+5 zero) -> 64 channels, 3x3, pad 1, bias + ReLU on the tensor pipe.
// The NHWC8 input makes one (pixel, tap) = 8 halves = 16 B = exactly one row of a UMMA core matrix, so the
// im2col A tile [128 px x K = 9 taps x 8 ch (+8 zero) = 80] is written by 128 producer threads straight into
// the canonical no-swizzle K-major layout: core matrix (row group g, tap t) at g * 1280 + t * 128, row r % 8
// at + 16 B.  Image borders are zero-filled by predication.  Weights [64 x 80] sit in shared memory in the
// same layout for the whole kernel.  M tile = one 128-pixel row segment; 5 MMAs (K = 16 each) per tile, so
// the kernel is bound by the 16 KB output store per tile, not by the tensor pipe.
//   warp 0: MMA issuer | warps 1-4: im2col producers | warps 5-8: epilogue (TMEM -> bias/ReLU -> TMA store)
// ---------------------------------------------------------------------------------------------
#ifndef SD_C1_CTAS
#define SD_C1_CTAS 2
#endif
constexpr int kC1CtasPerSm = SD_C1_CTAS;               // 2 CTAs x 4 stages (108 KB each); 3 CTAs x 2 stages measured slower (0.53 vs 0.48 ms per 256 tiles)
constexpr int kC1Stages = kC1CtasPerSm == 3 ? 2 : 4;
constexpr int kC1K = 80;
constexpr int kC1GroupBytes = (kC1K / 8) * 128;        // 1280: one 8-row group across K
constexpr int kC1ABytes = 16 * kC1GroupBytes;          // 20480
constexpr int kC1BBytes = 8 * kC1GroupBytes;           // 10240 (64 output channels)
constexpr int kC1Threads = 288;
constexpr int kC1SmemBytes = kC1Stages * kC1ABytes + kC1BBytes + 16384 + 1024 + 1024;

struct ConvFirstParams {
  CUtensorMap tmOut;
  const uint4* in;             // NHWC8 fp16, 16 B per pixel
  const uint4* w;              // 10240 B, canonical layout (host-packed)
  const float* bias;           // [64]
  int B, H, W;
  int* err_flag;
};

// K-major operand without swizzle: LBO = distance between core matrices adjacent in K, SBO = between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_none(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(kC1Threads, kC1CtasPerSm) conv_first_umma_kernel(const __grid_constant__ ConvFirstParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_base = smem_base + kC1Stages * kC1ABytes;
  const uint32_t out_base = w_base + kC1BBytes;                    // 1024-aligned (20480 * 4 + 10240 = 92160 = 90 KB)
  uint8_t* misc = smem_al + kC1Stages * kC1ABytes + kC1BBytes + 16384;
  const uint32_t bar_base = out_base + 16384;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kC1Stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kC1Stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kC1Stages + 2 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(misc + 248);
  float* s_bias = reinterpret_cast<float*>(misc + 256);
  constexpr uint32_t kIdesc = kIdescBase | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmOut);
    for (int s = 0; s < kC1Stages; ++s) { mbar_init(full_bar(s), 128); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights -> smem, zero the K-padding core matrix (tap slot 9) of every stage, bias
  for (int i = threadIdx.x; i < kC1BBytes / 16; i += kC1Threads)
    reinterpret_cast<uint4*>(smem_al + kC1Stages * kC1ABytes)[i] = __ldg(p.w + i);
  for (int i = threadIdx.x; i < kC1Stages * 128; i += kC1Threads) {
    const int st = i >> 7, r = i & 127;
    *reinterpret_cast<uint4*>(smem_al + st * kC1ABytes + (r >> 3) * kC1GroupBytes + 9 * 128 + (r & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int segs = p.W / 128;
  const int n_work = p.B * p.H * segs;
  const FastDiv fd_segs((uint32_t)segs), fd_h((uint32_t)p.H);     // a work item is one 128-pixel row segment: its index math must be cheap

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      const uint64_t bdesc = umma_desc_none(w_base, 128, kC1GroupBytes);
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u, p.err_flag, 2);
        mbar_wait(full_bar(stage), phase, p.err_flag, 3);
        tc_fence_after();
        const uint64_t adesc = umma_desc_none(smem_base + stage * kC1ABytes, 128, kC1GroupBytes);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 64);
#pragma unroll
        for (int k = 0; k < kC1K / 16; ++k)     // K = 16 halves = 2 core matrices = 256 B -> +16 in the (addr >> 4) field
          umma_f16(d_tmem, adesc + 16u * k, bdesc + 16u * k, kIdesc, (uint32_t)(k != 0));
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(as));
        if (++stage == kC1Stages) { stage = 0; phase ^= 1u; }
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp <= 4) {
    // ======================= im2col producers: thread = pixel =======================
    const int r = threadIdx.x - 32;
    int stage = 0; uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      uint32_t rest_u, sx_u, n_u, y_u;
      fd_segs.divmod((uint32_t)w, rest_u, sx_u); fd_h.divmod(rest_u, n_u, y_u);
      const int sx = (int)sx_u, y = (int)y_u, n = (int)n_u;
      const int x = sx * 128 + r;
      uint4 v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        v[t] = make_uint4(0u, 0u, 0u, 0u);
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) v[t] = __ldg(p.in + ((int64_t)n * p.H + yy) * p.W + xx);
      }
      mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1);
      uint8_t* dst = smem_al + stage * kC1ABytes + (r >> 3) * kC1GroupBytes + (r & 7) * 16;
#pragma unroll
      for (int t = 0; t < 9; ++t) *reinterpret_cast<uint4*>(dst + t * 128) = v[t];
      fence_async_smem();                         // generic-proxy writes -> visible to the tensor pipe
      mbar_arrive(full_bar(stage));
      if (++stage == kC1Stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ======================= epilogue (warps 5..8) =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 160;
    int as = 0; uint32_t aphase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      uint32_t rest_u, sx_u, n_u, y_u;
      fd_segs.divmod((uint32_t)w, rest_u, sx_u); fd_h.divmod(rest_u, n_u, y_u);
      const int sx = (int)sx_u, y = (int)y_u, n = (int)n_u;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 64);
      mbar_wait(tfull_bar(as), aphase, p.err_flag, 4);
      tc_fence_after();
      if (et == 0) tma_store_wait_read();
      epi_bar();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        const uint32_t rbase = out_base + (uint32_t)row * 128u;
        epi_pack32<true>(v, s_bias + c * 32, rbase, row, c);
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
      fence_async_smem();
      epi_bar();
      if (et == 0) { tma_store_4d(&p.tmOut, out_base, 0, sx * 128, y, n); tma_store_commit(); }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

}  // namespace sd
