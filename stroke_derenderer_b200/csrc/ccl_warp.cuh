// K7 + K8, second generation: connected-component labelling with OpenCV's numbering and the cv2
// `stats` rows in the same pass.  Replaces the strip-per-CTA kernels of round 1 (320 k thread-instructions per
// 128 x 128 strip, IPC ~1 behind five __syncthreads) with a WARP per strip and no CTA barrier at all:
//
//   1. ccl_warp_label_kernel (warp = strip of 128 x 128 px = 64 x 64 blocks of 2 x 2 px).  The mask streams in
//      with coalesced 16-byte loads; every load becomes 8 "even column" + 8 "odd column" bits (dp4a), so pixel row r
//      of the strip is two 64-bit words Xe / Xo whose bit k is the left / right pixel of block k.  Lane L owns block
//      rows 2L and 2L+1: occupancy, horizontal links, run starts and the contacts with the block row above are
//      plain 64-bit logic.  Union-find nodes are RUNS (maximal chains of linked blocks of one block row), numbered
//      by their ORDINAL in block-row-major order (= raster order of their first blocks, so the smallest ordinal of a
//      component is its first block in raster order, which is what OpenCV's numbering sorts by).  Every run first
//      takes ONE contact with the row above as its parent (plain stores), pointer jumping — lane-balanced over the
//      ordinals — flattens the resulting forest (depth halves per round), and only the contacts that are left, where
//      two trees meet, are real unions: shared-memory compare-and-swaps on roots only, finds halve their paths.
//      Parents are 16-bit and dense (4 KB for up to 2048 runs: 32 resident warps per SM; a strip with more runs —
//      salt-and-pepper noise — keeps its parents in a global scratch array, same code).  Outputs per strip: the
//      Xe / Xo words (2 KB), the run-start masks (512 B) and one 16-bit root ordinal per run (bit 15 = the root
//      touches a neighbouring strip): ~0.2 B/px instead of the 0.5 B/px per-block records of round 1.  Roots that
//      touch no other strip are final (root bitmap); the others register in the sparse global parent array.
//   2. ccl_seam_merge_kernel (thread = seam block row): 8-connectivity across strip seams on the sparse global parents.
//      ccl_line_kernel (CTA = line): seam roots that survived the merges enter the root bitmap; exclusive scan of the
//      bitmap (label = 1 + #roots before the root), island count; the last CTA to finish also scans the counts of all
//      lines into the stats row offsets.
//   3. ccl_strip_write2_kernel (CTA = strip, 256 threads): run roots -> final labels (shared-memory table), int32
//      labels out as 512-byte row segments, and, fused, the cv2 stats: every run's extent / area is reduced over the
//      lanes that hold runs of the same label (match_any + redux) before one set of global atomics per group.
// HBM traffic: 1 B/px mask read + 4 B/px labels written + ~0.4 B/px of records; no second pass over the labels.
#pragma once
#include "common.cuh"

namespace sd {

constexpr int kCw = 4;                       // warps (strips) per CTA
#ifndef SD_CCL_OCC_DEFAULT
#define SD_CCL_OCC_DEFAULT 8                 // label CTAs per SM the default instantiation is compiled for (SD_CCL_OCC overrides)
#endif
constexpr int kBigCoord = 0x3fffffff;        // stats rows keep (kBigCoord - max) so that every field is a min / add
constexpr int kStatFill = 0x7f7f7f7f;        // cudaMemsetAsync(0x7f) start value of every stats field

// -DSD_BOUNDS_CHECK (SD_EXTRA_NVCC_FLAGS): every index the CCL kernels compute is checked against the extent of the array
// it addresses; a violation prints the site and traps.  compute-sanitizer is closed on the pool this was developed on
// (profiles/r02_sanitize.md): tools/bounds_check.sh runs the CCL parity tests under this build instead.
#ifdef SD_BOUNDS_CHECK
#define CW_CHECK(cond)                                                                                        \
  do {                                                                                                        \
    if (!(cond)) {                                                                                            \
      printf("CW_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
      __trap();                                                                                               \
    }                                                                                                         \
  } while (0)
#else
#define CW_CHECK(cond) do { } while (0)
#endif

struct CclWarpWork {
  int64_t blk_total;    // extent of the block space (bounds checks)
  int n_strips;
  int* parent;          // [blk_total]       sparse: seam-touching local roots only (global block index -> parent)
  uint32_t* bitmap;     // [blk_total / 32]  bit = block is the root (first block) of a component
  int* prefix;          // [blk_total / 32]  exclusive count of root bits before this word, per line
  uint4* pix;           // [strips][128]     {Xe.lo, Xe.hi, Xo.lo, Xo.hi} of every pixel row of the strip
  uint2* rs;            // [strips][64]      run-start mask of every block row
  uint16_t* roots;      // [strips][4096]    one entry per run, block-row major: ordinal of its root run | touch << 15
  uint16_t* parent_fb;  // [strips][4096]    local parents of a strip with more runs than fit in shared memory (rare: noise)
  int* bnd_root;        // [strips][2][64]   global index of the root of each seam block (left / right column), -1 if none
  uint32_t* bnd_bits;   // [strips][2][4]    seam pixel columns: bit L of word r = pixel row 4L + r
  unsigned int* ticket; // [1]               lines finished (the last line CTA builds the stats offsets)
};

__device__ __forceinline__ uint32_t cw_nzflags(uint32_t w) {   // 0x80 in every non-zero byte
  return (w | ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
}
// 16 mask bytes -> 8 even-column bits (bit k = byte 2k non-zero) and 8 odd-column bits
__device__ __forceinline__ void cw_bits16(uint4 v, uint32_t& e, uint32_t& o) {
  const uint32_t f0 = cw_nzflags(v.x), f1 = cw_nzflags(v.y), f2 = cw_nzflags(v.z), f3 = cw_nzflags(v.w);
  e = __dp4a(f0, 0x00020001u, __dp4a(f1, 0x00080004u, __dp4a(f2, 0x00200010u, __dp4a(f3, 0x00800040u, 0u)))) >> 7;
  o = __dp4a(f0, 0x02000100u, __dp4a(f1, 0x08000400u, __dp4a(f2, 0x20001000u, __dp4a(f3, 0x80004000u, 0u)))) >> 7;
}
__device__ __forceinline__ uint64_t cw_shfl_up64(uint64_t v, int lane) {
  const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, 1), hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), 1);
  return lane ? (((uint64_t)hi << 32) | lo) : 0ull;
}
// block column of the run start that owns block k: highest run-start bit at or below k
__device__ __forceinline__ int cw_run_start(uint64_t rs, int k) { return 63 - __clzll((long long)(rs & ((2ull << k) - 1ull))); }

// Programmatic dependent launch: the kernels of the chain are launched with programmatic stream serialization, so the
// launch latency and prologue of kernel k+1 overlap the tail of kernel k.  Every kernel lets its dependents launch right
// away and waits for the COMPLETION (and memory flush) of its predecessor before it touches any data.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Union-find over the runs of one strip.  Parents are 16-bit and indexed by run ordinal: 4 KB of shared memory per strip
// for up to MAXRUNS = 2048 runs, which is what bounds the number of resident warps (the worst case, 4096 runs, is served
// from global memory through the same generic pointer).
typedef unsigned short cw_node_t;
// ordinal (within its block row, 0-based) of the run that owns block k: run starts at or below k, minus one
__device__ __forceinline__ int cw_run_index(uint64_t rs, int k) { return __popcll(rs & ((2ull << k) - 1ull)) - 1; }
// find with path halving.  Safe while unions are in flight because links are only ever created by a compare-and-swap
// on a ROOT (below): a non-root entry is never the target of a union, so re-pointing it at its grandparent (an
// ancestor for ever, parents only move towards the root) cannot undo a link; roots are never written here.
// (Round 1 linked with atomicMin on possibly stale roots, where a compressing store could drop a fresh link.)
__device__ __forceinline__ int cw_find(volatile cw_node_t* p, int a) {
  int q;
  while ((q = p[a]) != a) {
    const int g = p[q];
    if (g != q) p[a] = (cw_node_t)g;
    a = g;
  }
  return a;
}
// min-root union: the larger root is hung under the smaller one with a compare-and-swap that only succeeds while it
// still IS a root, so a link can never be overwritten.
__device__ __forceinline__ void cw_union(cw_node_t* p, int a, int b) {
  while (true) {
    a = cw_find(p, a);
    b = cw_find(p, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicCAS(&p[a], (cw_node_t)a, (cw_node_t)b);
    if (old == a) return;
    a = old;                                      // somebody re-parented a meanwhile; retry from its new parent
  }
}

// Contacts of a block row (top pixel row Te / To, links hl, run starts rs) with the pixel row above it (Ue / Uo; links
// hlU of that block row): bit k of vu / vl / vr = block k touches upper block k / k-1 / k+1, thinned so that a
// (run, upper run) pair that touches along several adjacent blocks is reported once.
struct CwContacts { uint64_t vu, vl, vr; };
__device__ __forceinline__ CwContacts cw_contact_masks(uint64_t Te, uint64_t To, uint64_t Ue, uint64_t Uo, uint64_t hl, uint64_t hlU) {
  const uint64_t vu0 = (Te | To) & (Ue | Uo);                 // block k - upper block k
  uint64_t vl = Te & (Uo << 1);                               // block k - upper block k-1 (diagonal)
  uint64_t vr = To & (Ue >> 1);                               // block k - upper block k+1 (diagonal)
  vl &= ~(vu0 & hlU) & ~((vu0 << 1) & hl);
  vr &= ~(vu0 & (hlU >> 1)) & ~((vu0 >> 1) & (hl >> 1));
  CwContacts c;
  c.vu = vu0 & ~((vu0 << 1) & hl & hlU);
  c.vl = vl; c.vr = vr;
  return c;
}
// Phase A: every run of the row (ordinals off, off + 1, ...) takes its FIRST contact with the row above (whose runs
// start at ordinal offU) as its parent: a plain store, the entry belongs to this lane alone, and a link always points to
// a smaller ordinal, so the forest is acyclic and a tree's root is its smallest run = the component's first block in
// raster order.  The contact is removed from the masks; what is left in them are the places where two trees meet (the
// bottom of a "V"), a small minority on handwriting.
__device__ __forceinline__ void cw_link_first(cw_node_t* parent, int off, int offU, uint64_t rs, uint64_t rsU, CwContacts& c) {
  for (uint64_t t = rs; t; t &= t - 1, ++off) {
    const int k = __ffsll((long long)t) - 1;
    const uint64_t above = t & (t - 1);                                              // run starts to the right of k
    const uint64_t ext = (above ? ((above & (~above + 1ull)) - 1ull) : ~0ull) & ~((1ull << k) - 1ull);   // blocks k .. next start - 1
    const uint64_t cu = c.vu & ext, cl = c.vl & ext, cr = c.vr & ext;
    int node = off;
    if (cu) { const int j = __ffsll((long long)cu) - 1; c.vu &= ~(1ull << j); node = offU + cw_run_index(rsU, j); }
    else if (cl) { const int j = __ffsll((long long)cl) - 1; c.vl &= ~(1ull << j); node = offU + cw_run_index(rsU, j - 1); }
    else if (cr) { const int j = __ffsll((long long)cr) - 1; c.vr &= ~(1ull << j); node = offU + cw_run_index(rsU, j + 1); }
    CW_CHECK(node >= 0 && node <= off && off < kStripBlocks);            // a link points to a smaller ordinal (or to itself)
    parent[off] = (cw_node_t)node;
  }
}
// Phase C: the contacts phase A left over, as real unions (compare-and-swap on roots) over the flattened trees
__device__ __forceinline__ void cw_union_rest(cw_node_t* parent, int off, int offU, uint64_t rs, uint64_t rsU, CwContacts c) {
  uint64_t vu = c.vu, vl = c.vl, vr = c.vr;
  while (vu | vl | vr) {
    int k, dk;
    if (vu) { k = __ffsll((long long)vu) - 1; vu &= vu - 1; dk = 0; }
    else if (vl) { k = __ffsll((long long)vl) - 1; vl &= vl - 1; dk = -1; }
    else { k = __ffsll((long long)vr) - 1; vr &= vr - 1; dk = 1; }
    CW_CHECK(k + dk >= 0 && k + dk < 64 && cw_run_index(rs, k) >= 0 && cw_run_index(rsU, k + dk) >= 0);
    cw_union(parent, off + cw_run_index(rs, k), offU + cw_run_index(rsU, k + dk));
  }
}
// position (block row * 64 + block column) of the run with ordinal r: row by bisection of the row offsets, column = the
// (r - rowoff[row])-th run start of that row.  Only used for the few roots that touch a strip seam.
__device__ __forceinline__ int cw_run_position(const uint16_t* rowoff, const uint64_t* rs, int r) {
  int lo = 0;
#pragma unroll
  for (int step = 32; step; step >>= 1) if (lo + step < 64 && rowoff[lo + step] <= r) lo += step;   // last row with rowoff <= r
  // rows without runs share the offset of the next row: step back is impossible (we took the LAST such row, which owns run r)
  uint64_t m = rs[lo];
  for (int n = r - rowoff[lo]; n > 0; --n) m &= m - 1;
  return lo * 64 + __ffsll((long long)m) - 1;
}

template <int MAXRUNS> struct __align__(16) CwLabelSmem {
  union {
    cw_node_t parent[MAXRUNS];                  // by run ordinal
    struct { uint8_t e[128][8], o[128][8]; } x; // Xe / Xo of every pixel row, one byte per 16-pixel load (until they are in registers)
  };
  uint64_t rs[64];                              // run starts of every block row
  uint32_t touch[kStripBlocks / 32];            // by run ordinal: the root touches a neighbouring strip
  uint16_t rowoff[64];                          // first ordinal of every block row
};

// warp-cooperative: #lines whose block offset is <= off, minus one (lines are sorted by offset)
__device__ __forceinline__ int cw_find_line(const sd_line* __restrict__ L, int n, int64_t off, int lane) {
  int cnt = 0;
  for (int i = lane; i < n; i += 32) cnt += (L[i].blk_off <= off) ? 1 : 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  return cnt - 1;
}

__device__ __forceinline__ int cw_uf_find(const int* __restrict__ parent, int a) {
  int p = __ldcg(parent + a);
  while (p != a) { a = p; p = __ldcg(parent + a); }
  return a;
}
__device__ __forceinline__ void cw_uf_union(int* parent, int a, int b) {
  while (true) {
    a = cw_uf_find(parent, a);
    b = cw_uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;
  }
}
__device__ __forceinline__ uint32_t cw_col_bit(uint4 p, int row) {
  const uint32_t w = (row & 2) ? ((row & 1) ? p.w : p.z) : ((row & 1) ? p.y : p.x);
  return (w >> (row >> 2)) & 1u;
}

// One block row of one strip seam: 8-connectivity between pixel column 127 of strip sg-1 and pixel column 0 of strip sg
// on the sparse global parents (bnd_root of the right side of a line's last strip is -1, so lines never merge).  All
// inputs are fetched up front (one round trip to L2) before the unions walk the parents.
__device__ __forceinline__ void cw_seam_row(const CclWarpWork& w, int64_t sg, int br) {
  const int a = __ldcg(w.bnd_root + (sg - 1) * 128 + 64 + br);
  const int* Rr = w.bnd_root + sg * 128;
  const int b0 = __ldcg(Rr + br), bm = br > 0 ? __ldcg(Rr + br - 1) : -1, bp = br < 63 ? __ldcg(Rr + br + 1) : -1;
  const uint4 Lb = __ldcg(reinterpret_cast<const uint4*>(w.bnd_bits + (sg - 1) * 8 + 4));
  const uint4 Rb = __ldcg(reinterpret_cast<const uint4*>(w.bnd_bits + sg * 8));
  if (a < 0) return;
  CW_CHECK(a < w.blk_total && b0 < w.blk_total && bm < w.blk_total && bp < w.blk_total && sg > 0 && sg < w.n_strips);
  const uint32_t a0 = cw_col_bit(Lb, 2 * br), a1 = cw_col_bit(Lb, 2 * br + 1);
  const uint32_t c0 = cw_col_bit(Rb, 2 * br), c1 = cw_col_bit(Rb, 2 * br + 1);
  if ((a0 | a1) & (c0 | c1)) cw_uf_union(w.parent, a, b0);
  if (br > 0 && a0 && cw_col_bit(Rb, 2 * br - 1)) cw_uf_union(w.parent, a, bm);
  if (br < 63 && a1 && cw_col_bit(Rb, 2 * br + 2)) cw_uf_union(w.parent, a, bp);
}

// MINB = resident CTAs per SM the register allocation aims at (8 -> 63 registers, no spills, 32 warps per SM; 10 -> 48
// registers, 40 warps; 12 -> 40 registers, 48 warps with MAXRUNS = 1536), MAXRUNS = runs whose parents fit in shared memory.
// Measured (profiles/r02_ccl_occ_merge_ab.json): more resident warps make the kernel SLOWER (36.8 / 46.7 / 50.2 us on 128
// text lines, 132 / 162 / 164 us on 512): the spills land on the same load/store pipe the union-find already queues on.
// 8 is the default; SD_CCL_OCC selects the others for A/B runs.
template <int MINB, int MAXRUNS>
__global__ void __launch_bounds__(32 * kCw, MINB) ccl_warp_label_kernel(const uint8_t* __restrict__ mask, const sd_line* __restrict__ L,
                                                                        int n_lines, int n_strips, CclWarpWork w) {
  extern __shared__ __align__(16) uint8_t cw_smem[];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  CwLabelSmem<MAXRUNS>& sm = reinterpret_cast<CwLabelSmem<MAXRUNS>*>(cw_smem)[wp];
  pdl_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0) *w.ticket = 0u;
  for (int strip = blockIdx.x * kCw + wp; strip < n_strips; strip += gridDim.x * kCw) {
    const int64_t blk0 = (int64_t)strip * kStripBlocks;
    const int li = cw_find_line(L, n_lines, blk0, lane);
    CW_CHECK(li >= 0 && li < n_lines);
    const sd_line ln = L[li];
    const int s = (int)((blk0 - ln.blk_off) >> 12), ns = ln.bw >> 6;
    CW_CHECK(s >= 0 && s < ns && ln.pitch == 2 * ln.bw && (ln.bw & 63) == 0);
    const uint8_t* src = mask + ln.px_off + s * 128;
    // all 128 rows of the strip on their way to L2 before the first batch of loads waits (a row = one 128-byte line)
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (int64_t)(i * 32 + lane) * ln.pitch));
    // ---- mask bytes -> Xe / Xo bits (a warp instruction moves 4 rows x 128 B) ----
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (b * 8 + i) * 4 + (lane >> 3);
        v[i] = __ldcs(reinterpret_cast<const uint4*>(src + (int64_t)row * ln.pitch + (lane & 7) * 16));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (b * 8 + i) * 4 + (lane >> 3);
        uint32_t e, o;
        cw_bits16(v[i], e, o);
        sm.x.e[row][lane & 7] = (uint8_t)e;
        sm.x.o[row][lane & 7] = (uint8_t)o;
      }
    }
    sm.touch[lane] = 0u; sm.touch[lane + 32] = 0u; sm.touch[lane + 64] = 0u; sm.touch[lane + 96] = 0u;
    __syncwarp();
    // ---- lane L: block rows a = 2L (pixel rows 4L, 4L+1) and b = 2L+1 (4L+2, 4L+3) ----
    const uint64_t* E = reinterpret_cast<const uint64_t*>(&sm.x.e[4 * lane][0]);
    const uint64_t* O = reinterpret_cast<const uint64_t*>(&sm.x.o[4 * lane][0]);
    const uint64_t Tea = E[0], Bea = E[1], Teb = E[2], Beb = E[3];
    const uint64_t Toa = O[0], Boa = O[1], Tob = O[2], Bob = O[3];
    {
      uint4* pp = w.pix + (int64_t)strip * 128 + 4 * lane;
      pp[0] = make_uint4((uint32_t)Tea, (uint32_t)(Tea >> 32), (uint32_t)Toa, (uint32_t)(Toa >> 32));
      pp[1] = make_uint4((uint32_t)Bea, (uint32_t)(Bea >> 32), (uint32_t)Boa, (uint32_t)(Boa >> 32));
      pp[2] = make_uint4((uint32_t)Teb, (uint32_t)(Teb >> 32), (uint32_t)Tob, (uint32_t)(Tob >> 32));
      pp[3] = make_uint4((uint32_t)Beb, (uint32_t)(Beb >> 32), (uint32_t)Bob, (uint32_t)(Bob >> 32));
    }
    const uint64_t Pea = Tea | Bea, Poa = Toa | Boa, Peb = Teb | Beb, Pob = Tob | Bob;
    const uint64_t occa = Pea | Poa, occb = Peb | Pob;
    const uint64_t hla = Pea & (Poa << 1), hlb = Peb & (Pob << 1);       // block k touches block k-1
    const uint64_t rsa = occa & ~hla, rsb = occb & ~hlb;                  // run starts
    w.rs[(int64_t)strip * 64 + 2 * lane] = make_uint2((uint32_t)rsa, (uint32_t)(rsa >> 32));
    w.rs[(int64_t)strip * 64 + 2 * lane + 1] = make_uint2((uint32_t)rsb, (uint32_t)(rsb >> 32));
    // run ordinals: the runs of the strip in block-row-major order (= the order of the root entries the write kernel reads)
    const int cnt_a = __popcll(rsa), cnt = cnt_a + __popcll(rsb);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const int n_runs = __shfl_sync(0xffffffffu, inc, 31);
    const int oa = inc - cnt, ob = oa + cnt_a;                            // first ordinals of the two block rows
    CW_CHECK(n_runs >= 0 && n_runs <= kStripBlocks && oa >= 0 && ob + __popcll(rsb) <= n_runs);
    // the block row above row a belongs to lane L-1 (its row b)
    const uint64_t Ue = cw_shfl_up64(Beb, lane), Uo = cw_shfl_up64(Bob, lane);
    const uint64_t hlU = cw_shfl_up64(hlb, lane), rsU = cw_shfl_up64(rsb, lane);
    const int oU = __shfl_up_sync(0xffffffffu, ob, 1);
    CwContacts ca = cw_contact_masks(Tea, Toa, Ue, Uo, hla, hlU);
    CwContacts cb = cw_contact_masks(Teb, Tob, Bea, Boa, hlb, hla);
    __syncwarp();                                                          // every lane has its Xe / Xo words: the parents may overwrite them
    sm.rs[2 * lane] = rsa; sm.rs[2 * lane + 1] = rsb;
    sm.rowoff[2 * lane] = (uint16_t)oa; sm.rowoff[2 * lane + 1] = (uint16_t)ob;
    cw_node_t* const par = n_runs <= MAXRUNS ? sm.parent : w.parent_fb + (int64_t)strip * kStripBlocks;
    // phase A: first contact of every run -> its parent (stores only)
    cw_link_first(par, oa, oU, rsa, rsU, ca);
    cw_link_first(par, ob, oa, rsb, rsa, cb);
    // phase B: pointer jumping flattens the forest (a vertical stroke is a chain of up to 64 runs: depth halves per
    // round).  Lane-balanced over the ordinals: text fills a few block rows with many runs and leaves the rest empty.
    {
      volatile cw_node_t* vp = par;
#pragma unroll 1
      for (int round = 0; round < 6; ++round) {
        __syncwarp();
        bool changed = false;
        for (int j = lane; j < n_runs; j += 32) {
          const int pa = vp[j];
          CW_CHECK(pa >= 0 && pa <= j);
          const int ga = vp[pa];
          if (ga != pa) { vp[j] = (cw_node_t)ga; changed = true; }
        }
        if (!__any_sync(0xffffffffu, changed)) break;
      }
    }
    __syncwarp();
    // phase C: the remaining contacts join trees
    cw_union_rest(par, oa, oU, rsa, rsU, ca);
    cw_union_rest(par, ob, oa, rsb, rsa, cb);
    __syncwarp();
    // ---- seam blocks: their roots touch a neighbouring strip ----
    const int gbase = (int)ln.blk_off + s * 64;                          // global index of local block i: gbase + (i >> 6) * bw + (i & 63)
    {
      volatile cw_node_t* vp = par;
      int* br = w.bnd_root + (int64_t)strip * 128;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t occ = h ? occb : occa, rs = h ? rsb : rsa;
        const int off = h ? ob : oa, row = 2 * lane + h;
        int left = -1, right = -1;
        if (s > 0 && (occ & 1ull)) {
          const int r = cw_find(vp, off);                                // block 0 has no left neighbour: it starts the row's first run
          atomicOr(&sm.touch[r >> 5], 1u << (r & 31));
          const int pos = cw_run_position(sm.rowoff, sm.rs, r);
          CW_CHECK(r >= 0 && r < n_runs && pos >= 0 && pos < kStripBlocks);
          left = gbase + (pos >> 6) * ln.bw + (pos & 63);
          CW_CHECK(left >= ln.blk_off && left < ln.blk_off + 64 * (int64_t)ln.bw && left < w.blk_total);
        }
        if (s < ns - 1 && (occ >> 63)) {
          const int r = cw_find(vp, off + cw_run_index(rs, 63));
          atomicOr(&sm.touch[r >> 5], 1u << (r & 31));
          const int pos = cw_run_position(sm.rowoff, sm.rs, r);
          CW_CHECK(r >= 0 && r < n_runs && pos >= 0 && pos < kStripBlocks);
          right = gbase + (pos >> 6) * ln.bw + (pos & 63);
          CW_CHECK(right >= ln.blk_off && right < ln.blk_off + 64 * (int64_t)ln.bw && right < w.blk_total);
        }
        br[row] = left; br[64 + row] = right;
      }
      // seam pixel columns 0 and 127 of the four pixel rows of this lane
      const uint32_t l0 = __ballot_sync(0xffffffffu, Tea & 1ull), l1 = __ballot_sync(0xffffffffu, Bea & 1ull);
      const uint32_t l2 = __ballot_sync(0xffffffffu, Teb & 1ull), l3 = __ballot_sync(0xffffffffu, Beb & 1ull);
      const uint32_t r0 = __ballot_sync(0xffffffffu, Toa >> 63), r1 = __ballot_sync(0xffffffffu, Boa >> 63);
      const uint32_t r2 = __ballot_sync(0xffffffffu, Tob >> 63), r3 = __ballot_sync(0xffffffffu, Bob >> 63);
      if (lane == 0) {
        uint4* bb = reinterpret_cast<uint4*>(w.bnd_bits + (int64_t)strip * 8);
        bb[0] = make_uint4(l0, l1, l2, l3);
        bb[1] = make_uint4(r0, r1, r2, r3);
      }
    }
    __syncwarp();
    // ---- one root entry per run (lane-balanced over the ordinals) ----
    {
      volatile cw_node_t* vp = par;
      uint16_t* out = w.roots + (int64_t)strip * kStripBlocks;
      for (int j = lane; j < n_runs; j += 32) {
        const int r = cw_find(vp, j);                                      // every union is done: roots are final
        CW_CHECK(r >= 0 && r <= j);
        const uint32_t tch = (sm.touch[r >> 5] >> (r & 31)) & 1u;
        out[j] = (uint16_t)(r | (tch << 15));
      }
      // ---- the roots themselves (row-owned: the position is known): interior roots -> bitmap, seam roots -> global parents
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = 2 * lane + h;
        int o = h ? ob : oa;
        uint64_t rootbits = 0ull;
        for (uint64_t t = h ? rsb : rsa; t; t &= t - 1, ++o) {
          if (vp[o] != o) continue;
          const int k = __ffsll((long long)t) - 1;
          if ((sm.touch[o >> 5] >> (o & 31)) & 1u) {
            const int g = gbase + row * ln.bw + k;
            CW_CHECK(g >= 0 && g < w.blk_total);
            w.parent[g] = g;
          }
          else rootbits |= 1ull << k;
        }
        *reinterpret_cast<uint2*>(w.bitmap + (ln.blk_off >> 5) + (int64_t)row * (ln.bw >> 5) + s * 2) =
            make_uint2((uint32_t)rootbits, (uint32_t)(rootbits >> 32));
      }
    }
    __syncwarp();                                                          // shared memory is reused by the next strip
  }
}

// thread = one block row of one strip seam.  (Measured alternative, rejected: the strip that finishes second at a seam
// merges it inside the label kernel — the serial global pointer chases on every warp's critical path cost +36 us on
// 66 Mpx against the 15 us of this kernel.)
__global__ void __launch_bounds__(256) ccl_seam_merge_kernel(CclWarpWork w, int n_strips) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_strips * 64 || (i >> 6) == 0) return;
  cw_seam_row(w, i >> 6, i & 63);
}

// CTA = line: seam roots that are still roots after the merges are component roots (root bitmap); then the exclusive
// scan of the bitmap (label = 1 + #roots before the root), island count; the last CTA to finish turns the counts of all
// lines into stats row offsets (stat_off[l] = sum over lines < l of (num - 1)).
// MERGE: the CTA first merges the seams of its own line (lines never share a component, so the seam unions of a line touch
// only that line's parents): one launch and one dependency edge fewer than ccl_seam_merge_kernel + this kernel.
template <bool MERGE>
__global__ void __launch_bounds__(1024) ccl_line_kernel(const sd_line* __restrict__ L, int n_lines, CclWarpWork w,
                                                        int* __restrict__ num_out, int64_t* __restrict__ stat_off) {
  const int l = blockIdx.x, tid = threadIdx.x;
  const sd_line ln = L[l];
  pdl_launch_dependents();
  pdl_wait();
  if (MERGE) {
    const int64_t first = ln.blk_off >> 12;                               // first strip of the line
    const int n_seam = ((ln.bw >> 6) - 1) * 64;
    for (int i = tid; i < n_seam; i += blockDim.x) cw_seam_row(w, first + 1 + (i >> 6), i & 63);
    __syncthreads();                                                      // the unions of this CTA are visible to its mark pass
  }
  const uint32_t* bitmap = w.bitmap + (ln.blk_off >> 5);
  int* prefix = w.prefix + (ln.blk_off >> 5);
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  __shared__ int s_last;
  if (tid == 0) s_carry = 0;
  {
    const int64_t first = (ln.blk_off >> 12) * 128;                       // bnd_root entries of this line's strips
    const int n_bnd = (ln.bw >> 6) * 128;
    for (int i = tid; i < n_bnd; i += blockDim.x) {
      const int k = __ldcg(w.bnd_root + first + i);
      CW_CHECK(k < w.blk_total && first + i < (int64_t)w.n_strips * 128);
      if (k >= 0 && __ldcg(w.parent + k) == k) atomicOr(&w.bitmap[k >> 5], 1u << (k & 31));
    }
  }
  __syncthreads();
  const int words = 2 * ln.bw;
  for (int c = 0; c < words; c += blockDim.x * 4) {
    const int i = c + tid * 4;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (i < words) v = __ldcg(reinterpret_cast<const uint4*>(bitmap + i));
    const int n0 = __popc(v.x), n1 = __popc(v.y), n2 = __popc(v.z), n3 = __popc(v.w);
    const int tot = n0 + n1 + n2 + n3;
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    int wv = (tid < 32) ? s_warp[tid] : 0;
    if (tid < 32) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wv, o);
        if (tid >= o) wv += t;
      }
    }
    const int carry = s_carry;
    __syncthreads();
    if (tid < 32) s_warp[tid] = wv;
    __syncthreads();
    const int woff = (tid >> 5) ? s_warp[(tid >> 5) - 1] : 0;
    const int ex = carry + woff + inc - tot;
    if (i < words) *reinterpret_cast<int4*>(prefix + i) = make_int4(ex, ex + n0, ex + n0 + n1, ex + n0 + n1 + n2);
    if (tid == blockDim.x - 1) s_carry = carry + woff + inc;
    __syncthreads();
  }
  if (tid == 0) {
    num_out[l] = s_carry + 1;                                             // cv2 counts the background label
    __threadfence();
    s_last = (atomicAdd(w.ticket, 1u) == (unsigned)gridDim.x - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!stat_off || !s_last) return;
  __threadfence();
  // every line's count is final: exclusive scan over the lines (warp 0; a few thousand lines at most)
  if (tid < 32) {
    int64_t carry = 0;
    for (int base = 0; base < n_lines; base += 32) {
      const int i = base + tid;
      const int64_t v = i < n_lines ? (int64_t)(__ldcg(num_out + i) - 1) : 0;
      int64_t inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (i < n_lines) stat_off[i] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (tid == 0) stat_off[n_lines] = carry;
  }
}

struct __align__(16) CwWriteSmem {
  int lab[kStripBlocks];           // run ordinal -> final label
  uint4 pix[128];
  uint2 rs[64];
  int rowoff[64];                  // first run ordinal of every block row
  int line;
};

// statistics of one run, merged over the lanes whose current run carries the same label, then one set of atomics
__device__ __forceinline__ void cw_stats_flush(int32_t* __restrict__ st_rows, int64_t cap_rows, int64_t row, uint32_t grp, int lane,
                                               int minx, int miny, int maxx, int maxy, int area) {
  minx = __reduce_min_sync(grp, minx); miny = __reduce_min_sync(grp, miny);
  maxx = __reduce_max_sync(grp, maxx); maxy = __reduce_max_sync(grp, maxy);
  area = __reduce_add_sync(grp, area);
  if (lane == __ffs(grp) - 1 && row < cap_rows) {
    int32_t* r = st_rows + row * 5;
    atomicMin(r + 0, minx); atomicMin(r + 1, miny);
    atomicMin(r + 2, kBigCoord - maxx); atomicMin(r + 3, kBigCoord - maxy);
    atomicAdd(r + 4, area);
  }
}

// CTA = strip, 256 threads.  Thread (block row br = tid >> 2, quarter q = tid & 3) owns the runs that START in its 16
// blocks: run roots -> final labels in shared memory, the fused cv2 stats of those runs; then warp wp expands block rows
// 8 wp .. 8 wp + 7, one 512-byte row segment per store instruction.
__global__ void __launch_bounds__(256) ccl_strip_write2_kernel(const sd_line* __restrict__ L, int n_lines, CclWarpWork w,
                                                               int* __restrict__ labels, const int64_t* __restrict__ stat_off,
                                                               int32_t* __restrict__ stats, int64_t cap_rows) {
  __shared__ CwWriteSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int strip = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const int64_t blk0 = (int64_t)strip * kStripBlocks;
  if (wp == 0) {
    const int li = cw_find_line(L, n_lines, blk0, lane);
    if (lane == 0) sm.line = li;
    // first run ordinal of every block row: exclusive scan of the run counts (two rows per lane)
    const uint2 r0 = __ldcs(w.rs + (int64_t)strip * 64 + 2 * lane), r1 = __ldcs(w.rs + (int64_t)strip * 64 + 2 * lane + 1);
    sm.rs[2 * lane] = r0; sm.rs[2 * lane + 1] = r1;
    const int c0 = __popc(r0.x) + __popc(r0.y), c1 = __popc(r1.x) + __popc(r1.y);
    int inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    sm.rowoff[2 * lane] = inc - c0 - c1; sm.rowoff[2 * lane + 1] = inc - c1;
  } else if (tid >= 128) {
    sm.pix[tid - 128] = __ldcs(w.pix + (int64_t)strip * 128 + (tid - 128));
  }
  __syncthreads();
  const int li = sm.line;
  CW_CHECK(li >= 0 && li < n_lines && strip < w.n_strips);
  const sd_line ln = L[li];
  const int s = (int)((blk0 - ln.blk_off) >> 12);
  CW_CHECK(s >= 0 && s < (ln.bw >> 6));
  const int gbase = (int)ln.blk_off + s * 64;
  const int br = tid >> 2, q = tid & 3;
  const uint2 r2 = sm.rs[br];
  const uint64_t rs = ((uint64_t)r2.y << 32) | r2.x;
  const uint64_t mine = rs & (0xFFFFull << (16 * q));                    // run starts inside this thread's 16 blocks
  const int ord0 = sm.rowoff[br] + __popcll(rs & ((1ull << (16 * q)) - 1ull));     // ordinal of this thread's first run
  const uint16_t* rt = w.roots + (int64_t)strip * kStripBlocks + ord0;
  // pass 1: runs that are roots (their entry is their own ordinal) compute their final label from their position
  {
    const uint16_t* p = rt;
    int ord = ord0;
    for (uint64_t t = mine; t; t &= t - 1, ++ord) {
      const uint32_t rr = __ldg(p++);
      if ((int)(rr & 0x7fffu) == ord) {
        const int k = __ffsll((long long)t) - 1;
        int g = gbase + br * ln.bw + k;
        if (rr >> 15) g = cw_uf_find(w.parent, g);
        CW_CHECK(ord >= 0 && ord < kStripBlocks && g >= ln.blk_off && g < ln.blk_off + 64 * (int64_t)ln.bw);
        sm.lab[ord] = 1 + __ldg(w.prefix + (g >> 5)) + __popc(__ldg(w.bitmap + (g >> 5)) & ((1u << (g & 31)) - 1u));
      }
    }
  }
  __syncthreads();
  // pass 2: every other run copies the label of its root; fused island statistics of the runs of this thread
  {
    const uint16_t* p = rt;
    const bool do_stats = stats != nullptr;
    const int64_t row0 = do_stats ? __ldg(stat_off + li) : 0;
    const uint4 t4 = sm.pix[2 * br], b4 = sm.pix[2 * br + 1];
    const uint64_t Te = ((uint64_t)t4.y << 32) | t4.x, To = ((uint64_t)t4.w << 32) | t4.z;
    const uint64_t Be = ((uint64_t)b4.y << 32) | b4.x, Bo = ((uint64_t)b4.w << 32) | b4.z;
    const uint64_t occ = Te | To | Be | Bo;
    int ord = ord0;
    for (uint64_t t = mine; t; t &= t - 1, ++ord) {
      const int k = __ffsll((long long)t) - 1;
      const int root = (int)(__ldg(p++) & 0x7fffu);
      CW_CHECK(root >= 0 && root <= ord && ord < kStripBlocks);
      const int lab = sm.lab[root];
      CW_CHECK(lab >= 1);
      if (root != ord) sm.lab[ord] = lab;
      if (do_stats) {
        // blocks of the run: occupied blocks from k up to the next run start of the ROW (it may lie in another quarter)
        const uint64_t above = rs & ~((2ull << k) - 1ull);
        const uint64_t upto = above ? ((above & (~above + 1ull)) - 1ull) : ~0ull;
        const uint64_t M = occ & upto & ~((1ull << k) - 1ull);
        const uint64_t Me = (Te | Be) & M, Mo = (To | Bo) & M;
        int minx = 1 << 20, maxx = -1;
        if (Me) { minx = 2 * (__ffsll((long long)Me) - 1); maxx = 2 * (63 - __clzll((long long)Me)); }
        if (Mo) { minx = min(minx, 2 * (__ffsll((long long)Mo) - 1) + 1); maxx = max(maxx, 2 * (63 - __clzll((long long)Mo)) + 1); }
        const bool top = ((Te | To) & M) != 0ull, bot = ((Be | Bo) & M) != 0ull;
        const int miny = 2 * br + (top ? 0 : 1), maxy = 2 * br + (bot ? 1 : 0);
        const int area = __popcll(Te & M) + __popcll(To & M) + __popcll(Be & M) + __popcll(Bo & M);
        // the lanes that are at a run right now (trip counts differ) pool the runs that carry the same label
        const uint32_t act = __activemask();
        const uint32_t grp = __match_any_sync(act, lab);
        cw_stats_flush(stats, cap_rows, row0 + lab - 1, grp, lane, s * 128 + minx, miny, s * 128 + maxx, maxy, area);
      }
    }
  }
  __syncthreads();
  // ---- expansion: warp wp writes block rows 8 wp .. 8 wp + 7, a store instruction = one 512-byte row segment ----
  int* out = labels + ln.px_off + s * 128 + lane * 4;
#pragma unroll 2
  for (int i = 0; i < 8; ++i) {
    const int b = wp * 8 + i;
    const uint2 q2 = sm.rs[b];
    const uint64_t rsb = ((uint64_t)q2.y << 32) | q2.x;
    const int off = sm.rowoff[b];
    const uint4 t4 = sm.pix[2 * b], b4 = sm.pix[2 * b + 1];
    const uint64_t Te = ((uint64_t)t4.y << 32) | t4.x, To = ((uint64_t)t4.w << 32) | t4.z;
    const uint64_t Be = ((uint64_t)b4.y << 32) | b4.x, Bo = ((uint64_t)b4.w << 32) | b4.z;
    const uint32_t te = (uint32_t)(Te >> (2 * lane)) & 3u, to = (uint32_t)(To >> (2 * lane)) & 3u;
    const uint32_t be = (uint32_t)(Be >> (2 * lane)) & 3u, bo = (uint32_t)(Bo >> (2 * lane)) & 3u;
    int l0 = 0, l1 = 0;
    if ((te | to | be | bo) & 1u) { CW_CHECK(cw_run_index(rsb, 2 * lane) >= 0 && off + cw_run_index(rsb, 2 * lane) < kStripBlocks); l0 = sm.lab[off + cw_run_index(rsb, 2 * lane)]; }
    if ((te | to | be | bo) & 2u) { CW_CHECK(cw_run_index(rsb, 2 * lane + 1) >= 0 && off + cw_run_index(rsb, 2 * lane + 1) < kStripBlocks); l1 = sm.lab[off + cw_run_index(rsb, 2 * lane + 1)]; }
    int4 a, c;
    a.x = (te & 1u) ? l0 : 0; a.y = (to & 1u) ? l0 : 0; a.z = (te & 2u) ? l1 : 0; a.w = (to & 2u) ? l1 : 0;
    c.x = (be & 1u) ? l0 : 0; c.y = (bo & 1u) ? l0 : 0; c.z = (be & 2u) ? l1 : 0; c.w = (bo & 2u) ? l1 : 0;
    __stcs(reinterpret_cast<int4*>(out + (int64_t)(2 * b) * ln.pitch), a);
    __stcs(reinterpret_cast<int4*>(out + (int64_t)(2 * b + 1) * ln.pitch), c);
  }
}

// stats rows -> cv2 layout (x, y, w, h, area); grid-stride over the rows actually used
__global__ void __launch_bounds__(256) ccl_stats_finish_kernel(int32_t* __restrict__ st, const int64_t* __restrict__ stat_off, int n_lines,
                                                               int64_t cap_rows) {
  pdl_wait();
  int64_t rows = stat_off[n_lines];
  if (rows > cap_rows) rows = cap_rows;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t* r = st + i * 5;
    const int minx = r[0], miny = r[1], maxx = kBigCoord - r[2], maxy = kBigCoord - r[3];
    r[2] = maxx - minx + 1; r[3] = maxy - miny + 1;
    r[4] = (int32_t)((uint32_t)r[4] - (uint32_t)kStatFill);
  }
}

}  // namespace sd
