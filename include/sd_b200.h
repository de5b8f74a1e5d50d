/* sd_b200.h — C ABI of the B200-native text-segmentation hot path.
 *
 * Drop-in boundary for parkseo7/Stroke-Derenderer's segmentation path.  The
 * reference is pure Python; its "FFI" for this path is the set of numpy / cv2 /
 * onnxruntime calls listed beside each entry point below (file:line relative to
 * the reference root).  A maintainer binds these with ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, a negative SD_E* code on failure;
 *    sd_last_error() returns a thread-local message.  No C++ exception crosses
 *    the ABI.
 *  - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch CUDA
 *    tensors); the library never frees caller memory.  h_* are host pointers.
 *  - every launch goes to the explicit `stream` (a cudaStream_t passed as
 *    void*); nothing uses the default stream, nothing synchronises unless the
 *    name ends in _sync.
 *  - one engine per GPU; an engine is not thread-safe.
 *  - there is NO CPU fallback: without a CUDA device the compute entry points
 *    fail with SD_ECUDA.
 */
#ifndef SD_B200_H
#define SD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_OK 0
#define SD_EINVAL (-1)
#define SD_ECUDA (-2)
#define SD_ESTATE (-3)
#define SD_ENOMEM (-4)

#define SD_TILE_H 128          /* evaluate_binarize.py:19 HEIGHT   */
#define SD_TILE_W 384          /* evaluate_binarize.py:20 WIDTH    */
#define SD_OVERLAP 64          /* evaluate_binarize.py:22 OVERLAP  */
#define SD_CIN_PAD 8           /* RGB padded to 8 halves (16 B) per pixel */
#define SD_CCL_CHUNK 4096      /* CCL strip: 64 x 64 blocks of 2x2 px; a line is a whole number of strips */

/* One text line of a batch (all lines already at height 128). */
typedef struct sd_line {
  int64_t img_off;    /* byte offset of the (128, width, 3) u8 RGB image in the packed input   */
  int64_t px_off;     /* element offset of the (128, pitch) plane in packed mask / label buffers */
  int64_t blk_off;    /* offset of the line in the CCL 2x2-block space (multiple of SD_CCL_CHUNK) */
  int32_t width;      /* W' = image width after resize_to_height                                  */
  int32_t n_tiles;    /* helper/split.py:19-26                                                    */
  int32_t wu;         /* w_unpad = W' // n_tiles (== max(W',1) when n_tiles == 1)                  */
  int32_t first_tile; /* index of the line's first tile in the batch tile stack                   */
  int32_t tile_w;     /* 384 */
  int32_t overlap;    /* 64  */
  int32_t pitch;      /* row pitch (elements) of the line's mask / label planes: round_up(W',128);
                         columns [W', pitch) of a mask plane must be zero                        */
  int32_t bw;         /* pitch / 2: 2x2-block columns (padded)                                    */
} sd_line;

/* Totals returned by sd_plan_lines. */
typedef struct sd_plan {
  int64_t img_bytes;   /* sum 128*W'*3           */
  int64_t px_total;    /* sum 128*pitch          */
  int64_t blk_total;   /* padded 2x2-block space */
  int32_t n_tiles;
  int32_t n_lines;
} sd_plan;

/* Where one tile's columns live in the packed line planes: column c < width of tile row y goes to
 * d_dst[y * pitch + c].  Filled by sd_tile_dst_table; consumed by sd_unet_forward_lines. */
typedef struct sd_tile_dst {
  uint8_t* d_dst;     /* device address of (row 0, first covered column) in the line's mask plane */
  int32_t pitch;      /* sd_line.pitch                                                            */
  int32_t width;      /* un-padded tile width (helper/split.py:31-34; <= tile_w)                   */
} sd_tile_dst;

/* Conv slots of the Attention-UNet (folded conv+BN), in execution order.
 * Weights are handed over as fp32 OIHW + fp32 bias; the library packs them. */
enum sd_slot {
  SD_CONV1_0 = 0, SD_CONV1_1, SD_CONV2_0, SD_CONV2_1, SD_CONV3_0, SD_CONV3_1,
  SD_CONV4_0, SD_CONV4_1, SD_CONV5_0, SD_CONV5_1,
  SD_UP5, SD_ATT5_G, SD_ATT5_X, SD_ATT5_PSI, SD_UPCONV5_0, SD_UPCONV5_1,
  SD_UP4, SD_ATT4_G, SD_ATT4_X, SD_ATT4_PSI, SD_UPCONV4_0, SD_UPCONV4_1,
  SD_UP3, SD_ATT3_G, SD_ATT3_X, SD_ATT3_PSI, SD_UPCONV3_0, SD_UPCONV3_1,
  SD_UP2, SD_ATT2_G, SD_ATT2_X, SD_ATT2_PSI, SD_UPCONV2_0, SD_UPCONV2_1,
  SD_HEAD, SD_NUM_SLOTS
};

/* Intermediate activations that sd_unet_read_tap can copy out (NHWC fp16). */
enum sd_tap {
  SD_TAP_X1 = 0, SD_TAP_X2, SD_TAP_X3, SD_TAP_X4, SD_TAP_X5,
  SD_TAP_D5U, SD_TAP_A4, SD_TAP_D5, SD_TAP_D4U, SD_TAP_A3, SD_TAP_D4,
  SD_TAP_D3U, SD_TAP_A2, SD_TAP_D3, SD_TAP_D2U, SD_TAP_A1, SD_TAP_D2, SD_NUM_TAPS
};

typedef struct sd_engine sd_engine;

const char* sd_last_error(void);
int sd_version(void);
/* "f16" or "bf16": the 16-bit operand type this library was built for (activations, packed weights, the tiles of
 * sd_tile_extract_f16 and the taps of sd_unet_read_tap); accumulation is fp32 either way. */
const char* sd_operand_dtype(void);
/* 1 if a CUDA device is usable from this process, else 0. */
int sd_cuda_available(void);

/* ---- host-side planning ------------------------------------------------- */
/* Tile geometry + packed-buffer offsets for a batch of line widths.
 * Replaces the bookkeeping of cut_and_stack (helper/split.py:57-79:
 * stack_indices, stack_widths, img_widths) and split_image (:16-37). */
int sd_plan_lines(const int32_t* h_widths, int n_lines, int tile_w, int overlap,
                  sd_line* h_lines_out, sd_plan* h_plan_out);

/* group_intervals + group_connections + add_to_group
 * (helper/partition.py:248-358), bit-exact incl. group and member order.
 * intervals: n pairs (a,b) sorted by a.  Writes member indices grouped
 * back-to-back into h_members (capacity n) and group start offsets into
 * h_group_start (capacity n+1); returns the number of groups (>=0) or <0. */
int sd_group_intervals(const int64_t* h_intervals_ab, int n, int64_t width,
                       int32_t* h_members, int32_t* h_group_start);

/* Batched island clustering of many lines from cv2-layout stats (the closed form of
 * get_binarized_islands + sort_islands + group_islands, helper/partition.py:17-26,31-69;
 * SURVEY.md A.5).  h_order holds, per line, np.argsort of the margin-expanded left
 * edges xs (the caller makes that exact numpy call: its tie order is observable,
 * SURVEY.md A.4).  Outputs: h_groups [n_groups][6] = (line, left, top, right,
 * bottom, canvas_off) in the reference's group order, h_group_of[row] = group id of
 * each island, h_line_group_start[n_lines+1].  Returns n_groups, or < 0. */
int64_t sd_group_lines(const int32_t* h_stats, const int64_t* h_stat_off, const int32_t* h_widths,
                       int n_lines, const int64_t* h_order, int margin, int img_h, int64_t target_w,
                       int64_t* h_groups, int32_t* h_group_of, int64_t* h_line_group_start,
                       int64_t* h_canvas_bytes);

/* Host: one sd_tile_dst per tile of the batch (stack order) for planes that start at device address
 * d_planes_base: the paste positions of reconstruct_images (helper/split.py:109-119: tile k of a line
 * lands at column k * wu with its un-padded width). */
int sd_tile_dst_table(const sd_line* h_lines, int n_lines, uint8_t* d_planes_base, sd_tile_dst* h_out);

/* ---- bandwidth-bound device stages --------------------------------------- */
/* K1a: split_image + pad_image + HWC->CHW stack (helper/split.py:10-54,81-84):
 * packed RGB lines -> (n_tiles, 3, 128, tile_w) u8, bit-exact. */
int sd_tile_extract_u8(const uint8_t* d_lines_rgb, const sd_line* d_lines, int n_lines,
                       int n_tiles, uint8_t* d_tiles_nchw, void* stream);
/* K1b: same cut, fused with `(x / 255.).astype(float32)`
 * (evaluate_binarize.py:99), emitted as NHWC fp16 with C padded to 8:
 * (n_tiles, 128, tile_w, 8) halves — the layout the first conv consumes. */
int sd_tile_extract_f16(const uint8_t* d_lines_rgb, const sd_line* d_lines, int n_lines,
                        int n_tiles, void* d_tiles_nhwc8, void* stream);

/* K6a: reconstruct_images (helper/split.py:89-124) on u8 tile outputs
 * (n_tiles, 128, tile_w), C == 1: un-pad, paste at stride wu, max on overlaps.
 * Output: packed (128, W') u8 planes at sd_line.px_off. */
int sd_glue_u8(const uint8_t* d_tiles, int n_tiles, const sd_line* d_lines, int n_lines,
               int64_t px_total, uint8_t* d_out, void* stream);
/* K6b: threshold (evaluate_binarize.py:103, strict >) + glue + main.py:108 in
 * one pass from fp16 probabilities (n_tiles, 128, tile_w); writes `on_value`
 * (255 to mirror binarize_images, 1 for the mask CCL consumes). */
int sd_glue_threshold_f16(const void* d_prob_f16, int n_tiles, const sd_line* d_lines,
                          int n_lines, int64_t px_total, float bin_thr, int on_value,
                          uint8_t* d_out, void* stream);

/* K7: cv2.connectedComponentsWithStats labels (helper/partition.py:14):
 * 8-connectivity, int32 labels numbered exactly as OpenCV numbers them
 * (SURVEY.md A.3).  d_mask: packed (128, W') u8, non-zero = foreground.
 * d_labels: packed int32 planes.  d_num: int32[n_lines] = N+1 (cv2's return
 * value, background included).  d_work: >= sd_ccl_workspace_bytes(). */
size_t sd_ccl_workspace_bytes(int64_t blk_total, int n_lines);
int sd_ccl_label(const uint8_t* d_mask, const sd_line* d_lines, int n_lines,
                 int64_t px_total, int64_t blk_total, int32_t* d_labels, int32_t* d_num,
                 void* d_work, void* stream);

/* K7 + K8 in one pass: the labels of sd_ccl_label plus the cv2 `stats` rows of sd_island_stats without a second
 * read of the labels (the island statistics are reduced while the labels are written).  d_stat_off: int64[n_lines+1]
 * OUT, row offset of every line (exclusive scan of num - 1; [n_lines] = total rows).  d_stats: int32[cap_rows][5];
 * rows beyond cap_rows are dropped: the caller checks d_stat_off[n_lines] <= cap_rows and retries with more room. */
int sd_ccl_label_stats(const uint8_t* d_mask, const sd_line* d_lines, int n_lines, int64_t px_total,
                       int64_t blk_total, int32_t* d_labels, int32_t* d_num, int64_t* d_stat_off,
                       int32_t* d_stats, int64_t cap_rows, void* d_work, void* stream);

/* K8: cv2 `stats` rows (x, y, w, h, area) int32 for labels 1..N of every line
 * (== cv2.boundingRect(labels == n), helper/partition.py:18-19).
 * d_stat_off[l] = row offset of line l in d_stats (row k-1 holds label k). */
int sd_island_stats(const int32_t* d_labels, const sd_line* d_lines, int n_lines,
                    int64_t px_total, const int64_t* d_stat_off, int64_t n_rows,
                    int32_t* d_stats, void* stream);

/* K9: group canvases (helper/partition.py:52-85): for group g with box
 * (left, top, right, bottom) in line `line`, canvas[y][x] =
 * group_of_label[label(top+y, left+x)] == g.  d_groups: int64[n_groups][6] =
 * (line, left, top, right, bottom, out_off).  d_group_of: int32 per stats row
 * (group id), d_stat_off as above. */
int sd_group_canvas(const int32_t* d_labels, const sd_line* d_lines,
                    const int64_t* d_groups, int n_groups, const int32_t* d_group_of,
                    const int64_t* d_stat_off, uint8_t* d_canvas, void* stream);
/* evaluate_strokes.py:202-222 + helper/partition.py:101-140 (resize_and_pad_image) + evaluate_strokes.py:58-69
 * (_normalize_image): per group, canvas -> cv2.normalize MINMAX -> cv2.resize INTER_LINEAR (8-bit fixed point,
 * bit-exact) to d_rs_dims[g] = (rs_w, rs_h) -> zero pad to size x size.  d_image_u8: (n_groups, size, size) =
 * the reference's `image`; d_input_f32 (optional): (n_groups, 3, size, size) = `image_input`, through d_lut =
 * 3 x 256 floats, lut[c][v] = (v / 255. - mean[c]) / std[c] computed by the caller in float64.  d_groups /
 * d_canvas as in sd_group_canvas. */
int sd_group_crops(const uint8_t* d_canvas, const int64_t* d_groups, const int32_t* d_rs_dims, int n_groups,
                   int size, uint8_t* d_image_u8, float* d_input_f32, const float* d_lut, void* stream);

/* Stroke-estimator front end, evaluate_strokes.py:72-91 (_encode_postprocess): encoder output d_enc (B, C, h, w) f32
 * -> d_out (B, 2h * 2w, C) f32, every value repeated on a 2 x 2 grid, channels last, positions flattened. */
int sd_encode_postprocess(const float* d_enc, int B, int C, int h, int w, float* d_out, void* stream);

/* common.py:85-93 / helper/split.py:127-135 (resize_to_height) for lines whose height is not 128:
 * cv2.resize(img, (dst_w, 128)) with the default INTER_LINEAR, bit-exact (8-bit fixed point, 2x area shortcut),
 * dst_w = int(w * (128 / h)) computed by the caller like the reference does.  Reads (src_h, src_w, 3) u8 at
 * d_src + src_off and writes (128, dst_w, 3) u8 at d_rgb + dst_off (= sd_line.img_off of the packed input).
 * Layout contract of d_src: every image is followed by at least 8 readable bytes (the kernel fetches pixel pairs
 * as aligned 32-bit words); images that do not start on a 4-byte boundary still work, through byte gathers. */
typedef struct sd_resize_job {
  int64_t src_off;
  int64_t dst_off;
  int32_t src_h, src_w, dst_w, reserved;
} sd_resize_job;
int sd_resize_lines(const uint8_t* d_src, const sd_resize_job* d_jobs, int n_jobs, int max_dst_w,
                    uint8_t* d_rgb, void* stream);

/* ---- host-side gather plumbing (results come back by D2H copies, no collective) ----------- */
/* Page-locks caller memory (e.g. a /dev/shm mapping shared by the ranks of a job) so that D2H copies land in it
 * directly at full PCIe rate; sd_host_unregister undoes it.  cudaHostRegister / cudaHostUnregister. */
int sd_host_register(void* h_ptr, size_t bytes);
int sd_host_unregister(void* h_ptr);
/* cudaMemcpyAsync device -> host on `stream` (asynchronous only when h_dst is page-locked). */
int sd_copy_d2h_async(void* h_dst, const void* d_src, size_t bytes, void* stream);

/* ---- Attention-UNet engine ------------------------------------------------ */
/* Replaces onnxruntime.InferenceSession (evaluate_binarize.py:48-53) and its
 * .run() (:62, :100).  max_tiles bounds one sd_unet_forward call. */
int sd_engine_create(int device, int max_tiles, int tile_h, int tile_w, sd_engine** out);
void sd_engine_destroy(sd_engine* e);
/* Folded conv+BN weights for one slot: w fp32 [cout][cin][k][k], b fp32 [cout]. */
int sd_engine_set_conv(sd_engine* e, int slot, const float* h_w, const float* h_b,
                       int cout, int cin, int k);
/* Packs to fp16, uploads, builds TMA descriptors.  impl: 0 = tcgen05 (product),
 * 1 = SIMT debug kernels (bring-up / cross-check only). */
int sd_engine_finalize(sd_engine* e, int impl);
int sd_engine_set_head_bias(sd_engine* e, float bias);
/* tiles: (n, tile_h, tile_w, 8) fp16 NHWC in [0,1].  Any output may be NULL.
 * d_prob_f32: (n, tile_h, tile_w) fp32 probabilities (ort.run output[0]);
 * d_prob_f16: same in fp16; d_mask_u8: 255 * (prob > bin_thr). */
int sd_unet_forward(sd_engine* e, const void* d_tiles_nhwc8, int n_tiles, float bin_thr,
                    float* d_prob_f32, void* d_prob_f16, uint8_t* d_mask_u8, void* stream);
/* The same forward with reconstruct_images (helper/split.py:89-124) + both thresholds
 * (evaluate_binarize.py:103, main.py:108) fused into the head: every tile ORs 255 * (prob > bin_thr) for its
 * un-padded columns straight into the packed line planes named by d_dst (one entry per tile, same order as
 * d_tiles).  The planes must be ZERO before the first tile of a line is processed (cudaMemsetAsync on the
 * same stream); tiles of one line may arrive in different calls. */
int sd_unet_forward_lines(sd_engine* e, const void* d_tiles_nhwc8, int n_tiles, float bin_thr,
                          const sd_tile_dst* d_dst, void* stream);
/* Copies an intermediate activation of the last forward (NHWC fp16) into
 * d_out; returns channels via *c, spatial dims via *h,*w. */
int sd_unet_read_tap(sd_engine* e, int tap, int n_tiles, void* d_out, size_t out_bytes,
                     int* c, int* h, int* w, void* stream);
/* Number of kernel launches issued by this library so far in this process. */
int64_t sd_launch_count(void);
/* Device time (ms) of each stage of the last forward when timing is enabled. */
int sd_engine_enable_timing(sd_engine* e, int on);
int sd_engine_layer_times(sd_engine* e, float* h_ms, int cap, int* n_out);
const char* sd_engine_layer_name(sd_engine* e, int i);
/* Debug builds only (-DSD_CONV_STATS): cycles spent in mbarrier waits of the tcgen05 conv kernels, per wait
 * code (1 producer<-empty slot, 2 MMA<-TMEM stage, 3 MMA<-full slot, 4 epilogue<-accumulator, 5 weights),
 * summed over the calling threads since the last reset; zeros in a normal build. */
int sd_debug_wait_cycles(unsigned long long* h_out8, int reset);
/* Code of the bounded mbarrier wait that timed out in a tcgen05 conv kernel (0 = none); readable even after the
 * trapped kernel has poisoned the CUDA context. */
int sd_engine_wait_error(sd_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* SD_B200_H */
