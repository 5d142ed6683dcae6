/*
 * microaligner_b200 -- C ABI of the B200 (sm_100a) non-linear registration hot path.
 *
 * The reference (VasylVaskivskyi/microaligner) is pure Python and has no FFI of its own: its
 * native boundary is the list of cv2 / scikit-learn calls it makes on this path.  Each entry
 * point below replaces one of those call sites (cited as path:line relative to the reference
 * checkout) and is what a ctypes binding in the reference would load -- see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - images are row-major with an explicit row pitch in BYTES; flows are dense (h, w, 2)
 *     float32, x-displacement first, exactly the layout cv2 / the reference use;
 *   - `stream` is a cudaStream_t passed as void*; every call is stream-ordered, asynchronous,
 *     allocation-free and re-entrant (scratch comes from the caller via *_workspace_bytes);
 *   - return value 0 = ok, negative = error (ma_last_error() gives the text, thread-local);
 *   - "tile geometry" (T, ov): tile (i, j) is the window [iT-ov, (i+1)T+ov) x [jT-ov, (j+1)T+ov)
 *     of the image, zero outside it (shared_modules/slicer.py:23-118); results are stitched
 *     from tile centres (shared_modules/stitcher.py:72-118).  Tiles are never materialised.
 */
#ifndef MICROALIGNER_B200_H
#define MICROALIGNER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MA_OK 0
#define MA_ERR_INVALID (-1)
#define MA_ERR_CUDA (-2)
#define MA_ERR_WORKSPACE (-3)

/* pixel types of single-channel images */
#define MA_U8 0
#define MA_U16 1
#define MA_F32 2

int ma_version(void);
const char* ma_last_error(void);

/* Process-wide tuning options (none changes a result); reserved for A/B measurements of kernel variants --
 * every variant measured so far has either been promoted to the default or deleted (profiles/r02_ab_variants.log). */
#define MA_OPT_COUNT 4
int ma_set_option(int option, int value);

/* ---- image pyramid: cv.pyrDown (optflow_reg/optflow_registrator.py:194) -------------------
 * dst is ((h+1)/2, (w+1)/2); 5x5 binomial, BORDER_REFLECT_101, (s+128)>>8. dtype MA_U8|MA_U16. */
int ma_pyrdown(const void* src, size_t src_pitch, int h, int w, int dtype,
               void* dst, size_t dst_pitch, void* stream);

/* same, restricted to destination rows [row_begin, row_end) */
int ma_pyrdown_rows(const void* src, size_t src_pitch, int h, int w, int dtype,
                    void* dst, size_t dst_pitch, int row_begin, int row_end, void* stream);

/* ---- flow up-sampling: cv.pyrUp(flow*scale, dstsize) (optflow_registrator.py:140,150,164,169,212)
 * src (h,w,2) -> dst (dh,dw,2), dh in {2h-1,2h}, dw in {2w-1,2w}; `scale` is the reference's
 * pre-multiplication of the flow (1, 2 or 4). */
int ma_pyrup_flow(const float* src, int h, int w, float* dst, int dh, int dw, float scale, void* stream);

/* same, restricted to destination rows [row_begin, row_end) */
int ma_pyrup_flow_rows(const float* src, int h, int w, float* dst, int dh, int dw, float scale,
                       int row_begin, int row_end, void* stream);

/* ---- Warper.warp (optflow_reg/warper.py:37-76): per-tile cv.remap(INTER_LINEAR, constant 0) of
 * `img` by map = tile-local grid - flow, tile centres written to `out`. dtype MA_U8|MA_U16. */
int ma_warp_tiles(const void* img, size_t img_pitch, int dtype, const float* flow, int h, int w,
                  int T, int ov, void* out, size_t out_pitch, void* stream);

/* same, restricted to output rows [row_begin, row_end) (multi-GPU row bands) */
int ma_warp_tiles_rows(const void* img, size_t img_pitch, int dtype, const float* flow, int h, int w,
                       int T, int ov, void* out, size_t out_pitch, int row_begin, int row_end, void* stream);

/* ---- transform_img_with_tmat (shared_modules/utils.py:98-114, exported at microaligner/__init__.py:20):
 * pad_to_shape (utils.py:53-66) fused with skimage.transform.warp(img, AffineTransform(inv), order=1,
 * mode='constant', cval=0, preserve_range=True).astype(dtype).  `inv3x3` is a HOST pointer to the 9 doubles
 * (row-major) of the output(x, y, 1) -> input matrix, i.e. pinv([[tmat], [0, 0, 1]]) as the reference computes
 * it; float64 arithmetic, metric / affine / projective coordinate transform chosen from the matrix like
 * skimage's _warp_fast.  The src_h x src_w page sits at (pad_top, pad_left) of the zero out_h x out_w frame.
 * dtype MA_U8 | MA_U16.  img and out must not overlap. */
int ma_warp_affine(const void* img, size_t img_pitch, int dtype, int src_h, int src_w, int pad_top, int pad_left,
                   const double* inv3x3, void* out, size_t out_pitch, int out_h, int out_w, void* stream);

/* ---- merge_two_flows per tile + stitch (optflow_reg/optflow_registrator.py:37-47, 217-240):
 * per tile: max(f1)==0 -> f2; max(f2)==0 -> f1; else f1 + remap(f2, map = -f1).
 * workspace: ma_merge_workspace_bytes(h, w, T). */
size_t ma_merge_workspace_bytes(int h, int w, int T);
int ma_merge_flows_tiles(const float* f1, const float* f2, int h, int w, int T, int ov,
                         float* out, void* workspace, void* stream);

/* same, restricted to the tiles of tile rows [tile_row_begin, tile_row_end) */
int ma_merge_flows_tile_rows(const float* f1, const float* f2, int h, int w, int T, int ov, float* out,
                             void* workspace, int tile_row_begin, int tile_row_end, void* stream);

/* ---- opt-in, NOT reference behaviour: proper composition of two flows in image
 * coordinates, out(p) = f2(p) + bilinear(f1, p - f2(p)), rows [row_begin, row_end). */
int ma_compose_flows_rows(const float* f1, const float* f2, int h, int w, float* out,
                          int row_begin, int row_end, void* stream);

/* ---- tiled Farneback: TileFlowCalc.calc_flow / farneback (optflow_reg/flow_calc.py:30-98) =
 * cv.calcOpticalFlowFarneback(mov, ref, None, 0.5, 0, win, iters, 1, 1.7, FARNEBACK_GAUSSIAN)
 * on every tile window in [tile_begin, tile_end) (row-major tile index), centres stitched into
 * flow_out (h,w,2).  T<=0 selects the reference's untiled branch (flow_calc.py:61-64): one
 * tile = the whole image, no overlap.  mov/ref share dtype and pitch.
 * Workspace: ma_farneback_workspace_bytes(h, w, T, ov, n_batch) holds n_batch tiles in flight;
 * the call loops over batches internally (all on `stream`). */
size_t ma_farneback_workspace_bytes(int h, int w, int T, int ov, int n_batch);
int ma_farneback_tiles(const void* mov, const void* ref, size_t pitch, int dtype, int h, int w,
                       int T, int ov, int win, int iters, int tile_begin, int tile_end,
                       float* flow_out, void* workspace, size_t workspace_bytes, void* stream);

/* same with option flags.  MA_FB_CONTRACT_FMA lets the window blur contract multiply+add into FMA: 1.5x fewer
 * FP32 pipe cycles in the two dominant kernels, results within ~1e-6 px of the default (inside the 0.01 / 0.1 px
 * contract per call) but no longer bit-identical to OpenCV's unfused CPU arithmetic.  Off by default. */
#define MA_FB_CONTRACT_FMA 1u
/* MA_FB_FULL_WINDOWS: compute every iteration on the whole tile window.  By default each iteration is restricted to
 * the dependency cone of the stitched tile centre (the centre grown by (iterations - 1 - it) * (win / 2) pixels):
 * same stitched flow bit for bit, ~13 % less window-blur work at the default parameters.  For A/B measurements and tests. */
#define MA_FB_FULL_WINDOWS 4u
int ma_farneback_tiles_ex(const void* mov, const void* ref, size_t pitch, int dtype, int h, int w,
                          int T, int ov, int win, int iters, int tile_begin, int tile_end,
                          float* flow_out, void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* ---- OptFlowRegistrator.dog (optflow_reg/optflow_registrator.py:249-274): min-max -> [0,1] f32,
 * 41-tap separable Gaussians sigma 5 and 9 (REFLECT_101), difference, min-max -> u8.
 * An all-zero input yields an all-zero u8 output.  No host synchronisation: both global
 * reductions stay on the device.  workspace: ma_dog_workspace_bytes(h, w). Needs h,w >= 21. */
size_t ma_dog_workspace_bytes(int h, int w);
int ma_dog_u8(const void* src, size_t src_pitch, int dtype, int h, int w,
              uint8_t* dst, size_t dst_pitch, void* workspace, void* stream);

/* The two phases of ma_dog_u8 on a band of rows, for row-sharded execution: the caller supplies the
 * GLOBAL min/max of the source (device float[2]) and later of the difference image (after reducing the
 * per-band values ma_dog_diff_rows returns in diff_minmax).  Memory scales with the band, not the image:
 * `diff` holds rows [row_begin, row_end) only ((row_end-row_begin) x ma_dog_diff_pitch_floats(w) floats) and
 * the workspace is ma_dog_band_workspace_bytes(w, row_end-row_begin).  Source rows [row_begin-20,
 * row_end+20) must be valid.  ma_dog_quantize_rows reads a diff plane whose first row is image row diff_row0. */
size_t ma_dog_diff_pitch_floats(int w);
size_t ma_dog_band_workspace_bytes(int w, int nrows);
int ma_dog_diff_rows(const void* src, size_t src_pitch, int dtype, int h, int w, const float* src_minmax,
                     int row_begin, int row_end, float* diff, float* diff_minmax, void* workspace, void* stream);
int ma_dog_quantize_rows(const float* diff, int diff_row0, int h, int w, const float* diff_minmax, int row_begin,
                         int row_end, uint8_t* dst, size_t dst_pitch, void* stream);

/* ---- mi_tiled (shared_modules/similarity_scoring.py:27-50): normalized mutual information
 * (sklearn, arithmetic mean, natural log) of two u8 label images over consecutive chunks of
 * `chunk` row-major elements of the dense n-element arrays; scores_out[c] (double, device) gets
 * one NMI per chunk c < ceil(n/chunk).  The joint histograms live in distributed shared memory (one four-CTA cluster
 * per chunk): ma_nmi_workspace_bytes returns 0 and `workspace` may be NULL (both kept for ABI stability). */
size_t ma_nmi_workspace_bytes(size_t n, size_t chunk);
int ma_nmi_chunks(const uint8_t* a, const uint8_t* b, size_t n, size_t chunk,
                  double* scores_out, void* workspace, void* stream);

/* same for chunks [chunk_begin, chunk_end) only (other entries of scores_out are left untouched) */
int ma_nmi_chunk_range(const uint8_t* a, const uint8_t* b, size_t n, size_t chunk, size_t chunk_begin, size_t chunk_end,
                       double* scores_out, void* workspace, void* stream);

/* the gate's two comparisons in one launch (check_if_higher_similarity, similarity_scoring.py:61-68: mi_tiled(ref, warped)
 * and mi_tiled(ref, unwarped)): scores0[c] = NMI(a, b0), scores1[c] = NMI(a, b1) for chunks [chunk_begin, chunk_end).
 * b1 / scores1 may be NULL (one comparison). */
int ma_nmi_chunk_range2(const uint8_t* a, const uint8_t* b0, const uint8_t* b1, size_t n, size_t chunk, size_t chunk_begin,
                        size_t chunk_end, double* scores0, double* scores1, void* stream);

/* ---- pipeline input prep (shared_modules/utils.py:75-95): z max-projection of n_pages images
 * followed by cv.normalize(.., 0, 255, NORM_MINMAX, CV_8U). pages_host is a HOST array of n_pages
 * device pointers (same pitch/dtype). workspace: ma_zmip_workspace_bytes(h, w, dtype). */
size_t ma_zmip_workspace_bytes(int h, int w, int dtype);
int ma_zmip_normalize_u8(const void* const* pages_host, int n_pages, size_t pitch, int dtype,
                         int h, int w, uint8_t* dst, size_t dst_pitch, void* workspace, void* stream);

/* ---- small helpers used by the host layer --------------------------------------------------*/
/* global min/max of an image as two floats (device, out2[0]=min, out2[1]=max). */
int ma_minmax(const void* src, size_t pitch, int dtype, int h, int w, float* out2, void* stream);

/* ---- accounting: kernels launched by this library so far, and an optional per-kernel profiler that
 * brackets every launch with CUDA events on the launching stream (off by default; bench.py turns it on
 * for the timed region).  ma_profile_read drains pending events (synchronises on them). `units` is the
 * number of pixels / tile-pixels the launches processed. */
long long ma_launch_count(void);
int ma_profile_kernels(void);
const char* ma_profile_kernel_name(int id);
void ma_profile_enable(int on);
void ma_profile_reset(void);
int ma_profile_read(int id, double* total_ms, long long* launches, double* units);

/* ---- host-side helper of the TIFF page reader (no CUDA): TIFF-flavoured LZW (MSB-first 9..12-bit codes, early
 * change), `n` compressed bytes -> at most `cap` bytes at dst.  Returns the bytes written, -1 on a corrupt stream.
 * Replaces what tifffile (utils.py:69-72 `TiffFile.pages[i].asarray()`) delegates to its compiled codec extension. */
long long ma_tiff_lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
