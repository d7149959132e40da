/*
 * hgs_raster.h -- C ABI of the B200 (sm_100a) Gaussian rasterization hot path.
 *
 * Drop-in boundary: Horizon-GS reaches this path only through four
 * Python callables of the third-party package gsplat,
 *     gaussian_renderer/render.py:40-54    gsplat.rasterization
 *     gaussian_renderer/render.py:56-76    gsplat.rasterization_2dgs
 *     gaussian_renderer/render.py:149-165  gsplat.cuda._wrapper.fully_fused_projection
 *     gaussian_renderer/render.py:171-186  gsplat.cuda._wrapper.fully_fused_projection_2dgs
 * gsplat binds its CUDA stages to Python through a torch extension (gsplat/cuda/_wrapper.py ->
 * gsplat/cuda/csrc, not vendored in the reference).  The functions below are what such a binding
 * would call, one per stage, with plain device pointers and sizes; horizongs_b200/_lib.py binds
 * them with ctypes and horizongs_b200/cuda/_wrapper.py mirrors gsplat's Python operator names.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; row-major, contiguous, float32
 *     unless typed otherwise; `stream` is a cudaStream_t passed as void*.
 *   - functions only enqueue work on `stream`: no allocation, no synchronisation, no global state;
 *     re-entrant.  Return 0 on success, a positive cudaError_t, or a negative HGS_ERR_* code.
 *   - C = cameras, N = Gaussians, flat Gaussian index = c*N + n, I = tile intersections,
 *     tile grid = ceil(W/tile) x ceil(H/tile), tile id = ty*tile_w + tx.
 *   - viewmats [C,4,4] world->camera (OpenCV axes), Ks [C,3,3], quats wxyz (normalised inside),
 *     scales/opacities post-activation (render.py:40-54 and scene/basic_model.py:328-361).
 */
#ifndef HGS_RASTER_H
#define HGS_RASTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGS_ABI_VERSION 2
int hgs_abi_version(void);
/* cumulative count of kernels this library has launched in the process (diagnostics: bench.py reports the
 * number launched inside its timed region) */
unsigned long long hgs_debug_launch_count(void);
/* human-readable text for a status code returned by any function below */
const char* hgs_status_string(int status);

/* ---- a3: fully_fused_projection (render.py:149-165; inside rasterization render.py:40) ----------
 * out: radii[C,N] i32 (0 = culled), means2d[C,N,2], depths[C,N], conics[C,N,3] (a,b,c of the inverse
 * blurred 2D covariance), compensations[C,N] or NULL, tiles_per_gauss[C,N] i32 or NULL (stage a8's
 * count pass fused in; needs tile_size).  Culled rows are written as zeros. */
int hgs_project3d_fwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                      const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                      float far_plane, float radius_clip, int tile_size, int32_t* radii, float* means2d,
                      float* depths, float* conics, float* compensations, int32_t* tiles_per_gauss, void* stream);
/* in: upstream gradients v_means2d[C,N,2], v_depths[C,N] (or NULL), v_conics[C,N,3]; each with a row
 * stride in floats (ld_* = 2, 1, 3 when dense) so that slices of the packed blend-gradient buffer can be
 * passed without a copy;  out (overwritten, summed over cameras): v_means[N,3], v_quats[N,4], v_scales[N,3].
 * vis_ids (or NULL): work list of the n_vis visible flat indices c*N+n from hgs_isect_bin_prepare -- one thread per
 * visible pair instead of one per Gaussian (same results; with C > 1 the sums over cameras use atomics).
 * flags bit 0: v_means already holds a gradient (the SH direction gradient of hgs_sh_bwd, zero in the rows of culled
 * Gaussians) and the projection gradient is ADDED to it instead of overwriting it; bit 1 (work-list path): the
 * caller has already zero-filled the outputs (e.g. on a second stream, while the blend backward runs). */
int hgs_project3d_bwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                      const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                      float far_plane, const int32_t* radii, const float* v_means2d, int ld_means2d,
                      const float* v_depths, int ld_depths, const float* v_conics, int ld_conics,
                      const int32_t* vis_ids, long long n_vis, float* v_means, float* v_quats, float* v_scales,
                      int flags, void* stream);

/* ---- a4: fully_fused_projection_2dgs (render.py:171-186; inside rasterization_2dgs render.py:62) --
 * out: radii[C,N], means2d[C,N,2], depths[C,N], ray_transforms[C,N,3,3] (rows M0,M1,M2 of (K [R|t] H)),
 * normals[C,N,3] (camera frame, facing the camera), tiles_per_gauss or NULL. */
int hgs_project2d_fwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                      const float* Ks, int C, int N, int width, int height, float near_plane, float far_plane,
                      float radius_clip, int tile_size, int32_t* radii, float* means2d, float* depths,
                      float* ray_transforms, float* normals, int32_t* tiles_per_gauss, void* stream);
/* upstream gradients (any may be NULL) with row strides in floats (2, 1, 9, 3 when dense); vis_ids as in
 * hgs_project3d_bwd. */
int hgs_project2d_bwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                      const float* Ks, int C, int N, int width, int height, float near_plane, float far_plane,
                      const int32_t* radii, const float* v_means2d, int ld_means2d, const float* v_depths,
                      int ld_depths, const float* v_ray_transforms, int ld_ray_transforms, const float* v_normals,
                      int ld_normals, const int32_t* vis_ids, long long n_vis, float* v_means, float* v_quats,
                      float* v_scales, int flags, void* stream);

/* ---- a7: spherical_harmonics (inside rasterization* when sh_degree is not None) ------------------
 * Direction of Gaussian n for camera c is dirs[c,n,:] if dirs != NULL, else means[n,:] - campos[c,:]
 * (normalised inside).  coeffs[N,K,3] is shared by all cameras.  radii (or NULL) masks culled rows to 0.
 * post != 0 fuses gsplat's `clamp_min(colors + 0.5, 0)`.  out: colors[C,N,3].
 * vis_ids (or NULL) = work list of visible flat indices (then radii is not read, and ONLY the listed rows of colors
 * are written: the rows of culled Gaussians are left untouched).  n_vis_dev (or NULL): the length
 * of the work list as a DEVICE value (counts_dev[0] of hgs_isect_bin_prepare); n_vis is then only an upper bound, so the
 * call can be enqueued before the host has read the count. */
int hgs_sh_fwd(int degree, int K, const float* dirs, const float* means, const float* campos, const float* coeffs,
               const int32_t* radii, const int32_t* vis_ids, long long n_vis, const long long* n_vis_dev, int C, int N,
               int post, float* colors, void* stream);
/* out (overwritten): v_coeffs[N,K,3] summed over cameras; v_dirs[C,N,3] or NULL; v_means[N,3] or NULL
 * (direction gradient summed over cameras).  `colors` is the forward output (needed for the clamp mask
 * when post != 0).  ld_v_colors = row stride of v_colors in floats (3 when dense). */
int hgs_sh_bwd(int degree, int K, const float* dirs, const float* means, const float* campos, const float* coeffs,
               const int32_t* radii, const int32_t* vis_ids, long long n_vis, const float* colors,
               const float* v_colors, int ld_v_colors, int C, int N, int post, float* v_coeffs, float* v_dirs,
               float* v_means, int outputs_zeroed, void* stream);
/* outputs_zeroed != 0 (one-camera work-list path only): v_coeffs / v_means were zero-filled by the caller. */

/* ---- a8-a10: tile intersection, tile|depth key sort, per-tile ranges (integer, bit-exact) --------
 * key = cam << (32 + tile_bits) | tile_id << 32 | (int64)(int32 bits of depth), value = flat index;
 * tile_bits = floor(log2(tile_w*tile_h)) + 1.  Sorted order equals a stable ascending sort of the keys
 * emitted Gaussian-major / tile row-major (what gsplat's isect_tiles(sort=True) returns). */
int hgs_isect_count(const float* means2d, const int32_t* radii, long long CN, int tile_size, int tile_w, int tile_h,
                    int32_t* tiles_per_gauss, void* stream);
/* workspace sizes in bytes for the calls below */
size_t hgs_scan_temp_bytes(long long n);
size_t hgs_isect_bin_temp_bytes(long long CN, int C, int tile_w, int tile_h);
size_t hgs_isect_bin_bucket_bytes(long long n_super_isects);
/* exclusive prefix sum of in[n] (i32) -> out[n] (i64 accumulate, stored i32; HGS_ERR_TOO_LARGE is
 * reported through *total >= 2^31 being left for the caller to check); total written to total_dev[0]. */
int hgs_exclusive_scan_i32(const int32_t* in, int32_t* out, long long* total_dev, long long n, void* temp,
                           size_t temp_bytes, void* stream);
/* Unsorted emission (gsplat isect_tiles(sort=False)): cum = exclusive scan of tiles_per_gauss. */
int hgs_isect_emit(const float* means2d, const int32_t* radii, const float* depths, const int32_t* cum, int C, int N,
                   int tile_size, int tile_w, int tile_h, long long* isect_ids, int32_t* flatten_ids, void* stream);
/* Sorted path by BINNING (replaces gsplat's emit + 64-bit CUB radix sort + isect_offset_encode; same results).
 * The binning / sorting unit is a SUPER-TILE of 2x2 tiles: a Gaussian touches ~2.5x fewer of them than tiles, so
 * that many fewer keys are scattered and sorted; every tile's range is an order-preserving selection of its
 * super-tile's sorted keys.
 * Phase 1 (no host round trip): one pass over tiles_per_gauss compacts the Gaussians with at least one tile, in
 * ascending flat-index order (single-pass scan, decoupled look-back), and histograms the (camera, super-tile)
 * bins; one more launch scans the histogram.
 * out: visible_ids[<= CN] i32 (ascending: the work list of the per-Gaussian kernels that only touch visible
 * Gaussians), counts_dev[0] = n_vis, counts_dev[1] = I,
 * counts_dev[2] = number of (Gaussian, super-tile) pairs.
 * temp (hgs_isect_bin_temp_bytes) must be passed unchanged to phase 2. */
int hgs_isect_bin_prepare(const float* means2d, const int32_t* radii, const float* depths,
                          const int32_t* tiles_per_gauss, int C, int N, int tile_size, int tile_w, int tile_h,
                          int32_t* visible_ids, long long* counts_dev, void* temp, size_t temp_bytes, void* stream);
/* Phase 1 with its first launch fused into the projection: hgs_project3d_fwd_bin = hgs_project3d_fwd (tiles_per_gauss
 * required) + the compaction / histogram of hgs_isect_bin_prepare in ONE kernel (the Gaussians' tile boxes and depths
 * are still in registers there: no second pass over tiles_per_gauss, means2d, radii, depths); hgs_isect_bin_scan is
 * the remaining scan launch.  Same outputs and temp as hgs_isect_bin_prepare.
 * Shading, optionally fused as well (shade != -2): a visible Gaussian evaluates its colour and writes its 64-byte
 * blend record (what hgs_sh_fwd on the work list + hgs_blend3d_pack produce, bit for bit) while centre, conic and depth
 * are in registers.  shade = -1: feats = colours [N,3]; shade = 0..4: feats = SH coefficients [N,K,3], direction
 * means - campos[c], colors_out [C,N,3] receives max(SH + 0.5, 0) for the visible rows (other rows untouched).
 * opacities [N]; depth_channel != 0 puts the camera depth in the record's fourth channel (RGB+D / RGB+ED);
 * records: hgs_blend3d_pack_bytes(C*N) bytes, 16-byte aligned. */
int hgs_project3d_fwd_bin(const float* means, const float* quats, const float* scales, const float* viewmats,
                          const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                          float far_plane, float radius_clip, int tile_size, int32_t* radii, float* means2d, float* depths,
                          float* conics, float* compensations, int32_t* tiles_per_gauss, int32_t* visible_ids,
                          long long* counts_dev, void* temp, size_t temp_bytes, int shade, const float* feats, int K,
                          const float* opacities, const float* campos, int depth_channel, float* colors_out,
                          void* records, void* stream);
int hgs_isect_bin_scan(int C, int N, int tile_size, int tile_w, int tile_h, long long* counts_dev, void* temp,
                       size_t temp_bytes, void* stream);
/* Phase 2 (counts read back by the caller to size the outputs; n_visible_bound >= counts_dev[0] sizes the grid,
 * the exact counts are read on the device): every visible Gaussian drops a key (depth bits << 32 | flat index << 4
 * | mask of the super-tile's tiles it touches; C*N < 2^28) into the ranges of the super-tiles it touches; one CTA per (camera, super-tile) sorts its range on the SM (bitonic network
 * in registers / shuffles / shared memory; ranges > 2048 keys in chunks merged through L2) and counts the keys of
 * each of its tiles; a last kernel sums those counts into the per-tile range starts and selects every tile's keys,
 * in order, from its super-tile's sorted range.
 * in: counts_dev, temp of phase 1 (it holds the tile boxes and depth bits of the visible Gaussians); bucket: hgs_isect_bin_bucket_bytes(counts[2]) bytes of scratch.
 * out: isect_ids[I] i64, flatten_ids[I] i32, isect_offsets[C*tile_h*tile_w] i32.
 * n_isects and bucket_bytes / 8 are CAPACITIES (n_super_isects only sizes the check of bucket_bytes): a caller that has
 * read counts_dev passes the exact counts; a caller that has not may pass a guess, enqueue the call without waiting,
 * and compare counts_dev with its guess afterwards -- when I > n_isects or counts[2] > bucket_bytes / 8 every kernel of
 * this call returns at once without touching temp or the outputs, and the call is simply repeated with larger buffers.
 * n_visible_bound must be >= counts[0] (C*N always is). */
int hgs_isect_bin_sorted(const long long* counts_dev, int C, int N, long long n_visible_bound, long long n_isects,
                         long long n_super_isects, int tile_size, int tile_w, int tile_h, int32_t* isect_offsets,
                         long long* isect_ids, int32_t* flatten_ids, void* temp, size_t temp_bytes, void* bucket,
                         size_t bucket_bytes, void* stream);
/* gsplat isect_offset_encode on already-sorted keys. */
int hgs_isect_offset_encode(const long long* isect_ids, long long n_isects, int C, int tile_w, int tile_h,
                            int32_t* isect_offsets, void* stream);

/* ---- backward of a3 + a7 of one rasterized view, fused (the three calls hgs_blend3d_unpack, hgs_sh_bwd and
 * hgs_project3d_bwd over the same work list, bit for bit): one pass over the 48-byte rows of the packed gradient
 * buffer vpack [N,12] that hgs_blend3d_bwd_packed accumulated (v_means2d 0..1 | v_conics 2..4 | v_opacity 5 |
 * v_colors 8..10 | v_depth 11).  One camera; SH colours (degree 0..4, coefficients coeffs [N,K,3], campos [3],
 * colors [N,3] = the clamped forward colours).  in: vis_ids[n_vis] ascending ids of the visible Gaussians.
 * out (all zero-filled by the caller, rows of visible Gaussians written): v_means2d [N,2], v_opacities [N],
 * v_coeffs [N,K,3], v_means [N,3] (projection part + SH view-direction part), v_quats [N,4], v_scales [N,3]. */
int hgs_gauss_bwd_fused(const float* vpack, int has_depth, const int32_t* vis_ids, long long n_vis, int N,
                        const float* means, const float* quats, const float* scales, const float* viewmat,
                        const float* Kmat, int width, int height, float eps2d, float near_plane, float far_plane,
                        int sh_degree, int K, const float* campos, const float* coeffs, const float* colors,
                        float* v_means2d, float* v_opacities, float* v_coeffs, float* v_means, float* v_quats,
                        float* v_scales, void* stream);

/* ---- a11: rasterize_to_pixels (3DGS alpha blending) ----------------------------------------------
 * colors[C,N,CH]; if depths != NULL an extra channel CH (the camera-space depth) is blended after the
 * colours (render modes RGB+D / RGB+ED), so the output has D = CH + 1 channels, else D = CH.
 * backgrounds[C,D] or NULL.  out: render_colors[C,H,W,D] (= sum c*alpha*T + T_final*bg),
 * render_alphas[C,H,W,1] (= 1 - T_final), last_ids[C,H,W] i32 (index into flatten_ids of the last
 * blended Gaussian, for the backward pass).  Supported: CH + (depths?1:0) <= 32, tile_size == 16. */
int hgs_blend3d_fwd(const float* means2d, const float* conics, const float* colors, const float* depths,
                    const float* opacities, const float* backgrounds, int C, int N, int CH, int width, int height,
                    int tile_size, const int32_t* isect_offsets, const int32_t* flatten_ids, long long n_isects,
                    float* render_colors, float* render_alphas, int32_t* last_ids, void* stream);
/* Gradients are ACCUMULATED (+=) into v_means2d[C,N,2], v_conics[C,N,3], v_colors[C,N,CH], v_depths[C,N]
 * (NULL iff depths == NULL), v_opacities[C,N]; the caller zero-fills them.  v_means2d_abs is NULL or
 * receives sum |v_means2d| (gsplat absgrad). */
int hgs_blend3d_bwd(const float* means2d, const float* conics, const float* colors, const float* depths,
                    const float* opacities, const float* backgrounds, int C, int N, int CH, int width, int height,
                    int tile_size, const int32_t* isect_offsets, const int32_t* flatten_ids, long long n_isects,
                    const float* render_alphas, const int32_t* last_ids, const float* v_render_colors,
                    const float* v_render_alphas, float* v_means2d, float* v_means2d_abs, float* v_conics,
                    float* v_colors, float* v_depths, float* v_opacities, void* stream);

/* Fast path of a11 for <= 4 render channels (every mode the reference uses).
 * hgs_blend3d_pack writes one 64-byte record per Gaussian with radii > 0 (radii may be NULL = all):
 * centre, conic scaled by -log2(e)/2, opacity, cut-off exponent, colour (+ depth as the last channel when
 * depths != NULL).  records: hgs_blend3d_pack_bytes(C*N) bytes, 64-byte aligned.
 * n_vis_dev (or NULL): device-side length of vis_ids (n_vis is then an upper bound), as in hgs_sh_fwd.
 * The blend kernels gather records with TMA bulk copies.  D = channels incl. the depth channel;
 * normalize_depth != 0 fuses expected-depth normalisation (last channel / max(alpha, 1e-10)).
 * vpack[C*N,12] (zero-filled by the caller) accumulates per-Gaussian gradients:
 *   [0:2] v_means2d, [2:5] v_conics, [5] v_opacities, [8:8+D] v_colors (+ v_depths in the last channel). */
size_t hgs_blend3d_pack_bytes(long long CN);
int hgs_blend3d_pack(const float* means2d, const float* conics, const float* colors, const float* depths,
                     const float* opacities, const int32_t* radii, const int32_t* vis_ids, long long n_vis,
                     const long long* n_vis_dev, long long CN, int CH, void* records, void* stream);
int hgs_blend3d_fwd_packed(const void* records, const float* backgrounds, int C, int D, int normalize_depth,
                           int width, int height, int tile_size, const int32_t* isect_offsets,
                           const int32_t* flatten_ids, long long n_isects, float* render_colors,
                           float* render_alphas, int32_t* last_ids, void* stream);
int hgs_blend3d_bwd_packed(const void* records, const float* backgrounds, int C, int D, int normalize_depth,
                           int width, int height, int tile_size, const int32_t* isect_offsets,
                           const int32_t* flatten_ids, long long n_isects, const float* render_colors,
                           const float* render_alphas, const int32_t* last_ids, const float* v_render_colors,
                           const float* v_render_alphas, float* vpack, void* stream);

/* Dense copies of two columns of vpack for the rows listed in vis_ids (the others are zero-filled): v_means2d[CN,2]
 * and v_opacities[CN] -- what autograd consumes as dense tensors (retain_grad of meta["means2d"], the opacity leaf). */
int hgs_blend3d_unpack(const float* vpack, const int32_t* vis_ids, long long n_vis, long long CN, float* v_means2d,
                       float* v_opacities, int outputs_zeroed, void* stream);
/* rows[ids[j]] of an [*, row_floats] float buffer (row_floats % 4 == 0, 16-byte aligned) := 0 for j < n_ids: zeroes the
 * accumulator rows of the visible Gaussians only, where every reader of the buffer goes through the same work list. */
int hgs_zero_rows(float* rows, int row_floats, const int32_t* ids, long long n_ids, void* stream);

/* Measurement aid (not on the product path): counters[0] += P_eval, the (pixel, Gaussian) pairs a per-pixel
 * front-to-back loop visits before the pixel stops; counters[1] += P_blend, the pairs actually blended.
 * counters: eight zero-initialised uint64 on the device ([2..4]: culling statistics). */
int hgs_blend3d_stats(const void* records, int C, int width, int height, int tile_size,
                      const int32_t* isect_offsets, const int32_t* flatten_ids, long long n_isects,
                      unsigned long long* counters, void* stream);

/* ---- a12: rasterize_to_pixels_2dgs ----------------------------------------------------------------
 * As a11 with ray_transforms[C,N,3,3] and normals[C,N,3]; additionally blends normals, and writes
 * render_distort[C,H,W,1] (NULL = distortion off), render_median[C,H,W,1], median_ids[C,H,W]. */
int hgs_blend2d_fwd(const float* means2d, const float* ray_transforms, const float* colors, const float* depths,
                    const float* normals, const float* opacities, const float* backgrounds, int C, int N, int CH,
                    int width, int height, int tile_size, const int32_t* isect_offsets, const int32_t* flatten_ids,
                    long long n_isects, float* render_colors, float* render_alphas, float* render_normals,
                    float* render_distort, float* render_median, int32_t* last_ids, int32_t* median_ids,
                    void* stream);
/* v_densify[C,N,2] (or NULL) receives the screen-space positional gradient used for densification. */
int hgs_blend2d_bwd(const float* means2d, const float* ray_transforms, const float* colors, const float* depths,
                    const float* normals, const float* opacities, const float* backgrounds, int C, int N, int CH,
                    int width, int height, int tile_size, const int32_t* isect_offsets, const int32_t* flatten_ids,
                    long long n_isects, const float* render_colors, const float* render_alphas,
                    const int32_t* last_ids, const int32_t* median_ids, const float* v_render_colors,
                    const float* v_render_alphas, const float* v_render_normals, const float* v_render_distort,
                    const float* v_render_median, float* v_means2d, float* v_ray_transforms, float* v_colors,
                    float* v_depths, float* v_normals, float* v_opacities, float* v_densify, void* stream);

/* Fast path of a12 (1, 3 or 4 colour channels incl. the depth channel): 112-byte surfel records (centre,
 * opacity, ray transform, normal, colour, footprint ellipse) gathered by TMA bulk copies; per-warp culling;
 * vpack[C*N,24] (zero-filled by the caller) accumulates the per-surfel gradients:
 *   [0:2] v_means2d, [2:11] v_ray_transforms, [11:14] v_normals, [14] v_opacities, [16:16+D] v_colors
 *   (+ v_depths in the last channel), [20:22] densification gradient.
 * render_distort / v_render_distort NULL = distortion off. */
size_t hgs_blend2d_pack_bytes(long long CN);
int hgs_blend2d_pack(const float* means2d, const float* ray_transforms, const float* colors, const float* depths,
                     const float* normals, const float* opacities, const int32_t* radii, const int32_t* vis_ids,
                     long long n_vis, long long CN, int CH, void* records, void* stream);
int hgs_blend2d_fwd_packed(const void* records, const float* backgrounds, int C, int D, int normalize_depth,
                           int width, int height, int tile_size, const int32_t* isect_offsets,
                           const int32_t* flatten_ids, long long n_isects, float* render_colors,
                           float* render_alphas, float* render_normals, float* render_distort,
                           float* render_median, int32_t* last_ids, int32_t* median_ids, void* stream);
int hgs_blend2d_bwd_packed(const void* records, const float* backgrounds, int C, int D, int normalize_depth,
                           int width, int height, int tile_size, const int32_t* isect_offsets,
                           const int32_t* flatten_ids, long long n_isects, const float* render_colors,
                           const float* render_alphas, const int32_t* last_ids, const int32_t* median_ids,
                           const float* v_render_colors, const float* v_render_alphas,
                           const float* v_render_normals, const float* v_render_distort,
                           const float* v_render_median, float* vpack, void* stream);

/* ---- a13: 2DGS post-ops of rasterization_2dgs (render.py:56-76; gsplat depth_to_normal + normal rotation) ---
 * fwd: normals_world[C,H,W,3] = R_c2w normals_cam (both NULL to skip); normals_from_depth[C,H,W,3] (or NULL) from the
 * z-depth map depth[(c*H*W + pixel) * ld_depth] (e.g. the last channel of render_colors: ld_depth = channels): back-
 * projected pixel centres, central differences, normalised cross product, one-pixel zero border.  The camera-to-
 * world transform is the closed form (R^T, -R^T t) of viewmats[C,4,4]; Ks[C,3,3].
 * bwd: v_normals_cam[C,H,W,3] = R_c2w^T v_normals_world (NULL in: zeros); v_depth[(..) * ld_v_depth] = gradient of
 * the depth map through normals_from_depth (v_normals_from_depth NULL: zeros).  Either output may be NULL. */
int hgs_normals_post_fwd(const float* normals_cam, const float* depth, int ld_depth, const float* viewmats,
                         const float* Ks, int C, int H, int W, float* normals_world, float* normals_from_depth,
                         void* stream);
int hgs_normals_post_bwd(const float* depth, int ld_depth, const float* viewmats, const float* Ks, int C, int H, int W,
                         const float* v_normals_world, const float* v_normals_from_depth, float* v_normals_cam,
                         float* v_depth, int ld_v_depth, void* stream);

/* ---- f2 (next row of SURVEY.md section 8): densification statistics -------------------------------
 * One pass over the view-space gradient (scene/basic_model.py:96-144, the part fed by the rasterizer): for
 * Gaussians with radii > 0 in a view, grad_accum[n] += (mode_max ? max : sum) of ||v_means2d * (W/2, H/2)||,
 * denom[n] += number of views that saw n, max_radii[n] = max(max_radii[n], radii) (or NULL).
 * v_means2d[C,N,2] with row stride ld_means2d floats. */
int hgs_densify_stats(const float* v_means2d, int ld_means2d, const int32_t* radii, const int32_t* vis_ids,
                      long long n_vis, int C, int N, int width, int height, int mode_max, float* grad_accum,
                      float* denom, float* max_radii, void* stream);

/* ---- f3, first part (next row of SURVEY.md section 8): fused photometric L1 loss -------------------------
 * L = mean_{p,c<3} |render_colors[p,c] - gt[p,c]| + w_depth * mean_p render_colors[p,3] (D == 4 only)
 *     + w_alpha * mean_p render_alphas[p] (render_alphas may be NULL)          utils/loss_utils.py:17-18, train.py:158
 * P = pixels (C*H*W), D = 3 or 4 channels.  fwd: partials[hgs_l1_loss_partials()] scratch, loss[1] out (device).
 * bwd: v_loss[1] = upstream gradient (device); v_render_colors[P,D], v_render_alphas[P] (or NULL) overwritten. */
int hgs_l1_loss_partials(void);
int hgs_l1_loss_fwd(const float* render_colors, const float* render_alphas, const float* gt, long long P, int D,
                    float w_depth, float w_alpha, float* partials, float* loss, void* stream);
int hgs_l1_loss_bwd(const float* render_colors, const float* gt, const float* v_loss, long long P, int D,
                    float w_depth, float w_alpha, float* v_render_colors, float* v_render_alphas, void* stream);

/* ---- f3, second part: fused SSIM (utils/loss_utils.py:20-60, train.py:159) -----------------------------------
 * mean SSIM between channels 0..2 of render_colors[C,H,W,D] (channels-last, D >= 3) and gt[C,H,W,3], 11x11
 * Gaussian window (sigma 1.5), zero padding, C1 = 0.01^2, C2 = 0.03^2 -- the reference's ssim(img1, img2).
 * fwd: dmaps[3][C*H*W*3] receives d ssim / d{mu1, E[x^2], E[xy]} per pixel (kept for bwd), partials[hgs_ssim_partials]
 * scratch, ssim_mean[1] out (device).  bwd: v_ssim[1] = d loss / d ssim_mean (device); the gradient w.r.t.
 * render_colors[..., 0:3] is ADDED to v_render_colors[C,H,W,D]. */
long long hgs_ssim_partials(int C, int H, int W);
int hgs_ssim_fwd(const float* render_colors, const float* gt, int C, int H, int W, int D, float* dmaps, float* partials,
                 float* ssim_mean, void* stream);
int hgs_ssim_bwd(const float* render_colors, const float* gt, const float* dmaps, const float* v_ssim, int C, int H,
                 int W, int D, float* v_render_colors, void* stream);

/* ---- f1 (next row of SURVEY.md section 8): fused anchor -> neural-Gaussian decode -------------------------------
 * scene/basic_model.py:297-371 generate_neural_gaussians for view_dim 3 or 0, appearance_dim 0, feat_dim 32,
 * n_offsets k <= 16, color_dim = 3 (color_attr 'RGB') or 3 (d + 1)^2 <= 48 (color_attr 'SH<d>', lod_model.py:58-61):
 * three MLPs Linear(32 + view_dim, 32) -> ReLU -> Linear(32, {k, 7k, color_dim k}) (scene/lod_model.py:67-84; Tanh on
 * the opacity, optional Sigmoid on the colour) on cat(anchor_feat, unit(anchor - cam_center)) -- or on anchor_feat
 * alone when view_dim == 0 (basic_model.py:313-316).  HGS_ERR_TOO_LARGE when the weights do not fit shared memory.
 * mlp_host: HOST array of 12 device pointers {W1[32,32+view_dim], b1[32], W2[out,32], b2[out]} x {opacity, cov, colour}.
 * vis[V] (int64): indices of the visible anchors.  count: opac_all[V,k] = tanh(opacity MLP), bits[V] = mask of the
 * offsets with opacity > 0, cnt[V] = their number.  fwd: row0[V] (int64) = exclusive scan of cnt; writes the kept
 * Gaussians at rows row0[v] + rank: xyz[M,3], color[M,color_dim], opacity[M], scales[M,3], quats[M,4] (the rasterizer's
 * inputs; color is [M, color_dim]).  bwd: gradients of those five tensors -> g_anchor[A,3], g_feat[A,32], g_offset[A,k,3], g_scaling[A,6]
 * (rows of visible anchors with a kept offset are OVERWRITTEN; zero-fill first) and += into mlp_grad_host (12
 * device pointers, same shapes as mlp_host). */
int hgs_decode_count(const float* const* mlp_host, const float* anchor, const float* feat, const float* cam_center,
                     const long long* vis, long long V, int feat_dim, int k, int view_dim, int color_dim,
                     float* opac_all, int32_t* bits, int32_t* cnt, void* stream);
int hgs_decode_fwd(const float* const* mlp_host, const float* anchor, const float* feat, const float* offset,
                   const float* scaling, const float* cam_center, const long long* vis, long long V, int feat_dim,
                   int k, int view_dim, int color_dim, int color_sigmoid, const float* opac_all, const int32_t* bits, const long long* row0,
                   float* xyz, float* color, float* opacity, float* scales, float* quats, void* stream);
int hgs_decode_bwd(const float* const* mlp_host, float* const* mlp_grad_host, const float* anchor, const float* feat,
                   const float* offset, const float* scaling, const float* cam_center, const long long* vis,
                   long long V, int feat_dim, int k, int view_dim, int color_dim, int color_sigmoid,
                   const float* opac_all, const int32_t* bits,
                   const long long* row0, const float* v_xyz, const float* v_color, const float* v_opacity,
                   const float* v_scales, const float* v_quats, float* g_anchor, float* g_feat, float* g_offset,
                   float* g_scaling, void* stream);

/* LOD level test + anchor prefilter in one pass (scene/lod_model.py:286-290 set_anchor_mask with
 * basic_model.py:192-203 map_to_int_level; gaussian_renderer/render.py:120-197 prefilter_voxel): visible[a] = 1 iff
 * level[a] <= clamp(f(log2(standard_dist / (|anchor - cam_center| * resolution_scale)) / log2(fork) + extra_level[a]),
 * 0, max_level) (f = floor / round / ceil for level_mode 0 / 1 / 2; level == NULL skips the test) AND projecting
 * the anchor as a Gaussian with scales = scaling[a, 0:3] (row stride ld_scaling), quats = rotation[a] gives
 * radii > 0 -- the arithmetic of hgs_project3d_fwd.  viewmat[16], Kmat[9], cam_center[3]: device pointers. */
int hgs_anchor_filter(const float* anchor, const int32_t* level, const float* extra_level, const float* scaling,
                      int ld_scaling, const float* rotation, const float* cam_center, float resolution_scale,
                      float standard_dist, float fork, int max_level, int level_mode, const float* viewmat,
                      const float* Kmat, int N, int width, int height, float eps2d, float near_plane, float far_plane,
                      float radius_clip, uint8_t* visible, void* stream);

/* ---- e: exchange of view-sharded gradients over NVLink peer memory -----------
 * New behaviour (the reference trains one view per iteration in one process; gaussian_renderer/render.py has no
 * collective): with one view per GPU only the Gaussians a view sees have non-zero gradient rows, so the SUM over
 * ranks is done as a sparse all-reduce.  `tensors_host` / `widths_host` are HOST arrays describing up to
 * HGS_EXCHANGE_MAX_TENSORS dense row-major float32 device tensors [N, widths[k]] (the parameter gradients and the
 * densification statistics); a record is one id plus the concatenated rows (padded to a multiple of 4 floats).
 * Each rank owns a mailbox of hgs_exchange_mailbox_bytes() in its own memory, mapped into every peer.
 *   push   : records of the n_rows Gaussians listed in ids[] (unique, ASCENDING: the rank's visible set as written
 *            by hgs_isect_bin_prepare; n_rows <= cap_rows, a multiple of 32) are gathered, staged in shared memory and
 *            stored with TMA bulk copies into slot (step & 1, rank) of EVERY mailbox in mailboxes_host[world] (peer
 *            pointers; [rank] is the local one), then flag `rank` of every mailbox is raised to step + 1 (release,
 *            system scope).  n_ids = N, the number of rows of each tensor.
 *   reduce : for every block of consecutive Gaussian ids (n_ids = N rows in each tensor) waits (acquire) for flag
 *            src = 0..world-1 of the local mailbox, merges the sources' records in that order (ids[] must be
 *            ASCENDING, as hgs_isect_bin_prepare's visible_ids are) and overwrites the touched rows of the tensors --
 *            the same order on every rank, so all replicas end with bit-identical sums; rows no rank listed are
 *            left as they are (zero).  *status_dev (device int, zero-initialised) is set to 1 if a peer's flag
 *            does not arrive within 20 s.
 * `step` must increase by 1 per exchange on all ranks; both calls only enqueue kernels on `stream`. */
#define HGS_EXCHANGE_MAX_TENSORS 8
#define HGS_EXCHANGE_MAX_RANKS 16
#define HGS_PEER_HANDLE_BYTES 64
int hgs_exchange_row_floats(const int* widths_host, int n_tensors);
size_t hgs_exchange_mailbox_bytes(int world, long long n_ids, long long cap_rows, int row_floats);
int hgs_exchange_push(float* const* tensors_host, const int* widths_host, int n_tensors, long long n_ids,
                      const int32_t* ids, long long n_rows, long long cap_rows, void* const* mailboxes_host, int world,
                      int rank, unsigned long long step, void* stream);
int hgs_exchange_reduce(float* const* tensors_host, const int* widths_host, int n_tensors, long long n_ids,
                        long long cap_rows, const void* mailbox, int world, int rank, unsigned long long step,
                        int* status_dev, void* stream);
/* ---- e, fused form: SH / projection backward fused with the exchange (csrc/exchange_vjp.cu) -----------------
 * The SH-coefficient part of a view's gradient is rank one -- v_coeffs[k][c] = basis_k(mean - camera position) *
 * v_colour[c] -- and every rank holds the means, so the ranks exchange one 64-byte record per visible Gaussian
 * {v_means 3, v_opacity, v_quats 4, v_scales 3, densification norm, v_colour 3} instead of 38 gradients, replacing
 * hgs_sh_bwd + hgs_project3d_bwd + hgs_densify_stats + the gradient all-reduce of a step (one camera per rank).
 *   push   : one thread per row of ids[n_rows] (ascending) runs the camera-specific backward from the row the
 *            blend backward left in vpack[N,12] (see hgs_blend3d_bwd_packed): projection VJP, SH direction
 *            gradient (sh_degree >= 1: coeffs[N,K,3]; 0 or -1: none), SH clamp mask (colors_fwd[N,3] = the
 *            forward colours, or NULL), densification norm; the records and this rank's camera (DEVICE pointers
 *            viewmat[16], Kmat[9], campos[3]) go into slot (step & 1, rank) of every mailbox, then the rank's
 *            flag is raised (release).  multicast_base (or NULL): an NVSwitch multicast mapping of the same
 *            mailboxes (same offsets): every record, header and flag is then stored ONCE (multimem.st) and the
 *            switch replicates it to all ranks, instead of one unicast copy per peer.
 *   reduce : waits for all flags, then for every Gaussian sums the records of the sources that listed it, in rank
 *            order, expanding the SH part with that source's direction, and OVERWRITES v_means[N,3], v_quats[N,4],
 *            v_scales[N,3], v_opacities[N], v_coeffs[N,K,3] (sh_degree -1: plain colours, K == 1) -- zeros for
 *            Gaussians nobody saw; grad_accum[N] / denom[N] (or NULL) receive += the densification statistics of
 *            all views.  Bit-identical on every rank. */
size_t hgs_exchange_vjp_mailbox_bytes(int world, long long n_ids, long long cap_rows);
int hgs_exchange_vjp_push(int sh_degree, int K, const float* vpack, const float* colors_fwd, const float* viewmat,
                          const float* Kmat, const float* campos, const float* means, const float* quats,
                          const float* scales, const float* coeffs, int width, int height, float eps2d,
                          float near_plane, float far_plane, long long n_ids, const int32_t* ids, long long n_rows,
                          long long cap_rows, void* const* mailboxes_host, void* multicast_base, int world, int rank,
                          unsigned long long step, void* stream);
/* the same push for rasterization_2dgs: vpack24[N,24] is the row layout of hgs_blend2d_bwd_packed (has_depth: the
 * depth channel was rendered, its gradient sits in column 19); the surfel projection VJP replaces the 3DGS one and
 * the densification norm uses v_means2d + the densification gradient (columns 20:22).  The reduce is shared. */
int hgs_exchange_vjp_push_2dgs(int sh_degree, int K, const float* vpack24, int has_depth, const float* colors_fwd,
                               const float* viewmat, const float* Kmat, const float* campos, const float* means,
                               const float* quats, const float* scales, const float* coeffs, int width, int height,
                               float near_plane, float far_plane, long long n_ids, const int32_t* ids, long long n_rows,
                               long long cap_rows, void* const* mailboxes_host, void* multicast_base, int world, int rank,
                               unsigned long long step, void* stream);
int hgs_exchange_vjp_reduce(int sh_degree, int K, const float* means, long long n_ids, long long cap_rows,
                            const void* mailbox, int world, int rank, unsigned long long step, float* v_means,
                            float* v_quats, float* v_scales, float* v_opacities, float* v_coeffs, float* grad_accum,
                            float* denom, int* status_dev, void* stream);
/* Peer memory management (these allocate / synchronise, unlike the stage functions): a zero-filled device
 * allocation, its 64-byte CUDA IPC handle (host buffer) and the mapping of a peer's handle into this process. */
int hgs_peer_alloc(size_t bytes, void** out);
int hgs_peer_free(void* ptr);
int hgs_peer_export(void* ptr, unsigned char* handle64_host);
int hgs_peer_import(const unsigned char* handle64_host, void** out);
int hgs_peer_close(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* HGS_RASTER_H */
