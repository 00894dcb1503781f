/*
 * ubs_b200.h -- C ABI of libubs_b200.so, the B200 (sm_100a) rasteriser for Universal Beta Splatting.
 *
 * Boundary contract
 *   - extern "C", plain pointers and sizes, no torch / ATen types.  Every pointer is a DEVICE pointer unless
 *     its name starts with `h_`.  The caller owns all memory (PyTorch allocates, passes data_ptr()).
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and asynchronous.  No entry point
 *     synchronises the host unless documented.
 *   - Every function returns 0 on success or a negative UBS_E* code; ubs_last_error() returns a thread-local
 *     message for the last failure.  No exceptions cross the ABI.
 *   - FP32 arithmetic throughout; radii/tile counts/offsets/ids are int32, intersection keys int64, matching
 *     the reference dtypes (reference: submodules/gsplat/cuda/csrc/bindings.h:34-273).
 *
 * Each entry point cites the reference interface it replaces (paths relative to
 * /root/reference/submodules/gsplat/cuda/csrc unless stated otherwise).
 */
#ifndef UBS_B200_H
#define UBS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UBS_OK 0
#define UBS_EINVAL (-1)   /* bad argument (null pointer, unsupported size)            */
#define UBS_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed               */
#define UBS_ENOSPC (-3)   /* caller-provided capacity / workspace too small            */
#define UBS_EUNSUPPORTED (-4)

#define UBS_MAX_CHANNELS 16 /* colour channels handled by one compositing launch      */
#define UBS_MAX_RANKS 8     /* GPUs of one NVSwitch box in the sharded train step     */

/* ---- library ------------------------------------------------------------------------------------------ */
const char *ubs_last_error(void);
unsigned long long ubs_launch_count(void); /* kernels launched by this library in this process (bench bookkeeping) */
int ubs_version(void);           /* ABI version, bumped on any signature change                          */
int ubs_device_sm_count(void);   /* multiProcessorCount of the current device (grid sizing), <0 on error */

/* ---- K1: skew parameters -> first-order rotation ------------------------------------------------------ */
/* replaces l_triangle_to_rotmat_{fwd,bwd}_tensor (bindings.h:68-75; l_triagnle_to_rotmat_fwd.cu:8-36,
 * l_triagnle_to_rotmat_bwd.cu:8-25).  l_triangle [N,3] -> rot [N,3,3] = I + skew(a01,a02,a12).            */
int ubs_l_triangle_to_rotmat_fwd(int64_t N, const float *l_triangle, float *rot, void *stream);
int ubs_l_triangle_to_rotmat_bwd(int64_t N, const float *v_rot, float *v_l_triangle, void *stream);

/* ---- K2: (rot, scale, l_triangle) -> covariance ------------------------------------------------------- */
/* replaces rot_scale_l_triangle_to_covar_{fwd,bwd}_tensor (bindings.h:49-66;
 * rot_scale_l_triangle_to_covar_fwd.cu:8-188, ..._bwd.cu:8-235).  The strictly-lower entries of L follow
 * torch.tril_indices(D, D, -1) order (the only layout the reference caller ever builds:
 * scene/beta_model.py:69-73); D in [3, 8].  spatial_block != 0 -> only the 3x3 block is produced/consumed.  */
int ubs_rot_scale_l_triangle_to_covar_fwd(int64_t N, int D, int spatial_block, const float *rot, /* [N,9] */
                                          const float *scale,                                     /* [N,D] */
                                          const float *l_triangle, /* [N, D(D-1)/2] */
                                          float *covar,            /* [N,d,d], d = spatial ? 3 : D */
                                          void *stream);
int ubs_rot_scale_l_triangle_to_covar_bwd(int64_t N, int D, int spatial_block, const float *rot, const float *scale,
                                          const float *l_triangle, const float *v_covar, /* [N,d,d] */
                                          float *v_rot,                                   /* [N,9] */
                                          float *v_scale,                                 /* [N,D] */
                                          float *v_l_triangle,                            /* [N,D(D-1)/2] */
                                          void *stream);

/* ---- K3/K4: conditioning on the query (view direction [+ time]) --------------------------------------- */
/* replaces cond_mean_convariance_opacity_{fwd,bwd}_tensor (bindings.h:34-47;
 * cond_mean_convariance_opacity_fwd.cu:106-290, ..._bwd.cu:101-563).  D in [4, 8], Cd = D-3.
 * No gradient is produced for `query` (reference: cuda/_wrapper.py:611).                                    */
int ubs_cond_mean_covar_opacity_fwd(int64_t N, int D, const float *means, /* [N,D]   */
                                    const float *covars,                  /* [N,D,D] */
                                    const float *opacities,               /* [N]     */
                                    const float *betas,                   /* [N,Cd]  */
                                    const float *query,                   /* [N,Cd]  */
                                    float *out_means,                     /* [N,3]   */
                                    float *out_covars,                    /* [N,3,3] */
                                    float *out_opacities,                 /* [N]     */
                                    void *stream);
int ubs_cond_mean_covar_opacity_bwd(int64_t N, int D, const float *means, const float *covars, const float *opacities,
                                    const float *betas, const float *query, const float *v_out_means, /* [N,3]   */
                                    const float *v_out_covars,                                        /* [N,3,3] */
                                    const float *v_out_opacities,                                     /* [N]     */
                                    float *v_means,                                                   /* [N,D]   */
                                    float *v_covars,                                                  /* [N,D,D] */
                                    float *v_opacities,                                               /* [N]     */
                                    float *v_betas,                                                   /* [N,Cd]  */
                                    void *stream);

/* ---- K5/K6: world -> screen projection ---------------------------------------------------------------- */
/* replaces fully_fused_projection_{fwd,bwd}_tensor, covars path, perspective camera (bindings.h:129-178;
 * fully_fused_projection_fwd.cu:19-177, ..._bwd.cu:19-240).  Culled entries get radii = 0 and ZEROED outputs
 * (the reference leaves them uninitialised).  compensations / v_compensations may be NULL.                   */
int ubs_projection_fwd(int C, int64_t N, const float *means, /* [N,3] */
                       const float *covars,                  /* [N,6] xx,xy,xz,yy,yz,zz */
                       const float *viewmats,                /* [C,4,4] world->camera, row-major */
                       const float *Ks,                      /* [C,3,3] */
                       int width, int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                       int32_t *radii,       /* [C,N]   */
                       float *means2d,       /* [C,N,2] */
                       float *depths,        /* [C,N]   */
                       float *conics,        /* [C,N,3] */
                       float *compensations, /* [C,N] or NULL */
                       void *stream);
int ubs_projection_bwd(int C, int64_t N, const float *means, const float *covars, const float *viewmats,
                       const float *Ks, int width, int height, float eps2d, const int32_t *radii,
                       const float *conics, const float *compensations, /* NULL ok */
                       const float *v_means2d, const float *v_depths, const float *v_conics,
                       const float *v_compensations, /* NULL ok */
                       float *v_means,               /* [N,3]  accumulated over cameras (zeroed by callee) */
                       float *v_covars,              /* [N,6]  (xx, xy+yx, xz+zx, yy, yz+zy, zz)          */
                       float *v_viewmats,            /* [C,4,4] or NULL                                   */
                       void *stream);

/* ---- K7-K9: tile intersection, onesweep radix sort, tile offsets --------------------------------------- */
/* replaces isect_tiles_tensor + isect_offset_encode_tensor (bindings.h:180-199; isect_tiles.cu:16-333).
 * Two-phase so that the caller can size the exact outputs (the reference syncs the host the same way,
 * isect_tiles.cu:180-181), or skip the sync by passing a capacity bound.
 *
 *   ubs_isect_workspace_bytes(C*N, capacity)  -> bytes of scratch the next two calls need
 *   ubs_isect_count(...)       tiles_per_gauss[C,N], *n_isects (device int64) ; no host sync
 *   ubs_isect_emit_sort(...)   emits (key,val) pairs, stable LSD onesweep sort on bits
 *                              [0, 32+tile_n_bits+cam_n_bits), writes offsets[C,th,tw].
 *                              capacity = size of isect_ids / flatten_ids in elements; if *n_isects exceeds it
 *                              nothing past capacity is written.  status (NULL or two device int32): status[0] = 1 if
 *                              THIS call's list was truncated else 0 (rewritten by every call), status[1] counts the
 *                              truncated calls.  The projection-backward entry points take status as their skip_flag.
 * key = cam << (32+tb) | tile << 32 | (int64)(int32)float_bits(depth) ; val = cam*N + prim.                  */
size_t ubs_isect_workspace_bytes(int64_t CN, int64_t capacity);
int ubs_isect_count(int C, int64_t N, const float *means2d, const int32_t *radii, int tile_size, int tile_width,
                    int tile_height, int32_t *tiles_per_gauss, /* [C,N] */
                    int64_t *n_isects,                         /* [1] device */
                    void *workspace, size_t workspace_bytes, void *stream);
int ubs_isect_emit_sort(int C, int64_t N, const float *means2d, const int32_t *radii, const float *depths,
                        int tile_size, int tile_width, int tile_height, int do_sort,
                        const int32_t *tiles_per_gauss, /* from ubs_isect_count */
                        const int64_t *n_isects,        /* [1] device, from ubs_isect_count */
                        int64_t capacity, int64_t *isect_ids, /* [capacity] sorted keys out */
                        int32_t *flatten_ids,                 /* [capacity] sorted vals out */
                        int32_t *offsets,                     /* [C,th,tw] or NULL */
                        int32_t *status,                      /* [2] device or NULL */
                        void *workspace, size_t workspace_bytes, void *stream);
/* stand-alone offset encoding of already sorted keys (isect_tiles.cu:287-366). n_isects is a host value.     */
int ubs_isect_offset_encode(int64_t n_isects, const int64_t *isect_ids, int C, int tile_width, int tile_height,
                            int32_t *offsets, void *stream);

/* ---- K7-K9, B200-first route: tile binning + per-tile segment sort -------------------------------------- */
/* Same outputs as ubs_isect_count + ubs_isect_emit_sort (tiles_per_gauss, *n_isects, isect_ids and flatten_ids in
 * the reference's sorted order, offsets[C,th,tw]; reference: isect_tiles.cu:99-333) without a global sort: per-tile
 * pair counts come from a 2-D prefix sum of rectangle corner deltas, their exclusive scan is `offsets`, pairs are
 * written straight into their tile's segment and each segment is sorted in shared memory by (depth bits, flatten id)
 * -- the total order the reference's stable sort produces, so the result is bit-exact (csrc/bin_sort.cu).
 * Requires depths >= +0 for every primitive with radii > 0 (near_plane > 0); otherwise use ubs_isect_emit_sort.
 *   deltas_ready = 1: ubs_fused_project_fwd(tile_delta = workspace) already accumulated the corner deltas
 *   deltas_ready = 0: computed here from means2d / radii (tiles_per_gauss [C,N] is then written if non-NULL)
 * capacity bounds the pair arrays exactly as in ubs_isect_emit_sort (status[0] / status[1] as there; offsets are then
 * clamped to capacity).  No host synchronisation.                                                            */
size_t ubs_isect_bin_workspace_bytes(int C, int tile_width, int tile_height, int64_t capacity);
int ubs_isect_bin_sort(int C, int64_t N, const float *means2d, const int32_t *radii, const float *depths,
                       int tile_size, int tile_width, int tile_height, int deltas_ready,
                       int32_t *tiles_per_gauss, /* [C,N] out when deltas_ready == 0, may be NULL */
                       int64_t capacity, int64_t *n_isects, /* [1] device, out */
                       int64_t *isect_ids,                  /* [capacity] */
                       int32_t *flatten_ids,                /* [capacity] */
                       int32_t *offsets,                    /* [C,th,tw]  */
                       int32_t *status,                     /* [2] device or NULL */
                       void *workspace, size_t workspace_bytes, void *stream);
/* stand-alone stable radix sort of (int64 key, int32 val) pairs on bits [begin_bit, end_bit); n on device.   */
size_t ubs_radix_sort_workspace_bytes(int64_t capacity);
int ubs_radix_sort_pairs(const int64_t *n_dev, int64_t capacity, int64_t *keys_in, int32_t *vals_in, int64_t *keys_out,
                         int32_t *vals_out, int begin_bit, int end_bit, void *workspace, size_t workspace_bytes,
                         void *stream);

/* ---- K10/K11: per-tile alpha compositing with the Beta kernel ------------------------------------------ */
/* replaces rasterize_to_pixels_{fwd,bwd}_tensor (bindings.h:201-252; rasterize_to_pixels_fwd.cu:16-191,
 * rasterize_to_pixels_bwd.cu:16-276).  channels in [1, UBS_MAX_CHANNELS]; tile_size must be 16.
 * n_isects is read from DEVICE memory (int64) so the call composes with the capacity-bounded sort; it is clamped
 * to isect_capacity (the number of elements of flatten_ids) inside the kernels.                               */
int ubs_rasterize_fwd(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity, const float *means2d,
                      const float *conics,
                      const float *colors,      /* [C,N,channels] */
                      const float *opacities,   /* [C,N] */
                      const float *betas,       /* [C,N] */
                      const float *backgrounds, /* [C,channels] or NULL */
                      const uint8_t *masks,     /* [C,th,tw] bool or NULL */
                      int channels, int width, int height, int tile_size, const int32_t *offsets,
                      const int32_t *flatten_ids, float *render_colors, /* [C,H,W,channels] */
                      float *render_alphas,                             /* [C,H,W,1] */
                      int32_t *last_ids,                                /* [C,H,W]   */
                      void *stream);
int ubs_rasterize_bwd(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity, const float *means2d,
                      const float *conics,
                      const float *colors, const float *opacities, const float *betas, const float *backgrounds,
                      const uint8_t *masks, int channels, int width, int height, int tile_size, const int32_t *offsets,
                      const int32_t *flatten_ids, const float *render_alphas, const int32_t *last_ids,
                      const float *v_render_colors, const float *v_render_alphas,
                      /* gradients, ACCUMULATED into (caller zero-fills): */
                      float *v_means2d, float *v_conics, float *v_colors, float *v_opacities, float *v_betas,
                      void *stream);

/* Diagnostic work counters for the compositing roofline (no reference counterpart):
 * counts[8] (device u64) = { E_test, E_acc, E_cull, pairs staged, E_any, E_cull4, E_any4, E_any8x2 } -- see
 * csrc/rasterize_fwd.cu.                                                                                    */
/* The same two kernels gathering from the 48-byte splat rows ubs_fused_project_fwd writes (two sectors per pair
 * instead of five: the compositing kernels are sensitive to what stays in L1).  colors: NULL = colours out of the rows
 * themselves -- channels 3: RGB; channels 4: RGB + depth (render modes "RGB+D" / "RGB+ED"); channels 1: depth
 * ("Depth" / "EDepth" / "Normal"; reference: rendering.py:131-142) -- else a [C,N,channels] array as above.
 * The gradient outputs of the backward stay separate arrays (they feed ubs_fused_project_bwd); with colours out of
 * the rows and channels 4 or 1, v_colors is [C,N,3] (RGB part) and the depth channel's gradient goes to v_depths [C,N]
 * (must be given and zeroed by the caller; NULL otherwise).                                                     */
int ubs_rasterize_fwd_splats(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                             const float *splats, /* [C,N,12] */
                             const float *colors, const float *backgrounds, const uint8_t *masks, int channels,
                             int width, int height, int tile_size, const int32_t *offsets,
                             const int32_t *flatten_ids, float *render_colors, float *render_alphas,
                             int32_t *last_ids, void *stream);
int ubs_rasterize_bwd_splats(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                             const float *splats, const float *colors, const float *backgrounds,
                             const uint8_t *masks, int channels, int width, int height, int tile_size,
                             const int32_t *offsets, const int32_t *flatten_ids, const float *render_alphas,
                             const int32_t *last_ids, const float *v_render_colors, const float *v_render_alphas,
                             float *v_means2d, float *v_conics, float *v_colors, float *v_opacities, float *v_betas,
                             float *v_depths, void *stream);
/* RGB compositing backward of the fused path: colours from the splat rows, all screen-space gradients of a primitive
 * accumulated (vector reductions) into ONE 48-byte row, v_rows [C,N,12], zeroed by the caller:
 *   0..2  v_colors            3, 4, 5  v_conics = (r3, 2 r4, r5)
 *   6, 7  moment form of v_means2d = (2a r6 + 2b r7, 2b r6 + 2c r7) with the primitive's conic (a, b, c)
 *   8     v_opacities         9  v_betas / ln 2        10  v_depths (not written here)        11  unused
 * ubs_fused_project_bwd* take these rows as `v_rows` with rows_form = 1 (two sectors per visible primitive instead
 * of the eight of separate arrays).
 * Same results as ubs_rasterize_bwd_splats(channels = 3) up to the order of the float additions
 * (replaces rasterize_to_pixels_bwd.cu:106-274 for the (16-tile, 3-channel) instantiation).                     */
int ubs_rasterize_bwd_rows(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity, const float *splats,
                           const float *backgrounds, const uint8_t *masks, int width, int height, int tile_size,
                           const int32_t *offsets, const int32_t *flatten_ids, const float *render_alphas,
                           const int32_t *last_ids, const float *v_render_colors, const float *v_render_alphas,
                           float *v_rows,
                           const int32_t *skip_flag, /* NULL, or the `status` of the frame's tile-list build: a
                                                        truncated frame leaves v_rows untouched (zero gradient) */
                           void *stream);
/* Rows for ubs_fused_project_bwd* (rows_form = 0) from the separate gradient arrays of ubs_rasterize_bwd[_splats]:
 * 0..2 v_colors | 3..5 v_conics | 6, 7 v_means2d | 8 v_opacities | 9 v_betas | 10 v_depths | 11 zero.
 * v_colors / v_depths may be NULL (zeros).  CN = C * N.                                                          */
int ubs_pack_gradient_rows(int64_t CN, const float *v_means2d, const float *v_depths, const float *v_conics,
                           const float *v_opacities, const float *v_betas, const float *v_colors, float *v_rows,
                           void *stream);
int ubs_rasterize_count(int C, const int64_t *n_isects, int64_t isect_capacity, const float *means2d,
                        const float *conics, const float *opacities, const float *betas, int width, int height,
                        int tile_size, const int32_t *offsets, const int32_t *flatten_ids,
                        unsigned long long *counts, void *stream);

/* ---- fused fast path: raw parameters -> screen-space records ------------------------------------------- */
/* New entry (no single reference counterpart): fuses the activations (scene/beta_model.py:36-52,103-121),
 * the view-direction / timestamp query (scene/beta_model.py:675-690), K1, K2, K3 and K5 in one pass over the
 * PACKED primitive records, and the tile count of K7.  One launch covers all C cameras.
 *
 * Packed record (floats, row stride UBS_RECORD_STRIDE(D), 16-byte aligned rows):
 *   [0,3) xyz | [3,D) mean | [D,D+3) rgb | D+3 opacity logit | [D+4, 2D+2) beta (D-2, raw) |
 *   [2D+2, 3D+2) scale (raw, pre-softplus) | [3D+2, 3D+2+D(D-1)/2) l_triangle | zero padding to the stride
 * D=6: 35 floats -> stride 36 (144 B).  D=7: 44 floats -> stride 44 (176 B).                                 */
#define UBS_RECORD_FLOATS(D) (3 * (D) + 2 + (D) * ((D)-1) / 2)
#define UBS_RECORD_STRIDE(D) ((UBS_RECORD_FLOATS(D) + 3) / 4 * 4)
int ubs_record_stride(int D);

/* tile_delta: NULL, or the start of a ubs_isect_bin_sort workspace: the kernel then also adds every visible
 * primitive's four tile-rectangle corner deltas to it (the grid is zeroed first) and the call to
 * ubs_isect_bin_sort(..., deltas_ready = 1, ...) that follows produces *n_isects; with tile_delta == NULL the
 * pair count is written to *n_isects here (block sums in `workspace`, for ubs_isect_emit_sort).             */
int ubs_fused_project_fwd(int C, int64_t N, int D, const float *records, /* [N, stride] */
                          const float *viewmats,                         /* [C,4,4] */
                          const float *Ks,                               /* [C,3,3] */
                          const float *cam_pos,                          /* [C,3] camera centres (world) */
                          const float *timestamps,                       /* [C] (D=7) or NULL */
                          const uint8_t *prim_mask,                      /* [N] bool or NULL (viewer filter) */
                          int width, int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                          int calc_compensations, int tile_size, int tile_width, int tile_height,
                          int32_t *radii,           /* [C,N]   */
                          float *means2d,           /* [C,N,2] */
                          float *depths,            /* [C,N]   */
                          float *conics,            /* [C,N,3]; NULL = render-only frame: conics, opacities, betas,
                                                       colors and tiles_per_gauss are not written (the compositing
                                                       kernels gather from `splats`; no backward, no `meta`) */
                          float *opacities,         /* [C,N] conditioned (x compensation if requested) */
                          float *betas,             /* [C,N] spatial beta = 4 exp(raw beta_0) */
                          float *colors,            /* [C,N,3] or NULL (rgb copied out of the record) */
                          int32_t *tiles_per_gauss, /* [C,N] */
                          float *splats,            /* NULL, or [C,N,12]: the visible primitives' screen-space records
                                                       once more as 48-byte rows (mean2d.xy, opacity, beta | conic abc,
                                                       depth | rgb, 0) for ubs_rasterize_{fwd,bwd}_splats; rows of
                                                       culled primitives are left untouched */
                          int32_t *tile_delta,      /* NULL or start of the bin-sort workspace (see above) */
                          int64_t *n_isects,        /* [1] device (may be NULL when tile_delta is given) */
                          void *workspace, size_t workspace_bytes, /* ubs_isect_workspace_bytes(C*N, cap); unused
                                                                      when tile_delta is given */
                          int activated,      /* != 0: the records hold ACTIVATED values (softplus'd scales, sigmoid'd
                                                 opacity, 4 exp'd betas) -- what ubs_pack_records builds from the
                                                 tensors the reference's operator chain passes around */
                          const float *query, /* NULL, or [N, D-3]: the conditioning query of every primitive as the
                                                 caller computed it (scene/beta_model.py:675-690), used for every
                                                 camera instead of the view direction from cam_pos [+ timestamp] */
                          void *stream);
/* backward of the above: consumes gradients w.r.t. the screen-space records and writes a packed gradient
 * record buffer (same layout as `records`; accumulated over cameras; zeroed by callee).                      */
int ubs_fused_project_bwd(int C, int64_t N, int D, const float *records, const float *viewmats, const float *Ks,
                          const float *cam_pos, const float *timestamps, int width, int height, float eps2d,
                          int calc_compensations, const int32_t *radii, const float *conics,
                          const float *v_rows, /* [C,N,12] screen-space gradient rows */
                          int rows_form,       /* 1: as ubs_rasterize_bwd_rows writes them (moment form),
                                                  0: as ubs_pack_gradient_rows builds them from separate arrays */
                          float *v_records,    /* [N, stride] */
                          float *v_viewmats,   /* NULL, or [C,4,4]: gradient of the world-to-camera matrices through
                                                  the projection (fully_fused_projection_bwd.cu:178-201; zeroed by
                                                  the callee) -- _wrapper.py:898 `viewmats_requires_grad` */
                          int activated, const float *query, /* as in ubs_fused_project_fwd */
                          const int32_t *skip_flag, /* NULL, or the `status` of the frame's tile-list build: when
                                                       status[0] != 0 (the frame lost pairs to the capacity bound) the
                                                       gradient records are all zero -- the view is dropped */
                          void *stream);

/* ubs_fused_project_bwd with ubs_unpack_records fused in (the drop-in route): the gradient tiles leave the kernel as
 * the reference's seven separate tensors (layouts as in ubs_unpack_records; any destination may be NULL) and no
 * gradient record buffer exists.                                                                               */
int ubs_fused_project_bwd_unpacked(int C, int64_t N, int D, const float *records, const float *viewmats,
                                   const float *Ks, const float *cam_pos, const float *timestamps, int width,
                                   int height, float eps2d, int calc_compensations, const int32_t *radii,
                                   const float *conics, const float *v_rows, int rows_form, float *v_mean,
                                   float *v_rgb, float *v_opacity, float *v_beta0, float *v_beta_c, float *v_scale,
                                   float *v_l_triangle, float *v_viewmats, int activated, const float *query,
                                   const int32_t *skip_flag, void *stream);

/* Packing of the reference's separate per-primitive tensors into records and back (the zero-edit drop-in route,
 * ubs_b200/dropin.py; csrc/pack.cu).  mean [N,D] = xyz | conditional mean (get_mean, scene/beta_model.py:111-113),
 * rgb [N,3], opacity [N], beta0 [N] (spatial beta), beta_c [N,D-3], scale [N,D], l_triangle [N,D(D-1)/2]; values are
 * copied as they are (activated or raw).  ubs_unpack_records: any destination may be NULL (gradient not wanted).  */
int ubs_pack_records(int64_t N, int D, const float *mean, const float *rgb, const float *opacity, const float *beta0,
                     const float *beta_c, const float *scale, const float *l_triangle, float *records, void *stream);
int ubs_unpack_records(int64_t N, int D, const float *records, float *mean, float *rgb, float *opacity, float *beta0,
                       float *beta_c, float *scale, float *l_triangle, void *stream);

/* ---- the train step around the render --------------------------------------------- */
/* Photometric loss and its gradient: (1 - lambda) * mean|img - gt| + lambda * (1 - mean SSIM(img, gt)).
 * replaces l1_loss + fused_ssim + autograd (train.py:118-121; utils/loss_utils.py:18-19,45-85; fused_ssim is the
 * third-party rahul-goel/fused-ssim@1272e21 of setup.py:26 -- same 11x11 sigma-1.5 zero-padded window).
 * Images are addressed by element strides (n, channel, y, x), so the rendered [C,H,W,ch] buffer and a [ch,H,W]
 * ground-truth image are both read in place; v_img is written with img's strides.
 * loss_out: NULL or [3] device floats = {L1, SSIM, loss}.  v_img: NULL (evaluate only) or d(grad_scale * loss)/d img.
 * workspace: ubs_l1_ssim_workspace_bytes() bytes (16 suffice when v_img == NULL).                             */
size_t ubs_l1_ssim_workspace_bytes(int C, int channels, int height, int width);
int ubs_l1_ssim_loss(int C, int channels, int height, int width, const float *img, int64_t img_sn, int64_t img_sc,
                     int64_t img_sy, int64_t img_sx, const float *gt, int64_t gt_sn, int64_t gt_sc, int64_t gt_sy,
                     int64_t gt_sx, float lambda_dssim, float grad_scale, float *loss_out, float *v_img,
                     void *workspace, size_t workspace_bytes, void *stream);

/* torch.optim.Adam(eps) over the packed records, one learning rate per record column (h_lr: HOST array of
 * UBS_RECORD_STRIDE(D) doubles; padding columns are never moved).  replaces optimizer.step() over the seven
 * parameter groups of scene/beta_model.py:239-268 (train.py:169).  `step` counts from 1.  All four buffers are the
 * full [N, stride] arrays; a row range lets the caller pipeline chunks behind their gradient all-reduce.  opacity_reg / scale_reg
 * != 0 add the gradient of  opacity_reg * mean|sigmoid(raw opacity)| + scale_reg * mean|softplus(raw scale)[:3]|
 * (train.py:122-124; `[:3]` selects the first three primitives there and here).                               */
int ubs_adam_step(int64_t N, int D, int64_t row_begin, int64_t row_count, /* update rows [begin, begin+count) of N */
                  float *records, const float *grads, float *exp_avg, float *exp_avg_sq, const double *h_lr, double beta1, double beta2, double eps, int64_t step, double opacity_reg,
                  double scale_reg, void *stream);

/* ubs_fused_project_bwd + ubs_adam_step in one launch, for single-GPU batch-1 training (the reference's default,
 * arguments/__init__.py:103): the gradient records never touch HBM; `records`, `exp_avg`, `exp_avg_sq` are updated
 * in place.  Same arguments as the two calls it replaces (no v_records).                                      */
int ubs_fused_project_bwd_adam(int C, int64_t N, int D, float *records, const float *viewmats, const float *Ks,
                               const float *cam_pos, const float *timestamps, int width, int height, float eps2d,
                               int calc_compensations, const int32_t *radii, const float *conics,
                               const float *v_rows, int rows_form, /* as in ubs_fused_project_bwd */
                               float *exp_avg, float *exp_avg_sq, const double *h_lr, double beta1, double beta2, double eps,
                               int64_t step, double opacity_reg, double scale_reg,
                               const int32_t *skip_flag, /* as above; a truncated frame leaves records and moments
                                                            untouched */
                               void *stream);

/* ---- sharded train step over the GPUs of one NVSwitch box (no reference counterpart: SURVEY.md 2.3) ------------ */
/* Rows are split into `world` shards of `shard_rows` (a multiple of 128) rows; rank g owns shard g, i.e. its Adam
 * moments and the duty to update its parameters.  Pointers in the h_* arrays are HOST arrays of `world` DEVICE
 * addresses of symmetric (peer-mapped) buffers, entry g = rank g's buffer.
 *
 * ubs_fused_project_bwd_scatter: ubs_fused_project_bwd for one camera whose gradient tiles go, by bulk stores over
 * NVLink, into slot `rank` of the owner's staging buffer [world][shard_rows][stride] instead of a local v_records.
 * ubs_reduce_adam_gather: the owner sums the `world` slots of its staging buffer, applies ubs_adam_step's update to
 * its rows (moments are [shard_rows, stride], local) and stores the new parameters into every rank's records.
 * The caller separates the two, and the next forward, with a barrier across the ranks.                            */
int ubs_fused_project_bwd_scatter(int64_t N, int D, const float *records, const float *viewmats, const float *Ks,
                                  const float *cam_pos, const float *timestamps, int width, int height, float eps2d,
                                  int calc_compensations, const int32_t *radii, const float *conics,
                                  const float *v_rows, int rows_form, /* as in ubs_fused_project_bwd */
                                  int world, int rank, int64_t shard_rows, float *const *h_staging,
                                  const int32_t *skip_flag, /* as in ubs_fused_project_bwd: zero tiles are sent */
                                  void *stream);
/* Pull form of the sharded step (the default): what crosses NVLink before the update are the 48-byte screen-space
 * gradient rows, not the 144/176-byte parameter gradients.  Every rank renders its own view and leaves its rows
 * (ubs_rasterize_bwd_rows, with skip_flag) in a peer-mapped [N,12] buffer; after a barrier the owner of a shard runs
 * this: per 128-row tile, the rows of all `world` views arrive by bulk copies from the ranks' buffers, the projection
 * backward sums the views' gradients in registers (camera c = rank c's view: viewmats [world,4,4], Ks [world,3,3],
 * cam_pos [world,3], timestamps [world]), Adam is applied (moments [shard_rows, stride], local) and the new record
 * tile is stored into every rank's records.  A barrier follows.  Equal to ubs_fused_project_bwd_adam over the
 * world-camera batch, restricted to the shard.                                                                      */
int ubs_fused_project_bwd_adam_pull(int64_t N, int D, int world, int rank, int64_t shard_rows,
                                    float *const *h_peer_records, const float *const *h_peer_rows,
                                    const float *viewmats, const float *Ks, const float *cam_pos,
                                    const float *timestamps, int width, int height, float eps2d,
                                    int calc_compensations, float *exp_avg_shard, float *exp_avg_sq_shard,
                                    const double *h_lr, double beta1, double beta2, double eps, int64_t step,
                                    double opacity_reg, double scale_reg, void *stream);
int ubs_reduce_adam_gather(int64_t N, int D, int world, int rank, int64_t shard_rows, const float *staging,
                           float *exp_avg_shard, float *exp_avg_sq_shard, float *const *h_peer_records,
                           float *mc_records, /* NULL, or the NVLS multicast address of the records buffers: one
                                                 multimem.st per element instead of `world` peer stores */
                           const double *h_lr, double beta1, double beta2, double eps, int64_t step,
                           double opacity_reg, double scale_reg, void *stream);

/* MCMC relocation given the sampled indices: rows dst_idx[i] <- rows src_idx[i] with the opacity rescaled to
 * 1 - (1 - o)^(1/(m+1)), m = multiplicity of the source among src_idx, clamped to [0.005, 1 - eps]; the sources
 * take the same opacity and their Adam moments are zeroed (exp_avg / exp_avg_sq may both be NULL).
 * replaces the tensor part of relocate_gs / add_new_gs (scene/beta_model.py:512-657); dst and src rows must be
 * disjoint sets (dead vs alive primitives, or freshly appended rows).  counts: [N] int32 scratch.  The moment
 * buffers hold rows [moment_row_begin, moment_row_begin + moment_row_count): (0, N) for whole-buffer moments, the
 * rank's own shard in the sharded step (sources outside it are some other rank's to reset).                    */
int ubs_mcmc_relocate(int64_t N, int D, float *records, float *exp_avg, float *exp_avg_sq, int64_t moment_row_begin,
                      int64_t moment_row_count, int64_t K, const int64_t *dst_idx, const int64_t *src_idx,
                      int32_t *counts, void *stream);

/* SGLD position noise of the MCMC densification step (train.py:156-163): xyz += Sigma_xyz (noise (1 - sigmoid(raw
 * opacity))^100 noise_lr xyz_lr), Sigma_xyz = BetaModel.get_xyz_covariance (scene/beta_model.py:143-152, the
 * spatial_block of rot_scale_l_triangle_to_covar).  noise: [N,3] N(0,1) draws made by the caller (torch.randn_like
 * defines them).  In place on the xyz columns of the records.                                                   */
int ubs_sgld_noise(int64_t N, int D, float *records, const float *noise, double noise_lr, double xyz_lr, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* UBS_B200_H */
