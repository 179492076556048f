/*
 * fatesplat.h -- C ABI of libfatesplat.so, the sm_100a 3D-Gaussian-splatting renderer that replaces the
 * native side of FateAvatar's rasterizer operators.
 *
 * Boundary being replaced (reference = /root/reference/submodules/diff-gaussian-rasterization, "DGR", and
 * /root/reference/submodules/simple-knn):
 *
 *   fs_forward        <- CudaRasterizer::Rasterizer::forward   DGR cuda_rasterizer/rasterizer.h:35-59,
 *                        rasterizer_impl.cu:198-336; bound to Python by RasterizeGaussiansCUDA,
 *                        DGR rasterize_points.cu:35-115 (pybind: ext.cpp:15)
 *   fs_backward       <- CudaRasterizer::Rasterizer::backward  DGR rasterizer.h:61-85, rasterizer_impl.cu:340-434;
 *                        RasterizeGaussiansBackwardCUDA, rasterize_points.cu:117-196 (ext.cpp:16)
 *   fs_mark_visible   <- CudaRasterizer::Rasterizer::markVisible DGR rasterizer.h:28-33, rasterizer_impl.cu:141-154;
 *                        markVisible, rasterize_points.cu:198-217 (ext.cpp:17)
 *   fs_knn_mean_dist2 <- SimpleKNN::knn  simple-knn/simple_knn.h, simple_knn.cu:186-222; distCUDA2, spatial.cu:15-26
 *   fs_pose_forward / fs_pose_backward <- the per-frame torch ops of model/fateavatar.py:225-258 (no native
 *                        counterpart upstream; see the declaration below)
 *   fs_flame_forward / fs_flame_backward <- flame/FLAME.py:131-204 + flame/lbs.py:24-100 (torch ops upstream)
 *
 * Differences from the reference native surface, all deliberate:
 *   - plain C, raw device pointers and sizes, explicit stream, no torch/glm/std::function types;
 *   - the three growable scratch buffers (geom/binning/img, DGR rasterize_points.cu:68-78) are ONE
 *     caller-owned workspace whose size comes from fs_workspace_bytes(); the library never allocates;
 *   - no host synchronisation: num_rendered is written to the workspace header on the device and, if the
 *     caller passes a pinned host pointer, copied there asynchronously on `stream`;
 *   - every call returns a status (0 = ok, <0 = error; text via fs_last_error()) instead of throwing.
 *
 * All pointers named `d_*` are device pointers; absent optional inputs are NULL exactly where the
 * reference receives an empty tensor (rasterize_points.cu:97-103).  Matrices are the 16 floats of the
 * reference's *transposed* view / full-projection tensors (points are row vectors).
 */
#ifndef FATESPLAT_H
#define FATESPLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS_OK 0
#define FS_ERR_INVALID_ARGUMENT (-1)
#define FS_ERR_WORKSPACE_TOO_SMALL (-2)
#define FS_ERR_CUDA (-3)
#define FS_ERR_UNSUPPORTED (-4)

/* Header at byte 0 of the workspace (device memory).  fs_forward also mirrors it to `h_info` (pinned). */
typedef struct fs_frame_info {
    uint32_t num_rendered; /* R: number of (Gaussian, tile) instances, == reference's return value          */
    uint32_t overflow;     /* 1 if R exceeded the workspace's instance capacity (frame is incomplete)       */
    uint32_t num_visible;  /* Gaussians with radii > 0                                                      */
    uint32_t max_tile_instances; /* heaviest tile's instance count                                         */
    uint32_t reserved[4];  /* [0] prefiltered-violation flag, [1] forward work counter, [2] depth segments, [3] non-empty tiles */
} fs_frame_info;

/* Byte offsets (from the workspace base) of the named per-frame arrays: the parity taps.
 * They play the role of GeometryState / BinningState / ImageState (DGR rasterizer_impl.h:30-66). */
typedef struct fs_workspace_layout {
    size_t total_bytes;
    size_t info;          /* fs_frame_info                                                   */
    size_t depths;        /* float  [P]                                                      */
    size_t cov3D;         /* float  [P][6]                                                   */
    size_t splat;         /* float4 [P][3]: {mx,my,ex,ey} {conic.x,conic.y,conic.z,opacity} {r,g,b,bits(gaussian id)} */
    size_t clamped;       /* uint8  [P][4] (3 used: SH colour clamp flags)                   */
    size_t rect;          /* uint16 [P][4] (x0,y0,x1,y1) tile rectangle                      */
    size_t tiles_touched; /* uint32 [P]                                                      */
    size_t tile_count;    /* uint32 [Tn]                                                     */
    size_t tile_cursor;   /* uint32 [Tn]                                                     */
    size_t ranges;        /* uint32 [Tn][2]  (start,end) into point_list; (0,0) if empty     */
    size_t big_tiles;     /* uint32 [Tn+1]   [0]=count, then ids of tiles too large for the smem sort */
    size_t work_order;    /* uint32 [Tn]     tile ids, heaviest first (work list of the blend kernels) */
    size_t tile_meta;     /* uint32 [Tn][4]  (range start, range end, first segment id, 0): one 16-byte load per unit */
    size_t seg_base;      /* uint32 [Tn+1]   first depth-segment id of every tile (segments of 256 list positions) */
    size_t seg_info;      /* uint32 [Smax][2] (tile, segment index inside the tile), Smax = Rcap/256 + Tn + 1 */
    size_t ckpt;          /* float4 [Smax][256] per-pixel (T, accumulated colour) before each segment boundary:
                             lets the backward blend start in the middle of a tile's list */
    size_t final_C;       /* float4 [H*W]    accumulated colour without background (backward: colour behind a
                             boundary = (final_C - ckpt.C) / ckpt.T) */
    size_t inst_keys;     /* uint64 [Rcap]   (depth_bits<<32 | gaussian) per tile segment    */
    size_t inst_keys_alt; /* uint64 [Rcap]   ping-pong for the large-tile global sort        */
    size_t point_list;    /* uint32 [Rcap]   sorted Gaussian ids (== reference point_list)   */
    size_t inst_splat;    /* float4 [Rcap][3] splat records gathered in sorted order         */
    size_t final_T;       /* float  [H*W]                                                    */
    size_t n_contrib;     /* uint32 [H*W]                                                    */
    size_t bwd_counter;   /* uint32 (256-byte slot) work counter of the backward blend, directly before grad_acc */
    size_t grad_acc;      /* float  [P][12]  backward accumulator: dmean2D.xy, dconic.xyw, dopacity, drgb, 3 pad */
    size_t instance_capacity; /* Rcap (count, not bytes)                                     */
    size_t pair_mask;     /* uint32 [Rcap][8] written by the forward blend: for tile t (list range [s,e)), 8x4 block b
                             and list position j the word at (8*s + b*(e-s) + j) has bit p set iff pixel p of the
                             block blended that instance -- the backward blend reads the forward's decisions
                             instead of re-deriving them */
} fs_workspace_layout;

/* Size/layout of the workspace for P Gaussians, a W x H image and room for `instance_capacity` instances. */
size_t fs_workspace_bytes(int P, int width, int height, size_t instance_capacity);
int fs_get_workspace_layout(int P, int width, int height, size_t instance_capacity, fs_workspace_layout* out);

/*
 * Forward render.  Argument meaning and order follow Rasterizer::forward (rasterizer.h:35-59).
 *   d_out_color [3][H][W], d_radii [P] (int32) are outputs owned by the caller.
 *   h_info: optional pinned host pointer; receives the frame header asynchronously on `stream`.
 * Empty input (P == 0) leaves out_color untouched like the reference (rasterize_points.cu:81).
 */
int fs_forward(int P, int D, int M, const float* d_background, int width, int height, const float* d_means3D,
               const float* d_shs, const float* d_colors_precomp, const float* d_opacities, const float* d_scales,
               float scale_modifier, const float* d_rotations, const float* d_cov3D_precomp,
               const float* d_viewmatrix, const float* d_projmatrix, const float* d_cam_pos, float tan_fovx,
               float tan_fovy, int prefiltered, float* d_out_color, int* d_radii, void* d_workspace,
               size_t workspace_bytes, size_t instance_capacity, fs_frame_info* h_info, void* stream);

/*
 * Backward.  Argument meaning follows Rasterizer::backward (rasterizer.h:61-85).  The workspace must be the
 * one fs_forward filled for the same inputs.  All dL_* outputs are [P,...] and are fully written
 * (zero for culled Gaussians) -- the caller does not need to zero-fill them.
 *   d_dL_dmean2D [P][3] (z = 0), d_dL_dcolors [P][3], d_dL_dopacity [P], d_dL_dmean3D [P][3],
 *   d_dL_dcov3D [P][6], d_dL_dsh [P][M][3], d_dL_dscale [P][3], d_dL_drot [P][4].
 */
int fs_backward(int P, int D, int M, const float* d_background, int width, int height, const float* d_means3D,
                const float* d_shs, const float* d_colors_precomp, const float* d_scales, float scale_modifier,
                const float* d_rotations, const float* d_cov3D_precomp, const float* d_viewmatrix,
                const float* d_projmatrix, const float* d_cam_pos, float tan_fovx, float tan_fovy,
                const int* d_radii, void* d_workspace, size_t workspace_bytes, size_t instance_capacity,
                const float* d_dL_dpix, float* d_dL_dmean2D, float* d_dL_dopacity, float* d_dL_dcolors,
                float* d_dL_dmean3D, float* d_dL_dcov3D, float* d_dL_dsh, float* d_dL_dscale, float* d_dL_drot,
                void* stream);

/*
 * Optional performance hint for the next fs_forward calls on this thread: the heaviest tile's instance count
 * observed in recent frames (fs_frame_info.max_tile_instances), 0 = unknown.  Only affects which sort kernels
 * are launched, never the result.
 */
void fs_set_tile_hint(uint32_t max_tile_instances);
/* Per calling thread (default on): when h_info is pinned host memory, fs_forward's scan kernel also stores R / overflow
 * straight into it (zero-copy store + system fence) long before the frame ends, for callers that block on R like the
 * reference (rasterizer_impl.cu:281).  Callers that never wait (no-host-sync mode, CUDA-graph replay) switch it off and
 * save the kernel that fence; the full header is still copied to h_info at the end of the frame. */
void fs_set_early_notify(int on);

/* present[i] = (view-space z of means3D[i] > 0.2); uint8 0/1 (rasterizer_impl.cu:54-66). */
int fs_mark_visible(int P, const float* d_means3D, const float* d_viewmatrix, const float* d_projmatrix,
                    uint8_t* d_present, void* stream);

/* Mean squared distance to the 3 nearest neighbours (simple_knn.cu:148-222). */
size_t fs_knn_workspace_bytes(int P);
int fs_knn_mean_dist2(int P, const float* d_points, float* d_mean_dist2, void* d_workspace, size_t workspace_bytes,
                      void* stream);

/*
 * Per-splat pose stage: mesh vertices -> the four activated tensors render() hands to
 * the rasterizer.  Replaces, fused into one kernel per direction, model/fateavatar.py:225-240,253-258,
 * volume_rendering/mesh_compute.py:18-59 (face frame / scale / normal), mesh_sampling.py:171-200 (barycentric
 * position), pytorch3d matrix_to_quaternion + quaternion_multiply, and the GaussianModel activations
 * (gaussian_model.py:105-128).  faces / face_index are int64 as in the reference's buffers.
 *   forward : means3D [N,3], scales [N,3] = exp(_scaling + log(face_scale/canonical)), rotations [N,4] =
 *             normalize(q_face (x) _rotation), opacities [N] = sigmoid(_opacity)
 *   backward: dL/dverts [V,3] (zero-filled here, then accumulated) and dL/d{_scaling,_rotation,_offset,_opacity}
 */
int fs_pose_forward(int N, int V, int F, const float* d_verts, const long long* d_faces,
                    const long long* d_face_index, const float* d_bary, const float* d_face_scale_canonical,
                    const float* d_scaling_raw, const float* d_rotation_raw, const float* d_offset_raw,
                    const float* d_opacity_raw, float shell_len, int resize_scale, float* d_means3D, float* d_scales,
                    float* d_rotations, float* d_opacities, void* stream);
int fs_pose_backward(int N, int V, int F, const float* d_verts, const long long* d_faces,
                     const long long* d_face_index, const float* d_bary, const float* d_face_scale_canonical,
                     const float* d_scaling_raw, const float* d_rotation_raw, const float* d_offset_raw,
                     const float* d_opacity_raw, float shell_len, int resize_scale, const float* d_dL_dmeans3D,
                     const float* d_dL_dscales, const float* d_dL_drotations, const float* d_dL_dopacities,
                     float* d_dL_dverts, float* d_dL_dscaling_raw, float* d_dL_drotation_raw, float* d_dL_doffset_raw,
                     float* d_dL_dopacity_raw, void* stream);

/*
 * FLAME linear blend skinning with personalised blendshape deltas, batch size 1.
 * Replaces flame/FLAME.py:156-204 (forward_with_delta_blendshape) AND, in the same pass, flame/FLAME.py:131-154
 * (forward: the same expression/pose without the deltas, FateAvatar's `verts_orig`, model/fateavatar.py:211-222),
 * i.e. flame/lbs.py:24-100 run twice (no native counterpart upstream: ~40 torch kernels per call).
 *   betas [L]            shape+expression coefficients (FLAME.py:180: zeros(n_shape) ++ expression)
 *   l0                   coefficients [0, l0) are KNOWN to be zero (l0 = n_shape upstream): their columns of
 *                        shapedirs are not read and their gradient rows are written as zeros; pass 0 if unknown
 *   pose [J*3]           axis-angle per joint;   parents_host [J]: HOST array, parents[0] ignored, parents[j] < j
 *   v_template [V,3], shapedirs [V,3,L], posedirs [(J-1)*9, V*3], J_regressor [J,V], lbs_weights [V,J]
 *   delta_vertex [V,3], delta_shapedirs [V,3,L], delta_posedirs [(J-1)*9, V*3]: trainable deltas, each may be NULL
 * Outputs: verts [V,3]; optional verts_orig [V,3] (no deltas), pose_feature [(J-1)*9], transforms [J,4,4]
 * (lbs.py "A", relative rigid transforms) for both paths.  The workspace (fs_flame_workspace_bytes, 256-byte
 * aligned) keeps v_posed, joints and the kinematic chain for fs_flame_backward.
 * Backward: dL/dverts [V,3] -> dL/d{delta_vertex, delta_shapedirs, delta_posedirs} (each optional), including the
 * path through the joint regression and the kinematic chain.  d_dL_dv_shaped / d_dL_dv_posed (optional, [V,3])
 * expose the factors of the two rank-1 gradients (delta_shapedirs grad = dL_dv_shaped (x) betas, delta_posedirs
 * grad = pose_feature (x) dL_dv_posed) for callers that all-reduce 62 KB instead of 26 MB;
 * d_factor_header (optional, [L + (J-1)*9]) receives [betas | pose_feature], the head of the factor record of
 * fs_flame_expand_grads, so a rank's whole record is produced by this one call.
 * fs_flame_backward_coeffs (optional, call it after fs_flame_backward on the same workspace) adds the gradients of
 * the COEFFICIENTS for callers that optimise per-frame tracking (train/base.py:113-151; off for INSTA data):
 * dL/dbetas [L] (entries below l0 are written as 0: the caller declared them constant) and dL/dpose [J*3],
 * through the blendshapes, the joint regression, the pose correctives, the kinematic chain and Rodrigues.
 */
#define FS_FLAME_MAX_JOINTS 8
size_t fs_flame_workspace_bytes(int V);
int fs_flame_forward(int V, int L, int l0, int J, const int* parents_host, const float* d_betas, const float* d_pose,
                     const float* d_v_template, const float* d_delta_vertex, const float* d_shapedirs,
                     const float* d_delta_shapedirs, const float* d_posedirs, const float* d_delta_posedirs,
                     const float* d_J_regressor, const float* d_lbs_weights, float* d_verts, float* d_verts_orig,
                     float* d_pose_feature, float* d_transforms, float* d_transforms_orig, void* d_workspace,
                     size_t workspace_bytes, void* stream);
int fs_flame_backward_coeffs(int V, int L, int l0, int J, const int* parents_host, const float* d_pose,
                             const float* d_shapedirs, const float* d_delta_shapedirs, const float* d_posedirs,
                             const float* d_delta_posedirs, const float* d_J_regressor, void* d_workspace,
                             size_t workspace_bytes, float* d_dL_dbetas, float* d_dL_dpose, void* stream);
int fs_flame_backward(int V, int L, int l0, int J, const int* parents_host, const float* d_betas,
                      const float* d_J_regressor, const float* d_lbs_weights, const float* d_dL_dverts,
                      void* d_workspace, size_t workspace_bytes, float* d_dL_ddelta_vertex,
                      float* d_dL_ddelta_shapedirs, float* d_dL_ddelta_posedirs, float* d_dL_dv_shaped,
                      float* d_dL_dv_posed, float* d_factor_header, void* stream);

/*
 * Data-parallel exchange of the FLAME delta gradients in factored form.  Each rank's
 * dL/d(delta_shapedirs) is the rank-1 product dL_dv_shaped (x) betas (24 MB dense), dL/d(delta_posedirs) is
 * pose_feature (x) dL_dv_posed and dL/d(delta_vertex) is dL_dv_shaped, so ranks all-gather one small record each
 *     [ betas L | pose_feature NP | dL_dv_shaped 3V | dL_dv_posed 3V ]        (floats; ~120 KB at FLAME sizes)
 * and expand the SUM over ranks locally: out = scale * sum_r (record_r's outer products).  d_factors holds N such
 * records `rank_stride` floats apart (the all-gather output).  Any output pointer may be NULL.  N <= 8.
 */
#define FS_FLAME_MAX_RANKS 8
int fs_flame_expand_grads(int N, int V, int L, int l0, int NP, const float* d_factors, size_t rank_stride, float scale,
                          float* d_dL_ddelta_vertex, float* d_dL_ddelta_shapedirs, float* d_dL_ddelta_posedirs,
                          void* stream);

/*
 * Peer-memory all-reduce (sum) of the frame-sharded step's flat gradient bucket, one kernel over
 * NVLink / NVSwitch instead of an NCCL call.  Every rank holds `n` floats (a multiple of 4, 16-byte aligned) at the
 * same offset of a symmetric, peer-mapped allocation (e.g. torch.distributed._symmetric_memory):
 *   d_multicast != NULL : multicast address of that region; the NVLink switch adds the N copies in flight
 *                         (multimem.ld_reduce), so a rank reads n floats once regardless of N;
 *   otherwise           : d_peer_ptrs, a DEVICE array of the N ranks' unicast addresses, summed in rank order
 *                         (bitwise identical result on every rank).
 * `offset` (floats, multiple of 4) is added to every base address: one allocation can hold several buckets.
 * The sum is written to the rank-local d_out (must not alias the symmetric region).  The caller puts a cross-rank
 * barrier before the call (all buckets complete) and keeps every rank from refilling a bucket that others may
 * still be reading (a second barrier, or two buckets used alternately).
 */
int fs_p2p_allreduce(int N, const float* d_multicast, const float* const* d_peer_ptrs, size_t offset, size_t n,
                     float* d_out, void* stream);

/*
 * Two-shot variant for larger N (needs multicast): rank `rank` reduces ITS 1/N slice of the n floats with one
 * in-switch reduction (multimem.ld_reduce on d_multicast_in) and broadcasts the result into every rank's output
 * region with multicast stores (multimem.st on d_multicast_out) -- n/N floats per rank and direction instead of n.
 * Needs a cross-rank barrier before (inputs complete) AND after (all slices have landed) the call.
 */
int fs_p2p_reduce_scatter_bcast(int N, int rank, const float* d_multicast_in, float* d_multicast_out, size_t n,
                                void* stream);

/*
 * Fused exchange of one frame-sharded step (SURVEY 8e + 8f N4): cross-rank barrier + all-reduce of the splat-gradient
 * part + expansion of the N factor records into the dense FLAME delta gradients, ONE kernel over peer memory.  It
 * replaces, per step, [signal-pad barrier, fs_p2p_allreduce | fs_p2p_reduce_scatter_bcast (+ barrier),
 * fs_flame_expand_grads] and -- like them -- stands where upstream would need an NCCL all-reduce (upstream trains on
 * one GPU, train/base.py:54-60).  CUDA-graph capturable: the barrier epoch lives in device memory.
 *
 * Every rank owns one symmetric, peer-mapped allocation with the same layout (float offsets, multiples of 4):
 *   [in_offset, +n_splat)                      this step's splat gradients (summed over ranks)
 *   [in_offset + rec_offset + r*rec_stride, ..) rank r's factor record [betas L | pose_feature NP | dL/dv_shaped 3V |
 *                                              dL/dv_posed 3V]; a rank fills only ITS slot, the others stay zero
 *   [out_offset, +n_splat)                     algo 2: where the summed splat gradients land on every rank
 *   [flags_offset, +fs_p2p_exchange_flag_floats())  zero-initialised once; owned by the library afterwards
 *   [gather_offset, +N*rec_stride)             local scratch: the N records are copied here once per step
 * d_peer_ptrs: DEVICE array of the N ranks' unicast base addresses; d_multicast: multicast base or NULL;
 * d_local_base: this rank's own base.  algo 0 = one-shot, unicast 128-bit peer loads summed in rank order (bitwise
 * identical on all ranks); 1 = one-shot, multimem.ld_reduce through the switch; 2 = two-shot (each rank reduces its
 * 1/N slice in the switch and multimem.st-broadcasts it; follow with fs_p2p_wait before reading out_offset).  For
 * algo 0/1 the sum is written to the rank-local d_out.  The dense delta gradients (any may be NULL; all NULL = no
 * FLAME part) are scale * sum over ranks of the records' outer products, as fs_flame_expand_grads.
 * The caller alternates between two input buckets (in_offset) on consecutive steps; no other synchronisation.
 */
size_t fs_p2p_exchange_flag_floats(void);
int fs_p2p_exchange(int N, int rank, int algo, const float* const* d_peer_ptrs, const float* d_multicast,
                    float* d_local_base, size_t in_offset, size_t n_splat, size_t rec_offset, size_t rec_stride,
                    size_t out_offset, size_t flags_offset, size_t gather_offset, float* d_out, int V, int L, int l0,
                    int NP, float scale,
                    float* d_dL_ddelta_vertex, float* d_dL_ddelta_shapedirs, float* d_dL_ddelta_posedirs, void* stream);
int fs_p2p_wait(int N, float* d_local_base, size_t flags_offset, void* stream);
/* Device-side timing of fs_p2p_exchange since the last reset (synchronises the device): mean ns this rank waited in the
 * barrier for its peers, mean ns from the barrier to the last CTA's exit, number of calls. */
int fs_p2p_exchange_timing(float* d_local_base, size_t flags_offset, double* wait_ns, double* work_ns, int* calls,
                           int reset);

/*
 * Densification statistics (model/fateavatar.py:734-737, gaussian_model.py:418-420), in place:
 *   xyz_gradient_accum[i] += hypot(viewspace_grad[i,0], viewspace_grad[i,1]);  denom[i] += 1   where update_filter[i]
 * viewspace_grad [P,3] is the .grad of the dummy screen-space tensor (fs_backward's dL_dmeans2D), update_filter [P]
 * is torch.bool / uint8 (radii > 0), xyz_gradient_accum and denom are [P,1] float.
 */
int fs_densify_stats(int P, const float* d_viewspace_grad, const uint8_t* d_update_filter,
                     float* d_xyz_gradient_accum, float* d_denom, void* stream);
/* Per-frame camera in closed form (replaces the two CPU inverses, the host round trip and the GPU inverse
 * of volume_rendering/camera_3dgs.py:53-72): d_cam_pose [4,4] row-major as the dataset yields it (camera-to-world rotation
 * in [:3,:3], world-to-camera translation in [:3,3]), d_projection_t = getProjectionMatrix(...)^T [4,4]  ->
 * world_view_transform [4,4], full_proj_transform [4,4], camera_center [3].  One launch, no host involvement. */
int fs_frame_camera(const float* d_cam_pose, const float* d_projection_t, float* d_world_view, float* d_full_proj,
                    float* d_camera_center, void* stream);
/* Fused L1 image loss of the optimise loop (train/loss.py:103-105, rgb_type 'l1'): *d_loss = mean |x - target| and
 * d_grad[i] = sign(x[i] - target[i]) / n in ONE pass (as torch ops: ~9 launches over the image).  Deterministic.
 * d_workspace: fs_l1_loss_workspace_bytes() bytes, zero-initialised once by the caller, reusable across calls. */
size_t fs_l1_loss_workspace_bytes(void);
int fs_l1_loss(size_t n, const float* d_x, const float* d_target, float* d_grad, float* d_loss, void* d_workspace,
               void* stream);
/* Frame-sharded form of the same statistic: this frame's increments, written (not accumulated) -- hypot(grad) and 1
 * where d_radii[i] > 0, else 0 -- so that they can travel in the gradient bucket and be summed over ranks. */
int fs_densify_stats_inc(int P, const float* d_viewspace_grad, const int* d_radii, float* d_accum_inc,
                         float* d_denom_inc, void* stream);

/*
 * Optimiser step and splat-set maintenance of the optimise loop, on the device.
 *
 * fs_adam_step replaces the two torch.optim.Adam instances over 8 parameter groups of train/optim.py:15-35 by ONE
 * launch over up to FS_ADAM_MAX_TENSORS tensors (torch.optim.Adam arithmetic, betas / eps shared, one learning rate and
 * one step counter per tensor).  d_steps: FS_ADAM_MAX_TENSORS + 1 device ints, zero-initialised by the caller; entry k
 * counts the steps tensor k has taken (advanced by the kernel, so a recorded launch stays valid across CUDA-graph
 * replays), the last entry is library scratch.
 *
 * fs_splat_soa names every array that has one row per splat in model/fateavatar.py (five parameters, their Adam
 * moments -- NULL before the first step --, splat sites, densification statistics, flags), each allocated for a
 * CAPACITY >= P rows, so that
 *   fs_splat_append  (= _uv_densify, :610-672): rows [P, P+n) become children of d_parents[j] (parameters copied,
 *                    scaling' = log(exp(scaling) * 0.75), moments 0, same face, barycentrics d_new_bary[j], flag 1) and
 *                    the statistics of all P+n rows restart at 0;
 *   fs_splat_prune   (= _prune_low_opacity_points, :674-713): rows with sigmoid(opacity) < min_opacity are dropped,
 *                    the survivors of `in` are written in order to `out` (another SoA of the same capacity), their
 *                    number to *d_new_P (device);
 *   fs_opacity_reset (= _reset_opacity, :715-732): opacity = inverse_sigmoid(min(sigmoid(opacity), cap)), moments 0
 * change P without reallocating or concatenating anything.
 */
#define FS_ADAM_MAX_TENSORS 8
typedef struct fs_adam_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    size_t n;
    float lr;
} fs_adam_tensor;
int fs_adam_step(int n_tensors, const fs_adam_tensor* h_tensors, int* d_steps, double beta1, double beta2, double eps,
                 void* stream);
typedef struct fs_splat_soa {
    float *opacity, *offset, *color, *rotation, *scaling;  /* [cap,1] [cap,1] [cap,3] [cap,4] [cap,3] */
    float* exp_avg[5];                                       /* same order and shapes; entries may be NULL */
    float* exp_avg_sq[5];
    long long* face_index;                                   /* [cap] */
    float* bary;                                             /* [cap,3] */
    float *accum, *denom;                                    /* xyz_gradient_accum, denom [cap,1] */
    float *max_radii2D, *sample_flag;                        /* [cap], may be NULL */
} fs_splat_soa;
int fs_splat_append(const fs_splat_soa* soa, int P, int n, const long long* d_parents, const float* d_new_bary,
                    void* stream);
size_t fs_splat_prune_workspace_bytes(int P);
int fs_splat_prune(const fs_splat_soa* in, const fs_splat_soa* out, int P, float min_opacity, int* d_new_P,
                   void* d_workspace, size_t workspace_bytes, void* stream);
int fs_opacity_reset(float* d_opacity, float* d_exp_avg, float* d_exp_avg_sq, int P, float cap, void* stream);

/*
 * Optional per-stage device timing for bench.py's roofline figures.  While enabled, every stage launch is
 * bracketed by CUDA events on the launching stream; fs_profile_read waits for them and returns, per stage id
 * (0 preprocess, 1 tile_scan, 2 scatter, 3 tile_sort, 4 big_tile_sort, 5 blend_forward, 6 blend_backward,
 * 7 preprocess_backward, 8 knn, 9 pose_forward, 10 pose_backward, 11 flame_forward, 12 flame_backward, 13 exchange), the summed milliseconds and the number of launches since the last read.
 */
#define FS_NUM_STAGES 14
void fs_profile_enable(int on);
int fs_profile_read(float* total_ms, int* counts, int n);

/* Number of kernels the last fs_forward / fs_backward / fs_knn_mean_dist2 call on this thread launched. */
int fs_last_launch_count(void);
const char* fs_last_error(void);
const char* fs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FATESPLAT_H */
