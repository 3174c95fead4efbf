/*
 * vkv.h — C ABI of the B200-native VkVolume hot path (libvkv.so).
 *
 * This is the drop-in boundary.  The reference (LDeakin/VkVolume) has no FFI for
 * this path: its components are C++ classes that record into a caller-owned
 * vkb::CommandBuffer and the caller submits + waits (src/volume_render.cpp:292-327).
 * Here a CUDA stream plays the command buffer ("void *stream" == cudaStream_t,
 * NULL == the default stream) and vkv_stream_synchronize plays compute_submit.
 * Every entry point below names the reference interface it replaces (file:line,
 * relative to the reference repository root).
 *
 * Stream model: a vkv_volume's device state (TF texture, masks, colour table,
 * tile-scheduling history, counters, maps, scratch frames) belongs to one stream
 * at a time, like the reference's single in-flight command buffer.  A call that
 * arrives on a different stream than the volume's previous call is ordered, by
 * the library, after everything that previous stream had been given (an event
 * record + wait): alternating streams on one volume serialises, it never races.
 * Different volumes are independent.  Calling into ONE volume from several host
 * threads at the same time is not supported (the reference is single-threaded).
 *
 * Rules of the boundary: plain C, opaque handles, plain pointers and sizes, int
 * status codes (0 == VKV_OK) with a thread-local message behind vkv_last_error();
 * nothing throws across it.  There is NO CPU fallback: every compute entry point
 * runs hand-written sm_100a kernels and fails with VKV_ERR_CUDA when no device
 * is usable.
 *
 * Struct layouts are the reference's (std140-compatible):
 *   vkv_transfer_function_uniform == TransferFunctionUniform (src/transfer_function.h:20-32)
 *   vkv_camera_uniform            == CameraUniform           (src/volume_render_subpass.h:32-39)
 *   vkv_ray_cast_uniform          == RayCastUniform          (src/volume_render_subpass.h:46-53)
 *   vkv_render_options            == VolumeRenderSubpass::Options (src/volume_render_subpass.h:74-81)
 *   vkv_volume_options            == Volume::Options         (src/volume_component.h:45-56)
 */
#ifndef VKV_H
#define VKV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VKV_API __attribute__((visibility("default")))
#else
#define VKV_API
#endif

/* ---- status codes ------------------------------------------------------- */
enum {
	VKV_OK            = 0,
	VKV_ERR_ARGUMENT  = 1, /* bad pointer / size / enum                        */
	VKV_ERR_CUDA      = 2, /* CUDA runtime error, message in vkv_last_error()  */
	VKV_ERR_STATE     = 3, /* call order violated (e.g. render before upload)  */
	VKV_ERR_IO        = 4, /* loader: file missing / size mismatch / bad type  */
	VKV_ERR_NOMEM     = 5
};

/* VolumeRenderSubpass::SkippingType (src/volume_render_subpass.h:58-64) */
enum {
	VKV_SKIP_NONE                 = 0,
	VKV_SKIP_BLOCK                = 1,
	VKV_SKIP_DISTANCE             = 2,
	VKV_SKIP_ANISOTROPIC_DISTANCE = 3
};

/* VolumeRenderSubpass::Test (src/volume_render_subpass.h:66-72) */
enum {
	VKV_TEST_NONE                = 0,
	VKV_TEST_RAY_ENTRY           = 1,
	VKV_TEST_RAY_EXIT            = 2,
	VKV_TEST_NUM_TEXTURE_SAMPLES = 3
};

/* Trilinear filter used by the ray caster for V and G.
 * HARDWARE: tex3D on a cudaArray (the north-star path; weights are 1.8 fixed point).
 * EXACT:    8 point fetches + fp32 lerp in the oracle's operation order (parity aid). */
enum {
	VKV_FILTER_HARDWARE = 0,
	VKV_FILTER_EXACT    = 1
};

/* ---- PODs mirrored from the reference ----------------------------------- */
typedef struct vkv_transfer_function_uniform {
	float    sampling_factor;
	float    voxel_alpha_factor;
	float    grad_magnitude_modifier;
	uint32_t use_gradient; /* VkBool32 */
	float    intensity_min;
	float    intensity_range_inv;
	float    gradient_min;
	float    gradient_range_inv;
} vkv_transfer_function_uniform;

typedef struct vkv_camera_uniform {
	float view[16];          /* column-major, as glm::mat4 */
	float proj[16];          /* Y-flipped reverse-Z projection */
	float view_proj_inv[16];
	float model[16];
	float model_inv[16];
} vkv_camera_uniform;

typedef struct vkv_ray_cast_uniform {
	float   plane[4];       /* clip plane, world space   */
	float   plane_tex[4];   /* clip plane, texture space */
	float   cam_pos_tex[4];
	float   block_size[4];  /* effective block size per axis (floats, w = 0) */
	int32_t front_index;
	int32_t _pad[3];
} vkv_ray_cast_uniform;

typedef struct vkv_volume_options {
	float   sampling_factor;          /* default 1 */
	float   voxel_alpha_factor;       /* default 1 */
	int32_t use_precomputed_gradient; /* default 1 */
	float   intensity_min;            /* default 0 */
	float   intensity_max;            /* default 1 */
	float   gradient_min;             /* default 0 */
	float   gradient_max;             /* default 1 */
} vkv_volume_options;

typedef struct vkv_render_options {
	int32_t skipping_type;         /* VKV_SKIP_*, default DISTANCE */
	float   clip_distance;         /* default 50                   */
	int32_t early_ray_termination; /* default 1                    */
	int32_t depth_attachment;      /* DEPTH_ATTACHMENT variant (volume_render.frag:122-136,151-165): the ray is
	                                * discarded behind / shortened at the depth already in `depth_dev`.
	                                * Needs load_framebuffer = 1 and a depth buffer.  default 0 */
	int32_t test;                  /* VKV_TEST_*                   */
	int32_t filter;                /* VKV_FILTER_* (extension; 0 = hardware) */
	int32_t load_framebuffer;      /* extension.  0: the target starts from the render pass clear — colour (0,0,0,1),
	                                * depth 0 (render_pipeline.cpp:38-39) — which is what a single volume over an empty
	                                * scene sees.  1: blend and depth-test (GREATER_OR_EQUAL, depth write on,
	                                * volume_render_subpass.cpp:176-190) over what `rgba8_dev` / `depth_dev` already
	                                * hold: the second and later volumes of VolumeRenderSubpass::draw's loop
	                                * (:219-293), or a volume drawn over rasterised geometry. */
} vkv_render_options;

/* Camera + scene-node description from which the host maths of
 * VolumeRenderSubpass::draw (src/volume_render_subpass.cpp:219-249) builds the
 * two uniforms.  Defaults of the reference: Sponza "main_camera", yfov 1.0,
 * znear 1, zfar 4000, node scale 100 (src/volume_render.cpp:237). */
typedef struct vkv_camera_desc {
	float translation[3];
	float rotation[4];     /* quaternion x, y, z, w */
	float yfov;            /* radians */
	float aspect;          /* width / height */
	float znear, zfar;
	float node_translation[3];
	float node_rotation[4]; /* quaternion x, y, z, w */
	float node_scale[3];
} vkv_camera_desc;

/* LoadVolume::Header (src/load_volume.h:29-39); strings are fixed-size here. */
typedef struct vkv_volume_header {
	uint32_t extent[3];
	float    voxel_size[3];
	float    normalisation_range[2];
	char     type[16];
	char     endianness[16];
	float    image_transform[16]; /* column-major */
} vkv_volume_header;

/* Per-frame sample counters — what the reference's SHOW_NUM_SAMPLES variant
 * sums per pixel (shaders/volume_render.frag:200-204,226,268). */
typedef struct vkv_sample_counts {
	uint64_t volume_samples;
	uint64_t distance_samples;
	uint64_t empty_samples;
	uint64_t covered_pixels;
} vkv_sample_counts;

typedef struct vkv_context vkv_context;
typedef struct vkv_volume  vkv_volume;

/* ---- library / context -------------------------------------------------- */
VKV_API const char *vkv_last_error(void);
VKV_API const char *vkv_version(void);

/* Replaces the Vulkan instance/device bring-up of VolumeRender::prepare
 * (src/volume_render.cpp:138-160): binds to CUDA device `device`. */
VKV_API int  vkv_context_create(int device, vkv_context **out);
VKV_API void vkv_context_destroy(vkv_context *ctx);
VKV_API int  vkv_context_device(const vkv_context *ctx);
VKV_API int  vkv_context_sm_count(const vkv_context *ctx);

/* compute_submit's fence wait (src/volume_render.cpp:301-327). */
VKV_API int vkv_stream_synchronize(vkv_context *ctx, void *stream);

/* ---- loader: LoadVolume (src/load_volume.{h,cpp}) ------------------------ */
/* LoadVolume::load_header (src/load_volume.cpp:33-86). */
VKV_API int vkv_load_header(const char *filename_header, vkv_volume_header *out);
/* LoadVolume::load_data (src/load_volume.cpp:88-172): out must hold W*H*D bytes. */
VKV_API int vkv_load_data(const char *filename_data, const vkv_volume_header *header, uint8_t *out, size_t out_size);

/* ---- Volume (src/volume_component.{h,cpp}) ------------------------------- */
/* Volume::load_from_file's allocation half (src/volume_component.cpp:66-96):
 * V, G (if use_precomputed_gradient), TF texture, distance-map swap image of
 * extent ceil(dim / block_size).  Distance maps are created on demand by
 * vkv_volume_set_number_of_distance_maps / vkv_compute_distance_map. */
VKV_API int  vkv_volume_create(vkv_context *ctx, uint32_t width, uint32_t height, uint32_t depth,
                               uint32_t distance_map_block_size, int use_precomputed_gradient, vkv_volume **out);
VKV_API void vkv_volume_destroy(vkv_volume *vol);

/* The staging upload of load_from_file (src/volume_component.cpp:99-136).
 * `voxels` is a HOST pointer to W*H*D bytes (x fastest). */
VKV_API int vkv_volume_upload(vkv_volume *vol, const uint8_t *voxels, void *stream);
/* Same, from a DEVICE pointer (data already resident in HBM). */
VKV_API int vkv_volume_upload_device(vkv_volume *vol, const uint8_t *voxels_dev, void *stream);
/* Loader fused with the upload: raw file-order voxels of `type` ("uint8_t", "int8_t",
 * "uint16_t", "int16_t") and `endianness` ("big"/"little") on the HOST are copied to
 * the device and normalised there exactly as LoadVolume::load_data_impl does
 * (src/load_volume.cpp:151-169). */
VKV_API int vkv_volume_upload_raw(vkv_volume *vol, const void *raw, size_t raw_bytes, const char *type,
                                  const char *endianness, float norm_lo, float norm_hi, void *stream);

/* Volume::set_number_of_distance_maps (src/volume_component.cpp:155-184). */
VKV_API int vkv_volume_set_number_of_distance_maps(vkv_volume *vol, size_t n);

/* Volume::get_transfer_function_uniform (src/volume_component.cpp:226-240). Pure host maths. */
VKV_API int vkv_transfer_function_uniform_from_options(const vkv_volume_options *opt, vkv_transfer_function_uniform *out);

/* Volume::update_transfer_function_texture (src/volume_component.cpp:242-278):
 * builds the 256x256 RGBA8 texture from the options ON THE DEVICE (bit-identical
 * to the reference's CPU loop) and derives the visibility masks the occupancy and
 * count kernels use. */
VKV_API int vkv_volume_update_transfer_function_texture(vkv_volume *vol, const vkv_volume_options *opt, void *stream);
/* Arbitrary texture from the HOST (what the staging copy would upload): 256*256*4 bytes,
 * texel (x = intensity, y = gradient) at [(y*256 + x)*4]. */
VKV_API int vkv_volume_set_transfer_function_texture(vkv_volume *vol, const uint8_t *rgba, void *stream);

/* ---- compute components --------------------------------------------------- */
/* ComputeGradientMap::compute (src/compute_gradient_map.cpp:57-81) +
 * shaders/gradient_map.comp, shaders/get_gradient_compute.glsl. */
VKV_API int vkv_compute_gradient_map(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, void *stream);

/* ComputeOccupiedVoxelCount::{initialise_buffer, compute, get_result}
 * (src/compute_occupied_voxel_count.cpp:67-156) + shaders/occupied_voxel_count*.comp.
 * The count lands in a device-side uint64; `count_out` (HOST pointer, may be NULL)
 * receives it after an internal stream synchronise (== get_result's map()). */
VKV_API int vkv_compute_occupied_voxel_count(vkv_volume *vol, const vkv_transfer_function_uniform *tfu,
                                             uint64_t *count_out, void *stream);

/* ComputeDistanceMap::compute (src/compute_distance_map.cpp:65-101) +
 * shaders/occupancy_map.comp, distance_map.comp, distance_map_anisotropic.comp.
 * skipping_type selects: NONE/BLOCK -> occupancy only (map 0), DISTANCE -> map 0,
 * ANISOTROPIC_DISTANCE -> maps 0..7 (index = 4[x-] + 2[y-] + [z-]). */
VKV_API int vkv_compute_distance_map(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, int skipping_type,
                                     void *stream);

/* Fused TF-change path of VolumeRender::update_transfer_function
 * (src/volume_render.cpp:392-445): TF texture + occupancy (+ count in the same
 * pass over V,G when count_out != NULL) + distance map, one stream, one sync
 * only if count_out != NULL. */
VKV_API int vkv_update_transfer_function(vkv_volume *vol, const vkv_volume_options *opt, int skipping_type,
                                         uint64_t *count_out, void *stream);

/* ---- ray caster ------------------------------------------------------------ */
/* Host maths of VolumeRenderSubpass::draw (src/volume_render_subpass.cpp:219-249). */
VKV_API int vkv_make_uniforms(const vkv_volume *vol, const vkv_camera_desc *cam, const float image_transform[16],
                              float clip_distance, vkv_camera_uniform *cam_out, vkv_ray_cast_uniform *ray_out);
/* The same host maths from the extents alone (volume extent W,H,D and map extent
 * ceil(dim / block_size), src/volume_render_subpass.cpp:243-249): needs no device
 * and no volume handle. */
VKV_API int vkv_make_uniforms_for_extent(const uint32_t extent[3], const uint32_t map_extent[3], const vkv_camera_desc *cam,
                                         const float image_transform[16], float clip_distance, vkv_camera_uniform *cam_out,
                                         vkv_ray_cast_uniform *ray_out);

/* VolumeRenderSubpass::draw (src/volume_render_subpass.cpp:159-294) + both vertex
 * shaders + shaders/volume_render.frag + the fixed-function blend / sRGB store
 * (SURVEY A.6): writes a width*height RGBA8 (sRGB-encoded RGB) framebuffer, row 0
 * at the top.  `rgba8_dev` is a DEVICE pointer; `depth_dev` (float per pixel,
 * reverse-Z gl_FragDepth) and `counts_dev` (vkv_sample_counts, accumulated into —
 * zero it first) may be NULL.  With opt->load_framebuffer both buffers are read
 * first and the volume is composited over them (see vkv_render_options); a scene
 * with several volumes is one call per volume in scene order, the first with
 * load_framebuffer = 0 (or 1 over a rasterised background), the rest with 1. */
VKV_API int vkv_render(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                       const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width,
                       int height, uint8_t *rgba8_dev, float *depth_dev, vkv_sample_counts *counts_dev,
                       void *stream);

/* Tile-sharded variant for multi-GPU image-space decomposition: the frame is cut
 * into tile_w x tile_h tiles numbered row-major; this call renders tiles
 * t = tile_first, tile_first + tile_stride, ...  and stores them into `rgba8_dev`
 * (full-frame pitch), which may be a peer-mapped pointer to rank 0's framebuffer
 * so the gather is fused into the kernel epilogue. */
VKV_API int vkv_render_tiles(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                             const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width,
                             int height, int tile_w, int tile_h, int tile_first, int tile_stride,
                             uint8_t *rgba8_dev, float *depth_dev, vkv_sample_counts *counts_dev, void *stream);

/* End-to-end call with HOST buffers: uniforms from host structs, framebuffer
 * (and counters) copied back to `rgba8_host` / `counts_host`, stream synchronised. */
VKV_API int vkv_render_to_host(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                               const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width,
                               int height, uint8_t *rgba8_host, vkv_sample_counts *counts_host, void *stream);

/* Pipelined form of vkv_render_to_host for sequences of frames (an orbit, an animation, the benchmark's e2e leg): the call
 * enqueues the ray casting on `stream` and the copy-out of the finished frame (and counters) on an internal copy stream and
 * returns; frame k leaves over PCIe while frame k + 1 is being cast (a ring of three device frames decouples them).
 * `rgba8_host` and `counts_host` must be page-locked and stay valid until vkv_render_to_host_wait(vol, stream) returns, which
 * is also when their contents are defined.  Renders over the clear colour only (no load_framebuffer / depth_attachment). */
VKV_API int vkv_render_to_host_async(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                                     const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width,
                                     int height, uint8_t *rgba8_host, vkv_sample_counts *counts_host, void *stream);
VKV_API int vkv_render_to_host_wait(vkv_volume *vol, void *stream);

/* Host-buffer call that also carries the depth attachment and composites: with
 * opt->load_framebuffer the contents of `rgba8_host` (and `depth_host`, float per
 * pixel, reverse-Z, may be NULL unless opt->depth_attachment) are uploaded, the
 * volume is blended and depth-tested over them as VolumeRenderSubpass::draw does
 * for each volume of the scene in turn (src/volume_render_subpass.cpp:176-190,
 * 219-293) — a rasterised background (src/volume_render.cpp:344-350 feeds the
 * Sponza depth to the DEPTH_ATTACHMENT shader variant) or the volumes drawn before
 * this one — and both buffers are copied back.  Without load_framebuffer it is
 * vkv_render_to_host plus the gl_FragDepth image. */
VKV_API int vkv_render_over_host(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                                 const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width,
                                 int height, uint8_t *rgba8_host, float *depth_host, vkv_sample_counts *counts_host,
                                 void *stream);

/* ---- resource access (Volume::get_* : src/volume_component.h:65-69) --------- */
VKV_API int      vkv_volume_extent(const vkv_volume *vol, uint32_t out[3]);
VKV_API int      vkv_volume_map_extent(const vkv_volume *vol, uint32_t out[3]);
VKV_API int      vkv_volume_block_size(const vkv_volume *vol, uint32_t out[3]); /* effective, SURVEY A.2 */
VKV_API size_t   vkv_volume_number_of_distance_maps(const vkv_volume *vol);
/* Device pointers to the linear copies (NULL if absent). */
VKV_API uint8_t *vkv_volume_device_voxels(vkv_volume *vol);
VKV_API uint8_t *vkv_volume_device_gradient(vkv_volume *vol);
VKV_API uint8_t *vkv_volume_device_distance_map(vkv_volume *vol, size_t idx);
VKV_API uint8_t *vkv_volume_device_transfer_function(vkv_volume *vol);
/* Debug read-backs to HOST memory (synchronous). */
VKV_API int vkv_volume_download_voxels(vkv_volume *vol, uint8_t *out, size_t out_size);
VKV_API int vkv_volume_download_gradient(vkv_volume *vol, uint8_t *out, size_t out_size);
/* The copy of the gradient map the ray caster samples (the texture array behind Volume::get_gradient). */
VKV_API int vkv_volume_download_gradient_texture(vkv_volume *vol, uint8_t *out, size_t out_size);
VKV_API int vkv_volume_download_distance_map(vkv_volume *vol, size_t idx, uint8_t *out, size_t out_size);
VKV_API int vkv_volume_download_transfer_function(vkv_volume *vol, uint8_t *out, size_t out_size);
/* Replace the gradient map / occupancy input (parity tests feed the oracle's G). */
VKV_API int vkv_volume_upload_gradient(vkv_volume *vol, const uint8_t *gradient, void *stream);

/* ---- z-slab sharding of the O(N) pass (multi-GPU) ------------- */
/* Occupancy (+ fused count when count_dev != NULL) for block slices
 * [zb_first, zb_first + zb_count) only, written into map `n_maps - 1` at its
 * place; the caller all-gathers the slab rows of the map between ranks and then
 * calls vkv_compute_distance_from_occupancy on every rank. */
VKV_API int vkv_compute_occupancy_slab(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, int skipping_type,
                                       uint32_t zb_first, uint32_t zb_count, uint64_t *count_dev, void *stream);
/* K3 alone: consumes (overwrites) the occupancy map in map `n_maps - 1`, as the reference does
 * (src/compute_distance_map.cpp:142-175, quirk A.8.3).  Fails with VKV_ERR_STATE when that map does not hold a fresh
 * occupancy map (e.g. on a second call in a row). */
VKV_API int vkv_compute_distance_from_occupancy(vkv_volume *vol, int skipping_type, void *stream);
/* For callers that write an occupancy map (0 = occupied, 255 = empty) into vkv_volume_device_distance_map(vol, n_maps - 1)
 * themselves — e.g. rows gathered from other ranks into a map this rank ran no slab of: declares it present. */
VKV_API int vkv_volume_mark_occupancy_present(vkv_volume *vol, int skipping_type);

/* ---- multi-GPU group: the TF-change rebuild sharded over the GPUs of one node ----
 * One process per GPU, every rank with a full replica of V and G.  vkv_update_transfer_function_sharded is
 * vkv_update_transfer_function (src/volume_render.cpp:392-445) with the work cut up: occupancy (+ count) on
 * the rank's z-slab of blocks; for the isotropic distance map the x and y passes of shaders/distance_map.comp on that slab,
 * an exchange of the xy-intermediate slabs, the z pass on the rank's share of the block rows and an exchange of the result
 * rows; every rank ends up with the whole map (and the whole count).  The exchanges are peer copies over NVLink into the
 * other ranks' buffers and the barriers are signal words in peer memory: nothing leaves the caller's stream, no host round
 * trip, no collective library on the data path.  Set-up: every rank calls vkv_volume_group_export, the application hands
 * all ranks' handle blobs to every rank (any transport: MPI, torch.distributed, a file), every rank calls
 * vkv_volume_group_open with the blobs in rank order.  All ranks must then make the same sequence of sharded calls
 * (it is a collective).  At most 8 ranks.  Volumes below 512 Mi voxels are rebuilt by every rank on its own replica instead
 * (same results, no communication: the barriers and exchanges would cost more than a second GPU saves; VKV_GROUP_ALWAYS=1
 * in the environment forces the sharded path). */
#define VKV_GROUP_HANDLE_BYTES (3 * 72)
VKV_API int vkv_volume_group_export(vkv_volume *vol, uint8_t handles_out[VKV_GROUP_HANDLE_BYTES]);
VKV_API int vkv_volume_group_open(vkv_volume *vol, int rank, int world, const uint8_t *all_handles);
VKV_API int vkv_volume_group_close(vkv_volume *vol);
VKV_API int vkv_update_transfer_function_sharded(vkv_volume *vol, const vkv_volume_options *opt, int skipping_type,
                                                 uint64_t *count_out, void *stream);

/* ---- cross-process peer mapping (CUDA IPC) for the fused tile gather --------
 * The blob is the CUDA IPC handle of the allocation containing dev_ptr plus dev_ptr's byte offset inside it, so
 * interior pointers (sub-allocations of a framework's caching allocator) map to the same bytes in the peer. */
#define VKV_IPC_HANDLE_BYTES 72
VKV_API int vkv_ipc_export(void *dev_ptr, uint8_t handle_out[VKV_IPC_HANDLE_BYTES]);
VKV_API int vkv_ipc_open(const uint8_t handle[VKV_IPC_HANDLE_BYTES], void **dev_ptr_out);
VKV_API int vkv_ipc_close(void *dev_ptr);

/* ---- bench / test utilities (not part of the reference surface) ------------- */
/* Seeded synthetic volumes generated in HBM (kinds: 0 blobs, 1 beetle shell + legs,
 * 2 sparse tubes, 3 hash-noise-modulated blobs).  dev_out: W*H*D bytes. */
VKV_API int vkv_synth_volume(vkv_context *ctx, int kind, uint64_t seed, uint32_t width, uint32_t height,
                             uint32_t depth, uint8_t *dev_out, void *stream);
/* tex3D throughput microbenchmark (the ray caster's roofline denominator):
 * returns filtered fetches per second for a working set of `extent`^3 texels. */
VKV_API int vkv_bench_tex3d(vkv_context *ctx, uint32_t extent, int fetches_per_thread, int coherent,
                            double *fetches_per_second_out);
/* Kernel-launch counter: number of libvkv kernels launched by this process. */
VKV_API uint64_t vkv_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VKV_H */
