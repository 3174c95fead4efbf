// common.cuh — internal declarations shared by the libvkv translation units.
//
// The product path: everything below runs on the device as hand-written sm_100a
// kernels.  There is no CPU fallback; a missing/failed CUDA device surfaces as
// VKV_ERR_CUDA from every entry point.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/vkv.h"

namespace vkv {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_kernel_launches;

#define VKV_CUDA_CHECK(expr)                                                                      \
	do {                                                                                          \
		cudaError_t _e = (expr);                                                                  \
		if (_e != cudaSuccess) {                                                                  \
			::vkv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
			return VKV_ERR_CUDA;                                                                  \
		}                                                                                         \
	} while (0)

#define VKV_REQUIRE(cond, code, msg)  \
	do {                              \
		if (!(cond)) {                \
			::vkv::set_error("%s", msg); \
			return code;              \
		}                             \
	} while (0)

// Launch bookkeeping: every kernel launch in the library goes through this so
// vkv_kernel_launch_count() is an honest count.
#define VKV_LAUNCHED()                                   \
	do {                                                 \
		::vkv::g_kernel_launches.fetch_add(1, std::memory_order_relaxed); \
		VKV_CUDA_CHECK(cudaGetLastError());              \
	} while (0)

static inline uint32_t rnd_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// Visibility masks derived from the transfer function (one bit per (gradient, intensity) texel):
//   .x : TF-texture alpha byte > 0            (occupancy, shaders/occupancy_map.comp:63-64)
//   .y : analytic alphaI*alphaG > 0           (voxel count, shaders/occupied_voxel_count.comp:39-43)
// word index = gradient * 8 + (intensity >> 5), bit = intensity & 31.
constexpr int kMaskWords = 256 * 256 / 32;

// What the O(N) pass knows about a visibility mask without looking texels up:
//   [v_lo, v_hi] x [g_lo, g_hi]  conservative byte ranges outside which no texel is visible (v_lo > v_hi: nothing is);
//   exact                        the visible set IS that rectangle, so a SIMD byte-range test classifies voxels exactly;
//   v_sure, g_sure               every texel with v >= v_sure and g >= g_sure is visible (256: no such rectangle).
struct TFRange {
	uint32_t v_lo, v_hi, g_lo, g_hi, exact, v_sure, g_sure, pad;
};
// tex_*: TF-texture mask (occupancy); ana_*: analytic mask (voxel count); *_row255: gradient row 255 only (use_gradient == false)
struct TFBounds {
	TFRange tex_all, ana_all, tex_row255, ana_row255;
};

}        // namespace vkv

struct vkv_context {
	int            device   = 0;
	int            sm_count = 0;
	cudaDeviceProp prop{};
};

// cudaFuncSetAttribute applies to the current device only: a process that drives several devices through the C ABI must opt in
// to the large dynamic shared-memory sizes once per (kernel instantiation, device), not once per process.
struct PerDeviceOnce {
	std::atomic<unsigned long long> mask{0ull};
	bool first(int device)
	{
		const unsigned long long bit = 1ull << (device & 63);
		return (mask.fetch_or(bit) & bit) == 0ull;
	}
};

// multi-GPU group (group.cu): one block per rank, written by the peers through their IPC mappings of it
constexpr int kGroupMax = 8;
struct GroupSignals {
	unsigned           arrived[kGroupMax];        // arrived[p]: last barrier sequence number rank p has signalled to this rank
	unsigned           pad[8];
	unsigned long long count[kGroupMax];          // count[p]: rank p's partial voxel count of the current rebuild
};

struct vkv_volume {
	vkv_context *ctx = nullptr;
	uint32_t     dim[3]{};         // W H D voxels
	uint32_t     dim_b[3]{};       // map extent = ceil(dim / bs_requested)   (volume_component.cpp:91-93)
	uint32_t     bs[3]{};          // effective block size = ceil(dim / dim_b) (compute_distance_map.cpp:108-113)
	uint32_t     bs_requested = 4;
	bool         precomputed_gradient = true;
	size_t       N = 0, M = 0;

	uint8_t *d_V = nullptr;        // linear voxels   (K1, K2 read these)
	uint8_t *d_G = nullptr;        // linear gradient
	cudaArray_t         a_V = nullptr, a_G = nullptr;        // 3D arrays (K4 samples these)
	cudaTextureObject_t t_V = 0, t_G = 0;
	cudaSurfaceObject_t s_G = 0;                     // surface over a_G: the gradient kernel stores into the array directly
	bool                G_array_synced = false;      // set by launch_gradient when the array needs no copy from d_G
	bool     has_V = false, has_G = false;

	uint8_t        *d_tf      = nullptr;        // 256*256 RGBA8
	uint2          *d_mask2   = nullptr;        // kMaskWords x {texture, analytic}
	vkv::TFBounds  *d_bounds  = nullptr;
	void           *d_tf_rows = nullptr;        // per-row mask statistics + ticket (tf_masks_kernel scratch)
	bool            has_tf    = false;
	vkv_transfer_function_uniform mask_tfu{};   // tfu the analytic mask was built for
	bool            mask_ana_valid = false;

	alignas(64) unsigned char tmap_V[128]{}, tmap_G[128]{};        // CUtensorMap over d_V / d_G (TMA-staged occupancy kernel)
	bool                      tmap_ok = false;

	std::vector<uint8_t *> d_maps;        // distance maps (map n-1 doubles as the occupancy map)
	uint8_t               *d_swap = nullptr;
	uint8_t               *d_tmp  = nullptr;        // second scratch map (anisotropic x-pass results)
	int                    maps_valid_for = -1;     // skipping type the maps currently hold
	int                    occupancy_in_map = -1;   // index of the map that holds a fresh occupancy map (K3 consumes it), or -1

	unsigned long long *d_count = nullptr;        // device-side count + render counters
	unsigned long long *h_count = nullptr;        // pinned
	vkv_sample_counts  *d_counts_scratch = nullptr;
	uint8_t            *d_fb_scratch = nullptr;        // framebuffer for vkv_render_to_host
	size_t              fb_scratch_bytes = 0;
	cudaStream_t        copy_stream = nullptr;         // vkv_render_to_host: D2H copies of finished bands overlap the next band
	cudaEvent_t         band_done[8]{};
	cudaEvent_t         copies_done = nullptr;
	// vkv_render_to_host_async: a ring of device frames so that the copy-out of frame k overlaps the ray casting of frame k + 1
	static constexpr int kAsyncSlots = 3;
	uint8_t            *d_async_fb[kAsyncSlots]{};
	vkv_sample_counts  *d_async_counts = nullptr;        // kAsyncSlots entries
	size_t              async_fb_bytes = 0;
	cudaEvent_t         async_rendered[kAsyncSlots]{}, async_copied[kAsyncSlots]{};
	unsigned            async_seq = 0;
	bool                async_ready = false;
	uint64_t            tf_version = 0;                // bumped by every TF-texture write
	void               *d_ctab = nullptr;              // ray caster: 256x256 float4 premultiplied colour table + the key it was built for
	uint64_t            ctab_tf_version = ~0ull;
	float               ctab_sampling = -1.0f, ctab_alpha = -1.0f;
	// ray caster tile scheduling from the previous frame's cost (raycast.cu): per tile of the last launch's tile list the
	// largest loop count of its warps, and the issue order derived from it
	unsigned           *d_tile_cost = nullptr, *d_tile_order = nullptr;
	int                 tile_hist_capacity = 0;
	int                 tile_hist_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};        // width, height, tile_w, tile_h, tile_first, tile_stride, my_tiles, skipping type of the history
	bool                tile_hist_valid = false;
	int                *h_tile_promote = nullptr;        // pinned + mapped: the last ordering pass's decision (1 = long tiles promoted)
	int                 tile_order_holdoff = 0;          // frames to go before the ordering pass is tried again
	bool                tile_order_prepared = false;     // d_tile_order holds (or will hold, once ev_order fires) the order for the next frame of this key
	cudaStream_t        side_stream = nullptr;           // the ordering pass runs here, beside the frame's long-ray pass
	cudaEvent_t         ev_march = nullptr, ev_order = nullptr;
	// multi-GPU group (group.cu): peers' map 0 / xy-intermediate / signal block, mapped through CUDA IPC
	int                 grp_rank = -1, grp_world = 0;
	uint8_t            *grp_map[8]{}, *grp_swap[8]{};
	GroupSignals       *d_grp = nullptr, *grp_sig[8]{};        // this rank's block and the peers' (grp_sig[rank] == d_grp)
	GroupSignals      **d_grp_sig = nullptr;                   // the same pointers, on the device
	uint8_t           **d_grp_ptrs = nullptr;                  // device copy of grp_map[8] followed by grp_swap[8]
	unsigned            grp_seq = 0;
	void               *d_lq = nullptr, *d_lrays = nullptr;        // ray caster: long-ray queue header + records (raycast.cu)
	int                 long_cap = 0;
	int                *h_long_hint = nullptr;        // pinned + mapped: long rays of the last frame that used the hand-over
	int                 long_holdoff = 0;             // frames to go before the hand-over is tried again
	unsigned            long_seq = 0;                 // launch sequence number: tags the queue records of a launch
	// the stream of the volume's previous call (api.cu ordered_stream): a call on another stream is ordered after it
	cudaStream_t        last_stream = nullptr;
	bool                last_stream_valid = false;
	cudaEvent_t         order_event = nullptr;
};

namespace vkv {
// kernels' host-side launchers (one per translation unit)
int launch_tf_texture(vkv_volume *vol, const vkv_volume_options *opt, cudaStream_t s);
int launch_tf_masks(vkv_volume *vol, const vkv_transfer_function_uniform *tfu_or_null, cudaStream_t s);
int launch_gradient(vkv_volume *vol, bool use_gradient, float modifier, cudaStream_t s);
int launch_occupancy(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, bool count, uint8_t *O, uint32_t zb_first, uint32_t zb_count,
                     unsigned long long *count_dev, cudaStream_t s);
int launch_distance(vkv_volume *vol, int skipping_type, cudaStream_t s);
bool distance_shardable(const vkv_volume *vol);
int launch_distance_xy_slab(vkv_volume *vol, uint32_t zb_first, uint32_t zb_count, bool split, cudaStream_t s);
int launch_distance_z_rows(vkv_volume *vol, uint32_t yb_first, uint32_t yb_count, bool split, cudaStream_t s);
int launch_normalise(const void *raw_dev, size_t n, int kind, bool big_endian, float lo, float hi, uint8_t *out,
                     cudaStream_t s);
int launch_render(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                  const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width, int height,
                  int tile_w, int tile_h, int tile_first, int tile_stride, int tile_limit, uint8_t *rgba8, float *depth,
                  vkv_sample_counts *counts, cudaStream_t s);
int sync_arrays_from_linear(vkv_volume *vol, bool gradient, cudaStream_t s);
int make_volume_tensor_maps(vkv_volume *vol);
}        // namespace vkv
