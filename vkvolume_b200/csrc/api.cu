// api.cu — the C ABI of include/vkv.h: contexts, volumes, uploads, orchestration.
// Each entry point names the reference interface it replaces in include/vkv.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>

#include "common.cuh"

namespace vkv {

static thread_local char t_error[512] = "";
std::atomic<uint64_t>    g_kernel_launches{0};

void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(t_error, sizeof t_error, fmt, ap);
	va_end(ap);
}

int sync_arrays_from_linear(vkv_volume *vol, bool gradient, cudaStream_t s)
{
	cudaArray_t dst = gradient ? vol->a_G : vol->a_V;
	uint8_t    *src = gradient ? vol->d_G : vol->d_V;
	if (!dst || !src) return VKV_OK;
	cudaMemcpy3DParms p{};
	p.srcPtr   = make_cudaPitchedPtr(src, vol->dim[0], vol->dim[0], vol->dim[1]);
	p.dstArray = dst;
	p.extent   = make_cudaExtent(vol->dim[0], vol->dim[1], vol->dim[2]);
	p.kind     = cudaMemcpyDeviceToDevice;
	VKV_CUDA_CHECK(cudaMemcpy3DAsync(&p, s));
	return VKV_OK;
}

static int make_texture(cudaArray_t arr, cudaTextureObject_t *out)
{
	cudaResourceDesc rd{};
	rd.resType         = cudaResourceTypeArray;
	rd.res.array.array = arr;
	cudaTextureDesc td{};
	// VK_FILTER_LINEAR + CLAMP_TO_EDGE + R8_UNORM (src/volume_component.cpp:139-148)
	td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
	td.filterMode       = cudaFilterModeLinear;
	td.readMode         = cudaReadModeNormalizedFloat;
	td.normalizedCoords = 1;
	VKV_CUDA_CHECK(cudaCreateTextureObject(out, &rd, &td, nullptr));
	return VKV_OK;
}

// Stream model of the boundary: a stream is a command buffer and a vkv_volume's device state (TF texture and masks, colour table,
// tile history, counters, maps, scratch frames) belongs to ONE stream at a time, like the reference's single in-flight command
// buffer (src/volume_render.cpp:292-327).  A call that arrives on another stream than the volume's previous call is ordered after
// everything that stream had been given (one event record + wait): alternating streams serialises, it never races.  Calls on one
// volume from several HOST threads at once are not supported (the reference is single-threaded).
static cudaStream_t ordered_stream(vkv_volume *vol, void *stream)
{
	cudaStream_t s = (cudaStream_t) stream;
	if (vol->last_stream_valid && vol->last_stream != s) {
		if (!vol->order_event) cudaEventCreateWithFlags(&vol->order_event, cudaEventDisableTiming);
		if (vol->order_event && cudaEventRecord(vol->order_event, vol->last_stream) == cudaSuccess) cudaStreamWaitEvent(s, vol->order_event, 0);
		else (void) cudaGetLastError();        // the previous stream no longer exists: its work finished when it was destroyed
	}
	vol->last_stream       = s;
	vol->last_stream_valid = true;
	return s;
}

struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard(int dev)
	{
		cudaGetDevice(&prev);
		if (prev != dev) cudaSetDevice(dev);
		else prev = -1;
	}
	~DeviceGuard()
	{
		if (prev >= 0) cudaSetDevice(prev);
	}
};

}        // namespace vkv

using namespace vkv;

extern "C" {

const char *vkv_last_error(void) { return t_error; }
const char *vkv_version(void) { return "vkvolume_b200 0.1 (sm_100a)"; }
uint64_t    vkv_kernel_launch_count(void) { return g_kernel_launches.load(); }

int vkv_context_create(int device, vkv_context **out)
{
	VKV_REQUIRE(out, VKV_ERR_ARGUMENT, "vkv_context_create: out is NULL");
	*out  = nullptr;
	int n = 0;
	VKV_CUDA_CHECK(cudaGetDeviceCount(&n));
	VKV_REQUIRE(device >= 0 && device < n, VKV_ERR_ARGUMENT, "vkv_context_create: no such CUDA device");
	DeviceGuard guard(device);        // the caller's current device is restored on every return path
	auto *ctx = new (std::nothrow) vkv_context();
	VKV_REQUIRE(ctx, VKV_ERR_NOMEM, "out of host memory");
	ctx->device = device;
	VKV_CUDA_CHECK(cudaGetDeviceProperties(&ctx->prop, device));
	ctx->sm_count = ctx->prop.multiProcessorCount;
	if (ctx->prop.major < 10) {
		set_error("vkv_context_create: device %d is sm_%d%d; this library carries sm_100a code only", device, ctx->prop.major,
		          ctx->prop.minor);
		delete ctx;
		return VKV_ERR_CUDA;
	}
	*out = ctx;
	return VKV_OK;
}

void vkv_context_destroy(vkv_context *ctx) { delete ctx; }
int  vkv_context_device(const vkv_context *ctx) { return ctx ? ctx->device : -1; }
int  vkv_context_sm_count(const vkv_context *ctx) { return ctx ? ctx->sm_count : 0; }

int vkv_stream_synchronize(vkv_context *ctx, void *stream)
{
	VKV_REQUIRE(ctx, VKV_ERR_ARGUMENT, "ctx is NULL");
	VKV_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) stream));
	return VKV_OK;
}

int vkv_volume_create(vkv_context *ctx, uint32_t width, uint32_t height, uint32_t depth, uint32_t block_size,
                      int use_precomputed_gradient, vkv_volume **out)
{
	VKV_REQUIRE(ctx && out, VKV_ERR_ARGUMENT, "vkv_volume_create: NULL argument");
	VKV_REQUIRE(width && height && depth && block_size, VKV_ERR_ARGUMENT, "vkv_volume_create: zero extent or block size");
	*out = nullptr;
	DeviceGuard guard(ctx->device);
	auto *vol = new (std::nothrow) vkv_volume();
	VKV_REQUIRE(vol, VKV_ERR_NOMEM, "out of host memory");
	vol->ctx    = ctx;
	vol->dim[0] = width; vol->dim[1] = height; vol->dim[2] = depth;
	vol->bs_requested = block_size;
	for (int a = 0; a < 3; ++a) {
		vol->dim_b[a] = rnd_up(vol->dim[a], block_size);        // volume_component.cpp:91-93
		vol->bs[a]    = rnd_up(vol->dim[a], vol->dim_b[a]);     // compute_distance_map.cpp:108-113
	}
	vol->N = (size_t) width * height * depth;
	vol->M = (size_t) vol->dim_b[0] * vol->dim_b[1] * vol->dim_b[2];
	if (vol->M >= (1ull << 32)) {        // the ray caster addresses map cells with 32-bit block indices
		set_error("vkv_volume_create: the distance map would have 2^32 or more blocks; use a larger block size");
		delete vol;
		return VKV_ERR_ARGUMENT;
	}
	vol->precomputed_gradient = use_precomputed_gradient != 0;
	auto fail = [&](cudaError_t e, const char *what) {
		set_error("vkv_volume_create: %s failed: %s", what, cudaGetErrorString(e));
		vkv_volume_destroy(vol);
		return VKV_ERR_CUDA;
	};
	cudaError_t e;
	if ((e = cudaMalloc(&vol->d_V, vol->N)) != cudaSuccess) return fail(e, "cudaMalloc(V)");
	cudaChannelFormatDesc fd  = cudaCreateChannelDesc<unsigned char>();
	cudaExtent            ext = make_cudaExtent(width, height, depth);
	if ((e = cudaMalloc3DArray(&vol->a_V, &fd, ext)) != cudaSuccess) return fail(e, "cudaMalloc3DArray(V)");
	if (make_texture(vol->a_V, &vol->t_V)) { vkv_volume_destroy(vol); return VKV_ERR_CUDA; }
	if (vol->precomputed_gradient) {
		if ((e = cudaMalloc(&vol->d_G, vol->N)) != cudaSuccess) return fail(e, "cudaMalloc(G)");
		if ((e = cudaMalloc3DArray(&vol->a_G, &fd, ext, cudaArraySurfaceLoadStore)) != cudaSuccess) return fail(e, "cudaMalloc3DArray(G)");
		if (make_texture(vol->a_G, &vol->t_G)) { vkv_volume_destroy(vol); return VKV_ERR_CUDA; }
		{
			cudaResourceDesc rd{};
			rd.resType         = cudaResourceTypeArray;
			rd.res.array.array = vol->a_G;
			if ((e = cudaCreateSurfaceObject(&vol->s_G, &rd)) != cudaSuccess) return fail(e, "cudaCreateSurfaceObject(G)");
		}
	}
	make_volume_tensor_maps(vol);
	if ((e = cudaMalloc(&vol->d_tf, 256 * 256 * 4)) != cudaSuccess) return fail(e, "cudaMalloc(tf)");
	if ((e = cudaMalloc(&vol->d_mask2, kMaskWords * sizeof(uint2))) != cudaSuccess) return fail(e, "cudaMalloc(mask)");
	if ((e = cudaMalloc(&vol->d_bounds, sizeof(TFBounds))) != cudaSuccess) return fail(e, "cudaMalloc(bounds)");
	if ((e = cudaMalloc(&vol->d_swap, vol->M)) != cudaSuccess) return fail(e, "cudaMalloc(swap)");
	if ((e = cudaMalloc(&vol->d_tmp, vol->M)) != cudaSuccess) return fail(e, "cudaMalloc(tmp)");
	if ((e = cudaMalloc(&vol->d_count, 8 * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc(count)");
	if ((e = cudaMalloc(&vol->d_counts_scratch, sizeof(vkv_sample_counts))) != cudaSuccess) return fail(e, "cudaMalloc(counts)");
	if ((e = cudaMallocHost(&vol->h_count, 8 * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMallocHost");
	*out = vol;
	return VKV_OK;
}

void vkv_volume_destroy(vkv_volume *vol)
{
	if (vol && vol->grp_world) vkv_volume_group_close(vol);
	if (vol) cudaFree(vol->d_grp);
	if (!vol) return;
	DeviceGuard guard(vol->ctx->device);
	if (vol->t_V) cudaDestroyTextureObject(vol->t_V);
	if (vol->t_G) cudaDestroyTextureObject(vol->t_G);
	if (vol->s_G) cudaDestroySurfaceObject(vol->s_G);
	if (vol->a_V) cudaFreeArray(vol->a_V);
	if (vol->a_G) cudaFreeArray(vol->a_G);
	cudaFree(vol->d_V); cudaFree(vol->d_G); cudaFree(vol->d_tf); cudaFree(vol->d_mask2); cudaFree(vol->d_bounds); cudaFree(vol->d_tf_rows);
	for (auto *m : vol->d_maps) cudaFree(m);
	cudaFree(vol->d_swap); cudaFree(vol->d_tmp); cudaFree(vol->d_count); cudaFree(vol->d_counts_scratch);
	cudaFree(vol->d_fb_scratch);
	cudaFree(vol->d_ctab);
	cudaFree(vol->d_tile_cost);
	cudaFree(vol->d_tile_order);
	if (vol->h_tile_promote) cudaFreeHost(vol->h_tile_promote);
	if (vol->copy_stream) {
		cudaStreamDestroy(vol->copy_stream);
		for (auto &e : vol->band_done) cudaEventDestroy(e);
		cudaEventDestroy(vol->copies_done);
	}
	for (auto *p : vol->d_async_fb) cudaFree(p);
	cudaFree(vol->d_async_counts);
	if (vol->async_ready)
		for (int i = 0; i < vkv_volume::kAsyncSlots; ++i) { cudaEventDestroy(vol->async_rendered[i]); cudaEventDestroy(vol->async_copied[i]); }
	if (vol->h_count) cudaFreeHost(vol->h_count);
	if (vol->order_event) cudaEventDestroy(vol->order_event);
	if (vol->side_stream) { cudaStreamSynchronize(vol->side_stream); cudaStreamDestroy(vol->side_stream); }
	if (vol->ev_march) cudaEventDestroy(vol->ev_march);
	if (vol->ev_order) cudaEventDestroy(vol->ev_order);
	cudaFree(vol->d_lq); cudaFree(vol->d_lrays);
	if (vol->h_long_hint) cudaFreeHost(vol->h_long_hint);
	delete vol;
}

int vkv_volume_upload(vkv_volume *vol, const uint8_t *voxels, void *stream)
{
	VKV_REQUIRE(vol && voxels, VKV_ERR_ARGUMENT, "vkv_volume_upload: NULL argument");
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	VKV_CUDA_CHECK(cudaMemcpyAsync(vol->d_V, voxels, vol->N, cudaMemcpyHostToDevice, s));
	vol->has_V = true;
	return sync_arrays_from_linear(vol, false, s);
}

int vkv_volume_upload_device(vkv_volume *vol, const uint8_t *voxels_dev, void *stream)
{
	VKV_REQUIRE(vol && voxels_dev, VKV_ERR_ARGUMENT, "vkv_volume_upload_device: NULL argument");
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	if (voxels_dev != vol->d_V) VKV_CUDA_CHECK(cudaMemcpyAsync(vol->d_V, voxels_dev, vol->N, cudaMemcpyDeviceToDevice, s));
	vol->has_V = true;
	return sync_arrays_from_linear(vol, false, s);
}

int vkv_volume_upload_gradient(vkv_volume *vol, const uint8_t *gradient, void *stream)
{
	VKV_REQUIRE(vol && gradient, VKV_ERR_ARGUMENT, "vkv_volume_upload_gradient: NULL argument");
	VKV_REQUIRE(vol->d_G, VKV_ERR_STATE, "volume was created without a precomputed gradient map");
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	VKV_CUDA_CHECK(cudaMemcpyAsync(vol->d_G, gradient, vol->N, cudaMemcpyDefault, s));
	vol->has_G = true;
	return sync_arrays_from_linear(vol, true, s);
}

int vkv_volume_upload_raw(vkv_volume *vol, const void *raw, size_t raw_bytes, const char *type, const char *endianness,
                          float lo, float hi, void *stream)
{
	VKV_REQUIRE(vol && raw && type && endianness, VKV_ERR_ARGUMENT, "vkv_volume_upload_raw: NULL argument");
	int kind;
	if (!strcmp(type, "uint8_t")) kind = 0;
	else if (!strcmp(type, "int8_t")) kind = 1;
	else if (!strcmp(type, "uint16_t")) kind = 2;
	else if (!strcmp(type, "int16_t")) kind = 3;
	else {
		set_error("unsupported image data type");        // load_volume.cpp:106-109
		return VKV_ERR_IO;
	}
	const size_t expect = vol->N * (kind >= 2 ? 2 : 1);
	VKV_REQUIRE(raw_bytes == expect, VKV_ERR_IO, "File size does not match expected size for the given image format/dimensions");
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s   = ordered_stream(vol, stream);
	void        *tmp = nullptr;
	VKV_CUDA_CHECK(cudaMallocAsync(&tmp, raw_bytes, s));
	VKV_CUDA_CHECK(cudaMemcpyAsync(tmp, raw, raw_bytes, cudaMemcpyHostToDevice, s));
	int rc = launch_normalise(tmp, vol->N, kind, strcmp(endianness, "big") == 0, lo, hi, vol->d_V, s);
	VKV_CUDA_CHECK(cudaFreeAsync(tmp, s));
	if (rc) return rc;
	vol->has_V = true;
	return sync_arrays_from_linear(vol, false, s);
}

int vkv_volume_set_number_of_distance_maps(vkv_volume *vol, size_t n)
{
	VKV_REQUIRE(vol, VKV_ERR_ARGUMENT, "vol is NULL");
	VKV_REQUIRE(n <= 8, VKV_ERR_ARGUMENT, "at most 8 distance maps");
	DeviceGuard guard(vol->ctx->device);
	if (n <= vol->d_maps.size()) return VKV_OK;        // volume_component.cpp:157-160: never shrinks
	while (vol->d_maps.size() < n) {
		uint8_t *m = nullptr;
		VKV_CUDA_CHECK(cudaMalloc(&m, vol->M));
		vol->d_maps.push_back(m);
	}
	return VKV_OK;
}

int vkv_transfer_function_uniform_from_options(const vkv_volume_options *o, vkv_transfer_function_uniform *u)
{
	VKV_REQUIRE(o && u, VKV_ERR_ARGUMENT, "NULL argument");
	u->sampling_factor         = o->sampling_factor;
	u->voxel_alpha_factor      = o->voxel_alpha_factor;
	u->grad_magnitude_modifier = 1.0f;
	u->use_gradient            = o->gradient_max != o->gradient_min ? 1u : 0u;
	u->intensity_min           = o->intensity_min;
	u->intensity_range_inv     = 1.0f / (o->intensity_max - o->intensity_min);
	u->gradient_min            = o->gradient_min;
	u->gradient_range_inv      = 1.0f / (o->gradient_max - o->gradient_min);
	return VKV_OK;
}

int vkv_volume_update_transfer_function_texture(vkv_volume *vol, const vkv_volume_options *opt, void *stream)
{
	VKV_REQUIRE(vol && opt, VKV_ERR_ARGUMENT, "NULL argument");
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	int          rc;
	(void) rc;
	return launch_tf_texture(vol, opt, s);        // texture, masks and bounds in one kernel
}

int vkv_volume_set_transfer_function_texture(vkv_volume *vol, const uint8_t *rgba, void *stream)
{
	VKV_REQUIRE(vol && rgba, VKV_ERR_ARGUMENT, "NULL argument");
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	VKV_CUDA_CHECK(cudaMemcpyAsync(vol->d_tf, rgba, 256 * 256 * 4, cudaMemcpyHostToDevice, s));
	vol->has_tf = true;
	++vol->tf_version;
	return launch_tf_masks(vol, nullptr, s);
}

int vkv_compute_gradient_map(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, void *stream)
{
	VKV_REQUIRE(vol && tfu, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(vol->has_V, VKV_ERR_STATE, "vkv_compute_gradient_map: no voxels uploaded");
	VKV_REQUIRE(vol->d_G, VKV_ERR_STATE, "volume was created without a precomputed gradient map");
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	int          rc;
	// quirk A.8.1: the map is all 1.0 when use_gradient is false at this moment
	if ((rc = launch_gradient(vol, tfu->use_gradient != 0, tfu->grad_magnitude_modifier, s))) return rc;
	return vol->G_array_synced ? VKV_OK : sync_arrays_from_linear(vol, true, s);
}

static int ensure_analytic_mask(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, cudaStream_t s)
{
	if (vol->mask_ana_valid && !memcmp(&vol->mask_tfu, tfu, sizeof *tfu)) return VKV_OK;
	return launch_tf_masks(vol, tfu, s);
}

static int check_gradient_inputs(vkv_volume *vol, const vkv_transfer_function_uniform *tfu)
{
	// volumes created with use_precomputed_gradient = 0 evaluate gradients on the fly (`--gradient_test`) and have no map
	if (tfu->use_gradient && vol->precomputed_gradient)
		VKV_REQUIRE(vol->has_G, VKV_ERR_STATE, "gradient map not computed yet (vkv_compute_gradient_map)");
	return VKV_OK;
}

int vkv_compute_occupied_voxel_count(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, uint64_t *count_out, void *stream)
{
	VKV_REQUIRE(vol && tfu, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(vol->has_V && vol->has_tf, VKV_ERR_STATE, "vkv_compute_occupied_voxel_count: upload voxels and a transfer function first");
	int rc;
	if ((rc = check_gradient_inputs(vol, tfu))) return rc;
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	if ((rc = ensure_analytic_mask(vol, tfu, s))) return rc;
	VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_count, 0, sizeof(unsigned long long), s));
	if ((rc = launch_occupancy(vol, tfu, true, nullptr, 0, vol->dim_b[2], vol->d_count, s))) return rc;
	if (count_out) {
		VKV_CUDA_CHECK(cudaMemcpyAsync(vol->h_count, vol->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
		VKV_CUDA_CHECK(cudaStreamSynchronize(s));
		*count_out = vol->h_count[0];
	}
	return VKV_OK;
}

static int n_maps_for(int skipping_type) { return skipping_type == VKV_SKIP_ANISOTROPIC_DISTANCE ? 8 : 1; }

int vkv_compute_occupancy_slab(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, int skipping_type, uint32_t zb_first,
                               uint32_t zb_count, uint64_t *count_dev, void *stream)
{
	VKV_REQUIRE(vol && tfu, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(skipping_type >= 0 && skipping_type <= 3, VKV_ERR_ARGUMENT, "bad skipping type");
	VKV_REQUIRE(vol->has_V && vol->has_tf, VKV_ERR_STATE, "upload voxels and a transfer function first");
	VKV_REQUIRE(zb_first + zb_count <= vol->dim_b[2], VKV_ERR_ARGUMENT, "slab exceeds the map depth");
	int rc;
	if ((rc = check_gradient_inputs(vol, tfu))) return rc;
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	const int    n = n_maps_for(skipping_type);
	if ((rc = vkv_volume_set_number_of_distance_maps(vol, n))) return rc;
	if (count_dev && (rc = ensure_analytic_mask(vol, tfu, s))) return rc;
	vol->maps_valid_for = -1;
	rc = launch_occupancy(vol, tfu, count_dev != nullptr, vol->d_maps[n - 1], zb_first, zb_count,
	                      reinterpret_cast<unsigned long long *>(count_dev), s);
	if (!rc) vol->occupancy_in_map = n - 1;        // map n-1 holds occupancy (rows of this and earlier slabs) until K3 consumes it
	return rc;
}

int vkv_volume_mark_occupancy_present(vkv_volume *vol, int skipping_type)
{
	VKV_REQUIRE(vol, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(skipping_type >= 0 && skipping_type <= 3, VKV_ERR_ARGUMENT, "bad skipping type");
	VKV_REQUIRE((int) vol->d_maps.size() >= n_maps_for(skipping_type), VKV_ERR_STATE, "the maps of this skipping type are not allocated");
	vol->occupancy_in_map = n_maps_for(skipping_type) - 1;
	vol->maps_valid_for   = -1;
	return VKV_OK;
}

int vkv_compute_distance_from_occupancy(vkv_volume *vol, int skipping_type, void *stream)
{
	VKV_REQUIRE(vol, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(skipping_type >= 0 && skipping_type <= 3, VKV_ERR_ARGUMENT, "bad skipping type");
	VKV_REQUIRE((int) vol->d_maps.size() >= n_maps_for(skipping_type), VKV_ERR_STATE, "occupancy map not computed");
	// K3 consumes the occupancy map in place (map n-1, reference quirk A.8.3): a second call without a fresh occupancy pass
	// would transform a distance map and call the result valid
	VKV_REQUIRE(vol->occupancy_in_map == n_maps_for(skipping_type) - 1, VKV_ERR_STATE,
	            "vkv_compute_distance_from_occupancy: map n-1 does not hold an occupancy map (run vkv_compute_occupancy_slab first)");
	DeviceGuard guard(vol->ctx->device);
	int         rc = launch_distance(vol, skipping_type, ordered_stream(vol, stream));
	if (skipping_type == VKV_SKIP_DISTANCE || skipping_type == VKV_SKIP_ANISOTROPIC_DISTANCE) vol->occupancy_in_map = -1;
	if (!rc) vol->maps_valid_for = skipping_type;
	return rc;
}

int vkv_compute_distance_map(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, int skipping_type, void *stream)
{
	int rc = vkv_compute_occupancy_slab(vol, tfu, skipping_type, 0, vol ? vol->dim_b[2] : 0, nullptr, stream);
	if (rc) return rc;
	return vkv_compute_distance_from_occupancy(vol, skipping_type, stream);
}

int vkv_update_transfer_function(vkv_volume *vol, const vkv_volume_options *opt, int skipping_type, uint64_t *count_out, void *stream)
{
	VKV_REQUIRE(vol && opt, VKV_ERR_ARGUMENT, "NULL argument");
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s = ordered_stream(vol, stream);
	int          rc;
	if ((rc = vkv_volume_update_transfer_function_texture(vol, opt, stream))) return rc;
	vkv_transfer_function_uniform u;
	vkv_transfer_function_uniform_from_options(opt, &u);
	if (count_out) VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_count, 0, sizeof(unsigned long long), s));
	if ((rc = vkv_compute_occupancy_slab(vol, &u, skipping_type, 0, vol->dim_b[2], count_out ? (uint64_t *) vol->d_count : nullptr, stream))) return rc;
	if ((rc = vkv_compute_distance_from_occupancy(vol, skipping_type, stream))) return rc;
	if (count_out) {
		VKV_CUDA_CHECK(cudaMemcpyAsync(vol->h_count, vol->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
		VKV_CUDA_CHECK(cudaStreamSynchronize(s));
		*count_out = vol->h_count[0];
	}
	return VKV_OK;
}

static int check_render(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width, int height,
                        int tile_w, int tile_h, int tile_stride, bool have_depth)
{
	VKV_REQUIRE(width > 0 && height > 0, VKV_ERR_ARGUMENT, "bad framebuffer extent");
	VKV_REQUIRE(tile_w > 0 && tile_h > 0 && tile_w % 16 == 0 && tile_h % 8 == 0, VKV_ERR_ARGUMENT, "tile extent must be a multiple of 16x8");
	VKV_REQUIRE(tile_stride > 0, VKV_ERR_ARGUMENT, "tile stride must be positive");
	VKV_REQUIRE(opt->skipping_type >= 0 && opt->skipping_type <= 3, VKV_ERR_ARGUMENT, "bad skipping type");
	VKV_REQUIRE(opt->test >= 0 && opt->test <= 3, VKV_ERR_ARGUMENT, "bad test mode");
	VKV_REQUIRE(!opt->depth_attachment || (opt->load_framebuffer && have_depth), VKV_ERR_ARGUMENT,
	            "depth_attachment needs load_framebuffer = 1 and a depth buffer holding the scene's depth");
	VKV_REQUIRE(vol->has_V && vol->has_tf, VKV_ERR_STATE, "vkv_render: upload voxels and a transfer function first");
	int rc;
	if ((rc = check_gradient_inputs(vol, tfu))) return rc;
	if (opt->skipping_type != VKV_SKIP_NONE) {
		VKV_REQUIRE((int) vol->d_maps.size() >= n_maps_for(opt->skipping_type), VKV_ERR_STATE,
		            "vkv_render: distance maps for this skipping type were not computed");
		VKV_REQUIRE(vol->maps_valid_for == opt->skipping_type ||
		                (opt->skipping_type == VKV_SKIP_BLOCK && vol->maps_valid_for == VKV_SKIP_NONE),
		            VKV_ERR_STATE, "vkv_render: maps hold a different skipping type; call vkv_compute_distance_map first");
	}
	return VKV_OK;
}

int vkv_render_tiles(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                     const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width, int height, int tile_w,
                     int tile_h, int tile_first, int tile_stride, uint8_t *rgba8_dev, float *depth_dev, vkv_sample_counts *counts_dev,
                     void *stream)
{
	VKV_REQUIRE(vol && cam && ray && tfu && opt && rgba8_dev, VKV_ERR_ARGUMENT, "vkv_render: NULL argument");
	VKV_REQUIRE(tile_first >= 0, VKV_ERR_ARGUMENT, "tile_first must be >= 0");
	int rc;
	if ((rc = check_render(vol, tfu, opt, width, height, tile_w, tile_h, tile_stride, depth_dev != nullptr))) return rc;
	DeviceGuard guard(vol->ctx->device);
	return launch_render(vol, cam, ray, tfu, opt, width, height, tile_w, tile_h, tile_first, tile_stride, -1, rgba8_dev, depth_dev,
	                     counts_dev, ordered_stream(vol, stream));
}

int vkv_render(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray, const vkv_transfer_function_uniform *tfu,
               const vkv_render_options *opt, int width, int height, uint8_t *rgba8_dev, float *depth_dev,
               vkv_sample_counts *counts_dev, void *stream)
{
	return vkv_render_tiles(vol, cam, ray, tfu, opt, width, height, 64, 32, 0, 1, rgba8_dev, depth_dev, counts_dev, stream);
}

int vkv_render_to_host(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                       const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width, int height,
                       uint8_t *rgba8_host, vkv_sample_counts *counts_host, void *stream)
{
	VKV_REQUIRE(vol && cam && ray && tfu && opt && rgba8_host, VKV_ERR_ARGUMENT, "vkv_render_to_host: NULL argument");
	if (opt->load_framebuffer || opt->depth_attachment) return vkv_render_over_host(vol, cam, ray, tfu, opt, width, height, rgba8_host, nullptr, counts_host, stream);
	constexpr int TW = 64, TH = 32, kBands = 1;        // measured on B200: cross-stream band overlap costs more than it hides (scripts/e2e_probe.py)
	int rc;
	if ((rc = check_render(vol, tfu, opt, width, height, TW, TH, 1, false))) return rc;
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s     = ordered_stream(vol, stream);
	const size_t bytes = (size_t) width * height * 4;
	// Page-locked destination (cudaHostAlloc / cudaHostRegister / torch pin_memory): the ray caster's epilogue stores
	// the RGBA8 pixels straight into it over PCIe, so the transfer of finished pixels overlaps the rays still marching
	// and no staging frame or copy follows the kernel.  Pageable destinations take the staged copy below.
	{
		cudaPointerAttributes attr{};
		const char           *knob = getenv("VKV_E2E_ZEROCOPY");
		const bool            want = knob && atoi(knob) != 0;        // opt-in: measured on B200 (scripts/e2e_probe.py) the 32-byte PCIe writes of the 8x4-pixel warp tiles cost more (0.35 ms) than kernel + bulk copy (0.31 ms)
		if (want && cudaPointerGetAttributes(&attr, rgba8_host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
			vkv_sample_counts *cd = counts_host ? vol->d_counts_scratch : nullptr;
			if (counts_host) VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_counts_scratch, 0, sizeof(vkv_sample_counts), s));
			if ((rc = launch_render(vol, cam, ray, tfu, opt, width, height, TW, TH, 0, 1, -1, static_cast<uint8_t *>(attr.devicePointer), nullptr, cd, s)))
				return rc;
			if (counts_host) VKV_CUDA_CHECK(cudaMemcpyAsync(vol->h_count, vol->d_counts_scratch, sizeof(vkv_sample_counts), cudaMemcpyDeviceToHost, s));
			VKV_CUDA_CHECK(cudaStreamSynchronize(s));
			if (counts_host) memcpy(counts_host, vol->h_count, sizeof(vkv_sample_counts));
			return VKV_OK;
		}
		cudaGetLastError();        // a pageable pointer makes cudaPointerGetAttributes fail on old drivers: not an error here
	}
	if (vol->fb_scratch_bytes < bytes) {
		cudaFree(vol->d_fb_scratch);
		vol->d_fb_scratch     = nullptr;
		vol->fb_scratch_bytes = 0;
		VKV_CUDA_CHECK(cudaMalloc(&vol->d_fb_scratch, bytes));
		vol->fb_scratch_bytes = bytes;
	}
	if (!vol->copy_stream) {
		VKV_CUDA_CHECK(cudaStreamCreateWithFlags(&vol->copy_stream, cudaStreamNonBlocking));
		for (auto &e : vol->band_done) VKV_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		VKV_CUDA_CHECK(cudaEventCreateWithFlags(&vol->copies_done, cudaEventDisableTiming));
	}
	vkv_sample_counts *counts_dev = counts_host ? vol->d_counts_scratch : nullptr;
	if (counts_host) VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_counts_scratch, 0, sizeof(vkv_sample_counts), s));
	// The frame is rendered in bands of tile rows; the device-to-host copy of a finished band runs on a second
	// stream while the next band renders, so the PCIe transfer overlaps the ray casting instead of following it.
	const int tiles_x = (width + TW - 1) / TW, tiles_y = (height + TH - 1) / TH;
	int       want    = kBands;
	if (const char *e = getenv("VKV_E2E_BANDS")) want = atoi(e) > 0 && atoi(e) <= 8 ? atoi(e) : kBands;        // tuning knob
	const int bands   = tiles_y < want ? tiles_y : want;
	int       row0    = 0;
	for (int b = 0; b < bands; ++b) {
		const int row1 = (int) ((long long) tiles_y * (b + 1) / bands);
		if ((rc = launch_render(vol, cam, ray, tfu, opt, width, height, TW, TH, row0 * tiles_x, 1, (row1 - row0) * tiles_x,
		                        vol->d_fb_scratch, nullptr, counts_dev, s)))
			return rc;
		VKV_CUDA_CHECK(cudaEventRecord(vol->band_done[b], s));
		VKV_CUDA_CHECK(cudaStreamWaitEvent(vol->copy_stream, vol->band_done[b], 0));
		const size_t y0 = (size_t) row0 * TH, y1 = (size_t) (row1 * TH < height ? row1 * TH : height);
		VKV_CUDA_CHECK(cudaMemcpyAsync(rgba8_host + y0 * width * 4, vol->d_fb_scratch + y0 * width * 4, (y1 - y0) * width * 4,
		                               cudaMemcpyDeviceToHost, vol->copy_stream));
		row0 = row1;
	}
	// counters go through the volume's pinned staging words (the caller's struct is usually pageable memory,
	// and a pageable async copy would serialise against the band copies)
	static_assert(sizeof(vkv_sample_counts) <= 8 * sizeof(unsigned long long), "pinned staging area too small");
	if (counts_host) VKV_CUDA_CHECK(cudaMemcpyAsync(vol->h_count, vol->d_counts_scratch, sizeof(vkv_sample_counts), cudaMemcpyDeviceToHost, s));
	VKV_CUDA_CHECK(cudaEventRecord(vol->copies_done, vol->copy_stream));
	VKV_CUDA_CHECK(cudaStreamWaitEvent(s, vol->copies_done, 0));
	VKV_CUDA_CHECK(cudaStreamSynchronize(s));
	if (counts_host) memcpy(counts_host, vol->h_count, sizeof(vkv_sample_counts));
	return VKV_OK;
}

int vkv_render_to_host_async(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                             const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width, int height,
                             uint8_t *rgba8_host, vkv_sample_counts *counts_host, void *stream)
{
	VKV_REQUIRE(vol && cam && ray && tfu && opt && rgba8_host, VKV_ERR_ARGUMENT, "vkv_render_to_host_async: NULL argument");
	VKV_REQUIRE(!opt->load_framebuffer && !opt->depth_attachment, VKV_ERR_ARGUMENT, "vkv_render_to_host_async renders over the clear colour only");
	{
		// pageable destinations would turn the asynchronous copies into blocking ones (and the caller's buffers may go away): refuse them
		cudaPointerAttributes a{};
		const bool fb_pinned = cudaPointerGetAttributes(&a, rgba8_host) == cudaSuccess && a.type == cudaMemoryTypeHost;
		bool       c_pinned  = true;
		if (counts_host) c_pinned = cudaPointerGetAttributes(&a, counts_host) == cudaSuccess && a.type == cudaMemoryTypeHost;
		cudaGetLastError();
		VKV_REQUIRE(fb_pinned && c_pinned, VKV_ERR_ARGUMENT, "vkv_render_to_host_async: rgba8_host and counts_host must be page-locked host memory");
	}
	int rc;
	if ((rc = check_render(vol, tfu, opt, width, height, 64, 32, 1, false))) return rc;
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s     = ordered_stream(vol, stream);
	const size_t bytes = (size_t) width * height * 4;
	constexpr int kSlots = vkv_volume::kAsyncSlots;
	if (!vol->copy_stream) {
		VKV_CUDA_CHECK(cudaStreamCreateWithFlags(&vol->copy_stream, cudaStreamNonBlocking));
		for (auto &e : vol->band_done) VKV_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		VKV_CUDA_CHECK(cudaEventCreateWithFlags(&vol->copies_done, cudaEventDisableTiming));
	}
	if (!vol->async_ready) {
		for (int i = 0; i < kSlots; ++i) {
			VKV_CUDA_CHECK(cudaEventCreateWithFlags(&vol->async_rendered[i], cudaEventDisableTiming));
			VKV_CUDA_CHECK(cudaEventCreateWithFlags(&vol->async_copied[i], cudaEventDisableTiming));
		}
		VKV_CUDA_CHECK(cudaMalloc(&vol->d_async_counts, kSlots * sizeof(vkv_sample_counts)));
		vol->async_ready = true;
	}
	if (vol->async_fb_bytes < bytes) {
		VKV_CUDA_CHECK(cudaStreamSynchronize(vol->copy_stream));
		for (auto *&p : vol->d_async_fb) { cudaFree(p); p = nullptr; }
		vol->async_fb_bytes = 0;
		for (auto *&p : vol->d_async_fb) VKV_CUDA_CHECK(cudaMalloc(&p, bytes));
		vol->async_fb_bytes = bytes;
		vol->async_seq      = 0;
	}
	const int slot = (int) (vol->async_seq % kSlots);
	// the frame that used this slot kSlots calls ago must have left the device before it is rendered over
	if (vol->async_seq >= (unsigned) kSlots) VKV_CUDA_CHECK(cudaStreamWaitEvent(s, vol->async_copied[slot], 0));
	vkv_sample_counts *counts_dev = counts_host ? vol->d_async_counts + slot : nullptr;
	if (counts_host) VKV_CUDA_CHECK(cudaMemsetAsync(counts_dev, 0, sizeof(vkv_sample_counts), s));
	if ((rc = launch_render(vol, cam, ray, tfu, opt, width, height, 64, 32, 0, 1, -1, vol->d_async_fb[slot], nullptr, counts_dev, s))) return rc;
	VKV_CUDA_CHECK(cudaEventRecord(vol->async_rendered[slot], s));
	VKV_CUDA_CHECK(cudaStreamWaitEvent(vol->copy_stream, vol->async_rendered[slot], 0));
	VKV_CUDA_CHECK(cudaMemcpyAsync(rgba8_host, vol->d_async_fb[slot], bytes, cudaMemcpyDeviceToHost, vol->copy_stream));
	if (counts_host) VKV_CUDA_CHECK(cudaMemcpyAsync(counts_host, counts_dev, sizeof(vkv_sample_counts), cudaMemcpyDeviceToHost, vol->copy_stream));
	VKV_CUDA_CHECK(cudaEventRecord(vol->async_copied[slot], vol->copy_stream));
	++vol->async_seq;
	return VKV_OK;
}

int vkv_render_to_host_wait(vkv_volume *vol, void *stream)
{
	VKV_REQUIRE(vol, VKV_ERR_ARGUMENT, "vol is NULL");
	DeviceGuard guard(vol->ctx->device);
	VKV_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) stream));
	if (vol->copy_stream) VKV_CUDA_CHECK(cudaStreamSynchronize(vol->copy_stream));
	return VKV_OK;
}

int vkv_render_over_host(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                         const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width, int height,
                         uint8_t *rgba8_host, float *depth_host, vkv_sample_counts *counts_host, void *stream)
{
	VKV_REQUIRE(vol && cam && ray && tfu && opt && rgba8_host, VKV_ERR_ARGUMENT, "vkv_render_over_host: NULL argument");
	int rc;
	if ((rc = check_render(vol, tfu, opt, width, height, 64, 32, 1, depth_host != nullptr))) return rc;
	DeviceGuard  guard(vol->ctx->device);
	cudaStream_t s      = ordered_stream(vol, stream);
	const size_t px     = (size_t) width * height;
	const size_t fbytes = px * 4, dbytes = depth_host ? px * sizeof(float) : 0;
	// one scratch allocation: colour, then depth
	if (vol->fb_scratch_bytes < fbytes + dbytes) {
		cudaFree(vol->d_fb_scratch);
		vol->d_fb_scratch     = nullptr;
		vol->fb_scratch_bytes = 0;
		VKV_CUDA_CHECK(cudaMalloc(&vol->d_fb_scratch, fbytes + dbytes));
		vol->fb_scratch_bytes = fbytes + dbytes;
	}
	uint8_t *d_rgba  = vol->d_fb_scratch;
	float   *d_depth = depth_host ? reinterpret_cast<float *>(vol->d_fb_scratch + fbytes) : nullptr;
	if (opt->load_framebuffer) {
		VKV_CUDA_CHECK(cudaMemcpyAsync(d_rgba, rgba8_host, fbytes, cudaMemcpyHostToDevice, s));
		if (d_depth) VKV_CUDA_CHECK(cudaMemcpyAsync(d_depth, depth_host, dbytes, cudaMemcpyHostToDevice, s));
	}
	vkv_sample_counts *counts_dev = counts_host ? vol->d_counts_scratch : nullptr;
	if (counts_host) VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_counts_scratch, 0, sizeof(vkv_sample_counts), s));
	if ((rc = launch_render(vol, cam, ray, tfu, opt, width, height, 64, 32, 0, 1, -1, d_rgba, d_depth, counts_dev, s))) return rc;
	VKV_CUDA_CHECK(cudaMemcpyAsync(rgba8_host, d_rgba, fbytes, cudaMemcpyDeviceToHost, s));
	if (d_depth) VKV_CUDA_CHECK(cudaMemcpyAsync(depth_host, d_depth, dbytes, cudaMemcpyDeviceToHost, s));
	if (counts_host) VKV_CUDA_CHECK(cudaMemcpyAsync(vol->h_count, vol->d_counts_scratch, sizeof(vkv_sample_counts), cudaMemcpyDeviceToHost, s));
	VKV_CUDA_CHECK(cudaStreamSynchronize(s));
	if (counts_host) memcpy(counts_host, vol->h_count, sizeof(vkv_sample_counts));
	return VKV_OK;
}

int vkv_volume_extent(const vkv_volume *vol, uint32_t out[3])
{
	VKV_REQUIRE(vol && out, VKV_ERR_ARGUMENT, "NULL argument");
	memcpy(out, vol->dim, sizeof vol->dim);
	return VKV_OK;
}
int vkv_volume_map_extent(const vkv_volume *vol, uint32_t out[3])
{
	VKV_REQUIRE(vol && out, VKV_ERR_ARGUMENT, "NULL argument");
	memcpy(out, vol->dim_b, sizeof vol->dim_b);
	return VKV_OK;
}
int vkv_volume_block_size(const vkv_volume *vol, uint32_t out[3])
{
	VKV_REQUIRE(vol && out, VKV_ERR_ARGUMENT, "NULL argument");
	memcpy(out, vol->bs, sizeof vol->bs);
	return VKV_OK;
}
size_t   vkv_volume_number_of_distance_maps(const vkv_volume *vol) { return vol ? vol->d_maps.size() : 0; }
uint8_t *vkv_volume_device_voxels(vkv_volume *vol) { return vol ? vol->d_V : nullptr; }
uint8_t *vkv_volume_device_gradient(vkv_volume *vol) { return vol ? vol->d_G : nullptr; }
uint8_t *vkv_volume_device_distance_map(vkv_volume *vol, size_t idx) { return vol && idx < vol->d_maps.size() ? vol->d_maps[idx] : nullptr; }
uint8_t *vkv_volume_device_transfer_function(vkv_volume *vol) { return vol ? vol->d_tf : nullptr; }

static int download(vkv_volume *vol, const uint8_t *src, size_t bytes, uint8_t *out, size_t out_size)
{
	VKV_REQUIRE(vol && out, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(src, VKV_ERR_STATE, "resource does not exist");
	VKV_REQUIRE(out_size >= bytes, VKV_ERR_ARGUMENT, "output buffer too small");
	DeviceGuard guard(vol->ctx->device);
	VKV_CUDA_CHECK(cudaDeviceSynchronize());
	VKV_CUDA_CHECK(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
	return VKV_OK;
}
int vkv_volume_download_voxels(vkv_volume *vol, uint8_t *out, size_t n) { return download(vol, vol ? vol->d_V : nullptr, vol ? vol->N : 0, out, n); }
int vkv_volume_download_gradient(vkv_volume *vol, uint8_t *out, size_t n) { return download(vol, vol ? vol->d_G : nullptr, vol ? vol->N : 0, out, n); }
int vkv_volume_download_gradient_texture(vkv_volume *vol, uint8_t *out, size_t n)
{
	VKV_REQUIRE(vol && out, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(vol->a_G, VKV_ERR_STATE, "resource does not exist");
	VKV_REQUIRE(n >= vol->N, VKV_ERR_ARGUMENT, "output buffer too small");
	DeviceGuard guard(vol->ctx->device);
	VKV_CUDA_CHECK(cudaDeviceSynchronize());
	cudaMemcpy3DParms p{};
	p.srcArray = vol->a_G;
	p.dstPtr   = make_cudaPitchedPtr(out, vol->dim[0], vol->dim[0], vol->dim[1]);
	p.extent   = make_cudaExtent(vol->dim[0], vol->dim[1], vol->dim[2]);
	p.kind     = cudaMemcpyDeviceToHost;
	VKV_CUDA_CHECK(cudaMemcpy3D(&p));
	return VKV_OK;
}
int vkv_volume_download_distance_map(vkv_volume *vol, size_t idx, uint8_t *out, size_t n)
{
	return download(vol, vkv_volume_device_distance_map(vol, idx), vol ? vol->M : 0, out, n);
}
int vkv_volume_download_transfer_function(vkv_volume *vol, uint8_t *out, size_t n) { return download(vol, vol ? vol->d_tf : nullptr, 256 * 256 * 4, out, n); }

// The handle blob is the cudaIpcMemHandle_t of the ALLOCATION that contains dev_ptr followed by dev_ptr's byte offset
// inside it: framework allocators (torch's caching allocator) hand out interior pointers of larger cudaMalloc blocks,
// and cudaIpcOpenMemHandle always maps the block's base.
static std::mutex                 g_ipc_mutex;
static std::map<void *, void *>   g_ipc_bases;        // pointer returned by vkv_ipc_open -> mapped base

int vkv_ipc_export(void *dev_ptr, uint8_t handle_out[VKV_IPC_HANDLE_BYTES])
{
	static_assert(sizeof(cudaIpcMemHandle_t) + sizeof(uint64_t) == VKV_IPC_HANDLE_BYTES, "IPC handle size");
	VKV_REQUIRE(dev_ptr && handle_out, VKV_ERR_ARGUMENT, "NULL argument");
	typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);        // cuMemGetAddressRange
	static range_fn get_range = nullptr;
	if (!get_range) {
		void                           *fn = nullptr;
		cudaDriverEntryPointQueryResult q;
		VKV_CUDA_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
		VKV_REQUIRE(q == cudaDriverEntryPointSuccess && fn, VKV_ERR_CUDA, "cuMemGetAddressRange is not available");
		get_range = reinterpret_cast<range_fn>(fn);
	}
	unsigned long long base = 0;
	size_t             size = 0;
	VKV_REQUIRE(get_range(&base, &size, (unsigned long long) (uintptr_t) dev_ptr) == 0 && base, VKV_ERR_CUDA,
	            "cuMemGetAddressRange failed: not a device allocation");
	cudaIpcMemHandle_t h;
	VKV_CUDA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>((uintptr_t) base)));
	const uint64_t offset = (uint64_t) ((uintptr_t) dev_ptr - (uintptr_t) base);
	memcpy(handle_out, &h, sizeof h);
	memcpy(handle_out + sizeof h, &offset, sizeof offset);
	return VKV_OK;
}
int vkv_ipc_open(const uint8_t handle[VKV_IPC_HANDLE_BYTES], void **dev_ptr_out)
{
	VKV_REQUIRE(handle && dev_ptr_out, VKV_ERR_ARGUMENT, "NULL argument");
	cudaIpcMemHandle_t h;
	uint64_t           offset = 0;
	memcpy(&h, handle, sizeof h);
	memcpy(&offset, handle + sizeof h, sizeof offset);
	void *base = nullptr;
	VKV_CUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
	*dev_ptr_out = static_cast<uint8_t *>(base) + offset;
	std::lock_guard<std::mutex> lock(g_ipc_mutex);
	g_ipc_bases[*dev_ptr_out] = base;
	return VKV_OK;
}
int vkv_ipc_close(void *dev_ptr)
{
	VKV_REQUIRE(dev_ptr, VKV_ERR_ARGUMENT, "NULL argument");
	void *base = nullptr;
	{
		std::lock_guard<std::mutex> lock(g_ipc_mutex);
		auto it = g_ipc_bases.find(dev_ptr);
		VKV_REQUIRE(it != g_ipc_bases.end(), VKV_ERR_ARGUMENT, "pointer was not returned by vkv_ipc_open");
		base = it->second;
		g_ipc_bases.erase(it);
	}
	VKV_CUDA_CHECK(cudaIpcCloseMemHandle(base));
	return VKV_OK;
}

}        // extern "C"
