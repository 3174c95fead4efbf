// vkvolume.h — C++ host layer that keeps the reference's component API on top of the C ABI (include/vkv.h).
//
// Same class names, method names, argument meaning and error behaviour as the reference's src/:
//   LoadVolume                (src/load_volume.h:25-47)
//   Volume                    (src/volume_component.h:31-93)
//   ComputeGradientMap        (src/compute_gradient_map.h:28-47)
//   ComputeOccupiedVoxelCount (src/compute_occupied_voxel_count.h:28-50)
//   ComputeDistanceMap        (src/compute_distance_map.h:28-51)
//   VolumeRenderSubpass       (src/volume_render_subpass.h:55-101) with Options / SkippingType / Test
//   VolumeRender              (src/volume_render.cpp:292-327, 392-445): compute_start / compute_submit /
//                             update_transfer_function and its three log lines
// Substitutions: vkb::RenderContext -> vkvolume::RenderContext (a CUDA device), vkb::CommandBuffer ->
// vkvolume::CommandBuffer (a CUDA stream: recording == enqueueing), vkb::BufferAllocation holding a
// TransferFunctionUniform -> the POD itself, Volume::Image -> a device pointer view.  The Vulkan subpass
// plumbing is replaced by a headless render target: VolumeRenderSubpass::draw writes an RGBA8 (sRGB)
// framebuffer in device memory.  Errors: the loader throws std::runtime_error with the reference's
// messages; every other failure throws std::runtime_error carrying vkv_last_error() (the reference
// abort()s on VK_CHECK failures).  Nothing here computes on the CPU: all voxel work happens in libvkv.so.
#pragma once

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vkv.h"
#include "vkv_math.h"

namespace vkvolume {

inline void check(int rc)
{
	if (rc != VKV_OK) throw std::runtime_error(vkv_last_error());
}

using TransferFunctionUniform = vkv_transfer_function_uniform;        // src/transfer_function.h:20-32
using CameraUniform           = vkv_camera_uniform;                   // src/volume_render_subpass.h:32-39
using RayCastUniform          = vkv_ray_cast_uniform;                 // src/volume_render_subpass.h:46-53

// ~ vkb::RenderContext + vkb::Device: one CUDA device.
class RenderContext {
  public:
	explicit RenderContext(int device = 0) { check(vkv_context_create(device, &ctx_)); }
	~RenderContext() { vkv_context_destroy(ctx_); }
	RenderContext(const RenderContext &) = delete;
	RenderContext &operator=(const RenderContext &) = delete;
	vkv_context *get() const { return ctx_; }

  private:
	vkv_context *ctx_ = nullptr;
};

// ~ vkb::CommandBuffer: work recorded into it is enqueued on a CUDA stream, in order.
struct CommandBuffer {
	RenderContext *render_context = nullptr;
	void          *stream         = nullptr;        // cudaStream_t; nullptr = default stream
};

class LoadVolume {
  public:
	struct Extent3D { uint32_t width, height, depth; };
	struct Header {
		Extent3D    extent{};
		float       voxel_size[3]{};
		float       normalisation_range[2]{};
		std::string type;
		std::string endianness;
		float       image_transform[16]{};        // column-major mat4
		float       tf_range[2]{};                // declared but never filled by the reference either
		float       alpha_factor = 0.0f;
	};

	static Header load_header(std::string filename_header)
	{
		vkv_volume_header h{};
		if (vkv_load_header(filename_header.c_str(), &h) != VKV_OK) throw std::runtime_error(vkv_last_error());
		Header out;
		out.extent = {h.extent[0], h.extent[1], h.extent[2]};
		for (int i = 0; i < 3; ++i) out.voxel_size[i] = h.voxel_size[i];
		out.normalisation_range[0] = h.normalisation_range[0];
		out.normalisation_range[1] = h.normalisation_range[1];
		out.type       = h.type;
		out.endianness = h.endianness;
		for (int i = 0; i < 16; ++i) out.image_transform[i] = h.image_transform[i];
		return out;
	}

	static std::vector<uint8_t> load_data(std::string filename_data, const Header &header)
	{
		vkv_volume_header h = to_c(header);
		std::vector<uint8_t> out((size_t) header.extent.width * header.extent.height * header.extent.depth);
		if (vkv_load_data(filename_data.c_str(), &h, out.data(), out.size()) != VKV_OK) throw std::runtime_error(vkv_last_error());
		return out;
	}

	static vkv_volume_header to_c(const Header &header)
	{
		vkv_volume_header h{};
		h.extent[0] = header.extent.width; h.extent[1] = header.extent.height; h.extent[2] = header.extent.depth;
		for (int i = 0; i < 3; ++i) h.voxel_size[i] = header.voxel_size[i];
		h.normalisation_range[0] = header.normalisation_range[0];
		h.normalisation_range[1] = header.normalisation_range[1];
		snprintf(h.type, sizeof h.type, "%s", header.type.c_str());
		snprintf(h.endianness, sizeof h.endianness, "%s", header.endianness.c_str());
		for (int i = 0; i < 16; ++i) h.image_transform[i] = header.image_transform[i];
		return h;
	}
};

// ~ vkb::sg::Node transform of the volume (translation, rotation quaternion xyzw, scale).
struct Node {
	float translation[3] = {0, 0, 0};
	float rotation[4]    = {0, 0, 0, 1};
	float scale[3]       = {100, 100, 100};        // src/volume_render.cpp:237
};

// ~ vkb::sg::PerspectiveCamera on a node; defaults = Sponza "main_camera" (SURVEY A.5).
struct Camera {
	float translation[3] = {-705.01f, 195.20f, -119.93f};
	float rotation[4]    = {-0.004728f, -0.775409f, -0.005807f, 0.631416f};
	float yfov = 1.0f, aspect = 1.0f, znear = 1.0f, zfar = 4000.0f;
};

class Volume {
  public:
	explicit Volume(const std::string &name) : name_(name) {}
	~Volume() { vkv_volume_destroy(vol_); }
	Volume(const Volume &) = delete;
	Volume &operator=(const Volume &) = delete;

	struct Options {        // src/volume_component.h:45-56
		float sampling_factor          = 1.0f;
		float voxel_alpha_factor       = 1.0f;
		bool  use_precomputed_gradient = true;
		float intensity_min = 0.0f, intensity_max = 1.0f, gradient_min = 0.0f, gradient_max = 1.0f;
	} options;

	struct Image {        // a resident device resource (linear copy; the ray caster also holds a cudaArray)
		uint8_t *data = nullptr;
		uint32_t extent[3]{};
	};

	// src/volume_component.cpp:55-153.  Reads "<filename>.header" + "<filename>"; the raw voxels are normalised on
	// the device (fused loader) instead of on the CPU.  Returns true like the reference.
	bool load_from_file(RenderContext &render_context, std::string filename, uint32_t distance_map_block_size = 4)
	{
		auto header = LoadVolume::load_header(filename + ".header");
		create(render_context, header.extent.width, header.extent.height, header.extent.depth, distance_map_block_size);
		set_image_transform(header.image_transform);
		std::vector<uint8_t> raw = read_file(filename, header);
		check(vkv_volume_upload_raw(vol_, raw.data(), raw.size(), header.type.c_str(), header.endianness.c_str(),
		                            header.normalisation_range[0], header.normalisation_range[1], nullptr));
		return true;
	}

	// Same allocation without a file (synthetic data, tests): voxels are W*H*D bytes on the host.
	void load_from_memory(RenderContext &render_context, const uint8_t *voxels, uint32_t w, uint32_t h, uint32_t d,
	                      const float image_transform[16], uint32_t distance_map_block_size = 4)
	{
		create(render_context, w, h, d, distance_map_block_size);
		set_image_transform(image_transform);
		check(vkv_volume_upload(vol_, voxels, nullptr));
	}

	void       set_image_transform(const float mat[16]) { for (int i = 0; i < 16; ++i) image_transform_[i] = mat[i]; }
	float     *get_image_transform() { return image_transform_; }
	void       set_number_of_distance_maps(RenderContext &, size_t n) { check(vkv_volume_set_number_of_distance_maps(vol_, n)); }
	const Image get_volume() const { return image(vkv_volume_device_voxels(vol_), false); }
	const Image get_gradient() const { return image(vkv_volume_device_gradient(vol_), false); }
	const Image get_transfer_function() const
	{
		Image i;
		i.data = vkv_volume_device_transfer_function(vol_);
		i.extent[0] = i.extent[1] = 256; i.extent[2] = 1;
		return i;
	}
	const Image get_distance_map(size_t idx = 0) const
	{
		uint8_t *p = vkv_volume_device_distance_map(vol_, idx);
		if (!p) throw std::out_of_range("distance map index");        // std::vector::at in the reference
		return image(p, true);
	}

	TransferFunctionUniform get_transfer_function_uniform()        // src/volume_component.cpp:226-240
	{
		TransferFunctionUniform u{};
		vkv_volume_options      o = c_options();
		check(vkv_transfer_function_uniform_from_options(&o, &u));
		return u;
	}
	void update_transfer_function_texture(CommandBuffer &command_buffer)        // src/volume_component.cpp:242-278
	{
		vkv_volume_options o = c_options();
		check(vkv_volume_update_transfer_function_texture(vol_, &o, command_buffer.stream));
	}

	void  set_node(Node &node) { node_ = &node; }
	Node *get_node() const { return node_; }
	vkv_volume        *handle() const { return vol_; }
	const std::string &get_name() const { return name_; }
	vkv_volume_options c_options() const
	{
		return vkv_volume_options{options.sampling_factor, options.voxel_alpha_factor, options.use_precomputed_gradient ? 1 : 0,
		                          options.intensity_min, options.intensity_max, options.gradient_min, options.gradient_max};
	}

  private:
	void create(RenderContext &rc, uint32_t w, uint32_t h, uint32_t d, uint32_t bs)
	{
		vkv_volume_destroy(vol_);
		vol_ = nullptr;
		check(vkv_volume_create(rc.get(), w, h, d, bs, options.use_precomputed_gradient ? 1 : 0, &vol_));
	}
	Image image(uint8_t *p, bool map) const
	{
		Image i;
		i.data = p;
		check(map ? vkv_volume_map_extent(vol_, i.extent) : vkv_volume_extent(vol_, i.extent));
		return i;
	}
	static std::vector<uint8_t> read_file(const std::string &filename, const LoadVolume::Header &header)
	{
		const size_t bpv = (header.type == "uint16_t" || header.type == "int16_t") ? 2 : 1;
		if (header.type != "uint8_t" && header.type != "int8_t" && header.type != "uint16_t" && header.type != "int16_t")
			throw std::runtime_error("unsupported image data type");
		const size_t n = (size_t) header.extent.width * header.extent.height * header.extent.depth * bpv;
		FILE *f = fopen(filename.c_str(), "rb");
		if (!f) throw std::runtime_error("Failed to open data file");
		fseek(f, 0, SEEK_END);
		const size_t actual = (size_t) ftell(f);
		fseek(f, 0, SEEK_SET);
		if (actual != n) {
			fclose(f);
			throw std::runtime_error("File size does not match expected size for the given image format/dimensions");
		}
		std::vector<uint8_t> raw(n);
		const size_t got = fread(raw.data(), 1, n, f);
		fclose(f);
		if (got != n) throw std::runtime_error("File error");
		return raw;
	}

	std::string name_;
	vkv_volume *vol_  = nullptr;
	Node       *node_ = nullptr;
	float       image_transform_[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
};

class ComputeGradientMap {
  public:
	explicit ComputeGradientMap(RenderContext &render_context) : render_context(render_context) {}
	// src/compute_gradient_map.cpp:57-81
	void compute(CommandBuffer &command_buffer, Volume &volume, TransferFunctionUniform &transfer_function_uniform)
	{
		check(vkv_compute_gradient_map(volume.handle(), &transfer_function_uniform, command_buffer.stream));
	}

  private:
	RenderContext &render_context;
};

class ComputeOccupiedVoxelCount {
  public:
	explicit ComputeOccupiedVoxelCount(RenderContext &render_context) : render_context(render_context) {}
	// The reference needs a #workgroups * 512/subgroup_size * 8-byte scratch buffer (268 MB at 1024^3); the fused
	// single-kernel reduction here needs only the 8-byte result.
	struct Buffer { uint64_t result = 0; void *stream = nullptr; RenderContext *rc = nullptr; bool pending = false; };
	Buffer initialise_buffer(RenderContext &device, Volume &) { Buffer b; b.rc = &device; return b; }        // :67-78
	void   compute(CommandBuffer &command_buffer, Volume &volume, Buffer &buffer, TransferFunctionUniform &tfu)        // :80-147
	{
		check(vkv_compute_occupied_voxel_count(volume.handle(), &tfu, &buffer.result, command_buffer.stream));
		buffer.stream  = command_buffer.stream;
		buffer.pending = false;        // the C ABI synchronises when given a host result pointer (== get_result's map())
	}
	uint64_t get_result(Buffer &buffer) const { return buffer.result; }        // :149-156

  private:
	RenderContext &render_context;
};

class VolumeRenderSubpass;

class ComputeDistanceMap {
  public:
	explicit ComputeDistanceMap(RenderContext &render_context) : render_context(render_context) {}
	// src/compute_distance_map.cpp:65-101 — skipping_type is VolumeRenderSubpass::SkippingType as int
	void compute(CommandBuffer &command_buffer, Volume &volume, TransferFunctionUniform &transfer_function_uniform, int skipping_type)
	{
		check(vkv_compute_distance_map(volume.handle(), &transfer_function_uniform, skipping_type, command_buffer.stream));
	}

  private:
	RenderContext &render_context;
};

class VolumeRenderSubpass {
  public:
	enum class SkippingType : int { None = 0, Block = 1, Distance = 2, AnisotropicDistance = 3 };
	enum class Test : int { None = 0, RayEntry = 1, RayExit = 2, NumTextureSamples = 3 };
	struct Options {        // src/volume_render_subpass.h:74-81
		SkippingType skipping_type         = SkippingType::Distance;
		float        clip_distance         = 50.0f;
		bool         early_ray_termination = true;
		bool         depth_attachment      = false;
		Test         test                  = Test::None;
	};

	VolumeRenderSubpass(RenderContext &render_context, std::vector<Volume *> volumes, Camera &camera, Options options, uint32_t width, uint32_t height) :
	    render_context(render_context), camera(camera), volumes(std::move(volumes)), options(options), width(width), height(height)
	{}
	~VolumeRenderSubpass();

	void prepare();                                    // allocates the headless render target
	void draw(CommandBuffer &command_buffer);          // src/volume_render_subpass.cpp:159-294
	uint8_t *get_framebuffer() const { return framebuffer; }        // device pointer, RGBA8, sRGB-encoded RGB, row 0 on top
	float   *get_depth_buffer() const { return depth_buffer; }      // device pointer, float per pixel, reverse-Z (0 = far)
	// The reference draws the volumes after a geometry subpass whose colour and depth are already in the render target
	// (src/volume_render.cpp:344-350).  Headless: the caller fills get_framebuffer() / get_depth_buffer() and says so here;
	// draw() then composites over them (and, with options.depth_attachment, clips the rays at that depth).
	void set_background_loaded(bool loaded) { background_loaded = loaded; }
	std::vector<uint8_t> read_framebuffer(CommandBuffer &command_buffer);
	vkv_sample_counts    read_sample_counts(CommandBuffer &command_buffer);
	void                 reset_sample_counts(CommandBuffer &command_buffer);

  private:
	RenderContext        &render_context;
	Camera               &camera;
	std::vector<Volume *> volumes;
	Options               options;
	uint32_t              width, height;
	uint8_t              *framebuffer  = nullptr;
	float                *depth_buffer = nullptr;
	vkv_sample_counts    *counts       = nullptr;
	bool                  background_loaded = false;
};

// ~ the orchestration half of VolumeRender (src/volume_render.cpp): everything except Vulkan bring-up and the GUI.
class VolumeRender {
  public:
	explicit VolumeRender(RenderContext &rc, bool benchmark_mode = false) :
	    render_context(rc), benchmark_mode(benchmark_mode), compute_distance_map(rc), compute_gradient_map(rc), compute_occupied_voxel_count(rc)
	{}
	CommandBuffer &compute_start()        // :292-299
	{
		command_buffer.render_context = &render_context;
		return command_buffer;
	}
	void compute_submit(CommandBuffer &cmd) { check(vkv_stream_synchronize(render_context.get(), cmd.stream)); }        // :301-327: submit + fence wait

	void compute_gradient(Volume &volume);                     // :203-216
	void update_transfer_function(Volume &volume);             // :392-445

	VolumeRenderSubpass::Options volume_render_options;
	float                        last_occupied_percent = 0.0f, last_update_ms = 0.0f, last_gradient_ms = 0.0f;

  private:
	RenderContext            &render_context;
	bool                      benchmark_mode;
	CommandBuffer             command_buffer;
	ComputeDistanceMap        compute_distance_map;
	ComputeGradientMap        compute_gradient_map;
	ComputeOccupiedVoxelCount compute_occupied_voxel_count;
};

}        // namespace vkvolume

// ---- implementation (header-only) -----------------------------------------------------------------------------------
#include <cuda_runtime_api.h>

namespace vkvolume {

inline VolumeRenderSubpass::~VolumeRenderSubpass()
{
	cudaFree(framebuffer);
	cudaFree(depth_buffer);
	cudaFree(counts);
}

inline void VolumeRenderSubpass::prepare()
{
	if (!framebuffer && cudaMalloc((void **) &framebuffer, (size_t) width * height * 4) != cudaSuccess) throw std::runtime_error("cudaMalloc(framebuffer) failed");
	if (!depth_buffer && cudaMalloc((void **) &depth_buffer, (size_t) width * height * sizeof(float)) != cudaSuccess) throw std::runtime_error("cudaMalloc(depth buffer) failed");
	if (!counts && cudaMalloc((void **) &counts, sizeof(vkv_sample_counts)) != cudaSuccess) throw std::runtime_error("cudaMalloc(counts) failed");
	cudaMemset(counts, 0, sizeof(vkv_sample_counts));
}

inline void VolumeRenderSubpass::draw(CommandBuffer &command_buffer)
{
	if (!framebuffer) prepare();
	if (options.depth_attachment && !background_loaded) throw std::runtime_error("depth_attachment: fill the depth buffer and call set_background_loaded(true) first");
	// the first volume starts from the render-pass clear unless a background was loaded; every later one blends and
	// depth-tests over what is already in the target (the loop of src/volume_render_subpass.cpp:219-293)
	bool load = background_loaded;
	for (auto volume : volumes) {
		TransferFunctionUniform tfu = volume->get_transfer_function_uniform();
		Node                    default_node;
		Node                   *n = volume->get_node() ? volume->get_node() : &default_node;
		vkv_camera_desc         cd{};
		for (int i = 0; i < 3; ++i) { cd.translation[i] = camera.translation[i]; cd.node_translation[i] = n->translation[i]; cd.node_scale[i] = n->scale[i]; }
		for (int i = 0; i < 4; ++i) { cd.rotation[i] = camera.rotation[i]; cd.node_rotation[i] = n->rotation[i]; }
		cd.yfov = camera.yfov; cd.aspect = camera.aspect; cd.znear = camera.znear; cd.zfar = camera.zfar;
		CameraUniform  camera_uniform;
		RayCastUniform ray_cast_uniform;
		check(vkv_make_uniforms(volume->handle(), &cd, volume->get_image_transform(), options.clip_distance, &camera_uniform, &ray_cast_uniform));
		vkv_render_options ro{(int) options.skipping_type, options.clip_distance, options.early_ray_termination ? 1 : 0,
		                      options.depth_attachment ? 1 : 0, (int) options.test, VKV_FILTER_HARDWARE, load ? 1 : 0};
		check(vkv_render(volume->handle(), &camera_uniform, &ray_cast_uniform, &tfu, &ro, (int) width, (int) height, framebuffer, depth_buffer, counts,
		                 command_buffer.stream));
		load = true;
	}
}

inline std::vector<uint8_t> VolumeRenderSubpass::read_framebuffer(CommandBuffer &command_buffer)
{
	std::vector<uint8_t> out((size_t) width * height * 4);
	cudaMemcpyAsync(out.data(), framebuffer, out.size(), cudaMemcpyDeviceToHost, (cudaStream_t) command_buffer.stream);
	cudaStreamSynchronize((cudaStream_t) command_buffer.stream);
	return out;
}

inline vkv_sample_counts VolumeRenderSubpass::read_sample_counts(CommandBuffer &command_buffer)
{
	vkv_sample_counts c{};
	cudaMemcpyAsync(&c, counts, sizeof c, cudaMemcpyDeviceToHost, (cudaStream_t) command_buffer.stream);
	cudaStreamSynchronize((cudaStream_t) command_buffer.stream);
	return c;
}

inline void VolumeRenderSubpass::reset_sample_counts(CommandBuffer &command_buffer)
{
	cudaMemsetAsync(counts, 0, sizeof(vkv_sample_counts), (cudaStream_t) command_buffer.stream);
}

inline void VolumeRender::compute_gradient(Volume &volume)
{
	if (!volume.options.use_precomputed_gradient) return;
	auto       tfu   = volume.get_transfer_function_uniform();
	const auto start = std::chrono::system_clock::now();
	auto      &cmd   = compute_start();
	compute_gradient_map.compute(cmd, volume, tfu);
	compute_submit(cmd);
	const std::chrono::duration<float, std::milli> dur = std::chrono::system_clock::now() - start;
	last_gradient_ms = dur.count();
	printf("[info] Updated gradient map in %gms\n", dur.count());
}

inline void VolumeRender::update_transfer_function(Volume &volume)
{
	auto tfu = volume.get_transfer_function_uniform();
	if (benchmark_mode) {
		auto buffer = compute_occupied_voxel_count.initialise_buffer(render_context, volume);
		const auto start = std::chrono::system_clock::now();
		{
			auto &cmd = compute_start();
			volume.update_transfer_function_texture(cmd);
			compute_occupied_voxel_count.compute(cmd, volume, buffer, tfu);
			compute_submit(cmd);
		}
		const uint64_t n_occupied = compute_occupied_voxel_count.get_result(buffer);
		const Volume::Image vimg  = volume.get_volume();
		const size_t   n_voxels   = (size_t) vimg.extent[0] * vimg.extent[1] * vimg.extent[2];
		last_occupied_percent     = 100.0f * (float) n_occupied / (float) n_voxels;
		const std::chrono::duration<float, std::milli> dur = std::chrono::system_clock::now() - start;
		printf("[info] Occupied voxels: %g%% in %gms\n", last_occupied_percent, dur.count());
		const auto start2 = std::chrono::system_clock::now();
		const int  runs   = 5;
		for (int i = 0; i < runs; ++i) {
			auto &cmd = compute_start();
			compute_distance_map.compute(cmd, volume, tfu, (int) volume_render_options.skipping_type);
			compute_submit(cmd);
		}
		const std::chrono::duration<float, std::milli> dur2 = std::chrono::system_clock::now() - start2;
		last_update_ms = dur2.count() / (float) runs;
		printf("[info] Updated occupancy/distance map in %gms\n", last_update_ms);
	} else {
		{
			auto &cmd = compute_start();
			volume.update_transfer_function_texture(cmd);
			compute_submit(cmd);
		}
		{
			auto &cmd = compute_start();
			compute_distance_map.compute(cmd, volume, tfu, (int) volume_render_options.skipping_type);
			compute_submit(cmd);
		}
	}
}

}        // namespace vkvolume
