// vrender_b200 — headless stand-in for the reference's `vrender` executable, built on the reference-named C++
// host classes (vkvolume.h) over libvkv.so.
//
// Accepts the reference's flags (src/volume_render.h:46-56, VS/app/plugins/*):
//   --imin= --imax= --gmin= --gmax= --skipmode= --blocksize= --gradient_test --width= --height=
//   --benchmark[=frames] --stop-after-frame= --screenshot-output=<file.png|file.ppm> <dataset>
// and prints the log lines scripts/benchmark.py greps (scripts/benchmark.py:55-60):
//   "Updated gradient map in {}ms", "Occupied voxels: {}% in {}ms", "Updated occupancy/distance map in {}ms",
//   "... (ran {} frames, averaged {} fps)".
// Benchmark mode applies the reference's silent changes (src/volume_render.cpp:177-183,224-234): clip distance 1,
// early ray termination off, NumTextureSamples view, node scale 100/|R*scale|.
// <dataset> is a raw volume with a "<dataset>.header" next to it, or "synth:<kind>:<W>x<H>x<D>" for a generated one.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vkvolume.h"

using namespace vkvolume;

// PNG writer for the screenshot (the reference goes through stb_image_write after forcing alpha to 255,
// VS/framework/common/utils.cpp:141-175): 8-bit RGBA, filter 0 on every row, zlib stream of stored (uncompressed) blocks.
static uint32_t crc32_update(uint32_t crc, const uint8_t *p, size_t n)
{
	static uint32_t table[256];
	if (!table[1])
		for (uint32_t i = 0; i < 256; ++i) {
			uint32_t c = i;
			for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
			table[i] = c;
		}
	for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
	return crc;
}
static void png_chunk(FILE *f, const char type[4], const std::vector<uint8_t> &data)
{
	const uint32_t n      = (uint32_t) data.size();
	const uint8_t  len[4] = {(uint8_t) (n >> 24), (uint8_t) (n >> 16), (uint8_t) (n >> 8), (uint8_t) n};
	fwrite(len, 1, 4, f);
	fwrite(type, 1, 4, f);
	if (n) fwrite(data.data(), 1, n, f);
	uint32_t crc = crc32_update(0xffffffffu, reinterpret_cast<const uint8_t *>(type), 4);
	crc          = crc32_update(crc, data.data(), n) ^ 0xffffffffu;
	const uint8_t c[4] = {(uint8_t) (crc >> 24), (uint8_t) (crc >> 16), (uint8_t) (crc >> 8), (uint8_t) crc};
	fwrite(c, 1, 4, f);
}
static void write_png_rgba(const std::string &path, const uint8_t *rgba, uint32_t width, uint32_t height)
{
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) throw std::runtime_error("cannot open screenshot file");
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	fwrite(sig, 1, 8, f);
	std::vector<uint8_t> ihdr = {(uint8_t) (width >> 24), (uint8_t) (width >> 16), (uint8_t) (width >> 8), (uint8_t) width,
	                             (uint8_t) (height >> 24), (uint8_t) (height >> 16), (uint8_t) (height >> 8), (uint8_t) height,
	                             8, 6, 0, 0, 0};        // 8 bits, colour type 6 (RGBA), deflate, adaptive filtering, no interlace
	png_chunk(f, "IHDR", ihdr);
	std::vector<uint8_t> raw;        // filter byte 0 + row
	raw.reserve((size_t) height * (1 + (size_t) width * 4));
	for (uint32_t y = 0; y < height; ++y) {
		raw.push_back(0);
		raw.insert(raw.end(), rgba + (size_t) y * width * 4, rgba + (size_t) (y + 1) * width * 4);
	}
	std::vector<uint8_t> z = {0x78, 0x01};
	uint32_t             a = 1, b = 0;        // Adler-32
	for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
		const size_t n     = std::min<size_t>(65535, raw.size() - off);
		const bool   final = off + n >= raw.size();
		z.push_back(final ? 1 : 0);
		z.push_back((uint8_t) n); z.push_back((uint8_t) (n >> 8));
		z.push_back((uint8_t) ~n); z.push_back((uint8_t) (~n >> 8));
		z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
		for (size_t i = 0; i < n; ++i) { a = (a + raw[off + i]) % 65521u; b = (b + a) % 65521u; }
		if (final) break;
	}
	const uint32_t adler = (b << 16) | a;
	z.push_back((uint8_t) (adler >> 24)); z.push_back((uint8_t) (adler >> 16)); z.push_back((uint8_t) (adler >> 8)); z.push_back((uint8_t) adler);
	png_chunk(f, "IDAT", z);
	png_chunk(f, "IEND", {});
	fclose(f);
}

static bool flag_value(const char *arg, const char *name, std::string &out)
{
	const size_t n = strlen(name);
	if (strncmp(arg, name, n) != 0) return false;
	if (arg[n] == '=') { out = arg + n + 1; return true; }
	if (arg[n] == 0) { out = ""; return true; }
	return false;
}

int main(int argc, char **argv)
{
	float       imin = 0.1f, imax = 1.0f, gmin = 0.0f, gmax = 0.2f;        // src/volume_render.cpp:67-70
	int         skipmode = 2;
	uint32_t    blocksize = 4, width = 1280, height = 720;
	bool        gradient_test = false, benchmark = false;
	long        frames = 1000, stop_after = -1;
	std::string dataset = "stag_beetle_832x832x494.uint16", screenshot, v;
	for (int i = 1; i < argc; ++i) {
		const char *a = argv[i];
		if (flag_value(a, "--imin", v)) imin = strtof(v.c_str(), nullptr);
		else if (flag_value(a, "--imax", v)) imax = strtof(v.c_str(), nullptr);
		else if (flag_value(a, "--gmin", v)) gmin = strtof(v.c_str(), nullptr);
		else if (flag_value(a, "--gmax", v)) gmax = strtof(v.c_str(), nullptr);
		else if (flag_value(a, "--skipmode", v)) { int s = atoi(v.c_str()); if (s >= 0 && s <= 3) skipmode = s; }
		else if (flag_value(a, "--blocksize", v)) blocksize = (uint32_t) atoi(v.c_str());
		else if (flag_value(a, "--width", v)) width = (uint32_t) atoi(v.c_str());
		else if (flag_value(a, "--height", v)) height = (uint32_t) atoi(v.c_str());
		else if (flag_value(a, "--gradient_test", v)) gradient_test = true;
		else if (flag_value(a, "--benchmark", v)) { benchmark = true; if (!v.empty()) frames = atol(v.c_str()); }
		else if (flag_value(a, "--stop-after-frame", v)) stop_after = atol(v.c_str());
		else if (flag_value(a, "--screenshot-output", v)) screenshot = v;
		else if (flag_value(a, "--headless", v)) {}
		else if (a[0] != '-') dataset = a;
		else { fprintf(stderr, "unknown flag %s\n", a); return 2; }
	}
	if (stop_after >= 0) frames = stop_after;
	try {
		RenderContext render_context(0);
		VolumeRender  app(render_context, benchmark);
		app.volume_render_options.skipping_type = (VolumeRenderSubpass::SkippingType) skipmode;
		if (benchmark) {
			app.volume_render_options.clip_distance         = 1.0f;
			app.volume_render_options.early_ray_termination = false;
			app.volume_render_options.test                  = VolumeRenderSubpass::Test::NumTextureSamples;
		}
		Volume volume(dataset);
		volume.options.intensity_min = imin; volume.options.intensity_max = imax;
		volume.options.gradient_min = gmin; volume.options.gradient_max = gmax;
		volume.options.use_precomputed_gradient = !gradient_test;
		if (dataset.rfind("synth:", 0) == 0) {
			int      kind = 1;
			unsigned w = 256, h = 256, d = 256;
			if (sscanf(dataset.c_str(), "synth:%d:%ux%ux%u", &kind, &w, &h, &d) != 4) { fprintf(stderr, "bad synth spec\n"); return 2; }
			std::vector<uint8_t> zero(1);
			const float          phys[3] = {0.001f * w, 0.001f * h, 0.001f * d};
			vkvm::Mat4           it      = vkvm::scale(phys[0], phys[1], phys[2]);
			float                itf[16];
			vkvm::to_float(it, itf);
			std::vector<uint8_t> tmp((size_t) w * h * d, 0);
			volume.load_from_memory(render_context, tmp.data(), w, h, d, itf, blocksize);
			check(vkv_synth_volume(render_context.get(), kind, 0x5EED0000u + kind, w, h, d, vkv_volume_device_voxels(volume.handle()), nullptr));
			check(vkv_volume_upload_device(volume.handle(), vkv_volume_device_voxels(volume.handle()), nullptr));
		} else {
			volume.load_from_file(render_context, dataset, blocksize);
		}
		app.compute_gradient(volume);
		app.update_transfer_function(volume);

		Node node;
		if (benchmark) {        // scale so each rotated physical axis spans 100 units (src/volume_render.cpp:224-234)
			const float *m = volume.get_image_transform();
			float        s[3];
			for (int c = 0; c < 3; ++c) s[c] = std::sqrt(m[c * 4] * m[c * 4] + m[c * 4 + 1] * m[c * 4 + 1] + m[c * 4 + 2] * m[c * 4 + 2]);
			float rs[3] = {0, 0, 0};        // |R * scale|: rotate the scale vector by the rotation part of the transform
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c) rs[r] += (m[c * 4 + r] / s[c]) * s[c];
			for (int r = 0; r < 3; ++r) node.scale[r] = 100.0f / std::fabs(rs[r]);
		}
		volume.set_node(node);
		Camera camera;
		camera.aspect = (float) width / (float) height;
		VolumeRenderSubpass subpass(render_context, {&volume}, camera, app.volume_render_options, width, height);
		subpass.prepare();
		auto &cmd = app.compute_start();
		const auto t0 = std::chrono::steady_clock::now();
		for (long f = 0; f < frames; ++f) subpass.draw(cmd);
		app.compute_submit(cmd);
		const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		if (benchmark)
			printf("[info] Benchmark for volume_render completed in %g seconds (ran %ld frames, averaged %g fps)\n", elapsed, frames, frames / elapsed);
		const vkv_sample_counts c = subpass.read_sample_counts(cmd);
		printf("[info] samples: volume %llu distance %llu covered pixels %llu over %ld frames\n", (unsigned long long) c.volume_samples,
		       (unsigned long long) c.distance_samples, (unsigned long long) c.covered_pixels, frames);
		if (!screenshot.empty() && screenshot.size() >= 4 && screenshot.compare(screenshot.size() - 4, 4, ".png") == 0) {
			auto fb = subpass.read_framebuffer(cmd);
			for (size_t p = 0; p < (size_t) width * height; ++p) fb[p * 4 + 3] = 255;        // utils.cpp:141-175: transparency removed
			write_png_rgba(screenshot, fb.data(), width, height);
		} else if (!screenshot.empty()) {        // binary PPM, alpha dropped
			auto  fb = subpass.read_framebuffer(cmd);
			FILE *f  = fopen(screenshot.c_str(), "wb");
			if (!f) throw std::runtime_error("cannot open screenshot file");
			fprintf(f, "P6\n%u %u\n255\n", width, height);
			for (size_t p = 0; p < (size_t) width * height; ++p) fwrite(&fb[p * 4], 1, 3, f);
			fclose(f);
		}
	} catch (const std::exception &e) {
		fprintf(stderr, "[error] %s\n", e.what());
		return 1;
	}
	return 0;
}
