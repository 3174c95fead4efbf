// host_selftest — drives the reference-named C++ host classes (vkvolume.h) end to end and dumps what they produced, so that
// tests/test_host_classes_gpu.py can compare the bytes with the oracle (not just parse log lines):
//
//   host_selftest <dataset> <outdir> imin imax gmin gmax skipmode blocksize width height clip
//                 cam_tx cam_ty cam_tz cam_qx cam_qy cam_qz cam_qw [second_dataset]
//
// The sequence is the reference application's (src/volume_render.cpp:163-245): Volume::load_from_file ->
// ComputeGradientMap::compute -> VolumeRender::update_transfer_function (TF texture, [count,] occupancy + distance map)
// -> VolumeRenderSubpass::prepare / draw (all volumes of the scene, in order).  Outputs, raw bytes:
//   voxels.u8, gradient.u8, tf.rgba, map<i>.u8 (every distance map of volume 0), frame.rgba, depth.f32, info.txt
#include <cstdlib>
#include <fstream>
#include <string>

#include "vkvolume.h"

using namespace vkvolume;

static void dump(const std::string &path, const void *p, size_t n)
{
	std::ofstream f(path, std::ios::binary);
	f.write(static_cast<const char *>(p), (std::streamsize) n);
	if (!f) throw std::runtime_error("cannot write " + path);
}

int main(int argc, char **argv)
{
	if (argc < 19) {
		fprintf(stderr, "usage: host_selftest <dataset> <outdir> imin imax gmin gmax skipmode blocksize width height clip tx ty tz qx qy qz qw [dataset2]\n");
		return 2;
	}
	try {
		const std::string dataset = argv[1], out = argv[2];
		const float    imin = (float) atof(argv[3]), imax = (float) atof(argv[4]), gmin = (float) atof(argv[5]), gmax = (float) atof(argv[6]);
		const int      skipmode = atoi(argv[7]);
		const uint32_t blocksize = (uint32_t) atoi(argv[8]), width = (uint32_t) atoi(argv[9]), height = (uint32_t) atoi(argv[10]);
		const float    clip = (float) atof(argv[11]);
		Camera camera;
		for (int i = 0; i < 3; ++i) camera.translation[i] = (float) atof(argv[12 + i]);
		for (int i = 0; i < 4; ++i) camera.rotation[i] = (float) atof(argv[15 + i]);
		camera.aspect = (float) width / (float) height;

		RenderContext render_context(0);
		VolumeRender  app(render_context, /*benchmark_mode=*/false);
		app.volume_render_options.skipping_type = (VolumeRenderSubpass::SkippingType) skipmode;
		app.volume_render_options.clip_distance = clip;

		std::vector<std::unique_ptr<Volume>> volumes;
		std::vector<Node>                    nodes(argc > 19 ? 2 : 1);
		for (int v = 0; v < (int) nodes.size(); ++v) {
			auto volume = std::make_unique<Volume>(v == 0 ? "volume" : "volume2");
			volume->options.intensity_min = imin; volume->options.intensity_max = imax;
			volume->options.gradient_min  = gmin; volume->options.gradient_max  = gmax;
			volume->load_from_file(render_context, v == 0 ? dataset : std::string(argv[19]), blocksize);
			if (v == 1) { nodes[v].translation[0] = 8.0f; nodes[v].translation[1] = -3.0f; }        // the second volume sits beside the first
			volume->set_node(nodes[v]);
			app.compute_gradient(*volume);
			app.update_transfer_function(*volume);
			volumes.push_back(std::move(volume));
		}

		std::vector<Volume *> scene;
		for (auto &v : volumes) scene.push_back(v.get());
		VolumeRenderSubpass subpass(render_context, scene, camera, app.volume_render_options, width, height);
		subpass.prepare();
		auto &cmd = app.compute_start();
		subpass.draw(cmd);
		app.compute_submit(cmd);

		Volume    &vol = *volumes[0];
		const auto vi  = vol.get_volume();
		const size_t n = (size_t) vi.extent[0] * vi.extent[1] * vi.extent[2];
		std::vector<uint8_t> buf(n);
		check(vkv_volume_download_voxels(vol.handle(), buf.data(), n));
		dump(out + "/voxels.u8", buf.data(), n);
		check(vkv_volume_download_gradient(vol.handle(), buf.data(), n));
		dump(out + "/gradient.u8", buf.data(), n);
		std::vector<uint8_t> tf(256 * 256 * 4);
		check(vkv_volume_download_transfer_function(vol.handle(), tf.data(), tf.size()));
		dump(out + "/tf.rgba", tf.data(), tf.size());
		const size_t n_maps = vkv_volume_number_of_distance_maps(vol.handle());
		const auto   mi     = vol.get_distance_map(0);
		const size_t m      = (size_t) mi.extent[0] * mi.extent[1] * mi.extent[2];
		std::vector<uint8_t> map(m);
		for (size_t i = 0; i < n_maps; ++i) {
			check(vkv_volume_download_distance_map(vol.handle(), i, map.data(), m));
			dump(out + "/map" + std::to_string(i) + ".u8", map.data(), m);
		}
		const auto frame = subpass.read_framebuffer(cmd);
		dump(out + "/frame.rgba", frame.data(), frame.size());
		std::vector<float> depth((size_t) width * height);
		cudaMemcpy(depth.data(), subpass.get_depth_buffer(), depth.size() * sizeof(float), cudaMemcpyDeviceToHost);
		dump(out + "/depth.f32", depth.data(), depth.size() * sizeof(float));
		const auto  c = subpass.read_sample_counts(cmd);
		std::ofstream info(out + "/info.txt");
		info << vi.extent[0] << " " << vi.extent[1] << " " << vi.extent[2] << "\n"
		     << mi.extent[0] << " " << mi.extent[1] << " " << mi.extent[2] << "\n"
		     << n_maps << "\n"
		     << c.volume_samples << " " << c.distance_samples << " " << c.empty_samples << " " << c.covered_pixels << "\n";
		for (int i = 0; i < 16; ++i) info << vol.get_image_transform()[i] << (i == 15 ? "\n" : " ");
		return 0;
	} catch (const std::exception &e) {
		fprintf(stderr, "host_selftest: %s\n", e.what());
		return 1;
	}
}
