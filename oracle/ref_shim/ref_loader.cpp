// ref_loader.cpp — compiles the reference's src/load_volume.cpp AS IS (included from where it lies under
// /root/reference; the path arrives as -DREF_LOAD_VOLUME_CPP) and exposes LoadVolume::load_header/load_data
// through a C interface for tests.  Test infrastructure.
#include <cstdint>
#include <cstring>
#include <exception>
#include REF_LOAD_VOLUME_CPP

static thread_local char g_msg[256];

extern "C" __attribute__((visibility("default"))) const char *ref_loader_error() { return g_msg; }

// out_f: extent(3 as float), voxel_size(3), normalisation(2), image_transform(16 column-major) = 24 floats
extern "C" __attribute__((visibility("default"))) int ref_load_header(const char *fn, float *out_f, char *type16, char *endian16)
{
	try {
		LoadVolume::Header h = LoadVolume::load_header(fn);
		out_f[0] = (float) h.extent.width; out_f[1] = (float) h.extent.height; out_f[2] = (float) h.extent.depth;
		for (int i = 0; i < 3; ++i) out_f[3 + i] = h.voxel_size[i];
		out_f[6] = h.normalisation_range.x; out_f[7] = h.normalisation_range.y;
		for (int c = 0; c < 4; ++c)
			for (int r = 0; r < 4; ++r) out_f[8 + c * 4 + r] = h.image_transform[c][r];
		std::strncpy(type16, h.type.c_str(), 15); type16[15] = 0;
		std::strncpy(endian16, h.endianness.c_str(), 15); endian16[15] = 0;
		return 0;
	} catch (const std::exception &e) {
		std::strncpy(g_msg, e.what(), sizeof g_msg - 1);
		return -1;
	}
}

extern "C" __attribute__((visibility("default"))) int ref_load_data(const char *fn_header, const char *fn_data, uint8_t *out, size_t out_size)
{
	try {
		LoadVolume::Header   h = LoadVolume::load_header(fn_header);
		std::vector<uint8_t> v = LoadVolume::load_data(fn_data, h);
		if (v.size() > out_size) return -2;
		std::memcpy(out, v.data(), v.size());
		return 0;
	} catch (const std::exception &e) {
		std::strncpy(g_msg, e.what(), sizeof g_msg - 1);
		return -1;
	}
}
