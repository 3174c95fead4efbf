// loader.cu — LoadVolume (src/load_volume.{h,cpp}): header parsing on the host, voxel
// normalisation on the device.
//
// The reference converts every voxel to uint8 in a single-threaded std::transform on the CPU
// (src/load_volume.cpp:151-169).  Here the raw file bytes are copied to HBM and one kernel does
// the endian swap + normalisation with the same fp32 expression (IEEE division, truncation), so
// the bytes are identical and the O(N) pass runs at HBM speed.
#include <fstream>
#include <sstream>
#include <string>

#include "../host/vkv_math.h"
#include "common.cuh"

namespace vkv {

// kind: 0 uint8, 1 int8, 2 uint16, 3 int16
template <int KIND>
__global__ void __launch_bounds__(256) normalise_kernel(const uint8_t *__restrict__ raw, size_t n, bool big, float lo, float hi,
                                                       uint8_t *__restrict__ out)
{
	const float range = hi - lo;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
		float v;
		if (KIND == 0) v = (float) raw[i];
		else if (KIND == 1) v = (float) (signed char) raw[i];
		else {
			const unsigned b0 = raw[2 * i], b1 = raw[2 * i + 1];
			const unsigned u  = big ? ((b0 << 8) | b1) : ((b1 << 8) | b0);
			v                 = KIND == 2 ? (float) u : (float) (short) u;
		}
		// static_cast<uint8_t>(255 * max(0.0f, min(1.0f, (float(v) - min) / (max - min))))
		float t = (v - lo) / range;
		t       = fmaxf(0.0f, fminf(1.0f, t));
		out[i]  = (uint8_t) (255.0f * t);
	}
}

int launch_normalise(const void *raw_dev, size_t n, int kind, bool big, float lo, float hi, uint8_t *out, cudaStream_t s)
{
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int      grid = sms * 8;
	const uint8_t *raw  = static_cast<const uint8_t *>(raw_dev);
	switch (kind) {
		case 0: normalise_kernel<0><<<grid, 256, 0, s>>>(raw, n, big, lo, hi, out); break;
		case 1: normalise_kernel<1><<<grid, 256, 0, s>>>(raw, n, big, lo, hi, out); break;
		case 2: normalise_kernel<2><<<grid, 256, 0, s>>>(raw, n, big, lo, hi, out); break;
		default: normalise_kernel<3><<<grid, 256, 0, s>>>(raw, n, big, lo, hi, out); break;
	}
	VKV_LAUNCHED();
	return VKV_OK;
}

}        // namespace vkv

using namespace vkv;

extern "C" {

int vkv_load_header(const char *filename_header, vkv_volume_header *h)
{
	VKV_REQUIRE(filename_header && h, VKV_ERR_ARGUMENT, "NULL argument");
	std::ifstream file(filename_header);
	if (!file.is_open()) {
		set_error("Failed to open header file");        // load_volume.cpp:36-39
		return VKV_ERR_IO;
	}
	memset(h, 0, sizeof *h);
	std::string line;
	// Five lines; each is read with operator>> so anything after the expected tokens
	// (e.g. "# extents") is ignored (load_volume.cpp:54-79).
	std::getline(file, line);
	std::istringstream ss(line);
	ss >> h->extent[0] >> h->extent[1] >> h->extent[2];
	std::getline(file, line);
	ss = std::istringstream(line);
	ss >> h->voxel_size[0] >> h->voxel_size[1] >> h->voxel_size[2];
	std::getline(file, line);
	ss = std::istringstream(line);
	ss >> h->normalisation_range[0] >> h->normalisation_range[1];
	std::getline(file, line);
	ss = std::istringstream(line);
	std::string type, endianness;
	ss >> type >> endianness;
	snprintf(h->type, sizeof h->type, "%s", type.c_str());
	snprintf(h->endianness, sizeof h->endianness, "%s", endianness.c_str());
	std::getline(file, line);
	ss = std::istringstream(line);
	float aa[4] = {0, 0, 0, 0};
	ss >> aa[0] >> aa[1] >> aa[2] >> aa[3];
	// image_transform = rotate(radians(angle), axis) * scale(voxel_size * extent)  (load_volume.cpp:82-83)
	const float phys[3] = {h->voxel_size[0] * (float) h->extent[0], h->voxel_size[1] * (float) h->extent[1],
	                       h->voxel_size[2] * (float) h->extent[2]};
	const float rad     = aa[3] * 0.01745329251994329576923690768489f;
	vkvm::Mat4  m       = vkvm::rotate(rad, aa[0], aa[1], aa[2]) * vkvm::scale(phys[0], phys[1], phys[2]);
	vkvm::to_float(m, h->image_transform);
	return VKV_OK;
}

int vkv_load_data(const char *filename_data, const vkv_volume_header *h, uint8_t *out, size_t out_size)
{
	VKV_REQUIRE(filename_data && h && out, VKV_ERR_ARGUMENT, "NULL argument");
	int               kind;
	const std::string type = h->type;
	if (type == "uint8_t") kind = 0;
	else if (type == "int8_t") kind = 1;
	else if (type == "uint16_t") kind = 2;
	else if (type == "int16_t") kind = 3;
	else {
		set_error("unsupported image data type");
		return VKV_ERR_IO;
	}
	const size_t n         = (size_t) h->extent[0] * h->extent[1] * h->extent[2];
	const size_t file_size = n * (kind >= 2 ? 2 : 1);
	VKV_REQUIRE(out_size >= n, VKV_ERR_ARGUMENT, "output buffer too small");
	std::ifstream file(filename_data, std::ios::binary);
	if (!file.is_open()) {
		set_error("Failed to open data file");
		return VKV_ERR_IO;
	}
	file.seekg(0, std::ios::end);
	if ((size_t) file.tellg() != file_size) {
		set_error("File size does not match expected size for the given image format/dimensions");
		return VKV_ERR_IO;
	}
	file.seekg(0, std::ios::beg);
	uint8_t *h_raw = nullptr;
	VKV_CUDA_CHECK(cudaMallocHost(&h_raw, file_size));
	// chunked read, 100 MB at a time (load_volume.cpp:134-147)
	size_t pos = 0, left = file_size;
	while (left > 0) {
		const size_t chunk = left < (size_t) 100000000 ? left : (size_t) 100000000;
		file.read(reinterpret_cast<char *>(h_raw) + pos, (std::streamsize) chunk);
		if (!file) {
			cudaFreeHost(h_raw);
			set_error("File error");
			return VKV_ERR_IO;
		}
		pos += chunk;
		left -= chunk;
	}
	uint8_t *d_raw = nullptr, *d_out = nullptr;
	auto     cleanup = [&]() { cudaFree(d_raw); cudaFree(d_out); cudaFreeHost(h_raw); };
	cudaError_t e;
	if ((e = cudaMalloc(&d_raw, file_size)) != cudaSuccess || (e = cudaMalloc(&d_out, n)) != cudaSuccess ||
	    (e = cudaMemcpy(d_raw, h_raw, file_size, cudaMemcpyHostToDevice)) != cudaSuccess) {
		set_error("vkv_load_data: %s", cudaGetErrorString(e));
		cleanup();
		return VKV_ERR_CUDA;
	}
	int rc = launch_normalise(d_raw, n, kind, std::string(h->endianness) == "big", h->normalisation_range[0],
	                          h->normalisation_range[1], d_out, nullptr);
	if (!rc && (e = cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost)) != cudaSuccess) {
		set_error("vkv_load_data: %s", cudaGetErrorString(e));
		rc = VKV_ERR_CUDA;
	}
	cleanup();
	return rc;
}

}        // extern "C"
