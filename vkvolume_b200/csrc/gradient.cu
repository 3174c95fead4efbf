// gradient.cu — K1 gradient-magnitude map.
//
// Replaces shaders/gradient_map.comp + shaders/get_gradient_compute.glsl (4-tap tetrahedral
// gradient, one invocation per voxel with four clamped imageLoads) and
// ComputeGradientMap::compute (src/compute_gradient_map.cpp:57-81).
//
// B200 design: each thread produces 16 consecutive voxels from four 16-byte vector loads
// (the four (y±1, z±1) rows the taps live on); the x±1 neighbours of the run come from the
// adjacent lanes by warp shuffle, so every input byte is loaded once per output row.  Results
// are packed into one 16-byte store.  The arithmetic follows the shader's operation order in
// fp32 without contraction and with IEEE division / square root so the stored byte is
// identical to the CPU oracle's (the byte feeds the occupancy LUT).  UNORM decode b/255 is
// served from a 256-entry shared-memory table built with a true division.
// Algorithmic bytes: read N + write N = 2 B/voxel.
#include "common.cuh"

namespace vkv {

__device__ __forceinline__ unsigned char gradient_byte(float a, float b, float c, float d, float modifier)
{
	// 0.25 * (k.xyy*a + k.yyx*b + k.yxy*c + k.xxx*d), k = (1,-1)  (get_gradient_compute.glsl:13-18)
	const float gx  = 0.25f * (((a - b) - c) + d);
	const float gy  = 0.25f * (((-a - b) + c) + d);
	const float gz  = 0.25f * (((-a + b) - c) + d);
	const float len = sqrtf((gx * gx + gy * gy) + gz * gz);
	const float g   = fminf(fmaxf(len * modifier, 0.0f), 1.0f);
	return (unsigned char) rintf(g * 255.0f);        // imageStore to r8: round to nearest
}

__device__ __forceinline__ unsigned byte_of(const unsigned w[4], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xffu; }

__global__ void __launch_bounds__(256) gradient_vec16_kernel(const uint8_t *__restrict__ V, uint8_t *__restrict__ G, uint32_t W,
                                                            uint32_t H, uint32_t D, float modifier)
{
	__shared__ float s_lut[256];
	s_lut[threadIdx.x] = (float) threadIdx.x / 255.0f;        // UNORM decode, exactly as imageLoad
	__syncthreads();
	const uint32_t nchunks = W / 16;
	const uint64_t total   = (uint64_t) nchunks * H * D;
	const int      lane    = threadIdx.x & 31;
	for (uint64_t base = (uint64_t) blockIdx.x * blockDim.x; base < total; base += (uint64_t) gridDim.x * blockDim.x) {
		const uint64_t idx    = base + threadIdx.x;
		const bool     active = idx < total;
		const uint64_t cidx   = active ? idx : total - 1;
		const uint32_t chunk  = (uint32_t) (cidx % nchunks);
		const uint64_t r      = cidx / nchunks;
		const uint32_t y = (uint32_t) (r % H), z = (uint32_t) (r / H);
		const uint32_t ym = y > 0 ? y - 1 : 0, yp = y + 1 < H ? y + 1 : H - 1;
		const uint32_t zm = z > 0 ? z - 1 : 0, zp = z + 1 < D ? z + 1 : D - 1;
		const uint32_t x = chunk * 16;
		// rows: A = (y-1, z-1) read at x+1 | B = (y-1, z+1) at x-1 | C = (y+1, z-1) at x-1 | E = (y+1, z+1) at x+1
		const uint8_t *rowp[4] = {V + ((size_t) zm * H + ym) * W + x, V + ((size_t) zp * H + ym) * W + x,
		                          V + ((size_t) zm * H + yp) * W + x, V + ((size_t) zp * H + yp) * W + x};
		unsigned win[4][4];        // the 16-byte window of each row, already shifted by its x offset
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint4    q    = __ldg(reinterpret_cast<const uint4 *>(rowp[k]));
			const unsigned w[4] = {q.x, q.y, q.z, q.w};
			const bool     plus = (k == 0 || k == 3);        // tap at x+1 (else x-1)
			// neighbour words from adjacent lanes: same row iff the neighbour is the adjacent chunk
			const unsigned up   = __shfl_up_sync(0xffffffffu, w[3], 1);
			const unsigned down = __shfl_down_sync(0xffffffffu, w[0], 1);
			if (plus) {
				unsigned right;        // byte x+16 (clamped to W-1) in bits 0..7
				if (chunk == nchunks - 1) right = w[3] >> 24;
				else if (lane == 31) right = rowp[k][16];
				else right = down;
				win[k][0] = __funnelshift_r(w[0], w[1], 8);
				win[k][1] = __funnelshift_r(w[1], w[2], 8);
				win[k][2] = __funnelshift_r(w[2], w[3], 8);
				win[k][3] = __funnelshift_r(w[3], right, 8);
			} else {
				unsigned left;        // byte x-1 (clamped to 0) in bits 24..31
				if (chunk == 0) left = w[0] << 24;
				else if (lane == 0) left = ((unsigned) *(rowp[k] - 1)) << 24;
				else left = up;
				win[k][0] = __funnelshift_r(left, w[0], 24);
				win[k][1] = __funnelshift_r(w[0], w[1], 24);
				win[k][2] = __funnelshift_r(w[1], w[2], 24);
				win[k][3] = __funnelshift_r(w[2], w[3], 24);
			}
		}
		unsigned out[4] = {0, 0, 0, 0};
#pragma unroll
		for (int i = 0; i < 16; ++i) {
			const float a = s_lut[byte_of(win[0], i)];        // (+1,-1,-1)
			const float b = s_lut[byte_of(win[1], i)];        // (-1,-1,+1)
			const float c = s_lut[byte_of(win[2], i)];        // (-1,+1,-1)
			const float d = s_lut[byte_of(win[3], i)];        // (+1,+1,+1)
			out[i >> 2] |= (unsigned) gradient_byte(a, b, c, d, modifier) << (8 * (i & 3));
		}
		if (active) *reinterpret_cast<uint4 *>(G + ((size_t) z * H + y) * W + x) = make_uint4(out[0], out[1], out[2], out[3]);
	}
}

// Any extents: one thread per voxel, clamped byte loads through L1.
__global__ void __launch_bounds__(256) gradient_scalar_kernel(const uint8_t *__restrict__ V, uint8_t *__restrict__ G, uint32_t W,
                                                             uint32_t H, uint32_t D, float modifier)
{
	const uint64_t total = (uint64_t) W * H * D;
	for (uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t) gridDim.x * blockDim.x) {
		const uint32_t x = (uint32_t) (idx % W);
		const uint64_t r = idx / W;
		const uint32_t y = (uint32_t) (r % H), z = (uint32_t) (r / H);
		const uint32_t xm = x > 0 ? x - 1 : 0, xp = x + 1 < W ? x + 1 : W - 1;
		const uint32_t ym = y > 0 ? y - 1 : 0, yp = y + 1 < H ? y + 1 : H - 1;
		const uint32_t zm = z > 0 ? z - 1 : 0, zp = z + 1 < D ? z + 1 : D - 1;
		const float a = (float) V[((size_t) zm * H + ym) * W + xp] / 255.0f;
		const float b = (float) V[((size_t) zp * H + ym) * W + xm] / 255.0f;
		const float c = (float) V[((size_t) zm * H + yp) * W + xm] / 255.0f;
		const float d = (float) V[((size_t) zp * H + yp) * W + xp] / 255.0f;
		G[idx] = gradient_byte(a, b, c, d, modifier);
	}
}

__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t *__restrict__ p, uint32_t value, size_t n_words)
{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t) gridDim.x * blockDim.x) p[i] = value;
}

int launch_gradient(vkv_volume *vol, bool use_gradient, float modifier, cudaStream_t s)
{
	const int grid = vol->ctx->sm_count * 8;
	if (!use_gradient) {
		// get_gradient returns 1.0 for every voxel (get_gradient_compute.glsl:6-7) -> byte 255
		VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_G, 255, vol->N, s));
	} else if (vol->dim[0] % 16 == 0 && reinterpret_cast<uintptr_t>(vol->d_V) % 16 == 0) {
		gradient_vec16_kernel<<<grid, 256, 0, s>>>(vol->d_V, vol->d_G, vol->dim[0], vol->dim[1], vol->dim[2], modifier);
		VKV_LAUNCHED();
	} else {
		gradient_scalar_kernel<<<grid, 256, 0, s>>>(vol->d_V, vol->d_G, vol->dim[0], vol->dim[1], vol->dim[2], modifier);
		VKV_LAUNCHED();
	}
	vol->has_G = true;
	return VKV_OK;
}

}        // namespace vkv
