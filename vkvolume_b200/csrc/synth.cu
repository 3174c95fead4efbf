// synth.cu — bench/test utilities that are NOT part of the reference surface:
// seeded synthetic volumes generated directly in HBM (BASELINE.json configs use synthetic
// stand-ins for datasets that are not in the reference repository) and a tex3D throughput
// microbenchmark that provides the ray caster's roofline denominator.
#include "common.cuh"

namespace vkv {

struct SynthPrim {
	float cx, cy, cz;        // centre / first end point, in units of max(dim) voxels
	float ex, ey, ez;        // second end point (capsules)
	float r, amp;
};
struct SynthParams {
	int       kind, n_prims;
	uint32_t  W, H, D;
	float     inv_max;
	uint32_t  seed_lo, seed_hi;
	SynthPrim prims[64];
};

__host__ __device__ inline uint64_t splitmix64(uint64_t &x)
{
	uint64_t z = (x += 0x9E3779B97F4A7C15ull);
	z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
	uint32_t h = x * 0x8da6b343u ^ y * 0xd8163841u ^ z * 0xcb1ab31fu ^ seed;
	h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
	return h;
}

__device__ __forceinline__ float capsule_dist2(float px, float py, float pz, const SynthPrim &c)
{
	const float bx = c.ex - c.cx, by = c.ey - c.cy, bz = c.ez - c.cz;
	const float ax = px - c.cx, ay = py - c.cy, az = pz - c.cz;
	float       t  = (ax * bx + ay * by + az * bz) / fmaxf(bx * bx + by * by + bz * bz, 1e-12f);
	t              = fminf(fmaxf(t, 0.0f), 1.0f);
	const float dx = ax - t * bx, dy = ay - t * by, dz = az - t * bz;
	return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(256) synth_kernel(const __grid_constant__ SynthParams P, uint8_t *__restrict__ out)
{
	const uint64_t total = (uint64_t) P.W * P.H * P.D;
	for (uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t) gridDim.x * blockDim.x) {
		const uint32_t x = (uint32_t) (idx % P.W);
		const uint64_t r = idx / P.W;
		const uint32_t y = (uint32_t) (r % P.H), z = (uint32_t) (r / P.H);
		const float    px = (x + 0.5f) * P.inv_max, py = (y + 0.5f) * P.inv_max, pz = (z + 0.5f) * P.inv_max;
		const uint32_t h  = hash3(x, y, z, P.seed_lo);
		float          v  = 0.0f;
		if (P.kind == 0 || P.kind == 3) {        // Gaussian blobs
			for (int i = 0; i < P.n_prims; ++i) {
				const SynthPrim &b  = P.prims[i];
				const float      dx = px - b.cx, dy = py - b.cy, dz = pz - b.cz;
				const float      d2 = dx * dx + dy * dy + dz * dz, s2 = b.r * b.r;
				if (d2 < 9.0f * s2) v += b.amp * __expf(-0.5f * d2 / s2);
			}
			if (P.kind == 3) v *= 0.5f + 0.5f * ((hash3(x >> 2, y >> 2, z >> 2, P.seed_hi) & 0xffffu) * (1.0f / 65535.0f));
			v += (float) (h & 7u) - 4.0f + 4.0f;        // 0..7 background noise
		} else if (P.kind == 1) {                      // ellipsoidal shell + capsule legs
			const SynthPrim &e  = P.prims[0];          // centre c, radii in (ex,ey,ez), thickness r
			const float      qx = (px - e.cx) / e.ex, qy = (py - e.cy) / e.ey, qz = (pz - e.cz) / e.ez;
			const float      rr = sqrtf(qx * qx + qy * qy + qz * qz);
			const float      sh = fabsf(rr - 1.0f) * fminf(e.ex, fminf(e.ey, e.ez));        // ~distance to the shell
			if (sh < e.r) v = e.amp * (1.0f - 0.6f * sh / e.r);
			for (int i = 1; i < P.n_prims; ++i) {
				const SynthPrim &c  = P.prims[i];
				const float      d2 = capsule_dist2(px, py, pz, c);
				if (d2 < c.r * c.r) v = fmaxf(v, c.amp * (1.0f - 0.5f * d2 / (c.r * c.r)));
			}
			v += (float) (h % 13u);        // background noise 0..12 stays below the TF threshold
		} else {                           // kind 2: sparse tubes
			for (int i = 0; i < P.n_prims; ++i) {
				const SynthPrim &c  = P.prims[i];
				const float      d2 = capsule_dist2(px, py, pz, c);
				if (d2 < c.r * c.r) v = fmaxf(v, c.amp * (1.0f - 0.5f * d2 / (c.r * c.r)));
			}
			v += (float) (h % 9u);
		}
		out[idx] = (uint8_t) fminf(fmaxf(v, 0.0f), 255.0f);
	}
}

// tex3D throughput: every thread issues `n` filtered fetches along a ray-like path.
__global__ void __launch_bounds__(256) texbench_kernel(cudaTextureObject_t tex, int n, int coherent, float inv_extent, float *__restrict__ sink)
{
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	float          acc = 0.0f;
	if (coherent) {
		// a warp is an 8x4 pixel tile marching along z with a one-texel step, like the ray caster
		const uint32_t w = tid >> 5, l = tid & 31;
		float          x = ((w * 8u) % 1024u + (l & 7)) * inv_extent, y = (((w * 8u) / 1024u) * 4u % 1024u + (l >> 3)) * inv_extent;
		x -= floorf(x); y -= floorf(y);
		float z = (hash3(w, 0, 0, 7u) & 1023u) * inv_extent;
#pragma unroll 4
		for (int i = 0; i < n; ++i) {
			acc += tex3D<float>(tex, x, y, z);
			z += inv_extent;
			if (z > 1.0f) z -= 1.0f;
		}
	} else {
		uint32_t h = hash3(tid, 1, 2, 3u);
#pragma unroll 4
		for (int i = 0; i < n; ++i) {
			h = h * 1664525u + 1013904223u;
			const float x = (h & 0x3ffu) * (1.0f / 1024.0f), y = ((h >> 10) & 0x3ffu) * (1.0f / 1024.0f), z = ((h >> 20) & 0x3ffu) * (1.0f / 1024.0f);
			acc += tex3D<float>(tex, x, y, z);
		}
	}
	sink[tid] = acc;
}

}        // namespace vkv

using namespace vkv;

extern "C" {

int vkv_synth_volume(vkv_context *ctx, int kind, uint64_t seed, uint32_t W, uint32_t H, uint32_t D, uint8_t *dev_out, void *stream)
{
	VKV_REQUIRE(ctx && dev_out, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(kind >= 0 && kind <= 3, VKV_ERR_ARGUMENT, "bad synthetic volume kind");
	SynthParams P{};
	P.kind = kind; P.W = W; P.H = H; P.D = D;
	const float mx = (float) std::max(W, std::max(H, D));
	P.inv_max      = 1.0f / mx;
	const float ext[3] = {W / mx, H / mx, D / mx};
	uint64_t    st = seed;
	auto        rnd = [&]() { return (float) ((splitmix64(st) >> 40) * (1.0 / 16777216.0)); };
	P.seed_lo = (uint32_t) splitmix64(st);
	P.seed_hi = (uint32_t) splitmix64(st);
	if (kind == 0 || kind == 3) {
		P.n_prims = 64;
		for (int i = 0; i < 64; ++i) {
			SynthPrim &b = P.prims[i];
			b.cx = rnd() * ext[0]; b.cy = rnd() * ext[1]; b.cz = rnd() * ext[2];
			// sigma 6..24 voxels at 256^3 (kind 0), tighter blobs for the sparse big volume (kind 3)
			b.r   = kind == 0 ? (6.0f + 18.0f * rnd()) / 256.0f : (0.006f + 0.02f * rnd());
			b.amp = 64.0f + 191.0f * rnd();
		}
	} else if (kind == 1) {
		P.n_prims    = 7;
		SynthPrim &e = P.prims[0];
		e.cx = 0.5f * ext[0]; e.cy = 0.5f * ext[1]; e.cz = 0.5f * ext[2];
		e.ex = 0.30f * ext[0]; e.ey = 0.22f * ext[1]; e.ez = 0.36f * ext[2];
		e.r   = 0.02f;
		e.amp = 200.0f;
		for (int i = 1; i < 7; ++i) {        // six legs
			SynthPrim &c = P.prims[i];
			const float side = (i & 1) ? 1.0f : -1.0f, along = ((i - 1) / 2 - 1) * 0.18f;
			c.cx = e.cx + side * 0.25f * ext[0]; c.cy = e.cy + 0.1f * ext[1]; c.cz = e.cz + along * ext[2];
			c.ex = e.cx + side * (0.42f + 0.04f * rnd()) * ext[0]; c.ey = e.cy + (0.30f + 0.1f * rnd()) * ext[1];
			c.ez = c.cz + (rnd() - 0.5f) * 0.1f;
			c.r   = 0.012f;
			c.amp = 170.0f;
		}
	} else {
		P.n_prims = 24;
		for (int i = 0; i < 24; ++i) {
			SynthPrim &c = P.prims[i];
			c.cx = rnd() * ext[0]; c.cy = rnd() * ext[1]; c.cz = rnd() * ext[2];
			c.ex = c.cx + (rnd() - 0.5f) * 0.6f; c.ey = c.cy + (rnd() - 0.5f) * 0.6f; c.ez = c.cz + (rnd() - 0.5f) * 0.6f;
			c.r   = 0.006f + 0.006f * rnd();
			c.amp = 150.0f + 100.0f * rnd();
		}
	}
	synth_kernel<<<ctx->sm_count * 8, 256, 0, (cudaStream_t) stream>>>(P, dev_out);
	VKV_LAUNCHED();
	return VKV_OK;
}

int vkv_bench_tex3d(vkv_context *ctx, uint32_t extent, int fetches_per_thread, int coherent, double *out)
{
	VKV_REQUIRE(ctx && out && extent > 0 && fetches_per_thread > 0, VKV_ERR_ARGUMENT, "bad argument");
	const size_t N = (size_t) extent * extent * extent;
	uint8_t     *lin = nullptr;
	cudaArray_t  arr = nullptr;
	float       *sink = nullptr;
	cudaTextureObject_t tex = 0;
	cudaEvent_t  e0 = nullptr, e1 = nullptr;
	int          rc = VKV_OK;
	auto         cleanup = [&]() {
        if (tex) cudaDestroyTextureObject(tex);
        if (arr) cudaFreeArray(arr);
        cudaFree(lin); cudaFree(sink);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
	};
#define TB_CHECK(x)                                                          \
	do {                                                                     \
		cudaError_t _e = (x);                                                \
		if (_e != cudaSuccess) {                                             \
			set_error("vkv_bench_tex3d: %s: %s", #x, cudaGetErrorString(_e)); \
			cleanup();                                                       \
			return VKV_ERR_CUDA;                                             \
		}                                                                    \
	} while (0)
	TB_CHECK(cudaMalloc(&lin, N));
	if ((rc = vkv_synth_volume(ctx, 0, 42, extent, extent, extent, lin, nullptr))) { cleanup(); return rc; }
	cudaChannelFormatDesc fd = cudaCreateChannelDesc<unsigned char>();
	TB_CHECK(cudaMalloc3DArray(&arr, &fd, make_cudaExtent(extent, extent, extent)));
	cudaMemcpy3DParms p{};
	p.srcPtr = make_cudaPitchedPtr(lin, extent, extent, extent);
	p.dstArray = arr;
	p.extent = make_cudaExtent(extent, extent, extent);
	p.kind = cudaMemcpyDeviceToDevice;
	TB_CHECK(cudaMemcpy3D(&p));
	cudaResourceDesc rd{};
	rd.resType = cudaResourceTypeArray;
	rd.res.array.array = arr;
	cudaTextureDesc td{};
	td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
	td.filterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeNormalizedFloat;
	td.normalizedCoords = 1;
	TB_CHECK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
	const int grid = ctx->sm_count * 16, threads = 256;
	TB_CHECK(cudaMalloc(&sink, (size_t) grid * threads * sizeof(float)));
	TB_CHECK(cudaEventCreate(&e0));
	TB_CHECK(cudaEventCreate(&e1));
	const float inv = 1.0f / (float) extent;
	for (int i = 0; i < 3; ++i) texbench_kernel<<<grid, threads>>>(tex, fetches_per_thread, coherent, inv, sink);
	TB_CHECK(cudaGetLastError());
	TB_CHECK(cudaEventRecord(e0));
	const int reps = 5;
	for (int i = 0; i < reps; ++i) texbench_kernel<<<grid, threads>>>(tex, fetches_per_thread, coherent, inv, sink);
	g_kernel_launches.fetch_add(3 + reps);
	TB_CHECK(cudaEventRecord(e1));
	TB_CHECK(cudaEventSynchronize(e1));
	float ms = 0;
	TB_CHECK(cudaEventElapsedTime(&ms, e0, e1));
	*out = (double) grid * threads * fetches_per_thread * reps / (ms * 1e-3);
#undef TB_CHECK
	cleanup();
	return VKV_OK;
}

}        // extern "C"
