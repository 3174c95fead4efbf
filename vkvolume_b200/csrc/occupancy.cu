// occupancy.cu — K2a occupancy map fused with K2b/K2c occupied-voxel count.
//
// Replaces shaders/occupancy_map.comp (one thread serially scanning a bs^3 block),
// shaders/occupied_voxel_count.comp + occupied_voxel_count_reduce.comp (a per-voxel pass
// over V,G followed by log_S(n) strided reduce dispatches) and their host drivers
// (src/compute_distance_map.cpp:103-140, src/compute_occupied_voxel_count.cpp:80-147).
//
// B200 design: this is the only O(N)-byte pass of a transfer-function change, so it is
// written as one HBM-streaming kernel: 16-byte vector loads of V and G (8 loads in flight
// per thread), visibility by bit lookup in a 16 KB shared-memory mask (texture mask for the
// occupancy map, analytic mask for the count — the reference really uses two different
// transfer functions), a SIMD byte-range prefilter that rejects four voxels
// per instruction when none of them can be visible, warp-shuffle OR across the rows of a
// block, and a persistent grid (a multiple of the SM count) so the count needs one
// atomicAdd per CTA.  Algorithmic bytes: 2 B/voxel + 1 B/block (1 B/voxel when the gradient
// is unused).
#include "common.cuh"

namespace vkv {

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}

// SWAR byte-range prefilter: bit 7 of every byte of the result is set iff lo <= byte <= hi.
// "byte >= lo" needs one AND, one ADD and one LOP3 per 4 voxels (no per-byte video instructions):
//   lo <= 128: the byte passes if its top bit is set or its low 7 bits reach lo        -> ((x & 0x7f..) + (128-lo)) | x
//   lo >  128: the byte passes if its top bit is set and its low 7 bits reach lo - 128 -> ((x & 0x7f..) + (256-lo)) & x
// "byte <= hi" is the same test on ~x with lo' = 255 - hi.
struct ByteGE {
	unsigned add, or_sel;        // or_sel = ~0 selects the OR form
	__device__ __forceinline__ void set(unsigned lo)
	{
		or_sel = lo <= 128u ? 0xffffffffu : 0u;
		add    = (lo <= 128u ? 128u - lo : 256u - lo) * 0x01010101u;
	}
	__device__ __forceinline__ unsigned test(unsigned x) const
	{
		const unsigned m = (x & 0x7f7f7f7fu) + add;
		return (m & x) | ((m | x) & or_sel);        // caller masks with 0x80808080
	}
};

// Looks up the 4 voxels of word (v,g); returns occupancy hit in bit 0 and the
// number of analytically visible voxels in bits 8.. .
template <bool USE_G, bool COUNT>
__device__ __forceinline__ void classify_word(unsigned v, unsigned g, unsigned cand, const uint2 *__restrict__ s_mask, bool &hit,
                                              unsigned &cnt)
{
#pragma unroll
	for (int b = 0; b < 4; ++b) {
		if ((cand >> (8 * b + 7)) & 1u) {
			const unsigned vb  = (v >> (8 * b)) & 0xffu;
			const unsigned gb  = USE_G ? ((g >> (8 * b)) & 0xffu) : 255u;
			const uint2    m   = s_mask[gb * 8 + (vb >> 5)];
			const unsigned sh  = vb & 31u;
			hit |= (m.x >> sh) & 1u;
			if (COUNT) cnt += (m.y >> sh) & 1u;
		}
	}
}

// Fast path: cubic effective block size BS in {2,4,8}, W % 16 == 0.
// A warp covers (32/BS * 16) voxels in x  x  BS rows in y  x  BS slices in z = 32/BS*16/BS blocks.
template <int BS, bool USE_G, bool COUNT>
__global__ void __launch_bounds__(256) occupancy_fast_kernel(const uint8_t *__restrict__ V, const uint8_t *__restrict__ G,
                                                            const uint2 *__restrict__ mask2, const TFBounds *__restrict__ bounds,
                                                            uint32_t W, uint32_t H, uint32_t D, uint32_t Wb, uint32_t Hb,
                                                            uint32_t zb_first, uint32_t zb_count, uint8_t *__restrict__ O,
                                                            unsigned long long *__restrict__ count)
{
	constexpr int TX  = 32 / BS;        // lanes along x
	constexpr int XV  = TX * 16;        // voxels along x per warp
	constexpr int BPT = 16 / BS;        // blocks per thread
	__shared__ uint2    s_mask[kMaskWords];
	for (int i = threadIdx.x; i < kMaskWords; i += blockDim.x) s_mask[i] = mask2[i];
	const int      bsel = (USE_G ? 0 : 2) + (COUNT ? 1 : 0);
	const unsigned vlo = bounds->v_lo[bsel], vhi = bounds->v_hi[bsel];
	const unsigned glo = bounds->g_lo[bsel], ghi = bounds->g_hi[bsel];
	const bool     none = vlo > vhi;        // nothing visible anywhere
	ByteGE         v_ge, v_le, g_ge, g_le;
	v_ge.set(vlo); v_le.set(255u - vhi); g_ge.set(glo); g_le.set(255u - ghi);
	const bool need_vlo = vlo > 0u, need_vhi = vhi < 255u, need_glo = USE_G && glo > 0u, need_ghi = USE_G && ghi < 255u;
	__syncthreads();

	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int      tx = lane % TX, ty = lane / TX;
	const uint32_t xchunks = (W + XV - 1) / XV;
	const uint32_t ntasks  = xchunks * Hb * zb_count;        // < 2^32 (checked by the launcher)
	unsigned long long local_count = 0;

	for (uint32_t task = blockIdx.x * 8 + warp; task < ntasks; task += gridDim.x * 8) {
		const uint32_t r  = task / xchunks, xc = task - r * xchunks;
		const uint32_t bq = r / Hb, by = r - bq * Hb, bz = zb_first + bq;
		const uint32_t x = xc * XV + tx * 16, y = by * BS + ty;
		const bool     in_xy = x < W && y < H;
		uint4          vv[BS], gg[BS];
#pragma unroll
		for (int zz = 0; zz < BS; ++zz) {
			const uint32_t z = bz * BS + zz;
			vv[zz] = make_uint4(0, 0, 0, 0);
			gg[zz] = make_uint4(0, 0, 0, 0);
			if (in_xy && z < D) {
				const size_t off = ((size_t) z * H + y) * W + x;
				vv[zz] = ldg_stream(reinterpret_cast<const uint4 *>(V + off));
				if (USE_G) gg[zz] = ldg_stream(reinterpret_cast<const uint4 *>(G + off));
			}
		}
		unsigned flags = 0;        // bit k: block k of this thread's 16-voxel run is occupied
		unsigned cnt   = 0;
		auto candidates = [&](unsigned vword, unsigned gword) -> unsigned {
			unsigned c = 0x80808080u;
			if (need_vlo) c &= v_ge.test(vword);
			if (need_vhi) c &= v_le.test(~vword);
			if (need_glo) c &= g_ge.test(gword);
			if (need_ghi) c &= g_le.test(~gword);
			return c;
		};
		// branch-free prefilter over all 16*BS voxels of this thread; the LUT path below is entered by few warps
		unsigned any = 0;
		if (!none) {
#pragma unroll
			for (int zz = 0; zz < BS; ++zz) {
				any |= candidates(vv[zz].x, gg[zz].x) | candidates(vv[zz].y, gg[zz].y) | candidates(vv[zz].z, gg[zz].z) | candidates(vv[zz].w, gg[zz].w);
			}
		}
		if (any) {
#pragma unroll
			for (int zz = 0; zz < BS; ++zz) {
				const bool valid = in_xy && (bz * BS + zz) < D;
				const unsigned vw[4] = {vv[zz].x, vv[zz].y, vv[zz].z, vv[zz].w};
				const unsigned gw[4] = {gg[zz].x, gg[zz].y, gg[zz].z, gg[zz].w};
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const unsigned cand = candidates(vw[k], gw[k]);
					if (valid && cand) {
						if (BS >= 4) {
							bool hit = false;
							classify_word<USE_G, COUNT>(vw[k], gw[k], cand, s_mask, hit, cnt);
							if (hit) flags |= 1u << ((k * 4) / BS);
						} else {        // BS == 2: two blocks per word
#pragma unroll
							for (int h = 0; h < 2; ++h) {
								bool hit = false;
								classify_word<USE_G, COUNT>(vw[k] >> (16 * h), gw[k] >> (16 * h), (cand >> (16 * h)) & 0x8080u, s_mask, hit, cnt);
								if (hit) flags |= 1u << (k * 2 + h);
							}
						}
					}
				}
			}
		}
		// OR across the BS rows of the block row (lanes ty = 0..BS-1 share tx)
#pragma unroll
		for (int o = TX; o < 32; o <<= 1) flags |= __shfl_xor_sync(0xffffffffu, flags, o);
		if (ty == 0 && x < W && O) {
			const size_t ob = ((size_t) bz * Hb + by) * Wb + (x / BS);
			if (BPT == 4) {
				uchar4 o4 = make_uchar4((flags & 1u) ? 0 : 255, (flags & 2u) ? 0 : 255, (flags & 4u) ? 0 : 255, (flags & 8u) ? 0 : 255);
				*reinterpret_cast<uchar4 *>(O + ob) = o4;
			} else {
#pragma unroll
				for (int k = 0; k < BPT; ++k) O[ob + k] = ((flags >> k) & 1u) ? 0 : 255;
			}
		}
		if (COUNT) local_count += cnt;
	}

	if (COUNT) {
		unsigned long long c = local_count;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
		__shared__ unsigned long long s_c[8];
		if (lane == 0) s_c[warp] = c;
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned long long t = 0;
			for (int i = 0; i < 8; ++i) t += s_c[i];
			if (t) atomicAdd(count, t);
		}
	}
}

// Generic path: any per-axis effective block size, any extents.  One CTA per
// (block row, block slice, x segment of whole blocks); coalesced byte loads; shared flag array.
template <bool USE_G, bool COUNT>
__global__ void __launch_bounds__(256) occupancy_generic_kernel(const uint8_t *__restrict__ V, const uint8_t *__restrict__ G,
                                                               const uint2 *__restrict__ mask2, uint32_t W, uint32_t H, uint32_t D,
                                                               uint32_t Wb, uint32_t Hb, uint32_t bsx, uint32_t bsy, uint32_t bsz,
                                                               uint32_t zb_first, uint32_t zb_count, uint8_t *__restrict__ O,
                                                               unsigned long long *__restrict__ count)
{
	constexpr uint32_t MAXB = 1024;        // blocks of x handled per CTA pass
	__shared__ uint2              s_mask[kMaskWords];
	__shared__ unsigned char      s_flag[MAXB];
	__shared__ unsigned long long s_c[8];
	for (int i = threadIdx.x; i < kMaskWords; i += blockDim.x) s_mask[i] = mask2[i];
	__syncthreads();
	const uint32_t     seg_blocks = max(1u, MAXB / bsx);
	const uint32_t     xsegs      = (Wb + seg_blocks - 1) / seg_blocks;
	const uint64_t     ntasks     = (uint64_t) xsegs * Hb * zb_count;
	unsigned long long local      = 0;
	for (uint64_t task = blockIdx.x; task < ntasks; task += gridDim.x) {
		const uint32_t xs = (uint32_t) (task % xsegs);
		const uint64_t r  = task / xsegs;
		const uint32_t by = (uint32_t) (r % Hb), bz = zb_first + (uint32_t) (r / Hb);
		const uint32_t b0 = xs * seg_blocks, b1 = min(Wb, b0 + seg_blocks);        // [b0, b1)
		const uint32_t x0 = b0 * bsx, x1 = min(W, b1 * bsx);
		for (uint32_t i = threadIdx.x; i < b1 - b0; i += blockDim.x) s_flag[i] = 255;
		__syncthreads();
		const uint32_t ye = min(H, by * bsy + bsy), ze = min(D, bz * bsz + bsz);
		for (uint32_t z = bz * bsz; z < ze; ++z)
			for (uint32_t y = by * bsy; y < ye; ++y) {
				const size_t row = ((size_t) z * H + y) * W;
				for (uint32_t x = x0 + threadIdx.x; x < x1; x += blockDim.x) {
					const unsigned vb = V[row + x];
					const unsigned gb = USE_G ? G[row + x] : 255u;
					const uint2    m  = s_mask[gb * 8 + (vb >> 5)];
					if ((m.x >> (vb & 31u)) & 1u) s_flag[x / bsx - b0] = 0;
					if (COUNT) local += (m.y >> (vb & 31u)) & 1u;
				}
			}
		__syncthreads();
		if (O) {
			uint8_t *dst = O + ((size_t) bz * Hb + by) * Wb + b0;
			for (uint32_t i = threadIdx.x; i < b1 - b0; i += blockDim.x) dst[i] = s_flag[i];
		}
		__syncthreads();
	}
	if (COUNT) {
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
		if (lane == 0) s_c[warp] = local;
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned long long t = 0;
			for (int i = 0; i < 8; ++i) t += s_c[i];
			if (t) atomicAdd(count, t);
		}
	}
}

template <int BS, bool USE_G, bool COUNT>
static int launch_fast_inst(vkv_volume *vol, uint8_t *O, uint32_t zb_first, uint32_t zb_count, unsigned long long *count_dev, cudaStream_t s)
{
	// persistent grid: exactly the number of CTAs that are resident at once (a multiple of the SM count)
	static int per_sm = 0;
	if (per_sm == 0) {
		VKV_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, occupancy_fast_kernel<BS, USE_G, COUNT>, 256, 0));
		if (per_sm < 1) per_sm = 1;
	}
	occupancy_fast_kernel<BS, USE_G, COUNT><<<vol->ctx->sm_count * per_sm, 256, 0, s>>>(vol->d_V, vol->d_G, vol->d_mask2, vol->d_bounds, vol->dim[0], vol->dim[1],
	                                                                                    vol->dim[2], vol->dim_b[0], vol->dim_b[1], zb_first, zb_count, O, count_dev);
	VKV_LAUNCHED();
	return VKV_OK;
}

template <int BS>
static int launch_fast(vkv_volume *vol, bool use_g, bool count, uint8_t *O, uint32_t zb_first, uint32_t zb_count,
                       unsigned long long *count_dev, int grid, cudaStream_t s)
{
	(void) grid;
	if (use_g && count) return launch_fast_inst<BS, true, true>(vol, O, zb_first, zb_count, count_dev, s);
	if (use_g) return launch_fast_inst<BS, true, false>(vol, O, zb_first, zb_count, count_dev, s);
	if (count) return launch_fast_inst<BS, false, true>(vol, O, zb_first, zb_count, count_dev, s);
	return launch_fast_inst<BS, false, false>(vol, O, zb_first, zb_count, count_dev, s);
}

int launch_occupancy(vkv_volume *vol, bool use_gradient, bool count, uint8_t *O, uint32_t zb_first, uint32_t zb_count,
                     unsigned long long *count_dev, cudaStream_t s)
{
	if (zb_count == 0) return VKV_OK;
	// gradient bytes are only read when the TF uses them AND a precomputed map exists
	const bool use_g = use_gradient;
	const int  grid  = vol->ctx->sm_count * 8;        // persistent: 8 CTAs/SM x 256 threads = full occupancy
	const bool cubic = vol->bs[0] == vol->bs[1] && vol->bs[1] == vol->bs[2];
	const uint64_t fast_tasks = (uint64_t) ((vol->dim[0] + 63) / 64) * vol->dim_b[1] * zb_count;        // upper bound over BS
	const bool fast  = fast_tasks < (1ull << 32) && cubic && (vol->bs[0] == 2 || vol->bs[0] == 4 || vol->bs[0] == 8) && vol->dim[0] % 16 == 0 &&
	                  vol->dim_b[0] * vol->bs[0] == vol->dim[0] && (reinterpret_cast<uintptr_t>(vol->d_V) % 16 == 0);
	if (fast) {
		switch (vol->bs[0]) {
			case 2: return launch_fast<2>(vol, use_g, count, O, zb_first, zb_count, count_dev, grid, s);
			case 4: return launch_fast<4>(vol, use_g, count, O, zb_first, zb_count, count_dev, grid, s);
			default: return launch_fast<8>(vol, use_g, count, O, zb_first, zb_count, count_dev, grid, s);
		}
	}
#define VKV_OCC_ARGS vol->d_V, vol->d_G, vol->d_mask2, vol->dim[0], vol->dim[1], vol->dim[2], vol->dim_b[0], vol->dim_b[1], vol->bs[0], vol->bs[1], vol->bs[2], zb_first, zb_count, O, count_dev
	if (use_g && count) occupancy_generic_kernel<true, true><<<grid, 256, 0, s>>>(VKV_OCC_ARGS);
	else if (use_g) occupancy_generic_kernel<true, false><<<grid, 256, 0, s>>>(VKV_OCC_ARGS);
	else if (count) occupancy_generic_kernel<false, true><<<grid, 256, 0, s>>>(VKV_OCC_ARGS);
	else occupancy_generic_kernel<false, false><<<grid, 256, 0, s>>>(VKV_OCC_ARGS);
#undef VKV_OCC_ARGS
	VKV_LAUNCHED();
	return VKV_OK;
}

}        // namespace vkv
