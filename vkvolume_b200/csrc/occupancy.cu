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
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

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

// The candidate rectangle of a pass: the texture mask's ranges, widened by the analytic mask's when the count is fused.
struct Prefilter {
	unsigned vlo, vhi, glo, ghi;
	bool     none;
	__device__ __forceinline__ void load(const TFBounds *b, bool use_g, bool count)
	{
		const TFRange &t = use_g ? b->tex_all : b->tex_row255;
		const TFRange &a = use_g ? b->ana_all : b->ana_row255;
		vlo = t.v_lo; vhi = t.v_hi; glo = t.g_lo; ghi = t.g_hi;
		bool empty = t.v_lo > t.v_hi;
		if (count && a.v_lo <= a.v_hi) {
			if (empty) { vlo = a.v_lo; vhi = a.v_hi; glo = a.g_lo; ghi = a.g_hi; empty = false; }
			else { vlo = min(vlo, a.v_lo); vhi = max(vhi, a.v_hi); glo = min(glo, a.g_lo); ghi = max(ghi, a.g_hi); }
		}
		none = empty;
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
	Prefilter pf;
	pf.load(bounds, USE_G, COUNT);
	const unsigned vlo = pf.vlo, vhi = pf.vhi, glo = pf.glo, ghi = pf.ghi;
	const bool     none = pf.none;        // nothing visible anywhere
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

// Byte-range membership of the 4 voxels of a (V, G) word pair: bit 7 of every byte of the result is set iff
// v_lo <= v <= v_hi (and g_lo <= g <= g_hi).  HI = false drops the upper tests (the built-in ramp transfer function
// never needs them: everything above the threshold is visible).  Branch-free; the caller ANDs with 0x80808080.
struct RangeTest {
	ByteGE vge, vle, gge, gle;
	bool   empty;
	__device__ __forceinline__ void set(unsigned vlo, unsigned vhi, unsigned glo, unsigned ghi)
	{
		empty = vlo > vhi;
		vge.set(empty ? 0u : vlo); vle.set(empty ? 0u : 255u - vhi); gge.set(empty ? 0u : glo); gle.set(empty ? 0u : 255u - ghi);
	}
	template <bool USE_G, bool HI>
	__device__ __forceinline__ unsigned bits(unsigned v, unsigned g) const
	{
		unsigned c = vge.test(v);
		if (HI) c &= vle.test(~v);
		if (USE_G) {
			c &= gge.test(g);
			if (HI) c &= gle.test(~g);
		}
		return c;
	}
};

// Same membership test with a "clean" result: the operand is split once into its low 7 bits and its top bits, the
// selector is 0x80808080 instead of ~0, so every bit but bit 7 of each byte of the result is zero and the result can be
// summed as it is: __dp4a(result, 0x01010101, acc) adds 128 per member byte — one instruction on the fma pipe instead of
// mask + POPC + add on the alu pipe.  Used by the fused voxel count.
struct ByteGE7 {
	unsigned add, sel;
	__device__ __forceinline__ void set(unsigned lo)
	{
		sel = lo <= 128u ? 0x80808080u : 0u;
		add = (lo <= 128u ? 128u - lo : 256u - lo) * 0x01010101u;
	}
	__device__ __forceinline__ unsigned test(unsigned xl, unsigned xh) const
	{
		const unsigned m = xl + add;
		return (m & xh) | ((m | xh) & sel);
	}
};
struct RangeTest7 {
	ByteGE7 vge, vle, gge, gle;
	__device__ __forceinline__ void set(unsigned vlo, unsigned vhi, unsigned glo, unsigned ghi)
	{
		vge.set(vlo); vle.set(255u - vhi); gge.set(glo); gle.set(255u - ghi);
	}
	template <bool USE_G, bool HI>
	__device__ __forceinline__ unsigned bits(unsigned v, unsigned g) const
	{
		const unsigned vl = v & 0x7f7f7f7fu, vh = v & 0x80808080u;
		unsigned       c  = vge.test(vl, vh);
		if (HI) c &= vle.test(vl ^ 0x7f7f7f7fu, vh ^ 0x80808080u);
		if (USE_G) {
			const unsigned gl = g & 0x7f7f7f7fu, gh = g & 0x80808080u;
			c &= gge.test(gl, gh);
			if (HI) c &= gle.test(gl ^ 0x7f7f7f7fu, gh ^ 0x80808080u);
		}
		return c;
	}
};

// ---- TMA-staged streaming variant (BS = 4) ------------------------------------------------------------------
// The register-staged kernel above keeps only 64-128 B per thread in flight and alternates "load" and
// "classify" phases inside every warp, which leaves HBM at ~40 % of its bandwidth.  Here the loads are taken
// out of the compute warps entirely: one producer warp streams the 4x4 rows of a block row (1024 voxels of x
// per task, V and — if used — G) into a ring of shared-memory stages with 1-D bulk TMA copies
// (cp.async.bulk, completion on an mbarrier), 128-160 KB in flight per SM, and two groups of 256 consumer threads
// alternate over the stages: each thread owns one 4-voxel word column = exactly one block of the map, reads its
// 16 (+16) words conflict-free from shared memory, runs the SWAR prefilter + bit-mask lookup and stores one map byte.
// Persistent grid: one CTA per SM.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	unsigned           ok, spins = 0;
	unsigned long long t0 = 0ull;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok)
		             : "r"(smem_u32(bar)), "r"(parity)
		             : "memory");
		// a lost arrival must surface as an error, never as a hung GPU — but only after 20 s of WALL time (%globaltimer), so that
		// time-slicing, a debugger, compute-sanitizer or first-touch page migration cannot trip it (a spin count could)
		if (!ok && (++spins & 1023u) == 0u) {
			unsigned long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if (t0 == 0ull) t0 = now;
			else if (now - t0 > 20000000000ull) __trap();
		}
	} while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

// 3-D tiled TMA: one instruction moves a whole (256 words x 4 rows x 4 slices) box = 16 KB; rows or columns
// outside the volume are zero-filled by the hardware and still count towards the transaction bytes.
__device__ __forceinline__ void tma_load_box(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
	                 smem_u32(smem_dst)),
	             "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
	             : "memory");
}

constexpr int kTmaXC       = 1024;                   // voxels of x per task
constexpr int kTmaRows     = 16;                     // 4 x 4 rows of a block row
constexpr int kTmaConsumers = 512;                   // two groups of 256
constexpr int kTmaThreads  = kTmaConsumers + 32;     // + one producer warp
// The stage count must be EVEN: the two consumer groups take the producer's tasks alternately, so with an even ring every slot
// belongs to one group for the whole launch.  With an odd ring a slot alternates between the groups, and a group that had already
// finished its first two tasks could reach a slot whose FIRST fill (another group's task) had not landed yet — bulk copies complete
// out of order, and on volumes past the TLB reach the first boxes are the slow ones; waiting for an odd phase on a barrier still in
// phase 0 succeeds at once, the group consumed a stage that was not its own and released it a second time (seen as an unspecified
// launch failure on 1024-voxel-wide volumes of 128 MB and more with a gradient transfer function, five stages).
template <bool USE_G> struct TmaCfg {
	static constexpr int    kStages     = USE_G ? 6 : 10;
	static_assert(kStages % 2 == 0, "every ring slot must stay with one consumer group");
	static constexpr int    kStageBytes = kTmaRows * kTmaXC * (USE_G ? 2 : 1);
	static constexpr size_t kSmemBytes  = (size_t) kStages * kStageBytes + 2 * kStages * sizeof(uint64_t) + 16;
};

template <bool USE_G, bool COUNT>
__global__ void __launch_bounds__(kTmaThreads, 1) occupancy_tma_kernel(const __grid_constant__ CUtensorMap map_v,
                                                                      const __grid_constant__ CUtensorMap map_g,
                                                                      const uint2 *__restrict__ mask2, const TFBounds *__restrict__ bounds,
                                                                      uint32_t W, uint32_t H, uint32_t D, uint32_t Wb, uint32_t Hb,
                                                                      uint32_t zb_first, uint32_t zb_count, uint8_t *__restrict__ O,
                                                                      unsigned long long *__restrict__ count)
{
	constexpr int BS = 4, S = TmaCfg<USE_G>::kStages;
	constexpr int kStageBytes = TmaCfg<USE_G>::kStageBytes;
	extern __shared__ __align__(128) uint8_t s_dyn[];
	__shared__ uint2              s_mask[kMaskWords];
	__shared__ unsigned long long s_c[kTmaConsumers / 32];
	__shared__ uint4              meta[TmaCfg<USE_G>::kStages];
	uint8_t  *stage_base = s_dyn;
	uint64_t *full  = reinterpret_cast<uint64_t *>(s_dyn + (size_t) S * kStageBytes);
	uint64_t *empty = full + S;

	if (threadIdx.x == 0) {
		for (int i = 0; i < S; ++i) {
			mbar_init(&full[i], 1);                  // the producer's arrive.expect_tx (+ the bytes)
			mbar_init(&empty[i], 256 / 32);          // one arrival per consumer warp of the group that read the stage
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = threadIdx.x; i < kMaskWords; i += blockDim.x) s_mask[i] = mask2[i];
	__syncthreads();

	const uint32_t xchunks = (W + kTmaXC - 1) / kTmaXC;
	const uint32_t ntasks  = xchunks * Hb * zb_count;        // < 2^32 (checked by the launcher)
	const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (warp == kTmaConsumers / 32) {
		// ---- producer: one lane issues one box load per array per task ----
		if (lane == 0) {
			uint32_t n = 0;
			for (uint32_t task = blockIdx.x; task < ntasks; task += gridDim.x, ++n) {
				const int      st = (int) (n % S);
				const unsigned ph = (n / S) & 1u;
				if (n >= (uint32_t) S) mbar_wait(&empty[st], ph ^ 1u);        // the consumers released this slot
				const uint32_t r  = task / xchunks, xc = task - r * xchunks;
				const uint32_t bq = r / Hb, by = r - bq * Hb, bz = zb_first + bq;
				uint8_t       *dst = stage_base + (size_t) st * kStageBytes;
				meta[st] = make_uint4(xc, by, bz, 0u);
				mbar_arrive_expect_tx(&full[st], (unsigned) kStageBytes);
				tma_load_box(dst, &map_v, (int) (xc * (kTmaXC / 4)), (int) (by * BS), (int) (bz * BS), &full[st]);
				if (USE_G) tma_load_box(dst + kTmaRows * kTmaXC, &map_g, (int) (xc * (kTmaXC / 4)), (int) (by * BS), (int) (bz * BS), &full[st]);
			}
		}
		return;
	}

	// ---- consumers ----
	const int group = warp >> 3;                 // 0 / 1: stages n with n % 2 == group
	const int t     = threadIdx.x & 255;         // word column inside the task = block inside the chunk
	const TFRange tex = USE_G ? bounds->tex_all : bounds->tex_row255;
	const TFRange ana = USE_G ? bounds->ana_all : bounds->ana_row255;
	RangeTest rt_tex, rt_sure, rt_ana;
	rt_tex.set(tex.v_lo, tex.v_hi, tex.g_lo, tex.g_hi);
	rt_sure.set(tex.v_sure, tex.v_sure < 256u ? 255u : 0u, USE_G ? tex.g_sure : 0u, 255u);
	rt_ana.set(ana.v_lo, ana.v_hi, ana.g_lo, ana.g_hi);
	const bool tex_exact = tex.exact != 0u, ana_exact = ana.exact != 0u;
	// Fused count, common case (the built-in ramp transfer function): the analytically visible set is exactly a byte
	// rectangle that contains every texel the texture shows.  Then ONE clean range test per word both counts the voxels
	// (dp4a) and tells whether the block can be occupied at all; the texture classification below only runs for the few
	// columns that saw a candidate.
	RangeTest7 rt_ana7;
	rt_ana7.set(rt_ana.empty ? 0u : ana.v_lo, rt_ana.empty ? 255u : ana.v_hi, rt_ana.empty ? 0u : ana.g_lo, rt_ana.empty ? 255u : ana.g_hi);
	const bool count_first = COUNT && ana_exact && !rt_ana.empty &&
	                         (rt_tex.empty || (ana.v_lo <= tex.v_lo && tex.v_hi <= ana.v_hi && (!USE_G || (ana.g_lo <= tex.g_lo && tex.g_hi <= ana.g_hi))));
	// upper tests are only needed when some range stops below 255 (never with the built-in ramp transfer function)
	const bool need_hi = (!rt_tex.empty && (tex.v_hi < 255u || (USE_G && tex.g_hi < 255u))) ||
	                     (COUNT && !rt_ana.empty && (ana.v_hi < 255u || (USE_G && ana.g_hi < 255u)));

	unsigned long long local_count = 0;
	int      st = group;
	unsigned ph = 0;
	for (uint32_t task = blockIdx.x + (uint32_t) group * gridDim.x; task < ntasks; task += 2 * gridDim.x) {
		mbar_wait(&full[st], ph);
		const uint4    m  = meta[st];        // written by the producer before it armed the barrier: x chunk, block row, block slice
		const uint32_t by = m.y, bz = m.z;
		const uint32_t x  = m.x * kTmaXC + (uint32_t) t * 4;
		const unsigned *sv = reinterpret_cast<const unsigned *>(stage_base + (size_t) st * kStageBytes) + t;
		const unsigned *sg = sv + kTmaRows * kTmaXC / 4;
		// words are read straight from the stage (conflict-free: consecutive threads, consecutive words); the slot goes back
		// to the producer once this warp has classified its columns
		auto vw = [&](int row) { return sv[row * (kTmaXC / 4)]; };
		auto gw = [&](int row) { return USE_G ? sg[row * (kTmaXC / 4)] : 0u; };
		const int      st_used = st;
		st += 2;
		if (st >= S) { st -= S; ph ^= 1u; }
		if (x >= W) {
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty[st_used]);
			continue;
		}
		// rows beyond the volume (zero-filled by the TMA) must not be classified: bit r of `rows` = row r exists
		const uint32_t ny = min((uint32_t) BS, H - by * BS), nz = min((uint32_t) BS, D - bz * BS);
		const unsigned rows = (ny == BS && nz == BS) ? 0xffffu : (((1u << ny) - 1u) * 0x1111u) & ((1u << (4 * nz)) - 1u);
		bool     hit = false;
		unsigned cnt = 0;
		auto classify = [&](auto hi_tag, auto edge_tag) {
			constexpr bool HI = decltype(hi_tag)::value, EDGE = decltype(edge_tag)::value;
			// With the fused count ahead and a texture whose visible set is not a rectangle (gradient transfer functions), the counted
			// candidates stand in for the texture's bounding rectangle — they contain it (count_first) — and the second pass over the
			// column only looks for the sure rectangle: one range test per word less (the count used to double the pass on such TFs).
			const bool ana_for_tex = COUNT && count_first && !tex_exact;
			unsigned   acc_a = 0;
			if (COUNT && count_first) {
				unsigned c128 = 0;        // 128 x (visible voxels of this column)
#pragma unroll
				for (int row = 0; row < kTmaRows; ++row) {
					if (EDGE && !((rows >> row) & 1u)) continue;
					const unsigned a = rt_ana7.bits<USE_G, HI>(vw(row), gw(row));
					c128 = __dp4a(a, 0x01010101u, c128);
					acc_a |= a;
				}
				cnt = c128 >> 7;
				if (c128 == 0u || rt_tex.empty) return;        // no analytically visible voxel -> no texture-visible one either
			}
			unsigned acc_t = 0, acc_s = 0;
			if (ana_for_tex) {
				acc_t = acc_a;
#pragma unroll
				for (int row = 0; row < kTmaRows; ++row) {
					if (EDGE && !((rows >> row) & 1u)) continue;
					acc_s |= rt_sure.bits<USE_G, false>(vw(row), gw(row));
				}
			} else {
#pragma unroll
				for (int row = 0; row < kTmaRows; ++row) {
					if (EDGE && !((rows >> row) & 1u)) continue;
					acc_t |= rt_tex.bits<USE_G, HI>(vw(row), gw(row));
					if (!tex_exact) acc_s |= rt_sure.bits<USE_G, false>(vw(row), gw(row));
				}
			}
			acc_t &= 0x80808080u; acc_s &= 0x80808080u;
			if (rt_tex.empty) acc_t = 0u;
			if (tex_exact) hit = acc_t != 0u;
			else {
				hit = acc_s != 0u && !rt_sure.empty;
				if (!hit && acc_t) {        // candidates only in the sliver between the sure and the outer rectangle: look them up
					unsigned dummy = 0;
#pragma unroll
					for (int row = 0; row < kTmaRows; ++row) {
						if (EDGE && !((rows >> row) & 1u)) continue;
						const unsigned cand = rt_tex.bits<USE_G, HI>(vw(row), gw(row)) & 0x80808080u;
						if (cand && !hit) classify_word<USE_G, false>(vw(row), gw(row), cand, s_mask, hit, dummy);
					}
				}
			}
			if (COUNT && !count_first && !rt_ana.empty) {
#pragma unroll
				for (int row = 0; row < kTmaRows; ++row) {
					if (EDGE && !((rows >> row) & 1u)) continue;
					const unsigned cand = rt_ana.bits<USE_G, HI>(vw(row), gw(row)) & 0x80808080u;
					if (ana_exact) cnt += __popc(cand);
					else if (cand) {
						bool h = false;
						classify_word<USE_G, true>(vw(row), gw(row), cand, s_mask, h, cnt);
					}
				}
			}
		};
		if (rows == 0xffffu) {
			if (need_hi) classify(std::true_type{}, std::false_type{});
			else classify(std::false_type{}, std::false_type{});
		} else {
			classify(std::true_type{}, std::true_type{});
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&empty[st_used]);
		if (O) O[((size_t) bz * Hb + by) * Wb + (x >> 2)] = hit ? 0 : 255;
		if (COUNT) local_count += cnt;
	}

	if (COUNT) {
		unsigned long long c = local_count;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
		if (lane == 0) s_c[warp] = c;
		// consumers only (the producer warp has left): named barrier 1 over the 512 consumer threads
		asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumers) : "memory");
		if (threadIdx.x == 0) {
			unsigned long long tsum = 0;
			for (int i = 0; i < kTmaConsumers / 32; ++i) tsum += s_c[i];
			if (tsum) atomicAdd(count, tsum);
		}
	}
}

template <bool USE_G, bool COUNT>
static int launch_tma_inst(vkv_volume *vol, uint8_t *O, uint32_t zb_first, uint32_t zb_count, unsigned long long *count_dev, cudaStream_t s)
{
	static PerDeviceOnce configured;
	if (configured.first(vol->ctx->device)) {
		VKV_CUDA_CHECK(cudaFuncSetAttribute(occupancy_tma_kernel<USE_G, COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                                    (int) TmaCfg<USE_G>::kSmemBytes));
	}
	occupancy_tma_kernel<USE_G, COUNT><<<vol->ctx->sm_count, kTmaThreads, TmaCfg<USE_G>::kSmemBytes, s>>>(
	    *reinterpret_cast<const CUtensorMap *>(vol->tmap_V), *reinterpret_cast<const CUtensorMap *>(USE_G ? vol->tmap_G : vol->tmap_V), vol->d_mask2,
	    vol->d_bounds, vol->dim[0], vol->dim[1], vol->dim[2], vol->dim_b[0], vol->dim_b[1], zb_first, zb_count, O,
	    count_dev);
	VKV_LAUNCHED();
	return VKV_OK;
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


// ---- on-the-fly gradient variant (PRECOMPUTED_GRADIENT undefined: `--gradient_test`, volumes created with
// use_precomputed_gradient = 0) -------------------------------------------------------------------------------
// shaders/get_gradient_compute.glsl:12-20 evaluated per voxel in the shader's fp32 operation order (the library is
// compiled without contraction), then the texture lookup of occupancy_map.comp (row = nearest texel of the FLOAT
// gradient, not of a stored byte) and the analytic TF of occupied_voxel_count.comp.  No gradient map exists in this
// mode, so the pass reads V only: 4 clamped taps per voxel, served by L1/L2 (each byte of V is touched by 4 voxels
// of 4 neighbouring rows).  One thread per voxel; the map slab is pre-set to EMPTY (255) and hits store OCCUPIED (0),
// one store per run of hit lanes that share a block.
template <bool COUNT>
__global__ void __launch_bounds__(256) occupancy_otf_kernel(const uint8_t *__restrict__ V, const uint2 *__restrict__ mask2, uint32_t W, uint32_t H,
                                                           uint32_t D, uint32_t Wb, uint32_t Hb, uint32_t bsx, uint32_t bsy, uint32_t bsz,
                                                           uint32_t z_first, uint32_t z_end, vkv_transfer_function_uniform tfu,
                                                           uint8_t *__restrict__ O, unsigned long long *__restrict__ count)
{
	__shared__ uint2              s_mask[kMaskWords];
	__shared__ unsigned long long s_c[8];
	for (int i = threadIdx.x; i < kMaskWords; i += blockDim.x) s_mask[i] = mask2[i];
	__syncthreads();
	const uint32_t     xchunks = (W + 255u) / 256u;
	const uint64_t     ntasks  = (uint64_t) xchunks * H * (z_end - z_first);
	const int          lane    = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int          mx = (int) W - 1, my = (int) H - 1, mz = (int) D - 1;
	const size_t       WH = (size_t) W * H;
	unsigned long long local = 0;
	for (uint64_t task = blockIdx.x; task < ntasks; task += gridDim.x) {
		const uint32_t xc = (uint32_t) (task % xchunks);
		const uint64_t r  = task / xchunks;
		const int      y = (int) (r % H), z = (int) z_first + (int) (r / H);
		const int      x = (int) (xc * 256u + threadIdx.x);
		bool           hit = false;
		if (x <= mx) {
			const int    xm = max(x - 1, 0), xp = min(x + 1, mx), ym = max(y - 1, 0), yp = min(y + 1, my), zm = max(z - 1, 0), zp = min(z + 1, mz);
			const float  intensity = (float) __ldg(V + (size_t) z * WH + (size_t) y * W + x) / 255.0f;
			// k = (1,-1): taps (+,-,-), (-,-,+), (-,+,-), (+,+,+)
			const float a = (float) __ldg(V + (size_t) zm * WH + (size_t) ym * W + xp) / 255.0f;
			const float b = (float) __ldg(V + (size_t) zp * WH + (size_t) ym * W + xm) / 255.0f;
			const float c = (float) __ldg(V + (size_t) zm * WH + (size_t) yp * W + xm) / 255.0f;
			const float d = (float) __ldg(V + (size_t) zp * WH + (size_t) yp * W + xp) / 255.0f;
			const float gx  = 0.25f * (((1.0f * a + -1.0f * b) + -1.0f * c) + 1.0f * d);
			const float gy  = 0.25f * (((-1.0f * a + -1.0f * b) + 1.0f * c) + 1.0f * d);
			const float gz  = 0.25f * (((-1.0f * a + 1.0f * b) + -1.0f * c) + 1.0f * d);
			const float len = sqrtf((gx * gx + gy * gy) + gz * gz);
			const float gradient = fminf(fmaxf(len * tfu.grad_magnitude_modifier, 0.0f), 1.0f);
			// texture(transfer_function, vec2(intensity, gradient)).a > 0: nearest texel, clamp to edge
			const int      ti = min(max((int) floorf(intensity * 256.0f), 0), 255), tg = min(max((int) floorf(gradient * 256.0f), 0), 255);
			hit               = (s_mask[tg * 8 + (ti >> 5)].x >> (ti & 31)) & 1u;
			if (COUNT) {
				const float aI = fminf(fmaxf((intensity - tfu.intensity_min) * tfu.intensity_range_inv, 0.0f), 1.0f);
				const float aG = fminf(fmaxf((gradient - tfu.gradient_min) * tfu.gradient_range_inv, 0.0f), 1.0f);
				local += (aI * aG > 0.0f) ? 1u : 0u;
			}
		}
		if (O) {
			const unsigned hits  = __ballot_sync(0xffffffffu, hit);
			const bool     first = lane == 0 || !((hits >> (lane - 1)) & 1u) || (uint32_t) (x - 1) / bsx != (uint32_t) x / bsx;
			if (hit && first) O[((size_t) ((uint32_t) z / bsz) * Hb + (uint32_t) y / bsy) * Wb + (uint32_t) x / bsx] = 0;
		}
	}
	if (COUNT) {
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

static int launch_occupancy_otf(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, bool count, uint8_t *O, uint32_t zb_first,
                                uint32_t zb_count, unsigned long long *count_dev, cudaStream_t s)
{
	const uint32_t z0 = zb_first * vol->bs[2], z1 = std::min<uint32_t>(vol->dim[2], (zb_first + zb_count) * vol->bs[2]);
	if (O) VKV_CUDA_CHECK(cudaMemsetAsync(O + (size_t) zb_first * vol->dim_b[0] * vol->dim_b[1], 255, (size_t) zb_count * vol->dim_b[0] * vol->dim_b[1], s));
	if (z1 <= z0) return VKV_OK;
	const int grid = vol->ctx->sm_count * 8;
#define VKV_OTF_ARGS vol->d_V, vol->d_mask2, vol->dim[0], vol->dim[1], vol->dim[2], vol->dim_b[0], vol->dim_b[1], vol->bs[0], vol->bs[1], vol->bs[2], z0, z1, *tfu, O, count_dev
	if (count) occupancy_otf_kernel<true><<<grid, 256, 0, s>>>(VKV_OTF_ARGS);
	else occupancy_otf_kernel<false><<<grid, 256, 0, s>>>(VKV_OTF_ARGS);
#undef VKV_OTF_ARGS
	VKV_LAUNCHED();
	return VKV_OK;
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

// Tensor maps over the linear V / G copies viewed as (W/4) x H x D 32-bit words, box 256 x 4 x 4 (SWIZZLE_NONE):
// the TMA-staged occupancy kernel's loads.  Needs W % 16 == 0; failure just leaves tmap_ok false (register path).
int make_volume_tensor_maps(vkv_volume *vol)
{
	vol->tmap_ok = false;
	static_assert(sizeof(CUtensorMap) == sizeof(vol->tmap_V), "CUtensorMap size");
	if (vol->dim[0] % 16 != 0) return VKV_OK;
	typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
	                              const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static encode_fn encode = nullptr;
	if (!encode) {
		void                            *fn = nullptr;
		cudaDriverEntryPointQueryResult  q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
			cudaGetLastError();
			return VKV_OK;
		}
		encode = reinterpret_cast<encode_fn>(fn);
	}
	const cuuint64_t gdim[3]    = {vol->dim[0] / 4, vol->dim[1], vol->dim[2]};
	const cuuint64_t gstride[2] = {(cuuint64_t) vol->dim[0], (cuuint64_t) vol->dim[0] * vol->dim[1]};
	const cuuint32_t box[3]     = {kTmaXC / 4, 4, 4};
	const cuuint32_t estr[3]    = {1, 1, 1};
	uint8_t *bases[2] = {vol->d_V, vol->d_G};
	unsigned char *maps[2] = {vol->tmap_V, vol->tmap_G};
	for (int i = 0; i < 2; ++i) {
		if (!bases[i]) continue;
		if (encode(reinterpret_cast<CUtensorMap *>(maps[i]), CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, bases[i], gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
			return VKV_OK;
	}
	vol->tmap_ok = true;
	return VKV_OK;
}

int launch_occupancy(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, bool count, uint8_t *O, uint32_t zb_first, uint32_t zb_count,
                     unsigned long long *count_dev, cudaStream_t s)
{
	if (zb_count == 0) return VKV_OK;
	const bool use_gradient = tfu->use_gradient != 0;
	// no gradient map in this volume: gradients are evaluated per voxel (get_gradient_compute.glsl:12-20)
	if (use_gradient && !vol->precomputed_gradient) return launch_occupancy_otf(vol, tfu, count, O, zb_first, zb_count, count_dev, s);
	// gradient bytes are only read when the TF uses them AND a precomputed map exists
	const bool use_g = use_gradient;
	const int  grid  = vol->ctx->sm_count * 8;        // persistent: 8 CTAs/SM x 256 threads = full occupancy
	const bool cubic = vol->bs[0] == vol->bs[1] && vol->bs[1] == vol->bs[2];
	const uint64_t fast_tasks = (uint64_t) ((vol->dim[0] + 63) / 64) * vol->dim_b[1] * zb_count;        // upper bound over BS
	const bool fast  = fast_tasks < (1ull << 32) && cubic && (vol->bs[0] == 2 || vol->bs[0] == 4 || vol->bs[0] == 8) && vol->dim[0] % 16 == 0 &&
	                  vol->dim_b[0] * vol->bs[0] == vol->dim[0] && (reinterpret_cast<uintptr_t>(vol->d_V) % 16 == 0);
	if (fast && vol->bs[0] == 4 && vol->tmap_ok && !getenv("VKV_OCC_NO_TMA")) {
		if (use_g && count) return launch_tma_inst<true, true>(vol, O, zb_first, zb_count, count_dev, s);
		if (use_g) return launch_tma_inst<true, false>(vol, O, zb_first, zb_count, count_dev, s);
		if (count) return launch_tma_inst<false, true>(vol, O, zb_first, zb_count, count_dev, s);
		return launch_tma_inst<false, false>(vol, O, zb_first, zb_count, count_dev, s);
	}
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
