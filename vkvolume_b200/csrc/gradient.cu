// gradient.cu — K1 gradient-magnitude map.
//
// Replaces shaders/gradient_map.comp + shaders/get_gradient_compute.glsl (4-tap tetrahedral
// gradient, one invocation per voxel with four clamped imageLoads) and
// ComputeGradientMap::compute (src/compute_gradient_map.cpp:57-81).
//
// B200 design (kernels in the order launch_gradient prefers them):
//   * gradient_walk_kernel — the column walk (round 2): a thread owns a 16-voxel chunk of x at one z and walks every second row of a
//     y segment (the rows loaded for row y serve row y + 2 again: two 16-byte loads per step), integer formulation of |g| on dp4a,
//     exact ties resolved from a shared-memory stash before the row is stored, warps shaped for the texture array's surface stores;
//   * gradient_flat_kernel / gradient_int_kernel — round 1's integer kernels (flat persistent walk, one row task per warp): fallbacks
//     for extents whose indices do not fit the walk's packed queue entries, and the A/B references (VKV_GRAD_FLAT, VKV_GRAD_V1);
//   * gradient_vec16_kernel — the shader's operation order in fp32 for every voxel (grad_magnitude_modifier != 1, VKV_GRAD_FP32);
//   * gradient_scalar_kernel — any extents (W % 16 != 0).
// Every kernel stores the byte the CPU oracle stores: fp32 without contraction, IEEE division and square root wherever the fp32
// chain is evaluated, UNORM decode b / 255 from a 256-entry table built with a true division (the byte feeds the occupancy LUT).
// Algorithmic bytes: read N + write N = 2 B/voxel (+ N for the copy of the map in the texture array the ray caster samples).
#include <algorithm>
#include <cstdlib>

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

// ---- integer formulation (grad_magnitude_modifier == 1, the only value the reference ever passes) -------------
// With A, B, C, D the four tap bytes, 255 * |g| = sqrt(S) / 4 where S = sx^2 + sy^2 + sz^2 is an INTEGER:
//     (sx, sy, sz) = M (A, B, C, D),  M^T M = 4 I - J   =>   S = 4 (A^2 + B^2 + C^2 + D^2) - (A + B + C + D)^2,
// two dp4a instructions on the packed taps.  The stored byte is rint(sqrt(S) / 4); the reference's fp32 chain
// (UNORM decode, three adds, 0.25, length, * 255, round) carries an absolute error below 1.5e-4 in 255 |g|, so whenever
// sqrt(S) / 4 is further than 1e-3 from a rounding boundary n + 1/2 the byte is decided by S alone.  The remaining
// voxels — in practice only exact ties S = (4n + 2)^2, 3-5 % of a volume — are queued per warp in shared memory and
// evaluated by all 32 lanes with the shader's exact fp32 operation order (the old path), then patched into the
// staged output row.  Byte-identical to the fp32 kernel (and to the CPU oracle) at a quarter of the instructions.
__device__ __forceinline__ unsigned prmt_(unsigned a, unsigned b, unsigned sel)
{
	unsigned r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}

constexpr int kGradQueue = 512;        // worst case: every voxel of a warp iteration is a tie

// Work decomposition: a task is one row (y, z) x one group of 32 chunks (512 voxels of x), numbered with the chunk
// group fastest; warp w of CTA b takes task 8 b + w, so at any moment the grid reads and writes a compact window of the
// volume (DRAM pages stay open, the (y +- 1, z +- 1) rows are shared through L2) and the (group, y, z) decode is two
// 32-bit divisions per warp task instead of 64-bit ones per thread.
// SURF: also write the 16 result bytes into the cudaArray the ray caster samples (the map exists twice: linear for the
// occupancy pass, array for the texture unit) — one surface store instead of a second pass over the map.
template <bool SURF>
__global__ void __launch_bounds__(256) gradient_int_kernel(const uint8_t *__restrict__ V, uint8_t *__restrict__ G, cudaSurfaceObject_t surf,
                                                          uint32_t W, uint32_t H, uint32_t D)
{
	__shared__ float          s_lut[256];
	__shared__ __align__(16) unsigned s_w[8][4][32 * 4];        // taps (A | B << 8 | C << 16 | D << 24) of the warp's 512 voxels: [j][lane][i]
	__shared__ unsigned short s_qp[8][kGradQueue];              // queued voxels: lane * 16 + j * 4 + i
	__shared__ __align__(16) unsigned char s_out[8][32 * 16];
	__shared__ unsigned       s_qn[8];
	s_lut[threadIdx.x] = (float) threadIdx.x / 255.0f;        // UNORM decode, exactly as imageLoad
	if (threadIdx.x < 8) s_qn[threadIdx.x] = 0u;
	__syncthreads();
	const uint32_t nchunks = W / 16, ngroups = (nchunks + 31) / 32;
	const uint32_t ntasks  = ngroups * H * D;        // < 2^32 (checked by the launcher)
	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t task = blockIdx.x * 8u + warp;
	if (task >= ntasks) return;
	const uint32_t row = task / ngroups, grp = task - row * ngroups;
	const uint32_t z = row / H, y = row - z * H;
	{
		const uint32_t chunk  = grp * 32 + lane;
		const bool     active = chunk < nchunks;
		const uint32_t cch    = active ? chunk : nchunks - 1;        // idle lanes mirror the last chunk (their loads stay in bounds)
		const uint32_t ym = y > 0 ? y - 1 : 0, yp = y + 1 < H ? y + 1 : H - 1;
		const uint32_t zm = z > 0 ? z - 1 : 0, zp = z + 1 < D ? z + 1 : D - 1;
		const uint32_t x = cch * 16;
		const size_t   r_mm = ((size_t) zm * H + ym) * W, r_pm = ((size_t) zp * H + ym) * W, r_mp = ((size_t) zm * H + yp) * W,
		               r_pp = ((size_t) zp * H + yp) * W;
		// rows: A = (y-1, z-1) read at x+1 | B = (y-1, z+1) at x-1 | C = (y+1, z-1) at x-1 | E = (y+1, z+1) at x+1
		const uint8_t *rowp[4] = {V + r_mm + x, V + r_pm + x, V + r_mp + x, V + r_pp + x};
		const bool     first = cch == 0, last = cch == nchunks - 1;
		unsigned win[4][4];        // the 16-byte window of each row, already shifted by its x offset
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint4    q    = __ldg(reinterpret_cast<const uint4 *>(rowp[k]));
			const unsigned w[4] = {q.x, q.y, q.z, q.w};
			const bool     plus = (k == 0 || k == 3);        // tap at x+1 (else x-1)
			if (plus) {
				// byte x+16 (clamped to W-1) in bits 0..7: from the next lane, from memory for the group's last lane
				const unsigned down = __shfl_down_sync(0xffffffffu, w[0], 1);
				unsigned       edge = 0;
				if (lane == 31 && !last) edge = rowp[k][16];
				const unsigned right = last ? (w[3] >> 24) : (lane == 31 ? edge : down);
				win[k][0] = __funnelshift_r(w[0], w[1], 8);
				win[k][1] = __funnelshift_r(w[1], w[2], 8);
				win[k][2] = __funnelshift_r(w[2], w[3], 8);
				win[k][3] = __funnelshift_r(w[3], right, 8);
			} else {
				// byte x-1 (clamped to 0) in bits 24..31
				const unsigned up   = __shfl_up_sync(0xffffffffu, w[3], 1);
				unsigned       edge = 0;
				if (lane == 0 && !first) edge = ((unsigned) *(rowp[k] - 1)) << 24;
				const unsigned left = first ? (w[0] << 24) : (lane == 0 ? edge : up);
				win[k][0] = __funnelshift_r(left, w[0], 24);
				win[k][1] = __funnelshift_r(w[0], w[1], 24);
				win[k][2] = __funnelshift_r(w[1], w[2], 24);
				win[k][3] = __funnelshift_r(w[2], w[3], 24);
			}
		}
		unsigned out[4], flags = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			// 4x4 byte transpose: taps of voxel i of this word group -> one word (A, B, C, D)
			const unsigned t0 = prmt_(win[0][j], win[1][j], 0x5140u), t1 = prmt_(win[0][j], win[1][j], 0x7362u);
			const unsigned u0 = prmt_(win[2][j], win[3][j], 0x5140u), u1 = prmt_(win[2][j], win[3][j], 0x7362u);
			const unsigned wv[4] = {prmt_(t0, u0, 0x5410u), prmt_(t0, u0, 0x7632u), prmt_(t1, u1, 0x5410u), prmt_(t1, u1, 0x7632u)};
			*reinterpret_cast<uint4 *>(&s_w[warp][j][lane * 4]) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
			float yv[4];
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const unsigned q  = __dp4a(wv[i], wv[i], 0u);                // A^2 + B^2 + C^2 + D^2
				const unsigned sm = __dp4a(wv[i], 0x01010101u, 0u);          // A + B + C + D
				const unsigned S  = 4u * q - sm * sm;                        // < 2^20
				const float    f  = __uint_as_float(S | 0x4b000000u) - 8388608.0f;        // exact int -> float
				float rt;
				asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(f));
				const float yy = __fmaf_rn(rt, 0.25f, 8388608.0f);           // low mantissa bits = rint(sqrt(S) / 4)
				const float dd = __fmaf_rn(rt, 0.25f, -(yy - 8388608.0f));   // distance from that integer, in [-1/2, 1/2]
				yv[i]          = yy;
				// too close to n + 1/2 (the sign bit of 0.499 - |dd|): decided by the fp32 chain.  Bit 15 - (4 j + i) of flags.
				flags = __funnelshift_l(__float_as_uint(0.499f - fabsf(dd)), flags, 1);
			}
			const unsigned p01 = prmt_(__float_as_uint(yv[0]), __float_as_uint(yv[1]), 0x0040u);
			const unsigned p23 = prmt_(__float_as_uint(yv[2]), __float_as_uint(yv[3]), 0x0040u);
			out[j]             = prmt_(p01, p23, 0x5410u);
		}
		*reinterpret_cast<uint4 *>(&s_out[warp][lane * 16]) = make_uint4(out[0], out[1], out[2], out[3]);
		flags &= 0xffffu;
		while (flags) {
			const unsigned i    = 15u - (unsigned) (__ffs(flags) - 1);
			const unsigned slot = atomicAdd(&s_qn[warp], 1u);
			s_qp[warp][slot]    = (unsigned short) (lane * 16 + i);
			flags &= flags - 1;
		}
		__syncwarp();
		const unsigned nq = s_qn[warp];
		for (unsigned e = lane; e < nq; e += 32) {
			const unsigned p  = s_qp[warp][e];
			const unsigned ln = p >> 4, ji = p & 15u;
			const unsigned w  = s_w[warp][ji >> 2][ln * 4 + (ji & 3u)];
			s_out[warp][p]    = gradient_byte(s_lut[w & 0xffu], s_lut[(w >> 8) & 0xffu], s_lut[(w >> 16) & 0xffu], s_lut[w >> 24], 1.0f);
		}
		__syncwarp();
		const uint4 o = *reinterpret_cast<const uint4 *>(&s_out[warp][lane * 16]);
		if (active) {
			*reinterpret_cast<uint4 *>(G + ((size_t) z * H + y) * W + x) = o;
			if (SURF) surf3Dwrite(o, surf, (int) x, (int) y, (int) z);        // x in bytes
		}
	}
}

// ---- flat persistent formulation of the integer kernel ---------------------------------------------------------
// Same arithmetic as gradient_int_kernel, a third of the instructions per voxel:
//   * work item = one 16-voxel chunk, numbered flat over the volume (no idle lanes for any W % 16 == 0); each thread keeps
//     its (chunk, y, z) and advances it by the grid stride with carries instead of dividing;
//   * the x +- 1 neighbours of a chunk are one extra byte load per row (an L1 hit on the neighbouring lane's line) instead
//     of shuffles with edge cases;
//   * S is built directly as a biased float (0x4b000000 folded into the multiply-add), one MUFU.SQRT, and ONE fused
//     multiply-add 64 sqrt(S) + (2^23 + 128) whose mantissa holds the stored byte in bits 8..15 and, in bits 0..7, the
//     distance from the rounding boundary in 1/256ths: byte 0 == 0 <=> within 1/512 of n + 1/2 (every exact tie
//     S = (4n + 2)^2 and < 0.4 % false positives) <=> the voxel is decided by the shader's fp32 chain instead;
//   * those voxels are queued per WORD (one shared-memory atomic per 4-voxel word with a tie, no per-voxel loop), the queue
//     persists across the thread's chunks and is drained 32 entries at a time by the whole warp, which re-reads the four
//     taps (L1/L2 hits), evaluates the fp32 chain and patches the byte in the linear map and in the texture array.
constexpr int kFlatQueue = 96;        // <= 31 carried + 32 pushed per iteration + 32 re-queued per pass

// One entry per 16-voxel chunk holding ties: bit 8 i + j of `tie` <=> voxel 4 j + i of the chunk needs the fp32 chain.
struct GradQueue {
	unsigned       tie[8][kFlatQueue];
	unsigned       yz[8][kFlatQueue];         // y | z << 16
	unsigned short cx[8][kFlatQueue];         // chunk index within the row
};

// Warp-converged push: the lanes with `push` set append their entry; n is the warp-uniform entry count (kept in a register,
// no shared-memory atomics: slots come from a ballot).
__device__ __forceinline__ void grad_push(GradQueue &q, unsigned &n, int warp, int lane, bool push, unsigned t, unsigned yz, unsigned cx)
{
	const unsigned m = __ballot_sync(0xffffffffu, push);
	if (push) {
		const unsigned slot = n + __popc(m & ((1u << lane) - 1u));
		q.tie[warp][slot] = t, q.yz[warp][slot] = yz, q.cx[warp][slot] = (unsigned short) cx;
	}
	n += __popc(m);
}

template <bool SURF>
__device__ __forceinline__ void grad_drain(GradQueue &q, unsigned &n, const float *s_lut, int warp, int lane, unsigned n_take, const uint8_t *__restrict__ V,
                                           uint8_t *__restrict__ G, cudaSurfaceObject_t surf, uint32_t W, uint32_t H, uint32_t D)
{
	// pops the top n_take (<= 32) entries, one tie each; entries holding more ties push the rest back
	__syncwarp();
	const bool mine = (unsigned) lane < n_take;
	unsigned   t = 0, yz = 0, cx = 0;
	if (mine) {
		const unsigned e = n - n_take + lane;
		t = q.tie[warp][e], yz = q.yz[warp][e], cx = q.cx[warp][e];
	}
	n -= n_take;
	__syncwarp();
	const unsigned rest = t & (t - 1u);
	grad_push(q, n, warp, lane, rest != 0u, rest, yz, cx);
	if (mine) {
		const unsigned pos = __ffs(t) - 1;
		const uint32_t x = cx * 16 + 4 * (pos & 7u) + (pos >> 3), y = yz & 0xffffu, z = yz >> 16;
		const uint32_t xm = x > 0 ? x - 1 : 0, xp = x + 1 < W ? x + 1 : W - 1;
		const uint32_t ym = y > 0 ? y - 1 : 0, yp = y + 1 < H ? y + 1 : H - 1;
		const uint32_t zmH = (z > 0 ? z - 1 : 0) * H, zpH = (z + 1 < D ? z + 1 : D - 1) * H;
		const uint8_t *Vp = V + xp, *Vm = V + xm;
		const float    a = s_lut[Vp[(size_t) (zmH + ym) * W]];
		const float    b = s_lut[Vm[(size_t) (zpH + ym) * W]];
		const float    c = s_lut[Vm[(size_t) (zmH + yp) * W]];
		const float    d = s_lut[Vp[(size_t) (zpH + yp) * W]];
		const unsigned char g = gradient_byte(a, b, c, d, 1.0f);
		G[(size_t) (z * H + y) * W + x] = g;
		if (SURF) surf3Dwrite(g, surf, (int) x, (int) y, (int) z);
	}
}

// The four tap rows of one chunk (+ the neighbouring words of the warp's outer lanes), loaded one iteration ahead.
struct GradRows {
	uint4    qA, qB, qC, qE;
	unsigned eA, eB, eC, eE;
};

__device__ __forceinline__ void grad_load_rows(GradRows &r, const uint8_t *__restrict__ V, uint32_t cx, uint32_t y, uint32_t z, uint32_t W, uint32_t H,
                                               uint32_t D, uint32_t nchunks, int lane)
{
	const uint32_t ym = y > 0 ? y - 1 : 0, yp = y + 1 < H ? y + 1 : H - 1;
	const uint32_t zmH = (z > 0 ? z - 1 : 0) * H, zpH = (z + 1 < D ? z + 1 : D - 1) * H;
	const uint8_t *Vx = V + cx * 16;
	// rows: A = (y-1, z-1) read at x+1 | B = (y-1, z+1) at x-1 | C = (y+1, z-1) at x-1 | E = (y+1, z+1) at x+1
	const uint8_t *pA = Vx + (size_t) (zmH + ym) * W, *pB = Vx + (size_t) (zpH + ym) * W;
	const uint8_t *pC = Vx + (size_t) (zmH + yp) * W, *pE = Vx + (size_t) (zpH + yp) * W;
	r.qA = __ldg(reinterpret_cast<const uint4 *>(pA)), r.qB = __ldg(reinterpret_cast<const uint4 *>(pB));
	r.qC = __ldg(reinterpret_cast<const uint4 *>(pC)), r.qE = __ldg(reinterpret_cast<const uint4 *>(pE));
	r.eA = r.eB = r.eC = r.eE = 0;
	// the words next to the warp's 512-voxel run come from memory unless they lie in another row (never dereferenced:
	// the edge voxel is clamped to itself instead); all other lanes get them from the adjacent lane by shuffle
	if (lane == 31 && cx != nchunks - 1) r.eA = __ldg(reinterpret_cast<const unsigned *>(pA + 16)), r.eE = __ldg(reinterpret_cast<const unsigned *>(pE + 16));
	if (lane == 0 && cx != 0) r.eB = __ldg(reinterpret_cast<const unsigned *>(pB - 4)), r.eC = __ldg(reinterpret_cast<const unsigned *>(pC - 4));
}

template <bool SURF>
__global__ void __launch_bounds__(256) gradient_flat_kernel(const uint8_t *__restrict__ V, uint8_t *__restrict__ G, cudaSurfaceObject_t surf,
                                                           uint32_t W, uint32_t H, uint32_t D, uint32_t niter, uint32_t dcx, uint32_t dy, uint32_t dz)
{
	__shared__ float     s_lut[256];
	__shared__ GradQueue q;
	s_lut[threadIdx.x] = (float) threadIdx.x / 255.0f;        // UNORM decode, exactly as imageLoad
	__syncthreads();
	const uint32_t nchunks = W / 16;
	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned       qn = 0;        // entries in this warp's queue (warp-uniform)
	// this thread's first chunk; later ones are (dcx, dy, dz) further on (the grid stride, decomposed by the launcher)
	uint32_t cx, y, z;
	{
		const uint32_t gid = blockIdx.x * 256u + threadIdx.x;
		const uint32_t r   = gid / nchunks;
		cx = gid - r * nchunks, z = r / H, y = r - z * H;
	}
	// lanes past the end of the volume (last iteration only) mirror its last chunk and store nothing
	GradRows cur;
	grad_load_rows(cur, V, z < D ? cx : nchunks - 1, z < D ? y : H - 1, z < D ? z : D - 1, W, H, D, nchunks, lane);
	for (uint32_t it = 0; it < niter; ++it) {
		// this thread's next chunk, and its rows requested before the current ones are consumed
		uint32_t ncx = cx + dcx;
		const bool c1 = ncx >= nchunks;
		ncx -= c1 ? nchunks : 0u;
		uint32_t ny = y + dy + (c1 ? 1u : 0u);
		const bool c2 = ny >= H;
		ny -= c2 ? H : 0u;
		const uint32_t nz = z + dz + (c2 ? 1u : 0u);
		GradRows       nxt;
		if (it + 1 < niter) grad_load_rows(nxt, V, nz < D ? ncx : nchunks - 1, nz < D ? ny : H - 1, nz < D ? nz : D - 1, W, H, D, nchunks, lane);
		const bool active = z < D;
		{
			const uint32_t cxe   = active ? cx : nchunks - 1;
			const bool     first = cxe == 0, last = cxe == nchunks - 1;
			const uint4    qA = cur.qA, qB = cur.qB, qC = cur.qC, qE = cur.qE;
			unsigned       eA = __shfl_down_sync(0xffffffffu, qA.x, 1), eE = __shfl_down_sync(0xffffffffu, qE.x, 1);
			unsigned       eB = __shfl_up_sync(0xffffffffu, qB.w, 1), eC = __shfl_up_sync(0xffffffffu, qC.w, 1);
			if (lane == 31) eA = cur.eA, eE = cur.eE;
			if (lane == 0) eB = cur.eB, eC = cur.eC;
			const unsigned sel_p = last ? 0x3321u : 0x4321u, sel_m = first ? 0x6544u : 0x6543u;
			// the 16-byte window of each row, shifted by its x offset
			const unsigned wA[4] = {__funnelshift_r(qA.x, qA.y, 8), __funnelshift_r(qA.y, qA.z, 8), __funnelshift_r(qA.z, qA.w, 8), prmt_(qA.w, eA, sel_p)};
			const unsigned wE[4] = {__funnelshift_r(qE.x, qE.y, 8), __funnelshift_r(qE.y, qE.z, 8), __funnelshift_r(qE.z, qE.w, 8), prmt_(qE.w, eE, sel_p)};
			const unsigned wB[4] = {prmt_(eB, qB.x, sel_m), __funnelshift_r(qB.x, qB.y, 24), __funnelshift_r(qB.y, qB.z, 24), __funnelshift_r(qB.z, qB.w, 24)};
			const unsigned wC[4] = {prmt_(eC, qC.x, sel_m), __funnelshift_r(qC.x, qC.y, 24), __funnelshift_r(qC.y, qC.z, 24), __funnelshift_r(qC.z, qC.w, 24)};
			unsigned out[4], tie[4];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				// 4x4 byte transpose: taps of voxel i of this word -> one word (A, B, C, D)
				const unsigned t0 = prmt_(wA[j], wB[j], 0x5140u), t1 = prmt_(wA[j], wB[j], 0x7362u);
				const unsigned u0 = prmt_(wC[j], wE[j], 0x5140u), u1 = prmt_(wC[j], wE[j], 0x7362u);
				const unsigned wv[4] = {prmt_(t0, u0, 0x5410u), prmt_(t0, u0, 0x7632u), prmt_(t1, u1, 0x5410u), prmt_(t1, u1, 0x7632u)};
				unsigned yv[4];
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const unsigned qq = __dp4a(wv[i], wv[i], 0u);                // A^2 + B^2 + C^2 + D^2
					const unsigned sm = __dp4a(wv[i], 0x01010101u, 0u);          // A + B + C + D
					unsigned       X, Sb;                                       // Sb = bits of the float 2^23 + S   (S = 4 qq - sm^2 < 2^18)
					asm("mad.lo.u32 %0, %1, %1, 0xb5000000;" : "=r"(X) : "r"(sm));        // sm^2 - 0x4b000000
					asm("{.reg .u32 t; shl.b32 t, %1, 2; sub.u32 %0, t, %2;}" : "=r"(Sb) : "r"(qq), "r"(X));
					const float f = __uint_as_float(Sb) - 8388608.0f;           // S, exactly
					float       rt;
					asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(f));
					yv[i] = __float_as_uint(__fmaf_rn(rt, 64.0f, 8388736.0f));   // low 16 mantissa bits = rint(64 sqrt(S)) + 128
				}
				const unsigned p01 = prmt_(yv[0], yv[1], 0x5410u), p23 = prmt_(yv[2], yv[3], 0x5410u);
				out[j]             = prmt_(p01, p23, 0x7531u);                  // rint(sqrt(S) / 4) where that is safe
				const unsigned lo  = prmt_(p01, p23, 0x6420u);
				tie[j]             = (lo - 0x01010101u) & ~lo & 0x80808080u;     // 0x80 in the bytes of lo that are zero
			}
			const uint4 o = make_uint4(out[0], out[1], out[2], out[3]);
			if (active) {
				*reinterpret_cast<uint4 *>(G + (size_t) (z * H + y) * W + cx * 16) = o;
				if (SURF) surf3Dwrite(o, surf, (int) (cx * 16), (int) y, (int) z);        // x in bytes
			}
			const unsigned t = (tie[0] >> 7) | (tie[1] >> 6) | (tie[2] >> 5) | (tie[3] >> 4);
			grad_push(q, qn, warp, lane, t != 0u && active, t, y | (z << 16), cx);
		}
		cx = ncx, y = ny, z = nz, cur = nxt;
		// (the __syncwarp at the top of grad_drain orders this iteration's row stores before the byte patches)
		while (qn >= 32u) grad_drain<SURF>(q, qn, s_lut, warp, lane, 32u, V, G, surf, W, H, D);
	}
	while (qn) grad_drain<SURF>(q, qn, s_lut, warp, lane, qn < 32u ? qn : 32u, V, G, surf, W, H, D);
}

// ---- column walk ------------------------------------------------------------------------------------------------
// The tap lattice (x +- 1, y +- 1, z +- 1) splits the rows of the volume by the parity of y: the voxels of row y read rows
// y - 1 and y + 1 only.  A thread therefore owns one column — a 16-voxel chunk of x at one z — and walks every second row
// of a y segment: the two rows (y + 1, z - 1), (y + 1, z + 1) it loads for row y serve row y + 2 again (as its y - 1 taps),
// so a step is TWO 16-byte loads instead of four, the addresses are one multiply-add per plane (no chunk -> (x, y, z)
// decode, no carries), and nothing but the row registers rotates (the loop is unrolled by the three row pairs in flight).
// The x -+ 1 byte of a row comes from the neighbouring lane as one indexed shuffle of an "edge word" built per row:
// byte 0 = the row's first voxel, byte 3 = its last, bytes 1/2 = the voxel before / after the warp's run (loaded by the outer
// lanes of a run only).  Which lane a thread asks and which byte it takes — neighbour, own word at a clamped volume edge, or the
// memory byte — never changes during the walk, so the clamps cost nothing in the loop.  The shift by the tap's x offset is
// folded into the first stage of the 4x4 byte transpose (9 + 9 permutes per chunk instead of 16 shifts + 16 permutes).
// Arithmetic and tie handling as in the flat kernel: same bytes, about 40 % fewer instructions per voxel.
struct GradPair {        // rows (y', z - 1) and (y', z + 1) of one column
	uint4    m, p;
	unsigned xm, xp;        // the word beside the warp's run (lanes 0 / 31 only; undefined elsewhere)
	unsigned em, ep;        // edge words
};

__device__ __forceinline__ void grad_pair_load(GradPair &r, const uint8_t *__restrict__ baseM, const uint8_t *__restrict__ baseP, uint32_t row, uint32_t W,
                                               bool need_x, int xoff)
{
	// (volatile: the compiler otherwise sinks these loads to the end of the step to shorten their live ranges, which leaves
	// them ~50 instructions instead of a whole step ahead of their first use)
	const size_t o = (size_t) row * W;
	asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.m.x), "=r"(r.m.y), "=r"(r.m.z), "=r"(r.m.w) : "l"(baseM + o));
	asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.p.x), "=r"(r.p.y), "=r"(r.p.z), "=r"(r.p.w) : "l"(baseP + o));
	r.xm = r.xp = 0;
	if (need_x) {
		asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r.xm) : "l"(baseM + o + xoff));
		asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r.xp) : "l"(baseP + o + xoff));
	}
}

// Ties of the column walk.  Every row step stashes the packed taps of its 16 voxels in shared memory (four 16-byte stores per lane
// into a ring of kWalkRing steps), and a lane with ties appends ONE entry (its tie mask, its lane, the step) to the warp's FIFO.  A
// pass takes the 32 OLDEST entries, a lane per entry: lowest tie of the mask, taps from the stash (one shared-memory load — no tap
// addresses, no global re-reads, nothing to wait for), the shader's fp32 chain, and the byte goes into the step's RESULT ROW, which
// is staged in shared memory as well and leaves for the linear map and the texture array a step after it was computed — whole
// 16-byte stores only, no byte patches in memory.  Entries with ties left go back to the HEAD of the queue, so the queue stays
// sorted by age and at the benchmark volume's tie rate (a pass every 1.6 steps) an entry is served while its step is still staged;
// an entry that has outlived the ring (sparse ties: the queue reaches 32 entries only every so many steps) re-reads its four taps
// from global memory and patches the stored byte in both maps instead.
// Steps staged in shared memory.  Shared memory is taken from the L1 that serves the walk's row loads (every row is read by the
// columns at z - 1 and z + 1): measured on B200 with 3 CTAs per SM, ring 3 (70 KB per CTA) 0.347 / 1.07 / 4.26 ms on the
// 832x832x494 / 1024^3 / 2048x2048x1024 volumes, ring 2 (45 KB) 0.328 / 1.10 / 4.20 ms; 2 CTAs per SM lose 5-10 % either way.
#ifndef VKV_WALK_RING
#define VKV_WALK_RING 2
#endif
#ifndef VKV_WALK_CTAS
#define VKV_WALK_CTAS 3
#endif
#ifndef VKV_WALK_Q
#define VKV_WALK_Q 64
#endif
constexpr int kWalkQ = VKV_WALK_Q, kWalkRing = VKV_WALK_RING;        // queue capacity per warp (<= 31 carried + 32 pushed); steps staged in shared memory
struct WalkShared {
	uint4    stash[kWalkRing][4][8][32];        // [step % ring][word of the chunk][warp][lane]: (A, B, E, C) bytes of voxels 4j .. 4j+3
	uint4    rows[kWalkRing][8][32];            // [step % ring][warp][lane]: the step's 16 result bytes, stored to memory ring - 1 steps later
	unsigned q[8][kWalkQ];                     // tie mask (bit 4 i + j <=> voxel 4 j + i) | lane << 16 | (step mod 2048) << 21
	float    lut[256];
};

template <bool SURF>
__device__ __forceinline__ void walk_tie_pass(WalkShared &S, unsigned &qhead, unsigned qtail, int warp, int lane, unsigned step_now, unsigned first_staged, uint32_t y_first,
                                              uint32_t cx, uint32_t z, const uint8_t *__restrict__ V, uint8_t *__restrict__ G, cudaSurfaceObject_t surf,
                                              uint32_t W, uint32_t H, uint32_t D, uint32_t dbg)
{
	__syncwarp();        // this step's row stores, stash and queue entries are visible to the whole warp
	const unsigned n_take = min(32u, qtail - qhead);
	const bool     mine   = (unsigned) lane < n_take;
	unsigned       ent = 0;
	if (mine) ent = S.q[warp][(qhead + lane) & (kWalkQ - 1)];
	__syncwarp();
	// entries with ties left return to the head, in order: the queue stays sorted by age
	const unsigned t = ent & 0xffffu, rest = t & (t - 1u);
	const bool     surv = mine && rest != 0u;
	const unsigned ms   = __ballot_sync(0xffffffffu, surv), nsurv = __popc(ms);
	if (surv) S.q[warp][(qhead + n_take - nsurv + __popc(ms & ((1u << lane) - 1u))) & (kWalkQ - 1)] = (ent & 0xffff0000u) | rest;
	qhead += n_take - nsurv;
	// the column of the lane that owns the entry (constant over its walk)
	const int      src  = (int) ((ent >> 16) & 31u);
	const uint32_t cx_s = __shfl_sync(0xffffffffu, cx, src), z_s = __shfl_sync(0xffffffffu, z, src);
	if (mine) {
		const unsigned es  = step_now - ((step_now - (ent >> 21)) & 0x7ffu);        // the entry's step (entries are younger than 2048 steps)
		const unsigned pos = __ffs(t) - 1, j = pos & 3u, i = pos >> 2;
		const uint32_t x = cx_s * 16 + 4 * j + i, y = y_first + 2u * es + ((unsigned) src >> 4);
		unsigned       wv;        // (A, B, E, C)
		const bool     staged = es >= first_staged;        // the step's taps and result row are still in shared memory
		if (staged) {
			wv = reinterpret_cast<const unsigned *>(&S.stash[es % kWalkRing][j][warp][src])[i];
		} else {
			const uint32_t xm = x > 0 ? x - 1 : 0, xp = x + 1 < W ? x + 1 : W - 1;
			const uint32_t ym = y > 0 ? y - 1 : 0, yp = y + 1 < H ? y + 1 : H - 1;
			const uint32_t zmH = (z_s > 0 ? z_s - 1 : 0) * H, zpH = (z_s + 1 < D ? z_s + 1 : D - 1) * H;
			const uint8_t *Vp = V + xp, *Vm = V + xm;
			wv = (unsigned) __ldg(Vp + (size_t) (zmH + ym) * W) | ((unsigned) __ldg(Vm + (size_t) (zpH + ym) * W) << 8) |
			     ((unsigned) __ldg(Vp + (size_t) (zpH + yp) * W) << 16) | ((unsigned) __ldg(Vm + (size_t) (zmH + yp) * W) << 24);
		}
		const unsigned char g = gradient_byte(S.lut[wv & 0xffu], S.lut[(wv >> 8) & 0xffu], S.lut[wv >> 24], S.lut[(wv >> 16) & 0xffu], 1.0f);
		if (staged) {
			reinterpret_cast<unsigned char *>(&S.rows[es % kWalkRing][warp][src])[4 * j + i] = g;
		} else {
			if (!(dbg & 16) || g == 77) G[(size_t) (z_s * H + y) * W + x] = g;
			if (SURF && (!(dbg & 32) || g == 77)) surf3Dwrite(g, surf, (int) x, (int) y, (int) z_s);
		}
	}
}

__device__ __forceinline__ unsigned grad_edge_word(const uint4 &q, unsigned x)
{
	return prmt_(prmt_(q.x, q.w, 0x7000u), x, 0x3470u);        // (q[0], x[3], x[0], q[15])
}

// ABL: the ablation instantiation (VKV_GRAD_DBG=<bits>, timing only — the map is wrong): 1 no tie queue, 2 no row surface stores,
// 4 no row stores to the linear map, 8 no arithmetic (rows copied through), 16 / 32 no tie patches to the linear map / the array.
template <bool SURF, bool ABL>
__global__ void __launch_bounds__(256, VKV_WALK_CTAS) gradient_walk_kernel(const uint8_t *__restrict__ V, uint8_t *__restrict__ G, cudaSurfaceObject_t surf,
                                                              uint32_t W, uint32_t H, uint32_t D, uint32_t ncols, uint32_t ncg, uint32_t nseg, uint32_t steps, uint32_t dbg_bits)
{
	const uint32_t dbg = ABL ? dbg_bits : 0u;
	extern __shared__ __align__(16) unsigned char walk_smem[];
	WalkShared &S = *reinterpret_cast<WalkShared *>(walk_smem);
	S.lut[threadIdx.x] = (float) threadIdx.x / 255.0f;        // UNORM decode, exactly as imageLoad
	__syncthreads();
	const uint32_t nchunks = W / 16;
	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned       qhead = 0, qtail = 0;        // this warp's FIFO (warp-uniform counters; entry = counter mod kWalkQ)
	// A warp is 16 columns x the two row parities (lanes 16..31 walk the odd rows): each step it completes two adjacent rows of
	// 256 B, which is what the block-linear texture array wants — surface stores of 512 B x 1 row per warp run at less than half
	// the rate of 256 B x 2 (scripts/ubench/sust_patterns.cu: 0.168 vs 0.094 ms for this volume, copy + surface 0.35 vs 0.18 ms).
	// column = (z, chunk) with the chunk fastest; warp = (segment, group of 16 columns); lanes past the last column mirror it.
	const uint32_t gw  = blockIdx.x * 8u + (uint32_t) warp, l16 = (uint32_t) lane & 15u;
	const uint32_t seg = gw / ncg, cg = gw - seg * ncg;
	uint32_t       col = cg * 16u + l16;
	const bool     valid = seg < nseg && col < ncols;
	col                  = col < ncols ? col : ncols - 1;
	const uint32_t z = col / nchunks, cx = col - z * nchunks;
	uint32_t       y = seg * (2u * steps) + ((uint32_t) lane >> 4);        // first row; then every second one
	const bool     first = cx == 0, last = cx == nchunks - 1;
	const uint32_t zm = z > 0 ? z - 1 : 0, zp = z + 1 < D ? z + 1 : D - 1;
	const uint8_t *baseM = V + (size_t) zm * H * W + cx * 16, *baseP = V + (size_t) zp * H * W + cx * 16;
	uint8_t       *baseG = G + (size_t) z * H * W + cx * 16;
	// where the x - 1 / x + 1 byte of a row comes from: the lane asked and the byte of its edge word (4 + k: operand b of prmt)
	const bool     mem_prev = l16 == 0 && !first, mem_next = l16 == 15 && !last;
	const int      src_prev = (first || l16 == 0) ? lane : lane - 1, src_next = (last || l16 == 15) ? lane : lane + 1;
	const unsigned kp = first ? 0u : mem_prev ? 1u : 3u, kn = last ? 3u : mem_next ? 2u : 0u;
	const unsigned sel_prev = 0x0100u | ((4u + kp) << 12) | ((4u + kp) << 4);        // (a0, e[kp], a1, e[kp]): voxel 0 in the high half
	const unsigned sel_next = 0x6060u | kn | (kn << 8);                              // (e[kn], b2, e[kn], b2): voxel 15 in the low half
	const bool     need_x = mem_prev || mem_next;
	const int      xoff   = mem_prev ? -4 : 16;
	const uint32_t Hm1    = H - 1;

	const uint32_t y_first = seg * (2u * steps);        // row of step 0, parity 0 (warp-uniform)
	unsigned       sidx    = 0;                         // steps walked (warp-uniform)
	GradPair r0, r1, r2;
	grad_pair_load(r0, baseM, baseP, min(y > 0 ? y - 1 : 0u, Hm1), W, need_x, xoff);
	grad_pair_load(r1, baseM, baseP, min(y + 1, Hm1), W, need_x, xoff);
	r0.em = grad_edge_word(r0.m, r0.xm), r0.ep = grad_edge_word(r0.p, r0.xp);

	// the result row of step r leaves shared memory: one 16-byte store to the linear map, one into the texture array
	auto store_row = [&](unsigned r) {
		const uint32_t yr = y_first + ((uint32_t) lane >> 4) + 2u * r;
		if (valid && yr < H) {
			const uint4 o = S.rows[r % kWalkRing][warp][lane];
			if (!(dbg & 4)) *reinterpret_cast<uint4 *>(baseG + (size_t) yr * W) = o;
			if (SURF && !(dbg & 2)) surf3Dwrite(o, surf, (int) (cx * 16), (int) yr, (int) z);        // x in bytes
		}
	};
	// one row of the walk: taps A, B on the pair `lo` (row y - 1), C, E on `hi` (row y + 1); `nx` receives row y + 3
	auto step = [&](GradPair &lo, GradPair &hi, GradPair &nx, const int slot) {
		grad_pair_load(nx, baseM, baseP, min(y + 3, Hm1), W, need_x, xoff);
		hi.em = grad_edge_word(hi.m, hi.xm), hi.ep = grad_edge_word(hi.p, hi.xp);
		const unsigned eA = __shfl_sync(0xffffffffu, lo.em, src_next), eB = __shfl_sync(0xffffffffu, lo.ep, src_prev);
		const unsigned eC = __shfl_sync(0xffffffffu, hi.em, src_prev), eE = __shfl_sync(0xffffffffu, hi.ep, src_next);
		const unsigned a[4] = {lo.m.x, lo.m.y, lo.m.z, lo.m.w}, b[4] = {lo.p.x, lo.p.y, lo.p.z, lo.p.w};        // A at x + 1, B at x - 1
		const unsigned c[4] = {hi.m.x, hi.m.y, hi.m.z, hi.m.w}, e[4] = {hi.p.x, hi.p.y, hi.p.z, hi.p.w};        // C at x - 1, E at x + 1
		// first transpose stage with the x shift folded in.  P[j] = voxels (4j - 1, 4j), Q[j] = voxels (4j + 1, 4j + 2);
		// a half is (tap at x + 1, tap at x - 1) of one voxel
		unsigned pab[5], pce[5], qab[4], qce[4];
		pab[0] = prmt_(a[0], eB, sel_prev), pce[0] = prmt_(e[0], eC, sel_prev);
		pab[4] = prmt_(eA, b[3], sel_next), pce[4] = prmt_(eE, c[3], sel_next);
#pragma unroll
		for (int j = 1; j < 4; ++j) pab[j] = prmt_(a[j], b[j - 1], 0x7160u), pce[j] = prmt_(e[j], c[j - 1], 0x7160u);
#pragma unroll
		for (int j = 0; j < 4; ++j) qab[j] = prmt_(a[j], b[j], 0x5342u), qce[j] = prmt_(e[j], c[j], 0x5342u);
		unsigned out[4] = {0, 0, 0, 0}, tie[4] = {0, 0, 0, 0};
		if (!(dbg & 8))
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			// second stage: the four taps of one voxel in one word (their order does not matter to the two dot products)
			const unsigned wv[4] = {prmt_(pab[j], pce[j], 0x7632u), prmt_(qab[j], qce[j], 0x5410u), prmt_(qab[j], qce[j], 0x7632u),
			                        prmt_(pab[j + 1], pce[j + 1], 0x5410u)};
			S.stash[slot][j][warp][lane] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
			unsigned yv[4];
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const unsigned qq = __dp4a(wv[i], wv[i], 0x12c00000u);       // A^2 + B^2 + C^2 + D^2 + 0x4b000000 / 4
				const unsigned sm = __dp4a(wv[i], 0x01010101u, 0u);          // A + B + C + D
				const unsigned Sb = (qq << 2) - sm * sm;                    // bits of the float 2^23 + S   (S = 4 qq - sm^2 < 2^18)
				const float    f  = __uint_as_float(Sb) - 8388608.0f;       // S, exactly
				float          rt;
				asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(f));
				yv[i] = __float_as_uint(__fmaf_rn(rt, 64.0f, 8388736.0f));   // low 16 mantissa bits = rint(64 sqrt(S)) + 128
			}
			const unsigned p01 = prmt_(yv[0], yv[1], 0x5410u), p23 = prmt_(yv[2], yv[3], 0x5410u);
			out[j]             = prmt_(p01, p23, 0x7531u);                  // rint(sqrt(S) / 4) where that is safe
			const unsigned lo8 = prmt_(p01, p23, 0x6420u);
			tie[j]             = (lo8 - 0x01010101u) & ~lo8 & 0x80808080u;   // 0x80 in the bytes of lo8 that are zero
		}
		const bool  active = valid && y < H;
		uint4 o      = make_uint4(out[0], out[1], out[2], out[3]);
		if (dbg & 8) o = make_uint4(lo.m.x ^ eA, lo.p.y ^ eB, hi.m.z ^ eC, hi.p.w ^ eE);
		S.rows[slot][warp][lane] = o;
		const unsigned t8 = (tie[0] >> 7) | (tie[1] >> 6) | (tie[2] >> 5) | (tie[3] >> 4);        // bit 8 i + j <=> voxel 4 j + i
		{
			const bool     push = t8 != 0u && active && !(dbg & 1);
			const unsigned m    = __ballot_sync(0xffffffffu, push);
			if (push) {
				const unsigned t16 = (t8 & 0xfu) | ((t8 >> 4) & 0xf0u) | ((t8 >> 8) & 0xf00u) | ((t8 >> 12) & 0xf000u);
				S.q[warp][(qtail + __popc(m & ((1u << lane) - 1u))) & (kWalkQ - 1)] = t16 | ((unsigned) lane << 16) | (sidx << 21);
			}
			qtail += __popc(m);
		}
		y += 2;
		const unsigned first_staged = sidx >= (unsigned) (kWalkRing - 1) ? sidx - (unsigned) (kWalkRing - 1) : 0u;
		while (qtail - qhead >= 32u) walk_tie_pass<SURF>(S, qhead, qtail, warp, lane, sidx, first_staged, y_first, cx, z, V, G, surf, W, H, D, dbg);
		__syncwarp();        // the passes' byte patches are in the staged rows; their reads of the oldest stash slot are done
		if (sidx >= (unsigned) (kWalkRing - 1)) store_row(sidx - (unsigned) (kWalkRing - 1));        // (its slot is the one the next step overwrites)
		++sidx;
	};
	for (uint32_t k = 0; k < steps; k += 3) {
		step(r0, r1, r2, kWalkRing == 3 ? 0 : (int) (sidx % kWalkRing));
		step(r1, r2, r0, kWalkRing == 3 ? 1 : (int) (sidx % kWalkRing));
		step(r2, r0, r1, kWalkRing == 3 ? 2 : (int) (sidx % kWalkRing));
	}
	// the last kWalkRing - 1 rows are still staged: the remaining ties first, then the rows
	while (qtail != qhead) walk_tie_pass<SURF>(S, qhead, qtail, warp, lane, sidx - 1u, sidx - (unsigned) (kWalkRing - 1), y_first, cx, z, V, G, surf, W, H, D, dbg);
	__syncwarp();
#pragma unroll
	for (int r = kWalkRing - 1; r >= 1; --r) store_row(sidx - (unsigned) r);
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
	vol->G_array_synced = false;
	if (!use_gradient) {
		// get_gradient returns 1.0 for every voxel (get_gradient_compute.glsl:6-7) -> byte 255
		VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_G, 255, vol->N, s));
	} else if (vol->dim[0] % 16 == 0 && reinterpret_cast<uintptr_t>(vol->d_V) % 16 == 0 && modifier == 1.0f && !getenv("VKV_GRAD_FP32") &&
	           (uint64_t) ((vol->dim[0] / 16 + 31) / 32) * vol->dim[1] * vol->dim[2] < (1ull << 31)) {
		const unsigned ntasks = ((vol->dim[0] / 16 + 31) / 32) * vol->dim[1] * vol->dim[2];
		const unsigned grid   = (ntasks + 7) / 8;
		const uint64_t chunks = (uint64_t) (vol->dim[0] / 16) * vol->dim[1] * vol->dim[2];
		// The flat persistent walk wins while the volume's working set (the z +- 1 tap slices, two slices apart) stays in L2; on
		// multi-gigabyte volumes the row-task kernel, whose warps sweep one row each and keep neighbouring rows together, is up
		// to 2x faster (measured: 2048x2048x1024 5.4 vs 6.0 ms, 4096x4096x2048 44 vs 86 ms; 832x832x494 0.50 vs 0.37 ms).
		const bool flat_ok = vol->N <= (2ull << 30) || getenv("VKV_GRAD_FLAT");
		// The column walk (round 2) replaces both wherever its packed indices fit: half the loads, no index arithmetic in the loop.
		const uint32_t nch = vol->dim[0] / 16;
		const uint64_t ncols = (uint64_t) nch * vol->dim[2];
		const uint64_t ncg   = (ncols + 15) / 16;        // warps per segment: 16 columns x 2 row parities
		auto           n_seg = [&](uint32_t st) { return (uint64_t) ((vol->dim[1] + 2 * st - 1) / (2 * st)); };
		const uint64_t resident = (uint64_t) vol->ctx->sm_count * VKV_WALK_CTAS * 8;        // warps
		// Rows per thread (a multiple of 3; a segment is 2 * steps rows, two parities): the count that walks the fewest rows — the
		// last segment of a column is padded, and every segment pays about a step and a half of prologue — among those that leave
		// at least six waves of warps (measured on B200: 832 rows 30 steps 0.316 ms against 0.327 at 24; 1024 rows 27 steps
		// 1.054 ms against 1.076); small volumes halve the count until the machine is filled.
		uint32_t steps = 0;
		double   best  = 0.0;
		for (uint32_t st = 12; st <= 48; st += 3) {
			if (ncg * n_seg(st) < 6 * resident) continue;
			const double cost = (double) n_seg(st) * ((double) st + 1.5);
			if (steps == 0 || cost <= best) steps = st, best = cost;
		}
		if (steps == 0) {
			steps = 24;
			while (steps > 3 && ncg * n_seg(steps) < 8 * resident) steps /= 2;
		}
		if (const char *e = getenv("VKV_GRAD_STEPS")) steps = (uint32_t) std::min(680, std::max(1, atoi(e))) * 3;        // (queue entries carry the step mod 2048)
		const bool walk_ok = !getenv("VKV_GRAD_V1") && !getenv("VKV_GRAD_FLAT") && nch <= 65535 && vol->dim[1] <= 65535 && vol->dim[2] <= 65535 &&
		                     ncg * n_seg(steps) < (1ull << 31) && ncols < (1ull << 31);
		if (walk_ok) {
			const uint64_t warps = ncg * n_seg(steps);
			const unsigned g     = (unsigned) ((warps + 7) / 8);
			const bool     surf  = vol->s_G && !getenv("VKV_GRAD_NOSURF");
			const uint32_t dbg   = (surf && getenv("VKV_GRAD_DBG")) ? (uint32_t) atoi(getenv("VKV_GRAD_DBG")) : 0u;
			static PerDeviceOnce walk_configured;
			if (walk_configured.first(vol->ctx->device)) {
				VKV_CUDA_CHECK(cudaFuncSetAttribute(gradient_walk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(WalkShared)));
				VKV_CUDA_CHECK(cudaFuncSetAttribute(gradient_walk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(WalkShared)));
				VKV_CUDA_CHECK(cudaFuncSetAttribute(gradient_walk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(WalkShared)));
			}
#define VKV_GRAD_WALK(SURF_, ABL_)                                                                                                                 \
	gradient_walk_kernel<SURF_, ABL_><<<g, 256, sizeof(WalkShared), s>>>(vol->d_V, vol->d_G, SURF_ ? vol->s_G : 0, vol->dim[0], vol->dim[1], vol->dim[2], (uint32_t) ncols, \
	                                                    (uint32_t) ncg, (uint32_t) n_seg(steps), steps, dbg)
			if (dbg)
				VKV_GRAD_WALK(true, true);
			else if (surf)
				VKV_GRAD_WALK(true, false);
			else
				VKV_GRAD_WALK(false, false);
#undef VKV_GRAD_WALK
			vol->G_array_synced = surf;        // the kernel wrote the array itself
		} else if (!getenv("VKV_GRAD_V1") && flat_ok && vol->dim[0] <= 65536 && vol->dim[1] <= 65536 && vol->dim[2] < 65536 && chunks < (1ull << 32)) {
			// persistent: every warp resident at once, each thread walks chunks gid, gid + stride, ...
			const unsigned nchunks = vol->dim[0] / 16;
			const char    *e_ctas      = getenv("VKV_GRAD_CTAS");
			const int      ctas_per_sm = e_ctas ? atoi(e_ctas) : 2;        // measured best on B200 (scripts/grad_probe.py)
			uint64_t g = (uint64_t) vol->ctx->sm_count * ctas_per_sm;
			if (g * 256 > chunks) g = (chunks + 255) / 256;
			const uint64_t stride = g * 256;
			const uint32_t niter  = (uint32_t) ((chunks + stride - 1) / stride);
			const uint32_t dcx = (uint32_t) (stride % nchunks), rr = (uint32_t) (stride / nchunks), dy = rr % vol->dim[1], dz = rr / vol->dim[1];
			const bool     surf = vol->s_G && !getenv("VKV_GRAD_NOSURF");
			if (surf)
				gradient_flat_kernel<true><<<(unsigned) g, 256, 0, s>>>(vol->d_V, vol->d_G, vol->s_G, vol->dim[0], vol->dim[1], vol->dim[2], niter, dcx, dy, dz);
			else
				gradient_flat_kernel<false><<<(unsigned) g, 256, 0, s>>>(vol->d_V, vol->d_G, 0, vol->dim[0], vol->dim[1], vol->dim[2], niter, dcx, dy, dz);
			vol->G_array_synced = surf;        // the kernel wrote the array itself
		} else if (vol->s_G && !getenv("VKV_GRAD_NOSURF")) {
			gradient_int_kernel<true><<<grid, 256, 0, s>>>(vol->d_V, vol->d_G, vol->s_G, vol->dim[0], vol->dim[1], vol->dim[2]);
			vol->G_array_synced = true;        // the kernel wrote the array itself
		} else {
			gradient_int_kernel<false><<<grid, 256, 0, s>>>(vol->d_V, vol->d_G, 0, vol->dim[0], vol->dim[1], vol->dim[2]);
		}
		VKV_LAUNCHED();
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
