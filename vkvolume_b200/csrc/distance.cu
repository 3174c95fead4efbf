// distance.cu — K3a isotropic and K3b anisotropic Chebyshev distance maps.
//
// Replaces shaders/distance_map.comp (3 dispatches) and shaders/distance_map_anisotropic.comp
// (14 dispatches) driven by src/compute_distance_map.cpp:142-290.  In the reference one
// invocation walks a whole row/column serially (only Wb*Hb threads in flight) and searches up
// to the current distance at every cell.
//
// The result has a closed form, D(p) = min(255, min over occupied q of the
// Chebyshev distance), optionally restricted to an octant, and it is separable:
//   x pass : g(x)   = distance to the nearest occupied cell in the row (both sides, or one side)
//   y pass : s(y)   = min_n max(|n|, g(y+n))
//   z pass : D(z)   = min_n max(|n|, s(z+n))
// so any exact evaluation is bit-identical to the reference's passes.
//
// B200 design: every output cell gets its own thread.
//   * x pass: the row is turned into a bit mask with warp ballots and each lane finds the
//     nearest set bit with clz/ffs over at most 8 words per side (255-cell cap) — O(1), coalesced.
//   * y / z passes: s(y) <= r iff the minimum of g over the window [y-r, y+r] is <= r — monotone in r — so
//     each cell does an 8-step binary search on r against a sparse range-minimum table built in shared
//     memory from the staged tile (32 adjacent columns x the whole line, coalesced 32-byte row segments).
//     One thread per cell, no serial scan, ~25 conflict-free shared-memory operations per cell whatever
//     the distances are.  Lines too long for shared memory use narrower tiles, then a global-memory search.
//   * the anisotropic build shares the 2 x passes and 4 y passes between the 8 octant maps
//     (same sharing as the reference's 14-dispatch schedule).
// The maps are small (M bytes, L2-resident below ~100 MB); algorithmic bytes 6 B/block
// (isotropic) and 28 B/block (anisotropic) as in SURVEY §8(d).
#include <cstdlib>

#include "common.cuh"

namespace vkv {

constexpr int kRowWordsMax = 64;          // x pass: rows up to 2048 cells use the ballot path
constexpr int kLineMax     = 1024;        // y/z pass: lines up to 1024 cells are staged in shared memory

// ---- x pass ---------------------------------------------------------------------------------
// DIR = 0 both sides, +1 towards +x only, -1 towards -x only (distance_map_anisotropic.comp:44-53).
template <int DIR>
__global__ void __launch_bounds__(256) xpass_ballot_kernel(const uint8_t *__restrict__ O, uint8_t *__restrict__ out, uint32_t Wb,
                                                          uint64_t nrows)
{
	__shared__ unsigned s_words[8][kRowWordsMax + 16];        // 8 guard words each side, kept zero
	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned      *words  = &s_words[warp][8];
	const uint32_t nwords = (Wb + 31) / 32;
	for (int i = lane; i < kRowWordsMax + 16; i += 32) s_words[warp][i] = 0;
	__syncwarp();
	for (uint64_t row = (uint64_t) blockIdx.x * 8 + warp; row < nrows; row += (uint64_t) gridDim.x * 8) {
		const uint8_t *src = O + row * Wb;
		for (uint32_t w = 0; w < nwords; ++w) {
			const uint32_t x   = w * 32 + lane;
			const bool     occ = x < Wb && src[x] == 0;
			const unsigned m   = __ballot_sync(0xffffffffu, occ);
			if (lane == 0) words[w] = m;
		}
		__syncwarp();
		for (uint32_t w = 0; w < nwords; ++w) {
			const uint32_t x = w * 32 + lane;
			unsigned       d = 255;
			if (DIR >= 0) {        // nearest occupied cell at x' >= x
				unsigned m = words[w] >> lane;
				if (m) d = min(d, (unsigned) (__ffs(m) - 1));
				else {
#pragma unroll
					for (int k = 1; k <= 8; ++k) {
						const unsigned mk = words[w + k];
						if (mk) { d = min(d, (unsigned) (k * 32 - lane + __ffs(mk) - 1)); break; }
					}
				}
			}
			if (DIR <= 0) {        // nearest occupied cell at x' <= x
				unsigned m = words[w] << (31 - lane);
				if (m) d = min(d, (unsigned) __clz(m));
				else {
#pragma unroll
					for (int k = 1; k <= 8; ++k) {
						const unsigned mk = words[(int) w - k];
						if (mk) { d = min(d, (unsigned) (lane + 1 + (k - 1) * 32 + __clz(mk))); break; }
					}
				}
			}
			if (x < Wb) out[row * Wb + x] = (uint8_t) d;
		}
		__syncwarp();
	}
}


// Same, four cells per lane: a warp handles 128 cells per step with one 32-bit load and one 32-bit
// store per lane; the 4-bit occupancy nibbles are merged into mask words with three xor-shuffles.
// The search for the nearest set bit is O(1) whatever the distance: a 64-bit mask of the row's NON-ZERO words
// (two ballots) names the word that holds it, so an empty stretch costs two bit scans instead of a loop over
// up to eight words per side; only the outer cells of a lane's four are searched, the inner ones follow from
// d(x) = occupied ? 0 : d(x +- 1) + 1.   Needs Wb % 4 == 0 (rows then start 4-byte aligned).
struct RowBits {
	const unsigned    *words;
	unsigned long long nz;        // bit w: words[w] != 0
	// distance from bit `bit` of word w to the nearest set bit at or after it (255: none within 254)
	__device__ __forceinline__ unsigned right(int w, int bit) const
	{
		const unsigned m = words[w] >> bit;
		if (m) return (unsigned) (__ffs(m) - 1);
		const unsigned long long rest = w < 63 ? nz >> (w + 1) : 0ull;
		if (!rest) return 255u;
		const int k = __ffsll((long long) rest);        // the set bit is in word w + k
		return min((unsigned) (k * 32 - bit + __ffs(words[w + k]) - 1), 255u);
	}
	// ... at or before it
	__device__ __forceinline__ unsigned left(int w, int bit) const
	{
		const unsigned m = words[w] << (31 - bit);
		if (m) return (unsigned) __clz(m);
		const unsigned long long rest = nz & ((1ull << w) - 1ull);
		if (!rest) return 255u;
		const int w2 = 63 - __clzll((long long) rest);
		return min((unsigned) (bit + 1 + (w - 1 - w2) * 32 + __clz(words[w2])), 255u);
	}
};

template <int DIR, int SEGS>        // SEGS: compile-time bound on the row's 128-cell segments (2, 4, 8 or 16)
__global__ void __launch_bounds__(256) xpass_vec4_kernel(const uint8_t *__restrict__ O, uint8_t *__restrict__ out, uint32_t Wb, uint64_t nrows)
{
	__shared__ unsigned s_words[8][kRowWordsMax];
	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned      *words = s_words[warp];
	const uint32_t nseg  = (Wb + 127) / 128;
	for (int i = lane; i < kRowWordsMax; i += 32) words[i] = 0;
	__syncwarp();
	for (uint64_t row = (uint64_t) blockIdx.x * 8 + warp; row < nrows; row += (uint64_t) gridDim.x * 8) {
		const uint8_t *src = O + row * Wb;
		// the whole row is requested before any of it is looked at: one DRAM round trip per row, not one per 128 cells
		unsigned cw[SEGS];
#pragma unroll
		for (uint32_t sgm = 0; sgm < (uint32_t) SEGS; ++sgm) {
			const uint32_t x = sgm * 128 + lane * 4;
			cw[sgm]          = (sgm < nseg && x < Wb) ? __ldg(reinterpret_cast<const unsigned *>(src + x)) : 0xffffffffu;
		}
#pragma unroll
		for (uint32_t sgm = 0; sgm < (uint32_t) SEGS; ++sgm) {
			if (sgm >= nseg) break;
			const unsigned c = cw[sgm];
			unsigned       t = (c & 0x7f7f7f7fu) + 0x7f7f7f7fu;        // exact zero-byte detector -> 0x80 per zero byte
			t                = ~(t | c | 0x7f7f7f7fu);
			const unsigned y = t >> 7;
			unsigned       v = ((y | (y >> 7) | (y >> 14) | (y >> 21)) & 0xfu) << (4 * (lane & 7));
			v |= __shfl_xor_sync(0xffffffffu, v, 1);
			v |= __shfl_xor_sync(0xffffffffu, v, 2);
			v |= __shfl_xor_sync(0xffffffffu, v, 4);
			if ((lane & 7) == 0) words[sgm * 4 + (lane >> 3)] = v;
		}
		__syncwarp();
		RowBits rb;
		rb.words = words;
		rb.nz    = (unsigned long long) __ballot_sync(0xffffffffu, words[lane] != 0u) |
		        ((unsigned long long) __ballot_sync(0xffffffffu, words[32 + lane] != 0u) << 32);
		for (uint32_t sgm = 0; sgm < nseg; ++sgm) {
			const uint32_t x = sgm * 128 + lane * 4;
			if (x < Wb) {
				const int      w = (int) (x >> 5), b = (int) (x & 31);
				const unsigned occ = (words[w] >> b) & 0xfu;
				unsigned       d[4] = {255u, 255u, 255u, 255u};
				if (DIR >= 0) {
					unsigned r = rb.right(w, b + 3);
					d[3]       = r;
#pragma unroll
					for (int k = 2; k >= 0; --k) {
						r    = ((occ >> k) & 1u) ? 0u : min(r + 1u, 255u);
						d[k] = r;
					}
				}
				if (DIR <= 0) {
					unsigned l = rb.left(w, b);
					d[0]       = min(d[0], l);
#pragma unroll
					for (int k = 1; k < 4; ++k) {
						l    = ((occ >> k) & 1u) ? 0u : min(l + 1u, 255u);
						d[k] = min(d[k], l);
					}
				}
				*reinterpret_cast<unsigned *>(out + row * Wb + x) = d[0] | (d[1] << 8) | (d[2] << 16) | (d[3] << 24);
			}
		}
		__syncwarp();
	}
}

// Rows longer than the ballot path allows: one thread per row, literal sweeps.
template <int DIR>
__global__ void __launch_bounds__(128) xpass_serial_kernel(const uint8_t *__restrict__ O, uint8_t *__restrict__ out, uint32_t Wb,
                                                          uint64_t nrows)
{
	const uint64_t row = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= nrows) return;
	const uint8_t *src = O + row * Wb;
	uint8_t       *dst = out + row * Wb;
	if (DIR >= 0) {        // backward sweep: distance to the next occupied cell at x' >= x
		unsigned g = 255;
		for (int64_t x = (int64_t) Wb - 1; x >= 0; --x) {
			g      = src[x] == 0 ? 0u : min(g + 1, 255u);
			dst[x] = (uint8_t) g;
		}
	}
	if (DIR <= 0) {
		unsigned g = 255;
		for (uint32_t x = 0; x < Wb; ++x) {
			g = src[x] == 0 ? 0u : min(g + 1, 255u);
			dst[x] = DIR == 0 ? (uint8_t) min((unsigned) dst[x], g) : (uint8_t) g;
		}
	}
}

// ---- y / z passes ---------------------------------------------------------------------------
// One CTA: 32 adjacent x columns of one line set.  `line_stride` is the element stride between
// consecutive cells of a line (Wb for y lines, Wb*Hb for z lines); `outer_stride` selects the
// slice (y pass: z) or row (z pass: y) handled by blockIdx.y.
// DIR = 0: two-sided search (distance_map.comp:72-108); DIR = +-1: one-sided (distance_map_anisotropic.comp:55-91).
// NOUT = 2 writes both one-sided results of the SAME staged input (dir +1 -> dst0, dir -1 -> dst1).
template <int DIR, int NOUT, bool STAGED>
__global__ void __launch_bounds__(256) minmax_pass_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst0,
                                                         uint8_t *__restrict__ dst1, uint32_t Wb, uint32_t L, size_t line_stride,
                                                         size_t outer_stride)
{
	extern __shared__ uint8_t s_tile[];        // STAGED: L x 32 bytes
	__shared__ unsigned       s_colmin[32];
	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t x    = blockIdx.x * 32 + lane;
	const bool     in_x = x < Wb;
	const size_t   base = (size_t) blockIdx.y * outer_stride + x;
	const uint8_t *col  = src + base;
	if (threadIdx.x < 32) s_colmin[threadIdx.x] = 255u;
	__syncthreads();
	unsigned cmin = 255u;
	for (uint32_t p = warp; p < L; p += 8) {
		const unsigned v = in_x ? (unsigned) col[(size_t) p * line_stride] : 255u;
		if (STAGED) s_tile[p * 32 + lane] = (uint8_t) v;
		cmin = min(cmin, v);
	}
	if (cmin < 255u) atomicMin(&s_colmin[lane], cmin);
	__syncthreads();
	const bool empty_col = s_colmin[lane] >= 255u;

	auto cell = [&](uint32_t p) -> unsigned {
		return STAGED ? (unsigned) s_tile[p * 32 + lane] : (unsigned) __ldg(col + (size_t) p * line_stride);
	};
	for (uint32_t p = warp; p < L; p += 8) {
		if (!in_x) continue;
		if (DIR == 0) {
			unsigned best = empty_col ? 255u : cell(p);
			if (!empty_col) {
				const unsigned reach = max(p, L - 1 - p);        // beyond this no cell exists on either side
				for (unsigned n = 1; n < best && n <= reach; ++n) {
					const unsigned a = p >= n ? cell(p - n) : 255u;
					const unsigned b = p + n < L ? cell(p + n) : 255u;
					best             = min(best, max(n, min(a, b)));
				}
			}
			dst0[base + (size_t) p * line_stride] = (uint8_t) best;
		} else {
#pragma unroll
			for (int o = 0; o < NOUT; ++o) {
				const int dir  = NOUT == 2 ? (o == 0 ? 1 : -1) : DIR;
				unsigned  best = empty_col ? 255u : cell(p);
				if (!empty_col) {
					const unsigned reach = dir > 0 ? L - 1 - p : p;
					for (unsigned n = 1; n < best && n < 255u && n <= reach; ++n) {
						const unsigned g = cell(dir > 0 ? p + n : p - n);
						best             = min(best, max(n, g));
					}
				}
				(o == 0 ? dst0 : dst1)[base + (size_t) p * line_stride] = (uint8_t) best;
			}
		}
	}
}


// ---- y / z passes: range-minimum table + binary search on the radius --------------------------------
// s(y) <= r  <=>  min of g over the window [y-r, y+r] (one-sided: [y, y+r] or [y-r, y]) is <= r, and that
// predicate is monotone in r, so s(y) is found by an 8-step binary search whose window minimum is an O(1)
// lookup in a sparse table M_k(y) = min g[y .. y+2^k) built in shared memory.  Every cell is an independent
// thread (no serial scan along the line), shared-memory accesses are 32 consecutive bytes per warp, and the
// cost per cell is ~25 shared-memory operations whatever the distances are (the reference's search costs
// O(distance) per cell on top of being serial along the line).  Exact, so bit-identical.
// MODE 0: two-sided -> dst0 | 1: towards +axis -> dst0 | 2: towards -axis -> dst0 | 3: both one-sided (+ -> dst0, - -> dst1)
template <int SIDE>        // SIDE 0 two-sided, +1 / -1 one-sided
__device__ __forceinline__ unsigned rmq_search(const uint8_t *__restrict__ T, int L, int TW, size_t level_stride, int y, int col, unsigned g0)
{
	unsigned lo = 0, hi = min(g0, 255u);        // s(y) <= g(y): the n = 0 term
	while (lo < hi) {
		const int r   = (int) ((lo + hi) >> 1);
		const int a   = SIDE > 0 ? y : max(y - r, 0);
		const int b   = SIDE < 0 ? y : min(y + r, L - 1);
		const int k   = 31 - __clz(b - a + 1);
		const uint8_t *Tk = T + (size_t) k * level_stride;
		const unsigned w  = min((unsigned) Tk[a * TW + col], (unsigned) Tk[(b - (1 << k) + 1) * TW + col]);
		if (w <= (unsigned) r) hi = (unsigned) r;
		else lo = (unsigned) r + 1;
	}
	return lo;
}

template <int MODE>
__global__ void __launch_bounds__(256) minmax_rmq_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst0, uint8_t *__restrict__ dst1,
                                                       uint32_t Wb, uint32_t L, size_t line_stride, size_t outer_stride, int TW, int nlev)
{
	extern __shared__ uint8_t T[];        // nlev levels of L x TW bytes
	const int      col  = threadIdx.x % TW, sub = threadIdx.x / TW, nsub = blockDim.x / TW;
	const uint32_t x    = blockIdx.x * TW + col;
	const bool     in_x = x < Wb;
	const size_t   base = (size_t) blockIdx.y * outer_stride + x;
	const size_t   lst  = (size_t) L * TW;
	for (uint32_t p = sub; p < L; p += nsub) T[p * TW + col] = in_x ? src[base + (size_t) p * line_stride] : (uint8_t) 255;
	__syncthreads();
	for (int k = 1; k < nlev; ++k) {
		const uint8_t *prev = T + (size_t) (k - 1) * lst;
		uint8_t       *cur  = T + (size_t) k * lst;
		const uint32_t half = 1u << (k - 1);
		for (uint32_t p = sub; p < L; p += nsub) {
			const unsigned a = prev[p * TW + col];
			const unsigned b = p + half < L ? prev[(p + half) * TW + col] : 255u;
			cur[p * TW + col] = (uint8_t) min(a, b);
		}
		__syncthreads();
	}
	if (!in_x) return;
	for (uint32_t p = sub; p < L; p += nsub) {
		const unsigned g0 = T[p * TW + col];
		const size_t   o  = base + (size_t) p * line_stride;
		if (MODE == 0) dst0[o] = (uint8_t) rmq_search<0>(T, (int) L, TW, lst, (int) p, col, g0);
		if (MODE == 1 || MODE == 3) dst0[o] = (uint8_t) rmq_search<1>(T, (int) L, TW, lst, (int) p, col, g0);
		if (MODE == 2) dst0[o] = (uint8_t) rmq_search<-1>(T, (int) L, TW, lst, (int) p, col, g0);
		if (MODE == 3) dst1[o] = (uint8_t) rmq_search<-1>(T, (int) L, TW, lst, (int) p, col, g0);
	}
}

// ---- y pass as a chamfer sweep ---------------------------------------------------------------------------
// After the x pass g(x, y) is the distance to the nearest occupied cell of the SAME row.  The 2-D Chebyshev
// distance within a slice then obeys the chamfer recurrence
//     F(x, y) = min( g(x, y), 1 + min( F(x-1, y-1), F(x, y-1), F(x+1, y-1) ) )            (rows 0 .. y)
// and the same from the other side, in place on F: one diagonal step towards a source lowers max(|dx|, |dy|) by
// exactly one and no step can lower it by more, and the base term g already covers the sources of the row itself, so
// there is NO dependency inside a row — a row is one parallel step.  Exact (the result is the closed form
// min over sources of max(|dx|, |dy|), saturating at 255 because g <= 255), O(1) per cell, no search at all.
// The octant-restricted maps use the one-sided x pass and only the neighbours {0, sx} (a step may not leave the
// quadrant), and keep the two sweep directions as two outputs (sources at y' >= y -> dst0, y' <= y -> dst1).
// One CTA per z slice; thread t owns cells t, t + T, ...; the previous row lives in shared memory (double-buffered,
// one __syncthreads per row); the rows of the input are prefetched eight at a time.
__device__ __forceinline__ uint32_t d_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void d_mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(d_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void d_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(d_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void d_mbar_wait(uint64_t *bar, unsigned parity)
{
	unsigned           ok, spins = 0;
	unsigned long long t0 = 0ull;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok)
		             : "r"(d_smem_u32(bar)), "r"(parity)
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
__device__ __forceinline__ void d_tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d_smem_u32(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(d_smem_u32(bar))
	             : "memory");
}

// The input rows arrive by 1-D bulk TMA copies, kSweepRows rows (one contiguous block of the slice) per copy, through a
// ring of kSweepSlots buffers with one mbarrier each: the serial row recurrence never waits on a global load.
// A thread owns FOUR adjacent cells.  The previous row lives in shared memory as 16-bit lanes (two u16x2 words per
// thread), so the whole row step is a handful of native packed instructions with a short dependency chain — the step
// latency, not the instruction count, is what bounds a sweep of Hb dependent rows:
//     neighbours x +- 1 : 16-bit funnel shifts against the adjacent words
//     min of the three  : VIMNMX3.U16x2
//     min(g, m + 1)     : VIADDMNMX.U16x2     (m <= 255, so m + 1 needs no saturation; the result is <= g <= 255)
// and a 1024-cell row is 8 warps instead of 32 — the per-row barrier is cheaper and several slices share an SM.
constexpr int kSweepRows  = 8;
constexpr int kSweepSlots = 6;
__device__ __forceinline__ unsigned d_prmt(unsigned a, unsigned b, unsigned sel)
{
	unsigned r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}
// `split` (isotropic, small maps): the two sweeps of a slice are independent when both start from g — the result is then
// min(up, down), which the z pass takes while it stages — so they run as two CTAs (blockIdx.x = 2 slice + sweep) and the serial
// chain of a slice is Hb row steps instead of 2 Hb; costs one more map of traffic, so only where the maps live in the L2.
template <int XDIR>        // XDIR 0: isotropic (two-sided x, 3 neighbours, in-place second sweep); +-1: one-sided
__global__ void __launch_bounds__(256) ysweep_kernel(const uint8_t *__restrict__ g, uint8_t *__restrict__ dst0, uint8_t *__restrict__ dst1,
                                                     uint32_t Wb, uint32_t Hb, int split)
{
	extern __shared__ __align__(128) uint8_t s_dyn[];        // kSweepSlots chunks of kSweepRows x Wb | 2 row buffers of Wb/4 + 2 uint2
	__shared__ __align__(8) uint64_t s_bar[kSweepSlots];
	const uint32_t t      = threadIdx.x;                     // group of 4 cells of the row
	const uint32_t W4     = Wb >> 2;
	const bool     active = t < W4;
	const size_t   slice  = (size_t) (split ? blockIdx.x >> 1 : blockIdx.x) * Wb * Hb;
	const uint32_t chunk_bytes = kSweepRows * Wb;
	uint8_t       *ring = s_dyn;
	uint2         *buf0 = reinterpret_cast<uint2 *>(s_dyn + (size_t) kSweepSlots * chunk_bytes) + 1, *buf1 = buf0 + (W4 + 2);
	const uint32_t nchunks = (Hb + kSweepRows - 1) / kSweepRows;
	constexpr unsigned kFar = 0x00ff00ffu;        // 255 in both lanes
	if (t == 0) {
		for (int i = 0; i < kSweepSlots; ++i) d_mbar_init(&s_bar[i], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	unsigned issued = 0;        // copies issued so far (thread 0): copy n lands in ring slot n % kSweepSlots, phase (n / kSweepSlots) & 1
	for (int sweep = split ? (int) (blockIdx.x & 1u) : 0; sweep < 2; ++sweep) {
		// sweep 0 walks y upwards (sources at y' <= y), sweep 1 downwards (sources at y' >= y)
		const uint8_t  *src  = (XDIR == 0 && sweep == 1 && !split) ? dst0 : g;
		uint8_t        *dst  = (XDIR == 0 && !split) ? dst0 : (sweep == 0 ? dst1 : dst0);
		const ptrdiff_t step = sweep == 0 ? (ptrdiff_t) W4 : -(ptrdiff_t) W4;        // in words
		unsigned       *dp   = reinterpret_cast<unsigned *>(dst + slice + (sweep == 0 ? 0 : (size_t) (Hb - 1) * Wb)) + t;
		// chunk c of this sweep = rows [lo, lo + n) of the slice, consumed upwards (sweep 0) or downwards (sweep 1)
		auto issue = [&](uint32_t c) {
			const uint32_t i0 = c * kSweepRows, n = min((uint32_t) kSweepRows, Hb - i0);
			const uint32_t lo = sweep == 0 ? i0 : Hb - i0 - n;
			const unsigned slot = issued % kSweepSlots;
			d_mbar_expect_tx(&s_bar[slot], n * Wb);
			d_tma_load_1d(ring + (size_t) slot * chunk_bytes, src + slice + (size_t) lo * Wb, n * Wb, &s_bar[slot]);
			++issued;
		};
		if (XDIR == 0 && sweep == 1 && !split) {
			// the second sweep re-reads what this CTA just wrote with ordinary stores: order them before the async-proxy reads
			__threadfence();
			asm volatile("fence.proxy.async;" ::: "memory");
		}
		if (t == 0) {
			buf0[-1] = make_uint2(kFar, kFar); buf1[-1] = make_uint2(kFar, kFar);
			buf0[W4] = make_uint2(kFar, kFar); buf1[W4] = make_uint2(kFar, kFar);
		}
		if (active) buf0[t] = make_uint2(kFar, kFar);        // "row -1": nothing behind the first row
		__syncthreads();
		const unsigned consumed = issued;       // uniform bookkeeping of the copy sequence number (every thread tracks it)
		if (t == 0)
			for (uint32_t c = 0; c < (uint32_t) kSweepSlots && c < nchunks; ++c) issue(c);
		for (uint32_t c = 0; c < nchunks; ++c) {
			const unsigned seq = consumed + c, slot = seq % kSweepSlots;
			d_mbar_wait(&s_bar[slot], (seq / kSweepSlots) & 1u);
			const uint32_t  n  = min((uint32_t) kSweepRows, Hb - c * kSweepRows);
			const unsigned *cb = reinterpret_cast<const unsigned *>(ring + (size_t) slot * chunk_bytes);
			// rows alternate between the two row buffers; a chunk has an even number of rows unless it is the last one, so the
			// buffer roles are compile-time constants of the 2-row unrolled loop
			const unsigned *cp = cb + (sweep == 0 ? 0 : (size_t) (n - 1) * W4) + t;        // this thread's word in the chunk's first row
			auto row_step = [&](const uint2 *prev, uint2 *now) {
				if (active) {
					const unsigned gw = *cp;
					const uint2    p  = prev[t];                       // cells 0,1 | 2,3 of this thread in the previous row
					const unsigned mid = __funnelshift_r(p.x, p.y, 16);        // cells 1,2
					unsigned       ma = p.x, mb = p.y;
					if (XDIR == 0) {
						ma = __vimin3_u16x2(p.x, __funnelshift_l(prev[(int) t - 1].y, p.x, 16), mid);        // cells -1,0 and 1,2
						mb = __vimin3_u16x2(p.y, mid, __funnelshift_r(p.y, prev[t + 1].x, 16));              // cells 1,2 and 3,4
					} else if (XDIR > 0) {
						ma = __vminu2(p.x, mid);
						mb = __vminu2(p.y, __funnelshift_r(p.y, prev[t + 1].x, 16));
					} else {
						ma = __vminu2(p.x, __funnelshift_l(prev[(int) t - 1].y, p.x, 16));
						mb = __vminu2(p.y, mid);
					}
					const unsigned va = __viaddmin_u16x2(ma, 0x00010001u, d_prmt(gw, 0u, 0x4140u));
					const unsigned vb = __viaddmin_u16x2(mb, 0x00010001u, d_prmt(gw, 0u, 0x4342u));
					now[t] = make_uint2(va, vb);
					*dp    = d_prmt(va, vb, 0x6420u);
					dp += step;
					cp += step;
				}
				__syncthreads();
			};
			uint32_t k = 0;
			for (; k + 2 <= n; k += 2) {
				row_step(buf0, buf1);
				row_step(buf1, buf0);
			}
			if (k < n) row_step(buf0, buf1);        // odd tail: only ever in the last chunk of a sweep
			// every thread is past the chunk: its ring slot may be refilled
			if (t == 0 && c + kSweepSlots < nchunks) issue(c + kSweepSlots);
		}
		issued = consumed + nchunks;        // keep every thread's view of the sequence number in step with thread 0's
		if (split) break;                   // one sweep per CTA
	}
}

// ---- z pass as a walk along the line --------------------------------------------------------------------------
// One-sided result towards +z: F(z) = min_{j >= z} max(j - z, h(j)).  Stepping from z + 1 to z every candidate's cost
// grows by at most one, so with r = F(z + 1):   F(z) = min( h(z), r      if some j in [z+1, z+r] has h(j) <= r
//                                                                 r + 1  otherwise ),
// i.e. ONE range-minimum query per cell instead of an 8-step binary search; the two-sided value is min(F, B) with B the
// mirror image.  A CTA owns TW adjacent columns: it stages them and builds the sparse range-minimum table four columns
// per thread (32-bit shared-memory accesses, packed byte minimum).  The walk itself is cut into segments of kWalkSeg
// cells: a walker (one thread per column, direction and segment) finds the value at the head of its segment with the
// 8-step binary search and walks from there, so a line of L cells is 2 L / kWalkSeg independent dependency chains of
// ~kWalkSeg + 8 queries instead of two chains of L — the CTA's shared-memory tables are shared by up to 1024 walkers
// (with one walker per column and direction a long line left most of the SM idle).  Result rows leave as 32-bit words.
// MODE 0: dst0 = min(F, B) | 3: dst0 = F (towards +z), dst1 = B (towards -z).   Needs Wb % 4 == 0 and TW % 4 == 0.
constexpr int kWalkSeg = 32;

// F(z) towards +z (SIDE > 0) or -z (SIDE < 0) from scratch: smallest r with min h over the r + 1 cells from z on <= r.
// Only radii <= 254 are ever tested (hi <= 255), so windows have at most 255 cells: table levels 0..7.
template <int SIDE>
__device__ __forceinline__ unsigned walk_head(const uint8_t *__restrict__ T, int L, int TW, size_t lst, int z, int col, unsigned h0)
{
	unsigned lo = 0, hi = min(h0, 255u);
	while (lo < hi) {
		const int r = (int) ((lo + hi) >> 1);
		const int a = SIDE > 0 ? z : max(z - r, 0);
		const int b = SIDE > 0 ? min(z + r, L - 1) : z;
		const int k = 31 - __clz(b - a + 1);
		const uint8_t *Tk = T + (size_t) k * lst;
		const unsigned w  = min((unsigned) Tk[a * TW + col], (unsigned) Tk[(b - (1 << k) + 1) * TW + col]);
		if (w <= (unsigned) r) hi = (unsigned) r;
		else lo = (unsigned) r + 1;
	}
	return lo;
}

// src2 (or null): a second input, the cell-wise minimum of the two is what gets transformed (the split y sweeps' outputs)
template <int MODE>
__global__ void __launch_bounds__(1024) zwalk_kernel(const uint8_t *__restrict__ src, const uint8_t *__restrict__ src2, uint8_t *__restrict__ dst0,
                                                     uint8_t *__restrict__ dst1, uint32_t Wb, uint32_t L, size_t line_stride, size_t outer_stride, int TW,
                                                     int nlev)
{
	extern __shared__ __align__(16) uint8_t T[];        // nlev levels of L x TW bytes, then the F and B lines (2 x L x TW)
	const int      nthreads = blockDim.x;
	const int      TW4 = TW >> 2;                       // words per row
	const uint32_t x0  = blockIdx.x * TW;
	const size_t   base = (size_t) blockIdx.y * outer_stride + x0;
	const size_t   lst  = (size_t) L * TW;
	uint8_t       *Fl = T + (size_t) nlev * lst, *Bl = Fl + lst;
	// stage: word w of row p = columns 4w .. 4w+3
	const int nwords = (int) L * TW4;
	for (int i = threadIdx.x; i < nwords; i += nthreads) {
		const int p = i / TW4, w = i - p * TW4;
		unsigned  v = 0xffffffffu;
		if (x0 + 4u * w < Wb) {
			v = __ldg(reinterpret_cast<const unsigned *>(src + base + (size_t) p * line_stride) + w);
			if (src2) v = __vminu4(v, __ldg(reinterpret_cast<const unsigned *>(src2 + base + (size_t) p * line_stride) + w));
		}
		reinterpret_cast<unsigned *>(T)[i] = v;
	}
	__syncthreads();
	for (int k = 1; k < nlev; ++k) {
		const unsigned *prev = reinterpret_cast<const unsigned *>(T + (size_t) (k - 1) * lst);
		unsigned       *cur  = reinterpret_cast<unsigned *>(T + (size_t) k * lst);
		const int       half = (1 << (k - 1)) * TW4;        // in words
		for (int i = threadIdx.x; i < nwords; i += nthreads) {
			const unsigned a = prev[i];
			const unsigned b = i + half < nwords ? prev[i + half] : 0xffffffffu;
			cur[i] = __vminu4(a, b);
		}
		__syncthreads();
	}
	{
		const int Li = (int) L, nseg = (Li + kWalkSeg - 1) / kWalkSeg;
		const int nwalk = 2 * nseg * TW;        // walker id = (direction * nseg + segment) * TW + column
		for (int wk = threadIdx.x; wk < nwalk; wk += nthreads) {
			const int col = wk % TW, sd = wk / TW, seg = sd % nseg;
			const int z_lo = seg * kWalkSeg, z_hi = min(z_lo + kWalkSeg, Li) - 1;        // inclusive
			if (sd < nseg) {        // F: towards +z, walking down from the segment's last cell
				unsigned r = T[z_hi * TW + col];
				if (z_hi != Li - 1) r = walk_head<1>(T, Li, TW, lst, z_hi, col, r);
				Fl[z_hi * TW + col] = (uint8_t) r;
				for (int z = z_hi - 1; z >= z_lo; --z) {
					const unsigned hz = T[z * TW + col];
					unsigned cand = min(r + 1u, 255u);
					if (r > 0u && hz > r) {        // h(z) <= r decides F(z) = h(z) whatever the window holds
						const int a = z + 1, b = min(z + (int) r, Li - 1);
						const int k = 31 - __clz(b - a + 1);
						const uint8_t *Tk = T + (size_t) k * lst;
						const unsigned w  = min((unsigned) Tk[a * TW + col], (unsigned) Tk[(b - (1 << k) + 1) * TW + col]);
						if (w <= r) cand = r;
					}
					r = min(hz, cand);
					Fl[z * TW + col] = (uint8_t) r;
				}
			} else {                // B: towards -z, walking up from the segment's first cell
				unsigned r = T[z_lo * TW + col];
				if (z_lo != 0) r = walk_head<-1>(T, Li, TW, lst, z_lo, col, r);
				Bl[z_lo * TW + col] = (uint8_t) r;
				for (int z = z_lo + 1; z <= z_hi; ++z) {
					const unsigned hz = T[z * TW + col];
					unsigned cand = min(r + 1u, 255u);
					if (r > 0u && hz > r) {
						const int b = z - 1, a = max(z - (int) r, 0);
						const int k = 31 - __clz(b - a + 1);
						const uint8_t *Tk = T + (size_t) k * lst;
						const unsigned w  = min((unsigned) Tk[a * TW + col], (unsigned) Tk[(b - (1 << k) + 1) * TW + col]);
						if (w <= r) cand = r;
					}
					r = min(hz, cand);
					Bl[z * TW + col] = (uint8_t) r;
				}
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < nwords; i += nthreads) {
		const int p = i / TW4, w = i - p * TW4;
		if (x0 + 4u * w >= Wb) continue;
		const unsigned f = reinterpret_cast<const unsigned *>(Fl)[i], b = reinterpret_cast<const unsigned *>(Bl)[i];
		const size_t   o = base + (size_t) p * line_stride;
		if (MODE == 0) reinterpret_cast<unsigned *>(dst0 + o)[w] = __vminu4(f, b);
		else {
			reinterpret_cast<unsigned *>(dst0 + o)[w] = f;
			reinterpret_cast<unsigned *>(dst1 + o)[w] = b;
		}
	}
}

// zb_first / zb_count: the block slices to transform (the whole map by default; a rank's z-slab in the sharded build)
template <int DIR>
static int run_xpass(const vkv_volume *vol, const uint8_t *O, uint8_t *out, cudaStream_t s, uint32_t zb_first = 0, uint32_t zb_count = 0xffffffffu)
{
	const uint32_t Wb    = vol->dim_b[0];
	if (zb_count == 0xffffffffu) zb_count = vol->dim_b[2];
	const uint64_t nrows = (uint64_t) vol->dim_b[1] * zb_count;
	O += (size_t) zb_first * vol->dim_b[1] * Wb;
	out += (size_t) zb_first * vol->dim_b[1] * Wb;
	if (Wb <= kRowWordsMax * 32) {
		const int grid = (int) std::min<uint64_t>((nrows + 7) / 8, (uint64_t) vol->ctx->sm_count * 8);
		if (Wb % 4 == 0) {
			const uint32_t nseg = (Wb + 127) / 128;
			if (nseg <= 2) xpass_vec4_kernel<DIR, 2><<<grid, 256, 0, s>>>(O, out, Wb, nrows);
			else if (nseg <= 4) xpass_vec4_kernel<DIR, 4><<<grid, 256, 0, s>>>(O, out, Wb, nrows);
			else if (nseg <= 8) xpass_vec4_kernel<DIR, 8><<<grid, 256, 0, s>>>(O, out, Wb, nrows);
			else xpass_vec4_kernel<DIR, 16><<<grid, 256, 0, s>>>(O, out, Wb, nrows);
		}
		else xpass_ballot_kernel<DIR><<<grid, 256, 0, s>>>(O, out, Wb, nrows);
	} else {
		xpass_serial_kernel<DIR><<<(unsigned) ((nrows + 127) / 128), 128, 0, s>>>(O, out, Wb, nrows);
	}
	VKV_LAUNCHED();
	return VKV_OK;
}

// axis 1 = y lines, axis 2 = z lines
template <int DIR, int NOUT>
static int run_minmax(const vkv_volume *vol, int axis, const uint8_t *src, uint8_t *dst0, uint8_t *dst1, cudaStream_t s)
{
	const uint32_t Wb = vol->dim_b[0], Hb = vol->dim_b[1], Db = vol->dim_b[2];
	const uint32_t L            = axis == 1 ? Hb : Db;
	const size_t   line_stride  = axis == 1 ? (size_t) Wb : (size_t) Wb * Hb;
	const size_t   outer_stride = axis == 1 ? (size_t) Wb * Hb : (size_t) Wb;
	const dim3     grid((Wb + 31) / 32, axis == 1 ? Db : Hb);
	constexpr int  MODE = DIR == 0 ? 0 : (NOUT == 2 ? 3 : (DIR > 0 ? 1 : 2));
	// table levels: the largest window is min(2*255+1, L) cells (one-sided: min(256, L))
	const uint32_t max_window = std::min<uint32_t>(DIR == 0 ? 511u : 256u, L);
	int            nlev       = 1;
	while ((2u << (nlev - 1)) <= max_window) ++nlev;        // nlev = floor(log2(max_window)) + 1
	int TW = 32;
	while (TW > 8 && (size_t) nlev * L * TW > (size_t) 200 * 1024) TW >>= 1;
	const size_t smem = (size_t) nlev * L * TW;
	if (smem <= (size_t) 200 * 1024) {
		static PerDeviceOnce configured;        // opt in to > 48 KB of dynamic shared memory once per instantiation
		if (configured.first(vol->ctx->device)) {
			VKV_CUDA_CHECK(cudaFuncSetAttribute(minmax_rmq_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		}
		const dim3 grid_rmq((Wb + TW - 1) / TW, axis == 1 ? Db : Hb);
		minmax_rmq_kernel<MODE><<<grid_rmq, 256, smem, s>>>(src, dst0, dst1, Wb, L, line_stride, outer_stride, TW, nlev);
	} else if (L <= (uint32_t) kLineMax) {
		minmax_pass_kernel<DIR, NOUT, true><<<grid, 256, (size_t) L * 32, s>>>(src, dst0, dst1, Wb, L, line_stride, outer_stride);
	} else {
		minmax_pass_kernel<DIR, NOUT, false><<<grid, 256, 0, s>>>(src, dst0, dst1, Wb, L, line_stride, outer_stride);
	}
	VKV_LAUNCHED();
	return VKV_OK;
}

template <int XDIR>
static int run_ysweep(const vkv_volume *vol, const uint8_t *g, uint8_t *dst0, uint8_t *dst1, cudaStream_t s, bool *done, uint32_t zb_first = 0,
                      uint32_t zb_count = 0xffffffffu, bool split = false)
{
	const uint32_t Wb = vol->dim_b[0], Hb = vol->dim_b[1];
	const uint32_t Db = zb_count == 0xffffffffu ? vol->dim_b[2] : zb_count;
	*done = false;
	if (Db == 0) { *done = true; return VKV_OK; }
	{
		const size_t off = (size_t) zb_first * Wb * Hb;        // slices are independent: a slab is a smaller map
		g += off; dst0 += off;
		if (dst1) dst1 += off;
	}
	// four cells per thread; bulk copies need 16-byte aligned row blocks (slice size, hence every block start and length)
	if (Wb > 1024u || Wb % 4 != 0 || ((size_t) Wb * Hb) % 16 != 0 || ((size_t) kSweepRows * Wb) % 16 != 0 || (reinterpret_cast<uintptr_t>(g) % 16) != 0 ||
	    (reinterpret_cast<uintptr_t>(dst0) % 16) != 0 || (dst1 && (reinterpret_cast<uintptr_t>(dst1) % 16) != 0))
		return VKV_OK;        // otherwise the search kernel
	const int    threads = (int) ((Wb / 4 + 31u) / 32u * 32u);
	const size_t smem    = (size_t) kSweepSlots * kSweepRows * Wb + 2 * ((size_t) Wb / 4 + 2) * sizeof(uint2) + 16;
	static PerDeviceOnce configured;
	if (configured.first(vol->ctx->device)) {
		VKV_CUDA_CHECK(cudaFuncSetAttribute(ysweep_kernel<XDIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
	}
	ysweep_kernel<XDIR><<<split ? 2 * Db : Db, threads, smem, s>>>(g, dst0, dst1, Wb, Hb, split ? 1 : 0);
	VKV_LAUNCHED();
	*done = true;
	return VKV_OK;
}

// z lines (axis 2); MODE 0 two-sided -> dst0, MODE 3 both one-sided results
// yb_first / yb_count: the block rows whose z lines are transformed (all by default; a rank's share in the sharded build)
template <int MODE>
static int run_zwalk(const vkv_volume *vol, const uint8_t *src, uint8_t *dst0, uint8_t *dst1, cudaStream_t s, bool *done, uint32_t yb_first = 0,
                     uint32_t yb_count = 0xffffffffu, const uint8_t *src2 = nullptr)
{
	const uint32_t Wb = vol->dim_b[0], Hb = vol->dim_b[1], L = vol->dim_b[2];
	const uint32_t rows = yb_count == 0xffffffffu ? Hb : yb_count;
	*done = false;
	if (rows == 0) { *done = true; return VKV_OK; }
	{
		const size_t off = (size_t) yb_first * Wb;
		src += off; dst0 += off;
		if (dst1) dst1 += off;
		if (src2) src2 += off;
	}
	if (Wb % 4 != 0 || Hb > 65535u) return VKV_OK;        // 32-bit column groups; otherwise the search kernel
	const uint32_t max_window = std::min<uint32_t>(255u, L);
	int            nlev       = 1;
	while ((2u << (nlev - 1)) <= max_window) ++nlev;
	// 32 columns per CTA (a warp of walkers reads 32 consecutive bytes of a table row) unless the tables then exceed what
	// lets a few CTAs share an SM; narrower tiles for long lines
	const size_t per_col = (size_t) (nlev + 2) * L;
	int          TW      = 32;
	while (TW > 8 && per_col * TW > (size_t) 100 * 1024) TW >>= 1;
	if (per_col * TW <= (size_t) 50 * 1024 && Wb >= 64 && L >= 4u * kWalkSeg) TW = 64;
	const size_t smem = per_col * TW;
	if (smem > (size_t) 200 * 1024) return VKV_OK;
	static PerDeviceOnce configured;
	if (configured.first(vol->ctx->device)) {
		VKV_CUDA_CHECK(cudaFuncSetAttribute(zwalk_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	}
	const int  nseg    = (int) ((L + kWalkSeg - 1) / kWalkSeg);
	const int  threads = std::min(1024, std::max(64, (2 * nseg * TW + 31) / 32 * 32));
	const dim3 grid((Wb + TW - 1) / TW, rows);
	zwalk_kernel<MODE><<<grid, threads, smem, s>>>(src, src2, dst0, dst1, Wb, L, (size_t) Wb * Hb, (size_t) Wb, TW, nlev);
	VKV_LAUNCHED();
	*done = true;
	return VKV_OK;
}

int launch_distance(vkv_volume *vol, int skipping_type, cudaStream_t s)
{
	int rc;
	if (skipping_type == VKV_SKIP_DISTANCE) {
		uint8_t *map = vol->d_maps[0];        // holds the occupancy map on entry, the distance map on exit
		const bool legacy = getenv("VKV_DIST_SEARCH") != nullptr;        // A/B switch: the binary-search kernels for every pass
		bool       done   = false;
		if ((rc = run_xpass<0>(vol, map, vol->d_tmp, s))) return rc;
		// small maps (L2-resident): the two y sweeps of a slice as two CTAs, their outputs (d_swap: sources above, the map itself —
		// its occupancy was consumed by the x pass — : sources below) minimised by the z walk while it stages.  The z walk reads a
		// column completely before it writes it, so using the map as the second input is safe.
		const bool split_ok = !legacy && !getenv("VKV_DIST_NOSPLIT") && vol->M <= (size_t) 48 << 20 && vol->dim_b[0] % 4 == 0 && vol->dim_b[1] <= 65535u && vol->dim_b[2] <= 1024u;
		bool       split    = false;
		if (split_ok) {
			if ((rc = run_ysweep<0>(vol, vol->d_tmp, vol->d_swap, map, s, &done, 0, 0xffffffffu, true))) return rc;
			split = done;
		}
		if (!split && !legacy && (rc = run_ysweep<0>(vol, vol->d_tmp, vol->d_swap, nullptr, s, &done))) return rc;
		if (!done && (rc = run_minmax<0, 1>(vol, 1, vol->d_tmp, vol->d_swap, nullptr, s))) return rc;
		done = false;
		if (!legacy && (rc = run_zwalk<0>(vol, vol->d_swap, map, nullptr, s, &done, 0, 0xffffffffu, split ? map : nullptr))) return rc;
		if (split && !done) {
			set_error("launch_distance: split y sweeps need the z walk kernel");
			return VKV_ERR_STATE;
		}
		if (!done && (rc = run_minmax<0, 1>(vol, 2, vol->d_swap, map, nullptr, s))) return rc;
	} else if (skipping_type == VKV_SKIP_ANISOTROPIC_DISTANCE) {
		// map index i = 4[x-] + 2[y-] + [z-]  (compute_distance_map.cpp:228-252); occupancy lives in map 7
		uint8_t *const *m = vol->d_maps.data();
		const bool legacy = getenv("VKV_DIST_SEARCH") != nullptr;
		bool       ok_y = false, ok_z = false;
		// x+ half: maps 0..3.  The y sweep gives both y directions of the same x pass (y+ -> swap, y- -> map 3, which the
		// last z walk of this half then overwrites in place: a CTA stages its columns before it writes them).
		if ((rc = run_xpass<1>(vol, m[7], vol->d_tmp, s))) return rc;
		if (!legacy && (rc = run_ysweep<1>(vol, vol->d_tmp, vol->d_swap, m[3], s, &ok_y))) return rc;
		if (ok_y) {
			if ((rc = run_zwalk<3>(vol, vol->d_swap, m[0], m[1], s, &ok_z))) return rc;
			if (!ok_z && (rc = run_minmax<1, 2>(vol, 2, vol->d_swap, m[0], m[1], s))) return rc;
			if (ok_z) { if ((rc = run_zwalk<3>(vol, m[3], m[2], m[3], s, &ok_z))) return rc; }
			else {
				VKV_CUDA_CHECK(cudaMemcpyAsync(vol->d_swap, m[3], vol->M, cudaMemcpyDeviceToDevice, s));
				if ((rc = run_minmax<1, 2>(vol, 2, vol->d_swap, m[2], m[3], s))) return rc;
			}
		} else {
			if ((rc = run_minmax<1, 1>(vol, 1, vol->d_tmp, vol->d_swap, nullptr, s))) return rc;
			if ((rc = run_minmax<1, 2>(vol, 2, vol->d_swap, m[0], m[1], s))) return rc;
			if ((rc = run_minmax<-1, 1>(vol, 1, vol->d_tmp, vol->d_swap, nullptr, s))) return rc;
			if ((rc = run_minmax<1, 2>(vol, 2, vol->d_swap, m[2], m[3], s))) return rc;
		}
		// x- half: maps 4..7 (map 7 — the occupancy — is consumed by the x pass and overwritten last, as in the reference)
		if ((rc = run_xpass<-1>(vol, m[7], vol->d_tmp, s))) return rc;
		ok_y = ok_z = false;
		if (!legacy && (rc = run_ysweep<-1>(vol, vol->d_tmp, vol->d_swap, m[7], s, &ok_y))) return rc;
		if (ok_y) {
			if ((rc = run_zwalk<3>(vol, vol->d_swap, m[4], m[5], s, &ok_z))) return rc;
			if (!ok_z && (rc = run_minmax<1, 2>(vol, 2, vol->d_swap, m[4], m[5], s))) return rc;
			if (ok_z) { if ((rc = run_zwalk<3>(vol, m[7], m[6], m[7], s, &ok_z))) return rc; }
			else {
				VKV_CUDA_CHECK(cudaMemcpyAsync(vol->d_swap, m[7], vol->M, cudaMemcpyDeviceToDevice, s));
				if ((rc = run_minmax<1, 2>(vol, 2, vol->d_swap, m[6], m[7], s))) return rc;
			}
		} else {
			if ((rc = run_minmax<1, 1>(vol, 1, vol->d_tmp, vol->d_swap, nullptr, s))) return rc;
			if ((rc = run_minmax<1, 2>(vol, 2, vol->d_swap, m[4], m[5], s))) return rc;
			if ((rc = run_minmax<-1, 1>(vol, 1, vol->d_tmp, vol->d_swap, nullptr, s))) return rc;
			if ((rc = run_minmax<1, 2>(vol, 2, vol->d_swap, m[6], m[7], s))) return rc;
		}
	}
	// NONE / BLOCK: the occupancy map itself is what the ray caster reads (compute_distance_map.cpp:96-99)
	return VKV_OK;
}

// ---- the isotropic transform cut in two for the multi-GPU build (group.cu) ------------------------------------------------------
// The x and y passes only ever look inside one z slice, the z pass only inside one column: a rank runs x + y on its own z-slab
// (occupancy in map 0 -> xy-intermediate in d_swap), the slabs are exchanged, and the z pass runs on the rank's share of the
// block rows (d_swap -> map 0) before the second exchange.  Only for shapes the sweep / walk kernels cover.
bool distance_shardable(const vkv_volume *vol)
{
	const uint32_t Wb = vol->dim_b[0], Hb = vol->dim_b[1];
	return Wb <= 1024u && Wb % 4 == 0 && ((size_t) Wb * Hb) % 16 == 0 && ((size_t) kSweepRows * Wb) % 16 == 0 && Hb <= 65535u && Wb <= kRowWordsMax * 32u &&
	       !getenv("VKV_DIST_SEARCH");
}

int launch_distance_xy_slab(vkv_volume *vol, uint32_t zb_first, uint32_t zb_count, cudaStream_t s)
{
	int  rc;
	bool done = false;
	if (zb_count == 0) return VKV_OK;
	if ((rc = run_xpass<0>(vol, vol->d_maps[0], vol->d_tmp, s, zb_first, zb_count))) return rc;
	if ((rc = run_ysweep<0>(vol, vol->d_tmp, vol->d_swap, nullptr, s, &done, zb_first, zb_count))) return rc;
	if (!done) {
		set_error("launch_distance_xy_slab: shape not covered by the sweep kernel");
		return VKV_ERR_STATE;
	}
	return VKV_OK;
}

int launch_distance_z_rows(vkv_volume *vol, uint32_t yb_first, uint32_t yb_count, cudaStream_t s)
{
	int  rc;
	bool done = false;
	if ((rc = run_zwalk<0>(vol, vol->d_swap, vol->d_maps[0], nullptr, s, &done, yb_first, yb_count))) return rc;
	if (!done) {
		set_error("launch_distance_z_rows: shape not covered by the walk kernel");
		return VKV_ERR_STATE;
	}
	return VKV_OK;
}

}        // namespace vkv
