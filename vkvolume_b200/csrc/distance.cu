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
// B200 design (round 2; every pass of round 1 turned out to be bound by instruction issue or by the latency of a short dependent
// chain, never by bytes — profiles/r2_k3_summary.md — so each pass is built around its instruction count):
//   * x pass: the row is turned into a bit mask and each lane finds the nearest set bit on either side of its 4, 8 or 16 adjacent
//     cells with two O(1) searches (a 64-bit mask of the non-zero mask words names the word that holds it); cells in between by
//     recurrence; one vector load and store per lane, the next row prefetched.  At the HBM roofline on maps past the L2.
//   * y pass: chamfer recurrence, a whole row per step.  Wide maps: a warp owns a strip of 256 cells in registers (u16x2 lanes,
//     neighbours by shuffle, VIMNMX3 / VIADDMNMX), strips exchange edge cells through shared memory every 16 rows.  Narrow
//     isotropic maps with few slices: one CTA per sweep, previous row in shared memory, rows streamed through a bulk-TMA ring.
//   * z pass: a walk along the line, F(z) = min(h(z), F(z+1) or F(z+1) + 1), decided by a per-value "last seen at step" table in
//     shared memory (no search, no range-minimum tables); two walker warps per 32 columns, one per direction, meeting in the middle.
//   * the anisotropic build shares the 2 x passes and 4 y passes between the 8 octant maps
//     (same sharing as the reference's 14-dispatch schedule).
//   * shapes the fast kernels do not cover (Wb % 4 != 0, rows over 2048 cells, ...) fall back to a binary search on the radius against
//     a sparse range-minimum table (minmax_rmq_kernel) or the literal search (minmax_pass_kernel).
// Algorithmic bytes 6 B/block (isotropic) and 28 B/block (anisotropic) as in SURVEY §8(d).
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

// CPL cells per lane (4, 8 or 16: one 4-, 8- or 16-byte load and store), SEGS: compile-time bound on the row's segments of 32 CPL
// cells.  The more cells a lane owns, the fewer nearest-bit searches a row costs (two per lane and segment); a row's words are
// all requested before any is looked at, and the next row's before this row is worked on (one DRAM round trip per row, hidden
// behind the previous row's arithmetic).  Rows without a single occupied cell are answered with one store per lane.
template <int CPL> struct XWords;
template <> struct XWords<4>  { using T = unsigned; };
template <> struct XWords<8>  { using T = uint2; };
template <> struct XWords<16> { using T = uint4; };

template <int DIR, int CPL, int SEGS>
__global__ void __launch_bounds__(256) xpass_lanes_kernel(const uint8_t *__restrict__ O, uint8_t *__restrict__ out, uint32_t Wb, uint64_t nrows)
{
	using V = typename XWords<CPL>::T;
	constexpr int WPL = CPL / 4, LPW = 32 / CPL;        // words per lane; lanes per 32-cell mask word
	__shared__ unsigned s_words[8][kRowWordsMax];
	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned      *words = s_words[warp];
	const uint32_t nseg  = (Wb + 32 * CPL - 1) / (32 * CPL);
	for (int i = lane; i < kRowWordsMax; i += 32) words[i] = 0;
	__syncwarp();
	union Cells { V v; unsigned w[WPL]; };
	auto request = [&](Cells (&c)[SEGS], uint64_t row) {
		const uint8_t *src = O + row * Wb;
#pragma unroll
		for (uint32_t sgm = 0; sgm < (uint32_t) SEGS; ++sgm) {
			const uint32_t x = (sgm * 32 + lane) * CPL;
#pragma unroll
			for (int w = 0; w < WPL; ++w) c[sgm].w[w] = 0xffffffffu;
			if (sgm < nseg && x < Wb) c[sgm].v = __ldg(reinterpret_cast<const V *>(src + x));
		}
	};
	const uint64_t stride = (uint64_t) gridDim.x * 8;
	uint64_t       row    = (uint64_t) blockIdx.x * 8 + warp;
	Cells cur[SEGS], nxt[SEGS];
	if (row < nrows) request(cur, row);
	for (; row < nrows; row += stride) {
		if (row + stride < nrows) request(nxt, row + stride);
		unsigned occ[SEGS];        // this lane's CPL occupancy bits per segment
#pragma unroll
		for (uint32_t sgm = 0; sgm < (uint32_t) SEGS; ++sgm) {
			if (sgm >= nseg) break;
			unsigned bits = 0;
#pragma unroll
			for (int w = 0; w < WPL; ++w) {
				const unsigned c = cur[sgm].w[w];
				unsigned       t = (c & 0x7f7f7f7fu) + 0x7f7f7f7fu;        // exact zero-byte detector -> 0x80 per zero byte
				t                = ~(t | c | 0x7f7f7f7fu);
				const unsigned y = t >> 7;
				bits |= ((y | (y >> 7) | (y >> 14) | (y >> 21)) & 0xfu) << (4 * w);
			}
			occ[sgm]   = bits;
			unsigned v = bits << (CPL * (lane % LPW));
#pragma unroll
			for (int o = 1; o < LPW; o <<= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
			if (lane % LPW == 0) words[sgm * CPL + lane / LPW] = v;
		}
		__syncwarp();
		RowBits rb;
		rb.words = words;
		rb.nz    = (unsigned long long) __ballot_sync(0xffffffffu, words[lane] != 0u) |
		        ((unsigned long long) __ballot_sync(0xffffffffu, words[32 + lane] != 0u) << 32);
		uint8_t *dst = out + row * Wb;
#pragma unroll
		for (uint32_t sgm = 0; sgm < (uint32_t) SEGS; ++sgm) {
			if (sgm >= nseg) break;
			const uint32_t x = (sgm * 32 + lane) * CPL;
			if (x < Wb) {
				Cells res;
				if (rb.nz == 0ull) {        // nothing in this row: every distance saturates
#pragma unroll
					for (int w = 0; w < WPL; ++w) res.w[w] = 0xffffffffu;
				} else {
					const int      w0 = (int) (x >> 5), b = (int) (x & 31);
					const unsigned oc = occ[sgm];
					unsigned       d[CPL];
#pragma unroll
					for (int k = 0; k < CPL; ++k) d[k] = 255u;
					if (DIR >= 0) {
						unsigned r = rb.right(w0, b + CPL - 1);
						d[CPL - 1] = r;
#pragma unroll
						for (int k = CPL - 2; k >= 0; --k) {
							r    = ((oc >> k) & 1u) ? 0u : __viaddmin_u32(r, 1u, 255u);
							d[k] = r;
						}
					}
					if (DIR <= 0) {
						unsigned l = rb.left(w0, b);
						d[0]       = min(d[0], l);
#pragma unroll
						for (int k = 1; k < CPL; ++k) {
							l    = ((oc >> k) & 1u) ? 0u : __viaddmin_u32(l, 1u, 255u);
							d[k] = min(d[k], l);
						}
					}
#pragma unroll
					for (int w = 0; w < WPL; ++w) res.w[w] = d[4 * w] | (d[4 * w + 1] << 8) | (d[4 * w + 2] << 16) | (d[4 * w + 3] << 24);
				}
				*reinterpret_cast<V *>(dst + x) = res.v;
			}
		}
		__syncwarp();
#pragma unroll
		for (uint32_t sgm = 0; sgm < (uint32_t) SEGS; ++sgm) cur[sgm] = nxt[sgm];
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
//
// What bounds a sweep is the latency of one row step times the rows of a slice, so the step is kept inside a WARP: a warp
// owns a strip of 256 adjacent cells, eight per lane, the previous row lives in registers as 16-bit lanes (four u16x2 words
// per lane), the x +- 1 neighbours across lanes come from two shuffles, and the step itself is a handful of native packed
// instructions (funnel shifts, VIMNMX3.U16x2, VIADDMNMX.U16x2: min(g, m + 1) needs no saturation because m <= 255) —
// no shared memory, no barrier.  Rows wider than a strip are covered by overlapping strips: a strip's outer 32 cells
// on either side are a halo it computes but does not own; what a strip does not know about its neighbours spoils one
// more halo cell per row, so every 32 rows the warps of a slice exchange their edge cells through shared memory (one
// barrier per 32 rows instead of one per row).  Input rows are requested eight rows ahead, straight into registers.
// (The kernel this replaces — one CTA per slice, previous row in shared memory, one barrier per row, rows streamed through a
// bulk-TMA ring — spent 430 cycles per row step and 0.6 warp instructions per cell: profiles/r2v_k3_c5.)
// `split` (isotropic, small maps): the two sweeps of a slice are independent when both start from g — the result is then
// min(up, down), which the z pass takes while it stages — so they run as two CTAs (blockIdx.x = 2 slice + sweep) and the serial
// chain of a slice is Hb row steps instead of 2 Hb; costs one more map of traffic, so only where the maps live in the L2.
constexpr int kStripHalo = 16, kStripBlock = 16;        // kStripBlock <= kStripHalo
__device__ __forceinline__ unsigned d_prmt(unsigned a, unsigned b, unsigned sel)
{
	unsigned r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}

// XDIR 0: isotropic (two-sided x, 3 neighbours, in-place second sweep); +-1: one-sided.
// CPL: cells per lane, 8 (strips of 256 cells, 224 owned; the one launched) or 4 (128 / 96).
template <int XDIR, int CPL>
__global__ void __launch_bounds__(512) ysweep_kernel(const uint8_t *__restrict__ g, uint8_t *__restrict__ dst0, uint8_t *__restrict__ dst1,
                                                     uint32_t Wb, uint32_t Hb, int split, int nstrips)
{
	constexpr int NW = CPL / 2, NL = CPL / 4;                  // u16x2 words of state, 32-bit words of cells per lane
	extern __shared__ __align__(16) unsigned s_edge[];        // 2 x (Wb / 2 + 8) u16x2 words: a row of F at a block boundary (double-buffered)
	constexpr unsigned kFar = 0x00ff00ffu;                     // 255 in both lanes
	const int      lane = threadIdx.x & 31, strip = threadIdx.x >> 5;
	const int      halo = nstrips > 1 ? kStripHalo : 0, own = 32 * CPL - 2 * halo;
	const int      x    = strip * own - halo + CPL * lane;        // the first of this lane's cells (may lie outside the row)
	const bool     owned = CPL * lane >= halo && CPL * lane < halo + own;
	// words outside the row read as "far": load some word of the row instead and force the bits on  (Wb % 4 == 0)
	bool     in[NL], st[NL];
	unsigned force[NL];
	int      xl[NL];
#pragma unroll
	for (int w = 0; w < NL; ++w) {
		in[w]    = x + 4 * w >= 0 && x + 4 * w < (int) Wb;
		force[w] = in[w] ? 0u : 0xffffffffu;
		xl[w]    = in[w] ? x + 4 * w : 0;
		st[w]    = owned && in[w];
	}
	const size_t   slice = (size_t) (split ? blockIdx.x >> 1 : blockIdx.x) * Wb * Hb;
	const int      edge_words = (int) (Wb >> 1) + 8;
	int            parity = 0;
	for (int sweep = split ? (int) (blockIdx.x & 1u) : 0; sweep < 2; ++sweep) {
		// sweep 0 walks y upwards (sources at y' <= y), sweep 1 downwards (sources at y' >= y)
		const bool      inplace = XDIR == 0 && sweep == 1 && !split;
		const uint8_t  *src = (inplace ? dst0 : g) + slice;
		uint8_t        *dst = ((XDIR == 0 && !split) ? dst0 : (sweep == 0 ? dst1 : dst0)) + slice;
		// The second sweep of the in-place variant reads what the first one wrote: its own cells, and — in the halo — cells of the
		// neighbouring warps, which may by then hold either sweep's value.  Both are right: the final value is min(first, 1 + ...),
		// so taking it as the base term reproduces it.
		if (inplace) __syncthreads();
		unsigned prev[NW];        // "row -1": nothing behind the first row
#pragma unroll
		for (int w = 0; w < NW; ++w) prev[w] = kFar;
		const ptrdiff_t rs = sweep == 0 ? (ptrdiff_t) Wb : -(ptrdiff_t) Wb;        // from one row of the sweep to the next
		const size_t    first = sweep == 0 ? 0 : (size_t) (Hb - 1) * Wb;
		const uint8_t  *lp = src + first;        // the next row to request
		uint8_t        *sp = dst + first;        // the next row to store
		uint32_t        requested = 0;
		// eight rows into registers (rows past the slice repeat its last row: requested, never used)
		auto load8 = [&](unsigned (&buf)[8][NL]) {
#pragma unroll
			for (int k = 0; k < 8; ++k) {
#pragma unroll
				for (int w = 0; w < NL; ++w)
					// (the isotropic variant's second sweep reads what the first wrote: L2 loads throughout — the rows are streamed once anyway,
					// and one load flavour keeps the instruction stream free of a predicated twin of every load)
					buf[k][w] = (XDIR == 0 ? __ldcg(reinterpret_cast<const unsigned *>(lp + xl[w])) : __ldg(reinterpret_cast<const unsigned *>(lp + xl[w]))) | force[w];
				if (++requested < Hb) lp += rs;
			}
		};
		auto step8 = [&](const unsigned (&buf)[8][NL], const uint32_t rows) {        // rows: how many of the eight exist (8 except in the last group)
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				if ((uint32_t) k >= rows) break;
				// the cells next to the lane's own live in the neighbouring lanes; beyond the strip: far (true at the row's ends,
				// repaired by the halo inside)
				unsigned left  = __shfl_up_sync(0xffffffffu, prev[NW - 1], 1);
				unsigned right = __shfl_down_sync(0xffffffffu, prev[0], 1);
				if (lane == 0) left = kFar;
				if (lane == 31) right = kFar;
				unsigned m[NW + 1];        // m[w]: cells 2w - 1, 2w
				m[0] = __funnelshift_l(left, prev[0], 16);
#pragma unroll
				for (int w = 1; w < NW; ++w) m[w] = __funnelshift_r(prev[w - 1], prev[w], 16);
				m[NW] = __funnelshift_r(prev[NW - 1], right, 16);
#pragma unroll
				for (int w = 0; w < NW; ++w) {
					unsigned n;
					if (XDIR == 0) n = __vimin3_u16x2(m[w], prev[w], m[w + 1]);
					else if (XDIR > 0) n = __vminu2(prev[w], m[w + 1]);
					else n = __vminu2(prev[w], m[w]);
					prev[w] = __viaddmin_u16x2(n, 0x00010001u, d_prmt(buf[k][w >> 1], 0u, (w & 1) ? 0x4342u : 0x4140u));
				}
#pragma unroll
				for (int w = 0; w < NL; ++w)
					if (st[w]) reinterpret_cast<unsigned *>(sp + x)[w] = d_prmt(prev[2 * w], prev[2 * w + 1], 0x6420u);
				sp += rs;
			}
		};
		// every kStripBlock rows the strips of the slice swap edges: owners publish the row they hold, halo lanes take it over
		auto exchange = [&]() {
			unsigned *e = s_edge + parity * edge_words + (x >> 1);
			if (owned) {
#pragma unroll
				for (int w = 0; w < NL; ++w)
					if (in[w]) *reinterpret_cast<uint2 *>(e + 2 * w) = make_uint2(prev[2 * w], prev[2 * w + 1]);
			}
			__syncthreads();
			if (!owned) {
#pragma unroll
				for (int w = 0; w < NL; ++w) {
					uint2 v = make_uint2(kFar, kFar);
					if (in[w]) v = *reinterpret_cast<const uint2 *>(e + 2 * w);
					prev[2 * w] = v.x; prev[2 * w + 1] = v.y;
				}
			}
			parity ^= 1;
		};
		unsigned ga[8][NL], gb[8][NL];
		load8(ga);
		for (uint32_t i0 = 0; i0 < Hb; i0 += 16) {
			load8(gb);
			step8(ga, min(8u, Hb - i0));
			load8(ga);
			if (i0 + 8 < Hb) step8(gb, min(8u, Hb - i0 - 8));
			if (nstrips > 1 && i0 + 16 < Hb) exchange();
		}
	}
}

// ---- y sweep, one CTA per slice (narrow maps) ------------------------------------------------------------------------------
// Round 1's formulation of the same recurrence, kept for maps a single strip wide with too few slices to fill the machine (the
// 208x208x124 map of config 2: 248 sweeps): there a sweep is a pure latency chain, and a lone warp walking a 256-cell strip issues
// ~95 ALU-pipe instructions per row at one per ~4 cycles (385 cycles per row measured, with or without its loads, stores and
// shuffles), while a CTA with four cells per thread, the previous row in shared memory and one barrier per row takes ~220.
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
// `split` (isotropic, small maps): the two sweeps of a slice are independent when both start from g — the result is then
// min(up, down), which the z pass takes while it stages — so they run as two CTAs (blockIdx.x = 2 slice + sweep) and the serial
// chain of a slice is Hb row steps instead of 2 Hb; costs one more map of traffic, so only where the maps live in the L2.
template <int XDIR>        // XDIR 0: isotropic (two-sided x, 3 neighbours, in-place second sweep); +-1: one-sided
__global__ void __launch_bounds__(256) ysweep_ring_kernel(const uint8_t *__restrict__ g, uint8_t *__restrict__ dst0, uint8_t *__restrict__ dst1,
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
// and the two-sided value is min(F, B) with B the mirror image.  No cell of [z+1, z+r] can have h < r (it would have made
// F(z + 1) smaller), so the window test asks one thing only: is the NEAREST cell behind the walker whose value is exactly r
// at most r steps away?  A walker therefore keeps, per value v, the step at which it last passed a cell with h == v (256 bytes of
// shared memory, step numbers mod 256) and a step is: one load of h(z), one load of last[r], one load of the cell that entry
// names to validate it (stale or never-written entries name some other cell: the test then fails on its value or its distance,
// and an entry that happens to name a cell with h == r within reach is a witness in its own right — so the table needs no
// initialisation), one store of last[h(z)], ~25 instructions in all, no search, no range-minimum tables, whatever the
// distances are.  (The table-based walk this replaces spent 5 warp instructions per cell — 2.65 G on the 1024x1024x512 map,
// 3.1 ms, profiles/r2v_k3_c5 — and 10 L bytes of shared memory per column; a variant tracking the window minimum with rescans
// was measured slower still: on a sparse map every step of a warp has some lane rescanning.)
// A CTA owns 32 adjacent columns, staged as 32-bit words by all its threads; warp 0 then walks them towards -z (F), warp 1 towards
// +z (B), a lane per column, each cell's result going straight to global memory (one 32-byte sector per warp and step).
// MODE 0: dst0 = min(F, B) (folded on the way, see walk_line) | 3: dst0 = F (towards +z), dst1 = B (towards -z).   Needs Wb % 4 == 0.
constexpr int kWalkTW = 32;

// SIDE > 0: F (sources at or after a cell: the walk starts at the line's last cell); SIDE < 0: B, the mirror image.
// `tcol` = the column's cell 0 in the staged tile (rows kWalkTW bytes apart), `last` = this walker's byte of the value table's
// row 0 (rows kWalkTW bytes apart), `res` = the column's cell 0 in the result map.
// FOLD: the two walkers of a column meet in the middle.  Up to there a walker is the first to reach its cells and stores its own
// value; past it (after a barrier of the two walker warps) the other walker's value is already in the result map, and what is
// stored is the minimum of the two — the two-sided result, without a second result map or a pass to combine them.  The other
// walker's values are requested a group of eight steps ahead (L2), off the walker's dependency chain.
template <int SIDE, bool FOLD>
__device__ __forceinline__ void walk_line(const uint8_t *__restrict__ tcol, uint8_t *__restrict__ last, uint8_t *res, size_t res_stride, int L, bool live)
{
	constexpr int   st = SIDE > 0 ? kWalkTW : -kWalkTW;        // one cell towards the walker's sources
	const int       z0 = SIDE > 0 ? L - 1 : 0;
	const uint8_t  *p  = tcol + z0 * kWalkTW;
	uint8_t        *o  = res + (size_t) z0 * res_stride;
	const ptrdiff_t os = SIDE > 0 ? -(ptrdiff_t) res_stride : (ptrdiff_t) res_stride;
	int             r  = *p;                                   // the line's end: nothing behind it
	last[r * kWalkTW] = 0;
	unsigned        e  = 0;                                    // last[r]: the step at which the nearest cell with h == r was passed
	// The dependency chain of a step is  e -> distance -> address -> validating load -> witness -> r.  The table entry of the NEXT
	// step is taken off it: r can only become min(h, r) or min(h, r + 1), both known before the witness is, so both entries are
	// requested up front and the witness selects one.
	auto step = [&](int k) {        // k: steps walked = cells behind the current one
		p -= st;
		const int      hz   = *p;
		const unsigned dist = ((unsigned) k - e) & 255u;             // how far back the table says the nearest cell with h == r lies
		const int      dc   = min((int) dist, k);                    // (never past the line's end: the cell there is as good a witness)
		const int      v    = p[dc * st];
		last[hz * kWalkTW]  = (uint8_t) k;
		const int      ra = min(hz, r), rb = min(hz, r + 1);         // (r == 255 always has a witness one step back: no overflow)
		const unsigned ea = last[ra * kWalkTW], eb = last[rb * kWalkTW];
		const bool     wit = ((int) dist <= r) & (v == r);
		r = wit ? ra : rb;
		e = wit ? ea : eb;
	};
	// cells z >= L / 2 are reached by F first, the others by B
	const int mine = FOLD ? (SIDE > 0 ? L - L / 2 : L / 2) : L;
	if (mine > 0 && live) *o = (uint8_t) r;
	int k = 1;
#pragma unroll 2
	for (; k < mine; ++k) {
		step(k);
		o += os;
		if (live) *o = (uint8_t) r;
	}
	if (FOLD) {
		asm volatile("bar.sync 1, 64;" ::: "memory");        // the two walker warps: everything stored so far is visible to the other
		if (mine == 0) {        // L == 1, the B walker: its only cell was F's
			if (live) *o = (uint8_t) min(r, (int) __ldcg(o));
			return;
		}
		// the other walker's values of the next eight cells, requested a group ahead
		constexpr int G = 8;
		auto request = [&](unsigned (&q)[G], int k0) {
#pragma unroll
			for (int j = 0; j < G; ++j) q[j] = (live && k0 + j < L) ? (unsigned) __ldcg(o + (ptrdiff_t) (k0 + j - (k - 1)) * os) : 255u;
		};
		unsigned cur[G], nxt[G];
		request(cur, k);
		while (k < L) {
			request(nxt, k + G);
#pragma unroll
			for (int j = 0; j < G; ++j) {
				if (k < L) {
					step(k);
					o += os;
					if (live) *o = (uint8_t) min((unsigned) r, cur[j]);
					++k;
				}
			}
#pragma unroll
			for (int j = 0; j < G; ++j) cur[j] = nxt[j];
		}
	}
}

// src2 (or null): a second input, the cell-wise minimum of the two is what gets transformed (the split y sweeps' outputs)
template <int MODE, int THREADS>        // THREADS: 64 (short lines: CTAs per SM are bound by their number) or 256 (long lines: by shared memory; more threads stage)
__global__ void __launch_bounds__(THREADS) zwalk_kernel(const uint8_t *__restrict__ src, const uint8_t *__restrict__ src2, uint8_t *dst0, uint8_t *dst1,
                                                             uint32_t Wb, uint32_t L, size_t line_stride, size_t outer_stride)
{
	extern __shared__ __align__(16) uint8_t T[];        // the staged tile (L x 32), then the two walkers' value tables (2 x 256 x 32)
	constexpr int  TW = kWalkTW, TW4 = TW >> 2;         // words per row
	const uint32_t x0   = blockIdx.x * TW;
	const size_t   base = (size_t) blockIdx.y * outer_stride + x0;
	const int      nwords = (int) L * TW4;
	uint8_t       *tables = T + (((size_t) L * TW + 15) & ~(size_t) 15);
	// stage: word w of row p = columns 4w .. 4w+3; four requests in flight per thread
	{
		const int  w  = threadIdx.x % TW4;
		const bool in = x0 + 4u * w < Wb;
		for (int p0 = threadIdx.x / TW4; p0 < (int) L; p0 += 4 * (THREADS / TW4)) {
			unsigned v[4];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const int p = p0 + j * (THREADS / TW4);
				v[j]        = 0xffffffffu;
				if (in && p < (int) L) v[j] = __ldg(reinterpret_cast<const unsigned *>(src + base + (size_t) p * line_stride) + w);
			}
			if (src2) {
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const int p = p0 + j * (THREADS / TW4);
					if (in && p < (int) L) v[j] = __vminu4(v[j], __ldg(reinterpret_cast<const unsigned *>(src2 + base + (size_t) p * line_stride) + w));
				}
			}
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const int p = p0 + j * (THREADS / TW4);
				if (p < (int) L) reinterpret_cast<unsigned *>(T)[p * TW4 + w] = v[j];
			}
		}
	}
	__syncthreads();
	const int  lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool live = x0 + lane < Wb;
	if (MODE == 0) {
		// both walkers store into dst0 and fold as they go
		if (warp == 0) walk_line<1, true>(T + lane, tables + lane, dst0 + base + lane, line_stride, (int) L, live);
		else if (warp == 1) walk_line<-1, true>(T + lane, tables + 256 * TW + lane, dst0 + base + lane, line_stride, (int) L, live);
	} else {
		if (warp == 0) walk_line<1, false>(T + lane, tables + lane, dst0 + base + lane, line_stride, (int) L, live);
		else if (warp == 1) walk_line<-1, false>(T + lane, tables + 256 * TW + lane, dst1 + base + lane, line_stride, (int) L, live);
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
			// cells per lane: as many as keep most of a warp busy on one row
#define VKV_XP(CPL, SEGS) xpass_lanes_kernel<DIR, CPL, SEGS><<<grid, 256, 0, s>>>(O, out, Wb, nrows)
			if (Wb % 16 == 0 && Wb >= 384) {
				const uint32_t nseg = (Wb + 511) / 512;
				if (nseg <= 1) VKV_XP(16, 1);
				else if (nseg <= 2) VKV_XP(16, 2);
				else VKV_XP(16, 4);
			} else if (Wb % 8 == 0 && Wb >= 160) {
				const uint32_t nseg = (Wb + 255) / 256;
				if (nseg <= 1) VKV_XP(8, 1);
				else if (nseg <= 2) VKV_XP(8, 2);
				else VKV_XP(8, 8);
			} else {
				const uint32_t nseg = (Wb + 127) / 128;
				if (nseg <= 2) VKV_XP(4, 2);
				else VKV_XP(4, 16);
			}
#undef VKV_XP
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
	// 32-bit words of four cells; up to 16 strips (512 threads) per slice.  (Four cells per lane — twice the warps, half the work per
	// warp and row — was measured on the 208x208x124 map, where a slice is one strip and most SMs hold a single warp: no gain,
	// 37.2 vs 38.9 us; the row step is a latency chain either way.)
	if (Wb % 4 != 0) return VKV_OK;        // otherwise the search kernel
	const uint32_t ctas    = split ? 2 * Db : Db;
	// narrow maps with few slices: the CTA-per-slice sweep (bulk copies need 16-byte aligned row blocks)
	const bool ring_ok = Wb <= 256u && ((size_t) Wb * Hb) % 16 == 0 && ((size_t) kSweepRows * Wb) % 16 == 0 && (reinterpret_cast<uintptr_t>(g) % 16) == 0 &&
	                     (reinterpret_cast<uintptr_t>(dst0) % 16) == 0 && (!dst1 || (reinterpret_cast<uintptr_t>(dst1) % 16) == 0);
	const char *force = getenv("VKV_YSWEEP");        // A/B: "ring" / "strip"
	// (measured, sweep kernel alone: isotropic 208x208x124 / 256x256x199 maps 22 / 28 us against 35 / 43 us with strips; the one-sided
	// sweeps of the octant maps the other way round, 39 / 44 against 27 / 34 us)
	const bool  ring  = ring_ok && (force ? force[0] == 'r' : (XDIR == 0 && ctas < (uint32_t) vol->ctx->sm_count * 4u));
	if (ring) {
		const int    threads = (int) ((Wb / 4 + 31u) / 32u * 32u);
		const size_t smem_r  = (size_t) kSweepSlots * kSweepRows * Wb + 2 * ((size_t) Wb / 4 + 2) * sizeof(uint2) + 16;
		static PerDeviceOnce configured;
		if (configured.first(vol->ctx->device)) {
			VKV_CUDA_CHECK(cudaFuncSetAttribute(ysweep_ring_kernel<XDIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
		}
		ysweep_ring_kernel<XDIR><<<ctas, threads, smem_r, s>>>(g, dst0, dst1, Wb, Hb, split ? 1 : 0);
		VKV_LAUNCHED();
		*done = true;
		return VKV_OK;
	}
	const int      nstrips = Wb <= 256u ? 1 : (int) ((Wb + (256 - 2 * kStripHalo) - 1) / (256 - 2 * kStripHalo));
	if (nstrips > 16) return VKV_OK;
	const size_t smem = 2 * ((size_t) Wb / 2 + 8) * sizeof(unsigned);
	ysweep_kernel<XDIR, 8><<<ctas, 32 * nstrips, smem, s>>>(g, dst0, dst1, Wb, Hb, split ? 1 : 0, nstrips);
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
	const size_t smem = (((size_t) L * kWalkTW + 15) & ~(size_t) 15) + 2 * 256 * kWalkTW;
	if (smem > (size_t) 200 * 1024) return VKV_OK;
	static PerDeviceOnce configured;
	if (configured.first(vol->ctx->device)) {
		VKV_CUDA_CHECK(cudaFuncSetAttribute(zwalk_kernel<MODE, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		VKV_CUDA_CHECK(cudaFuncSetAttribute(zwalk_kernel<MODE, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	}
	const dim3 grid((Wb + kWalkTW - 1) / kWalkTW, rows);
	if (smem * 8 > (size_t) 220 * 1024) zwalk_kernel<MODE, 256><<<grid, 256, smem, s>>>(src, src2, dst0, dst1, Wb, L, (size_t) Wb * Hb, (size_t) Wb);
	else zwalk_kernel<MODE, 64><<<grid, 64, smem, s>>>(src, src2, dst0, dst1, Wb, L, (size_t) Wb * Hb, (size_t) Wb);
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
	return Wb % 4 == 0 && Wb <= (uint32_t) (256 - 2 * kStripHalo) * 16u && Hb <= 65535u && Wb <= kRowWordsMax * 32u && (size_t) vol->dim_b[2] * kWalkTW + 2 * 256 * kWalkTW + 16 <= (size_t) 200 * 1024 &&
	       !getenv("VKV_DIST_SEARCH");
}

// `split`: the two y sweeps of a slice as two CTAs (half the serial chain — with a slab's few slices the sweep is nothing but that
// chain): sources above -> d_swap, sources below -> map 0 (its occupancy was consumed by the x pass); the z pass then takes the
// minimum of the two while it stages, and both slabs are exchanged.
int launch_distance_xy_slab(vkv_volume *vol, uint32_t zb_first, uint32_t zb_count, bool split, cudaStream_t s)
{
	int  rc;
	bool done = false;
	if (zb_count == 0) return VKV_OK;
	if ((rc = run_xpass<0>(vol, vol->d_maps[0], vol->d_tmp, s, zb_first, zb_count))) return rc;
	if ((rc = run_ysweep<0>(vol, vol->d_tmp, vol->d_swap, split ? vol->d_maps[0] : nullptr, s, &done, zb_first, zb_count, split))) return rc;
	if (!done) {
		set_error("launch_distance_xy_slab: shape not covered by the sweep kernel");
		return VKV_ERR_STATE;
	}
	return VKV_OK;
}

int launch_distance_z_rows(vkv_volume *vol, uint32_t yb_first, uint32_t yb_count, bool split, cudaStream_t s)
{
	int  rc;
	bool done = false;
	if ((rc = run_zwalk<0>(vol, vol->d_swap, vol->d_maps[0], nullptr, s, &done, yb_first, yb_count, split ? vol->d_maps[0] : nullptr))) return rc;
	if (!done) {
		set_error("launch_distance_z_rows: shape not covered by the walk kernel");
		return VKV_ERR_STATE;
	}
	return VKV_OK;
}

}        // namespace vkv
