// raycast.cu — K4v + K4: per-pixel ray caster with empty-space skipping and early ray termination.
//
// Replaces VolumeRenderSubpass::draw (src/volume_render_subpass.cpp:159-294): two indexed draws
// (clipped unit cube via shaders/volume_render_clipped.vert + HW user clip plane + back-face cull,
// box/plane polygon via shaders/volume_render_plane_intersection.vert), the rasteriser, the fragment
// shader shaders/volume_render.frag and the fixed-function blend + R8G8B8A8_SRGB store.
//
// B200 design:
//   * no geometry at all: each pixel computes its ray/box/clip-plane entry analytically in fp64
//     from an affine pixel->direction basis built on the host;
//   * one thread per pixel, a warp is an 8x4 pixel tile (coherent rays -> coherent texture and
//     distance-map fetches), a CTA is just two warps (16x4 pixels) so that a few long rays do not pin
//     a large slice of the register file;
//   * V and G are cudaArray 3D textures sampled with hardware trilinear filtering
//     (tex3D, normalised coordinates, clamp, UNORM->float); an EXACT variant does 8 point loads
//     from the linear copies and fp32 lerps in the oracle's operation order;
//   * the distance / occupancy maps are plain byte loads through L1 (texelFetch semantics);
//   * the transfer function is a 256x256 RGBA8 table read through the read-only path; the
//     opacity-correction pow() is hoisted into a 256-entry shared-memory table;
//   * blend-with-clear, sRGB encode and the RGBA8 store happen in the epilogue; the store
//     target may be a peer-mapped framebuffer (multi-GPU tile gather fused into the kernel).
// The march loop is the fragment shader's, statement for statement, in fp32 without contraction.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace vkv {

// Long rays (distance-map modes, production variants): a ray still marching after `long_T` loop iterations is suspended — its
// march state goes into a queue — and finished by raycast_long_kernel, one ray per WARP (see there).
struct LongRay {
	unsigned p;                 // pixel index py * width + px
	int      i, i_min, i_first_hit, n_steps;
	unsigned idx_last, occupied, pad;
	float    out[4];
	float    entry[3], step[3], pad2[2];        // written as five float4 by raycast_kernel, read the same way
};
static_assert(sizeof(LongRay) == 80, "LongRay layout");
struct LongQueue {
	unsigned count;             // slots reserved by raycast_kernel (may exceed capacity: the excess was not suspended)
	unsigned head;              // next slot to hand out
	unsigned done;              // CTAs of raycast_long_kernel that have exited (the last one resets the queue)
	unsigned pad;
};

struct RayParams {
	double o[3];                    // camera position, texture space
	double d0[3], ddx[3], ddy[3];   // far-plane direction of pixel (px,py): d0 + px*ddx + py*ddy
	double plane[4];                // clip plane, texture space
	double pvm_x[4], pvm_y[4];      // rows x and y of proj*view*model (the `position` varying under depth_attachment)
	double pvm_z[4], pvm_w[4];      // rows z and w of proj*view*model (depth output)
	double s0;                      // plane . (o, 1)
	float  fd0[3], fddx[3], fddy[3], fo[3], fplane[3], fs0;        // fp32 copies for the conservative rejection test
	float  vol_to_map[3];           // dim / block_size per axis
	float  cam_pos_tex[3];
	float  dimf[3];
	float  block_size[3];
	int    dim[3];
	int    dim_b[3];
	int    dim_max;
	float  sampling_factor, sampling_factor_inv, voxel_alpha_factor, grad_modifier;
	int    use_gradient;
	int    ert, test;
	int    depth_attachment;        // DEPTH_ATTACHMENT variant of the fragment shader (LOAD instantiations only)
	float  vpi[16], mi[16];         // view_proj_inv and model_inv in fp32, column-major (depth-buffer intersection)
	int    width, height;
	int    tile_w, tile_h, tiles_x, tile_first, tile_stride, my_tiles, seq_base;
	const unsigned *tile_order;     // issue order of this launch's tile list (from the previous frame's cost) or null: centre-out
	unsigned       *tile_cost;      // per tile of the list: largest warp loop count of this frame (atomicMax) or null
	unsigned tiles_x_magic;         // floor(2^32 / tiles_x) + 1: tile / tiles_x == umulhi(tile, magic) for every tile of a frame; 0: divide
	int    bbox[4];                 // conservative screen bounds (inclusive) of the unit cube: x0, y0, x1, y1
	cudaTextureObject_t tex_v, tex_g;
	const uint8_t *V, *G;
	const uchar4  *tf;
	const float4  *ctab;            // per TF texel: premultiplied colour + corrected opacity (w < 0: texel alpha byte is 0)
	const uint8_t *maps;            // map 0; map i at maps_stride * i (anisotropic)
	const uint8_t *map_ptrs[8];
	uint8_t       *rgba8;
	float         *depth;
	unsigned long long *counts;     // vkv_sample_counts or null
	unsigned long long *trace;      // debug (VKV_RC_TRACE): per warp {start ns, end ns, loop iterations} or null
	LongQueue *lq;                  // long-ray queue or null
	LongRay   *lrays;
	int        long_T;              // suspend after this many loop iterations (0: never)
	int        long_cap;            // records the queue holds
	int       *long_hint;           // mapped host int: rays the last frame handed over (raycast_long_kernel writes it)
	const TFBounds *bounds;         // conservative byte ranges of the visible TF texels (tf.cu): samples outside are empty without a table read
	int    flags;                   // debug (VKV_RC_FLAGS): experiment switches, 0 in production
};

__device__ __forceinline__ float clampf_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ int   clampi_(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ int   tf_texel(float c) { return clampi_((int) floorf(c * 256.0f), 0, 255); }
// clamp(x, 0, hi) for hi >= 0 in one VIMNMX.RELU
__device__ __forceinline__ int   clamp0_(int x, int hi) { return __vimin_s32_relu(x, hi); }
// k in [0,4) -> one of four registers without branches (the compiler otherwise builds a branch tree around the texture scoreboard)
__device__ __forceinline__ float pick4_(int k, float a, float b, float c, float d)
{
	float r;
	asm("{\n\t.reg .pred p0, p1, p2;\n\t.reg .f32 lo, hi;\n\t"
	    "setp.eq.s32 p0, %1, 0;\n\tsetp.eq.s32 p1, %1, 2;\n\tsetp.lt.s32 p2, %1, 2;\n\t"
	    "selp.f32 lo, %2, %3, p0;\n\tselp.f32 hi, %4, %5, p1;\n\tselp.f32 %0, lo, hi, p2;\n\t}"
	    : "=f"(r)
	    : "r"(k), "f"(a), "f"(b), "f"(c), "f"(d));
	return r;
}

// texture(sampler3D, pos), LINEAR / CLAMP_TO_EDGE / UNORM — fp32 restatement on the linear copy
__device__ __forceinline__ float sample_exact(const uint8_t *__restrict__ T, const int dim[3], float px, float py, float pz)
{
	const float u = px * (float) dim[0] - 0.5f, v = py * (float) dim[1] - 0.5f, w = pz * (float) dim[2] - 0.5f;
	const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
	const float a = u - fu, b = v - fv, c = w - fw;
	int         x0 = (int) fu, y0 = (int) fv, z0 = (int) fw;
	const int   mx = dim[0] - 1, my = dim[1] - 1, mz = dim[2] - 1;
	const int   x1 = clampi_(x0 + 1, 0, mx), y1 = clampi_(y0 + 1, 0, my), z1 = clampi_(z0 + 1, 0, mz);
	x0 = clampi_(x0, 0, mx); y0 = clampi_(y0, 0, my); z0 = clampi_(z0, 0, mz);
	const size_t W = dim[0], WH = (size_t) dim[0] * dim[1];
	auto ld = [&](int x, int y, int z) { return (float) __ldg(T + (size_t) z * WH + (size_t) y * W + x) / 255.0f; };
	const float t000 = ld(x0, y0, z0), t100 = ld(x1, y0, z0), t010 = ld(x0, y1, z0), t110 = ld(x1, y1, z0);
	const float t001 = ld(x0, y0, z1), t101 = ld(x1, y0, z1), t011 = ld(x0, y1, z1), t111 = ld(x1, y1, z1);
	const float c00 = t000 * (1.0f - a) + t100 * a, c10 = t010 * (1.0f - a) + t110 * a;
	const float c01 = t001 * (1.0f - a) + t101 * a, c11 = t011 * (1.0f - a) + t111 * a;
	const float c0 = c00 * (1.0f - b) + c10 * b, c1 = c01 * (1.0f - b) + c11 * b;
	return c0 * (1.0f - c) + c1 * c;
}

// get_gradient of the fragment shader without PRECOMPUTED_GRADIENT (volume_render.frag:91-97): four filtered taps of the
// VOLUME at pos +- 1/dim on a tetrahedron, combined in the shader's order (no contraction)
template <bool EXACT>
__device__ __forceinline__ float gradient_otf(cudaTextureObject_t tex, const uint8_t *__restrict__ V, const int dim[3], const float di[3],
                                              float modifier, float px, float py, float pz)
{
	const float xp = px + di[0] * 1.0f, xm = px + di[0] * -1.0f, yp = py + di[1] * 1.0f, ym = py + di[1] * -1.0f;
	const float zp = pz + di[2] * 1.0f, zm = pz + di[2] * -1.0f;
	float a, b, d, e;
	if (EXACT) {
		a = sample_exact(V, dim, xp, ym, zm); b = sample_exact(V, dim, xm, ym, zp);
		d = sample_exact(V, dim, xm, yp, zm); e = sample_exact(V, dim, xp, yp, zp);
	} else {
		a = tex3D<float>(tex, xp, ym, zm); b = tex3D<float>(tex, xm, ym, zp);
		d = tex3D<float>(tex, xm, yp, zm); e = tex3D<float>(tex, xp, yp, zp);
	}
	const float gx = (((1.0f * a + -1.0f * b) + -1.0f * d) + 1.0f * e) * 0.25f;
	const float gy = (((-1.0f * a + -1.0f * b) + 1.0f * d) + 1.0f * e) * 0.25f;
	const float gz = (((-1.0f * a + 1.0f * b) + -1.0f * d) + 1.0f * e) * 0.25f;
	return clampf_(sqrtf((gx * gx + gy * gy) + gz * gz) * modifier, 0.0f, 1.0f);
}

__device__ __forceinline__ float srgb_encode(float c) { return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f; }
// The hardware-filter production variants (kFast in rc_cast_pixel) take the liberties every Vulkan driver takes with the reference's
// GLSL: multiply-adds contracted (GLSL allows it without `precise`), divisions and pow at the 2-3 ulp the Vulkan spec asks of
// them.  The EXACT-filter variants keep the oracle's operation-for-operation arithmetic (-fmad=false, IEEE division).
template <bool FAST> __device__ __forceinline__ float madf_(float a, float b, float c) { return FAST ? __fmaf_rn(a, b, c) : a * b + c; }
template <bool FAST> __device__ __forceinline__ float divf_(float a, float b) { return FAST ? __fdividef(a, b) : a / b; }
template <bool FAST> __device__ __forceinline__ float srgb_encode_(float c)
{
	if (!FAST) return srgb_encode(c);
	return c <= 0.0031308f ? 12.92f * c : __fmaf_rn(1.055f, __powf(c, 1.0f / 2.4f), -0.055f);
}
// R8G8B8A8_SRGB load (sRGB EOTF of the Vulkan specification): what the blender reads back from the attachment
__device__ __forceinline__ float srgb_decode(unsigned b)
{
	const float c = (float) b / 255.0f;
	return c <= 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
}
__device__ __forceinline__ unsigned unorm8(float c) { return (unsigned) (clampf_(c, 0.0f, 1.0f) * 255.0f + 0.5f); }


// Blend over the render-pass clear (0,0,0,1) / depth 0 and store: rgb = src.rgb, a = src.a * (1 - src.a)
// (volume_render_subpass.cpp:176-190), R8G8B8A8_SRGB encode; pixels that fail the depth test keep the clear values.
template <bool FAST>
__device__ __forceinline__ unsigned pack_over_clear(bool pass, const float out[4])
{
	unsigned packed = 0xff000000u;
	if (pass) {
		// the built-in transfer function is grey (volume_component.cpp:250-261): one encode serves the three channels
		const unsigned er = unorm8(srgb_encode_<FAST>(clampf_(out[0], 0.0f, 1.0f)));
		const bool     grey = out[1] == out[0] && out[2] == out[0];
		const unsigned eg = grey ? er : unorm8(srgb_encode_<FAST>(clampf_(out[1], 0.0f, 1.0f)));
		const unsigned eb = grey ? er : unorm8(srgb_encode_<FAST>(clampf_(out[2], 0.0f, 1.0f)));
		packed = er | (eg << 8) | (eb << 16) | (unorm8(out[3] * (1.0f - out[3])) << 24);
	}
	return packed;
}

// Per TF texel: the fragment shader's `color` after opacity correction and premultiplication
// (volume_render.frag:279-285): ca = clamp(voxel_alpha_factor * (1 - pow(1 - a, 1/sampling_factor)), 0, 1),
// rgb = (byte / 255) * ca.  w = -1 marks texels whose alpha byte is 0 (voxel_occupied = false).
// Built once per (TF texture, sampling_factor, voxel_alpha_factor) and cached in the volume: one 16-byte
// read-only load replaces the RGBA8 load, four UNORM decodes, the pow() and three multiplies per sample.
__global__ void __launch_bounds__(256) ctab_kernel(const uchar4 *__restrict__ tf, float4 *__restrict__ ctab, float voxel_alpha_factor,
                                                   float sampling_factor_inv)
{
	const int    idx = blockIdx.x * blockDim.x + threadIdx.x;
	const uchar4 tx  = tf[idx];
	float4       e   = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
	if (tx.w > 0) {
		const float a  = (float) tx.w / 255.0f;
		const float ca = clampf_(voxel_alpha_factor * (1.0f - powf(1.0f - a, sampling_factor_inv)), 0.0f, 1.0f);
		e = make_float4(((float) tx.x / 255.0f) * ca, ((float) tx.y / 255.0f) * ca, ((float) tx.z / 255.0f) * ca, ca);
	}
	ctab[idx] = e;
}

// Tiles of a launch are issued from the middle of its tile list outwards (m, m-1, m+1, m-2, ...) unless a history says otherwise.
__device__ __forceinline__ int centre_out(int seq, int n) { return n / 2 + ((seq & 1) ? -((seq + 1) >> 1) : (seq >> 1)); }
constexpr int kTraceWords = 8;
// A CTA is kRcWarps warps, two across and kRcWarps / 2 down: 16 x (2 kRcWarps) pixels.
#ifndef VKV_RC_WARPS
#define VKV_RC_WARPS 2
#endif
constexpr int kRcWarps = VKV_RC_WARPS, kRcThreads = 32 * kRcWarps, kRcRows = 2 * kRcWarps;
#ifndef VKV_RC_MIN_CTAS
#define VKV_RC_MIN_CTAS 16
#endif
// ---- long rays: one ray per warp, 64 lattice steps at a time -------------------------------------------------------------
// A ray's march is a serial state machine (skip-map hop -> next position -> hop ...), ~1300 cycles per trip when 20 rays share a
// warp in lockstep: a frame's time used to be the ~180 trips of its longest rays (profiles/r1s_trace.md).  What a step of the
// machine READS, however, depends on the step index alone (pos = entry + i * step): so a warp takes one suspended ray and
// evaluates 32 consecutive lattice steps at once — lane l looks at step base + l: block index, skip-map byte, the hop that byte
// would cause, the filtered sample(s) and its colour-table entry, 32 independent fetches instead of a 32-deep dependent chain —
// parks them in shared memory and then REPLAYS the shader's state machine over the window with no memory access at all
// (a few ALU instructions per visited step, every lane computing the same state).  The replay visits exactly the steps
// the serial march visits, in the same order, with the same fp32 operations: frames and counters are bit-identical to the
// one-kernel march (tests/test_parity_gpu.py::test_long_ray_pass_is_bit_identical).
__device__ __forceinline__ void store_over_clear(const RayParams &P, size_t p, bool pass, const float out[4], float frag_depth)
{
	reinterpret_cast<unsigned *>(P.rgba8)[p] = pack_over_clear<true>(pass, out);        // long rays exist in the kFast variants only
	if (P.depth) P.depth[p] = pass ? frag_depth : 0.0f;
}
// 1 / (map cells the ray advances per trip along axis k): sdi of the shader (volume_render.frag:229)
template <bool FAST> __device__ __forceinline__ float sdt_inv_(const RayParams &P, float step_k, int k)
{
	if (FAST) return __fdividef(1.0f, step_k * P.vol_to_map[k]);
	const float sdt = step_k * P.dimf[k] / P.block_size[k];
	return 1.0f / sdt;
}
#ifndef VKV_RC_LONG_PER
#define VKV_RC_LONG_PER 2
#endif
constexpr int kLongPer = VKV_RC_LONG_PER, kLongWindow = 32 * kLongPer;        // lattice steps per lane and per window
struct LongConsts {
	int      back, dim_b1[3];
	TFRange  tb;
	unsigned tb_vspan, tb_gspan;
};
__device__ __forceinline__ LongConsts long_consts(const RayParams &P)
{
	LongConsts L;
	L.back      = (int) ceilf(P.sampling_factor);
	L.dim_b1[0] = P.dim_b[0] - 1; L.dim_b1[1] = P.dim_b[1] - 1; L.dim_b1[2] = P.dim_b[2] - 1;
	L.tb        = P.use_gradient ? P.bounds->tex_all : P.bounds->tex_row255;
	L.tb_vspan  = L.tb.v_hi - L.tb.v_lo;
	L.tb_gspan  = L.tb.g_hi - L.tb.g_lo;
	return L;
}

// Finishes suspended ray j with the whole warp.  s_c / s_idx / s_hop: this warp's kLongWindow-entry shared-memory window.
template <int SKIP>
__device__ __forceinline__ void rc_long_ray(const RayParams &P, const LongConsts &LC, const unsigned j, const int lane, float4 *s_c, unsigned *s_idx, int *s_hop,
                                            unsigned long long &c_vol, unsigned long long &c_dist, unsigned long long &c_empty)
{
	const int      back = LC.back;
	const int      dim_b1[3] = {LC.dim_b1[0], LC.dim_b1[1], LC.dim_b1[2]};
	const TFRange  tb = LC.tb;
	const unsigned tb_vspan = LC.tb_vspan, tb_gspan = LC.tb_gspan;
	const float4 *r = reinterpret_cast<const float4 *>(P.lrays + j);
	const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4];
	const unsigned p = __float_as_uint(r0.x);
	int            i = __float_as_int(r0.y), i_min = __float_as_int(r0.z), i_first_hit = __float_as_int(r0.w);
	const int      n_steps = __float_as_int(r1.x);
	unsigned       idx_last = __float_as_uint(r1.y);
	bool           voxel_occupied = __float_as_uint(r1.z) != 0u;
	float          out[4]   = {r2.x, r2.y, r2.z, r2.w};
	const float    entry[3] = {r3.x, r3.y, r3.z}, step[3] = {r3.w, r4.x, r4.y};
	float          sdt_inv[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) sdt_inv[k] = sdt_inv_<true>(P, step[k], k);        // the same arithmetic as the march that suspended the ray
	const bool neg[3] = {sdt_inv[0] < 0.0f, sdt_inv[1] < 0.0f, sdt_inv[2] < 0.0f};
	const uint8_t *__restrict__ Dm = P.map_ptrs[0];
	if (SKIP == VKV_SKIP_ANISOTROPIC_DISTANCE) Dm = P.map_ptrs[(step[2] < 0 ? 1 : 0) + (step[1] < 0 ? 2 : 0) + (step[0] < 0 ? 4 : 0)];        // sign(step) == sign(dir)
	unsigned n_vol = 0u, n_dist = 0u, n_empty = 0u;
	bool     done = false;
	while (i < n_steps && !done) {
		// the window starts where a step back (volume_render.frag:253-261) from the current step could land
		const int base = max(i - back, i_min);
		unsigned  my_idx[kLongPer];
		unsigned  vis_w[kLongPer];        // ballots: bit l of word h speaks for step base + 32 h + l
#pragma unroll
		for (int h = 0; h < kLongPer; ++h) {
			const int   slot   = 32 * h + lane;
			const float fi     = (float) (base + slot);
			const float pos[3] = {__fmaf_rn(fi, step[0], entry[0]), __fmaf_rn(fi, step[1], entry[1]), __fmaf_rn(fi, step[2], entry[2])};
			float       u[3];
			int         u_i[3];
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				u[k]   = P.vol_to_map[k] * pos[k];
				u_i[k] = clamp0_((int) u[k], dim_b1[k]);
			}
			const unsigned idx  = ((unsigned) u_i[2] * (unsigned) P.dim_b[1] + (unsigned) u_i[1]) * (unsigned) P.dim_b[0] + (unsigned) u_i[0];
			const unsigned dist = __ldg(Dm + idx);
			const float    intensity = tex3D<float>(P.tex_v, pos[0], pos[1], pos[2]);
			const float    gradient  = P.use_gradient ? tex3D<float>(P.tex_g, pos[0], pos[1], pos[2]) : 1.0f;
			int            hop       = 0;
			if (dist > 0u) {
				float       dxyz[3];
				const float fd = (float) dist, omfd = 1.0f - fd;
#pragma unroll
				for (int k = 0; k < 3; ++k) {
					const float rr  = clampf_((float) u_i[k] - u[k], -1.0f, 0.0f);
					const float bse = neg[k] ? omfd : fd;
					dxyz[k]         = (bse + rr) * sdt_inv[k];
				}
				hop = max((int) ceilf(fminf(fminf(dxyz[0], dxyz[1]), dxyz[2])), 1);
			}
			const int ti = tf_texel(intensity), tg = tf_texel(gradient);
			float4    c  = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
			if ((unsigned) ti - tb.v_lo <= tb_vspan && (unsigned) tg - tb.g_lo <= tb_gspan) c = __ldg(P.ctab + tg * 256 + ti);
			s_idx[slot] = idx;
			s_hop[slot] = hop;
			s_c[slot]   = c;
			my_idx[h]   = idx;
			vis_w[h]    = __ballot_sync(0xffffffffu, c.w >= 0.0f);
		}
		unsigned same_w[kLongPer];
		auto same_mask = [&](unsigned ref) {
#pragma unroll
			for (int h = 0; h < kLongPer; ++h) same_w[h] = __ballot_sync(0xffffffffu, my_idx[h] == ref);
		};
		// word h of a per-lane array without dynamic indexing (kLongPer is 2 or 4)
		auto word = [&](const unsigned (&m)[kLongPer], int h) {
			unsigned v = m[0];
#pragma unroll
			for (int q = 1; q < kLongPer; ++q) v = h == q ? m[q] : v;
			return v;
		};
		same_mask(idx_last);
		__syncwarp();
		// Replay of the fragment shader's loop body (volume_render.frag:215-312) over the window; every lane carries the same state.
		// One trip of this loop is one EVENT of the march: a skip-map consultation, a visible sample, or a whole run of empty
		// samples (the common case on a ray that grazes the surface: nothing but the counters changes along it).
		while (i < n_steps) {
			const int k = i - base;
			if (k >= kLongWindow) break;
			const int      kw = k >> 5, kb = k & 31;
			const unsigned sw = word(same_w, kw), vw = word(vis_w, kw);
			if (!voxel_occupied && !((sw >> kb) & 1u)) {
				++n_dist;
				const int hop = s_hop[k];
				if (hop > 0) {
					i += hop;
				} else {
					voxel_occupied = true;
					idx_last       = s_idx[k];
					same_mask(idx_last);
					i = max(i - back, i_min);
				}
			} else if ((vw >> kb) & 1u) {
				++n_vol;
				const float4   c   = s_c[k];
				const unsigned idx = s_idx[k];
				voxel_occupied = true;
				if (idx != idx_last) {
					idx_last = idx;
					same_mask(idx_last);
				}
				const float w = 1.0f - out[3];
				out[0] = __fmaf_rn(w, c.x, out[0]); out[1] = __fmaf_rn(w, c.y, out[1]); out[2] = __fmaf_rn(w, c.z, out[2]); out[3] = __fmaf_rn(w, c.w, out[3]);
				if (c.w > 0.0f) i_first_hit = i;
				if (out[3] > 0.99f && P.ert) {
					out[3] = 1.0f;
					done   = true;
					break;
				}
				++i;
				i_min = i;
			} else {
				// step k is sampled and empty; the steps after it are sampled too for as long as they stay in block idx_last
				// (voxel_occupied is false from here on) and are empty themselves: count them word by word
				int L = 0;
				{
					const unsigned run = ~vw & (sw | (1u << kb));
					const unsigned t   = ~(run >> kb);        // the zeros shifted in at the top end the count at the word's end
					L                  = t ? __ffs((int) t) - 1 : 32;
				}
				for (int h = kw + 1; L == 32 * (h - kw) - kb && h < kLongPer; ++h) {
					const unsigned t = ~(~word(vis_w, h) & word(same_w, h));
					L += t ? __ffs((int) t) - 1 : 32;
				}
				L = min(L, n_steps - i);
				n_vol += (unsigned) L;
				n_empty += (unsigned) L;
				i += L;
				i_min          = i;
				voxel_occupied = false;
			}
		}
		__syncwarp();
	}
	c_vol += n_vol; c_dist += n_dist; c_empty += n_empty;
	if (lane == 0) {
		float frag_depth = 0.0f;
		if (P.depth && out[3] > 0.0f && i_first_hit < n_steps) {
			const double pm[3] = {(double) (entry[0] + step[0] * (float) i_first_hit) - 0.5,
			                      (double) (entry[1] + step[1] * (float) i_first_hit) - 0.5,
			                      (double) (entry[2] + step[2] * (float) i_first_hit) - 0.5};
			const double z = P.pvm_z[0] * pm[0] + P.pvm_z[1] * pm[1] + P.pvm_z[2] * pm[2] + P.pvm_z[3];
			const double w = P.pvm_w[0] * pm[0] + P.pvm_w[1] * pm[1] + P.pvm_w[2] * pm[2] + P.pvm_w[3];
			frag_depth     = (float) (z / w);
		}
		store_over_clear(P, (size_t) p, frag_depth >= 0.0f, out, frag_depth);
	}
}

// A CTA is two warps = a 16x4 pixel tile; small CTAs keep the register file busy while long rays finish.
// grid = (CTAs per tile in x, CTAs per tile in y, tiles of this launch).
// OTF: the volume has no gradient map; gradients come from gradient_otf (always instantiated with COUNT).
// TRACE: debug instantiation (VKV_RC_TRACE) that also records the per-warp timeline.
// LOAD: blend and depth-test over the existing contents of the target (vkv_render_options::load_framebuffer) instead of the
// render-pass clear, and honour depth_attachment; these instantiations always count (COUNT).
struct RcTraceAcc {        // TRACE instantiation only
	unsigned  tr_lanes = 0, tr_d = 0, tr_r = 0, tr_mixed = 0;
	long long tc_top = 0, tc_req = 0, tc_d = 0, tc_v = 0, tc_mark = 0;        // cycles per loop section
};

// One pixel: analytic ray entry, the fragment shader's march, blend + store.  Counters accumulate into the caller's (a persistent
// warp casts many pixel blocks); n_iter_out is the number of loop trips this lane made (tile cost, trace).
// OTF: the volume has no gradient map; gradients come from gradient_otf (always instantiated with COUNT).
// TRACE: debug instantiation (VKV_RC_TRACE) that also records the per-warp timeline.
// LOAD: blend and depth-test over the existing contents of the target (vkv_render_options::load_framebuffer) instead of the
// render-pass clear, and honour depth_attachment; these instantiations always count (COUNT).
template <int SKIP, bool EXACT, bool COUNT, bool OTF, bool TRACE, bool LOAD, bool GRAD>
__device__ __forceinline__ void rc_cast_pixel(const RayParams &P, const int px, const int py, const int lane, unsigned &n_vol, unsigned &n_dist,
                                              unsigned &n_empty, unsigned &covered_acc, unsigned &n_iter_out, RcTraceAcc &tr)
{
	// (compiled into the distance-map production variants only: the other modes are throughput-bound and keep their old code)
	constexpr bool kHist = (SKIP == VKV_SKIP_DISTANCE || SKIP == VKV_SKIP_ANISOTROPIC_DISTANCE) && !OTF && !EXACT && !LOAD;
	// the same variants hand their long rays over (P.long_T loop iterations into the march; 0: never)
	constexpr bool kLong = kHist;
	// hardware-filter production variants: contracted multiply-adds, 2-ulp divisions, fp32 entry point away from the silhouette
	constexpr bool kFast = !EXACT && !OTF && !LOAD;
	// GRAD = false: instantiations for transfer functions that ignore the gradient (the headline configuration): the gradient fetches,
	// their batch registers and the second texel index are not compiled in at all (a predicated-off instruction still costs an issue slot)
	const bool   use_g = GRAD && P.use_gradient;
	const bool   in_frame = px < P.width && py < P.height;
	const size_t p        = (size_t) py * P.width + px;
	unsigned covered = 0u;        // this pixel
	unsigned n_iter  = 0u;
	bool     suspended = false;
	float    out[4] = {0.0f, 0.0f, 0.0f, 0.0f};
	float    frag_depth = 0.0f;
	float    dst_depth  = 0.0f;        // what the depth attachment holds (LOAD) or the clear value
	bool     discarded  = false;

	if (in_frame) {
		if (LOAD && P.depth) dst_depth = P.depth[p];
		// ---- analytic ray entry (replaces both vertex shaders + rasteriser) ----
		// (1) conservative fp32 rejection (approximate divisions; only ever used to say "certainly misses")
		bool  maybe = px >= P.bbox[0] && px <= P.bbox[2] && py >= P.bbox[1] && py <= P.bbox[3];
		float entry_fast[3] = {0.0f, 0.0f, 0.0f};
		if (maybe) {
			float tnf = -INFINITY, tff = INFINITY, sdf = 0.0f;
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				const float dk = P.fd0[k] + (float) px * P.fddx[k] + (float) py * P.fddy[k];
				const float rk = __fdividef(1.0f, dk);
				const float a0 = (0.0f - P.fo[k]) * rk, a1 = (1.0f - P.fo[k]) * rk;
				tnf = fmaxf(tnf, fminf(a0, a1));
				tff = fminf(tff, fmaxf(a0, a1));
				sdf += P.fplane[k] * dk;
			}
			const float t0f = fmaxf(fmaxf(tnf, __fdividef(-P.fs0, sdf)), 0.0f);
			// reject only when the miss is far outside fp32 rounding (relative 1e-3); NaNs fall through to fp64
			const float band = 1e-3f * (fabsf(tff) + fabsf(t0f)) + 1e-6f;
			if (t0f > tff + band) maybe = false;
			// kFast: a hit that is as far from the silhouette on the other side takes its entry point from the same fp32 numbers
			// (error ~1e-7 of the box: 1e-4 voxel); only the band in between (a pixel or two wide) pays for fp64
			else if (kFast && t0f < tff - band && sdf > 1e-12f) {
				covered = 1;
				maybe   = false;
#pragma unroll
				for (int k = 0; k < 3; ++k) entry_fast[k] = __fmaf_rn(t0f, P.fd0[k] + (float) px * P.fddx[k] + (float) py * P.fddy[k], P.fo[k]);
			}
		}
		float entry[3] = {entry_fast[0], entry_fast[1], entry_fast[2]};
		if (maybe) {
			// (2) exact decision and entry point in fp64
			double d[3], tn = -INFINITY, tf = INFINITY;
			bool   hit = true;
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				d[k] = P.d0[k] + (double) px * P.ddx[k] + (double) py * P.ddy[k];
				if (d[k] == 0.0) {
					if (P.o[k] < 0.0 || P.o[k] > 1.0) hit = false;
				} else {
					double t0, t1;
					if (kFast) {
						const double r = 1.0 / d[k];
						t0 = (0.0 - P.o[k]) * r; t1 = (1.0 - P.o[k]) * r;
					} else {
						t0 = (0.0 - P.o[k]) / d[k]; t1 = (1.0 - P.o[k]) / d[k];
					}
					if (t0 > t1) { const double t = t0; t0 = t1; t1 = t; }
					if (t0 > tn) tn = t0;
					if (t1 < tf) tf = t1;
				}
			}
			const double sd = P.plane[0] * d[0] + P.plane[1] * d[1] + P.plane[2] * d[2];
			if (hit && sd > 0.0) {
				const double t_clip = -P.s0 / sd;
				double       t0     = tn > t_clip ? tn : t_clip;
				if (t0 < 0.0) t0 = 0.0;
				if (t0 < tf) {
					covered = 1;
#pragma unroll
					for (int k = 0; k < 3; ++k) entry[k] = (float) (P.o[k] + t0 * d[k]);
				}
			}
		}

		float position[4] = {0.0f, 0.0f, 0.5f, 1.0f}, frag_depth_front = 0.0f;
		if (LOAD && covered && P.depth_attachment) {
			// the `position` varying = gl_Position of the entry point (volume_render_clipped.vert:58-62), then :122-136
			const double pm[3] = {(double) entry[0] - 0.5, (double) entry[1] - 0.5, (double) entry[2] - 0.5};
			position[0] = (float) (P.pvm_x[0] * pm[0] + P.pvm_x[1] * pm[1] + P.pvm_x[2] * pm[2] + P.pvm_x[3]);
			position[1] = (float) (P.pvm_y[0] * pm[0] + P.pvm_y[1] * pm[1] + P.pvm_y[2] * pm[2] + P.pvm_y[3]);
			position[2] = (float) (P.pvm_z[0] * pm[0] + P.pvm_z[1] * pm[1] + P.pvm_z[2] * pm[2] + P.pvm_z[3]);
			position[3] = (float) (P.pvm_w[0] * pm[0] + P.pvm_w[1] * pm[1] + P.pvm_w[2] * pm[2] + P.pvm_w[3]);
			frag_depth_front = position[2] / position[3];
			if (dst_depth > frag_depth_front) discarded = true;        // REVERSE_DEPTH: the front face is behind the scene
			else frag_depth = dst_depth;
		}

		if (covered && !discarded) {
			// ---- fragment shader main() (volume_render.frag:117-336) ----
			const float dv[3] = {entry[0] - P.cam_pos_tex[0], entry[1] - P.cam_pos_tex[1], entry[2] - P.cam_pos_tex[2]};
			const float dl    = sqrtf((dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2]);
			const float dli    = kFast ? __fdividef(1.0f, dl) : 0.0f;
			const float dir[3] = {kFast ? dv[0] * dli : dv[0] / dl, kFast ? dv[1] * dli : dv[1] / dl, kFast ? dv[2] * dli : dv[2] / dl};
			float t2[3];
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				const float inv   = divf_<kFast>(1.0f, dir[k]);
				const float t_min = -entry[k] * inv, t_max = (1.0f - entry[k]) * inv;
				t2[k]             = fmaxf(t_min, t_max);
			}
			const float t_far       = fminf(fminf(t2[0], t2[1]), t2[2]);
			float       ray_exit[3] = {t_far * dir[0] + entry[0], t_far * dir[1] + entry[1], t_far * dir[2] + entry[2]};
			const float ev[3]       = {entry[0] - ray_exit[0], entry[1] - ray_exit[1], entry[2] - ray_exit[2]};
			float       ray_distance = sqrtf((ev[0] * ev[0] + ev[1] * ev[1]) + ev[2] * ev[2]);
			if (LOAD && P.depth_attachment) {
				// :151-165 — stop the ray where it meets the depth buffer
				const float cad[4] = {position[0] * dst_depth / frag_depth_front, position[1] * dst_depth / frag_depth_front,
				                      position[2] * dst_depth / frag_depth_front, position[3]};
				float pad[4], pmv[3];
#pragma unroll
				for (int r = 0; r < 4; ++r) pad[r] = ((P.vpi[0 + r] * cad[0] + P.vpi[4 + r] * cad[1]) + P.vpi[8 + r] * cad[2]) + P.vpi[12 + r] * cad[3];
				const float pw = pad[3];
#pragma unroll
				for (int r = 0; r < 4; ++r) pad[r] = pad[r] / pw;
#pragma unroll
				for (int r = 0; r < 3; ++r) pmv[r] = ((P.mi[0 + r] * pad[0] + P.mi[4 + r] * pad[1]) + P.mi[8 + r] * pad[2]) + P.mi[12 + r] * pad[3];
				const float hit[3] = {pmv[0] + 0.5f, pmv[1] + 0.5f, pmv[2] + 0.5f};
				const float hv[3]  = {entry[0] - hit[0], entry[1] - hit[1], entry[2] - hit[2]};
				const float hd     = sqrtf((hv[0] * hv[0] + hv[1] * hv[1]) + hv[2] * hv[2]);
				if (hd < ray_distance) {
					ray_exit[0] = hit[0]; ray_exit[1] = hit[1]; ray_exit[2] = hit[2];
					ray_distance = hd;
				}
			}

			if (P.test == VKV_TEST_RAY_ENTRY) {
				out[0] = entry[0]; out[1] = entry[1]; out[2] = entry[2]; out[3] = 1.0f;
			} else if (P.test == VKV_TEST_RAY_EXIT) {
				out[0] = ray_exit[0]; out[1] = ray_exit[1]; out[2] = ray_exit[2]; out[3] = 1.0f;
			} else {
				const int n_steps = (int) ceilf((float) P.dim_max * ray_distance * P.sampling_factor);
				float     step[3];
#pragma unroll
				for (int k = 0; k < 3; ++k) step[k] = divf_<kFast>(dir[k] * ray_distance, (float) n_steps - 1.0f);
				bool inside = true;
#pragma unroll
				for (int k = 0; k < 3; ++k) {
					const float t = entry[k] + step[k];
					if (!(t > 0.0f && t < 1.0f)) inside = false;        // lessThanEqual 0 / greaterThanEqual 1 / NaN
				}
				if (inside) {
					float sdt_inv[3];
#pragma unroll
					for (int k = 0; k < 3; ++k) sdt_inv[k] = sdt_inv_<kFast>(P, step[k], k);
					// (st + sg * dist) of the shader is an exact small integer: dist when the ray advances along +k, 1 - dist along -k
					// (BLOCK_SKIP: 1 or 0).  sdt_inv is finite, non-zero and not NaN here (the `inside` test above saw to that).
					const bool neg[3] = {sdt_inv[0] < 0.0f, sdt_inv[1] < 0.0f, sdt_inv[2] < 0.0f};
					const uint8_t *__restrict__ Dm = P.map_ptrs[0];
					if (SKIP == VKV_SKIP_ANISOTROPIC_DISTANCE)
						Dm = P.map_ptrs[(dir[2] < 0 ? 1 : 0) + (dir[1] < 0 ? 2 : 0) + (dir[0] < 0 ? 4 : 0)];
					// u_last of the shader, kept as the linear index of the block (the map has < 2^32 blocks, so the
					// index identifies the block); it starts at block (0,0,0) (volume_render.frag:197)
					unsigned idx_last = 0u;
					int      i_min    = 0;
					bool     voxel_occupied = true;
					int      i_first_hit    = n_steps;
					const int back = (int) ceilf(P.sampling_factor);
					const int dim_b1[3] = {P.dim_b[0] - 1, P.dim_b[1] - 1, P.dim_b[2] - 1};
					const float dim_inv[3] = {1.0f / P.dimf[0], 1.0f / P.dimf[1], 1.0f / P.dimf[2]};
					// look-ahead cache of hardware-filtered samples i .. i+3: consecutive volume samples are the common case
					// inside occupied regions, and one batch of independent fetches replaces four dependent round trips
					// conservative visible rectangle of the TF texture in (intensity texel, gradient texel): a sample outside it is
					// empty (alpha byte 0) and needs no colour-table read — on the long grazing rays most samples are
					const TFRange tb    = use_g ? P.bounds->tex_all : P.bounds->tex_row255;
					const unsigned tb_vspan = tb.v_hi - tb.v_lo, tb_gspan = tb.g_hi - tb.g_lo;
					int      pre_base = -0x40000000;
					float pre_v0 = 0.0f, pre_v1 = 0.0f, pre_v2 = 0.0f, pre_v3 = 0.0f, pre_g0 = 1.0f, pre_g1 = 1.0f, pre_g2 = 1.0f, pre_g3 = 1.0f;
					// trips this lane may spend in the march before it is handed to raycast_long_kernel (kLong variants; never otherwise)
					unsigned trip_limit = (kLong && P.long_T > 0) ? (unsigned) P.long_T : 0xffffffffu;
					bool     ert_done   = false;
					int      i          = 0;
					for (;;) {
					for (; i < n_steps && (!kLong || n_iter < trip_limit);) {
						if (kHist || TRACE) ++n_iter;
						if (TRACE) tr.tc_mark = clock64();
						const float fi     = (float) i;
						const float pos[3] = {madf_<kFast>(fi, step[0], entry[0]), madf_<kFast>(fi, step[1], entry[1]), madf_<kFast>(fi, step[2], entry[2])};
						float    u[3];
						int      u_i[3] = {0, 0, 0};
						unsigned idx    = 0u;
						bool     do_skip = false;
						if (SKIP != VKV_SKIP_NONE) {
#pragma unroll
							for (int k = 0; k < 3; ++k) {
								u[k]   = P.vol_to_map[k] * pos[k];
								u_i[k] = clamp0_((int) u[k], dim_b1[k]);
							}
							idx     = ((unsigned) u_i[2] * (unsigned) P.dim_b[1] + (unsigned) u_i[1]) * (unsigned) P.dim_b[0] + (unsigned) u_i[0];
							do_skip = !voxel_occupied && idx != idx_last;
						}
						// ---- memory requests of this iteration, issued before any result is consumed ----
						// Every sampling lane keeps a batch of four hardware-filtered samples (steps pre_base .. pre_base + 3).  The
						// batches of a warp are kept aligned: when one lane runs out, every sampling lane that has used part of its
						// batch starts a new one, so the warp waits for one texture round trip per four iterations instead of one per
						// iteration with its lanes out of phase.  (Fetching early returns the same value: positions depend on the step
						// index alone.)  BLOCK mode is texture-throughput-bound rather than latency-bound (measured: alignment costs 10 %
						// there and gains 10 % with the distance maps), so its lanes refill on their own.
						if (TRACE) { const long long c = clock64(); tr.tc_top += c - tr.tc_mark; tr.tc_mark = c; }
						constexpr bool kLookAhead = !OTF && !EXACT;
						unsigned       dist       = 0u;
						int            k          = 0;
						if (SKIP != VKV_SKIP_NONE && do_skip) dist = __ldg(Dm + idx);
						bool any_refill = false;
						if (kLookAhead) {
							k                = i - pre_base;
							const bool vmode = !do_skip;
							const bool need  = vmode && (unsigned) k >= 4u;
							if (SKIP != VKV_SKIP_BLOCK || TRACE) any_refill = __any_sync(__activemask(), need);        // BLOCK refills per lane: no vote
							if (SKIP == VKV_SKIP_BLOCK ? need : (any_refill && vmode && k != 0)) {
								pre_base = i;
								k        = 0;
								const float f1 = (float) (i + 1), f2 = (float) (i + 2), f3 = (float) (i + 3);
								const float q1[3] = {madf_<kFast>(f1, step[0], entry[0]), madf_<kFast>(f1, step[1], entry[1]), madf_<kFast>(f1, step[2], entry[2])};
								const float q2[3] = {madf_<kFast>(f2, step[0], entry[0]), madf_<kFast>(f2, step[1], entry[1]), madf_<kFast>(f2, step[2], entry[2])};
								const float q3[3] = {madf_<kFast>(f3, step[0], entry[0]), madf_<kFast>(f3, step[1], entry[1]), madf_<kFast>(f3, step[2], entry[2])};
								pre_v0 = tex3D<float>(P.tex_v, pos[0], pos[1], pos[2]);
								pre_v1 = tex3D<float>(P.tex_v, q1[0], q1[1], q1[2]);
								pre_v2 = tex3D<float>(P.tex_v, q2[0], q2[1], q2[2]);
								pre_v3 = tex3D<float>(P.tex_v, q3[0], q3[1], q3[2]);
								if (use_g) {
									pre_g0 = tex3D<float>(P.tex_g, pos[0], pos[1], pos[2]);
									pre_g1 = tex3D<float>(P.tex_g, q1[0], q1[1], q1[2]);
									pre_g2 = tex3D<float>(P.tex_g, q2[0], q2[1], q2[2]);
									pre_g3 = tex3D<float>(P.tex_g, q3[0], q3[1], q3[2]);
								}
							}
						}
						if (TRACE) {
							{ const long long c = clock64(); tr.tc_req += c - tr.tc_mark; tr.tc_mark = c; }
							const unsigned am = __activemask();
							tr.tr_lanes += __popc(am);
							tr.tr_d += __any_sync(am, do_skip) ? 1u : 0u;
							tr.tr_r += any_refill ? 1u : 0u;
							tr.tr_mixed += (__any_sync(am, do_skip) && __any_sync(am, !do_skip)) ? 1u : 0u;
						}
						if (SKIP != VKV_SKIP_NONE && do_skip) {
							++n_dist;
							if (dist > 0u) {
								float       dxyz[3];
								const float fd = (float) dist, omfd = 1.0f - fd;        // exact (dist <= 255)
#pragma unroll
								for (int k = 0; k < 3; ++k) {
									const float rr   = clampf_((float) u_i[k] - u[k], -1.0f, 0.0f);
									const float base = SKIP == VKV_SKIP_BLOCK ? (neg[k] ? 0.0f : 1.0f) : (neg[k] ? omfd : fd);
									dxyz[k]          = (base + rr) * sdt_inv[k];
								}
								const float m       = fminf(fminf(dxyz[0], dxyz[1]), dxyz[2]);
								int         i_delta = (int) ceilf(m);
								if (i_delta < 1) i_delta = 1;
								i += i_delta;
							} else {
								voxel_occupied = true;
								idx_last       = idx;
								i              = max(i - back, i_min);
							}
							if (TRACE) { const long long c = clock64(); tr.tc_d += c - tr.tc_mark; tr.tc_mark = c; }
						} else {
							if (TRACE) tr.tc_mark = clock64();
							++n_vol;
							float intensity, gradient = 1.0f;
							if (OTF) {
								intensity = EXACT ? sample_exact(P.V, P.dim, pos[0], pos[1], pos[2]) : tex3D<float>(P.tex_v, pos[0], pos[1], pos[2]);
								if (use_g)
									gradient = gradient_otf<EXACT>(P.tex_v, P.V, P.dim, dim_inv, P.grad_modifier, pos[0], pos[1], pos[2]);
							} else if (EXACT) {
								intensity = sample_exact(P.V, P.dim, pos[0], pos[1], pos[2]);
								if (use_g) gradient = sample_exact(P.G, P.dim, pos[0], pos[1], pos[2]);
							} else {
								intensity = pick4_(k, pre_v0, pre_v1, pre_v2, pre_v3);
								if (use_g) gradient = pick4_(k, pre_g0, pre_g1, pre_g2, pre_g3);
							}
							const int ti = tf_texel(intensity), tg = GRAD ? tf_texel(gradient) : 255;        // gradient = 1.0 without a gradient TF
							float4    c;
							c = __ldg(P.ctab + tg * 256 + ti);        // (skipping the read for texels outside the TF's visible rectangle was measured: no change here)
							voxel_occupied = c.w >= 0.0f;
							if (voxel_occupied) {
								if (SKIP != VKV_SKIP_NONE) idx_last = idx;
								const float w = 1.0f - out[3];
								out[0] = madf_<kFast>(w, c.x, out[0]); out[1] = madf_<kFast>(w, c.y, out[1]); out[2] = madf_<kFast>(w, c.z, out[2]); out[3] = madf_<kFast>(w, c.w, out[3]);
								if (c.w > 0.0f) i_first_hit = i;
								if (out[3] > 0.99f && P.ert) {
									out[3]   = 1.0f;
									ert_done = true;
									break;
								}
							} else {
								if (COUNT) ++n_empty;
							}
							++i;
							if (SKIP != VKV_SKIP_NONE) i_min = i;
							if (TRACE) { const long long c = clock64(); tr.tc_v += c - tr.tc_mark; tr.tc_mark = c; }
						}
					}
					if (!kLong || ert_done || i >= n_steps) break;
					{
						// the trip limit was reached with the ray still marching.  Every such lane of the warp arrives here in the same trip
						// (they count trips together): one atomic for the warp, one 80-byte record per lane
						const unsigned am     = __activemask();
						const int      leader = __ffs(am) - 1;
						unsigned       slot   = 0u;
						if (lane == leader) slot = atomicAdd(&P.lq->count, (unsigned) __popc(am));
						slot = __shfl_sync(am, slot, leader) + (unsigned) __popc(am & ((1u << lane) - 1u));
						if (slot < (unsigned) P.long_cap) {
							float4 *r = reinterpret_cast<float4 *>(P.lrays + slot);
							r[0] = make_float4(__uint_as_float((unsigned) p), __int_as_float(i), __int_as_float(i_min), __int_as_float(i_first_hit));
							r[2] = make_float4(out[0], out[1], out[2], out[3]);
							r[3] = make_float4(entry[0], entry[1], entry[2], step[0]);
							r[4] = make_float4(step[1], step[2], 0.0f, 0.0f);
							r[1] = make_float4(__int_as_float(n_steps), __uint_as_float(idx_last), __uint_as_float(voxel_occupied ? 1u : 0u), 0.0f);
							suspended = true;
							break;
						}
						trip_limit = 0xffffffffu;        // queue full: this ray finishes here
					}
					}
					if (P.depth && out[3] > 0.0f && i_first_hit < n_steps) {
						const double pm[3] = {(double) (entry[0] + step[0] * (float) i_first_hit) - 0.5,
						                      (double) (entry[1] + step[1] * (float) i_first_hit) - 0.5,
						                      (double) (entry[2] + step[2] * (float) i_first_hit) - 0.5};
						const double z = P.pvm_z[0] * pm[0] + P.pvm_z[1] * pm[1] + P.pvm_z[2] * pm[2] + P.pvm_z[3];
						const double w = P.pvm_w[0] * pm[0] + P.pvm_w[1] * pm[1] + P.pvm_w[2] * pm[2] + P.pvm_w[3];
						frag_depth     = (float) (z / w);
					}
					if (P.test == VKV_TEST_NUM_TEXTURE_SAMPLES) {
						const unsigned n_max = (unsigned) (ceilf((float) P.dim_max * sqrtf(3.0f)) * P.sampling_factor);
						const float    s     = (float) (n_vol + n_dist) / (float) n_max;
						out[0] = out[1] = out[2] = s;
						out[3] = 1.0f;
					}
				}
			}
		}
		// depth test GREATER_OR_EQUAL + depth write, then blend: rgb = src.rgb + dst.rgb*(1-src.a), a = src.a*(1-src.a) + dst.a*0
		// (volume_render_subpass.cpp:176-190); R8G8B8A8_SRGB store.  Without LOAD the destination is the clear colour
		// (0,0,0,1) / depth 0 and uncovered pixels are written with it.
		const bool pass = covered && !discarded && frag_depth >= dst_depth;
		if (kLong && suspended) {
			// raycast_long_kernel finishes this ray and stores its pixel
		} else if (!LOAD) {
			reinterpret_cast<unsigned *>(P.rgba8)[p] = pack_over_clear<kFast>(pass, out);
			if (P.depth) P.depth[p] = pass ? frag_depth : 0.0f;
		} else if (pass) {
			const unsigned d8 = reinterpret_cast<const unsigned *>(P.rgba8)[p];
			const float    sa = clampf_(out[3], 0.0f, 1.0f), om = 1.0f - sa;
			const float    r  = clampf_(out[0], 0.0f, 1.0f) + srgb_decode(d8 & 0xffu) * om;
			const float    g  = clampf_(out[1], 0.0f, 1.0f) + srgb_decode((d8 >> 8) & 0xffu) * om;
			const float    b  = clampf_(out[2], 0.0f, 1.0f) + srgb_decode((d8 >> 16) & 0xffu) * om;
			reinterpret_cast<unsigned *>(P.rgba8)[p] = unorm8(srgb_encode(clampf_(r, 0.0f, 1.0f))) | (unorm8(srgb_encode(clampf_(g, 0.0f, 1.0f))) << 8) |
			                                           (unorm8(srgb_encode(clampf_(b, 0.0f, 1.0f))) << 16) | (unorm8(sa * om) << 24);
			if (P.depth) P.depth[p] = frag_depth;
		}
	}
	covered_acc += covered;
	n_iter_out = n_iter;
}

// A CTA is two warps = a 16x4 pixel tile; small CTAs keep the register file busy while long rays finish.
// grid = (CTAs per tile in x, CTAs per tile in y, tiles of this launch).
template <int SKIP, bool EXACT, bool COUNT, bool OTF = false, bool TRACE = false, bool LOAD = false, bool GRAD = true>
__global__ void __launch_bounds__(kRcThreads, ((OTF || LOAD) ? 8 : VKV_RC_MIN_CTAS) * 2 / kRcWarps) raycast_kernel(const __grid_constant__ RayParams P)
{
	__shared__ unsigned long long s_cnt[kRcWarps][4];

	// CTA -> tile -> pixel
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long t_start = 0;
	unsigned           n_iter  = 0;
	RcTraceAcc         tr;
	if (TRACE) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
	// Tiles are issued from the middle of this launch's tile list outwards (m, m-1, m+1, m-2, ...): the long rays sit
	// near the image centre, so their latency chains start at t = 0 and the cheap border tiles fill the tail.
	// With a history (the previous frame of this volume at this frame size) the tiles are issued in decreasing order of the loop
	// count their longest ray needed last time: the frame time is the critical path of a few hundred long warps (profiles/
	// r1s_trace.md), so they have to start at t = 0, not whenever the sweep reaches them.
	// (compiled into the distance-map production variants only: the other modes are throughput-bound and keep their old code)
	constexpr bool kHist = (SKIP == VKV_SKIP_DISTANCE || SKIP == VKV_SKIP_ANISOTROPIC_DISTANCE) && !OTF && !EXACT && !LOAD;
	const int seq        = P.seq_base + (int) blockIdx.z;
	const int local_tile = (kHist && P.tile_order) ? (int) P.tile_order[seq] : centre_out(seq, P.my_tiles);
	const int tile       = P.tile_first + local_tile * P.tile_stride;
	const int tile_y = P.tiles_x_magic ? (int) __umulhi((unsigned) tile, P.tiles_x_magic) : tile / P.tiles_x, tile_x = tile - tile_y * P.tiles_x;
	const int tx0 = tile_x * P.tile_w + (int) blockIdx.x * 16;
	const int ty0 = tile_y * P.tile_h + (int) blockIdx.y * kRcRows;
	const int px = tx0 + (warp & 1) * 8 + (lane & 7);
	const int py = ty0 + (warp >> 1) * 4 + (lane >> 3);
	const bool   in_frame = px < P.width && py < P.height;
	const size_t p        = (size_t) py * P.width + px;

	// CTA-uniform rejection against the projected bounds of the unit cube: such pixels keep the clear colour
	// (0,0,0,1) (render_pipeline.cpp:38) and depth 0.
	if (tx0 > P.bbox[2] || tx0 + 15 < P.bbox[0] || ty0 > P.bbox[3] || ty0 + kRcRows - 1 < P.bbox[1]) {
		if (in_frame && !LOAD) {
			reinterpret_cast<unsigned *>(P.rgba8)[p] = 0xff000000u;
			if (P.depth) P.depth[p] = 0.0f;
		}
		return;
	}

	unsigned n_vol = 0, n_dist = 0, n_empty = 0, covered = 0;
	rc_cast_pixel<SKIP, EXACT, COUNT, OTF, TRACE, LOAD, GRAD>(P, px, py, lane, n_vol, n_dist, n_empty, covered, n_iter, tr);

	if (kHist && P.tile_cost) {
		const unsigned it_warp = __reduce_max_sync(0xffffffffu, n_iter);
		if (lane == 0 && it_warp) atomicMax(P.tile_cost + local_tile, it_warp);
	}
	if (TRACE) {
		unsigned long long t_end;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
		// per-warp record (kTraceWords u64): start ns, end ns, {iterations | with a skip-map load | with a texture batch | with both kinds
		// of lane | mean live lanes}, then cycles of the longest-lived lanes in: loop head, request issue, skip branch, sample branch
		const unsigned it_max = __reduce_max_sync(0xffffffffu, n_iter);
		const unsigned m_d = __reduce_max_sync(0xffffffffu, tr.tr_d), m_r = __reduce_max_sync(0xffffffffu, tr.tr_r), m_m = __reduce_max_sync(0xffffffffu, tr.tr_mixed);
		const unsigned m_l = __reduce_max_sync(0xffffffffu, tr.tr_lanes);
		const unsigned c_top = __reduce_max_sync(0xffffffffu, (unsigned) tr.tc_top), c_req = __reduce_max_sync(0xffffffffu, (unsigned) tr.tc_req);
		const unsigned c_d = __reduce_max_sync(0xffffffffu, (unsigned) tr.tc_d), c_v = __reduce_max_sync(0xffffffffu, (unsigned) tr.tc_v);
		if (lane == 0) {
			const size_t w = (((size_t) blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * kRcWarps + warp;
			unsigned long long *t = P.trace + w * kTraceWords;
			t[0] = t_start; t[1] = t_end;
			t[2] = (unsigned long long) (it_max & 0xfffu) | ((unsigned long long) (m_d & 0xfffu) << 12) | ((unsigned long long) (m_r & 0xfffu) << 24) |
			       ((unsigned long long) (m_m & 0xfffu) << 36) | ((unsigned long long) (it_max ? m_l / it_max : 0u) << 48);
			t[3] = c_top; t[4] = c_req; t[5] = c_d; t[6] = c_v; t[7] = 0;
		}
	}
	if (COUNT) {
		unsigned long long c[4] = {n_vol, n_dist, n_empty, covered};
#pragma unroll
		for (int k = 0; k < 4; ++k) {
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
			if (lane == 0) s_cnt[warp][k] = c[k];
		}
		__syncthreads();
		if (threadIdx.x < 4) {
			unsigned long long t = 0;
#pragma unroll
			for (int w = 0; w < kRcWarps; ++w) t += s_cnt[w][threadIdx.x];
			if (t) atomicAdd(P.counts + threadIdx.x, t);
		}
	}
}

#ifndef VKV_RC_LONG_CTAS
#define VKV_RC_LONG_CTAS 4
#endif
constexpr int kLongWarps = 8, kLongCtasPerSm = VKV_RC_LONG_CTAS;
template <int SKIP>
__global__ void __launch_bounds__(32 * kLongWarps, kLongCtasPerSm) raycast_long_kernel(const __grid_constant__ RayParams P)
{
	__shared__ float4   s_c[kLongWarps][kLongWindow];
	__shared__ unsigned s_idx[kLongWarps][kLongWindow];
	__shared__ int      s_hop[kLongWarps][kLongWindow];        // > 0: trips the skip-map byte of this step's block lets the ray jump; 0: the block is occupied
	const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned n_rays = min(P.lq->count, (unsigned) P.long_cap);
	unsigned long long c_vol = 0ull, c_dist = 0ull, c_empty = 0ull;
	const LongConsts LC = long_consts(P);
	for (;;) {
		unsigned j = 0u;
		if (lane == 0) j = atomicAdd(&P.lq->head, 1u);
		j = __shfl_sync(0xffffffffu, j, 0);
		if (j >= n_rays) break;
		rc_long_ray<SKIP>(P, LC, j, lane, s_c[warp], s_idx[warp], s_hop[warp], c_vol, c_dist, c_empty);
	}
	if (P.counts && lane == 0 && (c_vol | c_dist | c_empty)) {
		atomicAdd(P.counts + 0, c_vol);
		atomicAdd(P.counts + 1, c_dist);
		atomicAdd(P.counts + 2, c_empty);
	}
	// the last CTA out publishes the frame's long-ray count (host-visible hint, read a frame or more later without synchronising)
	// and resets the queue for the next frame (every CTA read `count` before it got here)
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(&P.lq->done, 1u) == gridDim.x - 1u) {
			if (P.long_hint) *P.long_hint = (int) min(P.lq->count, 0x7fffffffu);
			P.lq->count = 0u;
			P.lq->head  = 0u;
			P.lq->done  = 0u;
		}
	}
}

// ---- tile issue order from the previous frame's cost --------------------------------------------------------------------
// One CTA.  cost[t] = largest warp loop count inside tile t of the launch's tile list last frame.  A frame whose longest rays are
// few is bound by their latency chains: its long tiles (cost >= 3/4, then >= 1/2 of the frame's maximum, after a 3x3 dilation over
// the tile grid that absorbs the camera motion between two frames — full-frame lists only: in a strided multi-GPU list the
// neighbours belong to other ranks) are promoted to the front; all other tiles keep the centre-out order, which keeps the tiles
// that run together neighbours in the volume (texture and L2 locality).  When more than a third of the tiles are long the frame is
// throughput-bound (ESS off, block skipping, small volumes) and nothing is promoted: measured, reordering costs those 5-13 %.
// Also clears cost[] for the frame about to start.

__global__ void __launch_bounds__(1024) tile_order_kernel(unsigned *__restrict__ cost, unsigned *__restrict__ order, int n, int tiles_x, int tiles_y, int dilate, int *__restrict__ decision)
{
	extern __shared__ unsigned s_cost[];        // n raw costs, then n dilated costs
	unsigned *s_dil = s_cost + n;
	__shared__ unsigned           s_max;
	__shared__ unsigned long long s_warp[32];
	__shared__ unsigned long long s_total;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_max = 0u;
	for (int t = threadIdx.x; t < n; t += blockDim.x) s_cost[t] = cost[t];
	__syncthreads();
	unsigned mx = 0u;
	for (int t = threadIdx.x; t < n; t += blockDim.x) {
		unsigned c = s_cost[t];
		if (dilate) {
			const int ty = t / tiles_x, tx = t - ty * tiles_x;
			for (int dy = -1; dy <= 1; ++dy)
				for (int dx = -1; dx <= 1; ++dx) {
					const int x = tx + dx, y = ty + dy;
					if (x >= 0 && x < tiles_x && y >= 0 && y < tiles_y) c = max(c, s_cost[y * tiles_x + x]);
				}
		}
		s_dil[t] = c;
		mx       = max(mx, c);
		cost[t]  = 0u;
	}
	mx = __reduce_max_sync(0xffffffffu, mx);
	if (lane == 0) atomicMax(&s_max, mx);
	__syncthreads();
	const unsigned hi = max(1u, s_max - s_max / 4u), lo = max(1u, s_max / 2u);
	auto group = [&](int t) { const unsigned c = s_dil[t]; return c >= hi ? 0 : (c >= lo ? 1 : 2); };
	// each thread owns a run of the centre-out sequence; counts per group packed 3 x 20 bits
	const int per = (n + (int) blockDim.x - 1) / (int) blockDim.x;
	const int s0 = min(n, (int) threadIdx.x * per), s1 = min(n, s0 + per);
	unsigned long long mine = 0ull;
	for (int q = s0; q < s1; ++q) mine += 1ull << (20 * group(centre_out(q, n)));
	unsigned long long incl = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += v;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		unsigned long long w = s_warp[lane], wi = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o);
			if (lane >= o) wi += v;
		}
		s_warp[lane] = wi - w;        // exclusive prefix of the warps
		if (lane == 31) s_total = wi;
	}
	__syncthreads();
	const unsigned long long excl = s_warp[warp] + incl - mine, total = s_total;
	const unsigned n0 = (unsigned) (total & 0xfffffu), n1 = (unsigned) ((total >> 20) & 0xfffffu);
	const bool     promote = (n0 + n1) * 3u <= (unsigned) n;
	if (threadIdx.x == 0) *decision = promote ? (int) (n0 + n1) : 0;        // host-visible hint (read a frame or more later, without synchronising): 0 = keep centre-out, else the number of promoted tiles
	unsigned       pos[3] = {(unsigned) (excl & 0xfffffu), n0 + (unsigned) ((excl >> 20) & 0xfffffu), n0 + n1 + (unsigned) ((excl >> 40) & 0xfffffu)};
	for (int q = s0; q < s1; ++q) {
		const int t = centre_out(q, n);
		if (promote) order[pos[group(t)]++] = (unsigned) t;
		else order[q] = (unsigned) t;
	}
}

// ---- host side: build RayParams in fp64 from the reference's uniforms ------------------------
static void mul_mv(const double *m, const double *v, double *out)
{
	for (int r = 0; r < 4; ++r) out[r] = m[0 + r] * v[0] + m[4 + r] * v[1] + m[8 + r] * v[2] + m[12 + r] * v[3];
}
static void mul_mm(const double *a, const double *b, double *out)
{
	double t[16];
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 4; ++r) {
			double s = 0;
			for (int k = 0; k < 4; ++k) s += a[k * 4 + r] * b[c * 4 + k];
			t[c * 4 + r] = s;
		}
	for (int i = 0; i < 16; ++i) out[i] = t[i];
}

static void far_dir(const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray, double px, double py, int W, int H, double d[3])
{
	double vpi[16], mi[16];
	for (int i = 0; i < 16; ++i) { vpi[i] = cam->view_proj_inv[i]; mi[i] = cam->model_inv[i]; }
	const double ndc[4] = {2.0 * (px + 0.5) / W - 1.0, 2.0 * (py + 0.5) / H - 1.0, 0.0, 1.0};        // z = 0: far plane (reverse-Z)
	double       wp[4], mp[4];
	mul_mv(vpi, ndc, wp);
	for (int k = 0; k < 3; ++k) wp[k] /= wp[3];
	wp[3] = 1.0;
	mul_mv(mi, wp, mp);
	for (int k = 0; k < 3; ++k) d[k] = mp[k] + 0.5 - (double) ray->cam_pos_tex[k];
}

// Conservative screen-space bounds (pixel indices, inclusive) of the unit cube [0,1]^3 in texture space.
// A pixel (px, py) looks along d0 + px*ddx + py*ddy; corner c is seen at the pixel solving
// [ddx ddy d0] (l*px, l*py, l)^T = c - o with l > 0.  Any corner at or behind the camera plane, a camera
// inside the cube or a degenerate basis fall back to the whole frame.
static void screen_bbox(const RayParams &P, int width, int height, int out[4])
{
	out[0] = 0; out[1] = 0; out[2] = width - 1; out[3] = height - 1;
	const double eps = 1e-6;
	if (P.o[0] >= -eps && P.o[0] <= 1 + eps && P.o[1] >= -eps && P.o[1] <= 1 + eps && P.o[2] >= -eps && P.o[2] <= 1 + eps) return;
	const double a[3] = {P.ddx[0], P.ddx[1], P.ddx[2]}, b[3] = {P.ddy[0], P.ddy[1], P.ddy[2]}, c[3] = {P.d0[0], P.d0[1], P.d0[2]};
	const double bc[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
	const double ca[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
	const double ab[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
	const double det   = a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2];
	const double scale = (std::fabs(a[0]) + std::fabs(a[1]) + std::fabs(a[2])) * (std::fabs(b[0]) + std::fabs(b[1]) + std::fabs(b[2])) *
	                     (std::fabs(c[0]) + std::fabs(c[1]) + std::fabs(c[2]));
	if (!(std::fabs(det) > 1e-12 * scale)) return;
	double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
	for (int corner = 0; corner < 8; ++corner) {
		const double r[3] = {(corner & 1 ? 1.0 : 0.0) - P.o[0], (corner & 2 ? 1.0 : 0.0) - P.o[1], (corner & 4 ? 1.0 : 0.0) - P.o[2]};
		const double lx = (r[0] * bc[0] + r[1] * bc[1] + r[2] * bc[2]) / det;
		const double ly = (r[0] * ca[0] + r[1] * ca[1] + r[2] * ca[2]) / det;
		const double l  = (r[0] * ab[0] + r[1] * ab[1] + r[2] * ab[2]) / det;
		if (!(l > 1e-9)) return;        // corner not strictly in front of the camera: no finite bound
		const double x = lx / l, y = ly / l;
		xmin = std::fmin(xmin, x); xmax = std::fmax(xmax, x);
		ymin = std::fmin(ymin, y); ymax = std::fmax(ymax, y);
	}
	if (!(std::isfinite(xmin) && std::isfinite(xmax) && std::isfinite(ymin) && std::isfinite(ymax))) return;
	// one pixel of slack on every side for rounding; an off-screen cube gives an empty rectangle
	const double x0 = std::floor(xmin) - 1.0, x1 = std::ceil(xmax) + 1.0, y0 = std::floor(ymin) - 1.0, y1 = std::ceil(ymax) + 1.0;
	out[0] = (int) std::fmax(x0, 0.0); out[2] = (int) std::fmin(x1, (double) width - 1.0);
	out[1] = (int) std::fmax(y0, 0.0); out[3] = (int) std::fmin(y1, (double) height - 1.0);
	if (x1 < 0.0 || y1 < 0.0 || x0 > width - 1.0 || y0 > height - 1.0) { out[0] = 1; out[2] = 0; out[1] = 1; out[3] = 0; }
}

int launch_render(vkv_volume *vol, const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                  const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int width, int height, int tile_w,
                  int tile_h, int tile_first, int tile_stride, int tile_limit, uint8_t *rgba8, float *depth,
                  vkv_sample_counts *counts, cudaStream_t s)
{
	RayParams P{};
	double    dx1[3], dy1[3];
	far_dir(cam, ray, 0, 0, width, height, P.d0);
	far_dir(cam, ray, 1, 0, width, height, dx1);
	far_dir(cam, ray, 0, 1, width, height, dy1);
	for (int k = 0; k < 3; ++k) {
		P.ddx[k] = dx1[k] - P.d0[k];
		P.ddy[k] = dy1[k] - P.d0[k];
		P.o[k]   = ray->cam_pos_tex[k];
		P.cam_pos_tex[k] = ray->cam_pos_tex[k];
		P.dimf[k]        = (float) vol->dim[k];
		P.dim[k]         = (int) vol->dim[k];
		P.dim_b[k]       = (int) vol->dim_b[k];
		P.block_size[k]  = ray->block_size[k];
		P.vol_to_map[k]  = P.dimf[k] / P.block_size[k];
		P.fd0[k] = (float) P.d0[k]; P.fddx[k] = (float) P.ddx[k]; P.fddy[k] = (float) P.ddy[k];
		P.fo[k]  = (float) P.o[k];
		P.fplane[k] = ray->plane_tex[k];
	}
	for (int k = 0; k < 4; ++k) P.plane[k] = ray->plane_tex[k];
	{
		// same operation order as the per-pixel evaluation used to have (no contraction)
		volatile double t = P.plane[0] * P.o[0];
		t = t + P.plane[1] * P.o[1];
		t = t + P.plane[2] * P.o[2];
		t = t + P.plane[3];
		P.s0  = t;
		P.fs0 = P.fplane[0] * P.fo[0] + P.fplane[1] * P.fo[1] + P.fplane[2] * P.fo[2] + ray->plane_tex[3];
	}
	{
		double pr[16], vw[16], md[16], pv[16], pvm[16];
		for (int i = 0; i < 16; ++i) { pr[i] = cam->proj[i]; vw[i] = cam->view[i]; md[i] = cam->model[i]; }
		mul_mm(pr, vw, pv);
		mul_mm(pv, md, pvm);
		for (int c = 0; c < 4; ++c) { P.pvm_x[c] = pvm[c * 4 + 0]; P.pvm_y[c] = pvm[c * 4 + 1]; P.pvm_z[c] = pvm[c * 4 + 2]; P.pvm_w[c] = pvm[c * 4 + 3]; }
		for (int i = 0; i < 16; ++i) { P.vpi[i] = cam->view_proj_inv[i]; P.mi[i] = cam->model_inv[i]; }
	}
	P.dim_max             = (int) std::max(vol->dim[0], std::max(vol->dim[1], vol->dim[2]));
	P.sampling_factor     = tfu->sampling_factor;
	P.sampling_factor_inv = 1.0f / tfu->sampling_factor;
	P.voxel_alpha_factor  = tfu->voxel_alpha_factor;
	P.grad_modifier       = tfu->grad_magnitude_modifier;
	P.use_gradient        = tfu->use_gradient ? 1 : 0;
	P.ert                 = opt->early_ray_termination ? 1 : 0;
	P.test                = opt->test;
	P.depth_attachment    = opt->depth_attachment ? 1 : 0;
	P.width = width; P.height = height;
	P.tile_w = tile_w; P.tile_h = tile_h;
	P.tiles_x         = (width + tile_w - 1) / tile_w;
	const int tiles_y = (height + tile_h - 1) / tile_h;
	const int n_tiles = P.tiles_x * tiles_y;
	// exact while tile * tiles_x < 2^32 (Granlund-Montgomery round-up reciprocal)
	P.tiles_x_magic = (P.tiles_x > 1 && (unsigned long long) n_tiles * (unsigned) P.tiles_x < (1ull << 32)) ? (unsigned) ((1ull << 32) / (unsigned) P.tiles_x) + 1u : 0u;
	P.tile_first = tile_first; P.tile_stride = tile_stride;
	int my_tiles = tile_first < n_tiles ? (n_tiles - tile_first + tile_stride - 1) / tile_stride : 0;
	if (tile_limit >= 0 && my_tiles > tile_limit) my_tiles = tile_limit;
	if (my_tiles == 0) return VKV_OK;
	P.my_tiles = my_tiles;
	screen_bbox(P, width, height, P.bbox);
	// premultiplied colour table: rebuilt when the TF texture, the sampling factor or the alpha factor changed
	if (!vol->d_ctab) VKV_CUDA_CHECK(cudaMalloc(&vol->d_ctab, 256 * 256 * sizeof(float4)));
	if (vol->ctab_tf_version != vol->tf_version || vol->ctab_sampling != tfu->sampling_factor || vol->ctab_alpha != tfu->voxel_alpha_factor) {
		ctab_kernel<<<256, 256, 0, s>>>(reinterpret_cast<const uchar4 *>(vol->d_tf), reinterpret_cast<float4 *>(vol->d_ctab),
		                                tfu->voxel_alpha_factor, 1.0f / tfu->sampling_factor);
		VKV_LAUNCHED();
		vol->ctab_tf_version = vol->tf_version;
		vol->ctab_sampling   = tfu->sampling_factor;
		vol->ctab_alpha      = tfu->voxel_alpha_factor;
	}
	P.ctab  = reinterpret_cast<const float4 *>(vol->d_ctab);
	P.tex_v = vol->t_V; P.tex_g = vol->t_G;
	P.V = vol->d_V; P.G = vol->d_G;
	P.tf = reinterpret_cast<const uchar4 *>(vol->d_tf);
	for (int i = 0; i < 8; ++i) P.map_ptrs[i] = i < (int) vol->d_maps.size() ? vol->d_maps[i] : nullptr;
	P.maps   = P.map_ptrs[0];
	P.rgba8  = rgba8;
	P.depth  = depth;
	P.counts = reinterpret_cast<unsigned long long *>(counts);
	P.bounds = vol->d_bounds;
	{
		const char *fl = getenv("VKV_RC_FLAGS");
		P.flags        = fl ? atoi(fl) : 0;
	}

	const bool exact = opt->filter == VKV_FILTER_EXACT;
	// on-the-fly gradients (volume created without a gradient map); the variant always counts, into scratch if need be
	const bool otf = P.use_gradient && !vol->precomputed_gradient;
	const bool load = opt->load_framebuffer != 0;
	// long rays go to raycast_long_kernel (distance-map production variants, plain view): VKV_RC_LONG_T=<trips> moves the hand-over
	// point, 0 keeps every ray in the one-kernel march
	int long_T = 64;
	if (const char *lt = getenv("VKV_RC_LONG_T")) long_T = atoi(lt);
	const bool use_long = long_T > 0 && (opt->skipping_type == VKV_SKIP_DISTANCE || opt->skipping_type == VKV_SKIP_ANISOTROPIC_DISTANCE) && !otf &&
	                      !exact && !load && opt->test == VKV_TEST_NONE && !getenv("VKV_RC_TRACE");
	// Frames with MANY long rays are throughput-bound: one ray per warp costs them more issue slots than it saves latency (measured:
	// 256^3 blobs at 512x512, +115 %).  The hand-over is therefore kept only while the previous frame's count stayed below
	// kLongMaxRays; a frame above it switches it off for 32 frames, then it is tried again.
	constexpr int kLongMaxRays = 20000;
	if (use_long && vol->h_long_hint) {
		if (vol->long_holdoff == 0 && *static_cast<volatile int *>(vol->h_long_hint) > kLongMaxRays && !getenv("VKV_RC_LONG_ALWAYS")) {
			vol->long_holdoff = 32;
			*vol->h_long_hint = 0;
		}
	}
	const bool long_now = use_long && vol->long_holdoff == 0;
	if (use_long && vol->long_holdoff > 0) --vol->long_holdoff;
	if (long_now) {
		if (!vol->h_long_hint) {
			VKV_CUDA_CHECK(cudaHostAlloc(&vol->h_long_hint, sizeof(int), cudaHostAllocMapped));
			*vol->h_long_hint = 0;
		}
		if (!vol->d_lq) {
			constexpr int kCap = 1 << 18;        // 262 144 rays x 80 B = 21 MB; a frame with more long rays keeps the rest in the march
			VKV_CUDA_CHECK(cudaMalloc(&vol->d_lq, sizeof(LongQueue)));
			VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_lq, 0, sizeof(LongQueue), s));
			VKV_CUDA_CHECK(cudaMalloc(&vol->d_lrays, (size_t) kCap * sizeof(LongRay)));
			vol->long_cap = kCap;
		}
		P.lq       = static_cast<LongQueue *>(vol->d_lq);
		P.lrays    = static_cast<LongRay *>(vol->d_lrays);
		P.long_T   = long_T;
		P.long_cap = vol->long_cap;
		P.long_hint = vol->h_long_hint;

		if (getenv("VKV_RC_DEBUG")) fprintf(stderr, "[vkv] long rays last frame: %d\n", *vol->h_long_hint);
	}
	if ((otf || load) && !P.counts) P.counts = reinterpret_cast<unsigned long long *>(vol->d_counts_scratch);
	// gridDim.z is limited to 65535: launch the tile list in chunks (one chunk up to 134 Mpixel with 64x32 tiles)
	// debug: VKV_RC_TRACE=<file> dumps per-warp {start ns, end ns, loop iterations} of this launch (synchronous; never set in production)
	const char         *trace_path = getenv("VKV_RC_TRACE");
	unsigned long long *d_trace    = nullptr;
	size_t              trace_n    = 0;
	if (trace_path && my_tiles <= 65535) {
		trace_n = (size_t) my_tiles * (tile_w / 16) * (tile_h / kRcRows) * kRcWarps * kTraceWords;
		VKV_CUDA_CHECK(cudaMalloc(&d_trace, trace_n * sizeof(unsigned long long)));
		VKV_CUDA_CHECK(cudaMemsetAsync(d_trace, 0, trace_n * sizeof(unsigned long long), s));
		P.trace = d_trace;
	}
	// tile scheduling history (see tile_order_kernel); VKV_RC_NO_HISTORY=1 keeps the centre-out order.  The ordering pass for frame
	// k + 1 needs nothing but frame k's march: it is launched on a side stream as soon as that march is, and runs beside frame k's
	// long-ray pass instead of in front of frame k + 1 (one CTA, ~5 us: 5 % of the headline frame when it sat on the critical path).
	bool use_hist = false;
	int  tiles_y_hist = tiles_y;
	{
		const char *no_hist  = getenv("VKV_RC_NO_HISTORY");
		// distance-map modes only: block skipping and ESS off are throughput-bound at every size measured (promotion costs them 2 %);
		// up to 6000 tiles: an 8K frame (16 200 tiles) is throughput-bound too (measured with the limit lifted: 1.394 vs 1.343 ms)
		use_hist = !(no_hist && atoi(no_hist) != 0) && my_tiles >= 64 && my_tiles <= 6000 &&
		           (opt->skipping_type == VKV_SKIP_DISTANCE || opt->skipping_type == VKV_SKIP_ANISOTROPIC_DISTANCE) && !otf && !exact && !load;
		const int   key[8]   = {width, height, tile_w, tile_h, tile_first, tile_stride, my_tiles, opt->skipping_type};
		if (use_hist) {
			if (vol->tile_hist_capacity < my_tiles) {
				if (vol->tile_order_prepared) VKV_CUDA_CHECK(cudaStreamSynchronize(vol->side_stream));
				vol->tile_order_prepared = false;
				cudaFree(vol->d_tile_cost);
				cudaFree(vol->d_tile_order);
				vol->d_tile_cost = vol->d_tile_order = nullptr;
				vol->tile_hist_capacity = 0;
				vol->tile_hist_valid    = false;
				VKV_CUDA_CHECK(cudaMalloc(&vol->d_tile_cost, (size_t) my_tiles * sizeof(unsigned)));
				VKV_CUDA_CHECK(cudaMalloc(&vol->d_tile_order, (size_t) my_tiles * sizeof(unsigned)));
				vol->tile_hist_capacity = my_tiles;
			}
			if (!vol->side_stream) {
				VKV_CUDA_CHECK(cudaStreamCreateWithFlags(&vol->side_stream, cudaStreamNonBlocking));
				VKV_CUDA_CHECK(cudaEventCreateWithFlags(&vol->ev_march, cudaEventDisableTiming));
				VKV_CUDA_CHECK(cudaEventCreateWithFlags(&vol->ev_order, cudaEventDisableTiming));
			}
			const bool same = vol->tile_hist_valid && std::equal(key, key + 8, vol->tile_hist_key);
			if (!vol->h_tile_promote) {
				VKV_CUDA_CHECK(cudaHostAlloc(&vol->h_tile_promote, sizeof(int), cudaHostAllocMapped));
				*vol->h_tile_promote = 1;
			}
			// whatever the side stream was given has to be through before this frame touches the cost and order arrays
			if (vol->tile_order_prepared) VKV_CUDA_CHECK(cudaStreamWaitEvent(s, vol->ev_order, 0));
			if (same && vol->tile_order_prepared) {
				P.tile_order = vol->d_tile_order;        // computed from the previous frame's cost, beside its long-ray pass
			} else if (!same) {
				VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_tile_cost, 0, (size_t) my_tiles * sizeof(unsigned), s));
				vol->tile_order_holdoff = 0;
			}        // (same, nothing prepared: a hold-off frame — centre-out order; the cost keeps accumulating, maximum over the frames in between)
			vol->tile_order_prepared = false;
			P.tile_cost = vol->d_tile_cost;
			std::copy(key, key + 8, vol->tile_hist_key);
		}
		vol->tile_hist_valid = use_hist;
	}
	// after the march of this frame has been launched: the ordering pass for the next frame of the same key
	auto prepare_next_order = [&]() -> int {
		if (!use_hist) return VKV_OK;
		// a frame that did not qualify (throughput-bound) is not asked again for a while: the pass costs ~5 us
		if (vol->tile_order_holdoff == 0 && *static_cast<volatile int *>(vol->h_tile_promote) == 0) {
			vol->tile_order_holdoff = 32;
			*vol->h_tile_promote    = 1;
		}
		if (vol->tile_order_holdoff > 0) {
			--vol->tile_order_holdoff;
			return VKV_OK;
		}
		const int full = (tile_first == 0 && tile_stride == 1 && my_tiles == n_tiles) ? 1 : 0;
		VKV_CUDA_CHECK(cudaEventRecord(vol->ev_march, s));
		VKV_CUDA_CHECK(cudaStreamWaitEvent(vol->side_stream, vol->ev_march, 0));
		tile_order_kernel<<<1, 1024, 2 * (size_t) my_tiles * sizeof(unsigned), vol->side_stream>>>(vol->d_tile_cost, vol->d_tile_order, my_tiles, P.tiles_x, tiles_y_hist, full, vol->h_tile_promote);
		VKV_LAUNCHED();
		VKV_CUDA_CHECK(cudaEventRecord(vol->ev_order, vol->side_stream));
		vol->tile_order_prepared = true;
		return VKV_OK;
	};
	// one launch of the march over `count` tiles of the launch's list starting at `base`, on stream `st`
	auto launch_march = [&](const RayParams &Q, int count, cudaStream_t st) -> int {
		const dim3 grid((unsigned) (tile_w / 16), (unsigned) (tile_h / kRcRows), (unsigned) count);
		const bool nograd = !Q.use_gradient && !getenv("VKV_RC_GRAD_GENERIC");        // (A/B: the generic instantiation)
#define VKV_RC(SK)                                                                      \
	do {                                                                                \
		if (load && otf && exact) raycast_kernel<SK, true, true, true, false, true><<<grid, kRcThreads, 0, st>>>(Q);  \
		else if (load && otf) raycast_kernel<SK, false, true, true, false, true><<<grid, kRcThreads, 0, st>>>(Q);     \
		else if (load && exact) raycast_kernel<SK, true, true, false, false, true><<<grid, kRcThreads, 0, st>>>(Q);   \
		else if (load) raycast_kernel<SK, false, true, false, false, true><<<grid, kRcThreads, 0, st>>>(Q);           \
		else if (d_trace && !otf && !exact) raycast_kernel<SK, false, false, false, true><<<grid, kRcThreads, 0, st>>>(Q); \
		else if (otf && exact) raycast_kernel<SK, true, true, true><<<grid, kRcThreads, 0, st>>>(Q);   \
		else if (otf) raycast_kernel<SK, false, true, true><<<grid, kRcThreads, 0, st>>>(Q);      \
		else if (exact && counts) raycast_kernel<SK, true, true><<<grid, kRcThreads, 0, st>>>(Q);      \
		else if (exact) raycast_kernel<SK, true, false><<<grid, kRcThreads, 0, st>>>(Q);          \
		else if (Q.counts && nograd) raycast_kernel<SK, false, true, false, false, false, false><<<grid, kRcThreads, 0, st>>>(Q); \
		else if (nograd) raycast_kernel<SK, false, false, false, false, false, false><<<grid, kRcThreads, 0, st>>>(Q); \
		else if (Q.counts) raycast_kernel<SK, false, true><<<grid, kRcThreads, 0, st>>>(Q);         \
		else raycast_kernel<SK, false, false><<<grid, kRcThreads, 0, st>>>(Q);                    \
	} while (0)
		switch (opt->skipping_type) {
			case VKV_SKIP_NONE: VKV_RC(VKV_SKIP_NONE); break;
			case VKV_SKIP_BLOCK: VKV_RC(VKV_SKIP_BLOCK); break;
			case VKV_SKIP_DISTANCE: VKV_RC(VKV_SKIP_DISTANCE); break;
			default: VKV_RC(VKV_SKIP_ANISOTROPIC_DISTANCE); break;
		}
#undef VKV_RC
		VKV_LAUNCHED();
		return VKV_OK;
	};
	auto launch_long = [&](cudaStream_t st) -> int {
		const unsigned grid2 = (unsigned) vol->ctx->sm_count * (unsigned) kLongCtasPerSm;
		if (opt->skipping_type == VKV_SKIP_DISTANCE) raycast_long_kernel<VKV_SKIP_DISTANCE><<<grid2, 32 * kLongWarps, 0, st>>>(P);
		else raycast_long_kernel<VKV_SKIP_ANISOTROPIC_DISTANCE><<<grid2, 32 * kLongWarps, 0, st>>>(P);
		VKV_LAUNCHED();
		return VKV_OK;
	};
	// (Measured and dropped: issuing the promoted tiles and the rest as two concurrent launches so that the second pass overlaps the
	// bulk — the promoted tiles' launch itself lasts ~60 us, its rays hand over only after long_T trips: 2 % — and serving the queue
	// from inside the march, by warps about to exit or by a persistent grid: slower, profiles/r2_trace.md.)
	int rc2;
	for (int base = 0; base < my_tiles; base += 65535) {
		P.seq_base = base;
		if ((rc2 = launch_march(P, std::min(65535, my_tiles - base), s))) return rc2;
	}
	if ((rc2 = prepare_next_order())) return rc2;
	if (long_now && (rc2 = launch_long(s))) return rc2;
	if (d_trace) {
		std::vector<unsigned long long> h(trace_n);
		VKV_CUDA_CHECK(cudaStreamSynchronize(s));
		VKV_CUDA_CHECK(cudaMemcpy(h.data(), d_trace, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
		cudaFree(d_trace);
		if (FILE *f = fopen(trace_path, "wb")) {
			fwrite(h.data(), sizeof(unsigned long long), trace_n, f);
			fclose(f);
		}
	}
	return VKV_OK;
}

}        // namespace vkv
