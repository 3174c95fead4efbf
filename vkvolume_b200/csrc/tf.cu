// tf.cu — transfer-function texture and the visibility masks derived from it.
//
// Replaces Volume::update_transfer_function_texture (src/volume_component.cpp:242-278):
// the reference fills the 256x256 RGBA8 texture in a CPU loop and uploads it through a
// staging buffer on every TF change; here one 65 536-thread kernel writes it in HBM with
// the same fp32 operation order (the library is compiled with -fmad=false, so no
// contraction) — bit-identical bytes, no host round trip — and, in the same pass, derives
// what the O(N) occupancy/count kernel needs to classify voxels without touching the texture:
// one bit per texel for each of the two transfer functions the reference really uses
// (texture alpha > 0 for the occupancy map, analytic alphaI*alphaG > 0 for the voxel count,
// SURVEY A.2/A.3), the byte ranges outside which nothing is visible, whether the visible set
// is exactly that rectangle (then a SIMD range test IS the classification) and, when it is
// not, the largest all-visible rectangle anchored at (255, 255) ("sure" thresholds).
#include "common.cuh"

namespace vkv {

__device__ __forceinline__ float clampf_dev(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// per gradient row: statistics of the two masks, written by the row's CTA, reduced by the last CTA to finish
struct TFRowStats {
	int lo_t, hi_t, cnt_t, first_t;        // texture mask: lowest / highest visible intensity, visible texels, lowest v with [v..255] all visible (256: none)
	int lo_a, hi_a, cnt_a, pad;            // analytic mask
};
static_assert(sizeof(TFRowStats) == 32, "TFRowStats layout");

// One CTA per gradient row g (256 texels), one warp per mask word.  FROM_OPTIONS fuses the texture
// generation (same arithmetic as the reference's CPU loop); otherwise the texture is read.
// `have_tfu == 0` leaves the analytic mask empty.
template <bool FROM_OPTIONS>
__global__ void __launch_bounds__(256) tf_masks_kernel(uchar4 *__restrict__ tex, float imin, float imax, float gmin, float gmax,
                                                        vkv_transfer_function_uniform tfu, int have_tfu, uint2 *__restrict__ mask2,
                                                        TFRowStats *__restrict__ rows, unsigned *__restrict__ ticket,
                                                        TFBounds *__restrict__ bounds)
{
	__shared__ unsigned s_mt[8], s_ma[8];
	__shared__ bool     s_last;
	const int g = blockIdx.x, v = threadIdx.x, idx = g * 256 + v;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned char alpha;
	if (FROM_OPTIONS) {
		const float i_inv        = 1.0f / (imax - imin);
		const float g_inv        = 1.0f / (gmax - gmin);
		const bool  use_gradient = gmax != gmin;
		const float alpha_i      = clampf_dev((((float) v / 255.0f) - imin) * i_inv, 0.0f, 1.0f);
		const float alpha_g      = use_gradient ? clampf_dev((((float) g / 255.0f) - gmin) * g_inv, 0.0f, 1.0f) : 1.0f;
		// static_cast<uint8_t>(clamp(alpha_i * alpha_g * 255, 0, 255)): truncation
		alpha    = (unsigned char) clampf_dev(alpha_i * alpha_g * 255.0f, 0.0f, 255.0f);
		tex[idx] = make_uchar4(alpha, alpha, alpha, alpha);
	} else {
		alpha = tex[idx].w;
	}
	bool ana = false;
	if (have_tfu) {
		// analytic TF (shaders/transfer_function.glsl:41-43) with g = 1.0 when gradients are unused
		// (shaders/get_gradient_compute.glsl:6-7); for row 255 float(255)/255 == 1.0 exactly.
		const float aG = clampf_dev((((float) g / 255.0f) - tfu.gradient_min) * tfu.gradient_range_inv, 0.0f, 1.0f);
		const float aI = clampf_dev((((float) v / 255.0f) - tfu.intensity_min) * tfu.intensity_range_inv, 0.0f, 1.0f);
		ana            = aI * aG > 0.0f;
	}
	const unsigned mt = __ballot_sync(0xffffffffu, alpha > 0);
	const unsigned ma = __ballot_sync(0xffffffffu, ana);
	if (lane == 0) {
		mask2[idx >> 5] = make_uint2(mt, ma);
		s_mt[warp] = mt; s_ma[warp] = ma;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		TFRowStats st;
		st.lo_t = st.lo_a = 256; st.hi_t = st.hi_a = -1; st.cnt_t = st.cnt_a = 0; st.first_t = 256; st.pad = 0;
		for (int w = 0; w < 8; ++w) {
			if (s_mt[w]) { st.lo_t = min(st.lo_t, w * 32 + __ffs(s_mt[w]) - 1); st.hi_t = w * 32 + 31 - __clz(s_mt[w]); st.cnt_t += __popc(s_mt[w]); }
			if (s_ma[w]) { st.lo_a = min(st.lo_a, w * 32 + __ffs(s_ma[w]) - 1); st.hi_a = w * 32 + 31 - __clz(s_ma[w]); st.cnt_a += __popc(s_ma[w]); }
		}
		for (int w = 7; w >= 0; --w) {        // run of ones ending at texel 255
			if (s_mt[w] == 0xffffffffu) { st.first_t = w * 32; continue; }
			const int lead = __clz(~s_mt[w]);
			if (lead > 0) st.first_t = w * 32 + 32 - lead;
			break;
		}
		rows[g] = st;
		__threadfence();
		s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (!s_last) return;

	// ---- the last CTA reduces the 256 rows: thread r folds row r with warp reductions + shared atomics ----
	__shared__ int                s_first[256], s_wmax[8];
	__shared__ unsigned           s_red[4][6];        // per set: v_lo, v_hi, g_lo, g_hi, total, unused
	__shared__ unsigned long long s_best;
	__threadfence();
	const TFRowStats me = rows[threadIdx.x];
	const int        r  = threadIdx.x;
	if (threadIdx.x < 4) {
		s_red[threadIdx.x][0] = 0xffffffffu; s_red[threadIdx.x][1] = 0u; s_red[threadIdx.x][2] = 0xffffffffu;
		s_red[threadIdx.x][3] = 0u; s_red[threadIdx.x][4] = 0u;
	}
	if (threadIdx.x == 0) { s_best = 0ull; *ticket = 0u; }        // self-resetting ticket
	s_first[r] = me.first_t;
	const int wm = __reduce_max_sync(0xffffffffu, me.first_t);
	if (lane == 0) s_wmax[warp] = wm;
	__syncthreads();
	auto fold = [&](int set, int lo, int hi, int cnt) {
		const unsigned vlo = cnt ? (unsigned) lo : 0xffffffffu, vhi = cnt ? (unsigned) hi : 0u;
		const unsigned glo = cnt ? (unsigned) r : 0xffffffffu, ghi = cnt ? (unsigned) r : 0u;
		const unsigned a0 = __reduce_min_sync(0xffffffffu, vlo), a1 = __reduce_max_sync(0xffffffffu, vhi);
		const unsigned a2 = __reduce_min_sync(0xffffffffu, glo), a3 = __reduce_max_sync(0xffffffffu, ghi);
		const unsigned a4 = __reduce_add_sync(0xffffffffu, (unsigned) cnt);
		if (lane == 0) {
			atomicMin(&s_red[set][0], a0); atomicMax(&s_red[set][1], a1); atomicMin(&s_red[set][2], a2);
			atomicMax(&s_red[set][3], a3); atomicAdd(&s_red[set][4], a4);
		}
	};
	fold(0, me.lo_t, me.hi_t, me.cnt_t);
	fold(1, me.lo_a, me.hi_a, me.cnt_a);
	fold(2, me.lo_t, me.hi_t, r == 255 ? me.cnt_t : 0);
	fold(3, me.lo_a, me.hi_a, r == 255 ? me.cnt_a : 0);
	{
		// largest all-visible rectangle [vs..255] x [gs..255] of the texture mask: suffix maximum of first_t from row gs up
		int F = 0;
		for (int q = r; q < (warp + 1) * 32; ++q) F = max(F, s_first[q]);
		for (int w = warp + 1; w < 8; ++w) F = max(F, s_wmax[w]);
		if (F < 256) {
			const unsigned long long area = (unsigned long long) (256 - F) * (256 - r);
			atomicMax(&s_best, (area << 20) | ((unsigned long long) F << 8) | (unsigned long long) r);
		}
	}
	__syncthreads();
	if (threadIdx.x < 4) {
		const int k = threadIdx.x;
		TFRange   o;
		o.v_lo = s_red[k][0]; o.v_hi = s_red[k][1]; o.g_lo = s_red[k][2]; o.g_hi = s_red[k][3];
		const unsigned tot = s_red[k][4];
		// the empty set is exactly the empty rectangle (v_lo > v_hi)
		o.exact  = tot == 0u ? 1u : (tot == (o.v_hi - o.v_lo + 1u) * (o.g_hi - o.g_lo + 1u) ? 1u : 0u);
		o.v_sure = 256u; o.g_sure = 256u; o.pad = 0u;
		if (k == 0 && s_best) { o.v_sure = (uint32_t) ((s_best >> 8) & 0xfffu); o.g_sure = (uint32_t) (s_best & 0xffu); }
		if (k == 2 && s_first[255] < 256) { o.v_sure = (uint32_t) s_first[255]; o.g_sure = 255u; }
		TFRange *dst = k == 0 ? &bounds->tex_all : (k == 1 ? &bounds->ana_all : (k == 2 ? &bounds->tex_row255 : &bounds->ana_row255));
		*dst = o;
	}
}

static int ensure_scratch(vkv_volume *vol, cudaStream_t s)
{
	if (vol->d_tf_rows) return VKV_OK;
	VKV_CUDA_CHECK(cudaMalloc(&vol->d_tf_rows, 256 * sizeof(TFRowStats) + 16));
	VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_tf_rows, 0, 256 * sizeof(TFRowStats) + 16, s));        // the ticket starts at 0 and resets itself
	return VKV_OK;
}

int launch_tf_texture(vkv_volume *vol, const vkv_volume_options *opt, cudaStream_t s)
{
	// texture + both masks + bounds in one pass; the analytic mask is built for the uniform these options imply
	vkv_transfer_function_uniform u;
	vkv_transfer_function_uniform_from_options(opt, &u);
	int rc;
	if ((rc = ensure_scratch(vol, s))) return rc;
	auto *rows = reinterpret_cast<TFRowStats *>(vol->d_tf_rows);
	tf_masks_kernel<true><<<256, 256, 0, s>>>(reinterpret_cast<uchar4 *>(vol->d_tf), opt->intensity_min, opt->intensity_max, opt->gradient_min,
	                                          opt->gradient_max, u, 1, vol->d_mask2, rows, reinterpret_cast<unsigned *>(rows + 256), vol->d_bounds);
	VKV_LAUNCHED();
	vol->has_tf = true;
	++vol->tf_version;
	vol->mask_ana_valid = true;
	vol->mask_tfu       = u;
	return VKV_OK;
}

int launch_tf_masks(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, cudaStream_t s)
{
	vkv_transfer_function_uniform u{};
	if (tfu) u = *tfu;
	int rc;
	if ((rc = ensure_scratch(vol, s))) return rc;
	auto *rows = reinterpret_cast<TFRowStats *>(vol->d_tf_rows);
	tf_masks_kernel<false><<<256, 256, 0, s>>>(reinterpret_cast<uchar4 *>(vol->d_tf), 0.0f, 0.0f, 0.0f, 0.0f, u, tfu ? 1 : 0, vol->d_mask2, rows,
	                                           reinterpret_cast<unsigned *>(rows + 256), vol->d_bounds);
	VKV_LAUNCHED();
	vol->mask_ana_valid = tfu != nullptr;
	if (tfu) vol->mask_tfu = *tfu;
	return VKV_OK;
}

}        // namespace vkv
