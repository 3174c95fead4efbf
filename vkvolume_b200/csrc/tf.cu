// tf.cu — transfer-function texture and the visibility masks derived from it.
//
// Replaces Volume::update_transfer_function_texture (src/volume_component.cpp:242-278):
// the reference fills the 256x256 RGBA8 texture in a CPU loop and uploads it through a
// staging buffer on every TF change; here one 65 536-thread kernel writes it in HBM with
// the same fp32 operation order (the library is compiled with -fmad=false, so no
// contraction) — bit-identical bytes, no host round trip.
#include "common.cuh"

namespace vkv {

__device__ __forceinline__ float clampf_dev(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// One thread per texel; texel (x = intensity, y = gradient).
__global__ void __launch_bounds__(256) tf_texture_kernel(uchar4 *__restrict__ tex, float imin, float imax, float gmin, float gmax)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= 256 * 256) return;
	const float i = (float) (idx & 255), g = (float) (idx >> 8);
	const float i_inv        = 1.0f / (imax - imin);
	const float g_inv        = 1.0f / (gmax - gmin);
	const bool  use_gradient = gmax != gmin;
	const float alpha_i      = clampf_dev(((i / 255.0f) - imin) * i_inv, 0.0f, 1.0f);
	const float alpha_g      = use_gradient ? clampf_dev(((g / 255.0f) - gmin) * g_inv, 0.0f, 1.0f) : 1.0f;
	// static_cast<uint8_t>(clamp(alpha_i * alpha_g * 255, 0, 255)): truncation
	const unsigned char a = (unsigned char) clampf_dev(alpha_i * alpha_g * 255.0f, 0.0f, 255.0f);
	tex[idx] = make_uchar4(a, a, a, a);
}

// Derives both bit masks and the conservative byte bounds.  One CTA of 1024 threads,
// two mask words per thread.  `have_tfu == 0` leaves the analytic mask empty.
__global__ void __launch_bounds__(1024) tf_masks_kernel(const uchar4 *__restrict__ tex, vkv_transfer_function_uniform tfu,
                                                         int have_tfu, uint2 *__restrict__ mask2, TFBounds *__restrict__ bounds)
{
	__shared__ unsigned s_vlo[4], s_vhi[4], s_glo[4], s_ghi[4];
	if (threadIdx.x < 4) {
		s_vlo[threadIdx.x] = 255u; s_vhi[threadIdx.x] = 0u;
		s_glo[threadIdx.x] = 255u; s_ghi[threadIdx.x] = 0u;
	}
	__syncthreads();
	for (int w = threadIdx.x; w < kMaskWords; w += blockDim.x) {
		const int g  = w >> 3;
		const int v0 = (w & 7) * 32;
		unsigned  mt = 0, ma = 0;
		// analytic TF (shaders/transfer_function.glsl:41-43) with g = 1.0 when gradients are unused
		// (shaders/get_gradient_compute.glsl:6-7); for row 255 float(255)/255 == 1.0 exactly.
		const float gradient = (float) g / 255.0f;
		const float aG       = clampf_dev((gradient - tfu.gradient_min) * tfu.gradient_range_inv, 0.0f, 1.0f);
		for (int b = 0; b < 32; ++b) {
			const int v = v0 + b;
			if (tex[g * 256 + v].w > 0) mt |= 1u << b;
			if (have_tfu) {
				const float intensity = (float) v / 255.0f;
				const float aI        = clampf_dev((intensity - tfu.intensity_min) * tfu.intensity_range_inv, 0.0f, 1.0f);
				if (aI * aG > 0.0f) ma |= 1u << b;
			}
		}
		mask2[w] = make_uint2(mt, ma);
		const unsigned sets[2] = {mt, mt | ma};
		for (int k = 0; k < 2; ++k) {
			const unsigned m = sets[k];
			if (!m) continue;
			const unsigned lo = v0 + (__ffs(m) - 1), hi = v0 + (31 - __clz(m));
			atomicMin(&s_vlo[k], lo); atomicMax(&s_vhi[k], hi);
			atomicMin(&s_glo[k], (unsigned) g); atomicMax(&s_ghi[k], (unsigned) g);
			if (g == 255) {
				atomicMin(&s_vlo[2 + k], lo); atomicMax(&s_vhi[2 + k], hi);
				atomicMin(&s_glo[2 + k], 255u); atomicMax(&s_ghi[2 + k], 255u);
			}
		}
	}
	__syncthreads();
	if (threadIdx.x < 4) {
		bounds->v_lo[threadIdx.x] = s_vlo[threadIdx.x]; bounds->v_hi[threadIdx.x] = s_vhi[threadIdx.x];
		bounds->g_lo[threadIdx.x] = s_glo[threadIdx.x]; bounds->g_hi[threadIdx.x] = s_ghi[threadIdx.x];
	}
}

int launch_tf_texture(vkv_volume *vol, const vkv_volume_options *opt, cudaStream_t s)
{
	tf_texture_kernel<<<256, 256, 0, s>>>(reinterpret_cast<uchar4 *>(vol->d_tf), opt->intensity_min, opt->intensity_max,
	                                       opt->gradient_min, opt->gradient_max);
	VKV_LAUNCHED();
	vol->has_tf = true;
	return VKV_OK;
}

int launch_tf_masks(vkv_volume *vol, const vkv_transfer_function_uniform *tfu, cudaStream_t s)
{
	vkv_transfer_function_uniform u{};
	if (tfu) u = *tfu;
	tf_masks_kernel<<<1, 1024, 0, s>>>(reinterpret_cast<const uchar4 *>(vol->d_tf), u, tfu ? 1 : 0, vol->d_mask2, vol->d_bounds);
	VKV_LAUNCHED();
	vol->mask_ana_valid = tfu != nullptr;
	if (tfu) vol->mask_tfu = *tfu;
	return VKV_OK;
}

}        // namespace vkv
