/*
 * vkv_oracle.c — CPU ORACLE for the VkVolume hot path.  TEST INFRASTRUCTURE ONLY
 * (see vkv_oracle.h for who may use it and for the parity status).
 *
 * Plain C restatement of the reference's shaders and host maths.  Citations are
 * file:line relative to the reference repository root.  Compile with
 * -O2 -ffp-contract=off -fopenmp (oracle/Makefile).
 */
#include "vkv_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* small helpers                                                               */
/* ------------------------------------------------------------------------- */
static inline float  clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline int    clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline uint32_t rnd_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

int orc_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#else
	(void) n;
#endif
}

/* column-major 4x4 in double: m[c*4 + r] */
static void m4_identity(double *m)
{
	memset(m, 0, 16 * sizeof(double));
	m[0] = m[5] = m[10] = m[15] = 1.0;
}
static void m4_mul(const double *a, const double *b, double *out)
{
	double t[16];
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 4; ++r) {
			double s = 0;
			for (int k = 0; k < 4; ++k) s += a[k * 4 + r] * b[c * 4 + k];
			t[c * 4 + r] = s;
		}
	memcpy(out, t, sizeof(t));
}
static void m4_mul_v(const double *m, const double *v, double *out)
{
	double t[4];
	for (int r = 0; r < 4; ++r) t[r] = m[0 + r] * v[0] + m[4 + r] * v[1] + m[8 + r] * v[2] + m[12 + r] * v[3];
	memcpy(out, t, sizeof(t));
}
static void m4_transpose(const double *m, double *out)
{
	double t[16];
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 4; ++r) t[r * 4 + c] = m[c * 4 + r];
	memcpy(out, t, sizeof(t));
}
/* general inverse by Gauss-Jordan with partial pivoting */
static int m4_inverse(const double *m, double *out)
{
	double a[4][8];
	for (int r = 0; r < 4; ++r)
		for (int c = 0; c < 4; ++c) {
			a[r][c]     = m[c * 4 + r];
			a[r][c + 4] = (r == c) ? 1.0 : 0.0;
		}
	for (int col = 0; col < 4; ++col) {
		int piv = col;
		for (int r = col + 1; r < 4; ++r)
			if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
		if (a[piv][col] == 0.0) return -1;
		if (piv != col)
			for (int c = 0; c < 8; ++c) {
				double t = a[col][c];
				a[col][c] = a[piv][c];
				a[piv][c] = t;
			}
		double inv = 1.0 / a[col][col];
		for (int c = 0; c < 8; ++c) a[col][c] *= inv;
		for (int r = 0; r < 4; ++r)
			if (r != col) {
				double f = a[r][col];
				if (f != 0.0)
					for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
			}
	}
	for (int r = 0; r < 4; ++r)
		for (int c = 0; c < 4; ++c) out[c * 4 + r] = a[r][c + 4];
	return 0;
}
static void m4_to_float(const double *m, float *out)
{
	for (int i = 0; i < 16; ++i) out[i] = (float) m[i];
}
static void m4_from_float(const float *m, double *out)
{
	for (int i = 0; i < 16; ++i) out[i] = (double) m[i];
}
/* glm::rotate(angle, axis) (glm/ext/matrix_transform.inl), used at src/load_volume.cpp:83 */
static void m4_rotate(double angle, const double axis_in[3], double *out)
{
	double c = cos(angle), s = sin(angle);
	double len = sqrt(axis_in[0] * axis_in[0] + axis_in[1] * axis_in[1] + axis_in[2] * axis_in[2]);
	double ax[3] = {axis_in[0] / len, axis_in[1] / len, axis_in[2] / len};
	double t[3]  = {(1 - c) * ax[0], (1 - c) * ax[1], (1 - c) * ax[2]};
	m4_identity(out);
	out[0]  = c + t[0] * ax[0];
	out[1]  = t[0] * ax[1] + s * ax[2];
	out[2]  = t[0] * ax[2] - s * ax[1];
	out[4]  = t[1] * ax[0] - s * ax[2];
	out[5]  = c + t[1] * ax[1];
	out[6]  = t[1] * ax[2] + s * ax[0];
	out[8]  = t[2] * ax[0] + s * ax[1];
	out[9]  = t[2] * ax[1] - s * ax[0];
	out[10] = c + t[2] * ax[2];
}
static void m4_scale(const double s[3], double *out)
{
	m4_identity(out);
	out[0] = s[0];
	out[5] = s[1];
	out[10] = s[2];
}
static void m4_translate(const double t[3], double *out)
{
	m4_identity(out);
	out[12] = t[0];
	out[13] = t[1];
	out[14] = t[2];
}
/* glm::mat4_cast(quat) (glm/gtc/quaternion.inl) */
static void m4_from_quat(const double q[4] /* x y z w */, double *out)
{
	double x = q[0], y = q[1], z = q[2], w = q[3];
	m4_identity(out);
	out[0]  = 1 - 2 * (y * y + z * z);
	out[1]  = 2 * (x * y + w * z);
	out[2]  = 2 * (x * z - w * y);
	out[4]  = 2 * (x * y - w * z);
	out[5]  = 1 - 2 * (x * x + z * z);
	out[6]  = 2 * (y * z + w * x);
	out[8]  = 2 * (x * z + w * y);
	out[9]  = 2 * (y * z - w * x);
	out[10] = 1 - 2 * (x * x + y * y);
}

/* ------------------------------------------------------------------------- */
/* a1: LoadVolume::load_header — src/load_volume.cpp:33-86                     */
/* ------------------------------------------------------------------------- */
static const char *next_line(const char *p, char *buf, size_t cap)
{
	size_t n = 0;
	while (*p && *p != '\n') {
		if (n + 1 < cap) buf[n++] = *p;
		++p;
	}
	buf[n] = 0;
	if (*p == '\n') ++p;
	return p;
}

int orc_parse_header(const char *text, vkv_volume_header *h)
{
	char        line[512];
	const char *p = text;
	memset(h, 0, sizeof(*h));
	/* ">>" extraction stops at the first token that does not parse, so a trailing
	 * "# comment" is simply never read (src/load_volume.cpp:56-79). */
	p = next_line(p, line, sizeof line);
	if (sscanf(line, "%u %u %u", &h->extent[0], &h->extent[1], &h->extent[2]) != 3) return -1;
	p = next_line(p, line, sizeof line);
	if (sscanf(line, "%f %f %f", &h->voxel_size[0], &h->voxel_size[1], &h->voxel_size[2]) != 3) return -1;
	p = next_line(p, line, sizeof line);
	if (sscanf(line, "%f %f", &h->normalisation_range[0], &h->normalisation_range[1]) != 2) return -1;
	p = next_line(p, line, sizeof line);
	if (sscanf(line, "%15s %15s", h->type, h->endianness) != 2) return -1;
	p = next_line(p, line, sizeof line);
	float aa[4];
	if (sscanf(line, "%f %f %f %f", &aa[0], &aa[1], &aa[2], &aa[3]) != 4) return -1;
	/* image_transform = rotate(radians(angle), axis) * scale(voxel_size * extent)  (:82-83) */
	float  phys[3];
	for (int i = 0; i < 3; ++i) phys[i] = h->voxel_size[i] * (float) h->extent[i];
	double R[16], S[16], M[16];
	double axis[3] = {aa[0], aa[1], aa[2]};
	double sc[3]   = {phys[0], phys[1], phys[2]};
	float  rad     = aa[3] * 0.01745329251994329576923690768489f; /* glm::radians */
	m4_rotate((double) rad, axis, R);
	m4_scale(sc, S);
	m4_mul(R, S, M);
	m4_to_float(M, h->image_transform);
	return 0;
}

/* ------------------------------------------------------------------------- */
/* a2: LoadVolume::load_data_impl<T> — src/load_volume.cpp:112-172             */
/* ------------------------------------------------------------------------- */
int orc_normalise(const void *raw, size_t n, const char *type, const char *endianness, float lo, float hi,
                  uint8_t *out)
{
	int big = strcmp(endianness, "big") == 0; /* anything else is treated as little (:151-162) */
	int kind;
	if (!strcmp(type, "uint8_t")) kind = 0;
	else if (!strcmp(type, "int8_t")) kind = 1;
	else if (!strcmp(type, "uint16_t")) kind = 2;
	else if (!strcmp(type, "int16_t")) kind = 3;
	else return -1; /* "unsupported image data type" (:106-109) */
	const uint8_t *b = (const uint8_t *) raw;
#pragma omp parallel for schedule(static)
	for (long long i = 0; i < (long long) n; ++i) {
		float v;
		switch (kind) {
			case 0: v = (float) b[i]; break;
			case 1: v = (float) (int8_t) b[i]; break;
			default: {
				uint16_t u = big ? (uint16_t) ((b[2 * i] << 8) | b[2 * i + 1]) : (uint16_t) ((b[2 * i + 1] << 8) | b[2 * i]);
				v          = kind == 2 ? (float) u : (float) (int16_t) u;
			}
		}
		/* static_cast<uint8_t>(255 * max(0, min(1, (float(v) - min) / (max - min))))  (:165-169) */
		float t = (v - lo) / (hi - lo);
		t       = fmaxf(0.0f, fminf(1.0f, t));
		out[i]  = (uint8_t) (255 * t);
	}
	return 0;
}

/* ------------------------------------------------------------------------- */
/* a4: Volume::get_transfer_function_uniform — src/volume_component.cpp:226-240 */
/* ------------------------------------------------------------------------- */
void orc_transfer_function_uniform(const vkv_volume_options *o, vkv_transfer_function_uniform *u)
{
	u->sampling_factor         = o->sampling_factor;
	u->voxel_alpha_factor      = o->voxel_alpha_factor;
	u->grad_magnitude_modifier = 1.0f;
	u->use_gradient            = o->gradient_max != o->gradient_min;
	u->intensity_min           = o->intensity_min;
	u->intensity_range_inv     = 1.0f / (o->intensity_max - o->intensity_min);
	u->gradient_min            = o->gradient_min;
	u->gradient_range_inv      = 1.0f / (o->gradient_max - o->gradient_min);
}

/* ------------------------------------------------------------------------- */
/* a5: Volume::update_transfer_function_texture — src/volume_component.cpp:242-261 */
/* ------------------------------------------------------------------------- */
void orc_transfer_function_texture(const vkv_volume_options *o, uint8_t *rgba)
{
	float  i_inv        = 1.0f / (o->intensity_max - o->intensity_min);
	float  g_inv        = 1.0f / (o->gradient_max - o->gradient_min);
	int    use_gradient = o->gradient_max != o->gradient_min;
	size_t idx          = 0;
	for (float g = 0; g < 256; ++g)
		for (float i = 0; i < 256; ++i, idx++) {
			float   alpha_i = clampf(((i / 255.0f) - o->intensity_min) * i_inv, 0.0f, 1.0f);
			float   alpha_g = use_gradient ? clampf(((g / 255.0f) - o->gradient_min) * g_inv, 0.0f, 1.0f) : 1.0f;
			uint8_t alpha   = (uint8_t) clampf(alpha_i * alpha_g * 255, 0, 255);
			rgba[idx * 4 + 0] = rgba[idx * 4 + 1] = rgba[idx * 4 + 2] = rgba[idx * 4 + 3] = alpha; /* glm::u8vec4(alpha) */
		}
}

/* ------------------------------------------------------------------------- */
/* K1: shaders/get_gradient_compute.glsl:5-23, shaders/gradient_map.comp:35-41 */
/* ------------------------------------------------------------------------- */
static inline float load_unorm(const uint8_t *V, const uint32_t dim[3], int x, int y, int z)
{
	/* imageLoad(image3D r8): UNORM decode = byte / 255 */
	return (float) V[((size_t) z * dim[1] + (size_t) y) * dim[0] + (size_t) x] / 255.0f;
}

/* on-the-fly tetrahedral gradient at integer voxel `pos` (the #else branch, :12-20) */
static float gradient_at(const uint8_t *V, const uint32_t dim[3], int x, int y, int z, float modifier)
{
	const int mx = (int) dim[0] - 1, my = (int) dim[1] - 1, mz = (int) dim[2] - 1;
	/* k = (1,-1): k.xyy = (+,-,-), k.yyx = (-,-,+), k.yxy = (-,+,-), k.xxx = (+,+,+) */
	float a = load_unorm(V, dim, clampi(x + 1, 0, mx), clampi(y - 1, 0, my), clampi(z - 1, 0, mz));
	float b = load_unorm(V, dim, clampi(x - 1, 0, mx), clampi(y - 1, 0, my), clampi(z + 1, 0, mz));
	float c = load_unorm(V, dim, clampi(x - 1, 0, mx), clampi(y + 1, 0, my), clampi(z - 1, 0, mz));
	float d = load_unorm(V, dim, clampi(x + 1, 0, mx), clampi(y + 1, 0, my), clampi(z + 1, 0, mz));
	/* 0.25 * (k.xyy*a + k.yyx*b + k.yxy*c + k.xxx*d), summed left to right */
	float gx = 0.25f * (((1.0f * a + -1.0f * b) + -1.0f * c) + 1.0f * d);
	float gy = 0.25f * (((-1.0f * a + -1.0f * b) + 1.0f * c) + 1.0f * d);
	float gz = 0.25f * (((-1.0f * a + 1.0f * b) + -1.0f * c) + 1.0f * d);
	float len = sqrtf((gx * gx + gy * gy) + gz * gz);
	return clampf(len * modifier, 0.0f, 1.0f);
}

static inline uint8_t store_unorm(float v)
{
	/* imageStore to r8: clamp, scale by 255, round to nearest (Vulkan spec float->UNORM) */
	return (uint8_t) rintf(clampf(v, 0.0f, 1.0f) * 255.0f);
}

void orc_gradient_map(const uint8_t *V, uint32_t W, uint32_t H, uint32_t D, int use_gradient, float modifier,
                      uint8_t *G, float *Gf)
{
	const uint32_t dim[3] = {W, H, D};
#pragma omp parallel for schedule(static)
	for (long long z = 0; z < (long long) D; ++z)
		for (uint32_t y = 0; y < H; ++y)
			for (uint32_t x = 0; x < W; ++x) {
				/* !use_gradient -> 1.0 (get_gradient_compute.glsl:6-7) */
				float  g   = use_gradient ? gradient_at(V, dim, (int) x, (int) y, (int) z, modifier) : 1.0f;
				size_t idx = ((size_t) z * H + y) * W + x;
				G[idx]     = store_unorm(g);
				if (Gf) Gf[idx] = g;
			}
}

/* ------------------------------------------------------------------------- */
/* A.2 block geometry: src/volume_component.cpp:91-93, compute_distance_map.cpp:108-113 */
/* ------------------------------------------------------------------------- */
void orc_map_extent(const uint32_t dim[3], uint32_t bs_requested, uint32_t dim_b[3], uint32_t bs_eff[3])
{
	for (int a = 0; a < 3; ++a) {
		dim_b[a]  = rnd_up(dim[a], bs_requested);
		bs_eff[a] = rnd_up(dim[a], dim_b[a]);
	}
}

/* texture(sampler2D, vec2(i,g)) with NEAREST / CLAMP_TO_EDGE on a 256x256 image:
 * texel = clamp(floor(coord * 256), 0, 255)  (Vulkan spec, nearest filtering) */
static inline int tf_texel(float coord)
{
	return clampi((int) floorf(coord * 256.0f), 0, 255);
}

/* ------------------------------------------------------------------------- */
/* K2a: shaders/occupancy_map.comp:45-73                                        */
/* ------------------------------------------------------------------------- */
void orc_occupancy_map(const uint8_t *V, const uint8_t *G, const uint8_t *tf, const uint32_t dim[3],
                       uint32_t bs_requested, int use_gradient, int precomputed, uint8_t *O)
{
	uint32_t dim_b[3], bs[3];
	orc_map_extent(dim, bs_requested, dim_b, bs);
#pragma omp parallel for schedule(dynamic, 1)
	for (long long bz = 0; bz < (long long) dim_b[2]; ++bz)
		for (uint32_t by = 0; by < dim_b[1]; ++by)
			for (uint32_t bx = 0; bx < dim_b[0]; ++bx) {
				uint32_t s[3] = {bx * bs[0], by * bs[1], (uint32_t) bz * bs[2]};
				uint32_t e[3];
				for (int a = 0; a < 3; ++a) e[a] = s[a] + bs[a] < dim[a] ? s[a] + bs[a] : dim[a];
				uint8_t result = 255; /* EMPTY */
				for (uint32_t z = s[2]; z < e[2] && result; ++z)
					for (uint32_t y = s[1]; y < e[1] && result; ++y)
						for (uint32_t x = s[0]; x < e[0]; ++x) {
							float intensity = load_unorm(V, dim, (int) x, (int) y, (int) z);
							float gradient;
							if (!use_gradient) gradient = 1.0f;
							else if (precomputed) gradient = load_unorm(G, dim, (int) x, (int) y, (int) z);
							else gradient = gradient_at(V, dim, (int) x, (int) y, (int) z, 1.0f);
							float alpha = (float) tf[((size_t) tf_texel(gradient) * 256 + tf_texel(intensity)) * 4 + 3] / 255.0f;
							if (alpha > 0.0f) {
								result = 0; /* OCCUPIED */
								break;
							}
						}
				O[((size_t) bz * dim_b[1] + by) * dim_b[0] + bx] = result;
			}
}

/* ------------------------------------------------------------------------- */
/* K2b: shaders/occupied_voxel_count.comp:28-56 with the analytic transfer      */
/* function of shaders/transfer_function.glsl:41-43                             */
/* ------------------------------------------------------------------------- */
static inline int voxel_counted(const uint8_t *V, const uint8_t *G, const uint32_t dim[3], int x, int y, int z,
                                const vkv_transfer_function_uniform *u, int precomputed)
{
	float intensity = load_unorm(V, dim, x, y, z);
	float gradient;
	if (!u->use_gradient) gradient = 1.0f;
	else if (precomputed) gradient = load_unorm(G, dim, x, y, z);
	else gradient = gradient_at(V, dim, x, y, z, u->grad_magnitude_modifier);
	float aI = clampf((intensity - u->intensity_min) * u->intensity_range_inv, 0.0f, 1.0f);
	float aG = clampf((gradient - u->gradient_min) * u->gradient_range_inv, 0.0f, 1.0f);
	return (aI * aG) > 0.0f;
}

uint64_t orc_occupied_voxel_count(const uint8_t *V, const uint8_t *G, const uint32_t dim[3],
                                  const vkv_transfer_function_uniform *u, int precomputed)
{
	uint64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
	for (long long z = 0; z < (long long) dim[2]; ++z)
		for (uint32_t y = 0; y < dim[1]; ++y)
			for (uint32_t x = 0; x < dim[0]; ++x) total += (uint64_t) voxel_counted(V, G, dim, (int) x, (int) y, (int) z, u, precomputed);
	return total;
}

/* The same, dispatched the way the reference does it: 8x8x8 workgroups, one partial sum per
 * subgroup (occupied_voxel_count.comp:43-55), then the strided in-place reduce of
 * occupied_voxel_count_reduce.comp:22-28 driven by compute_occupied_voxel_count.cpp:134-145. */
uint64_t orc_occupied_voxel_count_dispatch(const uint8_t *V, const uint8_t *G, const uint32_t dim[3],
                                           const vkv_transfer_function_uniform *u, int precomputed, uint32_t S)
{
	uint32_t wg[3]   = {rnd_up(dim[0], 8), rnd_up(dim[1], 8), rnd_up(dim[2], 8)};
	uint32_t n_sg    = 512 / S;
	size_t   n_elems = (size_t) wg[0] * wg[1] * wg[2] * n_sg;
	uint64_t *count  = (uint64_t *) calloc(n_elems, sizeof(uint64_t));
	if (!count) return (uint64_t) -1;
#pragma omp parallel for schedule(static)
	for (long long wz = 0; wz < (long long) wg[2]; ++wz)
		for (uint32_t wy = 0; wy < wg[1]; ++wy)
			for (uint32_t wx = 0; wx < wg[0]; ++wx) {
				size_t widx = ((size_t) wz * wg[1] + wy) * wg[0] + wx;
				for (uint32_t li = 0; li < 512; ++li) { /* gl_LocalInvocationIndex = z*64 + y*8 + x */
					uint32_t x = wx * 8 + (li & 7), y = wy * 8 + ((li >> 3) & 7), z = (uint32_t) wz * 8 + (li >> 6);
					if (x < dim[0] && y < dim[1] && z < dim[2])
						count[widx * n_sg + li / S] += (uint64_t) voxel_counted(V, G, dim, (int) x, (int) y, (int) z, u, precomputed);
				}
			}
	uint32_t stride = 1;
	while (stride < n_elems) {
		uint32_t groups = (uint32_t) ((n_elems + (uint64_t) S * stride - 1) / ((uint64_t) S * stride));
		for (uint32_t g = 0; g < groups; ++g) { /* one subgroup-sized workgroup each */
			uint64_t sum   = 0;
			uint32_t first = 0;
			for (uint32_t l = 0; l < S; ++l) {
				uint32_t sample = (g * S + l) * stride; /* uint arithmetic, as in the shader */
				if (l == 0) first = sample;
				if (sample < n_elems) sum += count[sample];
			}
			count[first] = sum;
		}
		stride *= S;
	}
	uint64_t r = count[0];
	free(count);
	return r;
}

/* ------------------------------------------------------------------------- */
/* K3a: shaders/distance_map.comp:44-109, dispatched by                        */
/* src/compute_distance_map.cpp:142-175 (stage 0 in place on the occupancy map) */
/* ------------------------------------------------------------------------- */
#define AT(buf, x, y, z) (buf)[((size_t) (z) * dy + (size_t) (y)) * dx + (size_t) (x)]

void orc_distance_map(const uint8_t *O, const uint32_t dim_b[3], uint8_t *dist)
{
	const int dx = (int) dim_b[0], dy = (int) dim_b[1], dz = (int) dim_b[2];
	size_t    M    = (size_t) dx * dy * dz;
	uint8_t  *swap = (uint8_t *) malloc(M);
	memcpy(dist, O, M); /* dist and dist_swap are the same image in stage 0 (:156-157) */
	/* stage 0 — "Transformation 1": two sweeps along x */
#pragma omp parallel for schedule(static)
	for (int z = 0; z < dz; ++z)
		for (int y = 0; y < dy; ++y) {
			uint32_t g1 = AT(dist, 0, y, z);
			for (int x = 1; x < dx; ++x) {
				uint32_t g = g1 + 1 < AT(dist, x, y, z) ? g1 + 1 : AT(dist, x, y, z);
				AT(dist, x, y, z) = (uint8_t) g;
				g1 = g;
			}
			for (int x = dx - 2; x >= 0; --x) {
				uint32_t g = g1 + 1 < AT(dist, x, y, z) ? g1 + 1 : AT(dist, x, y, z);
				AT(dist, x, y, z) = (uint8_t) g;
				g1 = g;
			}
		}
	/* stage 1 — "Transformation 2": dist -> swap along y */
#pragma omp parallel for schedule(static)
	for (int z = 0; z < dz; ++z)
		for (int x = 0; x < dx; ++x)
			for (int y = 0; y < dy; ++y) {
				uint32_t D = AT(dist, x, y, z);
				for (int n = 1; n < (int) D; ++n) {
					if (y >= n) {
						uint32_t Dn = AT(dist, x, y - n, z);
						uint32_t m  = (uint32_t) n > Dn ? (uint32_t) n : Dn;
						D           = D < m ? D : m;
					}
					if ((y + n) < dy && n < (int) D) {
						uint32_t Dn = AT(dist, x, y + n, z);
						uint32_t m  = (uint32_t) n > Dn ? (uint32_t) n : Dn;
						D           = D < m ? D : m;
					}
				}
				AT(swap, x, y, z) = (uint8_t) D;
			}
	/* stage 2 — "Transformation 3": swap -> dist along z */
#pragma omp parallel for schedule(static)
	for (int y = 0; y < dy; ++y)
		for (int x = 0; x < dx; ++x)
			for (int z = 0; z < dz; ++z) {
				uint32_t m_min = AT(swap, x, y, z);
				for (int n = 1; n < (int) m_min; ++n) {
					if (z >= n) {
						uint32_t g = AT(swap, x, y, z - n);
						uint32_t m = (uint32_t) n > g ? (uint32_t) n : g;
						m_min      = m_min < m ? m_min : m;
					}
					if ((z + n) < dz && n < (int) m_min) {
						uint32_t g = AT(swap, x, y, z + n);
						uint32_t m = (uint32_t) n > g ? (uint32_t) n : g;
						m_min      = m_min < m ? m_min : m;
					}
				}
				AT(dist, x, y, z) = (uint8_t) m_min;
			}
	free(swap);
}

/* ------------------------------------------------------------------------- */
/* K3b: shaders/distance_map_anisotropic.comp:31-92 in the 14-dispatch schedule */
/* of src/compute_distance_map.cpp:238-252                                      */
/* ------------------------------------------------------------------------- */
static void aniso_stage0(uint8_t *dist, const uint8_t *occ, int dx, int dy, int dz, int dir)
{
#pragma omp parallel for schedule(static)
	for (int z = 0; z < dz; ++z)
		for (int y = 0; y < dy; ++y) {
			int      start = dir > 0 ? dx - 1 : 0;
			int      end   = dir > 0 ? -1 : dx;
			uint32_t g1    = AT(occ, start, y, z);
			for (int x = start; x != end; x -= dir) {
				uint32_t o = AT(occ, x, y, z); /* read before write: safe when dist == occ (stage1(7,-1)) */
				uint32_t g = g1 + 1 < o ? g1 + 1 : o;
				AT(dist, x, y, z) = (uint8_t) g;
				g1 = g;
			}
		}
}
static void aniso_stage1(const uint8_t *dist, uint8_t *swap, int dx, int dy, int dz, int dir)
{
#pragma omp parallel for schedule(static)
	for (int z = 0; z < dz; ++z)
		for (int x = 0; x < dx; ++x)
			for (int y = 0; y < dy; ++y) {
				uint32_t m_min = AT(dist, x, y, z);
				for (int n = 1; n < (int) m_min && n < 255; ++n) {
					int yt = y + dir * n;
					if (yt < 0 || yt >= dy) break;
					uint32_t g = AT(dist, x, yt, z);
					uint32_t m = (uint32_t) n > g ? (uint32_t) n : g;
					if (m < m_min) m_min = m;
				}
				AT(swap, x, y, z) = (uint8_t) m_min;
			}
}
static void aniso_stage2(uint8_t *dist, const uint8_t *swap, int dx, int dy, int dz, int dir)
{
#pragma omp parallel for schedule(static)
	for (int y = 0; y < dy; ++y)
		for (int x = 0; x < dx; ++x)
			for (int z = 0; z < dz; ++z) {
				uint32_t m_min = AT(swap, x, y, z);
				for (int n = 1; n < (int) m_min && n < 255; ++n) {
					int zt = z + dir * n;
					if (zt < 0 || zt >= dz) break;
					uint32_t g = AT(swap, x, y, zt);
					uint32_t m = (uint32_t) n > g ? (uint32_t) n : g;
					if (m < m_min) m_min = m;
				}
				AT(dist, x, y, z) = (uint8_t) m_min;
			}
}

void orc_distance_map_anisotropic(const uint8_t *O, const uint32_t dim_b[3], uint8_t *D8)
{
	const int dx = (int) dim_b[0], dy = (int) dim_b[1], dz = (int) dim_b[2];
	size_t    M    = (size_t) dx * dy * dz;
	uint8_t  *swap = (uint8_t *) malloc(M);
#define MAP(i) (D8 + (size_t) (i) *M)
	memcpy(MAP(7), O, M); /* occupancy lives in map 7 (compute_distance_map.cpp:72,183) */
	aniso_stage0(MAP(3), MAP(7), dx, dy, dz, 1);
	aniso_stage1(MAP(3), swap, dx, dy, dz, 1);
	aniso_stage2(MAP(0), swap, dx, dy, dz, 1);
	aniso_stage2(MAP(1), swap, dx, dy, dz, -1);
	aniso_stage1(MAP(3), swap, dx, dy, dz, -1);
	aniso_stage2(MAP(2), swap, dx, dy, dz, 1);
	aniso_stage2(MAP(3), swap, dx, dy, dz, -1);
	aniso_stage0(MAP(7), MAP(7), dx, dy, dz, -1);
	aniso_stage1(MAP(7), swap, dx, dy, dz, 1);
	aniso_stage2(MAP(4), swap, dx, dy, dz, 1);
	aniso_stage2(MAP(5), swap, dx, dy, dz, -1);
	aniso_stage1(MAP(7), swap, dx, dy, dz, -1);
	aniso_stage2(MAP(6), swap, dx, dy, dz, 1);
	aniso_stage2(MAP(7), swap, dx, dy, dz, -1);
#undef MAP
	free(swap);
}

/* Closed form: D(p) = min(255, min over occupied q [in the octant] of the
 * Chebyshev distance).  Brute force; for small grids only. */
void orc_distance_map_closed_form(const uint8_t *O, const uint32_t dim_b[3], int octant, uint8_t *out)
{
	const int dx = (int) dim_b[0], dy = (int) dim_b[1], dz = (int) dim_b[2];
	int       sx = 0, sy = 0, sz = 0;
	if (octant >= 0) {
		sx = (octant & 4) ? -1 : 1;
		sy = (octant & 2) ? -1 : 1;
		sz = (octant & 1) ? -1 : 1;
	}
#pragma omp parallel for schedule(static)
	for (int z = 0; z < dz; ++z)
		for (int y = 0; y < dy; ++y)
			for (int x = 0; x < dx; ++x) {
				int best = 255;
				for (int qz = 0; qz < dz; ++qz)
					for (int qy = 0; qy < dy; ++qy)
						for (int qx = 0; qx < dx; ++qx) {
							if (AT(O, qx, qy, qz) != 0) continue;
							int ex = qx - x, ey = qy - y, ez = qz - z;
							if (octant >= 0 && (ex * sx < 0 || ey * sy < 0 || ez * sz < 0)) continue;
							int d = abs(ex);
							if (abs(ey) > d) d = abs(ey);
							if (abs(ez) > d) d = abs(ez);
							if (d < best) best = d;
						}
				AT(out, x, y, z) = (uint8_t) best;
			}
}

/* ------------------------------------------------------------------------- */
/* a13: host maths of VolumeRenderSubpass::draw — src/volume_render_subpass.cpp:219-249 */
/* ------------------------------------------------------------------------- */
void orc_make_uniforms(const uint32_t dim[3], const uint32_t dim_b[3], const vkv_camera_desc *cam,
                       const float image_transform[16], float clip_distance, vkv_camera_uniform *cu,
                       vkv_ray_cast_uniform *ru)
{
	double T[16], R[16], S[16], world[16], view[16], proj[16], vp[16], vp_inv[16];
	double t[3] = {cam->translation[0], cam->translation[1], cam->translation[2]};
	double q[4] = {cam->rotation[0], cam->rotation[1], cam->rotation[2], cam->rotation[3]};
	/* camera node world matrix = T * R * S (VS/framework/scene_graph/components/transform.cpp:92-97), S = 1 */
	m4_translate(t, T);
	m4_from_quat(q, R);
	m4_mul(T, R, world);
	m4_inverse(world, view); /* Camera::get_view */
	/* glm::perspective(fov, aspect, zfar, znear) = perspectiveRH_ZO with near/far swapped
	 * (VS/framework/scene_graph/components/perspective_camera.cpp:72-76), then proj[1][1] *= -1
	 * (VS/framework/rendering/subpass.cpp:29-36) */
	{
		float zNear = cam->zfar, zFar = cam->znear; /* swapped on purpose: reverse-Z */
		float tanHalf = tanf(cam->yfov / 2.0f);
		memset(proj, 0, sizeof proj);
		proj[0]  = 1.0f / (cam->aspect * tanHalf);
		proj[5]  = -(1.0f / tanHalf);
		proj[10] = zFar / (zNear - zFar);
		proj[11] = -1.0;
		proj[14] = -(zFar * zNear) / (zFar - zNear);
	}
	m4_mul(proj, view, vp);
	m4_inverse(vp, vp_inv);
	/* model = node.get_matrix() * image_transform */
	double nT[16], nR[16], nS[16], node[16], img[16], model[16], model_inv[16];
	double nt[3] = {cam->node_translation[0], cam->node_translation[1], cam->node_translation[2]};
	double nq[4] = {cam->node_rotation[0], cam->node_rotation[1], cam->node_rotation[2], cam->node_rotation[3]};
	double ns[3] = {cam->node_scale[0], cam->node_scale[1], cam->node_scale[2]};
	m4_translate(nt, nT);
	m4_from_quat(nq, nR);
	m4_scale(ns, nS);
	m4_mul(nT, nR, node);
	m4_mul(node, nS, node);
	m4_from_float(image_transform, img);
	m4_mul(node, img, model);
	m4_inverse(model, model_inv);
	(void) S;
	m4_to_float(view, cu->view);
	m4_to_float(proj, cu->proj);
	m4_to_float(vp_inv, cu->view_proj_inv);
	m4_to_float(model, cu->model);
	m4_to_float(model_inv, cu->model_inv);

	double half[3] = {0.5, 0.5, 0.5}, model_to_tex[16], global_to_tex[16], view_inv[16];
	m4_translate(half, model_to_tex);
	m4_mul(model_to_tex, model_inv, global_to_tex);
	m4_inverse(view, view_inv);
	double cam_pos[4] = {view_inv[12], view_inv[13], view_inv[14], 1.0};
	double cam_model[4], cam_tex[4];
	m4_mul_v(model_inv, cam_pos, cam_model);
	cam_model[3] = 1.0;
	m4_mul_v(model_to_tex, cam_model, cam_tex);
	double fwd[4] = {0, 0, -1, 0}, dir[4];
	m4_mul_v(view_inv, fwd, dir);
	double plane[4] = {dir[0], dir[1], dir[2],
	                   -(double) clip_distance - (cam_pos[0] * dir[0] + cam_pos[1] * dir[1] + cam_pos[2] * dir[2])};
	double g2t_inv[16], g2t_invT[16], plane_tex[4];
	m4_inverse(global_to_tex, g2t_inv);
	m4_transpose(g2t_inv, g2t_invT);
	m4_mul_v(g2t_invT, plane, plane_tex);
	for (int i = 0; i < 4; ++i) {
		ru->plane[i]       = (float) plane[i];
		ru->plane_tex[i]   = (float) plane_tex[i];
		ru->cam_pos_tex[i] = (float) cam_tex[i];
	}
	ru->front_index = (ru->plane_tex[0] < 0 ? 1 : 0) + (ru->plane_tex[1] < 0 ? 2 : 0) + (ru->plane_tex[2] < 0 ? 4 : 0);
	for (int a = 0; a < 3; ++a) ru->block_size[a] = (float) rnd_up(dim[a], dim_b[a]);
	ru->block_size[3] = 0;
	ru->_pad[0] = ru->_pad[1] = ru->_pad[2] = 0;
}

/* ------------------------------------------------------------------------- */
/* K4: shaders/volume_render.frag + analytic ray entry + store     */
/* ------------------------------------------------------------------------- */
typedef struct {
	const uint8_t *V, *G, *tf, *Dm;
	uint32_t       dim[3], dim_b[3];
	size_t         M;
	const vkv_camera_uniform            *cam;
	const vkv_ray_cast_uniform          *ray;
	const vkv_transfer_function_uniform *tfu;
	const vkv_render_options            *opt;
	int                                  precomputed;
} frag_ctx;

/* texture(sampler3D, pos): LINEAR / CLAMP_TO_EDGE, unnormalised = pos*size - 0.5, UNORM texels */
static float sample_trilinear(const uint8_t *T, const uint32_t dim[3], float px, float py, float pz)
{
	float u = px * (float) dim[0] - 0.5f, v = py * (float) dim[1] - 0.5f, w = pz * (float) dim[2] - 0.5f;
	float fu = floorf(u), fv = floorf(v), fw = floorf(w);
	float a = u - fu, b = v - fv, c = w - fw;
	int   x0 = (int) fu, y0 = (int) fv, z0 = (int) fw;
	int   mx = (int) dim[0] - 1, my = (int) dim[1] - 1, mz = (int) dim[2] - 1;
	int   x1 = clampi(x0 + 1, 0, mx), y1 = clampi(y0 + 1, 0, my), z1 = clampi(z0 + 1, 0, mz);
	x0 = clampi(x0, 0, mx);
	y0 = clampi(y0, 0, my);
	z0 = clampi(z0, 0, mz);
	float t000 = load_unorm(T, dim, x0, y0, z0), t100 = load_unorm(T, dim, x1, y0, z0);
	float t010 = load_unorm(T, dim, x0, y1, z0), t110 = load_unorm(T, dim, x1, y1, z0);
	float t001 = load_unorm(T, dim, x0, y0, z1), t101 = load_unorm(T, dim, x1, y0, z1);
	float t011 = load_unorm(T, dim, x0, y1, z1), t111 = load_unorm(T, dim, x1, y1, z1);
	float c00 = t000 * (1.0f - a) + t100 * a, c10 = t010 * (1.0f - a) + t110 * a;
	float c01 = t001 * (1.0f - a) + t101 * a, c11 = t011 * (1.0f - a) + t111 * a;
	float c0 = c00 * (1.0f - b) + c10 * b, c1 = c01 * (1.0f - b) + c11 * b;
	return c0 * (1.0f - c) + c1 * c;
}

/* get_gradient of the fragment shader (volume_render.frag:86-105) */
static float frag_gradient(const frag_ctx *c, const float pos[3], const float dim_inv[3])
{
	if (!c->tfu->use_gradient) return 1.0f;
	if (c->precomputed) return sample_trilinear(c->G, c->dim, pos[0], pos[1], pos[2]);
	float a = sample_trilinear(c->V, c->dim, pos[0] + dim_inv[0] * 1.0f, pos[1] + dim_inv[1] * -1.0f, pos[2] + dim_inv[2] * -1.0f);
	float b = sample_trilinear(c->V, c->dim, pos[0] + dim_inv[0] * -1.0f, pos[1] + dim_inv[1] * -1.0f, pos[2] + dim_inv[2] * 1.0f);
	float d = sample_trilinear(c->V, c->dim, pos[0] + dim_inv[0] * -1.0f, pos[1] + dim_inv[1] * 1.0f, pos[2] + dim_inv[2] * -1.0f);
	float e = sample_trilinear(c->V, c->dim, pos[0] + dim_inv[0] * 1.0f, pos[1] + dim_inv[1] * 1.0f, pos[2] + dim_inv[2] * 1.0f);
	float gx = (((1.0f * a + -1.0f * b) + -1.0f * d) + 1.0f * e) * 0.25f;
	float gy = (((-1.0f * a + -1.0f * b) + 1.0f * d) + 1.0f * e) * 0.25f;
	float gz = (((-1.0f * a + 1.0f * b) + -1.0f * d) + 1.0f * e) * 0.25f;
	return clampf(sqrtf((gx * gx + gy * gy) + gz * gz) * c->tfu->grad_magnitude_modifier, 0.0f, 1.0f);
}

static inline float glsl_step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
static inline float glsl_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

/* mat4 * vec4 in fp32, summed left to right as a GLSL compiler without contraction does */
static void m4f_mul_v(const float *m, const float v[4], float out[4])
{
	for (int r = 0; r < 4; ++r) out[r] = ((m[0 + r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r] * v[3];
}

/* main() of shaders/volume_render.frag:117-336 for one fragment whose interpolated varyings are `entry` (ray_entry) and
 * `position` (clip-space gl_Position of the entry point; read only under DEPTH_ATTACHMENT, where `depth_in` is what
 * subpassLoad(i_depth) returns).  `out` is the shader's out_color (premultiplied, before blending); returns the gl_FragDepth
 * value; *discarded is set when the shader executes `discard` (:133). */
static float frag_main(const frag_ctx *c, const float entry[3], const float position[4], float depth_in, float out[4],
                       uint32_t n_samples[3], int *discarded)
{
	const vkv_transfer_function_uniform *tfu = c->tfu;
	const vkv_render_options            *opt = c->opt;
	out[0] = out[1] = out[2] = out[3] = 0.0f;
	float frag_depth = 0.0f; /* REVERSE_DEPTH, no depth attachment (:139-141) */
	n_samples[0] = n_samples[1] = n_samples[2] = 0;
	*discarded = 0;
	float frag_depth_front = 0.0f;
	if (opt->depth_attachment) { /* :122-136 */
		frag_depth_front = position[2] / position[3];
		if (depth_in > frag_depth_front) { /* REVERSE_DEPTH: the front face is behind the scene */
			*discarded = 1;
			return frag_depth;
		}
		frag_depth = depth_in;
	}

	/* ray exit (:146-149) */
	float dv[3] = {entry[0] - c->ray->cam_pos_tex[0], entry[1] - c->ray->cam_pos_tex[1], entry[2] - c->ray->cam_pos_tex[2]};
	float dl    = sqrtf((dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2]);
	float dir[3] = {dv[0] / dl, dv[1] / dl, dv[2] / dl};
	float t_far;
	{
		float t2[3];
		for (int a = 0; a < 3; ++a) {
			float inv = 1.0f / dir[a];
			float t_min = -entry[a] * inv, t_max = (1.0f - entry[a]) * inv;
			t2[a] = fmaxf(t_min, t_max);
		}
		t_far = fminf(fminf(t2[0], t2[1]), t2[2]);
	}
	float ray_exit[3] = {t_far * dir[0] + entry[0], t_far * dir[1] + entry[1], t_far * dir[2] + entry[2]};
	float ev[3]       = {entry[0] - ray_exit[0], entry[1] - ray_exit[1], entry[2] - ray_exit[2]};
	float ray_distance = sqrtf((ev[0] * ev[0] + ev[1] * ev[1]) + ev[2] * ev[2]);

	if (opt->depth_attachment) { /* :151-165: stop the ray where it meets the depth buffer */
		float clip_at_depth[4] = {position[0] * depth_in / frag_depth_front, position[1] * depth_in / frag_depth_front,
		                          position[2] * depth_in / frag_depth_front, position[3]};
		float pos_at_depth[4], pm[4];
		m4f_mul_v(c->cam->view_proj_inv, clip_at_depth, pos_at_depth);
		const float w = pos_at_depth[3];
		for (int a = 0; a < 4; ++a) pos_at_depth[a] = pos_at_depth[a] / w;
		m4f_mul_v(c->cam->model_inv, pos_at_depth, pm);
		float hit[3] = {pm[0] + 0.5f, pm[1] + 0.5f, pm[2] + 0.5f};
		float hv[3]  = {entry[0] - hit[0], entry[1] - hit[1], entry[2] - hit[2]};
		float hd     = sqrtf((hv[0] * hv[0] + hv[1] * hv[1]) + hv[2] * hv[2]);
		if (hd < ray_distance) {
			ray_exit[0] = hit[0]; ray_exit[1] = hit[1]; ray_exit[2] = hit[2];
			ray_distance = hd;
		}
	}

	if (opt->test == VKV_TEST_RAY_ENTRY) { /* :168-170 */
		out[0] = entry[0]; out[1] = entry[1]; out[2] = entry[2]; out[3] = 1.0f;
		return frag_depth;
	}
	if (opt->test == VKV_TEST_RAY_EXIT) { /* :171-173 */
		out[0] = ray_exit[0]; out[1] = ray_exit[1]; out[2] = ray_exit[2]; out[3] = 1.0f;
		return frag_depth;
	}

	/* number of samples (:175-180) */
	const float dimf[3] = {(float) c->dim[0], (float) c->dim[1], (float) c->dim[2]};
	int   dim_max = (int) c->dim[0];
	if ((int) c->dim[1] > dim_max) dim_max = (int) c->dim[1];
	if ((int) c->dim[2] > dim_max) dim_max = (int) c->dim[2];
	int   n_steps = (int) ceilf((float) dim_max * ray_distance * tfu->sampling_factor);
	float step[3];
	for (int a = 0; a < 3; ++a) step[a] = dir[a] * ray_distance / ((float) n_steps - 1.0f);
	float sampling_factor_inv = 1.0f / tfu->sampling_factor;

	/* early exit (:184-187) */
	for (int a = 0; a < 3; ++a) {
		float t = entry[a] + step[a];
		if (t <= 0.0f || t >= 1.0f) return frag_depth;
		if (t != t) return frag_depth; /* NaN: neither comparison holds in GLSL either; treat as outside */
	}

	const int skip_on = opt->skipping_type != VKV_SKIP_NONE;
	float vol_to_map[3], sdt_inv[3];
	int   dim_map_1[3];
	for (int a = 0; a < 3; ++a) {
		vol_to_map[a] = dimf[a] / c->ray->block_size[a];
		dim_map_1[a]  = (int) c->dim_b[a] - 1;
		float sdt     = step[a] * dimf[a] / c->ray->block_size[a];
		sdt_inv[a]    = 1.0f / sdt;
	}
	int i_min = 0;
	int u_last[3] = {0, 0, 0};
	float dim_inv[3] = {1.0f / dimf[0], 1.0f / dimf[1], 1.0f / dimf[2]};
	const uint8_t *Dm = c->Dm;
	if (opt->skipping_type == VKV_SKIP_ANISOTROPIC_DISTANCE) {
		int idx = (dir[2] < 0 ? 1 : 0) + (dir[1] < 0 ? 2 : 0) + (dir[0] < 0 ? 4 : 0); /* :209 */
		Dm      = c->Dm + (size_t) idx * c->M;
	}

	int voxel_occupied = 1;
	int i_first_hit    = n_steps;
	for (int i = 0; i < n_steps;) {
		float pos[3] = {entry[0] + (float) i * step[0], entry[1] + (float) i * step[1], entry[2] + (float) i * step[2]};
		float u[3];
		int   u_i[3] = {0, 0, 0};
		int   do_skip = 0;
		if (skip_on) {
			for (int a = 0; a < 3; ++a) {
				u[a]   = vol_to_map[a] * pos[a];
				u_i[a] = clampi((int) u[a], 0, dim_map_1[a]);
			}
			do_skip = !voxel_occupied && (u_i[0] != u_last[0] || u_i[1] != u_last[1] || u_i[2] != u_last[2]);
		}
		if (do_skip) {
			n_samples[1]++;
			uint32_t dist = Dm[((size_t) u_i[2] * c->dim_b[1] + (size_t) u_i[1]) * c->dim_b[0] + (size_t) u_i[0]];
			if (dist > 0u) {
				float dxyz[3];
				for (int a = 0; a < 3; ++a) {
					float r = clampf((float) u_i[a] - u[a], -1.0f, 0.0f);
					if (opt->skipping_type == VKV_SKIP_BLOCK)
						dxyz[a] = (glsl_step(0.0f, sdt_inv[a]) + r) * sdt_inv[a]; /* :237 */
					else
						dxyz[a] = (glsl_step(0.0f, -sdt_inv[a]) + glsl_sign(sdt_inv[a]) * (float) dist + r) * sdt_inv[a]; /* :240 */
				}
				float m       = fminf(fminf(dxyz[0], dxyz[1]), dxyz[2]);
				int   i_delta = (int) ceilf(m);
				if (i_delta < 1) i_delta = 1;
				i += i_delta;
			} else {
				int i_delta    = -(int) ceilf(tfu->sampling_factor);
				voxel_occupied = 1;
				u_last[0] = u_i[0]; u_last[1] = u_i[1]; u_last[2] = u_i[2];
				i = i + i_delta > i_min ? i + i_delta : i_min;
			}
		} else {
			n_samples[0]++;
			float intensity = sample_trilinear(c->V, c->dim, pos[0], pos[1], pos[2]);
			float gradient  = frag_gradient(c, pos, dim_inv);
			const uint8_t *tx = c->tf + ((size_t) tf_texel(gradient) * 256 + (size_t) tf_texel(intensity)) * 4;
			float color[4] = {(float) tx[0] / 255.0f, (float) tx[1] / 255.0f, (float) tx[2] / 255.0f, (float) tx[3] / 255.0f};
			voxel_occupied = color[3] > 0.0f;
			if (voxel_occupied) {
				if (skip_on) { u_last[0] = u_i[0]; u_last[1] = u_i[1]; u_last[2] = u_i[2]; }
				color[3] = clampf(tfu->voxel_alpha_factor * (1.0f - powf(1.0f - color[3], sampling_factor_inv)), 0.0f, 1.0f);
				color[0] *= color[3]; color[1] *= color[3]; color[2] *= color[3];
				float w = 1.0f - out[3];
				out[0] = out[0] + w * color[0]; out[1] = out[1] + w * color[1];
				out[2] = out[2] + w * color[2]; out[3] = out[3] + w * color[3];
				if (color[3] > 0.0f) i_first_hit = i;
				if (out[3] > 0.99f && opt->early_ray_termination) {
					out[3] = 1.0f;
					break;
				}
			} else {
				n_samples[2]++;
			}
			++i;
			if (skip_on) i_min = i;
		}
	}

	/* depth (:314-321) */
	if (out[3] > 0.0f && i_first_hit < n_steps) {
		double pm[4] = {(double) (entry[0] + step[0] * (float) i_first_hit) - 0.5, (double) (entry[1] + step[1] * (float) i_first_hit) - 0.5,
		                (double) (entry[2] + step[2] * (float) i_first_hit) - 0.5, 1.0};
		double model[16], view[16], proj[16], w4[4], v4[4], p4[4];
		m4_from_float(c->cam->model, model);
		m4_from_float(c->cam->view, view);
		m4_from_float(c->cam->proj, proj);
		m4_mul_v(model, pm, w4);
		m4_mul_v(view, w4, v4);
		m4_mul_v(proj, v4, p4);
		frag_depth = (float) (p4[2] / p4[3]);
	}

	if (opt->test == VKV_TEST_NUM_TEXTURE_SAMPLES) { /* :323-335 */
		uint32_t n_max = (uint32_t) (ceilf((float) dim_max * sqrtf(3.0f)) * tfu->sampling_factor);
		float    s     = (float) (n_samples[0] + n_samples[1]) / (float) n_max;
		out[0] = out[1] = out[2] = s;
		out[3] = 1.0f;
	}
	return frag_depth;
}

static inline float srgb_encode(float c)
{
	return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}
static inline uint8_t unorm8(float c)
{
	return (uint8_t) (clampf(c, 0.0f, 1.0f) * 255.0f + 0.5f);
}
/* R8G8B8A8_SRGB load (Vulkan spec, "sRGB EOTF"): what the blender reads back from the attachment */
static inline float srgb_decode(uint8_t b)
{
	float c = (float) b / 255.0f;
	return c <= 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
}

/* Analytic replacement of both vertex shaders + rasteriser: returns 1 and the
 * interpolated ray_entry varying if the pixel is covered. */
static int pixel_entry(const frag_ctx *c, const double vp_inv[16], const double model_inv[16], int px, int py, int W,
                       int H, float entry[3])
{
	double ndc[4] = {2.0 * (px + 0.5) / W - 1.0, 2.0 * (py + 0.5) / H - 1.0, 0.0 /* far plane in reverse-Z */, 1.0};
	double wp[4], mp[4];
	m4_mul_v(vp_inv, ndc, wp);
	for (int a = 0; a < 3; ++a) wp[a] /= wp[3];
	wp[3] = 1.0;
	m4_mul_v(model_inv, wp, mp);
	double o[3] = {c->ray->cam_pos_tex[0], c->ray->cam_pos_tex[1], c->ray->cam_pos_tex[2]};
	double d[3] = {mp[0] + 0.5 - o[0], mp[1] + 0.5 - o[1], mp[2] + 0.5 - o[2]};
	double tn = -INFINITY, tf = INFINITY;
	for (int a = 0; a < 3; ++a) {
		if (d[a] == 0.0) {
			if (o[a] < 0.0 || o[a] > 1.0) return 0;
			continue;
		}
		double t0 = (0.0 - o[a]) / d[a], t1 = (1.0 - o[a]) / d[a];
		if (t0 > t1) { double t = t0; t0 = t1; t1 = t; }
		if (t0 > tn) tn = t0;
		if (t1 < tf) tf = t1;
	}
	/* clip plane: keep pi . p >= 0 (volume_render_clipped.vert:56) */
	const float *pl = c->ray->plane_tex;
	double s0 = pl[0] * o[0] + pl[1] * o[1] + pl[2] * o[2] + pl[3];
	double sd = pl[0] * d[0] + pl[1] * d[1] + pl[2] * d[2];
	if (!(sd > 0.0)) return 0;
	double t_clip = -s0 / sd;
	double t0     = tn > t_clip ? tn : t_clip;
	if (t0 < 0.0) t0 = 0.0;
	if (!(t0 < tf)) return 0;
	for (int a = 0; a < 3; ++a) entry[a] = (float) (o[a] + t0 * d[a]);
	return 1;
}

void orc_render(const uint8_t *V, const uint8_t *G, const uint8_t *tf, const uint8_t *Dm, const uint32_t dim[3],
                const uint32_t dim_b[3], const vkv_camera_uniform *cam, const vkv_ray_cast_uniform *ray,
                const vkv_transfer_function_uniform *tfu, const vkv_render_options *opt, int precomputed, int width,
                int height, int y_first, int y_count, uint8_t *rgba8, float *rgba_f, float *depth,
                vkv_sample_counts *counts)
{
	frag_ctx c;
	c.V = V; c.G = G; c.tf = tf; c.Dm = Dm;
	memcpy(c.dim, dim, sizeof c.dim);
	memcpy(c.dim_b, dim_b, sizeof c.dim_b);
	c.M = (size_t) dim_b[0] * dim_b[1] * dim_b[2];
	c.cam = cam; c.ray = ray; c.tfu = tfu; c.opt = opt; c.precomputed = precomputed;
	double vp_inv[16], model_inv[16];
	m4_from_float(cam->view_proj_inv, vp_inv);
	m4_from_float(cam->model_inv, model_inv);
	double pvm[16];
	{
		double pr[16], vw[16], md[16], pv[16];
		m4_from_float(cam->proj, pr);
		m4_from_float(cam->view, vw);
		m4_from_float(cam->model, md);
		m4_mul(pr, vw, pv);
		m4_mul(pv, md, pvm);
	}
	const int load = opt->load_framebuffer != 0;
	uint64_t nv = 0, nd = 0, ne = 0, ncov = 0;
	if (y_count < 0) { y_first = 0; y_count = height; }
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nv, nd, ne, ncov)
	for (int py = y_first; py < y_first + y_count; ++py)
		for (int px = 0; px < width; ++px) {
			size_t   p = (size_t) py * width + px;
			float    entry[3], out[4] = {0, 0, 0, 0}, fd = 0.0f;
			uint32_t ns[3] = {0, 0, 0};
			int      covered = pixel_entry(&c, vp_inv, model_inv, px, py, width, height, entry);
			int      discarded = 0, write = 0;
			/* destination: the render-pass clear (0,0,0,1) / depth 0 (render_pipeline.cpp:38-39), or what the target holds */
			float dst[4] = {0.0f, 0.0f, 0.0f, 1.0f}, dst_depth = 0.0f;
			if (load) {
				if (rgba8) {
					dst[0] = srgb_decode(rgba8[p * 4 + 0]); dst[1] = srgb_decode(rgba8[p * 4 + 1]); dst[2] = srgb_decode(rgba8[p * 4 + 2]);
					dst[3] = (float) rgba8[p * 4 + 3] / 255.0f;
				}
				if (depth) dst_depth = depth[p];
			}
			float r = dst[0], g = dst[1], b = dst[2], a = dst[3];
			if (covered) {
				/* the `position` varying: gl_Position of the entry point = proj * view * model * (ray_entry - 0.5)
				 * (volume_render_clipped.vert:58-62, volume_render_plane_intersection.vert:128) */
				float position[4] = {0.0f, 0.0f, 0.5f, 1.0f};
				if (opt->depth_attachment) {
					double pm[4] = {(double) entry[0] - 0.5, (double) entry[1] - 0.5, (double) entry[2] - 0.5, 1.0}, p4[4];
					m4_mul_v(pvm, pm, p4);
					for (int k = 0; k < 4; ++k) position[k] = (float) p4[k];
				}
				fd = frag_main(&c, entry, position, dst_depth, out, ns, &discarded);
				ncov++;
				/* depth test GREATER_OR_EQUAL with depth write, then blend: rgb = src.rgb + dst.rgb * (1 - src.a),
				 * a = src.a * (1 - src.a) + dst.a * 0  (src/volume_render_subpass.cpp:176-190 + pipeline_state.h:91-125) */
				if (!discarded && fd >= dst_depth) {
					const float sa = clampf(out[3], 0.0f, 1.0f);
					r = clampf(out[0], 0.0f, 1.0f) + dst[0] * (1.0f - sa);
					g = clampf(out[1], 0.0f, 1.0f) + dst[1] * (1.0f - sa);
					b = clampf(out[2], 0.0f, 1.0f) + dst[2] * (1.0f - sa);
					a = sa * (1.0f - sa);
					write = 1;
				}
			}
			nv += ns[0]; nd += ns[1]; ne += ns[2];
			if (rgba8 && (write || !load)) {
				/* R8G8B8A8_SRGB store (render_context.cpp:22): sRGB-encode RGB, alpha linear */
				rgba8[p * 4 + 0] = unorm8(srgb_encode(clampf(r, 0.0f, 1.0f)));
				rgba8[p * 4 + 1] = unorm8(srgb_encode(clampf(g, 0.0f, 1.0f)));
				rgba8[p * 4 + 2] = unorm8(srgb_encode(clampf(b, 0.0f, 1.0f)));
				rgba8[p * 4 + 3] = unorm8(a);
			}
			if (depth && (write || !load)) depth[p] = write ? fd : 0.0f;
			if (rgba_f) { rgba_f[p * 4 + 0] = out[0]; rgba_f[p * 4 + 1] = out[1]; rgba_f[p * 4 + 2] = out[2]; rgba_f[p * 4 + 3] = covered ? (discarded ? -2.0f : out[3]) : -1.0f; }
		}
	if (counts) {
		counts->volume_samples += nv;
		counts->distance_samples += nd;
		counts->empty_samples += ne;
		counts->covered_pixels += ncov;
	}
}

/* ------------------------------------------------------------------------- */
/* Synthetic inputs (NOT reference behaviour): CPU twin of vkv_synth_volume    */
/* (vkvolume_b200/csrc/synth.cu) so the reference arm of bench.py can build the */
/* same workload without touching the GPU.                                      */
/* ------------------------------------------------------------------------- */
typedef struct { float cx, cy, cz, ex, ey, ez, r, amp; } synth_prim;

static uint64_t splitmix64(uint64_t *x)
{
	uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
static inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
	uint32_t h = x * 0x8da6b343u ^ y * 0xd8163841u ^ z * 0xcb1ab31fu ^ seed;
	h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
	return h;
}
static inline float capsule_dist2(float px, float py, float pz, const synth_prim *c)
{
	float bx = c->ex - c->cx, by = c->ey - c->cy, bz = c->ez - c->cz;
	float ax = px - c->cx, ay = py - c->cy, az = pz - c->cz;
	float t  = (ax * bx + ay * by + az * bz) / fmaxf(bx * bx + by * by + bz * bz, 1e-12f);
	t        = fminf(fmaxf(t, 0.0f), 1.0f);
	float dx = ax - t * bx, dy = ay - t * by, dz = az - t * bz;
	return dx * dx + dy * dy + dz * dz;
}

int orc_synth_volume(int kind, uint64_t seed, uint32_t W, uint32_t H, uint32_t D, uint8_t *out)
{
	if (kind < 0 || kind > 3) return -1;
	synth_prim prims[64];
	int        n_prims;
	uint32_t   m = W > H ? W : H;
	if (D > m) m = D;
	float mx = (float) m, inv_max = 1.0f / mx;
	float ext[3] = {W / mx, H / mx, D / mx};
	uint64_t st = seed;
#define RND() ((float) ((splitmix64(&st) >> 40) * (1.0 / 16777216.0)))
	uint32_t seed_lo = (uint32_t) splitmix64(&st);
	uint32_t seed_hi = (uint32_t) splitmix64(&st);
	memset(prims, 0, sizeof prims);
	if (kind == 0 || kind == 3) {
		n_prims = 64;
		for (int i = 0; i < 64; ++i) {
			synth_prim *b = &prims[i];
			b->cx = RND() * ext[0]; b->cy = RND() * ext[1]; b->cz = RND() * ext[2];
			b->r   = kind == 0 ? (6.0f + 18.0f * RND()) / 256.0f : (0.006f + 0.02f * RND());
			b->amp = 64.0f + 191.0f * RND();
		}
	} else if (kind == 1) {
		n_prims = 7;
		synth_prim *e = &prims[0];
		e->cx = 0.5f * ext[0]; e->cy = 0.5f * ext[1]; e->cz = 0.5f * ext[2];
		e->ex = 0.30f * ext[0]; e->ey = 0.22f * ext[1]; e->ez = 0.36f * ext[2];
		e->r = 0.02f; e->amp = 200.0f;
		for (int i = 1; i < 7; ++i) {
			synth_prim *c = &prims[i];
			float side = (i & 1) ? 1.0f : -1.0f, along = ((i - 1) / 2 - 1) * 0.18f;
			c->cx = e->cx + side * 0.25f * ext[0]; c->cy = e->cy + 0.1f * ext[1]; c->cz = e->cz + along * ext[2];
			c->ex = e->cx + side * (0.42f + 0.04f * RND()) * ext[0]; c->ey = e->cy + (0.30f + 0.1f * RND()) * ext[1];
			c->ez = c->cz + (RND() - 0.5f) * 0.1f;
			c->r = 0.012f; c->amp = 170.0f;
		}
	} else {
		n_prims = 24;
		for (int i = 0; i < 24; ++i) {
			synth_prim *c = &prims[i];
			c->cx = RND() * ext[0]; c->cy = RND() * ext[1]; c->cz = RND() * ext[2];
			c->ex = c->cx + (RND() - 0.5f) * 0.6f; c->ey = c->cy + (RND() - 0.5f) * 0.6f; c->ez = c->cz + (RND() - 0.5f) * 0.6f;
			c->r   = 0.006f + 0.006f * RND();
			c->amp = 150.0f + 100.0f * RND();
		}
	}
#undef RND
#pragma omp parallel for schedule(static)
	for (long long z = 0; z < (long long) D; ++z)
		for (uint32_t y = 0; y < H; ++y)
			for (uint32_t x = 0; x < W; ++x) {
				float    px = (x + 0.5f) * inv_max, py = (y + 0.5f) * inv_max, pz = ((uint32_t) z + 0.5f) * inv_max;
				uint32_t h  = hash3(x, y, (uint32_t) z, seed_lo);
				float    v  = 0.0f;
				if (kind == 0 || kind == 3) {
					for (int i = 0; i < n_prims; ++i) {
						const synth_prim *b = &prims[i];
						float dx = px - b->cx, dy = py - b->cy, dz = pz - b->cz;
						float d2 = dx * dx + dy * dy + dz * dz, s2 = b->r * b->r;
						if (d2 < 9.0f * s2) v += b->amp * expf(-0.5f * d2 / s2);
					}
					if (kind == 3) v *= 0.5f + 0.5f * ((hash3(x >> 2, y >> 2, (uint32_t) z >> 2, seed_hi) & 0xffffu) * (1.0f / 65535.0f));
					v += (float) (h & 7u) - 4.0f + 4.0f;
				} else if (kind == 1) {
					const synth_prim *e = &prims[0];
					float qx = (px - e->cx) / e->ex, qy = (py - e->cy) / e->ey, qz = (pz - e->cz) / e->ez;
					float rr = sqrtf(qx * qx + qy * qy + qz * qz);
					float sh = fabsf(rr - 1.0f) * fminf(e->ex, fminf(e->ey, e->ez));
					if (sh < e->r) v = e->amp * (1.0f - 0.6f * sh / e->r);
					for (int i = 1; i < n_prims; ++i) {
						const synth_prim *c = &prims[i];
						float d2 = capsule_dist2(px, py, pz, c);
						if (d2 < c->r * c->r) v = fmaxf(v, c->amp * (1.0f - 0.5f * d2 / (c->r * c->r)));
					}
					v += (float) (h % 13u);
				} else {
					for (int i = 0; i < n_prims; ++i) {
						const synth_prim *c = &prims[i];
						float d2 = capsule_dist2(px, py, pz, c);
						if (d2 < c->r * c->r) v = fmaxf(v, c->amp * (1.0f - 0.5f * d2 / (c->r * c->r)));
					}
					v += (float) (h % 9u);
				}
				out[((size_t) z * H + y) * W + x] = (uint8_t) fminf(fmaxf(v, 0.0f), 255.0f);
			}
	return 0;
}
