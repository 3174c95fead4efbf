// vkv_math.h — minimal column-major 4x4 / vec maths for the host side (glm is not installed
// in this image; the reference uses glm only to build and invert 4x4 matrices on the host:
// src/volume_render_subpass.cpp:223-239, src/load_volume.cpp:82-83, src/volume_render.cpp:227-233).
// Conventions are glm's: m[c*4 + r], column vectors, quaternions as (x, y, z, w).
// All intermediate arithmetic is fp64; results are rounded to fp32 when stored in uniforms.
#pragma once

#include <cmath>
#include <cstring>

namespace vkvm {

struct Mat4 {
	double m[16];
	double       &at(int c, int r) { return m[c * 4 + r]; }
	const double &at(int c, int r) const { return m[c * 4 + r]; }
};

inline Mat4 identity()
{
	Mat4 r{};
	r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0;
	return r;
}

inline Mat4 from_float(const float *f)
{
	Mat4 r;
	for (int i = 0; i < 16; ++i) r.m[i] = f[i];
	return r;
}

inline void to_float(const Mat4 &a, float *f)
{
	for (int i = 0; i < 16; ++i) f[i] = (float) a.m[i];
}

inline Mat4 operator*(const Mat4 &a, const Mat4 &b)
{
	Mat4 r;
	for (int c = 0; c < 4; ++c)
		for (int row = 0; row < 4; ++row) {
			double s = 0;
			for (int k = 0; k < 4; ++k) s += a.at(k, row) * b.at(c, k);
			r.at(c, row) = s;
		}
	return r;
}

inline void mul(const Mat4 &a, const double v[4], double out[4])
{
	double t[4];
	for (int r = 0; r < 4; ++r) t[r] = a.at(0, r) * v[0] + a.at(1, r) * v[1] + a.at(2, r) * v[2] + a.at(3, r) * v[3];
	std::memcpy(out, t, sizeof t);
}

inline Mat4 transpose(const Mat4 &a)
{
	Mat4 r;
	for (int c = 0; c < 4; ++c)
		for (int row = 0; row < 4; ++row) r.at(row, c) = a.at(c, row);
	return r;
}

// glm::inverse equivalent: adjugate / determinant via 2x2 sub-determinants
inline Mat4 inverse(const Mat4 &a)
{
	const double *m = a.m;
	const double  s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[9] - m[8] * m[1], s2 = m[0] * m[13] - m[12] * m[1];
	const double  s3 = m[4] * m[9] - m[8] * m[5], s4 = m[4] * m[13] - m[12] * m[5], s5 = m[8] * m[13] - m[12] * m[9];
	const double  c5 = m[10] * m[15] - m[14] * m[11], c4 = m[6] * m[15] - m[14] * m[7], c3 = m[6] * m[11] - m[10] * m[7];
	const double  c2 = m[2] * m[15] - m[14] * m[3], c1 = m[2] * m[11] - m[10] * m[3], c0 = m[2] * m[7] - m[6] * m[3];
	const double  det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
	const double  id  = 1.0 / det;
	Mat4          r;
	r.m[0]  = (m[5] * c5 - m[9] * c4 + m[13] * c3) * id;
	r.m[4]  = (-m[4] * c5 + m[8] * c4 - m[12] * c3) * id;
	r.m[8]  = (m[7] * s5 - m[11] * s4 + m[15] * s3) * id;
	r.m[12] = (-m[6] * s5 + m[10] * s4 - m[14] * s3) * id;
	r.m[1]  = (-m[1] * c5 + m[9] * c2 - m[13] * c1) * id;
	r.m[5]  = (m[0] * c5 - m[8] * c2 + m[12] * c1) * id;
	r.m[9]  = (-m[3] * s5 + m[11] * s2 - m[15] * s1) * id;
	r.m[13] = (m[2] * s5 - m[10] * s2 + m[14] * s1) * id;
	r.m[2]  = (m[1] * c4 - m[5] * c2 + m[13] * c0) * id;
	r.m[6]  = (-m[0] * c4 + m[4] * c2 - m[12] * c0) * id;
	r.m[10] = (m[3] * s4 - m[7] * s2 + m[15] * s0) * id;
	r.m[14] = (-m[2] * s4 + m[6] * s2 - m[14] * s0) * id;
	r.m[3]  = (-m[1] * c3 + m[5] * c1 - m[9] * c0) * id;
	r.m[7]  = (m[0] * c3 - m[4] * c1 + m[8] * c0) * id;
	r.m[11] = (-m[3] * s3 + m[7] * s1 - m[11] * s0) * id;
	r.m[15] = (m[2] * s3 - m[6] * s1 + m[10] * s0) * id;
	return r;
}

inline Mat4 translate(double x, double y, double z)
{
	Mat4 r  = identity();
	r.m[12] = x;
	r.m[13] = y;
	r.m[14] = z;
	return r;
}

inline Mat4 scale(double x, double y, double z)
{
	Mat4 r  = identity();
	r.m[0]  = x;
	r.m[5]  = y;
	r.m[10] = z;
	return r;
}

// glm::rotate(angle, axis)
inline Mat4 rotate(double angle, double ax, double ay, double az)
{
	const double c = std::cos(angle), s = std::sin(angle);
	const double len = std::sqrt(ax * ax + ay * ay + az * az);
	ax /= len; ay /= len; az /= len;
	const double tx = (1 - c) * ax, ty = (1 - c) * ay, tz = (1 - c) * az;
	Mat4         r  = identity();
	r.m[0] = c + tx * ax;      r.m[1] = tx * ay + s * az; r.m[2]  = tx * az - s * ay;
	r.m[4] = ty * ax - s * az; r.m[5] = c + ty * ay;      r.m[6]  = ty * az + s * ax;
	r.m[8] = tz * ax + s * ay; r.m[9] = tz * ay - s * ax; r.m[10] = c + tz * az;
	return r;
}

// glm::mat4_cast(quat(w, x, y, z)), argument order here: x, y, z, w
inline Mat4 from_quat(double x, double y, double z, double w)
{
	Mat4 r = identity();
	r.m[0] = 1 - 2 * (y * y + z * z); r.m[1] = 2 * (x * y + w * z);     r.m[2]  = 2 * (x * z - w * y);
	r.m[4] = 2 * (x * y - w * z);     r.m[5] = 1 - 2 * (x * x + z * z); r.m[6]  = 2 * (y * z + w * x);
	r.m[8] = 2 * (x * z + w * y);     r.m[9] = 2 * (y * z - w * x);     r.m[10] = 1 - 2 * (x * x + y * y);
	return r;
}

// glm::perspective with GLM_FORCE_DEPTH_ZERO_TO_ONE, right-handed (perspectiveRH_ZO), evaluated in
// fp32 like glm does so the stored matrix is glm's.
inline Mat4 perspective_rh_zo(float fovy, float aspect, float z_near, float z_far)
{
	const float tan_half = std::tan(fovy / 2.0f);
	Mat4        r{};
	r.m[0]  = 1.0f / (aspect * tan_half);
	r.m[5]  = 1.0f / tan_half;
	r.m[10] = z_far / (z_near - z_far);
	r.m[11] = -1.0;
	r.m[14] = -(z_far * z_near) / (z_far - z_near);
	return r;
}

}        // namespace vkvm
