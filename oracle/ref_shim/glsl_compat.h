// glsl_compat.h — just enough of GLSL 4.60 in C++ to compile the reference's OWN shader sources
// (read from /root/reference/shaders at build time, never copied into the repository) and execute
// them on the CPU.  TEST INFRASTRUCTURE: used only to pin the oracle (oracle/vkv_oracle.c).
//
// What is real here: the shader source text (arithmetic, control flow, dispatch shape).
// What is emulated here because no Vulkan driver exists in this image: image load/store format
// conversion (R8_UNORM / R8_UINT), sampler filtering (Vulkan spec formulas: nearest, and linear with
// unnormalised coordinate u*size - 0.5, clamp-to-edge), subgroup operations.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

namespace glsl {

typedef unsigned int uint;

template <class A, class B>
struct promote {        // GLSL implicit conversions: int -> uint -> float (no doubles in these shaders)
	typedef typename std::conditional<std::is_floating_point<A>::value || std::is_floating_point<B>::value, float,
	                                  typename std::conditional<std::is_same<A, bool>::value && std::is_same<B, bool>::value, bool,
	                                                            typename std::conditional<(std::is_unsigned<A>::value && !std::is_same<A, bool>::value) ||
	                                                                                          (std::is_unsigned<B>::value && !std::is_same<B, bool>::value),
	                                                                                      uint, int>::type>::type>::type type;
};

template <class T, int N> struct vec;

template <class T, int N, int A, int B, int C>
struct swz3 {
	T d[N];
	operator vec<T, 3>() const;
	swz3 &operator=(const vec<T, 3> &v);
	template <class S> swz3 &operator*=(S s) { d[A] = T(d[A] * s); d[B] = T(d[B] * s); d[C] = T(d[C] * s); return *this; }
};

template <class T> struct vec<T, 2> {
	union {
		T d[2];
		struct { T x, y; };
		struct { T r, g; };
		swz3<T, 2, 0, 1, 1> xyy; swz3<T, 2, 1, 1, 0> yyx; swz3<T, 2, 1, 0, 1> yxy; swz3<T, 2, 0, 0, 0> xxx;
	};
	vec() : d{T(0), T(0)} {}
	vec(T a, T b) : d{a, b} {}
	template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> explicit vec(S s) : d{T(s), T(s)} {}
	vec(const vec &o) { d[0] = o.d[0]; d[1] = o.d[1]; }
	vec &operator=(const vec &o) { d[0] = o.d[0]; d[1] = o.d[1]; return *this; }
};

template <class T> struct vec<T, 3> {
	union {
		T d[3];
		struct { T x, y, z; };
		struct { T r, g, b; };
		swz3<T, 3, 0, 1, 2> xyz;
	};
	vec() : d{T(0), T(0), T(0)} {}
	vec(T a, T b, T c) : d{a, b, c} {}
	template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> explicit vec(S s) : d{T(s), T(s), T(s)} {}
	template <class U> vec(const vec<U, 3> &o) : d{T(o.d[0]), T(o.d[1]), T(o.d[2])} {}
	template <class U> explicit vec(const vec<U, 4> &o) : d{T(o.d[0]), T(o.d[1]), T(o.d[2])} {}
	template <class U, int N, int A, int B, int C> vec(const swz3<U, N, A, B, C> &s) : d{T(s.d[A]), T(s.d[B]), T(s.d[C])} {}
	vec(const vec &o) { for (int i = 0; i < 3; ++i) d[i] = o.d[i]; }
	vec &operator=(const vec &o) { for (int i = 0; i < 3; ++i) d[i] = o.d[i]; return *this; }
	template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> explicit operator S() const { return S(d[0]); }        // GLSL: scalar constructor takes the first component
	T &operator[](int i) { return d[i]; }
	const T &operator[](int i) const { return d[i]; }
};

template <class T> struct vec<T, 4> {
	union {
		T d[4];
		struct { T x, y, z, w; };
		struct { T r, g, b, a; };
		swz3<T, 4, 0, 1, 2> xyz; swz3<T, 4, 0, 1, 2> rgb;
	};
	vec() : d{T(0), T(0), T(0), T(0)} {}
	vec(T a_, T b_, T c_, T d_) : d{a_, b_, c_, d_} {}
	template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> explicit vec(S s) : d{T(s), T(s), T(s), T(s)} {}
	template <class U, class S> vec(const vec<U, 3> &v, S w_) : d{T(v.d[0]), T(v.d[1]), T(v.d[2]), T(w_)} {}
	template <class U, int N, int A, int B, int C, class S> vec(const swz3<U, N, A, B, C> &s, S w_) : d{T(s.d[A]), T(s.d[B]), T(s.d[C]), T(w_)} {}
	template <class U> vec(const vec<U, 4> &o) : d{T(o.d[0]), T(o.d[1]), T(o.d[2]), T(o.d[3])} {}
	vec(const vec &o) { for (int i = 0; i < 4; ++i) d[i] = o.d[i]; }
	vec &operator=(const vec &o) { for (int i = 0; i < 4; ++i) d[i] = o.d[i]; return *this; }
	T &operator[](int i) { return d[i]; }
	const T &operator[](int i) const { return d[i]; }
};

template <class T, int N, int A, int B, int C> swz3<T, N, A, B, C>::operator vec<T, 3>() const { return vec<T, 3>(d[A], d[B], d[C]); }
template <class T, int N, int A, int B, int C> swz3<T, N, A, B, C> &swz3<T, N, A, B, C>::operator=(const vec<T, 3> &v)
{
	d[A] = v.d[0]; d[B] = v.d[1]; d[C] = v.d[2];
	return *this;
}

typedef vec<float, 2> vec2; typedef vec<float, 3> vec3; typedef vec<float, 4> vec4;
typedef vec<int, 2> ivec2;  typedef vec<int, 3> ivec3;  typedef vec<int, 4> ivec4;
typedef vec<uint, 3> uvec3; typedef vec<uint, 4> uvec4;
typedef vec<bool, 3> bvec3;

// ---- generic component access ------------------------------------------------------------------
template <class X, class = void> struct vt { static const int n = -1; };
template <class X> struct vt<X, typename std::enable_if<std::is_arithmetic<X>::value>::type> {
	static const int n = 0; typedef X elem;
	static X get(const X &x, int) { return x; }
};
template <class T, int N> struct vt<vec<T, N>, void> {
	static const int n = N; typedef T elem;
	static T get(const vec<T, N> &v, int i) { return v.d[i]; }
};
template <class T, int N, int A, int B, int C> struct vt<swz3<T, N, A, B, C>, void> {
	static const int n = 3; typedef T elem;
	static T get(const swz3<T, N, A, B, C> &s, int i) { return s.d[i == 0 ? A : (i == 1 ? B : C)]; }
};
template <class L, class R> struct binop {
	static const int nl = vt<L>::n, nr = vt<R>::n;
	static const bool ok = nl >= 0 && nr >= 0 && (nl > 0 || nr > 0) && (nl == 0 || nr == 0 || nl == nr);
	static const int n = nl > nr ? nl : nr;
};
#define GLSL_BINOP(OP)                                                                                                   \
	template <class L, class R, class = typename std::enable_if<binop<L, R>::ok>::type>                                \
	vec<typename promote<typename vt<L>::elem, typename vt<R>::elem>::type, binop<L, R>::n> operator OP(const L &l, const R &r) \
	{                                                                                                                    \
		typedef typename promote<typename vt<L>::elem, typename vt<R>::elem>::type E;                                    \
		vec<E, binop<L, R>::n> o;                                                                                        \
		for (int i = 0; i < binop<L, R>::n; ++i) o.d[i] = E(E(vt<L>::get(l, i)) OP E(vt<R>::get(r, i)));                 \
		return o;                                                                                                        \
	}
GLSL_BINOP(+) GLSL_BINOP(-) GLSL_BINOP(*) GLSL_BINOP(/)
#undef GLSL_BINOP
template <class T, int N> vec<T, N> operator-(const vec<T, N> &v) { vec<T, N> o; for (int i = 0; i < N; ++i) o.d[i] = -v.d[i]; return o; }
template <class T, int N, class R> vec<T, N> &operator*=(vec<T, N> &v, const R &r) { v = v * r; return v; }
template <class T, int N, class R> vec<T, N> &operator+=(vec<T, N> &v, const R &r) { v = v + r; return v; }
template <class T, int N, class R> vec<T, N> &operator/=(vec<T, N> &v, const R &r) { v = v / r; return v; }

// ---- built-in functions (componentwise, with GLSL's implicit promotions) ---------------------------
inline float ceil(float x) { return std::ceil(x); }
inline float floor(float x) { return std::floor(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float pow(float x, float y) { return std::pow(x, y); }
template <class T, int N> vec<float, N> ceil(const vec<T, N> &v) { vec<float, N> o; for (int i = 0; i < N; ++i) o.d[i] = std::ceil(float(v.d[i])); return o; }

template <class A, class B, class = typename std::enable_if<vt<A>::n == 0 && vt<B>::n == 0>::type>
typename promote<A, B>::type min(A a, B b) { typedef typename promote<A, B>::type E; return E(a) < E(b) ? E(a) : E(b); }
template <class A, class B, class = typename std::enable_if<vt<A>::n == 0 && vt<B>::n == 0>::type>
typename promote<A, B>::type max(A a, B b) { typedef typename promote<A, B>::type E; return E(a) < E(b) ? E(b) : E(a); }
template <class L, class R, class = typename std::enable_if<binop<L, R>::ok>::type>
vec<typename promote<typename vt<L>::elem, typename vt<R>::elem>::type, binop<L, R>::n> min(const L &l, const R &r)
{
	typedef typename promote<typename vt<L>::elem, typename vt<R>::elem>::type E;
	vec<E, binop<L, R>::n> o;
	for (int i = 0; i < binop<L, R>::n; ++i) o.d[i] = min(E(vt<L>::get(l, i)), E(vt<R>::get(r, i)));
	return o;
}
template <class L, class R, class = typename std::enable_if<binop<L, R>::ok>::type>
vec<typename promote<typename vt<L>::elem, typename vt<R>::elem>::type, binop<L, R>::n> max(const L &l, const R &r)
{
	typedef typename promote<typename vt<L>::elem, typename vt<R>::elem>::type E;
	vec<E, binop<L, R>::n> o;
	for (int i = 0; i < binop<L, R>::n; ++i) o.d[i] = max(E(vt<L>::get(l, i)), E(vt<R>::get(r, i)));
	return o;
}
// clamp(x, lo, hi) = min(max(x, lo), hi)  (GLSL spec 8.3)
template <class X, class Lo, class Hi> auto clamp(const X &x, const Lo &lo, const Hi &hi) -> decltype(min(max(x, lo), hi)) { return min(max(x, lo), hi); }

template <class T, int N> float dot(const vec<T, N> &a, const vec<T, N> &b)
{
	float s = float(a.d[0]) * float(b.d[0]);
	for (int i = 1; i < N; ++i) s = s + float(a.d[i]) * float(b.d[i]);
	return s;
}
template <int N, int A, int B, int C> float dot(const vec<float, 3> &a, const swz3<float, N, A, B, C> &b) { return dot(a, vec<float, 3>(b)); }
template <int N> float length(const vec<float, N> &v) { return std::sqrt(dot(v, v)); }
template <int N> float distance(const vec<float, N> &a, const vec<float, N> &b) { return length(vec<float, N>(a - b)); }
template <int N> vec<float, N> normalize(const vec<float, N> &v) { return v / length(v); }
template <int N> vec<float, N> step(float edge, const vec<float, N> &x) { vec<float, N> o; for (int i = 0; i < N; ++i) o.d[i] = x.d[i] < edge ? 0.0f : 1.0f; return o; }
template <int N> vec<float, N> sign(const vec<float, N> &x) { vec<float, N> o; for (int i = 0; i < N; ++i) o.d[i] = x.d[i] > 0.0f ? 1.0f : (x.d[i] < 0.0f ? -1.0f : 0.0f); return o; }

#define GLSL_CMP(NAME, OP)                                                                                            \
	template <class A, class B, int N> vec<bool, N> NAME(const vec<A, N> &a, const vec<B, N> &b)                    \
	{                                                                                                                 \
		typedef typename promote<A, B>::type E;                                                                       \
		vec<bool, N> o;                                                                                               \
		for (int i = 0; i < N; ++i) o.d[i] = E(a.d[i]) OP E(b.d[i]);                                                  \
		return o;                                                                                                     \
	}
GLSL_CMP(lessThanEqual, <=) GLSL_CMP(greaterThanEqual, >=) GLSL_CMP(notEqual, !=)
#undef GLSL_CMP
template <int N> bool any(const vec<bool, N> &b) { bool r = false; for (int i = 0; i < N; ++i) r = r || b.d[i]; return r; }

struct mat4 {
	vec4 c[4];
	vec4 &operator[](int i) { return c[i]; }
	const vec4 &operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4 &m, const vec4 &v)
{
	vec4 o;
	for (int r = 0; r < 4; ++r) o.d[r] = ((m.c[0].d[r] * v.d[0] + m.c[1].d[r] * v.d[1]) + m.c[2].d[r] * v.d[2]) + m.c[3].d[r] * v.d[3];
	return o;
}
inline mat4 operator*(const mat4 &a, const mat4 &b)
{
	mat4 o;
	for (int c = 0; c < 4; ++c) o.c[c] = a * b.c[c];
	return o;
}

// ---- images (storage) -----------------------------------------------------------------------------------
struct image3D { uint8_t *data = nullptr; int w = 0, h = 0, d = 0; };         // r8  (UNORM)
struct uimage3D { uint8_t *data = nullptr; int w = 0, h = 0, d = 0; };        // r8ui
inline size_t texel_index(int w, int h, const ivec3 &p) { return (size_t(p.z) * h + size_t(p.y)) * w + size_t(p.x); }
inline ivec3 imageSize(const image3D &i) { return ivec3(i.w, i.h, i.d); }
inline ivec3 imageSize(const uimage3D &i) { return ivec3(i.w, i.h, i.d); }
inline vec4 imageLoad(const image3D &i, const ivec3 &p) { return vec4(float(i.data[texel_index(i.w, i.h, p)]) / 255.0f, 0.0f, 0.0f, 1.0f); }
inline uvec4 imageLoad(const uimage3D &i, const ivec3 &p) { return uvec4(uint(i.data[texel_index(i.w, i.h, p)]), 0u, 0u, 1u); }
inline void imageStore(image3D &i, const ivec3 &p, const vec4 &v)
{
	float c = v.x < 0.0f ? 0.0f : (v.x > 1.0f ? 1.0f : v.x);        // float -> UNORM8: clamp, scale, round to nearest
	i.data[texel_index(i.w, i.h, p)] = uint8_t(std::rint(c * 255.0f));
}
template <class T> inline void imageStore(uimage3D &i, const ivec3 &p, const vec<T, 4> &v) { i.data[texel_index(i.w, i.h, p)] = uint8_t(v.x); }

// ---- samplers ------------------------------------------------------------------------------------------------
struct sampler3D { const uint8_t *data = nullptr; int w = 0, h = 0, d = 0; };        // R8_UNORM, LINEAR, CLAMP_TO_EDGE
struct usampler3D { const uint8_t *data = nullptr; int w = 0, h = 0, d = 0; };       // R8_UINT, NEAREST
struct sampler2D { const uint8_t *rgba = nullptr; int w = 256, h = 256; };           // R8G8B8A8_UNORM, NEAREST, CLAMP_TO_EDGE
inline ivec3 textureSize(const sampler3D &s, int) { return ivec3(s.w, s.h, s.d); }
inline ivec3 textureSize(const usampler3D &s, int) { return ivec3(s.w, s.h, s.d); }
inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
inline vec4 texture(const sampler3D &s, const vec3 &p)
{
	// Vulkan spec "Texel Coordinate Systems": u = s*width - 0.5, i0 = floor(u), weights alpha = frac(u); CLAMP_TO_EDGE per texel
	float u = p.x * float(s.w) - 0.5f, v = p.y * float(s.h) - 0.5f, w = p.z * float(s.d) - 0.5f;
	float fu = std::floor(u), fv = std::floor(v), fw = std::floor(w);
	float a = u - fu, b = v - fv, c = w - fw;
	int   x0 = clampi(int(fu), 0, s.w - 1), x1 = clampi(int(fu) + 1, 0, s.w - 1);
	int   y0 = clampi(int(fv), 0, s.h - 1), y1 = clampi(int(fv) + 1, 0, s.h - 1);
	int   z0 = clampi(int(fw), 0, s.d - 1), z1 = clampi(int(fw) + 1, 0, s.d - 1);
	auto  t  = [&](int x, int y, int z) { return float(s.data[(size_t(z) * s.h + size_t(y)) * s.w + size_t(x)]) / 255.0f; };
	float c00 = t(x0, y0, z0) * (1.0f - a) + t(x1, y0, z0) * a, c10 = t(x0, y1, z0) * (1.0f - a) + t(x1, y1, z0) * a;
	float c01 = t(x0, y0, z1) * (1.0f - a) + t(x1, y0, z1) * a, c11 = t(x0, y1, z1) * (1.0f - a) + t(x1, y1, z1) * a;
	float c0 = c00 * (1.0f - b) + c10 * b, c1 = c01 * (1.0f - b) + c11 * b;
	return vec4(c0 * (1.0f - c) + c1 * c, 0.0f, 0.0f, 1.0f);
}
inline vec4 texture(const sampler2D &s, const vec2 &p)
{
	int x = clampi(int(std::floor(p.x * float(s.w))), 0, s.w - 1), y = clampi(int(std::floor(p.y * float(s.h))), 0, s.h - 1);
	const uint8_t *t = s.rgba + (size_t(y) * s.w + size_t(x)) * 4;
	return vec4(float(t[0]) / 255.0f, float(t[1]) / 255.0f, float(t[2]) / 255.0f, float(t[3]) / 255.0f);
}
// input attachment (DEPTH_ATTACHMENT): the harness stores the depth value this fragment's subpassLoad returns
struct subpassInput { float value = 0.0f; };
inline vec4 subpassLoad(const subpassInput &s) { return vec4(s.value, 0.0f, 0.0f, 1.0f); }
inline uvec4 texelFetch(const usampler3D &s, const ivec3 &p, int) { return uvec4(uint(s.data[(size_t(p.z) * s.h + size_t(p.y)) * s.w + size_t(p.x)]), 0u, 0u, 1u); }

// ---- invocation state + subgroup emulation (two-phase: gather the operands, then replay) ---------------------------
static uvec3 gl_GlobalInvocationID, gl_WorkGroupID, gl_NumWorkGroups;
static uint  gl_NumSubgroups, gl_SubgroupID;
static int   gl_VertexIndex;
static int      sg_phase;        // 0: collect operands, 1: replay with the reduced value
static uint64_t sg_sum;
static bool     sg_first;        // true for the lowest active invocation during replay
template <class T> inline T subgroupAdd(T x)
{
	if (sg_phase == 0) { sg_sum += uint64_t(x); return T(0); }
	return T(sg_sum);
}
inline bool subgroupElect() { return sg_phase == 1 && sg_first; }

}        // namespace glsl
