// Stand-in for <boost/endian/conversion.hpp>, which the reference's load_volume.cpp includes but which is
// absent from this image (the third_party/boost_endian submodule is not checked out).  Only the two in-place
// conversions load_volume.cpp calls are provided; semantics are Boost's (byte swap iff the host order differs).
#pragma once
#include <cstdint>
#include <cstring>
#include <type_traits>
namespace boost { namespace endian {
template <class T> inline void swap_inplace(T &v)
{
	unsigned char b[sizeof(T)];
	std::memcpy(b, &v, sizeof(T));
	for (size_t i = 0; i < sizeof(T) / 2; ++i) { unsigned char t = b[i]; b[i] = b[sizeof(T) - 1 - i]; b[sizeof(T) - 1 - i] = t; }
	std::memcpy(&v, b, sizeof(T));
}
inline bool host_is_little() { const uint16_t x = 1; unsigned char c; std::memcpy(&c, &x, 1); return c == 1; }
template <class T> inline void big_to_native_inplace(T &v) { if (host_is_little()) swap_inplace(v); }
template <class T> inline void little_to_native_inplace(T &v) { if (!host_is_little()) swap_inplace(v); }
}}
