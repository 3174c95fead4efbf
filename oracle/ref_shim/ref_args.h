// ref_args.h — one POD argument block shared by every generated reference-shader translation unit and
// by tests/ref_api.py (ctypes mirror).  Test infrastructure.
#pragma once
#include <stdint.h>

struct RefArgs {
	// volume / gradient (R8_UNORM images and samplers), W H D voxels
	uint8_t *V;
	uint8_t *G;
	int32_t  W, H, D;
	const uint8_t *tf_rgba;        // 256*256*4
	// TransferFunctionUniform
	float   sampling_factor, voxel_alpha_factor, grad_magnitude_modifier;
	int32_t use_gradient;
	float   intensity_min, intensity_range_inv, gradient_min, gradient_range_inv;
	// distance / occupancy maps (R8_UINT), extent Wb Hb Db
	uint8_t *maps[8];
	uint8_t *swap;
	int32_t  Wb, Hb, Db;
	int32_t  block_size[3];
	// occupied voxel count
	uint64_t *count;        // caller-allocated scratch of count_elements entries
	uint64_t  count_elements;
	uint32_t  subgroup_size;
	// ray caster uniforms (CameraUniform, RayCastUniform)
	float   view[16], proj[16], view_proj_inv[16], model[16], model_inv[16];
	float   plane[4], plane_tex[4], cam_pos_tex[4], block_size_f[4];
	int32_t front_index;
	// fragment batch
	int32_t      n_frag;
	const float *frag_entry;        // n_frag * 3
	float       *frag_out;          // n_frag * 4 (out_color)
	float       *frag_depth;        // n_frag
	// vertex outputs: up to 8 vertices: position_out (4), ray_entry (3), clip distance (1)
	float vert_out[8 * 8];
	// DEPTH_ATTACHMENT variants: the interpolated `position` varying (n_frag * 4), the depth subpassLoad returns (n_frag),
	// and whether the invocation executed `discard` (n_frag); all may be null for the other variants
	const float *frag_position;
	const float *frag_depth_in;
	int32_t     *frag_discarded;
};
