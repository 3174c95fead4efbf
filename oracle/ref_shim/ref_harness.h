// ref_harness.h — drivers that execute ONE transliterated reference shader the way the reference's host code
// dispatches it.  Included inside the anonymous namespace of each generated translation unit, after the shader
// text, so the shader's globals (bindings, push constants, in/out variables) are in scope.  Test infrastructure.
#include "ref_args.h"

static inline unsigned rnd_up(unsigned a, unsigned b) { return (a + b - 1) / b; }

template <class F> static void for_each_invocation_8x8x8(unsigned gx, unsigned gy, unsigned gz, F f)
{
	gl_NumWorkGroups = uvec3(gx, gy, gz);
	for (unsigned wz = 0; wz < gz; ++wz)
		for (unsigned wy = 0; wy < gy; ++wy)
			for (unsigned wx = 0; wx < gx; ++wx) {
				gl_WorkGroupID = uvec3(wx, wy, wz);
				f();
			}
}

#if defined(HARNESS_GRADIENT_MAP) || defined(HARNESS_OCCUPANCY_MAP) || defined(HARNESS_VOXEL_COUNT) || defined(HARNESS_FRAG)
static void set_tfu_common(const RefArgs *a)
{
	transfer_function_uniform.sampling_factor         = a->sampling_factor;
	transfer_function_uniform.voxel_alpha_factor      = a->voxel_alpha_factor;
	transfer_function_uniform.grad_magnitude_modifier = a->grad_magnitude_modifier;
	transfer_function_uniform.use_gradient            = a->use_gradient != 0;
}
#endif

#if defined(HARNESS_GRADIENT_MAP)
// ComputeGradientMap::compute: dispatch(rndUp(W,8), rndUp(H,8), rndUp(D,8))  (compute_gradient_map.cpp:76)
static void harness_entry(RefArgs *a)
{
	volume = image3D{a->V, a->W, a->H, a->D};
	gradient_map = image3D{a->G, a->W, a->H, a->D};
	set_tfu_common(a);
	for_each_invocation_8x8x8(rnd_up(a->W, 8), rnd_up(a->H, 8), rnd_up(a->D, 8), [&] {
		for (unsigned li = 0; li < 512; ++li) {
			gl_GlobalInvocationID = uvec3(gl_WorkGroupID.x * 8 + (li & 7), gl_WorkGroupID.y * 8 + ((li >> 3) & 7), gl_WorkGroupID.z * 8 + (li >> 6));
			shader_main();
		}
	});
}
#endif

#if defined(HARNESS_OCCUPANCY_MAP)
// ComputeDistanceMap::computeOccupancy (compute_distance_map.cpp:103-140)
static void harness_entry(RefArgs *a)
{
	volume = image3D{a->V, a->W, a->H, a->D};
#ifdef PRECOMPUTED_GRADIENT
	gradient_map = image3D{a->G, a->W, a->H, a->D};
#endif
	transfer_function = sampler2D{a->tf_rgba, 256, 256};
	occupancy_map = uimage3D{a->maps[0], a->Wb, a->Hb, a->Db};
	set_tfu_common(a);
	block_size = ivec4(int(rnd_up(a->W, a->Wb)), int(rnd_up(a->H, a->Hb)), int(rnd_up(a->D, a->Db)), 0);        // :108-113, :136
	for_each_invocation_8x8x8(rnd_up(a->Wb, 8), rnd_up(a->Hb, 8), rnd_up(a->Db, 8), [&] {
		for (unsigned li = 0; li < 512; ++li) {
			gl_GlobalInvocationID = uvec3(gl_WorkGroupID.x * 8 + (li & 7), gl_WorkGroupID.y * 8 + ((li >> 3) & 7), gl_WorkGroupID.z * 8 + (li >> 6));
			shader_main();
		}
	});
}
#endif

#if defined(HARNESS_VOXEL_COUNT)
// ComputeOccupiedVoxelCount::compute, first dispatch (compute_occupied_voxel_count.cpp:84-112).
// Subgroups are S consecutive local invocation indices; subgroupAdd/Elect are emulated in two phases.
static void harness_entry(RefArgs *a)
{
	volume = image3D{a->V, a->W, a->H, a->D};
#ifdef PRECOMPUTED_GRADIENT
	gradient_map = image3D{a->G, a->W, a->H, a->D};
#endif
	set_tfu_common(a);
	transfer_function_uniform.intensity_min       = a->intensity_min;
	transfer_function_uniform.intensity_range_inv = a->intensity_range_inv;
	transfer_function_uniform.gradient_min        = a->gradient_min;
	transfer_function_uniform.gradient_range_inv  = a->gradient_range_inv;
	count = a->count;
	const unsigned S = a->subgroup_size;
	gl_NumSubgroups  = 512 / S;
	for_each_invocation_8x8x8(rnd_up(a->W, 8), rnd_up(a->H, 8), rnd_up(a->D, 8), [&] {
		for (unsigned sg = 0; sg < 512 / S; ++sg) {
			gl_SubgroupID = sg;
			sg_sum        = 0;
			for (sg_phase = 0; sg_phase < 2; ++sg_phase)
				for (unsigned l = 0; l < S; ++l) {
					const unsigned li = sg * S + l;
					sg_first          = l == 0;
					gl_GlobalInvocationID = uvec3(gl_WorkGroupID.x * 8 + (li & 7), gl_WorkGroupID.y * 8 + ((li >> 3) & 7), gl_WorkGroupID.z * 8 + (li >> 6));
					shader_main();
				}
		}
	});
}
#endif

#if defined(HARNESS_VOXEL_COUNT_REDUCE)
// the reduce loop of ComputeOccupiedVoxelCount::compute (compute_occupied_voxel_count.cpp:122-146)
static void harness_entry(RefArgs *a)
{
	count = a->count;
	const uint64_t nElements     = a->count_elements;
	const uint32_t subgroup_size = SUBGROUP_SIZE;
	uint32_t       stride_host   = 1;
	while (stride_host < nElements) {
		bufferSize = nElements;        // push constants {n_elements, stride}
		stride     = stride_host;
		const uint32_t groups = uint32_t((nElements + uint64_t(subgroup_size) * stride_host - 1) / (uint64_t(subgroup_size) * stride_host));
		for (uint32_t g = 0; g < groups; ++g) {
			sg_sum = 0;
			for (sg_phase = 0; sg_phase < 2; ++sg_phase)
				for (uint32_t l = 0; l < subgroup_size; ++l) {
					sg_first              = l == 0;
					gl_GlobalInvocationID = uvec3(g * subgroup_size + l, 0, 0);
					shader_main();
				}
		}
		stride_host *= subgroup_size;
	}
}
#endif

#if defined(HARNESS_DISTANCE_MAP) || defined(HARNESS_DISTANCE_MAP_ANISO)
static void dispatch_2d(unsigned nx, unsigned ny)
{
	for (unsigned y = 0; y < rnd_up(ny, 8) * 8; ++y)
		for (unsigned x = 0; x < rnd_up(nx, 8) * 8; ++x) {
			gl_GlobalInvocationID = uvec3(x, y, 0);
			shader_main();
		}
}
#endif

#if defined(HARNESS_DISTANCE_MAP)
// ComputeDistanceMap::computeDistance (compute_distance_map.cpp:142-175): maps[0] holds the occupancy map on entry
static void harness_entry(RefArgs *a)
{
	uimage3D distance{a->maps[0], a->Wb, a->Hb, a->Db}, swap{a->swap, a->Wb, a->Hb, a->Db};
	dist = distance; dist_swap = distance;        // stage 0: both bindings are the distance image (:156-157)
	stage = 0;
	dispatch_2d(a->Hb, a->Db);
	dist_swap = swap;                             // :165
	stage = 1;
	dispatch_2d(a->Wb, a->Db);
	stage = 2;
	dispatch_2d(a->Wb, a->Hb);
}
#endif

#if defined(HARNESS_DISTANCE_MAP_ANISO)
// ComputeDistanceMap::computeDistanceAnisotropic (compute_distance_map.cpp:177-252): occupancy in maps[7]
static void harness_entry(RefArgs *a)
{
	auto img = [&](int i) { return uimage3D{a->maps[i], a->Wb, a->Hb, a->Db}; };
	uimage3D swap{a->swap, a->Wb, a->Hb, a->Db};
	auto stage1 = [&](int idx, int direction) { stage = 0; dir = direction; dist = img(idx); dist_swap = img(7); dispatch_2d(a->Hb, a->Db); };
	auto stage2 = [&](int idx, int direction) { stage = 1; dir = direction; dist = img(idx); dist_swap = swap; dispatch_2d(a->Wb, a->Db); };
	auto stage3 = [&](int idx, int direction) { stage = 2; dir = direction; dist = img(idx); dist_swap = swap; dispatch_2d(a->Wb, a->Hb); };
	stage1(3, 1);
	stage2(3, 1);
	stage3(0, 1);
	stage3(1, -1);
	stage2(3, -1);
	stage3(2, 1);
	stage3(3, -1);
	stage1(7, -1);
	stage2(7, 1);
	stage3(4, 1);
	stage3(5, -1);
	stage2(7, -1);
	stage3(6, 1);
	stage3(7, -1);
}
#endif

#if defined(HARNESS_FRAG) || defined(HARNESS_VERT_CLIPPED) || defined(HARNESS_VERT_PLANE)
static mat4 load_mat4(const float *m)
{
	mat4 r;
	for (int c = 0; c < 4; ++c) r.c[c] = vec4(m[c * 4 + 0], m[c * 4 + 1], m[c * 4 + 2], m[c * 4 + 3]);
	return r;
}
static void set_camera_uniforms(const RefArgs *a)
{
	camera_uniform.view          = load_mat4(a->view);
	camera_uniform.proj          = load_mat4(a->proj);
	camera_uniform.view_proj_inv = load_mat4(a->view_proj_inv);
	camera_uniform.model         = load_mat4(a->model);
	camera_uniform.model_inv     = load_mat4(a->model_inv);
	ray_cast_uniform.plane       = vec4(a->plane[0], a->plane[1], a->plane[2], a->plane[3]);
	ray_cast_uniform.plane_tex   = vec4(a->plane_tex[0], a->plane_tex[1], a->plane_tex[2], a->plane_tex[3]);
	ray_cast_uniform.cam_pos_tex = vec4(a->cam_pos_tex[0], a->cam_pos_tex[1], a->cam_pos_tex[2], a->cam_pos_tex[3]);
	ray_cast_uniform.block_size  = vec4(a->block_size_f[0], a->block_size_f[1], a->block_size_f[2], a->block_size_f[3]);
	ray_cast_uniform.front_index = a->front_index;
}
#endif

#if defined(HARNESS_FRAG)
// one invocation of volume_render.frag per fragment of the batch (inputs: the interpolated ray_entry varying)
static void harness_entry(RefArgs *a)
{
	set_camera_uniforms(a);
	set_tfu_common(a);
	volume = sampler3D{a->V, a->W, a->H, a->D};
#ifdef PRECOMPUTED_GRADIENT
	gradient_sampler = sampler3D{a->G, a->W, a->H, a->D};
#endif
	transfer_function = sampler2D{a->tf_rgba, 256, 256};
#ifdef ANISOTROPIC_DISTANCE
	for (int i = 0; i < 8; ++i) distance_map[i] = usampler3D{a->maps[i], a->Wb, a->Hb, a->Db};
#else
	distance_map[0] = usampler3D{a->maps[0], a->Wb, a->Hb, a->Db};
#endif
	for (int i = 0; i < a->n_frag; ++i) {
		ray_entry = vec3(a->frag_entry[i * 3 + 0], a->frag_entry[i * 3 + 1], a->frag_entry[i * 3 + 2]);
		position  = vec4(0.0f, 0.0f, 0.5f, 1.0f);        // only read under DEPTH_ATTACHMENT
		if (a->frag_position) position = vec4(a->frag_position[i * 4 + 0], a->frag_position[i * 4 + 1], a->frag_position[i * 4 + 2], a->frag_position[i * 4 + 3]);
#ifdef DEPTH_ATTACHMENT
		i_depth.value = a->frag_depth_in ? a->frag_depth_in[i] : 0.0f;
#endif
		shader_discarded = false;
		shader_main();
		if (a->frag_discarded) a->frag_discarded[i] = shader_discarded ? 1 : 0;
		for (int k = 0; k < 4; ++k) a->frag_out[i * 4 + k] = out_color[k];
		a->frag_depth[i] = gl_FragDepth;
	}
}
#endif

#if defined(HARNESS_VERT_CLIPPED)
// draw_indexed(36) over the 8 cube vertices (volume_render_subpass.cpp:109-117,285-287)
static void harness_entry(RefArgs *a)
{
	set_camera_uniforms(a);
	static const float verts[8][3] = {{-0.5f, -0.5f, -0.5f}, {-0.5f, -0.5f, 0.5f}, {-0.5f, 0.5f, -0.5f}, {-0.5f, 0.5f, 0.5f},
	                                  {0.5f, -0.5f, -0.5f},  {0.5f, -0.5f, 0.5f},  {0.5f, 0.5f, -0.5f},  {0.5f, 0.5f, 0.5f}};
	for (int i = 0; i < 8; ++i) {
		position = vec3(verts[i][0], verts[i][1], verts[i][2]);
		shader_main();
		float *o = a->vert_out + i * 8;
		for (int k = 0; k < 4; ++k) o[k] = gl_Position[k];
		for (int k = 0; k < 3; ++k) o[4 + k] = ray_entry[k];
		o[7] = gl_ClipDistance[0];
	}
}
#endif

#if defined(HARNESS_VERT_PLANE)
// draw_indexed(12) over vertex indices 0..5 (volume_render_subpass.cpp:131-139,289-292)
static void harness_entry(RefArgs *a)
{
	set_camera_uniforms(a);
	for (int i = 0; i < 6; ++i) {
		gl_VertexIndex = i;
		shader_main();
		float *o = a->vert_out + i * 8;
		for (int k = 0; k < 4; ++k) o[k] = gl_Position[k];
		for (int k = 0; k < 3; ++k) o[4 + k] = ray_entry[k];
		o[7] = 0.0f;
	}
}
#endif
