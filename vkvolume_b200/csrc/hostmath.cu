// hostmath.cu — host maths of VolumeRenderSubpass::draw (src/volume_render_subpass.cpp:219-249):
// CameraUniform and RayCastUniform from the camera node, the volume node and the image transform.
#include "../host/vkv_math.h"
#include "common.cuh"

using namespace vkv;
using namespace vkvm;

extern "C" int vkv_make_uniforms_for_extent(const uint32_t extent[3], const uint32_t map_extent[3], const vkv_camera_desc *cam,
                                            const float image_transform[16], float clip_distance, vkv_camera_uniform *cu,
                                            vkv_ray_cast_uniform *ru)
{
	VKV_REQUIRE(extent && map_extent && cam && image_transform && cu && ru, VKV_ERR_ARGUMENT, "vkv_make_uniforms: NULL argument");
	VKV_REQUIRE(map_extent[0] && map_extent[1] && map_extent[2], VKV_ERR_ARGUMENT, "vkv_make_uniforms: empty map extent");
	// camera.get_view(): inverse of the camera node's world matrix T*R*S, S = 1
	// (VS/framework/scene_graph/components/camera.cpp:36-45, transform.cpp:92-97)
	const Mat4 cam_world = translate(cam->translation[0], cam->translation[1], cam->translation[2]) *
	                       from_quat(cam->rotation[0], cam->rotation[1], cam->rotation[2], cam->rotation[3]);
	const Mat4 view = inverse(cam_world);
	// glm::perspective(fov, aspect, far, near) — reverse-Z — then Y flip (vulkan_style_projection)
	Mat4 proj  = perspective_rh_zo(cam->yfov, cam->aspect, cam->zfar, cam->znear);
	proj.m[5] *= -1.0;
	const Mat4 view_proj_inv = inverse(proj * view);
	// model = node.get_matrix() * image_transform  (volume_render_subpass.cpp:226)
	const Mat4 node = translate(cam->node_translation[0], cam->node_translation[1], cam->node_translation[2]) *
	                  from_quat(cam->node_rotation[0], cam->node_rotation[1], cam->node_rotation[2], cam->node_rotation[3]) *
	                  scale(cam->node_scale[0], cam->node_scale[1], cam->node_scale[2]);
	const Mat4 model     = node * from_float(image_transform);
	const Mat4 model_inv = inverse(model);
	to_float(view, cu->view);
	to_float(proj, cu->proj);
	to_float(view_proj_inv, cu->view_proj_inv);
	to_float(model, cu->model);
	to_float(model_inv, cu->model_inv);

	const Mat4   model_to_tex  = translate(0.5, 0.5, 0.5);
	const Mat4   global_to_tex = model_to_tex * model_inv;
	const Mat4   view_inv      = inverse(view);
	const double cam_pos[4]    = {view_inv.m[12], view_inv.m[13], view_inv.m[14], 1.0};
	double       cam_model[4], cam_tex[4], dir[4];
	mul(model_inv, cam_pos, cam_model);
	cam_model[3] = 1.0;
	mul(model_to_tex, cam_model, cam_tex);
	const double fwd[4] = {0, 0, -1, 0};
	mul(view_inv, fwd, dir);
	const double plane[4] = {dir[0], dir[1], dir[2],
	                         -(double) clip_distance - (cam_pos[0] * dir[0] + cam_pos[1] * dir[1] + cam_pos[2] * dir[2])};
	double       plane_tex[4];
	mul(transpose(inverse(global_to_tex)), plane, plane_tex);        // glm::inverseTranspose(global_to_tex) * plane
	for (int i = 0; i < 4; ++i) {
		ru->plane[i]       = (float) plane[i];
		ru->plane_tex[i]   = (float) plane_tex[i];
		ru->cam_pos_tex[i] = (float) cam_tex[i];
	}
	ru->front_index = (ru->plane_tex[0] < 0 ? 1 : 0) + (ru->plane_tex[1] < 0 ? 2 : 0) + (ru->plane_tex[2] < 0 ? 4 : 0);
	for (int a = 0; a < 3; ++a) ru->block_size[a] = (float) rnd_up(extent[a], map_extent[a]);
	ru->block_size[3] = 0.0f;
	ru->_pad[0] = ru->_pad[1] = ru->_pad[2] = 0;
	return VKV_OK;
}

extern "C" int vkv_make_uniforms(const vkv_volume *vol, const vkv_camera_desc *cam, const float image_transform[16],
                                 float clip_distance, vkv_camera_uniform *cu, vkv_ray_cast_uniform *ru)
{
	VKV_REQUIRE(vol, VKV_ERR_ARGUMENT, "vkv_make_uniforms: NULL argument");
	return vkv_make_uniforms_for_extent(vol->dim, vol->dim_b, cam, image_transform, clip_distance, cu, ru);
}
