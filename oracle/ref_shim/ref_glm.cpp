// ref_glm.cpp — the host maths of VolumeRenderSubpass::draw evaluated with the reference's own vendored glm
// (third_party/Vulkan-samples/third_party/glm), expression for expression as written at
// src/volume_render_subpass.cpp:221-249, with the camera/projection/transform conventions of
// VS/framework/scene_graph/components/{perspective_camera.cpp:72-76,transform.cpp:92-97,camera.cpp:36-45} and
// VS/framework/rendering/subpass.cpp:29-36.  Pins orc_make_uniforms / vkv_make_uniforms.  Test infrastructure.
#include <cstdint>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_inverse.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/gtx/transform.hpp>

static void store(const glm::mat4 &m, float *o)
{
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 4; ++r) o[c * 4 + r] = m[c][r];
}

// in: cam translation(3) rotation xyzw(4) yfov aspect znear zfar | node translation(3) rotation xyzw(4) scale(3) | image_transform(16) | clip_distance
//     | volume extent(3) | map extent(3)      (all floats, 45 values)
// out: view proj view_proj_inv model model_inv (80) | plane plane_tex cam_pos_tex block_size (16) | front_index (1)
extern "C" __attribute__((visibility("default"))) void ref_make_uniforms(const float *in, float *out)
{
	const glm::vec3 ct(in[0], in[1], in[2]);
	const glm::quat cq(in[6], in[3], in[4], in[5]);        // glm::quat(w, x, y, z)
	const float     yfov = in[7], aspect = in[8], znear = in[9], zfar = in[10];
	const glm::vec3 nt(in[11], in[12], in[13]);
	const glm::quat nq(in[17], in[14], in[15], in[16]);
	const glm::vec3 ns(in[18], in[19], in[20]);
	glm::mat4       image_transform;
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 4; ++r) image_transform[c][r] = in[21 + c * 4 + r];
	const float clip_distance = in[37];

	// Transform::get_matrix: translate * mat4_cast(rotation) * scale
	const glm::mat4 cam_world = glm::translate(glm::mat4(1.0), ct) * glm::mat4_cast(cq) * glm::scale(glm::mat4(1.0), glm::vec3(1.0f));
	const glm::mat4 node      = glm::translate(glm::mat4(1.0), nt) * glm::mat4_cast(nq) * glm::scale(glm::mat4(1.0), ns);
	const glm::mat4 view      = glm::inverse(cam_world);                                   // Camera::get_view
	glm::mat4       proj      = glm::perspective(yfov, aspect, zfar, znear);               // reversed depth
	proj[1][1] *= -1;                                                                      // vulkan_style_projection

	// --- src/volume_render_subpass.cpp:221-249 ---
	const glm::mat4 camera_view_proj_inv = glm::inverse(proj * view);
	const glm::mat4 model                = node * image_transform;
	const glm::mat4 model_inv            = glm::inverse(model);
	glm::mat4       model_to_tex         = glm::translate(glm::vec3(0.5f));
	glm::mat4       global_to_tex        = model_to_tex * model_inv;
	const glm::mat4 viewInv              = glm::inverse(view);
	const glm::vec3 cam_pos_global       = viewInv[3];
	const glm::vec3 cam_pos_model        = model_inv * glm::vec4(cam_pos_global, 1.0f);
	const glm::vec4 camera_pos_tex       = model_to_tex * glm::vec4(cam_pos_model, 1.0f);
	const glm::vec3 cam_dir_global       = glm::vec3(viewInv * glm::vec4(0, 0, -1, 0));
	const glm::vec4 plane                = glm::vec4(cam_dir_global, -clip_distance - glm::dot(cam_pos_global, cam_dir_global));
	const glm::vec4 plane_tex            = glm::inverseTranspose(global_to_tex) * plane;
	const int       front_index          = (plane_tex.x < 0 ? 1 : 0) + (plane_tex.y < 0 ? 2 : 0) + (plane_tex.z < 0 ? 4 : 0);
	auto            rndUp                = [](uint32_t a, uint32_t b) { return (a + b - 1) / b; };
	const glm::vec4 block_size(rndUp((uint32_t) in[38], (uint32_t) in[41]), rndUp((uint32_t) in[39], (uint32_t) in[42]),
	                           rndUp((uint32_t) in[40], (uint32_t) in[43]), 0);

	store(view, out); store(proj, out + 16); store(camera_view_proj_inv, out + 32); store(model, out + 48); store(model_inv, out + 64);
	for (int i = 0; i < 4; ++i) {
		out[80 + i] = plane[i]; out[84 + i] = plane_tex[i]; out[88 + i] = camera_pos_tex[i]; out[92 + i] = block_size[i];
	}
	out[96] = (float) front_index;
}
