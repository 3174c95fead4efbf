/*
 * vkv_oracle.h — CPU ORACLE for the VkVolume hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libvkv.so, vkvolume_b200/) never links,
 * imports or calls anything in oracle/.
 *
 * Each function is a plain-C restatement of one reference shader / host routine and
 * cites the file:line it follows (paths relative to the reference repository root).
 * Built with -ffp-contract=off so every fp32 operation rounds exactly once, in the
 * order written.
 *
 * Parity status: PINNED against the reference's own sources.  The reference ships no tests
 * or golden vectors for this path and its application cannot be built here,
 * but its ten GLSL shaders (compiled as C++ through oracle/ref_shim), src/load_volume.cpp
 * (as is) and its glm host maths can be, and are, executed on the CPU: oracle/_ref.  Their
 * outputs on seeded cases are committed as tests/golden/reference_outputs.npz and
 * tests/test_oracle_vs_reference.py holds this restatement to them (bit-exact for the
 * integer stages).  Not pinnable here (no Vulkan driver in this image): fixed-function
 * rasteriser coverage, sampler weight precision, blending and the sRGB store, which follow
 * the Vulkan specification's formulas.  See oracle/README.md.
 */
#ifndef VKV_ORACLE_H
#define VKV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/vkv.h" /* POD layouts only */

#ifdef __cplusplus
extern "C" {
#endif

/* loader */
int  orc_parse_header(const char *text, vkv_volume_header *out);
int  orc_normalise(const void *raw, size_t n_voxels, const char *type, const char *endianness, float lo, float hi,
                   uint8_t *out);
/* transfer function */
void orc_transfer_function_uniform(const vkv_volume_options *opt, vkv_transfer_function_uniform *out);
void orc_transfer_function_texture(const vkv_volume_options *opt, uint8_t *rgba /* 256*256*4 */);
/* K1 */
void orc_gradient_map(const uint8_t *V, uint32_t W, uint32_t H, uint32_t D, int use_gradient,
                      float grad_magnitude_modifier, uint8_t *G, float *G_float_or_null);
/* K2a */
void orc_map_extent(const uint32_t dim[3], uint32_t bs_requested, uint32_t dim_b[3], uint32_t bs_eff[3]);
void orc_occupancy_map(const uint8_t *V, const uint8_t *G, const uint8_t *tf_rgba, const uint32_t dim[3],
                       uint32_t bs_requested, int use_gradient, int precomputed_gradient, uint8_t *O);
/* K2b + K2c */
uint64_t orc_occupied_voxel_count(const uint8_t *V, const uint8_t *G, const uint32_t dim[3],
                                  const vkv_transfer_function_uniform *tfu, int precomputed_gradient);
uint64_t orc_occupied_voxel_count_dispatch(const uint8_t *V, const uint8_t *G, const uint32_t dim[3],
                                           const vkv_transfer_function_uniform *tfu, int precomputed_gradient,
                                           uint32_t subgroup_size);
/* K3a / K3b: literal pass-by-pass restatement, plus the closed form */
void orc_distance_map(const uint8_t *O, const uint32_t dim_b[3], uint8_t *Dout);
void orc_distance_map_anisotropic(const uint8_t *O, const uint32_t dim_b[3], uint8_t *D8 /* 8*M */);
void orc_distance_map_closed_form(const uint8_t *O, const uint32_t dim_b[3], int octant /* -1 = isotropic */,
                                  uint8_t *Dout);
/* host maths of VolumeRenderSubpass::draw */
void orc_make_uniforms(const uint32_t dim[3], const uint32_t dim_b[3], const vkv_camera_desc *cam,
                       const float image_transform[16], float clip_distance, vkv_camera_uniform *cam_out,
                       vkv_ray_cast_uniform *ray_out);
/* K4v + K4 + framebuffer conventions */
void orc_render(const uint8_t *V, const uint8_t *G, const uint8_t *tf_rgba, const uint8_t *Dmaps,
                const uint32_t dim[3], const uint32_t dim_b[3], const vkv_camera_uniform *cam,
                const vkv_ray_cast_uniform *ray, const vkv_transfer_function_uniform *tfu,
                const vkv_render_options *opt, int precomputed_gradient, int width, int height, int y_first,
                int y_count, uint8_t *rgba8, float *rgba_float_or_null, float *depth_or_null,
                vkv_sample_counts *counts_or_null);
/* synthetic inputs (CPU twin of vkv_synth_volume; not reference behaviour) */
int orc_synth_volume(int kind, uint64_t seed, uint32_t W, uint32_t H, uint32_t D, uint8_t *out);
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
