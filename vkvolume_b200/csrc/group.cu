// group.cu — the TF-change rebuild sharded over the GPUs of one node, exchanges done over NVLink peer memory.
//
// The reference has nothing of the kind (single GPU).  One process per GPU; every rank holds a full replica of V and G and
// ends up with a full replica of the maps (the ray caster needs them).  Per TF change, rank r of n:
//   0. barrier — every rank has finished reading the old maps (its frames are stream-ordered before this call);
//   1. K2a (+K2b) on its z-slab of blocks                                     -> rows of map 0, partial voxel count;
//   2. isotropic distance map: K3 x and y passes on that slab (they never look outside a z slice)
//                                                                             -> xy-intermediate slab in d_swap;
//      each block row of the slab pushed into the d_swap of the peer whose z pass needs it (remote stores over NVLink: 1/n of the
//      slab per peer), partial count into every peer's signal block;
//   3. barrier — all slabs have landed;
//   4. K3 z pass on its share of the block ROWS (a z line needs every slab, nothing else)   -> rows of map 0;
//      rows pushed into every peer's map 0;
//   5. barrier — every rank holds the whole map.
// Block skipping, the octant maps and shapes the sweep kernels do not cover exchange the occupancy slabs instead (step 2)
// and run K3 whole on every rank.  The barriers are signal words in peer memory (release store into each peer's block, acquire
// spin on the own block): no host round trip, no NCCL call on the data path, everything on the caller's stream.
#include <cstring>

#include "common.cuh"

namespace vkv {

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// One CTA, one thread per rank: tell every peer "I am at `seq`" (ordered after everything this stream did before: the kernel
// boundary orders the copies, the release store publishes them system-wide), then wait until every peer has said the same.
__global__ void __launch_bounds__(32) group_barrier_kernel(GroupSignals *const *__restrict__ sig, int rank, int world, unsigned seq,
                                                           const unsigned long long *__restrict__ my_count)
{
	const int p = threadIdx.x;
	if (p >= world) return;
	if (my_count) sig[p]->count[rank] = *my_count;        // partial voxel count, published together with the signal
	__threadfence_system();
	st_release_sys(&sig[p]->arrived[rank], seq);
	const unsigned     *mine = &sig[rank]->arrived[p];
	unsigned long long  t0   = 0ull;
	unsigned            spins = 0;
	while ((int) (ld_acquire_sys(mine) - seq) < 0) {
		__nanosleep(200);
		if ((++spins & 1023u) == 0u) {        // a peer that never arrives must surface as an error: 30 s of wall time
			unsigned long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if (t0 == 0ull) t0 = now;
			else if (now - t0 > 30000000000ull) __trap();
		}
	}
}

__global__ void __launch_bounds__(32) group_sum_count_kernel(const GroupSignals *__restrict__ mine, int world, unsigned long long *__restrict__ out)
{
	if (threadIdx.x == 0) {
		unsigned long long t = 0ull;
		for (int p = 0; p < world; ++p) t += mine->count[p];
		*out = t;
	}
}

// The exchange itself: this rank's piece of a buffer — n_chunks chunks of chunk_bytes at chunk_stride (one chunk: a z-slab; Db chunks:
// a range of block rows in every slice) — read once and stored into the same place of every peer's copy of the buffer through the
// peer mappings (remote stores over NVLink, like the ray caster's framebuffer epilogue).
template <typename T>
__global__ void __launch_bounds__(256) group_push_kernel(uint8_t *const *__restrict__ dsts, const uint8_t *__restrict__ src, size_t first, size_t chunk_bytes,
                                                        size_t n_chunks, size_t chunk_stride, int rank, int world)
{
	const size_t per   = chunk_bytes / sizeof(T);
	const size_t total = per * n_chunks;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
		const size_t c = i / per, e = i - c * per;
		const size_t off = first + c * chunk_stride + e * sizeof(T);
		const T      v   = *reinterpret_cast<const T *>(src + off);
		for (int p = 0; p < world; ++p)
			if (p != rank) *reinterpret_cast<T *>(dsts[p] + off) = v;
	}
}

static int group_push(vkv_volume *vol, uint8_t *const *dsts_dev, const uint8_t *src, size_t first, size_t chunk_bytes, size_t n_chunks, size_t chunk_stride,
                      cudaStream_t s)
{
	if (chunk_bytes == 0 || n_chunks == 0 || vol->grp_world < 2) return VKV_OK;
	const size_t total = chunk_bytes * n_chunks;
	const int    grid  = (int) std::min<size_t>((size_t) vol->ctx->sm_count * 8, (total / 16 + 255) / 256 + 1);
	const bool   a16   = first % 16 == 0 && chunk_bytes % 16 == 0 && chunk_stride % 16 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0;
	const bool   a4    = first % 4 == 0 && chunk_bytes % 4 == 0 && chunk_stride % 4 == 0 && reinterpret_cast<uintptr_t>(src) % 4 == 0;
	if (a16) group_push_kernel<uint4><<<grid, 256, 0, s>>>(dsts_dev, src, first, chunk_bytes, n_chunks, chunk_stride, vol->grp_rank, vol->grp_world);
	else if (a4) group_push_kernel<uint32_t><<<grid, 256, 0, s>>>(dsts_dev, src, first, chunk_bytes, n_chunks, chunk_stride, vol->grp_rank, vol->grp_world);
	else group_push_kernel<uint8_t><<<grid, 256, 0, s>>>(dsts_dev, src, first, chunk_bytes, n_chunks, chunk_stride, vol->grp_rank, vol->grp_world);
	VKV_LAUNCHED();
	return VKV_OK;
}

// The first exchange of the sharded distance build only has to feed the z pass: the peer that transforms block rows [y0, y0 + yc)
// needs those rows of this rank's slab and nothing else — 1/world of what a broadcast of the slab moves.  One warp per (slice, row):
// the row goes to the one peer that owns it (rows are dealt out `rows_per` at a time, like split()).
__global__ void __launch_bounds__(256) group_scatter_rows_kernel(uint8_t *const *__restrict__ dsts, const uint8_t *__restrict__ src, uint32_t z0, uint32_t zc,
                                                                uint32_t Wb, uint32_t Hb, uint32_t rows_per, int rank)
{
	const int      lane = threadIdx.x & 31;
	const uint64_t nrow = (uint64_t) zc * Hb;
	for (uint64_t r = (uint64_t) blockIdx.x * 8 + (threadIdx.x >> 5); r < nrow; r += (uint64_t) gridDim.x * 8) {
		const uint32_t z = z0 + (uint32_t) (r / Hb), y = (uint32_t) (r % Hb);
		const int      p = (int) (y / rows_per);
		if (p == rank) continue;
		const size_t off = ((size_t) z * Hb + y) * Wb;
		const uint4 *sp  = reinterpret_cast<const uint4 *>(src + off);
		uint4       *dp  = reinterpret_cast<uint4 *>(dsts[p] + off);
		for (uint32_t i = lane; i < Wb / 16; i += 32) dp[i] = sp[i];
	}
}

static int group_barrier(vkv_volume *vol, const unsigned long long *count_dev, cudaStream_t s)
{
	++vol->grp_seq;
	group_barrier_kernel<<<1, 32, 0, s>>>(vol->d_grp_sig, vol->grp_rank, vol->grp_world, vol->grp_seq, count_dev);
	VKV_LAUNCHED();
	return VKV_OK;
}

static void split(uint32_t total, int rank, int world, uint32_t *first, uint32_t *count)
{
	const uint32_t per = (total + (uint32_t) world - 1u) / (uint32_t) world;
	*first             = std::min(total, (uint32_t) rank * per);
	*count             = std::min(per, total - *first);
}

}        // namespace vkv

using namespace vkv;

extern "C" {

int vkv_volume_group_export(vkv_volume *vol, uint8_t handles_out[VKV_GROUP_HANDLE_BYTES])
{
	VKV_REQUIRE(vol && handles_out, VKV_ERR_ARGUMENT, "NULL argument");
	int rc;
	if ((rc = vkv_volume_set_number_of_distance_maps(vol, 1))) return rc;
	if (!vol->d_grp) {
		VKV_CUDA_CHECK(cudaMalloc(&vol->d_grp, sizeof(GroupSignals)));
		VKV_CUDA_CHECK(cudaMemset(vol->d_grp, 0, sizeof(GroupSignals)));
	}
	if ((rc = vkv_ipc_export(vol->d_maps[0], handles_out))) return rc;
	if ((rc = vkv_ipc_export(vol->d_swap, handles_out + VKV_IPC_HANDLE_BYTES))) return rc;
	return vkv_ipc_export(vol->d_grp, handles_out + 2 * VKV_IPC_HANDLE_BYTES);
}

int vkv_volume_group_open(vkv_volume *vol, int rank, int world, const uint8_t *all_handles)
{
	VKV_REQUIRE(vol && all_handles, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(world >= 1 && world <= kGroupMax && rank >= 0 && rank < world, VKV_ERR_ARGUMENT, "bad rank / world size (at most 8 ranks)");
	VKV_REQUIRE(vol->d_grp && !vol->d_maps.empty(), VKV_ERR_STATE, "call vkv_volume_group_export first");
	VKV_REQUIRE(vol->grp_world == 0, VKV_ERR_STATE, "the volume already belongs to a group");
	int rc;
	for (int p = 0; p < world; ++p) {
		if (p == rank) {
			vol->grp_map[p] = vol->d_maps[0]; vol->grp_swap[p] = vol->d_swap; vol->grp_sig[p] = vol->d_grp;
			continue;
		}
		const uint8_t *h = all_handles + (size_t) p * VKV_GROUP_HANDLE_BYTES;
		void          *q = nullptr;
		if ((rc = vkv_ipc_open(h, &q))) return rc;
		vol->grp_map[p] = static_cast<uint8_t *>(q);
		if ((rc = vkv_ipc_open(h + VKV_IPC_HANDLE_BYTES, &q))) return rc;
		vol->grp_swap[p] = static_cast<uint8_t *>(q);
		if ((rc = vkv_ipc_open(h + 2 * VKV_IPC_HANDLE_BYTES, &q))) return rc;
		vol->grp_sig[p] = static_cast<GroupSignals *>(q);
	}
	VKV_CUDA_CHECK(cudaMalloc(&vol->d_grp_sig, kGroupMax * sizeof(GroupSignals *)));
	VKV_CUDA_CHECK(cudaMemcpy(vol->d_grp_sig, vol->grp_sig, kGroupMax * sizeof(GroupSignals *), cudaMemcpyHostToDevice));
	VKV_CUDA_CHECK(cudaMalloc(&vol->d_grp_ptrs, 2 * kGroupMax * sizeof(uint8_t *)));        // [0..8): peers' map 0, [8..16): peers' d_swap
	VKV_CUDA_CHECK(cudaMemcpy(vol->d_grp_ptrs, vol->grp_map, kGroupMax * sizeof(uint8_t *), cudaMemcpyHostToDevice));
	VKV_CUDA_CHECK(cudaMemcpy(vol->d_grp_ptrs + kGroupMax, vol->grp_swap, kGroupMax * sizeof(uint8_t *), cudaMemcpyHostToDevice));
	vol->grp_rank = rank; vol->grp_world = world;
	return VKV_OK;
}

int vkv_volume_group_close(vkv_volume *vol)
{
	VKV_REQUIRE(vol, VKV_ERR_ARGUMENT, "NULL argument");
	if (vol->grp_world == 0) return VKV_OK;
	cudaDeviceSynchronize();
	for (int p = 0; p < vol->grp_world; ++p) {
		if (p == vol->grp_rank) continue;
		if (vol->grp_map[p]) vkv_ipc_close(vol->grp_map[p]);
		if (vol->grp_swap[p]) vkv_ipc_close(vol->grp_swap[p]);
		if (vol->grp_sig[p]) vkv_ipc_close(vol->grp_sig[p]);
	}
	cudaFree(vol->d_grp_sig);
	cudaFree(vol->d_grp_ptrs);
	vol->d_grp_sig  = nullptr;
	vol->d_grp_ptrs = nullptr;
	memset(vol->grp_map, 0, sizeof vol->grp_map); memset(vol->grp_swap, 0, sizeof vol->grp_swap); memset(vol->grp_sig, 0, sizeof vol->grp_sig);
	vol->grp_world = 0; vol->grp_rank = -1;
	return VKV_OK;
}

int vkv_update_transfer_function_sharded(vkv_volume *vol, const vkv_volume_options *opt, int skipping_type, uint64_t *count_out, void *stream)
{
	VKV_REQUIRE(vol && opt, VKV_ERR_ARGUMENT, "NULL argument");
	VKV_REQUIRE(vol->grp_world >= 1, VKV_ERR_STATE, "the volume belongs to no group (vkv_volume_group_export / _open)");
	VKV_REQUIRE(skipping_type >= 0 && skipping_type <= 3, VKV_ERR_ARGUMENT, "bad skipping type");
	const int    rank = vol->grp_rank, world = vol->grp_world;
	cudaStream_t s = (cudaStream_t) stream;
	int          rc;
	cudaSetDevice(vol->ctx->device);
	const uint32_t Wb = vol->dim_b[0], Hb = vol->dim_b[1], Db = vol->dim_b[2];
	const size_t   plane = (size_t) Wb * Hb;
	uint32_t       z0, zc, y0, yc;
	split(Db, rank, world, &z0, &zc);
	split(Hb, rank, world, &y0, &yc);
	const int  n_maps  = skipping_type == VKV_SKIP_ANISOTROPIC_DISTANCE ? 8 : 1;
	const bool sharded = skipping_type == VKV_SKIP_DISTANCE && distance_shardable(vol);
	// Small volumes are not worth the three barriers and two exchanges (~60 us against the ~40 us of occupancy pass a second GPU
	// saves on 342 M voxels): every rank then rebuilds its replica on its own, with no communication at all.
	const uint64_t n_voxels = (uint64_t) vol->dim[0] * vol->dim[1] * vol->dim[2];
	if (world == 1 || (n_voxels < (512ull << 20) && !getenv("VKV_GROUP_ALWAYS"))) return vkv_update_transfer_function(vol, opt, skipping_type, count_out, stream);
	// 0. nobody is still reading the old maps
	if ((rc = group_barrier(vol, nullptr, s))) return rc;
	// 1. TF texture + masks (replicated, a few microseconds), occupancy (+ count) of the own slab into map n-1
	if ((rc = vkv_volume_update_transfer_function_texture(vol, opt, stream))) return rc;
	vkv_transfer_function_uniform u;
	vkv_transfer_function_uniform_from_options(opt, &u);
	if (count_out) VKV_CUDA_CHECK(cudaMemsetAsync(vol->d_count, 0, sizeof(unsigned long long), s));
	if ((rc = vkv_compute_occupancy_slab(vol, &u, skipping_type, z0, zc, count_out ? (uint64_t *) vol->d_count : nullptr, stream))) return rc;
	uint8_t *const occ_map = vol->d_maps[n_maps - 1];
	if (sharded) {
		// 2. x + y passes on the slab; its block rows to the peers whose z pass needs them.  Few slices per rank make the y sweep a
		// pure latency chain: its two directions then run side by side into two maps (d_swap and map 0), both exchanged.
		const bool ysplit = world > 1 && (size_t) zc * plane <= ((size_t) 12 << 20) && !getenv("VKV_DIST_NOSPLIT");        // (33 MB slabs: measured slower)
		if ((rc = launch_distance_xy_slab(vol, z0, zc, ysplit, s))) return rc;
		if (Wb % 16 == 0 && reinterpret_cast<uintptr_t>(vol->d_swap) % 16 == 0 && reinterpret_cast<uintptr_t>(vol->d_maps[0]) % 16 == 0 && !getenv("VKV_GROUP_BROADCAST")) {
			if (zc) {
				const uint32_t rows_per = (Hb + (uint32_t) world - 1u) / (uint32_t) world;        // as split()
				const int      grid     = (int) std::min<uint64_t>(((uint64_t) zc * Hb + 7) / 8, (uint64_t) vol->ctx->sm_count * 8);
				group_scatter_rows_kernel<<<grid, 256, 0, s>>>(vol->d_grp_ptrs + kGroupMax, vol->d_swap, z0, zc, Wb, Hb, rows_per, rank);
				VKV_LAUNCHED();
				if (ysplit) {
					group_scatter_rows_kernel<<<grid, 256, 0, s>>>(vol->d_grp_ptrs, vol->d_maps[0], z0, zc, Wb, Hb, rows_per, rank);
					VKV_LAUNCHED();
				}
			}
		} else {
			if ((rc = group_push(vol, vol->d_grp_ptrs + kGroupMax, vol->d_swap, z0 * plane, zc * plane, 1, 0, s))) return rc;
			if (ysplit && (rc = group_push(vol, vol->d_grp_ptrs, vol->d_maps[0], z0 * plane, zc * plane, 1, 0, s))) return rc;
		}
		if ((rc = group_barrier(vol, count_out ? vol->d_count : nullptr, s))) return rc;
		// 4. z pass on the own block rows, result rows to every peer
		if ((rc = launch_distance_z_rows(vol, y0, yc, ysplit, s))) return rc;
		if ((rc = group_push(vol, vol->d_grp_ptrs, vol->d_maps[0], (size_t) y0 * Wb, (size_t) yc * Wb, Db, plane, s))) return rc;
		if ((rc = group_barrier(vol, nullptr, s))) return rc;
		vol->occupancy_in_map = -1;
		vol->maps_valid_for   = skipping_type;
	} else {
		// occupancy slabs to every peer, then the whole transform on every rank.  (The octant maps keep the occupancy in map 7,
		// which is not exported: it travels through the peers' d_swap and is copied into place after the barrier.)
		uint8_t *const stage = n_maps == 1 ? nullptr : vol->d_swap;
		// (source and destination offsets coincide: the slab sits at z0 * plane in this rank's occupancy map and in the peers' staging map)
		if (n_maps == 1) {
			if ((rc = group_push(vol, vol->d_grp_ptrs, occ_map, z0 * plane, zc * plane, 1, 0, s))) return rc;
		} else {
			// the peers' d_swap is the destination, the own map 7 the source: stage the slab in the own d_swap first (same offset)
			if (zc) VKV_CUDA_CHECK(cudaMemcpyAsync(vol->d_swap + z0 * plane, occ_map + z0 * plane, zc * plane, cudaMemcpyDeviceToDevice, s));
			if ((rc = group_push(vol, vol->d_grp_ptrs + kGroupMax, vol->d_swap, z0 * plane, zc * plane, 1, 0, s))) return rc;
		}
		if ((rc = group_barrier(vol, count_out ? vol->d_count : nullptr, s))) return rc;
		if (stage) {
			if (z0) VKV_CUDA_CHECK(cudaMemcpyAsync(occ_map, stage, z0 * plane, cudaMemcpyDeviceToDevice, s));
			if (z0 + zc < Db) VKV_CUDA_CHECK(cudaMemcpyAsync(occ_map + (z0 + zc) * plane, stage + (z0 + zc) * plane, (Db - z0 - zc) * plane, cudaMemcpyDeviceToDevice, s));
		}
		if ((rc = vkv_compute_distance_from_occupancy(vol, skipping_type, stream))) return rc;
		// the peers must not start pushing the next rebuild's slabs into d_swap / map 0 while this rank still transforms: barrier 0
		// of the next call sees to that; a trailing barrier keeps the contract "every rank holds the maps when the call's work is done"
		if ((rc = group_barrier(vol, nullptr, s))) return rc;
	}
	if (count_out) {
		group_sum_count_kernel<<<1, 32, 0, s>>>(vol->d_grp, world, vol->d_count);
		VKV_LAUNCHED();
		VKV_CUDA_CHECK(cudaMemcpyAsync(vol->h_count, vol->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
		VKV_CUDA_CHECK(cudaStreamSynchronize(s));
		*count_out = vol->h_count[0];
	}
	return VKV_OK;
}

}        // extern "C"
