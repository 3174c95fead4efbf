// Microbenchmark (GPU box): what a 3-D u8 cudaArray costs to fill through surface stores, by lane -> (x, y) mapping, beside the
// linear copy of the same bytes.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a sust_patterns.cu -o sust_patterns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// tile of a warp: TX lanes along x (16 B each), 32 / TX rows of y
template <int TX, bool LIN, bool SURF, bool READ>
__global__ void __launch_bounds__(256) fill_kernel(const uint8_t *__restrict__ V, uint8_t *__restrict__ G, cudaSurfaceObject_t surf, int W, int H, int D)
{
	constexpr int TY = 32 / TX;
	const int lane = threadIdx.x & 31;
	const int tiles_x = (W / 16 + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
	const long long ntiles = (long long) tiles_x * tiles_y * D;
	for (long long t = (long long) blockIdx.x * 8 + (threadIdx.x >> 5); t < ntiles; t += (long long) gridDim.x * 8) {
		const int tx = (int) (t % tiles_x), ty = (int) ((t / tiles_x) % tiles_y), z = (int) (t / ((long long) tiles_x * tiles_y));
		const int cx = tx * TX + lane % TX, y = ty * TY + lane / TX;
		if (cx * 16 >= W || y >= H) continue;
		const size_t o = ((size_t) z * H + y) * W + cx * 16;
		uint4 v = make_uint4(cx, y, z, 7);
		if (READ) v = __ldg(reinterpret_cast<const uint4 *>(V + o));
		if (LIN) *reinterpret_cast<uint4 *>(G + o) = v;
		if (SURF) surf3Dwrite(v, surf, cx * 16, y, z);
	}
}

template <int TX, bool LIN, bool SURF, bool READ>
float run(const char *name, const uint8_t *V, uint8_t *G, cudaSurfaceObject_t s, int W, int H, int D, uint8_t *flush, size_t nflush)
{
	cudaEvent_t a, b;
	cudaEventCreate(&a), cudaEventCreate(&b);
	float best = 1e9f;
	for (int r = 0; r < 5; ++r) {
		cudaMemsetAsync(flush, r, nflush);
		cudaEventRecord(a);
		fill_kernel<TX, LIN, SURF, READ><<<148 * 8, 256>>>(V, G, s, W, H, D);
		cudaEventRecord(b);
		cudaEventSynchronize(b);
		float ms;
		cudaEventElapsedTime(&ms, a, b);
		if (ms < best) best = ms;
	}
	printf("%-28s TX=%2d  %.3f ms\n", name, TX, best);
	return best;
}

int main(int argc, char **argv)
{
	int W = 832, H = 832, D = 494;
	if (argc > 3) W = atoi(argv[1]), H = atoi(argv[2]), D = atoi(argv[3]);
	const size_t N = (size_t) W * H * D;
	uint8_t *V, *G, *flush;
	CK(cudaMalloc(&V, N)); CK(cudaMalloc(&G, N)); CK(cudaMalloc(&flush, 256u << 20));
	CK(cudaMemset(V, 3, N));
	cudaArray_t arr;
	cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
	CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(W, H, D), cudaArraySurfaceLoadStore));
	cudaResourceDesc rd = {};
	rd.resType = cudaResourceTypeArray, rd.res.array.array = arr;
	cudaSurfaceObject_t s;
	CK(cudaCreateSurfaceObject(&s, &rd));
	printf("%d x %d x %d  (%.0f MB)\n", W, H, D, N / 1e6);
	run<32, true, false, true>("copy linear", V, G, s, W, H, D, flush, 256u << 20);
	run<32, true, false, false>("write linear", V, G, s, W, H, D, flush, 256u << 20);
	run<32, false, true, false>("write surf 512Bx1", V, G, s, W, H, D, flush, 256u << 20);
	run<16, false, true, false>("write surf 256Bx2", V, G, s, W, H, D, flush, 256u << 20);
	run<8, false, true, false>("write surf 128Bx4", V, G, s, W, H, D, flush, 256u << 20);
	run<4, false, true, false>("write surf 64Bx8", V, G, s, W, H, D, flush, 256u << 20);
	run<2, false, true, false>("write surf 32Bx16", V, G, s, W, H, D, flush, 256u << 20);
	run<32, true, true, true>("copy + surf 512Bx1", V, G, s, W, H, D, flush, 256u << 20);
	run<4, true, true, true>("copy + surf 64Bx8", V, G, s, W, H, D, flush, 256u << 20);
	run<8, true, true, true>("copy + surf 128Bx4", V, G, s, W, H, D, flush, 256u << 20);
	CK(cudaDeviceSynchronize());
	return 0;
}
