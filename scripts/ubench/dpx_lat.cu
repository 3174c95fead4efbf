// Dependent-chain latency of the packed 16-bit min/add instructions the y sweep uses (one warp, clock64 around 4096 dependent ops).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void chain(unsigned *out, unsigned a0, unsigned b, unsigned c, long long *cyc)
{
	unsigned a = a0 + threadIdx.x;
	long long t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < 4096; ++i) {
		if (OP == 0) a = __vimin3_u16x2(a, b, c + i);
		if (OP == 1) a = __viaddmin_u16x2(a, 0x00010001u, c + i);
		if (OP == 2) a = __vminu2(a, c + i);
		if (OP == 3) a = min(a, c + i);
		if (OP == 4) a = __funnelshift_r(a, c + i, 16);
		if (OP == 5) a = __shfl_up_sync(0xffffffffu, a, 1) + i;
		if (OP == 6) a = __vimin3_u32(a, b, c + i);
		if (OP == 7) a = __viaddmin_u32(a, 1u, c + i);
	}
	long long t1 = clock64();
	out[threadIdx.x] = a;
	if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
	unsigned *o; long long *c, h;
	cudaMalloc(&o, 128); cudaMalloc(&c, 8);
	const char *names[] = {"vimin3_u16x2", "viaddmin_u16x2", "vminu2", "min.u32", "funnelshift", "shfl_up+add", "vimin3_u32", "viaddmin_u32"};
#define RUN(OP) chain<OP><<<1, 32>>>(o, 7, 9, 1000, c); chain<OP><<<1, 32>>>(o, 7, 9, 1000, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-16s %.1f cycles per dependent op\n", names[OP], h / 4096.0);
	RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
	return 0;
}
