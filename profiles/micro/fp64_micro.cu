// Micro-benchmarks behind the Jacobi / QR kernel design: FP64 latency and issue rate per SM sub-partition on sm_100a,
// double rsqrt / reciprocal latency, shuffle-reduce latency, shared-memory round trip, __syncthreads cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_micro fp64_micro.cu && ./fp64_micro
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double* out, long long* cyc, int iters, double a, double b) {
	double x[CHAINS];
	for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
	__syncthreads();
	const long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
	}
	const long long t1 = clock64();
	double s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void rsqrt_kernel(double* out, long long* cyc, int iters, double a) {
	double x = 1.0 + threadIdx.x * 1e-3;
	const long long t0 = clock64();
	for (int i = 0; i < iters; ++i) x = rsqrt(x + a);
	const long long t1 = clock64();
	out[threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
	x = 1.0 + threadIdx.x * 1e-3;
	const long long t2 = clock64();
	for (int i = 0; i < iters; ++i) x = __drcp_rn(x + a);
	const long long t3 = clock64();
	out[threadIdx.x] += x;
	if (threadIdx.x == 0) cyc[1] = t3 - t2;
	x = 1.0 + threadIdx.x * 1e-3;
	const long long t4 = clock64();
	for (int i = 0; i < iters; ++i) x = sqrt(x + a);
	const long long t5 = clock64();
	out[threadIdx.x] += x;
	if (threadIdx.x == 0) cyc[2] = t5 - t4;
	x = 1.0 + threadIdx.x * 1e-3;
	const long long t6 = clock64();
	for (int i = 0; i < iters; ++i) x = a / (x + a);
	const long long t7 = clock64();
	out[threadIdx.x] += x;
	if (threadIdx.x == 0) cyc[3] = t7 - t6;
	float y = 1.0f + threadIdx.x * 1e-3f;
	const long long t8 = clock64();
	for (int i = 0; i < iters; ++i) y = rsqrtf(y + float(a));
	const long long t9 = clock64();
	out[threadIdx.x] += y;
	if (threadIdx.x == 0) cyc[4] = t9 - t8;
}

__global__ void shfl_kernel(double* out, long long* cyc, int iters) {
	double x = 1.0 + threadIdx.x * 1e-3;
	const long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
		x *= 1e-3;
	}
	const long long t1 = clock64();
	out[threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void smem_kernel(double* out, long long* cyc, int iters) {
	__shared__ double s[1024];
	double x = 1.0 + threadIdx.x * 1e-3;
	s[threadIdx.x] = x;
	__syncthreads();
	const long long t0 = clock64();
	for (int i = 0; i < iters; ++i) { s[threadIdx.x] = x; __syncwarp(); x = s[threadIdx.x ^ 1] + 1.0; __syncwarp(); }
	const long long t1 = clock64();
	out[threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
	__syncthreads();
	const long long t2 = clock64();
	for (int i = 0; i < iters; ++i) { __syncthreads(); }
	const long long t3 = clock64();
	if (threadIdx.x == 0) cyc[1] = t3 - t2;
}

template <int CHAINS>
__global__ void dmma_kernel(double* out, long long* cyc, int iters, double a, double b) {
	double d[CHAINS][2];
	for (int c = 0; c < CHAINS; ++c) { d[c][0] = threadIdx.x * 1e-3 + c; d[c][1] = 1.0; }
	__syncthreads();
	const long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int c = 0; c < CHAINS; ++c)
			asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[c][0]), "+d"(d[c][1]) : "d"(a), "d"(b));
	}
	const long long t1 = clock64();
	double s = 0; for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
	double* out; long long* cyc;
	cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 64);
	const int iters = 4096;
	printf("DFMA: cycles per dependent step of CHAINS independent fmas (one SM)\n");
#define RUN(CH, THREADS) { dfma_kernel<CH><<<1, THREADS>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize(); \
	printf("  chains %d warps/SM %2d (per SMSP %g): %.2f cycles/step  -> %.2f cycles per warp-instruction per SMSP\n", CH, THREADS / 32, THREADS / 128.0, double(*cyc) / iters, double(*cyc) / iters / (CH * (THREADS / 128.0 < 1 ? 1 : THREADS / 128.0))); }
	RUN(1, 32) RUN(2, 32) RUN(4, 32) RUN(8, 32) RUN(16, 32)
	RUN(8, 128) RUN(8, 256) RUN(8, 512) RUN(8, 1024) RUN(1, 256) RUN(1, 512) RUN(1, 1024) RUN(2, 256)
	printf("DMMA m8n8k4: cycles per step of CHAINS independent mma (one SM)\n");
#define RUNM(CH, THREADS) { dmma_kernel<CH><<<1, THREADS>>>(out, cyc, iters, 1e-3, 1e-3); cudaDeviceSynchronize(); \
	printf("  chains %d warps/SM %2d: %.2f cycles/step -> %.2f cycles per DMMA per SM\n", CH, THREADS / 32, double(*cyc) / iters, double(*cyc) / iters / (CH * (THREADS / 32))); }
	RUNM(1, 32) RUNM(2, 32) RUNM(4, 32) RUNM(8, 32) RUNM(1, 128) RUNM(4, 128) RUNM(1, 256) RUNM(2, 256) RUNM(4, 256) RUNM(8, 256)
	rsqrt_kernel<<<1, 32>>>(out, cyc, iters, 1e-9); cudaDeviceSynchronize();
	printf("dependent latency: rsqrt(double) %.1f  __drcp_rn %.1f  sqrt %.1f  div %.1f  rsqrtf %.1f cycles (incl. one DADD/FADD)\n", double(cyc[0]) / iters, double(cyc[1]) / iters, double(cyc[2]) / iters, double(cyc[3]) / iters, double(cyc[4]) / iters);
	shfl_kernel<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
	printf("5-stage double shuffle-add butterfly + DMUL: %.1f cycles\n", double(cyc[0]) / iters);
	for (int th : {32, 256, 512, 1024}) {
		smem_kernel<<<1, th>>>(out, cyc, iters); cudaDeviceSynchronize();
		printf("threads %4d: STS -> LDS -> DADD round trip %.1f cycles; __syncthreads %.1f cycles\n", th, double(cyc[0]) / iters, double(cyc[1]) / iters);
	}
	return 0;
}
