// Developer microbenchmark (B200): issue / pipe throughput of scalar FFMA versus packed FFMA2 (fma.rn.f32x2), per SM sub-partition.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_pipe fp32_pipe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
	float2 d;
	asm("{.reg .b64 a, b, c, d; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; mov.b64 c, {%6, %7}; fma.rn.f32x2 d, a, b, c; mov.b64 {%0, %1}, d;}"
	    : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
	return d;
}
template <int MODE>
__global__ void k(float* out, float s, float t, int iters, long long* cyc) {
	float a[16];
	float2 b[8];
#pragma unroll
	for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
	for (int i = 0; i < 8; i++) b[i] = make_float2(a[2 * i], a[2 * i + 1]);
	const float2 s2 = make_float2(s, s * 1.5f), t2 = make_float2(t, t * 0.5f);
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		if (MODE == 0) {        // 16 independent scalar FFMA (3 register operands)
#pragma unroll
			for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], s, t);
		} else if (MODE == 1) { // 8 independent FFMA2, full 64-bit operands
#pragma unroll
			for (int i = 0; i < 8; i++) b[i] = fma2(b[i], s2, t2);
		} else if (MODE == 2) { // 8 FFMA2 with broadcast scalar operands
#pragma unroll
			for (int i = 0; i < 8; i++) b[i] = fma2(b[i], make_float2(s, s), make_float2(t, t));
		} else if (MODE == 3) { // 8 scalar FFMA + 4 FFMA2 (same flops as mode 0)
#pragma unroll
			for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], s, t);
#pragma unroll
			for (int i = 4; i < 8; i++) b[i] = fma2(b[i], s2, t2);
		} else if (MODE == 4) { // 16 scalar FFMA with an immediate operand
#pragma unroll
			for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], 1.0001f, t);
		}
	}
	long long t1 = clock64();
	float r = 0.f;
#pragma unroll
	for (int i = 0; i < 16; i++) r += a[i];
#pragma unroll
	for (int i = 0; i < 8; i++) r += b[i].x + b[i].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_sm, int flops_per_iter_per_thread) {
	float* out; long long* cyc;
	cudaMalloc(&out, 148 * 1024 * 4 * 2); cudaMalloc(&cyc, 8);
	const int iters = 20000;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<MODE><<<148, warps_per_sm * 32>>>(out, 1.0001f, 0.0001f, 100, cyc);
	cudaEventRecord(e0);
	k<MODE><<<148, warps_per_sm * 32>>>(out, 1.0001f, 0.0001f, iters, cyc);
	cudaEventRecord(e1); cudaDeviceSynchronize();
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
	const double insts = (MODE == 0 || MODE == 4 ? 16.0 : MODE == 3 ? 12.0 : 8.0) * iters;  // per warp
	const double per_smsp = insts * warps_per_sm / 4.0;
	printf("%-34s warps/SM=%2d  cycles=%9lld  warp-inst/cycle/SMSP=%.3f  fma-lanes/cycle/SM=%.1f  TFLOP/s=%.1f\n", name, warps_per_sm, c,
	       per_smsp / c, (double)flops_per_iter_per_thread / 2 * iters * warps_per_sm * 32 / c,
	       (double)flops_per_iter_per_thread * iters * warps_per_sm * 32 * 148 / (ms * 1e-3) / 1e12);
	cudaFree(out); cudaFree(cyc);
}
int main() {
	for (int w : {4, 8, 16, 32}) {
		run<0>("FFMA x16 (3 regs)", w, 32);
		run<4>("FFMA x16 (immediate)", w, 32);
		run<1>("FFMA2 x8 (64-bit operands)", w, 32);
		run<2>("FFMA2 x8 (broadcast operands)", w, 32);
		run<3>("FFMA x8 + FFMA2 x4", w, 32);
	}
	return 0;
}
