// b200gs -- per-tile alpha blending: forward (K6) and backward (K7).
//
// Semantics: DGR/cuda_rasterizer/forward.cu:261-374 and backward.cu:399-557, generalised from 3
// colour channels to the 8 blended channels SDP-GS reads, [r,g,b | z | 1 | f0,f1,f2] (SURVEY.md
// Appendix D; background only under r,g,b).  One CTA per 16x16 tile as in the reference (tile
// ranges are part of the bit-exact contract), but:
//   * a warp owns a compact 8x4 pixel block, and before touching a batch each lane tests one
//     Gaussian's 1/255-alpha ellipse against the warp's block (exact minimum of the conic's
//     quadratic form over the block's visible edges, with a rounding margin).  Only Gaussians
//     that can reach alpha >= 1/255 somewhere in the block are evaluated per pixel.  The test is
//     conservative, so every (pixel, Gaussian) pair the reference blends is blended here with
//     the same arithmetic sequence (power = fma(q, -0.5, -(dy*(dx*b))), CUDA expf, ...);
//   * a batch's records (64 B per Gaussian: position, conic, opacity, colour, depth, feature)
//     are staged in shared memory once, so the inner loop never touches global memory (the
//     reference fetches colours from global per blended pair, forward.cu:355);
//   * the backward replaces the reference's 9 global float atomics per blended pair
//     (backward.cu:523,545-554) with a warp butterfly reduce-scatter, shared-memory accumulation
//     per batch and one 16-byte vector reduction (red.global.add.v4.f32) per (tile, Gaussian,
//     4 values).
#include "common.cuh"
#include "blend_common.cuh"

namespace {

template <bool EXT>
__global__ void __launch_bounds__(TILE_PIX) blend_forward_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const float4* __restrict__ rec, int W, int H,
	const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
	float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha, float* __restrict__ out_feat)
{
	__shared__ float4 s_g0[TILE_PIX], s_g1[TILE_PIX], s_g2[TILE_PIX];
	__shared__ float4 s_g3[EXT ? TILE_PIX : 1];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int bx = (warp & 1) * 8, by = (warp >> 1) * 4;
	const unsigned px = blockIdx.x * TILE_X + bx + (lane & 7), py = blockIdx.y * TILE_Y + by + (lane >> 3);
	const bool inside = px < (unsigned)W && py < (unsigned)H;
	const float pxf = (float)px, pyf = (float)py;
	PixelBlock pb;
	pb.X0 = (float)(blockIdx.x * TILE_X + bx); pb.X1 = pb.X0 + 7.f;
	pb.Y0 = (float)(blockIdx.y * TILE_Y + by); pb.Y1 = pb.Y0 + 3.f;

	const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
	const int n = (int)(range.y - range.x);

	bool done = !inside;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C[EXT ? 8 : 3];
#pragma unroll
	for (int ch = 0; ch < (EXT ? 8 : 3); ch++) C[ch] = 0.f;

	for (int base = 0; base < n; base += TILE_PIX) {
		if (__syncthreads_and(done)) break;
		const int cnt = min(TILE_PIX, n - base);
		if (tid < cnt) {
			const uint32_t id = point_list[range.x + base + tid];
			const float4* r = rec + 4 * (size_t)id;
			s_g0[tid] = __ldg(r); s_g1[tid] = __ldg(r + 1); s_g2[tid] = __ldg(r + 2);
			if (EXT) s_g3[tid] = __ldg(r + 3);
		}
		__syncthreads();
		if (__all_sync(0xFFFFFFFFu, done)) continue;
		for (int m = 0; m * 32 < cnt; m++) {
			const int jj = m * 32 + lane;
			bool keep = false;
			if (jj < cnt) keep = !cull_block(s_g0[jj], s_g1[jj], pb);
			unsigned mask = __ballot_sync(0xFFFFFFFFu, keep);
			while (mask) {
				const int j = m * 32 + (__ffs(mask) - 1);
				mask &= mask - 1;
				if (done) continue;
				const float4 g0 = s_g0[j];
				const float4 g1 = s_g1[j];
				const float dx = __fsub_rn(g0.x, pxf), dy = __fsub_rn(g0.y, pyf);
				const float power = pair_power(dx, dy, g0.z, g0.w, g1.x);
				if (power > 0.0f) continue;
				const float alpha = fminf(0.99f, __fmul_rn(g1.y, expf(power)));
				if (alpha < 1.0f / 255.0f) continue;
				const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
				if (test_T < 0.0001f) { done = true; continue; }
				const float4 g2 = s_g2[j];
				C[0] = __fmaf_rn(T, __fmul_rn(alpha, g2.x), C[0]);
				C[1] = __fmaf_rn(T, __fmul_rn(alpha, g2.y), C[1]);
				C[2] = __fmaf_rn(T, __fmul_rn(alpha, g2.z), C[2]);
				if (EXT) {
					const float4 g3 = s_g3[j];
					C[3] = __fmaf_rn(T, __fmul_rn(alpha, g2.w), C[3]);
					C[4] = __fmaf_rn(T, alpha, C[4]);
					C[5] = __fmaf_rn(T, __fmul_rn(alpha, g3.x), C[5]);
					C[6] = __fmaf_rn(T, __fmul_rn(alpha, g3.y), C[6]);
					C[7] = __fmaf_rn(T, __fmul_rn(alpha, g3.z), C[7]);
				}
				T = test_T;
				last_contributor = (uint32_t)(base + j + 1);
			}
		}
	}
	if (inside) {
		const size_t pix = (size_t)py * W + px, HW = (size_t)H * W;
		final_T[pix] = T;
		n_contrib[pix] = last_contributor;
		out_color[pix] = __fmaf_rn(__ldg(bg), T, C[0]);
		out_color[HW + pix] = __fmaf_rn(__ldg(bg + 1), T, C[1]);
		out_color[2 * HW + pix] = __fmaf_rn(__ldg(bg + 2), T, C[2]);
		if (EXT) {
			out_depth[pix] = C[3];
			out_alpha[pix] = C[4];
			out_feat[pix] = C[5];
			out_feat[HW + pix] = C[6];
			out_feat[2 * HW + pix] = C[7];
		}
	}
}

template <bool EXT>
__global__ void __launch_bounds__(TILE_PIX) blend_backward_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const float4* __restrict__ rec, int W, int H,
	const float* __restrict__ bg, const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
	const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha_map,
	const float* __restrict__ dL_dfeat, float* __restrict__ grec)
{
	constexpr int NC = EXT ? 8 : 3;
	__shared__ float4 s_g0[TILE_PIX], s_g1[TILE_PIX], s_g2[TILE_PIX];
	__shared__ float4 s_g3[EXT ? TILE_PIX : 1];
	__shared__ uint32_t s_id[TILE_PIX];
	__shared__ __align__(16) float s_acc[TILE_PIX * GREC_FLOATS];
	__shared__ uint32_t s_max[TILE_PIX / 32];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int bx = (warp & 1) * 8, by = (warp >> 1) * 4;
	const unsigned px = blockIdx.x * TILE_X + bx + (lane & 7), py = blockIdx.y * TILE_Y + by + (lane >> 3);
	const bool inside = px < (unsigned)W && py < (unsigned)H;
	const float pxf = (float)px, pyf = (float)py;
	PixelBlock pb;
	pb.X0 = (float)(blockIdx.x * TILE_X + bx); pb.X1 = pb.X0 + 7.f;
	pb.Y0 = (float)(blockIdx.y * TILE_Y + by); pb.Y1 = pb.Y0 + 3.f;
	const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
	const size_t pix = (size_t)py * W + px, HW = (size_t)H * W;

	const float T_final = inside ? final_T[pix] : 0.f;
	float T = T_final;
	const uint32_t last_contributor = inside ? n_contrib[pix] : 0u;
	float dpix[NC];
#pragma unroll
	for (int ch = 0; ch < NC; ch++) dpix[ch] = 0.f;
	if (inside) {
		if (dL_dcolor) { dpix[0] = dL_dcolor[pix]; dpix[1] = dL_dcolor[HW + pix]; dpix[2] = dL_dcolor[2 * HW + pix]; }
		if (EXT) {
			if (dL_ddepth) dpix[3] = dL_ddepth[pix];
			if (dL_dalpha_map) dpix[4] = dL_dalpha_map[pix];
			if (dL_dfeat) { dpix[5] = dL_dfeat[pix]; dpix[6] = dL_dfeat[HW + pix]; dpix[7] = dL_dfeat[2 * HW + pix]; }
		}
	}
	const float bg_dot_dpixel = __ldg(bg) * dpix[0] + __ldg(bg + 1) * dpix[1] + __ldg(bg + 2) * dpix[2];
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

	// entries at list positions >= max(n_contrib) over the tile are never replayed (backward.cu:487-488)
	uint32_t wmax = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, o));
	if (lane == 0) s_max[warp] = wmax;
	__syncthreads();
	uint32_t tmax = 0;
#pragma unroll
	for (int w = 0; w < TILE_PIX / 32; w++) tmax = max(tmax, s_max[w]);
	const int n = (int)tmax;  // <= range.y - range.x

	float accum_rec[NC], last_color[NC];
#pragma unroll
	for (int ch = 0; ch < NC; ch++) { accum_rec[ch] = 0.f; last_color[ch] = 0.f; }
	float last_alpha = 0.f;

	const int nbatch = (n + TILE_PIX - 1) / TILE_PIX;
	for (int b = nbatch - 1; b >= 0; b--) {
		const int base = b * TILE_PIX;
		const int cnt = min(TILE_PIX, n - base);
		__syncthreads();  // previous batch fully flushed
		if (tid < cnt) {
			const uint32_t id = point_list[range.x + base + tid];
			s_id[tid] = id;
			const float4* r = rec + 4 * (size_t)id;
			s_g0[tid] = __ldg(r); s_g1[tid] = __ldg(r + 1); s_g2[tid] = __ldg(r + 2);
			if (EXT) s_g3[tid] = __ldg(r + 3);
		}
#pragma unroll
		for (int k = 0; k < GREC_FLOATS / 4; k++)
			reinterpret_cast<float4*>(s_acc)[k * TILE_PIX + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
		__syncthreads();

		if ((int)wmax > base) {
			for (int m = (cnt - 1) / 32; m >= 0; m--) {
				const int jj = m * 32 + lane;
				bool keep = false;
				if (jj < cnt && (uint32_t)(base + jj) < wmax) keep = !cull_block(s_g0[jj], s_g1[jj], pb);
				unsigned mask = __ballot_sync(0xFFFFFFFFu, keep);
				while (mask) {
					const int bit = 31 - __clz(mask);
					mask &= ~(1u << bit);
					const int j = m * 32 + bit;
					const float4 g0 = s_g0[j];
					const float4 g1 = s_g1[j];
					const float dx = __fsub_rn(g0.x, pxf), dy = __fsub_rn(g0.y, pyf);
					const float power = pair_power(dx, dy, g0.z, g0.w, g1.x);
					const float G = expf(power);
					const float alpha = fminf(0.99f, __fmul_rn(g1.y, G));
					const bool active = ((uint32_t)(base + j) < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
					if (!__any_sync(0xFFFFFFFFu, active)) continue;
					float v[16];
#pragma unroll
					for (int k = 0; k < 16; k++) v[k] = 0.f;
					if (active) {
						T = T / (1.f - alpha);
						const float dchannel_dcolor = alpha * T;
						const float4 g2 = s_g2[j];
						float col[NC];
						col[0] = g2.x; col[1] = g2.y; col[2] = g2.z;
						if (EXT) {
							const float4 g3 = s_g3[j];
							col[3] = g2.w; col[4] = 1.0f; col[5] = g3.x; col[6] = g3.y; col[7] = g3.z;
						}
						float dL_dalpha = 0.0f;
#pragma unroll
						for (int ch = 0; ch < NC; ch++) {
							accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
							last_color[ch] = col[ch];
							dL_dalpha += (col[ch] - accum_rec[ch]) * dpix[ch];
						}
						dL_dalpha *= T;
						last_alpha = alpha;
						dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
						const float dL_dG = g1.y * dL_dalpha;  // min(0.99,.) clamp ignored in backward (backward.cu:538)
						const float gdx = G * dx, gdy = G * dy;
						const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
						const float dG_ddely = -gdy * g1.x - gdx * g0.w;
						v[0] = dL_dG * dG_ddelx * ddelx_dx;
						v[1] = dL_dG * dG_ddely * ddely_dy;
						v[2] = -0.5f * gdx * dx * dL_dG;
						v[3] = -0.5f * gdx * dy * dL_dG;
						v[4] = -0.5f * gdy * dy * dL_dG;
						v[5] = G * dL_dalpha;
						v[6] = dchannel_dcolor * dpix[0];
						v[7] = dchannel_dcolor * dpix[1];
						v[8] = dchannel_dcolor * dpix[2];
						if (EXT) {
							v[9] = dchannel_dcolor * dpix[3];
							v[10] = dchannel_dcolor * dpix[5];
							v[11] = dchannel_dcolor * dpix[6];
							v[12] = dchannel_dcolor * dpix[7];
						}
					}
					const float tot = warp_reduce_scatter16(v, lane);
					if ((lane & 1) == 0 && (lane >> 1) < (EXT ? 13 : 9)) atomicAdd(&s_acc[j * GREC_FLOATS + (lane >> 1)], tot);
				}
			}
		}
		__syncthreads();
		if (tid < cnt) {
			const float4* a = reinterpret_cast<const float4*>(s_acc) + 4 * tid;
			const float4 a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3];
			float* dst = grec + (size_t)s_id[tid] * GREC_FLOATS;
			if (a0.x != 0.f || a0.y != 0.f || a0.z != 0.f || a0.w != 0.f) red_add_v4(dst, a0.x, a0.y, a0.z, a0.w);
			if (a1.x != 0.f || a1.y != 0.f || a1.z != 0.f || a1.w != 0.f) red_add_v4(dst + 4, a1.x, a1.y, a1.z, a1.w);
			if (a2.x != 0.f || a2.y != 0.f || a2.z != 0.f || a2.w != 0.f) red_add_v4(dst + 8, a2.x, a2.y, a2.z, a2.w);
			if (EXT && (a3.x != 0.f)) red_add_v4(dst + 12, a3.x, a3.y, a3.z, a3.w);
		}
	}
}

}  // namespace

void launch_blend_forward_tile(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                          const b200gs_outputs_t& out, cudaStream_t stream) {
	const dim3 grid((v.width + TILE_X - 1) / TILE_X, (v.height + TILE_Y - 1) / TILE_Y);
	if (v.extended)
		blend_forward_kernel<true><<<grid, TILE_PIX, 0, stream>>>(is.ranges, bs.sorted_vals, gs.rec, v.width, v.height,
			v.background, is.final_T, is.n_contrib, out.color, out.depth, out.alpha, out.feature);
	else
		blend_forward_kernel<false><<<grid, TILE_PIX, 0, stream>>>(is.ranges, bs.sorted_vals, gs.rec, v.width, v.height,
			v.background, is.final_T, is.n_contrib, out.color, nullptr, nullptr, nullptr);
	count_launch();
}

void launch_blend_backward_tile(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                           const b200gs_grad_outputs_t& gout, float* grec, cudaStream_t stream) {
	const dim3 grid((v.width + TILE_X - 1) / TILE_X, (v.height + TILE_Y - 1) / TILE_Y);
	if (v.extended)
		blend_backward_kernel<true><<<grid, TILE_PIX, 0, stream>>>(is.ranges, bs.sorted_vals, gs.rec, v.width, v.height,
			v.background, is.final_T, is.n_contrib, gout.dL_dcolor, gout.dL_ddepth, gout.dL_dalpha, gout.dL_dfeature, grec);
	else
		blend_backward_kernel<false><<<grid, TILE_PIX, 0, stream>>>(is.ranges, bs.sorted_vals, gs.rec, v.width, v.height,
			v.background, is.final_T, is.n_contrib, gout.dL_dcolor, nullptr, nullptr, nullptr, grec);
	count_launch();
}
