// b200gs -- warp-autonomous alpha blending (default blend path) + heaviest-first tile schedule.
//
// Same semantics and per-pair arithmetic as blend.cu (DGR/cuda_rasterizer/forward.cu:261-374,
// backward.cu:399-557; 8 blended channels for SDP-GS), different decomposition.  ncu on the
// CTA-per-tile kernels (profiles/r01_v1_blend_ncu_summary.csv) showed SMs active only 50-70 % of the
// kernel's duration and `barrier` as the top stall: a tile's 8 warps wait for the slowest one twice
// per batch, and a few deep tiles decide the kernel time.  Here the unit of work is one warp = one
// 8x4 pixel block (8 units per tile, launched as 32-thread CTAs so the hardware CTA scheduler
// load-balances them), units are issued heaviest tile first (longest-processing-time order built by
// tile_ranges_schedule_kernel in binning.cu), and nothing ever waits on another warp:
//   * a round = 32 consecutive entries of the tile's sorted list, one per lane: the lane loads its
//     Gaussian's 32-byte geometry (g0,g1: one sector), runs the conservative block cull for it, and
//     only survivors fetch their payload (colour/depth/feature) and are appended to a ring in the
//     warp's shared memory; the per-pixel loop runs on dense batches of 32 survivors and reads them
//     as broadcasts.  All global loads are software-pipelined 1-3 rounds ahead in registers;
//   * a unit whose 32 pixels are saturated (forward) or whose deepest contributor is reached
//     (backward) exits immediately instead of idling until the rest of the tile is done;
//   * backward: the per-Gaussian sums over pixels are taken by a second phase in which a lane owns a
//     Gaussian (pixel moments of the pair weights left in shared memory by the first phase), so there
//     is no cross-lane reduction and no per-pair atomic: four 16-byte vector reductions per
//     (block, Gaussian) replace the reference's 9 float atomics per blended pair.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "blend_common.cuh"

namespace {

struct Unit {
	uint32_t tile;
	unsigned px, py;
	bool inside;
	float pxf, pyf;
	PixelBlock pb;
	uint2 range;
};

__device__ __forceinline__ Unit make_unit(const uint32_t* __restrict__ order, const uint2* __restrict__ ranges, int W, int H, int grid_x) {
	Unit u;
	const unsigned lane = threadIdx.x & 31;
	const uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	u.tile = __ldca(order + (unit >> 3));
	const int sub = unit & 7;
	const unsigned tx = u.tile % grid_x, ty = u.tile / grid_x;
	const unsigned bx = tx * TILE_X + (sub & 1) * 8, by = ty * TILE_Y + (sub >> 1) * 4;
	u.px = bx + (lane & 7);
	u.py = by + (lane >> 3);
	u.inside = u.px < (unsigned)W && u.py < (unsigned)H;
	u.pxf = (float)u.px;
	u.pyf = (float)u.py;
	u.pb.X0 = (float)bx; u.pb.X1 = u.pb.X0 + 7.f;
	u.pb.Y0 = (float)by; u.pb.Y1 = u.pb.Y0 + 3.f;
	u.range = __ldca(ranges + u.tile);
	return u;
}

// Register software pipeline shared by both kernels: a "round" is 32 consecutive list entries, one per
// lane.  ids are loaded 3 rounds ahead, geometry (g0,g1) 2 rounds ahead, cull + payload 1 round ahead, so
// the dependent chain point_list -> record -> payload (three L2 round trips) is off the critical path.
// Survivors of the cull are appended (in list order) to a 64-slot ring in shared memory; the per-pixel
// work runs on dense batches of 32 survivors regardless of how sparse the individual rounds were.
#define NOID 0xFFFFFFFFu

template <bool EXT>
__global__ void __launch_bounds__(32) blend_forward_warp_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ point_list,
	const float4* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg, float* __restrict__ final_T,
	uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
	float* __restrict__ out_alpha, float* __restrict__ out_feat)
{
	__shared__ float4 s_g0[64], s_g1[64], s_g2[64];
	__shared__ float4 s_g3[EXT ? 64 : 1];
	const unsigned lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	pdl_trigger();
	pdl_wait();
	const Unit u = make_unit(order, ranges, W, H, grid_x);
	const int n = (int)(u.range.y - u.range.x);
	constexpr int NC = EXT ? 8 : 3;

	bool done = !u.inside;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C[NC];
#pragma unroll
	for (int ch = 0; ch < NC; ch++) C[ch] = 0.f;

	// blend `count` staged survivors starting at ring slot `start` (front to back)
	auto process = [&](int start, int count) {
		for (int k0 = 0; k0 < count; k0 += 4) {
			if (__all_sync(0xFFFFFFFFu, done)) return;
			// 4 survivors at a time: their alphas are independent (ILP); only the T recurrence is serial
			float al[4];
			uint32_t ps[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				al[k] = 0.f;
				ps[k] = 0;
				if (k0 + k < count) {
					const float4 a = s_g0[start + k0 + k];
					const float4 b = s_g1[start + k0 + k];
					const float dx = __fsub_rn(a.x, u.pxf), dy = __fsub_rn(a.y, u.pyf);
					const float power = pair_power(dx, dy, a.z, a.w, b.x);
					const float alpha = fminf(0.99f, __fmul_rn(b.y, expf(power)));
					if (!(power > 0.0f) && !(alpha < 1.0f / 255.0f)) al[k] = alpha;  // al == 0 <=> skipped pair
					ps[k] = __float_as_uint(b.w);  // 1-based position in the tile's list
				}
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				if (k0 + k < count && !done && al[k] != 0.f) {
					const float alpha = al[k];
					const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
					if (test_T < 0.0001f) {
						done = true;
					} else {
						const float4 c = s_g2[start + k0 + k];
						// rgb: the reference's exact sequence fma(T, alpha*c, C) (forward.cu:355), images are bit-identical
						C[0] = __fmaf_rn(T, __fmul_rn(alpha, c.x), C[0]);
						C[1] = __fmaf_rn(T, __fmul_rn(alpha, c.y), C[1]);
						C[2] = __fmaf_rn(T, __fmul_rn(alpha, c.z), C[2]);
						if (EXT) {
							const float4 f = s_g3[start + k0 + k];
							const float w = __fmul_rn(alpha, T);
							C[3] = __fmaf_rn(w, c.w, C[3]);
							C[4] = __fadd_rn(C[4], w);
							C[5] = __fmaf_rn(w, f.x, C[5]);
							C[6] = __fmaf_rn(w, f.y, C[6]);
							C[7] = __fmaf_rn(w, f.z, C[7]);
						}
						T = test_T;
						last_contributor = ps[k];
					}
				}
			}
		}
	};

	auto load_id = [&](int base) -> uint32_t {
		const int i = base + (int)lane;
		return (i < n) ? __ldca(point_list + u.range.x + i) : NOID;
	};
	auto load_geo = [&](uint32_t id, float4& g0, float4& g1) {
		if (id != NOID) { const float4* r = rec + 4 * (size_t)id; g0 = __ldca(r); g1 = __ldca(r + 1); }
	};
	auto load_pay = [&](uint32_t id, bool keep, float4& g2, float4& g3) {
		if (keep) { const float4* r = rec + 4 * (size_t)id; g2 = __ldca(r + 2); if (EXT) g3 = __ldca(r + 3); }
	};
	float4 xg0, xg1, xg2, xg3, yg0, yg1;
	xg0 = xg1 = xg2 = xg3 = yg0 = yg1 = make_float4(0.f, 0.f, 0.f, 0.f);
	uint32_t id_x = load_id(0), id_y = load_id(32), id_z = load_id(64);
	load_geo(id_x, xg0, xg1);
	load_geo(id_y, yg0, yg1);
	bool xkeep = id_x != NOID && !cull_block(xg0, xg1, u.pb);
	load_pay(id_x, xkeep, xg2, xg3);
	int head = 0, tail = 0;

	for (int base = 0; base < n; base += 32) {
		if (__all_sync(0xFFFFFFFFu, done)) break;
		const uint32_t id_w = load_id(base + 96);                        // ids, round r+3
		float4 zg0 = make_float4(0.f, 0.f, 0.f, 0.f), zg1 = zg0;
		load_geo(id_z, zg0, zg1);                                        // geometry, round r+2
		const bool ykeep = id_y != NOID && !cull_block(yg0, yg1, u.pb);  // cull + payload, round r+1
		float4 yg2 = make_float4(0.f, 0.f, 0.f, 0.f), yg3 = yg2;
		load_pay(id_y, ykeep, yg2, yg3);

		const unsigned mask = __ballot_sync(0xFFFFFFFFu, xkeep);         // stage the survivors of round r
		if (xkeep) {
			const int slot = (head + __popc(mask & lt_mask)) & 63;
			xg1.w = __uint_as_float((uint32_t)(base + (int)lane + 1));
			s_g0[slot] = xg0; s_g1[slot] = xg1; s_g2[slot] = xg2;
			if (EXT) s_g3[slot] = xg3;
		}
		head += __popc(mask);
		if (head - tail >= 32) {
			__syncwarp();
			process(tail & 63, 32);
			tail += 32;
			__syncwarp();
		}
		xg0 = yg0; xg1 = yg1; xg2 = yg2; xg3 = yg3; xkeep = ykeep;
		yg0 = zg0; yg1 = zg1;
		id_y = id_z; id_z = id_w;
	}
	if (head > tail) {
		__syncwarp();
		process(tail & 63, head - tail);
	}

	if (u.inside) {
		const size_t pix = (size_t)u.py * W + u.px, HW = (size_t)H * W;
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

// Backward.  Per batch of up to 32 staged survivors (back to front):
//   phase 1 (lane = pixel): replay the blend recurrence and leave two weights per (pixel, Gaussian) pair in
//     shared memory: wg = G * dL/dG and wc = alpha * T.  The reference's per-channel accum_rec recurrence
//     (backward.cu:509-516) is carried as one scalar, A = sum_ch accum_rec[ch] * dL/dpixel[ch].
//   phase 2 (lane = Gaussian): each lane sums its Gaussian's column over the 32 pixels.  Every per-Gaussian
//     gradient of the reference (backward.cu:520-554) is a pixel-moment of wg -- sum wg * {1, dx, dy, dx^2,
//     dx*dy, dy^2} -- or sum wc * dL/dpixel[ch], so no cross-lane reduction and no per-pair atomics are
//     needed: one 64-byte gradient record per (block, Gaussian) goes out as four red.global.add.v4.f32.
template <bool EXT>
__global__ void __launch_bounds__(32) blend_backward_warp_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ point_list,
	const float4* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
	const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor,
	const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha_map, const float* __restrict__ dL_dfeat,
	float* __restrict__ grec)
{
	constexpr int NC = EXT ? 8 : 3;
	constexpr int WS = 33;  // padded row stride of the weight matrices: conflict-free for both phases
	__shared__ float4 s_g0[64], s_g1[64], s_g2[64];
	__shared__ float4 s_g3[EXT ? 64 : 1];
	__shared__ uint32_t s_id[64];
	__shared__ float s_wg[32 * WS], s_wc[32 * WS];
	__shared__ float4 s_dpix[32][2];
	const unsigned lane = threadIdx.x & 31;
	const unsigned gt_mask = lane == 31 ? 0u : (0xFFFFFFFFu << (lane + 1));
	pdl_trigger();
	pdl_wait();
	const Unit u = make_unit(order, ranges, W, H, grid_x);
	const size_t pix = (size_t)u.py * W + u.px, HW = (size_t)H * W;

	// everything the unit needs from the image planes is requested at once, ahead of the wmax == 0 exit test (ncu: 23 % of
	// the kernel's warp-time was spent in this prologue, serialised as n_contrib -> cotangents -> ids -> records)
	const float T_final = u.inside ? final_T[pix] : 0.f;
	float T = T_final;
	const uint32_t last_contributor = u.inside ? n_contrib[pix] : 0u;
	float dpix[NC];
#pragma unroll
	for (int ch = 0; ch < NC; ch++) dpix[ch] = 0.f;
	if (u.inside) {
		if (dL_dcolor) { dpix[0] = dL_dcolor[pix]; dpix[1] = dL_dcolor[HW + pix]; dpix[2] = dL_dcolor[2 * HW + pix]; }
		if (EXT) {
			if (dL_ddepth) dpix[3] = dL_ddepth[pix];
			if (dL_dalpha_map) dpix[4] = dL_dalpha_map[pix];
			if (dL_dfeat) { dpix[5] = dL_dfeat[pix]; dpix[6] = dL_dfeat[HW + pix]; dpix[7] = dL_dfeat[2 * HW + pix]; }
		}
	}
	uint32_t wmax = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, o));
	if (wmax == 0) return;  // nothing was blended into this block

	// cotangents of the channels that own a per-Gaussian gradient: r,g,b,z | f0,f1,f2
	s_dpix[lane][0] = make_float4(dpix[0], dpix[1], dpix[2], EXT ? dpix[3] : 0.f);
	s_dpix[lane][1] = EXT ? make_float4(dpix[5], dpix[6], dpix[7], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
	const float bg_dot_dpixel = __ldg(bg) * dpix[0] + __ldg(bg + 1) * dpix[1] + __ldg(bg + 2) * dpix[2];
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
	float A = 0.f, lastD = 0.f, last_alpha = 0.f;

	auto process = [&](int start, int count) {
		// ---- phase 1: lane = pixel
		for (int k0 = 0; k0 < count; k0 += 2) {
			float Gk[2], al[2], op[2];
			bool act[2];
#pragma unroll
			for (int k = 0; k < 2; k++) {
				Gk[k] = 0.f; al[k] = 0.f; op[k] = 0.f; act[k] = false;
				if (k0 + k < count) {
					const float4 a = s_g0[start + k0 + k];
					const float4 b = s_g1[start + k0 + k];
					const float dx = __fsub_rn(a.x, u.pxf), dy = __fsub_rn(a.y, u.pyf);
					const float power = pair_power(dx, dy, a.z, a.w, b.x);
					op[k] = b.y;
					Gk[k] = expf(power);
					al[k] = fminf(0.99f, __fmul_rn(b.y, Gk[k]));
					act[k] = (__float_as_uint(b.w) <= last_contributor) && !(power > 0.0f) && !(al[k] < 1.0f / 255.0f);
				}
			}
#pragma unroll
			for (int k = 0; k < 2; k++) {
				if (k0 + k < count) {
					float wg = 0.f, wc = 0.f;
					if (act[k]) {
						const float alpha = al[k];
						const float inv = __fdividef(1.0f, 1.0f - alpha);
						T *= inv;
						const float4 c = s_g2[start + k0 + k];
						float D = c.x * dpix[0] + c.y * dpix[1] + c.z * dpix[2];
						if (EXT) {
							const float4 f = s_g3[start + k0 + k];
							D += c.w * dpix[3] + dpix[4] + f.x * dpix[5] + f.y * dpix[6] + f.z * dpix[7];
						}
						A = last_alpha * lastD + (1.f - last_alpha) * A;
						lastD = D;
						last_alpha = alpha;
						const float dL_dalpha = (D - A) * T - (T_final * inv) * bg_dot_dpixel;
						wg = Gk[k] * (op[k] * dL_dalpha);  // G * dL/dG; clamp ignored as in backward.cu:538
						wc = alpha * T;
					}
					s_wg[lane * WS + k0 + k] = wg;
					s_wc[lane * WS + k0 + k] = wc;
				}
			}
		}
		__syncwarp();
		// ---- phase 2: lane = Gaussian
		if ((int)lane < count) {
			float S0 = 0.f, Cx = 0.f, Cy = 0.f, Cxx = 0.f, Cxy = 0.f, Cyy = 0.f;
			float g[7];
#pragma unroll
			for (int k = 0; k < 7; k++) g[k] = 0.f;
#pragma unroll
			for (int p = 0; p < 32; p++) {
				const float wg = s_wg[p * WS + lane], wc = s_wc[p * WS + lane];
				const float cx = (float)(p & 7), cy = (float)(p >> 3);
				S0 += wg;
				Cx = fmaf(wg, cx, Cx); Cy = fmaf(wg, cy, Cy);
				Cxx = fmaf(wg, cx * cx, Cxx); Cxy = fmaf(wg, cx * cy, Cxy); Cyy = fmaf(wg, cy * cy, Cyy);
				const float4 d0 = s_dpix[p][0];
				g[0] = fmaf(wc, d0.x, g[0]); g[1] = fmaf(wc, d0.y, g[1]); g[2] = fmaf(wc, d0.z, g[2]);
				if (EXT) {
					const float4 d1 = s_dpix[p][1];
					g[3] = fmaf(wc, d0.w, g[3]);
					g[4] = fmaf(wc, d1.x, g[4]); g[5] = fmaf(wc, d1.y, g[5]); g[6] = fmaf(wc, d1.z, g[6]);
				}
			}
			if (S0 != 0.f || g[0] != 0.f || g[1] != 0.f || g[2] != 0.f || (EXT && (g[3] != 0.f || g[4] != 0.f || g[5] != 0.f || g[6] != 0.f))) {
				const float4 a = s_g0[start + lane];
				const float4 b = s_g1[start + lane];
				const float ex = a.x - u.pb.X0, ey = a.y - u.pb.Y0;  // d = mean - pixel = (ex - cx, ey - cy)
				const float Sx = ex * S0 - Cx, Sy = ey * S0 - Cy;
				const float Sxx = ex * (ex * S0 - 2.f * Cx) + Cxx;
				const float Syy = ey * (ey * S0 - 2.f * Cy) + Cyy;
				const float Sxy = ex * (ey * S0 - Cy) - ey * Cx + Cxy;
				float* dst = grec + (size_t)s_id[start + lane] * GREC_FLOATS;
				red_add_v4(dst, -(a.z * Sx + a.w * Sy) * ddelx_dx, -(b.x * Sy + a.w * Sx) * ddely_dy, -0.5f * Sxx, -0.5f * Sxy);
				red_add_v4(dst + 4, -0.5f * Syy, S0 / b.y, g[0], g[1]);
				if (EXT) {
					red_add_v4(dst + 8, g[2], g[3], g[4], g[5]);
					red_add_v4(dst + 12, g[6], 0.f, 0.f, 0.f);
				} else {
					red_add_v4(dst + 8, g[2], 0.f, 0.f, 0.f);
				}
			}
		}
	};

	auto load_id = [&](int base) -> uint32_t {
		const int i = base + (int)lane;
		return (base >= 0 && (uint32_t)i < wmax) ? __ldca(point_list + u.range.x + i) : NOID;
	};
	auto load_geo = [&](uint32_t id, float4& g0, float4& g1) {
		if (id != NOID) { const float4* r = rec + 4 * (size_t)id; g0 = __ldca(r); g1 = __ldca(r + 1); }
	};
	auto load_pay = [&](uint32_t id, bool keep, float4& g2, float4& g3) {
		if (keep) { const float4* r = rec + 4 * (size_t)id; g2 = __ldca(r + 2); if (EXT) g3 = __ldca(r + 3); }
	};
	// list positions [0, wmax) back to front, 32 per round
	const int base0 = (int)((wmax - 1) & ~31u);
	float4 xg0, xg1, xg2, xg3, yg0, yg1;
	xg0 = xg1 = xg2 = xg3 = yg0 = yg1 = make_float4(0.f, 0.f, 0.f, 0.f);
	uint32_t id_x = load_id(base0), id_y = load_id(base0 - 32), id_z = load_id(base0 - 64);
	load_geo(id_x, xg0, xg1);
	load_geo(id_y, yg0, yg1);
	bool xkeep = id_x != NOID && !cull_block(xg0, xg1, u.pb);
	load_pay(id_x, xkeep, xg2, xg3);
	int head = 0, tail = 0;
	__syncwarp();  // s_dpix visible

	for (int base = base0; base >= 0; base -= 32) {
		const uint32_t id_w = load_id(base - 96);
		float4 zg0 = make_float4(0.f, 0.f, 0.f, 0.f), zg1 = zg0;
		load_geo(id_z, zg0, zg1);
		const bool ykeep = id_y != NOID && !cull_block(yg0, yg1, u.pb);
		float4 yg2 = make_float4(0.f, 0.f, 0.f, 0.f), yg3 = yg2;
		load_pay(id_y, ykeep, yg2, yg3);

		const unsigned mask = __ballot_sync(0xFFFFFFFFu, xkeep);
		if (xkeep) {  // deepest entry (highest lane) first
			const int slot = (head + __popc(mask & gt_mask)) & 63;
			xg1.w = __uint_as_float((uint32_t)(base + (int)lane + 1));
			s_g0[slot] = xg0; s_g1[slot] = xg1; s_g2[slot] = xg2; s_id[slot] = id_x;
			if (EXT) s_g3[slot] = xg3;
		}
		head += __popc(mask);
		if (head - tail >= 32) {
			__syncwarp();
			process(tail & 63, 32);
			tail += 32;
			__syncwarp();
		}
		xg0 = yg0; xg1 = yg1; xg2 = yg2; xg3 = yg3; xkeep = ykeep; id_x = id_y;
		yg0 = zg0; yg1 = zg1;
		id_y = id_z; id_z = id_w;
	}
	if (head > tail) {
		__syncwarp();
		process(tail & 63, head - tail);
	}
}

// B200GS_BLEND selects the blend kernels for A/B measurements: "warp" (this file), "async" (blend_async.cu),
// "pipe" (blend_pipe.cu), "tile" (CTA per tile, blend.cu).  Optional suffixes pick per direction, e.g. B200GS_BLEND=pipe,warp = forward pipe, backward warp.
int blend_variant(int dir) {
	static int v[2] = {-2, -2};
	if (v[0] == -2) {
		const char* e = getenv("B200GS_BLEND");
		auto parse = [](const char* s) { return (!s || !*s) ? -1 : (*s == 'a' ? 3 : (*s == 'p' ? 2 : (*s == 't' ? 1 : 0))); };
		const char* c = e ? strchr(e, ',') : nullptr;
		v[1] = c ? parse(c + 1) : parse(e);
		v[0] = parse(e);
	}
	return v[dir];  // -1: not forced, pick by shape
}

// Forward default (measured, profiles/): with few units per SM slot (small images: the kernel's duration is the walk of
// its longest lists) the cp.async list prefetch wins (64 vs 75 us at 504x378); with many waves of units (1297x840 and up)
// occupancy wins and the register-pipelined kernel, which needs less shared memory per warp, is faster (0.50 vs 0.56 ms at
// 1920x1080).
int forward_variant_for(unsigned units) {
	const int forced = blend_variant(0);
	if (forced >= 0) return forced;
	return units < 4u * 148u * 24u ? 3 : 0;
}

}  // namespace

void launch_blend_forward_tile(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                               const b200gs_outputs_t& out, cudaStream_t stream);
void launch_blend_backward_tile(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                                const b200gs_grad_outputs_t& gout, float* grec, cudaStream_t stream);

void launch_blend_forward_pipe(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                               const b200gs_outputs_t& out, cudaStream_t stream);
void launch_blend_forward_async(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                                const b200gs_outputs_t& out, cudaStream_t stream);
void launch_blend_backward_pipe(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                                const b200gs_grad_outputs_t& gout, float* grec, cudaStream_t stream);

void launch_blend_forward(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                          const b200gs_outputs_t& out, cudaStream_t stream) {
	const int gx = (v.width + TILE_X - 1) / TILE_X, gy = (v.height + TILE_Y - 1) / TILE_Y;
	const unsigned units = (unsigned)(gx * gy * 8);
	const int variant = forward_variant_for(units);
	if (variant == 3) return launch_blend_forward_async(v, gs, bs, is, out, stream);
	if (variant == 2) return launch_blend_forward_pipe(v, gs, bs, is, out, stream);
	if (variant == 1) return launch_blend_forward_tile(v, gs, bs, is, out, stream);
	if (v.extended)
		launch_k(PDL_BLEND_FWD, blend_forward_warp_kernel<true>, dim3(units), dim3(32), stream, is.ranges, is.tile_order, bs.sorted_vals, gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, out.color, out.depth, out.alpha, out.feature);
	else
		launch_k(PDL_BLEND_FWD, blend_forward_warp_kernel<false>, dim3(units), dim3(32), stream, is.ranges, is.tile_order, bs.sorted_vals, gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, out.color, nullptr, nullptr, nullptr);
	count_launch();
}

void launch_blend_backward(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                           const b200gs_grad_outputs_t& gout, float* grec, cudaStream_t stream) {
	if (blend_variant(1) == 2) return launch_blend_backward_pipe(v, gs, bs, is, gout, grec, stream);
	if (blend_variant(1) == 1) return launch_blend_backward_tile(v, gs, bs, is, gout, grec, stream);
	const int gx = (v.width + TILE_X - 1) / TILE_X, gy = (v.height + TILE_Y - 1) / TILE_Y;
	const unsigned units = (unsigned)(gx * gy * 8);
	if (v.extended)
		launch_k_first(blend_backward_warp_kernel<true>, dim3(units), dim3(32), stream, is.ranges, is.tile_order, bs.sorted_vals, gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, gout.dL_dcolor, gout.dL_ddepth, gout.dL_dalpha, gout.dL_dfeature, grec);
	else
		launch_k_first(blend_backward_warp_kernel<false>, dim3(units), dim3(32), stream, is.ranges, is.tile_order, bs.sorted_vals, gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, gout.dL_dcolor, nullptr, nullptr, nullptr, grec);
	count_launch();
}
