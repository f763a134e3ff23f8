// b200gs -- depth-pipelined alpha blending (default blend path).
//
// Same semantics and per-pair arithmetic as blend_warp.cu / blend.cu (DGR/cuda_rasterizer/forward.cu:261-374,
// backward.cu:399-557; 8 blended channels for SDP-GS).  ncu on the warp-autonomous kernels
// (profiles/r01_v11_ncu_full_summary.csv) shows what bounds them: 35 M (forward) / 46 M (backward) warp
// instructions are ~30 / 40 us of issue time over 592 schedulers, yet the kernels take 74 / 92 us, because the
// duration is the latency of the DEEPEST unit -- one warp walking a 3 700-entry tile list at ~1 250 cycles per
// 32 entries -- while most SMs have run out of work.  A unit's walk is only serial in the per-pixel
// transmittance recurrence; culling, staging and evaluating alpha (~2/3 of the instructions) are not.
//
// Here a unit (one 8x4 pixel block) is a CTA of NW warps that take the list's rounds (32 entries) in turn:
// warp w owns rounds w, w+NW, ...  For its round a warp
//   1. culls the 32 entries (one per lane) against the block, stages the survivors densely in ITS shared
//      memory slice and evaluates alpha of every (survivor, pixel) pair -- no dependence on earlier rounds;
//   2. waits for the TOKEN of the previous round (per pixel: T, the channel accumulators, last contributor,
//      done flag -- handed over through shared memory, `barrier.cta.arrive` by the producer warp /
//      `barrier.cta.sync` by the consumer, one named barrier per ring slot);
//   3. runs the serial part over its survivors (test_T, accumulate) and hands the token on.
// Every pixel sees exactly the reference's operation sequence (the colour image stays bit-identical); only the
// latency of a deep unit drops from (parallel + serial) to ~max(serial, (parallel + serial) / NW) per round.
// The backward kernel has the same shape with the token (T, A, lastD, last_alpha) flowing back to front and
// the per-Gaussian moment sums + vector reductions (see blend_warp.cu) done by the owner warp after it has
// passed the token on.
#include <cstdlib>
#include "common.cuh"
#include "blend_common.cuh"

namespace {

#define NOID 0xFFFFFFFFu

__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("barrier.cta.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("barrier.cta.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

struct PUnit {
	uint32_t tile;
	unsigned px, py;
	bool inside;
	float pxf, pyf;
	PixelBlock pb;
	uint2 range;
};

__device__ __forceinline__ PUnit make_punit(uint32_t unit, const uint32_t* order, const uint2* ranges, int W, int H, int grid_x) {
	PUnit u;
	const unsigned lane = threadIdx.x & 31;
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

template <bool EXT, int NW>
__global__ void __launch_bounds__(32 * NW) blend_forward_pipe_kernel(
	const uint2* ranges, const uint32_t* order, const uint32_t* point_list, const float4* rec, int W, int H, int grid_x,
	const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
	float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha, float* __restrict__ out_feat)
{
	constexpr int NC = EXT ? 8 : 3;
	constexpr int TOK = NC + 3;  // T, C[NC], last contributor, done
	__shared__ float4 s_g0[NW][32], s_g1[NW][32], s_g2[NW][32];
	__shared__ float4 s_g3[EXT ? NW : 1][32];
	__shared__ float s_alpha[NW][32][32];  // [warp][survivor][pixel]
	__shared__ float s_tok[TOK][32];
	__shared__ int s_stop;                 // > 0: every pixel is saturated; value = warps still to be told
	const unsigned lane = threadIdx.x & 31;
	const int w = threadIdx.x >> 5;
	const unsigned lt_mask = (1u << lane) - 1u;
	pdl_trigger();
	pdl_wait();
	const PUnit u = make_punit(blockIdx.x, order, ranges, W, H, grid_x);
	const int n = (int)(u.range.y - u.range.x);
	const int R = (n + 31) >> 5;
	if (w >= (R > 0 ? R : 1)) return;  // no round for this warp (warp 0 of an empty tile still writes the background)
	if (threadIdx.x == 0) s_stop = 0;

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
	// register pipeline over this warp's own rounds (stride S entries): ids 3 rounds ahead, geometry 2, cull + payload 1
	constexpr int S = 32 * NW;
	const int base0 = 32 * w;
	float4 xg0, xg1, xg2, xg3, yg0, yg1;
	xg0 = xg1 = xg2 = xg3 = yg0 = yg1 = make_float4(0.f, 0.f, 0.f, 0.f);
	uint32_t id_x = load_id(base0), id_y = load_id(base0 + S), id_z = load_id(base0 + 2 * S);
	load_geo(id_x, xg0, xg1);
	load_geo(id_y, yg0, yg1);
	bool xkeep = id_x != NOID && !cull_block(xg0, xg1, u.pb);
	load_pay(id_x, xkeep, xg2, xg3);

	float T = 1.0f;
	float C[NC];
#pragma unroll
	for (int ch = 0; ch < NC; ch++) C[ch] = 0.f;
	uint32_t last_contributor = 0;
	bool done = !u.inside;

	auto write_outputs = [&]() {
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
	};
	if (R == 0) { write_outputs(); return; }

	for (int r = w; r < R; r += NW) {
		const int base = 32 * r;
		// ---- advance the load pipeline
		const uint32_t id_w = load_id(base + 3 * S);
		float4 zg0 = make_float4(0.f, 0.f, 0.f, 0.f), zg1 = zg0;
		load_geo(id_z, zg0, zg1);
		const bool ykeep = id_y != NOID && !cull_block(yg0, yg1, u.pb);
		float4 yg2 = make_float4(0.f, 0.f, 0.f, 0.f), yg3 = yg2;
		load_pay(id_y, ykeep, yg2, yg3);

		// ---- 1. stage this round's survivors (list order) and evaluate alpha for every (survivor, pixel) pair
		const unsigned mask = __ballot_sync(0xFFFFFFFFu, xkeep);
		const int cnt = __popc(mask);
		if (xkeep) {
			const int slot = __popc(mask & lt_mask);
			xg1.w = __uint_as_float((uint32_t)(base + (int)lane + 1));  // 1-based position in the tile's list
			s_g0[w][slot] = xg0; s_g1[w][slot] = xg1; s_g2[w][slot] = xg2;
			if (EXT) s_g3[w][slot] = xg3;
		}
		__syncwarp();
		for (int k0 = 0; k0 < cnt; k0 += 4) {
#pragma unroll
			for (int k = 0; k < 4; k++) {
				if (k0 + k < cnt) {
					const float4 a = s_g0[w][k0 + k];
					const float4 b = s_g1[w][k0 + k];
					const float dx = __fsub_rn(a.x, u.pxf), dy = __fsub_rn(a.y, u.pyf);
					const float power = pair_power(dx, dy, a.z, a.w, b.x);
					const float alpha = fminf(0.99f, __fmul_rn(b.y, expf(power)));
					s_alpha[w][k0 + k][lane] = (!(power > 0.0f) && !(alpha < 1.0f / 255.0f)) ? alpha : 0.f;  // 0 <=> skipped pair
				}
			}
		}

		// ---- 2. token of the previous round
		if (NW > 1 && r > 0) {
			bar_sync(1 + (r - 1) % NW, 64);
			const int stop = *(volatile int*)&s_stop;
			if (stop > 0) {  // all 32 pixels saturated in an earlier round: tell the next owner (unless it is the one that found out), leave
				if (stop > 1 && r + 1 < R) {
					__syncwarp();
					if (lane == 0) *(volatile int*)&s_stop = stop - 1;
					__threadfence_block();
					bar_arrive(1 + r % NW, 64);
				}
				return;
			}
			T = s_tok[0][lane];
#pragma unroll
			for (int ch = 0; ch < NC; ch++) C[ch] = s_tok[1 + ch][lane];
			last_contributor = __float_as_uint(s_tok[NC + 1][lane]);
			done = s_tok[NC + 2][lane] != 0.f;
		}

		// ---- 3. the serial part: transmittance recurrence + accumulation, front to back.  Branch-free per pair, all
		// shared-memory operands of a group of 4 fetched up front: the loop-carried chain is one FMUL (T) and one
		// predicate (done) per pair.  A skipped pair has alpha == 0, and T * (1 - 0) == T exactly.
		for (int k0 = 0; k0 < cnt; k0 += 4) {
			if (__all_sync(0xFFFFFFFFu, done)) break;
			float al[4];
			float4 cc[4], ff[4];
			uint32_t ps[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const bool v = k0 + k < cnt;
				const int kk = min(k0 + k, cnt - 1);  // never a stale slot (its bits may be NaN; weight-0 accumulation needs finite payloads)
				al[k] = v ? s_alpha[w][k0 + k][lane] : 0.f;
				cc[k] = s_g2[w][kk];
				if (EXT) ff[k] = s_g3[w][kk];
				ps[k] = __float_as_uint(s_g1[w][kk].w);
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const float alpha = al[k];
				const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
				const bool live = !done && alpha != 0.f;
				const bool stop = live && test_T < 0.0001f;
				// a pair that is not blended accumulates with weight 0 (exact for finite payloads): no branch in the chain
				const float Tb = (live && !stop) ? T : 0.f;
				// rgb: the reference's exact sequence fma(T, alpha*c, C) (forward.cu:355), images are bit-identical
				C[0] = __fmaf_rn(Tb, __fmul_rn(alpha, cc[k].x), C[0]);
				C[1] = __fmaf_rn(Tb, __fmul_rn(alpha, cc[k].y), C[1]);
				C[2] = __fmaf_rn(Tb, __fmul_rn(alpha, cc[k].z), C[2]);
				if (EXT) {
					const float wt = __fmul_rn(alpha, Tb);
					C[3] = __fmaf_rn(wt, cc[k].w, C[3]);
					C[4] = __fadd_rn(C[4], wt);
					C[5] = __fmaf_rn(wt, ff[k].x, C[5]);
					C[6] = __fmaf_rn(wt, ff[k].y, C[6]);
					C[7] = __fmaf_rn(wt, ff[k].z, C[7]);
				}
				last_contributor = (live && !stop) ? ps[k] : last_contributor;
				T = (done || stop) ? T : test_T;
				done = done || stop;
			}
		}
		const bool all_done = __all_sync(0xFFFFFFFFu, done);
		if (all_done || r == R - 1) {
			if (NW > 1 && r + 1 < R) {  // stop the other warps
				if (lane == 0) *(volatile int*)&s_stop = NW - 1;
				__threadfence_block();
				bar_arrive(1 + r % NW, 64);
			}
			write_outputs();
			return;
		}
		if (NW > 1) {
			s_tok[0][lane] = T;
#pragma unroll
			for (int ch = 0; ch < NC; ch++) s_tok[1 + ch][lane] = C[ch];
			s_tok[NC + 1][lane] = __uint_as_float(last_contributor);
			s_tok[NC + 2][lane] = done ? 1.f : 0.f;
			__threadfence_block();
			bar_arrive(1 + r % NW, 64);
		}

		__syncwarp();  // this warp's staging slice is rewritten next
		xg0 = yg0; xg1 = yg1; xg2 = yg2; xg3 = yg3; xkeep = ykeep;
		yg0 = zg0; yg1 = zg1;
		id_y = id_z; id_z = id_w;
	}
}

constexpr int FWD_NW = 2;

// Backward.  Rounds run back to front; q = R-1-r is the position of a round in that sequence and warp q % NW owns it.
// Per round the owner warp
//   1. culls + stages the survivors (deepest first) and, per (survivor, pixel) pair, leaves go = opacity * G
//      (0 when the pair was not blended in the forward) and D = sum_ch payload[ch] * dL/dpixel[ch] in shared memory;
//   2. receives the token (T, A, lastD, last_alpha per pixel) of the round behind it;
//   3. serial part: replays the recurrence (backward.cu:509-538) and overwrites (go, D) with the two pair
//      weights wg = G * dL/dG and wc = alpha * T;  hands the token on;
//   4. lane = survivor: pixel moments of the weights -> the 13 per-Gaussian gradients, four 16-byte vector
//      reductions per (block, Gaussian) (same algebra as blend_warp.cu).  This part overlaps the next warp's step 3.
template <bool EXT, int NW>
__global__ void __launch_bounds__(32 * NW) blend_backward_pipe_kernel(
	const uint2* ranges, const uint32_t* order, const uint32_t* point_list, const float4* rec, int W, int H, int grid_x,
	const float* __restrict__ bg, const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
	const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha_map,
	const float* __restrict__ dL_dfeat, float* __restrict__ grec)
{
	constexpr int NC = EXT ? 8 : 3;
	constexpr int WS = 33;  // padded row stride: conflict-free for [survivor][pixel] by pixel (steps 1,3) and by survivor (step 4)
	__shared__ float4 s_g0[NW][32], s_g1[NW][32], s_g2[NW][32];
	__shared__ float4 s_g3[EXT ? NW : 1][32];
	__shared__ uint32_t s_id[NW][32];
	__shared__ float s_wg[NW][32 * WS], s_wc[NW][32 * WS];
	__shared__ float4 s_dpix[32][2];
	__shared__ float s_tok[4][32];
	const unsigned lane = threadIdx.x & 31;
	const int w = threadIdx.x >> 5;
	const unsigned gt_mask = lane == 31 ? 0u : (0xFFFFFFFFu << (lane + 1));
	pdl_trigger();
	pdl_wait();
	const PUnit u = make_punit(blockIdx.x, order, ranges, W, H, grid_x);
	const size_t pix = (size_t)u.py * W + u.px, HW = (size_t)H * W;

	const float T_final = u.inside ? final_T[pix] : 0.f;
	const uint32_t last_contributor = u.inside ? n_contrib[pix] : 0u;
	uint32_t wmax = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, o));
	if (wmax == 0) return;  // nothing was blended into this block
	const int R = (int)((wmax + 31) >> 5);
	if (w >= R) return;

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
	if (w == 0) {  // cotangents of the channels that own a per-Gaussian gradient: r,g,b,z | f0,f1,f2 (read by every warp's step 4)
		s_dpix[lane][0] = make_float4(dpix[0], dpix[1], dpix[2], EXT ? dpix[3] : 0.f);
		s_dpix[lane][1] = EXT ? make_float4(dpix[5], dpix[6], dpix[7], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
	}
	const float bg_dot_dpixel = __ldg(bg) * dpix[0] + __ldg(bg + 1) * dpix[1] + __ldg(bg + 2) * dpix[2];
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
	float T = T_final, A = 0.f, lastD = 0.f, last_alpha = 0.f;

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
	constexpr int S = 32 * NW;
	const int base0 = 32 * (R - 1 - w);
	float4 xg0, xg1, xg2, xg3, yg0, yg1;
	xg0 = xg1 = xg2 = xg3 = yg0 = yg1 = make_float4(0.f, 0.f, 0.f, 0.f);
	uint32_t id_x = load_id(base0), id_y = load_id(base0 - S), id_z = load_id(base0 - 2 * S);
	load_geo(id_x, xg0, xg1);
	load_geo(id_y, yg0, yg1);
	bool xkeep = id_x != NOID && !cull_block(xg0, xg1, u.pb);
	load_pay(id_x, xkeep, xg2, xg3);

	for (int q = w; q < R; q += NW) {
		const int base = 32 * (R - 1 - q);
		const uint32_t id_w = load_id(base - 3 * S);
		float4 zg0 = make_float4(0.f, 0.f, 0.f, 0.f), zg1 = zg0;
		load_geo(id_z, zg0, zg1);
		const bool ykeep = id_y != NOID && !cull_block(yg0, yg1, u.pb);
		float4 yg2 = make_float4(0.f, 0.f, 0.f, 0.f), yg3 = yg2;
		load_pay(id_y, ykeep, yg2, yg3);

		// ---- 1. stage (deepest entry = highest lane first), per-pair go and D
		const unsigned mask = __ballot_sync(0xFFFFFFFFu, xkeep);
		const int cnt = __popc(mask);
		if (xkeep) {
			const int slot = __popc(mask & gt_mask);
			xg1.w = __uint_as_float((uint32_t)(base + (int)lane + 1));
			s_g0[w][slot] = xg0; s_g1[w][slot] = xg1; s_g2[w][slot] = xg2; s_id[w][slot] = id_x;
			if (EXT) s_g3[w][slot] = xg3;
		}
		__syncwarp();
		for (int k0 = 0; k0 < cnt; k0 += 2) {
#pragma unroll
			for (int k = 0; k < 2; k++) {
				if (k0 + k < cnt) {
					const float4 a = s_g0[w][k0 + k];
					const float4 b = s_g1[w][k0 + k];
					const float dx = __fsub_rn(a.x, u.pxf), dy = __fsub_rn(a.y, u.pyf);
					const float power = pair_power(dx, dy, a.z, a.w, b.x);
					const float go = __fmul_rn(b.y, expf(power));
					const bool act = (__float_as_uint(b.w) <= last_contributor) && !(power > 0.0f) && !(fminf(0.99f, go) < 1.0f / 255.0f);
					const float4 c = s_g2[w][k0 + k];
					float D = c.x * dpix[0] + c.y * dpix[1] + c.z * dpix[2];
					if (EXT) {
						const float4 f = s_g3[w][k0 + k];
						D += c.w * dpix[3] + dpix[4] + f.x * dpix[5] + f.y * dpix[6] + f.z * dpix[7];
					}
					s_wg[w][(k0 + k) * WS + lane] = act ? go : 0.f;
					s_wc[w][(k0 + k) * WS + lane] = D;
				}
			}
		}

		// ---- 2. token of the round behind this one
		if (NW > 1 && q > 0) {
			bar_sync(1 + (q - 1) % NW, 64);
			T = s_tok[0][lane]; A = s_tok[1][lane]; lastD = s_tok[2][lane]; last_alpha = s_tok[3][lane];
		}

		// ---- 3. serial part, back to front (operands of 4 pairs fetched up front; 1/(1-alpha) is off the chain)
		for (int k0 = 0; k0 < cnt; k0 += 4) {
			float gv[4], Dv[4], iv[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const bool v = k0 + k < cnt;
				gv[k] = v ? s_wg[w][(k0 + k) * WS + lane] : 0.f;
				Dv[k] = v ? s_wc[w][(k0 + k) * WS + lane] : 0.f;
				iv[k] = __fdividef(1.0f, 1.0f - fminf(0.99f, gv[k]));
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				float wg = 0.f, wc = 0.f;
				if (gv[k] != 0.f) {
					const float alpha = fminf(0.99f, gv[k]);
					T *= iv[k];
					A = last_alpha * lastD + (1.f - last_alpha) * A;
					lastD = Dv[k];
					last_alpha = alpha;
					const float dL_dalpha = (Dv[k] - A) * T - (T_final * iv[k]) * bg_dot_dpixel;
					wg = gv[k] * dL_dalpha;  // G * dL/dG = (o G) * dL/dalpha; clamp ignored as in backward.cu:538
					wc = alpha * T;
				}
				if (k0 + k < cnt) {
					s_wg[w][(k0 + k) * WS + lane] = wg;
					s_wc[w][(k0 + k) * WS + lane] = wc;
				}
			}
		}
		if (NW > 1 && q + 1 < R) {
			s_tok[0][lane] = T; s_tok[1][lane] = A; s_tok[2][lane] = lastD; s_tok[3][lane] = last_alpha;
			__threadfence_block();
			bar_arrive(1 + q % NW, 64);
		}
		__syncwarp();

		// ---- 4. lane = survivor
		if ((int)lane < cnt) {
			float S0 = 0.f, Cx = 0.f, Cy = 0.f, Cxx = 0.f, Cxy = 0.f, Cyy = 0.f;
			float g[7];
#pragma unroll
			for (int k = 0; k < 7; k++) g[k] = 0.f;
#pragma unroll
			for (int p = 0; p < 32; p++) {
				const float wg = s_wg[w][lane * WS + p], wc = s_wc[w][lane * WS + p];
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
				const float4 a = s_g0[w][lane];
				const float4 b = s_g1[w][lane];
				const float ex = a.x - u.pb.X0, ey = a.y - u.pb.Y0;  // d = mean - pixel = (ex - cx, ey - cy)
				const float Sx = ex * S0 - Cx, Sy = ey * S0 - Cy;
				const float Sxx = ex * (ex * S0 - 2.f * Cx) + Cxx;
				const float Syy = ey * (ey * S0 - 2.f * Cy) + Cyy;
				const float Sxy = ex * (ey * S0 - Cy) - ey * Cx + Cxy;
				float* dst = grec + (size_t)s_id[w][lane] * GREC_FLOATS;
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
		__syncwarp();  // this warp's staging slice is rewritten next
		xg0 = yg0; xg1 = yg1; xg2 = yg2; xg3 = yg3; xkeep = ykeep; id_x = id_y;
		yg0 = zg0; yg1 = zg1;
		id_y = id_z; id_z = id_w;
	}
}

constexpr int BWD_NW = 2;

}  // namespace

template <bool EXT, int NW>
static void launch_fwd_nw(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is, const b200gs_outputs_t& out, cudaStream_t stream) {
	const int gx = (v.width + TILE_X - 1) / TILE_X, gy = (v.height + TILE_Y - 1) / TILE_Y;
	const unsigned units = (unsigned)(gx * gy * 8);
	launch_k(PDL_BLEND_FWD, blend_forward_pipe_kernel<EXT, NW>, dim3(units), dim3(32 * NW), stream, (const uint2*)is.ranges,
		(const uint32_t*)is.tile_order, (const uint32_t*)bs.sorted_vals, (const float4*)gs.rec, v.width, v.height, gx,
		v.background, is.final_T, is.n_contrib, out.color, EXT ? out.depth : nullptr, EXT ? out.alpha : nullptr, EXT ? out.feature : nullptr);
}
template <bool EXT, int NW>
static void launch_bwd_nw(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is, const b200gs_grad_outputs_t& gout, float* grec, cudaStream_t stream) {
	const int gx = (v.width + TILE_X - 1) / TILE_X, gy = (v.height + TILE_Y - 1) / TILE_Y;
	const unsigned units = (unsigned)(gx * gy * 8);
	// first kernel behind the gradient-record memset: plain stream ordering (launch_k_first)
	launch_k_first(blend_backward_pipe_kernel<EXT, NW>, dim3(units), dim3(32 * NW), stream, (const uint2*)is.ranges,
		(const uint32_t*)is.tile_order, (const uint32_t*)bs.sorted_vals, (const float4*)gs.rec, v.width, v.height, gx, v.background,
		(const float*)is.final_T, (const uint32_t*)is.n_contrib, gout.dL_dcolor, EXT ? gout.dL_ddepth : nullptr, EXT ? gout.dL_dalpha : nullptr,
		EXT ? gout.dL_dfeature : nullptr, grec);
}
static int env_nw(const char* name, int dflt) {
	const char* e = getenv(name);
	const int v = e ? atoi(e) : dflt;
	return (v == 1 || v == 2 || v == 4) ? v : dflt;
}

void launch_blend_forward_pipe(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                               const b200gs_outputs_t& out, cudaStream_t stream) {
	static const int nw = env_nw("B200GS_FWD_NW", FWD_NW);  // warps per pixel block (A/B measurements)
	if (v.extended) {
		if (nw == 1) launch_fwd_nw<true, 1>(v, gs, bs, is, out, stream);
		else if (nw == 2) launch_fwd_nw<true, 2>(v, gs, bs, is, out, stream);
		else launch_fwd_nw<true, 4>(v, gs, bs, is, out, stream);
	} else {
		if (nw == 1) launch_fwd_nw<false, 1>(v, gs, bs, is, out, stream);
		else if (nw == 2) launch_fwd_nw<false, 2>(v, gs, bs, is, out, stream);
		else launch_fwd_nw<false, 4>(v, gs, bs, is, out, stream);
	}
	count_launch();
}

void launch_blend_backward_pipe(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                                const b200gs_grad_outputs_t& gout, float* grec, cudaStream_t stream) {
	static const int nw = env_nw("B200GS_BWD_NW", BWD_NW);
	if (v.extended) {
		if (nw == 1) launch_bwd_nw<true, 1>(v, gs, bs, is, gout, grec, stream);
		else if (nw == 2) launch_bwd_nw<true, 2>(v, gs, bs, is, gout, grec, stream);
		else launch_bwd_nw<true, 4>(v, gs, bs, is, gout, grec, stream);
	} else {
		if (nw == 1) launch_bwd_nw<false, 1>(v, gs, bs, is, gout, grec, stream);
		else if (nw == 2) launch_bwd_nw<false, 2>(v, gs, bs, is, gout, grec, stream);
		else launch_bwd_nw<false, 4>(v, gs, bs, is, gout, grec, stream);
	}
	count_launch();
}
