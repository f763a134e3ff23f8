// b200gs -- warp-autonomous alpha blending (default blend path) + heaviest-first tile schedule.
//
// Same semantics and per-pair arithmetic as blend.cu (DGR/cuda_rasterizer/forward.cu:261-374,
// backward.cu:399-557; 8 blended channels for SDP-GS), different decomposition.  ncu on the
// CTA-per-tile kernels (profiles/r01_v1_blend_ncu_summary.csv) showed SMs active only 50-70 % of the
// kernel's duration and `barrier` as the top stall: a tile's 8 warps wait for the slowest one twice
// per batch, and a few deep tiles decide the kernel time.  Here the unit of work is one warp = one
// 8x4 pixel block (8 units per tile, launched as 32-thread CTAs so the hardware CTA scheduler
// load-balances them), units are issued heaviest tile first (longest-processing-time order from
// tile_schedule_kernel), and nothing ever waits on another warp:
//   * a round = 32 consecutive entries of the tile's sorted list, one per lane: the lane loads its
//     Gaussian's 32-byte geometry (g0,g1: one sector), runs the conservative block cull for it, and
//     only survivors fetch their payload (colour/depth/feature) and are staged in the warp's 2 KB of
//     shared memory, from where the per-pixel loop reads them as broadcasts;
//   * a unit whose 32 pixels are saturated (forward) or whose deepest contributor is reached
//     (backward) exits immediately instead of idling until the rest of the tile is done;
//   * backward: butterfly reduce-scatter over the warp, then ONE red.global.add.f32 instruction whose
//     13 active lanes hit the 13 consecutive floats of the Gaussian's 64-byte gradient record.
#include "common.cuh"
#include "blend_common.cuh"

namespace {

// ---- heaviest-first tile order -------------------------------------------------------------------
// key = 7 bits (position of the leading one and the next two bits of the tile's range length);
// counting sort by descending key in one CTA.  Order inside a bucket is arbitrary (units are independent).
__global__ void __launch_bounds__(1024) tile_schedule_kernel(const uint2* __restrict__ ranges, int tiles, uint32_t* __restrict__ order) {
	__shared__ uint32_t s_cnt[128];
	__shared__ uint32_t s_off[128];
	const int tid = threadIdx.x;
	if (tid < 128) s_cnt[tid] = 0;
	__syncthreads();
	auto key_of = [](uint2 r) -> uint32_t {
		const uint32_t n = r.y - r.x;
		if (n == 0) return 0u;
		const int msb = 31 - __clz(n);
		const uint32_t frac = msb >= 2 ? (n >> (msb - 2)) & 3u : (n << (2 - msb)) & 3u;
		return min(127u, (uint32_t)(msb + 1) * 4u + frac - 3u);
	};
	for (int t = tid; t < tiles; t += blockDim.x) atomicAdd(&s_cnt[key_of(ranges[t])], 1u);
	__syncthreads();
	if (tid == 0) {
		uint32_t run = 0;
		for (int k = 127; k >= 0; k--) { s_off[k] = run; run += s_cnt[k]; }
	}
	__syncthreads();
	for (int t = tid; t < tiles; t += blockDim.x) order[atomicAdd(&s_off[key_of(ranges[t])], 1u)] = (uint32_t)t;
}

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
	u.tile = order[unit >> 3];
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
	u.range = ranges[u.tile];
	return u;
}

template <bool EXT>
__global__ void __launch_bounds__(32) blend_forward_warp_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ point_list,
	const float4* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg, float* __restrict__ final_T,
	uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
	float* __restrict__ out_alpha, float* __restrict__ out_feat)
{
	__shared__ float4 s_g0[32], s_g1[32], s_g2[32];
	__shared__ float4 s_g3[EXT ? 32 : 1];
	const unsigned lane = threadIdx.x & 31;
	const Unit u = make_unit(order, ranges, W, H, grid_x);
	const int n = (int)(u.range.y - u.range.x);
	constexpr int NC = EXT ? 8 : 3;

	bool done = !u.inside;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C[NC];
#pragma unroll
	for (int ch = 0; ch < NC; ch++) C[ch] = 0.f;

	// Software pipeline over rounds of 32 list entries (one per lane), all in registers:
	//   ids are loaded 3 rounds ahead, geometry (g0,g1) 2 rounds ahead, cull + payload 1 round ahead,
	// so the only thing a round waits for is its own arithmetic -- the dependent chain
	// point_list -> record -> payload (three L2 round trips) is off the unit's critical path.
	const uint32_t NOID = 0xFFFFFFFFu;
	auto load_id = [&](int base) -> uint32_t {
		const int i = base + (int)lane;
		return (i >= 0 && i < n) ? __ldg(point_list + u.range.x + i) : NOID;
	};
	auto load_geo = [&](uint32_t id, float4& g0, float4& g1) {
		if (id != NOID) { const float4* r = rec + 4 * (size_t)id; g0 = __ldg(r); g1 = __ldg(r + 1); }
	};
	auto load_pay = [&](uint32_t id, bool keep, float4& g2, float4& g3) {
		if (keep) { const float4* r = rec + 4 * (size_t)id; g2 = __ldg(r + 2); if (EXT) g3 = __ldg(r + 3); }
	};
	float4 xg0, xg1, xg2, xg3, yg0, yg1;
	xg0 = xg1 = xg2 = xg3 = yg0 = yg1 = make_float4(0.f, 0.f, 0.f, 0.f);
	uint32_t id_x = load_id(0), id_y = load_id(32), id_z = load_id(64);
	load_geo(id_x, xg0, xg1);
	load_geo(id_y, yg0, yg1);
	bool xkeep = id_x != NOID && !cull_block(xg0, xg1, u.pb);
	load_pay(id_x, xkeep, xg2, xg3);

	for (int base = 0; base < n; base += 32) {
		if (__all_sync(0xFFFFFFFFu, done)) break;
		const uint32_t id_w = load_id(base + 96);                       // ids, round r+3
		float4 zg0 = make_float4(0.f, 0.f, 0.f, 0.f), zg1 = zg0;
		load_geo(id_z, zg0, zg1);                                       // geometry, round r+2
		const bool ykeep = id_y != NOID && !cull_block(yg0, yg1, u.pb);   // cull + payload, round r+1
		float4 yg2 = make_float4(0.f, 0.f, 0.f, 0.f), yg3 = yg2;
		load_pay(id_y, ykeep, yg2, yg3);

		unsigned mask = __ballot_sync(0xFFFFFFFFu, xkeep);              // consume round r
		if (mask != 0) {
			__syncwarp();  // readers of the previous round are done
			if (xkeep) {
				s_g0[lane] = xg0; s_g1[lane] = xg1; s_g2[lane] = xg2;
				if (EXT) s_g3[lane] = xg3;
			}
			__syncwarp();
		}
		while (mask) {
			// up to 4 survivors at a time: their alphas are independent (ILP), only the T recurrence is serial
			int jj[4];
			float al[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				jj[k] = mask ? (__ffs(mask) - 1) : -1;
				mask &= mask - 1;  // no-op once mask == 0
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				al[k] = 0.f;
				if (jj[k] >= 0) {
					const float4 a = s_g0[jj[k]];
					const float4 b = s_g1[jj[k]];
					const float dx = __fsub_rn(a.x, u.pxf), dy = __fsub_rn(a.y, u.pyf);
					const float power = pair_power(dx, dy, a.z, a.w, b.x);
					const float alpha = fminf(0.99f, __fmul_rn(b.y, expf(power)));
					if (!(power > 0.0f) && !(alpha < 1.0f / 255.0f)) al[k] = alpha;  // al == 0 <=> skipped pair
				}
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				if (jj[k] >= 0 && !done && al[k] != 0.f) {
					const float alpha = al[k];
					const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
					if (test_T < 0.0001f) {
						done = true;
					} else {
						const float4 c = s_g2[jj[k]];
						C[0] = __fmaf_rn(T, __fmul_rn(alpha, c.x), C[0]);
						C[1] = __fmaf_rn(T, __fmul_rn(alpha, c.y), C[1]);
						C[2] = __fmaf_rn(T, __fmul_rn(alpha, c.z), C[2]);
						if (EXT) {
							const float4 f = s_g3[jj[k]];
							C[3] = __fmaf_rn(T, __fmul_rn(alpha, c.w), C[3]);
							C[4] = __fmaf_rn(T, alpha, C[4]);
							C[5] = __fmaf_rn(T, __fmul_rn(alpha, f.x), C[5]);
							C[6] = __fmaf_rn(T, __fmul_rn(alpha, f.y), C[6]);
							C[7] = __fmaf_rn(T, __fmul_rn(alpha, f.z), C[7]);
						}
						T = test_T;
						last_contributor = (uint32_t)(base + jj[k] + 1);
					}
				}
			}
		}
		// rotate the pipeline registers
		xg0 = yg0; xg1 = yg1; xg2 = yg2; xg3 = yg3; xkeep = ykeep;
		yg0 = zg0; yg1 = zg1;
		id_y = id_z; id_z = id_w;
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

template <bool EXT>
__global__ void __launch_bounds__(32) blend_backward_warp_kernel(
	const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ point_list,
	const float4* __restrict__ rec, int W, int H, int grid_x, const float* __restrict__ bg,
	const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor,
	const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha_map, const float* __restrict__ dL_dfeat,
	float* __restrict__ grec)
{
	constexpr int NC = EXT ? 8 : 3;
	constexpr int NV = EXT ? 13 : 9;
	__shared__ float4 s_g0[32], s_g1[32], s_g2[32];
	__shared__ float4 s_g3[EXT ? 32 : 1];
	__shared__ uint32_t s_id[32];
	const unsigned lane = threadIdx.x & 31;
	const Unit u = make_unit(order, ranges, W, H, grid_x);
	const size_t pix = (size_t)u.py * W + u.px, HW = (size_t)H * W;

	const float T_final = u.inside ? final_T[pix] : 0.f;
	float T = T_final;
	const uint32_t last_contributor = u.inside ? n_contrib[pix] : 0u;
	uint32_t wmax = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, o));
	if (wmax == 0) return;  // nothing was blended into this block

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
	const float bg_dot_dpixel = __ldg(bg) * dpix[0] + __ldg(bg + 1) * dpix[1] + __ldg(bg + 2) * dpix[2];
	const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

	float accum_rec[NC], last_color[NC];
#pragma unroll
	for (int ch = 0; ch < NC; ch++) { accum_rec[ch] = 0.f; last_color[ch] = 0.f; }
	float last_alpha = 0.f;

	// list positions [0, wmax) back to front, 32 per round, with the same 3-deep register pipeline as the
	// forward (ids 3 rounds ahead, geometry 2, cull + payload 1)
	const uint32_t NOID = 0xFFFFFFFFu;
	auto load_id = [&](int base) -> uint32_t {
		const int i = base + (int)lane;
		return (base >= 0 && (uint32_t)i < wmax) ? __ldg(point_list + u.range.x + i) : NOID;
	};
	auto load_geo = [&](uint32_t id, float4& g0, float4& g1) {
		if (id != NOID) { const float4* r = rec + 4 * (size_t)id; g0 = __ldg(r); g1 = __ldg(r + 1); }
	};
	auto load_pay = [&](uint32_t id, bool keep, float4& g2, float4& g3) {
		if (keep) { const float4* r = rec + 4 * (size_t)id; g2 = __ldg(r + 2); if (EXT) g3 = __ldg(r + 3); }
	};
	const int base0 = (int)((wmax - 1) & ~31u);
	float4 xg0, xg1, xg2, xg3, yg0, yg1;
	xg0 = xg1 = xg2 = xg3 = yg0 = yg1 = make_float4(0.f, 0.f, 0.f, 0.f);
	uint32_t id_x = load_id(base0), id_y = load_id(base0 - 32), id_z = load_id(base0 - 64);
	load_geo(id_x, xg0, xg1);
	load_geo(id_y, yg0, yg1);
	bool xkeep = id_x != NOID && !cull_block(xg0, xg1, u.pb);
	load_pay(id_x, xkeep, xg2, xg3);

	for (int base = base0; base >= 0; base -= 32) {
		const uint32_t id_w = load_id(base - 96);
		float4 zg0 = make_float4(0.f, 0.f, 0.f, 0.f), zg1 = zg0;
		load_geo(id_z, zg0, zg1);
		const bool ykeep = id_y != NOID && !cull_block(yg0, yg1, u.pb);
		float4 yg2 = make_float4(0.f, 0.f, 0.f, 0.f), yg3 = yg2;
		load_pay(id_y, ykeep, yg2, yg3);

		unsigned mask = __ballot_sync(0xFFFFFFFFu, xkeep);
		if (mask != 0) {
			__syncwarp();
			if (xkeep) {
				s_g0[lane] = xg0; s_g1[lane] = xg1; s_g2[lane] = xg2; s_id[lane] = id_x;
				if (EXT) s_g3[lane] = xg3;
			}
			__syncwarp();
		}
		while (mask) {
			const int j = 31 - __clz(mask);
			mask &= ~(1u << j);
			const float4 a = s_g0[j];
			const float4 b = s_g1[j];
			const float dx = __fsub_rn(a.x, u.pxf), dy = __fsub_rn(a.y, u.pyf);
			const float power = pair_power(dx, dy, a.z, a.w, b.x);
			const float G = expf(power);
			const float alpha = fminf(0.99f, __fmul_rn(b.y, G));
			const bool active = ((uint32_t)(base + j) < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
			if (!__any_sync(0xFFFFFFFFu, active)) continue;
			float v[16];
#pragma unroll
			for (int k = 0; k < 16; k++) v[k] = 0.f;
			if (active) {
				T = T / (1.f - alpha);
				const float dchannel_dcolor = alpha * T;
				const float4 c = s_g2[j];
				float col[NC];
				col[0] = c.x; col[1] = c.y; col[2] = c.z;
				if (EXT) {
					const float4 f = s_g3[j];
					col[3] = c.w; col[4] = 1.0f; col[5] = f.x; col[6] = f.y; col[7] = f.z;
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
				const float dL_dG = b.y * dL_dalpha;  // min(0.99,.) clamp ignored in backward (backward.cu:538)
				const float gdx = G * dx, gdy = G * dy;
				const float dG_ddelx = -gdx * a.z - gdy * a.w;
				const float dG_ddely = -gdy * b.x - gdx * a.w;
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
			// lanes 0,2,..,2(NV-1) hold values 0..NV-1: one reduction instruction covering the 64-byte record
			if ((lane & 1) == 0 && (lane >> 1) < NV) atomicAdd(grec + (size_t)s_id[j] * GREC_FLOATS + (lane >> 1), tot);
		}
		xg0 = yg0; xg1 = yg1; xg2 = yg2; xg3 = yg3; xkeep = ykeep; id_x = id_y;
		yg0 = zg0; yg1 = zg1;
		id_y = id_z; id_z = id_w;
	}
}

bool use_tile_cta_path() {
	static int v = -1;
	if (v < 0) {
		const char* e = getenv("B200GS_BLEND");
		v = (e && e[0] == 't') ? 1 : 0;  // B200GS_BLEND=tile selects the CTA-per-tile kernels of blend.cu (A/B measurements)
	}
	return v == 1;
}

}  // namespace

void launch_blend_forward_tile(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                               const b200gs_outputs_t& out, cudaStream_t stream);
void launch_blend_backward_tile(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                                const b200gs_grad_outputs_t& gout, float* grec, cudaStream_t stream);

void launch_tile_schedule(const b200gs_view_t& v, ImageState& is, cudaStream_t stream) {
	if (use_tile_cta_path()) return;
	const int tiles = ((v.width + TILE_X - 1) / TILE_X) * ((v.height + TILE_Y - 1) / TILE_Y);
	tile_schedule_kernel<<<1, 1024, 0, stream>>>(is.ranges, tiles, is.tile_order);
	count_launch();
}

void launch_blend_forward(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                          const b200gs_outputs_t& out, cudaStream_t stream) {
	if (use_tile_cta_path()) return launch_blend_forward_tile(v, gs, bs, is, out, stream);
	const int gx = (v.width + TILE_X - 1) / TILE_X, gy = (v.height + TILE_Y - 1) / TILE_Y;
	const unsigned units = (unsigned)(gx * gy * 8);
	if (v.extended)
		blend_forward_warp_kernel<true><<<units, 32, 0, stream>>>(is.ranges, is.tile_order, bs.sorted_vals, gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, out.color, out.depth, out.alpha, out.feature);
	else
		blend_forward_warp_kernel<false><<<units, 32, 0, stream>>>(is.ranges, is.tile_order, bs.sorted_vals, gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, out.color, nullptr, nullptr, nullptr);
	count_launch();
}

void launch_blend_backward(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                           const b200gs_grad_outputs_t& gout, float* grec, cudaStream_t stream) {
	if (use_tile_cta_path()) return launch_blend_backward_tile(v, gs, bs, is, gout, grec, stream);
	const int gx = (v.width + TILE_X - 1) / TILE_X, gy = (v.height + TILE_Y - 1) / TILE_Y;
	const unsigned units = (unsigned)(gx * gy * 8);
	if (v.extended)
		blend_backward_warp_kernel<true><<<units, 32, 0, stream>>>(is.ranges, is.tile_order, bs.sorted_vals, gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, gout.dL_dcolor, gout.dL_ddepth, gout.dL_dalpha, gout.dL_dfeature, grec);
	else
		blend_backward_warp_kernel<false><<<units, 32, 0, stream>>>(is.ranges, is.tile_order, bs.sorted_vals, gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, gout.dL_dcolor, nullptr, nullptr, nullptr, grec);
	count_launch();
}
