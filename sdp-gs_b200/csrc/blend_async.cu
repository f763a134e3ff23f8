// b200gs -- warp-autonomous alpha blending with asynchronous (cp.async) list prefetch.
//
// Same unit of work, cull, survivor ring and per-pair arithmetic as blend_warp.cu (one warp = one 8x4 pixel
// block; DGR/cuda_rasterizer/forward.cu:261-374, backward.cu:399-557).  What changed is how a unit walks its
// tile's list.  Measured (B200, P=100k, 504x378): only ~8 % of a tile's entries survive a block's cull, so a
// round (32 entries) is ~110 warp instructions of walking plus 2-3 blended survivors -- yet the deepest unit
// needed ~1 250 cycles per round, i.e. one loaded L2 round trip: the register pipeline of blend_warp.cu
// (ids -> geometry -> payload, one round ahead each) cannot run faster than one dependent L2 access per round,
// and the kernel's duration is the walk of its longest list (3 700 entries = 116 rounds).  Splitting the serial
// recurrence over several warps (blend_pipe.cu) did not help for the same reason.
//
// Here every level of the dependent chain is fetched with cp.async into shared memory several rounds ahead
// (no registers are tied up by loads in flight):
//   ids       cp.async 4 B/lane   -> id ring,       DI rounds ahead
//   geometry  cp.async 2x16 B     -> raw ring,      D rounds ahead (address from the landed id)
//   payload   cp.async 2x16 B     -> survivor ring  (only for entries that survive the cull; consumed K+1 rounds later)
// One commit group per round, `cp.async.wait_group K` keeps K rounds of copies in flight.
#include "common.cuh"
#include "blend_common.cuh"

namespace {

#define NOID 0xFFFFFFFFu

__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sptr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sptr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int K_INFLIGHT = 2;             // commit groups (rounds) of copies allowed in flight
constexpr int GEO_AHEAD = K_INFLIGHT + 1; // geometry prefetch distance (rounds)
constexpr int ID_AHEAD = GEO_AHEAD + K_INFLIGHT + 1;
constexpr int NS_GEO = 4;                 // raw geometry stages (>= GEO_AHEAD + 1, power of two)
constexpr int NS_ID = 8;                  // id stages (>= ID_AHEAD + 1, power of two)
constexpr int RING = 128;                 // survivor slots: < 32 ready + (K+1) rounds whose payload is in flight + the round being staged
static_assert(NS_GEO >= GEO_AHEAD + 1 && NS_ID >= ID_AHEAD + 1, "ring depths");

struct AUnit {
	uint32_t tile;
	unsigned px, py;
	bool inside;
	float pxf, pyf;
	PixelBlock pb;
	uint2 range;
};

__device__ __forceinline__ AUnit make_aunit(const uint32_t* order, const uint2* ranges, int W, int H, int grid_x) {
	AUnit u;
	const unsigned lane = threadIdx.x & 31;
	const uint32_t unit = blockIdx.x;
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

// The list walker shared by forward and backward.  Entry i of the walk is list position pos(i); lanes handle
// entries 32r + lane of round r.  `FWD`: pos(i) = i, i < n.  Backward: the walk starts at the deepest
// contributor and runs towards the front: pos(i) = top - i.
template <bool EXT, bool FWD>
struct Walker {
	float4 (*raw0)[32];
	float4 (*raw1)[32];
	uint32_t (*rid)[32];
	float4 *g0, *g1, *g2, *g3;
	uint32_t* sid;  // backward only: Gaussian id per survivor slot
	const uint32_t* list;  // point_list + range.x
	const float4* rec;
	int n;      // entries in the walk
	int top;    // backward: list position of walk entry 0
	unsigned lane;

	__device__ __forceinline__ int pos(int i) const { return FWD ? i : top - i; }
	__device__ __forceinline__ void issue_ids(int r) {
		const int i = 32 * r + (int)lane;
		if (i < n) cp_async4(&rid[r & (NS_ID - 1)][lane], list + pos(i));
		else rid[r & (NS_ID - 1)][lane] = NOID;
	}
	__device__ __forceinline__ void issue_geo(int r) {
		const int i = 32 * r + (int)lane;
		if (i < n) {
			const uint32_t id = rid[r & (NS_ID - 1)][lane];
			const float4* src = rec + 4 * (size_t)id;
			cp_async16(&raw0[r & (NS_GEO - 1)][lane], src);
			cp_async16(&raw1[r & (NS_GEO - 1)][lane], src + 1);
		}
	}
	// cull round r, stage survivors at ring slots head.., request their payload.  Returns the number of survivors.
	__device__ __forceinline__ int cull_stage(int r, int head, const PixelBlock& pb) {
		const int i = 32 * r + (int)lane;
		bool keep = false;
		float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
		uint32_t id = NOID;
		if (i < n) {
			id = rid[r & (NS_ID - 1)][lane];
			a = raw0[r & (NS_GEO - 1)][lane];
			b = raw1[r & (NS_GEO - 1)][lane];
			keep = !cull_block(a, b, pb);
		}
		const unsigned mask = __ballot_sync(0xFFFFFFFFu, keep);
		if (keep) {
			// forward: list order; backward: the walk already runs deepest-first, so walk order
			const int slot = (head + __popc(mask & ((1u << lane) - 1u))) & (RING - 1);
			b.w = __uint_as_float((uint32_t)(pos(i) + 1));  // 1-based position in the tile's list
			g0[slot] = a; g1[slot] = b;
			if (!FWD) sid[slot] = id;
			const float4* src = rec + 4 * (size_t)id + 2;
			cp_async16(&g2[slot], src);
			if (EXT) cp_async16(&g3[slot], src + 1);
		}
		return __popc(mask);
	}
};

template <bool EXT>
__global__ void __launch_bounds__(32) blend_forward_async_kernel(
	const uint2* ranges, const uint32_t* order, const uint32_t* point_list, const float4* rec, int W, int H, int grid_x,
	const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
	float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha, float* __restrict__ out_feat)
{
	constexpr int NC = EXT ? 8 : 3;
	__shared__ float4 s_raw0[NS_GEO][32], s_raw1[NS_GEO][32];
	__shared__ uint32_t s_rid[NS_ID][32];
	__shared__ float4 s_g0[RING], s_g1[RING], s_g2[RING];
	__shared__ float4 s_g3[EXT ? RING : 1];
	const unsigned lane = threadIdx.x & 31;
	pdl_trigger();
	pdl_wait();
	const AUnit u = make_aunit(order, ranges, W, H, grid_x);
	Walker<EXT, true> wk;
	wk.raw0 = s_raw0; wk.raw1 = s_raw1; wk.rid = s_rid; wk.g0 = s_g0; wk.g1 = s_g1; wk.g2 = s_g2; wk.g3 = s_g3; wk.sid = nullptr;
	wk.list = point_list + u.range.x; wk.rec = rec; wk.n = (int)(u.range.y - u.range.x); wk.top = 0; wk.lane = lane;
	const int R = (wk.n + 31) >> 5;

	bool done = !u.inside;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C[NC];
#pragma unroll
	for (int ch = 0; ch < NC; ch++) C[ch] = 0.f;

	// blend `count` staged survivors starting at ring slot `start` (front to back).  Per group of 4: alphas (independent),
	// then the recurrence, branch-free: a pair that is not blended accumulates with weight 0 (exact for finite payloads),
	// so the loop-carried chain is one FMUL (T) and one predicate (done) per pair.
	auto process = [&](int start, int count) {
		for (int k0 = 0; k0 < count; k0 += 4) {
			if (__all_sync(0xFFFFFFFFu, done)) return;
			float al[4];
			float4 cc[4], ff[4];
			uint32_t ps[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int s = start + min(k0 + k, count - 1);  // past the end: re-read the last staged survivor (weight 0) -- never a stale slot, whose bits may be NaN
				const float4 a = s_g0[s];
				const float4 b = s_g1[s];
				cc[k] = s_g2[s];
				if (EXT) ff[k] = s_g3[s];
				const float dx = __fsub_rn(a.x, u.pxf), dy = __fsub_rn(a.y, u.pyf);
				const float power = pair_power(dx, dy, a.z, a.w, b.x);
				const float alpha = fminf(0.99f, __fmul_rn(b.y, expf(power)));
				al[k] = (k0 + k < count && !(power > 0.0f) && !(alpha < 1.0f / 255.0f)) ? alpha : 0.f;  // 0 <=> skipped pair
				ps[k] = __float_as_uint(b.w);
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const float alpha = al[k];
				const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
				const bool live = !done && alpha != 0.f;
				const bool stop = live && test_T < 0.0001f;
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
	};

	if (R > 0) {
		// prologue: ids of the first ID_AHEAD rounds, then geometry of the first GEO_AHEAD rounds
#pragma unroll
		for (int j = 0; j < ID_AHEAD; j++) wk.issue_ids(j);
		cp_commit();
		cp_wait<0>();
#pragma unroll
		for (int j = 0; j < GEO_AHEAD; j++) wk.issue_geo(j);
		cp_commit();
		cp_wait<0>();

		int head = 0, tail = 0;
		int h1 = 0, h2 = 0, h3 = 0;  // head after rounds r-1, r-2, r-3 (K_INFLIGHT + 1 = 3 back is what has landed)
		static_assert(K_INFLIGHT == 2, "head history below is written for K_INFLIGHT == 2");
		for (int r = 0; r < R; r++) {
			if (__all_sync(0xFFFFFFFFu, done)) break;
			wk.issue_ids(r + ID_AHEAD);
			cp_wait<K_INFLIGHT>();        // everything issued up to round r - K - 1 has landed: id(r + GEO_AHEAD), geometry(r), payload(<= r - K - 1)
			wk.issue_geo(r + GEO_AHEAD);
			const int ready = h3;         // survivors of rounds <= r - 3 have their payload
			while (ready - tail >= 32) {  // before staging: keeps the ring below < 32 + 3 rounds
				__syncwarp();
				process(tail & (RING - 1), 32);
				tail += 32;
				__syncwarp();
			}
			head += wk.cull_stage(r, head, u.pb);
			cp_commit();
			h3 = h2; h2 = h1; h1 = head;
		}
		cp_wait<0>();
		__syncwarp();
		while (head > tail && !__all_sync(0xFFFFFFFFu, done)) {
			const int c = min(32, head - tail);
			process(tail & (RING - 1), c);
			tail += c;
		}
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

}  // namespace

void launch_blend_forward_async(const b200gs_view_t& v, GeomState& gs, BinningState& bs, ImageState& is,
                                const b200gs_outputs_t& out, cudaStream_t stream) {
	const int gx = (v.width + TILE_X - 1) / TILE_X, gy = (v.height + TILE_Y - 1) / TILE_Y;
	const unsigned units = (unsigned)(gx * gy * 8);
	if (v.extended)
		launch_k(PDL_BLEND_FWD, blend_forward_async_kernel<true>, dim3(units), dim3(32), stream, (const uint2*)is.ranges,
			(const uint32_t*)is.tile_order, (const uint32_t*)bs.sorted_vals, (const float4*)gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, out.color, out.depth, out.alpha, out.feature);
	else
		launch_k(PDL_BLEND_FWD, blend_forward_async_kernel<false>, dim3(units), dim3(32), stream, (const uint2*)is.ranges,
			(const uint32_t*)is.tile_order, (const uint32_t*)bs.sorted_vals, (const float4*)gs.rec, v.width, v.height, gx,
			v.background, is.final_T, is.n_contrib, out.color, (float*)nullptr, (float*)nullptr, (float*)nullptr);
	count_launch();
}
