// b200gs -- ordering stages: depth order of the Gaussians, instance offsets, duplicate-with-keys,
// tile sort, tile ranges.
//
// The reference sorts all L (Gaussian, tile) instances by a 64-bit key (tile << 32 | depth bits)
// with a stable LSD radix sort over 32 + getHigherMsb(tiles) bits (DGR/cuda_rasterizer/
// rasterizer_impl.cu:70-111, 277-318): about six 8-bit passes over 12-byte pairs.  An LSD sort
// on (tile, depth) IS a stable sort by depth followed by a stable sort by tile, so the same
// permutation is produced here in two cheaper steps:
//   1. stable radix sort of the P Gaussians by their depth bits (4 passes over P, not L, items);
//   2. instances emitted in that order (y-major, x-minor tiles per Gaussian, as the reference
//      does), then a stable radix sort by tile id only (ceil(bit/8) = 2 passes over 8-byte pairs).
// point_list, tile ranges and (reconstructed) 64-bit keys are bit-identical to the reference.
//
// The radix pass is a single-read "onesweep" pass: per-tile digit counts are chained between
// CTAs with decoupled look-back (one 32-bit status word per (tile, digit)), tile ids are handed
// out by an atomic ticket so a CTA only ever waits on CTAs that already started.
#include <cstddef>
#include "common.cuh"

namespace {

constexpr uint32_t FLAG_LOCAL = 1u << 30;   // word holds this tile's own count
constexpr uint32_t FLAG_INCL = 2u << 30;    // word holds the inclusive count over tiles [0..t]
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

__device__ __forceinline__ int64_t load_count(const unsigned long long* n_dev, int64_t n_max) {
	if (n_dev == nullptr) return n_max;
	unsigned long long n = *n_dev;
	return n < (unsigned long long)n_max ? (int64_t)n : n_max;
}

// Digit histograms of every pass in one read of the keys: hist[pass][256].
__global__ void __launch_bounds__(256) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n_max,
                                                         const unsigned long long* __restrict__ n_dev, int end_bit,
                                                         uint32_t* __restrict__ hist) {
	__shared__ uint32_t s_hist[4][256];
	const int passes = (end_bit + 7) / 8;
	for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&s_hist[0][0])[i] = 0;
	__syncthreads();
	const int64_t n = load_count(n_dev, n_max);
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const uint32_t k = keys[i];
		for (int p = 0; p < passes; p++) {
			const int bits = min(8, end_bit - 8 * p);
			atomicAdd(&s_hist[p][(k >> (8 * p)) & ((1u << bits) - 1)], 1u);
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) {
		const uint32_t c = (&s_hist[0][0])[i];
		if (c) atomicAdd(hist + i, c);
	}
}

template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS) onesweep_pass_kernel(
	const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
	uint32_t* __restrict__ vals_out, int64_t n_max, const unsigned long long* __restrict__ n_dev, int shift, int bits,
	const uint32_t* __restrict__ hist /*[256] this pass*/, uint32_t* __restrict__ lookback /*[tiles][256]*/,
	unsigned int* __restrict__ ticket)
{
	constexpr int TILE = SORT_THREADS * ITEMS;
	constexpr int WARPS = SORT_THREADS / 32;
	__shared__ uint32_t s_warp_hist[WARPS][256];
	__shared__ uint32_t s_local_off[256];
	__shared__ uint32_t s_digit_base[256];
	__shared__ uint32_t s_keys[TILE];
	__shared__ uint32_t s_vals[TILE];
	__shared__ uint32_t s_tile;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	for (int i = tid; i < WARPS * 256; i += SORT_THREADS) (&s_warp_hist[0][0])[i] = 0;
	__syncthreads();
	const uint32_t tile = s_tile;
	const int64_t n = load_count(n_dev, n_max);
	const int64_t tile_base = (int64_t)tile * TILE;
	if (tile_base >= n) return;
	const int tile_n = (int)min((int64_t)TILE, n - tile_base);
	const uint32_t dmask = (1u << bits) - 1;

	// 1. load (warp-striped: warp w owns a contiguous 32*ITEMS chunk) and rank within the warp
	uint32_t key[ITEMS], val[ITEMS], rnk[ITEMS];
	const int warp_base = warp * 32 * ITEMS;
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		const int li = warp_base + j * 32 + lane;
		const bool valid = li < tile_n;
		key[j] = valid ? keys_in[tile_base + li] : 0xFFFFFFFFu;
		val[j] = valid ? vals_in[tile_base + li] : 0u;
	}
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		const int li = warp_base + j * 32 + lane;
		const bool valid = li < tile_n;
		const uint32_t d = (key[j] >> shift) & dmask;
		const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? d : (0x80000000u | lane));
		const uint32_t before = __popc(peers & ((1u << lane) - 1));
		uint32_t pre = 0;
		if (valid) pre = s_warp_hist[warp][d];
		__syncwarp();
		if (valid && before == 0) s_warp_hist[warp][d] = pre + __popc(peers);
		__syncwarp();
		rnk[j] = pre + before;
	}
	__syncthreads();

	// 2. per digit (thread d): exclusive scan over warps, tile total, decoupled look-back
	uint32_t tile_count = 0;
	{
		const int d = tid;
#pragma unroll
		for (int w = 0; w < WARPS; w++) {
			const uint32_t t = s_warp_hist[w][d];
			s_warp_hist[w][d] = tile_count;
			tile_count += t;
		}
		uint32_t excl = 0;
		uint32_t* my = lookback + (size_t)tile * 256 + d;
		if (tile == 0) {
			st_volatile(my, tile_count | FLAG_INCL);
		} else {
			st_volatile(my, tile_count | FLAG_LOCAL);
			int64_t t = (int64_t)tile - 1;
			while (true) {
				uint32_t w;
				do { w = ld_volatile(lookback + (size_t)t * 256 + d); } while ((w & FLAG_MASK) == 0);
				excl += w & VALUE_MASK;
				if (w & FLAG_INCL) break;
				t--;
			}
			st_volatile(my, (excl + tile_count) | FLAG_INCL);
		}
		// exclusive scan of the global digit histogram and of the tile counts across the 256 digits
		uint32_t g = hist[d], l = tile_count;
		uint32_t gi = g, li = l;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, gi, o), b = __shfl_up_sync(0xFFFFFFFFu, li, o);
			if (lane >= o) { gi += a; li += b; }
		}
		__shared__ uint32_t s_g[WARPS], s_l[WARPS];
		if (lane == 31) { s_g[warp] = gi; s_l[warp] = li; }
		__syncthreads();
		uint32_t gw = 0, lw = 0;
		for (int w = 0; w < warp; w++) { gw += s_g[w]; lw += s_l[w]; }
		const uint32_t g_excl = gw + gi - g, l_excl = lw + li - l;
		s_local_off[d] = l_excl;
		s_digit_base[d] = g_excl + excl - l_excl;  // global position = s_digit_base[d] + local position
	}
	__syncthreads();

	// 3. reorder inside the tile so each digit's run leaves the CTA as one contiguous write
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		const int li = warp_base + j * 32 + lane;
		if (li < tile_n) {
			const uint32_t d = (key[j] >> shift) & dmask;
			const uint32_t pos = s_local_off[d] + s_warp_hist[warp][d] + rnk[j];
			s_keys[pos] = key[j];
			s_vals[pos] = val[j];
		}
	}
	__syncthreads();
	for (int i = tid; i < tile_n; i += SORT_THREADS) {
		const uint32_t k = s_keys[i];
		const uint32_t d = (k >> shift) & dmask;
		const uint32_t dst = s_digit_base[d] + (uint32_t)i;  // mod 2^32: s_digit_base may have wrapped
		keys_out[dst] = k;
		vals_out[dst] = s_vals[i];
	}
}

// Inclusive scan of tiles-touched in depth order (K2 of the reference, rasterizer_impl.cu:277, fused
// with the gather through `order`), single pass with decoupled look-back.
__global__ void __launch_bounds__(SCAN_THREADS) scan_offsets_kernel(
	const uint32_t* __restrict__ order, const ushort4* __restrict__ rect, uint32_t* __restrict__ offsets, int P,
	unsigned long long* __restrict__ state, GeomHeader* __restrict__ hdr)
{
	constexpr unsigned long long F_LOCAL = 1ull << 62, F_INCL = 2ull << 62, F_MASK = 3ull << 62;
	constexpr int TILE = SCAN_THREADS * SCAN_ITEMS;
	__shared__ uint32_t s_warp[SCAN_THREADS / 32];
	__shared__ unsigned long long s_prefix;
	__shared__ uint32_t s_tile;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(&hdr->scan_ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	const int base = tile * TILE + tid * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS];
	uint32_t sum = 0;
#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; j++) {
		const int i = base + j;
		uint32_t n = 0;
		if (i < P) {
			const ushort4 r = rect[order[i]];
			n = (uint32_t)(r.z - r.x) * (uint32_t)(r.w - r.y);
		}
		sum += n;
		v[j] = sum;  // inclusive within the thread
	}
	uint32_t inc = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, inc, o);
		if (lane >= o) inc += a;
	}
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	uint32_t wpre = 0, total = 0;
	for (int w = 0; w < SCAN_THREADS / 32; w++) {
		if (w < warp) wpre += s_warp[w];
		total += s_warp[w];
	}
	if (tid == 0) {
		unsigned long long excl = 0;
		volatile unsigned long long* st = state;
		if (tile == 0) {
			st[0] = (unsigned long long)total | F_INCL;
		} else {
			st[tile] = (unsigned long long)total | F_LOCAL;
			int64_t t = (int64_t)tile - 1;
			while (true) {
				unsigned long long w;
				do { w = st[t]; } while ((w & F_MASK) == 0);
				excl += w & ~F_MASK;
				if (w & F_INCL) break;
				t--;
			}
			st[tile] = (excl + total) | F_INCL;
		}
		s_prefix = excl;
		if ((int64_t)(tile + 1) * TILE >= P) hdr->num_rendered = excl + total;  // last tile
	}
	__syncthreads();
	const uint32_t thread_excl = (uint32_t)s_prefix + wpre + inc - sum;
#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; j++) {
		const int i = base + j;
		if (i < P) offsets[i] = thread_excl + v[j];
	}
}

// K3: duplicateWithKeys (rasterizer_impl.cu:70-111) in depth order.  The depth half of the
// reference's key is implied by the emission order, so only the tile id is written as sort key.
__global__ void __launch_bounds__(256) emit_instances_kernel(
	const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets, const ushort4* __restrict__ rect, int P,
	uint32_t grid_x, int64_t capacity, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, GeomHeader* __restrict__ hdr)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31;
	uint32_t g = 0, n = 0, start = 0;
	ushort4 r = make_ushort4(0, 0, 0, 0);
	if (i < P) {
		g = order[i];
		r = rect[g];
		n = (uint32_t)(r.z - r.x) * (uint32_t)(r.w - r.y);
		start = offsets[i] - n;
		if (i == P - 1 && (unsigned long long)offsets[i] > (unsigned long long)capacity) atomicOr(&hdr->overflow, 1u);
	}
	constexpr uint32_t SMALL = 12;
	if (n > 0 && n <= SMALL) {
		const uint32_t w = r.z - r.x;
		uint32_t x = r.x, y = r.y;
		for (uint32_t k = 0; k < n; k++) {
			const uint32_t pos = start + k;
			if ((int64_t)pos < capacity) { keys[pos] = y * grid_x + x; vals[pos] = g; }
			if (++x == r.z) { x = r.x; y++; }
		}
		(void)w;
	}
	// large footprints: the whole warp emits one Gaussian's tiles together
	uint32_t big = __ballot_sync(0xFFFFFFFFu, n > SMALL);
	while (big) {
		const int src = __ffs(big) - 1;
		big &= big - 1;
		const uint32_t bg = __shfl_sync(0xFFFFFFFFu, g, src), bn = __shfl_sync(0xFFFFFFFFu, n, src);
		const uint32_t bstart = __shfl_sync(0xFFFFFFFFu, start, src);
		const uint32_t bx0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)r.x, src), by0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)r.y, src);
		const uint32_t bw = __shfl_sync(0xFFFFFFFFu, (uint32_t)(r.z - r.x), src);
		for (uint32_t k = lane; k < bn; k += 32) {
			const uint32_t pos = bstart + k;
			if ((int64_t)pos < capacity) { keys[pos] = (by0 + k / bw) * grid_x + (bx0 + k % bw); vals[pos] = bg; }
		}
	}
}

// K5: identifyTileRanges (rasterizer_impl.cu:116-138); `ranges` zero-initialised by the caller (:310)
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ tile_keys, int64_t n_max,
                                                          const unsigned long long* __restrict__ n_dev, uint2* __restrict__ ranges) {
	const int64_t n = load_count(n_dev, n_max);
	const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= n) return;
	const uint32_t cur = tile_keys[idx];
	if (idx == 0) ranges[cur].x = 0;
	else {
		const uint32_t prev = tile_keys[idx - 1];
		if (cur != prev) { ranges[prev].y = (uint32_t)idx; ranges[cur].x = (uint32_t)idx; }
	}
	if (idx == n - 1) ranges[cur].y = (uint32_t)n;
}

__global__ void __launch_bounds__(256) debug_keys_kernel(const uint32_t* __restrict__ tile_keys, const uint32_t* __restrict__ point_list,
                                                         const float* __restrict__ depths, uint64_t* __restrict__ out, int64_t L) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= L) return;
	out[i] = ((uint64_t)tile_keys[i] << 32) | (uint64_t)__float_as_uint(depths[point_list[i]]);
}

}  // namespace

int launch_radix_sort(uint32_t* key_a, uint32_t* key_b, uint32_t* val_a, uint32_t* val_b, int64_t n_max,
                      const unsigned long long* n_dev, int end_bit, uint32_t* hist, uint32_t* lookback,
                      unsigned int* tickets, cudaStream_t stream) {
	if (n_max <= 0 || end_bit <= 0) return 0;
	const int passes = (end_bit + 7) / 8;
	const int items = sort_items_for(n_max);
	const int64_t tiles = sort_tiles_for(n_max);
	{
		int64_t blocks = (n_max + 256 * 8 - 1) / (256 * 8);
		if (blocks > 148 * 8) blocks = 148 * 8;
		radix_hist_kernel<<<(unsigned)blocks, 256, 0, stream>>>(key_a, n_max, n_dev, end_bit, hist);
		count_launch();
	}
	uint32_t *ki = key_a, *ko = key_b, *vi = val_a, *vo = val_b;
	for (int p = 0; p < passes; p++) {
		const int bits = end_bit - 8 * p < 8 ? end_bit - 8 * p : 8;
		uint32_t* lb = lookback + (size_t)p * tiles * 256;
		if (items == SORT_ITEMS_SMALL)
			onesweep_pass_kernel<SORT_ITEMS_SMALL><<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(
				ki, vi, ko, vo, n_max, n_dev, 8 * p, bits, hist + 256 * p, lb, tickets + p);
		else
			onesweep_pass_kernel<SORT_ITEMS_LARGE><<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(
				ki, vi, ko, vo, n_max, n_dev, 8 * p, bits, hist + 256 * p, lb, tickets + p);
		count_launch();
		uint32_t* t = ki; ki = ko; ko = t;
		t = vi; vi = vo; vo = t;
	}
	return passes & 1;
}

// Depth order of the P Gaussians + instance offsets.  After this: gs.order = ids by (depth bits, id),
// gs.offsets = inclusive scan of tiles touched in that order, hdr->num_rendered = total.
void launch_depth_order(GeomState& gs, int P, cudaStream_t stream) {
	if (P <= 0) return;
	// keys: key_a (written by preprocess), values: order (identity written by preprocess); 4 passes -> result in (key_a, order)
	unsigned int* tickets = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(gs.hdr) + offsetof(GeomHeader, sort_ticket));
	launch_radix_sort(gs.key_a, gs.key_b, gs.order, gs.val_b, P, nullptr, 32, gs.hist, gs.lookback, tickets, stream);
}

void launch_offsets_scan(GeomState& gs, int P, cudaStream_t stream) {
	if (P <= 0) return;
	scan_offsets_kernel<<<(unsigned)scan_tiles_for(P), SCAN_THREADS, 0, stream>>>(gs.order, gs.rect, gs.offsets, P, gs.scan_state, gs.hdr);
	count_launch();
}

void launch_emit(const b200gs_view_t& v, GeomState& gs, BinningState& bs, int P, int64_t capacity, cudaStream_t stream) {
	const uint32_t gx = (v.width + TILE_X - 1) / TILE_X;
	emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(gs.order, gs.offsets, gs.rect, P, gx, capacity, bs.key_a,
	                                                            bs.val_a, gs.hdr);
	count_launch();
}

void launch_tile_sort(const b200gs_view_t& v, GeomState& gs, BinningState& bs, int64_t capacity, cudaStream_t stream) {
	const uint32_t gx = (v.width + TILE_X - 1) / TILE_X, gy = (v.height + TILE_Y - 1) / TILE_Y;
	const int bit = (int)higher_msb(gx * gy);
	const unsigned long long* n_dev = &gs.hdr->num_rendered;
	unsigned int* tickets = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(gs.hdr) + offsetof(GeomHeader, sort_ticket)) + 4;
	const int where = launch_radix_sort(bs.key_a, bs.key_b, bs.val_a, bs.val_b, capacity, n_dev, bit, gs.hist + 4 * 256,
	                                    bs.lookback, tickets, stream);
	bs.sorted_keys = where ? bs.key_b : bs.key_a;
	bs.sorted_vals = where ? bs.val_b : bs.val_a;
}

void launch_tile_ranges(GeomState& gs, BinningState& bs, ImageState& is, int64_t capacity, cudaStream_t stream) {
	const unsigned long long* n_dev = &gs.hdr->num_rendered;
	if (capacity > 0) {
		tile_ranges_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, stream>>>(bs.sorted_keys, capacity, n_dev, is.ranges);
		count_launch();
	}
}

void launch_debug_keys(const b200gs_view_t& v, GeomState& gs, BinningState& bs, uint64_t* keys_out, int64_t L,
                       cudaStream_t stream) {
	(void)v;
	if (L <= 0) return;
	debug_keys_kernel<<<(unsigned)((L + 255) / 256), 256, 0, stream>>>(bs.sorted_keys, bs.sorted_vals, gs.depths, keys_out, L);
	count_launch();
}

