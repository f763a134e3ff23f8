// b200gs -- C-ABI entry points, workspace layout and stage orchestration (see include/b200gs.h).
//
// Stage order mirrors CudaRasterizer::Rasterizer::forward/backward
// (DGR/cuda_rasterizer/rasterizer_impl.cu:198-336, 340-434); what differs is documented in
// binning.cu / blend.cu.  Nothing here allocates device memory or touches the legacy default
// stream: the caller owns the three byte workspaces and passes the stream.
#include <atomic>
#include <mutex>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- optional per-stage timing (CUDA events on the launch stream), used by bench.py for the roofline ----
enum Stage { ST_MEMSET = 0, ST_PREPROCESS, ST_DEPTH_SORT, ST_SCAN, ST_EMIT, ST_TILE_SORT, ST_RANGES, ST_BLEND_FWD,
             ST_BLEND_BWD, ST_PREPROCESS_BWD, ST_COUNT };
bool g_profile = false;
struct Pending { int stage; cudaEvent_t a, b; };
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_free_events;
std::mutex g_prof_mu;
double g_stage_ms[ST_COUNT] = {0};
long long g_stage_n[ST_COUNT] = {0};

cudaEvent_t get_event() {
	if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}
struct StageScope {
	cudaStream_t s; int stage; cudaEvent_t a;
	StageScope(cudaStream_t s_, int st) : s(s_), stage(st), a(nullptr) {
		if (g_profile) { std::lock_guard<std::mutex> l(g_prof_mu); a = get_event(); cudaEventRecord(a, s); }
	}
	~StageScope() {
		if (a) { std::lock_guard<std::mutex> l(g_prof_mu); cudaEvent_t b = get_event(); cudaEventRecord(b, s); g_pending.push_back({stage, a, b}); }
	}
};

template <typename T>
void carve(char*& p, T*& out, size_t count) {
	p = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p), 256));
	out = reinterpret_cast<T*>(p);
	p += count * sizeof(T);
}

int check_stage(const b200gs_view_t* v, cudaStream_t s, const char* what) {
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess && v->debug) e = cudaStreamSynchronize(s);  // CHECK_CUDA, auxiliary.h:166-173
	if (e != cudaSuccess) return fail(B200GS_E_CUDA, "[CUDA ERROR] in stage %s: %s", what, cudaGetErrorString(e));
	return 0;
}

}  // namespace

int train_fail(int code, const char* msg) { return fail(code, "%s", msg); }

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

unsigned pdl_mask() {
	static int v = -1;
	if (v < 0) {
		const char* e = getenv("B200GS_PDL");
		v = e ? (int)(strtoul(e, nullptr, 0) & 0xFFu) : 0xFF;
	}
	return (unsigned)v;
}

uint32_t higher_msb(uint32_t n) {
	uint32_t msb = sizeof(n) * 4, step = msb;
	while (step > 1) {
		step /= 2;
		if (n >> msb) msb += step; else msb -= step;
	}
	if (n >> msb) msb++;
	return msb;
}

// geom layout: [hdr | hist | depth look-back (4 passes) | scan state] (zeroed per forward) then the arrays
GeomState geom_from_chunk(char* base, int P) {
	GeomState g;
	char* p = base;
	const size_t n = (size_t)(P > 0 ? P : 1);
	carve(p, g.hdr, 1);
	carve(p, g.hist, 8 * 256);
	carve(p, g.lookback, (size_t)4 * sort_tiles_for(P) * 256);
	carve(p, g.scan_state, (size_t)scan_tiles_for(P) + 1);
	carve(p, g.depths, n);
	carve(p, g.rect, n);
	carve(p, g.rec, 4 * n);
	carve(p, g.clamped, n);
	carve(p, g.order, n);
	carve(p, g.key_a, n);
	carve(p, g.key_b, n);
	carve(p, g.val_b, n);
	g.bytes = align_up((size_t)(p - base), 256) + 256;
	return g;
}

static size_t geom_zero_len(int P) {
	GeomState g = geom_from_chunk(nullptr, P);
	return align_up(reinterpret_cast<size_t>(g.scan_state) + sizeof(unsigned long long) * ((size_t)scan_tiles_for(P) + 1), 256);
}

ImageState image_from_chunk(char* base, int W, int H) {
	ImageState s;
	char* p = base;
	const size_t N = (size_t)W * H, tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
	carve(p, s.ranges, tiles ? tiles : 1);
	carve(p, s.tile_order, tiles ? tiles : 1);
	carve(p, s.final_T, N ? N : 1);
	carve(p, s.n_contrib, N ? N : 1);
	carve(p, s.tile_count, (tiles ? tiles : 1) * (size_t)tile_count_stride());
	carve(p, s.tile_cost, 2 * (tiles ? tiles : 1));
	carve(p, s.tile_order_bwd, tiles ? tiles : 1);
	s.bytes = align_up((size_t)(p - base), 256) + 256;
	return s;
}

BinningState binning_from_chunk(char* base, int W, int H, int64_t capacity) {
	BinningState b;
	char* p = base;
	const size_t n = (size_t)(capacity > 0 ? capacity : 1);
	const size_t tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
	b.surv_words = n / 32 + tiles + 2;
	carve(p, b.lookback, (size_t)4 * sort_tiles_for(capacity) * 256);  // up to 4 tile passes (bit <= 32)
	carve(p, b.surv_bits, 8 * b.surv_words);
	carve(p, b.key_a, n);
	carve(p, b.key_b, n);
	carve(p, b.val_a, n);
	carve(p, b.val_b, n);
	b.sorted_keys = b.key_a;
	b.sorted_vals = b.val_a;
	b.bytes = align_up((size_t)(p - base), 256) + 256;
	return b;
}

static void resolve_sorted(const b200gs_view_t* v, BinningState& bs) {
	const uint32_t gx = (v->width + TILE_X - 1) / TILE_X, gy = (v->height + TILE_Y - 1) / TILE_Y;
	const int passes = ((int)higher_msb(gx * gy) + 7) / 8;
	bs.sorted_keys = (passes & 1) ? bs.key_b : bs.key_a;
	bs.sorted_vals = (passes & 1) ? bs.val_b : bs.val_a;
}

static int validate(const b200gs_view_t* v, const b200gs_gaussians_t* g, const b200gs_workspace_t* ws) {
	if (!v || !g || !ws) return fail(B200GS_E_ARG, "null argument");
	if (g->P < 0 || v->width <= 0 || v->height <= 0) return fail(B200GS_E_ARG, "bad sizes P=%d W=%d H=%d", g->P, v->width, v->height);
	if (((v->width + TILE_X - 1) / TILE_X) > 65535 || ((v->height + TILE_Y - 1) / TILE_Y) > 65535)
		return fail(B200GS_E_ARG, "image too large for 16-bit tile rects");
	if (g->P > 0) {
		if (!g->means3D || !g->opacities) return fail(B200GS_E_ARG, "means3D and opacities are required");
		if ((g->shs == nullptr) == (g->colors_precomp == nullptr))
			return fail(B200GS_E_ARG, "Please provide excatly one of either SHs or precomputed colors!");
		const bool sr = g->scales != nullptr && g->rotations != nullptr;
		if (sr == (g->cov3D_precomp != nullptr) || ((g->scales != nullptr) != (g->rotations != nullptr)))
			return fail(B200GS_E_ARG, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
		if (g->shs && v->sh_coeffs < (v->sh_degree + 1) * (v->sh_degree + 1))
			return fail(B200GS_E_ARG, "sh_coeffs=%d too small for sh_degree=%d", v->sh_coeffs, v->sh_degree);
		if (v->sh_degree < 0 || v->sh_degree > 3) return fail(B200GS_E_ARG, "sh_degree must be 0..3");
	}
	if (!v->background || !v->viewmatrix || !v->projmatrix || !v->campos) return fail(B200GS_E_ARG, "view pointers are required");
	if (!ws->geom || ws->geom_bytes < b200gs_geom_bytes(g->P)) return fail(B200GS_E_ARG, "geom workspace too small");
	if (!ws->image || ws->image_bytes < b200gs_image_bytes(v->width, v->height)) return fail(B200GS_E_ARG, "image workspace too small");
	return 0;
}

extern "C" {

int b200gs_version(void) { return B200GS_VERSION; }
const char* b200gs_last_error(void) { return g_err; }
int64_t b200gs_launch_count(void) { return g_launches.load(); }
void b200gs_profile_enable(int32_t on) { std::lock_guard<std::mutex> l(g_prof_mu); g_profile = on != 0; }
int b200gs_profile_read(double* ms_out, int64_t* n_out, int32_t reset) {
	std::lock_guard<std::mutex> l(g_prof_mu);
	for (auto& p : g_pending) {
		if (cudaEventSynchronize(p.b) != cudaSuccess) return fail(B200GS_E_CUDA, "profile event sync failed");
		float ms = 0.f;
		cudaEventElapsedTime(&ms, p.a, p.b);
		g_stage_ms[p.stage] += ms;
		g_stage_n[p.stage] += 1;
		g_free_events.push_back(p.a);
		g_free_events.push_back(p.b);
	}
	g_pending.clear();
	for (int i = 0; i < ST_COUNT; i++) {
		if (ms_out) ms_out[i] = g_stage_ms[i];
		if (n_out) n_out[i] = g_stage_n[i];
		if (reset) { g_stage_ms[i] = 0; g_stage_n[i] = 0; }
	}
	return 0;
}
void b200gs_abi_sizes(int64_t* out6) {
	out6[0] = sizeof(b200gs_view_t); out6[1] = sizeof(b200gs_gaussians_t); out6[2] = sizeof(b200gs_outputs_t);
	out6[3] = sizeof(b200gs_workspace_t); out6[4] = sizeof(b200gs_grad_outputs_t); out6[5] = sizeof(b200gs_grads_t);
}

size_t b200gs_geom_bytes(int32_t P) { return geom_from_chunk(nullptr, P).bytes; }
size_t b200gs_image_bytes(int32_t width, int32_t height) { return image_from_chunk(nullptr, width, height).bytes; }
size_t b200gs_binning_bytes(int64_t capacity, int32_t width, int32_t height) { return binning_from_chunk(nullptr, width, height, capacity).bytes; }
size_t b200gs_scratch_bytes(int32_t P) { return (size_t)(P > 0 ? P : 1) * GREC_FLOATS * sizeof(float); }

void b200gs_geom_layout(int32_t P, int64_t* off) {
	GeomState g = geom_from_chunk(nullptr, P);
	off[0] = (int64_t)reinterpret_cast<size_t>(g.hdr);
	off[1] = (int64_t)reinterpret_cast<size_t>(g.depths);
	off[2] = (int64_t)reinterpret_cast<size_t>(g.rect);
	off[3] = (int64_t)reinterpret_cast<size_t>(g.rec);
	off[4] = (int64_t)reinterpret_cast<size_t>(g.clamped);
	off[5] = (int64_t)reinterpret_cast<size_t>(g.order);
	off[6] = (int64_t)reinterpret_cast<size_t>(g.key_a);
}
void b200gs_image_layout(int32_t width, int32_t height, int64_t* off) {
	ImageState s = image_from_chunk(nullptr, width, height);
	off[0] = (int64_t)reinterpret_cast<size_t>(s.final_T);
	off[1] = (int64_t)reinterpret_cast<size_t>(s.n_contrib);
	off[2] = (int64_t)reinterpret_cast<size_t>(s.ranges);
}
void b200gs_binning_layout(int32_t width, int32_t height, int64_t capacity, int64_t* off) {
	BinningState b = binning_from_chunk(nullptr, width, height, capacity);
	b200gs_view_t v;
	memset(&v, 0, sizeof(v));
	v.width = width; v.height = height;
	resolve_sorted(&v, b);
	off[0] = (int64_t)reinterpret_cast<size_t>(b.sorted_vals);
	off[1] = (int64_t)reinterpret_cast<size_t>(b.sorted_keys);
}

int b200gs_workspace_init(const b200gs_workspace_t* ws, int32_t P, void* scratch, void* stream_) {
	if (!ws || !ws->geom || P < 0 || ws->geom_bytes < b200gs_geom_bytes(P)) return fail(B200GS_E_ARG, "workspace_init: bad arguments");
	cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
	cudaError_t e = cudaMemsetAsync(ws->geom, 0, geom_zero_len(P), stream);
	if (e == cudaSuccess && scratch) e = cudaMemsetAsync(scratch, 0, b200gs_scratch_bytes(P), stream);
	if (e == cudaSuccess && ws->image && ws->image_bytes) e = cudaMemsetAsync(ws->image, 0, ws->image_bytes, stream);  // no cost history yet
	if (e != cudaSuccess) return fail(B200GS_E_CUDA, "[CUDA ERROR] workspace_init: %s", cudaGetErrorString(e));
	return 0;
}

int b200gs_forward_preprocess(const b200gs_view_t* v, const b200gs_gaussians_t* g, const b200gs_outputs_t* out,
                              const b200gs_workspace_t* ws, void* stream_, int64_t* num_rendered_host) {
	if (int e = validate(v, g, ws)) return e;
	if (!out || (g->P > 0 && !out->radii)) return fail(B200GS_E_ARG, "radii output is required");
	cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
	const int P = g->P;
	char* gbase = reinterpret_cast<char*>(ws->geom);
	GeomState gs = geom_from_chunk(gbase, P);
	ImageState is = image_from_chunk(reinterpret_cast<char*>(ws->image), v->width, v->height);
	const size_t tiles = (size_t)((v->width + TILE_X - 1) / TILE_X) * ((v->height + TILE_Y - 1) / TILE_Y);
	if (!ws->persistent || P == 0) {  // persistent workspaces are left clean by the previous forward's blend kernel
		StageScope t(stream, ST_MEMSET);
		cudaMemsetAsync(gbase, 0, geom_zero_len(P), stream);
		if (P == 0) cudaMemsetAsync(is.ranges, 0, tiles * sizeof(uint2), stream);  // otherwise zeroed by the preprocess kernel (rasterizer_impl.cu:310)
	}
	if (P > 0) {
		{ StageScope t(stream, ST_PREPROCESS); launch_preprocess_forward(*v, *g, out->radii, gs, is, stream); }
		if (int e = check_stage(v, stream, "preprocess")) return e;
		{ StageScope t(stream, ST_DEPTH_SORT); launch_depth_order(gs, P, stream); }
		if (int e = check_stage(v, stream, "depth order")) return e;
	}
	if (num_rendered_host) {
		unsigned long long n = 0;
		cudaError_t e = cudaMemcpyAsync(&n, &gs.hdr->num_acc, sizeof(n), cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
		if (e != cudaSuccess) return fail(B200GS_E_CUDA, "[CUDA ERROR] reading num_rendered: %s", cudaGetErrorString(e));
		*num_rendered_host = (int64_t)n;
	}
	return 0;
}

static int forward_render_impl(const b200gs_view_t* v, const b200gs_gaussians_t* g, const b200gs_outputs_t* out,
                               const b200gs_workspace_t* ws, int64_t capacity, void* stream_, bool chained) {
	if (int e = validate(v, g, ws)) return e;
	if (!out || !out->color) return fail(B200GS_E_ARG, "color output is required");
	if (v->extended && (!out->depth || !out->alpha || !out->feature)) return fail(B200GS_E_ARG, "extended outputs are required");
	if (capacity < 0) return fail(B200GS_E_ARG, "negative capacity");
	if (capacity >= (1ll << 30)) return fail(B200GS_E_ARG, "capacity exceeds 2^30 instances (30-bit look-back counters)");
	if (!ws->binning || ws->binning_bytes < b200gs_binning_bytes(capacity, v->width, v->height)) return fail(B200GS_E_ARG, "binning workspace too small");
	cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
	const int P = g->P;
	GeomState gs = geom_from_chunk(reinterpret_cast<char*>(ws->geom), P);
	ImageState is = image_from_chunk(reinterpret_cast<char*>(ws->image), v->width, v->height);
	BinningState bs = binning_from_chunk(reinterpret_cast<char*>(ws->binning), v->width, v->height, capacity);
	resolve_sorted(v, bs);
	const size_t tiles = (size_t)((v->width + TILE_X - 1) / TILE_X) * ((v->height + TILE_Y - 1) / TILE_Y);
	if (P > 0 && capacity > 0) {
		{ StageScope t(stream, ST_EMIT); launch_scan_emit(*v, gs, bs, is, P, capacity, stream, chained, ws->persistent != 0); }
		if (int e = check_stage(v, stream, "instance offsets / duplicate-with-keys")) return e;
		{ StageScope t(stream, ST_TILE_SORT); launch_tile_sort(*v, gs, bs, capacity, stream); }
		if (int e = check_stage(v, stream, "tile sort")) return e;
	}
	const bool emitted = P > 0 && capacity > 0;
	if (P > 0 && !emitted) launch_finalize_header(gs, capacity, stream);  // the status words scan_emit assigns
	if (!emitted || !tile_counts_path((int)tiles)) {  // small tile grids: scan_emit's last CTA already built ranges + schedule
		{ StageScope t(stream, ST_RANGES); launch_tile_ranges(*v, gs, bs, is, emitted ? capacity : 0, stream, ws->persistent != 0); }
		if (int e = check_stage(v, stream, "tile ranges / schedule")) return e;
	}
	{
		StageScope t(stream, ST_BLEND_FWD);
		char* gbase = reinterpret_cast<char*>(ws->geom);
		launch_blend_forward(*v, gs, bs, is, *out, stream, reinterpret_cast<uint4*>(gbase + sizeof(GeomHeader)),
		                     (geom_zero_len(P) - sizeof(GeomHeader)) / sizeof(uint4), ws->persistent != 0);
	}
	return check_stage(v, stream, "blend forward");
}

int b200gs_forward_render(const b200gs_view_t* v, const b200gs_gaussians_t* g, const b200gs_outputs_t* out,
                          const b200gs_workspace_t* ws, int64_t capacity, void* stream) {
	return forward_render_impl(v, g, out, ws, capacity, stream, false);
}

int b200gs_forward(const b200gs_view_t* v, const b200gs_gaussians_t* g, const b200gs_outputs_t* out,
                   const b200gs_workspace_t* ws, int64_t capacity, void* stream) {
	if (int e = b200gs_forward_preprocess(v, g, out, ws, stream, nullptr)) return e;
	return forward_render_impl(v, g, out, ws, capacity, stream, g->P > 0);
}

int b200gs_forward_status(const b200gs_workspace_t* ws, void* stream_, int64_t* num_rendered, int32_t* overflow) {
	if (!ws || !ws->geom) return fail(B200GS_E_ARG, "null workspace");
	cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
	GeomHeader h;
	cudaError_t e = cudaMemcpyAsync(&h, ws->geom, sizeof(h), cudaMemcpyDeviceToHost, stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
	if (e != cudaSuccess) return fail(B200GS_E_CUDA, "[CUDA ERROR] reading status: %s", cudaGetErrorString(e));
	if (num_rendered) *num_rendered = (int64_t)h.num_rendered;
	if (overflow) *overflow = (int32_t)h.overflow;
	if (h.overflow & 2u) return fail(B200GS_E_ARG, "Point is filtered although prefiltered is set. This shouldn't happen!");
	if (h.overflow & 1u) return fail(B200GS_E_OVERFLOW, "num_rendered=%lld exceeds the binning capacity", (long long)h.num_rendered);
	return 0;
}

int b200gs_backward(const b200gs_view_t* v, const b200gs_gaussians_t* g, const int32_t* radii,
                    const b200gs_workspace_t* ws, int64_t capacity, const b200gs_grad_outputs_t* gout,
                    const b200gs_grads_t* grads, void* stream_) {
	if (int e = validate(v, g, ws)) return e;
	if (!gout || !grads) return fail(B200GS_E_ARG, "null gradient structs");
	const int P = g->P;
	if (P == 0) return 0;  // rasterize_points.cu:161
	if (!radii || !grads->scratch) return fail(B200GS_E_ARG, "radii and scratch are required");
	if (grads->scatter_bases) {
		if (grads->scatter_world < 2 || grads->scatter_rank < 0 || grads->scatter_rank >= grads->scatter_world || grads->scatter_shard_rows <= 0 ||
		    (grads->scatter_shard_rows & 127) || grads->scatter_shard_rows * grads->scatter_world < P)
			return fail(B200GS_E_ARG, "bad gradient scatter descriptor (shard_rows must be a multiple of 128 and cover P)");
		if (!g->shs || v->sh_coeffs != 16 || !g->scales || !g->rotations)
			return fail(B200GS_E_ARG, "gradient scatter needs shs with 16 coefficients and the scales/rotations path");
	}
	if (!ws->binning || ws->binning_bytes < b200gs_binning_bytes(capacity, v->width, v->height)) return fail(B200GS_E_ARG, "binning workspace too small");
	cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
	GeomState gs = geom_from_chunk(reinterpret_cast<char*>(ws->geom), P);
	ImageState is = image_from_chunk(reinterpret_cast<char*>(ws->image), v->width, v->height);
	BinningState bs = binning_from_chunk(reinterpret_cast<char*>(ws->binning), v->width, v->height, capacity);
	resolve_sorted(v, bs);
	float* grec = reinterpret_cast<float*>(grads->scratch);
	if (!ws->persistent) { StageScope t(stream, ST_MEMSET); cudaMemsetAsync(grec, 0, b200gs_scratch_bytes(P), stream); }
	{ StageScope t(stream, ST_BLEND_BWD); launch_blend_backward(*v, gs, bs, is, *gout, grec, stream, ws->persistent != 0); }
	if (int e = check_stage(v, stream, "blend backward")) return e;
	{ StageScope t(stream, ST_PREPROCESS_BWD); launch_preprocess_backward(*v, *g, radii, gs, grec, *grads, ws->persistent != 0, stream); }
	return check_stage(v, stream, "preprocess backward");
}

int b200gs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                        uint8_t* present, void* stream_) {
	(void)projmatrix;  // the reference computes p_proj but its frustum test only uses p_view.z (auxiliary.h:154)
	if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(B200GS_E_ARG, "bad arguments");
	if (P == 0) return 0;
	cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
	launch_mark_visible(P, means3D, viewmatrix, present, stream);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return fail(B200GS_E_CUDA, "[CUDA ERROR] mark_visible: %s", cudaGetErrorString(e));
	return 0;
}

int b200gs_debug_sorted_keys(const b200gs_view_t* v, int32_t P, const b200gs_workspace_t* ws, int64_t capacity,
                             uint64_t* keys_out, int64_t L, void* stream_) {
	if (!v || !ws || !keys_out || !ws->geom || !ws->binning) return fail(B200GS_E_ARG, "null argument");
	if (L > capacity) return fail(B200GS_E_ARG, "L exceeds capacity");
	cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
	GeomState gs = geom_from_chunk(reinterpret_cast<char*>(ws->geom), P);
	BinningState bs = binning_from_chunk(reinterpret_cast<char*>(ws->binning), v->width, v->height, capacity);
	resolve_sorted(v, bs);
	launch_debug_keys(*v, gs, bs, keys_out, L, stream);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return fail(B200GS_E_CUDA, "[CUDA ERROR] debug keys: %s", cudaGetErrorString(e));
	return 0;
}

}  // extern "C"
