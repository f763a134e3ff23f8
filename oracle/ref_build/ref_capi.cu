// ORACLE TOOLING ONLY -- not product code, never linked into libb200gs.so.
//
// Torch-free C-ABI over the UNMODIFIED reference rasterizer
// (CudaRasterizer::Rasterizer::{forward,backward,markVisible},
// DGR/cuda_rasterizer/rasterizer.h:31-84).  The reference's own binding
// (DGR/rasterize_points.cu:35-217) does the same three things through torch
// tensors and costs ~9 minutes of nvcc time because of torch/extension.h; this
// shim replaces only that binding so the reference kernels can be driven from
// ctypes with raw device pointers.  Built by oracle/Makefile into
// oracle/_ref/libref_rasterizer.so from the reference sources where they lie.
//
// Workspaces: the reference asks for its three byte buffers through
// std::function<char*(size_t)> callbacks (rasterizer_impl.cu:225-227, 238-240,
// 283-285).  Here the caller pre-allocates each buffer with the byte count the
// ref_required_* functions return; a callback that is asked for more than the
// capacity records an error and returns nullptr-free failure via return code.
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <cuda_runtime.h>
#include "rasterizer_impl.h"  // reference header: GeometryState/ImageState/BinningState, required<T>

using namespace CudaRasterizer;

namespace {
struct Arena {
	char* base;
	size_t cap;
	size_t asked;
	bool overflow;
};
std::function<char*(size_t)> arenaCallback(Arena* a) {
	return [a](size_t n) -> char* {
		a->asked = n;
		if (n > a->cap) { a->overflow = true; throw std::runtime_error("workspace too small"); }
		return a->base;
	};
}
char g_err[512] = "";
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err; }

size_t ref_required_geom(int P) { return required<GeometryState>(P); }
size_t ref_required_image(int N) { return required<ImageState>(N); }
size_t ref_required_binning(int R) { return required<BinningState>(R); }

// Byte offsets (from `base`) of every decodable sub-array, so Python does not
// have to re-derive the 128-byte bump allocation (rasterizer_impl.h:21-27).
// geom: depths, clamped, internal_radii, means2D, cov3D, conic_opacity, rgb,
//       tiles_touched, point_offsets           (9 entries)
void ref_geom_layout(char* base, int P, int64_t* off) {
	char* chunk = base;
	GeometryState g = GeometryState::fromChunk(chunk, P);
	off[0] = (char*)g.depths - base;
	off[1] = (char*)g.clamped - base;
	off[2] = (char*)g.internal_radii - base;
	off[3] = (char*)g.means2D - base;
	off[4] = (char*)g.cov3D - base;
	off[5] = (char*)g.conic_opacity - base;
	off[6] = (char*)g.rgb - base;
	off[7] = (char*)g.tiles_touched - base;
	off[8] = (char*)g.point_offsets - base;
}
// image: accum_alpha (final T), n_contrib, ranges   (3 entries)
void ref_image_layout(char* base, int N, int64_t* off) {
	char* chunk = base;
	ImageState s = ImageState::fromChunk(chunk, N);
	off[0] = (char*)s.accum_alpha - base;
	off[1] = (char*)s.n_contrib - base;
	off[2] = (char*)s.ranges - base;
}
// binning: point_list, point_list_unsorted, point_list_keys, point_list_keys_unsorted (4 entries)
void ref_binning_layout(char* base, int R, int64_t* off) {
	char* chunk = base;
	BinningState b = BinningState::fromChunk(chunk, R);
	off[0] = (char*)b.point_list - base;
	off[1] = (char*)b.point_list_unsorted - base;
	off[2] = (char*)b.point_list_keys - base;
	off[3] = (char*)b.point_list_keys_unsorted - base;
}

// Returns num_rendered (>= 0) or -1 on error.  If `binning_cap` is too small the
// call fails with -2 and *binning_needed holds the byte count to allocate
// (two-call protocol; the reference sizes this buffer after its own D2H sync).
int ref_forward(
	int P, int D, int M,
	const float* background, int width, int height,
	const float* means3D, const float* shs, const float* colors_precomp,
	const float* opacities, const float* scales, float scale_modifier,
	const float* rotations, const float* cov3D_precomp,
	const float* viewmatrix, const float* projmatrix, const float* cam_pos,
	float tan_fovx, float tan_fovy, int prefiltered,
	float* out_color, int* radii,
	char* geom_buf, size_t geom_cap,
	char* binning_buf, size_t binning_cap, size_t* binning_needed,
	char* image_buf, size_t image_cap,
	int debug)
{
	Arena g{geom_buf, geom_cap, 0, false}, b{binning_buf, binning_cap, 0, false}, i{image_buf, image_cap, 0, false};
	try {
		int n = Rasterizer::forward(
			arenaCallback(&g), arenaCallback(&b), arenaCallback(&i),
			P, D, M, background, width, height, means3D, shs, colors_precomp, opacities,
			scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
			tan_fovx, tan_fovy, prefiltered != 0, out_color, radii, debug != 0);
		if (binning_needed) *binning_needed = b.asked;
		return n;
	} catch (const std::exception& e) {
		if (binning_needed) *binning_needed = b.asked;
		snprintf(g_err, sizeof(g_err), "%s", e.what());
		return b.overflow ? -2 : -1;
	}
}

int ref_backward(
	int P, int D, int M, int R,
	const float* background, int width, int height,
	const float* means3D, const float* shs, const float* colors_precomp,
	const float* scales, float scale_modifier, const float* rotations,
	const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
	const float* campos, float tan_fovx, float tan_fovy, const int* radii,
	char* geom_buf, char* binning_buf, char* image_buf,
	const float* dL_dpix,
	float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
	float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
	int debug)
{
	try {
		Rasterizer::backward(P, D, M, R, background, width, height, means3D, shs, colors_precomp,
			scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
			tan_fovx, tan_fovy, radii, geom_buf, binning_buf, image_buf, dL_dpix,
			dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
			dL_dscale, dL_drot, debug != 0);
		return 0;
	} catch (const std::exception& e) {
		snprintf(g_err, sizeof(g_err), "%s", e.what());
		return -1;
	}
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
	Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
	return 0;
}

}  // extern "C"
