/* b200gs -- the one collective of the image-parallel training path: SUM of the fused
 * per-Gaussian gradient buffer over the ranks of one NVSwitch box, written as a kernel over peer-mapped memory
 * (NVLink 5 P2P loads / stores, or NVLS multimem when the buffer has a multicast mapping) instead of a library call.
 *
 * The reference has no distributed path (single process, cuda:0); this is what "the per-Gaussian gradients are
 * combined ... before the Adam step" (BASELINE.json north_star) costs on the wire: 62 floats per Gaussian.
 *
 * Every rank calls the function with the same n / grid on buffers that were allocated symmetrically (same offset in
 * every rank's peer-mapped allocation, e.g. torch.distributed._symmetric_memory).  Two-shot algorithm: rank r sums
 * slice r of all ranks' buffers in rank order 0..world-1 (bit-identical result everywhere) and writes it back to
 * every rank.  `flags` is a zero-initialised symmetric u32 array of b200gs_allreduce_flag_words(world) words used for
 * the two block-level barriers: one-way epoch flags (a rank stores the launch's epoch into its slot on every peer and
 * spins on its own memory; the epoch counter lives in the same array, so the call can be captured in a CUDA graph and
 * replayed).  All ranks must issue the same sequence of collective calls on a given flag array. */
#ifndef B200GS_COLLECTIVE_H_
#define B200GS_COLLECTIVE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

size_t b200gs_allreduce_flag_words(int32_t world);

/* buffers_dev / flags_dev: DEVICE arrays [world] of peer-mapped base pointers (entry `rank` is the local one).
 * offset_floats, n_floats: the region to reduce (both multiples of 4).  multicast_ptr: NVLS multicast mapping of the
 * same allocation or NULL (then plain P2P loads / stores are used). */
int b200gs_allreduce_sum_f32(void* const* buffers_dev, void* const* flags_dev, void* multicast_ptr, int64_t offset_floats,
                             int64_t n_floats, int32_t rank, int32_t world, void* stream);

/* Second half of the fused gradient exchange.  The first half happens inside b200gs_backward (b200gs_grads_t.scatter_*):
 * every rank s has pushed its gradient rows for the Gaussians owned by rank o into o's staging buffer,
 *     staging_o[s][c * Ps + li * w + k]   (floats; s = source rank, Ps = shard_rows, li = row within the shard,
 *                                          (c, w) = segment offset / width in floats per Gaussian: xyz (0,3), shs (3,48),
 *                                          opacity (51,1), scaling (52,3), rotation (55,4), feature (59,3); 64 * Ps floats per source).
 * This call: barrier (all pushes have landed), rank r sums its shard over the sources in rank order (bit-identical
 * result everywhere), writes the 62 reduced floats per Gaussian into EVERY rank's fused gradient buffer
 *     out_r[c * Pp + (o * Ps + li) * w + k],  Pp = world * Ps
 * over NVLink, barrier.  Same result as an all-reduce of the fused buffer, but the wire carries each byte once per
 * direction and half of it is hidden under the backward kernel.  `flags` as for b200gs_allreduce_sum_f32. */
int b200gs_gather_reduce_f32(void* const* staging_dev, void* const* out_dev, void* const* flags_dev, int64_t shard_rows,
                             int32_t rank, int32_t world, int32_t chained /* != 0: the previous operation on `stream` is
                             b200gs_backward (the kernel may then be launched programmatically behind it) */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GS_COLLECTIVE_H_ */
