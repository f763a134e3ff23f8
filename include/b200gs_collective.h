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
 * the two block-level barriers (self-resetting, so the call can be captured in a CUDA graph and replayed). */
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

#ifdef __cplusplus
}
#endif
#endif /* B200GS_COLLECTIVE_H_ */
