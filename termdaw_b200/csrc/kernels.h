// kernels.h — launchers of the sm_100a kernels in kernels.cu (one launch = one kernel class at one
// DAG level for every task in the batch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tdg_types.h"

namespace tdg {

void set_launch_waves(uint32_t waves);
cudaError_t launch_fx(int mode, const FxTask* t, uint32_t n, uint32_t max_chunks, bool vec, bool exact, bool grouped, const uint8_t* blob, cudaStream_t s);
// clock table rows [row0, depth) of one (sr, bl); per-chunk index ranges of every event-driven vertex of a pass
cudaError_t launch_env_table(float* clock, uint32_t bl, uint32_t row0, uint32_t depth, float fsr, cudaStream_t s);
cudaError_t launch_chunk_ranges(const RangeTask* t, uint32_t n, uint32_t max_chunks, bool any_segs, const uint8_t* blob, cudaStream_t s);   // + the voice / adsr segment tables
cudaError_t launch_norm_prefix(const NormPrefixTask* t, uint32_t n, cudaStream_t s);
cudaError_t launch_norm_total(const NormPrefixTask* t, uint32_t n, float* reduce_buf, cudaStream_t s);
cudaError_t launch_bp_total(const BpTask* t, uint32_t n, const uint8_t* blob, cudaStream_t s);
cudaError_t launch_norm_scale(const NormScaleTask* t, uint32_t n, uint32_t max_chunks, bool vec, cudaStream_t s);
cudaError_t launch_sample_loop(const LoopTask* t, uint32_t n, uint32_t max_chunks, bool vec, cudaStream_t s);
cudaError_t launch_sample_multi(const MultiTask* t, uint32_t n, uint32_t max_chunks, bool vec, const uint8_t* blob, cudaStream_t s);
cudaError_t launch_sample_lerp(const LerpTask* t, uint32_t n, uint32_t max_chunks, bool vec, const uint8_t* blob, cudaStream_t s);
cudaError_t launch_voice(int kind, const VoiceTask* t, uint32_t n, uint32_t max_chunks, bool exact, bool libm_sin, const uint8_t* blob, cudaStream_t s);
// band pass = three launches: tile aggregates, carry scan, apply
cudaError_t launch_bp_aggregate(const BpTask* t, uint32_t n, uint32_t max_tiles, bool vec, const uint8_t* blob, cudaStream_t s);
cudaError_t launch_bp_carry(const BpTask* t, uint32_t n, const uint8_t* blob, cudaStream_t s);
cudaError_t launch_bp_apply(const BpTask* t, uint32_t n, uint32_t max_tiles, bool vec, const uint8_t* blob, cudaStream_t s);
// ... or one launch (single pass with a deterministic grouped look-back) when no exchange step sits in between
cudaError_t launch_bp_fused(const BpTask* t, uint32_t n, uint32_t max_tiles, bool vec, const uint8_t* blob, cudaStream_t s);
cudaError_t launch_quantise(const QuantTask* t, uint32_t n, uint64_t max_blocks, cudaStream_t s);
cudaError_t launch_resample(const ResampleTask* t, uint32_t n, uint64_t max_out, uint32_t max_chunks, const uint8_t* blob, cudaStream_t s);   // positions (k_resample_index), then the windows
cudaError_t launch_interleave(const float* l, const float* r, float* out_pairs, uint64_t n, cudaStream_t s);
cudaError_t launch_deinterleave(const float* lin, uint64_t n_frames, uint32_t stride, uint32_t off_l, uint32_t off_r,
                                uint32_t mix, float* out_l, float* out_r, cudaStream_t s);
cudaError_t launch_absmax(const float* x, uint64_t n, int* out_bits, cudaStream_t s);
cudaError_t launch_scale_by_inv(float* x, uint64_t n, const float* max_a, const float* max_b, cudaStream_t s);

}  // namespace tdg
