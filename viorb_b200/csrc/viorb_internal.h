/* viorb_internal.h -- helpers shared by the translation units of libviorb_b200.so (not part of the ABI) */
#ifndef VIORB_INTERNAL_H
#define VIORB_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "viorb_gpu.h"

int viorb_fail(int code, const char* fmt, ...);
cudaStream_t viorb_ctx_stream(viorb_ctx* c);
int viorb_ctx_device(const viorb_ctx* c);
int viorb_ctx_bind(viorb_ctx* c);
/* per-context device scratch arena, grown on demand; contents are valid until the next call on the context */
int viorb_ctx_scratch(viorb_ctx* c, size_t bytes, uint8_t** out);
void viorb_ctx_add_launches(viorb_ctx* c, int n);
/* pinned host staging blocks (slot 0: searches, slot 1: frame-index builds) and the pool of frame-index device blocks */
int viorb_ctx_stage(viorb_ctx* c, int slot, size_t bytes, uint8_t** out);
int viorb_ctx_stage_mark(viorb_ctx* c, int slot);
int viorb_ctx_block_get(viorb_ctx* c, size_t bytes, uint8_t** out, size_t* got);
void viorb_ctx_block_put(viorb_ctx* c, uint8_t* p, size_t bytes);
viorb_ctx* viorb_extractor_ctx(viorb_extractor* e);

#define VCU(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return viorb_fail(VIORB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#endif
