// Shared pieces of the embedding-encoder kernels (encoder_gemm.cu, encoder_attn.cu,
// encoder_rows.cu): SURVEY 8(f) rank 4, the step in front of the retrieval path.  The reference
// obtains every embedding from an Ollama server over HTTP (app/main.py:134-180,
// app/embedding_gen.py:143-190; model mxbai-embed-large = a BERT-large encoder, 24 layers,
// hidden 1024, 16 heads of 64, FFN 4096, GELU, post-LayerNorm, CLS pooling); here the encoder
// runs on the same B200, so queries and chunk embeddings are born in HBM.
//
// Activations are fp16 where they feed a tensor-core operand (the Ollama model file is F16 too)
// and fp32 on the residual stream and in every LayerNorm.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "sqe_k2.cuh"       // make_sw128_desc, encode_tiled_fn, sqe_ptx.cuh

namespace sqe {
namespace enc {

constexpr int kHidden = 1024;
constexpr int kHeads = 16;
constexpr int kHeadDim = 64;
constexpr int kChunkK = 64;                    // fp16 elements per 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kBM = 128;                       // rows of X per CTA = TMEM lanes

// [rows, cols] fp16 row-major matrix with `ld` elements between rows; box = 64 elements (128 B,
// one swizzle row) x box_rows; 128-B swizzle; out-of-range elements read as zeros.
static int make_map_2d(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld,
                       uint32_t box_rows) {
    EncodeTiledFn encf = encode_tiled_fn();
    if (!encf) { set_error("encoder: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kChunkK), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encf(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("encoder: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return -2; }
    return 0;
}

// instruction descriptor of tcgen05.mma kind::f16 with fp16 A and B (both K-major), fp32 D:
// D fp32 [4,6) = 1, A fmt [7,10) = 0 (F16), B fmt [10,13) = 0, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace enc
}  // namespace sqe
