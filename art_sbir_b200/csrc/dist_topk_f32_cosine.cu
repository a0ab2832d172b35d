// K1 instantiations for fp32 (kind::tf32) embeddings, cosine metric — see dist_topk_kernel.cuh.
#define SBIR_K1_INST_TF32 true
#define SBIR_K1_INST_METRIC SBIR_COSINE
#define SBIR_K1_INST_NAME k1_launch_f32_cosine
#include "dist_topk_kernel.cuh"
