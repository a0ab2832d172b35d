// K1 instantiations for bf16 (kind::f16) embeddings, cosine metric — see dist_topk_kernel.cuh.
#define SBIR_K1_INST_TF32 false
#define SBIR_K1_INST_METRIC SBIR_COSINE
#define SBIR_K1_INST_NAME k1_launch_bf16_cosine
#include "dist_topk_kernel.cuh"
