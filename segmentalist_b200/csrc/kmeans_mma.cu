// placeholder: replaced by the tcgen05 filter GEMM (see next commit)
#include "common.cuh"
extern "C" int64_t segb_mma_x_tiles_bytes(int64_t n_emb, int32_t D) { return 16; }
extern "C" int64_t segb_mma_w_tiles_bytes(int32_t K_max, int32_t D) { return 16; }
extern "C" int64_t segb_mma_cand_bytes(int64_t n_emb) { return 16; }
extern "C" int segb_mma_pack_x(const float *, int64_t, int32_t, void *, float *, void *) { segb::set_error("mma path not built"); return SEGB_E_UNSUPPORTED; }
extern "C" int segb_mma_pack_means(const float *, int32_t, int32_t, void *, float *, void *) { segb::set_error("mma path not built"); return SEGB_E_UNSUPPORTED; }
extern "C" int segb_mma_filter(const void *, const void *, int64_t, int32_t, int32_t, void *, void *) { segb::set_error("mma path not built"); return SEGB_E_UNSUPPORTED; }
extern "C" int segb_mma_refine(const segb_kmeans *, const void *, const float *, const float *, int64_t, float *, int32_t *, int64_t *, void *) { segb::set_error("mma path not built"); return SEGB_E_UNSUPPORTED; }
