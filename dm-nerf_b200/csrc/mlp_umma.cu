// tcgen05 MLP kernel -- placeholder until the probe results are in (see tools/umma_probe.cu).
#include "umma_api.cuh"

namespace dmnerf {
int umma_weights_pack(UmmaWeights& w, const NetParams& p, cudaStream_t) { w.ins_num = p.ins_num; w.ready = false; return 0; }
void umma_weights_free(UmmaWeights& w) { if (w.image) cudaFree(w.image); if (w.bias) cudaFree(w.bias); w = UmmaWeights(); }
bool umma_available(const UmmaWeights& w) { return w.ready; }
int launch_mlp_umma(const UmmaWeights&, const NetParams&, const float*, const float*, const float*, const float*, int64_t,
                    int, float*, cudaStream_t) {
  set_error("tcgen05 MLP kernel not available in this build");
  return 3;
}
}  // namespace dmnerf
