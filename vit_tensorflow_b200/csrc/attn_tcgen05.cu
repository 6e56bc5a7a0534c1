// tcgen05 multi-head attention (bf16) -- placeholder dispatch until the kernel lands in this file.
#include "attention.cuh"

namespace vb {

template <>
bool attention_fast<__nv_bfloat16>(const __nv_bfloat16*, int, const __nv_bfloat16*, int, const __nv_bfloat16*, int, __nv_bfloat16*,
                                   int, int, int, int, int, int, int, const float*, const float*, const float*, const float*,
                                   cudaStream_t) {
  return false;
}

}  // namespace vb
