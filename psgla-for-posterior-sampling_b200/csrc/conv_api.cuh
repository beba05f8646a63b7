// Internal entry points shared by the denoiser drivers (conv_tc.cu, conv_gemm.cu, drunet.cu); not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/psgla_b200.h"

namespace psgla {

// validates a psgla_next_pre (inpainting "pre" of the next iteration, fused into the last layer's epilogue)
int check_next_pre(const psgla_next_pre* next, const psgla_img_shape& shape);

// Direction in which the next layer of a denoiser walks its work items: consecutive layers alternate, so that each starts on
// the part of its input the previous layer wrote last (still in L2).  PSGLA_CONV_ALTERNATE=0 pins it to "forward".
int next_layer_direction();

// conv_tc.cu -- layers whose weights stay resident in shared memory
void pack_conv3x3_swizzled(const float* w, int nout_real, int cin_real, int nout_pad, int cin_pad, uint8_t* dst);
int conv64_hidden(const void* in, void* out, const uint8_t* w, const float* bias, int B, int H, int W, int relu,
                  const void* res1, const void* res2, cudaStream_t st);
int conv_first16(const void* in16, void* out, const uint8_t* w, const float* bias, int B, int H, int W, int relu,
                 cudaStream_t st);
int conv_last_post(const void* in, const uint8_t* w, const float* bias, int B, int H, int W, const float* base,
                   float base_scale, float gain, float w_old, float w_new, float* x_out, float* sample, float* mean,
                   float* mean2, const psgla_next_pre* next, cudaStream_t st);

// conv_gemm.cu -- layers with streamed weights (mode: 0 conv3x3, 1 2x2 stride-2 down, 2 2x2 transposed up)
int conv_gemm_layer(int mode, int B, int Hin, int Win, int Cin, int Cout, const void* w, const void* in, const void* res1,
                    const void* res2, void* out, int relu, cudaStream_t st);

}  // namespace psgla
