// Kernels of the optional attention blocks (reference: medical_diffusion/models/utils/attention_blocks.py):
// LayerNorm, GEGLU gate, multi-head softmax attention core, per-sample channel bias.  The 1x1 convolutions /
// Linear layers around them run on the tcgen05 convolution kernel (conv_tc with ksize 1).
#pragma once

#include "mf_common.cuh"

namespace mf {

// LayerNorm over the channel dimension of every token (GEGLU.norm, attention_blocks.py:14,22): split -> split
int layernorm_split(const __half* in, long long in_plane, const float* gamma, const float* beta, __half* out,
                    long long out_plane, long long tokens, int C, float eps, cudaStream_t s);
// GEGLU gate (attention_blocks.py:23-24): in raw [tokens][2*Ch] -> out split [tokens][Ch] = a * gelu(gate)
int geglu_split(const float* in, __half* out, long long out_plane, long long tokens, int Ch, cudaStream_t s);
// softmax((q*s)^T (k*s)) v per head (compute_attention, attention_blocks.py:35-43), s = d^-0.25.
// q/k/v: raw fp32 rows of `row_stride` floats per token, head h occupies channels [h*d, (h+1)*d).
// out: split [B*N][heads*d]
int attention_core(const float* q, const float* k, const float* v, int row_stride, __half* out, long long out_plane,
                   int B, int N, int heads, int d, cudaStream_t s);
extern int g_attn_tc;   // 1 (default): tcgen05 attention core for N in {64,128,192,256}, d in {64,128}; 0: CUDA-core kernel
// out = in + bias[n][c] (one-token cross attention collapses to this): split -> split
int add_channel_bias_split(const __half* in, long long in_plane, const float* bias, int bias_stride, __half* out,
                           long long out_plane, int N, int HW, int C, cudaStream_t s);

}  // namespace mf
