// The width-128 decoder (src/variations/nrgbd.py:116-135) as a chain of 10 GEMM "layers" shared by the
// tensor-core builds (field_tc.cu: 3xTF32, field_bf.cu: 3xBF16):
//   0-4 forward  L1..L5:  h1 = relu(W1 f + b1), h2 = relu(W2 h1 + b2), [t; sdf] = W3 h2 + b3,
//                         hc = relu(W4 [t; f] + b4), rgb = sigmoid(W5 hc + b5)
//   5-9 dgrad    D5..D1:  the same weights transposed (reduction over the layer's outputs)
#pragma once
#include "field.cuh"

namespace pslam {
namespace declayers {
constexpr int kLayersFwd = 5, kLayersAll = 10;
// output columns N, reduction K of every layer (padded to the MMA granularity)
constexpr int hN[kLayersAll] = {128, 128, 144, 128, 16, 128, 144, 128, 128, 16};
constexpr int hK[kLayersAll] = {16, 128, 128, 144, 128, 16, 128, 144, 128, 128};
}  // namespace declayers

// source element of layer l at (output row n, reduction index k)
static __device__ __forceinline__ float tc_weight(const pslam_decoder_t &d, int l, int n, int k)
{
    switch (l) {
        case 0: return d.W1[n * 16 + k];
        case 1: return d.W2[n * 128 + k];
        case 2: return n < 128 ? d.W3[(1 + n) * 128 + k] : (n == 128 ? d.W3[k] : 0.0f);   // features first, sdf row at 128
        case 3: return d.W4[n * 144 + k];                                                  // k over [t(128); f(16)]
        case 4: return n < 3 ? d.W5[n * 128 + k] : 0.0f;
        // dgrad: B[n][k] = W[k][n] (reduction over the layer's outputs)
        case 5: return k < 3 ? d.W5[k * 128 + n] : 0.0f;                                   // g_hc[n] = sum_c g5[c] W5[c][n]
        case 6: return d.W4[k * 144 + n];                                                  // [g_t; g_f][n] = sum_k g_hc[k] W4[k][n]
        case 7: return k < 128 ? d.W3[(1 + k) * 128 + n] : (k == 128 ? d.W3[n] : 0.0f);    // g_h2[n] = sum_j g_o3[j] W3[j][n]
        case 8: return d.W2[k * 128 + n];
        case 9: return d.W1[k * 16 + n];                                                   // g_f[n] = sum_k g_h1[k] W1[k][n]
        default: return d.W4[k * 144 + 128 + n];                                           // (3xF16 stream, layer 10) g_f[n] += sum_k g_hc[k] W4[k][128 + n]
    }
}

static __device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace pslam
