// Interface between encodec.cu (C ABI, weight store, CUDA-core kernels) and codec_tc.cu (tensor-core decoder).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/vcb200_codec.h"

namespace vcb {

struct TcCodec;

// 0: built; 1: this configuration is outside what the tensor-core path covers (why -> *reason, static string); -1: error.
int tc_codec_build(const enc_config& cfg, const std::map<std::string, float*>& w_dev,
                   const std::map<std::string, std::vector<int64_t>>& shapes, TcCodec** out, const char** reason);
// whether (B, T) can run here (T long enough for every reflect padding)
bool tc_codec_accepts(const TcCodec* c, int B, int T);
int tc_codec_decode(TcCodec* c, const int64_t* codes_dev, float* wav_dev, int B, int T, cudaStream_t st, int64_t* launches);
void tc_codec_destroy(TcCodec* c);
// per-layer device times of the last decode when VCB_CODEC_PROFILE=1 (name, ms), in launch order
const std::vector<std::pair<std::string, float>>& tc_codec_profile(const TcCodec* c);

// debug: tensor `name` of the last decoded chunk as fp32 [B][C][halo + T] (hi + lo); dims = {B, C, halo + T, halo}
int tc_codec_debug_tensor(TcCodec* c, const char* name, float* host_out, int64_t cap, int32_t* dims);

}  // namespace vcb
