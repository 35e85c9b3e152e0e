// Engine: model container behind the opaque msam_handle.
#pragma once
#include <initializer_list>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/msam_b200.h"
#include "kernels.h"
#include "tensormap.h"

namespace msam {

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct EncBlock {
  bool global = false;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  __nv_bfloat16 *qkv_w = nullptr, *proj_w = nullptr, *fc1_w = nullptr, *fc2_w = nullptr, *rel_table = nullptr;
  float *qkv_b = nullptr, *proj_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
};

struct EncoderWeights {
  __nv_bfloat16* patch_w = nullptr;
  float *patch_b = nullptr, *pos_embed = nullptr;
  std::vector<EncBlock> blocks;
  __nv_bfloat16 *neck_conv1 = nullptr, *neck_conv2 = nullptr;
  float *neck_ln1_g = nullptr, *neck_ln1_b = nullptr, *neck_ln2_g = nullptr, *neck_ln2_b = nullptr;
};

struct EncoderWorkspace {
  __nv_bfloat16 *patches = nullptr, *xn = nullptr, *xn_win = nullptr, *qkv = nullptr, *attn = nullptr, *hidden = nullptr;
  float* x = nullptr;
  float *neck1 = nullptr, *neck2 = nullptr;
  __nv_bfloat16 *neck1b = nullptr, *neck_col = nullptr;
};

struct DecoderState;  // decoder.cu

struct Engine {
  msam_config cfg{};
  int device = 0, num_sms = 148;
  bool finalized = false;
  std::unordered_map<std::string, HostTensor> host_weights;
  std::vector<void*> allocs;
  EncoderWeights enc;
  EncoderWorkspace ws;
  DecoderState* dec = nullptr;

  void* dalloc(size_t bytes, bool zero = false);
  const std::vector<float>* host(const std::string& name, std::initializer_list<int64_t> shape);
  __nv_bfloat16* upload_bf16(const float* src, size_t n);
  float* upload_f32(const float* src, size_t n);
  __nv_bfloat16* up_bf16(const std::string& name, std::initializer_list<int64_t> shape);
  float* up_f32(const std::string& name, std::initializer_list<int64_t> shape);

  int finalize_encoder();
  int alloc_encoder_ws();
  int finalize_decoder();  // decoder.cu
  int encode(const uint8_t* u8, const float* f32, int B, int hh, int ww, float* out, cudaStream_t st);
};

}  // namespace msam
