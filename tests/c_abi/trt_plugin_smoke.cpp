// Drives aloception_oss_b200/csrc/trt_plugin/msda_trt_plugin.cpp the way TensorRT would -- creator -> plugin ->
// supportsFormatCombination / configurePlugin -> serialize -> deserializePlugin -> clone -> enqueue -- against the MOCK
// TensorRT header (tests/c_abi/mock_tensorrt/NvInferPlugin.h; TensorRT itself is not in this image) and compares enqueue()'s
// output with msda_forward on the same device buffers (bit-equal: same kernels).  The recipe of the reference's plugin test
// (alonet/torch2trt/plugins/ms_deform_im2col/test.py:104-113): M=8, D=32, levels 64^2 ... 8^2, scaled-down query count.
#include <NvInferPlugin.h>
#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "msda_b200.h"

extern "C" nvinfer1::IPluginCreator* msda_trt_plugin_creator(void);
using namespace nvinfer1;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

static uint32_t rng = 2024u;
static float frand() { rng = rng * 1664525u + 1013904223u; return (float)(rng >> 8) / 16777216.0f; }

static PluginTensorDesc desc(std::initializer_list<int> dims, DataType t) {
  PluginTensorDesc d{};
  d.dims.nbDims = (int32_t)dims.size();
  int i = 0;
  for (int v : dims) d.dims.d[i++] = v;
  d.type = t;
  d.format = TensorFormat::kLINEAR;
  d.scale = 1.f;
  return d;
}

int main() {
  const int B = 2, M = 8, D = 32, L = 4, P = 4, Lq = 500;
  const int hw[L] = {64, 32, 16, 8};
  std::vector<int32_t> shapes(2 * L), start(L);
  int S = 0;
  for (int l = 0; l < L; ++l) { shapes[2 * l] = shapes[2 * l + 1] = hw[l]; start[l] = S; S += hw[l] * hw[l]; }
  const size_t n_value = (size_t)B * S * M * D, n_attn = (size_t)B * Lq * M * L * P, n_loc = 2 * n_attn, n_out = (size_t)B * Lq * M * D;
  std::vector<float> value(n_value), loc(n_loc), attn(n_attn);
  for (auto& v : value) v = frand() - 0.5f;
  for (auto& v : loc) v = frand() * 1.2f - 0.1f;
  for (auto& v : attn) v = frand() * 0.1f;
  float *d_value, *d_loc, *d_attn, *d_out, *d_ref;
  int32_t *d_shapes, *d_start;
  CK(cudaMalloc((void**)&d_value, n_value * 4)); CK(cudaMalloc((void**)&d_loc, n_loc * 4)); CK(cudaMalloc((void**)&d_attn, n_attn * 4));
  CK(cudaMalloc((void**)&d_out, n_out * 4)); CK(cudaMalloc((void**)&d_ref, n_out * 4));
  CK(cudaMalloc((void**)&d_shapes, shapes.size() * 4)); CK(cudaMalloc((void**)&d_start, start.size() * 4));
  CK(cudaMemcpy(d_value, value.data(), n_value * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_loc, loc.data(), n_loc * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_attn, attn.data(), n_attn * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_shapes, shapes.data(), shapes.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_start, start.data(), start.size() * 4, cudaMemcpyHostToDevice));

  IPluginCreator* creator = msda_trt_plugin_creator();
  REQUIRE(std::strcmp(creator->getPluginName(), "MsDeformIm2ColTRT") == 0 && std::strcmp(creator->getPluginVersion(), "1") == 0);
  REQUIRE(creator->getFieldNames()->nbFields == 0);
  auto* plugin = static_cast<IPluginV2IOExt*>(creator->createPlugin("layer0", creator->getFieldNames()));
  REQUIRE(plugin != nullptr && plugin->getNbOutputs() == 1);
  // an unconfigured plugin refuses to run
  {
    const void* in[5] = {d_value, d_shapes, d_start, d_loc, d_attn};
    void* out[1] = {d_out};
    REQUIRE(plugin->enqueue(B, in, out, nullptr, nullptr) != 0);
  }
  PluginTensorDesc io[6] = {desc({S, M, D}, DataType::kFLOAT), desc({L, 2}, DataType::kINT32), desc({L}, DataType::kINT32),
                            desc({Lq, M, L, P, 2}, DataType::kFLOAT), desc({Lq, M, L, P}, DataType::kFLOAT),
                            desc({Lq, M * D}, DataType::kFLOAT)};
  for (int pos = 0; pos < 6; ++pos) REQUIRE(plugin->supportsFormatCombination(pos, io, 5, 1));
  {
    PluginTensorDesc bad[6];
    std::memcpy(bad, io, sizeof(io));
    bad[1].type = DataType::kFLOAT;  // level tensors must be int32
    REQUIRE(!plugin->supportsFormatCombination(1, bad, 5, 1));
    std::memcpy(bad, io, sizeof(io));
    bad[3].type = DataType::kHALF;   // one floating type for all
    REQUIRE(!plugin->supportsFormatCombination(3, bad, 5, 1));
  }
  const Dims in_dims[5] = {io[0].dims, io[1].dims, io[2].dims, io[3].dims, io[4].dims};
  const Dims od = plugin->getOutputDimensions(0, in_dims, 5);
  REQUIRE(od.nbDims == 2 && od.d[0] == Lq && od.d[1] == M * D);
  const DataType in_types[5] = {DataType::kFLOAT, DataType::kINT32, DataType::kINT32, DataType::kFLOAT, DataType::kFLOAT};
  REQUIRE(plugin->getOutputDataType(0, in_types, 5) == DataType::kFLOAT);
  plugin->configurePlugin(io, 5, io + 5, 1);
  plugin->setPluginNamespace("ns");
  REQUIRE(plugin->initialize() == 0 && plugin->getWorkspaceSize(B) == 0);

  // serialize -> deserialize -> clone: the copy must run and agree
  REQUIRE(plugin->getSerializationSize() == 7 * sizeof(int32_t) + sizeof(DataType));
  std::vector<char> blob(plugin->getSerializationSize());
  plugin->serialize(blob.data());
  int32_t first[7];
  std::memcpy(first, blob.data(), sizeof(first));
  REQUIRE(first[0] == 64 && first[1] == S && first[2] == M && first[3] == D && first[4] == L && first[5] == Lq && first[6] == P);
  auto* restored = static_cast<IPluginV2IOExt*>(creator->deserializePlugin("layer0", blob.data(), blob.size()));
  REQUIRE(restored != nullptr);
  auto* copy = static_cast<IPluginV2IOExt*>(restored->clone());
  REQUIRE(copy != nullptr && std::strcmp(plugin->getPluginNamespace(), "ns") == 0);

  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));
  const void* in[5] = {d_value, d_shapes, d_start, d_loc, d_attn};
  void* out[1] = {d_out};
  CK(cudaMemset(d_out, 0xff, n_out * 4));
  REQUIRE(copy->enqueue(B, in, out, nullptr, stream) == 0);
  const msda_dims dims = {B, S, M, D, L, Lq, P};
  REQUIRE(msda_forward(d_value, d_shapes, d_start, d_loc, d_attn, d_ref, &dims, MSDA_F32, stream) == 0);
  CK(cudaStreamSynchronize(stream));
  std::vector<float> got(n_out), want(n_out);
  CK(cudaMemcpy(got.data(), d_out, n_out * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(want.data(), d_ref, n_out * 4, cudaMemcpyDeviceToHost));
  REQUIRE(std::memcmp(got.data(), want.data(), n_out * 4) == 0);
  double s = 0;
  for (float v : want) s += (double)v * v;
  REQUIRE(s > 0);
  std::printf("enqueue      ok (%zu values, bit-equal to msda_forward)\n", n_out);
  copy->terminate();
  copy->destroy();
  restored->destroy();
  plugin->destroy();
  std::printf("TRT plugin smoke: OK\n");
  return 0;
}
