// msda_trt_plugin.cpp -- TensorRT plugin "MsDeformIm2ColTRT" (version "1") on the B200 operator.
//
// SURVEY.md section 8(f) row 2.  Replaces the reference's plugin library
// (alonet/torch2trt/plugins/ms_deform_im2col/sources/ms_deform_im2col_plugin.{h,cpp} + ms_deform_im2col_kernel.cu): same plugin
// name / version (the exporter's graph surgery renames `ms_deform_attn_forward` nodes to it, alonet/deformable_detr/
// trt_exporter.py:39-60), same five inputs without batch dimension
//     0 value (S, M, D)   1 spatial_shapes (L, 2) int32   2 level_start_index (L,) int32
//     3 sampling_loc (Lq, M, L, P, 2)   4 attn_weight (Lq, M, L, P)          ->  output (Lq, M * D)
// same serialized layout (seven int32 + the nvinfer1::DataType, ms_deform_im2col_plugin.cpp:120-142) so that engines built
// with either library deserialize with the other, float and half I/O.  enqueue() is ONE call into the C ABI
// (msda_im2col_inference, include/msda_b200.h) -- the kernels are the operator's own forward kernels.
//
// Build (needs the TensorRT >= 8 headers; not present in this image, where the file is compile- and run-checked against
// tests/c_abi/mock_tensorrt/NvInferPlugin.h, a minimal restatement of the interfaces used):
//   g++ -std=c++17 -O2 -fPIC -shared -I$TRT/include -I/usr/local/cuda/include -Iinclude msda_trt_plugin.cpp \
//       -Laloception_oss_b200 -lmsda_b200 -lnvinfer -o libmsda_trt_plugin.so
#include <NvInferPlugin.h>

#include <cstdint>
#include <cstring>
#include <new>
#include <string>

#include "msda_b200.h"

namespace msda_trt {

using namespace nvinfer1;

constexpr const char* kName = "MsDeformIm2ColTRT";
constexpr const char* kVersion = "1";

// What configurePlugin() learns from the tensor descriptors; (de)serialized field by field in this order.
struct Params {
  int32_t im2col_step = 64;  // kept for layout compatibility; the B200 launcher does not batch by it
  int32_t spatial_size = 0, num_heads = 0, channels = 0, num_levels = 0, num_query = 0, num_point = 0;
  DataType dtype = DataType::kFLOAT;
};
constexpr size_t kSerializedBytes = 7 * sizeof(int32_t) + sizeof(DataType);

class Plugin final : public IPluginV2IOExt {
 public:
  explicit Plugin(std::string name) : name_(std::move(name)) {}
  Plugin(std::string name, const void* data, size_t length) : name_(std::move(name)) {
    if (data != nullptr && length == kSerializedBytes) {
      const char* p = static_cast<const char*>(data);
      int32_t* ints[7] = {&p_.im2col_step, &p_.spatial_size, &p_.num_heads, &p_.channels, &p_.num_levels, &p_.num_query, &p_.num_point};
      for (int32_t* f : ints) { std::memcpy(f, p, sizeof(int32_t)); p += sizeof(int32_t); }
      std::memcpy(&p_.dtype, p, sizeof(DataType));
      valid_ = true;
    }
  }

  // ---- IPluginV2 ----
  const char* getPluginType() const noexcept override { return kName; }
  const char* getPluginVersion() const noexcept override { return kVersion; }
  int32_t getNbOutputs() const noexcept override { return 1; }
  Dims getOutputDimensions(int32_t, const Dims* inputs, int32_t) noexcept override {
    return Dims2(inputs[3].d[0], inputs[0].d[1] * inputs[0].d[2]);  // (Lq, M * D)
  }
  int32_t initialize() noexcept override { return 0; }
  void terminate() noexcept override {}
  size_t getWorkspaceSize(int32_t) const noexcept override { return 0; }
  int32_t enqueue(int32_t batchSize, const void* const* inputs, void* const* outputs, void*, cudaStream_t stream) noexcept override {
    if (!valid_) return 1;
    return msda_im2col_inference(stream, inputs[0], inputs[1], inputs[2], inputs[3], inputs[4], batchSize, p_.spatial_size,
                                 p_.num_heads, p_.channels, p_.num_levels, p_.num_query, p_.num_point, outputs[0],
                                 static_cast<int>(p_.dtype)) == 0 ? 0 : 1;
  }
  size_t getSerializationSize() const noexcept override { return kSerializedBytes; }
  void serialize(void* buffer) const noexcept override {
    char* p = static_cast<char*>(buffer);
    const int32_t ints[7] = {p_.im2col_step, p_.spatial_size, p_.num_heads, p_.channels, p_.num_levels, p_.num_query, p_.num_point};
    for (int32_t v : ints) { std::memcpy(p, &v, sizeof(int32_t)); p += sizeof(int32_t); }
    std::memcpy(p, &p_.dtype, sizeof(DataType));
  }
  void destroy() noexcept override { delete this; }
  void setPluginNamespace(const char* ns) noexcept override { namespace_ = ns ? ns : ""; }
  const char* getPluginNamespace() const noexcept override { return namespace_.c_str(); }

  // ---- IPluginV2Ext ----
  DataType getOutputDataType(int32_t, const DataType* inputTypes, int32_t) const noexcept override { return inputTypes[0]; }
  bool isOutputBroadcastAcrossBatch(int32_t, const bool*, int32_t) const noexcept override { return false; }
  bool canBroadcastInputAcrossBatch(int32_t) const noexcept override { return false; }
  IPluginV2Ext* clone() const noexcept override {
    Plugin* c = new (std::nothrow) Plugin(name_);
    if (c != nullptr) { c->p_ = p_; c->valid_ = valid_; c->namespace_ = namespace_; }
    return c;
  }

  // ---- IPluginV2IOExt ----
  void configurePlugin(const PluginTensorDesc* in, int32_t nbInput, const PluginTensorDesc* out, int32_t nbOutput) noexcept override {
    // value (S, M, D), spatial_shapes (L, 2), level_start_index (L,), sampling_loc (Lq, M, L, P, 2), attn_weight (Lq, M, L, P)
    valid_ = nbInput == 5 && nbOutput == 1 && out != nullptr && in[0].dims.nbDims == 3 && in[1].dims.nbDims == 2 &&
             in[2].dims.nbDims == 1 && in[3].dims.nbDims == 5 && in[4].dims.nbDims == 4;
    if (!valid_) return;
    p_.im2col_step = 64;
    p_.spatial_size = in[0].dims.d[0];
    p_.num_heads = in[0].dims.d[1];
    p_.channels = in[0].dims.d[2];
    p_.num_query = in[3].dims.d[0];
    p_.num_levels = in[3].dims.d[2];
    p_.num_point = in[3].dims.d[3];
    p_.dtype = in[0].type;
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t, int32_t) const noexcept override {
    if (inOut[pos].format != TensorFormat::kLINEAR) return false;
    if (pos == 1 || pos == 2) return inOut[pos].type == DataType::kINT32;  // level tensors
    const DataType t = inOut[pos].type;
    return (t == DataType::kFLOAT || t == DataType::kHALF) && t == inOut[0].type;  // one floating type for value, loc, attn, out
  }

 private:
  std::string name_, namespace_;
  Params p_;
  bool valid_ = false;
};

class Creator final : public IPluginCreator {
 public:
  Creator() { fields_.nbFields = 0; fields_.fields = nullptr; }  // the plugin takes no attributes: everything comes from the tensor shapes
  const char* getPluginName() const noexcept override { return kName; }
  const char* getPluginVersion() const noexcept override { return kVersion; }
  const PluginFieldCollection* getFieldNames() noexcept override { return &fields_; }
  IPluginV2* createPlugin(const char* name, const PluginFieldCollection*) noexcept override {
    return new (std::nothrow) Plugin(name ? name : "");
  }
  IPluginV2* deserializePlugin(const char* name, const void* data, size_t length) noexcept override {
    return new (std::nothrow) Plugin(name ? name : "", data, length);
  }
  void setPluginNamespace(const char* ns) noexcept override { namespace_ = ns ? ns : ""; }
  const char* getPluginNamespace() const noexcept override { return namespace_.c_str(); }

 private:
  PluginFieldCollection fields_;
  std::string namespace_;
};

REGISTER_TENSORRT_PLUGIN(Creator);

}  // namespace msda_trt

// test hook (plain C): the creator TensorRT's registry would hand out
extern "C" nvinfer1::IPluginCreator* msda_trt_plugin_creator(void) {
  static msda_trt::Creator creator;
  return &creator;
}
