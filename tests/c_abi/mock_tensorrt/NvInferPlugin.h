// MOCK of the TensorRT 8 plugin interfaces used by aloception_oss_b200/csrc/trt_plugin/msda_trt_plugin.cpp.
// TEST INFRASTRUCTURE: TensorRT is not installed in this image; this header restates, from the public TensorRT 8 API
// documentation, only the declarations that file needs (names, enumerators and virtual-method signatures of
// nvinfer1::IPluginV2 / IPluginV2Ext / IPluginV2IOExt / IPluginCreator, Dims, PluginTensorDesc, PluginFieldCollection and
// REGISTER_TENSORRT_PLUGIN), so that the plugin can be compiled and its logic exercised on the GPU box.  A real build uses
// the real <NvInferPlugin.h>.
#ifndef MOCK_NVINFER_PLUGIN_H_
#define MOCK_NVINFER_PLUGIN_H_

#include <cstddef>
#include <cstdint>

struct CUstream_st;
typedef CUstream_st* cudaStream_t;

namespace nvinfer1 {

enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4 };
enum class TensorFormat : int32_t { kLINEAR = 0, kCHW2 = 1, kHWC8 = 2, kCHW4 = 3, kCHW16 = 4, kCHW32 = 5 };

class Dims {
 public:
  static constexpr int32_t MAX_DIMS = 8;
  int32_t nbDims;
  int32_t d[MAX_DIMS];
};
class Dims2 : public Dims {
 public:
  Dims2(int32_t d0, int32_t d1) {
    nbDims = 2;
    d[0] = d0;
    d[1] = d1;
    for (int32_t i = 2; i < MAX_DIMS; ++i) d[i] = 0;
  }
};

struct PluginTensorDesc {
  Dims dims;
  DataType type;
  TensorFormat format;
  float scale;
};

enum class PluginFieldType : int32_t { kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5, kCHAR = 6, kDIMS = 7, kUNKNOWN = 8 };
struct PluginField {
  const char* name;
  const void* data;
  PluginFieldType type;
  int32_t length;
};
struct PluginFieldCollection {
  int32_t nbFields;
  const PluginField* fields;
};

class IPluginV2 {
 public:
  virtual const char* getPluginType() const noexcept = 0;
  virtual const char* getPluginVersion() const noexcept = 0;
  virtual int32_t getNbOutputs() const noexcept = 0;
  virtual Dims getOutputDimensions(int32_t index, const Dims* inputs, int32_t nbInputDims) noexcept = 0;
  virtual int32_t initialize() noexcept = 0;
  virtual void terminate() noexcept = 0;
  virtual size_t getWorkspaceSize(int32_t maxBatchSize) const noexcept = 0;
  virtual int32_t enqueue(int32_t batchSize, const void* const* inputs, void* const* outputs, void* workspace,
                          cudaStream_t stream) noexcept = 0;
  virtual size_t getSerializationSize() const noexcept = 0;
  virtual void serialize(void* buffer) const noexcept = 0;
  virtual void destroy() noexcept = 0;
  virtual void setPluginNamespace(const char* pluginNamespace) noexcept = 0;
  virtual const char* getPluginNamespace() const noexcept = 0;

 protected:
  virtual ~IPluginV2() noexcept = default;
};

class IPluginV2Ext : public IPluginV2 {
 public:
  virtual DataType getOutputDataType(int32_t index, const DataType* inputTypes, int32_t nbInputs) const noexcept = 0;
  virtual bool isOutputBroadcastAcrossBatch(int32_t outputIndex, const bool* inputIsBroadcasted, int32_t nbInputs) const noexcept = 0;
  virtual bool canBroadcastInputAcrossBatch(int32_t inputIndex) const noexcept = 0;
  virtual IPluginV2Ext* clone() const noexcept = 0;
};

class IPluginV2IOExt : public IPluginV2Ext {
 public:
  virtual void configurePlugin(const PluginTensorDesc* in, int32_t nbInput, const PluginTensorDesc* out, int32_t nbOutput) noexcept = 0;
  virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs, int32_t nbOutputs) const noexcept = 0;
};

class IPluginCreator {
 public:
  virtual const char* getPluginName() const noexcept = 0;
  virtual const char* getPluginVersion() const noexcept = 0;
  virtual const PluginFieldCollection* getFieldNames() noexcept = 0;
  virtual IPluginV2* createPlugin(const char* name, const PluginFieldCollection* fc) noexcept = 0;
  virtual IPluginV2* deserializePlugin(const char* name, const void* serialData, size_t serialLength) noexcept = 0;
  virtual void setPluginNamespace(const char* pluginNamespace) noexcept = 0;
  virtual const char* getPluginNamespace() const noexcept = 0;
  virtual ~IPluginCreator() = default;
};

}  // namespace nvinfer1

// the real macro instantiates a static registrar that hands the creator to TensorRT's plugin registry
#define REGISTER_TENSORRT_PLUGIN(name) static name mock_registered_##name {}

#endif  // MOCK_NVINFER_PLUGIN_H_
