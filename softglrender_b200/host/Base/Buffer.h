// Host-side image container handed to Texture::setImageData (role of Base/Buffer.h:21-133 in the reference).
// RendererCUDA copies the texels to HBM at upload time, so only a linear host layout is needed here; Tiled and
// Morton storage exist on the device side (csrc/sgl_texture.h).
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>
#include "GLMInc.h"

namespace SoftGL {

enum BufferLayout { Layout_Linear, Layout_Tiled, Layout_Morton };

template<typename T>
class Buffer {
 public:
  static std::shared_ptr<Buffer<T>> makeDefault(size_t w, size_t h) {
    auto b = std::make_shared<Buffer<T>>();
    b->create(w, h);
    return b;
  }

  void create(size_t w, size_t h) {
    width_ = w;
    height_ = h;
    data_.assign(w * h, T());
  }

  BufferLayout getLayout() const { return Layout_Linear; }
  size_t getWidth() const { return width_; }
  size_t getHeight() const { return height_; }
  bool empty() const { return data_.empty(); }
  T *getRawDataPtr() { return data_.data(); }
  const T *getRawDataPtr() const { return data_.data(); }
  size_t getRawDataSize() const { return data_.size(); }
  size_t getRawDataBytesSize() const { return data_.size() * sizeof(T); }

  T *get(size_t x, size_t y) { return (x < width_ && y < height_) ? &data_[x + y * width_] : nullptr; }
  void set(size_t x, size_t y, const T &v) {
    if (x < width_ && y < height_) data_[x + y * width_] = v;
  }

 private:
  size_t width_ = 0, height_ = 0;
  std::vector<T> data_;
};

}  // namespace SoftGL
