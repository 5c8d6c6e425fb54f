// Shared scaffolding of the model-level launch plans (seg.cu, srgan.cu): parameter table, bump allocator over the
// host-provided workspace, op list.
#pragma once
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "conv.cuh"

namespace wc {

class Bump {
 public:
  Bump() = default;
  Bump(void* base, size_t cap) : base_(static_cast<uint8_t*>(base)), cap_(cap) {}
  void* take(size_t bytes) {
    const size_t off = (off_ + 255) & ~static_cast<size_t>(255);
    off_ = off + bytes;
    if (!base_ || off_ > cap_) { overflow_ = true; return nullptr; }
    return base_ + off;
  }
  size_t used() const { return off_; }
  bool overflow() const { return overflow_; }
  static Bump dry() { return Bump(reinterpret_cast<void*>(static_cast<uintptr_t>(4096)), ~static_cast<size_t>(0) >> 1); }

 private:
  uint8_t* base_ = nullptr;
  size_t cap_ = 0, off_ = 0;
  bool overflow_ = false;
};

struct ParamTable {
  std::unordered_map<std::string, const float*> ptr;
  const float* get(const std::string& name, int* err) const {
    auto it = ptr.find(name);
    if (it == ptr.end()) {
      if (!*err) *err = fail("missing parameter '" + name + "'");
      return nullptr;
    }
    return it->second;
  }
};

using OpList = std::vector<std::function<int(cudaStream_t)>>;

inline Act make_act(Bump& bump, int B, int H, int W, int C) {
  Act a;
  a.B = B; a.H = H; a.W = W; a.C = C; a.ld = C;
  a.ptr = static_cast<__nv_bfloat16*>(bump.take(a.pixels() * C * sizeof(__nv_bfloat16)));
  return a;
}
inline Act slice_act(const Act& a, int c0, int C) {
  Act v = a;
  v.ptr = a.ptr ? a.ptr + c0 : nullptr;
  v.C = C;
  return v;
}

}  // namespace wc
