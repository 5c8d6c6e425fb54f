// Host-side helpers shared by the C-ABI translation units: error reporting, TMA tensor-map encoding
// through the driver entry point (no link-time dependency on libcuda), tensor views.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

namespace wc {

// ---- error plumbing: every C-ABI entry returns 0 on success, non-zero otherwise; text via wc_last_error()
void set_error(const std::string& msg);
int fail(const std::string& msg);  // sets error, returns 1
#define WC_CHECK_CUDA(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::wc::fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                        std::to_string(__LINE__));                                                   \
  } while (0)
#define WC_REQUIRE(cond, msg)                                                                  \
  do {                                                                                         \
    if (!(cond)) return ::wc::fail(std::string("requirement failed: ") + #cond + " - " + msg); \
  } while (0)

// Every kernel launch of the library goes through WC_LAUNCH_CHECK: counts it and surfaces launch errors.
void count_launch();
long long launch_count();
#define WC_LAUNCH_CHECK()               \
  do {                                  \
    ::wc::count_launch();               \
    WC_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

// Optional per-launch timing with CUDA events on the launching stream (bench.py roofline leg).
enum ProfClass : int { kProfIgemm = 0, kProfAttention = 1, kProfGroupNorm = 2, kProfBoundaryConv = 3, kProfScheduler = 4,
                       kProfOther = 5, kProfNumClasses = 8 };
bool profiling_enabled();
void prof_begin_launch(int cls, cudaStream_t st, double work);
void prof_annotate(int a, int b, int c, int d, int e, int f);  // shape info attached to the most recent launch record
void prof_end_launch(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int cls, cudaStream_t s, double work) : st(s), on(profiling_enabled()) {
    if (on) prof_begin_launch(cls, st, work);
  }
  void note(int a, int b = 0, int c = 0, int d = 0, int e = 0, int f = 0) const {
    if (on) prof_annotate(a, b, c, d, e, f);
  }
  ~ProfScope() {
    if (on) prof_end_launch(st);
  }
};

int num_sms();

// Time-embedding factor table (unet_base.py:22-24: 10000 ** (arange(half) / half), fp32): the host evaluates the reference's
// torch expression once and hands the table over, so the embedding's argument t / factor is bit-identical to the reference's.
// time_factor_table(half) returns the device copy for the current device, or nullptr (the kernels then use powf).
int set_time_factor_table(const float* host_table, int half);
const float* time_factor_table(int half);
#ifdef __CUDACC__
__device__ __forceinline__ float time_factor(const float* table, int j, int half) {
  return table ? table[j] : powf(10000.0f, static_cast<float>(j) / static_cast<float>(half));
}
#endif
const char* last_error_cstr();

// NHWC bf16 activation view: element (b,y,x,c) at ptr[((b*H + y)*W + x)*ld + c]; ld >= C lets a view
// address a channel slice of a wider (concatenated) buffer.
struct Act {
  __nv_bfloat16* ptr = nullptr;
  int B = 0, H = 0, W = 0, C = 0, ld = 0;
  size_t pixels() const { return static_cast<size_t>(B) * H * W; }
};

// Encode a tiled bf16 tensor map.  dims/strides are innermost-first; strides in ELEMENTS for dims 1..rank-1
// (dim 0 is contiguous).  swizzle_bytes in {0,32,64,128}.  Returns 0 on success.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                     const uint32_t* box, uint32_t swizzle_bytes);

}  // namespace wc
