// talkshow_b200 — shared host-side declarations (engine object, checkpoint lookup, workspace).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/talkshow_b200.h"

namespace ts {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

#define TS_CUDA(x)                                                                              \
  do {                                                                                          \
    cudaError_t _e = (x);                                                                       \
    if (_e != cudaSuccess) ts::fail(TS_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(_e)); \
  } while (0)

// ---- checkpoint view -----------------------------------------------------------------------
struct Ckpt {
  std::map<std::string, const ts_tensor*> m;
  Ckpt(const ts_tensor* t, int n) {
    for (int i = 0; i < n; ++i) m[t[i].name] = &t[i];
  }
  bool has(const std::string& k) const { return m.count(k) != 0; }
  // fp32 tensor with exactly this shape
  const float* f32(const std::string& k, std::initializer_list<int64_t> shape) const {
    auto it = m.find(k);
    if (it == m.end()) fail(TS_ERR_MISSING, "checkpoint tensor '%s' missing", k.c_str());
    const ts_tensor* t = it->second;
    if (t->dtype != 0) fail(TS_ERR_MISSING, "checkpoint tensor '%s' is not fp32", k.c_str());
    if (t->ndim != (int)shape.size()) fail(TS_ERR_MISSING, "checkpoint tensor '%s' has ndim %d", k.c_str(), t->ndim);
    int i = 0;
    for (int64_t d : shape) {
      if (t->shape[i] != d) fail(TS_ERR_MISSING, "checkpoint tensor '%s' dim %d is %lld, expected %lld", k.c_str(), i,
                                 (long long)t->shape[i], (long long)d);
      ++i;
    }
    return (const float*)t->data;
  }
  const ts_tensor* get(const std::string& k) const {
    auto it = m.find(k);
    if (it == m.end()) fail(TS_ERR_MISSING, "checkpoint tensor '%s' missing", k.c_str());
    return it->second;
  }
};

// ---- device buffer owned by the engine ------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) TS_CUDA(cudaFree(p));
    p = nullptr;
    bytes = 0;
    TS_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return (T*)p; }
};

// bump allocator over one DevBuf, reset at the start of every API call
struct Workspace {
  DevBuf buf;
  size_t off = 0;
  size_t need = 0;  // high-water mark of the sizing pass
  bool sizing = false;
  void begin_sizing() { sizing = true; off = 0; need = 0; }
  void begin(size_t) { sizing = false; off = 0; }
  template <class T>
  T* alloc(size_t n) {
    size_t b = (n * sizeof(T) + 255) & ~size_t(255);
    size_t o = off;
    off += b;
    if (off > need) need = off;
    if (sizing) return nullptr;
    if (off > buf.bytes) fail(TS_ERR_CUDA, "workspace overflow (%zu > %zu)", off, buf.bytes);
    return (T*)((char*)buf.p + o);
  }
};

// A conv/linear layer packed for the generic GEMM: W [N][K] row-major on the device, K = taps*Cin
// (tap-major, channel-minor, Cin padded to a multiple of 4), bias [N].
struct Layer {
  float* W = nullptr;
  float* W_hi = nullptr;  // 3xTF32 split copies for the tensor-core path (gemm_tc.cu), optional
  float* W_lo = nullptr;
  unsigned short* W_h16 = nullptr;   // fp16-split copies of W * 2^w_shift (power-of-two scale keeps the low plane normal)
  unsigned short* W_l16 = nullptr;
  float w_unscale = 1.f;             // 2^-w_shift, applied to the accumulator in the epilogue (exact)
  float* bias = nullptr;
  int N = 0, K = 0, taps = 1, cin = 0;  // cin = padded input channels
};

struct PixelPlan;  // pixelcnn.cu
struct ConvStacks; // convstack.cu
struct FaceNet;    // face.cu

}  // namespace ts

struct ts_engine {
  int device = 0;
  bool host_only = false;
  int sm_count = 0;
  std::string err;
  int64_t launches = 0;
  int pixel_mode = 0;
  int pixel_ctas = 0;        // > 0: persistent CTAs of the grid-wide PixelCNN plan built at the next load (default: one per SM)
  int pixel_fusion = 1;      // plan built at ts_load_pixelcnn: 0 plain 84-stage, 1 fused 52-stage, 2 fused + vert_to_horiz in the horizontal pass
  bool tc_pair = true;       // CTA-pair (cta_group::2) 256x256 tensor-core kernel (default)
  bool tc_onchip = false;    // experiment (mode 5): CTA-pair kernel takes plain fp32 operands and splits them hi / lo in shared memory
  bool tc_f16 = true;        // default: CTA-pair kernel on fp16-split operands (kind::f16, 3 products at twice the tf32 rate)
  bool tc_multicast = false;  // share operand boxes inside a thread-block cluster by TMA multicast
  bool tc_attr_set = false;  // cudaFuncSetAttribute(max dynamic smem) of the tcgen05 kernels done on this engine's device
  bool use_tc = true;  // dense contractions on the tcgen05 3xTF32 kernel when the geometry allows
  ts::PixelPlan* pix = nullptr;
  ts::ConvStacks* conv = nullptr;
  ts::FaceNet* face = nullptr;
  void* mfcc_tables = nullptr;  // ts::MfccTables (mfcc.cu)
  void* smplx = nullptr;        // ts::SmplxModel (lbs.cu)
  void* nccl = nullptr;         // ts::NcclApi (collective.cu): dlopen'ed NCCL + this engine's communicator
  ts::Workspace ws;
  // second stream of the fused body path (the two VQ decoders of a small batch run side by side) + its fork / join events
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  // largest batch whose two decoders run concurrently (ts_set_vq_parallel; 0 = one after the other).  Measured on a B200
  // (profiles/r02c_vq_decoders_side_by_side*.log): -0.6 ms per call at 1..16 samples (8-clip step 22.9 -> 22.3 ms), neutral at 32 / 64
  int vq_parallel_batch = 16;
  std::vector<void*> owned;  // device allocations that live as long as the engine
  // allocations of the weight sets, one slot per loadable module ("pixelcnn", "audioenc", "vq0", "vq1", "face"):
  // reloading a module frees the previous set (ts::LoadScope), a failed load frees its partial uploads
  std::map<std::string, std::vector<void*>> slot_mem;
  std::vector<void*>* alloc_sink = nullptr;
  float* upload(const std::vector<float>& h);
  unsigned short* upload(const std::vector<unsigned short>& h);
  void* dmalloc(size_t bytes);
};

namespace ts {
// Every entry point runs with the ENGINE's device current (allocations, tensor maps, launches) and restores the
// caller's device on exit: a process may hold one engine per GPU, or call torch.cuda.set_device between calls.
struct DeviceGuard {
  int prev = -1;
  bool active = false;
  explicit DeviceGuard(const ts_engine* e) {
    if (!e || e->host_only) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
    if (prev != e->device) {
      if (cudaSetDevice(e->device) != cudaSuccess) ts::fail(TS_ERR_CUDA, "cudaSetDevice(%d) failed", e->device);
      active = true;
    }
  }
  ~DeviceGuard() {
    if (active && prev >= 0) cudaSetDevice(prev);
  }
};
}  // namespace ts

namespace ts {
// Scope of one ts_load_*: every dmalloc/upload inside lands in the module's slot.  commit() (after the new module
// object has been installed) synchronises the device and frees the slot's previous allocations; leaving the scope
// without commit() (an exception part-way through) frees what this load had uploaded so far.
struct LoadScope {
  ts_engine* e;
  std::string slot;
  std::vector<void*> mem;
  bool done = false;
  LoadScope(ts_engine* e_, const char* slot_) : e(e_), slot(slot_) { e->alloc_sink = &mem; }
  void commit() {
    e->alloc_sink = nullptr;
    std::vector<void*>& old = e->slot_mem[slot];
    if (!old.empty() && !e->host_only) {
      cudaDeviceSynchronize();
      for (void* p : old) cudaFree(p);
    }
    old.swap(mem);
    mem.clear();
    done = true;
  }
  ~LoadScope() {
    e->alloc_sink = nullptr;
    if (!done && !e->host_only)
      for (void* p : mem) cudaFree(p);
  }
};
}  // namespace ts

namespace ts {
// entry points that launch kernels: a host-only (planning) engine has no device and no device copies of its weights
inline void require_device(const ts_engine* e) {
  if (e->host_only) fail(TS_ERR_UNSUPPORTED, "host-only engine cannot execute");
}
}  // namespace ts

// run `body`, convert exceptions to status codes
#define TS_API_BEGIN(e) \
  try {                 \
    if (!(e)) return TS_ERR_INVALID; \
    ts::DeviceGuard _ts_device_guard(e);
#define TS_API_END(e)                     \
  return TS_OK;                           \
  }                                       \
  catch (const ts::Error& ex) {           \
    (e)->err = ex.what();                 \
    return ex.code;                       \
  }                                       \
  catch (const std::exception& ex) {      \
    (e)->err = ex.what();                 \
    return TS_ERR_INVALID;                \
  }
