// Host runtime glue: error string, launch counter, TMA tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda).
#include "common.cuh"
#include "../../include/upgpt_b200.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <mutex>

namespace upgpt {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

static int g_pdl = -1;      // -1: not decided yet (UPGPT_PDL=0 disables; upgpt_set_pdl overrides)
bool pdl_enabled() {
  if (g_pdl < 0) { const char* e = getenv("UPGPT_PDL"); g_pdl = (e && e[0] == '0') ? 0 : 1; }
  return g_pdl != 0;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_encodeTiled)fn;
}

int make_tmap(CUtensorMap* out, int elem_bytes, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, bool swizzle128) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) {
    set_last_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return -3;
  }
  if (((uintptr_t)base & 15) != 0) {
    set_last_error("tensor map base %p not 16-byte aligned", base);
    return -3;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) {
        set_last_error("tensor map stride[%d]=%llu not a multiple of 16 bytes", i, (unsigned long long)gstr[i - 1]);
        return -3;
      }
    }
    if (bx[i] == 0 || bx[i] > 256) {
      set_last_error("tensor map box[%d]=%u out of range", i, bx[i]);
      return -3;
    }
  }
  CUresult r = g_encode(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                        const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", (int)r,
                   rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                   (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0],
                   rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return -3;
  }
  return 0;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128) {
  return make_tmap(out, 2, base, rank, dims, strides_bytes, box, swizzle128);
}



// ---- auxiliary streams: the parallel branch of a forked program (e.g. a ResBlock's skip 1x1 GEMM next to conv1 + GroupNorm) ----
// One non-blocking stream per (device, index); fork / join are event edges, so they are captured into CUDA graphs as plain
// dependencies. Events come from a small per-process ring (an event only carries the dependency between its record and the wait
// issued right after it).
static constexpr int kAuxStreams = kStreamSlots - 1, kAuxDevices = 64, kEventRing = 64;
static cudaStream_t g_aux[kAuxDevices][kAuxStreams] = {};
static cudaEvent_t g_events[kAuxDevices][kEventRing] = {};
static int g_event_next[kAuxDevices] = {};
static std::mutex g_aux_mu;

static cudaStream_t aux_stream(int idx) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kAuxDevices || idx < 0 || idx >= kAuxStreams) return nullptr;
  std::lock_guard<std::mutex> lk(g_aux_mu);
  if (!g_aux[dev][idx] && cudaStreamCreateWithFlags(&g_aux[dev][idx], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  return g_aux[dev][idx];
}

bool is_aux_stream(cudaStream_t s) {
  if (!s) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kAuxDevices) return false;
  for (int i = 0; i < kAuxStreams; ++i) if (g_aux[dev][i] == s) return true;
  return false;
}

int stream_slot(cudaStream_t s) {
  if (!s) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kAuxDevices) return 0;
  for (int i = 0; i < kAuxStreams; ++i) if (g_aux[dev][i] == s) return 1 + i;
  return 0;
}

static int edge(cudaStream_t from, cudaStream_t to) {
  int dev = 0;
  UPGPT_CHECK_CUDA(cudaGetDevice(&dev));
  UPGPT_REQUIRE(dev >= 0 && dev < kAuxDevices, "stream edge: device index %d out of range", dev);
  cudaEvent_t ev;
  {
    std::lock_guard<std::mutex> lk(g_aux_mu);
    const int i = g_event_next[dev];
    g_event_next[dev] = (i + 1) % kEventRing;
    if (!g_events[dev][i]) UPGPT_CHECK_CUDA(cudaEventCreateWithFlags(&g_events[dev][i], cudaEventDisableTiming));
    ev = g_events[dev][i];
  }
  UPGPT_CHECK_CUDA(cudaEventRecord(ev, from));
  UPGPT_CHECK_CUDA(cudaStreamWaitEvent(to, ev, 0));
  return 0;
}

}  // namespace upgpt

extern "C" void* upgpt_aux_stream(int idx) { return (void*)upgpt::aux_stream(idx); }
extern "C" int upgpt_stream_fork(void* main_stream, int aux_idx) {
  cudaStream_t aux = upgpt::aux_stream(aux_idx);
  UPGPT_REQUIRE(aux, "stream_fork: no auxiliary stream %d", aux_idx);
  return upgpt::edge((cudaStream_t)main_stream, aux);
}
extern "C" int upgpt_stream_join(void* main_stream, int aux_idx) {
  cudaStream_t aux = upgpt::aux_stream(aux_idx);
  UPGPT_REQUIRE(aux, "stream_join: no auxiliary stream %d", aux_idx);
  return upgpt::edge(aux, (cudaStream_t)main_stream);
}

namespace upgpt {
}  // namespace upgpt

extern "C" int upgpt_set_pdl(int on) { upgpt::g_pdl = on ? 1 : 0; return 0; }
extern "C" const char* upgpt_last_error(void) { return upgpt::g_err; }
extern "C" int upgpt_abi_version(void) { return 1; }
extern "C" long long upgpt_launch_count(void) { return upgpt::g_launches.load(); }

// ---- in-graph launch trace (common.cuh: trace_stamp): one device pointer per translation unit ----
namespace upgpt {
int trace_set_attention(unsigned long long*); int trace_set_clip(unsigned long long*); int trace_set_misc(unsigned long long*);
int trace_set_norm(unsigned long long*); int trace_set_tc_gemm(unsigned long long*);
}
extern "C" int upgpt_trace_set(unsigned long long* buf) {
  int rc = 0;
  rc |= upgpt::trace_set_attention(buf); rc |= upgpt::trace_set_clip(buf); rc |= upgpt::trace_set_misc(buf);
  rc |= upgpt::trace_set_norm(buf); rc |= upgpt::trace_set_tc_gemm(buf);
  if (rc) { upgpt::set_last_error("upgpt_trace_set: cudaMemcpyToSymbol failed"); return -2; }
  return 0;
}
