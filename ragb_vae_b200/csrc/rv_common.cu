// Error reporting, launch accounting and the per-category event profiler of librgbavae.
#include "rv_common.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace rv {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static std::atomic<int64_t> g_launches{0};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
struct ProfRec {
  int cat;
  cudaEvent_t e0, e1;
  double work;
};
static std::vector<ProfRec> g_prof;

LaunchScope::LaunchScope(int cat_, cudaStream_t stream_, double work) : cat(cat_), stream(stream_), timed(false) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on.load(std::memory_order_relaxed)) {
    if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) {
      timed = true;
      cudaEventRecord(e0, stream);
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof.push_back({cat, e0, e1, work});
    }
  }
}

LaunchScope::~LaunchScope() {
  if (timed) cudaEventRecord(e1, stream);
}

}  // namespace rv

extern "C" {

int rv_abi_version(void) { return RV_ABI_VERSION; }
const char* rv_last_error(void) { return rv::g_err; }
int64_t rv_launch_count(void) { return rv::g_launches.load(); }

int rv_prof_begin(void) {
  std::lock_guard<std::mutex> lk(rv::g_prof_mu);
  for (auto& r : rv::g_prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  rv::g_prof.clear();
  rv::g_prof_on.store(true);
  return 0;
}

int rv_prof_end(double* ms, int64_t* launches, double* work) {
  rv::g_prof_on.store(false);
  RV_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(rv::g_prof_mu);
  for (int c = 0; c < RV_PROF_CATEGORIES; ++c) {
    ms[c] = 0.0;
    launches[c] = 0;
    work[c] = 0.0;
  }
  for (auto& r : rv::g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess && r.cat >= 0 && r.cat < RV_PROF_CATEGORIES) {
      ms[r.cat] += t;
      launches[r.cat] += 1;
      work[r.cat] += r.work;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  rv::g_prof.clear();
  return 0;
}

}  // extern "C"
