// Library-wide state of the C ABI: version, thread-local error string, launch counter.
#include <stdarg.h>
#include <vector>
#include "common.cuh"
#include "tc_common.cuh"

namespace op3 {
static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- profiling -------------------------------------------------------------------------------------------
bool g_profile_on = false;
namespace {
struct ProfEntry { cudaEvent_t a, b; int kid; double work; };
std::vector<ProfEntry> g_entries;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
void profile_begin(cudaStream_t st, int kid, double work) {
  ProfEntry en{get_event(), get_event(), kid, work};
  cudaEventRecord(en.a, st);
  g_entries.push_back(en);
}
void profile_end(cudaStream_t st) { if (!g_entries.empty()) cudaEventRecord(g_entries.back().b, st); }
}  // namespace op3

extern "C" int op3_profile_enable(int on) {
  for (auto& e : op3::g_entries) { op3::g_pool.push_back(e.a); op3::g_pool.push_back(e.b); }
  op3::g_entries.clear();
  op3::g_profile_on = on != 0;
  return OP3_OK;
}

extern "C" int op3_profile_collect(double* ms, double* work, long long* launches) {
  for (int i = 0; i < op3::KID_COUNT; ++i) { ms[i] = 0; work[i] = 0; launches[i] = 0; }
  for (auto& e : op3::g_entries) {
    OP3_CHECK(cudaEventSynchronize(e.b));
    float t = 0.f;
    OP3_CHECK(cudaEventElapsedTime(&t, e.a, e.b));
    ms[e.kid] += t; work[e.kid] += e.work; launches[e.kid] += 1;
    op3::g_pool.push_back(e.a); op3::g_pool.push_back(e.b);
  }
  op3::g_entries.clear();
  return OP3_OK;
}

static int g_precision = OP3_PREC_FP32;
extern "C" int op3_set_precision(int mode) {
  if (mode < 0 || mode > 3) { op3::set_error("op3_set_precision: mode %d not in {0 fp32, 1 tf32x3, 2 tf32, 3 f16x3}", mode); return OP3_ERR_ARG; }
  g_precision = mode;
  return OP3_OK;
}
extern "C" int op3_get_precision(void) { return g_precision; }

static bool g_fused_mlp = true;
namespace op3 { bool fused_mlp_enabled() { return g_fused_mlp; } }
extern "C" int op3_set_fused_mlp(int on) { g_fused_mlp = on != 0; return OP3_OK; }
static bool g_tcq_pair = false;   // opt-in: same results bit for bit, not faster yet (conv5x5_tc.cu, TqRoles)
namespace op3 { bool tcq_pair_enabled() { return g_tcq_pair; } }
extern "C" int op3_set_tcq_pair(int on) { g_tcq_pair = on != 0; return OP3_OK; }

// ---- weight-gradient lane: a second stream per device for the conv weight-gradient kernels of the backward calls ----------
// Nothing downstream of a weight gradient runs before the optimizer, while the dgrad chain is the critical path and is broken
// up by small latency-bound kernels (class reductions, the refinement head, the dynamics chains) that leave most SMs idle.
// With the lane on, every conv wgrad of op3_decoder_bwd / op3_refine_bwd is forked onto the side stream (event edge from the
// launch stream, so it also becomes a parallel branch of a CUDA graph under capture) and fills those SMs; op3_wgrad_join makes
// the launch stream wait for the lane.  Accumulation order per parameter is unchanged (one FIFO lane), so results are bit-identical.
namespace op3 {
namespace {
struct WgradLane { cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, done = nullptr; bool pending = false; };
WgradLane g_lanes[OP3_MAX_DEVICES];
bool g_wgrad_overlap = false;
int lane_get(WgradLane** out) {
  int dev = 0;
  OP3_CHECK(cudaGetDevice(&dev));
  OP3_REQUIRE(dev >= 0 && dev < OP3_MAX_DEVICES, "weight-gradient lane: device %d out of range", dev);
  WgradLane& l = g_lanes[dev];
  if (!l.side) {
    OP3_CHECK(cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking));
    OP3_CHECK(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
    OP3_CHECK(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
  }
  *out = &l;
  return OP3_OK;
}
}  // namespace
int wgrad_stream(cudaStream_t st, cudaStream_t* out) {
  *out = st;
  if (!g_wgrad_overlap) return OP3_OK;
  WgradLane* l = nullptr;
  OP3_TRY(lane_get(&l));
  OP3_CHECK(cudaEventRecord(l->fork, st));
  OP3_CHECK(cudaStreamWaitEvent(l->side, l->fork, 0));
  l->pending = true;
  *out = l->side;
  return OP3_OK;
}
bool wgrad_overlap_enabled() { return g_wgrad_overlap; }
}  // namespace op3
extern "C" int op3_set_wgrad_overlap(int on) {
  if (on) { op3::WgradLane* l = nullptr; OP3_TRY(op3::lane_get(&l)); }   // streams / events are created outside any capture
  op3::g_wgrad_overlap = on != 0;
  return OP3_OK;
}
extern "C" int op3_wgrad_join(void* stream) {
  op3::WgradLane* l = nullptr;
  OP3_TRY(op3::lane_get(&l));
  if (!l->pending) return OP3_OK;
  OP3_CHECK(cudaEventRecord(l->done, l->side));
  OP3_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, l->done, 0));
  l->pending = false;
  return OP3_OK;
}

extern "C" const char* op3_version(void) { return "op3_b200 0.1.0 (sm_100a)"; }
extern "C" const char* op3_last_error(void) { return op3::g_err; }
extern "C" unsigned long long op3_launch_count(void) { return op3::g_launch_count; }
