// pcv_api.cu — the extern "C" surface of libpcv_attn.so (see include/pcv_attn.h).
// Argument validation, kernel-family dispatch and error reporting live here; kernels live in
// pcv_attn_tc.cu (tcgen05), pcv_attn_simt.cu (CUDA cores) and pcv_aux.cu.
#include "pcv_common.cuh"

#include <atomic>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

namespace pcv {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;

void prof_mark_begin(cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on) return;
  cudaEvent_t a, b;
  if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
  cudaEventRecord(a, stream);
  g_prof_events.emplace_back(a, b);
}
void prof_mark_end(cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on || g_prof_events.empty()) return;
  cudaEventRecord(g_prof_events.back().second, stream);
}

static int validate_attn(const pcv_attn_params* p) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "attn: params is NULL");
  PCV_REQUIRE(p->q && p->k && p->v, PCV_ERR_INVALID, "attn: q/k/v pointer is NULL");
  PCV_REQUIRE(p->B >= 1 && p->H >= 1 && p->N >= 1 && p->M >= 1, PCV_ERR_INVALID,
              "attn: B=%d H=%d N=%d M=%d must all be >= 1", p->B, p->H, p->N, p->M);
  PCV_REQUIRE(p->dqk >= 1 && p->dv >= 1, PCV_ERR_INVALID, "attn: dqk=%d dv=%d must be >= 1", p->dqk, p->dv);
  PCV_REQUIRE(p->dtype == PCV_BF16 || p->dtype == PCV_F16, PCV_ERR_INVALID, "attn: unknown dtype %d", p->dtype);
  PCV_REQUIRE(p->m_total >= p->M && p->m_offset >= 0 && p->m_offset + p->M <= p->m_total, PCV_ERR_INVALID,
              "attn: shard [%d,%d) outside m_total=%d", p->m_offset, p->m_offset + p->M, p->m_total);
  PCV_REQUIRE(!p->causal || p->m_total >= p->N, PCV_ERR_INVALID,
              "attn: causal attention needs m_total (%d) >= N (%d)", p->m_total, p->N);
  if (p->write_partial) {
    PCV_REQUIRE(p->part_o && p->part_m && p->part_l, PCV_ERR_INVALID, "attn: write_partial set but part_* NULL");
  } else {
    PCV_REQUIRE(p->out != nullptr, PCV_ERR_INVALID, "attn: out pointer is NULL");
  }
  PCV_REQUIRE(p->impl >= PCV_IMPL_AUTO && p->impl <= PCV_IMPL_DECODE, PCV_ERR_INVALID, "attn: unknown impl %d", p->impl);
  return PCV_OK;
}

// few query rows against a long cache: the streaming kernel (HBM-bound) beats a 128-row tensor-core tile
static bool use_decode(const pcv_attn_params& p, const char** why) {
  if (p.impl != PCV_IMPL_AUTO && p.impl != PCV_IMPL_DECODE) {
    *why = "another kernel was requested";
    return false;
  }
  return attn_decode_supported(p, why);
}

static bool use_tc(const pcv_attn_params& p, const char** why) {
  if (p.impl == PCV_IMPL_SIMT || p.impl == PCV_IMPL_DECODE) {
    *why = "another kernel was requested";
    return false;
  }
  return attn_tc_supported(p, why);
}

}  // namespace pcv

using namespace pcv;

extern "C" {

int pcv_abi_version(void) { return PCV_ABI_VERSION; }

const char* pcv_last_error(void) { return g_err; }

int pcv_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = true;
  return PCV_OK;
}

int pcv_profile_end(double* main_kernel_ms_total, int32_t* main_kernel_launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  double total = 0.0;
  int n = 0;
  for (auto& ev : g_prof_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
      total += ms;
      ++n;
    }
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  g_prof_events.clear();
  if (main_kernel_ms_total) *main_kernel_ms_total = total;
  if (main_kernel_launches) *main_kernel_launches = n;
  return PCV_OK;
}

int pcv_debug_read(uint32_t* out, int32_t n) {
  PCV_REQUIRE(out != nullptr && n >= 0, PCV_ERR_INVALID, "debug_read: bad argument");
  const int rc = debug_read(out, n);
  if (rc == PCV_OK && n > 0 && out[0] == 0u) return bwd_debug_read(out, n);  // nothing from the forward kernels
  return rc;
}

int pcv_debug_trace_read(uint64_t* out, int32_t n) {
  PCV_REQUIRE(out != nullptr, PCV_ERR_INVALID, "trace_read: bad argument");
  return debug_trace_read(reinterpret_cast<unsigned long long*>(out), n);
}

int pcv_debug_plan(int32_t B, int32_t H, int32_t N, int32_t M, int32_t workers, int32_t rows_per_unit,
                   int32_t rows_per_tile, int32_t* segs, int32_t max_segs, int32_t* counts) {
  PCV_REQUIRE(B > 0 && H > 0 && N > 0 && M > 0 && workers > 0, PCV_ERR_INVALID, "debug_plan: sizes must be positive");
  PCV_REQUIRE(rows_per_tile == 128 && (rows_per_unit == 128 || rows_per_unit == 256 || rows_per_unit == 512),
              PCV_ERR_INVALID, "debug_plan: rows_per_tile must be 128 and rows_per_unit 128, 256 or 512");
  PCV_REQUIRE(counts != nullptr && (segs != nullptr || max_segs == 0), PCV_ERR_INVALID, "debug_plan: NULL argument");
  return debug_plan(B, H, N, M, workers, rows_per_unit, rows_per_tile, segs, max_segs, counts);
}

uint64_t pcv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int pcv_get_device_info(pcv_device_info* info) {
  PCV_REQUIRE(info != nullptr, PCV_ERR_INVALID, "device_info: NULL argument");
  int dev = 0;
  PCV_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  PCV_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  info->device = dev;
  info->sm_major = prop.major;
  info->sm_minor = prop.minor;
  info->num_sms = prop.multiProcessorCount;
  info->smem_optin_bytes = (int)prop.sharedMemPerBlockOptin;
  info->tcgen05_ok = (prop.major == 10) ? 1 : 0;
  return PCV_OK;
}

int pcv_attn_supported_tcgen05(const pcv_attn_params* p) {
  if (validate_attn(p) != PCV_OK) return 0;
  const char* why = "";
  const bool ok = attn_tc_supported(*p, &why);
  if (!ok) set_error("tcgen05 path not applicable: %s", why);
  return ok ? 1 : 0;
}

int pcv_attn_workspace_bytes(const pcv_attn_params* p, size_t* bytes) {
  int rc = validate_attn(p);
  if (rc != PCV_OK) return rc;
  PCV_REQUIRE(bytes != nullptr, PCV_ERR_INVALID, "attn: bytes is NULL");
  const char* why = "";
  if (use_decode(*p, &why)) return attn_decode_workspace_bytes(*p, bytes);
  PCV_REQUIRE(p->impl != PCV_IMPL_DECODE, PCV_ERR_UNSUPPORTED, "attn: decode kernel requested but %s", why);
  if (use_tc(*p, &why)) return attn_tc_workspace_bytes(*p, bytes);
  PCV_REQUIRE(p->impl != PCV_IMPL_TCGEN05 && p->impl != PCV_IMPL_TCGEN05_PAIR, PCV_ERR_UNSUPPORTED,
              "attn: tcgen05 kernel requested but %s", why);
  return attn_simt_workspace_bytes(*p, bytes);
}

int pcv_attn_fwd(const pcv_attn_params* p, void* stream) {
  int rc = validate_attn(p);
  if (rc != PCV_OK) return rc;
  const char* why = "";
  if (use_decode(*p, &why)) return launch_attn_decode(*p, reinterpret_cast<cudaStream_t>(stream));
  PCV_REQUIRE(p->impl != PCV_IMPL_DECODE, PCV_ERR_UNSUPPORTED, "attn: decode kernel requested but %s", why);
  if (use_tc(*p, &why)) return launch_attn_tc(*p, reinterpret_cast<cudaStream_t>(stream));
  PCV_REQUIRE(p->impl != PCV_IMPL_TCGEN05 && p->impl != PCV_IMPL_TCGEN05_PAIR, PCV_ERR_UNSUPPORTED,
              "attn: tcgen05 kernel requested but %s", why);
  return launch_attn_simt(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_attn_fwd_sharded_supported(const pcv_attn_params* p) {
  if (validate_attn(p) != PCV_OK) return 0;
  const char* why = "";
  const bool ok = p->impl != PCV_IMPL_SIMT && attn_tc_fuse_supported(*p, &why);
  if (!ok) set_error("fused M-shard merge not applicable: %s", why);
  return ok ? 1 : 0;
}

int pcv_attn_fwd_sharded(const pcv_attn_params* p, const pcv_shard_fuse* f, void* stream) {
  PCV_REQUIRE(p != nullptr && f != nullptr, PCV_ERR_INVALID, "attn_fwd_sharded: NULL argument");
  PCV_REQUIRE(f->num_peers >= 1 && f->num_peers <= PCV_MAX_PEERS && f->rank >= 0 && f->rank < f->num_peers, PCV_ERR_INVALID,
              "attn_fwd_sharded: rank %d of %d", f->rank, f->num_peers);
  PCV_REQUIRE(f->epoch >= 1, PCV_ERR_INVALID, "attn_fwd_sharded: epoch must start at 1");
  for (int g = 0; g < f->num_peers; ++g)
    PCV_REQUIRE(f->part[g] && f->out[g] && f->flags[g], PCV_ERR_INVALID, "attn_fwd_sharded: NULL buffer of rank %d", g);
  pcv_attn_params q = *p;
  q.write_partial = 1;
  // validate_attn wants part_* for a partial launch; they are replaced by part[rank] inside the launcher
  q.part_o = reinterpret_cast<float*>(f->part[f->rank]);
  q.part_m = q.part_o;
  q.part_l = q.part_o;
  int rc = validate_attn(&q);
  if (rc != PCV_OK) return rc;
  return launch_attn_tc(q, reinterpret_cast<cudaStream_t>(stream), f);
}

int pcv_attn_combine(const pcv_combine_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "combine: params is NULL");
  return launch_combine(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_attn_combine_peers(const pcv_peer_combine_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "combine_peers: params is NULL");
  return launch_combine_peers(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_attn_merge_partials(const pcv_merge_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "merge_partials: params is NULL");
  return launch_merge_partials(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_partial_rescale(const pcv_rescale_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "rescale: params is NULL");
  return launch_rescale(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_rotary_apply(const pcv_rotary_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "rotary: params is NULL");
  return launch_rotary(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_kv_append(const pcv_kv_append_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "kv_append: params is NULL");
  return launch_kv_append(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_kv_project_supported(const pcv_kvproj_params* p) {
  if (p == nullptr) return 0;
  const char* why = "";
  const bool ok = kv_project_supported(*p, &why);
  if (!ok) set_error("kv_project not applicable: %s", why);
  return ok ? 1 : 0;
}

int pcv_ln_stats(const pcv_ln_stats_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "ln_stats: params is NULL");
  return launch_ln_stats(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_kv_project(const pcv_kvproj_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "kv_project: params is NULL");
  return launch_kv_project(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_attn_bwd_supported(const pcv_attn_bwd_params* p) {
  if (p == nullptr) return 0;
  const char* why = "";
  const bool ok = attn_bwd_supported(*p, &why);
  if (!ok) set_error("attn_bwd not applicable: %s", why);
  return ok ? 1 : 0;
}

int pcv_attn_bwd_workspace_bytes(const pcv_attn_bwd_params* p, size_t* bytes) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "attn_bwd_workspace_bytes: params is NULL");
  return attn_bwd_workspace_bytes(*p, bytes);
}

int pcv_attn_bwd(const pcv_attn_bwd_params* p, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "attn_bwd: params is NULL");
  return launch_attn_bwd(*p, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_attn_fwd_dropout_supported(const pcv_attn_params* p, float dropout_p) {
  if (p == nullptr) return 0;
  const char* why = "";
  const bool ok = attn_fwd_dropout_supported(*p, dropout_p, &why);
  if (!ok) set_error("attn_fwd_dropout not applicable: %s", why);
  return ok ? 1 : 0;
}

int pcv_attn_fwd_dropout_workspace_bytes(const pcv_attn_params* p, size_t* bytes) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "attn_fwd_dropout_workspace_bytes: params is NULL");
  return attn_fwd_dropout_workspace_bytes(*p, bytes);
}

int pcv_attn_fwd_dropout(const pcv_attn_params* p, const float* stat_m, const float* stat_l, float dropout_p,
                         uint64_t dropout_seed, void* stream) {
  PCV_REQUIRE(p != nullptr, PCV_ERR_INVALID, "attn_fwd_dropout: params is NULL");
  return launch_attn_fwd_dropout(*p, stat_m, stat_l, dropout_p, dropout_seed, reinterpret_cast<cudaStream_t>(stream));
}

int pcv_attn_dropout_mask(uint8_t* keep, int32_t B, int32_t H, int32_t N, int32_t M, float dropout_p,
                          uint64_t dropout_seed, void* stream) {
  return launch_dropout_mask(keep, B, H, N, M, dropout_p, dropout_seed, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
