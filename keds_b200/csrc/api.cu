// C ABI of libkeds_knn.so (see include/keds_knn.h). Host orchestration only: planning, buffers,
// TMA descriptors, launches. All arithmetic is in the kernels of score_topk_sm100.cuh and
// aux_kernels.cuh. There is no CPU path: without a CUDA device every compute call fails.
#include "../../include/keds_knn.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "aux_kernels.cuh"

using namespace keds;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(e_ == cudaErrorMemoryAllocation ? KEDS_ERR_OOM : KEDS_ERR_CUDA,        \
                  "%s failed: %s (%d) at %s:%d", #call, cudaGetErrorString(e_), (int)e_, \
                  __FILE__, __LINE__);                                                   \
  } while (0)

#define CKS(call)          \
  do {                     \
    int s_ = (call);       \
    if (s_ != 0) return s_; \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  // Bumped whenever the buffer moves: a CUDA graph captured over a handle bakes these addresses in,
  // and its owner compares keds_index_generation() before every replay (RetrievalStep.run()).
  uint64_t* gen = nullptr;
  bool borrowed = false;  // p points into memory owned by the caller (bound training parameters): never freed here
  void moved() {
    if (gen) ++*gen;
  }
  // view of caller-owned memory (replaces whatever was held)
  void borrow(void* ptr, size_t bytes) {
    release();
    p = ptr;
    cap = bytes;
    borrowed = true;
  }
  // grow without keeping contents
  int ensure(size_t bytes) {
    if (bytes <= cap && !borrowed) return 0;
    if (borrowed) {
      p = nullptr;
      cap = 0;
      borrowed = false;
    }
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    moved();
    CK(cudaMalloc(&p, bytes));
    cap = bytes;
    return 0;
  }
  // grow keeping the first `keep` bytes
  int grow_keep(size_t bytes, size_t keep) {
    if (bytes <= cap) return 0;
    void* np = nullptr;
    CK(cudaMalloc(&np, bytes));
    if (p && keep) CK(cudaMemcpy(np, p, keep, cudaMemcpyDeviceToDevice));
    if (p) cudaFree(p);
    p = np;
    cap = bytes;
    moved();
    return 0;
  }
  void release() {
    if (p && !borrowed) {
      cudaFree(p);
      moved();
    }
    p = nullptr;
    cap = 0;
    borrowed = false;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

// page-locked host staging (the numpy-in / numpy-out path: one DMA each way instead of the driver's
// pageable-memory staging, and no extra blocking copy for the status words)
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    cap = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 16-bit [rows][d_pad] row-major -> boxes of {BK columns x box_rows rows}, 128-byte swizzle,
// out-of-range rows read as zeros.
int encode_rows_map(CUtensorMap* tm, const void* base, int64_t rows, int d_pad, int box_rows, int fmt) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(KEDS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(d_pad) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, fmt == FMT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(KEDS_ERR_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
  return 0;
}

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// device that owns a device pointer (-1: not device memory)
int device_of(const void* p) {
  if (!p) return -1;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

// [0],[1] n_flagged per db, [2] err word, [3..5] spare, [6] largest candidate band |C| of the search
// (planner feedback), [7] spare
constexpr int CTRL_WORDS = 8;
constexpr int CTRL_BAND = 6;
// One block of status words per pass of a multi-pass call, so that no pass has to be read back
// before the next one starts. Behind the status words: the exact fallback's work counters, one set
// of four words per launch of that kernel (k_prep_rows zeroes the whole block, so the launches of a
// search need no memset between them).
constexpr int EXACT_CTR_SETS = 64;
constexpr int CTRL_STRIDE = CTRL_WORDS + 4 * EXACT_CTR_SETS;
constexpr int S_MAX = 192;
constexpr size_t CAND_BUDGET = size_t(1) << 30;
constexpr int64_t Q_PASS_MAX = 16384;
constexpr size_t TIMING_RING = 8192;
constexpr size_t TIMING_PAIRS = 5;  // k_prep_rows, k_score_topk, k_select_rerank, k_exact_fallback, (reserved)

// keds_index_label_hits: the search chain with the hit-counting kernel in the re-rank's place
struct HitsMode {
  int nks;
  int ks[HITS_MAX_CUTS];
  const long long* row_labels;
  const long long* qlabel;  // already offset to the pass's first query
  int* hits;                // [nq][nks], likewise
};

struct Plan {
  int exact_only = 0;
  bool pair = false;  // CTA-pair scoring kernel (two query tiles per work item)
  int sub = 1;        // candidate lines per (slice, query): 2 in the pair kernel (one per column half)
  int S = 1, n_qt = 1, n_qg = 1, n_items = 0, grid = 0;
  bool small_batch = true;  // one wave of re-rank blocks: the latency variant
};

}  // namespace

struct keds_index {
  int d = 0, d_pad = 0, metric = 0, device = 0, num_sms = 0;
  int64_t n = 0;
  int64_t id_offset = 0;
  float eps_scale = 1.f;
  // 16-bit operand format of x_bf16 / q_bf16 (FMT_FP16 or FMT_BF16; the buffer names predate the
  // choice) and what it was chosen from: running maxima over every row added so far
  int fmt = FMT_FP16;
  int fmt_forced = -1;  // -1: automatic (KEDS_OPERAND / keds_index_set_operand_format override it)
  float amax = 0.f, res_bf16 = 0.f, res_fp16 = 0.f;
  uint64_t generation = 0;        // bumped when a buffer a captured graph may hold is reallocated
  unsigned int* h_feedback = nullptr;  // mapped host word: largest candidate band of a recent search
  DevBuf x_f32, x_bf16, bias, dbstat, probe;
  CUtensorMap tm_x;   // {64 x 256}-row boxes (one CTA per tile)
  CUtensorMap tm_xh;  // {64 x 128}-row boxes (CTA pair: half a tile each)
  bool tm_x_ok = false;
  bool use_pair = true;
  // per-call scratch (one search in flight per handle)
  DevBuf q_f32, q_bf16, qstat, cand, cand_cnt, cand_theta, flagged[2], ctrl, exact_scratch, rk, theta0;
  bool warm_start = true;  // lists start at a finished list's threshold (KEDS_NO_WARM_START=1: every list cold)
  DevBuf D_stage[2], I_stage[2];
  HostBuf h_q, h_out;  // pinned staging for host-pointer calls
  CUtensorMap tm_q;
  const void* tm_q_base = nullptr;
  int64_t tm_q_rows = 0;
  int tm_q_fmt = -1;
  keds_search_stats stats;
  bool attrs_set = false;
  bool use_pdl = true;
  unsigned int* ctrl_cur = nullptr;  // status block of the pass being launched
  int ctrl_passes = 1;             // status blocks the last call used (finish_sync reads them all)
  int rerank_threads_large = 128;  // block size of the throughput re-rank variant (KEDS_RERANK_THREADS)
  int rerank_variant = 0;          // 0: by batch size; 1 / 2: force the latency / throughput variant
  // in-kernel timing of the scoring kernel (bench.py's roofline leg): {min start, max end} ns
  bool timing_on = false;
  DevBuf timing;
  size_t timing_launches = 0;
  // optional per-launch timing of the scoring kernel (bench.py's roofline leg)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_tag;
  size_t prof_used = 0;
};

namespace {

int set_kernel_attrs(keds_index* ix) {
  if (ix->attrs_set) return 0;
  CK(cudaFuncSetAttribute(k_score_topk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SCORE_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_score_topk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SCORE_PAIR_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_score_topk<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SCORE_SMEM_BYTES));

  CK(cudaFuncSetAttribute(k_score_topk<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SCORE_PAIR_SMEM_BYTES));
  const char* no_pair = getenv("KEDS_NO_PAIR");
  ix->use_pair = !(no_pair && no_pair[0] == '1');
  CK(cudaFuncSetAttribute(k_select_rerank<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_select_rerank<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_select_rerank<2, 6, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_exact_fallback, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(k_select_hits<2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const char* no_pdl = getenv("KEDS_NO_PDL");
  ix->use_pdl = !(no_pdl && no_pdl[0] == '1');
  if (const char* ws = getenv("KEDS_NO_WARM_START")) ix->warm_start = !(ws[0] == '1');
  if (const char* rv = getenv("KEDS_RERANK_VARIANT")) {
    if (!strcmp(rv, "latency")) ix->rerank_variant = 1;
    if (!strcmp(rv, "throughput")) ix->rerank_variant = 2;
  }
  if (const char* rt = getenv("KEDS_RERANK_THREADS")) {
    const int v = atoi(rt);
    if (v == 32 || v == 64 || v == 128 || v == 256) ix->rerank_threads_large = v;
  }
  ix->attrs_set = true;
  return 0;
}

// Launch on `st`; with pdl the kernel may be scheduled while its predecessor in the stream is
// still draining (programmatic dependent launch) -- every kernel of the search chain calls
// griddep_wait() before it touches anything an earlier kernel wrote.
// cluster > 1: thread-block cluster of that many CTAs along x (the CTA-pair scoring kernel).
template <typename... KArgs, typename... Args>
int launch_kc(bool pdl, int cluster, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
              cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
  return 0;
}

template <typename... KArgs, typename... Args>
int launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
             Args&&... args) {
  return launch_kc(pdl, 1, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}

// Pick the number of row slices: enough that no slice is expected to hold more than a third of
// LKEEP of the top-k, enough work items to fill the SMs, then whatever minimises the longest CTA.
// band_hint: the largest candidate band |C| a recent search on this handle reported (0: none).
// On i.i.d. data about 2k rows sit at or above tau; clustered embeddings put a whole cluster there,
// and a slice that holds LKEEP of them fails the certificate -- so the expected-fallback term
// below takes whichever is larger and the planner answers crowded bands with more, shorter slices.
Plan make_plan(const keds_index* ix, int n_db, int64_t nq, int k, int64_t n_min, int64_t n_max,
               uint32_t flags, unsigned int band_hint = 0) {
  Plan pl;
  pl.n_qt = static_cast<int>((nq + BM - 1) / BM);
  const int T_min = static_cast<int>((n_min + BN - 1) / BN);
  const int T_max = static_cast<int>((n_max + BN - 1) / BN);
  // a slice hands over its best LKEEP-1 rows and drops the rest (theta = its LKEEP-th best): for
  // the certificate to pass, theta must sit well below the k-th best score overall, i.e. the
  // top-k must be spread over many slices (expected share per slice <= LKEEP / 6)
  // more than one query tile: CTA pairs share each row tile (two query tiles per work item)
  pl.pair = ix->use_pair && pl.n_qt >= 2 && ix->num_sms >= 2;
  pl.sub = pl.pair ? ScoreCfg<true>::kSub : ScoreCfg<false>::kSub;
  // (in candidate lists: the pair kernel keeps one per column half of every slice)
  const int S_sel = std::max(1, ((6 * k + LKEEP - 1) / LKEEP + pl.sub - 1) / pl.sub);
  pl.n_qg = pl.pair ? (pl.n_qt + 1) / 2 : pl.n_qt;
  const int units = pl.pair ? ix->num_sms / 2 : ix->num_sms;
  const int groups = n_db * pl.n_qg;
  int S_hi = std::min(T_min, S_MAX);
  const long long by_mem = static_cast<long long>(CAND_BUDGET / (size_t(LKEEP) * BM * 8)) /
                           (static_cast<long long>(n_db) * pl.n_qt * pl.sub);
  S_hi = static_cast<int>(std::min<long long>(S_hi, by_mem));
  if ((flags & KEDS_SEARCH_EXACT_ONLY) || S_hi < S_sel || k > R_MAX / 2) {
    pl.exact_only = 1;
    return pl;
  }
  // re-rank variant: one wave of blocks (two per SM) runs the latency variant, more the
  // four-per-SM throughput variant
  pl.small_batch = nq * n_db <= 2ll * ix->num_sms;
  if (ix->rerank_variant == 1) pl.small_batch = true;   // experiments: KEDS_RERANK_VARIANT=latency|throughput
  if (ix->rerank_variant == 2) pl.small_batch = false;
  double best = 1e300;
  // one exact scan of every fp32 row, in the cost unit below (one bf16 row tile per work unit)
  const double scan_cost = 2.0 * T_max / units;
  for (int S = S_sel; S <= S_hi; ++S) {
    const long long items = static_cast<long long>(groups) * S;
    const long long G = std::min<long long>(items, units);
    const long long per_cta = (items + G - 1) / G;
    double cost = static_cast<double>(per_cta) * ((T_max + S - 1) / S) + 0.35 * per_cta;
    // expected fallbacks: a query is flagged when one list holds LKEEP or more of the rows at or
    // above tau (Poisson tail, five-fold margin) -- large k wants more lists than the SM count.
    // Rows at or above tau on i.i.d. data: ~2.4k with bf16 operands (wide band), k plus a few with
    // fp16 (measured 18-24 at k = 16, ~215 at k = 200); crowded bands arrive through band_hint.
    {
      const double band = ix->fmt == FMT_FP16 ? 1.15 * k + 12.0 : 2.0 * k;
      const double lam = std::max(band, 1.25 * static_cast<double>(band_hint)) / (S * pl.sub);
      double term = std::exp(-lam), tail = 0.0;  // term_i = e^-lam lam^i / i!
      for (int i = 1; i <= LKEEP + 40; ++i) {
        term *= lam / i;
        if (i >= LKEEP) tail += term;
      }
      cost += 5.0 * tail * static_cast<double>(nq) * n_db * S * pl.sub * scan_cost;
    }
    if (cost < best - 1e-9) {
      best = cost;
      pl.S = S;
    }
  }
  if (const char* dbg = getenv("KEDS_DEBUG_SLICES")) {  // experiments only: force the slice count
    const int v = atoi(dbg);
    if (v >= S_sel && v <= S_hi) pl.S = v;
  }
  pl.n_items = groups * pl.S;
  pl.grid = std::min(pl.n_items, units) * (pl.pair ? 2 : 1);
  return pl;
}

int ensure_q_map(keds_index* ix, int64_t rows_needed, cudaStream_t st) {
  const size_t bytes = static_cast<size_t>(rows_needed) * ix->d_pad * 2;
  if (bytes > ix->q_bf16.cap) {
    // round up so repeated slightly larger batches do not re-encode every call
    const int64_t rows = (rows_needed + 1023) / 1024 * 1024;
    CKS(ix->q_bf16.ensure(static_cast<size_t>(rows) * ix->d_pad * 2));
    // on the caller's stream: a legacy-stream memset is not ordered against a non-blocking stream
    CK(cudaMemsetAsync(ix->q_bf16.p, 0, ix->q_bf16.cap, st));
    ix->tm_q_base = nullptr;
  }
  if (ix->tm_q_base != ix->q_bf16.p || ix->tm_q_fmt != ix->fmt) {
    const int64_t rows = static_cast<int64_t>(ix->q_bf16.cap / (static_cast<size_t>(ix->d_pad) * 2));
    CKS(encode_rows_map(&ix->tm_q, ix->q_bf16.p, rows, ix->d_pad, BM, ix->fmt));
    ix->tm_q_base = ix->q_bf16.p;
    ix->tm_q_rows = rows;
    ix->tm_q_fmt = ix->fmt;
  }
  return 0;
}

// Stage marks (diagnostics; they sit between the kernels, so they suppress the programmatic
// overlap they measure around): tag 0 opens a search, tag t closes stage t
// (1 prep_rows, 2 score_topk, 3 select_rerank, 4 exact fallback pair).
constexpr int PROF_STAGES = 6;
int prof_mark(keds_index* a, cudaStream_t st, int tag) {
  if (!a->profiling) return 0;
  if (a->prof_used == a->prof_ev.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    a->prof_ev.push_back(e);
    a->prof_tag.push_back(0);
  }
  CK(cudaEventRecord(a->prof_ev[a->prof_used], st));
  a->prof_tag[a->prof_used] = tag;
  a->prof_used++;
  return 0;
}

// The exact-fallback kernel, ceil(nq / f_cap) passes; every launch returns at once when its slice
// of the flagged list is empty.
int launch_exact(keds_index* ix, keds_index* dbs[2], int n_db, const float* q_dev, int64_t nq, int k,
                 float* D[2], long long* I[2], int metric, const ConsumeParams& cons,
                 unsigned long long* timing, cudaStream_t st, const PeerOut* peer = nullptr) {
  ExactParams ep;
  memset(&ep, 0, sizeof ep);
  ep.cons = cons;
  ep.n_db = n_db;
  ep.d = ix->d;
  ep.metric = metric;
  ep.k = k;
  ep.q_f32 = q_dev;
  // Queries per launch: as many as the score scratch holds (it is allocated for the batch at hand,
  // never beyond the budget). Up to 1 GB always -- all of a 128-query batch at 2 x 0.5M rows, all of
  // 4,096 queries at 50k rows in ONE launch; large batches get up to 4 GB so that a search is followed
  // by about eight launches of this kernel, not sixty (each of them costs a few microseconds even
  // when nothing is queued: at 65,536 x 0.5M they added up to 2 ms of a 40-ms call).
  int64_t rows_sum = 0;
  for (int i = 0; i < n_db; ++i) rows_sum += dbs[i]->n;
  const long long row_bytes = 4 * std::max<int64_t>(rows_sum, 1);
  long long budget = std::max(1ll << 30, std::min(4ll << 30, static_cast<long long>(nq) * row_bytes / 8));
  long long fc = budget / row_bytes;
  fc = std::max(1ll, std::min<long long>(fc, nq));
  ep.f_cap = static_cast<int>(fc);
  // row chunks (phase-1 work units): about four per block, at least one 32-row group per warp
  const unsigned blocks_ = static_cast<unsigned>(ix->num_sms * 2);
  int64_t chunks_sum = 0;
  int chunks_max = 0;
  for (int i = 0; i < n_db; ++i) {
    const int64_t groups = (dbs[i]->n + 31) / 32;
    const int64_t want = 4ll * blocks_;
    ep.db[i].cg = std::max<int64_t>(EXACT_THREADS / 32, (groups + want - 1) / want);
    ep.db[i].chunks = static_cast<int>((groups + ep.db[i].cg - 1) / ep.db[i].cg);
    chunks_sum += ep.db[i].chunks;
    chunks_max = std::max(chunks_max, ep.db[i].chunks);
  }
  CKS(ix->exact_scratch.ensure(static_cast<size_t>(fc) * (rows_sum + chunks_sum) * 4));
  float* scratch = ix->exact_scratch.as<float>();
  for (int i = 0; i < n_db; ++i) {
    ExactDb& e = ep.db[i];
    e.x_f32 = dbs[i]->x_f32.as<float>();
    e.n_rows = dbs[i]->n;
    e.flagged = ix->flagged[i].as<int>();
    e.n_flagged = reinterpret_cast<int*>(ix->ctrl_cur) + i;
    e.scratch = scratch;
    scratch += static_cast<size_t>(fc) * dbs[i]->n;
    e.cmax = scratch;
    scratch += static_cast<size_t>(fc) * e.chunks;
    e.D = D[i];
    e.I = I[i];
    e.id_offset = dbs[i]->id_offset;
  }
  const int dq = (ix->d + 3) & ~3;
  const size_t smem_sc = static_cast<size_t>(EXACT_QG) * dq * 4 + (EXACT_THREADS / 32) * EXACT_QG * 4;
  ep.ck_cap = std::min(chunks_max, 8192);
  const size_t smem_sel = static_cast<size_t>(cons.part4) * 16 + static_cast<size_t>(k) * 20 + 256 * 4 + 16 + 16 + 32 +
                          static_cast<size_t>(EXACT_LIST_CAP) * 8 + static_cast<size_t>(ep.ck_cap) * 4;
  const size_t smem = std::max(smem_sc, smem_sel);
  if (smem > 160 * 1024) return fail(KEDS_ERR_ARG, "d=%d / k=%d too large for the exact fallback", ix->d, k);
  const unsigned blocks = blocks_;
  ep.err = ix->ctrl_cur + 2;
  ep.band_dev = ix->ctrl_cur + CTRL_BAND;
  const int passes = static_cast<int>((nq + fc - 1) / fc);
  for (int pass = 0; pass < passes; ++pass) {
    ep.pass = pass;
    ep.band_host = pass == passes - 1 ? ix->h_feedback : nullptr;
    if (peer) {
      ep.peer = *peer;
      ep.peer.publish = peer->publish && pass == passes - 1;  // only the step's very last launch publishes
    }
    ep.timing = pass == 0 ? timing : nullptr;
    // work counters of this launch: its own zeroed set (the last set is shared, with a memset, by
    // launches beyond EXACT_CTR_SETS)
    const int set = std::min(pass, EXACT_CTR_SETS - 1);
    ep.work = ix->ctrl_cur + CTRL_WORDS + 4 * set;
    ep.done = ep.work + 2;
    if (pass >= EXACT_CTR_SETS) CK(cudaMemsetAsync(ep.work, 0, 16, st));
    CKS(launch_k(ix->use_pdl, k_exact_fallback, dim3(blocks), dim3(EXACT_THREADS), smem, st, ep));
    ix->stats.launches += 1;
  }
  return 0;
}

// One pass (<= Q_PASS_MAX queries). q_dev: device fp32 [nq][d]. D/I: device outputs.
// cons_in (nullable): neighbour-consumer outputs for this pass, already offset to its first query.
int search_pass(keds_index* ix[2], int n_db, const float* q_dev, int64_t nq, int k, float* D[2],
                long long* I[2], uint32_t flags, cudaStream_t st, float* dump, int64_t ld_dump,
                const ConsumeParams* cons_in, const PeerOut* peer = nullptr, int pass_idx = 0,
                const float* q_src = nullptr, const HitsMode* hm = nullptr) {
  // q_src (nullable): the queries in page-locked host memory, addressed through the mapping;
  // k_prep_rows reads them there once and leaves the fp32 copy the other kernels use in q_dev
  keds_index* a = ix[0];
  const int metric = (flags & KEDS_SEARCH_FORCE_IP) ? METRIC_IP : a->metric;
  CKS(set_kernel_attrs(a));
  ConsumeParams cons;
  memset(&cons, 0, sizeof cons);
  if (cons_in) cons = *cons_in;
  // in-kernel %globaltimer stamps, one {min start, max end} pair per kernel of this chain
  // (prep, score, rerank, exact scores, exact select); the first TIMING_RING searches are timed
  unsigned long long* tchain = nullptr;
  if (a->timing_on && a->timing_launches < TIMING_RING && !dump) {
    tchain = a->timing.as<unsigned long long>() + TIMING_PAIRS * 2 * a->timing_launches;
    a->timing_launches++;
  }
  // (a multi-pass call sizes the status blocks of all its passes before the first launch)
  CKS(a->ctrl.ensure(static_cast<size_t>(pass_idx + 1) * CTRL_STRIDE * 4));
  a->ctrl_cur = a->ctrl.as<unsigned int>() + static_cast<size_t>(pass_idx) * CTRL_STRIDE;
  a->ctrl_passes = pass_idx + 1;
  for (int i = 0; i < n_db; ++i) CKS(a->flagged[i].ensure(static_cast<size_t>(nq) * 4));

  int64_t n_min = ix[0]->n, n_max = ix[0]->n;
  if (n_db > 1) {
    n_min = std::min(n_min, ix[1]->n);
    n_max = std::max(n_max, ix[1]->n);
  }
  const unsigned int band_hint = a->h_feedback ? *static_cast<volatile unsigned int*>(a->h_feedback) : 0u;
  Plan pl = make_plan(a, n_db, nq, k, n_min, n_max, flags, band_hint);
  if (dump && pl.exact_only) return fail(KEDS_ERR_ARG, "debug_scores needs at least one 256-row tile");
  a->stats.exact_only = pl.exact_only;
  a->stats.slices = pl.S;
  a->stats.items = pl.n_items;
  a->stats.grid = pl.grid;

  if (pl.exact_only) {
    if (q_src) CK(cudaMemcpyAsync(const_cast<float*>(q_dev), q_src, static_cast<size_t>(nq) * a->d * 4, cudaMemcpyDefault, st));
    CK(cudaMemsetAsync(a->ctrl_cur, 0, CTRL_STRIDE * 4, st));
    for (int i = 0; i < n_db; ++i) {
      k_flag_all<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, st>>>(
          a->flagged[i].as<int>(), reinterpret_cast<int*>(a->ctrl_cur) + i, static_cast<int>(nq));
      a->stats.launches++;
    }
  } else {
    // operands: bf16 queries + per-query residual norms
    const int64_t q_rows = static_cast<int64_t>(pl.n_qt) * BM;
    CKS(ensure_q_map(a, q_rows, st));
    CKS(a->qstat.ensure(static_cast<size_t>(nq) * sizeof(float4)));
    const int64_t theta_ld = q_rows;
    CKS(a->theta0.ensure(static_cast<size_t>(n_db) * theta_ld * 4));
    {
      CKS(prof_mark(a, st, 0));
      const int threads = 256;
      const long long warps_needed = nq;
      const unsigned blocks =
          static_cast<unsigned>(std::min<long long>((warps_needed * 32 + threads - 1) / threads, 4096));
      CKS(launch_k(a->use_pdl, k_prep_rows, dim3(blocks), dim3(threads), 0, st, q_src ? q_src : q_dev,
                   static_cast<long long>(nq), a->d, a->d_pad, a->fmt, a->q_bf16.as<uint16_t>(),
                   a->qstat.as<float4>(), static_cast<float*>(nullptr), static_cast<unsigned int*>(nullptr),
                   a->ctrl_cur, CTRL_STRIDE, a->theta0.as<float>(), static_cast<int>(n_db * theta_ld), tchain,
                   q_src ? const_cast<float*>(q_dev) : static_cast<float*>(nullptr)));
      a->stats.launches++;
    }
    // candidate lines are indexed by (db, slice, query tile) whatever the work-item grouping
    const size_t items = static_cast<size_t>(n_db) * pl.S * pl.sub * pl.n_qt;
    CKS(a->cand.ensure(items * LKEEP * BM * 8));
    CKS(a->cand_cnt.ensure(items * BM * 4));
    CKS(a->cand_theta.ensure(items * BM * 4));

    ScoreParams sp;
    memset(&sp, 0, sizeof sp);
    sp.n_db = n_db;
    sp.n_qt = pl.n_qt;
    sp.n_qg = pl.n_qg;
    sp.S = pl.S;
    sp.n_items = pl.n_items;
    sp.kblocks = a->d_pad / BK;
    sp.fmt_bits = a->fmt == FMT_BF16 ? kIdescBf16Bits : 0u;
    sp.nq = static_cast<int>(nq);
    for (int i = 0; i < n_db; ++i) {
      sp.n_rows[i] = static_cast<int>(ix[i]->n);
      sp.n_tiles[i] = static_cast<int>((ix[i]->n + BN - 1) / BN);
      sp.bias[i] = metric == METRIC_L2 ? ix[i]->bias.as<float>() : nullptr;
    }
    sp.cand = a->cand.as<uint2>();
    sp.cand_cnt = a->cand_cnt.as<int>();
    sp.cand_theta = a->cand_theta.as<float>();
    sp.theta0 = a->warm_start ? a->theta0.as<float>() : nullptr;
    sp.theta_ld = static_cast<int>(theta_ld);
    sp.err = a->ctrl_cur + 2;
    sp.dump = dump;
    sp.ld_dump = ld_dump;
    sp.timing = tchain ? tchain + 2 : nullptr;
    CKS(prof_mark(a, st, 1));
    if (pl.pair)
      CKS(launch_kc(a->use_pdl, 2, k_score_topk<true>, dim3(pl.grid), dim3(SCORE_PAIR_THREADS), SCORE_PAIR_SMEM_BYTES,
                    st, a->tm_q, ix[0]->tm_xh, n_db > 1 ? ix[1]->tm_xh : ix[0]->tm_xh, sp));
    else
      CKS(launch_k(a->use_pdl, k_score_topk<false>, dim3(pl.grid), dim3(SCORE_THREADS), SCORE_SMEM_BYTES, st,
                   a->tm_q, ix[0]->tm_x, n_db > 1 ? ix[1]->tm_x : ix[0]->tm_x, sp));
    CKS(prof_mark(a, st, 2));
    a->stats.launches++;
    CK(cudaGetLastError());
    if (dump) return 0;

    if (hm) {
      // label hits at the cut points instead of the ranked rows (label_hits.cuh): same lists, same
      // certificate, exact scores only for the rows inside the error band around a cut
      HitsParams hp;
      memset(&hp, 0, sizeof hp);
      hp.n_qt = pl.n_qt;
      hp.S = pl.S * pl.sub;
      hp.nq = static_cast<int>(nq);
      hp.d = a->d;
      hp.metric = metric;
      hp.nks = hm->nks;
      for (int j = 0; j < hm->nks; ++j) hp.ks[j] = hm->ks[j];
      unsigned int want = std::max<unsigned int>(static_cast<unsigned int>(k + k / 2 + 64), band_hint + band_hint / 2);
      int rmax = 256;
      while (rmax < static_cast<int>(std::min<unsigned int>(want, R_MAX))) rmax *= 2;
      hp.rmax = rmax;
      hp.cand = sp.cand;
      hp.cand_cnt = sp.cand_cnt;
      hp.cand_theta = sp.cand_theta;
      hp.q_f32 = q_dev;
      hp.qstat = a->qstat.as<float4>();
      hp.x_f32 = a->x_f32.as<float>();
      hp.dbstat = a->dbstat.as<unsigned int>();
      hp.row_labels = hm->row_labels;
      hp.qlabel = hm->qlabel;
      hp.hits = hm->hits;
      hp.flagged = a->flagged[0].as<int>();
      hp.n_flagged = reinterpret_cast<int*>(a->ctrl_cur);
      hp.eps_scale = a->eps_scale;
      hp.band_max = a->ctrl_cur + CTRL_BAND;
      hp.timing = tchain ? tchain + 4 : nullptr;
      const size_t hslots = static_cast<size_t>(pl.S) * pl.sub * LKEEP;
      // qvec | cs | keys, ids | c_sc, c_info, amb | hist | red | bcast | counters | n_cert, band_end, n_hit
      const size_t hsmem = static_cast<size_t>((a->d + 3) & ~3) * 4 + hslots * 8 + static_cast<size_t>(rmax) * 20 +
                           1024 * 4 + 32 * 4 + 16 + 16 + 3 * HITS_MAX_CUTS * 4;
      if (hsmem > 200 * 1024) return fail(KEDS_ERR_ARG, "label-hits shared memory %zu too large", hsmem);
      // <2 rows in flight per warp, 6 blocks per SM>: measured 0.55 ms at 10,000 x 50k against 0.68 (<3, 4>) / 0.57 (<1, 8>)
      CKS(launch_k(a->use_pdl, k_select_hits<2, 6>, dim3(static_cast<unsigned>(nq)), dim3(HITS_THREADS), hsmem, st, hp));
      CKS(prof_mark(a, st, 3));
      a->stats.launches++;
      CK(cudaGetLastError());
    } else {
    RerankParams rp;
    memset(&rp, 0, sizeof rp);
    rp.n_db = n_db;
    rp.n_qt = pl.n_qt;
    rp.S = pl.S * pl.sub;  // candidate lines per (database, query)
    rp.nq = static_cast<int>(nq);
    rp.k = k;
    rp.d = a->d;
    rp.metric = metric;
    rp.cand = sp.cand;
    rp.cand_cnt = sp.cand_cnt;
    rp.cand_theta = sp.cand_theta;
    rp.q_f32 = q_dev;
    rp.qstat = a->qstat.as<float4>();
    for (int i = 0; i < n_db; ++i) {
      rp.x_f32[i] = ix[i]->x_f32.as<float>();
      rp.dbstat[i] = ix[i]->dbstat.as<unsigned int>();
      rp.D[i] = D[i];
      rp.I[i] = I[i];
      rp.id_offset[i] = ix[i]->id_offset;
      rp.flagged[i] = a->flagged[i].as<int>();
      rp.n_flagged[i] = reinterpret_cast<int*>(a->ctrl_cur) + i;
    }
    rp.eps_scale = a->eps_scale;
    rp.band_max = a->ctrl_cur + CTRL_BAND;
    // one wave of blocks (two per SM): the latency variant; more: the four-per-SM throughput variant
    const bool small_batch = pl.small_batch;
    // Candidate capacity. A single wave has the shared memory to spare: full size. Large batches
    // live on blocks per SM, so they start small and follow the band a recent search reported (a
    // query that does not fit is answered by the exact fallback and raises the next call's figure).
    int rmax = R_MAX;
    if (!small_batch) {
      const unsigned int want =
          std::max<unsigned int>(static_cast<unsigned int>(k + k / 2 + 64), band_hint + band_hint / 2);
      rmax = 256;
      while (rmax < static_cast<int>(std::min<unsigned int>(want, R_MAX))) rmax *= 2;
    }
    rp.rmax = rmax;
    if (peer) rp.peer = *peer;
    rp.cons = cons;
    rp.timing = tchain ? tchain + 4 : nullptr;
    const size_t slots = static_cast<size_t>(pl.S) * pl.sub * LKEEP;
    // qvec | part | okey | keys, ids | smax | a_key, a_id, sel_id, sel_sc | hist | red | bcast | counters | top_*
    const size_t smem = static_cast<size_t>((a->d + 3) & ~3) * 4 + static_cast<size_t>(cons.part4) * 16 +
                        static_cast<size_t>(rmax) * 24 + slots * 8 + pl.S * pl.sub * 4 + 256 * 4 + 32 * 4 + 16 + 16 +
                        static_cast<size_t>(k) * 12;
    if (smem > 200 * 1024) return fail(KEDS_ERR_ARG, "re-rank shared memory %zu too large", smem);
    if (small_batch)
      CKS(launch_k(a->use_pdl, k_select_rerank<3, 2>, dim3(static_cast<unsigned>(nq), n_db),
                   dim3(RERANK_THREADS), smem, st, rp));
    else if (k >= 128 && a->rerank_threads_large == 128)
      // hundreds of fp32 rows per query: two rows per warp in flight (80 registers, six blocks per
      // SM) measured 0.82-0.87 ms against 0.95-1.0 at 10,000 x 50k, k = 200; at k = 16 / 64 the
      // 64-register variant below is 5-10 % faster
      CKS(launch_k(a->use_pdl, k_select_rerank<2, 6, 128>, dim3(static_cast<unsigned>(nq), n_db),
                   dim3(128), smem, st, rp));
    else
      CKS(launch_k(a->use_pdl, k_select_rerank<1, 4>, dim3(static_cast<unsigned>(nq), n_db),
                   dim3(a->rerank_threads_large), smem, st, rp));
    CKS(prof_mark(a, st, 3));
    a->stats.launches++;
    CK(cudaGetLastError());
    }
  }
  if (!(flags & KEDS_SEARCH_NO_FALLBACK) || pl.exact_only) {
    CKS(launch_exact(a, ix, n_db, q_dev, nq, k, D, I, metric, cons, tchain ? tchain + 6 : nullptr, st, peer));
    CKS(prof_mark(a, st, 4));
  }
  if (hm) {
    // the queued queries' exact rows (written by the fallback just now) -> their hits
    HitsParams hp;
    memset(&hp, 0, sizeof hp);
    hp.nks = hm->nks;
    for (int j = 0; j < hm->nks; ++j) hp.ks[j] = hm->ks[j];
    hp.hits = hm->hits;
    CKS(launch_k(a->use_pdl, k_hits_from_rows, dim3(static_cast<unsigned>(std::min<int64_t>((nq + 127) / 128, 64))),
                 dim3(128), 0, st, static_cast<const int*>(a->flagged[0].as<int>()),
                 static_cast<const int*>(reinterpret_cast<int*>(a->ctrl_cur)), static_cast<const long long*>(I[0]), k,
                 static_cast<long long>(a->id_offset), hm->row_labels, hm->qlabel, hp));
    a->stats.launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

__global__ void k_fill_pad(float* D, long long* I, long long n, float dv) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    D[i] = dv;
    I[i] = -1;
  }
}

// device alias of a page-locked host result block (nullptr in: nullptr out)
template <class T>
int mapped_alias(T* host, T** out) {
  *out = nullptr;
  if (!host) return 0;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
    cudaGetLastError();
    return fail(KEDS_ERR_ARG, "retrieve2_hostio: a host result pointer is not page-locked (cudaHostAlloc / pin_memory)");
  }
  *out = static_cast<T*>(at.devicePointer);
  return 0;
}

// Status of a call = its passes' status blocks folded: flagged queries summed, first error word.
void read_status(const keds_index* a, const uint32_t* blocks, uint32_t h[CTRL_WORDS]) {
  for (int i = 0; i < CTRL_WORDS; ++i) h[i] = 0;
  if (!blocks) return;
  for (int p = 0; p < a->ctrl_passes; ++p) {
    const uint32_t* b = blocks + static_cast<size_t>(p) * CTRL_STRIDE;
    h[0] += b[0];
    h[1] += b[1];
    if (h[2] == 0) h[2] = b[2];
  }
}

int finish_sync(keds_index* a, cudaStream_t st) {
  CK(cudaStreamSynchronize(st));
  uint32_t h[CTRL_WORDS] = {0};
  if (a->ctrl.p) {
    std::vector<uint32_t> all(static_cast<size_t>(a->ctrl_passes) * CTRL_STRIDE);
    CK(cudaMemcpy(all.data(), a->ctrl.p, all.size() * 4, cudaMemcpyDeviceToHost));
    read_status(a, all.data(), h);
  }
  a->stats.n_flagged[0] = static_cast<int32_t>(h[0]);
  a->stats.n_flagged[1] = static_cast<int32_t>(h[1]);
  a->stats.err_word = h[2];
  if (h[2] != 0)
    return fail(KEDS_ERR_KERNEL, "device watchdog fired (scoring pipeline or fallback barrier): error word 0x%x", h[2]);
  return 0;
}

int search_impl(keds_index* ix[2], int n_db, const float* q, int64_t nq, int k, float* D[2],
                int64_t* I[2], uint32_t flags, void* stream, const ConsumeParams* cons = nullptr,
                const PeerOut* peer = nullptr, bool q_mapped = false) {
  keds_index* a = ix[0];
  if (!a || !q || nq < 0 || k <= 0) return fail(KEDS_ERR_ARG, "search: null handle/query or bad nq/k");
  if (k > K_MAX) return fail(KEDS_ERR_ARG, "search: k=%d exceeds the maximum %d", k, K_MAX);
  for (int i = 0; i < n_db; ++i) {
    if (!ix[i] || !D[i] || !I[i]) return fail(KEDS_ERR_ARG, "search: null index or output pointer");
    if (ix[i]->d != a->d || ix[i]->metric != a->metric || ix[i]->device != a->device)
      return fail(KEDS_ERR_ARG, "search2: indices differ in d, metric or device");
    if (ix[i]->n > 0x7fffffffll - BN) return fail(KEDS_ERR_ARG, "index too large for 32-bit row ids");
  }
  DeviceGuard g(a->device);
  if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", a->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  memset(&a->stats, 0, sizeof a->stats);
  if (nq == 0) return 0;
  // one query operand feeds both databases: they must share the 16-bit format. bf16 is the one
  // every database can take, so a mixed pair settles there (once; the rows are re-rounded).
  if (n_db > 1 && ix[0]->fmt != ix[1]->fmt)
    for (int i = 0; i < 2; ++i)
      if (ix[i]->fmt != FMT_BF16) CKS(keds_index_set_operand_format(ix[i], FMT_BF16));

  // q_mapped: the queries sit in page-locked host memory and the first kernel of the chain reads
  // them through the mapping (no copy in front of the search); it leaves an fp32 copy in q_f32
  const float* q_map = nullptr;
  if (q_mapped && !is_device_ptr(q)) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, q) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
      cudaGetLastError();
      return fail(KEDS_ERR_ARG, "retrieve2_hostio: q is neither device memory nor page-locked host memory");
    }
    q_map = static_cast<const float*>(at.devicePointer);
    CKS(a->q_f32.ensure(static_cast<size_t>(nq) * a->d * 4));
  }
  const bool q_dev = q_map != nullptr || is_device_ptr(q);
  bool out_dev = true;
  for (int i = 0; i < n_db; ++i) {
    const bool dd = is_device_ptr(D[i]), di = is_device_ptr(I[i]);
    if (dd != di) return fail(KEDS_ERR_ARG, "search: D and I must both be host or both be device memory");
    out_dev = out_dev && dd;
    if (i > 0 && dd != is_device_ptr(D[0]))
      return fail(KEDS_ERR_ARG, "search2: outputs must all be host or all be device memory");
  }
  const float* qd = q_map ? a->q_f32.as<float>() : q;
  bool any_empty = false;
  for (int i = 0; i < n_db; ++i) any_empty = any_empty || ix[i]->n == 0;
  // Host in, host out (the reference's numpy call): the call synchronises before it returns, so
  // page-locked staging buffers can be reused call after call (one real DMA each way instead of
  // the driver's chunked pageable copies). Letting the kernels work on those buffers through the
  // mapping (as keds_retrieve2_hostio does for RetrievalStep) was measured here too and is no
  // faster for this path (two searches of 128 queries: 403-411 us with copies, 413-425 us mapped):
  // the call is dominated by its host side.
  // one database's I (8 B) and D (4 B) blocks in the staged result, labels first (8-byte aligned)
  const size_t res_per = (static_cast<size_t>(nq) * k * 12 + 15) & ~size_t(15);
  if (!q_dev) {
    const size_t qb = static_cast<size_t>(nq) * a->d * 4;
    CKS(a->q_f32.ensure(qb));
    const void* src = q;
    if (!out_dev && qb <= (size_t(64) << 20)) {
      CKS(a->h_q.ensure(qb));
      memcpy(a->h_q.p, q, qb);
      src = a->h_q.p;
    }
    CK(cudaMemcpyAsync(a->q_f32.p, src, qb, cudaMemcpyHostToDevice, st));
    qd = a->q_f32.as<float>();
  }
  float* Dd[2] = {nullptr, nullptr};
  long long* Id[2] = {nullptr, nullptr};
  for (int i = 0; i < n_db; ++i) {
    if (out_dev) {
      Dd[i] = D[i];
      Id[i] = reinterpret_cast<long long*>(I[i]);
    } else {
      CKS(a->D_stage[i].ensure(static_cast<size_t>(nq) * k * 4));
      CKS(a->I_stage[i].ensure(static_cast<size_t>(nq) * k * 8));
      Dd[i] = a->D_stage[i].as<float>();
      Id[i] = a->I_stage[i].as<long long>();
    }
  }
  // status blocks of every pass, sized before the first launch (growing them later would free
  // memory that kernels of an earlier pass still write)
  const int n_passes = static_cast<int>((nq + Q_PASS_MAX - 1) / Q_PASS_MAX);
  CKS(a->ctrl.ensure(static_cast<size_t>(n_passes) * CTRL_STRIDE * 4));
  // empty databases answer with padding only
  if (any_empty) {
    if (cons) return fail(KEDS_ERR_ARG, "retrieve2: both databases must hold rows");
    // no kernel of a chain may run on this handle: its status blocks read as clean
    CK(cudaMemsetAsync(a->ctrl.p, 0, static_cast<size_t>(n_passes) * CTRL_STRIDE * 4, st));
    a->ctrl_passes = n_passes;
    for (int i = 0; i < n_db; ++i) {
      if (ix[i]->n != 0) {
        keds_index* one[2] = {ix[i], nullptr};
        float* D1[2] = {Dd[i], nullptr};
        long long* I1[2] = {Id[i], nullptr};
        for (int64_t q0 = 0; q0 < nq; q0 += Q_PASS_MAX) {
          const int64_t nb = std::min<int64_t>(Q_PASS_MAX, nq - q0);
          float* Dp[2] = {D1[0] + q0 * k, nullptr};
          long long* Ip[2] = {I1[0] + q0 * k, nullptr};
          // (stream order keeps the passes apart: every kernel of a chain waits for its predecessor)
          CKS(one[0]->ctrl.ensure(static_cast<size_t>(n_passes) * CTRL_STRIDE * 4));
          CKS(search_pass(one, 1, qd + q0 * a->d, nb, k, Dp, Ip, flags, st, nullptr, 0, nullptr, nullptr,
                          static_cast<int>(q0 / Q_PASS_MAX)));
        }
      } else {
        const long long tot = static_cast<long long>(nq) * k;
        k_fill_pad<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, st>>>(
            Dd[i], Id[i], tot,
            (a->metric == METRIC_L2 && !(flags & KEDS_SEARCH_FORCE_IP)) ? FLT_MAX : -FLT_MAX);
      }
    }
  } else {
    for (int64_t q0 = 0; q0 < nq; q0 += Q_PASS_MAX) {
      const int64_t nb = std::min<int64_t>(Q_PASS_MAX, nq - q0);
      float* Dp[2] = {Dd[0] + q0 * k, n_db > 1 ? Dd[1] + q0 * k : nullptr};
      long long* Ip[2] = {Id[0] + q0 * k, n_db > 1 ? Id[1] + q0 * k : nullptr};
      ConsumeParams cp;
      if (cons) {
        cp = *cons;
        for (int s = 0; s < 2; ++s) {
          if (cp.feat[s]) cp.feat[s] += q0 * k * a->d;
          if (cp.pool[s]) cp.pool[s] += q0 * a->d;
          if (cp.host_D[s]) cp.host_D[s] += q0 * k;
          if (cp.host_I[s]) cp.host_I[s] += q0 * k;
        }
      }
      PeerOut pp;
      if (peer) {
        pp = *peer;
        for (int r = 0; r < pp.n; ++r) {
          if (pp.D[r]) pp.D[r] += q0 * k;
          if (pp.I[r]) pp.I[r] += q0 * k;
        }
        pp.publish = q0 + nb >= nq;  // the flags go out behind the step's last pass
      }
      // every pass has its own status block and the scratch is reused in stream order (each kernel
      // of a chain waits for its predecessor to complete), so the passes queue without a host round trip
      CKS(search_pass(ix, n_db, qd + q0 * a->d, nb, k, Dp, Ip, flags, st, nullptr, 0, cons ? &cp : nullptr,
                      peer ? &pp : nullptr, static_cast<int>(q0 / Q_PASS_MAX), q_map ? q_map + q0 * a->d : nullptr));
    }
  }
  if (!out_dev) {
    // results and status words into pinned memory with asynchronous copies, ONE synchronisation,
    // then plain memcpy into the caller's arrays
    const size_t db_ = static_cast<size_t>(nq) * k * 4, ib_ = static_cast<size_t>(nq) * k * 8;
    const size_t per = res_per;
    if (per * n_db > (size_t(256) << 20)) {
      // very large result blocks: not worth pinning that much host memory, copy straight out
      for (int i = 0; i < n_db; ++i) {
        CK(cudaMemcpyAsync(D[i], Dd[i], db_, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(I[i], Id[i], ib_, cudaMemcpyDeviceToHost, st));
      }
      return finish_sync(a, st);
    }
    const size_t ctrl_bytes = static_cast<size_t>(a->ctrl_passes) * CTRL_STRIDE * 4;
    CKS(a->h_out.ensure(per * n_db + ctrl_bytes));
    uint8_t* h = static_cast<uint8_t*>(a->h_out.p);
    for (int i = 0; i < n_db; ++i) {
      CK(cudaMemcpyAsync(h + per * i, Id[i], ib_, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(h + per * i + ib_, Dd[i], db_, cudaMemcpyDeviceToHost, st));
    }
    uint32_t* hall = reinterpret_cast<uint32_t*>(h + per * n_db);
    memset(hall, 0, ctrl_bytes);
    if (a->ctrl.p) CK(cudaMemcpyAsync(hall, a->ctrl.p, ctrl_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint32_t hc[CTRL_WORDS];
    read_status(a, hall, hc);
    for (int i = 0; i < n_db; ++i) {
      memcpy(I[i], h + per * i, ib_);
      memcpy(D[i], h + per * i + ib_, db_);
    }
    a->stats.n_flagged[0] = static_cast<int32_t>(hc[0]);
    a->stats.n_flagged[1] = static_cast<int32_t>(hc[1]);
    a->stats.err_word = hc[2];
    if (hc[2] != 0) return fail(KEDS_ERR_KERNEL, "device watchdog fired (scoring pipeline or fallback barrier): error word 0x%x", hc[2]);
    return 0;
  }
  return 0;
}

}  // namespace

extern "C" {

const char* keds_last_error(void) { return g_err.c_str(); }
const char* keds_version(void) { return "keds-knn-b200 0.1 (sm_100a)"; }

int keds_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int keds_index_create(int d, int metric, int device, keds_index_t** out) {
  if (!out) return fail(KEDS_ERR_ARG, "create: out is null");
  *out = nullptr;
  if (d <= 0 || d > 16384) return fail(KEDS_ERR_ARG, "create: bad dimension %d", d);
  if (metric != KEDS_METRIC_IP && metric != KEDS_METRIC_L2)
    return fail(KEDS_ERR_ARG, "create: metric must be 0 (IP) or 1 (L2)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KEDS_ERR_NO_GPU, "no CUDA device: libkeds_knn has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(KEDS_ERR_ARG, "create: device %d of %d", device, ndev);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(KEDS_ERR_NO_GPU, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  DeviceGuard g(device);
  if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", device);
  keds_index* ix = new keds_index();
  ix->d = d;
  ix->d_pad = (d + BK - 1) / BK * BK;
  ix->metric = metric;
  ix->device = device;
  ix->num_sms = prop.multiProcessorCount;
  memset(&ix->stats, 0, sizeof ix->stats);
  if (const char* op = getenv("KEDS_OPERAND")) {  // "bf16" / "fp16" pin the operand format; default: automatic
    if (!strcmp(op, "bf16")) ix->fmt_forced = FMT_BF16;
    if (!strcmp(op, "fp16")) ix->fmt_forced = FMT_FP16;
  }
  int s = ix->dbstat.ensure(16);
  if (s == 0 && cudaMemset(ix->dbstat.p, 0, 16) != cudaSuccess) s = fail(KEDS_ERR_CUDA, "memset failed");
  if (s == 0) s = ix->probe.ensure(16);
  if (s == 0) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
      s = fail(KEDS_ERR_CUDA, "cudaHostAlloc of the feedback word failed");
    else {
      memset(h, 0, 64);
      ix->h_feedback = static_cast<unsigned int*>(h);
    }
  }
  if (s != 0) {
    keds_index_free(ix);
    return s;
  }
  // every buffer whose address a captured search bakes in reports its moves
  DevBuf* tracked[] = {&ix->x_f32, &ix->x_bf16, &ix->bias, &ix->q_f32, &ix->q_bf16, &ix->qstat, &ix->cand,
                       &ix->cand_cnt, &ix->cand_theta, &ix->flagged[0], &ix->flagged[1], &ix->ctrl,
                       &ix->exact_scratch, &ix->D_stage[0], &ix->D_stage[1], &ix->I_stage[0], &ix->I_stage[1],
                       &ix->rk, &ix->theta0};
  for (DevBuf* b : tracked) b->gen = &ix->generation;
  *out = ix;
  return 0;
}

void keds_index_free(keds_index_t* ix) {
  if (!ix) return;
  DeviceGuard g(ix->device);
  DevBuf* bufs[] = {&ix->x_f32, &ix->x_bf16, &ix->bias, &ix->dbstat, &ix->q_f32, &ix->q_bf16,
                    &ix->qstat, &ix->cand, &ix->cand_cnt, &ix->cand_theta, &ix->flagged[0],
                    &ix->flagged[1], &ix->ctrl, &ix->exact_scratch, &ix->D_stage[0], &ix->D_stage[1],
                    &ix->I_stage[0], &ix->I_stage[1], &ix->timing, &ix->probe, &ix->rk, &ix->theta0};
  for (DevBuf* b : bufs) b->release();
  ix->h_q.release();
  ix->h_out.release();
  if (ix->h_feedback) cudaFreeHost(ix->h_feedback);
  for (cudaEvent_t e : ix->prof_ev) cudaEventDestroy(e);
  delete ix;
}

int keds_index_add(keds_index_t* ix, const float* x, int64_t n) { return keds_index_add_ex(ix, x, n, 0u); }

int keds_index_get_rows(const keds_index_t* ix, int64_t first, int64_t n, float* out) {
  if (!ix || !out || first < 0 || n < 0 || first + n > ix->n)
    return fail(KEDS_ERR_ARG, "get_rows: bad range [%lld, %lld) of %lld rows", (long long)first,
                (long long)(first + n), (long long)(ix ? ix->n : 0));
  if (n == 0) return 0;
  DeviceGuard g(ix->device);
  CK(cudaMemcpy(out, ix->x_f32.as<float>() + first * ix->d, static_cast<size_t>(n) * ix->d * 4,
                cudaMemcpyDefault));
  return 0;
}

// Operand format from what the rows look like (running maxima over every add): fp16 when nothing
// comes near its range limit and it keeps more of the rows than bf16 does (always the case for
// unit-norm embeddings: 11 significant bits against 8), bf16 otherwise.
static int choose_format(const keds_index* ix) {
  if (ix->fmt_forced >= 0) return ix->fmt_forced;
  return (ix->amax < 3.0e4f && ix->res_fp16 <= ix->res_bf16) ? FMT_FP16 : FMT_BF16;
}

// (re)build the 16-bit operand copy, the L2 bias and the certificate statistics of rows [r0, r1)
static int prep_db_rows(keds_index* ix, int64_t r0, int64_t r1) {
  const size_t d = ix->d, dp = ix->d_pad;
  const int64_t n = r1 - r0;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((n * 32 + 255) / 256, ix->num_sms * 16));
  k_prep_rows<<<blocks, 256>>>(ix->x_f32.as<float>() + r0 * d, n, ix->d, ix->d_pad, ix->fmt,
                               ix->x_bf16.as<uint16_t>() + r0 * dp, nullptr, ix->bias.as<float>() + r0,
                               ix->dbstat.as<unsigned int>(), nullptr, 0, nullptr, 0, nullptr, nullptr);
  CK(cudaGetLastError());
  return 0;
}

int keds_index_add_ex(keds_index_t* ix, const float* x, int64_t n, uint32_t flags) {
  if (!ix || (n > 0 && !x) || n < 0) return fail(KEDS_ERR_ARG, "add: null handle/data or negative n");
  if (n == 0) return 0;
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", ix->device);
  const int64_t n0 = ix->n, n1 = n0 + n;
  if (n1 > 0x7fffffffll - BN) return fail(KEDS_ERR_ARG, "add: more than 2^31 rows");
  const size_t d = ix->d, dp = ix->d_pad;
  const int64_t cap_rows = n0 == 0 ? n1 : std::max<int64_t>(n1, n0 + n0 / 2);
  const int64_t tiles_cap = (cap_rows + BN - 1) / BN;
  if (static_cast<size_t>(n1) * d * 4 > ix->x_f32.cap) {
    CKS(ix->x_f32.grow_keep(static_cast<size_t>(cap_rows) * d * 4, static_cast<size_t>(n0) * d * 4));
    CKS(ix->x_bf16.grow_keep(static_cast<size_t>(cap_rows) * dp * 2, static_cast<size_t>(n0) * dp * 2));
    CKS(ix->bias.grow_keep(static_cast<size_t>(tiles_cap) * BN * 4, static_cast<size_t>(n0) * 4));
  }
  CK(cudaMemcpy(ix->x_f32.as<float>() + n0 * d, x, static_cast<size_t>(n) * d * 4, cudaMemcpyDefault));
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((n * 32 + 255) / 256, ix->num_sms * 16));
  if (flags & KEDS_ADD_NORMALIZE) k_normalize_rows<<<blocks, 256>>>(ix->x_f32.as<float>() + n0 * d, n, ix->d);
  // which 16-bit format keeps more of these rows (one extra read of the upload)
  {
    CK(cudaMemset(ix->probe.p, 0, 16));
    k_probe_rows<<<blocks, 256>>>(ix->x_f32.as<float>() + n0 * d, n, ix->d, ix->probe.as<unsigned int>());
    CK(cudaGetLastError());
    float h[4] = {0.f, 0.f, 0.f, 0.f};
    CK(cudaMemcpy(h, ix->probe.p, 12, cudaMemcpyDeviceToHost));
    ix->amax = std::max(ix->amax, h[0]);
    ix->res_bf16 = std::max(ix->res_bf16, h[1]);
    ix->res_fp16 = std::max(ix->res_fp16, h[2]);
  }
  const int want = choose_format(ix);
  if (n0 > 0 && want != ix->fmt) {
    // rows added earlier were rounded to the other format: redo them from the fp32 master
    ix->fmt = want;
    CK(cudaMemset(ix->dbstat.p, 0, 16));
    CKS(prep_db_rows(ix, 0, n1));
  } else {
    ix->fmt = want;
    CKS(prep_db_rows(ix, n0, n1));
  }
  const int64_t padded = (n1 + BN - 1) / BN * BN;
  if (padded > n1)
    k_fill_f32<<<static_cast<unsigned>((padded - n1 + 255) / 256), 256>>>(ix->bias.as<float>() + n1,
                                                                          padded - n1, -INFINITY);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  ix->n = n1;
  CKS(encode_rows_map(&ix->tm_x, ix->x_bf16.p, n1, ix->d_pad, BN, ix->fmt));
  CKS(encode_rows_map(&ix->tm_xh, ix->x_bf16.p, n1, ix->d_pad, BN / 2, ix->fmt));
  ix->tm_x_ok = true;
  ix->generation++;  // the row count and the tensor maps are part of what a captured search holds
  return 0;
}

int keds_index_operand_format(const keds_index_t* ix) { return ix ? ix->fmt : -1; }

int keds_index_set_operand_format(keds_index_t* ix, int fmt) {
  if (!ix || fmt < -1 || fmt > FMT_FP16) return fail(KEDS_ERR_ARG, "set_operand_format: fmt must be -1 (auto), 0 (bf16) or 1 (fp16)");
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", ix->device);
  ix->fmt_forced = fmt;
  const int want = choose_format(ix);
  if (want == ix->fmt) return 0;
  ix->fmt = want;
  ix->generation++;
  if (ix->n == 0) return 0;
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(ix->dbstat.p, 0, 16));
  CKS(prep_db_rows(ix, 0, ix->n));
  CK(cudaDeviceSynchronize());
  CKS(encode_rows_map(&ix->tm_x, ix->x_bf16.p, ix->n, ix->d_pad, BN, ix->fmt));
  CKS(encode_rows_map(&ix->tm_xh, ix->x_bf16.p, ix->n, ix->d_pad, BN / 2, ix->fmt));
  return 0;
}

uint64_t keds_index_generation(const keds_index_t* ix) { return ix ? ix->generation : 0; }

int keds_index_reset(keds_index_t* ix) {
  if (!ix) return fail(KEDS_ERR_ARG, "reset: null handle");
  DeviceGuard g(ix->device);
  CK(cudaDeviceSynchronize());
  ix->x_f32.release();
  ix->x_bf16.release();
  ix->bias.release();
  ix->n = 0;
  ix->tm_x_ok = false;
  ix->amax = ix->res_bf16 = ix->res_fp16 = 0.f;
  ix->generation++;
  CK(cudaMemset(ix->dbstat.p, 0, 16));
  return 0;
}

int64_t keds_index_ntotal(const keds_index_t* ix) { return ix ? ix->n : -1; }
int keds_index_dim(const keds_index_t* ix) { return ix ? ix->d : -1; }
int keds_index_metric(const keds_index_t* ix) { return ix ? ix->metric : -1; }
int keds_index_device(const keds_index_t* ix) { return ix ? ix->device : -1; }
const float* keds_index_rows(const keds_index_t* ix) { return ix ? ix->x_f32.as<float>() : nullptr; }

int keds_index_set_id_offset(keds_index_t* ix, int64_t offset) {
  if (!ix) return fail(KEDS_ERR_ARG, "set_id_offset: null handle");
  ix->id_offset = offset;
  return 0;
}

int keds_index_set_pdl(keds_index_t* ix, int enable) {
  if (!ix) return fail(KEDS_ERR_ARG, "set_pdl: null handle");
  DeviceGuard g(ix->device);
  CKS(set_kernel_attrs(ix));
  ix->use_pdl = enable != 0;
  return 0;
}

int keds_index_set_eps_scale(keds_index_t* ix, float scale) {
  if (!ix || !(scale >= 0.f)) return fail(KEDS_ERR_ARG, "set_eps_scale: bad argument");
  ix->eps_scale = scale;
  return 0;
}

int keds_index_search(keds_index_t* ix, const float* q, int64_t nq, int k, float* D, int64_t* I,
                      void* stream) {
  return keds_index_search_ex(ix, q, nq, k, D, I, 0u, stream);
}

int keds_index_search_ex(keds_index_t* ix, const float* q, int64_t nq, int k, float* D, int64_t* I,
                         uint32_t flags, void* stream) {
  keds_index* v[2] = {ix, nullptr};
  float* Dv[2] = {D, nullptr};
  int64_t* Iv[2] = {I, nullptr};
  return search_impl(v, 1, q, nq, k, Dv, Iv, flags, stream);
}

int keds_index_search2(keds_index_t* a, keds_index_t* b, const float* q, int64_t nq, int k, float* Da,
                       int64_t* Ia, float* Db, int64_t* Ib, uint32_t flags, void* stream) {
  keds_index* v[2] = {a, b};
  float* Dv[2] = {Da, Db};
  int64_t* Iv[2] = {Ia, Ib};
  return search_impl(v, 2, q, nq, k, Dv, Iv, flags, stream);
}

}  // extern "C"

namespace {

int retrieve2_impl(keds_index_t* img, keds_index_t* txt, const float* q, int64_t nq, int k,
                   const int32_t* perm_img, const int32_t* perm_txt, int pool_mode, float tau,
                   float* D_img, int64_t* I_img, float* D_txt, int64_t* I_txt, float* feat_img,
                   float* feat_txt, float* pool_img, float* pool_txt, uint32_t flags, void* stream,
                   bool hostio, float* Dh_img, int64_t* Ih_img, float* Dh_txt, int64_t* Ih_txt) {
  if (!img || !txt || !q || !D_img || !I_img || !D_txt || !I_txt)
    return fail(KEDS_ERR_ARG, "retrieve2: null handle, query or result pointer");
  if (pool_mode < 0 || pool_mode > 2) return fail(KEDS_ERR_ARG, "retrieve2: pool_mode must be 0, 1 or 2");
  if ((!hostio && !is_device_ptr(q)) || !is_device_ptr(D_img) || !is_device_ptr(I_img) || !is_device_ptr(D_txt) ||
      !is_device_ptr(I_txt))
    return fail(KEDS_ERR_ARG, "retrieve2: device pointers only");
  if ((Dh_img == nullptr) != (Ih_img == nullptr) || (Dh_txt == nullptr) != (Ih_txt == nullptr))
    return fail(KEDS_ERR_ARG, "retrieve2_hostio: a host mirror needs both its D and its I block");
  if (k > 1024) return fail(KEDS_ERR_ARG, "retrieve2: k=%d too large for the fused consumer", k);
  keds_index* v[2] = {img, txt};
  float* Dv[2] = {D_img, D_txt};
  int64_t* Iv[2] = {I_img, I_txt};
  const bool want_pool = pool_mode != 0 && (pool_img || pool_txt);
  ConsumeParams cp;
  memset(&cp, 0, sizeof cp);
  {
    DeviceGuard g(img->device);
    if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", img->device);
    long long* ih[2] = {nullptr, nullptr};
    CKS(mapped_alias(Dh_img, &cp.host_D[0]));
    CKS(mapped_alias(reinterpret_cast<long long*>(Ih_img), &ih[0]));
    CKS(mapped_alias(Dh_txt, &cp.host_D[1]));
    CKS(mapped_alias(reinterpret_cast<long long*>(Ih_txt), &ih[1]));
    cp.host_I[0] = ih[0];
    cp.host_I[1] = ih[1];
  }
  cp.enabled = (feat_img || feat_txt || want_pool || Dh_img || Dh_txt) ? 1 : 0;
  cp.mode = want_pool ? pool_mode : 0;
  cp.tau = tau;
  cp.perm[0] = perm_img;
  cp.perm[1] = perm_txt;
  cp.feat[0] = feat_img;
  cp.feat[1] = feat_txt;
  cp.pool[0] = want_pool ? pool_img : nullptr;
  cp.pool[1] = want_pool ? pool_txt : nullptr;
  // one warp per neighbour row; per-warp partial pools meet in shared memory ([warps][d/4] float4)
  static_assert(RERANK_THREADS == EXACT_THREADS, "the consumer scratch is sized for one block shape");
  cp.part4 = (img->d & 3) == 0 ? (RERANK_THREADS / 32) * (img->d >> 2) : 0;
  return search_impl(v, 2, q, nq, k, Dv, Iv, flags, stream, cp.enabled ? &cp : nullptr, nullptr, hostio);
}

}  // namespace

extern "C" {

int keds_retrieve2(keds_index_t* img, keds_index_t* txt, const float* q, int64_t nq, int k,
                   const int32_t* perm_img, const int32_t* perm_txt, int pool_mode, float tau,
                   float* D_img, int64_t* I_img, float* D_txt, int64_t* I_txt, float* feat_img,
                   float* feat_txt, float* pool_img, float* pool_txt, uint32_t flags, void* stream) {
  return retrieve2_impl(img, txt, q, nq, k, perm_img, perm_txt, pool_mode, tau, D_img, I_img, D_txt, I_txt, feat_img,
                        feat_txt, pool_img, pool_txt, flags, stream, false, nullptr, nullptr, nullptr, nullptr);
}

int keds_retrieve2_hostio(keds_index_t* img, keds_index_t* txt, const float* q, int64_t nq, int k,
                          const int32_t* perm_img, const int32_t* perm_txt, int pool_mode, float tau,
                          float* D_img, int64_t* I_img, float* D_txt, int64_t* I_txt, float* Dh_img,
                          int64_t* Ih_img, float* Dh_txt, int64_t* Ih_txt, float* feat_img, float* feat_txt,
                          float* pool_img, float* pool_txt, uint32_t flags, void* stream) {
  return retrieve2_impl(img, txt, q, nq, k, perm_img, perm_txt, pool_mode, tau, D_img, I_img, D_txt, I_txt, feat_img,
                        feat_txt, pool_img, pool_txt, flags, stream, true, Dh_img, Ih_img, Dh_txt, Ih_txt);
}

int keds_index_sync(keds_index_t* ix, void* stream) {
  if (!ix) return fail(KEDS_ERR_ARG, "sync: null handle");
  DeviceGuard g(ix->device);
  return finish_sync(ix, static_cast<cudaStream_t>(stream));
}

int keds_index_set_profiling(keds_index_t* ix, int enable) {
  if (!ix) return fail(KEDS_ERR_ARG, "set_profiling: null handle");
  DeviceGuard g(ix->device);
  ix->profiling = enable == 2;  // stage marks (stream events between the kernels)
  ix->prof_used = 0;
  ix->timing_on = enable == 1;  // in-kernel timer of the scoring kernel, launch chain untouched
  ix->timing_launches = 0;
  if (ix->timing_on) {
    const size_t words = TIMING_RING * TIMING_PAIRS * 2;
    CKS(ix->timing.ensure(words * 8));
    std::vector<unsigned long long> init(words);
    for (size_t i = 0; i < words; i += 2) {
      init[i] = ~0ull;
      init[i + 1] = 0ull;
    }
    CK(cudaMemcpy(ix->timing.p, init.data(), words * 8, cudaMemcpyHostToDevice));
  }
  return 0;
}

int keds_index_profile_chain(keds_index_t* ix, double* dur_ms, double* gap_ms, int64_t* searches, int n) {
  if (!ix || !dur_ms || !gap_ms || !searches || n < static_cast<int>(TIMING_PAIRS))
    return fail(KEDS_ERR_ARG, "profile_chain: need room for %d kernels", (int)TIMING_PAIRS);
  if (!ix->timing_on) return fail(KEDS_ERR_ARG, "profile_chain: call set_profiling(1) first");
  DeviceGuard g(ix->device);
  CK(cudaDeviceSynchronize());
  const size_t words = TIMING_RING * TIMING_PAIRS * 2;
  std::vector<unsigned long long> t(words);
  CK(cudaMemcpy(t.data(), ix->timing.p, words * 8, cudaMemcpyDeviceToHost));
  const size_t used = std::min(ix->timing_launches, TIMING_RING);
  double dur[TIMING_PAIRS] = {0}, gap[TIMING_PAIRS] = {0};
  int64_t nd[TIMING_PAIRS] = {0}, ng[TIMING_PAIRS] = {0};
  for (size_t s = 0; s < used; ++s) {
    const unsigned long long* c = t.data() + s * TIMING_PAIRS * 2;
    unsigned long long prev_end = 0;
    for (size_t i = 0; i < TIMING_PAIRS; ++i) {
      const unsigned long long b = c[2 * i], e = c[2 * i + 1];
      if (e <= b || b == ~0ull) continue;  // kernel not part of this search
      dur[i] += static_cast<double>(e - b) * 1e-6;
      nd[i]++;
      if (prev_end != 0 && b >= prev_end) {
        gap[i] += static_cast<double>(b - prev_end) * 1e-6;
        ng[i]++;
      } else if (prev_end != 0) {
        ng[i]++;  // overlapped its predecessor (programmatic launch): gap 0
      }
      prev_end = e;
    }
  }
  for (size_t i = 0; i < TIMING_PAIRS; ++i) {
    dur_ms[i] = nd[i] ? dur[i] / nd[i] : 0.0;
    gap_ms[i] = ng[i] ? gap[i] / ng[i] : 0.0;
  }
  *searches = static_cast<int64_t>(used);
  return 0;
}

int keds_index_profile(keds_index_t* ix, double* score_ms_total, int64_t* score_launches) {
  if (!ix || !score_ms_total || !score_launches) return fail(KEDS_ERR_ARG, "profile: null argument");
  if (ix->timing_on) {
    double dur[TIMING_PAIRS], gap[TIMING_PAIRS];
    int64_t n = 0;
    CKS(keds_index_profile_chain(ix, dur, gap, &n, TIMING_PAIRS));
    *score_ms_total = dur[1] * static_cast<double>(n);
    *score_launches = n;
    return keds_index_set_profiling(ix, 1);  // re-arm the ring
  }
  double ms[PROF_STAGES];
  int64_t cnt[PROF_STAGES];
  CKS(keds_index_profile_stages(ix, ms, cnt, PROF_STAGES));
  *score_ms_total = ms[2];
  *score_launches = cnt[2];
  return 0;
}

int keds_index_profile_stages(keds_index_t* ix, double* ms_total, int64_t* launches, int n_stages) {
  if (!ix || !ms_total || !launches || n_stages < PROF_STAGES)
    return fail(KEDS_ERR_ARG, "profile_stages: need room for %d stages", PROF_STAGES);
  DeviceGuard g(ix->device);
  for (int t = 0; t < n_stages; ++t) {
    ms_total[t] = 0.0;
    launches[t] = 0;
  }
  for (size_t i = 1; i < ix->prof_used; ++i) {
    const int tag = ix->prof_tag[i];
    if (tag <= 0 || tag >= PROF_STAGES) continue;  // tag 0 opens a new search
    CK(cudaEventSynchronize(ix->prof_ev[i]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ix->prof_ev[i - 1], ix->prof_ev[i]));
    ms_total[tag] += ms;
    launches[tag]++;
  }
  ix->prof_used = 0;
  return 0;
}

int keds_index_last_stats(const keds_index_t* ix, keds_search_stats* out) {
  if (!ix || !out) return fail(KEDS_ERR_ARG, "last_stats: null argument");
  *out = ix->stats;
  return 0;
}

int keds_debug_scores(keds_index_t* ix, const float* q, int64_t nq, float* out, void* stream) {
  if (!ix || !q || !out || nq <= 0 || nq > Q_PASS_MAX) return fail(KEDS_ERR_ARG, "debug_scores: bad argument");
  if (!is_device_ptr(q) || !is_device_ptr(out)) return fail(KEDS_ERR_ARG, "debug_scores: device pointers only");
  if (ix->n == 0) return fail(KEDS_ERR_ARG, "debug_scores: empty index");
  DeviceGuard g(ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  memset(&ix->stats, 0, sizeof ix->stats);
  keds_index* v[2] = {ix, nullptr};
  float* Dv[2] = {nullptr, nullptr};
  long long* Iv[2] = {nullptr, nullptr};
  CKS(search_pass(v, 1, q, nq, 1, Dv, Iv, 0u, st, out, ix->n, nullptr));
  return finish_sync(ix, st);
}

int keds_debug_plan(int n_db, int64_t nq, int k, int64_t n_rows, int num_sms, int32_t out[6]) {
  if (!out || n_db < 1 || n_db > 2 || nq <= 0 || k <= 0 || n_rows <= 0 || num_sms <= 0)
    return fail(KEDS_ERR_ARG, "debug_plan: bad argument");
  keds_index ix;  // never touches the device: only the planner's inputs are read
  ix.num_sms = num_sms;
  const char* no_pair = getenv("KEDS_NO_PAIR");
  ix.use_pair = !(no_pair && no_pair[0] == '1');
  const Plan pl = make_plan(&ix, n_db, nq, k, n_rows, n_rows, 0u);
  out[0] = pl.exact_only;
  out[1] = pl.pair ? 1 : 0;
  out[2] = pl.S;
  out[3] = pl.n_qt;
  out[4] = pl.n_items;
  out[5] = pl.grid;
  return 0;
}

int keds_gather_pool(const float* base, int64_t n_base, const int64_t* I, const float* W,
                     const int32_t* perm, int64_t B, int k, int H, int d, float* out, void* stream) {
  if (!base || !I || !out || B < 0 || k <= 0 || d <= 0 || n_base < 0)
    return fail(KEDS_ERR_ARG, "gather_pool: bad argument");
  if (B == 0) return 0;
  const int dev = device_of(out);
  if (dev < 0 || device_of(base) != dev) return fail(KEDS_ERR_ARG, "gather_pool: base and out must be memory of one device");
  DeviceGuard g(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!W) {
    const long long warps = static_cast<long long>(B) * k;
    k_gather_rows<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, st>>>(
        base, static_cast<long long>(n_base), reinterpret_cast<const long long*>(I), perm, B, k, d, out);
  } else {
    if (H <= 0) return fail(KEDS_ERR_ARG, "gather_pool: H must be positive with weights");
    const size_t smem = static_cast<size_t>(k) * 8 + static_cast<size_t>(H) * k * 4;
    if (smem > 48 * 1024) return fail(KEDS_ERR_ARG, "gather_pool: H*k too large");
    k_weighted_pool<<<static_cast<unsigned>(B), 256, smem, st>>>(
        base, static_cast<long long>(n_base), reinterpret_cast<const long long*>(I), W, B, k, H, d, out);
  }
  CK(cudaGetLastError());
  return 0;
}

int keds_topk_merge(const float* Dp, const int64_t* Ip, int parts, int64_t nq, int k, int metric,
                    float* D, int64_t* I, void* stream) {
  return keds_topk_merge_strided(Dp, Ip, nq * k, nq * k, parts, nq, k, metric, D, I, stream);
}

static int merge_impl(const float* Dp, const int64_t* Ip, int64_t stride_d, int64_t stride_i, int parts,
                      int64_t nq, int k, int metric, float* D, int64_t* I, const MergeWait& mw,
                      void* stream) {
  if (!Dp || !Ip || !D || !I || parts <= 0 || nq < 0 || k <= 0)
    return fail(KEDS_ERR_ARG, "topk_merge: bad argument");
  if (nq == 0) return 0;
  const size_t tot = static_cast<size_t>(parts) * k;
  const size_t smem = tot * 8 + tot * 4;  // labels, rank scores
  if (smem > 200 * 1024) return fail(KEDS_ERR_ARG, "topk_merge: parts*k too large");
  const int dev = device_of(D);
  if (dev < 0) return fail(KEDS_ERR_ARG, "topk_merge: device pointers only");
  DeviceGuard g(dev);
  // the attribute is per device; IndexShards / IndexReplicas search from one host thread per GPU
  static std::atomic<bool> attr[64];
  if (dev >= 64 || !attr[dev].load(std::memory_order_acquire)) {
    CK(cudaFuncSetAttribute(k_topk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev < 64) attr[dev].store(true, std::memory_order_release);
  }
  // launched behind the local search / the push kernel with the programmatic attribute: its
  // launch latency hides under their tails (it starts with griddepcontrol.wait)
  CKS(launch_k(true, k_topk_merge, dim3(static_cast<unsigned>(nq)), dim3(128), smem,
               static_cast<cudaStream_t>(stream), Dp, reinterpret_cast<const long long*>(Ip),
               static_cast<long long>(stride_d), static_cast<long long>(stride_i), parts,
               static_cast<long long>(nq), k, metric, D, reinterpret_cast<long long*>(I), mw));
  return 0;
}

int keds_topk_merge_strided(const float* Dp, const int64_t* Ip, int64_t stride_d, int64_t stride_i,
                            int parts, int64_t nq, int k, int metric, float* D, int64_t* I,
                            void* stream) {
  MergeWait mw;
  memset(&mw, 0, sizeof mw);
  return merge_impl(Dp, Ip, stride_d, stride_i, parts, nq, k, metric, D, I, mw, stream);
}

int keds_p2p_push(const void* src, int64_t bytes, void* const* peer_dst, uint32_t* const* peer_flag,
                  int n_ranks, int my_rank, uint32_t epoch, uint32_t* ticket, void* stream) {
  if (!src || !peer_dst || !peer_flag || !ticket || bytes <= 0 || (bytes & 15) || n_ranks < 1 ||
      n_ranks > P2P_MAX_RANKS || my_rank < 0 || my_rank >= n_ranks)
    return fail(KEDS_ERR_ARG, "p2p_push: bad argument (bytes must be a positive multiple of 16, <= %d ranks)",
                P2P_MAX_RANKS);
  P2PPush p;
  memset(&p, 0, sizeof p);
  p.n_ranks = n_ranks;
  p.my_rank = my_rank;
  p.epoch = epoch;
  p.n16 = bytes / 16;
  p.src = static_cast<const uint4*>(src);
  for (int r = 0; r < n_ranks; ++r) {
    p.dst[r] = static_cast<uint4*>(peer_dst[r]);
    p.flag[r] = peer_flag[r];
    if (r != my_rank && (!p.dst[r] || !p.flag[r])) return fail(KEDS_ERR_ARG, "p2p_push: null peer pointer");
  }
  p.ticket = ticket;
  const int dev = device_of(src);
  if (dev < 0) return fail(KEDS_ERR_ARG, "p2p_push: src must be device memory");
  DeviceGuard g(dev);
  const unsigned blocks = static_cast<unsigned>(std::min<long long>(64, (p.n16 + 255) / 256));
  CKS(launch_k(true, k_p2p_push, dim3(std::max(1u, blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
  return 0;
}

int keds_topk_merge_wait(const float* Dp, const int64_t* Ip, int64_t stride_d, int64_t stride_i, int parts,
                         int64_t nq, int k, int metric, float* D, int64_t* I, const uint32_t* flags,
                         int my_rank, uint32_t epoch, uint32_t* err_word, void* stream) {
  if (!flags || !err_word || my_rank < 0 || my_rank >= parts)
    return fail(KEDS_ERR_ARG, "topk_merge_wait: bad argument");
  MergeWait mw;
  memset(&mw, 0, sizeof mw);
  mw.flags = flags;
  mw.my_rank = my_rank;
  mw.epoch = epoch;
  mw.err_word = err_word;
  return merge_impl(Dp, Ip, stride_d, stride_i, parts, nq, k, metric, D, I, mw, stream);
}

// ---- row-sharded search with the exchange fused into the search kernels ------------------------
struct keds_exchange {
  int n_ranks = 0, my_rank = 0, device = 0;
  uint8_t* base[P2P_MAX_RANKS] = {nullptr};  // rank r's symmetric buffer as mapped into this process
  int64_t buf_bytes = 0, slot_bytes = 0, i_off = 0;
  uint32_t epoch = 0;
  DevBuf words;  // [0] ticket, [1] error word, [2..3] pad, then {sum, max, count} wait statistics (u64)
};

namespace {
constexpr int64_t EX_FLAG_BYTES = 256;  // 2 parities x P2P_MAX_RANKS flag words at the start of every buffer
uint8_t* ex_slot(const keds_exchange* ex, int owner, int parity, int r) {
  return ex->base[owner] + EX_FLAG_BYTES + (static_cast<int64_t>(parity) * ex->n_ranks + r) * ex->slot_bytes;
}
uint32_t* ex_flag(const keds_exchange* ex, int owner, int parity, int r) {
  return reinterpret_cast<uint32_t*>(ex->base[owner]) + parity * P2P_MAX_RANKS + r;
}
}  // namespace

int keds_exchange_create(int n_ranks, int my_rank, int device, void* const* peer_base, int64_t buf_bytes,
                         keds_exchange_t** out) {
  if (!out) return fail(KEDS_ERR_ARG, "exchange_create: out is null");
  *out = nullptr;
  if (n_ranks < 1 || n_ranks > P2P_MAX_RANKS || my_rank < 0 || my_rank >= n_ranks || !peer_base)
    return fail(KEDS_ERR_ARG, "exchange_create: 1..%d ranks, my_rank inside", P2P_MAX_RANKS);
  const int64_t slot = (buf_bytes - EX_FLAG_BYTES) / (2 * n_ranks) / 48 * 48;  // D : I = 1 : 2, both 16-byte aligned
  if (slot < 48) return fail(KEDS_ERR_ARG, "exchange_create: buffer of %lld bytes is too small", (long long)buf_bytes);
  for (int r = 0; r < n_ranks; ++r)
    if (!peer_base[r] || (reinterpret_cast<uintptr_t>(peer_base[r]) & 15))
      return fail(KEDS_ERR_ARG, "exchange_create: peer buffer %d is null or not 16-byte aligned", r);
  DeviceGuard g(device);
  if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", device);
  keds_exchange* ex = new keds_exchange();
  ex->n_ranks = n_ranks;
  ex->my_rank = my_rank;
  ex->device = device;
  for (int r = 0; r < n_ranks; ++r) ex->base[r] = static_cast<uint8_t*>(peer_base[r]);
  ex->buf_bytes = buf_bytes;
  ex->slot_bytes = slot;
  ex->i_off = slot / 3;
  int s = ex->words.ensure(64);
  if (s == 0 && cudaMemset(ex->words.p, 0, 64) != cudaSuccess) s = fail(KEDS_ERR_CUDA, "memset failed");
  if (s != 0) {
    ex->words.release();
    delete ex;
    return s;
  }
  *out = ex;
  return 0;
}

void keds_exchange_free(keds_exchange_t* ex) {
  if (!ex) return;
  DeviceGuard g(ex->device);
  ex->words.release();
  delete ex;
}

int64_t keds_exchange_capacity(const keds_exchange_t* ex) { return ex ? ex->i_off / 4 : -1; }

int keds_index_search_sharded(keds_index_t* ix, keds_exchange_t* ex, const float* q, int64_t nq, int k, float* D,
                              int64_t* I, void* stream) {
  if (!ix || !ex || !q || !D || !I || nq <= 0 || k <= 0)
    return fail(KEDS_ERR_ARG, "search_sharded: null handle / pointer or bad nq / k");
  if (ex->device != ix->device) return fail(KEDS_ERR_ARG, "search_sharded: exchange and index live on different devices");
  if (!is_device_ptr(q) || !is_device_ptr(D) || !is_device_ptr(I))
    return fail(KEDS_ERR_ARG, "search_sharded: device pointers only");
  if (nq * k > ex->i_off / 4)
    return fail(KEDS_ERR_ARG, "search_sharded: nq*k = %lld exceeds the exchange capacity %lld", (long long)(nq * k),
                (long long)(ex->i_off / 4));
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t epoch = ++ex->epoch;
  const int parity = static_cast<int>(epoch & 1u);
  const int me = ex->my_rank, R = ex->n_ranks;
  // this rank's block lands in slot `me` of every rank's buffer (its own included: the merge reads it there)
  uint8_t* mine = ex_slot(ex, me, parity, me);
  float* Dl = reinterpret_cast<float*>(mine);
  int64_t* Il = reinterpret_cast<int64_t*>(mine + ex->i_off);
  PeerOut po;
  memset(&po, 0, sizeof po);
  po.n = R;
  po.my_rank = me;
  po.publish = 1;
  po.epoch = epoch;
  po.ticket = ex->words.as<unsigned int>();
  for (int r = 0; r < R; ++r) {
    uint8_t* there = ex_slot(ex, r, parity, me);
    po.D[r] = reinterpret_cast<float*>(there);
    po.I[r] = reinterpret_cast<long long*>(there + ex->i_off);
    po.flag[r] = ex_flag(ex, r, parity, me);
  }
  if (ix->n == 0) {
    // an empty shard still has to deliver (padding) and publish: plain fill + the stand-alone push
    const long long tot = static_cast<long long>(nq) * k;
    k_fill_pad<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, st>>>(
        Dl, reinterpret_cast<long long*>(Il), tot, ix->metric == METRIC_L2 ? FLT_MAX : -FLT_MAX);
    CK(cudaGetLastError());
    if (R > 1) {
      void* dst[P2P_MAX_RANKS];
      uint32_t* flg[P2P_MAX_RANKS];
      for (int r = 0; r < R; ++r) {
        dst[r] = ex_slot(ex, r, parity, me);
        flg[r] = ex_flag(ex, r, parity, me);
      }
      CKS(keds_p2p_push(mine, ex->slot_bytes, dst, flg, R, me, epoch, po.ticket, stream));
    }
  } else {
    keds_index* v[2] = {ix, nullptr};
    float* Dv[2] = {Dl, nullptr};
    int64_t* Iv[2] = {Il, nullptr};
    CKS(search_impl(v, 1, q, nq, k, Dv, Iv, 0u, stream, nullptr, R > 1 ? &po : nullptr));
  }
  MergeWait mw;
  memset(&mw, 0, sizeof mw);
  if (R > 1) {
    mw.flags = ex_flag(ex, me, parity, 0);
    mw.my_rank = me;
    mw.epoch = epoch;
    mw.err_word = ex->words.as<unsigned int>() + 1;
    mw.stats = reinterpret_cast<unsigned long long*>(ex->words.as<uint8_t>() + 16);
  }
  uint8_t* parts = ex_slot(ex, me, parity, 0);
  return merge_impl(reinterpret_cast<const float*>(parts), reinterpret_cast<const int64_t*>(parts + ex->i_off),
                    ex->slot_bytes / 4, ex->slot_bytes / 8, R, nq, k, ix->metric, D, I, mw, stream);
}

int keds_exchange_stats(keds_exchange_t* ex, void* stream, double* wait_us_avg, double* wait_us_max,
                        int64_t* steps, uint32_t* err_word) {
  if (!ex) return fail(KEDS_ERR_ARG, "exchange_stats: null handle");
  DeviceGuard g(ex->device);
  CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  uint8_t h[64];
  CK(cudaMemcpy(h, ex->words.p, 64, cudaMemcpyDeviceToHost));
  const uint32_t* w = reinterpret_cast<const uint32_t*>(h);
  const unsigned long long* st = reinterpret_cast<const unsigned long long*>(h + 16);
  if (wait_us_avg) *wait_us_avg = st[2] ? static_cast<double>(st[0]) / static_cast<double>(st[2]) * 1e-3 : 0.0;
  if (wait_us_max) *wait_us_max = static_cast<double>(st[1]) * 1e-3;
  if (steps) *steps = static_cast<int64_t>(st[2]);
  if (err_word) *err_word = w[1];
  CK(cudaMemset(ex->words.as<uint8_t>() + 16, 0, 24));  // statistics restart; ticket and error word stay
  if (w[1] != 0)
    return fail(KEDS_ERR_KERNEL, "sharded exchange: peer %u did not deliver within the watchdog window (error word 0x%x)",
                w[1] - 0x500u, w[1]);
  return 0;
}

// Gallery ranking against the rows of an index, on the tensor cores: the same scoring kernel with
// a counting epilogue between two small kernels (target scores + error band before, exact
// settlement of the band rows after), and the exact fp32 recount for queries whose band overflowed.
int keds_index_rank(keds_index_t* ix, const float* q, int64_t nq, const int64_t* target, const int64_t* exclude,
                    int64_t* rank_out, void* stream) {
  if (!ix || !q || !target || !rank_out || nq < 0) return fail(KEDS_ERR_ARG, "index_rank: null argument or bad nq");
  if (nq == 0) return 0;
  if (ix->n == 0) return fail(KEDS_ERR_ARG, "index_rank: empty index");
  if (!is_device_ptr(q) || !is_device_ptr(target) || !is_device_ptr(rank_out) || (exclude && !is_device_ptr(exclude)))
    return fail(KEDS_ERR_ARG, "index_rank: device pointers only");
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CKS(set_kernel_attrs(ix));
  const float* G = ix->x_f32.as<float>();
  const long long* tg = reinterpret_cast<const long long*>(target);
  const long long* ex = reinterpret_cast<const long long*>(exclude);
  long long* out = reinterpret_cast<long long*>(rank_out);
  const int d = ix->d;
  const size_t simt_smem = static_cast<size_t>((d + 3) & ~3) * 4 * 8;  // 8 queries per block
  if (simt_smem > 48 * 1024) return fail(KEDS_ERR_ARG, "index_rank: d too large");
  const int T = static_cast<int>((ix->n + BN - 1) / BN);
  if (T < 2) {
    // less than two row tiles: nothing for the tensor cores to win
    k_gallery_rank<<<static_cast<unsigned>((nq + 7) / 8), 256, simt_smem, st>>>(q, nq, G, ix->n, d, tg, ex, out,
                                                                                  nullptr, nullptr);
    CK(cudaGetLastError());
    return 0;
  }
  for (int64_t q0 = 0; q0 < nq; q0 += Q_PASS_MAX) {
    const int64_t nb = std::min<int64_t>(Q_PASS_MAX, nq - q0);
    const float* qp = q + q0 * d;
    const int n_qt = static_cast<int>((nb + BM - 1) / BM);
    const bool pair = ix->use_pair && n_qt >= 2 && ix->num_sms >= 2;
    const int sub = pair ? ScoreCfg<true>::kSub : ScoreCfg<false>::kSub;
    const int n_qg = pair ? (n_qt + 1) / 2 : n_qt;
    const int units = pair ? ix->num_sms / 2 : ix->num_sms;
    // as many candidate lists as the tiles allow: the band around a mid-ranked target is crowded
    const long long by_mem = static_cast<long long>(CAND_BUDGET / (size_t(LKEEP) * BM * 8)) / (static_cast<long long>(n_qt) * sub);
    const int S = static_cast<int>(std::max<long long>(1, std::min<long long>(std::min(T, S_MAX), by_mem)));
    const int n_items = n_qg * S;
    const int grid = std::min(n_items, units) * (pair ? 2 : 1);
    const int n_lists = S * sub;
    CKS(ix->ctrl.ensure(CTRL_STRIDE * 4));  // (finish_sync reads whole status blocks)
    ix->ctrl_passes = 1;
    CKS(ix->flagged[0].ensure(static_cast<size_t>(nb) * 4));
    CKS(ensure_q_map(ix, static_cast<int64_t>(n_qt) * BM, st));
    CKS(ix->qstat.ensure(static_cast<size_t>(nb) * sizeof(float4)));
    CKS(ix->rk.ensure(static_cast<size_t>(nb) * 5 * 4));
    const size_t lines = static_cast<size_t>(n_lists) * n_qt;
    CKS(ix->cand.ensure(lines * LKEEP * BM * 8));
    CKS(ix->cand_cnt.ensure(lines * BM * 4));
    CKS(ix->cand_theta.ensure(lines * BM * 4));
    float* rk_st = ix->rk.as<float>();
    float* rk_lo = rk_st + nb;
    float* rk_hi = rk_lo + nb;
    int* rk_t = reinterpret_cast<int*>(rk_hi + nb);
    int* rk_e = rk_t + nb;
    {
      const unsigned blocks = static_cast<unsigned>(std::min<long long>((nb * 32 + 255) / 256, 4096));
      CKS(launch_k(ix->use_pdl, k_prep_rows, dim3(blocks), dim3(256), 0, st, qp, static_cast<long long>(nb), d,
                   ix->d_pad, ix->fmt, ix->q_bf16.as<uint16_t>(), ix->qstat.as<float4>(), static_cast<float*>(nullptr),
                   static_cast<unsigned int*>(nullptr), ix->ctrl.as<unsigned int>(), CTRL_WORDS,
                   static_cast<float*>(nullptr), 0, static_cast<unsigned long long*>(nullptr),
                   static_cast<float*>(nullptr)));
    }
    const unsigned wblocks = static_cast<unsigned>((nb * 32 + 255) / 256);
    CKS(launch_k(ix->use_pdl, k_rank_targets, dim3(wblocks), dim3(256), 0, st, qp, static_cast<long long>(nb), G, d,
                 tg + q0, ex ? ex + q0 : nullptr, ix->qstat.as<float4>(), ix->dbstat.as<unsigned int>(), ix->eps_scale,
                 rk_st, rk_lo, rk_hi, rk_t, rk_e));
    ScoreParams sp;
    memset(&sp, 0, sizeof sp);
    sp.n_db = 1;
    sp.n_qt = n_qt;
    sp.n_qg = n_qg;
    sp.S = S;
    sp.n_items = n_items;
    sp.kblocks = ix->d_pad / BK;
    sp.fmt_bits = ix->fmt == FMT_BF16 ? kIdescBf16Bits : 0u;
    sp.nq = static_cast<int>(nb);
    sp.n_rows[0] = static_cast<int>(ix->n);
    sp.n_tiles[0] = T;
    sp.cand = ix->cand.as<uint2>();
    sp.cand_cnt = ix->cand_cnt.as<int>();
    sp.cand_theta = ix->cand_theta.as<float>();
    sp.err = ix->ctrl.as<uint32_t>() + 2;
    sp.rk_lo = rk_lo;
    sp.rk_hi = rk_hi;
    sp.rk_target = rk_t;
    sp.rk_exclude = ex ? rk_e : nullptr;
    if (pair)
      CKS(launch_kc(ix->use_pdl, 2, k_score_topk<true, true>, dim3(grid), dim3(SCORE_PAIR_THREADS), SCORE_PAIR_SMEM_BYTES,
                    st, ix->tm_q, ix->tm_xh, ix->tm_xh, sp));
    else
      CKS(launch_k(ix->use_pdl, k_score_topk<false, true>, dim3(grid), dim3(SCORE_THREADS), SCORE_SMEM_BYTES, st,
                   ix->tm_q, ix->tm_x, ix->tm_x, sp));
    CKS(launch_k(ix->use_pdl, k_rank_finish, dim3(wblocks), dim3(256), 0, st, qp, static_cast<long long>(nb), G, d,
                 n_lists, n_qt, static_cast<const uint2*>(sp.cand), static_cast<const int*>(sp.cand_cnt),
                 static_cast<const float*>(sp.cand_theta), static_cast<const float*>(rk_st),
                 static_cast<const int*>(rk_t), out + q0, ix->flagged[0].as<int>(), ix->ctrl.as<int>()));
    // exact recount of the queries whose band overflowed a list (none, normally: the blocks leave at once)
    CKS(launch_k(ix->use_pdl, k_gallery_rank, dim3(static_cast<unsigned>((nb + 7) / 8)), dim3(256), simt_smem, st, qp,
                 static_cast<long long>(nb), G, static_cast<long long>(ix->n), d, tg + q0, ex ? ex + q0 : nullptr,
                 out + q0, static_cast<const int*>(ix->flagged[0].as<int>()), static_cast<const int*>(ix->ctrl.as<int>())));
    ix->stats.launches = 5;
    ix->stats.slices = S;
    ix->stats.items = n_items;
    ix->stats.grid = grid;
    if (q0 + nb < nq) CKS(finish_sync(ix, st));  // the scratch is reused by the next pass
  }
  CK(cudaGetLastError());
  return 0;
}

int keds_index_label_hits(keds_index_t* ix, const float* q, int64_t nq, const int64_t* row_labels,
                          const int64_t* qlabel, const int32_t* ks, int nks, int32_t* hits, void* stream) {
  if (!ix || !q || !row_labels || !qlabel || !ks || !hits || nq < 0 || nks <= 0 || nks > HITS_MAX_CUTS)
    return fail(KEDS_ERR_ARG, "index_label_hits: null argument, bad nq or more than %d cut points", HITS_MAX_CUTS);
  for (int j = 0; j < nks; ++j)
    if (ks[j] <= 0 || (j > 0 && ks[j] <= ks[j - 1])) return fail(KEDS_ERR_ARG, "index_label_hits: ks must be ascending and positive");
  if (ks[nks - 1] > K_MAX) return fail(KEDS_ERR_ARG, "index_label_hits: k=%d exceeds the maximum %d", ks[nks - 1], K_MAX);
  if (nq == 0) return 0;
  if (ix->n == 0) return fail(KEDS_ERR_ARG, "index_label_hits: empty index");
  if (ix->n > 0x7fffffffll - BN) return fail(KEDS_ERR_ARG, "index too large for 32-bit row ids");
  if (!is_device_ptr(q) || !is_device_ptr(row_labels) || !is_device_ptr(qlabel) || !is_device_ptr(hits))
    return fail(KEDS_ERR_ARG, "index_label_hits: device pointers only");
  DeviceGuard g(ix->device);
  if (!g.ok) return fail(KEDS_ERR_NO_GPU, "cannot select CUDA device %d", ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  memset(&ix->stats, 0, sizeof ix->stats);
  const int kmax = ks[nks - 1];
  const int n_passes = static_cast<int>((nq + Q_PASS_MAX - 1) / Q_PASS_MAX);
  CKS(ix->ctrl.ensure(static_cast<size_t>(n_passes) * CTRL_STRIDE * 4));
  // exact rows of the queries the certificate queues (the fallback writes them, k_hits_from_rows counts them)
  const int64_t nb_max = std::min<int64_t>(nq, Q_PASS_MAX);
  CKS(ix->D_stage[0].ensure(static_cast<size_t>(nb_max) * kmax * 4));
  CKS(ix->I_stage[0].ensure(static_cast<size_t>(nb_max) * kmax * 8));
  keds_index* v[2] = {ix, nullptr};
  for (int64_t q0 = 0; q0 < nq; q0 += Q_PASS_MAX) {
    const int64_t nb = std::min<int64_t>(Q_PASS_MAX, nq - q0);
    HitsMode hm;
    hm.nks = nks;
    for (int j = 0; j < nks; ++j) hm.ks[j] = ks[j];
    hm.row_labels = reinterpret_cast<const long long*>(row_labels);
    hm.qlabel = reinterpret_cast<const long long*>(qlabel) + q0;
    hm.hits = hits + q0 * nks;
    float* Dp[2] = {ix->D_stage[0].as<float>(), nullptr};
    long long* Ip[2] = {ix->I_stage[0].as<long long>(), nullptr};
    CKS(search_pass(v, 1, q + q0 * ix->d, nb, kmax, Dp, Ip, 0u, st, nullptr, 0, nullptr, nullptr,
                    static_cast<int>(q0 / Q_PASS_MAX), nullptr, &hm));
  }
  CK(cudaGetLastError());
  return 0;
}

int keds_gallery_rank(const float* Q, int64_t nq, const float* G, int64_t ng, int d,
                      const int64_t* target, const int64_t* exclude, int64_t* rank_out, void* stream) {
  if (!Q || !G || !target || !rank_out || nq < 0 || ng <= 0 || d <= 0)
    return fail(KEDS_ERR_ARG, "gallery_rank: bad argument");
  if (nq == 0) return 0;
  const size_t smem = static_cast<size_t>((d + 3) & ~3) * 4 * 8;  // 8 queries per block
  if (smem > 48 * 1024) return fail(KEDS_ERR_ARG, "gallery_rank: d too large");
  const int dev = device_of(rank_out);
  if (dev < 0 || device_of(Q) != dev || device_of(G) != dev)
    return fail(KEDS_ERR_ARG, "gallery_rank: Q, G and rank_out must be memory of one device");
  DeviceGuard g(dev);
  if (ng >= 1024 && nq >= 64) {
    // worth a tensor-core pass: rank against a temporary index over the gallery (callers that score
    // many query sets against one gallery keep the index and call keds_index_rank themselves)
    keds_index_t* tmp = nullptr;
    CKS(keds_index_create(d, KEDS_METRIC_IP, dev, &tmp));
    int s = keds_index_add(tmp, G, ng);
    if (s == 0) s = keds_index_rank(tmp, Q, nq, target, exclude, rank_out, stream);
    if (s == 0) s = finish_sync(tmp, static_cast<cudaStream_t>(stream));
    keds_index_free(tmp);
    return s;
  }
  k_gallery_rank<<<static_cast<unsigned>((nq + 7) / 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      Q, nq, G, ng, d, reinterpret_cast<const long long*>(target),
      reinterpret_cast<const long long*>(exclude), reinterpret_cast<long long*>(rank_out), nullptr, nullptr);
  CK(cudaGetLastError());
  return 0;
}

int keds_label_hits(const int64_t* I, int64_t nq, int kmax, const int64_t* labels, const int64_t* qlabel,
                    const int32_t* ks, int nks, int32_t* hits, void* stream) {
  if (!I || !labels || !qlabel || !ks || !hits || nq < 0 || kmax <= 0 || nks <= 0)
    return fail(KEDS_ERR_ARG, "label_hits: bad argument");
  if (nq == 0) return 0;
  const int dev = device_of(hits);
  if (dev < 0 || device_of(I) != dev) return fail(KEDS_ERR_ARG, "label_hits: I and hits must be memory of one device");
  DeviceGuard g(dev);
  k_label_hits<<<static_cast<unsigned>((nq + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(I), nq, kmax, reinterpret_cast<const long long*>(labels),
      reinterpret_cast<const long long*>(qlabel), ks, nks, hits);
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// keds_consumer_*: the neighbour consumer (uses the helpers above)
#include "consumer_host.cuh"
// keds_clip_loss_*: the contrastive loss over the gathered features
#include "clip_loss_host.cuh"
// keds_consumer_bind_params / _forward_train / _backward: the consumer while it is being trained
#include "consumer_train_host.cuh"
