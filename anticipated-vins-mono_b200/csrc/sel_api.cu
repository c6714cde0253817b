// sel_api.cu -- C-ABI of the anticipated feature selector (include/bvio.h):
// FeatureSelector::select's numerical part, vins_estimator/src/feature_selector.cpp:139-170.
// Host code packs the inputs into one pinned slab, issues one H2D copy, launches the kernels of
// sel_kernels.cu (one per greedy round) and reads back the selected indices.  Multi-GPU: one
// process per GPU, candidates sharded by contiguous blocks, one ncclAllGather of the per-rank
// winner records per round.  NCCL is resolved with dlopen at bvio_comm_init so that the library
// has no link-time dependency on it.
#include "sel.h"
#include "ctx.h"
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <new>

using namespace bvio;

// ---- minimal NCCL surface (ABI-stable since NCCL 2.x) ---------------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_uid_t*) = nullptr;
  int (*CommInitRank)(ncclComm**, int, nccl_uid_t, int) = nullptr;
  int (*CommDestroy)(ncclComm*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static const int kNcclInt8 = 0, kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0;

static bool nccl_load() {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
  if (!lib) return false;
  NcclApi a;
  a.lib = lib;
  a.GetUniqueId = (int (*)(nccl_uid_t*))dlsym(lib, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(ncclComm**, int, nccl_uid_t, int))dlsym(lib, "ncclCommInitRank");
  a.CommDestroy = (int (*)(ncclComm*))dlsym(lib, "ncclCommDestroy");
  a.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm*, cudaStream_t))dlsym(lib, "ncclAllGather");
  a.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm*, cudaStream_t))dlsym(lib, "ncclAllReduce");
  a.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.AllReduce) return false;
  g_nccl = a;
  return true;
}

namespace bvio {
int sel_launch_counts(const SelProb& sp, unsigned long long* buf, int dir, cudaStream_t st);   // sel_kernels.cu
}

struct bvio_selprob {
  SelProb sp;
  Slab slab;
  bool from_cache = false;
  size_t in_bytes = 0, out_off = 0, out_bytes = 0;
  size_t o_out_idx = 0, o_out_val = 0, o_ctrl = 0;
  std::vector<int32_t> cand_id;
  unsigned long long* counts = nullptr;   // [2] device (sharded finalize)
  cudaGraphExec_t graph = nullptr;
  bool use_graph = true;
  bool sharded = false;
  int launches_per_run = 0;
};

static int sel_validate(bvio_ctx* ctx, const bvio_select_in* in) {
  if (!ctx || !in) return fail(ctx, BVIO_ERR_INVALID, "null argument");
  if (in->H < 1 || in->H > BVIO_HMAX) return fail(ctx, BVIO_ERR_INVALID, "H out of range [1,16]");
  if (!in->horizon_pos || !in->horizon_quat) return fail(ctx, BVIO_ERR_INVALID, "null horizon");
  if (in->N < 0 || in->U < 0 || in->C < 0 || in->kappa < 0 || in->nr_imu < 1)
    return fail(ctx, BVIO_ERR_INVALID, "negative size / nr_imu < 1");
  if (in->N > 0 && (!in->cand_id || !in->cand_xy || !in->cand_prob)) return fail(ctx, BVIO_ERR_INVALID, "null candidates");
  if (in->U > 0 && !in->used_xy) return fail(ctx, BVIO_ERR_INVALID, "null used_xy");
  if (in->C > 0 && (!in->cloud_xy || !in->cloud_depth)) return fail(ctx, BVIO_ERR_INVALID, "null cloud");
  return BVIO_OK;
}

static int sel_upload_impl(bvio_ctx* ctx, const bvio_select_in* in, bool use_cache, bool sharded, bvio_selprob** out, bool force_shard = false) {
  int rc = sel_validate(ctx, in);
  if (rc) return rc;
  if (!out) return fail(ctx, BVIO_ERR_INVALID, "null out");
  *out = nullptr;
  if (sharded && !ctx->comm) return fail(ctx, BVIO_ERR_NCCL, "bvio_comm_init has not been called");
  cudaSetDevice(ctx->device);
  if (sel_configure() != 0) return fail(ctx, BVIO_ERR_CUDA, "sel_configure failed");
  bvio_selprob* pr = new (std::nothrow) bvio_selprob();
  if (!pr) return fail(ctx, BVIO_ERR_INVALID, "out of host memory");
  SelProb& sp = pr->sp;
  memset(&sp, 0, sizeof sp);
  const int H = in->H, T = 3 * H, TT = T * (T + 1) / 2, D = 9 * (H + 1), N = in->N, U = in->U, C = in->C;
  if (sharded && ctx->world > 1 && !force_shard && !getenv("BVIO_SEL_FORCE_SHARD")) {
    // Planner: a greedy round costs one in-warp Cholesky chain however few candidates a warp holds, so sharding only
    // pays when one GPU needs more than one candidate per warp.  Below that every rank runs the whole selection itself
    // (identical results, no exchange) and the summary says transport 0.
    SelProb probe;
    memset(&probe, 0, sizeof probe);
    probe.H = H; probe.T = T; probe.TT = TT; probe.world = 1; probe.c0 = 0; probe.c1 = N; probe.kappa = in->kappa;
    if (sel_plan_persist(probe, ctx->sm_count) && probe.cpw <= 1) sharded = false;
  }
  if (sharded && ctx->world > 1 && !ctx->p2p_ready && !getenv("BVIO_SEL_NCCL")) {
    delete pr;
    return fail(ctx, BVIO_ERR_NCCL, "bvio_select_sharded: the peer-memory mailboxes could not be mapped (no CUDA IPC / peer access "
                                    "between the ranks' devices); set BVIO_SEL_NCCL=1 to run one ncclAllGather per greedy round instead");
  }
  sp.H = H; sp.T = T; sp.TT = TT; sp.D = D; sp.Do = D - T;
  sp.N = N; sp.U = U; sp.C = C; sp.kappa = in->kappa; sp.nr_imu = in->nr_imu;
  sp.rank = sharded ? ctx->rank : 0;
  sp.world = sharded ? ctx->world : 1;
  // contiguous blocks of ceil(N/world) candidates per rank (ascending id => rank-independent tie-breaking)
  const int per = (N + sp.world - 1) / sp.world;
  sp.c0 = std::min(N, sp.rank * per);
  sp.c1 = std::min(N, (sp.rank + 1) * per);
  const int nloc = sp.c1 - sp.c0;
  sp.grid_round = std::max(1, std::min((nloc + SEL_WARPS - 1) / SEL_WARPS, 4 * ctx->sm_count));
  sp.b0 = sp.c0; sp.b1 = sp.c1;
  // multi-GPU: all greedy rounds in the persistent kernel with the exchange over peer memory when the mailboxes
  // are mapped and the shard fits the co-resident grid on EVERY rank (the plan depends only on N / world sizes);
  // otherwise one kernel per round with an ncclAllGather in between
  sp.fused = (sharded && sp.world > 1 && ctx->p2p_ready && !getenv("BVIO_SEL_NCCL") && (sp.world - 1) * per < N) ? 1 : 0;
  if (sp.fused) {
    SelProb probe = sp;                       // the largest shard (rank 0's) decides for all ranks
    probe.c0 = 0; probe.c1 = std::min(N, per);
    if (!sel_plan_persist(probe, ctx->sm_count)) sp.fused = 0;
  }
  if (sp.fused) {
    sp.b0 = 0; sp.b1 = N;
    sp.epoch_base = ctx->sel_epoch;
    sp.mbox = (double*)ctx->mbox_local;
    sp.mflag = (unsigned long long*)((char*)ctx->mbox_local + SEL_MBOX_FLAG_OFF);
    for (int r = 0; r < sp.world; r++) {
      sp.peer_mbox[r] = (double*)ctx->mbox_peer[r];
      sp.peer_flag[r] = (unsigned long long*)((char*)ctx->mbox_peer[r] + SEL_MBOX_FLAG_OFF);
    }
  }
  sel_plan_persist(sp, ctx->sm_count);
  if (sp.fused && sp.grid_persist == 0) return (delete pr, fail(ctx, BVIO_ERR_INVALID, "fused selector plan mismatch"));
  sp.delta_imu = in->delta_imu; sp.acc_var = in->acc_var; sp.acc_bias_var = in->acc_bias_var;
  for (int i = 0; i < 4; i++) sp.q_ic[i] = in->q_ic[i];
  for (int i = 0; i < 3; i++) sp.t_ic[i] = in->t_ic[i];
  sp.cam = in->cam;
  for (int i = 0; i < 3; i++) sp.k1_pos[i] = in->state_k1_pos ? in->state_k1_pos[i] : in->horizon_pos[3 + i];
  for (int i = 0; i < 4; i++) sp.k1_quat[i] = in->state_k1_quat ? in->state_k1_quat[i] : in->horizon_quat[4 + i];
  sp.has_prior = in->omega_prior ? 1 : 0;
  for (int i = 0; i < 81; i++) sp.omega_prior[i] = in->omega_prior ? in->omega_prior[i] : 0.0;
  pr->sharded = sharded;
  // NCCL calls and the cooperative persistent kernel stay outside graph capture (5 launches anyway)
  pr->use_graph = !sharded && sp.grid_persist == 0;
  pr->cand_id.assign(in->cand_id, in->cand_id + N);

  Carver cv;
  const size_t Dd = sizeof(double), I = sizeof(int);
  size_t o_hpos = cv.take((size_t)(H + 1) * 3 * Dd), o_hquat = cv.take((size_t)(H + 1) * 4 * Dd);
  size_t o_cxy = cv.take((size_t)N * 2 * Dd), o_cp = cv.take((size_t)N * Dd);
  size_t o_uxy = cv.take((size_t)U * 2 * Dd);
  size_t o_clxy = cv.take((size_t)C * 2 * Dd), o_cld = cv.take((size_t)C * Dd);
  pr->in_bytes = cv.off;
  pr->out_off = cv.off;
  pr->o_out_idx = cv.take((size_t)std::max(1, in->kappa) * I);
  pr->o_out_val = cv.take((size_t)std::max(1, in->kappa) * Dd);
  pr->o_ctrl = cv.take(sizeof(SelCtrl));
  pr->out_bytes = cv.off - pr->out_off;
  const size_t h_bytes = cv.off;
  size_t o_Cc = cv.take((size_t)N * TT * Dd), o_Cu = cv.take((size_t)U * TT * Dd);
  size_t o_valid = cv.take((size_t)N * I), o_valid_u = cv.take((size_t)U * I), o_taken = cv.take((size_t)N * I);
  size_t o_depth = cv.take((size_t)(N + U) * Dd);
  size_t o_pair = cv.take((size_t)H * 324 * Dd), o_omega = cv.take((size_t)D * D * Dd), o_R = cv.take((size_t)TT * Dd);
  size_t o_blk = cv.take((size_t)std::max(sp.grid_round, 2 * sp.grid_persist) * 4 * Dd);
  size_t o_send = cv.take((size_t)(SEL_REC_HDR + TT) * Dd), o_all = cv.take((size_t)sp.world * (SEL_REC_HDR + TT) * Dd);
  size_t o_counts = cv.take(2 * sizeof(unsigned long long));
  const size_t d_bytes = cv.off;

  cudaError_t ce;
  if (use_cache) ce = slab_acquire(ctx->sel_cache, ctx->sel_cache_busy, d_bytes, h_bytes, pr->slab, pr->from_cache);
  else {
    bool busy = true;
    Slab none;
    ce = slab_acquire(none, busy, d_bytes, h_bytes, pr->slab, pr->from_cache);
  }
  if (ce != cudaSuccess) { delete pr; return fail(ctx, BVIO_ERR_CUDA, std::string("slab alloc: ") + cudaGetErrorString(ce)); }
  char* d = pr->slab.d;
  char* h = pr->slab.h;
  memcpy(h + o_hpos, in->horizon_pos, (size_t)(H + 1) * 3 * Dd);
  memcpy(h + o_hquat, in->horizon_quat, (size_t)(H + 1) * 4 * Dd);
  if (N) { memcpy(h + o_cxy, in->cand_xy, (size_t)N * 2 * Dd); memcpy(h + o_cp, in->cand_prob, (size_t)N * Dd); }
  if (U) memcpy(h + o_uxy, in->used_xy, (size_t)U * 2 * Dd);
  if (C) { memcpy(h + o_clxy, in->cloud_xy, (size_t)C * 2 * Dd); memcpy(h + o_cld, in->cloud_depth, (size_t)C * Dd); }
  sp.hpos = (const double*)(d + o_hpos); sp.hquat = (const double*)(d + o_hquat);
  sp.cand_xy = (const double2*)(d + o_cxy); sp.cand_prob = (const double*)(d + o_cp);
  sp.used_xy = (const double2*)(d + o_uxy);
  sp.cloud_xy = (const double2*)(d + o_clxy); sp.cloud_depth = (const double*)(d + o_cld);
  sp.out_idx = (int*)(d + pr->o_out_idx); sp.out_val = (double*)(d + pr->o_out_val); sp.ctrl = (SelCtrl*)(d + pr->o_ctrl);
  sp.Cc = (double*)(d + o_Cc); sp.Cu = (double*)(d + o_Cu);
  sp.valid = (int*)(d + o_valid); sp.valid_u = (int*)(d + o_valid_u); sp.taken = (int*)(d + o_taken);
  sp.depth = (double*)(d + o_depth); sp.pair = (double*)(d + o_pair); sp.omega = (double*)(d + o_omega);
  sp.R = (double*)(d + o_R); sp.blk_best = (double*)(d + o_blk);
  sp.rec_send = (double*)(d + o_send); sp.rec_all = (double*)(d + o_all);
  pr->counts = (unsigned long long*)(d + o_counts);
  cudaError_t e = cudaMemcpyAsync(d, h, pr->in_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    slab_release(pr->slab, ctx->sel_cache_busy, pr->from_cache);
    delete pr;
    return fail(ctx, BVIO_ERR_CUDA, std::string("select upload: ") + cudaGetErrorString(e));
  }
  *out = pr;
  return BVIO_OK;
}

static int sel_enqueue(bvio_ctx* ctx, bvio_selprob* pr, cudaStream_t st, int* nccl_rc) {
  const SelProb& sp = pr->sp;
  int n = 0;
  *nccl_rc = 0;
  n += sel_launch_reset(sp, st);
  n += sel_launch_build(sp, st);
  const size_t RS = SEL_REC_HDR + sp.TT;
  if (sp.grid_persist > 0) n += sel_launch_persist(sp, st);
  for (int it = 0; it < sp.kappa && sp.grid_persist == 0; it++) {
    n += sel_launch_round(sp, st);
    if (sp.world > 1) {
      int r = g_nccl.AllGather(sp.rec_send, sp.rec_all, RS, kNcclFloat64, ctx->comm, st);
      if (r) { *nccl_rc = r; return n; }
      n += sel_launch_apply(sp, st);
    }
  }
  n += sel_launch_final(sp, st);
  if (sp.world > 1 && !sp.fused) {
    n += sel_launch_counts(sp, pr->counts, 0, st);
    int r = g_nccl.AllReduce(pr->counts, pr->counts, 2, kNcclUint64, kNcclSum, ctx->comm, st);
    if (r) { *nccl_rc = r; return n; }
    n += sel_launch_counts(sp, pr->counts, 1, st);
  }
  return n;
}

extern "C" {

static void mbox_release(bvio_ctx* ctx) {
  for (int r = 0; r < 8; r++) {
    if (ctx->mbox_peer[r] && ctx->mbox_peer[r] != ctx->mbox_local) cudaIpcCloseMemHandle(ctx->mbox_peer[r]);
    ctx->mbox_peer[r] = nullptr;
  }
  if (ctx->mbox_local) cudaFree(ctx->mbox_local);
  ctx->mbox_local = nullptr;
  ctx->p2p_ready = false;
}

void bvio_sel_ctx_destroy(bvio_ctx* ctx) {
  if (!ctx) return;
  mbox_release(ctx);
  if (ctx->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
}

// Mailboxes of the fused exchange: every rank allocates 4 KB, publishes its CUDA IPC handle through one ncclAllGather
// and maps the others'.  Any failure (no peer access between the processes' devices, IPC disabled in the container)
// just leaves p2p_ready false: bvio_select_sharded then runs one kernel per round with an ncclAllGather in between.
static void mbox_setup(bvio_ctx* ctx) {
  mbox_release(ctx);
  if (ctx->world < 2 || ctx->world > SEL_MAX_WORLD) return;
  cudaIpcMemHandle_t mine, all[SEL_MAX_WORLD];
  char* dbuf = nullptr;
  bool ok = cudaMalloc(&ctx->mbox_local, SEL_MBOX_BYTES) == cudaSuccess &&
            cudaMemset(ctx->mbox_local, 0, SEL_MBOX_BYTES) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine, ctx->mbox_local) == cudaSuccess &&
            cudaMalloc((void**)&dbuf, sizeof(mine) * (ctx->world + 1)) == cudaSuccess;
  // the collective must be entered by every rank, whatever happened locally: a failed rank publishes zeros
  if (!ok) memset(&mine, 0, sizeof mine);
  int all_ok = 0;
  if (dbuf) {
    cudaMemcpyAsync(dbuf, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream);
    int r = g_nccl.AllGather(dbuf, dbuf + sizeof mine, sizeof mine, kNcclInt8, ctx->comm, ctx->stream);
    if (r == 0 && cudaMemcpyAsync(all, dbuf + sizeof mine, sizeof(mine) * ctx->world, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
        cudaStreamSynchronize(ctx->stream) == cudaSuccess) all_ok = 1;
  }
  if (all_ok && ok) {
    cudaIpcMemHandle_t zero;
    memset(&zero, 0, sizeof zero);
    for (int r = 0; r < ctx->world && ok; r++) {
      if (r == ctx->rank) { ctx->mbox_peer[r] = ctx->mbox_local; continue; }
      if (!memcmp(&all[r], &zero, sizeof zero)) { ok = false; break; }
      ok = cudaIpcOpenMemHandle(&ctx->mbox_peer[r], all[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    }
  } else ok = false;
  cudaGetLastError();
  if (dbuf) cudaFree(dbuf);
  // every rank must agree (a rank that could not map a peer would otherwise wait for records nobody sends)
  unsigned long long* flag = nullptr;
  unsigned long long hv = ok ? 1ull : 0ull, res = 0;
  if (cudaMalloc((void**)&flag, 2 * sizeof(unsigned long long)) == cudaSuccess) {
    cudaMemcpyAsync(flag, &hv, sizeof hv, cudaMemcpyHostToDevice, ctx->stream);
    // sum of the ok flags == world  <=>  everybody is ready
    if (g_nccl.AllReduce(flag, flag + 1, 1, kNcclUint64, kNcclSum, ctx->comm, ctx->stream) == 0 &&
        cudaMemcpyAsync(&res, flag + 1, sizeof res, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
        cudaStreamSynchronize(ctx->stream) == cudaSuccess && res == (unsigned long long)ctx->world) ctx->p2p_ready = true;
    cudaFree(flag);
  }
  if (!ctx->p2p_ready) mbox_release(ctx);
  ctx->sel_epoch = 0;
}

int bvio_select_upload(bvio_ctx* ctx, const bvio_select_in* in, bvio_selprob** out) {
  return sel_upload_impl(ctx, in, false, ctx && ctx->comm && ctx->world > 1, out);
}
int bvio_select_upload_mode(bvio_ctx* ctx, const bvio_select_in* in, int32_t mode, bvio_selprob** out) {
  if (mode < 0 || mode > 2) return fail(ctx, BVIO_ERR_INVALID, "bvio_select_upload_mode: mode must be 0, 1 or 2");
  return sel_upload_impl(ctx, in, false, mode != 0 && ctx && ctx->comm && ctx->world > 1, out, mode == 2);
}

int bvio_select_run(bvio_ctx* ctx, bvio_selprob* pr) {
  if (!ctx || !pr) return fail(ctx, BVIO_ERR_INVALID, "null problem");
  cudaSetDevice(ctx->device);
  int nrc = 0;
  BVIO_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  if (pr->use_graph) {
    if (!pr->graph) {
      cudaGraph_t g = nullptr;
      BVIO_CUDA_OK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      pr->launches_per_run = sel_enqueue(ctx, pr, ctx->stream, &nrc);
      // always leave capture mode, whatever happened inside, and never leak the graph
      cudaError_t ee = cudaStreamEndCapture(ctx->stream, &g);
      if (ee == cudaSuccess && !nrc) ee = cudaGraphInstantiate(&pr->graph, g, 0);
      if (g) cudaGraphDestroy(g);
      if (nrc) { pr->graph = nullptr; return fail(ctx, BVIO_ERR_NCCL, std::string("nccl: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?")); }
      if (ee != cudaSuccess) { pr->graph = nullptr; BVIO_CUDA_OK(ctx, ee); }
      BVIO_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    }
    BVIO_CUDA_OK(ctx, cudaGraphLaunch(pr->graph, ctx->stream));
    ctx->launches += pr->launches_per_run;
  } else {
    ctx->launches += sel_enqueue(ctx, pr, ctx->stream, &nrc);
    if (nrc) return fail(ctx, BVIO_ERR_NCCL, std::string("nccl: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?"));
  }
  if (pr->sp.fused) {            // every rank runs the same sequence of selections: epochs stay in lock step
    ctx->sel_epoch += (unsigned long long)pr->sp.kappa;
    pr->sp.epoch_base = ctx->sel_epoch;
  }
  BVIO_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  BVIO_CUDA_OK(ctx, cudaGetLastError());
  return BVIO_OK;
}

int bvio_select_fetch(bvio_ctx* ctx, bvio_selprob* pr, int32_t* out_ids, double* out_values, bvio_select_summary* summary) {
  if (!ctx || !pr) return fail(ctx, BVIO_ERR_INVALID, "null problem");
  cudaSetDevice(ctx->device);
  char* ho = pr->slab.h + pr->out_off;
  BVIO_CUDA_OK(ctx, cudaMemcpyAsync(ho, pr->slab.d + pr->out_off, pr->out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BVIO_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  const int* idx = (const int*)(pr->slab.h + pr->o_out_idx);
  const double* val = (const double*)(pr->slab.h + pr->o_out_val);
  const SelCtrl* c = (const SelCtrl*)(pr->slab.h + pr->o_ctrl);
  for (int i = 0; i < c->n_selected; i++) {
    if (out_ids) out_ids[i] = pr->cand_id[idx[i]];
    if (out_values) out_values[i] = val[i];
  }
  if (c->peer_timeout) return fail(ctx, BVIO_ERR_NCCL, "fused selector: a peer's round record did not arrive within 3 s");
  if (summary) {
    summary->n_selected = c->n_selected;
    summary->n_candidates_valid = c->n_valid;
    summary->candidates_scored = (int64_t)c->scored;
    summary->final_logdet = c->final_logdet;
    summary->min_margin = c->min_margin;
    summary->device_ms = ms;
    const SelProb& sp = pr->sp;
    summary->transport = sp.world == 1 ? 0 : (sp.fused ? 1 : 2);
    summary->world = sp.world;
    summary->grid = sp.grid_persist > 0 ? sp.grid_persist : sp.grid_round;
    summary->cpw = sp.grid_persist > 0 ? sp.cpw : 0;
    const double rounds = sp.kappa > 0 ? (double)sp.kappa : 1.0;
    summary->round_score_us = sp.grid_persist > 0 ? c->t_score * 1e-3 / rounds : 0.0;
    summary->round_barrier_us = sp.grid_persist > 0 ? c->t_barrier * 1e-3 / rounds : 0.0;
    summary->round_exchange_us = sp.grid_persist > 0 ? c->t_exchange * 1e-3 / rounds : 0.0;
  }
  return BVIO_OK;
}

void bvio_select_free(bvio_ctx* ctx, bvio_selprob* pr) {
  if (!pr) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (pr->graph) cudaGraphExecDestroy(pr->graph);
  bool dummy = true;
  slab_release(pr->slab, ctx ? ctx->sel_cache_busy : dummy, pr->from_cache);
  delete pr;
}

static int select_impl(bvio_ctx* ctx, const bvio_select_in* in, bool sharded, int32_t* out_ids, double* out_values,
                       bvio_select_summary* summary) {
  bvio_selprob* pr = nullptr;
  int rc = sel_upload_impl(ctx, in, true, sharded, &pr);
  if (rc) return rc;
  pr->use_graph = false;
  rc = bvio_select_run(ctx, pr);
  if (rc == BVIO_OK) rc = bvio_select_fetch(ctx, pr, out_ids, out_values, summary);
  bvio_select_free(ctx, pr);
  return rc;
}

int bvio_select(bvio_ctx* ctx, const bvio_select_in* in, int32_t* out_ids, double* out_values, bvio_select_summary* summary) {
  return select_impl(ctx, in, false, out_ids, out_values, summary);
}

int bvio_select_sharded(bvio_ctx* ctx, const bvio_select_in* in, int32_t* out_ids, double* out_values,
                        bvio_select_summary* summary) {
  return select_impl(ctx, in, true, out_ids, out_values, summary);
}

int bvio_nccl_unique_id(void* uid128) {
  if (!uid128 || !nccl_load()) return BVIO_ERR_NCCL;
  nccl_uid_t id;
  if (g_nccl.GetUniqueId(&id)) return BVIO_ERR_NCCL;
  memcpy(uid128, &id, sizeof id);
  return BVIO_OK;
}

int bvio_comm_init(bvio_ctx* ctx, const void* uid128, int32_t rank, int32_t world) {
  if (!ctx || !uid128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, BVIO_ERR_INVALID, "bad comm arguments");
  if (!nccl_load()) return fail(ctx, BVIO_ERR_NCCL, "libnccl.so.2 not found");
  cudaSetDevice(ctx->device);
  if (ctx->comm) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
  nccl_uid_t id;
  memcpy(&id, uid128, sizeof id);
  int r = g_nccl.CommInitRank(&ctx->comm, world, id, rank);
  if (r) return fail(ctx, BVIO_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  ctx->rank = rank; ctx->world = world;
  mbox_setup(ctx);
  return BVIO_OK;
}

int bvio_debug_build_delta(bvio_ctx* ctx, const bvio_select_in* in, double* Cout, int32_t* valid, double* omega) {
  bvio_selprob* pr = nullptr;
  int rc = sel_upload_impl(ctx, in, false, false, &pr);
  if (rc) return rc;
  const SelProb& sp = pr->sp;
  ctx->launches += sel_launch_reset(sp, ctx->stream);
  ctx->launches += sel_launch_build(sp, ctx->stream);
  cudaError_t e = cudaGetLastError();
  double* dfull = nullptr;
  if (e == cudaSuccess && Cout && sp.N) {
    e = cudaMalloc((void**)&dfull, sizeof(double) * (size_t)sp.N * sp.T * sp.T);
    if (e == cudaSuccess) { ctx->launches += sel_launch_expand(sp, dfull, ctx->stream); e = cudaGetLastError(); }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && dfull) e = cudaMemcpy(Cout, dfull, sizeof(double) * (size_t)sp.N * sp.T * sp.T, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && valid && sp.N) e = cudaMemcpy(valid, sp.valid, sizeof(int) * sp.N, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && omega) e = cudaMemcpy(omega, sp.omega, sizeof(double) * (size_t)sp.D * sp.D, cudaMemcpyDeviceToHost);
  if (dfull) cudaFree(dfull);
  bvio_select_free(ctx, pr);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("debug_build_delta: ") + cudaGetErrorString(e));
  return BVIO_OK;
}

}  // extern "C"
