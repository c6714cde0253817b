// ba_api.cu -- C-ABI of the bundle-adjustment path (include/bvio.h): context, batch upload,
// solve, download.  Host code only packs the caller's arrays into one pinned slab, issues one
// H2D copy, launches the kernels of ba_kernels.cu and copies the result back; there is no CPU
// compute path (Estimator::optimization(), vins_estimator/src/estimator.cpp:661-814).
#include "ba.h"
#include "ctx.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <chrono>
#include <new>
#include <thread>

using namespace bvio;

// host threads for packing / validating / scattering large batches: at most 16, and this process's share of the cores
// when several ranks live on one node (LOCAL_WORLD_SIZE is what torchrun / mpirun wrappers export); BVIO_PACK_THREADS
// overrides.  (Round 1: 8 ranks x 16 packing threads on a 32-core host cost 20 % of the end-to-end scaling.)
static int pack_threads() {
  static int cached = 0;
  if (cached) return cached;
  int n = (int)std::max(1u, std::thread::hardware_concurrency());
  if (const char* lw = getenv("LOCAL_WORLD_SIZE")) { const int w = atoi(lw); if (w > 1) n = std::max(1, n / w); }
  n = std::min(n, 16);
  if (const char* ev = getenv("BVIO_PACK_THREADS")) n = std::max(1, std::min(atoi(ev), 64));
  cached = n;
  return n;
}

struct bvio_batch {
  BaBatch bt;
  Slab slab;
 bool from_cache = false;
  int cache_slot = -1;                 // -1: own allocation, 0: ctx->ba_cache, 1 + i: ctx->ba_pipe[i]
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // device time of the last solve of this batch
  size_t in_bytes = 0;                 // [0, in_bytes): inputs, mirrored on the host
  size_t out_off = 0, out_bytes = 0;   // [out_off, out_off+out_bytes): outputs, mirrored on the host
  size_t o_pose_out = 0, o_sb_out = 0, o_invd_out = 0, o_ex_out = 0, o_td_out = 0, o_ctrl = 0;
  std::vector<int> lm_base;
  std::vector<int> perm;          // [total_L] device landmark -> caller landmark (within its window)
  cudaGraphExec_t graph = nullptr;
  bool use_graph = true;
  int launches_per_solve = 0;
  int debug = 0;
  int Kc = 0;          // the caller's K; bt.K = Kc + 1 when relocalization factors ride along as one more frame
  bool relo = false;
};

// a marginalization in flight on the context's second stream (bvio_marginalize_begin / _end)
struct bvio_marg_job {
  bvio_batch* bb = nullptr;
  bvio_prior_out* out = nullptr;
  int n = 0;
  size_t o_jac = 0, o_res = 0;
};

extern "C" {

int bvio_abi_version(void) { return BVIO_ABI_VERSION; }

void bvio_default_opts(bvio_opts* o) {
  if (!o) return;
  memset(o, 0, sizeof *o);
  o->max_iters = 8;               // config/euroc/euroc_config.yaml:55
  o->max_time_s = 0.0;
  o->estimate_extrinsic = 0;      // euroc_config.yaml:25
  o->estimate_td = 0;             // euroc_config.yaml:72
  o->focal_length = 460.0;        // parameters.h:13
  o->cauchy_a = 1.0;              // estimator.cpp:666
  o->G[0] = 0; o->G[1] = 0; o->G[2] = 9.81007;   // euroc_config.yaml:63
  o->TR = 0.0; o->ROW = 480.0;
  o->function_tolerance = 1e-6;   // Ceres defaults (estimator.cpp:794-806 leaves them untouched)
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_radius = 1e4;
  o->min_relative_decrease = 1e-3;
  o->strategy = BVIO_STRATEGY_LM;
  o->jacobi_scaling = 1;
}

int bvio_create(int device, bvio_ctx** out) {
  if (!out) return BVIO_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return BVIO_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return BVIO_ERR_CUDA;
  bvio_ctx* c = new (std::nothrow) bvio_ctx();
  if (!c) return BVIO_ERR_INVALID;
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return BVIO_ERR_CUDA; }
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      ba_configure() != 0) {
    delete c;
    return BVIO_ERR_CUDA;
  }
  *out = c;
  return BVIO_OK;
}

void bvio_sel_ctx_destroy(bvio_ctx* ctx);   // sel_api.cu

void bvio_destroy(bvio_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  bvio::ba_ws_prof_dump();
  bvio_sel_ctx_destroy(ctx);
  ctx->ba_cache.release();
  for (int i = 0; i < bvio_ctx::PIPE; i++) ctx->ba_pipe[i].release();
  ctx->sel_cache.release();
  if (ctx->marg_scratch) cudaFree(ctx->marg_scratch);
  if (ctx->marg_host) cudaFreeHost(ctx->marg_host);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

const char* bvio_last_error(const bvio_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* bvio_stream(bvio_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t bvio_launch_count(const bvio_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// pure (thread-safe) structural validation of one window; *msg names the first problem
static int validate_msg(const bvio_window* w, const bvio_opts* o, int K0, const char** msg) {
  if (!w || !o) { *msg = "null window/opts"; return BVIO_ERR_INVALID; }
  if (o->estimate_td && (!w->obs_vel || !w->obs_td || !w->obs_row || !w->para_td))
    { *msg = "estimate_td needs obs_vel / obs_td / obs_row / para_td"; return BVIO_ERR_INVALID; }
  if (o->estimate_td && !(o->ROW > 0)) { *msg = "estimate_td needs ROW > 0"; return BVIO_ERR_INVALID; }
  if (15 * w->K + (o->estimate_extrinsic ? 6 : 0) + (o->estimate_td ? 1 : 0) > 226 ||
      w->K + (o->estimate_extrinsic ? 1 : 0) + (o->estimate_td ? 1 : 0) > BVIO_KMAX)
    { *msg = "K too large with estimate_extrinsic / estimate_td (reduced system must fit one CTA's shared memory)"; return BVIO_ERR_INVALID; }
  if (o->strategy != BVIO_STRATEGY_LM && o->strategy != BVIO_STRATEGY_DOGLEG)
    { *msg = "unknown trust-region strategy"; return BVIO_ERR_INVALID; }
  if (w->K < 2 || w->K > BVIO_KMAX - 1) { *msg = "K out of range [2,15] (reduced system must fit one CTA's shared memory)"; return BVIO_ERR_INVALID; }
  if (w->K != K0) { *msg = "all windows of a batch must have the same K"; return BVIO_ERR_INVALID; }
  if (w->L < 0 || !w->para_pose || !w->para_speed_bias || !w->para_ex_pose || !w->preint)
    { *msg = "null state / preint arrays"; return BVIO_ERR_INVALID; }
  if (w->L > 0 && (!w->inv_depth || !w->lm_obs_offset || !w->obs_frame || !w->obs_xy))
    { *msg = "null landmark arrays"; return BVIO_ERR_INVALID; }
  if (w->L > 0 && w->lm_obs_offset[0] != 0) { *msg = "lm_obs_offset[0] != 0"; return BVIO_ERR_INVALID; }
  for (int l = 0; l < w->L; l++) {
    int o0 = w->lm_obs_offset[l], o1 = w->lm_obs_offset[l + 1], n = o1 - o0;
    if (n < 2 || n > BVIO_KMAX) { *msg = "landmark needs 2..16 observations"; return BVIO_ERR_INVALID; }
    for (int k = o0; k < o1; k++) {
      int f = w->obs_frame[k];
      if (f < 0 || f >= w->K || (k > o0 && f <= w->obs_frame[k - 1]))
        { *msg = "obs_frame must be strictly ascending within [0,K)"; return BVIO_ERR_INVALID; }
    }
  }
  if (w->n_relo < 0 || (w->n_relo > 0 && (!w->relo_pose || !w->relo_lm || !w->relo_xy)))
    { *msg = "relocalization: n_relo < 0 or null relo_pose / relo_lm / relo_xy"; return BVIO_ERR_INVALID; }
  if (w->n_relo > 0) {
    if (o->estimate_td) { *msg = "relocalization factors with estimate_td are not supported"; return BVIO_ERR_UNSUPPORTED; }
    if (w->K + 1 + (o->estimate_extrinsic ? 1 : 0) > BVIO_KMAX || 15 * (w->K + 1) + (o->estimate_extrinsic ? 6 : 0) > 226)
      { *msg = "relocalization: K too large (the loop-closure pose rides as one more frame)"; return BVIO_ERR_UNSUPPORTED; }
    for (int k = 0; k < w->n_relo; k++) {
      int l = w->relo_lm[k];
      if (l < 0 || l >= w->L || (k > 0 && l <= w->relo_lm[k - 1]))
        { *msg = "relo_lm must be strictly ascending within [0,L)"; return BVIO_ERR_INVALID; }
      if (w->lm_obs_offset[l + 1] - w->lm_obs_offset[l] + 1 > BVIO_KMAX)
        { *msg = "landmark needs 2..16 observations (relocalization match included)"; return BVIO_ERR_INVALID; }
    }
  }
  if (w->prior) {
    const bvio_prior* p = w->prior;
    if (p->n < 0 || p->n > 256 || p->nblocks < 0 || p->nblocks > PRIOR_MAXB)
      { *msg = "prior dimension out of range"; return BVIO_ERR_INVALID; }
    if (p->nblocks > 0 && (!p->block_kind || !p->block_frame || !p->block_idx || !p->x0))
      { *msg = "null prior block arrays"; return BVIO_ERR_INVALID; }
    if (p->n > 0 && (!p->lin_jac || !p->lin_res)) { *msg = "null prior lin_jac / lin_res"; return BVIO_ERR_INVALID; }
    for (int b = 0; b < p->nblocks; b++) {
      int kind = p->block_kind[b], loc = kind == BVIO_BLK_SPEEDBIAS ? 9 : (kind == BVIO_BLK_TD ? 1 : 6);
      if (kind < 0 || kind > 3 || p->block_idx[b] < 0 || p->block_idx[b] + loc > p->n)
        { *msg = "prior block out of range"; return BVIO_ERR_INVALID; }
      if ((kind == BVIO_BLK_POSE || kind == BVIO_BLK_SPEEDBIAS) && (p->block_frame[b] < 0 || p->block_frame[b] >= w->K))
        { *msg = "prior block frame out of range"; return BVIO_ERR_INVALID; }
    }
  }
  return BVIO_OK;
}

static int validate(bvio_ctx* ctx, const bvio_window* w, const bvio_opts* o, int K0) {
  const char* msg = "";
  int rc = validate_msg(w, o, K0, &msg);
  return rc ? fail(ctx, rc, msg) : BVIO_OK;
}

static Slab& cache_of(bvio_ctx* ctx, int slot) { return slot == 0 ? ctx->ba_cache : ctx->ba_pipe[slot - 1]; }
static bool& busy_of(bvio_ctx* ctx, int slot) { return slot == 0 ? ctx->ba_cache_busy : ctx->ba_pipe_busy[slot - 1]; }
static void release_slab(bvio_ctx* ctx, bvio_batch* bb) {
  bool dummy = true;
  if (bb->cache_slot >= 0 && ctx) slab_release(bb->slab, busy_of(ctx, bb->cache_slot), bb->from_cache);
  else slab_release(bb->slab, dummy, false);
}

// cache_slot: -1 = private allocation, 0 = the one-shot cache, 1.. = pipeline slots
static int upload_impl(bvio_ctx* ctx, const bvio_window* ws, int B, const bvio_opts* o, int cache_slot, int debug,
                       bvio_batch** out) {
  if (!ctx || !ws || B < 1 || !o || !out) return fail(ctx, BVIO_ERR_INVALID, "bad arguments");
  *out = nullptr;
  cudaSetDevice(ctx->device);
  // Relocalization (estimator.cpp:760-792): relo_Pose rides as frame Kc of a (Kc+1)-frame window -- no IMU link to it
  // (its preintegration slot carries sum_dt > 10, the reference's own "skip this factor" rule, estimator.cpp:705) and a
  // zero speed-bias block whose columns are identically zero: they add nothing to the gradient, the step, the model
  // decrease or any norm Ceres' loop looks at, so the iteration is the one of the reference's problem.
  const int Kc = ws[0].K;
  bool relo = false;
  for (int b = 0; b < B; b++) relo |= ws[b].n_relo > 0;
  const int K = Kc + (relo ? 1 : 0);
  int total_L = 0, total_obs = 0, nmax = 1, maxL = 0;
  {
    // structural validation walks every observation: spread large batches over host threads
    const int nthreads = B < 16 ? 1 : pack_threads();
    std::vector<int> rcs(nthreads, BVIO_OK);
    std::vector<const char*> msgs(nthreads, "");
    auto work = [&](int t) {
      for (int b = t; b < B && rcs[t] == BVIO_OK; b += nthreads) rcs[t] = validate_msg(ws + b, o, Kc, &msgs[t]);
    };
    ctx->pool.run(nthreads, work);
    for (int t = 0; t < nthreads; t++) if (rcs[t]) return fail(ctx, rcs[t], msgs[t]);
  }
  for (int b = 0; b < B; b++) {
    total_L += ws[b].L;
    total_obs += (ws[b].L ? ws[b].lm_obs_offset[ws[b].L] : 0) + ws[b].n_relo;
    maxL = std::max(maxL, ws[b].L);
    if (ws[b].prior) nmax = std::max(nmax, ws[b].prior->n);
  }
  bvio_batch* bb = new (std::nothrow) bvio_batch();
  if (!bb) return fail(ctx, BVIO_ERR_INVALID, "out of host memory");
  BaBatch& bt = bb->bt;
  memset(&bt, 0, sizeof bt);
  const int est_ex = o->estimate_extrinsic != 0, est_td = o->estimate_td != 0, XB = est_ex + est_td, KE = K + XB;
  bt.B = B; bt.K = K; bt.np = 15 * K + (est_ex ? 6 : 0) + est_td; bt.total_L = total_L; bt.total_obs = total_obs; bt.nmax = nmax;
  bt.est_ex = est_ex; bt.est_td = est_td;
  bb->Kc = Kc; bb->relo = relo;
  bt.solve_wide = B < ctx->sm_count && !getenv("BVIO_SOLVE_NARROW");
  bt.use_mma = XB == 0 && !getenv("BVIO_LEGACY_LINEARIZE");
  bt.use_ws = bt.use_mma && (getenv("BVIO_LIN_WS") ? atoi(getenv("BVIO_LIN_WS")) != 0 : !bt.solve_wide);
  bt.chunk_l = ba_pick_chunk(K, XB);
  int T = (maxL + 2 * bt.chunk_l - 1) / (2 * bt.chunk_l);   // >= 2 chunks per tile when there is a choice
  int Tcap = std::max(1, (10 * ctx->sm_count + B - 1) / B);   // ~5 waves of 2 CTAs/SM: measured optimum (tile record traffic vs tail)
  T = std::max(1, std::min(std::min(T, Tcap), 32));
  if (const char* ev = getenv("BVIO_TILES")) T = std::max(1, std::min(atoi(ev), 32));   // tuning knob
  bt.T = T;
  // the warp-specialised linearization keeps one CTA per SM busy for a whole tile: few, long tiles (about four waves) --
  // the pipeline fill / drain of a CTA and its tile record (HBM, re-read by ba_solve) are paid once per tile
  bt.TL = T;
  if (bt.use_ws) bt.TL = std::max(1, std::min(T, (4 * ctx->sm_count + B - 1) / B));
  if (const char* ev = getenv("BVIO_LIN_TILES")) bt.TL = std::max(1, std::min(atoi(ev), T));
  bt.undamped = debug;
  bt.max_iters = o->max_iters; bt.jacobi_scaling = o->jacobi_scaling;
  bt.strategy = o->strategy;
  bt.sqrt_info = o->focal_length / 1.5;   // estimator.cpp:17
  bt.cauchy_a = o->cauchy_a;
  for (int i = 0; i < 3; i++) bt.G[i] = o->G[i];
  bt.function_tolerance = o->function_tolerance; bt.gradient_tolerance = o->gradient_tolerance;
  bt.parameter_tolerance = o->parameter_tolerance; bt.initial_radius = o->initial_radius;
  bt.min_relative_decrease = o->min_relative_decrease;
  bt.max_time_s = o->max_time_s > 0.0 ? o->max_time_s : 0.0;
  bb->debug = debug;

  // ---- carve: inputs | outputs | scratch
  Carver cv;
  const size_t D = sizeof(double), I = sizeof(int);
  size_t o_lm_base = cv.take((B + 1) * I), o_lm_off = cv.take((total_L + 1) * I), o_obs_frame = cv.take(total_obs * I);
  size_t o_obs_xy = cv.take((size_t)total_obs * 2 * D);
  size_t o_pose0 = cv.take((size_t)B * K * 7 * D), o_sb0 = cv.take((size_t)B * K * 9 * D), o_ex = cv.take((size_t)B * 7 * D);
  size_t o_invd0 = cv.take((size_t)total_L * D);
  size_t o_td0 = cv.take((size_t)B * D);
  size_t o_obs_vel = est_td ? cv.take((size_t)total_obs * 2 * D) : 0, o_obs_shift = est_td ? cv.take((size_t)total_obs * D) : 0;
  size_t o_preint = cv.take((size_t)B * K * PREINT_DOUBLES * D);
  size_t o_pr_n = cv.take(B * I), o_pr_nb = cv.take(B * I);
  size_t o_pr_kind = cv.take((size_t)B * PRIOR_MAXB * I), o_pr_frame = cv.take((size_t)B * PRIOR_MAXB * I);
  size_t o_pr_idx = cv.take((size_t)B * PRIOR_MAXB * I);
  size_t o_pr_x0 = cv.take((size_t)B * PRIOR_MAXB * 9 * D);
  size_t o_pr_jac = cv.take((size_t)B * nmax * nmax * D), o_pr_res = cv.take((size_t)B * nmax * D);
  bb->in_bytes = cv.off;
  bb->out_off = cv.off;
  bb->o_pose_out = cv.take((size_t)B * K * 7 * D);
  bb->o_sb_out = cv.take((size_t)B * K * 9 * D);
  bb->o_invd_out = cv.take((size_t)total_L * D);
  bb->o_ex_out = cv.take((size_t)B * 7 * D);
  bb->o_td_out = cv.take((size_t)B * D);
  bb->o_ctrl = cv.take((size_t)B * sizeof(BaCtrl));
  bb->out_bytes = cv.off - bb->out_off;
  const size_t h_bytes = cv.off;
  size_t o_pose[2], o_sb[2], o_invd[2];
  for (int k = 0; k < 2; k++) {
    o_pose[k] = cv.take((size_t)B * K * 7 * D); o_sb[k] = cv.take((size_t)B * K * 9 * D);
    o_invd[k] = cv.take((size_t)total_L * D);
  }
  size_t o_imu = cv.take((size_t)B * K * IMU_REC * D), o_imu_out = cv.take((size_t)B * K * IMU_OUT * D);
  size_t o_pr_H = cv.take((size_t)B * nmax * nmax * D), o_pr_map = cv.take((size_t)B * nmax * I);
  size_t o_pr_out = cv.take((size_t)B * (nmax + 1) * D);
  size_t o_pr_inv = cv.take((size_t)B * bt.np * I);
  const size_t s0_len = (size_t)(bt.np + 1) * (bt.np + 2) / 2;
  size_t o_S0 = bt.solve_wide ? cv.take((size_t)B * s0_len * D) : 0;
  size_t o_h = cv.take((size_t)total_L * D), o_b = cv.take((size_t)total_L * D), o_sl2 = cv.take((size_t)total_L * D);
  size_t o_w = cv.take((size_t)total_obs * 6 * D);
  size_t o_tile = cv.take((size_t)B * T * tile_rec_doubles(KE) * D);
  size_t o_exs[2] = {cv.take((size_t)B * 7 * D), cv.take((size_t)B * 7 * D)};
  size_t o_wex = XB ? cv.take((size_t)total_L * XB * 6 * D) : 0;
  size_t o_tds[2] = {cv.take((size_t)B * D), cv.take((size_t)B * D)};
  size_t o_cost = cv.take((size_t)B * (T + 1) * COST_REC * D);
  size_t o_dp = cv.take((size_t)B * bt.np * D), o_sp = cv.take((size_t)B * bt.np * D);
  size_t o_svec = cv.take((size_t)B * 5 * bt.np * D);
  size_t o_dog_t = 0, o_dog_l = 0, o_dog_out = 0;
  if (bt.strategy == BVIO_STRATEGY_DOGLEG) {
    o_dog_t = cv.take((size_t)B * bt.np * D); o_dog_l = cv.take((size_t)total_L * 2 * D);
    o_dog_out = cv.take((size_t)B * T * DOG_REC * D);
  }
  size_t o_dbgS = 0, o_dbgg = 0;
  if (debug) { o_dbgS = cv.take((size_t)B * bt.np * bt.np * D); o_dbgg = cv.take((size_t)B * bt.np * D); }
  const size_t d_bytes = cv.off;

  cudaError_t ce;
  bb->cache_slot = cache_slot;
  if (cache_slot >= 0) ce = slab_acquire(cache_of(ctx, cache_slot), busy_of(ctx, cache_slot), d_bytes, h_bytes, bb->slab, bb->from_cache);
  else {
    bool dummy_busy = true;
    Slab none;
    ce = slab_acquire(none, dummy_busy, d_bytes, h_bytes, bb->slab, bb->from_cache);
  }
  if (ce == cudaSuccess) ce = cudaEventCreate(&bb->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&bb->ev1);
  if (ce != cudaSuccess) { delete bb; return fail(ctx, BVIO_ERR_CUDA, std::string("slab alloc: ") + cudaGetErrorString(ce)); }
  char* d = bb->slab.d;
  char* h = bb->slab.h;

  // ---- pack the host mirror
  int* h_lm_base = (int*)(h + o_lm_base);
  int* h_lm_off = (int*)(h + o_lm_off);
  int* h_obs_frame = (int*)(h + o_obs_frame);
  double* h_obs_xy = (double*)(h + o_obs_xy);
  double* h_pose0 = (double*)(h + o_pose0);
  double* h_sb0 = (double*)(h + o_sb0);
  double* h_ex = (double*)(h + o_ex);
  double* h_invd0 = (double*)(h + o_invd0);
  double* h_td0 = (double*)(h + o_td0);
  double* h_obs_vel = (double*)(h + o_obs_vel);
  double* h_obs_shift = (double*)(h + o_obs_shift);
  double* h_pre = (double*)(h + o_preint);
  int* h_pr_n = (int*)(h + o_pr_n);
  int* h_pr_nb = (int*)(h + o_pr_nb);
  int* h_pr_kind = (int*)(h + o_pr_kind);
  int* h_pr_frame = (int*)(h + o_pr_frame);
  int* h_pr_idx = (int*)(h + o_pr_idx);
  double* h_pr_x0 = (double*)(h + o_pr_x0);
  double* h_pr_jac = (double*)(h + o_pr_jac);
  double* h_pr_res = (double*)(h + o_pr_res);
  bb->lm_base.resize(B + 1);
  bb->perm.resize(total_L);
  std::vector<int> obs_base(B + 1);
  {
    int lb0 = 0, ob0 = 0;
    for (int b = 0; b < B; b++) {
      bb->lm_base[b] = lb0; obs_base[b] = ob0;
      lb0 += ws[b].L; ob0 += (ws[b].L ? ws[b].lm_obs_offset[ws[b].L] : 0) + ws[b].n_relo;
    }
    bb->lm_base[B] = lb0; obs_base[B] = ob0;
  }
  // windows are packed independently: spread large batches over a few host threads
  auto pack = [&](int b) {
    const bvio_window& w = ws[b];
    const int lb = bb->lm_base[b], ob = obs_base[b];
    h_lm_base[b] = lb;
    // device order: landmarks grouped by anchor frame (stable counting sort).  FeatureManager's list is
    // already in this order (features are appended as they are first seen); ba_linearize_mma walks the
    // anchors present in a chunk, so grouped input keeps that loop at 1-2 trips.  perm[device] = caller index.
    int* perm = bb->perm.data() + lb;
    bool grouped = true;                                   // already in FeatureManager's order?
    {
      int cnt[BVIO_KMAX + 1] = {0};
      int prev = -1;
      for (int l = 0; l < w.L; l++) {
        const int a = w.obs_frame[w.lm_obs_offset[l]];
        cnt[a + 1]++;
        grouped &= a >= prev;
        prev = a;
      }
      for (int f = 0; f < BVIO_KMAX; f++) cnt[f + 1] += cnt[f];
      if (grouped) for (int l = 0; l < w.L; l++) perm[l] = l;
      else for (int l = 0; l < w.L; l++) perm[cnt[w.obs_frame[w.lm_obs_offset[l]]]++] = l;
    }
    if (grouped && w.n_relo == 0 && w.L > 0) {
      // the caller's arrays are already the device layout: bulk copies instead of one small copy per landmark
      const int nobs = w.lm_obs_offset[w.L], o_first = w.lm_obs_offset[0];
      for (int j = 0; j < w.L; j++) h_lm_off[lb + j] = ob + (w.lm_obs_offset[j] - o_first);
      memcpy(h_obs_frame + ob, w.obs_frame + o_first, (size_t)(nobs - o_first) * I);
      memcpy(h_obs_xy + (size_t)2 * ob, w.obs_xy + (size_t)2 * o_first, (size_t)(nobs - o_first) * 2 * D);
      if (est_td) {
        memcpy(h_obs_vel + (size_t)2 * ob, w.obs_vel + (size_t)2 * o_first, (size_t)(nobs - o_first) * 2 * D);
        for (int k = o_first; k < nobs; k++) h_obs_shift[ob + k - o_first] = -w.obs_td[k] + o->TR / o->ROW * (w.obs_row[k] - o->ROW / 2);
      }
    }
    int run = ob;
    std::vector<int> relo_of;
    if (w.n_relo > 0) { relo_of.assign(w.L, -1); for (int k = 0; k < w.n_relo; k++) relo_of[w.relo_lm[k]] = k; }
    for (int j = 0; j < w.L && !(grouped && w.n_relo == 0); j++) {
      const int l = perm[j], o0 = w.lm_obs_offset[l], n = w.lm_obs_offset[l + 1] - o0;
      h_lm_off[lb + j] = run;
      memcpy(h_obs_frame + run, w.obs_frame + o0, n * I);
      memcpy(h_obs_xy + (size_t)2 * run, w.obs_xy + (size_t)2 * o0, (size_t)n * 2 * D);
      if (w.n_relo > 0 && relo_of[l] >= 0) {             // the match in the loop-closure frame = one more observation
        h_obs_frame[run + n] = Kc;
        h_obs_xy[2 * (size_t)(run + n)] = w.relo_xy[2 * relo_of[l]];
        h_obs_xy[2 * (size_t)(run + n) + 1] = w.relo_xy[2 * relo_of[l] + 1];
        run += 1;
      }
      if (est_td) {
        memcpy(h_obs_vel + (size_t)2 * run, w.obs_vel + (size_t)2 * o0, (size_t)n * 2 * D);
        // row centred like the factor's constructor (projection_td_factor.cpp:18-19)
        for (int k = 0; k < n; k++) h_obs_shift[run + k] = -w.obs_td[o0 + k] + o->TR / o->ROW * (w.obs_row[o0 + k] - o->ROW / 2);
      }
      run += n;
    }
    memcpy(h_pose0 + (size_t)b * K * 7, w.para_pose, (size_t)Kc * 7 * D);
    memcpy(h_sb0 + (size_t)b * K * 9, w.para_speed_bias, (size_t)Kc * 9 * D);
    if (relo) {
      static const double ident[7] = {0, 0, 0, 0, 0, 0, 1};
      memcpy(h_pose0 + ((size_t)b * K + Kc) * 7, w.n_relo > 0 ? w.relo_pose : ident, 7 * D);
      memset(h_sb0 + ((size_t)b * K + Kc) * 9, 0, 9 * D);
    }
    memcpy(h_ex + (size_t)b * 7, w.para_ex_pose, 7 * D);
    h_td0[b] = w.para_td ? w.para_td[0] : 0.0;
    if (grouped) memcpy(h_invd0 + lb, w.inv_depth, (size_t)w.L * D);
    else for (int j = 0; j < w.L; j++) h_invd0[lb + j] = w.inv_depth[perm[j]];
    memcpy(h_pre + (size_t)b * K * PREINT_DOUBLES, w.preint, (size_t)Kc * sizeof(bvio_preint));
    if (relo) {
      double* pr = h_pre + ((size_t)b * K + Kc) * PREINT_DOUBLES;
      memset(pr, 0, sizeof(bvio_preint));
      pr[6] = 1.0;                                         // delta_q = identity
      pr[16] = 1e9;                                        // sum_dt > 10: no IMU factor into the loop-closure frame
      for (int i = 0; i < 15; i++) pr[17 + 16 * i] = pr[242 + 16 * i] = 1.0;   // jacobian = covariance = I: ba_prepare's
                                                           // square-root information stays finite (it is never used)
    }
    h_pr_n[b] = 0; h_pr_nb[b] = 0;
    if (w.prior) {
      const bvio_prior* p = w.prior;
      h_pr_n[b] = p->n; h_pr_nb[b] = p->nblocks;
      const double* x0 = p->x0;
      for (int k = 0; k < p->nblocks; k++) {
        int kind = p->block_kind[k], gs = kind == BVIO_BLK_SPEEDBIAS ? 9 : (kind == BVIO_BLK_TD ? 1 : 7);
        h_pr_kind[b * PRIOR_MAXB + k] = kind;
        h_pr_frame[b * PRIOR_MAXB + k] = p->block_frame[k];
        h_pr_idx[b * PRIOR_MAXB + k] = p->block_idx[k];
        double* dst = h_pr_x0 + (size_t)(b * PRIOR_MAXB + k) * 9;
        for (int i = 0; i < 9; i++) dst[i] = i < gs ? x0[i] : 0.0;
        x0 += gs;
      }
      memcpy(h_pr_jac + (size_t)b * nmax * nmax, p->lin_jac, (size_t)p->n * p->n * D);
      memcpy(h_pr_res + (size_t)b * nmax, p->lin_res, (size_t)p->n * D);
    }
  };
  h_lm_base[B] = bb->lm_base[B];
  h_lm_off[total_L] = obs_base[B];

  bt.lm_base = (const int*)(d + o_lm_base); bt.lm_off = (const int*)(d + o_lm_off);
  bt.obs_frame = (const int*)(d + o_obs_frame); bt.obs_xy = (const double2*)(d + o_obs_xy);
  bt.pose0 = (const double*)(d + o_pose0); bt.sb0 = (const double*)(d + o_sb0); bt.ex = (const double*)(d + o_ex);
  bt.exs[0] = (double*)(d + o_exs[0]); bt.exs[1] = (double*)(d + o_exs[1]); bt.ex_out = (double*)(d + bb->o_ex_out);
  bt.wex = (double*)(d + o_wex);
  bt.td0 = (const double*)(d + o_td0); bt.tds[0] = (double*)(d + o_tds[0]); bt.tds[1] = (double*)(d + o_tds[1]);
  bt.td_out = (double*)(d + bb->o_td_out);
  bt.obs_vel = (const double2*)(d + o_obs_vel); bt.obs_shift = (const double*)(d + o_obs_shift);
  bt.invd0 = (const double*)(d + o_invd0); bt.preint_raw = (const double*)(d + o_preint);
  bt.pr_n = (const int*)(d + o_pr_n); bt.pr_nb = (const int*)(d + o_pr_nb);
  bt.pr_kind = (const int*)(d + o_pr_kind); bt.pr_frame = (const int*)(d + o_pr_frame); bt.pr_idx = (const int*)(d + o_pr_idx);
  bt.pr_x0 = (const double*)(d + o_pr_x0); bt.pr_jac = (const double*)(d + o_pr_jac); bt.pr_res = (const double*)(d + o_pr_res);
  bt.pose_out = (double*)(d + bb->o_pose_out); bt.sb_out = (double*)(d + bb->o_sb_out);
  bt.invd_out = (double*)(d + bb->o_invd_out); bt.ctrl = (BaCtrl*)(d + bb->o_ctrl);
  for (int k = 0; k < 2; k++) {
    bt.pose[k] = (double*)(d + o_pose[k]); bt.sb[k] = (double*)(d + o_sb[k]); bt.invd[k] = (double*)(d + o_invd[k]);
  }
  bt.imu = (double*)(d + o_imu); bt.imu_out = (double*)(d + o_imu_out);
  bt.pr_H = (double*)(d + o_pr_H); bt.pr_map = (int*)(d + o_pr_map); bt.pr_out = (double*)(d + o_pr_out);
  bt.pr_inv = (int*)(d + o_pr_inv); bt.S0 = bt.solve_wide ? (double*)(d + o_S0) : nullptr;
  bt.h = (double*)(d + o_h); bt.b = (double*)(d + o_b); bt.sl2 = (double*)(d + o_sl2); bt.w = (double*)(d + o_w);
  bt.tile_out = (double*)(d + o_tile); bt.cost_out = (double*)(d + o_cost);
  bt.delta_p = (double*)(d + o_dp); bt.scale_p = (double*)(d + o_sp); bt.solve_vec = (double*)(d + o_svec);
  bt.dog_t = (double*)(d + o_dog_t); bt.dog_l = (double*)(d + o_dog_l); bt.dog_out = (double*)(d + o_dog_out);
  bt.dbg_S = debug ? (double*)(d + o_dbgS) : nullptr;
  bt.dbg_g = debug ? (double*)(d + o_dbgg) : nullptr;

  // pipeline slots upload on the copy stream so that the H2D overlaps the previous sub-batch's kernels
  cudaStream_t up = cache_slot >= 1 ? ctx->copy_stream : ctx->stream;
  {
    // (packing and copying in four chunks of windows, so that the H2D of one chunk overlaps the packing of the next, was
    // measured: no gain -- the host side of a sub-batch is 2.9 ms, bound by host memory bandwidth, BVIO_DEBUG=1 prints it)
    int nthreads = pack_threads();
    if (B < 16) nthreads = 1;
    ctx->pool.run(nthreads, [&](int t) { for (int b = t; b < B; b += nthreads) pack(b); });
  }
  cudaError_t e = cudaMemcpyAsync(d, h, bb->in_bytes, cudaMemcpyHostToDevice, up);
  if (e == cudaSuccess) { ctx->launches += ba_launch_prepare(bt, up); e = cudaGetLastError(); }
  if (e == cudaSuccess && up != ctx->stream) {
    e = cudaEventRecord(bb->ev1, up);                      // ev1 is re-recorded by the solve afterwards
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, bb->ev1, 0);
  }
  if (e != cudaSuccess) {
    release_slab(ctx, bb);
    delete bb;
    return fail(ctx, BVIO_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e));
  }
  *out = bb;
  return BVIO_OK;
}

static int enqueue_solve(bvio_ctx* ctx, bvio_batch* bb, cudaStream_t st) {
  int n = 0;
  n += ba_launch_reset(bb->bt, st);
  for (int it = 0; it < bb->bt.max_iters; it++) n += ba_launch_iteration(bb->bt, st, true);
  n += ba_launch_iteration(bb->bt, st, false);   // final linearization: gradient norm of the result
  n += ba_launch_finish(bb->bt, st);
  return n;
}

extern "C" {

int bvio_batch_upload(bvio_ctx* ctx, const bvio_window* windows, int32_t B, const bvio_opts* opts, bvio_batch** out) {
  return upload_impl(ctx, windows, B, opts, -1, 0, out);
}

int bvio_batch_solve(bvio_ctx* ctx, bvio_batch* bb) {
  if (!ctx || !bb) return fail(ctx, BVIO_ERR_INVALID, "null batch");
  cudaSetDevice(ctx->device);
  BVIO_CUDA_OK(ctx, cudaEventRecord(bb->ev0, ctx->stream));
  if (bb->use_graph) {
    if (!bb->graph) {
      cudaGraph_t g = nullptr;
      BVIO_CUDA_OK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      bb->launches_per_solve = enqueue_solve(ctx, bb, ctx->stream);
      // always leave capture mode, whatever happened inside, and never leak the graph
      cudaError_t ec = cudaStreamEndCapture(ctx->stream, &g);
      if (ec == cudaSuccess) ec = cudaGraphInstantiate(&bb->graph, g, 0);
      if (g) cudaGraphDestroy(g);
      if (ec != cudaSuccess) { bb->graph = nullptr; BVIO_CUDA_OK(ctx, ec); }
      // the event recorded before the capture is still valid; re-record so ev0 directly precedes the launch
      BVIO_CUDA_OK(ctx, cudaEventRecord(bb->ev0, ctx->stream));
    }
    BVIO_CUDA_OK(ctx, cudaGraphLaunch(bb->graph, ctx->stream));
    ctx->launches += bb->launches_per_solve;
  } else {
    ctx->launches += enqueue_solve(ctx, bb, ctx->stream);
  }
  BVIO_CUDA_OK(ctx, cudaEventRecord(bb->ev1, ctx->stream));
  BVIO_CUDA_OK(ctx, cudaGetLastError());
  return BVIO_OK;
}

}  // extern "C"

// D2H of the outputs, asynchronous on the context stream
static int enqueue_d2h(bvio_ctx* ctx, bvio_batch* bb) {
  char* ho = bb->slab.h + bb->out_off;
  BVIO_CUDA_OK(ctx, cudaMemcpyAsync(ho, bb->slab.d + bb->out_off, bb->out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return BVIO_OK;
}

// after the stream has been synchronised: scatter the pinned output mirror into the caller's arrays
static int unpack_outputs(bvio_ctx* ctx, bvio_batch* bb, bvio_window* windows, bvio_summary* summaries) {
  float ms = 0;
  cudaEventElapsedTime(&ms, bb->ev0, bb->ev1);
  const BaBatch& bt = bb->bt;
  const double* pose = (const double*)(bb->slab.h + bb->o_pose_out);
  const double* sb = (const double*)(bb->slab.h + bb->o_sb_out);
  const double* invd = (const double*)(bb->slab.h + bb->o_invd_out);
  const double* exo = (const double*)(bb->slab.h + bb->o_ex_out);
  const double* tdo = (const double*)(bb->slab.h + bb->o_td_out);
  const BaCtrl* ctrl = (const BaCtrl*)(bb->slab.h + bb->o_ctrl);
  if (bt.B == 1 && getenv("BVIO_DEBUG")) {
    const unsigned long long* t = ctrl[0].stamps;
    fprintf(stderr, "[bvio] last ba_solve pass, us: assemble %.1f vectors+dogleg-pre %.1f cholesky %.1f back-subst %.1f step %.1f | solve device total %.3f ms\n",
            (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (t[3] - t[2]) * 1e-3, (t[4] - t[3]) * 1e-3, (t[5] - t[4]) * 1e-3, ms);
  }
  int rc = BVIO_OK;
  auto scatter = [&](int b) {
      bvio_window& w = windows[b];
      memcpy(w.para_pose, pose + (size_t)b * bt.K * 7, (size_t)bb->Kc * 7 * sizeof(double));
      memcpy(w.para_speed_bias, sb + (size_t)b * bt.K * 9, (size_t)bb->Kc * 9 * sizeof(double));
      if (bb->relo && w.n_relo > 0) memcpy(w.relo_pose, pose + ((size_t)b * bt.K + bb->Kc) * 7, 7 * sizeof(double));
      int L = bb->lm_base[b + 1] - bb->lm_base[b];
      const int* perm = bb->perm.data() + bb->lm_base[b];
      for (int j = 0; j < L; j++) w.inv_depth[perm[j]] = invd[bb->lm_base[b] + j];
      if (bt.est_ex) memcpy(w.para_ex_pose, exo + (size_t)b * 7, 7 * sizeof(double));
      if (bt.est_td) w.para_td[0] = tdo[b];
  };
  if (windows && bt.B >= 64) {
    const int nthreads = pack_threads();
    ctx->pool.run(nthreads, [&](int t) { for (int b = t; b < bt.B; b += nthreads) scatter(b); });
  }
  for (int b = 0; b < bt.B; b++) {
    if (windows && bt.B < 64) scatter(b);
    const BaCtrl& c = ctrl[b];
    if (summaries) {
      bvio_summary& s = summaries[b];
      s.iterations = c.iterations; s.num_accepted = c.accepted; s.num_rejected = c.rejected;
      s.termination = c.termination; s.initial_cost = c.initial_cost; s.final_cost = c.cost;
      s.final_radius = c.radius; s.final_gradient_max = c.gmax; s.device_ms = ms;
    }
    if (!(c.cost == c.cost) || c.cost > 1.79e308) rc = BVIO_ERR_NUMERIC;
  }
  if (rc) return fail(ctx, rc, "non-finite cost");
  return BVIO_OK;
}

extern "C" {

int bvio_batch_download(bvio_ctx* ctx, bvio_batch* bb, bvio_window* windows, bvio_summary* summaries) {
  if (!ctx || !bb) return fail(ctx, BVIO_ERR_INVALID, "null batch");
  cudaSetDevice(ctx->device);
  int rc = enqueue_d2h(ctx, bb);
  if (rc) return rc;
  BVIO_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return unpack_outputs(ctx, bb, windows, summaries);
}

void bvio_batch_free(bvio_ctx* ctx, bvio_batch* bb) {
  if (!bb) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (bb->graph) cudaGraphExecDestroy(bb->graph);
  if (bb->ev0) cudaEventDestroy(bb->ev0);
  if (bb->ev1) cudaEventDestroy(bb->ev1);
  release_slab(ctx, bb);
  delete bb;
}

// One-shot solve of B independent windows with host buffers.  Large batches are cut into sub-batches of about
// two windows per SM and software-pipelined on the context stream: while the GPU solves sub-batch i the host
// packs sub-batch i+1 into its own pinned slab, and the D2H of i trails its solve -- so the wall time tends to
// max(host packing, device solve) instead of their sum.
int bvio_optimize_batch(bvio_ctx* ctx, bvio_window* windows, int32_t B, const bvio_opts* opts, bvio_summary* summaries) {
  if (!ctx || !windows || B < 1 || !opts) return fail(ctx, BVIO_ERR_INVALID, "bad arguments");
  int per = std::max(1, 2 * ctx->sm_count);   // two windows per SM per sub-batch: measured optimum (74 ... 592 swept)
  if (const char* ev = getenv("BVIO_PIPE_WINDOWS")) per = std::max(1, atoi(ev));   // tuning knob: windows per sub-batch
  int S = std::min(bvio_ctx::PIPE, std::max(1, B / per));
  if (S == 1) {
    bvio_batch* bb = nullptr;
    int rc = upload_impl(ctx, windows, B, opts, 0, 0, &bb);
    if (rc) return rc;
    bb->use_graph = false;   // one-shot: direct launches beat capture + instantiate
    rc = bvio_batch_solve(ctx, bb);
    if (rc == BVIO_OK) rc = bvio_batch_download(ctx, bb, windows, summaries);
    bvio_batch_free(ctx, bb);
    return rc;
  }
  bvio_batch* sub[bvio_ctx::PIPE] = {nullptr};
  int b0[bvio_ctx::PIPE + 1];
  for (int i = 0; i <= S; i++) b0[i] = (int)((long long)B * i / S);
  int rc = BVIO_OK;
  const bool dbg = getenv("BVIO_DEBUG") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  double t_up[bvio_ctx::PIPE] = {0}, t_enq[bvio_ctx::PIPE] = {0};
  for (int i = 0; i < S && rc == BVIO_OK; i++) {
    const double ta = now();
    rc = upload_impl(ctx, windows + b0[i], b0[i + 1] - b0[i], opts, 1 + i, 0, &sub[i]);
    if (rc) break;
    t_up[i] = now() - ta;
    sub[i]->use_graph = false;
    rc = bvio_batch_solve(ctx, sub[i]);
    if (rc == BVIO_OK) rc = enqueue_d2h(ctx, sub[i]);
    t_enq[i] = now() - ta - t_up[i];
  }
  const double t1 = now();
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  const double t2 = now();
  if (rc == BVIO_OK && e != cudaSuccess) rc = fail(ctx, BVIO_ERR_CUDA, std::string("optimize_batch: ") + cudaGetErrorString(e));
  for (int i = 0; i < S; i++) {
    if (!sub[i]) continue;
    if (rc == BVIO_OK) rc = unpack_outputs(ctx, sub[i], windows + b0[i], summaries ? summaries + b0[i] : nullptr);
    bvio_batch_free(ctx, sub[i]);
  }
  if (dbg) {
    fprintf(stderr, "[bvio] optimize_batch B=%d in %d sub-batches: total %.2f ms = host until last enqueue %.2f (", B, S, now() - t0, t1 - t0);
    for (int i = 0; i < S; i++) fprintf(stderr, "upload %.2f + enqueue %.2f%s", t_up[i], t_enq[i], i + 1 < S ? ", " : "");
    fprintf(stderr, ") + wait for the GPU %.2f + unpack %.2f\n", t2 - t1, now() - t2);
  }
  return rc;
}

int bvio_optimize(bvio_ctx* ctx, bvio_window* window, const bvio_opts* opts, bvio_summary* summary) {
  return bvio_optimize_batch(ctx, window, 1, opts, summary);
}

int bvio_debug_linearize(bvio_ctx* ctx, const bvio_window* window, const bvio_opts* opts, double* S, double* g,
                         double* h, double* b, double* cost) {
  bvio_batch* bb = nullptr;
  int rc = upload_impl(ctx, window, 1, opts, -1, 1, &bb);
  if (rc) return rc;
  const BaBatch& bt = bb->bt;
  ctx->launches += ba_launch_reset(bt, ctx->stream);
  ctx->launches += ba_launch_iteration(bt, ctx->stream, true);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && S) e = cudaMemcpy(S, bt.dbg_S, sizeof(double) * bt.np * bt.np, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && g) e = cudaMemcpy(g, bt.dbg_g, sizeof(double) * bt.np, cudaMemcpyDeviceToHost);
  std::vector<double> tmp(bt.total_L);
  if (e == cudaSuccess && h && bt.total_L) {
    e = cudaMemcpy(tmp.data(), bt.h, sizeof(double) * bt.total_L, cudaMemcpyDeviceToHost);
    for (int j = 0; j < bt.total_L; j++) h[bb->perm[j]] = tmp[j];
  }
  if (e == cudaSuccess && b && bt.total_L) {
    e = cudaMemcpy(tmp.data(), bt.b, sizeof(double) * bt.total_L, cudaMemcpyDeviceToHost);
    for (int j = 0; j < bt.total_L; j++) b[bb->perm[j]] = tmp[j];
  }
  if (e == cudaSuccess && cost) {
    BaCtrl c;
    e = cudaMemcpy(&c, bt.ctrl, sizeof c, cudaMemcpyDeviceToHost);
    *cost = c.cost;
  }
  bvio_batch_free(ctx, bb);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("debug_linearize: ") + cudaGetErrorString(e));
  return BVIO_OK;
}

// Omega_PRIOR for bvio_select_in.omega_prior: the window's information on (position, velocity, accelerometer bias) of
// its newest frame = Schur complement of the undamped reduced matrix onto those nine dimensions (ba_marginal9_kernel).
int bvio_window_omega_prior(bvio_ctx* ctx, const bvio_window* window, const bvio_opts* opts, double* omega9) {
  if (!omega9) return fail(ctx, BVIO_ERR_INVALID, "null omega9");
  bvio_batch* bb = nullptr;
  int rc = upload_impl(ctx, window, 1, opts, -1, 1, &bb);
  if (rc) return rc;
  const BaBatch& bt = bb->bt;
  ctx->launches += ba_launch_reset(bt, ctx->stream);
  ctx->launches += ba_launch_iteration(bt, ctx->stream, true);
  double* dout = nullptr;
  cudaError_t e = cudaMalloc((void**)&dout, sizeof(double) * 81 + sizeof(int) * 2);
  int bad = 0;
  if (e == cudaSuccess) {
    ctx->launches += ba_launch_marginal9(bt.dbg_S, bt.np, bb->Kc - 1, dout, (int*)(dout + 81), ctx->stream);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy(omega9, dout, sizeof(double) * 81, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&bad, dout + 81, sizeof(int), cudaMemcpyDeviceToHost);
  if (dout) cudaFree(dout);
  bvio_batch_free(ctx, bb);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("window_omega_prior: ") + cudaGetErrorString(e));
  if (bad) return fail(ctx, BVIO_ERR_NUMERIC, "window_omega_prior: the window does not determine the other states (singular elimination)");
  return BVIO_OK;
}

// Per-kernel device time of one solve (bench.py roofline): direct launches with an event between
// kernels.  out_ms = {linearize, solve, cost, total of the three}, summed over all passes;
// out_launches = {linearize, solve, cost} launch counts.
int bvio_batch_solve_timed(bvio_ctx* ctx, bvio_batch* bb, double out_ms[4], int32_t out_launches[3]) {
  if (!ctx || !bb || !out_ms) return fail(ctx, BVIO_ERR_INVALID, "null argument");
  cudaSetDevice(ctx->device);
  const int passes = bb->bt.max_iters + 1;
  std::vector<cudaEvent_t> ev((size_t)passes * 4);
  for (auto& e : ev) BVIO_CUDA_OK(ctx, cudaEventCreate(&e));
  ctx->launches += ba_launch_reset(bb->bt, ctx->stream);
  for (int it = 0; it < passes; it++)
    ctx->launches += ba_launch_iteration(bb->bt, ctx->stream, it < bb->bt.max_iters, &ev[(size_t)it * 4]);
  ctx->launches += ba_launch_finish(bb->bt, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  out_ms[0] = out_ms[1] = out_ms[2] = out_ms[3] = 0;
  int nl[3] = {0, 0, 0};
  for (int it = 0; it < passes && e == cudaSuccess; it++) {
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ev[it * 4], ev[it * 4 + 1]);
    cudaEventElapsedTime(&b, ev[it * 4 + 1], ev[it * 4 + 2]);
    cudaEventElapsedTime(&c, ev[it * 4 + 2], ev[it * 4 + 3]);
    out_ms[0] += a; out_ms[1] += b; nl[0]++; nl[1]++;
    if (it < bb->bt.max_iters) { out_ms[2] += c; nl[2]++; }
  }
  out_ms[3] = out_ms[0] + out_ms[1] + out_ms[2];
  if (out_launches) { out_launches[0] = nl[0]; out_launches[1] = nl[1]; out_launches[2] = nl[2]; }
  for (auto& x : ev) cudaEventDestroy(x);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("solve_timed: ") + cudaGetErrorString(e));
  return BVIO_OK;
}

// Marginalization tail of Estimator::optimization() (estimator.cpp:816-991).  The host only decides
// which parameter blocks are present / dropped / kept (pure bookkeeping over the window structure,
// what MarginalizationInfo::addResidualBlockInfo does with its address maps); all arithmetic runs in
// ba_marginalize_kernel.  Kept blocks come out in frame order (Pose f, SpeedBias f ascending, then
// Ex_Pose) with the reference's addr_shift applied.  flag 1 and a prior that does not touch Pose[K-2]:
// out->n = -1 (the reference leaves last_marginalization_info untouched, estimator.cpp:926-928).
static int marg_impl(bvio_ctx* ctx, const bvio_window* w_in, const bvio_opts* opts, int32_t flag, bvio_prior_out* out,
                     bvio_marg_job** job) {
  if (!ctx || !w_in || !opts || !out || (flag != 0 && flag != 1)) return fail(ctx, BVIO_ERR_INVALID, "bad arguments");
  if (ctx->marg_inflight) return fail(ctx, BVIO_ERR_INVALID, "a marginalization is in flight: call bvio_marginalize_end first");
  const bool async = job != nullptr;
  if (job) *job = nullptr;
  cudaStream_t st = async ? ctx->copy_stream : ctx->stream;
  // the relocalization factors are not part of the marginalization (estimator.cpp:816-991)
  bvio_window w_copy = *w_in;
  w_copy.n_relo = 0; w_copy.relo_pose = nullptr; w_copy.relo_lm = nullptr; w_copy.relo_xy = nullptr;
  const bvio_window* w = &w_copy;
  const int K = w->K;
  if (opts->estimate_td && K > 14)
    return fail(ctx, BVIO_ERR_INVALID, "bvio_marginalize: K <= 14 with estimate_td");
  int rc = validate(ctx, w, opts, K);
  if (rc) return rc;
  const bvio_prior* pr = w->prior;
  std::vector<char> has_pose(K, 0), has_sb(K, 0);
  bool has_ex = false, has_td = false;
  if (pr) for (int b = 0; b < pr->nblocks; b++) {
    int kind = pr->block_kind[b], f = pr->block_frame[b];
    if (kind == BVIO_BLK_POSE) has_pose[f] = 1;
    else if (kind == BVIO_BLK_SPEEDBIAS) has_sb[f] = 1;
    else if (kind == BVIO_BLK_EXPOSE) has_ex = true;
    else if (kind == BVIO_BLK_TD) has_td = true;
  }
  std::vector<int> dropidx, keepidx;
  if (flag == 1) {
    if (!pr || !has_pose[K - 2]) { out->n = -1; out->nblocks = 0; return BVIO_OK; }   // (async: *job stays NULL, nothing to wait for)
    for (int i = 0; i < 6; i++) dropidx.push_back(15 * (K - 2) + i);
    has_pose[K - 2] = 0;
  } else {
    if (w->preint[1].sum_dt < 10.0) { has_pose[0] = has_sb[0] = has_pose[1] = has_sb[1] = 1; }
    for (int l = 0; l < w->L; l++) {
      int o0 = w->lm_obs_offset[l], o1 = w->lm_obs_offset[l + 1];
      if (w->obs_frame[o0] != 0) continue;
      has_ex = true;
      if (opts->estimate_td) has_td = true;        // para_Td is a block of every ProjectionTdFactor (estimator.cpp:863-871)
      for (int k = o0; k < o1; k++) has_pose[w->obs_frame[k]] = 1;
    }
    for (int i = 0; i < 15; i++) dropidx.push_back(i);
    has_pose[0] = has_sb[0] = 0;
  }
  struct Blk { int kind, frame, idx; };
  std::vector<Blk> blocks;
  int n = 0;
  for (int f = 0; f < K; f++) {
    if (has_pose[f]) { blocks.push_back({BVIO_BLK_POSE, f, n}); for (int i = 0; i < 6; i++) keepidx.push_back(15 * f + i); n += 6; }
    if (has_sb[f]) { blocks.push_back({BVIO_BLK_SPEEDBIAS, f, n}); for (int i = 0; i < 9; i++) keepidx.push_back(15 * f + 6 + i); n += 9; }
  }
  if (has_ex) { blocks.push_back({BVIO_BLK_EXPOSE, 0, n}); for (int i = 0; i < 6; i++) keepidx.push_back(15 * K + i); n += 6; }
  if (has_td) { blocks.push_back({BVIO_BLK_TD, 0, n}); keepidx.push_back(15 * K + 6); n += 1; }
  const int m = (int)dropidx.size();
  if (n > out->cap_n || (int)blocks.size() > out->cap_blocks) return fail(ctx, BVIO_ERR_INVALID, "bvio_prior_out capacity too small");
  out->n = n;
  out->nblocks = (int)blocks.size();
  double* x0 = out->x0;
  for (size_t i = 0; i < blocks.size(); i++) {
    const Blk& bl = blocks[i];
    const double* src = bl.kind == BVIO_BLK_POSE ? w->para_pose + 7 * bl.frame
                        : bl.kind == BVIO_BLK_SPEEDBIAS ? w->para_speed_bias + 9 * bl.frame
                        : bl.kind == BVIO_BLK_TD ? w->para_td : w->para_ex_pose;
    int gs = bl.kind == BVIO_BLK_SPEEDBIAS ? 9 : (bl.kind == BVIO_BLK_TD ? 1 : 7);
    int frame = bl.frame;
    if (bl.kind != BVIO_BLK_EXPOSE && bl.kind != BVIO_BLK_TD) frame = (flag == 0) ? frame - 1 : (frame == K - 1 ? K - 2 : frame);   // addr_shift
    out->block_kind[i] = bl.kind; out->block_frame[i] = frame; out->block_idx[i] = bl.idx;
    memcpy(x0, src, sizeof(double) * gs);
    x0 += gs;
  }
  if (n == 0) return BVIO_OK;
  const int M = 15 * K + 7;
  if (ba_marginalize_smem_bytes(K, pr ? pr->n : 1, n) > 220 * 1024)
    return fail(ctx, BVIO_ERR_UNSUPPORTED, "kept dimension too large for the single-CTA eigen-decomposition");
  bvio_batch* bb = nullptr;
  rc = upload_impl(ctx, w, 1, opts, async ? 1 : 0, 0, &bb);   // async: pipeline slot 0, uploaded and prepared on the copy stream
  if (rc) return rc;
  char* scratch = nullptr;
  Carver cv;
  size_t o_A = cv.take(sizeof(double) * M * M), o_b = cv.take(sizeof(double) * M), o_jac = cv.take(sizeof(double) * n * n);
  size_t o_res = cv.take(sizeof(double) * n), o_drop = cv.take(sizeof(int) * (m + 1)), o_keep = cv.take(sizeof(int) * n);
  size_t o_st = cv.take(sizeof(int) * 4 + sizeof(unsigned long long) * 8);
  int n0 = 0;                               // landmarks anchored at the dropped frame: the first n0 in device order
  if (flag == 0) for (int l = 0; l < w->L; l++) n0 += w->obs_frame[w->lm_obs_offset[l]] == 0;
  size_t o_part = cv.take(sizeof(double) * ba_marginalize_part_doubles(K, ba_marginalize_groups(n0)));
  cudaError_t e = cudaSuccess;
  const size_t stage_bytes = sizeof(double) * ((size_t)n * n + n);
  if (async && ctx->marg_host_bytes < stage_bytes) {       // pinned staging for the asynchronous D2H
    if (ctx->marg_host) cudaFreeHost(ctx->marg_host);
    ctx->marg_host = nullptr; ctx->marg_host_bytes = 0;
    e = cudaMallocHost((void**)&ctx->marg_host, stage_bytes + stage_bytes / 4);
    if (e == cudaSuccess) ctx->marg_host_bytes = stage_bytes + stage_bytes / 4;
  }
  if (ctx->marg_bytes < cv.off) {
    if (ctx->marg_scratch) cudaFree(ctx->marg_scratch);
    ctx->marg_scratch = nullptr; ctx->marg_bytes = 0;
    e = cudaMalloc((void**)&ctx->marg_scratch, cv.off + cv.off / 4);
    if (e == cudaSuccess) ctx->marg_bytes = cv.off + cv.off / 4;
  }
  scratch = ctx->marg_scratch;
  if (e == cudaSuccess) e = cudaMemcpyAsync(scratch + o_drop, dropidx.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(scratch + o_keep, keepidx.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    ctx->launches += ba_launch_marginalize(bb->bt, flag, m, n, n0, (const int*)(scratch + o_drop), (const int*)(scratch + o_keep),
                                           (double*)(scratch + o_A), (double*)(scratch + o_b), (double*)(scratch + o_part), (double*)(scratch + o_jac),
                                           (double*)(scratch + o_res), (int*)(scratch + o_st),
                                           getenv("BVIO_MARG_CHOLESKY") ? 1 : 0, st);
    e = cudaGetLastError();
  }
  if (async) {
    // the dropidx / keepidx vectors above are pageable host memory: their copies have completed (staged) on return
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->marg_host, scratch + o_jac, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->marg_host + sizeof(double) * (size_t)n * n, scratch + o_res, sizeof(double) * n, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) { cudaStreamSynchronize(st); bvio_batch_free(ctx, bb); return fail(ctx, BVIO_ERR_CUDA, std::string("marginalize: ") + cudaGetErrorString(e)); }
    bvio_marg_job* j = new (std::nothrow) bvio_marg_job();
    if (!j) { cudaStreamSynchronize(st); bvio_batch_free(ctx, bb); return fail(ctx, BVIO_ERR_INVALID, "out of host memory"); }
    j->bb = bb; j->out = out; j->n = n;
    ctx->marg_inflight = true;
    *job = j;
    return BVIO_OK;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out->lin_jac, scratch + o_jac, sizeof(double) * n * n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out->lin_res, scratch + o_res, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && getenv("BVIO_DEBUG")) {
    struct { int stv[4]; unsigned long long ph[8]; } dbg;
    memset(&dbg, 0, sizeof dbg);
    cudaMemcpy(&dbg, scratch + o_st, sizeof dbg, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[bvio] marginalize: m=%d n=%d sweeps/rank=%d log10 pivots min %.2f max %.2f | us: factors %.1f prior %.1f drop-eig %.1f "
            "schur %.1f kept-eig %.1f (SM clock %.0f MHz) out %.1f total %.1f\n", m, n, dbg.stv[0], dbg.stv[1] / 100.0, dbg.stv[2] / 100.0,
            (dbg.ph[1] - dbg.ph[0]) * 1e-3, (dbg.ph[2] - dbg.ph[1]) * 1e-3, (dbg.ph[3] - dbg.ph[2]) * 1e-3, (dbg.ph[4] - dbg.ph[3]) * 1e-3,
            (dbg.ph[5] - dbg.ph[4]) * 1e-3, (double)dbg.ph[7] / ((dbg.ph[5] - dbg.ph[4]) * 1e-3 + 1e-9),
            (dbg.ph[6] - dbg.ph[5]) * 1e-3, (dbg.ph[6] - dbg.ph[0]) * 1e-3);
  }
  bvio_batch_free(ctx, bb);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("marginalize: ") + cudaGetErrorString(e));
  for (int i = 0; i < n * n; i++) if (!(out->lin_jac[i] == out->lin_jac[i])) return fail(ctx, BVIO_ERR_NUMERIC, "non-finite prior");
  return BVIO_OK;
}

int bvio_marginalize(bvio_ctx* ctx, const bvio_window* w, const bvio_opts* opts, int32_t flag, bvio_prior_out* out) {
  return marg_impl(ctx, w, opts, flag, out, nullptr);
}

int bvio_marginalize_begin(bvio_ctx* ctx, const bvio_window* w, const bvio_opts* opts, int32_t flag, bvio_prior_out* out,
                           bvio_marg_job** job) {
  if (!job) return fail(ctx, BVIO_ERR_INVALID, "null job");
  return marg_impl(ctx, w, opts, flag, out, job);
}

int bvio_marginalize_end(bvio_ctx* ctx, bvio_marg_job* job) {
  if (!ctx) return BVIO_ERR_INVALID;
  if (!job) return BVIO_OK;                     // begin() had nothing to enqueue (out->n <= 0)
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
  const int n = job->n;
  if (e == cudaSuccess) {
    memcpy(job->out->lin_jac, ctx->marg_host, sizeof(double) * (size_t)n * n);
    memcpy(job->out->lin_res, ctx->marg_host + sizeof(double) * (size_t)n * n, sizeof(double) * n);
  }
  ctx->marg_inflight = false;
  bvio_batch_free(ctx, job->bb);
  bvio_prior_out* out = job->out;
  delete job;
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("marginalize_end: ") + cudaGetErrorString(e));
  for (int i = 0; i < n * n; i++) if (!(out->lin_jac[i] == out->lin_jac[i])) return fail(ctx, BVIO_ERR_NUMERIC, "non-finite prior");
  return BVIO_OK;
}

}  // extern "C"
