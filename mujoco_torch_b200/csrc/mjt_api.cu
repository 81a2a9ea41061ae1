// C ABI of libmjt_b200 (see include/mjt.h): model upload and kernel sequencing for step / forward.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mjt.h"
#include "k_collision.cuh"
#include "k_constraint.cuh"
#include "k_convex.cuh"
#include "k_hfield.cuh"
#include "k_sensor.cuh"
#include "k_smooth.cuh"
#include "mjt_solve_layout.h"
#include "mjt_common.cuh"

#ifndef MJT_CONSTRAINT_SPLIT
#define MJT_CONSTRAINT_SPLIT 1024   // CTAs per env block in k_constraint (capped by the number of row sources)
#endif
static thread_local std::string g_err;

// Optional per-kernel timing (bench.py's roofline leg): when enabled every launch is bracketed by CUDA events on the
// launching stream; mjt_profile_end() synchronises and returns "name:ms:count;..." totals.
struct ProfRec { const char* name; cudaEvent_t a, b; };
static thread_local std::vector<ProfRec>* g_prof = nullptr;
struct ProfScope {
  ProfRec r;
  cudaStream_t st;
  bool on;
  ProfScope(const char* name, cudaStream_t s) : st(s), on(g_prof != nullptr) {
    if (on) { r.name = name; cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, st); }
  }
  ~ProfScope() {
    if (on) { cudaEventRecord(r.b, st); g_prof->push_back(r); }
  }
};
#define LAUNCH(name, st, ...)            \
  do {                                   \
    ProfScope prof__(name, st);          \
    __VA_ARGS__;                         \
  } while (0)

static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
#define CUDA_OK(expr)                                                                            \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e__));    \
  } while (0)

struct mjtModel {
  int device;
  int* d_blob;
  std::vector<int> h;  // host copy of the blob
  int blob_words;
  int solve_words;  // reduced blob staged by k_solve (stored right after the full blob)
  bool f64;
  int sm_count;
  int max_smem;
  int solve_envs;  // envs (warps) per CTA in k_solve
  size_t solve_smem;
  size_t solve_blob_bytes, solve_region_bytes;  // smem = blob + envs * region
  int solve_lanes;                              // lanes per env in k_solve (32, or 16 for nv <= 16)
  // live-tile launch (models with many, mostly inactive rows): a first k_solve pass whose tile only has room for rcap
  // live rows / jcap J words per env; envs that overflow are flagged and solved by a second pass with the full tile
  bool tile;
  int tile_envs;
  size_t tile_smem, tile_region_bytes;
  // tuning knobs, read from the environment once at mjt_model_create (never on the launch path)
  bool has_hfield;        // some candidate pair is a height-field pair (selects the k_collision instantiation)
  int constraint_split;   // CTAs per env block in k_constraint
  int ovf_envs;           // envs per CTA of the live-tile overflow pass
  int hv(int k) const { return h[k]; }
  bool state_layout() const { return h[H_layout_state] != 0; }
};

static const char* kFieldNames[] = {

    "qpos", "qvel", "act", "time", "qacc", "qacc_warmstart", "act_dot", "xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis",
    "geom_xpos", "geom_xmat", "site_xpos", "site_xmat", "subtree_com", "cinert", "crb", "cdof", "cdof_dot", "cvel", "qM", "qLD",
    "ten_length", "ten_J", "ten_velocity", "actuator_length", "actuator_moment", "actuator_velocity", "actuator_force", "qfrc_bias",
    "qfrc_passive", "qfrc_gravcomp", "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "qfrc_constraint", "contact_dist", "contact_pos",
    "contact_frame", "efc_J", "efc_D", "efc_aref", "efc_frictionloss", "efc_force", "sensordata", "scr_cacc", "scr_cfrc", "scr_live", "scr_ovf",
    "scr_cand_dist", "scr_cand_pos", "scr_cand_frame", "scr_consrc"

};
static_assert(sizeof(kFieldNames) / sizeof(kFieldNames[0]) == NUM_D, "field name table out of sync with mjt_layout.h");

extern "C" int mjt_version(void) { return 1000; }
extern "C" const char* mjt_last_error(void) { return g_err.c_str(); }

extern "C" int mjt_model_create(const void* blob_host, size_t nbytes, int device, mjtModel** out) {
  if (!blob_host || !out || nbytes < MJT_HDR_WORDS * 4) return fail("mjt_model_create: bad arguments");
  const int* w = static_cast<const int*>(blob_host);
  if ((unsigned)w[H_magic] != MJT_MAGIC) return fail("mjt_model_create: bad blob magic (layout mismatch?)");
  if (((size_t)w[H_blob_words] + (size_t)w[H_solve_blob_words]) * 4 != nbytes) return fail("mjt_model_create: blob size mismatch");
  CUDA_OK(cudaSetDevice(device));
  mjtModel* m = new mjtModel();
  m->device = device;
  m->blob_words = w[H_blob_words];
  m->solve_words = w[H_solve_blob_words];
  m->h.assign(w, w + m->blob_words);
  m->f64 = w[H_is_f64] != 0;
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  m->sm_count = prop.multiProcessorCount;
  m->max_smem = (int)prop.sharedMemPerBlockOptin;
  CUDA_OK(cudaMalloc(&m->d_blob, nbytes));
  CUDA_OK(cudaMemcpy(m->d_blob, blob_host, nbytes, cudaMemcpyHostToDevice));
  // k_solve launch shape.  G lanes per env: a half warp for models with nv <= 16 (two envs per warp), else a full warp.
  // Pick the envs per CTA that maximise resident envs per SM under the shared-memory and register budgets.
  const int nv = w[H_nv], nefc = w[H_nefc];
  const size_t tsz = m->f64 ? 8 : 4;
  int G = nv <= 16 ? 16 : 32;
  if (const char* fg = getenv("MJT_SOLVE_G")) {   // experiment knob: half warps with two dofs per lane for 16 < nv <= 32
    if (atoi(fg) == 16 && nv <= 32) G = 16;
  }
  m->solve_lanes = G;
  const bool compact = nefc > MJT_COMPACT_MIN_NEFC;
  const size_t region = (size_t)solve_layout(nv, nefc, w[H_jc_words], w[H_nf], w[H_nsrc], (int)tsz, compact, G).words * tsz;
  const size_t blob_b = ((size_t)m->blob_words + 3) / 4 * 16;
  const size_t sblob_b = ((size_t)m->solve_words + 3) / 4 * 16;
  const size_t sm_total = (size_t)prop.sharedMemPerMultiprocessor;
  // f64 kernel: MJT_SOLVE_F64_THREADS x MJT_SOLVE_F64_MINB threads fit the register file; f32 kernel: 2 x 512
  const int tmax = m->f64 ? MJT_SOLVE_F64_THREADS : 512;
  // resident threads per SM the kernels are compiled for (k_solve's __launch_bounds__): f64 one 512-thread CTA at 128 registers; f32
  // two at 64 registers -- except the two-dofs-per-lane instantiation (nv > 32), whose tile fits one CTA per SM anyway and which
  // therefore keeps 128 registers instead of spilling
  const size_t tcap = (size_t)tmax * (m->f64 ? MJT_SOLVE_F64_MINB : (nv > G ? 1 : 2));
  const int estep = 32 / G, emax = tmax / G;
  const size_t max_smem = (size_t)m->max_smem;
  auto pick_envs = [&](size_t reg, const char* knob) {   // most resident envs per SM wins; ties go to the larger CTA
    int best_e = 0;
    size_t best_envs = 0;
    for (int e = emax; e >= estep; e -= estep) {
      const size_t need = sblob_b + reg * e + 1024;
      if (need > max_smem) continue;
      size_t ctas = sm_total / need;
      if (ctas > 32) ctas = 32;
      size_t envs = ctas * e;
      if (envs * G > tcap) envs = tcap / G;
      if (envs > best_envs) { best_envs = envs; best_e = e; }
    }
    if (const char* force = getenv(knob)) {   // tuning knob: force envs per CTA
      const int e = atoi(force);
      if (e >= estep && e <= emax && e % estep == 0 && sblob_b + reg * e + 1024 <= max_smem) best_e = e;
    }
    return best_e;
  };
  const int best_e = pick_envs(region, "MJT_SOLVE_E");
  if (best_e == 0) {
    delete m;
    return fail("mjt_model_create: model too large for the shared-memory solver tile (nv * nefc)");
  }
  m->solve_envs = best_e;
  m->solve_smem = sblob_b + region * best_e;
  m->solve_blob_bytes = sblob_b;
  m->solve_region_bytes = region;
  m->tile = false;
  m->tile_envs = 0;
  m->tile_smem = m->tile_region_bytes = 0;
  const char* tile_off = getenv("MJT_LIVE_TILE");   // A/B knob: MJT_LIVE_TILE=0 keeps the single full-tile launch
  if (compact && w[H_live_rcap] > 0 && !(tile_off && atoi(tile_off) == 0)) {
    const size_t treg = (size_t)solve_layout(nv, nefc, w[H_jc_words], w[H_nf], w[H_nsrc], (int)tsz, true, G, w[H_live_rcap], w[H_live_jcap]).words * tsz;
    const int te = pick_envs(treg, "MJT_TILE_E");
    if (te > 0 && treg < region) {
      m->tile = true;
      m->tile_envs = te;
      m->tile_region_bytes = treg;
      m->tile_smem = sblob_b + treg * te;
    }
  }
  if (blob_b + 64 > (size_t)m->max_smem) {
    delete m;
    return fail("mjt_model_create: model blob exceeds shared memory");
  }
  m->has_hfield = false;
  for (int p = 0; p < w[H_npair_cand]; ++p) m->has_hfield = m->has_hfield || (w[w[MJT_OFF_I + I_pair_fn] + p] >= FN_HFIELD_SPHERE);
  m->constraint_split = w[H_nsrc] < MJT_CONSTRAINT_SPLIT ? w[H_nsrc] : MJT_CONSTRAINT_SPLIT;
  if (const char* fs = getenv("MJT_CONSTRAINT_SPLIT")) { const int v = atoi(fs); if (v >= 1) m->constraint_split = v < w[H_nsrc] ? v : w[H_nsrc]; }
  m->ovf_envs = 32 / G;   // one warp per CTA: CTAs without a flagged env leave at once, flagged envs pack the SMs
  if (const char* fe = getenv("MJT_OVF_E")) { const int v = atoi(fe); if (v >= 32 / G && v <= best_e && v % (32 / G) == 0) m->ovf_envs = v; }
  *out = m;
  return 0;
}

extern "C" void mjt_model_destroy(mjtModel* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaFree(m->d_blob);
  delete m;
}

extern "C" int mjt_model_layout(const mjtModel* m) { return m->state_layout() ? 1 : 0; }
extern "C" int mjt_model_data_words(const mjtModel* m) { return m->hv(H_pub_words); }
extern "C" int mjt_model_dtype_size(const mjtModel* m) { return m->f64 ? 8 : 4; }
static int rk4_tail_words(const mjtModel* m) { return m->hv(H_nq) + 3 * m->hv(H_nv) + 2 * m->hv(H_na) + 1; }
// work buffer = [workspace (state layout only) | RK4: one more caller-side buffer for the stage passes | RK4 tail]
extern "C" int mjt_model_scratch_words(const mjtModel* m) {
  int n = m->state_layout() ? m->hv(H_data_words) : 0;
  if (m->hv(H_integrator) == INT_RK4) n += m->hv(H_pub_words) + rk4_tail_words(m);
  return n;
}
extern "C" int mjt_model_field(const mjtModel* m, const char* name, int* offset, int* words) {
  for (int k = 0; k < NUM_D; ++k)
    if (strcmp(name, kFieldNames[k]) == 0) {
      // offsets are those of the caller's buffer: the D table in the full layout, the P table in the state layout
      int off = m->hv(MJT_OFF_D + k);
      if (m->state_layout()) {
        off = -1;
        static const int pd[NUM_P] = {D_qpos, D_qvel, D_act, D_time, D_qacc_warmstart, D_qacc, D_act_dot, D_sensordata, D_qfrc_constraint};
        static_assert(P_qacc_warmstart == 4 && P_qacc == 5 && P_qfrc_constraint == 8, "pd[] must follow layout.py P_FIELDS");
        for (int q = 0; q < NUM_P; ++q)
          if (pd[q] == k) off = m->hv(MJT_OFF_P + q);
        if (off < 0) return -1;
      }
      if (offset) *offset = off;
      if (words) *words = m->hv(MJT_OFF_DSZ + k);
      return 0;
    }
  return -1;
}

static MjtIn to_in(const mjtStateIn* s) {
  MjtIn in;
  in.qpos = s->qpos; in.qvel = s->qvel; in.act = s->act; in.time = s->time; in.qacc = s->qacc; in.qacc_warmstart = s->qacc_warmstart;
  in.ctrl = s->ctrl; in.qfrc_applied = s->qfrc_applied; in.xfrc_applied = s->xfrc_applied; in.mocap_pos = s->mocap_pos;
  in.mocap_quat = s->mocap_quat; in.sensordata = s->sensordata; in.qfrc_constraint = s->qfrc_constraint; in.eq_active = s->eq_active;
  return in;
}

static cudaError_t launch_solve(int lanes, int dpl, int compact, int only_flagged, int grid, int threads, size_t smem, size_t smem_attr,
                                cudaStream_t st, const int* blob, const int* sblob, int solve_words, const MjtIn& in, float* out, float* pub, int B, int mode) {
  return mjt_launch_solve_f32(lanes, dpl, compact, 0, 0, only_flagged, grid, threads, smem, smem_attr, st, blob, sblob, solve_words, in, out, pub, B, mode);
}
static cudaError_t launch_solve(int lanes, int dpl, int compact, int only_flagged, int grid, int threads, size_t smem, size_t smem_attr,
                                cudaStream_t st, const int* blob, const int* sblob, int solve_words, const MjtIn& in, double* out, double* pub, int B, int mode) {
  return mjt_launch_solve_f64(lanes, dpl, compact, 0, 0, only_flagged, grid, threads, smem, smem_attr, st, blob, sblob, solve_words, in, out, pub, B, mode);
}

template <typename T>
static int set_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  }
  return 0;
}

// one forward pass (optionally followed by the in-kernel integrator) into `out`
// Threads per CTA of the thread-per-env kernels: full-size CTAs once the batch fills the GPU twice over, otherwise
// smaller CTAs so that a small batch still spreads over all SMs.
static int env_threads(const mjtModel* m, int B, int grid_x = 1) {
  // grid_x: CTAs that share one env block (k_collision: one per candidate pair, k_constraint: one per row-source group) -- they
  // count towards filling the GPU, and every CTA stages the whole model blob, so such kernels keep full-size CTAs much longer
  int t = MJT_ENV_THREADS;
  while (t > 32 && (long long)((B + t - 1) / t) * grid_x < 2 * m->sm_count) t >>= 1;
  return t;
}

// One forward pass (optionally followed by the in-kernel integrator).  `pub` is the caller's buffer; `ws` is where the derived
// quantities go: the same buffer in the full layout, the lent workspace in the state layout.  `upto` stops the sequence after
// a stage (the stage-level API: kinematics() needs k_smooth only, collision() k_smooth + k_collision ...).
enum { UPTO_ALL = 0, UPTO_SMOOTH = 1, UPTO_COLLISION = 2, UPTO_CONSTRAINT = 3 };
template <typename T>
static int run_forward(const mjtModel* m, const MjtIn& in, T* pub, T* ws, int B, int mode, int upto, cudaStream_t st, int* launches) {
  const int threads = env_threads(m, B);
  const int gx = (B + threads - 1) / threads;
  const size_t blob_b = ((size_t)m->blob_words + 3) / 4 * 16;
  const int nsens = m->hv(H_nsensordata);
  int n = 0;
  auto done = [&](void) {
    if (launches) *launches = n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
    return 0;
  };
  if (set_smem<T>((const void*)k_smooth<T>, blob_b)) return 1;
  LAUNCH("k_smooth", st, k_smooth<T><<<gx, threads, blob_b, st>>>(m->d_blob, m->blob_words, in, ws, B, mode));
  ++n;
  if (upto == UPTO_SMOOTH) return done();
  if (m->hv(H_npair_cand) > 0) {
    const int ct = env_threads(m, B, m->hv(H_npair_cand)), cgx = (B + ct - 1) / ct;
    dim3 grid(m->hv(H_npair_cand), cgx < 65535 ? cgx : 65535);
    if (m->has_hfield) {
      if (set_smem<T>((const void*)k_collision<T, true>, blob_b)) return 1;
      LAUNCH("k_collision", st, k_collision<T, true><<<grid, ct, blob_b, st>>>(m->d_blob, m->blob_words, ws, B));
    } else {
      if (set_smem<T>((const void*)k_collision<T, false>, blob_b)) return 1;
      LAUNCH("k_collision", st, k_collision<T, false><<<grid, ct, blob_b, st>>>(m->d_blob, m->blob_words, ws, B));
    }
    ++n;
  }
  if (m->hv(H_topk) > 0) {   // max_contact_points: keep the K closest of the candidate contacts
    if (set_smem<T>((const void*)k_topk<T>, blob_b)) return 1;
    LAUNCH("k_topk", st, k_topk<T><<<gx, threads, blob_b, st>>>(m->d_blob, m->blob_words, ws, B));
    ++n;
  }
  if (upto == UPTO_COLLISION) return done();
  if (m->hv(H_nsrc) > 0) {
    if (set_smem<T>((const void*)k_constraint<T>, blob_b)) return 1;
    // grid.x CTAs share one env block and split its row sources; each CTA stages the model blob once, so a few sources
    // per CTA amortise the staging while still spreading an env block's rows over several SMs
    const int ct = env_threads(m, B, m->constraint_split), cgx = (B + ct - 1) / ct;
    dim3 grid(m->constraint_split, cgx < 65535 ? cgx : 65535);
    LAUNCH("k_constraint", st, k_constraint<T><<<grid, ct, blob_b, st>>>(m->d_blob, m->blob_words, in, ws, B));
    ++n;
  }
  if (upto == UPTO_CONSTRAINT) return done();
  if (nsens > 0) {
    if (set_smem<T>((const void*)k_sensor<T>, blob_b)) return 1;
    LAUNCH("k_sensor", st, k_sensor<T><<<gx, threads, blob_b, st>>>(m->d_blob, m->blob_words, in, ws, pub, B, (1 << 1) | (1 << 2)));
    ++n;
  }
  {
    const int E = m->solve_envs;
    const size_t smem = m->solve_blob_bytes + m->solve_region_bytes * E;
    const int gs = (B + E - 1) / E;
    // COMPACT instantiation (live-row list) for models with more than two 32-row chunks of constraint rows
    const int compact = m->hv(H_nefc) > MJT_COMPACT_MIN_NEFC ? 1 : 0;
    const int lanes = m->solve_lanes;
    if (m->hv(H_nv) > 2 * lanes) return fail("mjt: nv > 64 not supported by the dense solver kernel");
    const int dpl = m->hv(H_nv) <= lanes ? 1 : 2;
    cudaError_t se;
    if (m->tile) {   // live-tile pass over every env, then the full tile for the envs it flagged
      const int TE = m->tile_envs;
      const size_t tsmem = m->solve_blob_bytes + m->tile_region_bytes * TE;
      ProfScope prof__("k_solve", st);
      se = launch_solve(lanes, dpl, 2, 0, (B + TE - 1) / TE, lanes * TE, tsmem, m->tile_smem, st, m->d_blob, m->d_blob + m->blob_words,
                        m->solve_words, in, ws, pub, B, mode);
      ++n;
      if (se == cudaSuccess) {
        const int E2 = m->ovf_envs;
        const size_t smem2 = m->solve_blob_bytes + m->solve_region_bytes * E2;
        se = launch_solve(lanes, dpl, compact, 1, (B + E2 - 1) / E2, lanes * E2, smem2, m->solve_smem, st, m->d_blob, m->d_blob + m->blob_words,
                          m->solve_words, in, ws, pub, B, mode);
      }
    } else {
      ProfScope prof__("k_solve", st);
      se = launch_solve(lanes, dpl, compact, 0, gs, lanes * E, smem, m->solve_smem, st, m->d_blob, m->d_blob + m->blob_words, m->solve_words,
                        in, ws, pub, B, mode);
    }
    if (se != cudaSuccess) return fail(std::string("k_solve launch: ") + cudaGetErrorString(se));
    ++n;
  }
  // acc-stage sensors run only when the solver ran (reference forward.py:390-401 returns early when nefc == 0)
  if (nsens > 0 && m->hv(H_nsens_acc) > 0 && m->hv(H_nefc) > 0) {
    LAUNCH("k_sensor", st, k_sensor<T><<<gx, threads, blob_b, st>>>(m->d_blob, m->blob_words, in, ws, pub, B, 1 << 3));
    ++n;
  }
  return done();
}

template <typename T>
static int run_step(const mjtModel* m, const MjtIn& in, T* out, T* work, int B, int flags, cudaStream_t st, int* launches) {
  int total = 0, n = 0;
  const int fixed = (flags & MJT_FLAG_FIXED_ITERATIONS) ? MJT_MODE_FIXED : 0;
  const bool split = m->state_layout();
  if (split && !work) return fail("mjt_step: the state layout needs the work buffer (mjt_model_scratch_words)");
  T* ws = split ? work : out;
  if (m->hv(H_integrator) != INT_RK4) {
    if (run_forward<T>(m, in, out, ws, B, MJT_MODE_STEP | MJT_MODE_CHECK | fixed, UPTO_ALL, st, &n)) return 1;
    if (launches) *launches = n;
    return 0;
  }
  if (!work) return fail("mjt_step: RK4 needs a work buffer (mjt_model_scratch_words)");
  const int threads = env_threads(m, B), gx = (B + threads - 1) / threads;
  const size_t blob_b = ((size_t)m->blob_words + 3) / 4 * 16;
  const int nq = m->hv(H_nq), nv = m->hv(H_nv), na = m->hv(H_na);
  T* stage_pub = work + (split ? (size_t)m->hv(H_data_words) * B : 0);   // caller-side buffer of the stage passes
  T* stage_ws = split ? work : stage_pub;                                  // state layout: every pass reuses the one workspace
  T* tail = stage_pub + (size_t)m->hv(H_pub_words) * B;
  if (set_smem<T>((const void*)k_rk4_stage<T>, blob_b)) return 1;
  if (run_forward<T>(m, in, out, ws, B, MJT_MODE_FORWARD | MJT_MODE_CHECK | fixed, UPTO_ALL, st, &n)) return 1;
  total += n;
  T* prev = out;
  for (int i = 0; i < 3; ++i) {
    LAUNCH("k_rk4_stage", st, k_rk4_stage<T><<<gx, threads, blob_b, st>>>(m->d_blob, m->blob_words, out, prev, tail, B, i));
    ++total;
    MjtIn sin = in;
    sin.qpos = tail;
    sin.qvel = tail + (size_t)nq * B;
    sin.act = tail + (size_t)(nq + nv) * B;
    sin.time = tail + (size_t)(nq + nv + na) * B;
    // warm start carried from the previous stage's solution (the running Data of the reference's loop)
    sin.qacc_warmstart = prev + (size_t)m->hv(MJT_OFF_P + P_qacc_warmstart) * B;
    sin.sensordata = m->hv(H_nsensordata) > 0 ? prev + (size_t)m->hv(MJT_OFF_P + P_sensordata) * B : nullptr;
    if (run_forward<T>(m, sin, stage_pub, stage_ws, B, MJT_MODE_FORWARD | fixed, UPTO_ALL, st, &n)) return 1;
    total += n;
    prev = stage_pub;
  }
  LAUNCH("k_rk4_stage", st, k_rk4_stage<T><<<gx, threads, blob_b, st>>>(m->d_blob, m->blob_words, out, prev, tail, B, 3));
  ++total;
  if (launches) *launches = total;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

static int check_args(const mjtModel* m, const mjtStateIn* in, void* out, int nenv) {
  if (!m || !in || !out) return fail("mjt: null argument");
  if (nenv <= 0) return fail("mjt: nenv must be positive");
  if (!in->qpos || !in->qvel || !in->time) return fail("mjt: qpos, qvel and time are required inputs");
  if (m->hv(H_nu) > 0 && !in->ctrl) return fail("mjt: ctrl is required for a model with actuators");
  cudaError_t e = cudaSetDevice(m->device);
  if (e != cudaSuccess) return fail(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  return 0;
}

extern "C" int mjt_step(const mjtModel* m, const mjtStateIn* in, void* out, void* work, int nenv, int flags, void* stream) {
  if (check_args(m, in, out, nenv)) return 1;
  if (flags & ~MJT_FLAG_FIXED_ITERATIONS) return fail("mjt_step: unknown flag (only MJT_FLAG_FIXED_ITERATIONS applies to a step)");
  const MjtIn i = to_in(in);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return m->f64 ? run_step<double>(m, i, (double*)out, (double*)work, nenv, flags, st, nullptr)
                : run_step<float>(m, i, (float*)out, (float*)work, nenv, flags, st, nullptr);
}

extern "C" int mjt_forward(const mjtModel* m, const mjtStateIn* in, void* out, void* work, int nenv, int flags, void* stream) {
  if (check_args(m, in, out, nenv)) return 1;
  const bool split = m->state_layout();
  if (split && !work) return fail("mjt_forward: the state layout needs the work buffer (mjt_model_scratch_words)");
  int upto = UPTO_ALL;
  if (flags & MJT_FLAG_UPTO_SMOOTH) upto = UPTO_SMOOTH;
  else if (flags & MJT_FLAG_UPTO_COLLISION) upto = UPTO_COLLISION;
  else if (flags & MJT_FLAG_UPTO_CONSTRAINT) upto = UPTO_CONSTRAINT;
  if (upto != UPTO_ALL && split) return fail("mjt_forward: partial forward passes need the full layout (their outputs are derived fields)");
  const int mode = MJT_MODE_FORWARD | ((flags & MJT_FLAG_FIXED_ITERATIONS) ? MJT_MODE_FIXED : 0);
  const MjtIn i = to_in(in);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return m->f64 ? run_forward<double>(m, i, (double*)out, split ? (double*)work : (double*)out, nenv, mode, upto, st, nullptr)
                : run_forward<float>(m, i, (float*)out, split ? (float*)work : (float*)out, nenv, mode, upto, st, nullptr);
}

extern "C" int mjt_step_launches(const mjtModel* m) {
  int per_forward = m->tile ? 3 : 2;  // k_smooth + k_solve (live-tile models: two k_solve passes)
  if (m->hv(H_npair_cand) > 0) ++per_forward;
  if (m->hv(H_topk) > 0) ++per_forward;
  if (m->hv(H_nsrc) > 0) ++per_forward;
  if (m->hv(H_nsensordata) > 0) ++per_forward;
  if (m->hv(H_nsensordata) > 0 && m->hv(H_nsens_acc) > 0 && m->hv(H_nefc) > 0) ++per_forward;
  return m->hv(H_integrator) == INT_RK4 ? 4 * per_forward + 4 : per_forward;
}

extern "C" void mjt_profile_begin(void) {
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  g_prof->clear();
}

extern "C" int mjt_profile_end(char* buf, size_t nbuf) {
  if (!g_prof) return fail("mjt_profile_end: profiling was not started");
  cudaDeviceSynchronize();
  struct Acc { std::string name; double ms; int n; };
  std::vector<Acc> acc;
  for (auto& r : *g_prof) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    bool found = false;
    for (auto& a : acc)
      if (a.name == r.name) { a.ms += ms; a.n += 1; found = true; }
    if (!found) acc.push_back({r.name, ms, 1});
  }
  delete g_prof;
  g_prof = nullptr;
  std::string out;
  for (auto& a : acc) out += a.name + ":" + std::to_string(a.ms) + ":" + std::to_string(a.n) + ";";
  if (buf && nbuf) {
    strncpy(buf, out.c_str(), nbuf - 1);
    buf[nbuf - 1] = 0;
  }
  return 0;
}

extern "C" int mjt_solve_config(const mjtModel* m, int* envs_per_cta, int* smem_bytes) {
  // the launch that processes every env: the live-tile pass when the model has one
  if (envs_per_cta) *envs_per_cta = m->tile ? m->tile_envs : m->solve_envs;
  if (smem_bytes) *smem_bytes = (int)(m->tile ? m->tile_smem : m->solve_smem);
  return 0;
}
