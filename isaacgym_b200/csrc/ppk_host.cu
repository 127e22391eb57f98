// Host-buffer session: the same fused step for state tensors that live in HOST memory (isaacgym's
// CPU pipeline, `use_gpu_pipeline: False`).  Only the rigid-body rows the step consumes cross
// PCIe (strided 2-D copies of the id runs), the env batch is cut into chunks and the chunks are
// pipelined over three streams so H2D of chunk c+1, the kernel of chunk c and D2H of chunk c-1
// overlap.  The whole per-step pipeline (copies + kernels + fork/join) is recorded into a CUDA graph
// the second time it is issued with the same buffers and replayed afterwards, so one step costs a
// single graph launch on the host.  The session owns its device staging; everything else follows
// include/ppk.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/ppk.h"

namespace {

constexpr int kRow = 13;
constexpr int kStreams = 3;
constexpr uint32_t kDeferCounterClear = 1u << 8;   // internal phase bit understood by launch_adof

struct Run { int first_row, rows, dst_row; };

}  // namespace

extern "C" int ppk_internal_adof_clear(const PpkBuffers* b, void* stream);

struct PpkHostSession {
  PpkTask host_task;     // as the caller described it (full rigid-body tensor)
  PpkTask dev_task;      // compact rigid-body tensor: only the staged rows, ids remapped
  int64_t max_envs = 0;
  int num_chunks = 1;
  std::vector<Run> runs;       // live rigid-body rows staged per env
  std::vector<Run> init_runs;  // reference-pose rows staged per env (ADOF)
  int dev_bodies = 0;
  int num_flags = 0;
  cudaStream_t streams[kStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_join[kStreams] = {nullptr, nullptr, nullptr};
  // device staging (sized for max_envs)
  float *rb = nullptr, *root = nullptr, *dof = nullptr, *force = nullptr, *pre = nullptr;
  float *init_root = nullptr, *init_dof = nullptr, *init_rb = nullptr, *reset_vel = nullptr, *reset_yz = nullptr;
  float *obs = nullptr, *rew = nullptr;
  int64_t *reset = nullptr, *progress = nullptr, *last_hitter = nullptr;
  uint8_t* flags[PPK_MAX_FLAGS] = {};
  double* stats = nullptr;
  uint32_t* scratch = nullptr;
  bool constants_uploaded = false;
  const void* const_src[3] = {};
  int64_t const_n = 0;
  double* stats_out = nullptr;     // [PPK_NUM_STATS] reduced on the device, copied to the caller's host `stats`
  bool zero_copy = false;   // small per-env buffers are pinned host memory: the kernel reads/writes them in place
  bool zc_launch = false;   // the launch table (reset_ball_vel / reset_ball_pos_yz) is pinned: read in place
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  // recorded pipeline
  PpkBuffers key;
  uint32_t key_phases = 0;
  int key_hits = 0;
  bool key_valid = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t graph_h2d = 0, graph_d2h = 0;
};

namespace {

#define CU(expr)                                   \
  do {                                             \
    if ((expr) != cudaSuccess) {                   \
      cudaGetLastError();                          \
      return PPK_ERR_CUDA;                         \
    }                                              \
  } while (0)

int flags_of(int variant) {
  switch (variant) {
    case PPK_TILT: return 3;
    case PPK_A4: return 6;
    case PPK_NES: return 2;
    case PPK_ALIGN: return 1;
    case PPK_ALIGN2: return 1;
    case PPK_ADOF: return 9;
    default: return 0;
  }
}

// Collapse an id list into runs of consecutive rows and remap the ids.
void add_ids(std::vector<Run>& runs, const int32_t* ids, int n, int32_t* remapped, int& next_row) {
  int j = 0;
  while (j < n) {
    int len = 1;
    while (j + len < n && ids[j + len] == ids[j] + len) ++len;
    runs.push_back({ids[j], len, next_row});
    for (int i = 0; i < len; ++i) remapped[j + i] = next_row + i;
    next_row += len;
    j += len;
  }
}

int find_row(const std::vector<Run>& runs, int row) {
  for (const Run& r : runs)
    if (row >= r.first_row && row < r.first_row + r.rows) return r.dst_row + (row - r.first_row);
  return -1;
}

template <typename T>
int dmalloc(T** p, size_t count) {
  return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T) + 64) == cudaSuccess ? PPK_OK : PPK_ERR_CUDA;
}

bool is_pinned_host(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost && a.devicePointer == p;
}

void drop_graph(PpkHostSession* s) {
  if (s->exec) cudaGraphExecDestroy(s->exec);
  if (s->graph) cudaGraphDestroy(s->graph);
  s->exec = nullptr;
  s->graph = nullptr;
}

// Enqueue one full step: fork from streams[0], chunks round-robin over the streams, join back.
// Used both eagerly and under stream capture (identical work either way).
int enqueue_step(PpkHostSession* s, const PpkBuffers* hb, uint32_t phases) {
  const int64_t n = hb->num_envs;
  const PpkTask& t = s->host_task;
  const int v = t.variant, A = t.num_actors, D = t.num_dofs, B = t.num_bodies, Bd = s->dev_bodies;
  const bool rew = phases & PPK_PHASE_REWARD, rst = phases & PPK_PHASE_RESET, obs = phases & PPK_PHASE_OBS;
  const bool two = (v == PPK_A4 || v == PPK_ALIGN2);
  const int obs_w = (v == PPK_BASE) ? 24 : (v == PPK_ADOF) ? 313 : two ? 2 * 94 : 80;
  const int rew_w = two ? 2 : 1;
  const int pre_stride = hb->pre_ball_stride > 0 ? hb->pre_ball_stride : 2;
  s->h2d_bytes = 0;
  s->d2h_bytes = 0;
  // chunk boundaries on multiples of 32 envs keep every chunk's tile grid aligned
  const int chunks = s->num_chunks;
  int64_t per = ((n + chunks - 1) / chunks + 31) / 32 * 32;
  if (per <= 0) per = 32;
  const bool adof_deferred = (v == PPK_ADOF) && rst;
  const bool zc = s->zero_copy;                      // kernel touches the small pinned host buffers directly
  const bool zc_rows = zc && v != PPK_BASE;          // ... and writes reset rows straight to the host tensors
  cudaStream_t origin = s->streams[0];
  if (adof_deferred) CU(cudaMemsetAsync(s->scratch, 0, sizeof(uint32_t), origin));
  // The launch table is the step's per-call INPUT (the caller refills it in place whenever it wants fresh draws,
  // TILT:857-862 draws per reset): it travels with every step that may reset, never cached.  BASE: one pair for all.
  if (rst && v == PPK_BASE && !s->zc_launch) {
    CU(cudaMemcpyAsync(s->reset_vel, hb->reset_ball_vel, sizeof(float) * 6, cudaMemcpyHostToDevice, origin));
    s->h2d_bytes += sizeof(float) * 6;
  }
  CU(cudaEventRecord(s->ev_fork, origin));
  for (int i = 1; i < kStreams; ++i) CU(cudaStreamWaitEvent(s->streams[i], s->ev_fork, 0));

  int ci = 0;
  for (int64_t lo = 0; lo < n; lo += per, ++ci) {
    const int64_t m = (n - lo < per) ? (n - lo) : per;
    cudaStream_t st = s->streams[ci % kStreams];
    // ---- H2D: only what the step reads
    for (const Run& r : s->runs) {
      CU(cudaMemcpy2DAsync(s->rb + ((size_t)lo * Bd + r.dst_row) * kRow, sizeof(float) * Bd * kRow,
                           hb->rigid_body_states + ((size_t)lo * B + r.first_row) * kRow, sizeof(float) * B * kRow,
                           sizeof(float) * r.rows * kRow, m, cudaMemcpyHostToDevice, st));
      s->h2d_bytes += sizeof(float) * r.rows * kRow * m;
    }
    CU(cudaMemcpyAsync(s->root + (size_t)lo * A * kRow, hb->root_states + (size_t)lo * A * kRow, sizeof(float) * m * A * kRow, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->dof + (size_t)lo * D * 2, hb->dof_states + (size_t)lo * D * 2, sizeof(float) * m * D * 2, cudaMemcpyHostToDevice, st));
    s->h2d_bytes += sizeof(float) * m * (A * kRow + D * 2);
    if (hb->dof_forces) { CU(cudaMemcpyAsync(s->force + (size_t)lo * D, hb->dof_forces + (size_t)lo * D, sizeof(float) * m * D, cudaMemcpyHostToDevice, st)); s->h2d_bytes += sizeof(float) * m * D; }
    if (rew && v != PPK_BASE && !zc) { CU(cudaMemcpyAsync(s->pre + (size_t)lo * pre_stride, hb->pre_ball_states + (size_t)lo * pre_stride, sizeof(float) * m * pre_stride, cudaMemcpyHostToDevice, st)); s->h2d_bytes += sizeof(float) * m * pre_stride; }
    if (!zc) {
      CU(cudaMemcpyAsync(s->progress + lo, hb->progress_buf + lo, sizeof(int64_t) * m, cudaMemcpyHostToDevice, st));
    }
    s->h2d_bytes += sizeof(int64_t) * m;
    if ((!rew || v == PPK_BASE) && !zc) { CU(cudaMemcpyAsync(s->reset + lo, hb->reset_buf + lo, sizeof(int64_t) * m, cudaMemcpyHostToDevice, st)); s->h2d_bytes += sizeof(int64_t) * m; }
    if (rew)
      for (int i = 0; i < s->num_flags; ++i) {
        if (!zc) CU(cudaMemcpyAsync(s->flags[i] + lo, hb->flags[i] + lo, m, cudaMemcpyHostToDevice, st));
        s->h2d_bytes += m;
      }
    if (zc && rew && v != PPK_BASE) s->h2d_bytes += sizeof(float) * m * 2;   // saved ball velocity read in place
    if (v == PPK_ALIGN2 && rew) {
      if (!zc) CU(cudaMemcpyAsync(s->last_hitter + lo, hb->last_hitter + lo, sizeof(int64_t) * m, cudaMemcpyHostToDevice, st));
      s->h2d_bytes += sizeof(int64_t) * m;
    }
    if (rst && v != PPK_BASE && !s->zc_launch) {
      CU(cudaMemcpyAsync(s->reset_vel + (size_t)lo * 3, hb->reset_ball_vel + (size_t)lo * 3, sizeof(float) * m * 3, cudaMemcpyHostToDevice, st));
      s->h2d_bytes += sizeof(float) * m * 3;
      if (v == PPK_ADOF) {
        CU(cudaMemcpyAsync(s->reset_yz + (size_t)lo * 2, hb->reset_ball_pos_yz + (size_t)lo * 2, sizeof(float) * m * 2, cudaMemcpyHostToDevice, st));
        s->h2d_bytes += sizeof(float) * m * 2;
      }
    }

    // ---- the fused step on the chunk
    PpkBuffers db;
    memset(&db, 0, sizeof(db));
    db.struct_size = sizeof(PpkBuffers);
    db.num_envs = m;
    db.rigid_body_states = s->rb + (size_t)lo * Bd * kRow;
    db.root_states = s->root + (size_t)lo * A * kRow;
    db.dof_states = s->dof + (size_t)lo * D * 2;
    db.dof_forces = s->force + (size_t)lo * D;
    db.pre_ball_states = s->pre + (size_t)lo * pre_stride;
    db.pre_ball_stride = pre_stride; db.pre_vx_offset = hb->pre_vx_offset; db.pre_vz_offset = hb->pre_vz_offset;
    db.initial_root_states = s->init_root + (size_t)lo * A * kRow;
    db.initial_dof_states = s->init_dof + (size_t)lo * D * 2;
    db.initial_body_states = s->init_rb ? s->init_rb + (size_t)lo * Bd * kRow : nullptr;
    db.reset_ball_vel = (v == PPK_BASE) ? s->reset_vel : s->reset_vel + (size_t)lo * 3;
    db.reset_ball_pos_yz = s->reset_yz + (size_t)lo * 2;
    db.obs_buf = s->obs + (size_t)lo * obs_w;
    db.rew_buf = s->rew + (size_t)lo * rew_w;
    db.reset_buf = s->reset + lo;
    db.progress_buf = s->progress + lo;
    for (int i = 0; i < s->num_flags; ++i) db.flags[i] = s->flags[i] + lo;
    db.stats = s->stats;
    db.scratch = s->scratch;
    db.last_hitter = s->last_hitter ? s->last_hitter + lo : nullptr;
    if (zc) {
      if (hb->last_hitter) db.last_hitter = hb->last_hitter + lo;
      if (hb->pre_ball_states) db.pre_ball_states = hb->pre_ball_states + (size_t)lo * pre_stride;
      db.rew_buf = hb->rew_buf ? hb->rew_buf + (size_t)lo * rew_w : db.rew_buf;
      db.reset_buf = hb->reset_buf + lo;
      db.progress_buf = hb->progress_buf + lo;
      for (int i = 0; i < s->num_flags; ++i) db.flags[i] = hb->flags[i] + lo;
    }
    if (zc_rows) {
      db.root_states_out = hb->root_states + (size_t)lo * A * kRow;
      db.dof_states_out = hb->dof_states + (size_t)lo * D * 2;
    }
    if (rst && s->zc_launch) {      // pinned launch table: only the rows of resetting envs are read, in place
      db.reset_ball_vel = (v == PPK_BASE) ? hb->reset_ball_vel : hb->reset_ball_vel + (size_t)lo * 3;
      if (hb->reset_ball_pos_yz) db.reset_ball_pos_yz = hb->reset_ball_pos_yz + (size_t)lo * 2;
    }
    int rc = ppk_post_physics_step(&s->dev_task, &db, phases | (adof_deferred ? kDeferCounterClear : 0u), st);
    if (rc != PPK_OK) return rc;

    // ---- D2H: everything the step wrote (ADOF counters wait for the shard-wide clear below)
    if (obs) { CU(cudaMemcpyAsync(hb->obs_buf + (size_t)lo * obs_w, s->obs + (size_t)lo * obs_w, sizeof(float) * m * obs_w, cudaMemcpyDeviceToHost, st)); s->d2h_bytes += sizeof(float) * m * obs_w; }
    if (rew) {
      if (!zc) {
        CU(cudaMemcpyAsync(hb->rew_buf + (size_t)lo * rew_w, s->rew + (size_t)lo * rew_w, sizeof(float) * m * rew_w, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hb->reset_buf + lo, s->reset + lo, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, st));
      }
      s->d2h_bytes += (sizeof(float) * rew_w + sizeof(int64_t)) * m;
    }
    if (phases & (PPK_PHASE_PROGRESS | PPK_PHASE_RESET)) {
      if (!zc) CU(cudaMemcpyAsync(hb->progress_buf + lo, s->progress + lo, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, st));
      s->d2h_bytes += sizeof(int64_t) * m;
    }
    if (rst && !zc_rows) {
      CU(cudaMemcpyAsync(hb->root_states + (size_t)lo * A * kRow, s->root + (size_t)lo * A * kRow, sizeof(float) * m * A * kRow, cudaMemcpyDeviceToHost, st));
      s->d2h_bytes += sizeof(float) * m * A * kRow;
      if (t.reset_dof) { CU(cudaMemcpyAsync(hb->dof_states + (size_t)lo * D * 2, s->dof + (size_t)lo * D * 2, sizeof(float) * m * D * 2, cudaMemcpyDeviceToHost, st)); s->d2h_bytes += sizeof(float) * m * D * 2; }
    }
    if (v == PPK_ALIGN2 && (rew || rst)) {
      if (!zc) CU(cudaMemcpyAsync(hb->last_hitter + lo, s->last_hitter + lo, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, st));
      s->d2h_bytes += sizeof(int64_t) * m;
    }
    if (rew || rst) {
      const int nf = adof_deferred ? 4 : s->num_flags;
      for (int i = 0; i < nf; ++i) {
        if (!zc) CU(cudaMemcpyAsync(hb->flags[i] + lo, s->flags[i] + lo, m, cudaMemcpyDeviceToHost, st));
        s->d2h_bytes += m;
      }
    }
  }
  // join
  for (int i = 1; i < kStreams; ++i) {
    CU(cudaEventRecord(s->ev_join[i], s->streams[i]));
    CU(cudaStreamWaitEvent(origin, s->ev_join[i], 0));
  }
  if (phases & PPK_PHASE_STATS) {
    // the logged sums (TILT:763-766, ADOF:1164-1168): fold the session's slots, hand the 8 doubles to the caller
    int rc = ppk_stats_reduce(s->stats, s->stats_out, origin);
    if (rc != PPK_OK) return rc;
    CU(cudaMemcpyAsync(hb->stats, s->stats_out, sizeof(double) * PPK_NUM_STATS, cudaMemcpyDeviceToHost, origin));
    s->d2h_bytes += sizeof(double) * PPK_NUM_STATS;
  }
  if (adof_deferred) {
    // ADOF:1162-1175: any reset in the shard clears the five counters of ALL envs
    PpkBuffers db;
    memset(&db, 0, sizeof(db));
    db.struct_size = sizeof(PpkBuffers);
    db.num_envs = n;
    for (int i = 0; i < s->num_flags; ++i) db.flags[i] = zc ? hb->flags[i] : s->flags[i];
    db.scratch = s->scratch;
    int rc = ppk_internal_adof_clear(&db, origin);
    if (rc != PPK_OK) return rc;
    for (int i = 4; i < 9; ++i) {
      if (!zc) CU(cudaMemcpyAsync(hb->flags[i], s->flags[i], n, cudaMemcpyDeviceToHost, origin));
      s->d2h_bytes += n;
    }
  }
  return PPK_OK;
}

}  // namespace

extern "C" {

int ppk_host_session_destroy(PpkHostSession* s) {
  if (!s) return PPK_ERR_NULL;
  for (cudaStream_t st : s->streams)
    if (st) cudaStreamSynchronize(st);
  drop_graph(s);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  for (cudaEvent_t e : s->ev_join)
    if (e) cudaEventDestroy(e);
  for (cudaStream_t st : s->streams)
    if (st) cudaStreamDestroy(st);
  void* ptrs[] = {s->rb, s->root, s->dof, s->force, s->pre, s->init_root, s->init_dof, s->init_rb, s->reset_vel,
                  s->reset_yz, s->obs, s->rew, s->reset, s->progress, s->last_hitter, s->stats, s->stats_out, s->scratch};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (uint8_t* f : s->flags)
    if (f) cudaFree(f);
  delete s;
  return PPK_OK;
}

int ppk_host_session_create(const PpkTask* task, int64_t max_envs, int32_t num_chunks, PpkHostSession** out) {
  if (!task || !out) return PPK_ERR_NULL;
  if (task->struct_size != sizeof(PpkTask)) return PPK_ERR_ABI;
  if (max_envs <= 0 || num_chunks <= 0 || task->variant < PPK_BASE || task->variant > PPK_ALIGN2) return PPK_ERR_SHAPE;
  PpkHostSession* s = new (std::nothrow) PpkHostSession();
  if (!s) return PPK_ERR_CUDA;
  memset(&s->key, 0, sizeof(s->key));
  s->host_task = *task;
  s->dev_task = *task;
  s->max_envs = max_envs;
  s->num_chunks = num_chunks;
  s->num_flags = flags_of(task->variant);
  PpkTask& d = s->dev_task;
  int next = 0;
  if (task->variant == PPK_ADOF) {
    // the ADOF kernel stages row windows [0,40) / [0,28) itself: ship those windows, ids unchanged
    s->runs.push_back({0, 40, 0});
    s->init_runs.push_back({0, 40, 0});   // same row count so both tensors share one env stride
    next = 40;
  } else if (task->variant == PPK_BASE) {
    s->runs.push_back({task->paddle_body[0], 1, 0});
    s->runs.push_back({task->paddle_body[1], 1, 1});
    d.paddle_body[0] = 0; d.paddle_body[1] = 1;
    next = 2;
  } else {
    add_ids(s->runs, task->body_ids[0], task->num_body_ids, d.body_ids[0], next);
    if (task->variant == PPK_A4 || task->variant == PPK_ALIGN2) add_ids(s->runs, task->body_ids[1], task->num_body_ids, d.body_ids[1], next);
    for (int h = 0; h < 2; ++h) {
      int row = find_row(s->runs, task->paddle_body[h]);
      if (row < 0 && (h == 0 || task->variant == PPK_A4 || task->variant == PPK_ALIGN2)) {   // paddle not among the obs bodies: stage it too
        s->runs.push_back({task->paddle_body[h], 1, next});
        row = next++;
      }
      d.paddle_body[h] = row < 0 ? 0 : row;
    }
  }
  if (task->variant != PPK_BASE) {
    ++next;                    // pad row: keeps the ADOF bulk staging windows in bounds
    if (next & 1) ++next;      // an even row count: env pairs are 16-byte multiples (tensor-map staging of the family kernel)
  }
  s->dev_bodies = next;
  d.num_bodies = next;
  const int A = task->num_actors, D = task->num_dofs;
  const size_t n = (size_t)max_envs;
  const int obs_w = (task->variant == PPK_BASE) ? 24 : (task->variant == PPK_ADOF) ? 313
                    : (task->variant == PPK_A4 || task->variant == PPK_ALIGN2) ? 2 * 94 : 80;
  int rc = PPK_OK;
  for (int i = 0; i < kStreams && rc == PPK_OK; ++i)
    if (cudaStreamCreateWithFlags(&s->streams[i], cudaStreamNonBlocking) != cudaSuccess) rc = PPK_ERR_CUDA;
  if (rc == PPK_OK && cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess) rc = PPK_ERR_CUDA;
  for (int i = 0; i < kStreams && rc == PPK_OK; ++i)
    if (cudaEventCreateWithFlags(&s->ev_join[i], cudaEventDisableTiming) != cudaSuccess) rc = PPK_ERR_CUDA;
  if (rc == PPK_OK) rc = dmalloc(&s->rb, n * next * kRow);
  if (rc == PPK_OK) rc = dmalloc(&s->root, n * A * kRow);
  if (rc == PPK_OK) rc = dmalloc(&s->dof, n * D * 2);
  if (rc == PPK_OK) rc = dmalloc(&s->force, n * D);
  if (rc == PPK_OK) rc = dmalloc(&s->pre, n * kRow);
  if (rc == PPK_OK) rc = dmalloc(&s->init_root, n * A * kRow);
  if (rc == PPK_OK) rc = dmalloc(&s->init_dof, n * D * 2);
  if (rc == PPK_OK && task->variant == PPK_ADOF) rc = dmalloc(&s->init_rb, n * next * kRow);
  if (rc == PPK_OK) rc = dmalloc(&s->reset_vel, n * 3 + 8);
  if (rc == PPK_OK) rc = dmalloc(&s->reset_yz, n * 2);
  if (rc == PPK_OK) rc = dmalloc(&s->obs, n * obs_w);
  if (rc == PPK_OK) rc = dmalloc(&s->rew, n * 2);
  if (rc == PPK_OK) rc = dmalloc(&s->reset, n);
  if (rc == PPK_OK) rc = dmalloc(&s->progress, n);
  if (rc == PPK_OK && task->variant == PPK_ALIGN2) rc = dmalloc(&s->last_hitter, n);
  for (int i = 0; i < s->num_flags && rc == PPK_OK; ++i) rc = dmalloc(&s->flags[i], n);
  if (rc == PPK_OK) rc = dmalloc(&s->stats, (size_t)PPK_STATS_SLOTS * PPK_NUM_STATS);
  if (rc == PPK_OK) rc = dmalloc(&s->stats_out, (size_t)PPK_NUM_STATS);
  if (rc == PPK_OK) rc = dmalloc(&s->scratch, 16);
  if (rc == PPK_OK && cudaMemset(s->stats, 0, sizeof(double) * PPK_STATS_SLOTS * PPK_NUM_STATS) != cudaSuccess) rc = PPK_ERR_CUDA;
  if (rc == PPK_OK && cudaMemset(s->scratch, 0, 64) != cudaSuccess) rc = PPK_ERR_CUDA;
  if (rc == PPK_OK && cudaMemset(s->rb, 0, n * next * kRow * sizeof(float)) != cudaSuccess) rc = PPK_ERR_CUDA;
  if (rc != PPK_OK) {
    cudaGetLastError();
    ppk_host_session_destroy(s);
    return rc;
  }
  *out = s;
  return PPK_OK;
}

int ppk_host_session_traffic(const PpkHostSession* s, int64_t* h2d, int64_t* d2h) {
  if (!s) return PPK_ERR_NULL;
  if (h2d) *h2d = s->h2d_bytes;
  if (d2h) *d2h = s->d2h_bytes;
  return PPK_OK;
}

int ppk_host_post_physics_step(PpkHostSession* s, const PpkBuffers* hb, uint32_t phases) {
  if (!s || !hb) return PPK_ERR_NULL;
  if (hb->struct_size != sizeof(PpkBuffers)) return PPK_ERR_ABI;
  const int64_t n = hb->num_envs;
  if (n < 0 || n > s->max_envs) return PPK_ERR_SHAPE;
  if (n == 0) return PPK_OK;
  const PpkTask& t = s->host_task;
  const int v = t.variant, A = t.num_actors, D = t.num_dofs, B = t.num_bodies, Bd = s->dev_bodies;
  const bool rew = phases & PPK_PHASE_REWARD, rst = phases & PPK_PHASE_RESET, obs = phases & PPK_PHASE_OBS;
  if (!hb->rigid_body_states || !hb->root_states || !hb->dof_states || !hb->progress_buf || !hb->reset_buf) return PPK_ERR_NULL;
  if (v != PPK_BASE && !hb->dof_forces) return PPK_ERR_NULL;
  if ((rew && !hb->rew_buf) || (obs && !hb->obs_buf)) return PPK_ERR_NULL;
  if (rew && v != PPK_BASE && (!hb->pre_ball_states || hb->pre_ball_stride <= 0 || hb->pre_ball_stride > kRow)) return PPK_ERR_NULL;
  if (rst && (!hb->initial_root_states || !hb->reset_ball_vel)) return PPK_ERR_NULL;
  if ((rew || rst))
    for (int i = 0; i < s->num_flags; ++i)
      if (!hb->flags[i]) return PPK_ERR_NULL;
  if (v == PPK_ALIGN2 && (rew || rst) && !hb->last_hitter) return PPK_ERR_NULL;
  if (v == PPK_ADOF && (!hb->initial_body_states || !hb->initial_dof_states || (rst && !hb->reset_ball_pos_yz))) return PPK_ERR_NULL;
  if ((phases & PPK_PHASE_STATS) && !hb->stats) return PPK_ERR_NULL;     // host [PPK_NUM_STATS] doubles receiving the sums

  // constant tensors (the initial states: TILT:186,214, ADOF:200) go up when their host pointers or the env count change
  const void* csrc[3] = {hb->initial_root_states, hb->initial_dof_states, hb->initial_body_states};
  int64_t const_bytes = 0;
  if (!s->constants_uploaded || s->const_n != n || memcmp(csrc, s->const_src, sizeof(csrc)) != 0) {
    cudaStream_t st = s->streams[0];
    if (hb->initial_root_states) { CU(cudaMemcpyAsync(s->init_root, hb->initial_root_states, sizeof(float) * n * A * kRow, cudaMemcpyHostToDevice, st)); const_bytes += sizeof(float) * n * A * kRow; }
    if (hb->initial_dof_states) { CU(cudaMemcpyAsync(s->init_dof, hb->initial_dof_states, sizeof(float) * n * D * 2, cudaMemcpyHostToDevice, st)); const_bytes += sizeof(float) * n * D * 2; }
    if (hb->initial_body_states && s->init_rb)
      for (const Run& r : s->init_runs) {
        CU(cudaMemcpy2DAsync(s->init_rb + (size_t)r.dst_row * kRow, sizeof(float) * Bd * kRow,
                             hb->initial_body_states + (size_t)r.first_row * kRow, sizeof(float) * B * kRow,
                             sizeof(float) * r.rows * kRow, n, cudaMemcpyHostToDevice, st));
        const_bytes += sizeof(float) * r.rows * kRow * n;
      }
    CU(cudaStreamSynchronize(st));
    memcpy(s->const_src, csrc, sizeof(csrc));
    s->const_n = n;
    s->constants_uploaded = true;
  }

  // same buffers as last time?  1st call: eager; 2nd: record into a graph; afterwards: replay
  const bool same = s->key_valid && s->key_phases == phases && memcmp(&s->key, hb, sizeof(PpkBuffers)) == 0;
  if (!same) {
    drop_graph(s);
    // Pinned host buffers are device-addressable (UVA): the per-env scalars, flags and the rows of
    // resetting envs are then read / written by the kernel in place instead of through ~20 small
    // DMA copies per chunk; the bulk tensors (rigid bodies, roots, DOFs, observations) stay on DMA.
    bool zc = is_pinned_host(hb->progress_buf) && is_pinned_host(hb->reset_buf) && is_pinned_host(hb->rew_buf) &&
              is_pinned_host(hb->pre_ball_states) && is_pinned_host(hb->root_states) && is_pinned_host(hb->dof_states);
    for (int i = 0; i < s->num_flags; ++i) zc = zc && is_pinned_host(hb->flags[i]);
    zc = zc && is_pinned_host(hb->last_hitter);
    s->zero_copy = zc;
    s->zc_launch = is_pinned_host(hb->reset_ball_vel) && is_pinned_host(hb->reset_ball_pos_yz) && hb->reset_ball_vel != nullptr;
    s->key = *hb;
    s->key_phases = phases;
    s->key_valid = true;
    s->key_hits = 0;
  }
  cudaStream_t origin = s->streams[0];
  if (same && s->exec) {
    CU(cudaGraphLaunch(s->exec, origin));
    CU(cudaStreamSynchronize(origin));
    s->h2d_bytes = s->graph_h2d;
    s->d2h_bytes = s->graph_d2h;
    return PPK_OK;
  }
  const bool record = same && s->key_hits >= 1;
  ++s->key_hits;
  if (record) CU(cudaStreamBeginCapture(origin, cudaStreamCaptureModeThreadLocal));
  int rc = enqueue_step(s, hb, phases);
  if (record) {
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(origin, &g);
    if (rc != PPK_OK || e != cudaSuccess || g == nullptr) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      s->key_valid = false;
      return rc != PPK_OK ? rc : PPK_ERR_CUDA;
    }
    s->graph = g;
    if (cudaGraphInstantiate(&s->exec, s->graph, 0) != cudaSuccess) {
      cudaGetLastError();
      drop_graph(s);
      s->key_valid = false;
      return PPK_ERR_CUDA;
    }
    s->graph_h2d = s->h2d_bytes;
    s->graph_d2h = s->d2h_bytes;
    CU(cudaGraphLaunch(s->exec, origin));
  } else if (rc != PPK_OK) {
    for (cudaStream_t st : s->streams) cudaStreamSynchronize(st);
    return rc;
  }
  CU(cudaStreamSynchronize(origin));
  s->h2d_bytes += const_bytes;
  return PPK_OK;
}

}  // extern "C"
