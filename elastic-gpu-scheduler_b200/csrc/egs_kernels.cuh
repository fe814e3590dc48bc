// egs_kernels.cuh -- sm_100a kernels of the scheduler core (integer / indexing work only;
// HBM- and latency-bound, no tensor cores).
//
//   k_evaluate : the FULL-EVALUATE ("score") kernel -- Trade on every node, no cache shortcut.
//                Roofline kernel: reads 2*G*4 B of rows, writes fit(1)+score(4)+gpu(C) per node.
//   k_pass     : one pod against all nodes honouring the option cache (node.go:61-85), with
//                fit/score digests, first-max selection and the bind (node.go:87-104) fused in
//                the last block.  EGS_MODE_RESCAN launches it once per pod.
//   k_gather_* : /scheduler/filter and /scheduler/priorities over an explicit candidate list.
//   k_bind / k_apply : single-node mutations (Bind, AddPod, ForgetPod).
#pragma once
#include "egs_device.cuh"

#define PASS_THREADS 256

struct OptTable {            // option cache of ONE request shape (slot)
  uint8_t *st;               // [N_pad]  OPT_*
  int32_t *sc;               // [N_pad]  option.Score
  uint8_t *al;               // [EGS_C][N_pad] GPU mask per container (option.Allocated)
  size_t plane;              // N_pad
};

struct Partial { unsigned long long key, fd, sd; int fit; int pad; };

struct PodOut {              // per-pod outputs of the batch loop (device pointers, may be null)
  int32_t *node, *status, *fit_count;
  uint8_t *alloc;            // [P][EGS_C]
  unsigned long long *fit_digest, *score_digest;
};

struct PassArgs {
  const int32_t *core, *mem, *mem_total;   // rows are written by the bind in the last block
  int32_t *core_w, *mem_w;
  int n, policy;
  Req req;
  OptTable t;
  uint8_t *all_st; size_t slot_stride; int n_slots;   // every slot's state plane (UNFIT memo reset)
  uint8_t *vec_fit; int32_t *vec_score;    // optional full vectors
  Partial *partials; unsigned int *ticket;
  int pod; PodOut out; int do_bind;
};

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { unsigned long long x = __shfl_xor_sync(0xffffffffu, v, o); v = x > v ? x : v; }
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide reduction of (max key, sum fd, sum sd, sum fit); result valid in thread 0
__device__ __forceinline__ void block_reduce(unsigned long long &key, unsigned long long &fd,
                                             unsigned long long &sd, int &fit, Partial *sm) {
  key = warp_max_u64(key); fd = warp_sum_u64(fd); sd = warp_sum_u64(sd); fit = warp_sum_i32(fit);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { sm[w].key = key; sm[w].fd = fd; sm[w].sd = sd; sm[w].fit = fit; }
  __syncthreads();
  if (w == 0) {
    key = lane < nw ? sm[lane].key : 0ull; fd = lane < nw ? sm[lane].fd : 0ull;
    sd = lane < nw ? sm[lane].sd : 0ull; fit = lane < nw ? sm[lane].fit : 0;
    key = warp_max_u64(key); fd = warp_sum_u64(fd); sd = warp_sum_u64(sd); fit = warp_sum_i32(fit);
  }
  __syncthreads();
}

// Rows of node w changed: every UNFIT memo of that node is stale (the reference would re-Trade).
__device__ __forceinline__ void memo_reset(uint8_t *all_st, size_t slot_stride, int n_slots, size_t w) {
  for (int s = 0; s < n_slots; s++) {
    uint8_t *p = all_st + (size_t)s * slot_stride + w;
    if (*p == OPT_UNFIT) *p = OPT_ABSENT;
  }
}

// NodeAllocator.Allocate (node.go:87-104) for the winner `w` of the current pod; one thread.
__device__ __forceinline__ int bind_winner(const PassArgs &a, uint32_t w, uint32_t &masks) {
  masks = 0;
  for (int c = 0; c < a.req.C; c++) masks |= (uint32_t)a.t.al[(size_t)c * a.t.plane + w] << (8 * c);
  a.t.st[w] = OPT_ABSENT;                                   // deferred delete, node.go:90-92
  bool ok = transact_row(a.core_w + (size_t)w * EGS_G, a.mem_w + (size_t)w * EGS_G, a.mem_total[w], a.req, masks);
  memo_reset(a.all_st, a.slot_stride, a.n_slots, w);
  return ok ? EGS_OK : EGS_ERR_TRANSACT;
}

// --------------------------------------------------------------------------------------------
// k_pass: one pod, all nodes.  One thread per node; block partials; the last block to finish
// (atomic ticket) folds the partials, picks the first max and binds it.
// --------------------------------------------------------------------------------------------
template <bool SINGLE>
__global__ void __launch_bounds__(PASS_THREADS) k_pass(PassArgs a) {
  __shared__ Partial sm[PASS_THREADS / 32];
  __shared__ bool is_last;
  const int i = blockIdx.x * PASS_THREADS + threadIdx.x;
  unsigned long long key = 0, fd = 0, sd = 0;
  int fit = 0;
  if (i < a.n) {
    uint8_t st = a.t.st[i];
    int score = 0;
    if (st == OPT_ABSENT) {                                 // cache miss -> Trade (node.go:67-71)
      int c[EGS_G], m[EGS_G];
      load_row(a.core, a.mem, (size_t)i, c, m);
      uint32_t masks;
      const int mt = SINGLE ? 0 : a.mem_total[i];
      if (trade_any(c, m, mt, a.req, SINGLE, a.policy, score, masks)) {
        st = OPT_CACHED;
        a.t.sc[i] = score;
        for (int k = 0; k < a.req.C; k++) a.t.al[(size_t)k * a.t.plane + i] = (uint8_t)(masks >> (8 * k));
      } else {
        st = OPT_UNFIT;                                     // not cached by the reference (node.go:68-70)
      }
      a.t.st[i] = st;
    } else if (st == OPT_CACHED) {
      score = a.t.sc[i];                                    // reused without re-validation (node.go:64-66)
    }
    if (st == OPT_CACHED) {
      fit = 1; key = cand_key(score, (uint32_t)i); fd = fit_term((uint32_t)i); sd = score_term((uint32_t)i, score);
    }
    if (a.vec_fit) a.vec_fit[i] = (uint8_t)fit;
    if (a.vec_score) a.vec_score[i] = fit ? score : 0;
  }
  block_reduce(key, fd, sd, fit, sm);
  if (threadIdx.x == 0) {
    Partial p; p.key = key; p.fd = fd; p.sd = sd; p.fit = fit; p.pad = 0;
    a.partials[blockIdx.x] = p;
    __threadfence();
    is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  key = 0; fd = 0; sd = 0; fit = 0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += PASS_THREADS) {
    const volatile Partial *q = a.partials + b;
    unsigned long long k2 = q->key;
    key = k2 > key ? k2 : key; fd += q->fd; sd += q->sd; fit += q->fit;
  }
  block_reduce(key, fd, sd, fit, sm);
  if (threadIdx.x == 0) {
    *a.ticket = 0;
    int node = -1, status = EGS_ERR_NOFIT;
    uint32_t masks = 0;
    if (key != 0 && a.do_bind) {
      node = (int)key_node(key);
      status = bind_winner(a, (uint32_t)node, masks);
      if (status != EGS_OK) masks = 0;
    } else if (key != 0) {
      node = (int)key_node(key); status = EGS_OK;
    }
    const int p = a.pod;
    if (a.out.node) a.out.node[p] = node;
    if (a.out.status) a.out.status[p] = status;
    if (a.out.alloc) for (int c = 0; c < EGS_C; c++) a.out.alloc[(size_t)p * EGS_C + c] = (uint8_t)(masks >> (8 * c));
    if (a.out.fit_count) a.out.fit_count[p] = fit;
    if (a.out.fit_digest) a.out.fit_digest[p] = fd;
    if (a.out.score_digest) a.out.score_digest[p] = sd;
  }
}

// --------------------------------------------------------------------------------------------
// k_evaluate: full evaluate of nodes [0,n) for one request, results to flat output planes.
// ITEMS nodes per thread, strided by the block so every warp-level load instruction covers a
// contiguous 32 x 32 B span; all row loads are issued before any Trade runs.
// --------------------------------------------------------------------------------------------
struct EvalArgs {
  const int32_t *core, *mem, *mem_total;
  int lo, n, policy;                                            // nodes [lo, lo + n)
  Req req;
  uint8_t *fit; int32_t *score; uint8_t *gpu; size_t plane;     // gpu: [C][plane]; all indexed by node id
  uint8_t v_fit, v_unfit;                                       // byte written for fit / unfit (1/0, or OPT_NEW/OPT_UNFIT
};                                                              // when the target is an option table)

template <bool SINGLE, int ITEMS>
__global__ void __launch_bounds__(256) k_evaluate(EvalArgs a) {
  const int base = blockIdx.x * (256 * ITEMS) + threadIdx.x;
  int c[ITEMS][EGS_G], m[ITEMS][EGS_G];
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const int j = base + it * 256;
    if (j < a.n) load_row(a.core, a.mem, (size_t)(a.lo + j), c[it], m[it]);
  }
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const int j = base + it * 256;
    if (j >= a.n) continue;
    const int i = a.lo + j;
    int score; uint32_t masks;
    const int mt = SINGLE ? 0 : a.mem_total[i];
    const bool ok = trade_any(c[it], m[it], mt, a.req, SINGLE, a.policy, score, masks);
    a.fit[i] = ok ? a.v_fit : a.v_unfit;
    a.score[i] = ok ? score : 0;
    if (SINGLE) a.gpu[i] = ok ? (uint8_t)masks : 0;
    else for (int k = 0; k < a.req.C; k++) a.gpu[(size_t)k * a.plane + i] = ok ? (uint8_t)(masks >> (8 * k)) : 0;
  }
}

// --------------------------------------------------------------------------------------------
// verbs over an explicit candidate list (ids == nullptr -> identity)
// --------------------------------------------------------------------------------------------
struct GatherArgs {
  const int32_t *core, *mem, *mem_total;
  int n, n_nodes, policy;
  Req req;
  OptTable t;
  const int32_t *ids;
  uint8_t *out_fit; int32_t *out_score; int *panic_flag;
};

// Assume per node (node.go:61-73)
template <bool SINGLE>
__global__ void __launch_bounds__(256) k_gather_filter(GatherArgs a) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= a.n) return;
  const int i = a.ids ? a.ids[j] : j;
  if (i < 0 || i >= a.n_nodes) { a.out_fit[j] = 0; return; }
  uint8_t st = a.t.st[i];
  if (st == OPT_ABSENT) {
    int c[EGS_G], m[EGS_G], score; uint32_t masks;
    load_row(a.core, a.mem, (size_t)i, c, m);
    if (trade_any(c, m, a.mem_total[i], a.req, SINGLE, a.policy, score, masks)) {
      st = OPT_CACHED;
      a.t.sc[i] = score;
      for (int k = 0; k < a.req.C; k++) a.t.al[(size_t)k * a.t.plane + i] = (uint8_t)(masks >> (8 * k));
    } else {
      st = OPT_UNFIT;
    }
    a.t.st[i] = st;
  }
  a.out_fit[j] = st == OPT_CACHED;
}

// Score per node (node.go:75-85): cached score; no entry -> Assume; fails -> 0, succeeds -> the
// reference dereferences nil (panic) -- flagged.
template <bool SINGLE>
__global__ void __launch_bounds__(256) k_gather_score(GatherArgs a) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= a.n) return;
  const int i = a.ids ? a.ids[j] : j;
  if (i < 0 || i >= a.n_nodes) { a.out_score[j] = 0; return; }   // scheduler.go:176-179
  uint8_t st = a.t.st[i];
  if (st == OPT_CACHED) { a.out_score[j] = a.t.sc[i]; return; }
  if (st == OPT_ABSENT) {
    int c[EGS_G], m[EGS_G], score; uint32_t masks;
    load_row(a.core, a.mem, (size_t)i, c, m);
    if (trade_any(c, m, a.mem_total[i], a.req, SINGLE, a.policy, score, masks)) {
      a.t.st[i] = OPT_CACHED; a.t.sc[i] = score;              // Assume cached it before the nil deref
      for (int k = 0; k < a.req.C; k++) a.t.al[(size_t)k * a.t.plane + i] = (uint8_t)(masks >> (8 * k));
      *a.panic_flag = 1;
    } else {
      a.t.st[i] = OPT_UNFIT;
    }
  }
  a.out_score[j] = 0;
}

struct BindArgs {
  int32_t *core, *mem; const int32_t *mem_total;
  int node; Req req; OptTable t;
  uint8_t *all_st; size_t slot_stride; int n_slots;
  int skip_transact;          // uid already in the node's podsMap (node.go:149)
  int consume;                // 1: Bind (delete the entry); 0: peek
  int32_t *result;            // [0] had entry, [1] status, [2] masks, [3] score
};
__global__ void k_bind(BindArgs a) {
  const size_t w = (size_t)a.node;
  const bool had = a.t.st[w] == OPT_CACHED;
  uint32_t masks = 0; int status = EGS_ERR_NO_OPTION, score = 0;
  if (had) {
    for (int c = 0; c < a.req.C; c++) masks |= (uint32_t)a.t.al[(size_t)c * a.t.plane + w] << (8 * c);
    score = a.t.sc[w];
    status = EGS_OK;
    if (a.consume) {
      a.t.st[w] = OPT_ABSENT;
      if (!a.skip_transact) {
        bool ok = transact_row(a.core + w * EGS_G, a.mem + w * EGS_G, a.mem_total[w], a.req, masks);
        memo_reset(a.all_st, a.slot_stride, a.n_slots, w);
        status = ok ? EGS_OK : EGS_ERR_TRANSACT;
      }
    }
  }
  a.result[0] = had; a.result[1] = status; a.result[2] = (int32_t)masks; a.result[3] = score;
}

// AddPod / ForgetPod with the option rebuilt from annotations (allocate.go:75-93):
// explicit index lists, Transact (gpu.go:153-175) or Cancel (gpu.go:177-191).
#define EGS_CA EGS_MAX_CONTAINERS_APPLY
struct ReqW { int C; int core[EGS_CA], mem[EGS_CA], cnt[EGS_CA]; };   // a pod as AddPod / ForgetPod see it (up to 8 containers)
struct ApplyArgs {
  int32_t *core, *mem; const int32_t *mem_total;
  int node; ReqW req;
  int n_idx[EGS_CA]; int8_t idx[EGS_CA][EGS_G];
  uint8_t *all_st; size_t slot_stride; int n_slots;
  int cancel;
};
__global__ void k_apply(ApplyArgs a) {
  int32_t *c = a.core + (size_t)a.node * EGS_G, *m = a.mem + (size_t)a.node * EGS_G;
  const int mt = a.mem_total[a.node];
  bool stop = false;
  for (int i = 0; i < a.req.C && !stop; i++) {
    const bool whole = a.req.cnt[i] > 0;
    const int lim = whole ? a.n_idx[i] : (a.n_idx[i] > 0 ? 1 : 0);
    for (int j = 0; j < lim; j++) {
      const int g = a.idx[i][j];
      if (a.cancel) {                                          // GPU.Sub gpu.go:41-49
        if (whole) { c[g] = EGS_CORE_PER_GPU; m[g] = mt; } else { c[g] += a.req.core[i]; m[g] += a.req.mem[i]; }
      } else {                                                 // CanAllocate + Add
        if (whole) {
          if (!(c[g] == EGS_CORE_PER_GPU && m[g] == mt)) { stop = true; break; }
          c[g] = 0; m[g] = 0;
        } else {
          if (!(c[g] >= a.req.core[i] && m[g] >= a.req.mem[i])) { stop = true; break; }
          c[g] -= a.req.core[i]; m[g] -= a.req.mem[i];
        }
      }
    }
  }
  memo_reset(a.all_st, a.slot_stride, a.n_slots, (size_t)a.node);
}

// Many AddPod / ForgetPod row updates in ONE launch: the host has grouped the records by node (record order kept
// inside a node); one thread per touched node applies its run with the arithmetic of k_apply.
struct ApplyOp { int node, cancel; ReqW req; int n_idx[EGS_CA]; int8_t idx[EGS_CA][EGS_G]; };
struct ApplyManyArgs {
  int32_t *core, *mem; const int32_t *mem_total;
  const ApplyOp *ops; const int32_t *group_off; int n_groups;   // group g = ops[group_off[g] .. group_off[g+1])
  uint8_t *all_st; size_t slot_stride; int n_slots;
};
__global__ void k_apply_many(ApplyManyArgs a) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= a.n_groups) return;
  const int node = a.ops[a.group_off[gi]].node;
  int32_t *c = a.core + (size_t)node * EGS_G, *m = a.mem + (size_t)node * EGS_G;
  const int mt = a.mem_total[node];
  for (int o = a.group_off[gi]; o < a.group_off[gi + 1]; o++) {
    const ApplyOp &op = a.ops[o];
    bool stop = false;
    for (int i = 0; i < op.req.C && !stop; i++) {
      const bool whole = op.req.cnt[i] > 0;
      const int lim = whole ? op.n_idx[i] : (op.n_idx[i] > 0 ? 1 : 0);
      for (int j = 0; j < lim; j++) {
        const int g = op.idx[i][j];
        if (op.cancel) {                                         // GPU.Sub gpu.go:41-49
          if (whole) { c[g] = EGS_CORE_PER_GPU; m[g] = mt; } else { c[g] += op.req.core[i]; m[g] += op.req.mem[i]; }
        } else {                                                 // CanAllocate + Add; first failure stops, no rollback (gpu.go:153-175)
          if (whole) {
            if (!(c[g] == EGS_CORE_PER_GPU && m[g] == mt)) { stop = true; break; }
            c[g] = 0; m[g] = 0;
          } else {
            if (!(c[g] >= op.req.core[i] && m[g] >= op.req.mem[i])) { stop = true; break; }
            c[g] -= op.req.core[i]; m[g] -= op.req.mem[i];
          }
        }
      }
    }
  }
  memo_reset(a.all_st, a.slot_stride, a.n_slots, (size_t)node);
}

// Rows of nodes [node0, node0+n) were overwritten from the host.  full != 0 (node_set: a fresh
// NodeAllocator, node.go:42-50) drops every option; full == 0 (state_load) only clears the UNFIT
// memos -- cached options stay, stale, exactly like the reference's map would.
__global__ void k_node_reset(uint8_t *all_st, size_t slot_stride, int n_slots, int node0, int n, int full) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int s = 0; s < n_slots; s++) {
    uint8_t *p = all_st + (size_t)s * slot_stride + node0 + i;
    if (full || *p == OPT_UNFIT) *p = OPT_ABSENT;
  }
}
