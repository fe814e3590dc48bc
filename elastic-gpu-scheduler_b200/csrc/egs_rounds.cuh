// egs_rounds.cuh -- EGS_MODE_ROUNDS: the exact round-based decision loop.
//
// Reference semantics (node.go:61-104): a bind changes ONE node's rows and deletes ONE cache
// entry; every other (node, shape) option -- even a stale one -- is by definition unchanged.
// So within a stretch of pods only nodes that WON in that stretch can change.  A round is:
//
//   k_select  (grid, N-proportional, shardable over GPUs): per node and per shape of the
//             round: Trade every absent option (the full-evaluate work), fold fit count and
//             digests, keep the top-RK candidates (score desc, node asc) per shape.
//   k_merge   (one CTA per shape): fold the per-CTA lists, gather each candidate's payload
//             (rows + its options for all round shapes) into the shard's candidate buffer.
//   [ncclAllGather of the candidate buffers when the node list is sharded]
//   k_resolve (ONE warp, shared-memory resident): replays the pods one after the other:
//             winner = max(tracked nodes' options, best untracked list head), Transact,
//             invalidate, outputs; stops early when a list runs dry or the tracked table is
//             full; writes the tracked nodes back.
//
// Option states: OPT_NEW marks an option select evaluated AHEAD of the shape's next filter.
// It is only valid while the node's rows stay unchanged; once a pod of that shape has run
// ("observed"), it is an ordinary cached option (OPT_CACHED).
#pragma once
#include <vector>
#include "egs_kernels.cuh"

#define OPT_NEW 3

#define RK 32        // candidates kept per (shard, shape) per round (one per lane)
#define RT 256       // tracked (touched) nodes per round
#define RS 32        // shapes per round
#define RD 8         // shards
#define SEL_THREADS 128
#define SEL_WARPS (SEL_THREADS / 32)

struct RoundSet { int n; int slot[RS]; };

struct alignas(16) Cand {           // payload of one candidate node (16-byte chunks: cp.async)
  unsigned long long key;           // cand_key(score, node); 0 = empty
  int32_t rc[EGS_G], rm[EGS_G];
  int32_t mt, pad;
  unsigned long long fterm, sbase;    // fit_term(node), score_base(node): hashed once, here
  int32_t sc[RS];
  uint32_t al[RS];
  uint8_t st[RS];
};
struct alignas(16) ShardBuf {       // what one shard contributes to a round
  int32_t len[RS], more[RS], fit[RS], pad0;
  unsigned long long fd[RS], sd[RS];
  alignas(16) Cand cand[RS][RK];
};

struct AggPart { unsigned long long fd, sd; int fit, pad; };

struct TableSet {                   // all option tables of the handle
  uint8_t *st; int32_t *sc; uint8_t *al; size_t n_pad; int n_slots;
};
__device__ __forceinline__ uint8_t *tb_st(const TableSet &t, int slot) { return t.st + (size_t)slot * t.n_pad; }
__device__ __forceinline__ int32_t *tb_sc(const TableSet &t, int slot) { return t.sc + (size_t)slot * t.n_pad; }
__device__ __forceinline__ uint8_t *tb_al(const TableSet &t, int slot) { return t.al + (size_t)slot * EGS_C * t.n_pad; }

struct SelectArgs {
  const int32_t *core, *mem, *mem_total;
  int lo, hi, policy;               // this shard's node range
  RoundSet set;
  Req reqs[RS];
  TableSet tb;
  const uint8_t *obs_pending;       // per slot: shape observed since its OPT_NEW options were made
  unsigned long long *cta_lists;    // [grid][RS][RK]
  AggPart *cta_agg;                 // [grid][RS]
};

// 32 keys, one per lane -> sorted descending across the lanes (bitonic network, 15 exchange steps)
__device__ __forceinline__ unsigned long long warp_sort_desc(unsigned long long v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool take_max = ((lane & j) == 0) == ((lane & k) == 0);
      v = take_max ? (o > v ? o : v) : (o < v ? o : v);
    }
  }
  return v;
}
// two descending lists (one key per lane each) -> the 32 largest of their union, descending:
// max(a[i], b[31-i]) is a bitonic sequence holding exactly those keys; 5 merge steps sort it
__device__ __forceinline__ unsigned long long merge_top32(unsigned long long a, unsigned long long b, int lane) {
  const unsigned long long br = __shfl_sync(0xffffffffu, b, 31 - lane);
  unsigned long long v = a > br ? a : br;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
    v = ((lane & j) == 0) ? (o > v ? o : v) : (o < v ? o : v);
  }
  return v;
}

// --------------------------------------------------------------------------------------------
// k_select
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_THREADS) k_select(SelectArgs a) {
  __shared__ unsigned long long s_list[SEL_WARPS][RS][RK];
  __shared__ AggPart s_agg[SEL_WARPS][RS];
  __shared__ Req s_reqs[RS];          // kernel params are indexed dynamically: stage them in smem
  __shared__ int s_slot[RS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * SEL_WARPS + warp, nwarps = gridDim.x * SEL_WARPS;
  if (threadIdx.x < RS) { s_slot[threadIdx.x] = a.set.slot[threadIdx.x]; s_reqs[threadIdx.x] = a.reqs[threadIdx.x]; }
  __syncthreads();
  for (int s = lane; s < RS; s += 32) { s_agg[warp][s].fd = 0; s_agg[warp][s].sd = 0; s_agg[warp][s].fit = 0; }
  for (int i = lane; i < RS * RK; i += 32) (&s_list[warp][0][0])[i] = 0;
  __syncwarp();
  static_assert(RK == 32, "top-K lists are one key per lane");
  const int n_chunks = (a.hi - a.lo + 127) / 128;
  for (int chunk = gwarp; chunk < n_chunks; chunk += nwarps) {
    const int base = a.lo + chunk * 128 + lane;               // lane handles nodes base + 32*j: tied scores arrive in key order
    unsigned long long h1[4], h2[4];                          // per-node digest hashes, shared by all shapes
#pragma unroll
    for (int j = 0; j < 4; j++) { h1[j] = fit_term((uint32_t)(base + 32 * j)); h2[j] = score_base((uint32_t)(base + 32 * j)); }
    for (int s = 0; s < a.set.n; s++) {
      const int slot = s_slot[s];
      uint8_t *stp = tb_st(a.tb, slot);
      int32_t *scp = tb_sc(a.tb, slot);
      uint8_t *alp = tb_al(a.tb, slot);
      const bool pending = a.obs_pending[slot] != 0;
      const Req &r = s_reqs[s];
      const bool single = req_is_single(r);
      uint8_t st[4]; int sc[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int i = base + 32 * j;
        st[j] = i < a.hi ? stp[i] : (uint8_t)OPT_UNFIT;
        sc[j] = i < a.hi ? scp[i] : 0;
      }
      unsigned long long key[4], fd = 0, sd = 0;
      int fit = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int i = base + 32 * j;
        key[j] = 0;
        if (i >= a.hi) continue;
        const uint8_t st0 = st[j];
        if (st[j] == OPT_NEW && pending) st[j] = OPT_CACHED;
        if (st[j] == OPT_ABSENT) {                             // full evaluate (gpu.go:65-129)
          int c[EGS_G], m[EGS_G]; uint32_t masks;
          load_row(a.core, a.mem, (size_t)i, c, m);
          if (trade_any(c, m, a.mem_total[i], r, single, a.policy, sc[j], masks)) {
            st[j] = OPT_NEW;
            scp[i] = sc[j];
            for (int k = 0; k < r.C; k++) alp[(size_t)k * a.tb.n_pad + i] = (uint8_t)(masks >> (8 * k));
          } else {
            st[j] = OPT_UNFIT;
          }
        }
        if (st[j] != st0) stp[i] = st[j];
        if (st[j] == OPT_CACHED || st[j] == OPT_NEW) {
          key[j] = cand_key(sc[j], (uint32_t)i);
          fit++; fd += h1[j]; sd += score_term_b(h2[j], sc[j]);
        }
      }
      fit = warp_sum_i32(fit); fd = warp_sum_u64(fd); sd = warp_sum_u64(sd);
      if (lane == 0) { s_agg[warp][s].fit += fit; s_agg[warp][s].fd += fd; s_agg[warp][s].sd += sd; }
      // top-32 of this warp for shape s: sort 32 keys, merge sorted lists
      unsigned long long L = s_list[warp][s][lane];
      bool changed = false;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const unsigned long long kth = __shfl_sync(0xffffffffu, L, RK - 1);
        if (__ballot_sync(0xffffffffu, key[j] > kth)) { L = merge_top32(L, warp_sort_desc(key[j], lane), lane); changed = true; }
      }
      if (changed) s_list[warp][s][lane] = L;
      __syncwarp();
    }
  }
  __syncthreads();
  // fold the warps of this CTA: warp w owns shapes s == w (mod SEL_WARPS)
  for (int s = warp; s < a.set.n; s += SEL_WARPS) {
    unsigned long long L = s_list[0][s][lane];
    for (int w = 1; w < SEL_WARPS; w++) L = merge_top32(L, s_list[w][s][lane], lane);
    if (lane < RK) a.cta_lists[((size_t)blockIdx.x * RS + s) * RK + lane] = L;
    if (lane == 0) {
      AggPart t; t.fd = 0; t.sd = 0; t.fit = 0; t.pad = 0;
      for (int w = 0; w < SEL_WARPS; w++) { t.fd += s_agg[w][s].fd; t.sd += s_agg[w][s].sd; t.fit += s_agg[w][s].fit; }
      a.cta_agg[(size_t)blockIdx.x * RS + s] = t;
    }
  }
}

// --------------------------------------------------------------------------------------------
// k_merge: one CTA per shape of the round
// --------------------------------------------------------------------------------------------
struct MergeArgs {
  const int32_t *core, *mem, *mem_total;
  RoundSet set;
  TableSet tb;
  uint8_t *obs_pending;
  const unsigned long long *cta_lists; const AggPart *cta_agg; int n_cta;
  ShardBuf *out;
};

__global__ void __launch_bounds__(256) k_merge(MergeArgs a) {
  __shared__ unsigned long long s_l[8][RK];
  __shared__ AggPart s_a[8];
  __shared__ unsigned long long s_final[RK];
  const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long L = 0;
  AggPart ag; ag.fd = 0; ag.sd = 0; ag.fit = 0; ag.pad = 0;
  for (int c = warp; c < a.n_cta; c += 8) {
    L = merge_top32(L, a.cta_lists[((size_t)c * RS + s) * RK + lane], lane);
    if (lane == 0) { const AggPart p = a.cta_agg[(size_t)c * RS + s]; ag.fd += p.fd; ag.sd += p.sd; ag.fit += p.fit; }
  }
  if (lane < RK) s_l[warp][lane] = L;
  if (lane == 0) s_a[warp] = ag;
  __syncthreads();
  if (warp == 0) {
    L = s_l[0][lane];
    for (int w = 1; w < 8; w++) L = merge_top32(L, s_l[w][lane], lane);
    if (lane < RK) s_final[lane] = L;
    if (lane == 0) {
      AggPart t = s_a[0];
      for (int w = 1; w < 8; w++) { t.fd += s_a[w].fd; t.sd += s_a[w].sd; t.fit += s_a[w].fit; }
      const int len = t.fit < RK ? t.fit : RK;
      a.out->len[s] = len; a.out->more[s] = t.fit > RK; a.out->fit[s] = t.fit; a.out->fd[s] = t.fd; a.out->sd[s] = t.sd;
      a.obs_pending[a.set.slot[s]] = 0;                        // consumed by this round's select
    }
  }
  __syncthreads();
  // payload: 16 threads per candidate, 16 candidates per pass
  for (int k = threadIdx.x >> 4; k < RK; k += 16) {
    const int f = threadIdx.x & 15;
    const unsigned long long key = s_final[k];
    Cand *cd = &a.out->cand[s][k];
    if (key == 0) { if (f == 0) cd->key = 0; continue; }
    const size_t node = key_node(key);
    if (f == 0) { cd->key = key; cd->mt = a.mem_total[node]; cd->pad = 0; }
    if (f == 1) cd->fterm = fit_term((uint32_t)node);
    if (f == 2) cd->sbase = score_base((uint32_t)node);
    if (f < EGS_G) cd->rc[f] = a.core[node * EGS_G + f]; else cd->rm[f - EGS_G] = a.mem[node * EGS_G + f - EGS_G];
    for (int s2 = f; s2 < RS; s2 += 16) {
      uint8_t st = OPT_UNFIT; int32_t sc = 0; uint32_t al = 0;
      if (s2 < a.set.n) {
        const int slot = a.set.slot[s2];
        st = tb_st(a.tb, slot)[node]; sc = tb_sc(a.tb, slot)[node];
        const uint8_t *alp = tb_al(a.tb, slot);
        for (int c = 0; c < EGS_C; c++) al |= (uint32_t)alp[(size_t)c * a.tb.n_pad + node] << (8 * c);
      }
      cd->st[s2] = st; cd->sc[s2] = sc; cd->al[s2] = al;
    }
  }
}

// --------------------------------------------------------------------------------------------
// k_resolve: one warp
// --------------------------------------------------------------------------------------------
struct ResolveArgs {
  int32_t *core, *mem;              // write-back targets
  int lo, hi, policy, n_shards;
  RoundSet set;
  Req reqs[RS];
  TableSet tb;
  uint8_t *obs_pending;
  const ShardBuf *bufs;             // [n_shards]
  const int32_t *pod_slot; int p0, p_limit;
  PodOut out;
  int32_t *done;                    // [0] pods resolved, [1] tracked nodes, [2] stop reason
  long long *prof;                  // EGS_RESOLVE_PROF: 12 counters
};

#define RTW (RT / 32)
struct ResolveSmem {                 // shape-major, padded: lanes = tracked slots OR lanes = shapes are both conflict-free
  unsigned long long lkey[RS][RD * RK];   // untracked candidate lists (sorted, consumed entries zeroed)
  unsigned long long tkey[RS][RT];        // cand_key of a tracked node's option when it is fit (CACHED/NEW), else 0
  uint32_t al[RS][RT + 1];                // option.Allocated masks
  uint8_t st[RS][RT + 4];                 // OPT_*
  unsigned pmask[RS][RTW];                // tracked slots whose option is ABSENT: Trade at the shape's next pod
  unsigned long long hkey[RS][RD];        // current head of each list (0 = none)
  unsigned long long afd[RS], asd[RS];
  int afit[RS];
  int cur[RS][RD], len[RS][RD], more[RS][RD];
  int observed[RS];
  int rq_single[RS], rq_core[RS], rq_mem[RS]; uint32_t rq_cmask[RS];
  int node[RT], mt[RT], dirty[RT];
  unsigned long long fterm[RT];           // fit_term(node) of each tracked slot
  unsigned long long sbase[RT];           // score_base(node) of each tracked slot
  int rc[RT][EGS_G], rm[RT][EGS_G];
  // per-pod outputs, flushed 32 pods at a time with coalesced stores
  int o_node[64], o_status[64], o_fit[64]; uint32_t o_alloc[64]; unsigned long long o_fd[64], o_sd[64];
  int8_t set_idx[2048];                   // option-table slot id -> index in the round's shape set (-1: not in it)
  Req reqs[RS];
  int hset[512];                          // open-addressed set of tracked node ids (-1 empty): lazy list maintenance
  int hpay_node[RS];                      // node whose payload sits (or is arriving) in hpay[s]; -1 none
  alignas(16) Cand hpay[RS];              // prefetched payload of each shape's best untracked head
};

__device__ __forceinline__ unsigned hset_slot(uint32_t node) { return (node * 2654435761u) >> 23; }   // 9 bits
__device__ __forceinline__ bool hset_has(const ResolveSmem &S, uint32_t node) {
  for (unsigned i = hset_slot(node);; i = (i + 1) & 511u) {
    const int v = S.hset[i];
    if (v == (int)node) return true;
    if (v < 0) return false;
  }
}
__device__ __forceinline__ void hset_add(ResolveSmem &S, uint32_t node) {   // one lane; at most RT (256) entries
  unsigned i = hset_slot(node);
  while (S.hset[i] >= 0) i = (i + 1) & 511u;
  S.hset[i] = (int)node;
}
// advance list (s, d) past entries whose node is tracked by now, and cache its head
__device__ __forceinline__ void list_head_update(ResolveSmem &S, int s, int d) {
  int c = S.cur[s][d];
  const int len = S.len[s][d];
  unsigned long long k = 0;
  while (c < len) { k = S.lkey[s][d * RK + c]; if (!hset_has(S, key_node(k))) break; c++; }
  S.cur[s][d] = c;
  S.hkey[s][d] = c < len ? k : 0ull;
}

// pods [p_first, p_first + n), n <= 32, sit in ring entries (rel0 + i) & 63
__device__ __forceinline__ void flush_outputs(const ResolveSmem &S, const PodOut &out, int p_first, int rel0, int n, int lane) {
  if (lane < n) {
    const size_t p = (size_t)p_first + lane;
    const int r = (rel0 + lane) & 63;
    if (out.node) out.node[p] = S.o_node[r];
    if (out.status) out.status[p] = S.o_status[r];
    if (out.fit_count) out.fit_count[p] = S.o_fit[r];
    if (out.fit_digest) out.fit_digest[p] = S.o_fd[r];
    if (out.score_digest) out.score_digest[p] = S.o_sd[r];
    if (out.alloc) reinterpret_cast<uint32_t *>(out.alloc)[p] = S.o_alloc[r];
  }
}

#ifdef EGS_RESOLVE_PROF
#define PROF_T(i) { long long now_ = clock64(); prof[i] += now_ - tprev; tprev = now_; }
#define PROF_C(i, v) { prof[i] += (v); }
#define PROF_PTR prof
#else
#define PROF_T(i)
#define PROF_C(i, v)
#define PROF_PTR nullptr
#endif
// Segmented (8-lane group) reductions with FULL-mask shuffles: xor offsets 1,2,4 never leave a group,
// so all four groups reduce at once.  (Collectives with a different member mask per group are issued
// one group at a time by the hardware.)
__device__ __forceinline__ int seg8_max(int v) {
  v = max(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = max(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return max(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ unsigned seg8_minu(unsigned v) {
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return min(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ int seg8_add(int v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v + __shfl_xor_sync(0xffffffffu, v, 4);
}
__device__ __forceinline__ unsigned long long seg8_max64(unsigned long long v) {
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, v, o); v = x > v ? x : v; }
  return v;
}

// Best untracked head of shape s over the shards -> (key, shard); warp-uniform result.
__device__ __forceinline__ unsigned long long best_head(const ResolveSmem &S, int s, int D, int &d_out) {
  unsigned long long b = 0; int bd = 0;
  for (int d = 0; d < D; d++) { const unsigned long long k = S.hkey[s][d]; if (k > b) { b = k; bd = d; } }
  d_out = bd;
  return b;
}
// Start the asynchronous copy of that head's payload into S.hpay[s] (23 x 16 B, one chunk per lane).
__device__ __forceinline__ void prefetch_head(ResolveSmem &S, const ResolveArgs &a, int s, int D, int lane) {
  int d; const unsigned long long k = best_head(S, s, D, d);
  if (k == 0) { if (lane == 0) S.hpay_node[s] = -1; return; }
  const char *src = reinterpret_cast<const char *>(&a.bufs[d].cand[s][S.cur[s][d]]);
  if (lane < (int)(sizeof(Cand) / 16)) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(reinterpret_cast<char *>(&S.hpay[s]) + lane * 16);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + lane * 16));
  }
  if (lane == 0) S.hpay_node[s] = (int)key_node(k);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct PodRec {                     // decision of one pod, made by an 8-lane group (fast path)
  int s, u, bk, from_head, t, d, fit, pad;
  unsigned long long win, fd, sd;
};

// Commit of one decided pod: NodeAllocator.Allocate (node.go:87-104) on the tracked copy, or NOFIT.
// (fitc, ofd, osd) are the shape's aggregates after this pod's filter; warp-uniform arguments.
__device__ __forceinline__ void commit_pod(ResolveSmem &S, const ResolveArgs &a, int lane, int rel, int s,
                                           unsigned long long win, int from_head, int t_in, int d, int fitc,
                                           unsigned long long ofd, unsigned long long osd, bool mono, int ns, int D,
                                           int &nT, int n_observed, long long *prof) {
#ifdef EGS_RESOLVE_PROF
  long long tprev = clock64();
#endif
  int o_node = -1, o_status = EGS_ERR_NOFIT; uint32_t o_masks = 0;
  if (win != 0) {
    int t = t_in;
    if (from_head) {
      // an untracked node wins: it becomes tracked.  Its payload was prefetched into shared memory when it
      // became the best head of this shape (cp.async); fall back to the candidate buffer otherwise.
      t = nT++;
      const uint32_t w = key_node(win);
      cp_async_wait_all();
      __syncwarp();
      PROF_T(12)
      {
        int rowv, mtv, scv; uint32_t alv; uint8_t st; unsigned long long ft, sb;
        if (S.hpay_node[s] == (int)w) {                            // warp-uniform: payload already in shared memory
          const Cand &cd = S.hpay[s];
          rowv = lane < 2 * EGS_G ? cd.rc[lane] : 0; mtv = cd.mt; st = cd.st[lane]; scv = cd.sc[lane]; alv = cd.al[lane]; ft = cd.fterm; sb = cd.sbase;
        } else {
          const Cand &cd = a.bufs[d].cand[s][S.cur[s][d]];
          rowv = lane < 2 * EGS_G ? cd.rc[lane] : 0; mtv = cd.mt; st = cd.st[lane]; scv = cd.sc[lane]; alv = cd.al[lane]; ft = cd.fterm; sb = cd.sbase;
        }
        if (lane < EGS_G) S.rc[t][lane] = rowv; else if (lane < 2 * EGS_G) S.rm[t][lane - EGS_G] = rowv;   // rc[8], rm[8] contiguous
        if (lane == 0) { S.node[t] = (int)w; S.mt[t] = mtv; S.dirty[t] = 0; S.fterm[t] = ft; S.sbase[t] = sb; hset_add(S, w); }
        if (st == OPT_NEW && S.observed[lane]) st = OPT_CACHED;   // lane == shape index
        S.st[lane][t] = st; S.al[lane][t] = alv;
        S.tkey[lane][t] = (st == OPT_CACHED || st == OPT_NEW) ? cand_key(scv, w) : 0ull;
        if (st == OPT_ABSENT) S.pmask[lane][t >> 5] |= 1u << (t & 31);
      }
      __syncwarp();
      PROF_T(13)
      // it leaves the untracked lists lazily: only lists whose HEAD is this node advance now (entries deeper
      // in a list are skipped when they surface); then the changed heads get their payloads prefetched
      for (int i = lane; i < ns * D; i += 32) {
        const unsigned long long hk = S.hkey[i / D][i % D];
        if (hk != 0 && key_node(hk) == w) list_head_update(S, i / D, i % D);
      }
      __syncwarp();
      PROF_T(14)
      if (lane < ns) {                                            // lane == shape: its own 23 x 16 B copies, in lockstep
        int dd; const unsigned long long k = best_head(S, lane, D, dd);
        const int want = k ? (int)key_node(k) : -1;
        if (want != S.hpay_node[lane]) {
          S.hpay_node[lane] = want;
          if (k) {
            const char *src = reinterpret_cast<const char *>(&a.bufs[dd].cand[lane][S.cur[lane][dd]]);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&S.hpay[lane]);
#pragma unroll
            for (int q = 0; q < (int)(sizeof(Cand) / 16); q++)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * 16), "l"(src + q * 16));
          }
        }
      }
      cp_async_commit();
      PROF_T(15) PROF_C(11, 1)
    }
    o_node = S.node[t];
    const int single = S.rq_single[s];
    const uint32_t masks = S.al[s][t] & S.rq_cmask[s];
    const unsigned pbit = 1u << (t & 31);
    int ok = 0;
    // deferred delete of the option (node.go:90-92) + aggregates: computed by every lane (broadcast loads),
    // stored by lane 0 -- no divergent region on the common single-container path
    const int nfit = fitc - 1;
    const unsigned long long nfd = ofd - S.fterm[t], nsd = osd - score_term_b(S.sbase[t], key_score(win));
    const unsigned npm = S.pmask[s][t >> 5] | pbit;
    if (single) {                                                 // GPUs.Transact gpu.go:164-171
      const int g = __ffs(masks) - 1;
      const int c = S.rc[t][g], m = S.rm[t][g], rc = S.rq_core[s], rm = S.rq_mem[s];
      ok = (c >= rc && m >= rm) ? 1 : 0;
      __syncwarp();                                               // all lanes have read before lane 0 writes
      if (lane == 0) {
        S.st[s][t] = OPT_ABSENT; S.tkey[s][t] = 0; S.pmask[s][t >> 5] = npm;
        S.afit[s] = nfit; S.afd[s] = nfd; S.asd[s] = nsd; S.dirty[t] = 1;
        if (ok) { S.rc[t][g] = c - rc; S.rm[t][g] = m - rm; }
      }
    } else {
      __syncwarp();
      if (lane == 0) {
        S.st[s][t] = OPT_ABSENT; S.tkey[s][t] = 0; S.pmask[s][t >> 5] = npm;
        S.afit[s] = nfit; S.afd[s] = nfd; S.asd[s] = nsd; S.dirty[t] = 1;
        ok = transact_row(S.rc[t], S.rm[t], S.mt[t], S.reqs[s], masks) ? 1 : 0;
      }
      ok = __shfl_sync(0xffffffffu, ok, 0);
    }
    // Rows changed.  (a) not-yet-observed NEW options of this node are void (only while some shape of the
    // round is unobserved); (b) UNFIT memos are void -- unless every request of the round is >= 0: rows
    // then only decrease and an option that did not fit can never fit (exact shortcut).
    if ((!mono || n_observed < ns) && lane < ns && lane != s) {
      const uint8_t v = S.st[lane][t];
      if (v == OPT_UNFIT && !mono) { S.st[lane][t] = OPT_ABSENT; S.pmask[lane][t >> 5] |= pbit; }
      else if (v == OPT_NEW && !S.observed[lane]) {
        const unsigned long long k2 = S.tkey[lane][t];
        S.st[lane][t] = OPT_ABSENT; S.tkey[lane][t] = 0; S.pmask[lane][t >> 5] |= pbit;
        S.afit[lane] -= 1; S.afd[lane] -= S.fterm[t]; S.asd[lane] -= score_term_b(S.sbase[t], key_score(k2));
      }
    }
    __syncwarp();
    o_status = ok ? EGS_OK : EGS_ERR_TRANSACT;
    o_masks = ok ? masks : 0;
  }
  if (lane == 0) {
    const int r = rel & 63;
    S.o_node[r] = o_node; S.o_status[r] = o_status; S.o_fit[r] = fitc; S.o_fd[r] = ofd; S.o_sd[r] = osd; S.o_alloc[r] = o_masks;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32) k_resolve(ResolveArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ResolveSmem &S = *reinterpret_cast<ResolveSmem *>(smem_raw);
  __shared__ PodRec recs[4];
  const int lane = threadIdx.x;
  const int D = a.n_shards, ns = a.set.n, DK = D * RK;
  const int grp = lane >> 3, gl = lane & 7;                      // 8-lane groups: one lane per GPU of a node
  const unsigned gmask = 0xFFu << (8 * grp);
  // ---- prologue
  for (int i = lane; i < 512; i += 32) S.hset[i] = -1;
  bool mono = true;                                              // all requests >= 0: rows only decrease in this round
  {
    const int s = lane;                                          // RS == 32: one shape per lane
    int fit = 0; unsigned long long fd = 0, sd = 0;
    for (int d = 0; d < D; d++) {
      const ShardBuf &b = a.bufs[d];
      S.cur[s][d] = 0; S.len[s][d] = s < ns ? b.len[s] : 0; S.more[s][d] = s < ns ? b.more[s] : 0;
      if (s < ns) { fit += b.fit[s]; fd += b.fd[s]; sd += b.sd[s]; }
    }
    S.afit[s] = fit; S.afd[s] = fd; S.asd[s] = sd; S.observed[s] = 0;
    for (int w = 0; w < RTW; w++) S.pmask[s][w] = 0;
    if (s < ns) {
      S.reqs[s] = a.reqs[s];
      const Req &r = a.reqs[s];
      S.rq_single[s] = req_is_single(r); S.rq_core[s] = r.core[0]; S.rq_mem[s] = r.mem[0];
      S.rq_cmask[s] = r.C >= 4 ? 0xFFFFFFFFu : ((1u << (8 * r.C)) - 1u);   // alloc planes >= C are never written
      for (int c = 0; c < r.C; c++) mono &= r.core[c] >= 0 && r.mem[c] >= 0;
    }
  }
  mono = __all_sync(0xffffffffu, mono);
  for (int s = 0; s < ns; s++)
    for (int e = lane; e < DK; e += 32) S.lkey[s][e] = a.bufs[e / RK].cand[s][e % RK].key;
  __syncwarp();
  for (int i = lane; i < ns * D; i += 32) list_head_update(S, i / D, i % D);
  __syncwarp();
  for (int s = 0; s < ns; s++) prefetch_head(S, a, s, D, lane);
  cp_async_commit();
  int nT = 0, done = 0, reason = 0, n_observed = 0, flushed = 0;
  for (int i = lane; i < 2048; i += 32) S.set_idx[i] = -1;
  __syncwarp();
  if (lane < ns && a.set.slot[lane] < 2048) S.set_idx[a.set.slot[lane]] = (int8_t)lane;
  __syncwarp();
  // pod -> shape index of the round, resolved 32 pods at a time (current block + the next one, so that a
  // 4-pod window may straddle the block boundary); -1: shape not in the set / past the limit
  auto load_block = [&](int blk) -> int {
    const int q = a.p0 + blk * 32 + lane;
    if (q >= a.p_limit) return -1;
    const int slot = a.pod_slot[q];
    return slot < 2048 ? (int)S.set_idx[slot] : -1;
  };
  int cur_blk = 0, myshape = load_block(0), nxshape = load_block(1);
  // ---- sequential replay
  int p = a.p0;
#ifdef EGS_RESOLVE_PROF
  long long prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tprev = clock64();
#endif
  while (p < a.p_limit) {
    const int rel = p - a.p0;
    if ((rel >> 5) != cur_blk) { cur_blk = rel >> 5; myshape = nxshape; nxshape = load_block(cur_blk + 1); }
    while (rel - flushed >= 32) { __syncwarp(); flush_outputs(S, a.out, a.p0 + flushed, flushed, 32, lane); flushed += 32; }
    const int ri = rel & 31;
    // ======== fast path: up to 4 consecutive pods with distinct single-container shapes, decided by four
    // 8-lane groups at once.  Preconditions make the decisions independent of each other's commits except
    // for the hazards checked below; requires rows to be monotone and every shape of the round observed.
    if (mono && n_observed == ns && nT + 4 <= RT) {
      int sq[4]; int W = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int i = ri + q;
        const int v0 = __shfl_sync(0xffffffffu, myshape, i & 31), v1 = __shfl_sync(0xffffffffu, nxshape, i & 31);
        sq[q] = i < 32 ? v0 : v1;
      }
      {
        bool okq = true;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          okq = okq && sq[q] >= 0;
          for (int i = 0; i < q; i++) okq = okq && sq[q] != sq[i];
          if (okq) W = q + 1;
        }
      }
      const int sg = grp == 0 ? sq[0] : grp == 1 ? sq[1] : grp == 2 ? sq[2] : max(sq[3], 0);
      const int sgs = max(sg, 0);                                  // safe index for inactive groups
      // per group: single shape? at most one pending slot? list dry?   (all lanes execute; no group masks)
      unsigned long long head = 0; int u;
      {
        const unsigned word = gl < ((nT + 31) >> 5) ? S.pmask[sgs][gl] : 0u;
        const int cnt = seg8_add(__popc(word));
        u = seg8_max(word ? gl * 32 + __ffs(word) - 1 : -1);
        bool dry = false;
        if (gl < D) { head = S.hkey[sgs][gl]; dry = head == 0 && S.more[sgs][gl] != 0; }
        const bool bad = cnt > 1 || !S.rq_single[sgs] || dry;
        const unsigned badm = __ballot_sync(0xffffffffu, bad);
#pragma unroll
        for (int q = 3; q >= 0; q--) if ((badm >> (8 * q)) & 0xFFu) W = min(W, q);
      }
      PROF_T(0)
      if (W >= 2) {
        // ---- decisions (read-only on shared state); groups >= W compute on a valid shape and are ignored
        const int rq_c = S.rq_core[sgs], rq_m = S.rq_mem[sgs];
        const int uu = max(u, 0);
        // Trade of the one absent option, lane == GPU
        const int c = S.rc[uu][gl], m = S.rm[uu][gl];
        const int cmin = (int)seg8_minu((unsigned)c), mmin = (int)seg8_minu((unsigned)m);
        const int c1 = seg8_max(c), m1 = seg8_max(m);
        const int c2 = seg8_max(c == c1 ? INT32_MIN : c), m2 = seg8_max(m == m1 ? INT32_MIN : m);
        const unsigned ceq = (__ballot_sync(0xffffffffu, c == c1) >> (8 * grp)) & 0xFFu;
        const unsigned meq = (__ballot_sync(0xffffffffu, m == m1) >> (8 * grp)) & 0xFFu;
        const int cex = (c == c1 && __popc(ceq) == 1) ? c2 : c1, mex = (m == m1 && __popc(meq) == 1) ? m2 : m1;
        const bool okt = c >= rq_c && m >= rq_m;
        const int nc = c - rq_c, nm = m - rq_m;
        const int x = (max(mex, nm) + max(cex, nc)) - (min(mmin, nm) + min(cmin, nc));
        const int key = (!okt || u < 0) ? -1 : (a.policy == EGS_BINPACK ? (x >> 2) * 8 + gl : gl);
        const int bk = seg8_max(key);
        const int sc = (bk >= 0 && a.policy == EGS_BINPACK) ? (bk >> 3) * 100 : 0;
        const uint32_t nd = (uint32_t)S.node[uu];
        const unsigned long long tradekey = bk >= 0 ? cand_key(sc, nd) : 0ull;
        unsigned long long best = 0; int best_t = -1;
        for (int t = gl; t < nT; t += 8) {                         // the pending slot holds key 0 (zeroed by its bind)
          const unsigned long long k = S.tkey[sgs][t];
          if (k > best) { best = k; best_t = t; }
        }
        if (tradekey > best) { best = tradekey; best_t = u; }
        const bool from_head = head > best;
        const unsigned long long mine = from_head ? head : best;
        const unsigned long long win = seg8_max64(mine);
        const unsigned wm = (__ballot_sync(0xffffffffu, mine == win && win != 0) >> (8 * grp)) & 0xFFu;
        const int ownl = wm ? __ffs(wm) - 1 : 0;                     // lane inside the group
        const int fh = __shfl_sync(0xffffffffu, (int)from_head, grp * 8 + ownl);
        const int tw = __shfl_sync(0xffffffffu, best_t, grp * 8 + ownl);
        // the shape's aggregates after this pod's filter (group-uniform)
        const int fit = S.afit[sgs] + (bk >= 0);
        const unsigned long long fd = S.afd[sgs] + (bk >= 0 ? S.fterm[uu] : 0ull);
        const unsigned long long sd = S.asd[sgs] + (bk >= 0 ? score_term_b(S.sbase[uu], sc) : 0ull);
        PROF_T(1)
        // ---- hazards: a head-win changes lists / the tracked set for everyone after it; a pending node
        // that an earlier pod of the group binds must be Traded on the rows AFTER that bind.
        int hfh[4], ht[4], hu[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          hfh[q] = __shfl_sync(0xffffffffu, (int)(win != 0 && fh), 8 * q);
          ht[q] = __shfl_sync(0xffffffffu, (win != 0 && !fh) ? tw : -2 - q, 8 * q);
          hu[q] = __shfl_sync(0xffffffffu, u, 8 * q);
        }
        int Wc = W;
#pragma unroll
        for (int j = 3; j >= 1; j--) {
          bool hz = false;
#pragma unroll
          for (int i = 0; i < j; i++) hz |= hfh[i] || (hu[j] >= 0 && hu[j] == ht[i]);
          if (hz && j < Wc) Wc = j;
        }
        // pods that can commit side by side: tracked wins on pairwise distinct slots (and NOFIT pods);
        // a head-win (always the last committed pod) goes through the sequential commit
        int Ws = Wc;
#pragma unroll
        for (int j = 3; j >= 0; j--) {
          bool dup = hfh[j];
#pragma unroll
          for (int i = 0; i < j; i++) dup |= ht[i] == ht[j];
          if (dup && j < Ws) Ws = j;
        }
        PROF_T(2)
        // ---- commits side by side: group leaders write disjoint rows (shape) and slots (node).
        // Phase 1, every lane (safe indices for idle groups): read what the bind needs.
        const bool act = grp < Ws;
        const bool bindp = act && win != 0;
        const int tb = bindp ? tw : 0;
        const bool same = bindp && u == tb;                          // the node just Traded wins again (the common case)
        const uint32_t masks = !bindp ? 1u : same ? (1u << (bk & 7)) : (S.al[sgs][tb] & 0xFFu);
        const int gsel = __ffs(masks | 0x100u) - 1 & 7;
        const int cc = S.rc[tb][gsel], mm = S.rm[tb][gsel];
        const int okb = (cc >= rq_c && mm >= rq_m) ? 1 : 0;          // GPUs.Transact gpu.go:164-171
        const int nodeb = S.node[tb];
        const unsigned long long nfd = fd - S.fterm[tb], nsd = sd - score_term_b(S.sbase[tb], key_score(win));
        __syncwarp();                                                // all reads done before any leader writes
        // Phase 2, group leaders.
        if (act && gl == 0) {
          const unsigned ubit = 1u << (u & 31);
          int o_node = -1, o_status = EGS_ERR_NOFIT; uint32_t o_masks = 0;
          if (u >= 0 && !same) {                                     // this pod's filter Traded slot u
            if (bk >= 0) { S.st[sgs][u] = OPT_CACHED; S.al[sgs][u] = 1u << (bk & 7); S.tkey[sgs][u] = tradekey; }
            else S.st[sgs][u] = OPT_UNFIT;
            S.pmask[sgs][u >> 5] &= ~ubit;
          }
          if (bindp) {
            S.st[sgs][tb] = OPT_ABSENT; S.tkey[sgs][tb] = 0; S.pmask[sgs][tb >> 5] |= 1u << (tb & 31);   // node.go:90-92
            S.afit[sgs] = fit - 1; S.afd[sgs] = nfd; S.asd[sgs] = nsd; S.dirty[tb] = 1;
            if (okb) { S.rc[tb][gsel] = cc - rq_c; S.rm[tb][gsel] = mm - rq_m; }
            o_node = nodeb; o_status = okb ? EGS_OK : EGS_ERR_TRANSACT; o_masks = okb ? masks : 0;
          } else if (u >= 0 && bk >= 0) {                            // nothing fits elsewhere, but the Trade result stands
            S.afit[sgs] = fit; S.afd[sgs] = fd; S.asd[sgs] = sd;
          }
          const int r = (rel + grp) & 63;
          S.o_node[r] = o_node; S.o_status[r] = o_status; S.o_fit[r] = fit; S.o_fd[r] = fd; S.o_sd[r] = sd; S.o_alloc[r] = o_masks;
        }
        __syncwarp();
        // ---- the rest (a head-win, or pods sharing a node) in pod order
        if (Ws < Wc) {
          if (gl == 0 && grp >= Ws && grp < Wc) {
            PodRec r;
            r.s = sg; r.u = u; r.bk = bk; r.from_head = fh; r.t = tw; r.d = ownl; r.pad = sc;
            r.fit = fit; r.fd = fd; r.sd = sd; r.win = win;
            recs[grp] = r;
          }
          __syncwarp();
          for (int q = Ws; q < Wc; q++) {
            const PodRec r = recs[q];
            if (r.u >= 0 && lane == 0) {                             // this pod's filter Traded slot u
              if (r.bk >= 0) { S.st[r.s][r.u] = OPT_CACHED; S.al[r.s][r.u] = 1u << (r.bk & 7); S.tkey[r.s][r.u] = cand_key(r.pad, (uint32_t)S.node[r.u]); }
              else S.st[r.s][r.u] = OPT_UNFIT;
              S.pmask[r.s][r.u >> 5] &= ~(1u << (r.u & 31));
            }
            __syncwarp();
            commit_pod(S, a, lane, rel + q, r.s, r.win, r.from_head, r.t, r.d, r.fit, r.fd, r.sd, mono, ns, D, nT, n_observed, PROF_PTR);
          }
        }
        PROF_T(3) PROF_C(6, 1) PROF_C(7, Wc) PROF_C(8, W)
        p += Wc; done += Wc;
        continue;
      }
    }
    // ======== general path: one pod
    const int s = __shfl_sync(0xffffffffu, myshape, ri);
    if (s < 0) { reason = 1; break; }                            // shape outside this round's set
    if (nT >= RT) { reason = 2; break; }                          // no free tracked slot for a new winner
    // best untracked candidate per shard (cached heads; consumed entries were skipped when they were zeroed)
    unsigned long long head = 0; bool dry = false;
    if (lane < D) { head = S.hkey[s][lane]; dry = head == 0 && S.more[s][lane] != 0; }
    if (__ballot_sync(0xffffffffu, dry)) { reason = 3; break; }   // a truncated list ran dry: next round
    if (!S.observed[s]) {                                          // first pod of this shape in the round:
      for (int t = lane; t < nT; t += 32) if (S.st[s][t] == OPT_NEW) S.st[s][t] = OPT_CACHED;   // NEW options are now ordinary
      __syncwarp();
      if (lane == 0) S.observed[s] = 1;
      n_observed++;
      __syncwarp();
    }
    const int single = S.rq_single[s];
    // tracked nodes: Trade absent options NOW (this pod's filter); best tracked option
    unsigned long long best = 0; int best_t = -1;
    const int nw = (nT + 31) >> 5;
    for (int w = 0; w < nw; w++) {
      unsigned word = S.pmask[s][w];
      if (word) {
        if (single) {
          // 8 lanes per pending node (lane == GPU), up to 4 nodes at a time
          const int rq_c = S.rq_core[s], rq_m = S.rq_mem[s];
          while (word) {
            int bsel = -1;
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const int b = word ? __ffs(word) - 1 : -1;
              if (word) word &= word - 1;
              if (q == grp) bsel = b;
            }
            bool okl = false; int sc = 0; int t = 0;
            if (bsel >= 0) {                                       // group-uniform
              t = w * 32 + bsel;
              const int c = S.rc[t][gl], m = S.rm[t][gl];
              const int cmin = (int)__reduce_min_sync(gmask, (unsigned)c), mmin = (int)__reduce_min_sync(gmask, (unsigned)m);
              const int c1 = __reduce_max_sync(gmask, c), m1 = __reduce_max_sync(gmask, m);
              const int c2 = __reduce_max_sync(gmask, c == c1 ? INT32_MIN : c), m2 = __reduce_max_sync(gmask, m == m1 ? INT32_MIN : m);
              const bool cu = __popc(__ballot_sync(gmask, c == c1)) == 1, mu = __popc(__ballot_sync(gmask, m == m1)) == 1;
              const int cex = (c == c1 && cu) ? c2 : c1, mex = (m == m1 && mu) ? m2 : m1;   // max over the OTHER GPUs
              const bool ok = c >= rq_c && m >= rq_m;                                         // gpu.go:55
              const int nc = c - rq_c, nm = m - rq_m;
              const int x = (max(mex, nm) + max(cex, nc)) - (min(mmin, nm) + min(cmin, nc));
              const int key = !ok ? -1 : (a.policy == EGS_BINPACK ? (x >> 2) * 8 + gl : gl);
              const int bk = __reduce_max_sync(gmask, key);
              if (gl == 0) {
                if (bk >= 0) {
                  sc = a.policy == EGS_BINPACK ? (bk >> 3) * 100 : 0;
                  S.st[s][t] = OPT_CACHED; S.al[s][t] = 1u << (bk & 7); S.tkey[s][t] = cand_key(sc, (uint32_t)S.node[t]);
                  okl = true;
                } else {
                  S.st[s][t] = OPT_UNFIT;
                }
              }
            }
            for (unsigned rem = __ballot_sync(0xffffffffu, okl); rem; rem &= rem - 1) {   // usually one leader
              if (lane == __ffs(rem) - 1) { S.afit[s] += 1; S.afd[s] += S.fterm[t]; S.asd[s] += score_term_b(S.sbase[t], sc); }
              __syncwarp();
            }
          }
        } else {
          const int t = w * 32 + lane;
          bool ok = false; int sc = 0;
          if ((word >> lane) & 1u) {
            uint32_t masks;
            ok = trade_general(S.rc[t], S.rm[t], S.mt[t], S.reqs[s], a.policy, sc, masks);
            if (ok) { S.st[s][t] = OPT_CACHED; S.al[s][t] = masks; S.tkey[s][t] = cand_key(sc, (uint32_t)S.node[t]); }
            else S.st[s][t] = OPT_UNFIT;
          }
          for (unsigned rem = word; rem; rem &= rem - 1) {
            if (lane == __ffs(rem) - 1 && ok) { S.afit[s] += 1; S.afd[s] += S.fterm[t]; S.asd[s] += score_term_b(S.sbase[t], sc); }
            __syncwarp();
          }
        }
        __syncwarp();
        if (lane == 0) S.pmask[s][w] = 0;
      }
      const int t = w * 32 + lane;
      const unsigned long long k = t < nT ? S.tkey[s][t] : 0ull;
      if (k > best) { best = k; best_t = t; }
    }
    __syncwarp();
    // winner = max over (tracked options, untracked list heads): two redux.sync steps on the key halves
    const bool from_head = head > best;
    const unsigned long long mine = from_head ? head : best;
    const unsigned hi = (unsigned)(mine >> 32);
    const unsigned m1 = __reduce_max_sync(0xffffffffu, hi);
    const unsigned m2 = __reduce_max_sync(0xffffffffu, hi == m1 ? (unsigned)mine : 0u);
    const unsigned long long win = ((unsigned long long)m1 << 32) | m2;
    int owner = 0, fh = 0, tw = -1;
    if (win != 0) {
      owner = __ffs(__ballot_sync(0xffffffffu, mine == win)) - 1;
      fh = __shfl_sync(0xffffffffu, (int)from_head, owner);
      tw = __shfl_sync(0xffffffffu, best_t, owner);
    }
    PROF_T(4)
    commit_pod(S, a, lane, rel, s, win, fh, tw, owner, S.afit[s], S.afd[s], S.asd[s], mono, ns, D, nT, n_observed, PROF_PTR);
    PROF_T(5) PROF_C(9, 1)
    p++; done++;
  }
  __syncwarp();
  while (done > flushed) { const int n = min(32, done - flushed); flush_outputs(S, a.out, a.p0 + flushed, flushed, n, lane); flushed += n; }
  // ---- epilogue: write the tracked nodes back (each shard its own nodes)
  for (int t = 0; t < nT; t++) {
    const int w = S.node[t];
    if (w < a.lo || w >= a.hi) continue;
    if (S.dirty[t]) {
      if (lane < EGS_G) a.core[(size_t)w * EGS_G + lane] = S.rc[t][lane];
      else if (lane < 2 * EGS_G) a.mem[(size_t)w * EGS_G + lane - EGS_G] = S.rm[t][lane - EGS_G];
    }
    if (lane < ns) {
      const int slot = a.set.slot[lane];
      const uint8_t st = S.st[lane][t];
      tb_st(a.tb, slot)[w] = st;
      if (st == OPT_CACHED || st == OPT_NEW) {
        tb_sc(a.tb, slot)[w] = key_score(S.tkey[lane][t]);
        uint8_t *alp = tb_al(a.tb, slot);
        const uint32_t am = S.al[lane][t];
        for (int c = 0; c < S.reqs[lane].C; c++) alp[(size_t)c * a.tb.n_pad + w] = (uint8_t)(am >> (8 * c));
      }
    }
    if (S.dirty[t]) {                                            // shapes outside the round set
      for (int slot = lane; slot < a.tb.n_slots; slot += 32) {
        bool in_set = false;
        for (int q = 0; q < ns; q++) in_set |= a.set.slot[q] == slot;
        if (in_set) continue;
        uint8_t *q = tb_st(a.tb, slot) + w;
        if (*q == OPT_UNFIT) *q = OPT_ABSENT;
        else if (*q == OPT_NEW) *q = a.obs_pending[slot] ? OPT_CACHED : OPT_ABSENT;
      }
    }
  }
  if (lane < ns && S.observed[lane]) a.obs_pending[a.set.slot[lane]] = 1;
  if (lane == 0) { a.done[0] = done; a.done[1] = nT; a.done[2] = reason; }
#ifdef EGS_RESOLVE_PROF
  if (lane == 0) for (int i = 0; i < 16; i++) atomicAdd((unsigned long long *)a.prof + i, (unsigned long long)prof[i]);
#endif
}

// End of a ROUNDS batch: no OPT_NEW may outlive it (the other code paths know three states).
__global__ void k_rounds_finalize(TableSet tb, uint8_t *obs_pending, int lo, int hi) {
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  for (int slot = 0; slot < tb.n_slots; slot++) {
    uint8_t *q = tb_st(tb, slot) + i;
    if (*q == OPT_NEW) *q = obs_pending[slot] ? OPT_CACHED : OPT_ABSENT;
  }
}
__global__ void k_clear_u8(uint8_t *p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

// ---------------------------------------------------------------------------------------------
struct egs_handle;

struct RoundsState {
  void *comm = nullptr;             // ncclComm_t
  int32_t *d_pod_slot = nullptr; int pod_cap = 0;
  uint8_t *d_obs = nullptr; int obs_cap = 0;
  unsigned long long *d_cta_lists = nullptr; AggPart *d_cta_agg = nullptr; int grid = 0;
  ShardBuf *d_bufs = nullptr;       // [RD]; own shard written at index `rank`
  int32_t *d_done = nullptr; int32_t *h_done = nullptr; long long *d_prof = nullptr;
  int64_t rounds = 0, pods = 0, tracked = 0; int64_t stops[4] = {0, 0, 0, 0};
};

static int batch_rescan(egs_handle *h, int P, const int32_t *c_off, const egs_unit *units,
                        const std::vector<int> &slots, PodOut out);
static int batch_rounds(egs_handle *h, int P, const int32_t *c_off, const egs_unit *units,
                        const std::vector<int> &slots, PodOut out);
static void rounds_free(RoundsState *r);
static int rounds_comm_unique_id(uint8_t out_id[128]);
static int rounds_comm_init(egs_handle *h, const uint8_t id[128]);
