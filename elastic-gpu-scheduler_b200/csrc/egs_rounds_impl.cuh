// egs_rounds_impl.cuh -- host side of EGS_MODE_ROUNDS (included at the end of egs_api.cu).
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is resolved at run time (see nccl_api)

// NCCL is NOT a link-time dependency: a process that also hosts PyTorch must end up with ONE
// libnccl (torch bundles its own, newer than the system one).  dlopen by soname returns the
// copy that is already mapped, else the system library.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return nullptr;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))dlsym(lib, "ncclAllGather");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  return api.ok ? &api : nullptr;
}

#define NCK(h, call)                                                                     \
  do {                                                                                   \
    ncclResult_t r_ = (call);                                                            \
    if (r_ != ncclSuccess) {                                                             \
      (h)->err = std::string(#call) + ": " + nccl_api()->GetErrorString(r_);                     \
      return EGS_ERR_COMM;                                                               \
    }                                                                                    \
  } while (0)

static void rounds_free(RoundsState *r) {
  if (r->comm && nccl_api()) nccl_api()->CommDestroy((ncclComm_t)r->comm);
  void *dev[] = {r->d_pod_slot, r->d_obs, r->d_cta_lists, r->d_cta_agg, r->d_bufs, r->d_done, r->d_prof};
  for (void *p : dev) if (p) cudaFree(p);
  if (r->h_done) cudaFreeHost(r->h_done);
  *r = RoundsState();
}

static int rounds_comm_unique_id(uint8_t out_id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  if (!nccl_api() || nccl_api()->GetUniqueId(&id) != ncclSuccess) return EGS_ERR_COMM;
  memcpy(out_id, &id, 128);
  return EGS_OK;
}

static int rounds_comm_init(egs_handle *h, const uint8_t id[128]) {
  if (h->world < 2) return fail(h, EGS_ERR_BAD_ARG, "egs_shard_set first");
  if (!nccl_api()) return fail(h, EGS_ERR_COMM, "libnccl.so.2 not found");
  if (h->rounds.comm) { nccl_api()->CommDestroy((ncclComm_t)h->rounds.comm); h->rounds.comm = nullptr; }
  ncclUniqueId uid; memcpy(&uid, id, 128);
  ncclComm_t comm;
  NCK(h, nccl_api()->CommInitRank(&comm, h->world, uid, h->rank));
  h->rounds.comm = comm;
  return EGS_OK;
}

static int rounds_ensure(egs_handle *h, int P) {
  RoundsState &R = h->rounds;
  if (!R.d_bufs) {
    CK(h, cudaMalloc(&R.d_bufs, sizeof(ShardBuf) * RD));
    CK(h, cudaMemsetAsync(R.d_bufs, 0, sizeof(ShardBuf) * RD, h->stream));
    CK(h, cudaMalloc(&R.d_done, sizeof(int32_t) * 4));
    CK(h, cudaMalloc(&R.d_prof, sizeof(long long) * 16));
    CK(h, cudaMemsetAsync(R.d_prof, 0, sizeof(long long) * 16, h->stream));
    CK(h, cudaMallocHost(&R.h_done, sizeof(int32_t) * 4));
    CK(h, cudaFuncSetAttribute(k_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ResolveSmem)));
  }
  const int chunks = (h->hi - h->lo + 127) / 128;
  const int grid = std::max(1, std::min((chunks + SEL_WARPS - 1) / SEL_WARPS, 296));
  if (grid > R.grid) {
    if (R.d_cta_lists) { cudaFree(R.d_cta_lists); cudaFree(R.d_cta_agg); }
    CK(h, cudaMalloc(&R.d_cta_lists, sizeof(unsigned long long) * (size_t)grid * RS * RK));
    CK(h, cudaMalloc(&R.d_cta_agg, sizeof(AggPart) * (size_t)grid * RS));
    R.grid = grid;
  }
  if (P > R.pod_cap) {
    if (R.d_pod_slot) cudaFree(R.d_pod_slot);
    R.pod_cap = 0;
    CK(h, cudaMalloc(&R.d_pod_slot, sizeof(int32_t) * (size_t)P));
    R.pod_cap = P;
  }
  const int ns = (int)h->shapes.size();
  if (ns > R.obs_cap) {
    const int cap = std::max(ns, 4096);
    uint8_t *n;
    CK(h, cudaMalloc(&n, (size_t)cap));
    CK(h, cudaMemsetAsync(n, 0, (size_t)cap, h->stream));
    if (R.d_obs) {
      CK(h, cudaMemcpyAsync(n, R.d_obs, (size_t)R.obs_cap, cudaMemcpyDeviceToDevice, h->stream));
      CK(h, cudaStreamSynchronize(h->stream));
      cudaFree(R.d_obs);
    }
    R.d_obs = n; R.obs_cap = cap;
  }
  return EGS_OK;
}

static int batch_rounds(egs_handle *h, int P, const int32_t *c_off, const egs_unit *units,
                        const std::vector<int> &slots, PodOut out) {
  (void)c_off; (void)units;
  RoundsState &R = h->rounds;
  if (h->world > 1 && !R.comm) return fail(h, EGS_ERR_COMM, "sharded handle without egs_comm_init");
  if (h->world > RD) return fail(h, EGS_ERR_BAD_ARG, "too many shards");
  TRY(rounds_ensure(h, P));
  // pod -> slot ids to the device (pinned staging)
  TRY(ensure_stage(h, sizeof(int32_t) * (size_t)P));
  CK(h, cudaStreamSynchronize(h->stream));
  memcpy(h->h_stage, slots.data(), sizeof(int32_t) * (size_t)P);
  CK(h, cudaMemcpyAsync(R.d_pod_slot, h->h_stage, sizeof(int32_t) * (size_t)P, cudaMemcpyHostToDevice, h->stream));

  TableSet tb; tb.st = h->d_st; tb.sc = h->d_sc; tb.al = h->d_al; tb.n_pad = (size_t)h->n_pad; tb.n_slots = (int)h->shapes.size();
  const int chunks = (h->hi - h->lo + 127) / 128;
  const int grid = std::max(1, std::min((chunks + SEL_WARPS - 1) / SEL_WARPS, 296));
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (h->timing) for (auto &e : ev) CK(h, cudaEventCreate(&e));

  // distinct shapes of the whole batch, in order of first appearance: when they fit one round set the set
  // is the same for every round and no per-round scan of the pod list is needed
  std::vector<int> batch_shapes;
  {
    std::vector<char> seen(h->shapes.size(), 0);
    for (int p = 0; p < P && (int)batch_shapes.size() <= RS; p++)
      if (!seen[slots[p]]) { seen[slots[p]] = 1; batch_shapes.push_back(slots[p]); }
  }
  const bool one_set = (int)batch_shapes.size() <= RS;

  // Cold shapes (option table still all-absent, e.g. a fresh or restored scheduler): ONE full-evaluate launch
  // per shape over this shard's nodes fills the table (OPT_NEW / OPT_UNFIT) -- the HBM-roofline kernel,
  // N * (8G + 5 + C) bytes each; k_select then finds nothing left to Trade for them.
  if (h->hi > h->lo) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int cold = 0;
    for (int slot : batch_shapes) {
      if ((int)batch_shapes.size() > RS && cold >= RS) break;
      if (!h->slot_cold[slot]) continue;
      if (!e0) { CK(h, cudaEventCreate(&e0)); CK(h, cudaEventCreate(&e1)); CK(h, cudaEventRecord(e0, h->stream)); }
      const Shape &sh = h->shapes[slot];
      EvalArgs ea;
      ea.core = h->d_core; ea.mem = h->d_mem; ea.mem_total = h->d_mem_total; ea.lo = h->lo; ea.n = h->hi - h->lo;
      ea.policy = h->policy; ea.req = make_req(sh.C, sh.u);
      ea.fit = tb.st + (size_t)slot * tb.n_pad; ea.score = tb.sc + (size_t)slot * tb.n_pad;
      ea.gpu = tb.al + (size_t)slot * EGS_C * tb.n_pad; ea.plane = tb.n_pad; ea.v_fit = OPT_NEW; ea.v_unfit = OPT_UNFIT;
      if (is_single(sh.C, sh.u)) k_evaluate<true, 2><<<(ea.n + 511) / 512, 256, 0, h->stream>>>(ea);
      else k_evaluate<false, 1><<<(ea.n + 255) / 256, 256, 0, h->stream>>>(ea);
      cold++;
    }
    if (e0) {
      CK(h, cudaEventRecord(e1, h->stream));
      CK(h, cudaEventSynchronize(e1));
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      h->k_launches[EGS_K_EVALUATE] += cold; h->k_ms[EGS_K_EVALUATE] += ms;
      cudaEventDestroy(e0); cudaEventDestroy(e1);
      CK(h, cudaGetLastError());
    }
  }
  for (int slot : batch_shapes) h->slot_cold[slot] = 0;
  for (int p = 0; p < P; p++) h->slot_cold[slots[p]] = 0;

  int p0 = 0;
  while (p0 < P) {
    // the round's shape set: distinct shapes in pod order until RS are collected
    SelectArgs sa; MergeArgs ma; ResolveArgs ra;
    RoundSet set; set.n = 0;
    int plim = p0;
    const int pcap = std::min(P, p0 + 16384);    // a round never gets further (tracked table, list depth)
    if (one_set) {
      for (int q : batch_shapes) set.slot[set.n++] = q;
      plim = pcap;
    }
    for (; plim < pcap; plim++) {
      const int slot = slots[plim];
      bool found = false;
      for (int q = 0; q < set.n; q++) if (set.slot[q] == slot) { found = true; break; }
      if (found) continue;
      if (set.n == RS) break;
      set.slot[set.n++] = slot;
    }
    for (int q = set.n; q < RS; q++) set.slot[q] = -1;
    sa.core = h->d_core; sa.mem = h->d_mem; sa.mem_total = h->d_mem_total;
    sa.lo = h->lo; sa.hi = h->hi; sa.policy = h->policy; sa.set = set; sa.tb = tb; sa.obs_pending = R.d_obs;
    sa.cta_lists = R.d_cta_lists; sa.cta_agg = R.d_cta_agg;
    memset(sa.reqs, 0, sizeof sa.reqs);
    for (int q = 0; q < set.n; q++) sa.reqs[q] = make_req(h->shapes[set.slot[q]].C, h->shapes[set.slot[q]].u);
    ma.core = h->d_core; ma.mem = h->d_mem; ma.mem_total = h->d_mem_total; ma.set = set; ma.tb = tb;
    ma.obs_pending = R.d_obs; ma.cta_lists = R.d_cta_lists; ma.cta_agg = R.d_cta_agg; ma.n_cta = grid;
    ma.out = R.d_bufs + h->rank;
    ra.core = h->d_core; ra.mem = h->d_mem; ra.lo = h->lo; ra.hi = h->hi; ra.policy = h->policy; ra.n_shards = h->world;
    ra.set = set; memcpy(ra.reqs, sa.reqs, sizeof ra.reqs); ra.tb = tb; ra.obs_pending = R.d_obs; ra.bufs = R.d_bufs;
    ra.pod_slot = R.d_pod_slot; ra.p0 = p0; ra.p_limit = plim; ra.out = out; ra.done = R.d_done; ra.prof = R.d_prof;

    if (h->timing) CK(h, cudaEventRecord(ev[0], h->stream));
    k_select<<<grid, SEL_THREADS, 0, h->stream>>>(sa);
    if (h->timing) CK(h, cudaEventRecord(ev[1], h->stream));
    k_merge<<<set.n, 256, 0, h->stream>>>(ma);
    if (h->world > 1)
      NCK(h, nccl_api()->AllGather(R.d_bufs + h->rank, R.d_bufs, sizeof(ShardBuf), ncclChar, (ncclComm_t)R.comm, h->stream));
    if (h->timing) CK(h, cudaEventRecord(ev[2], h->stream));
    k_resolve<<<1, 32, sizeof(ResolveSmem), h->stream>>>(ra);
    if (h->timing) CK(h, cudaEventRecord(ev[3], h->stream));
    CK(h, cudaMemcpyAsync(R.h_done, R.d_done, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    const int done = R.h_done[0];
    if (done < 1 || done > plim - p0) return fail(h, EGS_ERR_CUDA, "rounds: resolver made no progress");
    p0 += done;
    R.rounds++; R.pods += done; R.tracked += R.h_done[1]; R.stops[R.h_done[2] & 3]++;
    h->k_launches[EGS_K_SELECT] += 2; h->k_launches[EGS_K_RESOLVE] += 1;
    if (h->timing) {
      float a = 0, b = 0, c = 0;
      cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&c, ev[2], ev[3]);
      h->k_ms[EGS_K_SELECT] += a; h->k_ms[EGS_K_MERGE] += b; h->k_ms[EGS_K_RESOLVE] += c;
    }
  }
  if (h->timing) for (auto &e : ev) cudaEventDestroy(e);
  const int n = h->hi - h->lo;
  if (n > 0) k_rounds_finalize<<<(n + 255) / 256, 256, 0, h->stream>>>(tb, R.d_obs, h->lo, h->hi);   // a shard may be empty
  k_clear_u8<<<(R.obs_cap + 255) / 256, 256, 0, h->stream>>>(R.d_obs, R.obs_cap);
  CK(h, cudaGetLastError());
  return EGS_OK;
}
