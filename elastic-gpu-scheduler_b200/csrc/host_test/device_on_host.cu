// device_on_host.cu -- the kernels' integer arithmetic (egs_device.cuh: Trade fast path, general Trade,
// Transact) compiled for the HOST, so the CPU test suite can check the very source the GPU runs against
// the oracle (tests/test_device_arith_host.py).  Not part of libegs.
#include <cstring>

#include "../egs_device.cuh"

static Req make_req(int C, const egs_unit *u) {
  Req r; memset(&r, 0, sizeof r);
  r.C = C;
  for (int i = 0; i < C; i++) { r.core[i] = u[i].core; r.mem[i] = u[i].mem; r.cnt[i] = u[i].count; }
  return r;
}
static void rows(const int32_t *core, const int32_t *mem, int (&c)[EGS_G], int (&m)[EGS_G]) {
  for (int g = 0; g < EGS_G; g++) { c[g] = core[g]; m[g] = mem[g]; }
}

extern "C" {
// path: 0 = the dispatch the kernels use (fast path when req_is_single), 1 = always the general DFS
int egsdh_trade(const int32_t *core, const int32_t *mem, int mem_total, int C, const egs_unit *units, int policy,
                int path, int32_t *score, uint32_t *masks) {
  int c[EGS_G], m[EGS_G];
  rows(core, mem, c, m);
  const Req r = make_req(C, units);
  const bool single = path == 0 && req_is_single(r);
  int sc = 0; uint32_t mk = 0;
  const bool ok = trade_any(c, m, mem_total, r, single, policy, sc, mk);
  *score = sc; *masks = mk;
  return ok ? 1 : 0;
}
// the leaf-parallel Trade the resolver runs across a warp, serially: max over (score, leaf index)
int egsdh_trade_leaves(const int32_t *core, const int32_t *mem, int mem_total, int C, const egs_unit *units, int policy,
                       int32_t *score, uint32_t *masks) {
  int c[EGS_G], m[EGS_G];
  rows(core, mem, c, m);
  const Req r = make_req(C, units);
  int bits, nbranch, nleaf;
  trade_leaf_space(c, r, bits, nbranch, nleaf);
  long long best = -1; uint32_t bm = 0;
  for (int leaf = 0; leaf < nleaf; leaf++) {
    uint32_t mk = 0;
    const int sc = trade_leaf_eval(c, m, mem_total, r, policy, bits, nbranch, leaf, mk);
    if (sc < 0) continue;
    const long long key = ((long long)sc << 20) | leaf;
    if (key > best) { best = key; bm = mk; }
  }
  if (best < 0) return 0;
  *score = (int32_t)(best >> 20); *masks = bm;
  return 1;
}
int egsdh_transact(int32_t *core, int32_t *mem, int mem_total, int C, const egs_unit *units, uint32_t masks) {
  const Req r = make_req(C, units);
  return transact_row(core, mem, mem_total, r, masks) ? 1 : 0;
}
int egsdh_is_single(int C, const egs_unit *units) { const Req r = make_req(C, units); return req_is_single(r) ? 1 : 0; }
unsigned long long egsdh_cand_key(int32_t score, uint32_t node) { return cand_key(score, node); }
unsigned long long egsdh_fit_term(uint32_t node) { return fit_term(node); }
unsigned long long egsdh_score_term(uint32_t node, int32_t score) { return score_term(node, score); }
}
