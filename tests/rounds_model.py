"""Executable model of EGS_MODE_ROUNDS (csrc/egs_rounds.cuh) in plain Python.

It follows the device algorithm step for step -- option states ABSENT/CACHED/UNFIT/NEW, `k_select`
evaluating absent options ahead of time, top-K candidate lists per (shard, shape), the sequential
resolver with its tracked-node table, pending re-evaluation, invalidation rules, the monotone
shortcut, sticky observation flags and the end-of-batch finalize -- but with K, T, RS as parameters so
tests can force every early-termination path.  Trade / Transact come from the python mirror of the
reference (oracle/egs_oracle.py), so the model checks the ROUND STRUCTURE, not the arithmetic.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import egs_oracle as po

ABSENT, CACHED, UNFIT, NEW = 0, 1, 2, 3
MASK64 = (1 << 64) - 1


def fit_term(node: int) -> int:
    return po.fit_digest_term(node)


def score_term(node: int, score: int) -> int:
    return po.score_digest_term(node, score)


class Entry:
    __slots__ = ("st", "score", "alloc")

    def __init__(self):
        self.st, self.score, self.alloc = ABSENT, 0, None


class RoundsModel:
    def __init__(self, policy: int, K: int = 32, T: int = 256, RS: int = 32, shards: int = 1, window: int = 1):
        self.policy, self.K, self.T, self.RS, self.D, self.W = policy, K, T, RS, shards, window
        self.nodes: List[List[po.GPU]] = []
        self.tables: Dict[tuple, List[Entry]] = {}          # shape -> per-node entries
        self.obs_pending: Dict[tuple, bool] = {}
        self.stats = dict(rounds=0, dry=0, full=0, shape=0, windows=0, window_pods=0, cuts=0)

    # ---- state
    def add_node(self, core_alloc: int, mem_alloc: int) -> int:
        na = po.NodeAllocator.new(core_alloc, mem_alloc, self.policy)
        self.nodes.append(na.gpus if na else [])
        for t in self.tables.values():
            t.append(Entry())
        return len(self.nodes) - 1

    def set_rows(self, node: int, core, mem):
        for g, c, m in zip(self.nodes[node], core, mem):
            g.core_avail, g.mem_avail = c, m

    def rows(self, node: int):
        return [(g.core_avail, g.mem_avail) for g in self.nodes[node]]

    def _table(self, shape):
        if shape not in self.tables:
            self.tables[shape] = [Entry() for _ in self.nodes]
            self.obs_pending[shape] = False
        return self.tables[shape]

    def _trade(self, gpus, shape):
        if not gpus:
            return None
        return po.trade(gpus, po.RATERS[self.policy], list(shape))

    # ---- one batch, driver rule
    def schedule_batch(self, pods: Sequence[tuple]):
        out = []
        p0, P = 0, len(pods)
        N = len(self.nodes)
        bounds = [N * d // self.D for d in range(self.D + 1)]
        while p0 < P:
            # round set: distinct shapes in pod order
            shapes: List[tuple] = []
            plim = p0
            while plim < P:
                s = pods[plim]
                if s not in shapes:
                    if len(shapes) == self.RS:
                        break
                    shapes.append(s)
                plim += 1
            for s in shapes:
                self._table(s)
            mono = all(u[0] >= 0 and u[1] >= 0 for s in shapes for u in s)
            # ---- k_select per shard: evaluate ABSENT, convert NEW when observed since, aggregates, top-K
            lists = {s: [] for s in shapes}                      # per shape: per shard (keys, more)
            agg = {}
            for s in shapes:
                tab = self.tables[s]
                pend = self.obs_pending[s]
                fit, fd, sd = 0, 0, 0
                for d in range(self.D):
                    keys = []
                    for i in range(bounds[d], bounds[d + 1]):
                        e = tab[i]
                        if e.st == NEW and pend:
                            e.st = CACHED
                        if e.st == ABSENT:
                            opt = self._trade(self.nodes[i], s)
                            if opt is None:
                                e.st = UNFIT
                            else:
                                e.st, e.score, e.alloc = NEW, opt.score, opt.allocated
                        if e.st in (CACHED, NEW):
                            keys.append((-e.score, i))
                            fit += 1
                            fd = (fd + fit_term(i)) & MASK64
                            sd = (sd + score_term(i, e.score)) & MASK64
                    keys.sort()
                    lists[s].append((keys[:self.K], len(keys) > self.K))
                agg[s] = [fit, fd, sd]
                self.obs_pending[s] = False                       # consumed by this select (k_merge)
            # ---- k_resolve
            self.stats["rounds"] += 1
            tracked: Dict[int, dict] = {}                         # node -> {shape: [st, score, alloc]}, rows live in self.nodes copy
            rows_copy: Dict[int, List[po.GPU]] = {}
            dirty = set()
            observed = {s: False for s in shapes}
            cur = {s: [0] * self.D for s in shapes}
            done = 0
            p = p0
            while p < plim:
                # ---- 4-wide window (device fast path): mono round, every shape observed, distinct shapes,
                # <= 1 pending option per shape, room in the tracked table.  Decisions are taken from the SAME
                # state, then cut at the first hazard, then committed in order WITHOUT being recomputed.
                if self.W > 1 and mono and all(observed.values()) and len(tracked) + self.W <= self.T:
                    win_shapes = []
                    for q in range(self.W):
                        if p + q >= plim or pods[p + q] in win_shapes or pods[p + q] not in lists:
                            break
                        win_shapes.append(pods[p + q])
                    decs = []
                    for s in win_shapes:
                        pend = [n_ for n_, ent in tracked.items() if ent[s][0] == ABSENT]
                        heads, dry = [], False
                        for d in range(self.D):
                            keys, more = lists[s][d]
                            c = cur[s][d]
                            while c < len(keys) and keys[c][1] in tracked:
                                c += 1
                            if c < len(keys):
                                heads.append(keys[c])
                            elif more:
                                dry = True
                        if len(pend) > 1 or dry:
                            break
                        u = pend[0] if pend else None
                        topt = self._trade(rows_copy[u], s) if u is not None else None
                        cands = [(-ent[s][1], n_) for n_, ent in tracked.items() if ent[s][0] in (CACHED, NEW)] + heads
                        if topt is not None:
                            cands.append((-topt.score, u))
                        fitc, ofd, osd = agg[s]
                        if topt is not None:
                            fitc, ofd, osd = fitc + 1, (ofd + fit_term(u)) & MASK64, (osd + score_term(u, topt.score)) & MASK64
                        w = min(cands)[1] if cands else None
                        decs.append(dict(s=s, u=u, topt=topt, w=w, head=(w is not None and w not in tracked), agg=(fitc, ofd, osd)))
                    nW = len(decs)
                    for j in range(1, nW):                        # hazards
                        if any((decs[i]["head"] and not getattr(self, "relax_head", False)) or
                               (decs[j]["u"] is not None and decs[j]["u"] == decs[i]["w"] and not decs[i]["head"]) for i in range(j)):
                            nW = j
                            self.stats["cuts"] += 1
                            break
                    if nW >= 2:
                        self.stats["windows"] += 1
                        self.stats["window_pods"] += nW
                        for dec in decs[:nW]:
                            s, u, topt, w = dec["s"], dec["u"], dec["topt"], dec["w"]
                            fitc, ofd, osd = dec["agg"]
                            if u is not None:                     # the Trade result of this pod's filter
                                e = tracked[u][s]
                                if topt is None:
                                    e[0] = UNFIT
                                else:
                                    e[0], e[1], e[2] = CACHED, topt.score, topt.allocated
                            agg[s] = [fitc, ofd, osd]
                            if w is None:
                                out.append(dict(node=-1, status=po.EGS_ERR_NOFIT, alloc=None, fit_count=fitc, fit_digest=ofd, score_digest=osd))
                            else:
                                if w not in tracked:
                                    rows_copy[w] = self.nodes[w]
                                    tracked[w] = {s2: [CACHED if (self.tables[s2][w].st == NEW) else self.tables[s2][w].st,
                                                       self.tables[s2][w].score, self.tables[s2][w].alloc] for s2 in shapes}
                                e = tracked[w][s]
                                opt = po.GPUOption(request=list(s), allocated=e[2], score=e[1])
                                e[0] = ABSENT
                                agg[s] = [fitc - 1, (ofd - fit_term(w)) & MASK64, (osd - score_term(w, e[1])) & MASK64]
                                ok = po.transact(rows_copy[w], opt)
                                dirty.add(w)
                                out.append(dict(node=w, status=po.EGS_OK if ok else po.EGS_ERR_TRANSACT, alloc=opt.allocated if ok else None,
                                                fit_count=fitc, fit_digest=ofd, score_digest=osd))
                            p += 1
                            done += 1
                        continue
                s = pods[p]
                if s not in lists:
                    self.stats["shape"] += 1
                    break
                if len(tracked) >= self.T:
                    self.stats["full"] += 1
                    break
                # heads (skip candidates that became tracked)
                heads = []
                dry = False
                for d in range(self.D):
                    keys, more = lists[s][d]
                    c = cur[s][d]
                    while c < len(keys) and keys[c][1] in tracked:
                        c += 1
                    cur[s][d] = c
                    if c < len(keys):
                        heads.append(keys[c])
                    elif more:
                        dry = True
                if dry:
                    self.stats["dry"] += 1
                    break
                if not observed[s]:
                    for n_, ent in tracked.items():
                        if ent[s][0] == NEW:
                            ent[s][0] = CACHED
                    observed[s] = True
                # tracked nodes: Trade absent options now
                for n_, ent in tracked.items():
                    e = ent[s]
                    if e[0] == ABSENT:
                        opt = self._trade(rows_copy[n_], s)
                        if opt is None:
                            e[0] = UNFIT
                        else:
                            e[0], e[1], e[2] = CACHED, opt.score, opt.allocated
                            agg[s][0] += 1
                            agg[s][1] = (agg[s][1] + fit_term(n_)) & MASK64
                            agg[s][2] = (agg[s][2] + score_term(n_, opt.score)) & MASK64
                cands = [(-ent[s][1], n_) for n_, ent in tracked.items() if ent[s][0] in (CACHED, NEW)] + heads
                fitc, ofd, osd = agg[s]
                if not cands:
                    out.append(dict(node=-1, status=po.EGS_ERR_NOFIT, alloc=None, fit_count=fitc, fit_digest=ofd, score_digest=osd))
                    p += 1
                    done += 1
                    continue
                negs, w = min(cands)
                if w not in tracked:                              # head-win: the node becomes tracked
                    rows_copy[w] = self.nodes[w]                  # (the model mutates the node rows in place)
                    ent = {}
                    for s2 in shapes:
                        e = self.tables[s2][w]
                        st = e.st
                        if st == NEW and observed[s2]:
                            st = CACHED
                        ent[s2] = [st, e.score, e.alloc]
                    tracked[w] = ent
                e = tracked[w][s]
                opt = po.GPUOption(request=list(s), allocated=e[2], score=e[1])
                # deferred delete + aggregates, Transact
                e[0] = ABSENT
                agg[s][0] -= 1
                agg[s][1] = (agg[s][1] - fit_term(w)) & MASK64
                agg[s][2] = (agg[s][2] - score_term(w, e[1])) & MASK64
                ok = po.transact(rows_copy[w], opt)
                dirty.add(w)
                all_obs = all(observed.values())
                if not mono or not all_obs:
                    for s2 in shapes:
                        if s2 == s:
                            continue
                        e2 = tracked[w][s2]
                        if e2[0] == UNFIT and not mono:
                            e2[0] = ABSENT
                        elif e2[0] == NEW and not observed[s2]:
                            e2[0] = ABSENT
                            agg[s2][0] -= 1
                            agg[s2][1] = (agg[s2][1] - fit_term(w)) & MASK64
                            agg[s2][2] = (agg[s2][2] - score_term(w, e2[1])) & MASK64
                out.append(dict(node=w, status=po.EGS_OK if ok else po.EGS_ERR_TRANSACT, alloc=opt.allocated if ok else None,
                                fit_count=fitc, fit_digest=ofd, score_digest=osd))
                p += 1
                done += 1
            assert done >= 1, "resolver made no progress"
            if p == plim and plim < P:
                self.stats["shape"] += 1                          # the next pod's shape is outside this round's set
            # ---- epilogue: write back
            for w, ent in tracked.items():
                for s2 in shapes:
                    e = self.tables[s2][w]
                    e.st, e.score, e.alloc = ent[s2]
                if w in dirty:
                    for s2, tab in self.tables.items():
                        if s2 in shapes:
                            continue
                        e = tab[w]
                        if e.st == UNFIT:
                            e.st = ABSENT
                        elif e.st == NEW:
                            e.st = CACHED if self.obs_pending[s2] else ABSENT
            for s2 in shapes:
                if observed[s2]:
                    self.obs_pending[s2] = True
            p0 += done
        # ---- finalize: no NEW outlives the batch
        for s2, tab in self.tables.items():
            for e in tab:
                if e.st == NEW:
                    e.st = CACHED if self.obs_pending[s2] else ABSENT
            self.obs_pending[s2] = False
        return out
