"""CPU ORACLE (python mirror) -- TEST INFRASTRUCTURE ONLY, never on the product path.

A line-by-line restatement, in plain Python integers, of the reference's
Filter / Score / Allocate path (`pkg/scheduler` of elastic-ai/elastic-gpu-scheduler).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.

PARITY STATUS: "parity unpinned" by the reference itself -- its only test
(`pkg/scheduler/scheduler_test.go:11-24`) has no assertions and no Go toolchain
exists in this image, so the pins are the hand-traced known-answer vectors of
SURVEY.md section 8c (KA-0..KA-10, KA-T, hash KATs), checked in
tests/test_oracle_ka.py, plus the sha256 KATs of R8 which ARE checkable
(hashlib == Go crypto/sha256).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Nothing here is copied: Go pointers/slices/maps are restated
as Python lists/dicts.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

GPU_CORE_EACH_CARD = 100  # pkg/utils/types.go:6
NOT_NEED_GPU = -1         # pkg/scheduler/allocate.go:16
NOT_NEED_RATE = -2        # pkg/scheduler/allocate.go:17
SCORE_MIN = 0             # pkg/scheduler/rater.go:4

POLICY_BINPACK = 0
POLICY_SPREAD = 1

# status codes shared with include/egs.h
EGS_OK = 0
EGS_ERR_NOFIT = 1          # "no enough resource to allocate"   gpu.go:126
EGS_ERR_NO_OPTION = 2      # "cannot find option of GPU request" node.go:95
EGS_ERR_TRANSACT = 3       # "can't trade option ..."            gpu.go:160,168
EGS_ERR_BAD_ARG = 4
EGS_ERR_NO_GPU = 7         # "no gpu available on node %s"      node.go:29

MSG_NOFIT = "no enough resource to allocate"  # gpu.go:126

Unit = Tuple[int, int, int]  # (Core, Memory, GPUCount)  gpu.go:9-13


def unit_string(u: Unit) -> str:
    """GPUUnit.String, gpu.go:15-17."""
    return "(core: %d, memory: %d, gpu count: %d)" % (u[0], u[1], u[2])


def request_string(req: Sequence[Unit]) -> str:
    """GPURequest.String, allocate.go:22-28."""
    return "".join(unit_string(u) for u in req)


def request_hash(req: Sequence[Unit]) -> str:
    """GPURequest.Hash, allocate.go:30-33: first 8 hex chars of sha256."""
    return hashlib.sha256(request_string(req).encode()).hexdigest()[0:8]


def new_gpu_request(containers: Sequence[Tuple[int, int]]) -> List[Unit]:
    """NewGPURequest, allocate.go:35-58.  `containers` = per container
    (Requests[gpu-core], Requests[gpu-memory]) already through Quantity.Value()
    (pod.go:94-108; 0 when the key is absent)."""
    out: List[Unit] = []
    for core, mem in containers:
        if core == 0 and mem == 0:                       # allocate.go:41-45
            out.append((NOT_NEED_GPU, NOT_NEED_GPU, 0))
        elif core >= GPU_CORE_EACH_CARD:                 # allocate.go:46-49
            out.append((0, 0, core // GPU_CORE_EACH_CARD))
        else:                                            # allocate.go:50-53
            out.append((core, mem, 0))
    return out


@dataclass
class GPU:
    """gpu.go:19-25."""
    core_avail: int
    mem_avail: int
    core_total: int
    mem_total: int

    def add(self, u: Unit) -> None:          # gpu.go:31-39
        if u[2] > 0:
            self.core_avail = 0
            self.mem_avail = 0
        else:
            self.core_avail -= u[0]
            self.mem_avail -= u[1]

    def sub(self, u: Unit) -> None:          # gpu.go:41-49
        if u[2] > 0:
            self.core_avail = self.core_total
            self.mem_avail = self.mem_total
        else:
            self.core_avail += u[0]
            self.mem_avail += u[1]

    def can_allocate(self, u: Unit) -> bool:  # gpu.go:51-56
        if u[2] > 0:
            return self.core_avail == self.core_total and self.mem_avail == self.mem_total
        return self.core_avail >= u[0] and self.mem_avail >= u[1]


def go_div(a: int, b: int) -> int:
    """Go integer division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def rate_binpack(g: Sequence[GPU], indexes: Sequence[int]) -> int:
    """Binpack.Rate, rater.go:18-51."""
    seen = [0] * len(g)
    gpu_count = 0
    for i in indexes:
        if i < 0:
            continue
        if seen[i] == 0:
            seen[i] += 1
            gpu_count += 1
    max_m = min_m = g[0].mem_avail
    max_c = min_c = g[0].core_avail
    for x in g:
        max_m = max(max_m, x.mem_avail)
        min_m = min(min_m, x.mem_avail)
        max_c = max(max_c, x.core_avail)
        min_c = min(min_c, x.core_avail)
    rng = go_div(max_m + max_c - min_m - min_c, 2)
    return go_div(rng, gpu_count + 1) * 100


def rate_spread(g: Sequence[GPU], indexes: Sequence[int]) -> int:
    """Spread.Rate, rater.go:56-59 (a stub in the reference: constant 0)."""
    return 0


RATERS = {POLICY_BINPACK: rate_binpack, POLICY_SPREAD: rate_spread}


@dataclass
class GPUOption:
    """allocate.go:60-73."""
    request: List[Unit]
    allocated: List[Optional[List[int]]]
    score: int = 0


def get_free_gpus(g: Sequence[GPU]) -> List[int]:
    """GPUs.GetFreeGPUs, gpu.go:193-202."""
    return [i for i, x in enumerate(g)
            if x.core_avail == x.core_total and x.mem_avail == x.mem_total]


def trade(g: List[GPU], rater, request: List[Unit]) -> Optional[GPUOption]:
    """GPUs.Trade, gpu.go:65-129.  Returns None for the error
    "no enough resource to allocate" (gpu.go:125-127)."""
    indexes: List[Optional[List[int]]] = [None] * len(request)
    option = GPUOption(request=list(request), allocated=[None] * len(request), score=0)
    found = False

    def dfs(ci: int) -> None:
        nonlocal found
        if ci == len(request):                                    # gpu.go:73-93
            found = True
            rate_idx = [ix[0] if len(ix) == 1 else NOT_NEED_RATE for ix in indexes]
            s = rater(g, rate_idx)
            if option.score > s:                                  # gpu.go:85
                return
            option.allocated = [list(ix) for ix in indexes]
            option.score = s
            return
        u = request[ci]
        if u[2] > 0:                                              # gpu.go:95-109
            free = get_free_gpus(g)
            if len(free) < u[2]:
                return
            indexes[ci] = free[:u[2]]
            for gi in indexes[ci]:
                g[gi].add(u)
            dfs(ci + 1)
            for gi in indexes[ci]:
                g[gi].sub(u)
            return
        for i, gpu in enumerate(g):                               # gpu.go:110-122
            if not gpu.can_allocate(u):
                continue
            gpu.add(u)
            indexes[ci] = [i]
            dfs(ci + 1)
            gpu.sub(u)

    dfs(0)
    if not found:
        return None
    return option


def transact(g: List[GPU], option: GPUOption) -> bool:
    """GPUs.Transact, gpu.go:153-175.  False == error; earlier Adds stay applied."""
    for i in range(len(option.allocated)):
        alloc = option.allocated[i] or []
        u = option.request[i]
        if u[2] > 0:
            for j in alloc:
                if not g[j].can_allocate(u):
                    return False
                g[j].add(u)
        else:
            if len(alloc) > 0:
                if not g[alloc[0]].can_allocate(u):
                    return False
                g[alloc[0]].add(u)
    return True


def cancel(g: List[GPU], option: GPUOption) -> None:
    """GPUs.Cancel, gpu.go:177-191 (unchecked)."""
    for i in range(len(option.request)):
        alloc = option.allocated[i] or []
        u = option.request[i]
        if u[2] > 0:
            for j in alloc:
                g[j].sub(u)
        else:
            if len(alloc) > 0:
                g[alloc[0]].sub(u)


@dataclass
class NodeAllocator:
    """node.go:13-21."""
    gpus: List[GPU]
    policy: int
    allocated: Dict[object, GPUOption] = field(default_factory=dict)
    pods_map: Dict[int, bool] = field(default_factory=dict)
    faithful_hash: bool = False

    @staticmethod
    def new(core_allocatable: int, mem_allocatable: int, policy: int,
            faithful_hash: bool = False) -> Optional["NodeAllocator"]:
        """NewNodeAllocator, node.go:23-59 (without pod replay: callers use add())."""
        gpu_count = core_allocatable // GPU_CORE_EACH_CARD           # node.go:27
        if gpu_count == 0:                                            # node.go:28-30
            return None
        m = mem_allocatable // gpu_count                              # node.go:37-38
        gpus = [GPU(GPU_CORE_EACH_CARD, m, GPU_CORE_EACH_CARD, m) for _ in range(gpu_count)]
        return NodeAllocator(gpus=gpus, policy=policy, faithful_hash=faithful_hash)

    def key(self, req: Sequence[Unit]):
        # R8: the reference keys on the sha256 prefix; the tuple differs only on
        # a 32-bit prefix collision.
        return request_hash(req) if self.faithful_hash else tuple(req)

    def assume(self, req: List[Unit]) -> Optional[List[List[int]]]:
        """NodeAllocator.Assume, node.go:61-73."""
        k = self.key(req)
        opt = self.allocated.get(k)
        if opt is not None:
            return opt.allocated
        opt = trade(self.gpus, RATERS[self.policy], req)
        if opt is None:
            return None
        self.allocated[k] = opt
        return opt.allocated

    def score(self, req: List[Unit]) -> int:
        """NodeAllocator.Score, node.go:75-85.  The reference dereferences a nil
        option (panic) when the entry was missing but Assume then succeeds;
        the driver order filter->priorities never reaches that, so raise."""
        k = self.key(req)
        opt = self.allocated.get(k)
        if opt is None:
            ids = self.assume(req)
            if not ids:
                return SCORE_MIN
            raise RuntimeError("reference would panic: nil option (node.go:84)")
        return opt.score

    def add(self, uid: int, option: GPUOption) -> bool:
        """NodeAllocator.Add, node.go:148-160 (option already built)."""
        if uid not in self.pods_map:
            self.pods_map[uid] = True
            return transact(self.gpus, option)
        return True

    def allocate(self, req: List[Unit], uid: int):
        """NodeAllocator.Allocate, node.go:87-104 -> (status, allocated)."""
        k = self.key(req)
        opt = self.allocated.get(k)
        try:
            if opt is None:
                return EGS_ERR_NO_OPTION, None
            if not self.add(uid, opt):
                return EGS_ERR_TRANSACT, None
            return EGS_OK, opt.allocated
        finally:
            self.allocated.pop(k, None)                               # node.go:90-92

    def forget(self, req: List[Unit], alloc: List[Optional[List[int]]], uid: int) -> None:
        """NodeAllocator.Forget, node.go:129-140 (+NewGPUOptionFromPod allocate.go:75-93)."""
        if uid in self.pods_map:
            cancel(self.gpus, GPUOption(request=list(req), allocated=list(alloc)))
            del self.pods_map[uid]


MASK64 = (1 << 64) - 1


def splitmix64_next(state: List[int]) -> int:
    """splitmix64 (SURVEY.md 8d 'Deterministic inputs')."""
    state[0] = (state[0] + 0x9E3779B97F4A7C15) & MASK64
    z = state[0]
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def mix64(x: int) -> int:
    """Digest mixer = one splitmix64 output for state x (see include/egs.h)."""
    z = (x + 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def fit_digest_term(node: int) -> int:
    return mix64(2 * node + 1)


def score_digest_term(node: int, score: int) -> int:
    return (mix64(2 * node + 2) * (2 * (score & 0xFFFFFFFF) + 1)) & MASK64


class Scheduler:
    """GPUUnitScheduler over pre-loaded nodes (scheduler.go:41-60, :108-290);
    node ids are dense ints (name interning stays in the Go host)."""

    def __init__(self, policy: int, faithful_hash: bool = False):
        self.policy = policy
        self.faithful_hash = faithful_hash
        self.nodes: List[Optional[NodeAllocator]] = []
        self.pod_maps: Dict[int, bool] = {}
        self.released: Dict[int, bool] = {}

    def add_node(self, core_allocatable: int, mem_allocatable: int) -> int:
        na = NodeAllocator.new(core_allocatable, mem_allocatable, self.policy, self.faithful_hash)
        self.nodes.append(na)
        return len(self.nodes) - 1 if na is not None else -1

    def set_rows(self, node: int, core: Sequence[int], mem: Sequence[int]) -> None:
        for g, c, m in zip(self.nodes[node].gpus, core, mem):
            g.core_avail, g.mem_avail = c, m

    def rows(self, node: int):
        return [(g.core_avail, g.mem_avail) for g in self.nodes[node].gpus]

    def assume(self, node_ids: Sequence[int], req: List[Unit]) -> List[int]:
        """GPUUnitScheduler.Assume, scheduler.go:112-168 -> fit flag per candidate
        (the 4 goroutines touch disjoint nodes, so serial order is equivalent)."""
        return [1 if (self.nodes[n] is not None and self.nodes[n].assume(req) is not None) else 0
                for n in node_ids]

    def score(self, node_ids: Sequence[int], req: List[Unit]) -> List[int]:
        """GPUUnitScheduler.Score, scheduler.go:170-184."""
        return [self.nodes[n].score(req) if self.nodes[n] is not None else SCORE_MIN
                for n in node_ids]

    def bind(self, node: int, req: List[Unit], uid: int):
        """GPUUnitScheduler.Bind, scheduler.go:186-227 (apiserver calls succeed)."""
        st, alloc = self.nodes[node].allocate(req, uid)
        if st == EGS_OK:
            self.pod_maps[uid] = True                                 # scheduler.go:224
        return st, alloc

    def add_pod(self, node: int, req: List[Unit], alloc, uid: int) -> None:
        """AddPod, scheduler.go:229-245 (ni.Add's error is discarded at :242)."""
        if uid in self.pod_maps:
            return
        self.nodes[node].add(uid, GPUOption(request=list(req), allocated=list(alloc)))
        self.pod_maps[uid] = True

    def forget_pod(self, node: int, req: List[Unit], alloc, uid: int) -> None:
        """ForgetPod, scheduler.go:247-267 (node < 0 == empty NodeName)."""
        if node >= 0:
            self.nodes[node].forget(req, alloc, uid)
        if uid in self.pod_maps:
            del self.pod_maps[uid]
            self.released[uid] = True

    def schedule_one(self, req: List[Unit], uid: int):
        """Harness driver rule: filter all nodes in index order ->
        score the fit nodes -> first max -> bind."""
        n = len(self.nodes)
        ids = list(range(n))
        fit = self.assume(ids, req)
        fit_ids = [i for i in ids if fit[i]]
        scores = self.score(fit_ids, req)
        fd = sd = 0
        for i, s in zip(fit_ids, scores):
            fd = (fd + fit_digest_term(i)) & MASK64
            sd = (sd + score_digest_term(i, s)) & MASK64
        res = dict(fit=fit, scores=scores, fit_ids=fit_ids, fit_count=len(fit_ids),
                   fit_digest=fd, score_digest=sd, node=-1, status=EGS_ERR_NOFIT, alloc=None)
        if not fit_ids:
            return res
        best = max(scores)
        w = fit_ids[scores.index(best)]
        st, alloc = self.bind(w, req, uid)
        res.update(node=w, status=st, alloc=alloc)
        return res
