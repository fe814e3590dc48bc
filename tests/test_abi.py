"""CPU-side checks of the drop-in boundary: libegs builds for sm_100a, loads, and exports every
symbol include/egs.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(egs):
    L = egs.load()
    hdr = open(os.path.join(ROOT, "include", "egs.h")).read()
    declared = sorted(set(re.findall(r"\b(egs_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    assert sorted(egs.SYMBOLS) == declared, "capi.SYMBOLS out of step with include/egs.h"
    for name in declared:
        assert hasattr(L, name), f"libegs.so does not export {name}"


def test_status_strings_match_reference(egs):
    L = egs.load()
    assert L.egs_status_string(egs.EGS_ERR_NOFIT) == b"no enough resource to allocate"      # gpu.go:126
    assert L.egs_status_string(egs.EGS_ERR_NO_OPTION).startswith(b"cannot find option of GPU request")
    assert L.egs_status_string(egs.EGS_ERR_NO_GPU).startswith(b"no gpu available on node")
    assert L.egs_status_string(egs.EGS_ERR_NO_NODE).startswith(b"elastic gpu scheduler get node failed")


def test_mix64_matches_oracle(egs):
    import egs_oracle as po
    import oracle_c as oc
    L = egs.load()
    for x in [0, 1, 2**32, 2**64 - 1, 0xDEADBEEF12345678]:
        assert L.egs_mix64(x) == po.mix64(x) == oc.lib().egso_mix64(x)


def test_unit_from_requests(egs):
    from ka_vectors import UNIT_KA
    for inp, unit in UNIT_KA:
        assert egs.unit_from_requests(*inp) == unit


def test_sass_is_sm100a(egs):
    """The shipped library holds sm_100a SASS (no PTX-JIT, no other arch)."""
    import shutil
    import subprocess
    path = egs.lib_path()
    if not shutil.which("cuobjdump"):
        return
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_create_without_gpu_fails_loudly(egs):
    """No CPU fallback: on a box without a CUDA device egs_create must error, not emulate."""
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    assert egs.load().egs_create(0, 16, 8, 0, ctypes.byref(h)) == egs.EGS_ERR_CUDA


def test_workload_generators():
    import egs_b200
    w = egs_b200.workloads.config(0)
    assert w.n_nodes == 4 and w.gpus == 2 and w.mem_total == 16 and w.n_pods == 8
    assert list(w.units[:, 1]) == [4, 8, 4, 12, 16, 8, 4, 8]
    w = egs_b200.workloads.config(4, n_nodes=1000, n_pods=500)
    assert w.core.shape == (1000, 8) and ((w.core >= 0) & (w.core <= 100)).all()
    assert ((w.mem >= 81920 - 80 * 1024) & (w.mem <= 81920)).all()
    full = (w.core == 100) & (w.mem == 81920)
    assert 0.45 < full.mean() < 0.6
    assert len(np.unique(w.units, axis=0)) == 16
    # prefix stability
    w2 = egs_b200.workloads.config(4, n_nodes=100, n_pods=50)
    assert (w2.core == w.core[:100]).all() and (w2.units == w.units[:50]).all()


def test_mutation_record_layout_matches_the_c_struct(egs, tmp_path):
    """capi.MUTATION_DTYPE (numpy) must be egs_mutation (include/egs.h) byte for byte: size and every field offset,
    taken from a C program compiled against the header."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "egs.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %d %d\\n", sizeof(egs_mutation),'
                   'offsetof(egs_mutation,kind),offsetof(egs_mutation,node_id),offsetof(egs_mutation,n_containers),'
                   'offsetof(egs_mutation,units),offsetof(egs_mutation,n_idx),offsetof(egs_mutation,idx),'
                   'offsetof(egs_mutation,uid),sizeof(egs_unit),EGS_MAX_CONTAINERS_APPLY,EGS_MAX_GPUS);return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    v = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    dt = egs.MUTATION_DTYPE
    assert v[0] == dt.itemsize
    for off, name in zip(v[1:8], ["kind", "node_id", "n_containers", "units", "n_idx", "idx", "uid"]):
        assert dt.fields[name][1] == off, name
    assert v[8] == 12 and v[9] == egs.EGS_MAX_CONTAINERS_APPLY == 8 and v[10] == 8
    a = egs.mutations_array([(egs.EGS_MUT_ADD, 7, [(10, 4096, 0), (-1, -1, 0)], [[3], []], 99)])
    assert a[0]["node_id"] == 7 and a[0]["n_containers"] == 2 and a[0]["uid"] == 99
    assert a[0]["units"][1].tolist() == [-1, -1, 0] and a[0]["n_idx"][:2].tolist() == [1, 0] and a[0]["idx"][0][0] == 3
