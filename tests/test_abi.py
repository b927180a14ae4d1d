"""CPU checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol declared in
include/get_b200.h, the ctypes signature table covers all of them, and the host-side module mirror has the
reference's parameter names and shapes. No kernel is launched here."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from helpers import ROOT, load_golden


def _declared():
    src = open(os.path.join(ROOT, "include", "get_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^(?:int|int64_t|const char\*)\s+(get_\w+)\s*\(", src, flags=re.M)))


def test_library_builds_and_exports_every_declared_symbol():
    from get_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    declared = _declared()
    assert len(declared) >= 20
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared in include/get_b200.h but not exported: %s" % missing
    assert sorted(_lib.SIGNATURES.keys()) == declared, "ctypes table and header disagree"
    lib = _lib.load()                      # dlopen + symbol lookup only
    assert lib.get_b200_abi_version() == 2
    assert lib.get_b200_launch_count() == 0


def test_library_is_sm100a_only_and_contains_the_kernels():
    from get_b200 import build
    path = build.build()
    out = subprocess.run(["cuobjdump", "-lelf", path], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_descriptor_layouts_match_header():
    """sizeof/offsets of the ctypes mirrors vs the C structs (compiled with gcc)."""
    import ctypes
    import tempfile
    from get_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "get_b200.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(get_gemm_desc), sizeof(get_gemm_operand),
    offsetof(get_gemm_desc,B), offsetof(get_gemm_desc,K), offsetof(get_gemm_desc,C), offsetof(get_gemm_desc,bias0),
    offsetof(get_gemm_desc,group_rows), offsetof(get_gemm_desc,drop_out_p), offsetof(get_gemm_desc,workspace));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(get_gemm_bp_desc), sizeof(get_bp_tensor),
    offsetof(get_gemm_bp_desc,B), offsetof(get_gemm_bp_desc,K), offsetof(get_gemm_bp_desc,mode), offsetof(get_gemm_bp_desc,C),
    offsetof(get_gemm_bp_desc,bias), offsetof(get_gemm_bp_desc,planes_out), offsetof(get_gemm_bp_desc,group_rows),
    offsetof(get_gemm_bp_desc,drop_out_p), offsetof(get_gemm_bp_desc,split_k), offsetof(get_gemm_bp_desc,workspace_floats));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(get_bp_dst), offsetof(get_bp_dst,row0), sizeof(get_pack_job),
    offsetof(get_pack_job,dst), offsetof(get_pack_job,first_block), offsetof(get_pack_job,kind));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    D, B, T, J = _lib.GemmDesc, _lib.GemmBpDesc, _lib.BpDst, _lib.PackJob
    want = [ctypes.sizeof(D), ctypes.sizeof(_lib.GemmOperand), D.B.offset, D.K.offset, D.C.offset, D.bias0.offset,
            D.group_rows.offset, D.drop_out_p.offset, D.workspace.offset,
            ctypes.sizeof(B), ctypes.sizeof(_lib.BpTensor), B.B.offset, B.K.offset, B.mode.offset, B.C.offset, B.bias.offset,
            B.planes_out.offset, B.group_rows.offset, B.drop_out_p.offset, B.split_k.offset, B.workspace_floats.offset,
            ctypes.sizeof(T), T.row0.offset, ctypes.sizeof(J), J.dst.offset, J.first_block.offset, J.kind.offset]
    assert got == want


@pytest.mark.parametrize("name,over", [
    ("tiny_snopes", {}),
    ("tiny_politifact", dict(use_claim_source=True, heads_words=2, heads_evds=1, gsl_rate=0.3)),
    ("snopes_dims", None),
])
def test_state_dict_keys_and_shapes_match_reference(name, over):
    from get_b200 import synthetic
    from get_b200.model import Graph_basedSemantiStructure
    gold = load_golden(name)
    if over is None:
        w = synthetic.get_workload("snopes", vocab=400, n_article_sources=16)
    else:
        w = synthetic.get_workload("tiny", **over)
    model = Graph_basedSemantiStructure(synthetic.match_params(w))
    ref = dict(zip((str(s) for s in gold["out/param_names"]), (str(s) for s in gold["out/param_shapes"])))
    got = {k: ",".join(str(int(s)) for s in v.shape) for k, v in model.state_dict().items()}
    assert got == ref
    # parameters that never receive a gradient in the reference exist but are inert here too
    assert all(any(n.startswith(p) for p in ("bilstm.", "query_bilstm.", "trans.", "ggnn_with_gsl.word_scorer1."))
               for n in (str(s) for s in gold["out/no_grad_params"]))


def test_product_path_has_no_cpu_fallback_and_never_imports_the_oracle():
    import get_b200.model, get_b200.modules, get_b200.ops  # noqa: F401
    from get_b200.modules import GGNN
    with pytest.raises(RuntimeError):
        GGNN(4, 4)(torch.zeros(1, 3, 3), torch.zeros(1, 3, 4))
    pkg = os.path.join(ROOT, "get_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_dropout_host_mirror_statistics():
    from get_b200.dropout import keep_mask
    m = keep_mask(200000, 0.2, 42)
    assert set(np.unique(m.numpy()).tolist()) == {0.0, 1.25}
    assert abs(float((m == 0).float().mean()) - 0.2) < 0.005
    assert not torch.equal(m, keep_mask(200000, 0.2, 43))
    assert torch.equal(keep_mask(10, 0.0, 1), torch.ones(10))
