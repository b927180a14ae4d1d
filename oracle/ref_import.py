"""TEST / BASELINE INFRASTRUCTURE ONLY (never imported by get_b200).

Import the UNMODIFIED reference modules of the GET hot path:
  * in the build container from the read-only tree /root/reference;
  * on the GPU box from `oracle/_ref/`, a git-ignored verbatim copy of exactly the reference files this import needs,
    made by `python oracle/make_ref.py` (no reference source is ever committed to this repository).

Stubs the reference's missing third-party imports (nltk, hyperopt, keras, allennlp, tensorboardX, pytorch_transformers;
SURVEY.md Appendix B) and, when no GPU is present, shims `Tensor.cuda` to identity because Models/BiDAF/wrapper.py:221
hard-codes `.cuda()`. Used by tests/golden/make_golden.py (golden fixtures) and by bench.py's `--impl reference` and
`torch_cuda_baseline` legs (the reference's own modules timed on the host cores / on the GPU as the baseline)."""
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
LOCAL_COPY = os.path.join(HERE, "_ref")


def reference_root():
    env = os.environ.get("GET_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/Models"):
        return "/root/reference"
    if os.path.isdir(os.path.join(LOCAL_COPY, "Models")):
        return LOCAL_COPY
    return None


def available() -> bool:
    return reference_root() is not None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (gbss_module, wrapper_module, two_branches_attention_module, self_attention_module)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference modules not available: neither /root/reference nor oracle/_ref exists "
                           "(run `python oracle/make_ref.py` in the build container)")
    os.environ["HOME"] = tempfile.mkdtemp(prefix="get_ref_home_")     # `import matchzoo` creates ~/.matchzoo
    sys.dont_write_bytecode = True
    import torch

    class _Apply(object):
        pass

    _stub("nltk")
    hp = _stub("hyperopt", hp=types.SimpleNamespace())
    pyll = _stub("hyperopt.pyll", Apply=_Apply)
    base = _stub("hyperopt.pyll.base", Apply=_Apply)
    hp.pyll = pyll
    pyll.base = base
    _stub("keras")
    al = _stub("allennlp")
    alm = _stub("allennlp.modules")
    ale = _stub("allennlp.modules.elmo", batch_to_ids=lambda *a, **k: None, Elmo=object)
    al.modules = alm
    alm.elmo = ale
    _stub("tensorboardX", SummaryWriter=object)
    _stub("pytorch_transformers", BertModel=object)
    if root not in sys.path:
        sys.path.insert(0, root)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    from Models.FCWithEvidences import graph_based_semantic_structure as gbss
    from Models.BiDAF import wrapper
    from thirdparty import two_branches_attention as tba
    from thirdparty import self_attention as sa
    return gbss, wrapper, tba, sa
