"""Import the UNMODIFIED reference (read-only at /root/reference) inside THIS container only.

Test infrastructure for generating golden vectors (SURVEY.md App. B recipe). Never imported by
`-m gpu` tests, smoke() or bench.py: /root/reference does not exist on the GPU box.

Stubs the reference's missing third-party imports (nltk, hyperopt, keras, allennlp, tensorboardX,
pytorch_transformers) and, on CPU, shims `Tensor.cuda` to identity because
Models/BiDAF/wrapper.py:221 hard-codes `.cuda()`.
"""
import os
import sys
import tempfile
import types

REF_ROOT = os.environ.get("GET_REFERENCE_ROOT", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (gbss_module, wrapper_module, two_branches_attention_module, self_attention_module)."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree %s not present (only available in the build container)" % REF_ROOT)
    os.environ["HOME"] = tempfile.mkdtemp(prefix="get_ref_home_")
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
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    from Models.FCWithEvidences import graph_based_semantic_structure as gbss
    from Models.BiDAF import wrapper
    from thirdparty import two_branches_attention as tba
    from thirdparty import self_attention as sa
    return gbss, wrapper, tba, sa
